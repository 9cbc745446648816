"""Tetrahedral mesh container and synthetic mesh generators.

``pyvista`` is not available in this environment, so ``TetMesh`` provides the small part of the
``pv.UnstructuredGrid`` surface the reference touches (``points``, ``cells_dict``-like ``cells``,
``point_data``, ``cell_data``, ``n_points``, ``n_cells``).  Every ``from_pyvista`` constructor in this
package accepts either a ``TetMesh`` or a real ``pyvista.UnstructuredGrid``.  ``read_vtu`` / ``write_vtu`` (and
``TetMesh.load`` / ``TetMesh.save``) read and write VTK XML ``.vtu`` files without VTK.
"""

from ._mesh import TetMesh, as_tet_arrays
from ._generate import (cube_tet_mesh, cube_tet_slab, embedded_tetra_mesh, hash_uniform, lumped_vertex_volume,
                        morton_reorder)
from ._vtu import read_vtu, write_vtu
from ._device import (DeviceMesh, cube_tet_mesh_device, cube_tet_slab_device, hash_uniform_device,
                      morton_codes_device)

__all__ = [
    "DeviceMesh",
    "TetMesh",
    "cube_tet_mesh_device",
    "cube_tet_slab_device",
    "hash_uniform_device",
    "morton_codes_device",
    "as_tet_arrays",
    "cube_tet_mesh",
    "cube_tet_slab",
    "hash_uniform",
    "embedded_tetra_mesh",
    "lumped_vertex_volume",
    "morton_reorder",
    "read_vtu",
    "write_vtu",
]
