"""CPU oracle: a numpy restatement of liblaf/apple's FEM-elasticity hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``apple_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` do, and there only as the checker / reported CPU baseline.

Every function cites the reference file:line (relative to ``/root/reference/src/liblaf/apple``)
that it restates.  The reference itself cannot be imported here (``warp``, ``jax``, ``pyvista``,
``liblaf.*`` are absent), so the oracle is pinned by

* the reference's one known-answer test (``tests/forward/test_static_simulation.py:16-93``),
* central finite differences of its own energy (gradient, Hessian product, Hessian diagonal,
  Hessian quadratic form), and
* the derived golden energies recorded in SURVEY.md section 8(c).

PNCG parity is UNPINNED at the iteration level: the optimizer lives in the un-vendored
``liblaf-peach`` dependency.  ``oracle.pncg`` restates the recurrences of the reference's own
"PNCG-like" benchmark (``benches/bench_pncg_branching_backends.py``) and is validated at the
convergence level only.
"""

from . import fem, pncg, region  # noqa: F401
