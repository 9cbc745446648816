"""Rest-shape precompute (mirror of ``liblaf.apple.jax.fem``: ``Region``)."""

from ._region import Region

__all__ = ["Region"]
