from ._ext_force import ExternalForce

__all__ = ["ExternalForce"]
