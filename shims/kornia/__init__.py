"""Minimal `kornia`: the one function santurini/vsrlab imports (`kornia.geometry.transform.resize`)."""
from . import geometry  # noqa: F401
