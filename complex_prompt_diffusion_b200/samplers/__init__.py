"""Sampler plugin package: importing it registers "Euler", "Euler Ancestral" and "DPM++ 2m"
(interface of cpd/samplers/__init__.py)."""
from .registry import register, make, create, lookup
from . import euler, dpmpp  # noqa: F401  (registration side effect)

__all__ = ["register", "make", "create", "lookup"]
