"""Sampler plugin package: importing it registers "Euler", "Euler Ancestral", "DPM++ 2m" and the two-stage / multistep
samplers "Huen", "DPM2", "DPM2 Ancestral", "DPM++ 2s Ancestral", "LMS" (interface of cpd/samplers/__init__.py)."""
from .registry import register, make, create, lookup
from . import euler, dpmpp, multistage  # noqa: F401  (registration side effect)

__all__ = ["register", "make", "create", "lookup"]
