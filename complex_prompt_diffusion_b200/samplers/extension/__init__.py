from .registry import register, make, create, lookup
from .denoiser import Denoiser

__all__ = ["register", "make", "create", "lookup", "Denoiser"]
