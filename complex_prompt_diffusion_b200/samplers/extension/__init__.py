from .registry import register, make, create, lookup
from .denoiser import Denoiser
from . import threshold  # noqa: F401  (registers the thresholding / score-corrector extensions)

__all__ = ["register", "make", "create", "lookup", "Denoiser", "threshold"]
