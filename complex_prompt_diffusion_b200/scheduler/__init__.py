from .k import KScheduler
from .discrete import SigmaScheduler

__all__ = ["KScheduler", "SigmaScheduler"]
