"""SigmaScheduler - the object `Denoiser.scheduler` exposes (reference: cpd/scheduler/discrete.py:12-137).

The reference class is unusable as committed (no `append_zero`, `sigmas` is None and is overwritten by every
`get_sigmas` call - SURVEY.md 8-c D1/D2); this mirror keeps its public surface (`get_sigmas`, `get_scalings`,
`sigma_to_t`, `t_to_sigma`, `.sigmas`) on top of KScheduler's separate training table, which is the only
working combination in the reference (k.py:98,268-279).
"""
from .k import KScheduler


class SigmaScheduler(KScheduler):
    def __init__(self, **kwargs):
        super().__init__(kwargs.get("num_train_timesteps", 1000))
        self.algorithm = kwargs.get("sigma_algorithm", "default")
        self.total_steps = kwargs.get("steps", kwargs.get("total_steps", None))
        self.schedule = self.get_sigmas(self.algorithm, self.total_steps, **kwargs) if self.total_steps else None

    def get_scalings(self, x, convert_to_sigma=False):
        sigma = self.t_to_sigma(x) if convert_to_sigma else x
        return KScheduler.get_scalings(sigma)
