"""Oracle (test infrastructure): sigma schedules and sigma<->t maps.

Follows cpd/scheduler/k.py:30-116,157-208,216-279,556-576 and cpd/scheduler/discrete.py:21-137
(paths relative to /root/reference).  See oracle/__init__.py for the defect repairs (D1, D2).
"""
import math

import numpy as np
import torch


def make_betas(n_timestep=1000, linear_start=0.0008, linear_end=0.012, max_beta=0.999, decimals=4):
    """'quad' beta schedule, clamped, ROUNDED to 4 decimals.  k.py:40-42,166-169,205-208."""
    betas = np.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=np.float64) ** 2
    betas = torch.from_numpy(betas).clamp(max=max_beta)
    return np.around(betas, decimals=decimals)  # -> torch.float64 tensor (np.around dispatches to .round)


def training_tables(n_timestep=1000):
    """betas, alphas_cumprod and the 1000-entry training sigma table (fp64).  k.py:53-63,98-99."""
    betas = make_betas(n_timestep)
    alphas = 1.0 - betas.numpy()
    alphas_cumprod = np.cumprod(alphas, axis=0)
    sigmas = ((1 - alphas_cumprod) / alphas_cumprod) ** 0.5
    return betas, torch.from_numpy(alphas_cumprod), torch.from_numpy(sigmas)


def append_zero(x):
    """k.py:575-576."""
    return torch.cat([x, x.new_zeros([1])])


class OracleSchedule:
    """KScheduler table semantics (D2) + SigmaScheduler algorithm set (discrete.py:87-108)."""

    def __init__(self, num_train_timesteps=1000):
        self.betas, self.alphas_cumprod, self.sigmas = training_tables(num_train_timesteps)

    # --- inference sigma schedules -------------------------------------------------------------
    def get_sigmas_karras(self, n, sigma_min=0.1, sigma_max=10, rho=7.0, **_):
        """k.py:216-226 / discrete.py:21-32 (fp32 ramp)."""
        ramp = torch.linspace(0, 1, n)
        min_inv_rho = sigma_min ** (1 / rho)
        max_inv_rho = sigma_max ** (1 / rho)
        return (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho

    def get_sigmas_exponential(self, n, sigma_min=0.1, sigma_max=10, **_):
        """k.py:228-237."""
        return torch.linspace(math.log(sigma_max), math.log(sigma_min), n).exp()

    def get_sigmas_quad(self, n, sigma_min=0.1, sigma_max=10, **_):
        """k.py:239-248."""
        return torch.linspace(math.sqrt(sigma_max), math.sqrt(sigma_min), n) ** 2

    def get_sigmas_sigmoid(self, n, sigma_min=0.1, sigma_max=10.0, **_):
        """discrete.py:56-64 (multiplies by sigma_min; reproduced as written)."""
        return torch.sigmoid(torch.linspace(-6, 6, n)) * (sigma_max - sigma_min) * sigma_min

    def get_sigmas_vp(self, n, beta_d=19.9, beta_min=0.1, eps_s=1e-3, **_):
        """k.py:250-258."""
        t = torch.linspace(1, eps_s, n)
        return torch.sqrt(torch.exp(beta_d * t ** 2 / 2 + beta_min * t) - 1)

    def get_sigmas_linear(self, n, **_):
        """k.py:260-266: t_to_sigma(linspace(999, 0, n)) -> fp64."""
        if n is None:
            return append_zero(self.sigmas.flip(0))
        t_max = len(self.sigmas) - 1
        return self.t_to_sigma(torch.linspace(t_max, 0, n))

    def get_sigmas(self, algorithm, n, **kw):
        """discrete.py:87-108 / k.py:268-279.  Returns n+1 values (zero appended)."""
        kw = {k: v for k, v in kw.items() if k in ("sigma_min", "sigma_max", "rho", "beta_d", "beta_min", "eps_s")}
        if algorithm in ("linear", "default"):
            s = self.get_sigmas_linear(n, **kw)
        elif algorithm in ("karras",):
            s = self.get_sigmas_karras(n, **kw)
        elif algorithm in ("exp", "exponential"):
            s = self.get_sigmas_exponential(n, **kw)
        elif algorithm in ("quad", "quadratic"):
            s = self.get_sigmas_quad(n, **kw)
        elif algorithm in ("vp", "variance_preserving"):
            s = self.get_sigmas_vp(n, **kw)
        elif algorithm in ("sig", "sigmoid"):
            s = self.get_sigmas_sigmoid(n, **kw)
        else:
            raise NotImplementedError(algorithm)
        return append_zero(s)

    # --- sigma <-> t ---------------------------------------------------------------------------
    def sigma_to_t_idx(self, sigma):
        """k.py:556-567.  Returns (t fp64, low_idx int64, high_idx int64)."""
        sigma = sigma.cpu()
        dists = torch.abs(sigma - self.sigmas[:, None])
        low_idx, high_idx = torch.sort(torch.topk(dists, dim=0, k=2, largest=False).indices, dim=0)[0]
        low, high = self.sigmas[low_idx], self.sigmas[high_idx]
        w = (low - sigma) / (low - high)
        w = w.clamp(0, 1)
        t = (1 - w) * low_idx + w * high_idx
        return t.view(sigma.shape), low_idx, high_idx

    def sigma_to_t(self, sigma):
        return self.sigma_to_t_idx(sigma)[0]

    def t_to_sigma(self, t):
        """k.py:569-573."""
        t = t.cpu().float()
        low_idx, high_idx, w = t.floor().long(), t.ceil().long(), t.frac()
        return (1 - w) * self.sigmas[low_idx] + w * self.sigmas[high_idx]

    @staticmethod
    def get_scalings(sigma):
        """discrete.py:110-117: c_out = -sigma, c_in = 1/sqrt(sigma^2 + 1)."""
        c_out = -sigma
        c_in = 1 / (sigma ** 2 + 1 ** 2) ** 0.5
        return c_out, c_in
