"""KScheduler: the 1000-entry training-sigma table plus the sigma schedules of the k-diffusion loop.

Mirror of the reference interface `cpd/scheduler/k.py` (KScheduler, :30-116 tables, :216-279 schedules,
:556-576 sigma<->t), re-implemented: tables are built once in numpy fp64, schedules with the same torch
fp32 expressions so their bit patterns match the reference's, and `sigma_to_t` does a binary search on the
(monotone) table instead of a 1000-way |sigma - table| top-k, returning the same two neighbour indices.
Host-side only (a few hundred flops per sample() call); nothing here touches the GPU.
"""
import math

import numpy as np
import torch


def _beta_table(n, start, end, max_beta, decimals):
    # "quad"/"scaled_linear": linspace in sqrt space, squared; clamp; ROUND to `decimals` (k.py:166-169,207-208)
    b = np.linspace(start ** 0.5, end ** 0.5, n, dtype=np.float64) ** 2
    b = torch.from_numpy(b).clamp(max=max_beta)
    return torch.round(b, decimals=decimals)


class KScheduler:
    def __init__(self, num_train_timesteps: int = 1000, **kwargs):
        schedule = kwargs.get("beta_schedule", "quad")
        if schedule not in ("quad", "scaled_linear"):
            raise NotImplementedError(f"beta_schedule {schedule!r}: only the default 'quad' table is on the hot path")
        self.num_train_timesteps = num_train_timesteps
        self.betas = _beta_table(num_train_timesteps, kwargs.get("beta_start", 0.0008), kwargs.get("beta_end", 0.012),
                                 kwargs.get("beta_max", 0.999), 4)
        self.alphas = 1.0 - self.betas.numpy()
        self.alphas_cumprod = np.cumprod(self.alphas, axis=0)
        assert self.alphas_cumprod.shape[0] == num_train_timesteps
        self.sigmas = torch.from_numpy(((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5)  # fp64, increasing
        self._table = self.sigmas.numpy()
        self.quantize = kwargs.get("quantize", False)

    # ---- schedules (n values + appended zero) ----------------------------------------------------
    @staticmethod
    def append_zero(x):
        return torch.cat([x, x.new_zeros([1])])

    def get_sigmas_karras(self, n, **kwargs):
        lo, hi, rho = kwargs.get("sigma_min", 0.1), kwargs.get("sigma_max", 10), kwargs.get("rho", 7.0)
        ramp = torch.linspace(0, 1, n)
        a, b = hi ** (1 / rho), lo ** (1 / rho)
        return (a + ramp * (b - a)) ** rho

    def get_sigmas_exponential(self, n, **kwargs):
        lo, hi = kwargs.get("sigma_min", 0.1), kwargs.get("sigma_max", 10)
        return torch.linspace(math.log(hi), math.log(lo), n).exp()

    def get_sigmas_quad(self, n, **kwargs):
        lo, hi = kwargs.get("sigma_min", 0.1), kwargs.get("sigma_max", 10)
        return torch.linspace(math.sqrt(hi), math.sqrt(lo), n) ** 2

    def get_sigmas_vp(self, n, **kwargs):
        beta_d, beta_min, eps_s = kwargs.get("beta_d", 19.9), kwargs.get("beta_min", 0.1), kwargs.get("eps_s", 1e-3)
        t = torch.linspace(1, eps_s, n)
        return torch.sqrt(torch.exp(beta_d * t ** 2 / 2 + beta_min * t) - 1)

    def get_sigmas_sigmoid(self, n, **kwargs):
        # discrete.py:56-64 multiplies by sigma_min (sic); kept so the schedule matches the reference's.
        lo, hi = kwargs.get("sigma_min", 0.1), kwargs.get("sigma_max", 10.0)
        return torch.sigmoid(torch.linspace(-6, 6, n)) * (hi - lo) * lo

    def get_sigmas_linear(self, n, **kwargs):
        if n is None:
            return self.append_zero(self.sigmas.flip(0))
        return self.t_to_sigma(torch.linspace(len(self.sigmas) - 1, 0, n))

    _ALGS = {"linear": "linear", "default": "linear", "karras": "karras", "exp": "exponential", "exponential": "exponential",
             "quad": "quad", "quadratic": "quad", "vp": "vp", "variance_preserving": "vp", "sig": "sigmoid", "sigmoid": "sigmoid"}

    def get_sigmas(self, algorithm, n, **kwargs):
        if algorithm not in self._ALGS:
            raise NotImplementedError(f"unknown sigma algorithm {algorithm!r}")
        sig = getattr(self, "get_sigmas_" + self._ALGS[algorithm])(n, **kwargs)
        return self.append_zero(sig)

    # ---- sigma <-> t ------------------------------------------------------------------------------
    def sigma_to_idx(self, sigma):
        """(low_idx, high_idx): the two table entries nearest to each sigma, ascending - the integer
        contract of k.py:556-562 (topk of |sigma - table| with k = 2, sorted)."""
        s = np.atleast_1d(np.asarray(torch.as_tensor(sigma).detach().cpu().double().numpy()))
        tab = self._table
        n = len(tab)
        j = np.searchsorted(tab, s)  # tab[j-1] < s <= tab[j]
        low = np.empty(s.shape, dtype=np.int64)
        for i, (sv, jj) in enumerate(zip(s, j)):
            cands = [c for c in (jj - 2, jj - 1, jj, jj + 1) if 0 <= c < n]
            # two smallest distances; ties resolved to the lower index like a stable ascending sort
            cands.sort(key=lambda c: (abs(sv - tab[c]), c))
            low[i] = min(cands[0], cands[1])
            assert abs(cands[0] - cands[1]) == 1
        return torch.from_numpy(low), torch.from_numpy(low + 1)

    def sigma_to_t(self, sigma, quantize=None, device=None):
        sigma = torch.as_tensor(sigma).detach().cpu()
        low_idx, high_idx = self.sigma_to_idx(sigma)
        low_idx, high_idx = low_idx.view(sigma.shape), high_idx.view(sigma.shape)
        low, high = self.sigmas[low_idx], self.sigmas[high_idx]
        w = ((low - sigma) / (low - high)).clamp(0, 1)
        t = (1 - w) * low_idx + w * high_idx
        return t.view(sigma.shape)

    def t_to_sigma(self, t, device=None):
        t = torch.as_tensor(t).detach().cpu().float()
        low_idx, high_idx, w = t.floor().long(), t.ceil().long(), t.frac()
        return (1 - w) * self.sigmas[low_idx] + w * self.sigmas[high_idx]

    @staticmethod
    def get_scalings(sigma=None, **kwargs):
        if sigma is None:
            sigma = kwargs["sigma"]
        return -sigma, 1 / (sigma ** 2 + 1 ** 2) ** 0.5
