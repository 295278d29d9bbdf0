"""Euler and Euler-ancestral samplers (interface of cpd/samplers/euler.py:13-95).

Step scalars follow euler.py:49-54 (Euler), :84-92 and get_ancestral_step :97-102 (ancestral), evaluated on
0-dim fp32 torch tensors on the host; the update itself is part of the fused CUDA step kernel.
"""
import torch

from .._lib import CPD_EULER, CPD_EULER_ANCESTRAL
from .diffusion import DiffusionSamplerWrapper
from .k_diffusion import KDiffusionSampler
from .registry import register


def get_ancestral_step(sigma_from, sigma_to):
    """sigma_down / sigma_up of an ancestral step (Karras et al. 2022, as euler.py:97-102)."""
    sigma_up = (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5
    sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
    return sigma_down, sigma_up


class EulerDiffusionSampler(KDiffusionSampler):
    """Algorithm 2 (Euler steps) of Karras et al. (2022), gamma = 0."""

    def __init__(self, model):
        super().__init__(model, "euler")

    @torch.no_grad()
    def _sampling(self, x, sigmas, model_args=None, **kwargs):
        model_args = {} if model_args is None else model_args
        callback = kwargs.get("callback", None)
        den, plan = self._begin(x, model_args, kwargs)
        if self._step_graph_ok(plan, model_args, kwargs) and not kwargs.get("s_churn", 0.0) and not kwargs.get("rng_compat", True):
            # gamma = 0 on every step and no RNG draw to reproduce: one captured graph per step (k_diffusion._graph_loop)
            rows = [dict(dt=float(sigmas[i + 1] - sigmas[i] * 1)) for i in range(len(sigmas) - 1)]
            return self._graph_loop(x, sigmas, plan, CPD_EULER, rows, model_args)
        den_out = torch.empty_like(x) if callback is not None else None
        for i in range(len(sigmas) - 1):
            model_args["t_idx"] = i
            sigma_hat = self._churn(x, sigmas, i, kwargs)  # euler.py:42-46
            dt = sigmas[i + 1] - sigma_hat
            x_before = x.clone() if callback is not None else None
            den.fused_step(x, sigma_hat, plan, dict(sampler=CPD_EULER, dt=float(dt), denoised_out=den_out), **model_args)
            self._clip_sample(x, kwargs)
            self._callback(callback, x_before, i, sigmas[i], den_out, sigma_hat)
        return x


class EulerAncestralDiffusionSampler(KDiffusionSampler):
    """Ancestral sampling with Euler steps."""

    def __init__(self, model):
        super().__init__(model, "euler ancestral")

    @torch.no_grad()
    def _sampling(self, x, sigmas, model_args=None, **kwargs):
        model_args = {} if model_args is None else model_args
        callback = kwargs.get("callback", None)
        noise_sampler = kwargs.get("noise_sampler", None)  # default reproduces torch.randn_like(x) call order
        den, plan = self._begin(x, model_args, kwargs)
        if self._step_graph_ok(plan, model_args, kwargs):
            rows = []
            for i in range(len(sigmas) - 1):
                sigma_down, sigma_up = get_ancestral_step(sigmas[i], sigmas[i + 1])
                rows.append(dict(dt=float(sigma_down - sigmas[i]), sigma_up=float(sigma_up)))
            draw = noise_sampler if noise_sampler is not None else torch.randn_like
            return self._graph_loop(x, sigmas, plan, CPD_EULER_ANCESTRAL, rows, model_args, noise_fn=lambda t: draw(t).contiguous())
        den_out = torch.empty_like(x) if callback is not None else None
        for i in range(len(sigmas) - 1):
            model_args["t_idx"] = i
            sigma_down, sigma_up = get_ancestral_step(sigmas[i], sigmas[i + 1])
            dt = sigma_down - sigmas[i]
            noise = noise_sampler(x) if noise_sampler is not None else torch.randn_like(x)
            noise = noise.to(x.device, torch.float32).contiguous()
            x_before = x.clone() if callback is not None else None
            den.fused_step(x, sigmas[i], plan, dict(sampler=CPD_EULER_ANCESTRAL, dt=float(dt), sigma_up=float(sigma_up),
                                                    noise=noise, denoised_out=den_out), **model_args)
            self._clip_sample(x, kwargs)
            self._callback(callback, x_before, i, sigmas[i], den_out)
        return x


@register("Euler")
class EulerSamplerWrapper(DiffusionSamplerWrapper):
    def __init__(self, name, **kwargs):
        kwargs["constructor"] = EulerDiffusionSampler
        super().__init__(name, **kwargs)


@register("Euler Ancestral")
class EulerAncestralSamplerWrapper(DiffusionSamplerWrapper):
    def __init__(self, name, **kwargs):
        kwargs["constructor"] = EulerAncestralDiffusionSampler
        super().__init__(name, **kwargs)
