"""DPM-Solver++(2M) sampler (interface of cpd/samplers/dpmpp.py:12-56).

The multistep coefficients follow dpmpp.py:42-50, evaluated on 0-dim fp32 tensors on the host; the history
tensor (previous denoised) lives in HBM and is read and rewritten by the fused step kernel.
"""
import torch

from .._lib import CPD_DPMPP_2M
from .diffusion import DiffusionSamplerWrapper
from .k_diffusion import KDiffusionSampler
from .registry import register


def dpmpp_2m_scalars(sigmas, i, have_history):
    """(ratio, expm1, c1, c2, first) for step i: x' = ratio * x - expm1 * (c1 * den - c2 * old)."""
    t, t_next = sigmas[i].log().neg(), sigmas[i + 1].log().neg()
    h = t_next - t
    ratio = t_next.neg().exp() / t.neg().exp()
    em = (-h).expm1()
    if not have_history or bool(sigmas[i + 1] == 0):
        return float(ratio), float(em), 0.0, 0.0, 1
    h_last = t - sigmas[i - 1].log().neg()
    r = h_last / h
    return float(ratio), float(em), float(1 + 1 / (2 * r)), float(1 / (2 * r)), 0


class DPMPlusPlus2mDiffusionSampler(KDiffusionSampler):
    def __init__(self, model):
        super().__init__(model, "dpmpp 2m")

    @torch.no_grad()
    def _sampling(self, x, sigmas, model_args=None, **kwargs):
        model_args = {} if model_args is None else model_args
        callback = kwargs.get("callback", None)
        den, plan = self._begin(x, model_args, kwargs)
        if self._step_graph_ok(plan, model_args, kwargs):
            rows = []
            for i in range(len(sigmas) - 1):
                ratio, em, c1, c2, first = dpmpp_2m_scalars(sigmas, i, have_history=i > 0)
                rows.append(dict(dpm_ratio=ratio, dpm_expm1=em, dpm_c1=c1, dpm_c2=c2, dpm_first=first, write_old=1))
            return self._graph_loop(x, sigmas, plan, CPD_DPMPP_2M, rows, model_args)
        old = torch.empty_like(x)
        den_out = torch.empty_like(x) if callback is not None else None
        for i in range(len(sigmas) - 1):
            model_args["t_idx"] = i
            ratio, em, c1, c2, first = dpmpp_2m_scalars(sigmas, i, have_history=i > 0)
            x_before = x.clone() if callback is not None else None
            den.fused_step(x, sigmas[i], plan, dict(sampler=CPD_DPMPP_2M, dpm_ratio=ratio, dpm_expm1=em, dpm_c1=c1, dpm_c2=c2,
                                                    dpm_first=first, write_old=1, old_denoised=old, denoised_out=den_out),
                           **model_args)
            self._clip_sample(x, kwargs)
            self._callback(callback, x_before, i, sigmas[i], den_out)
        return x


@register("DPM++ 2m")
class DPMPlusPlus2mSamplerWrapper(DiffusionSamplerWrapper):
    def __init__(self, name, **kwargs):
        kwargs["constructor"] = DPMPlusPlus2mDiffusionSampler
        super().__init__(name, **kwargs)
