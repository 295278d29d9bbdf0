"""Oracle (test infrastructure): composable multi-prompt CFG denoiser.

Follows cpd/samplers/extension/denoiser.py:324-463 (_process_conditioning), :465-521
(_calculate_epsilon) and :528-544 (forward); paths relative to /root/reference.  Only the default
branches are restated (no attention guidance, blur, CLIP guidance, score corrector, depth mask);
defect repairs D3, D6, D7, D8 are described in oracle/__init__.py.
"""
import numpy as np
import torch

from .schedule import OracleSchedule


def _safe_to(x, dtype):
    """cpd/util.py:399-425 for the cases on the path: python number -> 1-element tensor; cast."""
    if isinstance(x, (int, float)):
        x = torch.Tensor([x])
    elif isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x.to(dtype)


def append_dims(x, target_dims):
    """denoiser.py:549-554."""
    return x[(...,) + (None,) * (target_dims - x.ndim)]


def combine_fp16(e_t_out, e_t_uncond, e_scales, e_masks):
    """denoiser.py:450-460: sum_k half(m_k)*half(w_k)*(half(e_k) - half(e_u)), python sum from int 0."""
    return sum([
        e_masks[i].to(torch.float16) * e_scales[i].to(torch.float16) *
        (e_t.to(torch.float16) - e_t_uncond.to(torch.float16))
        for i, e_t in enumerate(e_t_out)])


def guidance_scale(uc_scale, t_idx, total_steps, decay=False, decay_min=2, decay_start=None):
    """denoiser.py:475-494 (decaying_uc_scale)."""
    if decay_start is None:
        decay_start = int(total_steps * 0.2)
    if decay and decay_start < t_idx:
        decay_start = min(t_idx, decay_start)
        uc_scale = max(decay_min, uc_scale - (uc_scale * (np.log(t_idx + 1 - decay_start) / np.log(total_steps))))
    return uc_scale


class OracleDenoiser:
    """Denoiser restatement for ONE image (x:[1,4,h,w]) - the only batch the reference supports (D7)."""

    def __init__(self, unet, dtype=None):
        self.unet = unet
        self.dtype = dtype if dtype is not None else next(unet.parameters()).dtype
        self.scheduler = OracleSchedule(1000)
        self.trace = None  # optional list collecting per-step intermediates

    def process_conditioning(self, x, c, sigma, **kwargs):
        uc = kwargs.get("unconditional_conditioning", None)
        e_factors, e_scales, e_masks = [], [], []
        assert "and" in c
        for (scale, factor, _, mask) in c["and"]:  # denoiser.py:369-375
            e_factors.append(_safe_to(factor, self.dtype))
            e_scales.append(_safe_to(scale, self.dtype))
            e_masks.append(_safe_to(mask, self.dtype))
        for (scale, factor, _, mask) in c.get("not", []):  # denoiser.py:376-381
            e_factors.append(_safe_to(factor, self.dtype))
            e_scales.append(_safe_to(-scale, self.dtype))
            e_masks.append(_safe_to(mask, self.dtype))
        bs = 1 + len(e_factors)
        t_in = torch.cat([sigma] * bs)  # denoiser.py:384
        f_uc = torch.cat([uc] + e_factors)  # denoiser.py:385
        _, sc_in = [append_dims(s, x.ndim) for s in self.scheduler.get_scalings(t_in)]  # :390
        x_in = x * sc_in  # :391  ([1,4,h,w] * [bs,1,1,1])
        t_full, low_idx, high_idx = self.scheduler.sigma_to_t_idx(t_in)
        t_in = t_full.to(self.dtype)  # :393 (P3: cast to the UNet parameter dtype)
        inj = dict(inject_feats=kwargs.get("inject_feats", None), inject_feats_stop=kwargs.get("inject_feats_stop", 10),
                   inject_attns=kwargs.get("inject_attns", None), inject_attns_stop=kwargs.get("inject_attns_stop", 10))  # :353-356
        if inj["inject_feats"] is None and inj["inject_attns"] is None:
            inj = {}
        if kwargs.get("y") is not None:  # SDXL extension: vector conditioning rows [1 + N, adm] (row 0 = unconditional) or [1, adm]
            y = torch.as_tensor(kwargs["y"])
            y = y.expand(bs, -1) if y.shape[0] == 1 else y
            out, _skips = self.unet(x_in, t_in, f_uc, return_attn=True, y=_safe_to(y, self.dtype), **inj)
        else:
            out, _skips = self.unet(x_in, t_in, f_uc, return_attn=True, **inj)  # :397-402
        e_t_out = list(out.chunk(bs))  # :439
        e_t_uncond = e_t_out.pop(0)  # :440
        sum_e_t = combine_fp16(e_t_out, e_t_uncond, e_scales, e_masks)  # :450-460
        if self.trace is not None:
            self.trace.append({"low_idx": low_idx.clone(), "high_idx": high_idx.clone(), "t": t_full.clone(),
                               "unet_out": out.detach().float().clone()})
        return sum_e_t, e_t_uncond

    @torch.no_grad()
    def calculate_epsilon(self, x, **kwargs):
        sigma = kwargs.get("sigma")
        c = kwargs.get("conditioning")
        t_idx = kwargs.get("t_idx", 0)
        total_steps = kwargs.get("total_steps", 1000)
        uc_scale = guidance_scale(kwargs.get("unconditional_guidance_scale", 1.0), t_idx, total_steps,
                                  kwargs.get("decaying_uc_scale", False), kwargs.get("decaying_uc_scale_min", 2),
                                  kwargs.get("decaying_uc_scale_start", None))
        sum_e_t, e_t_uncond = self.process_conditioning(x, c, **kwargs)
        scaled_e_t = uc_scale * sum_e_t  # denoiser.py:514 (fp16 tensor * python float stays fp16)
        if kwargs.get("scaled_clip", kwargs.get("dynamic_scale_clip", False)):  # :499-512: dynamic scale clip
            from .samplers import threshold_apply
            thr = kwargs.get("scaled_clip_threshold", kwargs.get("dynamic_scale_clip_threshold", 99.5))
            scaled_e_t = threshold_apply(scaled_e_t, kwargs.get("scaled_clip_alg", "dynamic_thresholding"), thr).half()
        e_t = e_t_uncond + scaled_e_t  # :515
        corrector = kwargs.get("score_corrector", None)
        if corrector is not None:  # :517-518
            ck = dict(kwargs.get("corrector_kwargs", None) or {})
            ck["verbose"] = kwargs.get("verbose", False)  # :505
            e_t = corrector.modify_score(e_t, x, self.scheduler.sigma_to_t(sigma), c, **ck)
        return e_t

    def __call__(self, x, sigma, **kwargs):
        """denoiser.py:528-544.  Returns the denoised sample (gamma = 0)."""
        assert x.shape[1] == 4 and x.shape[0] == 1
        kwargs["sigma"] = sigma
        eps = self.calculate_epsilon(x.clone(), **kwargs)
        assert eps.shape == x.shape
        sigma_hat = sigma * (kwargs.get("gamma", 0) + 1)
        pred_type = kwargs.get("pred_type", "epsilon")
        if pred_type == "epsilon":
            sample = x - sigma_hat * eps  # :540
        elif pred_type == "velocity":
            sample = eps * (-sigma / (sigma ** 2 + 1) ** 0.5) + (x / (sigma ** 2 + 1))  # :542
        else:
            raise ValueError(pred_type)
        if self.trace is not None:
            self.trace[-1]["x"] = x.detach().float().clone()
            self.trace[-1]["sigma"] = torch.as_tensor(sigma).detach().clone()
            self.trace[-1]["eps"] = eps.detach().float().clone()
            self.trace[-1]["denoised"] = sample.detach().float().clone()
        return sample
