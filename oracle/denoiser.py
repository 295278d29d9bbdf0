"""Oracle (test infrastructure): composable multi-prompt CFG denoiser.

Follows cpd/samplers/extension/denoiser.py:324-463 (_process_conditioning), :465-521
(_calculate_epsilon) and :528-544 (forward); paths relative to /root/reference.  Restated branches:
the default path, the score-corrector hook, the scale clip, feature / skip injection, and (round 2) the unconditional blur
(:333-337, :441-442), attention guidance (:341-350, :410-435, :461-462) and the depth mask (:358-360, :386-388).  CLIP guidance
(:76-265: a backward pass through VAE + CLIP) is not restated.  Defect repairs D3, D6, D7, D8 are described in
oracle/__init__.py.

Third-party arithmetic on these branches: `torchvision.transforms.GaussianBlur(kernel_size=k)` (torchvision is not pinned by the
reference; 0.26.0 in the build container).  Its algorithm is restated in `gaussian_blur_random_sigma` with the same torch ops in
the same order: sigma ~ U(0.1, 2.0) from the global torch RNG (`torch.empty(1).uniform_(0.1, 2.0).item()`), kernel1d =
normalised exp(-0.5 (x / sigma)^2) on linspace(-(k-1)/2, (k-1)/2, k), kernel2d = mm(k1d[:, None], k1d[None, :]), reflect padding
of k // 2 and a depthwise conv2d.  Pinned against the shimmed reference by tests/golden/ref_sampling7.npz.
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


def gaussian_blur_random_sigma(img, kernel_size, sigma_range=(0.1, 2.0)):
    """torchvision.transforms.GaussianBlur(kernel_size)(img) (see the module docstring).  Returns (blurred, sigma)."""
    sigma = torch.empty(1).uniform_(sigma_range[0], sigma_range[1]).item()  # GaussianBlur.get_params
    ksize_half = (kernel_size - 1) * 0.5
    xs = torch.linspace(-ksize_half, ksize_half, steps=kernel_size, dtype=img.dtype)
    pdf = torch.exp(-0.5 * (xs / sigma).pow(2))
    k1 = pdf / pdf.sum()
    k2 = torch.mm(k1[:, None], k1[None, :])
    squeeze = img.ndim == 3
    t = img[None] if squeeze else img
    C = t.shape[-3]
    kernel = k2.expand(C, 1, kernel_size, kernel_size)
    pad = kernel_size // 2
    t = torch.nn.functional.pad(t, [pad, pad, pad, pad], mode="reflect")
    t = torch.nn.functional.conv2d(t, kernel, groups=C)
    return (t[0] if squeeze else t), sigma


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
        t_idx, total_steps = kwargs.get("t_idx", 0), kwargs.get("total_steps", 1000)
        # denoiser.py:333-337: the unconditional blur is active on the last `rounds` schedule indices
        uc_blur = kwargs.get("unconditional_guidance_blur", False)
        uc_blur_k = kwargs.get("unconditional_guidance_blur_k", 7)
        uc_blur_rounds = kwargs.get("unconditional_guidance_blur_rounds", int(total_steps / 10))
        uc_blur = uc_blur and (t_idx > (total_steps - uc_blur_rounds))
        # denoiser.py:341-350: attention guidance, likewise
        attn_guide = kwargs.get("attn_guide", kwargs.get("return_attn", False))
        attn_guide_mode = kwargs.get("attn_guide_mode", 2)
        attn_guide_rounds = kwargs.get("attn_guide_rounds", kwargs.get("return_attn_rounds", 4))
        attn_guide = attn_guide and (t_idx > (total_steps - attn_guide_rounds))
        attn_guide_idx = kwargs.get("attn_guide_idx", kwargs.get("return_attn_idx", -1))
        attn_guide_scale = kwargs.get("attn_guide_scale", 1.1)
        attn_guide_mask_threshold = kwargs.get("attn_guide_mask_threshold", kwargs.get("attn_mask_threshold", 90))
        attn_guide_blur_k = kwargs.get("attn_guide_blur_k", 31)
        depth_mask = kwargs.get("depth_mask", None)
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
        if depth_mask is not None:  # :358-360, :386-388: the depth map rides along as a fifth input channel (scaled by c_in too)
            x_depth = torch.cat([x[0], torch.as_tensor(depth_mask)[0]]).unsqueeze(0)
            x_in = x_depth * sc_in
        else:
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
        e_t_out_attn = None
        if attn_guide:  # :404-435
            attn = _skips[attn_guide_idx]
            mask = attn.mean(1, keepdims=True)  # :408
            s = np.percentile(mask.detach().cpu(), attn_guide_mask_threshold, axis=tuple(range(0, mask.ndim)))  # :411
            mask = mask.clone()
            mask[mask > s] = 1
            mask[mask < s] = 0
            mask = mask[0]
            sigma_hat = sigma * (kwargs.get("gamma", 0) + 1)  # gamma = 0 (D8)
            sample = x - (sigma_hat * out[0])  # :419
            blur_sample, blur_sigma = gaussian_blur_random_sigma(sample, attn_guide_blur_k)  # :420
            blur_x = blur_sample + (sigma_hat / out[0])  # :421 (a division, as written)
            masked_x = blur_x * mask  # :424
            if attn_guide_mode == 2:
                masked_x = masked_x * sc_in[0]  # :425-426
            guide_x = masked_x + (x * (1 - mask))  # :427
            if attn_guide_mode == 1:
                guide_x = guide_x * sc_in[0]  # :428-429
            t_guide = self.scheduler.sigma_to_t(sigma)  # :362: fp64, NOT cast to the model dtype
            attn_out = self.unet(guide_x, t_guide, uc)  # :430 (no return_attn: a bare tensor or (out, skips))
            if isinstance(attn_out, (tuple, list)):
                attn_out = attn_out[0]
            e_t_out_attn = attn_out[0]  # :435
            if self.trace is not None:
                self.guide_trace = dict(mask=mask.clone(), percentile=float(s), blur_sigma=blur_sigma, guide_x=guide_x.clone(),
                                        e_attn=e_t_out_attn.clone())
        e_t_out = list(out.chunk(bs))  # :439
        e_t_uncond = e_t_out.pop(0)  # :440
        if uc_blur:
            e_t_uncond, _ = gaussian_blur_random_sigma(e_t_uncond, uc_blur_k)  # :441-442
        sum_e_t = combine_fp16(e_t_out, e_t_uncond, e_scales, e_masks)  # :450-460
        if attn_guide:
            sum_e_t = e_t_out_attn + attn_guide_scale * (sum_e_t - e_t_out_attn)  # :461-462
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
