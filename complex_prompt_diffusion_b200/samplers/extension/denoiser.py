"""Denoiser: composable multi-prompt classifier-free guidance around the UNet.

Mirror of the reference interface `cpd/samplers/extension/denoiser.py` (`Denoiser(unet, vae, tokenizer,
clip_model, decode, quantize=False, **kw)`, `forward(x, sigma, **kw) -> denoised`, `.scheduler`), keeping the
kwargs names of SURVEY.md Appendix B.  The arithmetic after the UNet call - the fp16 weighted delta
(:450-460), e_t = e_u + s * sum (:510-515), eps/v -> denoised (:533-542) - and the sampler update that
follows it run as ONE CUDA kernel (`cpd_sampler_step`); nothing is computed on the CPU and no device->host
synchronisation happens inside a step (the reference has five).

Differences that are deliberate repairs of reference defects: the image batch may be > 1 and
means B independent batch-1 trajectories (D7); sigma is passed once (D3); no hard-coded `.cuda()` (D6).
Built on the device: dynamic / static scale clipping and every runnable thresholding extension, the score-corrector
hook, feature / skip injection, the decaying guidance scale, the unconditional blur (:333-337,441-442), attention guidance
(:341-350,404-435,461-462) and the depth mask (:358-360,386-388).  CLIP guidance (a backward pass through VAE + CLIP) and
gamma > 0 (defective in the reference, D8) raise NotImplementedError when requested instead of being silently ignored.
"""
import math

import torch

from ... import ops
from ..._lib import (CPD_DENOISE_ONLY, CPD_DPMPP_2M, CPD_EULER, CPD_EULER_ANCESTRAL, CPD_PRED_EPSILON, CPD_PRED_VELOCITY,
                     CPD_THRESH_DYNAMIC, CPD_THRESH_DYNANORMIC, CPD_THRESH_RENORM, CPD_THRESH_SCALED_DYNAMIC_PERC,
                     CPD_THRESH_SCALED_NORM, CPD_THRESH_SCALED_SPATIAL_NORM, CPD_THRESH_SPATIAL_NORM, CPD_THRESH_STATIC)
from ...scheduler.discrete import SigmaScheduler

_UNSUPPORTED_TRUTHY = ("clip_guidance",)  # a backward pass through VAE + CLIP (denoiser.py:76-265): out of the hot-path scope

# Thresholding extensions, all on the device (registered names of samplers/extension/threshold.py).  The first two are
# clamps (the bound feeds the fused step directly); the others rewrite the tensor (cpd_threshold_ex).
# "norm_thresholding" (threshold.py:182-205) reads an undefined x_max and cannot run in the reference either.
THRESHOLD_ALGS = {"dynamic_thresholding": CPD_THRESH_DYNAMIC, "static_thresholding": CPD_THRESH_STATIC,
                  "dynanormic_thresholding": CPD_THRESH_DYNANORMIC,
                  "scaled_dynamic_perc_thresholding": CPD_THRESH_SCALED_DYNAMIC_PERC,
                  "renorm_thresholding": CPD_THRESH_RENORM, "scaled_norm_thresholding": CPD_THRESH_SCALED_NORM,
                  "spatial_norm_thresholding": CPD_THRESH_SPATIAL_NORM,
                  "scaled_spatial_norm_thresholding": CPD_THRESH_SCALED_SPATIAL_NORM}
_CLAMP_ALGS = (CPD_THRESH_DYNAMIC, CPD_THRESH_STATIC)


def threshold_alg(name):
    if name == "none":  # the base ScoreCorrector (threshold.py:7-45): identity
        return None
    if name not in THRESHOLD_ALGS:
        raise NotImplementedError(f"thresholding algorithm {name!r} is not built on the device (have: {sorted(THRESHOLD_ALGS)})")
    return THRESHOLD_ALGS[name]


def apply_threshold(x, bound, name, threshold):
    """extension(x, threshold=...) of the reference, in place on x: [n_images, 4, h, w] fp32 (values end up rounded through
    fp16, D10).  bound: [n_images] fp32 scratch."""
    alg = threshold_alg(name)
    if alg is None:
        return x
    if alg in _CLAMP_ALGS:
        ops.threshold(x, bound, alg=alg, threshold=float(threshold), clamp_inplace=True)
    else:
        ops.threshold_ex(x, bound, alg=alg, threshold=float(threshold))
    return x


class ConditioningPlan:
    """The conditioning dict {"and": [(scale, emb, guide, mask)...], "not": [...]} (prompts.py:622-648) flattened
    into what the kernels need: context rows [1 + N, 77, D] (row 0 = unconditional), per-sub-prompt weights
    rounded to the model dtype (P4: scale is moved to the UNet dtype before the fp16 cast, denoiser.py:373-381)
    and masks (scalar or [hw] fp32 device tensors holding model-dtype-rounded values)."""

    def __init__(self, c, uc, dtype, device, hw_shape):
        if not isinstance(c, dict) or "and" not in c:
            raise ValueError("conditioning must be a dict with an 'and' list (CompositionalPrompt._build_embeddings)")
        if uc is None:
            raise ValueError("unconditional_conditioning is required")
        entries = [(s, f, m) for (s, f, _g, m) in c["and"]] + [(-s, f, m) for (s, f, _g, m) in c.get("not", [])]
        if not 1 <= len(entries) <= ops.CPD_MAX_SUBPROMPTS:
            raise ValueError(f"between 1 and {ops.CPD_MAX_SUBPROMPTS} weighted sub-prompts are supported, got {len(entries)}")
        self.n_sub = len(entries)
        self.weights, self.mask_scalars, self.masks = [], [], []
        factors = [torch.as_tensor(uc)]
        h, w = hw_shape
        for scale, factor, mask in entries:
            factors.append(torch.as_tensor(factor))
            self.weights.append(float(torch.tensor([float(scale)]).to(dtype).float()))
            if isinstance(mask, (int, float)):
                self.mask_scalars.append(float(torch.tensor([float(mask)]).to(dtype).float()))
                self.masks.append(None)
            else:
                m = torch.as_tensor(mask)
                if m.numel() == 1:
                    self.mask_scalars.append(float(m.reshape(1).to(dtype).float()))
                    self.masks.append(None)
                else:
                    if m.numel() != h * w:
                        raise ValueError(f"mask must be [1,1,{h},{w}] (prompts.py:855), got {tuple(m.shape)}")
                    self.mask_scalars.append(1.0)
                    self.masks.append(m.reshape(h * w).to(dtype).float().contiguous().to(device))
        for f in factors:
            if f.ndim != 3 or f.shape[0] != 1:
                raise ValueError(f"each embedding must be [1, tokens, dim], got {tuple(f.shape)}")
        self.context = torch.cat([f.to(device) for f in factors]).contiguous()  # [1 + N, tokens, D]


class ReferenceUNetAdapter:
    """Lets the Denoiser drive ANY model with the reference's UNet call signature (denoiser.py:397-407):
    `unet(x [R*B, 4, h, w], timesteps [R*B], context [R*B, tokens, D], return_attn=True, ...) -> (out, skips)`, a bare tensor,
    or an object with `.sample` - e.g. the reference's own UNetModel or a diffusers-style module already living on the GPU.
    It builds the batch of denoiser.py:384-393 (x * c_in broadcast over the rows, the timestep repeated, the context rows
    tiled per image) on the device.  This is the compatibility path; `models.unet.UNetModel` provides `forward_rows` itself."""

    def __init__(self, unet):
        self.unet = unet
        p = next(iter(unet.parameters()))
        self.dtype, self.device = p.dtype, p.device
        self.eps_dtype = p.dtype
        self._ctx, self._y = None, None

    def parameters(self):
        return self.unet.parameters()

    def set_context(self, ctx):
        self._ctx = ctx.to(self.device, self.dtype)

    def set_vector(self, y):
        self._y = y.to(self.device, self.dtype)

    @torch.no_grad()
    def forward_rows(self, x, c_in, t, rows_per_image, inject=None):
        B, R = x.shape[0], rows_per_image
        if self._ctx is None or self._ctx.shape[0] != R:
            raise RuntimeError("ReferenceUNetAdapter: set_context must be called with one context row per conditioning row")
        x_in = (x * torch.tensor(c_in, dtype=torch.float32, device=x.device)).to(self.dtype)  # denoiser.py:390-391
        x_in = x_in[:, None].expand(B, R, *x.shape[1:]).reshape(B * R, *x.shape[1:])
        t_in = torch.full((B * R,), t, dtype=self.dtype, device=self.device)  # :384,393
        ctx = self._ctx[None].expand(B, *self._ctx.shape).reshape(B * R, *self._ctx.shape[1:])
        kw = {"return_attn": True}
        if self._y is not None:
            kw["y"] = self._y[None].expand(B, *self._y.shape).reshape(B * R, -1)
        if inject:
            kw.update(inject_feats=inject.get("feats"), inject_feats_stop=inject.get("feats_stop", 10),
                      inject_attns=inject.get("attns"), inject_attns_stop=inject.get("attns_stop", 10))
        out = self.unet(x_in, t_in, ctx, **kw)
        if isinstance(out, (tuple, list)):
            out = out[0]
        if hasattr(out, "sample"):  # :404-405
            out = out.sample
        return out.contiguous()


class Denoiser(torch.nn.Module):
    def __init__(self, unet, vae=None, tokenizer=None, clip_model=None, decode=None, quantize=False, **kwargs):
        super().__init__()
        self.name = kwargs.get("name", "Denoiser")
        if not hasattr(unet, "forward_rows"):  # a model with the reference's call signature
            unet = ReferenceUNetAdapter(unet)
        p = next(iter(unet.parameters()))
        self.dtype, self.device = p.dtype, p.device
        if self.device.type != "cuda":
            raise RuntimeError("Denoiser needs a UNet on a CUDA device: the hot path has no CPU fallback")
        self.scheduler = SigmaScheduler(num_train_timesteps=1000, **kwargs)
        self.quantize = quantize
        self.sigma_data = 1.0
        self.unet, self.vae, self.tokenizer, self.clip_model, self.decode = unet, vae, tokenizer, clip_model, decode
        self._plan_key, self._plan = None, None
        self._part, self._group = None, None

    def set_row_partition(self, part, group=None):
        """Multi-GPU row sharding (used when the image batch is smaller than the GPU count): this rank
        evaluates only `part.rows` of the (1 + N) conditioning rows of its image; the eps rows are all-gathered over
        NCCL inside `part.group_ranks` every step and every rank of the group runs the (tiny) fused step redundantly, so
        x stays replicated without a broadcast.  `part` comes from `dist.partition`; None switches it off."""
        if part is not None and not part.needs_allgather:
            part = None
        self._part, self._group = part, group
        self._plan_key = None

    # ---- host-side (once per sample() call) -----------------------------------------------------------
    def _check_kwargs(self, kwargs):
        for key in _UNSUPPORTED_TRUTHY:
            v = kwargs.get(key, None)
            if v is not None and v is not False:
                raise NotImplementedError(f"Denoiser kwarg {key!r} is outside the B200 hot-path scope (SURVEY.md 8-f)")
        if kwargs.get("gamma", 0):
            raise NotImplementedError("gamma > 0 is not supported (the reference branch is defective, SURVEY.md D8)")
        if kwargs.get("pred_type", "epsilon") not in ("epsilon", "velocity"):
            raise ValueError(f"unknown pred_type {kwargs.get('pred_type')!r}")

    @staticmethod
    def _versions(c, uc, y):
        """Version counters of every tensor the plan is built from (in-place edits of an embedding invalidate the plan)."""
        out = []
        entries = list(c.get("and", [])) + list(c.get("not", []))
        items = [uc, y]
        for (scale, emb, _guide, mask) in entries:
            items += [scale, emb, mask]
        for t in items:
            # (object, version): the object itself is held by the key, so an entry REPLACED by a fresh tensor of the same
            # layout (version 0 again) is seen as a different object - never compare id()s of objects that may have died
            out.append((t, t._version) if isinstance(t, torch.Tensor) else (None, float(t) if isinstance(t, (int, float)) else None))
        return out

    @staticmethod
    def _same_versions(a, b):
        return len(a) == len(b) and all(x[0] is y[0] and x[1] == y[1] for x, y in zip(a, b))

    def plan_conditioning(self, c, uc, hw_shape, y=None, force=False):
        """`y` (extension for UNets with vector conditioning, i.e. SDXL - not expressible by the reference): tensor
        [1 + N, adm] with one row per UNet row (row 0 = unconditional) or [1, adm] shared by all rows.
        The plan is rebuilt at the start of every sample() call (`force`) and, between direct Denoiser calls, whenever the
        conditioning OBJECTS or their tensor versions change.  The objects are kept alive by the cache: `id()` / `data_ptr()`
        of a freed object are recycled by Python / the caching allocator and must never key a cache."""
        if isinstance(c, list):
            raise NotImplementedError("per-step conditioning lists are not supported")
        if not isinstance(c, dict) or "and" not in c:
            raise ValueError("conditioning must be a dict with an 'and' list (CompositionalPrompt._build_embeddings)")
        key = (c, uc, y, tuple(hw_shape), self._versions(c, uc, y))
        k0 = self._plan_key
        same = (not force and k0 is not None and k0[0] is c and k0[1] is uc and k0[2] is y and k0[3] == key[3] and self._same_versions(k0[4], key[4]))
        if not same:
            self._plan = ConditioningPlan(c, uc, self.dtype, self.device, hw_shape)
            self._plan_key = key
            rows = None if self._part is None else self._part.rows
            if rows is None:
                self.unet.set_context(self._plan.context)
            elif rows:
                self.unet.set_context(self._plan.context[rows].contiguous())
            if y is not None or getattr(self.unet, "adm", 0):
                if y is None:
                    raise ValueError("this UNet needs vector conditioning: pass y=[1 + N, adm] (row 0 = unconditional)")
                y = torch.as_tensor(y)
                if y.shape[0] == 1:
                    y = y.expand(1 + self._plan.n_sub, -1)
                if y.shape[0] != 1 + self._plan.n_sub:
                    raise ValueError(f"y must have 1 or {1 + self._plan.n_sub} rows, got {y.shape[0]}")
                if rows is None:
                    self.unet.set_vector(y.contiguous())
                elif rows:
                    self.unet.set_vector(y[rows].contiguous())
        return self._plan

    @staticmethod
    def guidance_scale(**kwargs):
        """unconditional_guidance_scale with the optional log decay of denoiser.py:475-494."""
        s = kwargs.get("unconditional_guidance_scale", 1.0)
        t_idx, total = kwargs.get("t_idx", 0), kwargs.get("total_steps", 1000)
        if kwargs.get("decaying_uc_scale", False):
            start = kwargs.get("decaying_uc_scale_start", int(total * 0.2))
            if start < t_idx:
                start = min(t_idx, start)
                s = max(kwargs.get("decaying_uc_scale_min", 2), s - s * (math.log(t_idx + 1 - start) / math.log(total)))
        return float(s)

    @staticmethod
    def guidance_branches(**kwargs):
        """Which optional branches are active at this schedule index (denoiser.py:333-337,341-343): both switch on for the last
        `rounds` indices only.  Returns (uc_blur_kernel_size or None, attention-guidance settings dict or None)."""
        t_idx, total = kwargs.get("t_idx", 0), kwargs.get("total_steps", 1000)
        blur_k = None
        if kwargs.get("unconditional_guidance_blur", False) and t_idx > (total - kwargs.get("unconditional_guidance_blur_rounds", int(total / 10))):
            blur_k = int(kwargs.get("unconditional_guidance_blur_k", 7))
        guide = None
        if kwargs.get("attn_guide", kwargs.get("return_attn", False)) and \
                t_idx > (total - kwargs.get("attn_guide_rounds", kwargs.get("return_attn_rounds", 4))):
            guide = dict(mode=kwargs.get("attn_guide_mode", 2), idx=kwargs.get("attn_guide_idx", kwargs.get("return_attn_idx", -1)),
                         scale=kwargs.get("attn_guide_scale", 1.1),
                         threshold=kwargs.get("attn_guide_mask_threshold", kwargs.get("attn_mask_threshold", 90)),
                         blur_k=int(kwargs.get("attn_guide_blur_k", 31)))
            if guide["mode"] not in (1, 2):
                raise ValueError(f"attn_guide_mode must be 1 or 2, got {guide['mode']}")
        return blur_k, guide

    @staticmethod
    def wants_host_between_kernels(kwargs):
        """True when a step needs host decisions between its kernels (the optional guidance branches, the depth mask): such steps
        cannot be replayed as one captured graph."""
        return bool(kwargs.get("unconditional_guidance_blur", False) or kwargs.get("attn_guide", kwargs.get("return_attn", False)) or
                    kwargs.get("depth_mask", None) is not None)

    @staticmethod
    def _random_blur_sigma():
        """torchvision GaussianBlur(kernel_size) draws its sigma from U(0.1, 2.0) with the global CPU generator on every call
        (GaussianBlur.get_params): the same draw, so a seeded run consumes the RNG like the reference."""
        return torch.empty(1).uniform_(0.1, 2.0).item()

    def _attention_guidance(self, x, eps, plan, guide, sig, c_in, **kwargs):
        """denoiser.py:404-435: saliency mask from the channel mean of a skip tensor, blurred denoised sample blended into the
        latent under the mask, one extra UNet evaluation of the guided latent on the unconditional context.  Returns e_attn
        [B, 4, h, w] (eps dtype).  Everything runs on the device."""
        from ...models.unet import UNetModel
        unet = self.unet
        if not isinstance(unet, UNetModel):
            raise NotImplementedError("attention guidance needs the library's UNetModel (it reads the skip tensors of the plan)")
        if self._part is not None:
            raise NotImplementedError("attention guidance together with row sharding")
        B, R = x.shape[0], 1 + plan.n_sub
        hw = x.shape[2] * x.shape[3]
        n_in = len(unet.inputs)
        idx = guide["idx"] % n_in  # index into the POPPED skip list: 0 = last input block ... -1 = the input conv's output
        src = unet._tap(0, n_in - 1 - idx, B * R)  # NCHW view of the NHWC buffer: [B*R, c, h, w]
        if tuple(src.shape[2:]) != tuple(x.shape[2:]):
            raise ValueError(f"attn_guide_idx={guide['idx']} selects a {tuple(src.shape[2:])} tensor; the mask must have the latent's size {tuple(x.shape[2:])}")
        cch = src.shape[1]
        nhwc = src.permute(0, 2, 3, 1)  # the plan's own contiguous NHWC memory
        mean = torch.empty(B * R, hw, dtype=torch.float32, device=x.device)
        ops.channel_mean(nhwc, mean, pixels=B * R * hw, c=cch)
        pct = torch.empty(B, dtype=torch.float32, device=x.device)
        for b in range(B):  # np.percentile over ALL rows of the image's mask tensor (:411)
            ops.percentile(mean[b * R:(b + 1) * R].reshape(-1), pct[b:b + 1], q=float(guide["threshold"]))
        L = 4 * hw
        sample = torch.empty_like(x)
        ops.attn_guide(0, sample, n_images=B, hw=hw, x=x, eps_u=eps, eps_stride=R * L, sigma_hat=sig)
        blurred = torch.empty_like(x)
        ops.gaussian_blur(sample, blurred, kernel_size=guide["blur_k"], sigma=self._random_blur_sigma())
        guide_x = torch.empty_like(x)
        ops.attn_guide(1, guide_x, n_images=B, hw=hw, x=x, eps_u=eps, eps_stride=R * L, mask_mean=mean, mask_img_stride=R * hw, pct=pct,
                       blur=blurred, sigma_hat=sig, c_in=c_in, mode=guide["mode"])
        # :362,430: the guided latent is evaluated at t = sigma_to_t(sigma) in fp64 -> fp32 (NOT rounded to the model dtype) on
        # the unconditional context; afterwards the prompt's context rows are cached again
        t64 = self.scheduler.sigma_to_t(torch.tensor([sig], dtype=torch.float32))
        uc_ctx = plan.context[0:1]
        e_attn = unet.forward(guide_x, timesteps=t64.to(torch.float32).expand(B), context=uc_ctx)
        unet.set_context(plan.context)
        return e_attn.contiguous()

    # ---- device-side --------------------------------------------------------------------------------
    @staticmethod
    def _inject(kwargs):
        """Feature / skip injection kwargs of denoiser.py:353-356, handed to the UNet like :397-402 (lists of tensors with
        one [R, c, h, w] entry per output block, applied to blocks below the *_stop index)."""
        if kwargs.get("inject_feats") is None and kwargs.get("inject_attns") is None:
            return None
        return dict(feats=kwargs.get("inject_feats"), feats_stop=kwargs.get("inject_feats_stop", 10),
                    attns=kwargs.get("inject_attns"), attns_stop=kwargs.get("inject_attns_stop", 10))

    @staticmethod
    def _one_sigma(sigma):
        """All images of a batch share the schedule (B independent trajectories, D7): a per-image sigma vector must hold one
        value.  Returns a 1-element fp32 CPU tensor."""
        sig = torch.as_tensor(sigma, dtype=torch.float32).reshape(-1).cpu()
        if sig.numel() > 1 and not bool((sig == sig[0]).all()):
            raise ValueError("Denoiser: the images of a batch must share one sigma (got a vector with different values)")
        return sig[:1]

    def step_scalars(self, sigma, **kwargs):
        """The per-step scalars of one Denoiser evaluation exactly as unet_rows / fused_step compute them (fp32 torch expressions
        on 0-dim tensors like the reference's): c_in (denoiser.py:390), t in the model dtype (:393), the guidance scale after its
        optional decay (:475-494), sigma_hat and the v-prediction coefficients (:540-542)."""
        sig = self._one_sigma(sigma)
        c_in = 1 / (sig ** 2 + 1 ** 2) ** 0.5
        t = self.scheduler.sigma_to_t(sig).to(self.dtype).float()
        s = float(sig[0])
        sig_t = torch.tensor([s], dtype=torch.float32)
        return dict(c_in=float(c_in), t=float(t), guidance=self.guidance_scale(**kwargs), sigma_hat=s,
                    v_c_eps=float(-sig_t / (sig_t ** 2 + 1) ** 0.5), v_c_x_div=float(sig_t ** 2 + 1))

    def unet_rows(self, x, sigma, plan, inject=None, depth_mask=None):
        """Run the UNet on the (1 + N) conditioning rows of every image: returns eps rows [B*(1+N), 4, h, w]
        (image-major).  x: [B,4,h,w] fp32; sigma: 0-dim/1-element fp32 CPU tensor (same for all images)."""
        sig = self._one_sigma(sigma)
        c_in = 1 / (sig ** 2 + 1 ** 2) ** 0.5  # get_scalings, fp32 like denoiser.py:390
        t = self.scheduler.sigma_to_t(sig).to(self.dtype).float()  # fp64 -> model dtype (P3, denoiser.py:393)
        if depth_mask is not None:
            # :358-360,386-388: the depth map rides along as an extra input channel of a depth-conditioned UNet (in_channels = 5)
            # and is scaled by c_in with the latent; the same map for every image of the batch
            d = torch.as_tensor(depth_mask).to(x.device, torch.float32)
            d = d.reshape(-1, *d.shape[-3:]) if d.ndim >= 3 else d.reshape(1, 1, *d.shape)
            x = torch.cat([x, d[:1].expand(x.shape[0], -1, -1, -1)], dim=1).contiguous()
        if self._part is None:
            if inject:
                return self.unet.forward_rows(x, float(c_in), float(t), rows_per_image=1 + plan.n_sub, inject=inject)
            return self.unet.forward_rows(x, float(c_in), float(t), rows_per_image=1 + plan.n_sub)
        if inject:
            raise NotImplementedError("feature injection together with row sharding")
        # row-sharded: my rows of my (single) image, then the NCCL all-gather of eps inside the image's rank group
        from ... import dist as D
        if x.shape[0] != 1:
            raise ValueError("row sharding works on one image per rank group (shard the batch across groups first)")
        L = x[0].numel()
        if self._part.rows:
            local = self.unet.forward_rows(x, float(c_in), float(t), rows_per_image=len(self._part.rows)).reshape(len(self._part.rows), L)
        else:
            local = torch.empty(0, L, dtype=self.unet.eps_dtype, device=self.device)
        full = D.allgather_eps_rows(local, self._part, group=self._group)
        return full.reshape(1 + plan.n_sub, *x.shape[1:]).contiguous()

    def fused_step(self, x, sigma, plan, step, eps=None, **kwargs):
        """One UNet evaluation + ONE fused kernel: CFG combine, denoised, sampler update (x updated in place).
        `step` is a dict of the fp32 scalars for cpd_sampler_step; `eps` = already evaluated UNet rows (skips the UNet)."""
        if eps is None:
            eps = self.unet_rows(x, sigma, plan, inject=self._inject(kwargs), depth_mask=kwargs.get("depth_mask", None))
        sig = float(self._one_sigma(sigma)[0])
        blur_k, guide = self.guidance_branches(**kwargs)
        e_attn = None
        if guide is not None:  # before the blur: the reference builds the guided latent from the unblurred out[0] (:419-435)
            sig32 = torch.tensor([sig], dtype=torch.float32)
            e_attn = self._attention_guidance(x, eps, plan, guide, sig, float(1 / (sig32 ** 2 + 1 ** 2) ** 0.5), **kwargs)
        if blur_k is not None:  # :441-442: e_t_uncond = GaussianBlur(k)(e_t_uncond), a new random sigma per call
            B, R = x.shape[0], 1 + plan.n_sub
            rows = eps.view(B, R, *x.shape[1:])
            if rows.dtype != torch.float32:
                raise NotImplementedError("the unconditional blur runs on fp32 eps rows (UNetModel(eps_dtype=torch.float32))")
            blurred = torch.empty_like(x)
            ops.gaussian_blur(rows[:, 0], blurred, kernel_size=blur_k, sigma=self._random_blur_sigma(), img_stride_src=R * x[0].numel())
            rows[:, 0].copy_(blurred)
        sig_t = torch.tensor([sig], dtype=torch.float32)
        pred = CPD_PRED_VELOCITY if kwargs.get("pred_type", "epsilon") == "velocity" else CPD_PRED_EPSILON
        common = dict(n_sub=plan.n_sub, weights=plan.weights, mask_scalars=plan.mask_scalars, masks=plan.masks,
                      guidance=self.guidance_scale(**kwargs), pred_type=pred, sigma_hat=sig,
                      v_c_eps=float(-sig_t / (sig_t ** 2 + 1) ** 0.5), v_c_x_div=float(sig_t ** 2 + 1))
        clip = scaled_in = None
        if e_attn is not None:
            # :461-462,514: sum_e_t = e_attn + scale * (sum_e_t - e_attn) in fp32 (an fp16 tensor minus an fp32 one promotes), then
            # the guidance scale.  A combine-only pass with guidance 1 hands out the fp16 sum itself.
            sum16 = torch.empty_like(x)
            ops.sampler_step(eps, x, sampler=CPD_DENOISE_ONLY, scaled_out=sum16, **dict(common, guidance=1.0))
            scaled_in = torch.empty_like(x)
            ops.attn_guide(2, scaled_in, n_images=x.shape[0], hw=x.shape[2] * x.shape[3], sum16=sum16, e_attn=e_attn, scale=guide["scale"],
                           guidance=common["guidance"])
            if kwargs.get("scaled_clip", kwargs.get("dynamic_scale_clip", False)):  # :499-512 on the mixed term
                bound = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
                apply_threshold(scaled_in, bound, kwargs.get("scaled_clip_alg", "dynamic_thresholding"),
                                kwargs.get("scaled_clip_threshold", kwargs.get("dynamic_scale_clip_threshold", 99.5)))
        elif kwargs.get("scaled_clip", kwargs.get("dynamic_scale_clip", False)):
            # Dynamic scale clip (denoiser.py:499-512): the scaled guidance term s * sum_e_t is thresholded before it is
            # added to e_u.  The reference runs np.percentile on a CPU copy every step; here a combine-only pass writes
            # the term, cpd_threshold finds the per-image bound on the device and the fused step clamps with it.
            alg = threshold_alg(kwargs.get("scaled_clip_alg", "dynamic_thresholding"))
            thr = kwargs.get("scaled_clip_threshold", kwargs.get("dynamic_scale_clip_threshold", 99.5))
            if alg is not None:
                scaled = torch.empty_like(x)
                bound = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
                ops.sampler_step(eps, x, sampler=CPD_DENOISE_ONLY, scaled_out=scaled, **common)
                if alg in _CLAMP_ALGS:
                    ops.threshold(scaled, bound, alg=alg, threshold=float(thr), clamp_inplace=False)
                    clip = bound
                else:  # not a clamp: rewrite the term and feed it back
                    ops.threshold_ex(scaled, bound, alg=alg, threshold=float(thr))
                    scaled_in = scaled
        corrector = kwargs.get("score_corrector", None)
        if corrector is not None:
            # denoiser.py:517-518: e_t = score_corrector.modify_score(e_t, x, t, c, **corrector_kwargs).  A combine-only
            # pass materialises e_t, the corrector rewrites it (the registered extensions do so on the device,
            # extension/threshold.py; any object with the reference's modify_score signature works), and the update runs
            # on the rewritten row (n_sub = 0).
            e_t = torch.empty_like(x)
            ops.sampler_step(eps, x, sampler=CPD_DENOISE_ONLY, eps_out=e_t, clip_scaled=clip, scaled_in=scaled_in, **common)
            ck = dict(kwargs.get("corrector_kwargs", None) or {})
            ck["verbose"] = kwargs.get("verbose", False)  # :505
            e_t = corrector.modify_score(e_t, x.clone(), self.scheduler.sigma_to_t(sig_t), kwargs.get("conditioning"), **ck)
            if tuple(e_t.shape) != tuple(x.shape):
                raise ValueError(f"score_corrector returned shape {tuple(e_t.shape)}, expected {tuple(x.shape)}")
            e_t = e_t.to(x.device, torch.float32).contiguous()
            single = dict(common, n_sub=0, weights=(), mask_scalars=(), masks=None)
            ops.sampler_step(e_t, x, **single, **step)
            return x
        ops.sampler_step(eps, x, clip_scaled=clip, scaled_in=scaled_in, **common, **step)
        return x

    @torch.no_grad()
    def evaluate(self, x, sigma, **kwargs):
        """One Denoiser evaluation with its intermediates (parity checks: tests, bench.py's `parity` key): the UNet rows
        [B * (1 + N), 4, h, w] (denoiser.py:397-402, :439), the combined e_t (:515) and the denoised sample (:540/:542)."""
        self._check_kwargs(kwargs)
        x = x.to(self.device, torch.float32).contiguous()
        plan = self.plan_conditioning(kwargs.get("conditioning"), kwargs.get("unconditional_conditioning"), x.shape[-2:],
                                      y=kwargs.get("y"))
        rows = self.unet_rows(x, sigma, plan, inject=self._inject(kwargs), depth_mask=kwargs.get("depth_mask", None)).clone()
        e_t, denoised = torch.empty_like(x), torch.empty_like(x)
        self.fused_step(x.clone(), sigma, plan, dict(sampler=CPD_DENOISE_ONLY, denoised_out=denoised, eps_out=e_t), eps=rows, **kwargs)
        return dict(rows=rows, e_t=e_t, denoised=denoised)

    @torch.no_grad()
    def forward(self, x, sigma, **kwargs):
        """Reference-compatible entry point: returns the denoised sample for x:[B,4,h,w] (fp32)."""
        if x.ndim != 4 or x.shape[1] != 4:
            raise ValueError(f"[denoiser] `x` has incorrect shape: {tuple(x.shape)}")
        self._check_kwargs(kwargs)
        x = x.to(self.device, torch.float32).contiguous()
        plan = self.plan_conditioning(kwargs.get("conditioning"), kwargs.get("unconditional_conditioning"), x.shape[-2:],
                                      y=kwargs.get("y"))
        denoised = torch.empty_like(x)
        scratch = x.clone()
        self.fused_step(scratch, sigma, plan, dict(sampler=CPD_DENOISE_ONLY, denoised_out=denoised), **kwargs)
        return denoised
