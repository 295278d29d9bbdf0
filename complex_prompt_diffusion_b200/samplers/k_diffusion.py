"""KDiffusionSampler: schedule, initial noise scaling and dispatch to the per-sampler loop
(interface of cpd/samplers/k_diffusion.py:22-97).

All per-step scalars are computed once on the host with the same fp32 torch expressions the reference
evaluates on 0-dim tensors, then each step is one UNet evaluation plus one fused CUDA kernel
(`Denoiser.fused_step`).  B images are B independent trajectories sharing the schedule.
"""
import torch

from .extension.denoiser import Denoiser


class KDiffusionSampler:
    def __init__(self, model, name="sample_heun"):
        self.name = name
        self.noise_sync = None  # dist.sample_sharded: replaces a locally drawn noise tensor by the row-sharding group's shared draw
        self.denoiser = Denoiser(model["unet"], model.get("vae"), model.get("tokenizer"), model.get("clip_new_model"),
                                 model.get("decode"))

    # ---- schedule helpers ----------------------------------------------------------------------------
    def _sigmas(self, steps, kwargs):
        sigmas = self.denoiser.scheduler.get_sigmas(kwargs.get("scheduler", "default"), steps, **kwargs)
        # The "linear"/"default" schedule is fp64 in the reference; its Euler loops then promote x to fp64 and
        # crash in the UNet (defect D9, see oracle/make_golden.py), DPM++ 2M keeps x fp32.  Here the loop always
        # runs on fp32 scalars.
        return sigmas

    def sample(self, steps, batch_size, shape, **kwargs):
        dev = self.denoiser.device
        decode = kwargs.get("decode", False)
        x_T = kwargs.get("x_T", None)
        sigmas = self._sigmas(steps, kwargs)
        if decode:  # img2img: truncated schedule, x_T is the encoded image (k_diffusion.py:64-70)
            strength = kwargs.get("denoising_strength", 0.0)
            t_enc = int((1 - min(strength, 0.999)) * steps)
            sigmas = sigmas[steps - t_enc - 1:]
            noise = torch.randn([batch_size] + list(shape))
            if self.noise_sync is not None:  # row-sharded group: every rank must start from the same noised image
                noise = self.noise_sync(noise.to(dev)).cpu()
            x = x_T.to(dev, torch.float32) + (noise * sigmas[0]).to(dev, torch.float32)
        else:
            if x_T is None:
                x_T = torch.randn([batch_size] + list(shape))
            x = (x_T.to("cpu", torch.float32) * sigmas[0]).to(torch.float32).to(dev) if not x_T.is_cuda \
                else (x_T.float() * sigmas[0].to(dev)).float()
        kwargs["total_steps"] = len(sigmas)
        return self._sampling(x.contiguous(), sigmas, model_args=kwargs, disable=kwargs.get("silent", False), **kwargs)

    def sample_img2img(self, x, noise, steps, **kwargs):
        dev = self.denoiser.device
        strength = kwargs.get("denoising_strength", 0.0)
        t_enc = int((1 - min(strength, 0.999)) * steps)
        sigmas = self._sigmas(steps, kwargs)
        xi = x.to(dev, torch.float32) + (noise.cpu().float() * sigmas[steps - t_enc - 1]).to(dev, torch.float32)
        sched = sigmas[steps - t_enc - 1:]
        kwargs["total_steps"] = len(sched)
        return self._sampling(xi.contiguous(), sched, model_args=kwargs, disable=not kwargs.get("verbose", False), **kwargs)

    # ---- shared loop plumbing -------------------------------------------------------------------------
    def _begin(self, x, model_args, kwargs):
        den = self.denoiser
        den._check_kwargs(model_args)
        plan = den.plan_conditioning(model_args.get("conditioning"), model_args.get("unconditional_conditioning"), x.shape[-2:],
                                     y=model_args.get("y"), force=True)  # once per sample() call: never a stale prompt
        return den, plan

    def _churn(self, x, sigmas, i, kwargs):
        """Stochastic churn of Karras et al. Algorithm 2 (euler.py:40-46, huen.py:38-43, dpm2.py:38-43): one noise tensor is
        drawn per step (also when gamma = 0: the reference consumes the RNG either way; `rng_compat=False` skips the unused
        draw), sigma_hat = sigma * (gamma + 1) and, when gamma > 0, x += (noise * s_noise) * sqrt(sigma_hat^2 - sigma^2) on the
        device.  Returns sigma_hat (0-dim fp32 tensor)."""
        from .. import ops
        s_churn, s_tmin = kwargs.get("s_churn", 0.0), kwargs.get("s_tmin", 0.0)
        s_tmax, s_noise = kwargs.get("s_tmax", float("inf")), kwargs.get("s_noise", 1.0)
        gamma = min(s_churn / (len(sigmas) - 1), 2 ** 0.5 - 1) if s_tmin <= sigmas[i] <= s_tmax else 0.0
        noise_sampler = kwargs.get("noise_sampler", None)
        noise = None
        if gamma > 0 or kwargs.get("rng_compat", True):
            noise = noise_sampler(x) if noise_sampler is not None else torch.randn_like(x)
        sigma_hat = sigmas[i] * (gamma + 1)
        if gamma > 0:
            scale = (sigma_hat ** 2 - sigmas[i] ** 2) ** 0.5
            ops.add_noise(x, noise.to(x.device, torch.float32).contiguous(), noise_mul=float(s_noise), scale=float(scale))
        return sigma_hat

    def _clip_sample(self, x, kwargs):
        """Sample thresholding after the update (euler.py:55-56, dpmpp.py:51-52): x <- half(clamp(x, -s, s)) with
        s = max(percentile(|x|), 1) per image, all on the device (the reference goes through np.percentile on the CPU)."""
        if not kwargs.get("clip_sample", False):
            return
        from .extension.denoiser import apply_threshold
        if getattr(self, "_clip_bound", None) is None or self._clip_bound.numel() < x.shape[0] or self._clip_bound.device != x.device:
            self._clip_bound = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
        apply_threshold(x, self._clip_bound, kwargs.get("clip_sample_alg", "dynamic_thresholding"), kwargs.get("clip_sample_thresh", 90))

    def _callback(self, callback, x_before, i, sigma, denoised, sigma_hat=None):
        if callback is not None:
            callback({"x": x_before, "i": i, "sigma": sigma, "sigma_hat": sigma if sigma_hat is None else sigma_hat, "eps": denoised})

    def _sampling(self, x, sigmas, model_args=None, **kwargs):
        raise NotImplementedError()
