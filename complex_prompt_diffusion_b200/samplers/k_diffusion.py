"""KDiffusionSampler: schedule, initial noise scaling and dispatch to the per-sampler loop
(interface of cpd/samplers/k_diffusion.py:22-97).

All per-step scalars are computed once on the host with the same fp32 torch expressions the reference
evaluates on 0-dim tensors, then each step is one UNet evaluation plus one fused CUDA kernel
(`Denoiser.fused_step`).  B images are B independent trajectories sharing the schedule.
"""
import ctypes as C

import torch

from .extension.denoiser import Denoiser

STEP_TABLE_ROWS = 1024  # capacity of the per-schedule device table of step scalars (rows of 64 bytes)


class KDiffusionSampler:
    def __init__(self, model, name="sample_heun"):
        self.name = name
        self.noise_sync = None  # dist.sample_sharded: replaces a locally drawn noise tensor by the row-sharding group's shared draw
        self._step_graphs = {}  # one captured CUDA graph per (shape, prompt layout, sampler): select scalars -> UNet -> fused step
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

    # ---- one CUDA graph per sampler step ---------------------------------------------------------------
    def _step_graph_ok(self, plan, model_args, kwargs):
        """The whole step - [cpd_step_select -> cpd_unet_forward -> cpd_sampler_step] - replays as ONE captured graph when
        nothing in it needs the host between the kernels: the library's UNet (not a foreign model behind the adapter), no row
        sharding, no per-step callback / thresholding / score corrector / injection, no spatial masks."""
        from ..models.unet import UNetModel
        den = self.denoiser
        if not isinstance(den.unet, UNetModel) or not den.unet.use_cuda_graph or den._part is not None:
            return False
        if not kwargs.get("step_graph", True) or kwargs.get("callback") is not None or kwargs.get("clip_sample", False):
            return False
        if model_args.get("scaled_clip", model_args.get("dynamic_scale_clip", False)) or model_args.get("score_corrector") is not None:
            return False
        if den._inject(model_args) is not None or any(m is not None for m in plan.masks) or den.wants_host_between_kernels(model_args):
            return False
        return True

    def _graph_loop(self, x, sigmas, plan, sampler, rows, model_args, noise_fn=None):
        """Run the schedule through one captured graph per step.  `rows[i]` = dict of the step's fused-step scalars (dt, sigma_up,
        dpm_*, write_old, noise_mul); the Denoiser's own scalars are added here.  All scalars of the schedule go to the device in
        ONE table at the start; per step the host only replays the graph (plus the noise draw of ancestral samplers)."""
        from .. import ops
        from .._lib import StepScalars, check, load, stream_ptr, CPD_PRED_EPSILON, CPD_PRED_VELOCITY
        den, unet, dev = self.denoiser, self.denoiser.unet, x.device
        n = len(sigmas) - 1
        if n > STEP_TABLE_ROWS:
            raise ValueError(f"more than {STEP_TABLE_ROWS} sampler steps")
        table = (StepScalars * n)()
        for i in range(n):
            model_args["t_idx"] = i
            sc = dict(den.step_scalars(rows[i].get("sigma", sigmas[i]), **model_args))
            sc.update({k: v for k, v in rows[i].items() if k != "sigma"})
            r = table[i]
            r.noise_mul, r.dpm_first = 1.0, 1
            for k, v in sc.items():
                setattr(r, k, v)
        pred = CPD_PRED_VELOCITY if model_args.get("pred_type", "epsilon") == "velocity" else CPD_PRED_EPSILON
        R = 1 + plan.n_sub
        key = (tuple(x.shape), plan.n_sub, tuple(plan.weights), tuple(plan.mask_scalars), int(sampler), pred, unet._ctx_shape, unet._y_rows,
               noise_fn is not None)
        st = self._step_graphs.get(key)
        lib = load()
        with torch.cuda.device(dev):
            if st is None:
                if len(self._step_graphs) >= 4:  # prompts with new weights re-capture: keep the cache small
                    self._step_graphs.pop(next(iter(self._step_graphs)))
                st = dict(x=torch.empty_like(x), old=torch.empty_like(x), noise=torch.zeros_like(x) if noise_fn is not None else None,
                          eps=torch.empty(x.shape[0] * R, *x.shape[1:], dtype=unet.eps_dtype, device=dev),
                          counter=torch.zeros(1, dtype=torch.int32, device=dev), cur=torch.zeros(16, dtype=torch.float32, device=dev),
                          table=torch.zeros(STEP_TABLE_ROWS * 16, dtype=torch.float32, device=dev),
                          host=torch.zeros(STEP_TABLE_ROWS * 16, dtype=torch.float32).pin_memory())
                st["x"].copy_(x)
                st["cur"][0] = 1.0

                def one_step():
                    check(lib.cpd_step_select(C.c_void_p(st["table"].data_ptr()), STEP_TABLE_ROWS, C.c_void_p(st["counter"].data_ptr()),
                                              C.c_void_p(st["cur"].data_ptr()), stream_ptr(dev)), "cpd_step_select")
                    ops._count()
                    unet._run(st["x"], R, st["cur"][0:1], st["cur"][1:2], 1, st["eps"], no_graph=True)
                    ops.sampler_step(st["eps"], st["x"], n_sub=plan.n_sub, weights=plan.weights, mask_scalars=plan.mask_scalars, masks=plan.masks,
                                     guidance=0.0, sampler=sampler, pred_type=pred, sigma_hat=1.0, old_denoised=st["old"], noise=st["noise"],
                                     dyn=st["cur"])
                # eager warm-up of the UNet shape (allocates the plan's workspace, times the GEMM variants), then the capture
                unet._run(st["x"], R, st["cur"][0:1], st["cur"][1:2], 1, st["eps"], no_graph=True)
                torch.cuda.synchronize(dev)
                n0 = ops.LAUNCHES
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    one_step()
                st.update(graph=graph, launches=ops.LAUNCHES - n0)
                self._step_graphs[key] = st
            st["host"][: n * 16].copy_(torch.frombuffer(bytearray(bytes(table)), dtype=torch.float32))
            st["table"][: n * 16].copy_(st["host"][: n * 16], non_blocking=True)
            st["x"].copy_(x, non_blocking=True)
            st["counter"].zero_()
            for i in range(n):
                if noise_fn is not None:
                    st["noise"].copy_(noise_fn(st["x"]).to(dev, torch.float32), non_blocking=True)
                st["graph"].replay()
                ops.LAUNCHES += st["launches"]
            return st["x"].clone()

    def _sampling(self, x, sigmas, model_args=None, **kwargs):
        raise NotImplementedError()
