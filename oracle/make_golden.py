"""Generate the committed golden fixtures under tests/golden from the SHIMMED, UNMODIFIED reference.

Run once in the build container (needs /root/reference):   python -m oracle.make_golden
TEST INFRASTRUCTURE ONLY; never imported by the product, the gpu tests, smoke() or bench.py.

What is pinned (the reference itself has no tests / golden vectors):
  tests/golden/schedule_kat.json   - KScheduler tables, get_sigmas(karras|linear|exp|quad|vp), sigma_to_t
                                     integer indices (cpd/scheduler/k.py run as-is).
  tests/golden/ref_sampling.npz    - reference Denoiser + Euler / Euler Ancestral / DPM++ 2m samplers driving
                                     the reference UNetModel (tiny config, seeded non-zero weights) for
                                     one image with N=3 weighted sub-prompts (2 conjunctions incl. a spatial
                                     mask, 1 negation): inputs, per-step UNet inputs/outputs, denoised tensors and final latents.
  tests/golden/schedule_kat2.json  - SigmaScheduler.get_sigmas_{karras,exponential,quad,vp,sigmoid} of discrete.py (--schedule2-only)
  tests/golden/ref_sampling2.npz   - Heun / DPM2 / DPM2-a / DPM++ 2S-a / LMS and the Denoiser's scale clip (--more-only)
  tests/golden/ref_sampling3.npz   - stochastic churn, s_churn > 0 (--churn-only)
  tests/golden/ref_sampling4.npz   - score_corrector hook with the registered thresholding extensions (--corrector-only)
  tests/golden/ref_sampling5.npz   - img2img branch, decode=True + denoising_strength (--img2img-only)
  tests/golden/ref_sampling6.npz   - decaying guidance scale, decaying_uc_scale* (--decay-only)
  tests/golden/ref_sampling7.npz   - unconditional blur, attention guidance, depth mask branches of the Denoiser (--guidance-only)
  tests/golden/ref_unet_inject.npz - UNetModel with return_attn / return_feat / inject_attns / inject_feats (--inject-only)
  tests/golden/ref_noise.npz       - NoiseGenerator seed modes and draws (--noise-only)
  tests/golden/ref_threshold.npz   - every runnable thresholding extension on seeded tensors (--threshold-only)
  tests/golden/ref_prompts.npz     - WeightedPrompt._parse_prompt / CompositionalPrompt._parse_mask_style (--prompts-only)
  tests/golden/ref_vae.npz         - first-stage decoder (--vae-only)
Runtime repairs applied to the reference objects (no source edits), see SURVEY.md 8-c:
  D1  SigmaScheduler.append_zero added (discrete.py:107 calls it, it is defined on another class, :765).
  D2  SigmaScheduler.sigmas is given KScheduler's training table and get_sigmas no longer overwrites it.
  D3  _process_conditioning receives sigma once (denoiser.py:508 passes it positionally AND in kwargs).
  D4-D6 in oracle/ref_shim.py.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def schedule_kats(K):
    ks = K.KScheduler()
    out = {
        "n_distinct_betas": int(torch.unique(ks.betas).numel()),
        "betas": {str(i): float(ks.betas[i]) for i in (0, 1, 499, 998, 999)},
        "alphas_cumprod": {str(i): float(ks.alphas_cumprod[i]) for i in (0, 1, 499, 998, 999)},
        "sigmas_table": {str(i): float(ks.sigmas[i]) for i in (0, 1, 2, 250, 499, 750, 998, 999)},
        "sigmas_table_sum": float(ks.sigmas.sum()),
        "get_sigmas": {},
        "sigma_to_t": {},
    }
    for alg, n in (("karras", 10), ("karras", 20), ("karras", 30), ("linear", 10), ("linear", 20),
                   ("exp", 10), ("quad", 10), ("vp", 10)):
        s = ks.get_sigmas(alg, n, device="cpu")
        key = f"{alg}_{n}"
        out["get_sigmas"][key] = {"dtype": str(s.dtype), "values": [float(v) for v in s],
                                  "bits": [int(v) for v in s.view(torch.int32 if s.dtype == torch.float32 else torch.int64)]}
        sig = s[:-1]
        dists = torch.abs(sig.cpu() - ks.sigmas[:, None])
        low_idx, high_idx = torch.sort(torch.topk(dists, dim=0, k=2, largest=False).indices, dim=0)[0]
        t = ks.sigma_to_t(sig, device="cpu")
        out["sigma_to_t"][key] = {"low_idx": low_idx.tolist(), "high_idx": high_idx.tolist(),
                                  "t": [float(v) for v in t], "t_dtype": str(t.dtype)}
    # t_to_sigma at a few fractional points
    tt = torch.tensor([0.0, 0.5, 10.25, 499.75, 998.5, 999.0])
    out["t_to_sigma"] = {"t": tt.tolist(), "sigma": [float(v) for v in ks.t_to_sigma(tt, device="cpu")]}
    return out


def build_reference_sampler(cls_name, unet):
    """Construct the reference sampler (registry name) around a UNet module with repairs D1-D3."""
    import cpd.samplers as S
    import cpd.scheduler.discrete as D
    import cpd.scheduler.k as K
    import cpd.samplers.extension.denoiser as DN

    if not hasattr(D.SigmaScheduler, "append_zero"):  # D1
        D.SigmaScheduler.append_zero = lambda self, x: torch.cat([x, x.new_zeros([1])])
    if not getattr(D.SigmaScheduler, "_d2", False):  # D2
        orig = D.SigmaScheduler.get_sigmas

        def get_sigmas(self, algorithm, n, **kw):
            table = self.sigmas
            out = orig(self, algorithm, n, **kw)
            self.sigmas = table
            return out

        D.SigmaScheduler.get_sigmas = get_sigmas
        D.SigmaScheduler._d2 = True
    if not getattr(DN.Denoiser, "_d3", False):  # D3
        orig_pc = DN.Denoiser._process_conditioning

        def _pc(self, x, c, *args, **kwargs):
            kwargs.pop("sigma", None)
            return orig_pc(self, x, c, args[0], **kwargs)

        DN.Denoiser._process_conditioning = _pc
        DN.Denoiser._d3 = True
    model = {"unet": unet, "vae": None, "tokenizer": None, "clip_new_model": torch.nn.Module(), "decode": None}
    wrapper = S.make({"name": cls_name, "args": {}}, {"model": model})
    wrapper.sampler.denoiser.scheduler.sigmas = K.KScheduler().sigmas  # D2: the 1000-entry training table
    return wrapper


def make_case_inputs(cfg, hw, seed=1234):
    g = torch.Generator().manual_seed(seed)
    D = cfg.context_dim
    uc = torch.randn(1, 77, D, generator=g)
    embs = [torch.randn(1, 77, D, generator=g) for _ in range(3)]
    mask = torch.zeros(1, 1, hw, hw, dtype=torch.uint8)
    mask[..., : hw // 2] = 1  # "left half valid" style mask (prompts.py:807-818)
    c = {"and": [(1.0, embs[0], None, 1), (0.6, embs[1], None, mask)], "not": [(0.4, embs[2], None, 1)]}
    x_T = torch.randn(1, 4, hw, hw, generator=g)
    return uc, embs, mask, c, x_T


def reference_sampling(ref_shim):
    from oracle.unet import UNetConfig, make_weights

    cfg = UNetConfig.tiny()
    hw, steps = 8, 6
    unet = ref_shim.build_reference_unet(cfg)
    unet.load_state_dict(make_weights(cfg, seed=0), strict=True)
    unet.eval()
    uc, embs, mask, c, x_T = make_case_inputs(cfg, hw)
    out = {"x_T": x_T.numpy(), "uc": uc.numpy(), "embs": torch.cat(embs).numpy(), "mask": mask.numpy(),
           "scales": np.array([1.0, 0.6, 0.4]), "steps": np.array(steps), "hw": np.array(hw), "guidance": np.array(7.5)}
    return _run_cases(ref_shim, unet, out, c, uc, x_T, hw, steps,
                      (("Euler", "karras", "epsilon", {}), ("DPM++ 2m", "karras", "epsilon", {}),
                       ("Euler Ancestral", "karras", "epsilon", {}), ("Euler", "exp", "velocity", {}),
                       ("DPM++ 2m", "linear", "velocity", {})))


MORE_CASES = (("Huen", "karras", "epsilon", {}), ("DPM2", "karras", "epsilon", {}), ("DPM2 Ancestral", "karras", "epsilon", {}),
              ("DPM++ 2s Ancestral", "karras", "epsilon", {}), ("LMS", "karras", "epsilon", {}), ("Huen", "exp", "velocity", {}),
              ("Euler", "karras", "epsilon", {"scaled_clip": True, "scaled_clip_threshold": 97.0}),
              ("DPM++ 2m", "karras", "epsilon", {"scaled_clip": True, "scaled_clip_alg": "static_thresholding",
                                                 "scaled_clip_threshold": 0.5}))


def reference_sampling_more(ref_shim):
    """tests/golden/ref_sampling2.npz: the remaining k-diffusion samplers (two-stage / multistep) and the dynamic /
    static scale clip of the Denoiser, same inputs as ref_sampling.npz."""
    from oracle.unet import UNetConfig, make_weights

    cfg = UNetConfig.tiny()
    hw, steps = 8, 6
    unet = ref_shim.build_reference_unet(cfg)
    unet.load_state_dict(make_weights(cfg, seed=0), strict=True)
    unet.eval()
    uc, embs, mask, c, x_T = make_case_inputs(cfg, hw)
    return _run_cases(ref_shim, unet, {}, c, uc, x_T, hw, steps, MORE_CASES)


def _run_cases(ref_shim, unet, out, c, uc, x_T, hw, steps, cases):
    for name, sched, pred, extra in cases:
        # NB ("Euler", "linear") cannot be generated: get_sigmas_linear returns fp64 sigmas, to_ode's
        # append_dims (euler.py:103-111) turns the 0-dim sigma into a 4-D fp64 tensor, x is promoted to fp64
        # and the second UNet call raises "Input type (double) and bias type (float)" (defect D9).
        wrapper = build_reference_sampler(name, unet)
        dens, noises, u_x, u_t, u_out = [], [], [], [], []
        orig_forward = unet.forward

        def rec_forward(x, t, ctx, **k):
            r = orig_forward(x, t, ctx, **k)
            u_x.append(x.clone()); u_t.append(t.clone()); u_out.append(r[0].clone())
            return r

        unet.forward = rec_forward
        real_randn_like = torch.randn_like

        def rec_randn_like(x, *a, **k):
            n = real_randn_like(x, *a, **k)
            noises.append(n.clone())
            return n

        torch.randn_like = rec_randn_like
        torch.manual_seed(77)
        call_extra = dict(extra)
        if isinstance(call_extra.get("score_corrector"), tuple):  # (registered name, threshold_x, threshold_e): manager.py:84-90
            from cpd.samplers.extension.registry import create
            import cpd.samplers.extension.threshold  # noqa: F401
            nm, tx, te = call_extra["score_corrector"]
            call_extra["score_corrector"] = create(nm, threshold_x=tx, threshold_e=te)
        try:
            res = wrapper.sampler.sample(steps=steps, batch_size=1, shape=[4, hw, hw], x_T=x_T.clone(), conditioning=c,
                                         unconditional_conditioning=uc, unconditional_guidance_scale=7.5,
                                         scheduler=sched, device="cpu", silent=True, pred_type=pred,
                                         callback=lambda d: dens.append(d["eps"].clone()), **call_extra)
        finally:
            torch.randn_like = real_randn_like
            unet.forward = orig_forward
        key = f"{name}|{sched}|{pred}".replace(" ", "_") + ("|" + "|".join(f"{k}={v}" for k, v in extra.items()) if extra else "")
        out[key + "|final"] = res.numpy()
        out[key + "|denoised"] = torch.stack(dens).numpy()
        out[key + "|unet_x"] = torch.stack(u_x).numpy()
        out[key + "|unet_t"] = torch.stack(u_t).numpy()
        out[key + "|unet_out"] = torch.stack(u_out).numpy()
        if noises:
            out[key + "|noise"] = torch.stack(noises).numpy()
        print(key, "final std", float(res.std()), "n noise draws", len(noises))
    return out


CHURN_CASES = (("Euler", "karras", "epsilon", {"s_churn": 4.0, "s_noise": 1.003}), ("Huen", "karras", "epsilon", {"s_churn": 4.0}),
               ("DPM2", "karras", "epsilon", {"s_churn": 9.0, "s_tmin": 0.5, "s_tmax": 6.0, "s_noise": 0.99}))


def reference_sampling_churn(ref_shim):
    """tests/golden/ref_sampling3.npz: stochastic churn (s_churn > 0: gamma > 0, noise added before the denoiser call)."""
    from oracle.unet import UNetConfig, make_weights

    cfg = UNetConfig.tiny()
    hw, steps = 8, 6
    unet = ref_shim.build_reference_unet(cfg)
    unet.load_state_dict(make_weights(cfg, seed=0), strict=True)
    unet.eval()
    uc, embs, mask, c, x_T = make_case_inputs(cfg, hw)
    return _run_cases(ref_shim, unet, {}, c, uc, x_T, hw, steps, CHURN_CASES)


CORRECTOR_CASES = (("Euler", "karras", "epsilon", {"score_corrector": ("static_thresholding", 1.5, 0.9)}),
                   ("DPM++ 2m", "karras", "epsilon", {"score_corrector": ("dynamic_thresholding", 95.0, 97.0)}),
                   ("Euler Ancestral", "karras", "epsilon", {"score_corrector": ("renorm_thresholding", None, 96.0)}),
                   ("Huen", "karras", "epsilon", {"score_corrector": ("scaled_dynamic_perc_thresholding", 90.0, 95.0),
                                                  "scaled_clip": True, "scaled_clip_alg": "dynanormic_thresholding",
                                                  "scaled_clip_threshold": 99.0}))


def reference_sampling_corrector(ref_shim):
    """tests/golden/ref_sampling4.npz: the score_corrector hook (denoiser.py:517-518) with the registered thresholding
    extensions built like manager.py:84-90, and a non-clamp scaled_clip_alg."""
    from oracle.unet import UNetConfig, make_weights

    cfg = UNetConfig.tiny()
    hw, steps = 8, 6
    unet = ref_shim.build_reference_unet(cfg)
    unet.load_state_dict(make_weights(cfg, seed=0), strict=True)
    unet.eval()
    uc, embs, mask, c, x_T = make_case_inputs(cfg, hw)
    return _run_cases(ref_shim, unet, {}, c, uc, x_T, hw, steps, CORRECTOR_CASES)


IMG2IMG_CASES = (("Euler", "karras", "epsilon", {"decode": True, "denoising_strength": 0.6}),
                 ("DPM++ 2m", "karras", "epsilon", {"decode": True, "denoising_strength": 0.35}),
                 ("Euler Ancestral", "exp", "epsilon", {"decode": True, "denoising_strength": 1.0}))


def reference_sampling_img2img(ref_shim):
    """tests/golden/ref_sampling5.npz: the decode=True branch of KDiffusionSampler.sample (k_diffusion.py:64-70: truncated
    schedule, x = x_T + randn * sigmas[0] with the global CPU generator seeded by torch.manual_seed(77) right before)."""
    from oracle.unet import UNetConfig, make_weights

    cfg = UNetConfig.tiny()
    hw, steps = 8, 6
    unet = ref_shim.build_reference_unet(cfg)
    unet.load_state_dict(make_weights(cfg, seed=0), strict=True)
    unet.eval()
    uc, embs, mask, c, x_T = make_case_inputs(cfg, hw)
    return _run_cases(ref_shim, unet, {}, c, uc, x_T, hw, steps, IMG2IMG_CASES)


GUIDANCE_CASES = (
    ("Euler", "karras", "epsilon", {"unconditional_guidance_blur": True, "unconditional_guidance_blur_rounds": 4}),
    ("DPM++ 2m", "karras", "epsilon", {"attn_guide": True}),  # defaults: mode 2, rounds 4, threshold 90, blur k 31, scale 1.1
    ("Euler", "karras", "epsilon", {"attn_guide": True, "attn_guide_mode": 1, "attn_guide_blur_k": 7, "attn_guide_mask_threshold": 75,
                                    "attn_guide_scale": 1.3, "attn_guide_rounds": 3, "unconditional_guidance_blur": True,
                                    "unconditional_guidance_blur_k": 5, "unconditional_guidance_blur_rounds": 3}),
    ("DPM++ 2m", "karras", "velocity", {"depth_mask": True}),
)


def reference_sampling_guidance(ref_shim):
    """tests/golden/ref_sampling7.npz: the Denoiser's unconditional blur (denoiser.py:333-337,441-442), attention guidance
    (:341-350,404-435,461-462) and depth mask (:358-360,386-388) branches on the shimmed reference with the tiny UNet at 16 x 16
    (the 31-tap blur needs reflect padding of 15 < 16).  Every UNet call is recorded (input, timestep, output and, for
    return_attn calls, the skip tensor the saliency mask is taken from) so that the oracle can be replayed bit-exactly.  The
    blur's random sigma comes from the global torch RNG (torchvision GaussianBlur.get_params), seeded with 77 per case."""
    import dataclasses
    from oracle.unet import UNetConfig, make_weights

    hw, steps = 16, 6
    out = {}
    for name, sched, pred, extra in GUIDANCE_CASES:
        cfg = UNetConfig.tiny()
        depth = bool(extra.get("depth_mask"))
        if depth:
            cfg = dataclasses.replace(cfg, in_channels=5)
        unet = ref_shim.build_reference_unet(cfg)
        unet.load_state_dict(make_weights(cfg, seed=0), strict=True)
        unet.eval()
        uc, embs, mask, c, x_T = make_case_inputs(UNetConfig.tiny(), hw)
        if "x_T" not in out:
            out.update({"x_T": x_T.numpy(), "uc": uc.numpy(), "embs": torch.cat(embs).numpy(), "mask": mask.numpy(),
                        "scales": np.array([1.0, 0.6, 0.4]), "steps": np.array(steps), "hw": np.array(hw), "guidance": np.array(7.5)})
        wrapper = build_reference_sampler(name, unet)
        calls, dens = [], []
        orig_forward = unet.forward
        idx = extra.get("attn_guide_idx", -1)

        def rec_forward(x, t, ctx, **k):
            r = orig_forward(x, t, ctx, **k)
            o = r[0] if isinstance(r, tuple) else r
            # the saliency source is only ever used through its channel mean (denoiser.py:408): record that ([R, 1, h, w]); a
            # one-channel tensor replays it bit-exactly (the mean over one channel is the value itself)
            calls.append((x.clone(), t.clone().double(), o.clone(), r[1][idx].mean(1, keepdims=True) if isinstance(r, tuple) else None))
            return r

        unet.forward = rec_forward
        call_extra = dict(extra)
        if depth:
            g = torch.Generator().manual_seed(4321)
            call_extra["depth_mask"] = torch.rand(1, 1, hw, hw, generator=g)
            out["depth_mask"] = call_extra["depth_mask"].numpy()
        torch.manual_seed(77)
        try:
            res = wrapper.sampler.sample(steps=steps, batch_size=1, shape=[4, hw, hw], x_T=x_T.clone(), conditioning=c,
                                         unconditional_conditioning=uc, unconditional_guidance_scale=7.5, scheduler=sched,
                                         device="cpu", silent=True, pred_type=pred, callback=lambda d: dens.append(d["eps"].clone()),
                                         **call_extra)
        finally:
            unet.forward = orig_forward
        key = f"{name}|{sched}|{pred}".replace(" ", "_") + "|" + "|".join(f"{k}={v}" for k, v in extra.items())
        out[key + "|final"] = res.numpy()
        out[key + "|denoised"] = torch.stack(dens).numpy()
        out[key + "|n_calls"] = np.array(len(calls))
        for i, (x, t, o, sk) in enumerate(calls):
            out[f"{key}|call{i}|x"], out[f"{key}|call{i}|t"], out[f"{key}|call{i}|out"] = x.numpy(), t.numpy(), o.numpy()
            if sk is not None and extra.get("attn_guide"):
                out[f"{key}|call{i}|skip"] = sk.numpy()
        print(key, "final std", float(res.std()), "unet calls", len(calls))
    return out


def schedule_kats_discrete():
    """tests/golden/schedule_kat2.json: SigmaScheduler.get_sigmas_{karras, exponential, quad, vp, sigmoid}
    (cpd/scheduler/discrete.py:21-85) called unbound (they read only their kwargs), float32 bit patterns as ints."""
    import cpd.scheduler.discrete as D

    S = D.SigmaScheduler
    out = {}
    for n in (1, 2, 10, 20, 30):
        for alg, fn in (("karras", S.get_sigmas_karras), ("exp", S.get_sigmas_exponential), ("quad", S.get_sigmas_quad),
                        ("vp", S.get_sigmas_vp), ("sigmoid", S.get_sigmas_sigmoid)):
            for tag, kw in (("default", {}), ("wide", {"sigma_min": 0.03, "sigma_max": 14.6, "rho": 5.0})):
                try:
                    v = fn(None, n, device="cpu", **kw)
                except Exception as e:  # noqa: BLE001
                    out[f"{alg}|{n}|{tag}"] = {"error": type(e).__name__}
                    continue
                out[f"{alg}|{n}|{tag}"] = {"dtype": str(v.dtype), "bits": v.to(torch.float32).view(torch.int32).tolist()}
    return out


DECAY_CASES = (("Euler", "karras", "epsilon", {"decaying_uc_scale": True}),
               ("DPM++ 2m", "karras", "epsilon", {"decaying_uc_scale": True, "decaying_uc_scale_start": 0, "decaying_uc_scale_min": 3}),
               ("Huen", "exp", "epsilon", {"decaying_uc_scale": True, "decaying_uc_scale_start": 2, "decaying_uc_scale_min": 0.5}))


def reference_sampling_decay(ref_shim):
    """tests/golden/ref_sampling6.npz: decaying guidance scale (denoiser.py:477-494; t_idx from the sampler loops, total_steps
    = len(sigmas) from KDiffusionSampler.sample)."""
    from oracle.unet import UNetConfig, make_weights

    cfg = UNetConfig.tiny()
    hw, steps = 8, 6
    unet = ref_shim.build_reference_unet(cfg)
    unet.load_state_dict(make_weights(cfg, seed=0), strict=True)
    unet.eval()
    uc, embs, mask, c, x_T = make_case_inputs(cfg, hw)
    return _run_cases(ref_shim, unet, {}, c, uc, x_T, hw, steps, DECAY_CASES)


NOISE_CASES = (("iter", 41, 5), ("constant", 7, 5), ("c", 123456, 5), ("loop", 9, 5), ("l", 0, 3), ("random", 5, 5))


def reference_noise(ref_shim):
    """tests/golden/ref_noise.npz: cpd/noise.py NoiseGenerator - the seed property under every seed mode and sample()
    (noise.py:34-46,86-93), six draws each; "random" mode after random.seed(2024); plus sample(seed=explicit)."""
    import random
    import cpd.noise as N

    out = {"build_cycle_mod_5": np.array(N.build_cycle_mod(5)), "build_cycle_mod_3": np.array(N.build_cycle_mod(3))}
    for mode, seed, cyc in NOISE_CASES:
        random.seed(2024)
        ng = N.NoiseGenerator((1, 4, 4, 4), "cpu", seed=seed, seed_mode=mode, cycle_size=cyc)
        draws, seeds = [], []
        for _ in range(6):
            draws.append(ng.sample().numpy())
            seeds.append(ng.last_seed)
        out[f"{mode}|{seed}|{cyc}|draws"] = np.stack(draws)
        out[f"{mode}|{seed}|{cyc}|seeds"] = np.array(seeds)
    ng = N.NoiseGenerator((2, 4, 8, 8), "cpu", seed=3)
    out["explicit|99"] = ng.sample(seed=99).numpy()
    out["explicit|after"] = np.array([ng.last_seed])
    return out


def reference_unet_injection(ref_shim):
    """tests/golden/ref_unet_inject.npz: the shimmed reference UNetModel (tiny config, seeded weights) with return_attn /
    return_feat and with inject_attns / inject_feats and their *_stop indices (unet.py:774-813).  The injected tensors are
    scaled copies of the model's own skip / feature tensors, so a test can rebuild them from its own forward pass."""
    from oracle.unet import UNetConfig, make_weights

    cfg = UNetConfig.tiny()
    unet = ref_shim.build_reference_unet(cfg)
    unet.load_state_dict(make_weights(cfg, seed=0), strict=True)
    unet.eval()
    g = torch.Generator().manual_seed(31)
    x = torch.randn(2, 4, 8, 8, generator=g)
    t = torch.tensor([731.0, 12.5])
    ctx = torch.randn(2, 77, cfg.context_dim, generator=g)
    with torch.no_grad():
        out, skips, feats = unet(x, t, ctx, return_attn=True, return_feat=True)
        inj_a = [s_ * 0.5 for s_ in skips]
        inj_f = [skips[0] * 0.3] + [f_ * 0.7 for f_ in feats[:-1]]
        out_a = unet(x, t, ctx, inject_attns=inj_a, inject_attns_stop=5)
        out_f = unet(x, t, ctx, inject_feats=inj_f, inject_feats_stop=3)
        out_af, skips_af = unet(x, t, ctx, return_attn=True, inject_attns=inj_a, inject_attns_stop=12, inject_feats=inj_f, inject_feats_stop=7)
    print("unet injection: plain std", float(out.std()), "attn-injected", float(out_a.std()), "feat-injected", float(out_f.std()))
    return {"x": x.numpy(), "t": t.numpy(), "ctx": ctx.numpy(), "out": out.numpy(), "out_a": out_a.numpy(), "out_f": out_f.numpy(),
            "out_af": out_af.numpy(), "n_skips": np.array(len(skips)), "n_feats": np.array(len(feats)),
            "skip_shapes": np.array([list(s_.shape) for s_ in skips]), "feat_shapes": np.array([list(f_.shape) for f_ in feats]),
            "skip0": skips[0].numpy(), "skip_last": skips[-1].numpy(), "feat_last": feats[-1].numpy(),
            "returned_skip0_af": skips_af[0].numpy()}


def reference_vae(ref_shim):
    """tests/golden/ref_vae.npz: the shimmed reference first-stage decoder (tiny config, seeded weights) on a seeded latent."""
    from oracle.vae import VAEConfig, make_weights

    cfg = VAEConfig.tiny()
    sd = make_weights(cfg, seed=0)
    decode = ref_shim.build_reference_vae_decode(cfg, sd)
    z = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(5))
    img = decode(z)
    print("vae decode", tuple(img.shape), "std", float(img.std()))
    return {"z": z.numpy(), "image": img.numpy()}


THRESHOLD_CASES = (("static_thresholding", 0.8), ("dynamic_thresholding", 99.0), ("dynamic_thresholding", 85.5),
                   ("dynanormic_thresholding", 99.0), ("dynanormic_thresholding", 0.855),
                   ("scaled_dynamic_perc_thresholding", 99.0), ("scaled_dynamic_perc_thresholding", 85.5),
                   ("renorm_thresholding", 99.66), ("renorm_thresholding", 70.0),
                   ("scaled_norm_thresholding", 60.0), ("scaled_norm_thresholding", 20.0),
                   ("spatial_norm_thresholding", 1.5), ("spatial_norm_thresholding", 0.6),
                   ("scaled_spatial_norm_thresholding", 60.0), ("scaled_spatial_norm_thresholding", 25.0))


def threshold_inputs():
    g = torch.Generator().manual_seed(77)
    return [torch.randn(1, 4, 16, 16, generator=g) * 2.5, torch.randn(1, 4, 32, 24, generator=g) * 0.7 + 0.3]


def reference_thresholds(ref_shim):
    """tests/golden/ref_threshold.npz: every runnable registered extension of samplers/extension/threshold.py called the
    way denoiser.py:511-512 calls it (create(name)(x, threshold=t)) on seeded single-image inputs."""
    from cpd.samplers.extension.registry import create
    import cpd.samplers.extension.threshold  # noqa: F401  (registers the classes)

    out = {}
    for j, x in enumerate(threshold_inputs()):
        out[f"x{j}"] = x.numpy()
        for k, (name, thr) in enumerate(THRESHOLD_CASES):
            y = create(name)(x.clone(), threshold=thr)
            assert y.dtype == torch.float16
            out[f"y{j}_{k}"] = y.numpy()  # fp16
    try:
        create("norm_thresholding")(threshold_inputs()[0], threshold=50.0)
        raise SystemExit("norm_thresholding unexpectedly runs")
    except NameError as e:
        print("norm_thresholding:", e)
    return out


PROMPT_STRINGS = ("a cat:1.5 a dog:0.5 trees", "x: y:abc z", "plain prompt without weights", ":5 rest", "a:2", "a:2 ", "red:1e-1 blue:-0.5",
                  "one:1 two:2 three:3 tail:", "spaces  inside:0.25  double", "", "colon at end:", "a:b:c 1:2")
MASK_SIZES = ("half", "third", "quarter", "fourth", "fifrth", "sixth", "seventh", "eigth", "ninth", "tenth", "2", "3", "5", "7", "10")
MASK_DIRECTIONS = ("left", "l", "west", "right", "r", "top", "t", "north", "bottom", "bot", "south")
MASK_MINORITIES = ("valid", "v", "show", "hidden", "h", "hide")
MASK_SHAPES = ((512, 512), (512, 768), (768, 512), (1024, 1024), (256, 320))


def mask_styles():
    styles = [f"{d}_{sz}_{m}" for d in ("left", "r", "top", "bot") for sz in MASK_SIZES for m in ("valid", "hidden")]
    styles += [f"{d}_third_{m}" for d in MASK_DIRECTIONS for m in MASK_MINORITIES]  # every alias once
    return styles + ["left", "top_third", "right_quarter", "b"]


def reference_prompts(ref_shim):
    """tests/golden/ref_prompts.npz: WeightedPrompt._parse_prompt (prompts.py:546-589) and
    CompositionalPrompt._parse_mask_style (:737-856) called unbound on stub instances (they only read self.opt.H / W).
    Every mask of the reference is constant along one axis (checked here), so the fixture keeps its profile along the other."""
    import types
    import cpd.embeddings.prompts as P

    out = {"prompt_strings": np.array(list(PROMPT_STRINGS))}
    parsed = [P.WeightedPrompt._parse_prompt(None, t) for t in PROMPT_STRINGS]
    out["prompt_parsed"] = np.array(json.dumps(parsed))
    styles, cases, lines, axes = mask_styles(), [], [], []
    for (H, W) in MASK_SHAPES:
        stub = types.SimpleNamespace(opt=types.SimpleNamespace(H=H, W=W))
        for st in styles:
            m = P.CompositionalPrompt._parse_mask_style(stub, st)
            assert m.dtype == torch.uint8 and tuple(m.shape) == (1, H // 8, W // 8)
            horizontal = bool((m == m[:, :1, :]).all())  # constant along rows: the profile runs along x
            line = m[0, 0, :] if horizontal else m[0, :, 0]
            full = line.view(1, 1, -1).expand_as(m) if horizontal else line.view(1, -1, 1).expand_as(m)
            assert torch.equal(full, m)
            pad = np.full(128, 255, dtype=np.uint8)
            pad[:line.numel()] = line.numpy()
            cases.append(f"{H}x{W}|{st}")
            lines.append(pad)
            axes.append(2 if horizontal else 1)
    out["mask_cases"] = np.array(cases)
    out["mask_lines"] = np.stack(lines)
    out["mask_axes"] = np.array(axes, dtype=np.uint8)
    print("prompt strings", len(parsed), "mask cases", len(cases))
    return out


def main():
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import ref_shim

    ref_shim.install()
    import cpd.scheduler.k as K

    os.makedirs(GOLD, exist_ok=True)
    only = [f for f in sys.argv[1:] if f.endswith("-only")]
    want = lambda flag: not only or flag in only
    if not only:
        with open(os.path.join(GOLD, "schedule_kat.json"), "w") as f:
            json.dump(schedule_kats(K), f, indent=1)
        np.savez_compressed(os.path.join(GOLD, "ref_sampling.npz"), **reference_sampling(ref_shim))
    if want("--guidance-only"):
        np.savez_compressed(os.path.join(GOLD, "ref_sampling7.npz"), **reference_sampling_guidance(ref_shim))
    if want("--more-only"):
        np.savez_compressed(os.path.join(GOLD, "ref_sampling2.npz"), **reference_sampling_more(ref_shim))
    if want("--vae-only"):
        np.savez_compressed(os.path.join(GOLD, "ref_vae.npz"), **reference_vae(ref_shim))
    if want("--churn-only"):
        np.savez_compressed(os.path.join(GOLD, "ref_sampling3.npz"), **reference_sampling_churn(ref_shim))
    if want("--corrector-only"):
        np.savez_compressed(os.path.join(GOLD, "ref_sampling4.npz"), **reference_sampling_corrector(ref_shim))
    if want("--inject-only"):
        np.savez_compressed(os.path.join(GOLD, "ref_unet_inject.npz"), **reference_unet_injection(ref_shim))
    if want("--noise-only"):
        np.savez_compressed(os.path.join(GOLD, "ref_noise.npz"), **reference_noise(ref_shim))
    if want("--decay-only"):
        np.savez_compressed(os.path.join(GOLD, "ref_sampling6.npz"), **reference_sampling_decay(ref_shim))
    if want("--schedule2-only"):
        with open(os.path.join(GOLD, "schedule_kat2.json"), "w") as f:
            json.dump(schedule_kats_discrete(), f)
    if want("--img2img-only"):
        np.savez_compressed(os.path.join(GOLD, "ref_sampling5.npz"), **reference_sampling_img2img(ref_shim))
    if want("--prompts-only"):
        np.savez_compressed(os.path.join(GOLD, "ref_prompts.npz"), **reference_prompts(ref_shim))
    if want("--threshold-only"):
        np.savez_compressed(os.path.join(GOLD, "ref_threshold.npz"), **reference_thresholds(ref_shim))
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
