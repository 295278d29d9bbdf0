"""Oracle (test infrastructure): the k-diffusion sampler loops and KDiffusionSampler.sample.

Follows cpd/samplers/k_diffusion.py:56-82, cpd/samplers/euler.py:24-57,71-105, cpd/samplers/dpmpp.py:23-56
(DPM++ 2M) and :70-113 (DPM++ 2S ancestral), cpd/samplers/huen.py:24-58, cpd/samplers/dpm2.py:24-108 and
cpd/samplers/lms.py:28-64 (paths relative to /root/reference).  Noise is INJECTED (``noise_sampler`` replaces
``torch.randn_like``) so the CUDA path and the oracle consume identical tensors (RNG row).

Sample thresholding (``clip_sample``, euler.py:55-56,93-94, dpmpp.py:51-52; threshold.py:47-88): repair D10 - the
reference's thresholding returns ``x.half()``, which silently turns x, ``s_in = x.new_ones(...)`` and therefore the sigma
handed to the Denoiser into fp16 for every later step.  Restated as what the extension computes (clamp to +-s, values
rounded through fp16) with x kept fp32.
"""
import numpy as np
import torch

from .denoiser import append_dims


def _minmax_scale(x):
    """threshold.py:131-133 (and :160-163, :220-223, :269-272): global min-max rescaling to [-1, 1]."""
    x_max, x_min = x.max(), x.min()
    y = (x - x_min) / (x_max - x_min)
    return 2 * y - 1., x_max, x_min


def _minmax_unscale(y, x_max, x_min):
    """threshold.py:144-145."""
    y = (y + 1) / 2
    return (x_max - x_min) * y + x_min


def threshold_apply(x, alg, threshold):
    """The registered thresholding extensions of threshold.py:7-286 on x: [B, C, h, w]; returns fp32 values rounded through
    fp16 (D10: the reference returns x.half()).  "norm_thresholding" (:182-205) reads an undefined x_max (NameError in the
    reference, D12) and is not restated.  Callers pass one image at a time (D7), so x.max() / x.min() / np.max over the
    per-image percentiles are per image."""
    x = x.float().clone()
    if alg == "none":  # :7-45
        return x
    if alg == "static_thresholding":  # :47-62
        torch.clamp_(x, -1 * threshold, threshold)
    elif alg == "dynamic_thresholding":  # :63-85
        s = np.percentile(np.abs(x.cpu()), threshold, axis=tuple(range(1, x.ndim)))
        s = np.max(np.append(s, 1.0))
        torch.clamp_(x, -1 * s, s)
    elif alg == "dynanormic_thresholding":  # :87-116
        q = threshold / 100 if 1 < threshold <= 100 else threshold
        s = torch.quantile(torch.abs(x).reshape((x.shape[0], -1)), q, dim=1)
        s = torch.maximum(s, torch.ones_like(s))[(...,) + (None,) * (x.ndim - 1)]
        x = torch.clamp(x, -s, s)
        x = x / s
    elif alg == "scaled_dynamic_perc_thresholding":  # :118-146
        x, x_max, x_min = _minmax_scale(x)
        s = np.percentile(np.abs(x.cpu()), threshold, axis=tuple(range(1, x.ndim)))
        s = np.max(np.append(s, 1.0))
        torch.clamp_(x, -1 * s, s)
        x = _minmax_unscale(x, x_max, x_min)
    elif alg == "renorm_thresholding":  # :148-180
        x, x_max, x_min = _minmax_scale(x)
        q = threshold / 100 if 1 < threshold <= 100 else threshold
        s = torch.quantile(x.flatten(1).abs(), q, dim=-1)
        s.clamp_(min=1.0)
        torch.clamp_(x, -1 * s, s)  # broadcasts [B] against the last axis in the reference: only valid for B = 1 (D7)
        x = _minmax_unscale(x, x_max, x_min)
    elif alg == "scaled_norm_thresholding":  # :207-237
        x, x_max, x_min = _minmax_scale(x)
        thr = (threshold / 100) * x_max
        s = x.pow(2).flatten(1).mean(1).sqrt().clamp(min=thr)
        x = x * (thr / s)
        x = _minmax_unscale(x, x_max, x_min)
    elif alg == "spatial_norm_thresholding":  # :239-254
        s = x.pow(2).mean(1, keepdim=True).sqrt().clamp(min=threshold)
        x = x * (threshold / s)
    elif alg == "scaled_spatial_norm_thresholding":  # :256-286
        x, x_max, x_min = _minmax_scale(x)
        thr = (threshold / 100) * x_max
        s = x.pow(2).mean(1, keepdim=True).sqrt().clamp(min=thr)
        x = x * (thr / s)
        x = _minmax_unscale(x, x_max, x_min)
    else:
        raise NotImplementedError(alg)
    return x.half().float()


class OracleScoreCorrector:
    """threshold.py:7-45 (ScoreCorrector.modify_score) over threshold_apply.  threshold_x thresholds x and drops the result
    (x is a clone, denoiser.py:524); threshold_e rewrites e_t and returns it as fp16 like the reference's `.half()`."""

    def __init__(self, alg, threshold_x=None, threshold_e=None):
        self.alg, self.threshold_x, self.threshold_e = alg, threshold_x, threshold_e

    def modify_score(self, e_t, x, t, c, **kwargs):
        if self.threshold_x:
            threshold_apply(x, self.alg, self.threshold_x)  # :20-23, result unused
        if self.threshold_e:
            e_t = threshold_apply(e_t, self.alg, self.threshold_e).half()  # :24-27
        return e_t


def _churn(x, sigmas, i, model_args, noise_sampler):
    """Stochastic churn of Karras et al. Algorithm 2 as in euler.py:40-45 / huen.py:38-43 / dpm2.py:38-43: one randn_like per
    step (drawn even when gamma = 0), sigma_hat = sigma * (gamma + 1), x += eps * sqrt(sigma_hat^2 - sigma^2) when gamma > 0."""
    s_churn, s_tmin = model_args.get("s_churn", 0.0), model_args.get("s_tmin", 0.0)
    s_tmax, s_noise = model_args.get("s_tmax", float("inf")), model_args.get("s_noise", 1.0)
    gamma = min(s_churn / (len(sigmas) - 1), 2 ** 0.5 - 1) if s_tmin <= sigmas[i] <= s_tmax else 0.0
    eps = None
    if noise_sampler is not None:
        eps = noise_sampler(x) * s_noise
    sigma_hat = sigmas[i] * (gamma + 1)
    if gamma > 0:
        x = x + eps * (sigma_hat ** 2 - sigmas[i] ** 2) ** 0.5
    return x, sigma_hat


def _clip(x, model_args):
    if model_args.get("clip_sample", False):
        return threshold_apply(x, model_args.get("clip_sample_alg", "dynamic_thresholding"), model_args.get("clip_sample_thresh", 90))
    return x


def to_ode(x, sigma, denoised):
    """euler.py:103-105."""
    return (x - denoised) / append_dims(sigma, x.ndim)


def get_ancestral_step(sigma_from, sigma_to):
    """euler.py:97-102."""
    sigma_up = (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5
    sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
    return sigma_down, sigma_up


@torch.no_grad()
def sample_euler(denoiser, x, sigmas, model_args, noise_sampler=None, callback=None):
    """euler.py:24-57 with s_churn = 0 (gamma = 0)."""
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        model_args["t_idx"] = i
        x, sigma_hat = _churn(x, sigmas, i, model_args, noise_sampler)  # euler.py:42-46
        den = denoiser(x, sigma_hat * s_in, **model_args)
        d = to_ode(x, sigma_hat, den)
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sigmas[i], "sigma_hat": sigma_hat, "eps": den})
        dt = sigmas[i + 1] - sigma_hat
        x = x + d * dt
        x = _clip(x, model_args)
    return x


@torch.no_grad()
def sample_euler_ancestral(denoiser, x, sigmas, model_args, noise_sampler, callback=None):
    """euler.py:71-95.  noise_sampler(x) stands for torch.randn_like(x) (one draw per step, after the model)."""
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        model_args["t_idx"] = i
        den = denoiser(x, sigmas[i] * s_in, **model_args)
        sigma_down, sigma_up = get_ancestral_step(sigmas[i], sigmas[i + 1])
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sigmas[i], "sigma_hat": sigmas[i], "eps": den})
        d = to_ode(x, sigmas[i], den)
        dt = sigma_down - sigmas[i]
        x = x + d * dt
        x = x + noise_sampler(x) * sigma_up
        x = _clip(x, model_args)
    return x


@torch.no_grad()
def sample_dpmpp_2m(denoiser, x, sigmas, model_args, callback=None):
    """dpmpp.py:23-56."""
    s_in = x.new_ones([x.shape[0]])
    sigma_fn = lambda t: t.neg().exp()
    t_fn = lambda sigma: sigma.log().neg()
    old_den = None
    for i in range(len(sigmas) - 1):
        model_args["t_idx"] = i
        den = denoiser(x, sigmas[i] * s_in, **model_args)
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sigmas[i], "sigma_hat": sigmas[i], "eps": den})
        t, t_next = t_fn(sigmas[i]), t_fn(sigmas[i + 1])
        h = t_next - t
        if old_den is None or sigmas[i + 1] == 0:
            x = (sigma_fn(t_next) / sigma_fn(t)) * x - (-h).expm1() * den
        else:
            h_last = t - t_fn(sigmas[i - 1])
            r = h_last / h
            den_d = (1 + 1 / (2 * r)) * den - (1 / (2 * r)) * old_den
            x = (sigma_fn(t_next) / sigma_fn(t)) * x - (-h).expm1() * den_d
        x = _clip(x, model_args)
        old_den = den
    return x


@torch.no_grad()
def sample_heun(denoiser, x, sigmas, model_args, noise_sampler=None, callback=None):
    """huen.py:24-58 ("Huen"), gamma = 0.  One randn_like per step is drawn and discarded (:40)."""
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        model_args["t_idx"] = i
        x, sigma_hat = _churn(x, sigmas, i, model_args, noise_sampler)  # huen.py:39-43
        den = denoiser(x, sigma_hat * s_in, **model_args)
        d = to_ode(x, sigma_hat, den)
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sigmas[i], "sigma_hat": sigma_hat, "eps": den})
        dt = sigmas[i + 1] - sigma_hat
        if sigmas[i + 1] == 0:
            x = x + d * dt
        else:
            x_2 = x + d * dt
            den_2 = denoiser(x_2, sigmas[i + 1] * s_in, **model_args)
            d_2 = to_ode(x_2, sigmas[i + 1], den_2)
            d_prime = (d + d_2) / 2
            x = x + d_prime * dt
    return x


@torch.no_grad()
def sample_dpm2(denoiser, x, sigmas, model_args, noise_sampler=None, callback=None):
    """dpm2.py:24-56: cube-root midpoint, second evaluation on EVERY step (no sigma_next == 0 special case)."""
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        model_args["t_idx"] = i
        x, sigma_hat = _churn(x, sigmas, i, model_args, noise_sampler)  # dpm2.py:39-43
        den = denoiser(x, sigma_hat * s_in, **model_args)
        d = to_ode(x, sigma_hat, den)
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sigmas[i], "sigma_hat": sigma_hat, "eps": den})
        sigma_mid = ((sigma_hat ** (1 / 3) + sigmas[i + 1] ** (1 / 3)) / 2) ** 3
        dt_1 = sigma_mid - sigma_hat
        dt_2 = sigmas[i + 1] - sigma_hat
        x_2 = x + d * dt_1
        den_2 = denoiser(x_2, sigma_mid * s_in, **model_args)
        d_2 = to_ode(x_2, sigma_mid, den_2)
        x = x + d_2 * dt_2
    return x


@torch.no_grad()
def sample_dpm2_ancestral(denoiser, x, sigmas, model_args, noise_sampler, callback=None):
    """dpm2.py:74-108 (t_idx is NOT updated in this loop, like the reference)."""
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        den = denoiser(x, sigmas[i] * s_in, **model_args)
        sigma_down, sigma_up = get_ancestral_step(sigmas[i], sigmas[i + 1])
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sigmas[i], "sigma_hat": sigmas[i], "eps": den})
        d = to_ode(x, sigmas[i], den)
        sigma_mid = ((sigmas[i] ** (1 / 3) + sigma_down ** (1 / 3)) / 2) ** 3
        dt_1 = sigma_mid - sigmas[i]
        dt_2 = sigma_down - sigmas[i]
        x_2 = x + d * dt_1
        den_2 = denoiser(x_2, sigma_mid * s_in, **model_args)
        d_2 = to_ode(x_2, sigma_mid, den_2)
        x = x + d_2 * dt_2
        x = x + noise_sampler(x) * sigma_up
    return x


def get_ancestral_step_eta(sigma_from, sigma_to, eta=1.0):
    """dpmpp.py:115-122."""
    if not eta:
        return sigma_to, 0.0
    sigma_up = min(sigma_to, eta * (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
    sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
    return sigma_down, sigma_up


@torch.no_grad()
def sample_dpmpp_2s_ancestral(denoiser, x, sigmas, model_args, noise_sampler, callback=None):
    """dpmpp.py:70-113 (eta = 1, temperature = 1 defaults; t_idx is not updated in this loop)."""
    eta, tmp = model_args.get("eta", 1.0), model_args.get("temperature", 1.0)
    s_in = x.new_ones([x.shape[0]])
    sigma_fn = lambda t: t.neg().exp()
    t_fn = lambda sigma: sigma.log().neg()
    for i in range(len(sigmas) - 1):
        den = denoiser(x, sigmas[i] * s_in, **model_args)
        sigma_down, sigma_up = get_ancestral_step_eta(sigmas[i], sigmas[i + 1], eta=eta)
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sigmas[i], "sigma_hat": sigmas[i], "eps": den})
        if sigma_down == 0:
            d = to_ode(x, sigmas[i], den)
            dt = sigma_down - sigmas[i]
            x = x + d * dt
        else:
            t, t_next = t_fn(sigmas[i]), t_fn(sigma_down)
            r = 1 / 2
            h = t_next - t
            s = t + r * h
            x_2 = (sigma_fn(s) / sigma_fn(t)) * x - (-h * r).expm1() * den
            den_2 = denoiser(x_2, sigma_fn(s) * s_in, **model_args)
            x = (sigma_fn(t_next) / sigma_fn(t)) * x - (-h).expm1() * den_2
        x = x + noise_sampler(x) * tmp * sigma_up
    return x


def linear_multistep_coeff(order, t, i, j):
    """lms.py:54-64 (scipy.integrate.quad, epsrel 1e-4)."""
    from scipy import integrate
    if order - 1 > i:
        raise ValueError(f"Order {order} too high for step {i}")

    def fn(tau):
        prod = 1.0
        for k in range(order):
            if j == k:
                continue
            prod *= (tau - t[i - k]) / (t[i - j] - t[i - k])
        return prod
    return integrate.quad(fn, t[i], t[i + 1], epsrel=1e-4)[0]


@torch.no_grad()
def sample_lms(denoiser, x, sigmas, model_args, noise_sampler=None, callback=None):
    """lms.py:28-52, order 4."""
    order = model_args.get("order", 4)
    s_in = x.new_ones([x.shape[0]])
    ds = []
    for i in range(len(sigmas) - 1):
        model_args["t_idx"] = i
        den = denoiser(x, sigmas[i] * s_in, **model_args)
        d = to_ode(x, sigmas[i], den)
        ds.append(d)
        if len(ds) > order:
            ds.pop(0)
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sigmas[i], "sigma_hat": sigmas[i], "eps": den})
        cur_order = min(i + 1, order)
        coeffs = [linear_multistep_coeff(cur_order, sigmas.cpu(), i, j) for j in range(cur_order)]
        x = x + sum(coeff * d for coeff, d in zip(coeffs, reversed(ds)))
    return x


SAMPLERS = {"Euler": sample_euler, "Euler Ancestral": sample_euler_ancestral, "DPM++ 2m": sample_dpmpp_2m,
            "Huen": sample_heun, "DPM2": sample_dpm2, "DPM2 Ancestral": sample_dpm2_ancestral,
            "DPM++ 2s Ancestral": sample_dpmpp_2s_ancestral, "LMS": sample_lms}


@torch.no_grad()
def sample(denoiser, name, steps, x_T, noise_sampler=None, callback=None, **kwargs):
    """KDiffusionSampler.sample, k_diffusion.py:56-82 (decode=False branch), for ONE image x_T:[1,4,h,w].

    kwargs are the reference's (conditioning, unconditional_conditioning, unconditional_guidance_scale,
    scheduler, sigma_min/max/rho, pred_type ...).  The same dict is model_args and **kwargs and is mutated
    with t_idx each step (k_diffusion.py:78-81, euler.py:41).
    """
    scheduler = kwargs.get("scheduler", "default")
    sigmas = denoiser.scheduler.get_sigmas(scheduler, steps, **kwargs)
    if kwargs.get("decode", False):  # img2img branch, k_diffusion.py:64-70 (the frame-to-frame step of cpd/animation.py:171-176)
        denoising_strength = kwargs.get("denoising_strength", 0.0)
        t_enc = int((1 - min(denoising_strength, 0.999)) * steps)
        sigmas = sigmas[steps - t_enc - 1:]
        noise = torch.randn(list(x_T.shape))
        noise = noise * sigmas[0]
        x = x_T + noise
    else:
        x = x_T * sigmas[0]  # k_diffusion.py:74
    kwargs["total_steps"] = len(sigmas)  # :76
    fn = SAMPLERS[name]
    if name == "DPM++ 2m":
        return fn(denoiser, x, sigmas, kwargs, callback=callback)
    return fn(denoiser, x, sigmas, kwargs, noise_sampler, callback=callback)


class OracleNoiseGenerator:
    """cpd/noise.py:12-46,86-93 restricted to seed modes iter/const (first 'iter' draw uses seed+1)."""

    def __init__(self, shape, device="cpu", seed=0, seed_mode="iter"):
        self.shape, self.device, self._seed, self.seed_mode = shape, device, seed, seed_mode

    @property
    def seed(self):
        if self.seed_mode == "iter":
            self._seed += 1
        return self._seed

    def sample(self, seed=None):
        if seed is None:
            seed = self.seed
        torch.manual_seed(seed)
        return torch.randn(self.shape, device=self.device)
