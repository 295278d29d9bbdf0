"""Oracle (test infrastructure): Euler / Euler-ancestral / DPM++ 2M loops and KDiffusionSampler.sample.

Follows cpd/samplers/k_diffusion.py:56-82, cpd/samplers/euler.py:24-57,71-105 and
cpd/samplers/dpmpp.py:23-56 (paths relative to /root/reference).  Noise is INJECTED
(``noise_sampler`` replaces ``torch.randn_like``) so the CUDA path and the oracle consume identical
tensors (RNG row).
"""
import torch

from .denoiser import append_dims


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
        if noise_sampler is not None:
            noise_sampler(x)  # euler.py:43 draws (and discards, gamma = 0) one randn_like per step
        sigma_hat = sigmas[i] * 1.0
        den = denoiser(x, sigma_hat * s_in, **model_args)
        d = to_ode(x, sigma_hat, den)
        if callback is not None:
            callback({"x": x, "i": i, "sigma": sigmas[i], "sigma_hat": sigma_hat, "eps": den})
        dt = sigmas[i + 1] - sigma_hat
        x = x + d * dt
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
        old_den = den
    return x


SAMPLERS = {"Euler": sample_euler, "Euler Ancestral": sample_euler_ancestral, "DPM++ 2m": sample_dpmpp_2m}


@torch.no_grad()
def sample(denoiser, name, steps, x_T, noise_sampler=None, callback=None, **kwargs):
    """KDiffusionSampler.sample, k_diffusion.py:56-82 (decode=False branch), for ONE image x_T:[1,4,h,w].

    kwargs are the reference's (conditioning, unconditional_conditioning, unconditional_guidance_scale,
    scheduler, sigma_min/max/rho, pred_type ...).  The same dict is model_args and **kwargs and is mutated
    with t_idx each step (k_diffusion.py:78-81, euler.py:41).
    """
    scheduler = kwargs.get("scheduler", "default")
    sigmas = denoiser.scheduler.get_sigmas(scheduler, steps, **kwargs)
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
