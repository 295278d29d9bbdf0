"""Two-stage and multistep k-diffusion samplers on the fused step kernel: "Huen" (Heun,
cpd/samplers/huen.py:11-58), "DPM2" and "DPM2 Ancestral" (cpd/samplers/dpm2.py:10-108), "DPM++ 2s Ancestral"
(cpd/samplers/dpmpp.py:58-113) and "LMS" (cpd/samplers/lms.py:12-64).

Every UNet evaluation is still followed by ONE fused kernel: `cpd_sampler_step` separates the UNet input (x), the sample
the update starts from (x_base) and the destination (x_out), so stage 1 writes x_2 next to x and stage 2 evaluates x_2
and updates x; Heun keeps d of stage 1, LMS a ring of the last four d tensors (d_out / d_prev).  Step scalars are the
fp32 0-dim tensor expressions of the reference, evaluated on the host.
"""
import torch

from .._lib import CPD_DPMPP_2M, CPD_EULER, CPD_EULER_ANCESTRAL, CPD_HEUN2, CPD_LMS
from .diffusion import DiffusionSamplerWrapper
from .euler import get_ancestral_step
from .k_diffusion import KDiffusionSampler
from .registry import register


def _noise(noise_sampler, x):
    n = noise_sampler(x) if noise_sampler is not None else torch.randn_like(x)
    return n.to(x.device, torch.float32).contiguous()


class HeunDiffusionSampler(KDiffusionSampler):
    """Algorithm 2 of Karras et al. (2022) with the Heun correction (huen.py:24-58), gamma = 0."""

    def __init__(self, model):
        super().__init__(model, "heun")

    @torch.no_grad()
    def _sampling(self, x, sigmas, model_args=None, **kwargs):
        model_args = {} if model_args is None else model_args
        callback = kwargs.get("callback", None)
        den, plan = self._begin(x, model_args, kwargs)
        den_out = torch.empty_like(x) if callback is not None else None
        x2, d1 = torch.empty_like(x), torch.empty_like(x)
        for i in range(len(sigmas) - 1):
            model_args["t_idx"] = i
            sigma_hat = self._churn(x, sigmas, i, kwargs)  # huen.py:39-43
            dt = sigmas[i + 1] - sigma_hat
            if sigmas[i + 1] == 0:  # Euler step (huen.py:48-50)
                x_before = x.clone() if callback is not None else None
                den.fused_step(x, sigma_hat, plan, dict(sampler=CPD_EULER, dt=float(dt), denoised_out=den_out), **model_args)
                self._callback(callback, x_before, i, sigmas[i], den_out, sigma_hat)
            else:
                # stage 1: d = to_ode(x), x_2 = x + d * dt (x itself is not touched)
                den.fused_step(x, sigma_hat, plan, dict(sampler=CPD_EULER, dt=float(dt), x_out=x2, d_out=d1, denoised_out=den_out),
                               **model_args)
                self._callback(callback, x, i, sigmas[i], den_out, sigma_hat)
                # stage 2: d_2 = to_ode(x_2, sigma_{i+1}); x = x + ((d + d_2) / 2) * dt
                den.fused_step(x2, sigmas[i + 1], plan, dict(sampler=CPD_HEUN2, dt=float(dt), x_base=x, x_out=x, d_prev=(d1,)),
                               **model_args)
        return x


def _sigma_mid(a, b):
    return ((a ** (1 / 3) + b ** (1 / 3)) / 2) ** 3  # dpm2.py:48,94 (cube-root midpoint)


class DPM2DiffusionSampler(KDiffusionSampler):
    """dpm2.py:24-56: second evaluation at the cube-root midpoint on every step."""

    def __init__(self, model):
        super().__init__(model, "dpm2")

    @torch.no_grad()
    def _sampling(self, x, sigmas, model_args=None, **kwargs):
        model_args = {} if model_args is None else model_args
        callback = kwargs.get("callback", None)
        den, plan = self._begin(x, model_args, kwargs)
        den_out = torch.empty_like(x) if callback is not None else None
        x2 = torch.empty_like(x)
        for i in range(len(sigmas) - 1):
            model_args["t_idx"] = i
            sigma_hat = self._churn(x, sigmas, i, kwargs)  # dpm2.py:39-43
            sigma_mid = _sigma_mid(sigma_hat, sigmas[i + 1])
            dt_1 = sigma_mid - sigma_hat
            dt_2 = sigmas[i + 1] - sigma_hat
            den.fused_step(x, sigma_hat, plan, dict(sampler=CPD_EULER, dt=float(dt_1), x_out=x2, denoised_out=den_out), **model_args)
            self._callback(callback, x, i, sigmas[i], den_out, sigma_hat)
            den.fused_step(x2, sigma_mid, plan, dict(sampler=CPD_EULER, dt=float(dt_2), x_base=x, x_out=x), **model_args)
        return x


class DPM2AncestralDiffusionSampler(KDiffusionSampler):
    """dpm2.py:74-108 (like the reference, t_idx is not advanced inside this loop)."""

    def __init__(self, model):
        super().__init__(model, "dpm2 ancestral")

    @torch.no_grad()
    def _sampling(self, x, sigmas, model_args=None, **kwargs):
        model_args = {} if model_args is None else model_args
        callback = kwargs.get("callback", None)
        noise_sampler = kwargs.get("noise_sampler", None)
        den, plan = self._begin(x, model_args, kwargs)
        den_out = torch.empty_like(x) if callback is not None else None
        x2 = torch.empty_like(x)
        for i in range(len(sigmas) - 1):
            sigma_down, sigma_up = get_ancestral_step(sigmas[i], sigmas[i + 1])
            sigma_mid = _sigma_mid(sigmas[i], sigma_down)
            dt_1 = sigma_mid - sigmas[i]
            dt_2 = sigma_down - sigmas[i]
            den.fused_step(x, sigmas[i], plan, dict(sampler=CPD_EULER, dt=float(dt_1), x_out=x2, denoised_out=den_out), **model_args)
            self._callback(callback, x, i, sigmas[i], den_out)
            noise = _noise(noise_sampler, x)
            den.fused_step(x2, sigma_mid, plan, dict(sampler=CPD_EULER_ANCESTRAL, dt=float(dt_2), sigma_up=float(sigma_up), noise=noise,
                                                     x_base=x, x_out=x), **model_args)
        return x


def get_ancestral_step_eta(sigma_from, sigma_to, eta=1.0):
    """dpmpp.py:115-122."""
    if not eta:
        return sigma_to, 0.0
    sigma_up = min(sigma_to, eta * (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
    sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
    return sigma_down, sigma_up


class DPMPlusPlus2sAncestralDiffusionSampler(KDiffusionSampler):
    """Ancestral sampling with DPM-Solver++(2S) second-order steps (dpmpp.py:70-113)."""

    def __init__(self, model):
        super().__init__(model, "dpmpp 2s ancestral")

    @torch.no_grad()
    def _sampling(self, x, sigmas, model_args=None, **kwargs):
        model_args = {} if model_args is None else model_args
        callback = kwargs.get("callback", None)
        noise_sampler = kwargs.get("noise_sampler", None)
        if kwargs.get("clip_sample", False):
            raise NotImplementedError("clip_sample inside DPM++ 2s Ancestral (applied BEFORE the update, dpmpp.py:92-93) is not built")
        eta, tmp = kwargs.get("eta", 1.0), kwargs.get("temperature", 1.0)
        den, plan = self._begin(x, model_args, kwargs)
        den_out = torch.empty_like(x) if callback is not None else None
        x2 = torch.empty_like(x)
        sigma_fn = lambda t: t.neg().exp()
        t_fn = lambda sigma: sigma.log().neg()
        for i in range(len(sigmas) - 1):
            sigma_down, sigma_up = get_ancestral_step_eta(sigmas[i], sigmas[i + 1], eta=eta)
            if sigma_down == 0:  # Euler step, then the (zero-scaled) noise addition of dpmpp.py:111
                dt = sigma_down - sigmas[i]
                x_before = x.clone() if callback is not None else None
                noise = _noise(noise_sampler, x)
                den.fused_step(x, sigmas[i], plan, dict(sampler=CPD_EULER_ANCESTRAL, dt=float(dt), sigma_up=float(sigma_up), noise=noise,
                                                        noise_mul=float(tmp), denoised_out=den_out), **model_args)
                self._callback(callback, x_before, i, sigmas[i], den_out)
                continue
            t, t_next = t_fn(sigmas[i]), t_fn(sigma_down)
            r = 1 / 2
            h = t_next - t
            s = t + r * h
            den.fused_step(x, sigmas[i], plan, dict(sampler=CPD_DPMPP_2M, dpm_ratio=float(sigma_fn(s) / sigma_fn(t)),
                                                    dpm_expm1=float((-h * r).expm1()), dpm_first=1, x_out=x2, denoised_out=den_out),
                           **model_args)
            self._callback(callback, x, i, sigmas[i], den_out)
            noise = _noise(noise_sampler, x)
            den.fused_step(x2, sigma_fn(s), plan, dict(sampler=CPD_DPMPP_2M, dpm_ratio=float(sigma_fn(t_next) / sigma_fn(t)),
                                                       dpm_expm1=float((-h).expm1()), dpm_first=1, x_base=x, x_out=x, noise=noise,
                                                       sigma_up=float(sigma_up), noise_mul=float(tmp)), **model_args)
        return x


def linear_multistep_coeff(order, t, i, j):
    """lms.py:54-64: integral of the j-th Lagrange basis polynomial over [t_i, t_{i+1}] (scipy quad, epsrel 1e-4)."""
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


class LMSDiffusionSampler(KDiffusionSampler):
    """Linear multistep sampler of order <= 4 (lms.py:28-52): the last four d tensors live in a ring in HBM."""

    def __init__(self, model):
        super().__init__(model, "lms")

    @torch.no_grad()
    def _sampling(self, x, sigmas, model_args=None, **kwargs):
        model_args = {} if model_args is None else model_args
        callback = kwargs.get("callback", None)
        order = kwargs.get("order", 4)
        if not 1 <= order <= 4:
            raise ValueError("LMS order must be in 1..4")
        den, plan = self._begin(x, model_args, kwargs)
        den_out = torch.empty_like(x) if callback is not None else None
        ring = [torch.empty_like(x) for _ in range(order)]
        sig_cpu = sigmas.cpu()
        for i in range(len(sigmas) - 1):
            model_args["t_idx"] = i
            cur_order = min(i + 1, order)
            coeffs = [linear_multistep_coeff(cur_order, sig_cpu, i, j) for j in range(cur_order)]
            prev = [ring[(i - k) % order] for k in range(1, cur_order)]  # d_{i-1}, d_{i-2}, ...
            x_before = x.clone() if callback is not None else None
            den.fused_step(x, sigmas[i], plan, dict(sampler=CPD_LMS, lms_coeff=coeffs, d_prev=prev, d_out=ring[i % order],
                                                    denoised_out=den_out), **model_args)
            self._callback(callback, x_before, i, sigmas[i], den_out)
        return x


def _wrapper(name, ctor):
    @register(name)
    class _W(DiffusionSamplerWrapper):
        def __init__(self, name, **kwargs):
            kwargs["constructor"] = ctor
            super().__init__(name, **kwargs)
    _W.__name__ = ctor.__name__.replace("DiffusionSampler", "SamplerWrapper")
    return _W


HeunSamplerWrapper = _wrapper("Huen", HeunDiffusionSampler)  # the reference registers the misspelt name (huen.py:11)
DPM2SamplerWrapper = _wrapper("DPM2", DPM2DiffusionSampler)
DPM2AncestralSamplerWrapper = _wrapper("DPM2 Ancestral", DPM2AncestralDiffusionSampler)
DPMPlusPlus2sAncestralSamplerWrapper = _wrapper("DPM++ 2s Ancestral", DPMPlusPlus2sAncestralDiffusionSampler)
LMSSamplerWrapper = _wrapper("LMS", LMSDiffusionSampler)
