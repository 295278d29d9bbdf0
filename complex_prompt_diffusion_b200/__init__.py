"""complex_prompt_diffusion_b200 - B200-native (sm_100a) drop-in for the denoising-loop hot path of
milesgray/complex_prompt_diffusion: sampler registry / wrappers / k-diffusion samplers, Denoiser,
sigma schedules, seeded noise and the UNet forward, with all device arithmetic in libcpd_b200.so.

Module layout mirrors the reference package `cpd` for the path only:
  samplers/ (registry, diffusion, k_diffusion, euler, dpmpp, extension/denoiser)  <- cpd/samplers
  scheduler/ (k, discrete)                                                       <- cpd/scheduler
  models/unet.py                                                                 <- cpd/models/unet.py (+attention, util)
  noise.py                                                                       <- cpd/noise.py
"""
from . import samplers, scheduler  # noqa: F401

__version__ = "0.1.0"
