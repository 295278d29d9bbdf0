#!/usr/bin/env python
"""First-stage decoder (VAE decode) at full size: SD config, B latents of 64x64 -> 512x512 images; CUDA events."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    from complex_prompt_diffusion_b200 import ops
    from complex_prompt_diffusion_b200.models.vae import VAEDecoder
    from complex_prompt_diffusion_b200.models import fixtures
    cfg = fixtures.VAE_PRESETS["sd"]
    dec = VAEDecoder(fixtures.random_state_dict(fixtures.vae_param_shapes(cfg), seed=0), device="cuda", **cfg)
    z = torch.randn(a.batch, 4, a.latent, a.latent, device="cuda")
    for _ in range(2):
        out = dec.decode(z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        out = dec.decode(z)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    fl = fixtures.vae_flops(cfg, a.latent, a.latent) * a.batch
    print(f"vae decode B={a.batch} {a.latent}x{a.latent} -> {tuple(out.shape)}: {ms:.2f} ms  {fl / ms / 1e9:.1f} TFLOP/s  "
          f"{a.batch / ms * 1e3:.1f} images/s  finite={bool(torch.isfinite(out).all())}")
    ops.PROFILE = []
    dec.decode(z)
    torch.cuda.synchronize()
    agg = {}
    for (k_, e_a, e_b, f, _l) in ops.PROFILE:
        agg[k_] = agg.get(k_, 0.0) + e_a.elapsed_time(e_b)
    ops.PROFILE = None
    print("per-op-kind ms:", {k: round(v, 2) for k, v in agg.items()})


if __name__ == "__main__":
    main()
