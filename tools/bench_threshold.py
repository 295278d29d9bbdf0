#!/usr/bin/env python
"""Micro-benchmark of the thresholding kernels (cpd_threshold / cpd_threshold_ex): us per call at SD-1.5 / SDXL latent sizes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from complex_prompt_diffusion_b200.samplers.extension.denoiser import apply_threshold
    for (n, hw) in ((4, 64), (2, 128)):
        x0 = torch.randn(n, 4, hw, hw, device="cuda")
        bound = torch.empty(n, device="cuda")
        for name, thr in (("dynamic_thresholding", 99.5), ("renorm_thresholding", 99.5), ("scaled_norm_thresholding", 60.0),
                          ("spatial_norm_thresholding", 1.5), ("static_thresholding", 1.0)):
            xs = [x0.clone() for _ in range(13)]
            for x in xs[:3]:
                apply_threshold(x, bound, name, thr)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for x in xs[3:]:
                apply_threshold(x, bound, name, thr)
            e1.record()
            torch.cuda.synchronize()
            print(f"{name:34s} n={n} {hw}x{hw}: {e0.elapsed_time(e1) * 1e3 / 10:8.1f} us per call")


if __name__ == "__main__":
    main()
