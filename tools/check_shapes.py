#!/usr/bin/env python
"""Smoke-run the UNet at the other BASELINE.json shapes (finite outputs, no launch failures) and time one evaluation."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from complex_prompt_diffusion_b200.models.unet import UNetModel
    from complex_prompt_diffusion_b200.models import fixtures
    for name, hw, imgs, rows in (("sd21", 96, 1, 2), ("sd15", 128, 1, 2), ("sd15", 32, 1, 2), ("sd15", 64, 1, 4), ("sd21", 64, 2, 2)):
        cfg = fixtures.UNET_PRESETS[name]
        unet = UNetModel(fixtures.random_state_dict(fixtures.unet_param_shapes(cfg), seed=0), device="cuda", **fixtures.unet_kwargs(name))
        x = torch.randn(imgs, 4, hw, hw, device="cuda")
        unet.set_context(torch.randn(rows, 77, cfg["context_dim"], device="cuda"))
        for _ in range(2):
            out = unet.forward_rows(x, 0.5, 500.0, rows)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            out = unet.forward_rows(x, 0.5, 500.0, rows)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 3 * 1e3
        fl = fixtures.unet_flops(cfg, hw, hw) * imgs * rows
        print(f"{name} {hw}x{hw} {imgs} image(s) x {rows} rows: finite={bool(torch.isfinite(out).all())} {ms:.2f} ms/eval "
              f"{fl / ms / 1e9:.0f} TFLOP/s")
        del unet
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
