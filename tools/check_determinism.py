#!/usr/bin/env python
"""Run-to-run determinism and batch invariance of one UNet evaluation (tiny and sd15 configs)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def main():
    from complex_prompt_diffusion_b200.models.unet import UNetModel
    from complex_prompt_diffusion_b200.models import fixtures
    for name, hw in (("tiny", 32), ("sd15", 32)):
        cfg = fixtures.UNET_PRESETS[name]
        unet = UNetModel(fixtures.random_state_dict(fixtures.unet_param_shapes(cfg), seed=0), device="cuda", **fixtures.unet_kwargs(name))
        g = torch.Generator().manual_seed(1)
        x = torch.randn(2, 4, hw, hw, generator=g).cuda()
        ctx = torch.randn(4, 77, cfg["context_dim"], generator=g).cuda()
        unet.set_context(ctx)
        a = unet.forward_rows(x, 0.5, 500.0, 4).clone()
        b = unet.forward_rows(x, 0.5, 500.0, 4).clone()
        print(f"{name}: same call twice (graph replay): bit-identical={torch.equal(a, b)} rel={rel(a, b):.2e}")
        one = unet.forward_rows(x[:1].contiguous(), 0.5, 500.0, 4).clone()
        print(f"{name}: image 0 alone vs inside a batch of 2: bit-identical={torch.equal(one, a[:4])} rel={rel(one, a[:4]):.2e}")
        two = unet.forward_rows(x[1:].contiguous(), 0.5, 500.0, 4).clone()
        print(f"{name}: image 1 alone vs inside a batch of 2: bit-identical={torch.equal(two, a[4:])} rel={rel(two, a[4:]):.2e}")
        unet.use_cuda_graph = False
        c = unet.forward_rows(x, 0.5, 500.0, 4).clone()
        print(f"{name}: eager vs graph: bit-identical={torch.equal(c, a)} rel={rel(c, a):.2e}")


if __name__ == "__main__":
    main()
