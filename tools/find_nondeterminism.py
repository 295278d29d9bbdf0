#!/usr/bin/env python
"""Run one eager UNet evaluation twice and report the first kernel launches whose outputs differ between the runs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from complex_prompt_diffusion_b200 import ops
    from complex_prompt_diffusion_b200.models.unet import UNetModel
    from complex_prompt_diffusion_b200.models import fixtures
    name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    cfg = fixtures.UNET_PRESETS[name]
    unet = UNetModel(fixtures.random_state_dict(fixtures.unet_param_shapes(cfg), seed=0), device="cuda", use_cuda_graph=False, **fixtures.unet_kwargs(name))
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 4, 32, 32, generator=g).cuda()
    ctx = torch.randn(4, 77, cfg["context_dim"], generator=g).cuda()
    unet.set_context(ctx)
    log, state = [], {"pass": 0, "i": 0, "bad": 0}

    def wrap(fname, out_index):
        orig = getattr(ops, fname)

        def f(*a, **k):
            r = orig(*a, **k)
            out = a[out_index]
            torch.cuda.synchronize()
            if state["pass"] == 0:
                log.append(out.clone())
            else:
                ref = log[state["i"]]
                if not torch.equal(ref, out):
                    d = (ref.float() - out.float())
                    nbad = int((d != 0).sum())
                    if state["bad"] < 12:
                        print(f"launch {state['i']:4d} {fname:12s} out{tuple(out.shape)}: {nbad} / {out.numel()} elements differ, "
                              f"max abs {d.abs().max().item():.3e}, kwargs { {kk: vv for kk, vv in k.items() if isinstance(vv, (int, float))} }")
                    state["bad"] += 1
                state["i"] += 1
            return r
        setattr(ops, fname, f)
    wrap("gemm_conv", 2)
    wrap("attention", 3)
    wrap("groupnorm", 3)
    wrap("layernorm", 3)
    wrap("conv_in", 3)
    wrap("conv_out", 3)
    wrap("upsample2x", 1)
    for p in (0, 1):
        state["pass"], state["i"] = p, 0
        unet.forward_rows(x, 0.5, 500.0, 4)
    print(f"{name}: {len(log)} launches compared, {state['bad']} differ")


if __name__ == "__main__":
    main()
