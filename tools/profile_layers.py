#!/usr/bin/env python
"""Per-layer CUDA-event timing of one UNet evaluation batch (default: SD-1.5, 64x64 latent, 16 rows = config 2's
per-step batch).  Prints every distinct (kind, shape) with launches, total ms, TFLOP/s or GB/s-relevant info."""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="sd15")
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--images", type=int, default=4)
    ap.add_argument("--rows", type=int, default=4)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--ncu-replay", action="store_true",
                    help="run under `ncu --profile-from-start off`: profile exactly ONE CUDA-graph replay of the evaluation")
    a = ap.parse_args()
    from complex_prompt_diffusion_b200 import ops
    from complex_prompt_diffusion_b200.models.unet import UNetModel
    from complex_prompt_diffusion_b200.models import fixtures
    cfg = fixtures.UNET_PRESETS[a.model]
    unet = UNetModel(fixtures.random_state_dict(fixtures.unet_param_shapes(cfg), seed=0), device="cuda", **fixtures.unet_kwargs(a.model))
    x = torch.randn(a.images, 4, a.latent, a.latent, device="cuda")
    ctx = torch.randn(a.rows, 77, cfg["context_dim"], device="cuda")
    unet.set_context(ctx)
    for _ in range(2):
        unet.forward_rows(x, 0.5, 500.0, a.rows)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        unet.forward_rows(x, 0.5, 500.0, a.rows)
    e1.record()
    torch.cuda.synchronize()
    print(f"whole forward ({a.images * a.rows} rows): {e0.elapsed_time(e1) / a.reps:.3f} ms")
    if a.ncu_replay:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        unet.forward_rows(x, 0.5, 500.0, a.rows)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    agg = collections.OrderedDict()
    unet.profile(True)  # plan-level records: eager launches bracketed by CUDA events inside cpd_unet_forward
    for _ in range(a.reps):
        unet.forward_rows(x, 0.5, 500.0, a.rows)
        torch.cuda.synchronize()
    for kind, label, us, fl in unet.profile_records():
        d = agg.setdefault((kind, label), [0, 0.0, 0.0])
        d[0] += 1
        d[1] += us / 1e3
        d[2] += fl
    unet.profile(False)
    tot = sum(v[1] for v in agg.values()) / a.reps
    print(f"sum of per-launch event times: {tot:.3f} ms")
    kinds = collections.defaultdict(float)
    for (kind, label), (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        kinds[kind] += ms / a.reps
        tf = fl / ms / 1e9 if ms > 0 and fl > 0 else 0.0
        print(f"{kind:12s} {label:40s} x{n // a.reps:3d} {ms / a.reps:8.3f} ms {100 * ms / a.reps / tot:5.1f}%  {tf:7.1f} TFLOP/s")
    print({k: round(v, 3) for k, v in kinds.items()})


if __name__ == "__main__":
    main()
