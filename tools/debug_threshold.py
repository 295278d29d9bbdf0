#!/usr/bin/env python
"""Test helper (not collected by pytest): per-algorithm mismatch counts of cpd_threshold / cpd_threshold_ex against the oracle (needs a GPU)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from complex_prompt_diffusion_b200.samplers.extension.denoiser import apply_threshold  # noqa: E402
from oracle.make_golden import THRESHOLD_CASES  # noqa: E402
from oracle.samplers import threshold_apply  # noqa: E402

for shape in [(3, 4, 64, 64), (2, 4, 128, 128), (5, 4, 24, 40)]:
    g = torch.Generator().manual_seed(shape[2])
    x = torch.randn(*shape, generator=g) * torch.tensor([0.5, 2.0, 1.0, 3.0, 0.8])[:shape[0]].view(-1, 1, 1, 1) + 0.1
    for name, thr in THRESHOLD_CASES:
        xd = x.cuda().clone()
        bound = torch.zeros(shape[0], device="cuda")
        apply_threshold(xd, bound, name, thr)
        ref = torch.cat([threshold_apply(x[b:b + 1], name, thr) for b in range(shape[0])])
        d = (xd.cpu() - ref)
        bad = (d != 0)
        msg = ""
        if bad.any():
            idx = bad.nonzero()[0].tolist()
            msg = f" first at {idx}: dev {float(xd.cpu()[tuple(idx)])!r} ref {float(ref[tuple(idx)])!r} x {float(x[tuple(idx)])!r}"
        print(f"{shape} {name:36s} {thr:7.3f} mismatches {int(bad.sum()):6d} per image {bad.flatten(1).sum(1).tolist()} bound {bound.cpu().tolist()}{msg}")
print("torch threads", torch.get_num_threads())
