#!/usr/bin/env python
"""Test helper (not collected by pytest): conv_out against torch with per-channel / per-region error breakdown (needs a GPU)."""
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from complex_prompt_diffusion_b200 import ops  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
for (n, h, w, cin) in [(1, 16, 16, 64), (2, 16, 16, 320), (1, 64, 64, 320)]:
    g = torch.Generator().manual_seed(1)
    a = torch.randn(n, h, w, cin, generator=g).to(torch.float16).cuda()
    wt = (torch.randn(4, cin, 3, 3, generator=g) / math.sqrt(9 * cin)).to(torch.bfloat16)
    b = torch.randn(4, generator=g)
    out = torch.full((n, 4, h, w), float("nan"), device="cuda")
    ops.conv_out(a, wt.permute(0, 2, 3, 1).contiguous().cuda(), b.cuda(), out, n=n, h=h, w=w, cin=cin, cout=4)
    ref = F.conv2d(a.float().permute(0, 3, 1, 2), wt.float().cuda(), b.cuda(), padding=1)
    torch.cuda.synchronize()
    err = (out - ref).abs()
    print((n, h, w, cin), "nan", int(torch.isnan(out).sum()), "rel", float((out - ref).norm() / ref.norm()),
          "per-channel max", [float(err[:, o].max()) for o in range(4)], "interior max", float(err[:, :, 1:-1, 1:-1].max()),
          "out[0,:,5,5]", out[0, :, 5, 5].tolist(), "ref", ref[0, :, 5, 5].tolist())
