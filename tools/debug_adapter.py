#!/usr/bin/env python
"""Test helper (not collected): where the reference-signature adapter and the oracle diverge on a toy model (needs a GPU)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from complex_prompt_diffusion_b200.samplers.extension.denoiser import Denoiser  # noqa: E402
from oracle.denoiser import OracleDenoiser  # noqa: E402


class TinyModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        gg = torch.Generator().manual_seed(11)
        self.conv = torch.nn.Conv2d(4, 4, 3, padding=1)
        self.proj = torch.nn.Linear(32, 4)
        for p_ in self.parameters():
            p_.data = torch.randn(p_.shape, generator=gg) * 0.2

    def forward(self, x, timesteps, context, return_attn=False, **kw):
        out = self.conv(x) + self.proj(context.mean(1))[:, :, None, None] + 1e-3 * timesteps[:, None, None, None]
        return out, [out] * 12


g = torch.Generator().manual_seed(5)
uc = torch.randn(1, 77, 32, generator=g)
e = [torch.randn(1, 77, 32, generator=g) for _ in range(2)]
c = {"and": [(1.0, e[0], None, 1)], "not": [(0.5, e[1], None, 1)]}
x = torch.randn(1, 4, 8, 8, generator=g) * 5
cpu, gpu = TinyModel().eval(), TinyModel().eval().cuda()
for a, b in zip(cpu.parameters(), gpu.parameters()):
    assert torch.equal(a, b.cpu())
kw = dict(conditioning=c, unconditional_conditioning=uc, unconditional_guidance_scale=4.0)
sigma = torch.tensor([3.7])
od = OracleDenoiser(cpu, dtype=torch.float32)
od.trace = []
ref = od(x.clone(), sigma, **dict(kw))
den = Denoiser(gpu)
out = den.forward(x.clone().cuda(), sigma, **dict(kw))
print("denoised rel", float((out.cpu() - ref).norm() / ref.norm()))
tr = od.trace[-1]
plan = den.plan_conditioning(c, uc, (8, 8))
rows = den.unet_rows(x.cuda(), sigma, plan)
print("unet rows rel", float((rows.cpu().reshape(-1) - tr["unet_out"].reshape(-1)).norm() / tr["unet_out"].norm()), "t oracle", tr["t"].tolist())
print("eps rel", float((out.cpu() - ref).abs().max()), "ref max", float(ref.abs().max()))
