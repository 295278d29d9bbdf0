#!/usr/bin/env python
"""clock64 timeline of the cross-attention kernel (attention_umma5.cu), CTA 0, first 8 work items.
Softmax thread phases: 0 loop top, 1 S ready, 2 S in registers, 3 exp done, 4 P stored + arrive, 5 P V done, 6 O read (o_free),
7 stores issued.  MMA issuer phases per (item, tile): 1 P seen, 2 P V issued, 3 next S issued (0 unused)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from complex_prompt_diffusion_b200 import ops, _lib  # noqa: E402

B, H, Nq, Nk, d = [int(v) for v in (sys.argv[1:6] if len(sys.argv) >= 6 else (16, 8, 4096, 77, 40))]
dpad, nk_pad, kvb = (d + 15) // 16 * 16, (Nk + 15) // 16 * 16, 4
ip = H * dpad
q = torch.randn(B * Nq, H, dpad, device="cuda").to(torch.float16)
k = torch.randn(kvb * nk_pad, H, dpad, device="cuda").to(torch.float16)
vt = torch.randn(H, dpad, kvb * nk_pad, device="cuda").to(torch.float16)
o = torch.empty(B * Nq, ip, device="cuda", dtype=torch.float16)
buf = torch.zeros(192, dtype=torch.int64, device="cuda")
lib = _lib.load()
lib.cpd_debug_attention_cross_timeline.argtypes = [C.c_void_p]
lib.cpd_debug_attention_cross_timeline.restype = None


def run():
    ops.attention(q, k, vt, o, ldq=ip, ldk=ip, ldvt=kvb * nk_pad, ldo=ip, batch=B, heads=H, nq=Nq, nk=Nk, nk_pad=nk_pad, dpad=dpad,
                  scale=d ** -0.5, d_head=d, kv_batch=kvb)


for _ in range(3):
    run()
torch.cuda.synchronize()
lib.cpd_debug_attention_cross_timeline(C.c_void_p(buf.data_ptr()))
run()
torch.cuda.synchronize()
lib.cpd_debug_attention_cross_timeline(C.c_void_p(0))
v = buf.cpu().tolist()
t0 = min(x for x in v if x > 0)
for t in range(2):
    for i in range(8):
        row = v[(t * 8 + i) * 8:(t * 8 + i) * 8 + 8]
        if row[0]:
            print(f"softmax tile {t} item {i}: " + " ".join(f"{x - t0:7d}" for x in row))
for i in range(8):
    for t in range(2):
        row = v[128 + (i * 2 + t) * 4:128 + (i * 2 + t) * 4 + 4]
        if any(row):
            print(f"issuer item {i} tile {t}: " + " ".join(f"{(x - t0) if x else -1:7d}" for x in row))
