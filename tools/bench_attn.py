#!/usr/bin/env python
"""Micro-benchmark of cpd_attention on the SD-1.5 shapes (16 rows): CUDA events, TFLOP/s by 4*B*H*Nq*Nk*d (real d)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SHAPES = [  # B, H, Nq, Nk, d
    (16, 8, 4096, 4096, 40), (16, 8, 1024, 1024, 80), (16, 8, 256, 256, 160), (16, 8, 4096, 77, 40), (16, 8, 1024, 77, 80),
    (8, 5, 9216, 9216, 64), (16, 8, 256, 77, 160), (8, 5, 9216, 77, 64), (16, 10, 1024, 77, 64),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", type=int, default=-1)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    from complex_prompt_diffusion_b200 import ops
    for i, (B, H, Nq, Nk, d) in enumerate(SHAPES):
        if a.only >= 0 and i != a.only:
            continue
        dpad = (d + 15) // 16 * 16
        nk_pad = (Nk + 15) // 16 * 16
        ip = H * dpad
        kvb = 4 if (Nk == 77 and B % 4 == 0) else B  # cross-attention: the context rows are shared by the images of a batch
        q = torch.zeros(B * Nq, H, dpad, device="cuda", dtype=torch.float16)
        k = torch.zeros(B * nk_pad, H, dpad, device="cuda", dtype=torch.float16)
        q[..., :d] = torch.randn(B * Nq, H, d, device="cuda")
        k[..., :d] = torch.randn(B * nk_pad, H, d, device="cuda")
        vt = torch.zeros(H, dpad, B * nk_pad, device="cuda", dtype=torch.float16)
        vt[:, :d] = torch.randn(H, d, B * nk_pad, device="cuda")
        o = torch.empty(B * Nq, ip, device="cuda", dtype=torch.float16)

        def run():
            ops.attention(q, k, vt, o, ldq=ip, ldk=ip, ldvt=B * nk_pad, ldo=ip, batch=B, heads=H, nq=Nq, nk=Nk, nk_pad=nk_pad, dpad=dpad,
                          scale=d ** -0.5, d_head=d, kv_batch=kvb)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / a.reps
        fl = 4.0 * B * H * Nq * Nk * d
        exps = B * H * Nq * nk_pad
        print(f"attn B{B} H{H} {Nq}x{Nk} d{d}: {us:9.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  {exps / us / 1e6:6.2f} Texp/s (MUFU peak ~4.3)")


if __name__ == "__main__":
    main()
