#!/usr/bin/env python
"""Micro-benchmark of cpd_gemm_conv on the UNet's dominant shapes (CUDA events, L2 flushed between launches by
cycling through several distinct operand sets larger than L2 in total)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SHAPES = [  # (name, n_img, h, w, cin, cout, ksize)
    ("conv3 64^2 320->320", 16, 64, 64, 320, 320, 3),
    ("conv3 32^2 640->640", 16, 32, 32, 640, 640, 3),
    ("conv3 16^2 1280->1280", 16, 16, 16, 1280, 1280, 3),
    ("conv3 8^2 1280->1280", 16, 8, 8, 1280, 1280, 3),
    ("lin 65536x320x320", 1, 1, 65536, 320, 320, 1),
    ("lin 16384x640x640", 1, 1, 16384, 640, 640, 1),
    ("lin 4096x1280x1280", 1, 1, 4096, 1280, 1280, 1),
    ("lin 65536x320x1280", 1, 1, 65536, 1280, 320, 1),
    ("lin 8192x8192x8192", 1, 1, 8192, 8192, 8192, 1),
    ("geglu 65536x320->2560", 1, 1, 65536, 320, 2560, -1),  # ksize -1: the GEGLU epilogue (ff.net.0, attention.py:92-100)
    ("lin 65536x320x768 (q|k)", 1, 1, 65536, 320, 768, 1),
    # small-M levels (config 3: two rows of SD-2.1 at 96 x 96; strong-scaled config 2)
    ("conv3 2x12^2 1280->1280", 2, 12, 12, 1280, 1280, 3),
    ("conv3 2x12^2 2560->1280", 2, 12, 12, 2560, 1280, 3),
    ("conv3 2x24^2 1280->1280", 2, 24, 24, 1280, 1280, 3),
    ("lin 18432x320x320", 1, 1, 18432, 320, 320, 1),
    ("conv3 2x96^2 320->320", 2, 96, 96, 320, 320, 3),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--only", type=int, default=-1)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    from complex_prompt_diffusion_b200 import ops
    for i, (name, n, h, w, cin, cout, ks) in enumerate(SHAPES):
        if a.only >= 0 and i != a.only:
            continue
        geglu = ks < 0
        ks = abs(ks)
        M = n * h * w
        sets = max(2, min(8, int(200e6 // (M * (cin + cout) * 2 + cout * cin * ks * ks * 2)) + 1))
        xs = [torch.randn(M, cin, device="cuda").to(torch.float16) for _ in range(sets)]
        ws = [(torch.randn(cout, ks * ks * cin, device="cuda") / (ks * ks * cin) ** 0.5).to(torch.float16) for _ in range(sets)]
        outs = [torch.empty(M, cout // 2 if geglu else cout, device="cuda", dtype=torch.float16) for _ in range(sets)]
        bias = torch.randn(cout, device="cuda")

        def run(j):
            if geglu:
                ops.gemm_conv(xs[j % sets], ws[j % sets], outs[j % sets], n_img=n, h=h, w=w, c0=cin, n_out=cout, bias=bias,
                              epilogue=ops.CPD_EPI_GEGLU, geglu_block=256)
            else:
                ops.gemm_conv(xs[j % sets], ws[j % sets], outs[j % sets], n_img=n, h=h, w=w, c0=cin, n_out=cout, ksize=ks, bias=bias,
                              variant=a.variant)
        for j in range(3):
            run(j)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for j in range(a.reps):
            run(j)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / a.reps
        fl = 2.0 * M * cout * cin * ks * ks
        print(f"{name:28s} M={M:6d} N={cout:5d} K={cin * ks * ks:6d}  {us:9.1f} us  {fl / us / 1e6:8.1f} TFLOP/s  (variant {a.variant})")


if __name__ == "__main__":
    main()
