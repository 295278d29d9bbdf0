#!/usr/bin/env python
"""Micro-benchmark of the small UNet kernels on the SD-1.5 shapes (16 rows): conv_out (320 -> 4, 64 x 64) and conv_in, CUDA events
around back-to-back launches on rotated buffers (inputs larger than L2 in total)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, n=20):
    for _ in range(3):
        fn(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


def main():
    from complex_prompt_diffusion_b200 import ops
    dev = "cuda"
    n, h, w, cin = 16, 64, 64, 320
    xs = [torch.randn(n, h, w, cin, device=dev).to(torch.float16) for _ in range(6)]
    wt = (torch.randn(4, 3, 3, cin, device=dev) / math.sqrt(9 * cin)).to(torch.bfloat16)
    b = torch.randn(4, device=dev)
    out = torch.empty(n, 4, h, w, device=dev)
    us = timed(lambda i: ops.conv_out(xs[i % 6], wt, b, out, n=n, h=h, w=w, cin=cin, cout=4))
    mb = n * h * w * cin * 2 / 1e6
    print(f"conv_out n={n} {h}x{w} cin={cin}: {us:7.1f} us  ({mb:.1f} MB read once = {mb / us * 1e3:.0f} GB/s; {2 * n * h * w * 9 * cin * 4 / us / 1e6:.1f} TFLOP/s)")
    x = torch.randn(1, 4, h, w, device=dev)
    wi = (torch.randn(320, 3, 3, 4, device=dev) / 6).to(torch.bfloat16)
    bi = torch.randn(320, device=dev)
    oi = torch.empty(n, h, w, 320, dtype=torch.float16, device=dev)
    us = timed(lambda i: ops.conv_in(x, wi, bi, oi, n=1, cin=4, h=h, w=w, cout=320, scale=0.5, rows_per_image=n))
    print(f"conv_in  1 image -> {n} rows {h}x{w} cout=320: {us:7.1f} us  ({mb:.1f} MB written = {mb / us * 1e3:.0f} GB/s)")


if __name__ == "__main__":
    main()
