#!/usr/bin/env python
"""Micro-benchmark of cpd_groupnorm / cpd_layernorm on the SD-1.5 shapes (16 rows): CUDA events around a CUDA graph of
launches (no host launch gaps), warm (same buffer, L2-resident when it fits) and cold (buffers rotated past the 126 MB L2).
GB/s = (one read + one write of the tensor) / time: the HBM-algorithmic traffic of the op."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

GN = [(16, 4096, 320), (16, 1024, 640), (16, 256, 1280), (16, 64, 1280), (16, 4096, 640), (16, 256, 2560), (16, 1024, 1920),
      (16, 4096, 960), (4, 4096, 320), (16, 1024, 960), (2, 9216, 320), (2, 2304, 640)]
LN = [(65536, 320), (16384, 640), (4096, 1280)]


def timed(fn_list, reps):
    """fn_list: closures launched round-robin; returns us per launch."""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for f in fn_list:
            f()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(reps):
                fn_list[i % len(fn_list)]()
        g.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        g.replay()
        e1.record(s)
        s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=24)
    ap.add_argument("--only", default="")
    ap.add_argument("--gn-index", type=int, default=-1, help="run only this entry of the GroupNorm shape list")
    a = ap.parse_args()
    from complex_prompt_diffusion_b200 import ops
    dev = "cuda"
    gn = GN if a.gn_index < 0 else [GN[a.gn_index]]
    for (n, hw, C) in gn if a.only in ("", "gn") else []:
        nbuf = max(2, int(400e6 // (n * hw * C * 2)) + 1)
        nbuf = min(nbuf, 24)
        xs = [torch.randn(n * hw, C, device=dev).to(torch.float16) for _ in range(nbuf)]
        out = torch.empty_like(xs[0])
        g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
        stats = torch.zeros(n * 64 * ops.GN_MAX_CHUNKS, device=dev, dtype=torch.float64)
        mk = lambda x: (lambda: ops.groupnorm(x, g, b, out, stats, n_img=n, hw=hw, c0=C))
        warm = timed([mk(xs[0])], a.reps)
        cold = timed([mk(x) for x in xs], a.reps)
        mb = 2 * n * hw * C * 2 / 1e6
        print(f"groupnorm n={n} hw={hw:5d} C={C:5d} ({mb:6.1f} MB r+w): warm {warm:7.1f} us {mb / warm * 1e3:7.0f} GB/s   cold {cold:7.1f} us {mb / cold * 1e3:7.0f} GB/s")
    for (rows, C) in LN if a.only in ("", "ln") else []:
        nbuf = min(24, max(2, int(400e6 // (rows * C * 2)) + 1))
        xs = [torch.randn(rows, C, device=dev).to(torch.float16) for _ in range(nbuf)]
        out = torch.empty_like(xs[0])
        g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
        mk = lambda x: (lambda: ops.layernorm(x, g, b, out, rows=rows, c=C))
        warm = timed([mk(xs[0])], a.reps)
        cold = timed([mk(x) for x in xs], a.reps)
        mb = 2 * rows * C * 2 / 1e6
        print(f"layernorm rows={rows:6d} C={C:5d} ({mb:6.1f} MB r+w): warm {warm:7.1f} us {mb / warm * 1e3:7.0f} GB/s   cold {cold:7.1f} us {mb / cold * 1e3:7.0f} GB/s")


if __name__ == "__main__":
    main()
