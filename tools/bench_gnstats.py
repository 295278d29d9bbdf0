#!/usr/bin/env python
"""A/B of the GroupNorm-statistics epilogue (cpd_gemm_params.gn_sums_out) on the 3x3 convolutions that would carry it, and of
cpd_groupnorm_apply against cpd_groupnorm on the tensors they write."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timeit(fn, reps=20):
    for j in range(3):
        fn(j)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(reps):
        fn(j)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    from complex_prompt_diffusion_b200 import ops
    dev, f16 = "cuda", torch.float16
    for (n, h, c) in [(16, 64, 320), (16, 32, 640), (16, 16, 1280)]:
        sets = 3
        xs = [torch.randn(n, h, h, c, device=dev).to(f16) for _ in range(sets)]
        w = (torch.randn(c, 9 * c, device=dev) / math.sqrt(9 * c)).to(f16)
        bias = torch.randn(c, device=dev)
        outs = [torch.empty(n * h * h, c, device=dev, dtype=f16) for _ in range(sets)]
        sums = torch.zeros(n, c, 2, dtype=torch.int64, device=dev)
        for v in (0, 160, 2160, 256, 128):
            if v > 1000 and c > 640:
                continue
            kw = dict(n_img=n, h=h, w=h, c0=c, n_out=c, ksize=3, bias=bias, variant=v)
            try:
                t0 = timeit(lambda j: ops.gemm_conv(xs[j % sets], w, outs[j % sets], **kw))
                t1 = timeit(lambda j: ops.gemm_conv(xs[j % sets], w, outs[j % sets], gn_sums_out=sums, **kw))
            except RuntimeError as e:
                print(f"conv {n}x{h}x{h}x{c} v{v}: {str(e)[:80]}")
                continue
            print(f"conv {n}x{h}x{h}x{c} v{v}: plain {t0:7.1f} us   with GroupNorm statistics {t1:7.1f} us", flush=True)
        gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        stats = torch.zeros(n * 64 * ops.GN_MAX_CHUNKS, dtype=torch.float64, device=dev)
        ys = [torch.empty_like(o) for o in outs]
        sums.zero_()
        ops.gemm_conv(xs[0], w, outs[0], gn_sums_out=sums, n_img=n, h=h, w=h, c0=c, n_out=c, ksize=3, bias=bias, variant=160)
        t_gn = timeit(lambda j: ops.groupnorm(outs[j % sets], gamma, beta, ys[j % sets], stats, n_img=n, hw=h * h, c0=c, silu=True))
        t_ap = timeit(lambda j: ops.groupnorm_apply(outs[j % sets], gamma, beta, ys[j % sets], sums, n_img=n, hw=h * h, c=c, silu=True))
        mb = 2 * n * h * h * c * 2 / 1e6
        print(f"groupnorm {n}x{h * h}x{c}: own statistics {t_gn:6.1f} us   apply only {t_ap:6.1f} us   ({mb:.0f} MB read + written)")


if __name__ == "__main__":
    main()
