#!/usr/bin/env python
"""A/B of the folded-LayerNorm / transposed-tail GEMM epilogues against the plain launches of the same shapes (CUDA events,
operands rotated over several sets)."""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--variants", default="0")
    ap.add_argument("--profile", default="", choices=["", "qkv", "producer"],
                    help="launch only the fused Q | K | V projection (folded LayerNorm + transposed tail) or the residual GEMM that emits "
                         "the partial sums, 6 times, at the 64 x 64 shape: for `ncu -k regex:gemm2_kernel -s 4 -c 1`")
    a = ap.parse_args()
    from complex_prompt_diffusion_b200 import ops
    dev = "cuda"
    f16 = torch.float16
    if a.profile:
        M, K, N, nt = (65536, 320, 1152, 384) if a.profile == "qkv" else (65536, 384, 320, 0)
        x = torch.randn(M, K, device=dev).to(f16)
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(f16)
        bias, gvec = torch.randn(N, device=dev), torch.randn(N, device=dev)
        sums = torch.rand(10, M, 2, device=dev) + 1.0
        sums[:, :, 1] += 400.0
        out = torch.empty(M, N - nt, device=dev, dtype=f16)
        out_t = torch.empty(max(nt, 1), M, device=dev, dtype=f16)
        res = torch.randn(M, N, device=dev).to(f16)
        for _ in range(6):
            if a.profile == "qkv":
                ops.gemm_conv(x, w, out, n_img=1, h=1, w=M, c0=K, n_out=N, bias=bias, ldd=N - nt, d_t=out_t, dt_col0=N - nt, ln_sums=sums, ln_parts=4,
                              ln_g=gvec, variant=192)
            else:
                ops.gemm_conv(x, w, out, n_img=1, h=1, w=M, c0=K, n_out=N, bias=bias, residual=res, ld_res=N, ln_sums_out=sums, variant=160)
        torch.cuda.synchronize()
        return

    def timeit(fn):
        for j in range(3):
            fn(j)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for j in range(a.reps):
            fn(j)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / a.reps

    for (M, K, N, geglu, nt) in [(65536, 320, 384, False, 0), (65536, 320, 768, False, 0), (65536, 320, 1152, False, 384),
                                 (16384, 640, 1920, False, 640), (4096, 1280, 3840, False, 1280), (65536, 320, 2560, True, 0),
                                 (16384, 640, 5120, True, 0)]:
        sets = 3
        xs = [torch.randn(M, K, device=dev).to(f16) for _ in range(sets)]
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(f16)
        bias = torch.randn(N, device=dev)
        gvec = torch.randn(N, device=dev)
        parts = 4
        sums = torch.rand(parts, M, 2, device=dev) + 1.0
        sums[:, :, 1] += 400.0
        outs = [torch.empty(M, (N // 2 if geglu else N) - nt, device=dev, dtype=f16) for _ in range(sets)]
        outs_full = [torch.empty(M, N // 2 if geglu else N, device=dev, dtype=f16) for _ in range(sets)]
        outs_t = [torch.empty(max(nt, 1), M, device=dev, dtype=f16) for _ in range(sets)]
        for v in [int(x) for x in a.variants.split(",")]:
            kw = dict(n_img=1, h=1, w=M, c0=K, n_out=N, bias=bias)
            if geglu:
                kw.update(epilogue=ops.CPD_EPI_GEGLU, geglu_block=256)
            else:
                kw.update(variant=v)
            t_plain = timeit(lambda j: ops.gemm_conv(xs[j % sets], w, outs_full[j % sets], **kw))
            t_ln = timeit(lambda j: ops.gemm_conv(xs[j % sets], w, outs_full[j % sets], ln_sums=sums, ln_parts=parts, ln_g=gvec, **kw))
            line = f"M={M} N={N} K={K}{' geglu' if geglu else ''} v{v}: plain {t_plain:7.1f} us   folded-LN consumer {t_ln:7.1f} us"
            if nt:
                t_t = timeit(lambda j: ops.gemm_conv(xs[j % sets], w, outs[j % sets], ldd=N - nt, d_t=outs_t[j % sets], dt_col0=N - nt, **kw))
                t_both = timeit(lambda j: ops.gemm_conv(xs[j % sets], w, outs[j % sets], ldd=N - nt, d_t=outs_t[j % sets], dt_col0=N - nt,
                                                        ln_sums=sums, ln_parts=parts, ln_g=gvec, **kw))
                line += f"   transposed tail {t_t:7.1f} us   both {t_both:7.1f} us"
            print(line, flush=True)
            if geglu:
                break
    # producer: 65536 x 320 x 384 with residual, with / without the partial sums
    M, K, N = 65536, 384, 320
    x = torch.randn(M, K, device=dev).to(f16)
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(f16)
    res = torch.randn(M, N, device=dev).to(f16)
    out = torch.empty(M, N, device=dev, dtype=f16)
    bias = torch.randn(N, device=dev)
    sums = torch.empty(10, M, 2, device=dev)
    for v in [int(x) for x in a.variants.split(",")]:
        t0 = timeit(lambda j: ops.gemm_conv(x, w, out, n_img=1, h=1, w=M, c0=K, n_out=N, bias=bias, residual=res, ld_res=N, variant=v))
        t1 = timeit(lambda j: ops.gemm_conv(x, w, out, n_img=1, h=1, w=M, c0=K, n_out=N, bias=bias, residual=res, ld_res=N, variant=v, ln_sums_out=sums))
        print(f"producer M={M} N={N} K={K} v{v}: plain {t0:7.1f} us   with partial sums {t1:7.1f} us")


if __name__ == "__main__":
    main()
