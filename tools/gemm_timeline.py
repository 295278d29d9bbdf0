#!/usr/bin/env python
"""Phase timeline of one gemm2_kernel launch.  Needs the TIMELINE=1 build that lives beside the product library:
    make -C complex_prompt_diffusion_b200/csrc TIMELINE=1 BUILD=build_tl OUT=../libcpd_b200_tl.so
    CPD_B200_LIB=complex_prompt_diffusion_b200/libcpd_b200_tl.so python tools/gemm_timeline.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
NAMES = ["entry", "cluster sync 1", "setup done (alloc, barriers, sync 2)", "first TMA issued", "first stage landed", "last MMA committed",
         "epilogue: accumulator ready", "first TMA store issued", "stores read", "cluster sync 3", "dealloc"]


def main():
    from complex_prompt_diffusion_b200 import ops
    lib = ops.load()
    for nm in ("cpd_debug_gemm_timeline", "cpd_debug_gemm_mma", "cpd_debug_gemm_epilogue"):
        getattr(lib, nm).restype = C.c_int
    from complex_prompt_diffusion_b200._lib import CPD_EPI_GEGLU
    # (M, N, K, variant (< 0: GEGLU), residual) linears, or - with --conv n,h,w,cin,cout,ksize,variant[;...] - convolutions
    cases = [(1, 1, M, K, N, 1, v, res) for (M, N, K, v, res) in
             [(256, 160, 64, 160, False), (65536, 320, 320, 2160, True), (65536, 320, 320, 160, True), (65536, 768, 320, 2128, False),
              (16384, 640, 640, 160, True), (65536, 2560, 320, -1, False), (65536, 320, 1280, 2160, True)]]
    if "--conv" in sys.argv:
        cases = [tuple(int(x) for x in c.split(",")) + (False,) for c in sys.argv[sys.argv.index("--conv") + 1].split(";")]
    for (n_img, hh, ww, cin, N, ks, v, res) in cases:
        M, K = n_img * hh * ww, ks * ks * cin
        a = torch.randn(M, cin, device="cuda").half()
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
        geglu = v < 0
        o = torch.empty(M, N // 2 if geglu else N, device="cuda", dtype=torch.float16)
        r = torch.randn(M, N, device="cuda").half() if res else None
        bias = torch.randn(N, device="cuda")

        def launch():
            if geglu:
                ops.gemm_conv(a, w, o, n_img=n_img, h=hh, w=ww, c0=cin, n_out=N, bias=bias, epilogue=CPD_EPI_GEGLU, geglu_block=256)
            else:
                ops.gemm_conv(a, w, o, n_img=n_img, h=hh, w=ww, c0=cin, n_out=N, ksize=ks, bias=bias, variant=v, residual=r, ld_res=N if res else 0)
        for _ in range(5):
            launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            launch()
        e1.record()
        torch.cuda.synchronize()
        print(f"== M={M} N={N} K={K} variant={v} residual={res}: {e0.elapsed_time(e1) * 100:.1f} us per launch")
        ts = (C.c_ulonglong * 16)()
        lib.cpd_debug_gemm_timeline(ts)
        print(f"M={M} N={N} K={K}:")
        for i, nm in enumerate(NAMES):
            print(f"   {nm:40s} +{(ts[i] - ts[0]) / 1e3:7.2f} us")
        for i, nm in ((13, "producer: loop entered"), (11, "producer: first tile coordinates done"), (12, "producer: first empty-slot wait passed")):
            print(f"   {nm:40s} +{(ts[i] - ts[0]) / 1e3:7.2f} us")
        mm = (C.c_longlong * 96)()
        lib.cpd_debug_gemm_mma(mm)
        for tile in range(3):
            row = [(mm[(tile * 16 + k) * 2], mm[(tile * 16 + k) * 2 + 1]) for k in range(min(16, max(1, K // 64)))]
            row = [x for x in row if x[0]]
            if not row:
                continue
            if row[0][0]:
                t0 = mm[0]
                print(f"   MMA issuer tile {tile}: " + " ".join(f"[{a - t0}->{b - t0}]" for a, b in row))
        C.memset(C.addressof(mm), 0, C.sizeof(mm))
        ep = (C.c_longlong * 128)()
        lib.cpd_debug_gemm_epilogue(ep)
        names = ["chunk top", "top2", "tmem ld done", "math done", "st.shared done", "proxy fence", "arrived", "end"]
        for tile in range(2):
            for ch in range(8):
                row = [ep[(tile * 8 + ch) * 8 + ph] for ph in range(8)]
                if row[0] == 0:
                    continue
                t0 = ep[(tile * 8) * 8]
                print(f"   epilogue tile {tile} chunk {ch}: " + "  ".join(f"{n}=+{v - t0}" for n, v in zip(names, row)))
        C.memset(C.addressof(ep), 0, C.sizeof(ep))


if __name__ == "__main__":
    main()
