#!/usr/bin/env python
"""Phase timeline of the softmax warps of attention2_kernel, key blocks 8..15 of CTA 0 (needs a TIMELINE=1 build:
make -C complex_prompt_diffusion_b200/csrc clean all TIMELINE=1).  Cycles relative to tile 0's first stamp."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
NAMES = ["loop top", "S ready", "S in regs", "max done", "gate passed", "exp done", "P stored+arrive"]


def main():
    from complex_prompt_diffusion_b200 import ops
    lib = ops.load()
    B, H, Nq, Nk, d = 16, 8, 4096, 4096, 40
    dpad, nk_pad = 48, Nk
    ip = H * dpad
    q = torch.randn(B * Nq, H, dpad, device="cuda").half()
    k = torch.randn(B * nk_pad, H, dpad, device="cuda").half()
    vt = torch.randn(H, dpad, B * nk_pad, device="cuda").half()
    o = torch.empty(B * Nq, ip, device="cuda", dtype=torch.float16)
    for _ in range(3):
        ops.attention(q, k, vt, o, ldq=ip, ldk=ip, ldvt=B * nk_pad, ldo=ip, batch=B, heads=H, nq=Nq, nk=Nk, nk_pad=nk_pad, dpad=dpad,
                      scale=d ** -0.5, d_head=d)
    torch.cuda.synchronize()
    if "--persistent" in sys.argv:
        soft, mma = (C.c_longlong * 256)(), (C.c_longlong * 128)()
        lib.cpd_debug_attn3_timeline(soft, mma)
        t0 = soft[0]
        ev = []
        names = ["top", "S ready", "S in regs", "max done", "exp done", "P arrive"]
        for t in range(4):
            for j in range(8):
                for ph in range(6):
                    v = soft[(t * 8 + j) * 8 + ph]
                    if v:
                        ev.append((v - t0, f"tile {t} blk {8 + j}: {names[ph]}"))
        mn = ["wait P", "P seen", "S issue", "S issued"]
        for j in range(8):
            for t in range(4):
                for ph in range(4):
                    v = mma[(j * 4 + t) * 4 + ph]
                    if v:
                        ev.append((v - t0, f"    MMA blk {8 + j} tile {t}: {mn[ph]}"))
        for v, n in sorted(ev):
            if 0 <= v < 14000:
                print(f"{v:7d}  {n}")
        return
    ts = (C.c_longlong * 128)()
    lib.cpd_debug_attn_timeline(ts)
    t0 = ts[0]
    for j in range(8):
        for t in range(2):
            row = [ts[(t * 8 + j) * 8 + ph] - t0 for ph in range(7)]
            print(f"blk {8 + j} tile {t}: " + "  ".join(f"{n}={v}" for n, v in zip(NAMES, row)))
    for t in range(2):
        per = (ts[(t * 8 + 7) * 8] - ts[(t * 8) * 8]) / 7
        print(f"tile {t}: {per:.0f} cycles per block")


if __name__ == "__main__":
    main()
