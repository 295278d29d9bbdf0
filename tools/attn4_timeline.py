#!/usr/bin/env python
"""Phase timeline of attention4_kernel (the 4096^2 d = 40 self-attention): softmax warps and MMA issuers of CTA 0, key blocks 8..15.
Needs the TIMELINE=1 build:  make -C complex_prompt_diffusion_b200/csrc TIMELINE=1 BUILD=build_tl OUT=../libcpd_b200_tl.so
  CPD_B200_LIB=complex_prompt_diffusion_b200/libcpd_b200_tl.so python tools/attn4_timeline.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


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
    soft, mma = (C.c_longlong * 256)(), (C.c_longlong * 128)()
    lib.cpd_debug_attn4_timeline.restype = C.c_int
    lib.cpd_debug_attn4_timeline(soft, mma)
    t0 = min(v for v in soft if v)
    sn = ["top", "S ready", "S in regs", "max + exchange done", "exp done", "PV(j-1) done seen", "P stored + arrive"]
    mn = ["top", "S(j) read by softmax", "S(j+1) issued", "P(j) seen", "PV(j) issued"]
    ev = []
    for t in range(2):
        for hf in range(2):
            for j in range(8):
                for ph in range(7):
                    v = soft[((t * 2 + hf) * 8 + j) * 8 + ph]
                    if v:
                        ev.append((v - t0, f"tile {t} half {hf} blk {8 + j}: {sn[ph]}"))
        for j in range(8):
            for ph in range(5):
                v = mma[(t * 8 + j) * 8 + ph]
                if v:
                    ev.append((v - t0, f"        issuer {t} blk {8 + j}: {mn[ph]}"))
    lim = int(sys.argv[1]) if len(sys.argv) > 1 else 9000
    for v, n in sorted(ev):
        if 0 <= v < lim:
            print(f"{v:7d}  {n}")
    for t in range(2):
        for hf in range(2):
            base = (t * 2 + hf) * 8
            per = (soft[(base + 7) * 8] - soft[base * 8]) / 7
            ph = [0.0] * 6
            for j in range(8):
                for p in range(6):
                    ph[p] += (soft[(base + j) * 8 + p + 1] - soft[(base + j) * 8 + p]) / 8
            print(f"tile {t} half {hf}: {per:.0f} cycles per block; mean phase lengths: " + ", ".join(f"{sn[p + 1]} {ph[p]:.0f}" for p in range(6)))


if __name__ == "__main__":
    main()
