#!/usr/bin/env python
"""Fused sampler-step kernel at a saturating batch (HBM roofline) and at the named shape (latency)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from complex_prompt_diffusion_b200 import ops
    from complex_prompt_diffusion_b200._lib import CPD_DPMPP_2M, CPD_PRED_EPSILON
    dev = "cuda"
    for tag, B, dt in (("saturating bf16 eps", 704, torch.bfloat16), ("saturating fp32 eps", 512, torch.float32), ("named shape fp32 eps", 4, torch.float32)):
        # three disjoint operand sets, rotated: no launch finds its operands in the 126 MB L2 (DRAM bytes = algorithmic bytes)
        nsets = 3 if B > 16 else 1
        sets = [(torch.randn(B * 4, 4, 64, 64, device=dev).to(dt), torch.randn(B, 4, 64, 64, device=dev), torch.randn(B, 4, 64, 64, device=dev))
                for _ in range(nsets)]
        eps = sets[0][0]
        args = dict(n_sub=3, weights=[1.0, 0.6, -0.4], mask_scalars=[1.0] * 3, masks=[None] * 3, guidance=7.5, sampler=CPD_DPMPP_2M,
                    pred_type=CPD_PRED_EPSILON, sigma_hat=2.0, dpm_ratio=0.8, dpm_expm1=-0.2, dpm_c1=1.5, dpm_c2=0.5, dpm_first=0,
                    write_old=1)
        for k in range(3 * nsets):
            ops.sampler_step(sets[k % nsets][0], sets[k % nsets][1], old_denoised=sets[k % nsets][2], **args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        n = 21
        for k in range(n):
            ops.sampler_step(sets[k % nsets][0], sets[k % nsets][1], old_denoised=sets[k % nsets][2], **args)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / n
        nbytes = (4 * eps.element_size() + 16) * 4 * 4096 * B
        print(f"{tag:24s} B={B:4d}: {us:8.1f} us  {nbytes / us / 1e3:7.1f} GB/s ({nbytes / 1e6:.1f} MB)")


if __name__ == "__main__":
    main()
