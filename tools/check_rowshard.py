#!/usr/bin/env python
"""torchrun --nproc-per-node 2 tools/check_rowshard.py : batch 1 on 2 GPUs (row sharding + NCCL all-gather of eps) must give
exactly the latents of the single-GPU run; batch 2 on 2 GPUs (image sharding) likewise."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from complex_prompt_diffusion_b200 import samplers, dist as D
    from complex_prompt_diffusion_b200.models.unet import UNetModel
    from complex_prompt_diffusion_b200.models import fixtures
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = fixtures.UNET_PRESETS["tiny"]
    unet = UNetModel(fixtures.random_state_dict(fixtures.unet_param_shapes(cfg), seed=0), device=dev, **fixtures.unet_kwargs("tiny"))
    g = torch.Generator().manual_seed(0)
    uc = torch.randn(1, 77, cfg["context_dim"], generator=g)
    embs = [torch.randn(1, 77, cfg["context_dim"], generator=g) for _ in range(3)]
    c = {"and": [(1.0, embs[0], None, 1), (0.6, embs[1], None, 1)], "not": [(0.4, embs[2], None, 1)]}
    x_T = torch.randn(2, 4, 32, 32, generator=g)
    wrapper = samplers.make({"name": "DPM++ 2m", "args": {}}, {"model": {"unet": unet}})
    kw = dict(unconditional_guidance_scale=7.5, scheduler="karras", rng_compat=False)
    ok = True
    for batch in (1, 2):
        wrapper.sampler.denoiser.set_row_partition(None)
        ref = wrapper.sampler.sample(steps=6, batch_size=batch, shape=[4, 32, 32], x_T=x_T[:batch].clone(), conditioning=c,
                                     unconditional_conditioning=uc, **dict(kw)).clone()
        got = D.sample_sharded(wrapper, steps=6, batch=batch, shape=[4, 32, 32], x_T=x_T[:batch].clone(), conditioning=c,
                               unconditional_conditioning=uc, **dict(kw))
        same = torch.equal(ref, got)
        err = ((ref - got).norm() / ref.norm()).item()
        if dist.get_rank() == 0:
            mode = "row-sharded + all-gather" if batch < dist.get_world_size() else "image-sharded"
            print(f"batch {batch} on {dist.get_world_size()} GPUs ({mode}): bit-identical={same} rel={err:.2e}")
        # Same kernels on both sides, but split-K layers add fp32 partials with atomics (order varies) and the tuned tile
        # variants depend on the batch: low-order bits differ and a random-weight UNet amplifies them over the steps.
        # CPD_GEMM_AUTOTUNE=0 (no split-K) makes the comparison bit-exact.
        ok = ok and err < 2e-2
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
