"""Frame sequences on the denoising loop (BASELINE.json configs[4]; the part of cpd/animation.py:125-178 that drives the
hot path).

The reference's `render_animation_step` re-prompts every frame, warps the previous frame in IMAGE space (cv2 / pytorch3d /
MiDaS - outside the hot-path scope), re-encodes it and calls `render(latent=..., decode=True, denoising_strength=...)`,
i.e. `KDiffusionSampler.sample` with the img2img branch (k_diffusion.py:64-70); the first frame of a sequence is a plain
txt2img sample from seeded noise (`seed_everything(seed); torch.randn(...)`, animation.py:160-161).  Frames inside such a
chain are sequential (frame i starts from frame i-1), so the unit that shards across GPUs is an independent SEGMENT
(key-frame to key-frame): `segments_for_rank` deals whole segments to ranks, no collective on the data path, and the
finished latents are gathered once at the end (`gather_frames`).
"""
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch


def segments(n_frames: int, segment_len: int) -> List[Tuple[int, int]]:
    """[start, stop) frame ranges of the independent segments (the last one may be shorter)."""
    if n_frames < 0 or segment_len <= 0:
        raise ValueError("n_frames >= 0 and segment_len > 0 required")
    return [(s, min(s + segment_len, n_frames)) for s in range(0, n_frames, segment_len)]


def segments_for_rank(n_frames: int, segment_len: int, world: int, rank: int) -> List[Tuple[int, int]]:
    """Round-robin deal of whole segments: rank r renders segments r, r + world, ...  Every frame is rendered by exactly
    one rank; with 64 frames in segments of 8 over 1 / 2 / 4 / 8 GPUs every rank gets 8 / 4 / 2 / 1 segments."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} / world {world}")
    return segments(n_frames, segment_len)[rank::world]


def render_sequence(wrapper, frames: Sequence[dict], *, steps: int, shape, segment_len: int, strength: float = 0.5,
                    world: int = 1, rank: int = 0, transform: Optional[Callable] = None, **sample_kwargs) -> Dict[int, torch.Tensor]:
    """Render the frames of this rank's segments.  `frames[i]` holds the per-frame kwargs of the sampler call
    (`conditioning`, `unconditional_conditioning`, optionally `seed`, `y`, ...).  Frame 0 of a segment: txt2img from
    `torch.manual_seed(seed); torch.randn` (animation.py:160-161); later frames: img2img from the previous latent
    (`decode=True`, `denoising_strength=strength`) after the optional latent-space `transform(prev_latent, i)` hook that
    stands where the reference warps the decoded image.  Returns {frame index: latent [1, C, h, w] on the device}."""
    sampler = wrapper.sampler
    out = {}
    for (s0, s1) in segments_for_rank(len(frames), segment_len, world, rank):
        prev = None
        for i in range(s0, s1):
            kw = dict(sample_kwargs)
            kw.update({k: v for k, v in frames[i].items() if k != "seed"})
            if prev is None:
                torch.manual_seed(int(frames[i].get("seed", i)))
                x_T = torch.randn([1] + list(shape))
                lat = sampler.sample(steps=steps, batch_size=1, shape=list(shape), x_T=x_T, **kw)
            else:
                src = transform(prev, i) if transform is not None else prev
                torch.manual_seed(int(frames[i].get("seed", i)))
                lat = sampler.sample(steps=steps, batch_size=1, shape=list(shape), x_T=src, decode=True,
                                     denoising_strength=strength, **kw)
            prev = lat.clone()
            out[i] = prev
    return out


def gather_frames(local: Dict[int, torch.Tensor], n_frames: int, group=None) -> Optional[List[torch.Tensor]]:
    """Collect every rank's finished latents (once, after the loops): returns the list of all frames on every rank."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [local[i] for i in range(n_frames)]
    boxes = [None] * dist.get_world_size(group)
    dist.all_gather_object(boxes, {i: t.cpu() for i, t in local.items()}, group=group)
    merged = {}
    for b in boxes:
        merged.update(b)
    if sorted(merged) != list(range(n_frames)):
        raise RuntimeError("frame segments do not cover the sequence exactly once")
    return [merged[i] for i in range(n_frames)]
