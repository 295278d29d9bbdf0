"""Multi-rank host logic on CPU: partition arithmetic and the gloo all-gather paths (world size 2)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from complex_prompt_diffusion_b200 import dist as D


@pytest.mark.parametrize("batch,rows,world", [(4, 4, 1), (4, 4, 2), (4, 4, 8), (8, 2, 8), (2, 2, 8), (3, 4, 8), (5, 4, 2), (1, 4, 8),
                                              (1, 2, 4), (0, 4, 2)])
def test_partition_covers_every_unit_once(batch, rows, world):
    owners = {}
    for r in range(world):
        p = D.partition(batch, rows, world, r)
        for img in p.images:
            for row in p.rows:
                if batch >= world:
                    owners.setdefault((img, row), []).append(r)
                else:
                    owners.setdefault((img, row), []).append(r)
    assert set(owners) == {(i, j) for i in range(batch) for j in range(rows)}
    assert all(len(v) == 1 for v in owners.values())
    sizes = [len(D.partition(batch, rows, world, r).images) * len(D.partition(batch, rows, world, r).rows) for r in range(world)]
    if batch >= world and batch:
        assert max(sizes) - min(sizes) <= rows


def test_config_layouts():
    # cfg 2: B=4, R=4 on 8 GPUs -> 2 ranks per image, 2 rows each, all-gather inside groups of 2
    p = D.partition(4, 4, 8, 5)
    assert p.images == [2] and p.rows == [2, 3] and p.group_ranks == [4, 5] and p.needs_allgather
    # cfg 3: B=8 on 8 GPUs -> one image per GPU, no communication
    p = D.partition(8, 2, 8, 3)
    assert p.images == [3] and p.rows == [0, 1] and not p.needs_allgather
    # cfg 4: B=2, R=2 on 8 GPUs -> 4 ranks per image, only 2 have a row
    ps = [D.partition(2, 2, 8, r) for r in range(8)]
    assert [len(p.rows) for p in ps] == [1, 1, 0, 0, 1, 1, 0, 0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # row sharding: 1 image, 3 rows over 2 ranks
        part = D.partition(1, 3, world, rank)
        L = 8
        full = torch.arange(3 * L, dtype=torch.float32).view(3, L)
        got = D.allgather_eps_rows(full[part.rows].clone(), part)
        ok1 = torch.equal(got, full)
        # image sharding: 5 images over 2 ranks
        pi = D.partition(5, 2, world, rank)
        x = torch.arange(5 * 4, dtype=torch.float32).view(5, 4)
        got2 = D.gather_images(x[pi.images].clone(), 5, world)
        ok2 = torch.equal(got2, x)
        # twice through the cached plan (preallocated buffers are reused), equal row counts: 4 rows over 2 ranks
        part4 = D.partition(1, 4, world, rank)
        full4 = torch.arange(4 * L, dtype=torch.float32).view(4, L) + 100.0
        for rep in range(2):
            ok1 = ok1 and torch.equal(D.allgather_eps_rows((full4 + rep)[part4.rows].clone(), part4), full4 + rep)
        # group-shared noise (row-sharded stochastic samplers): per-rank RNG streams differ, the draw must not
        torch.manual_seed(1234 + rank)
        draw = D.shared_noise_sampler(None, None, 0)
        n1, n2 = draw(torch.zeros(2, 3)), draw(torch.zeros(2, 3))
        gathered = [torch.empty_like(n1) for _ in range(world)]
        dist.all_gather(gathered, torch.cat([n1, n2])[:2])
        ok3 = all(torch.equal(g, gathered[0]) for g in gathered) and not torch.equal(n1, n2)
        custom = D.shared_noise_sampler(lambda x: torch.full_like(x, float(rank + 7)), None, 0)
        ok3 = ok3 and torch.equal(custom(torch.zeros(4)), torch.full((4,), 7.0))
        q.put((rank, ok1, ok2 and ok3))
    finally:
        dist.destroy_process_group()


def test_allgather_paths_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok1 and ok2 for _, ok1, ok2 in res), res


@pytest.mark.parametrize("n_frames,seg,world", [(64, 8, 1), (64, 8, 2), (64, 8, 4), (64, 8, 8), (10, 4, 3), (5, 8, 2), (0, 4, 2)])
def test_animation_segments_cover_every_frame_once(n_frames, seg, world):
    """BASELINE.json configs[4]: frames shard across ranks as whole independent segments."""
    from complex_prompt_diffusion_b200.animation import segments, segments_for_rank
    seen = []
    for r in range(world):
        for (a, b) in segments_for_rank(n_frames, seg, world, r):
            assert 0 <= a < b <= n_frames and b - a <= seg
            seen += list(range(a, b))
    assert sorted(seen) == list(range(n_frames))
    assert sum(b - a for a, b in segments(n_frames, seg)) == n_frames
    if n_frames == 64 and seg == 8:
        assert all(len(segments_for_rank(n_frames, seg, world, r)) == 8 // world for r in range(world))
