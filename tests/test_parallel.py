"""Frame-chunk sharding + the trajectory all-gather, world_size 2 and 3 over gloo on the CPU."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vistracker_b200 import parallel as PL


def test_batches_follow_the_reference_boundaries():
    assert PL.batch_bounds(0, 1500, 512) == [(0, 512), (512, 1024), (1024, 1500)]
    assert PL.batch_bounds(10, 20, 96) == [(10, 20)]
    # 1500 frames in joint-optimisation batches of 96 -> 16 batches -> 2 per GPU on 8 GPUs
    per_rank = [PL.rank_frames(0, 1500, 96, 8, r) for r in range(8)]
    assert all(len(p) == 2 for p in per_rank)
    flat = [b for p in per_rank for b in p]
    assert flat == PL.batch_bounds(0, 1500, 96) and flat[-1] == (1440, 1500)
    # more ranks than batches: the surplus ranks get nothing, nothing is lost
    got = [PL.rank_frames(0, 100, 64, 4, r) for r in range(4)]
    assert [b for p in got for b in p] == [(0, 64), (64, 100)] and got[2] == [] and got[3] == []


def _worker(rank, world, port, lens, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start = sum(lens[:rank])
        t = torch.arange(start, start + lens[rank], dtype=torch.float32)[:, None] * torch.ones(1, PL.SMPLT_WIDTH) + 0.25 * rank
        full = PL.gather_trajectory(t)
        # second collective of the path: the per-frame neural predictions [t, 13] that the object smoother / HVOP-Net consume
        from vistracker_b200.pipeline import pack_neural
        n = lens[rank]
        frames = torch.arange(start, start + n, dtype=torch.float32)
        neural = PL.gather_trajectory(pack_neural(frames[:, None, None] * torch.ones(n, 3, 3), torch.zeros(n, 3), frames / 100))
        ok13 = tuple(neural.shape) == (sum(lens), 13) and bool(torch.equal(neural[:, 0], torch.arange(sum(lens), dtype=torch.float32))) and \
            bool(torch.allclose(neural[:, 12], torch.arange(sum(lens), dtype=torch.float32) / 100))
        q.put((rank, full.shape, float(full[:, 0].sum()), bool((full[1:, 3] >= full[:-1, 3]).all()) and ok13))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("lens", [(5, 3), (4, 0, 7)])
def test_ragged_trajectory_all_gather(lens):
    world = len(lens)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, lens, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = sum(lens)
    expect = sum(float(i) + 0.25 * r for r in range(world) for i in range(sum(lens[:r]), sum(lens[:r]) + lens[r]))
    for rank, shape, s, ordered in res:
        assert tuple(shape) == (total, PL.SMPLT_WIDTH) and abs(s - expect) < 1e-3 and ordered


def test_single_process_is_a_no_op():
    t = torch.randn(4, PL.SMPLT_WIDTH)
    assert PL.gather_trajectory(t) is t
    p, b, tr = PL.unpack_smplt(PL.pack_smplt(torch.ones(2, 156), torch.zeros(2, 10), torch.ones(2, 3)))
    assert p.shape == (2, 156) and b.shape == (2, 10) and tr.shape == (2, 3)
