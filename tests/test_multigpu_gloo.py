"""CPU, world_size 2 over gloo: the sharding / result-gather host logic used by bench.py for N > 1."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fastspeech2_b200.parallel import Rank0Gather, gather_padded, shard_microbatches


def test_shard_microbatches_partition():
    for n in (1, 7, 8, 13):
        for world in (1, 2, 4, 8):
            parts = [shard_microbatches(n, r, world) for r in range(world)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T = 5 + 3 * rank
    local = torch.arange(2 * T * 4, dtype=torch.float32).reshape(2, T, 4) + 1000 * rank
    lens = torch.tensor([T, T - 2], dtype=torch.int64)
    out, all_lens = gather_padded(local, lens)
    q.put((rank, out, all_lens))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_padded_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out, lens in res:
        assert out.shape == (4, 8, 4) and lens.tolist() == [5, 3, 8, 6]
        assert torch.equal(out[:2, :5], torch.arange(40, dtype=torch.float32).reshape(2, 5, 4))
        assert torch.all(out[:2, 5:] == 0)
        assert torch.equal(out[2:], torch.arange(64, dtype=torch.float32).reshape(2, 8, 4) + 1000)


def _worker_rank0(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = Rank0Gather(align=4)
    results = []
    for step in range(3):                              # lengths differ per rank and per step; two buffers alternate
        T = 6 + 2 * rank - step
        local = torch.arange(2 * T, dtype=torch.float32).reshape(2, T) + 100 * rank + 1000 * step
        g.submit(local, torch.tensor([T, T - 1], dtype=torch.int64))
    g.flush()
    q.put((rank, g.last()))
    dist.barrier()
    dist.destroy_process_group()


def test_rank0_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_rank0, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[1] is None                              # only rank 0 receives
    payload, lens = res[0]
    assert lens.tolist() == [4, 3, 6, 5]               # step 2: T = 4 on rank 0, 6 on rank 1
    assert payload.shape[0] == 4 and payload.shape[1] >= 8
    assert torch.equal(payload[:2, :4], torch.arange(8, dtype=torch.float32).reshape(2, 4) + 2000)
    assert torch.equal(payload[2:, :6], torch.arange(12, dtype=torch.float32).reshape(2, 6) + 2100)
