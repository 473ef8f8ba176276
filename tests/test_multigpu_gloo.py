"""CPU, world_size 2 over gloo: the sharding / result-gather host logic used by bench.py for N > 1."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fastspeech2_b200.parallel import gather_padded, shard_microbatches


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
