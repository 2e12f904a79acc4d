"""Multi-GPU host logic on CPU: world_size-2 gloo run of the window sharding + host-side gather
(whisper-mlx_b200/sharding.py), the only cross-rank step of the path (no data-path collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_windows, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from whisper_mlx_b200.sharding import gather_by_index, plan_windows, shard_indices

    windows = plan_windows(n_windows * 3000 - 1234, [(0, n_windows * 3000 - 1234)])
    mine = shard_indices(len(windows), rank, world)
    # stand-in for decode: the "segments" of window i are a function of (seek, size) only
    local = {i: None if i % 5 == 4 else [{"seek": windows[i][0], "size": windows[i][1], "rank": rank}] for i in mine}
    merged = gather_by_index(local, len(windows))
    # timing reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, mine, merged, t.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_window_sharding():
    world, n_windows = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_windows, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, mine0, merged0, t0), (r1, mine1, merged1, t1) = out
    assert sorted(mine0 + mine1) == list(range(n_windows)) and not set(mine0) & set(mine1)
    assert abs(len(mine0) - len(mine1)) <= 1
    assert merged0 == merged1 and len(merged0) == n_windows
    assert t0 == t1 == 2.0
    for i, segs in enumerate(merged0):
        if i % 5 == 4:
            assert segs is None
        else:
            assert segs[0]["seek"] == i * 3000 and segs[0]["rank"] == (0 if i in mine0 else 1)
    assert merged0[-1][0]["size"] == 3000 - 1234


def test_shard_indices_cover_everything():
    from whisper_mlx_b200.sharding import plan_windows, shard_indices

    for n in (0, 1, 7, 120):
        for w in (1, 2, 4, 8):
            parts = [shard_indices(n, r, w) for r in range(w)]
            assert sum(parts, []) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)
    assert plan_windows(7000, [(0, 7000)]) == [(0, 3000), (3000, 3000), (6000, 1000)]
    assert plan_windows(7000, [(500, 4000), (6500, 7000)]) == [(500, 3000), (3500, 500), (6500, 500)]
