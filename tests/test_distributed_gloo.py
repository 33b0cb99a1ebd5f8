"""CPU-only, world_size 2 and 3 over gloo: the sharding / gather logic of the multi-GPU driver
(ocrfa_b200/distributed.py). The per-batch recognizer is replaced by a deterministic stand-in whose output depends
only on the global line index, so any mis-sharding, mis-ordering or padding leak shows up in the gathered result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ocrfa_b200.distributed import gather_decoded, max_over_ranks, recognize_sharded, shard_range

T = 9


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_recognize(lo, hi):
    idx = torch.arange(lo, hi, dtype=torch.int64)
    lengths = (idx % (T + 1)).to(torch.int32)
    labels = ((idx[:, None] * 7 + torch.arange(T)[None, :]) % 185).to(torch.int32)
    labels = labels * (torch.arange(T)[None, :] < lengths[:, None])
    conf = (idx.to(torch.float64) + 0.5) / 1000.0
    return labels, lengths, conf


def _worker(rank, world, port, n_lines, batch, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        calls = []

        def rec(lo, hi):
            calls.append((lo, hi))
            return _fake_recognize(lo, hi)

        labels, lengths, conf = recognize_sharded(n_lines, rec, T, batch_size=batch)
        slowest = max_over_ranks(float(rank + 1), torch.device("cpu"))
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), labels=labels.numpy(), lengths=lengths.numpy(),
                 conf=conf.numpy(), calls=np.array(calls, dtype=np.int64).reshape(-1, 2), slowest=slowest)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_lines,batch", [(2, 37, 8), (2, 1, 4), (3, 10, 256), (2, 0, 4)])
def test_sharded_recognition_gathers_global_order(tmp_path, world, n_lines, batch):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_lines, batch, str(tmp_path)), nprocs=world, join=True)
    exp_l, exp_n, exp_c = _fake_recognize(0, n_lines)
    covered = []
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(d["labels"], exp_l.numpy()) and np.array_equal(d["lengths"], exp_n.numpy())
        assert np.array_equal(d["conf"], exp_c.numpy())
        assert float(d["slowest"]) == float(world)
        lo, hi = shard_range(n_lines, r, world)
        for a, b in d["calls"].tolist():
            assert lo <= a < b <= hi and b - a <= batch
            covered += list(range(a, b))
    assert covered == list(range(n_lines))  # every line processed exactly once, in order, by exactly one rank


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 256, 1000003):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) == (n + world - 1) // world
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_single_process_gather_is_identity():
    la, ln, cf = _fake_recognize(0, 5)
    a, b, c = gather_decoded(la, ln, cf, 5)
    assert torch.equal(a, la) and torch.equal(b, ln) and torch.equal(c, cf)
