"""Host-side multi-rank logic on CPU: world_size = 2, gloo backend.  The per-rank compute is
replaced by the CPU oracle through a stand-in context, so only the sharding and the gather are under test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gaplac_b200 import shard, workloads as W


def test_shard_ranges_cover_the_batch():
    for B in (0, 1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            r = [shard.shard_range(B, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


class _OracleCtx:
    """Stand-in for gaplac_b200._lib.Context in this CPU test only."""

    def lml_batched(self, prog, X, Y, Theta, sigma2, jitter=0.0):
        from oracle import c_oracle as CO
        return CO.lml_batched(prog, X, Y, Theta, sigma2, jitter, x_batched=np.ndim(X) == 3)


def _per_item_problem(B, n=24):
    rng = np.random.default_rng(B)
    d = W.make_c2(n=n, B=B)
    Xb = rng.uniform(-5, 5, (B, n, 1))               # a different input set per item
    Yb = rng.standard_normal((B, n))
    return d["ops"], Xb, Yb, d["Theta"], rng.uniform(0.05, 0.2, B)


def _worker_per_item(rank, world, port, B, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ops, Xb, Yb, Th, s2 = _per_item_problem(B)
    lml, _ = shard.sharded_logpdf(_OracleCtx(), ops, Xb, Yb, Th, s2)
    np.save(os.path.join(out_dir, f"lml_{rank}.npy"), lml)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_per_item_inputs(tmp_path):
    """X (B, n, d), Y (B, n) and sigma2 (B,) are all sliced to the rank's block: rank 1 must read ITS items' X."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    B = 7
    mp.spawn(_worker_per_item, args=(2, port, B, str(tmp_path)), nprocs=2, join=True)
    from oracle import c_oracle as CO
    ops, Xb, Yb, Th, s2 = _per_item_problem(B)
    ref, _ = CO.lml_batched(ops, Xb, Yb, Th, s2, x_batched=True)
    for rank in range(2):
        assert np.array_equal(np.load(tmp_path / f"lml_{rank}.npy"), ref)


def _worker(rank, world, port, B, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = W.make_c2(n=48, B=B)
    lml, info = shard.sharded_logpdf(_OracleCtx(), d["ops"], d["X"], d["y"], d["Theta"], 0.0)
    np.save(os.path.join(out_dir, f"lml_{rank}.npy"), lml)
    np.save(os.path.join(out_dir, f"info_{rank}.npy"), info)
    t = torch.tensor([float(rank)])
    dist.all_reduce(t)
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [9, 16])
def test_two_ranks_gather_in_batch_order(tmp_path, B):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, B, str(tmp_path)), nprocs=2, join=True)
    from oracle import c_oracle as CO
    d = W.make_c2(n=48, B=B)
    ref, _ = CO.lml_batched(d["ops"], d["X"], d["y"], d["Theta"], 0.0)
    for rank in range(2):
        got = np.load(tmp_path / f"lml_{rank}.npy")
        assert got.shape == (B,)
        assert np.array_equal(got, ref)           # every rank holds the full vector, in batch order
        assert not np.load(tmp_path / f"info_{rank}.npy").any()
