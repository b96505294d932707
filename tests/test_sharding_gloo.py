"""N > 1 host logic on CPU: world_size-2 `gloo` processes exercise the leading-axis sharding arithmetic, the unique-id
handshake and the decomposition of the sharded reductions / row-sharded matmul against the oracle (the collectives
themselves run over NCCL on the GPU box: tests/test_multi_gpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from compute.scala_b200 import sharding
from oracle import reference as ref


def test_shard_rows_partitions_exactly():
    for rows in (0, 1, 7, 8, 16384, 16385):
        for world in (1, 2, 3, 8):
            blocks = [sharding.shard_rows(rows, world, r) for r in range(world)]
            assert blocks[0][0] == 0
            for (s0, n0), (s1, _) in zip(blocks, blocks[1:]):
                assert s0 + n0 == s1
            assert blocks[-1][0] + blocks[-1][1] == rows
            assert max(n for _, n in blocks) - min(n for _, n in blocks) <= 1
    assert sharding.shard_shape([16384, 16384], 8, 3) == [2048, 16384]
    assert sharding.shard_offsets([10, 4], 3) == [0, 16, 28, 40]
    with pytest.raises(ValueError):
        sharding.shard_rows(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, rows, cols):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # unique-id handshake: every rank ends up with rank 0's 128 bytes
        uid = sharding.exchange_unique_id(dist, lambda: bytes([7 + (i % 11) for i in range(128)]))
        assert uid == bytes([7 + (i % 11) for i in range(128)])
        # dataset E (exactly summable) of the full [rows, cols] tensor; this rank owns a contiguous row block
        full = (np.floor(ref.random_buffer(rows * cols, 5) * np.float32(9.0)) - np.float32(4.0)).astype(np.float32).reshape(rows, cols)
        start, n = sharding.shard_rows(rows, world, rank)
        mine = full[start : start + n]
        # full sum: local sum in the reference's CPU order, all-reduce of one float == the unsharded oracle
        part = torch.tensor([float(ref.sum_reference_cpu_order(mine.reshape(-1)))], dtype=torch.float32)
        dist.all_reduce(part)
        assert part.item() == float(ref.sum_reference_cpu_order(full.reshape(-1)))
        # axis 0 (sharded axis): partial column sums (fp32 left fold) + all-reduce of one row
        cols_part = torch.from_numpy(np.add.accumulate(mine, axis=0, dtype=np.float32)[-1].copy())
        dist.all_reduce(cols_part)
        assert np.array_equal(cols_part.numpy(), np.add.accumulate(full, axis=0, dtype=np.float32)[-1])
        # axis 1: purely local row sums + all-gather of equally sized blocks
        rows_part = torch.from_numpy(np.add.accumulate(mine, axis=1, dtype=np.float32)[:, -1].copy())
        if rows % world == 0:
            out = [torch.empty_like(rows_part) for _ in range(world)]
            dist.all_gather(out, rows_part)
            assert np.array_equal(torch.cat(out).numpy(), np.add.accumulate(full, axis=1, dtype=np.float32)[:, -1])
        # row-sharded matmul: A and C row blocks, B replicated, no exchange
        k = 24
        a = (np.floor(ref.random_buffer(rows * k, 9) * 9) - 4).astype(np.float32).reshape(rows, k)
        b = (np.floor(ref.random_buffer(k * cols, 10) * 9) - 4).astype(np.float32).reshape(k, cols)
        c_mine = torch.from_numpy((a[start : start + n].astype(np.float64) @ b.astype(np.float64)).astype(np.float32))
        if rows % world == 0:
            out = [torch.empty_like(c_mine) for _ in range(world)]
            dist.all_gather(out, c_mine)
            assert np.array_equal(torch.cat(out).numpy(), (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("rows,cols", [(64, 48), (33, 16)])
def test_world_size_2_gloo(rows, cols):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, rows, cols), nprocs=2, join=True)


def test_distribution_follows_the_lazy_graph():
    """host logic of the sharded Tensor API (no device work: nothing is evaluated): which operations keep a row block a row block,
    which produce partial sums, which need the sum, and which are refused because they would mix the sharded leading axis"""
    from compute.scala_b200 import cuda

    T = cuda.Tensor
    x = T.random([8, 6], seed=1).shard()
    assert x.distribution == "row block" and T.random([8, 6], seed=1).distribution == "whole"
    assert (x + x).distribution == "row block" and T.tanh(x).distribution == "row block" and (-x).distribution == "row block"
    assert (x * T.fill(2.0, [8, 6])).distribution == "row block"  # replicated operand
    assert x.sum().distribution == "whole"  # the global sum
    acc = x.split(0)[0]
    for p in x.split(0)[1:]:
        acc = acc + p
    assert acc.distribution == "partial sum"  # the fold of the local rows
    assert T.tanh(acc).distribution == "whole" and (acc * acc).distribution == "whole" and (acc - acc).distribution == "whole"
    assert acc.nonInline().distribution == "whole" and acc.reshape([2, 3]).distribution == "whole" and acc.broadcast([6, 2]).distribution == "whole"
    for view in (x.split(1)[0], x.permute([0, 1]), x.translate([0, 2]), x.reshape([8, 2, 3]), x.broadcast([8, 6, 4]), T.join([x, x]), T.join([x, x], 1), x.nonInline()):
        assert view.distribution == "row block"
    for refused in (lambda: x.permute([1, 0]), lambda: x.translate([1, 0]), lambda: x.reshape([6, 8]), lambda: x.reduce("max"), lambda: T.join([x, x], 0),
                    lambda: x.scale([4, 6]), lambda: x.shard(), lambda: T.scalar(1.0).shard()):
        with pytest.raises(cuda.ComputeCudaError):
            refused()
    # the matmul pattern over a row block of A and a replicated B stays a row block and is still the pattern
    b = T.random([6, 5], seed=2)
    prod = x.broadcast([8, 6, 5]) * b.reshape([1, 6, 5]).broadcast([8, 6, 5])
    c = prod.split(1)[0]
    for p in prod.split(1)[1:]:
        c = c + p
    assert c.distribution == "row block" and tuple(c.shape) == (8, 5)
    assert tuple(x.gather().shape) == (8, 6) and x.gather().distribution == "whole"  # one rank: the identity
    for rows in (0, 1, 7, 8, 16385):
        for world in (1, 2, 3, 8):
            for r in range(world):
                assert cuda.shard_rows(rows, world, r) == sharding.shard_rows(rows, world, r)
