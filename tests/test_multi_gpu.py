"""GPU parity for the sharded paths (needs >= 2 GPUs on the box; skipped otherwise): leading-axis shards per process,
NCCL (inside libcompute_cuda.so) only for combining partial reductions and gathering row blocks — checked on dataset E
(exact in any order) against numpy on the unsharded data."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _axis_sum(x, axis):
    parts = x.split(axis)
    acc = parts[0]
    for p in parts[1:]:
        acc = acc + p
    return acc


def _worker(rank, world, port, multicast):
    import torch.distributed as dist

    os.environ["CC_MULTICAST"] = "1" if multicast else "0"  # NVLS multicast stores, or one store per peer over CUDA IPC mappings

    from compute.scala_b200 import cuda, sharding
    from oracle import reference as ref

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # only carries the NCCL unique id
    try:
        cuda.init(rank)
        comm = sharding.Communicator(cuda, dist)
        T = cuda.Tensor
        rows, cols = 1024, 2048
        full = (np.floor(ref.random_buffer(rows * cols, 5) * np.float32(9.0)) - np.float32(4.0)).astype(np.float32).reshape(rows, cols)
        start, n = sharding.shard_rows(rows, world, rank)
        shard = T(full[start : start + n])
        assert comm.peer, "NVLink peer mailboxes should map on an HGX box"
        col_sums = _axis_sum(shard, 0)
        for route in (True, False, True, True):  # fused / one-shot kernels over peer memory, NCCL, and back (epoch parity)
            comm.route_peer(route)
            s = comm.full_sum(shard)
            assert s.to_host(1)[0] == np.float32(full.astype(np.int64).sum())
            c0 = comm.axis0_sum(col_sums)
            assert np.array_equal(c0.to_host(cols), full.astype(np.int64).sum(axis=0).astype(np.float32))
            s.release(), c0.release()
        # ragged one-shot all-reduces, back to back (mailbox double buffering)
        for length in (1, 3, 1000, 1025, 65536):
            v = np.arange(length, dtype=np.float32) % 17 + rank
            buf = cuda.Buffer.from_host(v)
            for rep in range(3):
                cuda.allreduce_sum(buf, length)
            got = buf.to_host(length)
            base = np.arange(length, dtype=np.float32) % 17
            expect = ((base * world + sum(range(world))) * world) * world
            assert np.array_equal(got, expect), (length, got[:4], expect[:4])
            buf.release()
        # ragged one-shot all-gathers over the same mailboxes, back to back, against the NCCL route
        for length in (1, 3, 1000, 1024, 1025, 4099, 65536, 65537):  # 65537 floats per rank: above the mailbox capacity -> NCCL
            v = (np.arange(length, dtype=np.float32) % 251) + 1000 * rank
            send = cuda.Buffer.from_host(v)
            expect = np.concatenate([(np.arange(length, dtype=np.float32) % 251) + 1000 * r for r in range(world)])
            for route in (True, False, True):
                comm.route_peer(route)
                recv = cuda.Buffer.from_host(np.full(length * world + 8, -1.0, np.float32))
                before = cuda.stats()["device_kernels"]
                cuda.allgather(send, recv, length)
                got = recv.to_host(length * world + 8)
                assert np.array_equal(got[: length * world], expect), (length, route)
                assert (got[length * world :] == -1.0).all()  # nothing written past the gathered vector
                if route and length <= 65536:
                    assert cuda.stats()["device_kernels"] == before + 1  # our kernel, not an NCCL call
                recv.release()
            send.release()
        comm.route_peer(True)
        s = comm.full_sum(shard)
        c0 = comm.axis0_sum(col_sums)
        c1 = comm.axis1_sum(_axis_sum(shard, 1), gather=True)
        assert np.array_equal(c1.to_host(rows), full.astype(np.int64).sum(axis=1).astype(np.float32))
        m, k, nn = 512, 256, 512
        a = (np.floor(ref.random_buffer(m * k, 9) * 9) - 4).astype(np.float32).reshape(m, k)
        b = (np.floor(ref.random_buffer(k * nn, 10) * 9) - 4).astype(np.float32).reshape(k, nn)
        ms, mn = sharding.shard_rows(m, world, rank)
        A, B = cuda.Buffer.from_host(a[ms : ms + mn]), cuda.Buffer.from_host(b)
        Cw = comm.matmul_rows(A, B, mn, nn, k, gather=True)
        assert np.array_equal(Cw.to_host(m * nn).reshape(m, nn), (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32))
        # the same exchange fused into the contraction's epilogue (TMA stores into every rank's copy over NVLink), vs NCCL
        for (m, k, nn) in ((512, 256, 512), (2 * 200, 96, 132), (2 * 1000, 300, 1000)):
            a = (np.floor(ref.random_buffer(m * k, 9) * 9) - 4).astype(np.float32).reshape(m, k)
            b = (np.floor(ref.random_buffer(k * nn, 10) * 9) - 4).astype(np.float32).reshape(k, nn)
            ms, mn = sharding.shard_rows(m, world, rank)
            A2, B2 = cuda.Buffer.from_host(a[ms : ms + mn]), cuda.Buffer.from_host(b)
            want = (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32)
            comm._require_equal_blocks(mn * nn)  # (its one-off agreement exchange is a kernel too: keep it out of the launch counts below)
            arena = comm.gather_arena(mn * nn * world)
            if not multicast:
                assert not arena.is_multicast
            elif rank == 0:
                print("symmetric arena has an NVLS multicast mapping:", arena.is_multicast)
            for rep, config in enumerate((None, None, "512", "256", "128", "64")):  # back to back: the entry barrier protects the arena
                # every tile configuration of the gather epilogue: CTA pairs (cta_group::2) and single CTAs
                if config:
                    os.environ["CC_GEMM_FORCE_CONFIG"] = config
                s0 = cuda.stats()["device_kernels"]
                Cf = comm.matmul_rows(A2, B2, mn, nn, k, gather=True, fused=True)
                used = cuda.stats()["device_kernels"] - s0
                os.environ.pop("CC_GEMM_FORCE_CONFIG", None)
                assert used == (5 if rep == 0 else 4), used  # barrier, split A, (split B once), contraction + gather, barrier
                assert np.array_equal(Cf.to_host(m * nn).reshape(m, nn), want), (m, k, nn, rep, config)
                Cf.release()
            Cn = comm.matmul_rows(A2, B2, mn, nn, k, gather=True, fused=False)
            assert np.array_equal(Cn.to_host(m * nn).reshape(m, nn), want)
            for x in (A2, B2, Cn):
                x.release()
        for x in (s, c0, c1, A, B, Cw):
            x.release()
        _api_level(cuda, comm, sharding, ref, rank, world)
        cuda.synchronize()
        comm.close()
    finally:
        dist.destroy_process_group()


def _api_level(cuda, comm, sharding, ref, rank, world):
    """the same exchanges through the Tensor API: a row block declared with .shard() carries its distribution through the lazy graph
    (include/compute_cuda.h: ct_shard / ct_gather); user code is the single-GPU code (benchmarks.scala:188-191, README.md:301-310)"""
    T = cuda.Tensor
    rows, cols = 1024, 2048
    full = (np.floor(ref.random_buffer(rows * cols, 5) * np.float32(9.0)) - np.float32(4.0)).astype(np.float32).reshape(rows, cols)
    start, n = sharding.shard_rows(rows, world, rank)
    x = T(full[start : start + n]).shard()
    i64 = full.astype(np.int64)
    for route in (True, False):
        comm.route_peer(route)
        assert x.sum().flatArray()[0] == np.float32(i64.sum())  # global sum: local fold + all-reduce of one float
        assert (x * x + x).sum().flatArray()[0] == np.float32((i64 * i64 + i64).sum())  # inline operand: fused fold, then all-reduce
        col = comm.fold(x.split(0))
        assert col.distribution == "partial sum"
        assert np.array_equal(col.flatArray(), i64.sum(axis=0).astype(np.float32))  # all-reduced when evaluated
        assert np.array_equal((col * col).flatArray(), (i64.sum(axis=0) ** 2).astype(np.float32))  # ... or used by anything but +
        row = comm.fold(x.split(1))
        assert row.distribution == "row block"
        assert np.array_equal(row.flatArray(), i64[start : start + n].sum(axis=1).astype(np.float32))  # stays sharded: no exchange
        assert np.array_equal(row.gather().flatArray(), i64.sum(axis=1).astype(np.float32))
        assert np.array_equal((x * x - x).gather().flatArray().reshape(rows, cols), full * full - full)  # elementwise: no exchange until gathered
        # compute step + collective = ONE kernel on the peer route (the reduction's final stage pushes / collects over the mailboxes itself),
        # reduction + collective call on the NCCL route; same values either way (checked above), here the launch counts
        counts = []
        for build in (lambda: comm.fold(x.split(0)), lambda: comm.fold(x.split(1)).gather(), lambda: x.sum()):
            s0 = cuda.stats()["device_kernels"]
            build().flatArray()
            counts.append(cuda.stats()["device_kernels"] - s0)
        if route:
            assert counts == [1, 1, 1], counts
    comm.route_peer(True)
    for (m, k, nn) in ((512, 256, 512), (2 * 200, 96, 132), (1024, 512, 1024)):
        a = (np.floor(ref.random_buffer(m * k, 9) * 9) - 4).astype(np.float32).reshape(m, k)
        b = (np.floor(ref.random_buffer(k * nn, 10) * 9) - 4).astype(np.float32).reshape(k, nn)
        ms, mn = sharding.shard_rows(m, world, rank)
        want = (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32)
        c = comm.matmul_pattern(T(a[ms : ms + mn]).shard(), T(b))
        assert c.distribution == "row block"
        big = m * k * nn >= 2**25
        assert c.compile().info.kind == (2 if big else 1)  # the pattern, on this rank's rows: the tcgen05 contraction from 2^25 multiply-adds
        assert np.array_equal(c.flatArray().reshape(mn, nn), want[ms : ms + mn])
        for zero_copy in (False, True):
            s0 = cuda.stats()["device_kernels"]
            g = c.gather(zero_copy=zero_copy)
            got = g.flatArray().reshape(m, nn)
            assert np.array_equal(got, want), (m, k, nn, zero_copy)
        assert g.distribution == "whole"
    # uneven row blocks: a gather is an IllegalArgument on EVERY rank (no hang); sums and sharded results still work
    urows = 2 * 100 + 1
    ufull = (np.floor(ref.random_buffer(urows * 64, 6) * 9) - 4).astype(np.float32).reshape(urows, 64)
    us, un = sharding.shard_rows(urows, world, rank)
    u = T(ufull[us : us + un]).shard()
    assert u.sum().flatArray()[0] == np.float32(ufull.astype(np.int64).sum())
    assert np.array_equal(comm.fold(u.split(0)).flatArray(), ufull.astype(np.int64).sum(axis=0).astype(np.float32))
    try:
        comm.fold(u.split(1)).gather().flatArray()
        raise AssertionError("uneven blocks were gathered")
    except cuda.ComputeCudaError as e:
        assert "equal row blocks" in str(e)
    try:
        comm.axis1_sum(comm.fold(u.split(1)))
        raise AssertionError("uneven blocks were gathered")
    except ValueError as e:
        assert "equal blocks" in str(e)


@pytest.mark.parametrize("multicast", [True, False], ids=["nvls-multicast", "ipc-peer-stores"])
def test_sharded_reductions_and_matmul_two_gpus(multicast):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under `gpurun --gpus 2`)")
    mp.spawn(_worker, args=(2, _free_port(), multicast), nprocs=2, join=True)
