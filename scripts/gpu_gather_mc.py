"""C5 at N ranks: row-sharded contraction un-gathered, gathered from the epilogue (NVLS multicast stores or one store per peer, by CC_MULTICAST),
and gathered by ncclAllGather after the kernel. Every gathered result is checked on dataset E before timing. Rank 0 writes
gpurun_out/gather_mc_n<N>_<route>.json.
  CC_MULTICAST=1|0 python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/gpu_gather_mc.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from compute.scala_b200 import cuda, sharding  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cuda.init(local, streams=1)
comm = sharding.Communicator(cuda, dist)
T = cuda.Tensor
out = {"n_gpus": world, "route": "nvls multicast" if os.environ.get("CC_MULTICAST", "1") != "0" else "one TMA store per peer (CUDA IPC)"}
for n5 in [int(x) for x in os.environ.get("GATHER_SIZES", "8192,4096").split(",")]:
    m5 = n5 // world

    def e_tensor(shape, seed):
        r = T.random(shape, seed=seed)
        return ((r * T.fill(9.0, shape)) - (r * T.fill(9.0, shape)) % T.fill(1.0, shape) - T.fill(4.0, shape)).doCache()

    A = e_tensor([m5, n5], 9 + 16 * rank).shard()
    B = e_tensor([n5, n5], 10)
    hb = bench.np_dataset_e(n5 * n5, 10).reshape(n5, n5).astype(np.float64)
    c = comm.matmul_pattern(A, B)
    sample = np.r_[0:2, m5 // 2:m5 // 2 + 2, m5 - 2:m5]

    def check(whole):
        for o in range(world):
            ha = np.stack([bench.np_dataset_e(n5, 9 + 16 * o, first=int(r) * n5) for r in sample]).astype(np.float64)
            if not np.array_equal(whole[o * m5 + sample].astype(np.float64), ha @ hb):
                return False
        return True

    def timed(expr, steps=10):
        for _ in range(3):
            expr.doBuffer().release()
        cuda.synchronize()
        dist.barrier()
        cuda.timer_start()
        for _ in range(steps):
            expr.doBuffer().release()
        ms = cuda.timer_stop() / steps
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    rec = {"un-gathered ms": timed(c)}
    cg = c.gather(zero_copy=True)
    ok = check(cg.flatArray().reshape(n5, n5))
    rec["fused gather ms"] = timed(cg)
    rec["fused gather verified"] = ok
    arena = comm.gather_arena(m5 * n5 * world)
    comm.route_peer(False)
    cn = c.gather(zero_copy=True)
    ok2 = check(cn.flatArray().reshape(n5, n5))
    rec["ncclAllGather after the kernel ms"] = timed(cn)
    rec["nccl verified"] = ok2
    comm.route_peer(True)
    rec["TFLOP/s fused"] = 2 * n5**3 / rec["fused gather ms"] / 1e9
    out[f"{n5}^3"] = rec
    del A, B, c, cg, cn
t = torch.tensor([1 if all(v.get("fused gather verified", True) and v.get("nccl verified", True) for v in out.values() if isinstance(v, dict)) else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
out["verified on every rank"] = bool(t.item())
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = ("mc" if os.environ.get("CC_MULTICAST", "1") != "0" else "ipc") + os.environ.get("GATHER_TAG", "")
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"gather_mc_n{world}_{tag}.json"), "w"), indent=1)
    print(json.dumps(out))
cuda.synchronize()
comm.close()
dist.destroy_process_group()
