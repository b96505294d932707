"""Row-sharded 8192^3 matmul + all-gather of C under torchrun: fused epilogue (per tile configuration) vs contraction + ncclAllGather."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from compute.scala_b200 import cuda, sharding

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cuda.init(local, streams=1)
comm = sharding.Communicator(cuda, dist)
T = cuda.Tensor
n = 8192
m = n // world
A = T.randomNormal([m, n], seed=9 + 16 * rank).doCache()
B = T.randomNormal([n, n], seed=10).doCache()
ab, bb = A.doBuffer(), B.doBuffer()


def measure(step, steps=10):
    for _ in range(3):
        step()
    cuda.synchronize()
    dist.barrier()
    cuda.timer_start()
    for _ in range(steps):
        step()
    ms = cuda.timer_stop() / steps
    t = torch.tensor([ms], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


out = {}
for cfg in ((None,) if "--quick" in sys.argv else (None, "512", "256", "128", "64")):
    if cfg:
        os.environ["CC_GEMM_FORCE_CONFIG"] = cfg
    else:
        os.environ.pop("CC_GEMM_FORCE_CONFIG", None)
    out[f"sharded only [{cfg or 'auto'}]"] = measure(lambda: comm.matmul_rows(ab, bb, m, n, n).release())
    out[f"fused gather [{cfg or 'auto'}]"] = measure(lambda: comm.matmul_rows(ab, bb, m, n, n, gather=True, fused=True).release())
os.environ.pop("CC_GEMM_FORCE_CONFIG", None)
out["contraction + ncclAllGather [auto]"] = measure(lambda: comm.matmul_rows(ab, bb, m, n, n, gather=True, fused=False).release())
if rank == 0:
    print(json.dumps({"world": world, "ms": out}), flush=True)
cuda.synchronize()
comm.close()
dist.destroy_process_group()
