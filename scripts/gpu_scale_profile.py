"""Per-kernel device times of the sharded legs at N ranks, from the library's built-in command profiler (cc_profile_*: a timing-event pair
around every command on its own stream) — the multi-rank stand-in for an ncu launch list, which must not wrap a multi-rank command.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/gpu_scale_profile.py
Rank 0 writes gpurun_out/scale_profile_n<N>.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from compute.scala_b200 import cuda, sharding  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cuda.init(local, streams=1)
comm = sharding.Communicator(cuda, dist)
T = cuda.Tensor
rows = 16384 // world
x = T.random([rows, 16384], seed=5 + 16 * rank).doCache().shard()
total, col, rowg = x.sum(), comm.fold(x.split(0)), comm.fold(x.split(1)).gather()
n5 = 8192
A = T.random([n5 // world, n5], seed=9 + 16 * rank).doCache().shard()
B = T.random([n5, n5], seed=10).doCache()
c = comm.matmul_pattern(A, B)
cg = c.gather(zero_copy=True)
legs = [("C3 full sum", total), ("C3 axis-0", col), ("C3 axis-1 + gather", rowg), ("C5 sharded", c), ("C5 + fused all-gather", cg)]
for _, e in legs:
    for _ in range(3):
        e.doBuffer().release()
cuda.synchronize()
dist.barrier()
out = {}
for name, e in legs:
    cuda.profile(True)
    for _ in range(20):
        e.doBuffer().release()
    cuda.synchronize()
    rep = cuda.profile_report()
    cuda.profile(False)
    out[name] = [{k: r[k] for k in ("name", "count", "avg_us", "min_us", "max_us") if k in r} for r in rep]
    dist.barrier()
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"n_gpus": world, "how": __doc__.split("\n")[0], "legs": out}, open(os.path.join(ROOT, "gpurun_out", f"scale_profile_n{world}.json"), "w"), indent=1)
    for k, v in out.items():
        print(k, [(r.get("name", "")[:50], round(r.get("avg_us", 0), 1)) for r in v])
comm.close()
dist.destroy_process_group()
