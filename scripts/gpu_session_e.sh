#!/bin/bash
# eight GPUs: bench at N=8 through the sharded Tensor API (every leg verified before it is timed), per-kernel times of the sharded legs
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n8.txt 2>&1; nproc >> gpurun_out/r02_topo_n8.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_scale_n8.json 2> gpurun_out/r02_scale_n8.err; echo "bench8=$?"; tail -c 1500 gpurun_out/r02_scale_n8.json; tail -3 gpurun_out/r02_scale_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 scripts/gpu_scale_profile.py > gpurun_out/r02_scale_profile_n8.log 2>&1; echo "prof8=$?"; tail -8 gpurun_out/r02_scale_profile_n8.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_scale_n4.json 2> gpurun_out/r02_scale_n4.err; echo "bench4=$?"
