#!/bin/bash
# round 2, second GPU session (two GPUs): full GPU tier incl. the 2-GPU sharding tests, bench at N=2 through the sharded Tensor API
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n2.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_n2.log 2>&1; echo "pytest=$?"; tail -12 gpurun_out/r02_pytest_gpu_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_scale_n2.json 2> gpurun_out/r02_scale_n2.err; echo "bench2=$?"; tail -c 5000 gpurun_out/r02_scale_n2.json; tail -5 gpurun_out/r02_scale_n2.err
