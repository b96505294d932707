#!/bin/bash
# two GPUs: multi-GPU + contraction tests on the new kernels (A through tensor memory, 256-byte gather stores), bench at N=2
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multi_gpu.py tests/test_gemm.py tests/test_full_size.py tests/test_resources_and_errors.py -m gpu -q -x > gpurun_out/r02_pytest_gpu_d.log 2>&1; echo "pytest=$?"; tail -8 gpurun_out/r02_pytest_gpu_d.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_scale_n2.json 2> gpurun_out/r02_scale_n2.err; echo "bench2=$?"; tail -c 2500 gpurun_out/r02_scale_n2.json | cut -c1-2500; tail -3 gpurun_out/r02_scale_n2.err
