#!/bin/bash
# round 2, first GPU session (one GPU): full GPU tier, smoke, both bench arms, the opt-in lowerings A/B'd, launch-bound re-timing
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nproc > gpurun_out/r02_host.txt; lscpu | grep -E "Model name|Socket|NUMA|Thread" >> gpurun_out/r02_host.txt; nvidia-smi topo -m >> gpurun_out/r02_host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest=$?"; tail -8 gpurun_out/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke=$?"; tail -2 gpurun_out/r02_smoke.log
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref=$?"; cat gpurun_out/r02_bench_ref.json | cut -c1-600
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench=$?"; tail -c 3000 gpurun_out/r02_bench_n1.json; tail -5 gpurun_out/r02_bench_n1.err
for k in fuse_col red_p2 red_p4 batched; do timeout 600 python scripts/gpu_knob_ab.py $k > gpurun_out/r02_knob_$k.log 2>&1; echo "knob $k=$?"; tail -12 gpurun_out/r02_knob_$k.log; done
CC_TEST_BATCHED=1 timeout 600 python -m pytest tests/test_gemm.py -m gpu -q -k batched > gpurun_out/r02_pytest_batched.log 2>&1; echo "batched=$?"; tail -5 gpurun_out/r02_pytest_batched.log
timeout 300 python scripts/gpu_launch_bound.py > gpurun_out/r02_launch_bound.log 2>&1; echo "lb=$?"; tail -5 gpurun_out/r02_launch_bound.log
