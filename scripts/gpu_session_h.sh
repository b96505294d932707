#!/bin/bash
# one GPU, end of round 2: the whole GPU tier + smoke + both bench arms + launch list + the reference's JMH grid on the final kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest=$?"; tail -4 gpurun_out/r02_pytest_gpu_final.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke=$?"; tail -1 gpurun_out/r02_smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench=$?"; tail -3 gpurun_out/r02_bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launch_list_bench.csv \
    python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --short-side > gpurun_out/r02_ncu_list.log 2>&1; echo "ncu_list=$?"
timeout 600 python scripts/gpu_reference_grid.py > gpurun_out/r02_reference_grid.log 2>&1; echo "grid=$?"; tail -2 gpurun_out/r02_reference_grid.log | cut -c1-300
timeout 300 python scripts/gpu_gemm_v2.py > gpurun_out/r02_gemm_v2.log 2>&1; echo "gemm_v2=$?"
timeout 300 python scripts/gpu_gemm_splitk.py > gpurun_out/r02_gemm_splitk.log 2>&1; echo "splitk=$?"
