#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python scripts/gpu_graph_debug.py 2>&1 | grep "sum\|keep" | cut -c1-200
timeout 900 python -m pytest tests/test_threads_and_events.py tests/test_resources_and_errors.py tests/test_parity_configs.py tests/test_cuda_goldens.py tests/test_edge_cases.py -m gpu -q -x > gpurun_out/r02_pytest_gpu_c2.log 2>&1; echo "pytest=$?"; tail -6 gpurun_out/r02_pytest_gpu_c2.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench=$?"; tail -c 1200 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
CC_NC_LOADS=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_n1_nc_loads.json 2> gpurun_out/r02_bench_n1_nc.err; echo "bench_nc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launch_list_bench.csv \
    python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --short-side > gpurun_out/r02_ncu_list.log 2>&1; echo "ncu_list=$?"
