#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench=$?"; tail -c 800 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
CC_NC_LOADS=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_n1_nc_loads.json 2> gpurun_out/r02_bench_n1_nc.err; echo "bench_nc=$?"
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launch_list_bench.csv \
    python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --short-side > gpurun_out/r02_ncu_list.log 2>&1; echo "ncu_list=$?"
timeout 600 python -m pytest tests/test_threads_and_events.py -m gpu -q -x 2>&1 | tail -2
