#!/bin/bash
# Run on the GPU box through gpurun: GPU tests, bench, ncu launch list and a full capture of the C2 kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest=$?"; tail -25 gpurun_out/pytest_gpu.log
python bench.py --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench=$?"; tail -c 6000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --short-side > gpurun_out/ncu_bench.log 2>&1; echo "ncu_list=$?"
  ncu --set full --clock-control none --import-source on -k regex:jit_kernel -s 3 -c 2 -f -o gpurun_out/c2_full \
      python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-side-configs > gpurun_out/ncu_full.log 2>&1; echo "ncu_full=$?"
  ls -la gpurun_out
fi
