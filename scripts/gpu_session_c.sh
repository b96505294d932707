#!/bin/bash
# round 2, third GPU session (one GPU): full GPU tier on the final defaults, bench, ncu launch list + full captures (C2 kernel, the CTA-pair
# contraction, the fused column second stage), SASS summary
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest=$?"; tail -6 gpurun_out/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke=$?"; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench=$?"; tail -c 1500 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launch_list_bench.csv \
    python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --short-side > gpurun_out/r02_ncu_list.log 2>&1; echo "ncu_list=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jit_kernel -s 3 -c 2 -f -o gpurun_out/r02_c2_full \
    python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-side-configs > gpurun_out/r02_ncu_c2.log 2>&1; echo "ncu_c2=$?"
cat > /tmp/prof_r02.py <<'PY'
import sys
sys.path.insert(0, ".")
from compute.scala_b200 import cuda
cuda.init(0, streams=1)
T = cuda.Tensor
which = sys.argv[1]
if which == "gemm":
    for (m, k, n) in [(8192, 8192, 8192)]:
        A, B = T.random([m, k], seed=9).doCache(), T.random([k, n], seed=10).doCache()
        a, b, c = A.doBuffer(), B.doBuffer(), cuda.Buffer.alloc(m * n)
        for _ in range(2): cuda.matmul_3xtf32(a, b, c, m, n, k)
elif which == "cols":
    x = T.random([16384, 4096], seed=5).doCache()
    parts = x.split(0)
    acc = parts[0]
    for p in parts[1:]: acc = acc + p
    for _ in range(3): acc.doBuffer().release()
cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_3xtf32' -s 1 -c 1 -f -o gpurun_out/r02_gemm_pair python /tmp/prof_r02.py gemm > gpurun_out/r02_ncu_gemm.log 2>&1; echo "ncu_gemm=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'reduce_cols' -s 1 -c 1 -f -o gpurun_out/r02_reduce_cols_fused python /tmp/prof_r02.py cols > gpurun_out/r02_ncu_cols.log 2>&1; echo "ncu_cols=$?"
ls -la gpurun_out/*.ncu-rep
