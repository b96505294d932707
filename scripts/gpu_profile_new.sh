#!/bin/bash
# ncu --set full captures of the kernels added after the first profile pass: the fused whole-tensor fold (reduce_all), the nested
# convolution reduction (reduce_cols with epilogue), and the 128 / 64-wide contraction tiles (1 GPU).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cat > /tmp/prof_new.py <<'PY'
import sys
sys.path.insert(0, ".")
from compute.scala_b200 import cuda
cuda.init(0, streams=1)
T = cuda.Tensor
which = sys.argv[1]
if which == "fold":
    n = 16384
    a, b, c = (T.random([n, n], seed=s).doCache() for s in (1, 2, 3))
    t = a * b + c
    e = (T.tanh(T.log(T.exp(t) + a) * b) + c).sum()
    for _ in range(3): e.doBuffer().release()
elif which == "conv":
    sys.argv = ["x"]
    src = open("scripts/gpu_conv.py").read()
    exec(src[src.index("def convolute"):src.index("out = {}")])
    for (b, h, w, d, ks) in [(128, 32, 32, 8, 3), (64, 56, 56, 64, 3)]:
        e = convolute(T.randomNormal([b, h, w, d], seed=1).doCache(), T.randomNormal([ks, ks, d, d], seed=2).doCache(), T.randomNormal([d], seed=3).doCache())
        for _ in range(3): e.doBuffer().release()
elif which == "gemm":
    for (m, k, n) in [(1024, 1024, 1024), (2048, 2048, 2048), (4096, 4096, 4096)]:
        A, B = T.randomNormal([m, k], seed=9).doCache(), T.randomNormal([k, n], seed=10).doCache()
        a, b, c = A.doBuffer(), B.doBuffer(), cuda.Buffer.alloc(m * n)
        for _ in range(2): cuda.matmul_3xtf32(a, b, c, m, n, k)
cuda.synchronize()
PY
for cfg in fold conv gemm; do
  ncu --set full --clock-control none --import-source on -k regex:'reduce_all|reduce_cols|gemm_3xtf32' -f -o gpurun_out/r01b_$cfg python /tmp/prof_new.py $cfg > gpurun_out/ncu_$cfg.log 2>&1
  echo "$cfg rc=$?"
done
ls -la gpurun_out/*.ncu-rep
