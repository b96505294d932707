#!/bin/bash
# ncu --set full captures of the dominant kernel of every config (1 GPU). Summaries are extracted on the CPU box.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cat > /tmp/prof_driver.py <<'PY'
import sys
sys.path.insert(0, ".")
from compute.scala_b200 import cuda
cuda.init(0, streams=1)
T = cuda.Tensor
which = sys.argv[1]
def axis(x, ax):
    parts = x.split(ax); acc = parts[0]
    for p in parts[1:]: acc = acc + p
    return acc
if which == "c3":
    x = T.random([16384, 16384], seed=5).doCache()
    for e in (x.sum(), axis(x, 0), axis(x, 1)):
        for _ in range(3): e.doBuffer().release()
elif which == "c4":
    t = T.random([512, 512, 512], seed=7).doCache(); m = T.random([512, 512], seed=8).doCache()
    for e in (t.permute([2, 0, 1]).translate([3, -5, 7]), T.join(t.split(1)), m.broadcast([512, 512, 512])):
        for _ in range(3): e.doBuffer().release()
elif which == "c5":
    n = 8192
    A, B = T.randomNormal([n, n], seed=9).doCache(), T.randomNormal([n, n], seed=10).doCache()
    a, b, c = A.doBuffer(), B.doBuffer(), cuda.Buffer.alloc(n * n)
    for _ in range(3): cuda.matmul_3xtf32(a, b, c, n, n, n)
elif which == "c1":
    a, b, c = (T.random([1024, 1024], seed=s).doCache() for s in (1, 2, 3))
    e = T.tanh(a * b + c)
    for _ in range(5): e.doBuffer().release()
cuda.synchronize()
PY
for cfg in c1 c3 c4 c5; do
  ncu --set full --clock-control none --import-source on -k regex:'jit_kernel|reduce_|gemm_3xtf32|split_' -f -o gpurun_out/r01_$cfg python /tmp/prof_driver.py $cfg > gpurun_out/ncu_$cfg.log 2>&1
  echo "$cfg rc=$?"
done
ls -la gpurun_out/*.ncu-rep
