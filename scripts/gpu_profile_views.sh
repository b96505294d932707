#!/bin/bash
# ncu --set full captures (one launch each) of the view / reduction templates tuned late in round 1: the 5-point stencil (L1-allocating
# loads), an odd-extent translation (scalar lanes, 8 per thread), the last-axis sum with a warp per row, the tanh x100 counted loop,
# join at dimension 0, and a max fold over an axis. 1 GPU.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cat > /tmp/prof_views.py <<'PY'
import sys
sys.path.insert(0, ".")
from compute.scala_b200 import cuda
cuda.init(0, streams=1)
T = cuda.Tensor
def chain(parts, f):
    acc = parts[0]
    for p in parts[1:]:
        acc = f(acc, p)
    return acc
N3 = [512, 512, 512]
x = T.random(N3, seed=1).doCache()
xo = T.random([1001, 1003, 127], seed=3).doCache()
xm = T.random([256, 512, 1024], seed=8).doCache()
xs = T.random([128, 128, 128], seed=9).doCache()
t100 = xs
for _ in range(100):
    t100 = T.tanh(t100)
cases = [
    x + x.translate([0, 0, 1]) + x.translate([0, 0, -1]) + x.translate([0, 1, 0]) + x.translate([0, -1, 0]),
    xo.translate([0, 0, 1]),
    chain(xm.split(2), lambda a, b: a + b),
    t100,
    T.join(x.split(0), 0),
    chain(xm.split(1), T.max),
]
for e in cases:
    for _ in range(2):
        e.doBuffer().release()
cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:'jit_kernel|reduce_rows|reduce_cols' -f -o gpurun_out/r01c_views python /tmp/prof_views.py > gpurun_out/ncu_views.log 2>&1
echo "rc=$?"
ncu -i gpurun_out/r01c_views.ncu-rep --page raw --csv > gpurun_out/r01c_views_raw.csv 2>/dev/null
ls -la gpurun_out/r01c_views*
