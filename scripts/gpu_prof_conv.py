"""one evaluation of the reference's 3 x 3 / depth 8 / batch 128 convolution (benchmarks.scala:463-556, :612-630) for ncu"""
import sys
sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
from compute.scala_b200 import cuda
import gpu_knob_ab as ab
cuda.init(0, streams=1)
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 128
label, build, _ = [w for w in ab.workloads("tile_owner", cuda.Tensor) if w[0] == f"conv 3x3 batch {batch} 32x32 depth 8"][0]
e = build()
for _ in range(3):
    e.doBuffer().release()
cuda.synchronize()
