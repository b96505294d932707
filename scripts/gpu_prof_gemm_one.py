import sys
sys.path.insert(0, ".")
from compute.scala_b200 import cuda
cuda.init(0, streams=1)
T = cuda.Tensor
m = k = n = int(sys.argv[1])
A, B = T.random([m, k], seed=9).doCache(), T.random([k, n], seed=10).doCache()
a, b, c = A.doBuffer(), B.doBuffer(), cuda.Buffer.alloc(m * n)
for _ in range(3): cuda.matmul_3xtf32(a, b, c, m, n, k)
cuda.synchronize()
