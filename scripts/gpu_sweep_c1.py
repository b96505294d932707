"""C1 (tanh(a*b+c), 1024^2, L2-resident) step time over the elementwise template's knobs. Run on the GPU box."""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from compute.scala_b200 import cuda  # noqa: E402

cuda.init(0, streams=1)
T = cuda.Tensor
n = 1024
a, b, c = (T.random([n, n], seed=s).doCache() for s in (1, 2, 3))
for U, mb, gm in itertools.product(("1", "2", "4"), ("0", "4", "8"), ("", "8", "4", "2", "1")):
    os.environ["CC_TUNE_U"] = U
    os.environ["CC_TUNE_MIN_BLOCKS"] = mb
    if gm:
        os.environ["CC_TUNE_GRID_MULT"] = gm
    else:
        os.environ.pop("CC_TUNE_GRID_MULT", None)
    cuda.kernel_cache_clear()
    e = T.tanh(a * b + c)
    for _ in range(20):
        e.doBuffer().release()
    cuda.synchronize()
    cuda.timer_start()
    for _ in range(500):
        e.doBuffer().release()
    us = cuda.timer_stop() / 500 * 1000
    print(f"U={U} min_blocks={mb} grid_mult={gm or 'inf'}: {us:.2f} us/step  {16 * n * n / us / 1e3:.0f} GB/s", flush=True)
