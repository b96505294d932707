"""TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the reference's hot path.

Nothing in the product package (`compute/scala_b200`) may import this package.
Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` use it, and only as the checker / the reported CPU baseline.

Parity status: the reference (Scala + LWJGL OpenCL on POCL) cannot be executed in
this image (no JVM, no OpenCL ICD), so the oracle is a *restatement*.  It is pinned
against every golden vector the reference's own tests hold for the path
(`tests/test_oracle_goldens.py`, source lines cited there).  For tanh/exp/log the
reference itself pins nothing (the arithmetic lives in the OpenCL driver's libm,
unpinned): **parity unpinned** for those three functions beyond "<= 2 ulp from the
correctly-rounded fp32 result".
"""
