"""Regenerates tests/golden/oracle_fixtures.npz from the CPU oracle (oracle/reference.py), which is itself pinned to the golden
vectors of the reference's own test-suite (tests/golden/reference_goldens.json, tests/test_oracle_goldens.py).  The reference (Scala
+ OpenCL) cannot run in this image, so these are oracle outputs, not reference outputs: they freeze the oracle's bits so that neither
the oracle nor the CUDA path can drift unnoticed.  For "ulp2" cases both oracle variants are stored (FP_CONTRACT off / on).

    python tests/golden/make_fixtures.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from cases import CASES  # noqa: E402

from oracle import reference as ref  # noqa: E402


def evaluate(build, contract):
    ref.CONTRACT[0] = contract
    try:
        t = build(ref.Tensor)
        return np.asarray(t.shape, np.int64), t.flat_array()
    finally:
        ref.CONTRACT[0] = False


def main():
    out = {}
    for name, (build, bar) in CASES.items():
        shape, strict = evaluate(build, False)
        out[name + "/shape"] = shape
        out[name + "/strict"] = strict
        if bar == "ulp2":
            out[name + "/contracted"] = evaluate(build, True)[1]
    np.savez_compressed(os.path.join(HERE, "oracle_fixtures.npz"), **out)
    print(f"{len(CASES)} cases, {sum(v.nbytes for v in out.values())} bytes of values")


if __name__ == "__main__":
    main()
