"""CPU-only: native unit tests of host-side helpers, built with AddressSanitizer + UBSan."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_smallvec_under_asan(tmp_path):
    """cc::SmallVec carries the launch path's short lists (argument buffers, hazard marks, kernel parameters): copy / move / growth /
    reuse after move, inline and on the heap, with the sanitizers watching"""
    exe = str(tmp_path / "smallvec_test")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-Wall", "-Wextra", "-Wno-self-move", "-Wno-self-assign-overloaded",
           "-I", os.path.join(ROOT, "compute", "scala_b200", "csrc"), os.path.join(HERE, "native", "smallvec_test.cpp"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1"))
    assert r.returncode == 0 and "smallvec ok" in r.stdout, (r.stdout, r.stderr[-3000:])
