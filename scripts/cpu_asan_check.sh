#!/bin/bash
# Host-side memory-safety check without a GPU: builds the C++ half of libcompute_cuda.so with AddressSanitizer into a scratch copy of the
# package, then runs (a) every driver-spy scenario of tests/driver_spy (normal paths, eight threads, injected driver failures) and (b) the
# compile-only graph fuzzer (default plans and each opt-in lowering) against it. libstdc++ is preloaded next to libasan so that C++
# exceptions work inside a Python process. Prints one line per run; any "AddressSanitizer" line is a finding.
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
W="${1:-/tmp/compute_cuda_asan}"
rm -rf "$W" && mkdir -p "$W/pkg/compute" "$W/pkg/tests" && cd "$W/pkg"
python -c "import sys; sys.path.insert(0, '$ROOT'); from compute.scala_b200 import build; build.build()" || exit 1
cp -r "$ROOT/compute/scala_b200" compute/ && cp "$ROOT/compute/__init__.py" compute/ 2>/dev/null
cp -r "$ROOT/include" "$ROOT/oracle" . && cp -r "$ROOT/tests/driver_spy" "$ROOT/tests/test_fuzz_differential.py" "$ROOT/tests/test_fuzz_compile_only.py" tests/
cd compute/scala_b200 && rm -f ./*.so
for f in ir codegen driver runtime tensor; do
  g++ -std=c++17 -O1 -g -fsanitize=address -fno-omit-frame-pointer -fPIC -I/usr/local/cuda/include -c csrc/$f.cpp -o build/$f.asan.o &
done
wait
nvcc -shared -o libcompute_cuda.so build/{ir,codegen,driver,runtime,tensor}.asan.o build/kernels_basic.cu.o build/gemm_3xtf32.cu.o \
  -gencode arch=compute_100a,code=sm_100a -cudart static -lnvrtc -ldl -lpthread -Xlinker -rpath,/usr/local/cuda/lib64 -Xcompiler -fsanitize=address || exit 1
gcc -O2 -std=c11 -shared -fPIC -I"$(python -c 'import sysconfig; print(sysconfig.get_paths()["include"])')" csrc/py_hotcalls.c \
  -o "_hotcalls$(python -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')" -L. -lcompute_cuda '-Wl,-rpath,$ORIGIN' || exit 1
cd "$W/pkg"
mkdir -p "$W/spy" && gcc -O2 -shared -fPIC -I/usr/local/cuda/include tests/driver_spy/spy_libcuda.c -o "$W/spy/libcuda.so.1" || exit 1
PRE="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libstdc++.so)"
run() { LD_PRELOAD="$PRE" ASAN_OPTIONS=detect_leaks=0 "$@" 2>&1 | grep -E "AddressSanitizer|SUMMARY|Traceback" | head -3; }
for sc in steady_loop independent_rotate first_use_of_uploaded_inputs read_back read_back_values two_launch_plan_and_fold structural_cache \
          balance_on_shutdown threads compile_does_not_block_launches same_structure_from_many_threads \
          "faults launch" "faults module" "faults alloc_once" "faults alloc_always" "faults d2h" "faults sync"; do
  echo "spy scenario: $sc $(run env LD_LIBRARY_PATH="$W/spy" python tests/driver_spy/scenarios.py $sc | tr '\n' ' ')"
done
cat > "$W/fuzz.py" <<PY
import sys
sys.path.insert(0, "$W/pkg"); sys.path.insert(0, "$W/pkg/tests")
import numpy as np
from compute.scala_b200 import cuda
assert "$W" in cuda._L()._name
from test_fuzz_compile_only import LazyGen, DIMS_WIDE
n = 0
for seed in range(91000, 91120):
    gen = LazyGen(cuda, seed, dims=DIMS_WIDE if seed % 2 else (1, 2, 3, 4, 5, 8, 12), max_rank=2 if seed % 2 else 4)
    p = gen.expr(depth=2)
    if int(np.prod(p.shape)) <= 2_000_000:
        p.g.compile(); n += 1
print("compiled", n, "random graphs")
PY
for knob in "CC_NOOP=1" "CC_FUSE_COL_STAGE=1" "CC_BATCHED_CONTRACTION=1 CC_TUNE_CONTRACTION_MIN_MACS=1" "CC_PDL=0"; do
  echo "compile fuzz [$knob]: $(LD_PRELOAD="$PRE" ASAN_OPTIONS=detect_leaks=0 env $knob python "$W/fuzz.py" 2>&1 | grep -E "AddressSanitizer|SUMMARY|Traceback|compiled" | head -3 | tr '\n' ' ')"
done

# ---- ThreadSanitizer over the multi-threaded scenarios (the runtime lock, the in-flight compile table, atomics on handles) ----------------
cd "$W/pkg/compute/scala_b200"
for f in ir codegen driver runtime tensor; do
  g++ -std=c++17 -O1 -g -fsanitize=thread -fno-omit-frame-pointer -fPIC -I/usr/local/cuda/include -c csrc/$f.cpp -o build/$f.tsan.o &
done
wait
nvcc -shared -o libcompute_cuda.so build/{ir,codegen,driver,runtime,tensor}.tsan.o build/kernels_basic.cu.o build/gemm_3xtf32.cu.o \
  -gencode arch=compute_100a,code=sm_100a -cudart static -lnvrtc -ldl -lpthread -Xlinker -rpath,/usr/local/cuda/lib64 -Xcompiler -fsanitize=thread || exit 1
cd "$W/pkg"
PRE="$(gcc -print-file-name=libtsan.so) $(gcc -print-file-name=libstdc++.so)"
for sc in threads same_structure_from_many_threads compile_does_not_block_launches; do
  echo "tsan scenario: $sc $(LD_PRELOAD="$PRE" TSAN_OPTIONS=report_signal_unsafe=0 LD_LIBRARY_PATH="$W/spy" python tests/driver_spy/scenarios.py $sc 2>&1 | grep -E "ThreadSanitizer|SUMMARY|Traceback" | sort | uniq -c | head -5 | tr '\n' ' ')"
done
