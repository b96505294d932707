#!/bin/bash
# compute-sanitizer over the GPU parity tests (every precompiled kernel and every NVRTC-generated template the tests reach),
# minus the BASELINE-size cases (minutes each under instrumentation). Run through gpurun; summaries land in gpurun_out/.
#   scripts/gpu_sanitizer.sh memcheck | racecheck | synccheck | initcheck
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TOOL=${1:-memcheck}
SKIP='not full_size and not beyond_2_31 and not 8192 and not 2_pow_28 and not 16384 and not 512_cubed and not concurrent_callers and not c_consumer and not pool_gives'
FILES="tests/test_cuda_goldens.py tests/test_edge_cases.py tests/test_parity_configs.py tests/test_gemm.py tests/test_threads_and_events.py tests/test_fuzz_differential.py tests/test_golden_fixtures.py"
timeout ${SANITIZER_TIMEOUT:-1500} compute-sanitizer --tool "$TOOL" --error-exitcode 99 --print-limit 20 \
    --log-file "gpurun_out/sanitizer_${TOOL}.log" \
    python -m pytest $FILES -q -m gpu -k "$SKIP" -x -p no:cacheprovider > "gpurun_out/sanitizer_${TOOL}_pytest.log" 2>&1
echo "rc=$?"
tail -5 "gpurun_out/sanitizer_${TOOL}_pytest.log"
grep -c "=========" "gpurun_out/sanitizer_${TOOL}.log"
tail -15 "gpurun_out/sanitizer_${TOOL}.log"
