#!/bin/bash
# sweep of the elementwise template's knobs on the C2 workload (device-resident value only)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for U in 2 4 8; do for G in 4 8 16 64 100000; do for MB in 0 6 8; do
  r=$(CC_TUNE_U=$U CC_TUNE_GRID_MULT=$G CC_TUNE_MIN_BLOCKS=$MB python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-side-configs --e2e-steps 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],4))")
  echo "U=$U grid_mult=$G min_blocks=$MB -> $r"
done; done; done
