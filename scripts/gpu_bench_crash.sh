#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; "$@" > gpurun_out/crash_$name.out 2> gpurun_out/crash_$name.err; echo "$name rc=$? out=$(wc -c < gpurun_out/crash_$name.out) err_tail: $(grep -v '^  File \"/opt' gpurun_out/crash_$name.err | tail -12 | cut -c1-160 | tr '\n' '|')"; }
A="bench.py --steps 5 --warmup 3 --no-cpu-baseline --short-side"
run plain python $A
run fh python -X faulthandler $A
run dbg env CC_GRAPH_DEBUG=1 python $A
run nopdl env CC_PDL=0 python -X faulthandler $A
run noside python -X faulthandler $A --no-side-configs
ulimit -c 0
run gdbbt gdb -batch -ex run -ex bt --args python $A
