"""CPU-only: which driver calls the runtime turns a command into.  A driver spy (tests/driver_spy, test infrastructure only: it counts
calls and executes nothing) stands in front of libcuda for a child process that drives the public API; the counters are the assertion.
This is the launch path's logic -- stream affinity, hazard events, pooled memory, the programmatic-dependent-launch attribute, the
structural kernel cache, direct-to-host small results, resource balance on shutdown -- checked where no GPU exists.  No values are
computed (kernels never run), so nothing here is a parity or performance claim."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SPY_SRC = os.path.join(HERE, "driver_spy", "spy_libcuda.c")
SCENARIOS = os.path.join(HERE, "driver_spy", "scenarios.py")


@pytest.fixture(scope="module")
def spy_dir():
    key = hashlib.sha1(open(SPY_SRC, "rb").read()).hexdigest()[:16]
    d = os.path.join(tempfile.gettempdir(), "compute_cuda_driver_spy_" + key)
    so = os.path.join(d, "libcuda.so.1")
    if not os.path.exists(so):
        os.makedirs(d, exist_ok=True)
        cmd = ["gcc", "-O2", "-Wall", "-Werror", "-shared", "-fPIC", "-I/usr/local/cuda/include", SPY_SRC, "-o", so + ".tmp"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        os.replace(so + ".tmp", so)
    return d


def run(spy_dir, scenario, *args, **env):
    e = dict(os.environ, LD_LIBRARY_PATH=spy_dir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""), **env)
    e.pop("CC_KERNEL_CACHE_DIR", None)
    r = subprocess.run([sys.executable, SCENARIOS, scenario, *args], env=e, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_a_steady_loop_is_one_launch_per_step_and_nothing_else(spy_dir):
    """BASELINE config 1 in miniature: the output block returns from the pool carrying the previous step's mark, the step follows it
    onto the same stream (no event record / wait), the kernel is launched with the dependent-launch attribute, memory comes from the pool"""
    r = run(spy_dir, "steady_loop")
    assert r["pdl_launches"] == 100 and r["plain_launches"] == 0 and r["cuLaunchKernelEx"] == 100
    assert r["streams_launched_on"] == 1
    assert r["cuEventRecord"] == 0 and r["cuStreamWaitEvent"] == 0
    assert r["cuMemAlloc"] == 0 and r["alloc_calls"] == 100 and r["pool_hits"] == 100
    assert r["compiles"] == 0 and r["cuModuleLoadData"] == 0
    assert r["cuCtxSetCurrent"] <= 200


def test_pdl_can_be_switched_off(spy_dir):
    r = run(spy_dir, "steady_loop", CC_PDL="0")
    assert r["pdl_launches"] == 0 and r["plain_launches"] == 100
    assert r.get("cuLaunchKernelEx", 0) == 0 and r["cuLaunchKernel"] == 100
    assert r["cuEventRecord"] == 0 and r["cuStreamWaitEvent"] == 0


def test_independent_commands_still_rotate_over_the_streams(spy_dir):
    """affinity follows HOT hazards only: eight independent expressions over one long-uploaded input spread over the four compute
    streams (the reference's several queues per device), each stream paying its one wait for the copy stream"""
    r = run(spy_dir, "independent_rotate")
    assert r["pdl_launches"] == 8
    assert r["streams_launched_on"] == 4
    assert r["cuEventRecord"] == 4 and r["cuStreamWaitEvent"] == 4


def test_hazards_on_uploaded_inputs_cost_one_wait_once(spy_dir):
    r = run(spy_dir, "first_use_of_uploaded_inputs")
    # three inputs written by the H2D stream: one event covers all of them
    assert r["first"]["cuEventRecord"] == 1 and r["first"]["cuStreamWaitEvent"] == 1 and r["first"]["pdl_launches"] == 1
    assert r["second"]["cuEventRecord"] == 0 and r["second"]["cuStreamWaitEvent"] == 0 and r["second"]["pdl_launches"] == 1
    assert r["second"]["cuMemAlloc"] == 0


def test_small_results_are_stored_by_the_kernel_large_ones_are_copied(spy_dir):
    r = run(spy_dir, "read_back")
    assert r["small"]["pdl_launches"] == 1 and r["small"]["cuMemcpyDtoHAsync"] == 0 and r["small"]["cuMemHostAlloc"] == 0
    assert r["big"]["pdl_launches"] == 1 and r["big"]["cuMemcpyDtoHAsync"] == 1 and r["big"]["cuMemHostAlloc"] == 0
    assert r["small"]["cuEventSynchronize"] == 1 and r["big"]["cuEventSynchronize"] == 1


@pytest.mark.parametrize("binding", ["native", "ctypes"])
def test_read_backs_land_in_the_callers_memory(spy_dir, binding):
    """flatArray / flatBuffer / flatArrayInto through the native binding: dtype, length, placement (the spy's copies deliver 42.0f),
    release semantics, the empty tensor, and a typed error from graph construction"""
    r = run(spy_dir, "read_back_values", **({"CC_PY_NO_HOTCALLS": "1"} if binding == "ctypes" else {}))
    assert r["native_binding"] is (binding == "native")
    assert r["flat_array"] == ["float32", [76800], 42.0, 42.0]
    assert r["flat_buffer"] == ["float32", [76800], 42.0, 42.0, 76800]
    assert r["released"] is True
    assert r["into"] == [42.0, 42.0, -1.0]  # nothing written past the requested floats
    assert r["small"] == ["float32", [16]]
    assert r["empty"] == [[0], 0]
    assert r["typed_error"] is True


def test_multi_launch_plans_and_folds_stay_on_their_stream(spy_dir):
    r = run(spy_dir, "two_launch_plan_and_fold", CC_FUSE_COL_STAGE="0")  # the separate second stage (the default fuses it, next test)
    assert r["axis_launches_per_step"] == 2
    # the second launch of each step reads the partials the first one just wrote: plain stream order, not an early-resident (PDL) grid
    assert r["axis"]["pdl_launches"] == 50 and r["axis"]["plain_launches"] == 50
    assert r["axis"]["cuEventRecord"] == 0 and r["axis"]["cuStreamWaitEvent"] == 0 and r["axis"]["cuMemAlloc"] == 0
    assert r["fold"]["pdl_launches"] == 50 and r["fold"]["cuEventRecord"] == 0 and r["fold"]["cuStreamWaitEvent"] == 0
    assert r["fold"]["cuMemsetD32Async"] == 0  # the fold's block counter resets itself


def test_fused_second_stage_is_one_launch_and_its_counters_are_given_back(spy_dir):
    """the default since round 2 (kernel side: tests/test_kernel_emulation.py): one launch per step, the per-stream block counters are
    allocated and cleared once, and freed at shutdown"""
    r = run(spy_dir, "two_launch_plan_and_fold")
    assert r["axis_launches_per_step"] == 1
    assert r["axis"]["pdl_launches"] == 50 and r["axis"]["cuMemAlloc"] == 0 and r["axis"]["cuMemsetD32Async"] == 0
    assert r["axis"]["cuEventRecord"] == 0 and r["axis"]["cuStreamWaitEvent"] == 0
    r = run(spy_dir, "balance_on_shutdown")
    assert r["cuMemAlloc"] == r["cuMemFree"] > 0 and r["live_tensors"] == 0


def test_pdl_is_not_used_for_a_kernel_that_reads_what_the_previous_command_wrote(spy_dir):
    """ld.global.nc requires read-only data for the grid's whole lifetime, and a PDL grid's lifetime starts while its predecessor runs"""
    r = run(spy_dir, "pdl_only_when_inputs_are_settled")
    assert r["settled"]["pdl_launches"] == 20 and r["settled"]["plain_launches"] == 0
    assert r["chained"]["pdl_launches"] <= 1 and r["chained"]["plain_launches"] >= 19


def test_a_captured_sequence_is_one_graph_launch_per_replay(spy_dir):
    """cc_graph_begin / cc_graph_end / cc_graph_launch: the 20 + 1 evaluations are recorded between cuStreamBeginCapture and
    cuStreamEndCapture on ONE stream with no event traffic; each replay is a single cuGraphLaunch and launches no kernel by itself"""
    r = run(spy_dir, "graph_capture_and_replay")
    cap, rep = r["captured"], r["replay"]
    assert cap["cuStreamBeginCapture"] == 1 and cap["cuStreamEndCapture"] == 1 and cap["cuGraphInstantiate"] == 1
    assert cap["pdl_launches"] == 21 and cap["streams_launched_on"] == 1
    assert cap.get("cuEventRecord", 0) == 0 and cap.get("cuStreamWaitEvent", 0) == 0 and cap.get("cuMemcpyDtoHAsync", 0) == 0
    # the capture's own pool: ONE block for the 20 elementwise outputs, the column sums and their partials; plus the runtime's lazily created
    # fold scratch + counter, which cc_graph_begin sets up before the capture starts (a memset + synchronise cannot be captured)
    assert cap.get("cuMemAlloc", 0) <= 5
    assert r["refused_copy"] is True
    assert r["commands"] == 21 and r["kernels_counted_while_capturing"] == 0 and r["kernels_counted_by_replays"] == 5 * 21
    assert rep["cuGraphLaunch"] == 5 and rep["pdl_launches"] == 0 and rep["plain_launches"] == 0
    assert r["launch_after"] == 1


def test_structurally_equal_expressions_share_one_module(spy_dir):
    r = run(spy_dir, "structural_cache")
    assert r["compiles"] == 2 and r["cache_hits"] == 1
    assert r["cuModuleLoadData"] == 2 and r["pdl_launches"] == 3


def test_every_driver_resource_is_given_back_on_shutdown(spy_dir):
    r = run(spy_dir, "balance_on_shutdown")
    assert r["live_tensors"] == 0
    assert r["bytes_in_use_before_shutdown"] <= 8192  # the runtime's own fold scratch
    assert r["cuMemAlloc"] == r["cuMemFree"] > 0
    assert r["cuMemHostAlloc"] == r["cuMemFreeHost"] > 0
    assert r["cuStreamCreate"] == r["cuStreamDestroy"] > 0
    assert r["cuEventCreate"] == r["cuEventDestroy"]
    assert r["cuModuleLoadData"] == r["cuModuleUnload"] > 0
    assert r["cuDevicePrimaryCtxRetain"] == r["cuDevicePrimaryCtxRelease"] == 1


def test_eight_threads_share_the_runtime(spy_dir):
    r = run(spy_dir, "threads")
    assert r["failures"] == [] and r["hung"] == 0
    assert r["launches"] == 8 * 200 * 2 and r["pdl_launches"] == 8 * 200 * 2
    assert r["compiles"] == 8  # one structure per thread, compiled once
    assert r["bytes_in_use_delta"] == 0
    assert r["live_tensors"] == 4  # a, b, a * b, shared


def test_the_jit_runs_outside_the_runtime_lock(spy_dir):
    """a thread stepping a cached plan is not held up by another thread's NVRTC compilations (Tensors.scala:1321-1329 builds programs
    asynchronously too); a structure requested by eight threads at once is compiled once"""
    r = run(spy_dir, "compile_does_not_block_launches")
    assert min(r["compile_ms"]) > 20.0, r  # the compilations were real (six of them, back to back)
    # held up by each compilation, the stepper would get in a handful of steps; free of them it makes hundreds of thousands
    assert r["steps"] > 1000, r
    assert r["worst_step_ms"] < max(r["compile_ms"]), r  # (loose on purpose: a loaded machine can deschedule the thread for a while)
    r = run(spy_dir, "same_structure_from_many_threads")
    assert r["failures"] == []
    assert r["compiles"] == 1 and r["nvrtc_compiles"] == 1 and r["cuModuleLoadData"] == 1
    assert r["launches"] == 16 and r["pdl_launches"] == 16


# ---- fault injection: a driver call fails in the middle of the work (checkErrorCode -> typed exception, OpenCL.scala:251-312) -------------
CC_ERR_CUDA, CC_ERR_OUT_OF_MEMORY = -4, -9


def _balanced(r, failed_allocs=0):
    assert r["live_tensors"] == 0
    assert r["bytes_in_use_after"] == r["bytes_in_use_base"], "a failed command leaked device memory"
    assert r["after_ok"] == 5, "the runtime was not usable after the failure"
    assert r["cuMemAlloc"] - failed_allocs == r["cuMemFree"]
    assert r.get("cuMemHostAlloc", 0) == r.get("cuMemFreeHost", 0)
    assert r["cuEventCreate"] == r["cuEventDestroy"]
    assert r["cuStreamCreate"] == r["cuStreamDestroy"]


def test_a_failed_launch_is_a_typed_error_and_leaks_nothing(spy_dir):
    r = run(spy_dir, "faults", "launch")
    assert r["ok"] == 29 and len(r["errors"]) == 1
    status, msg = r["errors"][0]
    assert status == CC_ERR_CUDA and "cuLaunchKernelEx" in msg and "(719)" in msg
    _balanced(r)
    assert r["cuModuleLoadData"] == r["cuModuleUnload"]


def test_a_failed_module_load_is_retried_by_the_next_evaluation(spy_dir):
    r = run(spy_dir, "faults", "module")
    assert r["ok"] == 4 and len(r["errors"]) == 1 and r["errors"][0][0] == CC_ERR_CUDA and "cuModuleLoadData" in r["errors"][0][1]
    _balanced(r)
    assert r["cuModuleLoadData"] - 1 == r["cuModuleUnload"]  # the failed load produced no module


def test_out_of_memory_trims_the_pool_and_retries(spy_dir):
    r = run(spy_dir, "faults", "alloc_once")
    assert r["errors"] == [] and r["ok"] == 6  # the caller never saw it (pool trimmed, allocation retried)
    _balanced(r, failed_allocs=1)


def test_persistent_out_of_memory_is_reported_as_such(spy_dir):
    r = run(spy_dir, "faults", "alloc_always")
    assert r["ok"] >= 1 and len(r["errors"]) >= 4
    assert all(st == CC_ERR_OUT_OF_MEMORY and "CUDA_ERROR_OUT_OF_MEMORY" in msg for st, msg in r["errors"])
    _balanced(r, failed_allocs=2 * len(r["errors"]))  # each failed evaluation tried twice (before and after trimming the pool)


@pytest.mark.parametrize("case,entry", [("d2h", "cuMemcpyDtoHAsync"), ("sync", "cuEventSynchronize")])
def test_a_failed_read_back_is_a_typed_error_and_leaks_nothing(spy_dir, case, entry):
    r = run(spy_dir, "faults", case)
    assert r["ok"] == 7 and len(r["errors"]) == 1 and r["errors"][0][0] == CC_ERR_CUDA and entry in r["errors"][0][1]
    _balanced(r)


def test_heap_allocations_per_launch_stay_low(spy_dir, tmp_path):
    """a launch-bound step is ~3 us, of which the library's own share was mostly malloc / free (14 allocations per launch before the
    short lists moved into cc::SmallVec): a regression guard on the counts, which are deterministic (C driver program, malloc counted by
    an LD_PRELOAD shim, the driver spy in front of libcuda)"""
    root = os.path.dirname(HERE)
    libdir = os.path.join(root, "compute", "scala_b200")
    shim, exe = str(tmp_path / "libmalloc_count.so"), str(tmp_path / "host_cost")
    for cmd in (["gcc", "-O2", "-shared", "-fPIC", os.path.join(HERE, "driver_spy", "malloc_count.c"), "-o", shim, "-ldl"],
                ["gcc", "-O2", "-std=c11", "-D_POSIX_C_SOURCE=200809L", "-I", os.path.join(root, "include"), os.path.join(HERE, "driver_spy", "host_cost.c"), "-o", exe,
                 "-L", libdir, "-lcompute_cuda", "-ldl", f"-Wl,-rpath,{libdir}"]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    env = dict(os.environ, LD_PRELOAD=shim, LD_LIBRARY_PATH=spy_dir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    env.pop("CC_KERNEL_CACHE_DIR", None)
    r = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.stdout, r.stderr[-2000:])
    counts = json.loads(r.stdout.strip().splitlines()[-1])
    assert counts["mallocs_per_steady_step"] <= 3.0, counts       # measured: 2 (the Buffer object and its registry node)
    assert counts["mallocs_per_fresh_expression"] <= 42.0, counts  # measured: 36 (16 for three nodes, 17 for the first evaluation, 3 to launch and release)
