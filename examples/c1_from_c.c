/* A plain-C consumer of include/compute_cuda.h: BASELINE config 1, tanh(a*b+c) on 1024x1024, built and evaluated through the
 * drop-in boundary exactly as a JVM caller would through LWJGL (scalars and pointers only, no C++ types, no Python).
 *
 *   gcc -O2 -std=c11 -D_POSIX_C_SOURCE=200809L -Iinclude examples/c1_from_c.c -o /tmp/c1_from_c -Lcompute/scala_b200 -lcompute_cuda -Wl,-rpath,$PWD/compute/scala_b200 -lm
 *   /tmp/c1_from_c [steps]
 *
 * Without a GPU it must fail loudly (CC_ERR_NO_DRIVER) — tests/test_abi_and_codegen.py builds and runs it for exactly that.
 * On a B200 it checks the result against libm's tanhf and reports the per-step time of the device-resident loop driven from C
 * (launch-rate bound: 16 MiB of L2-resident traffic per step) and of the read-back variants. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "compute_cuda.h"

#define CHECK(call)                                                                          \
  do {                                                                                       \
    int st_ = (call);                                                                        \
    if (st_ != CC_OK) {                                                                      \
      fprintf(stderr, "%s -> %d: %s\n", #call, st_, cc_last_error());                        \
      return st_ == CC_ERR_NO_DRIVER ? 3 : 1;                                                \
    }                                                                                        \
  } while (0)

static double now_us(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

int main(int argc, char** argv) {
  const int steps = argc > 1 ? atoi(argv[1]) : 20000;
  printf("%s\n", cc_version());
  CHECK(cc_init(-1));
  const int32_t shape[2] = {1024, 1024};
  const uint64_t n = 1024ull * 1024ull;
  ct_tensor a, b, c, ab, abc, r, ca, cb, ccached;
  CHECK(ct_random(shape, 2, 1, 0.f, &a));
  CHECK(ct_random(shape, 2, 2, 0.f, &b));
  CHECK(ct_random(shape, 2, 3, 0.f, &c));
  CHECK(ct_do_cache(a, &ca)); /* Tensor.random(...).doCache (T:642-666): inputs resident in HBM */
  CHECK(ct_do_cache(b, &cb));
  CHECK(ct_do_cache(c, &ccached));
  CHECK(ct_binary(CT_TIMES, ca, cb, &ab));
  CHECK(ct_binary(CT_PLUS, ab, ccached, &abc));
  CHECK(ct_unary(CT_TANH, abc, &r));

  /* correctness: against libm on the same inputs */
  float *ha = malloc(n * 4), *hb = malloc(n * 4), *hc = malloc(n * 4), *hr = malloc(n * 4);
  CHECK(ct_flat_array(ca, ha, n));
  CHECK(ct_flat_array(cb, hb, n));
  CHECK(ct_flat_array(ccached, hc, n));
  CHECK(ct_flat_array(r, hr, n));
  double worst = 0;
  for (uint64_t i = 0; i < n; ++i) {
    const double want = tanh((double)ha[i] * (double)hb[i] + (double)hc[i]);
    const double err = fabs((double)hr[i] - want) / (fabs(want) > 1e-30 ? fabs(want) : 1e-30);
    if (err > worst) worst = err;
  }
  printf("max relative error vs fp64 libm: %.3g (2 ulp of fp32 = %.3g)\n", worst, 2 * 1.1920929e-7);
  if (worst > 4 * 1.1920929e-7) return 2;

  /* device-resident loop driven from C: evaluate, drop the buffer (doBuffer + release) */
  cc_buffer out;
  for (int i = 0; i < 2000; ++i) {
    CHECK(ct_do_buffer(r, &out, NULL));
    CHECK(cc_buffer_release(out));
  }
  CHECK(cc_synchronize());
  CHECK(cc_timer_start());
  double t0 = now_us();
  for (int i = 0; i < steps; ++i) {
    CHECK(ct_do_buffer(r, &out, NULL));
    CHECK(cc_buffer_release(out));
  }
  double t_submit = now_us() - t0;
  float ms = 0;
  CHECK(cc_timer_stop(&ms));
  printf("C1 device-resident from C: %.2f us/step on the device (%.0f GB/s of 16 MiB algorithmic), %.2f us/step of host submission\n",
         ms * 1e3 / steps, 16.0 * 1048576 / (ms * 1e-3 / steps) / 1e9, t_submit / steps);

  /* a small result end to end: sum -> 1 float, stored into pinned host memory by the kernel itself */
  ct_tensor s;
  CHECK(ct_sum(r, &s));
  float* host = NULL;
  uint64_t got = 0;
  for (int i = 0; i < 200; ++i) {
    CHECK(ct_flat_buffer(s, &host, &got));
    CHECK(ct_flat_buffer_release(host));
  }
  t0 = now_us();
  const int small_steps = steps / 10 > 100 ? steps / 10 : 100;
  for (int i = 0; i < small_steps; ++i) {
    CHECK(ct_flat_buffer(s, &host, &got));
    CHECK(ct_flat_buffer_release(host));
  }
  printf("sum(tanh(a*b+c)).flatBuffer from C: %.2f us/call end to end (fold fused with the chain, result stored to host by the kernel)\n",
         (now_us() - t0) / small_steps);

  cc_stats_t st;
  CHECK(cc_stats(&st));
  printf("compiles=%llu launches=%llu device_kernels=%llu pool_hits=%llu/%llu\n", (unsigned long long)st.compiles, (unsigned long long)st.launches,
         (unsigned long long)st.device_kernels, (unsigned long long)st.pool_hits, (unsigned long long)st.alloc_calls);
  ct_tensor all[] = {s, r, abc, ab, ccached, cb, ca, c, b, a};
  for (size_t i = 0; i < sizeof all / sizeof all[0]; ++i) CHECK(ct_release(all[i]));
  free(ha), free(hb), free(hc), free(hr);
  CHECK(cc_shutdown());
  return 0;
}
