/* TEST INFRASTRUCTURE ONLY: drives BASELINE config 1 in miniature through the C ABI with the driver spy in front of libcuda and the
 * malloc counter preloaded, and prints how many heap allocations the library makes per steady-state step (a cached plan: evaluate +
 * release) and per freshly built expression (build three nodes, first evaluation, release everything). */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>

#include "compute_cuda.h"

#define CHECK(call)                                                         \
  do {                                                                      \
    int st_ = (call);                                                       \
    if (st_ != 0) {                                                         \
      fprintf(stderr, "%s -> %d: %s\n", #call, st_, cc_last_error());       \
      exit(1);                                                              \
    }                                                                       \
  } while (0)

int main(void) {
  long (*count)(void) = (long (*)(void))dlsym(RTLD_DEFAULT, "malloc_count_get");
  if (!count) return 2;
  CHECK(cc_init(-1));
  int32_t shape[2] = {64, 64};
  float* h = calloc(64 * 64, 4);
  ct_tensor a, b, c, ca, cb, cc_, ab, abc, r;
  CHECK(ct_from_host(h, shape, 2, 0.f, &a));
  CHECK(ct_from_host(h, shape, 2, 0.f, &b));
  CHECK(ct_from_host(h, shape, 2, 0.f, &c));
  CHECK(ct_do_cache(a, &ca));
  CHECK(ct_do_cache(b, &cb));
  CHECK(ct_do_cache(c, &cc_));
  CHECK(ct_binary(CT_TIMES, ca, cb, &ab));
  CHECK(ct_binary(CT_PLUS, ab, cc_, &abc));
  CHECK(ct_unary(CT_TANH, abc, &r));
  cc_buffer out;
  for (int i = 0; i < 64; ++i) {
    CHECK(ct_do_buffer(r, &out, NULL));
    CHECK(cc_buffer_release(out));
  }
  long m0 = count();
  for (int i = 0; i < 1000; ++i) {
    CHECK(ct_do_buffer(r, &out, NULL));
    CHECK(cc_buffer_release(out));
  }
  long steady = count() - m0;
  for (int i = 0; i < 64; ++i) {
    ct_tensor x, y, z;
    CHECK(ct_binary(CT_TIMES, ca, cb, &x)); CHECK(ct_binary(CT_PLUS, x, cc_, &y)); CHECK(ct_unary(CT_TANH, y, &z));
    CHECK(ct_do_buffer(z, &out, NULL)); CHECK(cc_buffer_release(out));
    CHECK(ct_release(z)); CHECK(ct_release(y)); CHECK(ct_release(x));
  }
  m0 = count();
  for (int i = 0; i < 1000; ++i) {
    ct_tensor x, y, z;
    CHECK(ct_binary(CT_TIMES, ca, cb, &x)); CHECK(ct_binary(CT_PLUS, x, cc_, &y)); CHECK(ct_unary(CT_TANH, y, &z));
    CHECK(ct_do_buffer(z, &out, NULL)); CHECK(cc_buffer_release(out));
    CHECK(ct_release(z)); CHECK(ct_release(y)); CHECK(ct_release(x));
  }
  long fresh = count() - m0;
  printf("{\"mallocs_per_steady_step\": %.2f, \"mallocs_per_fresh_expression\": %.2f}\n", steady / 1000.0, fresh / 1000.0);
  return 0;
}
