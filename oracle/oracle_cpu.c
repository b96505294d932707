/* TEST INFRASTRUCTURE ONLY — C restatement ("port") of the kernels the reference's `cpu` backend runs for the
 * BASELINE.json configurations, used (a) to cross-check the numpy oracle at sizes numpy is too slow for and (b) as
 * the CPU baseline bench.py times on the GPU box's host cores.  Never linked into libcompute_cuda.so.
 *
 * The reference generates one scalar OpenCL C work-item per output element and lets the CPU OpenCL driver spread
 * work-groups over all cores (README.md:26-34); the loops below are those work-items with an OpenMP `parallel for`
 * standing in for the driver's scheduling.  `sum` stays a single work-item, exactly as the reference launches it on
 * CPU devices (Tensors.scala:690-695).  Paths relative to /root/reference:
 *   T: Tensors/src/main/scala/com/thoughtworks/compute/Tensors.scala
 *   K: OpenCLKernelBuilder/src/main/scala/com/thoughtworks/compute/OpenCLKernelBuilder.scala
 * Parity: pinned through tests/test_oracle_c.py against oracle/reference.py, which is pinned against the reference's
 * golden vectors (tests/test_oracle_goldens.py).  exp/log/tanh come from the host libm: parity unpinned for those
 * (see oracle/__init__.py).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU baseline must still use all host cores (as POCL spreads work-groups over
 * every core, README.md:26-34), so bench.py sets the count explicitly */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n >= 1) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* T:106-117 */
static inline uint32_t wang_hash(uint32_t value) {
  value = (value ^ 61u) ^ (value >> 16);
  value *= 9u;
  value ^= value << 4;
  value *= 0x27d4eb2du;
  value ^= value >> 15;
  return value;
}

/* T:432-443 — buffer[i] = hash(i ^ seed) / 4294967296.0f */
void oracle_random(float* out, int64_t n, uint32_t seed) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) out[i] = (float)wang_hash((uint32_t)i ^ seed) / 4294967296.0f;
}

/* C1: _9 = a*b; _11 = _9 + c; _12 = tanh(_11) — K:532-553, 319-325. Compiled with -ffp-contract=off
 * or =fast by the build script to bracket FP_CONTRACT ON. */
void oracle_c1(const float* a, const float* b, const float* c, float* out, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const float t9 = a[i] * b[i];
    const float t11 = t9 + c[i];
    out[i] = tanhf(t11);
  }
}

/* C2: t=a*b+c; u=exp(t); v=log(u+a); w=tanh(v*b); out=w+c — one SSA line per node as K:271-325,516-571 emit them */
void oracle_c2(const float* a, const float* b, const float* c, float* out, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const float ab = a[i] * b[i];
    const float t = ab + c[i];
    const float u = expf(t);
    const float ua = u + a[i];
    const float v = logf(ua);
    const float vb = v * b[i];
    const float w = tanhf(vb);
    out[i] = w + c[i];
  }
}

/* T:313-351 with get_global_size(0) == 1 (T:690-695): float16 lanes accumulated sequentially, hi/lo tree, tail */
float oracle_sum_cpu_order(const float* buffer, int64_t length) {
  const int64_t vl = length / 16;
  float s;
  if (vl >= 1) {
    float acc[16];
    for (int l = 0; l < 16; ++l) acc[l] = buffer[l];
    for (int64_t v = 1; v < vl; ++v)
      for (int l = 0; l < 16; ++l) acc[l] += buffer[16 * v + l];
    float f8[8], f4[4], f2[2];
    for (int l = 0; l < 8; ++l) f8[l] = acc[8 + l] + acc[l];
    for (int l = 0; l < 4; ++l) f4[l] = f8[4 + l] + f8[l];
    for (int l = 0; l < 2; ++l) f2[l] = f4[2 + l] + f4[l];
    s = f2[0] + f2[1];
    for (int64_t i = 16 * vl; i < length; ++i) s += buffer[i];
  } else {
    s = 0.0f;
    for (int64_t i = 0; i < length; ++i) s += buffer[i];
  }
  return s;
}

double oracle_sum_fp64(const float* buffer, int64_t length) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < length; ++i) s += (double)buffer[i];
  return s;
}

/* per-axis sum as users write it, `t.split(axis).reduce(_ + _)` (README.md:301-310): one work-item per output element,
 * fp32 left fold over the split index. x is [rows, cols] row-major. */
void oracle_axis_sum_2d(const float* x, int64_t rows, int64_t cols, int axis, float* out) {
  if (axis == 0) {
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < cols; ++j) {
      float acc = x[j];
      for (int64_t t = 1; t < rows; ++t) acc = acc + x[t * cols + j];
      out[j] = acc;
    }
  } else {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < rows; ++i) {
      float acc = x[i * cols];
      for (int64_t t = 1; t < cols; ++t) acc = acc + x[i * cols + t];
      out[i] = acc;
    }
  }
}

/* matmul as split/broadcast/sum (benchmarks.scala:188-191): C[i,k] = left fold over t of A[i,t]*B[t,k] in fp32 */
void oracle_matmul_left_fold(const float* a, const float* b, float* c, int64_t m, int64_t kk, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < m; ++i) {
    float* row = c + i * n;
    for (int64_t k = 0; k < n; ++k) row[k] = a[i * kk] * b[k];
    for (int64_t t = 1; t < kk; ++t) {
      const float av = a[i * kk + t];
      const float* br = b + t * n;
      for (int64_t k = 0; k < n; ++k) {
        const float p = av * br[k];
        row[k] = row[k] + p;
      }
    }
  }
}

/* K:348-411 for a rank-3 output over a rank-r (r <= 3) source with an INTEGER matrix m[r][4] (row-major, constant last):
 * index_y = g0*m[y][0] + g1*m[y][1] + g2*m[y][2] + m[y][3]; two-sided bounds -> padding. */
void oracle_affine_gather_3d(const float* src, const int64_t* src_shape, int rank, const int64_t* m, const int64_t* out_shape,
                             float padding, float* out) {
  const int64_t d0 = out_shape[0], d1 = out_shape[1], d2 = out_shape[2];
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t g0 = 0; g0 < d0; ++g0)
    for (int64_t g1 = 0; g1 < d1; ++g1)
      for (int64_t g2 = 0; g2 < d2; ++g2) {
        int ok = 1;
        int64_t off = 0;
        for (int y = 0; y < rank; ++y) {
          const int64_t idx = g0 * m[y * 4 + 0] + g1 * m[y * 4 + 1] + g2 * m[y * 4 + 2] + m[y * 4 + 3];
          const int any = m[y * 4 + 0] || m[y * 4 + 1] || m[y * 4 + 2] || m[y * 4 + 3];
          if (any && (idx < 0 || idx >= src_shape[y])) ok = 0;
          off = off * src_shape[y] + (ok ? idx : 0);
        }
        out[(g0 * d1 + g1) * d2 + g2] = ok ? src[off] : padding;
      }
}
