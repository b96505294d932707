// ndat.h — N-dimensional affine transforms on row-major double matrices, `rows = dims of the space mapped TO` x
// `cols = dims of the space mapped FROM + 1` (constant column last). Host-side only, always in double — same
// arithmetic and accumulation order as NDimensionalAffineTransform.scala:13-127 so composed views are bit-identical.
#pragma once
#include <vector>

#include "common.h"

namespace cc {
namespace ndat {

using Matrix = std::vector<double>;

// N:13-26 — identity plus `offset` in the constant column
inline Matrix translate(const std::vector<double>& offset) {
  const size_t n = offset.size();
  Matrix m(n * (n + 1), 0.0);
  for (size_t i = 0; i < n; ++i) {
    m[i * (n + 1) + i] = 1.0;
    m[i * (n + 1) + n] = offset[i];
  }
  return m;
}

// N:48-95 — m02 = m12 o m01 with m01: l1 x (l0+1), m12: l2 x (l1+1); accumulated over index1 in order
inline Matrix concatenate_into(const Matrix& m01, const Matrix& m12, size_t l0, size_t l1, size_t l2) {
  CC_REQUIRE(m01.size() == l1 * (l0 + 1) && m12.size() == l2 * (l1 + 1), CC_ERR_ILLEGAL_ARGUMENT, "affine matrix sizes do not compose");
  Matrix m02((l0 + 1) * l2, 0.0);
  for (size_t i2 = 0; i2 < l2; ++i2) {
    for (size_t i0 = 0; i0 < l0; ++i0) {
      double acc = 0.0;
      for (size_t i1 = 0; i1 < l1; ++i1) acc = acc + m12[i2 * (l1 + 1) + i1] * m01[i1 * (l0 + 1) + i0];
      m02[i2 * (l0 + 1) + i0] = acc;
    }
    double acc = m12[i2 * (l1 + 1) + l1];
    for (size_t i1 = 0; i1 < l1; ++i1) acc = acc + m12[i2 * (l1 + 1) + i1] * m01[i1 * (l0 + 1) + l0];
    m02[i2 * (l0 + 1) + l0] = acc;
  }
  return m02;
}

// N:28-35 — apply m01 first (from a space of `length0` dims), then m12
inline Matrix pre_concatenate(const Matrix& m01, const Matrix& m12, size_t length0) {
  CC_REQUIRE(m01.size() % (length0 + 1) == 0, CC_ERR_ILLEGAL_ARGUMENT, "bad affine matrix");
  const size_t l1 = m01.size() / (length0 + 1);
  CC_REQUIRE(m12.size() % (l1 + 1) == 0, CC_ERR_ILLEGAL_ARGUMENT, "bad affine matrix");
  const size_t l2 = m12.size() / (l1 + 1);
  return concatenate_into(m01, m12, length0, l1, l2);
}

// N:37-46
inline Matrix concatenate(const Matrix& m12, const Matrix& m01, size_t length2) {
  CC_REQUIRE(length2 > 0 && m12.size() % length2 == 0, CC_ERR_ILLEGAL_ARGUMENT, "bad affine matrix");
  const size_t l1 = m12.size() / length2 - 1;
  CC_REQUIRE(l1 > 0 && m01.size() % l1 == 0, CC_ERR_ILLEGAL_ARGUMENT, "bad affine matrix");
  const size_t l0 = m01.size() / l1 - 1;
  return concatenate_into(m01, m12, l0, l1, length2);
}

// N:97-127
inline std::vector<double> transform(const Matrix& m, const std::vector<double>& source) {
  const size_t n = source.size();
  CC_REQUIRE(m.size() % (n + 1) == 0, CC_ERR_ILLEGAL_ARGUMENT, "bad affine matrix");
  const size_t rows = m.size() / (n + 1);
  std::vector<double> out(rows);
  for (size_t y = 0; y < rows; ++y) {
    double acc = m[y * (n + 1) + n];
    for (size_t x = 0; x < n; ++x) acc = acc + m[y * (n + 1) + x] * source[x];
    out[y] = acc;
  }
  return out;
}

}  // namespace ndat
}  // namespace cc
