// Unit test of cc::SmallVec (csrc/common.h), built with -fsanitize=address,undefined by tests/test_native_units.py.
#include <cassert>
#include <cstdio>
#include <utility>
#include <vector>

#include "common.h"

struct P {
  int a;
  unsigned long long b;
};

template <class V>
static void check(const V& v, int n, int first = 0) {
  assert((int)v.size() == n);
  int i = first;
  for (const P& p : v) {
    assert(p.a == i && p.b == (unsigned long long)i * 3);
    ++i;
  }
}

int main() {
  using V = cc::SmallVec<P, 4>;
  V v;
  assert(v.empty());
  for (int i = 0; i < 3; ++i) v.push_back(P{i, (unsigned long long)i * 3});
  check(v, 3);                       // inline
  V c(v);
  check(c, 3);
  V m(std::move(c));
  check(m, 3);
  assert(c.empty());
  c.push_back(P{0, 0});              // a moved-from vector is usable
  check(c, 1);
  for (int i = 3; i < 40; ++i) v.push_back(P{i, (unsigned long long)i * 3});
  check(v, 40);                      // on the heap (grown three times)
  V h(v);
  check(h, 40);
  V hm(std::move(h));
  check(hm, 40);
  assert(h.empty());
  h = v;                             // copy-assign heap into a moved-from vector
  check(h, 40);
  h = V{P{0, 0}, P{1, 3}};           // move-assign an inline temporary over a heap vector
  check(h, 2);
  m = std::move(hm);                 // move-assign heap over inline
  check(m, 40);
  assert(hm.empty());
  m = m;                             // self-assignment
  check(m, 40);
  m.clear();
  assert(m.empty());
  for (int i = 0; i < 5; ++i) m.push_back(P{i, (unsigned long long)i * 3});
  check(m, 5);                       // capacity kept after clear
  std::vector<P> sv{{0, 0}, {1, 3}, {2, 6}, {3, 9}, {4, 12}, {5, 15}};
  V fromv(sv);
  check(fromv, 6);
  assert(fromv.back().a == 5 && fromv[2].b == 6 && fromv.data() == fromv.begin());
  std::vector<V> pool;               // vectors of SmallVecs relocate them (Block::pending lives in one)
  for (int k = 0; k < 50; ++k) {
    V e;
    for (int i = 0; i < k % 9; ++i) e.push_back(P{i, (unsigned long long)i * 3});
    pool.push_back(std::move(e));
  }
  for (int k = 0; k < 50; ++k) check(pool[(size_t)k], k % 9);
  puts("smallvec ok");
  return 0;
}
