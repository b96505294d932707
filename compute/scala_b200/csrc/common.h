// common.h — status/error plumbing shared by every translation unit of libcompute_cuda.so.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <new>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../../include/compute_cuda.h"

namespace cc {

// Carries a cc_status through C++ code; converted to the integer + thread-local message at the C boundary
// (the counterpart of checkErrorCode -> typed JVM exception, OpenCL.scala:251-312).
struct Error : std::runtime_error {
  int status;
  Error(int s, const std::string& m) : std::runtime_error(m), status(s) {}
};

std::string strprintf(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
void set_last_error(const std::string& m);
const char* last_error_cstr();

[[noreturn]] inline void fail(int status, const std::string& m) { throw Error(status, m); }

#define CC_REQUIRE(cond, status, ...)                  \
  do {                                                 \
    if (!(cond)) ::cc::fail((status), ::cc::strprintf(__VA_ARGS__)); \
  } while (0)

// A vector of trivially copyable things with room for N of them inside the object. The launch path builds half a dozen short lists per
// command (argument buffers, hazard marks, kernel parameters); as std::vectors those were 12 of the 14 heap allocations of a launch
// (0.5 of its 0.65 us of host time).
template <class T, size_t N>
class SmallVec {
  static_assert(std::is_trivially_copyable<T>::value, "SmallVec is for trivially copyable element types");

 public:
  SmallVec() {}
  SmallVec(std::initializer_list<T> l) {
    for (const T& x : l) push_back(x);
  }
  SmallVec(const std::vector<T>& v) {
    for (const T& x : v) push_back(x);
  }
  SmallVec(const SmallVec& o) { append(o.p_, o.n_); }
  SmallVec(SmallVec&& o) noexcept { steal(o); }
  SmallVec& operator=(const SmallVec& o) {
    if (this != &o) {
      n_ = 0;
      append(o.p_, o.n_);
    }
    return *this;
  }
  SmallVec& operator=(SmallVec&& o) noexcept {
    if (this != &o) {
      if (p_ != inline_) free(p_);
      p_ = inline_, cap_ = N, n_ = 0;
      steal(o);
    }
    return *this;
  }
  ~SmallVec() {
    if (p_ != inline_) free(p_);
  }
  void push_back(const T& x) {
    if (n_ == cap_) grow(cap_ * 2);
    p_[n_++] = x;
  }
  void clear() { n_ = 0; }
  size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  T* begin() { return p_; }
  T* end() { return p_ + n_; }
  const T* begin() const { return p_; }
  const T* end() const { return p_ + n_; }
  T* data() { return p_; }
  const T* data() const { return p_; }
  T& operator[](size_t i) { return p_[i]; }
  const T& operator[](size_t i) const { return p_[i]; }
  T& back() { return p_[n_ - 1]; }
  void pop_back() { --n_; }

 private:
  void grow(size_t cap) {
    T* q = (T*)malloc(cap * sizeof(T));
    if (!q) throw std::bad_alloc();
    memcpy((void*)q, (const void*)p_, n_ * sizeof(T));
    if (p_ != inline_) free(p_);
    p_ = q, cap_ = cap;
  }
  void append(const T* src, size_t n) {
    if (n_ + n > cap_) grow(n_ + n);
    memcpy((void*)(p_ + n_), (const void*)src, n * sizeof(T));
    n_ += n;
  }
  void steal(SmallVec& o) {  // *this is empty and inline
    if (o.p_ != o.inline_) {
      p_ = o.p_, cap_ = o.cap_, n_ = o.n_;
      o.p_ = o.inline_, o.cap_ = N;
    } else {
      memcpy((void*)inline_, (const void*)o.inline_, o.n_ * sizeof(T));
      n_ = o.n_;
    }
    o.n_ = 0;
  }
  T inline_[N];
  T* p_ = inline_;
  size_t n_ = 0, cap_ = N;
};

// Wraps the body of an extern "C" entry point.
template <class F>
inline int guarded(F&& f) noexcept {
  try {
    f();
    return CC_OK;
  } catch (const Error& e) {
    set_last_error(e.what());
    return e.status;
  } catch (const std::bad_alloc&) {
    set_last_error("host allocation failed");
    return CC_ERR_OUT_OF_MEMORY;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return CC_ERR_ILLEGAL_ARGUMENT;
  }
}

}  // namespace cc
