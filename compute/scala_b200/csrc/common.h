// common.h — status/error plumbing shared by every translation unit of libcompute_cuda.so.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

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
