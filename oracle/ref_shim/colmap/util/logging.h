// oracle/ref_shim: stands in for <colmap/util/logging.h> (TEST INFRASTRUCTURE): the THROW_CHECK / CHECK / LOG macros
// of COLMAP + glog, reduced to "throw std::invalid_argument on failure" and a stream that is discarded.
#pragma once
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
namespace colmap_shim {
struct NullStream { template <typename T> NullStream &operator<<(const T &) { return *this; } NullStream &operator<<(std::ostream &(*)(std::ostream &)) { return *this; } };
struct FatalStream {
  std::ostringstream os; const char *what;
  explicit FatalStream(const char *w) : what(w) {}
  template <typename T> FatalStream &operator<<(const T &v) { os << v; return *this; }
  [[noreturn]] ~FatalStream() noexcept(false) { throw std::invalid_argument(std::string("check failed: ") + what + " " + os.str()); }
};
} // namespace colmap_shim
#define SHIM_CHECK_(cond, text) if (cond) {} else ::colmap_shim::FatalStream(text)
#define THROW_CHECK(c) SHIM_CHECK_((c), #c)
#define THROW_CHECK_MSG(c, msg) SHIM_CHECK_((c), #c) << msg
#define THROW_CHECK_EQ(a, b) SHIM_CHECK_(((a) == (b)), #a " == " #b)
#define THROW_CHECK_NE(a, b) SHIM_CHECK_(((a) != (b)), #a " != " #b)
#define THROW_CHECK_LT(a, b) SHIM_CHECK_(((a) < (b)), #a " < " #b)
#define THROW_CHECK_LE(a, b) SHIM_CHECK_(((a) <= (b)), #a " <= " #b)
#define THROW_CHECK_GT(a, b) SHIM_CHECK_(((a) > (b)), #a " > " #b)
#define THROW_CHECK_GE(a, b) SHIM_CHECK_(((a) >= (b)), #a " >= " #b)
#define THROW_CHECK_NOTNULL(p) (p)
#define CHECK(c) SHIM_CHECK_((c), #c)
#define CHECK_EQ(a, b) THROW_CHECK_EQ(a, b)
#define CHECK_NE(a, b) THROW_CHECK_NE(a, b)
#define CHECK_LT(a, b) THROW_CHECK_LT(a, b)
#define CHECK_LE(a, b) THROW_CHECK_LE(a, b)
#define CHECK_GT(a, b) THROW_CHECK_GT(a, b)
#define CHECK_GE(a, b) THROW_CHECK_GE(a, b)
#define CHECK_NOTNULL(p) (p)
#define LOG(sev) ::colmap_shim::NullStream()
#define VLOG(n) ::colmap_shim::NullStream()
#define LOG_FATAL_THROW(ex) ::colmap_shim::FatalStream(#ex)
