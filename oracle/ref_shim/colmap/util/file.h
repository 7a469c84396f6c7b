// oracle/ref_shim: stands in for <colmap/util/file.h> (TEST INFRASTRUCTURE).
#pragma once
#include <string>
namespace colmap {
inline std::string JoinPaths(const std::string &a, const std::string &b) { return a + "/" + b; }
} // namespace colmap
