// oracle/ref_shim: stands in for <ceres/ceres.h> where a header of the reference includes it without using it
// (base/infinite_line.h:11). TEST INFRASTRUCTURE.
#pragma once
namespace ceres {}
