// oracle/ref_shim: stands in for <colmap/sensor/bitmap.h> (TEST INFRASTRUCTURE): image files are never read here.
#pragma once
#include <string>
namespace colmap {
class Bitmap {
public:
  bool Read(const std::string &, bool = true) { return false; }
  int Width() const { return 0; }
  int Height() const { return 0; }
  bool ExifFocalLength(double *) const { return false; }
};
} // namespace colmap
