// lm_kernels.cuh — device layout of the batched line refinement (see lm_kernels.cu).
#pragma once
#include "lm_math.cuh"

namespace lm {

// One supporting 2D segment of a track, digested once per solve (168 B, resident in L1/L2 afterwards).
struct LMBlockDev {
  double p[4]; // x1,y1,x2,y2
  double k[4]; // fx,fy,cx,cy
  double R[9]; // ceres::QuaternionToRotation(qvec)
  double t[3];
  double w;    // ScaledLoss weight = |segment| / 30
};

struct LMParams {
  const LMBlockDev *blocks; // [n]
  const int64_t *sup_off;   // [T+1]
  const double *x0;         // [T][6] uvec, wvec
  const uint8_t *active;    // [T] 0 = parameter blocks held constant (count_images < min_num_images)
  double *x_out;            // [T][6]
  int32_t *iters;           // [T][2] iterations, successful steps
  double *cost;             // [T][2] initial, final
  int32_t *term;            // [T] termination code
  int64_t T;
  double geometric_alpha, cauchy_scale;
  int max_num_iterations, max_invalid;
};

void launch_lm_prepare(const double *segs, const int32_t *sup_view, const double *kvec, const double *qvec,
                       const double *tvec, int64_t n, LMBlockDev *out, cudaStream_t s);
void launch_lm_refine(const LMParams &p, cudaStream_t s);

} // namespace lm
