// lm_kernels.cuh — device layout of the batched line refinement (see lm_kernels.cu).
#pragma once
#include "lm_math.cuh"

namespace lm {

// One supporting 2D segment of a track, digested once per solve (200 B, resident in L1/L2 afterwards).
struct LMBlockDev {
  double p[4]; // x1,y1,x2,y2
  double k[4]; // fx,fy,cx,cy
  double R[9]; // ceres::QuaternionToRotation(qvec)
  double t[3];
  double w;    // ScaledLoss weight = |segment| / 30
  double vdir[3]; // GetDirectionFromVP(vp, kvec) of the support's VP (ceresbase/line_projection.h:125-135)
  double wvp;     // weight * vp_multiplier, 0 = no VP residual for this support
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
  const double *line3d;     // [n][6] track.line3d_list (start, end) or NULL
  double *seg_out;          // [T][6] output segment cut from the refined line; NaN when left to the host
  unsigned long long *next_track; // device counter, zero at launch: warps fetch tracks dynamically
  int num_outliers;
  int64_t T;
  double geometric_alpha, cauchy_scale;
  int max_num_iterations, max_invalid;
};

void launch_lm_prepare(const double *segs, const int32_t *sup_view, const double *kvec, const double *qvec,
                       const double *tvec, const double *sup_vp, double vp_multiplier, int64_t n, LMBlockDev *out,
                       cudaStream_t s);
void launch_lm_refine(const LMParams &p, cudaStream_t s);
void launch_lm_prologue(const double *line_init, const int64_t *sup_off, const int32_t *sup_view, int64_t T,
                        int min_num_images, double *x0, uint8_t *active, int *err, cudaStream_t s);

} // namespace lm
