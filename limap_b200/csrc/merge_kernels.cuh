// merge_kernels.cuh — post-triangulation track filters and the remerge pair test.
#pragma once
#include "tri_kernels.cuh"

namespace lm {

// Per-support checks of merging/merging_utils.cc:27-155; one thread per supporting 2D line.
struct SupportParams {
  const ViewD *views;        // [n_views]
  const int64_t *sup_off;    // [T+1]
  const int32_t *sup_view;   // [S]
  const double4 *segs;       // [S] x1,y1,x2,y2
  const double *track_line;  // [T][6]
  int64_t T, S;
  double th_angular2d, th_perp2d, th_sv_angular3d, th_overlap;
  uint8_t *flags;            // [S] bit0 reprojection ok, bit1 sensitivity ok, bit2 overlap ok
};
void launch_support_flags(const SupportParams &p, cudaStream_t s);

// All-pairs LineLinker3d::check_connection of RemergeLineTracks (merging/merging.cc:527-556).
struct RemergeParams {
  const double *lines;       // [T][7] start, end, uncertainty
  const float4 *dirf;        // [T] unit direction in fp32 (gate), w unused
  const float4 *ballf;       // [T] midpoint - origin, grown radius (gate)
  const uint8_t *active;     // [T]
  int64_t T;
  int all_active;
  LinkerDev<double> lk;      // after set_to_spatial_merging()
  float cos_gate;            // cos(th_angle) - margin; gate used only when use_gate
  int use_gate;
  int use_ball;              // use_innerseg: the ball gate is a necessary condition
  uint32_t *edges;           // [capacity][2] (a < b), unordered
  unsigned long long *counter; // [2]: edges found, pairs past the gate
  unsigned long long capacity;
};
void launch_remerge_dirs(const double *lines, int64_t T, const double origin[3], double th_innerseg, float4 *dirf,
                         float4 *ballf, cudaStream_t s);
void launch_remerge_pairs(const RemergeParams &p, cudaStream_t s);

} // namespace lm
