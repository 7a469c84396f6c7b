// sfm_kernels.cuh — visual-neighbour ranking from a sparse point model:
// SfmModel::{GetMaxIoUImages, GetMaxDiceCoeffImages, GetMaxOverlapImages} (pointsfm/sfm_model.cc:88-226) on top of
// COLMAP's ComputeSharedPoints / ComputeTriangulationAngles (colmap/mvs/model.cc, restated; see oracle/orc_sfm.cpp).
#pragma once
#include "lm_math.cuh"

namespace lm {

// (point, pair-of-track-entries) -> sort key: (min image << 16 | max image) << 32 | float bits of the triangulation angle
void launch_sfm_pair_keys(const double *centres, const double *xyz, const int64_t *track_off, const int32_t *track_img,
                          const int64_t *rec_off, int64_t n_points, int64_t n_rec, unsigned long long *keys,
                          unsigned int *num_points, cudaStream_t s);
// pair of every sorted record (high word of the key), for the run-length encoding
void launch_sfm_pair_ids(const unsigned long long *keys, int64_t n_rec, unsigned int *pair_ids, cudaStream_t s);
// per run (image pair): the 75th-percentile angle test and the similarity score; two directed records per surviving run
void launch_sfm_scores(const unsigned long long *keys, const unsigned int *run_pair, const unsigned int *run_len,
                       const unsigned int *run_start, int n_runs, const unsigned int *num_points, float min_angle_rad,
                       int mode, unsigned int *dir_val, unsigned long long *dir_key, unsigned int *n_dir, cudaStream_t s);
// first `num_images` destinations of every source image from the directed records sorted by (source, score desc, dest)
void launch_sfm_take(const unsigned int *dir_val, int64_t n_dir, int n_images, int num_images, int32_t *out_neighbors,
                     int32_t *out_count, cudaStream_t s);

} // namespace lm
