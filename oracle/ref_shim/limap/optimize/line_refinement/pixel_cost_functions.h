// oracle/ref_shim: shadows limap/optimize/line_refinement/pixel_cost_functions.h (heatmap / feature residuals: HDF5,
// boost, ceres cubic interpolation -- outside the hot path) so that cost_functions.h, which includes it,
// compiles unchanged for its geometric and VP functors. TEST INFRASTRUCTURE.
#pragma once
