// oracle/ref_shim: stands in for <JLinkage/include/VPSample.h> of B1ueber2y/JLinkage (an un-vendored submodule of the
// reference: TEST INFRASTRUCTURE). The library's hypothesis sampling and clustering are restated in oracle/orc_vp.h
// (J-Linkage with a counter-based RNG: DESIGN.md 3.4); VPSample::run / VPCluster::run forward to that restatement, so that
// limap's own wrapper (vplib/JLinkage/JLinkage.cc: length filter, cluster filtering, label renumbering, VP fit) compiles
// unchanged and is pinned by tests/test_ref_pinning.py. The seed and the image index of the draw come from
// ref_vp_set_context (oracle/ref_api.cpp); definitions live there (compiled without FMA contraction, like orc_vp.cpp).
#pragma once
#include <vector>
namespace VPSample {
std::vector<std::vector<float> *> *run(std::vector<std::vector<float> *> *pts, int num_models, int minimal_set, int sampling_type,
                                       int seed_mode);
}
