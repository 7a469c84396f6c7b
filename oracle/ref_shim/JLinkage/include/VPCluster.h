// oracle/ref_shim: stands in for <JLinkage/include/VPCluster.h> (see VPSample.h).
#pragma once
#include <vector>
namespace VPCluster {
int run(std::vector<unsigned int> &labels, std::vector<unsigned int> &label_count, std::vector<std::vector<float> *> *pts,
        std::vector<std::vector<float> *> *models, float inlier_threshold, int minimal_set);
}
