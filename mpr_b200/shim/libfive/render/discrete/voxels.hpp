// Minimal libfive::Voxels (include/libfive/render/discrete/voxels.hpp:21-47): bounds + sample
// positions at voxel centres.  Only used by drivers to describe the CPU comparison render.
#pragma once
#include <array>
#include <cmath>
#include <vector>
#include <Eigen/Eigen>

namespace libfive {
class Voxels {
public:
    Voxels(const Eigen::Vector3f& lo, const Eigen::Vector3f& hi, float res) : lower(lo), upper(hi) {
        for (int i = 0; i < 3; ++i) {
            const float extent = upper(i) - lower(i);
            const int n = extent > 0 ? int(std::ceil(extent * res)) : 1;
            const float mid = (upper(i) + lower(i)) / 2, half = extent > 0 ? n / res / 2 : 0;
            lower(i) = mid - half;
            upper(i) = mid + half;
            pts[i].resize(n);
            for (int k = 0; k < n; ++k) pts[i][k] = lower(i) + (upper(i) - lower(i)) * (k + 0.5f) / n;
        }
    }
    Eigen::Vector3f lower, upper;
    std::array<std::vector<float>, 3> pts;
};
}  // namespace libfive
