// approx/laplace.h -- same declarations as the reference (lib/approx/include/approx/laplace.h:11-28); implemented in
// cpp/src/approx_satfill.cpp on top of the C-ABI of libsatfill.so (include/satfill.h).  apply_laplace (cv::Mat) is
// only declared when OpenCV's C++ headers exist.
#pragma once

#include <unordered_map>

#include "utils.h"

namespace approx {

struct ConnectedComponents {
    MatX<int> matrix;
    std::unordered_map<int, std::vector<index_t>> region_map;
};

// 4-connectivity, background 0, labels 1..K by first pixel in row-major raster order; region_map[l] in raster order.
ConnectedComponents find_connected_components(MatX<bool> const& invalid);

// Laplace fill with Dirichlet boundary (laplace.cpp:122-132).  Throws std::runtime_error when the element counts
// differ (laplace.cpp:124-127); an empty mask is a no-op (laplace.cpp:41-44).
void fill_missing_portion_smooth_boundary(MatX<f64>& input_image, MatX<bool> const& invalid_pixels);

// Knobs the reference does not have (its Laplace runs Eigen defaults, laplace.cpp:113-114): tolerance <= 0 keeps epsilon.
struct LaplaceOptions {
    f64 tolerance = 0.0;
    long max_iterations = 0;
    bool multigrid = true;  // false: Eigen's DiagonalPreconditioner (the reference's own), ~50x more iterations
};
void set_laplace_options(LaplaceOptions const& options);

}  // namespace approx
