// approx/laplace.h -- same declarations as the reference (lib/approx/include/approx/laplace.h:11-31); implemented in
// cpp/src/approx_satfill.cpp on top of the C-ABI of libsatfill.so (include/satfill.h).  apply_laplace (cv::Mat) and the
// image-file helpers of approx/utils.h are declared -- and defined inline, on sa_apply_laplace_u8 -- only where OpenCV's
// C++ headers exist (__has_include(<opencv2/core.hpp>)); this image has none, so tests/fake_opencv/ holds a minimal cv::Mat
// that lets the CPU suite compile and run that code against the fake C-ABI.
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
LaplaceOptions const& laplace_options();

}  // namespace approx

namespace approx {
// approx::apply_laplace (laplace.h:31, laplace.cpp:134-168) on plain buffers: `image` and `invalid_image` are rows x cols x 3
// bytes in cv::imread order (B, G, R); mask = (R >= red_threshold) & (G <= 150) of invalid_image (laplace.cpp:141-146); every
// channel is filled with that mask -- in ONE batched solve on the device, where upstream re-assembles its matrix per channel
// -- and `out` receives rows x cols x 3 doubles (the CV_64FC3 layout of upstream's cv::merge).  The cv::Mat overload below
// is a thin wrapper; this one exists so that callers without OpenCV (and the tests) reach the same code.
void apply_laplace(const unsigned char* image, const unsigned char* invalid_image, Eigen::Index rows, Eigen::Index cols,
    f64 red_threshold, f64* out);
}  // namespace approx

#if __has_include(<opencv2/core.hpp>)
#include <opencv2/core.hpp>

#include <stdexcept>

namespace approx {

// cv::Mat apply_laplace(cv::Mat const& image, cv::Mat const& invalid_image, f64 red_threshold)   -- laplace.h:31.
// Throws like fill_missing_portion_smooth_boundary when the sizes differ (laplace.cpp:124-127).
inline cv::Mat apply_laplace(cv::Mat const& image, cv::Mat const& invalid_image, f64 red_threshold)
{
    if (image.rows != invalid_image.rows || image.cols != invalid_image.cols)
        throw std::runtime_error("Input image and mask need to be the same size");
    if (image.type() != CV_8UC3 || invalid_image.type() != CV_8UC3)
        throw std::runtime_error("apply_laplace: CV_8UC3 images (cv::imread(path, cv::IMREAD_COLOR))");
    const cv::Mat img = image.isContinuous() ? image : image.clone();
    const cv::Mat inv = invalid_image.isContinuous() ? invalid_image : invalid_image.clone();
    cv::Mat out(image.rows, image.cols, CV_64FC3);
    apply_laplace(img.data, inv.data, image.rows, image.cols, red_threshold, reinterpret_cast<f64*>(out.data));
    return out;
}

}  // namespace approx
#endif
