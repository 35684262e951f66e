// approx/poisson.h -- same declarations as the reference for the mask overloads (lib/approx/include/approx/poisson.h:
// 12-52); implemented in cpp/src/approx_satfill.cpp on top of libsatfill.so.
#pragma once

#include <optional>

#include "utils.h"

namespace approx {

struct PerfInfo {  // poisson.h:12-21
    long region_size = 0;
    f64 tolerance = 0.0;
    long max_iterations = 0;
    long iterations = 0;
    f64 error = 0.0;
    f64 solve_time = 0.0;
    void write(fs::path const& output) const;  // appends one CSV line (poisson.cpp:14-19)
};

// Offset / white-key overload (poisson.h:30-33, poisson.cpp:21-143): the replacement image is pasted at
// (start_row, start_column); its unknowns are the pixels that are NOT the white key (MultiChannelImage::valid_pixel).
// The reference solves with a direct factorisation; here the system goes through the same CG (tolerance 1e-12).
void blend_images_poisson(MultiChannelImage& input_images, MultiChannelImage const& replacement_images, int start_row,
    int start_column);
void blend_images_poisson(MultiChannelImage& input_images, MultiChannelImage const& replacement_images,
    MatX<bool> const& invalid_mask, f64 tolerance = 1e-6, std::optional<int> max_iterations = {});
std::vector<MatX<f64>> blend_images_poisson(std::vector<MatX<f64>> const& input_images,
    std::vector<MatX<f64>> const& replacement_images, MatX<bool> const& invalid_mask, f64 tolerance = 1e-6,
    std::optional<int> max_iterations = {});

// The record of the last blend (the reference appends it to a hard-coded path, poisson.cpp:287-289).
PerfInfo const& last_perf_info();

}  // namespace approx
