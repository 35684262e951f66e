// approx/utils.h -- API value types of the fill path (reference: lib/approx/include/approx/utils.h:15-106), without
// the Eigen sparse typedefs (there is no matrix any more), range-v3 or OpenCV.
#pragma once

#include <filesystem>
#include <optional>
#include <vector>

#include "utils/types.h"

namespace fs = std::filesystem;
using namespace utils;

namespace approx {

struct index_t {  // utils.h:19-27
    Eigen::Index row;
    Eigen::Index col;
    bool operator==(index_t other) const { return row == other.row && col == other.col; }
};

template <typename T>
bool within_bounds(MatX<T> const& image, index_t index)  // utils.h:29-33
{
    return index.row >= 0 && index.row < image.rows() && index.col >= 0 && index.col < image.cols();
}

// utils.h:35-50: in-image 4-neighbours in the order (-1,0) (+1,0) (0,-1) (0,+1); tests/approximation.h:9-33.
template <typename T>
std::vector<index_t> valid_neighbours(MatX<T> const& image, index_t index)
{
    std::vector<index_t> out;
    const index_t cand[4] = { { index.row - 1, index.col }, { index.row + 1, index.col }, { index.row, index.col - 1 },
        { index.row, index.col + 1 } };
    for (index_t c : cand)
        if (within_bounds(image, c))
            out.push_back(c);
    return out;
}

struct MultiChannelImage {  // utils.h:52-106
    explicit MultiChannelImage(std::vector<MatX<f64>> images) : images(std::move(images)) {}
    MultiChannelImage(std::initializer_list<MatX<f64>> images) : images(images) {}
    MultiChannelImage(size_t channels, Eigen::Index rows, Eigen::Index cols)
    {
        for (size_t c = 0; c < channels; ++c)
            images.emplace_back(MatX<f64>::Zero(rows, cols));
    }
    MultiChannelImage() = default;

    std::vector<MatX<f64>> images;

    f64 const& operator()(size_t c, Eigen::Index row, Eigen::Index col) const { return images.at(c)(row, col); }
    f64& operator()(size_t c, Eigen::Index row, Eigen::Index col) { return images.at(c)(row, col); }
    MatX<f64> const& operator[](size_t c) const { return images[c]; }
    MatX<f64>& operator[](size_t c) { return images[c]; }
    [[nodiscard]] Eigen::Index size() const { return images[0].size(); }
    [[nodiscard]] Eigen::Index rows() const { return images[0].rows(); }
    [[nodiscard]] Eigen::Index cols() const { return images[0].cols(); }
    [[nodiscard]] bool valid_pixel(Eigen::Index row, Eigen::Index col) const  // white key, utils.h:101-105
    {
        bool invalid = static_cast<int>(images[0](row, col)) == 1 && static_cast<int>(images[1](row, col)) == 1
            && static_cast<int>(images[2](row, col)) == 1;
        return !invalid;
    }
};

}  // namespace approx
