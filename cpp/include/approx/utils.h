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

// read_image / image_list_to_cv / write_image (utils.h:108-110, utils.cpp:16-68): the 8-bit image files either side of the
// fill.  Only with OpenCV's C++ headers (this image has none; tests/fake_opencv/ holds the few members used here so that the
// block is compiled and run by tests/test_host_api.py).  Channels are R, G, B in [0, 1], gamma 2.2 decoded on the way in and
// encoded (truncating, like the reference's static_cast<uchar>) on the way out; files hold B, G, R.
#if __has_include(<opencv2/imgcodecs.hpp>)
#include <cmath>
#include <opencv2/core.hpp>
#include <opencv2/imgcodecs.hpp>

#include "utils/error.h"
#include "utils/log.h"

namespace approx {

inline constexpr f64 image_gamma = 2.2;  // utils.cpp:9

inline MultiChannelImage read_image(fs::path path)  // utils.cpp:16-35
{
    cv::Mat const file = cv::imread(path.string(), cv::IMREAD_COLOR);
    if (file.empty())
        throw utils::IOError("Failed to open image", path);
    MultiChannelImage out(3, (Eigen::Index)file.rows, (Eigen::Index)file.cols);
    f64 decode[256];  // 256 possible values: one pow each instead of three per pixel
    for (int v = 0; v < 256; ++v)
        decode[v] = std::pow(v / 255.0, 1.0 / image_gamma);
    for (int r = 0; r < file.rows; ++r)
        for (int c = 0; c < file.cols; ++c) {
            cv::Vec3b const bgr = file.at<cv::Vec3b>(r, c);
            for (int ch = 0; ch < 3; ++ch)
                out[(size_t)ch](r, c) = decode[bgr[2 - ch]];
        }
    return out;
}

inline std::optional<cv::Mat> image_list_to_cv(std::vector<MatX<f64>> const& channels)  // utils.cpp:37-59
{
    if (channels.size() != 3) {
        utils::log(utils::LogLevel::warn, "approx", "Image with less than 3 channels is not supported. (%zu channels provided)", channels.size());
        return {};
    }
    cv::Mat file((int)channels[0].rows(), (int)channels[0].cols(), CV_8UC3);
    for (int r = 0; r < file.rows; ++r)
        for (int c = 0; c < file.cols; ++c) {
            cv::Vec3b bgr;
            for (int ch = 0; ch < 3; ++ch)
                bgr[2 - ch] = static_cast<unsigned char>(std::pow(channels[(size_t)ch](r, c), image_gamma) * 255.0);
            file.at<cv::Vec3b>(r, c) = bgr;
        }
    return file;
}

inline void write_image(std::vector<MatX<f64>> const& channels, fs::path const& output_path)  // utils.cpp:61-68
{
    if (auto const file = image_list_to_cv(channels))
        cv::imwrite(output_path.string(), *file);
}

}  // namespace approx
#endif
