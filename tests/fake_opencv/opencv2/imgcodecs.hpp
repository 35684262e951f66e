// TEST INFRASTRUCTURE -- cv::imread / cv::imwrite for 8-bit colour images over binary PPM ("P6", stored R G B, handed out
// B G R like OpenCV does), enough for approx::read_image / write_image (cpp/include/approx/utils.h) to be compiled and run.
#pragma once
#include <cstdio>
#include <string>

#include "core.hpp"
namespace cv {
enum { IMREAD_COLOR = 1 };
inline Mat imread(std::string const& path, int = IMREAD_COLOR)
{
    std::FILE* f = std::fopen(path.c_str(), "rb");
    if (!f)
        return Mat();
    int w = 0, h = 0, maxv = 0;
    if (std::fscanf(f, "P6 %d %d %d", &w, &h, &maxv) != 3 || maxv != 255 || std::fgetc(f) == EOF) {
        std::fclose(f);
        return Mat();
    }
    Mat m(h, w, CV_8UC3);
    for (int i = 0; i < w * h; ++i) {
        unsigned char rgb[3];
        if (std::fread(rgb, 1, 3, f) != 3) {
            std::fclose(f);
            return Mat();
        }
        m.data[3 * i + 0] = rgb[2], m.data[3 * i + 1] = rgb[1], m.data[3 * i + 2] = rgb[0];
    }
    std::fclose(f);
    return m;
}
inline bool imwrite(std::string const& path, Mat const& m)
{
    std::FILE* f = std::fopen(path.c_str(), "wb");
    if (!f)
        return false;
    std::fprintf(f, "P6\n%d %d\n255\n", m.cols, m.rows);
    for (int i = 0; i < m.rows * m.cols; ++i) {
        const unsigned char rgb[3] = { m.data[3 * i + 2], m.data[3 * i + 1], m.data[3 * i + 0] };
        std::fwrite(rgb, 1, 3, f);
    }
    return std::fclose(f) == 0;
}
}  // namespace cv
