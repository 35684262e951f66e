// TEST INFRASTRUCTURE -- a minimal cv::Mat (the members approx::apply_laplace(cv::Mat ...) touches), so that the
// OpenCV-gated part of cpp/include/approx/laplace.h compiles and runs in an image without OpenCV's C++ headers.
#pragma once
#include <cstring>
#include <memory>
#define CV_8U 0
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_64FC3 CV_MAKETYPE(CV_64F, 3)
namespace cv {
class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
    Mat() = default;
    Mat(int r, int c, int type) : rows(r), cols(c), type_(type), store_(new unsigned char[(size_t)r * c * elem()]) { data = store_.get(); }
    int type() const { return type_; }
    bool isContinuous() const { return true; }
    Mat clone() const
    {
        Mat m(rows, cols, type_);
        std::memcpy(m.data, data, (size_t)rows * cols * elem());
        return m;
    }
    size_t elem() const { return (size_t)(((type_ >> 3) + 1) * ((type_ & 7) == CV_64F ? 8 : 1)); }

private:
    int type_ = 0;
    std::shared_ptr<unsigned char[]> store_;
};
}  // namespace cv
