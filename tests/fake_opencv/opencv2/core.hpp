// TEST INFRASTRUCTURE -- a minimal cv::Mat (the members approx::apply_laplace(cv::Mat ...) and approx::read_image /
// image_list_to_cv touch), so that the OpenCV-gated parts of cpp/include/approx/{laplace,utils}.h compile and run in an image
// without OpenCV's C++ headers.
#pragma once
#include <cstring>
#include <memory>
#define CV_8U 0
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_64FC3 CV_MAKETYPE(CV_64F, 3)
namespace cv {
struct Vec3b {
    unsigned char v[3] = { 0, 0, 0 };
    unsigned char& operator[](int i) { return v[i]; }
    unsigned char const& operator[](int i) const { return v[i]; }
};
class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
    Mat() = default;
    Mat(int r, int c, int type) : rows(r), cols(c), type_(type), store_(new unsigned char[(size_t)r * c * elem()]) { data = store_.get(); }
    int type() const { return type_; }
    bool isContinuous() const { return true; }
    bool empty() const { return data == nullptr || rows * cols == 0; }
    template <typename T>
    T& at(int r, int c) { return reinterpret_cast<T*>(data)[(size_t)r * cols + c]; }
    template <typename T>
    T const& at(int r, int c) const { return reinterpret_cast<T const*>(data)[(size_t)r * cols + c]; }
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
