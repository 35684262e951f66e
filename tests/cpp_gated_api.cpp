// TEST INFRASTRUCTURE -- compiled by tests/test_host_api.py with -Itests/fake_opencv (a minimal cv::Mat) and run against the
// fake C-ABI: the parts of the reference's C++ API that are gated on headers this image lacks --
//   cv::Mat approx::apply_laplace(cv::Mat const&, cv::Mat const&, f64)                          (laplace.h:31)
//   std::string approx::find_good_close_image(std::string const&, f64, DataBase&)               (poisson.h:63)
//   approx::read_image / image_list_to_cv / write_image                                         (utils.h:108-110)
// Prints the filled image (doubles, one per line) after a header line, and the picker's answers.
#include <approx/laplace.h>
#include <approx/poisson.h>
#include <approx/utils.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

struct FakeRow {
    utils::Date date;
    double percent_invalid;
};
struct FakeDataBase {  // the two queries of approx::DataBase (db.h:30-31)
    std::vector<FakeRow> rows;
    FakeRow current;
    int queries = 0;
    std::vector<FakeRow> select_close_images(std::string const&) { ++queries; return rows; }
    FakeRow select_info_about_date(std::string const&) { ++queries; return current; }
};

int main(int argc, char** argv)
{
    const int rows = std::atoi(argv[1]), cols = std::atoi(argv[2]);
    cv::Mat image(rows, cols, CV_8UC3), invalid(rows, cols, CV_8UC3);
    unsigned seed = 12345;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return seed >> 24; };
    for (int i = 0; i < rows * cols * 3; ++i) {
        image.data[i] = (unsigned char)rnd();
        invalid.data[i] = 0;
    }
    for (int r = 3; r < rows - 3; ++r)
        for (int c = 4; c < cols - 5; ++c)
            if ((r / 5 + c / 7) % 3 == 0) {  // a few blobs marked red (R >= 220, G <= 150), B G R order
                invalid.data[(r * cols + c) * 3 + 2] = 255;
                invalid.data[(r * cols + c) * 3 + 1] = 20;
            }
    approx::set_laplace_options({ 1e-12, 0, true });
    cv::Mat out = approx::apply_laplace(image, invalid, 220.0);
    if (out.type() != CV_64FC3 || out.rows != rows || out.cols != cols)
        return 3;
    std::printf("image %d %d\n", rows, cols);
    for (int i = 0; i < rows * cols * 3; ++i)
        std::printf("%d %d %.10f\n", (int)image.data[i], (int)invalid.data[i], reinterpret_cast<double*>(out.data)[i]);
    bool threw = false;
    try {
        cv::Mat small(rows - 1, cols, CV_8UC3);
        approx::apply_laplace(image, small, 220.0);
    } catch (std::runtime_error const&) {  // laplace.cpp:124-127
        threw = true;
    }
    FakeDataBase db;
    db.rows = { { utils::Date("2019-05-12"), 0.50 }, { utils::Date("2019-05-20"), 0.10 }, { utils::Date("2019-06-01"), 0.02 } };
    db.current = { utils::Date("2019-05-22"), 0.30 };
    std::printf("picker %d %s %s", threw ? 1 : 0, approx::find_good_close_image("2019-05-22", 1.0, db).c_str(),
        approx::find_good_close_image("2019-05-22", 0.0, db).c_str());
    db.current.percent_invalid = 0.01;
    std::printf(" %s", approx::find_good_close_image("2019-05-22", 0.5, db).c_str());
    FakeDataBase empty;
    const std::string none = approx::find_good_close_image("2019-05-22", 0.5, empty);
    std::printf(" [%s] %d", none.c_str(), empty.queries);
    int q0 = db.queries;
    try {
        approx::find_good_close_image("2019-05-22", 1.5, db);
    } catch (utils::GenericError const&) {  // poisson.cpp:325-327, before the database is touched
        std::printf(" weight-error %d", db.queries - q0);
    }
    std::printf("\n");
    // image files (argv[3]: a directory to write into): a 256-level ramp survives write -> read bit for bit except where
    // the truncating encode lands one level low; the channel order is R, G, B on the C++ side and B, G, R in the file
    if (argc > 3) {
        const std::string dir = argv[3];
        approx::MultiChannelImage ramp(3, 4, 256);
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 256; ++c) {
                ramp[0](r, c) = std::pow(c / 255.0, 1.0 / approx::image_gamma);
                ramp[1](r, c) = std::pow((255 - c) / 255.0, 1.0 / approx::image_gamma);
                ramp[2](r, c) = r == 0 ? 0.0 : 1.0;
            }
        approx::write_image(ramp.images, dir + "/ramp.ppm");
        approx::MultiChannelImage back = approx::read_image(dir + "/ramp.ppm");
        int worst = 0, exact = 0;
        for (int ch = 0; ch < 3; ++ch)
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 256; ++c) {
                    const int want = (int)std::lround(std::pow(ramp[(size_t)ch](r, c), approx::image_gamma) * 255.0);
                    const int got = (int)std::lround(std::pow(back[(size_t)ch](r, c), approx::image_gamma) * 255.0);
                    worst = std::max(worst, std::abs(want - got));
                    exact += want == got;
                }
        auto const mat = approx::image_list_to_cv(ramp.images);
        const cv::Vec3b px = mat->at<cv::Vec3b>(2, 255);  // R = 1, G = 0, B = 1  ->  B G R = 255 0 255
        bool threw_io = false;
        try {
            approx::read_image(dir + "/missing.ppm");
        } catch (utils::IOError const&) {  // utils.cpp:19-21
            threw_io = true;
        }
        std::vector<MatX<f64>> two(2, MatX<f64>::Zero(2, 2));
        approx::write_image(two, dir + "/two.ppm");  // logged, nothing written (utils.cpp:39-42, 64-66)
        std::FILE* f = std::fopen((dir + "/two.ppm").c_str(), "rb");
        std::printf("images %d %d %d %d %d %d %d %d %d\n", (int)back.rows(), (int)back.cols(), worst, exact, (int)px[0], (int)px[1], (int)px[2],
            threw_io ? 1 : 0, f ? 1 : 0);
        if (f)
            std::fclose(f);
    }
    return 0;
}
