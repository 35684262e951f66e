// TEST INFRASTRUCTURE -- compiled by tests/test_host_api.py with -Itests/fake_opencv (a minimal cv::Mat) and run against the
// fake C-ABI: the parts of the reference's C++ API that are gated on headers this image lacks --
//   cv::Mat approx::apply_laplace(cv::Mat const&, cv::Mat const&, f64)                          (laplace.h:31)
//   std::string approx::find_good_close_image(std::string const&, f64, DataBase&)               (poisson.h:63)
// Prints the filled image (doubles, one per line) after a header line, and the picker's answers.
#include <approx/laplace.h>
#include <approx/poisson.h>

#include <cstdio>
#include <cstdlib>
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
    return 0;
}
