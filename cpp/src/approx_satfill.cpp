// Host-side C++ shim: the reference's `approx` API (laplace.h:20,28; poisson.h:41-52) implemented by calls into the
// C-ABI of libsatfill.so.  No arithmetic happens here -- Eigen is only the container type of the signatures.
#include <approx/laplace.h>
#include <approx/poisson.h>
#include <utils/log.h>

#include <satfill.h>

#include <climits>
#include <cstdio>
#include <fstream>
#include <algorithm>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>

namespace approx {
namespace {

struct Ctx {
    sa_ctx* h = nullptr;
    std::mutex lock;  // a context serves one host thread at a time
    Ctx()
    {
        int device = 0;
        if (const char* e = std::getenv("SATFILL_DEVICE"))
            device = std::atoi(e);
        if (sa_create(&h, device, nullptr) != SA_OK)
            throw std::runtime_error("satfill: no usable CUDA device (there is no CPU fallback)");
    }
    ~Ctx() { sa_destroy(h); }
};

Ctx& ctx()
{
    static Ctx c;
    return c;
}

LaplaceOptions g_laplace;
PerfInfo g_perf;

static_assert(sizeof(bool) == 1, "MatX<bool> is passed to the C-ABI as a byte mask");

}  // namespace

void set_laplace_options(LaplaceOptions const& options) { g_laplace = options; }
LaplaceOptions const& laplace_options() { return g_laplace; }
PerfInfo const& last_perf_info() { return g_perf; }

void PerfInfo::write(fs::path const& output) const
{
    std::ofstream f(output, std::ios::app);
    f << region_size << ',' << tolerance << ',' << max_iterations << ',' << iterations << ',' << error << ','
      << solve_time << '\n';
}

void fill_missing_portion_smooth_boundary(MatX<f64>& input_image, MatX<bool> const& invalid_pixels)
{
    if (input_image.size() != invalid_pixels.size())  // laplace.cpp:124-127
        throw std::runtime_error("Input image and mask need to be the same size");
    if (input_image.rows() != invalid_pixels.rows())
        throw std::runtime_error("Input image and mask need to be the same shape");
    Ctx& c = ctx();
    std::lock_guard<std::mutex> guard(c.lock);
    sa_options o;
    sa_default_options(&o, SA_LAPLACE);
    if (g_laplace.tolerance > 0)
        o.tolerance = g_laplace.tolerance;
    if (g_laplace.max_iterations > 0)
        o.max_iterations = g_laplace.max_iterations;
    o.precond = g_laplace.multigrid ? SA_PRECOND_MULTIGRID : SA_PRECOND_JACOBI;
    double* band = input_image.data();
    sa_stats st {};
    // MatX is column-major: row stride 1, column stride rows
    int rc = sa_laplace_fill(c.h, &band, 1, reinterpret_cast<const uint8_t*>(invalid_pixels.data()), input_image.rows(),
        input_image.cols(), 1, input_image.rows(), &o, &st);
    if (rc != SA_OK && rc != SA_EMPTY_MASK && rc != SA_NOT_CONVERGED)  // the reference never checks info() here
        throw std::runtime_error(std::string("satfill: ") + sa_last_error(c.h));
    g_perf = PerfInfo { (long)st.unknowns, st.tolerance, (long)st.max_iterations, (long)st.iterations, st.error, st.solve_ms * 1e-3 };
    if (rc == SA_EMPTY_MASK)  // laplace.cpp:41-44
        utils::log(utils::LogLevel::info, "approx", "No invalid pixels found: nothing to do");
    else  // laplace.cpp:129-131 logs the elapsed time of the call
        utils::log(utils::LogLevel::info, "approx", "Laplace fill: %lld unknowns, %lld iterations, residual %.3e, %.3f ms on the device",
            (long long)st.unknowns, (long long)st.iterations, st.error, st.setup_ms + st.solve_ms);
}

void apply_laplace(const unsigned char* image, const unsigned char* invalid_image, Eigen::Index rows, Eigen::Index cols,
    f64 red_threshold, f64* out)
{
    if (rows <= 0 || cols <= 0)
        return;
    if (!image || !invalid_image || !out)
        throw std::runtime_error("apply_laplace: null buffer");
    Ctx& c = ctx();
    std::lock_guard<std::mutex> guard(c.lock);
    sa_options o;
    sa_default_options(&o, SA_LAPLACE);
    if (g_laplace.tolerance > 0)
        o.tolerance = g_laplace.tolerance;
    if (g_laplace.max_iterations > 0)
        o.max_iterations = g_laplace.max_iterations;
    o.precond = g_laplace.multigrid ? SA_PRECOND_MULTIGRID : SA_PRECOND_JACOBI;
    sa_stats st[3] {};
    int rc = sa_apply_laplace_u8(c.h, image, invalid_image, rows, cols, 3, red_threshold, out, nullptr, &o, st);
    if (rc != SA_OK && rc != SA_EMPTY_MASK && rc != SA_NOT_CONVERGED)  // the reference never checks info() here
        throw std::runtime_error(std::string("satfill: ") + sa_last_error(c.h));
    if (rc == SA_EMPTY_MASK) {  // laplace.cpp:41-44 per channel: nothing to fill, the image comes back as doubles
        for (Eigen::Index i = 0; i < rows * cols * 3; ++i)
            out[i] = (f64)image[i];
    }
}

void blend_images_poisson(MultiChannelImage& input_images, MultiChannelImage const& replacement_images,
    MatX<bool> const& invalid_mask, f64 tolerance, std::optional<int> max_iterations)
{
    if (input_images.images.empty())
        return;
    if (input_images.images.size() != replacement_images.images.size() || input_images.rows() != replacement_images.rows()
        || input_images.cols() != replacement_images.cols()) {  // poisson.cpp:154-157: log and return
        utils::log(utils::LogLevel::err, "approx", "Input and replacement images must have the same dimensions");
        return;
    }
    if (invalid_mask.rows() != input_images.rows() || invalid_mask.cols() != input_images.cols()) {
        // poisson.cpp:158-160 logs and continues (out-of-bounds reads); return instead
        utils::log(utils::LogLevel::err, "approx", "Invalid mask must match the image dimensions");
        return;
    }
    Ctx& c = ctx();
    std::lock_guard<std::mutex> guard(c.lock);
    const int nb = (int)input_images.images.size();
    std::vector<double*> in(nb);
    std::vector<const double*> rep(nb);
    for (int b = 0; b < nb; ++b) {
        in[b] = input_images.images[b].data();
        rep[b] = replacement_images.images[b].data();
    }
    sa_options o;
    sa_default_options(&o, SA_POISSON);
    o.tolerance = tolerance;
    o.max_iterations = max_iterations ? *max_iterations : 0;  // 0 -> n / 2 (poisson.cpp:207)
    std::vector<sa_stats> st(nb);
    int rc = sa_poisson_blend(c.h, in.data(), rep.data(), nb, reinterpret_cast<const uint8_t*>(invalid_mask.data()),
        input_images.rows(), input_images.cols(), 1, input_images.rows(), &o, st.data());
    const sa_stats& last = st[nb - 1];  // the reference keeps the last band's numbers (poisson.cpp:259-261)
    g_perf = PerfInfo { (long)last.unknowns, last.tolerance, (long)last.max_iterations, (long)last.iterations, last.error,
        last.solve_ms * 1e-3 };
    if (rc == SA_NOT_CONVERGED)  // poisson.cpp:263-269: nothing was written
        utils::log(utils::LogLevel::err, "approx", "Failed to solve the linear system: no convergence");
    else if (rc != SA_OK && rc != SA_EMPTY_MASK)
        utils::log(utils::LogLevel::err, "approx", "%s", sa_last_error(c.h));
}

void blend_images_poisson(MultiChannelImage& input_images, MultiChannelImage const& replacement_images, int start_row,
    int start_column)
{
    // the three sanity checks of poisson.cpp:25-39: log and return
    if (input_images.images.empty() || replacement_images.images.size() < 3
        || replacement_images.images.size() < input_images.images.size())
        return;
    const Eigen::Index R = replacement_images.rows(), C = replacement_images.cols();
    if (replacement_images.size() > input_images.size()) {
        utils::log(utils::LogLevel::err, "approx", "Cannot solve problem: replacement image is larger than the input image");
        return;
    }
    if (start_row < 0 || start_column < 0 || start_row >= input_images.rows() || start_column >= input_images.cols()) {
        utils::log(utils::LogLevel::err, "approx", "Cannot solve problem: row/column is out of bounds");
        return;
    }
    if (start_row + R > input_images.rows() || start_column + C > input_images.cols()) {
        utils::log(utils::LogLevel::err, "approx", "Cannot solve problem: replacement image goes beyond the bounds of the input image");
        return;
    }
    // The system lives in the replacement's own rectangle (neighbours outside it are dropped, poisson.cpp:76,108), its
    // unknowns are the non-key pixels, its boundary values come from the input at the offset: exactly the mask overload
    // on the crop.
    MatX<bool> unknown(R, C);
    for (Eigen::Index col = 0; col < C; ++col)
        for (Eigen::Index row = 0; row < R; ++row)
            unknown(row, col) = replacement_images.valid_pixel(row, col);
    const size_t nb = input_images.images.size();
    MultiChannelImage crop, repl;
    for (size_t b = 0; b < nb; ++b) {
        crop.images.push_back(input_images.images[b].block(start_row, start_column, R, C));
        repl.images.push_back(replacement_images.images[b]);
    }
    blend_images_poisson(crop, repl, unknown, 1e-12, std::optional<int>(INT_MAX / 2));
    for (size_t b = 0; b < nb; ++b)  // poisson.cpp:126-139: only the unknowns are written
        for (Eigen::Index col = 0; col < C; ++col)
            for (Eigen::Index row = 0; row < R; ++row)
                if (unknown(row, col))
                    input_images.images[b](start_row + row, start_column + col) = crop.images[b](row, col);
}

std::vector<MatX<f64>> blend_images_poisson(std::vector<MatX<f64>> const& input_images,
    std::vector<MatX<f64>> const& replacement_images, MatX<bool> const& invalid_mask, f64 tolerance,
    std::optional<int> max_iterations)
{
    MultiChannelImage input(input_images);  // poisson.cpp:292-303
    MultiChannelImage replacement(replacement_images);
    blend_images_poisson(input, replacement, invalid_mask, tolerance, max_iterations);
    return input.images;
}

MatX<bool> preprocess_cloud_band(MatX<f64> const& cloud_band, int dilation_size)
{
    MatX<bool> mask = MatX<bool>::Zero(cloud_band.rows(), cloud_band.cols());
    if (cloud_band.size() == 0)
        return mask;
    Ctx& c = ctx();
    std::lock_guard<std::mutex> guard(c.lock);
    // column-major Eigen storage: row stride 1, column stride rows; the mask comes back in the same layout
    int rc = sa_morph_close_mask(c.h, cloud_band.data(), cloud_band.rows(), cloud_band.cols(), 1, cloud_band.rows(),
        dilation_size, reinterpret_cast<uint8_t*>(mask.data()));
    if (rc != SA_OK)
        throw std::runtime_error(std::string("preprocess_cloud_band: ") + sa_last_error(c.h));
    return mask;
}

void highlight_area_replaced(MultiChannelImage& input_images, MultiChannelImage const& replacement_images, int start_row,
    int start_col, Vec3<f64> const& color)
{
    for (Eigen::Index row = 0; row < replacement_images.rows(); ++row)  // poisson.cpp:305-321
        for (Eigen::Index col = 0; col < replacement_images.cols(); ++col)
            if (replacement_images.valid_pixel(row, col))
                for (int c = 0; c < 3; ++c)
                    input_images(c, row + start_row, col + start_col) = color[c];
}

std::string find_good_close_image(std::string const& date_string, f64 distance_weight, std::vector<DayInfo> close_images,
    f64 percent_invalid_of_date)
{
    if (distance_weight < 0 || distance_weight > 1)  // poisson.cpp:325-327
        throw utils::GenericError("Could not find close image: distance weight not between 0 and 1");
    utils::Date date(date_string);
    if (close_images.empty())  // poisson.cpp:331-334
        return {};
    std::stable_sort(close_images.begin(), close_images.end(), [&](DayInfo const& a, DayInfo const& b) {
        return a.distance(date, distance_weight) < b.distance(date, distance_weight);
    });
    if (percent_invalid_of_date < close_images[0].percent_invalid)  // poisson.cpp:340-343: use the Laplace fill instead
        return date_string;
    std::ostringstream iso;  // to_iso_extended_string: YYYY-MM-DD
    iso << close_images[0].date;
    return iso.str();
}

ConnectedComponents find_connected_components(MatX<bool> const& invalid)
{
    ConnectedComponents out;
    const Eigen::Index rows = invalid.rows(), cols = invalid.cols();
    out.matrix = MatX<int>::Zero(rows, cols);
    if (rows == 0 || cols == 0)
        return out;
    Ctx& c = ctx();
    std::lock_guard<std::mutex> guard(c.lock);
    std::vector<int32_t> labels((size_t)rows * cols);  // dense row-major table from the device
    int32_t k = 0;
    int rc = sa_label_components(c.h, reinterpret_cast<const uint8_t*>(invalid.data()), rows, cols, 1, rows, labels.data(), &k);
    if (rc != SA_OK)
        throw std::runtime_error(std::string("satfill: ") + sa_last_error(c.h));
    for (Eigen::Index r = 0; r < rows; ++r)  // container bookkeeping only: raster order = the contract's order
        for (Eigen::Index col = 0; col < cols; ++col) {
            int l = labels[(size_t)r * cols + col];
            out.matrix(r, col) = l;
            if (l)
                out.region_map[l].push_back({ r, col });
        }
    return out;
}

}  // namespace approx
