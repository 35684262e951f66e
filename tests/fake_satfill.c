/* TEST INFRASTRUCTURE -- a stand-in `libsatfill.so` for CPU-only tests of host programs that link the C-ABI
 * (tests/test_drivers.py runs cpp/src/poisson_main.cpp against it through LD_LIBRARY_PATH).  It implements the entry points
 * the C++ `approx` shim and the plain-C examples call, by calling the ORACLE (oracle/_build/liboracle.so: the plain-C
 * restatement of the reference), so that everything around the device calls -- GeoTIFF decode, memory layouts, band order, status handling,
 * the output file -- can be checked end to end on a machine without a GPU.  It is never built into, shipped with or
 * loaded by the product: the product library fails loudly without a device. */
#include <stdlib.h>
#include <string.h>

#include "satfill.h"

typedef struct {
    int64_t unknowns, system_size, iterations;
    double error, assemble_s, solve_s;
} so_stats;
int so_laplace_fill(double* img, const uint8_t* mask, int64_t rows, int64_t cols, int64_t rs, int64_t cs, int mode, double tol,
    int64_t max_it, so_stats* st);
int so_poisson_blend(double* const* inputs, const double* const* replacements, int nbands, const uint8_t* mask, int64_t rows,
    int64_t cols, int64_t rs, int64_t cs, double tol, int64_t max_it, so_stats* per_band);
int32_t so_label_components(const uint8_t* mask, int64_t rows, int64_t cols, int64_t rs, int64_t cs, int32_t* labels);
int64_t so_unknown_numbering(const uint8_t* mask, int64_t rows, int64_t cols, int64_t rs, int64_t cs, int32_t* number);

struct sa_ctx {
    int dummy;
};

int sa_create(sa_ctx** out, int device, void* stream)
{
    (void)device, (void)stream;
    *out = (sa_ctx*)calloc(1, sizeof(sa_ctx));
    return *out ? SA_OK : SA_OUT_OF_MEMORY;
}
void sa_destroy(sa_ctx* ctx) { free(ctx); }
const char* sa_last_error(const sa_ctx* ctx)
{
    (void)ctx;
    return "fake_satfill";
}
void sa_default_options(sa_options* o, int problem)
{
    memset(o, 0, sizeof(*o));
    o->tolerance = problem == SA_POISSON ? 1e-6 : 2.220446049250313e-16;
}

static void fill_stats(sa_stats* dst, const so_stats* src, int n, int status, const sa_options* o)
{
    if (!dst)
        return;
    for (int i = 0; i < n; ++i) {
        memset(&dst[i], 0, sizeof(dst[i]));
        dst[i].unknowns = src[i].unknowns;
        dst[i].iterations = src[i].iterations;
        dst[i].error = src[i].error;
        dst[i].tolerance = o ? o->tolerance : 0.0;
        dst[i].solve_ms = src[i].solve_s * 1e3;
        dst[i].status = status;
    }
}

int sa_laplace_fill(sa_ctx* ctx, double* const* images, int nbands, const uint8_t* mask, int64_t rows, int64_t cols,
    int64_t row_stride, int64_t col_stride, const sa_options* opts, sa_stats* stats)
{
    (void)ctx;
    int worst = SA_OK;
    for (int b = 0; b < nbands; ++b) {
        so_stats st;
        int rc = so_laplace_fill(images[b], mask, rows, cols, row_stride, col_stride, 1, opts ? opts->tolerance : 0.0,
            opts ? opts->max_iterations : 0, &st);
        fill_stats(stats ? stats + b : NULL, &st, 1, rc, opts);
        if (rc > worst)
            worst = rc;
    }
    return worst;
}

int sa_poisson_blend(sa_ctx* ctx, double* const* inputs, const double* const* replacements, int nbands, const uint8_t* mask,
    int64_t rows, int64_t cols, int64_t row_stride, int64_t col_stride, const sa_options* opts, sa_stats* stats)
{
    (void)ctx;
    so_stats* st = (so_stats*)calloc((size_t)(nbands > 0 ? nbands : 1), sizeof(so_stats));
    int64_t max_it = opts && opts->max_iterations > 0 ? opts->max_iterations : -1;
    int rc = so_poisson_blend(inputs, replacements, nbands, mask, rows, cols, row_stride, col_stride,
        opts ? opts->tolerance : 1e-6, max_it, st);
    fill_stats(stats, st, nbands, rc, opts);
    free(st);
    return rc;
}

/* approx::apply_laplace (laplace.cpp:134-168): mask from the marked image, every channel filled with it */
int sa_apply_laplace_u8(sa_ctx* ctx, const uint8_t* image, const uint8_t* invalid, int64_t rows, int64_t cols, int channels,
    double red_threshold, double* out, uint8_t* mask_out, const sa_options* opts, sa_stats* stats)
{
    (void)ctx;
    const int64_t n = rows * cols;
    uint8_t* mask = (uint8_t*)malloc((size_t)(n > 0 ? n : 1));
    double* plane = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    if (!mask || !plane)
        return SA_OUT_OF_MEMORY;
    for (int64_t i = 0; i < n; ++i)
        mask[i] = invalid[i * channels + 2] >= red_threshold && invalid[i * channels + 1] <= 150;
    int worst = SA_OK;
    for (int ch = 0; ch < channels; ++ch) {
        for (int64_t i = 0; i < n; ++i)
            plane[i] = image[i * channels + ch];
        so_stats st;
        int rc = so_laplace_fill(plane, mask, rows, cols, cols, 1, 1, opts ? opts->tolerance : 0.0, opts ? opts->max_iterations : 0, &st);
        fill_stats(stats ? stats + ch : NULL, &st, 1, rc, opts);
        if (rc > worst)
            worst = rc;
        for (int64_t i = 0; i < n; ++i)
            out[i * channels + ch] = plane[i];
    }
    if (mask_out)
        memcpy(mask_out, mask, (size_t)n);
    free(mask);
    free(plane);
    return worst;
}

int sa_label_components(sa_ctx* ctx, const uint8_t* mask, int64_t rows, int64_t cols, int64_t row_stride, int64_t col_stride,
    int32_t* labels, int32_t* out_num_labels)
{
    (void)ctx;
    *out_num_labels = so_label_components(mask, rows, cols, row_stride, col_stride, labels);
    return SA_OK;
}

int sa_unknown_numbering(sa_ctx* ctx, const uint8_t* mask, int64_t rows, int64_t cols, int64_t row_stride, int64_t col_stride,
    int32_t* numbering, int64_t* out_count)
{
    (void)ctx;
    *out_count = so_unknown_numbering(mask, rows, cols, row_stride, col_stride, numbering);
    return SA_OK;
}

/* cv::morphologyEx(MORPH_CLOSE) with a (2 radius + 1)^2 rectangle, windows clipped at the border, then != 0 */
int sa_morph_close_mask(sa_ctx* ctx, const double* band, int64_t rows, int64_t cols, int64_t row_stride, int64_t col_stride,
    int radius, uint8_t* mask_out)
{
    (void)ctx;
    double* dil = (double*)malloc(sizeof(double) * (size_t)(rows * cols));
    if (!dil)
        return SA_OUT_OF_MEMORY;
    for (int64_t r = 0; r < rows; ++r)
        for (int64_t c = 0; c < cols; ++c) {
            double m = band[r * row_stride + c * col_stride];
            for (int64_t y = r - radius < 0 ? 0 : r - radius; y <= r + radius && y < rows; ++y)
                for (int64_t x = c - radius < 0 ? 0 : c - radius; x <= c + radius && x < cols; ++x)
                    if (band[y * row_stride + x * col_stride] > m)
                        m = band[y * row_stride + x * col_stride];
            dil[r * cols + c] = m;
        }
    for (int64_t r = 0; r < rows; ++r)
        for (int64_t c = 0; c < cols; ++c) {
            double m = dil[r * cols + c];
            for (int64_t y = r - radius < 0 ? 0 : r - radius; y <= r + radius && y < rows; ++y)
                for (int64_t x = c - radius < 0 ? 0 : c - radius; x <= c + radius && x < cols; ++x)
                    if (dil[y * cols + x] < m)
                        m = dil[y * cols + x];
            mask_out[r * row_stride + c * col_stride] = m != 0.0;
        }
    free(dil);
    return SA_OK;
}
