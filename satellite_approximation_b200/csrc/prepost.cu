// The steps either side of the fill path, as plain HBM-bound element-wise / small-window kernels:
//
//   * approx::apply_laplace (lib/approx/source/laplace.cpp:134-168; laplace_main): mask from the red / green channels of
//     an 8-bit colour image, the three channels of the base image widened to double planes of ONE scene (so that the
//     three fills share one index, one hierarchy and every launch), and the planes interleaved back into the CV_64FC3
//     layout the reference returns;
//   * preprocess_cloud_band (executables/poisson-main.cpp:10-21): morphological closing with a (2 radius + 1)^2
//     rectangle -- separable, so four passes of a running max / min along one axis -- and the cast to bool.
//
// A rectangle is symmetric under transposition, so, like the solver, the morphology runs on the caller's buffer as it
// lies whether it is row- or column-major.
#include "common.cuh"

#include <cfloat>

namespace satfill {

namespace {

// 8-bit interleaved (rows x cols x CH) -> mask plane + CH double planes
template <int CH>
__global__ void __launch_bounds__(256) k_split_u8(const uint8_t* __restrict__ image, const uint8_t* __restrict__ invalid,
    int64_t rows, int64_t cols, int64_t pitch, int64_t plane, double red_threshold, double* __restrict__ u0,
    uint8_t* __restrict__ mask0)
{
    const int64_t n = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        const uint8_t* px = image + i * CH;
        const uint8_t* iv = invalid + i * CH;
        // cv::imread(IMREAD_COLOR) is B, G, R: channels_cv[2] is red, channels_cv[1] green (laplace.cpp:141-146)
        mask0[r * pitch + c] = ((double)iv[2] >= red_threshold && iv[1] <= 150) ? 1 : 0;
#pragma unroll
        for (int k = 0; k < CH; ++k)
            u0[(int64_t)k * plane + r * pitch + c] = (double)px[k];  // cv2eigen of an 8-bit channel, laplace.cpp:155-156
    }
}

// CH double planes -> interleaved rows x cols x CH doubles (cv::merge of the filled channels, laplace.cpp:164-165)
template <int CH>
__global__ void __launch_bounds__(256) k_merge_f64(const double* __restrict__ u0, int64_t rows, int64_t cols, int64_t pitch,
    int64_t plane, double* __restrict__ out)
{
    const int64_t n = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
#pragma unroll
        for (int k = 0; k < CH; ++k)
            out[i * CH + k] = u0[(int64_t)k * plane + r * pitch + c];
    }
}

// running max (MAX) or min over [x - radius, x + radius] clipped to the extent, along the fast (axis 1) or slow (axis 0)
// axis of a dense slow x fast array
template <bool MAX>
__global__ void __launch_bounds__(256) k_morph_pass(const double* __restrict__ src, double* __restrict__ dst, int64_t slow,
    int64_t fast, int radius, int axis)
{
    const int64_t n = slow * fast;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = i / fast, f = i - s * fast;
        const int64_t at = axis ? f : s, extent = axis ? fast : slow, step = axis ? 1 : fast;
        const int64_t lo = at - radius < 0 ? 0 : at - radius, hi = at + radius >= extent ? extent - 1 : at + radius;
        const double* p = src + i + (lo - at) * step;
        double v = *p;
        for (int64_t k = lo + 1; k <= hi; ++k) {
            p += step;
            const double w = *p;
            v = MAX ? (w > v ? w : v) : (w < v ? w : v);
        }
        dst[i] = v;
    }
}

__global__ void __launch_bounds__(256) k_nonzero_mask(const double* __restrict__ src, uint8_t* __restrict__ dst, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[i] != 0.0 ? 1 : 0;  // MatX<f64>::cast<bool>() (poisson-main.cpp:20)
}

unsigned grid_for(const sa_ctx* ctx, int64_t n)
{
    int64_t want = (n + 255) / 256, cap = (int64_t)ctx->sm_count * 16;
    return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace

int split_u8_scene(sa_scene* s, const uint8_t* d_image, const uint8_t* d_invalid, int channels, double red_threshold)
{
    sa_ctx* ctx = s->ctx;
    if (channels != 3)
        return fail(ctx, SA_BAD_ARGUMENT, "apply_laplace: cv::imread(IMREAD_COLOR) images have 3 channels");
    SA_LAUNCH(ctx, k_split_u8<3>, grid_for(ctx, s->rows * s->cols), 256, 0, d_image, d_invalid, s->rows, s->cols, s->pitch,
        s->plane, red_threshold, s->plane0(s->u, 0), s->mask0(s->mask));
    SA_CUDA(ctx, cudaGetLastError());
    return SA_OK;
}

int merge_f64_scene(sa_scene* s, int channels, double* d_out)
{
    sa_ctx* ctx = s->ctx;
    if (channels != 3)
        return fail(ctx, SA_BAD_ARGUMENT, "apply_laplace: 3 channels");
    SA_LAUNCH(ctx, k_merge_f64<3>, grid_for(ctx, s->rows * s->cols), 256, 0, s->plane0(s->u, 0), s->rows, s->cols, s->pitch,
        s->plane, d_out);
    SA_CUDA(ctx, cudaGetLastError());
    return SA_OK;
}

// d_a holds the band (dense slow x fast); d_b is scratch of the same size; the mask lands in d_mask
int morph_close_mask(sa_ctx* ctx, double* d_a, double* d_b, int64_t slow, int64_t fast, int radius, uint8_t* d_mask)
{
    const int64_t n = slow * fast;
    const unsigned g = grid_for(ctx, n);
    SA_LAUNCH(ctx, k_morph_pass<true>, g, 256, 0, d_a, d_b, slow, fast, radius, 1);   // dilate
    SA_LAUNCH(ctx, k_morph_pass<true>, g, 256, 0, d_b, d_a, slow, fast, radius, 0);
    SA_LAUNCH(ctx, k_morph_pass<false>, g, 256, 0, d_a, d_b, slow, fast, radius, 1);  // erode
    SA_LAUNCH(ctx, k_morph_pass<false>, g, 256, 0, d_b, d_a, slow, fast, radius, 0);
    SA_LAUNCH(ctx, k_nonzero_mask, g, 256, 0, d_a, d_mask, n);
    SA_CUDA(ctx, cudaGetLastError());
    return SA_OK;
}

}  // namespace satfill
