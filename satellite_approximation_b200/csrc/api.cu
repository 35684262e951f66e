// C-ABI of libsatfill.so (include/satfill.h): contexts, device-resident scenes and the host-pointer entry points that
// the C++ `approx` shim, the pybind11 module and the ctypes binding call.  No compute happens on the host: every
// entry point either moves bytes or launches kernels of this library, and fails with SA_CUDA_ERROR when there is no
// usable device.
#include "common.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <new>

using namespace satfill;

namespace {

enum Layout { LAYOUT_BAD = 0, LAYOUT_ROW_MAJOR = 1, LAYOUT_COL_MAJOR = 2 };

Layout classify(int64_t rows, int64_t cols, int64_t rs, int64_t cs)
{
    bool row_major = (cs == 1 || cols <= 1) && (rs >= cols || rows <= 1);
    bool col_major = (rs == 1 || rows <= 1) && (cs >= rows || cols <= 1);
    if (row_major)
        return LAYOUT_ROW_MAJOR;
    if (col_major)
        return LAYOUT_COL_MAJOR;
    return LAYOUT_BAD;
}

int check_ctx(sa_ctx* ctx)
{
    if (!ctx)
        return SA_BAD_ARGUMENT;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess)
        return fail(ctx, SA_CUDA_ERROR, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return SA_OK;
}

// Upload a host mask (any of the two layouts) as a dense row-major rows x cols device table (pitch = cols).
int upload_mask_row_major(sa_ctx* ctx, const uint8_t* mask, int64_t rows, int64_t cols, int64_t rs, int64_t cs,
    uint8_t** d_out)
{
    *d_out = nullptr;
    Layout lay = classify(rows, cols, rs, cs);
    if (lay == LAYOUT_BAD)
        return fail(ctx, SA_BAD_ARGUMENT, "mask strides: one of row_stride / col_stride must be 1");
    if (rows == 0 || cols == 0)
        return SA_OK;
    uint8_t* d = nullptr;
    SA_CUDA(ctx, cudaMallocAsync(&d, (size_t)(rows * cols), ctx->stream));
    if (lay == LAYOUT_ROW_MAJOR) {
        SA_CUDA(ctx, cudaMemcpy2DAsync(d, (size_t)cols, mask, (size_t)(rows > 1 ? rs : cols), (size_t)cols, (size_t)rows,
                         cudaMemcpyHostToDevice, ctx->stream));
    } else {
        uint8_t* t = nullptr;  // cols x rows, row-major = the column-major source as it lies in memory
        SA_CUDA(ctx, cudaMallocAsync(&t, (size_t)(rows * cols), ctx->stream));
        SA_CUDA(ctx, cudaMemcpy2DAsync(t, (size_t)rows, mask, (size_t)(cols > 1 ? cs : rows), (size_t)rows, (size_t)cols,
                         cudaMemcpyHostToDevice, ctx->stream));
        SA_TRY(transpose_u8(ctx, t, cols, rows, rows, d, cols));
        SA_CUDA(ctx, cudaFreeAsync(t, ctx->stream));
    }
    *d_out = d;
    return SA_OK;
}

int scene_alloc(sa_scene* s, bool transposed)
{
    sa_ctx* ctx = s->ctx;
    s->transposed = transposed;
    s->rows = transposed ? s->user_cols : s->user_rows;
    s->cols = transposed ? s->user_rows : s->user_cols;
    s->pitch = round_up(s->cols + 1, TILE_W);
    s->rows_p = round_up(s->rows, TILE_H);
    if (s->rows_p == 0)
        s->rows_p = TILE_H;
    s->plane = (s->rows_p + 2) * s->pitch;
    s->tiles_x = (int)(s->pitch / TILE_W);
    s->tiles_y = (int)(s->rows_p / TILE_H);
    size_t vec = (size_t)s->plane * s->nbands * sizeof(double);
    SA_CUDA(ctx, cudaMalloc(&s->u, vec));
    SA_CUDA(ctx, cudaMalloc(&s->r, vec));
    SA_CUDA(ctx, cudaMemsetAsync(s->u, 0, vec, ctx->stream));
    if (s->problem == SA_POISSON) {
        SA_CUDA(ctx, cudaMalloc(&s->g, vec));
        SA_CUDA(ctx, cudaMemsetAsync(s->g, 0, vec, ctx->stream));
    }
    SA_CUDA(ctx, cudaMalloc(&s->mask, (size_t)s->plane));
    SA_CUDA(ctx, cudaMalloc(&s->umask, (size_t)s->plane));
    SA_CUDA(ctx, cudaMemsetAsync(s->mask, 0, (size_t)s->plane, ctx->stream));
    SA_CUDA(ctx, cudaMemsetAsync(s->umask, 0, (size_t)s->plane, ctx->stream));
    SA_CUDA(ctx, cudaMalloc(&s->tile_list, sizeof(int32_t) * 3 * (size_t)s->tiles_x * s->tiles_y));
    {
        size_t words = (size_t)(s->tiles_x + 2) * (s->tiles_y + 2) * 32;
        s->tb_words = words;
        SA_CUDA(ctx, cudaMalloc(&s->tbits, 2 * words * sizeof(uint32_t)));
        SA_CUDA(ctx, cudaMemsetAsync(s->tbits, 0, 2 * words * sizeof(uint32_t), ctx->stream));
    }
    SA_CUDA(ctx, cudaMalloc(&s->d_counters, sizeof(int32_t) * 4));
    SA_CUDA(ctx, cudaMalloc(&s->d_count64, sizeof(unsigned long long)));
    SA_CUDA(ctx, cudaMalloc(&s->scal, sizeof(BandScalars) * s->nbands));
    s->oriented = true;
    return SA_OK;
}

// Decide / check the resident orientation of a scene from the strides of a source or destination buffer.
int scene_orient(sa_scene* s, int64_t rs, int64_t cs)
{
    Layout lay = classify(s->user_rows, s->user_cols, rs, cs);
    if (lay == LAYOUT_BAD)
        return fail(s->ctx, SA_BAD_ARGUMENT, "strides: one of row_stride / col_stride must be 1");
    bool transposed = lay == LAYOUT_COL_MAJOR;
    if (!s->oriented)
        return scene_alloc(s, transposed);
    if (transposed != s->transposed) {
        // a buffer that is valid in both layouts (a single row or column) is fine
        Layout alt = transposed ? LAYOUT_ROW_MAJOR : LAYOUT_COL_MAJOR;
        bool both = (alt == LAYOUT_ROW_MAJOR)
            ? ((cs == 1 || s->user_cols <= 1) && (rs >= s->user_cols || s->user_rows <= 1))
            : ((rs == 1 || s->user_rows <= 1) && (cs >= s->user_rows || s->user_cols <= 1));
        if (!both)
            return fail(s->ctx, SA_BAD_ARGUMENT, "all buffers of one scene must share one memory layout");
    }
    return SA_OK;
}

// source pitch (elements) along the resident slow axis
int64_t slow_stride(const sa_scene* s, int64_t rs, int64_t cs)
{
    int64_t st = s->transposed ? cs : rs;
    return s->rows > 1 ? st : s->cols;
}

template <typename T>
int copy_in(sa_scene* s, T* dst0, const T* src, int64_t rs, int64_t cs, int on_device, int64_t row_lo, int64_t row_hi,
    cudaStream_t stream = nullptr)
{
    sa_ctx* ctx = s->ctx;
    if (!stream)
        stream = ctx->stream;
    if (row_hi <= row_lo || s->cols == 0)
        return SA_OK;
    int64_t sp = slow_stride(s, rs, cs);
    SA_CUDA(ctx, cudaMemcpy2DAsync(dst0 + row_lo * s->pitch, (size_t)s->pitch * sizeof(T), src + row_lo * sp,
                     (size_t)sp * sizeof(T), (size_t)s->cols * sizeof(T), (size_t)(row_hi - row_lo),
                     on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
    return SA_OK;
}

template <typename T>
int copy_out(sa_scene* s, const T* src0, T* dst, int64_t rs, int64_t cs, int on_device, int64_t row_lo, int64_t row_hi,
    cudaStream_t stream = nullptr)
{
    sa_ctx* ctx = s->ctx;
    if (!stream)
        stream = ctx->stream;
    if (row_hi <= row_lo || s->cols == 0)
        return SA_OK;
    int64_t dp = slow_stride(s, rs, cs);
    SA_CUDA(ctx, cudaMemcpy2DAsync(dst + row_lo * dp, (size_t)dp * sizeof(T), src0 + row_lo * s->pitch,
                     (size_t)s->pitch * sizeof(T), (size_t)s->cols * sizeof(T), (size_t)(row_hi - row_lo),
                     on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, stream));
    return SA_OK;
}

void scene_free(sa_scene* s)
{
    if (!s)
        return;
    if (s->ctx)
        cudaSetDevice(s->ctx->device);
    free_hierarchy(s);
    cudaFree(s->u);
    cudaFree(s->g);
    cudaFree(s->r);
    cudaFree(s->p[0]);
    cudaFree(s->p[1]);
    cudaFree(s->z);
    cudaFree(s->t);
    cudaFree(s->mask);
    cudaFree(s->umask);
    cudaFree(s->tile_list);
    cudaFree(s->tbits);
    cudaFree(s->d_counters);
    cudaFree(s->d_count64);
    cudaFree(s->scal);
    delete s;
}

}  // namespace

// The context caches the scene of the last host-pointer fill so that repeated calls on same-shaped inputs do not
// re-allocate HBM.
struct sa_ctx_cache {
    sa_scene* scene = nullptr;
};
static sa_ctx_cache* cache_of(sa_ctx* ctx);

struct sa_ctx_full : sa_ctx {
    sa_ctx_cache cache;
};
static sa_ctx_cache* cache_of(sa_ctx* ctx) { return &static_cast<sa_ctx_full*>(ctx)->cache; }

extern "C" {

int sa_abi_version(void) { return SATFILL_ABI_VERSION; }

int sa_has_legacy_variants(void) { return SATFILL_LEGACY_VARIANTS ? 1 : 0; }

int64_t sa_scene_plane_elements(int64_t rows, int64_t cols)
{
    if (rows < 0 || cols < 0)
        return -1;
    // the larger of the two orientations a scene can be resident in (scene_alloc)
    auto plane = [](int64_t r, int64_t c) {
        int64_t rp = round_up(r, TILE_H);
        if (rp == 0)
            rp = TILE_H;
        return (rp + 2) * round_up(c + 1, TILE_W);
    };
    return std::max(plane(rows, cols), plane(cols, rows));
}

int sa_create(sa_ctx** out, int device, void* stream)
{
    if (!out)
        return SA_BAD_ARGUMENT;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count)
        return SA_CUDA_ERROR;  // no CPU fallback: the library is unusable without a device
    if (cudaSetDevice(device) != cudaSuccess)
        return SA_CUDA_ERROR;
    sa_ctx_full* ctx = new (std::nothrow) sa_ctx_full();
    if (!ctx)
        return SA_OUT_OF_MEMORY;
    ctx->device = device;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return SA_CUDA_ERROR;
        }
        ctx->owns_stream = true;
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    ctx->grid_sms = ctx->sm_count;
    ctx->pinned_bytes = 1 << 20;
    if (cudaMallocHost(&ctx->pinned, ctx->pinned_bytes) != cudaSuccess) {
        if (ctx->owns_stream)
            cudaStreamDestroy(ctx->stream);
        delete ctx;
        return SA_OUT_OF_MEMORY;
    }
    for (auto& ev : ctx->ev)
        cudaEventCreate(&ev);
    *out = ctx;
    return SA_OK;
}

void sa_destroy(sa_ctx* ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    scene_free(cache_of(ctx)->scene);
    dist_shutdown(ctx);
    for (cudaEvent_t e : ctx->io_ev)
        cudaEventDestroy(e);
    if (ctx->io_in)
        cudaStreamDestroy(ctx->io_in);
    if (ctx->io_out)
        cudaStreamDestroy(ctx->io_out);
    for (auto& ev : ctx->ev)
        if (ev)
            cudaEventDestroy(ev);
    for (auto& ev : ctx->ev_pool)
        cudaEventDestroy(ev);
    if (ctx->pinned)
        cudaFreeHost(ctx->pinned);
    cudaFree(ctx->d_barrier);
    if (ctx->owns_stream)
        cudaStreamDestroy(ctx->stream);
    delete static_cast<sa_ctx_full*>(ctx);
}

const char* sa_last_error(const sa_ctx* ctx) { return ctx ? ctx->error.c_str() : "null context"; }

int64_t sa_kernel_launches(const sa_ctx* ctx) { return ctx ? ctx->launches : 0; }

void sa_default_options(sa_options* o, int problem)
{
    if (!o)
        return;
    std::memset(o, 0, sizeof(*o));
    // Laplace: Eigen's default tolerance is machine epsilon (IterativeSolverBase.h:367-368); Poisson: 1e-6 (poisson.h:45)
    o->tolerance = problem == SA_POISSON ? 1e-6 : DBL_EPSILON;
    o->max_iterations = 0;
    // The red-black multigrid V-cycle (mg_rb.cu) is the default preconditioner: the drop-in call takes the fast path.  It
    // changes the path CG takes, not its fixed point or its stop rule, and the float cycle does not limit the attainable
    // residual (tests: 1e-14).  SA_PRECOND_JACOBI -- Eigen's DiagonalPreconditioner, the reference's own -- is the opt-in.
    o->precond = SA_PRECOND_MULTIGRID;
    o->mg_variant = SA_MG_RB32;
    o->check_every = 0;
    o->mg_levels = 0;
    o->mg_smooth = 2;
}

int sa_synchronize(sa_ctx* ctx)
{
    SA_TRY(check_ctx(ctx));
    SA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SA_OK;
}

/* ---- integer path ------------------------------------------------------------------------------------------- */

int sa_mask_scan(sa_ctx* ctx, const uint8_t* mask, int64_t rows, int64_t cols, int64_t row_stride, int64_t col_stride,
    int64_t* out_pixels, int64_t capacity, int64_t* out_count, int64_t bbox[4])
{
    SA_TRY(check_ctx(ctx));
    if (rows < 0 || cols < 0 || (rows * cols > 0 && !mask))
        return fail(ctx, SA_BAD_ARGUMENT, "mask_scan: bad arguments");
    uint8_t* d_mask = nullptr;
    SA_TRY(upload_mask_row_major(ctx, mask, rows, cols, row_stride, col_stride, &d_mask));
    int64_t* d_pixels = nullptr;
    if (!out_pixels)
        capacity = 0;
    if (capacity > rows * cols)
        capacity = rows * cols;
    if (capacity > 0)
        SA_CUDA(ctx, cudaMallocAsync(&d_pixels, (size_t)capacity * 2 * sizeof(int64_t), ctx->stream));
    int64_t count = 0;
    int st = device_numbering(ctx, d_mask, rows, cols, cols, nullptr, d_pixels, capacity, &count, bbox);
    if (st == SA_OK && d_pixels) {
        int64_t ncopy = count < capacity ? count : capacity;
        if (ncopy > 0) {
            SA_CUDA(ctx, cudaMemcpyAsync(out_pixels, d_pixels, (size_t)ncopy * 2 * sizeof(int64_t), cudaMemcpyDeviceToHost,
                             ctx->stream));
        }
    }
    if (d_pixels)
        cudaFreeAsync(d_pixels, ctx->stream);
    if (d_mask)
        cudaFreeAsync(d_mask, ctx->stream);
    SA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (out_count)
        *out_count = count;
    return st;
}

int sa_unknown_numbering(sa_ctx* ctx, const uint8_t* mask, int64_t rows, int64_t cols, int64_t row_stride,
    int64_t col_stride, int32_t* numbering, int64_t* out_count)
{
    SA_TRY(check_ctx(ctx));
    if (rows < 0 || cols < 0 || (rows * cols > 0 && (!mask || !numbering)))
        return fail(ctx, SA_BAD_ARGUMENT, "unknown_numbering: bad arguments");
    if (rows * cols > (int64_t)INT32_MAX)  // the reference stores the count in a 32-bit int (poisson.cpp:177)
        return fail(ctx, SA_BAD_ARGUMENT, "unknown_numbering: rows * cols must fit 32 bits");
    uint8_t* d_mask = nullptr;
    SA_TRY(upload_mask_row_major(ctx, mask, rows, cols, row_stride, col_stride, &d_mask));
    int32_t* d_num = nullptr;
    int64_t count = 0;
    int st = SA_OK;
    if (rows * cols > 0) {
        SA_CUDA(ctx, cudaMallocAsync(&d_num, (size_t)(rows * cols) * sizeof(int32_t), ctx->stream));
        st = device_numbering(ctx, d_mask, rows, cols, cols, d_num, nullptr, 0, &count, nullptr);
        if (st == SA_OK) {
            SA_CUDA(ctx, cudaMemcpyAsync(numbering, d_num, (size_t)(rows * cols) * sizeof(int32_t), cudaMemcpyDeviceToHost,
                             ctx->stream));
        }
        cudaFreeAsync(d_num, ctx->stream);
        cudaFreeAsync(d_mask, ctx->stream);
    }
    SA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (out_count)
        *out_count = count;
    return st;
}

int sa_label_components(sa_ctx* ctx, const uint8_t* mask, int64_t rows, int64_t cols, int64_t row_stride,
    int64_t col_stride, int32_t* labels, int32_t* out_num_labels)
{
    SA_TRY(check_ctx(ctx));
    if (rows < 0 || cols < 0 || (rows * cols > 0 && (!mask || !labels)))
        return fail(ctx, SA_BAD_ARGUMENT, "label_components: bad arguments");
    if (out_num_labels)
        *out_num_labels = 0;
    if (rows * cols == 0)
        return SA_OK;
    if (rows * cols > (int64_t)INT32_MAX)
        return fail(ctx, SA_BAD_ARGUMENT, "label_components: rows * cols must fit a 32-bit label table");
    uint8_t* d_mask = nullptr;
    SA_TRY(upload_mask_row_major(ctx, mask, rows, cols, row_stride, col_stride, &d_mask));
    int32_t* d_lab = nullptr;
    SA_CUDA(ctx, cudaMallocAsync(&d_lab, (size_t)(rows * cols) * sizeof(int32_t), ctx->stream));
    int32_t K = 0;
    int st = device_label_components(ctx, d_mask, rows, cols, cols, d_lab, &K);
    if (st == SA_OK) {
        SA_CUDA(ctx, cudaMemcpyAsync(labels, d_lab, (size_t)(rows * cols) * sizeof(int32_t), cudaMemcpyDeviceToHost,
                         ctx->stream));
    }
    cudaFreeAsync(d_lab, ctx->stream);
    cudaFreeAsync(d_mask, ctx->stream);
    SA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (out_num_labels)
        *out_num_labels = K;
    return st;
}

/* ---- scenes --------------------------------------------------------------------------------------------------- */

int sa_scene_create(sa_ctx* ctx, int problem, int64_t rows, int64_t cols, int nbands, sa_scene** out)
{
    SA_TRY(check_ctx(ctx));
    if (!out || rows < 0 || cols < 0 || nbands < 1 || (problem != SA_LAPLACE && problem != SA_POISSON))
        return fail(ctx, SA_BAD_ARGUMENT, "scene_create: bad arguments");
    // The strip kernels address a band plane with 32-bit element offsets (cg_strip.cu: TileBits::origin): the padded plane
    // -- (rows rounded up to 32 + 2 guard rows) x (cols + 1 rounded up to 32), in either orientation -- must hold fewer
    // than 2^31 elements (a 46000 x 46000 band; bigger holes are split by rows across GPUs, dist.cu).
    if (sa_scene_plane_elements(rows, cols) > (int64_t)INT32_MAX || rows > (1 << 20) || cols > (1 << 20))
        return fail(ctx, SA_BAD_ARGUMENT, "scene_create: scene too large for one device plane (32-bit plane offsets)");
    // per-band scalars are published through a fixed page-locked buffer (cg.cu: k_publish_scalars, two look-ahead slots)
    if (nbands > SA_MAX_BANDS)
        return fail(ctx, SA_BAD_ARGUMENT, "scene_create: too many bands in one scene (SA_MAX_BANDS)");
    sa_scene* s = new (std::nothrow) sa_scene();
    if (!s)
        return fail(ctx, SA_OUT_OF_MEMORY, "scene_create: host allocation failed");
    s->ctx = ctx;
    s->problem = problem;
    s->user_rows = rows;
    s->user_cols = cols;
    s->nbands = nbands;
    *out = s;
    return SA_OK;
}

void sa_scene_destroy(sa_scene* scene)
{
    if (scene && scene->ctx)
        cudaStreamSynchronize(scene->ctx->stream);
    scene_free(scene);
}

int sa_scene_set_mask(sa_scene* s, const uint8_t* src, int64_t row_stride, int64_t col_stride, int src_on_device)
{
    if (!s)
        return SA_BAD_ARGUMENT;
    SA_TRY(check_ctx(s->ctx));
    SA_TRY(scene_orient(s, row_stride, col_stride));
    SA_TRY(copy_in<uint8_t>(s, s->mask0(s->mask), src, row_stride, col_stride, src_on_device, 0, s->rows));
    s->mask_set = true;
    s->indexed = false;
    return SA_OK;
}

int sa_scene_set_band(sa_scene* s, int band, const double* src, int64_t row_stride, int64_t col_stride,
    int src_on_device)
{
    if (!s)
        return SA_BAD_ARGUMENT;
    SA_TRY(check_ctx(s->ctx));
    if (band < 0 || band >= s->nbands)
        return fail(s->ctx, SA_BAD_ARGUMENT, "scene_set_band: band out of range");
    SA_TRY(scene_orient(s, row_stride, col_stride));
    return copy_in<double>(s, s->plane0(s->u, band), src, row_stride, col_stride, src_on_device, 0, s->rows);
}

int sa_scene_set_guidance(sa_scene* s, int band, const double* src, int64_t row_stride, int64_t col_stride,
    int src_on_device)
{
    if (!s)
        return SA_BAD_ARGUMENT;
    SA_TRY(check_ctx(s->ctx));
    if (s->problem != SA_POISSON)
        return fail(s->ctx, SA_BAD_ARGUMENT, "scene_set_guidance: not a Poisson scene");
    if (band < 0 || band >= s->nbands)
        return fail(s->ctx, SA_BAD_ARGUMENT, "scene_set_guidance: band out of range");
    SA_TRY(scene_orient(s, row_stride, col_stride));
    return copy_in<double>(s, s->plane0(s->g, band), src, row_stride, col_stride, src_on_device, 0, s->rows);
}

int sa_scene_solve(sa_scene* s, const sa_options* opts, sa_stats* stats)
{
    if (!s)
        return SA_BAD_ARGUMENT;
    SA_TRY(check_ctx(s->ctx));
    if (!s->mask_set)
        return fail(s->ctx, SA_BAD_ARGUMENT, "scene_solve: no mask set");
    sa_options o;
    if (opts)
        o = *opts;
    else
        sa_default_options(&o, s->problem);
    if (!(o.tolerance > 0.0))
        o.tolerance = s->problem == SA_POISSON ? 1e-6 : DBL_EPSILON;
    return solve_scene(s, o, stats);
}

int sa_scene_get_band(sa_scene* s, int band, double* dst, int64_t row_stride, int64_t col_stride, int dst_on_device)
{
    if (!s)
        return SA_BAD_ARGUMENT;
    SA_TRY(check_ctx(s->ctx));
    if (!s->oriented || band < 0 || band >= s->nbands)
        return fail(s->ctx, SA_BAD_ARGUMENT, "scene_get_band: band out of range or empty scene");
    SA_TRY(scene_orient(s, row_stride, col_stride));
    SA_TRY(copy_out<double>(s, s->plane0(s->u, band), dst, row_stride, col_stride, dst_on_device, 0, s->rows));
    if (!dst_on_device)
        SA_CUDA(s->ctx, cudaStreamSynchronize(s->ctx->stream));
    return SA_OK;
}

int sa_scene_precondition(sa_scene* s, const sa_options* opts, const double* r, double* z, int64_t row_stride,
    int64_t col_stride)
{
    if (!s)
        return SA_BAD_ARGUMENT;
    SA_TRY(check_ctx(s->ctx));
    if (!s->mask_set || !r || !z)
        return fail(s->ctx, SA_BAD_ARGUMENT, "scene_precondition: no mask set or null buffers");
    SA_TRY(scene_orient(s, row_stride, col_stride));
    sa_options o;
    if (opts)
        o = *opts;
    else
        sa_default_options(&o, s->problem);
    o.precond = SA_PRECOND_MULTIGRID;
    SA_TRY(ensure_indexed(s));
    SA_TRY(copy_in<double>(s, s->plane0(s->r, 0), r, row_stride, col_stride, 0, 0, s->rows));
    SA_TRY(precondition_scene(s, o));
    SA_TRY(copy_out<double>(s, s->plane0(s->p[0], 0), z, row_stride, col_stride, 0, 0, s->rows));
    SA_CUDA(s->ctx, cudaStreamSynchronize(s->ctx->stream));
    // the work vectors must be zero outside the unknown set and consistent for the next solve: re-clear them
    s->indexed = false;
    return SA_OK;
}

int sa_scene_info(const sa_scene* s, int64_t* unknowns, int32_t* active_tiles, int32_t* total_tiles)
{
    if (!s)
        return SA_BAD_ARGUMENT;
    if (unknowns)
        *unknowns = s->n_unknowns;
    if (active_tiles)
        *active_tiles = s->n_active_tiles;
    if (total_tiles)
        *total_tiles = s->tiles_x * s->tiles_y;
    return SA_OK;
}

/* ---- row decomposition across GPUs ------------------------------------------------------------------------------------ */

int sa_dist_unique_id(void* id128)
{
    if (!id128)
        return SA_BAD_ARGUMENT;
    return dist_unique_id(id128);
}

int sa_dist_init(sa_ctx* ctx, const void* id128, int rank, int world)
{
    SA_TRY(check_ctx(ctx));
    if (!id128)
        return fail(ctx, SA_BAD_ARGUMENT, "dist_init: null id");
    if (ctx->comm)
        return fail(ctx, SA_BAD_ARGUMENT, "dist_init: the context already has a communicator");
    return dist_init(ctx, id128, rank, world);
}

int sa_dist_partition(int64_t rows, int world, int levels, int64_t* row_begin)
{
    if (rows < 0 || world < 1 || levels < 1 || levels > 8 || !row_begin)
        return SA_BAD_ARGUMENT;
    dist_partition(rows, world, levels, row_begin);
    return SA_OK;
}

int sa_dist_levels(int64_t rows, int world)
{
    if (rows < 0 || world < 1)
        return 0;
    return dist_choose_levels(rows, world);
}

int sa_scene_set_distributed(sa_scene* s, int on)
{
    if (!s)
        return SA_BAD_ARGUMENT;
    if (on && !s->ctx->comm)
        return fail(s->ctx, SA_BAD_ARGUMENT, "scene_set_distributed: the context has no communicator (sa_dist_init)");
    if (s->distributed != (on != 0)) {
        // a split scene only ever scrubs its own rows: switching modes clears the work vectors wholesale and re-indexes
        s->work_dirty |= 8;  // WORK_FULL
        s->indexed = false;
    }
    s->distributed = on != 0;
    s->dist_planned = false;
    s->dist_no_window = false;
    return SA_OK;
}

int sa_scene_owned_rows(const sa_scene* s, int64_t* lo, int64_t* hi, int* axis)
{
    if (!s)
        return SA_BAD_ARGUMENT;
    int64_t a = 0, b = s->rows;
    if (s->distributed && s->dist_planned && !s->dl.empty()) {
        a = s->dl[0].row_lo < s->rows ? s->dl[0].row_lo : s->rows;
        b = s->dl[0].row_hi < s->rows ? s->dl[0].row_hi : s->rows;
    }
    if (lo)
        *lo = a;
    if (hi)
        *hi = b;
    if (axis)
        *axis = s->transposed ? 1 : 0;
    return SA_OK;
}

int sa_dist_uses_peer_memory(const sa_ctx* ctx) { return ctx ? dist_uses_peer_memory(ctx) : 0; }

int sa_scene_allgather_band(sa_scene* s, int band)
{
    if (!s)
        return SA_BAD_ARGUMENT;
    SA_TRY(check_ctx(s->ctx));
    if (band < 0 || band >= s->nbands)
        return fail(s->ctx, SA_BAD_ARGUMENT, "scene_allgather_band: band out of range");
    return dist_allgather_band(s, band);
}

/* ---- float path, host pointers ---------------------------------------------------------------------------------- */

static int host_fill(sa_ctx* ctx, int problem, double* const* images, const double* const* guidance, int nbands,
    const uint8_t* mask, int64_t rows, int64_t cols, int64_t rs, int64_t cs, const sa_options* opts, sa_stats* stats)
{
    SA_TRY(check_ctx(ctx));
    if (nbands < 1 || rows < 0 || cols < 0 || !images || (problem == SA_POISSON && !guidance))
        return fail(ctx, SA_BAD_ARGUMENT, "fill: bad arguments");
    if (rows * cols == 0) {
        if (stats)
            for (int b = 0; b < nbands; ++b) {
                stats[b] = sa_stats {};
                stats[b].status = SA_EMPTY_MASK;
            }
        return SA_EMPTY_MASK;
    }
    if (!mask)
        return fail(ctx, SA_BAD_ARGUMENT, "fill: null mask");
    Layout lay = classify(rows, cols, rs, cs);
    if (lay == LAYOUT_BAD)
        return fail(ctx, SA_BAD_ARGUMENT, "strides: one of row_stride / col_stride must be 1");
    sa_ctx_cache* cache = cache_of(ctx);
    sa_scene* s = cache->scene;
    bool transposed = lay == LAYOUT_COL_MAJOR;
    if (s && (s->problem != problem || s->user_rows != rows || s->user_cols != cols || s->nbands != nbands
                 || s->transposed != transposed)) {
        cudaStreamSynchronize(ctx->stream);
        scene_free(s);
        s = cache->scene = nullptr;
    }
    if (!s) {
        SA_TRY(sa_scene_create(ctx, problem, rows, cols, nbands, &s));
        int st = scene_alloc(s, transposed);
        if (st != SA_OK) {
            scene_free(s);
            return st;
        }
        cache->scene = s;
    }
    // 1. mask -> unknown set and active tiles; only the band of rows that holds active tiles (plus one known row on
    //    either side for the boundary values) has to cross PCIe.
    SA_TRY(copy_in<uint8_t>(s, s->mask0(s->mask), mask, rs, cs, 0, 0, s->rows));
    s->mask_set = true;
    s->indexed = false;
    sa_options o;
    if (opts)
        o = *opts;
    else
        sa_default_options(&o, problem);
    if (!(o.tolerance > 0.0))
        o.tolerance = problem == SA_POISSON ? 1e-6 : DBL_EPSILON;
    SA_TRY(ensure_indexed(s, o.precond != SA_PRECOND_MULTIGRID ? WORK_JACOBI
                                 : (o.mg_variant != SA_MG_JACOBI64 && o.cg_variant == 0 ? (WORK_RB | (o.mg_variant == SA_MG_RB32 ? WORK_RBW : 0)) : WORK_J64)));
    if (s->n_unknowns == 0)
        return solve_scene(s, o, stats);  // fills stats, returns SA_EMPTY_MASK
    // 2a. Direct mode: when the caller's arrays are page-locked (cudaMallocHost / cudaHostRegister / torch pin_memory) the
    //     device can address them, and no image is copied at all: the set-up kernel reads, straight from host memory,
    //     only the pixels the equations look at -- the ring of known pixels around the unknown set, plus g on the unknown
    //     set for Poisson -- and a scatter kernel stores only the unknown pixels back.  Known pixels never cross PCIe.
    {
        bool direct = std::getenv("SATFILL_NO_DIRECT") == nullptr && o.cg_variant == 0
            && (o.precond != SA_PRECOND_MULTIGRID || o.mg_variant != SA_MG_JACOBI64) && s->cols % 2 == 0 && s->rows > 0;
        const int64_t sp = slow_stride(s, rs, cs);
        // the fetch / scatter kernels address the caller's array with 32-bit element offsets as well
        direct = direct && sp % 2 == 0 && (s->rows_p + 2) * sp <= (int64_t)INT32_MAX;
        std::vector<double*> dev_f((size_t)nbands, nullptr);
        std::vector<const double*> dev_g((size_t)nbands, nullptr);
        for (int b = 0; b < nbands && direct; ++b) {
            cudaPointerAttributes at {};
            if (cudaPointerGetAttributes(&at, images[b]) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer
                || ((uintptr_t)at.devicePointer & 15)) {
                direct = false;
                break;
            }
            dev_f[(size_t)b] = (double*)at.devicePointer;
            if (problem == SA_POISSON) {
                cudaPointerAttributes ag {};
                if (cudaPointerGetAttributes(&ag, guidance[b]) != cudaSuccess || ag.type != cudaMemoryTypeHost || !ag.devicePointer
                    || ((uintptr_t)ag.devicePointer & 15)) {
                    direct = false;
                    break;
                }
                dev_g[(size_t)b] = (const double*)ag.devicePointer;
            }
        }
        cudaGetLastError();  // cudaPointerGetAttributes on pageable memory may leave a sticky-free error behind
        ctx->last_fill_direct = direct;
        if (direct) {
            // Band windows: the unknowns of window c travel home (k_scatter_direct on io_out, a few CTAs) while window
            // c + 1 is solved.  A solve costs a few milliseconds whatever its size and a window's scatter has to hide
            // under the next solve, so a tile is cut into about five windows -- one when its bands are small.
            const int64_t band_bytes = s->rows * s->cols * (int64_t)sizeof(double);
            int64_t chunk_bytes = (int64_t)256 << 20;
            if (const char* e = std::getenv("SATFILL_CHUNK_BYTES"))
                chunk_bytes = std::max<int64_t>(1, std::atoll(e));
            int per_chunk = (int)std::min<int64_t>(std::min(nbands, HOST_BANDS_MAX), std::max<int64_t>(1, chunk_bytes / std::max<int64_t>(band_bytes, 1)));
            if (per_chunk >= nbands && nbands <= HOST_BANDS_MAX)
                per_chunk = nbands;
            else if (!std::getenv("SATFILL_CHUNK_BYTES"))
                per_chunk = std::min(HOST_BANDS_MAX, std::max(per_chunk, (nbands + 4) / 5));
            if (const char* e = std::getenv("SATFILL_CHUNK_BANDS"))  // tuning knob
                per_chunk = std::min(std::min(nbands, HOST_BANDS_MAX), std::max(1, std::atoi(e)));
            std::vector<int> edge { 0 };  // window c holds the bands [edge[c], edge[c + 1])
            while (edge.back() < nbands)
                edge.push_back(std::min(nbands, edge.back() + per_chunk));
            const int nch = (int)edge.size() - 1;
            if (!ctx->io_out) {
                SA_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->io_in, cudaStreamNonBlocking));
                SA_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->io_out, cudaStreamNonBlocking));
            }
            const Level lv = fine_level(s);
            auto window = [&](int b0, int b1) {
                HostBands hb {};
                for (int b = b0; b < b1; ++b) {
                    hb.f[b - b0] = dev_f[(size_t)b];
                    hb.g[b - b0] = dev_g[(size_t)b];
                }
                hb.pitch = sp;
                hb.rows = s->rows;
                hb.cols = s->cols;
                return hb;
            };
            int st = SA_OK;
            const bool dbg = std::getenv("SATFILL_DEBUG_IO") != nullptr;
            auto now_ms = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
            const double t_start = now_ms();
            SA_TRY(prepare_solve(s, o));
            // The way in: while window c is solved on the context's stream, k_fetch_direct pulls the ring of known pixels
            // of window c + 1 over PCIe on io_in.  Every read of host memory is a PCIe round trip and the bus takes only
            // so many at a time (measured: 5 - 7 GB/s with ~6000 threads, less with fewer AND with more), so the rings of
            // a 13-band tile need 30 - 45 ms however they are fetched: they have to travel beside the solves, not in
            // front of them.
            while ((int)ctx->io_ev.size() < nch) {
                cudaEvent_t e;
                SA_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                ctx->io_ev.push_back(e);
            }
            auto fetch = [&](int c) {
                const int b0 = edge[(size_t)c], b1 = edge[(size_t)c + 1];
                SA_TRY(launch_fetch_direct(ctx, ctx->io_in, lv, b1 - b0, problem == SA_POISSON, s->plane0(s->u, b0),
                    problem == SA_POISSON ? s->plane0(s->g, b0) : nullptr, window(b0, b1)));
                SA_CUDA(ctx, cudaEventRecord(ctx->io_ev[(size_t)c], ctx->io_in));
                return (int)SA_OK;
            };
            // The io kernels hold a few SMs for themselves (cg_strip.cu: io_ctas); the solve kernels size their grids --
            // CTAs that each walk a fixed share of the tiles -- for the SMs that are left, so that every solve CTA is
            // resident at once.
            struct GridGuard {
                sa_ctx* c;
                ~GridGuard() { c->grid_sms = c->sm_count; }
            } grid_guard { ctx };
            if (nch > 1) {
                int reserve = io_ctas(false) + io_ctas(true);
                if (const char* e = std::getenv("SATFILL_IO_RESERVE_SMS"))  // tuning knob
                    reserve = std::atoi(e);
                ctx->grid_sms = std::max(ctx->sm_count / 2, ctx->sm_count - std::max(0, reserve));
            }
            SA_TRY(fetch(0));
            for (int c = 0; c < nch; ++c) {
                const int b0 = edge[(size_t)c], b1 = edge[(size_t)c + 1];
                const HostBands hb = window(b0, b1);
                s->band0 = b0;
                s->band_n = b1 - b0;
                SA_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->io_ev[(size_t)c], 0));  // this window's ring has landed
                if (c + 1 < nch)
                    SA_TRY(fetch(c + 1));  // the next one trickles in while this window is solved
                int stc = solve_scene(s, o, stats ? stats + b0 : nullptr);  // returns with the context's stream drained
                s->band0 = 0;
                s->band_n = -1;
                if (stc != SA_OK && stc != SA_NOT_CONVERGED) {
                    cudaStreamSynchronize(ctx->io_in);
                    cudaStreamSynchronize(ctx->io_out);
                    return stc;
                }
                if (stc != SA_OK)
                    st = stc;
                if (dbg)
                    std::fprintf(stderr, "[satfill io] direct chunk %d: solved at %.1f ms (set-up %.1f ms, solve %.1f ms)\n", c,
                        now_ms() - t_start, stats ? stats[b0].setup_ms : 0.0, stats ? stats[b0].solve_ms : 0.0);
                if (problem == SA_LAPLACE)  // never looks at the solver status (laplace.cpp:113-119)
                    SA_TRY(launch_scatter_direct(ctx, ctx->io_out, lv, b1 - b0, s->plane0(s->u, b0), hb));
            }
            if (problem == SA_POISSON && st == SA_OK)  // nothing is written unless every band converged (poisson.cpp:263-269)
                for (int c = 0; c < nch; ++c) {
                    const int b0 = edge[(size_t)c], b1 = edge[(size_t)c + 1];
                    SA_TRY(launch_scatter_direct(ctx, ctx->io_out, lv, b1 - b0, s->plane0(s->u, b0), window(b0, b1)));
                }
            SA_CUDA(ctx, cudaStreamSynchronize(ctx->io_out));
            if (dbg)
                std::fprintf(stderr, "[satfill io] direct: all out at %.1f ms\n", now_ms() - t_start);
            return st;
        }
    }
    int32_t* h_cnt = (int32_t*)ctx->pinned;
    SA_CUDA(ctx, cudaMemcpyAsync(h_cnt, s->d_counters, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int64_t row_lo = (int64_t)(h_cnt[1] / s->tiles_x) * TILE_H - 1;
    int64_t row_hi = (int64_t)(h_cnt[2] / s->tiles_x + 1) * TILE_H + 1;
    if (row_lo < 0)
        row_lo = 0;
    if (row_hi > s->rows)
        row_hi = s->rows;
    // 2b. Pageable (or oddly laid out) host arrays: bands cross PCIe whole and are solved in chunks: chunk c + 1 .. are on their way in (io_in) and chunk c - 1 is on
    //    its way out (io_out) while chunk c is solved on the context's stream.  A chunk is at least ~256 MB of image so
    //    that its transfer hides the solve's fixed costs; small scenes go through as one chunk.
    const int64_t band_bytes = (row_hi - row_lo) * s->cols * (int64_t)sizeof(double);
    int64_t chunk_bytes = (int64_t)256 << 20;
    if (const char* e = std::getenv("SATFILL_CHUNK_BYTES"))  // tests force the chunked path on small scenes
        chunk_bytes = std::max<int64_t>(1, std::atoll(e));
    int per_chunk = (int)std::min<int64_t>(nbands, std::max<int64_t>(1, chunk_bytes / std::max<int64_t>(band_bytes, 1)));
    const bool windows_ok = o.precond != SA_PRECOND_MULTIGRID || o.mg_variant != SA_MG_JACOBI64;
    if (per_chunk >= nbands || !windows_ok || std::getenv("SATFILL_NO_PIPELINE"))
        per_chunk = nbands;
    const int nch = (nbands + per_chunk - 1) / per_chunk;
    if (nch > 1) {
        if (!ctx->io_in) {
            SA_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->io_in, cudaStreamNonBlocking));
            SA_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->io_out, cudaStreamNonBlocking));
        }
        while ((int)ctx->io_ev.size() < nch) {
            cudaEvent_t e;
            SA_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->io_ev.push_back(e);
        }
    }
    cudaStream_t sin = nch > 1 ? ctx->io_in : ctx->stream, sout = nch > 1 ? ctx->io_out : ctx->stream;
    for (int c = 0; c < nch; ++c) {
        for (int b = c * per_chunk; b < std::min(nbands, (c + 1) * per_chunk); ++b) {
            SA_TRY(copy_in<double>(s, s->plane0(s->u, b), images[b], rs, cs, 0, row_lo, row_hi, sin));
            if (problem == SA_POISSON)
                SA_TRY(copy_in<double>(s, s->plane0(s->g, b), guidance[b], rs, cs, 0, row_lo, row_hi, sin));
        }
        if (nch > 1)
            SA_CUDA(ctx, cudaEventRecord(ctx->io_ev[(size_t)c], sin));
    }
    // Laplace never looks at the solver status (laplace.cpp:113-119): a chunk leaves as soon as it is solved.  Poisson
    // writes nothing unless every band converged (poisson.cpp:263-269): its bands leave after the last chunk.
    int st = SA_OK;
    const bool dbg = std::getenv("SATFILL_DEBUG_IO") != nullptr;
    auto now_ms = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t_start = now_ms();
    for (int c = 0; c < nch; ++c) {
        const int b0 = c * per_chunk, b1 = std::min(nbands, (c + 1) * per_chunk);
        if (dbg) {
            cudaEventSynchronize(ctx->io_ev[(size_t)c]);
            std::fprintf(stderr, "[satfill io] chunk %d: in at %.1f ms", c, now_ms() - t_start);
        }
        if (nch > 1)
            SA_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->io_ev[(size_t)c], 0));
        s->band0 = b0;
        s->band_n = b1 - b0;
        int stc = solve_scene(s, o, stats ? stats + b0 : nullptr);  // returns with the context's stream drained
        s->band0 = 0;
        s->band_n = -1;
        if (dbg)
            std::fprintf(stderr, ", solved at %.1f ms\n", now_ms() - t_start);
        if (stc != SA_OK && stc != SA_NOT_CONVERGED) {
            cudaStreamSynchronize(sin);
            cudaStreamSynchronize(sout);
            return stc;
        }
        if (stc != SA_OK)
            st = stc;
        if (problem == SA_LAPLACE)
            for (int b = b0; b < b1; ++b)
                SA_TRY(copy_out<double>(s, s->plane0(s->u, b), images[b], rs, cs, 0, row_lo, row_hi, sout));
    }
    if (problem == SA_POISSON && st == SA_OK)
        for (int b = 0; b < nbands; ++b)
            SA_TRY(copy_out<double>(s, s->plane0(s->u, b), images[b], rs, cs, 0, row_lo, row_hi, sout));
    SA_CUDA(ctx, cudaStreamSynchronize(sout));
    if (nch > 1)
        SA_CUDA(ctx, cudaStreamSynchronize(sin));
    if (dbg)
        std::fprintf(stderr, "[satfill io] all out at %.1f ms\n", now_ms() - t_start);
    return st;
}

int sa_laplace_fill(sa_ctx* ctx, double* const* images, int nbands, const uint8_t* mask, int64_t rows, int64_t cols,
    int64_t row_stride, int64_t col_stride, const sa_options* opts, sa_stats* stats)
{
    return host_fill(ctx, SA_LAPLACE, images, nullptr, nbands, mask, rows, cols, row_stride, col_stride, opts, stats);
}

int sa_poisson_blend(sa_ctx* ctx, double* const* inputs, const double* const* replacements, int nbands,
    const uint8_t* mask, int64_t rows, int64_t cols, int64_t row_stride, int64_t col_stride, const sa_options* opts,
    sa_stats* stats)
{
    return host_fill(ctx, SA_POISSON, inputs, replacements, nbands, mask, rows, cols, row_stride, col_stride, opts, stats);
}

int sa_last_fill_direct(const sa_ctx* ctx) { return ctx && ctx->last_fill_direct ? 1 : 0; }

/* ---- the steps either side of the path ------------------------------------------------------------------------------ */

int sa_apply_laplace_u8(sa_ctx* ctx, const uint8_t* image, const uint8_t* invalid, int64_t rows, int64_t cols, int channels,
    double red_threshold, double* out, uint8_t* mask_out, const sa_options* opts, sa_stats* stats)
{
    SA_TRY(check_ctx(ctx));
    if (!image || !invalid || !out || rows < 0 || cols < 0)
        return fail(ctx, SA_BAD_ARGUMENT, "apply_laplace: bad arguments");
    if (channels != 3)
        return fail(ctx, SA_BAD_ARGUMENT, "apply_laplace: cv::imread(IMREAD_COLOR) images have 3 channels");
    if (rows * cols == 0)
        return SA_EMPTY_MASK;
    // the cached scene of the host-pointer entry points, row-major like cv::Mat
    sa_ctx_cache* cache = cache_of(ctx);
    sa_scene* s = cache->scene;
    if (s && (s->problem != SA_LAPLACE || s->user_rows != rows || s->user_cols != cols || s->nbands != channels || s->transposed)) {
        cudaStreamSynchronize(ctx->stream);
        scene_free(s);
        s = cache->scene = nullptr;
    }
    if (!s) {
        SA_TRY(sa_scene_create(ctx, SA_LAPLACE, rows, cols, channels, &s));
        int st = scene_alloc(s, false);
        if (st != SA_OK) {
            scene_free(s);
            return st;
        }
        cache->scene = s;
    }
    const size_t nbytes = (size_t)rows * cols * channels;
    uint8_t* d_u8 = nullptr;
    double* d_out = nullptr;
    SA_CUDA(ctx, cudaMalloc(&d_u8, 2 * nbytes));
    if (cudaMalloc(&d_out, nbytes * sizeof(double)) != cudaSuccess) {
        cudaFree(d_u8);
        return fail(ctx, SA_OUT_OF_MEMORY, "apply_laplace: output staging");
    }
    auto cleanup = [&](int st) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(d_u8);
        cudaFree(d_out);
        return st;
    };
    if (cudaMemcpyAsync(d_u8, image, nbytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess
        || cudaMemcpyAsync(d_u8 + nbytes, invalid, nbytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
        return cleanup(fail(ctx, SA_CUDA_ERROR, "apply_laplace: upload"));
    int st = split_u8_scene(s, d_u8, d_u8 + nbytes, channels, red_threshold);
    if (st != SA_OK)
        return cleanup(st);
    s->mask_set = true;
    s->indexed = false;
    sa_options o;
    if (opts)
        o = *opts;
    else
        sa_default_options(&o, SA_LAPLACE);
    if (!(o.tolerance > 0.0))
        o.tolerance = DBL_EPSILON;
    st = solve_scene(s, o, stats);
    if (st != SA_OK && st != SA_NOT_CONVERGED && st != SA_EMPTY_MASK)  // Laplace never looks at the solver status
        return cleanup(st);
    int st2 = merge_f64_scene(s, channels, d_out);
    if (st2 != SA_OK)
        return cleanup(st2);
    if (cudaMemcpyAsync(out, d_out, nbytes * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
        return cleanup(fail(ctx, SA_CUDA_ERROR, "apply_laplace: download"));
    if (mask_out
        && cudaMemcpy2DAsync(mask_out, (size_t)cols, s->mask0(s->mask), (size_t)s->pitch, (size_t)cols, (size_t)rows,
               cudaMemcpyDeviceToHost, ctx->stream)
            != cudaSuccess)
        return cleanup(fail(ctx, SA_CUDA_ERROR, "apply_laplace: mask download"));
    return cleanup(st);
}

int sa_morph_close_mask(sa_ctx* ctx, const double* band, int64_t rows, int64_t cols, int64_t row_stride, int64_t col_stride,
    int radius, uint8_t* mask_out)
{
    SA_TRY(check_ctx(ctx));
    if (!band || !mask_out || rows < 0 || cols < 0 || radius < 0)
        return fail(ctx, SA_BAD_ARGUMENT, "morph_close_mask: bad arguments");
    if (rows * cols == 0)
        return SA_OK;
    Layout lay = classify(rows, cols, row_stride, col_stride);
    if (lay == LAYOUT_BAD)
        return fail(ctx, SA_BAD_ARGUMENT, "strides: one of row_stride / col_stride must be 1");
    // a rectangle is transposition-symmetric: work on the buffer as it lies (slow x fast)
    const bool col_major = lay == LAYOUT_COL_MAJOR;
    const int64_t slow = col_major ? cols : rows, fast = col_major ? rows : cols;
    const int64_t sp = col_major ? col_stride : row_stride;  // host elements between slow-axis neighbours
    const size_t n = (size_t)slow * fast;
    double *d_a = nullptr, *d_b = nullptr;
    uint8_t* d_m = nullptr;
    SA_CUDA(ctx, cudaMalloc(&d_a, n * sizeof(double)));
    if (cudaMalloc(&d_b, n * sizeof(double)) != cudaSuccess || cudaMalloc(&d_m, n) != cudaSuccess) {
        cudaFree(d_a);
        cudaFree(d_b);
        return fail(ctx, SA_OUT_OF_MEMORY, "morph_close_mask: staging");
    }
    auto cleanup = [&](int st) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(d_a);
        cudaFree(d_b);
        cudaFree(d_m);
        return st;
    };
    if (cudaMemcpy2DAsync(d_a, (size_t)fast * sizeof(double), band, (size_t)(slow > 1 ? sp : fast) * sizeof(double),
            (size_t)fast * sizeof(double), (size_t)slow, cudaMemcpyHostToDevice, ctx->stream)
        != cudaSuccess)
        return cleanup(fail(ctx, SA_CUDA_ERROR, "morph_close_mask: upload"));
    int st = morph_close_mask(ctx, d_a, d_b, slow, fast, radius, d_m);
    if (st != SA_OK)
        return cleanup(st);
    if (cudaMemcpy2DAsync(mask_out, (size_t)(slow > 1 ? sp : fast), d_m, (size_t)fast, (size_t)fast, (size_t)slow,
            cudaMemcpyDeviceToHost, ctx->stream)
        != cudaSuccess)
        return cleanup(fail(ctx, SA_CUDA_ERROR, "morph_close_mask: download"));
    return cleanup(SA_OK);
}

}  // extern "C"
