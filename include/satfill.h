/* satfill.h -- C-ABI of libsatfill.so, the B200 (sm_100a) implementation of the Laplace / Poisson fill path of
 * ebiederstadt/satellite-approximation (lib/approx).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ / torch / Eigen types.  The C++ `approx`
 * shim (cpp/include/approx/ headers), the pybind11 module `satellite_approximation._core` and the ctypes binding
 * (satellite_approximation_b200/_capi.py) all sit on exactly these entry points.  Each entry point names the
 * reference interface it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   * Images are `double`, masks are one byte per pixel (non-zero = invalid), like the reference's
 *     utils::MatX<f64> / MatX<bool> (lib/utils/include/utils/types.h:31).
 *   * Every host image / mask argument carries explicit ELEMENT strides (row_stride, col_stride).  The reference's
 *     MatX is column-major (row_stride = 1, col_stride = rows); numpy C-order is (cols, 1).  One of the two strides
 *     must be 1 and the other >= the extent it jumps over.
 *   * Integer outputs are always in row-major raster order of (row, col), which is how the reference scans
 *     (laplace.cpp:34-40, poisson.cpp:169-176).
 *   * Host-pointer entry points own all device memory and copies; `sa_scene_*` entry points keep a scene resident
 *     in HBM so that a caller (bench.py) can time the solve alone.
 *   * Every function returns an sa_status; sa_last_error(ctx) holds a message for the last non-zero one.
 *   * A context is the unit of thread-safety: one host thread per context at a time.
 */
#ifndef SATFILL_H
#define SATFILL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SATFILL_ABI_VERSION 5

typedef enum sa_status {
    SA_OK = 0,
    SA_EMPTY_MASK = 1,     /* no invalid pixel: nothing done (laplace.cpp:41-44 logs and returns)                 */
    SA_NOT_CONVERGED = 2,  /* max_iterations reached first (Eigen::NoConvergence, poisson.cpp:263-269)            */
    SA_SIZE_MISMATCH = 3,  /* laplace.cpp:124-127 throws; poisson.cpp:154-157 logs and returns                    */
    SA_BAD_ARGUMENT = 4,
    SA_CUDA_ERROR = 5,
    SA_NCCL_ERROR = 6,
    SA_OUT_OF_MEMORY = 7
} sa_status;

typedef struct sa_ctx sa_ctx;
typedef struct sa_scene sa_scene;

typedef enum sa_problem {
    SA_LAPLACE = 0, /* laplace.cpp:31-120: unknowns = invalid pixels not on the image border, diagonal 4, x0 = 0   */
    SA_POISSON = 1  /* poisson.cpp:145-290: unknowns = invalid pixels, diagonal = #in-image neighbours, x0 = g     */
} sa_problem;

typedef enum sa_precond {
    SA_PRECOND_JACOBI = 0,   /* Eigen's DiagonalPreconditioner (BasicPreconditioners.h:39-86), matrix-free        */
    SA_PRECOND_MULTIGRID = 1 /* symmetric V-cycle on the masked grid (DESIGN.md "Multigrid")                       */
} sa_precond;

typedef enum sa_mg_variant {
    SA_MG_RB32 = 0,     /* red-black Gauss-Seidel V(1,1), float arithmetic inside the preconditioner: one warp per tile,
                         * neighbourhood in registers, coarse tail in one cooperative launch (mg_rbw.cu)              */
    SA_MG_JACOBI64 = 1, /* damped-Jacobi V(nu,nu), double (mg_fused.cu / mg.cu)                                      */
    SA_MG_RB32_CTA = 2  /* the same cycle as SA_MG_RB32 on its first-generation kernels (one CTA per tile and band,
                         * neighbourhood in shared memory, one launch per level: mg_rb.cu); kept as the tested reference */
} sa_mg_variant;

/* Solver knobs.  The reference exposes tolerance / max_iterations on Poisson only (poisson.h:45-46); Laplace runs
 * Eigen defaults (epsilon, 2N; laplace.cpp:113-114, IterativeSolverBase.h:251,367-368).  Zero-initialise, then
 * call sa_default_options(). */
typedef struct sa_options {
    double tolerance;       /* stop when |b_U - A_UU x|_2 <= tolerance * |b_U|_2 on the reduced system          */
    int64_t max_iterations; /* <= 0: reference default (Laplace 2n, Poisson n/2 with n = unknowns)              */
    int32_t precond;        /* sa_precond                                                                        */
    int32_t check_every;    /* host polls the device-side convergence flags every this many iterations; <= 0:
                             * automatic (32 for Jacobi, 1 for multigrid)                                         */
    int32_t mg_levels;      /* multigrid: maximum number of levels (<= 0: automatic)                             */
    int32_t mg_smooth;      /* multigrid: pre = post smoothing sweeps                                            */
    int32_t profile;        /* != 0: time every solver kernel with CUDA events (sa_stats.kernel_ms)              */
    int32_t mg_unfused;     /* != 0: run the V-cycle one sweep per kernel (reference path of the fused kernels)  */
    int32_t mg_variant;     /* sa_mg_variant; CG itself (iterate, residual, operator, dot products) is always double */
    int32_t cg_variant;     /* 0: strip kernels (cg_strip.cu); 1: first-generation staged-tile kernels (cg.cu)          */
} sa_options;

/* Per-band solve record (superset of approx::PerfInfo, poisson.h:12-21). */
typedef struct sa_stats {
    int64_t unknowns;       /* PerfInfo::region_size                                                            */
    int64_t iterations;     /* PerfInfo::iterations                                                             */
    int64_t max_iterations; /* PerfInfo::max_iterations actually applied                                        */
    double tolerance;       /* PerfInfo::tolerance                                                              */
    double error;           /* PerfInfo::error: sqrt(|r|^2 / |b|^2) at exit                                     */
    double solve_ms;        /* PerfInfo::solve_time: device time of the solve loop for the whole batch          */
    double setup_ms;        /* device time of mask indexing + right-hand side                                   */
    int32_t status;         /* sa_status of this band                                                           */
    int32_t active_tiles;   /* tiles of the block-sparse layout that contain an unknown                         */
    /* sa_options.profile: summed CUDA-event durations and launch counts of the solver kernels of the whole batch,
     * by class: 0 = CG direction (p update + p.Ap), 1 = CG update (x, r, norms), 2 = multigrid single sweeps (the
     * default red-black path: its cooperative tail kernel), 3 = multigrid single transfers (residual, restriction,
     * prolongation; the default red-black path has none and counts here the passes of the CG update that leave x alone --
     * it adds alpha p to x every other pass, class 1 then holds the passes that add two steps), 4 = fused multigrid descent
     * (pre-smoothing + residual + restriction) on level 0, 5 = fused multigrid ascent (prolongation + post-smoothing)
     * on level 0, 6 / 7 = the same two on the coarse levels */
    double kernel_ms[8];
    int64_t kernel_launches[8];
    /* unknowns x bands summed over the launches of each class (a multigrid launch on level l counts the unknowns of
     * level l): algorithmic bytes of a class = bytes per unknown x kernel_units */
    int64_t kernel_units[8];
} sa_stats;

/* ---- context ------------------------------------------------------------------------------------------------ */

/* device: CUDA ordinal.  stream: a cudaStream_t the library should launch on, or NULL for a (non-blocking) stream the
 * context creates and owns.  The legacy default stream has handle 0 = NULL: a caller that produces device inputs on
 * the default stream must synchronise before handing them over, or pass an explicit stream. */
int sa_create(sa_ctx** out, int device, void* stream);
void sa_destroy(sa_ctx* ctx);
const char* sa_last_error(const sa_ctx* ctx);
int sa_abi_version(void);
/* 1 when the library also holds the first-generation kernels (built with SATFILL_LEGACY_VARIANTS: cg_variant = 1,
 * SA_MG_JACOBI64, SA_MG_RB32_CTA -- the tested references of the product kernels); the product library returns 0 and
 * refuses those variants with SA_BAD_ARGUMENT. */
int sa_has_legacy_variants(void);
void sa_default_options(sa_options* opts, int problem);
/* number of kernels of this library launched through ctx since creation (bench.py's gpu_launches) */
int64_t sa_kernel_launches(const sa_ctx* ctx);

/* ---- integer path, host pointers ----------------------------------------------------------------------------- */

/* Replaces the invalid-pixel scan + bounding box of solve_matrix (laplace.cpp:33-52).
 * out_pixels: capacity `capacity` (row, col) int64 pairs, may be NULL to only count.  bbox = {min_row, max_row,
 * min_col, max_col}; for an empty mask {rows, -1, cols, -1}. */
int sa_mask_scan(sa_ctx* ctx, const uint8_t* mask, int64_t rows, int64_t cols, int64_t row_stride, int64_t col_stride,
    int64_t* out_pixels, int64_t capacity, int64_t* out_count, int64_t bbox[4]);

/* Replaces the unknown numbering of blend_images_poisson (poisson.cpp:162-177): numbering[col + row*cols] = k, the
 * number of invalid pixels strictly before (row, col) in raster order, or -1 for valid pixels. */
int sa_unknown_numbering(sa_ctx* ctx, const uint8_t* mask, int64_t rows, int64_t cols, int64_t row_stride,
    int64_t col_stride, int32_t* numbering, int64_t* out_count);

/* Replaces approx::find_connected_components (laplace.h:11-20, declared but never defined in the reference;
 * contract: tests/approximation.h:55-75 + SURVEY.md 8a A3).  4-connectivity, background 0, labels 1..K by first
 * pixel in raster order.  labels is a dense row-major rows x cols table. */
int sa_label_components(sa_ctx* ctx, const uint8_t* mask, int64_t rows, int64_t cols, int64_t row_stride,
    int64_t col_stride, int32_t* labels, int32_t* out_num_labels);

/* ---- float path, host pointers -------------------------------------------------------------------------------- */

/* Replaces approx::fill_missing_portion_smooth_boundary (laplace.h:28, laplace.cpp:122-132) for `nbands` images
 * sharing one mask (apply_laplace, laplace.cpp:152-162, calls it once per channel and re-assembles each time).
 * images[b] is modified in place at invalid pixels only.  The element-count check of laplace.cpp:124-127 is made by
 * the caller-side shim (one rows x cols here covers image and mask); stats: nbands entries or NULL.
 * Laplace unknowns are the invalid pixels that are not on the image border: border pixels are Dirichlet data and come
 * back unchanged (laplace.cpp:98-100). */
int sa_laplace_fill(sa_ctx* ctx, double* const* images, int nbands, const uint8_t* mask, int64_t rows, int64_t cols,
    int64_t row_stride, int64_t col_stride, const sa_options* opts, sa_stats* stats);

/* Replaces approx::blend_images_poisson, mask overload (poisson.h:41-46, poisson.cpp:145-290).  inputs[b] is
 * modified in place at invalid pixels only, and only if every band converged (poisson.cpp:263-269). */
int sa_poisson_blend(sa_ctx* ctx, double* const* inputs, const double* const* replacements, int nbands,
    const uint8_t* mask, int64_t rows, int64_t cols, int64_t row_stride, int64_t col_stride, const sa_options* opts,
    sa_stats* stats);

/* ---- device-resident scenes ------------------------------------------------------------------------------------ */

/* A scene = one mask + nbands images (+ nbands guidance images for SA_POISSON) + solver work vectors, all in HBM,
 * in the library's layout: row-major with the unit-stride axis of the source as the fast axis (the 5-point operator
 * is transpose-invariant, so a column-major source is solved as its transpose without a transposition). */
int sa_scene_create(sa_ctx* ctx, int problem, int64_t rows, int64_t cols, int nbands, sa_scene** out);
void sa_scene_destroy(sa_scene* scene);
/* Limits of sa_scene_create (and of the host-pointer fills, which keep a scene).  The kernels address a band plane
 * with 32-bit element offsets: sa_scene_plane_elements(rows, cols) -- the padded plane, in the larger of the two
 * orientations a scene can be resident in; pure host arithmetic, needs no device -- must not exceed INT32_MAX
 * (a 46000 x 46000 band; larger systems are split by rows across GPUs).  At most SA_MAX_BANDS bands per scene. */
#define SA_MAX_BANDS 512
int64_t sa_scene_plane_elements(int64_t rows, int64_t cols);
/* src_on_device != 0: `src` is a device pointer (same strides convention).  Uploads are asynchronous on the
 * context's stream when the host memory is pinned. */
int sa_scene_set_mask(sa_scene* scene, const uint8_t* src, int64_t row_stride, int64_t col_stride, int src_on_device);
int sa_scene_set_band(sa_scene* scene, int band, const double* src, int64_t row_stride, int64_t col_stride,
    int src_on_device);
int sa_scene_set_guidance(sa_scene* scene, int band, const double* src, int64_t row_stride, int64_t col_stride,
    int src_on_device);
/* Mask indexing + tile list + right-hand side + solve + write-back into the resident bands. */
int sa_scene_solve(sa_scene* scene, const sa_options* opts, sa_stats* stats);
int sa_scene_get_band(sa_scene* scene, int band, double* dst, int64_t row_stride, int64_t col_stride, int dst_on_device);
/* After a solve (or any call that indexed the mask): unknowns of the linear system, tiles of the block-sparse layout
 * that hold one, and tiles in total.  Any pointer may be NULL. */
int sa_scene_info(const sa_scene* scene, int64_t* unknowns, int32_t* active_tiles, int32_t* total_tiles);
/* Diagnostic hook for the parity tests: z = M^-1 r, one application of the multigrid preconditioner selected by `opts`
 * to band 0 of a scene whose mask is set.  r and z are host rows x cols tables (same strides); r is taken at the
 * unknown cells only, z is zero elsewhere.  Not used by the fill path itself. */
int sa_scene_precondition(sa_scene* scene, const sa_options* opts, const double* r, double* z, int64_t row_stride,
    int64_t col_stride);
/* Blocks until everything queued on the context's stream has finished. */
int sa_synchronize(sa_ctx* ctx);

/* ---- one system across several GPUs (a single very large hole, decomposed by rows) -------------------- */

/* One process per GPU.  The host language moves 128 bytes: rank 0 calls sa_dist_unique_id, broadcasts the id by whatever
 * it has (torch.distributed, MPI, a file), then every rank calls sa_dist_init on its context.  NCCL (libnccl.so.2) is
 * resolved at run time; without it these calls return SA_NCCL_ERROR and everything else keeps working. */
int sa_dist_unique_id(void* id128);
int sa_dist_init(sa_ctx* ctx, const void* id128, int rank, int world);
/* The row partition the solver uses (pure host logic): world + 1 row boundaries, aligned to 32 * 2^(levels - 1) rows so
 * that `levels` multigrid levels split at the same places.  sa_dist_levels: the number of levels it splits for a scene
 * of `rows` rows; coarser levels are replicated on every rank. */
int sa_dist_partition(int64_t rows, int world, int levels, int64_t* row_begin);
int sa_dist_levels(int64_t rows, int world);
/* Mark a scene of a context that went through sa_dist_init as ONE system shared by all ranks: every rank sets the same
 * mask and bands (only its own rows and one row around them are read), sa_scene_solve is collective, and afterwards
 * each rank holds the solution on the rows sa_scene_owned_rows reports (axis: 0 = rows, 1 = columns of the caller's
 * array -- the split runs along the slow axis of the memory layout).  sa_scene_allgather_band (collective) completes
 * a band on every rank. */
int sa_scene_set_distributed(sa_scene* scene, int on);
int sa_scene_owned_rows(const sa_scene* scene, int64_t* lo, int64_t* hi, int* axis);
int sa_scene_allgather_band(sa_scene* scene, int band);
/* 1 when the exchanges inside the iteration loop of a row-decomposed solve go over peer memory -- every rank's arena mapped by
 * all ranks of the node through CUDA IPC, two small kernels per exchange -- and 0 when they go over NCCL (the fallback: peers
 * that cannot map each other, or SATFILL_DIST_NCCL_ONLY in the environment).  Valid after the first distributed solve. */
int sa_dist_uses_peer_memory(const sa_ctx* ctx);

/* 1 if the last sa_laplace_fill / sa_poisson_blend of this context ran in direct mode: the caller's arrays were page-locked
 * (device-addressable), so no image was copied -- the set-up kernel read the known pixels that border the unknown set (and
 * the replacement image on the unknown set) straight from host memory and only the unknown pixels were stored back.
 * 0 if the images were copied whole (pageable memory, an odd fast extent, SATFILL_NO_DIRECT). */
int sa_last_fill_direct(const sa_ctx* ctx);

/* ---- the steps either side of the path (next rows) ------------------------------------------------------- */

/* approx::apply_laplace (lib/approx/source/laplace.cpp:134-168), the body of laplace_main
 * (executables/laplace-main.cpp:34-40).  `image` and `invalid` are interleaved 8-bit images as cv::imread(IMREAD_COLOR)
 * returns them: rows x cols x channels bytes, channel order B, G, R, channels == 3.  Invalid pixels are those of
 * `invalid` with R >= red_threshold and G <= 150 (laplace.cpp:141-146); every channel of `image` is filled over that
 * mask in one batched solve (the reference re-assembles and solves per channel, laplace.cpp:152-162).  `out` receives
 * rows x cols x channels doubles, interleaved like the CV_64FC3 matrix the reference returns (laplace.cpp:159-167);
 * `mask_out` (optional) rows x cols bytes.  stats: `channels` entries (optional).  Border semantics as sa_laplace_fill. */
int sa_apply_laplace_u8(sa_ctx* ctx, const uint8_t* image, const uint8_t* invalid, int64_t rows, int64_t cols, int channels,
    double red_threshold, double* out, uint8_t* mask_out, const sa_options* opts, sa_stats* stats);

/* preprocess_cloud_band (executables/poisson-main.cpp:10-21): cv::morphologyEx(MORPH_CLOSE) of a float64 band with a
 * (2 radius + 1)^2 rectangle (poisson_main: radius 5), then MatX<f64>::cast<bool>().  Dilation then erosion, windows
 * clipped at the image border (OpenCV's default border value for morphology leaves outside pixels out of the max / min).
 * mask_out has the band's layout (same element strides), one byte per pixel: 1 where the closed band is non-zero. */
int sa_morph_close_mask(sa_ctx* ctx, const double* band, int64_t rows, int64_t cols, int64_t row_stride, int64_t col_stride,
    int radius, uint8_t* mask_out);

#ifdef __cplusplus
}
#endif
#endif /* SATFILL_H */
