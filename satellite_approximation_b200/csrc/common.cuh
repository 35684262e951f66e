// Shared declarations of libsatfill.so (sm_100a only).  See include/satfill.h for the C-ABI and DESIGN.md for the
// layout.  Every image-shaped quantity of a scene is a row-major "plane": (rows_p + 2) x pitch elements per band,
// rows_p = rows rounded up to the 32-row tile, pitch = (cols + 1) rounded up to the 32-column tile, so that there is
// at least one zero column after the last image column, plus one zero guard row above and below.  Pointers address
// element (0, 0); the guards sit at negative / past-the-end offsets.  With the solver's convention that every work
// vector is ZERO at cells that are not unknowns, the 5-point operator needs no bounds checks and no neighbour-mask
// look-ups: a neighbour that is known, outside the image or in the padding simply contributes 0.
// Only the tiles listed in `tile_list` (tiles holding at least one unknown) are ever touched by solver kernels.
#pragma once

#include <cuda_runtime.h>

#include <cfloat>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/satfill.h"

// SATFILL_LEGACY_VARIANTS = 1 also builds the first-generation kernels -- cg_variant = 1 (cg.cu: tile + halo staged through
// shared memory), SA_MG_JACOBI64 (mg.cu, mg_fused.cu) and SA_MG_RB32_CTA (mg_rb.cu) -- the references the product kernels
// are tested against (lib/libsatfill_legacy.so, tests/test_gpu_legacy.py).  The product library ships the hot path only.
#ifndef SATFILL_LEGACY_VARIANTS
#define SATFILL_LEGACY_VARIANTS 0
#endif

namespace satfill {

constexpr int TILE_W = 32;
constexpr int TILE_H = 32;
constexpr int CG_BLOCK_X = 32;
#ifndef SATFILL_BLOCK_Y
#define SATFILL_BLOCK_Y 4
#endif
// Warps per tile CTA.  4 (8 rows per thread) rather than 8: the tile kernels are latency-bound chains (tile list ->
// mask -> data), so what matters is bytes in flight per SM = resident CTAs x loads per thread; fewer, fatter threads
// double both at the same thread count (profiles/: 2.4-3 TB/s with 8 warps).
constexpr int CG_BLOCK_Y = SATFILL_BLOCK_Y;
constexpr int CG_THREADS = CG_BLOCK_X * CG_BLOCK_Y;
constexpr int ROWS_PER_THREAD = TILE_H / CG_BLOCK_Y;
constexpr int MAX_LEVELS = 12;

// Per-band CG scalars, device resident.  Slot s = iteration & 3 (a ring of four so that the slot two iterations
// ahead can be zeroed while the current and previous ones are still being read; DESIGN.md "Reductions").
struct BandScalars {
    double rz[4];   // r.z entering iteration k
    double rr[4];   // |r|^2 entering iteration k
    double pq[4];   // p.Ap of iteration k
    double bnorm2;  // |b_U|^2
    double thr;     // max(tol^2 |b|^2, DBL_MIN)      (ConjugateGradient.h:50-51)
    double rr_exit; // |r|^2 when the band stopped
    int done;       // sticky: band has met thr (or had a zero right-hand side)
    int iters;      // Eigen-style iteration count at exit (ConjugateGradient.h:65-82)
    int zero_rhs;   // ConjugateGradient.h:43-49: x is set to zero
    // x += alpha p applied every other pass (k_update2, XM): the band's last pass left alpha_hist[pend_buf ^ 1 ...] -- see
    // k_flush_x.  pend != 0: the step of the band's last pass (alpha_hist[pend_pass & 1], direction in p buffer pend_buf)
    // has not been added to x yet.
    int pend;
    double alpha_hist[2];  // alpha of pass k in slot k & 1
    int pend_buf;
    int pad;
};

// One grid level of a scene (level 0 = the image; levels >= 1 = multigrid coarse grids).  Passed to kernels by value.
struct Level {
    int64_t rows, cols;   // logical extent
    int64_t pitch;        // elements per plane row
    int64_t plane;        // elements per band plane including the two guard rows
    int tiles_x, tiles_y;
    int n_tiles;          // active tiles
    const uint8_t* umask; // 1 = unknown of the linear system; addressed like a plane (pitch bytes per row)
    const int32_t* tile_list;
    const int32_t* tile_yx;  // the same list as packed tile coordinates (ty << 16 | tx): no division in the kernels
    int fixed_diag;       // != 0: the diagonal is 4 everywhere (Laplace: unknowns never touch the image border)
    // The same unknown set as one 32-bit word per tile row: word [((ty + 1) * tb_stride + tx + 1) * 32 + row], bit = col.
    // A ring of all-zero tiles surrounds the grid, so neighbourhoods of any tile can be read without bounds checks.
    // 128 B per tile: small enough to stay resident in L2 across kernels (15 MB for a 10980^2 scene).
    const uint32_t* tbits;
    const uint32_t* tbitsT;  // transposed: word index = column of the tile, bit = row (column masks for the fused kernels)
    int tb_stride;        // tiles_x + 2
    // coarse levels of the red-black cycle: 1 / diagonal of the boundary-corrected coarse operator (mg.cu), a float plane
    // shared by all bands; nullptr on level 0 (diagonal = neighbour count)
    const float* winv;
};

// Direct mode of the host-pointer entry points: the caller's page-locked arrays as the device addresses them (api.cu).
constexpr int HOST_BANDS_MAX = 16;
struct HostBands {
    double* f[HOST_BANDS_MAX];        // images: read by k_fetch_direct, written (unknown pixels only) by k_scatter_direct
    const double* g[HOST_BANDS_MAX];  // Poisson: replacement images
    int64_t pitch;                    // elements between rows of the caller's arrays (the slow-axis stride)
    int64_t rows, cols;               // resident extents
};

}  // namespace satfill

struct sa_ctx {
    // row decomposition across GPUs (dist.cu): NCCL communicator (void*: nccl.h stays out of this header), rank, size
    void* comm = nullptr;
    int rank = 0, world = 1;
    double* d_red = nullptr;  // packed per-band scalars for the all-reduce
    void* peer = nullptr;  // peer-memory arena of the row decomposition (dist.cu: PeerState)
    unsigned* d_barrier = nullptr;  // arrival counter of the grid-wide barrier of the cooperative tail kernel (mg_rbw.cu)
    // host-pointer entry points: PCIe transfers of the other band chunks run on these while a chunk is solved
    cudaStream_t io_in = nullptr, io_out = nullptr;
    bool last_fill_direct = false;
    std::vector<cudaEvent_t> io_ev;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    int64_t launches = 0;
    std::string error;
    int sm_count = 148;
    int grid_sms = 148;  // the SMs the solve kernels size their grids for (api.cu: fewer while io kernels run beside them)
    // pinned host scratch for polling convergence flags / small read-backs
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    cudaEvent_t ev[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    std::vector<cudaEvent_t> ev_pool;  // per-kernel timing (sa_options.profile)
};

struct sa_level_store {
    satfill::Level lv {};
    uint8_t* umask_alloc = nullptr;    // base of the allocation (guard row included)
    int32_t* tile_list = nullptr;      // 3 * tiles entries: list, per-tile flags, packed (ty, tx) list
    int32_t* d_counters = nullptr;     // {active tiles, first, last, -}
    uint32_t* tbits = nullptr;  // row words, then (at +tb_words) the transposed column words
    size_t tb_words = 0;
    int64_t rows_p = 0;
    double* x = nullptr;               // coarse levels: correction; nbands planes (allocation base)
    double* b = nullptr;               // coarse levels: restricted residual
    double* t = nullptr;               // scratch (second smoothing buffer)
    float* winv = nullptr;             // Level::winv (allocation base, guard row included)
    int64_t n_unknowns = 0;
};

// One distributed multigrid level as a rank sees it: the rows it owns and its slice of the raster-ordered tile list.
struct DistLevel {
    int64_t row_lo = 0, row_hi = 0, rows = 0;
    int tile_lo = 0, tile_hi = 0;
    std::vector<int64_t> bounds;  // row boundaries of all ranks at this level (world + 1)
};

struct sa_scene {
    // row decomposition (dist.cu)
    bool distributed = false, dist_planned = false, dist_mg = false;
    // The mask is indexed -- unknown set, tile lists, bit words, coarse levels that are split by rows -- only on the rank's own
    // tile rows (+ one tile row either side), so that the set-up of a row-decomposed solve scales with the ranks: see
    // dist_prepare_window (dist.cu).  dist_no_window: the scene turned out too small for that (fewer usable levels than planned).
    bool dist_windowed = false, dist_no_window = false, dist_mg_window = false;
    double dist_unit_frac = 1.0;  // share of the rows this rank owns: the profile counts the unknowns a rank processes
    int dist_levels = 0;                      // multigrid levels that are split by rows; coarser ones are replicated
    std::vector<DistLevel> dl;
    std::vector<int64_t> dist_gather_rows;    // rows of the first replicated level each rank produces (world + 1)
    sa_ctx* ctx = nullptr;
    int problem = SA_LAPLACE;
    int64_t user_rows = 0, user_cols = 0;  // as given to sa_scene_create
    bool transposed = false;               // the resident layout is the transpose of the user's (row, col) grid
    bool oriented = false;
    int64_t rows = 0, cols = 0;            // resident extents (fast axis = cols)
    int nbands = 0;
    int64_t pitch = 0;
    int64_t rows_p = 0;
    int64_t plane = 0;                     // (rows_p + 2) * pitch
    int tiles_x = 0, tiles_y = 0;
    // allocation bases (nbands planes each); the *0 accessors below skip the first guard row
    double* u = nullptr;   // image; unknown pixels hold the iterate x
    double* g = nullptr;   // guidance (SA_POISSON only)
    double* r = nullptr;
    double* p[2] = { nullptr, nullptr };
    double* z = nullptr;   // multigrid only: preconditioned residual
    double* t = nullptr;   // multigrid only: fine-level scratch (smoother ping-pong, residual)
    size_t p_bytes = 0;    // bytes of each of p[0], p[1]: float planes for the product path, double planes once a path needs them (cg.cu: ensure_p)
    uint8_t* mask = nullptr;   // normalised 0/1 invalid mask (allocation base)
    uint8_t* umask = nullptr;  // unknown set (allocation base)
    int32_t* tile_list = nullptr;
    uint32_t* tbits = nullptr;  // tile-row bit masks of umask (Level::tbits), then the transposed words (Level::tbitsT)
    size_t tb_words = 0;
    int32_t* d_counters = nullptr;
    unsigned long long* d_count64 = nullptr;
    satfill::BandScalars* scal = nullptr;
    int n_active_tiles = 0;
    int64_t n_unknowns = 0;
    bool mask_set = false;
    bool indexed = false;
    std::vector<sa_level_store> coarse;  // multigrid hierarchy below level 0
    bool hierarchy_built = false;
    // What has written the work vectors (r, p, z, t, coarse x / b / t) since they were last zero everywhere: a set of
    // WORK_* bits.  A solve leaves them non-zero only at the unknowns of ITS mask, so when the mask changes they are
    // scrubbed through the OLD tile lists (scrub_work_vectors) instead of being cleared plane by plane.
    int work_dirty = 8;        // WORK_FULL until the first clear
    bool ever_indexed = false;  // the tile lists / bit masks of a previous mask are valid
    bool stale_r = false;       // r was left unscrubbed by a mask change (the red-black path never reads r unmasked)
    bool stale_rb = false;      // ... and so were the cycle's masked-only vectors (float copy of r, red halves, coarse b)
    bool stale_all = false;     // ... and so was everything else (strip CG + warp-per-tile cycle before and after the change)

    // Band window of the next sa_scene_solve-like call: the host-pointer entry points (api.cu) solve a scene in chunks of
    // bands so that PCIe transfers of the other chunks overlap the solve.  band_n < 0: all bands.
    int band0 = 0, band_n = -1;
    int win_n() const { return band_n < 0 ? nbands : band_n; }

    double* plane0(double* base, int band) const { return base + (int64_t)band * plane + pitch; }
    // float planes of the red-black cycle (mg_rb.cu) inside the z allocation: z itself, then the float copy of the
    // CG residual that k_update / k_residual write for it
    // (both start at the first band of the window)
    float* rb_z() const { return (float*)z + pitch + (int64_t)band0 * plane; }
    float* rb_rf() const { return (float*)z + (int64_t)plane * nbands + pitch + (int64_t)band0 * plane; }
    uint8_t* mask0(uint8_t* base) const { return base + pitch; }
};

namespace satfill {

inline int fail(sa_ctx* ctx, int status, const std::string& msg)
{
    if (ctx)
        ctx->error = msg;
    return status;
}

#define SA_CUDA(ctx, expr)                                                                                       \
    do {                                                                                                         \
        cudaError_t e__ = (expr);                                                                                \
        if (e__ != cudaSuccess) {                                                                                \
            return satfill::fail((ctx), e__ == cudaErrorMemoryAllocation ? SA_OUT_OF_MEMORY : SA_CUDA_ERROR,     \
                std::string(#expr) + ": " + cudaGetErrorString(e__));                                            \
        }                                                                                                        \
    } while (0)

#define SA_TRY(expr)             \
    do {                         \
        int st__ = (expr);       \
        if (st__ != SA_OK)       \
            return st__;         \
    } while (0)

// Every kernel of the library is launched through this macro so that sa_kernel_launches() is exact.
#define SA_LAUNCH(ctx, kernel, grid, block, smem, ...)                     \
    do {                                                                   \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);   \
        (ctx)->launches += 1;                                              \
    } while (0)
// a launch of a first-generation kernel: compiled away in the product build (the request is refused before it gets there)
#if SATFILL_LEGACY_VARIANTS
#define SA_LAUNCH_LEGACY(...) SA_LAUNCH(__VA_ARGS__)
#else
#define SA_LAUNCH_LEGACY(...) \
    do {                      \
    } while (0)
#endif

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// Kernel classes reported in sa_stats.kernel_ms / kernel_launches when sa_options.profile is set.
enum KernelClass {
    KC_DIRECTION = 0, KC_UPDATE = 1, KC_SMOOTH = 2, KC_TRANSFER = 3, KC_MG_DOWN = 4, KC_MG_UP = 5, KC_MG_DOWN_COARSE = 6,
    KC_MG_UP_COARSE = 7, KC_COUNT = 8,
    KC_UPDATE_DEFERRED = KC_TRANSFER  // strip CG + red-black cycle: the passes of k_update2 that leave x alone (the slot is free there)
};

// Brackets individual launches with CUDA events on the launching stream; the pairs are resolved after the next
// stream synchronisation (flush).  Off unless profiling was requested: the events cost ~1 us of launch gap each.
struct KernelTimer {
    sa_ctx* ctx = nullptr;
    bool on = false;
    size_t used = 0;
    struct Pending { int cls; size_t e0; };
    std::vector<Pending> pending;
    double ms[KC_COUNT] = {};
    int64_t n[KC_COUNT] = {};
    int64_t units[KC_COUNT] = {};

    cudaEvent_t take()
    {
        if (used == ctx->ev_pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ctx->ev_pool.push_back(e);
        }
        return ctx->ev_pool[used++];
    }
    void begin(int cls, int64_t launch_units = 0)
    {
        if (!on)
            return;
        units[cls] += launch_units;
        pending.push_back({ cls, used });
        cudaEventRecord(take(), ctx->stream);
    }
    void end()
    {
        if (!on)
            return;
        cudaEventRecord(take(), ctx->stream);
    }
    void flush()  // call after a stream synchronisation
    {
        for (const Pending& p : pending) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, ctx->ev_pool[p.e0], ctx->ev_pool[p.e0 + 1]) == cudaSuccess) {
                ms[p.cls] += t;
                n[p.cls] += 1;
            }
        }
        pending.clear();
        used = 0;
    }
};

// ---- mask_index.cu ---------------------------------------------------------------------------------------------
int transpose_u8(sa_ctx* ctx, const uint8_t* src, int64_t src_rows, int64_t src_cols, int64_t src_pitch, uint8_t* dst,
    int64_t dst_pitch);
int transpose_i32(sa_ctx* ctx, const int32_t* src, int64_t src_rows, int64_t src_cols, int64_t src_pitch, int32_t* dst,
    int64_t dst_pitch);
// normalise to 0/1, build umask + active tile list + unknown count for a scene
int index_scene(sa_scene* s);
// number of unknowns of the whole mask (row-decomposed scenes index only their own rows)
int count_unknowns(sa_scene* s, int64_t* out);
// raster-order numbering / pixel list / bbox of a device mask (pitch bytes per row, non-zero = invalid)
int device_numbering(sa_ctx* ctx, const uint8_t* mask, int64_t rows, int64_t cols, int64_t pitch, int32_t* numbering,
    int64_t* out_pixels, int64_t capacity, int64_t* out_count, int64_t bbox[4]);
// flags[n_tiles] -> raster-ordered list of the flagged tiles; d_n_active (device, 4 ints) receives
// {count, first tile, last tile}
int compact_tile_flags(sa_ctx* ctx, const int32_t* flags, int n_tiles, int tiles_x, int32_t* tile_list, int32_t* tile_yx,
    int32_t* d_n_active, int first_tile = 0);
// exclusive scan helper shared with ccl.cu: in place over `n` 64-bit counters, total written to *total
int device_scan_u64(sa_ctx* ctx, unsigned long long* data, int64_t n, unsigned long long* total);

// ---- ccl.cu ----------------------------------------------------------------------------------------------------
int device_label_components(sa_ctx* ctx, const uint8_t* mask, int64_t rows, int64_t cols, int64_t pitch,
    int32_t* labels /* dense rows x cols, device */, int32_t* out_num_labels);

// ---- cg.cu -----------------------------------------------------------------------------------------------------
// index the mask if it changed and clear the work vectors (they must be zero outside the unknown set)
int ensure_indexed(sa_scene* s, int next_kind = 0);  // next_kind: WORK_* of the solve that follows (0: not known)
int solve_scene(sa_scene* s, const sa_options& o, sa_stats* stats);
Level fine_level(const sa_scene* s);
// z = M^-1 r on band 0: r is taken from s->r (masked here), the result is left as doubles in s->p[0]
int precondition_scene(sa_scene* s, const sa_options& o);

// ---- cg_strip.cu: the two kernels of a CG iteration, shared-memory-free generation ---------------------------------
int launch_setup2(sa_ctx* ctx, const Level& lv, int nbands, bool poisson, double* u, const double* g, double* r, float* rf,
    BandScalars* scal);
int prepare_solve(sa_scene* s, const sa_options& o);
int launch_fetch_direct(sa_ctx* ctx, cudaStream_t stream, const Level& lv, int nbands, bool poisson, double* u, double* g,
    const HostBands& src);
int launch_scatter_direct(sa_ctx* ctx, cudaStream_t stream, const Level& lv, int nbands, const double* u, const HostBands& dst);
int ensure_p(sa_scene* s, size_t elem_bytes);  // cg.cu: the two search-direction buffers, sized for float or double planes
int io_ctas(bool scatter);  // CTAs (one SM each) of the fetch / scatter kernel
int launch_direction2(sa_ctx* ctx, const Level& lv, int nbands, bool jacobi, const void* zin, bool z_is_float,
    const void* p_old, void* p_new, bool p_is_float, BandScalars* scal, int k);
// xm: 0 = x += alpha p;  1 = x is left alone (alpha and the direction stay behind for the next pass);  2 = x += the step of
// the previous pass (direction p_prev) and this one
int launch_update2(sa_ctx* ctx, const Level& lv, int nbands, bool jacobi, double* u, const void* p, bool p_is_float, double* r,
    float* rf, BandScalars* scal, int k, int xm = 0, const void* p_prev = nullptr);
int launch_flush_x(sa_ctx* ctx, const Level& lv, int nbands, double* u, const float* p0, const float* p1, const BandScalars* scal);

// work_dirty bits
enum { WORK_CLEAN = 0, WORK_JACOBI = 1, WORK_RB = 2, WORK_J64 = 4, WORK_FULL = 8, WORK_PF = 16 /* p planes hold floats */,
    WORK_RBW = 32 /* the red-black cycle ran on the warp-per-tile kernels (mg_rbw.cu) */ };
// cg_strip.cu: zero the given planes at the unknowns of `lv` (whole sectors)
struct ScrubPlanes {
    double* d[5];  // double planes, element (0, 0) of band 0
    float* f[4];   // float planes
    float* h[1];   // colour-split float half planes (mg_rb.cu: red cells)
    int nd, nf, nh;
};
int launch_scrub(sa_ctx* ctx, const Level& lv, int nbands, const ScrubPlanes& planes);

// ---- prepost.cu: the steps either side of the path
int split_u8_scene(sa_scene* s, const uint8_t* d_image, const uint8_t* d_invalid, int channels, double red_threshold);
int merge_f64_scene(sa_scene* s, int channels, double* d_out);
int morph_close_mask(sa_ctx* ctx, double* d_a, double* d_b, int64_t slow, int64_t fast, int radius, uint8_t* d_mask);

// ---- dist.cu: row decomposition of one system across GPUs -------------------------------------------------------------
enum DistWhat { DIST_SETUP = 0, DIST_RZ = 1, DIST_PQ = 2, DIST_RR = 3, DIST_RR_RZ = 4 };
void dist_partition(int64_t rows, int world, int levels, int64_t* row_begin);
int dist_choose_levels(int64_t rows, int world);
int dist_unique_id(void* out128);
int dist_init(sa_ctx* ctx, const void* id128, int rank, int world);
void dist_shutdown(sa_ctx* ctx);
int dist_plan_scene(sa_scene* s, bool multigrid);
int dist_prepare_window(sa_scene* s, bool multigrid);  // before index_scene: the rows of every split level this rank owns
constexpr int SA_RETRY_UNWINDOWED = -77;               // dist_plan_scene: index the whole mask and plan again
Level dist_level(const sa_scene* s, int l, const Level& full);
template <typename T>
int dist_halo(sa_scene* s, int l, T* base, int64_t pitch, int64_t plane, int above, int below);
int dist_gather(sa_scene* s, float* base, int64_t pitch, int64_t plane);
int dist_reduce(sa_scene* s, int what, int slot, int clear_slot);
int dist_reduce_pack(sa_scene* s, int what, int slot);
int dist_reduce_issue(sa_scene* s);
int dist_reduce_unpack(sa_scene* s, int what, int slot, int clear_slot);
int dist_group_begin(sa_scene* s);
int dist_group_end(sa_scene* s);
int dist_allgather_band(sa_scene* s, int band);
// one exchange step (halo rows of a vector and / or the sum of a group of per-band scalars), over peer memory or NCCL
int dist_step(sa_scene* s, int level, int vec, void* base, int elem_bytes, int64_t pitch, int64_t plane, int above, int below, int what,
    int slot, int clear_slot);
int dist_uses_peer_memory(const sa_ctx* ctx);
enum DistVec { DIST_VEC_RHS = 0, DIST_VEC_SOL = 1, DIST_VEC_DIR = 2 };

// ---- mg_fused.cu -------------------------------------------------------------------------------------------------
int launch_mg_down(sa_ctx* ctx, const Level& lf, const Level& lc, int nbands, const double* b, double* x_out, double* bc,
    const BandScalars* scal);
int launch_mg_up(sa_ctx* ctx, const Level& lf, const Level& lc, int nbands, const double* x_in, const double* b,
    const double* ec, double* x_out, BandScalars* scal, int rz_slot);

// ---- mg.cu -----------------------------------------------------------------------------------------------------
int build_hierarchy(sa_scene* s, const sa_options& o);
void free_hierarchy(sa_scene* s);
// z = M^{-1} r for every band that is not done (one symmetric V-cycle); r.z is accumulated into rz[rz_slot]
int apply_vcycle(sa_scene* s, const sa_options& o, KernelTimer& kt, int rz_slot, int live_bands);

// ---- mg_rb.cu ---------------------------------------------------------------------------------------------------
// the same for the red-black float cycle: z is a FLOAT plane in s->z
int apply_vcycle_rb(sa_scene* s, const sa_options& o, KernelTimer& kt, int rz_slot, int live_bands);

// ---- mg_rbw.cu: the same cycle, one warp per tile with the neighbourhood in registers, coarse tail in one launch ------------
int apply_vcycle_rbw(sa_scene* s, const sa_options& o, KernelTimer& kt, int rz_slot, int live_bands);

}  // namespace satfill
