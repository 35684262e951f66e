// Red-black multigrid V(1,1)-cycle in single precision: the default preconditioner of the CG solve (cg.cu).
//
//   z = M^-1 r :  on every level   pre-smoothing  = one red-black Gauss-Seidel sweep (red, then black) from zero,
//                                  post-smoothing = the reverse sweep (black, then red)  =>  M is symmetric positive
//                                  definite, as CG requires (the coarsest level runs K forward then K reverse sweeps).
//
// Same grid hierarchy, transfer operators and re-discretised 5-point operator as mg.cu (DESIGN.md "Multigrid"); only
// the smoother and the arithmetic type differ.  The preconditioner does not have to be accurate -- CG's iterate,
// residual, search direction and operator stay double (cg.cu) -- so it runs in float: half the HBM bytes and half
// the shared-memory wavefronts of the double Jacobi cycle (mg_fused.cu), which ncu showed to be bound by the
// shared-memory data pipe (profiles/r1_fused_*: l1tex data pipe 80 %, DRAM 25 %).  tools/mg_prototype.py: 11 CG
// iterations to 1e-6 against 10 for the double V(2,2) Jacobi cycle, at half the smoothing work.
//
// What red-black ordering buys, per level (red = (row + col) even):
//   * down: red x = b / d is pointwise; after the black half-sweep the residual of every black cell is zero and the
//     residual of a red cell is just the sum of its black neighbours.  The full-weighting restriction lands on red
//     cells and only sees its centre and four diagonal (red) neighbours.
//   * up:   the black half-sweep overwrites black cells without reading them, so only the RED half of the
//     pre-smoothed iterate ever has to travel between the two kernels; it is stored colour-split (half a plane).
//   => per unknown: down  R b 4 (8 on level 0: the CG residual is double) + W x_red 2 + W b_coarse 1
//                   up    R x_red 2 + R b 4 (8) + R e_coarse 1 + W x 4
//
// Kernel structure: one CTA per active 32 x 32 tile and band; the (32 + 2H)^2 neighbourhood is kept colour-split in
// shared memory (R[row][h], B[row][h], h = half column: the red cell of row i sits in column 2h + (i & 1)), so that a
// warp touches consecutive words in every access.  Thread (h, y) slides down rows [4y, 4y + 4) of half column h with
// the vertical neighbours in registers: 2 shared loads + 1 store per cell update.  The unknown set comes from the
// per-tile column bit masks (Level::tbitsT); all global loads are predicated on it and issued before the first use.
#include "common.cuh"
#include "tile.cuh"

namespace satfill {

namespace {

constexpr int RB_HP = 20;  // threads per row group (half columns, padded): lane-linear shared-memory addressing
constexpr int RB_RG = 4;   // rows per thread
constexpr int RB_S = 21;   // row stride of the colour-split arrays: RB_RG * RB_S = 84 = 20 (mod 32) => bank = thread id
constexpr unsigned long long EVEN_ROWS = 0x5555555555555555ull;

template <bool FIXED>
__device__ __forceinline__ float rb_winv(const Level& lv, int64_t r, int64_t c)
{
    if (FIXED)
        return 0.25f;
    int n = (r > 0) + (r < lv.rows - 1) + (c > 0) + (c < lv.cols - 1);
    return n == 4 ? 0.25f : (n == 3 ? (1.0f / 3.0f) : (n == 2 ? 0.5f : 1.0f));
}

// red / black unknown bits of the thread's half column: bit i <=> region row i
template <int H>
__device__ __forceinline__ void colour_masks(const Level& lv, int ty, int tx, int h, unsigned long long& red,
    unsigned long long& black)
{
    constexpr int W = TILE_W + 2 * H;
    unsigned long long cm0 = 2 * h < W ? region_col_mask<H>(lv, ty, tx, 2 * h) : 0ull;
    unsigned long long cm1 = 2 * h + 1 < W ? region_col_mask<H>(lv, ty, tx, 2 * h + 1) : 0ull;
    red = (cm0 & EVEN_ROWS) | (cm1 & ~EVEN_ROWS);
    black = (cm1 & EVEN_ROWS) | (cm0 & ~EVEN_ROWS);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// descent: pre-smoothing from zero, residual, restriction
// ---------------------------------------------------------------------------------------------------------------
template <bool FIXED, typename BT>
__global__ void __launch_bounds__(RB_HP * 10) k_rb_down(Level lf, Level lc, const BT* __restrict__ b,
    float* __restrict__ xr, float* __restrict__ bc, const BandScalars* __restrict__ scal)
{
    constexpr int H = 3, W = TILE_W + 2 * H;  // 38
    constexpr int NG = (W + RB_RG - 1) / RB_RG;  // 10 row groups
    constexpr int THREADS = RB_HP * NG;
    constexpr int ROWS = NG * RB_RG + 2;  // one spare row above and below: the sliding window needs no bounds checks
    __shared__ float R[ROWS * RB_S];
    __shared__ float B[ROWS * RB_S];
    if (scal[blockIdx.y].done)
        return;
    const int t = threadIdx.x, h = t % RB_HP, y = t / RB_HP;
    const int tile = lf.tile_list[blockIdx.x];
    const int ty = tile / lf.tiles_x, tx = tile % lf.tiles_x;
    const int64_t r0 = (int64_t)ty * TILE_H, c0 = (int64_t)tx * TILE_W;
    const int row0 = RB_RG * y;
    unsigned long long redm, blkm;
    colour_masks<H>(lf, ty, tx, h, redm, blkm);
    const unsigned rm = (unsigned)(redm >> row0) & 15u, bm = (unsigned)(blkm >> row0) & 15u;
    const int64_t gr = r0 - H + row0, gc = c0 - H + 2 * h;  // global position of (row0, column 2h)
    // ---- global loads: the right-hand side at the thread's 2 x 4 cells
    float bred[RB_RG], bblk[RB_RG];
    {
        const BT* bp = b + (int64_t)blockIdx.y * lf.plane + gr * lf.pitch + gc;
        BT vr[RB_RG], vb[RB_RG];
#pragma unroll
        for (int k = 0; k < RB_RG; ++k) {
            vr[k] = ((rm >> k) & 1) ? bp[k * lf.pitch + (k & 1)] : BT(0);
            vb[k] = ((bm >> k) & 1) ? bp[k * lf.pitch + 1 - (k & 1)] : BT(0);
        }
#pragma unroll
        for (int k = 0; k < RB_RG; ++k) {
            bred[k] = (float)vr[k];
            bblk[k] = (float)vb[k];
        }
    }
    const int sb = (row0 + 1) * RB_S + h;  // shared index of (row0, h)
    // ---- red half-sweep from zero: x = b / d (pointwise); the tile's own red cells go to HBM colour-split
    {
        float* xo = xr + (int64_t)blockIdx.y * (lf.plane >> 1) + gr * (lf.pitch >> 1);
#pragma unroll
        for (int k = 0; k < RB_RG; ++k) {
            const int row = row0 + k, col = 2 * h + (k & 1);
            float v = rb_winv<FIXED>(lf, gr + k, gc + (k & 1)) * bred[k];
            R[sb + k * RB_S] = v;
            if (((rm >> k) & 1) && row >= H && row < H + TILE_H && col >= H && col < H + TILE_W)
                xo[k * (lf.pitch >> 1) + ((gc + (k & 1)) >> 1)] = v;
        }
    }
    __syncthreads();
    // ---- black half-sweep on rows / columns 1 .. W-2
    {
        const float* p = R + sb;
        float n = p[-RB_S], c = p[0];
#pragma unroll
        for (int k = 0; k < RB_RG; ++k) {
            const int row = row0 + k, col = 2 * h + 1 - (k & 1);
            float s = p[(k + 1) * RB_S];
            float side = p[k * RB_S + ((k & 1) ? -1 : 1)];
            float v = rb_winv<FIXED>(lf, gr + k, gc + 1 - (k & 1)) * (bblk[k] + ((n + s) + (c + side)));
            bool on = ((bm >> k) & 1) && row >= 1 && row < W - 1 && col >= 1 && col < W - 1;
            B[sb + k * RB_S] = on ? v : 0.f;
            n = c;
            c = s;
        }
    }
    __syncthreads();
    // ---- residual: zero at black cells; at a red cell b - d x + sum(black neighbours) = sum(black neighbours)
    {
        const float* p = B + sb;
        float n = p[-RB_S], c = p[0];
#pragma unroll
        for (int k = 0; k < RB_RG; ++k) {
            const int row = row0 + k, col = 2 * h + (k & 1);
            float s = p[(k + 1) * RB_S];
            float side = p[k * RB_S + ((k & 1) ? 1 : -1)];
            float v = (n + s) + (c + side);
            bool on = ((rm >> k) & 1) && row >= 2 && row < W - 2 && col >= 2 && col < W - 2;
            R[sb + k * RB_S] = on ? v : 0.f;  // R is dead as an iterate: reuse it for the residual
            n = c;
            c = s;
        }
    }
    __syncthreads();
    // ---- full-weighting restriction: coarse (ci, cj) <-> tile cell (2ci, 2cj) = region (2ci + 3, 2cj + 3), a red
    //      cell in an odd row (half column cj + 1); its diagonal neighbours are the red cells of the rows above and
    //      below in half columns cj + 1 and cj + 2; its edge neighbours are black (zero residual).
    {
        float* bco = bc + (int64_t)blockIdx.y * lc.plane + (r0 >> 1) * lc.pitch + (c0 >> 1);
        const uint32_t* rowbits = lf.tbits + ((size_t)(ty + 1) * lf.tb_stride + (tx + 1)) * 32;
        for (int i = t; i < (TILE_H / 2) * (TILE_W / 2); i += THREADS) {
            int ci = i >> 4, cj = i & 15;
            if ((rowbits[2 * ci] >> (2 * cj)) & 1) {  // mask injection: coarse unknown <=> fine (2I, 2J) unknown
                const float* p = R + (2 * ci + H + 1) * RB_S + cj + 1;
                bco[ci * lc.pitch + cj] = p[0] + 0.25f * ((p[-RB_S] + p[-RB_S + 1]) + (p[RB_S] + p[RB_S + 1]));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// ascent: prolongation + correction (red cells only), post-smoothing black then red, (level 0) r.z
// ---------------------------------------------------------------------------------------------------------------
template <bool FIXED, typename BT, bool DOT>
__global__ void __launch_bounds__(RB_HP * 9) k_rb_up(Level lf, Level lc, const float* __restrict__ xr,
    const BT* __restrict__ b, const float* __restrict__ ec, float* __restrict__ x_out, BandScalars* __restrict__ scal,
    int slot)
{
    constexpr int H = 2, W = TILE_W + 2 * H;  // 36
    constexpr int NG = W / RB_RG;             // 9 row groups
    constexpr int THREADS = RB_HP * NG;
    constexpr int ROWS = NG * RB_RG + 2;
    constexpr int EW = W / 2 + 1, ES = EW + 2;  // 19 x 19 coarse cells cover the region
    __shared__ float R[ROWS * RB_S];
    __shared__ float B[ROWS * RB_S];
    __shared__ float E[EW * ES];
    __shared__ double s_red[(THREADS + 31) / 32];
    if (scal[blockIdx.y].done)
        return;
    const int t = threadIdx.x, h = t % RB_HP, y = t / RB_HP;
    const int tile = lf.tile_list[blockIdx.x];
    const int ty = tile / lf.tiles_x, tx = tile % lf.tiles_x;
    const int64_t r0 = (int64_t)ty * TILE_H, c0 = (int64_t)tx * TILE_W;
    const int row0 = RB_RG * y;
    unsigned long long redm, blkm;
    colour_masks<H>(lf, ty, tx, h, redm, blkm);
    const unsigned rm = (unsigned)(redm >> row0) & 15u, bm = (unsigned)(blkm >> row0) & 15u;
    const int64_t gr = r0 - H + row0, gc = c0 - H + 2 * h;
    const int64_t goff = (int64_t)blockIdx.y * lf.plane + gr * lf.pitch + gc;
    // ---- global loads first: red x on the whole region, b where an update needs it, the coarse correction
    float xv[RB_RG], bred[RB_RG], bblk[RB_RG];
    {
        const float* xp = xr + (int64_t)blockIdx.y * (lf.plane >> 1) + gr * (lf.pitch >> 1);
        const BT* bp = b + goff;
        BT vr[RB_RG], vb[RB_RG];
#pragma unroll
        for (int k = 0; k < RB_RG; ++k) {
            const int row = row0 + k, cr = 2 * h + (k & 1), cb = 2 * h + 1 - (k & 1);
            bool red = (rm >> k) & 1, blk = (bm >> k) & 1;
            xv[k] = red ? xp[k * (lf.pitch >> 1) + ((gc + (k & 1)) >> 1)] : 0.f;
            vr[k] = (red && row >= H && row < H + TILE_H && cr >= H && cr < H + TILE_W) ? bp[k * lf.pitch + (k & 1)] : BT(0);
            vb[k] = (blk && row >= 1 && row < W - 1 && cb >= 1 && cb < W - 1) ? bp[k * lf.pitch + 1 - (k & 1)] : BT(0);
        }
        const float* e = ec + (int64_t)blockIdx.y * lc.plane;
        const int64_t I0 = (r0 >> 1) - 1, J0 = (c0 >> 1) - 1;
        for (int i = t; i < EW * EW; i += THREADS) {
            int ei = i / EW, ej = i - ei * EW;
            int64_t I = I0 + ei, J = J0 + ej;
            E[ei * ES + ej] = (I >= 0 && I < lc.rows && J >= 0 && J < lc.cols) ? e[I * lc.pitch + J] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < RB_RG; ++k) {
            bred[k] = (float)vr[k];
            bblk[k] = (float)vb[k];
        }
    }
    __syncthreads();
    const int sb = (row0 + 1) * RB_S + h;
    // ---- R = x + P e at red cells (bilinear; region row / column parity = global parity).  Even rows: the red cell
    //      sits on a coarse point; odd rows: in the middle of four.
    if (h < W / 2) {
#pragma unroll
        for (int k = 0; k < RB_RG; ++k) {
            const float* p = E + ((row0 + k) >> 1) * ES + h;
            float pe = (k & 1) ? 0.25f * ((p[0] + p[1]) + (p[ES] + p[ES + 1])) : p[0];
            R[sb + k * RB_S] = ((rm >> k) & 1) ? xv[k] + pe : 0.f;
        }
    }
    __syncthreads();
    double acc = 0.0;
    float* xo = x_out + goff;
    // ---- black half-sweep on rows / columns 1 .. W-2; the tile's own black cells are final
    {
        const float* p = R + sb;
        float n = p[-RB_S], c = p[0];
#pragma unroll
        for (int k = 0; k < RB_RG; ++k) {
            const int row = row0 + k, col = 2 * h + 1 - (k & 1);
            float s = p[(k + 1) * RB_S];
            float side = p[k * RB_S + ((k & 1) ? -1 : 1)];
            float v = rb_winv<FIXED>(lf, gr + k, gc + 1 - (k & 1)) * (bblk[k] + ((n + s) + (c + side)));
            bool on = ((bm >> k) & 1) && row >= 1 && row < W - 1 && col >= 1 && col < W - 1;
            B[sb + k * RB_S] = on ? v : 0.f;
            if (on && row >= H && row < H + TILE_H && col >= H && col < H + TILE_W) {
                xo[k * lf.pitch + 1 - (k & 1)] = v;
                if (DOT)
                    acc += (double)bblk[k] * (double)v;
            }
            n = c;
            c = s;
        }
    }
    __syncthreads();
    // ---- red half-sweep on the tile itself
    {
        const float* p = B + sb;
        float n = p[-RB_S], c = p[0];
#pragma unroll
        for (int k = 0; k < RB_RG; ++k) {
            const int row = row0 + k, col = 2 * h + (k & 1);
            float s = p[(k + 1) * RB_S];
            float side = p[k * RB_S + ((k & 1) ? 1 : -1)];
            float v = rb_winv<FIXED>(lf, gr + k, gc + (k & 1)) * (bred[k] + ((n + s) + (c + side)));
            if (((rm >> k) & 1) && row >= H && row < H + TILE_H && col >= H && col < H + TILE_W) {
                xo[k * lf.pitch + (k & 1)] = v;
                if (DOT)
                    acc += (double)bred[k] * (double)v;
            }
            n = c;
            c = s;
        }
    }
    if (DOT) {
        for (int o = 16; o; o >>= 1)
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((t & 31) == 0)
            s_red[t >> 5] = acc;
        __syncthreads();
        if (t == 0) {
            double sum = 0.0;
            for (int w = 0; w < (THREADS + 31) / 32; ++w)
                sum += s_red[w];
            if (sum != 0.0)
                atomicAdd(&scal[blockIdx.y].rz[slot], sum);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// coarsest level: K forward (red, black) then K reverse (black, red) Gauss-Seidel sweeps from zero, in place in
// global memory, one CTA per band over the active tiles of the level (a handful; usually one).  Global writes of a
// CTA are visible to its own threads after __syncthreads().
// ---------------------------------------------------------------------------------------------------------------
template <bool FIXED, typename BT, bool DOT>
__global__ void __launch_bounds__(1024) k_rb_coarsest(Level lv, const BT* __restrict__ b, float* __restrict__ x,
    BandScalars* __restrict__ scal, int slot, int sweeps)
{
    __shared__ double s_red[32];
    if (scal[blockIdx.x].done)
        return;
    const int t = threadIdx.x, lr = t >> 5, lc = t & 31;
    const BT* bb = b + (int64_t)blockIdx.x * lv.plane;
    float* xb = x + (int64_t)blockIdx.x * lv.plane;
    for (int ti = 0; ti < lv.n_tiles; ++ti) {  // zero start
        int tile = lv.tile_list[ti];
        int64_t idx = ((int64_t)(tile / lv.tiles_x) * TILE_H + lr) * lv.pitch + (int64_t)(tile % lv.tiles_x) * TILE_W + lc;
        if (lv.umask[idx])
            xb[idx] = 0.f;
    }
    __syncthreads();
    for (int hs = 0; hs < 4 * sweeps; ++hs) {
        // half-sweep colours: forward sweeps red, black, red, black ...; reverse sweeps black, red, ...
        int colour = hs < 2 * sweeps ? (hs & 1) : 1 - (hs & 1);  // 0 = red
        for (int ti = 0; ti < lv.n_tiles; ++ti) {
            int tile = lv.tile_list[ti];
            int64_t r = (int64_t)(tile / lv.tiles_x) * TILE_H + lr, c = (int64_t)(tile % lv.tiles_x) * TILE_W + lc;
            int64_t idx = r * lv.pitch + c;
            if (((r + c) & 1) == colour && lv.umask[idx]) {
                float nb = (xb[idx - lv.pitch] + xb[idx + lv.pitch]) + (xb[idx - 1] + xb[idx + 1]);
                xb[idx] = rb_winv<FIXED>(lv, r, c) * ((float)bb[idx] + nb);
            }
        }
        __syncthreads();
    }
    if (DOT) {
        double acc = 0.0;
        for (int ti = 0; ti < lv.n_tiles; ++ti) {
            int tile = lv.tile_list[ti];
            int64_t idx = ((int64_t)(tile / lv.tiles_x) * TILE_H + lr) * lv.pitch + (int64_t)(tile % lv.tiles_x) * TILE_W + lc;
            if (lv.umask[idx])
                acc += (double)bb[idx] * (double)xb[idx];
        }
        for (int o = 16; o; o >>= 1)
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lc == 0)
            s_red[lr] = acc;
        __syncthreads();
        if (t == 0) {
            double sum = 0.0;
            for (int w = 0; w < 32; ++w)
                sum += s_red[w];
            if (sum != 0.0)
                atomicAdd(&scal[blockIdx.x].rz[slot], sum);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct RBLevel {
    Level lv;
    int64_t units;
    const void* b;  // level 0: the CG residual (double); coarse levels: float
    float* x;       // full plane (level 0: z)
    float* xr;      // colour-split half plane
};

template <typename BT>
int launch_down(sa_ctx* ctx, const RBLevel& F, const RBLevel& C, int nb, const BandScalars* scal)
{
    dim3 grid((unsigned)F.lv.n_tiles, (unsigned)nb);
    if (F.lv.fixed_diag)
        SA_LAUNCH(ctx, (k_rb_down<true, BT>), grid, RB_HP * 10, 0, F.lv, C.lv, (const BT*)F.b, F.xr, (float*)C.b, scal);
    else
        SA_LAUNCH(ctx, (k_rb_down<false, BT>), grid, RB_HP * 10, 0, F.lv, C.lv, (const BT*)F.b, F.xr, (float*)C.b, scal);
    return SA_OK;
}

template <typename BT, bool DOT>
int launch_up(sa_ctx* ctx, const RBLevel& F, const RBLevel& C, int nb, BandScalars* scal, int slot)
{
    dim3 grid((unsigned)F.lv.n_tiles, (unsigned)nb);
    if (F.lv.fixed_diag)
        SA_LAUNCH(ctx, (k_rb_up<true, BT, DOT>), grid, RB_HP * 9, 0, F.lv, C.lv, F.xr, (const BT*)F.b, C.x, F.x, scal, slot);
    else
        SA_LAUNCH(ctx, (k_rb_up<false, BT, DOT>), grid, RB_HP * 9, 0, F.lv, C.lv, F.xr, (const BT*)F.b, C.x, F.x, scal, slot);
    return SA_OK;
}

template <typename BT, bool DOT>
int launch_coarsest(sa_ctx* ctx, const RBLevel& L, int nb, BandScalars* scal, int slot, int sweeps)
{
    if (L.lv.fixed_diag)
        SA_LAUNCH(ctx, (k_rb_coarsest<true, BT, DOT>), nb, 1024, 0, L.lv, (const BT*)L.b, L.x, scal, slot, sweeps);
    else
        SA_LAUNCH(ctx, (k_rb_coarsest<false, BT, DOT>), nb, 1024, 0, L.lv, (const BT*)L.b, L.x, scal, slot, sweeps);
    return SA_OK;
}

}  // namespace

// z (float, in s->z) = M^-1 r for every band that is not done; r.z is accumulated into rz[rz_slot].
// Storage: the level buffers allocated by mg.cu (double-sized) are used as float planes.
int apply_vcycle_rb(sa_scene* s, const sa_options& o, KernelTimer& kt, int rz_slot, int live_bands)
{
    sa_ctx* ctx = s->ctx;
    const int nb = s->nbands;
    std::vector<RBLevel> L;
    L.push_back({ fine_level(s), s->n_unknowns * live_bands, s->plane0(s->r, 0), (float*)s->z + s->pitch,
        (float*)s->t + (s->pitch >> 1) });
    for (sa_level_store& c : s->coarse) {
        if (c.lv.n_tiles == 0)
            break;
        L.push_back({ c.lv, c.n_unknowns * live_bands, (float*)c.b + c.lv.pitch, (float*)c.x + c.lv.pitch,
            (float*)c.t + (c.lv.pitch >> 1) });
    }
    const int nl = (int)L.size();
    BandScalars* scal = s->scal;
    const int coarse_sweeps = 16;
    if (nl == 1) {
        kt.begin(KC_SMOOTH, L[0].units);
        SA_TRY((launch_coarsest<double, true>(ctx, L[0], nb, scal, rz_slot, coarse_sweeps)));
        kt.end();
        SA_CUDA(ctx, cudaGetLastError());
        return SA_OK;
    }
    for (int l = 0; l < nl - 1; ++l) {
        kt.begin(l == 0 ? KC_MG_DOWN : KC_MG_DOWN_COARSE, L[l].units);
        if (l == 0)
            SA_TRY(launch_down<double>(ctx, L[l], L[l + 1], nb, scal));
        else
            SA_TRY(launch_down<float>(ctx, L[l], L[l + 1], nb, scal));
        kt.end();
    }
    kt.begin(KC_SMOOTH, L[nl - 1].units);
    SA_TRY((launch_coarsest<float, false>(ctx, L[nl - 1], nb, scal, 0, coarse_sweeps)));
    kt.end();
    for (int l = nl - 2; l >= 0; --l) {
        kt.begin(l == 0 ? KC_MG_UP : KC_MG_UP_COARSE, L[l].units);
        if (l == 0)
            SA_TRY((launch_up<double, true>(ctx, L[l], L[l + 1], nb, scal, rz_slot)));
        else
            SA_TRY((launch_up<float, false>(ctx, L[l], L[l + 1], nb, scal, 0)));
        kt.end();
    }
    SA_CUDA(ctx, cudaGetLastError());
    return SA_OK;
}

}  // namespace satfill
