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

#if SATFILL_LEGACY_VARIANTS
namespace satfill {

namespace {

constexpr int RB_HP = 20;  // threads per row group (half columns): lane-linear shared-memory addressing
#ifndef SATFILL_RB_DOWN_RG
#define SATFILL_RB_DOWN_RG 14
#endif
#ifndef SATFILL_RB_UP_RG
#define SATFILL_RB_UP_RG 12
#endif
// rows per thread: even, so that the row parity is the parity of the unrolled row index
constexpr int RB_DOWN_RG = SATFILL_RB_DOWN_RG, RB_UP_RG = SATFILL_RB_UP_RG;
// row stride S of the colour-split arrays with RG * S = 20 (mod 32): shared-memory bank = thread id + constant
__host__ __device__ constexpr int rb_stride(int rg) { return rg == 4 ? 21 : (rg == 6 ? 30 : (rg == 10 ? 34 : (rg == 12 ? 23 : (rg == 14 ? 22 : -1)))); }
constexpr unsigned long long EVEN_ROWS = 0x5555555555555555ull;

template <bool FIXED>
__device__ __forceinline__ float rb_winv(const Level& lv, int64_t r, int64_t c)
{
    if (FIXED)
        return 0.25f;
    int n = (r > 0) + (r < lv.rows - 1) + (c > 0) + (c < lv.cols - 1);
    return n == 4 ? 0.25f : (n == 3 ? (1.0f / 3.0f) : (n == 2 ? 0.5f : 1.0f));
}

// red / black unknown bits of the thread's half column (frame columns 2h, 2h + 1): bit i <=> frame row i
template <int H>
__device__ __forceinline__ void colour_masks(const Level& lv, int ty, int tx, int h, unsigned long long& red,
    unsigned long long& black)
{
    unsigned long long cm0 = region_col_mask<H>(lv, ty, tx, 2 * h);
    unsigned long long cm1 = region_col_mask<H>(lv, ty, tx, 2 * h + 1);
    red = (cm0 & EVEN_ROWS) | (cm1 & ~EVEN_ROWS);
    black = (cm1 & EVEN_ROWS) | (cm0 & ~EVEN_ROWS);
}

// bit k <=> frame row row0 + k lies in [lo, hi)
template <int RG>
__device__ __forceinline__ unsigned rb_rows(int row0, int lo, int hi)
{
    return (unsigned)((((1ull << hi) - 1) & ~((1ull << lo) - 1)) >> row0) & ((1u << RG) - 1);
}
// bit k <=> the column of the thread's red (black) cell in row row0 + k lies in [lo, hi); red cells sit in column
// 2h + (k & 1), black cells in column 2h + 1 - (k & 1)
template <int RG>
__device__ __forceinline__ unsigned rb_cols(int h, int lo, int hi, bool red)
{
    constexpr unsigned KM = (1u << RG) - 1, EV = 0x55555555u & KM, OD = 0xAAAAAAAAu & KM;
    bool c0 = 2 * h >= lo && 2 * h < hi, c1 = 2 * h + 1 >= lo && 2 * h + 1 < hi;
    return red ? ((c0 ? EV : 0u) | (c1 ? OD : 0u)) : ((c1 ? EV : 0u) | (c0 ? OD : 0u));
}

// predicated read-only loads: one instruction, no branch; 0 when the predicate is off
__device__ __forceinline__ float ldg_if(const float* p, unsigned pred)
{
    float v;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\tmov.f32 %0, 0f00000000;\n\t@q ld.global.nc.f32 %0, [%1];\n\t}"
        : "=f"(v)
        : "l"(p), "r"(pred));
    return v;
}
__device__ __forceinline__ float2 ldg2_if(const float* p, unsigned pred)
{
    float2 v;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\tmov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\t"
        "@q ld.global.nc.v2.f32 {%0, %1}, [%2];\n\t}"
        : "=f"(v.x), "=f"(v.y)
        : "l"(p), "r"(pred));
    return v;
}

// A row walker: a byte address advanced by the row pitch with one 64-bit add (the compiler otherwise keeps an element
// index and re-derives the address -- four integer instructions per row and pointer)
template <typename T>
struct RowPtr {
    unsigned long long a;
    __device__ __forceinline__ RowPtr(const T* p) : a((unsigned long long)p) {}
    __device__ __forceinline__ T* get() const { return (T*)a; }
    __device__ __forceinline__ void step(unsigned long long bytes) { a += bytes; }
};

// Accesses predicated on bit k of a row mask.  SATFILL_RB_CXX_PRED = 1 writes them as C++ conditionals, which the
// compiler turns into one LOP3-to-predicate + a predicated access (5 % fewer instructions than the asm forms, which need
// the bit as a 0 / 1 register first) -- and which measured 15 % SLOWER on the GPU: the compiler then orders each load
// next to its use instead of issuing all of a thread's loads back to back.  Kept as a switch; the asm forms are used.
#ifndef SATFILL_RB_CXX_PRED
#define SATFILL_RB_CXX_PRED 0
#endif
__device__ __forceinline__ float ldg_bit(const float* p, unsigned mask, int k)
{
    if (!SATFILL_RB_CXX_PRED)
        return ldg_if(p, (mask >> k) & 1);
    float v = 0.f;
    if (mask & (1u << k))
        v = __ldg(p);
    return v;
}
__device__ __forceinline__ float2 ldg2_bit(const float* p, unsigned mask, int k)
{
    if (!SATFILL_RB_CXX_PRED)
        return ldg2_if(p, (mask >> k) & 1);
    float2 v = make_float2(0.f, 0.f);
    if (mask & (1u << k))
        v = __ldg(reinterpret_cast<const float2*>(p));
    return v;
}

// predicated stores: one instruction, no branch
__device__ __forceinline__ void stg_if(float* p, float v, unsigned pred)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.f32 [%0], %1;\n\t}" ::"l"(p), "f"(v), "r"(pred) : "memory");
}
__device__ __forceinline__ void stg2_if(float* p, float x, float y, unsigned pred)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q st.global.v2.f32 [%0], {%1, %2};\n\t}" ::"l"(p), "f"(x), "f"(y),
                 "r"(pred)
                 : "memory");
}

__device__ __forceinline__ void stg_bit(float* p, float v, unsigned mask, int k)
{
    if (!SATFILL_RB_CXX_PRED) {
        stg_if(p, v, (mask >> k) & 1);
        return;
    }
    if (mask & (1u << k))
        *p = v;
}
__device__ __forceinline__ void stg2_bit(float* p, float x, float y, unsigned mask, int k)
{
    if (!SATFILL_RB_CXX_PRED) {
        stg2_if(p, x, y, (mask >> k) & 1);
        return;
    }
    if (mask & (1u << k))
        *reinterpret_cast<float2*>(p) = make_float2(x, y);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// descent: pre-smoothing from zero, residual, restriction.  The dependence region has a halo of 3 cells; it is
// framed with a halo of 4 (40 x 40, outer ring unused) so that a half column is an 8-byte aligned pair of cells.
// ---------------------------------------------------------------------------------------------------------------
constexpr int RB_DOWN_NG = (TILE_W + 8 + RB_DOWN_RG - 1) / RB_DOWN_RG;
constexpr int RB_UP_NG = (TILE_W + 4 + RB_UP_RG - 1) / RB_UP_RG;

// WINV: the level carries its own 1 / diagonal plane (coarse levels, Level::winv)
template <bool FIXED, bool WINV>
__global__ void __launch_bounds__(RB_HP * RB_DOWN_NG) k_rb_down(Level lf, Level lc, const float* __restrict__ b,
    float* __restrict__ xr, float* __restrict__ bc, const BandScalars* __restrict__ scal)
{
    constexpr int H = 4, W = TILE_W + 2 * H;  // frame 40 x 40; cells used: frame rows / columns 1 .. 38
    constexpr int RG = RB_DOWN_RG, NG = RB_DOWN_NG, S = rb_stride(RG);
    static_assert(S > 0 && (RG * S) % 32 == 20 && RG % 2 == 0, "unsupported rows per thread");
    constexpr int THREADS = RB_HP * NG;
    constexpr int ROWS = NG * RG + 2;  // one spare row above and below: the sliding window needs no bounds checks
    constexpr unsigned KM = (1u << RG) - 1;
    __shared__ float R[ROWS * S];
    __shared__ float B[ROWS * S];
    if (scal[blockIdx.y].done)
        return;
    const int t = threadIdx.x, y = t / RB_HP, h = t - y * RB_HP;
    const int yx = lf.tile_yx[blockIdx.x];
    const int ty = yx >> 16, tx = yx & 0xffff;
    const int row0 = RG * y;
    const int pitch = (int)lf.pitch, pitch2 = pitch >> 1;
    unsigned long long redm, blkm;
    colour_masks<H>(lf, ty, tx, h, redm, blkm);
    const unsigned rm = (unsigned)(redm >> row0) & KM & rb_rows<RG>(row0, 1, W - 1) & rb_cols<RG>(h, 1, W - 1, true);
    const unsigned bm = (unsigned)(blkm >> row0) & KM & rb_rows<RG>(row0, 2, W - 2) & rb_cols<RG>(h, 2, W - 2, false);
    const int64_t gr = (int64_t)ty * TILE_H - H, gc = (int64_t)tx * TILE_W - H;  // global position of the frame origin
    // plane offset of the thread's (row0, column 2h); rows are walked by adding the pitch to a pointer (a fresh 64-bit
    // address per row costs half a dozen integer instructions, and this kernel is bound by instruction issue)
    const int64_t toff = (gr + row0) * lf.pitch + gc + 2 * h;
    // ---- global loads: the right-hand side at the thread's RG x 2 cells, one aligned pair per row
    float bred[RG], bblk[RG], wred[RG], wblk[RG];
    {
        RowPtr<const float> bp(b + (int64_t)blockIdx.y * lf.plane + toff), wp(WINV ? lf.winv + toff : nullptr);
        const unsigned long long pb = (unsigned long long)pitch * sizeof(float);
        const unsigned ld = rm | bm;
#pragma unroll
        for (int k = 0; k < RG; ++k, bp.step(pb), wp.step(pb)) {
            float2 v = ldg2_bit(bp.get(), ld, k);
            bred[k] = (k & 1) ? v.y : v.x;
            bblk[k] = (k & 1) ? v.x : v.y;
            if (WINV) {
                float2 w = ldg2_bit(wp.get(), ld, k);
                wred[k] = (k & 1) ? w.y : w.x;
                wblk[k] = (k & 1) ? w.x : w.y;
            } else {
                wred[k] = rb_winv<FIXED>(lf, gr + row0 + k, gc + 2 * h + (k & 1));
                wblk[k] = rb_winv<FIXED>(lf, gr + row0 + k, gc + 2 * h + 1 - (k & 1));
            }
        }
    }
    // mask injection for the restriction at the end (coarse unknown <=> fine (2I, 2J) unknown): the row words are
    // fetched now, so that their latency is long gone when they are used
    constexpr int NCOARSE = (TILE_H / 2) * (TILE_W / 2), CPT = (NCOARSE + THREADS - 1) / THREADS;
    uint32_t crow[CPT];
    {
        const uint32_t* rowbits = lf.tbits + ((size_t)(ty + 1) * lf.tb_stride + (tx + 1)) * 32;
#pragma unroll
        for (int q = 0; q < CPT; ++q) {
            int i = t + q * THREADS;
            crow[q] = i < NCOARSE ? __ldg(rowbits + 2 * (i >> 4)) : 0u;
        }
    }
    const int sb = (row0 + 1) * S + h;  // shared index of (row0, h)
    // ---- red half-sweep from zero: x = b / d (pointwise); the tile's own red cells go to HBM colour-split
    {
        RowPtr<float> xo(xr + (int64_t)blockIdx.y * (lf.plane >> 1) + (gr + row0) * pitch2 + (gc >> 1) + h);
        const unsigned long long pb2 = (unsigned long long)pitch2 * sizeof(float);
        const unsigned own = rm & rb_rows<RG>(row0, H, H + TILE_H) & rb_cols<RG>(h, H, H + TILE_W, true);
#pragma unroll
        for (int k = 0; k < RG; ++k, xo.step(pb2)) {
            float v = wred[k] * (((rm >> k) & 1) ? bred[k] : 0.f);
            R[sb + k * S] = v;
            stg_bit(xo.get(), v, own, k);
        }
    }
    __syncthreads();
    // ---- black half-sweep on frame rows / columns 2 .. 37
    {
        const float* p = R + sb;
        float n = p[-S], c = p[0];
#pragma unroll
        for (int k = 0; k < RG; ++k) {
            float s = p[(k + 1) * S];
            float side = p[k * S + ((k & 1) ? -1 : 1)];
            float v = wblk[k] * (bblk[k] + ((n + s) + (c + side)));
            B[sb + k * S] = ((bm >> k) & 1) ? v : 0.f;
            n = c;
            c = s;
        }
    }
    __syncthreads();
    // ---- residual: zero at black cells; at a red cell b - d x + sum(black neighbours) = sum(black neighbours)
    {
        const unsigned on = rm & rb_rows<RG>(row0, 3, W - 3) & rb_cols<RG>(h, 3, W - 3, true);
        const float* p = B + sb;
        float n = p[-S], c = p[0];
#pragma unroll
        for (int k = 0; k < RG; ++k) {
            float s = p[(k + 1) * S];
            float side = p[k * S + ((k & 1) ? 1 : -1)];
            float v = (n + s) + (c + side);
            R[sb + k * S] = ((on >> k) & 1) ? v : 0.f;  // R is dead as an iterate: reuse it for the residual
            n = c;
            c = s;
        }
    }
    __syncthreads();
    // ---- full-weighting restriction: coarse (ci, cj) <-> tile cell (2ci, 2cj) = frame (2ci + 4, 2cj + 4), a red
    //      cell in an even row (half column cj + 2); its diagonal neighbours are the red cells of the rows above and
    //      below in half columns cj + 1 and cj + 2; its edge neighbours are black (zero residual).
    {
        float* bco = bc + (int64_t)blockIdx.y * lc.plane + (int64_t)(ty * (TILE_H / 2)) * lc.pitch + tx * (TILE_W / 2);
        const int cpitch = (int)lc.pitch;
#pragma unroll
        for (int q = 0; q < CPT; ++q) {
            const int i = t + q * THREADS;
            const int ci = i >> 4, cj = i & 15;
            if (i < NCOARSE) {
                const float* p = R + (2 * ci + H + 1) * S + cj + 2;
                const float v = p[0] + 0.25f * ((p[-S - 1] + p[-S]) + (p[S - 1] + p[S]));
                stg_if(bco + (ci * cpitch + cj), v, (crow[q] >> (2 * cj)) & 1);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// ascent: prolongation + correction (red cells only), post-smoothing black then red, (level 0) r.z
// ---------------------------------------------------------------------------------------------------------------
template <bool FIXED, bool DOT, bool WINV>
__global__ void __launch_bounds__(RB_HP * RB_UP_NG) k_rb_up(Level lf, Level lc, const float* __restrict__ xr,
    const float* __restrict__ b, const float* __restrict__ ec, float* __restrict__ x_out, BandScalars* __restrict__ scal,
    int slot)
{
    constexpr int H = 2, W = TILE_W + 2 * H;  // 36
    constexpr int RG = RB_UP_RG, NG = RB_UP_NG, S = rb_stride(RG);
    static_assert(S > 0 && (RG * S) % 32 == 20 && RG % 2 == 0, "unsupported rows per thread");
    constexpr int THREADS = RB_HP * NG;
    constexpr int ROWS = NG * RG + 2;
    constexpr int EW = W / 2 + 1, ES = RB_HP + 1;  // 19 x 19 coarse cells cover the region
    constexpr int EROWS = (NG * RG) / 2 + 2;
    constexpr unsigned KM = (1u << RG) - 1;
    __shared__ float R[ROWS * S];
    __shared__ float B[ROWS * S];
    __shared__ float E[EROWS * ES];
    __shared__ float s_acc[DOT ? THREADS : 1];
    if (scal[blockIdx.y].done)
        return;
    const int t = threadIdx.x, y = t / RB_HP, h = t - y * RB_HP;
    const int yx = lf.tile_yx[blockIdx.x];
    const int ty = yx >> 16, tx = yx & 0xffff;
    const int row0 = RG * y;
    const int pitch = (int)lf.pitch, pitch2 = pitch >> 1;
    unsigned long long redm = 0, blkm = 0;
    if (h < W / 2)
        colour_masks<H>(lf, ty, tx, h, redm, blkm);
    const unsigned rm = (unsigned)(redm >> row0) & KM;
    const unsigned on_blk = (unsigned)(blkm >> row0) & KM & rb_rows<RG>(row0, 1, W - 1) & rb_cols<RG>(h, 1, W - 1, false);
    const unsigned own_rows = rb_rows<RG>(row0, H, H + TILE_H);
    const unsigned own_red = rm & own_rows & rb_cols<RG>(h, H, H + TILE_W, true);
    const unsigned own_blk = on_blk & own_rows & rb_cols<RG>(h, H, H + TILE_W, false);
    const int64_t gr = (int64_t)ty * TILE_H - H, gc = (int64_t)tx * TILE_W - H;
    // plane offset of the thread's (row0, column 2h); rows are walked by pointer increments (see k_rb_down)
    const int64_t toff = (gr + row0) * lf.pitch + gc + 2 * h;
    const int64_t goff = (int64_t)blockIdx.y * lf.plane + toff;
    // ---- global loads first: red x on the whole region, b where an update needs it, the coarse correction
    float xv[RG], bred[RG], bblk[RG], wred[RG], wblk[RG];
    {
        RowPtr<const float> xp(xr + (int64_t)blockIdx.y * (lf.plane >> 1) + (gr + row0) * pitch2 + (gc >> 1) + h);
        RowPtr<const float> bp(b + goff), wp(WINV ? lf.winv + toff : nullptr);
        const unsigned long long pb = (unsigned long long)pitch * sizeof(float), pb2 = (unsigned long long)pitch2 * sizeof(float);
        const unsigned ld = own_red | on_blk;
#pragma unroll
        for (int k = 0; k < RG; ++k, xp.step(pb2), bp.step(pb), wp.step(pb)) {
            xv[k] = ldg_bit(xp.get(), rm, k);
            float2 v = ldg2_bit(bp.get(), ld, k);
            bred[k] = (k & 1) ? v.y : v.x;
            bblk[k] = (k & 1) ? v.x : v.y;
            if (WINV) {
                float2 w = ldg2_bit(wp.get(), ld, k);
                wred[k] = (k & 1) ? w.y : w.x;
                wblk[k] = (k & 1) ? w.x : w.y;
            } else {
                wred[k] = rb_winv<FIXED>(lf, gr + row0 + k, gc + 2 * h + (k & 1));
                wblk[k] = rb_winv<FIXED>(lf, gr + row0 + k, gc + 2 * h + 1 - (k & 1));
            }
        }
        const int cpitch = (int)lc.pitch;
        const int I0 = ty * (TILE_H / 2) - 1, J = tx * (TILE_W / 2) - 1 + h;
        const float* e = ec + (int64_t)blockIdx.y * lc.plane + (int64_t)I0 * lc.pitch + J;
        const bool jok = h < EW && J >= 0 && J < lc.cols;
#pragma unroll
        for (int q = 0; q < (EW + NG - 1) / NG; ++q) {
            int ei = y + q * NG;
            if (ei < EW && h < EW)
                E[ei * ES + h] = ldg_if(e + ei * cpitch, jok && I0 + ei >= 0 && I0 + ei < lc.rows);
        }
    }
    __syncthreads();
    const int sb = (row0 + 1) * S + h;
    // ---- R = x + P e at red cells (bilinear; region row / column parity = global parity).  Even rows: the red cell
    //      sits on a coarse point; odd rows: in the middle of four.
    if (h < W / 2) {
        const float* p = E + (row0 >> 1) * ES + h;
#pragma unroll
        for (int k = 0; k < RG; ++k) {
            const float* q = p + (k >> 1) * ES;
            float pe = (k & 1) ? 0.25f * ((q[0] + q[1]) + (q[ES] + q[ES + 1])) : q[0];
            R[sb + k * S] = ((rm >> k) & 1) ? xv[k] + pe : 0.f;
        }
    }
    __syncthreads();
    float acc = 0.f;
    float vblk[RG];
    // ---- black half-sweep on rows / columns 1 .. W-2; the tile's own black cells are final
    {
        const float* p = R + sb;
        float n = p[-S], c = p[0];
#pragma unroll
        for (int k = 0; k < RG; ++k) {
            float s = p[(k + 1) * S];
            float side = p[k * S + ((k & 1) ? -1 : 1)];
            float v = wblk[k] * (bblk[k] + ((n + s) + (c + side)));
            v = ((on_blk >> k) & 1) ? v : 0.f;
            B[sb + k * S] = v;
            vblk[k] = v;
            if (DOT)
                acc += ((own_blk >> k) & 1) ? bblk[k] * v : 0.f;
            n = c;
            c = s;
        }
    }
    __syncthreads();
    // ---- red half-sweep on the tile itself; both cells of the half column leave as one aligned pair (a cell that
    //      is not an unknown is written as the zero it already holds)
    {
        RowPtr<float> xo(x_out + goff);
        const unsigned long long pb = (unsigned long long)pitch * sizeof(float);
        const unsigned st = own_red | own_blk;
        const float* p = B + sb;
        float n = p[-S], c = p[0];
#pragma unroll
        for (int k = 0; k < RG; ++k, xo.step(pb)) {
            float s = p[(k + 1) * S];
            float side = p[k * S + ((k & 1) ? 1 : -1)];
            float v = wred[k] * (bred[k] + ((n + s) + (c + side)));
            v = ((own_red >> k) & 1) ? v : 0.f;
            if (DOT)
                acc += bred[k] * v;
            stg2_bit(xo.get(), (k & 1) ? vblk[k] : v, (k & 1) ? v : vblk[k], st, k);
            n = c;
            c = s;
        }
    }
    if (DOT) {
        // per-thread partial sums (a dozen products) in float, everything above that in double.  The CTA is not a whole
        // number of warps, so the partial sums go through shared memory and the (always complete) first warp adds them.
        s_acc[t] = acc;
        __syncthreads();
        if (t < 32) {
            double a = 0.0;
            for (int i = t; i < THREADS; i += 32)
                a += (double)s_acc[i];
            for (int o = 16; o; o >>= 1)
                a += __shfl_xor_sync(0xffffffffu, a, o);
            if (t == 0 && a != 0.0)
                atomicAdd(&scal[blockIdx.y].rz[slot], a);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// coarsest level: K forward (red, black) then K reverse (black, red) Gauss-Seidel sweeps from zero, in place in
// global memory, one CTA per band over the active tiles of the level (a handful; usually one).  Global writes of a
// CTA are visible to its own threads after __syncthreads().
// ---------------------------------------------------------------------------------------------------------------
template <bool FIXED, bool DOT>
__global__ void __launch_bounds__(1024) k_rb_coarsest(Level lv, const float* __restrict__ b, float* __restrict__ x,
    BandScalars* __restrict__ scal, int slot, int sweeps)
{
    __shared__ double s_red[32];
    if (scal[blockIdx.x].done)
        return;
    const int t = threadIdx.x, lr = t >> 5, lc = t & 31;
    const float* bb = b + (int64_t)blockIdx.x * lv.plane;
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
                xb[idx] = (lv.winv ? lv.winv[idx] : rb_winv<FIXED>(lv, r, c)) * (bb[idx] + nb);
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
    float* b;  // level 0: the float copy of the CG residual
    float* x;       // full plane (level 0: z)
    float* xr;      // colour-split half plane
};

int launch_down(sa_ctx* ctx, const RBLevel& F, const RBLevel& C, int nb, const BandScalars* scal)
{
    if (F.lv.n_tiles == 0)
        return SA_OK;
    dim3 grid((unsigned)F.lv.n_tiles, (unsigned)nb);
    if (F.lv.winv)
        SA_LAUNCH(ctx, (k_rb_down<true, true>), grid, RB_HP * RB_DOWN_NG, 0, F.lv, C.lv, F.b, F.xr, C.b, scal);
    else if (F.lv.fixed_diag)
        SA_LAUNCH(ctx, (k_rb_down<true, false>), grid, RB_HP * RB_DOWN_NG, 0, F.lv, C.lv, F.b, F.xr, C.b, scal);
    else
        SA_LAUNCH(ctx, (k_rb_down<false, false>), grid, RB_HP * RB_DOWN_NG, 0, F.lv, C.lv, F.b, F.xr, C.b, scal);
    return SA_OK;
}

template <bool DOT>
int launch_up(sa_ctx* ctx, const RBLevel& F, const RBLevel& C, int nb, BandScalars* scal, int slot)
{
    if (F.lv.n_tiles == 0)
        return SA_OK;
    dim3 grid((unsigned)F.lv.n_tiles, (unsigned)nb);
    if (F.lv.winv)
        SA_LAUNCH(ctx, (k_rb_up<true, DOT, true>), grid, RB_HP * RB_UP_NG, 0, F.lv, C.lv, F.xr, F.b, C.x, F.x, scal, slot);
    else if (F.lv.fixed_diag)
        SA_LAUNCH(ctx, (k_rb_up<true, DOT, false>), grid, RB_HP * RB_UP_NG, 0, F.lv, C.lv, F.xr, F.b, C.x, F.x, scal, slot);
    else
        SA_LAUNCH(ctx, (k_rb_up<false, DOT, false>), grid, RB_HP * RB_UP_NG, 0, F.lv, C.lv, F.xr, F.b, C.x, F.x, scal, slot);
    return SA_OK;
}

template <bool DOT>
int launch_coarsest(sa_ctx* ctx, const RBLevel& L, int nb, BandScalars* scal, int slot, int sweeps)
{
    if (L.lv.fixed_diag)
        SA_LAUNCH(ctx, (k_rb_coarsest<true, DOT>), nb, 1024, 0, L.lv, L.b, L.x, scal, slot, sweeps);
    else
        SA_LAUNCH(ctx, (k_rb_coarsest<false, DOT>), nb, 1024, 0, L.lv, L.b, L.x, scal, slot, sweeps);
    return SA_OK;
}

}  // namespace

// z (float, in s->z) = M^-1 r for every band that is not done; r.z is accumulated into rz[rz_slot].
// Storage: the level buffers allocated by mg.cu (double-sized) are used as float planes.
int apply_vcycle_rb(sa_scene* s, const sa_options& o, KernelTimer& kt, int rz_slot, int live_bands)
{
    sa_ctx* ctx = s->ctx;
    const int nb = s->win_n(), b0 = s->band0;  // the band window (common.cuh): every base pointer starts at band b0
    std::vector<RBLevel> L;
    L.push_back({ fine_level(s), s->n_unknowns * live_bands, s->rb_rf(), s->rb_z(),
        (float*)s->t + (s->pitch >> 1) + (int64_t)b0 * (s->plane >> 1) });
    for (sa_level_store& c : s->coarse) {
        if (c.lv.n_tiles == 0)
            break;
        L.push_back({ c.lv, c.n_unknowns * live_bands, (float*)c.b + c.lv.pitch + (int64_t)b0 * c.lv.plane,
            (float*)c.x + c.lv.pitch + (int64_t)b0 * c.lv.plane, (float*)c.t + (c.lv.pitch >> 1) + (int64_t)b0 * (c.lv.plane >> 1) });
    }
    const int nl = (int)L.size();
    BandScalars* scal = s->scal + b0;
    const int coarse_sweeps = 16;
    // row decomposition (dist.cu): levels below dist_levels run on the rank's slice and exchange halo rows, the others
    // are replicated on every rank
    const bool dist = s->distributed && s->dist_planned && ctx->world > 1;
    const int dlv = dist ? s->dist_levels : 0;
    for (int l = 0; l < nl && l < dlv; ++l)
        L[l].lv = dist_level(s, l, L[l].lv);
    if (nl == 1) {
        kt.begin(KC_SMOOTH, L[0].units);
        SA_TRY((launch_coarsest<true>(ctx, L[0], nb, scal, rz_slot, coarse_sweeps)));
        kt.end();
        SA_CUDA(ctx, cudaGetLastError());
        return SA_OK;
    }
    for (int l = 0; l < nl - 1; ++l) {
        kt.begin(l == 0 ? KC_MG_DOWN : KC_MG_DOWN_COARSE, L[l].units);
        SA_TRY(launch_down(ctx, L[l], L[l + 1], nb, scal));
        kt.end();
        if (l < dlv) {
            // the ascent reads the red half of the iterate 2 rows beyond the slice; the next level's descent reads
            // its right-hand side 3 rows beyond -- or, if that level is replicated, everywhere
            SA_TRY(dist_group_begin(s));  // one NCCL launch for both exchanges
            SA_TRY(dist_halo<float>(s, l, L[l].xr, L[l].lv.pitch >> 1, L[l].lv.plane >> 1, 2, 2));
            if (l + 1 < dlv)
                SA_TRY(dist_halo<float>(s, l + 1, L[l + 1].b, L[l + 1].lv.pitch, L[l + 1].lv.plane, 3, 3));
            else
                SA_TRY(dist_gather(s, L[l + 1].b, L[l + 1].lv.pitch, L[l + 1].lv.plane));
            SA_TRY(dist_group_end(s));
        }
    }
    kt.begin(KC_SMOOTH, L[nl - 1].units);
    SA_TRY((launch_coarsest<false>(ctx, L[nl - 1], nb, scal, 0, coarse_sweeps)));
    kt.end();
    for (int l = nl - 2; l >= 0; --l) {
        kt.begin(l == 0 ? KC_MG_UP : KC_MG_UP_COARSE, L[l].units);
        if (l == 0)
            SA_TRY((launch_up<true>(ctx, L[l], L[l + 1], nb, scal, rz_slot)));
        else
            SA_TRY((launch_up<false>(ctx, L[l], L[l + 1], nb, scal, 0)));
        kt.end();
        // the finer level's ascent interpolates from up to 2 coarse rows beyond.  (Level 0: CG's direction needs 1 row of z;
        // the caller exchanges it in one group with the all-reduce of r.z -- cg.cu.)
        if (l < dlv && l > 0)
            SA_TRY(dist_halo<float>(s, l, L[l].x, L[l].lv.pitch, L[l].lv.plane, 2, 2));
    }
    SA_CUDA(ctx, cudaGetLastError());
    return SA_OK;
}

}  // namespace satfill
#else
namespace satfill {
int apply_vcycle_rb(sa_scene* s, const sa_options&, KernelTimer&, int, int)
{
    return fail(s->ctx, SA_BAD_ARGUMENT, "SA_MG_RB32_CTA needs a library built with SATFILL_LEGACY_VARIANTS");
}
}  // namespace satfill
#endif  // SATFILL_LEGACY_VARIANTS
