// Red-black multigrid V(1,1)-cycle, second generation: ONE WARP PER TILE, the tile's neighbourhood in REGISTERS.
//
// Same arithmetic as mg_rb.cu (same hierarchy, transfers, smoother and float storage: see the header of that file), a
// different machine mapping.  ncu on the first generation (one 60-thread CTA per tile and band, the (32 + 2H)^2
// neighbourhood colour-split in shared memory; profiles/r2a_*): the same 2.0 ns per tile whether the tile is full or
// half empty, instruction issue at 70 % and the shared-memory / L1 pipe at 70 %, DRAM at a third of its peak -- the
// kernels are bound by per-tile work, half of which is shared-memory traffic for values a thread already held.
//
// Here a lane owns a QUAD of four consecutive columns (one aligned 16-byte load per row) of RG consecutive rows; ten
// lanes span the 40-column frame, three row groups the 36 .. 42 frame rows: 30 of 32 lanes busy.  In a 5-point
// red-black sweep a cell's vertical neighbours and two of the three horizontal ones are the lane's own registers; the
// remaining one comes from the adjacent lane by ONE warp shuffle per row and sweep (plus two per sweep and lane for the
// rows of the group above / below).  No shared memory, no barrier: a warp never waits for another warp.
//   descent  (3 sweeps + restriction):  ~46 shuffles per lane instead of ~120 shared-memory accesses per thread
//   ascent   (prolongation + 2 sweeps): ~40 shuffles per lane
// and about 2.3x fewer instructions per tile.  The sweeps run in place: a red sweep only reads black cells and vice
// versa.  Cells outside the dependence cone of the tile (the outer rings of the frame) compute garbage that nothing
// consumes; only the unknown bits are ever needed as masks, and they come from the per-tile column words (Level::tbitsT)
// of the tile and its neighbours: three 16-byte loads per lane and tile, shared by all bands.
//
// Persistent kernels: a warp walks work items (tile, chunk of bands) of the raster-ordered tile list, so that
// neighbouring warps work on neighbouring tiles at the same time (their frame halos meet in L2) and the column masks
// of a tile serve all bands of the chunk.
//
// The coarse TAIL of the cycle -- every level from the first one that no longer fills the GPU down to the coarsest
// and back up -- runs in ONE cooperative launch with grid-wide barriers between levels (k_rbw_tail): on the 10980^2
// benchmark tile that replaces 2 x 7 + 1 launches of a few dozen CTAs each.
#include "common.cuh"
#include "tile.cuh"

#include <cstdlib>

namespace satfill {

namespace {

constexpr int RW_WARPS = 4;                 // warps per CTA of the tail kernel (the per-level kernels run one-warp CTAs)
constexpr int RW_THREADS = 32 * RW_WARPS;
// resident one-warp CTAs per SM the per-level kernels are compiled for (= the register budget: 65536 / 32 / CTAs)
#ifndef SATFILL_RBW_DOWN_CTAS
#define SATFILL_RBW_DOWN_CTAS 20
#endif
#ifndef SATFILL_RBW_UP_CTAS
#define SATFILL_RBW_UP_CTAS 16
#endif
#ifndef SATFILL_RBW_W_CTAS
#define SATFILL_RBW_W_CTAS 12
#endif
// The kernels are bound by memory latency at the occupancy their registers allow (ncu: 5 of 16 warps per SM wait on a
// long scoreboard per issue slot, issue slots 46 % busy, DRAM 45 - 64 %): while a band of a tile is computed, the rows of the
// NEXT band of the same tile are already on their way into L2 (prefetch.global.L2, predicated like the loads; `next` =
// the distance in elements to that band, 0 for the last band of the item).
#ifndef SATFILL_RBW_PREFETCH
#define SATFILL_RBW_PREFETCH 0  // measured: 1.6 - 2x SLOWER with the prefetches on (gpurun_out/r2f: 48 -> 80 ms descent, 54 -> 110 ms ascent)
#endif
constexpr int RBW_DOWN_CTAS = SATFILL_RBW_DOWN_CTAS, RBW_UP_CTAS = SATFILL_RBW_UP_CTAS, RBW_W_CTAS = SATFILL_RBW_W_CTAS;
constexpr unsigned FULLW = 0xffffffffu;
constexpr int DN_RG = 14, DN_HR = 4;        // descent: 3 x 14 = 42 >= 40 frame rows (halo 4: dependence cone 3, aligned)
constexpr int UP_RG = 12, UP_HR = 2;        // ascent:  3 x 12 = 36 frame rows (halo 2)
constexpr int HC = 4;                       // column halo of both frames: 40 columns = 10 quads

// Accesses predicated on (mask & bit) != 0: an AND with an immediate and a predicated access, no branch; loads give 0 when
// the predicate is off.  (Written as asm so that all loads of a lane are issued back to back: a C++ conditional makes the
// compiler order each load next to its use -- measured 15 % slower on the first generation.)
__device__ __forceinline__ float4 ldg4_if(const float* p, unsigned mask, unsigned bit)
{
    float4 v;
    asm("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %5, %6;\n\tsetp.ne.u32 q, t, 0;\n\tmov.f32 %0, 0f00000000;\n\t"
        "mov.f32 %1, 0f00000000;\n\tmov.f32 %2, 0f00000000;\n\tmov.f32 %3, 0f00000000;\n\t"
        "@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
        : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
        : "l"(p), "r"(mask), "r"(bit));
    return v;
}
__device__ __forceinline__ float2 ldg2_ifw(const float* p, unsigned mask, unsigned bit)
{
    float2 v;
    asm("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %3, %4;\n\tsetp.ne.u32 q, t, 0;\n\tmov.f32 %0, 0f00000000;\n\t"
        "mov.f32 %1, 0f00000000;\n\t@q ld.global.nc.v2.f32 {%0, %1}, [%2];\n\t}"
        : "=f"(v.x), "=f"(v.y)
        : "l"(p), "r"(mask), "r"(bit));
    return v;
}
__device__ __forceinline__ void stg4_if(float* p, float a, float b, float c, float d, unsigned mask, unsigned bit)
{
    asm volatile("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %5, %6;\n\tsetp.ne.u32 q, t, 0;\n\t"
                 "@q st.global.v4.f32 [%0], {%1, %2, %3, %4};\n\t}" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "r"(mask), "r"(bit)
                 : "memory");
}
__device__ __forceinline__ void stg2_ifw(float* p, float x, float y, unsigned mask, unsigned bit)
{
    asm volatile("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %3, %4;\n\tsetp.ne.u32 q, t, 0;\n\t"
                 "@q st.global.v2.f32 [%0], {%1, %2};\n\t}" ::"l"(p), "f"(x), "f"(y), "r"(mask), "r"(bit)
                 : "memory");
}
// the line holding p into L2, if (mask & bit) != 0
__device__ __forceinline__ void prefetch_l2_if(const void* p, unsigned mask, unsigned bit)
{
    asm volatile("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %1, %2;\n\tsetp.ne.u32 q, t, 0;\n\t@q prefetch.global.L2 [%0];\n\t}" ::"l"(p),
                 "r"(mask), "r"(bit));
}
// value if bit k of the mask is set, else 0.  ZW: the value already carries a factor 1 / d that is zero at every cell that
// is not an unknown (the 1 / d plane of a coarse level): no select needed
template <bool ZW>
__device__ __forceinline__ float keep(unsigned mask, int k, float v)
{
    return ZW ? v : ((mask & (1u << k)) ? v : 0.f);
}

// Unknown bits of the four columns of quad q (frame columns 4q .. 4q + 3 <-> tile columns 4q - 4 ..) of the frame of tile
// (ty, tx): bit i of cm[j] <=> frame row i (= tile row i - HR) of that column holds an unknown.  Quad 0 lies in the tile to
// the west, quad 9 in the tile to the east; a ring of all-zero tiles surrounds the grid (Level::tbitsT).
template <int HR>
__device__ __forceinline__ void quad_col_masks(const Level& lv, int ty, int tx, int q, unsigned long long cm[4])
{
    const int txx = tx + (q == 0 ? -1 : (q == 9 ? 1 : 0));
    const int col = (4 * q - HC) & 31;
    const uint32_t* w = lv.tbitsT + ((size_t)(ty + 1) * lv.tb_stride + (txx + 1)) * 32 + col;
    const size_t vs = (size_t)lv.tb_stride * 32;
    const uint4 C = __ldg(reinterpret_cast<const uint4*>(w));
    const uint4 N = __ldg(reinterpret_cast<const uint4*>(w - vs));
    const uint4 S = __ldg(reinterpret_cast<const uint4*>(w + vs));
    const unsigned c[4] = { C.x, C.y, C.z, C.w }, n[4] = { N.x, N.y, N.z, N.w }, s[4] = { S.x, S.y, S.z, S.w };
#pragma unroll
    for (int j = 0; j < 4; ++j)
        cm[j] = ((unsigned long long)n[j] >> (32 - HR)) | ((unsigned long long)c[j] << HR) | ((unsigned long long)s[j] << (32 + HR));
}

template <bool FIXED>
__device__ __forceinline__ float winv_of(const Level& lv, int64_t r, int64_t c)
{
    if (FIXED)
        return 0.25f;
    const int n = (r > 0) + (r < lv.rows - 1) + (c > 0) + (c < lv.cols - 1);
    return n == 4 ? 0.25f : (n == 3 ? (1.0f / 3.0f) : (n == 2 ? 0.5f : 1.0f));
}

// 1 / d from the coordinates (level 0 of a Poisson scene: d = in-image neighbour count, poisson.cpp:187-190) without a
// register per cell: row k of the lane has degree 2 unless it is the first or last image row, likewise the columns
struct WCoord {
    int top_k, bot_k, dc[4];
    __device__ __forceinline__ WCoord(const Level& lv, int64_t r0, int64_t c0)
    {
        top_k = (int)-r0;
        bot_k = (int)(lv.rows - 1 - r0);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            dc[j] = 2 - (c0 + j == 0) - (c0 + j == lv.cols - 1);
    }
    __device__ __forceinline__ float at(int k, int j) const
    {
        const int n = 2 - (k == top_k) - (k == bot_k) + dc[j];
        return n == 4 ? 0.25f : (n == 3 ? (1.0f / 3.0f) : (n == 2 ? 0.5f : 1.0f));
    }
};

// bit k set <=> k even / odd, for RG rows
template <int RG>
struct Par {
    static constexpr unsigned KM = (1u << RG) - 1, EV = 0x55555555u & KM, OD = 0xAAAAAAAAu & KM;
};

// ---------------------------------------------------------------------------------------------------------------
// descent of one tile and band: red half-sweep from zero (x = b / d, pointwise), black half-sweep, residual (zero at
// black cells; at a red cell the sum of its black neighbours), full-weighting restriction onto the coarse cells of
// the tile.  Frame: rows ty*32 - 4 .., columns tx*32 - 4 ..; lane = 10 g + q owns rows [14 g, 14 g + 14) of quad q.
// In a row of even frame parity the red cells are columns 0 and 2 of the quad, in an odd row columns 1 and 3.
// WMODE: 0 = 1/d is 1/4 everywhere, 1 = the level carries a 1/d plane (coarse levels), 2 = 1/d from the coordinates
// ---------------------------------------------------------------------------------------------------------------
template <int WMODE>
__device__ __forceinline__ void rbw_down_tile(const Level& lf, int64_t cpitch, int ty, int tx, int lane, unsigned alive,
    const unsigned long long cm[4], const float* __restrict__ b, float* __restrict__ bc, int64_t next)
{
    constexpr int RG = DN_RG, HR = DN_HR;
    constexpr unsigned KM = Par<RG>::KM, EV = Par<RG>::EV, OD = Par<RG>::OD;
    const bool live = lane < 30;
    const int g = live ? lane / 10 : 2, q = live ? lane - 10 * g : 9;
    const int row0 = RG * g;
    // rows of the lane that lie inside the 40-row frame
    // (alive = 0: the band has converged -- every access is predicated off and the arithmetic runs on zeros; a branch around
    // the tile would make the warp shuffles below conditional, which costs a WARPSYNC / ENDCOLLECTIVE pair around each)
    const unsigned rows_in = (live ? ((row0 + RG <= 32 + 2 * HR) ? KM : ((1u << (32 + 2 * HR - row0)) - 1)) : 0u) & alive;
    const unsigned c0 = (unsigned)(cm[0] >> row0) & rows_in, c1 = (unsigned)(cm[1] >> row0) & rows_in;
    const unsigned c2 = (unsigned)(cm[2] >> row0) & rows_in, c3 = (unsigned)(cm[3] >> row0) & rows_in;
    const unsigned anyq = c0 | c1 | c2 | c3;
    // per-row bits of the two red and the two black cells of the quad
    const unsigned rA = (c0 & EV) | (c1 & OD), rB = (c2 & EV) | (c3 & OD);
    const unsigned kA = (c1 & EV) | (c0 & OD), kB = (c3 & EV) | (c2 & OD);
    const int pitch = (int)lf.pitch;
    const int64_t gr = (int64_t)ty * TILE_H - HR, gc = (int64_t)tx * TILE_W - HC;
    const int64_t toff = (gr + row0) * lf.pitch + gc + 4 * q;
    const unsigned long long pb = (unsigned long long)pitch * sizeof(float);
    float v[RG][4], wv[WMODE == 1 ? RG : 1][4];
    const WCoord wc(lf, gr + row0, gc + 4 * q);
    auto W = [&](int k, int j) -> float { return WMODE == 0 ? 0.25f : (WMODE == 1 ? wv[WMODE == 1 ? k : 0][j] : wc.at(k, j)); };
    {
        unsigned long long bp = (unsigned long long)(b + toff), wp = (unsigned long long)(WMODE == 1 ? lf.winv + toff : nullptr);
#pragma unroll
        for (int k = 0; k < RG; ++k, bp += pb, wp += pb) {
            const float4 t = ldg4_if((const float*)bp, anyq, 1u << k);
            v[k][0] = t.x, v[k][1] = t.y, v[k][2] = t.z, v[k][3] = t.w;
            if (WMODE == 1) {
                const float4 w = ldg4_if((const float*)wp, anyq, 1u << k);
                wv[k][0] = w.x, wv[k][1] = w.y, wv[k][2] = w.z, wv[k][3] = w.w;
            }
        }
        if (SATFILL_RBW_PREFETCH && next) {
            unsigned long long np = (unsigned long long)(b + toff + next);
#pragma unroll
            for (int k = 0; k < RG; ++k, np += pb)
                prefetch_l2_if((const void*)np, anyq, 1u << k);
        }
    }
    // ---- red half-sweep from zero: x = b / d, pointwise.  Nothing of it is stored: the ascent recomputes it from the same
    //      b with the same multiplication (the first generation carried the red half of the iterate through HBM: 2 B
    //      written and 2 B read per unknown and level, a fifth of the cycle's traffic).
    {
#pragma unroll
        for (int k = 0; k < RG; ++k) {
            const int j0 = k & 1, j1 = 2 + (k & 1);
            // masked like every other cell update: the cycle's vectors are only ever trusted at the unknowns of the CURRENT
            // mask (a mask change leaves them unscrubbed: cg.cu, stale_rb)
            v[k][j0] = keep<WMODE == 1>(rA, k, v[k][j0] * W(k, j0));
            v[k][j1] = keep<WMODE == 1>(rB, k, v[k][j1] * W(k, j1));
        }
    }
    // ---- black half-sweep (in place: a black update reads red cells only)
    {
        // red cells of the adjacent row groups: row -1 (odd: columns 1, 3) and row RG (even: columns 0, 2)
        const float n1 = __shfl_up_sync(FULLW, v[RG - 1][1], 10), n3 = __shfl_up_sync(FULLW, v[RG - 1][3], 10);
        const float s0 = __shfl_down_sync(FULLW, v[0][0], 10), s2 = __shfl_down_sync(FULLW, v[0][2], 10);
#pragma unroll
        for (int k = 0; k < RG; ++k) {
            if ((k & 1) == 0) {  // black: columns 1 and 3; east of column 3 is the next lane's column 0
                const float e = __shfl_down_sync(FULLW, v[k][0], 1);
                const float na = k > 0 ? v[k > 0 ? k - 1 : 0][1] : n1, nb = k > 0 ? v[k > 0 ? k - 1 : 0][3] : n3;
                const float sa = v[k + 1][1], sb = v[k + 1][3];  // RG is even: row k + 1 exists
                const float xa = W(k, 1) * (v[k][1] + ((na + sa) + (v[k][0] + v[k][2])));
                const float xb = W(k, 3) * (v[k][3] + ((nb + sb) + (v[k][2] + e)));
                v[k][1] = keep<WMODE == 1>(kA, k, xa);
                v[k][3] = keep<WMODE == 1>(kB, k, xb);
            } else {  // black: columns 0 and 2; west of column 0 is the previous lane's column 3
                const float w = __shfl_up_sync(FULLW, v[k][3], 1);
                const float na = v[k - 1][0], nb = v[k - 1][2];
                const float sa = k < RG - 1 ? v[k < RG - 1 ? k + 1 : k][0] : s0, sb = k < RG - 1 ? v[k < RG - 1 ? k + 1 : k][2] : s2;
                const float xa = W(k, 0) * (v[k][0] + ((na + sa) + (w + v[k][1])));
                const float xb = W(k, 2) * (v[k][2] + ((nb + sb) + (v[k][1] + v[k][3])));
                v[k][0] = keep<WMODE == 1>(kA, k, xa);
                v[k][2] = keep<WMODE == 1>(kB, k, xb);
            }
        }
    }
    // ---- residual: zero at black cells; at a red cell  b - d x + sum(black neighbours) = sum(black neighbours)
    {
        // black cells of the adjacent row groups: row -1 (odd: columns 0, 2) and row RG (even: columns 1, 3)
        const float n0 = __shfl_up_sync(FULLW, v[RG - 1][0], 10), n2 = __shfl_up_sync(FULLW, v[RG - 1][2], 10);
        const float s1 = __shfl_down_sync(FULLW, v[0][1], 10), s3 = __shfl_down_sync(FULLW, v[0][3], 10);
#pragma unroll
        for (int k = 0; k < RG; ++k) {
            if ((k & 1) == 0) {  // red: columns 0 and 2; west of column 0 is the previous lane's column 3
                const float w = __shfl_up_sync(FULLW, v[k][3], 1);
                const float na = k > 0 ? v[k > 0 ? k - 1 : 0][0] : n0, nb = k > 0 ? v[k > 0 ? k - 1 : 0][2] : n2;
                const float ra = (na + v[k + 1][0]) + (w + v[k][1]);
                const float rb = (nb + v[k + 1][2]) + (v[k][1] + v[k][3]);
                v[k][0] = keep<false>(rA, k, ra);
                v[k][2] = keep<false>(rB, k, rb);
            } else {  // red: columns 1 and 3; east of column 3 is the next lane's column 0
                const float e = __shfl_down_sync(FULLW, v[k][0], 1);
                const float sa = k < RG - 1 ? v[k < RG - 1 ? k + 1 : k][1] : s1, sb = k < RG - 1 ? v[k < RG - 1 ? k + 1 : k][3] : s3;
                const float ra = (v[k - 1][1] + sa) + (v[k][0] + v[k][2]);
                const float rb = (v[k - 1][3] + sb) + (v[k][2] + e);
                v[k][1] = keep<false>(rA, k, ra);
                v[k][3] = keep<false>(rB, k, rb);
            }
        }
    }
    // ---- full-weighting restriction: coarse (ci, cj) <-> frame (2 ci + 4, 2 cj + 4), a red cell of an even row in
    //      column 0 or 2 of a quad; its edge neighbours are black (zero residual), its diagonal neighbours the red
    //      cells of the odd rows above and below: columns -1 (the previous lane's column 3), 1 and 3.
    {
        const float u1 = __shfl_up_sync(FULLW, v[RG - 1][1], 10), u3 = __shfl_up_sync(FULLW, v[RG - 1][3], 10);
        const float ul = __shfl_up_sync(FULLW, v[RG - 1][3], 11);
        float left[RG / 2];  // the previous lane's column 3 in the odd rows
#pragma unroll
        for (int m = 0; m < RG / 2; ++m)
            left[m] = __shfl_up_sync(FULLW, v[2 * m + 1][3], 1);
        const bool ownq = q >= 1 && q <= 8;
        const unsigned own_rows = ((((1ull << (HR + TILE_H)) - 1) & ~((1ull << HR) - 1)) >> row0) & KM;
        const unsigned st = ownq ? ((c0 | c2) & own_rows & EV) : 0u;
        // coarse row of frame row row0 + k: ty * 16 + (row0 + k - 4) / 2
        unsigned long long bo = (unsigned long long)(bc + ((int64_t)ty * (TILE_H / 2) + ((row0 - HR) >> 1)) * cpitch
            + (int64_t)tx * (TILE_W / 2) + 2 * (q - 1));
        const unsigned long long cb2 = (unsigned long long)cpitch * sizeof(float);
#pragma unroll
        for (int m = 0; m < RG / 2; ++m, bo += cb2) {
            const int k = 2 * m;
            const float a1 = m > 0 ? v[m > 0 ? k - 1 : 0][1] : u1, a3 = m > 0 ? v[m > 0 ? k - 1 : 0][3] : u3;
            const float al = m > 0 ? left[m > 0 ? m - 1 : 0] : ul;
            const float ca = v[k][0] + 0.25f * ((al + a1) + (left[m] + v[k + 1][1]));
            const float cb = v[k][2] + 0.25f * ((a1 + a3) + (v[k + 1][1] + v[k + 1][3]));
            // a coarse cell of the pair that is not an unknown gets the zero it already holds
            stg2_ifw((float*)bo, keep<false>(c0, k, ca), keep<false>(c2, k, cb), st, 1u << k);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// ascent of one tile and band: x = x_red + P e at red cells (bilinear prolongation of the coarse correction), black
// half-sweep, red half-sweep on the tile itself; returns the lane's share of b . x over the tile (level 0: r . z).
// Frame: rows ty*32 - 2 .., columns tx*32 - 4 ..; lane = 10 g + q owns rows [12 g, 12 g + 12) of quad q.
// ---------------------------------------------------------------------------------------------------------------
template <int WMODE, bool DOT>
__device__ __forceinline__ float rbw_up_tile(const Level& lf, int64_t cpitch, int ty, int tx, int lane, unsigned alive,
    const unsigned long long cm[4], const float* __restrict__ b, const float* __restrict__ ec, float* __restrict__ x_out, int64_t next)
{
    constexpr int RG = UP_RG, HR = UP_HR;
    constexpr unsigned KM = Par<RG>::KM, EV = Par<RG>::EV, OD = Par<RG>::OD;
    const bool live = lane < 30;
    const int g = live ? lane / 10 : 2, q = live ? lane - 10 * g : 9;
    const int row0 = RG * g;
    const unsigned liv = (live ? KM : 0u) & alive;  // 3 x 12 rows are exactly the 36-row frame
    const unsigned c0 = (unsigned)(cm[0] >> row0) & liv, c1 = (unsigned)(cm[1] >> row0) & liv;
    const unsigned c2 = (unsigned)(cm[2] >> row0) & liv, c3 = (unsigned)(cm[3] >> row0) & liv;
    const unsigned anyq = c0 | c1 | c2 | c3;
    const unsigned rA = (c0 & EV) | (c1 & OD), rB = (c2 & EV) | (c3 & OD);
    const unsigned kA = (c1 & EV) | (c0 & OD), kB = (c3 & EV) | (c2 & OD);
    // unknown bits of the coarse cells the lane interpolates from: coarse row m (m = 0 .. RG / 2) <-> frame row row0 + 2 m,
    // coarse columns <-> quad columns 0 and 2
    const unsigned cc = (live ? (unsigned)(((cm[0] | cm[2]) >> row0) & ((1u << (RG + 1)) - 1)) : 0u) & alive;
    const int pitch = (int)lf.pitch;
    const int64_t gr = (int64_t)ty * TILE_H - HR, gc = (int64_t)tx * TILE_W - HC;
    const int64_t toff = (gr + row0) * lf.pitch + gc + 4 * q;
    const unsigned long long pb = (unsigned long long)pitch * sizeof(float);
    float v[RG][4], bv[RG][4], wv[WMODE == 1 ? RG : 1][4];
    float2 e[RG / 2 + 1];
    const WCoord wc(lf, gr + row0, gc + 4 * q);
    auto W = [&](int k, int j) -> float { return WMODE == 0 ? 0.25f : (WMODE == 1 ? wv[WMODE == 1 ? k : 0][j] : wc.at(k, j)); };
    {
        unsigned long long bp = (unsigned long long)(b + toff), wp = (unsigned long long)(WMODE == 1 ? lf.winv + toff : nullptr);
#pragma unroll
        for (int k = 0; k < RG; ++k, bp += pb, wp += pb) {
            const float4 t = ldg4_if((const float*)bp, anyq, 1u << k);
            bv[k][0] = t.x, bv[k][1] = t.y, bv[k][2] = t.z, bv[k][3] = t.w;
            if (WMODE == 1) {
                const float4 w = ldg4_if((const float*)wp, anyq, 1u << k);
                wv[k][0] = w.x, wv[k][1] = w.y, wv[k][2] = w.z, wv[k][3] = w.w;
            }
        }
        // coarse correction: frame (row0 + 2 m, 4 q) <-> coarse (ty * 16 - 1 + row0 / 2 + m, tx * 16 - 2 + 2 q)
        unsigned long long ep = (unsigned long long)(ec + ((int64_t)ty * (TILE_H / 2) - (HR >> 1) + (row0 >> 1)) * cpitch
            + (int64_t)tx * (TILE_W / 2) - (HC >> 1) + 2 * q);
        const unsigned long long cb2 = (unsigned long long)cpitch * sizeof(float);
#pragma unroll
        for (int m = 0; m <= RG / 2; ++m, ep += cb2)
            e[m] = ldg2_ifw((const float*)ep, cc, 1u << (2 * m));
        if (SATFILL_RBW_PREFETCH && next) {
            unsigned long long nb = (unsigned long long)(b + toff + next);
#pragma unroll
            for (int k = 0; k < RG; ++k, nb += pb)
                prefetch_l2_if((const void*)nb, anyq, 1u << k);
        }
    }
    // ---- x = x_red + P e at the red cells, x_red = b / d being the descent's red half-sweep from zero, recomputed (the
    //      frame's row / column parity is the global one).  Even rows: the red cells sit on coarse points; odd rows: in
    //      the middle of four, the easternmost of them the next lane's first.
    {
#pragma unroll
        for (int k = 0; k < RG; ++k) {
            const int m = k >> 1;
            v[k][k & 1] = bv[k][k & 1] * W(k, k & 1);
            v[k][2 + (k & 1)] = bv[k][2 + (k & 1)] * W(k, 2 + (k & 1));
            v[k][1 - (k & 1)] = 0.f;
            v[k][3 - (k & 1)] = 0.f;
            if ((k & 1) == 0) {
                v[k][0] = keep<false>(rA, k, v[k][0] + e[m].x);
                v[k][2] = keep<false>(rB, k, v[k][2] + e[m].y);
            } else {
                const float ex0 = __shfl_down_sync(FULLW, e[m].x, 1), ex1 = __shfl_down_sync(FULLW, e[m + 1].x, 1);
                const float pa = 0.25f * ((e[m].x + e[m].y) + (e[m + 1].x + e[m + 1].y));
                const float pb_ = 0.25f * ((e[m].y + ex0) + (e[m + 1].y + ex1));
                v[k][1] = keep<false>(rA, k, v[k][1] + pa);
                v[k][3] = keep<false>(rB, k, v[k][3] + pb_);
            }
        }
    }
    float acc = 0.f;
    // rows / lanes of the tile itself
    const bool ownq = q >= 1 && q <= 8;
    const unsigned own_rows = ownq ? (unsigned)(((((1ull << (HR + TILE_H)) - 1) & ~((1ull << HR) - 1)) >> row0) & KM) : 0u;
    // ---- black half-sweep (in place)
    {
        const float n1 = __shfl_up_sync(FULLW, v[RG - 1][1], 10), n3 = __shfl_up_sync(FULLW, v[RG - 1][3], 10);
        const float s0 = __shfl_down_sync(FULLW, v[0][0], 10), s2 = __shfl_down_sync(FULLW, v[0][2], 10);
#pragma unroll
        for (int k = 0; k < RG; ++k) {
            float xa, xb;
            if ((k & 1) == 0) {
                const float ee = __shfl_down_sync(FULLW, v[k][0], 1);
                const float na = k > 0 ? v[k > 0 ? k - 1 : 0][1] : n1, nb = k > 0 ? v[k > 0 ? k - 1 : 0][3] : n3;
                xa = W(k, 1) * (bv[k][1] + ((na + v[k + 1][1]) + (v[k][0] + v[k][2])));
                xb = W(k, 3) * (bv[k][3] + ((nb + v[k + 1][3]) + (v[k][2] + ee)));
                xa = keep<WMODE == 1>(kA, k, xa);
                xb = keep<WMODE == 1>(kB, k, xb);
                v[k][1] = xa, v[k][3] = xb;
            } else {
                const float ww = __shfl_up_sync(FULLW, v[k][3], 1);
                const float sa = k < RG - 1 ? v[k < RG - 1 ? k + 1 : k][0] : s0, sb = k < RG - 1 ? v[k < RG - 1 ? k + 1 : k][2] : s2;
                xa = W(k, 0) * (bv[k][0] + ((v[k - 1][0] + sa) + (ww + v[k][1])));
                xb = W(k, 2) * (bv[k][2] + ((v[k - 1][2] + sb) + (v[k][1] + v[k][3])));
                xa = keep<WMODE == 1>(kA, k, xa);
                xb = keep<WMODE == 1>(kB, k, xb);
                v[k][0] = xa, v[k][2] = xb;
            }
        }
    }
    // ---- red half-sweep; the tile's own cells are final and leave as whole quads (a cell that is not an unknown is
    //      written as the zero it already holds)
    {
        const float n0 = __shfl_up_sync(FULLW, v[RG - 1][0], 10), n2 = __shfl_up_sync(FULLW, v[RG - 1][2], 10);
        const float s1 = __shfl_down_sync(FULLW, v[0][1], 10), s3 = __shfl_down_sync(FULLW, v[0][3], 10);
        unsigned long long xo = (unsigned long long)(x_out + toff);
        const unsigned st = anyq & own_rows;
#pragma unroll
        for (int k = 0; k < RG; ++k, xo += pb) {
            float xa, xb;
            if ((k & 1) == 0) {
                const float ww = __shfl_up_sync(FULLW, v[k][3], 1);
                const float na = k > 0 ? v[k > 0 ? k - 1 : 0][0] : n0, nb = k > 0 ? v[k > 0 ? k - 1 : 0][2] : n2;
                xa = W(k, 0) * (bv[k][0] + ((na + v[k + 1][0]) + (ww + v[k][1])));
                xb = W(k, 2) * (bv[k][2] + ((nb + v[k + 1][2]) + (v[k][1] + v[k][3])));
                xa = keep<WMODE == 1>(rA, k, xa);
                xb = keep<WMODE == 1>(rB, k, xb);
                if (DOT)  // b . x over the whole quad row: both colours are final here
                    acc += keep<false>(own_rows, k, fmaf(bv[k][0], xa, bv[k][1] * v[k][1]) + fmaf(bv[k][2], xb, bv[k][3] * v[k][3]));
                stg4_if((float*)xo, xa, v[k][1], xb, v[k][3], st, 1u << k);
            } else {
                const float ee = __shfl_down_sync(FULLW, v[k][0], 1);
                const float sa = k < RG - 1 ? v[k < RG - 1 ? k + 1 : k][1] : s1, sb = k < RG - 1 ? v[k < RG - 1 ? k + 1 : k][3] : s3;
                xa = W(k, 1) * (bv[k][1] + ((v[k - 1][1] + sa) + (v[k][0] + v[k][2])));
                xb = W(k, 3) * (bv[k][3] + ((v[k - 1][3] + sb) + (v[k][2] + ee)));
                xa = keep<WMODE == 1>(rA, k, xa);
                xb = keep<WMODE == 1>(rB, k, xb);
                if (DOT)
                    acc += keep<false>(own_rows, k, fmaf(bv[k][1], xa, bv[k][0] * v[k][0]) + fmaf(bv[k][3], xb, bv[k][2] * v[k][2]));
                stg4_if((float*)xo, v[k][0], xa, v[k][2], xb, st, 1u << k);
            }
        }
    }
    return acc;
}

// Work items of a level: (tile, chunk of bands), tile-major, so that consecutive warps work on neighbouring tiles.
// band_major: item = (band, tile), band-major -- the whole GPU sweeps one band's planes at a time (a level-0 plane of the
// benchmark tile is 0.5 GB: with the bands as the inner loop every warp touches a different 2 MB page per iteration and
// array; the first-generation kernels, one CTA per tile and band with the band in blockIdx.y, had this order for free).
struct Items {
    int n_tiles, nbands, bchunk, nchunks, band_major;
    __host__ __device__ int count() const { return band_major ? n_tiles * nbands : n_tiles * nchunks; }
    // item -> tile index and band range
    __device__ __forceinline__ void at(int i, int& ti, int& b0, int& b1) const
    {
        if (band_major) {
            b0 = i / n_tiles;
            ti = i - b0 * n_tiles;
            b1 = b0 + 1;
        } else {
            ti = i / nchunks;
            b0 = (i - ti * nchunks) * bchunk;
            b1 = min(nbands, b0 + bchunk);
        }
    }
};
inline Items make_items(int n_tiles, int nbands, int total_warps, bool band_major)
{
    // as many bands per item as still leave every warp a few items (the column masks of a tile serve the whole chunk)
    int bchunk = nbands;
    while (bchunk > 1 && (int64_t)n_tiles * ((nbands + bchunk - 1) / bchunk) < 3 * (int64_t)total_warps)
        bchunk = (bchunk + 1) / 2;
    return Items { n_tiles, nbands, bchunk, (nbands + bchunk - 1) / bchunk, band_major ? 1 : 0 };
}

template <int WMODE>
__device__ __forceinline__ void rbw_down_items(const Level& lf, const Level& lc, const Items it, int first, int stride, int lane,
    const float* __restrict__ b, float* __restrict__ bc, const BandScalars* __restrict__ scal)
{
    const int n = it.count();
    for (int i = first; i < n; i += stride) {
        int ti, b0, b1;
        it.at(i, ti, b0, b1);
        const int yx = lf.tile_yx[ti];
        const int ty = yx >> 16, tx = yx & 0xffff;
        unsigned long long cm[4];
        quad_col_masks<DN_HR>(lf, ty, tx, lane < 30 ? lane % 10 : 9, cm);
        for (int band = b0; band < b1; ++band) {
            const unsigned alive = scal[band].done ? 0u : ~0u;
            const int64_t next = band + 1 < b1 ? lf.plane : 0;
            if (WMODE == 2) {
                // 1 / d is 1 / 4 unless the frame touches the image border (warp-uniform)
                const bool inner = ty > 0 && tx > 0 && (int64_t)(ty + 1) * TILE_H + DN_HR < lf.rows && (int64_t)(tx + 1) * TILE_W + HC < lf.cols;
                if (inner)
                    rbw_down_tile<0>(lf, lc.pitch, ty, tx, lane, alive, cm, b + (int64_t)band * lf.plane, bc + (int64_t)band * lc.plane, next);
                else
                    rbw_down_tile<2>(lf, lc.pitch, ty, tx, lane, alive, cm, b + (int64_t)band * lf.plane, bc + (int64_t)band * lc.plane, next);
            } else {
                rbw_down_tile<WMODE>(lf, lc.pitch, ty, tx, lane, alive, cm, b + (int64_t)band * lf.plane, bc + (int64_t)band * lc.plane, next);
            }
        }
    }
}

// s_acc: per-warp, per-band partial sums of b . x (DOT), nbands doubles per warp
template <int WMODE, bool DOT>
__device__ __forceinline__ void rbw_up_items(const Level& lf, const Level& lc, const Items it, int first, int stride, int lane,
    const float* __restrict__ b, const float* __restrict__ ec, float* __restrict__ x_out, const BandScalars* __restrict__ scal,
    double* s_acc)
{
    const int n = it.count();
    for (int i = first; i < n; i += stride) {
        int ti, b0, b1;
        it.at(i, ti, b0, b1);
        const int yx = lf.tile_yx[ti];
        const int ty = yx >> 16, tx = yx & 0xffff;
        unsigned long long cm[4];
        quad_col_masks<UP_HR>(lf, ty, tx, lane < 30 ? lane % 10 : 9, cm);
        for (int band = b0; band < b1; ++band) {
            const unsigned alive = scal[band].done ? 0u : ~0u;
            const float* bb = b + (int64_t)band * lf.plane;
            const float* eb = ec + (int64_t)band * lc.plane;
            float* xb = x_out + (int64_t)band * lf.plane;
            const int64_t next = band + 1 < b1 ? lf.plane : 0;
            float acc;
            if (WMODE == 2) {
                const bool inner = ty > 0 && tx > 0 && (int64_t)(ty + 1) * TILE_H + UP_HR < lf.rows && (int64_t)(tx + 1) * TILE_W + HC < lf.cols;
                acc = inner ? rbw_up_tile<0, DOT>(lf, lc.pitch, ty, tx, lane, alive, cm, bb, eb, xb, next)
                            : rbw_up_tile<2, DOT>(lf, lc.pitch, ty, tx, lane, alive, cm, bb, eb, xb, next);
            } else {
                acc = rbw_up_tile<WMODE, DOT>(lf, lc.pitch, ty, tx, lane, alive, cm, bb, eb, xb, next);
            }
            if (DOT) {
                // a few dozen products per lane in float, everything above that in double
                double a = (double)acc;
#pragma unroll
                for (int o = 16; o; o >>= 1)
                    a += __shfl_xor_sync(FULLW, a, o);
                if (lane == 0)
                    s_acc[band] += a;
            }
        }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// one level per launch (the levels that fill the GPU)
// ---------------------------------------------------------------------------------------------------------------
// CTAs of ONE warp: item indices then depend on blockIdx only, the compiler sees every loop as uniform and emits the warp
// shuffles bare (with several warps per CTA the warp index comes from threadIdx and every shuffle gets a WARPSYNC /
// ENDCOLLECTIVE pair: ~100 of 1700 instructions per tile).
template <int WMODE>
__global__ void __launch_bounds__(32, WMODE == 0 ? RBW_DOWN_CTAS : RBW_W_CTAS) k_rbw_down(Level lf, Level lc, Items it,
    const float* __restrict__ b, float* __restrict__ bc, const BandScalars* __restrict__ scal)
{
    rbw_down_items<WMODE>(lf, lc, it, (int)blockIdx.x, (int)gridDim.x, (int)threadIdx.x, b, bc, scal);
}

template <int WMODE, bool DOT>
__global__ void __launch_bounds__(32, WMODE == 0 ? RBW_UP_CTAS : RBW_W_CTAS) k_rbw_up(Level lf, Level lc, Items it,
    const float* __restrict__ b, const float* __restrict__ ec, float* __restrict__ x_out, BandScalars* __restrict__ scal, int slot)
{
    extern __shared__ double s_acc[];  // nbands partial sums of b . x (DOT only)
    const int lane = (int)threadIdx.x;
    if (DOT) {
        for (int i = lane; i < it.nbands; i += 32)
            s_acc[i] = 0.0;
        __syncwarp();
    }
    rbw_up_items<WMODE, DOT>(lf, lc, it, (int)blockIdx.x, (int)gridDim.x, lane, b, ec, x_out, scal, s_acc);
    if (DOT) {
        __syncwarp();
        for (int i = lane; i < it.nbands; i += 32)
            if (s_acc[i] != 0.0)
                atomicAdd(&scal[i].rz[slot], s_acc[i]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// the coarse tail in one cooperative launch
// ---------------------------------------------------------------------------------------------------------------
struct TailLevel {
    Level lv;
    Items it;
    float* b;
    float* x;
};
struct TailArgs {
    TailLevel L[MAX_LEVELS];
    int n;        // levels in the tail; the last one is the coarsest (solved by relaxation, one CTA per band)
    int n_grid;   // the first n_grid of them are worked by the whole grid, the rest per band inside one CTA
    int nbands;
    int sweeps;
    int fixed;    // coarsest level only: 1 / d = 1 / 4 where the level has no 1 / d plane (a one-level hierarchy)
};

namespace {

// All CTAs of the (co-resident: cooperative launch) grid meet here.  `counter` counts arrivals monotonically.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += 1;
        const unsigned target = epoch * gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < target);
        __threadfence();
    }
    __syncthreads();
}

// K forward (red, black) then K reverse (black, red) Gauss-Seidel sweeps from zero on the coarsest level.  One active tile
// (every scene that is not extremely elongated): the tile lives in shared memory with a zero ring (`sm`: 3 x 34 x 34
// floats), 1 / d zero at every cell that is not an unknown, so a half-sweep is branch free and costs one barrier instead
// of a round trip to L2 per cell.  Otherwise in place in global memory: the CTA's own writes are visible to its threads
// after __syncthreads().
constexpr int CP = TILE_W + 2;
template <bool DOT>
__device__ void rbw_coarsest(const Level& lv, bool fixed, const float* __restrict__ bb, float* __restrict__ xb, int sweeps,
    double* dot_out, float* sm)
{
    if (lv.n_tiles == 1) {
        const int t = threadIdx.x, nt = blockDim.x;
        float *sx = sm, *sb = sm + CP * CP, *sw = sm + 2 * CP * CP;
        const int tile = lv.tile_list[0];
        const int64_t r0 = (int64_t)(tile / lv.tiles_x) * TILE_H, c0 = (int64_t)(tile % lv.tiles_x) * TILE_W;
        __syncthreads();  // the previous band's solve is over
        for (int i = t; i < 3 * CP * CP; i += nt)
            sm[i] = 0.f;
        __syncthreads();
        for (int i = t; i < TILE_H * TILE_W; i += nt) {
            const int lr = i >> 5, lc = i & 31;
            const int64_t idx = (r0 + lr) * lv.pitch + c0 + lc;
            if (lv.umask[idx]) {
                sb[(lr + 1) * CP + lc + 1] = bb[idx];
                sw[(lr + 1) * CP + lc + 1] = lv.winv ? lv.winv[idx] : (fixed ? 0.25f : winv_of<false>(lv, r0 + lr, c0 + lc));
            }
        }
        __syncthreads();
        for (int hs = 0; hs < 4 * sweeps; ++hs) {
            const int colour = hs < 2 * sweeps ? (hs & 1) : 1 - (hs & 1);  // 0 = red = (row + column) even
            for (int i = t; i < TILE_H * TILE_W / 2; i += nt) {
                const int lr = i >> 4, lc = 2 * (i & 15) + ((lr + colour) & 1);
                const int p = (lr + 1) * CP + lc + 1;
                sx[p] = sw[p] * (sb[p] + ((sx[p - CP] + sx[p + CP]) + (sx[p - 1] + sx[p + 1])));
            }
            __syncthreads();
        }
        double acc = 0.0;
        for (int i = t; i < TILE_H * TILE_W; i += nt) {
            const int lr = i >> 5, lc = i & 31;
            const int64_t idx = (r0 + lr) * lv.pitch + c0 + lc;
            const int p = (lr + 1) * CP + lc + 1;
            xb[idx] = sx[p];  // the whole tile: zero at every cell that is not an unknown (1 / d is zero there)
            acc += (double)sb[p] * (double)sx[p];
        }
        if (DOT) {
            for (int o = 16; o; o >>= 1)
                acc += __shfl_xor_sync(FULLW, acc, o);
            if ((t & 31) == 0 && acc != 0.0)
                atomicAdd(dot_out, acc);
        }
        return;
    }
    const int t = threadIdx.x, nt = blockDim.x;
    const int cells = lv.n_tiles * (TILE_H * TILE_W);
    for (int i = t; i < cells; i += nt) {
        const int tile = lv.tile_list[i >> 10], lr = (i >> 5) & 31, lc = i & 31;
        const int64_t idx = ((int64_t)(tile / lv.tiles_x) * TILE_H + lr) * lv.pitch + (int64_t)(tile % lv.tiles_x) * TILE_W + lc;
        xb[idx] = 0.f;  // every cell of the active tiles: the readers of x take whole pairs / quads of it
    }
    __syncthreads();
    for (int hs = 0; hs < 4 * sweeps; ++hs) {
        const int colour = hs < 2 * sweeps ? (hs & 1) : 1 - (hs & 1);  // 0 = red
        for (int i = t; i < cells; i += nt) {
            const int tile = lv.tile_list[i >> 10], lr = (i >> 5) & 31, lc = i & 31;
            const int64_t r = (int64_t)(tile / lv.tiles_x) * TILE_H + lr, c = (int64_t)(tile % lv.tiles_x) * TILE_W + lc;
            const int64_t idx = r * lv.pitch + c;
            if (((r + c) & 1) == colour && lv.umask[idx]) {
                // neighbours through the mask: the plane is only trusted at the unknowns of the current mask (cg.cu: stale_all)
                const float nb = ((lv.umask[idx - lv.pitch] ? xb[idx - lv.pitch] : 0.f) + (lv.umask[idx + lv.pitch] ? xb[idx + lv.pitch] : 0.f))
                    + ((lv.umask[idx - 1] ? xb[idx - 1] : 0.f) + (lv.umask[idx + 1] ? xb[idx + 1] : 0.f));
                const float w = lv.winv ? lv.winv[idx] : (fixed ? 0.25f : winv_of<false>(lv, r, c));
                xb[idx] = w * (bb[idx] + nb);
            }
        }
        __syncthreads();
    }
    if (DOT) {
        double acc = 0.0;
        for (int i = t; i < cells; i += nt) {
            const int tile = lv.tile_list[i >> 10], lr = (i >> 5) & 31, lc = i & 31;
            const int64_t idx = ((int64_t)(tile / lv.tiles_x) * TILE_H + lr) * lv.pitch + (int64_t)(tile % lv.tiles_x) * TILE_W + lc;
            if (lv.umask[idx])
                acc += (double)bb[idx] * (double)xb[idx];
        }
        for (int o = 16; o; o >>= 1)
            acc += __shfl_xor_sync(FULLW, acc, o);
        if ((t & 31) == 0 && acc != 0.0)
            atomicAdd(dot_out, acc);
    }
}

}  // namespace

// Levels [0, n_grid) of the tail are worked by the whole grid with a grid-wide barrier after each; the DEEP levels
// [n_grid, n) -- a handful of tiles each -- are worked per band by ONE CTA, whose warps only need __syncthreads between
// levels: a grid barrier costs ~4 us (one atomic per CTA and a round trip to L2 per poll), a CTA barrier nothing, and on a
// scene-sized problem (1697 x 1284: the reference's sample scene) ten of the seventeen phases of the tail are deep.
__global__ void __launch_bounds__(RW_THREADS, 3) k_rbw_tail(TailArgs A, BandScalars* __restrict__ scal, unsigned* __restrict__ barrier)
{
    __shared__ float s_coarsest[3 * CP * CP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wg = (int)blockIdx.x * RW_WARPS + warp;
    const int stride = (int)gridDim.x * RW_WARPS;
    unsigned epoch = 0;
    for (int l = 0; l < A.n_grid; ++l) {
        rbw_down_items<1>(A.L[l].lv, A.L[l + 1].lv, A.L[l].it, wg, stride, lane, A.L[l].b, A.L[l + 1].b, scal);
        grid_barrier(barrier, epoch);
    }
    for (int band = blockIdx.x; band < A.nbands; band += gridDim.x) {
        if (scal[band].done)
            continue;
        for (int l = A.n_grid; l + 1 < A.n; ++l) {
            const Level& lf = A.L[l].lv;
            const Level& lc = A.L[l + 1].lv;
            for (int ti = warp; ti < lf.n_tiles; ti += RW_WARPS) {
                const int yx = lf.tile_yx[ti];
                unsigned long long cm[4];
                quad_col_masks<DN_HR>(lf, yx >> 16, yx & 0xffff, lane < 30 ? lane % 10 : 9, cm);
                rbw_down_tile<1>(lf, lc.pitch, yx >> 16, yx & 0xffff, lane, ~0u, cm, A.L[l].b + (int64_t)band * lf.plane,
                    A.L[l + 1].b + (int64_t)band * lc.plane, 0);
            }
            __threadfence();
            __syncthreads();
        }
        {
            const TailLevel& C = A.L[A.n - 1];
            rbw_coarsest<false>(C.lv, A.fixed != 0, C.b + (int64_t)band * C.lv.plane, C.x + (int64_t)band * C.lv.plane, A.sweeps, nullptr,
                s_coarsest);
            __threadfence();
            __syncthreads();
        }
        for (int l = A.n - 2; l >= A.n_grid; --l) {
            const Level& lf = A.L[l].lv;
            const Level& lc = A.L[l + 1].lv;
            for (int ti = warp; ti < lf.n_tiles; ti += RW_WARPS) {
                const int yx = lf.tile_yx[ti];
                unsigned long long cm[4];
                quad_col_masks<UP_HR>(lf, yx >> 16, yx & 0xffff, lane < 30 ? lane % 10 : 9, cm);
                rbw_up_tile<1, false>(lf, lc.pitch, yx >> 16, yx & 0xffff, lane, ~0u, cm, A.L[l].b + (int64_t)band * lf.plane,
                    A.L[l + 1].x + (int64_t)band * lc.plane, A.L[l].x + (int64_t)band * lf.plane, 0);
            }
            __threadfence();
            __syncthreads();
        }
    }
    for (int l = A.n_grid - 1; l >= 0; --l) {
        grid_barrier(barrier, epoch);
        rbw_up_items<1, false>(A.L[l].lv, A.L[l + 1].lv, A.L[l].it, wg, stride, lane, A.L[l].b, A.L[l + 1].x, A.L[l].x, scal, nullptr);
    }
}

// a one-level hierarchy: the whole preconditioner is the relaxation on level 0 (tiny scenes)
__global__ void __launch_bounds__(RW_THREADS) k_rbw_coarsest_only(Level lv, int fixed, const float* __restrict__ b, float* __restrict__ x,
    BandScalars* __restrict__ scal, int slot, int sweeps)
{
    __shared__ float s_coarsest[3 * CP * CP];
    if (scal[blockIdx.x].done)
        return;
    rbw_coarsest<true>(lv, fixed != 0, b + (int64_t)blockIdx.x * lv.plane, x + (int64_t)blockIdx.x * lv.plane, sweeps,
        &scal[blockIdx.x].rz[slot], s_coarsest);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct RWLevel {
    Level lv;
    int64_t units;
    float* b;   // level 0: the float copy of the CG residual
    float* x;   // full plane (level 0: z)
};

int env_int(const char* name, int dflt)
{
    const char* e = std::getenv(name);
    return e && *e ? std::atoi(e) : dflt;
}

int wmode_of(const Level& lv) { return lv.winv ? 1 : (lv.fixed_diag ? 0 : 2); }

// band-major item order for levels whose band plane is larger than `SATFILL_RBW_BAND_MAJOR_MB` MB (default 256: level 0 of a 10980^2 tile)
bool band_major_for(const Level& lv)
{
    static const int mb = env_int("SATFILL_RBW_BAND_MAJOR_MB", 256);
    return mb >= 0 && (int64_t)lv.plane * (int64_t)sizeof(float) > (int64_t)mb << 20;
}

// resident CTAs per SM of a kernel (cached per kernel function)
template <typename K>
int ctas_per_sm(K kernel, size_t smem, int threads)
{
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1)
        n = 1;
    return n;
}

int launch_down_w(sa_ctx* ctx, const RWLevel& F, const RWLevel& C, int nb, const BandScalars* scal)
{
    if (F.lv.n_tiles == 0)
        return SA_OK;
    const int mode = wmode_of(F.lv);
    static int occ[3] = { 0, 0, 0 };
    if (!occ[mode])
        occ[mode] = mode == 0 ? ctas_per_sm(k_rbw_down<0>, 0, 32) : (mode == 1 ? ctas_per_sm(k_rbw_down<1>, 0, 32) : ctas_per_sm(k_rbw_down<2>, 0, 32));
    const int max_ctas = ctx->grid_sms * occ[mode];  // one warp each
    const Items it = make_items(F.lv.n_tiles, nb, max_ctas, band_major_for(F.lv));
    const unsigned grid = (unsigned)(it.count() < max_ctas ? it.count() : max_ctas);
    if (mode == 0)
        SA_LAUNCH(ctx, k_rbw_down<0>, grid, 32, 0, F.lv, C.lv, it, F.b, C.b, scal);
    else if (mode == 1)
        SA_LAUNCH(ctx, k_rbw_down<1>, grid, 32, 0, F.lv, C.lv, it, F.b, C.b, scal);
    else
        SA_LAUNCH(ctx, k_rbw_down<2>, grid, 32, 0, F.lv, C.lv, it, F.b, C.b, scal);
    return SA_OK;
}

template <bool DOT>
int launch_up_w(sa_ctx* ctx, const RWLevel& F, const RWLevel& C, int nb, BandScalars* scal, int slot)
{
    if (F.lv.n_tiles == 0)
        return SA_OK;
    const int mode = wmode_of(F.lv);
    const size_t smem = DOT ? sizeof(double) * (size_t)nb : 0;
    static int occ[3] = { 0, 0, 0 };
    static size_t occ_smem[3] = { 0, 0, 0 };
    if (!occ[mode] || occ_smem[mode] != smem) {
        occ[mode] = mode == 0 ? ctas_per_sm(k_rbw_up<0, DOT>, smem, 32) : (mode == 1 ? ctas_per_sm(k_rbw_up<1, DOT>, smem, 32) : ctas_per_sm(k_rbw_up<2, DOT>, smem, 32));
        occ_smem[mode] = smem;
    }
    const int max_ctas = ctx->grid_sms * occ[mode];  // one warp each
    const Items it = make_items(F.lv.n_tiles, nb, max_ctas, band_major_for(F.lv));
    const unsigned grid = (unsigned)(it.count() < max_ctas ? it.count() : max_ctas);
    if (mode == 0)
        SA_LAUNCH(ctx, (k_rbw_up<0, DOT>), grid, 32, smem, F.lv, C.lv, it, F.b, C.x, F.x, scal, slot);
    else if (mode == 1)
        SA_LAUNCH(ctx, (k_rbw_up<1, DOT>), grid, 32, smem, F.lv, C.lv, it, F.b, C.x, F.x, scal, slot);
    else
        SA_LAUNCH(ctx, (k_rbw_up<2, DOT>), grid, 32, smem, F.lv, C.lv, it, F.b, C.x, F.x, scal, slot);
    return SA_OK;
}

// levels [first, nl) of the cycle in one cooperative launch (first >= 1: every level of the tail carries a 1 / d plane)
int launch_tail(sa_ctx* ctx, const std::vector<RWLevel>& L, int first, int nb, BandScalars* scal, int sweeps, bool fixed)
{
    static int occ = 0;
    if (!occ)
        occ = ctas_per_sm(k_rbw_tail, 0, RW_THREADS);
    int tail_ctas = env_int("SATFILL_TAIL_CTAS_PER_SM", 2);
    if (tail_ctas > occ)
        tail_ctas = occ;
    const int max_ctas = ctx->grid_sms * tail_ctas;
    TailArgs A {};
    A.n = (int)L.size() - first;
    A.nbands = nb;
    A.sweeps = sweeps;
    A.fixed = fixed ? 1 : 0;
    int most = nb;
    const int deep_tiles = env_int("SATFILL_TAIL_DEEP_TILES", 8);
    A.n_grid = 0;
    for (int l = first; l + 1 < (int)L.size(); ++l)
        if (L[(size_t)l].lv.n_tiles > deep_tiles)
            A.n_grid = l - first + 1;
    for (int l = first; l < (int)L.size(); ++l) {
        TailLevel& T = A.L[l - first];
        T.lv = L[(size_t)l].lv;
        T.it = make_items(T.lv.n_tiles, nb, max_ctas * RW_WARPS, false);
        T.b = L[(size_t)l].b;
        T.x = L[(size_t)l].x;
        if (l - first < A.n_grid)
            most = std::max(most, (T.it.count() + RW_WARPS - 1) / RW_WARPS);
    }
    unsigned grid = (unsigned)std::min(most, max_ctas);
    if (!ctx->d_barrier)
        SA_CUDA(ctx, cudaMalloc(&ctx->d_barrier, sizeof(unsigned)));
    SA_CUDA(ctx, cudaMemsetAsync(ctx->d_barrier, 0, sizeof(unsigned), ctx->stream));
    BandScalars* sc = scal;
    unsigned* bar = ctx->d_barrier;
    void* args[] = { &A, &sc, &bar };
    SA_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)k_rbw_tail, dim3(grid), dim3(RW_THREADS), args, 0, ctx->stream));
    ctx->launches += 1;
    return SA_OK;
}

}  // namespace

// z (float, in s->z) = M^-1 r for every band that is not done; r.z is accumulated into rz[rz_slot].
// Storage: the level buffers allocated by mg.cu (double-sized) are used as float planes.
int apply_vcycle_rbw(sa_scene* s, const sa_options& o, KernelTimer& kt, int rz_slot, int live_bands)
{
    sa_ctx* ctx = s->ctx;
    const int nb = s->win_n(), b0 = s->band0;  // the band window (common.cuh): every base pointer starts at band b0
    std::vector<RWLevel> L;
    L.push_back({ fine_level(s), s->n_unknowns * live_bands, s->rb_rf(), s->rb_z() });
    for (sa_level_store& c : s->coarse) {
        if (c.lv.n_tiles == 0)
            break;
        L.push_back({ c.lv, c.n_unknowns * live_bands, (float*)c.b + c.lv.pitch + (int64_t)b0 * c.lv.plane,
            (float*)c.x + c.lv.pitch + (int64_t)b0 * c.lv.plane });
    }
    const int nl = (int)L.size();
    BandScalars* scal = s->scal + b0;
    const int coarse_sweeps = 16;
    const bool fixed = s->problem == SA_LAPLACE;
    // row decomposition (dist.cu): levels below dist_levels run on the rank's slice and exchange halo rows, the others
    // are replicated on every rank
    const bool dist = s->distributed && s->dist_planned && ctx->world > 1;
    const int dlv = dist ? s->dist_levels : 0;
    for (int l = 0; l < nl && l < dlv; ++l) {
        L[(size_t)l].lv = dist_level(s, l, L[(size_t)l].lv);
        // the profile counts what THIS rank processes (a windowed scene counted only its own unknowns on the split coarse levels)
        if (l == 0 || !s->dist_windowed)
            L[(size_t)l].units = (int64_t)((double)L[(size_t)l].units * s->dist_unit_frac);
    }
    if (nl == 1) {
        kt.begin(KC_SMOOTH, L[0].units);
        SA_LAUNCH(ctx, k_rbw_coarsest_only, nb, RW_THREADS, 0, L[0].lv, fixed ? 1 : 0, L[0].b, L[0].x, scal, rz_slot, coarse_sweeps);
        kt.end();
        SA_CUDA(ctx, cudaGetLastError());
        return SA_OK;
    }
    // the tail: from the first level (>= 1, replicated) whose work items no longer fill the GPU, down to the coarsest
    const int64_t tail_items = env_int("SATFILL_TAIL_ITEMS", 6144);
    int tail = nl - 1;
    while (tail - 1 >= 1 && tail - 1 >= dlv && (int64_t)L[(size_t)tail - 1].lv.n_tiles * nb <= tail_items)
        --tail;
    for (int l = 0; l < tail; ++l) {
        kt.begin(l == 0 ? KC_MG_DOWN : KC_MG_DOWN_COARSE, L[(size_t)l].units);
        SA_TRY(launch_down_w(ctx, L[(size_t)l], L[(size_t)l + 1], nb, scal));
        kt.end();
        if (l < dlv) {
            // the next level's descent reads its right-hand side 3 rows beyond the slice -- or, if that level is replicated,
            // everywhere (nothing of this level travels: the ascent recomputes the red half of the iterate from b, whose halo
            // rows it already holds)
            if (l + 1 < dlv)
                SA_TRY(dist_step(s, l + 1, DIST_VEC_RHS, L[(size_t)l + 1].b, 4, L[(size_t)l + 1].lv.pitch, L[(size_t)l + 1].lv.plane, 3, 3, -1,
                    0, -1));
            else
                SA_TRY(dist_gather(s, L[(size_t)l + 1].b, L[(size_t)l + 1].lv.pitch, L[(size_t)l + 1].lv.plane));
        }
    }
    {
        int64_t units = 0;
        for (int l = tail; l < nl; ++l)
            units += L[(size_t)l].units;
        kt.begin(KC_SMOOTH, units);
        SA_TRY(launch_tail(ctx, L, tail, nb, scal, coarse_sweeps, fixed));
        kt.end();
    }
    for (int l = tail - 1; l >= 0; --l) {
        kt.begin(l == 0 ? KC_MG_UP : KC_MG_UP_COARSE, L[(size_t)l].units);
        if (l == 0)
            SA_TRY((launch_up_w<true>(ctx, L[(size_t)l], L[(size_t)l + 1], nb, scal, rz_slot)));
        else
            SA_TRY((launch_up_w<false>(ctx, L[(size_t)l], L[(size_t)l + 1], nb, scal, 0)));
        kt.end();
        // the finer level's ascent interpolates from up to 2 coarse rows beyond.  (Level 0: CG's direction needs 1 row of z;
        // the caller exchanges it in one group with the all-reduce of r.z -- cg.cu.)
        if (l < dlv && l > 0)
            SA_TRY(dist_step(s, l, DIST_VEC_SOL, L[(size_t)l].x, 4, L[(size_t)l].lv.pitch, L[(size_t)l].lv.plane, 2, 2, -1, 0, -1));
    }
    SA_CUDA(ctx, cudaGetLastError());
    return SA_OK;
}

}  // namespace satfill
