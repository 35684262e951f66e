// Row decomposition of ONE linear system across the GPUs of a node (BASELINE.json configs[4]: a single
// 20000 x 20000 hole).  One process per GPU; NCCL over NVLink / NVSwitch for the two kinds of exchange the path has:
//
//   * halo rows: every kernel of the solver reads a vector with a halo of 1 to 3 rows (cg_strip.cu, mg_rb.cu); rows are
//     contiguous in memory, so a halo is one ncclSend / ncclRecv pair per neighbour and band, grouped;
//   * dot products: the CTA partial sums of a rank are added across ranks by one small ncclAllReduce per CG phase
//     (r.z, p.Ap, |r|^2), packed for all bands.
//
// Ownership is by whole tile rows, aligned so that the first `dist_levels` multigrid levels split at the same places
// (rank k owns tile rows [T_k 2^-l, T_{k+1} 2^-l) of level l); every coarser level is small (< 1 / 4^dist_levels of
// the fine grid) and is REPLICATED: its right-hand side is gathered once per cycle and every rank runs the same
// arithmetic on it, which removes all communication from the latency-bound bottom of the V-cycle.
//
// Every rank holds the whole mask (1 B / pixel) and indexes it itself, so tile lists, unknown counts and the coarse
// hierarchy need no communication; a rank's tiles are a contiguous slice of each raster-ordered tile list.  Planes are
// allocated at full size and addressed with global row indices -- only the rank's rows (+ halos) are ever touched --
// which keeps every kernel identical to the single-GPU path.
//
// NCCL is resolved at run time (dlopen of the libnccl.so.2 the process already has -- torch's -- or the system one), so
// libsatfill.so has no link-time dependency on it and single-GPU users never load it.
#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>

namespace satfill {

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl()
{
    static NcclApi api = [] {
        NcclApi a;
        // RTLD_NOLOAD first: reuse the library the process already loaded (torch.distributed's) so that there is one NCCL
        a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!a.handle)
            a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.handle)
            return a;
#define SA_SYM(name) *(void**)(&a.name) = dlsym(a.handle, "nccl" #name)
        SA_SYM(GetUniqueId);
        SA_SYM(CommInitRank);
        SA_SYM(CommDestroy);
        SA_SYM(AllReduce);
        SA_SYM(Broadcast);
        SA_SYM(Send);
        SA_SYM(Recv);
        SA_SYM(GroupStart);
        SA_SYM(GroupEnd);
        SA_SYM(GetErrorString);
#undef SA_SYM
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.Broadcast && a.Send && a.Recv
            && a.GroupStart && a.GroupEnd && a.GetErrorString;
        return a;
    }();
    return api;
}

#define SA_NCCL(ctx, expr)                                                                                        \
    do {                                                                                                          \
        ncclResult_t r__ = (expr);                                                                                \
        if (r__ != ncclSuccess)                                                                                   \
            return fail((ctx), SA_NCCL_ERROR, std::string(#expr) + ": " + nccl().GetErrorString(r__));            \
    } while (0)

// pack / unpack of the per-band scalars that one CG phase reduces across ranks
__global__ void k_pack(const BandScalars* __restrict__ scal, int nbands, int what, int slot, double* __restrict__ buf)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbands)
        return;
    const BandScalars& s = scal[b];
    double* o = buf + 3 * b;
    if (what == DIST_SETUP) {
        o[0] = s.bnorm2;
        o[1] = s.rr[0];
        o[2] = s.rz[0];
    } else if (what == DIST_RZ) {
        o[0] = s.rz[slot];
        o[1] = o[2] = 0.0;
    } else if (what == DIST_PQ) {
        o[0] = s.pq[slot];
        o[1] = o[2] = 0.0;
    } else {
        o[0] = s.rr[slot];
        o[1] = s.rz[slot];
        o[2] = 0.0;
    }
}

__global__ void k_unpack(BandScalars* __restrict__ scal, int nbands, int what, int slot, const double* __restrict__ buf,
    int clear_slot)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbands)
        return;
    BandScalars& s = scal[b];
    const double* o = buf + 3 * b;
    if (what == DIST_SETUP) {
        s.bnorm2 = o[0];
        s.rr[0] = o[1];
        s.rz[0] = o[2];
    } else if (what == DIST_RZ) {
        s.rz[slot] = o[0];
    } else if (what == DIST_PQ) {
        s.pq[slot] = o[0];
    } else {
        s.rr[slot] = o[0];
        if (what == DIST_RR_RZ)
            s.rz[slot] = o[1];
    }
    // the ring slot two iterations ahead: a rank whose slice holds no tile has no lead thread to recycle it
    if (clear_slot >= 0) {
        s.rz[clear_slot] = 0.0;
        s.rr[clear_slot] = 0.0;
        s.pq[clear_slot] = 0.0;
    }
}

}  // namespace

// ---- partition (host logic; also exported through the C-ABI so that it can be tested without a GPU) ---------------
// rows -> world + 1 row boundaries, multiples of 32 * 2^(levels - 1) rows (the last one is `rows` rounded up to it):
// blocks of aligned tile rows dealt out as evenly as possible, earlier ranks taking the remainder.
void dist_partition(int64_t rows, int world, int levels, int64_t* row_begin)
{
    const int64_t block = (int64_t)TILE_H << (levels > 1 ? levels - 1 : 0);
    const int64_t nblocks = (rows + block - 1) / block;
    int64_t at = 0;
    for (int k = 0; k < world; ++k) {
        row_begin[k] = at * block;
        at += nblocks / world + (k < nblocks % world ? 1 : 0);
    }
    row_begin[world] = nblocks * block;
}

// number of distributed multigrid levels for a scene: as many as keep at least two aligned blocks per rank, at most 4
int dist_choose_levels(int64_t rows, int world)
{
    int levels = 1;
    while (levels < 4) {
        int64_t block = (int64_t)TILE_H << levels;  // block size with one more level
        if ((rows + block - 1) / block < 2 * (int64_t)world)
            break;
        ++levels;
    }
    return levels;
}

int dist_unique_id(void* out128)
{
    if (!nccl().ok)
        return SA_NCCL_ERROR;
    ncclUniqueId id;
    if (nccl().GetUniqueId(&id) != ncclSuccess)
        return SA_NCCL_ERROR;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(out128, &id, 128);
    return SA_OK;
}

int dist_init(sa_ctx* ctx, const void* id128, int rank, int world)
{
    if (!nccl().ok)
        return fail(ctx, SA_NCCL_ERROR, "libnccl.so.2 could not be loaded");
    if (world < 1 || rank < 0 || rank >= world)
        return fail(ctx, SA_BAD_ARGUMENT, "dist_init: bad rank / world");
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    ncclComm_t comm = nullptr;
    SA_NCCL(ctx, nccl().CommInitRank(&comm, world, id, rank));
    ctx->comm = comm;
    ctx->rank = rank;
    ctx->world = world;
    SA_CUDA(ctx, cudaMalloc(&ctx->d_red, sizeof(double) * 3 * 64));
    return SA_OK;
}

void dist_shutdown(sa_ctx* ctx)
{
    if (ctx->comm && nccl().ok)
        nccl().CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
    cudaFree(ctx->d_red);
    ctx->d_red = nullptr;
}

// ---- per-scene plan ---------------------------------------------------------------------------------------------------
// Slices every level's tile list to the rank's tile rows.  Called after index_scene / build_hierarchy (the tile lists
// are raster ordered, so a slice is a contiguous range found by binary search on a host copy).
static int slice_tiles(sa_ctx* ctx, const int32_t* d_list, int n_tiles, int tiles_x, int ty_lo, int ty_hi, int* lo, int* hi)
{
    std::vector<int32_t> h((size_t)n_tiles);
    if (n_tiles)
        SA_CUDA(ctx, cudaMemcpyAsync(h.data(), d_list, sizeof(int32_t) * (size_t)n_tiles, cudaMemcpyDeviceToHost, ctx->stream));
    SA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *lo = (int)(std::lower_bound(h.begin(), h.end(), (int32_t)((int64_t)ty_lo * tiles_x)) - h.begin());
    *hi = (int)(std::lower_bound(h.begin(), h.end(), (int32_t)((int64_t)ty_hi * tiles_x)) - h.begin());
    return SA_OK;
}

int dist_plan_scene(sa_scene* s, bool multigrid)
{
    sa_ctx* ctx = s->ctx;
    const int world = ctx->world, rank = ctx->rank;
    int levels = 1;
    if (multigrid) {
        const int nl = 1 + (int)s->coarse.size();
        levels = dist_choose_levels(s->rows, world);
        // the coarsest level of the hierarchy is always replicated (its solver is one CTA per band)
        int usable = 0;
        for (int l = 1; l < nl; ++l)
            if (s->coarse[l - 1].lv.n_tiles > 0)
                usable = l;
        if (levels > usable)
            levels = usable;
        if (levels < 1)
            return fail(ctx, SA_BAD_ARGUMENT, "distributed solve: the scene is too small to be split by rows");
    }
    if ((s->rows + TILE_H - 1) / TILE_H < world)
        return fail(ctx, SA_BAD_ARGUMENT, "distributed solve: fewer tile rows than ranks");
    s->dist_levels = levels;
    s->dist_mg = multigrid;
    std::vector<int64_t> rb((size_t)world + 1);
    dist_partition(s->rows, world, levels, rb.data());
    s->dl.assign((size_t)levels, DistLevel {});
    for (int l = 0; l < levels; ++l) {
        DistLevel& d = s->dl[(size_t)l];
        const int64_t lrows = l == 0 ? s->rows : s->coarse[l - 1].lv.rows;
        const int lty = l == 0 ? s->tiles_y : s->coarse[l - 1].lv.tiles_y;
        const int ltx = l == 0 ? s->tiles_x : s->coarse[l - 1].lv.tiles_x;
        d.bounds.resize((size_t)world + 1);
        for (int k = 0; k <= world; ++k)
            d.bounds[(size_t)k] = std::min(rb[(size_t)k] >> l, (int64_t)lty * TILE_H);
        d.row_lo = d.bounds[(size_t)rank];
        d.row_hi = d.bounds[(size_t)rank + 1];
        d.rows = lrows;
        const int32_t* list = l == 0 ? s->tile_list : s->coarse[l - 1].tile_list;
        const int n = l == 0 ? s->n_active_tiles : s->coarse[l - 1].lv.n_tiles;
        SA_TRY(slice_tiles(ctx, list, n, ltx, (int)(d.row_lo / TILE_H), (int)(d.row_hi / TILE_H), &d.tile_lo, &d.tile_hi));
    }
    // the rows of the first replicated level that every rank produces (for the gather)
    s->dist_gather_rows.assign((size_t)world + 1, 0);
    if (multigrid) {
        const sa_level_store& R = s->coarse[(size_t)levels - 1];  // level `levels`
        for (int k = 0; k <= world; ++k)
            s->dist_gather_rows[(size_t)k] = std::min(rb[(size_t)k] >> levels, (int64_t)R.lv.tiles_y * TILE_H);
    }
    s->dist_planned = true;
    return SA_OK;
}

// The level as the rank sees it: its slice of the tile list.
Level dist_level(const sa_scene* s, int l, const Level& full)
{
    Level lv = full;
    if (!s->dist_planned || l >= s->dist_levels)
        return lv;
    const DistLevel& d = s->dl[(size_t)l];
    lv.tile_list = full.tile_list + d.tile_lo;
    lv.tile_yx = full.tile_yx + d.tile_lo;
    lv.n_tiles = d.tile_hi - d.tile_lo;
    return lv;
}

// ---- exchanges -----------------------------------------------------------------------------------------------------------
// Halo rows of a plane-shaped vector of level l (element (0, 0) of band 0 at `base`, `pitch` elements per row, `plane`
// elements per band): after the call the `above` rows before the rank's first row and the `below` rows after its last
// row hold the neighbours' values.
template <typename T>
int dist_halo(sa_scene* s, int l, T* base, int64_t pitch, int64_t plane, int above, int below)
{
    sa_ctx* ctx = s->ctx;
    if (!s->dist_planned || ctx->world == 1)
        return SA_OK;
    const DistLevel& d = s->dl[(size_t)l];
    const ncclComm_t comm = (ncclComm_t)ctx->comm;
    const ncclDataType_t dt = sizeof(T) == 8 ? ncclFloat64 : ncclFloat32;
    const int up = ctx->rank - 1, down = ctx->rank + 1;
    // ranks whose slice is empty at this level (row_lo == row_hi) neither own nor need rows; the partition hands out
    // whole blocks to the leading ranks, so empty slices only occur at the tail and never sit between two non-empty ones
    const bool mine = d.row_hi > d.row_lo;
    SA_NCCL(ctx, nccl().GroupStart());
    for (int b = 0; b < s->nbands; ++b) {
        T* p = base + (int64_t)b * plane;
        if (mine && up >= 0 && d.row_lo > 0) {
            // the rank above needs `below` rows from my top; I need `above` rows from its bottom
            SA_NCCL(ctx, nccl().Send(p + d.row_lo * pitch, (size_t)(below * pitch), dt, up, comm, ctx->stream));
            SA_NCCL(ctx, nccl().Recv(p + (d.row_lo - above) * pitch, (size_t)(above * pitch), dt, up, comm, ctx->stream));
        }
        if (mine && down < ctx->world && d.bounds[(size_t)down + 1] > d.bounds[(size_t)down]) {
            SA_NCCL(ctx, nccl().Send(p + (d.row_hi - above) * pitch, (size_t)(above * pitch), dt, down, comm, ctx->stream));
            SA_NCCL(ctx, nccl().Recv(p + d.row_hi * pitch, (size_t)(below * pitch), dt, down, comm, ctx->stream));
        }
    }
    SA_NCCL(ctx, nccl().GroupEnd());
    return SA_OK;
}
template int dist_halo<double>(sa_scene*, int, double*, int64_t, int64_t, int, int);
template int dist_halo<float>(sa_scene*, int, float*, int64_t, int64_t, int, int);

// Every rank's rows of the first replicated level -> all ranks (in place in the full plane).
int dist_gather(sa_scene* s, float* base, int64_t pitch, int64_t plane)
{
    sa_ctx* ctx = s->ctx;
    if (!s->dist_planned || ctx->world == 1)
        return SA_OK;
    const ncclComm_t comm = (ncclComm_t)ctx->comm;
    SA_NCCL(ctx, nccl().GroupStart());
    for (int b = 0; b < s->nbands; ++b)
        for (int k = 0; k < ctx->world; ++k) {
            int64_t lo = s->dist_gather_rows[(size_t)k], hi = s->dist_gather_rows[(size_t)k + 1];
            if (hi <= lo)
                continue;
            float* p = base + (int64_t)b * plane + lo * pitch;
            SA_NCCL(ctx, nccl().Broadcast(p, p, (size_t)((hi - lo) * pitch), ncclFloat32, k, comm, ctx->stream));
        }
    SA_NCCL(ctx, nccl().GroupEnd());
    return SA_OK;
}

// After a solve every rank holds the solution on its own rows only: broadcast them so that each rank has the whole band.
int dist_allgather_band(sa_scene* s, int band)
{
    sa_ctx* ctx = s->ctx;
    if (!s->dist_planned || ctx->world == 1)
        return SA_OK;
    const DistLevel& d = s->dl[0];
    SA_NCCL(ctx, nccl().GroupStart());
    for (int k = 0; k < ctx->world; ++k) {
        int64_t lo = d.bounds[(size_t)k], hi = std::min(d.bounds[(size_t)k + 1], s->rows_p);
        if (hi <= lo)
            continue;
        double* p = s->plane0(s->u, band) + lo * s->pitch;
        SA_NCCL(ctx, nccl().Broadcast(p, p, (size_t)((hi - lo) * s->pitch), ncclFloat64, k, (ncclComm_t)ctx->comm, ctx->stream));
    }
    SA_NCCL(ctx, nccl().GroupEnd());
    return SA_OK;
}

// NCCL groups nest: a pair of these around several exchanges turns them into one launch (17 NCCL launches per CG
// iteration are what separates the row-decomposed solve from linear scaling; grouped they are 10).
int dist_group_begin(sa_scene* s)
{
    sa_ctx* ctx = s->ctx;
    if (!s->dist_planned || ctx->world == 1)
        return SA_OK;
    SA_NCCL(ctx, nccl().GroupStart());
    return SA_OK;
}
int dist_group_end(sa_scene* s)
{
    sa_ctx* ctx = s->ctx;
    if (!s->dist_planned || ctx->world == 1)
        return SA_OK;
    SA_NCCL(ctx, nccl().GroupEnd());
    return SA_OK;
}

// Sum of the per-rank partial scalars of one CG phase, in three steps so that the all-reduce can share a group with the
// halo exchange of the same phase:  pack (kernel)  ->  issue (NCCL, may sit inside dist_group_begin / _end)  ->  unpack
// (kernel; clear_slot >= 0 also recycles that ring slot).
int dist_reduce_pack(sa_scene* s, int what, int slot)
{
    sa_ctx* ctx = s->ctx;
    if (!s->dist_planned || ctx->world == 1)
        return SA_OK;
    if (s->nbands > 64)
        return fail(ctx, SA_BAD_ARGUMENT, "distributed solve: at most 64 bands");
    SA_LAUNCH(ctx, k_pack, 1, 64, 0, s->scal, s->nbands, what, slot, ctx->d_red);
    return SA_OK;
}
int dist_reduce_issue(sa_scene* s)
{
    sa_ctx* ctx = s->ctx;
    if (!s->dist_planned || ctx->world == 1)
        return SA_OK;
    SA_NCCL(ctx, nccl().AllReduce(ctx->d_red, ctx->d_red, (size_t)(3 * s->nbands), ncclFloat64, ncclSum, (ncclComm_t)ctx->comm,
                     ctx->stream));
    return SA_OK;
}
int dist_reduce_unpack(sa_scene* s, int what, int slot, int clear_slot)
{
    sa_ctx* ctx = s->ctx;
    if (!s->dist_planned || ctx->world == 1)
        return SA_OK;
    SA_LAUNCH(ctx, k_unpack, 1, 64, 0, s->scal, s->nbands, what, slot, ctx->d_red, clear_slot);
    SA_CUDA(ctx, cudaGetLastError());
    return SA_OK;
}
int dist_reduce(sa_scene* s, int what, int slot, int clear_slot)
{
    SA_TRY(dist_reduce_pack(s, what, slot));
    SA_TRY(dist_reduce_issue(s));
    return dist_reduce_unpack(s, what, slot, clear_slot);
}

}  // namespace satfill
