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
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
        SA_SYM(AllGather);
        SA_SYM(Send);
        SA_SYM(Recv);
        SA_SYM(GroupStart);
        SA_SYM(GroupEnd);
        SA_SYM(GetErrorString);
#undef SA_SYM
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.Broadcast && a.AllGather && a.Send && a.Recv
            && a.GroupStart && a.GroupEnd && a.GetErrorString;
        return a;
    }();
    return api;
}

// Inside a GroupStart / GroupEnd pair: remember the first failure, keep going to the GroupEnd (a return from inside an open
// group would leave it open and the peers blocked in their half of the exchange).
#define SA_NCCL_IN_GROUP(first_error, expr)                                     \
    do {                                                                        \
        ncclResult_t r__ = (expr);                                              \
        if (r__ != ncclSuccess && (first_error) == ncclSuccess)                 \
            (first_error) = r__;                                                \
    } while (0)

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


// ---------------------------------------------------------------------------------------------------------------------------
// Peer-memory exchange (NVLink / NVSwitch, CUDA IPC): the exchanges of the iteration loop without a collective library.
//
// Inside the CG loop every exchange is tiny -- one to three halo rows per neighbour (3 to 240 KB) and 3 doubles per band --
// so what it costs is launch and protocol latency, not bandwidth: ten NCCL launches per iteration were half a millisecond
// against 0.9 ms of arithmetic at 8 GPUs.  Here every rank owns an ARENA in its HBM that all ranks of the node map (CUDA
// IPC handles, all-gathered once per plan over the NCCL communicator), and one exchange is two small kernels on the
// solver's stream:
//   k_peer_push   stores my boundary rows straight into the neighbours' arenas and my partial sums into every peer's
//                 arena (remote stores over NVLink), fences (system scope), and raises one flag per receiver to the epoch
//                 of this exchange;
//   k_peer_pull   waits (acquire, system scope) until the flags of my senders have reached the epoch, moves the rows from
//                 my arena into the halo rows of my plane and adds the partial sums of all ranks IN RANK ORDER -- every
//                 rank computes bit-identical sums, so all ranks take every decision of the solve together.
// No rank ever waits inside a push, so the pushes of all ranks always complete and the pulls cannot deadlock.  A slot of
// the arena is reused one CG iteration later; between two uses of a slot lies at least one all-rank reduction, which no
// rank can leave before every rank has pushed for it -- i.e. before every rank has, in stream order, drained the slot.
// The reduction slots themselves alternate between two buffers for the same reason.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int PEER_MAX_WORLD = 64, PEER_MAX_BANDS = 64, PEER_KINDS = 3 * 8;
constexpr size_t PEER_FLAG_BYTES = 4096;                                                 // halo flags [kind][2], then reduction flags [rank]
constexpr size_t PEER_AR_BYTES = sizeof(double) * 2 * PEER_MAX_WORLD * 3 * PEER_MAX_BANDS;  // [parity][rank][3 * band]
constexpr size_t PEER_HALO_OFF = ((PEER_FLAG_BYTES + PEER_AR_BYTES + 255) / 256) * 256;

struct PeerState {
    bool ok = false;
    char* arena = nullptr;  // my arena (device memory of this rank)
    size_t arena_bytes = 0;
    std::vector<char*> peer;  // every rank's arena as this process maps it ([rank] = arena)
    size_t slot_off[PEER_KINDS][2] = {};    // byte offset of the (kind, direction) staging slot in a rank's arena
    size_t slot_bytes[PEER_KINDS] = {};
    unsigned long long halo_epoch[PEER_KINDS] = {};
    unsigned long long ar_epoch = 0;
    unsigned* d_ticket = nullptr;
    size_t refused_bytes = 0;  // an arena of this size could not be shared (or sharing is switched off): stay on NCCL
};

struct PeerHalo {       // one exchange as the kernels see it
    char* base;         // element (0, 0) of band 0 of my plane
    int64_t row_bytes, plane_bytes;
    int nbands;
    int64_t send_up_row, send_up_rows;      // my first rows -> the rank above (its halo below)
    int64_t send_down_row, send_down_rows;  // my last rows -> the rank below (its halo above)
    int64_t recv_up_row, recv_up_rows;      // the rows above my slice <- the rank above
    int64_t recv_down_row, recv_down_rows;  // the rows below my slice <- the rank below
    char* up_slot;      // the slot in the arena of the rank above that receives from below
    char* down_slot;    // the slot in the arena of the rank below that receives from above
    char* my_from_up;   // my own slots
    char* my_from_down;
    unsigned long long* up_flag;    // flags to raise (in the neighbours' arenas) and to wait for (in mine)
    unsigned long long* down_flag;
    unsigned long long* my_up_flag;
    unsigned long long* my_down_flag;
    unsigned long long epoch;
};
struct PeerReduce {
    int on, what, slot, clear_slot, nbands, rank, world, parity;
    unsigned long long epoch;
    char* peer[PEER_MAX_WORLD];  // arenas
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// rows [row, row + rows) of every band, packed band after band
__device__ __forceinline__ void copy_rows(char* dst, bool dst_packed, const char* src, bool src_packed, int64_t rows, int64_t row_bytes,
    int64_t plane_bytes, int nbands, int t, int nt)
{
    const int64_t per_band = rows * row_bytes, n16 = per_band >> 4;  // rows are multiples of 128 bytes
    for (int b = 0; b < nbands; ++b) {
        const uint4* sp = reinterpret_cast<const uint4*>(src + (src_packed ? b * per_band : b * plane_bytes));
        uint4* dp = reinterpret_cast<uint4*>(dst + (dst_packed ? b * per_band : b * plane_bytes));
        for (int64_t i = t; i < n16; i += nt)
            dp[i] = sp[i];
    }
}

__global__ void __launch_bounds__(256) k_peer_push(PeerHalo H, PeerReduce R, const BandScalars* __restrict__ scal, unsigned* __restrict__ ticket)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    if (H.send_up_rows > 0)
        copy_rows(H.up_slot, true, H.base + H.send_up_row * H.row_bytes, false, H.send_up_rows, H.row_bytes, H.plane_bytes, H.nbands, t, nt);
    if (H.send_down_rows > 0)
        copy_rows(H.down_slot, true, H.base + H.send_down_row * H.row_bytes, false, H.send_down_rows, H.row_bytes, H.plane_bytes, H.nbands, t, nt);
    if (R.on && blockIdx.x == 0) {
        // my partial sums into slot [parity][my rank] of every rank's arena (my own included)
        for (int i = threadIdx.x; i < R.world * R.nbands; i += blockDim.x) {
            const int peer = i / R.nbands, b = i - peer * R.nbands;
            const BandScalars& sc = scal[b];
            double v0, v1 = 0.0, v2 = 0.0;
            if (R.what == DIST_SETUP)
                v0 = sc.bnorm2, v1 = sc.rr[0], v2 = sc.rz[0];
            else if (R.what == DIST_RZ)
                v0 = sc.rz[R.slot];
            else if (R.what == DIST_PQ)
                v0 = sc.pq[R.slot];
            else
                v0 = sc.rr[R.slot], v1 = sc.rz[R.slot];
            double* o = reinterpret_cast<double*>(R.peer[peer] + PEER_FLAG_BYTES) + ((size_t)(R.parity * PEER_MAX_WORLD + R.rank) * PEER_MAX_BANDS + b) * 3;
            o[0] = v0, o[1] = v1, o[2] = v2;
        }
    }
    // the last CTA to get here raises the flags: everything every CTA stored is then visible system-wide
    __threadfence_system();
    __syncthreads();
    __shared__ unsigned last;
    if (threadIdx.x == 0)
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        __threadfence_system();
        if (threadIdx.x == 0) {
            *ticket = 0;
            if (H.send_up_rows > 0)
                st_release_sys(H.up_flag, H.epoch);
            if (H.send_down_rows > 0)
                st_release_sys(H.down_flag, H.epoch);
        }
        if (R.on)
            for (int peer = threadIdx.x; peer < R.world; peer += blockDim.x)
                st_release_sys(reinterpret_cast<unsigned long long*>(R.peer[peer] + 2048) + R.rank, R.epoch);
    }
}

__global__ void __launch_bounds__(256) k_peer_pull(PeerHalo H, PeerReduce R, BandScalars* __restrict__ scal)
{
    if (threadIdx.x == 0) {
        if (H.recv_up_rows > 0)
            while (ld_acquire_sys(H.my_up_flag) < H.epoch) { }
        if (H.recv_down_rows > 0)
            while (ld_acquire_sys(H.my_down_flag) < H.epoch) { }
    }
    if (R.on && blockIdx.x == 0 && threadIdx.x < R.world)
        while (ld_acquire_sys(reinterpret_cast<const unsigned long long*>(R.peer[R.rank] + 2048) + threadIdx.x) < R.epoch) { }
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    if (H.recv_up_rows > 0)
        copy_rows(H.base + H.recv_up_row * H.row_bytes, false, H.my_from_up, true, H.recv_up_rows, H.row_bytes, H.plane_bytes, H.nbands, t, nt);
    if (H.recv_down_rows > 0)
        copy_rows(H.base + H.recv_down_row * H.row_bytes, false, H.my_from_down, true, H.recv_down_rows, H.row_bytes, H.plane_bytes, H.nbands, t, nt);
    if (R.on && blockIdx.x == 0) {
        for (int b = threadIdx.x; b < R.nbands; b += blockDim.x) {
            const double* base = reinterpret_cast<const double*>(R.peer[R.rank] + PEER_FLAG_BYTES) + (size_t)R.parity * PEER_MAX_WORLD * PEER_MAX_BANDS * 3;
            double v0 = 0.0, v1 = 0.0, v2 = 0.0;
            for (int r = 0; r < R.world; ++r) {  // rank order: the same sum, bit for bit, on every rank
                const double* o = base + ((size_t)r * PEER_MAX_BANDS + b) * 3;
                v0 += o[0], v1 += o[1], v2 += o[2];
            }
            BandScalars& sc = scal[b];
            if (R.what == DIST_SETUP)
                sc.bnorm2 = v0, sc.rr[0] = v1, sc.rz[0] = v2;
            else if (R.what == DIST_RZ)
                sc.rz[R.slot] = v0;
            else if (R.what == DIST_PQ)
                sc.pq[R.slot] = v0;
            else {
                sc.rr[R.slot] = v0;
                if (R.what == DIST_RR_RZ)
                    sc.rz[R.slot] = v1;
            }
            if (R.clear_slot >= 0) {  // see k_unpack
                sc.rz[R.clear_slot] = 0.0;
                sc.rr[R.clear_slot] = 0.0;
                sc.pq[R.clear_slot] = 0.0;
            }
        }
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

static void peer_release(sa_ctx* ctx)
{
    PeerState* P = (PeerState*)ctx->peer;
    if (!P)
        return;
    for (size_t r = 0; r < P->peer.size(); ++r)
        if (P->peer[r] && (int)r != ctx->rank)
            cudaIpcCloseMemHandle(P->peer[r]);
    cudaFree(P->arena);
    cudaFree(P->d_ticket);
    delete P;
    ctx->peer = nullptr;
}

void dist_shutdown(sa_ctx* ctx)
{
    if (ctx->comm)
        cudaStreamSynchronize(ctx->stream);
    peer_release(ctx);
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

static int peer_plan(sa_scene* s);

// extents of level l of the hierarchy of a rows-high grid (mg.cu: alloc_hierarchy)
static void level_rows(int64_t rows, int l, int64_t* lrows, int* ltiles_y)
{
    for (int i = 0; i < l; ++i)
        rows = (rows + 1) / 2;
    *lrows = rows;
    int64_t rp = round_up(rows, TILE_H);
    *ltiles_y = (int)((rp == 0 ? TILE_H : rp) / TILE_H);
}

// row boundaries of every split level and of the first replicated one
static void fill_bounds(sa_scene* s, int levels)
{
    const int world = s->ctx->world, rank = s->ctx->rank;
    std::vector<int64_t> rb((size_t)world + 1);
    dist_partition(s->rows, world, levels, rb.data());
    s->dl.assign((size_t)levels, DistLevel {});
    for (int l = 0; l < levels; ++l) {
        DistLevel& d = s->dl[(size_t)l];
        int64_t lrows;
        int lty;
        level_rows(s->rows, l, &lrows, &lty);
        d.bounds.resize((size_t)world + 1);
        for (int k = 0; k <= world; ++k)
            d.bounds[(size_t)k] = std::min(rb[(size_t)k] >> l, (int64_t)lty * TILE_H);
        d.row_lo = d.bounds[(size_t)rank];
        d.row_hi = d.bounds[(size_t)rank + 1];
        d.rows = lrows;
    }
    // the rows of the first replicated level that every rank produces (for the gather)
    s->dist_gather_rows.assign((size_t)world + 1, 0);
    int64_t rrows;
    int rty;
    level_rows(s->rows, levels, &rrows, &rty);
    for (int k = 0; k <= world; ++k)
        s->dist_gather_rows[(size_t)k] = std::min(rb[(size_t)k] >> levels, (int64_t)rty * TILE_H);
}

// Called BEFORE the mask is indexed: fixes the rows of every split level this rank owns, so that index_scene and
// build_hierarchy only have to look at those rows (+ one tile row either side) -- the set-up of the row-decomposed solve
// then scales with the ranks instead of being repeated by every rank on the whole mask.  The number of split levels is the
// one dist_choose_levels wants; dist_plan_scene checks it against the hierarchy that came out and, for scenes too small
// for it, asks for the unwindowed path (SA_RETRY_UNWINDOWED).
int dist_prepare_window(sa_scene* s, bool multigrid)
{
    sa_ctx* ctx = s->ctx;
    s->dist_windowed = false;
    if (!s->distributed || ctx->world <= 1 || s->dist_no_window || std::getenv("SATFILL_DIST_NO_WINDOW"))
        return SA_OK;
    if ((s->rows + TILE_H - 1) / TILE_H < ctx->world)
        return SA_OK;  // dist_plan_scene reports it
    const int levels = multigrid ? dist_choose_levels(s->rows, ctx->world) : 1;
    fill_bounds(s, levels);
    s->dist_levels = levels;
    s->dist_mg_window = multigrid;
    s->dist_windowed = true;
    return SA_OK;
}

int dist_plan_scene(sa_scene* s, bool multigrid)
{
    sa_ctx* ctx = s->ctx;
    const int world = ctx->world;
    if ((s->rows + TILE_H - 1) / TILE_H < world)
        return fail(ctx, SA_BAD_ARGUMENT, "distributed solve: fewer tile rows than ranks");
    const int nl = 1 + (int)s->coarse.size();
    if (s->dist_windowed) {
        if (s->dist_mg_window != multigrid)
            return SA_RETRY_UNWINDOWED;
        if (multigrid) {
            // the hierarchy must reach below the split levels: a replicated level with tiles (the replicated levels are the
            // same on every rank, so every rank comes to the same conclusion; a split level's own list may well be empty)
            bool deep_enough = false;
            for (int l = s->dist_levels; l < nl; ++l)
                if (s->coarse[(size_t)l - 1].lv.n_tiles > 0)
                    deep_enough = true;
            if (!deep_enough)
                return SA_RETRY_UNWINDOWED;
        }
        // the tile lists ARE the rank's slices
        for (int l = 0; l < s->dist_levels; ++l) {
            s->dl[(size_t)l].tile_lo = 0;
            s->dl[(size_t)l].tile_hi = l == 0 ? s->n_active_tiles : s->coarse[(size_t)l - 1].lv.n_tiles;
        }
        s->dist_mg = multigrid;
        s->dist_planned = true;
        s->dist_unit_frac = s->rows > 0 ? (double)(std::min(s->dl[0].row_hi, s->rows) - std::min(s->dl[0].row_lo, s->rows)) / (double)s->rows : 1.0;
        SA_TRY(peer_plan(s));
        return SA_OK;
    }
    int levels = 1;
    if (multigrid) {
        levels = dist_choose_levels(s->rows, world);
        // the coarsest level of the hierarchy is always replicated (its solver is one CTA per band)
        int usable = 0;
        for (int l = 1; l < nl; ++l)
            if (s->coarse[(size_t)l - 1].lv.n_tiles > 0)
                usable = l;
        if (levels > usable)
            levels = usable;
        if (levels < 1)
            return fail(ctx, SA_BAD_ARGUMENT, "distributed solve: the scene is too small to be split by rows");
    }
    s->dist_levels = levels;
    s->dist_mg = multigrid;
    fill_bounds(s, levels);
    for (int l = 0; l < levels; ++l) {
        DistLevel& d = s->dl[(size_t)l];
        const int ltx = l == 0 ? s->tiles_x : s->coarse[(size_t)l - 1].lv.tiles_x;
        const int32_t* list = l == 0 ? s->tile_list : s->coarse[(size_t)l - 1].tile_list;
        const int n = l == 0 ? s->n_active_tiles : s->coarse[(size_t)l - 1].lv.n_tiles;
        SA_TRY(slice_tiles(ctx, list, n, ltx, (int)(d.row_lo / TILE_H), (int)(d.row_hi / TILE_H), &d.tile_lo, &d.tile_hi));
    }
    s->dist_planned = true;
    s->dist_unit_frac = s->rows > 0 ? (double)(std::min(s->dl[0].row_hi, s->rows) - std::min(s->dl[0].row_lo, s->rows)) / (double)s->rows : 1.0;
    SA_TRY(peer_plan(s));
    return SA_OK;
}

// ---- peer-memory arena: layout for a scene, allocation, exchange of the IPC handles ----------------------------------------
// Collective: every rank calls it with the same scene shape (dist_plan_scene).  Falls back to NCCL for the whole
// communicator (P->ok = false on every rank) when any rank cannot map a peer's arena, or when SATFILL_DIST_NCCL_ONLY is set.
static int peer_plan(sa_scene* s)
{
    sa_ctx* ctx = s->ctx;
    const int world = ctx->world, rank = ctx->rank;
    if (world > PEER_MAX_WORLD || s->nbands > PEER_MAX_BANDS || s->dist_levels > PEER_KINDS / 3)
        return SA_OK;  // stays on NCCL
    PeerState* P = (PeerState*)ctx->peer;
    if (!P) {
        P = new PeerState();
        ctx->peer = P;
        SA_CUDA(ctx, cudaMalloc(&P->d_ticket, sizeof(unsigned)));
        SA_CUDA(ctx, cudaMemsetAsync(P->d_ticket, 0, sizeof(unsigned), ctx->stream));
    }
    // staging slots: (level, vector) x (from above, from below), up to 3 rows of doubles per band
    size_t off = PEER_HALO_OFF;
    for (int k = 0; k < PEER_KINDS; ++k)
        P->slot_bytes[k] = 0;
    for (int l = 0; l < s->dist_levels; ++l) {
        const int64_t pitch = l == 0 ? s->pitch : s->coarse[(size_t)l - 1].lv.pitch;
        for (int v = 0; v < 3; ++v) {
            const int k = l * 3 + v;
            P->slot_bytes[k] = (size_t)(3 * pitch * (int64_t)sizeof(double)) * (size_t)s->nbands;
            for (int d = 0; d < 2; ++d) {
                P->slot_off[k][d] = off;
                off += (P->slot_bytes[k] + 255) / 256 * 256;
            }
        }
    }
    if (off <= P->arena_bytes || off <= P->refused_bytes)
        return SA_OK;  // the arena of an earlier plan is large enough (every rank sees the same history of shapes)
    const bool forced_off = std::getenv("SATFILL_DIST_NCCL_ONLY") != nullptr;
    // a new, larger arena: allocate, all-gather the handles over the NCCL communicator, map the peers, then drop the old one
    // (the all-gather orders this after everything any rank still had in flight on the old arena)
    char* fresh = nullptr;
    const size_t bytes = off + (off >> 2);
    SA_CUDA(ctx, cudaMalloc(&fresh, bytes));
    SA_CUDA(ctx, cudaMemsetAsync(fresh, 0, bytes, ctx->stream));
    cudaIpcMemHandle_t mine {};
    int ok = forced_off ? 0 : (cudaIpcGetMemHandle(&mine, fresh) == cudaSuccess ? 1 : 0);
    cudaGetLastError();
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    char* d_h = nullptr;
    SA_CUDA(ctx, cudaMalloc(&d_h, (size_t)world * 64 + sizeof(int)));
    SA_CUDA(ctx, cudaMemcpyAsync(d_h + (size_t)rank * 64, &mine, 64, cudaMemcpyHostToDevice, ctx->stream));
    SA_NCCL(ctx, nccl().AllGather(d_h + (size_t)rank * 64, d_h, 64, ncclChar, (ncclComm_t)ctx->comm, ctx->stream));
    std::vector<cudaIpcMemHandle_t> all((size_t)world);
    SA_CUDA(ctx, cudaMemcpyAsync(all.data(), d_h, (size_t)world * 64, cudaMemcpyDeviceToHost, ctx->stream));
    SA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<char*> mapped((size_t)world, nullptr);
    mapped[(size_t)rank] = fresh;
    for (int r = 0; r < world && ok; ++r) {
        if (r == rank)
            continue;
        void* q = nullptr;
        if (cudaIpcOpenMemHandle(&q, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            ok = 0;
            cudaGetLastError();
        }
        mapped[(size_t)r] = (char*)q;
    }
    // every rank must agree
    int* d_ok = reinterpret_cast<int*>(d_h + (size_t)world * 64);
    SA_CUDA(ctx, cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    SA_NCCL(ctx, nccl().AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, (ncclComm_t)ctx->comm, ctx->stream));
    SA_CUDA(ctx, cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_h);
    // the old arena and its mappings
    for (size_t r = 0; r < P->peer.size(); ++r)
        if (P->peer[r] && (int)r != rank)
            cudaIpcCloseMemHandle(P->peer[r]);
    cudaFree(P->arena);
    if (!ok) {
        for (int r = 0; r < world; ++r)
            if (r != rank && mapped[(size_t)r])
                cudaIpcCloseMemHandle(mapped[(size_t)r]);
        cudaFree(fresh);
        P->arena = nullptr;
        P->arena_bytes = 0;
        P->peer.clear();
        P->ok = false;
        P->refused_bytes = off;
        return SA_OK;
    }
    P->arena = fresh;
    P->arena_bytes = bytes;
    P->peer = mapped;
    for (int k = 0; k < PEER_KINDS; ++k)
        P->halo_epoch[k] = 0;
    P->ar_epoch = 0;
    P->ok = true;
    return SA_OK;
}

int dist_uses_peer_memory(const sa_ctx* ctx)
{
    const PeerState* P = (const PeerState*)ctx->peer;
    return P && P->ok ? 1 : 0;
}

// One exchange step of the row-decomposed solve on the solver's stream: the halo rows of one vector of level `level`
// (base != nullptr: `above` rows wanted above the slice, `below` rows below it; vec 0 = right-hand side / residual,
// 1 = iterate / correction, 2 = search direction) and / or the sum over the ranks of one group of per-band scalars
// (what >= 0: DistWhat; clear_slot as in dist_reduce_unpack).  Over peer memory when the arena is mapped, else over NCCL.
int dist_step(sa_scene* s, int level, int vec, void* base, int elem_bytes, int64_t pitch, int64_t plane, int above, int below, int what,
    int slot, int clear_slot)
{
    sa_ctx* ctx = s->ctx;
    if (!s->dist_planned || ctx->world == 1)
        return SA_OK;
    PeerState* P = (PeerState*)ctx->peer;
    if (!P || !P->ok) {
        if (what >= 0)
            SA_TRY(dist_reduce_pack(s, what, slot));
        SA_TRY(dist_group_begin(s));
        int st = SA_OK;
        if (what >= 0)
            st = dist_reduce_issue(s);
        if (st == SA_OK && base)
            st = elem_bytes == 8 ? dist_halo<double>(s, level, (double*)base, pitch, plane, above, below)
                                 : dist_halo<float>(s, level, (float*)base, pitch, plane, above, below);
        int st2 = dist_group_end(s);  // always closes the group, also on an error inside it
        if (st != SA_OK)
            return st;
        SA_TRY(st2);
        if (what >= 0)
            SA_TRY(dist_reduce_unpack(s, what, slot, clear_slot));
        return SA_OK;
    }
    const int rank = ctx->rank, world = ctx->world;
    const DistLevel& d = s->dl[(size_t)level];
    const bool mine = d.row_hi > d.row_lo;
    const int up = rank - 1, down = rank + 1;
    const bool has_up = base && mine && up >= 0 && d.row_lo > 0;
    const bool has_down = base && mine && down < world && d.bounds[(size_t)down + 1] > d.bounds[(size_t)down];
    const int kind = level * 3 + vec;
    PeerHalo H {};
    H.base = (char*)base;
    H.row_bytes = pitch * elem_bytes;
    H.plane_bytes = plane * elem_bytes;
    H.nbands = s->nbands;
    if (base) {
        if ((size_t)(std::max(above, below) * H.row_bytes) * (size_t)s->nbands > P->slot_bytes[kind])
            return fail(ctx, SA_BAD_ARGUMENT, "distributed solve: halo rows do not fit their staging slot");
        H.epoch = ++P->halo_epoch[kind];
        unsigned long long* my_flags = reinterpret_cast<unsigned long long*>(P->arena);
        H.my_up_flag = my_flags + kind * 2 + 0;
        H.my_down_flag = my_flags + kind * 2 + 1;
        H.my_from_up = P->arena + P->slot_off[kind][0];
        H.my_from_down = P->arena + P->slot_off[kind][1];
        if (has_up) {  // the rank above receives my first rows as its halo BELOW
            H.send_up_row = d.row_lo, H.send_up_rows = below;
            H.recv_up_row = d.row_lo - above, H.recv_up_rows = above;
            H.up_slot = P->peer[(size_t)up] + P->slot_off[kind][1];
            H.up_flag = reinterpret_cast<unsigned long long*>(P->peer[(size_t)up]) + kind * 2 + 1;
        }
        if (has_down) {  // the rank below receives my last rows as its halo ABOVE
            H.send_down_row = d.row_hi - above, H.send_down_rows = above;
            H.recv_down_row = d.row_hi, H.recv_down_rows = below;
            H.down_slot = P->peer[(size_t)down] + P->slot_off[kind][0];
            H.down_flag = reinterpret_cast<unsigned long long*>(P->peer[(size_t)down]) + kind * 2 + 0;
        }
    }
    PeerReduce R {};
    R.on = what >= 0 ? 1 : 0;
    if (R.on) {
        R.what = what, R.slot = slot, R.clear_slot = clear_slot, R.nbands = s->nbands, R.rank = rank, R.world = world;
        R.epoch = ++P->ar_epoch;
        R.parity = (int)(R.epoch & 1);
        for (int r = 0; r < world; ++r)
            R.peer[r] = P->peer[(size_t)r];
    }
    const int64_t bytes = (int64_t)std::max(above, below) * H.row_bytes * s->nbands;
    const unsigned grid = !base ? 1u : (bytes > (1 << 20) ? 32u : (bytes > (64 << 10) ? 8u : 2u));
    SA_LAUNCH(ctx, k_peer_push, grid, 256, 0, H, R, s->scal, P->d_ticket);
    SA_LAUNCH(ctx, k_peer_pull, grid, 256, 0, H, R, s->scal);
    SA_CUDA(ctx, cudaGetLastError());
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
    ncclResult_t bad = ncclSuccess;
    SA_NCCL(ctx, nccl().GroupStart());
    for (int b = 0; b < s->nbands; ++b) {
        T* p = base + (int64_t)b * plane;
        if (mine && up >= 0 && d.row_lo > 0) {
            // the rank above needs `below` rows from my top; I need `above` rows from its bottom
            SA_NCCL_IN_GROUP(bad, nccl().Send(p + d.row_lo * pitch, (size_t)(below * pitch), dt, up, comm, ctx->stream));
            SA_NCCL_IN_GROUP(bad, nccl().Recv(p + (d.row_lo - above) * pitch, (size_t)(above * pitch), dt, up, comm, ctx->stream));
        }
        if (mine && down < ctx->world && d.bounds[(size_t)down + 1] > d.bounds[(size_t)down]) {
            SA_NCCL_IN_GROUP(bad, nccl().Send(p + (d.row_hi - above) * pitch, (size_t)(above * pitch), dt, down, comm, ctx->stream));
            SA_NCCL_IN_GROUP(bad, nccl().Recv(p + d.row_hi * pitch, (size_t)(below * pitch), dt, down, comm, ctx->stream));
        }
    }
    SA_NCCL_IN_GROUP(bad, nccl().GroupEnd());
    SA_NCCL(ctx, bad);
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
    ncclResult_t bad = ncclSuccess;
    SA_NCCL(ctx, nccl().GroupStart());
    for (int b = 0; b < s->nbands; ++b)
        for (int k = 0; k < ctx->world; ++k) {
            int64_t lo = s->dist_gather_rows[(size_t)k], hi = s->dist_gather_rows[(size_t)k + 1];
            if (hi <= lo)
                continue;
            float* p = base + (int64_t)b * plane + lo * pitch;
            SA_NCCL_IN_GROUP(bad, nccl().Broadcast(p, p, (size_t)((hi - lo) * pitch), ncclFloat32, k, comm, ctx->stream));
        }
    SA_NCCL_IN_GROUP(bad, nccl().GroupEnd());
    SA_NCCL(ctx, bad);
    return SA_OK;
}

// After a solve every rank holds the solution on its own rows only: broadcast them so that each rank has the whole band.
int dist_allgather_band(sa_scene* s, int band)
{
    sa_ctx* ctx = s->ctx;
    if (!s->dist_planned || ctx->world == 1)
        return SA_OK;
    const DistLevel& d = s->dl[0];
    ncclResult_t bad = ncclSuccess;
    SA_NCCL(ctx, nccl().GroupStart());
    for (int k = 0; k < ctx->world; ++k) {
        int64_t lo = d.bounds[(size_t)k], hi = std::min(d.bounds[(size_t)k + 1], s->rows_p);
        if (hi <= lo)
            continue;
        double* p = s->plane0(s->u, band) + lo * s->pitch;
        SA_NCCL_IN_GROUP(bad, nccl().Broadcast(p, p, (size_t)((hi - lo) * s->pitch), ncclFloat64, k, (ncclComm_t)ctx->comm, ctx->stream));
    }
    SA_NCCL_IN_GROUP(bad, nccl().GroupEnd());
    SA_NCCL(ctx, bad);
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
