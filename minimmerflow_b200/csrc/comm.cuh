// comm.cuh -- multi-GPU plumbing: one process per GPU, NCCL point-to-point halo exchange over
// NVLink and a scalar max all-reduce.  Replaces the reference's MPI layer
// (src/communications.{hpp,tpp,cpp}; call sites src/main.cpp:307-328, 391-395, 425-430, 461-466,
// 497-502).  NCCL is resolved with dlopen at mmf_comm_init time, so single-GPU use of the library
// has no NCCL dependency and a process that already loaded torch's NCCL shares that copy.
#pragma once

#include "mmf_common.cuh"
#include "uniform_path.cuh"

#include <dlfcn.h>
#include <nccl.h>

namespace mmf {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

inline NcclApi &nccl_api() { static NcclApi api; return api; }

static int nccl_load(mmf_ctx *ctx)
{
    NcclApi &a = nccl_api();
    if (a.lib) return MMF_OK;
    const char *names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char *n : names) {
        a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.lib) break;
    }
    if (!a.lib) return fail(ctx, MMF_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define MMF_NCCL_SYM(field, name)                                                                  \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, name));                              \
    if (!a.field) return fail(ctx, MMF_ERR_NCCL, "libnccl: missing symbol %s", name);
    MMF_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    MMF_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    MMF_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    MMF_NCCL_SYM(Send, "ncclSend")
    MMF_NCCL_SYM(Recv, "ncclRecv")
    MMF_NCCL_SYM(GroupStart, "ncclGroupStart")
    MMF_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    MMF_NCCL_SYM(AllReduce, "ncclAllReduce")
    MMF_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef MMF_NCCL_SYM
    return MMF_OK;
}

#define MMF_NCCL(ctx, call)                                                                        \
    do {                                                                                           \
        ncclResult_t r__ = (call);                                                                 \
        if (r__ != ncclSuccess) {                                                                  \
            return mmf::fail((ctx), MMF_ERR_NCCL, "%s failed at %s:%d: %s", #call, __FILE__,       \
                             __LINE__, mmf::nccl_api().GetErrorString(r__));                       \
        }                                                                                          \
    } while (0)

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, n_ranks = 1;
    // generic path ghost lists (GhostCommunicator's exchange lists, src/communications.cpp:621-630)
    std::vector<int> nbr;
    std::vector<int64_t> send_off, recv_off; // per neighbour offsets into the id arrays
    int32_t *d_send_ids = nullptr, *d_recv_ids = nullptr;
    double *d_send_buf = nullptr, *d_recv_buf = nullptr;
};

static int comm_unique_id(void *out)
{
    if (!out) return fail(nullptr, MMF_ERR_INVALID, "mmf_comm_unique_id: null buffer");
    int rc = nccl_load(nullptr);
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    MMF_NCCL(nullptr, nccl_api().GetUniqueId(&id));
    memcpy(out, &id, sizeof id);
    return MMF_OK;
}

static int comm_init(mmf_ctx *ctx, int rank, int n_ranks, const void *unique_id)
{
    if (ctx->comm) return fail(ctx, MMF_ERR_STATE, "mmf_comm_init: communicator already initialised");
    if (!unique_id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(ctx, MMF_ERR_INVALID, "mmf_comm_init: bad arguments");
    int rc = nccl_load(ctx);
    if (rc) return rc;
    Comm *c = new Comm();
    c->rank = rank;
    c->n_ranks = n_ranks;
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof id);
    ncclResult_t r = nccl_api().CommInitRank(&c->comm, n_ranks, id, rank);
    if (r != ncclSuccess) {
        delete c;
        return fail(ctx, MMF_ERR_NCCL, "ncclCommInitRank failed: %s", nccl_api().GetErrorString(r));
    }
    ctx->comm = c;
    return MMF_OK;
}

static void comm_ipc_close(mmf_ctx *ctx);

static void comm_destroy(mmf_ctx *ctx)
{
    if (!ctx->comm) return;
    if (ctx->uni && ctx->uni->p2p && ctx->comm->comm && ctx->d_ctl) {
        // With peer stores a neighbour may still be writing its last layers into this rank's ghost
        // cells: tearing a partitioned handle down is collective (like freeing a communicator) -- every
        // rank first drains its own streams, then all ranks meet in an all-reduce before any mapping goes.
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->comm_stream);
        if (nccl_api().AllReduce(&ctx->d_ctl->est_max, &ctx->d_ctl->est_max, 1, ncclDouble, ncclMax, ctx->comm->comm,
                                 ctx->stream) == ncclSuccess) {
            cudaStreamSynchronize(ctx->stream);
        }
    }
    comm_ipc_close(ctx);
    if (ctx->comm->comm) nccl_api().CommDestroy(ctx->comm->comm);
    delete ctx->comm;
    ctx->comm = nullptr;
}

// MPI_Allreduce(MPI_IN_PLACE, &maxEig, 1, MPI_DOUBLE, MPI_MAX) of src/main.cpp:393, in stream
int comm_allreduce_max_enqueue(mmf_ctx *ctx, double *d_value, int count)
{
    Comm *c = ctx->comm;
    MMF_NCCL(ctx, nccl_api().AllReduce(d_value, d_value, (size_t) count, ncclDouble, ncclMax, c->comm, ctx->stream));
    return MMF_OK;
}

// ---- generic path: list-driven pack / unpack ----------------------------------------------------
// PiercedStorageBufferStreamer<double>::write / read (src/communications.tpp:164-199): for each
// listed id, nFields consecutive values.

__global__ void __launch_bounds__(256) ghost_pack_kernel(const int32_t *__restrict__ ids, int64_t n, int64_t stride,
                                                         const double *__restrict__ S, double *__restrict__ buf)
{
    const int64_t q = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int64_t c = ids[q];
#pragma unroll
    for (int k = 0; k < NF; ++k) buf[q * NF + k] = S[k * stride + c];
}

__global__ void __launch_bounds__(256) ghost_unpack_kernel(const int32_t *__restrict__ ids, int64_t n, int64_t stride,
                                                           const double *__restrict__ buf, double *__restrict__ S)
{
    const int64_t q = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int64_t c = ids[q];
#pragma unroll
    for (int k = 0; k < NF; ++k) S[k * stride + c] = buf[q * NF + k];
}

static int comm_set_ghost_lists(mmf_ctx *ctx, int n_nbr, const int32_t *ranks, const int64_t *send_off,
                                const int64_t *send_ids, const int64_t *recv_off, const int64_t *recv_ids)
{
    Comm *c = ctx->comm;
    if (!c) return fail(ctx, MMF_ERR_STATE, "mmf_comm_set_ghost_lists: call mmf_comm_init first");
    if (ctx->path != MMF_PATH_GENERIC) return fail(ctx, MMF_ERR_STATE, "ghost lists apply to the generic path");
    if (n_nbr < 0 || (n_nbr > 0 && (!ranks || !send_off || !send_ids || !recv_off || !recv_ids))) {
        return fail(ctx, MMF_ERR_INVALID, "mmf_comm_set_ghost_lists: bad arguments");
    }
    c->nbr.assign(ranks, ranks + n_nbr);
    c->send_off.assign(send_off, send_off + n_nbr + 1);
    c->recv_off.assign(recv_off, recv_off + n_nbr + 1);
    const int64_t ns = n_nbr ? send_off[n_nbr] : 0, nr = n_nbr ? recv_off[n_nbr] : 0;
    std::vector<int32_t> s(ns), r(nr);
    for (int64_t i = 0; i < ns; ++i) {
        if (send_ids[i] < 0 || send_ids[i] >= ctx->n_cells) return fail(ctx, MMF_ERR_INVALID, "send id out of range");
        s[i] = (int32_t) send_ids[i];
    }
    for (int64_t i = 0; i < nr; ++i) {
        if (recv_ids[i] < 0 || recv_ids[i] >= ctx->n_cells) return fail(ctx, MMF_ERR_INVALID, "recv id out of range");
        r[i] = (int32_t) recv_ids[i];
    }
    int rc;
    if ((rc = dev_upload(ctx, &c->d_send_ids, s))) return rc;
    if ((rc = dev_upload(ctx, &c->d_recv_ids, r))) return rc;
    if ((rc = dev_alloc(ctx, &c->d_send_buf, (size_t) ns * NF))) return rc;
    if ((rc = dev_alloc(ctx, &c->d_recv_buf, (size_t) nr * NF))) return rc;
    return MMF_OK;
}

static int comm_generic_exchange_enqueue(mmf_ctx *ctx, double *S)
{
    Comm *c = ctx->comm;
    const int n_nbr = (int) c->nbr.size();
    if (n_nbr == 0) return MMF_OK;
    const int64_t ns = c->send_off[n_nbr], nr = c->recv_off[n_nbr];
    if (ns > 0) {
        ghost_pack_kernel<<<grid_for(ns, 256), 256, 0, ctx->stream>>>(c->d_send_ids, ns, ctx->gm.stride, S, c->d_send_buf);
        MMF_LAUNCH_CHECK(ctx);
    }
    MMF_NCCL(ctx, nccl_api().GroupStart());
    for (int q = 0; q < n_nbr; ++q) {
        const int64_t s0 = c->send_off[q], s1 = c->send_off[q + 1], r0 = c->recv_off[q], r1 = c->recv_off[q + 1];
        if (r1 > r0) MMF_NCCL(ctx, nccl_api().Recv(c->d_recv_buf + r0 * NF, (size_t) (r1 - r0) * NF, ncclDouble, c->nbr[q], c->comm, ctx->stream));
        if (s1 > s0) MMF_NCCL(ctx, nccl_api().Send(c->d_send_buf + s0 * NF, (size_t) (s1 - s0) * NF, ncclDouble, c->nbr[q], c->comm, ctx->stream));
    }
    MMF_NCCL(ctx, nccl_api().GroupEnd());
    if (nr > 0) {
        ghost_unpack_kernel<<<grid_for(nr, 256), 256, 0, ctx->stream>>>(c->d_recv_ids, nr, ctx->gm.stride, c->d_recv_buf, S);
        MMF_LAUNCH_CHECK(ctx);
    }
    return MMF_OK;
}

// ---- uniform path: face layers of the box -------------------------------------------------------

// layer = -1 / n (ghost) or 0 / n-1 (interior boundary) along `axis`; buf layout [k][b][a]
__global__ void __launch_bounds__(256) uniform_layer_kernel(const UniformGeom g, double *__restrict__ S,
                                                            double *__restrict__ buf, int axis, int layer, int to_buf)
{
    const int na = (axis == 0) ? g.ny : g.nx;
    const int nb = (axis == 2) ? g.ny : g.nz;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (a >= na || b >= nb) return;
    int i, j, k;
    if (axis == 0)      { i = layer; j = a; k = b; }
    else if (axis == 1) { i = a; j = layer; k = b; }
    else                { i = a; j = b; k = layer; }
    const long long o = uoff(g, i, j, k);
    const long long per = (long long) na * nb;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        if (to_buf) buf[f * per + (long long) b * na + a] = S[f * g.fs + o];
        else        S[f * g.fs + o] = buf[f * per + (long long) b * na + a];
    }
}

static int comm_set_box_neighbours(mmf_ctx *ctx, const int32_t ranks[6])
{
    Comm *c = ctx->comm;
    if (!c) return fail(ctx, MMF_ERR_STATE, "mmf_comm_set_box_neighbours: call mmf_comm_init first");
    if (ctx->path != MMF_PATH_UNIFORM) return fail(ctx, MMF_ERR_STATE, "box neighbours apply to the uniform path");
    UniformPath *u = ctx->uni;
    const UniformGeom &g = u->g;
    if (u->bodies) { // the wall-cell passes know no partition sides: a box with bodies is a single-GPU box
        return fail(ctx, MMF_ERR_INVALID, "a uniform box with bodies cannot be partitioned: describe the ranks' meshes "
                    "for the generic path (MMF_FLAG_FORCE_GENERIC or MMF_UNIFORM_BODIES=0)");
    }
    for (int s = 0; s < 6; ++s) {
        const bool partition_side = (g.bc[s] == -2);
        if (partition_side != (ranks[s] >= 0)) {
            return fail(ctx, MMF_ERR_INVALID, "side %d: neighbour rank given iff the side is a partition boundary", s);
        }
        if (ranks[s] >= c->n_ranks) return fail(ctx, MMF_ERR_INVALID, "side %d: neighbour rank out of range", s);
        u->nbr_rank[s] = ranks[s];
        if (ranks[s] >= 0 && !u->send_buf[s]) {
            const int axis = s >> 1;
            const size_t n = (size_t) NF * ((axis == 0) ? g.ny : g.nx) * ((axis == 2) ? g.ny : g.nz);
            int rc;
            if ((rc = dev_alloc(ctx, &u->send_buf[s], n))) return rc;
            if ((rc = dev_alloc(ctx, &u->recv_buf[s], n))) return rc;
        }
    }
    return MMF_OK;
}

// ---- uniform path: halo exchange by direct peer stores (NVLink P2P through CUDA IPC) -----------
// One kernel copies every partition-side boundary layer of S straight into the ghost layer of the
// neighbour's copy of the same array (no pack buffer, no NCCL rendezvous, no unpack); its last block
// then raises this rank's arrival counter in every neighbour.  The consumer is a one-warp kernel in
// front of the next stage that waits for the counters of all its neighbours.
struct PushArgs {
    double *dst[6];               // neighbour's array (same field layout, same box dimensions), or -- x sides
                                  // with compact_x -- the neighbour's compact ghost columns for this array
    int compact_x;                // x layers go to [field][k+1][j+1] columns (XGhost) instead of the padded array
    long long xg_fs;
    unsigned long long *flag[6];  // neighbour's arrival counter for the side this rank sits on
    int side_of_slot[6];          // blockIdx.z -> side: only sides that have a neighbour are launched
};

__global__ void __launch_bounds__(256) uniform_push_kernel(const UniformGeom g, const double *__restrict__ S, const PushArgs args,
                                                           unsigned int *__restrict__ done, unsigned long long seq)
{
    const int side = args.side_of_slot[blockIdx.z];
    {
        const int axis = side >> 1;
        const bool hi = side & 1;
        const int na = (axis == 0) ? g.ny : g.nx, nb = (axis == 2) ? g.ny : g.nz;
        const int n_ax = (axis == 0) ? g.nx : (axis == 1) ? g.ny : g.nz;
        const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
        if (a < na && b < nb) {
            const int own = hi ? n_ax - 1 : 0, ghost = hi ? -1 : n_ax; // my layer -> the neighbour's ghost layer
            long long so, go;
            if (axis == 0)      { so = uoff(g, own, a, b); go = uoff(g, ghost, a, b); }
            else if (axis == 1) { so = uoff(g, a, own, b); go = uoff(g, a, ghost, b); }
            else                { so = uoff(g, a, b, own); go = uoff(g, a, b, ghost); }
            double *d = args.dst[side];
            if (axis == 0 && args.compact_x) { // contiguous in j: coalesced stores over NVLink
                const long long co = (long long) (b + 1) * (g.ny + 2) + (a + 1);
#pragma unroll
                for (int f = 0; f < NF; ++f) d[f * args.xg_fs + co] = S[f * g.fs + so];
            } else {
#pragma unroll
                for (int f = 0; f < NF; ++f) d[f * g.fs + go] = S[f * g.fs + so];
            }
        }
    }
    // last block out raises the arrival counters: every block's stores are fenced at system scope
    // before it counts itself done, so they are performed before the counters move
    __shared__ bool last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
        last = (atomicAdd(done, 1u) == total - 1);
    }
    __syncthreads();
    if (last && threadIdx.x < 6 && args.flag[threadIdx.x]) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(args.flag[threadIdx.x]) = seq;
        if (threadIdx.x == 0) *done = 0;
    } else if (last && threadIdx.x == 0) {
        *done = 0;
    }
}

__global__ void uniform_wait_kernel(const unsigned long long *flags, unsigned int side_mask, unsigned long long seq,
                                    double *timeouts, unsigned long long timeout_ns)
{
    const int s = threadIdx.x;
    if (s < 6 && ((side_mask >> s) & 1u)) {
        if (!halo_spin(flags + s, seq, timeout_ns)) atomicAdd(timeouts, 1.0);
    }
    __threadfence_system();
}

constexpr int DMA_SEQ_RING = 1024; // pinned sequence numbers the arrival counters are copied from (comm_uniform_dma_push)
constexpr int IPC_HANDLES = 5; // U, Wa, Wb, arrival counters, compact x ghost columns
constexpr size_t IPC_BLOB_BYTES = 512; // = MMF_IPC_BLOB_BYTES
static_assert(IPC_HANDLES * sizeof(cudaIpcMemHandle_t) <= IPC_BLOB_BYTES, "IPC blob too small");

static int comm_ipc_export(mmf_ctx *ctx, void *out)
{
    if (ctx->path != MMF_PATH_UNIFORM) return fail(ctx, MMF_ERR_STATE, "mmf_comm_ipc_export: uniform path only");
    UniformPath *u = ctx->uni;
    int rc;
    if (!u->flags) {
        if ((rc = dev_alloc(ctx, &u->flags, 8))) return rc;
        if ((rc = dev_alloc(ctx, &u->push_count, 1))) return rc;
        MMF_CUDA(ctx, cudaMemset(u->flags, 0, 8 * sizeof(unsigned long long)));
        MMF_CUDA(ctx, cudaMemset(u->push_count, 0, sizeof(unsigned int)));
        // compact x ghost columns: [side][array][field][nz+2][ny+2], initialised with a benign state
        u->xg_fs = ((long long) (u->g.ny + 2) * (u->g.nz + 2) + 15) / 16 * 16;
        if ((rc = dev_alloc(ctx, &u->xghost, (size_t) 6 * NF * u->xg_fs))) return rc;
        for (int q = 0; q < 6; ++q) {
            fill_benign_kernel<<<grid_for(u->xg_fs, 256), 256, 0, ctx->stream>>>(u->xghost + (size_t) q * NF * u->xg_fs, u->xg_fs);
            MMF_LAUNCH_CHECK(ctx);
        }
        MMF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    memset(out, 0, IPC_BLOB_BYTES);
    cudaIpcMemHandle_t *h = static_cast<cudaIpcMemHandle_t *>(out);
    for (int a = 0; a < 3; ++a) MMF_CUDA(ctx, cudaIpcGetMemHandle(&h[a], u->arr[a]));
    MMF_CUDA(ctx, cudaIpcGetMemHandle(&h[3], u->flags));
    MMF_CUDA(ctx, cudaIpcGetMemHandle(&h[4], u->xghost));
    return MMF_OK;
}

static int comm_ipc_import(mmf_ctx *ctx, const void *all_ranks)
{
    Comm *c = ctx->comm;
    if (!c) return fail(ctx, MMF_ERR_STATE, "mmf_comm_ipc_import: call mmf_comm_init first");
    if (ctx->path != MMF_PATH_UNIFORM) return fail(ctx, MMF_ERR_STATE, "mmf_comm_ipc_import: uniform path only");
    UniformPath *u = ctx->uni;
    if (!u->flags) return fail(ctx, MMF_ERR_STATE, "mmf_comm_ipc_import: call mmf_comm_ipc_export first");
    const char *blob = static_cast<const char *>(all_ranks);
    for (int s = 0; s < 6; ++s) {
        const int r = u->nbr_rank[s];
        if (r < 0) continue;
        if (r == c->rank) return fail(ctx, MMF_ERR_INVALID, "mmf_comm_ipc_import: a box cannot neighbour itself");
        const cudaIpcMemHandle_t *h = reinterpret_cast<const cudaIpcMemHandle_t *>(blob + (size_t) r * IPC_BLOB_BYTES);
        // the same neighbour may sit on several sides only in degenerate grids; open per side anyway
        for (int a = 0; a < IPC_HANDLES; ++a) {
            void *p = nullptr;
            bool reused = false;
            for (int s2 = 0; s2 < s && !reused; ++s2) {
                if (u->nbr_rank[s2] == r) { p = u->ipc_opened[s2][a]; reused = true; }
            }
            if (!reused) {
                cudaIpcMemHandle_t hh;
                memcpy(&hh, &h[a], sizeof hh);
                MMF_CUDA(ctx, cudaIpcOpenMemHandle(&p, hh, cudaIpcMemLazyEnablePeerAccess));
                u->ipc_opened[s][a] = p;
            }
            if (a < 3)       u->peer_arr[s][a] = static_cast<double *>(p);
            else if (a == 3) u->peer_flags[s] = static_cast<unsigned long long *>(p) + (s ^ 1); // I sit on its opposite side
            else             u->peer_xghost[s] = static_cast<double *>(p);
        }
    }
    u->p2p = true;
    // the copy engines carry the exchange when no x side is a partition side (MMF_DMA_PUSH=0: the push kernel)
    u->dma_push = u->nbr_rank[0] < 0 && u->nbr_rank[1] < 0 && !(getenv("MMF_DMA_PUSH") && atoi(getenv("MMF_DMA_PUSH")) == 0);
    if (u->dma_push && !u->seq_ring) {
        MMF_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void **>(&u->seq_ring), DMA_SEQ_RING * sizeof(unsigned long long), cudaHostAllocDefault));
    }
    // an x partition side keeps its ghosts in compact columns only the rotate form reads: the stages switch to it (same
    // warps per CTA: 'h' at 12 warps updates 11 rows per tile, 'r' 10 -- tiles, z chunks and estimate buffers follow)
    if (uniform_use_xghost(ctx)) {
        bool changed = false;
        for (int st = 0; st < 4; ++st) {
            if (u->shape[st].form != 'r') { u->shape[st].form = 'r'; changed = true; }
        }
        if (changed) { if (int rc = uniform_setup_shapes(ctx)) return rc; }
    }
    return uniform_build_tile_orders(ctx);
}

static void comm_ipc_close(mmf_ctx *ctx)
{
    UniformPath *u = ctx->uni;
    if (!u) return;
    if (u->seq_ring) { cudaFreeHost(u->seq_ring); u->seq_ring = nullptr; }
    u->dma_push = false;
    for (int s = 0; s < 6; ++s)
        for (int a = 0; a < IPC_HANDLES; ++a)
            if (u->ipc_opened[s][a]) { cudaIpcCloseMemHandle(u->ipc_opened[s][a]); u->ipc_opened[s][a] = nullptr; }
    u->p2p = false;
}

// ---- uniform path: the same exchange by the COPY ENGINES ------------------------------------------
// A stage kernel holds every SM of the GPU (one CTA per SM, all registers), so a push KERNEL launched next to the
// following stage only gets in when the first CTAs retire (70 us at 256^3), and then competes with them; at 8 GPUs
// the boundary CTAs of the neighbours sat waiting for it (profiles/r02e_multigpu.md).  The copy engines need no SM:
// a z layer is one 2-D peer copy (a whole padded plane per field, fields one pitch apart), a y layer one 2-D copy
// per field (a padded row per plane), and the arrival counter follows in stream order as an 8-byte copy out of a
// pinned ring of sequence numbers.  The padded rows / planes carry their own ghost cells along: those land in edge
// and corner ghosts of the neighbour, which no interface touches.  x layers (8-byte elements a row apart) stay with
// the push kernel: the default decompositions have no partition side across x.
static int comm_uniform_dma_push(mmf_ctx *ctx, double *S, int a, bool defer_wait)
{
    UniformPath *u = ctx->uni;
    const UniformGeom &g = u->g;
    unsigned int mask = 0;
    for (int s = 2; s < 6; ++s) mask |= (u->nbr_rank[s] >= 0) ? (1u << s) : 0u;
    if (!mask) return MMF_OK;
    const unsigned long long seq = ++u->xchg_seq;
    const bool async = defer_wait && u->push_async;
    cudaStream_t ps = ctx->stream;
    if (async) { // next to the following stage
        MMF_CUDA(ctx, cudaEventRecord(u->ev_stage, ctx->stream));
        MMF_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, u->ev_stage, 0));
        ps = ctx->comm_stream;
    }
    const size_t row = (size_t) g.px * sizeof(double), plane = row * g.py, field = (size_t) g.fs * sizeof(double);
    for (int s = 2; s < 6; ++s) {
        if (u->nbr_rank[s] < 0) continue;
        const int axis = s >> 1;
        const bool hi = s & 1;
        const int n_ax = (axis == 1) ? g.ny : g.nz;
        const int own = hi ? n_ax - 1 : 0, ghost = hi ? -1 : n_ax; // my layer -> the neighbour's ghost layer
        const char *src = reinterpret_cast<const char *>(S);
        char *dst = reinterpret_cast<char *>(u->peer_arr[s][a]);
        if (axis == 2) {
            MMF_CUDA(ctx, cudaMemcpy2DAsync(dst + (size_t) (ghost + 1) * plane, field, src + (size_t) (own + 1) * plane, field,
                                            plane, NF, cudaMemcpyDefault, ps));
        } else if (field == plane * g.pz) { // the fields follow each other without a gap: one copy for all of them
            MMF_CUDA(ctx, cudaMemcpy2DAsync(dst + (size_t) (ghost + 1) * row, plane, src + (size_t) (own + 1) * row, plane,
                                            row, (size_t) g.pz * NF, cudaMemcpyDefault, ps));
        } else {
            for (int f = 0; f < NF; ++f) {
                MMF_CUDA(ctx, cudaMemcpy2DAsync(dst + f * field + (size_t) (ghost + 1) * row, plane,
                                                src + f * field + (size_t) (own + 1) * row, plane, row, g.pz, cudaMemcpyDefault, ps));
            }
        }
    }
    unsigned long long *slot = u->seq_ring + (seq % DMA_SEQ_RING);
    *slot = seq; // (the host runs at most a few steps ahead of the device: mmf_run synchronises every 8 steps)
    for (int s = 2; s < 6; ++s) {
        if (u->nbr_rank[s] < 0) continue;
        MMF_CUDA(ctx, cudaMemcpyAsync(u->peer_flags[s], slot, sizeof(unsigned long long), cudaMemcpyDefault, ps));
    }
    if (async) {
        MMF_CUDA(ctx, cudaEventRecord(u->ev_push[a], ctx->comm_stream));
        u->push_pending[a] = true;
    }
    u->arr_seq[a] = seq;
    if (!defer_wait) { // nobody downstream waits in-kernel: block the stream until all neighbours have delivered
        uniform_wait_kernel<<<1, 32, 0, ctx->stream>>>(u->flags, mask, seq, &ctx->d_ctl->halo_timeouts, halo_timeout_ns());
        MMF_LAUNCH_CHECK(ctx);
    }
    return MMF_OK;
}

static int comm_uniform_push_enqueue(mmf_ctx *ctx, double *S, bool defer_wait)
{
    UniformPath *u = ctx->uni;
    const UniformGeom &g = u->g;
    int a = -1;
    for (int q = 0; q < 3; ++q) if (S == u->arr[q]) a = q;
    if (a < 0) return fail(ctx, MMF_ERR_STATE, "peer exchange is defined for the U / W arrays only");
    if (u->dma_push) return comm_uniform_dma_push(ctx, S, a, defer_wait);
    PushArgs args{};
    unsigned int mask = 0;
    int n_slots = 0;
    // compact x ghost columns only where the readers know about them (v5 stage kernels with in-kernel waits)
    args.compact_x = uniform_use_xghost(ctx) ? 1 : 0;
    args.xg_fs = u->xg_fs;
    for (int s = 0; s < 6; ++s) {
        if (u->nbr_rank[s] < 0) continue;
        args.dst[s] = u->peer_arr[s][a];
        if (s < 2 && args.compact_x) { // my -x layer is the +x ghost column of the neighbour, and vice versa
            args.dst[s] = u->peer_xghost[s] + (size_t) (((s ^ 1) & 1) * 3 + a) * NF * u->xg_fs;
        }
        args.flag[s] = u->peer_flags[s];
        args.side_of_slot[n_slots++] = s;
        mask |= 1u << s;
    }
    if (!mask) return MMF_OK;
    const unsigned long long seq = ++u->xchg_seq;
    const int na = std::max(g.nx, g.ny), nb = std::max(g.ny, g.nz);
    dim3 grid((na + 255) / 256, nb, n_slots);
    const bool async = defer_wait && u->push_async;
    cudaStream_t ps = ctx->stream;
    if (!async) { // pushes share one completion counter: never two in flight
        for (int q = 0; q < 3; ++q) {
            if (u->push_pending[q]) {
                MMF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, u->ev_push[q], 0));
                u->push_pending[q] = false;
            }
        }
    }
    if (async) { // next to the interior tiles of the following stage
        MMF_CUDA(ctx, cudaEventRecord(u->ev_stage, ctx->stream));
        MMF_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, u->ev_stage, 0));
        ps = ctx->comm_stream;
    }
    uniform_push_kernel<<<grid, 256, 0, ps>>>(g, S, args, u->push_count, seq);
    MMF_LAUNCH_CHECK(ctx);
    if (async) {
        MMF_CUDA(ctx, cudaEventRecord(u->ev_push[a], ctx->comm_stream));
        u->push_pending[a] = true;
    }
    u->arr_seq[a] = seq;
    if (!defer_wait) { // nobody downstream waits in-kernel: block the stream until all neighbours have delivered
        uniform_wait_kernel<<<1, 32, 0, ctx->stream>>>(u->flags, mask, seq, &ctx->d_ctl->halo_timeouts, halo_timeout_ns());
        MMF_LAUNCH_CHECK(ctx);
    }
    return MMF_OK;
}

int comm_uniform_exchange_enqueue(mmf_ctx *ctx, double *S, bool defer_wait)
{
    Comm *c = ctx->comm;
    UniformPath *u = ctx->uni;
    if (u->p2p) return comm_uniform_push_enqueue(ctx, S, defer_wait);
    const UniformGeom &g = u->g;
    bool any = false;
    for (int s = 0; s < 6; ++s) {
        if (u->nbr_rank[s] < 0) continue;
        any = true;
        const int axis = s >> 1;
        const int n_ax = (axis == 0) ? g.nx : (axis == 1) ? g.ny : g.nz;
        const int na = (axis == 0) ? g.ny : g.nx, nb = (axis == 2) ? g.ny : g.nz;
        dim3 grid((na + 255) / 256, nb);
        uniform_layer_kernel<<<grid, 256, 0, ctx->stream>>>(g, S, u->send_buf[s], axis, (s & 1) ? n_ax - 1 : 0, 1);
        MMF_LAUNCH_CHECK(ctx);
    }
    if (!any) return MMF_OK;
    MMF_NCCL(ctx, nccl_api().GroupStart());
    for (int s = 0; s < 6; ++s) {
        if (u->nbr_rank[s] < 0) continue;
        const int axis = s >> 1;
        const size_t n = (size_t) NF * ((axis == 0) ? g.ny : g.nx) * ((axis == 2) ? g.ny : g.nz);
        MMF_NCCL(ctx, nccl_api().Recv(u->recv_buf[s], n, ncclDouble, u->nbr_rank[s], c->comm, ctx->stream));
        MMF_NCCL(ctx, nccl_api().Send(u->send_buf[s], n, ncclDouble, u->nbr_rank[s], c->comm, ctx->stream));
    }
    MMF_NCCL(ctx, nccl_api().GroupEnd());
    for (int s = 0; s < 6; ++s) {
        if (u->nbr_rank[s] < 0) continue;
        const int axis = s >> 1;
        const int n_ax = (axis == 0) ? g.nx : (axis == 1) ? g.ny : g.nz;
        const int na = (axis == 0) ? g.ny : g.nx, nb = (axis == 2) ? g.ny : g.nz;
        dim3 grid((na + 255) / 256, nb);
        uniform_layer_kernel<<<grid, 256, 0, ctx->stream>>>(g, S, u->recv_buf[s], axis, (s & 1) ? n_ax : -1, 0);
        MMF_LAUNCH_CHECK(ctx);
    }
    return MMF_OK;
}

static int comm_exchange_enqueue(mmf_ctx *ctx, int field)
{
    if (ctx->path == MMF_PATH_UNIFORM) return comm_uniform_exchange_enqueue(ctx, uniform_field_ptr(ctx, field), false);
    return comm_generic_exchange_enqueue(ctx, ctx->fields[field]);
}

} // namespace mmf
