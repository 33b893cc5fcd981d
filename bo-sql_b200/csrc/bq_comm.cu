// bq_comm.cu — the collectives of the multi-GPU operator layer, native: NCCL over NVLink / NVSwitch, one process per GPU.
//
// The reference is a single process (no counterpart).  The operator layer (bo-sql_b200/host/exchange.cpp) decides WHAT
// crosses NVLink; this file is HOW: every entry point enqueues on the context's stream, so collectives are ordered with
// the kernels that produce and consume their buffers and nothing synchronises the host except the small host exchange
// (row counts, statistics, outcome words), which has to return values to the CPU.
//
//   bq_comm_unique_id / bq_comm_init   rank 0 makes the id, the host passes its 128 bytes to every rank by any channel
//   bq_comm_all_gather                 equal shards: ncclAllGather
//   bq_comm_all_gather_v               ragged shards: grouped ncclBroadcast, one per contributing rank, straight into place
//   bq_comm_all_to_all_v               grouped ncclSend / ncclRecv
//   bq_comm_all_reduce_sum_u32         ncclAllReduce (in-switch reduction when NVLS is available)
//   bq_comm_host_all_gather_i64        n int64 per rank through a device staging buffer, one stream synchronisation
#include "bq_common.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

namespace bq {

// NCCL is bound at run time, when the first communicator is made: the library has no link-time dependency on libnccl, so
// loading it never decides which NCCL a host process ends up with (PyTorch bundles its own, newer libnccl.so.2; an older
// copy loaded first would leave torch's symbols unresolved).  dlopen by soname returns the copy the process already has.
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
};

static const NcclApi& nccl() {
    static NcclApi api;
    static bool loaded = false;
    if (loaded) return api;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) throw std::runtime_error(std::string("cannot load NCCL (libnccl.so.2): ") + dlerror());
    auto bind = [&](auto& fn, const char* name) {
        void* sym = dlsym(h, name);
        if (!sym) throw std::runtime_error(std::string("NCCL symbol missing: ") + name);
        fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(sym);
    };
    bind(api.GetUniqueId, "ncclGetUniqueId");
    bind(api.CommInitRank, "ncclCommInitRank");
    bind(api.CommDestroy, "ncclCommDestroy");
    bind(api.CommAbort, "ncclCommAbort");
    bind(api.GetErrorString, "ncclGetErrorString");
    bind(api.AllGather, "ncclAllGather");
    bind(api.AllReduce, "ncclAllReduce");
    bind(api.Broadcast, "ncclBroadcast");
    bind(api.Send, "ncclSend");
    bind(api.Recv, "ncclRecv");
    bind(api.GroupStart, "ncclGroupStart");
    bind(api.GroupEnd, "ncclGroupEnd");
    loaded = true;
    return api;
}

struct Comm {
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    int64_t* dev_stage = nullptr;      // [kStageWords] send | [kStageWords * world] receive
    int64_t* host_stage = nullptr;     // pinned, same layout
    size_t stage_words = 0;
    uint64_t calls[5] = {0, 0, 0, 0, 0};      // all_gather, all_gather_v, all_to_all_v, all_reduce_sum_u32, host_all_gather_i64
    uint64_t bytes_sent = 0;
};

constexpr size_t kStageWords = 512;    // int64 per rank per host exchange

#define BQ_NCCL(expr)                                                                                   \
    do {                                                                                                \
        ncclResult_t _r = (expr);                                                                       \
        if (_r != ncclSuccess) {                                                                        \
            throw std::runtime_error(std::string("NCCL error: ") + nccl().GetErrorString(_r) + " at " +    \
                                     __FILE__ + ":" + std::to_string(__LINE__) + " (" #expr ")");       \
        }                                                                                               \
    } while (0)

static Comm* comm_of(bq_ctx* ctx) {
    auto* c = static_cast<Comm*>(ctx->comm);
    if (!c || !c->comm) throw std::runtime_error("no communicator: call bq_comm_init first");
    return c;
}

}  // namespace bq

using namespace bq;

extern "C" {

int bq_comm_unique_id(void* id128) {
    return guarded([&] {
        static_assert(sizeof(ncclUniqueId) == BQ_COMM_ID_BYTES, "unique id size");
        ncclUniqueId id;
        BQ_NCCL(nccl().GetUniqueId(&id));
        std::memcpy(id128, &id, sizeof id);
    });
}

int bq_comm_init(bq_ctx* ctx, int world, int rank, const void* id128) {
    return guarded([&] {
        if (world < 1 || rank < 0 || rank >= world) throw std::runtime_error("bq_comm_init: bad world / rank");
        if (ctx->comm) throw std::runtime_error("bq_comm_init: this context already has a communicator");
        BQ_CUDA(cudaSetDevice(ctx->device));
        auto* c = new Comm();
        try {
            c->world = world;
            c->rank = rank;
            ncclUniqueId id;
            std::memcpy(&id, id128, sizeof id);
            BQ_NCCL(nccl().CommInitRank(&c->comm, world, id, rank));
            c->stage_words = kStageWords * (static_cast<size_t>(world) + 1);
            BQ_CUDA(cudaMalloc(&c->dev_stage, c->stage_words * 8));
            BQ_CUDA(cudaMallocHost(&c->host_stage, c->stage_words * 8));
        } catch (...) {
            if (c->comm) nccl().CommAbort(c->comm);
            if (c->dev_stage) cudaFree(c->dev_stage);
            if (c->host_stage) cudaFreeHost(c->host_stage);
            delete c;
            throw;
        }
        ctx->comm = c;
    });
}

void bq_comm_destroy(bq_ctx* ctx) {
    if (!ctx || !ctx->comm) return;
    auto* c = static_cast<Comm*>(ctx->comm);
    cudaStreamSynchronize(ctx->stream);
    if (c->comm) nccl().CommDestroy(c->comm);
    if (c->dev_stage) cudaFree(c->dev_stage);
    if (c->host_stage) cudaFreeHost(c->host_stage);
    delete c;
    ctx->comm = nullptr;
}

int bq_comm_world(bq_ctx* ctx) { return ctx->comm ? static_cast<Comm*>(ctx->comm)->world : 1; }
int bq_comm_rank(bq_ctx* ctx) { return ctx->comm ? static_cast<Comm*>(ctx->comm)->rank : 0; }

int bq_comm_stats(bq_ctx* ctx, uint64_t* calls, uint64_t* bytes_sent) {
    return guarded([&] {
        Comm* c = comm_of(ctx);
        if (calls)
            for (int i = 0; i < 5; ++i) calls[i] = c->calls[i];
        if (bytes_sent) *bytes_sent = c->bytes_sent;
    });
}

int bq_comm_all_gather(bq_ctx* ctx, const void* send, void* recv, size_t bytes) {
    return guarded([&] {
        Comm* c = comm_of(ctx);
        c->calls[0]++;
        c->bytes_sent += bytes;
        if (bytes) BQ_NCCL(nccl().AllGather(send, recv, bytes, ncclUint8, c->comm, ctx->stream));
    });
}

int bq_comm_all_gather_v(bq_ctx* ctx, const void* send, void* recv, const int64_t* bytes_by_rank) {
    return guarded([&] {
        Comm* c = comm_of(ctx);
        c->calls[1]++;
        c->bytes_sent += static_cast<uint64_t>(bytes_by_rank[c->rank]);
        bool equal = true;
        for (int r = 1; r < c->world; ++r) equal = equal && bytes_by_rank[r] == bytes_by_rank[0];
        if (equal) {
            if (bytes_by_rank[0]) BQ_NCCL(nccl().AllGather(send, recv, static_cast<size_t>(bytes_by_rank[0]), ncclUint8, c->comm, ctx->stream));
            return;
        }
        // one broadcast per contributing rank, straight into its slot of the output (no padding, no staging copy)
        BQ_NCCL(nccl().GroupStart());
        size_t off = 0;
        for (int r = 0; r < c->world; ++r) {
            const size_t n = static_cast<size_t>(bytes_by_rank[r]);
            if (n) {
                char* slot = static_cast<char*>(recv) + off;
                ncclResult_t rc = nccl().Broadcast(r == c->rank ? send : slot, slot, n, ncclUint8, r, c->comm, ctx->stream);
                if (rc != ncclSuccess) {
                    nccl().GroupEnd();
                    BQ_NCCL(rc);
                }
            }
            off += n;
        }
        BQ_NCCL(nccl().GroupEnd());
    });
}

int bq_comm_all_to_all_v(bq_ctx* ctx, const void* send, const int64_t* send_bytes, void* recv, const int64_t* recv_bytes) {
    return guarded([&] {
        Comm* c = comm_of(ctx);
        c->calls[2]++;
        BQ_NCCL(nccl().GroupStart());
        size_t so = 0, ro = 0;
        ncclResult_t rc = ncclSuccess;
        for (int r = 0; r < c->world && rc == ncclSuccess; ++r) {
            const size_t sn = static_cast<size_t>(send_bytes[r]), rn = static_cast<size_t>(recv_bytes[r]);
            if (r != c->rank) c->bytes_sent += sn;
            if (sn) rc = nccl().Send(static_cast<const char*>(send) + so, sn, ncclUint8, r, c->comm, ctx->stream);
            if (rn && rc == ncclSuccess) rc = nccl().Recv(static_cast<char*>(recv) + ro, rn, ncclUint8, r, c->comm, ctx->stream);
            so += sn;
            ro += rn;
        }
        ncclResult_t end = nccl().GroupEnd();
        BQ_NCCL(rc);
        BQ_NCCL(end);
    });
}

int bq_comm_all_reduce_sum_u32(bq_ctx* ctx, void* buf, size_t words) {
    return guarded([&] {
        Comm* c = comm_of(ctx);
        c->calls[3]++;
        c->bytes_sent += words * 4;
        if (words) BQ_NCCL(nccl().AllReduce(buf, buf, words, ncclUint32, ncclSum, c->comm, ctx->stream));
    });
}

int bq_comm_host_all_gather_i64(bq_ctx* ctx, const int64_t* mine, int32_t n, int64_t* all) {
    return guarded([&] {
        Comm* c = comm_of(ctx);
        c->calls[4]++;
        if (n < 0 || static_cast<size_t>(n) > kStageWords) throw std::runtime_error("host exchange: at most 512 words per rank");
        if (n == 0) return;
        const size_t w = static_cast<size_t>(n);
        std::memcpy(c->host_stage, mine, w * 8);
        BQ_CUDA(cudaMemcpyAsync(c->dev_stage, c->host_stage, w * 8, cudaMemcpyHostToDevice, ctx->stream));
        BQ_NCCL(nccl().AllGather(c->dev_stage, c->dev_stage + kStageWords, w, ncclInt64, c->comm, ctx->stream));
        BQ_CUDA(cudaMemcpyAsync(c->host_stage + kStageWords, c->dev_stage + kStageWords, w * 8 * static_cast<size_t>(c->world),
                                cudaMemcpyDeviceToHost, ctx->stream));
        BQ_CUDA(cudaStreamSynchronize(ctx->stream));
        std::memcpy(all, c->host_stage + kStageWords, w * 8 * static_cast<size_t>(c->world));
    });
}

}  // extern "C"
