// bq_ctx.cu — context, device-resident columns and relations (the storage/ side of the hot path).
//
// Replaces, on the device, what the reference keeps in std::vector<T> inside ColumnVector<T>
// (include/types.h:134-145) and Table (include/storage/table.h:20-30).  A column is one contiguous
// HBM allocation (cudaMalloc: 256-byte aligned, so every 128-bit load in the scan kernels is aligned).
#include "bq_common.cuh"

#include <cstring>
#include <mutex>
#include <set>

namespace bq {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }

static std::mutex g_live_mu;
static std::set<bq_ctx*> g_live;
static bool is_live(bq_ctx* ctx) {
    std::lock_guard<std::mutex> lk(g_live_mu);
    return ctx && g_live.count(ctx);
}

static constexpr size_t kBigBlock = 32ull << 20;          // allocations from this size up are recycled whole
static constexpr size_t kBigFreeCap = 64ull << 30;        // idle bytes kept before the oldest idle block goes back

static void big_release_idle(bq_ctx* ctx, size_t keep_bytes) {
    while (ctx->big_free_bytes > keep_bytes) {
        size_t oldest = ctx->big_blocks.size();
        for (size_t i = 0; i < ctx->big_blocks.size(); ++i)
            if (ctx->big_blocks[i].free && !ctx->big_blocks[i].exported &&          // a peer may still hold a mapping of an exported block
                (oldest == ctx->big_blocks.size() || ctx->big_blocks[i].stamp < ctx->big_blocks[oldest].stamp)) oldest = i;
        if (oldest == ctx->big_blocks.size()) break;
        cudaFree(ctx->big_blocks[oldest].p);              // synchronises the device: nothing can still be using it
        ctx->big_free_bytes -= ctx->big_blocks[oldest].bytes;
        ctx->big_blocks.erase(ctx->big_blocks.begin() + static_cast<long>(oldest));
    }
}

void* dev_alloc(bq_ctx* ctx, size_t bytes) {
    void* p = nullptr;
    if (bytes < kBigBlock) {
        BQ_CUDA(cudaMallocAsync(&p, bytes ? bytes : 16, ctx->stream));
        return p;
    }
    const size_t want = (bytes + kBigBlock - 1) / kBigBlock * kBigBlock;
    size_t best = ctx->big_blocks.size();
    for (size_t i = 0; i < ctx->big_blocks.size(); ++i) {
        const auto& b = ctx->big_blocks[i];
        if (b.free && b.bytes >= want && b.bytes <= want + want / 4 && (best == ctx->big_blocks.size() || b.bytes < ctx->big_blocks[best].bytes)) best = i;
    }
    if (best != ctx->big_blocks.size()) {
        auto& b = ctx->big_blocks[best];
        b.free = false;
        ctx->big_free_bytes -= b.bytes;
        return b.p;
    }
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        big_release_idle(ctx, 0);                          // out of memory: give every idle block back and try once more
        BQ_CUDA(cudaMalloc(&p, want));
    }
    ctx->big_blocks.push_back({p, want, false, 0});
    return p;
}

void dev_free(bq_ctx* ctx, void* p) {
    if (!p) return;
    if (!is_live(ctx)) {
        cudaFree(p);
        return;
    }
    for (auto& b : ctx->big_blocks)
        if (b.p == p) {
            b.free = true;
            b.stamp = ++ctx->big_clock;
            ctx->big_free_bytes += b.bytes;
            if (ctx->big_free_bytes > kBigFreeCap) big_release_idle(ctx, kBigFreeCap);
            return;
        }
    cudaFreeAsync(p, ctx->stream);
}

bq_col* new_col(bq_ctx* ctx, int type, size_t n) {
    if (type < 0 || type > 3) throw std::runtime_error("Unknown column type");
    auto* c = new bq_col();
    c->ctx = ctx;
    c->type = type;
    c->n = n;
    size_t bytes = n * width_of(type);
    // pad to a whole 16-byte vector so tail vector loads never leave the allocation
    size_t alloc = ((bytes + 255) / 256) * 256 + 256;
    try {
        c->ptr = dev_alloc(ctx, alloc);
    } catch (...) {
        delete c;
        throw;
    }
    return c;
}

void free_col(bq_col* c) {
    if (!c) return;
    if (c->owns && c->ptr) dev_free(c->ctx, c->ptr);
    delete c;
}

void* scratch(bq_ctx* ctx, size_t bytes) {
    if (bytes > ctx->scratch_bytes) {
        if (ctx->scratch) {
            dev_free(ctx, ctx->scratch);
            ctx->scratch = nullptr;
        }
        size_t want = bytes < (1u << 20) ? (1u << 20) : bytes;
        ctx->scratch = dev_alloc(ctx, want);
        ctx->scratch_bytes = want;
    }
    return ctx->scratch;
}

void* pinned(bq_ctx* ctx, size_t bytes) {
    if (bytes > ctx->pinned_bytes) {
        if (ctx->pinned) cudaFreeHost(ctx->pinned);
        size_t want = bytes < 4096 ? 4096 : bytes;
        BQ_CUDA(cudaMallocHost(&ctx->pinned, want));
        ctx->pinned_bytes = want;
    }
    return ctx->pinned;
}

// Persistent-style grids: a multiple of the SM count, capped by the work available.
int grid_for(bq_ctx* ctx, size_t rows, int blocks_per_sm) {
    size_t tiles = (rows + kTileRows - 1) / kTileRows;
    size_t want = static_cast<size_t>(ctx->sm_count) * blocks_per_sm;
    if (tiles < want) want = tiles;
    if (want < 1) want = 1;
    return static_cast<int>(want);
}

// ---- min/max of a column's integer key (cached in the handle) -----------------------------------
__global__ void __launch_bounds__(kBlock) k_minmax(const void* __restrict__ base, int kind, size_t n,
                                                   long long* __restrict__ out /* [min,max] */) {
    long long lo = INT64_MAX, hi = INT64_MIN;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        long long k = key_of(load_raw(base, kind, i), kind);
        lo = k < lo ? k : lo;
        hi = k > hi ? k : hi;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        long long l2 = __shfl_xor_sync(0xffffffffu, lo, o);
        long long h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out, lo);
        atomicMax(out + 1, hi);
    }
}

}  // namespace bq

using namespace bq;

extern "C" {

const char* bq_last_error(void) { return g_error.c_str(); }

int bq_ctx_create(int device, bq_ctx** out) {
    return guarded([&] {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0)
            throw std::runtime_error(std::string("no CUDA device: the bo-sql B200 hot path has no CPU fallback (") +
                                     cudaGetErrorString(e) + ")");
        if (device < 0 || device >= n) throw std::runtime_error("bad device ordinal");
        BQ_CUDA(cudaSetDevice(device));
        auto* ctx = new bq_ctx();
        ctx->device = device;
        cudaDeviceProp prop;
        BQ_CUDA(cudaGetDeviceProperties(&prop, device));
        ctx->sm_count = prop.multiProcessorCount;
        BQ_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
        ctx->stream = ctx->own_stream;
        // keep freed blocks in the pool: a query's temporaries are reused by the next query without driver calls
        cudaMemPool_t pool;
        BQ_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = UINT64_MAX;
        BQ_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        {
            std::lock_guard<std::mutex> lk(g_live_mu);
            g_live.insert(ctx);
        }
        *out = ctx;
    });
}

void bq_ctx_destroy(bq_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    bq_comm_destroy(ctx);
    {
        std::lock_guard<std::mutex> lk(g_live_mu);
        g_live.erase(ctx);
    }
    for (auto& m : ctx->ipc_mappings) cudaIpcCloseMemHandle(m.second);
    ctx->ipc_mappings.clear();
    for (auto& b : ctx->big_blocks)
        if (b.free) cudaFree(b.p);           // blocks still held by a column are freed by that column (dev_free, dead context)
    ctx->big_blocks.clear();
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int bq_ctx_set_stream(bq_ctx* ctx, void* s) {
    return guarded([&] {
        BQ_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->stream = s ? static_cast<cudaStream_t>(s) : ctx->own_stream;
    });
}

int bq_ctx_sync(bq_ctx* ctx) {
    return guarded([&] { BQ_CUDA(cudaStreamSynchronize(ctx->stream)); });
}

void* bq_ctx_stream(bq_ctx* ctx) { return ctx->stream; }

// ---- peer memory: one process per GPU, buffers shared through CUDA IPC ---------------------------------------------
int bq_col_alloc_shared(bq_ctx* ctx, int type, size_t n, bq_col** out) {
    return guarded([&] {
        if (type < 0 || type > 3) throw std::runtime_error("Unknown column type");
        auto* c = new bq_col();
        c->ctx = ctx;
        c->type = type;
        c->n = n;
        try {
            size_t bytes = n * width_of(type);
            c->ptr = dev_alloc(ctx, bytes < kBigBlock ? kBigBlock : bytes);      // IPC needs a cudaMalloc'ed block of its own
        } catch (...) {
            delete c;
            throw;
        }
        *out = c;
    });
}

int bq_col_ipc_export(bq_ctx* ctx, const bq_col* col, void* handle64) {
    return guarded([&] {
        static_assert(sizeof(cudaIpcMemHandle_t) == BQ_IPC_HANDLE_BYTES, "handle size");
        for (auto& b : ctx->big_blocks)
            if (b.p == col->ptr) {
                cudaIpcMemHandle_t h;
                BQ_CUDA(cudaIpcGetMemHandle(&h, col->ptr));
                std::memcpy(handle64, &h, sizeof h);
                b.exported = true;
                return;
            }
        throw std::runtime_error("only columns from bq_col_alloc_shared (or >= 32 MB) can be shared with a peer");
    });
}

int bq_ipc_open(bq_ctx* ctx, const void* handle64, void** device_ptr) {
    return guarded([&] {
        std::array<unsigned char, 64> key;
        std::memcpy(key.data(), handle64, 64);
        for (auto& m : ctx->ipc_mappings)
            if (m.first == key) {
                *device_ptr = m.second;
                return;
            }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handle64, sizeof h);
        void* p = nullptr;
        BQ_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->ipc_mappings.emplace_back(key, p);
        *device_ptr = p;
    });
}

size_t bq_ctx_ipc_mappings(bq_ctx* ctx) { return ctx->ipc_mappings.size(); }

int bq_ctx_pool_stats(bq_ctx* ctx, size_t* reserved_bytes, size_t* used_bytes) {
    return guarded([&] {
        cudaMemPool_t pool;
        BQ_CUDA(cudaDeviceGetMemPool(&pool, ctx->device));
        uint64_t r = 0, u = 0;
        BQ_CUDA(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &r));
        BQ_CUDA(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &u));
        for (const auto& b : ctx->big_blocks) {          // recycled large blocks count as held; the ones handed out as used
            r += b.bytes;
            if (!b.free) u += b.bytes;
        }
        if (reserved_bytes) *reserved_bytes = r;
        if (used_bytes) *used_bytes = u;
    });
}

int bq_copy_bytes(bq_ctx* ctx, void* dst, const void* src, size_t bytes) {
    return guarded([&] {
        if (bytes) BQ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    });
}

int bq_zero_bytes(bq_ctx* ctx, void* dst, size_t bytes) {
    return guarded([&] {
        if (bytes) BQ_CUDA(cudaMemsetAsync(dst, 0, bytes, ctx->stream));
    });
}

int bq_ctx_info(bq_ctx* ctx, int* sm_count, size_t* free_bytes, size_t* total_bytes) {
    return guarded([&] {
        BQ_CUDA(cudaSetDevice(ctx->device));
        if (sm_count) *sm_count = ctx->sm_count;
        size_t f = 0, t = 0;
        BQ_CUDA(cudaMemGetInfo(&f, &t));
        if (free_bytes) *free_bytes = f;
        if (total_bytes) *total_bytes = t;
    });
}

uint64_t bq_ctx_launches(bq_ctx* ctx) { return ctx->launches; }

int bq_ctx_profile(bq_ctx* ctx, int enable) {
    ctx->profile = enable != 0;
    return 0;
}

int bq_ctx_profile_read(bq_ctx* ctx, uint64_t* launches, double* total_ms) {
    return guarded([&] {
        BQ_CUDA(cudaStreamSynchronize(ctx->stream));
        double total = 0.0;
        for (auto& pr : ctx->profile_events) {
            float ms = 0.f;
            BQ_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
            total += ms;
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
        if (launches) *launches = ctx->profile_events.size();
        if (total_ms) *total_ms = total;
        ctx->profile_events.clear();
    });
}

int bq_col_wrap(bq_ctx* ctx, int type, void* device_ptr, size_t n, bq_col** out) {
    return guarded([&] {
        if (type < 0 || type > 3) throw std::runtime_error("Unknown column type");
        auto* c = new bq_col();
        c->ctx = ctx;
        c->type = type;
        c->n = n;
        c->ptr = device_ptr;
        c->owns = false;
        *out = c;
    });
}

int bq_col_alloc(bq_ctx* ctx, int type, size_t n, bq_col** out) {
    return guarded([&] { *out = new_col(ctx, type, n); });
}

int bq_col_upload(bq_ctx* ctx, int type, const void* host, size_t n, bq_col** out) {
    return guarded([&] {
        bq_col* c = new_col(ctx, type, n);
        if (n) {
            cudaError_t e = cudaMemcpyAsync(c->ptr, host, n * width_of(type), cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) {
                free_col(c);
                throw std::runtime_error(std::string("upload failed: ") + cudaGetErrorString(e));
            }
        }
        *out = c;
    });
}

int bq_col_write(bq_ctx* ctx, bq_col* col, size_t offset, const void* host, size_t n) {
    return guarded([&] {
        if (offset + n > col->n) throw std::runtime_error("bq_col_write out of range");
        size_t w = width_of(col->type);
        BQ_CUDA(cudaMemcpyAsync(static_cast<char*>(col->ptr) + offset * w, host, n * w, cudaMemcpyHostToDevice,
                                ctx->stream));
        col->has_minmax = false;
    });
}

int bq_col_read(bq_ctx* ctx, const bq_col* col, size_t offset, size_t n, void* host) {
    return guarded([&] {
        if (offset + n > col->n) throw std::runtime_error("bq_col_read out of range");
        if (!n) return;
        size_t w = width_of(col->type);
        BQ_CUDA(cudaMemcpyAsync(host, static_cast<const char*>(col->ptr) + offset * w, n * w, cudaMemcpyDeviceToHost,
                                ctx->stream));
        BQ_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

int bq_col_read_async(bq_ctx* ctx, const bq_col* col, size_t offset, size_t n, void* host) {
    return guarded([&] {
        if (offset + n > col->n) throw std::runtime_error("bq_col_read out of range");
        if (!n) return;
        size_t w = width_of(col->type);
        BQ_CUDA(cudaMemcpyAsync(host, static_cast<const char*>(col->ptr) + offset * w, n * w, cudaMemcpyDeviceToHost,
                                ctx->stream));
    });
}

void bq_col_free(bq_ctx* ctx, bq_col* col) {
    (void)ctx;          // the column remembers its owner; release is stream-ordered, no synchronisation
    free_col(col);
}

size_t bq_col_size(const bq_col* col) { return col->n; }
int bq_col_owns(const bq_col* col) { return col->owns && col->ptr ? 1 : 0; }
int bq_col_type(const bq_col* col) { return col->type; }
void* bq_col_ptr(const bq_col* col) { return col->ptr; }

int bq_col_set_stats(bq_col* col, int64_t min_key, int64_t max_key, size_t ndv) {
    col->has_minmax = true;
    col->min_key = min_key;
    col->max_key = max_key;
    col->ndv = ndv;
    return 0;
}

void bq_col_invalidate_stats(bq_col* col) { col->has_minmax = false; }

int bq_col_minmax(bq_ctx* ctx, bq_col* col, int64_t* min_key, int64_t* max_key) {
    return guarded([&] {
        if (!col->has_minmax) {
            if (col->n == 0) {
                col->min_key = 0;
                col->max_key = -1;
            } else {
                auto* d = static_cast<long long*>(scratch(ctx, 16));
                auto* h = static_cast<long long*>(pinned(ctx, 16));
                h[0] = INT64_MAX;
                h[1] = INT64_MIN;
                BQ_CUDA(cudaMemcpyAsync(d, h, 16, cudaMemcpyHostToDevice, ctx->stream));
                int grid = grid_for(ctx, col->n, 8);
                k_minmax<<<grid, kBlock, 0, ctx->stream>>>(col->ptr, col->type, col->n, d);
                ctx->launches++;
                BQ_CUDA(cudaGetLastError());
                BQ_CUDA(cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, ctx->stream));
                BQ_CUDA(cudaStreamSynchronize(ctx->stream));
                col->min_key = h[0];
                col->max_key = h[1];
            }
            col->has_minmax = true;
        }
        if (min_key) *min_key = col->min_key;
        if (max_key) *max_key = col->max_key;
    });
}

uint64_t bq_key_hash(int64_t key) { return key_hash(static_cast<uint64_t>(key)); }

int64_t bq_f64_key(double v) {
    uint64_t b;
    std::memcpy(&b, &v, 8);
    return f64_key_from_bits(b);
}
double bq_f64_from_key(int64_t k) {
    uint64_t b = f64_bits_from_key(k);
    double v;
    std::memcpy(&v, &b, 8);
    return v;
}

int bq_host_alloc(size_t bytes, void** out) {
    return guarded([&] { BQ_CUDA(cudaMallocHost(out, bytes ? bytes : 1)); });
}
void bq_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int bq_rel_create(bq_ctx* ctx, bq_col* const* cols, int n_cols, bq_rel** out) {
    (void)ctx;
    return guarded([&] {
        auto* r = new bq_rel();
        for (int i = 0; i < n_cols; ++i) {
            if (i && cols[i]->n != cols[0]->n) {
                delete r;
                throw std::runtime_error("relation columns differ in length");
            }
            r->cols.push_back(cols[i]);
        }
        r->rows = n_cols ? cols[0]->n : 0;
        *out = r;
    });
}
size_t bq_rel_rows(const bq_rel* rel) { return rel->rows; }
int bq_rel_cols(const bq_rel* rel) { return static_cast<int>(rel->cols.size()); }
bq_col* bq_rel_col(const bq_rel* rel, int i) { return rel->cols.at(i); }
void bq_rel_free(bq_ctx* ctx, bq_rel* rel) {
    (void)ctx;
    if (!rel) return;
    for (auto* c : rel->cols) free_col(c);
    delete rel;
}

void bq_rel_release(bq_rel* rel, bq_col** out_cols) {
    if (!rel) return;
    for (size_t i = 0; i < rel->cols.size(); ++i) out_cols[i] = rel->cols[i];
    delete rel;
}

}  // extern "C"
