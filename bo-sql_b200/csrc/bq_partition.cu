// bq_partition.cu — hash partitioning of (key, payload...) rows: the building block of the exchange step.
//
// The reference has no counterpart (one process, one unordered_map: src/exec/operator.cpp:739-762, 984-1014).  A
// GROUP BY over 10^8 distinct keys or a join against a build side far larger than L2 makes every table update a random
// HBM transaction; partitioning rows by hash(key) first turns that into a stream: rows of one partition arrive
// together, so the part of the table they touch (capacity / P slots) stays resident in the 126 MB L2 while it is hot.
// The same kernel produces the per-peer send buffers of the multi-GPU shuffle (P = number of ranks).
//
//   k_part_hist     each CTA owns a contiguous row range and counts its rows per partition (shared-memory atomics)
//   (exclusive scan over [partition][CTA])                      -> where each CTA's rows of each partition go
//   k_part_scatter  each CTA re-reads its range tile by tile, counting-sorts the tile by partition in shared memory and
//                   writes every partition's run with coalesced stores (mean run = tile / P rows)
// HBM traffic: keys twice + payload once in, everything once out (24 B + 16 B per 16-byte row).
#include "bq_common.cuh"
#include "bq_internal.cuh"

namespace bq {

constexpr int kPartTile = 2048;                 // rows staged per CTA iteration
constexpr int kPartThreads = 512;               // scatter CTA: 4 rows per thread keeps it at 64 registers, two CTAs per SM
constexpr int kPartRows = kPartTile / kPartThreads;
constexpr int kHistThreads = 1024;              // histogram CTA (two per SM: one wave, same grid as the scatter)
constexpr int kHistCells = 2048;                // shared counters per CTA: P partitions x (kHistCells / P) lane-private copies
constexpr int kMaxPartLog2 = 10;
constexpr int kMaxHotKeys = 16;

struct PartParams {
    const void* key;
    int key_kind;
    const void* pay[2];
    int pay_w[2];
    int n_pay;
    size_t row_begin, n;
    int log2p;                         // size of the partition tables (hash partitions, plus the hot partition's half)
    int shift;
    unsigned hash_mask;                // hash partitions - 1
    int n_hot;                         // rows whose key is one of hot[] go to partition hash_mask + 1 instead (skew handling)
    long long hot[kMaxHotKeys];
    size_t rows_per_block;
    unsigned* hist;                    // [P][G]
    const unsigned long long* base;    // [P][G]
    // where partition q's rows go: dest[q] is the address of this launch's FIRST row of partition q (local output column,
    // or a peer GPU's buffer mapped through CUDA IPC: the shuffle writes straight over NVLink, no staging copy)
    void* const* dest_key;             // [P]
    void* const* dest_pay[2];          // [P] each
};

BQ_D unsigned part_of(const PartParams& p, long long k) {
    unsigned q = static_cast<unsigned>(key_hash(static_cast<uint64_t>(k)) >> p.shift) & p.hash_mask;
    if (p.n_hot) {
        bool hot = false;
#pragma unroll
        for (int i = 0; i < kMaxHotKeys; ++i) hot = hot || (i < p.n_hot && k == p.hot[i]);
        if (hot) q = p.hash_mask + 1;
    }
    return q;
}

__global__ void __launch_bounds__(kHistThreads) k_part_hist(const __grid_constant__ PartParams p) {
    __shared__ unsigned h[kHistCells];
    const unsigned P = 1u << p.log2p;
    // lanes of a warp use different copies of the counters, so a skewed or tiny partition set does not serialise the atomics
    const unsigned copies = P >= kHistCells ? 1u : kHistCells / P;
    const unsigned mine = (threadIdx.x % (copies < 32 ? copies : 32)) * P;
    for (unsigned i = threadIdx.x; i < kHistCells; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const size_t lo = blockIdx.x * p.rows_per_block;
    const size_t hi = lo + p.rows_per_block < p.n ? lo + p.rows_per_block : p.n;
    size_t t = lo + threadIdx.x;
    for (; t + 3 * (size_t)kHistThreads < hi; t += 4 * (size_t)kHistThreads) {          // four independent loads in flight
        long long k0 = load_raw(p.key, p.key_kind, p.row_begin + t);
        long long k1 = load_raw(p.key, p.key_kind, p.row_begin + t + kHistThreads);
        long long k2 = load_raw(p.key, p.key_kind, p.row_begin + t + 2 * kHistThreads);
        long long k3 = load_raw(p.key, p.key_kind, p.row_begin + t + 3 * kHistThreads);
        atomicAdd(&h[mine + part_of(p, k0)], 1u);
        atomicAdd(&h[mine + part_of(p, k1)], 1u);
        atomicAdd(&h[mine + part_of(p, k2)], 1u);
        atomicAdd(&h[mine + part_of(p, k3)], 1u);
    }
    for (; t < hi; t += kHistThreads) atomicAdd(&h[mine + part_of(p, load_raw(p.key, p.key_kind, p.row_begin + t))], 1u);
    __syncthreads();
    const unsigned used = copies < 32 ? copies : 32;
    for (unsigned i = threadIdx.x; i < P; i += blockDim.x) {
        unsigned c = 0;
        for (unsigned r = 0; r < used; ++r) c += h[r * P + i];
        p.hist[static_cast<size_t>(i) * gridDim.x + blockIdx.x] = c;
    }
}

BQ_D void store_narrow(void* base, int width, size_t i, long long v) {
    if (width == 8) static_cast<long long*>(base)[i] = v;
    else static_cast<int*>(base)[i] = static_cast<int>(v);
}
BQ_D long long load_width(const void* base, int width, size_t i) {
    return width == 8 ? __ldg(static_cast<const long long*>(base) + i) : static_cast<long long>(__ldg(static_cast<const int*>(base) + i));
}

__global__ void __launch_bounds__(kPartThreads, 2) k_part_scatter(const __grid_constant__ PartParams p) {
    extern __shared__ unsigned char smem_raw[];
    const unsigned P = 1u << p.log2p;
    // layout: skey[T] spay0[T] spay1[T] (8 B each) | cursor[P] dkey[P] dpay0[P] dpay1[P] (8 B) | cnt[P] start[P] (4 B) | spart[T] (2 B)
    long long* skey = reinterpret_cast<long long*>(smem_raw);
    long long* spay0 = skey + kPartTile;
    long long* spay1 = spay0 + kPartTile;
    unsigned long long* cursor = reinterpret_cast<unsigned long long*>(spay1 + kPartTile);
    void** dkey = reinterpret_cast<void**>(cursor + P);
    void** dpay0 = dkey + P;
    void** dpay1 = dpay0 + P;
    unsigned* cnt = reinterpret_cast<unsigned*>(dpay1 + P);
    unsigned* start = cnt + P;
    unsigned short* spart = reinterpret_cast<unsigned short*>(start + P);
    __shared__ unsigned warp_tot[kPartThreads / 32];

    for (unsigned i = threadIdx.x; i < P; i += blockDim.x) {
        // rows of partition i written by the CTAs before this one
        cursor[i] = p.base[static_cast<size_t>(i) * gridDim.x + blockIdx.x] - p.base[static_cast<size_t>(i) * gridDim.x];
        dkey[i] = p.dest_key[i];
        dpay0[i] = p.n_pay > 0 ? p.dest_pay[0][i] : nullptr;
        dpay1[i] = p.n_pay > 1 ? p.dest_pay[1][i] : nullptr;
    }
    const size_t lo = blockIdx.x * p.rows_per_block;
    const size_t hi = lo + p.rows_per_block < p.n ? lo + p.rows_per_block : p.n;
    const int key_w = width_of(p.key_kind);

    for (size_t tile = lo; tile < hi; tile += kPartTile) {
        const unsigned tn = static_cast<unsigned>(hi - tile < (size_t)kPartTile ? hi - tile : (size_t)kPartTile);
        for (unsigned i = threadIdx.x; i < P; i += blockDim.x) cnt[i] = 0;
        __syncthreads();
        long long k[kPartRows], v0[kPartRows], v1[kPartRows];
        unsigned part[kPartRows], rank[kPartRows];
#pragma unroll
        for (int j = 0; j < kPartRows; ++j) {
            const unsigned x = j * kPartThreads + threadIdx.x;
            if (x < tn) {
                const size_t row = p.row_begin + tile + x;
                k[j] = load_raw(p.key, p.key_kind, row);
                v0[j] = p.n_pay > 0 ? load_width(p.pay[0], p.pay_w[0], row) : 0;
                v1[j] = p.n_pay > 1 ? load_width(p.pay[1], p.pay_w[1], row) : 0;
            }
        }
#pragma unroll
        for (int j = 0; j < kPartRows; ++j) {
            const unsigned x = j * kPartThreads + threadIdx.x;
            if (x < tn) {
                part[j] = part_of(p, k[j]);
                rank[j] = atomicAdd(&cnt[part[j]], 1u);
            }
        }
        __syncthreads();
        // exclusive scan of cnt[0..P) into start[]: each thread owns P/kPartThreads consecutive entries (P >= kPartThreads) or one
        {
            const unsigned per = P > (unsigned)kPartThreads ? P / kPartThreads : 1;
            const unsigned first = threadIdx.x * per;
            unsigned local = 0;
            if (first < P)
                for (unsigned i = 0; i < per; ++i) local += cnt[first + i];
            unsigned incl = local;
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            if (lane == 31) warp_tot[warp] = incl;
            __syncthreads();
            unsigned wpre = 0;
            for (int w = 0; w < warp; ++w) wpre += warp_tot[w];
            unsigned run = wpre + incl - local;
            if (first < P)
                for (unsigned i = 0; i < per; ++i) {
                    start[first + i] = run;
                    run += cnt[first + i];
                }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kPartRows; ++j) {
            const unsigned x = j * kPartThreads + threadIdx.x;
            if (x < tn) {
                const unsigned at = start[part[j]] + rank[j];
                skey[at] = k[j];
                spay0[at] = v0[j];
                spay1[at] = v1[j];
                spart[at] = static_cast<unsigned short>(part[j]);
            }
        }
        __syncthreads();
        for (unsigned x = threadIdx.x; x < tn; x += blockDim.x) {
            const unsigned q = spart[x];
            const size_t g = cursor[q] + (x - start[q]);
            store_narrow(dkey[q], key_w, g, skey[x]);
            if (p.n_pay > 0) store_narrow(dpay0[q], p.pay_w[0], g, spay0[x]);
            if (p.n_pay > 1) store_narrow(dpay1[q], p.pay_w[1], g, spay1[x]);
        }
        __syncthreads();
        for (unsigned i = threadIdx.x; i < P; i += blockDim.x) cursor[i] += cnt[i];
        __syncthreads();
    }
}

__global__ void k_part_offsets(const unsigned long long* __restrict__ base, unsigned G, unsigned P, size_t n, long long* __restrict__ out) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) out[i] = static_cast<long long>(base[static_cast<size_t>(i) * G]);
    if (i == P) out[P] = static_cast<long long>(n);
}

}  // namespace bq

using namespace bq;

// Counting half of a partition pass; kept between bq_partition_count and bq_partition_scatter.
struct bq_part_plan {
    bq_ctx* ctx = nullptr;
    PartParams p{};
    unsigned G = 0, P = 0;
    void* hist = nullptr;
    void* base = nullptr;
    std::vector<int64_t> counts;      // rows per partition (host)
};

static void part_plan_release(bq_part_plan* pl) {
    if (!pl) return;
    dev_free(pl->ctx, pl->hist);
    dev_free(pl->ctx, pl->base);
    delete pl;
}

// histogram + scan; `want_host_counts` costs one device-to-host copy (the shuffle needs the sizes on the host anyway)
static bq_part_plan* part_count(bq_ctx* ctx, const bq_col* key, size_t row_begin, size_t row_end, int log2_parts, int hash_shift,
                                bool want_host_counts, const int64_t* hot_keys = nullptr, int n_hot = 0) {
    if (n_hot < 0 || n_hot > kMaxHotKeys) throw std::runtime_error("at most 16 hot keys per partition pass");
    const int log2_hash = log2_parts;
    if (n_hot) ++log2_parts;           // the hot partition is index 2^log2_hash; the rest of the upper half stays empty
    if (log2_parts < 0 || log2_parts > kMaxPartLog2) throw std::runtime_error("partition count must be 1 .. 1024 (a power of two)");
    if (row_end < row_begin || row_end > key->n) throw std::runtime_error("bad row range");
    const size_t n = row_end - row_begin;
    if (n > 0xFFFFFFFFull) throw std::runtime_error("at most 2^32 rows per partition pass");
    auto* pl = new bq_part_plan();
    pl->ctx = ctx;
    pl->P = 1u << log2_parts;
    PartParams& p = pl->p;
    p.key = key->ptr;
    p.key_kind = key->type;
    p.row_begin = row_begin;
    p.n = n;
    p.log2p = log2_parts;
    p.shift = hash_shift;
    p.hash_mask = (1u << log2_hash) - 1;
    p.n_hot = n_hot;
    for (int i = 0; i < n_hot; ++i) p.hot[i] = hot_keys[i];
    unsigned G = static_cast<unsigned>(ctx->sm_count) * 2;        // what is resident at once: one wave for both kernels
    const size_t tiles = (n + kPartTile - 1) / kPartTile;
    if (tiles < G) G = static_cast<unsigned>(tiles ? tiles : 1);
    pl->G = G;
    p.rows_per_block = ((tiles + G - 1) / G) * kPartTile;
    try {
        const size_t cells = static_cast<size_t>(pl->P) * G;
        pl->hist = dev_alloc(ctx, cells * 4);
        pl->base = dev_alloc(ctx, (cells + 1) * 8);
        p.hist = static_cast<unsigned*>(pl->hist);
        p.base = static_cast<unsigned long long*>(pl->base);
        pl->counts.assign(pl->P, 0);
        if (n) {
            k_part_hist<<<G, kHistThreads, 0, ctx->stream>>>(p);
            ctx->launches++;
            BQ_CUDA(cudaGetLastError());
            exclusive_scan_u32(ctx, p.hist, cells, static_cast<unsigned long long*>(pl->base));
            if (want_host_counts) {
                // partition starts = base[q * G]; read them with one strided copy
                std::vector<unsigned long long> starts(pl->P);
                BQ_CUDA(cudaMemcpy2DAsync(starts.data(), 8, pl->base, static_cast<size_t>(G) * 8, 8, pl->P, cudaMemcpyDeviceToHost, ctx->stream));
                BQ_CUDA(cudaStreamSynchronize(ctx->stream));
                for (unsigned q = 0; q < pl->P; ++q)
                    pl->counts[q] = static_cast<int64_t>((q + 1 < pl->P ? starts[q + 1] : n) - starts[q]);
            }
        }
    } catch (...) {
        part_plan_release(pl);
        throw;
    }
    return pl;
}

static void part_scatter(bq_part_plan* pl, const bq_col* const* payload, int n_payload, size_t row_end, void* const* d_dest_key,
                         void* const* d_dest_pay0, void* const* d_dest_pay1) {
    bq_ctx* ctx = pl->ctx;
    PartParams& p = pl->p;
    if (n_payload < 0 || n_payload > 2) throw std::runtime_error("at most two payload columns per partition pass");
    p.n_pay = n_payload;
    for (int i = 0; i < n_payload; ++i) {
        if (payload[i]->n < row_end) throw std::runtime_error("payload column shorter than the row range");
        p.pay[i] = payload[i]->ptr;
        p.pay_w[i] = width_of(payload[i]->type);
    }
    p.dest_key = d_dest_key;
    p.dest_pay[0] = d_dest_pay0;
    p.dest_pay[1] = d_dest_pay1;
    if (!p.n) return;
    const size_t smem = static_cast<size_t>(kPartTile) * 24 + pl->P * 40 + kPartTile * 2;
    BQ_CUDA(cudaFuncSetAttribute(k_part_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    k_part_scatter<<<pl->G, kPartThreads, smem, ctx->stream>>>(p);
    ctx->launches++;
    BQ_CUDA(cudaGetLastError());
}

// device table of P destination addresses: dest[q] = base + start_row[q] * width (local outputs)
__global__ void k_part_local_dest(const unsigned long long* __restrict__ base, unsigned G, unsigned P, char* out, int width, void** dest) {
    unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < P) dest[q] = out + base[static_cast<size_t>(q) * G] * static_cast<unsigned long long>(width);
}

extern "C" int bq_partition(bq_ctx* ctx, const bq_col* key, const bq_col* const* payload, int n_payload, size_t row_begin,
                            size_t row_end, int log2_parts, int hash_shift, bq_col** out_key, bq_col** out_payload,
                            bq_col** out_offsets) {
    return guarded([&] {
        if (n_payload < 0 || n_payload > 2) throw std::runtime_error("at most two payload columns per partition pass");
        bq_part_plan* pl = part_count(ctx, key, row_begin, row_end, log2_parts, hash_shift, false);
        const size_t n = pl->p.n;
        const unsigned P = pl->P, G = pl->G;
        bq_col* ok = nullptr;
        bq_col* op[2] = {nullptr, nullptr};
        bq_col* off = nullptr;
        try {
            ok = new_col(ctx, key->type, n);
            for (int i = 0; i < n_payload; ++i) op[i] = new_col(ctx, payload[i]->type, n);
            off = new_col(ctx, BQ_INT64, P + 1);
            if (n) {
                DevBuf dest(ctx, static_cast<size_t>(P) * 8 * 3);
                void** d = dest.as<void*>();
                const unsigned blocks = (P + 255) / 256;
                k_part_local_dest<<<blocks, 256, 0, ctx->stream>>>(pl->p.base, G, P, static_cast<char*>(ok->ptr), width_of(key->type), d);
                for (int i = 0; i < n_payload; ++i)
                    k_part_local_dest<<<blocks, 256, 0, ctx->stream>>>(pl->p.base, G, P, static_cast<char*>(op[i]->ptr), width_of(payload[i]->type), d + (i + 1) * P);
                ctx->launches += 1 + n_payload;
                part_scatter(pl, payload, n_payload, row_end, d, d + P, d + 2 * P);
                k_part_offsets<<<(P + 1 + 255) / 256, 256, 0, ctx->stream>>>(pl->p.base, G, P, n, static_cast<long long*>(off->ptr));
                ctx->launches++;
                BQ_CUDA(cudaGetLastError());
            } else {
                BQ_CUDA(cudaMemsetAsync(off->ptr, 0, (P + 1) * 8, ctx->stream));
            }
        } catch (...) {
            free_col(ok);
            free_col(op[0]);
            free_col(op[1]);
            free_col(off);
            part_plan_release(pl);
            throw;
        }
        part_plan_release(pl);
        *out_key = ok;
        for (int i = 0; i < n_payload; ++i) out_payload[i] = op[i];
        *out_offsets = off;
    });
}

extern "C" int bq_partition_count(bq_ctx* ctx, const bq_col* key, size_t row_begin, size_t row_end, int log2_parts, int hash_shift,
                                  int64_t* host_counts, bq_part_plan** out) {
    return guarded([&] {
        bq_part_plan* pl = part_count(ctx, key, row_begin, row_end, log2_parts, hash_shift, true);
        for (unsigned q = 0; q < pl->P; ++q) host_counts[q] = pl->counts[q];
        *out = pl;
    });
}

extern "C" int bq_partition_count_hot(bq_ctx* ctx, const bq_col* key, size_t row_begin, size_t row_end, int log2_parts, int hash_shift,
                                      const int64_t* hot_keys, int n_hot, int64_t* host_counts, bq_part_plan** out) {
    return guarded([&] {
        bq_part_plan* pl = part_count(ctx, key, row_begin, row_end, log2_parts, hash_shift, true, hot_keys, n_hot);
        for (unsigned q = 0; q < pl->P; ++q) host_counts[q] = pl->counts[q];
        *out = pl;
    });
}

extern "C" int bq_partition_scatter(bq_ctx* ctx, bq_part_plan* plan, const bq_col* const* payload, int n_payload,
                                    void* const* dest_key, void* const* dest_pay0, void* const* dest_pay1) {
    return guarded([&] {
        if (!plan || plan->ctx != ctx) throw std::runtime_error("bad partition plan");
        const unsigned P = plan->P;
        // the destination tables travel to the device in one small stream-ordered copy
        std::vector<void*> host(static_cast<size_t>(P) * 3, nullptr);
        for (unsigned q = 0; q < P; ++q) {
            host[q] = dest_key[q];
            if (n_payload > 0) host[P + q] = dest_pay0[q];
            if (n_payload > 1) host[2 * P + q] = dest_pay1[q];
        }
        DevBuf dest(ctx, host.size() * 8);
        BQ_CUDA(cudaMemcpyAsync(dest.p, host.data(), host.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        BQ_CUDA(cudaStreamSynchronize(ctx->stream));       // `host` is pageable and about to go out of scope
        void** d = dest.as<void*>();
        part_scatter(plan, payload, n_payload, plan->p.row_begin + plan->p.n, d, d + P, d + 2 * P);
    });
}

extern "C" void bq_part_plan_free(bq_part_plan* plan) { part_plan_release(plan); }
