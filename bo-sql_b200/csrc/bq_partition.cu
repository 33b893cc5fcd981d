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
constexpr int kPartRows = kPartTile / kBlock;   // rows per thread per tile
constexpr int kMaxPartLog2 = 10;

struct PartParams {
    const void* key;
    int key_kind;
    const void* pay[2];
    int pay_w[2];
    int n_pay;
    size_t row_begin, n;
    int log2p;
    int shift;
    size_t rows_per_block;
    unsigned* hist;                    // [P][G]
    const unsigned long long* base;    // [P][G]
    void* out_key;
    void* out_pay[2];
};

BQ_D unsigned part_of(long long k, int shift, unsigned mask) {
    return static_cast<unsigned>(key_hash(static_cast<uint64_t>(k)) >> shift) & mask;
}

__global__ void __launch_bounds__(kBlock) k_part_hist(const __grid_constant__ PartParams p) {
    extern __shared__ unsigned h[];
    const unsigned P = 1u << p.log2p;
    for (unsigned i = threadIdx.x; i < P; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const size_t lo = blockIdx.x * p.rows_per_block;
    const size_t hi = lo + p.rows_per_block < p.n ? lo + p.rows_per_block : p.n;
    for (size_t t = lo + threadIdx.x; t < hi; t += blockDim.x) {
        long long k = load_raw(p.key, p.key_kind, p.row_begin + t);
        atomicAdd(&h[part_of(k, p.shift, P - 1)], 1u);
    }
    __syncthreads();
    for (unsigned i = threadIdx.x; i < P; i += blockDim.x) p.hist[static_cast<size_t>(i) * gridDim.x + blockIdx.x] = h[i];
}

BQ_D void store_narrow(void* base, int width, size_t i, long long v) {
    if (width == 8) static_cast<long long*>(base)[i] = v;
    else static_cast<int*>(base)[i] = static_cast<int>(v);
}
BQ_D long long load_width(const void* base, int width, size_t i) {
    return width == 8 ? __ldg(static_cast<const long long*>(base) + i) : static_cast<long long>(__ldg(static_cast<const int*>(base) + i));
}

__global__ void __launch_bounds__(kBlock) k_part_scatter(const __grid_constant__ PartParams p) {
    extern __shared__ unsigned char smem_raw[];
    const unsigned P = 1u << p.log2p;
    // layout: skey[T] spay0[T] spay1[T] (8 B each) | cursor[P] (8 B) | cnt[P] start[P] (4 B) | spart[T] (2 B)
    long long* skey = reinterpret_cast<long long*>(smem_raw);
    long long* spay0 = skey + kPartTile;
    long long* spay1 = spay0 + kPartTile;
    unsigned long long* cursor = reinterpret_cast<unsigned long long*>(spay1 + kPartTile);
    unsigned* cnt = reinterpret_cast<unsigned*>(cursor + P);
    unsigned* start = cnt + P;
    unsigned short* spart = reinterpret_cast<unsigned short*>(start + P);
    __shared__ unsigned warp_tot[kBlock / 32];

    for (unsigned i = threadIdx.x; i < P; i += blockDim.x) cursor[i] = p.base[static_cast<size_t>(i) * gridDim.x + blockIdx.x];
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
            const unsigned x = j * kBlock + threadIdx.x;
            if (x < tn) {
                const size_t row = p.row_begin + tile + x;
                k[j] = load_raw(p.key, p.key_kind, row);
                v0[j] = p.n_pay > 0 ? load_width(p.pay[0], p.pay_w[0], row) : 0;
                v1[j] = p.n_pay > 1 ? load_width(p.pay[1], p.pay_w[1], row) : 0;
            }
        }
#pragma unroll
        for (int j = 0; j < kPartRows; ++j) {
            const unsigned x = j * kBlock + threadIdx.x;
            if (x < tn) {
                part[j] = part_of(k[j], p.shift, P - 1);
                rank[j] = atomicAdd(&cnt[part[j]], 1u);
            }
        }
        __syncthreads();
        // exclusive scan of cnt[0..P) into start[]: each thread owns P/kBlock consecutive entries (P >= kBlock) or one
        {
            const unsigned per = P > (unsigned)kBlock ? P / kBlock : 1;
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
            const unsigned x = j * kBlock + threadIdx.x;
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
            store_narrow(p.out_key, key_w, g, skey[x]);
            if (p.n_pay > 0) store_narrow(p.out_pay[0], p.pay_w[0], g, spay0[x]);
            if (p.n_pay > 1) store_narrow(p.out_pay[1], p.pay_w[1], g, spay1[x]);
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

extern "C" int bq_partition(bq_ctx* ctx, const bq_col* key, const bq_col* const* payload, int n_payload, size_t row_begin,
                            size_t row_end, int log2_parts, int hash_shift, bq_col** out_key, bq_col** out_payload,
                            bq_col** out_offsets) {
    return guarded([&] {
        if (log2_parts < 0 || log2_parts > kMaxPartLog2) throw std::runtime_error("partition count must be 1 .. 1024 (a power of two)");
        if (n_payload < 0 || n_payload > 2) throw std::runtime_error("at most two payload columns per partition pass");
        if (row_end < row_begin || row_end > key->n) throw std::runtime_error("bad row range");
        const size_t n = row_end - row_begin;
        if (n > 0xFFFFFFFFull) throw std::runtime_error("at most 2^32 rows per partition pass");
        const unsigned P = 1u << log2_parts;
        PartParams p{};
        p.key = key->ptr;
        p.key_kind = key->type;
        p.n_pay = n_payload;
        for (int i = 0; i < n_payload; ++i) {
            if (payload[i]->n < row_end) throw std::runtime_error("payload column shorter than the row range");
            p.pay[i] = payload[i]->ptr;
            p.pay_w[i] = width_of(payload[i]->type);
        }
        p.row_begin = row_begin;
        p.n = n;
        p.log2p = log2_parts;
        p.shift = hash_shift;
        unsigned G = static_cast<unsigned>(ctx->sm_count) * 3;
        const size_t tiles = (n + kPartTile - 1) / kPartTile;
        if (tiles < G) G = static_cast<unsigned>(tiles ? tiles : 1);
        p.rows_per_block = ((tiles + G - 1) / G) * kPartTile;

        bq_col* ok = new_col(ctx, key->type, n);
        bq_col* op[2] = {nullptr, nullptr};
        bq_col* off = nullptr;
        try {
            for (int i = 0; i < n_payload; ++i) op[i] = new_col(ctx, payload[i]->type, n);
            off = new_col(ctx, BQ_INT64, P + 1);
            DevBuf hist(ctx, static_cast<size_t>(P) * G * 4), base(ctx, static_cast<size_t>(P) * G * 8);
            p.hist = hist.as<unsigned>();
            p.base = base.as<unsigned long long>();
            p.out_key = ok->ptr;
            for (int i = 0; i < n_payload; ++i) p.out_pay[i] = op[i]->ptr;
            if (n) {
                k_part_hist<<<G, kBlock, P * 4, ctx->stream>>>(p);
                ctx->launches++;
                BQ_CUDA(cudaGetLastError());
                exclusive_scan_u32(ctx, p.hist, static_cast<size_t>(P) * G, base.as<unsigned long long>());
                const size_t smem = static_cast<size_t>(kPartTile) * 24 + P * 16 + kPartTile * 2;
                BQ_CUDA(cudaFuncSetAttribute(k_part_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
                k_part_scatter<<<G, kBlock, smem, ctx->stream>>>(p);
                ctx->launches++;
                BQ_CUDA(cudaGetLastError());
                k_part_offsets<<<(P + 1 + 255) / 256, 256, 0, ctx->stream>>>(p.base, G, P, n, static_cast<long long*>(off->ptr));
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
            throw;
        }
        *out_key = ok;
        for (int i = 0; i < n_payload; ++i) out_payload[i] = op[i];
        *out_offsets = off;
    });
}
