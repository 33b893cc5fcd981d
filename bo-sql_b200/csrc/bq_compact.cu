// bq_compact.cu — stable stream compaction and device-wide scans (warp ballot / popc based).
//
// The reference builds `std::vector<size_t> selected` one row at a time (Selection::next,
// src/exec/operator.cpp:410-416) and walks unordered_map buckets to emit groups (:1010-1013).
// Here every producer writes one bit per candidate (ballot of the predicate); this file turns a bit
// vector into the ascending list of set positions:
//   k_bits_block_count : popc per 32-bit word, summed per block of kBlock words
//   k_scan_block_sums  : exclusive scan of the block sums (one CTA, serial over chunks)
//   k_bits_expand      : per word, exclusive offset = block offset + in-block scan, then its bits
// Order is preserved (stable), which Selection and HashJoin need: their output is in scan order.
#include "bq_common.cuh"
#include "bq_internal.cuh"

namespace bq {

__global__ void __launch_bounds__(kBlock) k_bits_block_count(const unsigned* __restrict__ bits, size_t n_words,
                                                             unsigned long long* __restrict__ block_sums) {
    __shared__ unsigned warp_tot[kBlock / 32];
    size_t w = blockIdx.x * (size_t)kBlock + threadIdx.x;
    unsigned c = w < n_words ? __popc(bits[w]) : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int i = 0; i < kBlock / 32; ++i) t += warp_tot[i];
        block_sums[blockIdx.x] = t;
    }
}

// In-place exclusive scan of n values by one CTA; the grand total goes to *total.
__global__ void __launch_bounds__(1024) k_scan_block_sums(unsigned long long* __restrict__ v, size_t n,
                                                          unsigned long long* __restrict__ total) {
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (size_t base = 0; base < n; base += 1024) {
        size_t i = base + threadIdx.x;
        unsigned long long x = i < n ? v[i] : 0ULL;
        unsigned long long incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned long long t = warp_tot[lane];
            unsigned long long ti = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long y = __shfl_up_sync(0xffffffffu, ti, o);
                if (lane >= o) ti += y;
            }
            warp_tot[lane] = ti - t;   // exclusive prefix of the warp totals
        }
        __syncthreads();
        unsigned long long excl = carry + warp_tot[warp] + (incl - x);
        if (i < n) v[i] = excl;
        __syncthreads();
        // chunk total = exclusive prefix of the last element + its value
        if (threadIdx.x == 1023) carry = excl + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(kBlock) k_bits_expand(const unsigned* __restrict__ bits, size_t n_words,
                                                        const unsigned long long* __restrict__ block_offsets,
                                                        unsigned* __restrict__ out, unsigned base_index) {
    __shared__ unsigned warp_tot[kBlock / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    size_t w = blockIdx.x * (size_t)kBlock + threadIdx.x;
    unsigned word = w < n_words ? bits[w] : 0u;
    unsigned c = __popc(word);
    unsigned incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    unsigned wpre = 0;
    for (int i = 0; i < warp; ++i) wpre += warp_tot[i];
    unsigned long long pos = block_offsets[blockIdx.x] + wpre + (incl - c);
    unsigned first = base_index + static_cast<unsigned>(w) * 32u;
    while (word) {
        int b = __ffs(word) - 1;
        out[pos++] = first + b;
        word &= word - 1;
    }
}

// rowids of the set bits of `bits` (n_bits candidates), ascending. Returns the count (host sync).
size_t compact_bits(bq_ctx* ctx, const unsigned* bits, size_t n_bits, unsigned base_index, bq_col** out_rowids,
                    const int* dev_flag, int* host_flag) {
    size_t n_words = (n_bits + 31) / 32;
    if (n_words == 0) {
        *out_rowids = new_col(ctx, BQ_STRING, 0);
        if (dev_flag && host_flag) {
            auto* h = static_cast<int*>(pinned(ctx, 16));
            BQ_CUDA(cudaMemcpyAsync(h, dev_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
            BQ_CUDA(cudaStreamSynchronize(ctx->stream));
            *host_flag = *h;
        }
        return 0;
    }
    size_t n_blocks = (n_words + kBlock - 1) / kBlock;
    DevBuf sums_buf(ctx, (n_blocks + 1) * sizeof(unsigned long long));
    auto* sums = sums_buf.as<unsigned long long>();
    bq_col* ids = nullptr;
    try {
        k_bits_block_count<<<(unsigned)n_blocks, kBlock, 0, ctx->stream>>>(bits, n_words, sums);
        k_scan_block_sums<<<1, 1024, 0, ctx->stream>>>(sums, n_blocks, sums + n_blocks);
        ctx->launches += 2;
        BQ_CUDA(cudaGetLastError());
        auto* h = static_cast<unsigned long long*>(pinned(ctx, 16));
        BQ_CUDA(cudaMemcpyAsync(h, sums + n_blocks, 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (dev_flag) BQ_CUDA(cudaMemcpyAsync(h + 1, dev_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
        BQ_CUDA(cudaStreamSynchronize(ctx->stream));      // the one host round trip: output size (+ the error word)
        size_t total = static_cast<size_t>(*h);
        if (dev_flag && host_flag) *host_flag = static_cast<int>(h[1] & 0xFFFFFFFFull);
        ids = new_col(ctx, BQ_STRING, total);
        if (total) {
            k_bits_expand<<<(unsigned)n_blocks, kBlock, 0, ctx->stream>>>(bits, n_words, sums,
                                                                        static_cast<unsigned*>(ids->ptr), base_index);
            ctx->launches++;
            BQ_CUDA(cudaGetLastError());
        }
        *out_rowids = ids;
        return total;
    } catch (...) {
        free_col(ids);
        throw;
    }
}

// ---- device-wide exclusive scan of uint32 counts into uint64 offsets (join materialisation) -----
__global__ void __launch_bounds__(kBlock) k_u32_block_sum(const unsigned* __restrict__ v, size_t n,
                                                          unsigned long long* __restrict__ block_sums) {
    __shared__ unsigned long long warp_tot[kBlock / 32];
    size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x;
    unsigned long long c = i < n ? v[i] : 0ULL;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int k = 0; k < kBlock / 32; ++k) t += warp_tot[k];
        block_sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kBlock) k_u32_block_scan(const unsigned* __restrict__ v, size_t n,
                                                           const unsigned long long* __restrict__ block_offsets,
                                                           unsigned long long* __restrict__ out) {
    __shared__ unsigned long long warp_tot[kBlock / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x;
    unsigned long long x = i < n ? v[i] : 0ULL;
    unsigned long long incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    unsigned long long wpre = 0;
    for (int k = 0; k < warp; ++k) wpre += warp_tot[k];
    if (i < n) out[i] = block_offsets[blockIdx.x] + wpre + (incl - x);
}

// offsets[i] = sum_{j<i} counts[j]; returns the total (host sync). `offsets` must hold n entries.
size_t exclusive_scan_u32(bq_ctx* ctx, const unsigned* counts, size_t n, unsigned long long* offsets, bool want_total) {
    if (n == 0) return 0;
    size_t n_blocks = (n + kBlock - 1) / kBlock;
    DevBuf sums_buf(ctx, (n_blocks + 1) * sizeof(unsigned long long));
    auto* sums = sums_buf.as<unsigned long long>();
    {
        k_u32_block_sum<<<(unsigned)n_blocks, kBlock, 0, ctx->stream>>>(counts, n, sums);
        k_scan_block_sums<<<1, 1024, 0, ctx->stream>>>(sums, n_blocks, sums + n_blocks);
        k_u32_block_scan<<<(unsigned)n_blocks, kBlock, 0, ctx->stream>>>(counts, n, sums, offsets);
        ctx->launches += 3;
        BQ_CUDA(cudaGetLastError());
        if (!want_total) return 0;           // sums_buf is released in stream order, after the kernels above
        auto* h = static_cast<unsigned long long*>(pinned(ctx, 8));
        BQ_CUDA(cudaMemcpyAsync(h, sums + n_blocks, 8, cudaMemcpyDeviceToHost, ctx->stream));
        BQ_CUDA(cudaStreamSynchronize(ctx->stream));
        return static_cast<size_t>(*h);
    }
}

}  // namespace bq
