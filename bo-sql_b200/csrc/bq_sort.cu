// bq_sort.cu — OrderBy / Limit on a device-resident relation.
//
// OrderBy::next materialises every row as two vector<Datum> and runs std::sort with a lexicographic
// comparator (src/exec/operator.cpp:1097-1122, compare_datum :294-317); Limit then copies a prefix
// (:579-613).  Here each sort column becomes an order-preserving uint64 (sign flip for integers, the
// monotone key for doubles, complement for DESC; StrId orders by id like the reference) and a
// row permutation is sorted:
//   n <= 4096 : one rank kernel (rank = number of rows that sort before me, ties by input position);
//   larger    : LSD radix sort, 8 bits per pass, least-significant sort column first; passes whose digit
//               is constant over the input are skipped (an OR-reduction finds the varying bytes).
// Both are stable, so ties keep input order (the reference's std::sort leaves tie order unspecified, H4).
// LIMIT keeps the first k entries of the permutation, so only k rows of payload are gathered.
#include "bq_common.cuh"
#include "bq_internal.cuh"

namespace bq {

BQ_D unsigned long long sort_key(long long raw, int kind, int asc) {
    long long k = key_of(raw, kind);
    unsigned long long u = static_cast<unsigned long long>(k) ^ 0x8000000000000000ULL;
    return asc ? u : ~u;
}

// keys[i] = sort_key(col[perm ? perm[i] : i])
__global__ void __launch_bounds__(kBlock) k_make_keys(const void* __restrict__ col, int kind, int asc,
                                                      const unsigned* __restrict__ perm, size_t n,
                                                      unsigned long long* __restrict__ keys) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = perm ? perm[i] : i;
        keys[i] = sort_key(load_raw(col, kind, r), kind, asc);
    }
}

__global__ void __launch_bounds__(kBlock) k_iota(unsigned* __restrict__ v, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        v[i] = static_cast<unsigned>(i);
}

// ---- small inputs: rank sort over up to 4 key arrays (keys[k*n + i]) -----------------------------
__global__ void __launch_bounds__(kBlock) k_rank_sort(const unsigned long long* __restrict__ keys, int n_keys, size_t n,
                                                      unsigned* __restrict__ perm) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long mine[4];
    for (int k = 0; k < n_keys; ++k) mine[k] = keys[k * n + i];
    unsigned rank = 0;
    for (size_t j = 0; j < n; ++j) {
        int cmp = 0;   // -1: j sorts before i
        for (int k = 0; k < n_keys && cmp == 0; ++k) {
            unsigned long long o = __ldg(keys + k * n + j);
            cmp = o < mine[k] ? -1 : (o > mine[k] ? 1 : 0);
        }
        if (cmp < 0 || (cmp == 0 && j < i)) ++rank;
    }
    perm[rank] = static_cast<unsigned>(i);
}

// ---- top-k (ORDER BY ... LIMIT k over a large input) ----------------------------------------------
// Each 2048-row tile sorts itself and keeps its first k rows; the survivors (k per tile, in tile order, so ties
// still resolve by input position) are reduced again until one tile remains: 100 000 groups -> 49 tiles -> 980
// candidates -> 1 tile.  The reference sorts everything and lets Limit copy a prefix (src/exec/operator.cpp:1115,
// :579-613).  A tile is a bitonic network over POSITIONS in shared memory (66 steps of 1024 compare-exchanges); the
// keys stay where they were loaded and the comparator reads them through the positions - (keys..., position) is a
// total order, which is what makes the network's result the stable order.  (The first generation ranked every row
// against every other row of a 512-row tile: 30 x the work, 94 us per round on Q2's 100 000 groups.)
constexpr int kTopkTile = 2048;

constexpr int kTopkThreads = 1024;      // one compare-exchange per thread per step: the network is a latency chain, so fill the SM

template <int NK>
__global__ void __launch_bounds__(kTopkThreads) k_tile_topk(const unsigned long long* __restrict__ keys, size_t n, unsigned k,
                                                      unsigned* __restrict__ cand /* positions, k per tile */) {
    extern __shared__ unsigned long long topk_smem[];
    unsigned long long* sk = topk_smem;                                       // [NK][kTopkTile]
    unsigned short* idx = reinterpret_cast<unsigned short*>(sk + NK * kTopkTile);
    const size_t tile = blockIdx.x;
    const size_t t0 = tile * kTopkTile;
    const unsigned tn = static_cast<unsigned>((n - t0) < (size_t)kTopkTile ? (n - t0) : (size_t)kTopkTile);
#pragma unroll
    for (int q = 0; q < NK; ++q)
        for (unsigned x = threadIdx.x; x < (unsigned)kTopkTile; x += blockDim.x) sk[q * kTopkTile + x] = x < tn ? keys[q * n + t0 + x] : ~0ULL;
    for (unsigned x = threadIdx.x; x < (unsigned)kTopkTile; x += blockDim.x) idx[x] = static_cast<unsigned short>(x);
    // a before b?  padding positions (>= tn) carry all-ones keys and the highest positions: they sort last
    auto before = [&](unsigned a, unsigned b) {
#pragma unroll
        for (int q = 0; q < NK; ++q) {
            const unsigned long long ka = sk[q * kTopkTile + a], kb = sk[q * kTopkTile + b];
            if (ka != kb) return ka < kb;
        }
        return a < b;
    };
    // a short last tile (and the final round, a few hundred candidates) sorts the smallest power of two that holds it
    unsigned span = 64;
    while (span < tn) span <<= 1;
    for (unsigned size = 2; size <= span; size <<= 1)
        for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (unsigned t = threadIdx.x; t < span / 2; t += blockDim.x) {
                const unsigned pos = 2 * t - (t & (stride - 1));
                const unsigned a = idx[pos], b = idx[pos + stride];
                const bool up = (pos & size) == 0;
                if (up ? before(b, a) : before(a, b)) {
                    idx[pos] = static_cast<unsigned short>(b);
                    idx[pos + stride] = static_cast<unsigned short>(a);
                }
            }
        }
    __syncthreads();
    const unsigned keep = k < tn ? k : tn;
    for (unsigned r = threadIdx.x; r < keep; r += blockDim.x) cand[tile * k + r] = static_cast<unsigned>(t0 + idx[r]);
}

static void launch_tile_topk(bq_ctx* ctx, const unsigned long long* keys, int n_keys, size_t n, unsigned k, unsigned* cand) {
    const unsigned tiles = static_cast<unsigned>((n + kTopkTile - 1) / kTopkTile);
    const size_t smem = static_cast<size_t>(n_keys) * kTopkTile * 8 + kTopkTile * 2;
    switch (n_keys) {
        case 1: k_tile_topk<1><<<tiles, kTopkThreads, smem, ctx->stream>>>(keys, n, k, cand); break;
        case 2: k_tile_topk<2><<<tiles, kTopkThreads, smem, ctx->stream>>>(keys, n, k, cand); break;
        case 3:
            BQ_CUDA(cudaFuncSetAttribute(k_tile_topk<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));      // beyond the 48 KB default
            k_tile_topk<3><<<tiles, kTopkThreads, smem, ctx->stream>>>(keys, n, k, cand); break;
        default:
            BQ_CUDA(cudaFuncSetAttribute(k_tile_topk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            k_tile_topk<4><<<tiles, kTopkThreads, smem, ctx->stream>>>(keys, n, k, cand); break;
    }
}

// out[i] = src[idx[i]] (src == nullptr: identity)
__global__ void __launch_bounds__(kBlock) k_compose(const unsigned* __restrict__ src, const unsigned* __restrict__ idx, size_t n,
                                                    unsigned* __restrict__ out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = src ? src[idx[i]] : idx[i];
}

// ---- radix sort passes ---------------------------------------------------------------------------
constexpr int kRadixRounds = 8;
constexpr int kRadixTile = kBlock * kRadixRounds;

__global__ void __launch_bounds__(kBlock) k_diff_bits(const unsigned long long* __restrict__ keys, size_t n,
                                                      unsigned long long* __restrict__ out) {
    unsigned long long first = keys[0], acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        acc |= keys[i] ^ first;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc |= __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicOr(out, acc);
}

__global__ void __launch_bounds__(kBlock) k_radix_hist(const unsigned long long* __restrict__ keys, size_t n, int shift,
                                                       unsigned* __restrict__ hist, unsigned n_blocks) {
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    size_t base = blockIdx.x * (size_t)kRadixTile;
#pragma unroll
    for (int r = 0; r < kRadixRounds; ++r) {
        size_t i = base + r * kBlock + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * (size_t)n_blocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(kBlock) k_radix_scatter(const unsigned long long* __restrict__ keys_in,
                                                          const unsigned* __restrict__ vals_in,
                                                          unsigned long long* __restrict__ keys_out,
                                                          unsigned* __restrict__ vals_out, size_t n, int shift,
                                                          const unsigned long long* __restrict__ offsets, unsigned n_blocks) {
    __shared__ unsigned running[256];
    __shared__ unsigned wc[kBlock / 32][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    running[threadIdx.x] = 0;
    size_t base = blockIdx.x * (size_t)kRadixTile;
    for (int r = 0; r < kRadixRounds; ++r) {
#pragma unroll
        for (int w = 0; w < kBlock / 32; ++w) wc[w][threadIdx.x] = 0;
        __syncthreads();
        size_t i = base + r * kBlock + threadIdx.x;
        bool valid = i < n;
        unsigned long long key = valid ? keys_in[i] : 0ULL;
        unsigned d = static_cast<unsigned>((key >> shift) & 255u);
        unsigned peers = __match_any_sync(0xffffffffu, valid ? d : 256u + lane);
        unsigned rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank_in_warp == 0) wc[warp][d] = __popc(peers);
        __syncthreads();
        if (valid) {
            unsigned pre = 0;
            for (int w = 0; w < warp; ++w) pre += wc[w][d];
            unsigned long long pos = offsets[d * (size_t)n_blocks + blockIdx.x] + running[d] + pre + rank_in_warp;
            keys_out[pos] = key;
            vals_out[pos] = vals_in[i];
        }
        __syncthreads();
        unsigned add = 0;
#pragma unroll
        for (int w = 0; w < kBlock / 32; ++w) add += wc[w][threadIdx.x];
        running[threadIdx.x] += add;
        __syncthreads();
    }
}

// Stable sort of (keys, vals) by keys ascending; result left in keys/vals (buffers may swap).
static void radix_sort_pairs(bq_ctx* ctx, unsigned long long*& keys, unsigned*& vals, unsigned long long*& keys_alt,
                             unsigned*& vals_alt, size_t n) {
    auto* d_diff = static_cast<unsigned long long*>(scratch(ctx, 16));
    BQ_CUDA(cudaMemsetAsync(d_diff, 0, 8, ctx->stream));
    k_diff_bits<<<grid_for(ctx, n, 8), kBlock, 0, ctx->stream>>>(keys, n, d_diff);
    ctx->launches++;
    auto* h = static_cast<unsigned long long*>(pinned(ctx, 8));
    BQ_CUDA(cudaMemcpyAsync(h, d_diff, 8, cudaMemcpyDeviceToHost, ctx->stream));
    BQ_CUDA(cudaStreamSynchronize(ctx->stream));
    const unsigned long long diff = *h;
    const unsigned n_blocks = static_cast<unsigned>((n + kRadixTile - 1) / kRadixTile);
    DevBuf hist(ctx, 256 * (size_t)n_blocks * 4), offs(ctx, 256 * (size_t)n_blocks * 8);
    for (int byte = 0; byte < 8; ++byte) {
        if (((diff >> (8 * byte)) & 0xFFull) == 0) continue;
        const int shift = 8 * byte;
        k_radix_hist<<<n_blocks, kBlock, 0, ctx->stream>>>(keys, n, shift, static_cast<unsigned*>(hist.p), n_blocks);
        ctx->launches++;
        BQ_CUDA(cudaGetLastError());
        exclusive_scan_u32(ctx, static_cast<unsigned*>(hist.p), 256 * (size_t)n_blocks, static_cast<unsigned long long*>(offs.p));
        k_radix_scatter<<<n_blocks, kBlock, 0, ctx->stream>>>(keys, vals, keys_alt, vals_alt, n, shift,
                                                              static_cast<unsigned long long*>(offs.p), n_blocks);
        ctx->launches++;
        BQ_CUDA(cudaGetLastError());
        std::swap(keys, keys_alt);
        std::swap(vals, vals_alt);
    }
}

}  // namespace bq

using namespace bq;

extern "C" int bq_rel_sort(bq_ctx* ctx, const bq_rel* rel, int n_keys, const int* key_cols, const int* asc,
                           int64_t limit, bq_rel** out) {
    return guarded([&] {
        const size_t n = rel->rows;
        const int nc = static_cast<int>(rel->cols.size());
        if (n_keys < 0 || n_keys > 64) throw std::runtime_error("at most 64 sort keys");
        // the rank kernel and the top-k tiles hold up to 4 keys per row in registers / shared memory; longer key lists
        // (src/exec/operator.cpp:1115-1122 loops over any number) take the LSD radix passes, one sort column at a time
        const bool few_keys = n_keys <= 4;
        for (int k = 0; k < n_keys; ++k)
            if (key_cols[k] < 0 || key_cols[k] >= nc) throw std::runtime_error("sort key column out of range");
        if (n > 0xFFFFFFFFull) throw std::runtime_error("row ids are 32-bit: at most 2^32 rows per sort");
        size_t m = (limit >= 0 && static_cast<size_t>(limit) < n) ? static_cast<size_t>(limit) : n;

        std::vector<bq_col*> cols;
        try {
            if (n == 0 || m == 0) {
                for (int c = 0; c < nc; ++c) cols.push_back(new_col(ctx, rel->cols[c]->type, 0));
            } else {
                DevBuf permA(ctx, n * 4), permB(ctx, n * 4);
                auto* perm = static_cast<unsigned*>(permA.p);
                auto* perm_alt = static_cast<unsigned*>(permB.p);
                if (n_keys == 0) {
                    k_iota<<<grid_for(ctx, n, 8), kBlock, 0, ctx->stream>>>(perm, n);
                    ctx->launches++;
                } else if (n <= 4096 && few_keys) {
                    DevBuf keys(ctx, static_cast<size_t>(n_keys) * n * 8);
                    for (int k = 0; k < n_keys; ++k) {
                        const bq_col* c = rel->cols[key_cols[k]];
                        k_make_keys<<<grid_for(ctx, n, 8), kBlock, 0, ctx->stream>>>(
                            c->ptr, c->type, asc[k], nullptr, n, static_cast<unsigned long long*>(keys.p) + k * n);
                        ctx->launches++;
                    }
                    k_rank_sort<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, ctx->stream>>>(
                        static_cast<unsigned long long*>(keys.p), n_keys, n, perm);
                    ctx->launches++;
                    BQ_CUDA(cudaGetLastError());
                } else if (m <= 256 && few_keys) {
                    // top-k: tiles keep their first m rows until one tile is left
                    const size_t max_cand = ((n + kTopkTile - 1) / kTopkTile) * m;
                    DevBuf keys(ctx, static_cast<size_t>(n_keys) * n * 8), candB(ctx, max_cand * 4), idsA(ctx, max_cand * 4), idsB(ctx, max_cand * 4);
                    auto* kbuf = static_cast<unsigned long long*>(keys.p);
                    auto* cand = static_cast<unsigned*>(candB.p);
                    unsigned* ids = nullptr;                    // current survivors (global row ids); nullptr = all rows
                    auto* ids_next = static_cast<unsigned*>(idsA.p);
                    auto* ids_other = static_cast<unsigned*>(idsB.p);
                    size_t cur = n;
                    while (true) {
                        for (int k = 0; k < n_keys; ++k) {
                            const bq_col* c = rel->cols[key_cols[k]];
                            k_make_keys<<<grid_for(ctx, cur, 8), kBlock, 0, ctx->stream>>>(c->ptr, c->type, asc[k], ids, cur, kbuf + k * cur);
                            ctx->launches++;
                        }
                        const size_t tiles = (cur + kTopkTile - 1) / kTopkTile;
                        launch_tile_topk(ctx, kbuf, n_keys, cur, (unsigned)m, cand);
                        const size_t last = cur - (tiles - 1) * kTopkTile;
                        const size_t next = (tiles - 1) * m + (m < last ? m : last);
                        if (tiles == 1) {        // the last tile's first rows are the answer, already in order
                            k_compose<<<grid_for(ctx, next, 8), kBlock, 0, ctx->stream>>>(ids, cand, next, perm);
                            ctx->launches += 2;
                            BQ_CUDA(cudaGetLastError());
                            break;
                        }
                        k_compose<<<grid_for(ctx, next, 8), kBlock, 0, ctx->stream>>>(ids, cand, next, ids_next);
                        ctx->launches += 2;
                        BQ_CUDA(cudaGetLastError());
                        ids = ids_next;
                        std::swap(ids_next, ids_other);
                        cur = next;
                    }
                } else {
                    DevBuf keysA(ctx, n * 8), keysB(ctx, n * 8);
                    auto* keys = static_cast<unsigned long long*>(keysA.p);
                    auto* keys_alt = static_cast<unsigned long long*>(keysB.p);
                    k_iota<<<grid_for(ctx, n, 8), kBlock, 0, ctx->stream>>>(perm, n);
                    ctx->launches++;
                    for (int k = n_keys - 1; k >= 0; --k) {      // least-significant sort column first
                        const bq_col* c = rel->cols[key_cols[k]];
                        k_make_keys<<<grid_for(ctx, n, 8), kBlock, 0, ctx->stream>>>(c->ptr, c->type, asc[k], perm, n, keys);
                        ctx->launches++;
                        BQ_CUDA(cudaGetLastError());
                        radix_sort_pairs(ctx, keys, perm, keys_alt, perm_alt, n);
                    }
                }
                // gather the first m rows of every column
                bq_col ids;
                ids.ctx = ctx;
                ids.type = BQ_STRING;
                ids.n = m;
                ids.ptr = perm;
                ids.owns = false;
                for (int c = 0; c < nc; ++c) {
                    bq_col* o = nullptr;
                    if (bq_gather(ctx, rel->cols[c], &ids, &o)) throw std::runtime_error(bq_last_error());
                    cols.push_back(o);
                }
            }
            auto* r = new bq_rel();
            r->cols = cols;
            r->rows = (n == 0) ? 0 : m;
            *out = r;
        } catch (...) {
            for (auto* c : cols) free_col(c);
            throw;
        }
    });
}
