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

#include <memory>

namespace bq {

BQ_D unsigned long long sort_key(long long raw, int kind, int asc) {
    long long k = key_of(raw, kind);
    unsigned long long u = static_cast<unsigned long long>(k) ^ 0x8000000000000000ULL;
    return asc ? u : ~u;
}

// keys[i] = sort_key(col[perm ? perm[i] : i])
// neg_zero (optional): set when a DOUBLE column holds -0.0, the one value whose bits the key does not keep (it sorts as +0.0)
__global__ void __launch_bounds__(kBlock) k_make_keys(const void* __restrict__ col, int kind, int asc,
                                                      const unsigned* __restrict__ perm, size_t n,
                                                      unsigned long long* __restrict__ keys, int* __restrict__ neg_zero = nullptr) {
    bool seen = false;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = perm ? perm[i] : i;
        const long long raw = load_raw(col, kind, r);
        seen = seen || (kind == BQ_DOUBLE && raw == INT64_MIN);
        keys[i] = sort_key(raw, kind, asc);
    }
    if (neg_zero && seen) atomicOr(neg_zero, 1);
}

// The inverse: a sorted key array IS the sorted column (unless the column held -0.0), so the first sort column of a full sort
// is written back from its keys in one streaming pass instead of being gathered through the permutation - a random 8-byte
// gather costs a 32-byte sector per row (2.6 ms per 10^8 rows against 0.25 ms).
__global__ void __launch_bounds__(kBlock) k_unmake_keys(const unsigned long long* __restrict__ keys, int kind, int asc, size_t n,
                                                        void* __restrict__ out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long u = asc ? keys[i] : ~keys[i];
        const long long k = static_cast<long long>(u ^ 0x8000000000000000ULL);
        const long long raw = kind == BQ_DOUBLE ? static_cast<long long>(f64_bits_from_key(k)) : k;
        if (kind == BQ_INT64 || kind == BQ_DOUBLE) static_cast<long long*>(out)[i] = raw;
        else if (kind == BQ_STRING) static_cast<unsigned*>(out)[i] = static_cast<unsigned>(raw);
        else static_cast<int*>(out)[i] = static_cast<int>(raw);
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

struct TopkCols {              // the sort columns as they lie: the keys are made inside the kernel
    const void* col[4];
    int kind[4];
    int asc[4];
};

// ids: the rows still in the running (nullptr = all rows, in order); out: the surviving ROW IDS, k per tile, in tile order
template <int NK>
__global__ void __launch_bounds__(kTopkThreads) k_tile_topk(const __grid_constant__ TopkCols cols, const unsigned* __restrict__ ids, size_t n,
                                                      unsigned k, unsigned* __restrict__ out) {
    extern __shared__ unsigned long long topk_smem[];
    unsigned long long* sk = topk_smem;                                       // [NK][kTopkTile]
    unsigned short* idx = reinterpret_cast<unsigned short*>(sk + NK * kTopkTile);
    const size_t tile = blockIdx.x;
    const size_t t0 = tile * kTopkTile;
    const unsigned tn = static_cast<unsigned>((n - t0) < (size_t)kTopkTile ? (n - t0) : (size_t)kTopkTile);
#pragma unroll
    for (int q = 0; q < NK; ++q)
        for (unsigned x = threadIdx.x; x < (unsigned)kTopkTile; x += blockDim.x) {
            unsigned long long key = ~0ULL;
            if (x < tn) {
                const size_t row = ids ? ids[t0 + x] : t0 + x;
                key = sort_key(load_raw(cols.col[q], cols.kind[q], row), cols.kind[q], cols.asc[q]);
            }
            sk[q * kTopkTile + x] = key;
        }
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
                const bool up = (pos & size) == 0;
                if (NK == 1) {
                    // one sort column: the key travels with its position, so a step reads two (key, position) pairs at
                    // regular addresses instead of chasing positions into the key array
                    const unsigned long long ka = sk[pos], kb = sk[pos + stride];
                    const unsigned short ia = idx[pos], ib = idx[pos + stride];
                    const bool b_first = kb < ka || (kb == ka && ib < ia);
                    if (up ? b_first : !b_first) {
                        sk[pos] = kb;
                        sk[pos + stride] = ka;
                        idx[pos] = ib;
                        idx[pos + stride] = ia;
                    }
                } else {
                    const unsigned a = idx[pos], b = idx[pos + stride];
                    if (up ? before(b, a) : before(a, b)) {
                        idx[pos] = static_cast<unsigned short>(b);
                        idx[pos + stride] = static_cast<unsigned short>(a);
                    }
                }
            }
        }
    __syncthreads();
    const unsigned keep = k < tn ? k : tn;
    for (unsigned r = threadIdx.x; r < keep; r += blockDim.x) out[tile * k + r] = ids ? ids[t0 + idx[r]] : static_cast<unsigned>(t0 + idx[r]);
}

static void launch_tile_topk(bq_ctx* ctx, const TopkCols& cols, int n_keys, const unsigned* ids, size_t n, unsigned k, unsigned* out) {
    const unsigned tiles = static_cast<unsigned>((n + kTopkTile - 1) / kTopkTile);
    const size_t smem = static_cast<size_t>(n_keys) * kTopkTile * 8 + kTopkTile * 2;
    switch (n_keys) {
        case 1: k_tile_topk<1><<<tiles, kTopkThreads, smem, ctx->stream>>>(cols, ids, n, k, out); break;
        case 2: k_tile_topk<2><<<tiles, kTopkThreads, smem, ctx->stream>>>(cols, ids, n, k, out); break;
        case 3:
            BQ_CUDA(cudaFuncSetAttribute(k_tile_topk<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));      // beyond the 48 KB default
            k_tile_topk<3><<<tiles, kTopkThreads, smem, ctx->stream>>>(cols, ids, n, k, out); break;
        default:
            BQ_CUDA(cudaFuncSetAttribute(k_tile_topk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            k_tile_topk<4><<<tiles, kTopkThreads, smem, ctx->stream>>>(cols, ids, n, k, out); break;
    }
    ctx->launches++;
    BQ_CUDA(cudaGetLastError());
}

// ---- radix sort passes ---------------------------------------------------------------------------
// LSD, 8 bits per pass, (uint64 key, uint32 row id) pairs, stable.  One pass = k_radix_hist (digit counts per 4096-key
// tile, 8 B/row read) -> exclusive scan of the [digit][tile] counts -> k_radix_scatter (12 B/row read, 12 B/row written).
// The scatter sorts its tile by digit INSIDE shared memory first and then writes it out in that order, so neighbouring
// threads store to neighbouring addresses of the same digit's run (runs average 16 keys = 128 B for random digits); the
// first generation stored every key straight from the thread that had loaded it - 32 different runs per warp store.
// Ranking inside the tile is match-based and needs no block-wide barrier per round: a warp owns 512 consecutive keys and
// keeps its own running count per digit (one writer per counter, no atomics); (warp, round, lane) IS the input order, so
// digit_base[d] + the earlier warps' counts + the running count + rank-in-group is the stable position.
constexpr int kSortItems = 16;                        // keys per thread
constexpr int kSortTile = kBlock * kSortItems;        // 4096 keys per CTA
constexpr size_t kSortSmem = static_cast<size_t>(kSortTile) * 12;      // the tile's keys and row ids in digit order: 48 KB
#ifndef BQ_SCAT_THREADS
#define BQ_SCAT_THREADS 512
#define BQ_SCAT_BLOCKS 3
#endif
constexpr int kScatThreads = BQ_SCAT_THREADS;         // scatter CTA: 512 threads x 8 keys, three CTAs per SM (measured: 0.87 ms per pass; 256 x 16 x 3: 0.99, 1024 x 4 x 2: 0.94)
constexpr int kScatItems = kSortTile / kScatThreads;
constexpr int kScatBlocks = BQ_SCAT_BLOCKS;

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
                                                       unsigned* __restrict__ hist /* [256][n_tiles] */, unsigned n_tiles) {
    __shared__ unsigned h[kBlock / 32][256];          // one copy per warp: same-digit collisions stay inside a warp
    const int warp = threadIdx.x >> 5;
#pragma unroll
    for (int w = 0; w < kBlock / 32; ++w) h[w][threadIdx.x] = 0;
    __syncthreads();
    const size_t base = blockIdx.x * (size_t)kSortTile;
    unsigned long long k[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const size_t i = base + r * kBlock + threadIdx.x;
        k[r] = i < n ? keys[i] : 0ULL;
    }
#pragma unroll
    for (int r = 0; r < kSortItems; ++r)
        if (base + r * kBlock + threadIdx.x < n) atomicAdd(&h[warp][(k[r] >> shift) & 255u], 1u);
    __syncthreads();
    unsigned t = 0;
#pragma unroll
    for (int w = 0; w < kBlock / 32; ++w) t += h[w][threadIdx.x];
    hist[threadIdx.x * (size_t)n_tiles + blockIdx.x] = t;
}

__global__ void __launch_bounds__(kScatThreads, kScatBlocks) k_radix_scatter(const unsigned long long* __restrict__ keys_in,
                                                             const unsigned* __restrict__ vals_in,
                                                             unsigned long long* __restrict__ keys_out,
                                                             unsigned* __restrict__ vals_out, size_t n, int shift,
                                                             const unsigned long long* __restrict__ offsets /* [256][n_tiles] */,
                                                             unsigned n_tiles) {
    extern __shared__ __align__(16) unsigned char sort_smem[];
    auto* skey = reinterpret_cast<unsigned long long*>(sort_smem);                          // tile in digit order
    auto* sval = reinterpret_cast<unsigned*>(sort_smem + kSortTile * 8);
    __shared__ unsigned short cnt[kScatThreads / 32][256];                                        // per warp: keys of digit d seen so far
    __shared__ unsigned digit_base[256];
    __shared__ unsigned long long goff[256];                                                // global offset of a digit's run minus its tile-local base
    __shared__ unsigned warp_tot[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t base = blockIdx.x * (size_t)kSortTile;
    const unsigned tile_n = static_cast<unsigned>((n - base) < (size_t)kSortTile ? (n - base) : (size_t)kSortTile);
    constexpr unsigned kWarpKeys = kScatItems * 32;                                         // a warp owns 512 consecutive keys of the tile

    for (int i = threadIdx.x; i < (kScatThreads / 32) * 256; i += kScatThreads) (&cnt[0][0])[i] = 0;
    unsigned long long k[kScatItems];
#pragma unroll
    for (int r = 0; r < kScatItems; ++r) {
        const unsigned x = warp * kWarpKeys + r * 32 + lane;
        k[r] = x < tile_n ? keys_in[base + x] : 0ULL;
    }
    __syncthreads();
    // Ranking.  Input order inside the tile is (warp, round, lane); a warp walks its rounds in order, so the running count
    // of digit d in cnt[warp][d] is touched by one warp only - no block-wide barrier per round.  Lanes holding the same
    // digit find each other with match.any; each of them reads the running count, the first of them adds the group size.
    unsigned short off_in_warp[kScatItems];
    unsigned peers[kScatItems];
    // all sixteen matches first: they do not depend on each other, and match.any has a long latency (ncu: a third of the
    // kernel's stall samples sat on its consumer when each round waited for its own match)
#pragma unroll
    for (int r = 0; r < kScatItems; ++r) {
        const bool valid = warp * kWarpKeys + r * 32 + lane < tile_n;
        const unsigned d = static_cast<unsigned>((k[r] >> shift) & 255u);
        peers[r] = __match_any_sync(0xffffffffu, valid ? d : 256u + lane);
    }
#pragma unroll
    for (int r = 0; r < kScatItems; ++r) {
        const bool valid = warp * kWarpKeys + r * 32 + lane < tile_n;
        const unsigned d = static_cast<unsigned>((k[r] >> shift) & 255u);
        const unsigned rk = __popc(peers[r] & ((1u << lane) - 1u));
        // only the first lane of a group touches the counter (one shared-memory access per distinct digit); the others
        // get the old value by shuffle
        unsigned before = 0;
        if (valid && rk == 0) {
            before = cnt[warp][d];
            cnt[warp][d] = static_cast<unsigned short>(before + __popc(peers[r]));
        }
        before = __shfl_sync(0xffffffffu, before, __ffs(peers[r]) - 1);
        __syncwarp();          // the next round's leaders read what this round's leaders wrote
        off_in_warp[r] = static_cast<unsigned short>(before + rk);
    }
    __syncthreads();
    {   // thread d (the first 256 threads): digit d's counts over the warps -> exclusive offsets per warp; then the digits' bases
        const bool owner = threadIdx.x < 256;
        unsigned run = 0;
        if (owner) {
#pragma unroll
            for (int w = 0; w < kScatThreads / 32; ++w) {
                const unsigned c = cnt[w][threadIdx.x];
                cnt[w][threadIdx.x] = static_cast<unsigned short>(run);
                run += c;
            }
        }
        unsigned incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (owner && lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (owner) {
            unsigned pre = 0;
            for (int w = 0; w < warp; ++w) pre += warp_tot[w];
            const unsigned excl = pre + incl - run;
            digit_base[threadIdx.x] = excl;
            goff[threadIdx.x] = offsets[threadIdx.x * (size_t)n_tiles + blockIdx.x] - excl;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kScatItems; ++r) {
        const unsigned x = warp * kWarpKeys + r * 32 + lane;
        if (x < tile_n) {
            const unsigned d = static_cast<unsigned>((k[r] >> shift) & 255u);
            const unsigned pos = digit_base[d] + cnt[warp][d] + off_in_warp[r];
            skey[pos] = k[r];
            sval[pos] = vals_in[base + x];
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kScatItems; ++r) {
        const unsigned j = r * kScatThreads + threadIdx.x;
        if (j < tile_n) {
            const unsigned long long key = skey[j];
            const unsigned long long g = goff[(key >> shift) & 255u] + j;
            keys_out[g] = key;
            vals_out[g] = sval[j];
        }
    }
}

// Stable sort of (keys, vals) by keys ascending; result left in keys/vals (buffers may swap).  One host round trip per
// sort: which of the eight key bytes vary at all (constant bytes are skipped - an int32-ranged key takes four passes).
static void radix_sort_pairs(bq_ctx* ctx, unsigned long long*& keys, unsigned*& vals, unsigned long long*& keys_alt,
                             unsigned*& vals_alt, size_t n, const int* dev_flag = nullptr, int* host_flag = nullptr) {
    auto* d_diff = static_cast<unsigned long long*>(scratch(ctx, 16));
    BQ_CUDA(cudaMemsetAsync(d_diff, 0, 8, ctx->stream));
    k_diff_bits<<<grid_for(ctx, n, 8), kBlock, 0, ctx->stream>>>(keys, n, d_diff);
    ctx->launches++;
    auto* h = static_cast<unsigned long long*>(pinned(ctx, 16));
    BQ_CUDA(cudaMemcpyAsync(h, d_diff, 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (dev_flag) BQ_CUDA(cudaMemcpyAsync(h + 1, dev_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    BQ_CUDA(cudaStreamSynchronize(ctx->stream));
    const unsigned long long diff = *h;
    if (dev_flag && host_flag) *host_flag = static_cast<int>(h[1] & 0xFFFFFFFFull);
    const unsigned n_tiles = static_cast<unsigned>((n + kSortTile - 1) / kSortTile);
    DevBuf hist(ctx, 256 * (size_t)n_tiles * 4), offs(ctx, 256 * (size_t)n_tiles * 8);
    BQ_CUDA(cudaFuncSetAttribute(k_radix_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSortSmem)));
    for (int byte = 0; byte < 8; ++byte) {
        if (((diff >> (8 * byte)) & 0xFFull) == 0) continue;
        const int shift = 8 * byte;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        if (ctx->profile) {
            BQ_CUDA(cudaEventCreate(&ev0));
            BQ_CUDA(cudaEventCreate(&ev1));
            BQ_CUDA(cudaEventRecord(ev0, ctx->stream));
        }
        k_radix_hist<<<n_tiles, kBlock, 0, ctx->stream>>>(keys, n, shift, static_cast<unsigned*>(hist.p), n_tiles);
        ctx->launches++;
        BQ_CUDA(cudaGetLastError());
        exclusive_scan_u32(ctx, static_cast<unsigned*>(hist.p), 256 * (size_t)n_tiles, static_cast<unsigned long long*>(offs.p), false);
        k_radix_scatter<<<n_tiles, kScatThreads, kSortSmem, ctx->stream>>>(keys, vals, keys_alt, vals_alt, n, shift,
                                                                   static_cast<unsigned long long*>(offs.p), n_tiles);
        ctx->launches++;
        BQ_CUDA(cudaGetLastError());
        if (ctx->profile) {
            BQ_CUDA(cudaEventRecord(ev1, ctx->stream));
            ctx->profile_events.emplace_back(ev0, ev1);
        }
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
                std::unique_ptr<DevBuf> sort_keys_a, sort_keys_b;      // a full sort's key buffers outlive the sort: see k_unmake_keys
                const unsigned long long* first_key_sorted = nullptr;
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
                    // top-k: tiles keep their first m rows until one tile is left (one launch per round: keys are made from
                    // the columns inside the kernel and the survivors leave as row ids)
                    const size_t max_cand = ((n + kTopkTile - 1) / kTopkTile) * m;
                    DevBuf idsA(ctx, max_cand * 4), idsB(ctx, max_cand * 4);
                    TopkCols tc{};
                    for (int k = 0; k < n_keys; ++k) {
                        const bq_col* c = rel->cols[key_cols[k]];
                        tc.col[k] = c->ptr;
                        tc.kind[k] = c->type;
                        tc.asc[k] = asc[k];
                    }
                    unsigned* ids = nullptr;                    // current survivors (global row ids); nullptr = all rows
                    auto* ids_next = static_cast<unsigned*>(idsA.p);
                    auto* ids_other = static_cast<unsigned*>(idsB.p);
                    size_t cur = n;
                    while (true) {
                        const size_t tiles = (cur + kTopkTile - 1) / kTopkTile;
                        const size_t last = cur - (tiles - 1) * kTopkTile;
                        const size_t next = (tiles - 1) * m + (m < last ? m : last);
                        if (tiles == 1) {        // the last tile's first rows are the answer, already in order
                            launch_tile_topk(ctx, tc, n_keys, ids, cur, (unsigned)m, perm);
                            break;
                        }
                        launch_tile_topk(ctx, tc, n_keys, ids, cur, (unsigned)m, ids_next);
                        ids = ids_next;
                        std::swap(ids_next, ids_other);
                        cur = next;
                    }
                } else {
                    sort_keys_a = std::make_unique<DevBuf>(ctx, n * 8);
                    sort_keys_b = std::make_unique<DevBuf>(ctx, n * 8);
                    auto* keys = static_cast<unsigned long long*>(sort_keys_a->p);
                    auto* keys_alt = static_cast<unsigned long long*>(sort_keys_b->p);
                    k_iota<<<grid_for(ctx, n, 8), kBlock, 0, ctx->stream>>>(perm, n);
                    ctx->launches++;
                    DevBuf flag_buf(ctx, 16);
                    int neg_zero = 0;
                    for (int k = n_keys - 1; k >= 0; --k) {      // least-significant sort column first
                        const bq_col* c = rel->cols[key_cols[k]];
                        if (k == 0) BQ_CUDA(cudaMemsetAsync(flag_buf.p, 0, 4, ctx->stream));
                        // (the first column sorted reads its rows in place: the permutation is still the identity)
                        k_make_keys<<<grid_for(ctx, n, 8), kBlock, 0, ctx->stream>>>(c->ptr, c->type, asc[k], k == n_keys - 1 ? nullptr : perm, n, keys,
                                                                                     k == 0 ? flag_buf.as<int>() : nullptr);
                        ctx->launches++;
                        BQ_CUDA(cudaGetLastError());
                        radix_sort_pairs(ctx, keys, perm, keys_alt, perm_alt, n, k == 0 ? flag_buf.as<int>() : nullptr, &neg_zero);
                    }
                    if (!neg_zero) first_key_sorted = keys;
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
                    if (first_key_sorted && c == key_cols[0]) {
                        o = new_col(ctx, rel->cols[c]->type, m);
                        cols.push_back(o);
                        k_unmake_keys<<<grid_for(ctx, m, 8), kBlock, 0, ctx->stream>>>(first_key_sorted, rel->cols[c]->type, asc[0], m, o->ptr);
                        ctx->launches++;
                        BQ_CUDA(cudaGetLastError());
                        continue;
                    }
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
