// bq_join.cu — hash-join build and the materialising probe.
//
// HashJoin::open drains the right child into unordered_map<Key, vector<row_id>> plus a copy of every
// build row (src/exec/operator.cpp:739-762, ~250 B per build row).  Here the build side stays where it is
// (device columns) and the table holds only what a probe needs, chosen from catalog statistics
// (include/catalog/catalog.h:16-21: min/max/ndv, row_count):
//   BITMAP  unique dense key, no build column read downstream: one bit per key of [min,max]
//           (Q2: 250 M orders -> 31 MB, resident in the 126 MB L2 for the whole probe)
//   DIRECT  unique dense key: uint32 build row id per key of [min,max]
//   HASH    anything else: open addressing in HBM, linear probing, capacity 2x rows (power of two),
//           slot = (key, build row id+1) claimed with atomicCAS; duplicate keys occupy separate slots.
// Build-side predicate ranges are applied before insertion (filter push-down below an inner join keeps
// the joined row set of the reference, which filters above the join, src/logical/planner.cpp:110-117).
//
// The fused probe lives in bq_scan.cu; this file also has the materialising probe that HashJoin::next
// needs when its rows are consumed as rows: count matches per probe row, exclusive scan, fill pairs in
// probe order with each row's matches in build insertion order (src/exec/operator.cpp:802-816).
#include "bq_common.cuh"
#include "bq_internal.cuh"

#include <algorithm>
#include <cstdlib>

namespace bq {

struct BuildParams {
    const void* key;
    int key_kind;
    DSlot s[3];
    const long long* mask;
    size_t row_begin, row_end;
    long long key_min;
    unsigned long long domain;
    unsigned* bitmap;
    unsigned* direct;
    JoinSlot* h_slots;
    unsigned* h_occ;
    unsigned long long h_mask;
    unsigned long long* n_inserted;
    int* flags;   // 1 = duplicate key seen (BITMAP/DIRECT), 2 = key outside [min,max], 4 = table full
};

BQ_D bool canon_join_key(long long& k, int kind) {
    if (kind == BQ_DOUBLE) {
        if (k == INT64_MIN) k = 0;                                             // -0.0 == 0.0 (KeyEqual, src/exec/operator.cpp:657)
        if ((k & 0x7FFFFFFFFFFFFFFFLL) > 0x7FF0000000000000LL) return false;   // NaN never equals anything
    }
    return true;
}

// lo <= key <= hi as one unsigned compare; the host hands the device clamped, non-empty ranges (normalise_slot)
BQ_D bool fast_pass(const DSlot& s, long long raw) {
    const unsigned long long x = static_cast<unsigned long long>(key_of(raw, s.kind));
    bool ok = ((x - static_cast<unsigned long long>(s.lo0)) <= static_cast<unsigned long long>(s.hi0 - s.lo0)) != (s.neg0 != 0);
    if (s.nr > 1) ok = ok && (((x - static_cast<unsigned long long>(s.lo1)) <= static_cast<unsigned long long>(s.hi1 - s.lo1)) != (s.neg1 != 0));
    return ok;
}

__global__ void __launch_bounds__(kBlock) k_popcount_words(const unsigned* __restrict__ w, size_t n, unsigned long long* __restrict__ out) {
    unsigned long long c = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) c += __popc(w[i]);
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// KEYK: key kind known at compile time (-1: read it from the parameters); NPRED: number of predicate slots in use
// (-1: test all three at run time).  A warp takes 128 consecutive rows per trip, four per lane, with all loads of the trip
// issued before the first use.  BITMAP inserts are reductions (red.or: nothing comes back, so no lane waits on the L2);
// duplicate keys are found afterwards by comparing the bitmap's popcount with the rows inserted (build_kind).
template <int KIND, int KEYK, int NPRED>
__global__ void __launch_bounds__(kBlock) k_join_build(const __grid_constant__ BuildParams p) {
    unsigned long long local = 0;
    const size_t n = p.row_end - p.row_begin;
    const int lane = threadIdx.x & 31;
    const int key_kind = KEYK >= 0 ? KEYK : p.key_kind;
    constexpr int R = 4;
    const size_t warps = static_cast<size_t>(gridDim.x) * (kBlock / 32);
    const size_t warp = static_cast<size_t>(blockIdx.x) * (kBlock / 32) + (threadIdx.x >> 5);
    const size_t n_chunks = (n + 32 * R - 1) / (32 * R);
    for (size_t c = warp; c < n_chunks; c += warps) {
        long long k[R];
        bool ok[R];
        size_t row[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const size_t t = c * (32 * R) + 32 * r + lane;
            ok[r] = t < n;
            row[r] = p.row_begin + (ok[r] ? t : 0);
            k[r] = load_raw(p.key, key_kind, row[r]);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const size_t i = row[r];
            if (NPRED < 0) {
#pragma unroll
                for (int s = 0; s < 3; ++s)
                    if (p.s[s].ptr && p.s[s].nr) ok[r] = ok[r] && fast_pass(p.s[s], load_raw(p.s[s].ptr, p.s[s].kind, i));
                if (p.mask && __ldg(p.mask + i) == 0) ok[r] = false;
            } else {
#pragma unroll
                for (int s = 0; s < 3; ++s)
                    if (s < NPRED) ok[r] = ok[r] && fast_pass(p.s[s], load_raw(p.s[s].ptr, p.s[s].kind, i));
            }
            ok[r] = ok[r] && canon_join_key(k[r], key_kind);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const size_t i = row[r];
            if (KIND == BQ_JOIN_AUTO) {
                local += ok[r] ? 1 : 0;          // count only: how many rows the build-side predicates let through (sizes the hash table)
            } else if (KIND == BQ_JOIN_BITMAP) {
                unsigned long long idx = static_cast<unsigned long long>(k[r] - p.key_min);
                if (ok[r] && idx >= p.domain) {
                    atomicOr(p.flags, 2);
                    ok[r] = false;
                }
                // Build keys usually arrive clustered (o.order_id = row + 1): when every inserting lane of the warp hits
                // the same bitmap word, the warp combines its bits into ONE reduction instead of up to 32 on one address.
                const unsigned word = static_cast<unsigned>(idx >> 5);
                const unsigned bit = ok[r] ? 1u << (idx & 31) : 0u;
                const unsigned active = __ballot_sync(0xffffffffu, ok[r]);
                if (active) {
                    const int first = __ffs(active) - 1;
                    const unsigned w0 = __shfl_sync(0xffffffffu, word, first);
                    const bool same = __all_sync(0xffffffffu, !ok[r] || word == w0);
                    if (same) {
                        const unsigned bits = __reduce_or_sync(0xffffffffu, bit);
                        if (lane == first) atomicOr(p.bitmap + w0, bits);
                    } else if (ok[r]) {
                        atomicOr(p.bitmap + word, bit);
                    }
                    local += ok[r] ? 1 : 0;
                }
            } else if (KIND == BQ_JOIN_DIRECT) {
                if (ok[r]) {
                    unsigned long long idx = static_cast<unsigned long long>(k[r] - p.key_min);
                    if (idx >= p.domain) {
                        atomicOr(p.flags, 2);
                    } else {
                        unsigned old = atomicCAS(p.direct + idx, 0u, static_cast<unsigned>(i) + 1u);
                        if (old) atomicOr(p.flags, 1);
                        local++;
                    }
                }
            } else if (ok[r]) {
                // open addressing over 16-byte slots {key, build row + 1}: the row word is claimed, then the key is stored
                unsigned long long h = key_hash(static_cast<uint64_t>(k[r])) & p.h_mask;
                bool placed = false;
                for (unsigned long long probes = 0; probes <= p.h_mask; ++probes) {
                    unsigned old = atomicCAS(&p.h_slots[h].row, 0u, static_cast<unsigned>(i) + 1u);
                    if (old == 0u) {
                        p.h_slots[h].key = k[r];
                        if (p.h_occ) atomicOr(p.h_occ + (h >> 5), 1u << (h & 31));
                        placed = true;
                        break;
                    }
                    h = (h + 1) & p.h_mask;
                }
                if (!placed) atomicOr(p.flags, 4);
                local++;
            }
        }
    }
    local = warp_sum(local);
    if (lane == 0 && local) atomicAdd(p.n_inserted, local);
}

// Bitmap build for the common shape (INT64 key, at most one predicate column, no mask; row_begin a multiple of 8).
// A lane owns EIGHT CONSECUTIVE rows of a 256-row chunk, so a chunk is four 128-bit key loads and two (4-byte predicate
// column) or four (8-byte) 128-bit predicate loads per lane, all issued before the first use.  The eight bits of a lane
// almost always fall into one bitmap word; lanes whose neighbours hit the same word fold their bits with five
// shuffle steps (a segmented OR: folding is only ever done between lanes that name the SAME word, so it is correct for
// any key order and optimal for clustered keys such as o.order_id = row + 1), and only the first lane of each run issues
// the reduction: 8 red.or per 256 sequential rows instead of one per inserted row.  Nothing comes back from the L2, so no
// lane waits on it; duplicates are found afterwards by popcount (build_kind).
// PW: width of the predicate column in bytes (0 = no predicate).
template <int PW>
__global__ void __launch_bounds__(kBlock) k_bitmap_build4(const __grid_constant__ BuildParams p) {
    constexpr int R = 8;                       // consecutive rows per lane: 64 B of keys + 32 / 64 B of predicate values in flight
    unsigned long long local = 0;
    const size_t n = p.row_end - p.row_begin;
    const int lane = threadIdx.x & 31;
    const size_t warps = static_cast<size_t>(gridDim.x) * (kBlock / 32);
    const size_t warp = static_cast<size_t>(blockIdx.x) * (kBlock / 32) + (threadIdx.x >> 5);
    const size_t n_chunks = (n + 32 * R - 1) / (32 * R);
    const DSlot& s0 = p.s[0];
    bool out_of_domain = false;
    auto i64_of = [](int lo, int hi) { return static_cast<long long>((static_cast<unsigned long long>(static_cast<unsigned>(hi)) << 32) | static_cast<unsigned>(lo)); };
    for (size_t c = warp; c < n_chunks; c += warps) {
        const size_t t0 = c * (32 * R) + R * static_cast<size_t>(lane);       // first of this lane's rows, relative to row_begin
        const size_t i0 = p.row_begin + t0;
        long long k[R];
        long long pv[R];
        bool ok[R];
        if (t0 + R <= n) {
            const int4* kp = reinterpret_cast<const int4*>(static_cast<const long long*>(p.key) + i0);
            int4 kv[R / 2];
#pragma unroll
            for (int v = 0; v < R / 2; ++v) kv[v] = ldg_stream(kp + v);
            if (PW == 4) {
                const int4* qp = reinterpret_cast<const int4*>(static_cast<const int*>(s0.ptr) + i0);
                int4 q[R / 4];
#pragma unroll
                for (int v = 0; v < R / 4; ++v) q[v] = ldg_stream(qp + v);
#pragma unroll
                for (int v = 0; v < R / 4; ++v) {
                    const int w4[4] = {q[v].x, q[v].y, q[v].z, q[v].w};
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        pv[4 * v + r] = s0.kind == BQ_STRING ? static_cast<long long>(static_cast<unsigned>(w4[r])) : static_cast<long long>(w4[r]);
                }
            } else if (PW == 8) {
                const int4* qp = reinterpret_cast<const int4*>(static_cast<const long long*>(s0.ptr) + i0);
                int4 q[R / 2];
#pragma unroll
                for (int v = 0; v < R / 2; ++v) q[v] = ldg_stream(qp + v);
#pragma unroll
                for (int v = 0; v < R / 2; ++v) {
                    pv[2 * v] = i64_of(q[v].x, q[v].y);
                    pv[2 * v + 1] = i64_of(q[v].z, q[v].w);
                }
            }
#pragma unroll
            for (int v = 0; v < R / 2; ++v) {
                k[2 * v] = i64_of(kv[v].x, kv[v].y);
                k[2 * v + 1] = i64_of(kv[v].z, kv[v].w);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) ok[r] = true;
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                ok[r] = t0 + r < n;
                const size_t i = ok[r] ? i0 + r : p.row_begin;
                k[r] = __ldg(static_cast<const long long*>(p.key) + i);
                if (PW) pv[r] = load_raw(s0.ptr, s0.kind, i);
            }
        }
        unsigned w[R], b[R];
        bool one_word = true;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (PW) ok[r] = ok[r] && fast_pass(s0, pv[r]);
            const unsigned long long idx = static_cast<unsigned long long>(k[r] - p.key_min);
            if (ok[r] && idx >= p.domain) {
                out_of_domain = true;
                ok[r] = false;
            }
            w[r] = static_cast<unsigned>(idx >> 5);
            b[r] = ok[r] ? 1u << (idx & 31) : 0u;
            local += ok[r] ? 1 : 0;
            one_word = one_word && w[r] == w[0];
        }
        // fold inside the lane, then across lanes that name the same word
        unsigned W = 0xF0000000u | static_cast<unsigned>(lane), B = 0;      // a word index no bitmap has (domain <= 2^32 keys)
        if (one_word) {
            W = w[0];
#pragma unroll
            for (int r = 0; r < R; ++r) B |= b[r];
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (b[r]) atomicOr(p.bitmap + w[r], b[r]);
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned ow = __shfl_down_sync(0xffffffffu, W, d);
            const unsigned ob = __shfl_down_sync(0xffffffffu, B, d);
            if (lane + d < 32 && ow == W) B |= ob;
        }
        const unsigned pw = __shfl_up_sync(0xffffffffu, W, 1);
        if ((lane == 0 || pw != W) && B) atomicOr(p.bitmap + W, B);
    }
    if (out_of_domain) atomicOr(p.flags, 2);
    local = warp_sum(local);
    if (lane == 0 && local) atomicAdd(p.n_inserted, local);
}

// ---- materialising probe ---------------------------------------------------------------------------
struct ProbeParams {
    const void* key;
    int key_kind;
    const unsigned* rowids;   // optional probe row subset (in order)
    size_t row_begin;
    size_t n;                 // probe items
    int jkind;
    long long key_min;
    unsigned long long domain;
    const unsigned* bitmap;
    const unsigned* direct;
    const JoinSlot* h_slots;
    unsigned long long h_mask;
};

BQ_D unsigned probe_matches(const ProbeParams& p, long long k, unsigned* out, unsigned cap_out) {
    // returns the number of matches; writes up to cap_out build row ids to `out` when out != nullptr
    if (!canon_join_key(k, p.key_kind)) return 0;
    if (p.jkind == BQ_JOIN_DIRECT) {
        unsigned long long idx = static_cast<unsigned long long>(k - p.key_min);
        if (idx >= p.domain) return 0;
        unsigned e = __ldg(p.direct + idx);
        if (!e) return 0;
        if (out && cap_out) out[0] = e - 1;
        return 1;
    }
    unsigned m = 0;
    unsigned long long h = key_hash(static_cast<uint64_t>(k)) & p.h_mask;
    for (unsigned long long probes = 0; probes <= p.h_mask; ++probes) {
        const JoinSlot e = load_join_slot(p.h_slots + h);         // one 16-byte load: key and row together
        if (!e.row) break;
        if (e.key == k) {
            if (out && m < cap_out) out[m] = e.row - 1;
            ++m;
        }
        h = (h + 1) & p.h_mask;
    }
    return m;
}

__global__ void __launch_bounds__(kBlock) k_probe_count(const __grid_constant__ ProbeParams p, unsigned* __restrict__ counts) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= p.n) return;
    size_t i = p.rowids ? p.rowids[t] : p.row_begin + t;
    counts[t] = probe_matches(p, load_raw(p.key, p.key_kind, i), nullptr, 0);
}

__global__ void __launch_bounds__(kBlock) k_probe_fill(const __grid_constant__ ProbeParams p,
                                                       const unsigned* __restrict__ counts,
                                                       const unsigned long long* __restrict__ offsets,
                                                       unsigned* __restrict__ out_probe, unsigned* __restrict__ out_build) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= p.n) return;
    unsigned m = counts[t];
    if (!m) return;
    size_t i = p.rowids ? p.rowids[t] : p.row_begin + t;
    unsigned* dst = out_build + offsets[t];
    probe_matches(p, load_raw(p.key, p.key_kind, i), dst, m);
    // build insertion order == ascending build row id (the build side is inserted in scan order)
    for (unsigned a = 1; a < m; ++a) {
        unsigned v = dst[a];
        unsigned b = a;
        while (b > 0 && dst[b - 1] > v) {
            dst[b] = dst[b - 1];
            --b;
        }
        dst[b] = v;
    }
    unsigned* pr = out_probe + offsets[t];
    for (unsigned a = 0; a < m; ++a) pr[a] = static_cast<unsigned>(i);
}

// ---- semi-join probe in key-range passes (bitmaps beyond L2) ----------------------------------------------------------
struct ProbeBitsParams {
    const void* key;
    int key_kind;
    size_t row_begin, row_end;
    long long key_min;
    unsigned long long domain;
    unsigned long long slice_lo, slice_len;      // this pass tests keys with slice_lo <= key - key_min < slice_lo + slice_len
    const unsigned* bitmap;
    unsigned* out;                               // bit i = row i
    unsigned* idx32;                             // key - key_min per row (0xFFFFFFFF = outside the domain): written by the
                                                 // first pass, read by the later ones (4 bytes per row instead of 8)
    int first;                                   // first pass stores every word, later passes OR into the words they hit
};

// A warp owns 32 * ROWS consecutive rows per trip (row_begin is a multiple of 128): lane t tests rows t, 32+t, 64+t, ..., so
// the key column is read with fully coalesced requests, ROWS of them in flight per lane before the first use, and each
// ballot is one finished word of the output; later passes merge their words with red.or (no read-modify-write on the
// critical path).  The pass is a latency chain (key load -> bitmap probe -> word) hidden by resident warps: five CTAs per
// SM.  Only rows whose key lies in the pass's slice touch the bitmap; the key stream is marked L2 evict-first and the
// bitmap evict-last ($BOSQL_PROBE_HINTS=0 switches the hints off): ncu shows the slice competing with the stream for L2
// (half of a 64 MB slice's probes went to DRAM without them).
// FIRST: reads the key column itself and leaves key - key_min as a uint32 per row for the passes that follow.
template <bool FIRST, int ROWS, bool HINTS>
__global__ void __launch_bounds__(kBlock, 5) k_probe_bits(const __grid_constant__ ProbeBitsParams p) {
    const int lane = threadIdx.x & 31;
    const size_t warps = static_cast<size_t>(gridDim.x) * (kBlock / 32);
    const size_t warp = static_cast<size_t>(blockIdx.x) * (kBlock / 32) + (threadIdx.x >> 5);
    constexpr size_t per = 32 * ROWS;
    const size_t n_chunks = (p.row_end - p.row_begin + per - 1) / per;
    const size_t n_words = (p.row_end + 31) / 32;
    uint64_t stream_policy = 0, keep_policy = 0;
    if (HINTS) {
        stream_policy = l2_policy_evict_first();
        keep_policy = l2_policy_evict_last();
    }
    for (size_t c = warp; c < n_chunks; c += warps) {
        const size_t base = p.row_begin + c * per;
        unsigned idx[ROWS];
        if (FIRST) {
            long long k[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const size_t i = base + 32 * r + lane;
                k[r] = 0;
                if (i < p.row_end) {
                    if (p.key_kind == BQ_INT64)
                        k[r] = HINTS ? ldg_stream_i64_hint(static_cast<const long long*>(p.key) + i, stream_policy) : __ldg(static_cast<const long long*>(p.key) + i);
                    else if (p.key_kind == BQ_STRING) k[r] = static_cast<unsigned>(__ldg(static_cast<const int*>(p.key) + i));
                    else k[r] = __ldg(static_cast<const int*>(p.key) + i);
                }
            }
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const bool in = base + 32 * r + lane < p.row_end;
                const unsigned long long d = static_cast<unsigned long long>(k[r] - p.key_min);
                idx[r] = (in && d < p.domain) ? static_cast<unsigned>(d) : 0xFFFFFFFFu;
                if (in && p.idx32) p.idx32[base + 32 * r + lane] = idx[r];
            }
        } else {
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const size_t i = base + 32 * r + lane;
                idx[r] = 0xFFFFFFFFu;
                if (i < p.row_end)
                    idx[r] = static_cast<unsigned>(HINTS ? ldg_stream_i32_hint(reinterpret_cast<const int*>(p.idx32) + i, stream_policy)
                                                         : __ldg(reinterpret_cast<const int*>(p.idx32) + i));
            }
        }
        unsigned mine_word = 0;
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const unsigned long long s = static_cast<unsigned long long>(idx[r]) - p.slice_lo;
            const bool mine = idx[r] != 0xFFFFFFFFu && s < p.slice_len;
            bool hit = false;
            if (mine) {
                const unsigned w = HINTS ? ldg_keep_u32(p.bitmap + (idx[r] >> 5), keep_policy) : __ldg(p.bitmap + (idx[r] >> 5));
                hit = (w >> (idx[r] & 31)) & 1u;
            }
            const unsigned word = __ballot_sync(0xffffffffu, hit);
            if (lane == r) mine_word = word;
        }
        // lanes 0..ROWS-1 hold the words of the chunk: one coalesced store, or a red.or into the earlier passes' words
        const size_t w = (base >> 5) + lane;
        if (lane < ROWS && w < n_words) {
            if (p.first) p.out[w] = mine_word;
            else if (mine_word) atomicOr(p.out + w, mine_word);
        }
    }
}

static size_t next_pow2(size_t v) {
    size_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

static void free_tables(bq_join* j) {
    dev_free(j->ctx, j->bitmap);
    dev_free(j->ctx, j->direct);
    dev_free(j->ctx, j->h_slots);
    dev_free(j->ctx, j->h_occ);
    j->h_occ = nullptr;
    j->bitmap = j->direct = nullptr;
    j->h_slots = nullptr;
}

// A bitmap built for an exchange carries its build counters BEHIND the bitmap words, in a form that survives a word-wise
// sum over ranks: the rows inserted as three 21-bit limbs and the number of ranks that saw a key outside the domain - so
// one all-reduce merges the bitmaps AND tells every rank what all ranks inserted, with no host exchange in between.
constexpr size_t kTrailerWords = BQ_JOIN_TRAILER_WORDS;
__global__ void k_pack_trailer(const unsigned long long* __restrict__ counters /* [0] inserted, [1] flags */, unsigned* __restrict__ trailer) {
    const unsigned long long ins = counters[0];
    trailer[0] = static_cast<unsigned>(ins & 0x1FFFFFu);
    trailer[1] = static_cast<unsigned>((ins >> 21) & 0x1FFFFFu);
    trailer[2] = static_cast<unsigned>(ins >> 42);
    trailer[3] = (counters[1] & 2ull) ? 1u : 0u;      // keys outside [key_min, key_max]: stale statistics
    for (size_t i = 4; i < kTrailerWords; ++i) trailer[i] = 0;
}

// Builds one table kind; returns the kernel's flag word.  nosync (BITMAP only): nothing is read back - the counters are
// packed behind the bitmap and 0 is returned; bq_join_bitmap_verdict reads them after the exchange.
static int build_kind(bq_ctx* ctx, const bq_join_spec* spec, int kind, bq_join* j, bool nosync = false) {
    BuildParams p{};
    p.key = spec->key->ptr;
    p.key_kind = spec->key->type;
    bool never = false;
    int n_pred = 0;
    for (int s = 0; s < 3; ++s) {
        if (spec->pred[s].from_build) throw std::runtime_error("bq_join_build: slots are build-side columns already");
        DSlot d = make_dslot(spec->pred[s], spec->row_end, "pred");
        never = normalise_slot(d) || never;
        if (d.ptr && d.nr) p.s[n_pred++] = d;          // compacted: slots [0, n_pred) carry ranges
    }
    if (spec->mask) {
        if (spec->mask->type != BQ_INT64 || spec->mask->n < spec->row_end) throw std::runtime_error("mask must be an INT64 column covering the row range");
        p.mask = static_cast<const long long*>(spec->mask->ptr);
    }
    p.row_begin = spec->row_begin;
    p.row_end = spec->row_end;
    const size_t n = spec->row_end - spec->row_begin;
    free_tables(j);
    j->kind = kind;
    j->bytes = 0;
    if (kind == BQ_JOIN_BITMAP || kind == BQ_JOIN_DIRECT) {
        j->key_min = spec->key_min;
        j->key_max = spec->key_max;
        p.key_min = spec->key_min;
        p.domain = static_cast<unsigned long long>(spec->key_max - spec->key_min) + 1ULL;
        if (kind == BQ_JOIN_BITMAP) {
            j->bitmap_words = (p.domain + 31) / 32;
            j->bytes = j->bitmap_words * 4;
            j->bitmap = static_cast<unsigned*>(dev_alloc(ctx, j->bytes + 4 + kTrailerWords * 4));
            BQ_CUDA(cudaMemsetAsync(j->bitmap, 0, j->bytes + 4 + kTrailerWords * 4, ctx->stream));
            p.bitmap = j->bitmap;
        } else {
            j->bytes = p.domain * 4;
            j->direct = static_cast<unsigned*>(dev_alloc(ctx, j->bytes + 4));
            BQ_CUDA(cudaMemsetAsync(j->direct, 0, j->bytes + 4, ctx->stream));
            p.direct = j->direct;
        }
    } else {
        // The table is sized from the rows that will be INSERTED, not from the rows scanned: with a selective build-side
        // predicate (Q2: one order in four) a table sized for every scanned row is four times larger than needed - 8.6 GB
        // for 250 M orders, which random probes cannot even keep in the TLBs.  One counting pass over the predicate columns
        // (and one host round trip) buys a table of 2x the inserted rows.
        size_t expect = n;
        if (n && !never && (n_pred > 0 || p.mask)) {
            auto* dcount = static_cast<unsigned long long*>(scratch(ctx, 32));
            BQ_CUDA(cudaMemsetAsync(dcount, 0, 16, ctx->stream));
            BuildParams pc = p;
            pc.n_inserted = dcount;
            pc.flags = reinterpret_cast<int*>(dcount + 1);
            const int grid = grid_for(ctx, n, 8);
            if (p.key_kind == BQ_INT64 && !p.mask && n_pred <= 1) k_join_build<BQ_JOIN_AUTO, BQ_INT64, 1><<<grid, kBlock, 0, ctx->stream>>>(pc);
            else k_join_build<BQ_JOIN_AUTO, -1, -1><<<grid, kBlock, 0, ctx->stream>>>(pc);
            ctx->launches++;
            BQ_CUDA(cudaGetLastError());
            auto* hc = static_cast<unsigned long long*>(pinned(ctx, 32));
            BQ_CUDA(cudaMemcpyAsync(hc, dcount, 8, cudaMemcpyDeviceToHost, ctx->stream));
            BQ_CUDA(cudaStreamSynchronize(ctx->stream));
            expect = static_cast<size_t>(hc[0]);
        }
        // load factor in (1/8, 1/4]: a probe of an absent key ends at its home slot three times out of four, and the
        // occupancy bits (cap / 8 bytes, L2-resident up to 2^29 slots) answer those without touching the table
        size_t cap = next_pow2(expect * 4 < 1024 ? 1024 : expect * 4);
        if (cap * sizeof(JoinSlot) > (32ull << 30)) cap = next_pow2(expect * 2);      // very large builds: half the memory, longer probe chains
        j->h_mask = cap - 1;
        j->bytes = cap * sizeof(JoinSlot);
        j->h_slots = static_cast<JoinSlot*>(dev_alloc(ctx, cap * sizeof(JoinSlot)));
        BQ_CUDA(cudaMemsetAsync(j->h_slots, 0, cap * sizeof(JoinSlot), ctx->stream));
        p.h_slots = j->h_slots;
        if (cap <= (1ull << 29)) {
            j->h_occ = static_cast<unsigned*>(dev_alloc(ctx, cap / 8 + 4));
            BQ_CUDA(cudaMemsetAsync(j->h_occ, 0, cap / 8 + 4, ctx->stream));
            j->bytes += cap / 8;
        }
        p.h_occ = j->h_occ;
        p.h_mask = j->h_mask;
    }
    // [0] rows inserted, [1] flags, [2] bits set in the finished bitmap - fetched in ONE host round trip
    auto* d = static_cast<unsigned long long*>(scratch(ctx, 32));
    BQ_CUDA(cudaMemsetAsync(d, 0, 32, ctx->stream));
    p.n_inserted = d;
    p.flags = reinterpret_cast<int*>(d + 1);
    if (n && !never) {
        int grid = grid_for(ctx, n, 8);
        const bool lean = p.key_kind == BQ_INT64 && !p.mask && n_pred <= 1;
        const int pw = n_pred ? width_of(p.s[0].kind) : 0;
#define BQ_BUILD(KIND)                                                                                         \
        if (lean && n_pred == 0) k_join_build<KIND, BQ_INT64, 0><<<grid, kBlock, 0, ctx->stream>>>(p);         \
        else if (lean) k_join_build<KIND, BQ_INT64, 1><<<grid, kBlock, 0, ctx->stream>>>(p);                   \
        else k_join_build<KIND, -1, -1><<<grid, kBlock, 0, ctx->stream>>>(p);
        if (kind == BQ_JOIN_BITMAP && lean && p.row_begin % 8 == 0) {
            if (pw == 0) k_bitmap_build4<0><<<grid, kBlock, 0, ctx->stream>>>(p);
            else if (pw == 4) k_bitmap_build4<4><<<grid, kBlock, 0, ctx->stream>>>(p);
            else k_bitmap_build4<8><<<grid, kBlock, 0, ctx->stream>>>(p);
        }
        else if (kind == BQ_JOIN_BITMAP) { BQ_BUILD(BQ_JOIN_BITMAP) }
        else if (kind == BQ_JOIN_DIRECT) { BQ_BUILD(BQ_JOIN_DIRECT) }
        else { BQ_BUILD(BQ_JOIN_HASH) }
#undef BQ_BUILD
        ctx->launches++;
        BQ_CUDA(cudaGetLastError());
        if (kind == BQ_JOIN_BITMAP && !nosync) {
            // the inserts were reductions: a key inserted twice shows as a bitmap with fewer bits than rows
            k_popcount_words<<<grid_for(ctx, j->bitmap_words, 8), kBlock, 0, ctx->stream>>>(j->bitmap, j->bitmap_words, d + 2);
            ctx->launches++;
            BQ_CUDA(cudaGetLastError());
        }
    }
    if (nosync) {
        k_pack_trailer<<<1, 1, 0, ctx->stream>>>(d, j->bitmap + j->bitmap_words);
        ctx->launches++;
        BQ_CUDA(cudaGetLastError());
        j->build_rows = 0;
        j->bitmap_bits = 0;
        return 0;
    }
    auto* h = static_cast<unsigned long long*>(pinned(ctx, 32));
    BQ_CUDA(cudaMemcpyAsync(h, d, 24, cudaMemcpyDeviceToHost, ctx->stream));
    BQ_CUDA(cudaStreamSynchronize(ctx->stream));
    j->build_rows = static_cast<size_t>(h[0]);
    int flags = static_cast<int>(h[1] & 0xFFFFFFFFull);
    if (kind == BQ_JOIN_BITMAP && !(flags & 2) && h[2] != j->build_rows) flags |= 1;
    j->bitmap_bits = h[2];
    return flags;
}

}  // namespace bq

using namespace bq;

extern "C" {

int bq_join_build(bq_ctx* ctx, const bq_join_spec* spec, bq_join** out) {
    return guarded([&] {
        if (!spec->key) throw std::runtime_error("join build needs a key column");
        if (spec->row_end < spec->row_begin || spec->row_end > spec->key->n) throw std::runtime_error("bad build row range");
        if (spec->row_end > 0xFFFFFFFEull) throw std::runtime_error("row ids are 32-bit: at most 2^32-1 build rows");
        auto* j = new bq_join();
        j->ctx = ctx;
        try {
            int kind = spec->kind;
            const size_t n = spec->row_end - spec->row_begin;
            if (kind == BQ_JOIN_AUTO) {
                kind = BQ_JOIN_HASH;
                bool int_key = spec->key->type != BQ_DOUBLE;
                if (int_key && spec->key_max >= spec->key_min) {
                    unsigned long long dom = static_cast<unsigned long long>(spec->key_max - spec->key_min) + 1ULL;
                    // dense enough that a direct-address table beats 12 B/row of hash slots
                    if (dom <= (1ULL << 32) && dom <= 8ULL * (n ? n : 1) + 1024) kind = spec->need_rows ? BQ_JOIN_DIRECT : BQ_JOIN_BITMAP;
                }
            }
            int flags = build_kind(ctx, spec, kind, j);
            if (kind != BQ_JOIN_HASH && (flags & 3)) {
                // duplicate keys (or stale catalog bounds): a bitmap / direct table cannot represent them
                flags = build_kind(ctx, spec, BQ_JOIN_HASH, j);
            }
            if (flags & 4) throw std::runtime_error("join table overflow");
            *out = j;
        } catch (...) {
            free_tables(j);
            delete j;
            throw;
        }
    });
}

void bq_join_free(bq_ctx* ctx, bq_join* j) {
    (void)ctx;
    if (!j) return;
    free_tables(j);
    delete j;
}

int bq_join_kind(const bq_join* j) { return j->kind; }
size_t bq_join_bytes(const bq_join* j) { return j->bytes; }
size_t bq_join_build_rows(const bq_join* j) { return j->build_rows; }

int bq_join_bitmap_popcount(bq_ctx* ctx, const bq_join* j, uint64_t* out) {
    return guarded([&] {
        if (j->kind != BQ_JOIN_BITMAP) throw std::runtime_error("not a bitmap join");
        auto* d = static_cast<unsigned long long*>(scratch(ctx, 16));
        BQ_CUDA(cudaMemsetAsync(d, 0, 8, ctx->stream));
        if (j->bitmap_words) {
            k_popcount_words<<<grid_for(ctx, j->bitmap_words, 8), kBlock, 0, ctx->stream>>>(j->bitmap, j->bitmap_words, d);
            ctx->launches++;
            BQ_CUDA(cudaGetLastError());
        }
        auto* h = static_cast<unsigned long long*>(pinned(ctx, 8));
        BQ_CUDA(cudaMemcpyAsync(h, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
        BQ_CUDA(cudaStreamSynchronize(ctx->stream));
        *out = *h;
    });
}

int bq_join_build_bitmap_nosync(bq_ctx* ctx, const bq_join_spec* spec, bq_join** out) {
    return guarded([&] {
        if (!spec->key) throw std::runtime_error("join build needs a key column");
        if (spec->key->type == BQ_DOUBLE) throw std::runtime_error("bitmap joins have integer keys");
        if (spec->row_end < spec->row_begin || spec->row_end > spec->key->n) throw std::runtime_error("bad build row range");
        if (spec->key_max < spec->key_min || static_cast<unsigned long long>(spec->key_max - spec->key_min) >= (1ULL << 32))
            throw std::runtime_error("a bitmap join needs a key domain of at most 2^32 keys");
        auto* j = new bq_join();
        j->ctx = ctx;
        try {
            build_kind(ctx, spec, BQ_JOIN_BITMAP, j, true);
            *out = j;
        } catch (...) {
            free_tables(j);
            delete j;
            throw;
        }
    });
}

int bq_join_bitmap_verdict(bq_ctx* ctx, bq_join* j, uint64_t* set_bits, uint64_t* inserted, int* flags) {
    return guarded([&] {
        if (j->kind != BQ_JOIN_BITMAP) throw std::runtime_error("not a bitmap join");
        auto* d = static_cast<unsigned long long*>(scratch(ctx, 16));
        BQ_CUDA(cudaMemsetAsync(d, 0, 8, ctx->stream));
        if (j->bitmap_words) {
            k_popcount_words<<<grid_for(ctx, j->bitmap_words, 8), kBlock, 0, ctx->stream>>>(j->bitmap, j->bitmap_words, d);
            ctx->launches++;
            BQ_CUDA(cudaGetLastError());
        }
        auto* h = static_cast<unsigned long long*>(pinned(ctx, 8 + kTrailerWords * 4));
        BQ_CUDA(cudaMemcpyAsync(h, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
        BQ_CUDA(cudaMemcpyAsync(h + 1, j->bitmap + j->bitmap_words, kTrailerWords * 4, cudaMemcpyDeviceToHost, ctx->stream));
        BQ_CUDA(cudaStreamSynchronize(ctx->stream));            // the one host round trip of a build that was exchanged
        const auto* t = reinterpret_cast<const unsigned*>(h + 1);
        const uint64_t ins = static_cast<uint64_t>(t[0]) + (static_cast<uint64_t>(t[1]) << 21) + (static_cast<uint64_t>(t[2]) << 42);
        j->build_rows = static_cast<size_t>(ins);
        j->bitmap_bits = h[0];
        if (set_bits) *set_bits = h[0];
        if (inserted) *inserted = ins;
        if (flags) *flags = t[3] ? 2 : 0;
    });
}

void* bq_join_bitmap_ptr(const bq_join* j, size_t* n_words) {
    if (n_words) *n_words = j->bitmap_words;
    return j->bitmap;
}

int bq_join_probe_bits(bq_ctx* ctx, const bq_join* j, const bq_col* probe_key, size_t row_begin, size_t row_end,
                       size_t slice_bytes, bq_col** out_bits) {
    return guarded([&] {
        if (j->kind != BQ_JOIN_BITMAP) throw std::runtime_error("key-range probe passes need a bitmap join");
        if (probe_key->type == BQ_DOUBLE) throw std::runtime_error("bitmap joins have integer keys");
        if (row_end < row_begin || row_end > probe_key->n) throw std::runtime_error("bad probe row range");
        if (row_begin % 128) throw std::runtime_error("probe passes need a row range starting at a multiple of 128");
        if (slice_bytes < 1024) slice_bytes = 1024;
        const unsigned long long domain = static_cast<unsigned long long>(j->key_max - j->key_min) + 1ULL;
        // whole 32-bit words per slice, equal slices
        unsigned long long passes = (j->bytes + slice_bytes - 1) / slice_bytes;
        if (passes < 1) passes = 1;
        const unsigned long long slice_keys = ((domain + passes - 1) / passes + 31) / 32 * 32;
        passes = (domain + slice_keys - 1) / slice_keys;
        const size_t words = (row_end + 31) / 32;
        bq_col* bits = new_col(ctx, BQ_STRING, (words + 31) / 32 * 32);       // whole groups of words (scan: four at once; passes: up to 16)
        unsigned* idx32 = nullptr;
        const char* hint_env = std::getenv("BOSQL_PROBE_HINTS");
        const bool hints = !(hint_env && *hint_env == '0');
        try {
            BQ_CUDA(cudaMemsetAsync(bits->ptr, 0, bits->n * 4, ctx->stream));
            if (row_end > row_begin) {
                // with several passes the first one leaves key - key_min as 4 bytes per row: the later passes read half the bytes
                if (passes > 1 && domain <= 0xFFFFFFFFull) idx32 = static_cast<unsigned*>(dev_alloc(ctx, row_end * 4));
                ProbeBitsParams p{};
                p.key = probe_key->ptr;
                p.key_kind = probe_key->type;
                p.row_begin = row_begin;
                p.row_end = row_end;
                p.key_min = j->key_min;
                p.domain = domain;
                p.bitmap = j->bitmap;
                p.out = static_cast<unsigned*>(bits->ptr);
                p.idx32 = idx32;
                const int grid = grid_for(ctx, row_end - row_begin, 5);
                for (unsigned long long pass = 0; pass < passes; ++pass) {
                    p.slice_lo = pass * slice_keys;
                    p.slice_len = std::min(slice_keys, domain - p.slice_lo);
                    p.first = pass == 0;
                    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
                    if (ctx->profile) {
                        BQ_CUDA(cudaEventCreate(&ev0));
                        BQ_CUDA(cudaEventCreate(&ev1));
                        BQ_CUDA(cudaEventRecord(ev0, ctx->stream));
                    }
                    const bool first_kind = pass == 0 || !idx32;
                    if (first_kind && hints) k_probe_bits<true, 8, true><<<grid, kBlock, 0, ctx->stream>>>(p);
                    else if (first_kind) k_probe_bits<true, 8, false><<<grid, kBlock, 0, ctx->stream>>>(p);
                    else if (hints) k_probe_bits<false, 16, true><<<grid, kBlock, 0, ctx->stream>>>(p);
                    else k_probe_bits<false, 16, false><<<grid, kBlock, 0, ctx->stream>>>(p);
                    if (ctx->profile) {
                        BQ_CUDA(cudaEventRecord(ev1, ctx->stream));
                        ctx->profile_events.emplace_back(ev0, ev1);
                    }
                    ctx->launches++;
                    BQ_CUDA(cudaGetLastError());
                }
            }
            dev_free(ctx, idx32);
        } catch (...) {
            dev_free(ctx, idx32);
            free_col(bits);
            throw;
        }
        *out_bits = bits;
    });
}

int bq_join_probe(bq_ctx* ctx, const bq_join* j, const bq_col* probe_key, const bq_col* probe_rowids,
                  size_t row_begin, size_t row_end, bq_col** out_probe_rows, bq_col** out_build_rows) {
    return guarded([&] {
        if (j->kind == BQ_JOIN_BITMAP) throw std::runtime_error("a bitmap join cannot materialise build rows");
        if (probe_rowids && probe_rowids->type != BQ_STRING) throw std::runtime_error("row ids must be a uint32 column");
        if (!probe_rowids && (row_end < row_begin || row_end > probe_key->n)) throw std::runtime_error("bad probe row range");
        ProbeParams p{};
        p.key = probe_key->ptr;
        p.key_kind = probe_key->type;
        p.rowids = probe_rowids ? static_cast<const unsigned*>(probe_rowids->ptr) : nullptr;
        p.row_begin = row_begin;
        p.n = probe_rowids ? probe_rowids->n : row_end - row_begin;
        p.jkind = j->kind;
        p.key_min = j->key_min;
        p.domain = static_cast<unsigned long long>(j->key_max - j->key_min) + 1ULL;
        p.bitmap = j->bitmap;
        p.direct = j->direct;
        p.h_slots = j->h_slots;
        p.h_mask = j->h_mask;
        bq_col *op = nullptr, *ob = nullptr;
        unsigned* counts = nullptr;
        unsigned long long* offsets = nullptr;
        try {
            size_t total = 0;
            if (p.n) {
                counts = static_cast<unsigned*>(dev_alloc(ctx, p.n * 4));
                offsets = static_cast<unsigned long long*>(dev_alloc(ctx, p.n * 8));
                unsigned blocks = static_cast<unsigned>((p.n + kBlock - 1) / kBlock);
                k_probe_count<<<blocks, kBlock, 0, ctx->stream>>>(p, counts);
                ctx->launches++;
                BQ_CUDA(cudaGetLastError());
                total = exclusive_scan_u32(ctx, counts, p.n, offsets);
                if (total > 0xFFFFFFFFull) throw std::runtime_error("join result exceeds 2^32 rows");
            }
            op = new_col(ctx, BQ_STRING, total);
            ob = new_col(ctx, BQ_STRING, total);
            if (total) {
                unsigned blocks = static_cast<unsigned>((p.n + kBlock - 1) / kBlock);
                k_probe_fill<<<blocks, kBlock, 0, ctx->stream>>>(p, counts, offsets, static_cast<unsigned*>(op->ptr),
                                                               static_cast<unsigned*>(ob->ptr));
                ctx->launches++;
                BQ_CUDA(cudaGetLastError());
            }
            dev_free(ctx, counts);
            dev_free(ctx, offsets);
            *out_probe_rows = op;
            *out_build_rows = ob;
        } catch (...) {
            dev_free(ctx, counts);
            dev_free(ctx, offsets);
            free_col(op);
            free_col(ob);
            throw;
        }
    });
}

}  // extern "C"
