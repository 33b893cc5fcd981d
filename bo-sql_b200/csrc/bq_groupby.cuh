// bq_groupby.cuh — k_group_tables: high-cardinality GROUP BY over hash-partitioned rows, one shared-memory table per CTA.
//
// Replaces the accumulate AND emit phases of HashAggregate::next (src/exec/operator.cpp:984-1014, 1016-1062) for the case
// the L2-resident table handles worst: tens of millions of groups, a handful of rows each (configuration 4).  There every
// row costs three random global transactions (key probe, count reduction, sum reduction), and the load/store units retire
// about one such lane-operation per 1.3 clocks per SM - 3.7 ms for 250 M rows however the table is laid out.
//
// Here the rows arrive ordered by bq_partition (P <= 1024 partitions).  CTA (q, r) owns the keys of partition q whose
// second hash falls into split r of S, and keeps them in an open-addressing table in its own shared memory (8192 slots;
// load factor <= 0.55 by the sizing rule in bq_group_tables_plan).  The S CTAs of a partition stream the same rows at the
// same time, so all but the first read them from L2.  Shared-memory atomics would be no faster than global reductions
// (two clocks per lane, and an f64 add is a compare-and-swap loop), so the table is updated WITHOUT atomics:
//
//   produce  each warp scans up to 96 keys (coalesced, loaded one iteration ahead), keeps the rows of its split and
//            appends (key, values) to its own queue by ballot + popc - no shared counter;
//   claim    every lane holding a row walks the probe sequence to its key or to the first empty slot and writes its
//            thread id into that slot's tag;                                                  -- barrier --
//   update   a lane that reads its own id back owns the slot until the next barrier: plain load-add-store of count and
//            sums (the key too, if the slot was empty).  Lanes that lost the tag go round once more; what is still
//            unplaced after two rounds (well under 1 %) is finished with atomics in a phase of its own.
//
// A slot has exactly one owner between two barriers (the last lane to write its tag before the barrier), so the
// non-atomic updates cannot collide, and every contended slot is won by somebody in every round.
// At the end each CTA numbers its groups (ballot / popc, one atomicAdd on the global cursor per CTA) and writes the
// finished output columns: no table in HBM, no presence pass, no separate emit.
//
// Measured on a B200 (profiles/README.md, "Configuration 4 taken apart"): 250 M rows / 12.5 M groups in 1024 partitions x 3
// splits take 6.8 ms here against 3.7 ms for the L2-resident table of bq_scan.cu, and the 1024-way partition this kernel needs
// takes 8.4 ms against 2.1 ms for the 32-way one that table needs.  The kernel is bound by instruction issue (17 warp
// instructions per row, three quarters of them in the key scan that every split repeats), the partition by half-written
// sectors.  It is therefore OPT-IN (BOSQL_GROUP_TABLES=1 in the operator layer); the default path is the L2-resident table.
//
// The file is plain CUDA C++ without inline PTX on purpose: tests/cpp/emu compiles this very source for the host with a
// fibre-based CUDA emulation and checks it against std::map on the CPU (-m "not gpu").
#pragma once

#include <cstddef>
#include <cstdint>

#include "bosql_b200.h"

#ifndef BQ_GROUPBY_FN
#define BQ_GROUPBY_FN __device__ __forceinline__
#endif
#ifndef BQ_GROUP_STAT
#define BQ_GROUP_STAT(counter)        // the emulation counts where rows are placed (round 0, round 1, atomic phase)
#endif

namespace bq {

constexpr int kGroupThreads = 1024;
constexpr int kGroupWarps = kGroupThreads / 32;
constexpr int kGroupDepth = 3;                     // steps of 32 rows a warp loads ahead (and scans per iteration) at most
constexpr int kGroupQueue = 96;                    // entries of a warp's queue: a step is scanned while 32 entries are free
constexpr int kGroupRounds = 2;                    // claim / update rounds before the atomic phase
constexpr unsigned kGroupMaxWalk = 512;            // a longer probe sequence means the table is (all but) full: give up
constexpr long long kGroupEmptyKey = INT64_MIN;    // empty marker; a real INT64_MIN key lives in the spare slot [slots]

struct GroupParams {
    const void* key;
    int key_kind;                       // BQ_INT64 / BQ_STRING / BQ_DATE32 (integer keys)
    const void* val[2];                 // aggregate arguments (plain columns)
    int val_kind[2];
    int nv;
    const long long* offsets;           // [P + 1] row offsets of the partitions (bq_partition)
    unsigned splits;                    // S: CTAs per partition
    unsigned slots;                     // table slots per CTA, a power of two >= 1024
    void* out_key;
    int n_out;
    int func[BQ_MAX_AGG_OUT];
    int v[BQ_MAX_AGG_OUT];
    int as_int[BQ_MAX_AGG_OUT];
    void* out[BQ_MAX_AGG_OUT];
    unsigned long long capacity;        // rows the output columns can hold
    unsigned long long* cursor;         // groups written so far (all CTAs)
    int* err;                           // 2 = a table filled up, 8 = output capacity exceeded (sizing bug)
};

// second hash (murmur3's 32-bit finaliser over the folded key): split = high bits, home slot = low bits.  Independent of
// key_hash, whose top bits chose the partition.
BQ_GROUPBY_FN uint32_t group_mix(uint64_t k) {
    uint32_t h = static_cast<uint32_t>(k) ^ (static_cast<uint32_t>(k >> 32) * 0x9E3779B1u);
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}

// dynamic shared memory of one CTA, in bytes
inline size_t group_smem_bytes(unsigned slots, int nv) {
    const size_t sums = nv > 1 ? 2 : 1;
    size_t b = static_cast<size_t>(slots + 1) * 8 * (1 + sums);                       // keys, sums
    b += static_cast<size_t>(kGroupWarps) * kGroupQueue * 8 * (1 + sums);             // queues
    b += static_cast<size_t>(slots + 2) * 4;                                          // counts
    b += static_cast<size_t>(slots + 2) * 2;                                          // tags
    return b;
}

#if defined(__CUDACC__) || defined(BQ_CUDA_EMU)

BQ_GROUPBY_FN long long group_load_int(const void* base, int kind, size_t i) {
    if (kind == BQ_INT64 || kind == BQ_DOUBLE) return __ldg(static_cast<const long long*>(base) + i);
    if (kind == BQ_STRING) return static_cast<long long>(__ldg(static_cast<const unsigned*>(base) + i));
    return static_cast<long long>(__ldg(static_cast<const int*>(base) + i));
}
// datum_as_double (src/exec/operator.cpp:280-292)
BQ_GROUPBY_FN double group_load_value(const void* base, int kind, size_t i) {
    if (kind == BQ_DOUBLE) return __ldg(static_cast<const double*>(base) + i);
    return static_cast<double>(group_load_int(base, kind, i));
}

// one finished group: [key] then the outputs, as HashAggregate's emit (src/exec/operator.cpp:1030-1050)
BQ_GROUPBY_FN void group_write(const GroupParams& p, unsigned long long at, long long key, unsigned cnt, double s0, double s1) {
    if (at >= p.capacity) {
        atomicOr(p.err, 8);
        return;
    }
    if (p.out_key) {
        if (p.key_kind == BQ_INT64) static_cast<long long*>(p.out_key)[at] = key;
        else if (p.key_kind == BQ_STRING) static_cast<unsigned*>(p.out_key)[at] = static_cast<unsigned>(key);
        else static_cast<int*>(p.out_key)[at] = static_cast<int>(key);
    }
    for (int o = 0; o < p.n_out; ++o) {
        const double s = p.v[o] == 0 ? s0 : s1;
        if (p.func[o] == BQ_AGG_COUNT) {
            static_cast<long long*>(p.out[o])[at] = static_cast<long long>(cnt);
        } else if (p.func[o] == BQ_AGG_SUM) {
            if (p.as_int[o]) static_cast<long long*>(p.out[o])[at] = static_cast<long long>(s);     // :1044
            else static_cast<double*>(p.out[o])[at] = s;
        } else {
            static_cast<double*>(p.out[o])[at] = cnt == 0 ? 0.0 : __ddiv_rn(s, static_cast<double>(cnt));   // :1047
        }
    }
}

__global__ void __launch_bounds__(kGroupThreads, 1) k_group_tables(const __grid_constant__ GroupParams p) {
#if defined(BQ_CUDA_EMU)
    unsigned char* group_smem = emu::dyn_smem();
#else
    extern __shared__ __align__(16) unsigned char group_smem[];
#endif
    __shared__ unsigned wtot[kGroupWarps];
    __shared__ unsigned long long s_base;
    __shared__ int s_stop;
    __shared__ int s_full;

    const unsigned slots = p.slots, mask = slots - 1;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const bool two = p.nv > 1;
    long long* skey = reinterpret_cast<long long*>(group_smem);                       // [slots + 1]
    double* ssum0 = reinterpret_cast<double*>(skey + slots + 1);                      // [slots + 1]
    double* ssum1 = ssum0 + (slots + 1);                                              // [slots + 1] when two
    long long* qkey = reinterpret_cast<long long*>(ssum1 + (two ? slots + 1 : 0));    // [warps][kGroupQueue]
    double* qv0 = reinterpret_cast<double*>(qkey + kGroupWarps * kGroupQueue);
    double* qv1 = qv0 + kGroupWarps * kGroupQueue;                                    // when two
    unsigned* scnt = reinterpret_cast<unsigned*>(qv1 + (two ? kGroupWarps * kGroupQueue : 0));   // [slots + 2]
    unsigned short* stag = reinterpret_cast<unsigned short*>(scnt + slots + 2);       // [slots + 2]

    const unsigned part = blockIdx.x / p.splits, split = blockIdx.x % p.splits;
    const size_t lo = static_cast<size_t>(p.offsets[part]), hi = static_cast<size_t>(p.offsets[part + 1]);

    for (unsigned i = tid; i <= slots; i += kGroupThreads) {
        skey[i] = kGroupEmptyKey;
        ssum0[i] = 0.0;
        if (two) ssum1[i] = 0.0;
        scnt[i] = 0;
    }
    if (tid == 0) {
        s_stop = (hi <= lo) || (*reinterpret_cast<volatile int*>(p.err) != 0);     // nothing to do / another CTA failed
        s_full = 0;
    }
    __syncthreads();
    if (s_stop) return;

    long long* wq_key = qkey + warp * kGroupQueue;
    double* wq_v0 = qv0 + warp * kGroupQueue;
    double* wq_v1 = qv1 + warp * kGroupQueue;
    // a warp streams its own contiguous share of the partition, 32 rows (one per lane) at a time
    const size_t share = ((hi - lo + kGroupWarps - 1) / kGroupWarps + 31) / 32 * 32;
    size_t cursor = lo + warp * share < hi ? lo + warp * share : hi;
    const size_t end = cursor + share < hi ? cursor + share : hi;
    const int depth = p.splits < static_cast<unsigned>(kGroupDepth) ? static_cast<int>(p.splits) : kGroupDepth;
    unsigned qn = 0;                         // entries in this warp's queue (same value in every lane)

    // The rows of the next `depth` steps are loaded one iteration ahead: every warp meets the others at the barriers of
    // each iteration, so a load issued where its value is needed would expose the memory latency once per iteration.
    // About one row in `splits` belongs to this CTA, hence `depth` = splits steps for a full wave of 32.
    long long pk[kGroupDepth];
    double pv0[kGroupDepth], pv1[kGroupDepth];
    auto prefetch = [&]() {
#pragma unroll
        for (int j = 0; j < kGroupDepth; ++j) {
            const size_t row = cursor + static_cast<size_t>(j) * 32 + lane;
            pk[j] = 0;
            pv0[j] = 0.0;
            pv1[j] = 0.0;
            if (j < depth && row < end) {
                pk[j] = group_load_int(p.key, p.key_kind, row);
                if (p.nv > 0) pv0[j] = group_load_value(p.val[0], p.val_kind[0], row);
                if (two) pv1[j] = group_load_value(p.val[1], p.val_kind[1], row);
            }
        }
    };
    prefetch();

    for (;;) {
        // ---- produce: the rows of this split among the prefetched steps go to the warp's queue, while they fit
        int consumed = 0;
#pragma unroll
        for (int j = 0; j < kGroupDepth; ++j) {
            const size_t first = cursor + static_cast<size_t>(j) * 32;
            if (j < depth && consumed == j && first < end && qn + 32u <= static_cast<unsigned>(kGroupQueue)) {      // warp-uniform
                const long long k = pk[j];
                const bool mine = first + lane < end &&
                                  (p.splits == 1 || __umulhi(group_mix(static_cast<uint64_t>(k)), p.splits) == split);
                const unsigned b = __ballot_sync(0xffffffffu, mine);
                if (mine) {
                    const unsigned pos = qn + __popc(b & ((1u << lane) - 1u));
                    wq_key[pos] = k;
                    wq_v0[pos] = pv0[j];
                    if (two) wq_v1[pos] = pv1[j];
                }
                qn += __popc(b);
                ++consumed;
            }
        }
        cursor += static_cast<size_t>(consumed) * 32;
        if (cursor > end) cursor = end;
        __syncwarp();
        prefetch();          // steps that did not fit are simply loaded again (L1 / L2)
        // ---- take up to 32 rows off the end of the queue
        const unsigned take = qn < 32u ? qn : 32u;
        bool pending = lane < take;
        long long k = 0;
        double v0 = 0.0, v1 = 0.0;
        unsigned s = 0;
        bool spare = false;
        if (pending) {
            const unsigned e = qn - take + lane;
            k = wq_key[e];
            v0 = wq_v0[e];
            if (two) v1 = wq_v1[e];
            spare = k == kGroupEmptyKey;
            s = spare ? slots : (group_mix(static_cast<uint64_t>(k)) & mask);
        }
        qn -= take;
        const bool more = qn > 0 || cursor < end;
        int again = 0;

        // ---- claim / update rounds
#pragma unroll
        for (int r = 0; r < kGroupRounds; ++r) {
            bool found = spare;
            if (pending && !spare) {
                unsigned walked = 0;
                for (;;) {
                    const long long cur = reinterpret_cast<volatile long long*>(skey)[s];
                    if (cur == k) { found = true; break; }
                    if (cur == kGroupEmptyKey) break;
                    s = (s + 1) & mask;
                    if (++walked > kGroupMaxWalk) { s_full = 1; pending = false; break; }
                }
            }
            if (pending) stag[s] = static_cast<unsigned short>(tid);
            if (r == 0) {
                again = __syncthreads_or(more ? 1 : 0);
                // s_full is written before this barrier (the walks above, the previous iteration's atomic phase) or after
                // the next one: every thread reads the same value here
                if (*reinterpret_cast<volatile int*>(&s_full)) {
                    if (tid == 0) atomicOr(p.err, 2);
                    return;
                }
            } else {
                __syncthreads();
            }
            if (pending && reinterpret_cast<volatile unsigned short*>(stag)[s] == tid) {
                // this lane owns slot s until the next barrier
                bool ok = found;
                if (!found) {
                    const long long cur = skey[s];               // seen empty during the walk; claimed by an owner of an earlier round?
                    if (cur == kGroupEmptyKey) {
                        skey[s] = k;
                        ok = true;
                    } else {
                        ok = cur == k;
                    }
                }
                if (ok) {
                    scnt[s] = scnt[s] + 1u;
                    ssum0[s] = __dadd_rn(ssum0[s], v0);
                    if (two) ssum1[s] = __dadd_rn(ssum1[s], v1);
                    pending = false;
                    BQ_GROUP_STAT(r);
                }
            }
            __syncthreads();        // the next round's tags (or the atomic phase) must not meet this round's owners
        }
        // ---- whatever is still unplaced: atomics, in a phase of its own
        if (pending) {
            bool done = false;
            if (spare) {
                done = true;
            } else {
                for (unsigned walked = 0; walked <= kGroupMaxWalk; ++walked) {
                    long long cur = reinterpret_cast<volatile long long*>(skey)[s];
                    if (cur == kGroupEmptyKey) {
                        const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(skey + s),
                                                                  static_cast<unsigned long long>(kGroupEmptyKey),
                                                                  static_cast<unsigned long long>(k));
                        cur = prev == static_cast<unsigned long long>(kGroupEmptyKey) ? k : static_cast<long long>(prev);
                    }
                    if (cur == k) { done = true; break; }
                    s = (s + 1) & mask;
                }
            }
            if (done) {
                BQ_GROUP_STAT(kGroupRounds);
                atomicAdd(scnt + s, 1u);
                atomicAdd(ssum0 + s, v0);
                if (two) atomicAdd(ssum1 + s, v1);
            } else {
                s_full = 1;
            }
        }
        if (!again) break;          // block-uniform: no warp had rows left beyond this wave
        // (the next iteration's claim barrier separates this atomic phase from its updates)
    }

    __syncthreads();
    if (s_full) {               // set by the last wave
        if (tid == 0) atomicOr(p.err, 2);
        return;
    }

    // ---- emit: number the groups warp by warp (slot order), one atomicAdd on the global cursor per CTA
    const unsigned per_warp = slots / kGroupWarps;
    unsigned total = 0;
    for (unsigned i = lane; i < per_warp; i += 32) total += __popc(__ballot_sync(0xffffffffu, scnt[warp * per_warp + i] > 0u));
    if (lane == 0) wtot[warp] = total;
    __syncthreads();
    unsigned before = 0, all = 0;
    for (unsigned w = 0; w < static_cast<unsigned>(kGroupWarps); ++w) {
        const unsigned t = wtot[w];
        all += t;
        if (w < warp) before += t;
    }
    const bool spare_used = scnt[slots] > 0u;
    if (tid == 0) s_base = atomicAdd(p.cursor, static_cast<unsigned long long>(all) + (spare_used ? 1ull : 0ull));
    __syncthreads();
    unsigned long long at = s_base + before;
    for (unsigned i = lane; i < per_warp; i += 32) {
        const unsigned slot = warp * per_warp + i;
        const unsigned c = scnt[slot];
        const unsigned b = __ballot_sync(0xffffffffu, c > 0u);
        if (c > 0u) group_write(p, at + __popc(b & ((1u << lane) - 1u)), skey[slot], c, ssum0[slot], two ? ssum1[slot] : 0.0);
        at += __popc(b);
    }
    if (tid == 0 && spare_used) group_write(p, s_base + all, kGroupEmptyKey, scnt[slots], ssum0[slots], two ? ssum1[slots] : 0.0);
}

#endif   // __CUDACC__ || BQ_CUDA_EMU

}  // namespace bq
