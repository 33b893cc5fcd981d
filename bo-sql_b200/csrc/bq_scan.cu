// bq_scan.cu — the fused hot path: scan -> selection -> [join probe] -> aggregate in ONE kernel.
//
// Replaces the per-row loops of ColumnarScan::next (src/exec/operator.cpp:345-384), Selection::next
// (:403-429, evaluate_predicate per row), HashJoin::next's probe (:764-837) and HashAggregate::next's
// accumulate phase (:984-1014) for pipelines whose predicate is a conjunction of `column OP literal`
// ranges (anything else arrives pre-evaluated as `mask`).  Nothing is materialised between operators.
//
// Memory behaviour (HBM-bound, the only roofline that applies to this path):
//  * every referenced column is read exactly once with coalesced vector loads on the read-only,
//    no-L1-allocate path: a warp owns a 128-row chunk; lane t holds rows {2t,2t+1,64+2t,65+2t}, so an
//    8-byte column is two fully coalesced 128-bit loads per lane and a 4-byte column two 64-bit loads;
//  * all loads of a chunk are issued before the first use (memory-level parallelism), the grid is a
//    multiple of the SM count and warps walk chunks grid-stride, so neighbouring warps stream
//    neighbouring DRAM pages;
//  * aggregation state: registers + shuffle (global aggregate), a per-CTA shared-memory table merged
//    once per CTA (dense low-cardinality GROUP BY: Q1), L2-resident global arrays with red.add
//    (dense medium cardinality: Q2's sku), or an open-addressing table in HBM (atomicCAS claim +
//    red.add; sized from catalog NDV).
//  * the slot layout is a template parameter for the shapes of the configurations (loads, widening and
//    range tests fold to straight-line code); any other shape runs the same source with run-time flags.
//
// Exactness: COUNT is an integer; SUM accumulates in double like AggState::sum
// (include/exec/operator.hpp:149-152) — exact for integer arguments while |sum| <= 2^53 (H1), within
// 1e-12 relative for DOUBLE arguments whose order of addition differs (H2).  `a*b` is rounded before it
// is added (H12): every FP op below is an explicit __d*_rn intrinsic, which is never contracted to FMA.
#include "bq_common.cuh"
#include "bq_internal.cuh"

#include <cstdlib>
#include <cstring>

namespace bq {

enum { S_KEY = 0, S_A = 1, S_B = 2, S_P0 = 3, S_P1 = 4, S_P2 = 5, S_JK = 6, N_SLOTS = 7 };
enum { G_NONE = 0, G_SMEM = 1, G_DENSE = 2, G_HASH = 3 };

constexpr long long kEmptyKey = INT64_MIN;   // empty marker of the group table (a real INT64_MIN key uses the spare slot)
constexpr uint32_t kGenericShape = 0xFFFFFFFFu;

struct DVExpr {
    int op, l_src, r_src, imm_is_f;
    long long imm_i;
    double imm_f;
};

struct ScanParams {
    DSlot s[N_SLOTS];
    unsigned present;          // bit per slot
    int nv;
    DVExpr v[2];
    size_t row_begin, row_end; // all rows
    size_t vec_begin;          // first row of the vector region (multiple of 4)
    size_t n_chunks;           // 128-row chunks in the vector region
    const long long* mask;     // optional 0/1 per row
    const unsigned* row_bits;  // BQ_JOIN_ROWBITS: bit i = row i joins (bq_join_probe_bits); row_begin is a multiple of 128
    int rowbits_skip;          // 1: loads of halves without a matching row are not issued (sector skipping)
    // group state
    long long key_min;
    unsigned long long key_domain;     // G_SMEM / G_DENSE: number of slots
    double* g_sum0;
    double* g_sum1;
    unsigned long long* g_cnt;
    long long* h_keys;                 // G_HASH: capacity+1 entries
    unsigned long long h_mask;
    int h_part_log2, h_part_shift;     // > 0: partition-major table for rows ordered by bq_partition
    // join
    int jmode;
    long long jk_min;
    unsigned long long jk_domain;
    const unsigned* j_bitmap;
    const unsigned* j_direct;
    const JoinSlot* jh_slots;
    const unsigned* jh_occ;            // optional: one bit per slot (occupied), L2-resident
    unsigned long long jh_mask;
    // G_NONE partials: [grid] x {cnt, sum0, sum1}
    unsigned long long* part_cnt;
    double* part_sum;                  // [grid][2]
    unsigned* ticket;
    int staged;                        // 1: qualifying rows are staged per warp and aggregated 32 at a time
    int need_count;                    // 0: no COUNT/AVG output -> skip the per-row count atomic of the global tables
    int* err;                          // 1 = integer division by zero, 2 = group table full, 4 = key outside dense domain
};

// ---- compile-time / run-time slot description ---------------------------------------------------
// SHAPE packs 4 bits per slot: 0 = absent, else 1+kind; bit 3 = value comes from the build side.
constexpr uint32_t shape_bits(int slot, int kind, bool from_build) {
    return static_cast<uint32_t>((1 + kind) | (from_build ? 8 : 0)) << (4 * slot);
}
template <uint32_t SHAPE>
struct Shape {
    static constexpr bool generic = (SHAPE == kGenericShape);
    BQ_D static bool present(const ScanParams& p, int s) {
        if (generic) return (p.present >> s) & 1u;
        return ((SHAPE >> (4 * s)) & 7u) != 0;
    }
    BQ_D static int kind(const ScanParams& p, int s) {
        if (generic) return p.s[s].kind;
        return static_cast<int>((SHAPE >> (4 * s)) & 7u) - 1;
    }
    BQ_D static bool from_build(const ScanParams& p, int s) {
        if (generic) return p.s[s].from_build != 0;
        return ((SHAPE >> (4 * s)) & 8u) != 0;
    }
    BQ_D static bool streamed(const ScanParams& p, int s) { return present(p, s) && !from_build(p, s); }
};

// ---- loads ---------------------------------------------------------------------------------------
BQ_D int2 ldg_stream2(const int2* p) {
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
BQ_D long long pack64(int lo, int hi) {
    return static_cast<long long>((static_cast<unsigned long long>(static_cast<unsigned>(hi)) << 32) |
                                  static_cast<unsigned>(lo));
}
// rows {base+2t, base+2t+1, base+64+2t, base+65+2t} of one column widened to 8-byte slot values
BQ_D void load_quad(const void* ptr, int kind, size_t base, int lane, long long (&raw)[4]) {
    if (kind == BQ_INT64 || kind == BQ_DOUBLE) {
        const int4* q = reinterpret_cast<const int4*>(static_cast<const char*>(ptr) + (base + 2 * lane) * 8);
        int4 a = ldg_stream(q);
        int4 b = ldg_stream(q + 32);
        raw[0] = pack64(a.x, a.y);
        raw[1] = pack64(a.z, a.w);
        raw[2] = pack64(b.x, b.y);
        raw[3] = pack64(b.z, b.w);
    } else {
        const int2* q = reinterpret_cast<const int2*>(static_cast<const char*>(ptr) + (base + 2 * lane) * 4);
        int2 a = ldg_stream2(q);
        int2 b = ldg_stream2(q + 32);
        if (kind == BQ_STRING) {
            raw[0] = static_cast<unsigned>(a.x);
            raw[1] = static_cast<unsigned>(a.y);
            raw[2] = static_cast<unsigned>(b.x);
            raw[3] = static_cast<unsigned>(b.y);
        } else {
            raw[0] = a.x;
            raw[1] = a.y;
            raw[2] = b.x;
            raw[3] = b.y;
        }
    }
}

// The same with each half predicated: a lane whose two rows of a half were ruled out beforehand (join-match bits) does not
// issue that load, so a 32-byte sector none of whose four rows survive is never fetched from HBM.
BQ_D void load_quad_pred(const void* ptr, int kind, size_t base, int lane, bool lo_half, bool hi_half, long long (&raw)[4]) {
    raw[0] = raw[1] = raw[2] = raw[3] = 0;
    if (kind == BQ_INT64 || kind == BQ_DOUBLE) {
        const int4* q = reinterpret_cast<const int4*>(static_cast<const char*>(ptr) + (base + 2 * lane) * 8);
        if (lo_half) {
            int4 a = ldg_stream(q);
            raw[0] = pack64(a.x, a.y);
            raw[1] = pack64(a.z, a.w);
        }
        if (hi_half) {
            int4 b = ldg_stream(q + 32);
            raw[2] = pack64(b.x, b.y);
            raw[3] = pack64(b.z, b.w);
        }
    } else {
        const int2* q = reinterpret_cast<const int2*>(static_cast<const char*>(ptr) + (base + 2 * lane) * 4);
        int2 a = make_int2(0, 0), b = make_int2(0, 0);
        if (lo_half) a = ldg_stream2(q);
        if (hi_half) b = ldg_stream2(q + 32);
        if (kind == BQ_STRING) {
            raw[0] = static_cast<unsigned>(a.x);
            raw[1] = static_cast<unsigned>(a.y);
            raw[2] = static_cast<unsigned>(b.x);
            raw[3] = static_cast<unsigned>(b.y);
        } else {
            raw[0] = a.x;
            raw[1] = a.y;
            raw[2] = b.x;
            raw[3] = b.y;
        }
    }
}

// ---- aggregate argument: numeric_binary + datum_as_double ---------------------------------------
BQ_D double eval_vexpr(const DVExpr& e, int op, int l_src, int r_src, long long a, int ka, long long b, int kb, int* err) {
    if (op == BQ_V_A) return as_double(a, ka);
    if (op == BQ_V_B) return as_double(b, kb);
    const long long imm = e.imm_is_f ? __double_as_longlong(e.imm_f) : e.imm_i;
    const int kimm = e.imm_is_f ? BQ_DOUBLE : BQ_INT64;
    long long l = a, r = b;
    int kl = ka, kr = kb;
    if (l_src == BQ_L_B) { l = b; kl = kb; }
    else if (l_src == BQ_L_IMM) { l = imm; kl = kimm; }
    if (r_src == BQ_R_A) { r = a; kr = ka; }
    else if (r_src == BQ_R_IMM) { r = imm; kr = kimm; }
    if (kl == BQ_DOUBLE || kr == BQ_DOUBLE) {       // src/exec/expression.cpp:34-44
        double x = as_double(l, kl), y = as_double(r, kr);
        switch (op) {
            case BQ_V_MUL: return __dmul_rn(x, y);
            case BQ_V_ADD: return __dadd_rn(x, y);
            case BQ_V_SUB: return __dsub_rn(x, y);
            default: return y == 0.0 ? __longlong_as_double(0x7FF0000000000000LL) : __ddiv_rn(x, y);
        }
    }
    long long z = 0;                                // src/exec/expression.cpp:45-56 (wraps like int64_t)
    switch (op) {
        case BQ_V_MUL: z = static_cast<long long>(static_cast<unsigned long long>(l) * static_cast<unsigned long long>(r)); break;
        case BQ_V_ADD: z = static_cast<long long>(static_cast<unsigned long long>(l) + static_cast<unsigned long long>(r)); break;
        case BQ_V_SUB: z = static_cast<long long>(static_cast<unsigned long long>(l) - static_cast<unsigned long long>(r)); break;
        default:
            if (r == 0) { *err = 1; z = 0; }
            else if (l == INT64_MIN && r == -1) z = INT64_MIN;
            else z = l / r;
    }
    return static_cast<double>(z);
}

// ---- group table claim (open addressing, linear probing) ----------------------------------------
// Partition-major variant: the table is split into 2^part_log2 regions; a key probes only inside the region of its
// partition, so rows ordered by partition touch one region (capacity / P slots) at a time.
BQ_D unsigned long long group_slot_part(long long* h_keys, unsigned long long h_mask, int part_log2, int part_shift, int* err,
                                        long long key) {
    if (key == kEmptyKey) return h_mask + 1;
    const unsigned long long hv = key_hash(static_cast<uint64_t>(key));
    const unsigned long long sub_mask = h_mask >> part_log2;
    const unsigned long long base = ((hv >> part_shift) & ((1ULL << part_log2) - 1)) * (sub_mask + 1);
    unsigned long long h = hv & sub_mask;
    for (unsigned long long probes = 0; probes <= sub_mask; ++probes) {
        long long cur = *reinterpret_cast<volatile long long*>(h_keys + base + h);
        if (cur == key) return base + h;
        if (cur == kEmptyKey) {
            long long prev = static_cast<long long>(atomicCAS(reinterpret_cast<unsigned long long*>(h_keys + base + h),
                                                              static_cast<unsigned long long>(kEmptyKey),
                                                              static_cast<unsigned long long>(key)));
            if (prev == kEmptyKey || prev == key) return base + h;
        }
        h = (h + 1) & sub_mask;
    }
    *err = 2;
    return h_mask + 1;
}

BQ_D unsigned long long group_slot(long long* h_keys, unsigned long long h_mask, int* err, long long key) {
    if (key == kEmptyKey) return h_mask + 1;        // spare slot for the one key that equals the marker
    unsigned long long h = key_hash(static_cast<uint64_t>(key)) & h_mask;
    for (unsigned long long probes = 0; probes <= h_mask; ++probes) {
        long long cur = *reinterpret_cast<volatile long long*>(h_keys + h);
        if (cur == key) return h;
        if (cur == kEmptyKey) {
            long long prev = static_cast<long long>(atomicCAS(reinterpret_cast<unsigned long long*>(h_keys + h),
                                                              static_cast<unsigned long long>(kEmptyKey),
                                                              static_cast<unsigned long long>(key)));
            if (prev == kEmptyKey || prev == key) return h;
        }
        h = (h + 1) & h_mask;
    }
    *err = 2;
    return h_mask + 1;
}

// XSHAPE: join kind and aggregate-argument forms known at compile time (specialised kernels).
//   bits 0-2 join kind (7 = run time) | bits 3-4 number of arguments | per argument 7 bits: op(3) l_src(2) r_src(2)
constexpr uint32_t kGenericX = 0xFFFFFFFFu;
constexpr uint32_t xshape(int jmode, int nv, int op0 = 0, int l0 = 0, int r0 = 0, int op1 = 0, int l1 = 0, int r1 = 0) {
    return static_cast<uint32_t>(jmode) | (static_cast<uint32_t>(nv) << 3) |
           (static_cast<uint32_t>(op0 | (l0 << 3) | (r0 << 5)) << 5) | (static_cast<uint32_t>(op1 | (l1 << 3) | (r1 << 5)) << 12);
}
template <uint32_t XSHAPE>
struct XShape {
    static constexpr bool generic = (XSHAPE == kGenericX);
    BQ_D static int jmode(const ScanParams& p) { return generic ? p.jmode : static_cast<int>(XSHAPE & 7u); }
    BQ_D static int nv(const ScanParams& p) { return generic ? p.nv : static_cast<int>((XSHAPE >> 3) & 3u); }
    BQ_D static int op(const ScanParams& p, int i) { return generic ? p.v[i].op : static_cast<int>((XSHAPE >> (5 + 7 * i)) & 7u); }
    BQ_D static int l_src(const ScanParams& p, int i) { return generic ? p.v[i].l_src : static_cast<int>((XSHAPE >> (8 + 7 * i)) & 3u); }
    BQ_D static int r_src(const ScanParams& p, int i) { return generic ? p.v[i].r_src : static_cast<int>((XSHAPE >> (10 + 7 * i)) & 3u); }
};

template <uint32_t SHAPE, uint32_t XSHAPE, int GMODE>
struct RowSink {
    using Sh = Shape<SHAPE>;
    using Xs = XShape<XSHAPE>;
    const ScanParams& p;
    double* s_sum0;
    double* s_sum1;
    unsigned* s_cnt;
    unsigned long long cnt = 0;
    double sum0 = 0.0, sum1 = 0.0;
    int err = 0;

    BQ_D RowSink(const ScanParams& pp, double* a, double* b, unsigned* c) : p(pp), s_sum0(a), s_sum1(b), s_cnt(c) {}

    BQ_D void values(long long a, long long b, double& v0, double& v1) {
        const int ka = Sh::present(p, S_A) ? Sh::kind(p, S_A) : BQ_INT64;
        const int kb = Sh::present(p, S_B) ? Sh::kind(p, S_B) : BQ_INT64;
        v0 = 0.0;
        v1 = 0.0;
        if (Xs::nv(p) > 0) v0 = eval_vexpr(p.v[0], Xs::op(p, 0), Xs::l_src(p, 0), Xs::r_src(p, 0), a, ka, b, kb, &err);
        if (Xs::nv(p) > 1) v1 = eval_vexpr(p.v[1], Xs::op(p, 1), Xs::l_src(p, 1), Xs::r_src(p, 1), a, ka, b, kb, &err);
    }

    // one qualifying (probe row, build row) pair
    BQ_D void add(long long key_raw, long long a, long long b) {
        double v0, v1;
        values(a, b, v0, v1);
        if (GMODE == G_NONE) {
            cnt += 1;
            sum0 = __dadd_rn(sum0, v0);
            sum1 = __dadd_rn(sum1, v1);
        } else if (GMODE == G_SMEM) {
            unsigned long long idx = static_cast<unsigned long long>(key_raw - p.key_min);
            if (idx < p.key_domain) {
                atomicAdd(s_cnt + idx, 1u);
                if (Xs::nv(p) > 0) atomicAdd(s_sum0 + idx, v0);
                if (Xs::nv(p) > 1) atomicAdd(s_sum1 + idx, v1);
            } else {
                err |= 4;       // key outside the catalog's [min,max]: stale statistics, the host re-plans
            }
        } else if (GMODE == G_DENSE) {
            unsigned long long idx = static_cast<unsigned long long>(key_raw - p.key_min);
            if (idx < p.key_domain) {
                if (p.need_count) {
                    atomicAdd(p.g_cnt + idx, 1ULL);
                    if (Xs::nv(p) > 0) atomicAdd(p.g_sum0 + idx, v0);
                } else {
                    // presence rides on the sum: slots start as -0.0 and v + 0.0 is never -0.0, so a touched slot
                    // can never read back as -0.0 (one atomic per row instead of two)
                    atomicAdd(p.g_sum0 + idx, __dadd_rn(v0, 0.0));
                }
                if (Xs::nv(p) > 1) atomicAdd(p.g_sum1 + idx, v1);
            } else {
                err |= 4;
            }
        } else {
            long long k = key_raw;
            if (Sh::kind(p, S_KEY) == BQ_DOUBLE && k == INT64_MIN) k = 0;   // -0.0 groups with +0.0
            unsigned long long idx = p.h_part_log2 > 0 ? group_slot_part(p.h_keys, p.h_mask, p.h_part_log2, p.h_part_shift, p.err, k)
                                                       : group_slot(p.h_keys, p.h_mask, p.err, k);
            if (p.need_count || idx > p.h_mask) atomicAdd(p.g_cnt + idx, 1ULL);     // else presence = the claimed key
            if (Xs::nv(p) > 0) atomicAdd(p.g_sum0 + idx, v0);
            if (Xs::nv(p) > 1) atomicAdd(p.g_sum1 + idx, v1);
        }
    }

    // branch-free accumulation for the global aggregate: a failed row adds +0.0 and 0
    BQ_D void add_masked(bool pass, long long a, long long b) {
        double v0, v1;
        values(a, b, v0, v1);
        cnt += pass ? 1ULL : 0ULL;
        sum0 = __dadd_rn(sum0, pass ? v0 : 0.0);
        sum1 = __dadd_rn(sum1, pass ? v1 : 0.0);
    }

    BQ_D void build_add(long long key_raw, long long a, long long b, unsigned brow) {
        if (Sh::present(p, S_KEY) && Sh::from_build(p, S_KEY)) key_raw = load_raw(p.s[S_KEY].ptr, Sh::kind(p, S_KEY), brow);
        if (Sh::present(p, S_A) && Sh::from_build(p, S_A)) a = load_raw(p.s[S_A].ptr, Sh::kind(p, S_A), brow);
        if (Sh::present(p, S_B) && Sh::from_build(p, S_B)) b = load_raw(p.s[S_B].ptr, Sh::kind(p, S_B), brow);
        add(key_raw, a, b);
    }

    // bitmap / direct-address probe: true = the row joins (brow = matched build row for DIRECT)
    BQ_D bool probe_dense(long long jk, unsigned& brow) {
        brow = 0;
        if (Sh::kind(p, S_JK) == BQ_DOUBLE) {
            if (jk == INT64_MIN) jk = 0;                                         // -0.0 == 0.0 (KeyEqual, :657)
            if ((jk & 0x7FFFFFFFFFFFFFFFLL) > 0x7FF0000000000000LL) return false;   // NaN never matches
        }
        unsigned long long idx = static_cast<unsigned long long>(jk - p.jk_min);
        if (idx >= p.jk_domain) return false;
        if (Xs::jmode(p) == BQ_JOIN_BITMAP) return (__ldg(p.j_bitmap + (idx >> 5)) >> (idx & 31)) & 1u;
        unsigned e = __ldg(p.j_direct + idx);
        brow = e - 1;
        return e != 0;
    }

    // hash probe: every matching build row contributes (duplicate build keys, src/exec/operator.cpp:802-816)
    BQ_D void probe_hash(long long key_raw, long long a, long long b, long long jk) {
        if (Sh::kind(p, S_JK) == BQ_DOUBLE) {
            if (jk == INT64_MIN) jk = 0;
            if ((jk & 0x7FFFFFFFFFFFFFFFLL) > 0x7FF0000000000000LL) return;
        }
        unsigned long long h = key_hash(static_cast<uint64_t>(jk)) & p.jh_mask;
        for (unsigned long long probes = 0; probes <= p.jh_mask; ++probes) {
            const JoinSlot e = load_join_slot(p.jh_slots + h);
            if (!e.row) return;
            if (e.key == jk) build_add(key_raw, a, b, e.row - 1);
            h = (h + 1) & p.jh_mask;
        }
    }

    // The same walk for a probe whose FIRST slot is already in registers: the vector loop issues the first-slot loads of a
    // lane's four rows together (four random sectors in flight per lane instead of one), then finishes each row here.  At the
    // tables' load factor (<= 0.5, usually far less) most probes end at that first slot: empty = no match.
    BQ_D void probe_hash_from(long long key_raw, long long a, long long b, long long jk, unsigned long long h, JoinSlot e) {
        for (unsigned long long probes = 0; probes <= p.jh_mask; ++probes) {
            if (!e.row) return;
            if (e.key == jk) build_add(key_raw, a, b, e.row - 1);
            h = (h + 1) & p.jh_mask;
            e = load_join_slot(p.jh_slots + h);
        }
    }

    // a row that passed every streamed range (scalar head/tail path and staged flushes)
    BQ_D void row(long long key_raw, long long a, long long b, long long jk) {
        if (Xs::jmode(p) == 0 || Xs::jmode(p) == BQ_JOIN_ROWBITS) {      // (match bits were tested by the caller)
            add(key_raw, a, b);
        } else if (Xs::jmode(p) == BQ_JOIN_HASH) {
            probe_hash(key_raw, a, b, jk);
        } else {
            unsigned brow;
            if (probe_dense(jk, brow)) build_add(key_raw, a, b, brow);
        }
    }
};

// One range test on a streamed value.  Ranges reaching the device are never empty (the host resolves those), so
// lo <= k <= hi is the single unsigned compare (k - lo) <= (hi - lo); 4-byte kinds compare in 32 bits.
struct RangeRegs {
    unsigned long long lo0, span0, lo1, span1;
    bool neg0, neg1;
    int nr;
};
BQ_D RangeRegs range_regs(const DSlot& s) {
    RangeRegs r;
    r.nr = s.nr;
    r.lo0 = static_cast<unsigned long long>(s.lo0);
    r.span0 = static_cast<unsigned long long>(s.hi0) - static_cast<unsigned long long>(s.lo0);
    r.lo1 = static_cast<unsigned long long>(s.lo1);
    r.span1 = static_cast<unsigned long long>(s.hi1) - static_cast<unsigned long long>(s.lo1);
    r.neg0 = s.neg0 != 0;
    r.neg1 = s.neg1 != 0;
    return r;
}
BQ_D bool range_pass(const RangeRegs& r, long long raw, int kind, int nr) {
    bool ok;
    if (kind == BQ_STRING || kind == BQ_DATE32) {
        const unsigned x = static_cast<unsigned>(raw);
        ok = ((x - static_cast<unsigned>(r.lo0)) <= static_cast<unsigned>(r.span0)) != r.neg0;
        if (nr > 1) ok = ok && (((x - static_cast<unsigned>(r.lo1)) <= static_cast<unsigned>(r.span1)) != r.neg1);
    } else {
        const unsigned long long x = static_cast<unsigned long long>(key_of(raw, kind));
        ok = ((x - r.lo0) <= r.span0) != r.neg0;
        if (nr > 1) ok = ok && (((x - r.lo1) <= r.span1) != r.neg1);
    }
    return ok;
}

constexpr int kStageCap = 64;     // staged qualifying rows per warp (flushed 32 at a time)

// RSHAPE: 2 bits per slot = how many ranges the slot carries (specialised kernels; the generic kernel reads it at run time)
constexpr uint32_t rshape_bits(int slot, int n_ranges) { return static_cast<uint32_t>(n_ranges) << (2 * slot); }

// resident CTAs per SM the compiler must leave room for.  The fused bitmap probe into a dense table fits six at 39 registers
// without a spill; measured against five at 48 registers on one box it is the same 6.53 ms - the probe is bound by the
// L1->L2 request rate (one random sector per row), not by warps in flight
constexpr int scan_min_blocks(uint32_t shape, uint32_t xs, int gmode, bool staged) {
    return shape == kGenericShape ? 2 : (xs == xshape(BQ_JOIN_BITMAP, 1, BQ_V_MUL, BQ_L_A, BQ_R_B) && gmode == G_DENSE && !staged) ? 6 : 4;
}
template <uint32_t SHAPE, uint32_t RSHAPE, uint32_t XSHAPE, int GMODE, bool STAGED>
__global__ void __launch_bounds__(kBlock, scan_min_blocks(SHAPE, XSHAPE, GMODE, STAGED)) k_scan(const __grid_constant__ ScanParams p) {
    using Sh = Shape<SHAPE>;
    using Xs = XShape<XSHAPE>;
    extern __shared__ double smem_dyn[];
    __shared__ double red_sum[kBlock / 32][2];
    __shared__ unsigned long long red_cnt[kBlock / 32];
    __shared__ int red_err;
    __shared__ bool is_last;
    // Qualifying rows are staged per warp and aggregated 32 at a time: at low selectivity the expensive part
    // (table update, join payload, atomics) then runs with full warps instead of once per row position.
    __shared__ long long st_key[STAGED ? kBlock / 32 : 1][STAGED ? kStageCap : 1];
    __shared__ long long st_a[STAGED ? kBlock / 32 : 1][STAGED ? kStageCap : 1];
    __shared__ long long st_b[STAGED ? kBlock / 32 : 1][STAGED ? kStageCap : 1];
    __shared__ long long st_x[STAGED ? kBlock / 32 : 1][STAGED ? kStageCap : 1];   // probe key (hash join) or matched build row

    double* s_sum0 = nullptr;
    double* s_sum1 = nullptr;
    unsigned* s_cnt = nullptr;
    if (GMODE == G_SMEM) {
        s_sum0 = smem_dyn;
        s_sum1 = smem_dyn + p.key_domain;
        s_cnt = reinterpret_cast<unsigned*>(smem_dyn + 2 * p.key_domain);
        for (unsigned long long i = threadIdx.x; i < p.key_domain; i += blockDim.x) {
            s_sum0[i] = 0.0;
            s_sum1[i] = 0.0;
            s_cnt[i] = 0u;
        }
    }
    if (threadIdx.x == 0) red_err = 0;
    __syncthreads();

    RowSink<SHAPE, XSHAPE, GMODE> sink(p, s_sum0, s_sum1, s_cnt);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const size_t warps_total = static_cast<size_t>(gridDim.x) * (kBlock / 32);
    const size_t warp_global = static_cast<size_t>(blockIdx.x) * (kBlock / 32) + warp;
    // global aggregate without a multi-match probe: branch-free accumulation in registers, no staging
    const bool direct_path = (GMODE == G_NONE) && !STAGED;
    const bool dense_join = Xs::jmode(p) == BQ_JOIN_BITMAP || Xs::jmode(p) == BQ_JOIN_DIRECT;

    RangeRegs rr[N_SLOTS];
#pragma unroll
    for (int s = 0; s < N_SLOTS; ++s)
        if (Sh::streamed(p, s)) rr[s] = range_regs(p.s[s]);
    int staged = 0;       // warp-uniform
    // number of ranges per slot: a compile-time constant in the specialised kernels
    auto nr_of = [&](int s) -> int { return Sh::generic ? rr[s].nr : static_cast<int>((RSHAPE >> (2 * s)) & 3u); };
    const long long* mask = Sh::generic ? p.mask : nullptr;      // specialised kernels never carry a mask column

    auto flush32 = [&](int first, int n) {
        // lanes [0,n) take staged entries [first, first+n)
        if (!STAGED) return;
        __syncwarp();
        if (lane < n) {
            const int e = first + lane;
            if (Xs::jmode(p) == BQ_JOIN_HASH) sink.probe_hash(st_key[warp][e], st_a[warp][e], st_b[warp][e], st_x[warp][e]);
            else if (dense_join) sink.build_add(st_key[warp][e], st_a[warp][e], st_b[warp][e], static_cast<unsigned>(st_x[warp][e]));
            else sink.add(st_key[warp][e], st_a[warp][e], st_b[warp][e]);
        }
        __syncwarp();
    };

    // match bits are fetched one trip ahead, so that the (predicated) column loads never wait for them
    uint4 bits_next = make_uint4(0, 0, 0, 0);
    if (Xs::jmode(p) == BQ_JOIN_ROWBITS && warp_global < p.n_chunks)
        bits_next = __ldg(reinterpret_cast<const uint4*>(p.row_bits + ((p.vec_begin + warp_global * 128) >> 5)));

    // ---- vector region: whole 128-row chunks -----------------------------------------------------
    for (size_t c = warp_global; c < p.n_chunks; c += warps_total) {
        const size_t base = p.vec_begin + c * 128;
        long long raw[N_SLOTS][4];
        long long mk[4] = {1, 1, 1, 1};
        bool pass[4] = {true, true, true, true};
        if (Xs::jmode(p) == BQ_JOIN_ROWBITS) {
            // join-match bits of the chunk's 128 rows (fetched one trip ahead): one 16-byte broadcast load; the bits ARE
            // the join, and only the halves with a match are read
            const uint4 w = bits_next;
            if (c + warps_total < p.n_chunks) bits_next = __ldg(reinterpret_cast<const uint4*>(p.row_bits + ((base + warps_total * 128) >> 5)));
            const unsigned lo_w = lane < 16 ? w.x : w.y, hi_w = lane < 16 ? w.z : w.w;
            const int sh = (2 * lane) & 31;
            pass[0] = (lo_w >> sh) & 1u;
            pass[1] = (lo_w >> (sh + 1)) & 1u;
            pass[2] = (hi_w >> sh) & 1u;
            pass[3] = (hi_w >> (sh + 1)) & 1u;
            const bool all = p.rowbits_skip == 0;
#pragma unroll
            for (int s = 0; s < N_SLOTS; ++s)
                if (Sh::streamed(p, s)) load_quad_pred(p.s[s].ptr, Sh::kind(p, s), base, lane, all || pass[0] || pass[1], all || pass[2] || pass[3], raw[s]);
        } else {
            // phase 1: issue every load of the chunk
#pragma unroll
            for (int s = 0; s < N_SLOTS; ++s) {
                if (Sh::streamed(p, s)) load_quad(p.s[s].ptr, Sh::kind(p, s), base, lane, raw[s]);
            }
        }
        if (mask) load_quad(mask, BQ_INT64, base, lane, mk);
        // phase 2: range tests
#pragma unroll
        for (int s = 0; s < N_SLOTS; ++s) {
            if (Sh::streamed(p, s) && nr_of(s) > 0) {
#pragma unroll
                for (int r = 0; r < 4; ++r) pass[r] = pass[r] && range_pass(rr[s], raw[s][r], Sh::kind(p, s), nr_of(s));
            }
        }
        if (mask) {
#pragma unroll
            for (int r = 0; r < 4; ++r) pass[r] = pass[r] && (mk[r] != 0);
        }
        // phase 2b: bitmap / direct-address probes of the survivors, four independent loads in flight
        unsigned brow[4] = {0, 0, 0, 0};
        if (dense_join) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (pass[r]) pass[r] = sink.probe_dense(raw[S_JK][r], brow[r]);
        }
        // phase 2c: hash-table probes straight from registers (dense / global aggregates: nothing to stage for)
        if (Xs::jmode(p) == BQ_JOIN_HASH && !STAGED) {
            JoinSlot first[4];
            unsigned long long hh[4];
            long long jkc[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                jkc[r] = raw[S_JK][r];
                if (Sh::kind(p, S_JK) == BQ_DOUBLE) {
                    if (jkc[r] == INT64_MIN) jkc[r] = 0;                                                  // -0.0 == 0.0 (KeyEqual, :657)
                    if ((jkc[r] & 0x7FFFFFFFFFFFFFFFLL) > 0x7FF0000000000000LL) pass[r] = false;         // NaN never matches
                }
                hh[r] = key_hash(static_cast<uint64_t>(jkc[r])) & p.jh_mask;
                first[r].row = 0;
                first[r].key = 0;
            }
            if (p.jh_occ) {
                // home slot empty = key absent (linear probing never skips an empty slot): answered from the L2-resident
                // occupancy bits, four loads in flight, before any table sector is requested from HBM
                unsigned occ[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) occ[r] = pass[r] ? __ldg(p.jh_occ + (hh[r] >> 5)) : 0u;
#pragma unroll
                for (int r = 0; r < 4; ++r) pass[r] = pass[r] && ((occ[r] >> (hh[r] & 31)) & 1u);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (pass[r]) first[r] = load_join_slot(p.jh_slots + hh[r]);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (pass[r] && first[r].row)
                    sink.probe_hash_from(Sh::streamed(p, S_KEY) ? raw[S_KEY][r] : 0, Sh::streamed(p, S_A) ? raw[S_A][r] : 0,
                                         Sh::streamed(p, S_B) ? raw[S_B][r] : 0, jkc[r], hh[r], first[r]);
            }
            continue;
        }
        // phase 3: aggregate the survivors
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const long long kr = Sh::streamed(p, S_KEY) ? raw[S_KEY][r] : 0;
            long long a = Sh::streamed(p, S_A) ? raw[S_A][r] : 0;
            long long b = Sh::streamed(p, S_B) ? raw[S_B][r] : 0;
            const long long jk = Sh::streamed(p, S_JK) ? raw[S_JK][r] : 0;
            if (direct_path) {
                if (Sh::present(p, S_A) && Sh::from_build(p, S_A)) a = pass[r] ? load_raw(p.s[S_A].ptr, Sh::kind(p, S_A), brow[r]) : 0;
                if (Sh::present(p, S_B) && Sh::from_build(p, S_B)) b = pass[r] ? load_raw(p.s[S_B].ptr, Sh::kind(p, S_B), brow[r]) : 0;
                sink.add_masked(pass[r], a, b);
            } else if (!STAGED) {
                // most rows qualify (join-only pipelines): update the table straight from registers
                if (pass[r]) {
                    if (dense_join) sink.build_add(kr, a, b, brow[r]);
                    else sink.add(kr, a, b);
                }
            } else {
                const unsigned ballot = __ballot_sync(0xffffffffu, pass[r]);
                if (ballot) {
                    if (pass[r]) {
                        const int pos = staged + __popc(ballot & lt_mask);
                        st_key[warp][pos] = kr;
                        st_a[warp][pos] = a;
                        st_b[warp][pos] = b;
                        st_x[warp][pos] = Xs::jmode(p) == BQ_JOIN_HASH ? jk : static_cast<long long>(brow[r]);
                    }
                    staged += __popc(ballot);
                    if (staged >= 32) {
                        flush32(staged - 32, 32);
                        staged -= 32;
                    }
                }
            }
        }
    }
    if (STAGED && staged > 0) flush32(0, staged);

    // ---- head and tail rows (fewer than 4 + 128): one row per thread, CTA 0 -----------------------
    if (blockIdx.x == 0) {
        const size_t vec_end = p.vec_begin + p.n_chunks * 128;
        const size_t n_head = p.vec_begin - p.row_begin;
        const size_t n_tail = p.row_end - vec_end;
        for (size_t t = threadIdx.x; t < n_head + n_tail; t += blockDim.x) {
            const size_t i = t < n_head ? p.row_begin + t : vec_end + (t - n_head);
            bool ok = true;
            long long val[N_SLOTS];
#pragma unroll
            for (int s = 0; s < N_SLOTS; ++s) {
                val[s] = 0;
                if (Sh::streamed(p, s)) {
                    val[s] = load_raw(p.s[s].ptr, Sh::kind(p, s), i);
                    if (nr_of(s) > 0) ok = ok && range_pass(rr[s], val[s], Sh::kind(p, s), nr_of(s));
                }
            }
            if (mask && __ldg(mask + i) == 0) ok = false;
            if (Xs::jmode(p) == BQ_JOIN_ROWBITS && !((__ldg(p.row_bits + (i >> 5)) >> (i & 31)) & 1u)) ok = false;
            if (ok) sink.row(val[S_KEY], val[S_A], val[S_B], val[S_JK]);
        }
    }

    // ---- epilogue ----------------------------------------------------------------------------------
    if (sink.err) atomicOr(&red_err, sink.err);
    if (GMODE == G_NONE) {
        unsigned long long c = warp_sum(sink.cnt);
        double s0 = warp_sum(sink.sum0), s1 = warp_sum(sink.sum1);
        if (lane == 0) {
            red_cnt[warp] = c;
            red_sum[warp][0] = s0;
            red_sum[warp][1] = s1;
        }
    }
    __syncthreads();
    if (GMODE == G_SMEM) {
        // merge the CTA's table into the global one: one red per touched group per CTA
        for (unsigned long long i = threadIdx.x; i < p.key_domain; i += blockDim.x) {
            unsigned c = s_cnt[i];
            if (c) {
                atomicAdd(p.g_cnt + i, static_cast<unsigned long long>(c));
                if (Xs::nv(p) > 0) atomicAdd(p.g_sum0 + i, s_sum0[i]);
                if (Xs::nv(p) > 1) atomicAdd(p.g_sum1 + i, s_sum1[i]);
            }
        }
    }
    if (threadIdx.x == 0) {
        if (red_err) atomicOr(p.err, red_err);
        if (GMODE == G_NONE) {
            unsigned long long c = 0;
            double s0 = 0.0, s1 = 0.0;
            for (int w = 0; w < kBlock / 32; ++w) {     // fixed order: deterministic per grid size
                c += red_cnt[w];
                s0 = __dadd_rn(s0, red_sum[w][0]);
                s1 = __dadd_rn(s1, red_sum[w][1]);
            }
            p.part_cnt[blockIdx.x] = c;
            p.part_sum[2 * blockIdx.x] = s0;
            p.part_sum[2 * blockIdx.x + 1] = s1;
            __threadfence();
            unsigned t = atomicAdd(p.ticket, 1u);
            is_last = (t == gridDim.x - 1);
        }
    }
    if (GMODE == G_NONE) {
        __syncthreads();
        if (is_last && threadIdx.x == 0) {
            // the last CTA folds the per-CTA partials in CTA order into slot 0 of the global state
            unsigned long long c = 0;
            double s0 = 0.0, s1 = 0.0;
            for (unsigned b = 0; b < gridDim.x; ++b) {
                c += *reinterpret_cast<volatile unsigned long long*>(p.part_cnt + b);
                s0 = __dadd_rn(s0, *reinterpret_cast<volatile double*>(p.part_sum + 2 * b));
                s1 = __dadd_rn(s1, *reinterpret_cast<volatile double*>(p.part_sum + 2 * b + 1));
            }
            p.g_cnt[0] = c;
            p.g_sum0[0] = s0;
            p.g_sum1[0] = s1;
            *p.ticket = 0;
        }
    }
}

// ---- emit: presence bits over the state, then one thread per surviving group ---------------------
__global__ void __launch_bounds__(kBlock) k_presence_bits(const unsigned long long* __restrict__ cnt, size_t n,
                                                          unsigned* __restrict__ bits) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    bool on = i < n && cnt[i] != 0ULL;
    unsigned b = __ballot_sync(0xffffffffu, on);
    if ((threadIdx.x & 31) == 0 && (i >> 5) < (n + 31) / 32) bits[i >> 5] = b;
}

// presence without counts: dense tables whose sum slots started as -0.0, hash tables by their claimed keys
__global__ void __launch_bounds__(kBlock) k_presence_sum(const double* __restrict__ sum0, size_t n, unsigned* __restrict__ bits) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    bool on = i < n && __double_as_longlong(sum0[i]) != INT64_MIN;
    unsigned b = __ballot_sync(0xffffffffu, on);
    if ((threadIdx.x & 31) == 0 && (i >> 5) < (n + 31) / 32) bits[i >> 5] = b;
}
__global__ void __launch_bounds__(kBlock) k_presence_key(const long long* __restrict__ keys, const unsigned long long* __restrict__ cnt,
                                                         size_t n, unsigned* __restrict__ bits) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    bool on = i < n && (i + 1 < n ? keys[i] != kEmptyKey : cnt[i] != 0ULL);      // last entry = the spare slot
    unsigned b = __ballot_sync(0xffffffffu, on);
    if ((threadIdx.x & 31) == 0 && (i >> 5) < (n + 31) / 32) bits[i >> 5] = b;
}

struct EmitParams {
    const unsigned* rowids;
    size_t n;
    int gmode;
    int key_type;
    long long key_min;
    const long long* h_keys;
    const unsigned long long* cnt;
    const double* sum0;
    const double* sum1;
    void* out_key;
    int n_out;
    int func[BQ_MAX_AGG_OUT];
    int v[BQ_MAX_AGG_OUT];
    int as_int[BQ_MAX_AGG_OUT];
    void* out[BQ_MAX_AGG_OUT];
};

__global__ void __launch_bounds__(kBlock) k_emit(const __grid_constant__ EmitParams p) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    size_t g = p.rowids ? p.rowids[i] : i;
    if (p.out_key) {
        long long k = (p.gmode == G_HASH) ? p.h_keys[g] : p.key_min + static_cast<long long>(g);
        switch (p.key_type) {
            case BQ_INT64:
            case BQ_DOUBLE: static_cast<long long*>(p.out_key)[i] = k; break;
            case BQ_STRING: static_cast<unsigned*>(p.out_key)[i] = static_cast<unsigned>(k); break;
            default: static_cast<int*>(p.out_key)[i] = static_cast<int>(k); break;
        }
    }
    unsigned long long c = p.cnt[g];
    for (int o = 0; o < p.n_out; ++o) {
        double s = p.v[o] == 0 ? p.sum0[g] : p.sum1[g];
        if (p.func[o] == BQ_AGG_COUNT) {
            static_cast<long long*>(p.out[o])[i] = static_cast<long long>(c);
        } else if (p.func[o] == BQ_AGG_SUM) {
            if (p.as_int[o]) static_cast<long long*>(p.out[o])[i] = static_cast<long long>(s);   // :1044
            else static_cast<double*>(p.out[o])[i] = s;
        } else {
            static_cast<double*>(p.out[o])[i] = c == 0 ? 0.0 : __ddiv_rn(s, static_cast<double>(c));   // :1047
        }
    }
}

// Dense / global states of up to a million slots: CTAs of 1024 threads find the groups that exist in their 16384-slot chunk,
// number them in slot order and write the output columns - presence test, ordered compaction and emit in one launch (two
// when there are several chunks: a counting launch first, whose per-chunk totals give every CTA its base), with the group
// count and the scan's error word left for ONE host round trip.  The general path below spends five launches and needs the
// count on the host BEFORE it can emit; for Q1's 31 groups that was most of the time after the scan.
constexpr size_t kFinishChunk = 16384;
constexpr size_t kFinishMaxSlots = kFinishChunk * 64;
BQ_D bool slot_present(const EmitParams& p, int presence, size_t g) {
    return presence == 1 ? (__double_as_longlong(p.sum0[g]) != INT64_MIN) : (p.cnt[g] != 0);
}
__global__ void __launch_bounds__(1024) k_finish_count(const __grid_constant__ EmitParams p, size_t slots, int presence,
                                                       unsigned* __restrict__ chunk_count) {
    __shared__ unsigned warp_tot[32];
    const size_t lo = blockIdx.x * kFinishChunk, hi = lo + kFinishChunk < slots ? lo + kFinishChunk : slots;
    unsigned c = 0;
    for (size_t g = lo + threadIdx.x; g < hi; g += 1024) c += slot_present(p, presence, g) ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int w = 0; w < 32; ++w) t += warp_tot[w];
        chunk_count[blockIdx.x] = t;
    }
}
// chunk_count == nullptr: a single chunk (grid of one).  Slot g of the chunk belongs to (round g / 1024, thread g % 1024), so
// every load is coalesced; the sixteen rounds' ballots stay in registers, their per-warp counts (16 x 32 values) are scanned
// ONCE by the first 512 threads, and the rounds are then emitted without another barrier.  (One barrier pair per round cost
// 2 us each; sixteen consecutive slots per thread made the loads and stores strided: 56 us for Q2's 100 000 skus.)
__global__ void __launch_bounds__(1024) k_finish_small(const __grid_constant__ EmitParams p, size_t slots, int presence,
                                                       const unsigned* __restrict__ chunk_count, unsigned long long* __restrict__ n_groups) {
    constexpr int kRounds = static_cast<int>(kFinishChunk / 1024);
    __shared__ unsigned cnt[kRounds * 32];          // [round][warp] -> exclusive offsets after the scan
    __shared__ unsigned round_tot[kRounds];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t lo = blockIdx.x * kFinishChunk, hi = lo + kFinishChunk < slots ? lo + kFinishChunk : slots;
    unsigned ballots[kRounds];
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const size_t g = lo + static_cast<size_t>(r) * 1024 + threadIdx.x;
        ballots[r] = __ballot_sync(0xffffffffu, g < hi && slot_present(p, presence, g));
        if (lane == 0) cnt[r * 32 + warp] = __popc(ballots[r]);
    }
    __syncthreads();
    unsigned v = 0, incl = 0;
    if (threadIdx.x < kRounds * 32) {               // warp w of these threads holds round w's 32 warp counts
        v = cnt[threadIdx.x];
        incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) round_tot[warp] = incl;
    }
    __syncthreads();
    if (threadIdx.x < kRounds * 32) {
        unsigned pre = 0;
        for (int w = 0; w < warp; ++w) pre += round_tot[w];
        cnt[threadIdx.x] = pre + incl - v;
    }
    unsigned before = 0, total = 0;
    if (chunk_count)
        for (unsigned q = 0; q < blockIdx.x; ++q) before += chunk_count[q];
    for (int r = 0; r < kRounds; ++r) total += round_tot[r];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        if (!((ballots[r] >> lane) & 1u)) continue;
        const size_t g = lo + static_cast<size_t>(r) * 1024 + threadIdx.x;
        const size_t i = static_cast<size_t>(before) + cnt[r * 32 + warp] + __popc(ballots[r] & ((1u << lane) - 1u));
        if (p.out_key) {
            const long long k = p.key_min + static_cast<long long>(g);
            switch (p.key_type) {
                case BQ_INT64:
                case BQ_DOUBLE: static_cast<long long*>(p.out_key)[i] = k; break;
                case BQ_STRING: static_cast<unsigned*>(p.out_key)[i] = static_cast<unsigned>(k); break;
                default: static_cast<int*>(p.out_key)[i] = static_cast<int>(k); break;
            }
        }
        const unsigned long long cg = p.cnt[g];
        for (int o = 0; o < p.n_out; ++o) {
            const double sv = p.v[o] == 0 ? p.sum0[g] : p.sum1[g];
            if (p.func[o] == BQ_AGG_COUNT) static_cast<long long*>(p.out[o])[i] = static_cast<long long>(cg);
            else if (p.func[o] == BQ_AGG_SUM) {
                if (p.as_int[o]) static_cast<long long*>(p.out[o])[i] = static_cast<long long>(sv);      // :1044
                else static_cast<double*>(p.out[o])[i] = sv;
            } else {
                static_cast<double*>(p.out[o])[i] = cg == 0 ? 0.0 : __ddiv_rn(sv, static_cast<double>(cg));   // :1047
            }
        }
    }
    if (threadIdx.x == 0 && blockIdx.x == gridDim.x - 1) *n_groups = static_cast<unsigned long long>(before) + total;
}

__global__ void __launch_bounds__(kBlock) k_fill_keys(long long* __restrict__ keys, size_t n, long long v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) keys[i] = v;
}

// ---- shape registry -------------------------------------------------------------------------------
using ScanKernel = void (*)(const ScanParams);
struct ShapeEntry {
    uint32_t shape;
    uint32_t rshape;
    uint32_t xshape;
    int gmode;
    bool staged;
    ScanKernel fn;
};
#define BQ_SHAPE(shape, rshape, xs, gmode, staged) {shape, rshape, xs, gmode, staged, k_scan<shape, rshape, xs, gmode, staged>}
constexpr uint32_t kX_A = xshape(0, 1, BQ_V_A);                                   // SUM(a), no join
constexpr uint32_t kX_AB = xshape(0, 2, BQ_V_A, 0, 0, BQ_V_B);                    // SUM(a), SUM(b), no join
constexpr uint32_t kX_Q2 = xshape(BQ_JOIN_BITMAP, 1, BQ_V_MUL, BQ_L_A, BQ_R_B);   // bitmap probe, SUM(a * b)
constexpr uint32_t kX_J5 = xshape(BQ_JOIN_DIRECT, 1, BQ_V_MUL, BQ_L_A, BQ_R_B);   // direct probe, SUM(a * b.w)
constexpr uint32_t kX_Q2H = xshape(BQ_JOIN_HASH, 1, BQ_V_MUL, BQ_L_A, BQ_R_B);    // hash-table probe, SUM(a * b)
constexpr uint32_t kX_Q2B = xshape(BQ_JOIN_ROWBITS, 1, BQ_V_MUL, BQ_L_A, BQ_R_B); // precomputed match bits, SUM(a * b)
constexpr uint32_t kR_Q1 = rshape_bits(S_KEY, 1) | rshape_bits(S_P0, 1);     // one (merged) range on the date, one on status
constexpr uint32_t kR_P0 = rshape_bits(S_P0, 1);                             // filter sweep: one range on the predicate column
constexpr uint32_t kR_A = rshape_bits(S_A, 1);
constexpr uint32_t kR_NONE = 0;

// Q1: status (STRING range) AND order_date (DATE32 ranges, also the group key), SUM(total DOUBLE)
constexpr uint32_t kShapeQ1 = shape_bits(S_KEY, BQ_DATE32, false) | shape_bits(S_A, BQ_DOUBLE, false) |
                              shape_bits(S_P0, BQ_STRING, false);
// filter sweep: predicate column of each type, SUM(v DOUBLE) [+ SUM(w INT64)]
constexpr uint32_t kShapeF_I64 = shape_bits(S_A, BQ_DOUBLE, false) | shape_bits(S_P0, BQ_INT64, false);
constexpr uint32_t kShapeF_F64 = shape_bits(S_A, BQ_DOUBLE, false) | shape_bits(S_P0, BQ_DOUBLE, false);
constexpr uint32_t kShapeF_STR = shape_bits(S_A, BQ_DOUBLE, false) | shape_bits(S_P0, BQ_STRING, false);
constexpr uint32_t kShapeF_DATE = shape_bits(S_A, BQ_DOUBLE, false) | shape_bits(S_P0, BQ_DATE32, false);
constexpr uint32_t kShapeF2_I64 = kShapeF_I64 | shape_bits(S_B, BQ_INT64, false);
constexpr uint32_t kShapeF2_F64 = kShapeF_F64 | shape_bits(S_B, BQ_INT64, false);
constexpr uint32_t kShapeF2_STR = kShapeF_STR | shape_bits(S_B, BQ_INT64, false);
constexpr uint32_t kShapeF2_DATE = kShapeF_DATE | shape_bits(S_B, BQ_INT64, false);
// predicate on the summed column itself
constexpr uint32_t kShapeA_F64 = shape_bits(S_A, BQ_DOUBLE, false);
// Q2: probe l.order_id, GROUP BY l.sku (INT64 or STRING), SUM(l.qty INT64 * l.price DOUBLE)
constexpr uint32_t kShapeQ2 = shape_bits(S_KEY, BQ_INT64, false) | shape_bits(S_A, BQ_INT64, false) |
                              shape_bits(S_B, BQ_DOUBLE, false) | shape_bits(S_JK, BQ_INT64, false);
constexpr uint32_t kShapeQ2S = shape_bits(S_KEY, BQ_STRING, false) | shape_bits(S_A, BQ_INT64, false) |
                               shape_bits(S_B, BQ_DOUBLE, false) | shape_bits(S_JK, BQ_INT64, false);
// Q2 after bq_join_probe_bits: the probe key is no longer read
constexpr uint32_t kShapeQ2B = shape_bits(S_KEY, BQ_INT64, false) | shape_bits(S_A, BQ_INT64, false) | shape_bits(S_B, BQ_DOUBLE, false);
constexpr uint32_t kShapeQ2SB = shape_bits(S_KEY, BQ_STRING, false) | shape_bits(S_A, BQ_INT64, false) | shape_bits(S_B, BQ_DOUBLE, false);
// high-cardinality GROUP BY k (INT64) SUM/COUNT/AVG(v DOUBLE)
constexpr uint32_t kShapeGB = shape_bits(S_KEY, BQ_INT64, false) | shape_bits(S_A, BQ_DOUBLE, false);
// skewed join: probe p.k, SUM(p.v * b.w) with b.w from the build side
constexpr uint32_t kShapeJ5 = shape_bits(S_A, BQ_DOUBLE, false) | shape_bits(S_B, BQ_DOUBLE, true) |
                              shape_bits(S_JK, BQ_INT64, false);

static const ShapeEntry kShapes[] = {
    BQ_SHAPE(kShapeQ1, kR_Q1, kX_A, G_SMEM, true),
    BQ_SHAPE(kShapeF_I64, kR_P0, kX_A, G_NONE, false),    BQ_SHAPE(kShapeF_F64, kR_P0, kX_A, G_NONE, false),
    BQ_SHAPE(kShapeF_STR, kR_P0, kX_A, G_NONE, false),    BQ_SHAPE(kShapeF_DATE, kR_P0, kX_A, G_NONE, false),
    BQ_SHAPE(kShapeF2_I64, kR_P0, kX_AB, G_NONE, false),  BQ_SHAPE(kShapeF2_F64, kR_P0, kX_AB, G_NONE, false),
    BQ_SHAPE(kShapeF2_STR, kR_P0, kX_AB, G_NONE, false),  BQ_SHAPE(kShapeF2_DATE, kR_P0, kX_AB, G_NONE, false),
    BQ_SHAPE(kShapeA_F64, kR_A, kX_A, G_NONE, false),
    BQ_SHAPE(kShapeQ2, kR_NONE, kX_Q2, G_DENSE, false),   BQ_SHAPE(kShapeQ2, kR_NONE, kX_Q2, G_HASH, true),
    BQ_SHAPE(kShapeQ2S, kR_NONE, kX_Q2, G_DENSE, false),  BQ_SHAPE(kShapeQ2S, kR_NONE, kX_Q2, G_SMEM, true),
    BQ_SHAPE(kShapeQ2, kR_NONE, kX_Q2H, G_DENSE, false),  BQ_SHAPE(kShapeJ5, kR_NONE, xshape(BQ_JOIN_HASH, 1, BQ_V_MUL, BQ_L_A, BQ_R_B), G_NONE, false),
    BQ_SHAPE(kShapeQ2B, kR_NONE, kX_Q2B, G_DENSE, false), BQ_SHAPE(kShapeQ2SB, kR_NONE, kX_Q2B, G_DENSE, false),
    BQ_SHAPE(kShapeGB, kR_NONE, kX_A, G_HASH, true),      BQ_SHAPE(kShapeGB, kR_NONE, kX_A, G_DENSE, false),
    BQ_SHAPE(kShapeJ5, kR_NONE, kX_J5, G_NONE, false),
    // any other slot / range / argument layout, join kind or a mask column: same source, run-time flags
    BQ_SHAPE(kGenericShape, 0, kGenericX, G_NONE, false),  BQ_SHAPE(kGenericShape, 0, kGenericX, G_NONE, true),
    BQ_SHAPE(kGenericShape, 0, kGenericX, G_SMEM, true),   BQ_SHAPE(kGenericShape, 0, kGenericX, G_DENSE, false),
    BQ_SHAPE(kGenericShape, 0, kGenericX, G_DENSE, true),  BQ_SHAPE(kGenericShape, 0, kGenericX, G_HASH, true),
};

static ScanKernel pick_kernel(uint32_t shape, uint32_t rshape, uint32_t xs, bool has_mask, int gmode, bool staged, bool* specialised) {
    for (const auto& e : kShapes)
        if (!has_mask && e.shape == shape && e.rshape == rshape && e.xshape == xs && e.gmode == gmode && e.staged == staged) {
            *specialised = true;
            return e.fn;
        }
    for (const auto& e : kShapes)
        if (e.shape == kGenericShape && e.gmode == gmode && e.staged == staged) {
            *specialised = false;
            return e.fn;
        }
    throw std::runtime_error("no scan kernel for group mode");
}

// Clamp a slot's ranges to its column's domain, drop always-true ranges; returns true when some range can never pass.
bool normalise_slot(DSlot& d) {
    bool never = false;
    if (!d.ptr || d.nr == 0) return false;
    long long dmin = INT64_MIN, dmax = INT64_MAX;
    if (d.kind == BQ_STRING) { dmin = 0; dmax = 0xFFFFFFFFLL; }
    if (d.kind == BQ_DATE32) { dmin = INT32_MIN; dmax = INT32_MAX; }
    long long lo[2] = {d.lo0, d.lo1}, hi[2] = {d.hi0, d.hi1};
    int neg[2] = {d.neg0, d.neg1};
    int kept = 0;
    for (int i = 0; i < d.nr; ++i) {
        long long l = lo[i] < dmin ? dmin : lo[i], h = hi[i] > dmax ? dmax : hi[i];
        const bool empty = l > h, full = (l == dmin && h == dmax);
        if (!neg[i] ? empty : full) never = true;
        if (!neg[i] ? full : empty) continue;          // always true
        if (empty) continue;
        lo[kept] = l; hi[kept] = h; neg[kept] = neg[i];
        ++kept;
    }
    d.nr = kept;
    d.lo0 = lo[0]; d.hi0 = hi[0]; d.neg0 = neg[0];
    d.lo1 = lo[1]; d.hi1 = hi[1]; d.neg1 = neg[1];
    return never;
}

DSlot make_dslot(const bq_slot& s, size_t need_rows, const char* what) {
    DSlot d{};
    if (!s.col) return d;
    if (!s.from_build && s.col->n < need_rows)
        throw std::runtime_error(std::string("column bound to slot '") + what + "' is shorter than the scanned row range");
    if (s.n_ranges < 0 || s.n_ranges > 2) throw std::runtime_error("a slot carries at most two ranges");
    d.ptr = s.col->ptr;
    d.kind = s.col->type;
    d.nr = s.n_ranges;
    d.lo0 = s.r[0].lo; d.hi0 = s.r[0].hi; d.neg0 = s.r[0].neg;
    d.lo1 = s.r[1].lo; d.hi1 = s.r[1].hi; d.neg1 = s.r[1].neg;
    d.from_build = s.from_build;
    return d;
}

static size_t next_pow2(size_t v) {
    size_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

struct AggState {
    int gmode = G_NONE;
    size_t slots = 1;              // entries of cnt/sum arrays
    int key_type = BQ_INT64;
    bool has_key = false;
    long long key_min = 0;
    unsigned long long* cnt = nullptr;
    double* sum0 = nullptr;
    double* sum1 = nullptr;
    long long* h_keys = nullptr;
    void* block = nullptr;         // one allocation behind the arrays
    int* err = nullptr;            // device error word (division by zero, table overflow, stale statistics)
    int presence = 0;              // 0 by count, 1 by sum0 != -0.0 (dense, no counts), 2 by claimed key (hash, no counts)
    int arrays = 3;                // how many of sum0 | cnt | sum1 (in this order behind the 16-byte header) carry data
    bq_ctx* ctx = nullptr;

    ~AggState() { dev_free(ctx, block); }
};

// Runs the fused kernel for `spec`, leaving the aggregate state on the device.
static void run_scan(bq_ctx* ctx, const bq_scan_spec* spec, AggState& st, bool need_count) {
    if (spec->row_end < spec->row_begin) throw std::runtime_error("bad row range");
    if (spec->row_end > 0xFFFFFFFFull) throw std::runtime_error("row ids are 32-bit: at most 2^32 rows per scan");
    ScanParams p{};
    const bq_slot* slots[N_SLOTS] = {&spec->key, &spec->a, &spec->b, &spec->pred[0], &spec->pred[1], &spec->pred[2], &spec->jkey};
    static const char* names[N_SLOTS] = {"key", "a", "b", "pred0", "pred1", "pred2", "jkey"};
    uint32_t shape = 0;
    for (int s = 0; s < N_SLOTS; ++s) {
        p.s[s] = make_dslot(*slots[s], spec->row_end, names[s]);
        if (slots[s]->col) {
            p.present |= 1u << s;
            shape |= shape_bits(s, p.s[s].kind, p.s[s].from_build != 0);
            if (p.s[s].from_build && p.s[s].nr) throw std::runtime_error("build-side predicates belong in bq_join_build");
            if (p.s[s].from_build && !spec->join) throw std::runtime_error("from_build slot without a join");
        }
    }
    // Ranges reach the kernel clamped to the column's domain and never empty (its unsigned range test needs lo <= hi):
    // an always-true range is dropped, an always-false one empties the whole result without a launch.
    bool never_passes = false;
    for (int s = 0; s < N_SLOTS; ++s) never_passes = normalise_slot(p.s[s]) || never_passes;
    if (spec->n_v < 0 || spec->n_v > 2) throw std::runtime_error("at most two aggregate arguments");
    p.nv = spec->n_v;
    for (int i = 0; i < spec->n_v; ++i) {
        const bq_vexpr& e = spec->v[i];
        if (e.op < BQ_V_A || e.op > BQ_V_DIV) throw std::runtime_error("bad aggregate argument op");
        const bool binary = e.op >= BQ_V_MUL;
        if (binary && (e.l_src < BQ_L_A || e.l_src > BQ_L_IMM || e.r_src < BQ_R_B || e.r_src > BQ_R_IMM))
            throw std::runtime_error("bad aggregate argument operand source");
        bool needs_a = e.op == BQ_V_A || (binary && (e.l_src == BQ_L_A || e.r_src == BQ_R_A));
        bool needs_b = e.op == BQ_V_B || (binary && (e.l_src == BQ_L_B || e.r_src == BQ_R_B));
        if (needs_a && !spec->a.col) throw std::runtime_error("aggregate argument reads slot a, which is empty");
        if (needs_b && !spec->b.col) throw std::runtime_error("aggregate argument reads slot b, which is empty");
        p.v[i] = DVExpr{e.op, e.l_src, e.r_src, e.imm_is_f, e.imm_i, e.imm_f};
    }
    p.row_begin = spec->row_begin;
    p.row_end = spec->row_end;
    p.vec_begin = ((spec->row_begin + 3) / 4) * 4;
    if (p.vec_begin > p.row_end) p.vec_begin = p.row_end;
    p.n_chunks = (p.row_end - p.vec_begin) / 128;
    if (spec->mask) {
        if (spec->mask->type != BQ_INT64 || spec->mask->n < spec->row_end) throw std::runtime_error("mask must be an INT64 column covering the row range");
        p.mask = static_cast<const long long*>(spec->mask->ptr);
    }

    // ---- join ----
    if (spec->join) {
        const bq_join* j = spec->join;
        if (!spec->jkey.col) throw std::runtime_error("join without a probe key slot");
        p.jmode = j->kind;
        p.jk_min = j->key_min;
        p.jk_domain = static_cast<unsigned long long>(j->key_max - j->key_min) + 1ULL;
        p.j_bitmap = j->bitmap;
        p.j_direct = j->direct;
        p.jh_slots = j->h_slots;
        p.jh_occ = j->h_occ;
        p.jh_mask = j->h_mask;
        if (j->kind == BQ_JOIN_BITMAP)
            for (int s = 0; s < N_SLOTS; ++s)
                if (p.s[s].from_build) throw std::runtime_error("a bitmap join carries no build columns");
    } else if (spec->jkey.col && !spec->row_bits) {
        throw std::runtime_error("probe key slot without a join");
    }
    if (spec->row_bits) {
        if (spec->row_bits->type != BQ_STRING || spec->row_bits->n * 32 < spec->row_end) throw std::runtime_error("row bits must be a uint32 column with one bit per row");
        if (spec->row_begin % 128) throw std::runtime_error("row bits need a row range starting at a multiple of 128");
        if (spec->join) throw std::runtime_error("row bits replace the join probe: pass one or the other");
        p.row_bits = static_cast<const unsigned*>(spec->row_bits->ptr);
        p.jmode = BQ_JOIN_ROWBITS;
        const char* skip = std::getenv("BOSQL_ROWBITS_SKIP");
        p.rowbits_skip = (skip && *skip == '0') ? 0 : 1;
    }

    // ---- group state ----
    st.has_key = spec->group_mode != BQ_GROUP_NONE;
    size_t smem = 0;
    if (spec->group_mode == BQ_GROUP_NONE) {
        st.gmode = G_NONE;
        st.slots = 1;
    } else {
        if (!spec->key.col) throw std::runtime_error("GROUP BY without a key slot");
        st.key_type = spec->key.col->type;
        if (spec->group_mode == BQ_GROUP_DENSE) {
            if (st.key_type == BQ_DOUBLE) throw std::runtime_error("dense grouping needs an integer key");
            if (spec->key_max < spec->key_min) throw std::runtime_error("dense grouping needs key_min <= key_max");
            unsigned long long dom = static_cast<unsigned long long>(spec->key_max - spec->key_min) + 1ULL;
            if (dom > (1ULL << 31)) throw std::runtime_error("dense domain too large");
            st.slots = dom;
            st.key_min = spec->key_min;
            size_t need = dom * (2 * sizeof(double) + sizeof(unsigned));
            if (need <= 48 * 1024) {
                st.gmode = G_SMEM;
                smem = need;
            } else {
                st.gmode = G_DENSE;
            }
        } else {
            st.gmode = G_HASH;
            size_t hint = spec->ndv_hint ? spec->ndv_hint : (spec->row_end - spec->row_begin);
            size_t cap = next_pow2(hint * 2 < 1024 ? 1024 : hint * 2);
            if (spec->hash_part_log2 > 0) {
                if (spec->hash_part_log2 > 10) throw std::runtime_error("at most 1024 hash partitions");
                // every region gets 2x its expected share; a skewed partition raises "group table overflow"
                while ((cap >> spec->hash_part_log2) < 1024) cap <<= 1;
                p.h_part_log2 = spec->hash_part_log2;
                p.h_part_shift = spec->hash_part_shift;
            }
            st.slots = cap + 1;
            p.h_mask = cap - 1;
        }
    }
    p.key_min = st.key_min;
    p.key_domain = (st.gmode == G_SMEM || st.gmode == G_DENSE) ? st.slots : 0;

    // Staging pays when few rows survive the predicates or the table update is expensive (shared-memory CAS, hash
    // claim, multi-match probe); a join-only pipeline into a dense global table updates straight from registers,
    // and the global aggregate accumulates branch-free in registers.
    bool has_ranges = p.mask != nullptr;
    for (int s = 0; s < N_SLOTS; ++s) has_ranges = has_ranges || p.s[s].nr > 0;
    bool staged = false;
    // (a hash-join probe into a dense / global aggregate runs from registers, four first-slot loads in flight per lane)
    if (st.gmode == G_SMEM || st.gmode == G_HASH) staged = true;
    else if (st.gmode == G_DENSE) staged = has_ranges;
    p.staged = staged ? 1 : 0;
    bool specialised = false;
    uint32_t rshape = 0;
    for (int s = 0; s < N_SLOTS; ++s) rshape |= rshape_bits(s, p.s[s].nr);
    uint32_t xs = xshape(p.jmode, p.nv, p.nv > 0 ? p.v[0].op : 0, p.nv > 0 && p.v[0].op >= BQ_V_MUL ? p.v[0].l_src : 0,
                         p.nv > 0 && p.v[0].op >= BQ_V_MUL ? p.v[0].r_src : 0, p.nv > 1 ? p.v[1].op : 0,
                         p.nv > 1 && p.v[1].op >= BQ_V_MUL ? p.v[1].l_src : 0, p.nv > 1 && p.v[1].op >= BQ_V_MUL ? p.v[1].r_src : 0);
    ScanKernel fn = pick_kernel(shape, rshape, xs, p.mask != nullptr, st.gmode, staged, &specialised);
    size_t rows = spec->row_end - spec->row_begin;
    if (smem > 0) BQ_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int blocks_per_sm = 4;
    BQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, fn, kBlock, smem));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    int grid = grid_for(ctx, rows, blocks_per_sm);      // a whole number of resident CTAs per SM: one wave

    // one allocation: ticket err | sum0 | cnt | sum1 | keys | part_cnt | part_sum.  Ranks exchange a PREFIX of it: the
    // header and the arrays in use (SUM only: 8 bytes per slot; with counts 16; two sums 24)
    size_t n = st.slots;
    size_t off_tk = 0, off_s0 = 16, off_cnt = off_s0 + n * 8, off_s1 = off_cnt + n * 8, off_keys = off_s1 + n * 8;
    size_t off_pc = off_keys + (st.gmode == G_HASH ? n * 8 : 0);
    size_t off_ps = off_pc + static_cast<size_t>(grid) * 8;
    size_t total = off_ps + static_cast<size_t>(grid) * 16;
    st.ctx = ctx;
    st.block = dev_alloc(ctx, total);
    char* base = static_cast<char*>(st.block);
    st.cnt = reinterpret_cast<unsigned long long*>(base + off_cnt);
    st.sum0 = reinterpret_cast<double*>(base + off_s0);
    st.sum1 = reinterpret_cast<double*>(base + off_s1);
    st.h_keys = st.gmode == G_HASH ? reinterpret_cast<long long*>(base + off_keys) : nullptr;
    // zero everything (0 bits == 0.0 and count 0), then mark hash keys empty
    BQ_CUDA(cudaMemsetAsync(st.block, 0, total, ctx->stream));
    if (st.gmode == G_HASH) {
        k_fill_keys<<<grid_for(ctx, n, 8), kBlock, 0, ctx->stream>>>(st.h_keys, n, kEmptyKey);
        ctx->launches++;
    }
    if (spec->n_v == 0) need_count = true;
    p.need_count = need_count ? 1 : 0;

    if (!need_count && st.gmode == G_DENSE) {
        k_fill_keys<<<grid_for(ctx, n, 8), kBlock, 0, ctx->stream>>>(reinterpret_cast<long long*>(st.sum0), n, INT64_MIN);   // -0.0
        ctx->launches++;
        st.presence = 1;
    } else if (!need_count && st.gmode == G_HASH) {
        st.presence = 2;
    }
    // what another rank needs of this state: the counts travel whenever presence is read from them (shared-memory tables
    // and global aggregates always keep counts)
    st.arrays = spec->n_v > 1 ? 3 : ((need_count || st.presence == 0) ? 2 : 1);
    p.g_cnt = st.cnt;
    p.g_sum0 = st.sum0;
    p.g_sum1 = st.sum1;
    p.h_keys = st.h_keys;
    p.part_cnt = reinterpret_cast<unsigned long long*>(base + off_pc);
    p.part_sum = reinterpret_cast<double*>(base + off_ps);
    p.ticket = reinterpret_cast<unsigned*>(base + off_tk);
    p.err = reinterpret_cast<int*>(base + off_tk + 8);
    st.err = p.err;
    if (!never_passes && (rows > 0 || st.gmode == G_NONE)) {
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        if (ctx->profile) {
            BQ_CUDA(cudaEventCreate(&ev0));
            BQ_CUDA(cudaEventCreate(&ev1));
            BQ_CUDA(cudaEventRecord(ev0, ctx->stream));
        }
        fn<<<grid, kBlock, smem, ctx->stream>>>(p);
        if (ctx->profile) {
            BQ_CUDA(cudaEventRecord(ev1, ctx->stream));
            ctx->profile_events.emplace_back(ev0, ev1);
        }
        ctx->launches++;
        BQ_CUDA(cudaGetLastError());
    }
    // the error word (division by zero / table overflow / stale statistics) is read back together with the group
    // count in emit_state: one host round trip per aggregate
}

static void check_scan_flags(int flags) {
    if (flags & 1) throw std::runtime_error("Division by zero");          // src/exec/expression.cpp:52
    if (flags & 2) throw std::runtime_error("group table overflow: ndv_hint too small");
    if (flags & 4) throw std::runtime_error("group key outside [key_min,key_max]: stale statistics");
}

// Turns the device state into a relation. partial = [key] count sum0 sum1, else [key] + outs.
static bq_rel* emit_state(bq_ctx* ctx, AggState& st, const bq_agg_out* outs, int n_out, bool partial) {
    bq_agg_out pouts[3] = {{BQ_AGG_COUNT, 0, 0, 0}, {BQ_AGG_SUM, 0, 0, 0}, {BQ_AGG_SUM, 1, 0, 0}};
    if (partial) {
        outs = pouts;
        n_out = 3;
    }
    if (n_out < 0 || n_out > BQ_MAX_AGG_OUT) throw std::runtime_error("too many aggregate outputs");
    if (st.gmode != G_HASH && st.slots <= kFinishMaxSlots) {
        // the columns are allocated for every slot (at most a million rows) and cut to the group count afterwards
        std::vector<bq_col*> cols;
        try {
            EmitParams e{};
            e.gmode = st.gmode;
            e.key_type = st.key_type;
            e.key_min = st.key_min;
            e.cnt = st.cnt;
            e.sum0 = st.sum0;
            e.sum1 = st.sum1;
            if (st.has_key) {
                cols.push_back(new_col(ctx, st.key_type, st.slots));
                e.out_key = cols.back()->ptr;
            }
            e.n_out = n_out;
            for (int o = 0; o < n_out; ++o) {
                int type = BQ_DOUBLE;
                if (outs[o].func == BQ_AGG_COUNT) type = BQ_INT64;
                else if (outs[o].func == BQ_AGG_SUM) type = outs[o].as_int ? BQ_INT64 : BQ_DOUBLE;
                else if (outs[o].func != BQ_AGG_AVG) throw std::runtime_error("unknown aggregate function");
                if (outs[o].v < 0 || outs[o].v > 1) throw std::runtime_error("bad aggregate argument index");
                cols.push_back(new_col(ctx, type, st.slots));
                e.func[o] = outs[o].func;
                e.v[o] = outs[o].v;
                e.as_int[o] = outs[o].as_int;
                e.out[o] = cols.back()->ptr;
            }
            const unsigned chunks = static_cast<unsigned>((st.slots + kFinishChunk - 1) / kFinishChunk);
            auto* d = static_cast<unsigned long long*>(scratch(ctx, 16 + 64 * 4));
            auto* chunk_count = reinterpret_cast<unsigned*>(d + 2);
            if (chunks > 1) {
                k_finish_count<<<chunks, 1024, 0, ctx->stream>>>(e, st.slots, st.presence, chunk_count);
                ctx->launches++;
            }
            k_finish_small<<<chunks, 1024, 0, ctx->stream>>>(e, st.slots, st.presence, chunks > 1 ? chunk_count : nullptr, d);
            ctx->launches++;
            BQ_CUDA(cudaGetLastError());
            auto* h = static_cast<unsigned long long*>(pinned(ctx, 16));
            BQ_CUDA(cudaMemcpyAsync(h, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
            BQ_CUDA(cudaMemcpyAsync(h + 1, st.err, 4, cudaMemcpyDeviceToHost, ctx->stream));
            BQ_CUDA(cudaStreamSynchronize(ctx->stream));          // the one host round trip: group count + the scan's error word
            const int flags = static_cast<int>(h[1] & 0xFFFFFFFFull);
            if (flags) check_scan_flags(flags);
            const size_t n_groups = static_cast<size_t>(h[0]);
            for (auto* c : cols) c->n = n_groups;
            auto* rel = new bq_rel();
            rel->cols = cols;
            rel->rows = n_groups;
            return rel;
        } catch (...) {
            for (auto* c : cols) free_col(c);
            throw;
        }
    }
    bq_col* rowids = nullptr;
    size_t n_groups = 0;
    {
        size_t n_words = (st.slots + 31) / 32;
        DevBuf bits_buf(ctx, n_words * 4 + 4);
        auto* bits = bits_buf.as<unsigned>();
        size_t blocks = (st.slots + kBlock - 1) / kBlock;
        if (st.presence == 1) k_presence_sum<<<(unsigned)blocks, kBlock, 0, ctx->stream>>>(st.sum0, st.slots, bits);
        else if (st.presence == 2) k_presence_key<<<(unsigned)blocks, kBlock, 0, ctx->stream>>>(st.h_keys, st.cnt, st.slots, bits);
        else k_presence_bits<<<(unsigned)blocks, kBlock, 0, ctx->stream>>>(st.cnt, st.slots, bits);
        ctx->launches++;
        BQ_CUDA(cudaGetLastError());
        int flags = 0;
        n_groups = compact_bits(ctx, bits, st.slots, 0, &rowids, st.err, &flags);
        if (flags) {
            free_col(rowids);
            check_scan_flags(flags);
        }
    }
    std::vector<bq_col*> cols;
    try {
        EmitParams e{};
        e.rowids = static_cast<const unsigned*>(rowids->ptr);
        e.n = n_groups;
        e.gmode = st.gmode;
        e.key_type = st.key_type;
        e.key_min = st.key_min;
        e.h_keys = st.h_keys;
        e.cnt = st.cnt;
        e.sum0 = st.sum0;
        e.sum1 = st.sum1;
        if (st.has_key) {
            cols.push_back(new_col(ctx, st.key_type, n_groups));
            e.out_key = cols.back()->ptr;
        }
        e.n_out = n_out;
        for (int o = 0; o < n_out; ++o) {
            int type = BQ_DOUBLE;
            if (outs[o].func == BQ_AGG_COUNT) type = BQ_INT64;
            else if (outs[o].func == BQ_AGG_SUM) type = outs[o].as_int ? BQ_INT64 : BQ_DOUBLE;
            else if (outs[o].func != BQ_AGG_AVG) throw std::runtime_error("unknown aggregate function");
            if (outs[o].v < 0 || outs[o].v > 1) throw std::runtime_error("bad aggregate argument index");
            cols.push_back(new_col(ctx, type, n_groups));
            e.func[o] = outs[o].func;
            e.v[o] = outs[o].v;
            e.as_int[o] = outs[o].as_int;
            e.out[o] = cols.back()->ptr;
        }
        if (n_groups) {
            k_emit<<<(unsigned)((n_groups + kBlock - 1) / kBlock), kBlock, 0, ctx->stream>>>(e);
            ctx->launches++;
            BQ_CUDA(cudaGetLastError());
        }
        free_col(rowids);
        auto* rel = new bq_rel();
        rel->cols = cols;
        rel->rows = n_groups;
        return rel;
    } catch (...) {
        free_col(rowids);
        for (auto* c : cols) free_col(c);
        throw;
    }
}

// ---- merging partial states (multi-GPU) -----------------------------------------------------------
struct MergeParams {
    const void* key;
    int key_type;
    const long long* cnt;
    const double* s0;
    const double* s1;
    size_t n;
    int has_key;
    long long* h_keys;
    unsigned long long h_mask;
    unsigned long long* g_cnt;
    double* g_sum0;
    double* g_sum1;
    int* err;
};

__global__ void __launch_bounds__(kBlock) k_merge_partial(const __grid_constant__ MergeParams m) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= m.n) return;
    if (m.cnt[i] == 0) return;                 // zero padding of a fixed-capacity exchange buffer: no group
    if (m.cnt[i] < 0) {                        // a rank whose local scan failed sends -(error flags): fail here alike
        atomicOr(m.err, static_cast<int>(-m.cnt[i]));
        return;
    }
    unsigned long long idx = 0;
    if (m.has_key) {
        long long k = load_raw(m.key, m.key_type, i);
        if (m.key_type == BQ_DOUBLE && k == INT64_MIN) k = 0;
        idx = group_slot(m.h_keys, m.h_mask, m.err, k);
    }
    atomicAdd(m.g_cnt + idx, static_cast<unsigned long long>(m.cnt[i]));
    atomicAdd(m.g_sum0 + idx, m.s0[i]);
    atomicAdd(m.g_sum1 + idx, m.s1[i]);
}

// Dense states of all ranks, gathered back to back (block r = rank r's cnt | sum0 | sum1 | ticket err), folded into this
// rank's state: counts add up, sums are added in rank order (the same bits on every rank), error flags are OR-ed.
__global__ void __launch_bounds__(kBlock) k_fold_dense(const unsigned char* __restrict__ gathered, size_t block_bytes, int world, size_t n,
                                                       int arrays, unsigned long long* __restrict__ cnt, double* __restrict__ sum0,
                                                       double* __restrict__ sum1, int* __restrict__ err) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) {
        unsigned long long c = 0;
        // -0.0 is the identity of IEEE addition AND the "slot never touched" mark of a state kept without counts
        // (run_scan): (-0.0) + (-0.0) = -0.0, (-0.0) + x = x, so presence survives the fold
        double s0 = -0.0, s1 = -0.0;
        for (int r = 0; r < world; ++r) {
            const unsigned char* b = gathered + static_cast<size_t>(r) * block_bytes + 16;
            s0 = __dadd_rn(s0, reinterpret_cast<const double*>(b)[i]);
            if (arrays > 1) c += reinterpret_cast<const unsigned long long*>(b + n * 8)[i];
            if (arrays > 2) s1 = __dadd_rn(s1, reinterpret_cast<const double*>(b + n * 16)[i]);
        }
        sum0[i] = s0;
        if (arrays > 1) cnt[i] = c;
        if (arrays > 2) sum1[i] = s1;
    }
    if (i == 0) {
        int e = 0;
        for (int r = 0; r < world; ++r) e |= *reinterpret_cast<const int*>(gathered + static_cast<size_t>(r) * block_bytes + 8);
        *err = e;
    }
}

}  // namespace bq

using namespace bq;

struct bq_agg_state {
    AggState st;
};

extern "C" {

int bq_scan_state(bq_ctx* ctx, const bq_scan_spec* spec, bq_agg_state** out) {
    return guarded([&] {
        auto* s = new bq_agg_state();
        try {
            // counts are kept only when an output needs them (COUNT, AVG): SUM-only states mark presence on the sums, which
            // saves one L2 reduction per qualifying row (Q2 across GPUs: 3.9 -> 3.1 ms per 500 M probe rows)
            bool need_count = spec->n_out == 0;
            for (int o = 0; o < spec->n_out; ++o) need_count = need_count || spec->out[o].func != BQ_AGG_SUM;
            run_scan(ctx, spec, s->st, need_count);
        } catch (...) {
            delete s;
            throw;
        }
        *out = s;
    });
}

int bq_agg_state_dense(const bq_agg_state* s, void** ptr, size_t* bytes) {
    const bool dense = s->st.gmode != G_HASH;
    if (ptr) *ptr = dense ? s->st.block : nullptr;
    if (bytes) *bytes = dense ? 16 + s->st.slots * 8 * static_cast<size_t>(s->st.arrays) : 0;
    return 0;
}

int bq_agg_state_fold(bq_ctx* ctx, bq_agg_state* s, const void* gathered, int world) {
    return guarded([&] {
        if (s->st.gmode == G_HASH) throw std::runtime_error("only dense aggregate states are exchanged as they are");
        const size_t n = s->st.slots;
        k_fold_dense<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, ctx->stream>>>(static_cast<const unsigned char*>(gathered), 16 + n * 8 * static_cast<size_t>(s->st.arrays),
                                                                                   world, n, s->st.arrays, s->st.cnt, s->st.sum0, s->st.sum1, s->st.err);
        ctx->launches++;
        BQ_CUDA(cudaGetLastError());
        // s->st.presence is unchanged: every rank ran the same plan, so all blocks carry counts or none does
    });
}

int bq_agg_state_emit(bq_ctx* ctx, bq_agg_state* s, const bq_agg_out* outs, int n_out, bq_rel** out) {
    return guarded([&] { *out = emit_state(ctx, s->st, outs, n_out, false); });
}

void bq_agg_state_free(bq_agg_state* s) { delete s; }

int bq_scan_aggregate(bq_ctx* ctx, const bq_scan_spec* spec, bq_rel** out) {
    return guarded([&] {
        AggState st;
        bool need_count = false;
        for (int o = 0; o < spec->n_out; ++o) need_count = need_count || spec->out[o].func != BQ_AGG_SUM;
        run_scan(ctx, spec, st, need_count);
        *out = emit_state(ctx, st, spec->out, spec->n_out, false);
    });
}

int bq_scan_partial(bq_ctx* ctx, const bq_scan_spec* spec, bq_rel** out) {
    return guarded([&] {
        AggState st;
        run_scan(ctx, spec, st, true);
        *out = emit_state(ctx, st, nullptr, 0, true);
    });
}

int bq_agg_finish(bq_ctx* ctx, const bq_rel* const* parts, int n_parts, int has_key, int key_type,
                  const bq_agg_out* outs, int n_out, bq_rel** out) {
    return guarded([&] {
        size_t total = 0;
        const int ncols = has_key ? 4 : 3;
        for (int i = 0; i < n_parts; ++i) {
            if (static_cast<int>(parts[i]->cols.size()) != ncols) throw std::runtime_error("partial relation has the wrong arity");
            total += parts[i]->rows;
        }
        AggState st;
        st.has_key = has_key != 0;
        st.key_type = key_type;
        st.gmode = has_key ? G_HASH : G_NONE;
        size_t cap = next_pow2(total * 2 < 1024 ? 1024 : total * 2);
        st.slots = has_key ? cap + 1 : 1;
        size_t n = st.slots;
        size_t bytes = n * 8 * 4 + 16;
        st.ctx = ctx;
        st.block = dev_alloc(ctx, bytes);
        char* base = static_cast<char*>(st.block);
        st.cnt = reinterpret_cast<unsigned long long*>(base);
        st.sum0 = reinterpret_cast<double*>(base + n * 8);
        st.sum1 = reinterpret_cast<double*>(base + n * 16);
        st.h_keys = has_key ? reinterpret_cast<long long*>(base + n * 24) : nullptr;
        int* err = reinterpret_cast<int*>(base + n * 32);
        st.err = err;
        BQ_CUDA(cudaMemsetAsync(st.block, 0, bytes, ctx->stream));
        if (has_key) {
            k_fill_keys<<<grid_for(ctx, n, 8), kBlock, 0, ctx->stream>>>(st.h_keys, n, kEmptyKey);
            ctx->launches++;
        }
        // parts are folded one launch after another, i.e. in index order per group
        for (int i = 0; i < n_parts; ++i) {
            const bq_rel* r = parts[i];
            if (!r->rows) continue;
            MergeParams m{};
            int c = 0;
            if (has_key) {
                m.key = r->cols[c]->ptr;
                m.key_type = r->cols[c]->type;
                ++c;
            }
            m.cnt = static_cast<const long long*>(r->cols[c]->ptr);
            m.s0 = static_cast<const double*>(r->cols[c + 1]->ptr);
            m.s1 = static_cast<const double*>(r->cols[c + 2]->ptr);
            m.n = r->rows;
            m.has_key = has_key;
            m.h_keys = st.h_keys;
            m.h_mask = cap - 1;
            m.g_cnt = st.cnt;
            m.g_sum0 = st.sum0;
            m.g_sum1 = st.sum1;
            m.err = err;
            k_merge_partial<<<(unsigned)((r->rows + kBlock - 1) / kBlock), kBlock, 0, ctx->stream>>>(m);
            ctx->launches++;
            BQ_CUDA(cudaGetLastError());
        }
        *out = emit_state(ctx, st, outs, n_out, false);
    });
}

}  // extern "C"
