// bq_common.cuh — shared device/host helpers for the sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <array>
#include <vector>

#include "bosql_b200.h"

#if defined(__CUDACC__)
#define BQ_HD __host__ __device__ __forceinline__
#define BQ_D __device__ __forceinline__
#else
#define BQ_HD inline
#define BQ_D inline
#endif

#define BQ_CUDA(expr)                                                                             \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + \
                                     __FILE__ + ":" + std::to_string(__LINE__) + " (" #expr ")");   \
        }                                                                                         \
    } while (0)

namespace bq {

constexpr int kBlock = 256;          // threads per CTA of the streaming kernels
constexpr int kRowsPerThread = 4;    // one 128-bit load of a 4-byte column, two of an 8-byte column
constexpr int kTileRows = kBlock * kRowsPerThread;

// ---- physical kinds = type ordinals of include/bosql_b200.h -------------------------------------
BQ_HD int width_of(int type) { return (type == BQ_INT64 || type == BQ_DOUBLE) ? 8 : 4; }

// Order-preserving int64 key of a double: IEEE '<' on non-NaN values == '<' on keys.
// -0.0 is folded onto +0.0 first (IEEE compares them equal); NaNs land outside [key(-inf), key(+inf)].
BQ_HD int64_t f64_key_from_bits(uint64_t b) {
    if (b == 0x8000000000000000ULL) b = 0;
    int64_t s = static_cast<int64_t>(b);
    return s ^ ((s >> 63) & 0x7FFFFFFFFFFFFFFFLL);
}
BQ_HD uint64_t f64_bits_from_key(int64_t k) {
    return static_cast<uint64_t>(k ^ ((k >> 63) & 0x7FFFFFFFFFFFFFFFLL));
}

// ---- counter-based hash (splitmix64 finaliser): value(row) = f(seed, stream, row) ---------------
BQ_HD uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
BQ_HD uint64_t row_hash(uint64_t seed, uint64_t stream, uint64_t row) {
    return mix64(mix64(seed ^ (stream * 0xD6E8FEB86659FD93ULL)) + row);
}

// hash used by the open-addressing tables (the reference's hash is not observable)
BQ_HD uint64_t key_hash(uint64_t k) {
    k ^= k >> 33;
    k *= 0xFF51AFD7ED558CCDULL;
    k ^= k >> 33;
    k *= 0xC4CEB9FE1A85EC53ULL;
    k ^= k >> 33;
    return k;
}

#if defined(__CUDACC__)
// ---- streaming loads: 128-bit, read-only path, no L1 allocation (each byte is read once) --------
BQ_D int4 ldg_stream(const int4* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// L2 residency hints (createpolicy + ld ... L2::cache_hint).  A kernel that streams tens of GB past a table it probes at
// random (a join bitmap, a direct-address table) marks the stream evict-first and the table evict-last, so the table is
// what the 126 MB L2 keeps: without the hints the stream keeps pushing table lines out.
BQ_D uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
BQ_D uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
BQ_D int4 ldg_stream_hint(const int4* p, uint64_t policy) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(policy));
    return r;
}
BQ_D int2 ldg_stream2_hint(const int2* p, uint64_t policy) {
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.s32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(policy));
    return r;
}
BQ_D long long ldg_stream_i64_hint(const long long* p, uint64_t policy) {
    long long r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(policy));
    return r;
}
BQ_D int ldg_stream_i32_hint(const int* p, uint64_t policy) {
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(policy));
    return r;
}
BQ_D unsigned ldg_keep_u32(const unsigned* p, uint64_t policy) {
    unsigned r;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(policy));
    return r;
}

// Widen one element of `kind` at row i to its 8-byte slot value (what BQ_OP_COL pushes).
BQ_D int64_t load_raw(const void* base, int kind, size_t i) {
    switch (kind) {
        case BQ_INT64:
        case BQ_DOUBLE: return __ldg(reinterpret_cast<const long long*>(base) + i);
        case BQ_STRING: return static_cast<int64_t>(__ldg(reinterpret_cast<const unsigned*>(base) + i));
        default: return static_cast<int64_t>(__ldg(reinterpret_cast<const int*>(base) + i));
    }
}
// The integer key on which ranges are tested.
BQ_D int64_t key_of(int64_t raw, int kind) {
    return kind == BQ_DOUBLE ? f64_key_from_bits(static_cast<uint64_t>(raw)) : raw;
}
// datum_as_double (src/exec/operator.cpp:280-292)
BQ_D double as_double(int64_t raw, int kind) {
    return kind == BQ_DOUBLE ? __longlong_as_double(raw) : static_cast<double>(raw);
}

BQ_D double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
BQ_D unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

// ---- host-side objects behind the opaque handles ------------------------------------------------
}  // namespace bq

struct bq_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    // small reusable device scratch (partials, tickets, error flags)
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    void* pinned = nullptr;      // small pinned staging for scalar results
    size_t pinned_bytes = 0;
    // Large blocks (>= 32 MB) bypass the driver's stream-ordered pool: when a query's multi-GB temporaries change size from
    // one step to the next (exchange buffers), that pool re-maps physical memory under new addresses at ~40 ms per GB.
    // Blocks here are recycled whole on the context's stream (same-stream ordering makes reuse safe).
    struct BigBlock { void* p; size_t bytes; bool free; uint64_t stamp; bool exported = false; };
    // peer allocations mapped through CUDA IPC, keyed by the 64-byte handle (recycled blocks keep their handle)
    std::vector<std::pair<std::array<unsigned char, 64>, void*>> ipc_mappings;
    std::vector<BigBlock> big_blocks;
    size_t big_free_bytes = 0;
    uint64_t big_clock = 0;
    void* comm = nullptr;        // bq::Comm (bq_comm.cu): this rank's NCCL communicator, when the process is one of several
    bool profile = false;        // bracket the fused scan kernel with events (bench.py roofline)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> profile_events;
};

struct bq_col {
    bq_ctx* ctx = nullptr;       // owner (frees go back to its stream-ordered pool while it is alive)
    int type = BQ_INT64;
    size_t n = 0;
    void* ptr = nullptr;
    bool owns = true;
    bool has_minmax = false;
    int64_t min_key = 0, max_key = 0;
    size_t ndv = 0;
};

struct bq_rel {
    std::vector<bq_col*> cols;
    size_t rows = 0;
};

namespace bq {
// One slot of the open-addressing join table: key and build row id + 1 (0 = empty) side by side, so a probe step is ONE
// 16-byte load (the first-generation layout kept keys and rows in two arrays: two random sectors per step).
struct alignas(16) JoinSlot {
    long long key;
    unsigned row;
    unsigned pad;
};
#if defined(__CUDACC__)
BQ_D JoinSlot load_join_slot(const JoinSlot* p) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(p));
    JoinSlot s;
    s.key = static_cast<long long>((static_cast<unsigned long long>(static_cast<unsigned>(v.y)) << 32) | static_cast<unsigned>(v.x));
    s.row = static_cast<unsigned>(v.z);
    s.pad = 0;
    return s;
}
#endif
}  // namespace bq

struct bq_join {
    bq_ctx* ctx = nullptr;
    int kind = BQ_JOIN_HASH;
    int64_t key_min = 0, key_max = 0;     // BITMAP / DIRECT domain
    unsigned* bitmap = nullptr;           // BITMAP: bit (key-key_min)
    size_t bitmap_words = 0;
    unsigned* direct = nullptr;           // DIRECT: build row id + 1 at [key-key_min], 0 = absent
    bq::JoinSlot* h_slots = nullptr;      // HASH: open addressing, linear probing over 16-byte slots
    unsigned* h_occ = nullptr;            // HASH: one bit per slot, set when the slot holds a key (tables of <= 2^29 slots: the bits
                                          // fit the L2, so a probe whose home slot is empty never touches the table in HBM)
    uint64_t h_mask = 0;
    size_t build_rows = 0;                // rows inserted
    uint64_t bitmap_bits = 0;             // BITMAP: bits set when the build finished (== build_rows unless a key repeats)
    size_t bytes = 0;
};

namespace bq {

void set_error(const std::string& msg);
bq_col* new_col(bq_ctx* ctx, int type, size_t n);
void free_col(bq_col* c);
// Stream-ordered device memory (cudaMallocAsync on the context's stream, pool never trimmed): allocation and
// release cost microseconds and never synchronise the device, unlike cudaMalloc/cudaFree.
void* dev_alloc(bq_ctx* ctx, size_t bytes);
void dev_free(bq_ctx* ctx, void* p);          // ctx may be dead or null: falls back to cudaFree
struct DevBuf {
    bq_ctx* ctx;
    void* p = nullptr;
    DevBuf(bq_ctx* c, size_t bytes) : ctx(c), p(dev_alloc(c, bytes)) {}
    ~DevBuf() { dev_free(ctx, p); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    template <typename T> T* as() const { return static_cast<T*>(p); }
};
void* scratch(bq_ctx* ctx, size_t bytes);     // device scratch, valid until the next call that asks for more
void* pinned(bq_ctx* ctx, size_t bytes);
int grid_for(bq_ctx* ctx, size_t rows, int blocks_per_sm);

template <typename F>
int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        set_error(e.what());
        return 1;
    }
}

}  // namespace bq
