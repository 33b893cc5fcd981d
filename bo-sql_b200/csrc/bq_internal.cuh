// bq_internal.cuh — functions shared between the translation units of libbosql_b200.so.
#pragma once
#include "bq_common.cuh"

namespace bq {

// bq_compact.cu
// dev_flag (optional): a device int fetched in the same host round trip as the count
size_t compact_bits(bq_ctx* ctx, const unsigned* bits, size_t n_bits, unsigned base_index, bq_col** out_rowids,
                    const int* dev_flag = nullptr, int* host_flag = nullptr);
// want_total = false: nothing is read back and the host does not wait (the caller does not need the grand total)
size_t exclusive_scan_u32(bq_ctx* ctx, const unsigned* counts, size_t n, unsigned long long* offsets, bool want_total = true);

// ---- device view of a predicate slot (bq_slot resolved to raw pointers) ------------------------
struct DSlot {
    const void* ptr;
    int kind;
    int nr;
    long long lo0, hi0, lo1, hi1;
    int neg0, neg1;
    int from_build;
    int pad;
};
DSlot make_dslot(const bq_slot& s, size_t need_rows, const char* what);
bool normalise_slot(DSlot& d);     // clamp / drop always-true ranges; true = some range can never pass

#if defined(__CUDACC__)
// one range: (lo <= k && k <= hi) != neg
BQ_D bool in_range(long long k, long long lo, long long hi, int neg) { return ((k >= lo) & (k <= hi)) != (neg != 0); }
BQ_D bool slot_pass(const DSlot& s, long long raw) {
    if (s.nr == 0) return true;
    long long k = key_of(raw, s.kind);
    bool ok = in_range(k, s.lo0, s.hi0, s.neg0);
    if (s.nr > 1) ok = ok && in_range(k, s.lo1, s.hi1, s.neg1);
    return ok;
}
#endif

}  // namespace bq
