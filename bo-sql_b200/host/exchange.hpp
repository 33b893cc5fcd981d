// exchange.hpp — the multi-GPU exchange points of a plan.  The reference is a single process; this is the
// part of the operator layer that has no counterpart there.  The collectives themselves are supplied by the host through
// bqx_set_exchange (include/bosql_b200_exec.h); this file packs, sizes and unpacks what they move.
#pragma once

#include <cstdint>
#include <functional>
#include <vector>

#include "bosql_b200_exec.h"
#include "gpu_device.hpp"

namespace bosql::gpu {

struct Exchange {
    bqx_exchange fn{};
    bool active = false;
    bool native = false;      // fn points at the library's own NCCL collectives (bqx_comm_init), not at host callbacks
    int world() const { return fn.world; }
    int rank() const { return fn.rank; }

    // all[r*n + i] = rank r's mine[i]
    std::vector<int64_t> host_gather(const std::vector<int64_t>& mine);
    int64_t host_sum(int64_t v);
    bool host_all(bool v);                                   // logical AND over ranks
    void minmax(int64_t& lo, int64_t& hi);                   // global [lo, hi]; ranks with lo > hi hold no rows
    void sum_words(void* device_words, size_t n_words);

    // The concatenation, in rank order, of every rank's first `rows` rows of `col` (replicated on every rank).
    DevColPtr all_gather_column(const DevColPtr& col, size_t rows, const std::vector<int64_t>& rows_by_rank);
};
Exchange& exchange();

// A step that can fail on this rank's rows alone (an expression program hitting an integer division by zero in its shard):
// across GPUs every rank learns the outcome before anyone moves on, so that all of them throw and none is left waiting
// inside the next collective.  Single-GPU: just runs the step.  Every rank must call this at the same point of the plan.
void agree_on(const std::function<void()>& step);

// Partial aggregate states of every rank, ready for bq_agg_finish.  `local` is this rank's [key] count sum0 sum1 relation
// (nullptr with error_flags != 0 when the local scan failed: the flags travel as a negative count so that every rank's
// merge reports the same error).  capacity >= 0: every rank has at most that many groups (one fixed-size all-gather, no
// size exchange); capacity < 0: sizes are exchanged first.
struct GatheredPartials {
    DevColPtr buffer;
    std::vector<bq_rel*> parts;
    GatheredPartials() = default;
    GatheredPartials(const GatheredPartials&) = delete;
    GatheredPartials& operator=(const GatheredPartials&) = delete;
    ~GatheredPartials();
};
void gather_partials(const DeviceRelation* local, int error_flags, bool has_key, TypeId key_type, int64_t capacity,
                     GatheredPartials& out);

// Dense / global aggregate states (bq_scan_state): every rank's raw state block is all-gathered (same size everywhere: the
// key domain comes from catalog statistics) and folded in rank order by one launch - no compaction, no size exchange.
void all_gather_fold_state(bq_agg_state* state);

// Hash-partitions rows [0, rows) of key + payload by rank and exchanges them all-to-all; returns the rows this rank owns.
struct Shuffled {
    DevColPtr key;
    std::vector<DevColPtr> payload;
    size_t rows = 0;
};
// hot_keys (<= 16, the same list on every rank): rows with one of these keys are NOT moved by hash.  keep_hot_local: they
// stay on the rank that holds them (the probe side of a skewed join); otherwise every rank receives all of them (the
// matching build rows, replicated), so the join still sees every pair exactly once.
Shuffled shuffle_by_key(const DevColPtr& key, const std::vector<DevColPtr>& payload, size_t rows,
                        const std::vector<int64_t>& hot_keys = {}, bool keep_hot_local = true);

// $BOSQL_TRACE=1: synchronise and print the time since the previous mark (stderr) - phase breakdowns for tuning.
struct PhaseTrace {
    PhaseTrace();
    void mark(const char* what);
    bool on = false;
    double last = 0.0;
};

// Concatenates every rank's relation (same schema) in rank order.
DeviceRelationPtr all_gather_relation(const DeviceRelationPtr& local, const std::vector<TypeId>& types);

}  // namespace bosql::gpu
