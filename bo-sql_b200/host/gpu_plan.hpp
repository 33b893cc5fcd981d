// gpu_plan.hpp — turning a Pipeline + aggregate/selection request into kernel-layer specs.
#pragma once

#include <map>
#include <string>
#include <vector>

#include "expr_compile.hpp"
#include "gpu_device.hpp"

namespace bosql::gpu {

ColumnLookup lookup_for(const std::vector<PipeCol>& cols);
ColumnLookup lookup_for(const std::vector<std::string>& names, const std::vector<TypeId>& types);

DeviceRelationPtr empty_relation(const std::vector<TypeId>& types);
// the rows of `parts`, in order, as one relation (device-to-device copies)
DeviceRelationPtr concat_relations(const std::vector<DeviceRelationPtr>& parts, const std::vector<TypeId>& types);

// min/max of the column's key from the catalog, else computed on the device (cached in the handle)
void resolve_stats(PipeCol& col, bool force_device = false);

// Evaluates `e` over the first `rows` rows of `cols` into a fresh device column of `out_type`.
DevColPtr eval_to_column(const Expr* e, const std::vector<PipeCol>& cols, size_t rows, Dictionary* dict,
                         TypeId out_type, bool as_predicate);

// Result of distributing predicate conjuncts over kernel slots.
struct SlotPlan {
    std::vector<std::vector<bq_range>> role_ranges;               // parallel to the caller's role columns
    std::vector<std::pair<int, std::vector<bq_range>>> pred;      // (column, ranges) for predicate-only slots
    DevColPtr mask;                                               // AND of everything no range expresses
};
// conjuncts: predicates over `cols`; role_cols: columns already bound to a slot (index into cols, -1 = unused).
SlotPlan plan_slots(const std::vector<const Conjunct*>& conjuncts, const std::vector<PipeCol>& cols, size_t rows,
                    const std::vector<int>& role_cols, int n_pred_slots);

bq_slot make_slot(const DevColPtr& col, const std::vector<bq_range>& ranges, bool from_build = false);

// What every rank's shard holds (rows scanned, build rows), as exchanged by an earlier run of the same plan.
struct RowCountCache {
    bool have = false;
    int64_t local_probe = -1, local_build = -1;
    std::vector<int64_t> all;       // [rank * 2] probe rows, [rank * 2 + 1] build rows
};

struct AggRequest {
    const std::vector<std::unique_ptr<Expr>>* group_exprs = nullptr;
    struct Agg {
        std::string func;      // COUNT / SUM / AVG
        const Expr* arg = nullptr;
        TypeId result_type = TypeId::INT64;
    };
    std::vector<Agg> aggs;
    std::vector<TypeId> group_types;
    Dictionary* dict = nullptr;
    RowCountCache* row_cache = nullptr;      // optional, owned by the operator that runs the plan repeatedly
};
// Runs the fused scan -> selection -> [join probe] -> aggregate pipeline.  Returns nullptr when the pipeline
// cannot be fused as described (the caller materialises the child and calls again on the plain relation).
DeviceRelationPtr run_aggregate(Pipeline& p, const AggRequest& req);

// Selection in two halves: the ascending row ids (uint32) of the rows passing `conjuncts`, and the gather of those rows.
DevColPtr select_rowids(const std::vector<PipeCol>& cols, size_t rows, const std::vector<const Conjunct*>& conjuncts);
DeviceRelationPtr gather_rows(const std::vector<PipeCol>& cols, const DevColPtr& rowids);
// Rows of `rel` passing `conjuncts`, in order (Selection).
DeviceRelationPtr run_selection(const std::vector<PipeCol>& cols, size_t rows, const std::vector<const Conjunct*>& conjuncts);

}  // namespace bosql::gpu
