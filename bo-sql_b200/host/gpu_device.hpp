// gpu_device.hpp — host-side handles over the kernel-layer C ABI (include/bosql_b200.h): the process-wide
// context, RAII device columns/relations, and the Pipeline description that fusable operators exchange.
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "bosql_b200.h"
#include "bosql_sql.hpp"
#include "bosql_types.hpp"

namespace bosql::gpu {

// One context per process (one process per GPU).  Device ordinal: bqx_init(), else $BOSQL_DEVICE, else
// $LOCAL_RANK, else 0.  Throws when no CUDA device is present: there is no CPU execution path.
bq_ctx* context();
void init_context(int device);
void shutdown_context();
[[noreturn]] void throw_last_error();
inline void check(int rc) { if (rc) throw_last_error(); }

// Shared ownership of a device column.
struct DevCol {
    bq_col* h = nullptr;
    bool owns = true;
    DevCol(bq_col* handle, bool own) : h(handle), owns(own) {}
    ~DevCol();
    DevCol(const DevCol&) = delete;
    DevCol& operator=(const DevCol&) = delete;
    size_t rows() const { return bq_col_size(h); }
    TypeId type() const { return static_cast<TypeId>(bq_col_type(h)); }
    std::shared_ptr<DevCol> parent;                // a view keeps the column it points into alive
};
using DevColPtr = std::shared_ptr<DevCol>;
DevColPtr adopt(bq_col* h);                       // takes ownership
// rows [begin, end) of `col` as a column of its own, without copying (ColumnarScan's zero-copy slices, on the device)
DevColPtr view_of(const DevColPtr& col, size_t begin, size_t end);

// The HBM mirror of a table column: uploaded on first use, cached in Column::device.
DevColPtr mirror_of(const Column& col);

struct DeviceRelation {
    std::vector<DevColPtr> cols;
    size_t rows = 0;
    // Across GPUs: false = these are the rows THIS rank produced from its shard (scan, selection, join, a GROUP BY whose
    // groups stay with their owner); true = every rank holds the same complete relation (merged aggregates).
    bool replicated = false;
    // The rows are in ascending order of column 0 and column 0 holds distinct values (the groups of a dense aggregate
    // state are emitted in key order): an ORDER BY on that column alone has nothing left to do.
    bool ordered_by_first = false;
};
using DeviceRelationPtr = std::shared_ptr<DeviceRelation>;
DeviceRelationPtr relation_from(bq_rel* rel);     // consumes the bq_rel shell, keeps its columns

// Host memory for result columns.  Small results are plain heap memory; large ones (>= 256 KB) are pinned buffers from a
// recycling pool, so the device-to-host copy of a large result runs at PCIe speed and repeated queries pay no allocation.
std::shared_ptr<void> host_buffer(size_t bytes);

// Catalog statistics attached to a pipeline column (include/catalog/catalog.h:16-21).
struct KeyStats {
    bool known = false;
    bool measured = false;                // computed on this process's rows (a shard, when running across GPUs)
    int64_t min_key = 0, max_key = -1;    // on the integer key (f64 key for DOUBLE)
    size_t ndv = 0;
    size_t table_rows = 0;
};

struct PipeCol {
    std::string name;
    TypeId type;
    DevColPtr dev;
    int side = 0;              // 0 = probe / base table, 1 = build side of the join
    KeyStats stats;
};

struct Conjunct {
    std::unique_ptr<Expr> expr;
    Dictionary* dict = nullptr;    // dictionary of the Selection that owns it (string literals resolve there)
};

// scan [-> selection]* [-> inner join with a scan [-> selection]* build side] [-> selection]*
struct Pipeline {
    std::vector<PipeCol> cols;             // output schema order: probe columns, then build columns
    size_t rows = 0;                       // probe/base rows
    std::vector<Conjunct> conjuncts;       // pending predicates over `cols`
    bool joined = false;
    int probe_key = -1, build_key = -1;    // indices into cols
    size_t build_rows = 0;
    bool cross_join = false;               // ON was not col = col: every row matches every row
    Dictionary* dict = nullptr;
};

// Splits a predicate into its top-level AND conjuncts (clones).
void split_conjuncts(const Expr* e, Dictionary* dict, std::vector<Conjunct>& out);

}  // namespace bosql::gpu
