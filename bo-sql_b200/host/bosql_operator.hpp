// bosql_operator.hpp — the reference's physical-operator interface (include/exec/operator.hpp:17-218),
// implemented on the GPU.
//
// Same class names, constructor signatures, open/next/close contract, output_names()/output_types()/
// dictionary() and error behaviour (std::runtime_error with the reference's messages), so that
// build_physical_plan (src/exec/physical_planner.cpp:9-124) and run_query (src/exec/execution.cpp:8-61)
// work unchanged above it.  Below it nothing is the same: operators are plan nodes; a blocking operator
// (aggregate, sort, join build) recognises the pipeline beneath it and runs it as ONE fused kernel over
// device-resident columns through the C ABI of include/bosql_b200.h; next() then pages the finished
// result out in <= 4096-row host batches.
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "bosql_sql.hpp"
#include "bosql_types.hpp"

namespace bosql {

namespace gpu {
struct DeviceRelation;     // columns in HBM + row count (RAII over bq_col handles)
struct Pipeline;           // scan [-> selection]* [-> join] description handed to a fused kernel
}

struct ExprBindings {      // reference: include/exec/expression.h:13-22
    const std::vector<std::string>* column_names = nullptr;
    const std::vector<TypeId>* column_types = nullptr;
    std::unordered_map<std::string, size_t> name_to_index;
    Dictionary* dictionary = nullptr;
};
ExprBindings make_bindings(const std::vector<std::string>& names, const std::vector<TypeId>& types,
                           Dictionary* dictionary = nullptr);

struct Operator {
    virtual ~Operator() = default;
    virtual void open() = 0;
    virtual bool next(ExecBatch& out) = 0;
    virtual void close() = 0;

    const std::vector<std::string>& output_names() const { return names_; }
    const std::vector<TypeId>& output_types() const { return types_; }
    Dictionary* dictionary() const { return dict_; }

    // ---- device-side protocol (not part of the reference interface) --------------------------------
    // The operator's whole output as a device relation (computed on first call after open()).
    virtual std::shared_ptr<gpu::DeviceRelation> device_result() = 0;
    // What a consumer that stops pulling after `want` (> 0) rows makes this operator compute (Limit::next stops calling
    // next() once it has its rows, src/exec/operator.cpp:577-613): a relation whose first min(want, total) rows are the
    // first rows of device_result(), for which no expression was evaluated on a row the reference would not have reached.
    // Blocking operators compute everything either way (the default).
    virtual std::shared_ptr<gpu::DeviceRelation> device_prefix(size_t want) { (void)want; return device_result(); }
    // Describe this subtree as a fusable pipeline; false = not a scan/selection/join chain.
    virtual bool describe(gpu::Pipeline&) { return false; }
    // After the first next(): the whole result as host columns, when this operator pages out of one materialised copy
    // (lets a caller that wants everything skip the 4096-row batches; the batches alias the same memory).
    bool host_result(std::vector<std::shared_ptr<void>>& cols, size_t& rows) const;

protected:
    std::vector<std::string> names_;
    std::vector<TypeId> types_;
    Dictionary* dict_ = nullptr;

    // result paging shared by every operator: whole columns are brought to the host once, then sliced
    bool page_out(ExecBatch& out);
    void reset_paging();
    std::shared_ptr<gpu::DeviceRelation> result_;
    std::vector<std::shared_ptr<void>> host_cols_;
    size_t emit_offset_ = 0;
    bool paged_ = false;
};

// Every operator overrides the reference's open / next / close and the device-side device_result().
#define BQ_OPERATOR_LIFECYCLE                   \
    void open() override;                       \
    bool next(ExecBatch& out) override;         \
    void close() override;                      \
    std::shared_ptr<gpu::DeviceRelation> device_result() override;

struct ColumnarScan : public Operator {
    ColumnarScan(Table* t, std::vector<size_t> idx, size_t batch = 4096);
    BQ_OPERATOR_LIFECYCLE
    bool describe(gpu::Pipeline&) override;
    std::shared_ptr<gpu::DeviceRelation> device_prefix(size_t want) override;
    // rows [begin, end) of the scan's output, zero-copy (the batches [begin, end) covers, src/exec/operator.cpp:345-384)
    std::shared_ptr<gpu::DeviceRelation> device_window(size_t begin, size_t end);
    size_t table_rows() const;
    size_t batch_rows() const { return batch_size; }
    // Catalog statistics for this table (the planner attaches them; the reference's scan sees only Table*).
    void set_table_meta(const TableMeta* meta) { meta_ = meta; }
private:
    Table* table;
    std::vector<size_t> indices;
    size_t offset;
    size_t batch_size;
    const TableMeta* meta_ = nullptr;
};

struct Selection : public Operator {
    Selection(std::unique_ptr<Operator> c, std::unique_ptr<Expr> pred);
    BQ_OPERATOR_LIFECYCLE
    bool describe(gpu::Pipeline&) override;
    std::shared_ptr<gpu::DeviceRelation> device_prefix(size_t want) override;
private:
    std::shared_ptr<gpu::DeviceRelation> select_from(const std::shared_ptr<gpu::DeviceRelation>& in);
    std::unique_ptr<Operator> input_;
    std::unique_ptr<Expr> predicate;
    ExprBindings bindings;
};

struct Project : public Operator {
    Project(std::unique_ptr<Operator> c, std::vector<std::unique_ptr<Expr>> exprs, std::vector<std::string> aliases);
    BQ_OPERATOR_LIFECYCLE
    std::shared_ptr<gpu::DeviceRelation> device_prefix(size_t want) override;
private:
    std::shared_ptr<gpu::DeviceRelation> project_from(const std::shared_ptr<gpu::DeviceRelation>& in);
    std::unique_ptr<Operator> input_;
    std::vector<std::unique_ptr<Expr>> expressions;
    std::vector<std::string> aliases;
    ExprBindings bindings;
    std::vector<std::string> input_names;
    std::vector<TypeId> input_types;
    std::vector<int> direct_indices;
};

struct HashJoin : public Operator {
    HashJoin(std::unique_ptr<Operator> left, std::unique_ptr<Operator> right, std::vector<std::string> left_keys,
             std::vector<std::string> right_keys, std::unique_ptr<Expr> residual);
    BQ_OPERATOR_LIFECYCLE
    bool describe(gpu::Pipeline&) override;
private:
    std::unique_ptr<Operator> left_child, right_child;
    std::vector<std::string> left_key_names, right_key_names;
    std::unique_ptr<Expr> residual_filter;       // stored and never evaluated, like the reference
    std::vector<size_t> left_key_indices, right_key_indices;
    std::vector<TypeId> left_key_types, right_key_types;
    std::vector<std::string> left_names, right_names;
    std::vector<TypeId> left_types, right_types;
};

struct AggregateSpec {
    std::string func_name;
    std::unique_ptr<Expr> arg;
    std::string alias;
};

struct HashAggregate : public Operator {
    HashAggregate(std::unique_ptr<Operator> input, std::vector<std::unique_ptr<Expr>> group_exprs,
                  std::vector<AggregateSpec> aggregates);
    BQ_OPERATOR_LIFECYCLE
private:
    std::unique_ptr<Operator> input_;
    std::vector<std::unique_ptr<Expr>> group_exprs;
    std::vector<AggregateSpec> aggregates;
    ExprBindings child_bindings;
    std::vector<TypeId> group_types, agg_types, agg_arg_types;
    bool child_consumed = false;
    // across GPUs: the row counts of every rank's shard, exchanged by the first run and reused by the next ones (a plan's
    // tables do not change under it: ColumnarScan holds the Table it was built over)
    std::shared_ptr<void> row_count_cache_;
};

struct OrderBy : public Operator {
    struct SortKey {
        std::unique_ptr<Expr> expr;
        bool asc;
    };
    OrderBy(std::unique_ptr<Operator> input, std::vector<SortKey> sort_keys);
    BQ_OPERATOR_LIFECYCLE
    // Limit above an OrderBy asks for the first k rows only (top-k instead of a full sort)
    std::shared_ptr<gpu::DeviceRelation> sorted_prefix(int64_t limit);
    std::shared_ptr<gpu::DeviceRelation> sort_relation(const std::shared_ptr<gpu::DeviceRelation>& in, int64_t limit);
    // across GPUs, no LIMIT: this rank's rows exchanged so that it holds one key range of the first sort key
    std::shared_ptr<gpu::DeviceRelation> range_partition(const std::shared_ptr<gpu::DeviceRelation>& in);
private:
    std::unique_ptr<Operator> input_;
    std::vector<SortKey> sort_keys;
    ExprBindings bindings;
    bool child_consumed = false;
};

struct Limit : public Operator {
    Limit(std::unique_ptr<Operator> c, int64_t n);
    BQ_OPERATOR_LIFECYCLE
private:
    std::unique_ptr<Operator> input_;
    int64_t limit;
};

// reference: include/exec/physical_planner.h:11
std::unique_ptr<Operator> build_physical_plan(const LogicalOp* logical, const Catalog& catalog);

}  // namespace bosql
