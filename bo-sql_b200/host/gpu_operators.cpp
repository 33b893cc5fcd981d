// gpu_operators.cpp — the seven physical operators of the reference (src/exec/operator.cpp) as GPU plan nodes.
//
// Constructors validate and name/type their output exactly like the reference's (cited per constructor), because
// callers observe output_names()/output_types()/dictionary() and the thrown messages.  Execution differs
// completely: nothing loops over rows on the host.  A blocking operator asks its child to describe() itself;
// scan / selection / join chains fold into one Pipeline that gpu_plan.cpp runs as a single fused kernel.  Children
// that cannot be described (an aggregate under a sort, a projection ...) are materialised as device relations and
// the same kernels run over those.  next() pages the finished device relation out in <= 4096-row batches.
#include <algorithm>
#include <cctype>
#include <cstring>

#include "bosql_operator.hpp"
#include "exchange.hpp"
#include "gpu_plan.hpp"

namespace bosql {

using gpu::check;
using gpu::context;
using gpu::DeviceRelation;
using gpu::DeviceRelationPtr;
using gpu::DevColPtr;
using gpu::PipeCol;
using gpu::Pipeline;

ExprBindings make_bindings(const std::vector<std::string>& names, const std::vector<TypeId>& types, Dictionary* dictionary) {
    ExprBindings b;
    b.column_names = &names;
    b.column_types = &types;
    b.dictionary = dictionary;
    for (size_t i = 0; i < names.size(); ++i) b.name_to_index[names[i]] = i;
    return b;
}

namespace {

constexpr size_t kBatchRows = 4096;    // include/exec/operator.hpp:34 and src/exec/operator.cpp:765,1020,1129

std::string upper(std::string s) {
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return static_cast<char>(std::toupper(c)); });
    return s;
}

// infer_type, src/exec/operator.cpp:84-146 (declared output type of a projection / aggregate argument)
TypeId infer_type(const Expr* e, const ExprBindings& b) {
    switch (e->type) {
        case ExprType::COLUMN_REF: {
            auto it = b.name_to_index.find(e->str_val);
            if (it == b.name_to_index.end()) throw std::runtime_error("Unknown column: " + e->str_val);
            return (*b.column_types)[it->second];
        }
        case ExprType::LITERAL_INT: return TypeId::INT64;
        case ExprType::LITERAL_DOUBLE: return TypeId::DOUBLE;
        case ExprType::LITERAL_STRING: return TypeId::STRING;
        case ExprType::BINARY_OP:
            if (e->op >= BinaryOp::ADD && e->op <= BinaryOp::DIV) {
                TypeId l = infer_type(e->left.get(), b), r = infer_type(e->right.get(), b);
                return (l == TypeId::DOUBLE || r == TypeId::DOUBLE) ? TypeId::DOUBLE : TypeId::INT64;
            }
            return TypeId::INT64;
        case ExprType::FUNC_CALL: {
            if (e->func_name.empty()) throw std::runtime_error("Function call unsupported in projection");
            std::string f = upper(e->func_name);
            if (f == "COUNT") return TypeId::INT64;
            TypeId arg = TypeId::INT64;
            if (!e->args.empty()) arg = infer_type(e->args[0].get(), b);
            if (f == "SUM") return arg == TypeId::DOUBLE ? TypeId::DOUBLE : TypeId::INT64;
            if (f == "AVG") return TypeId::DOUBLE;
            throw std::runtime_error("Function call unsupported in projection");
        }
    }
    throw std::runtime_error("Cannot infer expression type");
}

std::vector<PipeCol> pipe_cols(const std::vector<std::string>& names, const std::vector<TypeId>& types, const DeviceRelation& rel) {
    std::vector<PipeCol> out(names.size());
    for (size_t i = 0; i < names.size(); ++i) {
        out[i].name = names[i];
        out[i].type = types[i];
        out[i].dev = rel.cols.at(i);
    }
    return out;
}

gpu::KeyStats stats_from_meta(const ColumnMeta& m, size_t table_rows) {
    gpu::KeyStats s;
    const ColumnStats& c = m.stats;
    s.ndv = c.ndv;
    s.table_rows = table_rows;
    switch (m.type) {
        case TypeId::INT64: s.min_key = c.min_i64; s.max_key = c.max_i64; break;
        case TypeId::DOUBLE: s.min_key = gpu::f64_key(c.min_f64); s.max_key = gpu::f64_key(c.max_f64); break;
        case TypeId::DATE32: s.min_key = c.min_date; s.max_key = c.max_date; break;
        case TypeId::STRING: s.min_key = 0; s.max_key = c.ndv ? static_cast<int64_t>(c.ndv) - 1 : -1; break;
    }
    // hand-built metas (tests/test_execution.cpp:31-36) leave every statistic zero: that means "unknown".
    // STRING min/max are never recorded by the loader, and ids may come from a shared dictionary: measure them.
    s.known = m.type != TypeId::STRING && (s.max_key > s.min_key || (c.ndv == 1 && s.max_key == s.min_key));
    return s;
}

template <typename T>
std::shared_ptr<void> download(const DevColPtr& col, size_t rows) {
    std::shared_ptr<void> buf = gpu::host_buffer(rows * sizeof(T));
    if (rows) check(bq_col_read_async(context(), col->h, 0, rows, buf.get()));      // page_out waits once for all columns
    return buf;
}

}  // namespace

// ---- result paging -----------------------------------------------------------------------------------------
void Operator::reset_paging() {
    result_.reset();
    host_cols_.clear();
    emit_offset_ = 0;
    paged_ = false;
}

bool Operator::page_out(ExecBatch& out) {
    if (!paged_) {
        gpu::PhaseTrace trace;
        result_ = device_result();
        trace.mark("device_result total");
        host_cols_.clear();
        for (size_t c = 0; c < result_->cols.size(); ++c) {
            switch (types_[c]) {
                case TypeId::INT64: host_cols_.push_back(download<int64_t>(result_->cols[c], result_->rows)); break;
                case TypeId::DOUBLE: host_cols_.push_back(download<double>(result_->cols[c], result_->rows)); break;
                case TypeId::STRING: host_cols_.push_back(download<uint32_t>(result_->cols[c], result_->rows)); break;
                case TypeId::DATE32: host_cols_.push_back(download<int32_t>(result_->cols[c], result_->rows)); break;
            }
        }
        check(bq_ctx_sync(context()));
        trace.mark("result to host");
        emit_offset_ = 0;
        paged_ = true;
    }
    if (emit_offset_ >= result_->rows) return false;      // never an empty batch with `true`
    const size_t take = std::min(kBatchRows, result_->rows - emit_offset_);
    out.clear();
    for (size_t c = 0; c < host_cols_.size(); ++c) {
        const char* base = static_cast<const char*>(host_cols_[c].get()) + emit_offset_ * type_width(types_[c]);
        out.columns.push_back({base, types_[c], take, host_cols_[c]});
    }
    out.length = take;
    emit_offset_ += take;
    return true;
}

bool Operator::host_result(std::vector<std::shared_ptr<void>>& cols, size_t& rows) const {
    if (!paged_ || !result_) return false;
    cols = host_cols_;
    rows = result_->rows;
    return true;
}

// ---- ColumnarScan (src/exec/operator.cpp:321-386) -------------------------------------------------------------
ColumnarScan::ColumnarScan(Table* t, std::vector<size_t> idx, size_t batch)
    : table(t), indices(std::move(idx)), offset(0), batch_size(batch) {
    if (!table) throw std::runtime_error("Scan table is null");
    if (indices.empty())
        for (size_t i = 0; i < table->columns.size(); ++i) indices.push_back(i);      // empty list = all columns (:326-331)
    for (size_t i : indices) {
        names_.push_back(table->columns[i].name);
        types_.push_back(table->columns[i].data->type());
    }
    dict_ = table->dict.get();
}

void ColumnarScan::open() {
    offset = 0;
    reset_paging();
}

bool ColumnarScan::next(ExecBatch& out) {
    if (indices.empty()) return false;
    bool host_backed = true;
    for (size_t i : indices) host_backed = host_backed && table->columns[i].data->host_data() != nullptr;
    if (!host_backed) return page_out(out);        // device-only table: page rows out of HBM
    // host-backed table: zero-copy slices of ColumnVector<T>::data, as the reference (:345-384)
    const size_t rows = table->columns[indices[0]].data->size();
    if (offset >= rows) return false;
    const size_t take = std::min(batch_size, rows - offset);
    out.clear();
    for (size_t i : indices) {
        const Column& c = *table->columns[i].data;
        const char* base = static_cast<const char*>(c.host_data()) + offset * type_width(c.type());
        out.columns.push_back({base, c.type(), take, {}});
    }
    out.length = take;
    offset += take;
    return true;
}

void ColumnarScan::close() {}

DeviceRelationPtr ColumnarScan::device_result() {
    auto rel = std::make_shared<DeviceRelation>();
    for (size_t i : indices) rel->cols.push_back(gpu::mirror_of(*table->columns[i].data));
    rel->rows = table_rows();
    return rel;
}

size_t ColumnarScan::table_rows() const { return indices.empty() ? 0 : table->columns[indices[0]].data->size(); }

DeviceRelationPtr ColumnarScan::device_window(size_t begin, size_t end) {
    DeviceRelationPtr all = device_result();
    end = std::min(end, all->rows);
    begin = std::min(begin, end);
    if (begin == 0 && end == all->rows) return all;
    auto rel = std::make_shared<DeviceRelation>();
    for (auto& c : all->cols) rel->cols.push_back(gpu::view_of(c, begin, end));
    rel->rows = end - begin;
    return rel;
}

// the batches a consumer of `want` rows pulls: ceil(want / batch) of them
DeviceRelationPtr ColumnarScan::device_prefix(size_t want) {
    const size_t b = batch_size ? batch_size : kBatchRows;
    const size_t n = table_rows();
    const size_t batches = want / b + (want % b ? 1 : 0);
    return device_window(0, batches > n / b ? n : batches * b);
}

bool ColumnarScan::describe(Pipeline& p) {
    DeviceRelationPtr rel = device_result();
    p = Pipeline{};
    p.rows = rel->rows;
    p.dict = dict_;
    for (size_t k = 0; k < indices.size(); ++k) {
        PipeCol c;
        c.name = names_[k];
        c.type = types_[k];
        c.dev = rel->cols[k];
        if (meta_)
            for (const ColumnMeta& m : meta_->columns)
                if (m.name == c.name && m.type == c.type) c.stats = stats_from_meta(m, meta_->row_count);
        p.cols.push_back(std::move(c));
    }
    return true;
}

// ---- Selection (src/exec/operator.cpp:388-433) -------------------------------------------------------------------
Selection::Selection(std::unique_ptr<Operator> c, std::unique_ptr<Expr> pred) : input_(std::move(c)), predicate(std::move(pred)) {
    if (!input_) throw std::runtime_error("Selection input_ is null");
    names_ = input_->output_names();
    types_ = input_->output_types();
    dict_ = input_->dictionary();
    bindings = make_bindings(names_, types_, dict_);
}

void Selection::open() {
    input_->open();
    reset_paging();
}
bool Selection::next(ExecBatch& out) { return page_out(out); }
void Selection::close() { input_->close(); }

bool Selection::describe(Pipeline& p) {
    if (!input_->describe(p)) return false;
    if (predicate) gpu::split_conjuncts(predicate.get(), dict_, p.conjuncts);
    return true;
}

DeviceRelationPtr Selection::select_from(const DeviceRelationPtr& in) {
    if (!predicate) return in;                       // null predicate = pass-through (:406-409)
    std::vector<gpu::Conjunct> conj;
    gpu::split_conjuncts(predicate.get(), dict_, conj);
    std::vector<const gpu::Conjunct*> ptrs;
    for (const auto& c : conj) ptrs.push_back(&c);
    DeviceRelationPtr out = gpu::run_selection(pipe_cols(names_, types_, *in), in->rows, ptrs);
    out->replicated = in->replicated;
    return out;
}

DeviceRelationPtr Selection::device_result() { return select_from(input_->device_result()); }

// The reference's Selection evaluates its predicate batch by batch and is asked for no further batch once the consumer
// has its rows (:403-429 under Limit::next :577-613), so a row beyond that point is never evaluated - which matters when the
// predicate can throw there (integer division by zero, H10).  Over a scan the same is done in windows of whole batches that
// grow geometrically; a window that throws is redone batch by batch, so the first failing batch is reached exactly when
// the reference reaches it.  Across GPUs the number of windows would differ per rank while predicate programs exchange
// their outcome collectively, so the sharded path evaluates its whole shard.
DeviceRelationPtr Selection::device_prefix(size_t want) {
    auto* scan = dynamic_cast<ColumnarScan*>(input_.get());
    if (!scan || !predicate || gpu::exchange().active) return select_from(input_->device_prefix(predicate ? static_cast<size_t>(-1) : want));
    const size_t n = scan->table_rows();
    if (want >= n) return select_from(input_->device_result());      // every batch would be pulled anyway
    const size_t b = scan->batch_rows() ? scan->batch_rows() : kBatchRows;
    std::vector<gpu::Conjunct> conj;
    gpu::split_conjuncts(predicate.get(), dict_, conj);
    std::vector<const gpu::Conjunct*> ptrs;
    for (const auto& c : conj) ptrs.push_back(&c);
    std::vector<DeviceRelationPtr> parts;
    size_t have = 0, at = 0;
    size_t window = std::max(b, (want + b - 1) / b * b);
    // rows [begin, end) of the scan: the selected rows, cut after the batch in which the consumer has its `want` rows -
    // an operator above (a Project that can throw) must not see rows of batches the reference never pulls
    auto take = [&](size_t begin, size_t end) {
        DeviceRelationPtr in = scan->device_window(begin, end);
        std::vector<PipeCol> cols = pipe_cols(names_, types_, *in);
        DevColPtr ids = gpu::select_rowids(cols, in->rows, ptrs);
        if (have + ids->rows() >= want && end - begin > b) {
            uint32_t local = 0;                               // position, inside the window, of the row that completes `want`
            check(bq_col_read(context(), ids->h, want - have - 1, 1, &local));
            const size_t cut = std::min(end, begin + (local / b + 1) * b);
            if (cut < end) {
                in = scan->device_window(begin, cut);
                cols = pipe_cols(names_, types_, *in);
                ids = gpu::select_rowids(cols, in->rows, ptrs);
            }
        }
        have += ids->rows();
        if (ids->rows()) parts.push_back(gpu::gather_rows(cols, ids));
    };
    while (at < n && have < want) {
        const size_t end = std::min(n, at + window);
        try {
            take(at, end);
        } catch (const std::runtime_error& e) {
            if (std::string(e.what()) != "Division by zero" || end - at <= b) throw;
            for (size_t s = at; s < end && have < want; s += b) take(s, std::min(end, s + b));      // throws at the batch the reference throws at
        }
        at = end;
        window *= 4;
    }
    if (parts.empty()) return gpu::empty_relation(types_);
    if (parts.size() == 1) return parts[0];
    return gpu::concat_relations(parts, types_);
}

// ---- Project (src/exec/operator.cpp:435-559) ------------------------------------------------------------------------
Project::Project(std::unique_ptr<Operator> c, std::vector<std::unique_ptr<Expr>> exprs, std::vector<std::string> alias_list)
    : input_(std::move(c)), expressions(std::move(exprs)), aliases(std::move(alias_list)) {
    if (!input_) throw std::runtime_error("Project input_ is null");
    input_names = input_->output_names();
    input_types = input_->output_types();
    dict_ = input_->dictionary();
    bindings = make_bindings(input_names, input_types, dict_);
    direct_indices.assign(expressions.size(), -1);
    for (size_t i = 0; i < expressions.size(); ++i) {
        const Expr* e = expressions[i].get();
        types_.push_back(infer_type(e, bindings));
        const bool has_alias = i < aliases.size() && !aliases[i].empty();
        names_.push_back(has_alias ? aliases[i] : (e->type == ExprType::COLUMN_REF ? e->str_val : "expr"));
        if (e->type == ExprType::COLUMN_REF) {
            auto it = bindings.name_to_index.find(e->str_val);
            if (it != bindings.name_to_index.end()) direct_indices[i] = static_cast<int>(it->second);
            continue;
        }
        // an expression that the child already computed under that name (aggregate outputs), :468-490
        const std::string candidate = has_alias ? aliases[i] : e->to_string();
        auto it = std::find(input_names.begin(), input_names.end(), candidate);
        if (it != input_names.end()) {
            direct_indices[i] = static_cast<int>(it - input_names.begin());
        } else if (e->type == ExprType::FUNC_CALL) {
            const std::string want = upper(candidate);
            for (size_t k = 0; k < input_names.size(); ++k)
                if (upper(input_names[k]) == want) {
                    direct_indices[i] = static_cast<int>(k);
                    break;
                }
        }
    }
}

void Project::open() {
    input_->open();
    reset_paging();
}
bool Project::next(ExecBatch& out) { return page_out(out); }
void Project::close() { input_->close(); }

DeviceRelationPtr Project::device_result() { return project_from(input_->device_result()); }
// Project::next evaluates every row of every batch it is asked for (:498-555): the child's prefix, whole
DeviceRelationPtr Project::device_prefix(size_t want) { return project_from(input_->device_prefix(want)); }

DeviceRelationPtr Project::project_from(const DeviceRelationPtr& in) {
    auto out = std::make_shared<DeviceRelation>();
    out->replicated = in->replicated;
    out->rows = in->rows;
    std::vector<PipeCol> cols = pipe_cols(input_names, input_types, *in);
    for (size_t i = 0; i < expressions.size(); ++i) {
        if (direct_indices[i] >= 0) {
            out->cols.push_back(in->cols[direct_indices[i]]);
            continue;
        }
        // no row is ever evaluated, so nothing can throw (:505-551); across GPUs an empty shard still takes part in the
        // outcome exchange of a program that can fail on another rank's rows
        if (in->rows == 0 && !gpu::exchange().active) {
            bq_col* h = nullptr;
            check(bq_col_alloc(context(), static_cast<int>(types_[i]), 0, &h));
            out->cols.push_back(gpu::adopt(h));
            continue;
        }
        // evaluate_expr per row; the declared type equals the value's static type for every well-typed
        // expression, so the coercions of :512-551 are identities
        out->cols.push_back(gpu::eval_to_column(expressions[i].get(), cols, in->rows, dict_, types_[i], false));
    }
    return out;
}

// ---- Limit (src/exec/operator.cpp:561-620) -----------------------------------------------------------------------------
Limit::Limit(std::unique_ptr<Operator> c, int64_t n) : input_(std::move(c)), limit(n) {
    if (!input_) throw std::runtime_error("Limit input_ is null");
    names_ = input_->output_names();
    types_ = input_->output_types();
    dict_ = input_->dictionary();
}

void Limit::open() {
    input_->open();
    reset_paging();
}
bool Limit::next(ExecBatch& out) { return page_out(out); }
void Limit::close() { input_->close(); }

DeviceRelationPtr Limit::device_result() {
    const int64_t want = limit < 0 ? 0 : limit;
    if (want == 0) return gpu::empty_relation(types_);        // next() never pulls the child (:578-580): nothing is evaluated
    if (auto* ob = dynamic_cast<OrderBy*>(input_.get())) return ob->sorted_prefix(want);      // top-k
    // the child computes what a consumer of `want` rows would have made it compute, not its whole output
    DeviceRelationPtr in = input_->device_prefix(static_cast<size_t>(want));
    auto prefix = [&](const DeviceRelationPtr& rel) {
        if (static_cast<uint64_t>(want) >= rel->rows) return rel;
        auto out = std::make_shared<DeviceRelation>();
        out->rows = static_cast<size_t>(want);
        out->replicated = rel->replicated;
        for (auto& c : rel->cols) {
            bq_col* h = nullptr;
            check(bq_slice(context(), c->h, 0, out->rows, &h));      // copy_range, :51-82
            out->cols.push_back(gpu::adopt(h));
        }
        return DeviceRelationPtr(out);
    };
    if (!gpu::exchange().active || in->replicated) return prefix(in);
    // across GPUs: the first `limit` rows in table order = rank order; each rank contributes at most `limit` of its own
    DeviceRelationPtr all = gpu::all_gather_relation(prefix(in), types_);
    all->replicated = true;
    return prefix(all);
}

// ---- HashJoin (src/exec/operator.cpp:671-858) ------------------------------------------------------------------------------
HashJoin::HashJoin(std::unique_ptr<Operator> left, std::unique_ptr<Operator> right, std::vector<std::string> left_keys,
                   std::vector<std::string> right_keys, std::unique_ptr<Expr> residual)
    : left_child(std::move(left)), right_child(std::move(right)), left_key_names(std::move(left_keys)),
      right_key_names(std::move(right_keys)), residual_filter(std::move(residual)) {
    if (!left_child || !right_child) throw std::runtime_error("Join operands cannot be null");
    left_names = left_child->output_names();
    left_types = left_child->output_types();
    right_names = right_child->output_names();
    right_types = right_child->output_types();
    names_ = left_names;
    names_.insert(names_.end(), right_names.begin(), right_names.end());
    types_ = left_types;
    types_.insert(types_.end(), right_types.begin(), right_types.end());

    // dictionary choice: left if the left side has a STRING column, else right (:694-704)
    Dictionary* ld = left_child->dictionary();
    Dictionary* rd = right_child->dictionary();
    auto has_string = [](const std::vector<TypeId>& t) { return std::find(t.begin(), t.end(), TypeId::STRING) != t.end(); };
    if (has_string(left_types) && ld) dict_ = ld;
    else if (has_string(right_types) && rd) dict_ = rd;
    else dict_ = ld ? ld : rd;

    auto resolve = [](const std::vector<std::string>& keys, const std::vector<std::string>& cols) {
        std::vector<size_t> out;
        for (const auto& k : keys) {
            auto it = std::find(cols.begin(), cols.end(), k);
            if (it == cols.end()) throw std::runtime_error("Join key not found: " + k);
            out.push_back(static_cast<size_t>(it - cols.begin()));
        }
        return out;
    };
    left_key_indices = resolve(left_key_names, left_names);
    right_key_indices = resolve(right_key_names, right_names);
    if (left_key_indices.size() != right_key_indices.size()) throw std::runtime_error("Join key cardinality mismatch");
    for (size_t i : left_key_indices) left_key_types.push_back(left_types[i]);
    for (size_t i : right_key_indices) right_key_types.push_back(right_types[i]);
}

void HashJoin::open() {
    right_child->open();      // the reference drains the build side here (:748-759); ours is built lazily on the device
    right_child->close();
    left_child->open();
    reset_paging();
}
bool HashJoin::next(ExecBatch& out) { return page_out(out); }
void HashJoin::close() { left_child->close(); }

bool HashJoin::describe(Pipeline& p) {
    if (left_key_indices.size() > 1) return false;
    Pipeline l, r;
    if (!left_child->describe(l) || !right_child->describe(r)) return false;
    if (l.joined || r.joined) return false;
    p = Pipeline{};
    p.rows = l.rows;
    p.build_rows = r.rows;
    p.joined = true;
    p.dict = dict_;
    for (auto& c : l.cols) {
        c.side = 0;
        p.cols.push_back(std::move(c));
    }
    const size_t n_left = p.cols.size();
    for (auto& c : r.cols) {
        c.side = 1;
        p.cols.push_back(std::move(c));
    }
    if (left_key_indices.empty()) {
        p.cross_join = true;       // every row matches every row
    } else {
        p.probe_key = static_cast<int>(left_key_indices[0]);
        p.build_key = static_cast<int>(n_left + right_key_indices[0]);
    }
    // predicates sitting under the join stay bound to their side: they only name that side's columns, but a name
    // that also exists on the other side would rebind after the merge, so such plans are not fused
    auto names_clash = [&]() {
        for (size_t i = 0; i < n_left; ++i)
            for (size_t j = n_left; j < p.cols.size(); ++j)
                if (p.cols[i].name == p.cols[j].name) return true;
        return false;
    };
    if ((!l.conjuncts.empty() || !r.conjuncts.empty()) && names_clash()) return false;
    for (auto& c : l.conjuncts) p.conjuncts.push_back(std::move(c));
    for (auto& c : r.conjuncts) p.conjuncts.push_back(std::move(c));
    return true;
}

DeviceRelationPtr HashJoin::device_result() {
    DeviceRelationPtr l = left_child->device_result();
    DeviceRelationPtr r = right_child->device_result();
    // across GPUs both inputs are row shards: broadcast the build side, so this rank emits the join rows of its probe rows
    if (gpu::exchange().active) r = gpu::all_gather_relation(r, right_child->output_types());
    bq_ctx* ctx = context();
    DevColPtr probe_rows, build_rows;
    if (l->rows == 0 || r->rows == 0) return gpu::empty_relation(types_);
    if (left_key_indices.empty()) {
        // cross product in probe order: pair i = (i / n_right, i % n_right)
        const uint64_t total = static_cast<uint64_t>(l->rows) * r->rows;
        if (total > 0xFFFFFFFFull) throw std::runtime_error("join result exceeds 2^32 rows");
        bq_col* iota = nullptr;
        check(bq_col_alloc(ctx, BQ_INT64, total, &iota));
        DevColPtr idx = gpu::adopt(iota);
        bq_gen_spec g{};
        g.dist = BQ_GEN_SEQ;
        check(bq_col_generate(ctx, idx->h, &g, 0));
        const int64_t nr = static_cast<int64_t>(r->rows);
        bq_insn q[3] = {{BQ_OP_COL, 0, {0}}, {BQ_OP_IMM_I, 0, {nr}}, {BQ_OP_DIV_I, 0, {0}}};
        bq_insn m[7] = {{BQ_OP_COL, 0, {0}}, {BQ_OP_COL, 0, {0}}, {BQ_OP_IMM_I, 0, {nr}}, {BQ_OP_DIV_I, 0, {0}},
                        {BQ_OP_IMM_I, 0, {nr}}, {BQ_OP_MUL_I, 0, {0}}, {BQ_OP_SUB_I, 0, {0}}};
        const bq_col* cols[1] = {idx->h};
        bq_col *pr = nullptr, *br = nullptr;
        check(bq_eval(ctx, q, 3, cols, 1, 0, total, BQ_STRING, &pr));
        probe_rows = gpu::adopt(pr);
        check(bq_eval(ctx, m, 7, cols, 1, 0, total, BQ_STRING, &br));
        build_rows = gpu::adopt(br);
    } else {
        // KeyEqual: a component whose two sides differ in TypeId never matches (:652)
        for (size_t i = 0; i < left_key_types.size(); ++i)
            if (left_key_types[i] != right_key_types[i]) return gpu::empty_relation(types_);
        DevColPtr bk = r->cols[right_key_indices[0]];
        DevColPtr pk = l->cols[left_key_indices[0]];
        bq_join_spec js{};
        js.key = bk->h;
        js.row_begin = 0;
        js.row_end = r->rows;
        js.need_rows = 1;
        js.kind = BQ_JOIN_HASH;
        if (right_key_types[0] != TypeId::DOUBLE) {
            int64_t lo = 0, hi = -1;
            check(bq_col_minmax(ctx, bk->h, &lo, &hi));
            js.kind = BQ_JOIN_AUTO;
            js.key_min = lo;
            js.key_max = hi;
        }
        bq_join* j = nullptr;
        check(bq_join_build(ctx, &js, &j));
        bq_col *pr = nullptr, *br = nullptr;
        int rc = bq_join_probe(ctx, j, pk->h, nullptr, 0, l->rows, &pr, &br);
        bq_join_free(ctx, j);
        check(rc);
        probe_rows = gpu::adopt(pr);
        build_rows = gpu::adopt(br);
        // Several key columns (Key holds one Datum per column, include/exec/operator.hpp:107-117; build_key :847-858): the
        // table is keyed on the first column, the candidate pairs are then filtered on the remaining components with
        // KeyEqual's per-type `==` (:646-667).  The compaction is stable, so pairs stay in probe order and, per probe row,
        // in build insertion order (:802-816).
        for (size_t first = 1; first < left_key_indices.size() && probe_rows->rows() > 0; first += BQ_MAX_PROGRAM_COLS / 2) {
            const size_t last = std::min(left_key_indices.size(), first + BQ_MAX_PROGRAM_COLS / 2);
            std::vector<DevColPtr> sides;
            std::vector<const bq_col*> prog_cols;
            std::vector<bq_insn> code;
            for (size_t k = first; k < last; ++k) {
                bq_col *lg = nullptr, *rg = nullptr;
                check(bq_gather(ctx, l->cols[left_key_indices[k]]->h, probe_rows->h, &lg));
                sides.push_back(gpu::adopt(lg));
                check(bq_gather(ctx, r->cols[right_key_indices[k]]->h, build_rows->h, &rg));
                sides.push_back(gpu::adopt(rg));
                const int c = static_cast<int>(2 * (k - first));
                code.push_back({BQ_OP_COL, c, {0}});
                code.push_back({BQ_OP_COL, c + 1, {0}});
                code.push_back({left_key_types[k] == TypeId::DOUBLE ? BQ_OP_EQ_F : BQ_OP_EQ_I, 0, {0}});
                if (k > first) code.push_back({BQ_OP_AND, 0, {0}});
            }
            for (auto& c : sides) prog_cols.push_back(c->h);
            bq_col* m = nullptr;
            check(bq_eval(ctx, code.data(), static_cast<int>(code.size()), prog_cols.data(), static_cast<int>(prog_cols.size()), 0,
                          probe_rows->rows(), BQ_INT64, &m));
            DevColPtr mask = gpu::adopt(m);
            bq_select_spec ss{};
            ss.mask = mask->h;
            ss.row_begin = 0;
            ss.row_end = probe_rows->rows();
            bq_col* keep = nullptr;
            check(bq_select(ctx, &ss, &keep));
            DevColPtr kept = gpu::adopt(keep);
            bq_col *p2 = nullptr, *b2 = nullptr;
            check(bq_gather(ctx, probe_rows->h, kept->h, &p2));
            probe_rows = gpu::adopt(p2);
            check(bq_gather(ctx, build_rows->h, kept->h, &b2));
            build_rows = gpu::adopt(b2);
        }
    }
    auto out = std::make_shared<DeviceRelation>();
    out->rows = probe_rows->rows();
    for (auto& c : l->cols) {
        bq_col* g = nullptr;
        check(bq_gather(ctx, c->h, probe_rows->h, &g));
        out->cols.push_back(gpu::adopt(g));
    }
    for (auto& c : r->cols) {
        bq_col* g = nullptr;
        check(bq_gather(ctx, c->h, build_rows->h, &g));
        out->cols.push_back(gpu::adopt(g));
    }
    return out;
}

// ---- HashAggregate (src/exec/operator.cpp:907-1074) ------------------------------------------------------------------------
HashAggregate::HashAggregate(std::unique_ptr<Operator> child_op, std::vector<std::unique_ptr<Expr>> group_exprs_in,
                             std::vector<AggregateSpec> aggregates_in)
    : input_(std::move(child_op)), group_exprs(std::move(group_exprs_in)), aggregates(std::move(aggregates_in)) {
    if (!input_) throw std::runtime_error("HashAggregate input_ is null");
    dict_ = input_->dictionary();
    child_bindings = make_bindings(input_->output_names(), input_->output_types(), dict_);
    for (size_t i = 0; i < group_exprs.size(); ++i) {
        TypeId t = infer_type(group_exprs[i].get(), child_bindings);
        group_types.push_back(t);
        names_.push_back(group_exprs[i]->type == ExprType::COLUMN_REF ? group_exprs[i]->str_val : "group" + std::to_string(i + 1));
        types_.push_back(t);
    }
    for (const auto& a : aggregates) {
        TypeId arg = TypeId::INT64;
        if (a.arg && a.func_name != "COUNT") arg = infer_type(a.arg.get(), child_bindings);
        TypeId res = TypeId::INT64;                        // COUNT, and any unknown function (:940-947)
        if (a.func_name == "SUM") res = arg == TypeId::DOUBLE ? TypeId::DOUBLE : TypeId::INT64;
        else if (a.func_name == "AVG") res = TypeId::DOUBLE;
        agg_arg_types.push_back(arg);
        agg_types.push_back(res);
        names_.push_back(!a.alias.empty() ? a.alias : a.func_name + "(" + (a.arg ? a.arg->to_string() : std::string("*")) + ")");
        types_.push_back(res);
    }
}

void HashAggregate::open() {
    child_consumed = false;
    input_->open();
    reset_paging();
}

bool HashAggregate::next(ExecBatch& out) {
    bool more = page_out(out);
    if (!child_consumed) {
        input_->close();
        child_consumed = true;
    }
    return more;
}

void HashAggregate::close() {
    if (!child_consumed) {
        input_->close();
        child_consumed = true;
    }
    reset_paging();
}

DeviceRelationPtr HashAggregate::device_result() {
    gpu::AggRequest req;
    req.group_exprs = &group_exprs;
    req.group_types = group_types;
    req.dict = dict_;
    if (!row_count_cache_) row_count_cache_ = std::make_shared<gpu::RowCountCache>();
    req.row_cache = static_cast<gpu::RowCountCache*>(row_count_cache_.get());
    for (size_t i = 0; i < aggregates.size(); ++i) {
        const auto& a = aggregates[i];
        if (a.func_name != "COUNT" && a.func_name != "SUM" && a.func_name != "AVG")
            throw std::runtime_error("unsupported aggregate function: " + a.func_name);
        req.aggs.push_back({a.func_name, a.arg.get(), agg_types[i]});
    }
    Pipeline p;
    if (input_->describe(p)) {
        if (DeviceRelationPtr r = gpu::run_aggregate(p, req)) return r;
    }
    // not a fusable chain (or it mixes both join sides in one predicate): materialise the child, then the same kernels
    DeviceRelationPtr in = input_->device_result();
    Pipeline plain;
    plain.rows = in->rows;
    plain.dict = dict_;
    plain.cols = pipe_cols(input_->output_names(), input_->output_types(), *in);
    DeviceRelationPtr r = gpu::run_aggregate(plain, req);
    if (!r) throw std::runtime_error("internal: aggregate over a materialised relation was not planned");
    return r;
}

// ---- OrderBy (src/exec/operator.cpp:1076-1161) -----------------------------------------------------------------------------
OrderBy::OrderBy(std::unique_ptr<Operator> child_op, std::vector<SortKey> sort_keys_in)
    : input_(std::move(child_op)), sort_keys(std::move(sort_keys_in)) {
    if (!input_) throw std::runtime_error("OrderBy input_ is null");
    names_ = input_->output_names();
    types_ = input_->output_types();
    dict_ = input_->dictionary();
    bindings = make_bindings(names_, types_, dict_);
}

void OrderBy::open() {
    child_consumed = false;
    input_->open();
    reset_paging();
}

bool OrderBy::next(ExecBatch& out) {
    bool more = page_out(out);
    if (!child_consumed) {
        input_->close();
        child_consumed = true;
    }
    return more;
}

void OrderBy::close() {
    if (!child_consumed) {
        input_->close();
        child_consumed = true;
    }
    reset_paging();
}

DeviceRelationPtr OrderBy::device_result() { return sorted_prefix(-1); }

DeviceRelationPtr OrderBy::sorted_prefix(int64_t limit) {
    DeviceRelationPtr in = input_->device_result();
    if (!gpu::exchange().active || in->replicated) {
        DeviceRelationPtr out = sort_relation(in, limit);
        out->replicated = in->replicated;
        return out;
    }
    // Across GPUs the input is this rank's share of the rows: a local top-k first when there is a LIMIT (each rank's k
    // best rows are the only candidates), one all-gather, then the same sort on the gathered candidates - every rank ends
    // up with the complete ordered result (ORDER BY ... LIMIT k).  Without a LIMIT everything is gathered.
    if (limit < 0) {
        // no LIMIT: range-partition the rows by sampled splitters of the first sort key, exchange them all-to-all and sort
        // locally - rank r ends up with the r-th range of the ordered result (ORDER BY without LIMIT).  Small
        // inputs (and $BOSQL_SORT=gather) are gathered and sorted on every rank instead, which leaves the complete result
        // everywhere.
        const char* mode = std::getenv("BOSQL_SORT");
        const bool force_gather = mode && std::string(mode) == "gather", force_range = mode && std::string(mode) == "range";
        const int64_t total = gpu::exchange().host_sum(static_cast<int64_t>(in->rows));
        if (!force_gather && (force_range || total > (1 << 22)) && gpu::exchange().world() <= 16) {
            DeviceRelationPtr mine = range_partition(in);
            DeviceRelationPtr out = sort_relation(mine, -1);
            out->replicated = false;
            return out;
        }
    }
    DeviceRelationPtr local = limit >= 0 ? sort_relation(in, limit) : in;
    DeviceRelationPtr all = gpu::all_gather_relation(local, types_);
    DeviceRelationPtr out = sort_relation(all, limit);
    out->replicated = true;
    return out;
}

// Rows of this rank's shard redistributed so that rank r holds the r-th key range of the first sort key (all rows with equal
// first key land on one rank, so the later keys order them locally).  Splitters: every rank contributes 128 keys sampled at a
// regular stride; the pooled sample's quantiles cut the key space into one range per rank.
DeviceRelationPtr OrderBy::range_partition(const DeviceRelationPtr& in) {
    gpu::Exchange& x = gpu::exchange();
    bq_ctx* ctx = context();
    const int W = x.world();
    std::vector<PipeCol> cols = pipe_cols(names_, types_, *in);
    // the first sort key as a column (a plain reference, or evaluated - its outcome exchange runs on every rank)
    const SortKey& k0 = sort_keys.front();
    DevColPtr key;
    TypeId key_type;
    if (k0.expr->type == ExprType::COLUMN_REF) {
        auto it = bindings.name_to_index.find(k0.expr->str_val);
        if (it == bindings.name_to_index.end()) throw std::runtime_error("Unknown column: " + k0.expr->str_val);
        key = in->cols[it->second];
        key_type = types_[it->second];
    } else {
        key_type = gpu::value_type(k0.expr.get(), gpu::lookup_for(cols));
        key = gpu::eval_to_column(k0.expr.get(), cols, in->rows, dict_, key_type, false);
    }
    const bool fp = key_type == TypeId::DOUBLE;
    // ---- sample ----
    constexpr size_t kSample = 128;
    std::vector<int64_t> mine(kSample + 1, 0);
    const size_t take = std::min(kSample, in->rows);
    mine[kSample] = static_cast<int64_t>(take);
    if (take) {
        const int64_t stride = static_cast<int64_t>(in->rows / take);
        bq_col* iota = nullptr;
        check(bq_col_alloc(ctx, BQ_INT64, take, &iota));
        DevColPtr seq = gpu::adopt(iota);
        bq_gen_spec g{};
        g.dist = BQ_GEN_SEQ;
        check(bq_col_generate(ctx, seq->h, &g, 0));
        bq_insn prog[3] = {{BQ_OP_COL, 0, {0}}, {BQ_OP_IMM_I, 0, {stride}}, {BQ_OP_MUL_I, 0, {0}}};
        const bq_col* pc[1] = {seq->h};
        bq_col* ids = nullptr;
        check(bq_eval(ctx, prog, 3, pc, 1, 0, take, BQ_STRING, &ids));
        DevColPtr rowids = gpu::adopt(ids);
        bq_col* sampled = nullptr;
        check(bq_gather(ctx, key->h, rowids->h, &sampled));
        DevColPtr sk = gpu::adopt(sampled);
        if (type_width(key_type) == 8) {
            check(bq_col_read(ctx, sk->h, 0, take, mine.data()));          // INT64, or the bits of a DOUBLE
        } else {
            std::vector<int32_t> narrow(take);
            check(bq_col_read(ctx, sk->h, 0, take, narrow.data()));
            for (size_t i = 0; i < take; ++i)
                mine[i] = key_type == TypeId::STRING ? static_cast<int64_t>(static_cast<uint32_t>(narrow[i])) : static_cast<int64_t>(narrow[i]);
        }
    }
    auto all = x.host_gather(mine);
    std::vector<int64_t> pool;
    for (int r = 0; r < W; ++r) {
        const int64_t n = all[static_cast<size_t>(r) * (kSample + 1) + kSample];
        for (int64_t i = 0; i < n; ++i) pool.push_back(all[static_cast<size_t>(r) * (kSample + 1) + static_cast<size_t>(i)]);
    }
    auto as_f = [](int64_t b) {
        double d;
        std::memcpy(&d, &b, 8);
        return d;
    };
    // order of the OUTPUT: ascending keys go to rank 0 first; for DESC the largest do
    std::sort(pool.begin(), pool.end(), [&](int64_t a, int64_t b) {
        const bool lt = fp ? as_f(a) < as_f(b) : a < b;
        const bool gt = fp ? as_f(a) > as_f(b) : a > b;
        return k0.asc ? lt : gt;
    });
    std::vector<int64_t> split;                          // W - 1 splitters, in output order
    for (int r = 1; r < W && !pool.empty(); ++r) split.push_back(pool[std::min(pool.size() - 1, pool.size() * static_cast<size_t>(r) / static_cast<size_t>(W))]);
    // ---- destination rank of every row: how many splitters lie at or before its key (in output order) ----
    std::vector<bq_insn> code;
    for (size_t i = 0; i < split.size(); ++i) {
        code.push_back({BQ_OP_COL, 0, {0}});
        bq_insn imm{};
        imm.op = fp ? BQ_OP_IMM_F : BQ_OP_IMM_I;
        imm.imm.i = split[i];                            // (the bits of the double for IMM_F: same union)
        code.push_back(imm);
        const int ge = fp ? BQ_OP_GE_F : BQ_OP_GE_I, le = fp ? BQ_OP_LE_F : BQ_OP_LE_I;
        code.push_back({k0.asc ? ge : le, 0, {0}});
        if (i) code.push_back({BQ_OP_ADD_I, 0, {0}});
    }
    if (code.empty()) code.push_back({BQ_OP_IMM_I, 0, {0}});
    const bq_col* kc[1] = {key->h};
    bq_col* dest_h = nullptr;
    check(bq_eval(ctx, code.data(), static_cast<int>(code.size()), kc, 1, 0, in->rows, BQ_INT64, &dest_h));
    DevColPtr dest = gpu::adopt(dest_h);
    // ---- stable reorder by destination, sizes, all-to-all per column ----
    std::vector<int64_t> send_rows(static_cast<size_t>(W), 0);
    DeviceRelationPtr ordered = in;
    if (in->rows) {
        std::vector<bq_col*> hs;
        for (auto& c : in->cols) hs.push_back(c->h);
        hs.push_back(dest->h);
        bq_rel* shell = nullptr;
        check(bq_rel_create(ctx, hs.data(), static_cast<int>(hs.size()), &shell));
        const int by = static_cast<int>(hs.size()) - 1, asc = 1;
        bq_rel* sorted = nullptr;
        int rc = bq_rel_sort(ctx, shell, 1, &by, &asc, -1, &sorted);
        std::vector<bq_col*> back(hs.size());
        bq_rel_release(shell, back.data());
        check(rc);
        ordered = gpu::relation_from(sorted);
        // rows per destination: a dense COUNT over the (sorted) destination column
        bq_scan_spec cs{};
        cs.key = gpu::make_slot(ordered->cols.back(), {});
        cs.row_begin = 0;
        cs.row_end = in->rows;
        cs.group_mode = BQ_GROUP_DENSE;
        cs.key_min = 0;
        cs.key_max = W - 1;
        cs.n_out = 1;
        cs.out[0].func = BQ_AGG_COUNT;
        bq_rel* counted = nullptr;
        check(bq_scan_aggregate(ctx, &cs, &counted));
        DeviceRelationPtr cr = gpu::relation_from(counted);
        std::vector<int64_t> dk(cr->rows), dn(cr->rows);
        if (cr->rows) {
            check(bq_col_read(ctx, cr->cols[0]->h, 0, cr->rows, dk.data()));
            check(bq_col_read(ctx, cr->cols[1]->h, 0, cr->rows, dn.data()));
        }
        for (size_t i = 0; i < dk.size(); ++i) send_rows[static_cast<size_t>(dk[i])] = dn[i];
    }
    auto matrix = x.host_gather(send_rows);              // matrix[s*W + d] = rows rank s sends to rank d
    std::vector<int64_t> recv_rows(static_cast<size_t>(W));
    size_t total = 0;
    for (int s2 = 0; s2 < W; ++s2) {
        recv_rows[static_cast<size_t>(s2)] = matrix[static_cast<size_t>(s2) * W + static_cast<size_t>(x.rank())];
        total += static_cast<size_t>(recv_rows[static_cast<size_t>(s2)]);
    }
    if (total > 0xFFFFFFFFull) throw std::runtime_error("a rank would own more than 2^32 rows after the range partition");
    auto out = std::make_shared<DeviceRelation>();
    out->rows = total;
    for (size_t c = 0; c < types_.size(); ++c) {
        const size_t w = type_width(types_[c]);
        std::vector<int64_t> sb(static_cast<size_t>(W)), rb(static_cast<size_t>(W));
        for (int r = 0; r < W; ++r) {
            sb[static_cast<size_t>(r)] = send_rows[static_cast<size_t>(r)] * static_cast<int64_t>(w);
            rb[static_cast<size_t>(r)] = recv_rows[static_cast<size_t>(r)] * static_cast<int64_t>(w);
        }
        bq_col* recv = nullptr;
        check(bq_col_alloc(ctx, static_cast<int>(types_[c]), total, &recv));
        DevColPtr dst = gpu::adopt(recv);
        if (x.fn.all_to_all_v(x.fn.user, bq_col_ptr(ordered->cols[c]->h), sb.data(), bq_col_ptr(dst->h), rb.data(), bq_ctx_stream(ctx)))
            throw std::runtime_error("exchange callback failed: all_to_all_v");
        out->cols.push_back(dst);
    }
    check(bq_ctx_sync(ctx));          // `ordered` (the send side) is released when this returns
    return out;
}

DeviceRelationPtr OrderBy::sort_relation(const DeviceRelationPtr& in, int64_t limit) {
    // (across GPUs a rank with an empty shard still evaluates its sort-key programs: they exchange their outcome)
    const bool must_evaluate = gpu::exchange().active && !in->replicated;
    if ((in->rows == 0 || limit == 0) && !must_evaluate) return gpu::empty_relation(types_);
    bq_ctx* ctx = context();
    std::vector<PipeCol> cols = pipe_cols(names_, types_, *in);
    // sort columns: plain column references sort the column itself; anything else is evaluated first (:1105-1107)
    std::vector<DevColPtr> rel_cols = in->cols;
    std::vector<int> key_cols, asc;
    for (const auto& k : sort_keys) {
        int idx = -1;
        if (k.expr->type == ExprType::COLUMN_REF) {
            auto it = bindings.name_to_index.find(k.expr->str_val);
            if (it == bindings.name_to_index.end()) throw std::runtime_error("Unknown column: " + k.expr->str_val);
            idx = static_cast<int>(it->second);
        } else {
            TypeId t = gpu::value_type(k.expr.get(), gpu::lookup_for(cols));
            rel_cols.push_back(gpu::eval_to_column(k.expr.get(), cols, in->rows, dict_, t, false));
            idx = static_cast<int>(rel_cols.size()) - 1;
        }
        key_cols.push_back(idx);
        asc.push_back(k.asc ? 1 : 0);
    }
    if (in->rows == 0 || limit == 0) return gpu::empty_relation(types_);
    // Groups of a dense aggregate arrive in key order with distinct keys: ORDER BY <that key> [ASC] is the input itself
    // (the reference sorts regardless, src/exec/operator.cpp:1115; its result is this order - distinct keys leave no ties)
    if (in->ordered_by_first && key_cols.size() == 1 && key_cols[0] == 0 && asc[0] == 1 && rel_cols.size() == in->cols.size() &&
        (limit < 0 || static_cast<size_t>(limit) >= in->rows)) {
        auto same = std::make_shared<DeviceRelation>(*in);
        return same;
    }
    std::vector<bq_col*> hs;
    for (auto& c : rel_cols) hs.push_back(c->h);
    bq_rel* shell = nullptr;
    check(bq_rel_create(ctx, hs.data(), static_cast<int>(hs.size()), &shell));
    bq_rel* sorted = nullptr;
    int rc = bq_rel_sort(ctx, shell, static_cast<int>(key_cols.size()), key_cols.data(), asc.data(), limit, &sorted);
    std::vector<bq_col*> back(hs.size());
    bq_rel_release(shell, back.data());       // the inputs stay owned by their DevCols
    check(rc);
    DeviceRelationPtr out = gpu::relation_from(sorted);
    out->cols.resize(in->cols.size());        // drop evaluated sort keys
    return out;
}

}  // namespace bosql
