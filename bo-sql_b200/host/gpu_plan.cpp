// gpu_plan.cpp — the host side of the fused pipelines: slot assignment, join-table choice from catalog statistics,
// grouping strategy, and the calls into include/bosql_b200.h.
//
// What the reference does per row at run time (HashAggregate::next, src/exec/operator.cpp:984-1014;
// Selection::next :403-429; HashJoin::open/next :739-837) is decided here once per query.
#include "gpu_plan.hpp"
#include "exchange.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <map>

namespace bosql::gpu {

ColumnLookup lookup_for(const std::vector<PipeCol>& cols) {
    const std::vector<PipeCol>* c = &cols;
    ColumnLookup l;
    // the reference's name->index map is filled in column order, so the LAST duplicate name wins
    l.index_of = [c](const std::string& name) {
        for (int i = static_cast<int>(c->size()) - 1; i >= 0; --i)
            if ((*c)[i].name == name) return i;
        return -1;
    };
    l.type_of = [c](int i) { return (*c)[i].type; };
    return l;
}

ColumnLookup lookup_for(const std::vector<std::string>& names, const std::vector<TypeId>& types) {
    const auto* n = &names;
    const auto* t = &types;
    ColumnLookup l;
    l.index_of = [n](const std::string& name) {
        for (int i = static_cast<int>(n->size()) - 1; i >= 0; --i)
            if ((*n)[i] == name) return i;
        return -1;
    };
    l.type_of = [t](int i) { return (*t)[i]; };
    return l;
}

DeviceRelationPtr empty_relation(const std::vector<TypeId>& types) {
    auto rel = std::make_shared<DeviceRelation>();
    for (TypeId t : types) {
        bq_col* h = nullptr;
        check(bq_col_alloc(context(), static_cast<int>(t), 0, &h));
        rel->cols.push_back(adopt(h));
    }
    rel->rows = 0;
    return rel;
}

DeviceRelationPtr concat_relations(const std::vector<DeviceRelationPtr>& parts, const std::vector<TypeId>& types) {
    auto out = std::make_shared<DeviceRelation>();
    for (const auto& p : parts) out->rows += p->rows;
    for (size_t c = 0; c < types.size(); ++c) {
        bq_col* h = nullptr;
        check(bq_col_alloc(context(), static_cast<int>(types[c]), out->rows, &h));
        DevColPtr col = adopt(h);
        const size_t w = type_width(types[c]);
        size_t at = 0;
        for (const auto& p : parts) {
            if (!p->rows) continue;
            check(bq_copy_bytes(context(), static_cast<char*>(bq_col_ptr(h)) + at * w, bq_col_ptr(p->cols[c]->h), p->rows * w));
            at += p->rows;
        }
        out->cols.push_back(col);
    }
    return out;
}

void resolve_stats(PipeCol& col, bool force_device) {
    if (col.stats.known && !force_device) return;
    int64_t lo = 0, hi = -1;
    if (force_device) bq_col_invalidate_stats(col.dev->h);
    check(bq_col_minmax(context(), col.dev->h, &lo, &hi));
    col.stats.known = true;
    col.stats.measured = true;
    col.stats.min_key = lo;
    col.stats.max_key = hi;
}

// An expression program fails only through an integer division by zero in the rows it is given (src/exec/expression.cpp:52).
// Across GPUs that is a fact about ONE rank's shard, so programs containing such a division exchange their outcome: either
// every rank continues or every rank throws (none is left alone inside the next collective).  Whether a program divides is
// a property of the plan, so all ranks agree on whether the exchange happens - also a rank whose shard is empty.
constexpr bool kGroupTablesDefault = false;

static void run_program(const std::function<void()>& launch, const std::vector<bq_insn>& code) {
    bool divides = false;
    for (const bq_insn& in : code) divides = divides || in.op == BQ_OP_DIV_I;
    if (divides && exchange().active) agree_on(launch);
    else launch();
}

DevColPtr eval_to_column(const Expr* e, const std::vector<PipeCol>& cols, size_t rows, Dictionary* dict, TypeId out_type,
                         bool as_predicate) {
    Program prog = compile(e, lookup_for(cols), dict, as_predicate);
    std::vector<const bq_col*> cs;
    for (int idx : prog.columns) cs.push_back(cols[idx].dev->h);
    bq_col* out = nullptr;
    const bq_col* none = nullptr;
    run_program([&] {
        check(bq_eval(context(), prog.code.data(), static_cast<int>(prog.code.size()), cs.empty() ? &none : cs.data(),
                      static_cast<int>(cs.size()), 0, rows, static_cast<int>(out_type), &out));
    }, prog.code);
    return adopt(out);
}

// Runs a hand-assembled program (no Expr typing rules: used for packing several integer keys into one).
static DevColPtr eval_program(const std::vector<bq_insn>& code, const std::vector<DevColPtr>& cols, size_t rows, TypeId out_type) {
    std::vector<const bq_col*> cs;
    for (const auto& c : cols) cs.push_back(c->h);
    bq_col* out = nullptr;
    run_program([&] {
        check(bq_eval(context(), code.data(), static_cast<int>(code.size()), cs.data(), static_cast<int>(cs.size()), 0, rows,
                      static_cast<int>(out_type), &out));
    }, code);
    return adopt(out);
}
static bq_insn insn(int op, int arg = 0, int64_t imm = 0) {
    bq_insn in{};
    in.op = op;
    in.arg = arg;
    in.imm.i = imm;
    return in;
}

bq_slot make_slot(const DevColPtr& col, const std::vector<bq_range>& ranges, bool from_build) {
    bq_slot s{};
    s.col = col ? col->h : nullptr;
    s.n_ranges = static_cast<int32_t>(ranges.size());
    s.from_build = from_build ? 1 : 0;
    for (size_t i = 0; i < ranges.size() && i < 2; ++i) s.r[i] = ranges[i];
    return s;
}

namespace {

// Intersect all positive ranges into one; keep negated ranges as they are.
std::vector<bq_range> normalise(const std::vector<bq_range>& in) {
    bool have_pos = false;
    bq_range pos{};
    std::vector<bq_range> out;
    for (const auto& r : in) {
        if (r.neg) {
            out.push_back(r);
        } else if (!have_pos) {
            pos = r;
            have_pos = true;
        } else {
            pos.lo = std::max(pos.lo, r.lo);
            pos.hi = std::min(pos.hi, r.hi);
        }
    }
    if (have_pos) out.insert(out.begin(), pos);
    return out;
}

std::unique_ptr<Expr> and_of(const std::vector<const Expr*>& es) {
    std::unique_ptr<Expr> acc;
    for (const Expr* e : es) {
        if (!acc) {
            acc = e->clone();
            continue;
        }
        auto n = std::make_unique<Expr>();
        n->type = ExprType::BINARY_OP;
        n->op = BinaryOp::AND;
        n->left = std::move(acc);
        n->right = e->clone();
        acc = std::move(n);
    }
    return acc;
}

}  // namespace

SlotPlan plan_slots(const std::vector<const Conjunct*>& conjuncts, const std::vector<PipeCol>& cols, size_t rows,
                    const std::vector<int>& role_cols, int n_pred_slots) {
    SlotPlan plan;
    plan.role_ranges.resize(role_cols.size());
    const ColumnLookup look = lookup_for(cols);

    struct PerCol {
        std::vector<bq_range> ranges;
        std::vector<const Conjunct*> sources;
    };
    std::map<int, PerCol> by_col;
    std::vector<const Conjunct*> leftovers;
    for (const Conjunct* c : conjuncts) {
        ColumnRange cr;
        if (to_range(c->expr.get(), look, c->dict, cr)) {
            by_col[cr.column].ranges.push_back(cr.range);
            by_col[cr.column].sources.push_back(c);
        } else {
            leftovers.push_back(c);
        }
    }
    for (auto& [col, pc] : by_col) {
        std::vector<bq_range> norm = normalise(pc.ranges);
        bool placed = false;
        if (norm.size() <= 2) {
            for (size_t r = 0; r < role_cols.size() && !placed; ++r) {
                // the same device column may sit in a role slot under another index (e.g. a name bound twice)
                if (role_cols[r] >= 0 && cols[role_cols[r]].dev->h == cols[col].dev->h && plan.role_ranges[r].empty()) {
                    plan.role_ranges[r] = norm;
                    placed = true;
                }
            }
            if (!placed && static_cast<int>(plan.pred.size()) < n_pred_slots) {
                plan.pred.push_back({col, norm});
                placed = true;
            }
        }
        if (!placed) leftovers.insert(leftovers.end(), pc.sources.begin(), pc.sources.end());
    }
    if (!leftovers.empty()) {
        // one program: c1 AND c2 AND ...  (every conjunct still evaluated for every row, H11)
        std::vector<const Expr*> es;
        Dictionary* dict = nullptr;
        for (const Conjunct* c : leftovers) {
            es.push_back(c->expr.get());
            if (c->dict) dict = c->dict;
        }
        auto all = and_of(es);
        plan.mask = eval_to_column(all.get(), cols, rows, dict, TypeId::INT64, true);
    }
    return plan;
}

DevColPtr select_rowids(const std::vector<PipeCol>& cols, size_t rows, const std::vector<const Conjunct*>& conjuncts) {
    SlotPlan plan = plan_slots(conjuncts, cols, rows, {}, 4);
    bq_select_spec spec{};
    for (size_t i = 0; i < plan.pred.size(); ++i) spec.pred[i] = make_slot(cols[plan.pred[i].first].dev, plan.pred[i].second);
    spec.mask = plan.mask ? plan.mask->h : nullptr;
    spec.row_begin = 0;
    spec.row_end = rows;
    bq_col* ids = nullptr;
    check(bq_select(context(), &spec, &ids));
    return adopt(ids);
}

DeviceRelationPtr gather_rows(const std::vector<PipeCol>& cols, const DevColPtr& rowids) {
    auto out = std::make_shared<DeviceRelation>();
    out->rows = rowids->rows();
    for (const auto& c : cols) {
        bq_col* g = nullptr;
        check(bq_gather(context(), c.dev->h, rowids->h, &g));
        out->cols.push_back(adopt(g));
    }
    return out;
}

DeviceRelationPtr run_selection(const std::vector<PipeCol>& cols, size_t rows, const std::vector<const Conjunct*>& conjuncts) {
    if (rows == 0) {
        // across GPUs an empty shard still takes part in the outcome exchange of a predicate program
        if (exchange().active) plan_slots(conjuncts, cols, 0, {}, 4);
        std::vector<TypeId> types;
        for (const auto& c : cols) types.push_back(c.type);
        return empty_relation(types);
    }
    return gather_rows(cols, select_rowids(cols, rows, conjuncts));
}

// ---- aggregate ------------------------------------------------------------------------------------------
namespace {

bool is_int_literal(const Expr* e) { return e->type == ExprType::LITERAL_INT; }
bool is_arith_op(BinaryOp op) { return op == BinaryOp::ADD || op == BinaryOp::SUB || op == BinaryOp::MUL || op == BinaryOp::DIV; }

int vop_of(BinaryOp op) {
    switch (op) {
        case BinaryOp::MUL: return BQ_V_MUL;
        case BinaryOp::ADD: return BQ_V_ADD;
        case BinaryOp::SUB: return BQ_V_SUB;
        default: return BQ_V_DIV;
    }
}

// One aggregate argument in the fused kernel's vocabulary: a column, or left OP right where each operand is a
// numeric column or an integer literal.
struct ValueForm {
    bool unary = true;
    int op = BQ_V_A;                // BQ_V_MUL.. when binary
    int col_l = -1, col_r = -1;     // pipeline column of each operand (-1 = immediate / unused)
    int64_t imm = 0;
    TypeId type = TypeId::INT64;    // static type of the argument (decides SUM's result type)
};

bool numeric_col(const std::vector<PipeCol>& cols, int idx) {
    return cols[idx].type == TypeId::INT64 || cols[idx].type == TypeId::DOUBLE;
}

// false = needs a derived column (evaluated by bq_eval first)
bool simple_value_form(const Expr* arg, const std::vector<PipeCol>& cols, const ColumnLookup& look, ValueForm& out) {
    if (arg->type == ExprType::COLUMN_REF) {
        int idx = look.index_of(arg->str_val);
        if (idx < 0) throw std::runtime_error("Unknown column: " + arg->str_val);
        out.unary = true;
        out.col_l = idx;
        out.type = cols[idx].type;
        return true;
    }
    if (arg->type != ExprType::BINARY_OP || !is_arith_op(arg->op)) return false;
    auto col_of = [&](const Expr* e) {
        if (e->type != ExprType::COLUMN_REF) return -1;
        int idx = look.index_of(e->str_val);
        if (idx < 0) throw std::runtime_error("Unknown column: " + e->str_val);
        return numeric_col(cols, idx) ? idx : -1;
    };
    const Expr* l = arg->left.get();
    const Expr* r = arg->right.get();
    const int lc = col_of(l), rc = col_of(r);
    const bool l_ok = lc >= 0 || is_int_literal(l), r_ok = rc >= 0 || is_int_literal(r);
    if (!l_ok || !r_ok || (lc < 0 && rc < 0)) return false;
    out.unary = false;
    out.op = vop_of(arg->op);
    out.col_l = lc;
    out.col_r = rc;
    out.imm = lc < 0 ? l->i64_val : (rc < 0 ? r->i64_val : 0);
    const bool fp = (lc >= 0 && cols[lc].type == TypeId::DOUBLE) || (rc >= 0 && cols[rc].type == TypeId::DOUBLE);
    out.type = fp ? TypeId::DOUBLE : TypeId::INT64;
    return true;
}

struct Domain {
    int64_t lo = 0, hi = -1;
    uint64_t size() const { return hi < lo ? 0 : static_cast<uint64_t>(hi - lo) + 1; }
};

}  // namespace

// Heavy hitters of a (probe-side) join key, agreed by all ranks: every rank counts the keys of a sample of its rows
// (the fused GROUP BY kernel over the first 2^20 rows, top 16 by count), the candidates are pooled on the host, and a key
// is hot when its estimated share of all rows exceeds 1 / (8 * world) - enough to overload the rank that would own it.
// BOSQL_GROUP_TABLES=0|1 overrides the default choice between the shared-memory group tables (bq_partition_aggregate) and
// the L2-resident table (bq_scan_aggregate) for a high-cardinality GROUP BY over plain column arguments.
static bool group_tables_enabled() {
    const char* e = std::getenv("BOSQL_GROUP_TABLES");
    return e ? *e != '0' : kGroupTablesDefault;
}

static std::vector<int64_t> find_hot_keys(PipeCol& key, size_t rows) {
    Exchange& xch = exchange();
    bq_ctx* ctx = context();
    constexpr size_t kSample = 1u << 20;
    constexpr int kTop = 16;
    const size_t sample = std::min(rows, kSample);
    std::vector<int64_t> mine(2 * kTop + 1, 0);          // keys, counts, sample size
    mine[2 * kTop] = static_cast<int64_t>(sample);
    if (sample) {
        bq_scan_spec s{};
        s.key = make_slot(key.dev, {});
        s.row_begin = 0;
        s.row_end = sample;
        s.group_mode = BQ_GROUP_HASH;
        s.ndv_hint = sample;
        s.n_out = 1;
        s.out[0].func = BQ_AGG_COUNT;
        bq_rel* rel = nullptr;
        check(bq_scan_aggregate(ctx, &s, &rel));
        DeviceRelationPtr counted = relation_from(rel);
        if (counted->rows) {
            std::vector<bq_col*> hs = {counted->cols[0]->h, counted->cols[1]->h};
            bq_rel* shell = nullptr;
            check(bq_rel_create(ctx, hs.data(), 2, &shell));
            int by = 1, asc = 0;
            bq_rel* top = nullptr;
            int rc = bq_rel_sort(ctx, shell, 1, &by, &asc, kTop, &top);
            std::vector<bq_col*> back(2);
            bq_rel_release(shell, back.data());
            check(rc);
            DeviceRelationPtr t = relation_from(top);
            std::vector<int64_t> counts(t->rows);
            check(bq_col_read(ctx, t->cols[1]->h, 0, t->rows, counts.data()));
            if (key.type == TypeId::INT64) {
                std::vector<int64_t> keys(t->rows);
                check(bq_col_read(ctx, t->cols[0]->h, 0, t->rows, keys.data()));
                for (size_t i = 0; i < t->rows; ++i) mine[i] = keys[i];
            } else {
                std::vector<int32_t> keys(t->rows);                 // DATE32 / STRING ids: 4-byte keys, widened as the kernels do
                check(bq_col_read(ctx, t->cols[0]->h, 0, t->rows, keys.data()));
                for (size_t i = 0; i < t->rows; ++i)
                    mine[i] = key.type == TypeId::STRING ? static_cast<int64_t>(static_cast<uint32_t>(keys[i])) : static_cast<int64_t>(keys[i]);
            }
            for (size_t i = 0; i < t->rows; ++i) mine[kTop + i] = counts[i];
        }
    }
    auto all = xch.host_gather(mine);
    const size_t stride = mine.size();
    std::map<int64_t, int64_t> pooled;
    int64_t sampled = 0;
    for (int r = 0; r < xch.world(); ++r) {
        sampled += all[r * stride + 2 * kTop];
        for (int i = 0; i < kTop; ++i)
            if (all[r * stride + kTop + i] > 0) pooled[all[r * stride + i]] += all[r * stride + kTop + i];
    }
    std::vector<std::pair<int64_t, int64_t>> ranked;       // (count, key)
    for (const auto& kv : pooled)
        if (sampled > 0 && kv.second * 8 * xch.world() > sampled) ranked.emplace_back(kv.second, kv.first);
    std::sort(ranked.rbegin(), ranked.rend());
    std::vector<int64_t> hot;
    for (size_t i = 0; i < ranked.size() && i < static_cast<size_t>(kTop); ++i) hot.push_back(ranked[i].second);
    std::sort(hot.begin(), hot.end());                      // the same list, in the same order, on every rank
    return hot;
}

DeviceRelationPtr run_aggregate(Pipeline& p, const AggRequest& req) {
    bq_ctx* ctx = context();
    std::vector<TypeId> out_types = req.group_types;
    for (const auto& a : req.aggs) out_types.push_back(a.result_type);

    // Multi-GPU: every rank holds a row shard and must reach the same exchange points, so nothing below may leave early
    // on a condition only this rank sees (an empty shard still takes part in the collectives).
    Exchange& xch = exchange();
    const bool dist = xch.active;
    // Row counts over all ranks, fetched with ONE host exchange the first time a decision needs them (every rank reaches
    // that point together: the decisions that lead there depend only on catalog statistics and on these counts).  They
    // stay valid when rows are shuffled between ranks later on - a shuffle moves rows, it does not change their number.
    struct GlobalRows {
        bool have = false;
        uint64_t probe = 0, build = 0;
        std::vector<int64_t> build_by_rank;
    } global_rows;
    auto fetch_global_rows = [&]() -> const GlobalRows& {
        if (!global_rows.have) {
            std::vector<int64_t> all;
            RowCountCache* rc = req.row_cache;
            if (rc && rc->have && rc->local_probe == static_cast<int64_t>(p.rows) && rc->local_build == static_cast<int64_t>(p.build_rows)) {
                all = rc->all;                    // every rank ran this plan before, over the same tables: same answer, no exchange
            } else {
                all = xch.host_gather({static_cast<int64_t>(p.rows), static_cast<int64_t>(p.build_rows)});
                if (rc) *rc = RowCountCache{true, static_cast<int64_t>(p.rows), static_cast<int64_t>(p.build_rows), all};
            }
            for (int r = 0; r < xch.world(); ++r) {
                global_rows.probe += static_cast<uint64_t>(all[2 * static_cast<size_t>(r)]);
                global_rows.build += static_cast<uint64_t>(all[2 * static_cast<size_t>(r) + 1]);
                global_rows.build_by_rank.push_back(all[2 * static_cast<size_t>(r) + 1]);
            }
            global_rows.have = true;
        }
        return global_rows;
    };
    auto not_fusable = [&](const char* why) -> DeviceRelationPtr {
        if (dist) throw std::runtime_error(std::string("not supported across GPUs: ") + why);
        return nullptr;
    };

    // zero input rows -> zero output rows, even for a global aggregate (src/exec/operator.cpp:990-993, H5)
    if (!dist && (p.rows == 0 || (p.joined && p.build_rows == 0))) return empty_relation(out_types);
    if (p.cross_join) return not_fusable("a join whose ON clause is not column = column");

    ColumnLookup look = lookup_for(p.cols);

    // The reference evaluates aggregate arguments and GROUP BY expressions only for rows that survived the Selection and the
    // join, and a WHERE above a join only for joined rows (src/logical/planner.cpp builds Filter above HashJoin).  The fused
    // pipeline evaluates derived columns and predicate programs over ALL probe rows - harmless unless the expression can
    // throw for a row the reference never looks at (integer division by zero, H10).  Those plans are materialised step by
    // step instead (selection / join first, then the same kernels over the surviving rows).
    {
        const bool filtered = !p.conjuncts.empty() || p.joined;
        bool risky = false;
        if (filtered) {
            for (const auto& g : *req.group_exprs)
                if (g->type != ExprType::COLUMN_REF && may_throw_per_row(g.get(), look)) risky = true;
            for (const auto& a : req.aggs) {
                ValueForm form;
                if (a.arg && a.func != "COUNT" && may_throw_per_row(a.arg, look) && !simple_value_form(a.arg, p.cols, look, form)) risky = true;
            }
        }
        if (p.joined)
            for (const Conjunct& c : p.conjuncts)
                if (may_throw_per_row(c.expr.get(), look)) risky = true;
        if (risky) return nullptr;
    }

    // ---- split predicates by the side they read ---------------------------------------------------------
    std::vector<const Conjunct*> probe_conj, build_conj;
    for (const Conjunct& c : p.conjuncts) {
        std::vector<int> refs;
        referenced(c.expr.get(), look, refs);
        bool any_probe = false, any_build = false;
        for (int r : refs) (p.cols[r].side ? any_build : any_probe) = true;
        if (any_probe && any_build) return not_fusable("a predicate comparing columns of both join sides");   // materialise the join
        (any_build ? build_conj : probe_conj).push_back(&c);
    }

    // the columns a derived expression may read: everything on the probe side (same length, streamed)
    auto all_on_probe = [&](const Expr* e) {
        std::vector<int> refs;
        referenced(e, look, refs);
        for (int r : refs)
            if (p.cols[r].side) return false;
        return true;
    };

    // ---- group key -----------------------------------------------------------------------------------------
    const auto& gexprs = *req.group_exprs;
    int key_col = -1;
    std::vector<Domain> packed_domains;      // multi-column keys packed into one int64
    std::vector<uint64_t> packed_strides;
    if (gexprs.size() == 1 && gexprs[0]->type == ExprType::COLUMN_REF) {
        key_col = look.index_of(gexprs[0]->str_val);
        if (key_col < 0) throw std::runtime_error("Unknown column: " + gexprs[0]->str_val);
    } else if (gexprs.size() == 1) {
        if (!all_on_probe(gexprs[0].get())) return not_fusable("a GROUP BY expression over build-side columns");
        PipeCol derived;
        derived.name = "\x01group1";
        derived.type = req.group_types[0];
        derived.dev = eval_to_column(gexprs[0].get(), p.cols, p.rows, req.dict, derived.type, false);
        p.cols.push_back(std::move(derived));
        key_col = static_cast<int>(p.cols.size()) - 1;
        look = lookup_for(p.cols);
    } else if (gexprs.size() > 1) {
        // GROUP BY a, b, ...: pack (a-min_a, b-min_b, ...) into one integer, most significant first
        uint64_t total = 1;
        std::vector<int> key_cols;
        for (size_t gi = 0; gi < gexprs.size(); ++gi) {
            const auto& g = gexprs[gi];
            int idx = -1;
            if (g->type == ExprType::COLUMN_REF) {
                idx = look.index_of(g->str_val);
                if (idx < 0) throw std::runtime_error("Unknown column: " + g->str_val);
            } else {
                // evaluate_key_row evaluates every group expression per row (src/exec/operator.cpp:972-982): a derived column
                if (!all_on_probe(g.get())) return not_fusable("a GROUP BY expression over build-side columns");
                PipeCol derived;
                derived.name = "\x01group" + std::to_string(gi + 1);
                derived.type = req.group_types[gi];
                derived.dev = eval_to_column(g.get(), p.cols, p.rows, req.dict, derived.type, false);
                p.cols.push_back(std::move(derived));
                idx = static_cast<int>(p.cols.size()) - 1;
            }
            if (p.cols[idx].type == TypeId::DOUBLE) throw std::runtime_error("GROUP BY over several keys needs integer-typed keys on the GPU path");
            if (p.cols[idx].side) return not_fusable("several GROUP BY keys from the build side");
            key_cols.push_back(idx);
            // the packing is only injective over the TRUE value range: measure it (cached in the column handle) rather
            // than trust catalog bounds, which nothing downstream could catch if they were stale
            p.cols[idx].stats.known = false;
            resolve_stats(p.cols[idx]);
            if (dist) xch.minmax(p.cols[idx].stats.min_key, p.cols[idx].stats.max_key);
            Domain d{p.cols[idx].stats.min_key, p.cols[idx].stats.max_key};
            if (d.size() == 0) d = Domain{0, 0};
            if (total > (1ULL << 62) / d.size()) throw std::runtime_error("GROUP BY key domain too large to pack");
            total *= d.size();
            packed_domains.push_back(d);
        }
        if (key_cols.size() > BQ_MAX_PROGRAM_COLS) throw std::runtime_error("too many GROUP BY keys");
        std::vector<bq_insn> code;
        std::vector<DevColPtr> kcols;
        uint64_t stride = total;
        for (size_t i = 0; i < key_cols.size(); ++i) {
            stride /= packed_domains[i].size();
            packed_strides.push_back(stride);
            kcols.push_back(p.cols[key_cols[i]].dev);
            code.push_back(insn(BQ_OP_COL, static_cast<int>(i)));
            code.push_back(insn(BQ_OP_IMM_I, 0, packed_domains[i].lo));
            code.push_back(insn(BQ_OP_SUB_I));
            code.push_back(insn(BQ_OP_IMM_I, 0, static_cast<int64_t>(stride)));
            code.push_back(insn(BQ_OP_MUL_I));
            if (i) code.push_back(insn(BQ_OP_ADD_I));
        }
        PipeCol derived;
        derived.name = "\x01packed";
        derived.type = TypeId::INT64;
        derived.dev = eval_program(code, kcols, p.rows, TypeId::INT64);
        derived.stats.known = true;
        derived.stats.min_key = 0;
        derived.stats.max_key = static_cast<int64_t>(total - 1);
        p.cols.push_back(std::move(derived));
        key_col = static_cast<int>(p.cols.size()) - 1;
        look = lookup_for(p.cols);
    }

    // ---- aggregate arguments -----------------------------------------------------------------------------------
    struct Value {
        std::string text;
        ValueForm form;
    };
    std::vector<Value> values;                 // distinct arguments
    std::vector<int> agg_value(req.aggs.size(), -1);
    for (size_t i = 0; i < req.aggs.size(); ++i) {
        const auto& a = req.aggs[i];
        if (a.func == "COUNT") continue;
        if (a.func != "SUM" && a.func != "AVG") throw std::runtime_error("unsupported aggregate function: " + a.func);
        std::string text = a.arg->to_string();
        int found = -1;
        for (size_t v = 0; v < values.size(); ++v)
            if (values[v].text == text) found = static_cast<int>(v);
        if (found < 0) {
            Value val;
            val.text = text;
            if (!simple_value_form(a.arg, p.cols, look, val.form)) {
                if (!all_on_probe(a.arg)) return not_fusable("an aggregate argument mixing both join sides");
                PipeCol derived;
                derived.name = "\x01value" + std::to_string(values.size());
                derived.type = value_type(a.arg, look);
                derived.dev = eval_to_column(a.arg, p.cols, p.rows, req.dict, derived.type, false);
                p.cols.push_back(std::move(derived));
                look = lookup_for(p.cols);
                val.form = ValueForm{};
                val.form.col_l = static_cast<int>(p.cols.size()) - 1;
                val.form.type = p.cols.back().type;
            }
            values.push_back(std::move(val));
            found = static_cast<int>(values.size()) - 1;
        }
        agg_value[i] = found;
    }

    // ---- passes: each holds <= 2 arguments over <= 2 columns ---------------------------------------------------
    struct Pass {
        int col_a = -1, col_b = -1;
        std::vector<int> vals;     // indices into values
    };
    std::vector<Pass> passes;
    for (size_t v = 0; v < values.size(); ++v) {
        const ValueForm& f = values[v].form;
        bool placed = false;
        for (Pass& ps : passes) {
            if (ps.vals.size() >= 2) continue;
            // try to fit f's columns into the pass's (A, B)
            int a = ps.col_a, b = ps.col_b;
            auto add = [&](int c) {
                if (c < 0 || c == a || c == b) return true;
                if (a < 0) { a = c; return true; }
                if (b < 0) { b = c; return true; }
                return false;
            };
            if (add(f.col_l) && add(f.col_r)) {
                ps.col_a = a;
                ps.col_b = b;
                ps.vals.push_back(static_cast<int>(v));
                placed = true;
                break;
            }
        }
        if (!placed) {
            Pass ps;
            ps.col_a = f.col_l >= 0 ? f.col_l : f.col_r;
            ps.col_b = (f.col_r >= 0 && f.col_r != ps.col_a) ? f.col_r : -1;
            ps.vals.push_back(static_cast<int>(v));
            passes.push_back(ps);
        }
    }
    if (passes.empty()) passes.push_back(Pass{});      // COUNT only

    // ---- join build ------------------------------------------------------------------------------------------------
    bq_join* join = nullptr;
    struct JoinGuard {
        bq_join*& j;
        ~JoinGuard() { if (j) bq_join_free(nullptr, j); }
    } join_guard{join};
    if (p.joined) for (int attempt = 0;; ++attempt) {
        PipeCol& bk = p.cols[p.build_key];
        const PipeCol& pk = p.cols[p.probe_key];
        if (bk.type != pk.type) return empty_relation(out_types);     // KeyEqual: different TypeId never match (:652)
        bool need_rows = key_col >= 0 && p.cols[key_col].side;
        for (const auto& v : values)
            for (int c : {v.form.col_l, v.form.col_r})
                if (c >= 0 && p.cols[c].side) need_rows = true;
        // Across GPUs the build side is sharded too.  A semi-join over unique, dense keys needs only its bitmap: each
        // rank sets the bits of its own build rows over the GLOBAL key domain and the words are summed (= OR, keys
        // being unique) - the probe side never moves.  Anything else is a broadcast join: the build columns are
        // all-gathered and every rank builds the full table.
        bool dist_bitmap = false;
        uint64_t shuffled_global_build = 0;      // > 0: both sides were co-partitioned; the key domain is still the global one
        if (dist) {
            const std::vector<int64_t> rows_by_rank = fetch_global_rows().build_by_rank;
            const uint64_t global_build = global_rows.build;
            if (bk.type != TypeId::DOUBLE && !need_rows) {
                resolve_stats(bk);
                if (bk.stats.measured) xch.minmax(bk.stats.min_key, bk.stats.max_key);
                const uint64_t dom = bk.stats.max_key >= bk.stats.min_key ? static_cast<uint64_t>(bk.stats.max_key - bk.stats.min_key) + 1 : 0;
                dist_bitmap = attempt == 0 && dom > 0 && dom <= (1ULL << 32) && dom <= 8 * global_build + 1024 && bk.stats.ndv && bk.stats.ndv == global_build;
            }
            // Neither: either broadcast the build side, or CO-PARTITION both sides by hash(join key) so that every rank joins
            // the keys it owns.  The choice is bytes over NVLink; skewed probe keys are handled by keeping the
            // heavy hitters' probe rows where they are and replicating their (few) build rows to every rank.
            bool shuffled_join = false;
            shuffled_global_build = 0;
            if (!dist_bitmap) {
                PipeCol& pkc = p.cols[p.probe_key];
                std::vector<int> probe_refs, build_refs;          // columns read downstream, besides the two keys
                auto note = [&](int c) {
                    if (c < 0 || c == p.probe_key || c == p.build_key) return;
                    auto& list = p.cols[c].side ? build_refs : probe_refs;
                    if (std::find(list.begin(), list.end(), c) == list.end()) list.push_back(c);
                };
                note(key_col);
                for (const auto& v : values) {
                    note(v.form.col_l);
                    note(v.form.col_r);
                }
                auto row_bytes = [&](const std::vector<int>& refs) {
                    uint64_t b = 8;
                    for (int c : refs) b += (p.cols[c].type == TypeId::INT64 || p.cols[c].type == TypeId::DOUBLE) ? 8 : 4;
                    return b;
                };
                const uint64_t W = static_cast<uint64_t>(xch.world());
                const uint64_t global_probe = fetch_global_rows().probe;
                const uint64_t broadcast_in = global_build * row_bytes(build_refs) * (W - 1) / W;                  // per rank
                const uint64_t shuffle_out = (global_probe * row_bytes(probe_refs) + global_build * row_bytes(build_refs)) / W * (W - 1) / W;
                const char* force = std::getenv("BOSQL_JOIN");
                const bool possible = bk.type != TypeId::DOUBLE && pkc.type != TypeId::DOUBLE && probe_conj.empty() && build_conj.empty() &&
                                      probe_refs.size() <= 2 && build_refs.size() <= 2 && !(std::getenv("BOSQL_SHUFFLE") && std::string(std::getenv("BOSQL_SHUFFLE")) == "collective");
                // measured on configuration 5, 8 GPUs (profiles/README.md): at 0.63 of the broadcast's bytes the shuffle wins by
                // 7 % (13.7 vs 14.7 ms) - its two partition passes, heavy-hitter detection and size/handle exchanges eat most
                // of the saving - so it is chosen from 0.7 down
                bool want = shuffle_out * 10 < broadcast_in * 7;
                if (force && std::string(force) == "shuffle") want = true;
                if (force && std::string(force) == "broadcast") want = false;
                if (possible && want) {
                    std::vector<int64_t> hot = find_hot_keys(pkc, p.rows);
                    auto devs = [&](const std::vector<int>& refs) {
                        std::vector<DevColPtr> d;
                        for (int c : refs) d.push_back(p.cols[c].dev);
                        return d;
                    };
                    Shuffled sp = shuffle_by_key(pkc.dev, devs(probe_refs), p.rows, hot, true);
                    pkc.dev = sp.key;
                    for (size_t i = 0; i < probe_refs.size(); ++i) p.cols[probe_refs[i]].dev = sp.payload[i];
                    p.rows = sp.rows;
                    Shuffled sb = shuffle_by_key(bk.dev, devs(build_refs), p.build_rows, hot, false);
                    bk.dev = sb.key;
                    for (size_t i = 0; i < build_refs.size(); ++i) p.cols[build_refs[i]].dev = sb.payload[i];
                    p.build_rows = sb.rows;
                    // bounds from the catalog still hold for the keys a rank owns; measured (shard) bounds do not
                    if (bk.stats.measured) bk.stats.known = false;
                    shuffled_join = true;
                    shuffled_global_build = global_build;
                }
            }
            if (!dist_bitmap && !shuffled_join) {
                for (auto& c : p.cols)
                    if (c.side) c.dev = xch.all_gather_column(c.dev, p.build_rows, rows_by_rank);
                p.build_rows = static_cast<size_t>(global_build);
                if (bk.stats.measured) bk.stats.known = false;        // shard-local bounds no longer describe the column
            }
        }
        // build-side columns as a column set of their own (length build_rows)
        std::vector<PipeCol> bcols;
        std::vector<int> bmap(p.cols.size(), -1);
        for (size_t i = 0; i < p.cols.size(); ++i)
            if (p.cols[i].side) {
                bmap[i] = static_cast<int>(bcols.size());
                bcols.push_back(p.cols[i]);
            }
        SlotPlan bplan;
        bplan = plan_slots(build_conj, bcols, p.build_rows, {}, 3);
        bq_join_spec js{};
        js.key = bk.dev->h;
        for (size_t i = 0; i < bplan.pred.size(); ++i) js.pred[i] = make_slot(bcols[bplan.pred[i].first].dev, bplan.pred[i].second);
        js.mask = bplan.mask ? bplan.mask->h : nullptr;
        js.row_begin = 0;
        js.row_end = p.build_rows;
        js.kind = dist_bitmap ? BQ_JOIN_BITMAP : BQ_JOIN_AUTO;
        js.need_rows = need_rows ? 1 : 0;
        if (bk.type != TypeId::DOUBLE) {
            resolve_stats(bk);
            js.key_min = bk.stats.min_key;
            js.key_max = bk.stats.max_key;
            js.unique = (bk.stats.ndv && bk.stats.ndv == p.build_rows) ? 1 : 0;
        } else {
            js.kind = BQ_JOIN_HASH;
        }
        if (shuffled_global_build && js.kind == BQ_JOIN_AUTO && js.key_max >= js.key_min) {
            // a rank owns 1/world of a domain that is dense GLOBALLY: judge the density on the whole build side, or ranks
            // just under the threshold would fall back to a hash table (measured: 27 ms instead of 2 on configuration 5)
            const uint64_t dom = static_cast<uint64_t>(js.key_max - js.key_min) + 1;
            if (dom <= (1ULL << 32) && dom <= 8 * shuffled_global_build + 1024) js.kind = need_rows ? BQ_JOIN_DIRECT : BQ_JOIN_BITMAP;
        }
        if (dist_bitmap) {
            // Each rank sets the bits of its own build rows over the GLOBAL key domain; nothing is read back.  The build
            // counters ride behind the bitmap words through the same all-reduce, so the merged bitmap arrives together with
            // the number of rows all ranks inserted: ONE collective and one host round trip for the whole distributed build.
            // The sum of the ranks' words equals their OR only while no key was inserted twice - by one rank (a shard with
            // duplicates) or by two (statistics that call the key unique may be stale, or a dimension table may be partly
            // replicated): a doubly set bit would carry into its neighbour.  The merged bitmap must therefore hold exactly one
            // bit per inserted row; if it does not (or a rank met a key outside the catalog's bounds), every rank sees the same
            // numbers and the join is redone as a broadcast join.
            check(bq_join_build_bitmap_nosync(ctx, &js, &join));
            size_t words = 0;
            void* bits = bq_join_bitmap_ptr(join, &words);
            xch.sum_words(bits, words + BQ_JOIN_TRAILER_WORDS);
            uint64_t set_bits = 0, inserted = 0;
            int build_flags = 0;
            check(bq_join_bitmap_verdict(ctx, join, &set_bits, &inserted, &build_flags));
            if (build_flags || set_bits != inserted) {
                bq_join_free(ctx, join);
                join = nullptr;
                continue;
            }
        } else {
            check(bq_join_build(ctx, &js, &join));
        }
        break;
    }

    // A semi-join bitmap beyond the L2 (a 2-billion-key domain is 250 MB; across GPUs the bitmap spans the GLOBAL key domain)
    // is probed in key-range passes: each pass streams the probe key once (8 bytes per row the first time, 4 afterwards) and
    // tests only the keys of one L2-sized slice, leaving one match bit per row; the fused scan then reads the bits instead of
    // probing, and not the probe key at all.  The kernels mark their streams L2 evict-first and the bitmap evict-last, which
    // is what lets a slice be most of the 126 MB L2.  (Leaving the last slice to the scan - bq_join_probe_bits_but_last -
    // measured slower: the scan is bound by its group updates, not by the probe.)
    // $BOSQL_BITMAP_SLICE_MB sets the slice size (0 = always one fused probe).
    DevColPtr row_bits;
    if (join && bq_join_kind(join) == BQ_JOIN_BITMAP && p.rows > 0) {
        size_t slice_bytes = 80u << 20;
        if (const char* e = std::getenv("BOSQL_BITMAP_SLICE_MB")) slice_bytes = static_cast<size_t>(std::atoll(e)) << 20;
        if (const char* e = std::getenv("BOSQL_BITMAP_SLICE_KB")) slice_bytes = static_cast<size_t>(std::atoll(e)) << 10;      // tests: force passes on small tables
        if (slice_bytes > 0 && bq_join_bytes(join) > slice_bytes + slice_bytes / 4) {
            PhaseTrace ptrace;
            bq_col* b = nullptr;
            check(bq_join_probe_bits(ctx, join, p.cols[p.probe_key].dev->h, 0, p.rows, slice_bytes, &b));
            row_bits = adopt(b);
            ptrace.mark("probe in key-range passes");
        }
    }

    // ---- run the passes ---------------------------------------------------------------------------------------------------
    bool groups_stay_sharded = false;
    bool groups_in_key_order = true;
    std::vector<DeviceRelationPtr> pass_results;
    std::vector<std::vector<size_t>> pass_aggs;      // which aggregates each pass produced, in column order
    for (const Pass& ps : passes) {
        std::vector<int> roles = {key_col, ps.col_a, ps.col_b, p.joined ? p.probe_key : -1};
        // ranges may only ride on streamed (probe-side) role columns
        std::vector<int> range_roles = roles;
        for (int& r : range_roles)
            if (r >= 0 && p.cols[r].side) r = -1;
        SlotPlan plan;
        plan = plan_slots(probe_conj, p.cols, p.rows, range_roles, 3);

        bq_scan_spec s{};
        auto slot_for = [&](int role_index) {
            int c = roles[role_index];
            if (c < 0) return bq_slot{};
            return make_slot(p.cols[c].dev, plan.role_ranges[role_index], p.cols[c].side != 0);
        };
        s.key = slot_for(0);
        s.a = slot_for(1);
        s.b = slot_for(2);
        s.jkey = slot_for(3);
        for (size_t i = 0; i < plan.pred.size(); ++i) s.pred[i] = make_slot(p.cols[plan.pred[i].first].dev, plan.pred[i].second);
        s.mask = plan.mask ? plan.mask->h : nullptr;
        s.row_begin = 0;
        s.row_end = p.rows;
        s.join = join;
        if (row_bits) {
            s.join = nullptr;
            s.row_bits = row_bits->h;
            if (s.jkey.n_ranges == 0) s.jkey = bq_slot{};      // a range riding on the probe key keeps the column as a predicate slot
        }
        s.n_v = static_cast<int32_t>(ps.vals.size());
        for (size_t k = 0; k < ps.vals.size(); ++k) {
            const ValueForm& f = values[ps.vals[k]].form;
            bq_vexpr ve{};
            if (f.unary) {
                ve.op = f.col_l == ps.col_a ? BQ_V_A : BQ_V_B;
            } else {
                ve.op = f.op;
                ve.l_src = f.col_l < 0 ? BQ_L_IMM : (f.col_l == ps.col_a ? BQ_L_A : BQ_L_B);
                ve.r_src = f.col_r < 0 ? BQ_R_IMM : (f.col_r == ps.col_a ? BQ_R_A : BQ_R_B);
                ve.imm_i = f.imm;
            }
            s.v[k] = ve;
        }
        // outputs of this pass
        std::vector<size_t> agg_index;
        for (size_t i = 0; i < req.aggs.size(); ++i) {
            const auto& a = req.aggs[i];
            int vpos = -1;
            if (a.func != "COUNT") {
                for (size_t k = 0; k < ps.vals.size(); ++k)
                    if (ps.vals[k] == agg_value[i]) vpos = static_cast<int>(k);
                if (vpos < 0) continue;
            } else if (&ps != &passes.front()) {
                continue;                                  // COUNT comes out of the first pass
            }
            if (s.n_out >= BQ_MAX_AGG_OUT) throw std::runtime_error("too many aggregates in one query");
            bq_agg_out o{};
            o.func = a.func == "COUNT" ? BQ_AGG_COUNT : (a.func == "SUM" ? BQ_AGG_SUM : BQ_AGG_AVG);
            o.v = vpos < 0 ? 0 : vpos;
            o.as_int = (a.func == "SUM" && a.result_type != TypeId::DOUBLE) ? 1 : 0;
            s.out[s.n_out++] = o;
            agg_index.push_back(i);
        }

        // grouping strategy from statistics
        auto choose_group = [&](bool force_device_stats) {
            if (key_col < 0) {
                s.group_mode = BQ_GROUP_NONE;
                return;
            }
            PipeCol& kc = p.cols[key_col];
            if (kc.type == TypeId::DOUBLE) {
                s.group_mode = BQ_GROUP_HASH;
                s.ndv_hint = kc.stats.ndv ? kc.stats.ndv : 0;
                return;
            }
            resolve_stats(kc, force_device_stats);
            if (dist && kc.stats.measured) xch.minmax(kc.stats.min_key, kc.stats.max_key);
            Domain d{kc.stats.min_key, kc.stats.max_key};
            // every rank must choose the same table kind (the partial states are exchanged in the form the kind implies):
            // across GPUs the row count that the choice weighs is the global one, not this shard's
            uint64_t est_rows = p.rows;
            if (dist && d.size() > 4096 && d.size() <= (1ULL << 26)) est_rows = fetch_global_rows().probe;
            if (d.size() > 0 && d.size() <= (1ULL << 26) && d.size() <= 4 * est_rows + 4096) {
                s.group_mode = BQ_GROUP_DENSE;
                s.key_min = d.lo;
                s.key_max = d.hi;
            } else {
                s.group_mode = BQ_GROUP_HASH;
                uint64_t hint = kc.stats.ndv ? kc.stats.ndv : std::min<uint64_t>(std::max<uint64_t>(p.rows, 1), d.size() ? d.size() : std::max<uint64_t>(p.rows, 1));
                s.ndv_hint = static_cast<size_t>(hint);
            }
        };
        choose_group(false);
        const bool has_key = key_col >= 0;
        const bool int_key = has_key && p.cols[key_col].type != TypeId::DOUBLE;
        const bool plain_stream = !p.joined && probe_conj.empty() && !s.mask;      // key / a / b are read as they lie
        size_t cur_rows = p.rows;
        std::vector<DevColPtr> reordered;        // keeps shuffled / partitioned columns alive through the scan
        bool smem_tables = false;                // shared-memory group tables over the partitioned rows (bq_partition_aggregate)
        int smem_splits = 1;
        const bq_col* part_offsets = nullptr;

        // Across GPUs, a GROUP BY with more groups than fit a partial-state exchange moves the ROWS instead: hash-partition
        // by key, all-to-all, and every rank aggregates the keys it owns (SURVEY.md 8e, high-cardinality GROUP BY).
        bool shuffled = false;
        if (dist && int_key && plain_stream) {
            const PipeCol& kc = p.cols[key_col];
            const Domain d{kc.stats.min_key, kc.stats.max_key};
            uint64_t groups = kc.stats.ndv;                      // catalog statistics describe the whole table
            if (!groups) {
                groups = d.size();
                if (groups == 0 || groups * 32 > (96ull << 20)) {
                    const uint64_t all_rows = fetch_global_rows().probe;
                    groups = groups ? std::min(groups, all_rows) : all_rows;
                }
            }
            if (groups * 32 > (96ull << 20)) {
                std::vector<DevColPtr> pay;
                if (ps.col_a >= 0) pay.push_back(p.cols[ps.col_a].dev);
                if (ps.col_b >= 0) pay.push_back(p.cols[ps.col_b].dev);
                Shuffled sh = shuffle_by_key(kc.dev, pay, p.rows);
                reordered.push_back(sh.key);
                s.key.col = sh.key->h;
                size_t k = 0;
                if (ps.col_a >= 0) { reordered.push_back(sh.payload[k]); s.a.col = sh.payload[k++]->h; }
                if (ps.col_b >= 0) { reordered.push_back(sh.payload[k]); s.b.col = sh.payload[k++]->h; }
                cur_rows = sh.rows;
                s.row_end = cur_rows;
                shuffled = true;
                if (s.group_mode == BQ_GROUP_HASH) {
                    // the keys a rank owns: about groups / world of them, never more than its rows
                    const uint64_t share = groups / static_cast<uint64_t>(xch.world()) + groups / (4 * static_cast<uint64_t>(xch.world())) + 4096;
                    s.ndv_hint = static_cast<size_t>(std::min<uint64_t>(share, std::max<uint64_t>(cur_rows, 1)));
                }
            }
        }

        // High-cardinality GROUP BY: when the hash table would be far larger than L2, order the rows by hash partition
        // first (bq_partition) so each partition's table region is L2-resident while its rows stream by.  The reference
        // updates one unordered_map row by row (src/exec/operator.cpp:988-1005); here it is two streaming passes.
        if (s.group_mode == BQ_GROUP_HASH && plain_stream && int_key && cur_rows >= (1u << 22) &&
            static_cast<uint64_t>(s.ndv_hint) * 32 > (96ull << 20)) {
            int log2p = 4;
            while (log2p < 10 && ((static_cast<uint64_t>(s.ndv_hint) * 64) >> log2p) > (32ull << 20)) ++log2p;
            // Plain column arguments: partition finely enough for one SHARED-MEMORY table per (partition, split) and let
            // bq_partition_aggregate do accumulate and emit in one launch (csrc/bq_groupby.cuh).
            bool plain_args = true;
            for (int k = 0; k < s.n_v; ++k) {
                if (s.v[k].op != BQ_V_A && s.v[k].op != BQ_V_B) plain_args = false;
                const int src = s.v[k].op == BQ_V_A ? ps.col_a : ps.col_b;
                if (src < 0 || p.cols[src].type == TypeId::STRING) plain_args = false;
            }
            int tables_log2p = 0, tables_splits = 0;
            if (group_tables_enabled() && plain_args && bq_group_tables_plan(s.ndv_hint, s.n_v, &tables_log2p, &tables_splits)) {
                log2p = tables_log2p;
                smem_tables = true;
                smem_splits = tables_splits;
            }
            const bq_col* pay[2];
            int n_pay = 0;
            if (ps.col_a >= 0) pay[n_pay++] = s.a.col;
            if (ps.col_b >= 0) pay[n_pay++] = s.b.col;
            bq_col *ok = nullptr, *off = nullptr, *op[2] = {nullptr, nullptr};
            PhaseTrace ptrace;
            check(bq_partition(ctx, s.key.col, pay, n_pay, 0, cur_rows, log2p, 64 - log2p, &ok, op, &off));
            ptrace.mark("local L2 partition");
            reordered.push_back(adopt(ok));
            reordered.push_back(adopt(off));
            part_offsets = off;
            s.key.col = ok;
            int k = 0;
            if (ps.col_a >= 0) { reordered.push_back(adopt(op[k])); s.a.col = op[k++]; }
            if (ps.col_b >= 0) { reordered.push_back(adopt(op[k])); s.b.col = op[k++]; }
            s.hash_part_log2 = log2p;
            s.hash_part_shift = 64 - log2p;
        }

        // One shared-memory table per (partition, split) when the plan above chose it; a table that fills up (statistics too
        // low, a skewed partition) sends the same partitioned rows through the L2-resident table instead.  Local work only:
        // across GPUs every rank may decide for itself.
        auto aggregate_rows = [&](bq_rel** rel) -> int {
            if (smem_tables) {
                const bq_col* args[2] = {nullptr, nullptr};
                for (int k = 0; k < s.n_v; ++k) args[k] = s.v[k].op == BQ_V_A ? s.a.col : s.b.col;
                const int rc = bq_partition_aggregate(ctx, s.key.col, args, s.n_v, part_offsets, s.hash_part_log2, smem_splits, s.out, s.n_out, rel);
                if (!rc) return 0;
                if (!std::strstr(bq_last_error(), "table overflow")) return rc;
                smem_tables = false;
            }
            return bq_scan_aggregate(ctx, &s, rel);
        };
        // a hash table sized from a stale (too low) ndv overflows: the retry sizes it from the rows themselves and gives up
        // the partition-major layout, whose regions assume an even spread of the keys
        auto widen_table = [&] {
            smem_tables = false;
            s.ndv_hint = std::max<size_t>(cur_rows, 1);
            s.hash_part_log2 = 0;
            s.hash_part_shift = 0;
        };
        PhaseTrace trace;
        DeviceRelationPtr r;
        if (!dist) {
            bq_rel* rel = nullptr;
            int rc = aggregate_rows(&rel);
            if (rc && std::strstr(bq_last_error(), "stale statistics")) {
                choose_group(true);                 // the catalog's min/max were wrong: measure and retry
                rc = aggregate_rows(&rel);
            }
            if (rc && std::strstr(bq_last_error(), "table overflow")) {
                widen_table();                      // the catalog's ndv was too low (or a partition is skewed): size by rows
                rc = aggregate_rows(&rel);
            }
            check(rc);
            r = relation_from(rel);
        } else if (shuffled) {
            // every rank owns a disjoint set of keys: aggregate locally, agree on the outcome, then replicate the groups
            auto attempt = [&](std::string& message) {
                bq_rel* rel = nullptr;
                int rc = aggregate_rows(&rel);
                if (rc) message = bq_last_error();
                else r = relation_from(rel);
                int64_t mine = rc ? (message.find("stale statistics") != std::string::npos ? 2 : message.find("table overflow") != std::string::npos ? 3 : 1) : 0, worst = 0;
                for (int64_t f : xch.host_gather({mine})) worst = std::max(worst, f);
                return worst;
            };
            std::string message;
            int64_t outcome = attempt(message);
            if (outcome == 2 || outcome == 3) {       // the worst outcome over the ranks: every rank retries the same way
                if (outcome == 2) choose_group(true);
                if (s.group_mode == BQ_GROUP_HASH) widen_table();
                message.clear();
                outcome = attempt(message);
            }
            if (outcome) throw std::runtime_error(message.empty() ? "aggregation failed on another rank" : message);
            if (xch.fn.keep_sharded) groups_stay_sharded = true;
            if (!xch.fn.keep_sharded) {
                std::vector<TypeId> rel_types;
                if (has_key) rel_types.push_back(p.cols[key_col].type);
                for (int o = 0; o < s.n_out; ++o)
                    rel_types.push_back(s.out[o].func == BQ_AGG_COUNT ? TypeId::INT64 : (s.out[o].func == BQ_AGG_SUM && s.out[o].as_int ? TypeId::INT64 : TypeId::DOUBLE));
                r = all_gather_relation(r, rel_types);
            }
        } else {
            // local fused scan -> partial states [key] count sum0 sum1 -> one all-gather -> merge in rank order.  A failed
            // local scan (division by zero, stale bounds) travels as a poisoned partial, so every rank's merge fails alike.
            const TypeId key_type = has_key ? p.cols[key_col].type : TypeId::INT64;
            auto attempt = [&]() {
                // dense / global states travel as they are: one all-gather of the raw count | sum arrays, one fold launch
                const bool dense_state = s.group_mode == BQ_GROUP_NONE ||
                                         (s.group_mode == BQ_GROUP_DENSE && static_cast<uint64_t>(s.key_max - s.key_min) < (1ull << 18));
                if (dense_state) {
                    bq_agg_state* st = nullptr;
                    check(bq_scan_state(ctx, &s, &st));
                    struct StateGuard {
                        bq_agg_state* s;
                        ~StateGuard() { bq_agg_state_free(s); }
                    } state_guard{st};
                    all_gather_fold_state(st);
                    bq_rel* fin = nullptr;
                    int rc2 = bq_agg_state_emit(ctx, st, s.out, s.n_out, &fin);
                    if (!rc2) r = relation_from(fin);
                    return rc2;
                }
                bq_rel* rel = nullptr;
                int rc = bq_scan_partial(ctx, &s, &rel);
                int flags = 0;
                DeviceRelationPtr local;
                if (rc) {
                    const char* m = bq_last_error();
                    flags = std::strstr(m, "Division by zero") ? 1 : std::strstr(m, "table overflow") ? 2 : std::strstr(m, "stale statistics") ? 4 : 0;
                    if (!flags) check(rc);
                } else {
                    local = relation_from(rel);
                }
                int64_t cap = -1;
                if (!has_key) cap = 1;
                else if (int_key) {
                    const Domain d{p.cols[key_col].stats.min_key, p.cols[key_col].stats.max_key};
                    if (p.cols[key_col].stats.known && d.size() > 0 && d.size() <= 65536) cap = static_cast<int64_t>(d.size());
                }
                if (cap >= 0 && local && static_cast<int64_t>(local->rows) > cap) {
                    local.reset();
                    flags = 4;                      // more groups than the bounds allow: they are stale
                }
                GatheredPartials g;
                gather_partials(local.get(), flags, has_key, key_type, cap, g);
                bq_rel* fin = nullptr;
                int rc2 = bq_agg_finish(ctx, g.parts.data(), static_cast<int>(g.parts.size()), has_key ? 1 : 0, static_cast<int>(key_type), s.out, s.n_out, &fin);
                if (!rc2) r = relation_from(fin);
                return rc2;
            };
            int rc = attempt();
            if (rc && std::strstr(bq_last_error(), "stale statistics")) {
                choose_group(true);
                rc = attempt();
            }
            if (rc && std::strstr(bq_last_error(), "table overflow")) {
                widen_table();
                rc = attempt();
            }
            check(rc);
        }
        // with several passes over a keyed aggregate, bring every pass into key order so rows line up
        if (passes.size() > 1 && key_col >= 0 && r->rows > 1) {
            std::vector<bq_col*> hs;
            for (auto& c : r->cols) hs.push_back(c->h);
            bq_rel* shell = nullptr;
            // a non-owning shell around r's columns
            check(bq_rel_create(ctx, hs.data(), static_cast<int>(hs.size()), &shell));
            int kc0 = 0, asc = 1;
            bq_rel* sorted = nullptr;
            int rc2 = bq_rel_sort(ctx, shell, 1, &kc0, &asc, -1, &sorted);
            std::vector<bq_col*> back(hs.size());
            bq_rel_release(shell, back.data());
            check(rc2);
            r = relation_from(sorted);
        }
        trace.mark("aggregate (+ exchange)");
        // dense states (local, or all-gathered and folded) emit their groups in ascending key order; so does the re-sort above
        const bool dense_emit = s.group_mode == BQ_GROUP_DENSE && !shuffled &&
                                (!dist || static_cast<uint64_t>(s.key_max - s.key_min) < (1ull << 18));
        if (!(dense_emit || (passes.size() > 1 && key_col >= 0))) groups_in_key_order = false;
        pass_results.push_back(r);
        pass_aggs.push_back(agg_index);
    }

    // ---- stitch: [keys] then aggregates in declaration order ----------------------------------------------------
    auto out = std::make_shared<DeviceRelation>();
    out->replicated = dist && !groups_stay_sharded;
    out->ordered_by_first = key_col >= 0 && packed_domains.empty() && groups_in_key_order;
    out->rows = pass_results.front()->rows;
    for (auto& pr : pass_results)
        if (pr->rows != out->rows) throw std::runtime_error("internal: aggregate passes disagree on the group count");
    const int key_cols_in_rel = key_col >= 0 ? 1 : 0;
    std::vector<DevColPtr> agg_cols(req.aggs.size());
    for (size_t pi = 0; pi < pass_results.size(); ++pi)
        for (size_t k = 0; k < pass_aggs[pi].size(); ++k) agg_cols[pass_aggs[pi][k]] = pass_results[pi]->cols[key_cols_in_rel + k];
    if (key_col >= 0) {
        DevColPtr key = pass_results.front()->cols[0];
        if (packed_domains.empty()) {
            out->cols.push_back(key);
        } else {
            // unpack the combined key: k_i = (packed / stride_i) % size_i + min_i, with q % s = q - (q / s) * s
            for (size_t i = 0; i < packed_domains.size(); ++i) {
                const int64_t st = static_cast<int64_t>(packed_strides[i]);
                const int64_t size = static_cast<int64_t>(packed_domains[i].size());
                if (out->rows == 0) {
                    bq_col* h = nullptr;
                    check(bq_col_alloc(ctx, static_cast<int>(req.group_types[i]), 0, &h));
                    out->cols.push_back(adopt(h));
                    continue;
                }
                std::vector<bq_insn> code = {
                    insn(BQ_OP_COL, 0), insn(BQ_OP_IMM_I, 0, st), insn(BQ_OP_DIV_I),                       // q
                    insn(BQ_OP_COL, 0), insn(BQ_OP_IMM_I, 0, st), insn(BQ_OP_DIV_I),                       // q
                    insn(BQ_OP_IMM_I, 0, size), insn(BQ_OP_DIV_I), insn(BQ_OP_IMM_I, 0, size), insn(BQ_OP_MUL_I),
                    insn(BQ_OP_SUB_I),                                                                        // q % size
                    insn(BQ_OP_IMM_I, 0, packed_domains[i].lo), insn(BQ_OP_ADD_I)};
                out->cols.push_back(eval_program(code, {key}, out->rows, req.group_types[i]));
            }
        }
    }
    for (size_t i = 0; i < req.aggs.size(); ++i) {
        if (!agg_cols[i]) throw std::runtime_error("internal: aggregate output missing");
        out->cols.push_back(agg_cols[i]);
    }
    return out;
}

}  // namespace bosql::gpu
