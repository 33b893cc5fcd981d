// physical_planner.cpp — LogicalOp tree -> GPU operator tree, one node per node like the reference's
// build_physical_plan (src/exec/physical_planner.cpp:9-124): scans receive the union of referenced names and keep
// the ones their table has (:20-32), an empty match means all columns, a Project over an Aggregate is elided
// (:50-52), aggregate function names are upper-cased (:90).  The only addition: each scan is handed its table's
// catalog statistics (TableMeta: row_count, min/max/ndv), from which the kernels' tables are sized.
#include <algorithm>
#include <cctype>

#include "bosql_operator.hpp"

namespace bosql {

std::unique_ptr<Operator> build_physical_plan(const LogicalOp* logical, const Catalog& catalog) {
    auto child_plan = [&](size_t i) { return build_physical_plan(logical->children.at(i).get(), catalog); };
    switch (logical->type) {
        case LogicalOpType::SCAN: {
            const auto* scan = dynamic_cast<const LogicalScan*>(logical);
            if (!scan) throw std::runtime_error("Invalid LogicalScan");
            OptionalRef<const Table> table = catalog.get_table_data(scan->table_name);
            if (!table.has_value()) throw std::runtime_error("Table not found: " + scan->table_name);
            std::vector<size_t> indices;
            for (const auto& name : scan->columns)
                for (size_t i = 0; i < table->columns.size(); ++i)
                    if (table->columns[i].name == name) {
                        indices.push_back(i);
                        break;
                    }
            auto op = std::make_unique<ColumnarScan>(const_cast<Table*>(&table.value()), std::move(indices));
            OptionalRef<const TableMeta> meta = catalog.get_table_meta(scan->table_name);
            if (meta.has_value()) op->set_table_meta(&meta.value());
            return op;
        }
        case LogicalOpType::FILTER: {
            const auto* f = dynamic_cast<const LogicalFilter*>(logical);
            if (!f) throw std::runtime_error("Invalid LogicalFilter");
            return std::make_unique<Selection>(child_plan(0), f->predicate->clone());
        }
        case LogicalOpType::PROJECT: {
            const auto* p = dynamic_cast<const LogicalProject*>(logical);
            if (!p) throw std::runtime_error("Invalid LogicalProject");
            auto child = child_plan(0);
            if (p->select_list.empty() || p->children[0]->type == LogicalOpType::AGGREGATE) return child;
            std::vector<std::unique_ptr<Expr>> exprs;
            for (const auto& e : p->select_list) exprs.push_back(e->clone());
            return std::make_unique<Project>(std::move(child), std::move(exprs), p->aliases);
        }
        case LogicalOpType::HASH_JOIN: {
            const auto* j = dynamic_cast<const LogicalHashJoin*>(logical);
            if (!j) throw std::runtime_error("Invalid LogicalHashJoin");
            auto left = child_plan(0);
            auto right = child_plan(1);
            return std::make_unique<HashJoin>(std::move(left), std::move(right), j->left_keys, j->right_keys,
                                              j->join_filter ? j->join_filter->clone() : nullptr);
        }
        case LogicalOpType::AGGREGATE: {
            const auto* a = dynamic_cast<const LogicalAggregate*>(logical);
            if (!a) throw std::runtime_error("Invalid LogicalAggregate");
            auto child = child_plan(0);
            std::vector<std::unique_ptr<Expr>> keys;
            for (const auto& k : a->group_keys) keys.push_back(k->clone());
            std::vector<AggregateSpec> specs;
            for (const auto& agg : a->aggregates) {
                AggregateSpec s;
                s.func_name = agg.func_name;
                std::transform(s.func_name.begin(), s.func_name.end(), s.func_name.begin(),
                               [](unsigned char c) { return static_cast<char>(std::toupper(c)); });
                s.alias = agg.alias;
                if (agg.arg) s.arg = agg.arg->clone();
                specs.push_back(std::move(s));
            }
            return std::make_unique<HashAggregate>(std::move(child), std::move(keys), std::move(specs));
        }
        case LogicalOpType::ORDER: {
            const auto* o = dynamic_cast<const LogicalOrder*>(logical);
            if (!o) throw std::runtime_error("Invalid LogicalOrder");
            auto child = child_plan(0);
            std::vector<OrderBy::SortKey> keys;
            for (const auto& it : o->order_by) keys.push_back({it.expr->clone(), it.asc});
            return std::make_unique<OrderBy>(std::move(child), std::move(keys));
        }
        case LogicalOpType::LIMIT: {
            const auto* l = dynamic_cast<const LogicalLimit*>(logical);
            if (!l) throw std::runtime_error("Invalid LogicalLimit");
            return std::make_unique<Limit>(child_plan(0), l->limit);
        }
    }
    throw std::runtime_error("Unsupported logical operator");
}

}  // namespace bosql
