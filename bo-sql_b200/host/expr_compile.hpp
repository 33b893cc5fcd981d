// expr_compile.hpp — exec/expression.cpp compiled once per plan instead of interpreted once per row.
//
// The reference resolves column names (unordered_map lookup), string literals (linear dictionary search) and
// operand types (Datum tags) for EVERY row (src/exec/expression.cpp:153-206).  Here all three are resolved on
// the host when the plan is opened:
//   * to_ranges():  a `column OP literal` conjunct becomes an integer range on the column's order-preserving key,
//                   evaluated inside the fused kernels (bq_slot.r);
//   * compile():    anything else becomes a typed postfix program for bq_eval.
// Both reproduce compare_values / numeric_binary / is_truthy exactly (hazards H6-H11 of SURVEY.md 8a).
#pragma once

#include <functional>
#include <string>
#include <vector>

#include "bosql_b200.h"
#include "bosql_sql.hpp"
#include "bosql_types.hpp"

namespace bosql::gpu {

// name -> (index, type); index < 0 = unknown column
struct ColumnLookup {
    std::function<int(const std::string&)> index_of;
    std::function<TypeId(int)> type_of;
};

struct Program {
    std::vector<bq_insn> code;
    std::vector<int> columns;      // program column slot -> lookup index
    TypeId result = TypeId::INT64;
};

// Typed program for `e`.  as_predicate appends is_truthy (result INT64 0/1).
// Throws the reference's messages: "Unknown column: x", "String literal without dictionary binding",
// "Cannot coerce string to numeric", "Unsupported string comparison",
// "Function calls not supported in expression evaluation".
Program compile(const Expr* e, const ColumnLookup& cols, Dictionary* dict, bool as_predicate);

// The static result type of `e` under evaluate_internal's rules (what compile() would leave on the stack).
TypeId value_type(const Expr* e, const ColumnLookup& cols);

struct ColumnRange {
    int column = -1;       // lookup index
    bq_range range{};
};
// If `e` is `column OP literal` (or a bare column used as a truth value, or `int literal OP int column`),
// produce the equivalent key range.  Returns false when the conjunct needs the general program.
bool to_range(const Expr* e, const ColumnLookup& cols, Dictionary* dict, ColumnRange& out);

// True when evaluating `e` can throw for SOME row values (an integer division whose divisor is not a non-zero literal,
// src/exec/expression.cpp:52).  Such an expression must only be evaluated for the rows the reference would evaluate it for.
bool may_throw_per_row(const Expr* e, const ColumnLookup& cols);

// Collects the lookup indices of every column `e` references (throws "Unknown column: x").
void referenced(const Expr* e, const ColumnLookup& cols, std::vector<int>& out);

int64_t f64_key(double v);

}  // namespace bosql::gpu
