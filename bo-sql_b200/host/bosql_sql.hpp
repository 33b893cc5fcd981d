// bosql_sql.hpp — the callers above the hot path: AST, SQL parser and logical plan, API-compatible with the
// reference's include/parser/{ast,parser}.h and include/logical/{logical,planner}.h.
//
// These are NOT the product (SURVEY.md section 2 rows 10-11 are out of scope as a rewrite); they exist so the
// operator layer can be driven by the same SQL text as the reference in tests, bench.py and the CLI, with the
// same grammar quirks: qualified names are literal strings ("l.sku", src/parser/parser.cpp:283-289), there is no
// BETWEEN keyword (an unknown identifier after a predicate silently ends the statement, :135-157), numbers are
// unsigned integers only (:28-34), a single JOIN is planned (src/logical/planner.cpp:65-91) and the WHERE filter
// sits directly above the base relation (:110-117).  Two documented extensions are OFF by
// default and switched on per parse: BETWEEN and decimal literals.
#pragma once

#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "bosql_types.hpp"

namespace bosql {

enum class ExprType { COLUMN_REF, LITERAL_INT, LITERAL_DOUBLE, LITERAL_STRING, BINARY_OP, FUNC_CALL };
enum class BinaryOp { EQ, NE, LT, LE, GT, GE, ADD, SUB, MUL, DIV, AND, OR };

struct Expr {
    ExprType type = ExprType::COLUMN_REF;
    std::string str_val;
    i64 i64_val = 0;
    f64 f64_val = 0.0;
    BinaryOp op = BinaryOp::EQ;
    std::unique_ptr<Expr> left, right;
    std::string func_name;
    std::vector<std::unique_ptr<Expr>> args;

    std::string to_string() const;
    std::unique_ptr<Expr> clone() const;
};

struct SelectItem {
    std::string alias;
    std::unique_ptr<Expr> expr;
    std::string to_string() const;
};

enum class AggFunc { NONE, SUM, COUNT, AVG };

struct GroupByClause {
    std::vector<std::unique_ptr<Expr>> columns;
    std::unique_ptr<Expr> having;
    std::string to_string() const;
};

struct OrderByItem {
    std::unique_ptr<Expr> expr;
    bool asc = true;
    std::string to_string() const;
};

struct TableRef {
    std::string table_name;
    std::string alias;
    std::string to_string() const;
};

struct JoinItem {
    TableRef table_ref;
    std::unique_ptr<Expr> on_condition;
    std::string to_string() const;
};

struct SelectStmt {
    std::vector<SelectItem> select_list;
    TableRef from_table;
    std::unique_ptr<Expr> where_clause;
    std::vector<JoinItem> joins;
    GroupByClause group_by;
    std::vector<OrderByItem> order_by;
    int limit = -1;
    std::string to_string() const;
};

struct ParseOptions {
    bool between = false;           // col BETWEEN a AND b  ->  (col >= a) AND (col <= b)
    bool decimal_literals = false;  // 1.5 -> LITERAL_DOUBLE
    bool negative_literals = false; // -5 / -1.5 where an operand is expected -> a literal (the reference lexes '-' as MINUS only
                                    // and its primary() rejects it, src/parser/parser.cpp:278-331)
    bool keywords_any_case = false; // select / Select / SELECT (the reference matches keywords exactly, :82-103)
};

SelectStmt parse_sql(const std::string& sql);
SelectStmt parse_sql(const std::string& sql, const ParseOptions& opts);

// ---- logical plan ---------------------------------------------------------------------------------
enum class LogicalOpType { SCAN, FILTER, PROJECT, HASH_JOIN, AGGREGATE, ORDER, LIMIT };

struct LogicalOp {
    LogicalOpType type;
    std::vector<std::unique_ptr<LogicalOp>> children;
    explicit LogicalOp(LogicalOpType t) : type(t) {}
    virtual ~LogicalOp() = default;
    virtual std::string to_string(int indent = 0) const = 0;
protected:
    std::string with_children(std::string head, int indent) const;
};

// every node prints itself the way the reference's LogicalOp::to_string does (plan-shape tests compare the text)
#define BQ_PLAN_TEXT std::string to_string(int indent = 0) const override;

struct LogicalScan : LogicalOp {
    std::string table_name;
    std::vector<std::string> columns;
    LogicalScan(const std::string& table, const std::vector<std::string>& cols)
        : LogicalOp(LogicalOpType::SCAN), table_name(table), columns(cols) {}
    BQ_PLAN_TEXT
};

struct LogicalFilter : LogicalOp {
    std::unique_ptr<Expr> predicate;
    explicit LogicalFilter(std::unique_ptr<Expr> pred) : LogicalOp(LogicalOpType::FILTER), predicate(std::move(pred)) {}
    BQ_PLAN_TEXT
};

struct LogicalProject : LogicalOp {
    std::vector<std::unique_ptr<Expr>> select_list;
    std::vector<std::string> aliases;
    LogicalProject(std::vector<std::unique_ptr<Expr>>&& selects, std::vector<std::string>&& alias_list)
        : LogicalOp(LogicalOpType::PROJECT), select_list(std::move(selects)), aliases(std::move(alias_list)) {}
    BQ_PLAN_TEXT
};

struct LogicalHashJoin : LogicalOp {
    std::vector<std::string> left_keys, right_keys;
    std::unique_ptr<Expr> join_filter;
    LogicalHashJoin(std::vector<std::string> l, std::vector<std::string> r, std::unique_ptr<Expr> filter = nullptr)
        : LogicalOp(LogicalOpType::HASH_JOIN), left_keys(std::move(l)), right_keys(std::move(r)), join_filter(std::move(filter)) {}
    BQ_PLAN_TEXT
};

struct LogicalAggregate : LogicalOp {
    struct AggExpr {
        std::string func_name;
        std::unique_ptr<Expr> arg;
        std::string alias;
    };
    std::vector<std::unique_ptr<Expr>> group_keys;
    std::vector<AggExpr> aggregates;
    LogicalAggregate(std::vector<std::unique_ptr<Expr>>&& keys, std::vector<AggExpr>&& aggs)
        : LogicalOp(LogicalOpType::AGGREGATE), group_keys(std::move(keys)), aggregates(std::move(aggs)) {}
    BQ_PLAN_TEXT
};

struct LogicalOrder : LogicalOp {
    struct OrderItem {
        std::unique_ptr<Expr> expr;
        bool asc;
    };
    std::vector<OrderItem> order_by;
    explicit LogicalOrder(std::vector<OrderItem>&& order) : LogicalOp(LogicalOpType::ORDER), order_by(std::move(order)) {}
    BQ_PLAN_TEXT
};

struct LogicalLimit : LogicalOp {
    int64_t limit;
    explicit LogicalLimit(int64_t lim) : LogicalOp(LogicalOpType::LIMIT), limit(lim) {}
    BQ_PLAN_TEXT
};

class LogicalPlanner {
public:
    std::unique_ptr<LogicalOp> build_logical_plan(const SelectStmt& stmt);
};

}  // namespace bosql
