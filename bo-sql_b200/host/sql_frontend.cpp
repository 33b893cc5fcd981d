// sql_frontend.cpp — tokenizer, recursive-descent parser and logical planner behind bosql_sql.hpp.
//
// Behavioural contract = the reference's src/parser/parser.cpp and src/logical/planner.cpp (file:line cited per
// function); written from that behaviour, not from its text.  It feeds the GPU operators the same Expr trees and
// plan shapes the reference feeds its CPU operators, so one SQL string drives both in the differential tests.
#include <algorithm>
#include <cctype>
#include <charconv>
#include <set>

#include "bosql_sql.hpp"

namespace bosql {

// ---- Expr helpers (reference: src/parser/ast_to_string.cpp) -----------------------------------------
static const char* op_text(BinaryOp op) {
    static const char* names[] = {"=", "!=", "<", "<=", ">", ">=", "+", "-", "*", "/", "AND", "OR"};
    return names[static_cast<int>(op)];
}

static std::string shortest_double(double v) {
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, v);
    return std::string(buf, r.ptr);
}

std::string Expr::to_string() const {
    switch (type) {
        case ExprType::COLUMN_REF: return str_val;
        case ExprType::LITERAL_INT: return std::to_string(i64_val);
        case ExprType::LITERAL_DOUBLE: return shortest_double(f64_val);
        case ExprType::LITERAL_STRING: return "'" + str_val + "'";
        case ExprType::BINARY_OP: return "(" + left->to_string() + " " + op_text(op) + " " + right->to_string() + ")";
        case ExprType::FUNC_CALL: {
            std::string s = func_name + "(";
            for (size_t i = 0; i < args.size(); ++i) s += (i ? ", " : "") + args[i]->to_string();
            return s + ")";
        }
    }
    return "UNKNOWN_EXPR";
}

std::unique_ptr<Expr> Expr::clone() const {
    auto c = std::make_unique<Expr>();
    c->type = type;
    c->str_val = str_val;
    c->i64_val = i64_val;
    c->f64_val = f64_val;
    c->op = op;
    c->func_name = func_name;
    if (left) c->left = left->clone();
    if (right) c->right = right->clone();
    for (const auto& a : args) c->args.push_back(a->clone());
    return c;
}

std::string SelectItem::to_string() const { return alias.empty() ? expr->to_string() : expr->to_string() + " AS " + alias; }
std::string OrderByItem::to_string() const { return expr->to_string() + (asc ? " ASC" : " DESC"); }
std::string TableRef::to_string() const { return alias.empty() ? table_name : table_name + " " + alias; }
std::string JoinItem::to_string() const { return "JOIN " + table_ref.to_string() + " ON " + on_condition->to_string(); }

std::string GroupByClause::to_string() const {
    if (columns.empty()) return "";
    std::string s = "GROUP BY ";
    for (size_t i = 0; i < columns.size(); ++i) s += (i ? ", " : "") + columns[i]->to_string();
    if (having) s += " HAVING " + having->to_string();
    return s;
}

std::string SelectStmt::to_string() const {
    std::string s = "SELECT ";
    for (size_t i = 0; i < select_list.size(); ++i) s += (i ? ", " : "") + select_list[i].to_string();
    s += " FROM " + from_table.to_string();
    for (const auto& j : joins) s += " " + j.to_string();
    if (where_clause) s += " WHERE " + where_clause->to_string();
    if (!group_by.columns.empty()) s += " " + group_by.to_string();
    if (!order_by.empty()) {
        s += " ORDER BY ";
        for (size_t i = 0; i < order_by.size(); ++i) s += (i ? ", " : "") + order_by[i].to_string();
    }
    if (limit >= 0) s += " LIMIT " + std::to_string(limit);
    return s;
}

// ---- tokens (ordinals match the reference's TokenType so "Expected N got M" reads the same) --------
namespace {

enum class Tok {
    SELECT = 0, FROM, WHERE, INNER, JOIN, ON, GROUP, BY, HAVING, ORDER, ASC, DESC, LIMIT,
    IDENTIFIER, NUMBER, STRING_LITERAL, COMMA, LPAREN, RPAREN, EQ, NE, LT, LE, GT, GE, PLUS, MINUS, MUL, DIV,
    SUM, COUNT, AVG, AS, AND, OR,
    END,
    DECIMAL   // extension token (ParseOptions::decimal_literals)
};

struct Token {
    Tok type;
    std::string text;
};

Tok keyword_or_identifier(const std::string& word, bool any_case) {
    // exact, case-sensitive (src/parser/parser.cpp:82-103) unless the extension flag folds the word to upper case first
    std::string s = word;
    if (any_case) std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return static_cast<char>(std::toupper(c)); });
    static const std::pair<const char*, Tok> kw[] = {
        {"SELECT", Tok::SELECT}, {"FROM", Tok::FROM}, {"WHERE", Tok::WHERE}, {"INNER", Tok::INNER}, {"JOIN", Tok::JOIN},
        {"ON", Tok::ON}, {"GROUP", Tok::GROUP}, {"BY", Tok::BY}, {"HAVING", Tok::HAVING}, {"ORDER", Tok::ORDER},
        {"ASC", Tok::ASC}, {"DESC", Tok::DESC}, {"LIMIT", Tok::LIMIT}, {"SUM", Tok::SUM}, {"COUNT", Tok::COUNT},
        {"AVG", Tok::AVG}, {"AS", Tok::AS}, {"AND", Tok::AND}, {"OR", Tok::OR}};
    for (const auto& [text, tok] : kw)
        if (s == text) return tok;
    return Tok::IDENTIFIER;
}

// src/parser/parser.cpp:15-80
std::vector<Token> lex(const std::string& sql, const ParseOptions& opts) {
    std::vector<Token> out;
    const size_t n = sql.size();
    size_t i = 0;
    auto uc = [&](size_t k) { return static_cast<unsigned char>(sql[k]); };
    while (i < n) {
        const char c = sql[i];
        if (std::isspace(uc(i))) { ++i; continue; }
        if (std::isalpha(uc(i)) || c == '_') {
            size_t j = i + 1;
            while (j < n && (std::isalnum(uc(j)) || sql[j] == '_')) ++j;
            std::string word = sql.substr(i, j - i);
            out.push_back({keyword_or_identifier(word, opts.keywords_any_case), word});
            i = j;
            continue;
        }
        if (std::isdigit(uc(i))) {
            size_t j = i + 1;
            while (j < n && std::isdigit(uc(j))) ++j;
            if (opts.decimal_literals && j + 1 < n && sql[j] == '.' && std::isdigit(uc(j + 1))) {
                size_t k = j + 1;
                while (k < n && std::isdigit(uc(k))) ++k;
                out.push_back({Tok::DECIMAL, sql.substr(i, k - i)});
                i = k;
                continue;
            }
            out.push_back({Tok::NUMBER, sql.substr(i, j - i)});
            i = j;
            continue;
        }
        if (c == '\'') {
            size_t j = i + 1;
            while (j < n && sql[j] != '\'') ++j;
            out.push_back({Tok::STRING_LITERAL, "'" + sql.substr(i + 1, j - i - 1) + "'"});
            i = j < n ? j + 1 : n;
            continue;
        }
        const bool eq_next = i + 1 < n && sql[i + 1] == '=';
        switch (c) {
            case ',': out.push_back({Tok::COMMA, ","}); break;
            case '(': out.push_back({Tok::LPAREN, "("}); break;
            case ')': out.push_back({Tok::RPAREN, ")"}); break;
            case '=': out.push_back({Tok::EQ, "="}); break;
            case '<': if (eq_next) { out.push_back({Tok::LE, "<="}); ++i; } else out.push_back({Tok::LT, "<"}); break;
            case '>': if (eq_next) { out.push_back({Tok::GE, ">="}); ++i; } else out.push_back({Tok::GT, ">"}); break;
            case '!': if (eq_next) { out.push_back({Tok::NE, "!="}); ++i; } break;   // a lone '!' vanishes (:64-68)
            case '+': out.push_back({Tok::PLUS, "+"}); break;
            case '-': out.push_back({Tok::MINUS, "-"}); break;
            case '*': out.push_back({Tok::MUL, "*"}); break;
            case '/': out.push_back({Tok::DIV, "/"}); break;
            case '.': out.push_back({Tok::IDENTIFIER, "."}); break;               // qualified names (:73)
            case ';': break;
            default: throw std::runtime_error("Unknown token: " + std::string(1, c));
        }
        ++i;
    }
    out.push_back({Tok::END, ""});
    return out;
}

std::unique_ptr<Expr> binary(BinaryOp op, std::unique_ptr<Expr> l, std::unique_ptr<Expr> r) {
    auto e = std::make_unique<Expr>();
    e->type = ExprType::BINARY_OP;
    e->op = op;
    e->left = std::move(l);
    e->right = std::move(r);
    return e;
}

class Parser {
public:
    Parser(const std::string& sql, const ParseOptions& opts) : toks_(lex(sql, opts)), opts_(opts) {}

    // src/parser/parser.cpp:108-158
    SelectStmt select() {
        SelectStmt st;
        expect(Tok::SELECT);
        st.select_list = select_list();
        expect(Tok::FROM);
        st.from_table = table_ref();
        while (at(Tok::INNER) || at(Tok::JOIN)) {
            if (at(Tok::INNER)) ++pos_;
            expect(Tok::JOIN);
            JoinItem j;
            j.table_ref = table_ref();
            expect(Tok::ON);
            j.on_condition = expr();
            st.joins.push_back(std::move(j));
        }
        if (accept(Tok::WHERE)) st.where_clause = expr();
        if (accept(Tok::GROUP)) {
            expect(Tok::BY);
            st.group_by.columns = expr_list();
            if (accept(Tok::HAVING)) st.group_by.having = expr();
        }
        if (accept(Tok::ORDER)) {
            expect(Tok::BY);
            do {
                OrderByItem it;
                it.expr = expr();
                if (accept(Tok::DESC)) it.asc = false;
                else accept(Tok::ASC);
                st.order_by.push_back(std::move(it));
            } while (accept(Tok::COMMA));
        }
        if (accept(Tok::LIMIT)) st.limit = std::stoi(expect(Tok::NUMBER).text);
        return st;   // whatever follows is ignored, as in the reference
    }

private:
    std::vector<Token> toks_;
    ParseOptions opts_;
    size_t pos_ = 0;

    const Token& cur() const { return toks_[std::min(pos_, toks_.size() - 1)]; }
    bool at(Tok t) const { return cur().type == t; }
    bool accept(Tok t) {
        if (!at(t)) return false;
        ++pos_;
        return true;
    }
    Token expect(Tok t) {
        if (!at(t))
            throw std::runtime_error("Expected " + std::to_string(static_cast<int>(t)) + " got " +
                                     std::to_string(static_cast<int>(cur().type)));
        return toks_[pos_++];
    }

    TableRef table_ref() {
        TableRef r;
        r.table_name = expect(Tok::IDENTIFIER).text;
        if (at(Tok::IDENTIFIER)) r.alias = toks_[pos_++].text;
        return r;
    }

    // src/parser/parser.cpp:160-180 — a bare '*' adds nothing to the list
    std::vector<SelectItem> select_list() {
        std::vector<SelectItem> out;
        do {
            if (accept(Tok::MUL)) continue;
            SelectItem it;
            it.expr = expr();
            if (accept(Tok::AS)) it.alias = expect(Tok::IDENTIFIER).text;
            out.push_back(std::move(it));
        } while (accept(Tok::COMMA));
        return out;
    }

    std::vector<std::unique_ptr<Expr>> expr_list() {
        std::vector<std::unique_ptr<Expr>> out;
        do out.push_back(expr()); while (accept(Tok::COMMA));
        return out;
    }

    // precedence: OR < AND < comparison (one, non-associative) < + - < * /   (:182-268)
    std::unique_ptr<Expr> expr() {
        auto l = conjunction();
        while (accept(Tok::OR)) l = binary(BinaryOp::OR, std::move(l), conjunction());
        return l;
    }
    std::unique_ptr<Expr> conjunction() {
        auto l = comparison();
        while (accept(Tok::AND)) l = binary(BinaryOp::AND, std::move(l), comparison());
        return l;
    }
    std::unique_ptr<Expr> comparison() {
        auto l = sum();
        static const std::pair<Tok, BinaryOp> ops[] = {{Tok::EQ, BinaryOp::EQ}, {Tok::NE, BinaryOp::NE}, {Tok::LT, BinaryOp::LT},
                                                      {Tok::LE, BinaryOp::LE}, {Tok::GT, BinaryOp::GT}, {Tok::GE, BinaryOp::GE}};
        for (const auto& [tok, op] : ops)
            if (accept(tok)) return binary(op, std::move(l), sum());
        if (opts_.between && at(Tok::IDENTIFIER) && is_word(cur().text, "BETWEEN")) {      // extension, off by default
            ++pos_;
            auto lo = sum();
            expect(Tok::AND);
            auto hi = sum();
            auto l2 = l->clone();
            return binary(BinaryOp::AND, binary(BinaryOp::GE, std::move(l), std::move(lo)),
                          binary(BinaryOp::LE, std::move(l2), std::move(hi)));
        }
        return l;
    }
    std::unique_ptr<Expr> sum() {
        auto l = product();
        while (at(Tok::PLUS) || at(Tok::MINUS)) {
            BinaryOp op = at(Tok::PLUS) ? BinaryOp::ADD : BinaryOp::SUB;
            ++pos_;
            l = binary(op, std::move(l), product());
        }
        return l;
    }
    std::unique_ptr<Expr> product() {
        auto l = factor();
        while (at(Tok::MUL) || at(Tok::DIV)) {
            BinaryOp op = at(Tok::MUL) ? BinaryOp::MUL : BinaryOp::DIV;
            ++pos_;
            l = binary(op, std::move(l), factor());
        }
        return l;
    }
    bool is_word(const std::string& text, const char* upper) const {
        if (!opts_.keywords_any_case) return text == upper;
        std::string s = text;
        std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return static_cast<char>(std::toupper(c)); });
        return s == upper;
    }
    std::unique_ptr<Expr> factor() {
        // extension, off by default: a minus sign directly in front of a number where an operand is expected is part of the literal
        if (opts_.negative_literals && at(Tok::MINUS) && pos_ + 1 < toks_.size() &&
            (toks_[pos_ + 1].type == Tok::NUMBER || toks_[pos_ + 1].type == Tok::DECIMAL)) {
            ++pos_;
            auto e = primary();
            if (e->type == ExprType::LITERAL_INT) e->i64_val = -e->i64_val;
            else e->f64_val = -e->f64_val;
            return e;
        }
        if (accept(Tok::LPAREN)) {
            auto e = expr();
            expect(Tok::RPAREN);
            return e;
        }
        return primary();
    }

    // src/parser/parser.cpp:278-331
    std::unique_ptr<Expr> primary() {
        Token t = cur();
        ++pos_;
        auto e = std::make_unique<Expr>();
        switch (t.type) {
            case Tok::IDENTIFIER: case Tok::SUM: case Tok::COUNT: case Tok::AVG: {
                std::string name = t.text;
                if (opts_.keywords_any_case && t.type != Tok::IDENTIFIER)      // Sum(...) / sum(...) name the aggregate SUM
                    std::transform(name.begin(), name.end(), name.begin(), [](unsigned char c) { return static_cast<char>(std::toupper(c)); });
                if (at(Tok::IDENTIFIER) && cur().text == ".") {     // "l" "." "sku" -> the literal name "l.sku"
                    ++pos_;
                    name += "." + cur().text;
                    ++pos_;
                }
                if (accept(Tok::LPAREN)) {
                    e->type = ExprType::FUNC_CALL;
                    e->func_name = name;
                    if (!at(Tok::RPAREN)) e->args = expr_list();
                    expect(Tok::RPAREN);
                } else {
                    e->type = ExprType::COLUMN_REF;
                    e->str_val = name;
                }
                return e;
            }
            case Tok::MUL:
                e->type = ExprType::COLUMN_REF;     // COUNT(*)
                e->str_val = "*";
                return e;
            case Tok::NUMBER:
                e->type = ExprType::LITERAL_INT;
                e->i64_val = std::stoll(t.text);
                return e;
            case Tok::DECIMAL:
                e->type = ExprType::LITERAL_DOUBLE;
                e->f64_val = std::stod(t.text);
                return e;
            case Tok::STRING_LITERAL:
                e->type = ExprType::LITERAL_STRING;
                e->str_val = t.text.substr(1, t.text.size() - 2);
                return e;
            default:
                throw std::runtime_error("Unexpected token in expression");
        }
    }
};

}  // namespace

SelectStmt parse_sql(const std::string& sql, const ParseOptions& opts) { return Parser(sql, opts).select(); }
SelectStmt parse_sql(const std::string& sql) { return parse_sql(sql, ParseOptions{}); }

// ---- logical plan (reference: src/logical/logical.cpp, src/logical/planner.cpp) ---------------------
static std::string join_names(const std::vector<std::string>& v) {
    std::string s;
    for (size_t i = 0; i < v.size(); ++i) s += (i ? ", " : "") + v[i];
    return s;
}

std::string LogicalOp::with_children(std::string head, int indent) const {
    std::string s = std::string(indent, ' ') + head;
    for (const auto& c : children) s += "\n" + c->to_string(indent + 2);
    return s;
}

std::string LogicalScan::to_string(int indent) const {
    return with_children("LogicalScan(table=" + table_name + ", cols=" + join_names(columns) + ")", indent);
}
std::string LogicalFilter::to_string(int indent) const {
    return with_children("LogicalFilter(" + predicate->to_string() + ")", indent);
}
std::string LogicalProject::to_string(int indent) const {
    std::string s;
    for (size_t i = 0; i < select_list.size(); ++i) {
        s += (i ? ", " : "") + select_list[i]->to_string();
        if (!aliases[i].empty()) s += " AS " + aliases[i];
    }
    return with_children("LogicalProject(" + s + ")", indent);
}
std::string LogicalHashJoin::to_string(int indent) const {
    std::string s = "LogicalHashJoin(left_keys=" + join_names(left_keys) + ", right_keys=" + join_names(right_keys);
    if (join_filter) s += ", filter=" + join_filter->to_string();
    return with_children(s + ")", indent);
}
std::string LogicalAggregate::to_string(int indent) const {
    std::string k, a;
    for (size_t i = 0; i < group_keys.size(); ++i) k += (i ? ", " : "") + group_keys[i]->to_string();
    for (size_t i = 0; i < aggregates.size(); ++i) {
        a += (i ? ", " : "") + aggregates[i].func_name + "(" + aggregates[i].arg->to_string() + ")";
        if (!aggregates[i].alias.empty()) a += " AS " + aggregates[i].alias;
    }
    return with_children("LogicalAggregate(keys=" + k + ", aggs=" + a + ")", indent);
}
std::string LogicalOrder::to_string(int indent) const {
    std::string s;
    for (size_t i = 0; i < order_by.size(); ++i)
        s += (i ? ", " : "") + order_by[i].expr->to_string() + (order_by[i].asc ? " ASC" : " DESC");
    return with_children("LogicalOrder(by: " + s + ")", indent);
}
std::string LogicalLimit::to_string(int indent) const {
    return with_children("LogicalLimit(" + std::to_string(limit) + ")", indent);
}

static void referenced_columns(const Expr* e, std::set<std::string>& out) {
    if (!e) return;
    if (e->type == ExprType::COLUMN_REF) out.insert(e->str_val);
    if (e->type == ExprType::BINARY_OP) {
        referenced_columns(e->left.get(), out);
        referenced_columns(e->right.get(), out);
    }
    if (e->type == ExprType::FUNC_CALL)
        for (const auto& a : e->args) referenced_columns(a.get(), out);
}

// src/logical/planner.cpp:108-165: Scan|Join -> Filter -> Aggregate -> Project -> Order -> Limit, no rewrites.
std::unique_ptr<LogicalOp> LogicalPlanner::build_logical_plan(const SelectStmt& stmt) {
    std::set<std::string> names;     // every scan receives the union of all referenced names (:27-56)
    for (const auto& it : stmt.select_list) referenced_columns(it.expr.get(), names);
    referenced_columns(stmt.where_clause.get(), names);
    for (const auto& j : stmt.joins) referenced_columns(j.on_condition.get(), names);
    for (const auto& g : stmt.group_by.columns) referenced_columns(g.get(), names);
    for (const auto& o : stmt.order_by) referenced_columns(o.expr.get(), names);
    const std::vector<std::string> columns(names.begin(), names.end());

    std::unique_ptr<LogicalOp> plan;
    if (stmt.joins.empty()) {
        plan = std::make_unique<LogicalScan>(stmt.from_table.table_name, columns);
    } else {
        const JoinItem& j = stmt.joins.front();      // only the first JOIN is planned (:65-67)
        std::vector<std::string> lk, rk;
        const Expr* on = j.on_condition.get();
        if (on->type == ExprType::BINARY_OP && on->op == BinaryOp::EQ && on->left->type == ExprType::COLUMN_REF &&
            on->right->type == ExprType::COLUMN_REF) {
            lk.push_back(on->left->str_val);       // left/right of the '=' — not matched to tables (:76-80)
            rk.push_back(on->right->str_val);
        }
        auto join = std::make_unique<LogicalHashJoin>(lk, rk);
        join->children.push_back(std::make_unique<LogicalScan>(stmt.from_table.table_name, columns));
        join->children.push_back(std::make_unique<LogicalScan>(j.table_ref.table_name, columns));
        plan = std::move(join);
    }
    auto stack = [&plan](std::unique_ptr<LogicalOp> op) {
        op->children.push_back(std::move(plan));
        plan = std::move(op);
    };
    if (stmt.where_clause) stack(std::make_unique<LogicalFilter>(stmt.where_clause->clone()));

    std::vector<LogicalAggregate::AggExpr> aggs;      // :94-106 (exact upper-case names only)
    for (const auto& it : stmt.select_list) {
        const Expr* e = it.expr.get();
        if (e->type == ExprType::FUNC_CALL && (e->func_name == "SUM" || e->func_name == "COUNT" || e->func_name == "AVG")) {
            if (e->args.empty()) throw std::runtime_error("aggregate without an argument");
            LogicalAggregate::AggExpr a;
            a.func_name = e->func_name;
            a.arg = e->args[0]->clone();
            a.alias = it.alias;
            aggs.push_back(std::move(a));
        }
    }
    if (!stmt.group_by.columns.empty() || !aggs.empty()) {
        std::vector<std::unique_ptr<Expr>> keys;
        for (const auto& g : stmt.group_by.columns) keys.push_back(g->clone());
        stack(std::make_unique<LogicalAggregate>(std::move(keys), std::move(aggs)));
    }
    std::vector<std::unique_ptr<Expr>> sel;
    std::vector<std::string> aliases;
    for (const auto& it : stmt.select_list) {
        sel.push_back(it.expr->clone());
        aliases.push_back(it.alias);
    }
    stack(std::make_unique<LogicalProject>(std::move(sel), std::move(aliases)));
    if (!stmt.order_by.empty()) {
        std::vector<LogicalOrder::OrderItem> items;
        for (const auto& o : stmt.order_by) items.push_back({o.expr->clone(), o.asc});
        stack(std::make_unique<LogicalOrder>(std::move(items)));
    }
    if (stmt.limit >= 0) stack(std::make_unique<LogicalLimit>(stmt.limit));
    return plan;
}

}  // namespace bosql
