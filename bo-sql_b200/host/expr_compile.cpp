// expr_compile.cpp — see expr_compile.hpp.  Semantics follow src/exec/expression.cpp of the reference:
//   numeric_binary  :31-58   compare_values :60-120   is_truthy :10-22   evaluate_internal :153-206
#include "expr_compile.hpp"

#include <cmath>
#include <cstring>
#include <limits>

namespace bosql::gpu {

int64_t f64_key(double v) { return bq_f64_key(v); }

namespace {

bool is_arith(BinaryOp op) { return op == BinaryOp::ADD || op == BinaryOp::SUB || op == BinaryOp::MUL || op == BinaryOp::DIV; }
bool is_cmp(BinaryOp op) { return op <= BinaryOp::GE; }

struct Emitter {
    const ColumnLookup& cols;
    Dictionary* dict;
    Program prog;

    void emit(int op, int arg = 0) {
        bq_insn in{};
        in.op = op;
        in.arg = arg;
        in.imm.i = 0;
        prog.code.push_back(in);
    }
    void emit_i(int64_t v) {
        bq_insn in{};
        in.op = BQ_OP_IMM_I;
        in.imm.i = v;
        prog.code.push_back(in);
    }
    void emit_f(double v) {
        bq_insn in{};
        in.op = BQ_OP_IMM_F;
        in.imm.f = v;
        prog.code.push_back(in);
    }
    int column_slot(int lookup_index) {
        for (size_t i = 0; i < prog.columns.size(); ++i)
            if (prog.columns[i] == lookup_index) return static_cast<int>(i);
        prog.columns.push_back(lookup_index);
        return static_cast<int>(prog.columns.size()) - 1;
    }

    // leaves one value on the stack, returns its static type
    TypeId gen(const Expr* e) {
        switch (e->type) {
            case ExprType::COLUMN_REF: {
                int idx = cols.index_of(e->str_val);
                if (idx < 0) throw std::runtime_error("Unknown column: " + e->str_val);
                emit(BQ_OP_COL, column_slot(idx));
                return cols.type_of(idx);
            }
            case ExprType::LITERAL_INT: emit_i(e->i64_val); return TypeId::INT64;
            case ExprType::LITERAL_DOUBLE: emit_f(e->f64_val); return TypeId::DOUBLE;
            case ExprType::LITERAL_STRING:
                if (!dict) throw std::runtime_error("String literal without dictionary binding");
                emit_i(static_cast<int64_t>(dict->get_or_add(e->str_val)));    // may append (H8), once per plan
                return TypeId::STRING;
            case ExprType::FUNC_CALL:
                throw std::runtime_error("Function calls not supported in expression evaluation");
            case ExprType::BINARY_OP: break;
        }
        const TypeId l = gen(e->left.get());
        const TypeId r = gen(e->right.get());
        if (is_arith(e->op)) {
            if (l == TypeId::STRING || r == TypeId::STRING) throw std::runtime_error("Cannot coerce string to numeric");
            const bool fp = l == TypeId::DOUBLE || r == TypeId::DOUBLE;
            if (fp) {
                if (l != TypeId::DOUBLE) emit(BQ_OP_I2F_2);
                if (r != TypeId::DOUBLE) emit(BQ_OP_I2F);
            }
            static const int iop[] = {BQ_OP_ADD_I, BQ_OP_SUB_I, BQ_OP_MUL_I, BQ_OP_DIV_I};
            static const int fop[] = {BQ_OP_ADD_F, BQ_OP_SUB_F, BQ_OP_MUL_F, BQ_OP_DIV_F};
            const int k = static_cast<int>(e->op) - static_cast<int>(BinaryOp::ADD);
            emit(fp ? fop[k] : iop[k]);
            return fp ? TypeId::DOUBLE : TypeId::INT64;
        }
        if (is_cmp(e->op)) {
            const int k = static_cast<int>(e->op);   // EQ..GE = 0..5
            bool fp = false;
            switch (l) {                              // dispatch on the LEFT operand's type (:61)
                case TypeId::INT64:
                    if (r == TypeId::DOUBLE) emit(BQ_OP_F2I);                       // truncation, H6 (:64)
                    else if (r != TypeId::INT64) { emit(BQ_OP_ZX32); emit(BQ_OP_F2I); }   // reads f64_val of a 4-byte datum
                    break;
                case TypeId::DOUBLE:
                    fp = true;
                    if (r == TypeId::INT64) emit(BQ_OP_I2F);                        // (:79)
                    else if (r != TypeId::DOUBLE) { emit(BQ_OP_ZX32); emit(BQ_OP_I2F); }
                    break;
                case TypeId::DATE32:
                    emit(BQ_OP_SX32);                                               // low 32 bits of the right datum, H9 (:93-94)
                    break;
                case TypeId::STRING:
                    if (e->op != BinaryOp::EQ && e->op != BinaryOp::NE) throw std::runtime_error("Unsupported string comparison");
                    emit(BQ_OP_ZX32);
                    break;
            }
            emit((fp ? BQ_OP_EQ_F : BQ_OP_EQ_I) + k);
            return TypeId::INT64;
        }
        // AND / OR: both sides are always evaluated (H11); operands pass through is_truthy
        emit(l == TypeId::DOUBLE ? BQ_OP_TRUTHY_F_2 : BQ_OP_TRUTHY_I_2);
        emit(r == TypeId::DOUBLE ? BQ_OP_TRUTHY_F : BQ_OP_TRUTHY_I);
        emit(e->op == BinaryOp::AND ? BQ_OP_AND : BQ_OP_OR);
        return TypeId::INT64;
    }
};

const int64_t kI64Min = std::numeric_limits<int64_t>::min();
const int64_t kI64Max = std::numeric_limits<int64_t>::max();

bq_range make_range(int64_t lo, int64_t hi, bool neg) {
    bq_range r{};
    r.lo = lo;
    r.hi = hi;
    r.neg = neg ? 1 : 0;
    return r;
}
bq_range empty_range() { return make_range(1, 0, false); }

// k OP v on integer keys with domain [dmin, dmax]
bq_range range_for(BinaryOp op, int64_t v, int64_t dmin, int64_t dmax) {
    switch (op) {
        case BinaryOp::EQ: return make_range(v, v, false);
        case BinaryOp::NE: return make_range(v, v, true);
        case BinaryOp::LT: return v == kI64Min ? empty_range() : make_range(dmin, v - 1, false);
        case BinaryOp::LE: return make_range(dmin, v, false);
        case BinaryOp::GT: return v == kI64Max ? empty_range() : make_range(v + 1, dmax, false);
        case BinaryOp::GE: return make_range(v, dmax, false);
        default: throw std::runtime_error("Invalid comparison operator");
    }
}

// static_cast<int64_t>(double) as the reference's x86-64 build performs it
int64_t trunc_to_i64(double d) {
    if (!(d >= -9223372036854775808.0 && d < 9223372036854775808.0)) return kI64Min;
    return static_cast<int64_t>(d);
}

BinaryOp flip(BinaryOp op) {
    switch (op) {
        case BinaryOp::LT: return BinaryOp::GT;
        case BinaryOp::LE: return BinaryOp::GE;
        case BinaryOp::GT: return BinaryOp::LT;
        case BinaryOp::GE: return BinaryOp::LE;
        default: return op;
    }
}

}  // namespace

Program compile(const Expr* e, const ColumnLookup& cols, Dictionary* dict, bool as_predicate) {
    Emitter em{cols, dict, {}};
    TypeId t = em.gen(e);
    if (as_predicate) {
        em.emit(t == TypeId::DOUBLE ? BQ_OP_TRUTHY_F : BQ_OP_TRUTHY_I);
        t = TypeId::INT64;
    }
    em.prog.result = t;
    if (em.prog.code.size() > BQ_MAX_PROGRAM) throw std::runtime_error("expression too large for the device evaluator");
    if (em.prog.columns.size() > BQ_MAX_PROGRAM_COLS) throw std::runtime_error("expression references too many columns");
    return std::move(em.prog);
}

TypeId value_type(const Expr* e, const ColumnLookup& cols) {
    switch (e->type) {
        case ExprType::COLUMN_REF: {
            int idx = cols.index_of(e->str_val);
            if (idx < 0) throw std::runtime_error("Unknown column: " + e->str_val);
            return cols.type_of(idx);
        }
        case ExprType::LITERAL_INT: return TypeId::INT64;
        case ExprType::LITERAL_DOUBLE: return TypeId::DOUBLE;
        case ExprType::LITERAL_STRING: return TypeId::STRING;
        case ExprType::FUNC_CALL: throw std::runtime_error("Function calls not supported in expression evaluation");
        case ExprType::BINARY_OP:
            if (is_arith(e->op)) {
                TypeId l = value_type(e->left.get(), cols), r = value_type(e->right.get(), cols);
                return (l == TypeId::DOUBLE || r == TypeId::DOUBLE) ? TypeId::DOUBLE : TypeId::INT64;
            }
            return TypeId::INT64;
    }
    return TypeId::INT64;
}

bool may_throw_per_row(const Expr* e, const ColumnLookup& cols) {
    if (!e || e->type != ExprType::BINARY_OP) return false;
    if (may_throw_per_row(e->left.get(), cols) || may_throw_per_row(e->right.get(), cols)) return true;
    if (e->op != BinaryOp::DIV) return false;
    if (value_type(e->left.get(), cols) == TypeId::DOUBLE || value_type(e->right.get(), cols) == TypeId::DOUBLE) return false;   // +inf
    return !(e->right->type == ExprType::LITERAL_INT && e->right->i64_val != 0);
}

void referenced(const Expr* e, const ColumnLookup& cols, std::vector<int>& out) {
    if (!e) return;
    if (e->type == ExprType::COLUMN_REF) {
        int idx = cols.index_of(e->str_val);
        if (idx < 0) throw std::runtime_error("Unknown column: " + e->str_val);
        if (std::find(out.begin(), out.end(), idx) == out.end()) out.push_back(idx);
    } else if (e->type == ExprType::BINARY_OP) {
        referenced(e->left.get(), cols, out);
        referenced(e->right.get(), cols, out);
    } else if (e->type == ExprType::FUNC_CALL) {
        for (const auto& a : e->args) referenced(a.get(), cols, out);
    }
}

bool to_range(const Expr* e, const ColumnLookup& cols, Dictionary* dict, ColumnRange& out) {
    // a bare column as a truth value: is_truthy = (value != 0)   (:10-22; for StrId: id != 0, H7)
    if (e->type == ExprType::COLUMN_REF) {
        int idx = cols.index_of(e->str_val);
        if (idx < 0) throw std::runtime_error("Unknown column: " + e->str_val);
        out.column = idx;
        out.range = make_range(0, 0, true);     // key(0.0) == 0 as well
        return true;
    }
    if (e->type != ExprType::BINARY_OP || !is_cmp(e->op)) return false;
    const Expr* l = e->left.get();
    const Expr* r = e->right.get();
    BinaryOp op = e->op;
    const bool lit_r = r->type == ExprType::LITERAL_INT || r->type == ExprType::LITERAL_DOUBLE || r->type == ExprType::LITERAL_STRING;
    if (!(l->type == ExprType::COLUMN_REF && lit_r)) {
        // `int literal OP int column` is a pure integer compare either way round
        if (l->type == ExprType::LITERAL_INT && r->type == ExprType::COLUMN_REF) {
            int idx = cols.index_of(r->str_val);
            if (idx < 0) throw std::runtime_error("Unknown column: " + r->str_val);
            if (cols.type_of(idx) != TypeId::INT64) return false;
            out.column = idx;
            out.range = range_for(flip(op), l->i64_val, kI64Min, kI64Max);
            return true;
        }
        return false;
    }
    int idx = cols.index_of(l->str_val);
    if (idx < 0) throw std::runtime_error("Unknown column: " + l->str_val);
    out.column = idx;
    switch (cols.type_of(idx)) {
        case TypeId::INT64: {
            int64_t v;
            if (r->type == ExprType::LITERAL_INT) v = r->i64_val;
            else if (r->type == ExprType::LITERAL_DOUBLE) v = trunc_to_i64(r->f64_val);   // H6
            else return false;
            out.range = range_for(op, v, kI64Min, kI64Max);
            return true;
        }
        case TypeId::DOUBLE: {
            double v;
            if (r->type == ExprType::LITERAL_INT) v = static_cast<double>(r->i64_val);
            else if (r->type == ExprType::LITERAL_DOUBLE) v = r->f64_val;
            else return false;
            if (std::isnan(v)) {      // every comparison with NaN is false except !=
                out.range = (op == BinaryOp::NE) ? make_range(1, 0, true) : empty_range();
                return true;
            }
            out.range = range_for(op, f64_key(v), f64_key(-std::numeric_limits<double>::infinity()),
                                  f64_key(std::numeric_limits<double>::infinity()));
            return true;
        }
        case TypeId::DATE32: {
            if (r->type != ExprType::LITERAL_INT) return false;
            int64_t v = static_cast<int64_t>(static_cast<int32_t>(static_cast<uint64_t>(r->i64_val)));   // low 32 bits, H9
            out.range = range_for(op, v, std::numeric_limits<int32_t>::min(), std::numeric_limits<int32_t>::max());
            return true;
        }
        case TypeId::STRING: {
            if (op != BinaryOp::EQ && op != BinaryOp::NE) throw std::runtime_error("Unsupported string comparison");
            int64_t v;
            if (r->type == ExprType::LITERAL_STRING) {
                if (!dict) throw std::runtime_error("String literal without dictionary binding");
                v = static_cast<int64_t>(dict->get_or_add(r->str_val));      // unknown literal: appended, matches nothing (H8)
            } else if (r->type == ExprType::LITERAL_INT) {
                v = static_cast<int64_t>(static_cast<uint32_t>(static_cast<uint64_t>(r->i64_val)));
            } else {
                return false;
            }
            out.range = range_for(op, v, 0, 0xFFFFFFFFLL);
            return true;
        }
    }
    return false;
}

}  // namespace bosql::gpu
