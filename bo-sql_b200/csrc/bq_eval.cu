// bq_eval.cu — typed expression programs: the general evaluator.
//
// evaluate_internal (src/exec/expression.cpp:153-206) walks the Expr tree per row, boxing every value in a
// tagged Datum and looking columns up by name in an unordered_map.  The host compiler
// (bo-sql_b200/host/expr_compile.cpp) resolves names, literal dictionary ids and every operand type once
// per plan and emits a postfix program over 8-byte slots; this kernel runs it for one row per thread with
// a register/local stack.  Control flow is uniform across the warp (same program for every row).
//
// It serves Project (src/exec/operator.cpp:498-555) and any predicate / aggregate argument / group key
// that the fused range form of bq_scan.cu cannot express (OR, column-vs-column comparisons, arithmetic
// inside predicates ...): those are materialised into a column (a 0/1 INT64 mask for predicates) and
// handed to the fused kernel.
#include "bq_common.cuh"
#include "bq_internal.cuh"

namespace bq {

struct EvalParams {
    bq_insn prog[BQ_MAX_PROGRAM];
    int n_insn;
    const void* cols[BQ_MAX_PROGRAM_COLS];
    int kinds[BQ_MAX_PROGRAM_COLS];
    size_t row_begin, n;
    int out_type;
    void* out;
    int* err;
};

constexpr int kStack = 16;

__global__ void __launch_bounds__(kBlock) k_eval(const __grid_constant__ EvalParams p) {
    int err = 0;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < p.n; t += (size_t)gridDim.x * blockDim.x) {
        const size_t row = p.row_begin + t;
        long long st[kStack];
        int sp = 0;
        for (int pc = 0; pc < p.n_insn; ++pc) {
            const bq_insn& in = p.prog[pc];
            switch (in.op) {
                case BQ_OP_COL: st[sp++] = load_raw(p.cols[in.arg], p.kinds[in.arg], row); break;
                case BQ_OP_IMM_I: st[sp++] = in.imm.i; break;
                case BQ_OP_IMM_F: st[sp++] = __double_as_longlong(in.imm.f); break;
                case BQ_OP_I2F: st[sp - 1] = __double_as_longlong(static_cast<double>(st[sp - 1])); break;
                case BQ_OP_I2F_2: st[sp - 2] = __double_as_longlong(static_cast<double>(st[sp - 2])); break;
                case BQ_OP_F2I: {
                    // static_cast<int64_t>(double) as x86-64 cvttsd2si does it: out-of-range and NaN give INT64_MIN
                    double d = __longlong_as_double(st[sp - 1]);
                    long long v;
                    if (!(d >= -9223372036854775808.0 && d < 9223372036854775808.0)) v = INT64_MIN;
                    else v = static_cast<long long>(d);
                    st[sp - 1] = v;
                    break;
                }
                case BQ_OP_SX32: st[sp - 1] = static_cast<long long>(static_cast<int>(st[sp - 1])); break;
                case BQ_OP_ZX32: st[sp - 1] = static_cast<long long>(static_cast<unsigned>(st[sp - 1])); break;
                case BQ_OP_ADD_I: st[sp - 2] = static_cast<long long>(static_cast<unsigned long long>(st[sp - 2]) + static_cast<unsigned long long>(st[sp - 1])); --sp; break;
                case BQ_OP_SUB_I: st[sp - 2] = static_cast<long long>(static_cast<unsigned long long>(st[sp - 2]) - static_cast<unsigned long long>(st[sp - 1])); --sp; break;
                case BQ_OP_MUL_I: st[sp - 2] = static_cast<long long>(static_cast<unsigned long long>(st[sp - 2]) * static_cast<unsigned long long>(st[sp - 1])); --sp; break;
                case BQ_OP_DIV_I: {
                    long long l = st[sp - 2], r = st[sp - 1];
                    long long z = 0;
                    if (r == 0) err |= 1;                       // "Division by zero", src/exec/expression.cpp:52
                    else if (l == INT64_MIN && r == -1) z = INT64_MIN;
                    else z = l / r;
                    st[sp - 2] = z;
                    --sp;
                    break;
                }
                case BQ_OP_ADD_F: st[sp - 2] = __double_as_longlong(__dadd_rn(__longlong_as_double(st[sp - 2]), __longlong_as_double(st[sp - 1]))); --sp; break;
                case BQ_OP_SUB_F: st[sp - 2] = __double_as_longlong(__dsub_rn(__longlong_as_double(st[sp - 2]), __longlong_as_double(st[sp - 1]))); --sp; break;
                case BQ_OP_MUL_F: st[sp - 2] = __double_as_longlong(__dmul_rn(__longlong_as_double(st[sp - 2]), __longlong_as_double(st[sp - 1]))); --sp; break;
                case BQ_OP_DIV_F: {
                    double l = __longlong_as_double(st[sp - 2]), r = __longlong_as_double(st[sp - 1]);
                    double z = r == 0.0 ? __longlong_as_double(0x7FF0000000000000LL) : __ddiv_rn(l, r);   // :41
                    st[sp - 2] = __double_as_longlong(z);
                    --sp;
                    break;
                }
                case BQ_OP_EQ_I: st[sp - 2] = st[sp - 2] == st[sp - 1]; --sp; break;
                case BQ_OP_NE_I: st[sp - 2] = st[sp - 2] != st[sp - 1]; --sp; break;
                case BQ_OP_LT_I: st[sp - 2] = st[sp - 2] < st[sp - 1]; --sp; break;
                case BQ_OP_LE_I: st[sp - 2] = st[sp - 2] <= st[sp - 1]; --sp; break;
                case BQ_OP_GT_I: st[sp - 2] = st[sp - 2] > st[sp - 1]; --sp; break;
                case BQ_OP_GE_I: st[sp - 2] = st[sp - 2] >= st[sp - 1]; --sp; break;
                case BQ_OP_EQ_F: st[sp - 2] = __longlong_as_double(st[sp - 2]) == __longlong_as_double(st[sp - 1]); --sp; break;
                case BQ_OP_NE_F: st[sp - 2] = __longlong_as_double(st[sp - 2]) != __longlong_as_double(st[sp - 1]); --sp; break;
                case BQ_OP_LT_F: st[sp - 2] = __longlong_as_double(st[sp - 2]) < __longlong_as_double(st[sp - 1]); --sp; break;
                case BQ_OP_LE_F: st[sp - 2] = __longlong_as_double(st[sp - 2]) <= __longlong_as_double(st[sp - 1]); --sp; break;
                case BQ_OP_GT_F: st[sp - 2] = __longlong_as_double(st[sp - 2]) > __longlong_as_double(st[sp - 1]); --sp; break;
                case BQ_OP_GE_F: st[sp - 2] = __longlong_as_double(st[sp - 2]) >= __longlong_as_double(st[sp - 1]); --sp; break;
                case BQ_OP_TRUTHY_I: st[sp - 1] = st[sp - 1] != 0; break;
                case BQ_OP_TRUTHY_F: st[sp - 1] = __longlong_as_double(st[sp - 1]) != 0.0; break;
                case BQ_OP_TRUTHY_I_2: st[sp - 2] = st[sp - 2] != 0; break;
                case BQ_OP_TRUTHY_F_2: st[sp - 2] = __longlong_as_double(st[sp - 2]) != 0.0; break;
                case BQ_OP_AND: st[sp - 2] = (st[sp - 2] != 0) & (st[sp - 1] != 0); --sp; break;
                case BQ_OP_OR: st[sp - 2] = (st[sp - 2] != 0) | (st[sp - 1] != 0); --sp; break;
                default: break;
            }
        }
        long long v = st[0];
        switch (p.out_type) {
            case BQ_INT64:
            case BQ_DOUBLE: static_cast<long long*>(p.out)[t] = v; break;
            case BQ_STRING: static_cast<unsigned*>(p.out)[t] = static_cast<unsigned>(v); break;
            default: static_cast<int*>(p.out)[t] = static_cast<int>(v); break;
        }
    }
    if (err) atomicOr(p.err, err);
}

// Static check of a program: stack discipline, column references, opcode range.
static void validate(const bq_insn* prog, int n, int n_cols) {
    if (n <= 0 || n > BQ_MAX_PROGRAM) throw std::runtime_error("expression program too long");
    int sp = 0;
    for (int i = 0; i < n; ++i) {
        int op = prog[i].op;
        int need = 0, delta = 0;
        if (op == BQ_OP_COL) {
            if (prog[i].arg < 0 || prog[i].arg >= n_cols) throw std::runtime_error("expression program reads a missing column");
            delta = 1;
        } else if (op == BQ_OP_IMM_I || op == BQ_OP_IMM_F) {
            delta = 1;
        } else if (op == BQ_OP_I2F || op == BQ_OP_F2I || op == BQ_OP_SX32 || op == BQ_OP_ZX32 || op == BQ_OP_TRUTHY_I ||
                   op == BQ_OP_TRUTHY_F) {
            need = 1;
        } else if (op == BQ_OP_I2F_2 || op == BQ_OP_TRUTHY_I_2 || op == BQ_OP_TRUTHY_F_2) {
            need = 2;
        } else if ((op >= BQ_OP_ADD_I && op <= BQ_OP_GE_F) || op == BQ_OP_AND || op == BQ_OP_OR) {
            need = 2;
            delta = -1;
        } else {
            throw std::runtime_error("unknown opcode in expression program");
        }
        if (sp < need) throw std::runtime_error("expression program underflows its stack");
        sp += delta;
        if (sp > kStack) throw std::runtime_error("expression program overflows its stack");
    }
    if (sp != 1) throw std::runtime_error("expression program must leave exactly one value");
}

}  // namespace bq

using namespace bq;

extern "C" int bq_eval(bq_ctx* ctx, const bq_insn* prog, int n_insn, const bq_col* const* cols, int n_cols,
                       size_t row_begin, size_t row_end, int out_type, bq_col** out) {
    return guarded([&] {
        if (n_cols < 0 || n_cols > BQ_MAX_PROGRAM_COLS) throw std::runtime_error("too many program columns");
        if (row_end < row_begin) throw std::runtime_error("bad row range");
        validate(prog, n_insn, n_cols);
        EvalParams p{};
        for (int i = 0; i < n_insn; ++i) p.prog[i] = prog[i];
        p.n_insn = n_insn;
        for (int c = 0; c < n_cols; ++c) {
            if (cols[c]->n < row_end) throw std::runtime_error("program column shorter than the row range");
            p.cols[c] = cols[c]->ptr;
            p.kinds[c] = cols[c]->type;
        }
        p.row_begin = row_begin;
        p.n = row_end - row_begin;
        p.out_type = out_type;
        bq_col* o = new_col(ctx, out_type, p.n);
        try {
            p.out = o->ptr;
            auto* d = static_cast<int*>(scratch(ctx, 16));
            BQ_CUDA(cudaMemsetAsync(d, 0, 4, ctx->stream));
            p.err = d;
            if (p.n) {
                k_eval<<<grid_for(ctx, p.n, 8), kBlock, 0, ctx->stream>>>(p);
                ctx->launches++;
                BQ_CUDA(cudaGetLastError());
            }
            auto* h = static_cast<int*>(pinned(ctx, 8));
            BQ_CUDA(cudaMemcpyAsync(h, d, 4, cudaMemcpyDeviceToHost, ctx->stream));
            BQ_CUDA(cudaStreamSynchronize(ctx->stream));
            if (*h & 1) throw std::runtime_error("Division by zero");
        } catch (...) {
            free_col(o);
            throw;
        }
        *out = o;
    });
}
