// bq_groupby.cu — host side of k_group_tables (bq_groupby.cuh): sizing rule and the C entry points.
//
// HashAggregate over tens of millions of groups (src/exec/operator.cpp:984-1062): bq_partition orders the rows by
// key_hash, then one launch aggregates every partition in shared-memory tables and writes the finished output columns.
#include "bq_common.cuh"
#include "bq_internal.cuh"
#include "bq_groupby.cuh"

#include <cstring>

using namespace bq;

namespace {

constexpr unsigned kSlotsOneSum = 8192;      // 229.5 KB of shared memory with one sum array (or none)
constexpr unsigned kSlotsTwoSums = 4096;     // 196.6 KB with two
constexpr double kMaxLoad = 0.55;            // expected groups per table / slots
constexpr int kMaxSplits = 4;                // every split re-reads the partition's keys (from L2)

unsigned slots_for(int n_args) { return n_args > 1 ? kSlotsTwoSums : kSlotsOneSum; }

}  // namespace

extern "C" int bq_group_tables_plan(size_t ndv_hint, int n_args, int* log2_parts, int* splits) {
    if (!ndv_hint || n_args < 0 || n_args > 2) return 0;
    const double per_table = slots_for(n_args) * kMaxLoad;
    const double tables = static_cast<double>(ndv_hint) / per_table;
    int lp = 4;
    while (lp < 10 && static_cast<double>(1u << lp) < tables) ++lp;
    const double s = tables / static_cast<double>(1u << lp);
    int sp = 1;
    while (static_cast<double>(sp) < s) ++sp;
    if (sp > kMaxSplits) return 0;
    *log2_parts = lp;
    *splits = sp;
    return 1;
}

extern "C" int bq_partition_aggregate(bq_ctx* ctx, const bq_col* key, const bq_col* const* args, int n_args, const bq_col* offsets,
                                      int log2_parts, int splits, const bq_agg_out* outs, int n_out, bq_rel** out) {
    return guarded([&] {
        if (!key || !offsets) throw std::runtime_error("partition aggregate: key and offsets are required");
        if (key->type == BQ_DOUBLE) throw std::runtime_error("partition aggregate needs an integer key");
        if (n_args < 0 || n_args > 2) throw std::runtime_error("at most two aggregate arguments");
        if (log2_parts < 0 || log2_parts > 10) throw std::runtime_error("partition count must be 1 .. 1024 (a power of two)");
        if (splits < 1 || splits > 64) throw std::runtime_error("bad split count");
        if (n_out < 0 || n_out > BQ_MAX_AGG_OUT) throw std::runtime_error("too many aggregate outputs");
        const unsigned P = 1u << log2_parts;
        if (offsets->type != BQ_INT64 || offsets->n < static_cast<size_t>(P) + 1) throw std::runtime_error("offsets must be an INT64 column of 2^log2_parts + 1 entries");
        const size_t rows = key->n;
        GroupParams p{};
        p.key = key->ptr;
        p.key_kind = key->type;
        p.nv = n_args;
        for (int i = 0; i < n_args; ++i) {
            if (!args[i] || args[i]->n < rows) throw std::runtime_error("aggregate argument column shorter than the key column");
            if (args[i]->type == BQ_STRING) throw std::runtime_error("partition aggregate: arguments must be numeric columns");
            p.val[i] = args[i]->ptr;
            p.val_kind[i] = args[i]->type;
        }
        p.offsets = static_cast<const long long*>(offsets->ptr);
        p.splits = static_cast<unsigned>(splits);
        p.slots = slots_for(n_args);
        const unsigned long long tables = static_cast<unsigned long long>(P) * p.splits;
        unsigned long long cap = tables * (p.slots + 1ull);            // what the tables can hold ...
        if (cap > rows) cap = rows;                                    // ... and never more groups than rows
        p.capacity = cap;

        std::vector<bq_col*> cols;
        try {
            cols.push_back(new_col(ctx, key->type, cap));
            p.out_key = cols.back()->ptr;
            p.n_out = n_out;
            for (int o = 0; o < n_out; ++o) {
                int type = BQ_DOUBLE;
                if (outs[o].func == BQ_AGG_COUNT) type = BQ_INT64;
                else if (outs[o].func == BQ_AGG_SUM) type = outs[o].as_int ? BQ_INT64 : BQ_DOUBLE;
                else if (outs[o].func != BQ_AGG_AVG) throw std::runtime_error("unknown aggregate function");
                if (outs[o].func != BQ_AGG_COUNT && (outs[o].v < 0 || outs[o].v >= (n_args > 0 ? n_args : 1)))
                    throw std::runtime_error("bad aggregate argument index");
                cols.push_back(new_col(ctx, type, cap));
                p.func[o] = outs[o].func;
                p.v[o] = outs[o].func == BQ_AGG_COUNT ? 0 : outs[o].v;
                p.as_int[o] = outs[o].as_int;
                p.out[o] = cols.back()->ptr;
            }
            auto* d = static_cast<unsigned long long*>(scratch(ctx, 16));
            p.cursor = d;
            p.err = reinterpret_cast<int*>(d + 1);
            BQ_CUDA(cudaMemsetAsync(d, 0, 16, ctx->stream));
            size_t n_groups = 0;
            if (rows) {
                const size_t smem = group_smem_bytes(p.slots, n_args);
                BQ_CUDA(cudaFuncSetAttribute(k_group_tables, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
                k_group_tables<<<static_cast<unsigned>(tables), kGroupThreads, smem, ctx->stream>>>(p);
                ctx->launches++;
                BQ_CUDA(cudaGetLastError());
                auto* h = static_cast<unsigned long long*>(pinned(ctx, 16));
                BQ_CUDA(cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, ctx->stream));
                BQ_CUDA(cudaStreamSynchronize(ctx->stream));       // the one host round trip: group count + error word
                const int flags = static_cast<int>(h[1] & 0xFFFFFFFFull);
                if (flags & 2) throw std::runtime_error("group table overflow: a shared-memory table filled up");
                if (flags) throw std::runtime_error("internal: partition aggregate wrote past its output columns");
                n_groups = static_cast<size_t>(h[0]);
            }
            for (auto* c : cols) c->n = n_groups;
            auto* rel = new bq_rel();
            rel->cols = cols;
            rel->rows = n_groups;
            *out = rel;
        } catch (...) {
            for (auto* c : cols) free_col(c);
            throw;
        }
    });
}
