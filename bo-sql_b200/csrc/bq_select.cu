// bq_select.cu — selection vectors, gathers and slices.
//
// Selection::next builds `selected` row by row and then gathers every column (copy_selected)
// per 4096-row batch (src/exec/operator.cpp:403-429, 11-49).  Here one kernel evaluates the predicate
// for the whole row range and writes one ballot word per warp; bq_compact.cu turns the bit vector into
// ascending row ids (stable: output stays in scan order); gathers are one coalesced kernel per column.
#include "bq_common.cuh"
#include "bq_internal.cuh"

namespace bq {

struct PredParams {
    DSlot s[4];
    int n_slots;
    const long long* mask;
    size_t row_begin, row_end;
};

__global__ void __launch_bounds__(kBlock) k_pred_bits(const __grid_constant__ PredParams p, unsigned* __restrict__ bits) {
    const size_t n = p.row_end - p.row_begin;
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    bool ok = t < n;
    if (ok) {
        const size_t i = p.row_begin + t;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            if (s < p.n_slots && p.s[s].ptr) ok = ok && slot_pass(p.s[s], load_raw(p.s[s].ptr, p.s[s].kind, i));
        }
        if (p.mask && __ldg(p.mask + i) == 0) ok = false;
    }
    unsigned b = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && (t >> 5) < (n + 31) / 32) bits[t >> 5] = b;
}

template <typename T>
__global__ void __launch_bounds__(kBlock) k_gather(const T* __restrict__ src, const unsigned* __restrict__ ids, size_t n,
                                                   T* __restrict__ dst) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = __ldg(src + ids[i]);
}

}  // namespace bq

using namespace bq;

extern "C" {

int bq_select(bq_ctx* ctx, const bq_select_spec* spec, bq_col** out_rowids) {
    return guarded([&] {
        if (spec->row_end < spec->row_begin) throw std::runtime_error("bad row range");
        if (spec->row_end > 0xFFFFFFFFull) throw std::runtime_error("row ids are 32-bit: at most 2^32 rows per scan");
        PredParams p{};
        p.n_slots = 4;
        for (int s = 0; s < 4; ++s) {
            if (spec->pred[s].from_build) throw std::runtime_error("bq_select has no build side");
            p.s[s] = make_dslot(spec->pred[s], spec->row_end, "pred");
        }
        if (spec->mask) {
            if (spec->mask->type != BQ_INT64 || spec->mask->n < spec->row_end) throw std::runtime_error("mask must be an INT64 column covering the row range");
            p.mask = static_cast<const long long*>(spec->mask->ptr);
        }
        p.row_begin = spec->row_begin;
        p.row_end = spec->row_end;
        size_t n = spec->row_end - spec->row_begin;
        if (n == 0) {
            *out_rowids = new_col(ctx, BQ_STRING, 0);
            return;
        }
        size_t n_words = (n + 31) / 32;
        DevBuf bits_buf(ctx, n_words * 4 + 4);
        auto* bits = bits_buf.as<unsigned>();
        k_pred_bits<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, ctx->stream>>>(p, bits);
        ctx->launches++;
        BQ_CUDA(cudaGetLastError());
        compact_bits(ctx, bits, n, static_cast<unsigned>(spec->row_begin), out_rowids);
    });
}

int bq_gather(bq_ctx* ctx, const bq_col* col, const bq_col* rowids, bq_col** out) {
    return guarded([&] {
        if (rowids->type != BQ_STRING) throw std::runtime_error("row ids must be a uint32 column");
        size_t n = rowids->n;
        bq_col* o = new_col(ctx, col->type, n);
        if (n) {
            int grid = grid_for(ctx, n, 8);
            const auto* ids = static_cast<const unsigned*>(rowids->ptr);
            if (width_of(col->type) == 8)
                k_gather<long long><<<grid, kBlock, 0, ctx->stream>>>(static_cast<const long long*>(col->ptr), ids, n,
                                                                     static_cast<long long*>(o->ptr));
            else
                k_gather<int><<<grid, kBlock, 0, ctx->stream>>>(static_cast<const int*>(col->ptr), ids, n,
                                                               static_cast<int*>(o->ptr));
            ctx->launches++;
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) {
                free_col(o);
                throw std::runtime_error(cudaGetErrorString(e));
            }
        }
        *out = o;
    });
}

int bq_slice(bq_ctx* ctx, const bq_col* col, size_t begin, size_t end, bq_col** out) {
    return guarded([&] {
        if (end < begin || end > col->n) throw std::runtime_error("bad slice");
        size_t n = end - begin, w = width_of(col->type);
        bq_col* o = new_col(ctx, col->type, n);
        if (n) {
            cudaError_t e = cudaMemcpyAsync(o->ptr, static_cast<const char*>(col->ptr) + begin * w, n * w,
                                            cudaMemcpyDeviceToDevice, ctx->stream);
            if (e != cudaSuccess) {
                free_col(o);
                throw std::runtime_error(cudaGetErrorString(e));
            }
        }
        *out = o;
    });
}

}  // extern "C"
