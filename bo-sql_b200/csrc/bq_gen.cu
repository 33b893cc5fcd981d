// bq_gen.cu — deterministic synthetic columns generated directly in HBM.
//
// value(row) depends only on (seed, stream, global row index) through a counter-based hash, so a
// 1 B-row table is produced on the device at memory speed while any slice can be regenerated on the
// host for the oracle (oracle/datagen.py restates exactly this arithmetic in numpy; the test
// tests/test_datagen.py compares the two bit for bit).
#include "bq_common.cuh"

namespace bq {

struct GenParams {
    int dist;
    int type;
    uint64_t seed, stream;
    int64_t lo;
    uint64_t range;      // hi - lo + 1
    double div;
    int base_year, n_years;
    const uint64_t* cdf;
    size_t n_cdf;
    const uint64_t* starts;
    uint64_t modulus;
    uint64_t row0;
    size_t n;
};

BQ_D int64_t gen_value(const GenParams& p, uint64_t row, double* as_f) {
    uint64_t h = row_hash(p.seed, p.stream, row);
    int64_t v = 0;
    switch (p.dist) {
        case BQ_GEN_SEQ: v = p.lo + static_cast<int64_t>(row * p.modulus); break;             // modulus = stride (1 unless given)
        case BQ_GEN_UNIFORM: v = p.lo + static_cast<int64_t>((h % p.range) * p.modulus); break;
        case BQ_GEN_UNIFORM_DIV:
            v = p.lo + static_cast<int64_t>(h % p.range);
            *as_f = static_cast<double>(v) / p.div;
            break;
        case BQ_GEN_DATE: {
            uint64_t h2 = mix64(h);
            int64_t y = p.base_year + static_cast<int64_t>(h % static_cast<uint64_t>(p.n_years));
            int64_t m = 1 + static_cast<int64_t>(h2 % 12ULL);
            int64_t d = 1 + static_cast<int64_t>((h2 >> 32) % 28ULL);
            v = y * 10000 + m * 100 + d;
            break;
        }
        case BQ_GEN_BUCKETS:
        case BQ_GEN_TABLE: {
            uint64_t u = h >> 11;   // 53 uniform bits
            size_t lo = 0, hi = p.n_cdf;   // first i with cdf[i] > u
            while (lo < hi) {
                size_t mid = (lo + hi) >> 1;
                if (__ldg(p.cdf + mid) > u) hi = mid; else lo = mid + 1;
            }
            if (lo >= p.n_cdf) lo = p.n_cdf - 1;
            if (p.dist == BQ_GEN_BUCKETS) {
                const uint64_t s0 = __ldg(p.starts + lo), s1 = __ldg(p.starts + lo + 1);
                v = p.lo + static_cast<int64_t>(s0 + mix64(h) % (s1 - s0));
            } else {
                v = p.lo + static_cast<int64_t>(lo);
            }
            break;
        }
        case BQ_GEN_HASHED: {
            uint64_t id = h % p.range;
            v = p.lo + static_cast<int64_t>(mix64(id ^ (p.seed * 0x2545F4914F6CDD1DULL)) % p.modulus);
            break;
        }
    }
    return v;
}

__global__ void __launch_bounds__(kBlock) k_generate(GenParams p, void* __restrict__ out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < p.n; i += (size_t)gridDim.x * blockDim.x) {
        double f = 0.0;
        int64_t v = gen_value(p, p.row0 + i, &f);
        switch (p.type) {
            case BQ_INT64: static_cast<long long*>(out)[i] = v; break;
            case BQ_DOUBLE: static_cast<double*>(out)[i] = (p.dist == BQ_GEN_UNIFORM_DIV) ? f : static_cast<double>(v); break;
            case BQ_STRING: static_cast<unsigned*>(out)[i] = static_cast<unsigned>(v); break;
            default: static_cast<int*>(out)[i] = static_cast<int>(v); break;
        }
    }
}

}  // namespace bq

using namespace bq;

extern "C" int bq_col_generate(bq_ctx* ctx, bq_col* col, const bq_gen_spec* s, uint64_t global_row0) {
    return guarded([&] {
        GenParams p{};
        p.dist = s->dist;
        p.type = col->type;
        p.seed = s->seed;
        p.stream = s->stream;
        p.lo = s->lo;
        p.range = static_cast<uint64_t>(s->hi - s->lo) + 1ULL;
        p.div = s->div;
        p.base_year = s->base_year;
        p.n_years = s->n_years > 0 ? s->n_years : 1;
        p.modulus = s->modulus ? s->modulus : 1;
        p.row0 = global_row0;
        p.n = col->n;
        if (s->dist < BQ_GEN_SEQ || s->dist > BQ_GEN_BUCKETS) throw std::runtime_error("bad generator dist");
        const bool table = s->dist == BQ_GEN_TABLE || s->dist == BQ_GEN_BUCKETS;
        if (s->dist != BQ_GEN_SEQ && s->dist != BQ_GEN_DATE && !table && s->hi < s->lo)
            throw std::runtime_error("generator needs lo <= hi");
        uint64_t* d_cdf = nullptr;
        if (table) {
            if (!s->cdf || !s->n_cdf) throw std::runtime_error("BQ_GEN_TABLE needs a cdf");
            const bool buckets = s->dist == BQ_GEN_BUCKETS;
            if (buckets && !s->starts) throw std::runtime_error("BQ_GEN_BUCKETS needs bucket starts");
            if (buckets)
                for (size_t i = 0; i < s->n_cdf; ++i)
                    if (s->starts[i + 1] <= s->starts[i]) throw std::runtime_error("BQ_GEN_BUCKETS: bucket starts must ascend");
            d_cdf = static_cast<uint64_t*>(dev_alloc(ctx, (2 * s->n_cdf + 1) * 8));
            BQ_CUDA(cudaMemcpyAsync(d_cdf, s->cdf, s->n_cdf * 8, cudaMemcpyHostToDevice, ctx->stream));
            if (buckets) BQ_CUDA(cudaMemcpyAsync(d_cdf + s->n_cdf, s->starts, (s->n_cdf + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
            p.cdf = d_cdf;
            p.n_cdf = s->n_cdf;
            p.starts = d_cdf + s->n_cdf;
        }
        if (col->n) {
            int grid = grid_for(ctx, col->n, 8);
            k_generate<<<grid, kBlock, 0, ctx->stream>>>(p, col->ptr);
            ctx->launches++;
            BQ_CUDA(cudaGetLastError());
        }
        if (d_cdf) dev_free(ctx, d_cdf);
        col->has_minmax = false;
    });
}
