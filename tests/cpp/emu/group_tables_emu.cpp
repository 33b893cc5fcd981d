// group_tables_emu.cpp — k_group_tables (bo-sql_b200/csrc/bq_groupby.cuh), the very source the GPU runs, compiled for
// the host with the fibre-based CUDA emulation and checked against std::map.  TEST INFRASTRUCTURE (-m "not gpu").
//
// What the checker follows: HashAggregate's accumulate and emit phases, src/exec/operator.cpp:984-1014 (count += 1,
// sum += datum_as_double(value)) and :1030-1050 (COUNT -> int64, SUM -> double or static_cast<int64_t>, AVG -> sum / count).
// Values are dyadic (k / 8), so every order of addition gives the same bits and the comparison is exact.
#include "cuda_emu.hpp"

#define BQ_GROUPBY_FN inline
#include "../../../bo-sql_b200/csrc/bq_groupby.cuh"

#include <algorithm>
#include <map>
#include <random>
#include <string>

namespace {

// key_hash of bq_common.cuh (murmur3's 64-bit finaliser): bq_partition orders rows by its top bits
uint64_t key_hash(uint64_t k) {
    k ^= k >> 33;
    k *= 0xFF51AFD7ED558CCDULL;
    k ^= k >> 33;
    k *= 0xC4CEB9FE1A85EC53ULL;
    k ^= k >> 33;
    return k;
}

struct Agg {
    long long count = 0;
    double s0 = 0.0, s1 = 0.0;
};

struct Case {
    std::string name;
    size_t rows;
    long long ids;          // distinct keys (before the special ones)
    int log2p, splits;
    unsigned slots;
    int key_kind;           // BQ_INT64 or BQ_DATE32
    int nv;                 // aggregate arguments: 0, 1 (DOUBLE) or 2 (INT64, DOUBLE)
    double heavy;           // fraction of the rows that carry ONE key (tag conflicts, atomic phase)
    bool expect_overflow;
};

int failures = 0;
void expect(bool ok, const std::string& what) {
    if (!ok) {
        std::fprintf(stderr, "FAILED: %s\n", what.c_str());
        ++failures;
    }
}

void run(const Case& c) {
    std::mt19937_64 rng(1234 + c.rows * 31 + c.ids);
    std::vector<long long> key(c.rows);
    std::vector<long long> a(c.rows);
    std::vector<double> b(c.rows);
    for (size_t i = 0; i < c.rows; ++i) {
        long long k = static_cast<long long>(rng() % static_cast<uint64_t>(c.ids));
        if (c.key_kind == BQ_INT64) {
            k = k * 7919 - 1000000007LL * (k % 3);             // spread, some negative
            if (i % 977 == 5) k = INT64_MIN;                   // the key that equals the table's empty marker
            if (i % 1013 == 7) k = INT64_MAX;
        } else {
            k = 20200101 + k;
        }
        if (c.heavy > 0 && (rng() % 1000) < c.heavy * 1000) k = c.key_kind == BQ_INT64 ? 42 : 20240229;
        key[i] = k;
        a[i] = static_cast<long long>(rng() % 2001) - 1000;
        b[i] = static_cast<double>(static_cast<long long>(rng() % 51201) - 25600) / 8.0;
    }
    // what bq_partition leaves: rows ordered by the top log2p bits of key_hash, and the offsets
    const unsigned P = 1u << c.log2p;
    auto part_of = [&](long long k) { return c.log2p ? static_cast<unsigned>(key_hash(static_cast<uint64_t>(k)) >> (64 - c.log2p)) : 0u; };
    std::vector<long long> offsets(P + 1, 0);
    for (size_t i = 0; i < c.rows; ++i) offsets[part_of(key[i]) + 1]++;
    for (unsigned q = 0; q < P; ++q) offsets[q + 1] += offsets[q];
    std::vector<long long> pk(c.rows), pa(c.rows);
    std::vector<int> pk32(c.rows);
    std::vector<double> pb(c.rows);
    {
        std::vector<long long> cur(offsets.begin(), offsets.end() - 1);
        for (size_t i = 0; i < c.rows; ++i) {
            const size_t at = static_cast<size_t>(cur[part_of(key[i])]++);
            pk[at] = key[i];
            pk32[at] = static_cast<int>(key[i]);
            pa[at] = a[i];
            pb[at] = b[i];
        }
    }
    std::map<long long, Agg> want;
    for (size_t i = 0; i < c.rows; ++i) {
        Agg& g = want[key[i]];
        g.count += 1;
        if (c.nv == 1) g.s0 += b[i];
        if (c.nv == 2) {
            g.s0 += static_cast<double>(a[i]);
            g.s1 += b[i];
        }
    }

    bq::GroupParams p{};
    p.key = c.key_kind == BQ_INT64 ? static_cast<const void*>(pk.data()) : static_cast<const void*>(pk32.data());
    p.key_kind = c.key_kind;
    p.nv = c.nv;
    if (c.nv == 1) {
        p.val[0] = pb.data();
        p.val_kind[0] = BQ_DOUBLE;
    } else if (c.nv == 2) {
        p.val[0] = pa.data();
        p.val_kind[0] = BQ_INT64;
        p.val[1] = pb.data();
        p.val_kind[1] = BQ_DOUBLE;
    }
    p.offsets = offsets.data();
    p.splits = static_cast<unsigned>(c.splits);
    p.slots = c.slots;
    const size_t cap = std::min<size_t>(static_cast<size_t>(P) * c.splits * (c.slots + 1), c.rows);
    p.capacity = cap;
    std::vector<long long> out_key(cap + 1), out_cnt(cap + 1), out_isum(cap + 1);
    std::vector<int> out_key32(cap + 1);
    std::vector<double> out_sum(cap + 1), out_avg(cap + 1);
    p.out_key = c.key_kind == BQ_INT64 ? static_cast<void*>(out_key.data()) : static_cast<void*>(out_key32.data());
    // outputs: COUNT(*), then per case SUM / AVG over the arguments
    int n_out = 0;
    p.func[n_out] = BQ_AGG_COUNT; p.out[n_out] = out_cnt.data(); ++n_out;
    if (c.nv == 1) {
        p.func[n_out] = BQ_AGG_SUM; p.v[n_out] = 0; p.out[n_out] = out_sum.data(); ++n_out;
        p.func[n_out] = BQ_AGG_AVG; p.v[n_out] = 0; p.out[n_out] = out_avg.data(); ++n_out;
    } else if (c.nv == 2) {
        p.func[n_out] = BQ_AGG_SUM; p.v[n_out] = 0; p.as_int[n_out] = 1; p.out[n_out] = out_isum.data(); ++n_out;     // SUM(INT64) -> int64
        p.func[n_out] = BQ_AGG_SUM; p.v[n_out] = 1; p.out[n_out] = out_sum.data(); ++n_out;
        p.func[n_out] = BQ_AGG_AVG; p.v[n_out] = 1; p.out[n_out] = out_avg.data(); ++n_out;
    }
    p.n_out = n_out;
    unsigned long long cursor = 0;
    int err = 0;
    p.cursor = &cursor;
    p.err = &err;

    emu::launch(P * static_cast<unsigned>(c.splits), bq::kGroupThreads, bq::group_smem_bytes(c.slots, c.nv), [&] { bq::k_group_tables(p); });

    if (c.expect_overflow) {
        expect((err & 2) != 0, c.name + ": a table that cannot hold its groups must report overflow");
        expect((err & ~2) == 0, c.name + ": no other error bit");
        std::printf("%-28s overflow reported\n", c.name.c_str());
        return;
    }
    expect(err == 0, c.name + ": error word " + std::to_string(err));
    expect(cursor == want.size(), c.name + ": " + std::to_string(cursor) + " groups, expected " + std::to_string(want.size()));
    std::map<long long, size_t> seen;
    for (size_t g = 0; g < cursor && g < cap; ++g) {
        const long long k = c.key_kind == BQ_INT64 ? out_key[g] : static_cast<long long>(out_key32[g]);
        if (!seen.emplace(k, g).second) { expect(false, c.name + ": key " + std::to_string(k) + " emitted twice"); continue; }
        auto it = want.find(k);
        if (it == want.end()) { expect(false, c.name + ": unknown key " + std::to_string(k)); continue; }
        const Agg& w = it->second;
        bool ok = out_cnt[g] == w.count;
        if (c.nv == 1) ok = ok && out_sum[g] == w.s0 && out_avg[g] == w.s0 / static_cast<double>(w.count);
        if (c.nv == 2) ok = ok && out_isum[g] == static_cast<long long>(w.s0) && out_sum[g] == w.s1 && out_avg[g] == w.s1 / static_cast<double>(w.count);
        if (!ok) expect(false, c.name + ": wrong aggregates for key " + std::to_string(k));
    }
    std::printf("%-28s %zu rows -> %llu groups ok\n", c.name.c_str(), c.rows, cursor);
}

}  // namespace

int main() {
    const std::vector<Case> cases = {
        {"one sum, three splits", 60000, 5000, 2, 3, 1024, BQ_INT64, 1, 0.0, false},
        {"two sums, heavy key", 40000, 2500, 3, 1, 1024, BQ_INT64, 2, 0.30, false},
        {"count only, date keys", 30000, 3000, 1, 2, 2048, BQ_DATE32, 0, 0.05, false},
        {"one partition, ragged tail", 4099, 700, 0, 1, 1024, BQ_INT64, 1, 0.0, false},
        {"near-full tables", 50000, 3600, 2, 1, 1024, BQ_INT64, 1, 0.0, false},          // ~900 groups per 1024-slot table
        {"empty partitions", 37, 5, 4, 2, 1024, BQ_INT64, 1, 0.0, false},
        {"overflow", 30000, 9000, 1, 1, 1024, BQ_INT64, 1, 0.0, true},
    };
    for (const Case& c : cases) run(c);
    if (failures) {
        std::fprintf(stderr, "%d failure(s)\n", failures);
        return 1;
    }
    std::printf("group tables emulation ok\n");
    return 0;
}
