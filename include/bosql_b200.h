/* bosql_b200.h — C ABI of the B200-native bo-sql operator hot path (kernel layer).
 *
 * This is the drop-in boundary: plain pointers, sizes and opaque handles, no C++ or torch types.
 * The C++ operator mirror in bo-sql_b200/host/ (same class names and constructor signatures as the
 * reference's include/exec/operator.hpp:17-218) is the only product caller; tests and bench.py bind
 * the same symbols through ctypes.  Each entry point names the reference code it replaces
 * (paths relative to the reference repository).
 *
 * Conventions
 *  - every function returns 0 on success; on failure it returns nonzero and bq_last_error() holds the
 *    message (the C++ wrapper rethrows it as std::runtime_error, the reference's only error channel,
 *    SURVEY.md section 5).
 *  - type ordinals equal the reference's TypeId (include/types.h:17): INT64, DOUBLE, STRING, DATE32.
 *    Physical widths: 8, 8, 4 (uint32 dictionary id), 4 (int32 YYYYMMDD)  (include/types.h:11-14).
 *  - all kernels run on the context's stream (bq_ctx_set_stream); nothing synchronises the device
 *    unless it returns host-visible data.
 *  - there is NO CPU fallback anywhere behind this header: without a CUDA device bq_ctx_create fails.
 */
#ifndef BOSQL_B200_H
#define BOSQL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bq_ctx bq_ctx;   /* one per process per GPU */
typedef struct bq_col bq_col;   /* device-resident column: the HBM mirror of ColumnVector<T> (include/types.h:134-145) */
typedef struct bq_rel bq_rel;   /* device-resident relation: equal-length columns (an operator's whole output) */
typedef struct bq_join bq_join; /* built join table (HashJoin::open, src/exec/operator.cpp:739-762) */

enum { BQ_INT64 = 0, BQ_DOUBLE = 1, BQ_STRING = 2, BQ_DATE32 = 3 };

/* ---- context ------------------------------------------------------------------------------- */
int bq_ctx_create(int device, bq_ctx** out);
void bq_ctx_destroy(bq_ctx* ctx);
int bq_ctx_set_stream(bq_ctx* ctx, void* cuda_stream);    /* cudaStream_t; NULL = the context's own stream */
int bq_ctx_sync(bq_ctx* ctx);
void* bq_ctx_stream(bq_ctx* ctx);                         /* the cudaStream_t every launch is ordered on */
/* bytes the stream-ordered allocator holds from the driver / has handed out (diagnostics) */
int bq_ctx_pool_stats(bq_ctx* ctx, size_t* reserved_bytes, size_t* used_bytes);
/* stream-ordered device-to-device copy / zero fill of raw bytes (packing exchange buffers for the multi-GPU path) */
int bq_copy_bytes(bq_ctx* ctx, void* dst, const void* src, size_t bytes);
int bq_zero_bytes(bq_ctx* ctx, void* dst, size_t bytes);
int bq_ctx_info(bq_ctx* ctx, int* sm_count, size_t* free_bytes, size_t* total_bytes);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
uint64_t bq_ctx_launches(bq_ctx* ctx);
/* Kernel timing for the roofline: while enabled, every launch of the fused scan kernel (bq_scan.cu k_scan) is
 * bracketed by CUDA events on the context's stream.  bq_ctx_profile_read synchronises, returns the number of
 * bracketed launches and their summed duration, and clears the record. */
int bq_ctx_profile(bq_ctx* ctx, int enable);
int bq_ctx_profile_read(bq_ctx* ctx, uint64_t* launches, double* total_ms);
const char* bq_last_error(void);

/* ---- columns: storage/ becomes device resident ------------------------------------------------ */
int bq_col_alloc(bq_ctx* ctx, int type, size_t n, bq_col** out);
/* synchronous upload of a host ColumnVector<T>::data (pageable or pinned) */
int bq_col_upload(bq_ctx* ctx, int type, const void* host, size_t n, bq_col** out);
/* asynchronous H2D into an existing column (host should be pinned); ordered on the context stream */
int bq_col_write(bq_ctx* ctx, bq_col* col, size_t offset, const void* host, size_t n);
/* synchronous D2H of rows [offset, offset+n) */
int bq_col_read(bq_ctx* ctx, const bq_col* col, size_t offset, size_t n, void* host);
/* the same copy enqueued on the context's stream without waiting (host must be pinned: bq_host_alloc); the bytes are there
 * after the next bq_ctx_sync - several result columns cost one synchronisation instead of one each */
int bq_col_read_async(bq_ctx* ctx, const bq_col* col, size_t offset, size_t n, void* host);
void bq_col_free(bq_ctx* ctx, bq_col* col);
/* non-owning column over device memory the caller manages (e.g. a buffer filled by an NCCL collective) */
int bq_col_wrap(bq_ctx* ctx, int type, void* device_ptr, size_t n, bq_col** out);
size_t bq_col_size(const bq_col* col);
int bq_col_owns(const bq_col* col);                        /* 1: the handle owns its device memory (not a wrap / view) */
int bq_col_type(const bq_col* col);
void* bq_col_ptr(const bq_col* col);                       /* raw device pointer */
/* Catalog statistics (include/catalog/catalog.h:16-21): min/max on the column's integer key
 * (for DOUBLE: the order-preserving key of the value, see bq_f64_key) and NDV; used to size tables. */
int bq_col_set_stats(bq_col* col, int64_t min_key, int64_t max_key, size_t ndv);
void bq_col_invalidate_stats(bq_col* col);
/* min/max computed on the device and cached when the catalog gave none */
int bq_col_minmax(bq_ctx* ctx, bq_col* col, int64_t* min_key, int64_t* max_key);
/* order-preserving int64 key of a double (so that every predicate is an integer range test) */
int64_t bq_f64_key(double v);
double bq_f64_from_key(int64_t k);

/* pinned host memory for paging results / staging uploads */
int bq_host_alloc(size_t bytes, void** out);
void bq_host_free(void* p);

/* ---- deterministic synthetic columns ---------------------------------------------
 * value(row) = f(seed, stream, global_row) with a counter-based hash, so any slice can be regenerated
 * on the host (oracle/datagen.py restates the same arithmetic in numpy). */
enum {
    BQ_GEN_SEQ = 0,        /* lo + global_row * stride                          (unique keys; dense when stride = 1) */
    BQ_GEN_UNIFORM = 1,    /* lo + (hash % (hi-lo+1)) * stride                  (ints, ids; stride = modulus, default 1) */
    BQ_GEN_UNIFORM_DIV = 2,/* (double)(lo + hash % (hi-lo+1)) / div             (DOUBLE, k/div)          */
    BQ_GEN_DATE = 3,       /* base_year*10000 + 100*m + d, m in 1..12, d in 1..28, uniform over years    */
    BQ_GEN_TABLE = 4,      /* inverse-CDF lookup: smallest i with cdf[i] > hash>>11 (53-bit), value lo+i */
    BQ_GEN_HASHED = 5,     /* lo + mix(hash % (hi-lo+1)) % modulus: sparse keys drawn from hi-lo+1 ids   */
    BQ_GEN_BUCKETS = 6     /* inverse-CDF over n_cdf buckets, then uniform inside the bucket:
                              i as BQ_GEN_TABLE, value lo + starts[i] + mix(hash) % (starts[i+1] - starts[i]).
                              Heavy-tailed keys over a domain too large for one threshold per key (Zipf over 5e8 ids) */
};
typedef struct bq_gen_spec {
    int dist;
    uint64_t seed;
    uint64_t stream;       /* column id */
    int64_t lo, hi;
    double div;            /* BQ_GEN_UNIFORM_DIV */
    int32_t base_year;     /* BQ_GEN_DATE */
    int32_t n_years;       /* BQ_GEN_DATE */
    const uint64_t* cdf;   /* BQ_GEN_TABLE: HOST pointer to n_cdf ascending 53-bit thresholds */
    size_t n_cdf;
    uint64_t modulus;      /* BQ_GEN_HASHED; BQ_GEN_SEQ / BQ_GEN_UNIFORM: stride between values (0 = 1) */
    const uint64_t* starts; /* BQ_GEN_BUCKETS: HOST pointer to n_cdf + 1 ascending offsets */
} bq_gen_spec;
int bq_col_generate(bq_ctx* ctx, bq_col* col, const bq_gen_spec* spec, uint64_t global_row0);

/* ---- typed expression programs: exec/expression.cpp compiled once per plan ----------------------
 * A postfix program over 8-byte slots; every operand type is resolved on the host (so the device never
 * dispatches on a Datum tag) while keeping evaluate_internal's semantics (src/exec/expression.cpp:153-206):
 * INT64-left comparisons truncate a DOUBLE right operand (:64), DOUBLE / 0 = +inf (:41), AND/OR evaluate
 * both sides (:176-198).  INT64 / 0 raises the reference's "Division by zero" (:52) as an error. */
enum {
    BQ_OP_COL = 0,     /* push column `arg` widened to 8 bytes (u32 zero-extended, i32 sign-extended) */
    BQ_OP_IMM_I = 1, BQ_OP_IMM_F = 2,
    BQ_OP_I2F = 3,     /* top: int64 -> double                      */
    BQ_OP_I2F_2 = 4,   /* second from top: int64 -> double          */
    BQ_OP_F2I = 5,     /* top: static_cast<int64_t>(double)         */
    BQ_OP_SX32 = 6,    /* top: keep the low 32 bits, sign-extended (DATE32-left comparison, :93-94) */
    BQ_OP_ZX32 = 7,    /* top: keep the low 32 bits, zero-extended (STRING-left comparison, :107-108) */
    BQ_OP_ADD_I = 8, BQ_OP_SUB_I = 9, BQ_OP_MUL_I = 10, BQ_OP_DIV_I = 11,
    BQ_OP_ADD_F = 12, BQ_OP_SUB_F = 13, BQ_OP_MUL_F = 14, BQ_OP_DIV_F = 15,
    BQ_OP_EQ_I = 16, BQ_OP_NE_I = 17, BQ_OP_LT_I = 18, BQ_OP_LE_I = 19, BQ_OP_GT_I = 20, BQ_OP_GE_I = 21,
    BQ_OP_EQ_F = 22, BQ_OP_NE_F = 23, BQ_OP_LT_F = 24, BQ_OP_LE_F = 25, BQ_OP_GT_F = 26, BQ_OP_GE_F = 27,
    BQ_OP_TRUTHY_I = 28, /* top: int -> 0/1   (is_truthy, :10-22)  */
    BQ_OP_TRUTHY_F = 29, /* top: double != 0.0 -> 0/1               */
    BQ_OP_TRUTHY_I_2 = 30, BQ_OP_TRUTHY_F_2 = 31, /* same on the second from top */
    BQ_OP_AND = 32, BQ_OP_OR = 33                  /* on 0/1 ints */
};
typedef struct bq_insn {
    int32_t op;
    int32_t arg;
    union { int64_t i; double f; } imm;
} bq_insn;
#define BQ_MAX_PROGRAM 64
#define BQ_MAX_PROGRAM_COLS 8
/* Evaluate a program over rows [row_begin,row_end) of `cols`; the result column has `out_type`
 * (INT64 takes the slot as int64, DOUBLE as double, STRING/DATE32 its low 32 bits).
 * Replaces Project::next's per-row evaluate_expr (src/exec/operator.cpp:498-555). */
int bq_eval(bq_ctx* ctx, const bq_insn* prog, int n_insn, const bq_col* const* cols, int n_cols,
            size_t row_begin, size_t row_end, int out_type, bq_col** out);

/* ---- fused scan -> selection -> [join probe] -> aggregate ------------------------------------------
 * One kernel replaces ColumnarScan::next + Selection::next + HashJoin::next + HashAggregate::next's
 * accumulate phase (src/exec/operator.cpp:345, 403, 764, 984-1014) for the pipeline shapes of the
 * configurations.  Columns are bound to fixed slots so every value lives in a register:
 *   key   group key column                       a, b   aggregate argument columns
 *   pred  up to 3 predicate-only columns         jkey   probe-side join key
 * Each slot carries up to two integer ranges on the column's order-preserving key: a row passes a range
 * iff (lo <= key && key <= hi) != neg.  Every `col OP literal` conjunct of the reference's predicate
 * language reduces to such a range with compare_values' semantics (src/exec/expression.cpp:60-120);
 * the host compiler (bo-sql_b200/host/expr_compile.cpp) does the reduction, anything else arrives as `mask`. */
typedef struct bq_range { int64_t lo, hi; int32_t neg; int32_t pad; } bq_range;
typedef struct bq_slot {
    const bq_col* col;     /* NULL = slot unused */
    int32_t n_ranges;
    int32_t from_build;    /* 1: read the value from the join's build side at the matched row */
    bq_range r[2];
} bq_slot;
/* aggregate argument: datum_as_double(A) / datum_as_double(B) (src/exec/operator.cpp:280-292) or
 * numeric_binary(A, B|imm) (src/exec/expression.cpp:31-58) followed by datum_as_double */
enum { BQ_V_NONE = 0, BQ_V_A = 1, BQ_V_B = 2, BQ_V_MUL = 3, BQ_V_ADD = 4, BQ_V_SUB = 5, BQ_V_DIV = 6 };
/* operand sources of a binary argument; the zero value of both fields means `A op B` */
enum { BQ_L_A = 0, BQ_L_B = 1, BQ_L_IMM = 2 };
enum { BQ_R_B = 0, BQ_R_A = 1, BQ_R_IMM = 2 };
typedef struct bq_vexpr {
    int32_t op;            /* BQ_V_A / BQ_V_B: the slot's value; BQ_V_MUL..DIV: left OP right */
    int32_t l_src;         /* BQ_L_* */
    int32_t r_src;         /* BQ_R_* */
    int32_t imm_is_f;      /* type of the immediate operand: 0 INT64 (imm_i), 1 DOUBLE (imm_f) */
    int64_t imm_i;
    double imm_f;
} bq_vexpr;
enum { BQ_GROUP_NONE = 0, BQ_GROUP_DENSE = 1, BQ_GROUP_HASH = 2 };
enum { BQ_AGG_COUNT = 0, BQ_AGG_SUM = 1, BQ_AGG_AVG = 2 };
typedef struct bq_agg_out {
    int32_t func;          /* BQ_AGG_*                                                     */
    int32_t v;             /* which value expression (0/1); ignored for COUNT               */
    int32_t as_int;        /* SUM of a non-DOUBLE argument: static_cast<int64_t>(sum) (src/exec/operator.cpp:1044) */
    int32_t pad;
} bq_agg_out;
#define BQ_MAX_AGG_OUT 8
typedef struct bq_scan_spec {
    bq_slot key, a, b, pred[3], jkey;
    size_t row_begin, row_end;
    const bq_col* mask;      /* optional INT64 0/1 column from bq_eval: predicate parts no range expresses */
    int32_t n_v;
    bq_vexpr v[2];
    int32_t group_mode;      /* BQ_GROUP_* ; DENSE needs key_min/key_max (catalog stats)    */
    int64_t key_min, key_max;
    size_t ndv_hint;         /* HASH: table capacity = next pow2 >= 2*ndv_hint               */
    const bq_join* join;     /* optional: inner-join probe on jkey                          */
    const bq_col* row_bits;  /* optional, instead of join: per-row match bits from bq_join_probe_bits (a semi-join already
                              * probed in key-range passes); needs row_begin % 128 == 0.  jkey may stay bound for its ranges */
    int32_t n_out;
    bq_agg_out out[BQ_MAX_AGG_OUT];
    /* HASH grouping over rows that bq_partition has ordered by partition = (hash(key) >> hash_part_shift) & (2^log2 - 1):
     * the table is laid out partition-major, so the slots a partition touches stay L2-resident while its rows stream by.
     * 0 = rows are in no particular order (one table-wide probe sequence). */
    int32_t hash_part_log2;
    int32_t hash_part_shift;
} bq_scan_spec;
/* Result relation: [key] then one column per out[] — HashAggregate's emit layout (src/exec/operator.cpp:1016-1062).
 * Rows = groups that received at least one row (none at all for a global aggregate over zero rows, :990-993).
 * Emit order: ascending key (DENSE) / table order (HASH); the reference's order is unordered_map iteration order
 * and carries no meaning. */
int bq_scan_aggregate(bq_ctx* ctx, const bq_scan_spec* spec, bq_rel** out);

/* Partial aggregate state for multi-GPU merges: same pipeline, but instead of finished outputs the relation
 * holds [key] count sum0 sum1 (INT64, DOUBLE, DOUBLE), to be combined across ranks and finished by bq_agg_finish. */
int bq_scan_partial(bq_ctx* ctx, const bq_scan_spec* spec, bq_rel** out);
/* Merge `n_parts` partial relations (equal keys combined, parts added in index order) and emit final outputs. */
int bq_agg_finish(bq_ctx* ctx, const bq_rel* const* parts, int n_parts, int has_key, int key_type,
                  const bq_agg_out* outs, int n_out, bq_rel** out);

/* The same exchange for DENSE / global states without compaction: the state stays on the device as it is (a 16-byte
 * header with the error word, then the sum0 | count | sum1 arrays over the key domain, of which only the prefix in use is
 * exchanged: 8 bytes per slot for a SUM-only state, whose untouched slots are marked by -0.0), ranks all-gather the raw
 * blocks - same size on every rank, because the domain comes from catalog statistics and every rank runs the same plan -
 * and ONE launch folds them in rank order.  bq_agg_state_dense returns a NULL
 * pointer for a hash-table state (use bq_scan_partial / bq_agg_finish for those). */
typedef struct bq_agg_state bq_agg_state;
int bq_scan_state(bq_ctx* ctx, const bq_scan_spec* spec, bq_agg_state** out);
int bq_agg_state_dense(const bq_agg_state* s, void** device_ptr, size_t* bytes);
int bq_agg_state_fold(bq_ctx* ctx, bq_agg_state* s, const void* gathered_blocks, int world);
int bq_agg_state_emit(bq_ctx* ctx, bq_agg_state* s, const bq_agg_out* outs, int n_out, bq_rel** out);
void bq_agg_state_free(bq_agg_state* s);

/* ---- selection vectors: Selection::next + copy_selected (src/exec/operator.cpp:403-429, 11-49) ------
 * Rows of [row_begin,row_end) passing all slot ranges (and mask), in scan order, as a STRING-typed (uint32)
 * column of row ids — warp ballot/popc compaction, stable. */
typedef struct bq_select_spec {
    bq_slot pred[4];
    const bq_col* mask;
    size_t row_begin, row_end;
} bq_select_spec;
int bq_select(bq_ctx* ctx, const bq_select_spec* spec, bq_col** out_rowids);
/* out[i] = col[rowids[i]] */
int bq_gather(bq_ctx* ctx, const bq_col* col, const bq_col* rowids, bq_col** out);
/* out = col[begin:end) (copy_range, src/exec/operator.cpp:51-82) */
int bq_slice(bq_ctx* ctx, const bq_col* col, size_t begin, size_t end, bq_col** out);

/* ---- hash join ---------------------------------------------------------------------------------------
 * Build on the right child (HashJoin::open, src/exec/operator.cpp:739-762).  The table kind comes from catalog
 * statistics: BITMAP when the key is unique and dense and no build column is read downstream (semi-join),
 * DIRECT (row id array indexed by key-min) when unique and dense, HASH (open addressing, linear probing,
 * duplicate keys kept) otherwise.  Rows failing the build-side ranges / mask are not inserted — the
 * reference evaluates such predicates above the join (src/logical/planner.cpp:110-117); for an inner join the
 * row set is the same. */
enum { BQ_JOIN_AUTO = 0, BQ_JOIN_BITMAP = 1, BQ_JOIN_DIRECT = 2, BQ_JOIN_HASH = 3, BQ_JOIN_ROWBITS = 4 /* scan-side only */ };
typedef struct bq_join_spec {
    const bq_col* key;
    bq_slot pred[3];
    const bq_col* mask;
    size_t row_begin, row_end;
    int32_t kind;            /* BQ_JOIN_* */
    int32_t need_rows;       /* build columns are read downstream (rules out BITMAP) */
    int32_t unique;          /* catalog: ndv == row_count */
    int32_t pad;
    int64_t key_min, key_max;/* catalog min/max of the key */
} bq_join_spec;
int bq_join_build(bq_ctx* ctx, const bq_join_spec* spec, bq_join** out);
void bq_join_free(bq_ctx* ctx, bq_join* j);
int bq_join_kind(const bq_join* j);
size_t bq_join_bytes(const bq_join* j);
size_t bq_join_build_rows(const bq_join* j);               /* rows inserted (after the build-side predicates) */
/* BITMAP joins: raw words, so ranks can exchange/OR partial bitmaps (multi-GPU broadcast join) */
void* bq_join_bitmap_ptr(const bq_join* j, size_t* n_words);
/* The same build for a bitmap that ranks are about to merge, WITHOUT a host round trip: the kernel is enqueued and its
 * counters (rows inserted, out-of-domain flag) are packed into BQ_JOIN_TRAILER_WORDS uint32 words right behind the bitmap's
 * words, in limbs that a word-wise SUM over up to 2048 ranks cannot overflow.  Sum n_words + BQ_JOIN_TRAILER_WORDS words
 * across the ranks, then ask for the verdict: bits set in the merged bitmap, rows all ranks inserted, flags (2 = some rank
 * saw a key outside [key_min, key_max]).  set_bits != inserted means a key was inserted twice (by one rank or by two). */
#define BQ_JOIN_TRAILER_WORDS 8
int bq_join_build_bitmap_nosync(bq_ctx* ctx, const bq_join_spec* spec, bq_join** out);
int bq_join_bitmap_verdict(bq_ctx* ctx, bq_join* j, uint64_t* set_bits, uint64_t* inserted, int* flags);
/* number of set bits (after ranks merged their bitmaps by summing words: must equal the rows they inserted, or two ranks
 * held the same key and the sum was not an OR) */
int bq_join_bitmap_popcount(bq_ctx* ctx, const bq_join* j, uint64_t* out);
/* Semi-join probe in KEY-RANGE PASSES for bitmaps larger than L2 (a 2-billion-key domain is 250 MB): the bitmap is cut into
 * slices of about slice_bytes; pass j streams the probe key once and tests only the rows whose key falls into slice j, so the
 * slice stays L2-resident while it is hot (a probe into an HBM-resident bitmap costs a 32-byte DRAM sector per row instead).
 * Result: a uint32 column of ceil((row_end-row_begin)/32) words (+ padding), bit i = row row_begin+i has a match.  The fused
 * scan then takes it as bq_scan_spec.row_bits and no longer reads the probe key. */
int bq_join_probe_bits(bq_ctx* ctx, const bq_join* j, const bq_col* probe_key, size_t row_begin, size_t row_end,
                       size_t slice_bytes, bq_col** out_bits);
/* Materialising probe (HashJoin::next, src/exec/operator.cpp:764-837): all (probe row, build row) pairs in probe
 * order, matches of one probe row in build insertion order. `probe_rowids` (optional) restricts/ordering the probe rows. */
int bq_join_probe(bq_ctx* ctx, const bq_join* j, const bq_col* probe_key, const bq_col* probe_rowids,
                  size_t row_begin, size_t row_end, bq_col** out_probe_rows, bq_col** out_build_rows);

/* ---- hash partitioning: the exchange step (no counterpart in the single-process reference) -----------------------
 * Reorders rows [row_begin,row_end) of key and up to two payload columns so that rows of one partition
 * ((hash(key) >> hash_shift) & (2^log2_parts - 1)) are contiguous; out_offsets is an INT64 column of 2^log2_parts + 1 row
 * offsets.  Used (a) before a high-cardinality GROUP BY / join so each partition's table fits in L2, and (b) to fill
 * the per-peer send buffers of the multi-GPU all-to-all (2^log2_parts = number of ranks). */
int bq_partition(bq_ctx* ctx, const bq_col* key, const bq_col* const* payload, int n_payload, size_t row_begin, size_t row_end,
                 int log2_parts, int hash_shift, bq_col** out_key, bq_col** out_payload, bq_col** out_offsets);
/* High-cardinality GROUP BY over rows that bq_partition has ordered (HashAggregate's accumulate and emit phases,
 * src/exec/operator.cpp:984-1062, when tens of millions of groups receive a handful of rows each): one launch, `splits`
 * CTAs per partition, each keeping the groups of its share of the partition's keys in a SHARED-MEMORY table that is updated
 * without atomics (slot ownership by tag, see csrc/bq_groupby.cuh) and writing the finished output columns itself - no
 * table in HBM, no presence / compaction / emit passes.  `args` are the aggregate arguments as plain numeric columns
 * (outs[i].v indexes them), `offsets` is bq_partition's out_offsets for the same log2_parts with hash_shift = 64 - log2_parts.
 * Result: [key] then one column per outs[], groups in no particular order.  Fails with "group table
 * overflow" when a table cannot hold its keys (statistics too low, or a skewed partition): the caller then aggregates the
 * same partitioned rows with bq_scan_aggregate (hash_part_log2 = log2_parts).
 * bq_group_tables_plan: the sizing rule - how many partitions and splits make every table's expected load <= 0.55;
 * returns 0 when more than four splits would be needed (the caller keeps the L2-resident table of bq_scan_aggregate). */
int bq_group_tables_plan(size_t ndv_hint, int n_args, int* log2_parts, int* splits);
int bq_partition_aggregate(bq_ctx* ctx, const bq_col* key, const bq_col* const* args, int n_args, const bq_col* offsets,
                           int log2_parts, int splits, const bq_agg_out* outs, int n_out, bq_rel** out);
/* The same pass in two halves, for the multi-GPU shuffle: count first (host_counts[2^log2_parts] = rows per partition, one
 * host round trip), then scatter every partition to an address of the caller's choice - dest_key[q] / dest_pay*[q] is
 * where THIS launch's first row of partition q goes.  The addresses may lie in a peer GPU's memory (bq_ipc_open): the
 * kernel then writes the exchange straight over NVLink, fused with the partitioning - no send buffer, no collective. */
typedef struct bq_part_plan bq_part_plan;
int bq_partition_count(bq_ctx* ctx, const bq_col* key, size_t row_begin, size_t row_end, int log2_parts, int hash_shift,
                       int64_t* host_counts, bq_part_plan** out);
/* Skew handling: rows whose key is one of hot_keys[0..n_hot) (<= 16) are counted in, and later scattered to, partition
 * 2^log2_parts instead of their hash partition; host_counts and the destination tables then have 2^(log2_parts+1) entries
 * (the entries above the hot one stay empty).  A shuffle keeps that partition local. */
int bq_partition_count_hot(bq_ctx* ctx, const bq_col* key, size_t row_begin, size_t row_end, int log2_parts, int hash_shift,
                           const int64_t* hot_keys, int n_hot, int64_t* host_counts, bq_part_plan** out);
int bq_partition_scatter(bq_ctx* ctx, bq_part_plan* plan, const bq_col* const* payload, int n_payload,
                         void* const* dest_key, void* const* dest_pay0, void* const* dest_pay1);
void bq_part_plan_free(bq_part_plan* plan);
/* ---- peer memory (one process per GPU on one NVSwitch domain) ----------------------------------------------------
 * bq_col_alloc_shared: a column in a block of its own that can be exported; bq_col_ipc_export: its 64-byte CUDA IPC handle
 * (send it to the peers by any host channel); bq_ipc_open: the peer's view of that block (mapped once per handle, cached
 * for the life of the context).  Ordering between processes is the caller's job (a host barrier before and after). */
#define BQ_IPC_HANDLE_BYTES 64
int bq_col_alloc_shared(bq_ctx* ctx, int type, size_t n, bq_col** out);
int bq_col_ipc_export(bq_ctx* ctx, const bq_col* col, void* handle64);
int bq_ipc_open(bq_ctx* ctx, const void* handle64, void** device_ptr);
size_t bq_ctx_ipc_mappings(bq_ctx* ctx);                 /* peer blocks mapped so far (diagnostics) */
/* ---- collectives (one process per GPU; NCCL over NVLink / NVSwitch) -----------------------------------------------
 * The multi-GPU exchange points of a plan (bo-sql_b200/host/exchange.cpp) run on these.  Rank 0 makes a
 * unique id (bq_comm_unique_id) and hands its 128 bytes to every rank over any host channel; every rank then calls
 * bq_comm_init on its own context.  All collectives are enqueued on the context's stream and return at once; only
 * bq_comm_host_all_gather_i64 synchronises, because it returns values to the CPU (row counts, statistics, outcome words:
 * all[r*n + i] = rank r's mine[i], n <= 512).  Byte counts are exact; *_v forms take one count per rank. */
#define BQ_COMM_ID_BYTES 128
int bq_comm_unique_id(void* id128);
int bq_comm_init(bq_ctx* ctx, int world, int rank, const void* id128);
void bq_comm_destroy(bq_ctx* ctx);
int bq_comm_world(bq_ctx* ctx);
int bq_comm_rank(bq_ctx* ctx);
int bq_comm_all_gather(bq_ctx* ctx, const void* send, void* recv, size_t bytes);
int bq_comm_all_gather_v(bq_ctx* ctx, const void* send, void* recv, const int64_t* bytes_by_rank);
int bq_comm_all_to_all_v(bq_ctx* ctx, const void* send, const int64_t* send_bytes, void* recv, const int64_t* recv_bytes);
int bq_comm_all_reduce_sum_u32(bq_ctx* ctx, void* buf, size_t words);
int bq_comm_host_all_gather_i64(bq_ctx* ctx, const int64_t* mine, int32_t n, int64_t* all);
/* calls5: all_gather, all_gather_v, all_to_all_v, all_reduce_sum_u32, host_all_gather_i64 issued so far; payload bytes sent */
int bq_comm_stats(bq_ctx* ctx, uint64_t* calls5, uint64_t* bytes_sent);
/* the hash all tables and partitions use (so a caller can predict a key's partition) */
uint64_t bq_key_hash(int64_t key);

/* ---- OrderBy / Limit (src/exec/operator.cpp:1097-1151, 561-620) ---------------------------------------
 * Stable sort of the relation by its key columns (asc/desc each; rank sort / tile top-k for up to 4 keys, LSD radix passes
 * for any number); limit >= 0 keeps the first
 * `limit` rows (top-k when the relation is large).  Ties keep input order (the reference's std::sort leaves
 * them unspecified). */
int bq_rel_sort(bq_ctx* ctx, const bq_rel* rel, int n_keys, const int* key_cols, const int* asc,
                int64_t limit, bq_rel** out);

/* ---- relations -------------------------------------------------------------------------------------- */
int bq_rel_create(bq_ctx* ctx, bq_col* const* cols, int n_cols, bq_rel** out); /* takes ownership of cols */
size_t bq_rel_rows(const bq_rel* rel);
int bq_rel_cols(const bq_rel* rel);
bq_col* bq_rel_col(const bq_rel* rel, int i);     /* borrowed */
void bq_rel_free(bq_ctx* ctx, bq_rel* rel);
/* hand the columns (bq_rel_cols of them) back to the caller and destroy the shell without freeing them */
void bq_rel_release(bq_rel* rel, bq_col** out_cols);

#ifdef __cplusplus
}
#endif
#endif /* BOSQL_B200_H */
