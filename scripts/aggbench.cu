// aggbench.cu — development microbenchmark: what does one group-table update per row cost on this GPU, by table
// layout and by where the table lives?  It decides the design of the high-cardinality GROUP BY (bq_hashagg.cu).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/aggbench.bin scripts/aggbench.cu && scripts/aggbench.bin
// Every variant streams (key, value) rows (16 B/row, as the partitioned GROUP BY does) and updates slot = key & mask:
//   soa      LD keys[s]; RED cnt[s]; RED sum[s]            three arrays (what k_scan's G_HASH does today)
//   aos32    LD slot.key; RED slot.cnt; RED slot.sum        one 32-byte sector per group
//   aos16    RED slot.cnt; RED slot.sum                     direct-address, 16-byte slot, no key check
//   red1     RED sum[s]                                     one reduction per row
//   smem     per-CTA shared table (keys masked to it): LDS key; ATOMS cnt; ATOMS sum (f64)
//   smemsort per-CTA tile: sort-free "owner" scheme: each thread owns slots, scans the tile's keys in shared memory
// Table regions from L2-resident (8..64 MB) to HBM-sized (4 GB).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ int4 ldg_na(const int4* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// keys are generated so that consecutive windows of `window` rows fall into one table region of `region` slots
// (what partitioned input looks like): slot = region_id * region + random % region
__global__ void k_gen(long long* keys, double* vals, size_t n, size_t window, size_t region, size_t total_slots) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        size_t regions = total_slots / region;
        size_t rid = (i / window) % regions;
        keys[i] = (long long)(rid * region + mix(i) % region);
        vals[i] = (double)(mix(i ^ 0x55) % 6400) / 64.0;
    }
}

struct Slot32 { long long key; unsigned long long cnt; double sum; long long pad; };
struct Slot16 { unsigned long long cnt; double sum; };

template <int MODE>
__global__ void __launch_bounds__(256) k_agg(const long long* __restrict__ keys, const double* __restrict__ vals, size_t n,
                                             long long* tkeys, unsigned long long* tcnt, double* tsum, Slot32* t32, Slot16* t16,
                                             unsigned long long* sink) {
    // a warp owns 128 rows per trip: lane t reads rows {2t, 2t+1, 64+2t, 65+2t} with two 128-bit loads per column
    const int lane = threadIdx.x & 31;
    const size_t warps = (size_t)gridDim.x * (blockDim.x / 32);
    const size_t warp = (size_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    unsigned long long bad = 0;
    for (size_t c = warp; c < n / 128; c += warps) {
        const size_t base = c * 128;
        const int4* kq = reinterpret_cast<const int4*>(keys + base + 2 * lane);
        const int4* vq = reinterpret_cast<const int4*>(vals + base + 2 * lane);
        int4 k0 = ldg_na(kq), k1 = ldg_na(kq + 32), v0 = ldg_na(vq), v1 = ldg_na(vq + 32);
        long long k[4] = {(long long)(((unsigned long long)(unsigned)k0.y << 32) | (unsigned)k0.x), (long long)(((unsigned long long)(unsigned)k0.w << 32) | (unsigned)k0.z),
                          (long long)(((unsigned long long)(unsigned)k1.y << 32) | (unsigned)k1.x), (long long)(((unsigned long long)(unsigned)k1.w << 32) | (unsigned)k1.z)};
        double v[4] = {__hiloint2double(v0.y, v0.x), __hiloint2double(v0.w, v0.z), __hiloint2double(v1.y, v1.x), __hiloint2double(v1.w, v1.z)};
        if (MODE == 0) {
            long long cur[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) cur[r] = *reinterpret_cast<volatile long long*>(tkeys + k[r]);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                bad += cur[r] != k[r];
                atomicAdd(tcnt + k[r], 1ULL);
                atomicAdd(tsum + k[r], v[r]);
            }
        } else if (MODE == 1) {
            long long cur[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) cur[r] = *reinterpret_cast<volatile long long*>(&t32[k[r]].key);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                bad += cur[r] != k[r];
                atomicAdd(&t32[k[r]].cnt, 1ULL);
                atomicAdd(&t32[k[r]].sum, v[r]);
            }
        } else if (MODE == 2) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                atomicAdd(&t16[k[r]].cnt, 1ULL);
                atomicAdd(&t16[k[r]].sum, v[r]);
            }
        } else if (MODE == 3) {
#pragma unroll
            for (int r = 0; r < 4; ++r) atomicAdd(tsum + k[r], v[r]);
        } else if (MODE == 4) {
#pragma unroll
            for (int r = 0; r < 4; ++r) atomicAdd(tcnt + k[r], (unsigned long long)k[r] | 1ULL);
        } else if (MODE == 5) {
            // plain (non-atomic) read-modify-write: what the L2 does without the atomic unit
#pragma unroll
            for (int r = 0; r < 4; ++r) tsum[k[r]] = v[r];
        } else {
            // claim path like the product kernel: load key; CAS when empty; then two REDs
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                long long cur = *reinterpret_cast<volatile long long*>(tkeys + k[r]);
                if (cur != k[r]) atomicCAS(reinterpret_cast<unsigned long long*>(tkeys + k[r]), 0ULL, (unsigned long long)k[r]);
                atomicAdd(tcnt + k[r], 1ULL);
                atomicAdd(tsum + k[r], v[r]);
            }
        }
    }
    if (bad == 0xFFFFFFFFFFFFULL) *sink = bad;
}

// shared-memory table: each CTA owns a contiguous row range whose keys (masked) fall into its own table of S slots
template <int S>
__global__ void __launch_bounds__(256) k_agg_smem(const long long* __restrict__ keys, const double* __restrict__ vals, size_t n,
                                                  size_t rows_per_cta, unsigned long long* out_cnt, double* out_sum) {
    extern __shared__ unsigned char raw[];
    long long* skey = reinterpret_cast<long long*>(raw);
    double* ssum = reinterpret_cast<double*>(skey + S);
    unsigned* scnt = reinterpret_cast<unsigned*>(ssum + S);
    for (int i = threadIdx.x; i < S; i += blockDim.x) { skey[i] = i; ssum[i] = 0.0; scnt[i] = 0; }
    __syncthreads();
    const size_t lo = blockIdx.x * rows_per_cta, hi = lo + rows_per_cta < n ? lo + rows_per_cta : n;
    unsigned bad = 0;
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const int s = (int)(keys[i] & (S - 1));
        bad += skey[s] != s;
        atomicAdd(scnt + s, 1u);
        atomicAdd(ssum + s, vals[i]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        out_cnt[(size_t)blockIdx.x * S + i] = scnt[i] + bad;
        out_sum[(size_t)blockIdx.x * S + i] = ssum[i];
    }
}

int main(int argc, char** argv) {
    size_t n = argc > 1 ? (size_t)atof(argv[1]) : (size_t)256e6;
    n = n / 128 * 128;
    long long* keys; double* vals;
    CK(cudaMalloc(&keys, n * 8)); CK(cudaMalloc(&vals, n * 8));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    unsigned long long* sink; CK(cudaMalloc(&sink, 8));
    int sm = 0; CK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0));
    printf("SMs %d rows %zu\n", sm, n);
    const size_t total_slots = 1ull << 27;             // 128 Mi slots: 4 GB at 32 B
    void* table; CK(cudaMalloc(&table, total_slots * 32));
    const char* names[7] = {"soa ", "aos32", "aos16", "red1", "red1u64", "store", "soa+claim"};
    // region = slots touched by one window of rows; rows per slot = 10 inside a window
    for (size_t region : {(size_t)1 << 19, (size_t)1 << 21, total_slots}) {
        const size_t window = region == total_slots ? n : region * 5;       // load factor 0.5 at 10 rows per key
        k_gen<<<sm * 8, 256>>>(keys, vals, n, window, region, total_slots);
        CK(cudaDeviceSynchronize());
        for (int mode = 0; mode < 7; ++mode) {
            CK(cudaMemset(table, 0, total_slots * 32));
            // keys[s] = s so the key check passes
            long long* tkeys = (long long*)table;
            unsigned long long* tcnt = (unsigned long long*)table + total_slots;
            double* tsum = (double*)table + 2 * total_slots;
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaEventRecord(e0));
                switch (mode) {
                    case 0: k_agg<0><<<sm * 8, 256>>>(keys, vals, n, tkeys, tcnt, tsum, nullptr, nullptr, sink); break;
                    case 1: k_agg<1><<<sm * 8, 256>>>(keys, vals, n, nullptr, nullptr, nullptr, (Slot32*)table, nullptr, sink); break;
                    case 2: k_agg<2><<<sm * 8, 256>>>(keys, vals, n, nullptr, nullptr, nullptr, nullptr, (Slot16*)table, sink); break;
                    case 3: k_agg<3><<<sm * 8, 256>>>(keys, vals, n, nullptr, nullptr, tsum, nullptr, nullptr, sink); break;
                    case 4: k_agg<4><<<sm * 8, 256>>>(keys, vals, n, nullptr, tcnt, nullptr, nullptr, nullptr, sink); break;
                    case 5: k_agg<5><<<sm * 8, 256>>>(keys, vals, n, nullptr, nullptr, tsum, nullptr, nullptr, sink); break;
                    default: k_agg<6><<<sm * 8, 256>>>(keys, vals, n, tkeys, tcnt, tsum, nullptr, nullptr, sink); break;
                }
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
            CK(cudaGetLastError());
            const double bytes_region = (double)region * (mode == 1 ? 32 : mode == 2 ? 16 : mode == 0 ? 24 : 8);
            printf("region %9zu slots (%7.1f MB touched at a time)  %s  %8.3f ms  %7.2f Grows/s  stream %6.0f GB/s\n", region,
                   bytes_region / 1e6, names[mode], best, n / best / 1e6, 16.0 * n / best / 1e6);
        }
    }
    // shared-memory tables
    {
        unsigned long long* oc; double* os;
        const int ctas = sm * 2;
        CK(cudaMalloc(&oc, (size_t)ctas * 8192 * 8)); CK(cudaMalloc(&os, (size_t)ctas * 8192 * 8));
        const size_t per = (n + ctas - 1) / ctas;
        auto run = [&](auto kern, int S) {
            size_t smem = (size_t)S * 20;
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaEventRecord(e0));
                kern<<<ctas, 256, smem>>>(keys, vals, n, per, oc, os);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
            CK(cudaGetLastError());
            printf("smem table %5d slots  %8.3f ms  %7.2f Grows/s\n", S, best, n / best / 1e6);
        };
        run(k_agg_smem<1024>, 1024);
        run(k_agg_smem<4096>, 4096);
        run(k_agg_smem<8192>, 8192);
    }
    return 0;
}
