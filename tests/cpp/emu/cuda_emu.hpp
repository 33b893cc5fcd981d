// cuda_emu.hpp — a small CUDA execution model for the host: TEST INFRASTRUCTURE, never linked into the product.
//
// Lets g++ compile a kernel written in plain CUDA C++ (no inline PTX) and run it on the CPU, so its logic - indexing,
// barrier structure, warp collectives, the order-independence of its result - is checked where there is no GPU.
// Every CUDA thread of a block is a fibre (ucontext) with its own stack; a fibre runs until it reaches a block barrier or a
// warp collective, parks there, and the scheduler resumes the next one.  Blocks run one after another, so `__shared__`
// variables are plain statics and dynamic shared memory is one buffer.  Atomics are plain operations (one OS thread).
// What this does NOT model: memory races between two barriers (the fibres of a block run in a fixed order between
// synchronisation points; EMU_SHUFFLE reverses that order every sweep to shake out order dependence), timing, banks.
#pragma once

#include <ucontext.h>

#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define BQ_CUDA_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __grid_constant__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

namespace emu {

struct Dim3 {
    unsigned x = 1, y = 1, z = 1;
};

struct WarpState {
    unsigned arrived = 0;          // lanes parked in the current collective
    unsigned gen = 0;
    unsigned ballot = 0;
    unsigned result[2] = {0, 0};
    unsigned alive = 0;            // lanes that have not returned
};

struct Fibre {
    ucontext_t ctx;
    Dim3 tid;
    bool done = false;
    unsigned warp = 0, lane = 0;
};

struct Block {
    std::vector<Fibre> fibres;
    std::vector<WarpState> warps;
    unsigned alive = 0;
    unsigned arrived = 0, gen = 0;
    int or_acc = 0, or_result[2] = {0, 0};
    Dim3 bid, bdim, gdim;
    ucontext_t sched;
    unsigned long progress = 0;    // barriers / collectives completed + fibres finished (deadlock detection)
    Fibre* cur = nullptr;
    std::function<void()> body;
    std::vector<unsigned char> dyn_smem;
};

inline Block*& current_block() {
    static Block* b = nullptr;
    return b;
}
inline Fibre& self() { return *current_block()->cur; }
inline void yield() {
    Block* b = current_block();
    swapcontext(&b->cur->ctx, &b->sched);
}

inline void block_barrier_complete(Block* b) {
    b->or_result[b->gen & 1u] = b->or_acc;
    b->or_acc = 0;
    b->arrived = 0;
    b->gen++;
    b->progress++;
}
inline void warp_collective_complete(WarpState& w) {
    w.result[w.gen & 1u] = w.ballot;
    w.ballot = 0;
    w.arrived = 0;
    w.gen++;
    current_block()->progress++;
}

inline int syncthreads_or(int pred) {
    Block* b = current_block();
    const unsigned my = b->gen;
    b->or_acc |= pred ? 1 : 0;
    b->arrived++;
    if (b->arrived == b->alive) block_barrier_complete(b);
    else while (b->gen == my) yield();
    return b->or_result[my & 1u];
}

inline unsigned warp_vote(unsigned mask, bool pred) {
    Block* b = current_block();
    Fibre& f = self();
    WarpState& w = b->warps[f.warp];
    assert(mask == 0xffffffffu && "the emulation supports full-warp collectives only");
    (void)mask;
    const unsigned my = w.gen;
    if (pred) w.ballot |= 1u << f.lane;
    w.arrived |= 1u << f.lane;
    if (w.arrived == w.alive) warp_collective_complete(w);
    else while (w.gen == my) yield();
    return w.result[my & 1u];
}

inline void fibre_entry() {
    Block* b = current_block();
    b->body();
    Fibre& f = self();
    f.done = true;
    b->progress++;
    // a thread that returns no longer takes part in barriers or collectives (Volta+ semantics)
    b->alive--;
    WarpState& w = b->warps[f.warp];
    w.alive &= ~(1u << f.lane);
    if (b->alive && b->arrived == b->alive) block_barrier_complete(b);
    if (w.alive && (w.arrived & w.alive) == w.alive && w.arrived) warp_collective_complete(w);
    swapcontext(&f.ctx, &b->sched);
}

// Runs `body` (a call of the kernel function) for every thread of every block of a 1-D launch.
inline void launch(unsigned grid, unsigned block, size_t dyn_smem_bytes, const std::function<void()>& body) {
    constexpr size_t kStack = 64 * 1024;
    std::vector<unsigned char> stacks(static_cast<size_t>(block) * kStack);
    const bool shuffle = std::getenv("EMU_SHUFFLE") != nullptr;
    Block b;
    b.body = body;
    b.dyn_smem.assign(dyn_smem_bytes + 16, 0xCD);          // poisoned: shared memory starts undefined
    current_block() = &b;
    for (unsigned g = 0; g < grid; ++g) {
        b.bid.x = g;
        b.bdim.x = block;
        b.gdim.x = grid;
        b.fibres.assign(block, Fibre{});
        b.warps.assign((block + 31) / 32, WarpState{});
        b.alive = block;
        b.arrived = 0;
        b.or_acc = 0;
        for (unsigned t = 0; t < block; ++t) {
            Fibre& f = b.fibres[t];
            f.tid.x = t;
            f.warp = t / 32;
            f.lane = t % 32;
            b.warps[f.warp].alive |= 1u << f.lane;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = stacks.data() + static_cast<size_t>(t) * kStack;
            f.ctx.uc_stack.ss_size = kStack;
            f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, fibre_entry, 0);
        }
        unsigned sweep = 0;
        while (b.alive) {
            const unsigned long before = b.progress;
            const bool reverse = shuffle && (sweep++ & 1u);
            for (unsigned i = 0; i < block; ++i) {
                Fibre& f = b.fibres[reverse ? block - 1 - i : i];
                if (f.done) continue;
                b.cur = &f;
                swapcontext(&b.sched, &f.ctx);
            }
            if (b.progress == before) {
                std::fprintf(stderr, "cuda_emu: block %u is deadlocked (a barrier or warp collective not reached by every thread)\n", g);
                std::abort();
            }
        }
    }
    current_block() = nullptr;
}

inline unsigned char* dyn_smem() {
    auto* p = current_block()->dyn_smem.data();
    return reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(p) + 15) & ~static_cast<uintptr_t>(15));
}

}  // namespace emu

#define threadIdx (emu::self().tid)
#define blockIdx (emu::current_block()->bid)
#define blockDim (emu::current_block()->bdim)
#define gridDim (emu::current_block()->gdim)

// ---- the intrinsics the emulated kernels use --------------------------------------------------------------------
template <typename T>
inline T __ldg(const T* p) { return *p; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned __umulhi(unsigned a, unsigned b) { return static_cast<unsigned>((static_cast<uint64_t>(a) * b) >> 32); }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline void __syncthreads() { emu::syncthreads_or(0); }
inline int __syncthreads_or(int pred) { return emu::syncthreads_or(pred); }
inline unsigned __ballot_sync(unsigned mask, bool pred) { return emu::warp_vote(mask, pred); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_vote(mask, false); }

inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
inline double atomicAdd(double* p, double v) { double o = *p; *p = o + v; return o; }
inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long val) {
    unsigned long long o = *p;
    if (o == cmp) *p = val;
    return o;
}
inline int atomicOr(int* p, int v) { int o = *p; *p = o | v; return o; }
