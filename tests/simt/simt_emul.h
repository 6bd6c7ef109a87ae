// simt_emul.h -- a minimal SIMT interpreter for running the repo's CUDA kernel SOURCES on the CPU.  TESTS ONLY.
//
// Why: the build container has no GPU.  Compiling csrc/*.cu with g++ against this shim executes the very same
// kernel code (templates, dispatch, warp collectives, shared-memory staging) lane by lane, so that the `-m "not gpu"`
// suite can check the kernels' logic bit for bit against the oracle before a GPU is spent on them.
// It is NOT a product path and nothing under online-3d-bpp-drl_b200/ links it: libbpp_b200.so has no CPU fallback.
// It says nothing about performance, memory-model races or PTX; `compute-sanitizer` on the B200 covers those.
//
// Model: one OS thread.  A block's threads are ucontext fibers.  A lane runs until it reaches a warp collective
// (`__syncwarp`, `__ballot_sync`, `__shfl_sync`, `__reduce_*_sync`, `__all_sync`), a block barrier (`__syncthreads`)
// or returns; the warp scheduler resumes the 32 lanes round-robin, so a collective completes once every lane of the
// warp has arrived (full-mask, converged collectives only -- which is all the kernels use; anything else aborts).
// Warps of a block run one after another up to the next block barrier; blocks run one after another.
#pragma once

#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <vector>

// ---- CUDA language keywords ----------------------------------------------------------------------------------------
#define __device__
#define __global__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static              // blocks run one at a time, so a function-local static is per block
#define __constant__ static const
#ifndef __clang__
#define __builtin_assume(x) do { if (!(x)) __builtin_unreachable(); } while (0)
#endif

// ---- vector types ----------------------------------------------------------------------------------------------------
struct uint2 { uint32_t x, y; };
struct __attribute__((aligned(16))) uint4 { uint32_t x, y, z, w; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct dim3 { unsigned x, y, z; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

// ---- the few runtime names the launch code touches -------------------------------------------------------------------
typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
namespace simt { extern int g_blocks_per_sm; }
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *b, F, int, size_t)
{
    *b = simt::g_blocks_per_sm;
    return cudaSuccess;
}

namespace simt {

enum State { RUNNABLE, WAIT_WARP, WAIT_BLOCK, DONE };

struct Fiber {
    ucontext_t ctx;
    std::vector<char> stack;
    dim3 tid;
    int lane, warp;
    unsigned ncoll;
    State state;
};

struct Machine {
    ucontext_t sched;
    std::vector<Fiber> fibers;
    Fiber *cur = nullptr;
    dim3 block_idx{0, 0, 0}, grid_dim{1, 1, 1}, block_dim{1, 1, 1};
    std::vector<uint8_t> dyn_smem;
    uint64_t arg[64][2][32];                // [warp][collective parity][lane]
    std::function<void()> body;
    unsigned long long collectives = 0;
};

extern Machine M;
extern int g_blocks_per_sm;
extern int g_lane_order;        // 0: lanes resume 0..31; 1: 31..0; 2: odd lanes first -- a missing __syncwarp() between a
                                // shared-memory write and another lane's read survives at most one of these schedules

[[noreturn]] static inline void die(const char *msg)
{
    std::fprintf(stderr, "simt_emul: %s\n", msg);
    std::abort();
}

static inline void yield(State s)
{
    Fiber *f = M.cur;
    f->state = s;
    swapcontext(&f->ctx, &M.sched);
}

// Every lane deposits `v`, waits for the other 31, and gets the 32 deposited values back.
static inline const uint64_t *warp_exchange(uint64_t v)
{
    Fiber *f = M.cur;
    const unsigned par = f->ncoll++ & 1u;
    M.arg[f->warp][par][f->lane] = v;
    yield(WAIT_WARP);
    return M.arg[f->warp][par];
}

static inline void fiber_entry()
{
    M.body();
    M.cur->state = DONE;
    swapcontext(&M.cur->ctx, &M.sched);
    die("resumed a finished fiber");
}

static inline void resume(Fiber &f)
{
    M.cur = &f;
    f.state = RUNNABLE;
    swapcontext(&M.sched, &f.ctx);
    M.cur = nullptr;
}

// Runs one block of `threads` threads to completion.
static inline void run_block(int threads)
{
    if (threads % 32) die("block size must be a multiple of 32");
    const int warps = threads / 32;
    if (warps > 64) die("at most 64 warps per block");
    if ((int)M.fibers.size() < threads) M.fibers.resize(threads);
    for (int t = 0; t < threads; ++t) {
        Fiber &f = M.fibers[t];
        if (f.stack.empty()) f.stack.resize(256 * 1024);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack.data();
        f.ctx.uc_stack.ss_size = f.stack.size();
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        f.tid = dim3{(unsigned)t, 0, 0};
        f.lane = t & 31;
        f.warp = t >> 5;
        f.ncoll = 0;
        f.state = RUNNABLE;
    }
    std::vector<State> wstate(warps, RUNNABLE);
    for (;;) {
        bool all_done = true;
        for (int w = 0; w < warps; ++w) {
            if (wstate[w] == DONE) continue;
            all_done = false;
            if (wstate[w] == WAIT_BLOCK) continue;
            for (;;) {                                   // passes over the warp until it blocks or ends
                State s = RUNNABLE;
                for (int k = 0; k < 32; ++k) {
                    const int l = g_lane_order == 1 ? 31 - k : (g_lane_order == 2 ? ((k < 16 ? 2 * k + 1 : 2 * (k - 16))) : k);
                    Fiber &f = M.fibers[w * 32 + l];
                    resume(f);
                    if (k == 0) s = f.state;
                    else if (f.state != s) die("lanes of one warp diverged across a collective / barrier / exit");
                }
                if (s == WAIT_WARP) { ++M.collectives; continue; }
                wstate[w] = s;
                break;
            }
        }
        if (all_done) break;
        bool any_runnable = false;
        for (int w = 0; w < warps; ++w) any_runnable |= wstate[w] == RUNNABLE;
        if (!any_runnable)                               // every live warp sits at the barrier: release it
            for (int w = 0; w < warps; ++w)
                if (wstate[w] == WAIT_BLOCK) wstate[w] = RUNNABLE;
    }
}

// kernel<<<grid, block, smem>>>(args...)  ==  simt::launch(grid, block, smem, [&] { kernel(args...); })
static inline void launch(unsigned grid, unsigned block, size_t smem, std::function<void()> body)
{
    M.body = std::move(body);
    M.grid_dim = dim3{grid, 1, 1};
    M.block_dim = dim3{block, 1, 1};
    M.dyn_smem.assign(smem + 64, 0xCD);                  // poisoned: kernels must not rely on zeroed shared memory
    for (unsigned b = 0; b < grid; ++b) {
        M.block_idx = dim3{b, 0, 0};
        run_block((int)block);
    }
}

static inline uint8_t *dyn_smem()
{
    uintptr_t p = reinterpret_cast<uintptr_t>(M.dyn_smem.data());
    return reinterpret_cast<uint8_t *>((p + 15) & ~(uintptr_t)15);
}

}  // namespace simt

#define threadIdx (simt::M.cur->tid)
#define blockIdx (simt::M.block_idx)
#define gridDim (simt::M.grid_dim)
#define blockDim (simt::M.block_dim)

// ---- intrinsics ------------------------------------------------------------------------------------------------------
using std::max;
using std::min;
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline void __stcs(T *p, T v) { *p = v; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __fdiv_rn(float a, float b) { return a / b; }       // IEEE single division (SSE), round to nearest
static inline int atomicAdd(int *p, int v) { int o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }

// 16-bit-pair integer instructions (VIMNMX.U16x2, VIMNMX3.U16x2, VIADDMNMX.U16x2), PRMT and SHF.R
static inline uint32_t __vminu2(uint32_t a, uint32_t b)
{
    return std::min(a & 0xffffu, b & 0xffffu) | (std::min(a >> 16, b >> 16) << 16);
}
static inline uint32_t __vmaxu2(uint32_t a, uint32_t b)
{
    return std::max(a & 0xffffu, b & 0xffffu) | (std::max(a >> 16, b >> 16) << 16);
}
static inline uint32_t __vimin3_u16x2(uint32_t a, uint32_t b, uint32_t c) { return __vminu2(__vminu2(a, b), c); }
static inline uint32_t __vimax3_u16x2(uint32_t a, uint32_t b, uint32_t c) { return __vmaxu2(__vmaxu2(a, b), c); }
static inline uint32_t __viaddmin_u16x2(uint32_t a, uint32_t b, uint32_t c)
{
    const uint32_t lo = ((a & 0xffffu) + (b & 0xffffu)) & 0xffffu, hi = ((a >> 16) + (b >> 16)) & 0xffffu;   // halves wrap
    return std::min(lo, c & 0xffffu) | (std::min(hi, c >> 16) << 16);
}
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s)
{
    const uint64_t src = (uint64_t)x | ((uint64_t)y << 32);
    uint32_t r = 0;
    for (int k = 0; k < 4; ++k) {
        const uint32_t n = (s >> (4 * k)) & 0xfu;
        uint32_t b = (uint32_t)(src >> (8 * (n & 7u))) & 0xffu;
        if (n & 8u) b = (b & 0x80u) ? 0xffu : 0u;          // sign replication mode
        r |= b << (8 * k);
    }
    return r;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift)
{
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> (shift & 31u));
}

static inline void __syncwarp(unsigned = 0xffffffffu) { simt::warp_exchange(0); }
static inline void __syncthreads() { simt::yield(simt::WAIT_BLOCK); }
static inline void need_full(unsigned mask) { if (mask != 0xffffffffu) simt::die("only full-mask collectives are emulated"); }
static inline unsigned __ballot_sync(unsigned mask, int pred)
{
    need_full(mask);
    const uint64_t *a = simt::warp_exchange(pred ? 1u : 0u);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r |= (unsigned)(a[l] & 1u) << l;
    return r;
}
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == 0xffffffffu; }
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src)
{
    need_full(mask);
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    const uint64_t *a = simt::warp_exchange(raw);
    T out;
    std::memcpy(&out, &a[src & 31], sizeof(T));
    return out;
}
static inline int __reduce_max_sync(unsigned mask, int v)
{
    need_full(mask);
    const uint64_t *a = simt::warp_exchange((uint64_t)(int64_t)v);
    int m = (int)(int64_t)a[0];
    for (int l = 1; l < 32; ++l) m = std::max(m, (int)(int64_t)a[l]);
    return m;
}
static inline int __reduce_add_sync(unsigned mask, int v)
{
    need_full(mask);
    const uint64_t *a = simt::warp_exchange((uint64_t)(int64_t)v);
    int s = 0;
    for (int l = 0; l < 32; ++l) s += (int)(int64_t)a[l];
    return s;
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { return (unsigned)__reduce_add_sync(mask, (int)v); }
