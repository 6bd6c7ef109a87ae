// bpp_common.cuh -- shared device-side definitions of libbpp_b200 (sm_100a).
//
// Vocabulary follows the reference: a *bin* holds a heightmap ("plain", envs/bpp0/space.py:22), the
// *box* is the item waiting to be placed (box_creator.preview(1)[0], envs/bpp0/bin3D.py:68-70).
#pragma once

#include <cstdint>
#include <cstddef>
#ifdef BPP_SIMT_EMUL
// tests/simt/simt_emul.h: the no-GPU test suite interprets these kernel sources lane by lane on the CPU (tests only;
// the library itself is always built by nvcc for sm_100a and has no CPU path)
#include "simt_emul.h"
#define BPP_LAUNCH(kernel, grid, block, smem, stream, ...) simt::launch((unsigned)(grid), (unsigned)(block), (smem), [&] { kernel(__VA_ARGS__); })
#define BPP_DYNAMIC_SMEM(name) uint8_t *name = simt::dyn_smem()
#else
#include <cuda_runtime.h>
#define BPP_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define BPP_DYNAMIC_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#endif

namespace bpp {

#ifndef BPP_WARPS_PER_BLOCK
#define BPP_WARPS_PER_BLOCK 8
#endif
constexpr int kWarpsPerBlock = BPP_WARPS_PER_BLOCK;      // default CTA size; the 10x10 bulk-copy instantiations pick 6 / 7 / 8 (launch_rv)
#ifndef BPP_SMALL_CTA_BINS_PER_SM
#define BPP_SMALL_CTA_BINS_PER_SM 160                  // envs with at least this many bins per SM run the 6-warp CTAs
#endif
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr int kMaxSide = 32;          // L, W <= 32 (reciprocal-multiply position decode, smem tile)
constexpr unsigned kFull = 0xffffffffu;

// Build switch (on by default; -DBPP_PACKED_SWEEP=0 gives the round-1 scalar sweep for A/B runs, tools/ab_bench.py):
// the mask sweep of the compile-time bin sizes evaluates TWO candidate positions per lane with the
// 16-bit-pair integer instructions (VIMNMX3.U16x2, VIADDMNMX.U16x2); see sweep_loop_packed below.
#ifndef BPP_PACKED_SWEEP
#define BPP_PACKED_SWEEP 1
#endif
constexpr bool kPackedSweep = BPP_PACKED_SWEEP != 0;
// Build switch (on by default): the placement check of step() (Space.check_box at the chosen position) runs in the lane-per-bin phase
// straight from the heightmap in global memory for footprints of up to 32 cells; the warp-per-bin phase then only
// raises the footprint.  Larger footprints (the (10,10,10) sentinels) keep the warp-cooperative check.
#ifndef BPP_LANE_CHECK
#define BPP_LANE_CHECK 1
#endif
constexpr bool kLaneCheck = BPP_LANE_CHECK != 0;
// Build switch (on by default): with rotation enabled, an item with x == y has two identical orientations (acktr/utils.py:76-88 computes
// the same mask twice); the second half of the mask is then a copy of the first instead of a second sweep.
#ifndef BPP_SQUARE_COPY
#define BPP_SQUARE_COPY 1
#endif
constexpr bool kSquareCopy = BPP_SQUARE_COPY != 0;
// Build switch (on by default): observation / mask rows leave with ordinary write-back stores instead of streaming (evict-first) ones,
// so that a consumer launched right after the step (the policy network) and the next step's overwrite find them in L2.
#ifndef BPP_PLAIN_STORES
#define BPP_PLAIN_STORES 1
#endif
#if BPP_PLAIN_STORES && !defined(BPP_SIMT_EMUL)
// stated in PTX: the output pointers went through pinned_ptr(), a plain C++ store would become a generic ST
__device__ __forceinline__ void store_out(float4 *p, float4 v)
{
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void store_out(float *p, float v)
{
    asm volatile("st.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
#else
template <typename T>
__device__ __forceinline__ void store_out(T *p, T v)
{
    __stcs(p, v);
}
#endif
// Build switch (on by default): the rollout instantiations for 10x10 bins move heightmap tiles, observation and mask
// rows with bulk asynchronous copies (see "Bulk asynchronous copies" below and the kernel in bpp_step.cu).
#ifndef BPP_TMA
#define BPP_TMA 1
#endif
constexpr bool kTmaBuild = BPP_TMA != 0;
// Bulk-copy path for 20x20 bins: log2 of the bins per warp batch (0 = off: the ordinary path).  A 20x20 bin needs 1.6 KB
// per float32 plane, so shared memory only holds small batches and the table of constant planes covers the values 0..7
// (larger item sides take ordinary stores).  Measured at 262 144 bins (profiles/r2j_ab_tma20.jsonl): 8-bin batches at
// 2 CTAs per SM 0.414-0.424 ms against 0.476-0.482 ms on the ordinary path; 4-bin x 3 CTAs 0.453, 16-bin x 1 CTA 0.560.
#ifndef BPP_TMA20
#define BPP_TMA20 3
#endif
// constant observation planes kept per CTA on the bulk-copy path: values 0 .. tma_lut_planes(cs) - 1
__host__ __device__ constexpr int tma_lut_planes(int cs) { return cs <= 10 ? cs + 1 : 8; }
// bins per warp batch of that path (log2): bounded by shared memory -- a batch's tiles and staged records live there.
// Measured on a B200 (profiles/r2c_ab_tma_variants.jsonl): without rotation 32-bin batches at 3 CTAs per SM beat 16-bin
// batches at 5 (the lane-per-bin phases are amortised over twice the bins, 0.239 vs 0.246 ms); with rotation -- twice
// the sweep work per bin -- 16-bin batches at 4-5 CTAs win (0.315 vs 0.345 ms).
// tma_bpw_log2 = the batch size of big envs; envs too small to fill the SMs with such batches take 16-bin batches.
#ifndef BPP_TMA_BPW_LOG2
#define BPP_TMA_BPW_LOG2 5
#endif
#ifndef BPP_TMA_BPW_LOG2_ROT
#define BPP_TMA_BPW_LOG2_ROT 4
#endif
__host__ __device__ constexpr int tma_bpw_log2(int rot) { return rot ? BPP_TMA_BPW_LOG2_ROT : BPP_TMA_BPW_LOG2; }
// bytes of the per-warp "pairs" array that the packed sweep reads (one 32-bit word per heightmap cell)
__host__ __device__ constexpr int pair_bytes(int cs) { return kPackedSweep && cs > 0 ? 4 * ((cs * cs + 15) & ~15) : 0; }

// Per-bin scalar state: 32 bytes = one DRAM sector, read with two broadcast 128-bit loads.
struct __align__(16) BinScalars {
    uint32_t box;       // x | y << 8 | z << 16  (the waiting item)
    int32_t cnt;        // len(space.boxes)
    int32_t vol;        // sum of placed item volumes (Space.get_ratio numerator, space.py:146-151)
    uint32_t episode;   // auto-resets so far
    uint32_t ctr;       // reset+step calls so far: Philox draw index
    int32_t pos;        // position inside the replayed / pre-cut episode
    int32_t ep_len;     // cut1/cut2: items in the current episode buffer
    uint32_t traj;      // replay: trajectory row of the current episode; cut1/cut2: episode buffer being played (0/1)
};
static_assert(sizeof(BinScalars) == 32, "BinScalars must be one 32-byte sector");

enum Mode : int { kModeStep = 0, kModeReset = 1, kModeObserve = 2 };

struct EnvParams {
    int N, L, W, H, A, A_pad, M, M_pad;      // M = A * (1 + rot): mask length
    int bpw_log2;                            // log2(bins per warp batch), chosen per launch
    int gen_kind;
    int n_items;
    const uint32_t *item_set;                // packed x | y<<8 | z<<16
    int n_traj, traj_cap;
    const uint32_t *traj;                    // [n_traj][traj_cap] packed
    const int32_t *traj_len;
    uint32_t first_mod, gn_mod;              // first_traj mod n_traj, global_n mod n_traj
    uint32_t seed_lo, seed_hi;
    unsigned long long global_offset;
    float lwh_f;                             // L*W*H (< 2^24, exact in float32)
    uint8_t *hm;                             // [N][A_pad] heightmaps, uint8
    BinScalars *sc;                          // [N]
    // cut1/cut2: two episode buffers per bin (the one being played and the pre-generated next one)
    uint32_t *ep_items;                      // [2][ep_buf_bins][ep_cap] packed items
    int32_t *ep_len2;                        // [2][ep_buf_bins] items per buffer
    int32_t *gen_list;                       // [ep_buf_bins] env-wide ids of bins whose spare buffer must be regenerated
    int32_t *gen_count;                      // number of entries in gen_list (appended by the step kernel)
    uint8_t *gen_scratch;                    // [N][gen_scratch_stride] generator work space
    int ep_cap;
    long long ep_buf_bins;                   // bins of the whole env (buffer stride), not of a launch slice
    long long ep_bin0;                       // first bin of this launch slice inside the env
    long long gen_scratch_stride;
    int lo[3], hi[3], rot;                   // item range (cut1 per axis, cut2 lo[0]/hi[0]); enable_rotation
    unsigned long long *metrics;             // {episodes, sum counter, sum vol}
};

struct StepIO {
    const long long *actions;   // [N] or null (reset / observe)
    float *obs;                 // [N][4A]
    float *mask;                // [N][M]
    float *reward;              // [N]
    uint8_t *done;              // [N]
    int32_t *counter;           // [N]
    float *ratio;               // [N]
    long long *sample;          // [N]
    int episode_inc;            // reset: 1 if the env had been reset before (LoadBoxCreator.index += 1)
    // Packed per-bin results (include/bpp_b200.h bpp_result, 16 bytes: reward f32 | placed volume i32 | counter i32 |
    // next sampled action i32): one 128-bit store per bin instead of reward / counter / ratio / sample; `done` stays a
    // separate byte array (the host reads it as np.bool_ without a pass over the records).
    uint4 *packed;              // [N] or null
    // actions as int32 with a stride (in int32 units) instead of `actions`: half the bytes on the wire, and stride 4 /
    // offset 3 reads them straight out of the previous step's records (policy-free rollouts through host memory)
    const int32_t *actions32;
    int actions32_stride;
    int n_steps;                // >= 1: env steps per launch; steps after the first take the previous step's sampled
                                // action (policy-free rollouts; rs / replay only, needs `sample` or `packed`)
};

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Stream layout (DESIGN.md "RNG streams"):
//   key = (seed lo, seed hi); ctr = (n >> 2, domain << 28 | aux, gid lo, gid hi); draw = word n & 3.
// Domain 0 / aux 0 is exactly the curand() sequence after curand_init(seed, gid, 0).
// ---------------------------------------------------------------------------------------------------
enum Domain : uint32_t { kDomItem = 0, kDomSample = 1, kDomGen = 2 };

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ uint32_t draw(uint32_t seed_lo, uint32_t seed_hi, uint32_t domain, uint32_t aux,
                                         unsigned long long gid, uint32_t n)
{
    const uint4 o = philox4x32_10(make_uint4(n >> 2, (domain << 28) | (aux & 0x0FFFFFFFu), (uint32_t)gid,
                                             (uint32_t)(gid >> 32)),
                                  make_uint2(seed_lo, seed_hi));
    const uint32_t w = n & 3u;
    return w == 0 ? o.x : (w == 1 ? o.y : (w == 2 ? o.z : o.w));
}

// Keeps a loop-invariant value in a register: ptxas otherwise re-derives such values (S2R, LDC, integer
// arithmetic) inside the per-position loop to save registers, which costs issue slots in an issue-bound kernel.
__device__ __forceinline__ int pinned(int v)
{
    asm volatile("" : "+r"(v));
    return v;
}
__device__ __forceinline__ uint32_t pinned(uint32_t v)
{
    asm volatile("" : "+r"(v));
    return v;
}
template <typename T>
__device__ __forceinline__ T *pinned_ptr(T *ptr)      // global pointers only (the address space is forgotten)
{
    asm volatile("" : "+l"(ptr));
    return ptr;
}

// 128-bit global accesses with default caching, stated in PTX so that pointers that went through pinned_ptr() (and
// lost their address space) still compile to LDG/STG rather than generic LD/ST.
__device__ __forceinline__ uint4 ld_global_u4(const uint4 *p)
{
#ifdef BPP_SIMT_EMUL
    return *p;
#else
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
#endif
}
__device__ __forceinline__ void st_global_u4(uint4 *p, uint4 v)
{
#ifdef BPP_SIMT_EMUL
    *p = v;
#else
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#endif
}

// ---------------------------------------------------------------------------------------------------
// Bulk asynchronous copies (TMA unit, SASS UBLKCP) and the mbarrier that tracks a load's completion (SYNCS).
// The step kernel moves whole batches of heightmap tiles with them and lets the copy engine write the observation /
// mask rows from shared memory, so that none of that traffic goes through the load/store unit.
//   bulk_load  : global -> shared, completes on an mbarrier (complete_tx::bytes); issued by ONE lane
//   bulk_store : shared -> global, tracked by the issuing thread's bulk async-group (commit / wait_group.read)
//   fence_async_smem: makes this thread's ordinary shared-memory writes visible to the async proxy (before a
//                bulk_store of that memory is issued by some lane after a __syncwarp)
// Addresses and sizes are multiples of 16 bytes.  Under BPP_SIMT_EMUL (CPU interpretation, tests only) the copies
// happen at issue time and the waits are no-ops; a __syncwarp() always separates the issue of a load from the first
// wait on it, which is what makes that model faithful to the order the hardware guarantees.
// ---------------------------------------------------------------------------------------------------
// A shared-memory address as the copy engine takes it (32-bit shared-window address; the kernel converts its per-warp
// pointers ONCE -- a generic-to-shared conversion per use costs two instructions per copy).
#ifdef BPP_SIMT_EMUL
typedef uint8_t *saddr_t;
__device__ __forceinline__ saddr_t saddr(void *p) { return static_cast<uint8_t *>(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t) { *bar = 0ull; }
__device__ __forceinline__ void fence_mbar_init() {}
__device__ __forceinline__ void fence_async_smem() {}
__device__ __forceinline__ void bulk_load(void *sdst, const void *gsrc, uint32_t bytes, unsigned long long *bar)
{
    std::memcpy(sdst, gsrc, bytes);
    *bar += 1ull;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    if (((*bar + 1ull) & 1ull) != (unsigned long long)(parity & 1u)) simt::die("mbar_wait before the load of that phase was issued");
}
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) { std::memcpy(gdst, ssrc, bytes); }
__device__ __forceinline__ void bulk_load_s(saddr_t sdst, const void *gsrc, uint32_t bytes, saddr_t bar)
{
    bulk_load(sdst, gsrc, bytes, reinterpret_cast<unsigned long long *>(bar));
}
__device__ __forceinline__ void mbar_wait_s(saddr_t bar, uint32_t parity) { mbar_wait(reinterpret_cast<unsigned long long *>(bar), parity); }
__device__ __forceinline__ void bulk_commit() {}
template <int N> __device__ __forceinline__ void bulk_wait_read() {}
__device__ __forceinline__ void bulk_wait_all() {}
#else
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(void *sdst, const void *gsrc, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    asm volatile("{\n .reg .pred p;\n BPP_MBAR_WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 " @p bra BPP_MBAR_DONE_%=;\n bra BPP_MBAR_WAIT_%=;\n BPP_MBAR_DONE_%=:\n}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
typedef uint32_t saddr_t;
__device__ __forceinline__ saddr_t saddr(void *p) { return smem_u32(p); }
__device__ __forceinline__ void bulk_load_s(saddr_t sdst, const void *gsrc, uint32_t bytes, saddr_t bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sdst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(saddr_t bar, uint32_t parity)
{
    asm volatile("{\n .reg .pred p;\n BPP_MBAR_WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 " @p bra BPP_MBAR_DONE_%=;\n bra BPP_MBAR_WAIT_%=;\n BPP_MBAR_DONE_%=:\n}"
                 ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_store(void *gdst, saddr_t ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
#endif

// floor(q / d) == (q * c_rcp16[d]) >> 16 for q < 1024, 1 <= d <= 32  (c_rcp16[d] = ceil(2^16 / d))
__constant__ uint32_t c_rcp16[kMaxSide + 1] = {0, 65536, 32768, 21846, 16384, 13108, 10923, 9363, 8192, 7282, 6554, 5958, 5462, 5042, 4682, 4370, 4096, 3856, 3641, 3450, 3277, 3121, 2979, 2850, 2731, 2622, 2521, 2428, 2341, 2260, 2185, 2115, 2048};
__device__ __forceinline__ uint32_t rcp16(uint32_t d) { return c_rcp16[d]; }

// ---------------------------------------------------------------------------------------------------
// Feasibility of one footprint.  `w` points at the window's first cell in a row-major tile of pitch W.
// Returns the resting height max_h or -1.  The reference's float tests max_area/area > 0.95 | 0.85 | 0.50
// (envs/bpp0/space.py:137-141, acktr/utils.py:28-32) are the exact integer tests 20*ma > 19*a,
// 20*ma > 17*a, 2*ma > a, i.e. ma >= floor(19a/20)+1, floor(17a/20)+1, floor(a/2)+1: three thresholds
// that depend only on the footprint area and are computed once per orientation.
// ---------------------------------------------------------------------------------------------------
// The three thresholds travel packed in one word: full | most << 10 | half << 20 (a <= 1024 keeps each < 1024).
__device__ __forceinline__ uint32_t pack_thresholds(int a)
{
    return (uint32_t)((19 * a) / 20 + 1) | ((uint32_t)((17 * a) / 20 + 1) << 10) | ((uint32_t)(a / 2 + 1) << 20);
}

template <int VARIANT>
__device__ __forceinline__ bool accept(int m, int cnt, int thr_full, int thr_most, int thr_half, int zmax, int c00,
                                       int c10, int c01, int c11)
{
    // full implies most implies half (thresholds 0.95 > 0.85 > 0.50), so the three-way rule is one comparison of the
    // count against the threshold selected by the number of corners below the reference height:
    //   0 corners below -> half, 1 below -> most, otherwise -> full      (space.py:137-142, utils.py:28-33)
    const bool fits = m <= zmax;                         // max_h + z <= H, space.py:134-135, utils.py:20-21
    if (VARIANT == 1) {                                  // corners against the corner maximum rm; needs rm == max_h
        const int rm = max(max(c00, c10), max(c01, c11));
        const int below = min(rm - c00, 1) + min(rm - c10, 1) + min(rm - c01, 1) + min(rm - c11, 1);
        const bool top = rm == m;
        const int need = (top && below == 0) ? thr_half : ((top && below == 1) ? thr_most : thr_full);
        return fits && below <= 1 && cnt >= need;        // sc = 4 - below >= 3, space.py:124-125
    } else {                                             // corners against the window maximum
        const int below = min(m - c00, 1) + min(m - c10, 1) + min(m - c01, 1) + min(m - c11, 1);
        const int need = below == 0 ? thr_half : (below == 1 ? thr_most : thr_full);
        return fits && cnt >= need;
    }
}

// Same with the number of corners below the window maximum already known (rule v2 only needs that).
template <int VARIANT>
__device__ __forceinline__ bool accept_cb(int m, int cnt, int thr_full, int thr_most, int thr_half, int zmax, int c00,
                                          int c10, int c01, int c11, int corners_below)
{
    if (VARIANT == 1) return accept<1>(m, cnt, thr_full, thr_most, thr_half, zmax, c00, c10, c01, c11);
    const int need = corners_below == 0 ? thr_half : (corners_below == 1 ? thr_most : thr_full);
    return m <= zmax && cnt >= need;
}

template <int VARIANT>
__device__ __forceinline__ bool accept(int m, int cnt, uint32_t thr, int z, int H, int c00, int c10, int c01, int c11)
{
    return accept<VARIANT>(m, cnt, (int)(thr & 1023u), (int)((thr >> 10) & 1023u), (int)(thr >> 20), H - z, c00, c10,
                           c01, c11);
}

// (max, number of cells equal to the max) of an x-by-y window; one pass, any size.
template <typename T>
__device__ __forceinline__ void window_stats(const T *w, int W, int x, int y, int &m, int &cnt)
{
    m = -1;
    cnt = 0;
#pragma unroll 1
    for (int r = 0; r < x; ++r) {
        const T *row = w + r * W;
#pragma unroll 1
        for (int c = 0; c < y; ++c) {
            const int h = (int)row[c];
            cnt = h > m ? 1 : cnt + (h == m);
            m = max(m, h);
        }
    }
}

// Same for rows of at most 5 cells, with the row unrolled: lanes of a warp that hold different footprints stay
// converged (the row loop runs to the warp's largest x), which is what the lane-per-bin placement check needs.
// Always reads 5 cells per row -- up to 4 bytes past the footprint's last cell, hence kHmSlack -- and masks the ones
// outside the footprint to -1.  Per row: maximum, cells equal to it, then the merge into the running (maximum, count).
constexpr int kHmSlack = 16;          // readable bytes that must follow the last heightmap row of an env
template <typename T>
__device__ __forceinline__ void window_stats_rows5(const T *w, int W, int x, int y, int &m, int &cnt)
{
    m = -1;
    cnt = 0;
#pragma unroll 1
    for (int r = 0; r < x; ++r) {
        const T *row = w + r * W;
        int h[5];
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const int v = (int)row[c];
            h[c] = c < y ? v : -1;
        }
        const int rm = max(max(max(h[0], h[1]), max(h[2], h[3])), h[4]);
        int below = 0;
#pragma unroll
        for (int c = 0; c < 5; ++c) below += min(rm - h[c], 1);
        const int rc = 5 - below;
        cnt = rm > m ? rc : cnt + (rm == m ? rc : 0);
        m = max(m, rm);
    }
}

// Same for a compile-time footprint: all cells go to registers, max tree, then the cells below the max are
// counted with min(m - h, 1); the four corners come out of the same registers.
template <int X, int Y>
__device__ __forceinline__ void window_stats_fixed(const uint8_t *w, int W, int &m, int &cnt, int &c00, int &c10,
                                                   int &c01, int &c11, int &corners_below)
{
    int h[X * Y];
#pragma unroll
    for (int r = 0; r < X; ++r)
#pragma unroll
        for (int c = 0; c < Y; ++c) h[r * Y + c] = w[r * W + c];
    m = h[0];
#pragma unroll
    for (int k = 1; k < X * Y; ++k) m = max(m, h[k]);
    int t[X * Y];                        // 1 where the cell is below the window maximum
#pragma unroll
    for (int k = 0; k < X * Y; ++k) t[k] = min(m - h[k], 1);
    int below = 0;
#pragma unroll
    for (int k = 0; k < X * Y; ++k) below += t[k];
    cnt = X * Y - below;
    c00 = h[0]; c01 = h[Y - 1]; c10 = h[(X - 1) * Y]; c11 = h[X * Y - 1];
    corners_below = t[0] + t[Y - 1] + t[(X - 1) * Y] + t[X * Y - 1];   // rule v2 counts corners == window maximum
}

// One thread evaluates one footprint on int32 heightmaps (batched check_box entry point): max_h or -1.
template <int VARIANT, typename T>
__device__ __forceinline__ int check_window(const T *w, int W, int x, int y, int z, int H)
{
    const int c00 = (int)w[0], c10 = (int)w[(x - 1) * W], c01 = (int)w[y - 1], c11 = (int)w[(x - 1) * W + y - 1];
    int m, cnt;
    window_stats<T>(w, W, x, y, m, cnt);
    return accept<VARIANT>(m, cnt, pack_thresholds(x * y), z, H, c00, c10, c01, c11) ? m : -1;
}

// Candidate-position loop of one orientation.  X, Y > 0: footprint known at compile time; 0: any size.
template <int VARIANT, int X, int Y, int CS>
__device__ __forceinline__ int sweep_loop(const uint8_t *tile, uint8_t *smask, int W_, int x, int y, int npos_, int rcp_,
                                          int dW_, int thr_full_, int thr_most_, int thr_half_, int zmax, int lane)
{
    // With the footprint (X, Y) and the bin side (CS) known at compile time everything but the tile contents and z
    // is a constant: candidate count, row length, the reciprocal of the position decode, the three thresholds.
    constexpr bool kConst = X > 0 && CS > 0;
    constexpr int kNj = kConst ? CS - Y + 1 : 1;
    const int W = CS > 0 ? CS : W_;
    const int npos = kConst ? (CS - X + 1) * kNj : npos_;
    const int rcp = kConst ? (65536 + kNj - 1) / kNj : rcp_;
    const int dW = kConst ? CS - kNj : dW_;
    const int thr_full = X > 0 ? (19 * X * Y) / 20 + 1 : thr_full_;
    const int thr_most = X > 0 ? (17 * X * Y) / 20 + 1 : thr_most_;
    const int thr_half = X > 0 ? (X * Y) / 2 + 1 : thr_half_;
    int total = 0;
#pragma unroll 1
    for (int p0 = 0; p0 < npos; p0 += 32) {
        const int q = p0 + lane;
        bool ok = false;
        if (q < npos) {
            const int i = (int)(((uint32_t)q * (uint32_t)rcp) >> 16);
            const int idx = i * dW + q;                  // i*W + (q - i*nj)
            const uint8_t *w = tile + idx;
            int m, cnt, c00, c10, c01, c11, cb;
            if (X > 0) {
                window_stats_fixed<(X > 0 ? X : 1), (Y > 0 ? Y : 1)>(w, W, m, cnt, c00, c10, c01, c11, cb);
            } else {
                c00 = w[0]; c10 = w[(x - 1) * W]; c01 = w[y - 1]; c11 = w[(x - 1) * W + y - 1];
                window_stats<uint8_t>(w, W, x, y, m, cnt);
                cb = min(m - c00, 1) + min(m - c10, 1) + min(m - c01, 1) + min(m - c11, 1);
            }
            ok = accept_cb<VARIANT>(m, cnt, thr_full, thr_most, thr_half, zmax, c00, c10, c01, c11, cb);
            smask[idx] = ok ? 1 : 0;
        }
        total += __popc(__ballot_sync(kFull, ok));
    }
    return total;
}

// ---------------------------------------------------------------------------------------------------
// Packed sweep (compile-time footprint and bin side): one lane evaluates the two horizontally adjacent
// candidate positions (i, j) and (i, j + 1), j even, at once, one position per 16-bit half of a register.
// `pairs[k]` holds the complemented heights (255 - tile[k]) | (255 - tile[k + 1]) << 16 of cell k and its right
// neighbour, so the X*Y loads of a lane deliver the windows of both positions.  With g = 255 - h:
//   window maximum        m = 255 - min g                          VIMNMX3.U16x2
//   cell below maximum    t = min(g - min g, 1), kept as t + 256    VIADDMNMX.U16x2 (g + (256 - min g), 257)
//   cells == maximum      cnt = X*Y - sum t;  corners below maximum cb = sum of the four corner t
// and the three-way threshold rule of acktr/utils.py:28-33 becomes one comparison per half:
//   feasible  <=>  cnt >= need(cb),  need = half | most | full for cb = 0 | 1 | >= 2
//             <=>  cnt + max(0, b0 - cb * (b0 - b1)) >= full,   b0 = full - half, b1 = full - most   (needs b0 >= 2 * b1)
// evaluated with a bias of 0x8000 per half so that bit 15 / bit 31 are the two answers and no borrow ever crosses
// the halves.  Positions are enumerated in increasing cell index (pair, then half), like the scalar sweep.
// This is rule v2 (acktr/utils.py:8-35), and for footprints of at most 40 cells -- all 16 unrolled ones -- also rule v1
// (space.py:111-144): count / area > 0.95 then leaves at most one cell below the maximum, so at least three corners sit
// at it and v1's early "fewer than three corners at the corner maximum" reject cannot fire; otherwise both rules need
// three or four corners AT THE WINDOW MAXIMUM, which makes the corner maximum equal to it and the two corner counts the
// same ("v1 vs v2"; the v1 golden masks and the v1 fuzz tests check exactly this path).
// ---------------------------------------------------------------------------------------------------
template <int X, int Y, int CS>
__device__ __forceinline__ int sweep_loop_packed(const uint32_t *pairs, uint8_t *smask, int zmax, int lane)
{
    constexpr int kNi = CS - X + 1, kNj = CS - Y + 1, kNpj = (kNj + 1) / 2, kWork = kNi * kNpj;
    constexpr int a = X * Y, full = (19 * a) / 20 + 1, most = (17 * a) / 20 + 1, half = a / 2 + 1;
    constexpr int b0 = full - half, b1 = full - most, K = b0 - b1;
    static_assert(b0 >= 2 * b1 && b1 >= 0, "threshold spacing assumed by the clamped bonus");
    constexpr uint32_t kRcp = (65536 + kNpj - 1) / kNpj;
    constexpr uint32_t kOnes = 0x00010001u;
    // t is carried with an offset of 256 per cell: sum t' = below + 256 a, cb' = cb + 1024
    // v = cnt + bonus - full + 0x8000 = (a + b0 - full + 0x8000 + 256 a) - sum t' - min(cb K, b0)
    //   with min(cb K, b0) = min(cb' K, b0 + 1024 K) - 1024 K
    constexpr uint32_t kBase = (uint32_t)(a + b0 - full + 0x8000 + 256 * a + 1024 * K) * kOnes;
    constexpr uint32_t kPenCap = (uint32_t)(b0 + 1024 * K) * kOnes;
    static_assert(a + b0 - full + 0x8000 + 256 * a + 1024 * K < 0x10000 && 1028 * K < 0x10000, "16-bit halves");
    // fits: m <= zmax  <=>  min g >= 255 - zmax  <=>  bit 15 of (min g + 0x8000 - 255 + zmax)
    const uint32_t zc = (uint32_t)(0x8000 - 255 + zmax) * kOnes;
    int mine = 0;
#pragma unroll 1
    for (int u0 = 0; u0 < kWork; u0 += 32) {
        const int u = u0 + lane;
        if (u < kWork) {
            const int i = (int)(((uint32_t)u * kRcp) >> 16), jp = u - i * kNpj;
            const int idx = i * CS + 2 * jp;
            uint32_t g[a];
            const uint32_t *w = pairs + idx;
#pragma unroll
            for (int r = 0; r < X; ++r)
#pragma unroll
                for (int c = 0; c < Y; ++c) g[r * Y + c] = w[r * CS + c];
            uint32_t gm = g[0];
#pragma unroll
            for (int k = 1; k + 1 < a; k += 2) gm = __vimin3_u16x2(gm, g[k], g[k + 1]);
            if ((a & 1) == 0) gm = __vminu2(gm, g[a - 1]);
            const uint32_t neg = 0x01000100u - gm;                     // 256 - min g per half, in 1..256
            uint32_t t[a];
#pragma unroll
            for (int k = 0; k < a; ++k) t[k] = __viaddmin_u16x2(g[k], neg, 0x01010101u);
            uint32_t below = 0;
#pragma unroll
            for (int k = 0; k < a; ++k) below += t[k];
            const uint32_t cb = t[0] + t[Y - 1] + t[(X - 1) * Y] + t[a - 1];
            const uint32_t v = kBase - below - __vminu2(cb * (uint32_t)K, kPenCap);
            const uint32_t f = gm + zc;
            const uint32_t ok2 = v & f & (2 * jp + 1 < kNj ? 0x80008000u : 0x00008000u);
            smask[idx] = (uint8_t)((ok2 >> 15) & 1u);
            smask[idx + 1] = (uint8_t)(ok2 >> 31);                     // a non-candidate cell when j + 1 == nj: stays 0
            mine += __popc(ok2);
        }
    }
    return __reduce_add_sync(kFull, mine);
}

// pairs[k] = (255 - tile[k]) | (255 - tile[k + 1]) << 16 for the A cells of a tile (A % 4 == 0; the word after the
// tile's last one must be readable: it only feeds the right neighbour of the last cell, which no valid position uses).
__device__ __forceinline__ void build_pairs(const uint8_t *tile, uint32_t *pairs, int A, int lane)
{
    const uint32_t *t32 = reinterpret_cast<const uint32_t *>(tile);
    uint4 *p4 = reinterpret_cast<uint4 *>(pairs);
#pragma unroll 1
    for (int k = lane; k < (A >> 2); k += 32) {
        const uint32_t w0 = t32[k], w1 = t32[k + 1];
        uint4 o;
        o.x = 0x00ff00ffu - __byte_perm(w0, 0u, 0x4140);
        o.y = 0x00ff00ffu - __byte_perm(w0, 0u, 0x4241);
        o.z = 0x00ff00ffu - __byte_perm(w0, 0u, 0x4342);
        o.w = 0x00ff00ffu - __byte_perm(__funnelshift_r(w0, w1, 24), 0u, 0x4140);   // last byte of w0, first of w1
        p4[k] = o;
    }
}

// Mask of one orientation: lanes sweep the (L-x+1)(W-y+1) candidate positions of the tile, write
// smask[i*W+j] = feasible, and return the number of feasible positions (warp-uniform).
// acktr/utils.py:54-57 / envs/bpp0/bin3D.py:85-88.  smask must be zero on entry; thr = pack_thresholds(x*y).
// The dispatch on the footprint happens once per orientation: footprints with sides 2..5 (the reference's item range,
// --item-size-range, acktr/arguments.py:30) have their own copy of the loop with the window unrolled; this is as much
// specialised code as the instruction cache takes (profiles/README.md: 44 KB of SASS runs at 76 % issue utilisation,
// a 64 KB variant with two positions per lane ran 2x slower).  Anything else takes the generic loop.
template <int VARIANT, int CS = 0>
__device__ __forceinline__ int sweep_orientation(const uint8_t *tile, uint8_t *smask, int L, int W, int H,
                                                 int x, int y, int z, uint32_t thr, int lane,
                                                 const uint32_t *pairs = nullptr)
{
    if (x > L || y > W) return 0;                        // range(width - x + 1) empty, utils.py:54-55
    if (z > H) return 0;                                 // max_h + z > H everywhere, utils.py:20-21
    const int zmax = pinned(H - z);                      // m + z <= H  <=>  m <= H - z
    const int nj = W - y + 1;
    const bool spec = x >= 2 && x <= 5 && y >= 2 && y <= 5;          // footprints with their own loop below
    const int npos = (CS > 0 && spec) ? 0 : pinned((L - x + 1) * nj);     // (those know these values at compile time)
    const int rcp = (CS > 0 && spec) ? 0 : pinned((int)rcp16((uint32_t)nj));
    const int dW = pinned(W - nj);
    const int thr_full = pinned((int)(thr & 1023u)), thr_most = pinned((int)((thr >> 10) & 1023u)),
              thr_half = pinned((int)(thr >> 20));
#define BPP_SW_CASE(X, Y) case ((X) - 2) * 4 + ((Y) - 2): \
        if constexpr (kPackedSweep && CS > 0 && (VARIANT == 2 || (X) * (Y) <= 40)) return sweep_loop_packed<X, Y, CS>(pairs, smask, zmax, lane); \
        else return sweep_loop<VARIANT, X, Y, CS>(tile, smask, W, x, y, npos, rcp, dW, thr_full, thr_most, thr_half, zmax, lane);
#define BPP_SW_ROW(X) BPP_SW_CASE(X, 2) BPP_SW_CASE(X, 3) BPP_SW_CASE(X, 4) BPP_SW_CASE(X, 5)
    if (spec) {
        switch ((x - 2) * 4 + (y - 2)) {
            BPP_SW_ROW(2) BPP_SW_ROW(3) BPP_SW_ROW(4) BPP_SW_ROW(5)
        default: break;
        }
    }
#undef BPP_SW_ROW
#undef BPP_SW_CASE
    return sweep_loop<VARIANT, 0, 0, 0>(tile, smask, W, x, y, npos, rcp, dW, thr_full, thr_most, thr_half, zmax, lane);
}

}  // namespace bpp
