// bpp_step.cu -- the fused per-step kernel of the batched 3D bin-packing environment (sm_100a).
//
// One launch per env step (or several steps per launch: StepIO::n_steps).  Persistent CTAs of 8 warps; every warp takes
// batches of bins through three phases (details at the kernel): scalar work with one LANE per bin, heightmap-tile work
// with the whole WARP per bin (footprint update, mask sweep with two candidate positions per lane in 16-bit register
// halves and the footprint unrolled at compile time, float32 observation and mask rows), and the per-bin results again
// with one lane per bin.  On the rollout instantiations for 10x10 and 20x20 bins the copy engine moves the data: one
// bulk load / store per batch of tiles (cp.async.bulk + mbarrier), bulk stores of the staged rows and of constant planes.
// HBM-bound integer work by design (no tensor cores).  At the target workload (2^19 bins of 10x10): 0.229 ms per launch,
// 0.99 of the measured copy bandwidth by algorithmic bytes, 0.78 by the bytes really moved; issue slots 67 % busy, 24
// warps per SM -- see DESIGN.md section 3 and profiles/README.md.  Reference behaviour is cited inline (paths relative
// to the reference root).
#include <cstdlib>
#include <mutex>

#include "bpp_common.cuh"
#include "bpp_launch.h"
#include "bpp_params.h"

// Build switch: also the ordinary (LDG/STG) instantiations take the warp index and the item through a warp reduction, which
// moves their batch / row address arithmetic to the uniform datapath (54-60 instead of 64 registers).
#ifndef BPP_UNIFORM
#define BPP_UNIFORM 1
#endif
constexpr bool kUniformOps = BPP_UNIFORM != 0;
#ifndef BPP_MIN_BLOCKS
#define BPP_MIN_BLOCKS 4   // resident CTAs per SM the register allocator must allow (8 warps each)
#endif
// Resident CTAs per SM the register allocator must allow for the bulk-copy instantiations.  Their shared memory holds a
// whole batch of tiles and records, so the batch size decides: at 10x10, 32-bin batches fit 3 CTAs per SM, 16-bin batches 5
// (4 with rotation) -- both are built and the launch picks per env size (launch_rv below); at 20x20, 8-bin batches fit 2.
__host__ __device__ constexpr int tma_min_blocks(int rot, int bpw_log2, int cs = 10)
{
    return cs > 10 ? (bpw_log2 >= 4 ? 1 : (bpw_log2 >= 3 || rot ? 2 : 3))
                   : (bpw_log2 >= 5 ? 3 : (rot ? 4 : 5));
}

namespace bpp {

__device__ __forceinline__ float4 bytes_to_float4(uint32_t w)
{
    return make_float4((float)(w & 0xffu), (float)((w >> 8) & 0xffu), (float)((w >> 16) & 0xffu),
                       (float)(w >> 24));
}

// The three constant planes of an observation by ordinary stores (bulk-copy path, item sides beyond its table of planes).
__device__ __noinline__ void store_planes_slow(float *o, int A, int bx, int by, int bz, int lane)
{
    const float fx = (float)bx, fy = (float)by, fz = (float)bz;
    float4 *o4 = reinterpret_cast<float4 *>(o);
    const int q4 = A >> 2;
    for (int k = lane; k < q4; k += 32) {
        store_out(o4 + q4 + k, make_float4(fx, fx, fx, fx));
        store_out(o4 + 2 * q4 + k, make_float4(fy, fy, fy, fy));
        store_out(o4 + 3 * q4 + k, make_float4(fz, fz, fz, fz));
    }
}

constexpr int kRecQuads = 5;     // uint4 words per staged bin record
constexpr uint32_t kSentinel10 = 10u | (10u << 8) | (10u << 16);   // binCreator.py:62,69-71; mdCreator.py:161

// Item at position `pos` of replay trajectory `t` (LoadBoxCreator, envs/bpp0/binCreator.py:64-72).
__device__ __forceinline__ uint32_t replay_item(const EnvParams &p, uint32_t t, int pos)
{
    return pos < __ldg(p.traj_len + t) ? __ldg(p.traj + (size_t)t * p.traj_cap + pos) : kSentinel10;
}

// Item at position `pos` of one of the bin's two pre-cut episode buffers (cut1: cutCreator.py:112-128, sentinel
// bin_size :113-115; cut2: mdCreator.py:163-166, sentinel (10,10,10) :161).
__device__ __forceinline__ uint32_t episode_item(const EnvParams &p, long long bin, int buf, int pos, int ep_len)
{
    if (pos < ep_len) return __ldg(p.ep_items + ((size_t)buf * p.ep_buf_bins + (size_t)(p.ep_bin0 + bin)) * p.ep_cap + pos);
    return p.gen_kind == 2 ? ((uint32_t)p.L | ((uint32_t)p.W << 8) | ((uint32_t)p.H << 16)) : kSentinel10;
}

// The kernel works on batches of 32 bins per warp (fewer for small N) in three phases:
//   A  (lane = bin)   scalar work once per bin instead of once per lane: state record, action decode with bounds
//                     test, Philox draws, both candidates for the next item (placement succeeds / episode ends) and
//                     their count thresholds, the placement check itself for footprints of up to 32 cells;
//                     everything is staged as an 80-byte record per bin in shared memory;
//   B  (warp = bin)   for each bin of the batch in turn: record (two broadcast 128-bit loads), heightmap tile,
//                     footprint update (warp-cooperative check for larger footprints), mask sweep, tile /
//                     observation / mask write-back, action sampling;
//   A2 (lane = bin)   reward, done, info, new state record, coalesced stores, episode metrics, generator refill list.
//
// Data movement comes in two flavours.  TMA = false (any bin size, any mode): tiles arrive by 128-bit LDG (prefetched
// one bin ahead) and leave, like the observation and mask rows, by 128-bit STG.  TMA = true (the rollout
// instantiations for 10x10 bins): the copy engine does it -- ONE bulk load per batch brings the batch's contiguous
// tiles into shared memory (mbarrier-tracked; phase A's placement check then reads them there instead of from global
// memory), the bins are updated in place, ONE bulk store per batch writes them back; per bin the float32 heightmap
// plane and mask row are staged in shared memory and leave by bulk stores, and the three constant planes of the
// observation are copied straight out of a per-CTA table of constant planes -- no LDS / STG at all for 3/5 of the bytes.
//
// Template parameters: ROT = enable_rotation, VARIANT = mask rule, VEC4 = 128-bit output stores possible,
// CS = compile-time bin side (L == W == CS; 10 and 20 are the reference's configurations, 0 = any size read
// from EnvParams), FAST = the rollout case (mode == step and every output requested) with the checks folded.
template <int ROT, int VARIANT, bool VEC4, int CS, bool FAST, int TMA, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, TMA ? tma_min_blocks(ROT, TMA, CS) : BPP_MIN_BLOCKS) bpp_step_kernel(const EnvParams p, const StepIO io, const int mode_arg)
{
    static_assert(!TMA || (FAST && VEC4 && CS > 0), "the bulk-copy path is the rollout path of the compile-time bin sizes");
    constexpr int kWarpsPerBlock = WARPS, kThreads = WARPS * 32;      // (this instantiation's CTA size, see launch_rv)
    const int mode = FAST ? (int)kModeStep : mode_arg;
    const bool w_obs = FAST || io.obs != nullptr, w_mask = FAST || io.mask != nullptr;
    const bool w_sample = io.sample != nullptr || io.packed != nullptr;       // FAST: one of the two is there
    BPP_DYNAMIC_SMEM(smem);
    __shared__ unsigned long long blk_metrics[3];
    __shared__ float4 plane_lut[TMA ? 1 : 256];       // (v, v, v, v) for the constant x / y / z planes of the observation
    __shared__ __align__(8) unsigned long long tile_bar[kWarpsPerBlock];      // TMA: one mbarrier per warp

    const int lane = pinned((int)(threadIdx.x & 31));
    __builtin_assume(lane >= 0 && lane < 32);
    // the warp index through a warp reduction: the compiler then knows it -- and the batch, bin and row addresses derived
    // from it -- to be warp-uniform (uniform registers; the bulk copies take uniform operands)
    const int warp = (TMA || kUniformOps) ? __reduce_max_sync(kFull, (int)(threadIdx.x >> 5)) : (int)(threadIdx.x >> 5);
    const int L = CS ? CS : p.L, W = CS ? CS : p.W, H = p.H, A = CS ? CS * CS : p.A;
    const int A_pad = CS ? ((CS * CS + 15) & ~15) : p.A_pad;
    const int M = A * (ROT ? 2 : 1), M_pad = CS ? ((CS * CS * (ROT ? 2 : 1) + 15) & ~15) : p.M_pad;
    constexpr bool kPacked = kPackedSweep && CS > 0;      // (rule v1 takes the packed loops for the 16 unrolled footprints too)
    // Shared memory.  Without TMA, per warp: heightmap tile | mask tile | pairs; then the 32 staged records per warp.
    // With TMA, per warp: the batch's tiles | mask tile | pairs | f32 staging (heightmap plane, mask row) | records,
    // and once per CTA the table of constant observation planes (values 0 .. CS).
    constexpr int kTmaBins = TMA ? 1 << TMA : 32;        // TMA = log2 of the bins per warp batch (0: ordinary path)
    constexpr int kTmaTiles = TMA ? kTmaBins * ((CS * CS + 15) & ~15) : 0;
    constexpr int kTmaStage = TMA ? 4 * CS * CS * (ROT ? 3 : 2) : 0;
    constexpr int kTmaLutN = tma_lut_planes(CS);      // constant planes for the values 0 .. kTmaLutN - 1
    constexpr int kTmaLutBytes = TMA ? kTmaLutN * 4 * CS * CS : 0;
    const int warp_bytes = TMA ? kTmaTiles + M_pad + pair_bytes(CS) + kTmaStage + kTmaBins * (16 * kRecQuads)
                               : A_pad + M_pad + pair_bytes(CS);
    uint8_t *wbase = smem + kTmaLutBytes + pinned(warp * warp_bytes);
    uint8_t *tiles = wbase;                                             // TMA: the batch's tiles; else the one tile
    uint8_t *smask = wbase + (TMA ? kTmaTiles : A_pad);
    uint32_t *pairs = reinterpret_cast<uint32_t *>(smask + M_pad);
    float4 *stage_hm = reinterpret_cast<float4 *>(smask + M_pad + pair_bytes(CS));
    float4 *stage_mask = stage_hm + (CS * CS) / 4;
    uint4 *rec = TMA ? reinterpret_cast<uint4 *>(smask + M_pad + pair_bytes(CS) + kTmaStage)
                     : reinterpret_cast<uint4 *>(smem + kWarpsPerBlock * warp_bytes) + pinned(warp * (32 * kRecQuads));
    uint4 *smask4 = reinterpret_cast<uint4 *>(smask);
    // shared-window addresses for the copy engine, converted once (warp-uniform)
    const saddr_t sa_lut = saddr(smem), sa_tiles = saddr(tiles), sa_stage_hm = saddr(stage_hm), sa_stage_mask = saddr(stage_mask),
                  sa_bar = saddr(&tile_bar[warp]);
    const int nq_tile = A_pad >> 4, nq_mask = M_pad >> 4;     // 128-bit words per tile
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    const uint32_t rcpW = rcp16((uint32_t)W);
    const bool fresh_reset = mode == kModeReset && !io.episode_inc;

    if (threadIdx.x < 3) blk_metrics[threadIdx.x] = 0ull;
    if constexpr (TMA != 0) {
        float4 *lut = reinterpret_cast<float4 *>(smem);
        for (int k = threadIdx.x; k < kTmaLutBytes / 16; k += kThreads) {
            const float f = (float)(k / (CS * CS / 4));
            lut[k] = make_float4(f, f, f, f);
        }
        fence_async_smem();                               // the table is only ever read by the copy engine
        if (lane == 0) { mbar_init(&tile_bar[warp], 1); fence_mbar_init(); }
    } else {
        for (int v = threadIdx.x; v < 256; v += kThreads) {
            const float f = (float)v;
            plane_lut[v] = make_float4(f, f, f, f);
        }
    }
    __syncthreads();
    unsigned long long m_ep = 0, m_cnt = 0, m_vol = 0;
    uint32_t tile_phase = 0;                              // TMA: parity of the mbarrier phase the next load completes

    // bins per warp batch: 32 when there are plenty of bins, fewer (power of two) for small N so that the
    // batches still spread over all SMs instead of serialising 32 bins on a few warps
    const int bpw_log2 = p.bpw_log2;
    const long long n_batches = ((long long)p.N + (1 << bpw_log2) - 1) >> bpw_log2;
    const int n_steps = FAST ? max(io.n_steps, 1) : 1;
    // Several env steps per launch on an env small enough that every warp owns at most one batch: its tiles stay in shared
    // memory from the first step to the last (one bulk load, one bulk store per launch instead of per step).
    const bool resident = TMA && n_steps > 1 && n_batches <= (long long)gridDim.x * kWarpsPerBlock;
#pragma unroll 1
    for (int step = 0; step < n_steps; ++step) {
    for (long long batch = (long long)blockIdx.x * kWarpsPerBlock + warp; batch < n_batches;
         batch += (long long)gridDim.x * kWarpsPerBlock) {
        const long long bin0 = batch << bpw_log2;
        const int nb = (int)min((long long)(1 << bpw_log2), (long long)p.N - bin0);
        const bool have = lane < nb;
        const long long mybin = have ? bin0 + lane : bin0;
        const unsigned long long gid = p.global_offset + (unsigned long long)mybin;
        uint8_t *hm0 = pinned_ptr(p.hm + (size_t)bin0 * A_pad);
        if (TMA && (!resident || step == 0)) {
            // the batch's tiles: one bulk load (the bulk store of this warp's previous batch must have read them out)
            if (lane == 0) {
                if (n_steps > 1) bulk_wait_all(); else bulk_wait_read<0>();
                bulk_load_s(sa_tiles, hm0, (uint32_t)(nb * A_pad), sa_bar);
            }
        }

        // ---- phase A: lane = bin ----------------------------------------------------------------------
        {
        BinScalars s = {};
        long long action = -1;
        if (have) {
            const uint4 *sp = reinterpret_cast<const uint4 *>(p.sc + mybin);
            const uint4 a = sp[0], b = sp[1];
            s.box = a.x; s.cnt = (int)a.y; s.vol = (int)a.z; s.episode = a.w;
            s.ctr = b.x; s.pos = (int)b.y; s.ep_len = (int)b.z; s.traj = b.w;
            if (mode == kModeStep) {
                if (step == 0) action = io.actions32 ? (long long)io.actions32[(size_t)mybin * (size_t)io.actions32_stride] : io.actions[mybin];
                else action = io.sample ? io.sample[mybin] : (long long)(int)io.packed[mybin].w;   // written by this lane one step ago
            }
        }
        if (fresh_reset) { s.episode = 0; s.ctr = 0; }
        if (TMA && (!resident || step == 0)) {
            __syncwarp();
            mbar_wait_s(sa_bar, tile_phase);
            tile_phase ^= 1u;
        }
        // action decode, envs/bpp0/bin3D.py:100-105 (strict '>': action == A stays un-rotated and falls out
        // of bounds; without rotation an action > A trips the reference's assert, here it is infeasible),
        // Space.idx_to_position (space.py:153-156), orientation (space.py:166-172) and the bounds test of
        // check_box (space.py:112-115).  act = cell offset | x << 10 | y << 16 | in-bounds << 22
        uint32_t act = 0, thr_cur = 0;
        if (mode == kModeStep) {
            long long idx = action;
            bool flag = false;
            if (idx > A) { if (ROT) { idx -= A; flag = true; } else idx = -1; }
            if (idx >= 0 && idx < A) {
                const int bx = s.box & 0xff, by = (s.box >> 8) & 0xff;
                const int x = flag ? by : bx, y = flag ? bx : by;
                const int lx = (int)(((uint32_t)idx * rcpW) >> 16), ly = (int)idx - lx * W;
                if (lx + x <= L && ly + y <= W) {
                    act = (uint32_t)idx | ((uint32_t)x << 10) | ((uint32_t)y << 16) | (1u << 22);
                    thr_cur = pack_thresholds(x * y);
                    if (kLaneCheck && x * y <= 32) {
                        // Space.check_box (space.py:116-144) by this lane alone, on the bin's heightmap in global
                        // memory (the sectors are re-read from L1 by the tile load of phase B) -- with TMA, on the
                        // batch's tiles in shared memory.  The verdict and the new height of the footprint replace
                        // the in-bounds flag: act bit 22 = placed, bits 24..31 = max_h + z, bit 23 = decided here.
                        const uint8_t *w = (TMA ? tiles + lane * A_pad : p.hm + (size_t)mybin * A_pad) + idx;
                        const int z = (s.box >> 16) & 0xff;
                        const int c00 = w[0], c10 = w[(x - 1) * W], c01 = w[y - 1], c11 = w[(x - 1) * W + y - 1];
                        int m, cnt;
                        if (y <= 5) window_stats_rows5<uint8_t>(w, W, x, y, m, cnt);
                        else window_stats<uint8_t>(w, W, x, y, m, cnt);
                        const bool ok = accept<1>(m, cnt, thr_cur, z, H, c00, c10, c01, c11);
                        act = (act & 0x003fffffu) | (1u << 23) | (ok ? (1u << 22) | ((uint32_t)(m + z) << 24) : 0u);
                    }
                }
            }
        }
        // Philox draws of this step (counter n = s.ctr): item stream (rs) and action-sampler stream
        uint32_t u_samp = 0;
        uint32_t next_ok = s.box, next_new = s.box;      // waiting item if the placement succeeds / a new episode starts
        uint32_t traj_new = s.traj;
        int len_new = s.ep_len;
        if (have && mode != kModeObserve) {
            switch (p.gen_kind) {
            case 0: {   // rs: RandomBoxCreator.generate_box_size, envs/bpp0/binCreator.py:38-40 -- same draw either way
                const uint32_t u = draw(p.seed_lo, p.seed_hi, kDomItem, 0, gid, s.ctr);
                next_ok = next_new = __ldg(p.item_set + __umulhi(u, (uint32_t)p.n_items));
                break;
            }
            case 1: {   // replay: LoadBoxCreator.reset / generate_box_size, binCreator.py:51-72
                if (fresh_reset) traj_new = (uint32_t)((p.first_mod + gid % (unsigned long long)p.n_traj) % (unsigned long long)p.n_traj);
                else { traj_new = s.traj + p.gn_mod; if (traj_new >= (uint32_t)p.n_traj) traj_new -= (uint32_t)p.n_traj; }
                next_new = replay_item(p, traj_new, 0);
                if (mode == kModeStep) next_ok = replay_item(p, s.traj, s.pos + 1);
                break;
            }
            default: {  // cut1 / cut2: the episode being played and the pre-generated next one (bpp_gen.cu)
                const int cur = fresh_reset ? 1 : (int)(s.traj & 1u);    // a fresh reset starts on buffer 0
                traj_new = (uint32_t)(cur ^ 1);
                len_new = __ldg(p.ep_len2 + (size_t)(cur ^ 1) * p.ep_buf_bins + (size_t)(p.ep_bin0 + mybin));
                next_new = episode_item(p, mybin, cur ^ 1, 0, len_new);
                if (mode == kModeStep) next_ok = episode_item(p, mybin, cur, s.pos + 1, s.ep_len);
                break;
            }
            }
        }
        const uint32_t thr_ok = pack_thresholds((int)(next_ok & 0xff) * (int)((next_ok >> 8) & 0xff));
        const uint32_t thr_new = pack_thresholds((int)(next_new & 0xff) * (int)((next_new >> 8) & 0xff));
        if (have && w_sample) u_samp = draw(p.seed_lo, p.seed_hi, kDomSample, 0, gid, s.ctr);
        // stage the per-lane results: {box, act, thr_cur, next_ok} {next_new, thr_ok, thr_new, u_samp}
        // {cnt, vol, episode, ctr} {pos, ep_len, traj, traj_new} {len_new, result, -, -}
        if (!TMA || lane < kTmaBins) {
            uint4 *r = rec + lane * kRecQuads;
            r[0] = make_uint4(s.box, act, thr_cur, next_ok);
            r[1] = make_uint4(next_new, thr_ok, thr_new, u_samp);
            r[2] = make_uint4((uint32_t)s.cnt, (uint32_t)s.vol, s.episode, s.ctr);
            r[3] = make_uint4((uint32_t)s.pos, (uint32_t)s.ep_len, s.traj, traj_new);
            r[4] = make_uint4((uint32_t)len_new, 0u, 0u, 0u);
        }
        }
        __syncwarp();
        float *obs0 = w_obs ? pinned_ptr(io.obs + (size_t)bin0 * 4 * A) : nullptr;
        float *mask0 = w_mask ? pinned_ptr(io.mask + (size_t)bin0 * M) : nullptr;

        // ---- phase B: warp = bin ----------------------------------------------------------------------
        uint4 pre = zero4;        // next bin's tile words, prefetched while the current bin is processed
        const bool use_pre = !TMA && nq_tile <= 32 && mode != kModeReset;
        if (use_pre && lane < nq_tile) pre = ld_global_u4(reinterpret_cast<const uint4 *>(hm0) + lane);
#pragma unroll 1
        for (int b = 0; b < nb; ++b) {
            const uint4 r0 = rec[b * kRecQuads], r1 = rec[b * kRecQuads + 1];      // broadcast reads
            const uint32_t box = r0.x, bact = r0.y, b_thr_cur = r0.z, b_ok = r0.w;
            const uint32_t b_new = r1.x, b_thr_ok = r1.y, b_thr_new = r1.z, b_us = r1.w;
            uint8_t *ghm = hm0 + (unsigned)b * (unsigned)A_pad;
            uint8_t *tile = TMA ? tiles + (unsigned)b * (unsigned)A_pad : tiles;   // TMA: updated in place
            uint4 *tile4 = reinterpret_cast<uint4 *>(tile);
            if (!TMA) {
                if (use_pre) {
                    if (lane < nq_tile) tile4[lane] = pre;
                    if (b + 1 < nb && lane < nq_tile) pre = ld_global_u4(reinterpret_cast<const uint4 *>(ghm + A_pad) + lane);
                } else {
#pragma unroll 1
                    for (int k = lane; k < nq_tile; k += 32)
                        tile4[k] = mode == kModeReset ? zero4 : ld_global_u4(reinterpret_cast<const uint4 *>(ghm) + k);
                }
            }
            if (!TMA) {
#pragma unroll 1
                for (int k = lane; k < nq_mask; k += 32) smask4[k] = zero4;
                __syncwarp();
            }

            bool succ = false;
            if (mode == kModeStep) {
                if (kLaneCheck && (bact & (1u << 23))) {
                    // decided in phase A: raise the footprint (update_height_graph, space.py:36-46)
                    succ = (bact >> 22) & 1u;
                    if (succ) {
                        const int x = (bact >> 10) & 63, y = (bact >> 16) & 63, a = x * y;
                        const int r = (int)(((uint32_t)lane * rcp16((uint32_t)y)) >> 16);
                        if (lane < a) tile[(bact & 1023u) + r * (W - y) + lane] = (uint8_t)(bact >> 24);
                    }
                } else if ((bact >> 22) & 1u) {
                    // Space.check_box (space.py:116-144) with the lanes spread over the footprint
                    const int x = (bact >> 10) & 63, y = (bact >> 16) & 63, z = (box >> 16) & 0xff;
                    uint8_t *w = tile + (bact & 1023u);
                    const int a = x * y;
                    const uint32_t rcp = rcp16((uint32_t)y);
                    const int c00 = w[0], c10 = w[(x - 1) * W], c01 = w[y - 1], c11 = w[(x - 1) * W + y - 1];
                    if (a <= 32) {
                        const int r = (int)(((uint32_t)lane * rcp) >> 16), off = r * (W - y) + lane;
                        const int h = lane < a ? (int)w[off] : -1;
                        const int m = __reduce_max_sync(kFull, h);
                        const int cnt = __popc(__ballot_sync(kFull, h == m));
                        succ = accept<1>(m, cnt, b_thr_cur, z, H, c00, c10, c01, c11);
                        if (succ) {                                          // update_height_graph, space.py:36-46
                            __syncwarp();
                            if (lane < a) w[off] = (uint8_t)(m + z);
                        }
                    } else {
                        int m = -1;
#pragma unroll 1
                        for (int q = lane; q < a; q += 32) {
                            const int r = (int)(((uint32_t)q * rcp) >> 16);
                            m = max(m, (int)w[r * (W - y) + q]);
                        }
                        m = __reduce_max_sync(kFull, m);
                        int cnt = 0;
#pragma unroll 1
                        for (int q = lane; q < a; q += 32) {
                            const int r = (int)(((uint32_t)q * rcp) >> 16);
                            cnt += (w[r * (W - y) + q] == m);
                        }
                        cnt = __reduce_add_sync(kFull, cnt);
                        succ = accept<1>(m, cnt, b_thr_cur, z, H, c00, c10, c01, c11);
                        if (succ) {
                            __syncwarp();
#pragma unroll 1
                            for (int q = lane; q < a; q += 32) {
                                const int r = (int)(((uint32_t)q * rcp) >> 16);
                                w[r * (W - y) + q] = (uint8_t)(m + z);
                            }
                        }
                    }
                }
                if (!succ) {                                                     // auto-reset, shmem_vec_env.py:127-130
                    __syncwarp();
#pragma unroll 1
                    for (int k = lane; k < nq_tile; k += 32) tile4[k] = zero4;  // Space(*bin_size), bin3D.py:57
                }
            }
            const uint32_t nbox_v = mode == kModeObserve ? box : (succ ? b_ok : b_new);
            const uint32_t nbox = (TMA || kUniformOps) ? __reduce_max_sync(kFull, nbox_v) : nbox_v;     // (the same in every lane) warp-uniform for the compiler
            const uint32_t nthr = mode == kModeObserve ? pack_thresholds((int)(box & 0xff) * (int)((box >> 8) & 0xff))
                                                       : (succ ? b_thr_ok : b_thr_new);
            __syncwarp();
            if (TMA) {      // (the previous bin's action sampler has read the mask tile: every lane is past the barrier above)
#pragma unroll 1
                for (int k = lane; k < nq_mask; k += 32) smask4[k] = zero4;
                if (!kPacked) __syncwarp();
            }

            // ground-truth mask of the new state
            const int bx = nbox & 0xff, by = (nbox >> 8) & 0xff, bz = (nbox >> 16) & 0xff;
            int total = 0;
            if (w_mask || w_sample) {
                if (kPacked) {
                    build_pairs(tile, pairs, A, lane);
                    __syncwarp();
                }
                // with rotation, an item with x == y has two identical orientations: sweep once and copy the half
                // (one call site for both cases: the specialised sweep loops must not be duplicated in the binary)
                const bool square = kSquareCopy && ROT && VEC4 && bx == by;
#pragma unroll 1
                for (int o = 0; o < (ROT ? (square ? 1 : 2) : 1); ++o)
                    total += sweep_orientation<VARIANT, CS>(tile, smask + o * A, L, W, H, o ? by : bx, o ? bx : by, bz, nthr, lane, pairs);
                if (square) {
                    __syncwarp();
                    const uint32_t *s32 = reinterpret_cast<const uint32_t *>(smask);
                    uint32_t *d32 = reinterpret_cast<uint32_t *>(smask + A);
#pragma unroll 1
                    for (int k = lane; k < (A >> 2); k += 32) d32[k] = s32[k];
                    total *= 2;
                }
                if (TMA && lane == 0) bulk_wait_read<0>();         // the previous bin's rows have left the staging area
                __syncwarp();
            }

            // write back
            if (TMA) {
                // cur_observation, bin3D.py:61-66, and the mask with its all-ones fallback, utils.py:59-60,89-92:
                // the float32 heightmap plane and the mask row are staged, then five bulk stores (one lane) write the
                // bin's rows; the three constant planes come from the table of constant planes.
                const uint32_t *t32 = reinterpret_cast<const uint32_t *>(tile);
                const uint32_t *s32 = reinterpret_cast<const uint32_t *>(smask);
#pragma unroll 1
                for (int k = lane; k < (A >> 2); k += 32) stage_hm[k] = bytes_to_float4(t32[k]);
#pragma unroll 1
                for (int k = lane; k < (M >> 2); k += 32)
                    stage_mask[k] = total == 0 ? make_float4(1.f, 1.f, 1.f, 1.f) : bytes_to_float4(s32[k]);
                fence_async_smem();
                __syncwarp();
                float *o = obs0 + (unsigned)b * (unsigned)(4 * A);
                float *mo = mask0 + (unsigned)b * (unsigned)M;
                const bool in_lut = max(max(bx, by), bz) < kTmaLutN;
                if (lane == 0) {
                    bulk_store(o, sa_stage_hm, 4 * A);
                    if (in_lut) {
                        bulk_store(o + A, sa_lut + bx * (4 * A), 4 * A);
                        bulk_store(o + 2 * A, sa_lut + by * (4 * A), 4 * A);
                        bulk_store(o + 3 * A, sa_lut + bz * (4 * A), 4 * A);
                    }
                    bulk_store(mo, sa_stage_mask, 4 * M);
                    bulk_commit();
                }
                if (!in_lut) store_planes_slow(o, A, bx, by, bz, lane);   // item sides beyond the table (never with the reference's item sets)
            } else {
            if (mode != kModeObserve)
#pragma unroll 1
                for (int k = lane; k < nq_tile; k += 32) st_global_u4(reinterpret_cast<uint4 *>(ghm) + k, tile4[k]);
            if (w_obs) {                                                        // cur_observation, bin3D.py:61-66
                float *o = obs0 + (unsigned)b * (unsigned)(4 * A);
                const float fx = (float)bx, fy = (float)by, fz = (float)bz;
                if (VEC4) {
                    const int q4 = A >> 2;
                    float4 *o4 = reinterpret_cast<float4 *>(o);
                    const uint32_t *t32 = reinterpret_cast<const uint32_t *>(tile);
#pragma unroll 1
                    for (int k = lane; k < q4; k += 32) {
                        store_out(o4 + k, bytes_to_float4(t32[k]));
                        store_out(o4 + q4 + k, plane_lut[bx]);
                        store_out(o4 + 2 * q4 + k, plane_lut[by]);
                        store_out(o4 + 3 * q4 + k, plane_lut[bz]);
                    }
                } else {
#pragma unroll 1
                    for (int k = lane; k < A; k += 32) {
                        store_out(o + k, (float)tile[k]); store_out(o + A + k, fx);
                        store_out(o + 2 * A + k, fy); store_out(o + 3 * A + k, fz);
                    }
                }
            }
            if (w_mask) {                                                       // all-ones fallback, utils.py:59-60,89-92
                float *mo = mask0 + (unsigned)b * (unsigned)M;
                if (VEC4) {
                    float4 *m4 = reinterpret_cast<float4 *>(mo);
                    const uint32_t *s32 = reinterpret_cast<const uint32_t *>(smask);
#pragma unroll 1
                    for (int k = lane; k < (M >> 2); k += 32)
                        store_out(m4 + k, total == 0 ? make_float4(1.f, 1.f, 1.f, 1.f) : bytes_to_float4(s32[k]));
                } else {
#pragma unroll 1
                    for (int k = lane; k < M; k += 32) store_out(mo + k, total == 0 ? 1.f : (float)smask[k]);
                }
            }
            }
            uint32_t pick = 0;
            if (w_sample) {   // uniform over mask == 1 (over all M entries under the fallback)
                if (total == 0) {
                    pick = __umulhi(b_us, (uint32_t)M);
                } else {
                    int r = (int)__umulhi(b_us, (uint32_t)total);
#pragma unroll 1
                    for (int c0 = 0; c0 < M; c0 += 32) {
                        const bool bit = (c0 + lane < M) && smask[c0 + lane];
                        const unsigned bal = __ballot_sync(kFull, bit);
                        const int n = __popc(bal);
                        if (r < n) {
                            const int rank = __popc(bal & ((1u << lane) - 1u));
                            const unsigned sel = __ballot_sync(kFull, bit && rank == r);
                            pick = (uint32_t)(c0 + __ffs(sel) - 1);
                            break;
                        }
                        r -= n;
                    }
                }
            }
            if (lane == 0) reinterpret_cast<uint32_t *>(rec + b * kRecQuads + 4)[1] = pick | (succ ? 0x80000000u : 0u);
            if (!TMA) __syncwarp();     // (bulk-copy path: nothing of the next bin touches this bin's scratch before its first barrier)
        }
        if (TMA) {
            // the batch's tiles, updated in place, go back with one bulk store
            fence_async_smem();
            __syncwarp();
            if (lane == 0 && (!resident || step == n_steps - 1)) {
                bulk_store(hm0, sa_tiles, (uint32_t)(nb * A_pad));
                bulk_commit();
            }
        }

        // ---- phase A2: lane = bin ---------------------------------------------------------------------
        if (mode == kModeObserve) continue;
        BinScalars s;
        uint32_t next_ok, next_new, traj_new, my_res;
        int len_new;
        {
            const uint4 *r = rec + (TMA ? (lane & (kTmaBins - 1)) : lane) * kRecQuads;
            const uint4 a = r[0], b1 = r[1], c = r[2], d = r[3], e4 = r[4];
            s.box = a.x; next_ok = a.w; next_new = b1.x;
            s.cnt = (int)c.x; s.vol = (int)c.y; s.episode = c.z; s.ctr = c.w;
            s.pos = (int)d.x; s.ep_len = (int)d.y; s.traj = d.z; traj_new = d.w;
            len_new = (int)e4.x; my_res = e4.y;
        }
        const bool succ = my_res >> 31;
        bool finished = false;
        if (mode == kModeStep) {
            // reward = 10*v/LWH and ratio = vol/LWH: the reference divides in float64 and VecPyTorch rounds
            // to float32 (bin3D.py:44-46,121; space.py:146-151; acktr/envs.py:192).  10*v, vol and LWH are
            // integers < 2^24, so one IEEE float32 division gives the same float32 (no double rounding can
            // cross a float32 midpoint: tests/test_oracle_golden.py::test_f32_division_equals_reference_route).
            float reward = 0.f;
            if (succ) {                                                          // bin3D.py:114-127
                const int v = (int)(s.box & 0xff) * (int)((s.box >> 8) & 0xff) * (int)((s.box >> 16) & 0xff);
                s.cnt += 1; s.vol += v; s.pos += 1;
                reward = __fdiv_rn((float)(10 * v), p.lwh_f);
            }
            finished = have && !succ;                                            // bin3D.py:108-112
            if (have) {
                if (io.packed) {
                    st_global_u4(io.packed + mybin, make_uint4(__float_as_uint(reward), (uint32_t)s.vol, (uint32_t)s.cnt, my_res & 0x7fffffffu));
                } else {
                    if (FAST || io.reward) io.reward[mybin] = reward;
                    if (FAST || io.counter) io.counter[mybin] = s.cnt;
                    if (FAST || io.ratio) io.ratio[mybin] = __fdiv_rn((float)s.vol, p.lwh_f);
                }
                if (FAST || io.done) io.done[mybin] = succ ? 0 : 1;
            }
            const unsigned fin = __ballot_sync(kFull, finished);
            if (fin) {
                m_ep += (unsigned long long)__popc(fin);
                m_cnt += (unsigned long long)__reduce_add_sync(kFull, finished ? s.cnt : 0);
                m_vol += (unsigned long long)__reduce_add_sync(kFull, finished ? s.vol : 0);
            }
        }
        if (have) {
            if (mode == kModeReset || finished) {                                // PackingGame.reset, bin3D.py:55-59
                if (!fresh_reset) s.episode += 1;
                s.cnt = 0; s.vol = 0; s.pos = 0; s.traj = traj_new; s.ep_len = len_new;
                s.box = next_new;
            } else {
                s.box = next_ok;
            }
            if (io.sample) io.sample[mybin] = (long long)(my_res & 0x7fffffffu);
            uint4 *sp = reinterpret_cast<uint4 *>(p.sc + mybin);
            sp[0] = make_uint4(s.box, (uint32_t)s.cnt, (uint32_t)s.vol, s.episode);
            sp[1] = make_uint4(s.ctr + 1u, (uint32_t)s.pos, (uint32_t)s.ep_len, s.traj);
        }
        if (p.gen_kind >= 2 && !fresh_reset) {
            // bins that moved to their spare episode go on the refill list of bpp_gen_kernel (compacted, so that
            // the generator warps are full)
            const bool refill = have && (mode == kModeReset || finished);
            const unsigned rb = __ballot_sync(kFull, refill);
            if (rb) {
                int base = 0;
                if (lane == 0) base = atomicAdd(p.gen_count, __popc(rb));
                base = __shfl_sync(kFull, base, 0);
                if (refill) p.gen_list[base + __popc(rb & ((1u << lane) - 1u))] = (int32_t)(p.ep_bin0 + mybin);
            }
        }
        if (n_steps > 1) __syncwarp();      // the next step of this warp's bins reads what its lanes just wrote
    }
    }
    if (TMA && lane == 0) bulk_wait_all();  // shared memory must outlive the copy engine's reads

    // ---- episode metrics: registers -> shared -> one global atomic triple per CTA -------------------
    if (mode == kModeStep && p.metrics) {
        if (lane == 0 && m_ep) {
            atomicAdd(&blk_metrics[0], m_ep); atomicAdd(&blk_metrics[1], m_cnt); atomicAdd(&blk_metrics[2], m_vol);
        }
        __syncthreads();
        if (threadIdx.x < 3 && blk_metrics[threadIdx.x]) atomicAdd(p.metrics + threadIdx.x, blk_metrics[threadIdx.x]);
    }
}

// One launch: persistent CTAs (one wave of resident blocks, grid-stride over the warp batches).
template <int ROT, int VARIANT, bool VEC4, int CS, bool FAST, int TMA, int WARPS = bpp::kWarpsPerBlock>
static cudaError_t launch_t(const EnvParams &p, const StepIO &io, int mode, int num_sms, cudaStream_t st)
{
    // per device and instantiation: resident CTAs per SM for the shared-memory size of the last launch
    static int blocks_per_sm[64] = {0};
    static size_t smem_seen[64] = {0}, smem_opted[64] = {0};
    static std::mutex mu;                                 // two host threads may drive envs on one device
    constexpr int kWarpsPerBlock = WARPS, kThreads = WARPS * 32;
    const size_t smem = step_smem_bytes(p, TMA, ROT, WARPS);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    dev &= 63;
    int resident;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (blocks_per_sm[dev] == 0 || smem_seen[dev] != smem) {
            if (smem > 40 * 1024 && smem > smem_opted[dev]) {   // opt in to more than the default dynamic limit
                e = cudaFuncSetAttribute(bpp_step_kernel<ROT, VARIANT, VEC4, CS, FAST, TMA, WARPS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return e;
                smem_opted[dev] = smem;
            }
            int b = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, bpp_step_kernel<ROT, VARIANT, VEC4, CS, FAST, TMA, WARPS>, kThreads, smem);
            if (e != cudaSuccess) return e;
            blocks_per_sm[dev] = b > 0 ? b : 1;
            smem_seen[dev] = smem;
        }
        resident = blocks_per_sm[dev];
    }
    {   // experiment hook: BPP_MAX_CTAS=k caps the resident CTAs per SM the grid is sized for
        static const char *force = std::getenv("BPP_MAX_CTAS");
        if (force && force[0] >= '1' && force[0] <= '9' && resident > force[0] - '0') resident = force[0] - '0';
    }
    const long long cap = (long long)num_sms * resident;
    // largest power-of-two batch (<= 32 bins per warp; <= 2^TMA on the bulk-copy path, whose shared memory
    // holds a batch's tiles) that still gives at least 80 % of the resident warps a batch
    // (measured at 65 536 bins: 16 bins per warp beat 8 and 32)
    EnvParams q = p;
    q.bpw_log2 = TMA ? TMA : 5;
    while (q.bpw_log2 > 0 && (((long long)p.N + (1 << q.bpw_log2) - 1) >> q.bpw_log2) < (cap * kWarpsPerBlock * 4) / 5) --q.bpw_log2;
    if (q.bpw_log2 < 3) {
        // Small envs: that rule runs down to one or two bins per warp, where the lane-per-bin phases (about 1.5 bins'
        // worth of instructions per batch, executed by a whole warp) dominate the instruction volume.  Pick, among the
        // batch sizes up to 8, the one with the smallest max(instruction volume / issue rate, rounds x latency of one
        // batch) -- a model with measured constants (profiles/README.md, "Small envs"), not a law.
        const double i_a = 450.0, i_b = 300.0 * (ROT ? 1.6 : 1.0) * (double)(p.A > 0 ? p.A : 100) / 100.0;
        const double rate = 0.6 * 4.0 * (double)num_sms, dep = 6.0;      // issued warp-instructions per cycle; cycles per dependent instruction
        double best = -1.0;
        int best_log2 = q.bpw_log2;
        for (int l = q.bpw_log2; l <= 3 && (!TMA || l <= TMA); ++l) {
            const double nb = (double)(1 << l), batches_l = (double)(((long long)p.N + (1 << l) - 1) >> l);
            const double rounds = (double)(long long)((batches_l + (double)(cap * kWarpsPerBlock) - 1.0) / (double)(cap * kWarpsPerBlock));
            const double volume = batches_l * (i_a + nb * i_b) / rate, latency = rounds * (i_a + nb * i_b) * dep;
            const double t = volume > latency ? volume : latency;
            if (best < 0.0 || t < best) { best = t; best_log2 = l; }
        }
        q.bpw_log2 = best_log2;
    }
    {   // experiment hook: BPP_BPW_LOG2=k forces 2^k bins per warp batch
        static const char *force = std::getenv("BPP_BPW_LOG2");
        if (force && force[0] >= '0' && force[0] <= '5') q.bpw_log2 = force[0] - '0';
        if (TMA && q.bpw_log2 > TMA) q.bpw_log2 = TMA;
    }
    const long long batches = ((long long)p.N + (1 << q.bpw_log2) - 1) >> q.bpw_log2;
    const long long want = (batches + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const int grid = (int)(want < cap ? want : cap);
    BPP_LAUNCH((bpp_step_kernel<ROT, VARIANT, VEC4, CS, FAST, TMA, WARPS>), grid, kThreads, smem, st, q, io, mode);
    return cudaGetLastError();
}

template <int ROT, int VARIANT>
static cudaError_t launch_rv(const EnvParams &p, const StepIO &io, int mode, bool vec4, int num_sms, cudaStream_t st)
{
    const bool fast = mode == kModeStep && io.obs && io.mask && io.done &&
                      ((io.reward && io.counter && io.ratio && io.sample) || io.packed);
    const int cs = (vec4 && p.L == p.W && (p.L == 10 || p.L == 20)) ? p.L : 0;
    if (cs == 10 && fast && kTmaBuild) {
        // Bulk-copy path.  Batch size and CTA size per launch (measured on a B200: profiles/r2d_ab_trimmed.jsonl,
        // r2l_ab_cta_size.jsonl).  Without rotation: 32-bin batches in 8-warp CTAs (3 per SM) while such batches still give
        // 80 % of those warps work, else 16-bin batches in 6-warp CTAs.  With rotation (twice the sweep work per bin):
        // 16-bin batches, in 7-warp CTAs for big envs and 6-warp CTAs for mid-size ones.  The register budget per thread is
        // the same for every CTA size (tma_min_blocks); the smaller CTAs pack the SM's registers and shared memory with
        // a different number of warps (the occupancy query in launch_t sizes the grid) and spread a few thousand batches
        // more evenly over the SMs: -5 % on the rotation headline shape, -5 % on 65 536-bin launches against 8-warp CTAs.
        static const char *force = std::getenv("BPP_TMA_FORCE");      // test hook: "5" / "4" / "3" picks the big-env / mid-size / small-env instantiation
        const bool forced = force && force[0] >= '3' && force[0] <= '5';
        // Below BPP_SMALL_CTA_BINS_PER_SM bins per SM the batch rule of launch_t runs down to a few bins per warp and the
        // per-CTA setup (the table of constant planes) weighs as much as the packing: 8-warp CTAs stay (4096 bins: 9.0 us
        // against 9.6 us with 6-warp CTAs; level from 8192 to 24 576; 6-warp CTAs ahead from 32 768 bins,
        // profiles/r2l_ab_cta_threshold.jsonl).
        const bool small_ctas = forced ? force[0] == '4' : (long long)p.N >= (long long)num_sms * BPP_SMALL_CTA_BINS_PER_SM;
        if constexpr (ROT != 0) {
            const long long warps_big = (long long)num_sms * tma_min_blocks(1, 4) * 7;
            const bool big = forced ? force[0] == '5' : (((long long)p.N + 15) >> 4) >= warps_big;
            if (big) return launch_t<ROT, VARIANT, true, 10, true, 4, 7>(p, io, mode, num_sms, st);
            return small_ctas ? launch_t<ROT, VARIANT, true, 10, true, 4, 6>(p, io, mode, num_sms, st)
                              : launch_t<ROT, VARIANT, true, 10, true, 4, kWarpsPerBlock>(p, io, mode, num_sms, st);
        } else {
            constexpr int kBig = tma_bpw_log2(0);
            const long long warps_big = (long long)num_sms * tma_min_blocks(0, kBig) * kWarpsPerBlock;
            const bool big = forced ? force[0] == '5' : (((long long)p.N + (1 << kBig) - 1) >> kBig) >= (warps_big * 4) / 5;
            if (kBig > 4 && big) return launch_t<ROT, VARIANT, true, 10, true, kBig, kWarpsPerBlock>(p, io, mode, num_sms, st);
            return small_ctas ? launch_t<ROT, VARIANT, true, 10, true, 4, 6>(p, io, mode, num_sms, st)
                              : launch_t<ROT, VARIANT, true, 10, true, 4, kWarpsPerBlock>(p, io, mode, num_sms, st);
        }
    }
    if (cs == 10) return fast ? launch_t<ROT, VARIANT, true, 10, true, 0>(p, io, mode, num_sms, st)
                              : launch_t<ROT, VARIANT, true, 10, false, 0>(p, io, mode, num_sms, st);
    if (cs == 20 && fast && kTmaBuild && BPP_TMA20 > 0) return launch_t<ROT, VARIANT, true, 20, true, BPP_TMA20>(p, io, mode, num_sms, st);
    if (cs == 20) return fast ? launch_t<ROT, VARIANT, true, 20, true, 0>(p, io, mode, num_sms, st)
                              : launch_t<ROT, VARIANT, true, 20, false, 0>(p, io, mode, num_sms, st);
    return vec4 ? launch_t<ROT, VARIANT, true, 0, false, 0>(p, io, mode, num_sms, st)
                : launch_t<ROT, VARIANT, false, 0, false, 0>(p, io, mode, num_sms, st);
}

cudaError_t launch_step(const EnvParams &p, const StepIO &io, int mode, int rot, int variant, bool vec4, int num_sms,
                        cudaStream_t st)
{
    if (rot) return variant == 1 ? launch_rv<1, 1>(p, io, mode, vec4, num_sms, st)
                                 : launch_rv<1, 2>(p, io, mode, vec4, num_sms, st);
    return variant == 1 ? launch_rv<0, 1>(p, io, mode, vec4, num_sms, st)
                        : launch_rv<0, 2>(p, io, mode, vec4, num_sms, st);
}

}  // namespace bpp
