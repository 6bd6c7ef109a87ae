// bpp_step.cu -- the fused per-step kernel of the batched 3D bin-packing environment (sm_100a).
//
// One warp owns one bin for the duration of a step (grid-stride over bins, persistent CTAs):
//   1. two broadcast 128-bit loads fetch the bin's 32-byte scalar record, 128-bit loads stage its
//      L x W uint8 heightmap into a shared-memory tile;
//   2. the action is decoded and the placement checked with the env-side rule (lanes cover the
//      footprint, warp max / ballot reductions), the footprint is raised in the tile, reward / done /
//      info are produced, a finished episode is reset in place and the next item is fetched
//      (Philox draw or replay row);
//   3. lanes sweep every candidate position of the new state (one lane per position, footprint read
//      from the tile) to build the ground-truth mask, both orientations when rotation is enabled;
//   4. heightmap, scalars, float32 observation and float32 mask leave with 128-bit stores.
// HBM-bound integer work: no tensor cores.  Reference behaviour cited inline (paths relative to the
// reference root).
#include "bpp_common.cuh"
#include "bpp_launch.h"

namespace bpp {

__device__ __forceinline__ float4 bytes_to_float4(uint32_t w)
{
    return make_float4((float)(w & 0xffu), (float)((w >> 8) & 0xffu), (float)((w >> 16) & 0xffu),
                       (float)(w >> 24));
}

// Next item of the bin's sequence (box_creator.preview(1)[0]).
__device__ __forceinline__ uint32_t fetch_item(const EnvParams &p, const BinScalars &s, unsigned long long gid,
                                               long long bin)
{
    switch (p.gen_kind) {
    case 0: {   // rs: RandomBoxCreator.generate_box_size, envs/bpp0/binCreator.py:38-40
        const uint32_t u = draw(p.seed_lo, p.seed_hi, kDomItem, 0, gid, s.ctr);
        return __ldg(p.item_set + __umulhi(u, (uint32_t)p.n_items));
    }
    case 1: {   // replay: LoadBoxCreator, binCreator.py:51-72 (sentinel (10,10,10) past the row, hard-coded)
        if (s.pos < __ldg(p.traj_len + s.traj)) return __ldg(p.traj + (size_t)s.traj * p.traj_cap + s.pos);
        return 10u | (10u << 8) | (10u << 16);
    }
    default: {  // cut1 / cut2: pre-cut episode buffer; sentinels cutCreator.py:113-115 / mdCreator.py:161
        if (s.pos < s.ep_len) return __ldg(p.ep_items + (size_t)bin * p.ep_cap + s.pos);
        return p.gen_kind == 2 ? ((uint32_t)p.L | ((uint32_t)p.W << 8) | ((uint32_t)p.H << 16))
                               : (10u | (10u << 8) | (10u << 16));
    }
    }
}

template <int ROT, int VARIANT, bool VEC4>
__global__ void __launch_bounds__(kThreads) bpp_step_kernel(const EnvParams p, const StepIO io, const int mode)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ unsigned long long blk_metrics[3];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t *tile = smem + (size_t)warp * (p.A_pad + p.M_pad);
    uint8_t *smask = tile + p.A_pad;
    const int L = p.L, W = p.W, H = p.H, A = p.A;

    if (threadIdx.x < 3) blk_metrics[threadIdx.x] = 0ull;
    __syncthreads();
    unsigned long long m_ep = 0, m_cnt = 0, m_vol = 0;

    for (long long bin = (long long)blockIdx.x * kWarpsPerBlock + warp; bin < p.N;
         bin += (long long)gridDim.x * kWarpsPerBlock) {
        const unsigned long long gid = p.global_offset + (unsigned long long)bin;

        // ---- 1. load state ------------------------------------------------------------------------
        BinScalars s;
        {
            const uint4 *sp = reinterpret_cast<const uint4 *>(p.sc + bin);
            const uint4 a = sp[0], b = sp[1];
            s.box = a.x; s.cnt = (int)a.y; s.vol = (int)a.z; s.episode = a.w;
            s.ctr = b.x; s.pos = (int)b.y; s.ep_len = (int)b.z; s.traj = b.w;
        }
        uint4 *tile4 = reinterpret_cast<uint4 *>(tile);
        if (mode == kModeReset) {
            for (int k = lane; k < p.A_pad / 16; k += 32) tile4[k] = make_uint4(0, 0, 0, 0);
        } else {
            const uint4 *g = reinterpret_cast<const uint4 *>(p.hm + (size_t)bin * p.A_pad);
            for (int k = lane; k < p.A_pad / 16; k += 32) tile4[k] = g[k];
        }
        for (int k = lane; k < p.M_pad / 16; k += 32)
            reinterpret_cast<uint4 *>(smask)[k] = make_uint4(0, 0, 0, 0);
        __syncwarp();

        // ---- 2. step --------------------------------------------------------------------------------
        float reward = 0.f;
        int done = 0;
        bool new_episode = false;
        if (mode == kModeStep) {
            const int bx = s.box & 0xff, by = (s.box >> 8) & 0xff, bz = (s.box >> 16) & 0xff;
            long long idx = io.actions[bin];
            bool flag = false;
            // envs/bpp0/bin3D.py:100-105: strict '>' -- action == A stays un-rotated and falls out of
            // bounds; without rotation an action > A trips the reference's assert, here it is infeasible.
            if (idx > A) { if (ROT) { idx -= A; flag = true; } else idx = -1; }
            const int x = flag ? by : bx, y = flag ? bx : by, z = bz;      // space.py:166-172
            int new_h = -1;
            if (idx >= 0 && idx < A) {
                const int lx = (int)idx / W, ly = (int)idx - lx * W;         // space.py:153-156
                if (lx + x <= L && ly + y <= W) {                            // space.py:112-113
                    // Space.check_box (space.py:116-144), lanes over the footprint
                    const uint8_t *w = tile + lx * W + ly;
                    const int a = x * y;
                    const uint32_t rcp = rcp16((uint32_t)y);
                    int m = -1;
                    for (int q = lane; q < a; q += 32) {
                        const int r = (int)(((uint32_t)q * rcp) >> 16), c = q - r * y;
                        m = max(m, (int)w[r * W + c]);
                    }
                    m = __reduce_max_sync(kFull, m);
                    int cnt = 0;
                    for (int q = lane; q < a; q += 32) {
                        const int r = (int)(((uint32_t)q * rcp) >> 16), c = q - r * y;
                        cnt += (w[r * W + c] == m);
                    }
                    cnt = __reduce_add_sync(kFull, cnt);
                    new_h = accept<1>(m, cnt, a, z, H, w[0], w[(x - 1) * W], w[y - 1], w[(x - 1) * W + y - 1]);
                    if (new_h >= 0) {                                        // update_height_graph, space.py:36-46
                        __syncwarp();
                        const uint8_t top = (uint8_t)(new_h + z);
                        for (int q = lane; q < a; q += 32) {
                            const int r = (int)(((uint32_t)q * rcp) >> 16), c = q - r * y;
                            tile[lx * W + ly + r * W + c] = top;
                        }
                    }
                }
            }
            if (new_h < 0) {                                                 // bin3D.py:108-112
                done = 1;
                if (lane == 0) {
                    if (io.reward) io.reward[bin] = 0.f;
                    if (io.done) io.done[bin] = 1;
                    if (io.counter) io.counter[bin] = s.cnt;
                    if (io.ratio) io.ratio[bin] = (float)((double)s.vol / p.lwh);
                }
                m_ep += 1; m_cnt += (unsigned long long)s.cnt; m_vol += (unsigned long long)s.vol;
                new_episode = true;
                s.episode += 1;
                __syncwarp();
                for (int k = lane; k < p.A_pad / 16; k += 32) tile4[k] = make_uint4(0, 0, 0, 0);
            } else {                                                         // bin3D.py:114-127
                const int v = bx * by * bz;
                s.cnt += 1; s.vol += v; s.pos += 1;
                reward = (float)(((double)v / p.lwh) * 10.0);               // bin3D.py:44-46,121
                if (lane == 0) {
                    if (io.reward) io.reward[bin] = reward;
                    if (io.done) io.done[bin] = 0;
                    if (io.counter) io.counter[bin] = s.cnt;
                    if (io.ratio) io.ratio[bin] = (float)((double)s.vol / p.lwh);   // space.py:146-151
                }
            }
        } else if (mode == kModeReset) {
            new_episode = true;
            s.episode += (uint32_t)io.episode_inc;
            if (!io.episode_inc) { s.episode = 0; s.ctr = 0; }
        }
        if (new_episode) {                                                   // PackingGame.reset, bin3D.py:55-59
            s.cnt = 0; s.vol = 0; s.pos = 0;
            if (p.gen_kind == 1) {                                           // trajectory of (gid, episode)
                if (mode == kModeReset && !io.episode_inc)
                    s.traj = (uint32_t)((p.first_mod + gid % (unsigned long long)p.n_traj) % (unsigned long long)p.n_traj);
                else { s.traj += p.gn_mod; if (s.traj >= (uint32_t)p.n_traj) s.traj -= (uint32_t)p.n_traj; }
            }
        }
        if (mode != kModeObserve) s.box = fetch_item(p, s, gid, bin);
        __syncwarp();

        // ---- 3. ground-truth mask of the new state -------------------------------------------------
        const int bx = s.box & 0xff, by = (s.box >> 8) & 0xff, bz = (s.box >> 16) & 0xff;
        int total = 0;
        if (io.mask || io.sample) {
            total = sweep_orientation<VARIANT>(tile, smask, L, W, H, bx, by, bz, lane);
            if (ROT) total += sweep_orientation<VARIANT>(tile, smask + A, L, W, H, by, bx, bz, lane);
            __syncwarp();
        }

        // ---- 4. write back ---------------------------------------------------------------------------
        if (mode != kModeObserve) {
            uint4 *g = reinterpret_cast<uint4 *>(p.hm + (size_t)bin * p.A_pad);
            for (int k = lane; k < p.A_pad / 16; k += 32) g[k] = tile4[k];
            if (lane == 0) {
                uint4 *sp = reinterpret_cast<uint4 *>(p.sc + bin);
                sp[0] = make_uint4(s.box, (uint32_t)s.cnt, (uint32_t)s.vol, s.episode);
                sp[1] = make_uint4(s.ctr + 1u, (uint32_t)s.pos, (uint32_t)s.ep_len, s.traj);
            }
        }
        if (io.obs) {                                                        // cur_observation, bin3D.py:61-66
            float *o = io.obs + (size_t)bin * 4 * A;
            const float fx = (float)bx, fy = (float)by, fz = (float)bz;
            if (VEC4) {
                const int q4 = A >> 2;
                float4 *o4 = reinterpret_cast<float4 *>(o);
                const uint32_t *t32 = reinterpret_cast<const uint32_t *>(tile);
                for (int k = lane; k < q4; k += 32) __stcs(o4 + k, bytes_to_float4(t32[k]));
                for (int k = lane; k < q4; k += 32) __stcs(o4 + q4 + k, make_float4(fx, fx, fx, fx));
                for (int k = lane; k < q4; k += 32) __stcs(o4 + 2 * q4 + k, make_float4(fy, fy, fy, fy));
                for (int k = lane; k < q4; k += 32) __stcs(o4 + 3 * q4 + k, make_float4(fz, fz, fz, fz));
            } else {
                for (int k = lane; k < A; k += 32) {
                    __stcs(o + k, (float)tile[k]); __stcs(o + A + k, fx);
                    __stcs(o + 2 * A + k, fy); __stcs(o + 3 * A + k, fz);
                }
            }
        }
        if (io.mask) {                                                       // all-ones fallback, utils.py:59-60,89-92
            float *mo = io.mask + (size_t)bin * p.M;
            if (VEC4) {
                float4 *m4 = reinterpret_cast<float4 *>(mo);
                const uint32_t *s32 = reinterpret_cast<const uint32_t *>(smask);
                for (int k = lane; k < (p.M >> 2); k += 32)
                    __stcs(m4 + k, total == 0 ? make_float4(1.f, 1.f, 1.f, 1.f) : bytes_to_float4(s32[k]));
            } else {
                for (int k = lane; k < p.M; k += 32) __stcs(mo + k, total == 0 ? 1.f : (float)smask[k]);
            }
        }
        if (io.sample) {   // uniform over mask == 1 (over all M entries under the fallback)
            const uint32_t u = draw(p.seed_lo, p.seed_hi, kDomSample, 0, gid, s.ctr);
            long long pick;
            if (total == 0) {
                pick = (long long)__umulhi(u, (uint32_t)p.M);
            } else {
                int r = (int)__umulhi(u, (uint32_t)total);
                pick = -1;
                for (int c0 = 0; c0 < p.M; c0 += 32) {
                    const bool bit = (c0 + lane < p.M) && smask[c0 + lane];
                    const unsigned b = __ballot_sync(kFull, bit);
                    const int n = __popc(b);
                    if (r < n) {
                        const int rank = __popc(b & ((1u << lane) - 1u));
                        const unsigned sel = __ballot_sync(kFull, bit && rank == r);
                        pick = c0 + __ffs(sel) - 1;
                        break;
                    }
                    r -= n;
                }
            }
            if (lane == 0) io.sample[bin] = pick;
        }
        __syncwarp();
    }

    // ---- episode metrics: registers -> shared -> one global atomic triple per CTA -------------------
    if (mode == kModeStep && p.metrics) {
        if (lane == 0 && m_ep) {
            atomicAdd(&blk_metrics[0], m_ep); atomicAdd(&blk_metrics[1], m_cnt); atomicAdd(&blk_metrics[2], m_vol);
        }
        __syncthreads();
        if (threadIdx.x < 3 && blk_metrics[threadIdx.x]) atomicAdd(p.metrics + threadIdx.x, blk_metrics[threadIdx.x]);
    }
}

template <int ROT, int VARIANT, bool VEC4>
static cudaError_t launch_t(const EnvParams &p, const StepIO &io, int mode, int grid, size_t smem, cudaStream_t st)
{
    bpp_step_kernel<ROT, VARIANT, VEC4><<<grid, kThreads, smem, st>>>(p, io, mode);
    return cudaGetLastError();
}

template <int ROT, int VARIANT, bool VEC4>
static cudaError_t occupancy_t(int *blocks, size_t smem)
{
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, bpp_step_kernel<ROT, VARIANT, VEC4>, kThreads, smem);
}

#define BPP_DISPATCH(FN, ...)                                                         \
    do {                                                                              \
        if (rot) {                                                                    \
            if (variant == 1) return vec4 ? FN<1, 1, true>(__VA_ARGS__) : FN<1, 1, false>(__VA_ARGS__); \
            return vec4 ? FN<1, 2, true>(__VA_ARGS__) : FN<1, 2, false>(__VA_ARGS__);  \
        }                                                                             \
        if (variant == 1) return vec4 ? FN<0, 1, true>(__VA_ARGS__) : FN<0, 1, false>(__VA_ARGS__);     \
        return vec4 ? FN<0, 2, true>(__VA_ARGS__) : FN<0, 2, false>(__VA_ARGS__);      \
    } while (0)

cudaError_t launch_step(const EnvParams &p, const StepIO &io, int mode, int rot, int variant, bool vec4,
                        int grid, size_t smem, cudaStream_t st)
{
    BPP_DISPATCH(launch_t, p, io, mode, grid, smem, st);
}

cudaError_t step_occupancy(int rot, int variant, bool vec4, size_t smem, int *blocks_per_sm)
{
    BPP_DISPATCH(occupancy_t, blocks_per_sm, smem);
}

}  // namespace bpp
