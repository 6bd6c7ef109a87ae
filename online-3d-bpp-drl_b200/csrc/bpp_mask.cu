// bpp_mask.cu -- stateless kernels: ground-truth masks from observations / raw heightmaps, batched
// check_box, state import/export in the reference's dtypes, Philox self-test.
#ifndef BPP_SIMT_EMUL
#include <curand_kernel.h>     // only for the stream-layout self-test at the end of this file
#endif

#include "bpp_common.cuh"
#include "bpp_launch.h"

namespace bpp {

// acktr/utils.py:37-94 on a batch: one warp per observation row.
// SRC = 0: obs float32[n][4A] (heightmap = channel 0, box = element 0 of channels 1..3, utils.py:42-47)
// SRC = 1: plain int32[n][A] + box int32[n][3]
template <int ROT, int VARIANT, int SRC>
__global__ void __launch_bounds__(kThreads) bpp_mask_kernel(const void *src, const int32_t *box, int n, int L, int W,
                                                            int H, int A_pad, int M_pad, float *mask)
{
    BPP_DYNAMIC_SMEM(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *tile = smem + (size_t)warp * (A_pad + M_pad);
    uint8_t *smask = tile + A_pad;
    const int A = L * W, M = A * (ROT ? 2 : 1);
    for (long long row = (long long)blockIdx.x * kWarpsPerBlock + warp; row < n;
         row += (long long)gridDim.x * kWarpsPerBlock) {
        int bx, by, bz;
        if (SRC == 0) {
            const float *o = static_cast<const float *>(src) + (size_t)row * 4 * A;
            for (int k = lane; k < A; k += 32) tile[k] = (uint8_t)(int)o[k];
            bx = (int)o[A]; by = (int)o[2 * A]; bz = (int)o[3 * A];
        } else {
            const int32_t *pl = static_cast<const int32_t *>(src) + (size_t)row * A;
            for (int k = lane; k < A; k += 32) tile[k] = (uint8_t)pl[k];
            bx = box[row * 3]; by = box[row * 3 + 1]; bz = box[row * 3 + 2];
        }
        for (int k = lane; k < M; k += 32) smask[k] = 0;
        __syncwarp();
        int total = 0;
        // a non-positive dimension has no reference behaviour (numpy raises); treat as no position
        if (bx > 0 && by > 0 && bx <= 255 && by <= 255) {
            const uint32_t thr = pack_thresholds(bx * by);
            total = sweep_orientation<VARIANT>(tile, smask, L, W, H, bx, by, bz, thr, lane);
            if (ROT) total += sweep_orientation<VARIANT>(tile, smask + A, L, W, H, by, bx, bz, thr, lane);
        }
        __syncwarp();
        float *mo = mask + (size_t)row * M;
        for (int k = lane; k < M; k += 32) mo[k] = total == 0 ? 1.f : (float)smask[k];
        __syncwarp();
    }
}

template <int SRC>
static cudaError_t launch_mask_t(const void *src, const int32_t *box, int n, int L, int W, int H, int rot, int variant,
                                 float *mask, int grid, cudaStream_t st)
{
    const int A = L * W, A_pad = (A + 15) & ~15, M = A * (rot ? 2 : 1), M_pad = (M + 15) & ~15;
    const size_t smem = (size_t)kWarpsPerBlock * (A_pad + M_pad);
    if (rot) {
        if (variant == 1) BPP_LAUNCH((bpp_mask_kernel<1, 1, SRC>), grid, kThreads, smem, st, src, box, n, L, W, H, A_pad, M_pad, mask);
        else BPP_LAUNCH((bpp_mask_kernel<1, 2, SRC>), grid, kThreads, smem, st, src, box, n, L, W, H, A_pad, M_pad, mask);
    } else {
        if (variant == 1) BPP_LAUNCH((bpp_mask_kernel<0, 1, SRC>), grid, kThreads, smem, st, src, box, n, L, W, H, A_pad, M_pad, mask);
        else BPP_LAUNCH((bpp_mask_kernel<0, 2, SRC>), grid, kThreads, smem, st, src, box, n, L, W, H, A_pad, M_pad, mask);
    }
    return cudaGetLastError();
}

cudaError_t launch_mask_from_obs(const float *obs, int n, int L, int W, int H, int rot, int variant, float *mask,
                                 int grid, cudaStream_t st)
{
    return launch_mask_t<0>(obs, nullptr, n, L, W, H, rot, variant, mask, grid, st);
}

cudaError_t launch_mask_from_plain(const int32_t *plain, const int32_t *box, int n, int L, int W, int H, int rot,
                                   int variant, float *mask, int grid, cudaStream_t st)
{
    return launch_mask_t<1>(plain, box, n, L, W, H, rot, variant, mask, grid, st);
}

// envs/bpp0/space.py:111-144 / acktr/utils.py:8-35, one thread per query (x, y, lx, ly, z).
__global__ void bpp_check_box_kernel(const int32_t *plain, const int32_t *query, int n, int L, int W, int H,
                                     int variant, int32_t *out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int32_t *q = query + (size_t)t * 5;
    const int x = q[0], y = q[1], lx = q[2], ly = q[3], z = q[4];
    int r = -1;
    if (x > 0 && y > 0 && lx >= 0 && ly >= 0 && lx + x <= L && ly + y <= W) {
        const int32_t *w = plain + (size_t)t * L * W + lx * W + ly;
        r = variant == 1 ? check_window<1, int32_t>(w, W, x, y, z, H) : check_window<2, int32_t>(w, W, x, y, z, H);
    }
    out[t] = r;
}

cudaError_t launch_check_box(const int32_t *plain, const int32_t *query, int n, int L, int W, int H, int variant,
                             int32_t *out, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    BPP_LAUNCH((bpp_check_box_kernel), (n + 127) / 128, 128, 0, st, plain, query, n, L, W, H, variant, out);
    return cudaGetLastError();
}

// State in the reference's dtypes: Space.plain int32[N][A], next box int32[N][3], len(boxes), volume.
__global__ void bpp_export_kernel(EnvParams p, int32_t *plain, int32_t *box, int32_t *counter, int32_t *vol)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long cells = (long long)p.N * p.A;
    if (plain && t < cells) {
        const long long b = t / p.A;
        const int k = (int)(t - b * p.A);
        plain[t] = p.hm[b * p.A_pad + k];
    }
    if (t < p.N) {
        const BinScalars s = p.sc[t];
        if (box) { box[t * 3] = s.box & 0xff; box[t * 3 + 1] = (s.box >> 8) & 0xff; box[t * 3 + 2] = (s.box >> 16) & 0xff; }
        if (counter) counter[t] = s.cnt;
        if (vol) vol[t] = s.vol;
    }
}

__global__ void bpp_import_kernel(EnvParams p, const int32_t *plain, const int32_t *box, const int32_t *counter,
                                  const int32_t *vol, int32_t *bad)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long cells = (long long)p.N * p.A;
    if (plain && t < cells) {
        const long long b = t / p.A;
        const int k = (int)(t - b * p.A);
        const int32_t h = plain[t];
        if (h < 0 || h > p.H) { if (bad) atomicAdd(bad, 1); }      // a height outside [0, H] cannot come from Space.plain
        else p.hm[b * p.A_pad + k] = (uint8_t)h;
    }
    if (t < p.N) {
        BinScalars s = p.sc[t];
        if (box) s.box = (uint32_t)(box[t * 3] & 0xff) | ((uint32_t)(box[t * 3 + 1] & 0xff) << 8) |
                         ((uint32_t)(box[t * 3 + 2] & 0xff) << 16);
        if (counter) s.cnt = counter[t];
        if (vol) s.vol = vol[t];
        p.sc[t] = s;
    }
}

cudaError_t launch_export_state(const EnvParams &p, int32_t *plain, int32_t *box, int32_t *counter, int32_t *vol,
                                cudaStream_t st)
{
    const long long n = (long long)p.N * p.A;
    BPP_LAUNCH((bpp_export_kernel), (unsigned)((n + 255) / 256), 256, 0, st, p, plain, box, counter, vol);
    return cudaGetLastError();
}

cudaError_t launch_import_state(const EnvParams &p, const int32_t *plain, const int32_t *box, const int32_t *counter,
                                const int32_t *vol, int32_t *bad, cudaStream_t st)
{
    const long long n = (long long)p.N * p.A;
    BPP_LAUNCH((bpp_import_kernel), (unsigned)((n + 255) / 256), 256, 0, st, p, plain, box, counter, vol, bad);
    return cudaGetLastError();
}

__global__ void bpp_debug_draws_kernel(unsigned long long seed, uint32_t domain, uint32_t aux, unsigned long long gid,
                                       int n, uint32_t *out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = draw((uint32_t)seed, (uint32_t)(seed >> 32), domain, aux, gid, (uint32_t)t);
}

cudaError_t launch_debug_draws(unsigned long long seed, uint32_t domain, uint32_t aux, unsigned long long gid, int n,
                               uint32_t *out, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    BPP_LAUNCH((bpp_debug_draws_kernel), (n + 127) / 128, 128, 0, st, seed, domain, aux, gid, n, out);
    return cudaGetLastError();
}

#ifndef BPP_SIMT_EMUL
// cuRAND's own Philox4_32_10 stream, for the test that the hand-written generator above reproduces it:
// out[i] = i-th curand() after curand_init(seed, subsequence, 0).
__global__ void bpp_debug_curand_kernel(unsigned long long seed, unsigned long long subsequence, int n, uint32_t *out)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        curandStatePhilox4_32_10_t st;
        curand_init(seed, subsequence, 0ull, &st);
        for (int i = 0; i < n; ++i) out[i] = curand(&st);
    }
}

cudaError_t launch_debug_curand(unsigned long long seed, unsigned long long subsequence, int n, uint32_t *out,
                                cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    BPP_LAUNCH((bpp_debug_curand_kernel), 1, 32, 0, st, seed, subsequence, n, out);
    return cudaGetLastError();
}
#endif

}  // namespace bpp
