// bpp_params.h -- kernel parameters derived from a bpp_config (include/bpp_b200.h): geometry, padded row lengths, RNG
// key, replay offsets, cut-generator ranges and capacities, shared-memory budget of the step kernel.  Pure host
// arithmetic shared by bpp_create (bpp_api.cu) and by the CPU interpretation of the kernels in tests/simt/.
#pragma once

#include "../../include/bpp_b200.h"
#include "bpp_launch.h"

namespace bpp {

// Fills every field of `p` that does not point to device memory.  `cfg` must have passed bpp_create's validation.
inline void derive_params(const bpp_config &cfg, EnvParams &p)
{
    const int rot = cfg.rot ? 1 : 0;
    p.N = cfg.n_bins; p.L = cfg.L; p.W = cfg.W; p.H = cfg.H;
    p.A = cfg.L * cfg.W; p.A_pad = (p.A + 15) & ~15;
    p.M = p.A * (rot ? 2 : 1); p.M_pad = (p.M + 15) & ~15;
    p.gen_kind = cfg.gen_kind;
    p.seed_lo = (uint32_t)cfg.seed; p.seed_hi = (uint32_t)(cfg.seed >> 32);
    p.global_offset = (unsigned long long)cfg.global_offset;
    p.lwh_f = (float)(cfg.L * cfg.W * cfg.H);
    const long long global_n = cfg.global_n > 0 ? cfg.global_n : cfg.n_bins;
    switch (cfg.gen_kind) {
    case BPP_GEN_RS:
        p.n_items = cfg.n_items;
        break;
    case BPP_GEN_REPLAY: {
        p.n_traj = cfg.n_traj; p.traj_cap = cfg.traj_cap;
        const long long T = cfg.n_traj;
        p.first_mod = (uint32_t)(((cfg.first_traj % T) + T) % T);
        p.gn_mod = (uint32_t)(global_n % T);
        break;
    }
    case BPP_GEN_CUT1:
    case BPP_GEN_CUT2: {
        const int axes = cfg.gen_kind == BPP_GEN_CUT1 ? 3 : 1;     // cut2 uses one range for all axes (bin3D.py:33)
        for (int a = 0; a < 3; ++a) { p.lo[a] = cfg.lo[axes == 3 ? a : 0]; p.hi[a] = cfg.hi[axes == 3 ? a : 0]; }
        p.rot = rot;
        p.ep_cap = (int)((long long)(cfg.L / p.lo[0]) * (cfg.W / p.lo[1]) * (cfg.H / p.lo[2]));
        p.ep_buf_bins = cfg.n_bins;
        p.ep_bin0 = 0;
        p.gen_scratch_stride = (long long)gen_scratch_bytes_per_bin(cfg.gen_kind, p.ep_cap);
        break;
    }
    default:
        break;
    }
}

// Dynamic shared memory of one step-kernel CTA.  Ordinary path: per warp the heightmap tile, the mask tile, the packed
// sweep's pair array (compile-time bin sides only) and 32 staged bin records.  Bulk-copy path (tma): the table of
// constant observation planes, then per warp the batch's tiles, mask tile, pairs, float32 staging rows and records.
inline size_t step_smem_bytes(const EnvParams &p, int tma = 0, int rot = 0, int warps = kWarpsPerBlock)   // tma: log2 of the bins per warp batch, 0 = ordinary path
{
    const int cs = (p.L == p.W && (p.L == 10 || p.L == 20)) ? p.L : 0;
    if (tma) {
        const int bins = 1 << tma, A = cs * cs, M = A * (rot ? 2 : 1);
        const size_t per_warp = (size_t)bins * ((A + 15) & ~15) + ((M + 15) & ~15) + pair_bytes(cs) + 4 * (size_t)(A + M) + (size_t)bins * 80;
        return (size_t)tma_lut_planes(cs) * 4 * A + (size_t)warps * per_warp;
    }
    return (size_t)warps * (size_t)(p.A_pad + p.M_pad + pair_bytes(cs)) + (size_t)warps * 32 * 80;
}

}  // namespace bpp
