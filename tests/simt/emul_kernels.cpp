// emul_kernels.cpp -- the repo's CUDA kernel sources compiled by g++ against simt_emul.h (TESTS ONLY, see there).
//
// Mirrors what bpp_api.cu does around the kernels (derive_params, launch_step / launch_gen / launch_mask_*), with
// "device" buffers being plain host arrays owned by the Python test.  The kernels, their template dispatch and the
// per-launch batch-size rule are the product sources, included verbatim.
#define BPP_SIMT_EMUL 1
#include "../../online-3d-bpp-drl_b200/csrc/bpp_step.cu"
#include "../../online-3d-bpp-drl_b200/csrc/bpp_gen.cu"
#include "../../online-3d-bpp-drl_b200/csrc/bpp_mask.cu"
#include "../../online-3d-bpp-drl_b200/csrc/bpp_params.h"

namespace simt {
Machine M;
int g_blocks_per_sm = 4;
int g_lane_order = 0;
}

using namespace bpp;

extern "C" {

// "Device" memory of one emulated env; the Python side allocates everything (sizes from emul_sizes).
struct EmulBuffers {
    uint8_t *hm;                 // [N][A_pad]
    void *sc;                    // [N] BinScalars (32 B)
    const uint32_t *item_set;    // packed x | y<<8 | z<<16
    const uint32_t *traj;        // [n_traj][traj_cap] packed
    const int32_t *traj_len;
    uint64_t *metrics;           // [4]
    uint32_t *ep_items;          // cut1/cut2: [2][N][ep_cap]
    int32_t *ep_len2;            // [2][N]
    int32_t *gen_list;           // [N]
    int32_t *gen_count;          // [1]
    uint8_t *gen_scratch;        // [N][gen_scratch_stride]
};

struct EmulIO {
    const int64_t *actions;
    float *obs, *mask, *reward;
    uint8_t *done;
    int32_t *counter;
    float *ratio;
    int64_t *sample;
    void *packed;                // bpp_result[N] (16 B) or null
    int32_t n_steps;             // env steps per launch (0 / 1 = one)
    const int32_t *actions32;    // int32 actions with a stride, instead of `actions`
    int32_t actions32_stride;
};

static EnvParams make_params(const bpp_config *cfg, const EmulBuffers *b)
{
    EnvParams p;
    std::memset(&p, 0, sizeof(p));
    derive_params(*cfg, p);
    p.hm = b->hm; p.sc = static_cast<BinScalars *>(b->sc); p.metrics = reinterpret_cast<unsigned long long *>(b->metrics);
    p.item_set = b->item_set; p.traj = b->traj; p.traj_len = b->traj_len;
    p.ep_items = b->ep_items; p.ep_len2 = b->ep_len2; p.gen_list = b->gen_list; p.gen_count = b->gen_count;
    p.gen_scratch = b->gen_scratch;
    return p;
}

// out[0] = A_pad, [1] = M_pad, [2] = ep_cap, [3] = gen_scratch_stride, [4] = sizeof(BinScalars), [5] = step smem bytes
void emul_sizes(const bpp_config *cfg, int64_t *out)
{
    EnvParams p;
    std::memset(&p, 0, sizeof(p));
    derive_params(*cfg, p);
    out[0] = p.A_pad; out[1] = p.M_pad; out[2] = p.ep_cap; out[3] = p.gen_scratch_stride;
    out[4] = (int64_t)sizeof(BinScalars); out[5] = (int64_t)step_smem_bytes(p);
}

static inline bool aligned16(const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

// bpp_reset / bpp_step / bpp_observe of bpp_api.cu: mode 1 / 0 / 2 (reset takes was_reset for LoadBoxCreator.index += 1).
int emul_env_call(const bpp_config *cfg, const EmulBuffers *b, const EmulIO *io_in, int mode, int was_reset,
                  int num_sms, int blocks_per_sm)
{
    simt::g_blocks_per_sm = blocks_per_sm;
    EnvParams p = make_params(cfg, b);
    StepIO io{};
    io.actions = reinterpret_cast<const long long *>(io_in->actions);
    io.obs = io_in->obs; io.mask = io_in->mask; io.reward = io_in->reward; io.done = io_in->done;
    io.counter = io_in->counter; io.ratio = io_in->ratio; io.sample = reinterpret_cast<long long *>(io_in->sample);
    io.packed = static_cast<uint4 *>(io_in->packed);
    io.n_steps = io_in->n_steps > 1 ? io_in->n_steps : 1;
    io.actions32 = io_in->actions32; io.actions32_stride = io_in->actions32_stride;
    io.episode_inc = (mode == kModeReset && was_reset) ? 1 : 0;
    const bool cut = p.gen_kind >= 2;
    const int rot = cfg->rot ? 1 : 0;
    const bool vec4 = (p.A % 4 == 0) && (!io.obs || aligned16(io.obs)) && (!io.mask || aligned16(io.mask));
    if (mode == kModeReset && cut && !was_reset)
        if (launch_gen(p, 0, 0u, num_sms, nullptr) != cudaSuccess) return -2;
    if (launch_step(p, io, mode, rot, cfg->mask_variant, vec4, num_sms, nullptr) != cudaSuccess) return -2;
    if (cut && ((mode == kModeReset && was_reset) || mode == kModeStep))
        if (launch_gen(p, 1, 0u, num_sms, nullptr) != cudaSuccess) return -2;
    return 0;
}

int emul_mask_from_obs(const float *obs, int n, int L, int W, int H, int rot, int variant, float *mask, int grid)
{
    return launch_mask_from_obs(obs, n, L, W, H, rot, variant, mask, grid, nullptr);
}

int emul_mask_from_plain(const int32_t *plain, const int32_t *box, int n, int L, int W, int H, int rot, int variant,
                         float *mask, int grid)
{
    return launch_mask_from_plain(plain, box, n, L, W, H, rot, variant, mask, grid, nullptr);
}

int emul_check_box(const int32_t *plain, const int32_t *query, int n, int L, int W, int H, int variant, int32_t *out)
{
    return launch_check_box(plain, query, n, L, W, H, variant, out, nullptr);
}

int emul_export_state(const bpp_config *cfg, const EmulBuffers *b, int32_t *plain, int32_t *box, int32_t *counter, int32_t *vol)
{
    return launch_export_state(make_params(cfg, b), plain, box, counter, vol, nullptr);
}

int emul_import_state(const bpp_config *cfg, const EmulBuffers *b, const int32_t *plain, const int32_t *box,
                      const int32_t *counter, const int32_t *vol)
{
    return launch_import_state(make_params(cfg, b), plain, box, counter, vol, nullptr, nullptr);
}

int emul_debug_draws(unsigned long long seed, uint32_t domain, uint32_t aux, unsigned long long gid, int n, uint32_t *out)
{
    return launch_debug_draws(seed, domain, aux, gid, n, out, nullptr);
}

unsigned long long emul_collectives(void) { return simt::M.collectives; }
void emul_set_lane_order(int order) { simt::g_lane_order = order; }

}  // extern "C"
