// bpp_api.cu -- the C ABI of libbpp_b200.so (include/bpp_b200.h).  Host-side plumbing only: argument
// validation, device memory owned by the env, launches.  No CPU fallback exists: without a usable
// sm_100 device every entry point fails with BPP_ERR_NO_DEVICE / BPP_ERR_CUDA.
#include <atomic>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <sched.h>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/bpp_b200.h"
#include "bpp_launch.h"
#include "bpp_params.h"

using namespace bpp;

namespace {

thread_local std::string g_last_error;
std::atomic<unsigned long long> g_launches{0};

int fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}

#define BPP_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return fail(BPP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
    } while (0)

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }
inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int device_grid_cap(int device, int *num_sms)
{
    cudaDeviceProp prop;
    BPP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(BPP_ERR_NO_DEVICE, "libbpp_b200 is built for sm_100a only; device " + std::to_string(device) +
                                           " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor));
    *num_sms = prop.multiProcessorCount;
    return BPP_OK;
}

// Device-side alias of a pinned (page-locked, mapped) host buffer -- or the pointer itself when it already is device
// (or managed) memory --, nullptr for pageable host memory.
inline void *mapped_alias(void *ptr)
{
    if (!ptr) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return ptr;
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    return at.devicePointer;
}

}  // namespace

struct bpp_env {
    bpp_config cfg;
    EnvParams p;
    int rot, variant;
    int num_sms;
    bool was_reset;
    unsigned long long steps;
    // device memory owned by the env
    uint8_t *d_hm;
    BinScalars *d_sc;
    uint32_t *d_items, *d_traj, *d_ep_items;
    int32_t *d_ep_len2;
    int32_t *d_gen_list, *d_gen_count;
    uint8_t *d_gen_scratch;
    int32_t *d_traj_len;
    unsigned long long *d_metrics;
    // staging for the host-buffer entry point
    long long *d_actions;
    float *d_reward, *d_ratio;
    uint8_t *d_done;
    int32_t *d_counter;
    // worker streams / events for the chunk-pipelined host entry point
    cudaStream_t ws[2];
    cudaEvent_t ev_start, ev_done[2];
    // last host buffers seen by the host entry points and their device aliases (pointer attribute queries are slow);
    // re-validated every kZcRecheck calls: a caller may have freed a pinned buffer and passed ordinary memory at
    // the same address
    const void *zc_host[5];
    void *zc_dev[5];
    bool zc_valid, zc_ok;
    unsigned zc_age;
    bpp_result *d_results;      // staging of the packed host entry point (pageable buffers)
    int32_t *d_flag;            // import-state validation
};

extern "C" {

int32_t bpp_abi_version(void) { return BPP_ABI_VERSION; }
const char *bpp_last_error(void) { return g_last_error.c_str(); }
uint64_t bpp_launch_count(void) { return g_launches.load(); }

#define BPP_STR2(x) #x
#define BPP_STR(x) BPP_STR2(x)
const char *bpp_build_info(void)
{
    return "packed_sweep=" BPP_STR(BPP_PACKED_SWEEP) " lane_check=" BPP_STR(BPP_LANE_CHECK) " square_copy=" BPP_STR(BPP_SQUARE_COPY)
           " plain_stores=" BPP_STR(BPP_PLAIN_STORES) " tma=" BPP_STR(BPP_TMA)
           " tma_bpw_log2=" BPP_STR(BPP_TMA_BPW_LOG2) "/" BPP_STR(BPP_TMA_BPW_LOG2_ROT) " tma20_bpw_log2=" BPP_STR(BPP_TMA20)
           " warps_per_block=" BPP_STR(BPP_WARPS_PER_BLOCK) " (bulk-copy path 10x10: 8 / 7 / 6 by rotation and env count) sm_100a";
}

int32_t bpp_destroy(bpp_env *env)
{
    if (!env) return BPP_OK;
    cudaSetDevice(env->cfg.device);
    cudaFree(env->d_hm); cudaFree(env->d_sc); cudaFree(env->d_items); cudaFree(env->d_traj);
    cudaFree(env->d_ep_items); cudaFree(env->d_ep_len2); cudaFree(env->d_gen_list); cudaFree(env->d_gen_count); cudaFree(env->d_gen_scratch);
    cudaFree(env->d_traj_len); cudaFree(env->d_metrics); cudaFree(env->d_actions);
    cudaFree(env->d_reward); cudaFree(env->d_ratio); cudaFree(env->d_done); cudaFree(env->d_counter);
    cudaFree(env->d_results); cudaFree(env->d_flag);
    for (int k = 0; k < 2; ++k) {
        if (env->ws[k]) cudaStreamDestroy(env->ws[k]);
        if (env->ev_done[k]) cudaEventDestroy(env->ev_done[k]);
    }
    if (env->ev_start) cudaEventDestroy(env->ev_start);
    delete env;
    return BPP_OK;
}

int32_t bpp_create(const bpp_config *cfg, bpp_env **out)
{
    if (!cfg || !out) return fail(BPP_ERR_INVALID, "bpp_create: null argument");
    *out = nullptr;
    if (cfg->abi_version != BPP_ABI_VERSION) return fail(BPP_ERR_INVALID, "bpp_create: ABI version mismatch");
    if (cfg->n_bins <= 0) return fail(BPP_ERR_INVALID, "bpp_create: n_bins must be positive");
    if (cfg->L < 1 || cfg->W < 1 || cfg->L > kMaxSide || cfg->W > kMaxSide)
        return fail(BPP_ERR_INVALID, "bpp_create: L and W must be in [1, 32]");
    if (cfg->H < 1 || cfg->H > 255) return fail(BPP_ERR_INVALID, "bpp_create: H must be in [1, 255]");
    if (cfg->mask_variant != BPP_MASK_V1 && cfg->mask_variant != BPP_MASK_V2)
        return fail(BPP_ERR_INVALID, "bpp_create: mask_variant must be 1 or 2");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0)
        return fail(BPP_ERR_NO_DEVICE, "bpp_create: no CUDA device (this library has no CPU path)");
    if (cfg->device < 0 || cfg->device >= count) return fail(BPP_ERR_INVALID, "bpp_create: bad device ordinal");
    int num_sms = 0;
    if (int rc = device_grid_cap(cfg->device, &num_sms)) return rc;
    BPP_CUDA(cudaSetDevice(cfg->device));

    bpp_env *e = new (std::nothrow) bpp_env();
    if (!e) return fail(BPP_ERR_INVALID, "bpp_create: out of host memory");
    std::memset(static_cast<void *>(e), 0, sizeof(*e));
    e->cfg = *cfg;
    e->rot = cfg->rot ? 1 : 0;
    e->variant = cfg->mask_variant;
    e->num_sms = num_sms;
    EnvParams &p = e->p;
    p.gen_kind = cfg->gen_kind;      // derive_params() fills the rest once the generator arguments are validated

#define BPP_CREATE_CUDA(expr)                                                                       \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            std::string m = std::string(#expr) + ": " + cudaGetErrorString(_e);                     \
            bpp_destroy(e);                                                                         \
            return fail(BPP_ERR_CUDA, m);                                                           \
        }                                                                                           \
    } while (0)

    switch (cfg->gen_kind) {
    case BPP_GEN_RS: {
        if (cfg->n_items <= 0 || !cfg->item_set_host) { bpp_destroy(e); return fail(BPP_ERR_INVALID, "bpp_create: rs needs an item set"); }
        std::vector<uint32_t> packed((size_t)cfg->n_items);
        for (int i = 0; i < cfg->n_items; ++i) {
            const uint8_t *it = cfg->item_set_host + (size_t)i * 3;
            if (!it[0] || !it[1] || !it[2]) { bpp_destroy(e); return fail(BPP_ERR_INVALID, "bpp_create: item dims must be >= 1"); }
            packed[i] = it[0] | (it[1] << 8) | (it[2] << 16);
        }
        BPP_CREATE_CUDA(cudaMalloc(&e->d_items, packed.size() * 4));
        BPP_CREATE_CUDA(cudaMemcpy(e->d_items, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice));
        p.item_set = e->d_items;
        break;
    }
    case BPP_GEN_REPLAY: {
        if (cfg->n_traj <= 0 || cfg->traj_cap <= 0 || !cfg->traj_host || !cfg->traj_len_host) {
            bpp_destroy(e); return fail(BPP_ERR_INVALID, "bpp_create: replay needs trajectories");
        }
        const size_t rows = (size_t)cfg->n_traj, cap = (size_t)cfg->traj_cap;
        std::vector<uint32_t> packed(rows * cap, 0u);
        for (size_t t = 0; t < rows; ++t) {
            if (cfg->traj_len_host[t] < 0 || cfg->traj_len_host[t] > cfg->traj_cap) {
                bpp_destroy(e); return fail(BPP_ERR_INVALID, "bpp_create: traj_len out of range");
            }
            for (int k = 0; k < cfg->traj_len_host[t]; ++k) {
                const uint8_t *it = cfg->traj_host + (t * cap + k) * 3;
                if (!it[0] || !it[1] || !it[2]) { bpp_destroy(e); return fail(BPP_ERR_INVALID, "bpp_create: item dims must be >= 1"); }
                packed[t * cap + k] = it[0] | (it[1] << 8) | (it[2] << 16);
            }
        }
        BPP_CREATE_CUDA(cudaMalloc(&e->d_traj, packed.size() * 4));
        BPP_CREATE_CUDA(cudaMemcpy(e->d_traj, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice));
        BPP_CREATE_CUDA(cudaMalloc(&e->d_traj_len, rows * 4));
        BPP_CREATE_CUDA(cudaMemcpy(e->d_traj_len, cfg->traj_len_host, rows * 4, cudaMemcpyHostToDevice));
        p.traj = e->d_traj; p.traj_len = e->d_traj_len;
        break;
    }
    case BPP_GEN_CUT1:
    case BPP_GEN_CUT2: {
        // cutCreator.py:69-76 asserts lo <= dim - lo for every cut (dim > hi), mdCreator.py:73-76 needs a cut that
        // leaves both parts >= lo: both require hi + 1 >= 2 * lo.
        const int axes = cfg->gen_kind == BPP_GEN_CUT1 ? 3 : 1;
        for (int a = 0; a < axes; ++a) {
            if (cfg->lo[a] < 1 || cfg->hi[a] < cfg->lo[a] || cfg->hi[a] > 255 || cfg->hi[a] + 1 < 2 * cfg->lo[a]) {
                bpp_destroy(e);
                return fail(BPP_ERR_INVALID, "bpp_create: cut generators need 1 <= lo <= hi and hi + 1 >= 2 * lo per axis");
            }
        }
        const long long cap = (long long)(cfg->L / cfg->lo[0]) * (cfg->W / cfg->lo[axes == 3 ? 1 : 0]) * (cfg->H / cfg->lo[axes == 3 ? 2 : 0]);
        if (cap < 1 || cap > 60000) { bpp_destroy(e); return fail(BPP_ERR_INVALID, "bpp_create: episode capacity out of range"); }
        derive_params(*cfg, p);
        const size_t Nb = (size_t)cfg->n_bins;
        BPP_CREATE_CUDA(cudaMalloc(&e->d_ep_items, 2 * Nb * p.ep_cap * sizeof(uint32_t)));
        BPP_CREATE_CUDA(cudaMalloc(&e->d_ep_len2, 2 * Nb * sizeof(int32_t)));
        BPP_CREATE_CUDA(cudaMemset(e->d_ep_len2, 0, 2 * Nb * sizeof(int32_t)));
        BPP_CREATE_CUDA(cudaMalloc(&e->d_gen_list, Nb * sizeof(int32_t)));
        BPP_CREATE_CUDA(cudaMalloc(&e->d_gen_count, sizeof(int32_t)));
        BPP_CREATE_CUDA(cudaMemset(e->d_gen_count, 0, sizeof(int32_t)));
        BPP_CREATE_CUDA(cudaMalloc(&e->d_gen_scratch, Nb * (size_t)p.gen_scratch_stride));
        p.ep_items = e->d_ep_items; p.ep_len2 = e->d_ep_len2; p.gen_list = e->d_gen_list; p.gen_count = e->d_gen_count; p.gen_scratch = e->d_gen_scratch;
        break;
    }
    default:
        bpp_destroy(e);
        return fail(BPP_ERR_INVALID, "bpp_create: unknown generator kind");
    }

    derive_params(*cfg, p);
    const size_t N = (size_t)p.N;
    BPP_CREATE_CUDA(cudaMalloc(&e->d_hm, N * p.A_pad + kHmSlack));
    BPP_CREATE_CUDA(cudaMemset(e->d_hm, 0, N * p.A_pad + kHmSlack));
    BPP_CREATE_CUDA(cudaMalloc(&e->d_sc, N * sizeof(BinScalars)));
    BPP_CREATE_CUDA(cudaMemset(e->d_sc, 0, N * sizeof(BinScalars)));
    BPP_CREATE_CUDA(cudaMalloc(&e->d_metrics, 4 * sizeof(unsigned long long)));
    BPP_CREATE_CUDA(cudaMemset(e->d_metrics, 0, 4 * sizeof(unsigned long long)));
    BPP_CREATE_CUDA(cudaMalloc(&e->d_actions, N * sizeof(long long)));
    BPP_CREATE_CUDA(cudaMalloc(&e->d_reward, N * sizeof(float)));
    BPP_CREATE_CUDA(cudaMalloc(&e->d_ratio, N * sizeof(float)));
    BPP_CREATE_CUDA(cudaMalloc(&e->d_done, N));
    BPP_CREATE_CUDA(cudaMalloc(&e->d_counter, N * sizeof(int32_t)));
    BPP_CREATE_CUDA(cudaMalloc(&e->d_flag, sizeof(int32_t)));
    BPP_CREATE_CUDA(cudaMemset(e->d_flag, 0, sizeof(int32_t)));
    p.hm = e->d_hm; p.sc = e->d_sc; p.metrics = e->d_metrics;

    for (int k = 0; k < 2; ++k) {
        BPP_CREATE_CUDA(cudaStreamCreateWithFlags(&e->ws[k], cudaStreamNonBlocking));
        BPP_CREATE_CUDA(cudaEventCreateWithFlags(&e->ev_done[k], cudaEventDisableTiming));
    }
    BPP_CREATE_CUDA(cudaEventCreateWithFlags(&e->ev_start, cudaEventDisableTiming));
    *out = e;
    return BPP_OK;
}

// Launches the step kernel over the bins [b0, b0 + n) of the env (the whole env when b0 = 0, n = N).
static int run_step_range(bpp_env *env, const StepIO &io_all, int mode, long long b0, int n, cudaStream_t st)
{
    EnvParams p = env->p;
    StepIO io = io_all;
    p.N = n;
    p.hm += (size_t)b0 * p.A_pad;
    p.sc += b0;
    p.global_offset += (unsigned long long)b0;
    p.ep_bin0 = b0;
    if (io.actions) io.actions += b0;
    if (io.actions32) io.actions32 += (size_t)b0 * (size_t)io.actions32_stride;
    if (io.obs) io.obs += (size_t)b0 * 4 * p.A;
    if (io.mask) io.mask += (size_t)b0 * p.M;
    if (io.reward) io.reward += b0;
    if (io.done) io.done += b0;
    if (io.counter) io.counter += b0;
    if (io.ratio) io.ratio += b0;
    if (io.sample) io.sample += b0;
    if (io.packed) io.packed += b0;
    const bool vec4 = (p.A % 4 == 0) && (!io.obs || aligned16(io.obs)) && (!io.mask || aligned16(io.mask));
    BPP_CUDA(launch_step(p, io, mode, env->rot, env->variant, vec4, env->num_sms, st));
    g_launches.fetch_add(1);
    return BPP_OK;
}

static int run_step(bpp_env *env, const StepIO &io, int mode, void *stream)
{
    return run_step_range(env, io, mode, 0, env->p.N, as_stream(stream));
}

static int enter(bpp_env *env, const char *who)
{
    if (!env) return fail(BPP_ERR_INVALID, std::string(who) + ": null env");
    BPP_CUDA(cudaSetDevice(env->cfg.device));
    return BPP_OK;
}

int32_t bpp_reset(bpp_env *env, float *obs, float *mask, int64_t *sample, void *stream)
{
    if (int rc = enter(env, "bpp_reset")) return rc;
    StepIO io{};
    io.obs = obs; io.mask = mask; io.sample = reinterpret_cast<long long *>(sample);
    io.episode_inc = env->was_reset ? 1 : 0;
    const bool cut = env->p.gen_kind >= 2;
    if (cut && !env->was_reset) {                         // episodes 0 and 1 of every bin
        BPP_CUDA(launch_gen(env->p, 0, 0u, env->num_sms, as_stream(stream)));
        g_launches.fetch_add(1);
    }
    if (int rc = run_step(env, io, kModeReset, stream)) return rc;
    if (cut && env->was_reset) {                          // every bin moved to its spare episode: refill
        BPP_CUDA(launch_gen(env->p, 1, 0u, env->num_sms, as_stream(stream)));
        g_launches.fetch_add(1);
    }
    env->was_reset = true;
    return BPP_OK;
}

int32_t bpp_step(bpp_env *env, const int64_t *actions, float *obs, float *mask, float *reward, uint8_t *done,
                 int32_t *counter, float *ratio, int64_t *sample, void *stream)
{
    if (int rc = enter(env, "bpp_step")) return rc;
    if (!env->was_reset) return fail(BPP_ERR_STATE, "bpp_step: call bpp_reset first");
    if (!actions) return fail(BPP_ERR_INVALID, "bpp_step: actions is null");
    StepIO io{};
    io.actions = reinterpret_cast<const long long *>(actions);
    io.obs = obs; io.mask = mask; io.reward = reward; io.done = done; io.counter = counter; io.ratio = ratio;
    io.sample = reinterpret_cast<long long *>(sample);
    if (int rc = run_step(env, io, kModeStep, stream)) return rc;
    if (env->p.gen_kind >= 2) {                           // refill the episode buffers of bins that just reset
        BPP_CUDA(launch_gen(env->p, 1, 0u, env->num_sms, as_stream(stream)));
        g_launches.fetch_add(1);
    }
    env->steps += (unsigned long long)env->p.N;
    return BPP_OK;
}


// Device aliases of up to five host buffers (null entries allowed): true when every non-null one is pinned and mapped.
// The answer is cached per env, keyed on the pointer values, and re-validated every kZcRecheck calls.
constexpr unsigned kZcRecheck = 256;
static bool zero_copy_aliases(bpp_env *env, const void *const hosts[5])
{
    if (!env->zc_valid || std::memcmp(hosts, env->zc_host, 5 * sizeof(void *)) != 0 || ++env->zc_age >= kZcRecheck) {
        std::memcpy(env->zc_host, hosts, 5 * sizeof(void *));
        env->zc_ok = true;
        for (int k = 0; k < 5; ++k) {
            env->zc_dev[k] = hosts[k] ? mapped_alias(const_cast<void *>(hosts[k])) : nullptr;
            if (hosts[k] && !env->zc_dev[k]) env->zc_ok = false;
        }
        env->zc_valid = true;
        env->zc_age = 0;
    }
    return env->zc_ok;
}

int32_t bpp_step_host(bpp_env *env, const int64_t *actions_host, float *obs, float *mask, float *reward_host,
                      uint8_t *done_host, int32_t *counter_host, float *ratio_host, int64_t *sample, void *stream)
{
    if (int rc = enter(env, "bpp_step_host")) return rc;
    if (!env->was_reset) return fail(BPP_ERR_STATE, "bpp_step_host: call bpp_reset first");
    if (!actions_host) return fail(BPP_ERR_INVALID, "bpp_step_host: actions_host is null");
    cudaStream_t st = as_stream(stream);
    const long long N = env->p.N;
    StepIO io{};
    io.actions = env->d_actions;
    io.obs = obs; io.mask = mask;
    io.reward = reward_host ? env->d_reward : nullptr;
    io.done = done_host ? env->d_done : nullptr;
    io.counter = counter_host ? env->d_counter : nullptr;
    io.ratio = ratio_host ? env->d_ratio : nullptr;
    io.sample = reinterpret_cast<long long *>(sample);
    // Zero-copy path: when every host buffer is pinned and mapped, the kernel reads the actions and writes
    // reward / done / counter / ratio straight over PCIe (coalesced 128-byte transactions issued by the lane-per-bin
    // phases), so the transfers overlap the tile work of other warps and no staging copy is needed.
    {
        const void *hosts[5] = {actions_host, reward_host, ratio_host, done_host, counter_host};
        if (zero_copy_aliases(env, hosts)) {
            StepIO z = io;
            z.actions = static_cast<const long long *>(env->zc_dev[0]);
            z.reward = static_cast<float *>(env->zc_dev[1]);
            z.ratio = static_cast<float *>(env->zc_dev[2]);
            z.done = static_cast<uint8_t *>(env->zc_dev[3]);
            z.counter = static_cast<int32_t *>(env->zc_dev[4]);
            if (int rc = run_step(env, z, kModeStep, stream)) return rc;
            if (env->p.gen_kind >= 2) {
                BPP_CUDA(launch_gen(env->p, 1, 0u, env->num_sms, st));
                g_launches.fetch_add(1);
            }
            env->steps += (unsigned long long)N;
            BPP_CUDA(cudaStreamSynchronize(st));
            return BPP_OK;
        }
    }
    // Staged path (pageable host memory): bins are independent, so the env is stepped in up to 4 slices on two worker
    // streams; slice c+1's action upload and slice c-1's result download overlap slice c's kernel.
    const int chunks = N >= 131072 ? 4 : (N >= 32768 ? 2 : 1);
    BPP_CUDA(cudaEventRecord(env->ev_start, st));
    for (int k = 0; k < 2; ++k) BPP_CUDA(cudaStreamWaitEvent(env->ws[k], env->ev_start, 0));
    for (int c = 0; c < chunks; ++c) {
        const long long b0 = ((N * c / chunks) / 32) * 32;
        const long long b1 = c + 1 == chunks ? N : ((N * (c + 1) / chunks) / 32) * 32;
        const size_t n = (size_t)(b1 - b0);
        if (n == 0) continue;
        cudaStream_t w = env->ws[c & 1];
        BPP_CUDA(cudaMemcpyAsync(env->d_actions + b0, actions_host + b0, n * sizeof(long long), cudaMemcpyHostToDevice, w));
        if (int rc = run_step_range(env, io, kModeStep, b0, (int)n, w)) return rc;
        if (reward_host) BPP_CUDA(cudaMemcpyAsync(reward_host + b0, env->d_reward + b0, n * sizeof(float), cudaMemcpyDeviceToHost, w));
        if (done_host) BPP_CUDA(cudaMemcpyAsync(done_host + b0, env->d_done + b0, n, cudaMemcpyDeviceToHost, w));
        if (counter_host) BPP_CUDA(cudaMemcpyAsync(counter_host + b0, env->d_counter + b0, n * sizeof(int32_t), cudaMemcpyDeviceToHost, w));
        if (ratio_host) BPP_CUDA(cudaMemcpyAsync(ratio_host + b0, env->d_ratio + b0, n * sizeof(float), cudaMemcpyDeviceToHost, w));
    }
    for (int k = 0; k < 2; ++k) {
        BPP_CUDA(cudaEventRecord(env->ev_done[k], env->ws[k]));
        BPP_CUDA(cudaStreamWaitEvent(st, env->ev_done[k], 0));
    }
    if (env->p.gen_kind >= 2) {
        BPP_CUDA(launch_gen(env->p, 1, 0u, env->num_sms, st));
        g_launches.fetch_add(1);
    }
    env->steps += (unsigned long long)N;
    BPP_CUDA(cudaStreamSynchronize(st));
    return BPP_OK;
}


int32_t bpp_step_packed(bpp_env *env, const int64_t *actions, float *obs, float *mask, bpp_result *results,
                        uint8_t *done, int64_t *sample, int32_t n_steps, void *stream)
{
    if (int rc = enter(env, "bpp_step_packed")) return rc;
    if (!env->was_reset) return fail(BPP_ERR_STATE, "bpp_step_packed: call bpp_reset first");
    if (!actions || !results || !done || !obs || !mask) return fail(BPP_ERR_INVALID, "bpp_step_packed: null pointer");
    if (!aligned16(results)) return fail(BPP_ERR_INVALID, "bpp_step_packed: results must be 16-byte aligned");
    if (n_steps < 1) return fail(BPP_ERR_INVALID, "bpp_step_packed: n_steps must be >= 1");
    if (n_steps > 1 && env->p.gen_kind >= 2)
        return fail(BPP_ERR_INVALID, "bpp_step_packed: several steps per launch need the rs or replay generator "
                                     "(cut1 / cut2 refill their episode buffers between steps)");
    StepIO io{};
    io.actions = reinterpret_cast<const long long *>(actions);
    io.obs = obs; io.mask = mask; io.done = done;
    io.packed = reinterpret_cast<uint4 *>(results);
    io.sample = reinterpret_cast<long long *>(sample);
    io.n_steps = n_steps;
    if (int rc = run_step(env, io, kModeStep, stream)) return rc;
    if (env->p.gen_kind >= 2) {
        BPP_CUDA(launch_gen(env->p, 1, 0u, env->num_sms, as_stream(stream)));
        g_launches.fetch_add(1);
    }
    env->steps += (unsigned long long)env->p.N * (unsigned long long)n_steps;
    return BPP_OK;
}

int32_t bpp_step_packed_host(bpp_env *env, const int64_t *actions_host, float *obs, float *mask,
                             bpp_result *results_host, uint8_t *done_host, int64_t *sample, void *stream)
{
    if (int rc = enter(env, "bpp_step_packed_host")) return rc;
    if (!env->was_reset) return fail(BPP_ERR_STATE, "bpp_step_packed_host: call bpp_reset first");
    if (!actions_host || !results_host || !done_host || !obs || !mask)
        return fail(BPP_ERR_INVALID, "bpp_step_packed_host: null pointer");
    cudaStream_t st = as_stream(stream);
    const size_t N = (size_t)env->p.N;
    StepIO io{};
    io.obs = obs; io.mask = mask; io.n_steps = 1;
    io.sample = reinterpret_cast<long long *>(sample);
    const void *hosts[5] = {actions_host, results_host, nullptr, done_host, nullptr};
    const bool zc = zero_copy_aliases(env, hosts) && aligned16(env->zc_dev[1]);
    if (zc) {       // the kernel reads the actions and writes one 16-byte record + one done byte per bin over PCIe
        io.actions = static_cast<const long long *>(env->zc_dev[0]);
        io.packed = static_cast<uint4 *>(env->zc_dev[1]);
        io.done = static_cast<uint8_t *>(env->zc_dev[3]);
    } else {        // ordinary host memory: staged through device buffers
        if (!env->d_results) BPP_CUDA(cudaMalloc(&env->d_results, N * sizeof(bpp_result)));
        BPP_CUDA(cudaMemcpyAsync(env->d_actions, actions_host, N * sizeof(long long), cudaMemcpyDefault, st));   // host or device actions
        io.actions = env->d_actions;
        io.packed = reinterpret_cast<uint4 *>(env->d_results);
        io.done = env->d_done;
    }
    if (int rc = run_step(env, io, kModeStep, stream)) return rc;
    if (env->p.gen_kind >= 2) {
        BPP_CUDA(launch_gen(env->p, 1, 0u, env->num_sms, st));
        g_launches.fetch_add(1);
    }
    if (!zc) {
        BPP_CUDA(cudaMemcpyAsync(results_host, env->d_results, N * sizeof(bpp_result), cudaMemcpyDeviceToHost, st));
        BPP_CUDA(cudaMemcpyAsync(done_host, env->d_done, N, cudaMemcpyDeviceToHost, st));
    }
    env->steps += (unsigned long long)N;
    BPP_CUDA(cudaStreamSynchronize(st));
    return BPP_OK;
}


int32_t bpp_step_packed_host_i32(bpp_env *env, const int32_t *actions_host, int32_t stride, float *obs, float *mask,
                                 bpp_result *results_host, uint8_t *done_host, int64_t *sample, void *stream)
{
    if (int rc = enter(env, "bpp_step_packed_host_i32")) return rc;
    if (!env->was_reset) return fail(BPP_ERR_STATE, "bpp_step_packed_host_i32: call bpp_reset first");
    if (!actions_host || !results_host || !done_host || !obs || !mask || stride < 1)
        return fail(BPP_ERR_INVALID, "bpp_step_packed_host_i32: bad arguments");
    cudaStream_t st = as_stream(stream);
    StepIO io{};
    io.obs = obs; io.mask = mask; io.n_steps = 1;
    io.sample = reinterpret_cast<long long *>(sample);
    const void *hosts[5] = {actions_host, results_host, nullptr, done_host, nullptr};
    if (!zero_copy_aliases(env, hosts) || !aligned16(env->zc_dev[1]))
        return fail(BPP_ERR_INVALID, "bpp_step_packed_host_i32: actions, results and done must be pinned mapped host memory "
                                     "(bpp_host_alloc) or device memory");
    io.actions32 = static_cast<const int32_t *>(env->zc_dev[0]);
    io.actions32_stride = stride;
    io.actions = reinterpret_cast<const long long *>(io.actions32);     // non-null marker for the launch code; never read
    io.packed = static_cast<uint4 *>(env->zc_dev[1]);
    io.done = static_cast<uint8_t *>(env->zc_dev[3]);
    if (int rc = run_step(env, io, kModeStep, stream)) return rc;
    if (env->p.gen_kind >= 2) {
        BPP_CUDA(launch_gen(env->p, 1, 0u, env->num_sms, st));
        g_launches.fetch_add(1);
    }
    env->steps += (unsigned long long)env->p.N;
    BPP_CUDA(cudaStreamSynchronize(st));
    return BPP_OK;
}

// CPUs of the NUMA node a device hangs off: /sys/bus/pci/devices/<bdf>/local_cpulist ("0-31,64-95")
static bool device_local_cpus(int device, cpu_set_t *set)
{
    char bdf[32] = {0};
    if (cudaDeviceGetPCIBusId(bdf, sizeof(bdf), device) != cudaSuccess) { cudaGetLastError(); return false; }
    for (char *c = bdf; *c; ++c) *c = (char)std::tolower((unsigned char)*c);
    const std::string path = std::string("/sys/bus/pci/devices/") + bdf + "/local_cpulist";
    FILE *f = std::fopen(path.c_str(), "r");
    if (!f) return false;
    char buf[4096] = {0};
    const size_t n = std::fread(buf, 1, sizeof(buf) - 1, f);
    std::fclose(f);
    if (n == 0) return false;
    CPU_ZERO(set);
    int count = 0;
    for (char *q = buf; *q && *q != '\n';) {
        char *e = nullptr;
        const long a = std::strtol(q, &e, 10);
        if (e == q) break;
        long b = a;
        if (*e == '-') { q = e + 1; b = std::strtol(q, &e, 10); }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET((int)c, set); ++count; }
        q = (*e == ',') ? e + 1 : e;
        if (*e != ',') break;
    }
    return count > 0;
}

int32_t bpp_bind_thread_to_device(int32_t device)
{
    cpu_set_t local, allowed, both;
    if (!device_local_cpus(device, &local)) return fail(BPP_ERR_INVALID, "bpp_bind_thread_to_device: no local_cpulist for this device");
    if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return fail(BPP_ERR_INVALID, "bpp_bind_thread_to_device: sched_getaffinity failed");
    CPU_AND(&both, &local, &allowed);
    if (CPU_COUNT(&both) == 0) return fail(BPP_ERR_INVALID, "bpp_bind_thread_to_device: none of the device-local CPUs is allowed here");
    if (sched_setaffinity(0, sizeof(both), &both) != 0) return fail(BPP_ERR_INVALID, "bpp_bind_thread_to_device: sched_setaffinity failed");
    return BPP_OK;
}

int32_t bpp_host_alloc(int32_t device, uint64_t bytes, void **out)
{
    if (!out || bytes == 0) return fail(BPP_ERR_INVALID, "bpp_host_alloc: bad arguments");
    *out = nullptr;
    BPP_CUDA(cudaSetDevice(device));
    // first touch decides the NUMA node of the pages: do it from a CPU next to the device, then restore the affinity
    cpu_set_t saved, local, both;
    const bool have_saved = sched_getaffinity(0, sizeof(saved), &saved) == 0;
    bool bound = false;
    if (have_saved && device_local_cpus(device, &local)) {
        CPU_AND(&both, &local, &saved);
        if (CPU_COUNT(&both) > 0) bound = sched_setaffinity(0, sizeof(both), &both) == 0;
    }
    void *ptr = nullptr;
    const cudaError_t e = cudaHostAlloc(&ptr, (size_t)bytes, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e == cudaSuccess) std::memset(ptr, 0, (size_t)bytes);
    if (bound) sched_setaffinity(0, sizeof(saved), &saved);
    if (e != cudaSuccess) return fail(BPP_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
    *out = ptr;
    return BPP_OK;
}

int32_t bpp_host_free(void *ptr)
{
    if (!ptr) return BPP_OK;
    BPP_CUDA(cudaFreeHost(ptr));
    return BPP_OK;
}

int32_t bpp_observe(bpp_env *env, float *obs, float *mask, void *stream)
{
    if (int rc = enter(env, "bpp_observe")) return rc;
    if (!env->was_reset) return fail(BPP_ERR_STATE, "bpp_observe: call bpp_reset first");
    StepIO io{};
    io.obs = obs; io.mask = mask;
    return run_step(env, io, kModeObserve, stream);
}

static int stateless_grid(int n, int *grid)
{
    int dev = 0, sms = 0;
    BPP_CUDA(cudaGetDevice(&dev));
    if (int rc = device_grid_cap(dev, &sms)) return rc;
    const long long want = ((long long)n + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const long long cap = (long long)sms * 8;
    *grid = (int)(want < cap ? want : cap);
    return BPP_OK;
}

static int check_dims(const char *who, int n, int L, int W, int H, int variant)
{
    if (n < 0) return fail(BPP_ERR_INVALID, std::string(who) + ": negative n");
    if (L < 1 || W < 1 || L > kMaxSide || W > kMaxSide) return fail(BPP_ERR_INVALID, std::string(who) + ": L and W must be in [1, 32]");
    if (H < 1 || H > 255) return fail(BPP_ERR_INVALID, std::string(who) + ": H must be in [1, 255]");
    if (variant != 1 && variant != 2) return fail(BPP_ERR_INVALID, std::string(who) + ": variant must be 1 or 2");
    return BPP_OK;
}

int32_t bpp_mask_from_obs(const float *obs, int32_t n, int32_t L, int32_t W, int32_t H, int32_t rot, int32_t variant,
                          float *mask, void *stream)
{
    if (int rc = check_dims("bpp_mask_from_obs", n, L, W, H, variant)) return rc;
    if (n == 0) return BPP_OK;
    if (!obs || !mask) return fail(BPP_ERR_INVALID, "bpp_mask_from_obs: null pointer");
    int grid = 0;
    if (int rc = stateless_grid(n, &grid)) return rc;
    BPP_CUDA(launch_mask_from_obs(obs, n, L, W, H, rot ? 1 : 0, variant, mask, grid, as_stream(stream)));
    g_launches.fetch_add(1);
    return BPP_OK;
}

int32_t bpp_mask_from_plain(const int32_t *plain, const int32_t *box, int32_t n, int32_t L, int32_t W, int32_t H,
                            int32_t rot, int32_t variant, float *mask, void *stream)
{
    if (int rc = check_dims("bpp_mask_from_plain", n, L, W, H, variant)) return rc;
    if (n == 0) return BPP_OK;
    if (!plain || !box || !mask) return fail(BPP_ERR_INVALID, "bpp_mask_from_plain: null pointer");
    int grid = 0;
    if (int rc = stateless_grid(n, &grid)) return rc;
    BPP_CUDA(launch_mask_from_plain(plain, box, n, L, W, H, rot ? 1 : 0, variant, mask, grid, as_stream(stream)));
    g_launches.fetch_add(1);
    return BPP_OK;
}

int32_t bpp_check_box(const int32_t *plain, const int32_t *query, int32_t n, int32_t L, int32_t W, int32_t H,
                      int32_t variant, int32_t *out, void *stream)
{
    if (int rc = check_dims("bpp_check_box", n, L, W, H, variant)) return rc;
    if (n == 0) return BPP_OK;
    if (!plain || !query || !out) return fail(BPP_ERR_INVALID, "bpp_check_box: null pointer");
    int grid = 0;
    if (int rc = stateless_grid(n, &grid)) return rc;   // device capability check
    BPP_CUDA(launch_check_box(plain, query, n, L, W, H, variant, out, as_stream(stream)));
    g_launches.fetch_add(1);
    return BPP_OK;
}

int32_t bpp_export_state(bpp_env *env, int32_t *plain, int32_t *box, int32_t *counter, int32_t *vol, void *stream)
{
    if (int rc = enter(env, "bpp_export_state")) return rc;
    BPP_CUDA(launch_export_state(env->p, plain, box, counter, vol, as_stream(stream)));
    g_launches.fetch_add(1);
    return BPP_OK;
}

int32_t bpp_import_state(bpp_env *env, const int32_t *plain, const int32_t *box, const int32_t *counter,
                         const int32_t *vol, void *stream)
{
    if (int rc = enter(env, "bpp_import_state")) return rc;
    // the per-bin generator state (Philox counters, replay row, pre-cut episodes) only exists after a reset
    if (!env->was_reset) return fail(BPP_ERR_STATE, "bpp_import_state: call bpp_reset first");
    cudaStream_t st = as_stream(stream);
    BPP_CUDA(cudaMemsetAsync(env->d_flag, 0, sizeof(int32_t), st));
    BPP_CUDA(launch_import_state(env->p, plain, box, counter, vol, env->d_flag, st));
    g_launches.fetch_add(1);
    int32_t bad = 0;
    BPP_CUDA(cudaMemcpyAsync(&bad, env->d_flag, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    BPP_CUDA(cudaStreamSynchronize(st));
    if (bad) return fail(BPP_ERR_INVALID, "bpp_import_state: " + std::to_string(bad) + " heights outside [0, H] (left unchanged)");
    return BPP_OK;
}

namespace {
struct StateHeader {
    uint32_t magic, n_bins, A_pad, gen_kind;
    uint64_t steps, ep_bytes;
    uint32_t was_reset, pad;
};
struct StatePart { void *ptr; size_t bytes; };
// device arrays that make up the env state, in blob order
int state_parts(bpp_env *env, StatePart *parts)
{
    const EnvParams &p = env->p;
    const size_t N = (size_t)p.N;
    int n = 0;
    parts[n++] = {env->d_hm, N * p.A_pad};
    parts[n++] = {env->d_sc, N * sizeof(BinScalars)};
    parts[n++] = {env->d_metrics, 4 * sizeof(unsigned long long)};
    if (p.gen_kind >= 2) {
        parts[n++] = {env->d_ep_items, 2 * N * p.ep_cap * sizeof(uint32_t)};
        parts[n++] = {env->d_ep_len2, 2 * N * sizeof(int32_t)};
    }
    return n;
}
}  // namespace

int32_t bpp_episode_items_host(bpp_env *env, int64_t bin, uint8_t *items_host, int32_t cap_items, int32_t *len_host,
                               void *stream)
{
    if (int rc = enter(env, "bpp_episode_items_host")) return rc;
    if (env->p.gen_kind < 2) return fail(BPP_ERR_INVALID, "bpp_episode_items_host: only for the cut1 / cut2 generators");
    if (bin < 0 || bin >= env->p.N || !items_host || !len_host || cap_items < 0)
        return fail(BPP_ERR_INVALID, "bpp_episode_items_host: bad arguments");
    if (!env->was_reset) return fail(BPP_ERR_STATE, "bpp_episode_items_host: call bpp_reset first");
    cudaStream_t st = as_stream(stream);
    BinScalars s;
    BPP_CUDA(cudaMemcpyAsync(&s, env->d_sc + bin, sizeof(s), cudaMemcpyDeviceToHost, st));
    BPP_CUDA(cudaStreamSynchronize(st));
    const int buf = (int)(s.traj & 1u);
    const int n = s.ep_len < cap_items ? s.ep_len : cap_items;
    std::vector<uint32_t> packed((size_t)(n > 0 ? n : 1));
    if (n > 0) {
        BPP_CUDA(cudaMemcpyAsync(packed.data(), env->d_ep_items + ((size_t)buf * env->p.N + (size_t)bin) * env->p.ep_cap,
                                 (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        BPP_CUDA(cudaStreamSynchronize(st));
    }
    for (int k = 0; k < n; ++k) {
        items_host[3 * k] = (uint8_t)(packed[k] & 0xff);
        items_host[3 * k + 1] = (uint8_t)((packed[k] >> 8) & 0xff);
        items_host[3 * k + 2] = (uint8_t)((packed[k] >> 16) & 0xff);
    }
    *len_host = s.ep_len;
    return BPP_OK;
}

int64_t bpp_state_bytes(bpp_env *env)
{
    if (!env) return fail(BPP_ERR_INVALID, "bpp_state_bytes: null env");
    StatePart parts[8];
    const int n = state_parts(env, parts);
    size_t total = 256;                                   // header, padded
    for (int k = 0; k < n; ++k) total += (parts[k].bytes + 255) & ~(size_t)255;
    return (int64_t)total;
}

int32_t bpp_save_state(bpp_env *env, void *blob, void *stream)
{
    if (int rc = enter(env, "bpp_save_state")) return rc;
    if (!blob) return fail(BPP_ERR_INVALID, "bpp_save_state: null blob");
    StatePart parts[8];
    const int n = state_parts(env, parts);
    StateHeader h{0x42505031u, (uint32_t)env->p.N, (uint32_t)env->p.A_pad, (uint32_t)env->p.gen_kind, env->steps,
                  (uint64_t)bpp_state_bytes(env), env->was_reset ? 1u : 0u, 0u};
    cudaStream_t st = as_stream(stream);
    BPP_CUDA(cudaMemcpyAsync(blob, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    BPP_CUDA(cudaStreamSynchronize(st));                  // h lives on this stack frame
    size_t off = 256;
    for (int k = 0; k < n; ++k) {
        BPP_CUDA(cudaMemcpyAsync(static_cast<uint8_t *>(blob) + off, parts[k].ptr, parts[k].bytes, cudaMemcpyDeviceToDevice, st));
        off += (parts[k].bytes + 255) & ~(size_t)255;
    }
    return BPP_OK;
}

int32_t bpp_load_state(bpp_env *env, const void *blob, void *stream)
{
    if (int rc = enter(env, "bpp_load_state")) return rc;
    if (!blob) return fail(BPP_ERR_INVALID, "bpp_load_state: null blob");
    cudaStream_t st = as_stream(stream);
    StateHeader h{};
    BPP_CUDA(cudaMemcpyAsync(&h, blob, sizeof(h), cudaMemcpyDeviceToHost, st));
    BPP_CUDA(cudaStreamSynchronize(st));
    if (h.magic != 0x42505031u || h.n_bins != (uint32_t)env->p.N || h.A_pad != (uint32_t)env->p.A_pad ||
        h.gen_kind != (uint32_t)env->p.gen_kind || h.ep_bytes != (uint64_t)bpp_state_bytes(env))
        return fail(BPP_ERR_INVALID, "bpp_load_state: blob does not match this env's configuration");
    StatePart parts[8];
    const int n = state_parts(env, parts);
    size_t off = 256;
    for (int k = 0; k < n; ++k) {
        BPP_CUDA(cudaMemcpyAsync(parts[k].ptr, static_cast<const uint8_t *>(blob) + off, parts[k].bytes, cudaMemcpyDeviceToDevice, st));
        off += (parts[k].bytes + 255) & ~(size_t)255;
    }
    if (env->p.gen_kind >= 2) BPP_CUDA(cudaMemsetAsync(env->d_gen_count, 0, sizeof(int32_t), st));
    env->steps = h.steps;
    env->was_reset = h.was_reset != 0;
    return BPP_OK;
}

int32_t bpp_metrics_host(bpp_env *env, uint64_t *out_host, void *stream)
{
    if (int rc = enter(env, "bpp_metrics_host")) return rc;
    if (!out_host) return fail(BPP_ERR_INVALID, "bpp_metrics_host: null pointer");
    unsigned long long tmp[4] = {0, 0, 0, 0};
    BPP_CUDA(cudaMemcpyAsync(tmp, env->d_metrics, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, as_stream(stream)));
    BPP_CUDA(cudaStreamSynchronize(as_stream(stream)));
    out_host[0] = tmp[0]; out_host[1] = tmp[1]; out_host[2] = tmp[2]; out_host[3] = env->steps;
    return BPP_OK;
}

int32_t bpp_debug_draws(uint64_t seed, uint32_t domain, uint32_t aux, uint64_t gid, int32_t n, uint32_t *out, void *stream)
{
    if (n < 0 || (n > 0 && !out)) return fail(BPP_ERR_INVALID, "bpp_debug_draws: bad arguments");
    BPP_CUDA(launch_debug_draws(seed, domain, aux, gid, n, out, as_stream(stream)));
    g_launches.fetch_add(1);
    return BPP_OK;
}

int32_t bpp_debug_curand_draws(uint64_t seed, uint64_t subsequence, int32_t n, uint32_t *out, void *stream)
{
    if (n < 0 || (n > 0 && !out)) return fail(BPP_ERR_INVALID, "bpp_debug_curand_draws: bad arguments");
    BPP_CUDA(launch_debug_curand(seed, subsequence, n, out, as_stream(stream)));
    g_launches.fetch_add(1);
    return BPP_OK;
}

}  // extern "C"
