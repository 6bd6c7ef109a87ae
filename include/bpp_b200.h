/*
 * bpp_b200.h -- C ABI of libbpp_b200.so, the B200 (sm_100a) batched 3D bin-packing environment.
 *
 * This is the drop-in boundary for ONE hot path of alexfrom0815/Online-3D-BPP-DRL: what a rollout
 * does to N independent bins per step -- action decode, placement feasibility, heightmap update,
 * reward, termination, auto-reset, next item, float32 observation and ground-truth action mask.
 * The reference is pure Python and has no FFI; each entry point below names the Python interface it
 * replaces (paths relative to the reference root).  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.  `stream` is a cudaStream_t passed
 *     as void* (NULL = legacy default stream).  Unless a name ends in `_host`, every data pointer
 *     is a DEVICE pointer on the env's device and is caller-owned; calls are asynchronous on
 *     `stream` and never synchronise.
 *   - Return value: 0 (BPP_OK) or a negative bpp_status; bpp_last_error() gives the message of the
 *     last failure on the calling thread.  In-episode events (infeasible or out-of-range action)
 *     are NOT errors: they end the episode exactly as envs/bpp0/space.py:111-115 does.
 *   - Layouts: heightmap index lx*W + ly (lx < L = container_size[0], ly < W = container_size[1]);
 *     observation float32[4*L*W] = [heightmap | x plane | y plane | z plane]
 *     (envs/bpp0/bin3D.py:61-66); mask float32[L*W*(1+rot)], rotated orientation second
 *     (acktr/utils.py:64-94).
 *   - One env per device context, not thread-safe (the reference's VecEnv is not either:
 *     baselines/common/vec_env/shmem_vec_env.py:58-64).
 */
#ifndef BPP_B200_H
#define BPP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BPP_ABI_VERSION 2

typedef enum {
    BPP_OK = 0,
    BPP_ERR_INVALID = -1,      /* bad argument / unsupported configuration */
    BPP_ERR_CUDA = -2,         /* a CUDA runtime call failed               */
    BPP_ERR_NO_DEVICE = -3,    /* no usable sm_100 device                  */
    BPP_ERR_STATE = -4         /* call order violation (e.g. step before reset) */
} bpp_status;

/* Item-sequence generators: --item-seq rs | cut1 | cut2 (acktr/arguments.py:24-27) and the dataset
 * replay of test mode (envs/bpp0/binCreator.py:42-72). */
typedef enum { BPP_GEN_RS = 0, BPP_GEN_REPLAY = 1, BPP_GEN_CUT1 = 2, BPP_GEN_CUT2 = 3 } bpp_gen_kind;

/* Which feasibility rule fills the mask: 1 = Space.check_box (envs/bpp0/space.py:111-144, the
 * env-side mask of bin3D.py:72-93), 2 = acktr.utils.check_box (acktr/utils.py:8-35, the training
 * mask of main.py:163-168).  step() itself always places with rule 1, like space.py:174. */
typedef enum { BPP_MASK_V1 = 1, BPP_MASK_V2 = 2 } bpp_mask_variant;

typedef struct {
    int32_t abi_version;      /* BPP_ABI_VERSION                                               */
    int32_t device;           /* CUDA device ordinal                                           */
    int32_t n_bins;           /* bins owned by this env (this rank's shard)                    */
    int32_t L, W, H;          /* container_size (envs/bpp0/bin3D.py:10)                        */
    int32_t rot;              /* enable_rotation                                               */
    int32_t mask_variant;     /* bpp_mask_variant                                              */
    int32_t gen_kind;         /* bpp_gen_kind                                                  */
    int32_t n_items;          /* rs: item-set size; HOST pointer item_set[n_items][3]          */
    const uint8_t *item_set_host;   /* box_set = args.box_size_set, acktr/arguments.py:122-128 */
    int32_t n_traj, traj_cap; /* replay: trajectories, padded row length                       */
    const uint8_t *traj_host; /* HOST [n_traj][traj_cap][3]                                    */
    const int32_t *traj_len_host;   /* HOST [n_traj]                                           */
    int32_t first_traj;       /* trajectory of (global bin 0, episode 0); the reference starts at 1
                                 (binCreator.py:55-56).  Bin g, episode e plays
                                 (first_traj + g + e * global_n) mod n_traj.                   */
    int32_t lo[3], hi[3];     /* cut1: per-axis item range; cut2 uses lo[0], hi[0]             */
    uint64_t seed;            /* Philox4x32-10 key                                             */
    int64_t global_offset;    /* global id of local bin 0 (rank * n_bins when sharded)         */
    int64_t global_n;         /* bins over all shards (0 => n_bins)                            */
} bpp_config;

typedef struct bpp_env bpp_env;

/* -- library ------------------------------------------------------------------------------------ */
int32_t bpp_abi_version(void);
const char *bpp_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
uint64_t bpp_launch_count(void);
/* The build switches this library was compiled with ("packed_sweep=1 lane_check=1 ..."), for bench provenance. */
const char *bpp_build_info(void);

/* -- lifetime: replaces gym.make('Bpp-v0', ...) x N in acktr/envs.py:31-41,86-101 ---------------- */
int32_t bpp_create(const bpp_config *cfg, bpp_env **out);
int32_t bpp_destroy(bpp_env *env);

/* -- VecEnv.reset(): baselines/common/vec_env/shmem_vec_env.py:61-67 -> PackingGame.reset
 *    (envs/bpp0/bin3D.py:55-59).  obs float32[N][4LW], mask float32[N][LW(1+rot)], sample int64[N]
 *    (a uniformly random valid action per bin, for policy-free rollouts); any of them may be NULL. */
int32_t bpp_reset(bpp_env *env, float *obs, float *mask, int64_t *sample, void *stream);

/* -- VecEnv.step(): shmem_vec_env.py:69-79,127-130 -> PackingGame.step (bin3D.py:95-127) with the
 *    worker's auto-reset, plus the trainer's mask loop main.py:163-169, fused into one launch.
 *    actions int64[N] (the LongTensor[N,1] of acktr/envs.py:182-187, squeezed).
 *    reward float32[N], done uint8[N], counter int32[N] (info['counter']), ratio float32[N]
 *    (info['ratio']) describe the step just taken; obs/mask/sample describe the state AFTER it,
 *    i.e. of the fresh episode when done.  Outputs may be NULL. */
int32_t bpp_step(bpp_env *env, const int64_t *actions, float *obs, float *mask, float *reward,
                 uint8_t *done, int32_t *counter, float *ratio, int64_t *sample, void *stream);

/* -- host-buffer variant of step: what VecPyTorch.step_async/step_wait (acktr/envs.py:182-193) do
 *    around the env.  actions/reward/done/counter/ratio are HOST buffers (pinned for full speed);
 *    obs/mask/sample stay on the device (as the reference's obs does, acktr/envs.py:191).
 *    Copies H2D, launches, copies D2H and synchronises `stream` before returning. */
int32_t bpp_step_host(bpp_env *env, const int64_t *actions_host, float *obs, float *mask,
                      float *reward_host, uint8_t *done_host, int32_t *counter_host,
                      float *ratio_host, int64_t *sample, void *stream);

/* -- step with packed results.  What VecPyTorch.step_wait hands back per env (acktr/envs.py:189-193: reward, and through
 *    infos the counter / ratio of bin3D.py:111,123-125) travels as ONE 16-byte record per bin, written with a single
 *    128-bit store (or one PCIe write when `results` is pinned host memory):
 *      reward       10 * item volume / (L*W*H) of the step, 0 when the episode ended   (bin3D.py:121)
 *      vol          placed volume after the step; ratio = vol / (L*W*H) and Monitor's episode return r = 10 * vol / (L*W*H)
 *                   follow exactly on the host (space.py:146-151, baselines/bench/monitor.py:64)
 *      counter      len(space.boxes) after the step (info['counter'])
 *      next_action  a uniformly random valid action for the state after the step (policy stand-in; ignore otherwise)
 *    `done` stays a separate byte array so that the host can take it as np.bool_[N] without a pass over the records.
 *    n_steps >= 1 env steps run in ONE launch: step 1 takes `actions`, every later step the action the previous one
 *    sampled (policy-free rollouts on small envs, where a launch per step costs more than the step); rs / replay
 *    generators only.  results / done describe the LAST step; the episode metrics count all of them.
 *    `sample` (int64[N], may be NULL) receives next_action as well, for callers that feed it back as `actions`. */
typedef struct {
    float reward;
    int32_t vol;
    int32_t counter;
    int32_t next_action;
} bpp_result;

int32_t bpp_step_packed(bpp_env *env, const int64_t *actions, float *obs, float *mask, bpp_result *results,
                        uint8_t *done, int64_t *sample, int32_t n_steps, void *stream);

/* Same with HOST results / done (pinned and mapped, e.g. from bpp_host_alloc: the kernel writes them in place over
 * PCIe; ordinary memory is staged through device copies).  actions_host may be pinned host memory (read in place),
 * ordinary host memory (staged) or a device pointer (a policy that lives on the GPU).  `sample` (device
 * memory, or pinned mapped host memory such as the caller's next action buffer; may be NULL) receives the next sampled
 * action as int64 as well.  One step; synchronises `stream`. */
int32_t bpp_step_packed_host(bpp_env *env, const int64_t *actions_host, float *obs, float *mask,
                             bpp_result *results_host, uint8_t *done_host, int64_t *sample, void *stream);

/* The same with int32 actions read with a stride (in int32 units): half the bytes on the wire, and
 * (actions_host = &results_prev[0].next_action, stride = 4) takes a policy-free rollout's actions straight out of the
 * records of the previous step.  actions / results / done must be pinned mapped host memory or device memory. */
int32_t bpp_step_packed_host_i32(bpp_env *env, const int32_t *actions_host, int32_t stride, float *obs, float *mask,
                                 bpp_result *results_host, uint8_t *done_host, int64_t *sample, void *stream);

/* -- pinned, device-mapped host memory placed on the NUMA node of `device` (the calling thread is bound to that node's
 *    CPUs while the pages are allocated and touched, then gets its affinity back), and a helper that binds the calling
 *    thread there for good -- what a per-GPU rollout worker should do before it starts launching. */
int32_t bpp_host_alloc(int32_t device, uint64_t bytes, void **out);
int32_t bpp_host_free(void *ptr);
int32_t bpp_bind_thread_to_device(int32_t device);

/* -- stateless mask: acktr/utils.py:37-62 get_possible_position / :64-94 get_rotation_mask on a batch
 *    of observations (float32[n][4LW], as delivered by VecPyTorch), and the same on raw state
 *    (plain int32[n][LW] = Space.plain, box int32[n][3]).  mask float32[n][LW(1+rot)]. */
int32_t bpp_mask_from_obs(const float *obs, int32_t n, int32_t L, int32_t W, int32_t H, int32_t rot,
                          int32_t variant, float *mask, void *stream);
int32_t bpp_mask_from_plain(const int32_t *plain, const int32_t *box, int32_t n, int32_t L, int32_t W,
                            int32_t H, int32_t rot, int32_t variant, float *mask, void *stream);

/* -- batched check_box: envs/bpp0/space.py:111-144 (variant 1) / acktr/utils.py:8-35 (variant 2).
 *    query int32[n][5] = (x, y, lx, ly, z) evaluated on plain int32[n][LW]; out int32[n] = max_h or -1. */
int32_t bpp_check_box(const int32_t *plain, const int32_t *query, int32_t n, int32_t L, int32_t W,
                      int32_t H, int32_t variant, int32_t *out, void *stream);

/* -- state in reference dtypes (Space.plain int32[N][LW], next box int32[N][3], len(space.boxes)
 *    int32[N], placed volume int32[N]); NULL pointers are skipped.  import needs an env that has been reset (the item
 *    generators' per-bin state comes from there) and heights within [0, H], else BPP_ERR_STATE / BPP_ERR_INVALID. */
int32_t bpp_export_state(bpp_env *env, int32_t *plain, int32_t *box, int32_t *counter, int32_t *vol,
                         void *stream);
int32_t bpp_import_state(bpp_env *env, const int32_t *plain, const int32_t *box, const int32_t *counter,
                         const int32_t *vol, void *stream);
/* obs/mask of the current state without stepping (cur_observation, bin3D.py:61-66). */
int32_t bpp_observe(bpp_env *env, float *obs, float *mask, void *stream);

/* -- cut1 / cut2 only: the item sequence of the episode bin `bin` is currently playing (what CuttingBoxCreator /
 *    MDlayerBoxCreator would hand out one by one, cutCreator.py:112-128, mdCreator.py:163-166), for host-side callers
 *    that need box_creator.preview(k).  items_host receives up to cap_items triples, *len_host the episode length.
 *    Synchronises `stream`. */
int32_t bpp_episode_items_host(bpp_env *env, int64_t bin, uint8_t *items_host, int32_t cap_items, int32_t *len_host,
                               void *stream);

/* -- checkpoint / resume (the reference checkpoints only the policy, main.py:186-191; the env state here is a few
 *    device arrays, so a rollout can be resumed bit-exactly): bpp_state_bytes() is the size of an opaque blob,
 *    bpp_save_state / bpp_load_state copy it to / from a caller-owned DEVICE buffer on `stream`.  A blob can only be
 *    loaded into an env created with the same configuration. */
int64_t bpp_state_bytes(bpp_env *env);
int32_t bpp_save_state(bpp_env *env, void *blob, void *stream);
int32_t bpp_load_state(bpp_env *env, const void *blob, void *stream);

/* -- episode metrics accumulated on the device since creation (what main.py:159-162 collects from
 *    infos): out_host[4] = {episodes, sum of counter, sum of placed volume, env-steps}.
 *    Synchronises `stream`.  Cross-rank aggregation is one NCCL all-reduce of these 4 integers. */
int32_t bpp_metrics_host(bpp_env *env, uint64_t *out_host, void *stream);

/* -- Philox4x32-10 self-test hook: out[n] = the n-th draw of stream (seed, domain, aux, gid) computed
 *    on the device (tests compare with the oracle and with cuRAND's stream layout). */
int32_t bpp_debug_draws(uint64_t seed, uint32_t domain, uint32_t aux, uint64_t gid, int32_t n,
                        uint32_t *out, void *stream);

/* out[i] = i-th curand() of cuRAND's own Philox4_32_10 after curand_init(seed, subsequence, 0). */
int32_t bpp_debug_curand_draws(uint64_t seed, uint64_t subsequence, int32_t n, uint32_t *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BPP_B200_H */
