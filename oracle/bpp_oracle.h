/*
 * bpp_oracle.h -- CPU restatement ("oracle") of the reference's batched 3D-BPP hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product path (online-3d-bpp-drl_b200/) never links, imports or falls back to this code.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against
 *   - tests/golden/ref_rollouts_*.npz  (traces recorded from the imported Python reference
 *     by tests/golden/make_golden.py: masks v1/v2/rotation, heightmaps, reward, done, info),
 *   - the reference's own known-answer data (user_study/trajs/traj0.txt, user_study/cut_2.txt
 *     sample, dataset/cut_2.pt) re-encoded as tests/golden/ npz files,
 *   - and, when /root/reference is importable, live differential fuzzing.
 * The item GENERATORS (rs / cut1 / cut2) are pinned distributionally only: the reference draws
 * from CPython/NumPy MT19937 streams, this restatement (and the CUDA product) from Philox4x32-10.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 */
#ifndef BPP_ORACLE_H
#define BPP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_GEN_RS = 0, ORC_GEN_REPLAY = 1, ORC_GEN_CUT1 = 2, ORC_GEN_CUT2 = 3 };
enum { ORC_MASK_V1 = 1, ORC_MASK_V2 = 2 };

typedef struct {
    int32_t n_bins;          /* bins owned by this instance (local shard)              */
    int32_t L, W, H;         /* container_size = (L "width", W "length", H "height")   */
    int32_t rot;             /* enable_rotation                                        */
    int32_t mask_variant;    /* ORC_MASK_V1 (Space.check_box) / ORC_MASK_V2 (acktr.utils.check_box) */
    int32_t gen_kind;        /* ORC_GEN_*                                              */
    int32_t n_items;         /* rs: size of the item set                               */
    const uint8_t *item_set; /* rs: [n_items,3]                                        */
    int32_t T, K;            /* replay: number of trajectories, padded row length      */
    const uint8_t *traj;     /* replay: [T,K,3]                                        */
    const int32_t *traj_len; /* replay: [T]                                            */
    int32_t first_traj;      /* replay: trajectory of (bin 0, episode 0); reference = 1 */
    int32_t lo[3], hi[3];    /* cut1: per-axis item range; cut2 uses lo[0], hi[0]      */
    uint64_t seed;
    int64_t global_offset;   /* global id of local bin 0                               */
    int64_t global_n;        /* total bins over all shards                             */
    int32_t n_threads;       /* OpenMP threads for the batched entry points (<=0: all) */
} orc_config;

typedef struct orc_env orc_env;

/* ---- stateless pieces ------------------------------------------------------------------ */
/* Space.check_box, envs/bpp0/space.py:111-144 ("rule v1"). plain is int32[L*W] row-major.   */
int32_t orc_check_box_v1(const int32_t *plain, int32_t L, int32_t W, int32_t H,
                         int32_t x, int32_t y, int32_t lx, int32_t ly, int32_t z);
/* acktr.utils.check_box, acktr/utils.py:8-35 ("rule v2").                                   */
int32_t orc_check_box_v2(const int32_t *plain, int32_t L, int32_t W, int32_t H,
                         int32_t x, int32_t y, int32_t lx, int32_t ly, int32_t z);
/* get_possible_position / get_rotation_mask (acktr/utils.py:37-94, envs/bpp0/bin3D.py:72-93).
 * mask_out is int32[L*W*(1+rot)].                                                          */
void orc_mask(const int32_t *plain, int32_t L, int32_t W, int32_t H,
              int32_t bx, int32_t by, int32_t bz, int32_t rot, int32_t variant,
              int32_t *mask_out);
/* Philox4x32-10 block (Salmon et al. SC'11; same stream layout as cuRAND's Philox4_32_10). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* The n-th 32-bit draw of stream (seed, domain, aux, global bin id).                        */
uint32_t orc_draw(uint64_t seed, uint32_t domain, uint32_t aux, uint64_t gid, uint32_t n);

/* cut1 / cut2 episode generators (envs/bpp0/cutCreator.py:32-128, mdCreator.py:59-166).
 * Write up to cap items into items[cap][3], positions (lx,ly,lz) into pos[cap][3] if non-NULL;
 * return the number of items (without sentinel).                                            */
int32_t orc_gen_cut1(const orc_config *cfg, uint64_t gid, uint32_t episode,
                     uint8_t *items, uint8_t *pos, int32_t cap);
int32_t orc_gen_cut2(const orc_config *cfg, uint64_t gid, uint32_t episode,
                     uint8_t *items, uint8_t *pos, int32_t cap);

/* ---- batched environment ---------------------------------------------------------------- */
orc_env *orc_create(const orc_config *cfg);
void orc_destroy(orc_env *e);
/* PackingGame.reset (envs/bpp0/bin3D.py:55-59) for every bin. Any output pointer may be NULL. */
void orc_reset(orc_env *e, float *obs, float *mask, int64_t *sample);
/* PackingGame.step (bin3D.py:95-127) + worker auto-reset (shmem_vec_env.py:127-130).        */
void orc_step(orc_env *e, const int64_t *actions, float *obs, float *mask, float *reward,
              uint8_t *done, int32_t *counter, float *ratio, int64_t *sample);
/* obs/mask of the current state, nothing advances                                         */
void orc_observe(orc_env *e, float *obs, float *mask);
/* state access (reference dtypes): plain int32[N,L*W], box int32[N,3], counter/vol int32[N] */
void orc_get_state(const orc_env *e, int32_t *plain, int32_t *box, int32_t *counter, int32_t *vol);
void orc_set_state(orc_env *e, const int32_t *plain, const int32_t *box, const int32_t *counter,
                   const int32_t *vol);
/* episode accumulators since creation: {n_episodes, sum_counter, sum_vol, n_steps}          */
void orc_metrics(const orc_env *e, uint64_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* BPP_ORACLE_H */
