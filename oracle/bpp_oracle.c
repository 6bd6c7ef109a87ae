/*
 * bpp_oracle.c -- CPU restatement of the reference hot path (see bpp_oracle.h for status).
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into or called from the product.
 *
 * Data model follows the reference literally: the heightmap ("plain") is int32[L][W], row-major,
 * lx indexes the first axis (size L = container_size[0]) and ly the second (size W).
 * Thresholds `max_area/area > 0.95 | 0.85 | 0.50` (float64 in the reference) are restated as the
 * exact integer tests 20*ma > 19*a, 20*ma > 17*a, 2*ma > a; tests/test_oracle_golden.py checks the
 * equivalence against Python float division for every 0 <= ma <= a <= 4096.
 */
#include "bpp_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10                                                                               */
/* ------------------------------------------------------------------------------------------ */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Stream layout shared with the CUDA product (DESIGN.md "RNG streams"):
 *   key = (seed lo, seed hi); ctr = (n >> 2, domain << 28 | aux, gid lo, gid hi); word n & 3.
 * For domain 0 / aux 0 this is the n-th curand() output after curand_init(seed, gid, 0).   */
uint32_t orc_draw(uint64_t seed, uint32_t domain, uint32_t aux, uint64_t gid, uint32_t n)
{
    uint32_t ctr[4] = { n >> 2, (domain << 28) | (aux & 0x0FFFFFFFu),
                        (uint32_t)gid, (uint32_t)(gid >> 32) };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t out[4];
    orc_philox4x32_10(ctr, key, out);
    return out[n & 3u];
}

static inline uint32_t mulhi32(uint32_t u, uint32_t n) { return (uint32_t)(((uint64_t)u * n) >> 32); }

enum { DOM_ITEM = 0, DOM_SAMPLE = 1, DOM_GEN = 2 };

/* ------------------------------------------------------------------------------------------ */
/* Feasibility rules                                                                           */
/* ------------------------------------------------------------------------------------------ */

/* envs/bpp0/space.py:111-144 */
int32_t orc_check_box_v1(const int32_t *plain, int32_t L, int32_t W, int32_t H,
                         int32_t x, int32_t y, int32_t lx, int32_t ly, int32_t z)
{
    if (lx + x > L || ly + y > W) return -1;          /* space.py:112-113 */
    if (lx < 0 || ly < 0) return -1;                   /* space.py:114-115 */
    const int32_t *rec = plain + (int64_t)lx * W + ly; /* rec[r][c] = rec[r*W + c], space.py:117 */
    int32_t r00 = rec[0], r10 = rec[(x - 1) * W], r01 = rec[y - 1], r11 = rec[(x - 1) * W + y - 1];
    int32_t rm = r00;
    if (r10 > rm) rm = r10;
    if (r01 > rm) rm = r01;
    if (r11 > rm) rm = r11;                            /* space.py:122 */
    int32_t sc = (r00 == rm) + (r10 == rm) + (r01 == rm) + (r11 == rm);
    if (sc < 3) return -1;                             /* space.py:124-125 */
    int32_t max_h = rec[0];
    for (int32_t r = 0; r < x; ++r)
        for (int32_t c = 0; c < y; ++c)
            if (rec[r * W + c] > max_h) max_h = rec[r * W + c];   /* space.py:127 */
    int32_t ma = 0;
    for (int32_t r = 0; r < x; ++r)
        for (int32_t c = 0; c < y; ++c)
            ma += (rec[r * W + c] == max_h);           /* space.py:129 */
    int32_t a = x * y;
    if (max_h + z > H) return -1;                      /* space.py:134-135 */
    if (20 * ma > 19 * a) return max_h;                /* > 0.95, space.py:137 */
    if (rm == max_h && sc == 3 && 20 * ma > 17 * a) return max_h;  /* > 0.85, :139 */
    if (rm == max_h && sc == 4 && 2 * ma > a) return max_h;        /* > 0.50, :141 */
    return -1;
}

/* acktr/utils.py:8-35 */
int32_t orc_check_box_v2(const int32_t *plain, int32_t L, int32_t W, int32_t H,
                         int32_t x, int32_t y, int32_t lx, int32_t ly, int32_t z)
{
    if (lx + x > L || ly + y > W) return -1;           /* utils.py:9-10 */
    if (lx < 0 || ly < 0) return -1;                   /* utils.py:11-12 */
    const int32_t *rec = plain + (int64_t)lx * W + ly;
    int32_t max_h = rec[0];
    for (int32_t r = 0; r < x; ++r)
        for (int32_t c = 0; c < y; ++c)
            if (rec[r * W + c] > max_h) max_h = rec[r * W + c];   /* utils.py:15 */
    int32_t ma = 0;
    for (int32_t r = 0; r < x; ++r)
        for (int32_t c = 0; c < y; ++c)
            ma += (rec[r * W + c] == max_h);           /* utils.py:16 */
    int32_t a = x * y;
    if (max_h + z > H) return -1;                      /* utils.py:20-21 */
    int32_t nc = (rec[0] == max_h) + (rec[(x - 1) * W] == max_h) + (rec[y - 1] == max_h)
               + (rec[(x - 1) * W + y - 1] == max_h);  /* utils.py:23-26 */
    if (20 * ma > 19 * a) return max_h;                /* utils.py:28 */
    if (nc == 3 && 20 * ma > 17 * a) return max_h;     /* utils.py:30 */
    if (nc == 4 && 2 * ma > a) return max_h;           /* utils.py:32 */
    return -1;
}

static int32_t check_box(int32_t variant, const int32_t *plain, int32_t L, int32_t W, int32_t H,
                         int32_t x, int32_t y, int32_t lx, int32_t ly, int32_t z)
{
    return variant == ORC_MASK_V1 ? orc_check_box_v1(plain, L, W, H, x, y, lx, ly, z)
                                  : orc_check_box_v2(plain, L, W, H, x, y, lx, ly, z);
}

/* acktr/utils.py:37-62 (no rotation), :64-94 (rotation: second half is the (y,x) orientation and
 * the all-ones fallback applies to the combined vector); envs/bpp0/bin3D.py:72-93 with variant v1. */
void orc_mask(const int32_t *plain, int32_t L, int32_t W, int32_t H,
              int32_t bx, int32_t by, int32_t bz, int32_t rot, int32_t variant, int32_t *mask_out)
{
    const int32_t A = L * W;
    const int32_t n = A * (rot ? 2 : 1);
    int32_t total = 0;
    memset(mask_out, 0, sizeof(int32_t) * (size_t)n);
    for (int32_t o = 0; o < (rot ? 2 : 1); ++o) {
        int32_t x = o ? by : bx, y = o ? bx : by;
        for (int32_t i = 0; i < L - x + 1; ++i)
            for (int32_t j = 0; j < W - y + 1; ++j)
                if (check_box(variant, plain, L, W, H, x, y, i, j, bz) >= 0) {
                    mask_out[o * A + i * W + j] = 1;
                    ++total;
                }
    }
    if (total == 0)
        for (int32_t k = 0; k < n; ++k) mask_out[k] = 1;
}

/* ------------------------------------------------------------------------------------------ */
/* Episode generators                                                                          */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int32_t x, y, z, lx, ly, lz; } piece_t;

typedef struct { uint64_t seed, gid; uint32_t aux, k; } gen_rng;
static inline uint32_t gen_next(gen_rng *g) { return orc_draw(g->seed, DOM_GEN, g->aux, g->gid, g->k++); }

/* envs/bpp0/cutCreator.py:32-128.
 * _cut_box (:79-95) sweeps the piece list level by level: every out-of-range piece of the current
 * list is replaced by its two halves (random.choice of the offending axis :68, random.randint of
 * the cut :76), in list order, until a sweep splits nothing.  The draws are consumed in exactly
 * that order (tests/test_generator_identity.py feeds the same stream to the unmodified reference).
 * Then generate_box_size (:112-128) pops a uniformly random ready piece (_add_candidate :98-107)
 * until none is left.  The emission order does not depend on the agent, so the whole episode is
 * produced at reset. */
int32_t orc_gen_cut1(const orc_config *cfg, uint64_t gid, uint32_t episode,
                     uint8_t *items, uint8_t *pos, int32_t cap)
{
    const int32_t L = cfg->L, W = cfg->W, H = cfg->H;
    gen_rng g = { cfg->seed, gid, episode, 0 };
    /* pieces are disjoint with every side >= lo, so a list never holds more than cap of them */
    piece_t *leaf = (piece_t *)malloc(sizeof(piece_t) * (size_t)(cap + 2));
    piece_t *next = (piece_t *)malloc(sizeof(piece_t) * (size_t)(cap + 2));
    int32_t n_leaf = 1;
    leaf[0] = (piece_t){ L, W, H, 0, 0, 0 };
    for (int again = 1; again;) {
        again = 0;
        int32_t n_next = 0;
        for (int32_t i = 0; i < n_leaf; ++i) {          /* for box in self.meta_list, :84 */
            piece_t b = leaf[i];
            int32_t dim[3] = { b.x, b.y, b.z };
            int32_t bad[3], nbad = 0;
            for (int32_t a = 0; a < 3; ++a)             /* _check_box, cutCreator.py:54-58 */
                if (dim[a] < cfg->lo[a] || dim[a] > cfg->hi[a]) bad[nbad++] = a;
            if (nbad == 0) { if (n_next < cap) next[n_next++] = b; continue; }
            int32_t df = bad[mulhi32(gen_next(&g), (uint32_t)nbad)];        /* :68 */
            int32_t lo = cfg->lo[df];
            int32_t span = dim[df] - 2 * lo + 1;        /* randint(lo, dim-lo), :69-76 */
            if (span < 1) span = 1;                     /* reference asserts; unreachable for valid configs */
            int32_t cut = lo + (int32_t)mulhi32(gen_next(&g), (uint32_t)span);
            piece_t p1 = b, p2 = b;                     /* MetaBox.split, :16-26 */
            if (df == 0) { p1.x = cut; p2.x = b.x - cut; p2.lx = b.lx + cut; }
            else if (df == 1) { p1.y = cut; p2.y = b.y - cut; p2.ly = b.ly + cut; }
            else { p1.z = cut; p2.z = b.z - cut; p2.lz = b.lz + cut; }
            if (n_next + 2 <= cap) { next[n_next++] = p1; next[n_next++] = p2; }
            again = 1;
        }
        piece_t *t = leaf; leaf = next; next = t;       /* self.meta_list = copy.deepcopy(new_list), :94 */
        n_leaf = n_next;
    }
    /* ordering phase */
    int32_t *shadow = (int32_t *)calloc((size_t)(L * W), sizeof(int32_t));
    int32_t *meta = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_leaf + 1));
    int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_leaf + 1));
    int32_t n_meta = n_leaf, n_cand = 0, n_out = 0;
    for (int32_t i = 0; i < n_leaf; ++i) meta[i] = i;
    for (;;) {
        /* _add_candidate, :98-107: ready <=> shadow == lz on the whole footprint */
        int32_t keep = 0;
        for (int32_t m = 0; m < n_meta; ++m) {
            const piece_t *p = &leaf[meta[m]];
            int32_t ready = 1;
            for (int32_t r = 0; r < p->x && ready; ++r)
                for (int32_t c = 0; c < p->y; ++c)
                    if (shadow[(p->lx + r) * W + p->ly + c] != p->lz) { ready = 0; break; }
            if (ready) cand[n_cand++] = meta[m]; else meta[keep++] = meta[m];
        }
        n_meta = keep;
        if (n_cand == 0) break;                         /* :113-115 -> sentinel from now on */
        int32_t idx = (int32_t)mulhi32(gen_next(&g), (uint32_t)n_cand);   /* :116 */
        const piece_t p = leaf[cand[idx]];
        for (int32_t m = idx; m + 1 < n_cand; ++m) cand[m] = cand[m + 1];  /* list.pop(idx) */
        --n_cand;
        int32_t ox = p.x, oy = p.y;
        if (cfg->rot) {                                 /* :120-126: rd < 0.5 keeps, else swaps */
            uint32_t coin = gen_next(&g);
            if (coin >= 0x80000000u) { ox = p.y; oy = p.x; }
        }
        if (n_out < cap) {
            items[n_out * 3 + 0] = (uint8_t)ox; items[n_out * 3 + 1] = (uint8_t)oy;
            items[n_out * 3 + 2] = (uint8_t)p.z;
            if (pos) { pos[n_out * 3 + 0] = (uint8_t)p.lx; pos[n_out * 3 + 1] = (uint8_t)p.ly;
                       pos[n_out * 3 + 2] = (uint8_t)p.lz; }
            ++n_out;
        }
        for (int32_t r = 0; r < p.x; ++r)               /* _update, :109-110 */
            for (int32_t c = 0; c < p.y; ++c) shadow[(p.lx + r) * W + p.ly + c] += p.z;
    }
    free(leaf); free(next); free(shadow); free(meta); free(cand);
    return n_out;
}

/* envs/bpp0/mdCreator.py:59-166.  gen_benchmark (:117-135) iterates `invalid_box` with a Python
 * for-loop while removing the current element and appending its halves: after a removal the
 * iterator index still advances, so the element that slid into the vacated slot waits for the
 * next sweep.  The index arithmetic below reproduces that traversal, which fixes the order of
 * the valid list and therefore the tie order of the final stable sort by low_bound (:137-138). */
int32_t orc_gen_cut2(const orc_config *cfg, uint64_t gid, uint32_t episode,
                     uint8_t *items, uint8_t *pos, int32_t cap)
{
    const int32_t lo = cfg->lo[0], hi = cfg->hi[0];
    gen_rng g = { cfg->seed, gid, episode, 0 };
    piece_t *inv = (piece_t *)malloc(sizeof(piece_t) * (size_t)(cap + 64));
    piece_t *val = (piece_t *)malloc(sizeof(piece_t) * (size_t)(cap + 64));
    int32_t n_inv = 0, n_val = 0;
    inv[n_inv++] = (piece_t){ cfg->L, cfg->W, cfg->H, 0, 0, 0 };
    while (n_inv > 0) {
        int32_t i = 0;
        while (i < n_inv) {
            piece_t b = inv[i];
            int32_t dim[3] = { b.x, b.y, b.z };
            int32_t fl[3], nf = 0;
            for (int32_t a = 0; a < 3; ++a) if (dim[a] > hi) fl[nf++] = a;   /* :60-66 */
            if (nf == 0) {                              /* unreachable for lo <= hi, hi+1 >= 2*lo (the
                                                           reference would raise IndexError here) */
                for (int32_t m = i; m + 1 < n_inv; ++m) inv[m] = inv[m + 1];
                --n_inv; if (n_val < cap) val[n_val++] = b; ++i; continue;
            }
            int32_t ax = fl[mulhi32(gen_next(&g), (uint32_t)nf)];              /* :68 */
            int32_t cut = 1 + (int32_t)mulhi32(gen_next(&g), (uint32_t)dim[ax]);  /* randint(1, dim) */
            if (cut < lo || dim[ax] - cut < lo) { ++i; continue; }           /* rejected: retry next sweep */
            piece_t p1 = b, p2 = b;
            if (ax == 0) { p1.x = cut; p2.x = b.x - cut; p2.lx = b.lx + cut; }          /* :70-79 */
            else if (ax == 1) { p1.y = cut; p2.y = b.y - cut; p2.ly = b.ly + cut; }     /* :80-89 */
            else { p1.z = b.z - cut; p2.z = cut; p2.lz = b.lz + b.z - cut; }             /* :90-99 */
            for (int32_t m = i; m + 1 < n_inv; ++m) inv[m] = inv[m + 1];    /* invalid_box.remove(box) */
            --n_inv;
            piece_t sub[2] = { p1, p2 };
            for (int32_t s = 0; s < 2; ++s) {           /* is_valid, :110-115 */
                const piece_t *q = &sub[s];
                int ok = q->x >= lo && q->x <= hi && q->y >= lo && q->y <= hi && q->z >= lo && q->z <= hi;
                if (ok) { if (n_val < cap) val[n_val++] = *q; }
                else inv[n_inv++] = *q;
            }
            ++i;                                        /* iterator advances past the vacated slot */
        }
    }
    /* depart_box: stable sort by low_bound (= lz), :137-138 */
    int32_t n_out = 0;
    for (int32_t level = 0; level <= cfg->H && n_out < cap; ++level)
        for (int32_t v = 0; v < n_val && n_out < cap; ++v)
            if (val[v].lz == level) {
                items[n_out * 3 + 0] = (uint8_t)val[v].x; items[n_out * 3 + 1] = (uint8_t)val[v].y;
                items[n_out * 3 + 2] = (uint8_t)val[v].z;
                if (pos) { pos[n_out * 3 + 0] = (uint8_t)val[v].lx; pos[n_out * 3 + 1] = (uint8_t)val[v].ly;
                           pos[n_out * 3 + 2] = (uint8_t)val[v].lz; }
                ++n_out;
            }
    free(inv); free(val);
    return n_out;
}

/* ------------------------------------------------------------------------------------------ */
/* Batched environment                                                                         */
/* ------------------------------------------------------------------------------------------ */
struct orc_env {
    orc_config cfg;
    uint8_t *item_set, *traj;
    int32_t *traj_len;
    int32_t A, R, cap;
    int32_t *plain;     /* [N][A]   Space.plain, space.py:22   */
    int32_t *box;       /* [N][3]   box_creator.box_list[0]    */
    int32_t *counter;   /* [N]      len(space.boxes)           */
    int32_t *vol;       /* [N]      sum of placed volumes      */
    uint32_t *episode;  /* [N]      resets so far              */
    uint32_t *ctr;      /* [N]      reset+step calls so far    */
    int32_t *pos;       /* [N]      replay / cut position      */
    uint8_t *ep_items;  /* [N][cap][3] cut1/cut2 episode items */
    int32_t *ep_len;    /* [N]                                 */
    int fresh;
    uint64_t met[4];
};

static int32_t cut_cap(const orc_config *c)
{
    int32_t l0 = c->lo[0] > 0 ? c->lo[0] : 1, l1 = c->lo[1] > 0 ? c->lo[1] : 1, l2 = c->lo[2] > 0 ? c->lo[2] : 1;
    if (c->gen_kind == ORC_GEN_CUT2) { l1 = l0; l2 = l0; }
    return (c->L / l0) * (c->W / l1) * (c->H / l2);
}

orc_env *orc_create(const orc_config *cfg)
{
    orc_env *e = (orc_env *)calloc(1, sizeof(orc_env));
    e->cfg = *cfg;
    const int32_t N = cfg->n_bins;
    e->A = cfg->L * cfg->W;
    e->R = cfg->rot ? 2 : 1;
    if (cfg->gen_kind == ORC_GEN_RS) {
        e->item_set = (uint8_t *)malloc((size_t)cfg->n_items * 3);
        memcpy(e->item_set, cfg->item_set, (size_t)cfg->n_items * 3);
        e->cfg.item_set = e->item_set;
    }
    if (cfg->gen_kind == ORC_GEN_REPLAY) {
        e->traj = (uint8_t *)malloc((size_t)cfg->T * cfg->K * 3);
        memcpy(e->traj, cfg->traj, (size_t)cfg->T * cfg->K * 3);
        e->traj_len = (int32_t *)malloc(sizeof(int32_t) * (size_t)cfg->T);
        memcpy(e->traj_len, cfg->traj_len, sizeof(int32_t) * (size_t)cfg->T);
        e->cfg.traj = e->traj; e->cfg.traj_len = e->traj_len;
    }
    e->plain = (int32_t *)calloc((size_t)N * e->A, sizeof(int32_t));
    e->box = (int32_t *)calloc((size_t)N * 3, sizeof(int32_t));
    e->counter = (int32_t *)calloc((size_t)N, sizeof(int32_t));
    e->vol = (int32_t *)calloc((size_t)N, sizeof(int32_t));
    e->episode = (uint32_t *)calloc((size_t)N, sizeof(uint32_t));
    e->ctr = (uint32_t *)calloc((size_t)N, sizeof(uint32_t));
    e->pos = (int32_t *)calloc((size_t)N, sizeof(int32_t));
    e->ep_len = (int32_t *)calloc((size_t)N, sizeof(int32_t));
    if (cfg->gen_kind == ORC_GEN_CUT1 || cfg->gen_kind == ORC_GEN_CUT2) {
        e->cap = cut_cap(cfg);
        e->ep_items = (uint8_t *)calloc((size_t)N * e->cap * 3, 1);
    }
    e->fresh = 1;
    return e;
}

void orc_destroy(orc_env *e)
{
    if (!e) return;
    free(e->item_set); free(e->traj); free(e->traj_len); free(e->plain); free(e->box);
    free(e->counter); free(e->vol); free(e->episode); free(e->ctr); free(e->pos);
    free(e->ep_items); free(e->ep_len); free(e);
}

/* box_creator.reset() + generate_box_size(): make the first item of episode e->episode[b]. */
static void begin_episode(orc_env *e, int32_t b)
{
    const orc_config *c = &e->cfg;
    const uint64_t gid = (uint64_t)(c->global_offset + b);
    e->pos[b] = 0;
    if (c->gen_kind == ORC_GEN_CUT1)
        e->ep_len[b] = orc_gen_cut1(c, gid, e->episode[b], e->ep_items + (size_t)b * e->cap * 3, NULL, e->cap);
    else if (c->gen_kind == ORC_GEN_CUT2)
        e->ep_len[b] = orc_gen_cut2(c, gid, e->episode[b], e->ep_items + (size_t)b * e->cap * 3, NULL, e->cap);
}

/* preview(1)[0] for the bin's current generator position. */
static void fetch_item(orc_env *e, int32_t b)
{
    const orc_config *c = &e->cfg;
    const uint64_t gid = (uint64_t)(c->global_offset + b);
    int32_t *box = e->box + (size_t)b * 3;
    switch (c->gen_kind) {
    case ORC_GEN_RS: {                                  /* binCreator.py:38-40 */
        uint32_t idx = mulhi32(orc_draw(c->seed, DOM_ITEM, 0, gid, e->ctr[b]), (uint32_t)c->n_items);
        box[0] = c->item_set[idx * 3]; box[1] = c->item_set[idx * 3 + 1]; box[2] = c->item_set[idx * 3 + 2];
        break;
    }
    case ORC_GEN_REPLAY: {                              /* binCreator.py:51-72 */
        uint64_t T = (uint64_t)c->T;
        uint64_t t = (((uint64_t)c->first_traj + gid) % T
                      + ((uint64_t)e->episode[b] % T) * ((uint64_t)c->global_n % T)) % T;
        int32_t p = e->pos[b];
        if (p < c->traj_len[t]) {
            const uint8_t *it = c->traj + ((size_t)t * c->K + p) * 3;
            box[0] = it[0]; box[1] = it[1]; box[2] = it[2];
        } else { box[0] = 10; box[1] = 10; box[2] = 10; }  /* :62, :69-71 sentinel, hard-coded */
        break;
    }
    case ORC_GEN_CUT1: case ORC_GEN_CUT2: {
        int32_t p = e->pos[b];
        if (p < e->ep_len[b]) {
            const uint8_t *it = e->ep_items + ((size_t)b * e->cap + p) * 3;
            box[0] = it[0]; box[1] = it[1]; box[2] = it[2];
        } else if (c->gen_kind == ORC_GEN_CUT1) {       /* cutCreator.py:113-115: bin_size */
            box[0] = c->L; box[1] = c->W; box[2] = c->H;
        } else { box[0] = 10; box[1] = 10; box[2] = 10; }  /* mdCreator.py:161 */
        break;
    }
    }
}

/* cur_observation (bin3D.py:61-66) -> float32 as VecPyTorch delivers it (acktr/envs.py:191);
 * GT mask; uniform-valid sample. */
static void observe(orc_env *e, int32_t b, float *obs, float *mask, int64_t *sample, int32_t *scratch)
{
    const orc_config *c = &e->cfg;
    const int32_t A = e->A, n = A * e->R;
    const int32_t *plain = e->plain + (size_t)b * A;
    const int32_t *box = e->box + (size_t)b * 3;
    if (obs) {
        float *o = obs + (size_t)b * 4 * A;
        for (int32_t k = 0; k < A; ++k) {
            o[k] = (float)plain[k]; o[A + k] = (float)box[0]; o[2 * A + k] = (float)box[1];
            o[3 * A + k] = (float)box[2];
        }
    }
    if (mask || sample) {
        orc_mask(plain, c->L, c->W, c->H, box[0], box[1], box[2], c->rot, c->mask_variant, scratch);
        if (mask) for (int32_t k = 0; k < n; ++k) mask[(size_t)b * n + k] = (float)scratch[k];
        if (sample) {
            /* uniform over mask == 1 (the fallback mask is all ones, so it is uniform over n then) */
            int32_t nv = 0;
            for (int32_t k = 0; k < n; ++k) nv += scratch[k];
            uint32_t r = mulhi32(orc_draw(c->seed, DOM_SAMPLE, 0, (uint64_t)(c->global_offset + b), e->ctr[b]),
                                 (uint32_t)nv);
            int32_t k = 0;
            for (;; ++k) if (scratch[k]) { if (r == 0) break; --r; }
            sample[b] = k;
        }
    }
}

void orc_reset(orc_env *e, float *obs, float *mask, int64_t *sample)
{
    const int32_t N = e->cfg.n_bins, A = e->A;
    const int fresh = e->fresh;
    int nt = e->cfg.n_threads;
#ifdef _OPENMP
    if (nt <= 0) nt = omp_get_max_threads();
#else
    nt = 1;
#endif
    #pragma omp parallel num_threads(nt)
    {
        int32_t *scratch = (int32_t *)malloc(sizeof(int32_t) * (size_t)A * 2);
        #pragma omp for schedule(dynamic, 128)
        for (int32_t b = 0; b < N; ++b) {
            if (!fresh) e->episode[b] += 1;             /* LoadBoxCreator.reset: index += 1 */
            memset(e->plain + (size_t)b * A, 0, sizeof(int32_t) * (size_t)A);   /* Space(*bin_size) */
            e->counter[b] = 0; e->vol[b] = 0;
            begin_episode(e, b);
            fetch_item(e, b);
            observe(e, b, obs, mask, sample, scratch);
            e->ctr[b] += 1;
        }
        free(scratch);
    }
    e->fresh = 0;
}

void orc_step(orc_env *e, const int64_t *actions, float *obs, float *mask, float *reward,
              uint8_t *done, int32_t *counter, float *ratio, int64_t *sample)
{
    const orc_config *c = &e->cfg;
    const int32_t N = c->n_bins, A = e->A, L = c->L, W = c->W, H = c->H;
    const double lwh = (double)L * (double)W * (double)H;
    uint64_t m_ep = 0, m_cnt = 0, m_vol = 0;
    int nt = c->n_threads;
#ifdef _OPENMP
    if (nt <= 0) nt = omp_get_max_threads();
#else
    nt = 1;
#endif
    #pragma omp parallel num_threads(nt) reduction(+:m_ep,m_cnt,m_vol)
    {
        int32_t *scratch = (int32_t *)malloc(sizeof(int32_t) * (size_t)A * 2);
        #pragma omp for schedule(dynamic, 128)
        for (int32_t b = 0; b < N; ++b) {
            int32_t *plain = e->plain + (size_t)b * A;
            int32_t *box = e->box + (size_t)b * 3;
            /* action decode, bin3D.py:96-105: strict '>' (action == A stays un-rotated and is
             * out of bounds); without rotation an action > A trips the reference's assert, here
             * it is an infeasible placement. */
            int64_t idx = actions[b];
            int flag = 0;
            if (idx > A) { if (c->rot) { idx -= A; flag = 1; } else idx = -1; }
            int32_t x = flag ? box[1] : box[0], y = flag ? box[0] : box[1], z = box[2];  /* space.py:166-172 */
            int32_t new_h = -1;
            if (idx >= 0 && idx < (int64_t)A) {          /* idx_to_position, space.py:153-156 */
                int32_t lx = (int32_t)(idx / W), ly = (int32_t)(idx % W);
                new_h = orc_check_box_v1(plain, L, W, H, x, y, lx, ly, z);            /* space.py:174 */
                if (new_h != -1)
                    for (int32_t r = 0; r < x; ++r)     /* update_height_graph, space.py:36-46 */
                        for (int32_t cc = 0; cc < y; ++cc) plain[(lx + r) * W + ly + cc] = new_h + z;
            }
            if (new_h == -1) {                          /* bin3D.py:108-112 */
                if (reward) reward[b] = 0.0f;
                if (done) done[b] = 1;
                if (counter) counter[b] = e->counter[b];
                if (ratio) ratio[b] = (float)((double)e->vol[b] / lwh);
                m_ep += 1; m_cnt += (uint64_t)e->counter[b]; m_vol += (uint64_t)e->vol[b];
                /* worker auto-reset, shmem_vec_env.py:127-130 -> PackingGame.reset, bin3D.py:55-59 */
                e->episode[b] += 1;
                memset(plain, 0, sizeof(int32_t) * (size_t)A);
                e->counter[b] = 0; e->vol[b] = 0;
                begin_episode(e, b);
            } else {                                    /* bin3D.py:114-127 */
                int32_t v = box[0] * box[1] * box[2];
                e->counter[b] += 1; e->vol[b] += v;
                if (reward) reward[b] = (float)(((double)v / lwh) * 10.0);   /* :44-46, :121 */
                if (done) done[b] = 0;
                if (counter) counter[b] = e->counter[b];
                if (ratio) ratio[b] = (float)((double)e->vol[b] / lwh);       /* space.py:146-151 */
                e->pos[b] += 1;                         /* box_creator.drop_box(); generate_box_size() */
            }
            fetch_item(e, b);
            observe(e, b, obs, mask, sample, scratch);
            e->ctr[b] += 1;
        }
        free(scratch);
    }
    e->met[0] += m_ep; e->met[1] += m_cnt; e->met[2] += m_vol; e->met[3] += (uint64_t)N;
}

/* Re-emit observation and mask of the current state without advancing anything. */
void orc_observe(orc_env *e, float *obs, float *mask)
{
    const int32_t N = e->cfg.n_bins;
    int32_t *scratch = (int32_t *)malloc(sizeof(int32_t) * (size_t)e->A * 2);
    for (int32_t b = 0; b < N; ++b) observe(e, b, obs, mask, NULL, scratch);
    free(scratch);
}

void orc_get_state(const orc_env *e, int32_t *plain, int32_t *box, int32_t *counter, int32_t *vol)
{
    const size_t N = (size_t)e->cfg.n_bins;
    if (plain) memcpy(plain, e->plain, sizeof(int32_t) * N * e->A);
    if (box) memcpy(box, e->box, sizeof(int32_t) * N * 3);
    if (counter) memcpy(counter, e->counter, sizeof(int32_t) * N);
    if (vol) memcpy(vol, e->vol, sizeof(int32_t) * N);
}

void orc_set_state(orc_env *e, const int32_t *plain, const int32_t *box, const int32_t *counter,
                   const int32_t *vol)
{
    const size_t N = (size_t)e->cfg.n_bins;
    if (plain) memcpy(e->plain, plain, sizeof(int32_t) * N * e->A);
    if (box) memcpy(e->box, box, sizeof(int32_t) * N * 3);
    if (counter) memcpy(e->counter, counter, sizeof(int32_t) * N);
    if (vol) memcpy(e->vol, vol, sizeof(int32_t) * N);
}

void orc_metrics(const orc_env *e, uint64_t out[4])
{
    for (int i = 0; i < 4; ++i) out[i] = e->met[i];
}
