// bpp_gen.cu -- on-device episode generators for --item-seq cut1 / cut2 (sm_100a).
//
// cut1: envs/bpp0/cutCreator.py:32-128 -- cut the bin recursively into pieces whose sides lie in the item range,
//       then emit them in a random *support-ready* order (a piece is ready when the shadow heightmap under its whole
//       footprint equals its lz); optional random x/y swap when rotation is enabled.
// cut2: envs/bpp0/mdCreator.py:59-166 -- split any piece with a side > high at a random position, rejecting cuts
//       that leave a side < low (the piece is retried in a later sweep; the sweep order reproduces the reference's
//       remove-while-iterating traversal), then stable-sort by the z of the bottom face.
// The emission order does not depend on the agent, so a whole episode (10^3: ~26 items, 20^3: ~210) is produced at
// once into the bin's spare episode buffer while the current episode is being played; the step kernel swaps buffers
// at auto-reset.  One warp per episode: the parts that are sequential in the order of the random draws run on one
// lane (or uniformly on all), list maintenance and footprint work use the 32 lanes.  Philox draws come from the bin's
// (domain 2, aux = episode) stream in exactly the order of the CPU oracle, so device and oracle agree bit for bit.
#include "bpp_common.cuh"
#include "bpp_launch.h"

namespace bpp {

namespace {

// A cuboid piece of the bin: size (x, y, z) and position (lx, ly, lz), one byte each, in one 64-bit word.
typedef unsigned long long Piece;
__device__ __forceinline__ Piece mk_piece(int x, int y, int z, int lx, int ly, int lz)
{
    return (Piece)(unsigned)(x | (y << 8) | (z << 16) | (lx << 24)) | ((Piece)(unsigned)(ly | (lz << 8)) << 32);
}
__device__ __forceinline__ int px(Piece q) { return (int)(q & 0xff); }
__device__ __forceinline__ int py(Piece q) { return (int)((q >> 8) & 0xff); }
__device__ __forceinline__ int pz(Piece q) { return (int)((q >> 16) & 0xff); }
__device__ __forceinline__ int plx(Piece q) { return (int)((q >> 24) & 0xff); }
__device__ __forceinline__ int ply(Piece q) { return (int)((q >> 32) & 0xff); }
__device__ __forceinline__ int plz(Piece q) { return (int)((q >> 40) & 0xff); }

// The bin's generator stream (domain 2, aux = episode); every lane of the warp keeps the same state.
struct GenRng {
    uint32_t seed_lo, seed_hi, aux, k, blk_idx;
    unsigned long long gid;
    uint4 blk;
    __device__ __forceinline__ uint32_t next()
    {
        const uint32_t b = k >> 2;
        if (b != blk_idx) {
            blk = philox4x32_10(make_uint4(b, ((uint32_t)kDomGen << 28) | (aux & 0x0FFFFFFFu), (uint32_t)gid, (uint32_t)(gid >> 32)),
                                make_uint2(seed_lo, seed_hi));
            blk_idx = b;
        }
        const uint32_t w = k & 3u;
        ++k;
        return w == 0 ? blk.x : (w == 1 ? blk.y : (w == 2 ? blk.z : blk.w));
    }
};

__device__ __forceinline__ uint32_t pack_item(int x, int y, int z) { return (uint32_t)x | ((uint32_t)y << 8) | ((uint32_t)z << 16); }

// list.pop(idx) / list.remove(item at idx) on an array in (shared) memory, by the whole warp
template <typename T>
__device__ __forceinline__ void warp_erase(T *a, int idx, int n, int lane)
{
    for (int k0 = idx; k0 + 1 < n; k0 += 32) {
        const int k = k0 + lane;
        T v = T();
        if (k + 1 < n) v = a[k + 1];
        __syncwarp();
        if (k + 1 < n) a[k] = v;
        __syncwarp();
    }
}

// cut1 ordering phase for episodes of up to 64 pieces, held in registers: lane l owns pieces l and 32 + l.
//
// The reference keeps a shadow heightmap and calls a waiting piece ready when the map equals its lz on its whole
// footprint (_add_candidate, cutCreator.py:98-107).  Pieces are emitted bottom-up and partition the bin exactly, so the
// map equals lz on a cell exactly when the piece directly below it there -- whose top face is at lz -- has been
// emitted: a piece is ready once the pieces whose top faces tile its bottom face have all been emitted.  Every piece
// therefore carries the area of its bottom face still uncovered; emitting a piece whose top is at a waiting piece's lz
// subtracts their overlap, and 0 means ready.  No shadow map, no footprint loops.  The candidate list is a position per
// piece: candidates.pop(idx) selects the piece at position idx and moves the later ones down; pieces that become ready are
// appended in meta_list order, i.e. by piece index (slot-major, lane-minor) -- the reference's orders exactly.
__device__ int order_cut1_registers(const EnvParams &p, GenRng &g, uint32_t *out, const Piece *leaf, int n_leaf, int cap, int lane)
{
    const unsigned lt = (1u << lane) - 1u;
    // per slot: the piece's item word (x | y << 8 | z << 16), its footprint bounds, the z of its bottom and top face
    uint32_t item[2];
    int x0[2], x1[2], y0[2], y1[2], zb[2], zt[2];
    int rem[2], pos[2];                                  // uncovered bottom area (waiting pieces), candidate-list position
    unsigned c[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int i = s * 32 + lane;
        const bool valid = i < n_leaf;
        const Piece q = valid ? leaf[i] : 0;
        item[s] = (uint32_t)q & 0x00ffffffu;
        x0[s] = plx(q); x1[s] = x0[s] + px(q); y0[s] = ply(q); y1[s] = y0[s] + py(q); zb[s] = plz(q); zt[s] = zb[s] + pz(q);
        rem[s] = (valid && zb[s] > 0) ? px(q) * py(q) : 0;
        c[s] = __ballot_sync(kFull, valid && zb[s] == 0);               // first scan: shadow == 0, ready <=> lz == 0
    }
    pos[0] = (c[0] >> lane) & 1u ? __popc(c[0] & lt) : -1;
    pos[1] = (c[1] >> lane) & 1u ? __popc(c[0]) + __popc(c[1] & lt) : -1;
    int n_cand = __popc(c[0]) + __popc(c[1]), n_out = 0;
    while (n_cand > 0) {                                                  // :113-115: no candidate -> sentinel from now on
        const int idx = (int)__umulhi(g.next(), (uint32_t)n_cand);        // random.randint(0, len(candidates) - 1), :116
        const bool mine0 = pos[0] == idx, mine1 = pos[1] == idx;
        const int src = __ffs(__ballot_sync(kFull, mine0 || mine1)) - 1;
        // candidates.pop(idx): the chosen piece's item word, footprint (4 x 8 bits) and top face reach every lane
        const uint32_t e_item = __shfl_sync(kFull, mine0 ? item[0] : item[1], src);
        const uint32_t e_box = __shfl_sync(kFull, mine0 ? (uint32_t)(x0[0] | (x1[0] << 8) | (y0[0] << 16) | (y1[0] << 24))
                                                        : (uint32_t)(x0[1] | (x1[1] << 8) | (y0[1] << 16) | (y1[1] << 24)), src);
        const int top = __shfl_sync(kFull, mine0 ? zt[0] : zt[1], src);
#pragma unroll
        for (int s = 0; s < 2; ++s) pos[s] = pos[s] == idx ? -1 : (pos[s] > idx ? pos[s] - 1 : pos[s]);
        --n_cand;
        uint32_t emit = e_item;
        if (p.rot) {                                                      // :120-126
            const uint32_t coin = g.next();
            if (coin >= 0x80000000u) emit = (e_item & 0x00ff0000u) | ((e_item & 0xffu) << 8) | ((e_item >> 8) & 0xffu);
        }
        if (n_out < cap) {
            if (lane == 0) out[n_out] = emit;
            ++n_out;
        }
        // _update + _add_candidate, :109-110, :98-107
        const int ex0 = e_box & 0xff, ex1 = (e_box >> 8) & 0xff, ey0 = (e_box >> 16) & 0xff, ey1 = e_box >> 24;
        bool fresh[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            fresh[s] = false;
            if (rem[s] > 0 && zb[s] == top) {
                const int wx = min(ex1, x1[s]) - max(ex0, x0[s]), wy = min(ey1, y1[s]) - max(ey0, y0[s]);
                if (wx > 0 && wy > 0) {
                    rem[s] -= wx * wy;
                    fresh[s] = rem[s] == 0;
                }
            }
        }
        const unsigned f0 = __ballot_sync(kFull, fresh[0]), f1 = __ballot_sync(kFull, fresh[1]);
        if (fresh[0]) pos[0] = n_cand + __popc(f0 & lt);                  // candidates.append, in meta_list order
        if (fresh[1]) pos[1] = n_cand + __popc(f0) + __popc(f1 & lt);
        n_cand += __popc(f0) + __popc(f1);
    }
    return n_out;
}

// cut1, one warp per episode.  Phase 1, _cut_box (cutCreator.py:79-95), sweeps the piece list level by level, and
// within one sweep the pieces are independent of each other except for WHICH two draws a piece that must be split
// consumes (2 per split, in list order) and WHERE its one or two successors land in the next list -- two prefix sums.
// So a sweep runs with one lane per piece: ballot of the lanes that split, draw index and output slot from the
// ballots, one Philox evaluation per splitting lane (its two draws share a 4-word block because the index is even).
// The draws are consumed in exactly the reference's order (tests/test_generator_identity.py).  Pieces are disjoint
// with every side >= lo, so neither list ever holds more than cap = (L/lo)(W/lo)(H/lo) of them.
// Phase 2 (support-ready ordering) uses the warp per piece: lanes test the waiting pieces against the piece just
// emitted, check / raise footprints of the shadow heightmap cell-parallel, and compact the lists with ballots.
// The lists keep exactly the reference's order (scan order of meta_list, appends to candidates, pop(idx)).
__device__ int gen_cut1(const EnvParams &p, GenRng &g, uint32_t *out, Piece *leaf, uint16_t *lists, uint8_t *shadow,
                        Piece *next, int cap, int lane)
{
    const int L = p.L, W = p.W, H = p.H;
    const unsigned lt = (1u << lane) - 1u;
    int n_leaf = 1;
    if (lane == 0) leaf[0] = mk_piece(L, W, H, 0, 0, 0);
    __syncwarp();
    for (;;) {
        int n_next = 0;
        bool again = false;
        for (int c0 = 0; c0 < n_leaf; c0 += 32) {                         // for box in self.meta_list, :84
            const int i = c0 + lane;
            const bool valid = i < n_leaf;
            const Piece b = valid ? leaf[i] : 0;
            // _check_box, :54-58: bit a of `bad` = side a lies outside [lo, hi]
            const unsigned bad = (unsigned)(px(b) < p.lo[0] || px(b) > p.hi[0]) | ((unsigned)(py(b) < p.lo[1] || py(b) > p.hi[1]) << 1) |
                                 ((unsigned)(pz(b) < p.lo[2] || pz(b) > p.hi[2]) << 2);
            const bool split = valid && bad != 0u;
            const unsigned sb = __ballot_sync(kFull, split), vb = __ballot_sync(kFull, valid);
            const int slot = n_next + __popc(vb & lt) + __popc(sb & lt);
            if (split) {
                const uint32_t k = g.k + 2u * (uint32_t)__popc(sb & lt);  // even: both draws come from one block
                const uint4 blk = philox4x32_10(make_uint4(k >> 2, ((uint32_t)kDomGen << 28) | (g.aux & 0x0FFFFFFFu),
                                                           (uint32_t)g.gid, (uint32_t)(g.gid >> 32)),
                                                make_uint2(g.seed_lo, g.seed_hi));
                const uint32_t u0 = (k & 2u) ? blk.z : blk.x, u1 = (k & 2u) ? blk.w : blk.y;
                const int r = (int)__umulhi(u0, (uint32_t)__popc(bad));   // random.choice(df_list), :68: the r-th offending axis
                const unsigned rest = r == 0 ? bad : (r == 1 ? (bad & (bad - 1u)) : 4u);
                const int df = __ffs(rest) - 1;
                const int side = df == 0 ? px(b) : (df == 1 ? py(b) : pz(b));
                const int lo = df == 0 ? p.lo[0] : (df == 1 ? p.lo[1] : p.lo[2]);
                int span = side - 2 * lo + 1;                             // randint(lo, side - lo), :69-76
                if (span < 1) span = 1;
                const int cut = lo + (int)__umulhi(u1, (uint32_t)span);
                Piece p1, p2;                                             // MetaBox.split, :16-26
                if (df == 0) { p1 = mk_piece(cut, py(b), pz(b), plx(b), ply(b), plz(b)); p2 = mk_piece(px(b) - cut, py(b), pz(b), plx(b) + cut, ply(b), plz(b)); }
                else if (df == 1) { p1 = mk_piece(px(b), cut, pz(b), plx(b), ply(b), plz(b)); p2 = mk_piece(px(b), py(b) - cut, pz(b), plx(b), ply(b) + cut, plz(b)); }
                else { p1 = mk_piece(px(b), py(b), cut, plx(b), ply(b), plz(b)); p2 = mk_piece(px(b), py(b), pz(b) - cut, plx(b), ply(b), plz(b) + cut); }
                if (slot + 2 <= cap) { next[slot] = p1; next[slot + 1] = p2; }
            } else if (valid && slot < cap) {
                next[slot] = b;
            }
            n_next += __popc(vb) + __popc(sb);
            g.k += 2u * (uint32_t)__popc(sb);
            again = again || sb != 0u;
        }
        __syncwarp();
        Piece *t = leaf; leaf = next; next = t;                           // self.meta_list = copy.deepcopy(new_list), :94
        n_leaf = min(n_next, cap);
        if (!again) break;
    }
    g.blk_idx = 0xffffffffu;

    // ordering phase: generate_box_size / _add_candidate / _update, cutCreator.py:98-128
    if (n_leaf <= 64) return order_cut1_registers(p, g, out, leaf, n_leaf, cap, lane);
    uint16_t *meta = lists, *cand = lists + cap;
    for (int i = lane; i < n_leaf; i += 32) meta[i] = (uint16_t)i;
    for (int k = lane; k < L * W; k += 32) shadow[k] = 0;
    __syncwarp();
    int n_meta = n_leaf, n_cand = 0, n_out = 0;
    bool first = true;
    Piece last = 0;
    for (;;) {
        // A piece is ready when the shadow heightmap equals its lz on its whole footprint.  The shadow only changes
        // under the piece just emitted, so after the first scan (shadow == 0: ready <=> lz == 0) a waiting piece can
        // only become ready if it overlaps the emitted footprint and sits exactly on the new height there.
        int keep = 0;
        for (int m0 = 0; m0 < n_meta; m0 += 32) {
            const int m = m0 + lane;
            const bool valid = m < n_meta;
            const int pid = valid ? (int)meta[m] : 0;
            const Piece q = valid ? leaf[pid] : 0;
            bool ready;
            if (first) {
                ready = valid && plz(q) == 0;
            } else {
                const bool pre = valid && plz(q) == plz(last) + pz(last) && plx(q) < plx(last) + px(last) &&
                                 plx(last) < plx(q) + px(q) && ply(q) < ply(last) + py(last) && ply(last) < ply(q) + py(q);
                ready = false;
                unsigned pm = __ballot_sync(kFull, pre);
                while (pm) {                                              // footprint check, lanes over the cells
                    const int src = __ffs(pm) - 1;
                    pm &= pm - 1;
                    const Piece qq = __shfl_sync(kFull, q, src);
                    const int y = py(qq), a = px(qq) * y, base = plx(qq) * W + ply(qq), lz = plz(qq);
                    const uint32_t rcp = rcp16((uint32_t)y);
                    bool ok = true;
                    for (int c = lane; c < a; c += 32) {
                        const int r = (int)(((uint32_t)c * rcp) >> 16);
                        if (shadow[base + r * W + (c - r * y)] != lz) ok = false;
                    }
                    const bool all_ok = __all_sync(kFull, ok);
                    if (lane == src) ready = all_ok;
                }
            }
            const unsigned rb = __ballot_sync(kFull, ready), kb = __ballot_sync(kFull, valid && !ready);
            __syncwarp();                                                 // this chunk of meta has been read by every lane
            if (ready) cand[n_cand + __popc(rb & lt)] = (uint16_t)pid;    // candidates.append, in scan order
            if (valid && !ready) meta[keep + __popc(kb & lt)] = (uint16_t)pid;
            n_cand += __popc(rb);
            keep += __popc(kb);
        }
        n_meta = keep;
        first = false;
        __syncwarp();
        if (n_cand == 0) break;                                           // :113-115 -> sentinel from now on
        const int idx = (int)__umulhi(g.next(), (uint32_t)n_cand);        // :116
        const Piece q = leaf[cand[idx]];
        __syncwarp();
        warp_erase(cand, idx, n_cand, lane);                              // candidates.pop(idx)
        --n_cand;
        int ox = px(q), oy = py(q);
        if (p.rot) {                                                      // :120-126
            const uint32_t coin = g.next();
            if (coin >= 0x80000000u) { ox = py(q); oy = px(q); }
        }
        if (n_out < cap) {
            if (lane == 0) out[n_out] = pack_item(ox, oy, pz(q));
            ++n_out;
        }
        {                                                                 // _update, :109-110
            const int y = py(q), a = px(q) * y, base = plx(q) * W + ply(q);
            const uint32_t rcp = rcp16((uint32_t)y);
            for (int c = lane; c < a; c += 32) {
                const int r = (int)(((uint32_t)c * rcp) >> 16);
                shadow[base + r * W + (c - r * y)] += (uint8_t)pz(q);
            }
        }
        last = q;
        __syncwarp();
    }
    return n_out;
}

// cut2, one warp per episode.  gen_benchmark's sweep (mdCreator.py:117-135) is sequential in the order of the random
// draws, so every lane follows it with the same state; lane 0 alone reads and writes the two lists (the current piece
// reaches the others by shuffle), so the sweep needs no barrier.  What the 32 lanes do share out:
//   * the draws: 128 at a time, one Philox block per lane, parked in shared memory (a visit takes two);
//   * invalid_box.remove(box) leaves a tombstone -- within a sweep the elements ahead of the iterator are all live, so
//     "the iterator skips the element that slid into the vacated slot" is i += 2 -- and the list is compacted between
//     sweeps with one ballot per 32 entries;
//   * the stable sort by low_bound (:137-138): one ballot per level and 32 entries.
// A piece is two words here: sizes x | y << 8 | z << 16 and position lx | ly << 8 | lz << 16 (a cut along axis a then
// is an add / subtract at bit 8a, and a field is one extract).  `inv`: cap + 64 pieces (size word 0 = tombstone);
// `val`: cap words x | y << 8 | z << 16 | lz << 24; `draws`: 128 words.
__device__ int gen_cut2(const EnvParams &p, GenRng &g, uint32_t *out, Piece *inv_, uint32_t *val, uint32_t *draws, int cap, int lane)
{
    uint2 *inv = reinterpret_cast<uint2 *>(inv_);
    const int lo = p.lo[0], hi = p.hi[0];
    const unsigned lt = (1u << lane) - 1u;
    int n_inv = 1, n_val = 0;
    uint32_t k = 0;                                                       // draws consumed (even)
    if (lane == 0) inv[0] = make_uint2((uint32_t)p.L | ((uint32_t)p.W << 8) | ((uint32_t)p.H << 16), 0u);
    while (n_inv > 0) {                                                   // gen_benchmark, mdCreator.py:117-135
        int i = 0, n_phys = n_inv;
        while (i < n_phys) {
            if ((k & 127u) == 0u) {                                       // next 128 draws of the (domain 2, episode) stream
                __syncwarp();
                const uint4 blk = philox4x32_10(make_uint4((k >> 2) + (uint32_t)lane, ((uint32_t)kDomGen << 28) | (g.aux & 0x0FFFFFFFu),
                                                           (uint32_t)g.gid, (uint32_t)(g.gid >> 32)),
                                                make_uint2(g.seed_lo, g.seed_hi));
                reinterpret_cast<uint4 *>(draws)[lane] = blk;
                __syncwarp();
            }
            uint2 b = make_uint2(0u, 0u);
            if (lane == 0) b = inv[i];
            const uint32_t d = __shfl_sync(kFull, b.x, 0), o = __shfl_sync(kFull, b.y, 0);
            const int x = d & 0xff, y = (d >> 8) & 0xff, z = (d >> 16) & 0xff;
            const unsigned fl = (unsigned)(x > hi) | ((unsigned)(y > hi) << 1) | ((unsigned)(z > hi) << 2);   // benchmark_split, :60-66
            if (fl == 0u) {                                               // unreachable for validated ranges (no draw is taken)
                if (lane == 0) inv[i].x = 0u;
                if (n_val < cap) { if (lane == 0) val[n_val] = d | ((o >> 16) << 24); ++n_val; }
                i += 2;
                continue;
            }
            const uint32_t u0 = draws[k & 127u], u1 = draws[(k + 1u) & 127u];
            k += 2u;
            const int r = (int)__umulhi(u0, (uint32_t)__popc(fl));        // random.choice(flags), :68: the r-th oversize axis
            const unsigned rest = r == 0 ? fl : (r == 1 ? (fl & (fl - 1u)) : 4u);
            const int ax = __ffs(rest) - 1, sh = 8 * ax;
            const int side = (int)((d >> sh) & 0xffu);
            const int cut = 1 + (int)__umulhi(u1, (uint32_t)side);        // randint(1, side)
            if (cut < lo || side - cut < lo) { ++i; continue; }            // rejected, retried in the next sweep
            // :70-99: along x / y the first part has size `cut` and the second starts `cut` further; along z the first part
            // keeps the bottom with size z - cut and the second (size cut) sits on top of it
            const int a = ax == 2 ? side - cut : cut;
            const uint32_t d1 = d - ((uint32_t)(side - a) << sh), d2 = d - ((uint32_t)a << sh), o2 = o + ((uint32_t)a << sh);
            if (lane == 0) inv[i].x = 0u;                                 // invalid_box.remove(box): tombstone
            // is_valid, :110-115: all three sizes within [lo, hi]; both new sizes are >= lo by the test above
            const unsigned in_range = (unsigned)(x >= lo && x <= hi) | ((unsigned)(y >= lo && y <= hi) << 1) | ((unsigned)(z >= lo && z <= hi) << 2);
            const bool others = (in_range | (1u << ax)) == 7u;
            const bool ok1 = others && a <= hi, ok2 = others && side - a <= hi;
            if (ok1) { if (n_val < cap) { if (lane == 0) val[n_val] = d1 | ((o >> 16) << 24); ++n_val; } }
            else if (n_phys < cap + 64) { if (lane == 0) inv[n_phys] = make_uint2(d1, o); ++n_phys; }
            if (ok2) { if (n_val < cap) { if (lane == 0) val[n_val] = d2 | ((o2 >> 16) << 24); ++n_val; } }
            else if (n_phys < cap + 64) { if (lane == 0) inv[n_phys] = make_uint2(d2, o2); ++n_phys; }
            i += 2;                                                       // the iterator skips the element that slid into slot i
        }
        int keep = 0;                                                     // drop the tombstones, 32 entries per ballot
        __syncwarp();
        for (int j0 = 0; j0 < n_phys; j0 += 32) {
            const int j = j0 + lane;
            const uint2 q = j < n_phys ? inv[j] : make_uint2(0u, 0u);
            const unsigned live = __ballot_sync(kFull, q.x != 0u);
            __syncwarp();                                                 // every lane has read its entry of this chunk
            if (q.x != 0u) inv[keep + __popc(live & lt)] = q;
            keep += __popc(live);
        }
        __syncwarp();
        n_inv = keep;
    }
    g.k = k;
    __syncwarp();
    int n_out = 0;                                                        // depart_box: stable sort by low_bound, :137-138
    for (int level = 0; level <= p.H && n_out < cap; ++level)
        for (int v0 = 0; v0 < n_val; v0 += 32) {
            const int v = v0 + lane;
            const uint32_t q = v < n_val ? val[v] : 0u;
            const bool hit = v < n_val && (int)(q >> 24) == level;
            const unsigned hb = __ballot_sync(kFull, hit);
            const int pos = n_out + __popc(hb & lt);
            if (hit && pos < cap) out[pos] = q & 0x00ffffffu;
            n_out = min(cap, n_out + __popc(hb));
        }
    return n_out;
}

}  // namespace

// One warp per episode.  mode 0 (init): every bin gets episode e0 into buffer 0 and e0 + 1 into buffer 1;
// mode 1 (refill): the bins on the compacted refill list (appended by the step kernel at auto-reset) get episode
// sc.episode + 1 into the buffer they are not playing.  p is the whole env (ep_bin0 == 0).
__global__ void __launch_bounds__(128) bpp_gen_kernel(const EnvParams p, const int mode, const uint32_t e0, const int use_smem)
{
    BPP_DYNAMIC_SMEM(gen_smem);
    const long long n_work = mode == 0 ? (long long)p.N : (long long)*p.gen_count;
    const int lane = threadIdx.x & 31;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long wi = warp_id; wi < n_work; wi += n_warps) {
        const long long bin = mode == 0 ? wi : (long long)p.gen_list[wi];
        const BinScalars s = p.sc[bin];
        const int cur = mode == 0 ? 0 : (int)(s.traj & 1u);
        // work space: shared memory when the launch could afford it (these are chains of dependent accesses), else the
        // bin's global scratch
        uint8_t *scratch = use_smem ? gen_smem + (size_t)(threadIdx.x >> 5) * (size_t)p.gen_scratch_stride
                                    : p.gen_scratch + (size_t)bin * (size_t)p.gen_scratch_stride;
        for (int which = (mode == 0 ? 0 : 1); which < 2; ++which) {
            const int buf = mode == 0 ? which : (cur ^ 1);
            const uint32_t episode = mode == 0 ? e0 + (uint32_t)which : s.episode + 1u;
            GenRng g;
            g.seed_lo = p.seed_lo; g.seed_hi = p.seed_hi; g.aux = episode; g.k = 0; g.blk_idx = 0xffffffffu;
            g.gid = p.global_offset + (unsigned long long)bin;
            uint32_t *out = p.ep_items + ((size_t)buf * p.ep_buf_bins + (size_t)bin) * p.ep_cap;
            int n;
            if (p.gen_kind == 2) {
                Piece *leaf = reinterpret_cast<Piece *>(scratch);
                const size_t o1 = ((size_t)p.ep_cap * sizeof(Piece) + 15) & ~(size_t)15;
                const size_t o2 = 2 * o1;
                const size_t o3 = o2 + (((size_t)p.ep_cap * 4 + 15) & ~(size_t)15);
                n = gen_cut1(p, g, out, leaf, reinterpret_cast<uint16_t *>(scratch + o2), scratch + o3,
                             reinterpret_cast<Piece *>(scratch + o1), p.ep_cap, lane);
            } else {
                // invalid list | valid list | draws, each 16-byte aligned (the draws are stored 128 bits at a time)
                Piece *inv = reinterpret_cast<Piece *>(scratch);
                const size_t o1 = ((size_t)(p.ep_cap + 64) * sizeof(Piece) + 15) & ~(size_t)15;
                uint32_t *val = reinterpret_cast<uint32_t *>(scratch + o1);
                uint32_t *draws = reinterpret_cast<uint32_t *>(scratch + o1 + (((size_t)p.ep_cap * 4 + 15) & ~(size_t)15));
                n = gen_cut2(p, g, out, inv, val, draws, p.ep_cap, lane);
            }
            if (lane == 0) p.ep_len2[(size_t)buf * p.ep_buf_bins + (size_t)bin] = n;
            __syncwarp();
        }
    }
}

cudaError_t launch_gen(const EnvParams &p, int mode, uint32_t e0, int num_sms, cudaStream_t st)
{
    long long want = ((long long)p.N + 3) / 4;            // 4 warps per block, one episode per warp
    const long long cap = (long long)num_sms * 16;
    if (mode == 1 && want > cap) want = cap;              // the list length is only known on the device
    const size_t smem = (size_t)4 * (size_t)p.gen_scratch_stride;
    const int use_smem = smem <= 200 * 1024;
    if (use_smem && smem > 48 * 1024) {
        cudaError_t ea = cudaFuncSetAttribute(bpp_gen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ea != cudaSuccess) return ea;
    }
    BPP_LAUNCH((bpp_gen_kernel), (unsigned)want, 128, use_smem ? smem : 0, st, p, mode, e0, use_smem);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || mode == 0) return e;
    return cudaMemsetAsync(p.gen_count, 0, sizeof(int32_t), st);     // list consumed
}

size_t gen_scratch_bytes_per_bin(int gen_kind, int cap)
{
    if (gen_kind == 2)      // piece list | piece list of the next sweep | meta + candidate lists | shadow heightmap (<= 1024)
        return 2 * (((size_t)cap * sizeof(Piece) + 15) & ~(size_t)15) + (((size_t)cap * 4 + 15) & ~(size_t)15) + 1024;
    // cut2: invalid list (cap + 64 pieces) | valid list (cap words) | 128 draws
    return (((size_t)(cap + 64) * sizeof(Piece) + 15) & ~(size_t)15) + (((size_t)cap * 4 + 15) & ~(size_t)15) + 512;
}

}  // namespace bpp
