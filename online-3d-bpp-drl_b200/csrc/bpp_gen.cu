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
// at auto-reset.  One thread per episode (the algorithms are sequential); Philox draws come from the bin's
// (domain 2, aux = episode) stream in exactly the order of the CPU oracle, so device and oracle agree bit for bit.
#include "bpp_common.cuh"
#include "bpp_launch.h"

namespace bpp {

namespace {

struct Piece {
    uint8_t x, y, z, lx, ly, lz;
};

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

constexpr int kGenLanes = 4;     // episodes generated per warp (see bpp_gen_kernel)
constexpr int kStack = 96;      // DFS stack of the cut1 split tree: depth <= (L+W+H)/min(lo) <= 96 for L,W<=32,H<=255,lo>=... checked on the host

__device__ __forceinline__ uint32_t pack_item(int x, int y, int z) { return (uint32_t)x | ((uint32_t)y << 8) | ((uint32_t)z << 16); }

// cut1 into `out` (packed items), using `leaf` (cap pieces) and `lists` (2*cap uint16) as scratch.  Returns the length.
__device__ int gen_cut1(const EnvParams &p, GenRng &g, uint32_t *out, Piece *leaf, uint16_t *lists, uint8_t *shadow,
                        Piece *stack, int cap)
{
    const int L = p.L, W = p.W, H = p.H;
    int sp = 0, n_leaf = 0;
    stack[sp++] = Piece{(uint8_t)L, (uint8_t)W, (uint8_t)H, 0, 0, 0};
    while (sp > 0) {
        const Piece b = stack[--sp];
        const int dim[3] = {b.x, b.y, b.z};
        int bad[3], nbad = 0;
        for (int a = 0; a < 3; ++a)                                   // _check_box, cutCreator.py:54-58
            if (dim[a] < p.lo[a] || dim[a] > p.hi[a]) bad[nbad++] = a;
        if (nbad == 0) { if (n_leaf < cap) leaf[n_leaf++] = b; continue; }
        const int df = bad[__umulhi(g.next(), (uint32_t)nbad)];          // random.choice, :68
        const int lo = p.lo[df];
        int span = dim[df] - 2 * lo + 1;                              // randint(lo, dim - lo), :69-76
        if (span < 1) span = 1;
        const int cut = lo + (int)__umulhi(g.next(), (uint32_t)span);
        Piece p1 = b, p2 = b;                                         // MetaBox.split, :16-26
        if (df == 0) { p1.x = (uint8_t)cut; p2.x = (uint8_t)(b.x - cut); p2.lx = (uint8_t)(b.lx + cut); }
        else if (df == 1) { p1.y = (uint8_t)cut; p2.y = (uint8_t)(b.y - cut); p2.ly = (uint8_t)(b.ly + cut); }
        else { p1.z = (uint8_t)cut; p2.z = (uint8_t)(b.z - cut); p2.lz = (uint8_t)(b.lz + cut); }
        if (sp + 2 <= kStack) { stack[sp++] = p2; stack[sp++] = p1; }  // p1 first => in-order leaves
    }
    // ordering phase: generate_box_size / _add_candidate / _update, cutCreator.py:98-128
    for (int k = 0; k < L * W; ++k) shadow[k] = 0;
    uint16_t *meta = lists, *cand = lists + cap;
    int n_meta = n_leaf, n_cand = 0, n_out = 0;
    for (int i = 0; i < n_leaf; ++i) meta[i] = (uint16_t)i;
    // A piece is ready when the shadow heightmap equals its lz on its whole footprint.  The shadow only changes under
    // the piece just emitted, so after the first scan (shadow == 0: ready <=> lz == 0) a waiting piece can only
    // become ready if it overlaps the emitted footprint and sits exactly on the new height there; everything else
    // is skipped with six comparisons.  The scan order (and therefore the candidate order) is the reference's.
    Piece last = Piece{0, 0, 0, 0, 0, 0};
    bool first = true;
    for (;;) {
        int keep = 0;
        for (int m = 0; m < n_meta; ++m) {
            const Piece q = leaf[meta[m]];
            bool ready;
            if (first) {
                ready = q.lz == 0;
            } else {
                ready = q.lz == last.lz + last.z && q.lx < last.lx + last.x && last.lx < q.lx + q.x &&
                        q.ly < last.ly + last.y && last.ly < q.ly + q.y;
                for (int r = 0; r < q.x && ready; ++r)
                    for (int c = 0; c < q.y; ++c)
                        if (shadow[(q.lx + r) * W + q.ly + c] != q.lz) { ready = false; break; }
            }
            if (ready) cand[n_cand++] = meta[m]; else meta[keep++] = meta[m];
        }
        first = false;
        n_meta = keep;
        if (n_cand == 0) break;                                       // :113-115 -> sentinel from now on
        const int idx = (int)__umulhi(g.next(), (uint32_t)n_cand);    // :116
        const Piece q = leaf[cand[idx]];
        for (int m = idx; m + 1 < n_cand; ++m) cand[m] = cand[m + 1];  // list.pop(idx)
        --n_cand;
        int ox = q.x, oy = q.y;
        if (p.rot) {                                                  // :120-126
            const uint32_t coin = g.next();
            if (coin >= 0x80000000u) { ox = q.y; oy = q.x; }
        }
        if (n_out < cap) out[n_out++] = pack_item(ox, oy, q.z);
        for (int r = 0; r < q.x; ++r)                                 // _update, :109-110
            for (int c = 0; c < q.y; ++c) shadow[(q.lx + r) * W + q.ly + c] += q.z;
        last = q;
    }
    return n_out;
}

// cut2 into `out`, using `inv` / `val` (cap + 64 pieces each) as scratch.  Returns the length.
__device__ int gen_cut2(const EnvParams &p, GenRng &g, uint32_t *out, Piece *inv, Piece *val, int cap)
{
    const int lo = p.lo[0], hi = p.hi[0];
    int n_inv = 0, n_val = 0;
    inv[n_inv++] = Piece{(uint8_t)p.L, (uint8_t)p.W, (uint8_t)p.H, 0, 0, 0};
    while (n_inv > 0) {                                               // gen_benchmark, mdCreator.py:117-135
        int i = 0;
        while (i < n_inv) {
            const Piece b = inv[i];
            const int dim[3] = {b.x, b.y, b.z};
            int fl[3], nf = 0;
            for (int a = 0; a < 3; ++a) if (dim[a] > hi) fl[nf++] = a;   // benchmark_split, :60-66
            if (nf == 0) {                                            // unreachable for validated ranges
                for (int m = i; m + 1 < n_inv; ++m) inv[m] = inv[m + 1];
                --n_inv; if (n_val < cap) val[n_val++] = b; ++i; continue;
            }
            const int ax = fl[__umulhi(g.next(), (uint32_t)nf)];        // :68
            const int cut = 1 + (int)__umulhi(g.next(), (uint32_t)dim[ax]);   // randint(1, dim)
            if (cut < lo || dim[ax] - cut < lo) { ++i; continue; }     // rejected, retried in the next sweep
            Piece p1 = b, p2 = b;
            if (ax == 0) { p1.x = (uint8_t)cut; p2.x = (uint8_t)(b.x - cut); p2.lx = (uint8_t)(b.lx + cut); }          // :70-79
            else if (ax == 1) { p1.y = (uint8_t)cut; p2.y = (uint8_t)(b.y - cut); p2.ly = (uint8_t)(b.ly + cut); }     // :80-89
            else { p1.z = (uint8_t)(b.z - cut); p2.z = (uint8_t)cut; p2.lz = (uint8_t)(b.lz + b.z - cut); }             // :90-99
            for (int m = i; m + 1 < n_inv; ++m) inv[m] = inv[m + 1];  // invalid_box.remove(box)
            --n_inv;
            const Piece sub[2] = {p1, p2};
            for (int s = 0; s < 2; ++s) {                             // is_valid, :110-115
                const Piece q = sub[s];
                const bool ok = q.x >= lo && q.x <= hi && q.y >= lo && q.y <= hi && q.z >= lo && q.z <= hi;
                if (ok) { if (n_val < cap) val[n_val++] = q; }
                else if (n_inv < cap + 64) inv[n_inv++] = q;
            }
            ++i;                                                      // the iterator skips the element that slid into slot i
        }
    }
    int n_out = 0;                                                    // depart_box: stable sort by low_bound, :137-138
    for (int level = 0; level <= p.H && n_out < cap; ++level)
        for (int v = 0; v < n_val && n_out < cap; ++v)
            if (val[v].lz == level) out[n_out++] = pack_item(val[v].x, val[v].y, val[v].z);
    return n_out;
}

}  // namespace

// One thread per episode.  mode 0 (init): every bin gets episode e0 into buffer 0 and e0 + 1 into buffer 1;
// mode 1 (refill): the bins on the compacted refill list (appended by the step kernel at auto-reset) get episode
// sc.episode + 1 into the buffer they are not playing.  p is the whole env (ep_bin0 == 0).
__global__ void __launch_bounds__(128) bpp_gen_kernel(const EnvParams p, const int mode, const uint32_t e0, const int use_smem)
{
    extern __shared__ __align__(16) uint8_t gen_smem[];
    // The per-episode work is long and serial, so throughput comes from many resident warps rather than from full
    // ones: only kGenLanes lanes of a warp take an episode each.
    const long long n_work = mode == 0 ? (long long)p.N : (long long)*p.gen_count;
    const int lane = threadIdx.x & 31;
    if (lane >= kGenLanes) return;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long wi = warp_id * kGenLanes + lane; wi < n_work; wi += n_warps * kGenLanes) {
        const long long bin = mode == 0 ? wi : (long long)p.gen_list[wi];
        const BinScalars s = p.sc[bin];
        const int cur = mode == 0 ? 0 : (int)(s.traj & 1u);
        // work space: shared memory when the launch could afford it (global-memory scratch costs a DRAM/L2 round trip
        // per dependent access of these strictly serial algorithms), else the bin's global scratch
        uint8_t *scratch = use_smem ? gen_smem + (size_t)((threadIdx.x >> 5) * kGenLanes + lane) * (size_t)p.gen_scratch_stride
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
                const size_t o2 = o1 + (((size_t)p.ep_cap * 4 + 15) & ~(size_t)15);
                const size_t o3 = o2 + (size_t)p.A_pad;
                n = gen_cut1(p, g, out, leaf, reinterpret_cast<uint16_t *>(scratch + o1), scratch + o2,
                             reinterpret_cast<Piece *>(scratch + o3), p.ep_cap);
            } else {
                Piece *inv = reinterpret_cast<Piece *>(scratch);
                Piece *val = inv + p.ep_cap + 64;
                n = gen_cut2(p, g, out, inv, val, p.ep_cap);
            }
            p.ep_len2[(size_t)buf * p.ep_buf_bins + (size_t)bin] = n;
        }
    }
}

cudaError_t launch_gen(const EnvParams &p, int mode, uint32_t e0, int num_sms, cudaStream_t st)
{
    long long want = ((long long)p.N + 4 * kGenLanes - 1) / (4 * kGenLanes);      // 4 warps per block
    const long long cap = (long long)num_sms * 12;
    if (mode == 1 && want > cap) want = cap;              // the list length is only known on the device
    const size_t smem = (size_t)4 * kGenLanes * (size_t)p.gen_scratch_stride;
    const int use_smem = smem <= 200 * 1024;
    if (use_smem && smem > 48 * 1024) {
        cudaError_t ea = cudaFuncSetAttribute(bpp_gen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ea != cudaSuccess) return ea;
    }
    bpp_gen_kernel<<<(unsigned)want, 128, use_smem ? smem : 0, st>>>(p, mode, e0, use_smem);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || mode == 0) return e;
    return cudaMemsetAsync(p.gen_count, 0, sizeof(int32_t), st);     // list consumed
}

size_t gen_scratch_bytes_per_bin(int gen_kind, int cap)
{
    if (gen_kind == 2)      // leaf pieces | meta + candidate lists | shadow heightmap (<= 1024) | split stack
        return (((size_t)cap * sizeof(Piece) + 15) & ~(size_t)15) + (((size_t)cap * 4 + 15) & ~(size_t)15) + 1024 +
               (((size_t)kStack * sizeof(Piece) + 15) & ~(size_t)15);
    return (((size_t)(cap + 64) * 2 * sizeof(Piece)) + 15) & ~(size_t)15;
}

}  // namespace bpp
