// tma_probe.cu -- microbenchmark behind the step kernel's output path (sm_100a): how fast can a persistent grid of
// 8-warp CTAs write rows of a few hundred bytes per "bin" to HBM
//   (a) with 128-bit STG from registers (what the round-1 kernel does: five 400-byte rows per bin),
//   (b) with cp.async.bulk.global.shared::cta (UBLKCP) issued by one lane per row, rows taken from shared memory,
//   (c) the same with the rows staged by STS first,
// and how fast 32-bin batches of heightmap tiles (3 584 B) arrive by cp.async.bulk + mbarrier vs per-bin LDG.128.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/tma_probe tools/tma_probe.cu && tools/_build/tma_probe
// Prints one JSON line per experiment.  Not part of the product; profiles/README.md quotes its results.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int kWarps = 8, kThreads = kWarps * 32;
constexpr int kRow = 400, kRowsPerBin = 5, kBinBytes = kRow * kRowsPerBin;   // obs f32[4][100] + mask f32[100]

// (a) STG.128 from registers, rows of 400 B = 25 float4 per row (lanes 0..24), like the round-1 kernel
__global__ void __launch_bounds__(kThreads, 4) k_stg(float *out, int n_bins)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long b = (long long)blockIdx.x * kWarps + warp; b < n_bins; b += (long long)gridDim.x * kWarps) {
        float4 *o = reinterpret_cast<float4 *>(out + b * (kBinBytes / 4));
        const float v = (float)(b & 7);
        if (lane < 25)
#pragma unroll
            for (int r = 0; r < kRowsPerBin; ++r) o[r * 25 + lane] = make_float4(v, v, v, v);
    }
}

// (b)/(c)/(d) rows by bulk copy.  MODE 0: rows_per_copy rows in one copy from a constant shared block (no staging);
// MODE 1: stage `staged` rows with STS.128 per bin (double-buffered), the other rows straight from a constant LUT.
template <int MODE>
__global__ void __launch_bounds__(kThreads, 4) k_bulk(float *out, int n_bins, int rows_per_copy, int staged)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 *lut = reinterpret_cast<float4 *>(smem);                 // 8 values x 2000 B of constants
    for (int k = threadIdx.x; k < 8 * kBinBytes / 16; k += kThreads) {
        const float v = (float)(k / (kBinBytes / 16));
        lut[k] = make_float4(v, v, v, v);
    }
    uint8_t *stage = smem + 8 * kBinBytes + warp * (2 * kBinBytes);  // two staging buffers per warp
    __syncthreads();
    fence_async_smem();
    int it = 0;
    for (long long b = (long long)blockIdx.x * kWarps + warp; b < n_bins; b += (long long)gridDim.x * kWarps, ++it) {
        uint8_t *o = reinterpret_cast<uint8_t *>(out) + b * kBinBytes;
        const uint8_t *src = smem + (b & 7) * kBinBytes;
        if (MODE == 1) {
            uint8_t *st = stage + (it & 1) * kBinBytes;
            if (lane == 0) bulk_wait_read<1>();                      // the copy that read this buffer two bins ago is done
            __syncwarp();
            const float v = (float)(b & 7);
            float4 *s4 = reinterpret_cast<float4 *>(st);
            for (int k = lane; k < staged * 25; k += 32) s4[k] = make_float4(v, v, v, v);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (staged > 0) bulk_store(o, st, staged * kRow);
                for (int r = staged; r < kRowsPerBin; ++r) bulk_store(o + r * kRow, src + r * kRow, kRow);
                bulk_commit();
            }
        } else {
            if (lane == 0) {
                for (int r = 0; r < kRowsPerBin; r += rows_per_copy) bulk_store(o + r * kRow, src + r * kRow, rows_per_copy * kRow);
                bulk_commit();
                bulk_wait_read<8>();
            }
        }
        __syncwarp();
    }
    if (lane == 0) bulk_wait_read<0>();
}


// (e) the same rows, but every warp owns BATCHES of `bpb` consecutive bins which it writes one after the other (the step
// kernel's assignment: neighbouring warps are bpb bins = bpb * 2000 bytes apart), by five bulk stores per bin
__global__ void __launch_bounds__(kThreads, 4) k_bulk_batched(float *out, int n_bins, int bpb)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 *lut = reinterpret_cast<float4 *>(smem);
    for (int k = threadIdx.x; k < 8 * kBinBytes / 16; k += kThreads) {
        const float v = (float)(k / (kBinBytes / 16));
        lut[k] = make_float4(v, v, v, v);
    }
    __syncthreads();
    fence_async_smem();
    const long long n_batches = n_bins / bpb;
    for (long long batch = (long long)blockIdx.x * kWarps + warp; batch < n_batches; batch += (long long)gridDim.x * kWarps) {
        for (int j = 0; j < bpb; ++j) {
            const long long b = batch * bpb + j;
            uint8_t *o = reinterpret_cast<uint8_t *>(out) + b * kBinBytes;
            const uint8_t *src = smem + (b & 7) * kBinBytes;
            if (lane == 0) {
                for (int r = 0; r < kRowsPerBin; ++r) bulk_store(o + r * kRow, src + r * kRow, kRow);
                bulk_commit();
                bulk_wait_read<8>();
            }
            __syncwarp();
        }
    }
    if (lane == 0) bulk_wait_read<0>();
}

// tile loads: 112-byte u8 tiles.  MODE 0: per bin, lanes 0..6 LDG.128 -> STS.128 (round-1 kernel, prefetch one ahead);
// MODE 1: one bulk load of 32 tiles (3 584 B) per batch + mbarrier, double-buffered.  The warp sums the bytes.
template <int MODE>
__global__ void __launch_bounds__(kThreads, 4) k_tiles(const uint8_t *hm, unsigned *sink, int n_bins)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[kWarps][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *buf = smem + warp * (2 * 3584);
    if (lane == 0) { mbar_init(&bars[warp][0], 1); mbar_init(&bars[warp][1], 1); }
    __syncthreads();
    unsigned acc = 0;
    const long long n_batches = n_bins / 32, stride = (long long)gridDim.x * kWarps;
    long long batch = (long long)blockIdx.x * kWarps + warp;
    if (MODE == 1 && batch < n_batches && lane == 0) {
        mbar_expect_tx(&bars[warp][0], 3584);
        bulk_load(buf, hm + batch * 3584, 3584, &bars[warp][0]);
    }
    int it = 0;
    for (; batch < n_batches; batch += stride, ++it) {
        if (MODE == 1) {
            const int cur = it & 1;
            if (batch + stride < n_batches && lane == 0) {
                mbar_expect_tx(&bars[warp][cur ^ 1], 3584);
                bulk_load(buf + (cur ^ 1) * 3584, hm + (batch + stride) * 3584, 3584, &bars[warp][cur ^ 1]);
            }
            mbar_wait(&bars[warp][cur], (it >> 1) & 1);
            const uint4 *t4 = reinterpret_cast<const uint4 *>(buf + cur * 3584);
            for (int b = 0; b < 32; ++b) {
                if (lane < 7) { const uint4 v = t4[b * 7 + lane]; acc += v.x ^ v.y ^ v.z ^ v.w; }
            }
            __syncwarp();
        } else {
            const uint4 *g4 = reinterpret_cast<const uint4 *>(hm + batch * 3584);
            uint4 *t4 = reinterpret_cast<uint4 *>(buf);
            uint4 pre = make_uint4(0, 0, 0, 0);
            if (lane < 7) pre = g4[lane];
            for (int b = 0; b < 32; ++b) {
                if (lane < 7) t4[lane] = pre;
                if (b + 1 < 32 && lane < 7) pre = g4[(b + 1) * 7 + lane];
                __syncwarp();
                if (lane < 7) { const uint4 v = t4[lane]; acc += v.x ^ v.y ^ v.z ^ v.w; }
                __syncwarp();
            }
        }
    }
    if (acc == 0x12345u) sink[0] = acc;
}

template <typename F>
static float time_ms(F launch, int reps = 20)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i) launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms / reps < best) best = ms / reps;
    }
    CK(cudaGetLastError());
    return best;
}

int main()
{
    const int n_bins = 1 << 19;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int grid = prop.multiProcessorCount * 4;
    float *out; CK(cudaMalloc(&out, (size_t)n_bins * kBinBytes));
    uint8_t *hm; CK(cudaMalloc(&hm, (size_t)n_bins * 112)); CK(cudaMemset(hm, 1, (size_t)n_bins * 112));
    unsigned *sink; CK(cudaMalloc(&sink, 4));
    const double gb = (double)n_bins * kBinBytes / 1e9;
    const size_t smem_bulk = 8 * kBinBytes + kWarps * 2 * kBinBytes;      // 16 000 + 32 000
    CK(cudaFuncSetAttribute(k_bulk<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bulk));
    CK(cudaFuncSetAttribute(k_bulk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bulk));
    CK(cudaFuncSetAttribute(k_tiles<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWarps * 2 * 3584));
    CK(cudaFuncSetAttribute(k_tiles<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWarps * 2 * 3584));

    float ms = time_ms([&] { k_stg<<<grid, kThreads>>>(out, n_bins); });
    printf("{\"exp\": \"stg128_5x400B_rows\", \"ms\": %.4f, \"GBps\": %.0f}\n", ms, gb / (ms * 1e-3));
    for (int rpc : {1, 5}) {
        ms = time_ms([&] { k_bulk<0><<<grid, kThreads, smem_bulk>>>(out, n_bins, rpc, 0); });
        printf("{\"exp\": \"bulk_store_%dx%dB_per_bin_no_staging\", \"ms\": %.4f, \"GBps\": %.0f, \"copies_per_us_per_sm\": %.1f}\n",
               5 / rpc, rpc * kRow, ms, gb / (ms * 1e-3), (double)n_bins * (5 / rpc) / prop.multiProcessorCount / (ms * 1e3));
    }
    for (int staged : {0, 2, 5}) {
        ms = time_ms([&] { k_bulk<1><<<grid, kThreads, smem_bulk>>>(out, n_bins, 1, staged); });
        printf("{\"exp\": \"bulk_store_staged%d_rows_sts_then_%d_copies\", \"ms\": %.4f, \"GBps\": %.0f}\n", staged,
               (staged > 0) + 5 - staged, ms, gb / (ms * 1e-3));
    }
    CK(cudaFuncSetAttribute(k_bulk_batched, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bulk));
    for (int g : {3, 4}) for (int bpb : {1, 16, 32}) {
        ms = time_ms([&] { k_bulk_batched<<<prop.multiProcessorCount * g, kThreads, smem_bulk>>>(out, n_bins, bpb); });
        printf("{\"exp\": \"bulk_store_5x400B_batches_of_%d_bins_per_warp_%d_ctas_per_sm\", \"ms\": %.4f, \"GBps\": %.0f}\n", bpb, g, ms, gb / (ms * 1e-3));
    }
    const double gbt = (double)n_bins * 112 / 1e9;
    ms = time_ms([&] { k_tiles<0><<<grid, kThreads, kWarps * 2 * 3584>>>(hm, sink, n_bins); });
    printf("{\"exp\": \"tiles_ldg128_sts_per_bin\", \"ms\": %.4f, \"GBps\": %.0f}\n", ms, gbt / (ms * 1e-3));
    ms = time_ms([&] { k_tiles<1><<<grid, kThreads, kWarps * 2 * 3584>>>(hm, sink, n_bins); });
    printf("{\"exp\": \"tiles_bulk_load_3584B_per_batch\", \"ms\": %.4f, \"GBps\": %.0f}\n", ms, gbt / (ms * 1e-3));
    return 0;
}
