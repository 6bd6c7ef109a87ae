// bpp_launch.h -- host-side declarations shared by the translation units of libbpp_b200.
#pragma once

#include "bpp_common.cuh"      // brings in <cuda_runtime.h>

namespace bpp {

// bpp_step.cu
cudaError_t launch_step(const EnvParams &p, const StepIO &io, int mode, int rot, int variant, bool vec4, int num_sms,
                        cudaStream_t st);

// bpp_gen.cu
cudaError_t launch_gen(const EnvParams &p, int mode, uint32_t e0, int num_sms, cudaStream_t st);
size_t gen_scratch_bytes_per_bin(int gen_kind, int cap);

// bpp_mask.cu
cudaError_t launch_mask_from_obs(const float *obs, int n, int L, int W, int H, int rot, int variant, float *mask,
                                 int grid, cudaStream_t st);
cudaError_t launch_mask_from_plain(const int32_t *plain, const int32_t *box, int n, int L, int W, int H, int rot,
                                   int variant, float *mask, int grid, cudaStream_t st);
cudaError_t launch_check_box(const int32_t *plain, const int32_t *query, int n, int L, int W, int H, int variant,
                             int32_t *out, cudaStream_t st);
cudaError_t launch_export_state(const EnvParams &p, int32_t *plain, int32_t *box, int32_t *counter, int32_t *vol,
                                cudaStream_t st);
cudaError_t launch_import_state(const EnvParams &p, const int32_t *plain, const int32_t *box, const int32_t *counter,
                                const int32_t *vol, int32_t *bad, cudaStream_t st);
cudaError_t launch_debug_draws(unsigned long long seed, uint32_t domain, uint32_t aux, unsigned long long gid, int n,
                               uint32_t *out, cudaStream_t st);

cudaError_t launch_debug_curand(unsigned long long seed, unsigned long long subsequence, int n, uint32_t *out,
                                cudaStream_t st);

}  // namespace bpp
