// vsr_internal.h -- what vsr_ga.cu needs from vsr_api.cu (not part of the C ABI; hidden from the .so's exports)
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include <string>

#include "../../include/vsr_eval.h"

namespace vsr_internal {
#define VSR_HIDDEN __attribute__((visibility("hidden")))
// device / stream the GA state lives on (the context's first device)
VSR_HIDDEN int primary_device(const vsr_ctx* ctx);
VSR_HIDDEN cudaStream_t primary_stream(const vsr_ctx* ctx);
VSR_HIDDEN bool ready(const vsr_ctx* ctx);
VSR_HIDDEN int genome_len(const vsr_ctx* ctx);
// Evaluate robots [lo, hi) of a batch whose genomes ([.][G] doubles, row 0 = robot 0) and fitness vector live on the
// primary device; work is enqueued on / joined back into `st` (a stream of the primary device).  With several devices in
// the context the range is split by contiguous index, genomes go out and fitness values come back by peer copies.
VSR_HIDDEN int eval_range(vsr_ctx* ctx, const double* d_genomes, size_t lo, size_t hi, double* d_fitness, cudaStream_t st);
VSR_HIDDEN const std::string& last_error(const vsr_ctx* ctx);
// episode-kernel instantiations (vsr_kernel_inst.cu): launch shapes 0..3 of the FP32 build, the FP64 parity build
VSR_HIDDEN const void* kernel_f32_s0(int ctrl, int full_scaff);
VSR_HIDDEN const void* kernel_f32_s1(int ctrl, int full_scaff);
VSR_HIDDEN const void* kernel_f32_s2(int ctrl, int full_scaff);
VSR_HIDDEN const void* kernel_f32_s3(int ctrl, int full_scaff);
VSR_HIDDEN const void* kernel_f32_s3_area(void);
VSR_HIDDEN const void* kernel_f32_s2_area(void);
VSR_HIDDEN const void* kernel_f64(int ctrl);
VSR_HIDDEN const void* kernel_obs(int precision, int ctrl); // per-observation summaries (vsr_eval_ex): generic scaffolding / sensors, 8 warps (FP64: 4)
VSR_HIDDEN const void* kernel_obs_scaff(int ctrl);          // ... default scaffolding folded in, generic sensors, 8 warps
VSR_HIDDEN const void* kernel_obs_area(int warps);          // ... the default robot (one AreaRatio reading into a centralized MLP), 8 or 10 warps
} // namespace vsr_internal
