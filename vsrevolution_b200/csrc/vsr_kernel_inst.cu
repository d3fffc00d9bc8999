// vsr_kernel_inst.cu -- instantiates vsr_episode_kernel for ONE launch shape per object file (-DVSR_INST=0..3: FP32 shapes
// {4,2} {5,2} {8,1} {10,1} warps x min CTAs/SM; 4: the FP64 parity build; 5-7: the observation builds), so the shapes compile in parallel.
#include "vsr_internal.h"
#include "vsr_kernel.cuh"

#ifndef VSR_INST
#error "compile with -DVSR_INST=0..7"
#endif

namespace {
template <typename R, int W, int B, int SCAFF>
const void* pick(int ctrl) {
  switch (ctrl) {
    case VSR_CTRL_PHASE_SIN: return (const void*)vsr::vsr_episode_kernel<R, VSR_CTRL_PHASE_SIN, W, B, SCAFF>;
    case VSR_CTRL_MLP_CENTRALIZED: return (const void*)vsr::vsr_episode_kernel<R, VSR_CTRL_MLP_CENTRALIZED, W, B, SCAFF>;
    default: return (const void*)vsr::vsr_episode_kernel<R, VSR_CTRL_MLP_DISTRIBUTED_HOMO, W, B, SCAFF>;
  }
}
// full_scaff: all four scaffoldings enabled (the Voxel default) -> the build with the mask folded at compile time
template <typename R, int W, int B>
const void* pick2(int ctrl, int full_scaff) { return full_scaff ? pick<R, W, B, 0xF>(ctrl) : pick<R, W, B, -1>(ctrl); }
} // namespace

namespace vsr_internal {
#if VSR_INST == 0
const void* kernel_f32_s0(int ctrl, int full_scaff) { return pick2<float, 4, 2>(ctrl, full_scaff); }
#elif VSR_INST == 1
const void* kernel_f32_s1(int ctrl, int full_scaff) { return pick2<float, 5, 2>(ctrl, full_scaff); }
#elif VSR_INST == 2
const void* kernel_f32_s2(int ctrl, int full_scaff) { return pick2<float, 8, 1>(ctrl, full_scaff); }
// the sub-wave (latency) partner of kernel_f32_s3_area: same specialisation, so a robot gets the same bits in either launch shape
const void* kernel_f32_s2_area(void) { return (const void*)vsr::vsr_episode_kernel<float, VSR_CTRL_MLP_CENTRALIZED, 8, 1, 0xF, 1>; }
#elif VSR_INST == 3
const void* kernel_f32_s3(int ctrl, int full_scaff) { return pick2<float, 10, 1>(ctrl, full_scaff); }
// the reference's default robot (CentralizedSensing MLP over one AreaRatio reading per voxel): sensor loop folded away
const void* kernel_f32_s3_area(void) { return (const void*)vsr::vsr_episode_kernel<float, VSR_CTRL_MLP_CENTRALIZED, 10, 1, 0xF, 1>; }
#elif VSR_INST == 4
const void* kernel_f64(int ctrl) { return pick<double, 4, 1, -1>(ctrl); }
#elif VSR_INST == 5
// the builds that also produce the per-observation summaries of vsr_eval_ex (posture, footprints, centre + angle): the flexible
// 8-warp FP32 shape and the FP64 parity build, generic scaffolding / sensors
template <typename R, int W>
const void* pick_obs(int ctrl) {
  switch (ctrl) {
    case VSR_CTRL_PHASE_SIN: return (const void*)vsr::vsr_episode_kernel<R, VSR_CTRL_PHASE_SIN, W, 1, -1, 0, 1>;
    case VSR_CTRL_MLP_CENTRALIZED: return (const void*)vsr::vsr_episode_kernel<R, VSR_CTRL_MLP_CENTRALIZED, W, 1, -1, 0, 1>;
    default: return (const void*)vsr::vsr_episode_kernel<R, VSR_CTRL_MLP_DISTRIBUTED_HOMO, W, 1, -1, 0, 1>;
  }
}
const void* kernel_obs(int precision, int ctrl) { return precision == 64 ? pick_obs<double, 4>(ctrl) : pick_obs<float, 8>(ctrl); }
#elif VSR_INST == 6
// ... with the Voxel default scaffolding folded in (the generic-scaffolding build is ~20 % slower), 8-warp FP32 shape
const void* kernel_obs_scaff(int ctrl) {
  switch (ctrl) {
    case VSR_CTRL_PHASE_SIN: return (const void*)vsr::vsr_episode_kernel<float, VSR_CTRL_PHASE_SIN, 8, 1, 0xF, 0, 1>;
    case VSR_CTRL_MLP_CENTRALIZED: return (const void*)vsr::vsr_episode_kernel<float, VSR_CTRL_MLP_CENTRALIZED, 8, 1, 0xF, 0, 1>;
    default: return (const void*)vsr::vsr_episode_kernel<float, VSR_CTRL_MLP_DISTRIBUTED_HOMO, 8, 1, 0xF, 0, 1>;
  }
}
#else
// VSR_INST == 7: ... and for the reference's default robot (the specialisation of kernel_f32_s3_area / kernel_f32_s2_area): the 8-warp
// shape for batches of up to 8 robot-warps per SM, the 10-warp / 168-register shape above (4096 bipeds = 9.2 warps per SM stay one wave)
const void* kernel_obs_area(int warps) {
  return warps == 10 ? (const void*)vsr::vsr_episode_kernel<float, VSR_CTRL_MLP_CENTRALIZED, 10, 1, 0xF, 1, 1>
                     : (const void*)vsr::vsr_episode_kernel<float, VSR_CTRL_MLP_CENTRALIZED, 8, 1, 0xF, 1, 1>;
}
#endif
} // namespace vsr_internal
