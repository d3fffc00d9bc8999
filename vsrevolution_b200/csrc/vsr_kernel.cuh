// vsr_kernel.cuh -- K1 `vsr_episode_kernel`: one full Locomotion episode per robot, lane = voxel.
//
// Replaces, for a whole population at once, Locomotion.apply(robot) (2dhmsr 1.12.1) and the dyn4j
// World.step(1) it drives; call site /root/reference/src/main/java/it/units/erallab/locomotion/Starter.java:83-94.
//
// Mapping:
//   * one warp lane per voxel; a robot with n voxels owns n consecutive lanes; floor(32/n) robots per warp
//   * the 4 corner masses of a voxel (x,y,angle,vx,vy,w) and the voxel's 18 spring + 4 rope joints live in
//     the lane's REGISTERS for the whole episode; intra-voxel joints are solved without any memory traffic
//   * inter-voxel weld joints are solved by the lane that created them (Robot.assemble order); the partner
//     masses travel by warp shuffle (horizontal round, then vertical round: a 2-colouring of the welds)
//   * ground contacts are per corner mass (ground is static), kept in a per-warp shared-memory pool
//   * controller (MLP / PhaseSin / DistributedSensing) runs per robot on the same lanes once per step
//   * positions are stored relative to a per-lane integer origin so FP32 keeps ~1e-7 absolute resolution
//   * HBM traffic: genome in (G doubles), one vsr_result out per robot; nothing else during the episode
//   * the warps of a CTA are re-aligned once per step (__syncthreads) so they share instruction fetches:
//     the step body is several thousand instructions and the first version of this kernel spent 80% of
//     its cycles in `no_instruction` stalls (profiles/r1_ncu_summary.md)
//
// Constraint order == the oracle's canonical order (oracle/vsr_oracle.c header).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vsr_eval.h"

namespace vsr {

constexpr int kMaxVox = 32;
constexpr int kMaxSensors = 8;   // distinct sensors over all voxels (position-dependent values are per voxel: vs_param)
constexpr int kMaxVoxSensors = 8; // sensors per voxel
constexpr int kMaxReadings = 12;  // readings per voxel (spinedTouchSighted's top-right voxel: vxy + a + cpg + 5 lidar rays = 9)
constexpr int kMaxLayers = 8;
constexpr int kMaxMan = 3;       // manifolds per corner mass (ground polygons it can touch at once)
constexpr int kMaxAvg = 6;       // averaged readings per voxel
constexpr int kPark = 20;         // X Y A (12) + polygon cache (6) + origin (2)
// contact-pool slot (36 R words, 16-byte aligned vectors; stride 36 = 4 x odd -> 128-bit accesses of 8 consecutive slots
// hit all 32 banks once):
//   geometry (rebuilt by every detect):   [0..3] nx ny px0 py0 | [4..7] d0 px1 py1 d1
//   accumulated impulses, double-buffered by detect parity p for warm starting:   [8 + 4p .. 11 + 4p] jn0 jt0 jn1 jt1
//   solver part (Jacobian rows rn = r x n, rt = r x t):   [16..19] rn0 rt0 mN0 mT0 | [20..23] rn1 rt1 mN1 mT1 |
//   [24..27] vb0 vb1 K12 iK11 | [28..31] iK12 iK22 - - | [32..35] lx0 ly0 lx1 ly1
constexpr int kPI = 8, kPS = 16;
constexpr int kPoolStride = 36;
// phase-alignment barriers of the specialised build (bits: KParams::sync_mask): step start, before the velocity loop, before
// the spring initialisation, before / after detect (tools/ab_run.sh sweeps on B200, one 10-warp CTA per SM); the generic
// build reads the mask from P (tuning only)
#ifndef VSR_SYNC_MASK_DEFAULT
#define VSR_SYNC_MASK_DEFAULT 115
#endif
constexpr int kSyncMaskDefault = VSR_SYNC_MASK_DEFAULT;
#ifndef VSR_POS_V2
#define VSR_POS_V2 1 // position iterations: branch-free rotations (deferred exact refresh), branch-free ropes, unrolled contact points
#endif
#ifndef VSR_REFRESH_CS
#define VSR_REFRESH_CS 1 // FP32: exact sin / cos from the accumulated angles after the position iterations
#endif
#ifndef VSR_CACHE_MASS
#define VSR_CACHE_MASS 0 // contact phases: the touching mass + its manifold slots cached in a register per step (A/B r2: -1.7 %: not kept)
#endif
#ifndef VSR_SOFT_SYNC
#define VSR_SOFT_SYNC 0 // > 0: bounded-skew phase alignment (warps may run this many boundaries ahead); 0: __syncthreads
#endif

struct SensorK {
  int kind, axes, rotated, soft_norm;
  double dom_min, dom_inv; // soft norm: (x - dom_min) * dom_inv
  int avg;                 // index into the averaged-reading table or -1
  int off, dim;            // offset inside a uniform voxel's reading vector (unused with per-voxel lists) / number of readings
  int noise_row;           // first row of the Noisy table for this sensor (rows depend on the sensor's dimension only)
  int n_rays;              // LIDAR
  double param;            // TIME_SIN: frequency; LIDAR: ray length (CONSTANT values and LIDAR sides are per voxel: KParams::vs_param)
};

// FP32 copies of the scalar constants: the FP32 kernel reads them straight from the constant bank
// (no double->float conversion, no long-lived register)
struct KConstF {
  float c_dt, c_half, c_im, c_ii, c_mu, c_lin_tol, c_ang_tol, c_max_corr, c_baum, c_gdt, c_lin_f, c_ang_f, c_sp_sigma, c_sp_boversig,
      c_sp_cfrac, c_rope_lower, c_rope_E, c_wa, c_wA, c_wh, c_wDA, c_wDh, c_w_i2i, c_rest_vel, c_rest_e, c_max_trans, c_max_trans2, c_max_rot,
      c_inv_area2, c_mass_off, c_base_pad;
  float c_range_min[3], c_range_rest[3], c_range_max[3];
};

struct KParams {
  KConstF cf;
  // dyn4j settings
  double dt, g, baumgarte, lin_tol, ang_tol, max_lin_corr, max_trans, max_rot, rest_vel;
  int vel_iters, pos_iters;
  // voxel
  double half, inv_m, inv_i, mu, rest_e, lin_f, ang_f;
  double range_min[3], range_rest[3], range_max[3];
  double rope_lower, side_len;
  double sp_sigma, sp_beta, sp_cfrac; // soft-constraint constants
  double weld_anchor;                 // |local anchor 2| of a weld = mass side
  int scaff_mask, limit_contraction;
  // body
  int n_vox, lanes_per_robot, robots_per_warp;
  int8_t nb_l[kMaxVox], nb_r[kMaxVox], nb_d[kMaxVox], nb_u[kMaxVox];
  double vox_x0[kMaxVox], vox_y0[kMaxVox]; // voxel centre after Locomotion's initial placement
  // terrain
  const double* tx;
  const double* ty;
  int n_t;
  double base_y, ty_max; // bottom of the ground body; highest terrain vertex
  // coarse height map of the (shared) terrain for the contact broadphase: thmax[b] >= every terrain height over
  // [thx0 + (b - 1) / thinv, thx0 + (b + 2) / thinv]; NULL (per-evaluation terrains) = test against ty_max
  const float* thmax;
  float thx0, thinv;
  int thn;
  // per-evaluation terrains (`-rnd`, Starter.java:89-93): robot i runs on vertices [toff[i], toff[i+1]) of tx / ty, its ground
  // body ends at tbase[i] and its initial placement is shifted by (tplace[2i], tplace[2i+1]); NULL = one terrain for all
  const int* toff;
  const double* tbase;
  const double* tplace;
  // sensors: the distinct sensors of the body, and every voxel's own list (RobotUtils.buildSensorizingFunction: `uniform-...` gives the
  // same list everywhere, `spinedTouch-...` / px / py / l5 depend on the voxel's grid position)
  int n_sensors, sens_dim, n_avg, hist_cap;
  SensorK sensors[kMaxSensors];
  int vs_max;                                   // longest per-voxel list
  uint8_t vs_n[kMaxVox];                        // sensors of voxel v
  uint8_t vs_idx[kMaxVox][kMaxVoxSensors];      // ... and which (index into sensors[])
  double vs_param[kMaxVox][kMaxVoxSensors];     // CONSTANT: the value; LIDAR: the side (0 N, 1 E, 2 S, 3 W)
  uint8_t vox_dim[kMaxVox];                     // readings of voxel v
  uint16_t vox_in_off[kMaxVox];                 // CentralizedSensing: first MLP input of voxel v
  const uint8_t* win_tab; // [n_avg][n_steps+1] window length (in readings) at step k
  const double* noise;    // Noisy sensors: [sens_dim][n_steps+1] additive noise of reading r at step k (NULL = none)
  void* hist;             // R[hist_cap][n_avg][total_lanes] ring of raw readings
  // controller
  int ctrl_kind, n_layers, layers[kMaxLayers], signals, in_act, genome_len, max_neurons;
  int w_stride; // heterogeneous DistributedSensing: weights of voxel v start at v * w_stride inside the robot's row (0 = one shared MLP)
  int time_sin; // PHASE_SIN kernel running TimeFunctions sinusoids: genome row = (frequency, phase, amplitude) per voxel
  double freq, amp;
  // episode
  int n_steps, k_first, k_last;
  double t_first, t_last; // accumulated times of the first/last kept observation (host-emulated, same IEEE adds)
  // io
  const double* genomes;
  double* fitness;
  vsr_result* results;
  int* overflow_count;   // device counter: += 1 per warp whose contact pool overflowed (checked by the host even when no vsr_result is requested)
  double* centers;       // optional [n_robots][k_last - k_first + 1][2]: robot centre (x, y) at every kept observation (vsr_eval_centers)
  void* energy;          // V4<R>[total lanes]: per-voxel energy accumulators (sum f^2 dt, sum |da|, sum |d(a^2)|, last a); used only with results
  int n_robots;
  // robot transformation `broken-<ratio>-<seed>` (RobotUtils.buildRobotTransformation, Starter.java:126,428): bit v of frozen[robot] =
  // the actuator of voxel v is frozen at 0 from the first step (BreakableVoxel, ACTUATOR FROZEN, trigger TIME >= 0); NULL = identity
  const uint32_t* frozen;
  // per-observation summaries (OBS builds only; vsr_eval_observed): average posture counters [n_robots][64], footprints
  // [n_robots][K] bytes for up to two segment counts, centre trajectory [n_robots][K][3] = (x, y, angle of the central voxel)
  int post_n, fp_n0, fp_n1, central_vox;
  unsigned int* post_cnt;
  uint8_t *footprints0, *footprints1;
  double* centers3;
  // trace (robot 0 only)
  double *tr_vertices, *tr_bodies, *tr_ctrl;
  int trace_steps;
  // shared-memory layout (bytes from the CTA's dynamic smem base; per warp strides)
  int pool_slots;
  int smem_warp_stride; // bytes per warp
  int off_pool_p, off_pool_m, off_pool_cnt, off_w, off_act, off_sig, off_bias, off_arm, off_eff, off_nrm, off_park, off_cold, off_stage, off_pool_o, off_mslots, off_rope;
  int sync_mask; // phase-alignment barriers (bit0 step start, bit1 before velocity loop, bit2 every velocity
                 // iteration, bit3 before position phase, bit4 before detect, bit5 before sensing, bit6 before spring
                 // init, bit7 before weld init, bit8 after the velocity loop, bit9 after the position phase, bit10 after control)
};

static_assert(sizeof(KParams) <= 8192, "KParams is passed by value as a __grid_constant__ kernel parameter");

// ------------------------------------------------------------------------------------------------
// math: the FP32 build uses single-instruction MUFU reciprocals / rsqrt and short polynomial
// sincos / tanh (all within a few ulp); the FP64 build uses the precise libm calls.
// ------------------------------------------------------------------------------------------------
template <class R> struct M;
// A/B switches for the FP32 approximations (profiles/r2_fp32_approx_ab.md: which one ends step-wise tracking of the oracle first):
//   -DVSR_IEEE_DIV      correctly rounded reciprocal / reciprocal square root instead of the MUFU approximations
//   -DVSR_FAST_SINCOS   the Cody-Waite sincos polynomials instead of CUDA libm sincosf (the round-1 form)
//   -DVSR_NO_SMALL_ROTATE  every position-solver rotation re-evaluates sin / cos of the angle (no small-angle update)
//   -DVSR_FAST_TANH     ex2.approx + rcp.approx tanh instead of tanhf (the round-1 form)
//   -DVSR_REFRESH_CS=0  no exact sin / cos refresh after the position iterations (the round-1 behaviour)
template <> struct M<float> {
#ifdef VSR_IEEE_DIV
  static __device__ __forceinline__ float rcp(float x) { return __frcp_rn(x); }
  static __device__ __forceinline__ float rsqrt_(float x) { return __frcp_rn(__fsqrt_rn(x)); }
#else
  static __device__ __forceinline__ float rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
  static __device__ __forceinline__ float rsqrt_(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#endif
  static __device__ __forceinline__ float abs_(float x) { return fabsf(x); }
  static __device__ __forceinline__ float max_(float a, float b) { return fmaxf(a, b); }
  static __device__ __forceinline__ float min_(float a, float b) { return fminf(a, b); }
  static __device__ __forceinline__ float rint_(float a) { return rintf(a); }
  // Cody-Waite reduction by pi/2 + minimax polynomials on [-pi/4, pi/4]; |a| < ~1e5
  static __device__ __forceinline__ void sincos_(float a, float* sn, float* cs) {
#ifndef VSR_FAST_SINCOS
    // CUDA libm sincosf (<= 2 ulp over the whole range): +0.6 % kernel time on config 2, and the step count over which the FP32 kernel
    // tracks the oracle grows 1.4-3.3x (profiles/r2_fp32_approx_ab.md); -DVSR_FAST_SINCOS restores the Cody-Waite polynomials below
    sincosf(a, sn, cs);
    return;
#endif
    const float k = rintf(a * 0.636619772367581343f);
    float r = fmaf(k, -1.57079625129699707031f, a);
    r = fmaf(k, -7.54978941586159635335e-08f, r);
    const float r2 = r * r;
    float s = fmaf(r2, 2.86567956e-6f, -1.98559923e-4f);
    s = fmaf(s, r2, 8.33338592e-3f);
    s = fmaf(s, r2, -1.66666672e-1f);
    s = fmaf(s * r2, r, r);
    float c = fmaf(r2, 2.44677067e-5f, -1.38877297e-3f);
    c = fmaf(c, r2, 4.16666567e-2f);
    c = fmaf(c, r2, -0.5f);
    c = fmaf(c, r2, 1.0f);
    const int q = (int)k;
    const float ss = (q & 1) ? c : s, cc = (q & 1) ? s : c;
    *sn = (q & 2) ? -ss : ss;
    *cs = ((q + 1) & 2) ? -cc : cc;
  }
  // rotate (c,s) by a small angle d (position-solver corrections); falls back to the full evaluation
  static __device__ __forceinline__ void rotate(float& a, float& c, float& s, float d) {
    a += d;
#ifdef VSR_NO_SMALL_ROTATE
    sincos_(a, &s, &c);
    return;
#endif
    if (fabsf(d) < 0.125f) {
      const float d2 = d * d;
      const float sd = d * fmaf(d2, fmaf(d2, 8.33333333e-3f, -1.66666667e-1f), 1.0f);
      const float cd = fmaf(d2, fmaf(d2, fmaf(d2, -1.38888889e-3f, 4.16666667e-2f), -0.5f), 1.0f);
      const float c2 = c * cd - s * sd;
      s = s * cd + c * sd;
      c = c2;
    } else {
      sincos_(a, &s, &c);
    }
  }
  // branch-free variant for the position solver: always the small-angle update; `big` collects whether any step left the
  // polynomial's range, in which case the caller re-evaluates sin / cos of the (exactly accumulated) angles once per iteration
  static __device__ __forceinline__ void rotate_fast(float& a, float& c, float& s, float d, bool& big) {
    a += d;
    big |= !(fabsf(d) < 0.125f);
    const float d2 = d * d;
    const float sd = d * fmaf(d2, fmaf(d2, 8.33333333e-3f, -1.66666667e-1f), 1.0f);
    const float cd = fmaf(d2, fmaf(d2, fmaf(d2, -1.38888889e-3f, 4.16666667e-2f), -0.5f), 1.0f);
    const float c2 = c * cd - s * sd;
    s = s * cd + c * sd;
    c = c2;
  }
  static __device__ __forceinline__ float tanh_(float x) {
#ifndef VSR_FAST_TANH
    // CUDA libm tanhf (<= 2 ulp): the controller's 27 tanh per robot-step cost nothing measurable, while the 2e-7 absolute error of the
    // ex2.approx / rcp.approx form below cut the number of steps the FP32 kernel tracks the oracle by 2.6x on the MLP robots
    // (profiles/r2_fp32_approx_ab.md); -DVSR_FAST_TANH restores it
    return tanhf(x);
#endif
    // 1 - 2/(e^{2x}+1); ex2.approx + rcp.approx, abs error ~2e-7
    const float xc = fminf(fmaxf(x, -15.f), 15.f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(xc * 2.88539008177792681472f));
    return 1.0f - 2.0f * rcp(e + 1.0f);
  }
  static __device__ __forceinline__ float atan2_(float y, float x) { return atan2f(y, x); }
  // a * sin(2 pi f t + phase): the turn count is reduced in double, the sine is FP32
  static __device__ __forceinline__ float phase_sin(double amp, double freq, double t, float phase) {
    double u = freq * t;
    u -= floor(u);
    float s, c;
    sincos_(fmaf((float)u, 6.28318530717958647692f, phase), &s, &c);
    return (float)amp * s;
  }
};
template <> struct M<double> {
  static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
  static __device__ __forceinline__ double rsqrt_(double x) { return 1.0 / sqrt(x); }
  static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
  static __device__ __forceinline__ double max_(double a, double b) { return fmax(a, b); }
  static __device__ __forceinline__ double min_(double a, double b) { return fmin(a, b); }
  static __device__ __forceinline__ double rint_(double a) { return rint(a); }
  static __device__ __forceinline__ void sincos_(double a, double* s, double* c) { sincos(a, s, c); }
  static __device__ __forceinline__ void rotate(double& a, double& c, double& s, double d) { a += d; sincos(a, &s, &c); }
  static __device__ __forceinline__ void rotate_fast(double& a, double& c, double& s, double d, bool&) { a += d; sincos(a, &s, &c); }
  static __device__ __forceinline__ double tanh_(double x) { return tanh(x); }
  static __device__ __forceinline__ double atan2_(double y, double x) { return atan2(y, x); }
  static __device__ __forceinline__ double phase_sin(double amp, double freq, double t, double phase) {
    return amp * sin(2.0 * 3.14159265358979323846 * freq * t + phase);
  }
};

template <class R> struct alignas(4 * sizeof(R)) V4 { R x, y, z, w; };
template <class R> struct alignas(2 * sizeof(R)) V2 { R x, y; };
// compiler barrier on a value: stops loop-invariant code motion from pinning derived terms in registers
__device__ __forceinline__ void opaque(float& x) { asm volatile("" : "+f"(x)); }
__device__ __forceinline__ void opaque(double& x) { asm volatile("" : "+d"(x)); }

#define FULL 0xffffffffu
#define SEL4(arr, k) ((k) == 0 ? arr[0] : ((k) == 1 ? arr[1] : ((k) == 2 ? arr[2] : arr[3])))
#define PUT4(arr, k, v) { arr[0] = (k) == 0 ? (v) : arr[0]; arr[1] = (k) == 1 ? (v) : arr[1]; arr[2] = (k) == 2 ? (v) : arr[2]; arr[3] = (k) == 3 ? (v) : arr[3]; }

// spring table: X(index, bodyA, bodyB, sx1, sy1, sx2, sy2, rangeKind, groupBit)
// canonical order (oracle build_robot): SIDE_INTERNAL, SIDE_EXTERNAL, SIDE_CROSS a, SIDE_CROSS b, CENTRAL_CROSS;
// pairs (0,1) (2,3) (1,2) (3,0) inside each group.
#define VSR_SPRINGS(X)                                                                                       \
  X(0, 0, 1, +1, -1, -1, -1, 0, 1) X(1, 2, 3, -1, +1, +1, +1, 0, 1) X(2, 1, 2, -1, -1, -1, +1, 0, 1)       \
  X(3, 3, 0, +1, +1, +1, -1, 0, 1) X(4, 0, 1, +1, +1, -1, +1, 0, 2) X(5, 2, 3, -1, -1, +1, -1, 0, 2)       \
  X(6, 1, 2, +1, -1, +1, +1, 0, 2) X(7, 3, 0, -1, +1, -1, -1, 0, 2) X(8, 0, 1, +1, +1, -1, -1, 1, 4)       \
  X(9, 2, 3, -1, +1, +1, -1, 1, 4) X(10, 1, 2, +1, -1, -1, +1, 1, 4) X(11, 3, 0, -1, +1, +1, -1, 1, 4)     \
  X(12, 0, 1, +1, -1, -1, +1, 1, 4) X(13, 2, 3, -1, -1, +1, +1, 1, 4) X(14, 1, 2, -1, -1, +1, +1, 1, 4)    \
  X(15, 3, 0, +1, +1, -1, -1, 1, 4) X(16, 0, 2, 0, 0, 0, 0, 2, 8) X(17, 1, 3, 0, 0, 0, 0, 2, 8)

// the same 18 springs as 9 pairs of body-disjoint (hence independent) solves, in canonical order:
// X(i, a, b, j, c, d): spring i joins bodies (a,b), spring j joins (c,d)
#define VSR_SPRING_PAIRS(X)                                                                             \
  X(0, 0, 1, 1, 2, 3, 1) X(2, 1, 2, 3, 3, 0, 1) X(4, 0, 1, 5, 2, 3, 2) X(6, 1, 2, 7, 3, 0, 2)           \
  X(8, 0, 1, 9, 2, 3, 4) X(10, 1, 2, 11, 3, 0, 4) X(12, 0, 1, 13, 2, 3, 4) X(14, 1, 2, 15, 3, 0, 4)

// rope table: X(index, bodyA, bodyB)
#define VSR_ROPES(X) X(0, 0, 1) X(1, 2, 3) X(2, 1, 2) X(3, 3, 0)

// ------------------------------------------------------------------------------------------------
// narrowphase: corner-mass box vs one ground polygon (SAT + reference-face clipping)
// ground polygon CCW: V0=(x0,base) V1=(x1,base) V2=(x1,y1) V3=(x0,y0); box centre (bcx,bcy), half side hh
// ------------------------------------------------------------------------------------------------
// -DVSR_NARROW_NOINLINE: the general narrowphase as a real function (flat / wide terrains take collide_box_top; this body then leaves
// the step's instruction stream)
#ifdef VSR_NARROW_NOINLINE
#define VSR_NARROW_INLINE __noinline__
#else
#define VSR_NARROW_INLINE __forceinline__
#endif
template <class R>
__device__ VSR_NARROW_INLINE int collide_box_ground(R x0, R y0, R x1, R y1, R base, R bcx, R bcy, R c, R s, R hh, R lin_tol,
                                                  R& onx, R& ony, R* opx, R* opy, R* odepth, int* oid) {
  const R hc = hh * c, hs = hh * s;
  const R bx[4] = {bcx - hc + hs, bcx + hc + hs, bcx + hc - hs, bcx - hc - hs};
  const R by[4] = {bcy - hs - hc, bcy + hs - hc, bcy + hs + hc, bcy - hs + hc};
  // ground top-edge normal (edge 2: V2 -> V3); edges 1 / 3 are the vertical sides (+x / -x); edge 0 (bottom,
  // ground_thickness below the surface) can never be the axis of least penetration.
  R tnx, tny;
  {
    const R ex = x0 - x1, ey = y0 - y1;
    const R il = M<R>::rsqrt_(ex * ex + ey * ey);
    tnx = ey * il; tny = -ex * il;
  }
  // separation of the box from the ground's edges 1, 2, 3 (first maximum wins, like the oracle's scan)
  const R bminx = M<R>::min_(M<R>::min_(bx[0], bx[1]), M<R>::min_(bx[2], bx[3]));
  const R bmaxx = M<R>::max_(M<R>::max_(bx[0], bx[1]), M<R>::max_(bx[2], bx[3]));
  R st = tnx * (bx[0] - x1) + tny * (by[0] - y1);
#pragma unroll
  for (int j = 1; j < 4; j++) st = M<R>::min_(st, tnx * (bx[j] - x1) + tny * (by[j] - y1));
  R sA = bminx - x1;
  int eA = 1;
  if (st > sA) { sA = st; eA = 2; }
  if (x0 - bmaxx > sA) { sA = x0 - bmaxx; eA = 3; }
  if (sA > 0) return 0;
  // separation of the ground from the box's edges: normals n0=(s,-c) n1=(c,s) n2=-n0 n3=-n1, each at distance hh
  // from the centre: sep_i = min_j n_i.(g_j - centre) - hh
  const R gx[4] = {x0, x1, x1, x0}, gy[4] = {base, base, y1, y0};
  R umin = 1e30f, umax = -1e30f, wmin = 1e30f, wmax = -1e30f;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const R dx = gx[j] - bcx, dy = gy[j] - bcy;
    const R u = s * dx - c * dy, w = c * dx + s * dy;
    umin = M<R>::min_(umin, u); umax = M<R>::max_(umax, u);
    wmin = M<R>::min_(wmin, w); wmax = M<R>::max_(wmax, w);
  }
  R sB = umin - hh;
  int eB = 0;
  if (wmin - hh > sB) { sB = wmin - hh; eB = 1; }
  if (-umax - hh > sB) { sB = -umax - hh; eB = 2; }
  if (-wmax - hh > sB) { sB = -wmax - hh; eB = 3; }
  if (sB > 0) return 0;
  const int flip = sB > sA + R(0.1) * lin_tol;
  const int e = flip ? eB : eA, e1 = (e + 1) & 3;
  const R bnx[4] = {s, c, -s, -c}, bny[4] = {-c, s, c, -s};
  const R gnx[4] = {0, 1, tnx, -1}, gny[4] = {-1, 0, tny, 0};
  R nx, ny, v1x, v1y, v2x, v2y;
  // reference edge (select chains keep every array index static -> registers, no local memory)
  if (flip) {
    nx = SEL4(bnx, e); ny = SEL4(bny, e);
    v1x = SEL4(bx, e); v1y = SEL4(by, e); v2x = SEL4(bx, e1); v2y = SEL4(by, e1);
  } else {
    nx = SEL4(gnx, e); ny = SEL4(gny, e);
    v1x = SEL4(gx, e); v1y = SEL4(gy, e); v2x = SEL4(gx, e1); v2y = SEL4(gy, e1);
  }
  // incident edge: most anti-parallel normal of the other polygon, first minimum wins
  int ie = 0;
  R mind = 1e30f;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const R d = flip ? nx * gnx[j] + ny * gny[j] : nx * bnx[j] + ny * bny[j];
    if (d < mind) { mind = d; ie = j; }
  }
  const int ie1 = (ie + 1) & 3;
  R cx0, cy0, cx1, cy1;
  if (flip) {
    cx0 = SEL4(gx, ie); cy0 = SEL4(gy, ie); cx1 = SEL4(gx, ie1); cy1 = SEL4(gy, ie1);
  } else {
    cx0 = SEL4(bx, ie); cy0 = SEL4(by, ie); cx1 = SEL4(bx, ie1); cy1 = SEL4(by, ie1);
  }
  R tx = v2x - v1x, ty = v2y - v1y;
  const R itl = M<R>::rsqrt_(tx * tx + ty * ty);
  tx *= itl; ty *= itl;
  // clip the incident edge against the two side planes of the reference edge
#pragma unroll
  for (int side = 0; side < 2; side++) {
    const R o = side == 0 ? tx * v1x + ty * v1y : tx * v2x + ty * v2y;
    const R p0 = tx * cx0 + ty * cy0, p1 = tx * cx1 + ty * cy1;
    const R d0 = side == 0 ? p0 - o : o - p0, d1 = side == 0 ? p1 - o : o - p1;
    if (d0 < 0 && d1 < 0) return 0;
    if (d0 < 0 || d1 < 0) {
      const R u = d0 * M<R>::rcp(d0 - d1);
      const R ix = cx0 + u * (cx1 - cx0), iy = cy0 + u * (cy1 - cy0);
      if (d0 < 0) { cx0 = ix; cy0 = iy; } else { cx1 = ix; cy1 = iy; }
    }
  }
  int np = 0;
  const R sep0 = nx * (cx0 - v1x) + ny * (cy0 - v1y), sep1 = nx * (cx1 - v1x) + ny * (cy1 - v1y);
  const int idb = (flip << 5) | (e << 3) | (ie << 1);
  if (sep0 <= 0) { opx[0] = cx0; opy[0] = cy0; odepth[0] = -sep0; oid[0] = idb; np = 1; }
  if (sep1 <= 0) {
    if (np) { opx[1] = cx1; opy[1] = cy1; odepth[1] = -sep1; oid[1] = idb | 1; }
    else { opx[0] = cx1; opy[0] = cy1; odepth[0] = -sep1; oid[0] = idb | 1; }
    np++;
  }
  onx = flip ? -nx : nx;
  ony = flip ? -ny : ny;
  return np;
}

// Fast narrowphase: the box lies at least kFastMargin inside the polygon's x-range, so the
// ground's top edge is provably the SAT reference face (no flip, side edges cannot win) and the incident box
// edge needs no clipping.  Same formulas as the general path -> identical normal / points / depths / ids.
// Returns -1 when the penetration is too deep for that argument (caller falls back to the general path).
constexpr float kFastMargin = 0.96f, kFastDepth = 0.05f; // margin > 2*half + depth (0.9 + 0.05), DESIGN.md 4.4
template <class R>
__device__ __forceinline__ int collide_box_top(R x1, R y1, R pnx, R pny, R bcx, R bcy, R c, R s, R hh,
                                               R& onx, R& ony, R* opx, R* opy, R* odepth, int* oid) {
  const R hc = hh * c, hs = hh * s;
  const R bx[4] = {bcx - hc + hs, bcx + hc + hs, bcx + hc - hs, bcx - hc - hs};
  const R by[4] = {bcy - hs - hc, bcy + hs - hc, bcy + hs + hc, bcy - hs + hc};
  R sep[4];
#pragma unroll
  for (int j = 0; j < 4; j++) sep[j] = pnx * (bx[j] - x1) + pny * (by[j] - y1);
  const R sA = M<R>::min_(M<R>::min_(sep[0], sep[1]), M<R>::min_(sep[2], sep[3]));
  if (sA > 0) return 0;
  if (sA < -R(kFastDepth)) return -1;
  // incident edge: most anti-parallel box normal {(s,-c),(c,s),(-s,c),(-c,-s)}, first minimum wins
  const R d0 = pnx * s - pny * c, d1 = pnx * c + pny * s;
  int ie = 0;
  R mind = d0;
  if (d1 < mind) { mind = d1; ie = 1; }
  if (-d0 < mind) { mind = -d0; ie = 2; }
  if (-d1 < mind) { mind = -d1; ie = 3; }
  const int ie1 = (ie + 1) & 3;
  const R s0 = SEL4(sep, ie), s1 = SEL4(sep, ie1);
  const int idb = (2 << 3) | (ie << 1);
  int np = 0;
  if (s0 <= 0) { opx[0] = SEL4(bx, ie); opy[0] = SEL4(by, ie); odepth[0] = -s0; oid[0] = idb; np = 1; }
  if (s1 <= 0) {
    const R cx1 = SEL4(bx, ie1), cy1 = SEL4(by, ie1);
    if (np) { opx[1] = cx1; opy[1] = cy1; odepth[1] = -s1; oid[1] = idb | 1; }
    else { opx[0] = cx1; opy[0] = cy1; odepth[0] = -s1; oid[0] = idb | 1; }
    np++;
  }
  onx = pnx; ony = pny;
  return np;
}

// ------------------------------------------------------------------------------------------------
// the episode kernel
// ------------------------------------------------------------------------------------------------
// SCAFF >= 0: the scaffolding mask is a compile-time constant (0xF, the Voxel default: no branch per spring); -1: read from P
// SENS: 0 = generic sensor loop; 1 = exactly one soft-normalised, non-averaged AreaRatio sensor (optionally Noisy)
// OBS: 1 = the build that also produces the per-observation summaries of vsr_eval_observed (posture, footprints, centre + angle)
template <class R, int CTRL, int WARPS, int MINB, int SCAFF = -1, int SENS = 0, int OBS = 0>
__global__ void __launch_bounds__(WARPS * 32, MINB) vsr_episode_kernel(const __grid_constant__ KParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // the 255-register builds (8 and 4 warps) may be launched with fewer warps per CTA than their launch bound, for controllers
  // whose weights need more shared memory per warp; the 10-warp build is always launched full
  const int cta_warps = WARPS == 10 ? WARPS : (int)(blockDim.x >> 5);
  const int gw = blockIdx.x * cta_warps + warp; // global robot-warp
  unsigned char* wsm = smem_raw + (size_t)warp * P.smem_warp_stride;
  R* pool = reinterpret_cast<R*>(wsm + P.off_pool_p);      // [S][kPoolStride]: persist[2][12], solve[20]
  int* poolM = reinterpret_cast<int*>(wsm + P.off_pool_m); // [2][S]
  int* poolCnt = reinterpret_cast<int*>(wsm + P.off_pool_cnt);
  unsigned char* tml = reinterpret_cast<unsigned char*>(wsm + P.off_pool_o); // [128] list of the corner masses (lane*4 + k) that have manifolds
  V4<R>* stage = reinterpret_cast<V4<R>*>(wsm + P.off_stage);  // [3][32] corner-mass velocities (vx, vy, w) x 4 masses per lane
  R* wts = reinterpret_cast<R*>(wsm + P.off_w);            // [robots_per_warp][genome_len]
  R* act = reinterpret_cast<R*>(wsm + P.off_act);          // centralized: [rpw][2][maxN]; distributed: [2][maxN][32]
  R* sig = reinterpret_cast<R*>(wsm + P.off_sig);          // distributed: [2][32][4*signals]
  R* sbias = reinterpret_cast<R*>(wsm + P.off_bias);       // [18][32] per-step spring bias / sigma
  R* spark = reinterpret_cast<R*>(wsm + P.off_park);       // [kPark][32] cold state parked across the velocity loop
  V2<R>* seff = reinterpret_cast<V2<R>*>(wsm + P.off_eff);   // [8][32] effective masses E of springs 0..15, pairs packed
  V4<R>* snrm = reinterpret_cast<V4<R>*>(wsm + P.off_nrm);   // [2][32] unit normals of springs 12..15 (pairs 6, 7): register relief
  V4<R>* srope = reinterpret_cast<V4<R>*>(wsm + P.off_rope); // [2][32] unit normals of the 4 contraction-limit ropes
  V4<R>* sarm = reinterpret_cast<V4<R>*>(wsm + P.off_arm); // [8][32] spring lever arms (r x n) of springs 0..15, pairs packed
  const int S = P.pool_slots;

  const int L = P.lanes_per_robot;
  const int rw = lane / L;                     // robot slot inside the warp
  const int v = lane - rw * L;                 // voxel index inside the robot
  const bool lane_used = rw < P.robots_per_warp;
  const int robot = gw * P.robots_per_warp + rw;
  const bool valid = lane_used && robot < P.n_robots;
  const int base_lane = rw * L;
  const int gl = gw * 32 + lane; // global lane (history ring column)
  const int total_lanes = gridDim.x * cta_warps * 32;

  // neighbours as lane ids packed in one register (self when absent, so shuffles stay in range) + validity bits
  const int vi = valid ? v : 0;
  unsigned nbr, nbv;
  {
    const int l_ = (valid && P.nb_l[vi] >= 0) ? base_lane + P.nb_l[vi] : -1, r_ = (valid && P.nb_r[vi] >= 0) ? base_lane + P.nb_r[vi] : -1;
    const int d_ = (valid && P.nb_d[vi] >= 0) ? base_lane + P.nb_d[vi] : -1, u_ = (valid && P.nb_u[vi] >= 0) ? base_lane + P.nb_u[vi] : -1;
    nbr = (unsigned)(l_ < 0 ? lane : l_) | ((unsigned)(r_ < 0 ? lane : r_) << 8) | ((unsigned)(d_ < 0 ? lane : d_) << 16) | ((unsigned)(u_ < 0 ? lane : u_) << 24);
    nbv = (l_ >= 0 ? 1u : 0u) | (r_ >= 0 ? 2u : 0u) | (d_ >= 0 ? 4u : 0u) | (u_ >= 0 ? 8u : 0u);
  }
#define ln_s ((int)(nbr & 31u))
#define rn_s ((int)((nbr >> 8) & 31u))
#define dn_s ((int)((nbr >> 16) & 31u))
#define un_s ((int)(nbr >> 24))
#define ln (((nbv & 1u) != 0) ? ln_s : -1)
#define rn (((nbv & 2u) != 0) ? rn_s : -1)
#define dn (((nbv & 4u) != 0) ? dn_s : -1)
#define un (((nbv & 8u) != 0) ? un_s : -1)

  // constants: macros, not variables -- every use reads the constant bank directly (FP32: KConstF floats;
  // FP64: derived from the double fields), so none of them occupies a long-lived register
#define KC(f, d) (sizeof(R) == 4 ? (R)P.cf.c_##f : (R)(d))
#define K_dt KC(dt, P.dt)
#define K_half KC(half, P.half)
#define K_im KC(im, P.inv_m)
#define K_ii KC(ii, P.inv_i)
#define K_mu KC(mu, P.mu)
#define K_lin_tol KC(lin_tol, P.lin_tol)
#define K_ang_tol KC(ang_tol, P.ang_tol)
#define K_max_corr KC(max_corr, P.max_lin_corr)
#define K_baum KC(baum, P.baumgarte)
#define K_lin_f KC(lin_f, P.lin_f)
#define K_ang_f KC(ang_f, P.ang_f)
#define K_sp_sigma KC(sp_sigma, P.sp_sigma)
#define K_sp_boversig KC(sp_boversig, P.sp_beta / P.sp_sigma)
#define K_sp_cfrac KC(sp_cfrac, P.sp_cfrac)
#define K_rope_lower KC(rope_lower, P.rope_lower)
#define K_rope_E KC(rope_E, 1.0 / (2.0 * P.inv_m))
#define K_wa KC(wa, P.weld_anchor)
#define K_wA KC(wA, 2.0 * P.inv_m)
#define K_wh KC(wh, 0.5 * P.inv_i)
#define K_wDA KC(wDA, 1.0 / (2.0 * P.inv_m + 0.5 * P.inv_i * P.weld_anchor * P.weld_anchor))
#define K_wDh KC(wDh, (0.5 * P.inv_i) / ((2.0 * P.inv_m) * (2.0 * P.inv_m + 0.5 * P.inv_i * P.weld_anchor * P.weld_anchor)))
#define K_w_i2i KC(w_i2i, 1.0 / (2.0 * P.inv_i))
#define K_rest_vel KC(rest_vel, P.rest_vel)
#define K_rest_e KC(rest_e, P.rest_e)
#define K_max_trans KC(max_trans, P.max_trans)
#define K_max_trans2 KC(max_trans2, P.max_trans * P.max_trans)
#define K_max_rot KC(max_rot, P.max_rot)
#define K_inv_area2 KC(inv_area2, 0.5 / (P.side_len * P.side_len))
  const R gdt = valid ? KC(gdt, P.g * P.dt) : R(0);
  const int scaff = SCAFF >= 0 ? SCAFF : P.scaff_mask;

  // ---- state ----------------------------------------------------------------------------------
  // per-lane integer origin (world = origin + local); corner masses NW NE SE SW
  R ox, oy;
  R X[4], Y[4], A[4], CS[4], SN[4], VX[4], VY[4], W[4];
  {
    double cx0 = valid ? P.vox_x0[vi] : 0.0, cy0 = valid ? P.vox_y0[vi] : 1.0e4;
    if (P.tplace != nullptr && valid) { cx0 += P.tplace[2 * robot]; cy0 += P.tplace[2 * robot + 1]; }
    ox = (R)rint(cx0); oy = (R)rint(cy0);
    const R lx = (R)(cx0 - (double)ox), ly = (R)(cy0 - (double)oy);
    const R off = KC(mass_off, P.side_len / 2 - P.half);
    const R sxk[4] = {-1, +1, +1, -1}, syk[4] = {+1, +1, -1, -1};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      X[k] = lx + sxk[k] * off; Y[k] = ly + syk[k] * off;
      A[k] = 0; CS[k] = 1; SN[k] = 0; VX[k] = 0; VY[k] = 0; W[k] = 0;
    }
  }
  R rest[3] = {KC(range_rest[0], P.range_rest[0]), KC(range_rest[1], P.range_rest[1]), KC(range_rest[2], P.range_rest[2])};
  R simp[18]; // accumulated spring impulses (warm start); during the velocity loop: impulse + bias/sigma
#pragma unroll
  for (int i = 0; i < 18; i++) simp[i] = 0;
  R rimp[4] = {0, 0, 0, 0};
  R wimp[4][3];
#pragma unroll
  for (int i = 0; i < 4; i++) wimp[i][0] = wimp[i][1] = wimp[i][2] = 0;
  bool touching = false; // any corner mass of this voxel has a manifold (Touch sensor; gates the velocity reloads)
  int par = 0; // parity of the current persistent pool buffer
  int kh = 1;  // terrain polygon hint
  int overflow = 0;

  // ---- controller state --------------------------------------------------------------------------
  R phase = 0;
  if (CTRL == VSR_CTRL_PHASE_SIN) {
    if (valid && !P.time_sin) phase = (R)P.genomes[(size_t)robot * P.genome_len + v];
  } else {
    // K2 genome upload: the warp's robots are consecutive rows -> one contiguous range of doubles,
    // read with 128-bit loads (double2) after a scalar head for 16-byte alignment.
    const int r0 = gw * P.robots_per_warp;
    int nr = P.n_robots - r0;
    nr = nr < 0 ? 0 : (nr > P.robots_per_warp ? P.robots_per_warp : nr);
    const size_t first = (size_t)r0 * P.genome_len, count = (size_t)nr * P.genome_len;
    const double* src = P.genomes + first;
    const size_t head = ((reinterpret_cast<uintptr_t>(src) & 15) && count) ? 1 : 0;
    for (size_t i = lane; i < head; i += 32) wts[i] = (R)src[i];
    const size_t nvec = (count - head) / 2;
    const double2* src2 = reinterpret_cast<const double2*>(src + head);
    for (size_t i = lane; i < nvec; i += 32) {
      const double2 d = __ldg(&src2[i]);
      wts[head + 2 * i] = (R)d.x;
      wts[head + 2 * i + 1] = (R)d.y;
    }
    for (size_t i = head + 2 * nvec + lane; i < count; i += 32) wts[i] = (R)src[i];
    if (CTRL == VSR_CTRL_MLP_DISTRIBUTED_HOMO)
      for (int i = lane; i < 2 * 32 * 4 * P.signals; i += 32) sig[i] = 0;
  }
  if (lane == 0) *poolCnt = 0;
  __syncwarp();
  // cold per-lane state kept in shared memory (touched once per step at most): running sums of the Average
  // sensors, and the first/last observation
  double* savg = reinterpret_cast<double*>(wsm + P.off_cold);           // [n_avg][32]
  int* savg_cnt = reinterpret_cast<int*>(savg + P.n_avg * 32);          // [n_avg][32]
  for (int i = 0; i < P.n_avg; i++) { savg[i * 32 + lane] = 0; savg_cnt[i * 32 + lane] = 0; }
  int sig_par = 0;
  unsigned pcnt[4] = {0u, 0u, 0u, 0u}; // OBS builds: on-chip average-posture counters (16 bits each) of the cells this lane counts
  R ctrl_f = 0;
  const bool frozen = P.frozen != nullptr && valid && ((__ldg(&P.frozen[robot]) >> v) & 1u);

  auto ld4 = [](const R* q) -> V4<R> { return *reinterpret_cast<const V4<R>*>(q); };
  auto st4 = [](R* q, const V4<R>& v) { *reinterpret_cast<V4<R>*>(q) = v; };
  int n_tm = 0; // warp-uniform: corner masses of this warp that have at least one manifold (length of tml)
#if VSR_CACHE_MASS
  unsigned cms = 0; // byte 0: the touching mass (owner lane * 4 + k) this lane serves, bytes 1..3: its manifold slots (massSlots entry)
#define TM_MI(i) ((i) < 32 ? (int)(cms & 0xffu) : (int)tml[i])
#define TM_SLOTS(i, mi) ((i) < 32 ? (cms >> 8) : massSlots[mi])
#else
#define TM_MI(i) ((int)tml[i])
#define TM_SLOTS(i, mi) (massSlots[mi])
#endif
  // per-lane cached ground polygon (the one under the voxel centre), in the lane's local frame
  int pk = 0;
  R px0 = 0, px1 = 0, py0 = 0, py1 = 0, pnx = 0, pny = 1;
  R* stg = reinterpret_cast<R*>(stage);
  // mass -> manifold slots of the current pool parity: byte r = 1 + slot of the rank-r manifold of that mass (0 = none)
  unsigned* massSlots = reinterpret_cast<unsigned*>(wsm + P.off_mslots); // [32 lanes][4 masses]
  // detect-time aliases inside the staging area: CS [32][4] | SN [32][4] | candidate list (128 bytes) | polygon hint [32]
  unsigned char* cand = reinterpret_cast<unsigned char*>(stg + 256);
  int* spk = reinterpret_cast<int*>(cand + 128);
  for (int i = lane; i < 128; i += 32) massSlots[i] = 0;

  // Ground contact detection, TRANSPOSED: the voxel lanes run a broadphase over their 4 corner masses and park the mass
  // poses in shared memory; the surviving (lane, mass) candidates are compacted over the warp and every lane then runs the
  // narrowphase of ONE candidate (lane = candidate, not lane = voxel): the few masses near the ground no longer serialise
  // on their owner lane while the other lanes idle.  The candidate lane also builds everything of the contact constraint
  // that depends on geometry only (Jacobian rows, effective masses, block matrix, local anchors); the restitution bias
  // needs the next step's velocities and is added at the top of the step.  Fills pool parity par^1, then flips par.
  auto detect = [&]() {
    __syncwarp();
    const int np_ = par ^ 1;
    unsigned cmask = 0;
    {
      const double oxd = (double)ox, oyd = (double)oy;
      if (valid) {
        const R cxl = R(0.25) * (X[0] + X[1] + X[2] + X[3]);
        if (!pk || cxl < px0 || cxl > px1) { // (re)load the polygon under the voxel centre
          const double cxw = oxd + (double)cxl;
          const int t0 = P.toff ? __ldg(&P.toff[robot]) : 0, tn = P.toff ? __ldg(&P.toff[robot + 1]) - t0 : P.n_t;
          const double* tx = P.tx + t0;
          const double* ty = P.ty + t0;
          int kk = kh;
          while (kk > 1 && __ldg(&tx[kk - 1]) > cxw) kk--;
          while (kk < tn - 1 && __ldg(&tx[kk]) <= cxw) kk++;
          kh = kk; pk = kk;
          px0 = (R)(__ldg(&tx[kk - 1]) - oxd); px1 = (R)(__ldg(&tx[kk]) - oxd);
          py0 = (R)(__ldg(&ty[kk - 1]) - oyd); py1 = (R)(__ldg(&ty[kk]) - oyd);
          const R ex = px0 - px1, ey = py0 - py1;
          const R il = M<R>::rsqrt_(ex * ex + ey * ey);
          pnx = ey * il; pny = -ex * il;
        }
        // broadphase: a mass well inside the cached polygon's x-range is tested against that polygon's AABB, any other mass against
        // the highest ground within a few units of it (height map of the shared terrain: on narrow polygons -- steps, short hills --
        // only the masses near the ground stay candidates) or, without a map, against the highest terrain vertex.  A filter only: the
        // narrowphase decides, results do not depend on it.  The build specialised for the default robot (flat-terrain benchmarks,
        // every mass takes the `fast` test) leaves the map out: the extra code costs it 2 % through register allocation alone.
        const R pymax = M<R>::max_(py0, py1), ytop = (R)(P.ty_max - oyd);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const R ext = K_half * (M<R>::abs_(CS[k]) + M<R>::abs_(SN[k]));
          const bool fast = X[k] - ext >= px0 + R(kFastMargin) && X[k] + ext <= px1 - R(kFastMargin);
          R ylim = fast ? pymax : ytop;
          if (SENS != 1 && !fast && P.thmax != nullptr) {
            int b = (int)(((float)ox + (float)X[k] - P.thx0) * P.thinv);
            b = b < 0 ? 0 : (b >= P.thn ? P.thn - 1 : b);
            ylim = (R)(__ldg(&P.thmax[b]) - (float)oy);
          }
          if (!(Y[k] - ext > ylim)) cmask |= 1u << k;
        }
      }
    }
    // park the poses (read back after the velocity loop; read by the candidate lanes below)
    {
#pragma unroll
      for (int k = 0; k < 4; k++) { spark[k * 32 + lane] = X[k]; spark[(4 + k) * 32 + lane] = Y[k]; spark[(8 + k) * 32 + lane] = A[k]; }
      spark[12 * 32 + lane] = px0; spark[13 * 32 + lane] = px1; spark[14 * 32 + lane] = py0; spark[15 * 32 + lane] = py1;
      spark[16 * 32 + lane] = pnx; spark[17 * 32 + lane] = pny;
      spark[18 * 32 + lane] = ox; spark[19 * 32 + lane] = oy;
      V4<R> t4;
      t4.x = CS[0]; t4.y = CS[1]; t4.z = CS[2]; t4.w = CS[3]; stage[0 * 32 + lane] = t4;
      t4.x = SN[0]; t4.y = SN[1]; t4.z = SN[2]; t4.w = SN[3]; stage[1 * 32 + lane] = t4;
      spk[lane] = pk;
      uint4 ms = reinterpret_cast<uint4*>(massSlots)[lane]; // masses that are not candidates lose their manifolds
      if (!(cmask & 1u)) ms.x = 0;
      if (!(cmask & 2u)) ms.y = 0;
      if (!(cmask & 4u)) ms.z = 0;
      if (!(cmask & 8u)) ms.w = 0;
      reinterpret_cast<uint4*>(massSlots)[lane] = ms;
    }
    int ncand = 0;
    {
      const unsigned lt = (1u << lane) - 1u;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const bool mine = (cmask >> k) & 1u;
        const unsigned b = __ballot_sync(FULL, mine);
        if (mine) cand[ncand + __popc(b & lt)] = (unsigned char)(lane | (k << 5));
        ncand += __popc(b);
      }
    }
    if (lane == 0) *poolCnt = 0;
    __syncwarp();
#pragma unroll 1
    for (int c0 = 0; c0 < ncand; c0 += 32) { // warp-uniform trip count: the rounds below vote over the whole warp
      const int ci = c0 + lane;
      const bool act = ci < ncand;
      const int c = act ? cand[ci] : 0, ol = c & 31, k = c >> 5, mi = ol * 4 + k;
      const R xk = spark[k * 32 + ol], yk = spark[(4 + k) * 32 + ol], ck = stg[mi], sk = stg[128 + mi];
      const R qx0 = spark[12 * 32 + ol], qx1 = spark[13 * 32 + ol], qy0 = spark[14 * 32 + ol], qy1 = spark[15 * 32 + ol];
      const R qnx = spark[16 * 32 + ol], qny = spark[17 * 32 + ol];
      const double oxd = (double)spark[18 * 32 + ol], oyd = (double)spark[19 * 32 + ol];
      const int pko = spk[ol];
      const unsigned prev = massSlots[mi];
      const R ext = K_half * (M<R>::abs_(ck) + M<R>::abs_(sk));
      const bool fast = xk - ext >= qx0 + R(kFastMargin) && xk + ext <= qx1 - R(kFastMargin);
      double ax0 = 0, ax1 = 0, ay0 = 0, basey = P.base_y;
      const double* tx = P.tx;
      const double* ty = P.ty;
      int kk = pko, kend = pko;
      if (!fast) {
        int tn = P.n_t;
        if (P.toff) { // the owner's own terrain
          const int orobot = gw * P.robots_per_warp + ol / L, t0 = __ldg(&P.toff[orobot]);
          tn = __ldg(&P.toff[orobot + 1]) - t0; tx += t0; ty += t0; basey = __ldg(&P.tbase[orobot]);
        }
        ax0 = oxd + (double)(xk - ext); ax1 = oxd + (double)(xk + ext); ay0 = oyd + (double)(yk - ext);
        // smallest polygon index kk (polygon kk spans tx[kk-1]..tx[kk]) with tx[kk] >= ax0
        while (kk > 1 && __ldg(&tx[kk - 1]) >= ax0) kk--;
        while (kk < tn - 1 && __ldg(&tx[kk]) < ax0) kk++;
        kend = tn - 1;
      }
      int cnt = 0;
      unsigned news = 0;
      bool more = act;
      // Rounds: every lane first walks (cheaply, divergently) to its next polygon that passes the AABB tests, then the warp
      // re-converges and the narrowphase runs ONCE per round for all lanes that found one (a mass near a polygon boundary
      // or on narrow steps meets 2-3 polygons; without the re-convergence the narrowphase body was replayed per lane group).
#pragma unroll 1
      for (;;) {
        bool have = false;
        int kuse = kk;
        R gx0 = qx0, gy0 = qy0, gx1 = qx1, gy1 = qy1;
        if (more) {
          if (fast) { have = true; more = false; }
          else {
            while (kk <= kend) {
              const double x0 = __ldg(&tx[kk - 1]);
              if (x0 > ax1) { kk = kend + 1; break; }
              const double x1 = __ldg(&tx[kk]), y0 = __ldg(&ty[kk - 1]), y1 = __ldg(&ty[kk]);
              const int kc = kk++;
              if (ax1 < x0 || ax0 > x1 || ay0 > fmax(y0, y1)) continue;
              have = true; kuse = kc;
              gx0 = (R)(x0 - oxd); gy0 = (R)(y0 - oyd); gx1 = (R)(x1 - oxd); gy1 = (R)(y1 - oyd);
              break;
            }
            more = have;
          }
        }
        if (!__any_sync(FULL, have)) break;
        if (!have) continue;
        R nx, ny, px[2], py[2], dp[2];
        int id[2];
        int np = -1;
        if (fast) np = collide_box_top<R>(qx1, qy1, qnx, qny, xk, yk, ck, sk, K_half, nx, ny, px, py, dp, id);
        if (np < 0)
          np = collide_box_ground<R>(gx0, gy0, gx1, gy1, (R)(basey - oyd), xk, yk, ck, sk, K_half, K_lin_tol, nx, ny, px, py, dp, id);
        if (!np) continue;
        const int s = atomicAdd(poolCnt, 1);
        if (s >= S) { overflow++; continue; }
        R jn0 = 0, jt0 = 0, jn1 = 0, jt1 = 0;
        // warm start from the previous manifold of the same (mass, polygon)
        const int key = kuse | (k << 14);
        for (int o = 0; o < kMaxMan; o++) {
          const int os = (int)((prev >> (8 * o)) & 0xffu) - 1;
          if (os < 0) break;
          const int ometa = poolM[par * S + os];
          if ((ometa & 0xffff) != key) continue;
          const int onp = (ometa >> 16) & 3;
          const V4<R> oi = ld4(pool + os * kPoolStride + kPI + 4 * par);
          for (int r = 0; r < onp; r++) {
            const int oidr = (ometa >> (18 + 6 * r)) & 63;
            const R ojn = r == 0 ? oi.x : oi.z, ojt = r == 0 ? oi.y : oi.w;
            if (oidr == id[0]) { jn0 = ojn; jt0 = ojt; }
            if (np > 1 && oidr == id[1]) { jn1 = ojn; jt1 = ojt; }
          }
        }
        R* sp = pool + s * kPoolStride;
        V4<R> v4;
        v4.x = nx; v4.y = ny; v4.z = px[0]; v4.w = py[0]; st4(sp, v4);
        v4.x = dp[0]; v4.y = np > 1 ? px[1] : R(0); v4.z = np > 1 ? py[1] : R(0); v4.w = np > 1 ? dp[1] : R(0); st4(sp + 4, v4);
        v4.x = jn0; v4.y = jt0; v4.z = jn1; v4.w = jt1; st4(sp + kPI + 4 * np_, v4);
        // geometry part of the contact constraint (ContactConstraint initialisation): Jacobian rows of a contact point:
        // normal (nx, ny, rn), tangent t = (ny, -nx): (ny, -nx, rt)
        int meta = key | (np << 16) | (id[0] << 18) | ((np > 1 ? id[1] : 0) << 24);
        const R r0x = px[0] - xk, r0y = py[0] - yk, r1x = np > 1 ? px[1] - xk : -xk, r1y = np > 1 ? py[1] - yk : -yk;
        const R rn0 = r0x * ny - r0y * nx, rt0 = -(r0x * nx + r0y * ny);
        const R rn1 = np > 1 ? r1x * ny - r1y * nx : R(0), rt1 = np > 1 ? -(r1x * nx + r1y * ny) : R(0);
        v4.x = rn0; v4.y = rt0; v4.z = M<R>::rcp(K_im + K_ii * rn0 * rn0); v4.w = M<R>::rcp(K_im + K_ii * rt0 * rt0); st4(sp + kPS, v4);
        v4.x = rn1; v4.y = rt1; v4.z = M<R>::rcp(K_im + K_ii * rn1 * rn1); v4.w = M<R>::rcp(K_im + K_ii * rt1 * rt1); st4(sp + kPS + 4, v4);
        R K12 = 0, i11 = 0, i12 = 0, i22 = 0;
        if (np == 2) {
          const R K11 = K_im + K_ii * rn0 * rn0, K22 = K_im + K_ii * rn1 * rn1;
          K12 = K_im + K_ii * rn0 * rn1;
          const R det = K11 * K22 - K12 * K12;
          if (K11 * K11 < R(100) * det) {
            meta |= 1 << 30;
            const R idet = det != 0 ? M<R>::rcp(det) : R(0);
            i11 = K22 * idet; i12 = -K12 * idet; i22 = K11 * idet;
          } else {
            meta = (meta & ~(3 << 16)) | (1 << 16); // ill conditioned: keep the first point only
          }
        }
        v4.x = 0; v4.y = 0; v4.z = K12; v4.w = i11; st4(sp + kPS + 8, v4); // .x .y: restitution bias, set at the top of the step
        v4.x = i12; v4.y = i22; v4.z = 0; v4.w = 0; st4(sp + kPS + 12, v4);
        v4.x = ck * r0x + sk * r0y; v4.y = -sk * r0x + ck * r0y; v4.z = ck * r1x + sk * r1y; v4.w = -sk * r1x + ck * r1y; st4(sp + kPS + 16, v4);
        poolM[np_ * S + s] = meta;
        news |= (unsigned)(s + 1) << (8 * cnt);
        cnt++;
        if (cnt >= kMaxMan) more = false;
      }
      if (act) massSlots[mi] = news;
    }
    par ^= 1;
    __syncwarp();
    // list of the touching masses: the contact phases of the step run lane = touching mass
    {
      const uint4 ms = reinterpret_cast<const uint4*>(massSlots)[lane];
      const unsigned w4[4] = {ms.x, ms.y, ms.z, ms.w};
      touching = valid && (ms.x | ms.y | ms.z | ms.w) != 0;
      const unsigned lt = (1u << lane) - 1u;
      int n = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const bool mine = valid && w4[k] != 0;
        const unsigned b = __ballot_sync(FULL, mine);
        if (mine) tml[n + __popc(b & lt)] = (unsigned char)(lane * 4 + k);
        n += __popc(b);
      }
      n_tm = n;
    }
    __syncwarp();
#if VSR_CACHE_MASS
    // the touching mass this lane serves in every contact phase of the coming step (restitution / warm start, 10 velocity iterations,
    // the position iterations) and its manifold slots, in ONE register: two dependent shared-memory loads less on each of those chains
    cms = lane < n_tm ? (unsigned)tml[lane] : 0u;
    cms |= lane < n_tm ? massSlots[cms] << 8 : 0u;
#endif
  };

  detect(); // World.step's first-step detect()

  double t = 0.0;
#if VSR_SOFT_SYNC > 0
  volatile int* sphase = reinterpret_cast<volatile int*>(smem_raw + (size_t)cta_warps * P.smem_warp_stride); // [cta_warps] boundaries passed
  int my_phase = 0;
  if (lane == 0) sphase[warp] = 0;
  __syncthreads();
#endif
  // =================================================================================================
#pragma unroll 1
  for (int step = 1; step <= P.n_steps; step++) {
#if VSR_SOFT_SYNC > 0
    // bounded-skew alignment: a warp may run at most VSR_SOFT_SYNC phase boundaries ahead of the slowest warp of its CTA, so the
    // warps still share (almost) one instruction stream, but a warp that needs extra position iterations or narrowphase rounds in
    // one step no longer stalls the other nine: the slack is paid back when another warp straggles.  Pure timing: no data crosses
    // warps, so the volatile counters need no fence.
#define PHASE_SYNC(bit)                                                                     \
    if ((SCAFF >= 0 ? kSyncMaskDefault : P.sync_mask) & (bit)) {                              \
      my_phase++;                                                                           \
      __syncwarp();                                                                         \
      if (lane == 0) sphase[warp] = my_phase;                                               \
      for (;;) {                                                                            \
        const int p_ = lane < cta_warps ? sphase[lane] : 0x7fffffff;                        \
        if (__reduce_min_sync(FULL, p_) >= my_phase - VSR_SOFT_SYNC) break;                 \
      }                                                                                     \
    }
#else
#define PHASE_SYNC(bit) if ((SCAFF >= 0 ? kSyncMaskDefault : P.sync_mask) & (bit)) __syncthreads(); /* re-align the CTA's warps: shared instruction stream */
#endif
    PHASE_SYNC(1)
    t += P.dt;
    // ---- integrate velocities: gravity, then damping ---------------------------------------------
#pragma unroll
    for (int k = 0; k < 4; k++) {
      VY[k] += gdt;
      VX[k] *= K_lin_f; VY[k] *= K_lin_f; W[k] *= K_ang_f;
    }
    // ---- contacts: restitution bias (from the pre-warm-start velocities) + warm start, lane = touching mass -----
    // (the geometry part of the constraint was built by detect(); cold state X Y A / polygon / origin is parked there too)
    if (n_tm > 0) {
      {
        V4<R> t4;
        t4.x = VX[0]; t4.y = VX[1]; t4.z = VX[2]; t4.w = VX[3]; stage[0 * 32 + lane] = t4;
        t4.x = VY[0]; t4.y = VY[1]; t4.z = VY[2]; t4.w = VY[3]; stage[1 * 32 + lane] = t4;
        t4.x = W[0]; t4.y = W[1]; t4.z = W[2]; t4.w = W[3]; stage[2 * 32 + lane] = t4;
      }
      __syncwarp();
#pragma unroll 1
      for (int i = lane; i < n_tm; i += 32) {
        const int mi = TM_MI(i);
        const unsigned slots = TM_SLOTS(i, mi);
        const R vx0 = stg[mi], vy0 = stg[128 + mi], w0 = stg[256 + mi]; // pre-warm-start velocity of this mass
        R vxk = vx0, vyk = vy0, wk = w0;
#pragma unroll 1
        for (unsigned sl = slots; sl != 0; sl >>= 8) { // slot bytes are packed from rank 0 upwards
          const int s = (int)(sl & 0xffu) - 1;
          const int meta = poolM[par * S + s];
          const int np = (meta >> 16) & 3;
          R* sp = pool + s * kPoolStride;
          const V2<R> nn = *reinterpret_cast<const V2<R>*>(sp);
          const V4<R> imp = ld4(sp + kPI + 4 * par);
          const V2<R> j0 = *reinterpret_cast<const V2<R>*>(sp + kPS), j1 = *reinterpret_cast<const V2<R>*>(sp + kPS + 4);
          const R nx = nn.x, ny = nn.y;
          {
            const R vn = nx * vx0 + ny * vy0;
            const R vn0 = vn + j0.x * w0, vn1 = vn + j1.x * w0;
            V2<R> vb;
            vb.x = vn0 < -K_rest_vel ? -K_rest_e * vn0 : R(0);
            vb.y = (np > 1 && vn1 < -K_rest_vel) ? -K_rest_e * vn1 : R(0);
            *reinterpret_cast<V2<R>*>(sp + kPS + 8) = vb;
          }
          {
            const R Jx = nx * imp.x + ny * imp.y, Jy = ny * imp.x - nx * imp.y;
            vxk += K_im * Jx; vyk += K_im * Jy; wk += K_ii * (j0.x * imp.x + j0.y * imp.y);
          }
          if (np > 1) {
            const R Jx = nx * imp.z + ny * imp.w, Jy = ny * imp.z - nx * imp.w;
            vxk += K_im * Jx; vyk += K_im * Jy; wk += K_ii * (j1.x * imp.z + j1.y * imp.w);
          }
        }
        stg[mi] = vxk; stg[128 + mi] = vyk; stg[256 + mi] = wk;
      }
      __syncwarp();
      if (touching) {
        V4<R> t4;
        t4 = stage[0 * 32 + lane]; VX[0] = t4.x; VX[1] = t4.y; VX[2] = t4.z; VX[3] = t4.w;
        t4 = stage[1 * 32 + lane]; VY[0] = t4.x; VY[1] = t4.y; VY[2] = t4.z; VY[3] = t4.w;
        t4 = stage[2 * 32 + lane]; W[0] = t4.x; W[1] = t4.y; W[2] = t4.z; W[3] = t4.w;
      }
      __syncwarp(); // the staging area is reused by the velocity loop
    }
    // ---- springs: initialise + warm start ---------------------------------------------------------
    PHASE_SYNC(64)
    R snx[18], sny[18], sE16 = 0, sE17 = 0; // entries 12..15 unused (those normals live in shared memory)
    V4<R> nrm4;
    V4<R> arm4;
    V2<R> eff2;
    {
      R HC[4], HS[4];
#pragma unroll
      for (int k = 0; k < 4; k++) { HC[k] = K_half * CS[k]; HS[k] = K_half * SN[k]; }
#define SPRING_INIT(i, a, b, sx1, sy1, sx2, sy2, kind, bit)                                      \
  if (scaff & bit) {                                                                             \
    const R r1x = (sx1)*HC[a] - (sy1)*HS[a], r1y = (sx1)*HS[a] + (sy1)*HC[a];                    \
    const R r2x = (sx2)*HC[b] - (sy2)*HS[b], r2y = (sx2)*HS[b] + (sy2)*HC[b];                    \
    R nx = (X[a] + r1x) - (X[b] + r2x), ny = (Y[a] + r1y) - (Y[b] + r2y);                        \
    const R l2 = nx * nx + ny * ny;                                                              \
    const R il = M<R>::rsqrt_(l2);                                                               \
    const R len = l2 * il;                                                                       \
    if (!(len >= K_lin_tol)) { nx = 0; ny = 0; } else { nx *= il; ny *= il; }                      \
    const R cr1 = r1x * ny - r1y * nx, cr2 = r2x * ny - r2y * nx;                                \
    const R E = K_sp_cfrac * M<R>::rcp(K_im + K_ii * cr1 * cr1 + K_im + K_ii * cr2 * cr2);                 \
    if ((i) < 12 || (i) >= 16) { snx[i] = nx; sny[i] = ny; }                                     \
    else if (((i) & 1) == 0) { nrm4.x = nx; nrm4.y = ny; }                                       \
    else { nrm4.z = nx; nrm4.w = ny; snrm[(((i) - 12) >> 1) * 32 + lane] = nrm4; }               \
    if ((i) < 16) {                                                                              \
      if (((i) & 1) == 0) { arm4.x = cr1; arm4.y = cr2; eff2.x = E; }                            \
      else { arm4.z = cr1; arm4.w = cr2; eff2.y = E; sarm[((i) >> 1) * 32 + lane] = arm4; seff[((i) >> 1) * 32 + lane] = eff2; } \
    } else if ((i) == 16) sE16 = E; else sE17 = E;                                               \
    const R m_ = simp[i] * K_im, q_ = simp[i] * K_ii;                                                \
    VX[a] += nx * m_; VY[a] += ny * m_; W[a] += q_ * cr1;                                        \
    VX[b] -= nx * m_; VY[b] -= ny * m_; W[b] -= q_ * cr2;                                        \
    const R bs = E * ((len - rest[kind]) * K_sp_boversig); /* bias / sigma */                      \
    sbias[i * 32 + lane] = bs;                                                                   \
    simp[i] += bs;                                                                               \
  }
      VSR_SPRINGS(SPRING_INIT)
#undef SPRING_INIT
    }
    // ---- ropes: initialise + warm start -------------------------------------------------------------
    bool ract[4];
    bool any_rope = false;
    V4<R> rn4;
#define ROPE_INIT(i, a, b)                                                                  \
  {                                                                                         \
    R nx = X[a] - X[b], ny = Y[a] - Y[b];                                                   \
    const R l2 = nx * nx + ny * ny;                                                         \
    const R il = M<R>::rsqrt_(l2);                                                          \
    const R len = l2 * il;                                                                  \
    if (!(len >= K_lin_tol)) { nx = 0; ny = 0; } else { nx *= il; ny *= il; }                 \
    ract[i] = P.limit_contraction && len < K_rope_lower;                                      \
    if (((i) & 1) == 0) { rn4.x = nx; rn4.y = ny; } else { rn4.z = nx; rn4.w = ny; srope[((i) >> 1) * 32 + lane] = rn4; } \
    if (!ract[i]) rimp[i] = 0;                                                              \
    const R m_ = rimp[i] * K_im;                                                              \
    VX[a] += nx * m_; VY[a] += ny * m_; VX[b] -= nx * m_; VY[b] -= ny * m_;                 \
    any_rope |= ract[i];                                                                    \
  }
    VSR_ROPES(ROPE_INIT)
#undef ROPE_INIT
    const bool warp_rope = __any_sync(FULL, any_rope);
    // ---- welds: r2 of the partner mass, warm start ----------------------------------------------------
    PHASE_SYNC(128)
    // owner welds: 0 = (my NW, left NE) 1 = (my SW, left SE) 2 = (my SW, below NW) 3 = (my SE, below NE)
    // local anchor 2 = body1 - body2 at assembly: horizontal (+m, 0), vertical (0, +m)
    R wr2x[4], wr2y[4];
    {
      const R c1 = __shfl_sync(FULL, CS[1], ln_s), s1 = __shfl_sync(FULL, SN[1], ln_s);
      const R c2 = __shfl_sync(FULL, CS[2], ln_s), s2 = __shfl_sync(FULL, SN[2], ln_s);
      wr2x[0] = K_wa * c1; wr2y[0] = K_wa * s1; wr2x[1] = K_wa * c2; wr2y[1] = K_wa * s2;
      const R c0 = __shfl_sync(FULL, CS[0], dn_s), s0 = __shfl_sync(FULL, SN[0], dn_s);
      const R c1d = __shfl_sync(FULL, CS[1], dn_s), s1d = __shfl_sync(FULL, SN[1], dn_s);
      wr2x[2] = -K_wa * s0; wr2y[2] = K_wa * c0; wr2x[3] = -K_wa * s1d; wr2y[3] = K_wa * c1d;
    }
    // apply impulse l (3) of owner weld to body1 (mine) and return the partner's delta
#define WELD_APPLY(wi, k1, lx_, ly_, lz_, dvx, dvy, dw)                   \
  {                                                                      \
    VX[k1] += K_im * (lx_); VY[k1] += K_im * (ly_); W[k1] += K_ii * (lz_);     \
    dvx = -K_im * (lx_); dvy = -K_im * (ly_);                                \
    dw = -K_ii * (wr2x[wi] * (ly_) - wr2y[wi] * (lx_) + (lz_));            \
  }
    {
      // horizontal round warm start
      R d0x = 0, d0y = 0, d0w = 0, d1x = 0, d1y = 0, d1w = 0;
      if (ln >= 0) {
        WELD_APPLY(0, 0, wimp[0][0], wimp[0][1], wimp[0][2], d0x, d0y, d0w)
        WELD_APPLY(1, 3, wimp[1][0], wimp[1][1], wimp[1][2], d1x, d1y, d1w)
      }
      const R a0 = __shfl_sync(FULL, d0x, rn_s), a1 = __shfl_sync(FULL, d0y, rn_s), a2 = __shfl_sync(FULL, d0w, rn_s);
      const R b0 = __shfl_sync(FULL, d1x, rn_s), b1 = __shfl_sync(FULL, d1y, rn_s), b2 = __shfl_sync(FULL, d1w, rn_s);
      if (rn >= 0) { VX[1] += a0; VY[1] += a1; W[1] += a2; VX[2] += b0; VY[2] += b1; W[2] += b2; }
      // vertical round warm start
      d0x = d0y = d0w = d1x = d1y = d1w = 0;
      if (dn >= 0) {
        WELD_APPLY(2, 3, wimp[2][0], wimp[2][1], wimp[2][2], d0x, d0y, d0w)
        WELD_APPLY(3, 2, wimp[3][0], wimp[3][1], wimp[3][2], d1x, d1y, d1w)
      }
      const R e0 = __shfl_sync(FULL, d0x, un_s), e1 = __shfl_sync(FULL, d0y, un_s), e2 = __shfl_sync(FULL, d0w, un_s);
      const R f0 = __shfl_sync(FULL, d1x, un_s), f1 = __shfl_sync(FULL, d1y, un_s), f2 = __shfl_sync(FULL, d1w, un_s);
      if (un >= 0) { VX[0] += e0; VY[0] += e1; W[0] += e2; VX[1] += f0; VY[1] += f1; W[1] += f2; }
    }
    // solve K l = -C for a weld with r1 = 0 and partner arm (r2x, r2y)
    // (the 2x2 block left after eliminating the angular row is D * (A I + h r2 r2^T), so no matrix entries are formed)
#define WELD_SOLVE(r2x, r2y, Cx, Cy, Cz, lx_, ly_, lz_)                                  \
  {                                                                                     \
    const R bx_ = -(Cx) - (r2y) * (Cz) * R(0.5), by_ = -(Cy) + (r2x) * (Cz) * R(0.5);   \
    const R hd_ = K_wDh * ((r2x) * bx_ + (r2y) * by_);                                  \
    lx_ = K_wDA * bx_ + hd_ * (r2x);                                                    \
    ly_ = K_wDA * by_ + hd_ * (r2y);                                                    \
    lz_ = -(Cz) * K_w_i2i + ((r2y)*lx_ - (r2x)*ly_) * R(0.5);                             \
  }

    // ---- velocity iterations ------------------------------------------------------------------------
    PHASE_SYNC(2)
#pragma unroll 1
    for (int it = 0; it < P.vel_iters; it++) {
      PHASE_SYNC(4)
      // springs: lambda' = E*Jv + sigma*(impulse + bias/sigma); impulse' -= lambda'  (DESIGN.md 4.2).
      // Two body-disjoint springs are solved side by side (explicit ILP 2: the lane's dependency chain is
      // the bottleneck, not the FMA pipe); one LDS.128 brings both springs' lever arms.
#define SPRING_SOLVE2(i, a, b, j, c, d, bit)                                                     \
  if (scaff & bit) {                                                                            \
    const V4<R> arm = sarm[((i) >> 1) * 32 + lane];                                             \
    const V2<R> eff = seff[((i) >> 1) * 32 + lane];                                             \
    R nxi, nyi, nxj, nyj;                                                                       \
    if ((i) >= 12) { const V4<R> n4 = snrm[(((i) - 12) >> 1) * 32 + lane]; nxi = n4.x; nyi = n4.y; nxj = n4.z; nyj = n4.w; } \
    else { nxi = snx[(i) < 12 ? (i) : 0]; nyi = sny[(i) < 12 ? (i) : 0]; nxj = snx[(j) < 12 ? (j) : 0]; nyj = sny[(j) < 12 ? (j) : 0]; } \
    R jv0 = nxi * (VX[a] - VX[b]), jv1 = nxj * (VX[c] - VX[d]);                                 \
    jv0 += nyi * (VY[a] - VY[b]); jv1 += nyj * (VY[c] - VY[d]);                                 \
    jv0 += arm.x * W[a]; jv1 += arm.z * W[c];                                                   \
    jv0 -= arm.y * W[b]; jv1 -= arm.w * W[d];                                                   \
    const R l0 = eff.x * jv0 + K_sp_sigma * simp[i], l1 = eff.y * jv1 + K_sp_sigma * simp[j];   \
    simp[i] -= l0; simp[j] -= l1;                                                               \
    const R m0 = l0 * K_im, m1 = l1 * K_im, q0 = l0 * K_ii, q1 = l1 * K_ii;                     \
    VX[a] -= nxi * m0; VX[c] -= nxj * m1; VY[a] -= nyi * m0; VY[c] -= nyj * m1;                 \
    VX[b] += nxi * m0; VX[d] += nxj * m1; VY[b] += nyi * m0; VY[d] += nyj * m1;                 \
    W[a] -= q0 * arm.x; W[c] -= q1 * arm.z; W[b] += q0 * arm.y; W[d] += q1 * arm.w;             \
  }
      VSR_SPRING_PAIRS(SPRING_SOLVE2)
#undef SPRING_SOLVE2
      if (scaff & 8) { // central cross (0,2) and (1,3): anchors at the centres, r x n = 0
        R jv0 = snx[16] * (VX[0] - VX[2]), jv1 = snx[17] * (VX[1] - VX[3]);
        jv0 += sny[16] * (VY[0] - VY[2]); jv1 += sny[17] * (VY[1] - VY[3]);
        const R l0 = sE16 * jv0 + K_sp_sigma * simp[16], l1 = sE17 * jv1 + K_sp_sigma * simp[17];
        simp[16] -= l0; simp[17] -= l1;
        const R m0 = l0 * K_im, m1 = l1 * K_im;
        VX[0] -= snx[16] * m0; VX[1] -= snx[17] * m1; VY[0] -= sny[16] * m0; VY[1] -= sny[17] * m1;
        VX[2] += snx[16] * m0; VX[3] += snx[17] * m1; VY[2] += sny[16] * m0; VY[3] += sny[17] * m1;
      }
      if (warp_rope) {
      // branch-free: an inactive rope keeps a zero impulse (few lanes have an active rope: a branch per rope diverges)
#define ROPE_SOLVE(i, a, b)                                                        \
  {                                                                                \
    const V2<R> n_ = *reinterpret_cast<const V2<R>*>(reinterpret_cast<const R*>(srope + ((i) >> 1) * 32 + lane) + ((i) & 1) * 2); \
    const R jv = n_.x * (VX[a] - VX[b]) + n_.y * (VY[a] - VY[b]);                  \
    const R old = rimp[i];                                                         \
    const R nw = ract[i] ? M<R>::max_(old - K_rope_E * jv, R(0)) : R(0);           \
    rimp[i] = nw;                                                                  \
    const R m_ = (nw - old) * K_im;                                                \
    VX[a] += n_.x * m_; VY[a] += n_.y * m_; VX[b] -= n_.x * m_; VY[b] -= n_.y * m_;         \
  }
        VSR_ROPES(ROPE_SOLVE)
#undef ROPE_SOLVE
      }
#ifndef VSR_PROBE_NO_WELD_LOOP
      // welds, horizontal round: partner = left neighbour's NE(1) / SE(2)
      {
        const R p0x = __shfl_sync(FULL, VX[1], ln_s), p0y = __shfl_sync(FULL, VY[1], ln_s), p0w = __shfl_sync(FULL, W[1], ln_s);
        const R p1x = __shfl_sync(FULL, VX[2], ln_s), p1y = __shfl_sync(FULL, VY[2], ln_s), p1w = __shfl_sync(FULL, W[2], ln_s);
        R d0x = 0, d0y = 0, d0w = 0, d1x = 0, d1y = 0, d1w = 0;
        if (ln >= 0) {
          R lx_, ly_, lz_;
          WELD_SOLVE(wr2x[0], wr2y[0], VX[0] - (p0x - p0w * wr2y[0]), VY[0] - (p0y + p0w * wr2x[0]), W[0] - p0w, lx_, ly_, lz_)
          wimp[0][0] += lx_; wimp[0][1] += ly_; wimp[0][2] += lz_;
          WELD_APPLY(0, 0, lx_, ly_, lz_, d0x, d0y, d0w)
          WELD_SOLVE(wr2x[1], wr2y[1], VX[3] - (p1x - p1w * wr2y[1]), VY[3] - (p1y + p1w * wr2x[1]), W[3] - p1w, lx_, ly_, lz_)
          wimp[1][0] += lx_; wimp[1][1] += ly_; wimp[1][2] += lz_;
          WELD_APPLY(1, 3, lx_, ly_, lz_, d1x, d1y, d1w)
        }
        const R a0 = __shfl_sync(FULL, d0x, rn_s), a1 = __shfl_sync(FULL, d0y, rn_s), a2 = __shfl_sync(FULL, d0w, rn_s);
        const R b0 = __shfl_sync(FULL, d1x, rn_s), b1 = __shfl_sync(FULL, d1y, rn_s), b2 = __shfl_sync(FULL, d1w, rn_s);
        if (rn >= 0) { VX[1] += a0; VY[1] += a1; W[1] += a2; VX[2] += b0; VY[2] += b1; W[2] += b2; }
      }
      // welds, vertical round: partner = below neighbour's NW(0) / NE(1)
      {
        const R p0x = __shfl_sync(FULL, VX[0], dn_s), p0y = __shfl_sync(FULL, VY[0], dn_s), p0w = __shfl_sync(FULL, W[0], dn_s);
        const R p1x = __shfl_sync(FULL, VX[1], dn_s), p1y = __shfl_sync(FULL, VY[1], dn_s), p1w = __shfl_sync(FULL, W[1], dn_s);
        R d0x = 0, d0y = 0, d0w = 0, d1x = 0, d1y = 0, d1w = 0;
        if (dn >= 0) {
          R lx_, ly_, lz_;
          WELD_SOLVE(wr2x[2], wr2y[2], VX[3] - (p0x - p0w * wr2y[2]), VY[3] - (p0y + p0w * wr2x[2]), W[3] - p0w, lx_, ly_, lz_)
          wimp[2][0] += lx_; wimp[2][1] += ly_; wimp[2][2] += lz_;
          WELD_APPLY(2, 3, lx_, ly_, lz_, d0x, d0y, d0w)
          WELD_SOLVE(wr2x[3], wr2y[3], VX[2] - (p1x - p1w * wr2y[3]), VY[2] - (p1y + p1w * wr2x[3]), W[2] - p1w, lx_, ly_, lz_)
          wimp[3][0] += lx_; wimp[3][1] += ly_; wimp[3][2] += lz_;
          WELD_APPLY(3, 2, lx_, ly_, lz_, d1x, d1y, d1w)
        }
        const R e0 = __shfl_sync(FULL, d0x, un_s), e1 = __shfl_sync(FULL, d0y, un_s), e2 = __shfl_sync(FULL, d0w, un_s);
        const R f0 = __shfl_sync(FULL, d1x, un_s), f1 = __shfl_sync(FULL, d1y, un_s), f2 = __shfl_sync(FULL, d1w, un_s);
        if (un >= 0) { VX[0] += e0; VY[0] += e1; W[0] += e2; VX[1] += f0; VY[1] += f1; W[1] += f2; }
      }
#endif
#ifndef VSR_PROBE_NO_CONTACT_LOOP
      // contacts, transposed: every lane solves the manifolds of ONE touching corner mass, in rank (polygon) order; the
      // masses' velocities travel through shared memory.  The ground is static, so different masses are independent and
      // each body still sees its constraints in the canonical order.
      if (n_tm > 0) {
        {
          V4<R> t4;
          t4.x = VX[0]; t4.y = VX[1]; t4.z = VX[2]; t4.w = VX[3]; stage[0 * 32 + lane] = t4;
          t4.x = VY[0]; t4.y = VY[1]; t4.z = VY[2]; t4.w = VY[3]; stage[1 * 32 + lane] = t4;
          t4.x = W[0]; t4.y = W[1]; t4.z = W[2]; t4.w = W[3]; stage[2 * 32 + lane] = t4;
        }
        __syncwarp();
#pragma unroll 1
        for (int i = lane; i < n_tm; i += 32) {
          const int mi = TM_MI(i); // corner mass = owner lane * 4 + k
          const unsigned slots = TM_SLOTS(i, mi);
          R vxk = stg[mi], vyk = stg[128 + mi], wk = stg[256 + mi];
#pragma unroll 1
          for (unsigned sl = slots; sl != 0; sl >>= 8) {
            const int s = (int)(sl & 0xffu) - 1;
            const int meta = poolM[par * S + s];
            const int np = (meta >> 16) & 3;
            R* sp = pool + s * kPoolStride;
            const V2<R> nn = *reinterpret_cast<const V2<R>*>(sp);
            V4<R> imp = ld4(sp + kPI + 4 * par);
            const V4<R> j0 = ld4(sp + kPS), j1 = ld4(sp + kPS + 4), cc = ld4(sp + kPS + 8);
            const R nx = nn.x, ny = nn.y;
            // all rows of a manifold share n and t = (ny, -nx): solve in (normal, tangent, angular) velocity coordinates
            R vn_ = nx * vxk + ny * vyk, vt_ = ny * vxk - nx * vyk;
            // friction rows first, clamped by mu * (accumulated normal impulse)
            {
              R l_ = -j0.w * (vt_ + j0.y * wk);
              const R maxf = K_mu * imp.x;
              const R nj = M<R>::max_(-maxf, M<R>::min_(imp.y + l_, maxf));
              l_ = nj - imp.y; imp.y = nj;
              vt_ += K_im * l_; wk += (K_ii * l_) * j0.y;
            }
            if (np > 1) {
              R l_ = -j1.w * (vt_ + j1.y * wk);
              const R maxf = K_mu * imp.z;
              const R nj = M<R>::max_(-maxf, M<R>::min_(imp.w + l_, maxf));
              l_ = nj - imp.w; imp.w = nj;
              vt_ += K_im * l_; wk += (K_ii * l_) * j1.y;
            }
            if (!(meta & (1 << 30))) { // normal rows one by one (single point, or ill-conditioned pair reduced to one)
              R l_ = -j0.z * (vn_ + j0.x * wk - cc.x);
              const R nj = M<R>::max_(imp.x + l_, R(0));
              l_ = nj - imp.x; imp.x = nj;
              vn_ += K_im * l_; wk += (K_ii * l_) * j0.x;
            } else { // 2-point block solver
              const V2<R> ik = *reinterpret_cast<const V2<R>*>(sp + kPS + 12);
              const R a1 = imp.x, a2 = imp.z, K12 = cc.z;
              const R vn1 = vn_ + j0.x * wk, vn2 = vn_ + j1.x * wk;
              const R K11 = K_im + K_ii * j0.x * j0.x, K22 = K_im + K_ii * j1.x * j1.x;
              const R b1 = vn1 - cc.x - (K11 * a1 + K12 * a2);
              const R b2 = vn2 - cc.y - (K12 * a1 + K22 * a2);
              R x1 = -(cc.w * b1 + ik.x * b2), x2 = -(ik.x * b1 + ik.y * b2);
              if (!(x1 >= 0 && x2 >= 0)) {
                x1 = -b1 * j0.z; x2 = 0;
                if (!(x1 >= 0 && K12 * x1 + b2 >= 0)) {
                  x1 = 0; x2 = -b2 * j1.z;
                  if (!(x2 >= 0 && K12 * x2 + b1 >= 0)) {
                    x1 = 0; x2 = 0;
                    if (!(b1 >= 0 && b2 >= 0)) { x1 = a1; x2 = a2; }
                  }
                }
              }
              const R d1 = x1 - a1, d2 = x2 - a2;
              imp.x = x1; imp.z = x2;
              vn_ += K_im * (d1 + d2); wk += K_ii * (j0.x * d1 + j1.x * d2);
            }
            vxk = nx * vn_ + ny * vt_; vyk = ny * vn_ - nx * vt_;
            st4(sp + kPI + 4 * par, imp);
          }
          stg[mi] = vxk; stg[128 + mi] = vyk; stg[256 + mi] = wk;
        }
        __syncwarp();
        if (touching) {
          V4<R> t4;
          t4 = stage[0 * 32 + lane]; VX[0] = t4.x; VX[1] = t4.y; VX[2] = t4.z; VX[3] = t4.w;
          t4 = stage[1 * 32 + lane]; VY[0] = t4.x; VY[1] = t4.y; VY[2] = t4.z; VY[3] = t4.w;
          t4 = stage[2 * 32 + lane]; W[0] = t4.x; W[1] = t4.y; W[2] = t4.z; W[3] = t4.w;
        }
      }
#endif
    } // velocity iterations
    // spring accumulators back to pure impulses (warm start of the next step)
    PHASE_SYNC(256)
#define SPRING_FIN(i, a, b, sx1, sy1, sx2, sy2, kind, bit) if (scaff & bit) simp[i] -= sbias[i * 32 + lane];
    VSR_SPRINGS(SPRING_FIN)
#undef SPRING_FIN

    // ---- integrate positions (per-step clamps) -----------------------------------------------------
    PHASE_SYNC(8)
    {
      asm volatile("" ::: "memory");
#pragma unroll
      for (int k = 0; k < 4; k++) { X[k] = spark[(k) * 32 + lane]; Y[k] = spark[(4 + k) * 32 + lane]; A[k] = spark[(8 + k) * 32 + lane]; }
      px0 = spark[12 * 32 + lane]; px1 = spark[13 * 32 + lane]; py0 = spark[14 * 32 + lane]; py1 = spark[15 * 32 + lane];
      pnx = spark[16 * 32 + lane]; pny = spark[17 * 32 + lane];
      ox = spark[18 * 32 + lane]; oy = spark[19 * 32 + lane];
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      R txx = VX[k] * K_dt, tyy = VY[k] * K_dt, r = W[k] * K_dt;
      const R t2 = txx * txx + tyy * tyy;
      if (t2 > K_max_trans2) {
        const R ratio = K_max_trans * M<R>::rsqrt_(t2);
        VX[k] *= ratio; VY[k] *= ratio; txx = VX[k] * K_dt; tyy = VY[k] * K_dt;
      }
      if (M<R>::abs_(r) > K_max_rot) {
        const R ratio = K_max_rot * M<R>::rcp(M<R>::abs_(r));
        W[k] *= ratio; r = W[k] * K_dt;
      }
      X[k] += txx; Y[k] += tyy; A[k] += r;
      M<R>::sincos_(A[k], &SN[k], &CS[k]);
    }

    // ---- position iterations: contacts, ropes, welds (H then V); per-robot early exit ----------------
    {
      // origin offsets of the weld partners (integers, exact)
      const R lox = __shfl_sync(FULL, ox, ln_s) - ox, loy = __shfl_sync(FULL, oy, ln_s) - oy;
      const R dox = __shfl_sync(FULL, ox, dn_s) - ox, doy = __shfl_sync(FULL, oy, dn_s) - oy;
      const unsigned robot_mask = lane_used ? (((L >= 32) ? FULL : ((1u << L) - 1u)) << base_lane) : (1u << lane);
      bool active = true;
      bool bigrot = false; // FP32: some rotation step of this iteration left the small-angle polynomial's range (see M<float>::rotate_fast)
#pragma unroll 1
      for (int it = 0; it < P.pos_iters; it++) {
        bool ok = true;
        // contacts, lane = touching mass: the masses' poses travel through shared memory (X Y A in the staging area, CS SN
        // in the spring-bias area, which is dead outside the velocity loop)
        if (n_tm > 0) {
          V4<R>* pcs = reinterpret_cast<V4<R>*>(sbias);
          R* pc = sbias;
          {
            V4<R> t4;
            t4.x = X[0]; t4.y = X[1]; t4.z = X[2]; t4.w = X[3]; stage[0 * 32 + lane] = t4;
            t4.x = Y[0]; t4.y = Y[1]; t4.z = Y[2]; t4.w = Y[3]; stage[1 * 32 + lane] = t4;
            t4.x = A[0]; t4.y = A[1]; t4.z = A[2]; t4.w = A[3]; stage[2 * 32 + lane] = t4;
            t4.x = CS[0]; t4.y = CS[1]; t4.z = CS[2]; t4.w = CS[3]; pcs[0 * 32 + lane] = t4;
            t4.x = SN[0]; t4.y = SN[1]; t4.z = SN[2]; t4.w = SN[3]; pcs[1 * 32 + lane] = t4;
          }
          const unsigned actm = __ballot_sync(FULL, active);
          __syncwarp();
          unsigned mybad = 0;
          unsigned bigm = 0; // owner lanes of the masses whose rotation step left the small-angle range (FP32, VSR_POS_V2)
#pragma unroll 1
          for (int i = lane; i < n_tm; i += 32) {
            const int mi = TM_MI(i), ol = mi >> 2;
            if (!((actm >> ol) & 1u)) continue;
            const unsigned slots = TM_SLOTS(i, mi);
            R xk = stg[mi], yk = stg[128 + mi], ak = stg[256 + mi], ck = pc[mi], sk = pc[128 + mi];
#pragma unroll 1
            for (unsigned sl = slots; sl != 0; sl >>= 8) {
              const int s = (int)(sl & 0xffu) - 1;
              const int np = (poolM[par * S + s] >> 16) & 3;
              const R* sp = pool + s * kPoolStride;
              const V4<R> pa = ld4(sp), pb = ld4(sp + 4), pl = ld4(sp + kPS + 16);
              const R nx = pa.x, ny = pa.y;
#if VSR_POS_V2
              // the (at most two) points of a manifold, unrolled: the second one is masked by np instead of looped over
#pragma unroll
              for (int q = 0; q < 2; q++) {
                const R lx = q == 0 ? pl.x : pl.z, ly = q == 0 ? pl.y : pl.w;
                const R cpx = q == 0 ? pa.z : pb.y, cpy = q == 0 ? pa.w : pb.z, cdp = q == 0 ? pb.x : pb.w;
                const bool on = q < np;
                const R rx = ck * lx - sk * ly, ry = sk * lx + ck * ly;
                const R sep = nx * (xk + rx - cpx) + ny * (yk + ry - cpy) - cdp;
                if (on && sep < -R(3) * K_lin_tol) mybad |= 1u << ol;
                R C = K_baum * (sep + K_lin_tol);
                C = M<R>::min_(C, R(0));
                C = M<R>::max_(C, -K_max_corr);
                const R rn_ = rx * ny - ry * nx;
                const R l_ = on ? -C * M<R>::rcp(K_im + K_ii * rn_ * rn_) : R(0);
                xk += K_im * nx * l_; yk += K_im * ny * l_;
                if (sizeof(R) == 4) { bool big_ = false; M<R>::rotate_fast(ak, ck, sk, K_ii * rn_ * l_, big_); if (big_) bigm |= 1u << ol; }
                else if (on) M<R>::rotate(ak, ck, sk, K_ii * rn_ * l_);
              }
#else
              for (int q = 0; q < np; q++) {
                const R lx = q == 0 ? pl.x : pl.z, ly = q == 0 ? pl.y : pl.w;
                const R cpx = q == 0 ? pa.z : pb.y, cpy = q == 0 ? pa.w : pb.z, cdp = q == 0 ? pb.x : pb.w;
                const R rx = ck * lx - sk * ly, ry = sk * lx + ck * ly;
                const R sep = nx * (xk + rx - cpx) + ny * (yk + ry - cpy) - cdp;
                if (sep < -R(3) * K_lin_tol) mybad |= 1u << ol;
                R C = K_baum * (sep + K_lin_tol);
                C = M<R>::min_(C, R(0));
                C = M<R>::max_(C, -K_max_corr);
                const R rn_ = rx * ny - ry * nx;
                const R l_ = -C * M<R>::rcp(K_im + K_ii * rn_ * rn_);
                xk += K_im * nx * l_; yk += K_im * ny * l_;
                M<R>::rotate(ak, ck, sk, K_ii * rn_ * l_);
              }
#endif
            }
            stg[mi] = xk; stg[128 + mi] = yk; stg[256 + mi] = ak; pc[mi] = ck; pc[128 + mi] = sk;
          }
          __syncwarp();
          if ((__reduce_or_sync(FULL, mybad) >> lane) & 1u) ok = false;
#if VSR_POS_V2
          if (sizeof(R) == 4 && __any_sync(FULL, bigm != 0) && ((__reduce_or_sync(FULL, bigm) >> lane) & 1u)) bigrot = true;
#endif
          if (touching && active) {
            V4<R> t4;
            t4 = stage[0 * 32 + lane]; X[0] = t4.x; X[1] = t4.y; X[2] = t4.z; X[3] = t4.w;
            t4 = stage[1 * 32 + lane]; Y[0] = t4.x; Y[1] = t4.y; Y[2] = t4.z; Y[3] = t4.w;
            t4 = stage[2 * 32 + lane]; A[0] = t4.x; A[1] = t4.y; A[2] = t4.z; A[3] = t4.w;
            t4 = pcs[0 * 32 + lane]; CS[0] = t4.x; CS[1] = t4.y; CS[2] = t4.z; CS[3] = t4.w;
            t4 = pcs[1 * 32 + lane]; SN[0] = t4.x; SN[1] = t4.y; SN[2] = t4.z; SN[3] = t4.w;
          }
          __syncwarp();
        }
        // ropes
#if VSR_POS_V2
        // branch-free: an inactive rope (or robot) applies a zero correction; ropes (0,1) / (2,3) and (1,2) / (3,0) touch disjoint masses,
        // so the straight-line code gives the scheduler two independent chains
        if (warp_rope) {
#define ROPE_POS(i, a, b)                                                                  \
  {                                                                                        \
    const bool on_ = ract[i] && active;                                                    \
    R nx = X[a] - X[b], ny = Y[a] - Y[b];                                                  \
    const R l2 = nx * nx + ny * ny;                                                        \
    const R il = M<R>::rsqrt_(l2);                                                         \
    const R len = l2 * il;                                                                 \
    if (!(len >= K_lin_tol)) { nx = 0; ny = 0; } else { nx *= il; ny *= il; }                \
    R C = len - K_rope_lower;                                                                \
    if (on_ && !(-C < K_lin_tol)) ok = false;                                                \
    C = M<R>::min_(C, R(0));                                                               \
    C = M<R>::max_(C, -K_max_corr);                                                          \
    const R l_ = -K_rope_E * C;                                                              \
    const R m_ = on_ ? l_ * K_im : R(0);                                                     \
    X[a] += nx * m_; Y[a] += ny * m_; X[b] -= nx * m_; Y[b] -= ny * m_;                    \
  }
          VSR_ROPES(ROPE_POS)
#undef ROPE_POS
        }
#else
        if (warp_rope) {
#define ROPE_POS(i, a, b)                                                                  \
  if (ract[i] && active) {                                                                 \
    R nx = X[a] - X[b], ny = Y[a] - Y[b];                                                  \
    const R l2 = nx * nx + ny * ny;                                                        \
    const R il = M<R>::rsqrt_(l2);                                                         \
    const R len = l2 * il;                                                                 \
    if (!(len >= K_lin_tol)) { nx = 0; ny = 0; } else { nx *= il; ny *= il; }                \
    R C = len - K_rope_lower;                                                                \
    if (!(-C < K_lin_tol)) ok = false;                                                       \
    C = M<R>::min_(C, R(0));                                                               \
    C = M<R>::max_(C, -K_max_corr);                                                          \
    const R l_ = -K_rope_E * C;                                                              \
    const R m_ = l_ * K_im;                                                                  \
    X[a] += nx * m_; Y[a] += ny * m_; X[b] -= nx * m_; Y[b] -= ny * m_;                    \
  }
          VSR_ROPES(ROPE_POS)
#undef ROPE_POS
        }
#endif
        // weld position solve for owner weld (body1 = mine k1; partner pose + rotation pulled)
#if VSR_POS_V2
#define POS_ROT(a_, c_, s_, d_) { if (sizeof(R) == 4) M<R>::rotate_fast(a_, c_, s_, d_, bigrot); else M<R>::rotate(a_, c_, s_, d_); }
        // local anchor 2 is (m, 0) for a horizontal weld and (0, m) for a vertical one: r2 = R(partner) * anchor without the zero products
#define WELD_R2(pc_, ps_, hz_) const R r2x = (hz_) ? (pc_) * K_wa : -(ps_) * K_wa, r2y = (hz_) ? (ps_) * K_wa : (pc_) * K_wa;
#else
#define POS_ROT(a_, c_, s_, d_) M<R>::rotate(a_, c_, s_, d_);
#define WELD_R2(pc_, ps_, hz_) const R r2x = (pc_) * ((hz_) ? K_wa : R(0)) - (ps_) * ((hz_) ? R(0) : K_wa), r2y = (ps_) * ((hz_) ? K_wa : R(0)) + (pc_) * ((hz_) ? R(0) : K_wa);
#endif
#define WELD_POS(k1, px_, py_, pa_, pc_, ps_, hz_, dpx, dpy, dpa)                                 \
  {                                                                                             \
    WELD_R2(pc_, ps_, hz_)                                                                      \
    const R Cx = X[k1] - ((px_) + r2x), Cy = Y[k1] - ((py_) + r2y), Cz = A[k1] - (pa_);         \
    if (!(Cx * Cx + Cy * Cy <= K_lin_tol * K_lin_tol && M<R>::abs_(Cz) <= K_ang_tol)) ok = false;     \
    R lx_, ly_, lz_;                                                                            \
    WELD_SOLVE(r2x, r2y, Cx, Cy, Cz, lx_, ly_, lz_)                                             \
    X[k1] += K_im * lx_; Y[k1] += K_im * ly_;                                                       \
    POS_ROT(A[k1], CS[k1], SN[k1], K_ii * lz_)                                                    \
    dpx = -K_im * lx_; dpy = -K_im * ly_; dpa = -K_ii * (r2x * ly_ - r2y * lx_ + lz_);                \
  }
        {
          const R p0x = __shfl_sync(FULL, X[1], ln_s) + lox, p0y = __shfl_sync(FULL, Y[1], ln_s) + loy, p0a = __shfl_sync(FULL, A[1], ln_s);
          const R p0c = __shfl_sync(FULL, CS[1], ln_s), p0s = __shfl_sync(FULL, SN[1], ln_s);
          const R p1x = __shfl_sync(FULL, X[2], ln_s) + lox, p1y = __shfl_sync(FULL, Y[2], ln_s) + loy, p1a = __shfl_sync(FULL, A[2], ln_s);
          const R p1c = __shfl_sync(FULL, CS[2], ln_s), p1s = __shfl_sync(FULL, SN[2], ln_s);
          R d0x = 0, d0y = 0, d0a = 0, d1x = 0, d1y = 0, d1a = 0;
          if (ln >= 0 && active) {
            WELD_POS(0, p0x, p0y, p0a, p0c, p0s, 1, d0x, d0y, d0a)
            WELD_POS(3, p1x, p1y, p1a, p1c, p1s, 1, d1x, d1y, d1a)
          }
          const R a0 = __shfl_sync(FULL, d0x, rn_s), a1 = __shfl_sync(FULL, d0y, rn_s), a2 = __shfl_sync(FULL, d0a, rn_s);
          const R b0 = __shfl_sync(FULL, d1x, rn_s), b1 = __shfl_sync(FULL, d1y, rn_s), b2 = __shfl_sync(FULL, d1a, rn_s);
          if (rn >= 0 && active) {
            X[1] += a0; Y[1] += a1; POS_ROT(A[1], CS[1], SN[1], a2)
            X[2] += b0; Y[2] += b1; POS_ROT(A[2], CS[2], SN[2], b2)
          }
        }
        {
          const R p0x = __shfl_sync(FULL, X[0], dn_s) + dox, p0y = __shfl_sync(FULL, Y[0], dn_s) + doy, p0a = __shfl_sync(FULL, A[0], dn_s);
          const R p0c = __shfl_sync(FULL, CS[0], dn_s), p0s = __shfl_sync(FULL, SN[0], dn_s);
          const R p1x = __shfl_sync(FULL, X[1], dn_s) + dox, p1y = __shfl_sync(FULL, Y[1], dn_s) + doy, p1a = __shfl_sync(FULL, A[1], dn_s);
          const R p1c = __shfl_sync(FULL, CS[1], dn_s), p1s = __shfl_sync(FULL, SN[1], dn_s);
          R d0x = 0, d0y = 0, d0a = 0, d1x = 0, d1y = 0, d1a = 0;
          if (dn >= 0 && active) {
            WELD_POS(3, p0x, p0y, p0a, p0c, p0s, 0, d0x, d0y, d0a)
            WELD_POS(2, p1x, p1y, p1a, p1c, p1s, 0, d1x, d1y, d1a)
          }
          const R e0 = __shfl_sync(FULL, d0x, un_s), e1 = __shfl_sync(FULL, d0y, un_s), e2 = __shfl_sync(FULL, d0a, un_s);
          const R f0 = __shfl_sync(FULL, d1x, un_s), f1 = __shfl_sync(FULL, d1y, un_s), f2 = __shfl_sync(FULL, d1a, un_s);
          if (un >= 0 && active) {
            X[0] += e0; Y[0] += e1; POS_ROT(A[0], CS[0], SN[0], e2)
            X[1] += f0; Y[1] += f1; POS_ROT(A[1], CS[1], SN[1], f2)
          }
        }
#undef WELD_POS
#undef WELD_R2
#undef POS_ROT
#if VSR_POS_V2
        // a rotation step beyond the small-angle polynomial's range (rare: falls, hard landings): refresh sin / cos of the exactly
        // accumulated angles of the voxel's masses before the next iteration (mass lanes report through `bigm` to the owner lane)
        // (per lane: a robot's result must not depend on which robots share its warp)
        if (sizeof(R) == 4 && __any_sync(FULL, bigrot)) {
          if (bigrot) {
#pragma unroll
            for (int k = 0; k < 4; k++) M<R>::sincos_(A[k], &SN[k], &CS[k]);
          }
          bigrot = false;
        }
#endif
        // robot-wide convergence
        const unsigned bad = __ballot_sync(FULL, active && !ok);
        if (!(bad & robot_mask)) active = false;
        if (!__any_sync(FULL, active)) break;
      }
#if VSR_REFRESH_CS
      // (cos, sin) were rotated incrementally through the iterations (FP32: each update rounds): re-evaluate them once from the exactly
      // accumulated angles, so that detect, sensing and the next step's spring initialisation start from sin / cos accurate to 2 ulp
      // (profiles/r2_fp32_approx_ab.md: this is what ended step-wise tracking of the oracle first)
      if (sizeof(R) == 4) {
#pragma unroll
        for (int k = 0; k < 4; k++) M<R>::sincos_(A[k], &SN[k], &CS[k]);
      }
#endif
    }

    // ---- re-centre the lane origin (integer shifts; keeps |local| small) -------------------------------
    PHASE_SYNC(512)
    {
      const R mx = R(0.25) * (X[0] + X[1] + X[2] + X[3]), my = R(0.25) * (Y[0] + Y[1] + Y[2] + Y[3]);
      if (M<R>::abs_(mx) > R(1) || M<R>::abs_(my) > R(1)) {
        const R sx = M<R>::rint_(mx), sy = M<R>::rint_(my);
        ox += sx; oy += sy;
#pragma unroll
        for (int k = 0; k < 4; k++) { X[k] -= sx; Y[k] -= sy; }
        pk = 0; // cached polygon is expressed in the old frame
      }
    }
    // ---- detect for the next step / touch sensors -----------------------------------------------------
    PHASE_SYNC(16)
    detect();
    PHASE_SYNC(32)

    // ---- robot.act(t): sensors -> controller -> applyForce ---------------------------------------------
    // outer voxel-poly vertices (NW NE SE SW), local frame
    R qx[4], qy[4];
    {
      const R sxk[4] = {-1, +1, +1, -1}, syk[4] = {+1, +1, -1, -1};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const R lx = sxk[k] * K_half, ly = syk[k] * K_half;
        qx[k] = X[k] + CS[k] * lx - SN[k] * ly;
        qy[k] = Y[k] + SN[k] * lx + CS[k] * ly;
      }
    }
    R f = 0;
    if (CTRL == VSR_CTRL_PHASE_SIN) {
      if (P.time_sin) {
        // TimeFunctions of per-voxel sinusoids (BodyBrainSinusoidal.java:87-119): the triple is re-read once per step (L1/L2)
        if (valid) {
          const double* g3 = P.genomes + (size_t)robot * P.genome_len + 3 * v;
          f = M<R>::phase_sin(__ldg(g3 + 2), __ldg(g3), t, (R)__ldg(g3 + 1));
        }
      } else
        f = M<R>::phase_sin(P.amp, P.freq, t, phase);
    } else {
      // sensor readings of this voxel
      R rd[kMaxReadings];
#pragma unroll
      for (int i = 0; i < kMaxReadings; i++) rd[i] = 0;
      R vang = 0;
      bool have_ang = false;
      if (SENS == 1) {
        // specialised build: one soft-normalised AreaRatio reading per voxel (`uniform-a-<sigma>`, the reference default,
        // Starter.java:152), optionally Noisy -- no sensor loop, none of the velocity / angle / Average code in the stream
        const SensorK& sk = P.sensors[0];
        R a2 = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) { const int j = (i + 1) & 3; a2 += qx[i] * qy[j] - qx[j] * qy[i]; }
        R val = M<R>::abs_(a2) * K_inv_area2;
        const R vv = (val - (R)sk.dom_min) * (R)sk.dom_inv;
        val = M<R>::tanh_((vv * R(2) - R(1)) * R(2)) * R(0.5) + R(0.5);
        if (P.noise) val += (R)__ldg(&P.noise[step]);
        rd[0] = val;
      } else {
      const int nsv = valid ? (int)P.vs_n[vi] : 0;
      int ri = 0; // next reading of this voxel
#pragma unroll 1
      for (int sj = 0; sj < P.vs_max; sj++) { // warp-uniform trip count; voxels with shorter lists idle
        if (sj >= nsv) continue;
        const SensorK& sk = P.sensors[P.vs_idx[vi][sj]];
        R raw0 = 0, raw1 = 0;
        R rays[kMaxVoxSensors];
        if (sk.kind == VSR_SENSOR_AREA_RATIO) {
          R a2 = 0;
#pragma unroll
          for (int i = 0; i < 4; i++) { const int j = (i + 1) & 3; a2 += qx[i] * qy[j] - qx[j] * qy[i]; }
          raw0 = M<R>::abs_(a2) * K_inv_area2;
        } else if (sk.kind == VSR_SENSOR_TOUCH) {
          raw0 = touching ? R(1) : R(0);
        } else if (sk.kind == VSR_SENSOR_CONSTANT) {
          raw0 = (R)P.vs_param[vi][sj];
        } else if (sk.kind == VSR_SENSOR_TIME_SIN) {
          raw0 = M<R>::phase_sin(1.0, sk.param, t, R(0));
        } else {
          if (!have_ang) {
            vang = (M<R>::atan2_(Y[1] - Y[0], X[1] - X[0]) + M<R>::atan2_(Y[2] - Y[3], X[2] - X[3])) * R(0.5);
            have_ang = true;
          }
          if (sk.kind == VSR_SENSOR_ANGLE) raw0 = vang;
          else if (sk.kind == VSR_SENSOR_LIDAR) {
            // Lidar [U]: rays from the voxel centre, directions of one side's sector rotated by the voxel angle, first hit on the
            // ground's surface (the polyline + the two outer walls) within the ray length; restated in oracle raycast_ground
            const double oxd = (double)ox, oyd = (double)oy;
            const R cx = R(0.25) * (X[0] + X[1] + X[2] + X[3]), cy = R(0.25) * (Y[0] + Y[1] + Y[2] + Y[3]);
            const R len = (R)sk.param;
            const int side = (int)P.vs_param[vi][sj];
            const int t0 = P.toff ? __ldg(&P.toff[robot]) : 0, tn = P.toff ? __ldg(&P.toff[robot + 1]) - t0 : P.n_t;
            const double* tx = P.tx + t0;
            const double* ty = P.ty + t0;
            const double basey = P.toff ? __ldg(&P.tbase[robot]) : P.base_y;
            // first segment that can matter: the one whose right end is not left of cx - len
            int k0 = kh;
            const double xl = oxd + (double)(cx - len), xr = oxd + (double)(cx + len);
            while (k0 > 1 && __ldg(&tx[k0 - 1]) > xl) k0--;
            while (k0 < tn - 1 && __ldg(&tx[k0]) < xl) k0++;
            const double startA = side == 0 ? 0.25 : (side == 1 ? -0.25 : (side == 2 ? 1.25 : 0.75)); // in units of pi
            const double endA = side == 0 ? 0.75 : (side == 1 ? 0.25 : (side == 2 ? 1.75 : 1.25));
#pragma unroll 1
            for (int i = 0; i < sk.n_rays; i++) {
              const double rel = sk.n_rays == 1 ? (endA - startA) / 2.0 * 3.14159265358979323846
                                                : (startA * 3.14159265358979323846 + (endA - startA) * 3.14159265358979323846 / (double)(sk.n_rays - 1) * (double)i);
              R dx, dy;
              M<R>::sincos_((R)rel + vang, &dy, &dx);
              R best = len;
#pragma unroll 1
              for (int k = k0; k < tn; k++) {
                const double x0d = __ldg(&tx[k - 1]);
                if (x0d > xr) break;
                const R x0 = (R)(x0d - oxd), y0 = (R)(__ldg(&ty[k - 1]) - oyd);
                const R ex = (R)(__ldg(&tx[k]) - oxd) - x0, ey = (R)(__ldg(&ty[k]) - oyd) - y0;
                const R den = dx * ey - dy * ex;
                if (den == R(0)) continue;
                const R tt = ((x0 - cx) * ey - (y0 - cy) * ex) / den, uu = ((x0 - cx) * dy - (y0 - cy) * dx) / den;
                if (tt >= R(0) && tt < best && uu >= R(0) && uu <= R(1)) best = tt;
              }
#pragma unroll
              for (int wside = 0; wside < 2; wside++) { // outer walls of the first / last ground polygon
                const int kw = wside == 0 ? 0 : tn - 1;
                const R xw = (R)(__ldg(&tx[kw]) - oxd), yt = (R)(__ldg(&ty[kw]) - oyd);
                if (dx == R(0)) continue;
                const R tt = (xw - cx) / dx, yy = cy + tt * dy;
                if (tt >= R(0) && tt < best && yy <= yt && yy >= (R)(basey - oyd)) best = tt;
              }
#pragma unroll
              for (int z = 0; z < kMaxVoxSensors; z++) if (z == i) rays[z] = best;
            }
          } else {
            R vx = R(0.25) * (VX[0] + VX[1] + VX[2] + VX[3]), vy = R(0.25) * (VY[0] + VY[1] + VY[2] + VY[3]);
            if (sk.rotated) {
              R sn_, cs_;
              M<R>::sincos_(-vang, &sn_, &cs_);
              const R rx = cs_ * vx - sn_ * vy, ry = sn_ * vx + cs_ * vy;
              vx = rx; vy = ry;
            }
            if ((sk.axes & 3) == 3) { raw0 = vx; raw1 = vy; }
            else raw0 = (sk.axes & 1) ? vx : vy;
          }
        }
        for (int d = 0; d < sk.dim; d++) {
          R val = d == 0 ? raw0 : raw1;
          if (sk.kind == VSR_SENSOR_LIDAR) {
#pragma unroll
            for (int z = 0; z < kMaxVoxSensors; z++) if (z == d) val = rays[z];
          }
          if (sk.avg >= 0) {
            // Average: mean of the last win readings (window table emulates the t_j >= t - w test in double)
            const int ai = sk.avg + d;
            const int slot = (step - 1) % P.hist_cap;
            R* hist = reinterpret_cast<R*>(P.hist);
            hist[((size_t)slot * P.n_avg + ai) * total_lanes + gl] = val;
            double sum = savg[ai * 32 + lane] + (double)val;
            int cnt = savg_cnt[ai * 32 + lane] + 1;
            const int win = P.win_tab[(size_t)ai * (P.n_steps + 1) + step];
            while (cnt > win) {
              const int oslot = (step - cnt + P.hist_cap * 4) % P.hist_cap;
              sum -= (double)hist[((size_t)oslot * P.n_avg + ai) * total_lanes + gl];
              cnt--;
            }
            savg[ai * 32 + lane] = sum; savg_cnt[ai * 32 + lane] = cnt;
            val = (R)(sum / cnt);
          }
          if (sk.soft_norm == 1) {
            const R vv = (val - (R)sk.dom_min) * (R)sk.dom_inv;
            val = M<R>::tanh_((vv * R(2) - R(1)) * R(2)) * R(0.5) + R(0.5);
          } else if (sk.soft_norm == 2) { // Normalization: linear map of the domain on [0, 1], clipped
            val = M<R>::min_(M<R>::max_((val - (R)sk.dom_min) * (R)sk.dom_inv, R(0)), R(1));
          }
          if (P.noise) val += (R)__ldg(&P.noise[(size_t)(sk.noise_row + d) * (P.n_steps + 1) + step]); // Noisy wrapper (host-replayed java.util.Random)
#pragma unroll
          for (int z = 0; z < kMaxReadings; z++) if (z == ri) rd[z] = val;
          ri++;
        }
      }
      }
      const int D = P.sens_dim, maxN = P.max_neurons;
      if (CTRL == VSR_CTRL_MLP_CENTRALIZED) {
        R* a0 = act + (size_t)(lane_used ? rw : 0) * 2 * maxN;
        const R* w_ = wts + (size_t)(lane_used ? rw : 0) * P.genome_len;
        if (valid) {
#pragma unroll
          for (int d = 0; d < (SENS == 1 ? 1 : kMaxReadings); d++) if (d < (int)P.vox_dim[vi]) a0[P.vox_in_off[vi] + d] = P.in_act ? M<R>::tanh_(rd[d]) : rd[d];
        }
        __syncwarp();
        int cur = 0, o = 0;
#pragma unroll 1
        for (int l = 1; l < P.n_layers; l++) {
          const int nf = P.layers[l - 1], nt = P.layers[l];
          const R* in = a0 + cur * maxN;
          R* out = a0 + (cur ^ 1) * maxN;
          if (valid)
            for (int j = v; j < nt; j += L) {
              const R* wj = w_ + o + j * (nf + 1);
              R sum = wj[0];
              for (int k = 0; k < nf; k++) sum += in[k] * wj[1 + k];
              out[j] = M<R>::tanh_(sum);
            }
          __syncwarp();
          o += nt * (nf + 1);
          cur ^= 1;
        }
        if (valid) f = a0[cur * maxN + v];
        __syncwarp();
      } else { // homogeneous DistributedSensing: every voxel runs the same MLP on its own inputs
        const int sg = P.signals;
        R* a0 = act; // [2][maxN][32]
        const R* w_ = wts + (size_t)(lane_used ? rw : 0) * P.genome_len + (size_t)(lane_used ? v : 0) * P.w_stride;
        const R* sin_ = sig + (size_t)sig_par * 32 * 4 * sg;
        R* sout = sig + (size_t)(sig_par ^ 1) * 32 * 4 * sg;
        int ni = 0;
#pragma unroll
        for (int d = 0; d < kMaxReadings; d++) if (d < D) a0[(ni++) * 32 + lane] = P.in_act ? M<R>::tanh_(rd[d]) : rd[d];
        // Dir N(0,-1) E(1,0) S(0,1) W(-1,0); read the neighbour's signal emitted towards me (opposite side)
#pragma unroll
        for (int d = 0; d < 4; d++) {
          const int nbl = d == 0 ? dn : (d == 1 ? rn : (d == 2 ? un : ln));
          const int opp = (d + 2) & 3;
          for (int s_ = 0; s_ < sg; s_++) {
            const R x_ = nbl >= 0 ? sin_[(nbl * 4 + opp) * sg + s_] : R(0);
            a0[(ni++) * 32 + lane] = P.in_act ? M<R>::tanh_(x_) : x_;
          }
        }
        int cur = 0, o = 0;
#pragma unroll 1
        for (int l = 1; l < P.n_layers; l++) {
          const int nf = P.layers[l - 1], nt = P.layers[l];
          const R* in = a0 + (size_t)cur * maxN * 32;
          R* out = a0 + (size_t)(cur ^ 1) * maxN * 32;
          for (int j = 0; j < nt; j++) {
            const R* wj = w_ + o + j * (nf + 1);
            R sum = wj[0];
            for (int k = 0; k < nf; k++) sum += in[k * 32 + lane] * wj[1 + k];
            out[j * 32 + lane] = M<R>::tanh_(sum);
          }
          o += nt * (nf + 1);
          cur ^= 1;
        }
        const R* res = a0 + (size_t)cur * maxN * 32;
        f = res[lane];
        for (int i = 0; i < 4 * sg; i++) sout[(lane * 4) * sg + i] = res[(1 + i) * 32 + lane];
        sig_par ^= 1;
        __syncwarp();
      }
    }
    PHASE_SYNC(1024)
    // Voxel.applyForce(f), ForceMethod.DISTANCE (a frozen actuator applies its last force before freezing: 0)
    if (frozen) f = 0;
    if (M<R>::abs_(f) > R(1)) f = f > 0 ? R(1) : R(-1);
    ctrl_f = f;
#pragma unroll
    for (int kq = 0; kq < 3; kq++) {
      const R rr = KC(range_rest[kq], P.range_rest[kq]);
      rest[kq] = f >= 0 ? rr - (rr - KC(range_min[kq], P.range_min[kq])) * f : rr + (KC(range_max[kq], P.range_max[kq]) - rr) * -f;
    }

    // ---- energy terms of the kept observations (Outcome.getControlPower / getAreaRatioPower inputs), only when the caller
    // asked for vsr_result records: per-voxel accumulators live in global memory (L2), nothing is added to the fitness-only path
    if (P.results != nullptr && step >= P.k_first && step <= P.k_last) {
      R a2 = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) { const int j = (i + 1) & 3; a2 += qx[i] * qy[j] - qx[j] * qy[i]; }
      const R area = M<R>::abs_(a2) * K_inv_area2;
      V4<R>* eacc = reinterpret_cast<V4<R>*>(P.energy) + gl;
      V4<R> e4;
      if (step == P.k_first) { e4.x = 0; e4.y = 0; e4.z = 0; }
      else {
        e4 = *eacc;
        e4.x += f * f * K_dt; e4.y += M<R>::abs_(area - e4.w); e4.z += M<R>::abs_(area * area - e4.w * e4.w);
      }
      e4.w = area;
      *eacc = e4;
    }
    // ---- observation: robot centre x at the first / last kept step --------------------------------------
    const bool obs_all = P.centers != nullptr && step >= P.k_first && step <= P.k_last;
    if (step == P.k_first || step == P.k_last || obs_all) {
      const double cxv = (double)ox + (double)(R(0.25) * (qx[0] + qx[1] + qx[2] + qx[3]));
      double sum = 0;
      for (int j = 0; j < L; j++) sum += __shfl_sync(FULL, cxv, (base_lane + j) & 31);
      sum /= L;
      if (obs_all) { // centre trajectory of the kept observations (inputs of Outcome's centre spectra, NamedFunctions.java:139-172)
        const double cyv = (double)oy + (double)(R(0.25) * (qy[0] + qy[1] + qy[2] + qy[3]));
        double sumy = 0;
        for (int j = 0; j < L; j++) sumy += __shfl_sync(FULL, cyv, (base_lane + j) & 31);
        if (valid && v == 0) {
          double* c = P.centers + ((size_t)robot * (size_t)(P.k_last - P.k_first + 1) + (size_t)(step - P.k_first)) * 2;
          c[0] = sum; c[1] = sumy / L;
        }
      }
      double en0 = 0, en1 = 0, en2 = 0;
      if (P.results != nullptr && step == P.k_last) { // per-robot sums of the voxel accumulators, in voxel order
        const V4<R> e4 = reinterpret_cast<const V4<R>*>(P.energy)[gl];
        for (int j = 0; j < L; j++) {
          en0 += __shfl_sync(FULL, (double)e4.x, (base_lane + j) & 31);
          en1 += __shfl_sync(FULL, (double)e4.y, (base_lane + j) & 31);
          en2 += __shfl_sync(FULL, (double)e4.z, (base_lane + j) & 31);
        }
      }
      // K3 result pack: the first observation waits in the fitness slot (global) until the last one arrives
      if (valid && v == 0) {
        if (step == P.k_first) P.fitness[robot] = sum;
        if (step == P.k_last) {
          const double cx_first = P.k_last == P.k_first ? sum : P.fitness[robot];
          const double dist = sum - cx_first, tm = P.t_last - P.t_first;
          const double vel = dist / tm;
          P.fitness[robot] = vel;
          if (P.results) {
            vsr_result r;
            r.velocity = vel; r.distance = dist; r.time = tm; r.first_x = cx_first; r.last_x = sum;
            r.n_steps = P.n_steps; r.k_first = P.k_first; r.k_last = P.k_last; r.contact_overflow = 0;
            r.control_energy = en0; r.area_ratio_abs = en1; r.area_ratio_sq = en2; r.reserved = 0;
            P.results[robot] = r;
          }
        }
      }
    }
    // ---- per-observation summaries (OBS builds; BehaviorUtils of 2dhmsr [U], restated in oracle/vsr_oracle.c) ----------------
    if (OBS && step >= P.k_first && step <= P.k_last) {
      const unsigned robot_mask = lane_used ? (((L >= 32) ? FULL : ((1u << L) - 1u)) << base_lane) : (1u << lane);
      // voxel polygon relative to the origin of the robot's first voxel (origins differ by integers: exact in FP32)
      const R oxr = __shfl_sync(FULL, ox, base_lane & 31), oyr = __shfl_sync(FULL, oy, base_lane & 31);
      R rx[4], ry[4];
#pragma unroll
      for (int k = 0; k < 4; k++) { rx[k] = qx[k] + (ox - oxr); ry[k] = qy[k] + (oy - oyr); }
      const R vminx = M<R>::min_(M<R>::min_(rx[0], rx[1]), M<R>::min_(rx[2], rx[3])), vmaxx = M<R>::max_(M<R>::max_(rx[0], rx[1]), M<R>::max_(rx[2], rx[3]));
      const R vminy = M<R>::min_(M<R>::min_(ry[0], ry[1]), M<R>::min_(ry[2], ry[3])), vmaxy = M<R>::max_(M<R>::max_(ry[0], ry[1]), M<R>::max_(ry[2], ry[3]));
      R minx = vminx, maxx = vmaxx, miny = vminy, maxy = vmaxy;
      for (int j = 0; j < L; j++) {
        const int src = (base_lane + j) & 31;
        minx = M<R>::min_(minx, __shfl_sync(FULL, vminx, src)); maxx = M<R>::max_(maxx, __shfl_sync(FULL, vmaxx, src));
        miny = M<R>::min_(miny, __shfl_sync(FULL, vminy, src)); maxy = M<R>::max_(maxy, __shfl_sync(FULL, vmaxy, src));
      }
      const int ko = step - P.k_first, K = P.k_last - P.k_first + 1;
      // BehaviorUtils.computeFootprint(polies, n): x-range of every touching voxel on n segments of the robot's x-extent
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int n = q == 0 ? P.fp_n0 : P.fp_n1;
        uint8_t* dst = q == 0 ? P.footprints0 : P.footprints1;
        if (n <= 0 || dst == nullptr) continue;
        unsigned bits = 0;
        if (valid && touching) {
          const R sc = R(n - 1) / (maxx - minx);
          int i0 = (int)floor((double)((vminx - minx) * sc + R(0.5))), i1 = (int)floor((double)((vmaxx - minx) * sc + R(0.5)));
          i0 = i0 < 0 ? 0 : i0; i1 = i1 > n - 1 ? n - 1 : i1;
          for (int x = i0; x <= i1; x++) bits |= 1u << x;
        }
        bits = __reduce_or_sync(robot_mask, bits);
        if (valid && v == 0) dst[(size_t)robot * K + ko] = (uint8_t)bits;
      }
      // BehaviorUtils.computePosture(polies, n): cells of the squared bounding box whose centre lies inside a voxel polygon
      if (P.post_n > 0) {
        const int n = P.post_n;
        R bx0 = minx, bx1 = maxx, by0 = miny, by1 = maxy;
        const R wdt = bx1 - bx0, hgt = by1 - by0;
        if (hgt < wdt) { const R d = wdt - hgt; by0 -= d / 2; by1 += d / 2; }
        else if (hgt > wdt) { const R d = hgt - wdt; bx0 -= d / 2; bx1 += d / 2; }
        const R side = bx1 - bx0;
        unsigned lo = 0, hi = 0, colm = 0, rowm = 0;
        if (valid) {
          // rows / columns whose cell centre lies inside the voxel's bounding box (a cell outside it cannot be inside the polygon)
          for (int i = 0; i < n; i++) {
            const R px = bx0 + (R(i) + R(0.5)) / R(n) * side, py = by0 + (R(i) + R(0.5)) / R(n) * side;
            if (!(px < vminx || px > vmaxx)) colm |= 1u << i;
            if (!(py < vminy || py > vmaxy)) rowm |= 1u << i;
          }
        }
        // every lane walks ITS OWN candidate cells (rows x columns, typically 9-16 of the 64), one per trip: the trip count is the
        // largest candidate count of the warp, not the union of everybody's rows times the widest column set
        {
          unsigned rm = rowm, cm = colm;
          bool more = rm != 0 && cm != 0;
          while (__any_sync(FULL, more)) {
            if (more) {
              const int iy = __ffs(rm) - 1, ix = __ffs(cm) - 1;
              const R px = bx0 + (R(ix) + R(0.5)) / R(n) * side, py = by0 + (R(iy) + R(0.5)) / R(n) * side;
              bool in = false;
#pragma unroll
              for (int i = 0; i < 4; i++) {
                const int j = (i + 3) & 3;
                if (((ry[i] > py) != (ry[j] > py)) && (px < (rx[j] - rx[i]) * (py - ry[i]) / (ry[j] - ry[i]) + rx[i])) in = !in;
              }
              const int c = iy * n + ix;
              if (in) { if (c < 32) lo |= 1u << c; else hi |= 1u << (c - 32); }
              cm &= cm - 1;
              if (cm == 0) { cm = colm; rm &= rm - 1; more = rm != 0; }
            }
          }
        }
        lo = __reduce_or_sync(robot_mask, lo); hi = __reduce_or_sync(robot_mask, hi);
        // cell c of a robot is counted by its lane c mod L.  Bodies of 8 voxels or more (at most 8 cells per lane) keep the counters
        // on chip, two 16-bit counters per register, and store them once at the last kept observation; smaller bodies (or more than
        // 65535 observations) count in L2
        const bool reg_cnt = L >= 8 && K <= 65535;
        if (valid) {
          if (reg_cnt) {
#pragma unroll
            for (int q = 0; q < 8; q++) {
              const int c = v + q * L;
              const unsigned bit = c < 32 ? (lo >> c) & 1u : (c < 64 ? (hi >> (c - 32)) & 1u : 0u);
              pcnt[q >> 1] += bit << (16 * (q & 1));
            }
            if (step == P.k_last) {
#pragma unroll
              for (int q = 0; q < 8; q++) {
                const int c = v + q * L;
                if (c < n * n) P.post_cnt[(size_t)robot * 64 + c] = (pcnt[q >> 1] >> (16 * (q & 1))) & 0xffffu;
              }
            }
          } else {
            for (int c = v; c < n * n; c += L)
              if ((c < 32 ? (lo >> c) : (hi >> (c - 32))) & 1u) atomicAdd(&P.post_cnt[(size_t)robot * 64 + c], 1u);
          }
        }
      }
      if (P.centers3 != nullptr) {
        const double cxv = (double)ox + (double)(R(0.25) * (qx[0] + qx[1] + qx[2] + qx[3]));
        const double cyv = (double)oy + (double)(R(0.25) * (qy[0] + qy[1] + qy[2] + qy[3]));
        // VoxelPoly.getAngle() of the central voxel: mean direction of the polygon's upper and lower sides
        const double ang = (double)((M<R>::atan2_(qy[1] - qy[0], qx[1] - qx[0]) + M<R>::atan2_(qy[2] - qy[3], qx[2] - qx[3])) * R(0.5));
        double sx = 0, sy = 0;
        for (int j = 0; j < L; j++) { sx += __shfl_sync(FULL, cxv, (base_lane + j) & 31); sy += __shfl_sync(FULL, cyv, (base_lane + j) & 31); }
        const double ca = __shfl_sync(FULL, ang, (base_lane + P.central_vox) & 31);
        if (valid && v == 0) {
          double* c = P.centers3 + ((size_t)robot * K + ko) * 3;
          c[0] = sx / L; c[1] = sy / L; c[2] = ca;
        }
      }
    }
    // ---- trace (parity tests; robot 0 only) ---------------------------------------------------------------
    if (P.trace_steps > 0 && step <= P.trace_steps && valid && robot == 0) {
      const size_t nb = (size_t)P.n_vox * 4;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (P.tr_vertices) {
          double* d = P.tr_vertices + (((size_t)(step - 1) * P.n_vox + v) * 4 + k) * 2;
          d[0] = (double)ox + (double)qx[k]; d[1] = (double)oy + (double)qy[k];
        }
        if (P.tr_bodies) {
          double* d = P.tr_bodies + ((size_t)(step - 1) * nb + v * 4 + k) * 6;
          d[0] = (double)ox + (double)X[k]; d[1] = (double)oy + (double)Y[k]; d[2] = A[k];
          d[3] = VX[k]; d[4] = VY[k]; d[5] = W[k];
        }
      }
      if (P.tr_ctrl) P.tr_ctrl[(size_t)(step - 1) * P.n_vox + v] = ctrl_f;
    }
  } // steps

  // ---- epilogue: contact-pool overflow count (fitness / result were written at step k_last) ---------------
  // A dropped manifold means missing ground contact for the rest of the episode: the robots of such a warp get a NaN fitness
  // (NaN ranks last in the generation loop and in the host comparator), so a solver can never select on an unreliable value;
  // vsr_result keeps the finite numbers next to contact_overflow for diagnosis.
  const int ovf = __reduce_add_sync(FULL, overflow);
  if (valid && v == 0 && ovf) {
    P.fitness[robot] = __longlong_as_double(0x7ff8000000000000LL);
    if (P.results) P.results[robot].contact_overflow = ovf;
  }
  if (ovf && lane == 0 && P.overflow_count) atomicAdd(P.overflow_count, 1);
#undef TM_MI
#undef TM_SLOTS
#undef WELD_APPLY
#undef WELD_SOLVE
#undef ln_s
#undef rn_s
#undef dn_s
#undef un_s
#undef ln
#undef rn
#undef dn
#undef un
#undef K_dt
#undef K_half
#undef K_im
#undef K_ii
#undef K_mu
#undef K_lin_tol
#undef K_ang_tol
#undef K_max_corr
#undef K_baum
#undef K_lin_f
#undef K_ang_f
#undef K_sp_sigma
#undef K_sp_boversig
#undef K_sp_cfrac
#undef K_rope_lower
#undef K_rope_E
#undef K_wa
#undef K_wA
#undef K_wh
#undef K_wDA
#undef K_wDh
#undef K_w_i2i
#undef K_rest_vel
#undef K_rest_e
#undef K_max_trans
#undef K_max_trans2
#undef K_max_rot
#undef K_inv_area2
#undef KC
}

} // namespace vsr
