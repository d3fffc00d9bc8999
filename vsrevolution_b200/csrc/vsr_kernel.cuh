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
//
// Constraint order == the oracle's canonical order (oracle/vsr_oracle.c header).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vsr_eval.h"

namespace vsr {

constexpr int kMaxVox = 32;
constexpr int kMaxSensors = 6;
constexpr int kMaxLayers = 8;
constexpr int kMaxMan = 3;       // manifolds per corner mass (ground polygons it can touch at once)
constexpr int kMaxAvg = 6;       // averaged readings per voxel
constexpr int kPoolPersist = 12; // NX NY (PX PY D JN JT)x2
constexpr int kPoolSolve = 18;   // (RX RY MN MT VB LX LY)x2 K12 IK11 IK12 IK22

struct SensorK {
  int kind, axes, rotated, soft_norm;
  double dom_min, dom_max;
  int avg;      // index into the averaged-reading table or -1
  int off, dim; // offset / size inside the voxel's reading vector
};

struct KParams {
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
  double base_y;
  // sensors
  int n_sensors, sens_dim, n_avg, hist_cap;
  SensorK sensors[kMaxSensors];
  const uint8_t* win_tab; // [n_avg][n_steps+1] window length (in readings) at step k
  void* hist;             // R[hist_cap][n_avg][total_lanes] ring of raw readings
  // controller
  int ctrl_kind, n_layers, layers[kMaxLayers], signals, in_act, genome_len, max_neurons;
  double freq, amp;
  // episode
  int n_steps, k_first, k_last;
  // io
  const double* genomes;
  double* fitness;
  vsr_result* results;
  int n_robots;
  // trace (robot 0 only)
  double *tr_vertices, *tr_bodies, *tr_ctrl;
  int trace_steps;
  // shared-memory layout (bytes from the CTA's dynamic smem base; per warp strides)
  int pool_slots;
  int smem_warp_stride; // bytes per warp
  int off_pool_p, off_pool_s, off_pool_m, off_pool_cnt, off_w, off_act, off_sig;
};

// ------------------------------------------------------------------------------------------------
// small math helpers (precise versions; the FP32 build uses no fast-math intrinsics except rsqrt-free sqrt)
// ------------------------------------------------------------------------------------------------
template <class R> struct M;
template <> struct M<float> {
  static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
  static __device__ __forceinline__ void sincos_(float a, float* s, float* c) { sincosf(a, s, c); }
  static __device__ __forceinline__ float tanh_(float x) { return tanhf(x); }
  static __device__ __forceinline__ float atan2_(float y, float x) { return atan2f(y, x); }
  static __device__ __forceinline__ float abs_(float x) { return fabsf(x); }
  static __device__ __forceinline__ float max_(float a, float b) { return fmaxf(a, b); }
  static __device__ __forceinline__ float min_(float a, float b) { return fminf(a, b); }
  static __device__ __forceinline__ float rint_(float a) { return rintf(a); }
};
template <> struct M<double> {
  static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
  static __device__ __forceinline__ void sincos_(double a, double* s, double* c) { sincos(a, s, c); }
  static __device__ __forceinline__ double tanh_(double x) { return tanh(x); }
  static __device__ __forceinline__ double atan2_(double y, double x) { return atan2(y, x); }
  static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
  static __device__ __forceinline__ double max_(double a, double b) { return fmax(a, b); }
  static __device__ __forceinline__ double min_(double a, double b) { return fmin(a, b); }
  static __device__ __forceinline__ double rint_(double a) { return rint(a); }
};

#define FULL 0xffffffffu

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

// rope table: X(index, bodyA, bodyB)
#define VSR_ROPES(X) X(0, 0, 1) X(1, 2, 3) X(2, 1, 2) X(3, 3, 0)

// ------------------------------------------------------------------------------------------------
// narrowphase: corner-mass box vs one ground polygon (SAT + reference-face clipping)
// ground polygon CCW: V0=(x0,base) V1=(x1,base) V2=(x1,y1) V3=(x0,y0)
// ------------------------------------------------------------------------------------------------
template <class R>
__device__ __noinline__ int collide_box_ground(R x0, R y0, R x1, R y1, R base, const R* bx, const R* by, R c, R s,
                                               R lin_tol, R* onx, R* ony, R* opx, R* opy, R* odepth, int* oid) {
  const R gx[4] = {x0, x1, x1, x0}, gy[4] = {base, base, y1, y0};
  R gnx[4], gny[4];
  {
    R ex = x0 - x1, ey = y0 - y1;
    R len = M<R>::sqrt_(ex * ex + ey * ey);
    gnx[0] = 0; gny[0] = -1;
    gnx[1] = 1; gny[1] = 0;
    gnx[2] = ey / len; gny[2] = -ex / len;
    gnx[3] = -1; gny[3] = 0;
  }
  const R bnx[4] = {s, c, -s, -c}, bny[4] = {-c, s, c, -s};
  // separation of the box from each ground edge (bottom edge 0 can never be the maximum: it is
  // ground_thickness below the surface)
  R sA = -1e30f;
  int eA = 1;
#pragma unroll
  for (int i = 1; i < 4; i++) {
    R sm = 1e30f;
#pragma unroll
    for (int j = 0; j < 4; j++) sm = M<R>::min_(sm, gnx[i] * (bx[j] - gx[i]) + gny[i] * (by[j] - gy[i]));
    if (sm > sA) { sA = sm; eA = i; }
  }
  if (sA > 0) return 0;
  R sB = -1e30f;
  int eB = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    R sm = 1e30f;
#pragma unroll
    for (int j = 0; j < 4; j++) sm = M<R>::min_(sm, bnx[i] * (gx[j] - bx[i]) + bny[i] * (gy[j] - by[i]));
    if (sm > sB) { sB = sm; eB = i; }
  }
  if (sB > 0) return 0;
  const int flip = sB > sA + R(0.1) * lin_tol;
  const int e = flip ? eB : eA;
  R nx, ny, v1x, v1y, v2x, v2y;
  if (flip) {
    nx = bnx[e]; ny = bny[e];
    v1x = bx[e]; v1y = by[e]; v2x = bx[(e + 1) & 3]; v2y = by[(e + 1) & 3];
  } else {
    nx = gnx[e]; ny = gny[e];
    v1x = gx[e]; v1y = gy[e]; v2x = gx[(e + 1) & 3]; v2y = gy[(e + 1) & 3];
  }
  int ie = 0;
  R mind = 1e30f;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    R d = flip ? nx * gnx[j] + ny * gny[j] : nx * bnx[j] + ny * bny[j];
    if (d < mind) { mind = d; ie = j; }
  }
  R cx[2], cy[2];
  if (flip) {
    cx[0] = gx[ie]; cy[0] = gy[ie]; cx[1] = gx[(ie + 1) & 3]; cy[1] = gy[(ie + 1) & 3];
  } else {
    cx[0] = bx[ie]; cy[0] = by[ie]; cx[1] = bx[(ie + 1) & 3]; cy[1] = by[(ie + 1) & 3];
  }
  R tx = v2x - v1x, ty = v2y - v1y;
  R tl = M<R>::sqrt_(tx * tx + ty * ty);
  tx /= tl; ty /= tl;
  {
    R o = tx * v1x + ty * v1y;
    R d0 = tx * cx[0] + ty * cy[0] - o, d1 = tx * cx[1] + ty * cy[1] - o;
    if (d0 < 0 && d1 < 0) return 0;
    if (d0 < 0 || d1 < 0) {
      R u = d0 / (d0 - d1);
      R ix = cx[0] + u * (cx[1] - cx[0]), iy = cy[0] + u * (cy[1] - cy[0]);
      if (d0 < 0) { cx[0] = ix; cy[0] = iy; } else { cx[1] = ix; cy[1] = iy; }
    }
  }
  {
    R o = tx * v2x + ty * v2y;
    R d0 = o - (tx * cx[0] + ty * cy[0]), d1 = o - (tx * cx[1] + ty * cy[1]);
    if (d0 < 0 && d1 < 0) return 0;
    if (d0 < 0 || d1 < 0) {
      R u = d0 / (d0 - d1);
      R ix = cx[0] + u * (cx[1] - cx[0]), iy = cy[0] + u * (cy[1] - cy[0]);
      if (d0 < 0) { cx[0] = ix; cy[0] = iy; } else { cx[1] = ix; cy[1] = iy; }
    }
  }
  int np = 0;
#pragma unroll
  for (int q = 0; q < 2; q++) {
    R sep = nx * (cx[q] - v1x) + ny * (cy[q] - v1y);
    if (sep <= 0) {
      opx[np] = cx[q]; opy[np] = cy[q]; odepth[np] = -sep;
      oid[np] = (flip << 5) | (e << 3) | (ie << 1) | q;
      np++;
    }
  }
  *onx = flip ? -nx : nx;
  *ony = flip ? -ny : ny;
  return np;
}

// ------------------------------------------------------------------------------------------------
// the episode kernel
// ------------------------------------------------------------------------------------------------
template <class R, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) vsr_episode_kernel(const __grid_constant__ KParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* wsm = smem_raw + (size_t)warp * P.smem_warp_stride;
  R* poolP = reinterpret_cast<R*>(wsm + P.off_pool_p);   // [2][kPoolPersist][S]
  R* poolS = reinterpret_cast<R*>(wsm + P.off_pool_s);   // [kPoolSolve][S]
  int* poolM = reinterpret_cast<int*>(wsm + P.off_pool_m); // [2][S]
  int* poolCnt = reinterpret_cast<int*>(wsm + P.off_pool_cnt);
  R* wts = reinterpret_cast<R*>(wsm + P.off_w);           // [robots_per_warp][genome_len]
  R* act = reinterpret_cast<R*>(wsm + P.off_act);         // centralized: [rpw][2][maxN]; distributed: [2][maxN][32]
  R* sig = reinterpret_cast<R*>(wsm + P.off_sig);         // distributed: [2][32][4*signals]
  const int S = P.pool_slots;

  const int L = P.lanes_per_robot;
  const int rw = lane / L;                     // robot slot inside the warp
  const int v = lane - rw * L;                 // voxel index inside the robot
  const bool lane_used = rw < P.robots_per_warp;
  const int robot = (blockIdx.x * WARPS + warp) * P.robots_per_warp + rw;
  const bool valid = lane_used && robot < P.n_robots;
  const int base_lane = rw * L;
  const int gl = (blockIdx.x * WARPS + warp) * 32 + lane; // global lane (history ring column)
  const int total_lanes = gridDim.x * WARPS * 32;

  // neighbours as lane ids (self when absent, so shuffles stay in range)
  const int vi = valid ? v : 0;
  const int ln = (valid && P.nb_l[vi] >= 0) ? base_lane + P.nb_l[vi] : -1;
  const int rn = (valid && P.nb_r[vi] >= 0) ? base_lane + P.nb_r[vi] : -1;
  const int dn = (valid && P.nb_d[vi] >= 0) ? base_lane + P.nb_d[vi] : -1;
  const int un = (valid && P.nb_u[vi] >= 0) ? base_lane + P.nb_u[vi] : -1;
  const int ln_s = ln < 0 ? lane : ln, rn_s = rn < 0 ? lane : rn, dn_s = dn < 0 ? lane : dn, un_s = un < 0 ? lane : un;

  // constants
  const R dt = (R)P.dt, half = (R)P.half, im = (R)P.inv_m, ii = (R)P.inv_i, mu = (R)P.mu;
  const R lin_tol = (R)P.lin_tol, ang_tol = (R)P.ang_tol, max_corr = (R)P.max_lin_corr, baum = (R)P.baumgarte;
  const R gdt = valid ? (R)(P.g * P.dt) : R(0);
  const R lin_f = (R)P.lin_f, ang_f = (R)P.ang_f;
  const R sp_sigma = (R)P.sp_sigma, sp_beta = (R)P.sp_beta, sp_cfrac = (R)P.sp_cfrac;
  const R rope_lower = (R)P.rope_lower, rope_E = R(1) / (im + im);
  const int scaff = P.scaff_mask;

  // ---- state ----------------------------------------------------------------------------------
  // per-lane integer origin (world = origin + local); corner masses NW NE SE SW
  R ox, oy;
  R X[4], Y[4], A[4], CS[4], SN[4], VX[4], VY[4], W[4];
  {
    const double cx0 = valid ? P.vox_x0[vi] : 0.0, cy0 = valid ? P.vox_y0[vi] : 1.0e4;
    ox = (R)rint(cx0); oy = (R)rint(cy0);
    const R lx = (R)(cx0 - (double)ox), ly = (R)(cy0 - (double)oy);
    const R off = (R)(P.side_len / 2 - P.half);
    const R sxk[4] = {-1, +1, +1, -1}, syk[4] = {+1, +1, -1, -1};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      X[k] = lx + sxk[k] * off; Y[k] = ly + syk[k] * off;
      A[k] = 0; CS[k] = 1; SN[k] = 0; VX[k] = 0; VY[k] = 0; W[k] = 0;
    }
  }
  R rest[3] = {(R)P.range_rest[0], (R)P.range_rest[1], (R)P.range_rest[2]};
  R simp[18];
#pragma unroll
  for (int i = 0; i < 18; i++) simp[i] = 0;
  R rimp[4] = {0, 0, 0, 0};
  R wimp[4][3];
#pragma unroll
  for (int i = 0; i < 4; i++) wimp[i][0] = wimp[i][1] = wimp[i][2] = 0;
  // contact slot lists: per mass [cnt:8 | s2:8 | s1:8 | s0:8]
  unsigned clist[4] = {0, 0, 0, 0};
  int par = 0; // parity of the current persistent pool buffer
  int kh = 1;  // terrain polygon hint
  int overflow = 0;

  // ---- controller state --------------------------------------------------------------------------
  R phase = 0;
  if (P.ctrl_kind == VSR_CTRL_PHASE_SIN) {
    if (valid) phase = (R)P.genomes[(size_t)robot * P.genome_len + v];
  } else {
    // genome upload: the warp's robots are consecutive rows -> one contiguous range of doubles.
    const int r0 = (blockIdx.x * WARPS + warp) * P.robots_per_warp;
    int nr = P.n_robots - r0;
    nr = nr < 0 ? 0 : (nr > P.robots_per_warp ? P.robots_per_warp : nr);
    const size_t first = (size_t)r0 * P.genome_len, count = (size_t)nr * P.genome_len;
    const double* src = P.genomes + first;
    // 128-bit vectorised body (double2), scalar head/tail for alignment
    size_t head = ((reinterpret_cast<uintptr_t>(src) & 15) && count) ? 1 : 0;
    for (size_t i = lane; i < head; i += 32) wts[i] = (R)src[i];
    const size_t nvec = (count - head) / 2;
    const double2* src2 = reinterpret_cast<const double2*>(src + head);
    for (size_t i = lane; i < nvec; i += 32) {
      double2 d = __ldg(&src2[i]);
      wts[head + 2 * i] = (R)d.x;
      wts[head + 2 * i + 1] = (R)d.y;
    }
    for (size_t i = head + 2 * nvec + lane; i < count; i += 32) wts[i] = (R)src[i];
    if (P.ctrl_kind == VSR_CTRL_MLP_DISTRIBUTED_HOMO)
      for (int i = lane; i < 2 * 32 * 4 * P.signals; i += 32) sig[i] = 0;
  }
  if (lane == 0) *poolCnt = 0;
  __syncwarp();
  double avg_sum[kMaxAvg];
  int avg_cnt[kMaxAvg];
#pragma unroll
  for (int i = 0; i < kMaxAvg; i++) { avg_sum[i] = 0; avg_cnt[i] = 0; }
  int sig_par = 0;
  double cx_first = 0, cx_last = 0, t_first = 0, t_last = 0;
  R ctrl_f = 0;

  // ---- helpers (lambdas capture state by reference; everything is inlined) -----------------------
  auto pP = [&](int p, int f, int s) -> R& { return poolP[(p * kPoolPersist + f) * S + s]; };
  auto pS = [&](int f, int s) -> R& { return poolS[f * S + s]; };

  // detection of all 4 corner masses of this lane against the ground; fills the pool (parity par^1)
  auto detect = [&]() {
    __syncwarp();
    if (lane == 0) *poolCnt = 0;
    __syncwarp();
    const int np_ = par ^ 1;
    const double oxd = (double)ox, oyd = (double)oy;
    unsigned nlist[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      nlist[k] = 0;
      if (!valid) continue;
      const R ext = half * (M<R>::abs_(CS[k]) + M<R>::abs_(SN[k]));
      const double ax0 = oxd + (double)(X[k] - ext), ax1 = oxd + (double)(X[k] + ext), ay0 = oyd + (double)(Y[k] - ext);
      // smallest polygon index kk (polygon kk spans tx[kk-1]..tx[kk]) with tx[kk] >= ax0
      int kk = kh;
      while (kk > 1 && __ldg(&P.tx[kk - 1]) >= ax0) kk--;
      while (kk < P.n_t - 1 && __ldg(&P.tx[kk]) < ax0) kk++;
      if (k == 0) kh = kk;
      R bx[4], by[4];
      {
        const R hc = half * CS[k], hs = half * SN[k];
        bx[0] = X[k] - hc + hs; by[0] = Y[k] - hs - hc;
        bx[1] = X[k] + hc + hs; by[1] = Y[k] + hs - hc;
        bx[2] = X[k] + hc - hs; by[2] = Y[k] + hs + hc;
        bx[3] = X[k] - hc - hs; by[3] = Y[k] - hs + hc;
      }
      int cnt = 0;
      for (; kk <= P.n_t - 1 && cnt < kMaxMan; kk++) {
        const double x0 = __ldg(&P.tx[kk - 1]);
        if (x0 > ax1) break;
        const double x1 = __ldg(&P.tx[kk]), y0 = __ldg(&P.ty[kk - 1]), y1 = __ldg(&P.ty[kk]);
        if (ax1 < x0 || ax0 > x1) continue;
        if (ay0 > fmax(y0, y1)) continue;
        R nx, ny, px[2], py[2], dp[2];
        int id[2];
        const int np = collide_box_ground<R>((R)(x0 - oxd), (R)(y0 - oyd), (R)(x1 - oxd), (R)(y1 - oyd),
                                             (R)(P.base_y - oyd), bx, by, CS[k], SN[k], lin_tol, &nx, &ny, px, py, dp, id);
        if (!np) continue;
        const int s = atomicAdd(poolCnt, 1);
        if (s >= S) { overflow++; continue; }
        R jn[2] = {0, 0}, jt[2] = {0, 0};
        // warm start from the previous manifold with the same polygon
        const int ocnt = clist[k] >> 24;
        for (int o = 0; o < ocnt; o++) {
          const int os = (clist[k] >> (8 * o)) & 0xff;
          const int ometa = poolM[par * S + os];
          if ((ometa & 0xffff) != kk) continue;
          const int onp = (ometa >> 16) & 3;
          for (int q = 0; q < np; q++)
            for (int r = 0; r < onp; r++)
              if (((ometa >> (18 + 6 * r)) & 63) == id[q]) {
                jn[q] = pP(par, 2 + 5 * r + 3, os);
                jt[q] = pP(par, 2 + 5 * r + 4, os);
              }
        }
        pP(np_, 0, s) = nx; pP(np_, 1, s) = ny;
        for (int q = 0; q < np; q++) {
          pP(np_, 2 + 5 * q + 0, s) = px[q]; pP(np_, 2 + 5 * q + 1, s) = py[q]; pP(np_, 2 + 5 * q + 2, s) = dp[q];
          pP(np_, 2 + 5 * q + 3, s) = jn[q]; pP(np_, 2 + 5 * q + 4, s) = jt[q];
        }
        poolM[np_ * S + s] = kk | (np << 16) | (id[0] << 18) | ((np > 1 ? id[1] : 0) << 24);
        nlist[k] |= (unsigned)s << (8 * cnt);
        cnt++;
      }
      nlist[k] |= (unsigned)cnt << 24;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) clist[k] = nlist[k];
    par ^= 1;
    __syncwarp();
  };

  detect(); // World.step's first-step detect()

  double t = 0.0;
  // =================================================================================================
  for (int step = 1; step <= P.n_steps; step++) {
    t += P.dt;
    // ---- integrate velocities: gravity, then damping ---------------------------------------------
#pragma unroll
    for (int k = 0; k < 4; k++) {
      VY[k] += gdt;
      VX[k] *= lin_f; VY[k] *= lin_f; W[k] *= ang_f;
    }
    // ---- contacts: initialise (+ restitution bias from pre-warm-start velocities), then warm start --
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int cnt = clist[k] >> 24;
      for (int o = 0; o < cnt; o++) {
        const int s = (clist[k] >> (8 * o)) & 0xff;
        int meta = poolM[par * S + s];
        int np = (meta >> 16) & 3;
        const R nx = pP(par, 0, s), ny = pP(par, 1, s), tx = ny, ty = -nx;
        R rn[2] = {0, 0};
        for (int q = 0; q < np; q++) {
          const R rx = pP(par, 2 + 5 * q, s) - X[k], ry = pP(par, 2 + 5 * q + 1, s) - Y[k];
          rn[q] = rx * ny - ry * nx;
          const R rt = rx * ty - ry * tx;
          pS(7 * q + 0, s) = rx; pS(7 * q + 1, s) = ry;
          pS(7 * q + 2, s) = R(1) / (im + ii * rn[q] * rn[q]);
          pS(7 * q + 3, s) = R(1) / (im + ii * rt * rt);
          const R vx = VX[k] - W[k] * ry, vy = VY[k] + W[k] * rx;
          const R vn = nx * vx + ny * vy;
          pS(7 * q + 4, s) = vn < -(R)P.rest_vel ? -(R)P.rest_e * vn : R(0);
          pS(7 * q + 5, s) = CS[k] * rx + SN[k] * ry;
          pS(7 * q + 6, s) = -SN[k] * rx + CS[k] * ry;
        }
        meta &= ~(1 << 30);
        if (np == 2) {
          const R K11 = im + ii * rn[0] * rn[0], K22 = im + ii * rn[1] * rn[1], K12 = im + ii * rn[0] * rn[1];
          const R det = K11 * K22 - K12 * K12;
          if (K11 * K11 < R(100) * det) {
            meta |= 1 << 30;
            const R idet = det != 0 ? R(1) / det : R(0);
            pS(14, s) = K12; pS(15, s) = K22 * idet; pS(16, s) = -K12 * idet; pS(17, s) = K11 * idet;
          } else {
            np = 1;
            meta = (meta & ~(3 << 16)) | (1 << 16);
          }
        }
        poolM[par * S + s] = meta;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int cnt = clist[k] >> 24;
      for (int o = 0; o < cnt; o++) {
        const int s = (clist[k] >> (8 * o)) & 0xff;
        const int np = (poolM[par * S + s] >> 16) & 3;
        const R nx = pP(par, 0, s), ny = pP(par, 1, s), tx = ny, ty = -nx;
        for (int q = 0; q < np; q++) {
          const R jn = pP(par, 2 + 5 * q + 3, s), jt = pP(par, 2 + 5 * q + 4, s);
          const R Jx = nx * jn + tx * jt, Jy = ny * jn + ty * jt;
          VX[k] += im * Jx; VY[k] += im * Jy; W[k] += ii * (pS(7 * q, s) * Jy - pS(7 * q + 1, s) * Jx);
        }
      }
    }
    // ---- springs: initialise + warm start ---------------------------------------------------------
    R HC[4], HS[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { HC[k] = half * CS[k]; HS[k] = half * SN[k]; }
    R snx[18], sny[18], sc1[18], sc2[18], sE[18], sB[18];
#define SPRING_INIT(i, a, b, sx1, sy1, sx2, sy2, kind, bit)                                     \
  if (scaff & bit) {                                                                            \
    const R r1x = (sx1)*HC[a] - (sy1)*HS[a], r1y = (sx1)*HS[a] + (sy1)*HC[a];                   \
    const R r2x = (sx2)*HC[b] - (sy2)*HS[b], r2y = (sx2)*HS[b] + (sy2)*HC[b];                   \
    R nx = (X[a] + r1x) - (X[b] + r2x), ny = (Y[a] + r1y) - (Y[b] + r2y);                       \
    const R len = M<R>::sqrt_(nx * nx + ny * ny);                                               \
    if (len < lin_tol) { nx = 0; ny = 0; } else { const R il = R(1) / len; nx *= il; ny *= il; } \
    const R cr1 = r1x * ny - r1y * nx, cr2 = r2x * ny - r2y * nx;                               \
    const R inv_mass = im + ii * cr1 * cr1 + im + ii * cr2 * cr2;                               \
    const R E = sp_cfrac / inv_mass;                                                            \
    snx[i] = nx; sny[i] = ny; sc1[i] = cr1; sc2[i] = cr2; sE[i] = E;                            \
    sB[i] = E * ((len - rest[kind]) * sp_beta);                                                 \
    const R m_ = simp[i] * im, q_ = simp[i] * ii;                                               \
    VX[a] += nx * m_; VY[a] += ny * m_; W[a] += q_ * cr1;                                       \
    VX[b] -= nx * m_; VY[b] -= ny * m_; W[b] -= q_ * cr2;                                       \
  }
    VSR_SPRINGS(SPRING_INIT)
#undef SPRING_INIT
    // ---- ropes: initialise + warm start -------------------------------------------------------------
    R rnx[4], rny[4];
    bool ract[4];
    bool any_rope = false;
#define ROPE_INIT(i, a, b)                                                                  \
  {                                                                                         \
    R nx = X[a] - X[b], ny = Y[a] - Y[b];                                                   \
    const R len = M<R>::sqrt_(nx * nx + ny * ny);                                           \
    if (len < lin_tol) { nx = 0; ny = 0; } else { const R il = R(1) / len; nx *= il; ny *= il; } \
    rnx[i] = nx; rny[i] = ny;                                                               \
    ract[i] = P.limit_contraction && len < rope_lower;                                      \
    if (!ract[i]) rimp[i] = 0;                                                              \
    const R m_ = rimp[i] * im;                                                              \
    VX[a] += nx * m_; VY[a] += ny * m_; VX[b] -= nx * m_; VY[b] -= ny * m_;                 \
    any_rope |= ract[i];                                                                    \
  }
    VSR_ROPES(ROPE_INIT)
#undef ROPE_INIT
    const bool warp_rope = __any_sync(FULL, any_rope);
    // ---- welds: r2 of the partner mass, warm start ----------------------------------------------------
    // owner welds: 0 = (my NW, left NE) 1 = (my SW, left SE) 2 = (my SW, below NW) 3 = (my SE, below NE)
    // local anchor 2 = body1 - body2 at assembly: horizontal (+m, 0), vertical (0, +m)
    const R wa = (R)P.weld_anchor;
    R wr2x[4], wr2y[4];
    {
      const R c1 = __shfl_sync(FULL, CS[1], ln_s), s1 = __shfl_sync(FULL, SN[1], ln_s);
      const R c2 = __shfl_sync(FULL, CS[2], ln_s), s2 = __shfl_sync(FULL, SN[2], ln_s);
      wr2x[0] = wa * c1; wr2y[0] = wa * s1; wr2x[1] = wa * c2; wr2y[1] = wa * s2;
      const R c0 = __shfl_sync(FULL, CS[0], dn_s), s0 = __shfl_sync(FULL, SN[0], dn_s);
      const R c1d = __shfl_sync(FULL, CS[1], dn_s), s1d = __shfl_sync(FULL, SN[1], dn_s);
      wr2x[2] = -wa * s0; wr2y[2] = wa * c0; wr2x[3] = -wa * s1d; wr2y[3] = wa * c1d;
    }
    // apply impulse l (3) of owner weld to body1 (mine) and return the partner's delta
#define WELD_APPLY(wi, k1, lx_, ly_, lz_, dvx, dvy, dw)                   \
  {                                                                      \
    VX[k1] += im * (lx_); VY[k1] += im * (ly_); W[k1] += ii * (lz_);     \
    dvx = -im * (lx_); dvy = -im * (ly_);                                \
    dw = -ii * (wr2x[wi] * (ly_) - wr2y[wi] * (lx_) + (lz_));            \
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
    // weld solve constants (r1 = 0 closed form)
    const R wA = im + im, wh = ii * R(0.5);
    const R wD = R(1) / (wA * (wA + wh * wa * wa));
    // solve K l = -C for a weld with r1 = 0 and partner arm (r2x, r2y)
#define WELD_SOLVE(r2x, r2y, Cx, Cy, Cz, lx_, ly_, lz_)                                  \
  {                                                                                     \
    const R bx_ = -(Cx) - (r2y) * (Cz) * R(0.5), by_ = -(Cy) + (r2x) * (Cz) * R(0.5);   \
    const R k00 = wA + wh * (r2x) * (r2x), k01 = wh * (r2x) * (r2y), k11 = wA + wh * (r2y) * (r2y); \
    lx_ = wD * (k00 * bx_ + k01 * by_);                                                 \
    ly_ = wD * (k01 * bx_ + k11 * by_);                                                 \
    lz_ = -(Cz) / (ii + ii) + ((r2y)*lx_ - (r2x)*ly_) * R(0.5);                         \
  }

    // ---- velocity iterations ------------------------------------------------------------------------
    for (int it = 0; it < P.vel_iters; it++) {
#define SPRING_SOLVE(i, a, b, sx1, sy1, sx2, sy2, kind, bit)                             \
  if (scaff & bit) {                                                                    \
    R jv = snx[i] * (VX[a] - VX[b]);                                                    \
    jv += sny[i] * (VY[a] - VY[b]);                                                     \
    jv += sc1[i] * W[a];                                                                \
    jv -= sc2[i] * W[b];                                                                \
    const R l_ = sE[i] * jv + sB[i] + sp_sigma * simp[i]; /* = -lambda */               \
    simp[i] -= l_;                                                                      \
    const R m_ = l_ * im, q_ = l_ * ii;                                                 \
    VX[a] -= snx[i] * m_; VY[a] -= sny[i] * m_; W[a] -= q_ * sc1[i];                    \
    VX[b] += snx[i] * m_; VY[b] += sny[i] * m_; W[b] += q_ * sc2[i];                    \
  }
      VSR_SPRINGS(SPRING_SOLVE)
#undef SPRING_SOLVE
      if (warp_rope) {
#define ROPE_SOLVE(i, a, b)                                                        \
  if (ract[i]) {                                                                   \
    const R jv = rnx[i] * (VX[a] - VX[b]) + rny[i] * (VY[a] - VY[b]);              \
    R l_ = -rope_E * jv;                                                           \
    const R old = rimp[i];                                                         \
    rimp[i] = M<R>::max_(old + l_, R(0));                                          \
    l_ = rimp[i] - old;                                                            \
    const R m_ = l_ * im;                                                          \
    VX[a] += rnx[i] * m_; VY[a] += rny[i] * m_; VX[b] -= rnx[i] * m_; VY[b] -= rny[i] * m_; \
  }
        VSR_ROPES(ROPE_SOLVE)
#undef ROPE_SOLVE
      }
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
      // contacts of my 4 corner masses
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int cnt = clist[k] >> 24;
        for (int o = 0; o < cnt; o++) {
          const int s = (clist[k] >> (8 * o)) & 0xff;
          const int meta = poolM[par * S + s];
          const int np = (meta >> 16) & 3;
          const R nx = pP(par, 0, s), ny = pP(par, 1, s), tx = ny, ty = -nx;
          for (int q = 0; q < np; q++) { // friction first
            const R rx = pS(7 * q, s), ry = pS(7 * q + 1, s);
            const R vx = VX[k] - W[k] * ry, vy = VY[k] + W[k] * rx;
            R l_ = -pS(7 * q + 3, s) * (tx * vx + ty * vy);
            const R maxf = mu * pP(par, 2 + 5 * q + 3, s);
            const R old = pP(par, 2 + 5 * q + 4, s);
            const R nj = M<R>::max_(-maxf, M<R>::min_(old + l_, maxf));
            l_ = nj - old;
            pP(par, 2 + 5 * q + 4, s) = nj;
            const R Jx = tx * l_, Jy = ty * l_;
            VX[k] += im * Jx; VY[k] += im * Jy; W[k] += ii * (rx * Jy - ry * Jx);
          }
          if (!(meta & (1 << 30))) {
            for (int q = 0; q < np; q++) {
              const R rx = pS(7 * q, s), ry = pS(7 * q + 1, s);
              const R vx = VX[k] - W[k] * ry, vy = VY[k] + W[k] * rx;
              R l_ = -pS(7 * q + 2, s) * (nx * vx + ny * vy - pS(7 * q + 4, s));
              const R old = pP(par, 2 + 5 * q + 3, s);
              const R nj = M<R>::max_(old + l_, R(0));
              l_ = nj - old;
              pP(par, 2 + 5 * q + 3, s) = nj;
              const R Jx = nx * l_, Jy = ny * l_;
              VX[k] += im * Jx; VY[k] += im * Jy; W[k] += ii * (rx * Jy - ry * Jx);
            }
          } else { // 2-point block solver
            const R r1x = pS(0, s), r1y = pS(1, s), r2x = pS(7, s), r2y = pS(8, s);
            const R a1 = pP(par, 5, s), a2 = pP(par, 10, s);
            const R vn1 = nx * (VX[k] - W[k] * r1y) + ny * (VY[k] + W[k] * r1x);
            const R vn2 = nx * (VX[k] - W[k] * r2y) + ny * (VY[k] + W[k] * r2x);
            const R K11 = R(1) / pS(2, s), K22 = R(1) / pS(9, s), K12 = pS(14, s);
            const R b1 = vn1 - pS(4, s) - (K11 * a1 + K12 * a2);
            const R b2 = vn2 - pS(11, s) - (K12 * a1 + K22 * a2);
            R x1 = -(pS(15, s) * b1 + pS(16, s) * b2), x2 = -(pS(16, s) * b1 + pS(17, s) * b2);
            if (!(x1 >= 0 && x2 >= 0)) {
              x1 = -b1 * pS(2, s); x2 = 0;
              if (!(x1 >= 0 && K12 * x1 + b2 >= 0)) {
                x1 = 0; x2 = -b2 * pS(9, s);
                if (!(x2 >= 0 && K12 * x2 + b1 >= 0)) {
                  x1 = 0; x2 = 0;
                  if (!(b1 >= 0 && b2 >= 0)) { x1 = a1; x2 = a2; }
                }
              }
            }
            const R d1 = x1 - a1, d2 = x2 - a2;
            pP(par, 5, s) = x1; pP(par, 10, s) = x2;
            const R J1x = nx * d1, J1y = ny * d1, J2x = nx * d2, J2y = ny * d2;
            VX[k] += im * (J1x + J2x); VY[k] += im * (J1y + J2y);
            W[k] += ii * ((r1x * J1y - r1y * J1x) + (r2x * J2y - r2y * J2x));
          }
        }
      }
    } // velocity iterations

    // ---- integrate positions (per-step clamps) -----------------------------------------------------
#pragma unroll
    for (int k = 0; k < 4; k++) {
      R txx = VX[k] * dt, tyy = VY[k] * dt, r = W[k] * dt;
      const R t2 = txx * txx + tyy * tyy;
      if (t2 > (R)(P.max_trans * P.max_trans)) {
        const R ratio = (R)P.max_trans / M<R>::sqrt_(t2);
        VX[k] *= ratio; VY[k] *= ratio; txx = VX[k] * dt; tyy = VY[k] * dt;
      }
      if (M<R>::abs_(r) > (R)P.max_rot) {
        const R ratio = (R)P.max_rot / M<R>::abs_(r);
        W[k] *= ratio; r = W[k] * dt;
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
      for (int it = 0; it < P.pos_iters; it++) {
        bool ok = true;
        // contacts
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int cnt = active ? (clist[k] >> 24) : 0;
          for (int o = 0; o < cnt; o++) {
            const int s = (clist[k] >> (8 * o)) & 0xff;
            const int np = (poolM[par * S + s] >> 16) & 3;
            const R nx = pP(par, 0, s), ny = pP(par, 1, s);
            for (int q = 0; q < np; q++) {
              const R lx = pS(7 * q + 5, s), ly = pS(7 * q + 6, s);
              const R rx = CS[k] * lx - SN[k] * ly, ry = SN[k] * lx + CS[k] * ly;
              const R sep = nx * (X[k] + rx - pP(par, 2 + 5 * q, s)) + ny * (Y[k] + ry - pP(par, 2 + 5 * q + 1, s)) -
                            pP(par, 2 + 5 * q + 2, s);
              if (sep < -R(3) * lin_tol) ok = false;
              R C = baum * (sep + lin_tol);
              C = M<R>::min_(C, R(0));
              C = M<R>::max_(C, -max_corr);
              const R rn_ = rx * ny - ry * nx;
              const R K = im + ii * rn_ * rn_;
              const R l_ = K > 0 ? -C / K : R(0);
              X[k] += im * nx * l_; Y[k] += im * ny * l_; A[k] += ii * rn_ * l_;
              M<R>::sincos_(A[k], &SN[k], &CS[k]);
            }
          }
        }
        // ropes
        if (warp_rope) {
#define ROPE_POS(i, a, b)                                                                  \
  if (ract[i] && active) {                                                                 \
    R nx = X[a] - X[b], ny = Y[a] - Y[b];                                                  \
    const R len = M<R>::sqrt_(nx * nx + ny * ny);                                          \
    if (len < lin_tol) { nx = 0; ny = 0; } else { const R il = R(1) / len; nx *= il; ny *= il; } \
    R C = len - rope_lower;                                                                \
    if (!(-C < lin_tol)) ok = false;                                                       \
    C = M<R>::min_(C, R(0));                                                               \
    C = M<R>::max_(C, -max_corr);                                                          \
    const R l_ = -rope_E * C;                                                              \
    const R m_ = l_ * im;                                                                  \
    X[a] += nx * m_; Y[a] += ny * m_; X[b] -= nx * m_; Y[b] -= ny * m_;                    \
  }
          VSR_ROPES(ROPE_POS)
#undef ROPE_POS
        }
        // weld position solve for owner weld (body1 = mine k1, partner pose pulled)
#define WELD_POS(k1, px_, py_, pa_, ax_, ay_, dpx, dpy, dpa)                                     \
  {                                                                                             \
    R s2, c2;                                                                                   \
    M<R>::sincos_(pa_, &s2, &c2);                                                               \
    const R r2x = c2 * (ax_) - s2 * (ay_), r2y = s2 * (ax_) + c2 * (ay_);                       \
    const R Cx = X[k1] - ((px_) + r2x), Cy = Y[k1] - ((py_) + r2y), Cz = A[k1] - (pa_);         \
    if (!(M<R>::sqrt_(Cx * Cx + Cy * Cy) <= lin_tol && M<R>::abs_(Cz) <= ang_tol)) ok = false;  \
    R lx_, ly_, lz_;                                                                            \
    WELD_SOLVE(r2x, r2y, Cx, Cy, Cz, lx_, ly_, lz_)                                             \
    X[k1] += im * lx_; Y[k1] += im * ly_; A[k1] += ii * lz_;                                    \
    M<R>::sincos_(A[k1], &SN[k1], &CS[k1]);                                                     \
    dpx = -im * lx_; dpy = -im * ly_; dpa = -ii * (r2x * ly_ - r2y * lx_ + lz_);                \
  }
        {
          const R p0x = __shfl_sync(FULL, X[1], ln_s) + lox, p0y = __shfl_sync(FULL, Y[1], ln_s) + loy, p0a = __shfl_sync(FULL, A[1], ln_s);
          const R p1x = __shfl_sync(FULL, X[2], ln_s) + lox, p1y = __shfl_sync(FULL, Y[2], ln_s) + loy, p1a = __shfl_sync(FULL, A[2], ln_s);
          R d0x = 0, d0y = 0, d0a = 0, d1x = 0, d1y = 0, d1a = 0;
          if (ln >= 0 && active) {
            WELD_POS(0, p0x, p0y, p0a, wa, R(0), d0x, d0y, d0a)
            WELD_POS(3, p1x, p1y, p1a, wa, R(0), d1x, d1y, d1a)
          }
          const R a0 = __shfl_sync(FULL, d0x, rn_s), a1 = __shfl_sync(FULL, d0y, rn_s), a2 = __shfl_sync(FULL, d0a, rn_s);
          const R b0 = __shfl_sync(FULL, d1x, rn_s), b1 = __shfl_sync(FULL, d1y, rn_s), b2 = __shfl_sync(FULL, d1a, rn_s);
          if (rn >= 0 && active) {
            X[1] += a0; Y[1] += a1; A[1] += a2; M<R>::sincos_(A[1], &SN[1], &CS[1]);
            X[2] += b0; Y[2] += b1; A[2] += b2; M<R>::sincos_(A[2], &SN[2], &CS[2]);
          }
        }
        {
          const R p0x = __shfl_sync(FULL, X[0], dn_s) + dox, p0y = __shfl_sync(FULL, Y[0], dn_s) + doy, p0a = __shfl_sync(FULL, A[0], dn_s);
          const R p1x = __shfl_sync(FULL, X[1], dn_s) + dox, p1y = __shfl_sync(FULL, Y[1], dn_s) + doy, p1a = __shfl_sync(FULL, A[1], dn_s);
          R d0x = 0, d0y = 0, d0a = 0, d1x = 0, d1y = 0, d1a = 0;
          if (dn >= 0 && active) {
            WELD_POS(3, p0x, p0y, p0a, R(0), wa, d0x, d0y, d0a)
            WELD_POS(2, p1x, p1y, p1a, R(0), wa, d1x, d1y, d1a)
          }
          const R e0 = __shfl_sync(FULL, d0x, un_s), e1 = __shfl_sync(FULL, d0y, un_s), e2 = __shfl_sync(FULL, d0a, un_s);
          const R f0 = __shfl_sync(FULL, d1x, un_s), f1 = __shfl_sync(FULL, d1y, un_s), f2 = __shfl_sync(FULL, d1a, un_s);
          if (un >= 0 && active) {
            X[0] += e0; Y[0] += e1; A[0] += e2; M<R>::sincos_(A[0], &SN[0], &CS[0]);
            X[1] += f0; Y[1] += f1; A[1] += f2; M<R>::sincos_(A[1], &SN[1], &CS[1]);
          }
        }
#undef WELD_POS
        // robot-wide convergence
        const unsigned bad = __ballot_sync(FULL, active && !ok);
        if (!(bad & robot_mask)) active = false;
        if (!__any_sync(FULL, active)) break;
      }
    }

    // ---- re-centre the lane origin (integer shifts; keeps |local| small) -------------------------------
    {
      const R mx = R(0.25) * (X[0] + X[1] + X[2] + X[3]), my = R(0.25) * (Y[0] + Y[1] + Y[2] + Y[3]);
      if (M<R>::abs_(mx) > R(1) || M<R>::abs_(my) > R(1)) {
        const R sx = M<R>::rint_(mx), sy = M<R>::rint_(my);
        ox += sx; oy += sy;
#pragma unroll
        for (int k = 0; k < 4; k++) { X[k] -= sx; Y[k] -= sy; }
      }
    }
    // ---- detect for the next step / touch sensors -----------------------------------------------------
    detect();

    // ---- robot.act(t): sensors -> controller -> applyForce ---------------------------------------------
    // outer voxel-poly vertices (NW NE SE SW), local frame
    R qx[4], qy[4];
    {
      const R sxk[4] = {-1, +1, +1, -1}, syk[4] = {+1, +1, -1, -1};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const R lx = sxk[k] * half, ly = syk[k] * half;
        qx[k] = X[k] + CS[k] * lx - SN[k] * ly;
        qy[k] = Y[k] + SN[k] * lx + CS[k] * ly;
      }
    }
    R f = 0;
    if (P.ctrl_kind == VSR_CTRL_PHASE_SIN) {
      f = (R)(P.amp * sin(2.0 * 3.14159265358979323846 * P.freq * t + (double)phase));
    } else {
      // sensor readings of this voxel
      R rd[8];
      R vang = 0;
      bool have_ang = false;
      for (int si = 0; si < P.n_sensors; si++) {
        const SensorK& sk = P.sensors[si];
        R raw[2] = {0, 0};
        if (sk.kind == VSR_SENSOR_AREA_RATIO) {
          R a2 = 0;
#pragma unroll
          for (int i = 0; i < 4; i++) { const int j = (i + 1) & 3; a2 += qx[i] * qy[j] - qx[j] * qy[i]; }
          raw[0] = M<R>::abs_(a2) * R(0.5) / (R)(P.side_len * P.side_len);
        } else if (sk.kind == VSR_SENSOR_TOUCH) {
          raw[0] = ((clist[0] | clist[1] | clist[2] | clist[3]) >> 24) ? R(1) : R(0);
        } else {
          if (!have_ang) {
            vang = (M<R>::atan2_(Y[1] - Y[0], X[1] - X[0]) + M<R>::atan2_(Y[2] - Y[3], X[2] - X[3])) * R(0.5);
            have_ang = true;
          }
          if (sk.kind == VSR_SENSOR_ANGLE) raw[0] = vang;
          else {
            R vx = R(0.25) * (VX[0] + VX[1] + VX[2] + VX[3]), vy = R(0.25) * (VY[0] + VY[1] + VY[2] + VY[3]);
            if (sk.rotated) {
              R sn_, cs_;
              M<R>::sincos_(-vang, &sn_, &cs_);
              const R rx = cs_ * vx - sn_ * vy, ry = sn_ * vx + cs_ * vy;
              vx = rx; vy = ry;
            }
            int o = 0;
            if (sk.axes & 1) raw[o++] = vx;
            if (sk.axes & 2) raw[o++] = vy;
          }
        }
        for (int d = 0; d < sk.dim; d++) {
          R val = raw[d];
          if (sk.avg >= 0) {
            // Average: mean of the last win readings (window table emulates the t_j >= t - w test in double)
            const int ai = sk.avg + d;
            const int slot = (step - 1) % P.hist_cap;
            R* hist = reinterpret_cast<R*>(P.hist);
            hist[((size_t)slot * P.n_avg + ai) * total_lanes + gl] = val;
            double sum = avg_sum[ai] + (double)val;
            int cnt = avg_cnt[ai] + 1;
            const int win = P.win_tab[(size_t)ai * (P.n_steps + 1) + step];
            while (cnt > win) {
              const int oslot = (step - cnt + P.hist_cap * 4) % P.hist_cap;
              sum -= (double)hist[((size_t)oslot * P.n_avg + ai) * total_lanes + gl];
              cnt--;
            }
            avg_sum[ai] = sum; avg_cnt[ai] = cnt;
            val = (R)(sum / cnt);
          }
          if (sk.soft_norm) {
            const R vv = (val - (R)sk.dom_min) / (R)(sk.dom_max - sk.dom_min);
            val = M<R>::tanh_((vv * R(2) - R(1)) * R(2)) * R(0.5) + R(0.5);
          }
          rd[sk.off + d] = val;
        }
      }
      const int D = P.sens_dim, maxN = P.max_neurons;
      if (P.ctrl_kind == VSR_CTRL_MLP_CENTRALIZED) {
        R* a0 = act + (size_t)(lane_used ? rw : 0) * 2 * maxN;
        const R* w_ = wts + (size_t)(lane_used ? rw : 0) * P.genome_len;
        if (valid)
          for (int d = 0; d < D; d++) a0[v * D + d] = P.in_act ? M<R>::tanh_(rd[d]) : rd[d];
        __syncwarp();
        int cur = 0, o = 0;
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
        R* a0 = act;                    // [2][maxN][32]
        const R* w_ = wts + (size_t)(lane_used ? rw : 0) * P.genome_len;
        const R* sin_ = sig + (size_t)sig_par * 32 * 4 * sg;
        R* sout = sig + (size_t)(sig_par ^ 1) * 32 * 4 * sg;
        int ni = 0;
        for (int d = 0; d < D; d++) a0[(ni++) * 32 + lane] = P.in_act ? M<R>::tanh_(rd[d]) : rd[d];
        // Dir N(0,-1) E(1,0) S(0,1) W(-1,0); read the neighbour's signal emitted towards me (opposite side)
        const int nbl[4] = {dn, rn, un, ln};
        const int opp[4] = {2, 3, 0, 1};
        for (int d = 0; d < 4; d++)
          for (int s_ = 0; s_ < sg; s_++) {
            const R x_ = nbl[d] >= 0 ? sin_[(nbl[d] * 4 + opp[d]) * sg + s_] : R(0);
            a0[(ni++) * 32 + lane] = P.in_act ? M<R>::tanh_(x_) : x_;
          }
        int cur = 0, o = 0;
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
    // Voxel.applyForce(f), ForceMethod.DISTANCE
    if (M<R>::abs_(f) > R(1)) f = f > 0 ? R(1) : R(-1);
    ctrl_f = f;
#pragma unroll
    for (int kq = 0; kq < 3; kq++) {
      const R rr = (R)P.range_rest[kq];
      rest[kq] = f >= 0 ? rr - (rr - (R)P.range_min[kq]) * f : rr + ((R)P.range_max[kq] - rr) * -f;
    }

    // ---- observation: robot centre x at the first / last kept step --------------------------------------
    if (step == P.k_first || step == P.k_last) {
      const double cxv = (double)ox + (double)(R(0.25) * (qx[0] + qx[1] + qx[2] + qx[3]));
      double sum = 0;
      for (int j = 0; j < L; j++) sum += __shfl_sync(FULL, cxv, (base_lane + j) & 31);
      sum /= L;
      if (step == P.k_first) { cx_first = sum; t_first = t; }
      if (step == P.k_last) { cx_last = sum; t_last = t; }
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

  // ---- epilogue: K3 result pack --------------------------------------------------------------------------
  const int ovf = __reduce_add_sync(FULL, overflow);
  if (valid && v == 0) {
    const double dist = cx_last - cx_first, tm = t_last - t_first;
    const double vel = dist / tm;
    P.fitness[robot] = vel;
    if (P.results) {
      vsr_result r;
      r.velocity = vel; r.distance = dist; r.time = tm; r.first_x = cx_first; r.last_x = cx_last;
      r.n_steps = P.n_steps; r.k_first = P.k_first; r.k_last = P.k_last; r.contact_overflow = ovf;
      P.results[robot] = r;
    }
  }
#undef WELD_APPLY
#undef WELD_SOLVE
}

} // namespace vsr
