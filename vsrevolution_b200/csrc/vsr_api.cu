// vsr_api.cu -- C ABI of libvsr_eval.so: contexts, model assembly, device buffers, kernel launches.
// Boundary documented in include/vsr_eval.h; replaces Starter.buildLocomotionTask (Starter.java:83-94).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "vsr_internal.h"
#include "vsr_kernel.cuh"

constexpr float kHeightBucket = 4.0f; // width of a bucket of the terrain height map (corner masses are under 1 unit wide)
namespace {

// launch shapes (warps per CTA, min CTAs per SM -> register cap): 0/2 = 255 registers, 8 warps/SM (3552 bipeds per wave);
// 1/3 = 168 registers, 10 warps/SM (4440 bipeds per wave).  Default 3: ONE CTA of 10 warps per SM, so that all resident
// warps are phase-aligned by the same barriers and share one instruction stream.
// VSR_LAUNCH_SHAPE overrides (tuning only).
struct Shape { int warps, minb; };
constexpr Shape kShapes[] = {{4, 2}, {5, 2}, {8, 1}, {10, 1}};
constexpr int kNumShapes = sizeof(kShapes) / sizeof(kShapes[0]);
constexpr int kDefaultShape = 3;

thread_local std::string g_create_error;

struct DeviceState {
  int dev = 0;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev_copy = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_join = nullptr; // eval_range: this device's slice is back on the primary device
  double *d_tx = nullptr, *d_ty = nullptr;
  float* d_thmax = nullptr; // coarse height map of the terrain (contact broadphase)
  uint8_t* d_win = nullptr;
  double* d_noise = nullptr;
  void* d_hist = nullptr;
  size_t hist_bytes = 0;
  void* d_energy = nullptr; // per-lane energy accumulators (results mode only)
  size_t energy_bytes = 0;
  // per-evaluation terrains (vsr_eval_terrains): vertices of all robots of the shard, offsets, ground bottoms, placement shifts
  double *d_rtx = nullptr, *d_rty = nullptr, *d_rbase = nullptr, *d_rplace = nullptr;
  int* d_roff = nullptr;
  size_t rt_vertices = 0, rt_robots = 0;
  double* d_genomes = nullptr;
  double* d_fitness = nullptr;
  vsr_result* d_results = nullptr;
  double* h_genomes = nullptr; // pinned staging
  double* h_fitness = nullptr;
  vsr_result* h_results = nullptr;
  size_t cap = 0; // robots the buffers hold
  int* d_ovf = nullptr;        // warps whose contact pool overflowed, accumulated over all launches (vsr_contact_overflows)
  double* d_centers = nullptr; // vsr_eval_centers: [robots of the shard][kept observations][2]
  size_t centers_bytes = 0;
  double* cur_centers = nullptr; // non-null only while vsr_eval_centers enqueues
  bool timed = false;
  // vsr_eval_ex: robot transformations and per-observation summaries
  uint32_t* d_frozen = nullptr; size_t frozen_cap = 0;
  unsigned int* d_post = nullptr; size_t post_cap = 0;       // [robots][64] posture counters
  uint8_t* d_fp[2] = {nullptr, nullptr}; size_t fp_bytes[2] = {0, 0}; // [robots][K]
  double* d_cen3 = nullptr; size_t cen3_bytes = 0;           // [robots][K][3]
  bool obs_attr = false; // dynamic shared-memory opt-in of the observation kernel done on this device
};

// what one launch carries besides genomes -> fitness (vsr_eval_ex / vsr_eval_terrains)
struct LaunchExtras {
  bool robot_terrains = false;
  double rt_ymax = 0;
  const uint32_t* d_frozen = nullptr;
  bool obs = false;
  int post_n = 0, fp_n[2] = {0, 0};
  unsigned int* d_post = nullptr;
  uint8_t* d_fp[2] = {nullptr, nullptr};
  double* d_cen3 = nullptr;
};

} // namespace

// validated host-side view of the per-robot inputs of one batch (vsr_eval_ex: the caller's arrays; vsr_stage_eval_opts: owned copies)
struct HostExtras {
  int n = 0;
  const uint32_t* frozen = nullptr;
  const double *tx = nullptr, *ty = nullptr;
  const int32_t* off = nullptr;
  std::vector<double> base, place; // per robot: bottom of its ground body; placement shift relative to the context's terrain
  double ymax = 0;
  // owned storage (staging only)
  std::vector<uint32_t> o_frozen;
  std::vector<double> o_tx, o_ty;
  std::vector<int32_t> o_off;
  void clear() { *this = HostExtras(); }
};

struct vsr_ctx {
  vsr_env_desc env;
  std::vector<double> tx, ty;
  std::vector<float> thmax; float thx0 = 0; // coarse height map (build_height_map)
  bool use_height_map = true;
  // body
  bool has_body = false, has_ctrl = false;
  int w = 0, h = 0, n_vox = 0;
  std::vector<uint8_t> mask;
  std::vector<int> cell_of_vox, vox_of_cell;
  std::vector<vsr_sensor_desc> sensors;      // the distinct sensors of the body (position-dependent fields zeroed)
  std::vector<vsr_sensor_desc> vsens;        // every voxel's own list, concatenated in voxel (grid) order
  std::vector<int> vsens_off;                // [n_vox + 1]
  std::vector<int> vsens_u;                  // [vsens.size()] index into `sensors`
  bool uniform_sensors = true;               // every voxel carries the same list
  bool full_scaff = false;                   // all four scaffoldings enabled: builds with the mask folded at compile time
  bool area_robot = false;                   // FP32, the reference default: one soft-normalised AreaRatio reading per voxel into a centralized MLP
  // controller
  vsr_ctrl_kind kind = VSR_CTRL_PHASE_SIN;
  std::vector<int> layers;
  int signals = 0, in_act = 1;
  double freq = 1, amp = 1;
  // derived
  bool ready = false;
  vsr::KParams kp;
  std::vector<uint8_t> win_tab;
  std::vector<double> noise_tab; // Noisy sensors: [sens_dim][n_steps+1]
  size_t smem_bytes = 0;
  // sub-wave batches (fewer robot-warps than fill the 10-warp kernel's wave): the 255-register build launched with just enough warps
  // per CTA to spread the batch over every SM (tools/latency_probe.sh: 1.2-1.5x lower episode latency below ~3500 bipeds)
  const void* kernel_lat = nullptr;
  int warps_lat = 0;        // most warps per CTA the latency kernel may get (shared memory / launch bound)
  size_t smem_warp = 0;     // shared-memory bytes per warp
  int n_sm = 148;
  const void* last_kernel = nullptr; // launch shape of the most recent episode-kernel launch (vsr_kernel_info)
  int last_warps = 0;
  size_t last_smem = 0;
  int shape = kDefaultShape, warps = 4;
  const void* kernel = nullptr;
  std::vector<DeviceState> devs;
  std::string err;
  int64_t launches = 0;
  double place_dx = 0, place_dy = 0; // initial placement on the context's own terrain
  int central_vox = 0;               // BehaviorUtils.getCentralElement of the body grid (voxel order)
  HostExtras staged;                 // vsr_stage_eval_opts: per-robot inputs of the next device-resident batches (n = 0: none)
};

namespace {

int fail(vsr_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  else g_create_error = msg;
  return code;
}

#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess)                                                                           \
      return fail(ctx, VSR_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));               \
  } while (0)

double ground_y_at(const double* tx, const double* ty, int n, double x) {
  for (int i = 1; i < n; i++)
    if (x >= tx[i - 1] && x <= tx[i]) {
      double u = (x - tx[i - 1]) / (tx[i] - tx[i - 1]);
      return ty[i - 1] + u * (ty[i] - ty[i - 1]);
    }
  return -INFINITY;
}

// Locomotion.apply: translate the robot so that its bounding box starts place_x_gap right of the terrain's second vertex and
// its lowest voxel is place_y_gap above the ground under that voxel's centre [U]; (dx, dy) is added to the grid positions
bool placement(const vsr_ctx* ctx, const double* tx, const double* ty, int nt, double* dx_out, double* dy_out) {
  const vsr_env_desc& e = ctx->env;
  const double L = e.side_length;
  double minx = INFINITY;
  for (int v = 0; v < ctx->n_vox; v++) minx = std::fmin(minx, (ctx->cell_of_vox[v] % ctx->w) * L - L / 2);
  const double dx = (tx[1] + e.place_x_gap) - minx;
  double min_gap = INFINITY;
  for (int v = 0; v < ctx->n_vox; v++) {
    int cell = ctx->cell_of_vox[v], x = cell % ctx->w, y = cell / ctx->w;
    double cx = x * L + dx, miny = y * L - L / 2;
    min_gap = std::fmin(min_gap, miny - ground_y_at(tx, ty, nt, cx));
  }
  const double dy = e.place_y_gap - min_gap;
  *dx_out = dx; *dy_out = dy;
  return std::isfinite(dy) && std::isfinite(dx);
}

int sensor_dim(const vsr_sensor_desc& s) {
  if (s.kind == VSR_SENSOR_VELOCITY) return ((s.axes & 1) ? 1 : 0) + ((s.axes & 2) ? 1 : 0);
  if (s.kind == VSR_SENSOR_LIDAR) return s.n_rays;
  return 1;
}

int voxel_dim(const vsr_ctx* c, int v) {
  int dim = 0;
  for (int i = c->vsens_off[v]; i < c->vsens_off[v + 1]; i++) dim += sensor_dim(c->vsens[i]);
  return dim;
}

int genome_len_of(const vsr_ctx* c) {
  if (!c->has_body || !c->has_ctrl) return VSR_ESTATE;
  int total = 0, dim = voxel_dim(c, 0);
  bool same = true;
  for (int v = 0; v < c->n_vox; v++) { const int d = voxel_dim(c, v); total += d; same &= d == dim; }
  if (c->kind == VSR_CTRL_PHASE_SIN) return c->n_vox;
  if (c->kind == VSR_CTRL_TIME_SIN) return 3 * c->n_vox;
  int n_in = c->layers.front(), n_out = c->layers.back();
  if (c->kind == VSR_CTRL_MLP_CENTRALIZED) {
    if (n_in != total || n_out != c->n_vox) return VSR_EDIM;
  } else {
    // DistributedSensing: one MLP input size, so every voxel must produce the same number of readings
    if (!same || n_in != dim + 4 * c->signals || n_out != 1 + 4 * c->signals) return VSR_EDIM;
  }
  const int nw = vsr_mlp_count_weights(c->layers.data(), (int)c->layers.size());
  return c->kind == VSR_CTRL_MLP_DISTRIBUTED_HETERO ? c->n_vox * nw : nw;
}

// the episode kernel is instantiated for three controller families; the two extra kinds run inside them
bool is_sin(vsr_ctrl_kind k) { return k == VSR_CTRL_PHASE_SIN || k == VSR_CTRL_TIME_SIN; }
bool is_dist(vsr_ctrl_kind k) { return k == VSR_CTRL_MLP_DISTRIBUTED_HOMO || k == VSR_CTRL_MLP_DISTRIBUTED_HETERO; }
int kernel_ctrl(vsr_ctrl_kind k) {
  return is_sin(k) ? VSR_CTRL_PHASE_SIN : (is_dist(k) ? VSR_CTRL_MLP_DISTRIBUTED_HOMO : VSR_CTRL_MLP_CENTRALIZED);
}

size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

// kernel instantiations live in vsr_kernel_inst.cu (one object per launch shape, compiled in parallel)
constexpr int kWarps64 = 4;
const void* kernel_ptr(int precision, int ctrl, int shape, int full_scaff) {
  if (precision == 64) return vsr_internal::kernel_f64(ctrl);
  switch (shape) {
    case 0: return vsr_internal::kernel_f32_s0(ctrl, full_scaff);
    case 1: return vsr_internal::kernel_f32_s1(ctrl, full_scaff);
    case 2: return vsr_internal::kernel_f32_s2(ctrl, full_scaff);
    default: return vsr_internal::kernel_f32_s3(ctrl, full_scaff);
  }
}

// (re)derive everything that depends on env + body + controller
int finalize(vsr_ctx* ctx) {
  ctx->ready = false;
  if (!ctx->has_body || !ctx->has_ctrl) return VSR_OK;
  const vsr_env_desc& e = ctx->env;
  int glen = genome_len_of(ctx);
  if (glen < 0) return fail(ctx, glen, "controller I/O dimensions do not match the body/sensors (VSR_EDIM)");
  vsr::KParams& k = ctx->kp;
  std::memset(&k, 0, sizeof(k));
  k.dt = e.dt; k.g = e.gravity_y; k.baumgarte = e.baumgarte; k.lin_tol = e.linear_tol; k.ang_tol = e.angular_tol;
  k.max_lin_corr = e.max_linear_corr; k.max_trans = e.max_translation; k.max_rot = e.max_rotation;
  k.rest_vel = e.restitution_velocity; k.vel_iters = e.vel_iters; k.pos_iters = e.pos_iters;
  const double L = e.side_length, m = L * e.mass_side_ratio;
  k.side_len = L;
  k.half = m / 2;
  const double density = (e.voxel_mass / 4) / (m * m), mass = density * m * m, inertia = mass * (m * m + m * m) / 12.0;
  k.inv_m = 1.0 / mass; k.inv_i = 1.0 / inertia;
  k.mu = std::sqrt(e.voxel_friction * e.ground_friction);
  k.rest_e = std::fmax(e.voxel_restitution, e.ground_restitution);
  k.lin_f = std::fmin(std::fmax(1.0 - e.dt * e.lin_damping, 0.0), 1.0);
  k.ang_f = std::fmin(std::fmax(1.0 - e.dt * e.ang_damping, 0.0), 1.0);
  const double min_side = std::sqrt(L * L * (1.0 - e.area_ratio_max_delta)), max_side = std::sqrt(L * L * (1.0 + e.area_ratio_max_delta));
  k.range_min[0] = min_side - 2 * m; k.range_rest[0] = L - 2 * m; k.range_max[0] = max_side - 2 * m;
  k.range_min[1] = std::sqrt(m * m + k.range_min[0] * k.range_min[0]);
  k.range_rest[1] = std::sqrt(m * m + k.range_rest[0] * k.range_rest[0]);
  k.range_max[1] = std::sqrt(m * m + k.range_max[0] * k.range_max[0]);
  k.range_min[2] = (min_side - m) * std::sqrt(2.0); k.range_rest[2] = (L - m) * std::sqrt(2.0); k.range_max[2] = (max_side - m) * std::sqrt(2.0);
  k.rope_lower = min_side - m;
  {
    const double w = 2.0 * M_PI * e.spring_f, c = e.dt * (2.0 * e.spring_d * w + e.dt * w * w);
    k.sp_sigma = 1.0 / (c + 1.0);
    k.sp_cfrac = c / (c + 1.0);
    k.sp_beta = e.dt * w * w / c;
  }
  k.weld_anchor = m;
  {
    vsr::KConstF& c = k.cf;
    c.c_dt = (float)k.dt; c.c_half = (float)k.half; c.c_im = (float)k.inv_m; c.c_ii = (float)k.inv_i; c.c_mu = (float)k.mu;
    c.c_lin_tol = (float)k.lin_tol; c.c_ang_tol = (float)k.ang_tol; c.c_max_corr = (float)k.max_lin_corr; c.c_baum = (float)k.baumgarte;
    c.c_gdt = (float)(k.g * k.dt); c.c_lin_f = (float)k.lin_f; c.c_ang_f = (float)k.ang_f;
    c.c_sp_sigma = (float)k.sp_sigma; c.c_sp_boversig = (float)(k.sp_beta / k.sp_sigma); c.c_sp_cfrac = (float)k.sp_cfrac;
    c.c_rope_lower = (float)k.rope_lower; c.c_rope_E = (float)(1.0 / (2.0 * k.inv_m));
    c.c_wa = (float)m; c.c_wA = (float)(2.0 * k.inv_m); c.c_wh = (float)(0.5 * k.inv_i);
    c.c_wDA = (float)(1.0 / (2.0 * k.inv_m + 0.5 * k.inv_i * m * m));
    c.c_wDh = (float)((0.5 * k.inv_i) / ((2.0 * k.inv_m) * (2.0 * k.inv_m + 0.5 * k.inv_i * m * m)));
    c.c_w_i2i = (float)(1.0 / (2.0 * k.inv_i));
    c.c_rest_vel = (float)k.rest_vel; c.c_rest_e = (float)k.rest_e;
    c.c_max_trans = (float)k.max_trans; c.c_max_trans2 = (float)(k.max_trans * k.max_trans); c.c_max_rot = (float)k.max_rot;
    c.c_inv_area2 = (float)(0.5 / (L * L)); c.c_mass_off = (float)(L / 2 - k.half); c.c_base_pad = 0;
    for (int i = 0; i < 3; i++) { c.c_range_min[i] = (float)k.range_min[i]; c.c_range_rest[i] = (float)k.range_rest[i]; c.c_range_max[i] = (float)k.range_max[i]; }
  }
  k.scaff_mask = e.scaffolding_mask; k.limit_contraction = e.limit_contraction;
  // body
  k.n_vox = ctx->n_vox;
  k.lanes_per_robot = ctx->n_vox;
  k.robots_per_warp = 32 / ctx->n_vox;
  for (int v = 0; v < ctx->n_vox; v++) {
    int cell = ctx->cell_of_vox[v], x = cell % ctx->w, y = cell / ctx->w;
    auto nb = [&](int ax, int ay) -> int8_t {
      if (ax < 0 || ax >= ctx->w || ay < 0 || ay >= ctx->h) return -1;
      return (int8_t)ctx->vox_of_cell[ay * ctx->w + ax];
    };
    k.nb_l[v] = nb(x - 1, y); k.nb_r[v] = nb(x + 1, y); k.nb_d[v] = nb(x, y - 1); k.nb_u[v] = nb(x, y + 1);
  }
  { // BehaviorUtils.getCentralElement [U]: the voxel closest to the mean grid position, first minimum in grid order
    double mx = 0, my = 0;
    for (int v = 0; v < ctx->n_vox; v++) { mx += ctx->cell_of_vox[v] % ctx->w; my += ctx->cell_of_vox[v] / ctx->w; }
    mx /= ctx->n_vox; my /= ctx->n_vox;
    double best = INFINITY;
    for (int v = 0; v < ctx->n_vox; v++) {
      const double ddx = ctx->cell_of_vox[v] % ctx->w - mx, ddy = ctx->cell_of_vox[v] / ctx->w - my, dd = ddx * ddx + ddy * ddy;
      if (dd < best) { best = dd; ctx->central_vox = v; }
    }
  }
  // Locomotion.apply initial placement
  {
    double dx, dy;
    if (!placement(ctx, ctx->tx.data(), ctx->ty.data(), (int)ctx->tx.size(), &dx, &dy)) return fail(ctx, VSR_EINVAL, "robot is placed outside the terrain");
    ctx->place_dx = dx; ctx->place_dy = dy;
    for (int v = 0; v < ctx->n_vox; v++) {
      int cell = ctx->cell_of_vox[v];
      k.vox_x0[v] = (cell % ctx->w) * L + dx;
      k.vox_y0[v] = (cell / ctx->w) * L + dy;
    }
  }
  k.n_t = (int)ctx->tx.size();
  k.base_y = *std::min_element(ctx->ty.begin(), ctx->ty.end()) - e.ground_thickness;
  k.ty_max = *std::max_element(ctx->ty.begin(), ctx->ty.end());
  k.thx0 = ctx->thx0; k.thinv = 1.0f / kHeightBucket; k.thn = (int)ctx->thmax.size(); k.thmax = nullptr; // device pointer set per launch
  // episode
  int ns, kf, kl;
  int rc = vsr_episode_steps(e.final_t, e.transient_t, e.dt, &ns, &kf, &kl);
  if (rc) return fail(ctx, rc, "bad episode times");
  if (!kf) return fail(ctx, VSR_EINVAL, "no observation falls inside [transient_t, final_t)");
  // a single kept observation has distance 0 over time 0: Outcome.getVelocity() would be 0/0
  if (kf == kl) return fail(ctx, VSR_EINVAL, "only one observation falls inside [transient_t, final_t): velocity is undefined");
  k.n_steps = ns; k.k_first = kf; k.k_last = kl;
  {
    volatile double t = 0.0;
    for (int s = 1; s <= ns; s++) { t = t + e.dt; if (s == kf) k.t_first = t; if (s == kl) k.t_last = t; }
  }
  // sensors: the distinct sensors (k.sensors) and every voxel's list of them (k.vs_*)
  k.n_sensors = (int)ctx->sensors.size();
  int navg = 0, nrows = 0;
  std::vector<double> windows;
  std::vector<int> row_of_dim(16, -1); // Noisy rows depend on the sensor's dimension only (every wrapper replays `new Random(0)`)
  for (int i = 0; i < k.n_sensors; i++) {
    const vsr_sensor_desc& s = ctx->sensors[i];
    vsr::SensorK& sk = k.sensors[i];
    sk.kind = s.kind; sk.axes = s.axes; sk.rotated = s.rotated; sk.soft_norm = s.soft_norm;
    sk.dom_min = s.dom_min; sk.dom_inv = s.soft_norm ? 1.0 / (s.dom_max - s.dom_min) : 1.0;
    sk.dim = sensor_dim(s); sk.off = 0;
    sk.n_rays = s.n_rays; sk.param = s.param;
    sk.avg = -1;
    if (s.avg_window > 0) {
      sk.avg = navg;
      for (int d = 0; d < sk.dim; d++) windows.push_back(s.avg_window);
      navg += sk.dim;
    }
    sk.noise_row = 0;
    if (s.noise_sigma > 0) {
      if (row_of_dim[sk.dim] < 0) { row_of_dim[sk.dim] = nrows; nrows += sk.dim; }
      sk.noise_row = row_of_dim[sk.dim];
    }
  }
  if (navg > vsr::kMaxAvg) return fail(ctx, VSR_EUNSUPPORTED, "more than 6 averaged readings over the body's distinct sensors");
  k.vs_max = 0;
  int total_dim = 0;
  for (int v = 0; v < ctx->n_vox; v++) {
    const int n = ctx->vsens_off[v + 1] - ctx->vsens_off[v], dim = voxel_dim(ctx, v);
    if (n > vsr::kMaxVoxSensors || dim > vsr::kMaxReadings) return fail(ctx, VSR_EUNSUPPORTED, "more than 8 sensors / 12 readings on one voxel");
    k.vs_n[v] = (uint8_t)n; k.vs_max = std::max(k.vs_max, n);
    k.vox_dim[v] = (uint8_t)dim; k.vox_in_off[v] = (uint16_t)total_dim;
    total_dim += dim;
    for (int j = 0; j < n; j++) {
      const vsr_sensor_desc& s = ctx->vsens[ctx->vsens_off[v] + j];
      k.vs_idx[v][j] = (uint8_t)ctx->vsens_u[ctx->vsens_off[v] + j];
      k.vs_param[v][j] = s.kind == VSR_SENSOR_CONSTANT ? s.param : (s.kind == VSR_SENSOR_LIDAR ? (double)s.side : 0.0);
    }
  }
  if (ctx->uniform_sensors) { // offsets inside the (common) reading vector, for the specialised builds
    int off = 0;
    for (int j = 0; j < (int)k.vs_n[0]; j++) { k.sensors[k.vs_idx[0][j]].off = off; off += k.sensors[k.vs_idx[0][j]].dim; }
  }
  k.sens_dim = voxel_dim(ctx, 0); k.n_avg = navg;
  // Average windows: count of stored readings with t_j >= t_k - window, emulated in double
  ctx->win_tab.assign((size_t)std::max(navg, 1) * (ns + 1), 1);
  int max_win = 1;
  if (navg) {
    std::vector<double> ts(ns + 1, 0.0);
    volatile double t = 0.0;
    for (int s = 1; s <= ns; s++) { t = t + e.dt; ts[s] = t; }
    for (int a = 0; a < navg; a++)
      for (int s = 1; s <= ns; s++) {
        int cnt = 0;
        for (int j = s; j >= 1 && ts[j] >= ts[s] - windows[a]; j--) cnt++;
        if (cnt > 250) return fail(ctx, VSR_EUNSUPPORTED, "Average window longer than 250 steps");
        ctx->win_tab[(size_t)a * (ns + 1) + s] = (uint8_t)cnt;
        max_win = std::max(max_win, cnt);
      }
  }
  k.hist_cap = max_win + 2;
  // Noisy wrappers: `new Noisy(sensor, sigma, 0)` on every voxel [U] -> every wrapper replays java.util.Random(0): one Gaussian per
  // reading per step, scaled by sigma * |domain| of the wrapped sensor ([0,1] after a normalisation).  The table holds, per distinct
  // (dimension, sigma scale) ... rows are per dimension; the scale is folded in per sensor below, so sensors of one dimension must
  // share sigma * |domain| (true for every configuration RobotUtils builds: one sigma for the whole body)
  ctx->noise_tab.clear();
  if (nrows > 0) {
    ctx->noise_tab.assign((size_t)nrows * (ns + 1), 0.0);
    std::vector<double> scale_of_dim(16, -1.0);
    for (int i = 0; i < k.n_sensors; i++) {
      const vsr_sensor_desc& sd = ctx->sensors[i];
      if (!(sd.noise_sigma > 0)) continue;
      const int dim = k.sensors[i].dim;
      const double scale = sd.noise_sigma * (sd.soft_norm ? 1.0 : std::fabs(sd.dom_max - sd.dom_min));
      if (scale_of_dim[dim] >= 0) {
        if (scale_of_dim[dim] != scale) return fail(ctx, VSR_EUNSUPPORTED, "Noisy sensors of the same dimension with different sigma * |domain|");
        continue;
      }
      scale_of_dim[dim] = scale;
      std::vector<double> g((size_t)dim * ns);
      vsr_java_random_gaussians(0, g.data(), dim * ns);
      for (int st = 1; st <= ns; st++)
        for (int d = 0; d < dim; d++) ctx->noise_tab[(size_t)(k.sensors[i].noise_row + d) * (ns + 1) + st] = g[(size_t)(st - 1) * dim + d] * scale;
    }
  }
  // controller
  k.ctrl_kind = kernel_ctrl(ctx->kind);
  k.w_stride = ctx->kind == VSR_CTRL_MLP_DISTRIBUTED_HETERO ? glen / ctx->n_vox : 0;
  k.time_sin = ctx->kind == VSR_CTRL_TIME_SIN ? 1 : 0;
  k.signals = ctx->signals; k.in_act = ctx->in_act; k.freq = ctx->freq; k.amp = ctx->amp;
  k.genome_len = glen;
  k.n_layers = (int)ctx->layers.size();
  k.max_neurons = 1;
  for (int i = 0; i < k.n_layers; i++) { k.layers[i] = ctx->layers[i]; k.max_neurons = std::max(k.max_neurons, ctx->layers[i]); }
  // shared memory layout (per warp)
  const size_t rs = e.precision == 64 ? 8 : 4;
  // contact-pool slots per warp.  Wide polygons (flat, hilly): 6/5 manifolds per voxel lane covers every bottom mass of
  // every robot touching plus polygon-boundary doubles.  Narrow polygons (steppy: segments of 0.5-1.5): a mass can touch
  // up to 3 polygons, so 2 per lane (capped by the 8-bit slot ids).  Overflow is never silent: it is counted in
  // vsr_result.contact_overflow and the tests assert it stays 0.
  double min_seg = INFINITY;
  for (size_t i = 1; i < ctx->tx.size(); i++) min_seg = std::fmin(min_seg, ctx->tx[i] - ctx->tx[i - 1]);
  const bool narrow = min_seg < 2.0 * m;
  // Small bodies need more than 6/5 per voxel: whichever side of a w x h body lies on the ground, up to max(w, h) voxels rest on
  // two corner masses each, plus two polygon-boundary doubles (a 1x1 robot: 4 manifolds for 1 voxel, 32 robots per warp).
  const int resting = k.robots_per_warp * std::min(2 * std::max(ctx->w, ctx->h) + 2, 4 * ctx->n_vox);
  int S = narrow ? std::max(16, std::min(96, k.robots_per_warp * ctx->n_vox * 2))
                 : std::max(16, std::min(254, std::max(std::min(64, k.robots_per_warp * ctx->n_vox * 6 / 5), resting)));
  if (const char* ev = std::getenv("VSR_POOL_SLOTS")) S = std::max(8, std::min(255, std::atoi(ev)));
  k.pool_slots = S;
  size_t o = 0;
  k.off_pool_p = (int)o; o = align16(o + (size_t)vsr::kPoolStride * S * rs);
  k.off_pool_m = (int)o; o = align16(o + 2 * S * 4);
  k.off_pool_cnt = (int)o; o = align16(o + 4);
  k.off_pool_o = (int)o; o = align16(o + 128); // list of touching masses
  k.off_mslots = (int)o; o = align16(o + 32 * 4 * 4); // mass -> manifold slots
  o = (o + 31) & ~size_t(31);
  // velocity staging of the transposed contact solve and the controller's activation scratch are used in disjoint
  // phases of a step: one region, sized for the larger
  size_t act_bytes = 0;
  if (ctx->kind == VSR_CTRL_MLP_CENTRALIZED) act_bytes = (size_t)k.robots_per_warp * 2 * k.max_neurons * rs;
  if (is_dist(ctx->kind)) act_bytes = (size_t)2 * k.max_neurons * 32 * rs;
  k.off_stage = (int)o; k.off_act = (int)o;
  o = align16(o + std::max((size_t)3 * 32 * 4 * rs, act_bytes));
  k.off_w = (int)o;
  if (!is_sin(ctx->kind)) o = align16(o + (size_t)k.robots_per_warp * glen * rs);
  k.off_sig = (int)o;
  if (is_dist(ctx->kind)) o = align16(o + (size_t)2 * 32 * 4 * std::max(1, k.signals) * rs);
  o = (o + 31) & ~size_t(31); // also staged as V4 rows by the position phase
  k.off_bias = (int)o;
  o = align16(o + (size_t)18 * 32 * rs);
  o = (o + 31) & ~size_t(31);
  k.off_arm = (int)o;
  o = align16(o + (size_t)8 * 32 * 4 * rs);
  k.off_eff = (int)o;
  o = align16(o + (size_t)8 * 32 * 2 * rs);
  o = (o + 31) & ~size_t(31);
  k.off_nrm = (int)o;
  o = align16(o + (size_t)2 * 32 * 4 * rs);
  k.off_rope = (int)o;
  o = align16(o + (size_t)2 * 32 * 4 * rs);
  k.off_park = (int)o;
  o = align16(o + (size_t)vsr::kPark * 32 * rs);
  k.off_cold = (int)o;
  o = align16(o + (size_t)navg * 32 * 12); // Average running sums + counts
  o = (o + 31) & ~size_t(31);
  k.smem_warp_stride = (int)o;
  k.sync_mask = 115; // step start, before spring init, before velocity loop, before/after detect (tools/ab_run.sh sweeps)
  if (const char* ev = std::getenv("VSR_SYNC_MASK")) k.sync_mask = std::atoi(ev);
  // launch shape: one CTA of 10 warps (168 registers) when its shared memory fits the 227 KB a CTA may use; otherwise the
  // 255-register kernel with 8 warps or fewer -- e.g. narrow-step terrains with their larger contact pool, large MLPs
  ctx->shape = kDefaultShape;
  if (10 * o + 1024 > 227 * 1024) ctx->shape = 2;
  if (const char* ev = std::getenv("VSR_LAUNCH_SHAPE")) ctx->shape = std::max(0, std::min(kNumShapes - 1, std::atoi(ev)));
  ctx->warps = e.precision == 64 ? kWarps64 : kShapes[ctx->shape].warps;
  // large controllers (weights of every robot of the warp live in shared memory): launch the chosen kernel with fewer warps
  // per CTA than its launch bound (the kernel indexes by blockDim)
  // (the 10-warp build folds its warp count at compile time, so a forced shape 3 that does not fit falls back to shape 2)
  if (e.precision != 64 && ctx->shape == 3 && 10 * o + 1024 > 227 * 1024) { ctx->shape = 2; ctx->warps = kShapes[2].warps; }
  while (ctx->warps > 1 && (size_t)ctx->warps * o + 1024 > 227 * 1024) ctx->warps--;
  if (const char* ev = std::getenv("VSR_WARPS")) // tuning: fewer warps per CTA (the 255-register builds index by blockDim)
    if (e.precision != 64 && ctx->shape != 3) ctx->warps = std::max(1, std::min(ctx->warps, std::atoi(ev)));
  const bool full_scaff = e.scaffolding_mask == 0xF && !std::getenv("VSR_GENERIC_SCAFF");
  ctx->kernel = kernel_ptr(e.precision, kernel_ctrl(ctx->kind), ctx->shape, full_scaff);
  // the reference's default robot: one soft-normalised, non-averaged AreaRatio reading per voxel into a centralized MLP
  ctx->full_scaff = full_scaff;
  ctx->area_robot = e.precision == 32 && full_scaff && ctx->kind == VSR_CTRL_MLP_CENTRALIZED && ctx->uniform_sensors && ctx->sensors.size() == 1 &&
                    ctx->sensors[0].kind == VSR_SENSOR_AREA_RATIO && ctx->sensors[0].soft_norm == 1 && !(ctx->sensors[0].avg_window > 0) &&
                    !std::getenv("VSR_GENERIC_SENSORS");
  if (ctx->area_robot && ctx->shape == 3 && ctx->warps == 10) ctx->kernel = vsr_internal::kernel_f32_s3_area();
  ctx->smem_bytes = o * ctx->warps + 128; // + the CTA's phase counters (bounded-skew alignment)
  ctx->smem_warp = o;
  ctx->kernel_lat = nullptr; ctx->warps_lat = 0; ctx->last_kernel = nullptr;
  if (e.precision == 32 && !std::getenv("VSR_LAUNCH_SHAPE") && !std::getenv("VSR_WARPS") && !std::getenv("VSR_NO_LATENCY_SHAPE")) {
    if (ctx->shape == 3) { // default wave kernel = 10 warps at 168 registers; latency kernel = the 8-warp build at 255
      ctx->kernel_lat = ctx->kernel == vsr_internal::kernel_f32_s3_area() ? vsr_internal::kernel_f32_s2_area()
                                                                          : kernel_ptr(32, kernel_ctrl(ctx->kind), 2, full_scaff);
      ctx->warps_lat = kShapes[2].warps;
      while (ctx->warps_lat > 1 && (size_t)ctx->warps_lat * o + 1024 > 227 * 1024) ctx->warps_lat--;
    } else if (ctx->shape == 2) { // already the flexible 255-register build: only the warp count varies
      ctx->kernel_lat = ctx->kernel;
      ctx->warps_lat = ctx->warps;
    }
  }
  if (k.n_t > 16383) return fail(ctx, VSR_EUNSUPPORTED, "terrains with more than 16383 vertices are not supported");
  if (ctx->smem_bytes > 227 * 1024) return fail(ctx, VSR_EUNSUPPORTED, "controller too large for shared memory");
  // per-device uploads
  for (auto& d : ctx->devs) {
    CK(cudaSetDevice(d.dev));
    // the attribute belongs to the kernel function, not to this context: several live contexts (one per morphology) share an
    // instantiation, so always opt in to the CTA maximum instead of this context's own footprint
    CK(cudaFuncSetAttribute(ctx->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (ctx->kernel_lat) CK(cudaFuncSetAttribute(ctx->kernel_lat, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int n_sm = 0;
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, d.dev) == cudaSuccess && n_sm > 0) ctx->n_sm = n_sm;
    if (d.d_win) { cudaFree(d.d_win); d.d_win = nullptr; }
    CK(cudaMalloc(&d.d_win, ctx->win_tab.size()));
    CK(cudaMemcpy(d.d_win, ctx->win_tab.data(), ctx->win_tab.size(), cudaMemcpyHostToDevice));
    if (d.d_noise) { cudaFree(d.d_noise); d.d_noise = nullptr; }
    if (!ctx->noise_tab.empty()) {
      CK(cudaMalloc(&d.d_noise, ctx->noise_tab.size() * sizeof(double)));
      CK(cudaMemcpy(d.d_noise, ctx->noise_tab.data(), ctx->noise_tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
  }
  ctx->ready = true;
  return VSR_OK;
}

int ensure_capacity(vsr_ctx* ctx, DeviceState& d, size_t n) {
  CK(cudaSetDevice(d.dev));
  const vsr::KParams& k = ctx->kp;
  if (n > d.cap) {
    size_t cap = std::max(n, d.cap * 2);
    if (d.d_genomes) { cudaFree(d.d_genomes); cudaFree(d.d_fitness); cudaFree(d.d_results); cudaFreeHost(d.h_genomes); cudaFreeHost(d.h_fitness); cudaFreeHost(d.h_results); }
    d.cap = 0;
    CK(cudaMalloc(&d.d_genomes, cap * k.genome_len * sizeof(double)));
    CK(cudaMalloc(&d.d_fitness, cap * sizeof(double)));
    CK(cudaMalloc(&d.d_results, cap * sizeof(vsr_result)));
    CK(cudaMallocHost(&d.h_genomes, cap * k.genome_len * sizeof(double)));
    CK(cudaMallocHost(&d.h_fitness, cap * sizeof(double)));
    CK(cudaMallocHost(&d.h_results, cap * sizeof(vsr_result)));
    d.cap = cap;
  }
  return VSR_OK;
}

// launch shape of one batch: kernel, warps per CTA, grid
struct LaunchPlan { const void* kernel; int warps, grid; size_t smem; };
LaunchPlan plan_for(const vsr_ctx* ctx, size_t n) {
  const vsr::KParams& k = ctx->kp;
  const size_t robot_warps = (n + k.robots_per_warp - 1) / k.robots_per_warp;
  LaunchPlan p{ctx->kernel, ctx->warps, (int)((robot_warps + ctx->warps - 1) / ctx->warps), ctx->smem_bytes};
  // a batch that the latency kernel can hold in one wave (<= warps_lat warps on every SM) is spread evenly over the SMs
  if (ctx->kernel_lat && robot_warps <= (size_t)ctx->warps_lat * ctx->n_sm) {
    p.kernel = ctx->kernel_lat;
    p.warps = (int)std::max<size_t>(1, (robot_warps + ctx->n_sm - 1) / ctx->n_sm);
    p.grid = (int)((robot_warps + p.warps - 1) / p.warps);
    p.smem = ctx->smem_warp * p.warps + 128;
  }
  return p;
}

int ensure_hist(vsr_ctx* ctx, DeviceState& d, int cta_warps_total) {
  const vsr::KParams& k = ctx->kp;
  if (!k.n_avg) return VSR_OK;
  const size_t rs = ctx->env.precision == 64 ? 8 : 4;
  size_t bytes = (size_t)k.hist_cap * k.n_avg * cta_warps_total * 32 * rs;
  if (bytes > d.hist_bytes) {
    if (d.d_hist) cudaFree(d.d_hist);
    d.hist_bytes = 0;
    CK(cudaMalloc(&d.d_hist, bytes));
    d.hist_bytes = bytes;
  }
  return VSR_OK;
}

int ensure_energy(vsr_ctx* ctx, DeviceState& d, int cta_warps_total) {
  const size_t rs = ctx->env.precision == 64 ? 8 : 4;
  const size_t bytes = (size_t)cta_warps_total * 32 * 4 * rs;
  if (bytes > d.energy_bytes) {
    if (d.d_energy) cudaFree(d.d_energy);
    d.energy_bytes = 0;
    CK(cudaMalloc(&d.d_energy, bytes));
    d.energy_bytes = bytes;
  }
  return VSR_OK;
}

// enqueue the episode kernel for n robots whose genomes are at d_genomes (device)
int enqueue(vsr_ctx* ctx, DeviceState& d, const double* d_genomes, size_t n, double* d_fitness, vsr_result* d_results,
            cudaStream_t st, double* trv, double* trb, double* trc, int trace_steps, bool time_it, const LaunchExtras* x = nullptr) {
  if (n == 0) return VSR_OK;
  vsr::KParams k = ctx->kp;
  // traces (robot 0 only, parity tests) always use the context's default kernel
  LaunchPlan lp = trace_steps > 0 ? LaunchPlan{ctx->kernel, ctx->warps, (int)((((n + k.robots_per_warp - 1) / k.robots_per_warp) + ctx->warps - 1) / ctx->warps), ctx->smem_bytes}
                                  : plan_for(ctx, n);
  if (x && x->obs) {
    // the observation build, spread over the SMs: FP64 = the parity build's shape; FP32 = the flexible 8-warp shape at 255 registers, with
    // the same specialisations as the plain kernels (default scaffolding; the default robot's sensor), and for the default robot the
    // 10-warp / 168-register shape when the batch needs more than 8 robot-warps per SM
    const size_t robot_warps = (n + k.robots_per_warp - 1) / k.robots_per_warp;
    const bool f32 = ctx->env.precision != 64;
    int maxw = f32 ? 8 : kWarps64;
    while (maxw > 1 && (size_t)maxw * ctx->smem_warp + 1024 > 227 * 1024) maxw--;
    lp.kernel = !f32 || !ctx->full_scaff ? vsr_internal::kernel_obs(ctx->env.precision, kernel_ctrl(ctx->kind))
                : (ctx->area_robot ? vsr_internal::kernel_obs_area(8) : vsr_internal::kernel_obs_scaff(kernel_ctrl(ctx->kind)));
    // several waves: equal ones (two waves of 5-warp CTAs, not one of 8-warp CTAs and a nearly empty one)
    const size_t waves = std::max<size_t>(1, (robot_warps + (size_t)maxw * ctx->n_sm - 1) / ((size_t)maxw * ctx->n_sm));
    lp.warps = (int)std::min<size_t>((size_t)maxw, std::max<size_t>(1, (robot_warps + waves * ctx->n_sm - 1) / (waves * ctx->n_sm)));
    if (f32 && ctx->area_robot && maxw == 8 && robot_warps > (size_t)8 * ctx->n_sm && (size_t)10 * ctx->smem_warp + 1024 <= 227 * 1024 &&
        !std::getenv("VSR_NO_OBS_WAVE")) {
      lp.kernel = vsr_internal::kernel_obs_area(10); // folds its warp count at compile time: always launched full
      lp.warps = 10;
    }
    lp.grid = (int)((robot_warps + lp.warps - 1) / lp.warps);
    lp.smem = ctx->smem_warp * lp.warps + 128;
    if (!d.obs_attr) {
      const void* ks[] = {vsr_internal::kernel_obs(ctx->env.precision, VSR_CTRL_PHASE_SIN), vsr_internal::kernel_obs(ctx->env.precision, VSR_CTRL_MLP_CENTRALIZED),
                          vsr_internal::kernel_obs(ctx->env.precision, VSR_CTRL_MLP_DISTRIBUTED_HOMO), vsr_internal::kernel_obs_scaff(VSR_CTRL_PHASE_SIN),
                          vsr_internal::kernel_obs_scaff(VSR_CTRL_MLP_CENTRALIZED), vsr_internal::kernel_obs_scaff(VSR_CTRL_MLP_DISTRIBUTED_HOMO),
                          vsr_internal::kernel_obs_area(8), vsr_internal::kernel_obs_area(10)};
      for (const void* kf : ks) CK(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      d.obs_attr = true;
    }
  }
  const int grid = lp.grid;
  int rc = ensure_hist(ctx, d, grid * lp.warps);
  if (rc) return rc;
  if (d_results) {
    rc = ensure_energy(ctx, d, grid * lp.warps);
    if (rc) return rc;
  }
  k.energy = d.d_energy;
  k.tx = d.d_tx; k.ty = d.d_ty; k.thmax = ctx->use_height_map ? d.d_thmax : nullptr; k.win_tab = d.d_win; k.hist = d.d_hist; k.noise = d.d_noise;
  k.genomes = d_genomes; k.fitness = d_fitness; k.results = d_results; k.n_robots = (int)n;
  k.centers = d.cur_centers;
  k.overflow_count = d.d_ovf;
  if (x && x->robot_terrains) { k.tx = d.d_rtx; k.ty = d.d_rty; k.toff = d.d_roff; k.tbase = d.d_rbase; k.tplace = d.d_rplace; k.ty_max = x->rt_ymax; k.thmax = nullptr; }
  if (x) {
    k.frozen = x->d_frozen;
    if (x->obs) {
      k.post_n = x->post_n; k.fp_n0 = x->fp_n[0]; k.fp_n1 = x->fp_n[1]; k.central_vox = ctx->central_vox;
      k.post_cnt = x->d_post; k.footprints0 = x->d_fp[0]; k.footprints1 = x->d_fp[1]; k.centers3 = x->d_cen3;
    }
  }
  k.tr_vertices = trv; k.tr_bodies = trb; k.tr_ctrl = trc; k.trace_steps = trace_steps;
  if (time_it) CK(cudaEventRecord(d.ev0, st));
  void* args[] = {&k};
  CK(cudaLaunchKernel(lp.kernel, dim3(grid), dim3(lp.warps * 32), args, lp.smem, st));
  ctx->last_kernel = lp.kernel; ctx->last_warps = lp.warps; ctx->last_smem = lp.smem;
  CK(cudaGetLastError());
  if (time_it) { CK(cudaEventRecord(d.ev1, st)); d.timed = true; }
  ctx->launches++;
  return VSR_OK;
}

// validate the per-robot inputs of a batch of n robots and derive what the kernel needs per robot (ground bottom, placement shift)
int build_host_extras(vsr_ctx* ctx, int n, const vsr_eval_opts& o, HostExtras& h) {
  h.clear();
  h.n = n;
  const bool terrains = o.terrain_x || o.terrain_y || o.terrain_off;
  if (terrains && (!o.terrain_x || !o.terrain_y || !o.terrain_off)) return fail(ctx, VSR_EINVAL, "per-robot terrains need terrain_x, terrain_y and terrain_off");
  if (o.frozen_masks) {
    if (ctx->n_vox < 32)
      for (int i = 0; i < n; i++)
        if (o.frozen_masks[i] >> ctx->n_vox) return fail(ctx, VSR_EINVAL, "frozen mask names a voxel the body does not have");
    h.frozen = o.frozen_masks;
  }
  if (terrains) {
    h.base.resize((size_t)n); h.place.resize((size_t)2 * n);
    h.ymax = -INFINITY;
    if (n > 0 && o.terrain_off[0] != 0) return fail(ctx, VSR_EINVAL, "terrain_off[0] must be 0");
    for (int i = 0; i < n; i++) {
      const int a = o.terrain_off[i], b = o.terrain_off[i + 1], nt = b - a;
      if (nt < 2 || nt > 16383) return fail(ctx, VSR_EINVAL, "every terrain needs 2..16383 vertices");
      double ymin = INFINITY;
      for (int j = a; j < b; j++) {
        if (j > a && !(o.terrain_x[j] > o.terrain_x[j - 1])) return fail(ctx, VSR_EINVAL, "terrain xs must be strictly increasing");
        ymin = std::fmin(ymin, o.terrain_y[j]); h.ymax = std::fmax(h.ymax, o.terrain_y[j]);
      }
      h.base[i] = ymin - ctx->env.ground_thickness;
      double dx, dy;
      if (!placement(ctx, o.terrain_x + a, o.terrain_y + a, nt, &dx, &dy)) return fail(ctx, VSR_EINVAL, "robot is placed outside the terrain");
      h.place[2 * (size_t)i] = dx - ctx->place_dx; h.place[2 * (size_t)i + 1] = dy - ctx->place_dy;
    }
    h.tx = o.terrain_x; h.ty = o.terrain_y; h.off = o.terrain_off;
  }
  return VSR_OK;
}

// upload robots [lo, hi) of the batch's per-robot inputs to device d and point the launch at them (small synchronous copies)
int upload_extras(vsr_ctx* ctx, DeviceState& d, const HostExtras& h, size_t lo, size_t hi, LaunchExtras& x) {
  const size_t cnt = hi - lo;
  if (h.tx) {
    const size_t v0 = (size_t)h.off[lo], nv = (size_t)h.off[hi] - v0;
    if (nv > d.rt_vertices) {
      cudaFree(d.d_rtx); cudaFree(d.d_rty); d.d_rtx = d.d_rty = nullptr; d.rt_vertices = 0;
      CK(cudaMalloc(&d.d_rtx, nv * sizeof(double))); CK(cudaMalloc(&d.d_rty, nv * sizeof(double)));
      d.rt_vertices = nv;
    }
    if (cnt > d.rt_robots) {
      cudaFree(d.d_rbase); cudaFree(d.d_rplace); cudaFree(d.d_roff); d.d_rbase = d.d_rplace = nullptr; d.d_roff = nullptr; d.rt_robots = 0;
      CK(cudaMalloc(&d.d_rbase, cnt * sizeof(double))); CK(cudaMalloc(&d.d_rplace, 2 * cnt * sizeof(double)));
      CK(cudaMalloc(&d.d_roff, (cnt + 1) * sizeof(int32_t)));
      d.rt_robots = cnt;
    }
    std::vector<int32_t> off(cnt + 1);
    for (size_t r = 0; r <= cnt; r++) off[r] = h.off[lo + r] - (int32_t)v0;
    CK(cudaMemcpy(d.d_rtx, h.tx + v0, nv * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.d_rty, h.ty + v0, nv * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.d_rbase, h.base.data() + lo, cnt * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.d_rplace, h.place.data() + 2 * lo, 2 * cnt * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.d_roff, off.data(), (cnt + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
    x.robot_terrains = true; x.rt_ymax = h.ymax;
  }
  if (h.frozen) {
    if (cnt > d.frozen_cap) {
      cudaFree(d.d_frozen); d.d_frozen = nullptr; d.frozen_cap = 0;
      CK(cudaMalloc(&d.d_frozen, cnt * sizeof(uint32_t)));
      d.frozen_cap = cnt;
    }
    CK(cudaMemcpy(d.d_frozen, h.frozen + lo, cnt * sizeof(uint32_t), cudaMemcpyHostToDevice));
    x.d_frozen = d.d_frozen;
  }
  return VSR_OK;
}

// Coarse height map for the contact broadphase: bucket b holds an upper bound of the terrain height over the bucket and its two
// neighbours ([x0 + (b - 1) B, x0 + (b + 2) B]), so a corner mass (extent < 1) anywhere inside bucket b that floats above
// thmax[b] cannot touch the ground.  The bound of a polygon's top edge is the larger of its two vertices; the float value is
// rounded up and padded, the test is a filter only (the narrowphase decides), so results do not depend on it.
void build_height_map(vsr_ctx* c) {
  const size_t nt = c->tx.size();
  c->thmax.clear();
  c->thx0 = nt ? (float)std::floor(c->tx.front()) : 0.0f;
  if (nt < 2) { c->thmax.assign(1, 3.0e38f); return; }
  const double x0 = c->thx0, B = kHeightBucket;
  const int nb = std::max(1, (int)std::ceil((c->tx.back() - x0) / B) + 1);
  std::vector<double> h(nb, -1.0e300);
  for (size_t i = 1; i < nt; i++) {
    const double top = std::fmax(c->ty[i - 1], c->ty[i]);
    // buckets whose widened range intersects [tx[i-1], tx[i]]: (b - 1) B <= tx[i] - x0  and  (b + 2) B >= tx[i-1] - x0
    const int b0 = std::max(0, (int)std::floor((c->tx[i - 1] - x0) / B) - 2), b1 = std::min(nb - 1, (int)std::floor((c->tx[i] - x0) / B) + 1);
    for (int b = b0; b <= b1; b++) h[b] = std::fmax(h[b], top);
  }
  c->thmax.resize(nb);
  for (int b = 0; b < nb; b++) c->thmax[b] = h[b] < -1.0e299 ? -3.0e38f : std::nextafter((float)h[b], 3.0e38f) + 1.0e-3f;
}

} // namespace

// ---- internal surface for vsr_ga.cu (vsr_internal.h) ---------------------------------------------------
namespace vsr_internal {
int primary_device(const vsr_ctx* ctx) { return ctx->devs[0].dev; }
cudaStream_t primary_stream(const vsr_ctx* ctx) { return ctx->devs[0].stream; }
bool ready(const vsr_ctx* ctx) { return ctx && ctx->ready; }
int genome_len(const vsr_ctx* ctx) { return ctx->kp.genome_len; }
const std::string& last_error(const vsr_ctx* ctx) { return ctx->err; }

int eval_range(vsr_ctx* ctx, const double* d_genomes, size_t lo, size_t hi, double* d_fitness, cudaStream_t st) {
  if (hi <= lo) return VSR_OK;
  const size_t G = (size_t)ctx->kp.genome_len, nd = ctx->devs.size(), cnt = hi - lo, per = (cnt + nd - 1) / nd;
  DeviceState& d0 = ctx->devs[0];
  // per-robot inputs staged for this batch (vsr_stage_eval_opts): robot i of the batch uses entry i
  const HostExtras* hx = ctx->staged.n > 0 ? &ctx->staged : nullptr;
  if (hx && hi > (size_t)hx->n) return fail(ctx, VSR_EINVAL, "the staged per-robot inputs cover fewer robots than the batch evaluates");
  cudaEvent_t fork = nullptr;
  if (nd > 1) {
    // the genomes are produced on `st`: the other devices' streams wait for them
    CK(cudaSetDevice(d0.dev));
    CK(cudaEventRecord(d0.ev_join, st));
    fork = d0.ev_join;
  }
  for (size_t i = 1; i < nd; i++) {
    DeviceState& d = ctx->devs[i];
    const size_t a = std::min(hi, lo + i * per), b = std::min(hi, a + per);
    if (b <= a) continue;
    int rc = ensure_capacity(ctx, d, b - a);
    if (rc) return rc;
    CK(cudaSetDevice(d.dev));
    CK(cudaStreamWaitEvent(d.stream, fork, 0));
    CK(cudaMemcpyPeerAsync(d.d_genomes, d.dev, d_genomes + a * G, d0.dev, (b - a) * G * sizeof(double), d.stream));
    LaunchExtras x;
    if (hx) { rc = upload_extras(ctx, d, *hx, a, b, x); if (rc) return rc; }
    rc = enqueue(ctx, d, d.d_genomes, b - a, d.d_fitness, nullptr, d.stream, nullptr, nullptr, nullptr, 0, false, hx ? &x : nullptr);
    if (rc) return rc;
    CK(cudaMemcpyPeerAsync(d_fitness + a, d0.dev, d.d_fitness, d.dev, (b - a) * sizeof(double), d.stream));
    CK(cudaEventRecord(d.ev_join, d.stream));
  }
  CK(cudaSetDevice(d0.dev));
  {
    const size_t b = std::min(hi, lo + per);
    LaunchExtras x;
    if (hx) { int rc = upload_extras(ctx, d0, *hx, lo, b, x); if (rc) return rc; }
    int rc = enqueue(ctx, d0, d_genomes + lo * G, b - lo, d_fitness + lo, nullptr, st, nullptr, nullptr, nullptr, 0, true, hx ? &x : nullptr);
    if (rc) return rc;
  }
  for (size_t i = 1; i < nd; i++) {
    const size_t a = std::min(hi, lo + i * per), b = std::min(hi, a + per);
    if (b > a) CK(cudaStreamWaitEvent(st, ctx->devs[i].ev_join, 0));
  }
  return VSR_OK;
}
} // namespace vsr_internal

extern "C" {

int vsr_env_default(vsr_env_desc* e) {
  if (!e) return VSR_EINVAL;
  std::memset(e, 0, sizeof(*e));
  e->dt = 1.0 / 60.0; e->vel_iters = 10; e->pos_iters = 10; e->gravity_y = -9.8; e->baumgarte = 0.2;
  e->linear_tol = 0.005; e->angular_tol = 2.0 * M_PI / 180.0; e->max_linear_corr = 0.2; e->max_translation = 2.0;
  e->max_rotation = 0.5 * M_PI; e->restitution_velocity = 1.0;
  e->side_length = 3.0; e->mass_side_ratio = 0.3; e->spring_f = 8.0; e->spring_d = 0.3; e->lin_damping = 1.0;
  e->ang_damping = 1.0; e->voxel_mass = 1.0; e->area_ratio_max_delta = 0.225; e->voxel_friction = 100.0;
  e->voxel_restitution = 0.1; e->limit_contraction = 1; e->scaffolding_mask = 0xF;
  e->ground_friction = 0.2; e->ground_restitution = 0.0; e->ground_thickness = 50.0;
  e->place_x_gap = 1.0; e->place_y_gap = 1.0; e->final_t = 30.0; e->transient_t = 1.0;
  e->precision = 32;
  return VSR_OK;
}


int vsr_ctx_create(const vsr_env_desc* env, vsr_ctx** out) {
  vsr_ctx* ctx = nullptr; // for CK/fail before the context exists
  if (!env || !out) return fail(nullptr, VSR_EINVAL, "null argument");
  *out = nullptr;
  if (!env->terrain_x || !env->terrain_y || env->n_terrain < 2) return fail(nullptr, VSR_EINVAL, "terrain polyline missing");
  for (int i = 1; i < env->n_terrain; i++)
    if (!(env->terrain_x[i] > env->terrain_x[i - 1])) return fail(nullptr, VSR_EINVAL, "terrain xs must be strictly increasing");
  if (env->precision != 32 && env->precision != 64) return fail(nullptr, VSR_EINVAL, "precision must be 32 or 64");
  if (!(env->dt > 0) || env->vel_iters < 0 || env->pos_iters < 0 || !(env->final_t > 0)) return fail(nullptr, VSR_EINVAL, "bad settings");
  if (env->n_devices < 0 || env->n_devices > 8) return fail(nullptr, VSR_EINVAL, "n_devices out of range");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(nullptr, VSR_ECUDA, std::string("no CUDA device: ") + (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0") +
                                        " (this library has no CPU fallback)");
  vsr_ctx* c = new (std::nothrow) vsr_ctx();
  if (!c) return fail(nullptr, VSR_ENOMEM, "out of host memory");
  c->env = *env;
  c->tx.assign(env->terrain_x, env->terrain_x + env->n_terrain);
  c->ty.assign(env->terrain_y, env->terrain_y + env->n_terrain);
  c->env.terrain_x = nullptr; c->env.terrain_y = nullptr;
  build_height_map(c);
  c->use_height_map = !std::getenv("VSR_NO_HEIGHT_MAP"); // test switch: the broadphase then tests against the highest terrain vertex
  std::vector<int> ids;
  if (env->n_devices == 0) {
    int cur = 0;
    cudaGetDevice(&cur);
    ids.push_back(cur);
  } else {
    for (int i = 0; i < env->n_devices; i++) ids.push_back(env->devices[i]);
  }
  ctx = c;
  for (int id : ids) {
    if (id < 0 || id >= ndev) { int rc = fail(nullptr, VSR_EINVAL, "device id out of range"); vsr_ctx_destroy(c); return rc; }
    DeviceState d;
    d.dev = id;
    c->devs.push_back(d);
  }
  for (auto& d : c->devs) {
    auto ck = [&](cudaError_t e_, const char* what) -> bool {
      if (e_ == cudaSuccess) return true;
      g_create_error = std::string(what) + ": " + cudaGetErrorString(e_);
      return false;
    };
    bool ok = ck(cudaSetDevice(d.dev), "cudaSetDevice") && ck(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking), "cudaStreamCreate") &&
              ck(cudaStreamCreateWithFlags(&d.copy_stream, cudaStreamNonBlocking), "cudaStreamCreate") &&
              ck(cudaEventCreateWithFlags(&d.ev_copy, cudaEventDisableTiming), "cudaEventCreate") &&
              ck(cudaEventCreate(&d.ev0), "cudaEventCreate") && ck(cudaEventCreate(&d.ev1), "cudaEventCreate") &&
              ck(cudaEventCreateWithFlags(&d.ev_join, cudaEventDisableTiming), "cudaEventCreate") &&
              ck(cudaMalloc(&d.d_tx, sizeof(double) * c->tx.size()), "cudaMalloc") && ck(cudaMalloc(&d.d_ty, sizeof(double) * c->ty.size()), "cudaMalloc") &&
              ck(cudaMemcpy(d.d_tx, c->tx.data(), sizeof(double) * c->tx.size(), cudaMemcpyHostToDevice), "cudaMemcpy") &&
              ck(cudaMemcpy(d.d_ty, c->ty.data(), sizeof(double) * c->ty.size(), cudaMemcpyHostToDevice), "cudaMemcpy") &&
              ck(cudaMalloc(&d.d_ovf, sizeof(int)), "cudaMalloc") && ck(cudaMemset(d.d_ovf, 0, sizeof(int)), "cudaMemset") &&
              ck(cudaMalloc(&d.d_thmax, sizeof(float) * c->thmax.size()), "cudaMalloc") &&
              ck(cudaMemcpy(d.d_thmax, c->thmax.data(), sizeof(float) * c->thmax.size(), cudaMemcpyHostToDevice), "cudaMemcpy");
    if (!ok) { vsr_ctx_destroy(c); return VSR_ECUDA; }
  }
  *out = c;
  return VSR_OK;
}

void vsr_ctx_destroy(vsr_ctx* c) {
  if (!c) return;
  for (auto& d : c->devs) {
    cudaSetDevice(d.dev);
    if (d.stream) cudaStreamSynchronize(d.stream);
    cudaFree(d.d_tx); cudaFree(d.d_ty); cudaFree(d.d_thmax); cudaFree(d.d_win); cudaFree(d.d_hist); cudaFree(d.d_noise); cudaFree(d.d_energy);
    cudaFree(d.d_rtx); cudaFree(d.d_rty); cudaFree(d.d_rbase); cudaFree(d.d_rplace); cudaFree(d.d_roff);
    cudaFree(d.d_genomes); cudaFree(d.d_fitness); cudaFree(d.d_results); cudaFree(d.d_centers); cudaFree(d.d_ovf);
    cudaFree(d.d_frozen); cudaFree(d.d_post); cudaFree(d.d_fp[0]); cudaFree(d.d_fp[1]); cudaFree(d.d_cen3);
    cudaFreeHost(d.h_genomes); cudaFreeHost(d.h_fitness); cudaFreeHost(d.h_results);
    if (d.ev0) cudaEventDestroy(d.ev0);
    if (d.ev1) cudaEventDestroy(d.ev1);
    if (d.ev_join) cudaEventDestroy(d.ev_join);
    if (d.ev_copy) cudaEventDestroy(d.ev_copy);
    if (d.copy_stream) cudaStreamDestroy(d.copy_stream);
    if (d.stream) cudaStreamDestroy(d.stream);
  }
  delete c;
}

const char* vsr_last_error(const vsr_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int vsr_body_set(vsr_ctx* ctx, int32_t w, int32_t h, const uint8_t* mask, const vsr_sensor_desc* sensors, int32_t n_sensors) {
  if (!ctx) return VSR_EINVAL;
  if (w <= 0 || h <= 0 || !mask || n_sensors < 0 || (n_sensors > 0 && !sensors)) return fail(ctx, VSR_EINVAL, "bad body arguments");
  int nv = 0;
  for (int i = 0; i < w * h; i++) nv += mask[i] ? 1 : 0;
  if (nv == 0) return fail(ctx, VSR_EINVAL, "Target robot has no voxels");
  // the same list on every voxel
  std::vector<vsr_sensor_desc> all;
  std::vector<int32_t> off(nv + 1);
  for (int v = 0; v < nv; v++) {
    off[v] = (int32_t)all.size();
    all.insert(all.end(), sensors, sensors + n_sensors);
  }
  off[nv] = (int32_t)all.size();
  return vsr_body_set_voxels(ctx, w, h, mask, all.data(), off.data());
}

int vsr_body_set_voxels(vsr_ctx* ctx, int32_t w, int32_t h, const uint8_t* mask, const vsr_sensor_desc* sensors, const int32_t* sensor_off) {
  if (!ctx) return VSR_EINVAL;
  if (w <= 0 || h <= 0 || !mask || !sensor_off) return fail(ctx, VSR_EINVAL, "bad body arguments");
  int nv = 0;
  for (int i = 0; i < w * h; i++) nv += mask[i] ? 1 : 0;
  if (nv == 0) return fail(ctx, VSR_EINVAL, "Target robot has no voxels");
  if (nv > vsr::kMaxVox) return fail(ctx, VSR_EUNSUPPORTED, "bodies with more than 32 voxels are not supported by this build");
  if (sensor_off[0] != 0) return fail(ctx, VSR_EINVAL, "sensor_off[0] must be 0");
  for (int v = 0; v < nv; v++)
    if (sensor_off[v + 1] < sensor_off[v]) return fail(ctx, VSR_EINVAL, "sensor_off must be non-decreasing");
  const int total = sensor_off[nv];
  if (total > 0 && !sensors) return fail(ctx, VSR_EINVAL, "bad body arguments");
  for (int v = 0; v < nv; v++)
    if (sensor_off[v + 1] - sensor_off[v] > vsr::kMaxVoxSensors) return fail(ctx, VSR_EUNSUPPORTED, "too many sensors per voxel");
  for (int i = 0; i < total; i++) {
    const vsr_sensor_desc& d = sensors[i];
    if (d.noise_sigma < 0) return fail(ctx, VSR_EINVAL, "negative noise sigma");
    if (d.kind < 0 || d.kind > VSR_SENSOR_LIDAR) return fail(ctx, VSR_EINVAL, "unknown sensor kind");
    if (d.kind == VSR_SENSOR_VELOCITY && !(d.axes & 3)) return fail(ctx, VSR_EINVAL, "velocity sensor without axes");
    if (d.soft_norm < 0 || d.soft_norm > 2) return fail(ctx, VSR_EINVAL, "soft_norm must be 0 (none), 1 (SoftNormalization) or 2 (Normalization)");
    if (d.soft_norm && !(d.dom_max > d.dom_min)) return fail(ctx, VSR_EINVAL, "empty sensor domain");
    if (d.kind == VSR_SENSOR_LIDAR && (d.n_rays < 1 || d.n_rays > 8 || d.side < 0 || d.side > 3 || !(d.param > 0)))
      return fail(ctx, VSR_EINVAL, "lidar needs 1..8 rays, a side in 0..3 and a positive ray length");
    if (d.kind == VSR_SENSOR_LIDAR && d.avg_window > 0) return fail(ctx, VSR_EUNSUPPORTED, "averaged lidar readings are not built");
  }
  // the distinct sensors of the body: position-dependent fields (CONSTANT value, LIDAR side) are kept per voxel, not per sensor
  std::vector<vsr_sensor_desc> uniq;
  std::vector<int> which((size_t)total);
  for (int i = 0; i < total; i++) {
    vsr_sensor_desc d = sensors[i];
    if (d.kind == VSR_SENSOR_CONSTANT) d.param = 0;
    if (d.kind == VSR_SENSOR_LIDAR) d.side = 0;
    int u = -1;
    for (size_t j = 0; j < uniq.size() && u < 0; j++)
      if (uniq[j].kind == d.kind && uniq[j].axes == d.axes && uniq[j].rotated == d.rotated && uniq[j].soft_norm == d.soft_norm &&
          uniq[j].dom_min == d.dom_min && uniq[j].dom_max == d.dom_max && uniq[j].avg_window == d.avg_window &&
          uniq[j].noise_sigma == d.noise_sigma && uniq[j].param == d.param && uniq[j].n_rays == d.n_rays) u = (int)j;
    if (u < 0) { u = (int)uniq.size(); uniq.push_back(d); }
    which[i] = u;
  }
  if ((int)uniq.size() > vsr::kMaxSensors) return fail(ctx, VSR_EUNSUPPORTED, "more than 8 distinct sensors on one body");
  bool uniform = true;
  for (int v = 1; v < nv && uniform; v++) {
    const int n = sensor_off[v + 1] - sensor_off[v];
    if (n != sensor_off[1] - sensor_off[0]) { uniform = false; break; }
    for (int j = 0; j < n; j++)
      if (which[sensor_off[v] + j] != which[sensor_off[0] + j] || sensors[sensor_off[v] + j].param != sensors[sensor_off[0] + j].param ||
          sensors[sensor_off[v] + j].side != sensors[sensor_off[0] + j].side) uniform = false;
  }
  ctx->w = w; ctx->h = h; ctx->n_vox = nv;
  ctx->mask.assign(mask, mask + w * h);
  ctx->vox_of_cell.assign(w * h, -1);
  ctx->cell_of_vox.clear();
  for (int i = 0; i < w * h; i++)
    if (mask[i]) { ctx->vox_of_cell[i] = (int)ctx->cell_of_vox.size(); ctx->cell_of_vox.push_back(i); }
  ctx->sensors = uniq;
  ctx->vsens.assign(sensors, sensors + total);
  ctx->vsens_off.assign(sensor_off, sensor_off + nv + 1);
  ctx->vsens_u = which;
  ctx->uniform_sensors = uniform;
  ctx->has_body = true;
  return finalize(ctx);
}

int vsr_controller_set(vsr_ctx* ctx, vsr_ctrl_kind kind, const int32_t* layers, int32_t n_layers, int32_t signals, double freq,
                       double amp, int32_t in_act) {
  if (!ctx) return VSR_EINVAL;
  if ((int)kind < 0 || (int)kind > VSR_CTRL_TIME_SIN) return fail(ctx, VSR_EINVAL, "unknown controller kind");
  if (!is_sin(kind)) {
    if (!layers || n_layers < 2 || n_layers > vsr::kMaxLayers) return fail(ctx, VSR_EINVAL, "MLP needs 2..8 layers");
    for (int i = 0; i < n_layers; i++)
      if (layers[i] <= 0 || layers[i] > 255) return fail(ctx, VSR_EUNSUPPORTED, "layer size must be in 1..255");
    if (is_dist(kind) && (signals < 0 || signals > 4)) return fail(ctx, VSR_EUNSUPPORTED, "signals must be in 0..4");
  }
  ctx->kind = kind;
  ctx->layers.clear();
  if (!is_sin(kind)) ctx->layers.assign(layers, layers + n_layers);
  ctx->signals = is_dist(kind) ? signals : 0;
  ctx->freq = freq; ctx->amp = amp; ctx->in_act = in_act ? 1 : 0;
  ctx->has_ctrl = true;
  return finalize(ctx);
}

int vsr_genome_len(const vsr_ctx* ctx) { return ctx ? genome_len_of(ctx) : VSR_EINVAL; }
int vsr_n_voxels(const vsr_ctx* ctx) { return ctx && ctx->has_body ? ctx->n_vox : VSR_ESTATE; }

int vsr_eval(vsr_ctx* ctx, const double* genomes, int32_t n, double* fitness_out, vsr_result* results_out) {
  if (!ctx) return VSR_EINVAL;
  if (!ctx->ready) return fail(ctx, VSR_ESTATE, "body and controller must be set before vsr_eval");
  if (n < 0 || (n > 0 && (!genomes || !fitness_out))) return fail(ctx, VSR_EINVAL, "bad vsr_eval arguments");
  if (n == 0) return VSR_OK;
  const int G = ctx->kp.genome_len;
  const size_t nd = ctx->devs.size(), per = ((size_t)n + nd - 1) / nd;
  // page-locked caller buffers (cudaHostAlloc / cudaHostRegister, e.g. a JNI direct buffer registered once) are copied
  // from directly; pageable ones go through the context's pinned staging buffer
  bool pinned_in = false;
  {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, genomes) == cudaSuccess) pinned_in = pa.type == cudaMemoryTypeHost;
    else cudaGetLastError();
  }
  // shard by contiguous index ranges; one stream per device; no inter-device traffic.  A shard larger than kChunk robots
  // is cut into chunks whose uploads overlap the previous chunk's episodes (copy stream = the device's second stream).
  constexpr size_t kChunk = 65536;
  for (size_t i = 0; i < nd; i++) {
    DeviceState& d = ctx->devs[i];
    size_t lo = std::min((size_t)n, i * per), hi = std::min((size_t)n, lo + per), cnt = hi - lo;
    if (!cnt) continue;
    int rc = ensure_capacity(ctx, d, cnt);
    if (rc) return rc;
    CK(cudaSetDevice(d.dev));
    for (size_t c0 = 0; c0 < cnt; c0 += kChunk) {
      const size_t cn = std::min(kChunk, cnt - c0);
      const double* src = genomes + (lo + c0) * G;
      if (!pinned_in) {
        std::memcpy(d.h_genomes + c0 * G, src, cn * G * sizeof(double));
        src = d.h_genomes + c0 * G;
      }
      CK(cudaMemcpyAsync(d.d_genomes + c0 * G, src, cn * G * sizeof(double), cudaMemcpyHostToDevice, d.copy_stream));
      CK(cudaEventRecord(d.ev_copy, d.copy_stream));
      CK(cudaStreamWaitEvent(d.stream, d.ev_copy, 0));
      rc = enqueue(ctx, d, d.d_genomes + c0 * G, cn, d.d_fitness + c0, results_out ? d.d_results + c0 : nullptr, d.stream, nullptr, nullptr,
                   nullptr, 0, i == 0 && c0 + cn == cnt);
      if (rc) return rc;
    }
    CK(cudaMemcpyAsync(d.h_fitness, d.d_fitness, cnt * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
    if (results_out) CK(cudaMemcpyAsync(d.h_results, d.d_results, cnt * sizeof(vsr_result), cudaMemcpyDeviceToHost, d.stream));
  }
  for (size_t i = 0; i < nd; i++) {
    DeviceState& d = ctx->devs[i];
    size_t lo = std::min((size_t)n, i * per), hi = std::min((size_t)n, lo + per), cnt = hi - lo;
    if (!cnt) continue;
    CK(cudaSetDevice(d.dev));
    CK(cudaStreamSynchronize(d.stream));
    std::memcpy(fitness_out + lo, d.h_fitness, cnt * sizeof(double));
    if (results_out) std::memcpy(results_out + lo, d.h_results, cnt * sizeof(vsr_result));
  }
  return VSR_OK;
}

int vsr_eval_centers(vsr_ctx* ctx, const double* genomes, int32_t n, double* fitness_out, vsr_result* results_out, double* centers_out) {
  if (!ctx) return VSR_EINVAL;
  if (!ctx->ready) return fail(ctx, VSR_ESTATE, "body and controller must be set before vsr_eval_centers");
  if (n < 0 || (n > 0 && (!genomes || !fitness_out || !centers_out))) return fail(ctx, VSR_EINVAL, "bad vsr_eval_centers arguments");
  if (n == 0) return VSR_OK;
  const size_t G = (size_t)ctx->kp.genome_len, kept = (size_t)(ctx->kp.k_last - ctx->kp.k_first + 1);
  const size_t nd = ctx->devs.size(), per = ((size_t)n + nd - 1) / nd;
  for (size_t i = 0; i < nd; i++) {
    DeviceState& d = ctx->devs[i];
    const size_t lo = std::min((size_t)n, i * per), hi = std::min((size_t)n, lo + per), cnt = hi - lo;
    if (!cnt) continue;
    int rc = ensure_capacity(ctx, d, cnt);
    if (rc) return rc;
    CK(cudaSetDevice(d.dev));
    const size_t bytes = cnt * kept * 2 * sizeof(double);
    if (bytes > d.centers_bytes) {
      cudaFree(d.d_centers); d.d_centers = nullptr; d.centers_bytes = 0;
      if (cudaMalloc(&d.d_centers, bytes) != cudaSuccess) { cudaGetLastError(); return fail(ctx, VSR_ENOMEM, "centre trajectories do not fit the device"); }
      d.centers_bytes = bytes;
    }
    std::memcpy(d.h_genomes, genomes + lo * G, cnt * G * sizeof(double));
    CK(cudaMemcpyAsync(d.d_genomes, d.h_genomes, cnt * G * sizeof(double), cudaMemcpyHostToDevice, d.stream));
    d.cur_centers = d.d_centers;
    rc = enqueue(ctx, d, d.d_genomes, cnt, d.d_fitness, results_out ? d.d_results : nullptr, d.stream, nullptr, nullptr, nullptr, 0, i == 0);
    d.cur_centers = nullptr;
    if (rc) return rc;
    CK(cudaMemcpyAsync(d.h_fitness, d.d_fitness, cnt * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
    if (results_out) CK(cudaMemcpyAsync(d.h_results, d.d_results, cnt * sizeof(vsr_result), cudaMemcpyDeviceToHost, d.stream));
  }
  for (size_t i = 0; i < nd; i++) {
    DeviceState& d = ctx->devs[i];
    const size_t lo = std::min((size_t)n, i * per), hi = std::min((size_t)n, lo + per), cnt = hi - lo;
    if (!cnt) continue;
    CK(cudaSetDevice(d.dev));
    CK(cudaStreamSynchronize(d.stream));
    std::memcpy(fitness_out + lo, d.h_fitness, cnt * sizeof(double));
    if (results_out) std::memcpy(results_out + lo, d.h_results, cnt * sizeof(vsr_result));
    CK(cudaMemcpy(centers_out + lo * kept * 2, d.d_centers, cnt * kept * 2 * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return VSR_OK;
}

int vsr_eval_terrains(vsr_ctx* ctx, const double* genomes, int32_t n, const double* terrain_x, const double* terrain_y,
                      const int32_t* terrain_off, double* fitness_out, vsr_result* results_out) {
  if (!ctx) return VSR_EINVAL;
  if (n > 0 && (!terrain_x || !terrain_y || !terrain_off)) return fail(ctx, VSR_EINVAL, "bad vsr_eval_terrains arguments");
  vsr_eval_opts o;
  std::memset(&o, 0, sizeof(o));
  o.terrain_x = terrain_x; o.terrain_y = terrain_y; o.terrain_off = terrain_off;
  return vsr_eval_ex(ctx, genomes, n, &o, fitness_out, results_out);
}

int vsr_eval_ex(vsr_ctx* ctx, const double* genomes, int32_t n, const vsr_eval_opts* opts, double* fitness_out, vsr_result* results_out) {
  if (!ctx) return VSR_EINVAL;
  if (!ctx->ready) return fail(ctx, VSR_ESTATE, "body and controller must be set before vsr_eval_ex");
  if (n < 0 || (n > 0 && (!genomes || !fitness_out))) return fail(ctx, VSR_EINVAL, "bad vsr_eval_ex arguments");
  vsr_eval_opts none;
  std::memset(&none, 0, sizeof(none));
  const vsr_eval_opts& o = opts ? *opts : none;
  if (o.posture_n < 0 || o.posture_n > 8) return fail(ctx, VSR_EUNSUPPORTED, "posture grids of side 1..8 are supported");
  for (int q = 0; q < 2; q++)
    if (o.footprint_n[q] < 0 || o.footprint_n[q] > 8) return fail(ctx, VSR_EUNSUPPORTED, "footprints of 1..8 segments are supported");
  if ((o.posture_n > 0) != (o.posture_out != nullptr)) return fail(ctx, VSR_EINVAL, "posture_n and posture_out go together");
  for (int q = 0; q < 2; q++)
    if ((o.footprint_n[q] > 0) != (o.footprints_out[q] != nullptr)) return fail(ctx, VSR_EINVAL, "footprint_n and footprints_out go together");
  if (n == 0) return VSR_OK;
  const bool obs = o.posture_n > 0 || o.footprint_n[0] > 0 || o.footprint_n[1] > 0 || o.centers_out != nullptr;
  const int G = ctx->kp.genome_len;
  const size_t K = (size_t)(ctx->kp.k_last - ctx->kp.k_first + 1), P2 = (size_t)o.posture_n * o.posture_n;
  HostExtras hx;
  {
    int rc = build_host_extras(ctx, n, o, hx);
    if (rc) return rc;
  }
  const size_t nd = ctx->devs.size(), per = ((size_t)n + nd - 1) / nd;
  for (size_t i = 0; i < nd; i++) {
    DeviceState& d = ctx->devs[i];
    size_t lo = std::min((size_t)n, i * per), hi = std::min((size_t)n, lo + per), cnt = hi - lo;
    if (!cnt) continue;
    int rc = ensure_capacity(ctx, d, cnt);
    if (rc) return rc;
    CK(cudaSetDevice(d.dev));
    LaunchExtras x;
    rc = upload_extras(ctx, d, hx, lo, hi, x);
    if (rc) return rc;
    if (obs) {
      x.obs = true; x.post_n = o.posture_n; x.fp_n[0] = o.footprint_n[0]; x.fp_n[1] = o.footprint_n[1];
      if (o.posture_n > 0) {
        if (cnt > d.post_cap) {
          cudaFree(d.d_post); d.d_post = nullptr; d.post_cap = 0;
          CK(cudaMalloc(&d.d_post, cnt * 64 * sizeof(unsigned int)));
          d.post_cap = cnt;
        }
        CK(cudaMemsetAsync(d.d_post, 0, cnt * 64 * sizeof(unsigned int), d.stream));
        x.d_post = d.d_post;
      }
      for (int q = 0; q < 2; q++)
        if (o.footprint_n[q] > 0) {
          if (cnt * K > d.fp_bytes[q]) {
            cudaFree(d.d_fp[q]); d.d_fp[q] = nullptr; d.fp_bytes[q] = 0;
            if (cudaMalloc(&d.d_fp[q], cnt * K) != cudaSuccess) { cudaGetLastError(); return fail(ctx, VSR_ENOMEM, "footprint series do not fit the device"); }
            d.fp_bytes[q] = cnt * K;
          }
          x.d_fp[q] = d.d_fp[q];
        }
      if (o.centers_out) {
        const size_t bytes = cnt * K * 3 * sizeof(double);
        if (bytes > d.cen3_bytes) {
          cudaFree(d.d_cen3); d.d_cen3 = nullptr; d.cen3_bytes = 0;
          if (cudaMalloc(&d.d_cen3, bytes) != cudaSuccess) { cudaGetLastError(); return fail(ctx, VSR_ENOMEM, "centre trajectories do not fit the device"); }
          d.cen3_bytes = bytes;
        }
        x.d_cen3 = d.d_cen3;
      }
    }
    std::memcpy(d.h_genomes, genomes + lo * G, cnt * G * sizeof(double));
    CK(cudaMemcpyAsync(d.d_genomes, d.h_genomes, cnt * G * sizeof(double), cudaMemcpyHostToDevice, d.stream));
    rc = enqueue(ctx, d, d.d_genomes, cnt, d.d_fitness, results_out ? d.d_results : nullptr, d.stream, nullptr, nullptr, nullptr, 0, i == 0, &x);
    if (rc) return rc;
    CK(cudaMemcpyAsync(d.h_fitness, d.d_fitness, cnt * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
    if (results_out) CK(cudaMemcpyAsync(d.h_results, d.d_results, cnt * sizeof(vsr_result), cudaMemcpyDeviceToHost, d.stream));
  }
  std::vector<unsigned int> cnts;
  for (size_t i = 0; i < nd; i++) {
    DeviceState& d = ctx->devs[i];
    size_t lo = std::min((size_t)n, i * per), hi = std::min((size_t)n, lo + per), cnt = hi - lo;
    if (!cnt) continue;
    CK(cudaSetDevice(d.dev));
    CK(cudaStreamSynchronize(d.stream));
    std::memcpy(fitness_out + lo, d.h_fitness, cnt * sizeof(double));
    if (results_out) std::memcpy(results_out + lo, d.h_results, cnt * sizeof(vsr_result));
    if (o.posture_n > 0) { // average posture: cells set in MORE than half of the kept observations
      cnts.resize(cnt * 64);
      CK(cudaMemcpy(cnts.data(), d.d_post, cnt * 64 * sizeof(unsigned int), cudaMemcpyDeviceToHost));
      for (size_t r = 0; r < cnt; r++)
        for (size_t c = 0; c < P2; c++) o.posture_out[(lo + r) * P2 + c] = (uint8_t)(2 * (size_t)cnts[r * 64 + c] > K);
    }
    for (int q = 0; q < 2; q++)
      if (o.footprint_n[q] > 0) CK(cudaMemcpy(o.footprints_out[q] + lo * K, d.d_fp[q], cnt * K, cudaMemcpyDeviceToHost));
    if (o.centers_out) CK(cudaMemcpy(o.centers_out + lo * K * 3, d.d_cen3, cnt * K * 3 * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return VSR_OK;
}

int vsr_stage_eval_opts(vsr_ctx* ctx, int32_t n, const vsr_eval_opts* opts) {
  if (!ctx) return VSR_EINVAL;
  if (!ctx->ready) return fail(ctx, VSR_ESTATE, "body and controller must be set before vsr_stage_eval_opts");
  if (n < 0) return fail(ctx, VSR_EINVAL, "negative batch size");
  if (n == 0 || !opts) { ctx->staged.clear(); return VSR_OK; }
  if (opts->posture_n || opts->footprint_n[0] || opts->footprint_n[1] || opts->centers_out)
    return fail(ctx, VSR_EUNSUPPORTED, "observation requests cannot be staged: use vsr_eval_ex");
  HostExtras h;
  int rc = build_host_extras(ctx, n, *opts, h);
  if (rc) return rc;
  // own the data: the caller's arrays may go away before the device-resident batch runs
  if (h.frozen) { h.o_frozen.assign(h.frozen, h.frozen + n); }
  if (h.tx) {
    const size_t nv = (size_t)h.off[n];
    h.o_tx.assign(h.tx, h.tx + nv); h.o_ty.assign(h.ty, h.ty + nv); h.o_off.assign(h.off, h.off + n + 1);
  }
  ctx->staged = std::move(h);
  HostExtras& t = ctx->staged;
  t.frozen = t.o_frozen.empty() ? nullptr : t.o_frozen.data();
  t.tx = t.o_tx.empty() ? nullptr : t.o_tx.data(); t.ty = t.o_ty.empty() ? nullptr : t.o_ty.data();
  t.off = t.o_off.empty() ? nullptr : t.o_off.data();
  if (!t.frozen && !t.tx) t.clear();
  return VSR_OK;
}

int vsr_ctx_episode(const vsr_ctx* ctx, int32_t* n_steps, int32_t* k_first, int32_t* k_last) {
  if (!ctx || !ctx->ready) return VSR_ESTATE;
  if (n_steps) *n_steps = ctx->kp.n_steps;
  if (k_first) *k_first = ctx->kp.k_first;
  if (k_last) *k_last = ctx->kp.k_last;
  return VSR_OK;
}

int vsr_central_voxel(const vsr_ctx* ctx) { return ctx && ctx->ready ? ctx->central_vox : VSR_ESTATE; }

int vsr_eval_device(vsr_ctx* ctx, const double* d_genomes, int32_t n, double* d_fitness, vsr_result* d_results, void* stream) {
  if (!ctx) return VSR_EINVAL;
  if (!ctx->ready) return fail(ctx, VSR_ESTATE, "body and controller must be set before vsr_eval_device");
  if (n < 0 || (n > 0 && (!d_genomes || !d_fitness))) return fail(ctx, VSR_EINVAL, "bad vsr_eval_device arguments");
  DeviceState& d = ctx->devs[0];
  CK(cudaSetDevice(d.dev));
  LaunchExtras x;
  if (ctx->staged.n > 0 && n > 0) {
    if (n > ctx->staged.n) return fail(ctx, VSR_EINVAL, "the staged per-robot inputs cover fewer robots than the batch evaluates");
    int rc = upload_extras(ctx, d, ctx->staged, 0, (size_t)n, x);
    if (rc) return rc;
  }
  return enqueue(ctx, d, d_genomes, (size_t)n, d_fitness, d_results, (cudaStream_t)stream, nullptr, nullptr, nullptr, 0, true,
                 ctx->staged.n > 0 ? &x : nullptr);
}

int vsr_trace(vsr_ctx* ctx, const double* genome, int32_t n_steps, double* vertices_out, double* bodies_out, double* ctrl_out,
              vsr_result* result_out) {
  if (!ctx) return VSR_EINVAL;
  if (!ctx->ready) return fail(ctx, VSR_ESTATE, "body and controller must be set before vsr_trace");
  if (!genome || n_steps < 0) return fail(ctx, VSR_EINVAL, "bad vsr_trace arguments");
  DeviceState& d = ctx->devs[0];
  int rc = ensure_capacity(ctx, d, 1);
  if (rc) return rc;
  const int G = ctx->kp.genome_len, nv = ctx->n_vox;
  n_steps = std::min(n_steps, ctx->kp.n_steps);
  double *dv = nullptr, *db = nullptr, *dc = nullptr;
  const size_t bv = (size_t)n_steps * nv * 8 * sizeof(double), bb = (size_t)n_steps * nv * 24 * sizeof(double), bc = (size_t)n_steps * nv * sizeof(double);
  if (vertices_out && n_steps) CK(cudaMalloc(&dv, bv));
  if (bodies_out && n_steps) CK(cudaMalloc(&db, bb));
  if (ctrl_out && n_steps) CK(cudaMalloc(&dc, bc));
  CK(cudaMemcpyAsync(d.d_genomes, genome, G * sizeof(double), cudaMemcpyHostToDevice, d.stream));
  rc = enqueue(ctx, d, d.d_genomes, 1, d.d_fitness, d.d_results, d.stream, dv, db, dc, n_steps, false);
  if (rc == VSR_OK) {
    cudaError_t e_ = cudaStreamSynchronize(d.stream);
    if (e_ != cudaSuccess) rc = fail(ctx, VSR_ECUDA, std::string("trace kernel: ") + cudaGetErrorString(e_));
  }
  if (rc == VSR_OK) {
    if (dv) cudaMemcpy(vertices_out, dv, bv, cudaMemcpyDeviceToHost);
    if (db) cudaMemcpy(bodies_out, db, bb, cudaMemcpyDeviceToHost);
    if (dc) cudaMemcpy(ctrl_out, dc, bc, cudaMemcpyDeviceToHost);
    if (result_out) cudaMemcpy(result_out, d.d_results, sizeof(vsr_result), cudaMemcpyDeviceToHost);
  }
  cudaFree(dv); cudaFree(db); cudaFree(dc);
  return rc;
}

double vsr_last_kernel_ms(vsr_ctx* ctx) {
  if (!ctx || ctx->devs.empty() || !ctx->devs[0].timed) return -1.0;
  DeviceState& d = ctx->devs[0];
  cudaSetDevice(d.dev);
  if (cudaEventSynchronize(d.ev1) != cudaSuccess) return -1.0;
  float ms = 0;
  if (cudaEventElapsedTime(&ms, d.ev0, d.ev1) != cudaSuccess) return -1.0;
  return ms;
}

int64_t vsr_kernel_launches(const vsr_ctx* ctx) { return ctx ? ctx->launches : 0; }

int64_t vsr_contact_overflows(vsr_ctx* ctx) {
  if (!ctx) return VSR_EINVAL;
  int64_t total = 0;
  for (auto& d : ctx->devs) {
    int v = 0;
    CK(cudaSetDevice(d.dev));
    CK(cudaStreamSynchronize(d.stream));
    CK(cudaMemcpy(&v, d.d_ovf, sizeof(int), cudaMemcpyDeviceToHost));
    total += v;
  }
  return total;
}

int vsr_kernel_info(const vsr_ctx* ctx_c, int32_t* regs, int32_t* smem, int32_t* threads, int32_t* rpw, int32_t* ctas_per_sm) {
  vsr_ctx* ctx = const_cast<vsr_ctx*>(ctx_c);
  if (!ctx) return VSR_EINVAL;
  if (!ctx->ready) return fail(ctx, VSR_ESTATE, "body and controller must be set first");
  CK(cudaSetDevice(ctx->devs[0].dev));
  // the shape of the most recent launch (sub-wave batches use another build and warp count), else the full-wave default
  const void* kern = ctx->last_kernel ? ctx->last_kernel : ctx->kernel;
  const int warps = ctx->last_kernel ? ctx->last_warps : ctx->warps;
  const size_t sm = ctx->last_kernel ? ctx->last_smem : ctx->smem_bytes;
  cudaFuncAttributes a;
  CK(cudaFuncGetAttributes(&a, kern));
  int nb = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, warps * 32, sm));
  if (regs) *regs = a.numRegs;
  if (smem) *smem = (int)sm;
  if (threads) *threads = warps * 32;
  if (rpw) *rpw = ctx->kp.robots_per_warp;
  if (ctas_per_sm) *ctas_per_sm = nb;
  return VSR_OK;
}

int vsr_input_map(const vsr_ctx* ctx, int32_t* cell, int32_t* sensor, int32_t* reading, int32_t cap) {
  if (!ctx || !ctx->has_body) return VSR_ESTATE;
  int n = 0;
  // CentralizedSensing: voxels in Grid.values() order (row-major y*w+x, nulls skipped), sensors in list order
  // (`sensor` = position in the voxel's own list)
  for (int v = 0; v < ctx->n_vox; v++)
    for (int s = ctx->vsens_off[v]; s < ctx->vsens_off[v + 1]; s++)
      for (int r = 0; r < sensor_dim(ctx->vsens[s]); r++) {
        if (n < cap) {
          if (cell) cell[n] = ctx->cell_of_vox[v];
          if (sensor) sensor[n] = s - ctx->vsens_off[v];
          if (reading) reading[n] = r;
        }
        n++;
      }
  return n;
}

int vsr_voxel_cells(const vsr_ctx* ctx, int32_t* cell, int32_t cap) {
  if (!ctx || !ctx->has_body) return VSR_ESTATE;
  for (int v = 0; v < ctx->n_vox && v < cap; v++)
    if (cell) cell[v] = ctx->cell_of_vox[v];
  return ctx->n_vox;
}

int vsr_counts(const vsr_ctx* ctx, int32_t* n_voxels, int32_t* n_springs, int32_t* n_ropes, int32_t* n_welds) {
  if (!ctx || !ctx->has_body) return VSR_ESTATE;
  int per = 0, sm = ctx->env.scaffolding_mask;
  if (sm & 1) per += 4;
  if (sm & 2) per += 4;
  if (sm & 4) per += 8;
  if (sm & 8) per += 2;
  int nw = 0;
  for (int v = 0; v < ctx->n_vox; v++) {
    int cell = ctx->cell_of_vox[v], x = cell % ctx->w, y = cell / ctx->w;
    if (x > 0 && ctx->mask[cell - 1]) nw += 2;
    if (y > 0 && ctx->mask[cell - ctx->w]) nw += 2;
  }
  if (n_voxels) *n_voxels = ctx->n_vox;
  if (n_springs) *n_springs = ctx->n_vox * per;
  if (n_ropes) *n_ropes = ctx->env.limit_contraction ? ctx->n_vox * 4 : 0;
  if (n_welds) *n_welds = nw;
  return VSR_OK;
}

} // extern "C"
