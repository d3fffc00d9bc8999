// vsr_names.cpp -- integer / naming layer of libvsr_eval.so (host only, no CUDA).
//
// Restates the name -> structure functions that surround the hot path so that the Java host (or the
// Python mirror) can hand plain arrays to the evaluator:
//   vsr_shape_mask          2dhmsr RobotUtils.buildShape(String)        (call site locomotion/Starter.java:405)
//   vsr_terrain             2dhmsr Locomotion.createTerrain(String)     (call site locomotion/Starter.java:87,91)
//   vsr_sensor_config       2dhmsr RobotUtils.buildSensorizingFunction  (call site locomotion/Starter.java:404)
//   vsr_mlp_inner_neurons   builder/function/MLP.java:28-42
//   vsr_mlp_count_weights   2dhmsr MultiLayerPerceptron.countWeights    (used at builder/function/MLP.java:55,76)
//   vsr_episode_steps       2dhmsr Locomotion.apply loop + Outcome.subOutcome (locomotion/Starter.java:430-433)
//   JavaRandom              java.util.Random (terrain seeds, Starter.java:91 "-rnd" replacement)
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vsr_eval.h"

namespace {

struct JavaRandom {
  uint64_t seed;
  bool have = false;
  double next_g = 0;
  explicit JavaRandom(int64_t s) : seed(((uint64_t)s ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1)) {}
  int32_t next(int bits) {
    seed = (seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
    return (int32_t)((int64_t)seed >> (48 - bits));
  }
  int32_t next_int(int32_t bound) { // java.util.Random.nextInt(int bound)
    int32_t r = next(31);
    const int32_t m = bound - 1;
    if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
    for (int32_t u = r;; u = next(31)) { // Java: `u - (r = u % bound) + m < 0` detects the int overflow
      r = u % bound;
      if ((int64_t)u - r + m < (1LL << 31)) break;
    }
    return r;
  }
  double next_double() { return (double)(((int64_t)next(26) << 27) + next(27)) * (1.0 / (double)(1LL << 53)); }
  double next_gaussian() {
    if (have) { have = false; return next_g; }
    double v1, v2, s;
    do {
      v1 = 2 * next_double() - 1;
      v2 = 2 * next_double() - 1;
      s = v1 * v1 + v2 * v2;
    } while (s >= 1 || s == 0);
    double mul = std::sqrt(-2 * std::log(s) / s);
    next_g = v2 * mul;
    have = true;
    return v1 * mul;
  }
};

std::vector<std::string> split(const std::string& s, char c) {
  std::vector<std::string> out;
  size_t p = 0;
  for (;;) {
    size_t q = s.find(c, p);
    out.push_back(s.substr(p, q == std::string::npos ? q : q - p));
    if (q == std::string::npos) break;
    p = q + 1;
  }
  return out;
}

bool parse_int(const std::string& s, long* v) {
  char* e = nullptr;
  *v = std::strtol(s.c_str(), &e, 10);
  return e && *e == 0 && !s.empty();
}
bool parse_double(const std::string& s, double* v) {
  char* e = nullptr;
  *v = std::strtod(s.c_str(), &e);
  return e && *e == 0 && !s.empty();
}

constexpr double kBorderH = 100.0, kLength = 2000.0, kBorderW = 10.0;

} // namespace

extern "C" {

int vsr_java_random_gaussians(int64_t seed, double* out, int32_t n) {
  if (!out || n < 0) return VSR_EINVAL;
  JavaRandom r(seed);
  for (int i = 0; i < n; i++) out[i] = r.next_gaussian();
  return VSR_OK;
}

int vsr_java_random_ints(int64_t seed, int32_t bound, int32_t* out, int32_t n) {
  if (!out || n < 0 || bound <= 0) return VSR_EINVAL;
  JavaRandom r(seed);
  for (int i = 0; i < n; i++) out[i] = r.next_int(bound);
  return VSR_OK;
}

// RobotUtils.buildRobotTransformation(name, externalRandom) [U] (call sites Starter.java:126,428; CLI default `identity`, :153):
//   identity | broken-<ratio>-<seed|rnd> | breakable-<time|area>-<mean>/<std>-<mean>/<std>-<seed|rnd>
// kind: 0 identity, 1 broken; breakable (malfunctions triggered during the episode) is outside this build
int vsr_robot_transformation(const char* name, int32_t* kind, double* ratio, int64_t* seed, int32_t* seed_is_rnd) {
  if (!name || !kind) return VSR_EINVAL;
  const std::string s(name);
  if (s == "identity") { *kind = 0; return VSR_OK; }
  auto p = split(s, '-');
  if (p[0] == "broken") {
    double r;
    long sd = 0;
    if (p.size() != 3 || !parse_double(p[1], &r) || r < 0) return VSR_EINVAL;
    const bool rnd = p[2] == "rnd";
    if (!rnd && (!parse_int(p[2], &sd) || sd < 0)) return VSR_EINVAL;
    *kind = 1;
    if (ratio) *ratio = r;
    if (seed) *seed = sd;
    if (seed_is_rnd) *seed_is_rnd = rnd ? 1 : 0;
    return VSR_OK;
  }
  if (p[0] == "breakable") return VSR_EUNSUPPORTED;
  return VSR_EINVAL;
}

int vsr_broken_masks(const vsr_ctx* ctx, double ratio, int64_t seed, int32_t n, uint32_t* masks_out) {
  uint64_t state = ((uint64_t)seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1);
  return vsr_broken_masks_stream(ctx, ratio, &state, n, masks_out);
}

int vsr_broken_masks_stream(const vsr_ctx* ctx, double ratio, uint64_t* lcg_state, int32_t n, uint32_t* masks_out) {
  const int nv = vsr_n_voxels(ctx);
  if (nv < 0) return nv;
  if (!lcg_state || n < 0 || (n > 0 && !masks_out) || !(ratio >= 0)) return VSR_EINVAL;
  JavaRandom r(0);
  r.seed = *lcg_state & ((1ULL << 48) - 1);
  struct Save { JavaRandom& r; uint64_t* out; ~Save() { *out = r.seed; } } save{r, lcg_state};
  for (int i = 0; i < n; i++) {
    uint32_t m = 0;
    for (int v = 0; v < nv; v++) // Grid.create(body, v -> v == null ? null : random.nextDouble() > ratio ? v : new BreakableVoxel(..., random.nextInt(), ...))
      if (!(r.next_double() > ratio)) { m |= 1u << v; (void)r.next(32); }
    masks_out[i] = m;
  }
  return VSR_OK;
}

int vsr_shape_mask(const char* name, int32_t* w, int32_t* h, uint8_t* mask, int32_t cap) {
  if (!name || !w || !h) return VSR_EINVAL;
  auto p = split(name, '-');
  if (p.size() != 2) return VSR_EINVAL;
  auto d = split(p[1], 'x');
  long W, H;
  if (d.size() != 2 || !parse_int(d[0], &W) || !parse_int(d[1], &H) || W <= 0 || H <= 0) return VSR_EINVAL;
  const std::string& k = p[0];
  if (k != "box" && k != "worm" && k != "biped" && k != "tripod" && k != "comb") return VSR_EINVAL;
  *w = (int32_t)W;
  *h = (int32_t)H;
  if (!mask) return VSR_OK;
  if (cap < W * H) return VSR_EINVAL;
  for (long y = 0; y < H; y++)
    for (long x = 0; x < W; x++) {
      bool v = true;
      if (k == "biped") v = !(y < H / 2 && x >= W / 4 && x < W * 3 / 4);
      else if (k == "tripod") v = !(y < H / 2 && x != 0 && x != W - 1 && x != W / 2);
      else if (k == "comb") v = y >= H / 2 || x % 2 == 0;
      mask[y * W + x] = v ? 1 : 0;
    }
  return VSR_OK;
}

int vsr_terrain(const char* name, double* xs_out, double* ys_out, int32_t cap) {
  if (!name) return VSR_EINVAL;
  auto p = split(name, '-');
  std::vector<double> xs, ys;
  if (p.size() == 1 && p[0] == "flat") {
    xs = {0.0, kBorderW, kLength - kBorderW, kLength};
    ys = {kBorderH, 5.0, 5.0, kBorderH};
  } else if (p.size() == 4 && (p[0] == "hilly" || p[0] == "steppy")) {
    double h, w;
    long seed;
    if (!parse_double(p[1], &h) || !parse_double(p[2], &w) || !parse_int(p[3], &seed)) return VSR_EINVAL;
    JavaRandom rnd(seed);
    xs = {0.0, kBorderW};
    ys = {kBorderH, 0.0};
    while (xs.back() < kLength - kBorderW) {
      xs.push_back(xs.back() + std::fmax(1.0, (rnd.next_gaussian() * 0.25 + 1) * w));
      if (p[0] == "hilly") {
        ys.push_back(ys.back() + rnd.next_gaussian() * h);
      } else {
        xs.push_back(xs.back() + 0.5);
        ys.push_back(ys.back());
        ys.push_back(ys.back() + rnd.next_gaussian() * h);
      }
    }
    xs.push_back(xs.back() + kBorderW);
    ys.push_back(kBorderH);
  } else if (p.size() == 2 && (p[0] == "downhill" || p[0] == "uphill")) {
    double a;
    if (!parse_double(p[1], &a)) return VSR_EINVAL;
    double dy = (kLength - 2 * kBorderW) * std::sin(a / 180 * M_PI);
    xs = {0.0, kBorderW, kLength - kBorderW, kLength};
    if (p[0] == "downhill") ys = {kBorderH + dy, 5 + dy, 5.0, kBorderH};
    else ys = {kBorderH, 5.0, 5 + dy, kBorderH + dy};
  } else {
    return VSR_EINVAL;
  }
  int n = (int)xs.size();
  if (xs_out && ys_out) {
    if (cap < n) return VSR_EINVAL;
    std::memcpy(xs_out, xs.data(), sizeof(double) * n);
    std::memcpy(ys_out, ys.data(), sizeof(double) * n);
  }
  return n;
}

namespace {
// Lidar.Side of a voxel at relative grid position (x, y) in [0, 1]^2 [U] (RobotUtils.lidarSide): the side that faces outwards
int lidar_side(double x, double y) {
  if (y > x && y > 1 - x) return 2;   // S
  if (y >= x && y <= 1 - x) return 3; // W
  if (y < x && y < 1 - x) return 0;   // N
  return 1;                           // E
}

// one entry of RobotUtils' PREDEFINED_SENSORS [U] for the voxel at (x, y) of a w x h grid; false = not built here
bool predefined_sensor(const std::string& s, int x, int y, int w, int h, double sigma, vsr_sensor_desc* d) {
  std::memset(d, 0, sizeof(*d));
  d->dom_min = 0; d->dom_max = 1; d->noise_sigma = sigma;
  const double rx = w > 1 ? (double)x / ((double)w - 1.0) : 0.0 / 0.0, ry = h > 1 ? (double)y / ((double)h - 1.0) : 0.0 / 0.0; // Java: 0d / 0d = NaN
  if (s == "a") { d->kind = VSR_SENSOR_AREA_RATIO; d->soft_norm = 1; d->dom_min = 0.5; d->dom_max = 1.5; }
  else if (s == "t") { d->kind = VSR_SENSOR_TOUCH; d->avg_window = 0.25; }
  else if (s == "r") { d->kind = VSR_SENSOR_ANGLE; d->soft_norm = 1; d->dom_min = -M_PI; d->dom_max = M_PI; }
  else if (s == "vx" || s == "vy" || s == "vxy") {
    d->kind = VSR_SENSOR_VELOCITY;
    d->axes = s == "vx" ? 1 : (s == "vy" ? 2 : 3);
    d->rotated = 1; d->soft_norm = 1; d->dom_min = -8; d->dom_max = 8; d->avg_window = 0.5;
  }
  else if (s == "px") { d->kind = VSR_SENSOR_CONSTANT; d->param = rx; }
  else if (s == "py") { d->kind = VSR_SENSOR_CONSTANT; d->param = ry; }
  else if (s == "cpg") { d->kind = VSR_SENSOR_TIME_SIN; d->param = -1.0; d->soft_norm = 2; d->dom_min = -1; d->dom_max = 1; }
  else if (s == "l1" || s == "l5" || s == "lf5") {
    d->kind = VSR_SENSOR_LIDAR; d->param = 10.0; d->soft_norm = 2; d->dom_min = 0; d->dom_max = 10.0;
    d->n_rays = s == "l1" ? 1 : 5;
    d->side = s == "lf5" ? 1 : lidar_side(rx, ry);
  }
  else return false;
  return true;
}
} // namespace

int vsr_sensor_config(const char* name, vsr_sensor_desc* out, int32_t cap) {
  if (!name) return VSR_EINVAL;
  auto p = split(name, '-');
  double sigma;
  if (p.size() != 3 || p[0] != "uniform" || !parse_double(p[2], &sigma)) return VSR_EINVAL;
  auto names = split(p[1], '+');
  int n = 0;
  for (auto& s : names) {
    vsr_sensor_desc d;
    // the position-dependent entries (px, py, l1, l5) differ between voxels: they need vsr_sensor_config_voxels
    if (s == "px" || s == "py" || s == "l1" || s == "l5" || !predefined_sensor(s, 0, 0, 2, 2, sigma, &d)) return VSR_EUNSUPPORTED;
    if (out) {
      if (n >= cap) return VSR_EINVAL;
      out[n] = d;
    }
    n++;
  }
  return n;
}

int vsr_sensor_config_voxels(const char* name, int32_t w, int32_t h, const uint8_t* mask, vsr_sensor_desc* out, int32_t cap, int32_t* off_out) {
  if (!name || !mask || w <= 0 || h <= 0) return VSR_EINVAL;
  const std::string nm(name);
  auto p = split(nm, '-');
  enum { UNIFORM, SPINED, SPINED_SIGHTED, EMPTY } form;
  double sigma = 0;
  bool cpg = false;
  std::vector<std::string> names;
  if (nm == "empty") form = EMPTY;
  else if (p[0] == "uniform" && p.size() == 3 && parse_double(p[2], &sigma)) { form = UNIFORM; names = split(p[1], '+'); }
  else if ((p[0] == "spinedTouch" || p[0] == "spinedTouchSighted") && p.size() == 4 && (p[1] == "t" || p[1] == "f") &&
           (p[2] == "t" || p[2] == "f") && parse_double(p[3], &sigma)) {
    form = p[0] == "spinedTouch" ? SPINED : SPINED_SIGHTED;
    cpg = p[1] == "t";
    if (p[2] == "t") return VSR_EUNSUPPORTED; // the Malfunction sensor
  } else return VSR_EINVAL;
  if (sigma < 0) return VSR_EINVAL;
  int n = 0, v = 0;
  auto emit = [&](const vsr_sensor_desc& d) -> bool {
    if (out) {
      if (n >= cap) return false;
      out[n] = d;
    }
    n++;
    return true;
  };
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      if (!mask[y * w + x]) continue;
      if (off_out) off_out[v] = n;
      v++;
      vsr_sensor_desc d;
      if (form == UNIFORM) {
        for (auto& s : names) {
          if (!predefined_sensor(s, x, y, w, h, sigma, &d)) return VSR_EUNSUPPORTED;
          if (!emit(d)) return VSR_EINVAL;
        }
      } else if (form == SPINED || form == SPINED_SIGHTED) {
        // [U] spinedTouch: touch on the bottom row, (hard-normalised, un-averaged) velocity on the top row, area ratio everywhere,
        // the central pattern generator on the top-right voxel; Sighted: a 5-ray front lidar on the right column
        if (y == 0) { predefined_sensor("t", x, y, w, h, sigma, &d); if (!emit(d)) return VSR_EINVAL; }
        if (y == h - 1) {
          predefined_sensor("vxy", x, y, w, h, sigma, &d);
          d.soft_norm = 2; d.avg_window = 0;
          if (!emit(d)) return VSR_EINVAL;
        }
        predefined_sensor("a", x, y, w, h, sigma, &d);
        d.soft_norm = 2;
        if (!emit(d)) return VSR_EINVAL;
        if (cpg && x == w - 1 && y == h - 1) { predefined_sensor("cpg", x, y, w, h, sigma, &d); if (!emit(d)) return VSR_EINVAL; }
        if (form == SPINED_SIGHTED && x == w - 1) { predefined_sensor("lf5", x, y, w, h, sigma, &d); if (!emit(d)) return VSR_EINVAL; }
      }
    }
  if (off_out) off_out[v] = n;
  return n;
}

int vsr_mlp_inner_neurons(int32_t n_in, int32_t n_out, double ratio, int32_t n_inner, int32_t* out) {
  if (n_inner < 0 || (n_inner > 0 && !out)) return VSR_EINVAL;
  // Java: (int) Math.max(2, Math.round(nIn * r)); Math.round = floor(x + 0.5); int '/' truncates toward zero (as C++)
  int center = (int)std::fmax(2.0, std::floor(n_in * ratio + 0.5));
  if (n_inner > 1) {
    for (int i = 0; i < n_inner / 2; i++) out[i] = n_in + (center - n_in) / (n_inner / 2 + 1) * (i + 1);
    for (int i = n_inner / 2; i < n_inner; i++) out[i] = center + (n_out - center) / (n_inner / 2 + 1) * (i - n_inner / 2);
  } else if (n_inner > 0) {
    out[0] = center;
  }
  return n_inner;
}

int vsr_mlp_count_weights(const int32_t* neurons, int32_t n_layers) {
  if (!neurons || n_layers < 2) return VSR_EINVAL;
  int n = 0;
  for (int i = 0; i + 1 < n_layers; i++) n += neurons[i + 1] * (neurons[i] + 1);
  return n;
}

int vsr_mlp_weight_index(const int32_t* neurons, int32_t n_layers, int32_t l, int32_t j, int32_t k) {
  if (!neurons || l < 0 || l + 1 >= n_layers || j < 0 || j >= neurons[l + 1] || k < 0 || k > neurons[l]) return VSR_EINVAL;
  int o = 0;
  for (int i = 0; i < l; i++) o += neurons[i + 1] * (neurons[i] + 1);
  return o + j * (neurons[l] + 1) + k;
}

int vsr_episode_steps(double final_t, double transient_t, double dt, int32_t* n_steps, int32_t* k_first, int32_t* k_last) {
  if (!(dt > 0) || !(final_t > 0) || final_t / dt > 1e8) return VSR_EINVAL;
  volatile double t = 0.0; // volatile: forbid extended-precision accumulation
  int k = 0, kf = 0, kl = 0;
  while (t < final_t) {
    t = t + dt;
    k++;
    if (t >= transient_t && t < final_t) {
      if (!kf) kf = k;
      kl = k;
    }
  }
  if (n_steps) *n_steps = k;
  if (k_first) *k_first = kf;
  if (k_last) *k_last = kl;
  return VSR_OK;
}

} // extern "C"
