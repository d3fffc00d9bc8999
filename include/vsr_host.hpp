// vsr_host.hpp -- C++ host side above the C ABI of libvsr_eval.so (include/vsr_eval.h): the reference's plugin interface for the
// hot path, restated with the reference's own names, argument meaning and error behaviour, so that code written against
// VSREvolution's Java reads the same here.  Header only, C++17, no dependency but the C ABI.
//
//   reference (Java, /root/reference/src/main/java/it/units/erallab/...)            here
//   builder/PrototypedFunctionBuilder.java:12-29  buildFor / exampleFor / compose     PrototypedFunctionBuilder<A, B>
//   builder/NamedProvider.java:44-60              "name;k=v;k=v" parsing              parseParams
//   builder/function/MLP.java:28-89               innerNeurons, weights -> MLP        MLP
//   builder/robot/BrainCentralized.java:25-57     function -> Robot(CentralizedSensing) BrainCentralized
//   builder/robot/BrainPhaseValues.java:26-60     values -> Robot(PhaseSin)           BrainPhaseValues
//   builder/robot/BrainHomoDistributed.java:24-96 function -> Robot(DistributedSensing) BrainHomoDistributed
//   locomotion/Starter.java:192-247 getMapper     "robotMapper<functionMapper"        buildMapper
//   locomotion/Starter.java:83-94,427-435         Locomotion + subOutcome             Locomotion (apply / applyAll), Outcome
//   2dhmsr Grid / Voxel / Robot / controllers (un-vendored dependency)                the carrier types below
//
// Robots are carriers: nothing is simulated on the host.  Locomotion flattens a population into the C ABI's (body, sensors, controller
// structure, genome matrix) form, one vsr_ctx per distinct structure, and evaluates it with one launch per structure.  Errors of the
// mappers are IllegalArgumentException (the Java type and message); a failed library call is VsrException with the C status.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <atomic>
#include <exception>
#include <memory>
#include <mutex>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "vsr_eval.h"

namespace vsr_host {

struct IllegalArgumentException : std::invalid_argument {
  using std::invalid_argument::invalid_argument;
};
struct VsrException : std::runtime_error {
  int status;
  VsrException(int status_, const std::string& what) : std::runtime_error(what), status(status_) {}
};
namespace detail {
template <class... Args>
std::string format(const char* fmt, Args... args) {
  char buf[512];
  std::snprintf(buf, sizeof buf, fmt, args...);
  return buf;
}
inline int check(int rc, const vsr_ctx* ctx, const char* what) {
  if (rc < 0) {
    const char* msg = vsr_last_error(ctx);
    const std::string text = std::string(what) + ": " + (msg ? msg : "") + " (status " + std::to_string(rc) + ")";
    if (rc == VSR_EINVAL || rc == VSR_EDIM) throw IllegalArgumentException(text);
    throw VsrException(rc, text);
  }
  return rc;
}
} // namespace detail

// 2dhmsr util/Grid: w x h cells, absent cells are empty; iteration order is row-major (y outer, x inner), y = 0 the bottom row --
// the order BrainPhaseValues consumes its values in and the kernel numbers the voxels in.
template <class T>
class Grid {
 public:
  struct Entry { int x, y; const std::optional<T>& value; };
  Grid(int w, int h) : w_(w), h_(h), cells_((size_t)w * h) {}
  int getW() const { return w_; }
  int getH() const { return h_; }
  const std::optional<T>& get(int x, int y) const {
    static const std::optional<T> none;
    return (x < 0 || x >= w_ || y < 0 || y >= h_) ? none : cells_[(size_t)y * w_ + x];
  }
  void set(int x, int y, T v) { cells_[(size_t)y * w_ + x] = std::move(v); }
  size_t count() const { size_t n = 0; for (const auto& c : cells_) n += c.has_value(); return n; }
  template <class F> void forEach(F f) const { for (int i = 0; i < w_ * h_; i++) f(Entry{i % w_, i / w_, cells_[i]}); }
  std::vector<T> values() const { std::vector<T> out; for (const auto& c : cells_) if (c) out.push_back(*c); return out; } // non-null only
 private:
  int w_, h_;
  std::vector<std::optional<T>> cells_;
};

inline int sensorDim(const vsr_sensor_desc& s) {
  if (s.kind == VSR_SENSOR_LIDAR) return s.n_rays;
  if (s.kind == VSR_SENSOR_VELOCITY) return ((s.axes & 1) ? 1 : 0) + ((s.axes & 2) ? 1 : 0);
  return 1;
}
// 2dhmsr Voxel: here the list of its sensors (RobotUtils.buildSensorizingFunction)
struct Voxel {
  std::vector<vsr_sensor_desc> sensors;
  // 2dhmsr BreakableVoxel as `broken-<ratio>-<seed>` builds it [U]: the actuator is frozen from t = 0 (it applies no force, sensors work)
  bool actuatorFrozen = false;
  int sensorDim() const { int d = 0; for (const auto& s : sensors) d += vsr_host::sensorDim(s); return d; }
};

// java.util.Random (the 48-bit LCG of the Java class library): the run's shared Random of locomotion/Starter.java:89-93,126,428
class JavaRandom {
 public:
  explicit JavaRandom(int64_t seed) : state_(((uint64_t)seed ^ 0x5DEECE66DULL) & kMask) {}
  double nextDouble() { return (double)(((int64_t)next(26) << 27) + next(27)) * (1.0 / (double)(1LL << 53)); }
  int32_t nextInt() { return next(32); }
  int32_t nextInt(int32_t bound) {
    if (bound <= 0) throw IllegalArgumentException("bound must be positive");
    int32_t r = next(31);
    const int32_t m = bound - 1;
    if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
    for (int32_t u = r;; u = next(31)) { // Java: `u - (r = u % bound) + m < 0` on wrapping ints = the sum exceeds 2^31 - 1
      r = u % bound;
      if ((int64_t)u - r + m < (1LL << 31)) return r;
    }
  }
 private:
  static constexpr uint64_t kMask = (1ULL << 48) - 1;
  int32_t next(int bits) {
    state_ = (state_ * 0x5DEECE66DULL + 0xBULL) & kMask;
    return (int32_t)((int64_t)state_ >> (48 - bits)); // bits == 32: the truncation to int is Java's
  }
  uint64_t state_;
};

// ---- functions (2dhmsr core/controllers) -------------------------------------------------------------------------------------------
struct TimedRealFunction {
  virtual ~TimedRealFunction() = default;
  virtual int getInputDimension() const = 0;
  virtual int getOutputDimension() const = 0;
};
using FunctionPtr = std::shared_ptr<const TimedRealFunction>;
// RealFunction.build(d -> d, nIn, nOut): the dimension carrier the mappers use as an example
class RealFunction : public TimedRealFunction {
 public:
  static FunctionPtr build(int nIn, int nOut) { return std::make_shared<RealFunction>(nIn, nOut); }
  RealFunction(int nIn, int nOut) : nIn_(nIn), nOut_(nOut) {}
  int getInputDimension() const override { return nIn_; }
  int getOutputDimension() const override { return nOut_; }
 private:
  int nIn_, nOut_;
};
// MultiLayerPerceptron(TANH, nIn, inner, nOut, weights): weights flat by layer -> to-neuron -> (bias, from-neurons)
class MultiLayerPerceptron : public TimedRealFunction {
 public:
  static std::vector<int> countNeurons(int nIn, const std::vector<int>& inner, int nOut) {
    std::vector<int> n{nIn};
    n.insert(n.end(), inner.begin(), inner.end());
    n.push_back(nOut);
    return n;
  }
  static int countWeights(const std::vector<int>& neurons) {
    const int c = vsr_mlp_count_weights(neurons.data(), (int)neurons.size());
    if (c < 0) throw IllegalArgumentException("MultiLayerPerceptron needs at least an input and an output layer");
    return c;
  }
  static int countWeights(int nIn, const std::vector<int>& inner, int nOut) { return countWeights(countNeurons(nIn, inner, nOut)); }
  MultiLayerPerceptron(int nIn, const std::vector<int>& inner, int nOut, std::vector<double> weights)
      : neurons_(countNeurons(nIn, inner, nOut)), weights_(std::move(weights)) {
    if ((int)weights_.size() != countWeights(neurons_))
      throw IllegalArgumentException(detail::format("Wrong number of weights: %d expected, %d found", countWeights(neurons_), (int)weights_.size()));
  }
  int getInputDimension() const override { return neurons_.front(); }
  int getOutputDimension() const override { return neurons_.back(); }
  const std::vector<int>& getNeurons() const { return neurons_; }
  const std::vector<double>& getParams() const { return weights_; }
 private:
  std::vector<int> neurons_;
  std::vector<double> weights_;
};

// ---- controllers -----------------------------------------------------------------------------------------------------------------------
struct Controller { virtual ~Controller() = default; };
using ControllerPtr = std::shared_ptr<const Controller>;
struct PhaseSin : Controller {
  double frequency, amplitude;
  Grid<double> phases;
  PhaseSin(double frequency_, double amplitude_, Grid<double> phases_) : frequency(frequency_), amplitude(amplitude_), phases(std::move(phases_)) {}
};
struct CentralizedSensing : Controller {
  FunctionPtr function;
  explicit CentralizedSensing(const Grid<Voxel>&, FunctionPtr f) : function(std::move(f)) {}
  static int nOfInputs(const Grid<Voxel>& body) { int n = 0; for (const auto& v : body.values()) n += v.sensorDim(); return n; }
  static int nOfOutputs(const Grid<Voxel>& body) { return (int)body.count(); }
};
class DistributedSensing : public Controller {
 public:
  DistributedSensing(const Grid<Voxel>& body, int signals_) : signals(signals_), functions_(body.getW(), body.getH()) {}
  static int nOfInputs(const Voxel& v, int signals) { return v.sensorDim() + 4 * signals; }
  static int nOfOutputs(const Voxel&, int signals) { return 1 + 4 * signals; }
  Grid<FunctionPtr>& getFunctions() { return functions_; }
  const Grid<FunctionPtr>& getFunctions() const { return functions_; }
  int signals;
 private:
  Grid<FunctionPtr> functions_;
};

class Robot {
 public:
  Robot(ControllerPtr controller, Grid<Voxel> voxels) : controller_(std::move(controller)), voxels_(std::move(voxels)) {}
  const Grid<Voxel>& getVoxels() const { return voxels_; }
  const ControllerPtr& getController() const { return controller_; }
 private:
  ControllerPtr controller_;
  Grid<Voxel> voxels_;
};

// RobotUtils.buildShape / buildSensorizingFunction (locomotion/Starter.java:404-405): the library's name layer does the parsing
struct RobotUtils {
  static Grid<bool> buildShape(const std::string& name) {
    int32_t w = 0, h = 0;
    detail::check(vsr_shape_mask(name.c_str(), &w, &h, nullptr, 0), nullptr, "RobotUtils.buildShape");
    std::vector<uint8_t> mask((size_t)w * h);
    detail::check(vsr_shape_mask(name.c_str(), &w, &h, mask.data(), w * h), nullptr, "RobotUtils.buildShape");
    Grid<bool> g(w, h);
    for (int i = 0; i < w * h; i++) if (mask[i]) g.set(i % w, i / w, true);
    return g;
  }
  static std::function<Grid<Voxel>(const Grid<bool>&)> buildSensorizingFunction(const std::string& name) {
    return [name](const Grid<bool>& shape) {
      const int w = shape.getW(), h = shape.getH();
      std::vector<uint8_t> mask((size_t)w * h, 0);
      shape.forEach([&](auto e) { mask[(size_t)e.y * w + e.x] = e.value.has_value() && *e.value; });
      const int nv = (int)shape.count();
      std::vector<int32_t> off(nv + 1, 0);
      const int total = detail::check(vsr_sensor_config_voxels(name.c_str(), w, h, mask.data(), nullptr, 0, off.data()), nullptr, "RobotUtils.buildSensorizingFunction");
      std::vector<vsr_sensor_desc> all(total > 0 ? total : 1);
      detail::check(vsr_sensor_config_voxels(name.c_str(), w, h, mask.data(), all.data(), total, off.data()), nullptr, "RobotUtils.buildSensorizingFunction");
      Grid<Voxel> body(w, h);
      int v = 0;
      shape.forEach([&](auto e) {
        if (!(e.value.has_value() && *e.value)) return;
        body.set(e.x, e.y, Voxel{std::vector<vsr_sensor_desc>(all.begin() + off[v], all.begin() + off[v + 1])});
        v++;
      });
      return body;
    };
  }
};

// RobotUtils.buildRobotTransformation(name, externalRandom) [U] (locomotion/Starter.java:126,428): `identity`, or
// `broken-<ratio>-<seed|rnd>`: every voxel becomes, with probability `ratio`, a voxel with a frozen actuator; the Random is created once
// here (seed) or is the caller's (`rnd`), so consecutive robots continue one stream.  `breakable-...` is outside this build.
inline std::function<Robot(const Robot&)> buildRobotTransformation(const std::string& name, std::shared_ptr<JavaRandom> externalRandom = nullptr) {
  int32_t kind = 0, seedIsRnd = 0;
  double ratio = 0;
  int64_t seed = 0;
  detail::check(vsr_robot_transformation(name.c_str(), &kind, &ratio, &seed, &seedIsRnd), nullptr, "RobotUtils.buildRobotTransformation");
  if (kind == 0) return [](const Robot& r) { return r; };
  std::shared_ptr<JavaRandom> random = seedIsRnd ? externalRandom : std::make_shared<JavaRandom>(seed);
  if (!random) throw IllegalArgumentException("transformation with seed `rnd` needs the run's Random");
  return [ratio, random](const Robot& robot) {
    Grid<Voxel> body = robot.getVoxels();
    robot.getVoxels().forEach([&](auto e) {
      if (!e.value || random->nextDouble() > ratio) return;
      Voxel broken = *e.value;
      broken.actuatorFrozen = true;
      random->nextInt(); // the BreakableVoxel's own seed
      body.set(e.x, e.y, broken);
    });
    return Robot(robot.getController(), body);
  };
}

// ---- builders (builder/*.java) ------------------------------------------------------------------------------------------------------------
using Params = std::map<std::string, std::string>;
// NamedProvider.build(String): "name;k=v;k=v" -> {NAME: name, k: v, ...}
inline Params parseParams(const std::string& s) {
  Params p;
  std::stringstream ss(s);
  std::string tok;
  while (std::getline(ss, tok, ';')) {
    const size_t eq = tok.find('=');
    if (eq == std::string::npos) p["NAME"] = tok; else p[tok.substr(0, eq)] = tok.substr(eq + 1);
  }
  return p;
}
inline std::string getOrDefault(const Params& p, const std::string& k, const std::string& d) { auto it = p.find(k); return it == p.end() ? d : it->second; }

template <class A, class B>
struct PrototypedFunctionBuilder {
  std::function<std::function<B(const A&)>(const B&)> buildFor;
  std::function<A(const B&)> exampleFor;
  // this.compose(other): C -> A -> B, the A built for this builder's own example of b (PrototypedFunctionBuilder.java:17-29)
  template <class C>
  PrototypedFunctionBuilder<C, B> compose(const PrototypedFunctionBuilder<C, A>& other) const {
    const PrototypedFunctionBuilder<A, B> self = *this;
    return {[self, other](const B& b) {
              auto outer = self.buildFor(b);
              auto inner = other.buildFor(self.exampleFor(b));
              return std::function<B(const C&)>([outer, inner](const C& c) { return outer(inner(c)); });
            },
            [self, other](const B& b) { return other.exampleFor(self.exampleFor(b)); }};
  }
};

struct MLP {
  static std::vector<int> innerNeurons(int nOfInputs, int nOfOutputs, double innerLayerRatio, int nOfInnerLayers) {
    std::vector<int> inner((size_t)(nOfInnerLayers > 0 ? nOfInnerLayers : 0));
    detail::check(vsr_mlp_inner_neurons(nOfInputs, nOfOutputs, innerLayerRatio, nOfInnerLayers, inner.data()), nullptr, "MLP.innerNeurons");
    return inner;
  }
  static PrototypedFunctionBuilder<std::vector<double>, FunctionPtr> build(const Params& params) {
    const double r = std::stod(getOrDefault(params, "r", "0.65"));
    const int nIL = std::stoi(getOrDefault(params, "nIL", "1"));
    return {[r, nIL](const FunctionPtr& function) {
              return std::function<FunctionPtr(const std::vector<double>&)>([r, nIL, function](const std::vector<double>& values) -> FunctionPtr {
                const int nIn = function->getInputDimension(), nOut = function->getOutputDimension();
                const std::vector<int> inner = innerNeurons(nIn, nOut, r, nIL);
                const int nW = MultiLayerPerceptron::countWeights(nIn, inner, nOut);
                if (nW != (int)values.size())
                  throw IllegalArgumentException(detail::format("Wrong number of values for weights: %d expected, %d found", nW, (int)values.size()));
                return std::make_shared<MultiLayerPerceptron>(nIn, inner, nOut, values);
              });
            },
            [r, nIL](const FunctionPtr& function) {
              const int nIn = function->getInputDimension(), nOut = function->getOutputDimension();
              return std::vector<double>((size_t)MultiLayerPerceptron::countWeights(nIn, innerNeurons(nIn, nOut, r, nIL), nOut), 0.0);
            }};
  }
};

struct BrainCentralized {
  static PrototypedFunctionBuilder<FunctionPtr, Robot> build(const Params&) {
    return {[](const Robot& robot) {
              const Grid<Voxel> body = robot.getVoxels();
              return std::function<Robot(const FunctionPtr&)>([body](const FunctionPtr& function) {
                if (function->getInputDimension() != CentralizedSensing::nOfInputs(body))
                  throw IllegalArgumentException(detail::format("Wrong number of function input args: %d expected, %d found",
                                                                CentralizedSensing::nOfInputs(body), function->getInputDimension()));
                if (function->getOutputDimension() != CentralizedSensing::nOfOutputs(body))
                  throw IllegalArgumentException(detail::format("Wrong number of function output args: %d expected, %d found",
                                                                CentralizedSensing::nOfOutputs(body), function->getOutputDimension()));
                return Robot(std::make_shared<CentralizedSensing>(body, function), body);
              });
            },
            [](const Robot& robot) { return RealFunction::build(CentralizedSensing::nOfInputs(robot.getVoxels()), CentralizedSensing::nOfOutputs(robot.getVoxels())); }};
  }
};

struct BrainPhaseValues {
  static PrototypedFunctionBuilder<std::vector<double>, Robot> build(const Params& params) {
    const double frequency = std::stod(getOrDefault(params, "f", "1")), amplitude = std::stod(getOrDefault(params, "a", "1"));
    return {[frequency, amplitude](const Robot& robot) {
              const Grid<Voxel> body = robot.getVoxels();
              return std::function<Robot(const std::vector<double>&)>([frequency, amplitude, body](const std::vector<double>& values) {
                const int nVox = (int)body.count();
                if (nVox != (int)values.size())
                  throw IllegalArgumentException(detail::format("Wrong number of values: %d expected, %d found", nVox, (int)values.size()));
                Grid<double> phases(body.getW(), body.getH());
                int c = 0;
                body.forEach([&](auto e) { if (e.value) phases.set(e.x, e.y, values[c++]); });
                return Robot(std::make_shared<PhaseSin>(frequency, amplitude, phases), body);
              });
            },
            [](const Robot& robot) { return std::vector<double>(robot.getVoxels().count(), 0.0); }};
  }
};

struct BrainHomoDistributed {
  static std::pair<int, int> getIODim(const Robot& robot, int signals) {
    const Grid<Voxel>& body = robot.getVoxels();
    const std::vector<Voxel> voxels = body.values();
    if (voxels.empty()) throw IllegalArgumentException("Target robot has no voxels");
    const int nIn = DistributedSensing::nOfInputs(voxels.front(), signals), nOut = DistributedSensing::nOfOutputs(voxels.front(), signals);
    std::string where, have;
    body.forEach([&](auto e) {
      if (!e.value || DistributedSensing::nOfInputs(*e.value, signals) == nIn) return;
      where += (where.empty() ? "" : ",") + detail::format("(%d,%d)", e.x, e.y);
      have += (have.empty() ? "" : ",") + std::to_string(DistributedSensing::nOfInputs(*e.value, signals));
    });
    if (!where.empty())
      throw IllegalArgumentException(detail::format("Cannot build BrainHomoDistributed robot mapper for this body: all voxels should have %d inputs, but voxels at positions %s have %s",
                                                    nIn, where.c_str(), have.c_str()));
    return {nIn, nOut};
  }
  static PrototypedFunctionBuilder<FunctionPtr, Robot> build(const Params& params) {
    const int signals = std::stoi(getOrDefault(params, "s", "1"));
    return {[signals](const Robot& robot) {
              const auto dim = getIODim(robot, signals);
              const Grid<Voxel> body = robot.getVoxels();
              return std::function<Robot(const FunctionPtr&)>([signals, dim, body](const FunctionPtr& function) {
                if (function->getInputDimension() != dim.first)
                  throw IllegalArgumentException(detail::format("Wrong number of function input args: %d expected, %d found", dim.first, function->getInputDimension()));
                if (function->getOutputDimension() != dim.second)
                  throw IllegalArgumentException(detail::format("Wrong number of function output args: %d expected, %d found", dim.second, function->getOutputDimension()));
                auto controller = std::make_shared<DistributedSensing>(body, signals);
                body.forEach([&](auto e) { if (e.value) controller->getFunctions().set(e.x, e.y, function); });
                return Robot(controller, body);
              });
            },
            [signals](const Robot& robot) { const auto dim = getIODim(robot, signals); return RealFunction::build(dim.first, dim.second); }};
  }
};

// Starter.getMapper (locomotion/Starter.java:192-247) for the mappers of the hot path: "brainCentralized<mlp;r=0.65;nIL=1",
// "brainHomoDist;s=1<mlp;r=0.65;nIL=1", "brainPhaseVals;f=1"
inline PrototypedFunctionBuilder<std::vector<double>, Robot> buildMapper(const std::string& name) {
  const size_t pipe = name.find('<');
  const Params robotParams = parseParams(name.substr(0, pipe));
  const std::string robotName = getOrDefault(robotParams, "NAME", "");
  if (robotName == "brainPhaseVals") {
    if (pipe != std::string::npos) throw IllegalArgumentException("brainPhaseVals takes no function mapper");
    return BrainPhaseValues::build(robotParams);
  }
  if (pipe == std::string::npos) throw IllegalArgumentException("Unknown mapper name: " + name);
  const Params functionParams = parseParams(name.substr(pipe + 1));
  if (getOrDefault(functionParams, "NAME", "") != "mlp") throw IllegalArgumentException("Unknown mapper name: " + name.substr(pipe + 1));
  if (robotName == "brainCentralized") return BrainCentralized::build(robotParams).compose(MLP::build(functionParams));
  if (robotName == "brainHomoDist") return BrainHomoDistributed::build(robotParams).compose(MLP::build(functionParams));
  throw IllegalArgumentException("Unknown mapper name: " + name);
}

// ---- the task ------------------------------------------------------------------------------------------------------------------------------
// Outcome as far as the hot path reads it (locomotion/NamedFunctions.java:66-72,135-137), already restricted to [transientT, finalT)
struct Outcome {
  double velocity = 0, distance = 0, time = 0, controlEnergy = 0, areaRatioEnergy = 0;
  double getVelocity() const { return velocity; }
  double getDistance() const { return distance; }
  double getTime() const { return time; }
  double getControlPower() const { return controlEnergy / time; }
  double getAreaRatioPower() const { return areaRatioEnergy / time; }
  double getCorrectedEfficiency() const { return distance / (1.0 + controlEnergy); }
  // Outcome.getAveragePosture(n) (Starter.java:180-190 numSpeciated features, NamedFunctions.java:356): [y * n + x], y = 0 the bottom row;
  // recorded by Locomotion::applyAllObserved(robots, n)
  std::vector<bool> averagePosture;
  int postureN = 0;
  const std::vector<bool>& getAveragePosture(int n) const {
    if (n != postureN) throw VsrException(VSR_ESTATE, "average posture of side " + std::to_string(n) + " was not recorded: use Locomotion::applyAllObserved");
    return averagePosture;
  }
};

// new Locomotion(finalT, Locomotion.createTerrain(name), new Settings()) followed by .subOutcome(transientT, finalT)
// (locomotion/Starter.java:83-94,427-435); `devices` shards every batch by index over those GPUs (empty: the current device)
class Locomotion {
 public:
  Locomotion(double finalT, const std::string& terrainName, double transientT = 0.0, std::vector<int> devices = {}, int precision = 32)
      : finalT_(finalT), transientT_(transientT), devices_(std::move(devices)), precision_(precision) {
    createTerrain(terrainName, xs_, ys_);
  }
  Locomotion(const Locomotion&) = delete;
  Locomotion& operator=(const Locomotion&) = delete;

  static void createTerrain(const std::string& name, std::vector<double>& xs, std::vector<double>& ys) {
    const int n = detail::check(vsr_terrain(name.c_str(), nullptr, nullptr, 0), nullptr, "Locomotion.createTerrain");
    xs.resize(n); ys.resize(n);
    detail::check(vsr_terrain(name.c_str(), xs.data(), ys.data(), n), nullptr, "Locomotion.createTerrain");
  }

  virtual ~Locomotion() { for (auto& kv : ctxs_) vsr_ctx_destroy(kv.second); }

  struct Terrain { std::vector<double> xs, ys; };
  // per-evaluation terrains of a `-rnd` task (RandomTerrainLocomotion), none for a shared terrain
  virtual std::vector<Terrain> terrainsFor(size_t) { return {}; }
  // bit v (voxel order) = the actuator of voxel v is frozen (what a `broken-...` transformation left behind)
  static uint32_t frozenMask(const Robot& robot) {
    uint32_t m = 0;
    int v = 0;
    robot.getVoxels().forEach([&](auto e) { if (e.value) { if (e.value->actuatorFrozen) m |= 1u << v; v++; } });
    return m;
  }

  Outcome apply(const Robot& robot) { return applyAll({robot}).front(); }

  // one launch per distinct (body, sensors, controller structure); the outcomes come back in the order of `robots`.
  // `terrains`: one polyline per robot (what a `-rnd` task draws itself, in robot order, before any grouping)
  std::vector<Outcome> applyAll(const std::vector<Robot>& robots, std::vector<Terrain> terrains = {}) { return evaluate(robots, std::move(terrains), 0); }
  // ... that also keeps every robot's average posture on the postureN x postureN grid (the `numSpeciated` features), same launch
  std::vector<Outcome> applyAllObserved(const std::vector<Robot>& robots, int postureN = 8) { return evaluate(robots, {}, postureN); }

 private:
  std::vector<Outcome> evaluate(const std::vector<Robot>& robots, std::vector<Terrain> terrains, int postureN) {
    if (terrains.empty()) terrains = terrainsFor(robots.size());
    if (!terrains.empty() && terrains.size() != robots.size()) throw IllegalArgumentException("one terrain per robot expected");
    std::vector<Outcome> out(robots.size());
    std::vector<vsr_ctx*> order;
    std::map<vsr_ctx*, std::vector<size_t>> groups;
    for (size_t i = 0; i < robots.size(); i++) {
      vsr_ctx* ctx = contextOf(robots[i]);
      if (!groups.count(ctx)) order.push_back(ctx);
      groups[ctx].push_back(i);
    }
    for (vsr_ctx* ctx : order) {
      const std::vector<size_t>& idx = groups[ctx];
      const int G = detail::check(vsr_genome_len(ctx), ctx, "vsr_genome_len");
      std::vector<double> genomes(idx.size() * (size_t)G), fitness(idx.size());
      std::vector<vsr_result> results(idx.size());
      for (size_t k = 0; k < idx.size(); k++) {
        const std::vector<double> g = genomeOf(robots[idx[k]]);
        if ((int)g.size() != G) throw IllegalArgumentException(detail::format("Wrong number of controller parameters: %d expected, %d found", G, (int)g.size()));
        std::memcpy(&genomes[k * (size_t)G], g.data(), sizeof(double) * G);
      }
      std::vector<uint32_t> frozen(idx.size());
      bool anyFrozen = false;
      for (size_t k = 0; k < idx.size(); k++) anyFrozen |= (frozen[k] = frozenMask(robots[idx[k]])) != 0;
      std::vector<uint8_t> posture(postureN > 0 ? idx.size() * (size_t)postureN * postureN : 0);
      if (!anyFrozen && terrains.empty() && postureN <= 0) {
        detail::check(vsr_eval(ctx, genomes.data(), (int32_t)idx.size(), fitness.data(), results.data()), ctx, "vsr_eval");
      } else { // per-evaluation inputs (transformation masks, terrains) and / or observation summaries
        vsr_eval_opts opts;
        std::memset(&opts, 0, sizeof opts);
        if (postureN > 0) { opts.posture_n = postureN; opts.posture_out = posture.data(); }
        std::vector<double> tx, ty;
        std::vector<int32_t> toff{0};
        if (anyFrozen) opts.frozen_masks = frozen.data();
        if (!terrains.empty()) {
          for (size_t i : idx) {
            tx.insert(tx.end(), terrains[i].xs.begin(), terrains[i].xs.end());
            ty.insert(ty.end(), terrains[i].ys.begin(), terrains[i].ys.end());
            toff.push_back((int32_t)tx.size());
          }
          opts.terrain_x = tx.data(); opts.terrain_y = ty.data(); opts.terrain_off = toff.data();
        }
        detail::check(vsr_eval_ex(ctx, genomes.data(), (int32_t)idx.size(), &opts, fitness.data(), results.data()), ctx, "vsr_eval_ex");
      }
      for (size_t k = 0; k < idx.size(); k++) {
        const vsr_result& r = results[k];
        Outcome& o = out[idx[k]];
        o.velocity = fitness[k]; // NaN if the contact pool overflowed
        o.distance = r.distance; o.time = r.time; o.controlEnergy = r.control_energy; o.areaRatioEnergy = r.area_ratio_abs;
        if (postureN > 0) {
          o.postureN = postureN;
          o.averagePosture.assign(posture.begin() + k * (size_t)postureN * postureN, posture.begin() + (k + 1) * (size_t)postureN * postureN);
        }
      }
    }
    return out;
  }


 public:
  // the controller parameters in the order the kernel reads them (== the mappers' genome order)
  static std::vector<double> genomeOf(const Robot& robot) {
    const Controller* c = robot.getController().get();
    if (auto* p = dynamic_cast<const PhaseSin*>(c)) {
      std::vector<double> g;
      robot.getVoxels().forEach([&](auto e) { if (e.value) g.push_back(p->phases.get(e.x, e.y).value_or(0.0)); });
      return g;
    }
    if (auto* cs = dynamic_cast<const CentralizedSensing*>(c)) return mlpOf(cs->function).getParams();
    if (auto* ds = dynamic_cast<const DistributedSensing*>(c)) return mlpOf(firstFunction(*ds)).getParams();
    throw VsrException(VSR_EUNSUPPORTED, "controller type outside the hot path");
  }

 private:
  static const MultiLayerPerceptron& mlpOf(const FunctionPtr& f) {
    auto* m = dynamic_cast<const MultiLayerPerceptron*>(f.get());
    if (!m) throw VsrException(VSR_EUNSUPPORTED, "only MultiLayerPerceptron functions are built (builder/function/MLP.java)");
    return *m;
  }
  static FunctionPtr firstFunction(const DistributedSensing& ds) {
    const std::vector<FunctionPtr> fs = ds.getFunctions().values();
    if (fs.empty()) throw IllegalArgumentException("Target robot has no voxels");
    for (const auto& f : fs)
      if (f != fs.front() && mlpOf(f).getParams() != mlpOf(fs.front()).getParams())
        throw VsrException(VSR_EUNSUPPORTED, "heterogeneous DistributedSensing: use the C ABI's VSR_CTRL_MLP_DISTRIBUTED_HETERO");
    return fs.front();
  }
  template <class T> static void put(std::string& key, const T& v) { key.append(reinterpret_cast<const char*>(&v), sizeof v); }

  vsr_ctx* contextOf(const Robot& robot) {
    const Grid<Voxel>& body = robot.getVoxels();
    const int w = body.getW(), h = body.getH();
    if (body.count() == 0) throw IllegalArgumentException("Target robot has no voxels");
    std::vector<uint8_t> mask((size_t)w * h, 0);
    std::vector<vsr_sensor_desc> sensors;
    std::vector<int32_t> off{0};
    body.forEach([&](auto e) {
      if (!e.value) return;
      mask[(size_t)e.y * w + e.x] = 1;
      sensors.insert(sensors.end(), e.value->sensors.begin(), e.value->sensors.end());
      off.push_back((int32_t)sensors.size());
    });
    // structure key: body, every voxel's sensor list, controller kind + sizes
    std::string key;
    put(key, w); put(key, h);
    key.append(mask.begin(), mask.end());
    for (int32_t o : off) put(key, o);
    for (const auto& s : sensors) { put(key, s.kind); put(key, s.axes); put(key, s.rotated); put(key, s.soft_norm); put(key, s.dom_min); put(key, s.dom_max);
                                    put(key, s.avg_window); put(key, s.noise_sigma); put(key, s.param); put(key, s.n_rays); put(key, s.side); }
    const Controller* c = robot.getController().get();
    vsr_ctrl_kind kind;
    std::vector<int32_t> layers;
    int signals = 0;
    double freq = 1.0, amp = 1.0;
    if (auto* p = dynamic_cast<const PhaseSin*>(c)) { kind = VSR_CTRL_PHASE_SIN; freq = p->frequency; amp = p->amplitude; }
    else if (auto* cs = dynamic_cast<const CentralizedSensing*>(c)) { kind = VSR_CTRL_MLP_CENTRALIZED; layers.assign(mlpOf(cs->function).getNeurons().begin(), mlpOf(cs->function).getNeurons().end()); }
    else if (auto* ds = dynamic_cast<const DistributedSensing*>(c)) {
      kind = VSR_CTRL_MLP_DISTRIBUTED_HOMO; signals = ds->signals;
      const FunctionPtr f = firstFunction(*ds);
      layers.assign(mlpOf(f).getNeurons().begin(), mlpOf(f).getNeurons().end());
    } else throw VsrException(VSR_EUNSUPPORTED, "controller type outside the hot path");
    put(key, (int)kind); put(key, signals); put(key, freq); put(key, amp);
    for (int32_t l : layers) put(key, l);
    auto it = ctxs_.find(key);
    if (it != ctxs_.end()) return it->second;

    vsr_env_desc env;
    detail::check(vsr_env_default(&env), nullptr, "vsr_env_default");
    env.final_t = finalT_; env.transient_t = transientT_; env.precision = precision_;
    env.terrain_x = xs_.data(); env.terrain_y = ys_.data(); env.n_terrain = (int32_t)xs_.size();
    env.n_devices = (int32_t)devices_.size();
    for (size_t i = 0; i < devices_.size() && i < 8; i++) env.devices[i] = devices_[i];
    vsr_ctx* ctx = nullptr;
    detail::check(vsr_ctx_create(&env, &ctx), nullptr, "vsr_ctx_create");
    try {
      detail::check(vsr_body_set_voxels(ctx, w, h, mask.data(), sensors.data(), off.data()), ctx, "vsr_body_set_voxels");
      detail::check(vsr_controller_set(ctx, kind, layers.data(), (int32_t)layers.size(), signals, freq, amp, 1), ctx, "vsr_controller_set");
    } catch (...) { vsr_ctx_destroy(ctx); throw; }
    ctxs_.emplace(key, ctx);
    return ctx;
  }

  double finalT_, transientT_;
  std::vector<int> devices_;
  int precision_;
  std::vector<double> xs_, ys_;
  std::map<std::string, vsr_ctx*> ctxs_;
};

// Batching behind JGEA's one-robot-per-call contract (`qualityFunction.apply(robot)` once per individual from a thread pool,
// locomotion/Starter.java:427-436; SURVEY 8b "granularity mismatch"), with no extra threads and no time-outs: apply(robot) only records
// the robot and returns a LazyOutcome; JGEA builds its individuals without reading the fitness object and reads it when it orders the
// new population -- the first getter called on any outcome of the batch sends the whole generation to the device in one applyAll.
// The Java form is java/it/units/erallab/gpu/{BatchedLocomotion,GpuOutcome}.java.  Thread-safe like the Java one.
class LazyBatchedLocomotion {
  struct Batch {
    Locomotion* task;
    std::mutex lock;
    std::vector<Robot> robots;
    std::vector<Outcome> outcomes;
    std::exception_ptr error;
    bool done = false;
  };
 public:
  class LazyOutcome {
   public:
    double getVelocity() const { return resolved().getVelocity(); }
    double getDistance() const { return resolved().getDistance(); }
    double getTime() const { return resolved().getTime(); }
    double getControlPower() const { return resolved().getControlPower(); }
    double getAreaRatioPower() const { return resolved().getAreaRatioPower(); }
    double getCorrectedEfficiency() const { return resolved().getCorrectedEfficiency(); }
   private:
    friend class LazyBatchedLocomotion;
    LazyOutcome(LazyBatchedLocomotion* owner, std::shared_ptr<Batch> batch, size_t slot) : owner_(owner), batch_(std::move(batch)), slot_(slot) {}
    const Outcome& resolved() const {
      owner_->close(batch_); // robots applied from now on belong to the next batch
      std::lock_guard<std::mutex> g(batch_->lock);
      if (!batch_->done) {
        try { batch_->outcomes = batch_->task->applyAll(batch_->robots); owner_->flushed_++; }
        catch (...) { batch_->error = std::current_exception(); } // every outcome of the batch reports it, like Locomotion.apply throwing
        batch_->done = true;
      }
      if (batch_->error) std::rethrow_exception(batch_->error);
      return batch_->outcomes[slot_];
    }
    LazyBatchedLocomotion* owner_;
    std::shared_ptr<Batch> batch_;
    size_t slot_;
  };
  explicit LazyBatchedLocomotion(Locomotion& task) : task_(&task) { open_ = newBatch(); }
  LazyOutcome apply(const Robot& robot) {
    std::lock_guard<std::mutex> g(lock_);
    open_->robots.push_back(robot);
    return LazyOutcome(this, open_, open_->robots.size() - 1);
  }
  int batchesFlushed() const { return flushed_; }
 private:
  std::shared_ptr<Batch> newBatch() { auto b = std::make_shared<Batch>(); b->task = task_; return b; }
  void close(const std::shared_ptr<Batch>& batch) {
    std::lock_guard<std::mutex> g(lock_);
    if (open_ == batch) open_ = newBatch();
  }
  Locomotion* task_;
  std::mutex lock_;
  std::shared_ptr<Batch> open_;
  std::atomic<int> flushed_{0};
};

// r -> new Locomotion(T, createTerrain(name.replace("-rnd", "-" + random.nextInt(10000))), settings).apply(r) (Starter.java:89-93):
// every evaluation gets its own terrain, the batch still goes to the device in one call.  Seeds are drawn in robot order (the
// reference draws them in thread-completion order, which is not reproducible).
inline std::string replaceRnd(const std::string& terrainName, int seed) {
  const size_t at = terrainName.find("-rnd");
  if (at == std::string::npos) throw IllegalArgumentException("terrain name without -rnd");
  return terrainName.substr(0, at) + "-" + std::to_string(seed) + terrainName.substr(at + 4);
}
class RandomTerrainLocomotion : public Locomotion {
 public:
  RandomTerrainLocomotion(double finalT, const std::string& terrainName, std::shared_ptr<JavaRandom> random, double transientT = 0.0)
      : Locomotion(finalT, replaceRnd(terrainName, 0), transientT), terrainName_(terrainName), random_(std::move(random)) {}
  std::vector<Terrain> terrainsFor(size_t n) override {
    std::vector<Terrain> ts(n);
    for (auto& t : ts) createTerrain(replaceRnd(terrainName_, random_->nextInt(10000)), t.xs, t.ys);
    return ts;
  }
 private:
  std::string terrainName_;
  std::shared_ptr<JavaRandom> random_;
};

// Starter.buildLocomotionTask(terrainName, episodeT, random, cacheOutcome) (Starter.java:83-94): one shared Locomotion for a fixed
// terrain, a per-evaluation terrain for `-rnd` names; the subOutcome(transientT, episodeT) of the quality function (:430-433) is
// folded in; the fitness is Outcome::getVelocity, maximised (:174)
inline std::unique_ptr<Locomotion> buildLocomotionTask(const std::string& terrainName, double episodeT, double transientT = 0.0,
                                                       std::shared_ptr<JavaRandom> random = nullptr) {
  if (terrainName.find("-rnd") != std::string::npos)
    return std::make_unique<RandomTerrainLocomotion>(episodeT, terrainName, random ? random : std::make_shared<JavaRandom>(0), transientT);
  return std::make_unique<Locomotion>(episodeT, terrainName, transientT);
}

// Starter.validate (locomotion/Starter.java:122-135) for a batch of robots.  The reference validates ONE robot per call: random = new
// Random(seed); robot = buildRobotTransformation(name, random).apply(robot); buildLocomotionTask(terrain, T, random, false).apply(robot)
// -- every robot restarts the same stream: the transformation's draws (only a `-rnd` seeded one takes them from `random`), then the
// `-rnd` terrain seed.  The batch does exactly that per robot and evaluates all of them in one launch per morphology.
struct ValidationOutcome {
  std::string terrainName, transformationName;
  int64_t seed;
  Outcome outcome;
};
inline std::vector<ValidationOutcome> validate(const std::vector<Robot>& robots, const std::string& terrainName, const std::string& transformationName,
                                               int64_t seed, double episodeTime, double transientTime) {
  const bool rnd = terrainName.find("-rnd") != std::string::npos;
  std::vector<Robot> transformed;
  std::vector<Locomotion::Terrain> terrains;
  for (const Robot& r : robots) {
    auto random = std::make_shared<JavaRandom>(seed);
    transformed.push_back(buildRobotTransformation(transformationName, random)(r));
    if (rnd) {
      terrains.emplace_back();
      Locomotion::createTerrain(replaceRnd(terrainName, random->nextInt(10000)), terrains.back().xs, terrains.back().ys);
    }
  }
  Locomotion task(episodeTime, rnd ? replaceRnd(terrainName, 0) : terrainName, transientTime);
  std::vector<ValidationOutcome> out;
  for (const Outcome& o : task.applyAll(transformed, terrains)) out.push_back({terrainName, transformationName, seed, o});
  return out;
}

} // namespace vsr_host
