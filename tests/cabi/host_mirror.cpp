// host_mirror.cpp -- include/vsr_host.hpp (the C++ host side: the reference's mapper / task interface over the C ABI) used the way
// locomotion/Starter.java:402-436 uses the Java one: buildShape + buildSensorizingFunction -> target robot, getMapper(...).buildFor(target)
// -> genotype -> Robot, Locomotion -> Outcome::getVelocity.
//
//   host_mirror check                                                       host-side checks only (dimensions, error messages); no GPU needed
//   host_mirror eval <shape> <sensors> <mapper> <terrain> <T> <transient> <n> <seed> <out.bin>
//   host_mirror validate <shape> <sensors> <mapper> <terrain> <T> <transient> <n> <seed> <out.bin> <transformation> <validation seed>
//
//   g++ -std=c++17 -O2 -pthread -Iinclude tests/cabi/host_mirror.cpp -Lvsrevolution_b200 -lvsr_eval -Wl,-rpath,$PWD/vsrevolution_b200 -o host_mirror
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <optional>
#include <string>
#include <thread>

#include "vsr_host.hpp"

using namespace vsr_host;

static Robot targetRobot(const std::string& shape, const std::string& sensors) {
  // Starter.java:402-406: new Robot(Controller.empty(), buildSensorizingFunction(sensorConfig).apply(buildShape(shape)))
  return Robot(nullptr, RobotUtils::buildSensorizingFunction(sensors.empty() ? "empty" : sensors)(RobotUtils::buildShape(shape)));
}

#define EXPECT(cond)                                                      \
  do {                                                                    \
    if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } \
  } while (0)

template <class F>
static std::string messageOf(F f) {
  try { f(); } catch (const IllegalArgumentException& e) { return e.what(); }
  return "";
}

static int check() {
  const Robot biped = targetRobot("biped-4x3", "uniform-a-0"), worm = targetRobot("worm-8x2", ""), comb = targetRobot("comb-7x2", "uniform-t+vxy+a-0");
  EXPECT(biped.getVoxels().count() == 10 && worm.getVoxels().count() == 16 && comb.getVoxels().count() == 11);
  EXPECT(!biped.getVoxels().get(1, 0).has_value() && biped.getVoxels().get(0, 0).has_value()); // the notch of the biped
  EXPECT(CentralizedSensing::nOfInputs(biped.getVoxels()) == 10 && CentralizedSensing::nOfOutputs(biped.getVoxels()) == 10);
  EXPECT(DistributedSensing::nOfInputs(comb.getVoxels().values().front(), 1) == 8 && DistributedSensing::nOfOutputs(comb.getVoxels().values().front(), 1) == 5);
  // genotype sizes of the BASELINE configs (157, 16, 75) from the mappers' own examples
  const auto central = buildMapper("brainCentralized<mlp;r=0.65;nIL=1"), phases = buildMapper("brainPhaseVals;f=1"), homo = buildMapper("brainHomoDist;s=1<mlp;r=0.65;nIL=1");
  EXPECT(central.exampleFor(biped).size() == 157 && phases.exampleFor(worm).size() == 16 && homo.exampleFor(comb).size() == 75);
  EXPECT((MLP::innerNeurons(10, 10, 0.65, 1) == std::vector<int>{7}) && (MLP::innerNeurons(8, 5, 0.65, 1) == std::vector<int>{5}));
  EXPECT((MLP::innerNeurons(10, 4, 0.65, 3) == std::vector<int>{9, 7, 6}));
  // the mapped robots carry what the mappers built
  const Robot r = central.buildFor(biped)(std::vector<double>(157, 0.25));
  const auto* cs = dynamic_cast<const CentralizedSensing*>(r.getController().get());
  EXPECT(cs && cs->function->getInputDimension() == 10 && cs->function->getOutputDimension() == 10 && Locomotion::genomeOf(r).size() == 157);
  std::vector<double> ph(16);
  for (int i = 0; i < 16; i++) ph[i] = 0.1 * i;
  const Robot w = phases.buildFor(worm)(ph);
  const auto* ps = dynamic_cast<const PhaseSin*>(w.getController().get());
  EXPECT(ps && ps->frequency == 1.0 && ps->phases.get(3, 1).value() == 0.1 * (8 + 3) && Locomotion::genomeOf(w) == ph); // row-major, y = 0 first
  // error behaviour: the Java exception type and message
  EXPECT(messageOf([&] { central.buildFor(biped)(std::vector<double>(156, 0.0)); }) == "Wrong number of values for weights: 157 expected, 156 found");
  EXPECT(messageOf([&] { phases.buildFor(worm)(std::vector<double>(15, 0.0)); }) == "Wrong number of values: 16 expected, 15 found");
  EXPECT(messageOf([&] { BrainCentralized::build({}).buildFor(biped)(RealFunction::build(9, 10)); }) == "Wrong number of function input args: 10 expected, 9 found");
  EXPECT(messageOf([&] { BrainHomoDistributed::build({{"s", "2"}}).buildFor(comb)(RealFunction::build(12, 8)); }) == "Wrong number of function output args: 9 expected, 8 found");
  EXPECT(!messageOf([&] { RobotUtils::buildShape("hexapod-4x3"); }).empty() && !messageOf([&] { buildMapper("brainQuantum<mlp"); }).empty());
  EXPECT(!messageOf([&] { Locomotion(10.0, "lunar-3"); }).empty());
  // java.util.Random: the known first double of new Random(0), and nextInt(bound) against the library's own replay
  EXPECT(JavaRandom(0).nextDouble() == 0.730967787376657);
  {
    JavaRandom jr(42);
    int32_t lib[50];
    EXPECT(vsr_java_random_ints(42, 10000, lib, 50) >= 0);
    for (int i = 0; i < 50; i++) EXPECT(jr.nextInt(10000) == lib[i]);
    JavaRandom p2(7);
    EXPECT(vsr_java_random_ints(7, 1024, lib, 10) >= 0); // power-of-two bound: the other branch of nextInt(bound)
    for (int i = 0; i < 10; i++) EXPECT(p2.nextInt(1024) == lib[i]);
  }
  // robot transformations: identity keeps the robot, broken-<ratio>-<seed> freezes actuators reproducibly, ratio 0 / 1 are the extremes
  EXPECT(Locomotion::frozenMask(buildRobotTransformation("identity")(r)) == 0);
  EXPECT(Locomotion::frozenMask(buildRobotTransformation("broken-0-1")(r)) == 0 && Locomotion::frozenMask(buildRobotTransformation("broken-1-1")(r)) == 0x3ffu);
  const uint32_t m1 = Locomotion::frozenMask(buildRobotTransformation("broken-0.5-7")(r)), m2 = Locomotion::frozenMask(buildRobotTransformation("broken-0.5-7")(r));
  EXPECT(m1 == m2 && m1 != 0 && m1 != 0x3ffu);
  EXPECT(!messageOf([&] { buildRobotTransformation("broken-0.5-rnd"); }).empty() && !messageOf([&] { buildRobotTransformation("melted-0.5-1"); }).empty());
  EXPECT(replaceRnd("hilly-1-10-rnd", 123) == "hilly-1-10-123");
  // the task itself needs a GPU: without one the library refuses (no CPU fallback)
  try {
    Locomotion task(2.0, "flat", 0.5);
    const Outcome o = task.apply(r);
    std::printf("gpu: velocity %.17g distance %.17g time %.17g\n", o.getVelocity(), o.getDistance(), o.getTime());
    EXPECT(o.getTime() > 1.4 && o.getTime() < 1.6 && o.getVelocity() == o.getDistance() / o.getTime());
    // JGEA's contract, one apply(robot) per individual: a generation is ONE launch, started by the first fitness that is read
    LazyBatchedLocomotion lazy(task);
    std::vector<LazyBatchedLocomotion::LazyOutcome> generation;
    std::vector<Robot> robots;
    for (int i = 0; i < 12; i++) {
      robots.push_back(central.buildFor(biped)(std::vector<double>(157, 0.05 * i - 0.3)));
      generation.push_back(lazy.apply(robots.back()));
    }
    EXPECT(lazy.batchesFlushed() == 0);
    const std::vector<Outcome> direct = task.applyAll(robots);
    for (int i = 11; i >= 0; i--) EXPECT(generation[i].getVelocity() == direct[i].getVelocity());
    EXPECT(lazy.batchesFlushed() == 1);
    const auto next = lazy.apply(robots[3]); // the next generation opens a new batch
    EXPECT(next.getVelocity() == direct[3].getVelocity() && lazy.batchesFlushed() == 2);
    { // JGEA applies the quality function from a thread pool (Starter.java:424-436): concurrent apply(), then concurrent reads
      std::vector<std::optional<LazyBatchedLocomotion::LazyOutcome>> slots(12);
      std::vector<std::thread> pool;
      for (int t = 0; t < 4; t++) pool.emplace_back([&, t] { for (int i = t; i < 12; i += 4) slots[i] = lazy.apply(robots[i]); });
      for (auto& th : pool) th.join();
      pool.clear();
      std::vector<double> got(12);
      for (int t = 0; t < 4; t++) pool.emplace_back([&, t] { for (int i = t; i < 12; i += 4) got[i] = slots[i]->getVelocity(); });
      for (auto& th : pool) th.join();
      for (int i = 0; i < 12; i++) EXPECT(got[i] == direct[i].getVelocity());
      EXPECT(lazy.batchesFlushed() == 3);
    }
    std::printf("lazy batching ok\n");
    // Outcome.getAveragePosture(8) of a population in the same launch; the velocity does not depend on what else is recorded
    const std::vector<Outcome> observed = task.applyAllObserved(robots, 8);
    for (int i = 0; i < 12; i++) {
      EXPECT(observed[i].getVelocity() == direct[i].getVelocity() && observed[i].getAveragePosture(8).size() == 64);
      int cells = 0;
      for (bool b : observed[i].getAveragePosture(8)) cells += b;
      EXPECT(cells >= 20 && cells <= 50); // 10 of the 12 cells of the 4 x 3 bounding box, squared to 4 x 4: about 40 of 64 cells
    }
    bool threw = false;
    try { direct[0].getAveragePosture(8); } catch (const VsrException& e) { threw = e.status == VSR_ESTATE; }
    EXPECT(threw);
    std::printf("average posture ok\n");
  } catch (const VsrException& e) {
    EXPECT(e.status == VSR_ECUDA);
    std::printf("no gpu: %s\n", e.what());
  }
  std::printf("host mirror checks ok\n");
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 2 && std::string(argv[1]) == "check") return check();
  const std::string mode = argc >= 2 ? argv[1] : "";
  if (argc < 11 || (mode != "eval" && mode != "validate") || (mode == "validate" && argc < 13)) {
    std::fprintf(stderr, "usage: %s check | eval shape sensors mapper terrain T transient n seed out.bin\n"
                         "       %s validate shape sensors mapper terrain T transient n seed out.bin transformation validationSeed\n", argv[0], argv[0]);
    return 2;
  }
  try {
    const Robot target = targetRobot(argv[2], argv[3]);
    const auto mapper = buildMapper(argv[4]);
    // eval: the run's shared Random(1) draws the `-rnd` terrain seeds (Starter.java:89-93)
    auto task = buildLocomotionTask(argv[5], std::atof(argv[6]), std::atof(argv[7]), std::make_shared<JavaRandom>(1));
    const int n = std::atoi(argv[8]);
    unsigned long long s = std::strtoull(argv[9], nullptr, 10);
    const size_t G = mapper.exampleFor(target).size();
    const auto build = mapper.buildFor(target);
    std::vector<Robot> robots;
    for (int i = 0; i < n; i++) {
      std::vector<double> genotype(G);
      for (auto& g : genotype) { // Knuth MMIX LCG, top 53 bits -> U(-1, 1): the Python test regenerates the same values
        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
        g = 2.0 * ((double)(s >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
      }
      robots.push_back(build(genotype));
    }
    std::vector<double> v;
    if (mode == "validate") { // Starter.validate (Starter.java:122-135) on every robot of the population
      for (const ValidationOutcome& vo : validate(robots, argv[5], argv[11], std::atoll(argv[12]), std::atof(argv[6]), std::atof(argv[7]))) v.push_back(vo.outcome.getVelocity());
    } else {
      for (const Outcome& o : task->applyAll(robots)) v.push_back(o.getVelocity());
    }
    std::FILE* f = std::fopen(argv[10], "wb");
    if (!f || std::fwrite(v.data(), sizeof(double), v.size(), f) != v.size()) { std::fprintf(stderr, "cannot write %s\n", argv[10]); return 1; }
    std::fclose(f);
    std::printf("genotype %zu robots %d best %.6f\n", G, n, *std::max_element(v.begin(), v.end()));
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
