import it.units.erallab.hmsrobots.core.controllers.CentralizedSensing;
import it.units.erallab.hmsrobots.core.controllers.Controller;
import it.units.erallab.hmsrobots.core.controllers.DistributedSensing;
import it.units.erallab.hmsrobots.core.controllers.MultiLayerPerceptron;
import it.units.erallab.hmsrobots.core.controllers.PhaseSin;
import it.units.erallab.hmsrobots.core.objects.Robot;
import it.units.erallab.hmsrobots.core.objects.Voxel;
import it.units.erallab.hmsrobots.core.snapshots.VoxelPoly;
import it.units.erallab.hmsrobots.tasks.locomotion.Locomotion;
import it.units.erallab.hmsrobots.tasks.locomotion.Outcome;
import it.units.erallab.hmsrobots.util.Grid;
import it.units.erallab.hmsrobots.util.RobotUtils;
import org.dyn4j.dynamics.Settings;

import java.io.BufferedReader;
import java.io.IOException;
import java.io.PrintWriter;
import java.nio.file.Files;
import java.nio.file.Path;
import java.util.ArrayList;
import java.util.List;
import java.util.Map;

/**
 * THE PARITY PIN (SURVEY.md 8c-4, BASELINE.md section 3): runs the REAL {@code new Locomotion(T, createTerrain(name), new Settings())
 * .apply(robot)} of 2dhmsr 1.12.1 (call site locomotion/Starter.java:87-93) on the genomes of the committed golden fixtures and dumps,
 * per robot, every observation's voxel-polygon vertices -- the quantity tests/golden/*.npz hold for the oracle.
 * tests/tools/import_reference_traces.py turns the CSVs into tests/golden/ref_*.npz and reports how far the oracle (and therefore every
 * [U] constant of DESIGN.md section 2) is from the reference.
 *
 * NOT COMPILED IN THE BUILD CONTAINER (no JDK / jars); bench.py --impl reference compiles and runs it when it finds `java`, `javac` and
 * the jars under baseline/_ref/.  Usage: java -cp <jars>:. DumpLocomotion <tests/golden dir with *.genomes.csv> <out dir>
 *
 * Input: one CSV per config written by tests/golden/make_golden.py (--csv): header line
 *   shape;sensorConfig;mapper;terrain;finalT  then one genome per line (comma-separated doubles).
 * Output <config>.ref.csv: robot,t,voxelIndex(grid order y-outer x-inner),vertexIndex,x,y  for every observation.
 */
public class DumpLocomotion {

  public static void main(String[] args) throws IOException {
    Path in = Path.of(args[0]), out = Path.of(args[1]);
    Files.createDirectories(out);
    try (var files = Files.list(in)) {
      for (Path f : (Iterable<Path>) files.filter(p -> p.toString().endsWith(".genomes.csv"))::iterator) {
        dump(f, out.resolve(f.getFileName().toString().replace(".genomes.csv", ".ref.csv")));
      }
    }
  }

  private static void dump(Path genomes, Path dst) throws IOException {
    try (BufferedReader r = Files.newBufferedReader(genomes); PrintWriter w = new PrintWriter(Files.newBufferedWriter(dst))) {
      String[] h = r.readLine().split(";");
      String shape = h[0], sensorConfig = h[1], mapper = h[2], terrain = h[3];
      double finalT = Double.parseDouble(h[4]);
      Grid<Voxel> body = (Grid<Voxel>) RobotUtils.buildSensorizingFunction(sensorConfig).apply(RobotUtils.buildShape(shape));
      Locomotion locomotion = new Locomotion(finalT, Locomotion.createTerrain(terrain), new Settings());
      w.println("robot,t,voxel,vertex,x,y");
      String line;
      int robotIndex = 0;
      while ((line = r.readLine()) != null) {
        if (line.isBlank()) {
          continue;
        }
        double[] g = java.util.Arrays.stream(line.split(",")).mapToDouble(Double::parseDouble).toArray();
        Robot robot = build(mapper, body, g);
        Outcome outcome = locomotion.apply(robot);
        for (Map.Entry<Double, Outcome.Observation> e : outcome.getObservations().entrySet()) {
          Grid<VoxelPoly> polies = e.getValue().getVoxelPolies(); // [U] Observation.getVoxelPolies()
          int v = 0;
          for (int y = 0; y < polies.getH(); y++) {
            for (int x = 0; x < polies.getW(); x++) {
              VoxelPoly p = polies.get(x, y);
              if (p == null) {
                continue;
              }
              for (int k = 0; k < p.getVertexes().length; k++) { // [U] Poly.getVertexes(): Point2[]
                w.printf("%d,%.17g,%d,%d,%.17g,%.17g%n", robotIndex, e.getKey(), v, k, p.getVertexes()[k].x, p.getVertexes()[k].y);
              }
              v++;
            }
          }
        }
        w.printf("# robot %d velocity %.17g distance %.17g time %.17g%n", robotIndex,
            outcome.subOutcome(1d, finalT).getVelocity(), outcome.subOutcome(1d, finalT).getDistance(), outcome.subOutcome(1d, finalT).getTime());
        robotIndex++;
      }
    }
  }

  /** genome -> Robot exactly as the reference's mappers do (builder/robot/BrainCentralized.java:36-46, BrainPhaseValues.java:38-50,
   *  BrainHomoDistributed.java:60-90), for the three controllers of the BASELINE configs. */
  private static Robot build(String mapper, Grid<Voxel> body, double[] g) {
    Grid<Voxel> b = it.units.erallab.hmsrobots.util.SerializationUtils.clone(body);
    if (mapper.startsWith("brainPhaseVals")) {
      double f = Double.parseDouble(mapper.replaceAll(".*f=([0-9.]+).*", "$1"));
      int[] c = {0};
      Grid<Double> phases = Grid.create(b.getW(), b.getH(), (x, y) -> 0d);
      for (Grid.Entry<Voxel> e : b) { // Grid iteration order, as BrainPhaseValues.java:40-46
        if (e.value() != null) {
          phases.set(e.key().x(), e.key().y(), g[c[0]++]);
        }
      }
      return new Robot(new PhaseSin(f, 1d, phases), b);
    }
    if (mapper.startsWith("brainCentralized")) {
      int nIn = CentralizedSensing.nOfInputs(b), nOut = CentralizedSensing.nOfOutputs(b);
      int[] inner = it.units.erallab.builder.function.MLP.innerNeurons(nIn, nOut, 0.65, 1); // needs VSREvolution on the class path
      MultiLayerPerceptron mlp = new MultiLayerPerceptron(MultiLayerPerceptron.ActivationFunction.TANH, nIn, inner, nOut, g);
      return new Robot(new CentralizedSensing(b, mlp), b);
    }
    if (mapper.startsWith("brainHomoDist")) {
      int s = Integer.parseInt(mapper.replaceAll(".*s=([0-9]+).*", "$1"));
      Voxel first = b.values().stream().filter(v -> v != null).findFirst().orElseThrow();
      int nIn = DistributedSensing.nOfInputs(first, s), nOut = DistributedSensing.nOfOutputs(first, s);
      int[] inner = it.units.erallab.builder.function.MLP.innerNeurons(nIn, nOut, 0.65, 1);
      DistributedSensing ds = new DistributedSensing(b, s);
      for (Grid.Entry<Voxel> e : b) {
        if (e.value() != null) {
          ds.getFunctions().set(e.key().x(), e.key().y(),
              new MultiLayerPerceptron(MultiLayerPerceptron.ActivationFunction.TANH, nIn, inner, nOut, g.clone()));
        }
      }
      return new Robot(ds, b);
    }
    throw new IllegalArgumentException("unsupported mapper " + mapper);
  }
}
