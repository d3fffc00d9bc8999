package it.units.erallab.gpu;

import it.units.erallab.hmsrobots.core.controllers.CentralizedSensing;
import it.units.erallab.hmsrobots.core.controllers.DistributedSensing;
import it.units.erallab.hmsrobots.core.controllers.MultiLayerPerceptron;
import it.units.erallab.hmsrobots.core.controllers.PhaseSin;
import it.units.erallab.hmsrobots.core.objects.BreakableVoxel;
import it.units.erallab.hmsrobots.core.objects.Robot;
import it.units.erallab.hmsrobots.core.objects.Voxel;
import it.units.erallab.hmsrobots.core.sensors.Angle;
import it.units.erallab.hmsrobots.core.sensors.AreaRatio;
import it.units.erallab.hmsrobots.core.sensors.Average;
import it.units.erallab.hmsrobots.core.sensors.CompositeSensor;
import it.units.erallab.hmsrobots.core.sensors.Noisy;
import it.units.erallab.hmsrobots.core.sensors.Sensor;
import it.units.erallab.hmsrobots.core.sensors.SoftNormalization;
import it.units.erallab.hmsrobots.core.sensors.Touch;
import it.units.erallab.hmsrobots.core.sensors.Velocity;
import it.units.erallab.hmsrobots.tasks.locomotion.Outcome;
import it.units.erallab.hmsrobots.util.Grid;

import java.util.ArrayList;
import java.util.List;
import java.util.function.Function;

/**
 * Drop-in for the task built by Starter.buildLocomotionTask (locomotion/Starter.java:83-94): a {@code Function<Robot, Outcome>} -- the
 * exact type of {@code Problem.qualityFunction} (Starter.java:75-77) -- with the constructor arguments of
 * {@code new Locomotion(finalT, terrain, settings)} plus the transient time that Starter applies afterwards through subOutcome (:430-433).
 *
 * NOT COMPILED IN THE BUILD CONTAINER (no JDK, 2dhmsr/JGEA jars absent). Three ways to batch behind JGEA's one-robot-per-call contract,
 * all keep JGEA's API and none needs a solver subclass (JGEA's AbstractPopulationBasedIterativeSolver.map is static [U], so a subclass
 * could not replace `executor.invokeAll` (Starter.java:436) without re-implementing init() and update()):
 *  (1) LAZY (default): apply(robot) only records the robot and returns a {@link GpuOutcome}; the first getter called on any outcome of
 *      the batch -- JGEA sorting the new population -- evaluates the whole batch in ONE launch. One generation = one launch, no threads.
 *  (2) applyAll(List<Robot>): the whole population in one call, for a host that already maps a population at once.
 *  (3) BatchedExecutor (below): an ExecutorService whose invokeAll announces the batch size, for eager (non-lazy) semantics.
 */
public class BatchedLocomotion implements Function<Robot, Outcome>, AutoCloseable {

  static final int POSTURE_N = 8;           // Starter.java:185, NamedFunctions.java:356
  static final int[] FOOTPRINT_N = {4, 3};  // NamedFunctions.java:180, :392

  private final long ctx;
  private final double transientT, finalT;
  private final boolean observe;
  private boolean configured = false;
  private GpuOutcome.Pending open;
  private volatile double lastBatchSeconds = 0;

  /**
   * @param observe also record average posture, footprints and the centre trajectories (the detailed / visual outcome functions and the
   *                numSpeciated features); false = fitness and the basic / energy getters only (fastest)
   */
  public BatchedLocomotion(double finalT, double[][] groundProfile, double transientT, boolean observe) {
    this.ctx = VsrEval.create(finalT, transientT, groundProfile[0], groundProfile[1], 32, new int[0]);
    this.finalT = finalT;
    this.transientT = transientT;
    this.observe = observe;
    this.open = new GpuOutcome.Pending(this);
  }

  /** (1) record the robot, return a lazy outcome. Thread-safe: JGEA calls this from its executor's threads. */
  @Override
  public synchronized Outcome apply(Robot robot) {
    GpuOutcome.Pending batch = open;
    batch.robots.add(robot);
    return new GpuOutcome(batch, batch.robots.size() - 1, transientT, finalT);
  }

  /** (2) the whole population in one call. */
  public List<Outcome> applyAll(List<Robot> robots) {
    List<Outcome> outcomes = new ArrayList<>(robots.size());
    GpuOutcome.Pending batch;
    synchronized (this) { // a private batch: robots recorded by concurrent apply() calls stay in theirs
      batch = new GpuOutcome.Pending(this);
      batch.robots.addAll(robots);
    }
    for (int i = 0; i < robots.size(); i++) {
      outcomes.add(new GpuOutcome(batch, i, transientT, finalT));
    }
    if (!robots.isEmpty()) {
      batch.resolve();
    }
    return outcomes;
  }

  double lastBatchSeconds() {
    return lastBatchSeconds;
  }

  /** Called by Pending.resolve(): close the batch, flatten the robots, one vsr_eval_ex launch, keep the raw arrays. */
  void evaluate(GpuOutcome.Pending batch) {
    List<Robot> robots;
    synchronized (this) {
      if (open == batch) {
        open = new GpuOutcome.Pending(this); // robots applied from now on belong to the next batch
      }
      robots = batch.robots;
      if (!configured && !robots.isEmpty()) {
        RobotFlattener.configure(ctx, robots.get(0)); // vsr_body_set + vsr_controller_set from the first robot
        configured = true;
      }
    }
    int n = robots.size();
    if (n == 0) {
      batch.results = new double[0];
      return;
    }
    synchronized (VsrEval.class) { // a vsr_ctx is not re-entrant
      long t0 = System.nanoTime();
      int g = VsrEval.genomeLength(ctx);
      double[] genomes = new double[n * g];
      int[] frozen = new int[n];
      boolean anyFrozen = false;
      for (int i = 0; i < n; i++) {
        System.arraycopy(RobotFlattener.parameters(robots.get(i)), 0, genomes, i * g, g);
        frozen[i] = RobotFlattener.frozenMask(robots.get(i));
        anyFrozen |= frozen[i] != 0;
      }
      int[] ep = VsrEval.episode(ctx);
      batch.kept = ep[2] - ep[1] + 1;
      if (observe) {
        batch.posture = new byte[n * POSTURE_N * POSTURE_N];
        batch.footprints = new byte[][]{new byte[n * batch.kept], new byte[n * batch.kept]};
        batch.centers = new double[n * batch.kept * 3];
        batch.results = VsrEval.evalEx(ctx, genomes, n, anyFrozen ? frozen : null, POSTURE_N, FOOTPRINT_N[0], FOOTPRINT_N[1],
            batch.posture, batch.footprints[0], batch.footprints[1], batch.centers, batch.kept);
      } else {
        batch.results = VsrEval.evalEx(ctx, genomes, n, anyFrozen ? frozen : null, 0, 0, 0, null, null, null, null, batch.kept);
      }
      lastBatchSeconds = (System.nanoTime() - t0) * 1e-9;
    }
  }

  /**
   * (3) Eager batching for hosts that must not hand out lazy outcomes: wraps the run's ExecutorService; invokeAll(tasks) -- what JGEA's
   * map() calls once per generation (Starter.java:436) -- runs the callables (each records its robot through apply()) and then resolves
   * the batch before returning the futures, so every Individual leaves invokeAll with an evaluated fitness.
   */
  public java.util.concurrent.ExecutorService eagerExecutor(java.util.concurrent.ExecutorService inner) {
    BatchedLocomotion self = this;
    return new java.util.concurrent.AbstractExecutorService() {
      @Override
      public <T> List<java.util.concurrent.Future<T>> invokeAll(java.util.Collection<? extends java.util.concurrent.Callable<T>> tasks)
          throws InterruptedException {
        List<java.util.concurrent.Future<T>> futures = inner.invokeAll(tasks);
        GpuOutcome.Pending batch;
        synchronized (self) {
          batch = self.open;
        }
        if (!batch.robots.isEmpty()) {
          batch.resolve();
        }
        return futures;
      }

      @Override
      public void execute(Runnable command) {
        inner.execute(command);
      }

      @Override
      public void shutdown() {
        inner.shutdown();
      }

      @Override
      public List<Runnable> shutdownNow() {
        return inner.shutdownNow();
      }

      @Override
      public boolean isShutdown() {
        return inner.isShutdown();
      }

      @Override
      public boolean isTerminated() {
        return inner.isTerminated();
      }

      @Override
      public boolean awaitTermination(long timeout, java.util.concurrent.TimeUnit unit) throws InterruptedException {
        return inner.awaitTermination(timeout, unit);
      }
    };
  }

  /** Lidar.Side of a one-sided Lidar from its first ray direction [U]: N (pi/4, 3pi/4) = 0, E (-pi/4, pi/4) = 1, S = 2, W = 3. */
  static int lidarSide(double[] rayDirections) {
    double a = rayDirections[rayDirections.length / 2];
    a = Math.atan2(Math.sin(a), Math.cos(a));
    if (a > Math.PI / 4 && a <= 3 * Math.PI / 4) {
      return 0;
    }
    if (a > -Math.PI / 4 && a <= Math.PI / 4) {
      return 1;
    }
    return a < 0 ? 2 : 3;
  }

  @Override
  public void close() {
    VsrEval.destroy(ctx);
  }

  /** Flattens the 2dhmsr objects into the plain arrays of the C ABI (the only place that touches Grid iteration order). */
  static final class RobotFlattener {
    /**
     * vsr_body_set + vsr_controller_set from the first robot of the run. Written against the 2dhmsr 1.12.1 API as recalled
     * (accessor names marked [U] must be checked against the jar): Grid.getW/getH/get(x, y), Voxel.getSensors(),
     * CompositeSensor.getSensor(), Sensor.getDomains() (DoubleRange min()/max()), Average.getInterval() [U], Noisy.getSigma() [U],
     * Velocity.isRotated()/getAxes() [U], MultiLayerPerceptron.getNeurons() (int[] incl. input and output layer),
     * PhaseSin.getFrequency()/getAmplitude() [U], DistributedSensing.getSignals() [U].
     */
    static void configure(long ctx, Robot robot) {
      Grid<Voxel> body = robot.getVoxels();
      int w = body.getW(), h = body.getH();
      byte[] mask = new byte[w * h];
      Voxel first = null;
      for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
          Voxel v = body.get(x, y);
          mask[y * w + x] = (byte) (v == null ? 0 : 1); // y = 0 is the bottom row, as in Grid
          if (v != null && first == null) {
            first = v;
          }
        }
      }
      if (first == null) {
        throw new IllegalArgumentException("Target robot has no voxels");
      }
      // every voxel's own sensor list (RobotUtils.buildSensorizingFunction: uniform, spinedTouch, px / py / l5 ...), voxels in the
      // kernel's order (y outer, x inner), wrappers peeled from the outside in
      List<Sensor> sensors = new ArrayList<>();
      List<Integer> offsets = new ArrayList<>(List.of(0));
      for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
          if (body.get(x, y) != null) {
            sensors.addAll(body.get(x, y).getSensors());
            offsets.add(sensors.size());
          }
        }
      }
      int[] sensorOff = offsets.stream().mapToInt(Integer::intValue).toArray();
      int[] si = new int[6 * sensors.size()];
      double[] sd = new double[5 * sensors.size()];
      for (int i = 0; i < sensors.size(); i++) {
        Sensor s = sensors.get(i);
        double window = 0, sigma = 0;
        int norm = 0;
        while (s instanceof CompositeSensor cs) {
          if (s instanceof Noisy n) {
            sigma = n.getSigma();
          } else if (s instanceof SoftNormalization) {
            norm = 1;
          } else if (s instanceof it.units.erallab.hmsrobots.core.sensors.Normalization) {
            norm = 2;
          } else if (s instanceof Average a) {
            window = a.getInterval();
          } else {
            throw new IllegalArgumentException("Unsupported sensor wrapper " + s.getClass().getSimpleName());
          }
          s = cs.getSensor();
        }
        int kind, axes = 0, rotated = 0, nRays = 0, side = 0;
        double param = 0;
        if (s instanceof AreaRatio) {
          kind = 0;
        } else if (s instanceof Touch) {
          kind = 1;
        } else if (s instanceof Velocity vel) {
          kind = 2;
          axes = (vel.getAxes().contains(Velocity.Axis.X) ? 1 : 0) | (vel.getAxes().contains(Velocity.Axis.Y) ? 2 : 0);
          rotated = vel.isRotated() ? 1 : 0;
        } else if (s instanceof Angle) {
          kind = 3;
        } else if (s instanceof it.units.erallab.hmsrobots.core.sensors.Constant k) {
          kind = 4;
          param = k.getValues()[0]; // [U]
        } else if (s instanceof it.units.erallab.hmsrobots.core.sensors.Lidar l) {
          kind = 6;
          param = l.getRayLength(); // [U]
          nRays = l.getRayDirections().length; // [U]; one side per Lidar, as RobotUtils builds them
          side = BatchedLocomotion.lidarSide(l.getRayDirections());
        } else {
          // TimeFunction sensors hold an opaque lambda: only RobotUtils' `cpg` (sin(-2 pi t)) is known to the device (kind 5, param -1)
          throw new IllegalArgumentException("Unsupported sensor " + s.getClass().getSimpleName());
        }
        si[6 * i] = kind; si[6 * i + 1] = axes; si[6 * i + 2] = rotated; si[6 * i + 3] = norm; si[6 * i + 4] = nRays; si[6 * i + 5] = side;
        sd[5 * i] = s.getDomains()[0].min(); sd[5 * i + 1] = s.getDomains()[0].max(); sd[5 * i + 2] = window; sd[5 * i + 3] = sigma;
        sd[5 * i + 4] = param;
      }
      VsrEval.setBody(ctx, w, h, mask, si, sd, sensorOff);
      Object c = robot.getController();
      if (c instanceof PhaseSin ps) {
        VsrEval.setController(ctx, 0, new int[0], 0, ps.getFrequency(), ps.getAmplitude(), true);
      } else if (c instanceof CentralizedSensing cs) {
        VsrEval.setController(ctx, 1, ((MultiLayerPerceptron) cs.getFunction()).getNeurons(), 0, 1d, 1d, true);
      } else if (c instanceof DistributedSensing ds) {
        List<MultiLayerPerceptron> fs = ds.getFunctions().values().stream().filter(f -> f != null).map(f -> (MultiLayerPerceptron) f).toList();
        boolean homo = fs.stream().allMatch(f -> java.util.Arrays.equals(f.getParams(), fs.get(0).getParams()));
        VsrEval.setController(ctx, homo ? 2 : 3, fs.get(0).getNeurons(), ds.getSignals(), 1d, 1d, true);
      } else {
        throw new IllegalArgumentException("Unsupported controller " + c.getClass().getSimpleName());
      }
    }

    static double[] parameters(Robot robot) {
      Object c = robot.getController();
      if (c instanceof CentralizedSensing cs) {
        return ((MultiLayerPerceptron) cs.getFunction()).getParams();
      }
      if (c instanceof PhaseSin ps) {
        // the phases of the voxels in the kernel's voxel order: y outer, x inner, y = 0 the bottom row -- an EXPLICIT loop, so that the
        // order does not depend on how Grid.stream() / Grid.iterator() walk the cells (brainPhaseFun fills the non-voxel cells too: they
        // are skipped by the body, builder/robot/BrainPhaseFunction.java:36-43)
        Grid<Voxel> body = robot.getVoxels();
        double[] phases = new double[(int) body.values().stream().filter(v -> v != null).count()];
        int c = 0;
        for (int y = 0; y < body.getH(); y++) {
          for (int x = 0; x < body.getW(); x++) {
            if (body.get(x, y) != null) {
              phases[c++] = ps.getPhases().get(x, y);
            }
          }
        }
        return phases;
      }
      if (c instanceof DistributedSensing ds) {
        // kind 2 (brainHomoDist): the shared weights once; kind 3 (brainHeteroDist): every voxel's weights, Grid iteration order
        // (== FunctionsGrid's slicing order, builder/misc/FunctionsGrid.java:31-46)
        List<double[]> ws = new ArrayList<>();
        for (int y = 0; y < ds.getFunctions().getH(); y++) { // explicit y-outer / x-inner order, as for the phases above
          for (int x = 0; x < ds.getFunctions().getW(); x++) {
            if (ds.getFunctions().get(x, y) != null) {
              ws.add(((MultiLayerPerceptron) ds.getFunctions().get(x, y)).getParams());
            }
          }
        }
        boolean homo = ws.stream().allMatch(w -> java.util.Arrays.equals(w, ws.get(0)));
        return homo ? ws.get(0) : ws.stream().flatMapToDouble(java.util.Arrays::stream).toArray();
      }
      throw new IllegalArgumentException("Unsupported controller " + c.getClass().getSimpleName());
    }

    /**
     * bit v (voxel order: y outer, x inner) = voxel v is a BreakableVoxel, i.e. what `broken-<ratio>-<seed>` of
     * RobotUtils.buildRobotTransformation put there (Starter.java:126,428): ACTUATOR FROZEN from t = 0 [U]. Other BreakableVoxel
     * configurations (`breakable-...`: malfunctions triggered during the episode) are outside the device build and are rejected.
     */
    static int frozenMask(Robot robot) {
      Grid<Voxel> body = robot.getVoxels();
      int mask = 0, v = 0;
      for (int y = 0; y < body.getH(); y++) {
        for (int x = 0; x < body.getW(); x++) {
          Voxel voxel = body.get(x, y);
          if (voxel == null) {
            continue;
          }
          if (voxel instanceof BreakableVoxel) {
            mask |= 1 << v;
          }
          v++;
        }
      }
      return mask;
    }
  }
}
