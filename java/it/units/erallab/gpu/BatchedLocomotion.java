package it.units.erallab.gpu;

import it.units.erallab.hmsrobots.core.controllers.CentralizedSensing;
import it.units.erallab.hmsrobots.core.controllers.DistributedSensing;
import it.units.erallab.hmsrobots.core.controllers.MultiLayerPerceptron;
import it.units.erallab.hmsrobots.core.controllers.PhaseSin;
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
import it.units.erallab.hmsrobots.util.Grid;

import java.util.ArrayList;
import java.util.List;
import java.util.function.Function;

/**
 * Drop-in for the task built by Starter.buildLocomotionTask (locomotion/Starter.java:83-94):
 * {@code Function<Robot, GpuOutcome>} with the same constructor arguments as {@code new Locomotion(finalT, terrain, settings)}
 * plus the transient time that Starter applies afterwards through subOutcome (:430-433).
 *
 * NOT COMPILED IN THE BUILD CONTAINER (no JDK, 2dhmsr/JGEA jars absent). Two ways to use it, both keep JGEA's API:
 *  (1) applyAll(List<Robot>) from a thin evolver subclass that maps the whole population at once (preferred);
 *  (2) apply(Robot) from JGEA's executor threads: calls are micro-batched (flush when `expectedBatch` robots are
 *      waiting, or by the first waiter whose `maxWaitMillis` expires), each caller blocks until its batch is evaluated.
 */
public class BatchedLocomotion implements Function<Robot, GpuOutcome>, AutoCloseable {

  private final long ctx;
  private final double transientT, finalT;
  private final int expectedBatch;
  private final long maxWaitMillis;
  private boolean configured = false;

  public BatchedLocomotion(double finalT, double[][] groundProfile, double transientT, int expectedBatch, long maxWaitMillis) {
    this.ctx = VsrEval.create(finalT, transientT, groundProfile[0], groundProfile[1], 32, new int[0]);
    this.finalT = finalT;
    this.transientT = transientT;
    this.expectedBatch = expectedBatch;
    this.maxWaitMillis = maxWaitMillis;
  }

  /** The whole population in one vsr_eval call. */
  public synchronized List<GpuOutcome> applyAll(List<Robot> robots) {
    if (robots.isEmpty()) {
      return List.of();
    }
    if (!configured) {
      RobotFlattener.configure(ctx, robots.get(0)); // vsr_body_set + vsr_controller_set from the first robot
      configured = true;
    }
    int g = VsrEval.genomeLength(ctx);
    double[] genomes = new double[robots.size() * g];
    for (int i = 0; i < robots.size(); i++) {
      System.arraycopy(RobotFlattener.parameters(robots.get(i)), 0, genomes, i * g, g);
    }
    double[] r = VsrEval.eval(ctx, genomes, robots.size());
    List<GpuOutcome> outcomes = new ArrayList<>(robots.size());
    for (int i = 0; i < robots.size(); i++) {
      outcomes.add(new GpuOutcome(r[i * 8], r[i * 8 + 1], r[i * 8 + 2], r[i * 8 + 5], r[i * 8 + 6], transientT, finalT));
    }
    return outcomes;
  }

  /** One open batch: callers append, the filler (or the first waiter whose wait expires) closes and evaluates it. */
  private static final class Batch {
    final List<Robot> robots = new ArrayList<>();
    final long t0 = System.currentTimeMillis();
    List<GpuOutcome> outcomes;
    RuntimeException error;
    boolean done;
  }

  private Batch open = new Batch();
  private final Object queueLock = new Object();

  private void flush(Batch batch) {
    try {
      batch.outcomes = applyAll(batch.robots); // synchronized: a vsr_ctx is not re-entrant
    } catch (RuntimeException e) {
      batch.error = e;
    }
    synchronized (batch) {
      batch.done = true;
      batch.notifyAll();
    }
  }

  /** Same protocol as vsrevolution_b200/locomotion.py:MicroBatchedLocomotion (exercised there by tests/test_host_mirror_cpu.py). */
  @Override
  public GpuOutcome apply(Robot robot) {
    Batch batch;
    int slot;
    boolean full;
    synchronized (queueLock) {
      batch = open;
      slot = batch.robots.size();
      batch.robots.add(robot);
      full = batch.robots.size() >= expectedBatch;
      if (full) {
        open = new Batch();
      }
    }
    if (full) {
      flush(batch);
    } else {
      boolean leader = false;
      synchronized (batch) {
        while (!batch.done && !leader) {
          long left = batch.t0 + maxWaitMillis - System.currentTimeMillis();
          if (left > 0) {
            try {
              batch.wait(left);
            } catch (InterruptedException e) {
              Thread.currentThread().interrupt();
              throw new RuntimeException(e);
            }
          } else {
            synchronized (queueLock) { // timeout: the first waiter that still finds the batch open closes it
              if (open == batch) {
                open = new Batch();
                leader = true;
              }
            }
            if (!leader) {
              try {
                batch.wait();
              } catch (InterruptedException e) {
                Thread.currentThread().interrupt();
                throw new RuntimeException(e);
              }
            }
          }
        }
      }
      if (leader) {
        flush(batch);
      }
    }
    if (batch.error != null) {
      throw batch.error;
    }
    return batch.outcomes.get(slot);
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
      // the uniform per-voxel sensor list (buildSensorizingFunction("uniform-...")), wrappers peeled from the outside in
      List<Sensor> sensors = first.getSensors();
      int[] si = new int[4 * sensors.size()];
      double[] sd = new double[4 * sensors.size()];
      for (int i = 0; i < sensors.size(); i++) {
        Sensor s = sensors.get(i);
        double window = 0, sigma = 0;
        boolean softNorm = false;
        while (s instanceof CompositeSensor cs) {
          if (s instanceof Noisy n) {
            sigma = n.getSigma();
          } else if (s instanceof SoftNormalization) {
            softNorm = true;
          } else if (s instanceof Average a) {
            window = a.getInterval();
          } else {
            throw new IllegalArgumentException("Unsupported sensor wrapper " + s.getClass().getSimpleName());
          }
          s = cs.getSensor();
        }
        int kind, axes = 0, rotated = 0;
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
        } else {
          throw new IllegalArgumentException("Unsupported sensor " + s.getClass().getSimpleName());
        }
        si[4 * i] = kind; si[4 * i + 1] = axes; si[4 * i + 2] = rotated; si[4 * i + 3] = softNorm ? 1 : 0;
        sd[4 * i] = s.getDomains()[0].min(); sd[4 * i + 1] = s.getDomains()[0].max(); sd[4 * i + 2] = window; sd[4 * i + 3] = sigma;
      }
      VsrEval.setBody(ctx, w, h, mask, si, sd);
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
        // the phases of the voxels in Grid iteration order (brainPhaseFun fills the non-voxel cells too: skip them by the body)
        Grid<Voxel> body = robot.getVoxels();
        return body.stream().filter(e -> e.value() != null).mapToDouble(e -> ps.getPhases().get(e.key().x(), e.key().y())).toArray();
      }
      if (c instanceof DistributedSensing ds) {
        // kind 2 (brainHomoDist): the shared weights once; kind 3 (brainHeteroDist): every voxel's weights, Grid iteration order
        // (== FunctionsGrid's slicing order, builder/misc/FunctionsGrid.java:31-46)
        List<double[]> ws = ds.getFunctions().values().stream().filter(f -> f != null)
            .map(f -> ((MultiLayerPerceptron) f).getParams()).toList();
        boolean homo = ws.stream().allMatch(w -> java.util.Arrays.equals(w, ws.get(0)));
        return homo ? ws.get(0) : ws.stream().flatMapToDouble(java.util.Arrays::stream).toArray();
      }
      throw new IllegalArgumentException("Unsupported controller " + c.getClass().getSimpleName());
    }
  }
}
