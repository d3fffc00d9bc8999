package it.units.erallab.gpu;

import it.units.erallab.hmsrobots.core.controllers.CentralizedSensing;
import it.units.erallab.hmsrobots.core.controllers.DistributedSensing;
import it.units.erallab.hmsrobots.core.controllers.MultiLayerPerceptron;
import it.units.erallab.hmsrobots.core.controllers.PhaseSin;
import it.units.erallab.hmsrobots.core.objects.Robot;

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
    static void configure(long ctx, Robot robot) {
      // body mask in y*w+x order, per-voxel sensor list (uniform), controller structure: see INTEGRATION.md section 3
      throw new UnsupportedOperationException("written against the 2dhmsr 1.12.1 API; compile with the jars to finish");
    }

    static double[] parameters(Robot robot) {
      Object c = robot.getController();
      if (c instanceof CentralizedSensing cs) {
        return ((MultiLayerPerceptron) cs.getFunction()).getParams();
      }
      if (c instanceof PhaseSin) {
        throw new UnsupportedOperationException("phases: Grid values in iteration order, nulls skipped");
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
