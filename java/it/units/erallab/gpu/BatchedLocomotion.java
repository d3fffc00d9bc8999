package it.units.erallab.gpu;

import it.units.erallab.hmsrobots.core.controllers.CentralizedSensing;
import it.units.erallab.hmsrobots.core.controllers.DistributedSensing;
import it.units.erallab.hmsrobots.core.controllers.MultiLayerPerceptron;
import it.units.erallab.hmsrobots.core.controllers.PhaseSin;
import it.units.erallab.hmsrobots.core.objects.Robot;

import java.util.ArrayList;
import java.util.List;
import java.util.concurrent.CompletableFuture;
import java.util.function.Function;

/**
 * Drop-in for the task built by Starter.buildLocomotionTask (locomotion/Starter.java:83-94):
 * {@code Function<Robot, GpuOutcome>} with the same constructor arguments as {@code new Locomotion(finalT, terrain, settings)}
 * plus the transient time that Starter applies afterwards through subOutcome (:430-433).
 *
 * NOT COMPILED IN THE BUILD CONTAINER (no JDK, 2dhmsr/JGEA jars absent). Two ways to use it, both keep JGEA's API:
 *  (1) applyAll(List<Robot>) from a thin evolver subclass that maps the whole population at once (preferred);
 *  (2) apply(Robot) from JGEA's executor threads: calls are micro-batched (flush when `expectedBatch` robots are
 *      waiting or after `maxWaitMillis`), each caller blocks on its future.
 */
public class BatchedLocomotion implements Function<Robot, GpuOutcome>, AutoCloseable {

  private final long ctx;
  private final double transientT, finalT;
  private final int expectedBatch;
  private final long maxWaitMillis;
  private boolean configured = false;
  private final List<Robot> pending = new ArrayList<>();
  private final List<CompletableFuture<GpuOutcome>> futures = new ArrayList<>();
  private long firstEnqueued = 0;

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

  @Override
  public GpuOutcome apply(Robot robot) {
    CompletableFuture<GpuOutcome> future = new CompletableFuture<>();
    List<Robot> batch = null;
    List<CompletableFuture<GpuOutcome>> batchFutures = null;
    synchronized (this) {
      if (pending.isEmpty()) {
        firstEnqueued = System.currentTimeMillis();
      }
      pending.add(robot);
      futures.add(future);
      if (pending.size() >= expectedBatch || System.currentTimeMillis() - firstEnqueued >= maxWaitMillis) {
        batch = new ArrayList<>(pending);
        batchFutures = new ArrayList<>(futures);
        pending.clear();
        futures.clear();
      }
    }
    if (batch != null) {
      try {
        List<GpuOutcome> outcomes = applyAll(batch);
        for (int i = 0; i < outcomes.size(); i++) {
          batchFutures.get(i).complete(outcomes.get(i));
        }
      } catch (RuntimeException e) {
        batchFutures.forEach(f -> f.completeExceptionally(e));
      }
    }
    return future.join();
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
        return ((MultiLayerPerceptron) ds.getFunctions().values().stream().filter(f -> f != null).findFirst().orElseThrow()).getParams();
      }
      throw new IllegalArgumentException("Unsupported controller " + c.getClass().getSimpleName());
    }
  }
}
