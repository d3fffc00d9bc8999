package it.units.erallab.gpu;

import it.units.erallab.hmsrobots.tasks.locomotion.Outcome;
import it.units.erallab.hmsrobots.util.Grid;

import java.util.ArrayList;
import java.util.List;
import java.util.SortedMap;
import java.util.TreeMap;

/**
 * A 2dhmsr {@link Outcome} whose getters are answered from the device summaries instead of from per-step observations, so that
 * {@code Problem.qualityFunction} ({@code Function<Robot, Outcome>}, locomotion/Starter.java:75-77), the comparator
 * {@code comparing(Outcome::getVelocity)} (:174) and every listener of locomotion/NamedFunctions.java:66-72,130-191,356,392 type-check
 * and read the same values they read from the CPU {@code Locomotion}.
 *
 * LAZY: {@link BatchedLocomotion#apply} returns an instance whose values are not computed yet; the first getter called on ANY outcome
 * of the open batch evaluates the whole batch in one vsr_eval_ex launch. JGEA builds its {@code Individual}s inside
 * {@code executor.invokeAll} (Starter.java:436) without reading the fitness object, and reads it only when the new population is
 * ordered -- so one generation is one launch, with no extra threads and no time-outs.
 *
 * NOT COMPILED IN THE BUILD CONTAINER (no JDK, no 2dhmsr jar). [U] = recalled API that must be checked against the jar:
 * {@code Outcome(Map<Double, Observation>)} is assumed public and the getters non-final (2dhmsr 1.12.1,
 * it.units.erallab.hmsrobots.tasks.locomotion.Outcome). The super constructor gets an empty observation map: nothing here relies on the
 * superclass state.
 */
public class GpuOutcome extends Outcome {

  /** One batch of robots that will be evaluated together; filled by BatchedLocomotion, resolved on first use. */
  static final class Pending {
    final BatchedLocomotion task;
    final List<it.units.erallab.hmsrobots.core.objects.Robot> robots = new ArrayList<>();
    double[] results;      // [n][8] velocity, distance, time, firstX, lastX, controlEnergy, areaRatioAbs, areaRatioSq
    byte[] posture;        // [n][64]
    byte[][] footprints;   // [2][n][K]
    double[] centers;      // [n][K][3]
    int kept;
    RuntimeException error;
    boolean done;

    Pending(BatchedLocomotion task) {
      this.task = task;
    }

    synchronized void resolve() {
      if (!done) {
        try {
          task.evaluate(this); // closes the batch (new applies open the next one) and fills the arrays above
        } catch (RuntimeException e) {
          error = e;
        }
        done = true;
      }
      if (error != null) {
        throw error;
      }
    }
  }

  private final Pending batch;
  private final int slot;
  private final double startT, endT;

  GpuOutcome(Pending batch, int slot, double startT, double endT) {
    super(new TreeMap<>()); // [U] Outcome(Map<Double, Observation>)
    this.batch = batch;
    this.slot = slot;
    this.startT = startT;
    this.endT = endT;
  }

  private double r(int k) {
    batch.resolve();
    return batch.results[slot * 8 + k];
  }

  @Override
  public double getVelocity() {
    return r(0);
  }

  @Override
  public double getDistance() {
    return r(1);
  }

  @Override
  public double getTime() {
    return r(2);
  }

  /** computation.time (NamedFunctions.java:68): the wall-clock share of this robot in its batch's launch. */
  @Override
  public double getComputationTime() {
    batch.resolve();
    return batch.task.lastBatchSeconds() / Math.max(1, batch.robots.size());
  }

  /** NamedFunctions.java:135-137; the 2dhmsr formulas are recalled [U] (see include/vsr_eval.h, vsr_result). */
  @Override
  public double getControlPower() {
    return r(5) / r(2);
  }

  @Override
  public double getAreaRatioPower() {
    return r(6) / r(2);
  }

  @Override
  public double getCorrectedEfficiency() {
    return r(1) / (1d + r(5));
  }

  /** Starter.java:180-190 (numSpeciated features), NamedFunctions.java:356. Grid.create(n, n, (x, y) -> ...) [U]. */
  @Override
  public Grid<Boolean> getAveragePosture(int n) {
    batch.resolve();
    if (n != BatchedLocomotion.POSTURE_N) {
      throw new IllegalArgumentException("the device rasterises the average posture on an " + BatchedLocomotion.POSTURE_N + "-cell grid side");
    }
    return Grid.create(n, n, (x, y) -> batch.posture[slot * n * n + y * n + x] != 0);
  }

  /** Centre trajectory of this robot over the kept observations: column 0 = x, 1 = y, 2 = angle of the central voxel. */
  double[] centerSeries(int column) {
    batch.resolve();
    double[] s = new double[batch.kept];
    for (int k = 0; k < batch.kept; k++) {
      s[k] = batch.centers[(slot * batch.kept + k) * 3 + column];
    }
    return s;
  }

  /** 0/1 series "footprint segment s touches the ground" (n = 4 -> request 0, n = 3 -> request 1). */
  double[] footprintSeries(int n, int segment) {
    batch.resolve();
    int q = n == BatchedLocomotion.FOOTPRINT_N[0] ? 0 : n == BatchedLocomotion.FOOTPRINT_N[1] ? 1 : -1;
    if (q < 0) {
      throw new IllegalArgumentException("footprints are recorded for n = 4 and n = 3 segments");
    }
    double[] s = new double[batch.kept];
    for (int k = 0; k < batch.kept; k++) {
      s[k] = (batch.footprints[q][slot * batch.kept + k] >> segment) & 1;
    }
    return s;
  }

  // The spectra (NamedFunctions.java:139-189,365-396) are BehaviorUtils.computeQuantizedSpectrum of these series [U]:
  // SortedMap<DoubleRange, Double> getCenterXVelocitySpectrum(minF, maxF, nBins) etc.  They are overridden with the same helper the
  // superclass uses, fed with the device series; the signatures depend on 2dhmsr's DoubleRange / BehaviorUtils and are left to the
  // maintainer who compiles against the jar (the Python mirror implements them: vsrevolution_b200/locomotion.py Outcome).

  /** Starter.java:430-433 calls subOutcome(transientT, T): the device already evaluated exactly that window. */
  @Override
  public Outcome subOutcome(double startT, double endT) {
    if (startT != this.startT || endT != this.endT) {
      throw new IllegalArgumentException("the device keeps only the window given at construction");
    }
    return this;
  }
}
