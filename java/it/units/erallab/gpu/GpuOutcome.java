package it.units.erallab.gpu;

/** The Outcome getters the reference reads on the hot path (locomotion/Starter.java:174, NamedFunctions.java:68-70). */
public record GpuOutcome(double velocity, double distance, double time, double controlEnergy, double areaRatioEnergy,
                         double startT, double endT) {
  public double getVelocity() {
    return velocity;
  }

  public double getDistance() {
    return distance;
  }

  public double getTime() {
    return time;
  }

  /** NamedFunctions.java:135-137: control.power, area.ratio.power, corrected.efficiency (2dhmsr formulas recalled, see vsr_eval.h). */
  public double getControlPower() {
    return controlEnergy / time;
  }

  public double getAreaRatioPower() {
    return areaRatioEnergy / time;
  }

  public double getCorrectedEfficiency() {
    return distance / (1d + controlEnergy);
  }

  /** Starter.java:430-433 calls subOutcome(transientT, T): the device already evaluated exactly that window. */
  public GpuOutcome subOutcome(double startT, double endT) {
    if (startT != this.startT || endT != this.endT) {
      throw new IllegalArgumentException("the device keeps only the window given at construction");
    }
    return this;
  }
}
