package it.units.erallab.gpu;

/**
 * JNI binding of libvsr_eval.so (include/vsr_eval.h). NOT COMPILED IN THE BUILD CONTAINER (no JDK): reviewed by hand
 * against the header. One native method per C entry point that the locomotion path needs.
 */
public final class VsrEval {
  static {
    System.loadLibrary("vsr_eval_jni"); // java/jni/vsr_eval_jni.c, links libvsr_eval.so
  }

  private VsrEval() {}

  /** vsr_env_default + vsr_ctx_create; terrain = Locomotion.createTerrain(name) ({xs, ys}). Returns the vsr_ctx handle. */
  public static native long create(double finalT, double transientT, double[] terrainXs, double[] terrainYs, int precision, int[] devices);

  /** vsr_body_set: mask[y*w+x], y = 0 bottom row; sensors flattened as {kind, axes, rotated, softNorm} ints + {min, max, window, sigma} doubles. */
  public static native void setBody(long ctx, int w, int h, byte[] mask, int[] sensorInts, double[] sensorDoubles);

  /** vsr_controller_set: kind 0 PhaseSin, 1 CentralizedSensing MLP, 2 homogeneous DistributedSensing MLP. */
  public static native void setController(long ctx, int kind, int[] layers, int signals, double freq, double amp, boolean inputActivation);

  public static native int genomeLength(long ctx);

  /** vsr_eval: genomes[n*G] row-major -> results[n*5] = {velocity, distance, time, firstX, lastX}. */
  public static native double[] eval(long ctx, double[] genomes, int n);

  public static native void destroy(long ctx);
}
