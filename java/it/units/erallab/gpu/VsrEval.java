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

  /** vsr_body_set_voxels: mask[y*w+x], y = 0 bottom row; sensors flattened as {kind, axes, rotated, norm, nRays, side} ints +
   *  {min, max, window, sigma, param} doubles; sensorOff[nVoxels + 1] = every voxel's slice of them (voxels in grid order). */
  public static native void setBody(long ctx, int w, int h, byte[] mask, int[] sensorInts, double[] sensorDoubles, int[] sensorOff);

  /** vsr_controller_set: kind 0 PhaseSin, 1 CentralizedSensing MLP, 2 homogeneous DistributedSensing MLP, 3 heterogeneous
   *  DistributedSensing (one MLP per voxel, weights concatenated in Grid order), 4 TimeFunctions of per-voxel sinusoids
   *  ((frequency, phase, amplitude) per voxel). */
  public static native void setController(long ctx, int kind, int[] layers, int signals, double freq, double amp, boolean inputActivation);

  public static native int genomeLength(long ctx);

  /** vsr_eval: genomes[n*G] row-major -> results[n*8] = {velocity, distance, time, firstX, lastX, controlEnergy, areaRatioAbs, areaRatioSq}. */
  public static native double[] eval(long ctx, double[] genomes, int n);

  /** vsr_eval_ex: frozen-actuator masks (`broken-<ratio>-<seed>` robots) and the per-observation summaries; null / 0 = not requested.
   *  Returns results[n*8] like eval; fills postureOut[n*pN*pN], footprints{0,1}Out[n*kept], centersOut[n*kept*3]. */
  public static native double[] evalEx(long ctx, double[] genomes, int n, int[] frozenMasks, int postureN, int footprintN0, int footprintN1,
                                       byte[] postureOut, byte[] footprints0Out, byte[] footprints1Out, double[] centersOut, int kept);

  /** {nSteps, kFirst, kLast} of the context's episode: kept = kLast - kFirst + 1 observations. */
  public static native int[] episode(long ctx);

  /** vsr_eval_terrains: robot i on the polyline xs/ys[off[i]..off[i+1]) -- the -rnd terrains of Starter.java:89-93. */
  public static native double[] evalTerrains(long ctx, double[] genomes, int n, double[] terrainXs, double[] terrainYs, int[] terrainOff);

  /** include/vsr_ga.h: the device-resident numGA loop (builder/solver/DoublesStandard.java:41-71); returns the vsr_ga handle. */
  public static native long gaCreate(long ctx, int nPop, double xOverProb, double tournamentRate, int minNTournament, double sigmaMut,
                                     long seed, boolean remap);

  /** The solver ES (builder/solver/SimpleES.java:41-49: nParents = round(nPop * parentsRate), nElites = 1) on the same loop. */
  public static native long esCreate(long ctx, int nPop, int nParents, int nElites, double sigma, long seed, boolean remap);

  /** Warps whose contact pool overflowed in any launch so far; 0 = every fitness value is reliable. */
  public static native long contactOverflows(long ctx);

  public static native void gaInit(long ga);

  public static native void gaStep(long ga);

  /** {fitness, genome[0..G)} of the best individual. */
  public static native double[] gaBest(long ga, int genomeLength);

  /** {fitness[nPop], genomes[nPop][G]} flattened, best first. */
  public static native double[] gaPopulation(long ga, int nPop, int genomeLength);

  public static native long gaEvaluations(long ga);

  public static native void gaDestroy(long ga);

  public static native void destroy(long ctx);
}
