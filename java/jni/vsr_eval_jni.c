/* JNI glue between it.units.erallab.gpu.VsrEval and libvsr_eval.so (include/vsr_eval.h).
 * NOT COMPILED IN THE BUILD CONTAINER (no JDK): type-checked there against a minimal jni.h stand-in
 * (tests/jni_stub/jni.h, tests/test_cabi_cpu.py).  Build on a JDK host:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include vsr_eval_jni.c \
 *       -L../../vsrevolution_b200 -lvsr_eval -o libvsr_eval_jni.so */
#include <jni.h>
#include <stdlib.h>
#include <string.h>

#include "vsr_eval.h"
#include "vsr_ga.h"

static jdoubleArray pack_results(JNIEnv* env, const vsr_result* res, jint n);

static void throw_status(JNIEnv* env, vsr_ctx* ctx, int rc) {
  const char* cls = (rc == VSR_EINVAL || rc == VSR_EDIM) ? "java/lang/IllegalArgumentException" : "java/lang/RuntimeException";
  (*env)->ThrowNew(env, (*env)->FindClass(env, cls), vsr_last_error(ctx));
}

static int throw_iae(JNIEnv* env, const char* msg) {
  (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/IllegalArgumentException"), msg);
  return 1;
}
/* every length that reaches native code is checked against the Java arrays first: a wrong `n` must raise
 * IllegalArgumentException, never read out of bounds */
static int bad_batch(JNIEnv* env, vsr_ctx* ctx, jdoubleArray genomes, jint n) {
  if (!ctx) return throw_iae(env, "null context handle");
  if (!genomes || n < 0) return throw_iae(env, "genomes is null or n is negative");
  const int g = vsr_genome_len(ctx);
  if (g < 0) { throw_status(env, ctx, g); return 1; }
  if ((jlong)(*env)->GetArrayLength(env, genomes) < (jlong)n * (jlong)g) return throw_iae(env, "genomes.length < n * genomeLength");
  return 0;
}

JNIEXPORT jlong JNICALL Java_it_units_erallab_gpu_VsrEval_create(JNIEnv* env, jclass c, jdouble finalT, jdouble transientT,
                                                                 jdoubleArray xs, jdoubleArray ys, jint precision, jintArray devices) {
  vsr_env_desc e;
  vsr_env_default(&e);
  e.final_t = finalT; e.transient_t = transientT; e.precision = precision;
  if (!xs || !ys || !devices) { throw_iae(env, "xs, ys and devices must not be null"); return 0; }
  if ((*env)->GetArrayLength(env, xs) != (*env)->GetArrayLength(env, ys)) { throw_iae(env, "xs.length != ys.length"); return 0; }
  jsize n = (*env)->GetArrayLength(env, xs);
  jdouble* px = (*env)->GetDoubleArrayElements(env, xs, NULL);
  jdouble* py = (*env)->GetDoubleArrayElements(env, ys, NULL);
  e.terrain_x = px; e.terrain_y = py; e.n_terrain = n;
  jsize nd = (*env)->GetArrayLength(env, devices);
  jint* pd = (*env)->GetIntArrayElements(env, devices, NULL);
  e.n_devices = nd > 8 ? 8 : nd;
  for (int i = 0; i < e.n_devices; i++) e.devices[i] = pd[i];
  vsr_ctx* ctx = NULL;
  int rc = vsr_ctx_create(&e, &ctx);
  (*env)->ReleaseIntArrayElements(env, devices, pd, JNI_ABORT);
  (*env)->ReleaseDoubleArrayElements(env, xs, px, JNI_ABORT);
  (*env)->ReleaseDoubleArrayElements(env, ys, py, JNI_ABORT);
  if (rc) { throw_status(env, NULL, rc); return 0; }
  return (jlong)(intptr_t)ctx;
}

/* vsr_body_set_voxels: sensors flattened as 6 ints {kind, axes, rotated, norm, nRays, side} + 5 doubles {min, max, window, sigma, param}
 * each; sensorOff[nVoxels + 1] gives every voxel's slice (voxels in grid order: y outer, x inner) */
JNIEXPORT void JNICALL Java_it_units_erallab_gpu_VsrEval_setBody(JNIEnv* env, jclass c, jlong h, jint w, jint hh, jbyteArray mask,
                                                                 jintArray si, jdoubleArray sd, jintArray sensorOff) {
  vsr_ctx* ctx = (vsr_ctx*)(intptr_t)h;
  if (!ctx || !mask || !si || !sd || !sensorOff) { throw_iae(env, "null argument"); return; }
  if (w <= 0 || hh <= 0 || (jlong)(*env)->GetArrayLength(env, mask) < (jlong)w * hh) { throw_iae(env, "mask.length < w * h"); return; }
  if ((*env)->GetArrayLength(env, si) % 6 || (*env)->GetArrayLength(env, sd) % 5 ||
      (*env)->GetArrayLength(env, sd) / 5 != (*env)->GetArrayLength(env, si) / 6) {
    throw_iae(env, "sensor arrays must hold 6 ints and 5 doubles per sensor"); return;
  }
  jsize ns = (*env)->GetArrayLength(env, si) / 6;
  jbyte* pm = (*env)->GetByteArrayElements(env, mask, NULL);
  jint* pi = (*env)->GetIntArrayElements(env, si, NULL);
  jdouble* pd = (*env)->GetDoubleArrayElements(env, sd, NULL);
  jint* po = (*env)->GetIntArrayElements(env, sensorOff, NULL);
  int nv = 0, bad = 0;
  for (int i = 0; i < w * hh; i++) nv += pm[i] ? 1 : 0;
  if ((*env)->GetArrayLength(env, sensorOff) < nv + 1) bad = 1;
  for (int v = 0; !bad && v <= nv; v++) if (po[v] < 0 || po[v] > ns || (v > 0 && po[v] < po[v - 1])) bad = 1;
  vsr_sensor_desc* s = (vsr_sensor_desc*)calloc(ns ? ns : 1, sizeof(*s));
  for (int i = 0; i < ns; i++) {
    s[i].kind = pi[6 * i]; s[i].axes = pi[6 * i + 1]; s[i].rotated = pi[6 * i + 2]; s[i].soft_norm = pi[6 * i + 3];
    s[i].n_rays = pi[6 * i + 4]; s[i].side = pi[6 * i + 5];
    s[i].dom_min = pd[5 * i]; s[i].dom_max = pd[5 * i + 1]; s[i].avg_window = pd[5 * i + 2]; s[i].noise_sigma = pd[5 * i + 3];
    s[i].param = pd[5 * i + 4];
  }
  int rc = bad ? VSR_EINVAL : vsr_body_set_voxels(ctx, w, hh, (const uint8_t*)pm, s, (const int32_t*)po);
  free(s);
  (*env)->ReleaseIntArrayElements(env, sensorOff, po, JNI_ABORT);
  (*env)->ReleaseDoubleArrayElements(env, sd, pd, JNI_ABORT);
  (*env)->ReleaseIntArrayElements(env, si, pi, JNI_ABORT);
  (*env)->ReleaseByteArrayElements(env, mask, pm, JNI_ABORT);
  if (bad) throw_iae(env, "sensorOff must hold nVoxels + 1 non-decreasing offsets into the sensor arrays");
  else if (rc) throw_status(env, ctx, rc);
}

JNIEXPORT void JNICALL Java_it_units_erallab_gpu_VsrEval_setController(JNIEnv* env, jclass c, jlong h, jint kind, jintArray layers,
                                                                       jint signals, jdouble freq, jdouble amp, jboolean inAct) {
  vsr_ctx* ctx = (vsr_ctx*)(intptr_t)h;
  if (!ctx || !layers) { throw_iae(env, "null argument"); return; }
  jsize nl = (*env)->GetArrayLength(env, layers);
  jint* pl = (*env)->GetIntArrayElements(env, layers, NULL);
  int rc = vsr_controller_set(ctx, (vsr_ctrl_kind)kind, (const int32_t*)pl, nl, signals, freq, amp, inAct ? 1 : 0);
  (*env)->ReleaseIntArrayElements(env, layers, pl, JNI_ABORT);
  if (rc) throw_status(env, ctx, rc);
}

JNIEXPORT jint JNICALL Java_it_units_erallab_gpu_VsrEval_genomeLength(JNIEnv* env, jclass c, jlong h) {
  int g = vsr_genome_len((vsr_ctx*)(intptr_t)h);
  if (g < 0) throw_status(env, (vsr_ctx*)(intptr_t)h, g);
  return g;
}

JNIEXPORT jdoubleArray JNICALL Java_it_units_erallab_gpu_VsrEval_eval(JNIEnv* env, jclass c, jlong h, jdoubleArray genomes, jint n) {
  vsr_ctx* ctx = (vsr_ctx*)(intptr_t)h;
  if (bad_batch(env, ctx, genomes, n)) return NULL;
  jdouble* pg = (*env)->GetDoubleArrayElements(env, genomes, NULL);
  double* fit = (double*)malloc(sizeof(double) * (n ? n : 1));
  vsr_result* res = (vsr_result*)malloc(sizeof(vsr_result) * (n ? n : 1));
  int rc = vsr_eval(ctx, pg, n, fit, res);
  (*env)->ReleaseDoubleArrayElements(env, genomes, pg, JNI_ABORT);
  jdoubleArray out = NULL;
  if (rc) throw_status(env, ctx, rc);
  else out = pack_results(env, res, n);
  free(fit); free(res);
  return out;
}

static jdoubleArray pack_results(JNIEnv* env, const vsr_result* res, jint n) {
  jdoubleArray out = (*env)->NewDoubleArray(env, n * 8);
  for (int i = 0; i < n; i++) {
    jdouble v[8] = {res[i].velocity, res[i].distance, res[i].time, res[i].first_x, res[i].last_x,
                    res[i].control_energy, res[i].area_ratio_abs, res[i].area_ratio_sq};
    (*env)->SetDoubleArrayRegion(env, out, i * 8, 8, v);
  }
  return out;
}

/* vsr_eval_terrains: one polyline per robot (the -rnd terrains of Starter.buildLocomotionTask, Starter.java:89-93) */
JNIEXPORT jdoubleArray JNICALL Java_it_units_erallab_gpu_VsrEval_evalTerrains(JNIEnv* env, jclass c, jlong h, jdoubleArray genomes, jint n,
                                                                              jdoubleArray xs, jdoubleArray ys, jintArray off) {
  (void)c;
  vsr_ctx* ctx = (vsr_ctx*)(intptr_t)h;
  if (bad_batch(env, ctx, genomes, n)) return NULL;
  if (!xs || !ys || !off) { throw_iae(env, "terrain arrays must not be null"); return NULL; }
  if ((*env)->GetArrayLength(env, off) < n + 1) { throw_iae(env, "off.length < n + 1"); return NULL; }
  {
    /* off[] itself is caller data: its last entry bounds the reads of xs / ys */
    jint last = 0;
    jint* po_ = (*env)->GetIntArrayElements(env, off, NULL);
    for (jint i = 0; i <= n; i++) if (po_[i] < 0 || (i > 0 && po_[i] < po_[i - 1])) last = -1;
    if (last == 0) last = po_[n];
    (*env)->ReleaseIntArrayElements(env, off, po_, JNI_ABORT);
    if (last < 0 || (*env)->GetArrayLength(env, xs) < last || (*env)->GetArrayLength(env, ys) < last) {
      throw_iae(env, "off must be non-decreasing and off[n] <= xs.length, ys.length"); return NULL;
    }
  }
  jdouble* pg = (*env)->GetDoubleArrayElements(env, genomes, NULL);
  jdouble* px = (*env)->GetDoubleArrayElements(env, xs, NULL);
  jdouble* py = (*env)->GetDoubleArrayElements(env, ys, NULL);
  jint* po = (*env)->GetIntArrayElements(env, off, NULL);
  double* fit = (double*)malloc(sizeof(double) * (n ? n : 1));
  vsr_result* res = (vsr_result*)malloc(sizeof(vsr_result) * (n ? n : 1));
  int rc = vsr_eval_terrains(ctx, pg, n, px, py, (const int32_t*)po, fit, res);
  (*env)->ReleaseIntArrayElements(env, off, po, JNI_ABORT);
  (*env)->ReleaseDoubleArrayElements(env, ys, py, JNI_ABORT);
  (*env)->ReleaseDoubleArrayElements(env, xs, px, JNI_ABORT);
  (*env)->ReleaseDoubleArrayElements(env, genomes, pg, JNI_ABORT);
  jdoubleArray out = NULL;
  if (rc) throw_status(env, ctx, rc);
  else out = pack_results(env, res, n);
  free(fit); free(res);
  return out;
}

/* vsr_eval_ex: the general batched call -- frozen-actuator masks of `broken-<ratio>-<seed>` robots (Starter.java:126,428), average
 * posture (Starter.java:180-190), footprints (NamedFunctions.java:179-180,392), centre + central-voxel-angle trajectories
 * (NamedFunctions.java:139-172).  Output arrays may be null (= not requested); lengths are checked against n and K. */
JNIEXPORT jdoubleArray JNICALL Java_it_units_erallab_gpu_VsrEval_evalEx(JNIEnv* env, jclass c, jlong h, jdoubleArray genomes, jint n,
                                                                        jintArray frozenMasks, jint postureN, jint fpN0, jint fpN1,
                                                                        jbyteArray postureOut, jbyteArray fp0Out, jbyteArray fp1Out,
                                                                        jdoubleArray centersOut, jint kept) {
  (void)c;
  vsr_ctx* ctx = (vsr_ctx*)(intptr_t)h;
  if (bad_batch(env, ctx, genomes, n)) return NULL;
  if (postureN < 0 || postureN > 8 || fpN0 < 0 || fpN0 > 8 || fpN1 < 0 || fpN1 > 8 || kept < 0) { throw_iae(env, "bad observation request"); return NULL; }
  if (frozenMasks && (*env)->GetArrayLength(env, frozenMasks) < n) { throw_iae(env, "frozenMasks.length < n"); return NULL; }
  if ((postureN > 0) != (postureOut != NULL) || (fpN0 > 0) != (fp0Out != NULL) || (fpN1 > 0) != (fp1Out != NULL)) {
    throw_iae(env, "observation sizes and output arrays go together"); return NULL;
  }
  if (postureOut && (jlong)(*env)->GetArrayLength(env, postureOut) < (jlong)n * postureN * postureN) { throw_iae(env, "postureOut too short"); return NULL; }
  if (fp0Out && (jlong)(*env)->GetArrayLength(env, fp0Out) < (jlong)n * kept) { throw_iae(env, "footprints0Out too short"); return NULL; }
  if (fp1Out && (jlong)(*env)->GetArrayLength(env, fp1Out) < (jlong)n * kept) { throw_iae(env, "footprints1Out too short"); return NULL; }
  if (centersOut && (jlong)(*env)->GetArrayLength(env, centersOut) < (jlong)n * kept * 3) { throw_iae(env, "centersOut too short"); return NULL; }
  {
    /* `kept` must be the library's own K, or the native writes would not match the Java arrays */
    int32_t ns = 0, kf = 0, kl = 0;
    if (vsr_ctx_episode(ctx, &ns, &kf, &kl) || kl - kf + 1 != kept) { throw_iae(env, "kept != k_last - k_first + 1 of this context"); return NULL; }
  }
  vsr_eval_opts o;
  memset(&o, 0, sizeof(o));
  jdouble* pg = (*env)->GetDoubleArrayElements(env, genomes, NULL);
  jint* pf = frozenMasks ? (*env)->GetIntArrayElements(env, frozenMasks, NULL) : NULL;
  jbyte* pp = postureOut ? (*env)->GetByteArrayElements(env, postureOut, NULL) : NULL;
  jbyte* p0 = fp0Out ? (*env)->GetByteArrayElements(env, fp0Out, NULL) : NULL;
  jbyte* p1 = fp1Out ? (*env)->GetByteArrayElements(env, fp1Out, NULL) : NULL;
  jdouble* pc = centersOut ? (*env)->GetDoubleArrayElements(env, centersOut, NULL) : NULL;
  o.frozen_masks = (const uint32_t*)pf;
  o.posture_n = postureN; o.footprint_n[0] = fpN0; o.footprint_n[1] = fpN1;
  o.posture_out = (uint8_t*)pp; o.footprints_out[0] = (uint8_t*)p0; o.footprints_out[1] = (uint8_t*)p1; o.centers_out = pc;
  double* fit = (double*)malloc(sizeof(double) * (n ? n : 1));
  vsr_result* res = (vsr_result*)malloc(sizeof(vsr_result) * (n ? n : 1));
  int rc = vsr_eval_ex(ctx, pg, n, &o, fit, res);
  if (pc) (*env)->ReleaseDoubleArrayElements(env, centersOut, pc, 0); /* mode 0: copy back */
  if (p1) (*env)->ReleaseByteArrayElements(env, fp1Out, p1, 0);
  if (p0) (*env)->ReleaseByteArrayElements(env, fp0Out, p0, 0);
  if (pp) (*env)->ReleaseByteArrayElements(env, postureOut, pp, 0);
  if (pf) (*env)->ReleaseIntArrayElements(env, frozenMasks, pf, JNI_ABORT);
  (*env)->ReleaseDoubleArrayElements(env, genomes, pg, JNI_ABORT);
  jdoubleArray out = NULL;
  if (rc) throw_status(env, ctx, rc);
  else out = pack_results(env, res, n);
  free(fit); free(res);
  return out;
}

/* {n_steps, k_first, k_last} of the context's episode (vsr_episode_steps on its own final_t / transient_t / dt) */
JNIEXPORT jintArray JNICALL Java_it_units_erallab_gpu_VsrEval_episode(JNIEnv* env, jclass c, jlong h) {
  (void)c;
  int32_t v[3] = {0, 0, 0};
  int rc = vsr_ctx_episode((vsr_ctx*)(intptr_t)h, &v[0], &v[1], &v[2]);
  if (rc) { throw_status(env, (vsr_ctx*)(intptr_t)h, rc); return NULL; }
  jintArray out = (*env)->NewIntArray(env, 3);
  (*env)->SetIntArrayRegion(env, out, 0, 3, (const jint*)v);
  return out;
}

/* ---- the device-resident numGA loop (include/vsr_ga.h; DoublesStandard.java:41-71) ---- */
static void throw_ga(JNIEnv* env, vsr_ga* ga, int rc) {
  const char* cls = (rc == VSR_EINVAL || rc == VSR_EDIM) ? "java/lang/IllegalArgumentException" : "java/lang/RuntimeException";
  (*env)->ThrowNew(env, (*env)->FindClass(env, cls), ga ? vsr_ga_last_error(ga) : "vsr_ga_create failed");
}

JNIEXPORT jlong JNICALL Java_it_units_erallab_gpu_VsrEval_gaCreate(JNIEnv* env, jclass c, jlong h, jint nPop, jdouble xOverProb,
                                                                   jdouble tournamentRate, jint minNTournament, jdouble sigmaMut,
                                                                   jlong seed, jboolean remap) {
  (void)c;
  vsr_ga_desc d;
  vsr_ga_default(&d);
  d.n_pop = nPop; d.xover_prob = xOverProb; d.tournament_rate = tournamentRate; d.min_tournament = minNTournament;
  d.sigma_mut = sigmaMut; d.seed = (uint64_t)seed; d.remap = remap ? 1 : 0;
  vsr_ga* ga = NULL;
  int rc = vsr_ga_create((vsr_ctx*)(intptr_t)h, &d, &ga);
  if (rc) { throw_ga(env, NULL, rc); return 0; }
  return (jlong)(intptr_t)ga;
}

/* the solver `ES` = new SimpleES(sigma, parentsRate) (builder/solver/SimpleES.java:41-49) on the same loop */
JNIEXPORT jlong JNICALL Java_it_units_erallab_gpu_VsrEval_esCreate(JNIEnv* env, jclass c, jlong h, jint nPop, jint nParents, jint nElites,
                                                                   jdouble sigma, jlong seed, jboolean remap) {
  (void)c;
  vsr_ga_desc d;
  if (vsr_es_default(&d, nPop)) { throw_ga(env, NULL, VSR_EINVAL); return 0; }
  d.n_parents = nParents; d.n_elites = nElites; d.sigma_mut = sigma; d.seed = (uint64_t)seed; d.remap = remap ? 1 : 0;
  vsr_ga* ga = NULL;
  int rc = vsr_ga_create((vsr_ctx*)(intptr_t)h, &d, &ga);
  if (rc) { throw_ga(env, NULL, rc); return 0; }
  return (jlong)(intptr_t)ga;
}

/* warps whose contact pool overflowed in any launch of the context so far (0 = every fitness value is reliable) */
JNIEXPORT jlong JNICALL Java_it_units_erallab_gpu_VsrEval_contactOverflows(JNIEnv* env, jclass c, jlong h) {
  (void)env; (void)c;
  return (jlong)vsr_contact_overflows((vsr_ctx*)(intptr_t)h);
}

JNIEXPORT void JNICALL Java_it_units_erallab_gpu_VsrEval_gaInit(JNIEnv* env, jclass c, jlong g) {
  (void)c;
  int rc = vsr_ga_init((vsr_ga*)(intptr_t)g);
  if (rc) throw_ga(env, (vsr_ga*)(intptr_t)g, rc);
}

JNIEXPORT void JNICALL Java_it_units_erallab_gpu_VsrEval_gaStep(JNIEnv* env, jclass c, jlong g) {
  (void)c;
  int rc = vsr_ga_step((vsr_ga*)(intptr_t)g);
  if (rc) throw_ga(env, (vsr_ga*)(intptr_t)g, rc);
}

/* best individual: {fitness, genome[0..G)} */
JNIEXPORT jdoubleArray JNICALL Java_it_units_erallab_gpu_VsrEval_gaBest(JNIEnv* env, jclass c, jlong g, jint genomeLengthHint) {
  (void)c; (void)genomeLengthHint;
  if (!g) { throw_iae(env, "null solver handle"); return NULL; }
  const jint genomeLength = vsr_ga_genome_len((vsr_ga*)(intptr_t)g); /* buffer sizes come from the native handle, not from Java */
  if (genomeLength < 0) { throw_ga(env, (vsr_ga*)(intptr_t)g, genomeLength); return NULL; }
  double* buf = (double*)malloc(sizeof(double) * (genomeLength + 1));
  int rc = vsr_ga_best((vsr_ga*)(intptr_t)g, buf + 1, buf);
  jdoubleArray out = NULL;
  if (rc) throw_ga(env, (vsr_ga*)(intptr_t)g, rc);
  else { out = (*env)->NewDoubleArray(env, genomeLength + 1); (*env)->SetDoubleArrayRegion(env, out, 0, genomeLength + 1, buf); }
  free(buf);
  return out;
}

/* whole population, best first: {fitness[nPop], genomes[nPop][G]} flattened */
JNIEXPORT jdoubleArray JNICALL Java_it_units_erallab_gpu_VsrEval_gaPopulation(JNIEnv* env, jclass c, jlong g, jint nPopHint, jint genomeLengthHint) {
  (void)c; (void)nPopHint; (void)genomeLengthHint;
  if (!g) { throw_iae(env, "null solver handle"); return NULL; }
  const jint nPop = vsr_ga_n_pop((vsr_ga*)(intptr_t)g), genomeLength = vsr_ga_genome_len((vsr_ga*)(intptr_t)g);
  if (nPop < 0 || genomeLength < 0) { throw_ga(env, (vsr_ga*)(intptr_t)g, VSR_EINVAL); return NULL; }
  const size_t tot = (size_t)nPop * ((size_t)genomeLength + 1);
  double* buf = (double*)malloc(sizeof(double) * (tot ? tot : 1));
  int rc = vsr_ga_population((vsr_ga*)(intptr_t)g, buf + nPop, buf);
  jdoubleArray out = NULL;
  if (rc) throw_ga(env, (vsr_ga*)(intptr_t)g, rc);
  else { out = (*env)->NewDoubleArray(env, (jsize)tot); (*env)->SetDoubleArrayRegion(env, out, 0, (jsize)tot, buf); }
  free(buf);
  return out;
}

JNIEXPORT jlong JNICALL Java_it_units_erallab_gpu_VsrEval_gaEvaluations(JNIEnv* env, jclass c, jlong g) {
  (void)env; (void)c;
  return (jlong)vsr_ga_evaluations((vsr_ga*)(intptr_t)g);
}

JNIEXPORT void JNICALL Java_it_units_erallab_gpu_VsrEval_gaDestroy(JNIEnv* env, jclass c, jlong g) {
  (void)env; (void)c;
  vsr_ga_destroy((vsr_ga*)(intptr_t)g);
}

JNIEXPORT void JNICALL Java_it_units_erallab_gpu_VsrEval_destroy(JNIEnv* env, jclass c, jlong h) { vsr_ctx_destroy((vsr_ctx*)(intptr_t)h); }
