/* JNI glue between it.units.erallab.gpu.VsrEval and libvsr_eval.so (include/vsr_eval.h).
 * NOT COMPILED IN THE BUILD CONTAINER (no jni.h).  Build on a JDK host:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include vsr_eval_jni.c \
 *       -L../../vsrevolution_b200 -lvsr_eval -o libvsr_eval_jni.so */
#include <jni.h>
#include <stdlib.h>
#include <string.h>

#include "vsr_eval.h"

static void throw_status(JNIEnv* env, vsr_ctx* ctx, int rc) {
  const char* cls = (rc == VSR_EINVAL || rc == VSR_EDIM) ? "java/lang/IllegalArgumentException" : "java/lang/RuntimeException";
  (*env)->ThrowNew(env, (*env)->FindClass(env, cls), vsr_last_error(ctx));
}

JNIEXPORT jlong JNICALL Java_it_units_erallab_gpu_VsrEval_create(JNIEnv* env, jclass c, jdouble finalT, jdouble transientT,
                                                                 jdoubleArray xs, jdoubleArray ys, jint precision, jintArray devices) {
  vsr_env_desc e;
  vsr_env_default(&e);
  e.final_t = finalT; e.transient_t = transientT; e.precision = precision;
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

JNIEXPORT void JNICALL Java_it_units_erallab_gpu_VsrEval_setBody(JNIEnv* env, jclass c, jlong h, jint w, jint hh, jbyteArray mask,
                                                                 jintArray si, jdoubleArray sd) {
  vsr_ctx* ctx = (vsr_ctx*)(intptr_t)h;
  jsize ns = (*env)->GetArrayLength(env, si) / 4;
  jbyte* pm = (*env)->GetByteArrayElements(env, mask, NULL);
  jint* pi = (*env)->GetIntArrayElements(env, si, NULL);
  jdouble* pd = (*env)->GetDoubleArrayElements(env, sd, NULL);
  vsr_sensor_desc* s = (vsr_sensor_desc*)calloc(ns ? ns : 1, sizeof(*s));
  for (int i = 0; i < ns; i++) {
    s[i].kind = pi[4 * i]; s[i].axes = pi[4 * i + 1]; s[i].rotated = pi[4 * i + 2]; s[i].soft_norm = pi[4 * i + 3];
    s[i].dom_min = pd[4 * i]; s[i].dom_max = pd[4 * i + 1]; s[i].avg_window = pd[4 * i + 2]; s[i].noise_sigma = pd[4 * i + 3];
  }
  int rc = vsr_body_set(ctx, w, hh, (const uint8_t*)pm, s, ns);
  free(s);
  (*env)->ReleaseDoubleArrayElements(env, sd, pd, JNI_ABORT);
  (*env)->ReleaseIntArrayElements(env, si, pi, JNI_ABORT);
  (*env)->ReleaseByteArrayElements(env, mask, pm, JNI_ABORT);
  if (rc) throw_status(env, ctx, rc);
}

JNIEXPORT void JNICALL Java_it_units_erallab_gpu_VsrEval_setController(JNIEnv* env, jclass c, jlong h, jint kind, jintArray layers,
                                                                       jint signals, jdouble freq, jdouble amp, jboolean inAct) {
  vsr_ctx* ctx = (vsr_ctx*)(intptr_t)h;
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
  jdouble* pg = (*env)->GetDoubleArrayElements(env, genomes, NULL);
  double* fit = (double*)malloc(sizeof(double) * (n ? n : 1));
  vsr_result* res = (vsr_result*)malloc(sizeof(vsr_result) * (n ? n : 1));
  int rc = vsr_eval(ctx, pg, n, fit, res);
  (*env)->ReleaseDoubleArrayElements(env, genomes, pg, JNI_ABORT);
  jdoubleArray out = NULL;
  if (rc) throw_status(env, ctx, rc);
  else {
    out = (*env)->NewDoubleArray(env, n * 8);
    for (int i = 0; i < n; i++) {
      jdouble v[8] = {res[i].velocity, res[i].distance, res[i].time, res[i].first_x, res[i].last_x,
                      res[i].control_energy, res[i].area_ratio_abs, res[i].area_ratio_sq};
      (*env)->SetDoubleArrayRegion(env, out, i * 8, 8, v);
    }
  }
  free(fit); free(res);
  return out;
}

JNIEXPORT void JNICALL Java_it_units_erallab_gpu_VsrEval_destroy(JNIEnv* env, jclass c, jlong h) { vsr_ctx_destroy((vsr_ctx*)(intptr_t)h); }
