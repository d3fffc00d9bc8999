/*
 * eval_population.c -- the C ABI of libvsr_eval.so used from plain C, no Python in the process: what a cgo / JNI / Panama binding does.
 *
 *   eval_population <shape> <sensor config> <terrain> <T> <transient> <n> <seed> [out.bin]
 *
 * builds the reference's default problem for that body -- RobotUtils.buildShape + buildSensorizingFunction (locomotion/Starter.java:402-406),
 * brainCentralized<mlp;r=0.65;nIL=1 (builder/robot/BrainCentralized.java:25-57, builder/function/MLP.java:28-89),
 * new Locomotion(T, createTerrain(terrain), new Settings()) with subOutcome(transient, T) (:83-94, :427-435) -- evaluates n genomes
 * ~ U(-1, 1) (a 64-bit LCG, so that the Python test can regenerate them) and prints / writes the fitness vector (Outcome::getVelocity).
 *
 *   gcc -O2 -Iinclude tests/cabi/eval_population.c -Lvsrevolution_b200 -lvsr_eval -Wl,-rpath,$PWD/vsrevolution_b200 -o eval_population
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vsr_eval.h"

#define CHECK(call)                                                                              \
  do {                                                                                           \
    int rc_ = (call);                                                                            \
    if (rc_ < 0) {                                                                               \
      fprintf(stderr, "%s failed: %d (%s)\n", #call, rc_, vsr_last_error(ctx));                  \
      return 1;                                                                                  \
    }                                                                                            \
  } while (0)

static uint64_t lcg_state;
static double lcg_uniform(void) { /* Knuth MMIX LCG, top 53 bits -> [0, 1) */
  lcg_state = lcg_state * 6364136223846793005ULL + 1442695040888963407ULL;
  return (double)(lcg_state >> 11) * (1.0 / 9007199254740992.0);
}

int main(int argc, char** argv) {
  if (argc < 8) {
    fprintf(stderr, "usage: %s shape sensors terrain T transient n seed [out.bin]\n", argv[0]);
    return 2;
  }
  const char *shape = argv[1], *sensors_name = argv[2], *terrain = argv[3];
  const double T = atof(argv[4]), transient = atof(argv[5]);
  const int n = atoi(argv[6]);
  lcg_state = strtoull(argv[7], NULL, 10);
  vsr_ctx* ctx = NULL;

  /* names -> structures (host side of the library, no GPU involved) */
  int32_t w = 0, h = 0;
  CHECK(vsr_shape_mask(shape, &w, &h, NULL, 0));
  uint8_t* mask = (uint8_t*)malloc((size_t)w * h);
  CHECK(vsr_shape_mask(shape, &w, &h, mask, w * h));
  const int n_sens = vsr_sensor_config(sensors_name, NULL, 0);
  if (n_sens < 0) { fprintf(stderr, "bad sensor config %s\n", sensors_name); return 1; }
  vsr_sensor_desc* sens = (vsr_sensor_desc*)calloc(n_sens > 0 ? n_sens : 1, sizeof(vsr_sensor_desc));
  CHECK(vsr_sensor_config(sensors_name, sens, n_sens));
  const int n_t = vsr_terrain(terrain, NULL, NULL, 0);
  if (n_t < 2) { fprintf(stderr, "bad terrain %s\n", terrain); return 1; }
  double *tx = (double*)malloc(sizeof(double) * n_t), *ty = (double*)malloc(sizeof(double) * n_t);
  CHECK(vsr_terrain(terrain, tx, ty, n_t));

  /* new Locomotion(T, terrain, new Settings()) + subOutcome(transient, T) */
  vsr_env_desc env;
  CHECK(vsr_env_default(&env));
  env.final_t = T; env.transient_t = transient;
  env.terrain_x = tx; env.terrain_y = ty; env.n_terrain = n_t;
  CHECK(vsr_ctx_create(&env, &ctx));
  CHECK(vsr_body_set(ctx, w, h, mask, sens, n_sens));

  /* brainCentralized<mlp;r=0.65;nIL=1: inputs = voxels x sensor readings, outputs = voxels */
  const int n_vox = vsr_n_voxels(ctx);
  int dim = 0;
  for (int i = 0; i < n_sens; i++) dim += sens[i].kind == VSR_SENSOR_VELOCITY ? ((sens[i].axes & 1) + ((sens[i].axes >> 1) & 1)) : 1;
  int32_t layers[3] = {n_vox * dim, 0, n_vox};
  CHECK(vsr_mlp_inner_neurons(layers[0], layers[2], 0.65, 1, &layers[1]));
  CHECK(vsr_controller_set(ctx, VSR_CTRL_MLP_CENTRALIZED, layers, 3, 0, 1.0, 1.0, 1));
  const int G = vsr_genome_len(ctx);
  if (G != vsr_mlp_count_weights(layers, 3)) { fprintf(stderr, "genome length mismatch\n"); return 1; }

  double* genomes = (double*)malloc(sizeof(double) * (size_t)n * G);
  for (size_t i = 0; i < (size_t)n * G; i++) genomes[i] = 2.0 * lcg_uniform() - 1.0;
  double* fitness = (double*)malloc(sizeof(double) * (n > 0 ? n : 1));
  vsr_result* results = (vsr_result*)malloc(sizeof(vsr_result) * (n > 0 ? n : 1));
  CHECK(vsr_eval(ctx, genomes, n, fitness, results));

  double best = -1e300, sum = 0;
  for (int i = 0; i < n; i++) { sum += fitness[i]; if (fitness[i] > best) best = fitness[i]; }
  printf("shape %s (%dx%d, %d voxels) mlp %d-%d-%d genome %d robots %d steps %d kept [%d, %d] best %.17g mean %.17g overflows %lld launches %lld\n",
         shape, w, h, n_vox, layers[0], layers[1], layers[2], G, n, n ? results[0].n_steps : 0, n ? results[0].k_first : 0,
         n ? results[0].k_last : 0, best, n ? sum / n : 0.0, (long long)vsr_contact_overflows(ctx), (long long)vsr_kernel_launches(ctx));
  if (argc > 8) {
    FILE* f = fopen(argv[8], "wb");
    if (!f || fwrite(fitness, sizeof(double), n, f) != (size_t)n) { fprintf(stderr, "cannot write %s\n", argv[8]); return 1; }
    fclose(f);
  }
  /* error path: a wrong-sized MLP is VSR_EDIM with the mapper's message, not a crash */
  int32_t bad[3] = {layers[0] + 1, layers[1], layers[2]};
  if (vsr_controller_set(ctx, VSR_CTRL_MLP_CENTRALIZED, bad, 3, 0, 1.0, 1.0, 1) != VSR_EDIM) { fprintf(stderr, "expected VSR_EDIM\n"); return 1; }
  vsr_ctx_destroy(ctx);
  free(mask); free(sens); free(tx); free(ty); free(genomes); free(fitness); free(results);
  return 0;
}
