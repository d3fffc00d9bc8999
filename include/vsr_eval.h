/*
 * vsr_eval.h -- C ABI of libvsr_eval.so: the B200 batched evaluator that replaces, for a whole
 * population at once, the per-individual fitness evaluation of VSREvolution
 *
 *     qualityFunction = transformation . Locomotion.apply . subOutcome(transientT, T)
 *     (reference: src/main/java/it/units/erallab/locomotion/Starter.java:83-94 buildLocomotionTask,
 *      :427-435 Problem.qualityFunction, :174 fitness = Outcome::getVelocity)
 *
 * Plain C types only: pointers, sizes, PODs.  No C++ exception crosses this boundary; every call
 * returns 0 (VSR_OK) or a negative vsr_status and vsr_last_error() gives the message.  The caller
 * owns every buffer it passes; the library keeps no pointer after a call returns.  A vsr_ctx is not
 * re-entrant (one vsr_eval in flight per context); create one context per host thread or serialise.
 *
 * There is NO CPU fallback: if no CUDA device is usable, vsr_ctx_create fails with VSR_ECUDA.
 *
 * INTEGRATION.md shows the JNI/Panama stub that binds these entry points from the Java host.
 */
#ifndef VSR_EVAL_H
#define VSR_EVAL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  VSR_OK = 0,
  VSR_EINVAL = -1,   /* bad argument (-> IllegalArgumentException in the Java shim)              */
  VSR_ECUDA = -2,    /* CUDA runtime failure or no device                                        */
  VSR_ENOMEM = -3,   /* host or device allocation failed                                         */
  VSR_EDIM = -4,     /* dimension mismatch (genome length, MLP I/O size; mirrors the mappers'    */
                     /* IllegalArgumentException, e.g. builder/function/MLP.java:56-62)          */
  VSR_EUNSUPPORTED = -5, /* valid in the reference but outside this build (e.g. >32 voxels)      */
  VSR_ESTATE = -6    /* call order (body/controller not set)                                     */
} vsr_status;

/* ---- environment: replaces `new Locomotion(finalT, groundProfile, new Settings())`,
 *      Starter.java:67,87 -------------------------------------------------------------------- */
typedef struct {
  /* dyn4j Settings() (Starter.java:67 PHYSICS_SETTINGS) */
  double dt;
  int32_t vel_iters, pos_iters;
  double gravity_y, baumgarte, linear_tol, angular_tol, max_linear_corr, max_translation, max_rotation,
      restitution_velocity;
  /* 2dhmsr Voxel defaults */
  double side_length, mass_side_ratio, spring_f, spring_d, lin_damping, ang_damping, voxel_mass,
      area_ratio_max_delta, voxel_friction, voxel_restitution;
  int32_t limit_contraction, scaffolding_mask;
  /* 2dhmsr Ground */
  double ground_friction, ground_restitution, ground_thickness;
  /* 2dhmsr Locomotion */
  double place_x_gap, place_y_gap;
  double final_t;      /* Starter.java:142 episodeTime */
  double transient_t;  /* Starter.java:143 episodeTransientTime (subOutcome lower bound) */
  /* terrain polyline == Locomotion.createTerrain(name) (Starter.java:87); copied at create */
  const double* terrain_x;
  const double* terrain_y;
  int32_t n_terrain;
  /* execution */
  int32_t precision;   /* 32 (default, FP32 state) or 64 (FP64 kernel build, for parity studies) */
  int32_t n_devices;   /* 0 = current device only; else devices[0..n_devices) shard the batch by index */
  int32_t devices[8];
} vsr_env_desc;

/* ---- sensors: one entry of the uniform per-voxel list built by
 *      RobotUtils.buildSensorizingFunction("uniform-<s1+s2..>-<sigma>") (Starter.java:404-405) - */
/* CONSTANT: Constant(value) -- the `px` / `py` predefined sensors (the voxel's relative grid position); TIME_SIN:
 * TimeFunction(t -> sin(2 pi param t), -1, 1) -- `cpg` (param = -1); LIDAR: Lidar(rayLength = param, {side: n_rays}) -- `l1`, `l5`, `lf5`:
 * distance to the ground along n_rays rays of one side's sector, rotating with the voxel (the rays see the ground only) [U] */
enum { VSR_SENSOR_AREA_RATIO = 0, VSR_SENSOR_TOUCH = 1, VSR_SENSOR_VELOCITY = 2, VSR_SENSOR_ANGLE = 3, VSR_SENSOR_CONSTANT = 4,
       VSR_SENSOR_TIME_SIN = 5, VSR_SENSOR_LIDAR = 6 };
typedef struct {
  int32_t kind;
  int32_t axes;       /* VELOCITY: bit0 X, bit1 Y */
  int32_t rotated;    /* VELOCITY: voxel frame */
  int32_t soft_norm;  /* 1: SoftNormalization wrapper; 2: Normalization wrapper ((x - min) / (max - min), clipped to [0, 1]) */
  double dom_min, dom_max;
  double avg_window;  /* Average wrapper window [s], 0 = none */
  double noise_sigma; /* Noisy(sensor, sigma, seed 0) wrapper: + nextGaussian()*sigma*|domain| (java.util.Random replayed on the host) */
  double param;       /* CONSTANT: the value; TIME_SIN: frequency; LIDAR: ray length */
  int32_t n_rays;     /* LIDAR: number of rays = readings (1..8) */
  int32_t side;       /* LIDAR: 0 N, 1 E, 2 S, 3 W */
} vsr_sensor_desc;

/* ---- controllers: CentralizedSensing (builder/robot/BrainCentralized.java:42-45), PhaseSin
 *      (builder/robot/BrainPhaseValues.java:47-50), homogeneous DistributedSensing
 *      (builder/robot/BrainHomoDistributed.java:76-87) ----------------------------------------- */
/*      MLP_DISTRIBUTED_HETERO: DistributedSensing with one MLP per voxel (builder/robot/BrainHeteroDistributed.java:76-83);
 *      the genome row is the voxels' weight vectors concatenated in grid order, as builder/misc/FunctionsGrid.java:31-46
 *      slices it (`brainHeteroDist<funsGrid`); every voxel has the same sensors, so layers[] is shared and
 *      G = n_voxels * vsr_mlp_count_weights(layers).
 *      TIME_SIN: TimeFunctions of per-voxel sinusoids t -> amplitude * sin(2 pi frequency t + phase)
 *      (builder/robot/BodyBrainSinusoidal.java:87-119) on the fixed body of the context; the genome row is
 *      (frequency, phase, amplitude) per voxel in grid order, G = 3 * n_voxels (freq / amp arguments ignored). */
typedef enum {
  VSR_CTRL_PHASE_SIN = 0,
  VSR_CTRL_MLP_CENTRALIZED = 1,
  VSR_CTRL_MLP_DISTRIBUTED_HOMO = 2,
  VSR_CTRL_MLP_DISTRIBUTED_HETERO = 3,
  VSR_CTRL_TIME_SIN = 4
} vsr_ctrl_kind;

/* ---- per-robot result: the Outcome getters the host reads (locomotion/NamedFunctions.java:68-70) */
typedef struct {
  double velocity;   /* Outcome.getVelocity()  == the fitness (Starter.java:174)                   */
  double distance;   /* Outcome.getDistance()                                                       */
  double time;       /* Outcome.getTime()                                                           */
  double first_x, last_x;
  int32_t n_steps, k_first, k_last;
  int32_t contact_overflow; /* >0: the per-warp contact pool overflowed (results unreliable)        */
  /* energy terms over the kept observations (k_first, k_last], summed over the voxels; they feed Outcome.getControlPower()
   * = control_energy / time, getAreaRatioPower() = area-ratio energy / time and getCorrectedEfficiency() = distance /
   * (1 + control_energy) (locomotion/NamedFunctions.java:135-137; the 2dhmsr formulas are recalled [U], so both candidate
   * area-ratio sums are returned): sum f^2 dt;  sum |a_k - a_{k-1}|;  sum |a_k^2 - a_{k-1}^2|                        */
  double control_energy, area_ratio_abs, area_ratio_sq;
  double reserved;
} vsr_result;

typedef struct vsr_ctx vsr_ctx;

/* defaults = `new Settings()` + 2dhmsr constants; terrain NULL, final_t 30, transient_t 1 */
int vsr_env_default(vsr_env_desc* env);

int vsr_ctx_create(const vsr_env_desc* env, vsr_ctx** out);
void vsr_ctx_destroy(vsr_ctx* ctx);
const char* vsr_last_error(const vsr_ctx* ctx); /* ctx may be NULL: error of the last failed create */

/* body grid: mask[y*w+x] != 0 where a voxel exists, y = 0 is the bottom row; same sensors on every voxel.
 * Replaces RobotUtils.buildSensorizingFunction(cfg).apply(RobotUtils.buildShape(shape)) (Starter.java:402-406). */
int vsr_body_set(vsr_ctx* ctx, int32_t w, int32_t h, const uint8_t* mask, const vsr_sensor_desc* sensors, int32_t n_sensors);

/* The same with one sensor list PER VOXEL: what RobotUtils.buildSensorizingFunction returns for names other than `uniform-...`
 * (`spinedTouch-<cpg>-<malfunction>-<sigma>`, `spinedTouchSighted-...`) and for the position-dependent predefined sensors (`px`, `py`,
 * `l1`, `l5`).  Voxel v (grid order: y outer, x inner, y = 0 bottom) carries sensors[sensor_off[v] .. sensor_off[v + 1]); at most 8
 * sensors and 12 readings per voxel.  CentralizedSensing concatenates the voxels' readings in grid order (vsr_input_map); DistributedSensing needs the
 * same number of readings on every voxel (one MLP input size), as in the reference (builder/robot/BrainHomoDistributed.java:52-58). */
int vsr_body_set_voxels(vsr_ctx* ctx, int32_t w, int32_t h, const uint8_t* mask, const vsr_sensor_desc* sensors, const int32_t* sensor_off);

/* controller: layers[] = MLP neurons per layer incl. input and output (ignored for PHASE_SIN / TIME_SIN) */
int vsr_controller_set(vsr_ctx* ctx, vsr_ctrl_kind kind, const int32_t* layers, int32_t n_layers, int32_t signals,
                       double freq, double amp, int32_t mlp_input_activation);

int vsr_genome_len(const vsr_ctx* ctx); /* >=0, or negative status */
int vsr_n_voxels(const vsr_ctx* ctx);

/*
 * Evaluate n robots.  genomes: [n][G] row-major doubles in HOST memory (G = vsr_genome_len).
 * fitness_out: [n] doubles (velocity).  results_out: [n] or NULL.
 * Includes the H2D upload of the genomes and the D2H read of the results.
 */
int vsr_eval(vsr_ctx* ctx, const double* genomes, int32_t n, double* fitness_out, vsr_result* results_out);

/*
 * Device-resident variant on the context's first device: d_genomes / d_fitness / d_results are DEVICE
 * pointers (d_results may be NULL).  Enqueues on `stream` (a cudaStream_t, NULL = default stream) and
 * does not synchronise.
 */
int vsr_eval_device(vsr_ctx* ctx, const double* d_genomes, int32_t n, double* d_fitness, vsr_result* d_results, void* stream);

/*
 * Parity trace of ONE robot for the first n_steps steps (all outputs optional, HOST memory):
 *  vertices_out [n_steps][n_voxels][4][2], bodies_out [n_steps][n_voxels*4][6] (x,y,angle,vx,vy,w),
 *  ctrl_out [n_steps][n_voxels].
 */
int vsr_trace(vsr_ctx* ctx, const double* genome, int32_t n_steps, double* vertices_out, double* bodies_out,
              double* ctrl_out, vsr_result* result_out);

/*
 * Evaluate n robots, each on its OWN terrain: the `-rnd` terrains of Starter.buildLocomotionTask (Starter.java:89-93), where
 * every evaluation builds `Locomotion.createTerrain(name.replace("-rnd", "-" + random.nextInt(10000)))`.  Robot i runs on the
 * polyline terrain_x/terrain_y[terrain_off[i] .. terrain_off[i+1]) (HOST arrays; terrain_off has n+1 entries, starts at 0);
 * its initial placement is recomputed for that polyline.  The context's own terrain only fixes the reference placement.
 */
int vsr_eval_terrains(vsr_ctx* ctx, const double* genomes, int32_t n, const double* terrain_x, const double* terrain_y,
                      const int32_t* terrain_off, double* fitness_out, vsr_result* results_out);

/*
 * The general batched call: vsr_eval plus, each optional,
 *   - per-robot terrains (as vsr_eval_terrains);
 *   - the robot transformation `broken-<ratio>-<seed>` of RobotUtils.buildRobotTransformation(transformationName, random)
 *     (Starter.java:126,428; `identity` is the default, :153): frozen_masks[i] bit v = voxel v (grid order) of robot i was replaced by a
 *     BreakableVoxel whose ACTUATOR is FROZEN from t = 0 -- it applies the last force before freezing, i.e. none; sensors keep working.
 *     Which voxels break is decided on the host (`random.nextDouble() > ratio ? v : broken`, one nextInt() consumed per broken voxel):
 *     vsr_broken_masks replays it;
 *   - the per-observation summaries behind the Outcome getters the reference logs: Outcome.getAveragePosture(n) (Starter.java:180-190:
 *     the feature vector `numSpeciated` clusters on; NamedFunctions.java:356), Outcome.getFootprintsSpectra(n, ...)
 *     (NamedFunctions.java:179-180 n = 4, :392 n = 3) and the centre trajectory with the angle of the central voxel
 *     (getCenterXVelocitySpectrum / getCenterYVelocitySpectrum / getCenterAngleSpectrum, NamedFunctions.java:139-172), for the whole
 *     population in one launch.  K = k_last - k_first + 1 kept observations (vsr_episode_steps).  The 2dhmsr BehaviorUtils
 *     definitions are recalled [U] and restated in oracle/vsr_oracle.c (vo_extras).
 * All pointers are HOST memory, caller-owned; NULL / 0 = not requested.  Observation requests run the observation build of the
 * episode kernel (8 warps per CTA at 255 registers); fitness values are those of vsr_eval.
 */
typedef struct {
  const uint32_t* frozen_masks;     /* [n] or NULL */
  const double* terrain_x;          /* per-robot terrains: all three or none */
  const double* terrain_y;
  const int32_t* terrain_off;       /* [n + 1] */
  int32_t posture_n;                /* 0, or grid side 1..8 */
  int32_t footprint_n[2];           /* 0, or segments 1..8 (two requests: the reference asks for 3 and for 4) */
  uint8_t* posture_out;             /* [n][posture_n * posture_n], index y * n + x, y = 0 bottom row, 1 = cell set on average */
  uint8_t* footprints_out[2];       /* [n][K], bit s = segment s touches the ground at that observation */
  double* centers_out;              /* [n][K][3] = robot centre x, y, angle of the central voxel */
} vsr_eval_opts;
int vsr_eval_ex(vsr_ctx* ctx, const double* genomes, int32_t n, const vsr_eval_opts* opts, double* fitness_out, vsr_result* results_out);
/*
 * Stage the per-robot inputs of a vsr_eval_opts record -- frozen masks and / or per-robot terrains, no observation requests -- for the
 * device-resident paths -- vsr_eval_device and the generation loop of vsr_ga.h: entry i applies to robot i of every following batch
 * (a batch may evaluate fewer than n robots, never more) until the next call; n = 0 or opts = NULL clears.  The data is copied.
 * This is how `-rnd` terrains (Starter.java:89-93) and robot transformations (Starter.java:427-429) reach the generation loop: the
 * host draws the seeds / masks of a generation, stages them, then runs the generation.
 */
int vsr_stage_eval_opts(vsr_ctx* ctx, int32_t n, const vsr_eval_opts* opts);
/* `broken-<ratio>-<seed>` replayed for n robots of this context's body, in robot order, from ONE java.util.Random(seed) shared by
 * the run (RobotUtils.buildRobotTransformation creates it once; with seed `rnd` the caller passes the run's shared seed and the
 * number of values already drawn is *not* tracked here): per voxel in grid order nextDouble() > ratio keeps the voxel, else the voxel
 * breaks and one nextInt() (the BreakableVoxel's own seed) is consumed.  masks_out [n]. */
int vsr_broken_masks(const vsr_ctx* ctx, double ratio, int64_t seed, int32_t n, uint32_t* masks_out);
/* the same on a running stream: *lcg_state is the raw 48-bit state of the shared java.util.Random -- (seed ^ 0x5DEECE66D) & (2^48 - 1)
 * right after `new Random(seed)` -- and is advanced, so consecutive batches (generations) continue where the last one stopped */
int vsr_broken_masks_stream(const vsr_ctx* ctx, double ratio, uint64_t* lcg_state, int32_t n, uint32_t* masks_out);
/* RobotUtils.buildRobotTransformation(name, random): parses `identity` (kind 0) and `broken-<ratio>-<seed|rnd>` (kind 1);
 * `breakable-...` (malfunctions triggered during the episode) returns VSR_EUNSUPPORTED, anything else VSR_EINVAL */
int vsr_robot_transformation(const char* name, int32_t* kind, double* ratio, int64_t* seed, int32_t* seed_is_rnd);
/* the episode of this context: simulated steps and the 1-based first / last kept observation (K = k_last - k_first + 1) */
int vsr_ctx_episode(const vsr_ctx* ctx, int32_t* n_steps, int32_t* k_first, int32_t* k_last);
/* index (voxel order) of BehaviorUtils.getCentralElement of the body: the voxel whose angle centers_out[..][2] records */
int vsr_central_voxel(const vsr_ctx* ctx);

/* timing / accounting of the most recent vsr_eval / vsr_eval_device on device 0 of the context */
/*
 * vsr_eval plus the robot-centre trajectory: centers_out [n][K][2] doubles = (x, y) of the robot centre (mean of the voxel
 * polygon centres, the point Outcome.getDistance / getVelocity measure) at each of the K = k_last - k_first + 1 kept observations
 * (vsr_episode_steps gives k_first / k_last).  These series are the inputs of Outcome.getCenterXVelocitySpectrum /
 * getCenterYVelocitySpectrum / getCenterXPositionSpectrum (locomotion/NamedFunctions.java:139-162), for the whole population in one
 * launch; the full per-step polygons of ONE robot (posture, footprints, video: NamedFunctions.java:163-189, VideoMaker.java:129-133)
 * come from vsr_trace.  Costs two extra warp reductions per kept step, only in this entry point.
 */
int vsr_eval_centers(vsr_ctx* ctx, const double* genomes, int32_t n, double* fitness_out, vsr_result* results_out, double* centers_out);

double vsr_last_kernel_ms(vsr_ctx* ctx);    /* CUDA-event duration of the episode kernel; syncs the stream */
int64_t vsr_kernel_launches(const vsr_ctx* ctx); /* total episode-kernel launches so far */
/* warps whose contact pool overflowed, over every launch of this context so far (0 = every fitness value is reliable; also counts the
 * fitness-only and generation-loop launches, which have no vsr_result to carry contact_overflow); synchronises the streams */
int64_t vsr_contact_overflows(vsr_ctx* ctx);
/* launch shape of the most recent episode-kernel launch (before any launch: the full-wave default).  Batches that need at most 8
 * robot-warps per SM run the 255-register build spread over every SM, larger ones the 10-warp build; results are bit-identical. */
int vsr_kernel_info(const vsr_ctx* ctx, int32_t* regs_per_thread, int32_t* smem_bytes, int32_t* threads_per_cta,
                    int32_t* robots_per_warp, int32_t* ctas_per_sm);

/* ---- integer / naming layer (bit-exact parity surface) -------------------------- */
/* RobotUtils.buildShape(name): writes w*h bytes, returns 0 and sets *w,*h; cap = capacity of mask */
int vsr_shape_mask(const char* name, int32_t* w, int32_t* h, uint8_t* mask, int32_t cap);
/* Locomotion.createTerrain(name): returns number of points (or negative); writes up to cap */
int vsr_terrain(const char* name, double* xs, double* ys, int32_t cap);
/* RobotUtils.buildSensorizingFunction("uniform-a+t-0"): returns number of sensors per voxel */
int vsr_sensor_config(const char* name, vsr_sensor_desc* out, int32_t cap);
/* RobotUtils.buildSensorizingFunction(name).apply(body) with per-voxel lists [U]: `uniform-<s1+s2..>-<sigma>` over the predefined
 * sensors a, t, r, vx, vy, vxy, px, py, cpg, l1, l5, lf5; `spinedTouch-<t|f>-<t|f>-<sigma>`; `spinedTouchSighted-<t|f>-<t|f>-<sigma>`;
 * `empty`.  Writes the concatenated lists (voxels in grid order) and off_out[n_voxels + 1]; returns the total number of sensors.
 * The predefined sensors ax / ay / axy (Trend) and m (Malfunction) are outside this build: VSR_EUNSUPPORTED. */
int vsr_sensor_config_voxels(const char* name, int32_t w, int32_t h, const uint8_t* mask, vsr_sensor_desc* out, int32_t cap, int32_t* off_out);
/* builder/function/MLP.java:28-42 innerNeurons */
int vsr_mlp_inner_neurons(int32_t n_in, int32_t n_out, double ratio, int32_t n_inner, int32_t* out);
/* MultiLayerPerceptron.countWeights(neurons) */
int vsr_mlp_count_weights(const int32_t* neurons, int32_t n_layers);
/* flat genome index of weight (layer l, to-neuron j, from k; k = 0 is the bias) */
int vsr_mlp_weight_index(const int32_t* neurons, int32_t n_layers, int32_t l, int32_t j, int32_t k);
/* Locomotion.apply's `while (t < finalT) t += dt` and subOutcome's kept key range, in IEEE double */
int vsr_episode_steps(double final_t, double transient_t, double dt, int32_t* n_steps, int32_t* k_first, int32_t* k_last);
/* CentralizedSensing input order: for input i -> cell index (y*w+x), sensor index, reading index */
int vsr_input_map(const vsr_ctx* ctx, int32_t* cell, int32_t* sensor, int32_t* reading, int32_t cap);
/* voxel order used for controller outputs / PhaseSin genome: voxel c -> cell index */
int vsr_voxel_cells(const vsr_ctx* ctx, int32_t* cell, int32_t cap);
/* joint counts of the body */
int vsr_counts(const vsr_ctx* ctx, int32_t* n_voxels, int32_t* n_springs, int32_t* n_ropes, int32_t* n_welds);
/* java.util.Random replay (terrain builders): fills out[n] with nextGaussian() of Random(seed) */
int vsr_java_random_gaussians(int64_t seed, double* out, int32_t n);
/* ... and with nextInt(bound): the per-evaluation terrain seeds `random.nextInt(10000)` of Starter.java:91 */
int vsr_java_random_ints(int64_t seed, int32_t bound, int32_t* out, int32_t n);

#ifdef __cplusplus
}
#endif
#endif
