/*
 * vsr_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Double-precision, single-robot, strictly sequential restatement of the path
 *   Locomotion.apply(robot) -> Outcome            (2dhmsr 1.12.1, un-vendored)
 *   World.step(1)                                 (dyn4j 3.4.x, un-vendored)
 * as called from /root/reference/src/main/java/it/units/erallab/locomotion/Starter.java:83-94,427-435.
 *
 * PARITY UNPINNED: neither dependency's source, nor a JVM, nor any golden vector exists in
 * /root/reference.  Every constant and formula here is restated from the
 * published algorithm of those libraries and is tagged [U] in docs/PHYSICS_SPEC.md.  The oracle is
 * pinned only by closed-form known-answer tests (tests/test_oracle_known_answers.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product (libvsr_eval.so) never links or calls it.
 */
#ifndef VSR_ORACLE_H
#define VSR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Physics + task parameters.  Defaults (vo_default_params) restate dyn4j Settings() and
 * 2dhmsr Voxel/Ground/Locomotion constants; see docs/PHYSICS_SPEC.md for the provenance of each. */
typedef struct {
  /* dyn4j Settings */
  double dt;                   /* step frequency, 1/60 s                                   */
  int    vel_iters;            /* velocity constraint solver iterations, 10                */
  int    pos_iters;            /* position constraint solver iterations, 10                */
  double gravity_y;            /* World.EARTH_GRAVITY = (0,-9.8)                            */
  double baumgarte;            /* 0.2                                                      */
  double linear_tol;           /* 0.005                                                    */
  double angular_tol;          /* 2 degrees                                                */
  double max_linear_corr;      /* 0.2                                                      */
  double max_translation;      /* 2.0 per step                                             */
  double max_rotation;         /* pi/2 per step                                            */
  double restitution_velocity; /* 1.0                                                      */
  /* 2dhmsr Voxel */
  double side_length;          /* 3                                                        */
  double mass_side_ratio;      /* 0.3                                                      */
  double spring_f;             /* 8 Hz                                                     */
  double spring_d;             /* 0.3                                                      */
  double lin_damping;          /* 1                                                        */
  double ang_damping;          /* 1                                                        */
  double voxel_mass;           /* 1 (0.25 per corner mass)                                 */
  double area_ratio_max_delta; /* 0.225                                                    */
  double voxel_friction;       /* 100                                                      */
  double voxel_restitution;    /* 0.1                                                      */
  int    limit_contraction;    /* 1: four lower-limit rope joints per voxel                */
  int    scaffolding_mask;     /* bit0 SIDE_INTERNAL bit1 SIDE_EXTERNAL bit2 SIDE_CROSS bit3 CENTRAL_CROSS */
  /* 2dhmsr Ground */
  double ground_friction;      /* dyn4j fixture default 0.2                                */
  double ground_restitution;   /* dyn4j fixture default 0.0                                */
  double ground_thickness;     /* 50 below min(ys)                                         */
  /* 2dhmsr Locomotion */
  double place_x_gap;          /* 1                                                        */
  double place_y_gap;          /* 1                                                        */
  double final_t;              /* episode time                                             */
  double transient_t;          /* Starter.java:143 episodeTransientTime                    */
} vo_params;

/* CONSTANT: Constant(value) (the `px` / `py` entries of RobotUtils' predefined sensors: the voxel's relative grid position);
 * TIME_SIN: TimeFunction(t -> sin(2 pi param t), -1, 1) (the `cpg` entry, param = -1); LIDAR: Lidar(rayLength = param, {side: n_rays})
 * (the `l1` / `l5` / `lf5` entries): distance to the ground along n_rays rays of one side's sector, rotated with the voxel [U] */
enum { VO_SENSOR_AREA_RATIO = 0, VO_SENSOR_TOUCH = 1, VO_SENSOR_VELOCITY = 2, VO_SENSOR_ANGLE = 3, VO_SENSOR_CONSTANT = 4,
       VO_SENSOR_TIME_SIN = 5, VO_SENSOR_LIDAR = 6 };

/* One sensor of a voxel's sensor list: Noisy(SoftNormalization | Normalization(Average(base))) with each wrapper optional. */
typedef struct {
  int    kind;
  int    axes;        /* VELOCITY: bit0 = X, bit1 = Y (readings emitted in that order) */
  int    rotated;     /* VELOCITY: express in the voxel frame */
  int    soft_norm;   /* 1: wrap in SoftNormalization; 2: wrap in Normalization ((x - min) / (max - min) clipped to [0, 1]) */
  double dom_min, dom_max;
  double avg_window;  /* >0: wrap in Average(window seconds) */
  double noise_sigma; /* >0: Noisy(sensor, sigma, seed 0): + nextGaussian()*sigma*|domain| per reading per step */
  double param;       /* CONSTANT: the value; TIME_SIN: frequency; LIDAR: ray length */
  int    n_rays;      /* LIDAR: rays (= readings) */
  int    side;        /* LIDAR: 0 N, 1 E, 2 S, 3 W */
} vo_sensor;

/* DISTRIBUTED_HETERO: one MLP per voxel, genome = the voxels' weight vectors concatenated in grid order
 * (brainHeteroDist<funsGrid: builder/robot/BrainHeteroDistributed.java:76-83, builder/misc/FunctionsGrid.java:31-46).
 * TIME_SIN: TimeFunctions of per-voxel sinusoids, genome = (frequency, phase, amplitude) per voxel in grid order
 * (builder/robot/BodyBrainSinusoidal.java:87-119: t -> amplitude * sin(2 pi freq t + phase)). */
enum { VO_CTRL_PHASE_SIN = 0, VO_CTRL_MLP_CENTRALIZED = 1, VO_CTRL_MLP_DISTRIBUTED_HOMO = 2,
       VO_CTRL_MLP_DISTRIBUTED_HETERO = 3, VO_CTRL_TIME_SIN = 4 };

typedef struct {
  int    kind;
  double freq, amp;        /* PHASE_SIN */
  int    n_layers;         /* MLP: number of entries of layers[] (input, hidden..., output) */
  int    layers[8];
  int    signals;          /* DISTRIBUTED: signals per side */
  int    input_activation; /* MLP: apply tanh to the input layer (2dhmsr quirk, [U]) */
} vo_ctrl;

/* Per-episode result (what Outcome getters used by the host need). */
typedef struct {
  double velocity;     /* Outcome.getVelocity() over [transient_t, final_t]      */
  double distance;     /* Outcome.getDistance()                                  */
  double time;         /* Outcome.getTime()                                      */
  double first_x, last_x;
  int    n_steps;      /* steps simulated                                        */
  int    k_first, k_last; /* 1-based step indices of first/last kept observation  */
  /* operation tallies (totals over the whole episode) for the roofline's algorithmic work */
  int64_t n_spring_vel, n_rope_vel, n_weld_vel, n_contact_pt_vel, n_block_vel;
  int64_t n_rope_pos, n_weld_pos, n_contact_pt_pos;
  int64_t n_detect_pairs, n_pos_iters, n_manifolds;
  /* energy terms over the kept observations (k_first, k_last], summed over the voxels [U] (2dhmsr Voxel.getControlEnergy /
   * getAreaRatioEnergy feed Outcome.getControlPower / getAreaRatioPower / getCorrectedEfficiency, NamedFunctions.java:135-137):
   *   control_energy = sum f^2 dt;  area_ratio_abs = sum |a_k - a_{k-1}|;  area_ratio_sq = sum |a_k^2 - a_{k-1}^2|        */
  double control_energy, area_ratio_abs, area_ratio_sq;
} vo_result;

/* Optional extras of one episode (vo_run_ex): the robot transformation `broken-<ratio>-<seed>` and the per-observation
 * summaries behind the Outcome getters the host logs ([U]: 2dhmsr RobotUtils.buildRobotTransformation, BehaviorUtils):
 *   frozen_mask   bit v = voxel v is a BreakableVoxel whose ACTUATOR is FROZEN from t = 0 (trigger TIME >= 0, restore time
 *                 infinite): applyForce(f) applies the last force before freezing, i.e. 0 -- the voxel is passive, its sensors work
 *   posture       Outcome.getAveragePosture(n) (Starter.java:180-190, NamedFunctions.java:356): per kept observation the voxel
 *                 polygons are rasterised on the n x n grid of the squared bounding box (cell centre inside any polygon);
 *                 average posture = cells set in MORE than half of the observations; [y*n + x], y = 0 bottom
 *   footprints    BehaviorUtils.computeFootprint(polies, n) behind Outcome.getFootprintsSpectra(n, ...) (NamedFunctions.java:179-180,
 *                 392): x-range of every ground-touching voxel polygon mapped on n segments of the robot's x-extent with
 *                 round((x - min) / (max - min) * (n - 1)); one byte per kept observation, bit s = segment s
 *   centers       per kept observation: robot centre x, y and the angle of the central voxel (getCentralElement: the voxel closest to
 *                 the mean grid position) -- inputs of getCenterX/YVelocitySpectrum and getCenterAngleSpectrum */
typedef struct {
  uint32_t frozen_mask;
  int      posture_n;        /* 0 = off, else 1..8 */
  uint8_t* posture_out;      /* [posture_n * posture_n] */
  int32_t* posture_counts;   /* optional [posture_n * posture_n]: observations in which the cell was set */
  int      footprint_n[2];   /* 0 = off, else 1..8 */
  uint8_t* footprints_out[2]; /* [kept observations] each */
  double*  centers_out;      /* optional [kept observations][3] */
  /* per-voxel sensor lists (RobotUtils.buildSensorizingFunction names other than `uniform-...`, e.g. `spinedTouch-...`): when
   * vox_sensor_off is non-NULL the sensors of voxel v (grid order) are vox_sensors[vox_sensor_off[v] .. vox_sensor_off[v + 1]) and the
   * uniform list passed to vo_run_ex is ignored; CentralizedSensing concatenates the voxels' readings in grid order */
  const vo_sensor* vox_sensors;
  const int*       vox_sensor_off;
} vo_extras;

void vo_default_params(vo_params* p);

/* counts: voxels, springs, ropes, welds for a mask (w*h, row-major y*w+x, y=0 bottom) */
int vo_counts(const vo_params* p, int w, int h, const uint8_t* mask, int* n_vox, int* n_springs, int* n_ropes, int* n_welds);

/* Genome length for the given body/sensors/controller (or <0 on error). */
int vo_genome_len(int w, int h, const uint8_t* mask, const vo_sensor* sensors, int n_sensors, const vo_ctrl* ctrl);
/* ... with per-voxel sensor lists (distributed controllers need the same reading count on every voxel) */
int vo_genome_len_vox(int w, int h, const uint8_t* mask, const vo_sensor* vox_sensors, const int* vox_sensor_off, const vo_ctrl* ctrl);

/*
 * Run one episode.
 *  trace_steps : number of leading steps for which to record traces (0 = none)
 *  trace_vertices : [trace_steps][n_vox][4][2] outer voxel-poly vertices after each step (may be NULL)
 *  trace_bodies   : [trace_steps][n_vox*4][6] (x,y,angle,vx,vy,w) of every corner mass (may be NULL)
 *  trace_ctrl     : [trace_steps][n_vox] applied control value f (may be NULL)
 * returns 0 on success, negative on error.
 */
int vo_run(const vo_params* p, int w, int h, const uint8_t* mask,
           const vo_sensor* sensors, int n_sensors, const vo_ctrl* ctrl,
           const double* terrain_x, const double* terrain_y, int n_terrain,
           const double* genome, int genome_len,
           int trace_steps, double* trace_vertices, double* trace_bodies, double* trace_ctrl,
           vo_result* out);

int vo_run_ex(const vo_params* p, int w, int h, const uint8_t* mask,
              const vo_sensor* sensors, int n_sensors, const vo_ctrl* ctrl,
              const double* terrain_x, const double* terrain_y, int n_terrain,
              const double* genome, int genome_len,
              int trace_steps, double* trace_vertices, double* trace_bodies, double* trace_ctrl,
              const vo_extras* extras, vo_result* out);
/* index (in voxel order) of BehaviorUtils.getCentralElement of the body grid */
int vo_central_voxel(int w, int h, const uint8_t* mask);

/* Evaluate n genomes ([n][genome_len] row-major) on n_threads host threads; velocity -> fitness_out[n]. */
int vo_eval_batch(const vo_params* p, int w, int h, const uint8_t* mask,
                  const vo_sensor* sensors, int n_sensors, const vo_ctrl* ctrl,
                  const double* terrain_x, const double* terrain_y, int n_terrain,
                  const double* genomes, int genome_len, int n, int n_threads,
                  double* fitness_out, vo_result* results_out);

/* ... with per-voxel sensor lists (see vo_extras) */
int vo_eval_batch_vox(const vo_params* p, int w, int h, const uint8_t* mask,
                      const vo_sensor* vox_sensors, const int* vox_sensor_off, const vo_ctrl* ctrl,
                      const double* terrain_x, const double* terrain_y, int n_terrain,
                      const double* genomes, int genome_len, int n, int n_threads,
                      double* fitness_out, vo_result* results_out);

/* Low-level test hooks for known-answer tests (single bodies / single joints, no robot). */
int vo_test_free_fall(const vo_params* p, int n_steps, double* y_out, double* vy_out);
int vo_test_spring_pair(const vo_params* p, double rest, double initial_len, int n_steps, double* len_out);
int vo_test_box_drop(const vo_params* p, double angle0, double height, double vx0, int n_steps, double* state_out /*[n][6]*/);

#ifdef __cplusplus
}
#endif
#endif
