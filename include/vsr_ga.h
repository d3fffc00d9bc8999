/*
 * vsr_ga.h -- C ABI of the device-resident generation loop of libvsr_eval.so (BASELINE config 5).
 *
 * Restates the solver the reference names `numGA`:
 *     new DoublesStandard(0.75, 0.05, 3, 0.35)                       locomotion/Starter.java:178
 *     builder/solver/DoublesStandard.java:41-71  ->  JGEA StandardEvolver(
 *         factory   = FixedLengthListFactory(G, UniformDoubleFactory(-1, 1)),                 :52-53
 *         operators = { GaussianMutation(sigma): 1 - pX,  UniformCrossover.andThen(GaussianMutation(sigma)): pX },  :54-57
 *         parent selector     = Tournament(max(minNTournament, ceil(nPop * tournamentRate))),                 :65
 *         unsurvival selector = Last(),  nOffspring = nPop,  overlapping = true,  remap)                    :66-69
 * with the fitness = Outcome::getVelocity, maximised (Starter.java:174, comparator :434).
 *
 * One generation == JGEA StandardEvolver.update():
 *   1. nPop offspring; for each: operator ~ {mutation: 1-pX, crossover+mutation: pX}; every parent = the best of
 *      nTournament individuals drawn uniformly with replacement; UniformCrossover takes each gene from parent 1 or 2
 *      with probability 1/2; GaussianMutation adds N(0, sigma^2) to every gene
 *   2. the offspring are evaluated (here: ONE launch of the episode kernel over all of them)
 *   3. overlapping: parents + offspring are merged and the worst are removed until nPop remain ((mu + lambda) truncation)
 * Semantics parity, not stream parity: JGEA consumes one sequential java.util.Random; here every draw is a pure function
 * Philox4x32-10(key = seed, counter = (generation, individual, stream, block)), so the result does not depend on the launch
 * geometry, on the number of GPUs, or on the order of evaluation -- and oracle/ga.py reproduces it index for index.
 * Deterministic tie rules replace JGEA's random tie-breaks: equal fitness -> lower population index wins; NaN fitness ranks last.
 *
 * Genomes never leave the device: population [nPop][G] + offspring [nPop][G] doubles stay in HBM, only the best individual
 * (or, on request, the population) is copied back.
 *
 * Multi-GPU: the GA state is replicated (same seed => every rank generates the identical offspring), each rank
 * EVALUATES only its contiguous shard [lo, hi) and the ranks exchange 8 bytes per robot (fitness all-gather) between
 * vsr_ga_*_begin and vsr_ga_*_end.  Inside one process, a context created with n_devices > 1 shards the evaluation over its
 * devices with peer copies and needs no exchange.
 */
#ifndef VSR_GA_H
#define VSR_GA_H

#include <stdint.h>

#include "vsr_eval.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int32_t n_pop;           /* nPop (solver parameter, DoublesStandard.java:42)                                   */
  double xover_prob;       /* xOverProb = 0.75                                                                   */
  double tournament_rate;  /* tournamentRate = 0.05                                                              */
  int32_t min_tournament;  /* minNTournament = 3                                                                 */
  double sigma_mut;        /* sigmaMut = 0.35                                                                    */
  double init_min, init_max; /* UniformDoubleFactory(-1, 1)                                                      */
  uint64_t seed;
  int32_t remap;           /* `remap`: re-evaluate the parents every generation (DoublesStandard.java:45); 0 = false */
  /* which JGEA solver the loop restates: VSR_SOLVER_NUM_GA (above) or VSR_SOLVER_ES, the reference's `ES` =
   * new SimpleES(0.35, 0.4) (locomotion/Starter.java:179) -> builder/solver/SimpleES.java:41-49 ->
   * JGEA SimpleEvolutionaryStrategy(mapper, UniformDoubleFactory(-1, 1) genotypes, nPop, stop,
   *                                 nOfParents = round(nPop * parentsRate), nOfElites = 1, sigma, remap).
   * One ES generation: the population is sorted; means[j] = mean of gene j over the best n_parents; the next population is the
   * n_elites best individuals, kept with their fitness (re-evaluated if remap), plus nPop - n_elites offspring
   * means + sigma_mut * N(0, 1) per gene; non-overlapping (the rest of the old population is dropped).  xover_prob,
   * tournament_rate and min_tournament are ignored.  JGEA 2.0.1 is un-vendored: semantics [U], as for numGA.                */
  int32_t solver;
  int32_t n_parents;       /* ES: nOfParents = (int) Math.round(nPop * parentsRate), SimpleES.java:46                     */
  int32_t n_elites;        /* ES: nOfElites = 1, SimpleES.java:47                                                          */
} vsr_ga_desc;

enum { VSR_SOLVER_NUM_GA = 0, VSR_SOLVER_ES = 1 };

typedef struct vsr_ga vsr_ga;

int vsr_ga_default(vsr_ga_desc* d);            /* numGA constants; n_pop = 0 (caller sets), seed 0                */
int vsr_ga_tournament_size(const vsr_ga_desc* d); /* max(minNTournament, ceil(nPop * tournamentRate))             */
/* ES constants of Starter.java:179 for a population of n_pop: sigma 0.35, n_parents = round(0.4 n_pop), n_elites = 1 */
int vsr_es_default(vsr_ga_desc* d, int32_t n_pop);
/* ES only: the means of the most recent generation step ([G] doubles, HOST buffer)                                  */
int vsr_ga_means(vsr_ga* ga, double* means_out);

/* ctx must be ready (body + controller set); the GA borrows it (destroy the GA first). */
int vsr_ga_create(vsr_ctx* ctx, const vsr_ga_desc* d, vsr_ga** out);
void vsr_ga_destroy(vsr_ga* ga);
const char* vsr_ga_last_error(const vsr_ga* ga);

/* evaluation shard of THIS process: robots [lo, hi) of every batch are evaluated here (default: all) */
int vsr_ga_set_shard(vsr_ga* ga, int32_t rank, int32_t world);

/* generation 0: draw the initial population and evaluate it.  init == init_begin + init_end.                    */
int vsr_ga_init(vsr_ga* ga);
int vsr_ga_init_begin(vsr_ga* ga);
int vsr_ga_init_end(vsr_ga* ga);
/* one generation.  step == step_begin + step_end.  Between begin and end the caller of a multi-process run all-gathers
 * the pending fitness vector (vsr_ga_pending): begin enqueues selection, variation and this rank's evaluations; end ranks,
 * merges and truncates.  Nothing synchronises the host except vsr_ga_sync / the copy-out calls.                   */
int vsr_ga_step(vsr_ga* ga);
int vsr_ga_step_begin(vsr_ga* ga);
int vsr_ga_step_end(vsr_ga* ga);
/* DEVICE pointer of the fitness vector begin() is filling ([count] doubles; count = nPop, or 2*nPop with remap: parents
 * first), this rank's shard [lo, hi) of it, and the stream the work is enqueued on (cudaStream_t)                 */
int vsr_ga_pending(vsr_ga* ga, double** d_fitness, int32_t* count, int32_t* lo, int32_t* hi, void** stream);
int vsr_ga_sync(vsr_ga* ga);

/* read-out (HOST buffers; each call synchronises).  The population is kept sorted: index 0 is the best individual. */
int vsr_ga_best(vsr_ga* ga, double* genome_out, double* fitness_out);
int vsr_ga_population(vsr_ga* ga, double* genomes_out /* [nPop][G] or NULL */, double* fitness_out /* [nPop] or NULL */);
/* the offspring of the most recent step_begin: genomes [nPop][G], parents [nPop][2] (population indices before the merge;
 * parents[i][1] = -1 for mutation-only), fitness [nPop]; any pointer may be NULL                                   */
int vsr_ga_offspring(vsr_ga* ga, double* genomes_out, int32_t* parents_out, double* fitness_out);
/* where the current population came from: origin[r] = row of the merged set (rows 0..nPop-1 = previous population, rows
 * nPop..2nPop-1 = offspring) that now sits at rank r; after init / set_population: the row of the unsorted input.
 * ES: the row of the new generation (rows 0..n_elites-1 = the elites, then the offspring in draw order)               */
int vsr_ga_survivors(vsr_ga* ga, int32_t* origin_out /* [nPop] */);
/* overwrite the population (HOST [nPop][G] genomes, [nPop] fitness values; sorted internally): resume / tests       */
int vsr_ga_set_population(vsr_ga* ga, const double* genomes, const double* fitness);
int64_t vsr_ga_generation(const vsr_ga* ga);   /* generations completed (0 after init)                            */
int64_t vsr_ga_evaluations(const vsr_ga* ga);  /* fitness evaluations so far (JGEA nOfFitnessEvaluations), whole job */
int64_t vsr_ga_kernel_launches(const vsr_ga* ga); /* GA kernels launched so far (episode kernels are counted by the ctx) */
/* sizes of the read-out buffers, from the native handle (a binding must not trust caller-supplied sizes) */
int32_t vsr_ga_n_pop(const vsr_ga* ga);
int32_t vsr_ga_genome_len(const vsr_ga* ga);

#ifdef __cplusplus
}
#endif
#endif
