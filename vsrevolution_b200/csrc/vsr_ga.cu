// vsr_ga.cu -- device-resident generation loop (include/vsr_ga.h).  Restates the solver the reference names `numGA`
// (locomotion/Starter.java:178 -> builder/solver/DoublesStandard.java:41-71 -> JGEA StandardEvolver): tournament parent
// selection, uniform crossover, Gaussian mutation, overlapping (mu + lambda) truncation with worst-first removal.
// Population and offspring genomes stay in HBM for the whole run; the episode kernel reads them in place.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/vsr_ga.h"
#include "vsr_internal.h"

namespace {

// ---- Philox4x32-10 (Salmon et al., SC'11): every draw is a pure function of (seed; generation, individual, stream, block)
struct U4 { uint32_t x, y, z, w; };
constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u, kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;

__host__ __device__ inline U4 philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)kPhiloxM0 * c0, p1 = (uint64_t)kPhiloxM1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += kPhiloxW0; k1 += kPhiloxW1;
  }
  return U4{c0, c1, c2, c3};
}

// streams (third counter word)
enum : uint32_t { S_INIT = 0, S_OP = 1, S_PARENT1 = 2, S_PARENT2 = 3, S_XOVER = 4, S_MUT = 5 };

__host__ __device__ inline double u53(uint32_t hi, uint32_t lo) { // [0, 1): 53 random bits
  return (double)(((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

struct GaK {
  uint32_t k0, k1;
  int n_pop, G, n_tour;
  double p_xover, sigma, init_min, init_span;
};

// generation 0: genes ~ U(init_min, init_max)   (UniformDoubleFactory, DoublesStandard.java:52-53)
__global__ void ga_init_kernel(GaK p, double* __restrict__ genomes) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x, total = (size_t)p.n_pop * p.G;
  if (idx >= total) return;
  const uint32_t i = (uint32_t)(idx / p.G), g = (uint32_t)(idx % p.G);
  const U4 r = philox(p.k0, p.k1, 0u, i, S_INIT, g);
  genomes[idx] = p.init_min + p.init_span * u53(r.x, r.y);
}

// parent selection.  The population is kept sorted best-first with ties broken towards the lower index, so the winner of a
// tournament (JGEA Tournament: nTournament uniform draws with replacement, keep the best) is the smallest drawn index.
// One thread per (offspring, parent slot).
__global__ void ga_select_kernel(GaK p, uint32_t gen, int32_t* __restrict__ parents) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * p.n_pop) return;
  const uint32_t i = (uint32_t)(t >> 1), slot = (uint32_t)(t & 1);
  if (slot == 1) {
    const U4 op = philox(p.k0, p.k1, gen, i, S_OP, 0u);
    if (!((double)op.x * (1.0 / 4294967296.0) < p.p_xover)) { // GaussianMutation alone: one parent
      parents[t] = -1;
      return;
    }
  }
  uint32_t best = 0xFFFFFFFFu;
  for (int d = 0; d < p.n_tour; d += 4) {
    const U4 r = philox(p.k0, p.k1, gen, i, slot ? S_PARENT2 : S_PARENT1, (uint32_t)(d >> 2));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int q = 0; q < 4; q++)
      if (d + q < p.n_tour) best = min(best, (uint32_t)(((uint64_t)w[q] * (uint64_t)p.n_pop) >> 32));
  }
  parents[t] = (int32_t)best;
}

// variation: UniformCrossover (each gene from either parent with probability 1/2) then GaussianMutation (every gene += N(0, sigma^2))
__global__ void ga_vary_kernel(GaK p, uint32_t gen, const double* __restrict__ pop, const int32_t* __restrict__ parents,
                               double* __restrict__ off) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x, total = (size_t)p.n_pop * p.G;
  if (idx >= total) return;
  const uint32_t i = (uint32_t)(idx / p.G), g = (uint32_t)(idx % p.G);
  const int32_t p1 = parents[2 * i], p2 = parents[2 * i + 1];
  int32_t src = p1;
  if (p2 >= 0) {
    const U4 x = philox(p.k0, p.k1, gen, i, S_XOVER, g >> 7); // 128 coin flips per block
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
    if ((w[(g >> 5) & 3] >> (g & 31)) & 1u) src = p2;
  }
  const U4 r = philox(p.k0, p.k1, gen, i, S_MUT, g);
  const double u1 = u53(r.x, r.y) + (1.0 / 9007199254740992.0); // (0, 1]
  const double u2 = u53(r.z, r.w);
  const double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
  off[idx] = pop[(size_t)src * p.G + g] + p.sigma * z;
}

// ---- ES (JGEA SimpleEvolutionaryStrategy; builder/solver/SimpleES.java:41-49) ---------------------------------------------------
// means[g] = mean of gene g over the best n_par individuals (the population is sorted: rows 0..n_par-1), summed in rank order so
// that the result does not depend on the launch geometry (and oracle/ga.py reproduces it bit for bit)
__global__ void es_means_kernel(const double* __restrict__ pop, int G, int n_par, double* __restrict__ means) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  double s = 0;
  for (int i = 0; i < n_par; i++) s += pop[(size_t)i * G + g];
  means[g] = s / (double)n_par;
}
// next generation, rows [0, n_elites) = the elites (genome and fitness), rows [n_elites, n_pop) = means + sigma * N(0, 1)
__global__ void es_sample_kernel(GaK p, uint32_t gen, int n_elites, const double* __restrict__ pop, const double* __restrict__ pop_fit,
                                 const double* __restrict__ means, double* __restrict__ nxt, double* __restrict__ nxt_fit) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x, total = (size_t)p.n_pop * p.G;
  if (idx >= total) return;
  const uint32_t i = (uint32_t)(idx / p.G), g = (uint32_t)(idx % p.G);
  if ((int)i < n_elites) {
    nxt[idx] = pop[idx];
    if (g == 0) nxt_fit[i] = pop_fit[i];
    return;
  }
  const U4 r = philox(p.k0, p.k1, gen, i - (uint32_t)n_elites, S_MUT, g);
  const double u1 = u53(r.x, r.y) + (1.0 / 9007199254740992.0); // (0, 1]
  const double u2 = u53(r.z, r.w);
  const double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
  nxt[idx] = means[g] + p.sigma * z;
}

// rank[i] = number of individuals that beat i: higher fitness first, NaN as -inf, ties towards the lower index (a strict
// total order, so the ranks are a permutation).  Tiled all-pairs count; blockIdx.y splits the opponents.
constexpr int kRankTile = 2048;
constexpr int kAllPairsMax = 8192; // larger merged sets are ranked by the bitonic sort below
__global__ void ga_rank_kernel(const double* __restrict__ fit, int m, int32_t* __restrict__ rank) {
  __shared__ double tile[kRankTile];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j0 = blockIdx.y * kRankTile, j1 = min(m, j0 + kRankTile);
  for (int j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
    double f = fit[j];
    tile[j - j0] = f == f ? f : -INFINITY;
  }
  __syncthreads();
  if (i >= m) return;
  double fi = fit[i];
  fi = fi == fi ? fi : -INFINITY;
  int c = 0;
  for (int j = j0; j < j1; j++) {
    const double fj = tile[j - j0];
    c += (fj > fi) || (fj == fi && j < i);
  }
  if (c) atomicAdd(&rank[i], c);
}

// ---- large populations: ranks by a bitonic sort of (key, index) pairs instead of the all-pairs count (O(m log^2 m)) ----------
// key = bit pattern that orders doubles descending (best first), NaN as -inf, -0 as +0; padding entries sort last
__device__ inline uint64_t order_key(double f) {
  if (!(f == f)) f = -INFINITY;
  if (f == 0.0) f = 0.0;
  const uint64_t u = (uint64_t)__double_as_longlong(f);
  const uint64_t asc = (u >> 63) ? ~u : (u | 0x8000000000000000ull);
  return ~asc;
}
__global__ void ga_keys_kernel(const double* __restrict__ fit, int m, int P, uint64_t* __restrict__ keys, uint32_t* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  keys[i] = i < m ? order_key(fit[i]) : ~0ull;
  idx[i] = (uint32_t)i;
}
__device__ inline void cmp_swap(uint64_t& ka, uint32_t& ia, uint64_t& kb, uint32_t& ib, bool up) {
  const bool a_first = ka < kb || (ka == kb && ia < ib);
  if (a_first != up) { const uint64_t tk = ka; ka = kb; kb = tk; const uint32_t ti = ia; ia = ib; ib = ti; }
}
// one (k, j) step in global memory: thread t owns the pair (i, i | j) with bit j of i clear
__global__ void ga_bitonic_global(uint64_t* __restrict__ keys, uint32_t* __restrict__ idx, int P, int k, int j) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P / 2) return;
  const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
  uint64_t ka = keys[i], kb = keys[l];
  uint32_t ia = idx[i], ib = idx[l];
  cmp_swap(ka, ia, kb, ib, (i & k) == 0);
  keys[i] = ka; keys[l] = kb; idx[i] = ia; idx[l] = ib;
}
// all steps j = j0 .. 1 of stage k for a tile of 2048 elements in shared memory (j0 <= 1024)
constexpr int kSortTile = 2048;
__global__ void ga_bitonic_shared(uint64_t* __restrict__ keys, uint32_t* __restrict__ idx, int k, int j0) {
  __shared__ uint64_t sk[kSortTile];
  __shared__ uint32_t si[kSortTile];
  const int base = blockIdx.x * kSortTile;
  for (int e = threadIdx.x; e < kSortTile; e += blockDim.x) { sk[e] = keys[base + e]; si[e] = idx[base + e]; }
  __syncthreads();
  for (int j = j0; j > 0; j >>= 1) {
    const int t = threadIdx.x;
    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
    cmp_swap(sk[i], si[i], sk[l], si[l], ((base + i) & k) == 0);
    __syncthreads();
  }
  for (int e = threadIdx.x; e < kSortTile; e += blockDim.x) { keys[base + e] = sk[e]; idx[base + e] = si[e]; }
}
__global__ void ga_rank_from_sorted(const uint32_t* __restrict__ idx, int m, int P, int32_t* __restrict__ rank) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= P) return;
  const uint32_t i = idx[r];
  if (i < (uint32_t)m) rank[i] = r; // padding sorts last, so the first m sorted entries are the real ones
}

// survivors: row i of the merged set goes to row rank[i] of the next population if rank[i] < n_keep (Last() removes the worst)
__global__ void ga_scatter_kernel(const double* __restrict__ src, const double* __restrict__ src_fit, const int32_t* __restrict__ rank,
                                  int G, int n_keep, double* __restrict__ dst, double* __restrict__ dst_fit, int32_t* __restrict__ origin) {
  const int i = blockIdx.x, r = rank[i];
  if (r >= n_keep) return;
  const double* s = src + (size_t)i * G;
  double* d = dst + (size_t)r * G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) d[g] = s[g];
  if (threadIdx.x == 0) { dst_fit[r] = src_fit[i]; origin[r] = i; }
}

} // namespace

struct vsr_ga {
  vsr_ctx* ctx = nullptr;
  vsr_ga_desc d;
  GaK k;
  int G = 0, dev = 0;
  cudaStream_t st = nullptr;
  double* genomes[2] = {nullptr, nullptr}; // [2 nPop][G]: population rows then offspring rows
  double* fit[2] = {nullptr, nullptr};     // [2 nPop + 8]
  int32_t *parents = nullptr, *rank = nullptr, *origin = nullptr;
  double* means = nullptr; // ES: [G]
  uint64_t* sort_keys = nullptr; // bitonic sort scratch: next power of two >= 2 nPop entries
  uint32_t* sort_idx = nullptr;
  int sort_cap = 0;
  int cur = 0;        // buffer holding the current population
  int off_buf = -1;   // buffer holding the most recent offspring
  int shard_rank = 0, shard_world = 1;
  int pending = 0;    // 0 none, 1 init, 2 step
  int pend_count = 0, pend_base = 0;
  int64_t gen = -1, evals = 0, launches = 0;
  std::string err;
};

namespace {

int gfail(vsr_ga* ga, int code, const std::string& msg) {
  if (ga) ga->err = msg;
  return code;
}

#define GCK(call)                                                                                   \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) return gfail(ga, VSR_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

void shard_of(const vsr_ga* ga, int count, int* lo, int* hi) {
  const int per = (count + ga->shard_world - 1) / ga->shard_world;
  *lo = std::min(count, ga->shard_rank * per);
  *hi = std::min(count, *lo + per);
}

// evaluate this process's shard of rows [base, base + count) of genomes[cur]
int eval_pending(vsr_ga* ga, int base, int count, int pending) {
  int lo, hi;
  shard_of(ga, count, &lo, &hi);
  int rc = vsr_internal::eval_range(ga->ctx, ga->genomes[ga->cur] + (size_t)base * ga->G, (size_t)lo, (size_t)hi, ga->fit[ga->cur] + base, ga->st);
  if (rc) return gfail(ga, rc, vsr_internal::last_error(ga->ctx));
  ga->pending = pending; ga->pend_base = base; ga->pend_count = count;
  return VSR_OK;
}

// rank rows [base, base + m) of buffer `cur`, keep the best n_pop in the other buffer, swap
int truncate_into_next(vsr_ga* ga, int m, int base = 0) {
  const int n = ga->d.n_pop, nxt = ga->cur ^ 1;
  const double* src_fit = ga->fit[ga->cur] + base;
  const double* src_gen = ga->genomes[ga->cur] + (size_t)base * ga->G;
  GCK(cudaSetDevice(ga->dev));
  if (m <= kAllPairsMax) {
    GCK(cudaMemsetAsync(ga->rank, 0, sizeof(int32_t) * m, ga->st));
    dim3 grid((m + 255) / 256, (m + kRankTile - 1) / kRankTile);
    ga_rank_kernel<<<grid, 256, 0, ga->st>>>(src_fit, m, ga->rank);
    ga->launches += 1;
  } else {
    int P = kSortTile;
    while (P < m) P <<= 1;
    ga_keys_kernel<<<(P + 255) / 256, 256, 0, ga->st>>>(src_fit, m, P, ga->sort_keys, ga->sort_idx);
    ga->launches += 1;
    for (int k = 2; k <= P; k <<= 1) {
      int j = k >> 1;
      for (; j >= kSortTile; j >>= 1) { // partners farther apart than a tile: one step per launch
        ga_bitonic_global<<<(P / 2 + 255) / 256, 256, 0, ga->st>>>(ga->sort_keys, ga->sort_idx, P, k, j);
        ga->launches += 1;
      }
      ga_bitonic_shared<<<P / kSortTile, kSortTile / 2, 0, ga->st>>>(ga->sort_keys, ga->sort_idx, k, j);
      ga->launches += 1;
    }
    ga_rank_from_sorted<<<(P + 255) / 256, 256, 0, ga->st>>>(ga->sort_idx, m, P, ga->rank);
    ga->launches += 1;
  }
  ga_scatter_kernel<<<m, 64, 0, ga->st>>>(src_gen, src_fit, ga->rank, ga->G, n, ga->genomes[nxt], ga->fit[nxt], ga->origin);
  GCK(cudaGetLastError());
  ga->launches += 1;
  ga->cur = nxt;
  return VSR_OK;
}

} // namespace

extern "C" {

int vsr_ga_default(vsr_ga_desc* d) {
  if (!d) return VSR_EINVAL;
  std::memset(d, 0, sizeof(*d));
  d->xover_prob = 0.75; d->tournament_rate = 0.05; d->min_tournament = 3; d->sigma_mut = 0.35; // Starter.java:178
  d->init_min = -1.0; d->init_max = 1.0;                                                        // DoublesStandard.java:53
  return VSR_OK;
}

int vsr_es_default(vsr_ga_desc* d, int32_t n_pop) {
  if (!d || n_pop <= 0) return VSR_EINVAL;
  std::memset(d, 0, sizeof(*d));
  d->solver = VSR_SOLVER_ES; d->n_pop = n_pop;
  d->sigma_mut = 0.35;                                       // Starter.java:179: new SimpleES(0.35, 0.4)
  d->n_parents = (int32_t)std::floor(n_pop * 0.4 + 0.5);     // (int) Math.round(nPop * parentsRate), SimpleES.java:46
  d->n_elites = 1;                                           // SimpleES.java:47
  d->init_min = -1.0; d->init_max = 1.0;                     // SimpleES.java:43
  d->min_tournament = 1;
  return VSR_OK;
}

int vsr_ga_tournament_size(const vsr_ga_desc* d) {
  if (!d || d->n_pop <= 0) return VSR_EINVAL;
  return std::max(d->min_tournament, (int)std::ceil((double)d->n_pop * d->tournament_rate)); // DoublesStandard.java:65
}

int vsr_ga_create(vsr_ctx* ctx, const vsr_ga_desc* d, vsr_ga** out) {
  if (!ctx || !d || !out) return VSR_EINVAL;
  *out = nullptr;
  if (!vsr_internal::ready(ctx)) return VSR_ESTATE;
  if (d->n_pop <= 0 || d->n_pop > (1 << 22) || !(d->xover_prob >= 0 && d->xover_prob <= 1) || !(d->tournament_rate >= 0) ||
      d->min_tournament < 1 || !(d->sigma_mut >= 0) || !(d->init_max >= d->init_min))
    return VSR_EINVAL;
  if (d->solver != VSR_SOLVER_NUM_GA && d->solver != VSR_SOLVER_ES) return VSR_EINVAL;
  if (d->solver == VSR_SOLVER_ES && (d->n_parents < 1 || d->n_parents > d->n_pop || d->n_elites < 0 || d->n_elites >= d->n_pop))
    return VSR_EINVAL;
  vsr_ga* ga = new (std::nothrow) vsr_ga();
  if (!ga) return VSR_ENOMEM;
  ga->ctx = ctx; ga->d = *d; ga->G = vsr_internal::genome_len(ctx);
  ga->dev = vsr_internal::primary_device(ctx);
  ga->st = vsr_internal::primary_stream(ctx);
  GaK& k = ga->k;
  k.k0 = (uint32_t)d->seed; k.k1 = (uint32_t)(d->seed >> 32);
  k.n_pop = d->n_pop; k.G = ga->G; k.n_tour = vsr_ga_tournament_size(d);
  k.p_xover = d->xover_prob; k.sigma = d->sigma_mut; k.init_min = d->init_min; k.init_span = d->init_max - d->init_min;
  const size_t n = (size_t)d->n_pop, rows = 2 * n;
  bool ok = cudaSetDevice(ga->dev) == cudaSuccess;
  for (int b = 0; b < 2 && ok; b++) {
    ok = ok && cudaMalloc(&ga->genomes[b], rows * ga->G * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMalloc(&ga->fit[b], (rows + 8) * sizeof(double)) == cudaSuccess;
    ok = ok && cudaMemset(ga->fit[b], 0, (rows + 8) * sizeof(double)) == cudaSuccess;
  }
  ok = ok && cudaMalloc(&ga->parents, rows * sizeof(int32_t)) == cudaSuccess;
  ok = ok && cudaMalloc(&ga->rank, rows * sizeof(int32_t)) == cudaSuccess;
  ok = ok && cudaMalloc(&ga->origin, n * sizeof(int32_t)) == cudaSuccess;
  ok = ok && cudaMemset(ga->parents, 0xFF, rows * sizeof(int32_t)) == cudaSuccess; // ES: no parent indices (-1)
  ok = ok && cudaMalloc(&ga->means, (size_t)ga->G * sizeof(double)) == cudaSuccess;
  if (rows > (size_t)kAllPairsMax) {
    int P = kSortTile;
    while ((size_t)P < rows) P <<= 1;
    ga->sort_cap = P;
    ok = ok && cudaMalloc(&ga->sort_keys, (size_t)P * sizeof(uint64_t)) == cudaSuccess;
    ok = ok && cudaMalloc(&ga->sort_idx, (size_t)P * sizeof(uint32_t)) == cudaSuccess;
  }
  if (!ok) { cudaGetLastError(); vsr_ga_destroy(ga); return VSR_ENOMEM; }
  *out = ga;
  return VSR_OK;
}

void vsr_ga_destroy(vsr_ga* ga) {
  if (!ga) return;
  cudaSetDevice(ga->dev);
  if (ga->st) cudaStreamSynchronize(ga->st);
  for (int b = 0; b < 2; b++) { cudaFree(ga->genomes[b]); cudaFree(ga->fit[b]); }
  cudaFree(ga->parents); cudaFree(ga->rank); cudaFree(ga->origin); cudaFree(ga->means); cudaFree(ga->sort_keys); cudaFree(ga->sort_idx);
  delete ga;
}

const char* vsr_ga_last_error(const vsr_ga* ga) { return ga ? ga->err.c_str() : "null vsr_ga"; }

int vsr_ga_set_shard(vsr_ga* ga, int32_t rank, int32_t world) {
  if (!ga) return VSR_EINVAL;
  if (world < 1 || world > 8 || rank < 0 || rank >= world) return gfail(ga, VSR_EINVAL, "bad shard (rank, world)");
  if (ga->pending) return gfail(ga, VSR_ESTATE, "a generation is in flight");
  ga->shard_rank = rank; ga->shard_world = world;
  return VSR_OK;
}

int vsr_ga_init_begin(vsr_ga* ga) {
  if (!ga) return VSR_EINVAL;
  if (ga->pending) return gfail(ga, VSR_ESTATE, "a generation is in flight");
  GCK(cudaSetDevice(ga->dev));
  const size_t total = (size_t)ga->d.n_pop * ga->G;
  ga_init_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ga->st>>>(ga->k, ga->genomes[ga->cur]);
  GCK(cudaGetLastError());
  ga->launches++;
  ga->off_buf = -1;
  return eval_pending(ga, 0, ga->d.n_pop, 1);
}

int vsr_ga_init_end(vsr_ga* ga) {
  if (!ga) return VSR_EINVAL;
  if (ga->pending != 1) return gfail(ga, VSR_ESTATE, "vsr_ga_init_end without vsr_ga_init_begin");
  int rc = truncate_into_next(ga, ga->d.n_pop);
  if (rc) return rc;
  ga->pending = 0; ga->gen = 0; ga->evals += ga->d.n_pop;
  return VSR_OK;
}

int vsr_ga_init(vsr_ga* ga) {
  int rc = vsr_ga_init_begin(ga);
  return rc ? rc : vsr_ga_init_end(ga);
}

int vsr_ga_step_begin(vsr_ga* ga) {
  if (!ga) return VSR_EINVAL;
  if (ga->pending) return gfail(ga, VSR_ESTATE, "a generation is in flight");
  if (ga->gen < 0) return gfail(ga, VSR_ESTATE, "vsr_ga_init (or vsr_ga_set_population) must come first");
  GCK(cudaSetDevice(ga->dev));
  const int n = ga->d.n_pop;
  const uint32_t gen = (uint32_t)(ga->gen + 1);
  const size_t total = (size_t)n * ga->G;
  double* pop = ga->genomes[ga->cur];
  double* off = pop + total;
  if (ga->d.solver == VSR_SOLVER_ES) {
    const int e = ga->d.n_elites;
    es_means_kernel<<<(ga->G + 63) / 64, 64, 0, ga->st>>>(pop, ga->G, ga->d.n_parents, ga->means);
    es_sample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ga->st>>>(ga->k, gen, e, pop, ga->fit[ga->cur], ga->means, off, ga->fit[ga->cur] + n);
    GCK(cudaGetLastError());
    ga->launches += 2;
    ga->off_buf = ga->cur;
    // the elites keep their fitness unless remap (SimpleEvolutionaryStrategy.update: map(offspringGenotypes, elites, ...))
    return ga->d.remap ? eval_pending(ga, n, n, 2) : eval_pending(ga, n + e, n - e, 2);
  }
  ga_select_kernel<<<(2 * n + 127) / 128, 128, 0, ga->st>>>(ga->k, gen, ga->parents);
  ga_vary_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ga->st>>>(ga->k, gen, pop, ga->parents, off);
  GCK(cudaGetLastError());
  ga->launches += 2;
  ga->off_buf = ga->cur;
  return ga->d.remap ? eval_pending(ga, 0, 2 * n, 2) : eval_pending(ga, n, n, 2);
}

int vsr_ga_step_end(vsr_ga* ga) {
  if (!ga) return VSR_EINVAL;
  if (ga->pending != 2) return gfail(ga, VSR_ESTATE, "vsr_ga_step_end without vsr_ga_step_begin");
  const int n = ga->d.n_pop;
  int rc = ga->d.solver == VSR_SOLVER_ES ? truncate_into_next(ga, n, n) : truncate_into_next(ga, 2 * n);
  if (rc) return rc;
  ga->evals += ga->pend_count;
  ga->pending = 0; ga->gen++;
  return VSR_OK;
}

int vsr_ga_step(vsr_ga* ga) {
  int rc = vsr_ga_step_begin(ga);
  return rc ? rc : vsr_ga_step_end(ga);
}

int vsr_ga_pending(vsr_ga* ga, double** d_fitness, int32_t* count, int32_t* lo, int32_t* hi, void** stream) {
  if (!ga) return VSR_EINVAL;
  if (!ga->pending) return gfail(ga, VSR_ESTATE, "no generation in flight");
  int l, h;
  shard_of(ga, ga->pend_count, &l, &h);
  if (d_fitness) *d_fitness = ga->fit[ga->cur] + ga->pend_base;
  if (count) *count = ga->pend_count;
  if (lo) *lo = l;
  if (hi) *hi = h;
  if (stream) *stream = (void*)ga->st;
  return VSR_OK;
}

int vsr_ga_sync(vsr_ga* ga) {
  if (!ga) return VSR_EINVAL;
  GCK(cudaSetDevice(ga->dev));
  GCK(cudaStreamSynchronize(ga->st));
  return VSR_OK;
}

int vsr_ga_best(vsr_ga* ga, double* genome_out, double* fitness_out) {
  if (!ga) return VSR_EINVAL;
  if (ga->gen < 0 || ga->pending) return gfail(ga, VSR_ESTATE, "no completed generation");
  GCK(cudaSetDevice(ga->dev));
  if (genome_out) GCK(cudaMemcpyAsync(genome_out, ga->genomes[ga->cur], ga->G * sizeof(double), cudaMemcpyDeviceToHost, ga->st));
  if (fitness_out) GCK(cudaMemcpyAsync(fitness_out, ga->fit[ga->cur], sizeof(double), cudaMemcpyDeviceToHost, ga->st));
  GCK(cudaStreamSynchronize(ga->st));
  return VSR_OK;
}

int vsr_ga_population(vsr_ga* ga, double* genomes_out, double* fitness_out) {
  if (!ga) return VSR_EINVAL;
  if (ga->gen < 0 || ga->pending) return gfail(ga, VSR_ESTATE, "no completed generation");
  GCK(cudaSetDevice(ga->dev));
  const size_t n = (size_t)ga->d.n_pop;
  if (genomes_out) GCK(cudaMemcpyAsync(genomes_out, ga->genomes[ga->cur], n * ga->G * sizeof(double), cudaMemcpyDeviceToHost, ga->st));
  if (fitness_out) GCK(cudaMemcpyAsync(fitness_out, ga->fit[ga->cur], n * sizeof(double), cudaMemcpyDeviceToHost, ga->st));
  GCK(cudaStreamSynchronize(ga->st));
  return VSR_OK;
}

int vsr_ga_offspring(vsr_ga* ga, double* genomes_out, int32_t* parents_out, double* fitness_out) {
  if (!ga) return VSR_EINVAL;
  if (ga->off_buf < 0) return gfail(ga, VSR_ESTATE, "no offspring yet");
  GCK(cudaSetDevice(ga->dev));
  const size_t n = (size_t)ga->d.n_pop;
  if (genomes_out) GCK(cudaMemcpyAsync(genomes_out, ga->genomes[ga->off_buf] + n * ga->G, n * ga->G * sizeof(double), cudaMemcpyDeviceToHost, ga->st));
  if (parents_out) GCK(cudaMemcpyAsync(parents_out, ga->parents, 2 * n * sizeof(int32_t), cudaMemcpyDeviceToHost, ga->st));
  if (fitness_out) GCK(cudaMemcpyAsync(fitness_out, ga->fit[ga->off_buf] + n, n * sizeof(double), cudaMemcpyDeviceToHost, ga->st));
  GCK(cudaStreamSynchronize(ga->st));
  return VSR_OK;
}

int vsr_ga_means(vsr_ga* ga, double* means_out) {
  if (!ga || !means_out) return VSR_EINVAL;
  if (ga->d.solver != VSR_SOLVER_ES || ga->gen < 1) return gfail(ga, VSR_ESTATE, "no ES generation step yet");
  GCK(cudaSetDevice(ga->dev));
  GCK(cudaMemcpyAsync(means_out, ga->means, (size_t)ga->G * sizeof(double), cudaMemcpyDeviceToHost, ga->st));
  GCK(cudaStreamSynchronize(ga->st));
  return VSR_OK;
}

int vsr_ga_survivors(vsr_ga* ga, int32_t* origin_out) {
  if (!ga || !origin_out) return VSR_EINVAL;
  if (ga->gen < 0 || ga->pending) return gfail(ga, VSR_ESTATE, "no completed generation");
  GCK(cudaSetDevice(ga->dev));
  GCK(cudaMemcpyAsync(origin_out, ga->origin, (size_t)ga->d.n_pop * sizeof(int32_t), cudaMemcpyDeviceToHost, ga->st));
  GCK(cudaStreamSynchronize(ga->st));
  return VSR_OK;
}

int vsr_ga_set_population(vsr_ga* ga, const double* genomes, const double* fitness) {
  if (!ga || !genomes || !fitness) return VSR_EINVAL;
  if (ga->pending) return gfail(ga, VSR_ESTATE, "a generation is in flight");
  GCK(cudaSetDevice(ga->dev));
  const size_t n = (size_t)ga->d.n_pop;
  GCK(cudaMemcpyAsync(ga->genomes[ga->cur], genomes, n * ga->G * sizeof(double), cudaMemcpyHostToDevice, ga->st));
  GCK(cudaMemcpyAsync(ga->fit[ga->cur], fitness, n * sizeof(double), cudaMemcpyHostToDevice, ga->st));
  int rc = truncate_into_next(ga, ga->d.n_pop);
  if (rc) return rc;
  GCK(cudaStreamSynchronize(ga->st)); // the caller's buffers are free again
  if (ga->gen < 0) ga->gen = 0;
  ga->off_buf = -1;
  return VSR_OK;
}

int64_t vsr_ga_generation(const vsr_ga* ga) { return ga ? ga->gen : -1; }
int64_t vsr_ga_evaluations(const vsr_ga* ga) { return ga ? ga->evals : 0; }
int64_t vsr_ga_kernel_launches(const vsr_ga* ga) { return ga ? ga->launches : 0; }
int32_t vsr_ga_n_pop(const vsr_ga* ga) { return ga ? ga->d.n_pop : VSR_EINVAL; }
int32_t vsr_ga_genome_len(const vsr_ga* ga) { return ga ? ga->G : VSR_EINVAL; }

} // extern "C"
