// FFMA vs FFMA2 (fma.rn.f32x2) throughput on sm_100a at the episode kernel's occupancy (10 warps/SM) and at full occupancy.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k_scalar(float* out, int iters, float a, float b) {
  float x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int i = 0; i < ILP; i++) x[i] = fmaf(x[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> // ILP = number of independent float2 chains
__global__ void k_packed(float* out, int iters, float a, float b) {
  float2 x[ILP];
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int i = 0; i < ILP; i++) x[i] = __ffma2_rn(x[i], a2, b2);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class K>
void run(const char* name, K kern, int threads, int flops_per_inner, float* d) {
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<148, threads>>>(d, 100, 0.999f, 0.001f);
  cudaEventRecord(e0);
  kern<<<148, threads>>>(d, iters, 0.999f, 0.001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fl = 148.0 * threads * (double)iters * 8 * flops_per_inner * 2;
  printf("%-28s threads/SM %4d  %7.2f TFLOP/s  (%.3f ms)\n", name, threads, fl / ms / 1e9, ms);
}
int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  for (int th : {320, 1024}) {
    run("scalar FFMA ILP1", k_scalar<1>, th, 1, d);
    run("scalar FFMA ILP2", k_scalar<2>, th, 2, d);
    run("scalar FFMA ILP4", k_scalar<4>, th, 4, d);
    run("scalar FFMA ILP8", k_scalar<8>, th, 8, d);
    run("packed FFMA2 ILP1 (2 fma)", k_packed<1>, th, 2, d);
    run("packed FFMA2 ILP2 (4 fma)", k_packed<2>, th, 4, d);
    run("packed FFMA2 ILP4 (8 fma)", k_packed<4>, th, 8, d);
  }
  return 0;
}
