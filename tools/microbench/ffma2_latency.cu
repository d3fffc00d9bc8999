// Dependent-issue latency and pipe occupancy of FFMA against FFMA2 (fma.rn.f32x2) on sm_100a, measured with clock64() on ONE warp
// per SM (latency) and on 1..4 warps per SM sub-partition (pipe occupancy): is a packed FMA one pipe cycle or two?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_latency ffma2_latency.cu && ./ffma2_latency
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
template <int ILP, bool PACKED>
__global__ void k(float* out, long long* cyc, int iters) {
  float x[ILP]; u64 y[ILP];
  const float a = 0.999f, b = 0.001f;
  u64 a2, b2; asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a)); asm("mov.b64 %0, {%1, %1};" : "=l"(b2) : "f"(b));
#pragma unroll
  for (int i = 0; i < ILP; i++) { x[i] = threadIdx.x + i; asm("mov.b64 %0, {%1, %1};" : "=l"(y[i]) : "f"(x[i])); }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 16; r++)
#pragma unroll
      for (int i = 0; i < ILP; i++) { if (PACKED) y[i] = fma2(y[i], a2, b2); else x[i] = fma1(x[i], a, b); }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(y[i])); s += x[i] + lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP, bool PACKED>
void run(const char* name, int threads, float* d, long long* dc) {
  const int iters = 4000;
  k<ILP, PACKED><<<148, threads>>>(d, dc, 100);
  k<ILP, PACKED><<<148, threads>>>(d, dc, iters);
  long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  const double per = (double)c / (iters * 16.0 * ILP);
  printf("%-22s warps/SMSP %4.2f  ILP %d : %6.2f cycles per instruction per warp, %5.2f pipe-instr/cycle/SMSP\n", name, threads / 128.0, ILP, per * ILP,
         (threads / 128.0) / per);
}
int main() {
  float* d; long long* dc; cudaMalloc(&d, 148 * 1024 * 4); cudaMalloc(&dc, 8);
  for (int th : {32, 128, 256, 384, 512}) {
    run<1, false>("FFMA  dependent", th, d, dc); run<1, true>("FFMA2 dependent", th, d, dc);
    run<2, false>("FFMA  2 chains", th, d, dc); run<2, true>("FFMA2 2 chains", th, d, dc);
    run<4, false>("FFMA  4 chains", th, d, dc); run<4, true>("FFMA2 4 chains", th, d, dc);
  }
  return 0;
}
