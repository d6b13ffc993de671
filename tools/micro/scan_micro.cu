// Micro-benchmark: the 528-window scan of exact_umma_kernel<0> in isolation (no TMEM / MMA / stagers): W warps per
// block, one block per SM, every thread scans `iters` rows.  Prints cycles per 128-row tile-equivalent per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scan_micro scan_micro.cu && ./scan_micro
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return (v > bv) || (v == bv && i < bi); }

template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32, 1) scan_kernel(const float* __restrict__ dots, const float* __restrict__ scale,
                                                             int iters, float* __restrict__ out_v, int* __restrict__ out_i,
                                                             long long* cycles) {
  __shared__ __align__(16) float sc[32 * 32];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sc[i] = scale[i];
  __syncthreads();
  float d0[32];
  for (int i = 0; i < 32; ++i) d0[i] = dots[(blockIdx.x * blockDim.x + threadIdx.x) * 32 + i];
  float accv = 0.f; int acci = 0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    float d[32], run[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) d[i] = d0[i] + (float)it * 1e-3f;
    float bv[8]; int bi[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { bv[k] = -INFINITY; bi[k] = 0x7fffffff; }
#pragma unroll
    for (int w = 1; w <= 32; ++w) {
      float scw[32];
#pragma unroll
      for (int s4 = 0; s4 + w <= 32; s4 += 4)
        *reinterpret_cast<float4*>(&scw[s4]) = *reinterpret_cast<const float4*>(&sc[(w - 1) * 32 + s4]);
#pragma unroll
      for (int s = 0; s + w <= 32; ++s) {
        run[s] = (w == 1) ? d[s] : __fadd_rn(run[s], d[s + w - 1]);
        const int pi = (w - 1) * 32 - ((w - 1) * (w - 2)) / 2 + s;
        const float v = __fmul_rn(run[s], scw[s]);
        if (v > bv[s & 7]) { bv[s & 7] = v; bi[s & 7] = pi; }
      }
    }
#pragma unroll
    for (int k = 1; k < 8; ++k)
      if (better(bv[k], bi[k], bv[0], bi[0])) { bv[0] = bv[k]; bi[0] = bi[k]; }
    accv += bv[0]; acci += bi[0];
  }
  const long long t1 = clock64();
  out_v[blockIdx.x * blockDim.x + threadIdx.x] = accv;
  out_i[blockIdx.x * blockDim.x + threadIdx.x] = acci;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int kWarps>
void run(const float* dots, const float* scale, float* ov, int* oi, long long* cyc, int iters) {
  scan_kernel<kWarps><<<148, kWarps * 32>>>(dots, scale, iters, ov, oi, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  // a scheduler owns kWarps / 4 warps; one warp-scan = 32 rows = a quarter of a 128-row tile
  printf("%d warps/SM: %.0f cycles per warp-scan, %.0f cycles per scheduler per 128-row tile (one scan warp each: x1; "
         "scans per scheduler in flight: %d)\n", kWarps, avg / iters, avg / iters / (kWarps / 4.0), kWarps / 4);
}

int main() {
  const int n = 148 * 16 * 32 * 32;
  float *dots, *scale, *ov; int* oi; long long* cyc;
  cudaMalloc(&dots, n * 4); cudaMalloc(&scale, 4096); cudaMalloc(&ov, 148 * 512 * 4); cudaMalloc(&oi, 148 * 512 * 4);
  cudaMalloc(&cyc, 148 * 8);
  float* h = new float[n];
  for (int i = 0; i < n; ++i) h[i] = (float)((i * 2654435761u) >> 8 & 0xffff) / 65536.f - 0.5f;
  cudaMemcpy(dots, h, n * 4, cudaMemcpyHostToDevice);
  for (int i = 0; i < 1024; ++i) h[i] = 1.0f / (1 + (i >> 5));
  cudaMemcpy(scale, h, 4096, cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 2; ++rep) {
    run<4>(dots, scale, ov, oi, cyc, 200);
    run<8>(dots, scale, ov, oi, cyc, 200);
    run<16>(dots, scale, ov, oi, cyc, 200);
  }
  return 0;
}
