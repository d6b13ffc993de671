// Micro-benchmark: the 528-window scan of exact_umma_kernel<0> in isolation (no TMEM / MMA / stagers): 4 warps per
// block (one per scheduler), one block per SM, every thread scans `iters` rows held in shared memory.
// Variants isolate what the scan pays for.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scan_micro scan_micro.cu && ./scan_micro
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return (v > bv) || (v == bv && i < bi); }

// kVariant 0: the kernel's scan (8 running maxima, value + index by compare/select)
//          1: no scale multiply / no scale loads (window sums only)
//          2: values only: running max by fmaxf (no index)
//          3: kernel's scan, scale row loaded one w ahead (software prefetch)
//          4: 16 running maxima instead of 8
template <int kVariant, int kWarps>
__global__ void __launch_bounds__(kWarps * 32, 1) scan_kernel(const float* __restrict__ dots, const float* __restrict__ scale,
                                                             int iters, float* __restrict__ out_v, int* __restrict__ out_i,
                                                             long long* cycles) {
  __shared__ __align__(16) float sc[32 * 32];
  __shared__ float sd[32][kWarps * 32 + 1];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sc[i] = scale[i];
  for (int i = 0; i < 32; ++i) sd[i][threadIdx.x] = dots[(blockIdx.x * blockDim.x + threadIdx.x) * 32 + i];
  __syncthreads();
  float accv = 0.f; int acci = 0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    asm volatile("" ::: "memory");   // like the kernel's mbarrier wait: the scale loads may not be hoisted out of the tile loop
    float d[32], run[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) d[i] = sd[i][threadIdx.x] + (float)it * 1e-3f;
    constexpr int kC = kVariant == 4 ? 16 : 8;
    float bv[kC]; int bi[kC];
#pragma unroll
    for (int k = 0; k < kC; ++k) { bv[k] = -INFINITY; bi[k] = 0x7fffffff; }
    float nxt[32];
    if (kVariant == 3) {
#pragma unroll
      for (int s4 = 0; s4 < 32; s4 += 4) *reinterpret_cast<float4*>(&nxt[s4]) = *reinterpret_cast<const float4*>(&sc[s4]);
    }
#pragma unroll
    for (int w = 1; w <= 32; ++w) {
      float scw[32];
      if (kVariant == 3) {
#pragma unroll
        for (int s = 0; s + w <= 32; ++s) scw[s] = nxt[s];
        if (w < 32) {
#pragma unroll
          for (int s4 = 0; s4 + w + 1 <= 32; s4 += 4)
            *reinterpret_cast<float4*>(&nxt[s4]) = *reinterpret_cast<const float4*>(&sc[w * 32 + s4]);
        }
      } else if (kVariant != 1) {
#pragma unroll
        for (int s4 = 0; s4 + w <= 32; s4 += 4)
          *reinterpret_cast<float4*>(&scw[s4]) = *reinterpret_cast<const float4*>(&sc[(w - 1) * 32 + s4]);
      }
#pragma unroll
      for (int s = 0; s + w <= 32; ++s) {
        run[s] = (w == 1) ? d[s] : __fadd_rn(run[s], d[s + w - 1]);
        const int pi = (w - 1) * 32 - ((w - 1) * (w - 2)) / 2 + s;
        const float v = kVariant == 1 ? run[s] : __fmul_rn(run[s], scw[s]);
        if (kVariant == 2) bv[s & 7] = fmaxf(bv[s & 7], v);
        else if (v > bv[s % kC]) { bv[s % kC] = v; bi[s % kC] = pi; }
      }
    }
#pragma unroll
    for (int k = 1; k < kC; ++k)
      if (better(bv[k], bi[k], bv[0], bi[0])) { bv[0] = bv[k]; bi[0] = bi[k]; }
    accv += bv[0]; acci += bi[0];
  }
  const long long t1 = clock64();
  out_v[blockIdx.x * blockDim.x + threadIdx.x] = accv;
  out_i[blockIdx.x * blockDim.x + threadIdx.x] = acci;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// Start-split scan: the windows of one row are shared by two threads by START position: [0, kSplit) and [kSplit, 32)
// (275 + 253 windows at kSplit = 10).  No catch-up work, and each thread holds only its own run[] (and the tail of d[]).
template <int kLo, int kHi>
__device__ __forceinline__ void scan_range(const float (*sd)[8 * 32 + 1], const float* sc, int it, float& obv, int& obi) {
  float d[32], run[32];
#pragma unroll
  for (int i = kLo; i < 32; ++i) d[i] = sd[i][threadIdx.x & 127] + (float)it * 1e-3f;
  float bv[8]; int bi[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { bv[k] = -INFINITY; bi[k] = 0x7fffffff; }
#pragma unroll
  for (int w = 1; w <= 32 - kLo; ++w) {
    float scw[32];
#pragma unroll
    for (int s4 = (kLo / 4) * 4; s4 + w <= 32 && s4 < kHi; s4 += 4)
      *reinterpret_cast<float4*>(&scw[s4]) = *reinterpret_cast<const float4*>(&sc[(w - 1) * 32 + s4]);
#pragma unroll
    for (int s = kLo; s < kHi && s + w <= 32; ++s) {
      run[s] = (w == 1) ? d[s] : __fadd_rn(run[s], d[s + w - 1]);
      const int pi = (w - 1) * 32 - ((w - 1) * (w - 2)) / 2 + s;
      const float v = __fmul_rn(run[s], scw[s]);
      if (v > bv[s & 7]) { bv[s & 7] = v; bi[s & 7] = pi; }
    }
  }
#pragma unroll
  for (int k = 1; k < 8; ++k)
    if (better(bv[k], bi[k], bv[0], bi[0])) { bv[0] = bv[k]; bi[0] = bi[k]; }
  obv = bv[0]; obi = bi[0];
}

// Start-major scan: one chain per window start s (running sum over w, local first maximum by strict >), kIl starts
// interleaved for ILP, scale read from a TRANSPOSED table sc_t[s][w - 1] in 4-wide loads consumed immediately; chains
// merge into the global best with the full (value desc, proposal asc) rule.  Small register footprint.
template <int kIl>
__global__ void __launch_bounds__(128, 1) scan_smajor_kernel(const float* __restrict__ dots, const float* __restrict__ scale,
                                                            int iters, float* __restrict__ out_v, int* __restrict__ out_i,
                                                            long long* cycles) {
  __shared__ __align__(16) float sct[32 * 32];
  __shared__ float sd[32][4 * 32 + 1];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sct[(i & 31) * 32 + (i >> 5)] = scale[i];   // [s][w - 1]
  for (int i = 0; i < 32; ++i) sd[i][threadIdx.x] = dots[(blockIdx.x * blockDim.x + threadIdx.x) * 32 + i];
  __syncthreads();
  float accv = 0.f; int acci = 0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    asm volatile("" ::: "memory");
    float d[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) d[i] = sd[i][threadIdx.x] + (float)it * 1e-3f;
    float gv = -INFINITY; int gi = 0x7fffffff;
#pragma unroll
    for (int s0 = 0; s0 < 32; s0 += kIl) {
      float run[kIl], bv[kIl]; int bi[kIl];
#pragma unroll
      for (int c = 0; c < kIl; ++c) { bv[c] = -INFINITY; bi[c] = 0x7fffffff; }
#pragma unroll
      for (int w0 = 0; w0 < 32 - s0; w0 += 4) {          // window lengths w0 + 1 .. w0 + 4
        float4 sc4[kIl];
#pragma unroll
        for (int c = 0; c < kIl; ++c)
          if (s0 + c + w0 < 32) sc4[c] = *reinterpret_cast<const float4*>(&sct[(s0 + c) * 32 + w0]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int w = w0 + j + 1;
#pragma unroll
          for (int c = 0; c < kIl; ++c) {
            const int s = s0 + c;
            if (s + w <= 32) {
              run[c] = (w == 1) ? d[s] : __fadd_rn(run[c], d[s + w - 1]);
              const int pi = (w - 1) * 32 - ((w - 1) * (w - 2)) / 2 + s;
              const float scl = j == 0 ? sc4[c].x : (j == 1 ? sc4[c].y : (j == 2 ? sc4[c].z : sc4[c].w));
              const float v = __fmul_rn(run[c], scl);
              if (v > bv[c]) { bv[c] = v; bi[c] = pi; }
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < kIl; ++c)
        if (better(bv[c], bi[c], gv, gi)) { gv = bv[c]; gi = bi[c]; }
    }
    accv += gv; acci += gi;
  }
  const long long t1 = clock64();
  out_v[blockIdx.x * blockDim.x + threadIdx.x] = accv;
  out_i[blockIdx.x * blockDim.x + threadIdx.x] = acci;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int kIl>
void run_smajor(const float* dots, const float* scale, float* ov, int* oi, long long* cyc, int iters) {
  for (int rep = 0; rep < 2; ++rep) {
    scan_smajor_kernel<kIl><<<148, 128>>>(dots, scale, iters, ov, oi, cyc);
    cudaDeviceSynchronize();
  }
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("start-major scan, %d starts interleaved, 1 warp per scheduler:  %6.0f cycles per scheduler per 128-row tile\n", kIl, avg / iters);
}

template <int kSplit>
__global__ void __launch_bounds__(256, 1) scan_split_kernel(const float* __restrict__ dots, const float* __restrict__ scale,
                                                           int iters, float* __restrict__ out_v, int* __restrict__ out_i,
                                                           long long* cycles) {
  __shared__ __align__(16) float sc[32 * 32];
  __shared__ float sd[32][8 * 32 + 1];
  __shared__ float xv[128]; __shared__ int xi[128];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sc[i] = scale[i];
  if (threadIdx.x < 128)
    for (int i = 0; i < 32; ++i) sd[i][threadIdx.x] = dots[(blockIdx.x * 128 + threadIdx.x) * 32 + i];
  __syncthreads();
  float accv = 0.f; int acci = 0;
  const int part = threadIdx.x >> 7, quarter = (threadIdx.x >> 5) & 3;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    asm volatile("" ::: "memory");
    float bv; int bi;
    if (part == 0) scan_range<0, kSplit>(sd, sc, it, bv, bi);
    else scan_range<kSplit, 32>(sd, sc, it, bv, bi);
    if (part == 1) { xv[threadIdx.x & 127] = bv; xi[threadIdx.x & 127] = bi; }
    asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
    if (part == 0) {
      const float ov = xv[threadIdx.x]; const int oi = xi[threadIdx.x];
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
      accv += bv; acci += bi;
    }
    asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
  }
  const long long t1 = clock64();
  if (part == 0) { out_v[blockIdx.x * 128 + threadIdx.x] = accv; out_i[blockIdx.x * 128 + threadIdx.x] = acci; }
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int kSplit>
void run_split(const float* dots, const float* scale, float* ov, int* oi, long long* cyc, int iters) {
  for (int rep = 0; rep < 2; ++rep) {
    scan_split_kernel<kSplit><<<148, 256>>>(dots, scale, iters, ov, oi, cyc);
    cudaDeviceSynchronize();
  }
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("start-split at s = %2d, 2 warps per scheduler:            %6.0f cycles per scheduler per 128-row tile\n", kSplit, avg / iters);
}

template <int kVariant, int kWarps>
void run(const char* name, const float* dots, const float* scale, float* ov, int* oi, long long* cyc, int iters) {
  for (int rep = 0; rep < 2; ++rep) {
    scan_kernel<kVariant, kWarps><<<148, kWarps * 32>>>(dots, scale, iters, ov, oi, cyc);
    cudaDeviceSynchronize();
  }
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("%-44s %d warps/SM: %6.0f cycles per warp-scan, %6.0f per scheduler per 128-row tile\n", name, kWarps,
         avg / iters, avg / iters * (kWarps / 4.0) / (kWarps / 4.0) / 1.0 * 1.0 / (kWarps / 4.0) * (kWarps / 4.0) / (kWarps / 4.0));
}

int main() {
  const int n = 148 * 8 * 32 * 32;
  float *dots, *scale, *ov; int* oi; long long* cyc;
  cudaMalloc(&dots, n * 4); cudaMalloc(&scale, 4096); cudaMalloc(&ov, 148 * 256 * 4); cudaMalloc(&oi, 148 * 256 * 4);
  cudaMalloc(&cyc, 148 * 8);
  float* h = new float[n];
  for (int i = 0; i < n; ++i) h[i] = (float)((i * 2654435761u) >> 8 & 0xffff) / 65536.f - 0.5f;
  cudaMemcpy(dots, h, n * 4, cudaMemcpyHostToDevice);
  for (int i = 0; i < 1024; ++i) h[i] = 1.0f / (1 + (i >> 5));
  cudaMemcpy(scale, h, 4096, cudaMemcpyHostToDevice);
  run<0, 4>("0 kernel scan", dots, scale, ov, oi, cyc, 200);
  run<1, 4>("1 no scale", dots, scale, ov, oi, cyc, 200);
  run<2, 4>("2 values only (fmaxf)", dots, scale, ov, oi, cyc, 200);
  run<3, 4>("3 scale row prefetched one w ahead", dots, scale, ov, oi, cyc, 200);
  run<4, 4>("4 16 running maxima", dots, scale, ov, oi, cyc, 200);
  run_smajor<4>(dots, scale, ov, oi, cyc, 200);
  run_smajor<8>(dots, scale, ov, oi, cyc, 200);
  run_smajor<16>(dots, scale, ov, oi, cyc, 200);
  run_split<10>(dots, scale, ov, oi, cyc, 200);
  run_split<8>(dots, scale, ov, oi, cyc, 200);
  run_split<12>(dots, scale, ov, oi, cyc, 200);
  run<0, 8>("0 kernel scan", dots, scale, ov, oi, cyc, 200);
  run<2, 8>("2 values only (fmaxf)", dots, scale, ov, oi, cyc, 200);
  return 0;
}
