// The clip-window scan of the exact (fp32-grade) clip-scale kernel, as device functions (shared by
// dkd_exact_umma.cu and tools/micro/scan2_micro.cu).
//
//   in : d[0..31]  the query's 32 per-clip dot products (thread = one (query, video) pair)
//        sc        shared memory, prop_scale of the video as [w - 1][s] (32 x 32 floats, warp-uniform reads)
//   out: max over the T(T+1)/2 windows (length w = 1..T, start s) of  fl(fl(d[s] + ... + d[s+w-1]) * scale[w][s])
//        with the window sum accumulated SEQUENTIALLY in s..s+w-1 order, and the FIRST proposal index
//        p(w, s) = (w-1) T - (w-1)(w-2)/2 + s attaining it (torch.max tie rule).
//
// v1 (round 1): 8 independent running (value, index) maxima, compare + 2 selects per window: 3 ALU-pipe + 2 FMA-pipe
//     instructions per window, ~7.4 k cycles per warp-scan measured.
// v2 (round 2): two phases.
//     phase 1  values only.  Row maximum m_w over the starts by 3-input max (0.5 ALU instruction per window), the
//              first best row tracked once per row.  With kPacked the window sums and products are formed two at a
//              time (add.rn.f32x2 / mul.rn.f32x2: IEEE results per lane, half the FMA-pipe issue slots).
//     phase 2  the first start s of row w* whose value equals the maximum: a sliding window sum (inexact, 2 adds per
//              step) filters the starts whose value can equal the maximum, the sequential sum of phase 1 is recomputed
//              for those only (from a per-thread column of shared memory, dynamic indexing) and compared for equality.
//     Same values, same tie rule => bit-identical results to v1 (checked row by row in tools/micro/scan2_micro.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dkd {

__device__ __forceinline__ bool scan_better(float v, int i, float bv, int bi) { return (v > bv) || (v == bv && i < bi); }

template <bool kT32>
__device__ __forceinline__ void window_scan_v1(const float (&d)[32], const float* __restrict__ sc, int T, float& out_v,
                                               int& out_i) {
  float run[32];
  float bv[8];
  int bi[8];
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
      if (kT32 || s + w <= T) {
        const int pi = kT32 ? ((w - 1) * 32 - ((w - 1) * (w - 2)) / 2 + s) : ((w - 1) * T - ((w - 1) * (w - 2)) / 2 + s);
        const float v = __fmul_rn(run[s], scw[s]);
        if (v > bv[s & 7]) { bv[s & 7] = v; bi[s & 7] = pi; }
      }
    }
  }
#pragma unroll
  for (int k = 1; k < 8; ++k)
    if (scan_better(bv[k], bi[k], bv[0], bi[0])) { bv[0] = bv[k]; bi[0] = bi[k]; }
  out_v = bv[0];
  out_i = bi[0];
}

__device__ __forceinline__ float max3f(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ uint64_t pack2f(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2f(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t add2f(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2f(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// T = 32 only.  dcol: this thread's column of a shared-memory scratch of 32 rows (dcol[i * dstride] = d[i]); it is
// written here and only read by the same thread, so no barrier is needed.
template <bool kPacked>
__device__ __forceinline__ void window_scan_v2(const float (&d)[32], const float* __restrict__ sc, float* dcol, int dstride,
                                               float& out_v, int& out_i) {
  // |d| mass: bounds every partial window sum, hence the error of the sliding sums of phase 2
  float mass = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) { dcol[i * dstride] = d[i]; mass += fabsf(d[i]); }
  float best = -INFINITY;
  int bw = 1;
  if (!kPacked) {
    float run[32];
#pragma unroll
    for (int w = 1; w <= 32; ++w) {
      float scw[32];
#pragma unroll
      for (int s4 = 0; s4 + w <= 32; s4 += 4)
        *reinterpret_cast<float4*>(&scw[s4]) = *reinterpret_cast<const float4*>(&sc[(w - 1) * 32 + s4]);
      float m0 = -INFINITY, m1 = -INFINITY;            // two chains per row, merged once
#pragma unroll
      for (int s = 0; s + w <= 32; s += 2) {
        run[s] = (w == 1) ? d[s] : __fadd_rn(run[s], d[s + w - 1]);
        const float v0 = __fmul_rn(run[s], scw[s]);
        if (s + 1 + w <= 32) {
          run[s + 1] = (w == 1) ? d[s + 1] : __fadd_rn(run[s + 1], d[s + w]);
          const float v1 = __fmul_rn(run[s + 1], scw[s + 1]);
          if ((s >> 1) & 1) m1 = max3f(m1, v0, v1); else m0 = max3f(m0, v0, v1);
        } else {
          if ((s >> 1) & 1) m1 = fmaxf(m1, v0); else m0 = fmaxf(m0, v0);
        }
      }
      const float mw = fmaxf(m0, m1);
      if (mw > best) { best = mw; bw = w; }
    }
  } else {
    // pairs (run[2k], run[2k+1]); the addend pair of step w is (d[2k+w-1], d[2k+w]) — even- or odd-aligned in d
    uint64_t run2[16];
    uint64_t de[16], dod[16];                            // (d[2j], d[2j+1]) and (d[2j+1], d[2j+2]); d[32] := 0
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      de[j] = pack2f(d[2 * j], d[2 * j + 1]);
      dod[j] = pack2f(d[2 * j + 1], j < 15 ? d[2 * j + 2] : 0.f);
    }
#pragma unroll
    for (int w = 1; w <= 32; ++w) {
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int s4 = 0; s4 + w <= 32; s4 += 4) {
        const ulonglong2 sc2 = *reinterpret_cast<const ulonglong2*>(&sc[(w - 1) * 32 + s4]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int s = s4 + 2 * h;
          if (s + w <= 32) {
            const int k = s >> 1;
            const int j = s + w - 1;                     // first clip added in this step
            if (w == 1) run2[k] = de[k];
            else run2[k] = add2f(run2[k], (j & 1) ? dod[j >> 1] : de[j >> 1]);
            float v0, v1;
            unpack2f(mul2f(run2[k], h == 0 ? sc2.x : sc2.y), v0, v1);
            if (s + 1 + w <= 32) { if (k & 1) m1 = max3f(m1, v0, v1); else m0 = max3f(m0, v0, v1); }
            else { if (k & 1) m1 = fmaxf(m1, v0); else m0 = fmaxf(m0, v0); }
          }
        }
      }
      const float mw = fmaxf(m0, m1);
      if (mw > best) { best = mw; bw = w; }
    }
  }
  // ---- phase 2: first start of row bw whose value equals `best`
  const float* scr = sc + (bw - 1) * 32;
  const float tol = 1.0e-5f * mass;                      // >= 93 roundings of 2^-24 * mass (sliding + sequential sums)
  float a = dcol[0];
  for (int i = 1; i < bw; ++i) a = __fadd_rn(a, dcol[i * dstride]);       // exact value of start 0
  int s_found = -1;
  const int ns = 33 - bw;
  for (int s = 0; s < ns; ++s) {
    const float scl = scr[s];
    if (fabsf(__fmul_rn(a, scl) - best) <= tol * fabsf(scl)) {
      float r = dcol[s * dstride];
      for (int i = 1; i < bw; ++i) r = __fadd_rn(r, dcol[(s + i) * dstride]);
      if (__fmul_rn(r, scl) == best) { s_found = s; break; }
    }
    if (s + 1 < ns) a = __fadd_rn(__fadd_rn(a, -dcol[s * dstride]), dcol[(s + bw) * dstride]);
  }
  if (s_found < 0) {                                      // unreachable unless the filter bound is violated: full search
    for (int s = 0; s < ns && s_found < 0; ++s) {
      float r = dcol[s * dstride];
      for (int i = 1; i < bw; ++i) r = __fadd_rn(r, dcol[(s + i) * dstride]);
      if (__fmul_rn(r, scr[s]) == best) s_found = s;
    }
    if (s_found < 0) s_found = 0;
  }
  out_v = best;
  out_i = (bw - 1) * 32 - ((bw - 1) * (bw - 2)) / 2 + s_found;
}

}  // namespace dkd
