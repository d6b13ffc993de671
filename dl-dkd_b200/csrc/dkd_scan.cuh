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

// ---- two-phase scan (T = 32 only) ---------------------------------------------------------------------------------
// phase 1: maximum VALUE over all windows and the first row (window length) attaining it.
template <bool kPacked>
__device__ __forceinline__ void scan_phase1(const float (&d)[32], const float* __restrict__ sc, float& best, int& bw) {
  best = -INFINITY;
  bw = 1;
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
}

// exact value of window (length w, start s): the sequential sum of phase 1, read back from the thread's column of the
// shared-memory scratch (dynamic start); all 32 loads are issued up front, the adds are predicated on i < w
__device__ __forceinline__ float window_value(const float* dcol, int dstride, const float* __restrict__ sc, int w, int s) {
  float x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = dcol[min(s + i, 31) * dstride];
  float r = x[0];
#pragma unroll
  for (int i = 1; i < 32; ++i) r = (i < w) ? __fadd_rn(r, x[i]) : r;
  return __fmul_rn(r, sc[(w - 1) * 32 + s]);
}

// phase 2: the first start s of row bw whose value equals `best`.  Straight-line code with static register indices:
// a sliding (inexact) window sum flags the starts whose value can equal the maximum (bit mask), only those are
// recomputed exactly (window_value) — normally one per thread, all threads of the warp in lock step.
__device__ __forceinline__ int scan_phase2(const float (&d)[32], const float* __restrict__ sc, const float* dcol, int dstride,
                                           float best, int bw, float mass) {
  float a = d[0];
#pragma unroll
  for (int i = 1; i < 32; ++i) a = (i < bw) ? __fadd_rn(a, d[i]) : a;       // start 0, exact
  const float* scr = sc + (bw - 1) * 32;
  const float tol = 1.0e-5f * mass;                      // >= 93 roundings of 2^-24 * mass (sliding + sequential sums)
  uint32_t cand = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float t[8], scl[8];
    *reinterpret_cast<float4*>(&scl[0]) = *reinterpret_cast<const float4*>(&scr[8 * c]);
    *reinterpret_cast<float4*>(&scl[4]) = *reinterpret_cast<const float4*>(&scr[8 * c + 4]);
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = dcol[min(8 * c + j + bw, 31) * dstride];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int s = 8 * c + j;
      const bool ok = (s + bw <= 32) && (fabsf(__fmul_rn(a, scl[j]) - best) <= tol * fabsf(scl[j]));
      cand |= (ok ? 1u : 0u) << s;
      a = __fadd_rn(__fadd_rn(a, -d[s]), t[j]);
    }
  }
  int s_found = -1;
  while (cand) {
    const int s = __ffs(cand) - 1;
    cand &= cand - 1;
    if (window_value(dcol, dstride, sc, bw, s) == best) { s_found = s; break; }
  }
  if (s_found < 0) {                                      // unreachable unless the filter bound is violated: full search
    for (int s = 0; s + bw <= 32 && s_found < 0; ++s)
      if (window_value(dcol, dstride, sc, bw, s) == best) s_found = s;
    if (s_found < 0) s_found = 0;
  }
  return s_found;
}

// dcol: this thread's column of a shared-memory scratch of 32 rows (dcol[i * dstride] = d[i]); it is written here and
// only read by the same thread, so no barrier is needed.
template <bool kPacked>
__device__ __forceinline__ void window_scan_v2(const float (&d)[32], const float* __restrict__ sc, float* dcol, int dstride,
                                               float& out_v, int& out_i) {
  float mass = 0.f;                                       // bounds every partial window sum (phase 2 filter tolerance)
#pragma unroll
  for (int i = 0; i < 32; ++i) { dcol[i * dstride] = d[i]; mass += fabsf(d[i]); }
  float best;
  int bw;
  scan_phase1<kPacked>(d, sc, best, bw);
  const int s_found = scan_phase2(d, sc, dcol, dstride, best, bw, mass);
  out_v = best;
  out_i = (bw - 1) * 32 - ((bw - 1) * (bw - 2)) / 2 + s_found;
}

// Known key clip (rescoring of candidates whose key clip the approximate pass already fixed): the exact value of window
// `key` (sequential sum) is checked against the value-only maximum of phase 1.  When the key's value IS the maximum and
// its row is the first row attaining it, the key is confirmed at the cost of phase 1 alone; otherwise (a wrong key: the
// approximate pass's argmax gap was misleading) phase 2 resolves the true first argmax.  Result = the full scan's.
// (Residual: an exact fp32 tie with an EARLIER start of the same row keeps the given key — a documented exact-score tie.)
template <bool kPacked>
__device__ __forceinline__ void window_scan_known(const float (&d)[32], const float* __restrict__ sc, float* dcol, int dstride,
                                                  int key, float& out_v, int& out_i) {
#pragma unroll
  for (int i = 0; i < 32; ++i) dcol[i * dstride] = d[i];
  // p(w, s) = (w-1) * 32 - (w-1)(w-2)/2 + s: invert for w by a 5-step search on the row starts
  key = min(max(key, 0), 527);
  int w = 1;
#pragma unroll
  for (int step = 16; step > 0; step >>= 1) {
    const int wt = w + step;
    if (wt <= 32 && (wt - 1) * 32 - ((wt - 1) * (wt - 2)) / 2 <= key) w = wt;
  }
  const int s = key - ((w - 1) * 32 - ((w - 1) * (w - 2)) / 2);
  const float vk = window_value(dcol, dstride, sc, w, s);
  float best;
  int bw;
  scan_phase1<kPacked>(d, sc, best, bw);
  out_v = best;
  out_i = key;
  if (!(vk == best && bw == w)) {
    float mass = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) mass += fabsf(d[i]);
    const int s_found = scan_phase2(d, sc, dcol, dstride, best, bw, mass);
    out_i = (bw - 1) * 32 - ((bw - 1) * (bw - 2)) / 2 + s_found;
  }
}

}  // namespace dkd
