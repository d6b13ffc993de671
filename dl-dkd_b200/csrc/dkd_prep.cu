// Corpus-side preparation kernels (query independent, run once per corpus / shard):
// row normalisation + bf16 cast, clip downsample, clip-proposal builder, key-clip attention table.
// All are HBM-bound streaming kernels; arithmetic is fp32.
#include <cuda_fp16.h>

#include "dkd_common.cuh"

namespace dkd {

// ------------------------------------------------------------------------------------------
// L2-normalise rows.  One warp per row, grid-stride.  F.normalize semantics
// (method/model.py:318-319): x / max(||x||_2, eps) with a true division.
__global__ void normalize_rows_kernel(const float* __restrict__ x, int64_t rows, int D, float eps,
                                      float* __restrict__ of32, __nv_bfloat16* __restrict__ obf,
                                      __half* __restrict__ ohf, int64_t rows_pad) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = warp0; r < rows_pad; r += nwarp) {
    if (r >= rows) {  // zero padding rows
      for (int d = lane; d < D; d += 32) {
        if (of32) of32[r * D + d] = 0.f;
        if (obf) obf[r * D + d] = __float2bfloat16(0.f);
        if (ohf) ohf[r * D + d] = __float2half(0.f);
      }
      continue;
    }
    const float* xr = x + r * D;
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) {
      float v = xr[d];
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    const float denom = fmaxf(sqrtf(ss), eps);
    for (int d = lane; d < D; d += 32) {
      float v = __fdiv_rn(xr[d], denom);
      if (of32) of32[r * D + d] = v;
      if (obf) obf[r * D + d] = __float2bfloat16(v);
      if (ohf) ohf[r * D + d] = __float2half_rn(v);
    }
  }
}

// ------------------------------------------------------------------------------------------
// average_to_fixed_length (method/data_provider.py:30-50) on the device.
// idxs[i] = min(round_half_even(i / T * n), n - 1) in fp32; clip i = mean(frames[s:e]) or frames[s].
__global__ void downsample_clips_kernel(const float* __restrict__ frames,
                                        const int32_t* __restrict__ lengths, int L, int D, int T,
                                        float* __restrict__ clips) {
  const int n = blockIdx.x;
  int len = lengths[n];
  len = len < 1 ? 1 : (len > L ? L : len);
  const float* f = frames + (int64_t)n * L * D;
  float* c = clips + (int64_t)n * T * D;
  for (int i = 0; i < T; ++i) {
    // torch: arange(0, T+1, 1.0) / T * n, rounded half-to-even, clamped to n-1
    float fs = __fmul_rn(__fdiv_rn((float)i, (float)T), (float)len);
    float fe = __fmul_rn(__fdiv_rn((float)(i + 1), (float)T), (float)len);
    int s = min((int)rintf(fs), len - 1);
    int e = min((int)rintf(fe), len - 1);
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      float v;
      if (s < e) {
        float acc = f[(int64_t)s * D + d];
        for (int r = s + 1; r < e; ++r) acc = __fadd_rn(acc, f[(int64_t)r * D + d]);
        v = __fdiv_rn(acc, (float)(e - s));
      } else {
        v = f[(int64_t)s * D + d];
      }
      c[(int64_t)i * D + d] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Clip-proposal builder.  One block per video, 16 warps; warp j owns window starts j and T-1-j
// (balanced: T+1 windows per warp).  The T x D clip tile lives in shared memory (48 KB at
// T=32, D=384); every window sum is a running fp32 sum over the window (the AvgPool1d order),
// each lane holding D/32 features as bf16x2-friendly pairs.  Per window: one warp reduction
// for the norm, one coalesced bf16 row store.
// kMeans = false (the production path: bf16 rows + scale only) never forms the mean: with
// scale = 1 / (w * ||sum / w||) = 1 / ||sum|| the normalised row is sum * scale, so a window costs one
// division instead of 4 per feature pair (the IEEE divisions made the first version issue bound at
// 1.7 TB/s).  kMeans = true also writes the un-normalised fp32 means with the exact sum / w division
// (tests, small shapes).
template <int kPairs, bool kMeans, bool kHalf = false>  // D = 64 * kPairs; kHalf: rows as IEEE half instead of bf16
__global__ void __launch_bounds__(512)
build_proposals_kernel(const float* __restrict__ clips, int T, int D,
                       __nv_bfloat16* __restrict__ prop_bf16, float* __restrict__ prop_scale,
                       float* __restrict__ prop_f32) {
  extern __shared__ float sclips[];  // T * D
  const int n = blockIdx.x;
  const int P = T * (T + 1) / 2;
  const float* c = clips + (int64_t)n * T * D;
  for (int i = threadIdx.x * 4; i < T * D; i += blockDim.x * 4) {
    *reinterpret_cast<float4*>(&sclips[i]) = *reinterpret_cast<const float4*>(&c[i]);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  // warp k owns starts k and T-1-k: T+1 windows per warp, balanced.
  for (int sidx = warp; sidx < 2 * ((T + 1) / 2); sidx += nwarps) {
    const int half = (T + 1) / 2;
    const int s = (sidx < half) ? sidx : (T - 1 - (sidx - half));
    if (sidx >= half && s == sidx - half) continue;  // odd T: middle start already done
    float acc[2 * kPairs];
#pragma unroll
    for (int j = 0; j < 2 * kPairs; ++j) acc[j] = 0.f;
    for (int w = 1; w <= T - s; ++w) {
      const float* row = &sclips[(s + w - 1) * D];
      float ss = 0.f;
      const float fw = (float)w;
#pragma unroll
      for (int j = 0; j < kPairs; ++j) {
        float2 v = *reinterpret_cast<const float2*>(&row[64 * j + 2 * lane]);
        acc[2 * j] = (w == 1) ? v.x : __fadd_rn(acc[2 * j], v.x);
        acc[2 * j + 1] = (w == 1) ? v.y : __fadd_rn(acc[2 * j + 1], v.y);
        ss = fmaf(acc[2 * j], acc[2 * j], ss);
        ss = fmaf(acc[2 * j + 1], acc[2 * j + 1], ss);
      }
      ss = warp_sum(ss);
      // ||mean|| = ||sum|| / w, clamped at 1e-12 like F.normalize:  scale = 1 / (w * max(||sum|| / w, 1e-12))
      const float scale = __fdiv_rn(1.0f, fmaxf(sqrtf(ss), __fmul_rn(fw, 1e-12f)));
      const int p = prop_index(w, s, T);
      const int64_t ro = ((int64_t)n * P + p) * D;
      if (prop_scale && lane == 0) prop_scale[(int64_t)n * P + p] = scale;
#pragma unroll
      for (int j = 0; j < kPairs; ++j) {
        if (kMeans && prop_f32) {
          *reinterpret_cast<float2*>(&prop_f32[ro + 64 * j + 2 * lane]) =
              make_float2(__fdiv_rn(acc[2 * j], fw), __fdiv_rn(acc[2 * j + 1], fw));
        }
        if (prop_bf16) {
          const float a = __fmul_rn(acc[2 * j], scale), b = __fmul_rn(acc[2 * j + 1], scale);
          if (kHalf) {
            *reinterpret_cast<__half2*>(&prop_bf16[ro + 64 * j + 2 * lane]) = __floats2half2_rn(a, b);
          } else {
            *reinterpret_cast<__nv_bfloat162*>(&prop_bf16[ro + 64 * j + 2 * lane]) = __floats2bfloat162_rn(a, b);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Key-clip attention table.  Block = (video n, chunk of kPC proposals), 256 threads.
//  phase 1: logits[p][l] = (sum_{i in window p} E[n][l][i]) / w           (E tile in smem)
//  phase 2: softmax over valid frames (one warp per proposal row, accurate expf)
//  phase 3: g[p][:] = sum_l a[p][l] * val[n][l][:]  — val streamed through smem in 64-feature
//           chunks, each thread keeps 3 proposals x 4 features x (D/64) chunks in registers
//  phase 4: row norm (16-lane reduction), write fp32 and fp16 rows.
constexpr int kPC = 48;   // proposals per block (528 = 11 * 48)
constexpr int kLmax = 128;

template <int kChunks>  // D = 64 * kChunks
__global__ void __launch_bounds__(256)
frame_attn_table_kernel(const float* __restrict__ E, const float* __restrict__ val,
                        const int32_t* __restrict__ lengths, int L, int T, int D,
                        float* __restrict__ table_f32, __half* __restrict__ table_f16) {
  extern __shared__ __align__(16) float smem_tab[];
  float (*sV)[64] = reinterpret_cast<float (*)[64]>(smem_tab);                        // kLmax x 64
  float (*sE)[33] = reinterpret_cast<float (*)[33]>(smem_tab + kLmax * 64);           // kLmax x 33
  float (*sA)[kLmax + 1] = reinterpret_cast<float (*)[kLmax + 1]>(smem_tab + kLmax * 64 + kLmax * 33);
  // chunk index fastest: the blocks of one video are launched together and share its val / E rows through L2
  // (video-major launch order re-read val from DRAM once per chunk: 11 x the algorithmic bytes).
  const int P = T * (T + 1) / 2;
  const int nchunk = (P + kPC - 1) / kPC;
  const int n = blockIdx.x / nchunk;
  const int p0 = (blockIdx.x % nchunk) * kPC;
  int len = lengths[n];
  len = len < 1 ? 1 : (len > L ? L : len);
  const int tid = threadIdx.x;

  for (int i = tid; i < kLmax * 32; i += 256) {
    int l = i >> 5, c = i & 31;
    sE[l][c] = (l < L && c < T) ? E[((int64_t)n * L + l) * T + c] : 0.f;
  }
  __syncthreads();
  // phase 1
  for (int i = tid; i < kPC * kLmax; i += 256) {
    const int pl = i / kLmax, l = i % kLmax;
    const int p = p0 + pl;
    float v = 0.f;
    if (p < P && l < len) {
      // invert p -> (w, s): windows of length w occupy [off_w, off_w + T - w + 1)
      int w = 1, off = 0;
      while (off + (T - w + 1) <= p) { off += T - w + 1; ++w; }
      const int s = p - off;
      float acc = sE[l][s];
      for (int k = 1; k < w; ++k) acc = __fadd_rn(acc, sE[l][s + k]);
      v = __fdiv_rn(acc, (float)w);
    }
    sA[pl][l] = v;
  }
  __syncthreads();
  // phase 2
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int pl = warp; pl < kPC; pl += 8) {
      float mx = -INFINITY;
      for (int l = lane; l < len; l += 32) mx = fmaxf(mx, sA[pl][l]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int l = lane; l < kLmax; l += 32) {
        float e = (l < len) ? expf(sA[pl][l] - mx) : 0.f;
        sA[pl][l] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      for (int l = lane; l < kLmax; l += 32) sA[pl][l] = __fdiv_rn(sA[pl][l], sum);
    }
  }
  // phase 3
  const int tp = tid >> 4;  // 0..15 -> proposals tp, tp+16, tp+32
  const int td = tid & 15;  // float4 column
  float acc[kChunks][3][4];
#pragma unroll
  for (int c = 0; c < kChunks; ++c)
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[c][a][b] = 0.f;

#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    __syncthreads();  // sA ready (first iter) / previous sV consumed
    for (int i = tid; i < kLmax * 16; i += 256) {
      const int l = i >> 4, q4 = i & 15;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (l < len) v = *reinterpret_cast<const float4*>(&val[((int64_t)n * L + l) * D + c * 64 + q4 * 4]);
      *reinterpret_cast<float4*>(&sV[l][q4 * 4]) = v;
    }
    __syncthreads();
    for (int l = 0; l < len; ++l) {
      const float4 v = *reinterpret_cast<const float4*>(&sV[l][td * 4]);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float w = sA[tp + 16 * a][l];
        acc[c][a][0] = fmaf(w, v.x, acc[c][a][0]);
        acc[c][a][1] = fmaf(w, v.y, acc[c][a][1]);
        acc[c][a][2] = fmaf(w, v.z, acc[c][a][2]);
        acc[c][a][3] = fmaf(w, v.w, acc[c][a][3]);
      }
    }
  }
  // phase 4
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c)
#pragma unroll
      for (int b = 0; b < 4; ++b) ss = fmaf(acc[c][a][b], acc[c][a][b], ss);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
    const int p = p0 + tp + 16 * a;
    if (p < P) {
      const int64_t ro = ((int64_t)n * P + p) * D;
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        float4 o = make_float4(__fdiv_rn(acc[c][a][0], denom), __fdiv_rn(acc[c][a][1], denom),
                               __fdiv_rn(acc[c][a][2], denom), __fdiv_rn(acc[c][a][3], denom));
        if (table_f32) *reinterpret_cast<float4*>(&table_f32[ro + c * 64 + td * 4]) = o;
        if (table_f16) {
          __half2 lo = __floats2half2_rn(o.x, o.y), hi = __floats2half2_rn(o.z, o.w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&lo);
          pk.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(&table_f16[ro + c * 64 + td * 4]) = pk;
        }
      }
    }
  }
}


// ------------------------------------------------------------------------------------------
// Key-clip attention table, v2 (D <= 384): same phases, but the a.V product — 95 % of the work — runs on 8 x 8
// register tiles over a whole-D val tile in shared memory.  v1 keeps 3 x 4 tiles and re-stages val per 64-feature
// chunk: 4 LDS per 12 FMA, LSU bound at 70 % wavefront utilisation (profiles/r1_ncu_v13.md).  Here a thread
// owns 8 proposals x 8 features (two float4 columns kD/2 apart, so every LDS.128 of a warp is contiguous):
// 4 LDS.128 per 64 FMA, FMA bound.  Block = (video, 64 proposals), 8 * kD/8 threads, one block per SM
// (val tile 128 x kD x 4 B = 192 KB at kD = 384 + transposed weights 32 KB = 224 KB of the 227 KB).
constexpr int kPC2 = 64;  // proposals per block: 8 proposal groups x kD/8 feature groups = 12 warps at kD = 384 — a
                          // multiple of the 4 schedulers (48 proposals = 9 warps left one scheduler with 3 warps and
                          // the others waiting a third of the a.V loop at the barrier)
constexpr int kAtLd = kPC2 + 4;  // sAt row stride: 16-byte aligned rows, 4-way instead of 32-way conflicts on the transposed stores
template <int kD>
__global__ void __launch_bounds__(kPC2 / 8 * kD / 8)
frame_attn_table_v2_kernel(const float* __restrict__ E, const float* __restrict__ val,
                           const int32_t* __restrict__ lengths, int L, int T,
                           float* __restrict__ table_f32, __half* __restrict__ table_f16) {
  constexpr int kFG = kD / 8;            // feature groups (threads per proposal group)
  constexpr int kThreads = kPC2 / 8 * kFG;
  extern __shared__ __align__(16) float smem_t2[];
  float* sV = smem_t2;                               // kLmax x kD   (phase 3)
  float* sAt = smem_t2 + kLmax * kD;                 // kLmax x kAtLd  transposed softmax weights
  float (*sE)[33] = reinterpret_cast<float (*)[33]>(smem_t2);                     // phases 1-2 input, aliased on sV
  static_assert(kLmax * 33 <= kLmax * kD, "the E tile must fit the val tile");
  const int P = T * (T + 1) / 2;
  const int nchunk = (P + kPC2 - 1) / kPC2;
  const int n = blockIdx.x / nchunk;
  const int p0 = (blockIdx.x % nchunk) * kPC2;
  int len = lengths[n];
  len = len < 1 ? 1 : (len > L ? L : len);
  const int tid = threadIdx.x;

  // async copies: the E tile (4-byte cp.async into the padded rows) and, behind it, every val row that does not
  // alias the E tile — those land while phases 1-2 run
  constexpr int kAliased = (kLmax * 33 + kD - 1) / kD;   // val rows overlapping the E tile
  for (int i = tid; i < kLmax * 32; i += kThreads) {
    const int l = i >> 5, c = i & 31;
    if (l < L && c < T) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&sE[l][c]);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(&E[((int64_t)n * L + l) * T + c]));
    } else {
      sE[l][c] = 0.f;
    }
  }
  asm volatile("cp.async.commit_group;");
  for (int i = tid + kAliased * (kD / 4); i < len * (kD / 4); i += kThreads) {
    const int l = i / (kD / 4), c4 = i % (kD / 4);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&sV[l * kD + c4 * 4]);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(&val[((int64_t)n * L + l) * kD + c4 * 4]));
  }
  asm volatile("cp.async.commit_group;");
  asm volatile("cp.async.wait_group 1;" ::: "memory");   // E tile landed
  __syncthreads();
  // phases 1 + 2, one warp per PAIR of adjacent proposals (same window length except at a length boundary): the
  // windows are warp uniform, every lane owns 4 frames of both rows (8 independent LDS/FADD chains), the logits
  // never leave registers.  v1 spends as many instructions inverting p -> (w, s) per (p, l) entry and walking
  // dependent LDS chains as on the a.V product.  Same arithmetic as v1: sequential window sum, / w, accurate
  // expf, e / sum.
  {
    constexpr int kQ = kLmax / 32;
    const int warp = tid >> 5, lane = tid & 31;
    for (int pr = warp; pr < kPC2 / 2; pr += kThreads / 32) {
      int w[2], s0[2];
      float lg[2][kQ];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int p = min(p0 + 2 * pr + r, P - 1);
        int ww = 1, off = 0;
        while (off + (T - ww + 1) <= p) { off += T - ww + 1; ++ww; }
        w[r] = ww;
        s0[r] = p - off;
#pragma unroll
        for (int k = 0; k < kQ; ++k) lg[r][k] = sE[lane + 32 * k][s0[r]];
      }
      const int wmax = max(w[0], w[1]);
      for (int i = 1; i < wmax; ++i) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          if (i < w[r]) {
#pragma unroll
            for (int k = 0; k < kQ; ++k) lg[r][k] = __fadd_rn(lg[r][k], sE[lane + 32 * k][s0[r] + i]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float fw = (float)w[r];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < kQ; ++k) {
          lg[r][k] = __fdiv_rn(lg[r][k], fw);
          if (lane + 32 * k < len) mx = fmaxf(mx, lg[r][k]);
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < kQ; ++k) {
          lg[r][k] = (lane + 32 * k < len) ? expf(lg[r][k] - mx) : 0.f;
          sum += lg[r][k];
        }
        sum = warp_sum(sum);
        const bool live = p0 + 2 * pr + r < P;
#pragma unroll
        for (int k = 0; k < kQ; ++k)
          sAt[(lane + 32 * k) * kAtLd + 2 * pr + r] = live ? __fdiv_rn(lg[r][k], sum) : 0.f;
      }
    }
  }
  __syncthreads();   // sE dead from here: the first val rows may overwrite it
  for (int i = tid; i < min(len, kAliased) * (kD / 4); i += kThreads) {
    const int l = i / (kD / 4), c4 = i % (kD / 4);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&sV[l * kD + c4 * 4]);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(&val[((int64_t)n * L + l) * kD + c4 * 4]));
  }
  asm volatile("cp.async.commit_group;");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // phase 3: g[8 proposals][8 features] per thread
  const int fg = tid % kFG, pg = tid / kFG;
  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
#pragma unroll 4
  for (int l = 0; l < len; ++l) {
    const float4 v0 = *reinterpret_cast<const float4*>(&sV[l * kD + fg * 4]);
    const float4 v1 = *reinterpret_cast<const float4*>(&sV[l * kD + kD / 2 + fg * 4]);
    const float4 w0 = *reinterpret_cast<const float4*>(&sAt[l * kAtLd + pg * 8]);
    const float4 w1 = *reinterpret_cast<const float4*>(&sAt[l * kAtLd + pg * 8 + 4]);
    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(w[a], v[b], acc[a][b]);
  }
  // phase 4: row norms — partial sums per (proposal, feature group) into smem, summed in a fixed order
  __syncthreads();   // all reads of sAt done: reuse it for the partials (kPC2 x kFG floats <= kLmax x kPC2)
  float* part = sAt;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    float ss = 0.f;
#pragma unroll
    for (int b = 0; b < 8; ++b) ss = fmaf(acc[a][b], acc[a][b], ss);
    part[(pg * 8 + a) * kFG + fg] = ss;
  }
  __syncthreads();
  float* snorm = part + kPC2 * kFG;       // kPC2 floats behind the partials
  static_assert(kPC2 * kFG + kPC2 <= kLmax * kPC2, "partials + norms must fit the sAt region");
  if (tid < kPC2) {
    float ss = 0.f;
    for (int k = 0; k < kFG; ++k) ss += part[tid * kFG + k];
    snorm[tid] = fmaxf(sqrtf(ss), 1e-12f);
  }
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int pl = pg * 8 + a;
    const float denom = snorm[pl];
    const int p = p0 + pl;
    if (p >= P) continue;
    const int64_t ro = ((int64_t)n * P + p) * kD;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col = h * (kD / 2) + fg * 4;
      float4 o = make_float4(__fdiv_rn(acc[a][4 * h + 0], denom), __fdiv_rn(acc[a][4 * h + 1], denom),
                             __fdiv_rn(acc[a][4 * h + 2], denom), __fdiv_rn(acc[a][4 * h + 3], denom));
      if (table_f32) *reinterpret_cast<float4*>(&table_f32[ro + col]) = o;
      if (table_f16) {
        __half2 lo = __floats2half2_rn(o.x, o.y), hi = __floats2half2_rn(o.z, o.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&lo);
        pk.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(&table_f16[ro + col]) = pk;
      }
    }
  }
}

}  // namespace dkd

using namespace dkd;

extern "C" int dkd_normalize_rows(const float* x, int64_t rows, int32_t D, float eps, float* out_f32,
                                  uint16_t* out_bf16, uint16_t* out_f16, int64_t rows_out_pad, void* stream) {
  if (!x || rows < 0 || D <= 0 || (!out_f32 && !out_bf16 && !out_f16)) return DKD_ERR_ARG;
  if (rows_out_pad < rows) rows_out_pad = rows;
  if (rows_out_pad == 0) return DKD_OK;
  const int wpb = 8;
  int64_t blocks = (rows_out_pad + wpb - 1) / wpb;
  if (blocks > 148 * 64) blocks = 148 * 64;
  normalize_rows_kernel<<<(unsigned)blocks, wpb * 32, 0, (cudaStream_t)stream>>>(
      x, rows, D, eps, out_f32, reinterpret_cast<__nv_bfloat16*>(out_bf16), reinterpret_cast<__half*>(out_f16),
      rows_out_pad);
  DKD_LAUNCH_CHECK();
  return DKD_OK;
}

extern "C" int dkd_downsample_clips(const float* frames, const int32_t* lengths, int32_t Nv, int32_t L,
                                    int32_t D, int32_t T, float* clips, void* stream) {
  if (!frames || !lengths || !clips || Nv < 0 || L <= 0 || D <= 0 || T <= 0) return DKD_ERR_ARG;
  if (Nv == 0) return DKD_OK;
  downsample_clips_kernel<<<Nv, 128, 0, (cudaStream_t)stream>>>(frames, lengths, L, D, T, clips);
  DKD_LAUNCH_CHECK();
  return DKD_OK;
}

template <int kPairs>
static int launch_build_proposals(const float* clips, int Nv, int T, int D, uint16_t* pb, float* ps,
                                  float* pf, cudaStream_t st, bool half) {
  size_t smem = (size_t)T * D * sizeof(float);
  auto* bf = reinterpret_cast<__nv_bfloat16*>(pb);
  if (half) {
    if (pf) return DKD_ERR_ARG;
    DKD_CUDA_TRY(cudaFuncSetAttribute(build_proposals_kernel<kPairs, false, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    build_proposals_kernel<kPairs, false, true><<<Nv, 512, smem, st>>>(clips, T, D, bf, ps, pf);
  } else if (pf) {
    DKD_CUDA_TRY(cudaFuncSetAttribute(build_proposals_kernel<kPairs, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    build_proposals_kernel<kPairs, true><<<Nv, 512, smem, st>>>(clips, T, D, bf, ps, pf);
  } else {
    DKD_CUDA_TRY(cudaFuncSetAttribute(build_proposals_kernel<kPairs, false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    build_proposals_kernel<kPairs, false><<<Nv, 512, smem, st>>>(clips, T, D, bf, ps, pf);
  }
  DKD_LAUNCH_CHECK();
  return DKD_OK;
}

static int build_proposals_any(const float* clips, int32_t Nv, int32_t T, int32_t D, uint16_t* prop_rows,
                               float* prop_scale, float* prop_f32, void* stream, bool half) {
  if (!clips || Nv < 0 || T <= 0 || D <= 0) return DKD_ERR_ARG;
  if (T > 32 || D % 64 != 0 || D > 512) return DKD_ERR_SHAPE;
  if (Nv == 0) return DKD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  switch (D / 64) {
    case 1: return launch_build_proposals<1>(clips, Nv, T, D, prop_rows, prop_scale, prop_f32, st, half);
    case 2: return launch_build_proposals<2>(clips, Nv, T, D, prop_rows, prop_scale, prop_f32, st, half);
    case 3: return launch_build_proposals<3>(clips, Nv, T, D, prop_rows, prop_scale, prop_f32, st, half);
    case 4: return launch_build_proposals<4>(clips, Nv, T, D, prop_rows, prop_scale, prop_f32, st, half);
    case 5: return launch_build_proposals<5>(clips, Nv, T, D, prop_rows, prop_scale, prop_f32, st, half);
    case 6: return launch_build_proposals<6>(clips, Nv, T, D, prop_rows, prop_scale, prop_f32, st, half);
    case 7: return launch_build_proposals<7>(clips, Nv, T, D, prop_rows, prop_scale, prop_f32, st, half);
    case 8: return launch_build_proposals<8>(clips, Nv, T, D, prop_rows, prop_scale, prop_f32, st, half);
  }
  return DKD_ERR_SHAPE;
}

extern "C" int dkd_build_proposals(const float* clips, int32_t Nv, int32_t T, int32_t D,
                                   uint16_t* prop_bf16, float* prop_scale, float* prop_f32,
                                   void* stream) {
  return build_proposals_any(clips, Nv, T, D, prop_bf16, prop_scale, prop_f32, stream, false);
}

extern "C" int dkd_build_proposals_f16(const float* clips, int32_t Nv, int32_t T, int32_t D,
                                       uint16_t* prop_f16, float* prop_scale, void* stream) {
  if (!prop_f16) return DKD_ERR_ARG;
  return build_proposals_any(clips, Nv, T, D, prop_f16, prop_scale, nullptr, stream, true);
}

template <int kChunks>
static int launch_frame_table(const float* E, const float* val, const int32_t* lengths, int Nv, int L,
                              int T, int D, float* tf, uint16_t* tb, cudaStream_t st) {
  const int P = T * (T + 1) / 2;
  const unsigned grid = (unsigned)Nv * (unsigned)((P + kPC - 1) / kPC);
  const size_t smem = sizeof(float) * (kLmax * 64 + kLmax * 33 + kPC * (kLmax + 1));
  DKD_CUDA_TRY(cudaFuncSetAttribute(frame_attn_table_kernel<kChunks>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  frame_attn_table_kernel<kChunks><<<grid, 256, smem, st>>>(E, val, lengths, L, T, D, tf,
                                                            reinterpret_cast<__half*>(tb));
  DKD_LAUNCH_CHECK();
  return DKD_OK;
}

template <int kD>
static int launch_frame_table_v2(const float* E, const float* val, const int32_t* lengths, int Nv, int L, int T,
                                 float* tf, uint16_t* tb, cudaStream_t st) {
  const int P = T * (T + 1) / 2;
  const unsigned grid = (unsigned)Nv * (unsigned)((P + kPC2 - 1) / kPC2);
  const size_t smem = sizeof(float) * ((size_t)kLmax * kD + (size_t)kLmax * kAtLd);
  DKD_CUDA_TRY(cudaFuncSetAttribute(frame_attn_table_v2_kernel<kD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
  frame_attn_table_v2_kernel<kD><<<grid, kPC2 / 8 * kD / 8, smem, st>>>(E, val, lengths, L, T, tf,
                                                                reinterpret_cast<__half*>(tb));
  DKD_LAUNCH_CHECK();
  return DKD_OK;
}

extern "C" int dkd_frame_attn_table(const float* E, const float* val, const int32_t* lengths, int32_t Nv,
                                    int32_t L, int32_t T, int32_t D, float* table_f32,
                                    uint16_t* table_f16, void* stream) {
  if (!E || !val || !lengths || Nv < 0 || (!table_f32 && !table_f16)) return DKD_ERR_ARG;
  if (L <= 0 || L > kLmax || T <= 0 || T > 32 || D % 64 != 0 || D > 512) return DKD_ERR_SHAPE;
  if (Nv == 0) return DKD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (D == 384) return launch_frame_table_v2<384>(E, val, lengths, Nv, L, T, table_f32, table_f16, st);
  if (D == 256) return launch_frame_table_v2<256>(E, val, lengths, Nv, L, T, table_f32, table_f16, st);
  switch (D / 64) {
    case 1: return launch_frame_table<1>(E, val, lengths, Nv, L, T, D, table_f32, table_f16, st);
    case 2: return launch_frame_table<2>(E, val, lengths, Nv, L, T, D, table_f32, table_f16, st);
    case 3: return launch_frame_table<3>(E, val, lengths, Nv, L, T, D, table_f32, table_f16, st);
    case 4: return launch_frame_table<4>(E, val, lengths, Nv, L, T, D, table_f32, table_f16, st);
    case 6: return launch_frame_table<6>(E, val, lengths, Nv, L, T, D, table_f32, table_f16, st);
    case 8: return launch_frame_table<8>(E, val, lengths, Nv, L, T, D, table_f32, table_f16, st);
  }
  return DKD_ERR_SHAPE;
}
