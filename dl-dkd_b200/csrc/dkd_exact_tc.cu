// Exact (fp32-grade) clip-scale scores on the tensor cores: per-clip dot products by error-compensated
// TF32 (3xTF32: x = hi + lo, hi.hi + hi.lo + lo.hi accumulated in fp32 — relative error per product
// ~2^-21, i.e. inside the fp32 summation-order noise of the reference's own einsum), then the
// T(T+1)/2 window cosines from running sums of the T per-clip dots (SURVEY §7 "linearity"):
//
//   d[i]        = qn[m] . clips[n, i]                       i = 0..T-1      (tensor cores)
//   S[p(w, s)]  = (d[s] + d[s+1] + ... + d[s+w-1]) * prop_scale[n, p]       (fp32, running sum)
//   out_max     = max_p S,   out_arg = first argmax_p       (torch.max tie rule)
//
// This is the fp32 flavour of get_clip_scale_scores used (a) as the exact drop-in path,
// (b) to re-resolve the pairs whose bf16 argmax is ambiguous, (c) to rescore the top-K candidates.
// (b) and (c) address the pairs through a CSR by video (vid_ptr / q_list).
//
// Block = one video x kWarps tiles of 16 list entries (grid.y blocks stride over the video's list).
// The video's T x D clip tile is split once into tf32 hi / lo planes in shared memory; every warp
// streams its 16 query rows from global memory (16 B per lane per 16 features, software-prefetched)
// and issues mma.sync.m16n8k8.tf32.  The k index inside a 16-feature chunk is permuted identically
// for both operands (lane t owns features 4t..4t+3) so that each operand fragment is one 128-bit load.
#include "dkd_common.cuh"

namespace dkd {

constexpr int kTcWarps = 16;          // warps per block
constexpr int kTcRows = 16;           // list entries per warp tile (MMA M)
constexpr int kTcDLd = 34;            // per-warp dots tile stride (words): 2*row + half -> 32 distinct banks

struct ClipTcParams {
  const float* q; int M;
  const float* clips; const float* scale;
  int Nv, T, D;
  float* out_max; int32_t* out_arg; int64_t ld_out;
  const int32_t* vid_ptr; const int32_t* q_list;
};

__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = tf32_rna(x);
  lo = tf32_rna(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float& c0, float& c1, float& c2, float& c3, uint32_t a0, uint32_t a1,
                                         uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kTcWarps * 32, 1)
clip_exact_tc_kernel(const ClipTcParams p) {
  extern __shared__ __align__(16) float smem_tc[];
  const int D = p.D, T = p.T, P = T * (T + 1) / 2;
  const int ldc = ((D + 31) & ~31) + 16;               // == 16 (mod 32) words: conflict-free LDS.128 fragments
  float* sHi = smem_tc;                                 // 32 x ldc   tf32 hi plane of the clips
  float* sLo = sHi + 32 * ldc;                          // 32 x ldc   tf32 lo plane
  float* sScale = sLo + 32 * ldc;                       // 528
  float* sDots = sScale + 528;                          // kTcWarps x 16 x kTcDLd

  const int n = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int e0 = 0, count = p.M;
  if (p.vid_ptr) { e0 = p.vid_ptr[n]; count = p.vid_ptr[n + 1] - e0; }
  const int rows_per_block = kTcWarps * kTcRows;
  if ((int)blockIdx.y * rows_per_block >= count) return;

  // ---- stage the clip tile as tf32 hi / lo planes (rows >= T are zero) ----
  const float* cbase = p.clips + (int64_t)n * T * D;
  const int d4 = D >> 2;
  for (int i = tid; i < 32 * d4; i += kTcWarps * 32) {
    const int r = i / d4, c4 = i - r * d4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < T) v = *reinterpret_cast<const float4*>(&cbase[(int64_t)r * D + c4 * 4]);
    uint32_t h[4], l[4];
    split_tf32(v.x, h[0], l[0]); split_tf32(v.y, h[1], l[1]);
    split_tf32(v.z, h[2], l[2]); split_tf32(v.w, h[3], l[3]);
    *reinterpret_cast<uint4*>(&sHi[r * ldc + c4 * 4]) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(&sLo[r * ldc + c4 * 4]) = make_uint4(l[0], l[1], l[2], l[3]);
  }
  for (int i = tid; i < P; i += kTcWarps * 32) sScale[i] = p.scale[(int64_t)n * P + i];
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  float* myDots = sDots + warp * (kTcRows * kTcDLd);
  const int nchunks = D >> 4;

  for (int tile = blockIdx.y * kTcWarps + warp; tile * kTcRows < count; tile += gridDim.y * kTcWarps) {
    const int r_lo = tile * kTcRows + g, r_hi = r_lo + 8;
    const int64_t q_lo = (r_lo < count) ? (p.q_list ? (int64_t)p.q_list[e0 + r_lo] : (int64_t)r_lo) : 0;
    const int64_t q_hi = (r_hi < count) ? (p.q_list ? (int64_t)p.q_list[e0 + r_hi] : (int64_t)r_hi) : 0;
    const float4* pa_lo = reinterpret_cast<const float4*>(p.q + q_lo * D) + t;
    const float4* pa_hi = reinterpret_cast<const float4*>(p.q + q_hi * D) + t;

    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;

    float4 va = __ldg(pa_lo), vb = __ldg(pa_hi);
    for (int c = 0; c < nchunks; ++c) {
      float4 na = va, nb = vb;
      if (c + 1 < nchunks) { na = __ldg(pa_lo + 4 * (c + 1)); nb = __ldg(pa_hi + 4 * (c + 1)); }
      // A fragments of the two k-steps of this chunk: step 0 uses features (4t, 4t+1), step 1 (4t+2, 4t+3)
      uint32_t ah[2][4], al[2][4];
      split_tf32(va.x, ah[0][0], al[0][0]); split_tf32(vb.x, ah[0][1], al[0][1]);
      split_tf32(va.y, ah[0][2], al[0][2]); split_tf32(vb.y, ah[0][3], al[0][3]);
      split_tf32(va.z, ah[1][0], al[1][0]); split_tf32(vb.z, ah[1][1], al[1][1]);
      split_tf32(va.w, ah[1][2], al[1][2]); split_tf32(vb.w, ah[1][3], al[1][3]);
      const int kof = c * 16 + t * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 bh = *reinterpret_cast<const uint4*>(&sHi[(8 * j + g) * ldc + kof]);
        const uint4 bl = *reinterpret_cast<const uint4*>(&sLo[(8 * j + g) * ldc + kof]);
        // small terms first
        mma_tf32(acc[j][0], acc[j][1], acc[j][2], acc[j][3], al[0][0], al[0][1], al[0][2], al[0][3], bh.x, bh.y);
        mma_tf32(acc[j][0], acc[j][1], acc[j][2], acc[j][3], ah[0][0], ah[0][1], ah[0][2], ah[0][3], bl.x, bl.y);
        mma_tf32(acc[j][0], acc[j][1], acc[j][2], acc[j][3], al[1][0], al[1][1], al[1][2], al[1][3], bh.z, bh.w);
        mma_tf32(acc[j][0], acc[j][1], acc[j][2], acc[j][3], ah[1][0], ah[1][1], ah[1][2], ah[1][3], bl.z, bl.w);
        mma_tf32(acc[j][0], acc[j][1], acc[j][2], acc[j][3], ah[0][0], ah[0][1], ah[0][2], ah[0][3], bh.x, bh.y);
        mma_tf32(acc[j][0], acc[j][1], acc[j][2], acc[j][3], ah[1][0], ah[1][1], ah[1][2], ah[1][3], bh.z, bh.w);
      }
      va = na; vb = nb;
    }

    // ---- dots tile -> shared, then 2 lanes per list entry scan the T(T+1)/2 windows ----
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      *reinterpret_cast<float2*>(&myDots[g * kTcDLd + 8 * j + 2 * t]) = make_float2(acc[j][0], acc[j][1]);
      *reinterpret_cast<float2*>(&myDots[(g + 8) * kTcDLd + 8 * j + 2 * t]) = make_float2(acc[j][2], acc[j][3]);
    }
    __syncwarp();
    const int row = lane >> 1, half = lane & 1;
    const float* drow = myDots + row * kTcDLd;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    // explicit unroll factors: nvcc 12.9 cicc crashes on the default unrolling of this loop nest for sm_100a
#pragma unroll 1
    for (int s = half; s < T; s += 2) {
      float run = 0.f;
      int pi = s;                      // p(1, s) = s
#pragma unroll 4
      for (int w = 1; w <= T - s; ++w) {
        const float d = drow[s + w - 1];
        run = (w == 1) ? d : __fadd_rn(run, d);
        const float v = __fmul_rn(run, sScale[pi]);
        if (better(v, pi, bv, bi)) { bv = v; bi = pi; }
        pi += T - w + 1;               // p(w+1, s) - p(w, s) = T - (w - 1)
      }
    }
    {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, 1);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, 1);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    const int r = tile * kTcRows + row;
    if (r < count && half == 0) {
      const int64_t o = p.vid_ptr ? (int64_t)(e0 + r)
                                  : (int64_t)r * p.ld_out + n;
      p.out_max[o] = bv;
      if (p.out_arg) p.out_arg[o] = bi;
    }
  }
}

}  // namespace dkd

using namespace dkd;

extern "C" int dkd_clip_score_f32(const float* qn, int32_t M, const float* clips, const float* prop_scale,
                                  int32_t Nv, int32_t T, int32_t D, float* out_max, int32_t* out_arg,
                                  int64_t ld_out, const int32_t* vid_ptr, const int32_t* q_list,
                                  void* stream) {
  if (!qn || !clips || !prop_scale || !out_max || M < 0 || Nv < 0) return DKD_ERR_ARG;
  if ((vid_ptr == nullptr) != (q_list == nullptr)) return DKD_ERR_ARG;
  if (T <= 0 || T > 32 || D <= 0 || D % 16 != 0 || D > 512) return DKD_ERR_SHAPE;
  if (!vid_ptr && ld_out < Nv) return DKD_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(qn) & 15) || (reinterpret_cast<uintptr_t>(clips) & 15)) return DKD_ERR_ALIGN;
  if (M == 0 || Nv == 0) return DKD_OK;
  ClipTcParams p{};
  p.q = qn; p.M = M; p.clips = clips; p.scale = prop_scale; p.Nv = Nv; p.T = T; p.D = D;
  p.out_max = out_max; p.out_arg = out_arg; p.ld_out = ld_out; p.vid_ptr = vid_ptr; p.q_list = q_list;
  const int ldc = ((D + 31) & ~31) + 16;
  const size_t smem = sizeof(float) * ((size_t)2 * 32 * ldc + 528 + (size_t)kTcWarps * kTcRows * kTcDLd);
  DKD_CUDA_TRY(cudaFuncSetAttribute(clip_exact_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int rows_per_block = kTcWarps * kTcRows;
  int tiles = (M + rows_per_block - 1) / rows_per_block;
  // dense lists are long: a few blocks per video keep the clip tile resident while striding over the list
  dim3 grid(Nv, tiles < 4 ? tiles : 4);
  clip_exact_tc_kernel<<<grid, kTcWarps * 32, smem, (cudaStream_t)stream>>>(p);
  DKD_LAUNCH_CHECK();
  return DKD_OK;
}
