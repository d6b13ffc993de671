// Training-step similarity kernels (BASELINE.json configs[4]): the in-batch query x video
// similarity of DLDKD.forward (method/model.py:109-157) with a hand-written backward, and the masked-softmax KL
// loss over the positive video's frame curve (compute_kl_loss, method/model.py:184-197), fused forward+backward.
//
// The reference runs get_sim_scores (:307-329) and get_unnormalized_sim_scores (:331-350) back to back on the
// same operands — two einsums, two (M, L, N) tensors, two normalisations — and then reads only the (M, N) maxima
// plus, for the KL term, the column of the positive video.  Here ONE pass of fp32 dot products produces
//   max_n / arg_n   max_l cos(q_m, x_nl)          (masked frames exactly -1e10, first argmax)
//   max_u / arg_u   max_l q_m . x_nl              (same masking)
//   curve[m, l]     cos(q_m, x_{labels[m], l})    (the only part of the per-frame tensor the losses read)
// with cos = (q . x) * rq[m] * rx[n, l], rq / rx = 1 / max(||.||, 1e-12) (F.normalize's denominator).  The
// (M, L, N) tensor never exists.  Backward is two gather-style kernels (one block per query for grad_q, one block
// per video and 128-feature slab for grad_x, accumulating in shared memory): no atomics, deterministic.
#include "dkd_common.cuh"

namespace dkd {

__global__ void row_inv_norms_kernel(const float* __restrict__ x, int64_t rows, int D, float eps,
                                     float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* xr = x + r * D;
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) ss = fmaf(xr[d], xr[d], ss);
  ss = warp_sum(ss);
  if (lane == 0) out[r] = __fdiv_rn(1.0f, fmaxf(sqrtf(ss), eps));
}

constexpr int kTM = 64;       // query rows per block
constexpr int kKC = 32;       // K chunk
constexpr int kLd = kKC + 4;  // padded smem stride
constexpr int kLmax = 128;    // frames per video (max_ctx_l)

struct TrainSimParams {
  const float *q, *x, *rq, *rx;
  const uint8_t* mask;
  const int32_t* labels;
  int M, N, L, D;
  float *max_n, *max_u, *curve;
  int32_t *arg_n, *arg_u;
};

// grid (N, ceil(M / 64)), 256 threads: thread (tm, ti) owns query rows tm, tm+32 and frames ti + 8 j.
__global__ void __launch_bounds__(256)
train_sim_fwd_kernel(const TrainSimParams p) {
  constexpr int kJ = kLmax / 8;
  __shared__ __align__(16) float sQ[kTM * kLd];
  __shared__ __align__(16) float sX[kLmax * kLd];
  const int n = blockIdx.x;
  const int r0 = blockIdx.y * kTM;
  const int tid = threadIdx.x;
  const int tm = tid >> 3, ti = tid & 7;
  const float* xbase = p.x + (int64_t)n * p.L * p.D;

  float acc[2][kJ];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int j = 0; j < kJ; ++j) acc[a][j] = 0.f;

  for (int kc = 0; kc < p.D; kc += kKC) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int idx = tid + t * 256;
      const int r = idx >> 3, c4 = idx & 7;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < p.M) v = *reinterpret_cast<const float4*>(&p.q[(int64_t)(r0 + r) * p.D + kc + c4 * 4]);
      *reinterpret_cast<float4*>(&sQ[r * kLd + c4 * 4]) = v;
    }
#pragma unroll
    for (int t = 0; t < kLmax / 32; ++t) {
      const int idx = tid + t * 256;
      const int r = idx >> 3, c4 = idx & 7;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < p.L) v = *reinterpret_cast<const float4*>(&xbase[(int64_t)r * p.D + kc + c4 * 4]);
      *reinterpret_cast<float4*>(&sX[r * kLd + c4 * 4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kKC; k += 4) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sQ[tm * kLd + k]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sQ[(tm + 32) * kLd + k]);
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        const float4 b = *reinterpret_cast<const float4*>(&sX[(ti + 8 * j) * kLd + k]);
        acc[0][j] = fmaf(a0.x, b.x, acc[0][j]);
        acc[0][j] = fmaf(a0.y, b.y, acc[0][j]);
        acc[0][j] = fmaf(a0.z, b.z, acc[0][j]);
        acc[0][j] = fmaf(a0.w, b.w, acc[0][j]);
        acc[1][j] = fmaf(a1.x, b.x, acc[1][j]);
        acc[1][j] = fmaf(a1.y, b.y, acc[1][j]);
        acc[1][j] = fmaf(a1.z, b.z, acc[1][j]);
        acc[1][j] = fmaf(a1.w, b.w, acc[1][j]);
      }
    }
    __syncthreads();
  }

  // per-frame scale and mask of this thread's 16 frames
  float rxv[kJ];
  bool ok[kJ];
#pragma unroll
  for (int j = 0; j < kJ; ++j) {
    const int row = ti + 8 * j;
    const bool in = row < p.L;
    rxv[j] = in ? p.rx[(int64_t)n * p.L + row] : 0.f;
    ok[j] = in && (!p.mask || p.mask[(int64_t)n * p.L + row] != 0);
  }
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int m = r0 + tm + 32 * a;
    const bool live = m < p.M;
    const float rqm = live ? p.rq[m] : 0.f;
    const bool pos = live && p.curve && p.labels && p.labels[m] == n;
    float bn = -INFINITY, bu = -INFINITY;
    int in_ = 0x7fffffff, iu = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      const int row = ti + 8 * j;
      if (row >= p.L) continue;
      const float dot = acc[a][j];
      const float cn = ok[j] ? __fmul_rn(__fmul_rn(dot, rqm), rxv[j]) : DKD_MASKED_SCORE;
      const float cu = ok[j] ? dot : DKD_MASKED_SCORE;
      if (better(cn, row, bn, in_)) { bn = cn; in_ = row; }
      if (better(cu, row, bu, iu)) { bu = cu; iu = row; }
      if (pos) p.curve[(int64_t)m * p.L + row] = cn;
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const float vn = __shfl_xor_sync(0xffffffffu, bn, o);
      const int jn = __shfl_xor_sync(0xffffffffu, in_, o);
      const float vu = __shfl_xor_sync(0xffffffffu, bu, o);
      const int ju = __shfl_xor_sync(0xffffffffu, iu, o);
      if (better(vn, jn, bn, in_)) { bn = vn; in_ = jn; }
      if (better(vu, ju, bu, iu)) { bu = vu; iu = ju; }
    }
    if (live && ti == 0) {
      const int64_t o = (int64_t)m * p.N + n;
      if (p.max_n) p.max_n[o] = bn;
      if (p.arg_n) p.arg_n[o] = in_;
      if (p.max_u) p.max_u[o] = bu;
      if (p.arg_u) p.arg_u[o] = iu;
    }
  }
}

struct TrainBwdParams {
  const float *q, *x, *rq, *rx;
  const uint8_t* mask;
  const int32_t* labels;
  int M, N, L, D;
  const float *max_n, *curve;
  const int32_t *arg_n, *arg_u;
  const float *g_n, *g_u, *g_c;  // upstream gradients of max_n, max_u (M, N) and curve (M, L); any may be null
  float *grad_q, *grad_x;
};

// d cos / d q = rq rx x - cos rq^2 q ;  d (q.x) / d q = x.  One block per query, thread d owns features d, d+128, ...
__global__ void __launch_bounds__(128)
train_sim_bwd_q_kernel(const TrainBwdParams p) {
  constexpr int kPer = 4;  // D <= 512
  const int m = blockIdx.x;
  const int tid = threadIdx.x;
  const float rqm = p.rq[m];
  float acc[kPer] = {0.f, 0.f, 0.f, 0.f};
  float sumcos = 0.f;
  for (int n = 0; n < p.N; ++n) {
    const int64_t o = (int64_t)m * p.N + n;
    if (p.g_n) {
      const float g = p.g_n[o];
      const int l = p.arg_n[o];
      if (g != 0.f && (!p.mask || p.mask[(int64_t)n * p.L + l] != 0)) {
        const float c = g * rqm * p.rx[(int64_t)n * p.L + l];
        const float* xr = p.x + ((int64_t)n * p.L + l) * p.D;
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
          const int d = tid + 128 * k;
          if (d < p.D) acc[k] = fmaf(c, xr[d], acc[k]);
        }
        sumcos = fmaf(g, p.max_n[o], sumcos);
      }
    }
    if (p.g_u) {
      const float g = p.g_u[o];
      const int l = p.arg_u[o];
      if (g != 0.f && (!p.mask || p.mask[(int64_t)n * p.L + l] != 0)) {
        const float* xr = p.x + ((int64_t)n * p.L + l) * p.D;
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
          const int d = tid + 128 * k;
          if (d < p.D) acc[k] = fmaf(g, xr[d], acc[k]);
        }
      }
    }
  }
  if (p.g_c && p.labels) {
    const int n = p.labels[m];
    for (int l = 0; l < p.L; ++l) {
      if (p.mask && p.mask[(int64_t)n * p.L + l] == 0) continue;
      const float g = p.g_c[(int64_t)m * p.L + l];
      if (g == 0.f) continue;
      const float c = g * rqm * p.rx[(int64_t)n * p.L + l];
      const float* xr = p.x + ((int64_t)n * p.L + l) * p.D;
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        const int d = tid + 128 * k;
        if (d < p.D) acc[k] = fmaf(c, xr[d], acc[k]);
      }
      sumcos = fmaf(g, p.curve[(int64_t)m * p.L + l], sumcos);
    }
  }
  const float s = sumcos * rqm * rqm;
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    const int d = tid + 128 * k;
    if (d < p.D) p.grad_q[(int64_t)m * p.D + d] = acc[k] - s * p.q[(int64_t)m * p.D + d];
  }
}

// d cos / d x = rq rx q - cos rx^2 x ;  d (q.x) / d x = q.  Block = (video n, slab of 128 features): the (L, 128)
// gradient tile accumulates in shared memory, column d owned by thread d (no conflicts, no atomics); the scalar
// sum of g * cos per frame by thread 0.  The per-query scalars of this video are staged through shared memory in
// pieces of kStage queries.
constexpr int kStage = 512;

__global__ void __launch_bounds__(128)
train_sim_bwd_x_kernel(const TrainBwdParams p) {
  extern __shared__ __align__(16) float smem_bx[];
  float* sx = smem_bx;                  // L x 128
  float* sc = sx + p.L * 128;           // L
  float* s_gn = sc + p.L;               // kStage each
  float* s_gu = s_gn + kStage;
  float* s_cs = s_gu + kStage;          // g_n * max_n
  int* s_ln = reinterpret_cast<int*>(s_cs + kStage);
  int* s_lu = s_ln + kStage;
  const int n = blockIdx.x;
  const int d0 = blockIdx.y * 128;
  const int tid = threadIdx.x;
  const int d = d0 + tid;
  const bool din = d < p.D;
  for (int i = tid; i < p.L * 128; i += 128) sx[i] = 0.f;
  for (int i = tid; i < p.L; i += 128) sc[i] = 0.f;
  for (int m0 = 0; m0 < p.M; m0 += kStage) {
    __syncthreads();
    for (int i = tid; i < kStage; i += 128) {
      const int m = m0 + i;
      float gn = 0.f, gu = 0.f, cs = 0.f;
      int ln = 0, lu = 0;
      if (m < p.M) {
        const int64_t o = (int64_t)m * p.N + n;
        if (p.g_n) {
          ln = p.arg_n[o];
          gn = p.g_n[o];
          if (p.mask && p.mask[(int64_t)n * p.L + ln] == 0) gn = 0.f;
          cs = gn * p.max_n[o];
          gn *= p.rq[m] * p.rx[(int64_t)n * p.L + ln];
        }
        if (p.g_u) {
          lu = p.arg_u[o];
          gu = p.g_u[o];
          if (p.mask && p.mask[(int64_t)n * p.L + lu] == 0) gu = 0.f;
        }
      }
      s_gn[i] = gn; s_gu[i] = gu; s_cs[i] = cs; s_ln[i] = ln; s_lu[i] = lu;
    }
    __syncthreads();
    const int cnt = min(kStage, p.M - m0);
    for (int i = 0; i < cnt; ++i) {
      const int m = m0 + i;
      const float gn = s_gn[i], gu = s_gu[i];
      const bool pos = p.g_c && p.labels && p.labels[m] == n;
      if (gn == 0.f && gu == 0.f && !pos) continue;
      const float qv = din ? p.q[(int64_t)m * p.D + d] : 0.f;
      if (gn != 0.f) {
        sx[s_ln[i] * 128 + tid] = fmaf(gn, qv, sx[s_ln[i] * 128 + tid]);
        if (tid == 0) sc[s_ln[i]] += s_cs[i];
      }
      if (gu != 0.f) sx[s_lu[i] * 128 + tid] = fmaf(gu, qv, sx[s_lu[i] * 128 + tid]);
      if (pos) {
        const float rqm = p.rq[m];
        for (int l = 0; l < p.L; ++l) {
          if (p.mask && p.mask[(int64_t)n * p.L + l] == 0) continue;
          const float g = p.g_c[(int64_t)m * p.L + l];
          if (g == 0.f) continue;
          sx[l * 128 + tid] = fmaf(g * rqm * p.rx[(int64_t)n * p.L + l], qv, sx[l * 128 + tid]);
          if (tid == 0) sc[l] = fmaf(g, p.curve[(int64_t)m * p.L + l], sc[l]);
        }
      }
    }
  }
  __syncthreads();
  if (din) {
    for (int l = 0; l < p.L; ++l) {
      const int64_t o = ((int64_t)n * p.L + l) * p.D + d;
      const float r = p.rx[(int64_t)n * p.L + l];
      p.grad_x[o] = sx[l * 128 + tid] - sc[l] * r * r * p.x[o];
    }
  }
}

// KL(softmax(target / temp) || softmax(pred / temp)) over the first lens[m] frames, reduction 'sum'
// (F.kl_div(log_softmax(p / temp), softmax(t / temp), reduction='sum'), method/model.py:192-195), and its
// gradient with respect to pred: (softmax(pred / temp) - softmax(target / temp)) / temp.  One warp per query.
__global__ void __launch_bounds__(128)
kl_curve_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                const int32_t* __restrict__ lens, int M, int L, float temp, float* __restrict__ loss,
                float* __restrict__ dpred) {
  constexpr int kPer = kLmax / 32;
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  int len = lens[m];
  len = len < 0 ? 0 : (len > L ? L : len);
  float pv[kPer], tv[kPer];
  float mp = -INFINITY, mt = -INFINITY;
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    const int l = lane + 32 * k;
    const bool in = l < len;
    pv[k] = in ? __fdiv_rn(pred[(int64_t)m * L + l], temp) : -INFINITY;
    tv[k] = in ? __fdiv_rn(target[(int64_t)m * L + l], temp) : -INFINITY;
    mp = fmaxf(mp, pv[k]);
    mt = fmaxf(mt, tv[k]);
  }
  mp = warp_max(mp);
  mt = warp_max(mt);
  float sp = 0.f, st = 0.f;
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    const int l = lane + 32 * k;
    if (l < len) {
      sp += expf(pv[k] - mp);
      st += expf(tv[k] - mt);
    }
  }
  sp = warp_sum(sp);
  st = warp_sum(st);
  const float lsp = logf(sp), lst = logf(st);
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    const int l = lane + 32 * k;
    float g = 0.f;
    if (l < len) {
      const float logp = pv[k] - mp - lsp;
      const float logt = tv[k] - mt - lst;
      const float t = __fdiv_rn(expf(tv[k] - mt), st);
      if (t > 0.f) acc = fmaf(t, logt - logp, acc);
      g = __fdiv_rn(expf(logp) - t, temp);
    }
    if (dpred && l < L) dpred[(int64_t)m * L + l] = g;
  }
  acc = warp_sum(acc);
  if (lane == 0 && loss) loss[m] = len > 0 ? acc : 0.f;
}

}  // namespace dkd

using namespace dkd;

extern "C" int dkd_row_inv_norms(const float* x, int64_t rows, int32_t D, float eps, float* out, void* stream) {
  if (!x || !out || rows < 0 || D <= 0) return DKD_ERR_ARG;
  if (rows == 0) return DKD_OK;
  const int wpb = 8;
  row_inv_norms_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(x, rows, D, eps, out);
  DKD_LAUNCH_CHECK();
  return DKD_OK;
}

extern "C" int dkd_train_sim_fwd(const float* q, const float* x, const float* rq, const float* rx,
                                 const uint8_t* mask, const int32_t* labels, int32_t M, int32_t N, int32_t L,
                                 int32_t D, float* max_n, int32_t* arg_n, float* max_u, int32_t* arg_u,
                                 float* curve, void* stream) {
  if (!q || !x || !rq || !rx || M < 0 || N < 0) return DKD_ERR_ARG;
  if (curve && !labels) return DKD_ERR_ARG;
  if (L <= 0 || L > kLmax || D <= 0 || D % 32 != 0) return DKD_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(x)) & 15) return DKD_ERR_ALIGN;
  if (M == 0 || N == 0) return DKD_OK;
  TrainSimParams p{q, x, rq, rx, mask, labels, M, N, L, D, max_n, max_u, curve, arg_n, arg_u};
  dim3 grid(N, (M + kTM - 1) / kTM);
  train_sim_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  DKD_LAUNCH_CHECK();
  return DKD_OK;
}

extern "C" int dkd_train_sim_bwd(const float* q, const float* x, const float* rq, const float* rx,
                                 const uint8_t* mask, const int32_t* labels, int32_t M, int32_t N, int32_t L,
                                 int32_t D, const float* max_n, const int32_t* arg_n, const int32_t* arg_u,
                                 const float* curve, const float* g_max_n, const float* g_max_u,
                                 const float* g_curve, float* grad_q, float* grad_x, void* stream) {
  if (!q || !x || !rq || !rx || M < 0 || N < 0 || (!grad_q && !grad_x)) return DKD_ERR_ARG;
  if ((g_max_n && (!arg_n || !max_n)) || (g_max_u && !arg_u) || (g_curve && (!curve || !labels))) return DKD_ERR_ARG;
  if (L <= 0 || L > kLmax || D <= 0 || D > 512) return DKD_ERR_SHAPE;
  if (M == 0 || N == 0) return DKD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  TrainBwdParams p{q, x, rq, rx, mask, labels, M, N, L, D, max_n, curve, arg_n, arg_u, g_max_n, g_max_u, g_curve,
                   grad_q, grad_x};
  if (grad_q) {
    train_sim_bwd_q_kernel<<<M, 128, 0, st>>>(p);
    DKD_LAUNCH_CHECK();
  }
  if (grad_x) {
    const size_t smem = sizeof(float) * ((size_t)L * 128 + L + 5 * kStage);
    DKD_CUDA_TRY(cudaFuncSetAttribute(train_sim_bwd_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(N, (D + 127) / 128);
    train_sim_bwd_x_kernel<<<grid, 128, smem, st>>>(p);
    DKD_LAUNCH_CHECK();
  }
  return DKD_OK;
}

extern "C" int dkd_kl_curve_loss(const float* pred, const float* target, const int32_t* lens, int32_t M,
                                 int32_t L, float temp, float* loss, float* dpred, void* stream) {
  if (!pred || !target || !lens || M < 0 || (!loss && !dpred) || !(temp > 0.f)) return DKD_ERR_ARG;
  if (L <= 0 || L > kLmax) return DKD_ERR_SHAPE;
  if (M == 0) return DKD_OK;
  kl_curve_kernel<<<(M + 3) / 4, 128, 0, (cudaStream_t)stream>>>(pred, target, lens, M, L, temp, loss, dpred);
  DKD_LAUNCH_CHECK();
  return DKD_OK;
}
