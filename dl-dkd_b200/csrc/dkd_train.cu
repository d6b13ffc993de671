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

// curve[m, l] = qn[m] . xn[labels[m], l] (both L2-normalised), masked frames -1e10: one warp per (query, 4 frames)
__global__ void __launch_bounds__(256)
train_curve_kernel(const float* __restrict__ qn, const float* __restrict__ xn, const uint8_t* __restrict__ mask,
                   const int32_t* __restrict__ labels, int M, int L, int D, float* __restrict__ curve) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x;
  const int n = labels[m];
  const float4* q4 = reinterpret_cast<const float4*>(qn + (int64_t)m * D);
  for (int l = blockIdx.y * 8 + warp; l < L; l += gridDim.y * 8) {
    const float4* x4 = reinterpret_cast<const float4*>(xn + ((int64_t)n * L + l) * D);
    float acc = 0.f;
    for (int i = lane; i < (D >> 2); i += 32) {
      const float4 a = __ldg(q4 + i), b = __ldg(x4 + i);
      acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) curve[(int64_t)m * L + l] = (!mask || mask[(int64_t)n * L + l]) ? acc : DKD_MASKED_SCORE;
  }
}

extern "C" int dkd_train_curve(const float* qn, const float* xn, const uint8_t* mask, const int32_t* labels, int32_t M,
                               int32_t L, int32_t D, float* curve, void* stream) {
  if (!qn || !xn || !labels || !curve || M < 0) return DKD_ERR_ARG;
  if (L <= 0 || D <= 0 || D % 4 != 0) return DKD_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(qn) | reinterpret_cast<uintptr_t>(xn)) & 15) return DKD_ERR_ALIGN;
  if (M == 0) return DKD_OK;
  dim3 grid(M, (L + 31) / 32 < 4 ? (L + 31) / 32 : 4);
  train_curve_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(qn, xn, mask, labels, M, L, D, curve);
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

// ------------------------------------------------------------------------------------------------------------
// Fused triplet + NCE losses on the (M, N) in-batch score matrices, forward value AND gradient in one pass
// (get_clip_triplet_loss method/model.py:352-388; clip_nce / clip_nce_soft method/model_components.py:106-233).
// The reference builds these from ~150 small tensor ops and two Python loops over the videos; here a row kernel
// (one warp per query), a column kernel (one block per video) and a 1-block reduction produce both loss terms and
// d loss / d scores.  No float atomics: per-row / per-column partial losses are summed in a fixed order, every
// gradient entry is written by the row kernel and updated by exactly one thread of the column kernel.
namespace dkd {

struct LossParams {
  const float *s_n, *s_u, *sims;   // sims == nullptr: hard labels (clip_nce); sims == s_u: self distillation
  const int32_t *labels, *t2v_draw, *v2t_pick;
  int M, N;
  float margin, alpha, belta;
  int soft;                        // label_style == 'soft'
  float *g_n, *g_u;                // (M, N) gradients of (triplet, nce) w.r.t. (s_n, s_u)
  float *row_part, *col_part;      // (2, M) and (2, N) partial losses
};

__device__ __forceinline__ float lse_warp_row(const float* row, int N, int lane, float* mx_out) {
  float mx = -INFINITY;
  for (int j = lane; j < N; j += 32) mx = fmaxf(mx, row[j]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int j = lane; j < N; j += 32) s += expf(row[j] - mx);
  s = warp_sum(s);
  *mx_out = mx;
  return mx + logf(s);
}

// part weights of clip_nce_soft (reduction 'mean'): the hard part counts only if hardQ != 0 and hardV != 0, the soft
// part only if softQ != 0 and softV != 0 (method/model_components.py:190-203)
__device__ __forceinline__ void nce_parts(const LossParams& p, int& hardQ, int& hardV, float& wh, float& ws) {
  hardQ = (int)floorf(p.alpha * (float)p.M);
  hardV = (int)floorf(p.alpha * (float)p.N);
  const bool hard_on = hardQ != 0 && hardV != 0;
  const bool soft_on = (p.M - hardQ) != 0 && (p.N - hardV) != 0;
  wh = hard_on ? p.alpha : 0.f;
  ws = soft_on ? (1.f - p.alpha) : 0.f;
}

__global__ void __launch_bounds__(128)
loss_rows_kernel(const LossParams p) {
  extern __shared__ float smem_lr[];          // 4 warps x 2 rows of N
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 4 + warp;
  if (m >= p.M) return;
  float* rn = smem_lr + (size_t)warp * 2 * p.N;
  float* ru = rn + p.N;
  const int lab = p.labels[m];
  for (int j = lane; j < p.N; j += 32) {
    rn[j] = p.s_n[(int64_t)m * p.N + j];
    ru[j] = p.s_u[(int64_t)m * p.N + j];
  }
  __syncwarp();
  // ---- triplet, text -> video: positive vs the negative at position t2v_draw[m] of the descending order in which
  // the positive (masked to 999) comes first
  {
    const int r = p.t2v_draw[m] - 1;
    int pick = -1;
    for (int j = lane; j < p.N; j += 32) {
      if (j == lab) continue;
      const float v = rn[j];
      int c = 0;
      for (int k = 0; k < p.N; ++k) {
        if (k == lab) continue;
        const float w = rn[k];
        c += (w > v) || (w == v && k < j);
      }
      if (c == r) pick = j;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pick = max(pick, __shfl_xor_sync(0xffffffffu, pick, o));
    float loss = 0.f;
    bool active = false;
    if (pick >= 0) {
      const float d = p.margin + rn[pick] - rn[lab];
      active = d > 0.f;
      loss = active ? d / (float)p.M : 0.f;
    }
    const float g = active ? 1.f / (float)p.M : 0.f;
    for (int j = lane; j < p.N; j += 32)
      p.g_n[(int64_t)m * p.N + j] = (j == pick ? g : 0.f) - (j == lab ? g : 0.f);
    if (lane == 0) p.row_part[m] = loss;
  }
  // ---- NCE, text -> video
  {
    float mx;
    const float lse = lse_warp_row(ru, p.N, lane, &mx);
    if (!p.soft) {  // clip_nce: mean over queries of lse - s[m, lab]
      const float a = 1.f / (float)p.M;
      for (int j = lane; j < p.N; j += 32)
        p.g_u[(int64_t)m * p.N + j] = a * (expf(ru[j] - lse) - (j == lab ? 1.f : 0.f));
      if (lane == 0) p.row_part[p.M + m] = a * (lse - ru[lab]);
    } else {
      int hardQ, hardV;
      float wh, ws;
      nce_parts(p, hardQ, hardV, wh, ws);
      const bool soft_row = m >= hardQ;
      const float a = soft_row ? ws / (float)(p.M - hardQ) : wh / (float)hardQ;
      const float* z = p.sims + (int64_t)m * p.N;
      float zl = 0.f, zmx;
      if (soft_row) {
        zmx = -INFINITY;
        for (int j = lane; j < p.N; j += 32) zmx = fmaxf(zmx, z[j]);
        zmx = warp_max(zmx);
        float s = 0.f;
        for (int j = lane; j < p.N; j += 32) s += expf(z[j] - zmx);
        zl = zmx + logf(warp_sum(s));
      }
      // dot = sum_n IQ[n] s[n];  ps = sum_n P[n] s[n] (self-distillation term)
      float dot = 0.f, ps = 0.f;
      for (int j = lane; j < p.N; j += 32) {
        const float one = (j == lab) ? 1.f : 0.f;
        const float P = soft_row ? expf(z[j] - zl) : 0.f;
        const float iq = soft_row ? fmaxf((1.f - p.belta) * P + p.belta * one, 0.f) : one;
        dot = fmaf(iq, ru[j], dot);
        ps = fmaf(P, ru[j], ps);
      }
      dot = warp_sum(dot);
      ps = warp_sum(ps);
      const bool self = soft_row && (p.sims == p.s_u);
      for (int j = lane; j < p.N; j += 32) {
        const float one = (j == lab) ? 1.f : 0.f;
        const float P = soft_row ? expf(z[j] - zl) : 0.f;
        const float iq = soft_row ? fmaxf((1.f - p.belta) * P + p.belta * one, 0.f) : one;
        float g = expf(ru[j] - lse) - iq;                              // sum_n IQ = 1
        if (self) g -= (1.f - p.belta) * P * (ru[j] - ps);
        p.g_u[(int64_t)m * p.N + j] = a * g;
      }
      if (lane == 0) p.row_part[p.M + m] = a * (lse - dot);
    }
  }
}

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = is_max ? -INFINITY : 0.f;
  for (int w = 0; w < nw; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

__global__ void __launch_bounds__(256)
loss_cols_kernel(const LossParams p) {
  extern __shared__ float smem_lc[];          // 3 columns of M + 8 reduction slots
  float* cn = smem_lc;
  float* cu = cn + p.M;
  float* cz = cu + p.M;
  float* red = cz + p.M;
  __shared__ int s_pick;
  const int i = blockIdx.x;
  const int tid = threadIdx.x;
  const bool has_z = p.soft != 0;
  for (int m = tid; m < p.M; m += blockDim.x) {
    cn[m] = p.s_n[(int64_t)m * p.N + i];
    cu[m] = p.s_u[(int64_t)m * p.N + i];
    cz[m] = has_z ? p.sims[(int64_t)m * p.N + i] : 0.f;
  }
  if (tid == 0) s_pick = -1;
  __syncthreads();
  // ---- triplet, video -> text
  float cnt_f = 0.f, possum = 0.f;
  for (int m = tid; m < p.M; m += blockDim.x)
    if (p.labels[m] == i) { cnt_f += 1.f; possum += cn[m]; }
  const float cnt = block_reduce(cnt_f, red, false);
  possum = block_reduce(possum, red, false);
  const int r = p.v2t_pick[i];
  for (int m = tid; m < p.M; m += blockDim.x) {
    if (p.labels[m] == i) continue;
    const float v = cn[m];
    int c = 0;
    for (int k = 0; k < p.M; ++k) {
      if (p.labels[k] == i) continue;
      const float w = cn[k];
      c += (w > v) || (w == v && k < m);
    }
    if (c == r) s_pick = m;                   // ranks are distinct: one writer
  }
  __syncthreads();
  const int pick = s_pick;
  float trip = 0.f;
  bool active = false;
  if (cnt == 0.f) {
    trip = __int_as_float(0x7fc00000);        // torch.mean of an empty slice (method/model.py:361) is nan
  } else if (pick >= 0) {
    const float d = p.margin + cn[pick] - possum / cnt;
    active = d > 0.f;
    trip = active ? d / (float)p.N : 0.f;
  }
  if (active) {
    const float g = 1.f / (float)p.N;
    for (int m = tid; m < p.M; m += blockDim.x) {
      const bool pos = p.labels[m] == i;
      if (pos || m == pick) p.g_n[(int64_t)m * p.N + i] += pos ? -g / cnt : g;
    }
  }
  // ---- NCE, video -> text (only videos that own a caption: label_dict keys)
  float nce = 0.f;
  if (cnt > 0.f) {
    float mxl = -INFINITY;
    for (int m = tid; m < p.M; m += blockDim.x) mxl = fmaxf(mxl, cu[m]);
    const float mx = block_reduce(mxl, red, true);
    float sl = 0.f;
    for (int m = tid; m < p.M; m += blockDim.x) sl += expf(cu[m] - mx);
    const float den = mx + logf(block_reduce(sl, red, false));
    if (!p.soft) {  // clip_nce: nominator = logsumexp over the video's own captions; mean over all N videos
      float pl = 0.f;
      for (int m = tid; m < p.M; m += blockDim.x)
        if (p.labels[m] == i) pl += expf(cu[m] - mx);
      const float nom = mx + logf(block_reduce(pl, red, false));
      const float b = 1.f / (float)p.N;
      for (int m = tid; m < p.M; m += blockDim.x) {
        const float wpos = (p.labels[m] == i) ? expf(cu[m] - nom) : 0.f;
        p.g_u[(int64_t)m * p.N + i] += b * (expf(cu[m] - den) - wpos);
      }
      nce = b * (den - nom);
    } else {
      int hardQ, hardV;
      float wh, ws;
      nce_parts(p, hardQ, hardV, wh, ws);
      const bool soft_col = i >= hardV;
      const float b = soft_col ? ws / (float)(p.N - hardV) : wh / (float)hardV;
      float zl = 0.f;
      if (soft_col) {
        float zm = -INFINITY;
        for (int m = tid; m < p.M; m += blockDim.x) zm = fmaxf(zm, cz[m]);
        const float zmx = block_reduce(zm, red, true);
        float zs = 0.f;
        for (int m = tid; m < p.M; m += blockDim.x) zs += expf(cz[m] - zmx);
        zl = zmx + logf(block_reduce(zs, red, false));
      }
      // nominator: logsumexp_m( log(IV[m] + 1e-12) + s[m] ), computed relative to mx (log(IV + eps) <= ~0)
      float nl = 0.f;
      for (int m = tid; m < p.M; m += blockDim.x) {
        const float one = (p.labels[m] == i) ? 1.f : 0.f;
        const float iv = soft_col ? fmaxf((1.f - p.belta) * expf(cz[m] - zl) + p.belta * one, 0.f) : one;
        nl += (iv + 1e-12f) * expf(cu[m] - mx);
      }
      const float nsum = block_reduce(nl, red, false);
      const float nom = mx + logf(nsum);
      const bool self = soft_col && (p.sims == p.s_u);
      float up = 0.f;                          // sum_m U[m] Pv[m],  U = W / (IV + eps) = exp(s - mx) / nsum
      if (self) {
        float ul = 0.f;
        for (int m = tid; m < p.M; m += blockDim.x) ul += expf(cu[m] - mx) / nsum * expf(cz[m] - zl);
        up = block_reduce(ul, red, false);
      }
      for (int m = tid; m < p.M; m += blockDim.x) {
        const float one = (p.labels[m] == i) ? 1.f : 0.f;
        const float Pv = soft_col ? expf(cz[m] - zl) : 0.f;
        const float iv = soft_col ? fmaxf((1.f - p.belta) * Pv + p.belta * one, 0.f) : one;
        const float e = expf(cu[m] - mx) / nsum;
        float g = expf(cu[m] - den) - (iv + 1e-12f) * e;
        if (self) g -= (1.f - p.belta) * Pv * (e - up);
        p.g_u[(int64_t)m * p.N + i] += b * g;
      }
      nce = b * (den - nom);
    }
  }
  if (tid == 0) {
    p.col_part[i] = trip;
    p.col_part[p.N + i] = nce;
  }
}

// out[0] = triplet, out[1] = nce: fixed-order sums of the row and column partials (one block).
__global__ void __launch_bounds__(256)
loss_reduce_kernel(const float* __restrict__ row_part, const float* __restrict__ col_part, int M, int N,
                   float* __restrict__ out) {
  __shared__ float red[8];
  for (int t = 0; t < 2; ++t) {
    float a = 0.f;
    for (int m = threadIdx.x; m < M; m += blockDim.x) a += row_part[t * M + m];
    for (int n = threadIdx.x; n < N; n += blockDim.x) a += col_part[t * N + n];
    const float s = block_reduce(a, red, false);
    if (threadIdx.x == 0) out[t] = s;
  }
}

}  // namespace dkd

extern "C" int64_t dkd_train_losses_workspace_floats(int32_t M, int32_t N) { return 2 * ((int64_t)M + N); }

extern "C" int dkd_train_losses(const float* s_n, const float* s_u, const float* sims, const int32_t* labels,
                                const int32_t* t2v_draw, const int32_t* v2t_pick, int32_t M, int32_t N,
                                float margin, int32_t soft, float alpha, float belta, float* out_terms,
                                float* g_n, float* g_u, float* workspace, void* stream) {
  if (!s_n || !s_u || !labels || !t2v_draw || !v2t_pick || !out_terms || !g_n || !g_u || !workspace) return DKD_ERR_ARG;
  if (soft && !sims) return DKD_ERR_ARG;
  if (M <= 0 || N <= 1 || N > 2048 || M > 8192) return DKD_ERR_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  LossParams p{s_n, s_u, soft ? sims : nullptr, labels, t2v_draw, v2t_pick, M, N, margin, alpha, belta, soft,
               g_n, g_u, workspace, workspace + 2 * (int64_t)M};
  const size_t smem_r = sizeof(float) * 4 * 2 * (size_t)N;
  const size_t smem_c = sizeof(float) * (3 * (size_t)M + 8);
  DKD_CUDA_TRY(cudaFuncSetAttribute(loss_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
  loss_rows_kernel<<<(M + 3) / 4, 128, smem_r, st>>>(p);
  DKD_LAUNCH_CHECK();
  loss_cols_kernel<<<N, 256, smem_c, st>>>(p);
  DKD_LAUNCH_CHECK();
  loss_reduce_kernel<<<1, 256, 0, st>>>(p.row_part, p.col_part, M, N, out_terms);
  DKD_LAUNCH_CHECK();
  return DKD_OK;
}
