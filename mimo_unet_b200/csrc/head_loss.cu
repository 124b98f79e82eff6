// Output heads, Laplace negative log-likelihood, loss-buffer softmax weighting and ensemble aggregation.
// All memory-bound: coalesced, vectorised where alignment allows, warp-shuffle + per-block partial reductions
// (finalised deterministically by a one-block kernel, never by float atomics across the grid).
//
// Replaces: OutConv (components.py:123-129), LaplaceNLL.forward (losses.py:132-164) and its autograd,
// _calculate_train_loss (mimo_unet.py:223-247), LossBuffer (loss_buffer.py:18-74) and compute_uncertainties
// (models/utils.py:76-101).
#include "common.cuh"
#include "ops.h"

namespace mimo {
namespace {

constexpr int kBlock = 256;

__device__ __forceinline__ float block_sum(float v, float* sh /*[32]*/) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;  // valid in warp 0
}

// ------------------------------------------------------------------------------------------------
// 1x1 head: out[b, s, k, h, w] = bias[k] + sum_c W[k][c] * feat[b, h, w, c]      (fp32 out, NCHW planes)
// ------------------------------------------------------------------------------------------------
__global__ void head_fwd_kernel(ActView f, const float* __restrict__ W, const float* __restrict__ bias, int K,
                                float* __restrict__ out, long long out_bstride /*elements between batch entries*/) {
  extern __shared__ float shw[];  // [K][C] + [K]
  const int C = f.C;
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) shw[i] = W[i];
  for (int i = threadIdx.x; i < K; i += blockDim.x) shw[K * C + i] = bias[i];
  __syncthreads();
  const long long HW = (long long)f.H * f.W, total = (long long)f.N * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % f.W), h = (int)((i / f.W) % f.H), n = (int)(i / HW);
    const bf16* src = f.base + f.pix(n, h, w);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = (k < K) ? shw[K * C + k] : 0.f;
    for (int c0 = 0; c0 < C; c0 += 8) {
      float v[8];
      load8(src + c0, min(8, C - c0), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (c0 + j < C) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < K) acc[k] = fmaf(v[j], shw[k * C + c0 + j], acc[k]);
        }
      }
    }
    float* dst = out + (long long)n * out_bstride + (long long)h * f.W + w;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < K) dst[k * HW] = acc[k];
  }
}

// Fast path (K in {1, 2, 4}, 16-byte aligned feature view): one thread per pixel, the pixel's channels are read with
// 16-byte loads, the weights sit in shared memory as [channel][K] so one LDS feeds K FMAs (the generic kernel above is
// instruction-issue bound: scalar loads of the partial channel group, one LDS per FMA, 64-bit index math; ncu IPC 3.1).
template <int K>
__global__ void __launch_bounds__(256)
head_fwd_fast_kernel(ActView f, const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ out,
                     long long out_bstride) {
  extern __shared__ float shw[];  // [groups*8][K], zero past C
  const int C = f.C, groups = (C + 7) >> 3;
  for (int i = threadIdx.x; i < groups * 8 * K; i += blockDim.x) {
    const int c = i / K, k = i - c * K;
    shw[i] = c < C ? W[k * C + c] : 0.f;
  }
  float b[K];
#pragma unroll
  for (int k = 0; k < K; ++k) b[k] = bias[k];
  __syncthreads();
  const uint4 tail_mask = group_mask(C - (groups - 1) * 8);
  const unsigned HW = (unsigned)f.H * (unsigned)f.W, total = (unsigned)f.N * HW;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned n = i / HW, r = i - n * HW;
    const unsigned h = r / (unsigned)f.W, w = r - h * (unsigned)f.W;
    const uint4* src = reinterpret_cast<const uint4*>(f.base + f.pix((int)n, (int)h, (int)w));
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = b[k];
    for (int g = 0; g < groups; ++g) {
      float v[8];
      uint4 r = src[g];
      if (g == groups - 1) r = and4(r, tail_mask);   // pad channels of the buffer may hold anything (0 * NaN)
      unpack8(r, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = fmaf(v[j], shw[(g * 8 + j) * K + k], acc[k]);
      }
    }
    float* dst = out + (long long)n * out_bstride + r;
#pragma unroll
    for (int k = 0; k < K; ++k) dst[(size_t)k * HW] = acc[k];
  }
}

// backward of the head: G[n,h,w,c] = gs * sum_k dOut[n,k,h,w] * W[k][c]  (bf16, unpadded NHWC)
//                       dW[k][c]  (+)= gs * sum_pix dOut * feat ;  db[k] (+)= gs * sum_pix dOut
// gs = *grad_scale (device scalar, e.g. the AMP loss scale) or 1.
__global__ void head_bwd_kernel(ActView f, const float* __restrict__ W, int K, const float* __restrict__ dout,
                                long long out_bstride, const float* __restrict__ grad_scale, ActView G,
                                float* __restrict__ part /*[gridDim.x][K*C + K]*/) {
  extern __shared__ float sh[];
  const int C = f.C;
  float* shw = sh;                 // [K][C]
  float* sd = shw + K * C;         // [K][kBlock]   dOut tile
  float* sf = sd + K * kBlock;     // [kBlock][C+1] feature tile
  const float gs = grad_scale ? *grad_scale : 1.f;
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) shw[i] = W[i];
  const long long HW = (long long)f.H * f.W, total = (long long)f.N * HW;
  const int pairs = K * C + K;
  // each thread owns up to 2 (k,c) accumulators (pairs <= 2*kBlock enforced by the host)
  float acc0 = 0.f, acc1 = 0.f;
  const long long tiles = (total + kBlock - 1) / kBlock;
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    __syncthreads();
    const long long i = t * kBlock + threadIdx.x;
    const bool ok = i < total;
    int w = 0, h = 0, n = 0;
    if (ok) { w = (int)(i % f.W); h = (int)((i / f.W) % f.H); n = (int)(i / HW); }
    float d[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      d[k] = (ok && k < K) ? gs * dout[(long long)n * out_bstride + k * HW + (long long)h * f.W + w] : 0.f;
      if (k < K) sd[k * kBlock + threadIdx.x] = d[k];
    }
    const bf16* src = f.base + f.pix(n, h, w);
    bf16* gdst = G.base + G.pix(n, h, w);
    for (int c0 = 0; c0 < C; c0 += 8) {
      const int nv = min(8, C - c0);
      float v[8], g[8];
      if (ok) load8(src + c0, nv, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (!ok) v[j] = 0.f;
        if (j < nv) sf[threadIdx.x * (C + 1) + c0 + j] = v[j];
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < K && j < nv) a = fmaf(d[k], shw[k * C + c0 + j], a);
        g[j] = a;
      }
      // pad channels of G are written as zeros (G.cpitch multiple of 8)
      if (ok) store8(gdst + c0, min(8, G.cpitch - c0), g);
    }
    __syncthreads();
    for (int rep = 0; rep < 2; ++rep) {
      const int pr = threadIdx.x + rep * kBlock;
      if (pr >= pairs) break;
      float a = 0.f;
      if (pr < K * C) {
        const int k = pr / C, c = pr - k * C;
        for (int p = 0; p < kBlock; ++p) a = fmaf(sd[k * kBlock + p], sf[p * (C + 1) + c], a);
      } else {
        const int k = pr - K * C;
        for (int p = 0; p < kBlock; ++p) a += sd[k * kBlock + p];
      }
      if (rep == 0) acc0 += a; else acc1 += a;
    }
  }
  if ((int)threadIdx.x < pairs) part[(size_t)blockIdx.x * pairs + threadIdx.x] = acc0;
  if ((int)threadIdx.x + kBlock < pairs) part[(size_t)blockIdx.x * pairs + threadIdx.x + kBlock] = acc1;
}

// Fast path of the head backward (K <= 4, 16-byte aligned views): block = (pixel lanes) x (8-channel groups), every
// thread keeps ONE channel group: W[k][c..c+8) and its dW / db accumulators live in registers (K*8 FMAs per pixel, no
// shared-memory tiles), G is written with one 16-byte store. The per-thread accumulators are combined through shared
// memory in a fixed order; the partial layout [block][K*C + K] is the one head_bwd_finalize_kernel reduces.
template <int K>
__global__ void __launch_bounds__(256, 3)
head_bwd_fast_kernel(ActView f, const float* __restrict__ W, const float* __restrict__ dout, long long out_bstride,
                     const float* __restrict__ grad_scale, ActView G, float* __restrict__ part) {
  extern __shared__ float sh[];  // [lanes][groups][K*8 + K]
  const int C = f.C;
  const int groups = (C + 7) >> 3;
  const int lanes = blockDim.x / groups;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups;
  const int c = g * 8, nv = min(8, C - c);
  const float gs = grad_scale ? *grad_scale : 1.f;
  const long long HW = (long long)f.H * f.W;
  float w[K][8], aw[K][8], ab[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    ab[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { w[k][j] = (j < nv) ? W[k * C + c + j] : 0.f; aw[k][j] = 0.f; }
  }
  if (pl < lanes) {
    const uint4 mask = group_mask(nv);
    const int rows = f.N * f.H;
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
      const int n = row / f.H, h = row - n * f.H;
      const bf16* src = f.base + f.pix(n, h, 0) + c;
      bf16* gdst = G.base + G.pix(n, h, 0) + c;
      const float* dp = dout + (long long)n * out_bstride + (long long)h * f.W;
      for (int x = pl; x < f.W; x += lanes) {
        const uint4 raw = and4(*reinterpret_cast<const uint4*>(src + (size_t)x * f.cpitch), mask);
        float d[K];
#pragma unroll
        for (int k = 0; k < K; ++k) d[k] = gs * dp[k * HW + x];
        float v[8], o[8];
        unpack8(raw, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a = 0.f;
#pragma unroll
          for (int k = 0; k < K; ++k) {
            a = fmaf(d[k], w[k][j], a);
            aw[k][j] = fmaf(d[k], v[j], aw[k][j]);
          }
          o[j] = a;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) ab[k] += d[k];
        *reinterpret_cast<uint4*>(gdst + (size_t)x * G.cpitch) = pack8(o);  // pad channels get zeros (w == 0 there)
      }
    }
    float* mine = sh + ((size_t)pl * groups + g) * (K * 8 + K);
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int j = 0; j < 8; ++j) mine[k * 8 + j] = aw[k][j];
      mine[K * 8 + k] = ab[k];
    }
  }
  __syncthreads();
  const int pairs = K * C + K;
  for (int i = threadIdx.x; i < pairs; i += blockDim.x) {
    float acc = 0.f;
    if (i < K * C) {
      const int k = i / C, ch = i - k * C;
      const float* srcp = sh + (size_t)(ch >> 3) * (K * 8 + K) + k * 8 + (ch & 7);
      for (int p = 0; p < lanes; ++p) acc += srcp[(size_t)p * groups * (K * 8 + K)];
    } else {
      const int k = i - K * C;
      const float* srcp = sh + K * 8 + k;  // group 0 of every pixel lane saw every pixel exactly once
      for (int p = 0; p < lanes; ++p) acc += srcp[(size_t)p * groups * (K * 8 + K)];
    }
    part[(size_t)blockIdx.x * pairs + i] = acc;
  }
}

__global__ void head_bwd_finalize_kernel(const float* __restrict__ part, int nparts, int K, int C, float* __restrict__ dW,
                                         float* __restrict__ db, int accumulate) {
  // one warp per entry: lanes stride over the partial rows (independent loads in flight), fixed-order butterfly
  const int pairs = K * C + K;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= pairs) return;
  const int lane = threadIdx.x & 31;
  double a = 0.0;
  for (int p = lane; p < nparts; p += 32) a += part[(size_t)p * pairs + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) {
    float* dst = (i < K * C) ? dW + i : db + (i - K * C);
    *dst = (accumulate ? *dst : 0.f) + (float)a;
  }
}

// ------------------------------------------------------------------------------------------------
// Laplace NLL element math
// ------------------------------------------------------------------------------------------------
// kind 0: LaplaceNLL (mimo/losses.py:132-164)   log(s) + |d| / s,   s = clamp(exp(log_s))
// kind 1: GaussianNLL (mimo/losses.py:48-79)    log(v) + d^2 / v,   v = clamp(exp(log_var))
// In both the clamp is applied in place under no_grad on a clone, i.e. it changes the VALUE the loss is evaluated at but is
// invisible to autograd: d/d log_p = dloss/dp (at the clamped value) * p_raw.
__device__ __forceinline__ void nll_elem(int kind, float mu, float ls, float y, float m, float eps_min, float eps_max, float* loss,
                                         float* g_mu, float* g_ls) {
  const float d = mu - y;
  const float s_raw = expf(ls);
  const float s_c = fminf(fmaxf(s_raw, eps_min), eps_max);
  const float inv = 1.f / s_c;
  if (kind == 0) {
    const float ad = fabsf(d);
    *loss = (logf(s_c) + ad * inv) * m;
    *g_mu = ((d > 0.f) ? inv : ((d < 0.f) ? -inv : 0.f)) * m;
    *g_ls = (inv - ad * inv * inv) * s_raw * m;
  } else {
    const float dd = d * d;
    *loss = (logf(s_c) + dd * inv) * m;
    *g_mu = 2.f * d * inv * m;
    *g_ls = (inv - dd * inv * inv) * s_raw * m;
  }
}
#define laplace_elem(...) nll_elem(kind, __VA_ARGS__)

// generic elementwise forward over [rows][cols] with per-tensor row strides (covers the strided p1/p2 views)
__global__ void laplace_fwd_kernel(const float* __restrict__ mu, long long mu_rs, const float* __restrict__ ls, long long ls_rs,
                                   const float* __restrict__ y, long long y_rs, const float* __restrict__ mask, long long m_rs,
                                   long long rows, long long cols, float eps_min, float eps_max, float* __restrict__ out_elem,
                                   float* __restrict__ part /*[gridDim.x] or null*/, int kind) {
  __shared__ float sh[32];
  const long long total = rows * cols;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols, c = i - r * cols;
    float l, a, b;
    laplace_elem(mu[r * mu_rs + c], ls[r * ls_rs + c], y[r * y_rs + c], mask ? mask[r * m_rs + c] : 1.f, eps_min, eps_max, &l, &a, &b);
    if (out_elem) out_elem[i] = l;
    acc += l;
  }
  if (part) {
    const float s = block_sum(acc, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
  }
}

__global__ void sum_partials_kernel(const float* __restrict__ part, int n, float scale, float* __restrict__ out) {
  __shared__ float sh[32];
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += part[i];
  const float s = block_sum(a, sh);
  if (threadIdx.x == 0) *out = s * scale;
}

// generic elementwise backward: g_mu = up * dl/dmu, g_ls = up * dl/dlog_s, up = upstream[i] (elementwise) or
// *upstream * scalar (reduce_mean path)
__global__ void laplace_bwd_kernel(const float* __restrict__ mu, long long mu_rs, const float* __restrict__ ls, long long ls_rs,
                                   const float* __restrict__ y, long long y_rs, const float* __restrict__ mask, long long m_rs,
                                   long long rows, long long cols, float eps_min, float eps_max, const float* __restrict__ up,
                                   int up_is_scalar, float up_scale, float* __restrict__ g_mu, float* __restrict__ g_ls, int kind) {
  const long long total = rows * cols;
  const float us = up_is_scalar ? (*up) * up_scale : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols, c = i - r * cols;
    float l, a, b;
    laplace_elem(mu[r * mu_rs + c], ls[r * ls_rs + c], y[r * y_rs + c], mask ? mask[r * m_rs + c] : 1.f, eps_min, eps_max, &l, &a, &b);
    const float u = up_is_scalar ? us : up[i];
    g_mu[i] = a * u;
    g_ls[i] = b * u;
  }
}

// ------------------------------------------------------------------------------------------------
// Deep evidential regression head + loss (reference mimo/models/evidential_unet.py:74-96, mimo/losses.py:195-271) for the
// M = 1, out_channels = 4 network: raw = (mu, log v, log alpha, log beta) per pixel.
//   head:  v = softplus(raw1), alpha = softplus(raw2) + 1, beta = softplus(raw3)          (nn.Softplus: threshold 20)
//   loss:  L = G(alpha - 1/2) / (4 G(alpha) v sqrt(beta)) * (2 beta (1 + v) + (2 alpha - 1) v d^2) + d^2 (2 alpha + v),  d = y - mu
// All tensors [B][4][HW] / [B][HW] fp32, contiguous. One elementwise pass each way.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }
// digamma for x > 0: recurrence up to x >= 6, then the asymptotic series
__device__ __forceinline__ float digamma_f(float x) {
  float r = 0.f;
  while (x < 6.f) { r -= 1.f / x; x += 1.f; }
  const float i = 1.f / x, i2 = i * i;
  return r + logf(x) - 0.5f * i - i2 * (1.f / 12.f - i2 * (1.f / 120.f - i2 * (1.f / 252.f)));
}

__global__ void evidential_head_kernel(const float* __restrict__ raw, float* __restrict__ out, long long B, long long HW) {
  const long long total = B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, p = i - b * HW;
    const float* r = raw + b * 4 * HW + p;
    float* o = out + b * 4 * HW + p;
    o[0] = r[0];
    o[HW] = softplus_f(r[HW]);
    o[2 * HW] = softplus_f(r[2 * HW]) + 1.f;
    o[3 * HW] = softplus_f(r[3 * HW]);
  }
}
__global__ void evidential_head_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ g_out, float* __restrict__ g_raw,
                                           long long B, long long HW) {
  const long long total = B * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, p = i - b * HW;
    const long long o = b * 4 * HW + p;
    g_raw[o] = g_out[o];
#pragma unroll
    for (int c = 1; c < 4; ++c) {
      const float x = raw[o + c * HW];
      g_raw[o + c * HW] = g_out[o + c * HW] * (x > 20.f ? 1.f : sigmoid_f(x));
    }
  }
}

__device__ __forceinline__ void evidential_elem(float mu, float v, float alpha, float beta, float y, float m, float* loss, float g[4]) {
  const float d = y - mu, dd = d * d;
  const float R = expf(lgammaf(alpha - 0.5f)) / expf(lgammaf(alpha));   // as the reference: Gamma(x) = exp(lgamma(x))
  const float coeff = R / (4.f * v * sqrtf(beta));
  const float second = 2.f * beta * (1.f + v) + (2.f * alpha - 1.f) * v * dd;
  *loss = (coeff * second + dd * (2.f * alpha + v)) * m;
  g[0] = (coeff * (2.f * alpha - 1.f) * v + (2.f * alpha + v)) * (-2.f * d) * m;
  g[1] = (-coeff / v * second + coeff * (2.f * beta + (2.f * alpha - 1.f) * dd) + dd) * m;
  g[2] = (coeff * (digamma_f(alpha - 0.5f) - digamma_f(alpha)) * second + coeff * 2.f * v * dd + 2.f * dd) * m;
  g[3] = (-coeff / (2.f * beta) * second + coeff * 2.f * (1.f + v)) * m;
}

// mode 0: elementwise loss (and, if part, per-block sums for the mean); mode 1: gradients w.r.t. the four parameters,
// scaled by up[i] (elementwise upstream) or *up * up_scale (mean)
__global__ void evidential_loss_kernel(const float* __restrict__ par, const float* __restrict__ y, const float* __restrict__ mask,
                                       long long B, long long HW, int mode, float* __restrict__ out_elem, float* __restrict__ part,
                                       const float* __restrict__ up, int up_is_scalar, float up_scale, float* __restrict__ g_par) {
  __shared__ float sh[32];
  const long long total = B * HW;
  const float us = (mode == 1 && up_is_scalar) ? (*up) * up_scale : 0.f;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, p = i - b * HW;
    const long long o = b * 4 * HW + p;
    float l, g[4];
    evidential_elem(par[o], par[o + HW], par[o + 2 * HW], par[o + 3 * HW], y[i], mask ? mask[i] : 1.f, &l, g);
    if (mode == 0) {
      if (out_elem) out_elem[i] = l;
      acc += l;
    } else {
      const float u = up_is_scalar ? us : up[i];
#pragma unroll
      for (int c = 0; c < 4; ++c) g_par[o + c * HW] = g[c] * u;
    }
  }
  if (mode == 0 && part) {
    const float sres = block_sum(acc, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = sres;
  }
}

// ------------------------------------------------------------------------------------------------
// Fused training loss: one pass over out[B,S,2C,H,W] and the labels
//   weights w[s] = S * softmax(mean_rows(buffer) / T)     (read BEFORE the new loss enters the buffer)
//   loss[s]     = mean_{b,c,h,w} l
//   dOut        = w[s] / (S * B*C*H*W) * dl/d(mu, log_s)  (gradient of mean_s(w_s * loss_s))
// grid = (blocks_per_row, B*S). Partials per (row, block) are finalised by laplace_train_finalize_kernel,
// which also advances the loss buffer -- no host round trip (the reference syncs twice per step here).
// ------------------------------------------------------------------------------------------------
struct LossBufferState {  // device-resident mirror of loss_buffer.py:LossBuffer
  int index;
  int size;
  int S;
  float temperature;
  // followed by float buf[size * S]
};

__device__ __forceinline__ void buffer_weights(const LossBufferState* st, const float* buf, float* w /*[S]*/) {
  // mean over ALL rows (zeros included, loss_buffer.py:61-62), softmax with temperature, times S
  const int S = st->S;
  float mx = -INFINITY;
  for (int s = 0; s < S; ++s) {
    float m = 0.f;
    for (int r = 0; r < st->size; ++r) m += buf[r * S + s];
    m = st->size > 0 ? m / (float)st->size : 0.f;
    w[s] = m / st->temperature;
    mx = fmaxf(mx, w[s]);
  }
  float den = 0.f;
  for (int s = 0; s < S; ++s) { w[s] = expf(w[s] - mx); den += w[s]; }
  for (int s = 0; s < S; ++s) w[s] = w[s] / den * (float)S;
}

// Persistent grid over (row = b*S + s, chunk) items: the loss-buffer weights are computed once per BLOCK (a serial
// softmax over the buffer rows, ~2 us of latency), not once per 16 KB chunk -- with one short-lived block per chunk the
// prologue cost as much as the streaming itself (ncu: 3.9 TB/s at 210 MB, 2.5 TB/s at 52 MB).
__global__ void __launch_bounds__(256, 4)
laplace_train_kernel(const float* __restrict__ out, const float* __restrict__ y, long long y_bs, long long y_ss,
                     const float* __restrict__ mask, long long m_bs, long long m_ss,
                     const long long* __restrict__ gather /*[S][B] or null*/, int B, int S, int C, long long HW,
                     float eps_min, float eps_max, const LossBufferState* __restrict__ lb,
                     const float* __restrict__ fixed_w /*[S] or null*/, float* __restrict__ dout,
                     float* __restrict__ part /*[B*S][nblk]*/, int nblk, float* __restrict__ mpart /*[B*S][nblk][4] or null*/, int kind) {
  __shared__ float sh[32];
  __shared__ float sw[64];
  if (threadIdx.x == 0) {
    if (fixed_w) { for (int s = 0; s < S; ++s) sw[s] = fixed_w[s]; }
    else if (lb) buffer_weights(lb, reinterpret_cast<const float*>(lb + 1), sw);
    else { for (int s = 0; s < S; ++s) sw[s] = 1.f; }
  }
  __syncthreads();
  const long long n = (long long)C * HW;
  const float inv_cnt = 1.f / ((float)S * (float)B * (float)n);
  const int items = B * S * nblk;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int row = item / nblk, chunk = item - row * nblk;  // row = b*S + s
    const int b = row / S, s = row - b * S;
    const float coef = sw[s] * inv_cnt;
    const long long src_b = gather ? gather[(long long)s * B + b] : b;
    const float* mu = out + (long long)row * 2 * n;
    const float* ls = mu + n;
    const float* yy = y + src_b * y_bs + s * y_ss;
    const float* mm = mask ? mask + src_b * m_bs + s * m_ss : nullptr;
    float* gmu = dout ? dout + (long long)row * 2 * n : nullptr;
    float* gls = dout ? gmu + n : nullptr;
    float acc = 0.f;
    // regression metrics of the location channel (reference mimo/metrics.py:22-34, called every step from
    // mimo_unet.py:135): sum |e|, sum e^2, sum y, sum y^2 with e = mu - y; the mask does not enter (as in the reference)
    float m_ae = 0.f, m_se = 0.f, m_y = 0.f, m_yy = 0.f;
#define MIMO_METRIC(MU, Y)                                   \
  {                                                          \
    const float e_ = (MU) - (Y);                             \
    m_ae += fabsf(e_); m_se = fmaf(e_, e_, m_se);            \
    m_y += (Y); m_yy = fmaf((Y), (Y), m_yy);                 \
  }
    const bool vec = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(mu) & 15) == 0) && ((reinterpret_cast<uintptr_t>(yy) & 15) == 0) &&
                     (!mm || (reinterpret_cast<uintptr_t>(mm) & 15) == 0) && (!dout || (reinterpret_cast<uintptr_t>(gmu) & 15) == 0);
    if (vec) {
      // chunk `chunk` of the row: 16-byte groups [lo, hi); two independent groups per thread and trip, streaming cache hints
      const long long n4 = n >> 2;
      const long long per = (n4 + nblk - 1) / nblk;
      const long long lo = chunk * per, hi = (lo + per < n4) ? lo + per : n4;
      for (long long i0 = lo + threadIdx.x; i0 < hi; i0 += 2 * blockDim.x) {
        const long long i1 = i0 + blockDim.x;
        const bool two = i1 < hi;
        const float4 a0 = __ldcs(reinterpret_cast<const float4*>(mu) + i0);
        const float4 l0 = __ldcs(reinterpret_cast<const float4*>(ls) + i0);
        const float4 t0 = __ldcs(reinterpret_cast<const float4*>(yy) + i0);
        const float4 m0 = mm ? __ldcs(reinterpret_cast<const float4*>(mm) + i0) : make_float4(1.f, 1.f, 1.f, 1.f);
        float4 a1 = a0, l1 = l0, t1 = t0, m1 = m0;
        if (two) {
          a1 = __ldcs(reinterpret_cast<const float4*>(mu) + i1);
          l1 = __ldcs(reinterpret_cast<const float4*>(ls) + i1);
          t1 = __ldcs(reinterpret_cast<const float4*>(yy) + i1);
          if (mm) m1 = __ldcs(reinterpret_cast<const float4*>(mm) + i1);
        }
        float4 ga, gl;
        float l;
        laplace_elem(a0.x, l0.x, t0.x, m0.x, eps_min, eps_max, &l, &ga.x, &gl.x); acc += l;
        laplace_elem(a0.y, l0.y, t0.y, m0.y, eps_min, eps_max, &l, &ga.y, &gl.y); acc += l;
        laplace_elem(a0.z, l0.z, t0.z, m0.z, eps_min, eps_max, &l, &ga.z, &gl.z); acc += l;
        laplace_elem(a0.w, l0.w, t0.w, m0.w, eps_min, eps_max, &l, &ga.w, &gl.w); acc += l;
        if (mpart) { MIMO_METRIC(a0.x, t0.x) MIMO_METRIC(a0.y, t0.y) MIMO_METRIC(a0.z, t0.z) MIMO_METRIC(a0.w, t0.w) }
        if (dout) {
          ga.x *= coef; ga.y *= coef; ga.z *= coef; ga.w *= coef;
          gl.x *= coef; gl.y *= coef; gl.z *= coef; gl.w *= coef;
          __stcs(reinterpret_cast<float4*>(gmu) + i0, ga);
          __stcs(reinterpret_cast<float4*>(gls) + i0, gl);
        }
        if (two) {
          laplace_elem(a1.x, l1.x, t1.x, m1.x, eps_min, eps_max, &l, &ga.x, &gl.x); acc += l;
          laplace_elem(a1.y, l1.y, t1.y, m1.y, eps_min, eps_max, &l, &ga.y, &gl.y); acc += l;
          laplace_elem(a1.z, l1.z, t1.z, m1.z, eps_min, eps_max, &l, &ga.z, &gl.z); acc += l;
          laplace_elem(a1.w, l1.w, t1.w, m1.w, eps_min, eps_max, &l, &ga.w, &gl.w); acc += l;
          if (mpart) { MIMO_METRIC(a1.x, t1.x) MIMO_METRIC(a1.y, t1.y) MIMO_METRIC(a1.z, t1.z) MIMO_METRIC(a1.w, t1.w) }
          if (dout) {
            ga.x *= coef; ga.y *= coef; ga.z *= coef; ga.w *= coef;
            gl.x *= coef; gl.y *= coef; gl.z *= coef; gl.w *= coef;
            __stcs(reinterpret_cast<float4*>(gmu) + i1, ga);
            __stcs(reinterpret_cast<float4*>(gls) + i1, gl);
          }
        }
      }
    } else {
      const long long per = (n + nblk - 1) / nblk;
      const long long lo = chunk * per, hi = (lo + per < n) ? lo + per : n;
      for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        float l, ga, gl;
        laplace_elem(mu[i], ls[i], yy[i], mm ? mm[i] : 1.f, eps_min, eps_max, &l, &ga, &gl);
        acc += l;
        if (mpart) MIMO_METRIC(mu[i], yy[i])
        if (dout) { gmu[i] = ga * coef; gls[i] = gl * coef; }
      }
    }
#undef MIMO_METRIC
    const float sres = block_sum(acc, sh);
    if (threadIdx.x == 0) part[(size_t)row * nblk + chunk] = sres;
    if (mpart) {
      const float r0 = block_sum(m_ae, sh), r1 = block_sum(m_se, sh), r2 = block_sum(m_y, sh), r3 = block_sum(m_yy, sh);
      if (threadIdx.x == 0) {
        float* mp = mpart + ((size_t)row * nblk + chunk) * 4;
        mp[0] = r0; mp[1] = r1; mp[2] = r2; mp[3] = r3;
      }
    }
  }
}

// one block: loss[s] = sum over (b, blocks) / (B*n); weights; weighted mean; buffer update
__global__ void laplace_train_finalize_kernel(const float* __restrict__ part, int nblk, int B, int S, double count,
                                              LossBufferState* lb, const float* __restrict__ fixed_w, int update_buffer,
                                              float* __restrict__ loss /*[S]*/, float* __restrict__ weights /*[S]*/,
                                              float* __restrict__ weighted /*[1]*/, const float* __restrict__ mpart,
                                              float* __restrict__ metrics /*[4]: mae, mse, rmse, r2*/) {
  __shared__ float sl[64];
  __shared__ double sm[8][4];
  // one warp per subnetwork (fixed lane-strided order + butterfly: deterministic), all subnetworks in parallel
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int s = warp; s < S; s += nwarps) {
    double a = 0.0;
    for (int i = lane; i < B * nblk; i += 32) {
      const int b = i / nblk, k = i - b * nblk;
      a += (double)part[(size_t)(b * S + s) * nblk + k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) sl[s] = (float)(a / count);
  }
  if (mpart && metrics) {
    // fixed thread-strided order + butterfly + fixed warp order: deterministic
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    const int total = B * S * nblk;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
#pragma unroll
      for (int k = 0; k < 4; ++k) a[k] += (double)mpart[(size_t)i * 4 + k];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
      if (lane == 0 && warp < 8) sm[warp][k] = a[k];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && mpart && metrics) {
    double t[4] = {0.0, 0.0, 0.0, 0.0};
    for (int w2 = 0; w2 < nwarps && w2 < 8; ++w2)
      for (int k = 0; k < 4; ++k) t[k] += sm[w2][k];
    const double nn = count * (double)S;   // all B*S*C*H*W elements
    const double mse = t[1] / nn;
    metrics[0] = (float)(t[0] / nn);
    metrics[1] = (float)mse;
    metrics[2] = (float)sqrt(mse);
    metrics[3] = (float)(1.0 - t[1] / (t[3] - t[2] * t[2] / nn));   // r2 = 1 - SS_res / SS_tot
  }
  if (threadIdx.x == 0) {
    float w[64];
    float* buf = lb ? reinterpret_cast<float*>(lb + 1) : nullptr;
    if (fixed_w) for (int s = 0; s < S; ++s) w[s] = fixed_w[s];
    else if (lb) buffer_weights(lb, buf, w);
    else for (int s = 0; s < S; ++s) w[s] = 1.f;
    float tot = 0.f;
    for (int s = 0; s < S; ++s) {
      loss[s] = sl[s];
      if (weights) weights[s] = w[s];
      tot += w[s] * sl[s];
    }
    if (weighted) *weighted = tot / (float)S;
    if (lb && update_buffer && lb->size > 0) {  // loss_buffer.py:43-52
      for (int s = 0; s < S; ++s) buf[lb->index * S + s] = sl[s];
      lb->index = (lb->index + 1) % lb->size;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Validation-step math in ONE pass over (p1, p2, label) (reference mimo_unet.py:146-183 with LaplaceNLL, losses.py:132-192;
// SURVEY 8 f2): per pixel of [B][inner] (inner = C * H * W) and all S subnetworks
//   nll_s      = log(clamp(exp(ls_s))) + |mu_s - y| / clamp(exp(ls_s))               (-> val_loss[s] = mean over b, inner)
//   mean       = sum_s mu_s / S,  alea = sum_s 2 exp(2 ls_s) / S,  epi = sum_s (mu_s - mean)^2 / (S - 1)   (utils.py:76-101)
//   param      = clamp(sqrt(alea + epi) / sqrt(2)),  nll_c = log(param) + |mean - y| / param                (-> val_loss_combined)
//   maps       mean, sqrt(alea), sqrt(epi), mean - y;   sums of |e|, e^2, y, y^2 (metrics.py:22-34) and of clip(std, 0, 5)
// The label is the un-repeated [B][inner] tensor (repeat_subnetworks is a stride-0 view of it: every subnetwork sees the same label).
// Partials: one row of S + 8 floats per block, combined in a fixed order by validation_finalize_kernel (deterministic).
// ------------------------------------------------------------------------------------------------
constexpr int kValExtra = 8;   // combined nll, sum|e|, sum e^2, sum y, sum y^2, sum clip(alea_std), sum clip(epi_std), (spare)

__global__ void __launch_bounds__(256)
validation_laplace_kernel(const float* __restrict__ p1, const float* __restrict__ p2, long long p_bs, long long p_ss,
                          const float* __restrict__ y, const float* __restrict__ mask, int B, int S, long long inner, float eps_min,
                          float eps_max, float* __restrict__ mean_o, float* __restrict__ alea_std_o, float* __restrict__ epi_std_o,
                          float* __restrict__ err_o, float* __restrict__ part) {
  extern __shared__ float red[];   // [warps][S + kValExtra]
  const int nacc = S + kValExtra;
  float comb = 0.f, sae = 0.f, sse = 0.f, sy = 0.f, syy = 0.f, sal = 0.f, sep = 0.f;
  float nll_acc[16];
#pragma unroll
  for (int s = 0; s < 16; ++s) nll_acc[s] = 0.f;
  const long long total = (long long)B * inner;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / inner, r = i - b * inner;
    const float yv = y[i];
    const float m = mask ? mask[i] : 1.f;
    const float* q1 = p1 + b * p_bs + r;
    const float* q2 = p2 + b * p_bs + r;
    float mu[16], sum = 0.f, al = 0.f;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
      if (s < S) {
        mu[s] = q1[s * p_ss];
        const float ls = q2[s * p_ss];
        const float sc = fminf(fmaxf(expf(ls), eps_min), eps_max);
        nll_acc[s] += (logf(sc) + fabsf(mu[s] - yv) / sc) * m;
        sum += mu[s];
        const float e2 = expf(ls) * 1.41421356237309515f;   // std = exp(log_scale) * sqrt(2), no clamp (losses.py:166-167)
        al += e2 * e2;
      }
    }
    const float mean = sum / (float)S;
    al /= (float)S;
    float ep = 0.f;
    if (S > 1) {
#pragma unroll
      for (int s = 0; s < 16; ++s)
        if (s < S) { const float d = mu[s] - mean; ep = fmaf(d, d, ep); }
      ep /= (float)(S - 1);
    }
    const float cstd = sqrtf(al + ep);
    const float param = fminf(fmaxf(cstd / 1.41421356237309515f, eps_min), eps_max);
    const float e = mean - yv;
    comb += (logf(param) + fabsf(e) / param) * m;
    const float as = sqrtf(al), es = sqrtf(ep);
    mean_o[i] = mean; alea_std_o[i] = as; epi_std_o[i] = es; err_o[i] = e;
    sae += fabsf(e); sse = fmaf(e, e, sse); sy += yv; syy = fmaf(yv, yv, syy);
    sal += fminf(fmaxf(as, 0.f), 5.f); sep += fminf(fmaxf(es, 0.f), 5.f);
  }
  // block reduction: lanes by shuffle, warps through shared memory in a fixed order
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  auto put = [&](int k, float v) {
    v = warp_sum(v);
    if (lane == 0) red[warp * nacc + k] = v;
  };
#pragma unroll
  for (int s = 0; s < 16; ++s)
    if (s < S) put(s, nll_acc[s]);
  put(S + 0, comb); put(S + 1, sae); put(S + 2, sse); put(S + 3, sy); put(S + 4, syy); put(S + 5, sal); put(S + 6, sep); put(S + 7, 0.f);
  __syncthreads();
  for (int k = threadIdx.x; k < nacc; k += blockDim.x) {
    float a = 0.f;
    for (int w = 0; w < nw; ++w) a += red[w * nacc + k];
    part[(size_t)blockIdx.x * nacc + k] = a;
  }
}

// scalars: [S] val_loss, then val_loss_combined, mae, mse, rmse, r2, mean clip(aleatoric_std, 0, 5), mean clip(epistemic_std, 0, 5)
__global__ void validation_finalize_kernel(const float* __restrict__ part, int nblk, int S, double count, float* __restrict__ scalars) {
  const int nacc = S + kValExtra;
  __shared__ double acc[16 + kValExtra];
  for (int k = threadIdx.x; k < nacc; k += blockDim.x) {
    double a = 0.0;
    for (int b = 0; b < nblk; ++b) a += (double)part[(size_t)b * nacc + k];
    acc[k] = a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) scalars[s] = (float)(acc[s] / count);
    const double comb = acc[S], sae = acc[S + 1], sse = acc[S + 2], sy = acc[S + 3], syy = acc[S + 4];
    scalars[S + 0] = (float)(comb / count);
    scalars[S + 1] = (float)(sae / count);
    const double mse = sse / count;
    scalars[S + 2] = (float)mse;
    scalars[S + 3] = (float)sqrt(mse);
    const double ss_tot = syy - sy * sy / count;           // torchmetrics r2_score: 1 - SS_res / SS_tot
    scalars[S + 4] = (float)(1.0 - sse / ss_tot);
    scalars[S + 5] = (float)(acc[S + 5] / count);
    scalars[S + 6] = (float)(acc[S + 6] / count);
  }
}

__global__ void lossbuffer_weights_kernel(const LossBufferState* lb, float* w) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float t[64];
    buffer_weights(lb, reinterpret_cast<const float*>(lb + 1), t);
    for (int s = 0; s < lb->S; ++s) w[s] = t[s];
  }
}
__global__ void lossbuffer_add_kernel(LossBufferState* lb, const float* loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0 && lb->size > 0) {
    float* buf = reinterpret_cast<float*>(lb + 1);
    for (int s = 0; s < lb->S; ++s) buf[lb->index * lb->S + s] = loss[s];
    lb->index = (lb->index + 1) % lb->size;
  }
}
__global__ void lossbuffer_init_kernel(LossBufferState* lb, int S, int size, float T) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { lb->index = 0; lb->size = size; lb->S = S; lb->temperature = T; }
  float* buf = reinterpret_cast<float*>(lb + 1);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S * size; i += gridDim.x * blockDim.x) buf[i] = 0.f;
}

__global__ void scale_by_scalar_kernel(float* __restrict__ x, long long n, const float* __restrict__ s) {
  const float v = *s;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] *= v;
}

// ------------------------------------------------------------------------------------------------
// Ensemble aggregation (models/utils.py:76-101): per output element over the S members
//   mean = 1/S sum mu ; aleatoric = 1/S sum 2*exp(2*log_s) ; epistemic = sum (mu-mean)^2 / (S-1)   (0 if S == 1)
// p1/p2 element (b, s, j) at base + b*bs + s*ss + j.
// ------------------------------------------------------------------------------------------------
__global__ void aggregate_kernel(const float* __restrict__ p1, long long p1_bs, long long p1_ss, const float* __restrict__ p2,
                                 long long p2_bs, long long p2_ss, int B, int S, long long inner, float* __restrict__ mean,
                                 float* __restrict__ alea, float* __restrict__ epi) {
  const long long total = (long long)B * inner;
  const float invS = 1.f / (float)S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / inner, j = i - b * inner;
    const float* a = p1 + b * p1_bs + j;
    const float* l = p2 + b * p2_bs + j;
    float m = 0.f, al = 0.f;
    for (int s = 0; s < S; ++s) {
      m += a[s * p1_ss];
      const float sd = expf(l[s * p2_ss]) * 1.41421356237309515f;  // losses.py:166-167
      al += sd * sd;
    }
    m *= invS;
    float e = 0.f;
    if (S > 1) {
      for (int s = 0; s < S; ++s) { const float d = a[s * p1_ss] - m; e += d * d; }
      e /= (float)(S - 1);
    }
    mean[i] = m; alea[i] = al * invS; epi[i] = e;
  }
}

// 16-byte variant: four consecutive output elements per thread. The epistemic variance uses the exact two-pass form
// (sum of squared deviations from the mean); the second pass re-reads the members from L1/L2 when S exceeds the
// register-resident limit, otherwise they stay in registers.
template <int SREG>
__global__ void __launch_bounds__(256)
aggregate_vec4_kernel(const float* __restrict__ p1, long long p1_bs, long long p1_ss, const float* __restrict__ p2,
                      long long p2_bs, long long p2_ss, int B, int S, long long inner4, float* __restrict__ mean,
                      float* __restrict__ alea, float* __restrict__ epi) {
  const long long total = (long long)B * inner4;
  const float invS = 1.f / (float)S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / inner4, j = (i - b * inner4) * 4;
    const float* a = p1 + b * p1_bs + j;
    const float* l = p2 + b * p2_bs + j;
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f), al = m, e = m;
    if (SREG > 0) {
      float4 va[SREG > 0 ? SREG : 1], vl[SREG > 0 ? SREG : 1];
#pragma unroll
      for (int s = 0; s < SREG; ++s) {
        va[s] = __ldcs(reinterpret_cast<const float4*>(a + s * p1_ss));
        vl[s] = __ldcs(reinterpret_cast<const float4*>(l + s * p2_ss));
      }
#pragma unroll
      for (int s = 0; s < SREG; ++s) {
        m.x += va[s].x; m.y += va[s].y; m.z += va[s].z; m.w += va[s].w;
        // (exp(l) * sqrt2)^2 = 2 exp(l)^2, evaluated like the reference: std first, then squared
        float sd;
        sd = expf(vl[s].x) * 1.41421356237309515f; al.x += sd * sd;
        sd = expf(vl[s].y) * 1.41421356237309515f; al.y += sd * sd;
        sd = expf(vl[s].z) * 1.41421356237309515f; al.z += sd * sd;
        sd = expf(vl[s].w) * 1.41421356237309515f; al.w += sd * sd;
      }
      m.x *= invS; m.y *= invS; m.z *= invS; m.w *= invS;
      if (SREG > 1) {
#pragma unroll
        for (int s = 0; s < SREG; ++s) {
          float d;
          d = va[s].x - m.x; e.x += d * d; d = va[s].y - m.y; e.y += d * d;
          d = va[s].z - m.z; e.z += d * d; d = va[s].w - m.w; e.w += d * d;
        }
        const float r = 1.f / (float)(SREG - 1);
        e.x *= r; e.y *= r; e.z *= r; e.w *= r;
      }
    } else {
      for (int s = 0; s < S; ++s) {
        const float4 va = *reinterpret_cast<const float4*>(a + s * p1_ss);
        const float4 vl = __ldcs(reinterpret_cast<const float4*>(l + s * p2_ss));
        m.x += va.x; m.y += va.y; m.z += va.z; m.w += va.w;
        float sd;
        sd = expf(vl.x) * 1.41421356237309515f; al.x += sd * sd;
        sd = expf(vl.y) * 1.41421356237309515f; al.y += sd * sd;
        sd = expf(vl.z) * 1.41421356237309515f; al.z += sd * sd;
        sd = expf(vl.w) * 1.41421356237309515f; al.w += sd * sd;
      }
      m.x *= invS; m.y *= invS; m.z *= invS; m.w *= invS;
      if (S > 1) {
        for (int s = 0; s < S; ++s) {
          const float4 va = *reinterpret_cast<const float4*>(a + s * p1_ss);
          float d;
          d = va.x - m.x; e.x += d * d; d = va.y - m.y; e.y += d * d;
          d = va.z - m.z; e.z += d * d; d = va.w - m.w; e.w += d * d;
        }
        const float r = 1.f / (float)(S - 1);
        e.x *= r; e.y *= r; e.z *= r; e.w *= r;
      }
    }
    al.x *= invS; al.y *= invS; al.z *= invS; al.w *= invS;
    __stcs(reinterpret_cast<float4*>(mean + i * 4), m);
    __stcs(reinterpret_cast<float4*>(alea + i * 4), al);
    __stcs(reinterpret_cast<float4*>(epi + i * 4), e);
  }
}

// Large member counts (MC dropout: S = 16, 32): one output element per thread, the S locations stay in registers between
// the mean and the squared-deviation pass (the vec4 kernel would need 4 S registers), all 2 S loads are independent.
template <int SREG>
__global__ void __launch_bounds__(256)
aggregate_reg_kernel(const float* __restrict__ p1, long long p1_bs, long long p1_ss, const float* __restrict__ p2, long long p2_bs,
                     long long p2_ss, int B, long long inner, float* __restrict__ mean, float* __restrict__ alea,
                     float* __restrict__ epi) {
  const long long total = (long long)B * inner;
  const float invS = 1.f / (float)SREG;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / inner, j = i - b * inner;
    const float* a = p1 + b * p1_bs + j;
    const float* l = p2 + b * p2_bs + j;
    float va[SREG], vl[SREG];
#pragma unroll
    for (int s = 0; s < SREG; ++s) { va[s] = __ldcs(a + s * p1_ss); vl[s] = __ldcs(l + s * p2_ss); }
    float m = 0.f, al = 0.f, e = 0.f;
#pragma unroll
    for (int s = 0; s < SREG; ++s) {
      m += va[s];
      const float sd = expf(vl[s]) * 1.41421356237309515f;
      al += sd * sd;
    }
    m *= invS;
#pragma unroll
    for (int s = 0; s < SREG; ++s) { const float d = va[s] - m; e += d * d; }
    __stcs(mean + i, m); __stcs(alea + i, al * invS); __stcs(epi + i, e / (float)(SREG - 1));
  }
}

inline int grid_for(long long work, int per_thread = 1) {
  long long g = ceil_div_ll(work, (long long)kBlock * per_thread);
  const long long cap = (long long)num_sms() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

// ===================================== launchers =====================================
int head_fwd_launch(const ActView& f, const float* W, const float* bias, int K, float* out, long long out_bstride, cudaStream_t st) {
  MIMO_CHECK(K >= 1 && K <= 8, MIMO_ERR_ARG, "head: out_channels must be in [1,8] (got %d)", K);
  const long long total = (long long)f.N * f.H * f.W;
  {
    const int groups = (f.C + 7) / 8;
    const bool aligned = ((uintptr_t)f.base % 16) == 0 && f.cpitch % 8 == 0 && f.c_off % 8 == 0 && f.c_off + groups * 8 <= f.cpitch;
    if (aligned && (K == 1 || K == 2 || K == 4) && total < (1ll << 31) && groups * 8 * K * sizeof(float) <= 40 * 1024) {
      const size_t shb = (size_t)groups * 8 * K * sizeof(float);
      const int grid = grid_for(total);
      if (K == 1) head_fwd_fast_kernel<1><<<grid, kBlock, shb, st>>>(f, W, bias, out, out_bstride);
      else if (K == 2) head_fwd_fast_kernel<2><<<grid, kBlock, shb, st>>>(f, W, bias, out, out_bstride);
      else head_fwd_fast_kernel<4><<<grid, kBlock, shb, st>>>(f, W, bias, out, out_bstride);
      MIMO_LAUNCH_CHECK();
      return MIMO_OK;
    }
  }
  head_fwd_kernel<<<grid_for(total), kBlock, (K * f.C + K) * sizeof(float), st>>>(f, W, bias, K, out, out_bstride);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int head_bwd_parts() { return 3 * num_sms(); }

int head_bwd_launch(const ActView& f, const float* W, int K, const float* dout, long long out_bstride, const float* grad_scale,
                    const ActView& G, float* part, float* dW, float* db, int accumulate, cudaStream_t st) {
  MIMO_CHECK(K >= 1 && K <= 8, MIMO_ERR_ARG, "head: out_channels must be in [1,8] (got %d)", K);
  MIMO_CHECK(K * f.C + K <= 2 * kBlock, MIMO_ERR_ARG, "head: K*C too large");
  MIMO_CHECK(G.pad == 0 && G.c_off == 0 && G.C == f.C, MIMO_ERR_ARG, "head_bwd: G view mismatch");
  const int nparts = head_bwd_parts();
  {
    // fast path: register accumulators, 16-byte loads / stores
    const int groups = (f.C + 7) / 8;
    const bool aligned = ((uintptr_t)f.base % 16) == 0 && f.cpitch % 8 == 0 && f.c_off % 8 == 0 && f.c_off + groups * 8 <= f.cpitch &&
                         ((uintptr_t)G.base % 16) == 0 && G.cpitch % 8 == 0 && groups * 8 <= G.cpitch;
    if (aligned && (K == 1 || K == 2 || K == 4) && groups <= kBlock) {
      const int rows = f.N * f.H;
      const int grid = rows < nparts ? rows : nparts;
      const int lanes = kBlock / groups;
      const size_t shb = (size_t)lanes * groups * (K * 8 + K) * sizeof(float);
      if (shb <= 48 * 1024) {
        if (K == 1) head_bwd_fast_kernel<1><<<grid, kBlock, shb, st>>>(f, W, dout, out_bstride, grad_scale, G, part);
        else if (K == 2) head_bwd_fast_kernel<2><<<grid, kBlock, shb, st>>>(f, W, dout, out_bstride, grad_scale, G, part);
        else head_bwd_fast_kernel<4><<<grid, kBlock, shb, st>>>(f, W, dout, out_bstride, grad_scale, G, part);
        MIMO_LAUNCH_CHECK();
        head_bwd_finalize_kernel<<<ceil_div(K * f.C + K, 8), 256, 0, st>>>(part, grid, K, f.C, dW, db, accumulate);
        MIMO_LAUNCH_CHECK();
        return MIMO_OK;
      }
    }
  }
  const size_t smem = ((size_t)K * f.C + (size_t)K * kBlock + (size_t)kBlock * (f.C + 1)) * sizeof(float);
  MIMO_CUDA(cudaFuncSetAttribute(head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));   // per device -> per launch
  MIMO_CHECK(smem <= 160 * 1024, MIMO_ERR_ARG, "head_bwd: feature count too large for shared memory");
  head_bwd_kernel<<<nparts, kBlock, smem, st>>>(f, W, K, dout, out_bstride, grad_scale, G, part);
  MIMO_LAUNCH_CHECK();
  head_bwd_finalize_kernel<<<ceil_div(K * f.C + K, 8), 256, 0, st>>>(part, nparts, K, f.C, dW, db, accumulate);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int laplace_parts() { return num_sms() * 8; }

int evidential_head_launch(const float* raw, float* out, long long B, long long HW, cudaStream_t st) {
  evidential_head_kernel<<<grid_for(B * HW, 4), kBlock, 0, st>>>(raw, out, B, HW);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}
int evidential_head_bwd_launch(const float* raw, const float* g_out, float* g_raw, long long B, long long HW, cudaStream_t st) {
  evidential_head_bwd_kernel<<<grid_for(B * HW, 4), kBlock, 0, st>>>(raw, g_out, g_raw, B, HW);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}
int evidential_loss_fwd_launch(const float* par, const float* y, const float* mask, long long B, long long HW, float* out_elem,
                               float* part, float* out_mean, cudaStream_t st) {
  const int grid = grid_for(B * HW, 4);
  evidential_loss_kernel<<<grid, kBlock, 0, st>>>(par, y, mask, B, HW, 0, out_elem, out_mean ? part : nullptr, nullptr, 0, 0.f, nullptr);
  MIMO_LAUNCH_CHECK();
  if (out_mean) {
    MIMO_CHECK(part != nullptr, MIMO_ERR_ARG, "evidential_loss_fwd: partial buffer required for the mean");
    sum_partials_kernel<<<1, kBlock, 0, st>>>(part, grid, (float)(1.0 / (double)(B * HW)), out_mean);
    MIMO_LAUNCH_CHECK();
  }
  return MIMO_OK;
}
int evidential_loss_bwd_launch(const float* par, const float* y, const float* mask, long long B, long long HW, const float* up,
                               int up_is_scalar, float up_scale, float* g_par, cudaStream_t st) {
  evidential_loss_kernel<<<grid_for(B * HW, 4), kBlock, 0, st>>>(par, y, mask, B, HW, 1, nullptr, nullptr, up, up_is_scalar, up_scale, g_par);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}



int laplace_fwd_launch(const float* mu, long long mu_rs, const float* ls, long long ls_rs, const float* y, long long y_rs,
                       const float* mask, long long m_rs, long long rows, long long cols, float eps_min, float eps_max,
                       float* out_elem, float* part, float* out_mean, cudaStream_t st, int kind) {
  const int grid = grid_for(rows * cols, 4);
  laplace_fwd_kernel<<<grid, kBlock, 0, st>>>(mu, mu_rs, ls, ls_rs, y, y_rs, mask, m_rs, rows, cols, eps_min, eps_max, out_elem,
                                              out_mean ? part : nullptr, kind);
  MIMO_LAUNCH_CHECK();
  if (out_mean) {
    MIMO_CHECK(part != nullptr, MIMO_ERR_ARG, "laplace_fwd: partial buffer required for the mean");
    sum_partials_kernel<<<1, kBlock, 0, st>>>(part, grid, (float)(1.0 / (double)(rows * cols)), out_mean);
    MIMO_LAUNCH_CHECK();
  }
  return MIMO_OK;
}

int laplace_bwd_launch(const float* mu, long long mu_rs, const float* ls, long long ls_rs, const float* y, long long y_rs,
                       const float* mask, long long m_rs, long long rows, long long cols, float eps_min, float eps_max,
                       const float* up, int up_is_scalar, float up_scale, float* g_mu, float* g_ls, cudaStream_t st, int kind) {
  laplace_bwd_kernel<<<grid_for(rows * cols, 4), kBlock, 0, st>>>(mu, mu_rs, ls, ls_rs, y, y_rs, mask, m_rs, rows, cols, eps_min,
                                                                 eps_max, up, up_is_scalar, up_scale, g_mu, g_ls, kind);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

size_t lossbuffer_bytes(int S, int size) { return sizeof(LossBufferState) + sizeof(float) * (size_t)S * (size > 0 ? size : 1); }

int lossbuffer_init_launch(void* state, int S, int size, float T, cudaStream_t st) {
  MIMO_CHECK(T > 0.f, MIMO_ERR_ARG, "Temperature should be positive.");
  MIMO_CHECK(S >= 1 && S <= 64, MIMO_ERR_ARG, "loss buffer supports 1..64 subnetworks");
  lossbuffer_init_kernel<<<1, 256, 0, st>>>((LossBufferState*)state, S, size, T);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}
int lossbuffer_weights_launch(const void* state, float* w, cudaStream_t st) {
  lossbuffer_weights_kernel<<<1, 32, 0, st>>>((const LossBufferState*)state, w);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}
int lossbuffer_add_launch(void* state, const float* loss, cudaStream_t st) {
  lossbuffer_add_kernel<<<1, 32, 0, st>>>((LossBufferState*)state, loss);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int laplace_train_blocks(long long n) {
  long long per_row = ceil_div_ll(n, (long long)kBlock * 16);
  if (per_row < 1) per_row = 1;
  if (per_row > 64) per_row = 64;
  return (int)per_row;
}

int laplace_train_launch(const float* out, const float* y, long long y_bs, long long y_ss, const float* mask, long long m_bs,
                         long long m_ss, const long long* gather, int B, int S, int C, long long HW, float eps_min, float eps_max,
                         void* lb_state, const float* fixed_w, int update_buffer, float* dout, float* part, float* loss,
                         float* weights, float* weighted, cudaStream_t st, float* mpart, float* metrics, int kind) {
  MIMO_CHECK(S >= 1 && S <= 64, MIMO_ERR_ARG, "laplace_train: 1..64 subnetworks supported");
  MIMO_CHECK((mpart == nullptr) == (metrics == nullptr), MIMO_ERR_ARG, "laplace_train: metrics need both the scratch and the output");
  const long long n = (long long)C * HW;
  const int nblk = laplace_train_blocks(n);
  const long long items = (long long)B * S * nblk;
  MIMO_CHECK(items < (1ll << 31), MIMO_ERR_ARG, "laplace_train: too many work items");
  int grid = 4 * num_sms();
  if (grid > items) grid = (int)items;
  laplace_train_kernel<<<grid, kBlock, 0, st>>>(out, y, y_bs, y_ss, mask, m_bs, m_ss, gather, B, S, C, HW, eps_min, eps_max,
                                                (const LossBufferState*)lb_state, fixed_w, dout, part, nblk, mpart, kind);
  MIMO_LAUNCH_CHECK();
  laplace_train_finalize_kernel<<<1, kBlock, 0, st>>>(part, nblk, B, S, (double)B * (double)n, (LossBufferState*)lb_state, fixed_w,
                                                      update_buffer, loss, weights, weighted, mpart, metrics);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int validation_scratch_floats(int S) { return 2 * num_sms() * (S + kValExtra); }

int validation_laplace_launch(const float* p1, const float* p2, long long p_bs, long long p_ss, const float* y, const float* mask, int B,
                              int S, long long inner, float eps_min, float eps_max, float* mean, float* alea_std, float* epi_std,
                              float* err, float* scratch, float* scalars, cudaStream_t st) {
  MIMO_CHECK(S >= 1 && S <= 16, MIMO_ERR_ARG, "validation: 1..16 subnetworks supported (got %d)", S);
  MIMO_CHECK(B >= 1 && inner >= 1, MIMO_ERR_ARG, "validation: empty input");
  const long long total = (long long)B * inner;
  int grid = 2 * num_sms();
  if ((long long)grid * kBlock > total) grid = (int)((total + kBlock - 1) / kBlock);
  const size_t smem = (size_t)(kBlock / 32) * (S + kValExtra) * sizeof(float);
  validation_laplace_kernel<<<grid, kBlock, smem, st>>>(p1, p2, p_bs, p_ss, y, mask, B, S, inner, eps_min, eps_max, mean, alea_std,
                                                        epi_std, err, scratch);
  MIMO_LAUNCH_CHECK();
  validation_finalize_kernel<<<1, 64, 0, st>>>(scratch, grid, S, (double)total, scalars);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int scale_by_scalar_launch(float* x, long long n, const float* s, cudaStream_t st) {
  scale_by_scalar_kernel<<<grid_for(n, 4), kBlock, 0, st>>>(x, n, s);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int aggregate_launch(const float* p1, long long p1_bs, long long p1_ss, const float* p2, long long p2_bs, long long p2_ss, int B,
                     int S, long long inner, float* mean, float* alea, float* epi, cudaStream_t st) {
  MIMO_CHECK(S >= 1, MIMO_ERR_ARG, "aggregate: S must be >= 1");
  const bool vec = (inner % 4) == 0 && (p1_bs % 4) == 0 && (p1_ss % 4) == 0 && (p2_bs % 4) == 0 && (p2_ss % 4) == 0 &&
                   (((uintptr_t)p1 | (uintptr_t)p2 | (uintptr_t)mean | (uintptr_t)alea | (uintptr_t)epi) % 16) == 0;
  if (S == 16 || S == 32) {
    const int grid = grid_for((long long)B * inner, 1);
    if (S == 16) aggregate_reg_kernel<16><<<grid, kBlock, 0, st>>>(p1, p1_bs, p1_ss, p2, p2_bs, p2_ss, B, inner, mean, alea, epi);
    else aggregate_reg_kernel<32><<<grid, kBlock, 0, st>>>(p1, p1_bs, p1_ss, p2, p2_bs, p2_ss, B, inner, mean, alea, epi);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
  }
  if (vec) {
    const long long inner4 = inner / 4;
    const int grid = grid_for((long long)B * inner4, 1);
#define MIMO_AGG(SR) aggregate_vec4_kernel<SR><<<grid, kBlock, 0, st>>>(p1, p1_bs, p1_ss, p2, p2_bs, p2_ss, B, S, inner4, mean, alea, epi)
    switch (S) {
      case 1: MIMO_AGG(1); break;
      case 2: MIMO_AGG(2); break;
      case 3: MIMO_AGG(3); break;
      case 4: MIMO_AGG(4); break;
      case 8: MIMO_AGG(8); break;
      default: MIMO_AGG(0); break;
    }
#undef MIMO_AGG
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
  }
  aggregate_kernel<<<grid_for((long long)B * inner, 2), kBlock, 0, st>>>(p1, p1_bs, p1_ss, p2, p2_bs, p2_ss, B, S, inner, mean, alea, epi);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

}  // namespace mimo
