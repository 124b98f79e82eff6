// Memory-bound NHWC bf16 kernels around the tensor-core convolutions: layout packing, BatchNorm
// (training statistics finalize / apply / backward), ReLU, Dropout2d, 2x2 max-pool, bilinear x2
// (align_corners=True) up-sampling, reflect-halo handling and their backward passes.
//
// Thread mapping everywhere: one thread = one pixel (or 2x2 pixel cell) x one group of 8 channels, channel
// groups fastest, so a warp touches consecutive 16-byte chunks (coalesced, vectorised when aligned).
#include "common.cuh"
#include "ops.h"

#include <stdlib.h>

namespace mimo {
namespace {

constexpr int kBlock = 256;

// MIMO_ELEM_REV=1 (default 0): the BN apply pass and the second BN-backward pass walk their tensors back to front, hoping for L2
// reuse between a producer's tail and the consumer's head. Measured on B200 (C2, tools/gpu/r2_rev.sh): 8.351 -> 8.342 ms/step, i.e.
// nothing: a 126 MB streaming write does not leave a useful tail in L2. Kept as an experiment knob.
inline int elementwise_reverse() {
  static const int v = [] { const char* e = getenv("MIMO_ELEM_REV"); return e ? atoi(e) : 0; }();
  return v;
}

inline int grid_for(long long work) {
  long long g = ceil_div_ll(work, kBlock);
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// Writes value vector to interior pixel (h,w) of a padded view and to every halo cell that mirrors it.
__device__ __forceinline__ void store_with_halo(const ActView& o, int n, int h, int w, int c, int nvalid, const float v[8]) {
  store8(o.base + o.pix(n, h, w) + c, nvalid, v);
  if (o.pad == 0) return;
  // reflect: halo row -1 mirrors row 1, halo row H mirrors row H-2 (same for columns)
  const int hh = (h == 1) ? -1 : ((h == o.H - 2) ? o.H : -2);
  const int ww = (w == 1) ? -1 : ((w == o.W - 2) ? o.W : -2);
  const int hh2 = (o.H == 3 && h == 1) ? o.H : -2;  // H == 3: row 1 mirrors into both halos
  const int ww2 = (o.W == 3 && w == 1) ? o.W : -2;
  const int hs[3] = {h, hh, hh2};
  const int ws[3] = {w, ww, ww2};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (hs[a] == -2) continue;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      if (ws[b] == -2 || (a == 0 && b == 0)) continue;
      store8(o.base + o.pix(n, hs[a], ws[b]) + c, nvalid, v);
    }
  }
}

// Packed variant for the hot kernels: `v` holds 8 bf16 (channels >= nvalid are zero). The halo logic only runs for
// pixels on the two outermost rings (a rarely taken branch).
// VEC: the address is 16-byte aligned. nvalid < 8 (last group of the view) is still written with one 16-byte store when
// `tail_ok`: the view spans its whole buffer, so the channels past C are pad channels nobody else owns (they get zeros).
template <bool VEC>
__device__ __forceinline__ void store_group(bf16* p, int nvalid, const uint4& v, bool tail_ok = false) {
  if (VEC && (nvalid == 8 || tail_ok)) {
    *reinterpret_cast<uint4*>(p) = v;
  } else {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    if ((reinterpret_cast<uintptr_t>(p) & 3) == 0) {
      // slices at an even channel offset (the 42-channel subnetwork slices of the stacked buffers): 32-bit stores of channel pairs
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (2 * i + 1 < nvalid) reinterpret_cast<uint32_t*>(p)[i] = w[i];
        else if (2 * i < nvalid) reinterpret_cast<unsigned short*>(p)[2 * i] = (unsigned short)w[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < nvalid) reinterpret_cast<unsigned short*>(p)[i] = (unsigned short)(w[i >> 1] >> ((i & 1) * 16));
    }
  }
}
template <bool VEC>
__device__ __forceinline__ void store_group_halo(const ActView& o, int n, int h, int w, int c, int nvalid, const uint4& v) {
  const bool tail_ok = o.c_off == 0 && ((o.C + 7) & ~7) == o.cpitch;
  store_group<VEC>(o.base + o.pix(n, h, w) + c, nvalid, v, tail_ok);
  if (o.pad != 1) return;
  if (h > 1 && h < o.H - 2 && w > 1 && w < o.W - 2) return;
  const int hh = (h == 1) ? -1 : ((h == o.H - 2) ? o.H : -2);
  const int ww = (w == 1) ? -1 : ((w == o.W - 2) ? o.W : -2);
  const int hh2 = (o.H == 3 && h == 1) ? o.H : -2;
  const int ww2 = (o.W == 3 && w == 1) ? o.W : -2;
  const int hs[3] = {h, hh, hh2};
  const int ws[3] = {w, ww, ww2};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (hs[a] == -2) continue;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      if (ws[b] == -2 || (a == 0 && b == 0)) continue;
      store_group<VEC>(o.base + o.pix(n, hs[a], ws[b]) + c, nvalid, v, tail_ok);
    }
  }
}
__device__ __forceinline__ uint32_t bf162_max_nan(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2_nan(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint4 max4_nan(const uint4& a, const uint4& b) {
  return make_uint4(bf162_max_nan(a.x, b.x), bf162_max_nan(a.y, b.y), bf162_max_nan(a.z, b.z), bf162_max_nan(a.w, b.w));
}

// ------------------------------------------------------------------------------------------------
// input packing: fp32 NCHW (strided) -> bf16 NHWC with reflect halo; optional batch gather
// (folds models/utils.py:38-41 index_select+stack of apply_input_transform into the load)
// ------------------------------------------------------------------------------------------------
__global__ void pack_input_kernel(const float* __restrict__ x, long long sb, long long sc, const long long* __restrict__ gather,
                                  ActView o) {
  const int Hp = o.H + 2 * o.pad, Wp = o.W + 2 * o.pad;
  const long long total = (long long)o.N * Hp * Wp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int wp = (int)(i % Wp), hp = (int)((i / Wp) % Hp), n = (int)(i / ((long long)Wp * Hp));
    const int h = reflect1(hp - o.pad, o.H), w = reflect1(wp - o.pad, o.W);
    const long long b = gather ? gather[n] : n;
    const float* src = x + b * sb + (long long)h * o.W + w;
    bf16* dst = o.base + ((long long)(n * Hp + hp) * Wp + wp) * o.cpitch + o.c_off;
    for (int c = 0; c < o.C; ++c) dst[c] = __float2bfloat16_rn(src[c * sc]);
  }
}

// OIHW fp32 -> bf16 [9][cout][cin_pitch] (fprop) and flipped/transposed [9][cin][cout_pitch] (dgrad)
__global__ void weight_pack_kernel(const float* __restrict__ w, int cout, int cin, bf16* __restrict__ wf, int cin_pitch,
                                   bf16* __restrict__ wd, int cout_pitch) {
  const int total_f = 9 * cout * cin_pitch;
  const int total_d = wd ? 9 * cin * cout_pitch : 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total_f + total_d; i += gridDim.x * blockDim.x) {
    if (i < total_f) {
      const int ci = i % cin_pitch, co = (i / cin_pitch) % cout, tap = i / (cin_pitch * cout);
      wf[i] = __float2bfloat16_rn(ci < cin ? w[((size_t)co * cin + ci) * 9 + tap] : 0.f);
    } else {
      const int j = i - total_f;
      const int co = j % cout_pitch, ci = (j / cout_pitch) % cin, tap = j / (cout_pitch * cin);
      // dgrad tap (a,b) uses W[kh=2-a][kw=2-b]  ->  flat tap index 8 - tap
      wd[j] = __float2bfloat16_rn(co < cout ? w[((size_t)co * cin + ci) * 9 + (8 - tap)] : 0.f);
    }
  }
}

constexpr int kMaxPackJobs = 48;
struct PackBatch { WeightPackJob job[kMaxPackJobs]; int first_block[kMaxPackJobs + 1]; int n; };

// all layers of the network in one launch: block ranges per layer, same element math as weight_pack_kernel. One index decode per
// (output channel, input position) pair, nine taps per thread (the per-element div / mod chain made the first version instruction
// bound: 69 us for 58 MB).
__global__ void weight_pack_batched_kernel(const PackBatch b) {
  int j = 0;
  while (j + 1 < b.n && (int)blockIdx.x >= b.first_block[j + 1]) ++j;
  const WeightPackJob& q = b.job[j];
  const int nblk = b.first_block[j + 1] - b.first_block[j];
  const int cout = q.cout, cin = q.cin, cin_pitch = q.cin_pitch, cout_pitch = q.cout_pitch;
  const int cin_phys = q.sl.phys_count(cin);
  const int pairs_f = cout * cin_pitch;                      // fprop pack: (co, cp), cp fastest
  const int pairs_d = q.wd ? cin_phys * cout_pitch : 0;      // dgrad pack: (cp, co), co fastest
  for (int i = ((int)blockIdx.x - b.first_block[j]) * blockDim.x + threadIdx.x; i < pairs_f + pairs_d; i += nblk * blockDim.x) {
    if (i < pairs_f) {
      const int co = i / cin_pitch, cp = i - co * cin_pitch;
      const int ci = q.sl.source(cp, cin);
      const float* src = q.w + ((size_t)co * cin + (ci >= 0 ? ci : 0)) * 9;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) q.wf[(size_t)tap * pairs_f + i] = __float2bfloat16_rn(ci >= 0 ? src[tap] : 0.f);
    } else {
      const int k = i - pairs_f;
      const int cp = k / cout_pitch, co = k - cp * cout_pitch;
      const int ci = q.sl.source(cp, cin);
      const bool ok = co < cout && ci >= 0;
      const float* src = q.w + ((size_t)(ok ? co : 0) * cin + (ok ? ci : 0)) * 9;
      // dgrad tap (a, b) uses W[kh = 2 - a][kw = 2 - b]  ->  flat tap index 8 - tap
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) q.wd[(size_t)tap * pairs_d + k] = __float2bfloat16_rn(ok ? src[8 - tap] : 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm statistics finalize (training): reduce per-tile partials, produce scale/shift, saved mean /
// invstd for backward, and update the running estimates (momentum 0.1, unbiased running variance,
// conv bias re-added to the running mean because the stored conv output is bias-free).
// One block per channel.
// ------------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const float* __restrict__ psum, const float* __restrict__ psq, int tiles, int cpitch, int C,
                                   double count, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ conv_bias, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, long long* __restrict__ nbt, float momentum, float eps, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  const int c = blockIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int t = threadIdx.x; t < tiles; t += blockDim.x) {
    s += (double)psum[(size_t)t * cpitch + c];
    q += (double)psq[(size_t)t * cpitch + c];
  }
  __shared__ double sh_s[32], sh_q[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh_s[warp] = s; sh_q[warp] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double S = 0.0, Q = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { S += sh_s[i]; Q += sh_q[i]; }
    const double mean = S / count;
    double var = Q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma[c], b = beta[c];
    scale[c] = g * invstd;
    shift[c] = b - (float)mean * g * invstd;
    save_mean[c] = (float)mean;
    save_invstd[c] = invstd;
    if (running_mean != nullptr) {
      const float bias = conv_bias ? conv_bias[c] : 0.f;
      const double unbiased = count > 1.0 ? var * (count / (count - 1.0)) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * ((float)mean + bias);
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
      if (c == 0 && nbt != nullptr) *nbt += 1;
    }
  }
}

// eval-mode affine from the running statistics: z = scale*y + shift with the conv bias folded in
__global__ void bn_eval_affine_kernel(int C, const float* gamma, const float* beta, const float* conv_bias, const float* rm,
                                      const float* rv, float eps, float* scale, float* shift, float* save_mean,
                                      float* save_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = rsqrtf(rv[c] + eps);
  const float bias = conv_bias ? conv_bias[c] : 0.f;
  scale[c] = gamma[c] * invstd;
  shift[c] = beta[c] + (bias - rm[c]) * gamma[c] * invstd;
  save_mean[c] = rm[c] - bias;  // so that x_hat = (y - save_mean) * invstd also holds in eval mode
  save_invstd[c] = invstd;
}

// ------------------------------------------------------------------------------------------------
// BN apply + ReLU (+ Dropout2d keep-scale) (+ 2x2 max-pool) with reflect-halo writes.
// One thread = one 2x2 pixel cell x 8 channels.
// ------------------------------------------------------------------------------------------------
// Block = 256 threads = (cell lanes) x (8-channel groups): every thread keeps ONE channel group (scale/shift in
// registers) and walks rows of 2x2 cells; 16-byte loads, packed bf16 stores, the 2x2 max is taken on the packed
// (stored) values with the NaN-propagating bf16x2 max.
template <bool VEC_O, bool VEC_P>
__global__ void __launch_bounds__(256, 2)
bn_relu_apply_kernel(const bf16* __restrict__ y, int ycp, const float* __restrict__ scale,
                     const float* __restrict__ shift, const float* __restrict__ drop /*[N][C] or null*/,
                     ActView o, ActView pool, int do_pool, int rev) {
  const int cells_h = (o.H + 1) >> 1, cells_w = (o.W + 1) >> 1;
  const int groups = (o.C + 7) >> 3;
  const int lanes = blockDim.x / groups;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups;
  if (pl >= lanes) return;
  const int c = g * 8, nv = min(8, o.C - c);
  const uint4 mask = group_mask(nv);
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc[k] = (k < nv) ? scale[c + k] : 0.f; sh[k] = (k < nv) ? shift[c + k] : 0.f; }
  const int cell_rows = o.N * cells_h;
  for (int cr0 = blockIdx.x; cr0 < cell_rows; cr0 += gridDim.x) {
    // rev: walk the tensor back to front, so the rows the producing convolution wrote LAST (still in L2) are read first and
    // the rows written last here are the ones the consuming convolution reads first
    const int cr = rev ? cell_rows - 1 - cr0 : cr0;
    const int n = cr / cells_h, chh = cr - n * cells_h;
    const int h0 = chh * 2;
    const bool row1 = h0 + 1 < o.H;
    const bf16* y0 = y + (size_t)(n * o.H + h0) * o.W * ycp + c;
    const bf16* y1 = y0 + (size_t)o.W * ycp;
    float dr[8];
    if (drop) {
#pragma unroll
      for (int k = 0; k < 8; ++k) dr[k] = k < nv ? drop[(size_t)n * o.C + c + k] : 0.f;
    }
    for (int cw = pl; cw < cells_w; cw += lanes) {
      const int w0 = cw * 2;
      const bool col1 = w0 + 1 < o.W;
      // issue all loads of the cell first (y is dense with ycp % 8 == 0: every 8-channel group is one aligned 16-byte load)
      uint4 raw[4];
      raw[0] = *reinterpret_cast<const uint4*>(y0 + (size_t)w0 * ycp);
      if (col1) raw[1] = *reinterpret_cast<const uint4*>(y0 + (size_t)(w0 + 1) * ycp);
      if (row1) raw[2] = *reinterpret_cast<const uint4*>(y1 + (size_t)w0 * ycp);
      if (row1 && col1) raw[3] = *reinterpret_cast<const uint4*>(y1 + (size_t)(w0 + 1) * ycp);
      uint4 mx = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int dy = j >> 1, dx = j & 1;
        if ((dy && !row1) || (dx && !col1)) continue;
        float v[8];
        unpack8(raw[j], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          v[k] = fmaxf(fmaf(v[k], sc[k], sh[k]), 0.f);
          if (drop) v[k] *= dr[k];
        }
        // the activation is STORED in bf16: pool over the stored value so indices/values agree with it
        const uint4 pk = and4(pack8(v), mask);
        mx = (j == 0) ? pk : max4_nan(mx, pk);
        store_group_halo<VEC_O>(o, n, h0 + dy, w0 + dx, c, nv, pk);
      }
      if (do_pool && row1 && col1 && chh < pool.H && cw < pool.W) store_group_halo<VEC_P>(pool, n, chh, cw, c, nv, mx);
    }
  }
}

// stand-alone 2x2 max-pool (component-level API; optional int64 indices in PyTorch's h*W+w convention,
// written for an NCHW [N][C][Hp][Wp] index tensor like nn.MaxPool2d(return_indices=True))
__global__ void maxpool_kernel(ActView in, ActView o, long long* __restrict__ idx_nchw) {
  const int groups = (o.C + 7) >> 3;
  const long long total = (long long)o.N * o.H * o.W * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long r = i / groups;
    const int w = (int)(r % o.W); r /= o.W;
    const int h = (int)(r % o.H);
    const int n = (int)(r / o.H);
    const int c = g * 8, nv = min(8, o.C - c);
    float mx[8]; int am[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { mx[k] = -INFINITY; am[k] = (2 * h) * in.W + 2 * w; }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float v[8];
        load8(in.base + in.pix(n, 2 * h + dy, 2 * w + dx) + c, nv, v);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (v[k] > mx[k] || v[k] != v[k]) { mx[k] = v[k]; am[k] = (2 * h + dy) * in.W + 2 * w + dx; }
      }
    store_with_halo(o, n, h, w, c, nv, mx);
    if (idx_nchw)
      for (int k = 0; k < nv; ++k) idx_nchw[(((size_t)n * o.C + c + k) * o.H + h) * o.W + w] = am[k];
  }
}

// ------------------------------------------------------------------------------------------------
// Reflect halo of a pad==1 view from its interior (row -1 := row 1, row H := row H-2, same for columns, corners from the
// mirrored rows): the fused inference epilogue of the conv kernels writes only interior pixels. One thread per (halo cell, 8-channel
// group); the halo is 2 (H + W + 2) of the (H+2)(W+2) cells, so this is a few percent of one pass over the tensor.
// ------------------------------------------------------------------------------------------------
__global__ void halo_fill_kernel(ActView v) {
  const int groups = (v.C + 7) >> 3;
  const int ring = 2 * (v.W + 2) + 2 * v.H;
  const long long total = (long long)v.N * ring * groups;
  const bool vec = ((v.c_off | v.cpitch) & 7) == 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long r = i / groups;
    const int k = (int)(r % ring);
    const int n = (int)(r / ring);
    int h, w;   // halo cell in interior coordinates (-1 .. H, -1 .. W)
    if (k < v.W + 2) { h = -1; w = k - 1; }
    else if (k < 2 * (v.W + 2)) { h = v.H; w = k - (v.W + 2) - 1; }
    else { const int t = k - 2 * (v.W + 2); h = t >> 1; w = (t & 1) ? v.W : -1; }
    const int sh = reflect1(h, v.H), sw = reflect1(w, v.W);
    const int c = g * 8, nv = min(8, v.C - c);
    const bf16* src = v.base + v.pix(n, sh, sw) + c;
    bf16* dst = v.base + v.pix(n, h, w) + c;
    if (vec && nv == 8) {
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
    } else {
      for (int j = 0; j < nv; ++j) dst[j] = src[j];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// bilinear x2, align_corners=True, then zero F.pad to the skip size (components.py:78,112-115), written
// into a channel slice of the (padded, reflect-halo) concat buffer.
// ------------------------------------------------------------------------------------------------
__global__ void upsample_kernel(ActView in, ActView o, int off_h, int off_w) {
  const int uh = 2 * in.H, uw = 2 * in.W;
  const float rh = uh > 1 ? (float)(in.H - 1) / (float)(uh - 1) : 0.f;
  const float rw = uw > 1 ? (float)(in.W - 1) / (float)(uw - 1) : 0.f;
  const int groups = (o.C + 7) >> 3;
  const long long total = (long long)o.N * o.H * o.W * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long r = i / groups;
    const int w = (int)(r % o.W); r /= o.W;
    const int h = (int)(r % o.H);
    const int n = (int)(r / o.H);
    const int c = g * 8, nv = min(8, o.C - c);
    float out[8];
    const int y = h - off_h, x = w - off_w;
    if (y < 0 || y >= uh || x < 0 || x >= uw) {
#pragma unroll
      for (int k = 0; k < 8; ++k) out[k] = 0.f;
    } else {
      const float sy = rh * y, sx = rw * x;
      const int y0 = (int)sy, x0 = (int)sx;
      const int y1 = min(y0 + 1, in.H - 1), x1 = min(x0 + 1, in.W - 1);
      const float ly = sy - y0, lx = sx - x0;
      float a[8], b[8], cc[8], d[8];
      load8(in.base + in.pix(n, y0, x0) + c, nv, a);
      load8(in.base + in.pix(n, y0, x1) + c, nv, b);
      load8(in.base + in.pix(n, y1, x0) + c, nv, cc);
      load8(in.base + in.pix(n, y1, x1) + c, nv, d);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        out[k] = (1.f - ly) * ((1.f - lx) * a[k] + lx * b[k]) + ly * ((1.f - lx) * cc[k] + lx * d[k]);
    }
    store_with_halo(o, n, h, w, c, nv, out);
  }
}

// ------------------------------------------------------------------------------------------------
// Decoder concat in ONE pass (reference components.py:110-119: up(x1), F.pad, cat([x2_skip, x1_up])): whole pixel lines of the
// haloed concat buffer `o` (c_off == 0) are written with 16-byte stores,
//   channels [0, f)          <- `skip`, the DENSE copy of the encoder output (f = skip.C),
//   channels [f, f + in.C)   <- bilinear x2 (align_corners) of `in`, zero outside the up-sampled area,
//   the rest of the pitch    <- 0.
// Why: with f = 21 the two producers used to write 42 and 84 bytes of every 128-byte line at different times of the forward
// pass; the partial-sector writes made bn_relu_apply of in_convs.c2 103 us (37 us for its dense twin) and the up-sampling 133 us
// (ncu: 104 MB read for a 63 MB input: sector fills). Thread = (pixel, 16-byte word k of the line); word k holds buffer channels
// [8k, 8k + 8): its first n_skip channels come from skip word k, the others are up channels t + i with t = 8k - f, i.e. the
// aligned source groups q = floor(t / 8) and q + 1 funnel-shifted by r = t - 8q (the same r for every word of the line).
// ------------------------------------------------------------------------------------------------
template <int WPP>
__global__ void __launch_bounds__(256, 3)
upsample_concat_kernel(ActView in, ActView skip, ActView o, int off_h, int off_w) {
  const int uh = 2 * in.H, uw = 2 * in.W;
  const float rh = uh > 1 ? (float)(in.H - 1) / (float)(uh - 1) : 0.f;
  const float rw = uw > 1 ? (float)(in.W - 1) / (float)(uw - 1) : 0.f;
  const int words = o.cpitch >> 3;
  constexpr int lanes = 256 / WPP;
  const int k = threadIdx.x % WPP, pl = threadIdx.x / WPP;
  const int f = skip.C, cup = in.C, gup = (cup + 7) >> 3;
  const int n_skip = min(max(f - 8 * k, 0), 8);
  const int t = 8 * k - f;
  const int q = t >= 0 ? (t >> 3) : -((-t + 7) >> 3);   // floor(t / 8)
  const int r = t - 8 * q;                                // 0 .. 7
  // every lane blends ONE aligned source group: `own` = q (r == 0) or q + 1 (r != 0: the word is the funnel of the left neighbour's
  // group and this lane's; a lane whose own group does not exist contributes zeros)
  const int g_own = r ? q + 1 : q;
  const bool use_own = g_own >= 0 && g_own < gup;
  const bool use_left = r != 0 && q >= 0 && q < gup;   // the left neighbour (word k - 1) owns group q
  const uint4 mask_o = group_mask(use_own ? min(8, cup - 8 * g_own) : 0);
  const uint4 mask_s = group_mask(n_skip);
  const bool active = k < words;
  const int rows = o.N * o.H;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / o.H, h = row - n * o.H;
    const int y = h - off_h;
    const bool yin = y >= 0 && y < uh;
    const float sy = rh * y;
    const int y0 = yin ? (int)sy : 0;
    const int y1 = min(y0 + 1, in.H - 1);
    const float ly = sy - y0;
    const bf16* r0 = in.base + in.pix(n, y0, 0);
    const bf16* r1 = in.base + in.pix(n, y1, 0);
    const bf16* srow = skip.base + skip.pix(n, h, 0) + 8 * k;
    // bilinear sample of this lane's source group at output column w (zeros outside the up-sampled area / for idle lanes)
    auto sample = [&](int w) -> uint4 {
      uint4 own = make_uint4(0, 0, 0, 0);
      const int x = w - off_w;
      if (active && w < o.W && use_own && yin && x >= 0 && x < uw) {
        const float sx = rw * x;
        const int x0 = (int)sx, x1 = min(x0 + 1, in.W - 1);
        const float lx = sx - x0;
        float a[8], b[8], cc[8], d[8], out[8];
        unpack8(and4(*reinterpret_cast<const uint4*>(r0 + (size_t)x0 * in.cpitch + 8 * g_own), mask_o), a);
        unpack8(and4(*reinterpret_cast<const uint4*>(r0 + (size_t)x1 * in.cpitch + 8 * g_own), mask_o), b);
        unpack8(and4(*reinterpret_cast<const uint4*>(r1 + (size_t)x0 * in.cpitch + 8 * g_own), mask_o), cc);
        unpack8(and4(*reinterpret_cast<const uint4*>(r1 + (size_t)x1 * in.cpitch + 8 * g_own), mask_o), d);
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i] = (1.f - ly) * ((1.f - lx) * a[i] + lx * b[i]) + ly * ((1.f - lx) * cc[i] + lx * d[i]);
        own = pack8(out);
      }
      return own;
    };
    // funnel with the left neighbour's group, merge the skip channels, store the pixel and the halo cells that mirror it
    auto emit = [&](int w, const uint4& own, const uint4& sw) {
      uint4 word = own;
      if (r != 0) {
        uint4 left;
        left.x = __shfl_up_sync(0xffffffffu, own.x, 1);
        left.y = __shfl_up_sync(0xffffffffu, own.y, 1);
        left.z = __shfl_up_sync(0xffffffffu, own.z, 1);
        left.w = __shfl_up_sync(0xffffffffu, own.w, 1);
        if (!use_left) left = make_uint4(0, 0, 0, 0);   // also covers word 0, whose left lane belongs to another pixel
        word = funnel8(left, own, r);
      }
      if (!active || w >= o.W) return;
      if (n_skip)
        word = make_uint4((sw.x & mask_s.x) | (word.x & ~mask_s.x), (sw.y & mask_s.y) | (word.y & ~mask_s.y),
                          (sw.z & mask_s.z) | (word.z & ~mask_s.z), (sw.w & mask_s.w) | (word.w & ~mask_s.w));
      if (h > 1 && h < o.H - 2 && w > 1 && w < o.W - 2) {   // interior: one target (the common case)
        *reinterpret_cast<uint4*>(o.base + o.pix(n, h, w) + 8 * k) = word;
        return;
      }
      // border: the pixel itself + the halo cells that mirror it
      const int hh = (h == 1) ? -1 : ((h == o.H - 2) ? o.H : -2);
      const int ww = (w == 1) ? -1 : ((w == o.W - 2) ? o.W : -2);
      const int hh2 = (o.H == 3 && h == 1) ? o.H : -2;
      const int ww2 = (o.W == 3 && w == 1) ? o.W : -2;
      const int hs[3] = {h, hh, hh2};
      const int ws[3] = {w, ww, ww2};
#pragma unroll 1
      for (int ai = 0; ai < 3; ++ai) {
        if (hs[ai] == -2) continue;
#pragma unroll 1
        for (int bi = 0; bi < 3; ++bi) {
          if (ws[bi] == -2) continue;
          *reinterpret_cast<uint4*>(o.base + o.pix(n, hs[ai], ws[bi]) + 8 * k) = word;
        }
      }
    };
    auto skip_word = [&](int w) -> uint4 {
      return (active && n_skip && w < o.W) ? *reinterpret_cast<const uint4*>(srow + (size_t)w * skip.cpitch) : make_uint4(0, 0, 0, 0);
    };
    // two pixels per thread and trip: ten independent 16-byte loads in flight (the kernel is load-latency bound otherwise);
    // uniform trip count: the shuffles need converged warps
    for (int w0 = 0; w0 < o.W; w0 += 2 * lanes) {
      const int wa = w0 + pl, wb = wa + lanes;
      const uint4 sa = skip_word(wa), sb = skip_word(wb);
      const uint4 oa = sample(wa);
      const uint4 ob = sample(wb);
      emit(wa, oa, sa);
      emit(wb, ob, sb);
    }
  }
}

// Fast path of the up-sampling: block = (pixel lanes) x GPP lanes per pixel (GPP = power of two >= #channel groups, so
// the groups of a pixel never straddle a warp), row-based 32-bit indexing, 16-byte loads of the four taps. Destination
// slices whose channel offset is NOT a multiple of 8 (decoder concat: 21 + 42 channels) are still written with 16-byte
// stores: each thread takes the missing leading channels of the next group from its neighbour lane (shuffle) and funnel-
// shifts the pair onto the buffer's 16-byte grid; only the first e = 8 - c_off % 8 channels of a pixel are scalar.
template <int GPP>
__global__ void __launch_bounds__(256, 3)
upsample_fast_kernel(ActView in, ActView o, int off_h, int off_w) {
  const int uh = 2 * in.H, uw = 2 * in.W;
  const float rh = uh > 1 ? (float)(in.H - 1) / (float)(uh - 1) : 0.f;
  const float rw = uw > 1 ? (float)(in.W - 1) / (float)(uw - 1) : 0.f;
  const int groups = (o.C + 7) >> 3;
  constexpr int lanes = 256 / GPP;
  const int g = threadIdx.x % GPP, pl = threadIdx.x / GPP;
  const int c = g * 8, nv = max(0, min(8, o.C - c));
  const uint4 mask = group_mask(nv);
  const int e = (8 - (o.c_off & 7)) & 7;                       // element shift onto the buffer's 16-byte grid
  const bool last_slice = o.c_off + o.C > o.cpitch - 8;         // the overhang of the last word falls on pad channels
  const bool tail_ok = e == 0 && o.c_off + ((o.C + 7) & ~7) <= o.cpitch && last_slice;
  const int rows = o.N * o.H;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / o.H, h = row - n * o.H;
    const int y = h - off_h;
    const bool yin = y >= 0 && y < uh;
    const float sy = rh * y;
    const int y0 = yin ? (int)sy : 0;
    const int y1 = min(y0 + 1, in.H - 1);
    const float ly = sy - y0;
    const bf16* r0 = in.base + in.pix(n, y0, 0) + c;
    const bf16* r1 = in.base + in.pix(n, y1, 0) + c;
    // bilinear sample of this thread's channel group at output column w (zero outside the up-sampled area / inactive lanes)
    auto sample = [&](int w) -> uint4 {
      uint4 own = make_uint4(0, 0, 0, 0);
      const int x = w - off_w;
      if (w < o.W && g < groups && yin && x >= 0 && x < uw) {
        const float sx = rw * x;
        const int x0 = (int)sx, x1 = min(x0 + 1, in.W - 1);
        const float lx = sx - x0;
        float a[8], b[8], cc[8], d[8], out[8];
        unpack8(load_group<true>(r0 + (size_t)x0 * in.cpitch, nv, mask), a);
        unpack8(load_group<true>(r0 + (size_t)x1 * in.cpitch, nv, mask), b);
        unpack8(load_group<true>(r1 + (size_t)x0 * in.cpitch, nv, mask), cc);
        unpack8(load_group<true>(r1 + (size_t)x1 * in.cpitch, nv, mask), d);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          out[k] = (1.f - ly) * ((1.f - lx) * a[k] + lx * b[k]) + ly * ((1.f - lx) * cc[k] + lx * d[k]);
        own = pack8(out);
      }
      return own;
    };
    // shuffle + funnel onto the buffer's 16-byte grid, then store the pixel and the halo cells that mirror it
    auto emit = [&](int w, const uint4& own) {
      uint4 word = own;
      if (e != 0) {
        uint4 next;
        next.x = __shfl_down_sync(0xffffffffu, own.x, 1);
        next.y = __shfl_down_sync(0xffffffffu, own.y, 1);
        next.z = __shfl_down_sync(0xffffffffu, own.z, 1);
        next.w = __shfl_down_sync(0xffffffffu, own.w, 1);
        if (g == GPP - 1) next = make_uint4(0, 0, 0, 0);  // the next lane belongs to another pixel
        word = funnel8(own, next, e);
      }
      if (!(w < o.W && g < groups)) return;
      const int valid = min(8, o.C - (c + e));  // view channels [c+e, c+e+8) carried by `word`
      const bool word_vec = valid == 8 || (valid > 0 && last_slice && o.c_off + c + e + 8 <= o.cpitch);
      auto put = [&](bf16* px) {
        if (e == 0) {
          store_group<true>(px + c, nv, own, tail_ok);
        } else {
          if (word_vec) *reinterpret_cast<uint4*>(px + c + e) = word;
          else if (valid > 0) store_group<false>(px + c + e, valid, word);
          if (g == 0) store_group<false>(px, min(e, o.C), own);  // leading channels [0, e) of the pixel
        }
      };
      if (o.pad != 1 || (h > 1 && h < o.H - 2 && w > 1 && w < o.W - 2)) {   // interior: one target (the common case)
        put(o.base + o.pix(n, h, w));
        return;
      }
      // border: the pixel itself + the halo cells that mirror it (outer two rings only)
      const int hh = (h == 1) ? -1 : ((h == o.H - 2) ? o.H : -2);
      const int ww = (w == 1) ? -1 : ((w == o.W - 2) ? o.W : -2);
      const int hh2 = (o.H == 3 && h == 1) ? o.H : -2;
      const int ww2 = (o.W == 3 && w == 1) ? o.W : -2;
      const int hs[3] = {h, hh, hh2};
      const int ws[3] = {w, ww, ww2};
#pragma unroll 1
      for (int ai = 0; ai < 3; ++ai) {
        if (hs[ai] == -2) continue;
#pragma unroll 1
        for (int bi = 0; bi < 3; ++bi) {
          if (ws[bi] == -2) continue;
          put(o.base + o.pix(n, hs[ai], ws[bi]));
        }
      }
    };
    // two pixels per thread and trip: eight independent 16-byte loads in flight (the kernel is load-latency bound otherwise);
    // uniform trip count: the shuffles need converged warps
    for (int w0 = 0; w0 < o.W; w0 += 2 * lanes) {
      const int wa = w0 + pl, wb = wa + lanes;
      const uint4 oa = sample(wa);
      const uint4 ob = sample(wb);
      emit(wa, oa);
      emit(wb, ob);
    }
  }
}

// backward of the above as a gather: every source pixel sums the destination pixels that sampled it
__global__ void upsample_bwd_kernel(ActView gdst /*unpadded, skip-sized*/, ActView gsrc /*unpadded*/, int off_h, int off_w,
                                    int accumulate) {
  const int ih = gsrc.H, iw = gsrc.W, uh = 2 * ih, uw = 2 * iw;
  const float rh = uh > 1 ? (float)(ih - 1) / (float)(uh - 1) : 0.f;
  const float rw = uw > 1 ? (float)(iw - 1) / (float)(uw - 1) : 0.f;
  const int groups = (gsrc.C + 7) >> 3;
  const long long total = (long long)gsrc.N * ih * iw * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long r = i / groups;
    const int sx = (int)(r % iw); r /= iw;
    const int sy = (int)(r % ih);
    const int n = (int)(r / ih);
    const int c = g * 8, nv = min(8, gsrc.C - c);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    // destination rows whose y0 or y1 equals sy lie in a small window around sy / rh
    const int ylo = max(0, (int)floorf((sy - 1) / fmaxf(rh, 1e-6f)) - 1), yhi = min(uh - 1, (int)ceilf((sy + 1) / fmaxf(rh, 1e-6f)) + 1);
    const int xlo = max(0, (int)floorf((sx - 1) / fmaxf(rw, 1e-6f)) - 1), xhi = min(uw - 1, (int)ceilf((sx + 1) / fmaxf(rw, 1e-6f)) + 1);
    for (int y = ylo; y <= yhi; ++y) {
      const float fy = rh * y;
      const int y0 = (int)fy, y1 = min(y0 + 1, ih - 1);
      const float ly = fy - y0;
      const float wy = (y0 == sy ? 1.f - ly : 0.f) + (y1 == sy ? ly : 0.f);
      if (wy == 0.f) continue;
      const int dh = y + off_h;
      if (dh < 0 || dh >= gdst.H) continue;
      for (int x = xlo; x <= xhi; ++x) {
        const float fx = rw * x;
        const int x0 = (int)fx, x1 = min(x0 + 1, iw - 1);
        const float lx = fx - x0;
        const float wx = (x0 == sx ? 1.f - lx : 0.f) + (x1 == sx ? lx : 0.f);
        if (wx == 0.f) continue;
        const int dw = x + off_w;
        if (dw < 0 || dw >= gdst.W) continue;
        float v[8];
        load8(gdst.base + gdst.pix(n, dh, dw) + c, nv, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += wy * wx * v[k];
      }
    }
    bf16* dst = gsrc.base + gsrc.pix(n, sy, sx) + c;
    if (accumulate) {
      float old[8];
      load8(dst, nv, old);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += old[k];
    }
    store8(dst, nv, acc);
  }
}

// Same gather with the candidate window cut to the destination pixels that can have a non-zero weight
// (|r * d - s| < 1 in both directions: at most 8 per direction for maps of 4+ pixels), the column weights hoisted out of
// the row loop and 16-byte loads: the generic kernel above spends ~50 candidate iterations per source pixel on index
// and weight arithmetic (ncu: issue-bound, IPC 3.1).
template <bool VEC>
__global__ void __launch_bounds__(256)
upsample_bwd_fast_kernel(ActView gdst, ActView gsrc, int off_h, int off_w, int accumulate) {
  const int ih = gsrc.H, iw = gsrc.W, uh = 2 * ih, uw = 2 * iw;
  const float rh = (float)(ih - 1) / (float)(uh - 1), rw = (float)(iw - 1) / (float)(uw - 1);
  const float inv_rh = 1.f / rh, inv_rw = 1.f / rw;
  const int groups = (gsrc.C + 7) >> 3;
  const unsigned total = (unsigned)gsrc.N * ih * iw * groups;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned g = i % groups;
    unsigned r = i / groups;
    const int sx = (int)(r % iw); r /= iw;
    const int sy = (int)(r % ih);
    const int n = (int)(r / ih);
    const int c = g * 8, nv = min(8, gsrc.C - c);
    const uint4 mask = group_mask(nv);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    const int ylo = max(0, (int)floorf((sy - 1) * inv_rh)), yhi = min(uh - 1, (int)ceilf((sy + 1) * inv_rh));
    const int xlo = max(0, (int)floorf((sx - 1) * inv_rw)), xhi = min(uw - 1, (int)ceilf((sx + 1) * inv_rw));
    float wxv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int x = xlo + j;
      const float fx = rw * x;
      const int x0 = (int)fx, x1 = min(x0 + 1, iw - 1);
      const float lx = fx - x0;
      const int dw = x + off_w;
      const bool ok = x <= xhi && dw >= 0 && dw < gdst.W;
      wxv[j] = ok ? ((x0 == sx ? 1.f - lx : 0.f) + (x1 == sx ? lx : 0.f)) : 0.f;
    }
    for (int y = ylo; y <= yhi; ++y) {
      const float fy = rh * y;
      const int y0 = (int)fy, y1 = min(y0 + 1, ih - 1);
      const float ly = fy - y0;
      const float wy = (y0 == sy ? 1.f - ly : 0.f) + (y1 == sy ? ly : 0.f);
      const int dh = y + off_h;
      if (wy == 0.f || dh < 0 || dh >= gdst.H) continue;
      const bf16* rowp = gdst.base + gdst.pix(n, dh, xlo + off_w) + c;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (wxv[j] != 0.f) {
          float v[8];
          unpack8(load_group<VEC>(rowp + (size_t)j * gdst.cpitch, nv, mask), v);
          const float w = wy * wxv[j];
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = fmaf(w, v[k], acc[k]);
        }
      }
    }
    bf16* dst = gsrc.base + gsrc.pix(n, sy, sx) + c;
    if (accumulate) {
      float old[8];
      load8(dst, nv, old);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += old[k];
    }
    store8(dst, nv, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// gradient gather: G[n,h,w,c] = fold(dpad)[h,w] (+ max-pool backward of gpool through act)
//   dpad  : dgrad output over the PADDED domain (H+2)x(W+2) (unpadded-style buffer of that size); the adjoint
//           of the reflect halo adds each halo cell onto the interior pixel it mirrors.
//   gpool : gradient w.r.t. the pooled map (unpadded, floor(H/2) x floor(W/2)); routed to the FIRST maximum of
//           each 2x2 window of `act` (row-major window order), like nn.MaxPool2d's backward.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fold_read(const ActView& dp /*H+2,W+2 unpadded*/, int n, int h, int w, int H, int W, int c, int nv,
                                          float acc[8]) {
  // rows of the padded domain that map to interior row h: h+1 always; 0 if h==1; H+1 if h==H-2
  int hs[3] = {h + 1, (h == 1) ? 0 : -1, (h == H - 2) ? H + 1 : -1};
  int ws[3] = {w + 1, (w == 1) ? 0 : -1, (w == W - 2) ? W + 1 : -1};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (hs[a] < 0) continue;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      if (ws[b] < 0) continue;
      float v[8];
      load8(dp.base + dp.pix(n, hs[a], ws[b]) + c, nv, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v[k];
    }
  }
}

// fold of the padded-domain gradient at interior pixel (h,w): dpad(h+1,w+1) plus, on the second ring, the halo cells that
// mirror it. `center` has already been loaded.
template <bool VEC>
__device__ __forceinline__ void fold_border(const ActView& dp, int n, int h, int w, int H, int W, int c, int nv, const uint4& mask,
                                            float acc[8]) {
  if (h != 1 && h != H - 2 && w != 1 && w != W - 2) return;
  const int hs[3] = {h + 1, (h == 1) ? 0 : -1, (h == H - 2) ? H + 1 : -1};
  const int ws[3] = {w + 1, (w == 1) ? 0 : -1, (w == W - 2) ? W + 1 : -1};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (hs[a] < 0) continue;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      if (ws[b] < 0 || (a == 0 && b == 0)) continue;
      float v[8];
      unpack8(load_group<VEC>(dp.base + dp.pix(n, hs[a], ws[b]) + c, nv, mask), v);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v[k];
    }
  }
}

// G = fold(dpad) (+ G when accumulating). Block = (pixel lanes) x (8-channel groups), one image row per block step.
template <bool VEC_IN, bool VEC_OUT>
__global__ void __launch_bounds__(256, 4)
grad_fold_kernel(ActView dpad, ActView dpad2, int two, ActView gout, int accumulate) {
  const int groups = (gout.C + 7) >> 3;
  const int lanes = blockDim.x / groups;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups;
  if (pl >= lanes) return;
  const int c = g * 8, nv = min(8, gout.C - c);
  const uint4 mask = group_mask(nv);
  const bool tail_ok = gout.c_off == 0 && ((gout.C + 7) & ~7) == gout.cpitch;
  const int rows = gout.N * gout.H;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / gout.H, h = row - n * gout.H;
    const bf16* src = dpad.base + dpad.pix(n, h + 1, 1) + c;
    const bf16* src2 = dpad2.base + dpad2.pix(n, h + 1, 1) + c;
    bf16* dst = gout.base + gout.pix(n, h, 0) + c;
    for (int w = pl; w < gout.W; w += lanes) {
      float acc[8];
      unpack8(load_group<VEC_IN>(src + (size_t)w * dpad.cpitch, nv, mask), acc);
      fold_border<VEC_IN>(dpad, n, h, w, gout.H, gout.W, c, nv, mask, acc);
      if (two) {   // sum of two sources (the same slice of two decoders' gradients) in one pass
        float t2[8];
        unpack8(load_group<VEC_IN>(src2 + (size_t)w * dpad2.cpitch, nv, mask), t2);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += t2[k];
        fold_border<VEC_IN>(dpad2, n, h, w, gout.H, gout.W, c, nv, mask, acc);
      }
      if (accumulate) {
        float old[8];
        unpack8(load_group<VEC_OUT>(dst + (size_t)w * gout.cpitch, nv, mask), old);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += old[k];
      }
      store_group<VEC_OUT>(dst + (size_t)w * gout.cpitch, nv, pack8(acc), tail_ok);
    }
  }
}

// G = [fold(dpad)] + max-pool backward of gpool through act, one thread per 2x2 cell x 8 channels: the winner of
// every window is found once (first maximum in row-major window order, like nn.MaxPool2d) instead of once per pixel.
template <bool VEC_D, bool VEC_A, bool VEC_P, bool VEC_OUT>
__global__ void __launch_bounds__(256, 3)
grad_gather_pool_kernel(ActView dpad, int has_dpad, ActView gpool, ActView act, ActView gout, int accumulate) {
  const int cells_h = (gout.H + 1) >> 1, cells_w = (gout.W + 1) >> 1;
  const int groups = (gout.C + 7) >> 3;
  const int lanes = blockDim.x / groups;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups;
  if (pl >= lanes) return;
  const int c = g * 8, nv = min(8, gout.C - c);
  const uint4 mask = group_mask(nv);
  const bool tail_ok = gout.c_off == 0 && ((gout.C + 7) & ~7) == gout.cpitch;
  const int cell_rows = gout.N * cells_h;
  for (int cr = blockIdx.x; cr < cell_rows; cr += gridDim.x) {
    const int n = cr / cells_h, chh = cr - n * cells_h;
    const int h0 = chh * 2;
    const bool row1 = h0 + 1 < gout.H;
    for (int cw = pl; cw < cells_w; cw += lanes) {
      const int w0 = cw * 2;
      const bool col1 = w0 + 1 < gout.W;
      const bool pooled = row1 && col1 && chh < gpool.H && cw < gpool.W;
      float out[4][8];
      // fold part
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int dy = j >> 1, dx = j & 1;
#pragma unroll
        for (int k = 0; k < 8; ++k) out[j][k] = 0.f;
        if ((dy && !row1) || (dx && !col1)) continue;
        if (has_dpad) {
          unpack8(load_group<VEC_D>(dpad.base + dpad.pix(n, h0 + dy + 1, w0 + dx + 1) + c, nv, mask), out[j]);
          fold_border<VEC_D>(dpad, n, h0 + dy, w0 + dx, gout.H, gout.W, c, nv, mask, out[j]);
        }
      }
      if (pooled) {
        float gp[8], a[4][8];
        unpack8(load_group<VEC_P>(gpool.base + gpool.pix(n, chh, cw) + c, nv, mask), gp);
#pragma unroll
        for (int j = 0; j < 4; ++j) unpack8(load_group<VEC_A>(act.base + act.pix(n, h0 + (j >> 1), w0 + (j & 1)) + c, nv, mask), a[j]);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          int win = 0;
          float m = a[0][k];
          if (a[1][k] > m) { m = a[1][k]; win = 1; }
          if (a[2][k] > m) { m = a[2][k]; win = 2; }
          if (a[3][k] > m) { m = a[3][k]; win = 3; }
#pragma unroll
          for (int j = 0; j < 4; ++j) out[j][k] += (win == j) ? gp[k] : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int dy = j >> 1, dx = j & 1;
        if ((dy && !row1) || (dx && !col1)) continue;
        bf16* dst = gout.base + gout.pix(n, h0 + dy, w0 + dx) + c;
        if (accumulate) {
          float old[8];
          unpack8(load_group<VEC_OUT>(dst, nv, mask), old);
#pragma unroll
          for (int k = 0; k < 8; ++k) out[j][k] += old[k];
        }
        store_group<VEC_OUT>(dst, nv, pack8(out[j]), tail_ok);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// BN + ReLU (+dropout) backward.
//   dz = G * drop * [scale*y + shift > 0]
//   pass 1 (reduce): per-block partials of sum dz and sum dz*y; the finalize kernel forms (in double)
//                    s1 = sum dz, s2 = sum dz * x_hat = invstd * (sum dz*y - mean * s1), x_hat = (y - mean) * invstd
//   pass 2 (apply) : dy = scale * (dz - (s1 + x_hat * s2) / count)     [training]
//                    dy = scale * dz                                     [eval: running stats are constants]
// ------------------------------------------------------------------------------------------------
// Block = 256 threads = (pixel lanes) x (8-channel groups), every thread keeps ONE channel group for the whole kernel
// (per-channel coefficients live in registers) and walks whole image rows (row = n*H + h): no per-element 64-bit index
// arithmetic, 16-byte loads, ~9 instructions per element. Per-thread accumulators are combined in shared memory in a
// fixed order (deterministic), one partial row per block.
template <bool VEC>
__global__ void __launch_bounds__(256, 3)
bn_bwd_reduce_kernel(ActView G, const bf16* __restrict__ y, int ycp, const float* __restrict__ scale,
                     const float* __restrict__ shift, const float* __restrict__ drop, int C,
                     float* __restrict__ part /*[gridDim.x][2][C]: sum dz, sum dz*y*/) {
  extern __shared__ float sh[];  // [pix_lanes][groups*16]
  const int groups = (C + 7) >> 3;
  const int pix_lanes = blockDim.x / groups;  // >= 1 (host guarantees groups <= blockDim.x)
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups;
  const int rows = G.N * G.H;
  const int c = g * 8, nv = min(8, C - c);
  float a1[8], a2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { a1[k] = 0.f; a2[k] = 0.f; }
  if (pl < pix_lanes) {
    const uint4 mask = group_mask(nv);
    float sc[8], sf[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = k < nv ? scale[c + k] : 0.f; sf[k] = k < nv ? shift[c + k] : 0.f; }
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
      const int n = row / G.H, h = row - n * G.H;
      const bf16* gp = G.base + G.pix(n, h, 0) + c;
      const bf16* yp = y + (size_t)row * G.W * ycp + c;
      float dr[8];
      if (drop) {
#pragma unroll
        for (int k = 0; k < 8; ++k) dr[k] = k < nv ? drop[(size_t)n * C + c + k] : 0.f;
      }
      for (int w = pl; w < G.W; w += 2 * pix_lanes) {
        const int w2 = w + pix_lanes;
        const bool has2 = w2 < G.W;
        // issue all loads first
        const uint4 gr = load_group<VEC>(gp + (size_t)w * G.cpitch, nv, mask);
        const uint4 yr = load_group<true>(yp + (size_t)w * ycp, nv, mask);
        uint4 gr2 = make_uint4(0, 0, 0, 0), yr2 = make_uint4(0, 0, 0, 0);
        if (has2) {
          gr2 = load_group<VEC>(gp + (size_t)w2 * G.cpitch, nv, mask);
          yr2 = load_group<true>(yp + (size_t)w2 * ycp, nv, mask);
        }
        float gv[8], yv[8];
        unpack8(gr, gv);
        unpack8(yr, yv);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float z = fmaf(yv[k], sc[k], sf[k]);
          float gg = gv[k];
          if (drop) gg *= dr[k];
          const float dz = (z > 0.f) ? gg : 0.f;
          a1[k] += dz;
          a2[k] = fmaf(dz, yv[k], a2[k]);
        }
        unpack8(gr2, gv);   // all-zero when !has2: contributes nothing
        unpack8(yr2, yv);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float z = fmaf(yv[k], sc[k], sf[k]);
          float gg = gv[k];
          if (drop) gg *= dr[k];
          const float dz = (z > 0.f) ? gg : 0.f;
          a1[k] += dz;
          a2[k] = fmaf(dz, yv[k], a2[k]);
        }
      }
    }
    float* mine = sh + (size_t)pl * groups * 16 + g * 16;
#pragma unroll
    for (int k = 0; k < 8; ++k) { mine[k] = a1[k]; mine[8 + k] = a2[k]; }
  }
  __syncthreads();
  // fixed-order combine over the pixel lanes: thread i < 2*C owns (which = i / C, channel = i % C)
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    const int which = i / C, ch = i - which * C;
    const float* src = sh + (ch >> 3) * 16 + which * 8 + (ch & 7);
    float acc = 0.f;
    for (int p = 0; p < pix_lanes; ++p) acc += src[(size_t)p * groups * 16];
    part[(size_t)blockIdx.x * 2 * C + i] = acc;
  }
}

// reduce partials -> s1,s2 ; write dgamma (= s2), dbeta (= s1), dbias_conv (eval only: scale * s1).
// Block = 8 channels x 32 partial lanes, combined in a fixed order.
__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(const float* __restrict__ part, int nparts, int C, float* __restrict__ s1s2 /*[2][C]*/,
                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias,
                       const float* __restrict__ scale, const float* __restrict__ mean, const float* __restrict__ invstd,
                       int training, float grad_scale, int accumulate) {
  __shared__ double sa[32][8], sb[32][8];
  const int cl = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + cl;
  double a = 0.0, b = 0.0;
  if (c < C) {
    // independent loads, unrolled so their latencies overlap (the grid is only C/8 blocks)
    int p = pl;
    for (; p + 96 < nparts; p += 128) {
      float va[4], vb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { va[u] = part[(size_t)(p + 32 * u) * 2 * C + c]; vb[u] = part[(size_t)(p + 32 * u) * 2 * C + C + c]; }
#pragma unroll
      for (int u = 0; u < 4; ++u) { a += va[u]; b += vb[u]; }
    }
    for (; p < nparts; p += 32) { a += part[(size_t)p * 2 * C + c]; b += part[(size_t)p * 2 * C + C + c]; }
  }
  sa[pl][cl] = a; sb[pl][cl] = b;
  __syncthreads();
  if (pl != 0 || c >= C) return;
  a = 0.0; b = 0.0;
  for (int p = 0; p < 32; ++p) { a += sa[p][cl]; b += sb[p][cl]; }
  b = (double)invstd[c] * (b - (double)mean[c] * a);
  s1s2[c] = (float)a; s1s2[C + c] = (float)b;
  if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)b * grad_scale;
  if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)a * grad_scale;
  // training: the conv bias cancels inside BatchNorm -> exactly zero gradient
  if (dbias) dbias[c] = (accumulate ? dbias[c] : 0.f) + (training ? 0.f : (float)a * scale[c] * grad_scale);
}

// dy = scale * (dz - (s1 + x_hat * s2) / count) rewritten as  dy = scale * dz + A * y + B  with per-channel
//   A = -scale * s2 * invstd / count,  B = -scale * (s1 - s2 * invstd * mean) / count   (A = B = 0 in eval mode),
// so the per-element work is 3 FMAs + one select. Same thread layout as the reduce kernel.
template <bool VEC>
__global__ void __launch_bounds__(256, 4)
bn_bwd_apply_kernel(ActView G, const bf16* __restrict__ y, int ycp, const float* __restrict__ scale,
                    const float* __restrict__ shift, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ drop,
                    const float* __restrict__ s1s2, int C, float inv_count, int training, ActView dy) {
  const int groups = (C + 7) >> 3;
  const int pix_lanes = blockDim.x / groups;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups;
  if (pl >= pix_lanes) return;
  const int rows = G.N * G.H;
  const int c = g * 8, nv = min(8, C - c);
  const uint4 mask = group_mask(nv);
  float sc[8], sf[8], ca[8], cb[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = k < nv ? scale[c + k] : 0.f;
    sf[k] = k < nv ? shift[c + k] : 0.f;
    ca[k] = 0.f; cb[k] = 0.f;
    if (training && k < nv) {
      const float is = invstd[c + k], s1 = s1s2[c + k], s2 = s1s2[C + c + k];
      ca[k] = -sc[k] * s2 * is * inv_count;
      cb[k] = -sc[k] * (s1 - s2 * is * mean[c + k]) * inv_count;
    }
  }
  const bool vec_store = c + 8 <= dy.cpitch;  // pad channels of dy are written as zeros (the tensor-core kernels read them)
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / G.H, h = row - n * G.H;
    const bf16* gp = G.base + G.pix(n, h, 0) + c;
    const bf16* yp = y + (size_t)row * G.W * ycp + c;
    bf16* dp = dy.base + dy.pix(n, h, 0) + c;
    float dr[8];
    if (drop) {
#pragma unroll
      for (int k = 0; k < 8; ++k) dr[k] = k < nv ? drop[(size_t)n * C + c + k] : 0.f;
    }
    for (int w = pl; w < G.W; w += pix_lanes) {
      const uint4 gr = load_group<VEC>(gp + (size_t)w * G.cpitch, nv, mask);
      const uint4 yr = load_group<true>(yp + (size_t)w * ycp, nv, mask);
      float gv[8], yv[8], out[8];
      unpack8(gr, gv);
      unpack8(yr, yv);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float z = fmaf(yv[k], sc[k], sf[k]);
        float gg = gv[k];
        if (drop) gg *= dr[k];
        const float dz = (z > 0.f) ? gg : 0.f;
        out[k] = fmaf(sc[k], dz, fmaf(ca[k], yv[k], cb[k]));
      }
      if (vec_store) *reinterpret_cast<uint4*>(dp + (size_t)w * dy.cpitch) = pack8(out);
      else store8(dp + (size_t)w * dy.cpitch, min(8, dy.cpitch - c), out);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Bulk-copy pipelined variants of the two BN-backward passes for dense operands (G and y are whole [N][H][W][cp]
// buffers with the same pitch: every full-resolution layer). The register-prefetch kernels above keep only ~48 KB per
// SM in flight and spill (ncu: 2.9 TB/s, 50 % "no eligible warp"); here one thread streams contiguous row chunks of G
// and y into a 3-stage shared-memory ring with cp.async.bulk (mbarrier complete_tx), so ~130 KB per SM are in
// flight without holding registers, and the 256 threads read 16-byte groups conflict-free out of shared memory.
//   MODE 0: per-block partials of sum dz, sum dz*y      MODE 1: dy = scale*dz + A*y + B  (see bn_bwd_apply_kernel)
// A chunk is `rows_per_chunk` whole image rows, or one of `segs` segments of a row when a row exceeds the stage.
// ------------------------------------------------------------------------------------------------
constexpr int kBulkStages = 3;
constexpr int kBulkStageCap = 16 * 1024;   // bytes per tensor per stage

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct BulkArgs {
  const bf16* G; const bf16* y;   // dense [rows][W][cp]; fold mode: G = padded-domain gradient [N][H+2][W+2][cp]
  int fold;                       // 1: the upstream gradient is fold_reflect(G) (adjoint of the reflect halo), formed on the fly
  int rev;                        // 1: walk the chunks back to front (pass 2 then starts on the data pass 1 left in L2)
  int capG;                       // bytes of the G part of a stage
  int cp, C, W, H, rows;          // rows = N*H
  int rows_per_chunk, segs, seg_w, n_chunks;
  const float *scale, *shift, *mean, *invstd, *drop, *s1s2;
  float inv_count; int training;
  float* part;                    // MODE 0: [gridDim.x][2][C]
  ActView dy;                     // MODE 1
};

template <int MODE>
__global__ void __launch_bounds__(256, 2)
bn_bwd_bulk_kernel(const BulkArgs a) {
  extern __shared__ __align__(128) uint8_t bsm[];
  uint8_t* sY = bsm;
  uint8_t* sG = bsm + kBulkStages * kBulkStageCap;
  uint64_t* full = reinterpret_cast<uint64_t*>(sG + (size_t)kBulkStages * a.capG);
  float* red = reinterpret_cast<float*>(bsm);   // MODE 0 epilogue: [pix_lanes][groups*16], aliases the (drained) ring
  const int C = a.C, cp = a.cp, W = a.W;
  const int groups = cp >> 3;                 // whole pixels are streamed: pad channels are zero in G (dz = 0)
  const int pix_lanes = blockDim.x / groups;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups;
  const int c = g * 8, nv = max(0, min(8, C - c));
  const bool lane_on = pl < pix_lanes;
  const size_t row_elems = (size_t)W * cp;

  const uint32_t prow_bytes = (uint32_t)((W + 2) * cp * 2);   // fold mode: one padded-domain row
  auto issue = [&](int chunk, int stage) {   // one thread
    if (a.fold) {
      // chunk = image row (n, h): padded row h+1 (W+2 pixels) and, for h == 1 / h == H-2, the halo row 0 / H+1 that folds onto it
      const int n = chunk / a.H, h = chunk - n * a.H;
      const bf16* img = a.G + (size_t)n * (a.H + 2) * (W + 2) * cp;
      const int extra = (h == 1) ? 0 : ((h == a.H - 2) ? a.H + 1 : -1);
      const uint32_t ybytes = (uint32_t)(row_elems * 2);
      mbar_arrive_expect_tx(&full[stage], prow_bytes * (extra >= 0 ? 2u : 1u) + ybytes);
      bulk_load(sG + (size_t)stage * a.capG, img + (size_t)(h + 1) * (W + 2) * cp, prow_bytes, &full[stage]);
      if (extra >= 0) bulk_load(sG + (size_t)stage * a.capG + prow_bytes, img + (size_t)extra * (W + 2) * cp, prow_bytes, &full[stage]);
      bulk_load(sY + (size_t)stage * kBulkStageCap, a.y + (size_t)chunk * row_elems, ybytes, &full[stage]);
      return;
    }
    int row0, w0, npx_rows, npx_w;
    if (a.segs > 1) { row0 = chunk / a.segs; const int sg = chunk - row0 * a.segs; w0 = sg * a.seg_w; npx_w = min(a.seg_w, W - w0); npx_rows = 1; }
    else { row0 = chunk * a.rows_per_chunk; w0 = 0; npx_w = W; npx_rows = min(a.rows_per_chunk, a.rows - row0); }
    const uint32_t bytes = (uint32_t)((size_t)npx_rows * npx_w * cp * 2);
    const size_t off = (size_t)row0 * row_elems + (size_t)w0 * cp;
    mbar_arrive_expect_tx(&full[stage], 2 * bytes);
    bulk_load(sG + (size_t)stage * a.capG, a.G + off, bytes, &full[stage]);
    bulk_load(sY + (size_t)stage * kBulkStageCap, a.y + off, bytes, &full[stage]);
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < kBulkStages; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const int my_n = (a.n_chunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto chunk_of = [&](int i) { const int k = blockIdx.x + i * gridDim.x; return a.rev ? a.n_chunks - 1 - k : k; };
  if (threadIdx.x == 0)
    for (int i = 0; i < kBulkStages && i < my_n; ++i) issue(chunk_of(i), i);

  float sc[8], sf[8], ca[8], cb[8], a1[8], a2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = (lane_on && k < nv) ? a.scale[c + k] : 0.f;
    sf[k] = (lane_on && k < nv) ? a.shift[c + k] : 0.f;
    ca[k] = 0.f; cb[k] = 0.f; a1[k] = 0.f; a2[k] = 0.f;
    if (MODE == 1 && a.training && lane_on && k < nv) {
      const float is = a.invstd[c + k], s1 = a.s1s2[c + k], s2 = a.s1s2[C + c + k];
      ca[k] = -sc[k] * s2 * is * a.inv_count;
      cb[k] = -sc[k] * (s1 - s2 * is * a.mean[c + k]) * a.inv_count;
    }
  }

  int stage = 0; uint32_t phase = 0;
  for (int i = 0; i < my_n; ++i) {
    const int chunk = chunk_of(i);
    int row0, w0, npx_rows, npx_w;
    if (a.fold) { row0 = chunk; w0 = 0; npx_w = W; npx_rows = 1; }
    else if (a.segs > 1) { row0 = chunk / a.segs; const int sg = chunk - row0 * a.segs; w0 = sg * a.seg_w; npx_w = min(a.seg_w, W - w0); npx_rows = 1; }
    else { row0 = chunk * a.rows_per_chunk; w0 = 0; npx_w = W; npx_rows = min(a.rows_per_chunk, a.rows - row0); }
    mbar_wait(&full[stage], phase);
    if (lane_on) {
      const uint4* gs = reinterpret_cast<const uint4*>(sG + (size_t)stage * a.capG);
      const uint4* ys = reinterpret_cast<const uint4*>(sY + (size_t)stage * kBulkStageCap);
      for (int rr = 0; rr < npx_rows; ++rr) {
        const int row = row0 + rr;
        const int n = row / a.H, h = row - n * a.H;
        float dr[8];
        if (a.drop) {
#pragma unroll
          for (int k = 0; k < 8; ++k) dr[k] = k < nv ? a.drop[(size_t)n * C + c + k] : 0.f;
        }
        bf16* dp = nullptr;
        if (MODE == 1) dp = a.dy.base + a.dy.pix(n, h, w0) + c;
        const int base = rr * npx_w * groups + g;
        const bool has_extra = a.fold && ((h == 1) || (h == a.H - 2));
        for (int w = pl; w < npx_w; w += pix_lanes) {
          float gv[8], yv[8], out[8];
          if (a.fold) {
            // G(h, w) = dpad(h+1, w+1) + the halo cells that mirror onto (h, w): columns 0 / W+1 for w == 1 / W-2, and the
            // whole extra row (with its own two corner columns) for h == 1 / H-2. Interior pixels take the first load only.
            unpack8(gs[(w + 1) * groups + g], gv);
            const bool edge_w = (w == 1) || (w == W - 2);
            if (has_extra || edge_w) {
              const int wm = (w == 1) ? 0 : ((w == W - 2) ? W + 1 : -1);
              const int wm2 = (W == 3 && w == 1) ? W + 1 : -1;
              const int nrow = has_extra ? 2 : 1;
              for (int r = 0; r < nrow; ++r) {
                const uint4* rowp = gs + (size_t)r * (W + 2) * groups + g;
                float t[8];
                if (r == 1) {
                  unpack8(rowp[(w + 1) * groups], t);
#pragma unroll
                  for (int k = 0; k < 8; ++k) gv[k] += t[k];
                }
                if (wm >= 0) {
                  unpack8(rowp[wm * groups], t);
#pragma unroll
                  for (int k = 0; k < 8; ++k) gv[k] += t[k];
                }
                if (wm2 >= 0) {
                  unpack8(rowp[wm2 * groups], t);
#pragma unroll
                  for (int k = 0; k < 8; ++k) gv[k] += t[k];
                }
              }
            }
          } else {
            unpack8(gs[base + w * groups], gv);
          }
          unpack8(ys[base + w * groups], yv);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float z = fmaf(yv[k], sc[k], sf[k]);
            float gg = gv[k];
            if (a.drop) gg *= dr[k];
            const float dz = (z > 0.f) ? gg : 0.f;
            if (MODE == 0) { a1[k] += dz; a2[k] = fmaf(dz, yv[k], a2[k]); }
            else out[k] = fmaf(sc[k], dz, fmaf(ca[k], yv[k], cb[k]));
          }
          if (MODE == 1) *reinterpret_cast<uint4*>(dp + (size_t)w * a.dy.cpitch) = pack8(out);
        }
      }
    }
    __syncthreads();   // every thread is done with this stage: refill it
    if (threadIdx.x == 0 && i + kBulkStages < my_n) issue(chunk_of(i + kBulkStages), stage);
    if (++stage == kBulkStages) { stage = 0; phase ^= 1; }
  }
  if (MODE == 0) {
    if (lane_on) {
      float* mine = red + (size_t)pl * groups * 16 + g * 16;
#pragma unroll
      for (int k = 0; k < 8; ++k) { mine[k] = a1[k]; mine[8 + k] = a2[k]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {   // fixed-order combine over the pixel lanes
      const int which = i / C, ch = i - which * C;
      const float* src = red + (ch >> 3) * 16 + which * 8 + (ch & 7);
      float acc = 0.f;
      for (int p = 0; p < pix_lanes; ++p) acc += src[(size_t)p * groups * 16];
      a.part[(size_t)blockIdx.x * 2 * C + i] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Element-wise dropout of nn.Dropout (reference model.py:239 center_dropout on the core centre, model.py:294 final_dropouts
// in front of the heads): a[n,h,w,c] *= keep[n,h,w,c] * scale, in place on the interior of the view. keep is a dense bf16
// 0/1 tensor [N][H][W][mask_cp] (mask_cp multiple of 8, >= C); the 1/(1-p) scale is applied in fp32. The same kernel
// multiplies the upstream gradient in backward.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mask_mul_kernel(ActView a, const bf16* __restrict__ keep, int mask_cp, float scale) {
  const int groups = (a.C + 7) >> 3;
  const long long total = (long long)a.N * a.H * a.W * groups;
  const bool vec = ((reinterpret_cast<uintptr_t>(a.base) & 15) == 0) && (a.cpitch & 7) == 0 && (a.c_off & 7) == 0 &&
                   ((reinterpret_cast<uintptr_t>(keep) & 15) == 0);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long r = i / groups;
    const int w = (int)(r % a.W); r /= a.W;
    const int h = (int)(r % a.H);
    const int n = (int)(r / a.H);
    const int c = g * 8, nv = min(8, a.C - c);
    bf16* px = a.base + a.pix(n, h, w) + c;
    const bf16* mk = keep + (((long long)n * a.H + h) * a.W + w) * mask_cp + c;
    float v[8], m[8];
    if (vec && nv == 8) {
      unpack8(*reinterpret_cast<const uint4*>(px), v);
      unpack8(*reinterpret_cast<const uint4*>(mk), m);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = v[k] * m[k] * scale;
      *reinterpret_cast<uint4*>(px) = pack8(v);
    } else {
      for (int k = 0; k < nv; ++k) px[k] = __float2bfloat16_rn(__bfloat162float(px[k]) * __bfloat162float(mk[k]) * scale);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Component-level up-sampling alternatives of `Up` (components.py:86-98). The reference cannot run them at
// whole-model level, so they are plain CUDA-core kernels, not tensor-core paths.
// ------------------------------------------------------------------------------------------------
// nn.MaxUnpool2d(2): scatter by the pooling indices, expressed as a gather over the output
__global__ void maxunpool_kernel(ActView in, const long long* __restrict__ idx_nchw, ActView o, int off_h, int off_w) {
  const int groups = (o.C + 7) >> 3;
  const long long total = (long long)o.N * o.H * o.W * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long r = i / groups;
    const int w = (int)(r % o.W); r /= o.W;
    const int h = (int)(r % o.H);
    const int n = (int)(r / o.H);
    const int c = g * 8, nv = min(8, o.C - c);
    float v[8], out[8];
    const int y = h - off_h, x = w - off_w;  // position inside the exact-2x unpooled map (rest is F.pad zeros)
    const int ph = y >> 1, pw = x >> 1;
    const bool inside = y >= 0 && x >= 0 && ph < in.H && pw < in.W;
    if (inside) load8(in.base + in.pix(n, ph, pw) + c, nv, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      out[k] = 0.f;
      if (inside && k < nv) {
        const long long id = idx_nchw[(((size_t)n * o.C + c + k) * in.H + ph) * in.W + pw];
        if (id == (long long)y * (2 * in.W) + x) out[k] = v[k];
      }
    }
    store_with_halo(o, n, h, w, c, nv, out);
  }
}

// nn.ConvTranspose2d(cin, cout, kernel_size=2, stride=2): every input pixel writes a disjoint 2x2 block
__global__ void convtranspose2x2_kernel(ActView in, const float* __restrict__ wt /*[cin][cout][2][2]*/, const float* __restrict__ bias,
                                        ActView o, int off_h, int off_w) {
  const int groups = (o.C + 7) >> 3;
  const long long total = (long long)o.N * o.H * o.W * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long r = i / groups;
    const int w = (int)(r % o.W); r /= o.W;
    const int h = (int)(r % o.H);
    const int n = (int)(r / o.H);
    const int c = g * 8, nv = min(8, o.C - c);
    const int y = h - off_h, x = w - off_w;
    const bool inside = y >= 0 && x >= 0 && (y >> 1) < in.H && (x >> 1) < in.W;
    const int tap = (y & 1) * 2 + (x & 1);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = (inside && k < nv && bias) ? bias[c + k] : 0.f;
    const bf16* src = in.base + in.pix(n, inside ? (y >> 1) : 0, inside ? (x >> 1) : 0);
    for (int ci = 0; inside && ci < in.C; ++ci) {
      const float x = __bfloat162float(src[ci]);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < nv) acc[k] = fmaf(x, wt[((size_t)ci * o.C + c + k) * 4 + tap], acc[k]);
    }
    store_with_halo(o, n, h, w, c, nv, acc);
  }
}

// NHWC bf16 view -> fp32 NCHW tensor (component-level API boundary)
__global__ void unpack_nchw_kernel(ActView in, float* __restrict__ out) {
  const long long total = (long long)in.N * in.C * in.H * in.W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % in.W);
    long long r = i / in.W;
    const int h = (int)(r % in.H); r /= in.H;
    const int c = (int)(r % in.C);
    const int n = (int)(r / in.C);
    out[i] = __bfloat162float(in.base[in.pix(n, h, w) + c]);
  }
}

}  // namespace

// ===================================== launchers =====================================
// true when every 8-channel group of the view starts 16-byte aligned and may be over-READ up to 8 channels
static bool view_vec_ok(const ActView& v) {
  return ((uintptr_t)v.base % 16) == 0 && v.cpitch % 8 == 0 && v.c_off % 8 == 0 && v.c_off + round_up(v.C, 8) <= v.cpitch;
}
// true when full 8-channel groups of the view can be written with 16-byte stores (partial groups stay scalar)
static bool view_store_vec_ok(const ActView& v) {
  return ((uintptr_t)v.base % 16) == 0 && v.cpitch % 8 == 0 && v.c_off % 8 == 0;
}

int pack_input_launch(const float* x, long long sb, long long sc, const long long* gather, const ActView& o, cudaStream_t st) {
  MIMO_CHECK(o.H >= 2 && o.W >= 2, MIMO_ERR_ARG, "pack_input: H,W must be >= 2");
  const long long total = (long long)o.N * (o.H + 2 * o.pad) * (o.W + 2 * o.pad);
  pack_input_kernel<<<grid_for(total), kBlock, 0, st>>>(x, sb, sc, gather, o);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int weight_pack_launch(const float* w, int cout, int cin, bf16* wf, int cin_pitch, bf16* wd, int cout_pitch, cudaStream_t st) {
  const long long total = 9LL * cout * cin_pitch + (wd ? 9LL * cin * cout_pitch : 0);
  weight_pack_kernel<<<grid_for(total), kBlock, 0, st>>>(w, cout, cin, wf, cin_pitch, wd, cout_pitch);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int weight_pack_batched_launch(const WeightPackJob* jobs, int n, cudaStream_t st) {
  for (int base = 0; base < n; base += kMaxPackJobs) {
    PackBatch b{};
    b.n = n - base < kMaxPackJobs ? n - base : kMaxPackJobs;
    int blocks = 0;
    for (int j = 0; j < b.n; ++j) {
      const WeightPackJob& q = jobs[base + j];
      b.job[j] = q;
      b.first_block[j] = blocks;
      const long long pairs = (long long)q.cout * q.cin_pitch + (q.wd ? (long long)q.sl.phys_count(q.cin) * q.cout_pitch : 0);
      int nb = (int)ceil_div_ll(pairs, kBlock);
      if (nb > 2 * num_sms()) nb = 2 * num_sms();
      blocks += nb < 1 ? 1 : nb;
    }
    b.first_block[b.n] = blocks;
    weight_pack_batched_kernel<<<blocks, kBlock, 0, st>>>(b);
    MIMO_LAUNCH_CHECK();
  }
  return MIMO_OK;
}

int bn_finalize_launch(const float* psum, const float* psq, int tiles, int cpitch, int C, double count, const float* gamma,
                       const float* beta, const float* conv_bias, float* rm, float* rv, long long* nbt, float momentum, float eps,
                       float* scale, float* shift, float* save_mean, float* save_invstd, cudaStream_t st) {
  static const int skip = getenv("MIMO_DEBUG_SKIP_FINALIZE") ? atoi(getenv("MIMO_DEBUG_SKIP_FINALIZE")) : 0;   // timing experiments only
  if (skip) return MIMO_OK;
  bn_finalize_kernel<<<C, 128, 0, st>>>(psum, psq, tiles, cpitch, C, count, gamma, beta, conv_bias, rm, rv, nbt, momentum, eps, scale,
                                        shift, save_mean, save_invstd);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int bn_eval_affine_launch(int C, const float* gamma, const float* beta, const float* conv_bias, const float* rm, const float* rv,
                          float eps, float* scale, float* shift, float* save_mean, float* save_invstd, cudaStream_t st) {
  bn_eval_affine_kernel<<<ceil_div(C, 128), 128, 0, st>>>(C, gamma, beta, conv_bias, rm, rv, eps, scale, shift, save_mean, save_invstd);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int bn_relu_apply_launch(const bf16* y, int ycp, const float* scale, const float* shift, const float* drop, const ActView& o,
                         const ActView* pool, cudaStream_t st) {
  MIMO_CHECK(o.H >= 2 && o.W >= 2, MIMO_ERR_ARG, "bn_relu_apply: H,W must be >= 2");
  ActView pv = pool ? *pool : o;
  if (pool) MIMO_CHECK(pool->H == o.H / 2 && pool->W == o.W / 2 && pool->C == o.C && pool->N == o.N, MIMO_ERR_ARG, "bn_relu_apply: pool view shape mismatch");
  MIMO_CHECK(ycp % 8 == 0 && ycp >= round_up(o.C, 8) && ((uintptr_t)y % 16) == 0, MIMO_ERR_ALIGN, "bn_relu_apply: y must be 16-byte aligned with a pitch >= round_up(C, 8)");
  MIMO_CHECK((o.C + 7) / 8 <= kBlock, MIMO_ERR_ARG, "bn_relu_apply: too many channels (%d)", o.C);
  const int cell_rows = o.N * ((o.H + 1) / 2);
  const int grid = cell_rows < 8 * num_sms() ? cell_rows : 8 * num_sms();
  const bool vo = view_store_vec_ok(o), vp = view_store_vec_ok(pv);
  if (vo && vp) bn_relu_apply_kernel<true, true><<<grid, kBlock, 0, st>>>(y, ycp, scale, shift, drop, o, pv, pool ? 1 : 0, elementwise_reverse());
  else if (vo) bn_relu_apply_kernel<true, false><<<grid, kBlock, 0, st>>>(y, ycp, scale, shift, drop, o, pv, pool ? 1 : 0, elementwise_reverse());
  else if (vp) bn_relu_apply_kernel<false, true><<<grid, kBlock, 0, st>>>(y, ycp, scale, shift, drop, o, pv, pool ? 1 : 0, elementwise_reverse());
  else bn_relu_apply_kernel<false, false><<<grid, kBlock, 0, st>>>(y, ycp, scale, shift, drop, o, pv, pool ? 1 : 0, elementwise_reverse());
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int maxpool_launch(const ActView& in, const ActView& o, long long* idx_nchw, cudaStream_t st) {
  MIMO_CHECK(o.H == in.H / 2 && o.W == in.W / 2 && o.C == in.C && o.N == in.N, MIMO_ERR_ARG, "maxpool: shape mismatch");
  const long long total = (long long)o.N * o.H * o.W * ((o.C + 7) / 8);
  maxpool_kernel<<<grid_for(total), kBlock, 0, st>>>(in, o, idx_nchw);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int halo_fill_launch(const ActView& v, cudaStream_t st) {
  MIMO_CHECK(v.pad == 1, MIMO_ERR_ARG, "halo_fill: needs a pad == 1 view");
  const long long total = (long long)v.N * (2 * (v.W + 2) + 2 * v.H) * ((v.C + 7) / 8);
  halo_fill_kernel<<<grid_for(total), kBlock, 0, st>>>(v);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int upsample_launch(const ActView& in, const ActView& o, cudaStream_t st) {
  MIMO_CHECK(o.N == in.N && o.C == in.C, MIMO_ERR_ARG, "upsample: N/C mismatch");
  const int dY = o.H - 2 * in.H, dX = o.W - 2 * in.W;
  MIMO_CHECK(dY >= 0 && dX >= 0, MIMO_ERR_ARG, "upsample: skip smaller than the up-sampled map");
  const int groups = (o.C + 7) / 8;
  if (groups <= 32 && view_vec_ok(in) && ((uintptr_t)o.base % 16) == 0 && o.cpitch % 8 == 0 && o.pad <= 1) {
    const int rows = o.N * o.H;
    const int grid = rows < 8 * num_sms() ? rows : 8 * num_sms();
    int gpp = 1;
    while (gpp < groups) gpp <<= 1;
    switch (gpp) {
      case 1: upsample_fast_kernel<1><<<grid, kBlock, 0, st>>>(in, o, dY / 2, dX / 2); break;
      case 2: upsample_fast_kernel<2><<<grid, kBlock, 0, st>>>(in, o, dY / 2, dX / 2); break;
      case 4: upsample_fast_kernel<4><<<grid, kBlock, 0, st>>>(in, o, dY / 2, dX / 2); break;
      case 8: upsample_fast_kernel<8><<<grid, kBlock, 0, st>>>(in, o, dY / 2, dX / 2); break;
      case 16: upsample_fast_kernel<16><<<grid, kBlock, 0, st>>>(in, o, dY / 2, dX / 2); break;
      default: upsample_fast_kernel<32><<<grid, kBlock, 0, st>>>(in, o, dY / 2, dX / 2); break;
    }
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
  }
  const long long total = (long long)o.N * o.H * o.W * groups;
  upsample_kernel<<<grid_for(total), kBlock, 0, st>>>(in, o, dY / 2, dX / 2);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

// skip: dense (pad 0) NHWC copy of the skip tensor; o: the WHOLE haloed concat buffer (c_off 0, C = skip.C + in.C)
bool upsample_concat_ok(const ActView& in, const ActView& skip, const ActView& o) {
  if (o.pad != 1 || o.c_off != 0 || o.C != skip.C + in.C || o.C > o.cpitch || (o.cpitch & 7) || (o.cpitch >> 3) > 32) return false;
  if (skip.pad != 0 || skip.N != o.N || skip.H != o.H || skip.W != o.W || (skip.c_off & 7) || (skip.cpitch & 7)) return false;
  if (skip.c_off + ((skip.C + 7) & ~7) > skip.cpitch) return false;
  if (in.N != o.N || (in.c_off & 7) || (in.cpitch & 7) || in.c_off + ((in.C + 7) & ~7) > in.cpitch) return false;
  if (((uintptr_t)o.base | (uintptr_t)skip.base | (uintptr_t)in.base) & 15) return false;
  return o.H >= 2 * in.H && o.W >= 2 * in.W;
}

int upsample_concat_launch(const ActView& in, const ActView& skip, const ActView& o, cudaStream_t st) {
  MIMO_CHECK(upsample_concat_ok(in, skip, o), MIMO_ERR_ARG, "upsample_concat: unsupported views");
  const int dY = o.H - 2 * in.H, dX = o.W - 2 * in.W;
  const int words = o.cpitch >> 3;
  const int rows = o.N * o.H;
  const int grid = rows < 8 * num_sms() ? rows : 8 * num_sms();
  int wpp = 1;
  while (wpp < words) wpp <<= 1;
  const ActView inv = [&] { ActView v = in; v.base = in.base + in.c_off; v.c_off = 0; return v; }();
  const ActView sv = [&] { ActView v = skip; v.base = skip.base + skip.c_off; v.c_off = 0; return v; }();
  switch (wpp) {
    case 1: upsample_concat_kernel<1><<<grid, kBlock, 0, st>>>(inv, sv, o, dY / 2, dX / 2); break;
    case 2: upsample_concat_kernel<2><<<grid, kBlock, 0, st>>>(inv, sv, o, dY / 2, dX / 2); break;
    case 4: upsample_concat_kernel<4><<<grid, kBlock, 0, st>>>(inv, sv, o, dY / 2, dX / 2); break;
    case 8: upsample_concat_kernel<8><<<grid, kBlock, 0, st>>>(inv, sv, o, dY / 2, dX / 2); break;
    case 16: upsample_concat_kernel<16><<<grid, kBlock, 0, st>>>(inv, sv, o, dY / 2, dX / 2); break;
    default: upsample_concat_kernel<32><<<grid, kBlock, 0, st>>>(inv, sv, o, dY / 2, dX / 2); break;
  }
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int upsample_bwd_launch(const ActView& gdst, const ActView& gsrc, int accumulate, cudaStream_t st) {
  MIMO_CHECK(gdst.N == gsrc.N && gdst.C == gsrc.C && gdst.pad == 0 && gsrc.pad == 0, MIMO_ERR_ARG, "upsample_bwd: view mismatch");
  const int dY = gdst.H - 2 * gsrc.H, dX = gdst.W - 2 * gsrc.W;
  MIMO_CHECK(dY >= 0 && dX >= 0, MIMO_ERR_ARG, "upsample_bwd: bad sizes");
  const long long total = (long long)gsrc.N * gsrc.H * gsrc.W * ((gsrc.C + 7) / 8);
  if (gsrc.H >= 4 && gsrc.W >= 4 && total < (1ll << 31)) {   // candidate window <= 8 destination pixels per direction
    if (view_vec_ok(gdst)) upsample_bwd_fast_kernel<true><<<grid_for(total), kBlock, 0, st>>>(gdst, gsrc, dY / 2, dX / 2, accumulate);
    else upsample_bwd_fast_kernel<false><<<grid_for(total), kBlock, 0, st>>>(gdst, gsrc, dY / 2, dX / 2, accumulate);
    MIMO_LAUNCH_CHECK();
    return MIMO_OK;
  }
  upsample_bwd_kernel<<<grid_for(total), kBlock, 0, st>>>(gdst, gsrc, dY / 2, dX / 2, accumulate);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int grad_gather_launch(const ActView* dpad, const ActView* gpool, const ActView* act, const ActView& gout, int accumulate,
                       cudaStream_t st, const ActView* dpad2) {
  MIMO_CHECK(gout.pad == 0, MIMO_ERR_ARG, "grad_gather: output must be unpadded");
  if (dpad) MIMO_CHECK(dpad->pad == 0 && dpad->H == gout.H + 2 && dpad->W == gout.W + 2 && dpad->C == gout.C, MIMO_ERR_ARG, "grad_gather: dpad shape mismatch");
  if (gpool) MIMO_CHECK(act && gpool->pad == 0 && gpool->H == gout.H / 2 && gpool->W == gout.W / 2 && gpool->C == gout.C && act->C == gout.C && act->H == gout.H, MIMO_ERR_ARG, "grad_gather: pool shape mismatch");
  MIMO_CHECK((gout.C + 7) / 8 <= kBlock, MIMO_ERR_ARG, "grad_gather: too many channels (%d)", gout.C);
  const bool vo = view_vec_ok(gout);
  if (dpad2) MIMO_CHECK(!gpool && dpad && dpad2->pad == 0 && dpad2->H == dpad->H && dpad2->W == dpad->W && dpad2->C == dpad->C && dpad2->N == dpad->N &&
                        dpad2->cpitch == dpad->cpitch && dpad2->c_off == dpad->c_off, MIMO_ERR_ARG, "grad_gather: second source must match the first");
  if (!gpool) {
    const bool vi = view_vec_ok(*dpad) && (!dpad2 || view_vec_ok(*dpad2));
    const int rows = gout.N * gout.H;
    const int grid = rows < 8 * num_sms() ? rows : 8 * num_sms();
    if (vi && vo) grad_fold_kernel<true, true><<<grid, kBlock, 0, st>>>(*dpad, dpad2 ? *dpad2 : *dpad, dpad2 ? 1 : 0, gout, accumulate);
    else if (vi) grad_fold_kernel<true, false><<<grid, kBlock, 0, st>>>(*dpad, dpad2 ? *dpad2 : *dpad, dpad2 ? 1 : 0, gout, accumulate);
    else if (vo) grad_fold_kernel<false, true><<<grid, kBlock, 0, st>>>(*dpad, dpad2 ? *dpad2 : *dpad, dpad2 ? 1 : 0, gout, accumulate);
    else grad_fold_kernel<false, false><<<grid, kBlock, 0, st>>>(*dpad, dpad2 ? *dpad2 : *dpad, dpad2 ? 1 : 0, gout, accumulate);
  } else {
    const bool vec = vo && (!dpad || view_vec_ok(*dpad)) && view_vec_ok(*gpool) && view_vec_ok(*act);
    const int cell_rows = gout.N * ((gout.H + 1) / 2);
    const int grid = cell_rows < 8 * num_sms() ? cell_rows : 8 * num_sms();
    if (vec) grad_gather_pool_kernel<true, true, true, true><<<grid, kBlock, 0, st>>>(dpad ? *dpad : gout, dpad ? 1 : 0, *gpool, *act, gout, accumulate);
    else grad_gather_pool_kernel<false, false, false, false><<<grid, kBlock, 0, st>>>(dpad ? *dpad : gout, dpad ? 1 : 0, *gpool, *act, gout, accumulate);
  }
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int bn_bwd_parts(int C) { (void)C; return 4 * num_sms(); }

int bn_bwd_launch(const ActView& G, const bf16* y, int ycp, const float* scale, const float* shift, const float* mean,
                  const float* invstd, const float* drop, int training, float* part, float* s1s2, float* dgamma, float* dbeta,
                  float* dbias, float grad_scale, int accumulate, const ActView& dy, cudaStream_t st, const ActView* fold_src) {
  const int C = G.C;
  MIMO_CHECK(G.pad == 0, MIMO_ERR_ARG, "bn_bwd: G must be unpadded");
  MIMO_CHECK(dy.pad != 1 && dy.c_off == 0 && dy.N == G.N && dy.H == G.H && dy.W == G.W && dy.cpitch >= C && dy.cpitch % 8 == 0 &&
                 ((uintptr_t)dy.base % 16) == 0,
             MIMO_ERR_ARG, "bn_bwd: dy must be a whole dense or zero-tail buffer of the same shape");
  MIMO_CHECK(ycp % 8 == 0 && ycp >= round_up(C, 8) && ((uintptr_t)y % 16) == 0, MIMO_ERR_ALIGN, "bn_bwd: y must be 16-byte aligned with a pitch >= round_up(C, 8)");
  const int groups = (C + 7) / 8;
  MIMO_CHECK(groups <= kBlock, MIMO_ERR_ARG, "bn_bwd: too many channels (%d)", C);
  const int rows = G.N * G.H;
  {
    // dense operands with one pitch: bulk-copy pipelined kernels (every full-resolution layer). With `fold_src` the
    // upstream gradient is fold_reflect(*fold_src) and is formed inside the two passes (no grad_fold launch, no G buffer).
    static const int bulk_on = getenv("MIMO_BN_BULK") ? atoi(getenv("MIMO_BN_BULK")) : 1;
    const int cp = round_up(C, 8);
    const size_t row_bytes = (size_t)G.W * cp * 2;
    const size_t prow_bytes = (size_t)(G.W + 2) * cp * 2;
    const bool common = bulk_on && ycp == cp && dy.cpitch == cp && (cp >> 3) <= kBlock && (size_t)G.N * (G.H + 2) * (G.W + 2) * cp < (1ull << 31);
    bool fold = false;
    if (fold_src) {
      const ActView& d = *fold_src;
      fold = common && d.pad == 0 && d.c_off == 0 && d.cpitch == cp && d.C == C && d.N == G.N && d.H == G.H + 2 && d.W == G.W + 2 &&
             ((uintptr_t)d.base % 16) == 0 && G.H >= 4 && G.W >= 3 && row_bytes <= (size_t)kBulkStageCap && 2 * prow_bytes <= 50 * 1024;
      if (!fold) {
        int rc = grad_gather_launch(fold_src, nullptr, nullptr, G, 0, st);
        if (rc) return rc;
      }
    }
    const bool dense = common && G.c_off == 0 && G.cpitch == cp && ((uintptr_t)G.base % 16) == 0;
    if (fold || dense) {
      BulkArgs a{};
      a.fold = fold ? 1 : 0;
      a.G = fold ? fold_src->base : G.base; a.y = y; a.cp = cp; a.C = C; a.W = G.W; a.H = G.H; a.rows = rows;
      if (fold) {
        a.rows_per_chunk = 1; a.segs = 1; a.seg_w = G.W; a.n_chunks = rows;
        a.capG = round_up((int)(2 * prow_bytes), 128);
      } else if (row_bytes <= (size_t)kBulkStageCap) {
        a.rows_per_chunk = (int)(kBulkStageCap / row_bytes); a.segs = 1; a.seg_w = G.W;
        a.n_chunks = ceil_div(rows, a.rows_per_chunk);
        a.capG = kBulkStageCap;
      } else {
        a.rows_per_chunk = 1; a.segs = (int)ceil_div_ll((long long)row_bytes, kBulkStageCap);
        a.seg_w = ceil_div(G.W, a.segs); a.segs = ceil_div(G.W, a.seg_w);
        a.n_chunks = rows * a.segs;
        a.capG = kBulkStageCap;
      }
      a.scale = scale; a.shift = shift; a.mean = mean; a.invstd = invstd; a.drop = drop; a.s1s2 = s1s2;
      a.inv_count = 1.f / (float)((long long)G.N * G.H * G.W); a.training = training;
      a.part = part; a.dy = dy;
      const size_t smem = (size_t)kBulkStages * (kBulkStageCap + a.capG) + kBulkStages * 8 + 64;
      // per launch (cheap): the attribute is per DEVICE, a process-wide "done" flag would skip the other GPUs of the process
      MIMO_CUDA(cudaFuncSetAttribute(bn_bwd_bulk_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      MIMO_CUDA(cudaFuncSetAttribute(bn_bwd_bulk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      int grid = (smem <= 112 * 1024 ? 2 : 1) * num_sms();
      if (grid > a.n_chunks) grid = a.n_chunks;
      if (grid > bn_bwd_parts(C)) grid = bn_bwd_parts(C);
      bn_bwd_bulk_kernel<0><<<grid, kBlock, smem, st>>>(a);
      MIMO_LAUNCH_CHECK();
      static const int skipf = getenv("MIMO_DEBUG_SKIP_FINALIZE") ? atoi(getenv("MIMO_DEBUG_SKIP_FINALIZE")) : 0;   // timing experiments only
      if (!skipf)
      bn_bwd_finalize_kernel<<<ceil_div(C, 8), 256, 0, st>>>(part, grid, C, s1s2, dgamma, dbeta, dbias, scale, mean, invstd, training, grad_scale, accumulate);
      MIMO_LAUNCH_CHECK();
      a.rev = elementwise_reverse();
      bn_bwd_bulk_kernel<1><<<grid, kBlock, smem, st>>>(a);
      MIMO_LAUNCH_CHECK();
      return MIMO_OK;
    }
  }
  const int nparts = rows < bn_bwd_parts(C) ? rows : bn_bwd_parts(C);
  const int pix_lanes = kBlock / groups;
  const size_t sh_bytes = (size_t)pix_lanes * groups * 16 * sizeof(float);
  const bool vec = view_vec_ok(G);
  if (vec) bn_bwd_reduce_kernel<true><<<nparts, kBlock, sh_bytes, st>>>(G, y, ycp, scale, shift, drop, C, part);
  else bn_bwd_reduce_kernel<false><<<nparts, kBlock, sh_bytes, st>>>(G, y, ycp, scale, shift, drop, C, part);
  MIMO_LAUNCH_CHECK();
  bn_bwd_finalize_kernel<<<ceil_div(C, 8), 256, 0, st>>>(part, nparts, C, s1s2, dgamma, dbeta, dbias, scale, mean, invstd, training, grad_scale, accumulate);
  MIMO_LAUNCH_CHECK();
  const long long npix = (long long)G.N * G.H * G.W;
  const int grid = rows < 8 * num_sms() ? rows : 8 * num_sms();
  if (vec) bn_bwd_apply_kernel<true><<<grid, kBlock, 0, st>>>(G, y, ycp, scale, shift, mean, invstd, drop, s1s2, C, 1.f / (float)npix, training, dy);
  else bn_bwd_apply_kernel<false><<<grid, kBlock, 0, st>>>(G, y, ycp, scale, shift, mean, invstd, drop, s1s2, C, 1.f / (float)npix, training, dy);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int mask_mul_launch(const ActView& a, const bf16* keep, int mask_cp, float scale, cudaStream_t st) {
  MIMO_CHECK(keep != nullptr && mask_cp % 8 == 0 && mask_cp >= a.C, MIMO_ERR_ARG, "mask_mul: keep mask must be dense bf16 [N][H][W][cp], cp %% 8 == 0, cp >= C");
  const long long total = (long long)a.N * a.H * a.W * ((a.C + 7) / 8);
  mask_mul_kernel<<<grid_for(total), kBlock, 0, st>>>(a, keep, mask_cp, scale);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int maxunpool_launch(const ActView& in, const long long* idx_nchw, const ActView& o, cudaStream_t st) {
  MIMO_CHECK(o.N == in.N && o.C == in.C && o.H >= 2 * in.H && o.W >= 2 * in.W, MIMO_ERR_ARG, "maxunpool: shape mismatch");
  const long long total = (long long)o.N * o.H * o.W * ((o.C + 7) / 8);
  maxunpool_kernel<<<grid_for(total), kBlock, 0, st>>>(in, idx_nchw, o, (o.H - 2 * in.H) / 2, (o.W - 2 * in.W) / 2);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int convtranspose2x2_launch(const ActView& in, const float* wt, const float* bias, const ActView& o, cudaStream_t st) {
  MIMO_CHECK(o.N == in.N && o.H >= 2 * in.H && o.W >= 2 * in.W, MIMO_ERR_ARG, "convtranspose2x2: output smaller than twice the input");
  const long long total = (long long)o.N * o.H * o.W * ((o.C + 7) / 8);
  convtranspose2x2_kernel<<<grid_for(total), kBlock, 0, st>>>(in, wt, bias, o, (o.H - 2 * in.H) / 2, (o.W - 2 * in.W) / 2);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int unpack_nchw_launch(const ActView& in, float* out, cudaStream_t st) {
  const long long total = (long long)in.N * in.C * in.H * in.W;
  unpack_nchw_kernel<<<grid_for(total), kBlock, 0, st>>>(in, out);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

}  // namespace mimo
