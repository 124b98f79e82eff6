// CUDA-core 3x3 convolution for the network's FIRST layer (reference components.py:23: nn.Conv2d(in_channels, mid, 3, padding=1,
// padding_mode="reflect") of every encoder InConv, model.py:126-130): <= 4 input channels, <= 32 output channels, full resolution.
//
// Why not the tensor cores: with 3 input channels the implicit GEMM has K = 27. The tcgen05 kernels stage 64-channel TMA boxes and
// issue one K = 16 MMA per tap (3 of 16 K-lanes used, 43 fixed cycles each): 68 us per layer at C2 for 1.5 GFLOP. A direct
// convolution on the FMA pipe needs 0.85 G fp32 FMAs; the layer's HBM traffic (8 B in, 48 B out per pixel) is ~13 us.
// Arithmetic: bf16 inputs and weights, fp32 accumulation, bf16 output: the same rounding points as the tensor-core kernels.
//
//   conv3x3_thin_kernel : thread = two horizontally adjacent output pixels x all output channels; weights as fp32 in shared memory
//                         ([tap][ci][co], broadcast 16-byte reads); accumulators are channel PAIRS updated with the packed
//                         fma.rn.f32x2 (the 3-register FFMA issues every other cycle on this part: scalar version 60 us, packed 53 us);
//                         training epilogue = bf16 store + per-thread BatchNorm partial sums of the stored values, reduced once per
//                         CTA into the CTA's row of the statistics buffers (same contract as the tensor-core kernels:
//                         <= conv3x3_stat_rows() CTAs); inference epilogue = fused affine / ReLU / Dropout2d factor / haloed
//                         destination (ConvFuse).
//
// The matching weight gradient (warp = tap, lane = pixel, 72 accumulators per lane) was built and measured at 114..125 us against
// 109 us of conv3x3_wgrad_flat_kernel (one 9-warp block per SM at 167 registers: load latency is not hidden); it was removed again,
// the first layer's wgrad stays on the tensor-core kernel (profiles/r02_findings.md).
#include "common.cuh"
#include "conv_epilogue.cuh"
#include "ops.h"

#include <stdlib.h>

namespace mimo {
namespace {

constexpr int kThinThreads = 384;

struct ThinParams {
  const bf16* x;       // haloed input [N][H+2][W+2][x_cpitch], already offset to the view's first channel
  int x_cpitch;
  int N, H, W;
  const bf16* w;       // packed [9][cout][cin_pitch]
  int cin, cin_pitch, cout;
  EpiArgs epi;
};

__device__ __forceinline__ void load4(const bf16* p, float f[4]) {
  const uint2 r = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&r.x), b = *reinterpret_cast<const __nv_bfloat162*>(&r.y);
  f[0] = __low2float(a); f[1] = __high2float(a); f[2] = __low2float(b); f[3] = __high2float(b);
}

template <int CIN, int CO8, bool FUSE>
__global__ void __launch_bounds__(kThinThreads, 1)
conv3x3_thin_kernel(const ThinParams p) {
  constexpr int CO = CO8 * 8;
  __shared__ __align__(16) float wsm[9 * CIN][CO];
  __shared__ float red[kThinThreads / 32][2 * CO];
  for (int i = threadIdx.x; i < 9 * CIN * CO; i += kThinThreads) {
    const int co = i % CO, k = i / CO, ci = k % CIN, tap = k / CIN;
    wsm[k][co] = (co < p.cout && ci < p.cin) ? __bfloat162float(p.w[((size_t)tap * p.cout + co) * p.cin_pitch + ci]) : 0.f;
  }
  __syncthreads();
  const EpiArgs& e = p.epi;
  const int H = p.H, W = p.W, wb = W + 2, xcp = p.x_cpitch;
  const int pw_n = (W + 1) >> 1;
  const int total = p.N * H * pw_n;
  const bool stats = !FUSE && e.stat_sum != nullptr;
  const int out_cmax = FUSE ? e.out_cmax : e.out_cpitch;
  float ssum[CO], ssq[CO];
#pragma unroll
  for (int c = 0; c < CO; ++c) { ssum[c] = 0.f; ssq[c] = 0.f; }
  if (FUSE) {
    // the per-channel affine lives in the registers the (unused) statistics would occupy
#pragma unroll
    for (int c = 0; c < CO; ++c) {
      ssum[c] = c < e.cout ? __ldg(e.scale + c) : 0.f;
      ssq[c] = c < e.cout ? __ldg(e.bias + c) : 0.f;
    }
  }
  for (int idx = blockIdx.x * kThinThreads + threadIdx.x; idx < total; idx += gridDim.x * kThinThreads) {
    const int pw = idx % pw_n;
    const int r = idx / pw_n;
    const int h = r % H, n = r / H;
    const int w0 = pw * 2;
    const bool two = w0 + 1 < W;
    // 3 rows x 4 columns of the haloed input around the pixel pair (column w0 + 3 only exists for a full pair)
    float in[3][4][CIN];
    const bf16* xp = p.x + ((size_t)(n * (H + 2) + h) * wb + w0) * xcp;
#pragma unroll
    for (int rr = 0; rr < 3; ++rr)
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        float f[4] = {0.f, 0.f, 0.f, 0.f};
        if (cc < 3 || two) load4(xp + ((size_t)rr * wb + cc) * xcp, f);
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) in[rr][cc][ci] = ci < p.cin ? f[ci] : 0.f;   // pad channels may hold anything (0 * NaN = NaN)
      }
    // accumulators as channel PAIRS: the 3-register FFMA issues every other cycle on this part, the packed fma.rn.f32x2 does two
    // FMAs per issue (ncu on the scalar version: 43 M warp instructions, issue slots 64 % busy, 60 us)
    float2 acc[2][CO / 2];
#pragma unroll
    for (int c = 0; c < CO / 2; ++c) { acc[0][c] = make_float2(0.f, 0.f); acc[1][c] = make_float2(0.f, 0.f); }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float2 a0 = make_float2(in[kh][kw][ci], in[kh][kw][ci]), a1 = make_float2(in[kh][kw + 1][ci], in[kh][kw + 1][ci]);
          const float4* wr = reinterpret_cast<const float4*>(wsm[(kh * 3 + kw) * CIN + ci]);
#pragma unroll
          for (int c4 = 0; c4 < CO / 4; ++c4) {
            const float4 wv = wr[c4];
            const float2 w01 = make_float2(wv.x, wv.y), w23 = make_float2(wv.z, wv.w);
            acc[0][c4 * 2 + 0] = __ffma2_rn(a0, w01, acc[0][c4 * 2 + 0]); acc[1][c4 * 2 + 0] = __ffma2_rn(a1, w01, acc[1][c4 * 2 + 0]);
            acc[0][c4 * 2 + 1] = __ffma2_rn(a0, w23, acc[0][c4 * 2 + 1]); acc[1][c4 * 2 + 1] = __ffma2_rn(a1, w23, acc[1][c4 * 2 + 1]);
          }
        }
    const float* drow = (FUSE && e.drop != nullptr) ? e.drop + (size_t)n * e.cout : nullptr;
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      if (px == 1 && !two) break;
      const int w = w0 + px;
      const size_t pix = (FUSE && e.halo) ? ((size_t)n * (H + 2) + h + 1) * wb + w + 1 : ((size_t)n * H + h) * W + w;
      bf16* dst = e.out + pix * e.out_cpitch;
#pragma unroll
      for (int c8 = 0; c8 < CO8; ++c8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = c8 * 8 + j;
          float t = (c & 1) ? acc[px][c >> 1].y : acc[px][c >> 1].x;
          if (FUSE) {
            t = fmaf(t, ssum[c], ssq[c]);
            if (e.relu & 1) t = fmaxf(t, 0.f);
            if (drow != nullptr) t *= (c < e.cout) ? __ldg(drow + c) : 0.f;
          } else {
            if (e.bias != nullptr) t += (c < e.cout) ? __ldg(e.bias + c) : 0.f;
            if (e.relu & 1) t = fmaxf(t, 0.f);
          }
          v[j] = t;
        }
        const uint4 pk = pack8(v);
        if (c8 * 8 < out_cmax) *reinterpret_cast<uint4*>(dst + c8 * 8) = pk;
        if (stats) {
          float f[8];
          unpack8(pk, f);   // statistics of the values as stored
#pragma unroll
          for (int j = 0; j < 8; ++j) { ssum[c8 * 8 + j] += f[j]; ssq[c8 * 8 + j] = fmaf(f[j], f[j], ssq[c8 * 8 + j]); }
        }
      }
    }
  }
  if (stats) {
    // one partial row per CTA: lanes, then warps (fixed order)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < CO; ++c) {
      const float s = warp_sum(ssum[c]), q = warp_sum(ssq[c]);
      if (lane == 0) { red[warp][c] = s; red[warp][CO + c] = q; }
    }
    __syncthreads();
    for (int col = threadIdx.x; col < e.out_cpitch; col += kThinThreads) {
      float s = 0.f, q = 0.f;
      if (col < CO)
        for (int wgi = 0; wgi < kThinThreads / 32; ++wgi) { s += red[wgi][col]; q += red[wgi][CO + col]; }
      e.stat_sum[(size_t)blockIdx.x * e.out_cpitch + col] = s;
      e.stat_sq[(size_t)blockIdx.x * e.out_cpitch + col] = q;
      for (int rr = blockIdx.x + gridDim.x; rr < e.stat_rows; rr += gridDim.x) {
        e.stat_sum[(size_t)rr * e.out_cpitch + col] = 0.f;
        e.stat_sq[(size_t)rr * e.out_cpitch + col] = 0.f;
      }
    }
  }
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

template <int CIN, int CO8>
int launch_thin(const ThinParams& p, bool fuse, int grid, cudaStream_t stream) {
  if (fuse) conv3x3_thin_kernel<CIN, CO8, true><<<grid, kThinThreads, 0, stream>>>(p);
  else conv3x3_thin_kernel<CIN, CO8, false><<<grid, kThinThreads, 0, stream>>>(p);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

template <int CIN>
int launch_thin_co(const ThinParams& p, int co8, bool fuse, int grid, cudaStream_t stream) {
  switch (co8) {
    case 1: return launch_thin<CIN, 1>(p, fuse, grid, stream);
    case 2: return launch_thin<CIN, 2>(p, fuse, grid, stream);
    case 3: return launch_thin<CIN, 3>(p, fuse, grid, stream);
    default: return launch_thin<CIN, 4>(p, fuse, grid, stream);
  }
}

}  // namespace

// fprop of a haloed input with at most 4 channels into at most 32 output channels
bool conv3x3_thin_ok(const ActView& in, int mode, int cout) {
  static const int enabled = env_int("MIMO_CONV_THIN", 1);
  if (!enabled || mode != 0 || in.pad != 1) return false;
  if (in.C > 4 || cout > 32) return false;
  if ((in.cpitch | in.c_off) & 3) return false;   // 8-byte loads of four channels
  return (long long)in.N * (in.H + 2) * (in.W + 2) * in.cpitch < (1ll << 31);
}

int conv3x3_thin_launch(const ActView& in, const bf16* wpacked, int cout, int cin_pitch, bf16* out, int out_cpitch, float* stat_sum,
                        float* stat_sq, const float* bias, int relu, cudaStream_t stream, const ConvFuse* fuse) {
  note_kernel(11);
  const int co8 = ceil_div(cout, 8);
  MIMO_CHECK(out_cpitch >= co8 * 8, MIMO_ERR_ARG, "conv3x3_thin: out_cpitch %d < round_up(cout = %d, 8)", out_cpitch, cout);
  MIMO_CHECK(fuse == nullptr || stat_sum == nullptr, MIMO_ERR_ARG, "conv3x3_thin: the fused inference epilogue takes no statistics");
  ThinParams p{};
  p.x = in.base + in.c_off;
  p.x_cpitch = in.cpitch;
  p.N = in.N; p.H = in.H; p.W = in.W;
  p.w = wpacked; p.cin = in.C; p.cin_pitch = cin_pitch; p.cout = cout;
  p.epi.cout = cout;
  p.epi.out_cpitch = out_cpitch;
  p.epi.stat_rows = conv3x3_stat_rows();
  p.epi.out = out;
  p.epi.stat_sum = stat_sum;
  p.epi.stat_sq = stat_sq;
  p.epi.bias = fuse ? fuse->shift : bias;
  p.epi.relu = fuse ? (fuse->relu ? 1 : 0) : relu;
  p.epi.scale = fuse ? fuse->scale : nullptr;
  p.epi.drop = fuse ? fuse->drop : nullptr;
  p.epi.halo = fuse ? fuse->halo : 0;
  p.epi.out_cmax = co8 * 8;
  const long long work = (long long)in.N * in.H * ((in.W + 1) / 2);
  int grid = (int)ceil_div_ll(work, kThinThreads);
  if (grid > conv3x3_stat_rows()) grid = conv3x3_stat_rows();   // one statistics row per CTA
  if (in.C <= 3) return launch_thin_co<3>(p, co8, fuse != nullptr, grid, stream);
  return launch_thin_co<4>(p, co8, fuse != nullptr, grid, stream);
}

}  // namespace mimo
