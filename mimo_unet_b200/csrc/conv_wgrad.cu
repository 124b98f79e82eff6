// Weight gradient of the reflect-padded 3x3 convolution on tcgen05 tensor cores.
//
// Replaces cuDNN wgrad in the reference's autograd graph for nn.Conv2d (components.py:23,26).
//
//   dW[kh][kw][co][ci] = sum_{n,h,w} dY[n,h,w,co] * Xpad[n,h+kh,w+kw,ci]
//
// GEMM view per (kh, kw): D[co, ci] = A^T B with the reduction (K) dimension running over PIXELS.
// Both operands sit in NHWC memory, i.e. the M/N dimension (channels) is the contiguous one: they are fed to
// the tensor core as MN-major operands (SWIZZLE_128B), loaded by the very same 4-D TMA boxes the forward
// kernel uses (64 channels x a tw*th*tn = 64 pixel tile).
//
// One CTA work item = (128-co tile, 64-ci chunk, kernel row kh, pixel range): 3 accumulators (kw = 0..2) of
// 128 x 64 fp32 in TMEM. Pixels are split over CTAs (split-K); partial results are combined with vectorised
// fp32 reductions (red.global.add.v4.f32) into a packed fp32 buffer [9][cout][cin_pitch] that a small kernel
// transposes into PyTorch's OIHW .grad layout.
#include "common.cuh"
#include "ops.h"

namespace mimo {
namespace {

constexpr int kPix = 64;                     // pixels per k-block
constexpr int kTileBytes = kPix * 64 * 2;    // one (64 px x 64 ch) box = 8 KB
constexpr int kStageBytes = 5 * kTileBytes;  // 2 A boxes (co 0..63, 64..127) + 3 B boxes (kw = 0..2)
constexpr int kStages = 5;
constexpr int kThreads = 192;

struct WgradParams {
  int n_img, H, W;
  int tw, th, tn, tiles_w, tiles_h, tiles_n, p_tiles;
  int co_tiles, ci_chunks, ksplit;
  int cout, cin, cin_pitch;
  float* dw;  // [9][cout][cin_pitch]
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                     const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* done_bar = empty_bar + kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work item decode: kh fastest so CTAs that run together share the same dY / X pixels in L2
  int item = blockIdx.x;
  const int kh = item % 3; item /= 3;
  const int cic = item % p.ci_chunks; item /= p.ci_chunks;
  const int cot = item % p.co_tiles; item /= p.co_tiles;
  const int ks = item;  // split index
  const int per = (p.p_tiles + p.ksplit - 1) / p.ksplit;
  const int pt_begin = ks * per;
  const int pt_end = min(p.p_tiles, pt_begin + per);
  const int n_kb = max(0, pt_end - pt_begin);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_dy);
    prefetch_tmap(&tmap_x);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      mbar_init(done_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // TMA producer: warp-uniform loop, one elected lane issues
    int stage = 0; uint32_t phase = 0;
    for (int pt = pt_begin; pt < pt_end; ++pt) {
      const int twi = pt % p.tiles_w;
      const int thi = (pt / p.tiles_w) % p.tiles_h;
      const int tni = pt / (p.tiles_w * p.tiles_h);
      const int w0 = twi * p.tw, h0 = thi * p.th, n0 = tni * p.tn;
      uint8_t* st = smem + (size_t)stage * kStageBytes;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
        tma_load_4d(&tmap_dy, &full_bar[stage], st, cot * 128, w0, h0, n0);
        tma_load_4d(&tmap_dy, &full_bar[stage], st + kTileBytes, cot * 128 + 64, w0, h0, n0);
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)  // X is the padded activation: pixel (h,w) + tap (kh,kw) -> (h+kh, w+kw)
          tma_load_4d(&tmap_x, &full_bar[stage], st + (2 + kw) * kTileBytes, cic * 64, w0 + kw, h0 + kh, n0);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform loop, precomputed descriptor words (see common.cuh), tcgen05 under elect_one()
    const uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);  // both operands MN-major
    constexpr uint32_t hi = desc_hi(1024, kLayoutSW128);
    // MN-major SW128: 64 channels contiguous (one 128 B row per pixel); next 64-channel block at LBO;
    // 8-pixel groups at SBO = 1024 B; a 16-pixel k-step advances the start address by 2048 B.
    const uint32_t lo0 = desc_lo(smem_u32(smem), kTileBytes);
    int stage = 0; uint32_t phase = 0;
    for (int kb = 0; kb < n_kb; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t a_lo = lo0 + (uint32_t)stage * (kStageBytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
          for (int k = 0; k < kPix / 16; ++k)
            umma_bf16_w(tmem_base + kw * 64, a_lo + k * (2048 >> 4), hi, a_lo + (((2 + kw) * kTileBytes + k * 2048) >> 4), hi, idesc,
                        (kb | k) != 0);
        }
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int co = cot * 128 + q * 32 + lane;
    if (n_kb > 0) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int kw = 0; kw < 3; ++kw) {
        float* dst_row = p.dw + ((size_t)(kh * 3 + kw) * p.cout + co) * p.cin_pitch + cic * 64;
#pragma unroll 1
        for (int c = 0; c < 64; c += 16) {
          float v[16];
          tmem_ld16(t_addr + kw * 64 + c, v);
          if (co < p.cout) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const int ci = cic * 64 + c + i;
              if (ci + 3 < p.cin_pitch) {
                red_add_v4(dst_row + c + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
              } else {
                for (int j = 0; j < 4; ++j)
                  if (ci + j < p.cin_pitch) atomicAdd(dst_row + c + i + j, v[i + j]);
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

void pick_tile64(int W, int H, int N, int* tw_o, int* th_o, int* tn_o) {
  double best = 1e30;
  for (int tw = 1; tw <= 64; tw <<= 1)
    for (int th = 1; tw * th <= 64; th <<= 1) {
      const int tn = 64 / (tw * th);
      const double tiles = (double)ceil_div(W, tw) * ceil_div(H, th) * ceil_div(N, tn);
      const double cost = tiles * (1.0 + 0.5 / tw);
      if (cost < best - 1e-9) { best = cost; *tw_o = tw; *th_o = th; *tn_o = tn; }
    }
}

__global__ void wgrad_unpack_kernel(const float* __restrict__ packed, float* __restrict__ grad, int cout, int cin,
                                    int cin_pitch, float scale, int accumulate) {
  // grad is OIHW [cout][cin][3][3]; packed is [9][cout][cin_pitch]
  const int total = cout * cin * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % 9;
    const int ci = (i / 9) % cin;
    const int co = i / (9 * cin);
    const float v = packed[((size_t)tap * cout + co) * cin_pitch + ci] * scale;
    grad[i] = accumulate ? grad[i] + v : v;
  }
}

constexpr int kMaxBatchJobs = 48;
struct UnpackBatch { WgradUnpackJob job[kMaxBatchJobs]; int first_block[kMaxBatchJobs + 1]; int n; };

// one index decode per (co, ci) pair, nine taps per thread: 36 contiguous bytes of the OIHW gradient
__global__ void wgrad_unpack_batched_kernel(const UnpackBatch b, float scale, int accumulate) {
  int j = 0;
  while (j + 1 < b.n && (int)blockIdx.x >= b.first_block[j + 1]) ++j;
  const WgradUnpackJob& q = b.job[j];
  const int nblk = b.first_block[j + 1] - b.first_block[j];
  const int pairs = q.cout * q.cin;
  const size_t tap_stride = (size_t)q.cout * q.cin_pitch;
  for (int i = ((int)blockIdx.x - b.first_block[j]) * blockDim.x + threadIdx.x; i < pairs; i += nblk * blockDim.x) {
    const int co = i / q.cin, ci = i - co * q.cin;
    const float* src = q.packed + (size_t)co * q.cin_pitch + q.sl.position(ci);
    float* dst = q.grad + (size_t)i * 9;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float v = src[tap * tap_stride] * scale;
      dst[tap] = accumulate ? dst[tap] + v : v;
    }
  }
}

}  // namespace

// dy     : unpadded view [N][H][W][cout...]           (pad == 0, c_off % 8 == 0)
// x      : padded activation view [N][H+2][W+2][cin]   (pad == 1, c_off % 8 == 0)
// dw     : fp32 [9][cout][cin_pitch], cin_pitch % 4 == 0; zeroed here before accumulation
int conv3x3_wgrad_launch(const ActView& dy, const ActView& x, float* dw, int cin_pitch, cudaStream_t stream, bool pre_zeroed) {
  MIMO_CHECK((dy.pad == 0 || dy.pad == 2) && x.pad == 1, MIMO_ERR_ARG, "wgrad: dy must be dense or zero-tailed and x haloed");
  MIMO_CHECK(dy.N == x.N && dy.H == x.H && dy.W == x.W, MIMO_ERR_ARG, "wgrad: dy/x shape mismatch");
  MIMO_CHECK(dy.cpitch % 8 == 0 && dy.c_off % 8 == 0 && x.cpitch % 8 == 0 && x.c_off % 8 == 0, MIMO_ERR_ALIGN,
             "wgrad: channel pitch/offset must be multiples of 8");
  MIMO_CHECK(cin_pitch % 4 == 0 && cin_pitch >= x.C, MIMO_ERR_ALIGN, "wgrad: cin_pitch %d invalid", cin_pitch);
  if (conv3x3_wgrad_flat_ok(dy, x)) return conv3x3_wgrad_flat_launch(dy, x, dw, cin_pitch, stream, pre_zeroed);
  if (conv3x3_wgrad_flatk_ok(dy, x)) return conv3x3_wgrad_flatk_launch(dy, x, dw, cin_pitch, stream, pre_zeroed);
  note_kernel(6);
  WgradParams p{};
  p.n_img = dy.N; p.H = dy.H; p.W = dy.W;
  pick_tile64(p.W, p.H, p.n_img, &p.tw, &p.th, &p.tn);
  p.tiles_w = ceil_div(p.W, p.tw); p.tiles_h = ceil_div(p.H, p.th); p.tiles_n = ceil_div(p.n_img, p.tn);
  p.p_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  p.cout = dy.C; p.cin = x.C; p.cin_pitch = cin_pitch;
  p.co_tiles = ceil_div(p.cout, 128);
  p.ci_chunks = ceil_div(p.cin, 64);
  const int base_items = 3 * p.co_tiles * p.ci_chunks;
  // split the pixel range so that about two waves of CTAs exist, but keep >= 8 k-blocks per CTA
  int ksplit = ceil_div(2 * num_sms(), base_items);
  const int max_split = p.p_tiles / 8 > 0 ? p.p_tiles / 8 : 1;
  if (ksplit > max_split) ksplit = max_split;
  if (ksplit < 1) ksplit = 1;
  // avoid empty splits
  { const int per = ceil_div(p.p_tiles, ksplit); ksplit = ceil_div(p.p_tiles, per); }
  p.ksplit = ksplit;
  p.dw = dw;

  CUtensorMap tm_dy, tm_x;
  {
    uint64_t dims[4] = {(uint64_t)dy.C, (uint64_t)dy.W, (uint64_t)dy.H, (uint64_t)dy.N};
    uint64_t strides[3] = {(uint64_t)dy.cpitch * 2, (uint64_t)dy.wb() * dy.cpitch * 2, (uint64_t)dy.hb() * dy.wb() * dy.cpitch * 2};
    uint32_t box[4] = {64, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tn};
    int rc = encode_tmap_bf16(&tm_dy, dy.base + dy.c_off, 4, dims, strides, box, 1);
    if (rc) return rc;
  }
  {
    const int Hb = x.H + 2, Wb = x.W + 2;
    uint64_t dims[4] = {(uint64_t)x.C, (uint64_t)Wb, (uint64_t)Hb, (uint64_t)x.N};
    uint64_t strides[3] = {(uint64_t)x.cpitch * 2, (uint64_t)Wb * x.cpitch * 2, (uint64_t)Hb * Wb * x.cpitch * 2};
    uint32_t box[4] = {64, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tn};
    int rc = encode_tmap_bf16(&tm_x, x.base + x.c_off, 4, dims, strides, box, 1);
    if (rc) return rc;
  }
  if (!pre_zeroed) MIMO_CUDA(cudaMemsetAsync(dw, 0, (size_t)9 * p.cout * cin_pitch * sizeof(float), stream));
  const size_t smem_bytes = (size_t)kStages * kStageBytes + (2 * kStages + 1) * 8 + 16 + 1024;
  MIMO_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));   // per device -> per launch
  const int grid = base_items * p.ksplit;
  conv3x3_wgrad_kernel<<<grid, kThreads, smem_bytes, stream>>>(tm_dy, tm_x, p);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int wgrad_unpack_launch(const float* packed, float* grad_oihw, int cout, int cin, int cin_pitch, float scale, int accumulate,
                        cudaStream_t stream) {
  const int total = cout * cin * 9;
  const int block = 256;
  int grid = ceil_div(total, block);
  if (grid > 4 * num_sms()) grid = 4 * num_sms();
  wgrad_unpack_kernel<<<grid, block, 0, stream>>>(packed, grad_oihw, cout, cin, cin_pitch, scale, accumulate);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

int wgrad_unpack_batched_launch(const WgradUnpackJob* jobs, int n, float scale, int accumulate, cudaStream_t stream) {
  for (int base = 0; base < n; base += kMaxBatchJobs) {
    UnpackBatch b{};
    b.n = n - base < kMaxBatchJobs ? n - base : kMaxBatchJobs;
    int blocks = 0;
    for (int j = 0; j < b.n; ++j) {
      b.job[j] = jobs[base + j];
      b.first_block[j] = blocks;
      int nb = ceil_div(jobs[base + j].cout * jobs[base + j].cin, 256);
      if (nb > num_sms()) nb = num_sms();
      blocks += nb < 1 ? 1 : nb;
    }
    b.first_block[b.n] = blocks;
    wgrad_unpack_batched_kernel<<<blocks, 256, 0, stream>>>(b, scale, accumulate);
    MIMO_LAUNCH_CHECK();
  }
  return MIMO_OK;
}

}  // namespace mimo
