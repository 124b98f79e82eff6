// "Flat" weight gradient of the reflect-padded 3x3 convolution for the HBM-bound layers (few channels, many pixels);
// companion of conv_flat.cu, replaces cuDNN wgrad for the reference's nn.Conv2d(k=3) layers (components.py:23,26)
// with <= 64 input and output channels.
//
//   dW[kh][kw][co][ci] = sum_P dY[P][co] * X[P + kh*(W+2) + kw][ci]
//
// where P runs over ALL positions of the [N][H+2][W+2] buffer grid: dY lives in the zero-tail layout (pad == 2), so
// the positions that fall on the tail contribute exactly zero, and X is the reflect-haloed conv input (pad == 1) with
// the same row pitch. Per tile of 128 consecutive positions: ONE TMA load of dY (128 rows) and THREE of X (130 rows,
// one per kh); the kw shift is a +128-byte shift of the UMMA descriptor start address inside the SWIZZLE_128B segment.
// Both operands are MN-major (channels contiguous), the reduction (K) dimension runs over positions.
// All nine tap accumulators (9 x N fp32 columns, N <= 48) stay in TMEM for the whole persistent CTA; they are
// flushed once at the end with fp32 reductions into the packed [9][cout][cin_pitch] gradient.
//
// Orientation: the operand with <= 48 (padded) channels is the UMMA N operand; the other one is the M operand (its
// 64-channel block fills accumulator lanes 0..63, lanes 64..127 are don't-care).
#include "common.cuh"
#include "ops.h"

#include <stdlib.h>

namespace mimo {
namespace {

constexpr int kBlockK = 128;                 // positions per tile
constexpr int kSegRows = kBlockK + 2;
constexpr int kSegBytes = 17 * 1024;         // 130 x 128 B rounded up to the swizzle repeat
constexpr int kDyBytes = kBlockK * 128;      // 16 KB
constexpr int kStageBytes = kDyBytes + 3 * kSegBytes;
constexpr int kStages = 3;
constexpr int kThreads = 192;

struct WgFlatParams {
  int wb;
  long long total_pos;
  int m_tiles;
  int n_cols;      // UMMA N (multiple of 16, <= 48)
  int swap;        // 0: M = dY (co), N = X (ci);  1: M = X (ci), N = dY (co)
  int cout, cin, cin_pitch;
  float* dw;       // [9][cout][cin_pitch], zeroed by the launcher
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_wgrad_flat_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                          const WgFlatParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* done_bar = empty_bar + kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x) {
      const long long p0 = (long long)t * kBlockK;
      uint8_t* st = smem + (size_t)stage * kStageBytes;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[stage], kDyBytes + 3 * kSegRows * 128);
        tma_load_2d(&tmap_dy, &full_bar[stage], st, 0, (int)p0);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) tma_load_2d(&tmap_x, &full_bar[stage], st + kDyBytes + kh * kSegBytes, 0, (int)(p0 + (long long)kh * p.wb));
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // warp-uniform loop; descriptor words precomputed, per-MMA work = 32-bit adds with immediates (common.cuh)
    const uint32_t idesc = make_idesc_bf16(128, p.n_cols, 1, 1);  // both operands MN-major
    constexpr uint32_t hi = desc_hi(1024, kLayoutSW128);
    // MN-major SW128: 64 channels contiguous (one 128 B row per position); 8-position groups at SBO = 1024 B;
    // a 16-position k-step advances the start address by 2048 B. The M operand's second 64-channel block
    // (LBO) only feeds the don't-care accumulator lanes 64..127.
    const uint32_t lo0 = desc_lo(smem_u32(smem), 1024);
    const uint32_t ncols = (uint32_t)p.n_cols;
    // offsets (16-byte units) of the dY tile / the X segments inside a stage, by operand role
    const uint32_t m_off = p.swap ? (kDyBytes >> 4) : 0u;
    const uint32_t n_off = p.swap ? 0u : (kDyBytes >> 4);
    const uint32_t m_x = p.swap ? 1u : 0u, n_x = p.swap ? 0u : 1u;  // which operand carries the tap shift
    int stage = 0; uint32_t phase = 0;
    uint32_t accumulate = 0;
    for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t st_lo = lo0 + (uint32_t)stage * (kStageBytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const uint32_t shift = (uint32_t)((kh * kSegBytes + kw * 128) >> 4);  // X rows shifted by kh rows + kw positions
            const uint32_t m_lo = st_lo + m_off + m_x * shift;
            const uint32_t n_lo = st_lo + n_off + n_x * shift;
            const uint32_t d_addr = tmem_base + (uint32_t)(kh * 3 + kw) * ncols;
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_bf16_w(d_addr, m_lo + k * (2048 >> 4), hi, n_lo + k * (2048 >> 4), hi, idesc, accumulate | (uint32_t)k);
          }
        }
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      accumulate = 1;
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  } else if ((warp & 3) < 2) {
    // ===================== flush (TMEM lanes 0..63 carry the M operand's channels) =====================
    const int q = warp & 3;
    const int m = q * 32 + lane;  // channel index of the M operand
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
#pragma unroll 1
      for (int c = 0; c < p.n_cols; c += 16) {
        float v[16];
        tmem_ld16(t_addr + tap * p.n_cols + c, v);
        if (!p.swap) {
          // m = co, columns = ci
          if (m < p.cout) {
            float* dst_row = p.dw + ((size_t)tap * p.cout + m) * p.cin_pitch;
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              if (c + i + 3 < p.cin_pitch) red_add_v4(dst_row + c + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
        } else {
          // m = ci, columns = co
          if (m < p.cin) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c + i < p.cout) atomicAdd(p.dw + ((size_t)tap * p.cout + c + i) * p.cin_pitch + m, v[i]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

bool conv3x3_wgrad_flat_ok(const ActView& dy, const ActView& x) {
  static const int enabled = getenv("MIMO_WGRAD_FLAT") ? atoi(getenv("MIMO_WGRAD_FLAT")) : 1;
  if (!enabled) return false;
  if (dy.pad != 2 || x.pad != 1) return false;
  if (dy.C > 64 || x.C > 64) return false;
  if (round_up(dy.C, 16) > 48 && round_up(x.C, 16) > 48) return false;  // 9 accumulators must fit 512 TMEM columns
  if ((long long)x.N * x.hb() * x.wb() >= (1ll << 31) - 256) return false;
  return true;
}

int conv3x3_wgrad_flat_launch(const ActView& dy, const ActView& x, float* dw, int cin_pitch, cudaStream_t stream) {
  WgFlatParams p{};
  p.wb = x.wb();
  p.total_pos = (long long)x.N * x.hb() * x.wb();
  p.m_tiles = (int)ceil_div_ll(p.total_pos, kBlockK);
  p.cout = dy.C; p.cin = x.C; p.cin_pitch = cin_pitch;
  // N operand: prefer X (ci) so the flush can use vector reductions along ci
  p.swap = round_up(x.C, 16) <= 48 ? 0 : 1;
  p.n_cols = p.swap ? round_up(dy.C, 16) : round_up(x.C, 16);
  p.dw = dw;

  CUtensorMap tm_dy, tm_x;
  {
    uint64_t dims[2] = {(uint64_t)dy.C, (uint64_t)p.total_pos};
    uint64_t strides[1] = {(uint64_t)dy.cpitch * 2};
    uint32_t box[2] = {64, (uint32_t)kBlockK};
    int rc = encode_tmap_bf16(&tm_dy, dy.base + dy.c_off, 2, dims, strides, box, 1);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)x.C, (uint64_t)p.total_pos};
    uint64_t strides[1] = {(uint64_t)x.cpitch * 2};
    uint32_t box[2] = {64, (uint32_t)kSegRows};
    int rc = encode_tmap_bf16(&tm_x, x.base + x.c_off, 2, dims, strides, box, 1);
    if (rc) return rc;
  }
  MIMO_CUDA(cudaMemsetAsync(dw, 0, (size_t)9 * p.cout * cin_pitch * sizeof(float), stream));
  const size_t smem_bytes = (size_t)kStages * kStageBytes + (2 * kStages + 1) * 8 + 16 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    MIMO_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    attr_set = true;
  }
  const int grid = p.m_tiles < num_sms() ? p.m_tiles : num_sms();
  conv3x3_wgrad_flat_kernel<<<grid, kThreads, smem_bytes, stream>>>(tm_dy, tm_x, p);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

}  // namespace mimo
