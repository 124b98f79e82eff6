// "Flat" weight gradient of the reflect-padded 3x3 convolution for the HBM-bound layers (few channels, many pixels);
// companion of conv_flat.cu, replaces cuDNN wgrad for the reference's nn.Conv2d(k=3) layers (components.py:23,26)
// with <= 64 input and output channels.
//
//   dW[kh][kw][co][ci] = sum_P dY[P][co] * X[P + kh*(W+2) + kw][ci]
//
// where P runs over ALL positions of the [N][H+2][W+2] buffer grid: dY lives in the zero-tail layout (pad == 2), so
// the positions that fall on the tail contribute exactly zero, and X is the reflect-haloed conv input (pad == 1) with
// the same row pitch. Every CTA owns a contiguous range of 128-position tiles; per tile ONE TMA load of dY (128 rows)
// and ONE of X: X streams through a shared-memory ring of 128-row chunks exactly like the input of conv_flat.cu, and the
// (kh, kw) shift of a tap is a row offset of the UMMA descriptor start address inside the SWIZZLE_128B ring.
// Both operands are MN-major (channels contiguous), the reduction (K) dimension runs over positions.
// The tap accumulators stay in TMEM for the whole persistent CTA; they are flushed once at the end with fp32
// reductions into the packed [9][cout][cin_pitch] gradient.
//
// Orientation: X is the UMMA M operand, dY the N operand (N = round_up(cout, 16) <= 64). An MN-major M operand of
// 128 rows is two 64-channel blocks LBO bytes apart; with LBO = 128 B (one position) the second block is X shifted by
// one more position, i.e. the NEXT kw tap: one MMA accumulates taps (kh,0) and (kh,1) in lanes 0..63 / 64..127, a
// second one tap (kh,2) (its upper lanes are don't-care). 48 instead of 72 MMAs per tile: these MMAs are bound by the
// shared-memory operand reads (~64 cycles each at M=128), not by the math.
#include "common.cuh"
#include "ops.h"

#include <stdlib.h>

namespace mimo {
namespace {

constexpr int kBlockK = 128;                 // positions per tile
constexpr int kChunkBytes = kBlockK * 128;   // 16 KB: one dY tile / one X ring slot
constexpr int kDyStages = 3;
constexpr int kMaxSlots = 10;
constexpr int kThreads = 192;

struct WgFlatParams {
  int wb;
  int m_tiles, tiles_per_cta;
  int slots;       // X ring slots S (+ mirror of slot 0 stored as slot S)
  int nc;          // X chunks a tile touches: ceil((2*wb + 131) / 128)
  int n_cols;      // UMMA N = round_up(cout, 16) <= 64
  int cout, cin, cin_pitch;
  float* dw;       // [9][cout][cin_pitch], zeroed by the launcher
};

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_wgrad_flat_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                          const WgFlatParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [dY stages: 3 x 16 KB][X ring: (S+1) x 16 KB][barriers]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_dy = smem;
  uint8_t* smem_x = smem + (size_t)kDyStages * kChunkBytes;
  const int S = p.slots;
  uint64_t* x_full = reinterpret_cast<uint64_t*>(smem_x + (size_t)(S + 1) * kChunkBytes);
  uint64_t* x_empty = x_full + kMaxSlots;
  uint64_t* dy_full = x_empty + kMaxSlots;
  uint64_t* dy_empty = dy_full + kDyStages;
  uint64_t* done_bar = dy_empty + kDyStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t_begin = blockIdx.x * p.tiles_per_cta;
  const int n_tiles = min(p.tiles_per_cta, p.m_tiles - t_begin);
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_dy);
    prefetch_tmap(&tmap_x);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 1); }
      for (int s = 0; s < kDyStages; ++s) { mbar_init(&dy_full[s], 1); mbar_init(&dy_empty[s], 1); }
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
    // step c loads X chunk c (rows [128 (t_begin + c), +128)) and, once the first tile's window is on its way, the dY
    // tile c - (nc - 1)
    const int n_chunks = n_tiles + p.nc - 1;
    int xs = 0; uint32_t xphase = 0;
    int ds = 0; uint32_t dphase = 0;
    long long row0 = (long long)t_begin * kBlockK;
    for (int c = 0; c < n_chunks; ++c, row0 += kBlockK) {
      mbar_wait(&x_empty[xs], xphase ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&x_full[xs], xs == 0 ? 2 * kChunkBytes : kChunkBytes);
        tma_load_2d(&tmap_x, &x_full[xs], smem_x + (size_t)xs * kChunkBytes, 0, (int)row0);
        if (xs == 0) tma_load_2d(&tmap_x, &x_full[xs], smem_x + (size_t)S * kChunkBytes, 0, (int)row0);  // mirror
      }
      __syncwarp();
      if (++xs == S) { xs = 0; xphase ^= 1; }
      const int i = c - (p.nc - 1);
      if (i >= 0) {
        mbar_wait(&dy_empty[ds], dphase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&dy_full[ds], kChunkBytes);
          tma_load_2d(&tmap_dy, &dy_full[ds], smem_dy + (size_t)ds * kChunkBytes, 0, (int)((long long)(t_begin + i) * kBlockK));
        }
        __syncwarp();
        if (++ds == kDyStages) { ds = 0; dphase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // warp-uniform loop; descriptor words precomputed, per-MMA work = 32-bit adds with immediates (common.cuh)
    const uint32_t idesc = make_idesc_bf16(128, p.n_cols, 1, 1);  // both operands MN-major
    constexpr uint32_t hi = desc_hi(1024, kLayoutSW128);
    // MN-major SW128: 64 channels contiguous (one 128 B row per position); 8-position groups at SBO = 1024 B;
    // a 16-position k-step advances the start address by 2048 B. M operand (X): LBO = 128 B -> lanes 64..127 read
    // the rows one position further (next kw tap). N operand (dY): single 64-channel block, LBO unused.
    const uint32_t x_lo0 = desc_lo(smem_u32(smem_x), 128);
    const uint32_t y_lo0 = desc_lo(smem_u32(smem_dy), 1024);
    const uint32_t ncols = (uint32_t)p.n_cols;
    const uint32_t ring_rows = (uint32_t)S * kBlockK;
    uint32_t win_row[6];  // window start (rows) of accumulator (kh, pair): kh*wb + 2*pair
#pragma unroll
    for (int a = 0; a < 6; ++a) win_row[a] = (uint32_t)((a >> 1) * p.wb + (a & 1) * 2);
    int waited = 0, wslot = 0; uint32_t wphase = 0;
    int slot = 0;
    int ds = 0; uint32_t dphase = 0;
    uint32_t accumulate = 0;
    for (int i = 0; i < n_tiles; ++i) {
      while (waited < i + p.nc) {
        mbar_wait(&x_full[wslot], wphase);
        ++waited;
        if (++wslot == S) { wslot = 0; wphase ^= 1; }
      }
      mbar_wait(&dy_full[ds], dphase);
      tc_fence_after();
      const uint32_t base_row = (uint32_t)slot * kBlockK;
      const uint32_t n_lo = y_lo0 + (uint32_t)ds * (kChunkBytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          uint32_t r = base_row + win_row[a];
          if (r >= ring_rows) r -= ring_rows;
          const uint32_t m_lo = x_lo0 + r * 8;
          const uint32_t d_addr = tmem_base + (uint32_t)a * ncols;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16_w(d_addr, m_lo + k * (2048 >> 4), hi, n_lo + k * (2048 >> 4), hi, idesc, accumulate | (uint32_t)k);
        }
        umma_commit(&x_empty[slot]);   // X chunk i is dead once these MMAs have read it
        umma_commit(&dy_empty[ds]);
      }
      __syncwarp();
      accumulate = 1;
      if (++slot == S) slot = 0;
      if (++ds == kDyStages) { ds = 0; dphase ^= 1; }
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  } else {
    // ===================== flush =====================
    // accumulator (kh, pair): lanes 0..63 = tap (kh, 2*pair), lanes 64..127 = tap (kh, 2*pair+1); lane % 64 = ci, column = co
    const int q = warp & 3;
    const int lane_m = q * 32 + lane;
    const int ci = lane_m & 63;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int a = 0; a < 6; ++a) {
      const int kh = a >> 1, kw = (a & 1) * 2 + (lane_m >> 6);
      const bool lane_ok = kw < 3 && ci < p.cin;
      float* dst = p.dw + ((size_t)(kh * 3 + kw) * p.cout) * p.cin_pitch + ci;
#pragma unroll 1
      for (int c = 0; c < p.n_cols; c += 16) {
        float v[16];
        tmem_ld16(t_addr + a * p.n_cols + c, v);
        if (lane_ok) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c + i < p.cout) atomicAdd(dst + (size_t)(c + i) * p.cin_pitch, v[i]);
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

int wg_plan_slots(int nc, size_t* smem_bytes) {
  const int fixed = kDyStages * kChunkBytes + (2 * kMaxSlots + 2 * kDyStages + 1) * 8 + 16 + 1024;
  int slots = (227 * 1024 - fixed) / kChunkBytes - 1;
  if (slots > kMaxSlots) slots = kMaxSlots;
  if (slots < nc + 1) return 0;
  *smem_bytes = (size_t)fixed + (size_t)(slots + 1) * kChunkBytes;
  return slots;
}

}  // namespace

bool conv3x3_wgrad_flat_ok(const ActView& dy, const ActView& x) {
  static const int enabled = getenv("MIMO_WGRAD_FLAT") ? atoi(getenv("MIMO_WGRAD_FLAT")) : 1;
  if (!enabled) return false;
  if (dy.pad != 2 || x.pad != 1) return false;
  if (dy.C > 64 || x.C > 64) return false;  // 6 accumulators x round_up(cout, 16) <= 384 TMEM columns
  if ((long long)x.N * x.hb() * x.wb() >= (1ll << 31) - 4096) return false;
  size_t smem;
  return wg_plan_slots(ceil_div(2 * x.wb() + 131, 128), &smem) > 0;
}

int conv3x3_wgrad_flat_launch(const ActView& dy, const ActView& x, float* dw, int cin_pitch, cudaStream_t stream, bool pre_zeroed) {
  note_kernel(4);
  WgFlatParams p{};
  p.wb = x.wb();
  const long long total_pos = (long long)x.N * x.hb() * x.wb();
  p.m_tiles = (int)ceil_div_ll(total_pos, kBlockK);
  p.tiles_per_cta = ceil_div(p.m_tiles, num_sms());
  const int grid = ceil_div(p.m_tiles, p.tiles_per_cta);
  p.nc = ceil_div(2 * p.wb + 131, 128);
  p.cout = dy.C; p.cin = x.C; p.cin_pitch = cin_pitch;
  p.n_cols = round_up(dy.C, 16);
  p.dw = dw;
  size_t smem_bytes = 0;
  p.slots = wg_plan_slots(p.nc, &smem_bytes);
  MIMO_CHECK(p.slots > 0, MIMO_ERR_ARG, "wgrad_flat: ring of %d chunks does not fit shared memory", p.nc);

  CUtensorMap tm_dy, tm_x;
  {
    uint64_t dims[2] = {(uint64_t)dy.C, (uint64_t)total_pos};
    uint64_t strides[1] = {(uint64_t)dy.cpitch * 2};
    uint32_t box[2] = {64, (uint32_t)kBlockK};
    int rc = encode_tmap_bf16(&tm_dy, dy.base + dy.c_off, 2, dims, strides, box, 1);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)x.C, (uint64_t)total_pos};
    uint64_t strides[1] = {(uint64_t)x.cpitch * 2};
    uint32_t box[2] = {64, (uint32_t)kBlockK};
    int rc = encode_tmap_bf16(&tm_x, x.base + x.c_off, 2, dims, strides, box, 1);
    if (rc) return rc;
  }
  if (!pre_zeroed) MIMO_CUDA(cudaMemsetAsync(dw, 0, (size_t)9 * p.cout * cin_pitch * sizeof(float), stream));
  MIMO_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));   // per launch: the attribute is per DEVICE, a process-wide "done" flag would skip the other GPUs
  conv3x3_wgrad_flat_kernel<<<grid, kThreads, smem_bytes, stream>>>(tm_dy, tm_x, p);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

}  // namespace mimo
