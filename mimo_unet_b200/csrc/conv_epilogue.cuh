// Shared epilogue of the tcgen05 convolution kernels (conv_igemm.cu, conv_flat.cu): 4 warps drain one 128-row fp32
// accumulator tile from TMEM, round it to bf16, stage it in shared memory and write it to global memory with
// coalesced 16-byte stores. For training-mode BatchNorm the per-channel sum / sum-of-squares of the STORED values
// are reduced with a transposed warp-shuffle butterfly (no serial loop over rows) and accumulated per CTA across all
// of its tiles (static schedule -> deterministic); each CTA writes ONE partial row at the end.
#pragma once
#include "common.cuh"

namespace mimo {

struct EpiArgs {
  int block_n;      // accumulator columns per tile (multiple of 16)
  int cout;         // true number of output channels
  int out_cpitch;   // channels per output pixel in memory (multiple of 8, >= cout)
  int stage_pitch;  // staging row pitch in bytes (block_n * 2 + 16)
  int stat_rows;    // rows of the statistics buffers (>= gridDim.x); rows without a CTA are zero-filled
  bf16* out;        // [pixels][out_cpitch]
  float* stat_sum;  // [stat_rows][out_cpitch] or nullptr
  float* stat_sq;
  const float* bias;  // optional per-cout bias (eval path) or nullptr; with `scale` it is the shift of v * scale + bias
  int relu;
  // ---- fused inference epilogue (eval-mode BatchNorm folded into the conv: v * scale[c] + bias[c] -> ReLU -> Dropout2d) ----
  const float* scale; // optional per-cout scale
  const float* drop;  // optional [N][cout] Dropout2d keep/scale factors (MC dropout), applied after the ReLU
  int halo;           // 1: `out` is the interior of a haloed [N][H+2][W+2] buffer (the consumer's conv input); 0: dense [N][H][W]
  int out_cmax;       // channels that may be written starting at `out` (multiple of 8; 0 = out_cpitch)
};

// what a launcher needs to know to fuse the eval-mode BatchNorm / ReLU / Dropout2d into the conv epilogue
struct ConvFuse {
  const float* scale;
  const float* shift;
  const float* drop;
  int relu;
  int halo;
};

// shared-memory scratch of the epilogue warps
struct EpiSmem {
  uint8_t* stage;   // [128][stage_pitch]
  int* row_pix;     // [128] output pixel index of each accumulator row, -1 = not stored
  float* wsum;      // [2][4][block_n] per-warp column partials (sum, sq)
  float* acc;       // [2][out_cpitch] per-CTA running statistics
};

__host__ __device__ inline size_t epi_smem_bytes(int block_n, int out_cpitch) {
  return (size_t)128 * (block_n * 2 + 16) + 128 * 4 + (size_t)8 * block_n * 4 + (size_t)2 * out_cpitch * 4;
}

__device__ __forceinline__ EpiSmem epi_carve(uint8_t* base, int block_n, int out_cpitch) {
  EpiSmem s;
  s.stage = base;
  s.row_pix = reinterpret_cast<int*>(base + (size_t)128 * (block_n * 2 + 16));
  s.wsum = reinterpret_cast<float*>(s.row_pix + 128);
  s.acc = s.wsum + 8 * block_n;
  return s;
}

// Transposed butterfly: every lane contributes 16 values (one per column); on return lane l holds the sum over the
// 32 lanes of column (l >> 1). 16 shuffles instead of 16 x 5.
__device__ __forceinline__ float warp_colsum16(const float v[16], int lane) {
  float a[8], b[4], c[2];
  {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float keep = up ? v[8 + i] : v[i];
      const float send = up ? v[i] : v[8 + i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float keep = up ? a[4 + i] : a[i];
      const float send = up ? a[i] : a[4 + i];
      b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = (lane & 4) != 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float keep = up ? b[2 + i] : b[i];
      const float send = up ? b[i] : b[2 + i];
      c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  const bool up = (lane & 2) != 0;
  const float keep = up ? c[1] : c[0];
  const float send = up ? c[0] : c[1];
  float d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;
}

// called once by the 128 epilogue threads before the first tile
__device__ __forceinline__ void epi_init(const EpiArgs& e, const EpiSmem& s, int et) {
  if (e.stat_sum != nullptr)
    for (int i = et; i < 2 * e.out_cpitch; i += 128) s.acc[i] = 0.f;
  named_bar_sync(1, 128);
}

// One tile. t_addr: TMEM address of this warp's lane quarter and this tile's accumulator; my_pix: output pixel of
// accumulator row (q*32+lane) or -1; co0: first output channel of the tile. The caller has already waited on the
// accumulator-full barrier; `tmem_empty_bar` is arrived on (one arrive per warp) as soon as TMEM has been drained.
__device__ __forceinline__ void epi_tile(const EpiArgs& e, const EpiSmem& s, uint32_t t_addr, int my_pix, int co0,
                                         uint64_t* tmem_empty_bar, int q, int lane, int et) {
  const int row = q * 32 + lane;
  s.row_pix[row] = my_pix;
  const float vmask = my_pix >= 0 ? 1.f : 0.f;
  uint8_t* my_row = s.stage + (size_t)row * e.stage_pitch;
  const bool stats = e.stat_sum != nullptr;
  for (int c = 0; c < e.block_n; c += 16) {
    float v[16];
    tmem_ld16(t_addr + c, v);
    if (e.bias != nullptr) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += (co0 + c + i < e.cout) ? __ldg(e.bias + co0 + c + i) : 0.f;
    }
    if (e.relu) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    bf16x8 lo, hi;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      lo.v[i] = __float2bfloat16_rn(v[i]);
      hi.v[i] = __float2bfloat16_rn(v[8 + i]);
    }
    *reinterpret_cast<bf16x8*>(my_row + c * 2) = lo;
    *reinterpret_cast<bf16x8*>(my_row + c * 2 + 16) = hi;
    if (stats) {
      // statistics of the values as stored (bf16), rows that are not stored contribute nothing
      float f[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        f[i] = __bfloat162float(lo.v[i]) * vmask;
        f[8 + i] = __bfloat162float(hi.v[i]) * vmask;
      }
      const float cs = warp_colsum16(f, lane);
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] *= f[i];
      const float cq = warp_colsum16(f, lane);
      if ((lane & 1) == 0) s.wsum[q * e.block_n + c + (lane >> 1)] = cs;
      else s.wsum[(4 + q) * e.block_n + c + (lane >> 1)] = cq;
    }
  }
  // all TMEM reads of this warp are done -> hand the accumulator back to the MMA warp
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(tmem_empty_bar);

  named_bar_sync(1, 128);  // staging tile + warp partials complete

  // ---- coalesced stores: 16-byte chunks, consecutive threads -> consecutive chunks of a pixel ----
  const int n_store = min(e.block_n, e.out_cpitch - co0);  // channels this tile owns in memory (multiple of 8)
  const int chunks = n_store >> 3;
  for (int id = et; id < 128 * chunks; id += 128) {
    const int r = id / chunks, ch = id - r * chunks;
    const int pix = s.row_pix[r];
    if (pix < 0) continue;
    const bf16x8 val = *reinterpret_cast<const bf16x8*>(s.stage + (size_t)r * e.stage_pitch + ch * 16);
    *reinterpret_cast<bf16x8*>(e.out + (size_t)pix * e.out_cpitch + co0 + ch * 8) = val;
  }
  if (stats) {
    for (int col = et; col < n_store; col += 128) {
      const float* ws = s.wsum + col;
      const int bn = e.block_n;
      s.acc[co0 + col] += (ws[0] + ws[bn]) + (ws[2 * bn] + ws[3 * bn]);
      s.acc[e.out_cpitch + co0 + col] += (ws[4 * bn] + ws[5 * bn]) + (ws[6 * bn] + ws[7 * bn]);
    }
  }
  named_bar_sync(1, 128);  // staging buffer + partials free for the next tile
}

// called once by the 128 epilogue threads after the last tile
__device__ __forceinline__ void epi_finish(const EpiArgs& e, const EpiSmem& s, int et) {
  if (e.stat_sum == nullptr) return;
  for (int col = et; col < e.out_cpitch; col += 128) {
    e.stat_sum[(size_t)blockIdx.x * e.out_cpitch + col] = s.acc[col];
    e.stat_sq[(size_t)blockIdx.x * e.out_cpitch + col] = s.acc[e.out_cpitch + col];
    for (int r = blockIdx.x + gridDim.x; r < e.stat_rows; r += gridDim.x) {
      e.stat_sum[(size_t)r * e.out_cpitch + col] = 0.f;
      e.stat_sq[(size_t)r * e.out_cpitch + col] = 0.f;
    }
  }
}

}  // namespace mimo
