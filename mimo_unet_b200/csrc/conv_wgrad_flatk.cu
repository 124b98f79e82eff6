// "Flat, stream-K" weight gradient of the reflect-padded 3x3 convolution for the layers with more than 64 input or
// output channels (the U-Net core: reference model.py:190-243, nn.Conv2d of components.py:23,26) -- replaces cuDNN
// wgrad there. Companion of conv_wgrad_flat.cu (same flattened addressing) and successor of the 4-D split-K kernel
// (conv_wgrad.cu), which was bound by the TMA / L2 row rate (5 boxes of 64 rows per 12 MMAs) and by whole-wave
// quantisation of its one-shot CTAs.
//
//   dW[kh][kw][co][ci] = sum_P dY[P][co] * X[P + kh*(W+2) + kw][ci]
//
// P runs over ALL positions of the [N][H+2][W+2] grid: dY lives in the zero-tail layout (pad == 2) so tail positions
// contribute exactly zero, X is the reflect-haloed conv input (pad == 1) with the same row pitch.
//
// GEMM per work item (128-co tile, one or two 64-ci chunks, kernel row kh), K = positions in blocks of 128:
//   M operand = dY  (MN-major, 128 co = two 64-channel boxes LBO = 16 KB apart)
//   N operand = X   (MN-major, N = 192 = THREE kw taps x 64 ci: the 64-channel blocks of an MN-major operand are LBO
//                    bytes apart, and with LBO = 128 B = one position the second / third block is X shifted by one /
//                    two positions, i.e. the next kw tap; SWIZZLE_128B is absolute-address based on B200, so the
//                    row-shifted blocks read what TMA wrote)
//   -> ONE M128 x N192 x K16 MMA per 16 positions and ci chunk does all three kw taps (96 math cycles, above the
//      ~64-cycle shared-memory A-operand floor), and a stage needs 256 dY rows + 130 X rows per ci chunk instead of 640.
// Accumulators: 192 fp32 TMEM columns per ci chunk (lane = co, column = kw*64 + ci).
//
// Scheduling is stream-K: the (item, k-block) space, weighted by the number of ci chunks, is cut into gridDim.x equal
// contiguous ranges, so every persistent CTA issues the same number of MMAs whatever the layer shape; a CTA flushes
// its accumulators with red.global.add.v4.f32 into the packed fp32 [9][cout][cin_pitch] gradient whenever its range
// crosses an item boundary (at most a few flushes per CTA).
#include "common.cuh"
#include "ops.h"

#include <stdlib.h>

namespace mimo {
namespace {

constexpr int kBlockK = 128;                    // positions per k-block
constexpr int kDyBox = kBlockK * 128;           // 16 KB: 128 positions x 64 channels
constexpr int kXSlot = 17 * 1024;               // 130 rows (128 + the kw = 1, 2 shifts), rounded to the swizzle repeat
constexpr int kStageBytes = 2 * kDyBox + 2 * kXSlot;   // 66 KB
constexpr int kStages = 3;
constexpr int kThreads = 192;
constexpr int kAccCols = 192;                   // per ci chunk: 3 kw x 64 ci

struct WgFlatKParams {
  int wb;
  int n_kb;            // k-blocks (128 positions) per item
  int n_pairs;         // ci chunk pairs per co tile
  int ci_chunks;
  int n2_items;        // items with two ci chunks: co_tiles * n_pairs * 3
  long long W2, Wtot;  // weight (k-block x ci chunks) of the two-chunk items / of everything
  int cout, cin_pitch;
  float* dw;           // [9][cout][cin_pitch], zeroed by the launcher
  int ko;              // diagnostic knock-outs (env MIMO_WGK_KO): 1 no reductions, 2 no memset, 4 no MMAs
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Walks the segments (item, k-block range) of this CTA's slice of the weighted stream-K space. All three warp roles
// run the same walk, so they agree on the sequence without communicating.
struct SegWalk {
  long long w, w_end;
  int kb0, kb1, nci, cot, cic0, kh;
  __host__ __device__ SegWalk(const WgFlatKParams& p, unsigned cta, unsigned grid) {
    w = cut(p, cta, grid);
    w_end = cut(p, cta + 1, grid);
  }
  static __host__ __device__ long long cut(const WgFlatKParams& p, unsigned c, unsigned grid) {
    long long v = (long long)c * p.Wtot / (long long)grid;
    if (v < p.W2) v &= ~1ll;   // a k-block of a two-chunk item weighs 2: cuts fall on k-block edges
    return v;
  }
  __host__ __device__ bool next(const WgFlatKParams& p) {
    if (w >= w_end) return false;
    long long seg_end;
    if (w < p.W2) {
      const long long per = 2ll * p.n_kb;
      const int item = (int)(w / per);
      kb0 = (int)((w - item * per) >> 1);
      seg_end = (item + 1) * per;
      if (seg_end > w_end) seg_end = w_end;
      kb1 = kb0 + (int)((seg_end - w) >> 1);
      nci = 2;
      kh = item % 3;
      const int r = item / 3;
      cot = r / p.n_pairs;
      cic0 = 2 * (r - cot * p.n_pairs);
    } else {
      const long long v = w - p.W2;
      const int item = (int)(v / p.n_kb);
      kb0 = (int)(v - (long long)item * p.n_kb);
      seg_end = p.W2 + (long long)(item + 1) * p.n_kb;
      if (seg_end > w_end) seg_end = w_end;
      kb1 = kb0 + (int)(seg_end - w);
      nci = 1;
      kh = item % 3;
      cot = item / 3;
      cic0 = p.ci_chunks - 1;
    }
    w = seg_end;
    return true;
  }
};

// shape -> schedule parameters (shared by the launcher and the host-side schedule dump used by the CPU tests)
void plan_schedule(WgFlatKParams& p, int cout, int cin, long long total_pos) {
  p.n_kb = (int)ceil_div_ll(total_pos, kBlockK);
  const int co_tiles = ceil_div(cout, 128);
  p.ci_chunks = ceil_div(cin, 64);
  p.n_pairs = p.ci_chunks / 2;
  p.n2_items = co_tiles * p.n_pairs * 3;
  const int n1_items = (p.ci_chunks & 1) ? co_tiles * 3 : 0;
  p.W2 = (long long)p.n2_items * p.n_kb * 2;
  p.Wtot = p.W2 + (long long)n1_items * p.n_kb;
}

int plan_grid(const WgFlatKParams& p, int sms) {
  // every CTA should own at least a few k-blocks; tiny layers use fewer CTAs
  long long grid = p.Wtot / 4;
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  return (int)grid;
}

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_wgrad_flatk_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                           const __grid_constant__ CUtensorMap tmap_x2, const WgFlatKParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (dY box co 0..63 | dY box co 64..127 | X slot chunk 0 | X slot chunk 1)][barriers]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* acc_full = empty_bar + kStages;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_dy);
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_x2);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      mbar_init(acc_full, 1);
      mbar_init(acc_empty, 4);  // one arrive per flush warp
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
    // ===================== TMA producer (warp-uniform loops, one elected lane issues) =====================
    SegWalk sw(p, blockIdx.x, gridDim.x);
    int stage = 0; uint32_t phase = 0;
    while (sw.next(p)) {
      const uint32_t tx = (uint32_t)(2 * kDyBox + sw.nci * (kDyBox + 2 * 128));
      for (int kb = sw.kb0; kb < sw.kb1; ++kb) {
        uint8_t* st = smem + (size_t)stage * kStageBytes;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          const int row = kb * kBlockK;
          const int xrow = row + sw.kh * p.wb;
          mbar_arrive_expect_tx(&full_bar[stage], tx);
          tma_load_2d(&tmap_dy, &full_bar[stage], st, sw.cot * 128, row);
          tma_load_2d(&tmap_dy, &full_bar[stage], st + kDyBox, sw.cot * 128 + 64, row);
          for (int j = 0; j < sw.nci; ++j) {
            uint8_t* xs = st + 2 * kDyBox + j * kXSlot;
            tma_load_2d(&tmap_x, &full_bar[stage], xs, (sw.cic0 + j) * 64, xrow);
            tma_load_2d(&tmap_x2, &full_bar[stage], xs + kDyBox, (sw.cic0 + j) * 64, xrow + kBlockK);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loops, precomputed descriptor words) =====================
    const uint32_t idesc = make_idesc_bf16(128, kAccCols, 1, 1);  // both operands MN-major
    constexpr uint32_t hi = desc_hi(1024, kLayoutSW128);
    // MN-major SW128: 64 channels contiguous (one 128 B row per position), 8-position groups at SBO = 1024 B, a
    // 16-position k-step advances the start address by 2048 B. dY: second 64-co block at LBO = 16 KB (the other box).
    // X: 64-ci block j at LBO * j = j positions further = kw tap j.
    const uint32_t a_lo0 = desc_lo(smem_u32(smem), kDyBox);
    const uint32_t b_lo0 = desc_lo(smem_u32(smem) + 2 * kDyBox, 128);
    SegWalk sw(p, blockIdx.x, gridDim.x);
    int stage = 0; uint32_t phase = 0;
    uint32_t n = 0;
    while (sw.next(p)) {
      mbar_wait(acc_empty, (n & 1u) ^ 1u);
      tc_fence_after();
      for (int kb = sw.kb0; kb < sw.kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + (uint32_t)stage * (kStageBytes >> 4);
        const uint32_t b_lo = b_lo0 + (uint32_t)stage * (kStageBytes >> 4);
        const uint32_t acc = kb != sw.kb0;   // the first k-step of a segment overwrites the accumulator
        if (elect_one()) {
          if (!(p.ko & 4)) {
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16_w(tmem_base, a_lo + k * (2048 >> 4), hi, b_lo + k * (2048 >> 4), hi, idesc, acc | (uint32_t)k);
          if (sw.nci == 2) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_bf16_w(tmem_base + kAccCols, a_lo + k * (2048 >> 4), hi, b_lo + (kXSlot >> 4) + k * (2048 >> 4), hi, idesc,
                          acc | (uint32_t)k);
          }
          }
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(acc_full);
      __syncwarp();
      ++n;
    }
  } else {
    // ===================== flush (4 warps): lane = co, column = kw * 64 + ci =====================
    const int q = warp & 3;
    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    SegWalk sw(p, blockIdx.x, gridDim.x);
    uint32_t n = 0;
    while (sw.next(p)) {
      mbar_wait(acc_full, n & 1u);
      tc_fence_after();
      const int co = sw.cot * 128 + q * 32 + lane;
      for (int j = 0; j < sw.nci; ++j) {
        const int ci0 = (sw.cic0 + j) * 64;
#pragma unroll 1
        for (int kw = 0; kw < 3; ++kw) {
          float* dst_row = p.dw + ((size_t)(sw.kh * 3 + kw) * p.cout + (co < p.cout ? co : 0)) * p.cin_pitch + ci0;
#pragma unroll 1
          for (int c = 0; c < 64; c += 16) {
            float v[16];
            tmem_ld16(t_addr + j * kAccCols + kw * 64 + c, v);
            if (co < p.cout && !(p.ko & 1)) {
#pragma unroll
              for (int i = 0; i < 16; i += 4)
                if (ci0 + c + i + 3 < p.cin_pitch) red_add_v4(dst_row + c + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      ++n;
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

bool conv3x3_wgrad_flatk_ok(const ActView& dy, const ActView& x) {
  static const int enabled = getenv("MIMO_WGRAD_FLATK") ? atoi(getenv("MIMO_WGRAD_FLATK")) : 1;
  if (!enabled) return false;
  if (dy.pad != 2 || x.pad != 1) return false;
  return (long long)x.N * x.hb() * x.wb() < (1ll << 31) - 4096;
}

// dy: zero-tail view (pad == 2), x: haloed view (pad == 1), both with c_off % 8 == 0; cin_pitch % 4 == 0
int conv3x3_wgrad_flatk_launch(const ActView& dy, const ActView& x, float* dw, int cin_pitch, cudaStream_t stream, bool pre_zeroed) {
  note_kernel(5);
  WgFlatKParams p{};
  p.wb = x.wb();
  const long long total_pos = (long long)x.N * x.hb() * x.wb();
  plan_schedule(p, dy.C, x.C, total_pos);
  p.cout = dy.C; p.cin_pitch = cin_pitch;
  p.dw = dw;

  CUtensorMap tm_dy, tm_x, tm_x2;
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
    uint32_t box2[2] = {64, 2};
    rc = encode_tmap_bf16(&tm_x2, x.base + x.c_off, 2, dims, strides, box2, 1);
    if (rc) return rc;
  }
  p.ko = getenv("MIMO_WGK_KO") ? atoi(getenv("MIMO_WGK_KO")) : 0;
  if (!(p.ko & 2) && !pre_zeroed) MIMO_CUDA(cudaMemsetAsync(dw, 0, (size_t)9 * p.cout * cin_pitch * sizeof(float), stream));
  const size_t smem_bytes = (size_t)kStages * kStageBytes + (2 * kStages + 2) * 8 + 16 + 1024;
  MIMO_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_flatk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));   // per launch: the attribute is per DEVICE, a process-wide "done" flag would skip the other GPUs
  const int grid = plan_grid(p, num_sms());
  conv3x3_wgrad_flatk_kernel<<<grid, kThreads, smem_bytes, stream>>>(tm_dy, tm_x, tm_x2, p);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

// Host-side dump of the stream-K schedule (pure arithmetic, no device): segment `j` of CTA `cta` as
// {co tile, first ci chunk, ci chunks, kh, first k-block, end k-block}; returns the number of segments of that CTA written
// (at most max_segs), or the grid size when out == nullptr.
int conv3x3_wgrad_flatk_schedule(int cout, int cin, long long total_pos, int sms, int cta, int* out, int max_segs) {
  WgFlatKParams p{};
  plan_schedule(p, cout, cin, total_pos);
  const int grid = plan_grid(p, sms);
  if (out == nullptr) return grid;
  if (cta < 0 || cta >= grid) return 0;
  SegWalk sw(p, (unsigned)cta, (unsigned)grid);
  int n = 0;
  while (sw.next(p)) {
    if (n < max_segs) {
      int* o = out + 6 * n;
      o[0] = sw.cot; o[1] = sw.cic0; o[2] = sw.nci; o[3] = sw.kh; o[4] = sw.kb0; o[5] = sw.kb1;
    }
    ++n;
  }
  return n;
}

}  // namespace mimo
