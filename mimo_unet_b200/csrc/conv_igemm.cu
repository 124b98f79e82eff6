// 3x3 convolution as an implicit GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), fed by TMA.
//
// Replaces: nn.Conv2d(k=3, padding=1, padding_mode="reflect") forward (reference components.py:23,26)
// and its input-gradient (cuDNN dgrad in the reference's autograd graph).
//
//   GEMM view   D[pixel, cout] = sum_{tap, cin} A[pixel + tap, cin] * Wp[tap][cout][cin]
//   A operand   128 output pixels x 64 input channels per k-block, loaded by ONE 4-D TMA box
//               (64 ch, tw, th, tn) straight out of the NHWC activation buffer. The tap shift is just a
//               different box origin, so there is no im2col buffer. Reflect padding is materialised by the
//               producer of the activation (1-pixel halo), so fprop is a "valid" conv over the padded
//               buffer; dgrad runs the same kernel over the unpadded dY with origin offset -2 and relies on
//               TMA's zero fill for out-of-bounds (negative) coordinates.
//   B operand   BLOCK_N couts x 64 cins of the packed bf16 weights for that tap (3-D TMA box).
//   Accumulator fp32 in TMEM, double buffered (2 x BLOCK_N columns) so the epilogue of tile i overlaps the
//               MMAs of tile i+1. Persistent CTAs, static round-robin tile schedule.
//   Epilogue    tcgen05.ld -> bf16 -> smem staging -> coalesced 16-byte global stores, plus (fprop, training)
//               per-channel sum / sum-of-squares partials of the *stored* bf16 values for BatchNorm.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane),
// warps 2..5 = epilogue (TMEM lane quarter = warp_idx % 4).
#include "common.cuh"
#include "conv_epilogue.cuh"
#include "ops.h"

#include <stdlib.h>

namespace mimo {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                      // bf16 elements = one 128-byte swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;

struct ConvParams {
  int n_img, out_h, out_w;    // output domain
  int tw, th, tn;             // pixel tile (tw*th*tn == 128), powers of two
  int tw_shift, th_shift;
  int tiles_w, tiles_h, tiles_n, m_tiles, n_tiles;
  int block_n;                // multiple of 16, <= 256
  int cin_chunks;             // ceil(Cin / 64)
  int origin;                 // 0 (fprop over padded input) or -2 (dgrad over unpadded dY)
  int cout;                   // true number of output channels
  int out_cpitch;             // channels per output pixel in memory (multiple of 8, >= cout)
  int stages;
  int b_stage_bytes;
  int ko;                     // diagnostic knock-outs (env MIMO_KO): 1 no global stores, 2 no A loads, 4 no MMAs
  EpiArgs epi;                // output tensor, BatchNorm statistics rows, optional fused bias / ReLU
};

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_igemm_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_w,
                     const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x A][stages x B][staging][barriers]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + (size_t)p.stages * kABytes;
  uint8_t* smem_epi = smem_b + (size_t)p.stages * p.b_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + ((epi_smem_bytes(p.block_n, p.epi.out_cpitch) + 15) & ~size_t(15)));
  uint64_t* full_bar = bars;                     // [stages]
  uint64_t* empty_bar = bars + kMaxStages;       // [stages]
  uint64_t* tmem_full = bars + 2 * kMaxStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles;
  const int k_blocks = 9 * p.cin_chunks;
  const uint32_t tmem_cols = (2 * p.block_n <= 32) ? 32 : (2 * p.block_n <= 64) ? 64 : (2 * p.block_n <= 128) ? 128
                             : (2 * p.block_n <= 256) ? 256 : 512;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_in);
    prefetch_tmap(&tmap_w);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tmem_full[a], 1);
        mbar_init(&tmem_empty[a], 4);  // one arrive per epilogue warp
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int mt = t / p.n_tiles, nt = t - mt * p.n_tiles;
      const int twi = mt % p.tiles_w;
      const int thi = (mt / p.tiles_w) % p.tiles_h;
      const int tni = mt / (p.tiles_w * p.tiles_h);
      const int w0 = twi * p.tw + p.origin, h0 = thi * p.th + p.origin, n0 = tni * p.tn;
      const int co0 = nt * p.block_n;
      int kh = 0, kw = 0, cc = 0;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[stage], ((p.ko & 2) ? 0 : kABytes) + p.b_stage_bytes);
          if (!(p.ko & 2)) tma_load_4d(&tmap_in, &full_bar[stage], smem_a + (size_t)stage * kABytes, cc * kBlockK, w0 + kw, h0 + kh, n0);
          tma_load_3d(&tmap_w, &full_bar[stage], smem_b + (size_t)stage * p.b_stage_bytes, cc * kBlockK, co0, kh * 3 + kw);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
        if (++cc == p.cin_chunks) { cc = 0; if (++kw == 3) { kw = 0; ++kh; } }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The loop is warp-uniform and only the tcgen05 instructions sit under elect_one(); descriptor words are
    // precomputed (high word constant, low word advanced by 32-bit adds), so issuing one MMA costs a handful of
    // uniform-datapath instructions instead of a 64-bit descriptor build + uniformisation loop.
    const uint32_t idesc = make_idesc_bf16(kBlockM, p.block_n, 0, 0);
    constexpr uint32_t hi = desc_hi(1024, kLayoutSW128);
    const uint32_t a_lo0 = desc_lo(smem_u32(smem_a), 16);
    const uint32_t b_lo0 = desc_lo(smem_u32(smem_b), 16);
    const uint32_t b_step = (uint32_t)p.b_stage_bytes >> 4;
    const bool no_mma = (p.ko & 4) != 0;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t it = 0;  // tile counter of this CTA: accumulator = it & 1, its use count = it >> 1
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const uint32_t acc = it & 1u;
      mbar_wait(&tmem_empty[acc], ((it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_addr = tmem_base + acc * (uint32_t)p.block_n;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + (uint32_t)stage * (kABytes >> 4);
        const uint32_t b_lo = b_lo0 + (uint32_t)stage * b_step;
        if (elect_one()) {
          if (!no_mma) {
            umma_bf16_w(d_addr, a_lo, hi, b_lo, hi, idesc, kb != 0);
            umma_bf16_w(d_addr, a_lo + 2, hi, b_lo + 2, hi, idesc, 1);
            umma_bf16_w(d_addr, a_lo + 4, hi, b_lo + 4, hi, idesc, 1);
            umma_bf16_w(d_addr, a_lo + 6, hi, b_lo + 6, hi, idesc, 1);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
      __syncwarp();
    }
  } else {
    // ===================== epilogue (4 warps, 128 threads) =====================
    const int q = warp & 3;                  // TMEM lane quarter owned by this warp
    const int row = q * 32 + lane;           // accumulator row == pixel index inside the tile
    const int et = threadIdx.x - 64;         // 0..127
    const EpiSmem es = epi_carve(smem_epi, p.block_n, p.epi.out_cpitch);
    epi_init(p.epi, es, et);
    uint32_t it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const uint32_t acc = it & 1u;
      const int mt = t / p.n_tiles, nt = t - mt * p.n_tiles;
      const int twi = mt % p.tiles_w;
      const int thi = (mt / p.tiles_w) % p.tiles_h;
      const int tni = mt / (p.tiles_w * p.tiles_h);
      // my row's pixel
      const int pw = twi * p.tw + (row & (p.tw - 1));
      const int ph = thi * p.th + ((row >> p.tw_shift) & (p.th - 1));
      const int pn = tni * p.tn + (row >> (p.tw_shift + p.th_shift));
      const bool valid = (pw < p.out_w) && (ph < p.out_h) && (pn < p.n_img);
      const int my_pix = (valid && !(p.ko & 1)) ? (pn * p.out_h + ph) * p.out_w + pw : -1;
      mbar_wait(&tmem_full[acc], (it >> 1) & 1u);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)p.block_n;
      epi_tile(p.epi, es, t_addr, my_pix, nt * p.block_n, &tmem_empty[acc], q, lane, et);
    }
    epi_finish(p.epi, es, et);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

int pick_tile(int W, int H, int N, int* tw_o, int* th_o, int* tn_o) {
  double best = 1e30;
  for (int tw = 1; tw <= 128; tw <<= 1) {
    for (int th = 1; tw * th <= 128; th <<= 1) {
      const int tn = 128 / (tw * th);
      const double tiles = (double)ceil_div(W, tw) * ceil_div(H, th) * ceil_div(N, tn);
      const double cost = tiles * (1.0 + 0.5 / tw);
      if (cost < best - 1e-9) {
        best = cost;
        *tw_o = tw; *th_o = th; *tn_o = tn;
      }
    }
  }
  return 0;
}

int ilog2(int v) { int s = 0; while ((1 << s) < v) ++s; return s; }

}  // namespace

int conv3x3_block_n(int cout) {
  const int c16 = round_up(cout, 16);
  if (c16 <= 256) return c16;
  const int parts = ceil_div(c16, 256);
  return round_up(ceil_div(c16, parts), 16);
}

// rows of the BatchNorm partial-statistics buffers: one per persistent CTA
int conv3x3_stat_rows() { return num_sms(); }

// in      : activation view. mode 0 (fprop): pad == 1, output domain = in.H x in.W
//                            mode 1 (dgrad): pad == 0, output domain = (in.H+2) x (in.W+2)
// wpacked : bf16 [9][cout][cin_pitch] (cin contiguous), cin_pitch multiple of 8
// out     : bf16 [N][out_h][out_w][out_cpitch]
// Small-channel layers that the CTA-pair kernel serves better than conv3x3_flat_kernel (measured, tools/gpu/r2_c2flat.sh): with a
// narrow row pitch four tiles per CTA share one segment halo (64 x 80 maps: 37 -> 30 us, 35 -> 29 us), and with more than 32
// input or output channels the pair MMA (39 cycles for 256 rows at small N) beats the 43-cycle fixed cost per single-CTA MMA
// (63 -> 31 @ 128 x 160: 94 -> 83 us, its dgrad 117 -> 97 us). Wide full-resolution maps with <= 32 channels stay on the flat
// kernel: there the 2 (W + 2) + 2 halo rows per segment cost more than the instruction shape gains (21 -> 21: 79 vs 85 us).
static bool prefer_c2(const ActView& in, int mode, int cout) {
  static const int enabled = getenv("MIMO_C2_SMALL") ? atoi(getenv("MIMO_C2_SMALL")) : 1;
  if (!enabled || !conv3x3_c2_ok(in, mode, cout)) return false;
  return 2 * in.wb() + 2 <= 256 || in.C > 32 || cout > 32;
}

bool conv3x3_fuse_ok(const ActView& in, int cout) {
  if (conv3x3_flat2_ok(in, 0, cout)) return false;   // (the opt-in experiment has no fused epilogue)
  return conv3x3_thin_ok(in, 0, cout) || conv3x3_flat_ok(in, 0, cout) || conv3x3_c2_ok(in, 0, cout);
}

int conv3x3_launch(const ActView& in, int mode, const bf16* wpacked, int cout, int cin_pitch, bf16* out, int out_cpitch,
                   float* stat_sum, float* stat_sq, const float* bias, int relu, cudaStream_t stream, const ConvFuse* fuse) {
  MIMO_CHECK(mode == 0 || mode == 1, MIMO_ERR_ARG, "conv3x3: bad mode %d", mode);
  MIMO_CHECK(mode == 0 ? in.pad == 1 : (in.pad == 0 || in.pad == 2), MIMO_ERR_ARG, "conv3x3: mode %d got a pad=%d input", mode, in.pad);
  MIMO_CHECK(in.cpitch % 8 == 0 && in.c_off % 8 == 0, MIMO_ERR_ALIGN, "conv3x3: input cpitch/c_off must be multiples of 8 (got %d/%d)", in.cpitch, in.c_off);
  MIMO_CHECK(cin_pitch % 8 == 0 && cin_pitch >= in.C, MIMO_ERR_ALIGN, "conv3x3: weight cin pitch %d invalid for C=%d", cin_pitch, in.C);
  MIMO_CHECK(out_cpitch % 8 == 0 && out_cpitch >= cout, MIMO_ERR_ALIGN, "conv3x3: out_cpitch %d invalid for cout=%d", out_cpitch, cout);
  MIMO_CHECK(((uintptr_t)in.base % 16) == 0 && ((uintptr_t)wpacked % 16) == 0 && ((uintptr_t)out % 16) == 0, MIMO_ERR_ALIGN,
             "conv3x3: pointers must be 16-byte aligned");
  MIMO_CHECK(in.H >= 2 && in.W >= 2, MIMO_ERR_ARG, "conv3x3: reflect padding needs H,W >= 2");
  if (conv3x3_thin_ok(in, mode, cout))
    return conv3x3_thin_launch(in, wpacked, cout, cin_pitch, out, out_cpitch, stat_sum, stat_sq, bias, relu, stream, fuse);
  if (fuse != nullptr) {
    if (conv3x3_flat_ok(in, mode, cout) && !conv3x3_flat2_ok(in, mode, cout) && !prefer_c2(in, mode, cout))
      return conv3x3_flat_launch(in, mode, wpacked, cout, cin_pitch, out, out_cpitch, stat_sum, stat_sq, bias, relu, stream, fuse);
    if (conv3x3_c2_ok(in, mode, cout))
      return conv3x3_c2_launch(in, mode, wpacked, cout, cin_pitch, out, out_cpitch, stat_sum, stat_sq, bias, relu, stream, fuse);
    set_error("conv3x3: no kernel with a fused inference epilogue for this layer (check conv3x3_fuse_ok first)");
    return MIMO_ERR_ARG;
  }
  if (conv3x3_flat2_ok(in, mode, cout))
    return conv3x3_flat2_launch(in, mode, wpacked, cout, cin_pitch, out, out_cpitch, stat_sum, stat_sq, bias, relu, stream);
  if (conv3x3_flat_ok(in, mode, cout) && !prefer_c2(in, mode, cout))
    return conv3x3_flat_launch(in, mode, wpacked, cout, cin_pitch, out, out_cpitch, stat_sum, stat_sq, bias, relu, stream);
  if (conv3x3_c2_ok(in, mode, cout))
    return conv3x3_c2_launch(in, mode, wpacked, cout, cin_pitch, out, out_cpitch, stat_sum, stat_sq, bias, relu, stream);
  if (conv3x3_flatk_ok(in, mode, cout))
    return conv3x3_flatk_launch(in, mode, wpacked, cout, cin_pitch, out, out_cpitch, stat_sum, stat_sq, bias, relu, stream);

  note_kernel(3);
  ConvParams p{};
  p.n_img = in.N;
  p.out_h = in.H + (mode == 1 ? 2 : 0);
  p.out_w = in.W + (mode == 1 ? 2 : 0);
  pick_tile(p.out_w, p.out_h, p.n_img, &p.tw, &p.th, &p.tn);
  p.tw_shift = ilog2(p.tw);
  p.th_shift = ilog2(p.th);
  p.tiles_w = ceil_div(p.out_w, p.tw);
  p.tiles_h = ceil_div(p.out_h, p.th);
  p.tiles_n = ceil_div(p.n_img, p.tn);
  p.m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  p.block_n = conv3x3_block_n(cout);
  p.n_tiles = ceil_div(round_up(cout, 16), p.block_n);
  p.cin_chunks = ceil_div(in.C, kBlockK);
  p.origin = (mode == 1) ? -2 : 0;
  p.b_stage_bytes = p.block_n * kBlockK * 2;
  {
    static const int ko = getenv("MIMO_KO") ? atoi(getenv("MIMO_KO")) : 0;
    p.ko = ko;
    if (ko & 8) { stat_sum = nullptr; stat_sq = nullptr; }
  }
  p.epi.block_n = p.block_n;
  p.epi.cout = cout;
  p.epi.out_cpitch = out_cpitch;
  p.epi.stage_pitch = p.block_n * 2 + 16;
  p.epi.stat_rows = conv3x3_stat_rows();
  p.epi.out = out;
  p.epi.stat_sum = stat_sum;
  p.epi.stat_sq = stat_sq;
  p.epi.bias = bias;
  p.epi.relu = relu;
  // the last n-tile may own fewer than block_n channels in memory; it must still be a multiple of 8
  MIMO_CHECK((p.n_tiles - 1) * p.block_n < out_cpitch, MIMO_ERR_ARG, "conv3x3: n-tiling exceeds out_cpitch");

  const int smem_budget = 227 * 1024 - 1024 /*align slack*/;
  const int fixed = (int)((epi_smem_bytes(p.block_n, out_cpitch) + 15) & ~size_t(15)) + (2 * kMaxStages + 4) * 8 + 16 + 64;
  int stages = (smem_budget - fixed) / (kABytes + p.b_stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  MIMO_CHECK(stages >= 2, MIMO_ERR_ARG, "conv3x3: not enough shared memory for block_n=%d", p.block_n);
  p.stages = stages;
  const size_t smem_bytes = (size_t)stages * (kABytes + p.b_stage_bytes) + fixed + 1024;

  // --- tensor maps ---
  CUtensorMap tm_in, tm_w;
  {
    // fprop reads the whole haloed buffer; dgrad reads the H x W interior (zero fill outside) whatever the buffer pitch is
    const int Hb = in.hb(), Wb = in.wb();
    uint64_t dims[4] = {(uint64_t)in.C, (uint64_t)(mode == 0 ? Wb : in.W), (uint64_t)(mode == 0 ? Hb : in.H), (uint64_t)in.N};
    uint64_t strides[3] = {(uint64_t)in.cpitch * 2, (uint64_t)Wb * in.cpitch * 2, (uint64_t)Hb * Wb * in.cpitch * 2};
    uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tn};
    int rc = encode_tmap_bf16(&tm_in, in.base + in.c_off, 4, dims, strides, box, 1);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)cin_pitch, (uint64_t)cout, 9};
    uint64_t strides[2] = {(uint64_t)cin_pitch * 2, (uint64_t)cout * cin_pitch * 2};
    uint32_t box[3] = {(uint32_t)kBlockK, (uint32_t)p.block_n, 1};
    int rc = encode_tmap_bf16(&tm_w, wpacked, 3, dims, strides, box, 1);
    if (rc) return rc;
  }

  MIMO_CUDA(cudaFuncSetAttribute(conv3x3_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));   // per launch: the attribute is per DEVICE, a process-wide "done" flag would skip the other GPUs
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  conv3x3_igemm_kernel<<<grid, kThreads, smem_bytes, stream>>>(tm_in, tm_w, p);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

}  // namespace mimo
