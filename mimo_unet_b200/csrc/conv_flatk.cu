// "Flat, K-chunked" tcgen05 implicit-GEMM 3x3 convolution for the mid layers (up to 256 input channels, <= 192 output
// channels, large feature maps: core.up3 / core.up2 / core.down2 of the reference U-Net, model.py:190-243): the layers
// where the 4-D kernel (conv_igemm.cu) is bound by the TMA row rate because it fetches the A tile once per TAP, and
// where the single-chunk ring kernel (conv_flat.cu) does not apply.
//
// Same flattened addressing as conv_flat.cu (the [N][H+2][W+2][C] buffer is one list of pixel rows, a tap is a row
// offset), but K is streamed in (64-channel chunk, kernel row kh) blocks and the work item is a PAIR of tiles:
//   * A: one contiguous segment of 258 rows per (chunk, kh) serves both 128-position tiles of the pair and all three kw
//     taps (row-shifted UMMA windows; SWIZZLE_128B is absolute-address based on B200, see conv_flat.cu);
//   * B: the three kw taps of the chunk for that kh (3 x block_n rows), shared by the two tiles;
//   -> (258 + 3*block_n) TMA rows per 256 positions x 64 channels x 3 taps instead of 3 x 2 x (128 + block_n).
//   * four TMEM accumulators (2 tiles x double buffering), epilogue without shared-memory staging or block barriers:
//     16-byte stores straight from registers, BatchNorm statistics reduced with the transposed warp butterfly and kept in
//     registers across all items of the CTA.
// mode 0 (fprop): in = pad==1 view, output dense [N][H][W]; mode 1 (dgrad): in = pad==2 (zero tail) view of dY, output =
// padded-domain gradient [N][H+2][W+2] (see conv_flat.cu).
#include "common.cuh"
#include "conv_epilogue.cuh"
#include "ops.h"

#include <stdlib.h>

namespace mimo {
namespace {

constexpr int kBlockM = 128;
constexpr int kPairRows = 2 * kBlockM + 2;       // 258 rows: two tiles + the kw = 1, 2 shifts
constexpr int kABytes = 33 * 1024;               // 256 rows x 128 B + a 2-row box at +32 KB, rounded to the swizzle repeat
constexpr int kThreads = 192;
constexpr int kMaxStages = 6;
constexpr int kMaxChunks16 = 12;                 // 16-column accumulator chunks over all n-tiles (cout <= 192)

struct FlatKParams {
  int wb, img_pix;
  long long total_pos;
  int origin;
  int out_h, out_w, n_img;
  int n_pairs, n_tiles_n, block_n;
  int cin_chunks, ks_last;     // 64-channel chunks; 16-channel k-steps that carry data in the last chunk
  int stages, stage_bytes;
  EpiArgs epi;
};

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_flatk_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a2,
                     const __grid_constant__ CUtensorMap tmap_w, const FlatKParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* smem_epi = reinterpret_cast<float*>(smem + (size_t)p.stages * p.stage_bytes);  // [2][4][16 * kMaxChunks16]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + 8 * 16 * kMaxChunks16);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tmem_full = bars + 2 * kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_items = p.n_pairs * p.n_tiles_n;
  const int k_blocks = 3 * p.cin_chunks;
  const uint32_t cols = 4u * (uint32_t)p.block_n;   // 2 tiles x 2 buffers
  const uint32_t tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_a2);
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
        mbar_init(&tmem_empty[a], 4);
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
    const uint32_t tx = (uint32_t)(kPairRows * 128 + 3 * p.block_n * 128);
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int pair = it / p.n_tiles_n, nt = it - pair * p.n_tiles_n;
      const long long p0 = (long long)pair * (2 * kBlockM) + p.origin;
      const int co0 = nt * p.block_n;
      for (int cc = 0; cc < p.cin_chunks; ++cc) {
        for (int kh = 0; kh < 3; ++kh) {
          uint8_t* st = smem + (size_t)stage * p.stage_bytes;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            const int row = (int)(p0 + (long long)kh * p.wb);
            mbar_arrive_expect_tx(&full_bar[stage], tx);
            tma_load_2d(&tmap_a, &full_bar[stage], st, cc * 64, row);
            tma_load_2d(&tmap_a2, &full_bar[stage], st + 256 * 128, cc * 64, row + 256);
            tma_load_3d(&tmap_w, &full_bar[stage], st + kABytes, cc * 64, co0, kh * 3);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, precomputed descriptor words) =====================
    const uint32_t idesc = make_idesc_bf16(kBlockM, p.block_n, 0, 0);
    constexpr uint32_t hi = desc_hi(1024, kLayoutSW128);
    const uint32_t lo0 = desc_lo(smem_u32(smem), 16);
    const uint32_t stage_step = (uint32_t)p.stage_bytes >> 4;
    const uint32_t b_tap = (uint32_t)(p.block_n * 128) >> 4;
    const uint32_t bn = (uint32_t)p.block_n;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t n = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
      const uint32_t buf = n & 1u;
      mbar_wait(&tmem_empty[buf], ((n >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d0 = tmem_base + buf * 2u * bn;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_lo = lo0 + (uint32_t)stage * stage_step;
        const uint32_t b_lo = a_lo + (kABytes >> 4);
        const uint32_t ks = (kb >= k_blocks - 3) ? (uint32_t)p.ks_last : 4u;   // the last chunk's three kh blocks
        if (elect_one()) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if ((uint32_t)k < ks)
                  umma_bf16_w(d0 + (uint32_t)half * bn, a_lo + (((half * kBlockM + kw) * 128 + k * 32) >> 4), hi,
                              b_lo + (uint32_t)kw * b_tap + ((k * 32) >> 4), hi, idesc, (kb | kw | k) != 0);
              }
            }
          }
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&tmem_full[buf]);
      __syncwarp();
    }
  } else {
    // ===================== epilogue (4 warps) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const EpiArgs& e = p.epi;
    const bool stats = e.stat_sum != nullptr;
    const int nchunks = p.block_n >> 4;
    float acc_s[kMaxChunks16], acc_q[kMaxChunks16];   // column (lane >> 1) of 16-column chunk j over all n-tiles
#pragma unroll
    for (int j = 0; j < kMaxChunks16; ++j) { acc_s[j] = 0.f; acc_q[j] = 0.f; }
    uint32_t n = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
      const uint32_t buf = n & 1u;
      const int pair = it / p.n_tiles_n, nt = it - pair * p.n_tiles_n;
      const int co0 = nt * p.block_n;
      mbar_wait(&tmem_full[buf], (n >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const long long pos = (long long)pair * (2 * kBlockM) + half * kBlockM + row;
        bool valid = false;
        size_t my_pix = 0;
        if (pos < p.total_pos) {
          const unsigned up = (unsigned)pos;
          const unsigned ni = up / (unsigned)p.img_pix;
          const unsigned rem = up - ni * (unsigned)p.img_pix;
          const unsigned hp = rem / (unsigned)p.wb, wp = rem - hp * (unsigned)p.wb;
          valid = (int)hp < p.out_h && (int)wp < p.out_w;
          my_pix = ((size_t)ni * p.out_h + hp) * p.out_w + wp;
        }
        const float vmask = valid ? 1.f : 0.f;
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 2u * (uint32_t)p.block_n + (uint32_t)(half * p.block_n);
        bf16* dst = e.out + my_pix * e.out_cpitch + co0;
#pragma unroll
        for (int j = 0; j < kMaxChunks16; ++j) {
          if (j < nchunks) {
            float v[16];
            tmem_ld16(t_addr + j * 16, v);
            if (e.bias != nullptr) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += (co0 + j * 16 + i < e.cout) ? __ldg(e.bias + co0 + j * 16 + i) : 0.f;
            }
            if (e.relu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            const uint4 lo = pack8(v), hi8 = pack8(v + 8);
            if (valid) {
              if (co0 + j * 16 < e.out_cpitch) *reinterpret_cast<uint4*>(dst + j * 16) = lo;
              if (co0 + j * 16 + 8 < e.out_cpitch) *reinterpret_cast<uint4*>(dst + j * 16 + 8) = hi8;
            }
            if (stats) {
              float f[16];
              unpack8(lo, f);
              unpack8(hi8, f + 8);
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] *= vmask;   // statistics of the values as stored; unstored rows count as 0
              const float cs = warp_colsum16(f, lane);
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] *= f[i];
              const float cq = warp_colsum16(f, lane);
              // n-tile nt owns accumulator slots [nt * nchunks, (nt + 1) * nchunks): static indexing over the product
#pragma unroll
              for (int t2 = 0; t2 < kMaxChunks16; ++t2)
                if (t2 == nt * nchunks + j) { acc_s[t2] += cs; acc_q[t2] += cq; }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    }
    if (stats) {
      // combine the four warps (fixed order) and write this CTA's partial row
      const int total_chunks = nchunks * p.n_tiles_n;
#pragma unroll
      for (int j = 0; j < kMaxChunks16; ++j) {
        if (j < total_chunks) {
          if ((lane & 1) == 0) smem_epi[(q * kMaxChunks16 + j) * 16 + (lane >> 1)] = acc_s[j];
          else smem_epi[((4 + q) * kMaxChunks16 + j) * 16 + (lane >> 1)] = acc_q[j];
        }
      }
      named_bar_sync(1, 128);
      for (int col = et; col < e.out_cpitch; col += 128) {
        // column col of the output = n-tile col / block_n, chunk (col % block_n) / 16
        const int nt = col / p.block_n, cc = col - nt * p.block_n;
        const int j = nt * nchunks + (cc >> 4), i = cc & 15;
        float s_ = 0.f, q_ = 0.f;
        if (j < total_chunks) {
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            s_ += smem_epi[(w * kMaxChunks16 + j) * 16 + i];
            q_ += smem_epi[((4 + w) * kMaxChunks16 + j) * 16 + i];
          }
        }
        e.stat_sum[(size_t)blockIdx.x * e.out_cpitch + col] = s_;
        e.stat_sq[(size_t)blockIdx.x * e.out_cpitch + col] = q_;
        for (int rr = blockIdx.x + gridDim.x; rr < e.stat_rows; rr += gridDim.x) {
          e.stat_sum[(size_t)rr * e.out_cpitch + col] = 0.f;
          e.stat_sq[(size_t)rr * e.out_cpitch + col] = 0.f;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

void plan_n(int cout, int* block_n, int* n_tiles) {
  const int c16 = round_up(cout, 16);
  *n_tiles = ceil_div(c16, 128);
  *block_n = round_up(ceil_div(c16, *n_tiles), 16);
}

}  // namespace

bool conv3x3_flatk_ok(const ActView& in, int mode, int cout) {
  static const int enabled = env_int("MIMO_CONV_FLATK", 1);
  const int min_items = env_int("MIMO_FLATK_MIN_ITEMS", 2 * 148);  // (read per call: tests force the path with 1)
  if (!enabled) return false;
  if (in.C > 256 || cout > 192) return false;   // (layers with <= 64 in AND <= 64 out channels took the ring kernel)
  if (mode == 0 && in.pad != 1) return false;
  if (mode == 1 && in.pad != 2) return false;
  const long long total_pos = (long long)in.N * in.hb() * in.wb();
  if (total_pos >= (1ll << 31) - 4096) return false;
  int block_n, n_tiles;
  plan_n(cout, &block_n, &n_tiles);
  if (n_tiles * (block_n >> 4) > kMaxChunks16) return false;
  return ceil_div_ll(total_pos, 2 * kBlockM) * n_tiles >= min_items;   // enough pairs to keep the persistent grid busy
}

int conv3x3_flatk_launch(const ActView& in, int mode, const bf16* wpacked, int cout, int cin_pitch, bf16* out, int out_cpitch,
                         float* stat_sum, float* stat_sq, const float* bias, int relu, cudaStream_t stream) {
  note_kernel(2);
  FlatKParams p{};
  p.wb = in.wb();
  p.img_pix = in.hb() * in.wb();
  p.total_pos = (long long)in.N * p.img_pix;
  p.origin = mode == 0 ? 0 : -(2 * p.wb + 2);
  p.out_h = mode == 0 ? in.H : in.H + 2;
  p.out_w = mode == 0 ? in.W : in.W + 2;
  p.n_img = in.N;
  p.n_pairs = (int)ceil_div_ll(p.total_pos, 2 * kBlockM);
  plan_n(cout, &p.block_n, &p.n_tiles_n);
  p.cin_chunks = ceil_div(in.C, 64);
  p.ks_last = ceil_div(in.C - (p.cin_chunks - 1) * 64, 16);
  p.stage_bytes = kABytes + ((3 * p.block_n * 128 + 1023) & ~1023);
  const int fixed = 8 * 16 * kMaxChunks16 * 4 + (2 * kMaxStages + 4) * 8 + 16 + 64 + 1024;
  int stages = (227 * 1024 - fixed) / p.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  MIMO_CHECK(stages >= 2, MIMO_ERR_ARG, "conv3x3_flatk: not enough shared memory for block_n=%d", p.block_n);
  p.stages = stages;
  const size_t smem_bytes = (size_t)stages * p.stage_bytes + fixed;
  p.epi.block_n = p.block_n;
  p.epi.cout = cout;
  p.epi.out_cpitch = out_cpitch;
  p.epi.stage_pitch = 0;
  p.epi.stat_rows = conv3x3_stat_rows();
  p.epi.out = out;
  p.epi.stat_sum = stat_sum;
  p.epi.stat_sq = stat_sq;
  p.epi.bias = bias;
  p.epi.relu = relu;
  MIMO_CHECK((p.n_tiles_n - 1) * p.block_n < out_cpitch, MIMO_ERR_ARG, "conv3x3_flatk: n-tiling exceeds out_cpitch");

  CUtensorMap tm_a, tm_a2, tm_w;
  {
    uint64_t dims[2] = {(uint64_t)in.C, (uint64_t)p.total_pos};
    uint64_t strides[1] = {(uint64_t)in.cpitch * 2};
    uint32_t box[2] = {64, 256};
    int rc = encode_tmap_bf16(&tm_a, in.base + in.c_off, 2, dims, strides, box, 1);
    if (rc) return rc;
    uint32_t box2[2] = {64, 2};
    rc = encode_tmap_bf16(&tm_a2, in.base + in.c_off, 2, dims, strides, box2, 1);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)cin_pitch, (uint64_t)cout, 9};
    uint64_t strides[2] = {(uint64_t)cin_pitch * 2, (uint64_t)cout * cin_pitch * 2};
    uint32_t box[3] = {64, (uint32_t)p.block_n, 3};
    int rc = encode_tmap_bf16(&tm_w, wpacked, 3, dims, strides, box, 1);
    if (rc) return rc;
  }
  MIMO_CUDA(cudaFuncSetAttribute(conv3x3_flatk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));   // per launch: the attribute is per DEVICE, a process-wide "done" flag would skip the other GPUs
  const int items = p.n_pairs * p.n_tiles_n;
  const int grid = items < num_sms() ? items : num_sms();
  conv3x3_flatk_kernel<<<grid, kThreads, smem_bytes, stream>>>(tm_a, tm_a2, tm_w, p);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

}  // namespace mimo
