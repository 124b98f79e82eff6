// "Flat" variant of the tcgen05 implicit-GEMM 3x3 convolution for the HBM-bound layers (few channels, many pixels):
// the full-resolution encoder / decoder convolutions and the first encoder level (reference components.py:23,26;
// shapes in SURVEY App. A: 3->21, 21->21, 63->31, 31->21 @ HxW and 21->42, 42->42 @ H/2 x W/2).
//
// Idea: the input buffer [N][H+2][W+2][C] is ONE long list of pixels. For an output position P (flattened, top-left
// aligned with the buffer) the nine taps are the rows  P + kh*(W+2) + kw  of that list, so
//   * a tile is 128 CONSECUTIVE positions, whatever W is (no partial tiles at row ends; the 2 positions per row and
//     2 rows per image that fall on the halo are computed and simply not stored: 3 % waste at 128x160);
//   * the A operand of all nine taps comes from THREE TMA loads per tile (one 130-row segment per kernel row kh)
//     instead of nine 128-row boxes: the kw shift is a +128-byte (one pixel row) shift of the UMMA descriptor start
//     address inside the SWIZZLE_128B segment, declared through the descriptor's base-offset field;
//   * the packed weights of all nine taps (<= 72 KB) are loaded ONCE per persistent CTA and stay resident.
// mode 0 (fprop): in = pad==1 (reflect halo) view, output = dense [N][H][W] (+ BatchNorm statistics).
// mode 1 (dgrad): in = pad==2 (zero tail) view of dY; output = padded-domain gradient [N][H+2][W+2]; tap (a,b) reads
//                 position Q + (a-2)*(W+2) + (b-2): rows/columns "before" an image row are the zero tail of the
//                 previous row / image (TMA zero-fills negative coordinates), so EVERY position is a valid output.
#include "common.cuh"
#include "conv_epilogue.cuh"
#include "ops.h"

#include <stdlib.h>

namespace mimo {
namespace {

constexpr int kBlockM = 128;
constexpr int kSegRows = kBlockM + 2;             // positions P .. P+129 cover kw = 0..2
constexpr int kSegBytes = 17 * 1024;              // 130 rows x 128 B rounded up to the 1024-byte swizzle repeat
constexpr int kStageBytes = 3 * kSegBytes;        // kh = 0..2
constexpr int kThreads = 192;
constexpr int kMaxStages = 4;

struct FlatParams {
  int wb;              // buffer row pitch in pixels (W + 2)
  int img_pix;         // pixels per image in the buffer ((H+2)*(W+2))
  long long total_pos; // N * img_pix
  int origin;          // first tap row offset: 0 (fprop) or -(2*wb + 2) (dgrad)
  int out_h, out_w;    // stored output domain inside the (H+2)x(W+2) position grid
  int n_img;
  int m_tiles;
  int block_n;
  int k_steps;         // ceil(C / 16): 16-channel MMA k-steps that carry data
  int stages;
  int bo_mode;         // diagnostic: 1 sets the descriptor base-offset field to kw (measured WRONG on B200: the
                       // hardware swizzles on absolute smem address bits, so shifted windows need base offset 0)
  int pf_dist;         // L2 prefetch distance in tiles of this CTA (0 = off)
  int ko;              // diagnostic knock-outs: 1 no global stores, 2 no A loads, 4 no MMAs
  long long* trace;    // diagnostic (env MIMO_FLAT_TRACE): CTA 0 records clock64() of its pipeline events, [3 roles][64 tiles][4]
  EpiArgs epi;
};

// issue one 32-lane x 16-column TMEM load without waiting (pair with tmem_ld_wait)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_flat_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_w, const FlatParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [B: 9 x block_n x 128 B][stages x 3 segments][epilogue scratch][barriers]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem;
  const int b_bytes = 9 * BN * 128;                  // BN % 8 == 0 -> every tap starts 1024-byte aligned
  uint8_t* smem_a = smem_b + ((b_bytes + 1023) & ~1023);
  float* smem_epi = reinterpret_cast<float*>(smem_a + (size_t)p.stages * kStageBytes);  // [2][4][BN] end-of-kernel partials
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + 8 * BN);
  uint64_t* full_bar = bars;                    // [stages]
  uint64_t* empty_bar = bars + kMaxStages;      // [stages]
  uint64_t* tmem_full = bars + 2 * kMaxStages;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint64_t* b_full = tmem_empty + 2;            // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t tmem_cols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : 256;

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
        mbar_init(&tmem_empty[a], 4);
      }
      mbar_init(b_full, 1);
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
    {
      if (elect_one()) {
        mbar_arrive_expect_tx(b_full, (uint32_t)b_bytes);
        tma_load_3d(&tmap_w, b_full, smem_b, 0, 0, 0);  // box (64 cin, block_n cout, 9 taps)
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x, ++it) {
        const long long p0 = (long long)t * kBlockM + p.origin;
        uint8_t* st = smem_a + (size_t)stage * kStageBytes;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          if (p.trace && blockIdx.x == 0 && it < 64) p.trace[(0 * 64 + it) * 4 + 0] = clock64();
          if (p.ko & 2) {
            mbar_arrive(&full_bar[stage]);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], 3 * kSegRows * 128);
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) tma_load_2d(&tmap_in, &full_bar[stage], st + kh * kSegBytes, 0, (int)(p0 + (long long)kh * p.wb));
          }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // One thread, so every instruction counts: descriptor words are precomputed, per-MMA work is two 32-bit adds with
    // immediate offsets (tap / k-step loops fully unrolled), no local-memory state. The loop is warp-uniform; only the
    // tcgen05 instructions sit under elect_one().
    {
      const uint32_t idesc = make_idesc_bf16(kBlockM, BN, 0, 0);
      constexpr uint32_t hi = desc_hi(1024, kLayoutSW128);
      const uint32_t b_lo0 = desc_lo(smem_u32(smem_b), 16);
      const uint32_t a_lo0 = desc_lo(smem_u32(smem_a), 16);
      const uint32_t ks = (uint32_t)p.k_steps;
      const bool no_mma = (p.ko & 4) != 0;
      mbar_wait(b_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;  // tile counter of this CTA: accumulator = it & 1, its use count = it >> 1
      for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x, ++it) {
        const uint32_t acc = it & 1u;
        const bool tr = p.trace && blockIdx.x == 0 && it < 64 && lane == 0;
        long long* trow = p.trace + (1 * 64 + it) * 4;
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1u) ^ 1u);
        if (tr) trow[0] = clock64();
        mbar_wait(&full_bar[stage], phase);
        if (tr) trow[1] = clock64();
        tc_fence_after();
        const uint32_t d_addr = tmem_base + acc * BN;
        const uint32_t a_lo = a_lo0 + (uint32_t)stage * (kStageBytes >> 4);
        if (elect_one()) {
         if (!no_mma) {
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                // A: segment kh, shifted by kw positions (+128 B) inside the swizzled segment; k-step = +32 B
                if ((uint32_t)k < ks)
                  umma_bf16_w(d_addr, a_lo + ((kh * kSegBytes + kw * 128 + k * 32) >> 4), hi,
                              b_lo0 + (((kh * 3 + kw) * BN * 128 + k * 32) >> 4), hi, idesc, (kh | kw | k) != 0);
              }
            }
          }
         }
         umma_commit(&empty_bar[stage]);   // smem slot free once these MMAs have read it
         umma_commit(&tmem_full[acc]);     // accumulator complete -> epilogue
        }
        __syncwarp();
        if (tr) trow[3] = clock64();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (4 warps, 128 threads) =====================
    // Thread (q, lane) owns accumulator row q*32+lane of every tile = one output position. No shared-memory staging and
    // no block barrier per tile: the row is stored straight from registers (16-byte stores; consecutive positions are
    // consecutive pixels in memory), and the BatchNorm statistics are accumulated per THREAD in registers over all of
    // the CTA's tiles, then reduced across threads once at the end of the kernel (fixed order -> deterministic).
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const EpiArgs& e = p.epi;
    const bool stats = e.stat_sum != nullptr;
    float ssum[BN], ssq[BN];
#pragma unroll
    for (int i = 0; i < BN; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
    // position of my row in the first tile, then advanced incrementally (no divisions in the loop)
    int pn, ph, pw;
    {
      const long long pos0 = (long long)blockIdx.x * kBlockM + row;
      pn = (int)(pos0 / p.img_pix);
      const int rem = (int)(pos0 - (long long)pn * p.img_pix);
      ph = rem / p.wb;
      pw = rem - ph * p.wb;
    }
    const int hb = p.img_pix / p.wb;
    int dn, dh, dw;
    {
      const long long d = (long long)gridDim.x * kBlockM;
      dn = (int)(d / p.img_pix);
      const int rem = (int)(d - (long long)dn * p.img_pix);
      dh = rem / p.wb;
      dw = rem - dh * p.wb;
    }
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x, ++it) {
      const uint32_t acc = it & 1u;
      const bool valid = pn < p.n_img && ph < p.out_h && pw < p.out_w && !(p.ko & 1);
      const size_t my_pix = (size_t)(pn * p.out_h + ph) * p.out_w + pw;
      const bool tr = p.trace && blockIdx.x == 0 && et == 0 && it < 64;
      long long* trow = p.trace + (2 * 64 + it) * 4;
      if (tr) trow[0] = clock64();
      mbar_wait(&tmem_full[acc], (it >> 1) & 1u);
      if (tr) trow[1] = clock64();
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      uint32_t r[BN];
#pragma unroll
      for (int c = 0; c < BN; c += 16) tmem_ld16_nowait(t_addr + c, r + c);
      tmem_ld_wait();
      // TMEM drained -> hand the accumulator back to the MMA warp before doing the math / stores
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (tr) trow[2] = clock64();
      if (valid) {
        bf16* dst = e.out + my_pix * e.out_cpitch;
#pragma unroll
        for (int c = 0; c < BN; c += 8) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[c + i]);
          if (e.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += (c + i < e.cout) ? __ldg(e.bias + c + i) : 0.f;
          }
          if (e.relu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          bf16x8 o;
#pragma unroll
          for (int i = 0; i < 8; ++i) o.v[i] = __float2bfloat16_rn(v[i]);
          if (c < e.out_cpitch) *reinterpret_cast<bf16x8*>(dst + c) = o;
          if (stats) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float f = __bfloat162float(o.v[i]);  // statistics of the values as stored
              ssum[c + i] += f;
              ssq[c + i] = fmaf(f, f, ssq[c + i]);
            }
          }
        }
      }
      if (tr) trow[3] = clock64();
      // advance my position by gridDim.x tiles
      pw += dw;
      if (pw >= p.wb) { pw -= p.wb; ++ph; }
      ph += dh;
      if (ph >= hb) { ph -= hb; ++pn; }
      pn += dn;
    }
    if (stats) {
      // cross-thread reduction, once per CTA: butterfly inside each warp, then the four warps through shared memory
#pragma unroll
      for (int c = 0; c < BN; c += 16) {
        const float cs = warp_colsum16(ssum + c, lane);
        const float cq = warp_colsum16(ssq + c, lane);
        if ((lane & 1) == 0) smem_epi[q * BN + c + (lane >> 1)] = cs;
        else smem_epi[(4 + q) * BN + c + (lane >> 1)] = cq;
      }
      named_bar_sync(1, 128);
      for (int col = et; col < e.out_cpitch; col += 128) {
        float s_ = 0.f, q_ = 0.f;
        if (col < BN) {
          s_ = (smem_epi[col] + smem_epi[BN + col]) + (smem_epi[2 * BN + col] + smem_epi[3 * BN + col]);
          q_ = (smem_epi[4 * BN + col] + smem_epi[5 * BN + col]) + (smem_epi[6 * BN + col] + smem_epi[7 * BN + col]);
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

}  // namespace

// The flat kernel handles single-chunk inputs (C <= 64) with at most 64 output channels.
bool conv3x3_flat_ok(const ActView& in, int mode, int cout) {
  static const int enabled = env_int("MIMO_CONV_FLAT", 1);
  if (!enabled) return false;
  if (in.C > 64 || round_up(cout, 16) > 64) return false;
  if (mode == 0 && in.pad != 1) return false;
  if (mode == 1 && in.pad != 2) return false;
  // int32 pixel indices inside the kernel
  if ((long long)in.N * in.hb() * in.wb() >= (1ll << 31) - 256) return false;
  return true;
}

int conv3x3_flat_launch(const ActView& in, int mode, const bf16* wpacked, int cout, int cin_pitch, bf16* out, int out_cpitch,
                        float* stat_sum, float* stat_sq, const float* bias, int relu, cudaStream_t stream) {
  FlatParams p{};
  p.wb = in.wb();
  p.img_pix = in.hb() * in.wb();
  p.total_pos = (long long)in.N * p.img_pix;
  p.origin = mode == 0 ? 0 : -(2 * p.wb + 2);
  p.out_h = mode == 0 ? in.H : in.H + 2;
  p.out_w = mode == 0 ? in.W : in.W + 2;
  p.n_img = in.N;
  p.m_tiles = (int)ceil_div_ll(p.total_pos, kBlockM);
  p.block_n = round_up(cout, 16);
  p.k_steps = ceil_div(in.C, 16);
  p.bo_mode = env_int("MIMO_FLAT_BO", 0);
  p.pf_dist = env_int("MIMO_FLAT_PF", 0);
  p.ko = env_int("MIMO_FLAT_KO", 0);
  p.trace = nullptr;
  if (env_int("MIMO_FLAT_TRACE", 0)) {
    static long long* trace_buf = nullptr;
    if (!trace_buf) cudaMalloc(&trace_buf, 3 * 64 * 4 * sizeof(long long));
    cudaMemsetAsync(trace_buf, 0, 3 * 64 * 4 * sizeof(long long), stream);
    p.trace = trace_buf;
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

  const int b_bytes = (9 * p.block_n * 128 + 1023) & ~1023;
  const int fixed = b_bytes + 8 * p.block_n * 4 + (2 * kMaxStages + 5) * 8 + 16 + 64;
  const int smem_budget = 227 * 1024 - 1024;
  int stages = (smem_budget - fixed) / kStageBytes;
  if (stages > kMaxStages) stages = kMaxStages;
  {
    const int forced = env_int("MIMO_FLAT_STAGES", 0);
    if (forced >= 1 && forced < stages) stages = forced;
  }
  {
    static bool told = false;
    if (!told && (p.ko || p.pf_dist || p.bo_mode || env_int("MIMO_FLAT_STAGES", 0))) {
      told = true;
      fprintf(stderr, "mimo_b200: conv3x3_flat diagnostics active: ko=%d pf=%d bo=%d stages=%d\n", p.ko, p.pf_dist, p.bo_mode, stages);
    }
  }
  MIMO_CHECK(stages >= 1, MIMO_ERR_ARG, "conv3x3_flat: not enough shared memory for block_n=%d", p.block_n);
  p.stages = stages;
  const size_t smem_bytes = (size_t)stages * kStageBytes + fixed + 1024;

  CUtensorMap tm_in, tm_w;
  {
    // the whole buffer as a list of pixels: (C channels, N*(H+2)*(W+2) rows); rows outside are zero-filled
    uint64_t dims[2] = {(uint64_t)in.C, (uint64_t)p.total_pos};
    uint64_t strides[1] = {(uint64_t)in.cpitch * 2};
    uint32_t box[2] = {64, (uint32_t)kSegRows};
    int rc = encode_tmap_bf16(&tm_in, in.base + in.c_off, 2, dims, strides, box, 1);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)cin_pitch, (uint64_t)cout, 9};
    uint64_t strides[2] = {(uint64_t)cin_pitch * 2, (uint64_t)cout * cin_pitch * 2};
    uint32_t box[3] = {64, (uint32_t)p.block_n, 9};
    int rc = encode_tmap_bf16(&tm_w, wpacked, 3, dims, strides, box, 1);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    MIMO_CUDA(cudaFuncSetAttribute(conv3x3_flat_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    MIMO_CUDA(cudaFuncSetAttribute(conv3x3_flat_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    MIMO_CUDA(cudaFuncSetAttribute(conv3x3_flat_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    MIMO_CUDA(cudaFuncSetAttribute(conv3x3_flat_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int grid = p.m_tiles < num_sms() ? p.m_tiles : num_sms();
  switch (p.block_n) {
    case 16: conv3x3_flat_kernel<16><<<grid, kThreads, smem_bytes, stream>>>(tm_in, tm_w, p); break;
    case 32: conv3x3_flat_kernel<32><<<grid, kThreads, smem_bytes, stream>>>(tm_in, tm_w, p); break;
    case 48: conv3x3_flat_kernel<48><<<grid, kThreads, smem_bytes, stream>>>(tm_in, tm_w, p); break;
    default: conv3x3_flat_kernel<64><<<grid, kThreads, smem_bytes, stream>>>(tm_in, tm_w, p); break;
  }
  MIMO_LAUNCH_CHECK();
  if (p.trace) {
    // diagnostic only (synchronises!): dump CTA 0's pipeline timeline relative to its first event
    static long long host[3 * 64 * 4];
    cudaDeviceSynchronize();
    cudaMemcpy(host, p.trace, sizeof(host), cudaMemcpyDeviceToHost);
    long long t0 = host[0];
    for (int i = 0; i < 3 * 64 * 4; ++i) if (host[i] > 0 && host[i] < t0) t0 = host[i];
    static int dumps = 0;
    if (dumps++ < 2) {
      fprintf(stderr, "# flat trace (cycles since first event): tile | producer: empty_ok | mma: tmem_empty_ok full_ok issued committed | epi: loop_top tmem_full_ok released done\n");
      for (int i = 0; i < 24; ++i) {
        fprintf(stderr, "%3d | %7lld | %7lld %7lld %7lld %7lld | %7lld %7lld %7lld %7lld\n", i, host[(0 * 64 + i) * 4] - t0,
                host[(64 + i) * 4 + 0] - t0, host[(64 + i) * 4 + 1] - t0, host[(64 + i) * 4 + 2] - t0, host[(64 + i) * 4 + 3] - t0,
                host[(128 + i) * 4 + 0] - t0, host[(128 + i) * 4 + 1] - t0, host[(128 + i) * 4 + 2] - t0, host[(128 + i) * 4 + 3] - t0);
      }
    }
  }
  return MIMO_OK;
}

}  // namespace mimo
