// "Flat" sliding-window variant of the tcgen05 implicit-GEMM 3x3 convolution for the HBM-bound layers (few channels,
// many pixels): the full-resolution encoder / decoder convolutions and the first encoder level (reference
// components.py:23,26; shapes in SURVEY App. A: 3->21, 21->21, 63->31, 31->21 @ HxW and 21->42, 42->42 @ H/2 x W/2).
//
// Idea: the input buffer [N][H+2][W+2][C] is ONE long list of pixel rows (128 B each in shared memory). For an output
// position P (flattened, top-left aligned with the buffer) the nine taps are the rows  P + kh*(W+2) + kw  of that
// list, so
//   * a tile is 128 CONSECUTIVE positions, whatever W is (the 2 positions per row and 2 rows per image that fall on
//     the halo are computed and simply not stored: 3 % waste at 128x160);
//   * every CTA owns a CONTIGUOUS range of tiles and streams the pixel list ONCE through a ring of 128-row chunks in
//     shared memory: tile t needs the rows [128 t, 128 t + 2 (W+2) + 130), i.e. the chunks t .. t+NC-1, and moving to
//     tile t+1 needs exactly one new chunk. One TMA load (128 rows) per tile instead of nine (4-D kernel) or three
//     (one per kernel row): the TMA row rate (measured ~4 cycles per 128-byte row per SM) was the limiter;
//   * each tap's A operand is a 128-row window at an ARBITRARY row offset inside the ring: the UMMA descriptor start
//     address is simply advanced by 128 B per row. B200's SWIZZLE_128B is a function of the absolute shared-memory
//     address bits, so row-shifted windows need no base-offset field (measured; tests/test_kernels_gpu.py). Windows that
//     run past the last ring slot continue into a mirror copy of slot 0 stored after it;
//   * the packed weights of all nine taps (<= 72 KB) are loaded ONCE per persistent CTA and stay resident.
// mode 0 (fprop): in = pad==1 (reflect halo) view, output = dense [N][H][W] (+ BatchNorm statistics).
// mode 1 (dgrad): in = pad==2 (zero tail) view of dY; output = padded-domain gradient [N][H+2][W+2]; tap (a,b) reads
//                 position Q + (a-2)*(W+2) + (b-2): rows/columns "before" an image row are the zero tail of the
//                 previous row / image (TMA zero-fills negative coordinates), so EVERY position is a valid output.
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, then 4 or 8 epilogue warps (two sets of four
// alternate tiles / TMEM accumulators). The epilogue has no shared-memory staging and no per-tile block barrier: rows go
// straight from registers to global memory and the BatchNorm statistics are accumulated per thread in registers.
#include "common.cuh"
#include "conv_epilogue.cuh"
#include "ops.h"

#include <stdlib.h>

namespace mimo {
namespace {

constexpr int kBlockM = 128;
constexpr int kChunkBytes = kBlockM * 128;  // one ring slot: 128 rows x 128 B
constexpr int kMaxSlots = 12;

struct FlatParams {
  int wb;              // buffer row pitch in pixels (W + 2)
  int img_pix;         // pixels per image in the buffer ((H+2)*(W+2))
  int origin;          // first tap row offset: 0 (fprop) or -(2*wb + 2) (dgrad)
  int out_h, out_w;    // stored output domain inside the (H+2)x(W+2) position grid
  int n_img;
  int m_tiles;         // all tiles
  int tiles_per_cta;   // contiguous tile range per CTA
  int slots;           // ring slots S (the mirror of slot 0 is stored as slot S)
  int nc;              // chunks a tile touches: ceil((2*wb + 130) / 128)
  int k_steps;         // ceil(C / 16): 16-channel MMA k-steps that carry data
  int ko;              // diagnostic knock-outs (env MIMO_FLAT_KO): 1 no global stores, 2 no TMA loads, 4 no MMAs
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

template <int BN, int SETS>
__global__ void __launch_bounds__(64 + 128 * SETS, 1)
conv3x3_flat_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_w, const FlatParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [B: 9 x BN x 128 B][ring: (S+1) x 16 KB][end-of-kernel statistics partials][barriers]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem;
  constexpr int b_bytes = 9 * BN * 128;  // BN % 8 == 0 -> every tap starts 1024-byte aligned
  uint8_t* smem_a = smem_b + ((b_bytes + 1023) & ~1023);
  float* smem_epi = reinterpret_cast<float*>(smem_a + (size_t)(p.slots + 1) * kChunkBytes);  // [2][4*SETS][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + 8 * SETS * BN);
  uint64_t* full_bar = bars;                    // [slots]
  uint64_t* empty_bar = bars + kMaxSlots;       // [slots]
  uint64_t* tmem_full = bars + 2 * kMaxSlots;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint64_t* b_full = tmem_empty + 2;            // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t tmem_cols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : 256;

  // this CTA's contiguous tile range
  const int t_begin = blockIdx.x * p.tiles_per_cta;
  const int n_tiles = min(p.tiles_per_cta, p.m_tiles - t_begin);
  const int S = p.slots;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_in);
    prefetch_tmap(&tmap_w);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tmem_full[a], 1);
        mbar_init(&tmem_empty[a], 4);  // one arrive per epilogue warp of the set that drains this accumulator
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
    if (elect_one()) {
      mbar_arrive_expect_tx(b_full, (uint32_t)b_bytes);
      tma_load_3d(&tmap_w, b_full, smem_b, 0, 0, 0);  // box (64 cin, BN cout, 9 taps)
    }
    __syncwarp();
    const int n_chunks = n_tiles + p.nc - 1;   // chunk c (relative) = rows [128 (t_begin + c) + origin, +128)
    int slot = 0;
    uint32_t phase = 0;
    long long row0 = (long long)t_begin * kBlockM + p.origin;
    for (int c = 0; c < n_chunks; ++c, row0 += kBlockM) {
      mbar_wait(&empty_bar[slot], phase ^ 1);  // the chunk that lived here (c - S) is dead once tile c - S has been multiplied
      if (p.trace && blockIdx.x == 0 && c < 64 && lane == 0) p.trace[(0 * 64 + c) * 4 + 0] = clock64();
      if (elect_one()) {
        if (p.ko & 2) {
          mbar_arrive(&full_bar[slot]);
        } else {
          mbar_arrive_expect_tx(&full_bar[slot], slot == 0 ? 2 * kChunkBytes : kChunkBytes);
          tma_load_2d(&tmap_in, &full_bar[slot], smem_a + (size_t)slot * kChunkBytes, 0, (int)row0);
          if (slot == 0) tma_load_2d(&tmap_in, &full_bar[slot], smem_a + (size_t)S * kChunkBytes, 0, (int)row0);  // mirror
        }
      }
      __syncwarp();
      if (++slot == S) { slot = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // One thread issues, so every instruction counts: the loop is warp-uniform (tcgen05 under elect_one()), descriptor
    // words are precomputed (common.cuh), per-MMA work is a couple of 32-bit adds.
    const uint32_t idesc = make_idesc_bf16(kBlockM, BN, 0, 0);
    constexpr uint32_t hi = desc_hi(1024, kLayoutSW128);
    const uint32_t b_lo0 = desc_lo(smem_u32(smem_b), 16);
    const uint32_t a_lo0 = desc_lo(smem_u32(smem_a), 16);
    const uint32_t ks = (uint32_t)p.k_steps;
    const bool no_mma = (p.ko & 4) != 0;
    const uint32_t ring_rows = (uint32_t)S * kBlockM;
    // row offset of every tap inside the tile's window
    uint32_t tap_row[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) tap_row[kh * 3 + kw] = (uint32_t)(kh * p.wb + kw);
    mbar_wait(b_full, 0);
    int waited = 0;          // chunks whose arrival has been observed
    int wslot = 0;
    uint32_t wphase = 0;
    int slot = 0;            // ring slot of chunk i (the first chunk of tile i)
    for (int i = 0; i < n_tiles; ++i) {
      const uint32_t acc = (uint32_t)i & 1u;
      const bool tr = p.trace && blockIdx.x == 0 && i < 64 && lane == 0;
      long long* trow = p.trace + (1 * 64 + i) * 4;
      if (tr) trow[0] = clock64();
      mbar_wait(&tmem_empty[acc], (((uint32_t)i >> 1) & 1u) ^ 1u);
      if (tr) trow[1] = clock64();
      while (waited < i + p.nc) {
        mbar_wait(&full_bar[wslot], wphase);
        ++waited;
        if (++wslot == S) { wslot = 0; wphase ^= 1; }
      }
      if (tr) trow[2] = clock64();
      tc_fence_after();
      const uint32_t d_addr = tmem_base + acc * BN;
      const uint32_t base_row = (uint32_t)slot * kBlockM;
      if (elect_one()) {
        if (!no_mma) {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            uint32_t r = base_row + tap_row[tap];
            if (r >= ring_rows) r -= ring_rows;          // windows that START past the ring end wrap; windows that only
            const uint32_t a_lo = a_lo0 + r * 8;         // END past it continue into the mirror of slot 0
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if ((uint32_t)k < ks)
                umma_bf16_w(d_addr, a_lo + k * 2, hi, b_lo0 + ((tap * BN * 128 + k * 32) >> 4), hi, idesc, (tap | k) != 0);
            }
          }
        }
        umma_commit(&empty_bar[slot]);    // chunk i is dead once these MMAs have read it
        umma_commit(&tmem_full[acc]);     // accumulator complete -> epilogue
      }
      __syncwarp();
      if (tr) trow[3] = clock64();
      if (++slot == S) slot = 0;
    }
  } else {
    // ===================== epilogue (SETS x 4 warps) =====================
    // Thread (set, q, lane) owns accumulator row q*32+lane of every tile of its set (tiles with i % SETS == set; with
    // SETS == 2 a set always drains TMEM accumulator `set`) = one output position. Rows are stored straight from
    // registers (16-byte stores; consecutive positions are consecutive pixels in memory); BatchNorm statistics are
    // accumulated per THREAD in registers over all tiles, then reduced across threads once at the end of the kernel
    // (fixed order -> deterministic).
    const int ew = warp - 2;
    const int q = warp & 3;             // TMEM lane quarter this warp may access (warp id % 4)
    const int set = ew >> 2;            // 0 or 1
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;    // 0 .. 128*SETS-1
    const EpiArgs& e = p.epi;
    const bool stats = e.stat_sum != nullptr && e.scale == nullptr;
    float ssum[BN], ssq[BN];
#pragma unroll
    for (int i = 0; i < BN; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
    // position of my row in my first tile, then advanced incrementally (no divisions in the loop)
    int pn, ph, pw;
    {
      const long long pos0 = (long long)(t_begin + set) * kBlockM + row;
      pn = (int)(pos0 / p.img_pix);
      const int rem = (int)(pos0 - (long long)pn * p.img_pix);
      ph = rem / p.wb;
      pw = rem - ph * p.wb;
    }
    const int hb = p.img_pix / p.wb;
    int dn, dh, dw;
    {
      const int d = SETS * kBlockM;
      dn = d / p.img_pix;
      const int rem = d - dn * p.img_pix;
      dh = rem / p.wb;
      dw = rem - dh * p.wb;
    }
    // fused inference epilogue: the per-channel affine lives in the registers the (unused) statistics would occupy
    const bool affine = e.scale != nullptr;
    if (affine) {
#pragma unroll
      for (int i = 0; i < BN; ++i) {
        ssum[i] = i < e.cout ? __ldg(e.scale + i) : 0.f;
        ssq[i] = i < e.cout ? __ldg(e.bias + i) : 0.f;
      }
    }
    for (int i = set; i < n_tiles; i += SETS) {
      const uint32_t acc = (uint32_t)i & 1u;
      const bool valid = pn < p.n_img && ph < p.out_h && pw < p.out_w && !(p.ko & 1);
      const size_t my_pix = e.halo ? ((size_t)pn * (p.out_h + 2) + ph + 1) * (p.out_w + 2) + pw + 1 : (size_t)(pn * p.out_h + ph) * p.out_w + pw;
      const float* drow = (e.drop != nullptr && valid) ? e.drop + (size_t)pn * e.cout : nullptr;
      const bool tr = p.trace && blockIdx.x == 0 && i < 64 && (et & 127) == 0;
      long long* trow = p.trace + (2 * 64 + i) * 4;
      if (tr) trow[0] = clock64();
      mbar_wait(&tmem_full[acc], ((uint32_t)i >> 1) & 1u);
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
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[c + j]);
          if (affine) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], ssum[c + j], ssq[c + j]);
          } else if (e.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += (c + j < e.cout) ? __ldg(e.bias + c + j) : 0.f;
          }
          if (e.relu & 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (drow != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] *= (c + j < e.cout) ? __ldg(drow + c + j) : 0.f;
          }
          bf16x8 o;
#pragma unroll
          for (int j = 0; j < 8; ++j) o.v[j] = __float2bfloat16_rn(v[j]);
          if (c < e.out_cmax) *reinterpret_cast<bf16x8*>(dst + c) = o;
          if (stats) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float f = __bfloat162float(o.v[j]);  // statistics of the values as stored
              ssum[c + j] += f;
              ssq[c + j] = fmaf(f, f, ssq[c + j]);
            }
          }
        }
      }
      if (tr) trow[3] = clock64();
      // advance my position by SETS tiles
      pw += dw;
      if (pw >= p.wb) { pw -= p.wb; ++ph; }
      ph += dh;
      if (ph >= hb) { ph -= hb; ++pn; }
      pn += dn;
    }
    if (stats) {
      // cross-thread reduction, once per CTA: butterfly inside each warp, then the warps through shared memory
      constexpr int NW = 4 * SETS;
#pragma unroll
      for (int c = 0; c < BN; c += 16) {
        const float cs = warp_colsum16(ssum + c, lane);
        const float cq = warp_colsum16(ssq + c, lane);
        if ((lane & 1) == 0) smem_epi[ew * BN + c + (lane >> 1)] = cs;
        else smem_epi[(NW + ew) * BN + c + (lane >> 1)] = cq;
      }
      named_bar_sync(1, 128 * SETS);
      for (int col = et; col < e.out_cpitch; col += 128 * SETS) {
        float s_ = 0.f, q_ = 0.f;
        if (col < BN) {
#pragma unroll
          for (int w = 0; w < NW; ++w) {
            s_ += smem_epi[w * BN + col];
            q_ += smem_epi[(NW + w) * BN + col];
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

// shared-memory plan: returns the number of ring slots that fit (0 = does not fit)
int plan_slots(int block_n, int sets, int nc, size_t* smem_bytes) {
  const int b_bytes = (9 * block_n * 128 + 1023) & ~1023;
  const int fixed = b_bytes + 8 * sets * block_n * 4 + (2 * kMaxSlots + 5) * 8 + 16 + 64 + 1024 /*alignment slack*/;
  const int budget = 227 * 1024;
  int slots = (budget - fixed) / kChunkBytes - 1;  // one extra slot holds the mirror of slot 0
  if (slots > kMaxSlots) slots = kMaxSlots;
  if (slots < nc + 1) return 0;                    // need the tile's window plus at least one chunk of prefetch
  *smem_bytes = (size_t)fixed + (size_t)(slots + 1) * kChunkBytes;
  return slots;
}

template <int BN, int SETS>
int launch_flat(const CUtensorMap& tm_in, const CUtensorMap& tm_w, const FlatParams& p, size_t smem_bytes, int grid, cudaStream_t stream) {
  MIMO_CUDA(cudaFuncSetAttribute(conv3x3_flat_kernel<BN, SETS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));   // per launch: the attribute is per DEVICE, a process-wide "done" flag would skip the other GPUs
  conv3x3_flat_kernel<BN, SETS><<<grid, 64 + 128 * SETS, smem_bytes, stream>>>(tm_in, tm_w, p);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

}  // namespace

// The flat kernel handles single-chunk inputs (C <= 64) with at most 64 output channels whose ring fits shared memory.
bool conv3x3_flat_ok(const ActView& in, int mode, int cout) {
  static const int enabled = env_int("MIMO_CONV_FLAT", 1);
  if (!enabled) return false;
  if (in.C > 64 || round_up(cout, 16) > 64) return false;
  if (mode == 0 && in.pad != 1) return false;
  if (mode == 1 && in.pad != 2) return false;
  // int32 pixel indices inside the kernel
  if ((long long)in.N * in.hb() * in.wb() >= (1ll << 31) - 4096) return false;
  size_t smem;
  const int bn = round_up(cout, 16);
  return plan_slots(bn, bn <= 48 ? 2 : 1, ceil_div(2 * in.wb() + 130, 128), &smem) > 0;
}

int conv3x3_flat_launch(const ActView& in, int mode, const bf16* wpacked, int cout, int cin_pitch, bf16* out, int out_cpitch,
                        float* stat_sum, float* stat_sq, const float* bias, int relu, cudaStream_t stream, const ConvFuse* fuse) {
  note_kernel(1);
  FlatParams p{};
  const int block_n = round_up(cout, 16);
  const int sets = block_n <= 48 ? 2 : 1;   // 8 epilogue warps unless the per-thread statistics registers do not fit
  p.wb = in.wb();
  p.img_pix = in.hb() * in.wb();
  const long long total_pos = (long long)in.N * p.img_pix;
  p.origin = mode == 0 ? 0 : -(2 * p.wb + 2);
  p.out_h = mode == 0 ? in.H : in.H + 2;
  p.out_w = mode == 0 ? in.W : in.W + 2;
  p.n_img = in.N;
  p.m_tiles = (int)ceil_div_ll(total_pos, kBlockM);
  p.tiles_per_cta = ceil_div(p.m_tiles, num_sms());
  const int grid = ceil_div(p.m_tiles, p.tiles_per_cta);
  p.nc = ceil_div(2 * p.wb + 130, 128);
  p.k_steps = ceil_div(in.C, 16);
  p.ko = env_int("MIMO_FLAT_KO", 0);
  p.trace = nullptr;
  if (env_int("MIMO_FLAT_TRACE", 0)) {
    static long long* trace_buf = nullptr;
    if (!trace_buf) cudaMalloc(&trace_buf, 3 * 64 * 4 * sizeof(long long));
    cudaMemsetAsync(trace_buf, 0, 3 * 64 * 4 * sizeof(long long), stream);
    p.trace = trace_buf;
  }
  size_t smem_bytes = 0;
  p.slots = plan_slots(block_n, sets, p.nc, &smem_bytes);
  MIMO_CHECK(p.slots > 0, MIMO_ERR_ARG, "conv3x3_flat: ring of %d chunks does not fit shared memory (block_n=%d)", p.nc, block_n);
  p.epi.block_n = block_n;
  p.epi.cout = cout;
  p.epi.out_cpitch = out_cpitch;
  p.epi.stage_pitch = 0;
  p.epi.stat_rows = conv3x3_stat_rows();
  p.epi.out = out;
  p.epi.stat_sum = stat_sum;
  p.epi.stat_sq = stat_sq;
  p.epi.bias = fuse ? fuse->shift : bias;
  p.epi.relu = fuse ? ((fuse->relu ? 1 : 0) | 2) : relu;   // bit 1: fused affine vectors are padded / aligned for 16-byte loads
  p.epi.scale = fuse ? fuse->scale : nullptr;
  p.epi.drop = fuse ? fuse->drop : nullptr;
  p.epi.halo = fuse ? fuse->halo : 0;
  p.epi.out_cmax = fuse ? round_up(cout, 8) : out_cpitch;
  {
    static bool told = false;
    if (!told && p.ko) {
      told = true;
      fprintf(stderr, "mimo_b200: conv3x3_flat diagnostics active: ko=%d\n", p.ko);
    }
  }

  CUtensorMap tm_in, tm_w;
  {
    // the whole buffer as a list of pixels: (C channels, N*(H+2)*(W+2) rows); rows outside are zero-filled
    uint64_t dims[2] = {(uint64_t)in.C, (uint64_t)total_pos};
    uint64_t strides[1] = {(uint64_t)in.cpitch * 2};
    uint32_t box[2] = {64, (uint32_t)kBlockM};
    int rc = encode_tmap_bf16(&tm_in, in.base + in.c_off, 2, dims, strides, box, 1);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)cin_pitch, (uint64_t)cout, 9};
    uint64_t strides[2] = {(uint64_t)cin_pitch * 2, (uint64_t)cout * cin_pitch * 2};
    uint32_t box[3] = {64, (uint32_t)block_n, 9};
    int rc = encode_tmap_bf16(&tm_w, wpacked, 3, dims, strides, box, 1);
    if (rc) return rc;
  }
  int rc;
  switch (block_n) {
    case 16: rc = launch_flat<16, 2>(tm_in, tm_w, p, smem_bytes, grid, stream); break;
    case 32: rc = launch_flat<32, 2>(tm_in, tm_w, p, smem_bytes, grid, stream); break;
    case 48: rc = launch_flat<48, 2>(tm_in, tm_w, p, smem_bytes, grid, stream); break;
    default: rc = launch_flat<64, 1>(tm_in, tm_w, p, smem_bytes, grid, stream); break;
  }
  if (rc == MIMO_OK && p.trace) {
    // diagnostic only (synchronises!): dump CTA 0's pipeline timeline relative to its first event
    static long long host[3 * 64 * 4];
    cudaDeviceSynchronize();
    cudaMemcpy(host, p.trace, sizeof(host), cudaMemcpyDeviceToHost);
    long long t0 = host[0];
    for (int i = 0; i < 3 * 64 * 4; ++i) if (host[i] > 0 && host[i] < t0) t0 = host[i];
    static int dumps = 0;
    if (dumps++ < 2) {
      fprintf(stderr, "# flat trace (cycles): i | producer: chunk_slot_free | mma: loop_top tmem_empty_ok chunks_ok issued | epi: loop_top tmem_full_ok released done\n");
      for (int i = 0; i < 40; ++i)
        fprintf(stderr, "%3d | %7lld | %7lld %7lld %7lld %7lld | %7lld %7lld %7lld %7lld\n", i, host[(0 * 64 + i) * 4] - t0,
                host[(64 + i) * 4 + 0] - t0, host[(64 + i) * 4 + 1] - t0, host[(64 + i) * 4 + 2] - t0, host[(64 + i) * 4 + 3] - t0,
                host[(128 + i) * 4 + 0] - t0, host[(128 + i) * 4 + 1] - t0, host[(128 + i) * 4 + 2] - t0, host[(128 + i) * 4 + 3] - t0);
    }
  }
  return rc;
}

}  // namespace mimo
