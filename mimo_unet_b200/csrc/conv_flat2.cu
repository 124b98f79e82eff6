// CTA-pair sliding-window tcgen05 3x3 convolution for the HBM-bound layers (<= 64 input and <= 64 output channels): the
// full-resolution encoder / decoder convolutions and the first encoder level of the reference U-Net (components.py:23,26;
// shapes in SURVEY App. A: 3->21, 21->21, 63->31, 31->21 @ HxW and 21->42, 42->42 @ H/2 x W/2), forward and dgrad.
// Successor of conv_flat.cu (same flattened addressing, same ring of 128-row chunks, resident weights, register
// statistics); what changed and why (measured, profiles/r02_findings.md):
//   * tcgen05.mma.cta_group::2: a single-CTA M128 x N32 x K16 MMA costs 59 cycles (43 cycles of fixed cost per instruction),
//     the pair's M256 x N32 one 39 cycles for twice the positions: 3x fewer tensor-pipe cycles per pixel. Each CTA of the
//     pair streams its own contiguous tile range through its own ring; the leader multiplies tile i of both ranges at once.
//   * the activation rows are no longer fetched by tensor-map TMA. A pixel of these layers is 16..128 bytes, and a tiled TMA
//     load costs 4..9 cycles per box ROW whatever its length (128 rows = 1150 cycles per tile at 48 B / pixel: slower than
//     the MMAs). The rows of the flattened buffer are CONTIGUOUS in global memory, so a chunk of 128 pixels is ONE
//     cp.async.bulk (no per-row cost) into a compact staging ring (deep: ~10 chunks in flight per SM), and two warps
//     re-lay it out into the SWIZZLE_128B operand ring (piece g of row r lands at r*128 + ((g ^ (r & 7)) << 4)), zeroing
//     pad channels and rows outside the buffer on the way; fence.proxy.async + an (if needed remote) mbarrier arrive
//     publish the chunk to the MMA warp.
//   * eight TMEM accumulators instead of two: the MMA -> commit -> epilogue -> (remote) arrive -> MMA round trip costs a few
//     thousand cycles across the pair, a tile only ~700.
// mode 0 (fprop): in = pad==1 (reflect halo) view, output = dense [N][H][W] (+ BatchNorm statistics).
// mode 1 (dgrad): in = pad==2 (zero tail) view of dY; output = padded-domain gradient [N][H+2][W+2].
//
// Warp roles (64 + 128*SETS + 64 threads): warp 0 = weight TMA + bulk-copy producer, warp 1 = TMEM allocator + MMA issuer
// (leader CTA only), then 4*SETS epilogue warps (sets alternate tiles), then 2 re-layout warps.
#include "common.cuh"
#include "conv_epilogue.cuh"
#include "ops.h"

#include <stdlib.h>

namespace mimo {
namespace {

#ifdef MIMO_FLAT2_POLL
#define MBAR_WAIT mbar_wait_poll
#else
#define MBAR_WAIT mbar_wait
#endif

constexpr int kBlockM = 128;
constexpr int kChunkBytes = kBlockM * 128;  // one ring slot: 128 rows x 128 B
constexpr int kMaxSlots = 12;
constexpr int kLoaderWarps = 2;                // 64 re-layout threads: the register budget of the 8 epilogue warps (per-thread statistics) matters more
constexpr int kMaxStage = 16;               // staging slots (compact copies of chunks in flight)
constexpr int kAcc = 8;                     // TMEM accumulators

struct Flat2Params {
  int wb;              // buffer row pitch in pixels (W + 2)
  int img_pix;         // pixels per image in the buffer ((H+2)*(W+2))
  long long total_pos;
  int origin;          // first tap row offset: 0 (fprop) or -(2*wb + 2) (dgrad)
  int out_h, out_w;    // stored output domain inside the (H+2)x(W+2) position grid
  int n_img;
  int tiles_per_cta;   // contiguous tile range per CTA (every CTA runs exactly this many; tiles past the end are empty)
  int slots;           // operand ring slots S (the mirror of slot 0 is stored as slot S)
  int st_slots, st_bytes;   // staging ring: slots and bytes per slot (128 pixels as they lie in global memory)
  int nc;              // chunks a tile touches: ceil((2*wb + 130) / 128)
  int k_steps;         // ceil(C / 16): 16-channel MMA k-steps that carry data
  int groups;          // 16-byte pieces per pixel in global memory (cpitch / 8)
  int c_valid;         // channels of the view; pieces holding channels >= c_valid are masked to zero in shared memory
  const bf16* in;      // first channel of the view at flattened position 0
  int ko;              // diagnostic knock-outs (env MIMO_FLAT2_KO): 1 no global stores, 2 no loads, 4 no MMAs, 8 no re-layout
  long long* trace;    // diagnostic (env MIMO_FLAT2_TRACE): CTA 0 records clock64() of its pipeline events, [4 roles][64][4]
  EpiArgs epi;
};

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// One warp re-lays one staged chunk (128 pixels of G 16-byte pieces, as they lie in global memory) out into a SWIZZLE_128B
// operand slot (shared-space addresses; branch-free body, four independent loads in flight). all_in: every row of the chunk
// lies inside the buffer (warp-uniform).
template <int G>
__device__ __forceinline__ void relayout_chunk(uint32_t src, uint32_t dst, uint32_t mirror, int lane, int last_g, const uint4& tail_mask,
                                               bool all_in, long long row0, long long total_pos) {
  constexpr int kIters = kBlockM * G / 32;   // 4 * G
#pragma unroll
  for (int k0 = 0; k0 < kIters; k0 += 4) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = lds128(src + (uint32_t)(lane + 32 * (k0 + u)) * 16u);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int item = lane + 32 * (k0 + u);
      const int r = item / G, g = item - r * G;
      const bool keep = g <= last_g && (all_in || (row0 + r >= 0 && row0 + r < total_pos));
      const uint32_t m = keep ? 0xffffffffu : 0u;
      const bool edge = g == last_g;
      uint4 w;
      w.x = v[u].x & m & (edge ? tail_mask.x : 0xffffffffu);
      w.y = v[u].y & m & (edge ? tail_mask.y : 0xffffffffu);
      w.z = v[u].z & m & (edge ? tail_mask.z : 0xffffffffu);
      w.w = v[u].w & m & (edge ? tail_mask.w : 0xffffffffu);
      const uint32_t off = (uint32_t)r * 128u + (uint32_t)((g ^ (r & 7)) << 4);
      sts128(dst + off, w);
      if (mirror) sts128(mirror + off, w);
    }
  }
}

// CG = CTAs per MMA group: 2 = CTA pair (tcgen05 cta_group::2, launched as 2-CTA clusters), 1 = the same pipeline with
// single-CTA MMAs (cta_group::1, no cluster) -- it keeps the bulk-copy + re-layout loader, whose per-row cost (none) is what
// the TMA-tiled loader of conv_flat.cu is bound by.
template <int BN, int SETS, int CG>
__global__ void __launch_bounds__(64 + 128 * SETS + 32 * kLoaderWarps, 1)
conv3x3_flat2_kernel(const __grid_constant__ CUtensorMap tmap_w, const Flat2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [B: 9 x BN/2 x 128 B][ring: (S+1) x 16 KB][end-of-kernel statistics partials][barriers]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_b = smem;
  constexpr int b_tap_bytes = (BN / CG) * 128;    // BN/CG % 8 == 0 -> every tap starts 1024-byte aligned
  constexpr int b_bytes = 9 * b_tap_bytes;
  uint8_t* smem_a = smem_b + ((b_bytes + 1023) & ~1023);
  uint8_t* smem_st = smem_a + (size_t)(p.slots + 1) * kChunkBytes;                           // staging ring
  float* smem_epi = reinterpret_cast<float*>(smem_st + (size_t)p.st_slots * p.st_bytes);     // [2][4*SETS][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + 8 * SETS * BN);
  uint64_t* full_bar = bars;                    // [slots]   leader only: 2 CTAs x re-layout warps arrive
  uint64_t* empty_bar = bars + kMaxSlots;       // [slots]   per CTA (multicast commit)
  uint64_t* st_full = bars + 2 * kMaxSlots;     // [st_slots] per CTA: bulk copy landed
  uint64_t* st_empty = st_full + kMaxStage;     // [st_slots] per CTA: re-layout warps are done with the slot
  uint64_t* tmem_full = st_empty + kMaxStage;   // [kAcc]    per CTA (multicast commit)
  uint64_t* tmem_empty = tmem_full + kAcc;      // [kAcc]    leader only: 4 warps x 2 CTAs
  uint64_t* b_full = tmem_empty + kAcc;         // [1]       leader only
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  constexpr uint32_t tmem_cols = (kAcc * BN <= 128) ? 128 : (kAcc * BN <= 256) ? 256 : 512;

  const int t_begin = blockIdx.x * p.tiles_per_cta;   // this CTA's contiguous tile range
  const int n_tiles = p.tiles_per_cta;
  const int S = p.slots;

  // the ring is zero-initialised once: the 16-byte pieces of a row that no copy ever writes (K padding up to the next
  // k-step) are read by the MMAs and must be finite
  for (int i = threadIdx.x; i < (S + 1) * (kChunkBytes / 16); i += blockDim.x)
    reinterpret_cast<uint4*>(smem_a)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();

  if (warp == 0 && lane == 0) prefetch_tmap(&tmap_w);
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(&full_bar[s], 2 * CG);   // slot s of an even/odd chunk PAIR (index s/2 used): 2 chunks x CG CTAs
      }
      for (int a = 0; a < kAcc; ++a) {
        mbar_init(&tmem_full[a], 1);
        mbar_init(&tmem_empty[a], 8 * CG);  // accumulator PAIR (index a/2 used): 2 tiles x 4 warps x CG CTAs
      }
      for (int a = 0; a < p.st_slots; ++a) {
        mbar_init(&st_full[a], 1);
        mbar_init(&st_empty[a], 1);
      }
      mbar_init(b_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    if constexpr (CG == 2) { tmem_alloc2(tmem_ptr, tmem_cols); tmem_relinquish2(); }
    else { tmem_alloc(tmem_ptr, tmem_cols); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== resident weights: each CTA loads ITS half of the rows of all nine taps, once =====================
    if (elect_one()) {
      if (rank == 0) mbar_arrive_expect_tx(b_full, (uint32_t)CG * (uint32_t)b_bytes);
      if constexpr (CG == 2) tma_load_3d_cg2(&tmap_w, mapa_shared(smem_u32(b_full), 0), smem_b, 0, (int)rank * (BN / 2), 0);   // box (64 cin, BN/2 cout, 9 taps)
      else tma_load_3d(&tmap_w, b_full, smem_b, 0, 0, 0);
    }
    __syncwarp();
    // ===================== bulk-copy producer: chunk c = rows [128 (t_begin + c) + origin, +128) of the pixel list =====================
    const int n_chunks = (n_tiles + p.nc - 1 + 1) & ~1;
    const size_t row_bytes = (size_t)p.groups * 16;
    const uint8_t* gbase = reinterpret_cast<const uint8_t*>(p.in);
    long long row0 = (long long)t_begin * kBlockM + p.origin;
    int ss = 0;
    uint32_t sphase = 0;
    for (int c = 0; c < n_chunks; ++c, row0 += kBlockM) {
      if (p.trace && blockIdx.x == 0 && c < 64 && lane == 0) p.trace[(3 * 64 + c) * 4 + 0] = clock64();
      MBAR_WAIT(&st_empty[ss], sphase ^ 1);
      if (p.trace && blockIdx.x == 0 && c < 64 && lane == 0) p.trace[(3 * 64 + c) * 4 + 1] = clock64();
      if (elect_one()) {
        const long long lo = row0 < 0 ? 0 : row0;
        const long long hi = row0 + kBlockM > p.total_pos ? p.total_pos : row0 + kBlockM;
        if (hi > lo && !(p.ko & 2)) {
          const uint32_t bytes = (uint32_t)((hi - lo) * (long long)row_bytes);
          mbar_arrive_expect_tx(&st_full[ss], bytes);
          bulk_load(smem_st + (size_t)ss * p.st_bytes + (size_t)(lo - row0) * row_bytes, gbase + lo * (long long)row_bytes, bytes, &st_full[ss]);
        } else {
          mbar_arrive(&st_full[ss]);
        }
      }
      __syncwarp();
      if (++ss == p.st_slots) { ss = 0; sphase ^= 1; }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===================== MMA issuer (leader CTA) =====================
      const uint32_t idesc = make_idesc_bf16(CG * kBlockM, BN, 0, 0);
      constexpr uint32_t hi = desc_hi(1024, kLayoutSW128);
      const uint32_t b_lo0 = desc_lo(smem_u32(smem_b), 16);
      const uint32_t a_lo0 = desc_lo(smem_u32(smem_a), 16);
      const uint32_t ks = (uint32_t)p.k_steps;
      const bool no_mma = (p.ko & 4) != 0;
      const uint32_t ring_rows = (uint32_t)S * kBlockM;
      uint32_t tap_row[9];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) tap_row[kh * 3 + kw] = (uint32_t)(kh * p.wb + kw);
      MBAR_WAIT(b_full, 0);
      // Tiles are multiplied in PAIRS (2j, 2j+1): per pair ONE wait for the two accumulators, ONE wait per chunk pair and ONE
      // commit. An mbarrier wait costs the issuing warp 120..200 cycles even when the phase is long complete, and the issuing
      // thread blocks while the tensor queue is full, so per-tile waits (350 cycles) sat serially next to the 700 MMA cycles.
      int waited = 0;          // chunk PAIRS whose arrival has been observed
      int slot = 0;            // ring slot of chunk i (the first chunk of tile i)
      for (int i = 0; i < n_tiles; i += 2) {
        const uint32_t ap = ((uint32_t)i >> 1) % (kAcc / 2);          // accumulator pair
        const uint32_t ause = ((uint32_t)i >> 1) / (kAcc / 2);
        const bool tr = p.trace && blockIdx.x == 0 && i < 64 && lane == 0;
        long long* trow = p.trace + (1 * 64 + i) * 4;
        if (tr) trow[0] = clock64();
        MBAR_WAIT(&tmem_empty[2 * ap], (ause & 1u) ^ 1u);
        if (tr) trow[1] = clock64();
        // tile i + 1 reads chunks up to i + nc: chunk pairs up to (i + nc) / 2
        while (waited <= (i + p.nc) / 2) {
          const int ps = waited % (S / 2);
          MBAR_WAIT(&full_bar[2 * ps], (uint32_t)(waited / (S / 2)) & 1u);
          ++waited;
        }
        if (tr) trow[2] = clock64();
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const uint32_t d_addr = tmem_base + (2 * ap + (uint32_t)t) * BN;
            uint32_t s0 = (uint32_t)slot + (uint32_t)t;
            if (s0 >= (uint32_t)S) s0 -= (uint32_t)S;
            const uint32_t base_row = s0 * kBlockM;
            if (!no_mma) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                uint32_t r = base_row + tap_row[tap];
                if (r >= ring_rows) r -= ring_rows;          // windows that START past the ring end wrap; windows that only
                const uint32_t a_lo = a_lo0 + r * 8;         // END past it continue into the mirror of slot 0
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  if ((uint32_t)k < ks) {
                    if constexpr (CG == 2) umma2_bf16_w(d_addr, a_lo + k * 2, hi, b_lo0 + ((tap * b_tap_bytes + k * 32) >> 4), hi, idesc, (tap | k) != 0);
                    else umma_bf16_w(d_addr, a_lo + k * 2, hi, b_lo0 + ((tap * b_tap_bytes + k * 32) >> 4), hi, idesc, (tap | k) != 0);
                  }
                }
              }
            }
          }
          // ONE commit per tile pair: both accumulators complete -> both epilogue sets of both CTAs; it also tells the re-layout
          // warps that chunks i and i + 1 are dead
          if constexpr (CG == 2) umma_commit2_mc(&tmem_full[2 * ap], 3);
          else umma_commit(&tmem_full[2 * ap]);
        }
        __syncwarp();
        if (tr) trow[3] = clock64();
        slot += 2;
        if (slot >= S) slot -= S;
      }
    }
  } else if (warp < 2 + 4 * SETS) {
    // ===================== epilogue (SETS x 4 warps per CTA, each CTA drains its own rows) =====================
    const int ew = warp - 2;
    const int q = warp & 3;             // TMEM lane quarter this warp may access (warp id % 4)
    const int set = ew >> 2;            // 0 or 1
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;    // 0 .. 128*SETS-1
    const EpiArgs& e = p.epi;
    const bool stats = e.stat_sum != nullptr;
    const uint32_t tempty0 = mapa_shared(smem_u32(tmem_empty), 0);
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
    for (int i = set; i < n_tiles; i += SETS) {
      const uint32_t acc = (uint32_t)i % kAcc;                      // accumulator of tile i; barriers belong to the PAIR
      const uint32_t ap = ((uint32_t)i >> 1) % (kAcc / 2), ause = ((uint32_t)i >> 1) / (kAcc / 2);
      const bool valid = pn < p.n_img && ph < p.out_h && pw < p.out_w && !(p.ko & 1);
      const size_t my_pix = (size_t)(pn * p.out_h + ph) * p.out_w + pw;
      const bool tr = p.trace && blockIdx.x == 0 && i < 64 && (et & 127) == 0;
      long long* trow = p.trace + (2 * 64 + i) * 4;
      if (tr) trow[0] = clock64();
      MBAR_WAIT(&tmem_full[2 * ap], ause & 1u);
      if (tr) trow[1] = clock64();
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      uint32_t r[BN];
#pragma unroll
      for (int c = 0; c < BN; c += 16) tmem_ld16_nowait(t_addr + c, r + c);
      tmem_ld_wait();
      // TMEM drained -> hand the accumulator back to the leader's MMA warp before doing the math / stores
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty0 + 2u * ap * 8u);
      if (tr) trow[2] = clock64();
      if (valid) {
        bf16* dst = e.out + my_pix * e.out_cpitch;
#pragma unroll
        for (int c = 0; c < BN; c += 8) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[c + j]);
          if (e.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += (c + j < e.cout) ? __ldg(e.bias + c + j) : 0.f;
          }
          if (e.relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          const uint4 o = pack8(v);
          if (c < e.out_cpitch) *reinterpret_cast<uint4*>(dst + c) = o;
          if (stats) {
            float f[8];
            unpack8(o, f);   // statistics of the values as stored
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              ssum[c + j] += f[j];
              ssq[c + j] = fmaf(f[j], f[j], ssq[c + j]);
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
  } else {
    // ===================== re-layout warps: staging slot (pixels as in global memory) -> SWIZZLE_128B operand ring =====================
    // warp w handles the chunks c = w, w + kLoaderWarps, ... on its own, so the per-chunk latencies (barrier waits, proxy
    // fence, remote arrive) of the warps overlap
    const int lw = warp - (2 + 4 * SETS);
    const int G = p.groups;
    const int n_chunks = (n_tiles + p.nc - 1 + 1) & ~1;   // whole chunk pairs (a trailing chunk past the range is zero-filled)
    const uint32_t full0 = mapa_shared(smem_u32(full_bar), 0);
    // pieces that hold channels past the view (pad channels / garbage in memory) are zeroed, the boundary piece is masked
    const int last_g = (p.c_valid - 1) >> 3;                   // last piece with valid channels
    const uint4 tail_mask = group_mask(p.c_valid - last_g * 8);
    for (int c = lw; c < n_chunks; c += kLoaderWarps) {
      const int slot = c % S, ss = c % p.st_slots;
      const long long row0 = (long long)(t_begin + c) * kBlockM + p.origin;
      const bool tr = p.trace && blockIdx.x == 0 && c < 64 && lane == 0;
      long long* trow = p.trace + (0 * 64 + c) * 4;
      if (tr) trow[0] = clock64();
      MBAR_WAIT(&st_full[ss], (uint32_t)(c / p.st_slots) & 1u);
      if (c >= S) {
        // the chunk that lived in this slot (c - S) is dead once tile c - S has been multiplied: its accumulator-full barrier
        // (fewer than kAcc / 2 tile pairs can have completed past it, so the parity is unambiguous)
        const int tp = (c - S) >> 1;   // tile pair of tile c - S
        MBAR_WAIT(&tmem_full[2 * (tp % (kAcc / 2))], (uint32_t)(tp / (kAcc / 2)) & 1u);
      }
      if (tr) trow[0] = clock64();   // (trace: waits done)
      const uint32_t src = smem_u32(smem_st) + (uint32_t)ss * (uint32_t)p.st_bytes;
      const uint32_t dst = smem_u32(smem_a) + (uint32_t)slot * kChunkBytes;
      const uint32_t mirror = slot == 0 ? smem_u32(smem_a) + (uint32_t)S * kChunkBytes : 0u;   // mirror of slot 0
      const bool all_in = row0 >= 0 && row0 + kBlockM <= p.total_pos && !(p.ko & 2);
      const long long tp = (p.ko & 2) ? 0 : p.total_pos;
      if (!(p.ko & 8))   // (knock-out 8: no re-layout traffic at all)
      switch (G) {
        case 1: relayout_chunk<1>(src, dst, mirror, lane, last_g, tail_mask, all_in, row0, tp); break;
        case 2: relayout_chunk<2>(src, dst, mirror, lane, last_g, tail_mask, all_in, row0, tp); break;
        case 3: relayout_chunk<3>(src, dst, mirror, lane, last_g, tail_mask, all_in, row0, tp); break;
        case 4: relayout_chunk<4>(src, dst, mirror, lane, last_g, tail_mask, all_in, row0, tp); break;
        case 5: relayout_chunk<5>(src, dst, mirror, lane, last_g, tail_mask, all_in, row0, tp); break;
        case 6: relayout_chunk<6>(src, dst, mirror, lane, last_g, tail_mask, all_in, row0, tp); break;
        case 7: relayout_chunk<7>(src, dst, mirror, lane, last_g, tail_mask, all_in, row0, tp); break;
        default: relayout_chunk<8>(src, dst, mirror, lane, last_g, tail_mask, all_in, row0, tp); break;
      }
      if (tr) trow[1] = clock64();   // (trace: loop done)
      fence_proxy_async();
      if (tr) trow[2] = clock64();   // (trace: fence done)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(full0 + (uint32_t)(slot & ~1) * 8u);   // the pair's barrier lives at the even slot
        mbar_arrive(&st_empty[ss]);
      }
      if (tr) trow[3] = clock64();
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc2(tmem_base, tmem_cols);
    else tmem_dealloc(tmem_base, tmem_cols);
  }
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// shared-memory plan: operand ring of nc + 3 slots (+ mirror), the rest goes to the staging ring; false = does not fit
bool plan_smem(int block_n, int cg, int sets, int nc, int st_bytes, int* slots, int* st_slots, size_t* smem_bytes) {
  const int b_bytes = (9 * (block_n / cg) * 128 + 1023) & ~1023;
  const int fixed = b_bytes + 8 * sets * block_n * 4 + (2 * kMaxSlots + 2 * kMaxStage + 2 * kAcc + 1) * 8 + 16 + 64 + 1024 /*alignment slack*/;
  const int budget = 227 * 1024 - fixed;
  static const int s_extra = env_int("MIMO_FLAT2_SLACK", 4);
  int S = (nc + s_extra + 1) & ~1;   // even: the slots are handed over in pairs; window of a tile pair (nc + 1) + prefetch
  if (S > kMaxSlots) S = kMaxSlots;
  if (S < nc + 2) return false;
  int ns = (budget - (S + 1) * kChunkBytes) / st_bytes;
  while (ns < 3 && S - 2 >= nc + 2) { S -= 2; ns = (budget - (S + 1) * kChunkBytes) / st_bytes; }
  if (ns < 2) return false;
  if (ns > kMaxStage) ns = kMaxStage;
  *slots = S;
  *st_slots = ns;
  *smem_bytes = (size_t)fixed + (size_t)(S + 1) * kChunkBytes + (size_t)ns * st_bytes;
  return true;
}

template <int BN, int SETS, int CG>
int launch_flat2(const CUtensorMap& tm_w, const Flat2Params& p, size_t smem_bytes, int grid, cudaStream_t stream) {
  MIMO_CUDA(cudaFuncSetAttribute(conv3x3_flat2_kernel<BN, SETS, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)grid);
  lc.blockDim = dim3(64 + 128 * SETS + 32 * kLoaderWarps);
  lc.dynamicSmemBytes = smem_bytes;
  lc.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  lc.attrs = at;
  lc.numAttrs = CG == 2 ? 1 : 0;
  MIMO_CUDA(cudaLaunchKernelEx(&lc, conv3x3_flat2_kernel<BN, SETS, CG>, tm_w, p));
  return MIMO_OK;
}

template <int CG>
int dispatch_flat2(int block_n, const CUtensorMap& tm_w, const Flat2Params& p, size_t smem_bytes, int grid, cudaStream_t stream) {
  switch (block_n) {
    case 16: return launch_flat2<16, 2, CG>(tm_w, p, smem_bytes, grid, stream);
    case 32: return launch_flat2<32, 2, CG>(tm_w, p, smem_bytes, grid, stream);
    case 48: return launch_flat2<48, 2, CG>(tm_w, p, smem_bytes, grid, stream);
    default: return launch_flat2<64, 1, CG>(tm_w, p, smem_bytes, grid, stream);
  }
}

// 0: kernel not used; 1: single-CTA MMAs; 2: CTA pairs (env MIMO_CONV_FLAT2)
int flat2_mode() {
  static const int mode = env_int("MIMO_CONV_FLAT2", 0);
  return mode;
}

}  // namespace

bool conv3x3_flat2_ok(const ActView& in, int mode, int cout) {
  // opt-in (MIMO_CONV_FLAT2 = 1 single-CTA MMAs, 2 CTA pairs): see profiles/r02_findings.md
  if (flat2_mode() == 0) return false;
  if (flat2_mode() == 1 && round_up(cout, 16) == 16) return false;   // (BN / CG must stay a multiple of 8 rows: fine; BN=16 untested here)
  if (in.C > 64 || round_up(cout, 16) > 64) return false;
  if (mode == 0 && in.pad != 1) return false;
  if (mode == 1 && in.pad != 2) return false;
  if (in.c_off != 0 || in.cpitch > 64) return false;   // the loaders copy whole pixels (cpitch channels from channel 0)
  if ((long long)in.N * in.hb() * in.wb() >= (1ll << 31) - 65536) return false;
  size_t smem;
  int slots, st_slots;
  const int bn = round_up(cout, 16);
  return plan_smem(bn, flat2_mode() == 2 ? 2 : 1, bn <= 48 ? 2 : 1, ceil_div(2 * in.wb() + 130, 128), kBlockM * in.cpitch * 2, &slots, &st_slots, &smem);
}

int conv3x3_flat2_launch(const ActView& in, int mode, const bf16* wpacked, int cout, int cin_pitch, bf16* out, int out_cpitch,
                         float* stat_sum, float* stat_sq, const float* bias, int relu, cudaStream_t stream) {
  note_kernel(9);
  const int CG = flat2_mode() == 2 ? 2 : 1;
  Flat2Params p{};
  const int block_n = round_up(cout, 16);
  const int sets = block_n <= 48 ? 2 : 1;   // 8 epilogue warps unless the per-thread statistics registers do not fit
  p.wb = in.wb();
  p.img_pix = in.hb() * in.wb();
  p.total_pos = (long long)in.N * p.img_pix;
  p.origin = mode == 0 ? 0 : -(2 * p.wb + 2);
  p.out_h = mode == 0 ? in.H : in.H + 2;
  p.out_w = mode == 0 ? in.W : in.W + 2;
  p.n_img = in.N;
  const int m_tiles = (int)ceil_div_ll(p.total_pos, kBlockM);
  int grid = num_sms() & ~1;
  if (grid > round_up(m_tiles, 2)) grid = round_up(m_tiles, 2);
  p.tiles_per_cta = round_up(ceil_div(m_tiles, grid), 2);   // tiles are multiplied in pairs
  grid = round_up(ceil_div(m_tiles, p.tiles_per_cta), 2);
  p.nc = ceil_div(2 * p.wb + 130, 128);
  p.k_steps = ceil_div(in.C, 16);
  p.groups = in.cpitch / 8;
  p.c_valid = in.C;
  p.in = in.base;
  p.ko = env_int("MIMO_FLAT2_KO", 0);
  p.trace = nullptr;
  if (env_int("MIMO_FLAT2_TRACE", 0)) {
    static long long* trace_buf = nullptr;
    if (!trace_buf) cudaMalloc(&trace_buf, 4 * 64 * 4 * sizeof(long long));
    cudaMemsetAsync(trace_buf, 0, 4 * 64 * 4 * sizeof(long long), stream);
    p.trace = trace_buf;
  }
  MIMO_CHECK(in.cpitch <= 64 && in.c_off == 0, MIMO_ERR_ARG, "conv3x3_flat2: needs a whole-pixel view with <= 64 channels per pixel");
  size_t smem_bytes = 0;
  p.st_bytes = kBlockM * in.cpitch * 2;
  MIMO_CHECK(plan_smem(block_n, CG, sets, p.nc, p.st_bytes, &p.slots, &p.st_slots, &smem_bytes), MIMO_ERR_ARG,
             "conv3x3_flat2: ring of %d chunks does not fit shared memory (block_n=%d)", p.nc, block_n);
  p.epi.block_n = block_n;
  p.epi.cout = cout;
  p.epi.out_cpitch = out_cpitch;
  p.epi.stage_pitch = 0;
  p.epi.stat_rows = conv3x3_stat_rows();
  p.epi.out = out;
  p.epi.stat_sum = stat_sum;
  p.epi.stat_sq = stat_sq;
  p.epi.bias = bias;
  p.epi.relu = relu;

  CUtensorMap tm_w;
  {
    uint64_t dims[3] = {(uint64_t)cin_pitch, (uint64_t)cout, 9};
    uint64_t strides[2] = {(uint64_t)cin_pitch * 2, (uint64_t)cout * cin_pitch * 2};
    uint32_t box[3] = {64, (uint32_t)(block_n / CG), 9};
    int rc = encode_tmap_bf16(&tm_w, wpacked, 3, dims, strides, box, 1);
    if (rc) return rc;
  }
  const int rc = CG == 2 ? dispatch_flat2<2>(block_n, tm_w, p, smem_bytes, grid, stream) : dispatch_flat2<1>(block_n, tm_w, p, smem_bytes, grid, stream);
  if (rc == MIMO_OK && p.trace) {
    // diagnostic only (synchronises!): dump CTA 0's pipeline timeline relative to its first event
    static long long host[4 * 64 * 4];
    cudaDeviceSynchronize();
    cudaMemcpy(host, p.trace, sizeof(host), cudaMemcpyDeviceToHost);
    long long t0 = 0;
    for (int i = 0; i < 4 * 64 * 4; ++i) if (host[i] > 0 && (t0 == 0 || host[i] < t0)) t0 = host[i];
    static int dumps = 0;
    if (dumps++ < 1) {
      fprintf(stderr, "# flat2 trace (cycles, CTA 0): i | bulk: top slot_free | relayout: top staged slot_dead published | mma: top acc_free chunks_ok issued | epi: top acc_full released done\n");
      for (int i = 0; i < 48; ++i) {
        fprintf(stderr, "%3d | %7lld %7lld |", i, host[(3 * 64 + i) * 4] - t0, host[(3 * 64 + i) * 4 + 1] - t0);
        for (int role = 0; role < 3; ++role)
          fprintf(stderr, " %7lld %7lld %7lld %7lld |", host[(role * 64 + i) * 4] - t0, host[(role * 64 + i) * 4 + 1] - t0,
                  host[(role * 64 + i) * 4 + 2] - t0, host[(role * 64 + i) * 4 + 3] - t0);
        fprintf(stderr, "\n");
      }
    }
  }
  return rc;
}

}  // namespace mimo
