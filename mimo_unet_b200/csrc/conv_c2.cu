// CTA-pair ("c2") tcgen05 implicit-GEMM 3x3 convolution: forward and input-gradient of every layer with more than 64
// input or output channels (the shared core of the reference U-Net, model.py:190-243, ~72 % of its FLOPs), and of the
// small-channel layers where it beats conv3x3_flat_kernel (maps up to 126 columns wide, or > 32 channels: prefer_c2 in conv_igemm.cu).
//
// Replaces: nn.Conv2d(k=3, padding=1, padding_mode="reflect") forward (reference components.py:23,26) and its cuDNN dgrad.
//
// Why a pair: measured on B200 (profiles/r02_umma_probe.txt) a single-CTA M128 x N x K16 MMA costs 43 + N/2 cycles, i.e. the
// tensor pipe idles 43 cycles per instruction (75 % of peak at N = 256, 53 % at N = 96); the cta_group::2 M256 instruction
// costs max(N/2, 40): full rate from N = 96 on. Besides, each CTA of the pair stages only HALF of the weight tile.
// Why chunk-level A segments: knock-outs (profiles/r02_findings.md) show the kernel is bound by the bytes TMA moves from
// L2 into shared memory (~20 B/cycle/SM for data that is not requested by many SMs at once), not by the tensor pipe.
//
//   GEMM view   D[position, cout] = sum_{tap, cin} A[position + tap shift, cin] * Wp[tap][cout][cin]
//   addressing  flat (conv_flat.cu): the haloed NHWC buffer [N][H+2][W+2][C] is ONE list of pixel rows, a tile is 128
//               consecutive positions, tap (kh, kw) is the row offset kh*(W+2) + kw; positions in the halo columns / rows
//               compute garbage that is not stored. An item is 2*T consecutive tiles: CTA r of the pair owns tiles
//               [r*T, (r+1)*T) and the pair's M = 256 MMA multiplies tile t of both CTAs at once.
//   A ring      per 64-channel chunk ONE contiguous segment of T*128 + 2*(W+2) + 2 rows per CTA: all nine taps of all T tiles
//               are row-shifted UMMA windows into it (SWIZZLE_128B is absolute-address based), so every input row is staged
//               (T*128 + 2*(W+2) + 2) / (T*128) times per chunk instead of three times (one segment per kernel row).
//   B ring      per (chunk, kh) the three kw taps x (block_n / 2) weight rows per CTA, shared by the T tiles -- or, for layers with
//               at most two K chunks, the whole weight slab of the pair's n-tile RESIDENT in shared memory for the kernel's lifetime
//               (b_resident: one load, no per-block barrier / commit; the launcher keeps the number of pairs a multiple of the n-tiles).
//   pipeline    both CTAs' TMA loads signal the LEADER's full barriers (cp.async.bulk.tensor ... .cta_group::2); the leader's
//               elected thread issues tcgen05.mma.cta_group::2 and frees ring slots in both CTAs with multicast commits;
//               accumulators are double buffered in TMEM (2 x T x block_n columns), the epilogue warps of both CTAs hand
//               them back with a (remote) arrive on the leader's barrier.
//   epilogue    8 warps per CTA (two per TMEM lane quarter: even / odd 16-column chunks), chunk-outer / tile-inner: tcgen05.ld of two
//               tiles in flight -> bf16 -> 16-byte global stores straight from registers; BatchNorm sum / sum-of-squares of the
//               stored values accumulated over the T tiles in registers, ONE transposed warp butterfly per chunk into per-quarter
//               shared-memory accumulators (fixed order -> deterministic), one partial row per CTA at the end.
//   variants    template <DIAG, EPI>: the knock-outs / cycle trace (MIMO_C2_KO, MIMO_C2_TRACE=2) exist only in the DIAG instantiations
//               (in a one-warp role every parameter test is a constant-bank load + branch on the critical path); EPI = 0 plain store
//               (dgrad), 1 + statistics (training fprop), 2 generic, 3 fused inference epilogue.
//   plan        n-tiling, streamed / resident weights, T and the segment mode are chosen per launch by a cost model in cycles
//               (conv3x3_c2_launch).
// mode 0 (fprop): in = pad==1 view, output dense [N][H][W]; mode 1 (dgrad): in = pad==2 (zero tail) view of dY, output =
// padded-domain gradient [N][H+2][W+2].
#include "common.cuh"
#include "conv_epilogue.cuh"
#include "ops.h"

#include <stdio.h>
#include <stdlib.h>

namespace mimo {
namespace {

constexpr int kBlockM = 128;                     // positions per tile
constexpr int kThreads = 320;                    // warp 0 TMA, warp 1 TMEM + MMA (leader only), warps 2..9 epilogue
constexpr int kMaxASlots = 4;
constexpr int kMaxBSlots = 8;

struct C2Params {
  int wb, img_pix;
  long long total_pos;
  int origin;
  int out_h, out_w, n_img;
  int T;                              // tiles per CTA per item
  int acc_stages;                     // 2: two TMEM accumulator sets of T tiles (epilogue overlaps the next item's MMAs); 1: one set
  int n_mitems, n_tiles_n, block_n;   // block_n = N of the pair's MMA (multiple of 16); each CTA stages block_n / 2 weight rows
  int cin_chunks, ks_last;            // 64-channel chunks; 16-channel k-steps that carry data in the last chunk
  int split, ks_split;                // virtual concat: chunks [0, split) come from the first input view (k-steps of its last chunk), the rest from the second
  int seg_kh;                         // 0: one A segment per chunk serves all nine taps (T*128 + 2*wb + 2 rows); 1: one per (chunk, kh)
                                      //    (T*128 + 2 rows, three kw taps): cheaper when 2*wb is large against T*128 (wide maps)
  int seg_rows, seg_full, seg_rem;    // A segment rows = seg_full boxes of 128 rows + one box of seg_rem rows
  int a_slots, a_slot_bytes, b_slots, b_slot_bytes, b_tap_bytes;
  int b_resident;                     // 1: all 3 * cin_chunks weight blocks of the pair's n-tile stay in shared memory for the whole kernel
                                      //    (thin-K layers: no weight streaming, no per-block barrier / commit); b_slots = 3 * cin_chunks
  int acc_cols;                       // n_tiles_n * block_n: columns of the per-warp statistics accumulators
  long long* trace;                   // diagnostic (env MIMO_C2_TRACE=2): per-CTA cycle counters, 16 per CTA
  int ko;                             // diagnostic knock-outs (env MIMO_C2_KO): 1 no MMAs, 2 no A loads, 4 no B loads, 8 no stats, 16 no stores, 32 no TMEM loads
  EpiArgs epi;
};

// DIAG: knock-out bits and cycle trace compiled in (diagnostic launches only: every ko test in the single-warp roles is a
// constant-bank load + branch on the critical path). EPI: 0 plain store (dgrad), 1 store + BatchNorm statistics (training fprop),
// 2 generic (bias / affine, ReLU, Dropout2d factor, halo destination, optional statistics), 3 fused inference epilogue (16-byte aligned
// affine vectors, ReLU, Dropout2d factor, halo destination, no statistics).
template <bool DIAG, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv3x3_c2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a2,
                  const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_b2,
                  const __grid_constant__ CUtensorMap tmap_w, const C2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + (size_t)p.a_slots * p.a_slot_bytes;
  float* smem_acc = reinterpret_cast<float*>(smem_b + (size_t)p.b_slots * p.b_slot_bytes);   // [4 warps][2][acc_cols]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_acc + 8 * p.acc_cols);
  uint64_t* a_full = bars;                            // leader only
  uint64_t* a_empty = a_full + kMaxASlots;            // per CTA (multicast commit)
  uint64_t* b_full = a_empty + kMaxASlots;            // leader only
  uint64_t* b_empty = b_full + kMaxBSlots;            // per CTA (multicast commit)
  uint64_t* tmem_full = b_empty + kMaxBSlots;         // per CTA (multicast commit)
  uint64_t* tmem_empty = tmem_full + 2;               // leader only: 8 epilogue warps x 2 CTAs
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int ko = DIAG ? p.ko : 0;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  // diagnostic: ko bit 64 = polling waits on the ring barriers, bit 128 = polling waits on the accumulator hand-over barriers
  auto WAIT = [&](uint64_t* bar, uint32_t ph) { if (ko & 64) mbar_wait_poll(bar, ph); else mbar_wait(bar, ph); };
  auto WAITX = [&](uint64_t* bar, uint32_t ph) { if (ko & 128) mbar_wait_poll(bar, ph); else mbar_wait(bar, ph); };
  const bool tracing = DIAG && p.trace != nullptr;
  long long tw0 = 0, tw1 = 0, tw2 = 0, twork = 0;
  unsigned long long tcommit = 0, tidx = 0, tarrive = 0;   // cycles spent in the role's waits / work (lane 0 of each role reports them)
  const long long t_begin = tracing ? clock64() : 0;
  auto TW = [&](long long& acc, uint64_t* bar, uint32_t ph, bool x) {
    const long long t0 = tracing ? clock64() : 0;
    if (x) WAITX(bar, ph); else WAIT(bar, ph);
    if (tracing) acc += clock64() - t0;
  };
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_items = p.n_mitems * p.n_tiles_n;
  const int T = p.T;
  const uint32_t cols = (uint32_t)(p.acc_stages * T * p.block_n);   // accumulators of T tiles, double-buffered when 2 sets fit in 512 columns
  const uint32_t tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_a2);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_b2);
    prefetch_tmap(&tmap_w);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.a_slots; ++s) {
        mbar_init(&a_full[s], 1);
        mbar_init(&a_empty[s], 1);
      }
      const int b_bars = p.b_resident ? 1 : p.b_slots;
      for (int s = 0; s < b_bars; ++s) {
        mbar_init(&b_full[s], 1);
        mbar_init(&b_empty[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tmem_full[a], 1);
        mbar_init(&tmem_empty[a], 16);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc2(tmem_ptr, tmem_cols);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before any remote arrive / TMA completion targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; completion on the leader's barriers) =====================
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    const uint32_t a_tx = (ko & 2) ? 0u : (uint32_t)p.seg_rows * 128u;   // per CTA
    const uint32_t b_tx = (ko & 4) ? 0u : 3u * (uint32_t)p.b_tap_bytes;
    const uint32_t a_full0 = mapa_shared(smem_u32(a_full), 0);
    const uint32_t b_full0 = mapa_shared(smem_u32(b_full), 0);
    const int b_rows = p.block_n >> 1;
    if (p.b_resident && pair < n_items) {
      // the launcher makes n_pairs a multiple of n_tiles_n: every item of this pair has the same n-tile -> one weight load per kernel
      const int co0 = (pair % p.n_tiles_n) * p.block_n + (int)rank * b_rows;
      if (elect_one()) {
        if (rank == 0) mbar_arrive_expect_tx(&b_full[0], 2u * b_tx * 3u * (uint32_t)p.cin_chunks);
        if (!(ko & 4))
          for (int cc = 0; cc < p.cin_chunks; ++cc)
            for (int kh = 0; kh < 3; ++kh)
              tma_load_3d_cg2(&tmap_w, b_full0, smem_b + (size_t)(cc * 3 + kh) * p.b_slot_bytes, cc * 64, co0, kh * 3);
      }
      __syncwarp();
    }
    for (int it = pair; it < n_items; it += n_pairs) {
      const int mi = it / p.n_tiles_n, nt = it - mi * p.n_tiles_n;
      const long long p0 = ((long long)mi * 2 + rank) * (T * kBlockM) + p.origin;
      const int co0 = nt * p.block_n + (int)rank * b_rows;
      for (int cc = 0; cc < p.cin_chunks; ++cc) {
        for (int kh = 0; kh < 3; ++kh) {
          if (p.seg_kh || kh == 0) {
            TW(tw0, &a_empty[sa], pa ^ 1, false);
            if (elect_one()) {
              uint8_t* st = smem_a + (size_t)sa * p.a_slot_bytes;
              const uint32_t fb = a_full0 + (uint32_t)sa * 8u;
              if (rank == 0) mbar_arrive_expect_tx(&a_full[sa], 2u * a_tx);
              if (!(ko & 2)) {
                const bool second = cc >= p.split;   // virtual concat: the chunk lives in the second input view
                const CUtensorMap* m1 = second ? &tmap_b : &tmap_a;
                const CUtensorMap* m2 = second ? &tmap_b2 : &tmap_a2;
                const int ch0 = (second ? cc - p.split : cc) * 64;
                const int r0 = (int)p0 + (p.seg_kh ? kh * p.wb : 0);
                for (int i = 0; i < p.seg_full; ++i) tma_load_2d_cg2(m1, fb, st + (size_t)i * (kBlockM * 128), ch0, r0 + i * kBlockM);
                if (p.seg_rem) tma_load_2d_cg2(m2, fb, st + (size_t)p.seg_full * (kBlockM * 128), ch0, r0 + p.seg_full * kBlockM);
              }
            }
            __syncwarp();
            if (++sa == p.a_slots) { sa = 0; pa ^= 1; }
          }
          if (p.b_resident) continue;
          TW(tw1, &b_empty[sb], pb ^ 1, false);
          if (elect_one()) {
            if (rank == 0) mbar_arrive_expect_tx(&b_full[sb], 2u * b_tx);
            if (!(ko & 4)) tma_load_3d_cg2(&tmap_w, b_full0 + (uint32_t)sb * 8u, smem_b + (size_t)sb * p.b_slot_bytes, cc * 64, co0, kh * 3);
          }
          __syncwarp();
          if (++sb == p.b_slots) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===================== MMA issuer (leader CTA, warp-uniform loop, one elected lane) =====================
      const uint32_t idesc = make_idesc_bf16(2 * kBlockM, p.block_n, 0, 0);
      constexpr uint32_t hi = desc_hi(1024, kLayoutSW128);
      const uint32_t a_lo0 = desc_lo(smem_u32(smem_a), 16);
      const uint32_t b_lo0 = desc_lo(smem_u32(smem_b), 16);
      const uint32_t a_step = (uint32_t)p.a_slot_bytes >> 4, b_step = (uint32_t)p.b_slot_bytes >> 4;
      const uint32_t b_tap = (uint32_t)p.b_tap_bytes >> 4;
      const uint32_t bn = (uint32_t)p.block_n;
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      uint32_t n = 0;
      const bool resident = p.b_resident != 0;
      if (resident && pair < n_items) {
        TW(tw1, &b_full[0], 0, false);
        tc_fence_after();
      }
      for (int it = pair; it < n_items; it += n_pairs, ++n) {
        const uint32_t buf = p.acc_stages == 2 ? (n & 1u) : 0u;
        TW(tw2, &tmem_empty[buf], ((p.acc_stages == 2 ? (n >> 1) : n) & 1u) ^ 1u, true);
        tc_fence_after();
        const uint32_t d0 = tmem_base + buf * (uint32_t)T * bn;
        for (int cc = 0; cc < p.cin_chunks; ++cc) {
          if (!p.seg_kh) TW(tw0, &a_full[sa], pa, false);
          uint32_t a_lo = a_lo0 + (uint32_t)sa * a_step;
          const uint32_t ks = (cc == p.cin_chunks - 1) ? (uint32_t)p.ks_last : (cc == p.split - 1 ? (uint32_t)p.ks_split : 4u);
          for (int kh = 0; kh < 3; ++kh) {
            if (p.seg_kh) {
              TW(tw0, &a_full[sa], pa, false);
              a_lo = a_lo0 + (uint32_t)sa * a_step;
            }
            if (!resident) TW(tw1, &b_full[sb], pb, false);
            tc_fence_after();
            const long long tq0 = tracing ? clock64() : 0;
            const uint32_t b_lo = b_lo0 + (uint32_t)(resident ? cc * 3 + kh : sb) * b_step;
            const uint32_t a_kh = a_lo + (p.seg_kh ? 0u : (uint32_t)(kh * p.wb) * 8u);   // 128 B per row = 8 descriptor units
            if (elect_one()) {
              for (int t = 0; t < T; ++t) {
                const uint32_t a_t = a_kh + (uint32_t)t * (kBlockM * 8u);
                const uint32_t d_t = d0 + (uint32_t)t * bn;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    if ((uint32_t)k < ks && !(ko & 1))
                      umma2_bf16_w(d_t, a_t + (uint32_t)(kw * 8 + k * 2), hi, b_lo + (uint32_t)kw * b_tap + (uint32_t)(k * 2), hi, idesc,
                                   (cc | kh | kw | k) != 0);
                  }
                }
              }
              const long long tc0 = tracing ? clock64() : 0;
              if (!resident) umma_commit2_mc(&b_empty[sb], 3);   // frees the weight block in both CTAs
              if (p.seg_kh || kh == 2) umma_commit2_mc(&a_empty[sa], 3);    // ... and the A segment after its last use
              if (tracing) tcommit += clock64() - tc0;
            }
            __syncwarp();
            if (tracing) twork += clock64() - tq0;
            if (!resident && ++sb == p.b_slots) { sb = 0; pb ^= 1; }
            if (p.seg_kh && ++sa == p.a_slots) { sa = 0; pa ^= 1; }
          }
          if (!p.seg_kh && ++sa == p.a_slots) { sa = 0; pa ^= 1; }
        }
        if (elect_one()) umma_commit2_mc(&tmem_full[buf], 3);
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (8 warps per CTA, each CTA drains its own 128 accumulator rows per tile) =====================
    // Two warps share a TMEM lane quarter and split the 16-column chunks (even / odd). A warp walks chunk-outer, tile-inner: the T
    // loads of a chunk are in flight together, the per-channel vectors are read once per chunk, and the BatchNorm statistics are
    // reduced across lanes once per chunk for all T tiles (the 32 shuffles of the two transposed butterflies were the largest part of
    // the drain: ~900 cycles per chunk and tile with one warp per quarter, measured with MIMO_C2_TRACE=2).
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int grp = (warp - 2) >> 2;         // 0: even chunks, 1: odd chunks
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const EpiArgs& e = p.epi;
    const bool stats = (EPI == 1 || (EPI == 2 && e.stat_sum != nullptr)) && !(ko & 8);
    // loop invariants in registers (the parameter struct lives in the constant bank)
    const int block_n = p.block_n, wb = p.wb, img_pix = p.img_pix, out_h = p.out_h, out_w = p.out_w, out_cpitch = e.out_cpitch;
    const int out_cmax = EPI >= 2 ? e.out_cmax : out_cpitch;
    const int halo = EPI >= 2 ? e.halo : 0;
    const long long total_pos = p.total_pos;
    bf16* const out = e.out;
    const int n_tiles_n = p.n_tiles_n, acc_stages = p.acc_stages;
    const int nchunks = block_n >> 4;
    float* wacc_s = smem_acc + (size_t)q * 2 * p.acc_cols;   // shared by the two warps of the quarter (disjoint columns)
    float* wacc_q = wacc_s + p.acc_cols;
    for (int i = et; i < 8 * p.acc_cols; i += 256) smem_acc[i] = 0.f;
    named_bar_sync(1, 256);
    const uint32_t tempty0 = mapa_shared(smem_u32(tmem_empty), 0);
    uint32_t n = 0;
    for (int it = pair; it < n_items; it += n_pairs, ++n) {
      const uint32_t buf = acc_stages == 2 ? (n & 1u) : 0u;
      const int mi = it / n_tiles_n, nt = it - mi * n_tiles_n;
      const int co0 = nt * block_n;
      // output pixel of this thread's row in each of the T tiles
      const long long ti0 = tracing ? clock64() : 0;
      bf16* dst[4];
      const float* drow[4];
      float vmask[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        dst[t] = nullptr; drow[t] = nullptr; vmask[t] = 0.f;
        if (t < T) {
          const long long pos = (((long long)mi * 2 + rank) * T + t) * kBlockM + row;
          if (pos < total_pos) {
            const unsigned up = (unsigned)pos;
            const unsigned ni = up / (unsigned)img_pix;
            const unsigned rem = up - ni * (unsigned)img_pix;
            const unsigned hp = rem / (unsigned)wb, wp = rem - hp * (unsigned)wb;
            if ((int)hp < out_h && (int)wp < out_w) {
              const size_t pix = halo ? ((size_t)ni * (out_h + 2) + hp + 1) * (out_w + 2) + wp + 1 : ((size_t)ni * out_h + hp) * out_w + wp;
              dst[t] = out + pix * out_cpitch + co0;
              vmask[t] = 1.f;
              if (EPI >= 2 && e.drop != nullptr) drow[t] = e.drop + (size_t)ni * e.cout;
            }
          }
        }
      }
      if (tracing) tidx += clock64() - ti0;
      TW(tw2, &tmem_full[buf], (acc_stages == 2 ? (n >> 1) : n) & 1u, true);
      tc_fence_after();
      const long long tq0 = tracing ? clock64() : 0;
      const uint32_t t_addr0 = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)T * (uint32_t)block_n;
#pragma unroll 1
      for (int j = grp; j < nchunks; j += 2) {
        const int c0 = co0 + j * 16;
        // per-channel vectors of this chunk (fused inference epilogue / bias): vec4 is set by the launcher when scale / shift hold
        // round_up(cout, 16) floats per n-tile column range (the executor's BatchNorm vectors are zero-padded up to the channel pitch)
        float sc[16], sh[16];
        const bool affine = EPI == 3 || (EPI == 2 && e.scale != nullptr);
        const bool biased = EPI == 2 && !affine && e.bias != nullptr;
        const bool relu = EPI >= 2 && (e.relu & 1);
        if (affine) {
          if (EPI == 3 || (e.relu & 2)) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(e.scale + c0 + i));
              const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + c0 + i));
              sc[i] = a.x; sc[i + 1] = a.y; sc[i + 2] = a.z; sc[i + 3] = a.w;
              sh[i] = b.x; sh[i + 1] = b.y; sh[i + 2] = b.z; sh[i + 3] = b.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              sc[i] = (c0 + i < e.cout) ? __ldg(e.scale + c0 + i) : 0.f;
              sh[i] = (c0 + i < e.cout) ? __ldg(e.bias + c0 + i) : 0.f;
            }
          }
        } else if (biased) {
#pragma unroll
          for (int i = 0; i < 16; ++i) sh[i] = (c0 + i < e.cout) ? __ldg(e.bias + c0 + i) : 0.f;
        }
        float ssum[16], ssq[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
        // two tiles at a time: their TMEM loads are in flight together (four would cost 32 more registers: spills at the 168 cap)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          uint32_t raw[2][16];
          if ((t & 1) == 0 && t < T && !(ko & 32)) {
            tmem_ld16_nowait(t_addr0 + (uint32_t)t * (uint32_t)block_n + j * 16, raw[0]);
            if (t + 1 < T) tmem_ld16_nowait(t_addr0 + (uint32_t)(t + 1) * (uint32_t)block_n + j * 16, raw[1]);
            tmem_ld_wait();
          }
          if (t < T) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = (ko & 32) ? (float)(i + j) : __uint_as_float(raw[t & 1][i]);
            if (affine) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], sc[i], sh[i]);
            } else if (biased) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += sh[i];
            }
            if (relu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            if (EPI >= 2 && drow[t] != nullptr) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] *= (c0 + i < e.cout) ? __ldg(drow[t] + c0 + i) : 0.f;
            }
            const uint4 lo = pack8(v), hi8 = pack8(v + 8);
            if (dst[t] != nullptr && !(ko & 16)) {
              if (c0 < out_cmax) *reinterpret_cast<uint4*>(dst[t] + j * 16) = lo;
              if (c0 + 8 < out_cmax) *reinterpret_cast<uint4*>(dst[t] + j * 16 + 8) = hi8;
            }
            if (stats) {
              // statistics of the values as stored (bf16); rows that are not stored count as 0
              float f[16];
              unpack8(lo, f);
              unpack8(hi8, f + 8);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float x = f[i] * vmask[t];
                ssum[i] += x;
                ssq[i] = fmaf(x, x, ssq[i]);
              }
            }
          }
        }
        if (stats) {
          const float cs = warp_colsum16(ssum, lane);
          const float cq = warp_colsum16(ssq, lane);
          // lane l holds column (l >> 1): even lanes keep the sums, odd lanes the squares (accumulators private to the quarter's chunk owner)
          const int col = c0 + (lane >> 1);
          if ((lane & 1) == 0) wacc_s[col] += cs;
          else wacc_q[col] += cq;
        }
      }
      const long long ta0 = tracing ? clock64() : 0;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty0 + buf * 8u);
      if (tracing) { twork += ta0 - tq0; tarrive += clock64() - ta0; }
    }
    if (stats) {
      // combine the four quarters (fixed order) and write this CTA's partial row
      named_bar_sync(1, 256);
      for (int col = et; col < e.out_cpitch; col += 256) {
        float s_ = 0.f, q_ = 0.f;
        if (col < p.acc_cols) {
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            s_ += smem_acc[(size_t)w * 2 * p.acc_cols + col];
            q_ += smem_acc[(size_t)w * 2 * p.acc_cols + p.acc_cols + col];
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

  if (tracing && lane == 0 && warp <= 2) {
    long long* tr = p.trace + (size_t)blockIdx.x * 16 + warp * 5;
    tr[0] = clock64() - t_begin; tr[1] = tw0; tr[2] = tw1; tr[3] = tw2; tr[4] = twork;
    if (warp == 2) { p.trace[(size_t)blockIdx.x * 16 + 15] = (long long)tidx; p.trace[(size_t)blockIdx.x * 16 + 11] = (long long)tarrive; }
  }
  if (tracing && warp == 1 && tcommit) atomicAdd(reinterpret_cast<unsigned long long*>(p.trace + (size_t)blockIdx.x * 16 + 12), tcommit);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA frees its tensor memory / exits while the pair's MMAs or remote arrives may still target it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, tmem_cols);
  }
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// n-tiling of the output channels: `split` = 1 is the coarsest tiling (fewest n-tiles with block_n <= 256), 2 halves the tiles
void plan_n(int cout, int split, int* block_n, int* n_tiles) {
  const int c16 = round_up(cout, 16);
  *n_tiles = ceil_div(c16, 256) * split;
  *block_n = round_up(ceil_div(c16, *n_tiles), 16);
}

// shared-memory plan for T tiles per CTA; returns false when it does not fit
bool plan_smem(int T, int seg_kh, int wb, int block_n, int acc_cols, int resident_blocks, C2Params* p, size_t* smem_bytes) {
  p->T = T;
  p->seg_kh = seg_kh;
  p->seg_rows = T * kBlockM + (seg_kh ? 2 : 2 * wb + 2);
  p->seg_full = p->seg_rows / kBlockM;
  p->seg_rem = p->seg_rows % kBlockM;
  p->a_slot_bytes = round_up(p->seg_rows * 128, 1024);
  p->b_tap_bytes = (block_n / 2) * 128;
  p->b_slot_bytes = round_up(3 * p->b_tap_bytes, 1024);
  p->b_resident = resident_blocks > 0 ? 1 : 0;
  const int fixed = 8 * acc_cols * 4 + (2 * kMaxASlots + 2 * kMaxBSlots + 4) * 8 + 16 + 64 + 1024;
  const int budget = 227 * 1024 - fixed;
  int a_slots = seg_kh ? 3 : 2, b_slots = 3;
  if (resident_blocks > 0) {
    // every weight block of the n-tile resident; the rest of the memory is the A ring
    b_slots = resident_blocks;
    if (a_slots * p->a_slot_bytes + b_slots * p->b_slot_bytes > budget) return false;
    while (a_slots < kMaxASlots && (a_slots + 1) * p->a_slot_bytes + b_slots * p->b_slot_bytes <= budget) ++a_slots;
  } else {
    // at least two A segments and three weight blocks in flight; then weight blocks up to 6, then a third A segment
    if (a_slots * p->a_slot_bytes + b_slots * p->b_slot_bytes > budget) return false;
    while (b_slots < 6 && a_slots * p->a_slot_bytes + (b_slots + 1) * p->b_slot_bytes <= budget) ++b_slots;
    while (a_slots < (seg_kh ? kMaxASlots : 3) && (a_slots + 1) * p->a_slot_bytes + b_slots * p->b_slot_bytes <= budget) ++a_slots;
    while (b_slots < kMaxBSlots && a_slots * p->a_slot_bytes + (b_slots + 1) * p->b_slot_bytes <= budget) ++b_slots;
  }
  p->a_slots = a_slots;
  p->b_slots = b_slots;
  *smem_bytes = (size_t)a_slots * p->a_slot_bytes + (size_t)b_slots * p->b_slot_bytes + fixed;
  return true;
}

}  // namespace

bool conv3x3_c2_ok(const ActView& in, int mode, int cout) {
  static const int enabled = env_int("MIMO_CONV_C2", 1);
  if (!enabled) return false;
  if (mode == 0 && in.pad != 1) return false;
  if (mode == 1 && in.pad != 2) return false;
  const long long total_pos = (long long)in.N * in.hb() * in.wb();
  if (total_pos >= (1ll << 31) - 65536) return false;
  int block_n, n_tiles;
  plan_n(cout, 1, &block_n, &n_tiles);
  C2Params p{};
  size_t smem_bytes;
  return plan_smem(1, 0, in.wb(), block_n, block_n * n_tiles, 0, &p, &smem_bytes) || plan_smem(1, 1, in.wb(), block_n, block_n * n_tiles, 0, &p, &smem_bytes);
}

int conv3x3_c2_launch(const ActView& in, int mode, const bf16* wpacked, int cout, int cin_pitch, bf16* out, int out_cpitch,
                      float* stat_sum, float* stat_sq, const float* bias, int relu, cudaStream_t stream, const ConvFuse* fuse,
                      const ActView* in2) {
  note_kernel(7);
  C2Params p{};
  p.wb = in.wb();
  p.img_pix = in.hb() * in.wb();
  p.total_pos = (long long)in.N * p.img_pix;
  p.origin = mode == 0 ? 0 : -(2 * p.wb + 2);
  p.out_h = mode == 0 ? in.H : in.H + 2;
  p.out_w = mode == 0 ? in.W : in.W + 2;
  p.n_img = in.N;
  { static const int ko = env_int("MIMO_C2_KO", 0); p.ko = ko; }
  p.cin_chunks = ceil_div(in.C, 64);
  p.ks_last = ceil_div(in.C - (p.cin_chunks - 1) * 64, 16);
  p.split = p.cin_chunks;
  p.ks_split = p.ks_last;
  if (in2 != nullptr) {
    MIMO_CHECK(in2->N == in.N && in2->H == in.H && in2->W == in.W && in2->pad == in.pad, MIMO_ERR_ARG, "conv3x3_c2: the two views of a virtual concat differ in shape");
    MIMO_CHECK(in2->cpitch % 8 == 0 && in2->c_off % 8 == 0 && ((uintptr_t)in2->base % 16) == 0, MIMO_ERR_ALIGN, "conv3x3_c2: second view not 16-byte aligned");
    const int c2chunks = ceil_div(in2->C, 64);
    p.cin_chunks = p.split + c2chunks;
    p.ks_last = ceil_div(in2->C - (c2chunks - 1) * 64, 16);
    MIMO_CHECK(cin_pitch >= p.split * 64 + in2->C, MIMO_ERR_ARG, "conv3x3_c2: weight pitch %d too small for the chunk-aligned virtual concat", cin_pitch);
  }
  // Plan: n-tiling (coarsest, or halved so that a thin-K layer's weights fit), weights streamed through a ring or RESIDENT, tiles per
  // CTA T (more tiles share one weight block and one segment halo: fewer bytes through TMA per FLOP, but coarser items: wave
  // quantisation over the 74 CTA pairs), A segment mode. Pick the combination with the lowest modelled cost.
  const int max_pairs = num_sms() / 2;
  size_t smem_bytes = 0;
  {
    static const int forced_t = env_int("MIMO_C2_T", 0);
    static const int forced_kh = env_int("MIMO_C2_SEGKH", -1);
    static const int allow_single = env_int("MIMO_C2_SINGLE", 0);
    static const int allow_resident = env_int("MIMO_C2_RESIDENT", 1);
    struct Choice { int T, kh, stages, split, resident; };
    Choice bestc{0, 0, 2, 1, 0};
    double best = 1e300;
    const double k_steps = 4.0 * (p.cin_chunks - 1 - (in2 != nullptr ? 1 : 0)) + p.ks_last + (in2 != nullptr ? p.ks_split : 0);
    for (int split = 1; split <= 2; ++split) {
      int block_n, n_tiles;
      plan_n(cout, split, &block_n, &n_tiles);
      if ((block_n / 2) % 8 != 0 || (n_tiles - 1) * block_n >= out_cpitch) continue;
      for (int resident = 0; resident <= (allow_resident ? 1 : 0); ++resident) {
        if (split == 2 && !resident) continue;                 // halving the n-tiles only pays when it makes the weights resident
        if (resident && max_pairs % n_tiles != 0) continue;   // a pair must keep one n-tile for the whole kernel
        // measured (tools/gpu/r2_c2.sh): with three or more K chunks the resident weights leave room for T = 1 only and the per-item
        // overheads eat the saved weight traffic (168 -> 84 @ 64x80: 99 us streamed, 106 us resident)
        if (resident && p.cin_chunks > 2) continue;
        for (int T = 1; T <= 4; T *= 2) {
          if (forced_t && T != forced_t) continue;
          if (T * block_n > 512) continue;
          // two accumulator sets (the epilogue of one item overlaps the MMAs of the next) when they fit in the 512 TMEM columns
          const int stages = 2 * T * block_n <= 512 ? 2 : 1;
          if (stages == 1 && !allow_single) continue;
          for (int kh = 0; kh < 2; ++kh) {
            if (forced_kh >= 0 && kh != forced_kh) continue;
            C2Params q = p;
            size_t sb;
            if (!plan_smem(T, kh, p.wb, block_n, block_n * n_tiles, resident ? 3 * p.cin_chunks : 0, &q, &sb)) continue;
            const long long items = ceil_div_ll(p.total_pos, 2ll * T * kBlockM) * n_tiles;
            const long long waves = ceil_div_ll(items, max_pairs);
            const double a_rows = kh ? 3.0 * q.seg_rows : (double)q.seg_rows;
            const double b_rows = 9.0 * (block_n / 2);
            // cycles per item: TMA delivers ~20 B/cycle/SM (6.4 cycles per 128-byte row), a pair MMA costs max(N/2, 40) cycles, a
            // tcgen05.commit ~400 cycles of the issuing thread, the epilogue ~5 cycles per accumulator column and tile (exposed only
            // with a single accumulator set)
            // ... plus ~300 cycles per TMA instruction / barrier round trip, and ~2000 cycles per item for the accumulator hand-over
            // and the pipeline ramp (without them T = 1 ties with T = 4 on wide maps, where the per-(chunk, kh) segments have no halo
            // to share: measured 352 us vs 234 us on C3's 90 -> 45 @ 256 x 256)
            const double n_loads = p.cin_chunks * ((kh ? 3.0 : 1.0) * (q.seg_full + (q.seg_rem ? 1 : 0)) + (resident ? 0.0 : 3.0));
            const double load = 6.4 * p.cin_chunks * (a_rows + (resident ? 0.0 : b_rows)) + 300.0 * n_loads;
            const double mma = (double)T * 9.0 * k_steps * (block_n / 2 > 40 ? block_n / 2 : 40);
            const double issue = 400.0 * (p.cin_chunks * ((kh ? 3.0 : 1.0) + (resident ? 0.0 : 3.0)) + 1.0) + (double)T * 9.0 * k_steps * 16.0;
            const double epi = stages == 1 ? 5.0 * T * block_n : 0.0;
            double item = load > mma ? load : mma;
            if (issue > item) item = issue;
            const double cost = (double)waves * (item + epi + 2000.0) + (resident ? 6.4 * p.cin_chunks * b_rows : 0.0);
            if (cost < best) { best = cost; bestc = Choice{T, kh, stages, split, resident}; }
          }
        }
      }
    }
    MIMO_CHECK(bestc.T > 0, MIMO_ERR_ARG, "conv3x3_c2: not enough shared memory for cout=%d, row pitch %d", cout, p.wb);
    plan_n(cout, bestc.split, &p.block_n, &p.n_tiles_n);
    p.acc_cols = p.block_n * p.n_tiles_n;
    plan_smem(bestc.T, bestc.kh, p.wb, p.block_n, p.acc_cols, bestc.resident ? 3 * p.cin_chunks : 0, &p, &smem_bytes);
    p.acc_stages = bestc.stages;
    static const int trace = env_int("MIMO_C2_TRACE", 0);
    if (trace)
      fprintf(stderr, "c2 plan: mode %d cin %d cout %d pos %lld wb %d -> block_n %d x %d, T %d, seg_kh %d, acc_stages %d, resident %d, a_slots %d b_slots %d\n",
              mode, in.C, cout, p.total_pos, p.wb, p.block_n, p.n_tiles_n, p.T, p.seg_kh, p.acc_stages, p.b_resident, p.a_slots, p.b_slots);
  }
  p.n_mitems = (int)ceil_div_ll(p.total_pos, 2ll * p.T * kBlockM);
  p.epi.block_n = p.block_n;
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
  MIMO_CHECK((p.n_tiles_n - 1) * p.block_n < out_cpitch, MIMO_ERR_ARG, "conv3x3_c2: n-tiling exceeds out_cpitch");

  CUtensorMap tm_a, tm_a2, tm_b, tm_b2, tm_w;
  for (int v = 0; v < 2; ++v) {
    const ActView& av = (v == 1 && in2 != nullptr) ? *in2 : in;   // without a second view the maps are duplicates (never read)
    uint64_t dims[2] = {(uint64_t)av.C, (uint64_t)p.total_pos};
    uint64_t strides[1] = {(uint64_t)av.cpitch * 2};
    uint32_t box[2] = {64, (uint32_t)kBlockM};
    int rc = encode_tmap_bf16(v ? &tm_b : &tm_a, av.base + av.c_off, 2, dims, strides, box, 1);
    if (rc) return rc;
    uint32_t box2[2] = {64, (uint32_t)(p.seg_rem ? p.seg_rem : 1)};
    rc = encode_tmap_bf16(v ? &tm_b2 : &tm_a2, av.base + av.c_off, 2, dims, strides, box2, 1);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)cin_pitch, (uint64_t)cout, 9};
    uint64_t strides[2] = {(uint64_t)cin_pitch * 2, (uint64_t)cout * cin_pitch * 2};
    uint32_t box[3] = {64, (uint32_t)(p.block_n / 2), 3};
    int rc = encode_tmap_bf16(&tm_w, wpacked, 3, dims, strides, box, 1);
    if (rc) return rc;
  }
  static const int trace_mode = env_int("MIMO_C2_TRACE", 0);
  const bool diag = p.ko != 0 || trace_mode == 2;
  const int epi = (fuse != nullptr && stat_sum == nullptr) ? 3 : (fuse != nullptr || bias != nullptr || relu != 0) ? 2 : (stat_sum != nullptr ? 1 : 0);
  using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const C2Params);
  static const KernelFn table[2][4] = {
      {conv3x3_c2_kernel<false, 0>, conv3x3_c2_kernel<false, 1>, conv3x3_c2_kernel<false, 2>, conv3x3_c2_kernel<false, 3>},
      {conv3x3_c2_kernel<true, 0>, conv3x3_c2_kernel<true, 1>, conv3x3_c2_kernel<true, 2>, conv3x3_c2_kernel<true, 3>}};
  const KernelFn kernel = table[diag ? 1 : 0][epi];
  MIMO_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const int items = p.n_mitems * p.n_tiles_n;
  const int grid = 2 * (items < max_pairs ? items : max_pairs);   // items and max_pairs are multiples of n_tiles_n in resident mode
  static long long* trace_buf = nullptr;
  if (trace_mode == 2) {   // diagnostic only: synchronous, not graph-capturable
    if (!trace_buf) MIMO_CUDA(cudaMalloc(&trace_buf, 512 * 16 * sizeof(long long)));
    MIMO_CUDA(cudaMemsetAsync(trace_buf, 0, 512 * 16 * sizeof(long long), stream));
    p.trace = trace_buf;
  }
  kernel<<<grid, kThreads, smem_bytes, stream>>>(tm_a, tm_a2, tm_b, tm_b2, tm_w, p);
  MIMO_LAUNCH_CHECK();
  if (trace_mode == 2) {
    static long long host[512 * 16];
    MIMO_CUDA(cudaStreamSynchronize(stream));
    MIMO_CUDA(cudaMemcpy(host, trace_buf, sizeof(host), cudaMemcpyDeviceToHost));
    for (int b : {0, 1, grid / 2, grid / 2 + 1, grid - 2, grid - 1}) {
      const long long* t = host + (size_t)b * 16;
      fprintf(stderr, "c2 trace cta %3d: loader total %lld wait a_empty %lld b_empty %lld | issuer total %lld wait a_full %lld b_full %lld tmem_empty %lld issue %lld | "
                      "epilogue(w2) total %lld wait tmem_full %lld work %lld idx %lld arrive %lld | commits %lld\n", b, t[0], t[1], t[2], t[5], t[6], t[7], t[8], t[9], t[10], t[13], t[14], t[15], t[11], t[12]);
    }
  }
  return MIMO_OK;
}

}  // namespace mimo
