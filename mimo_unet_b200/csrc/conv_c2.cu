// CTA-pair ("c2") tcgen05 implicit-GEMM 3x3 convolution: forward and input-gradient of every layer with more than 64
// input or output channels (the shared core of the reference U-Net, model.py:190-243, ~72 % of its FLOPs).
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
//   B ring      per (chunk, kh) the three kw taps x (block_n / 2) weight rows per CTA, shared by the T tiles.
//   pipeline    both CTAs' TMA loads signal the LEADER's full barriers (cp.async.bulk.tensor ... .cta_group::2); the leader's
//               elected thread issues tcgen05.mma.cta_group::2 and frees ring slots in both CTAs with multicast commits;
//               accumulators are double buffered in TMEM (2 x T x block_n columns), the epilogue warps of both CTAs hand
//               them back with a (remote) arrive on the leader's barrier.
//   epilogue    per CTA its own 128 rows per tile: tcgen05.ld -> bf16 -> 16-byte global stores straight from registers;
//               BatchNorm sum / sum-of-squares of the stored values via the transposed warp butterfly into per-warp
//               shared-memory accumulators (fixed order -> deterministic), one partial row per CTA at the end.
// mode 0 (fprop): in = pad==1 view, output dense [N][H][W]; mode 1 (dgrad): in = pad==2 (zero tail) view of dY, output =
// padded-domain gradient [N][H+2][W+2].
#include "common.cuh"
#include "conv_epilogue.cuh"
#include "ops.h"

#include <stdlib.h>

namespace mimo {
namespace {

constexpr int kBlockM = 128;                     // positions per tile
constexpr int kThreads = 192;                    // warp 0 TMA, warp 1 TMEM + MMA (leader only), warps 2..5 epilogue
constexpr int kMaxASlots = 4;
constexpr int kMaxBSlots = 8;

struct C2Params {
  int wb, img_pix;
  long long total_pos;
  int origin;
  int out_h, out_w, n_img;
  int T;                              // tiles per CTA per item
  int n_mitems, n_tiles_n, block_n;   // block_n = N of the pair's MMA (multiple of 16); each CTA stages block_n / 2 weight rows
  int cin_chunks, ks_last;            // 64-channel chunks; 16-channel k-steps that carry data in the last chunk
  int split, ks_split;                // virtual concat: chunks [0, split) come from the first input view (k-steps of its last chunk), the rest from the second
  int seg_kh;                         // 0: one A segment per chunk serves all nine taps (T*128 + 2*wb + 2 rows); 1: one per (chunk, kh)
                                      //    (T*128 + 2 rows, three kw taps): cheaper when 2*wb is large against T*128 (wide maps)
  int seg_rows, seg_full, seg_rem;    // A segment rows = seg_full boxes of 128 rows + one box of seg_rem rows
  int a_slots, a_slot_bytes, b_slots, b_slot_bytes, b_tap_bytes;
  int acc_cols;                       // n_tiles_n * block_n: columns of the per-warp statistics accumulators
  int ko;                             // diagnostic knock-outs (env MIMO_C2_KO): 1 no MMAs, 2 no A loads, 4 no B loads, 8 no stats, 16 no stores, 32 no TMEM loads
  EpiArgs epi;
};

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
  uint64_t* tmem_empty = tmem_full + 2;               // leader only: 4 epilogue warps x 2 CTAs
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_items = p.n_mitems * p.n_tiles_n;
  const int T = p.T;
  const uint32_t cols = 2u * (uint32_t)(T * p.block_n);   // double-buffered accumulators of T tiles
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
      for (int s = 0; s < p.b_slots; ++s) {
        mbar_init(&b_full[s], 1);
        mbar_init(&b_empty[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tmem_full[a], 1);
        mbar_init(&tmem_empty[a], 8);
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
    const uint32_t a_tx = (p.ko & 2) ? 0u : (uint32_t)p.seg_rows * 128u;   // per CTA
    const uint32_t b_tx = (p.ko & 4) ? 0u : 3u * (uint32_t)p.b_tap_bytes;
    const uint32_t a_full0 = mapa_shared(smem_u32(a_full), 0);
    const uint32_t b_full0 = mapa_shared(smem_u32(b_full), 0);
    const int b_rows = p.block_n >> 1;
    for (int it = pair; it < n_items; it += n_pairs) {
      const int mi = it / p.n_tiles_n, nt = it - mi * p.n_tiles_n;
      const long long p0 = ((long long)mi * 2 + rank) * (T * kBlockM) + p.origin;
      const int co0 = nt * p.block_n + (int)rank * b_rows;
      for (int cc = 0; cc < p.cin_chunks; ++cc) {
        for (int kh = 0; kh < 3; ++kh) {
          if (p.seg_kh || kh == 0) {
            mbar_wait(&a_empty[sa], pa ^ 1);
            if (elect_one()) {
              uint8_t* st = smem_a + (size_t)sa * p.a_slot_bytes;
              const uint32_t fb = a_full0 + (uint32_t)sa * 8u;
              if (rank == 0) mbar_arrive_expect_tx(&a_full[sa], 2u * a_tx);
              if (!(p.ko & 2)) {
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
          mbar_wait(&b_empty[sb], pb ^ 1);
          if (elect_one()) {
            if (rank == 0) mbar_arrive_expect_tx(&b_full[sb], 2u * b_tx);
            if (!(p.ko & 4)) tma_load_3d_cg2(&tmap_w, b_full0 + (uint32_t)sb * 8u, smem_b + (size_t)sb * p.b_slot_bytes, cc * 64, co0, kh * 3);
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
      for (int it = pair; it < n_items; it += n_pairs, ++n) {
        const uint32_t buf = n & 1u;
        mbar_wait(&tmem_empty[buf], ((n >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d0 = tmem_base + buf * (uint32_t)T * bn;
        for (int cc = 0; cc < p.cin_chunks; ++cc) {
          if (!p.seg_kh) mbar_wait(&a_full[sa], pa);
          uint32_t a_lo = a_lo0 + (uint32_t)sa * a_step;
          const uint32_t ks = (cc == p.cin_chunks - 1) ? (uint32_t)p.ks_last : (cc == p.split - 1 ? (uint32_t)p.ks_split : 4u);
          for (int kh = 0; kh < 3; ++kh) {
            if (p.seg_kh) {
              mbar_wait(&a_full[sa], pa);
              a_lo = a_lo0 + (uint32_t)sa * a_step;
            }
            mbar_wait(&b_full[sb], pb);
            tc_fence_after();
            const uint32_t b_lo = b_lo0 + (uint32_t)sb * b_step;
            const uint32_t a_kh = a_lo + (p.seg_kh ? 0u : (uint32_t)(kh * p.wb) * 8u);   // 128 B per row = 8 descriptor units
            if (elect_one()) {
              for (int t = 0; t < T; ++t) {
                const uint32_t a_t = a_kh + (uint32_t)t * (kBlockM * 8u);
                const uint32_t d_t = d0 + (uint32_t)t * bn;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    if ((uint32_t)k < ks && !(p.ko & 1))
                      umma2_bf16_w(d_t, a_t + (uint32_t)(kw * 8 + k * 2), hi, b_lo + (uint32_t)kw * b_tap + (uint32_t)(k * 2), hi, idesc,
                                   (cc | kh | kw | k) != 0);
                  }
                }
              }
              umma_commit2_mc(&b_empty[sb], 3);                 // frees the weight block in both CTAs
              if (p.seg_kh || kh == 2) umma_commit2_mc(&a_empty[sa], 3);    // ... and the A segment after its last use
            }
            __syncwarp();
            if (++sb == p.b_slots) { sb = 0; pb ^= 1; }
            if (p.seg_kh && ++sa == p.a_slots) { sa = 0; pa ^= 1; }
          }
          if (!p.seg_kh && ++sa == p.a_slots) { sa = 0; pa ^= 1; }
        }
        if (elect_one()) umma_commit2_mc(&tmem_full[buf], 3);
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (4 warps per CTA, each CTA drains its own 128 accumulator rows per tile) =====================
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const EpiArgs& e = p.epi;
    const bool stats = e.stat_sum != nullptr && !(p.ko & 8);
    const int nchunks = p.block_n >> 4;
    float* wacc_s = smem_acc + (size_t)q * 2 * p.acc_cols;
    float* wacc_q = wacc_s + p.acc_cols;
    for (int i = lane; i < 2 * p.acc_cols; i += 32) wacc_s[i] = 0.f;
    __syncwarp();
    const uint32_t tempty0 = mapa_shared(smem_u32(tmem_empty), 0);
    uint32_t n = 0;
    for (int it = pair; it < n_items; it += n_pairs, ++n) {
      const uint32_t buf = n & 1u;
      const int mi = it / p.n_tiles_n, nt = it - mi * p.n_tiles_n;
      const int co0 = nt * p.block_n;
      mbar_wait(&tmem_full[buf], (n >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < T; ++t) {
        const long long pos = (((long long)mi * 2 + rank) * T + t) * kBlockM + row;
        bool valid = false;
        size_t my_pix = 0;
        unsigned img = 0;
        if (pos < p.total_pos) {
          const unsigned up = (unsigned)pos;
          const unsigned ni = up / (unsigned)p.img_pix;
          const unsigned rem = up - ni * (unsigned)p.img_pix;
          const unsigned hp = rem / (unsigned)p.wb, wp = rem - hp * (unsigned)p.wb;
          valid = (int)hp < p.out_h && (int)wp < p.out_w;
          img = ni;
          my_pix = e.halo ? ((size_t)ni * (p.out_h + 2) + hp + 1) * (p.out_w + 2) + wp + 1 : ((size_t)ni * p.out_h + hp) * p.out_w + wp;
        }
        const float* drow = (e.drop != nullptr && valid) ? e.drop + (size_t)img * e.cout : nullptr;
        const float vmask = valid ? 1.f : 0.f;
        bf16* dst = e.out + my_pix * e.out_cpitch + co0;
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (buf * (uint32_t)T + (uint32_t)t) * (uint32_t)p.block_n;
#pragma unroll 1
        for (int j = 0; j < nchunks; ++j) {
          float v[16];
          if (p.ko & 32) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = (float)(i + j);
          } else {
            tmem_ld16(t_addr + j * 16, v);
          }
          if (e.scale != nullptr) {
            // vec4 is set by the launcher when scale / shift hold round_up(cout, 16) floats per n-tile column range (the
            // executor's BatchNorm vectors are padded with zeros up to the channel pitch) and are 16-byte aligned
            const int c0 = co0 + j * 16;
            if (e.relu & 2) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const float4 sc = __ldg(reinterpret_cast<const float4*>(e.scale + c0 + i));
                const float4 sh = __ldg(reinterpret_cast<const float4*>(e.bias + c0 + i));
                v[i] = fmaf(v[i], sc.x, sh.x); v[i + 1] = fmaf(v[i + 1], sc.y, sh.y);
                v[i + 2] = fmaf(v[i + 2], sc.z, sh.z); v[i + 3] = fmaf(v[i + 3], sc.w, sh.w);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = (c0 + i < e.cout) ? fmaf(v[i], __ldg(e.scale + c0 + i), __ldg(e.bias + c0 + i)) : 0.f;
            }
          } else if (e.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += (co0 + j * 16 + i < e.cout) ? __ldg(e.bias + co0 + j * 16 + i) : 0.f;
          }
          if (e.relu & 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (drow != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] *= (co0 + j * 16 + i < e.cout) ? __ldg(drow + co0 + j * 16 + i) : 0.f;
          }
          const uint4 lo = pack8(v), hi8 = pack8(v + 8);
          if (valid && !(p.ko & 16)) {
            if (co0 + j * 16 < e.out_cmax) *reinterpret_cast<uint4*>(dst + j * 16) = lo;
            if (co0 + j * 16 + 8 < e.out_cmax) *reinterpret_cast<uint4*>(dst + j * 16 + 8) = hi8;
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
            // lane l holds column (l >> 1): even lanes keep the sums, odd lanes the squares (private per-warp accumulators)
            const int col = co0 + j * 16 + (lane >> 1);
            if ((lane & 1) == 0) wacc_s[col] += cs;
            else wacc_q[col] += cq;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty0 + buf * 8u);
    }
    if (stats) {
      // combine the four warps (fixed order) and write this CTA's partial row
      named_bar_sync(1, 128);
      for (int col = et; col < e.out_cpitch; col += 128) {
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

void plan_n(int cout, int* block_n, int* n_tiles) {
  const int c16 = round_up(cout, 16);
  *n_tiles = ceil_div(c16, 256);
  *block_n = round_up(ceil_div(c16, *n_tiles), 16);
}

// shared-memory plan for T tiles per CTA; returns false when it does not fit
bool plan_smem(int T, int seg_kh, int wb, int block_n, int acc_cols, C2Params* p, size_t* smem_bytes) {
  p->T = T;
  p->seg_kh = seg_kh;
  p->seg_rows = T * kBlockM + (seg_kh ? 2 : 2 * wb + 2);
  p->seg_full = p->seg_rows / kBlockM;
  p->seg_rem = p->seg_rows % kBlockM;
  p->a_slot_bytes = round_up(p->seg_rows * 128, 1024);
  p->b_tap_bytes = (block_n / 2) * 128;
  p->b_slot_bytes = round_up(3 * p->b_tap_bytes, 1024);
  const int fixed = 8 * acc_cols * 4 + (2 * kMaxASlots + 2 * kMaxBSlots + 4) * 8 + 16 + 64 + 1024;
  const int budget = 227 * 1024 - fixed;
  // at least two A segments and three weight blocks in flight; then weight blocks up to 6, then a third A segment
  int a_slots = seg_kh ? 3 : 2, b_slots = 3;
  if (a_slots * p->a_slot_bytes + b_slots * p->b_slot_bytes > budget) return false;
  while (b_slots < 6 && a_slots * p->a_slot_bytes + (b_slots + 1) * p->b_slot_bytes <= budget) ++b_slots;
  while (a_slots < (seg_kh ? kMaxASlots : 3) && (a_slots + 1) * p->a_slot_bytes + b_slots * p->b_slot_bytes <= budget) ++a_slots;
  while (b_slots < kMaxBSlots && a_slots * p->a_slot_bytes + (b_slots + 1) * p->b_slot_bytes <= budget) ++b_slots;
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
  plan_n(cout, &block_n, &n_tiles);
  C2Params p{};
  size_t smem_bytes;
  return plan_smem(1, 0, in.wb(), block_n, block_n * n_tiles, &p, &smem_bytes) || plan_smem(1, 1, in.wb(), block_n, block_n * n_tiles, &p, &smem_bytes);
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
  plan_n(cout, &p.block_n, &p.n_tiles_n);
  p.acc_cols = p.block_n * p.n_tiles_n;
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
  MIMO_CHECK((p.block_n / 2) % 8 == 0, MIMO_ERR_ARG, "conv3x3_c2: block_n/2 must be a multiple of 8");
  // tiles per CTA: more tiles share one weight block and one segment halo (fewer bytes through TMA per FLOP) but make the
  // items coarser (wave quantisation over the 74 CTA pairs). Pick the T with the lowest modelled cost.
  const int max_pairs = num_sms() / 2;
  size_t smem_bytes = 0;
  {
    static const int forced_t = env_int("MIMO_C2_T", 0);
    static const int forced_kh = env_int("MIMO_C2_SEGKH", -1);
    int best_t = 0, best_kh = 0;
    double best = 1e300;
    for (int T = 1; T <= 4; T *= 2) {
      if (forced_t && T != forced_t) continue;
      if (2 * T * p.block_n > 512) continue;
      for (int kh = 0; kh < 2; ++kh) {
        if (forced_kh >= 0 && kh != forced_kh) continue;
        C2Params q = p;
        size_t sb;
        if (!plan_smem(T, kh, p.wb, p.block_n, p.acc_cols, &q, &sb)) continue;
        const long long items = ceil_div_ll(p.total_pos, 2ll * T * kBlockM) * p.n_tiles_n;
        const long long waves = ceil_div_ll(items, max_pairs);
        const double a_rows = kh ? 3.0 * q.seg_rows : (double)q.seg_rows;
        const double rows_per_item = p.cin_chunks * (a_rows + 9.0 * (p.block_n / 2));   // TMA rows per CTA per item
        const double cost = (double)waves * rows_per_item;
        if (cost < best) { best = cost; best_t = T; best_kh = kh; }
      }
    }
    MIMO_CHECK(best_t > 0, MIMO_ERR_ARG, "conv3x3_c2: not enough shared memory for block_n=%d, row pitch %d", p.block_n, p.wb);
    plan_smem(best_t, best_kh, p.wb, p.block_n, p.acc_cols, &p, &smem_bytes);
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
  MIMO_CUDA(cudaFuncSetAttribute(conv3x3_c2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const int items = p.n_mitems * p.n_tiles_n;
  const int grid = 2 * (items < max_pairs ? items : max_pairs);
  conv3x3_c2_kernel<<<grid, kThreads, smem_bytes, stream>>>(tm_a, tm_a2, tm_b, tm_b2, tm_w, p);
  MIMO_LAUNCH_CHECK();
  return MIMO_OK;
}

}  // namespace mimo
