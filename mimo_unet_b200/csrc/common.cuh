// Shared device/host helpers for the mimo_unet_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mimo_b200.h"

#if defined(__CUDA_ARCH__) && !defined(__CUDA_ARCH_FEAT_SM100_ALL)
#error "mimo_unet_b200 kernels must be compiled with -gencode arch=compute_100a,code=sm_100a"
#endif

typedef __nv_bfloat16 bf16;

namespace mimo {

// ---------------------------------------------------------------------------------------------
// error plumbing (C-ABI: return codes + thread-local message, no exceptions across the boundary)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* last_error();

#define MIMO_CHECK(cond, code, ...)                 \
  do {                                              \
    if (!(cond)) {                                  \
      ::mimo::set_error(__VA_ARGS__);               \
      return (code);                                \
    }                                               \
  } while (0)

#define MIMO_CUDA(expr)                                                                     \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::mimo::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MIMO_ERR_CUDA;                                                                 \
    }                                                                                       \
  } while (0)

#define MIMO_LAUNCH_CHECK()                                                                 \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) {                                                                \
      ::mimo::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MIMO_ERR_CUDA;                                                                 \
    }                                                                                       \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

int num_sms();

// conv launchers record which kernel they dispatched to: 1 flat, 2 flatk, 3 igemm (4-D), 4 wgrad_flat, 5 wgrad_flatk, 6 wgrad (4-D),
// 7 c2 (CTA-pair fprop/dgrad), 8 wgrad_c2
void note_kernel(int id);
int last_kernel();
const char* conv_kernel_name(int id);

// ---------------------------------------------------------------------------------------------
// Activation view: NHWC bf16 with an optional halo and an optional channel slice.
//   pad == 0: dense [N][H][W]
//   pad == 1: one-pixel halo all around, buffer [N][H+2][W+2], interior at (1,1)   (reflect halo of conv inputs)
//   pad == 2: two-pixel ZERO TAIL, buffer [N][H+2][W+2], interior at (0,0); the two extra columns / rows after every
//             image row / image stay zero, so a flattened pixel index can run across row ends (conv_flat.cu)
// element (n,h,w,c) lives at base[((n*hb() + h+org())*wb() + w+org())*cpitch + c_off + c]
// ---------------------------------------------------------------------------------------------
struct ActView {
  bf16* base;
  int N, H, W;
  int pad;     // 0, 1 or 2 (see above)
  int cpitch;  // channels per pixel in memory (multiple of 8)
  int c_off;   // first channel of this view
  int C;       // channels in this view
  __host__ __device__ inline int hb() const { return H + (pad ? 2 : 0); }
  __host__ __device__ inline int wb() const { return W + (pad ? 2 : 0); }
  __host__ __device__ inline int org() const { return pad == 1 ? 1 : 0; }
  __host__ __device__ inline long long pix(int n, int h, int w) const {
    return ((long long)(n * hb() + h + org()) * wb() + (w + org())) * cpitch + c_off;
  }
};

static inline ActView make_view(const mimo_act_t& a) {
  ActView v;
  v.base = (bf16*)a.ptr; v.N = a.n; v.H = a.h; v.W = a.w; v.pad = a.pad; v.cpitch = a.cpitch; v.c_off = a.c_off; v.C = a.c;
  return v;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// reflect index for a 1-pixel halo: -1 -> 1, n -> n-2
__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

struct alignas(16) bf16x8 {
  bf16 v[8];
};

// Loads up to 8 channels [c, c+8) of a pixel (ptr points at channel c). Vector path when aligned.
__device__ __forceinline__ void load8(const bf16* p, int nvalid, float out[8]) {
  if (nvalid >= 8 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
    bf16x8 t = *reinterpret_cast<const bf16x8*>(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = __bfloat162float(t.v[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = (i < nvalid) ? __bfloat162float(p[i]) : 0.f;
  }
}

// ---- register-level bf16 <-> fp32 helpers for the memory-bound kernels (they are instruction-issue bound, not HBM
// bound, unless the per-element instruction count is kept to a handful) ----
// 8 bf16 in a uint4 -> 8 floats: low half = x << 16, high half = x & 0xffff0000 (one ALU op per element)
__device__ __forceinline__ void unpack8(const uint4& r, float f[8]) {
  f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
  f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
  f[4] = __uint_as_float(r.z << 16); f[5] = __uint_as_float(r.z & 0xffff0000u);
  f[6] = __uint_as_float(r.w << 16); f[7] = __uint_as_float(r.w & 0xffff0000u);
}
// two floats -> packed bf16x2 (round to nearest even), `lo` in the low half: one F2FP instruction
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint4 pack8(const float f[8]) {
  return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
}
// word masks that zero the channels >= nvalid of an 8-channel group (loop-invariant per thread)
__device__ __forceinline__ uint4 group_mask(int nvalid) {
  uint32_t m[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) m[i] = (2 * i + 1 < nvalid) ? 0xffffffffu : ((2 * i < nvalid) ? 0x0000ffffu : 0u);
  return make_uint4(m[0], m[1], m[2], m[3]);
}
__device__ __forceinline__ uint4 and4(const uint4& a, const uint4& m) { return make_uint4(a.x & m.x, a.y & m.y, a.z & m.z, a.w & m.w); }
// 16-byte load of an 8-channel group when the view is 16-byte aligned (VEC), generic path otherwise; channels >= nvalid
// come back as zero either way (the vector path may over-read pad channels inside the pixel's pitch and masks them)
// own = view channels [c, c+8), next = [c+8, c+16) (packed bf16): returns the 8 channels starting at c + e, 0 < e < 8
__device__ __forceinline__ uint4 funnel8(const uint4& own, const uint4& next, int e) {
  const uint32_t w[8] = {own.x, own.y, own.z, own.w, next.x, next.y, next.z, next.w};
  uint4 r = own;
  switch (e) {
#define MIMO_F8(E)                                                                                                       \
  case E:                                                                                                                \
    r.x = __funnelshift_r(w[(E >> 1) + 0], w[(E >> 1) + 1], (E & 1) * 16);                                               \
    r.y = __funnelshift_r(w[(E >> 1) + 1], w[(E >> 1) + 2], (E & 1) * 16);                                               \
    r.z = __funnelshift_r(w[(E >> 1) + 2], w[(E >> 1) + 3], (E & 1) * 16);                                               \
    r.w = __funnelshift_r(w[(E >> 1) + 3], w[(E >> 1) + 4 > 7 ? 7 : (E >> 1) + 4], (E & 1) * 16);                        \
    break;
    MIMO_F8(1) MIMO_F8(2) MIMO_F8(3) MIMO_F8(4) MIMO_F8(5) MIMO_F8(6) MIMO_F8(7)
#undef MIMO_F8
    default: break;
  }
  return r;
}

template <bool VEC>
__device__ __forceinline__ uint4 load_group(const bf16* p, int nvalid, const uint4& mask) {
  if (VEC) return and4(*reinterpret_cast<const uint4*>(p), mask);
  // Unaligned view (channel offset not a multiple of 8, e.g. the [21, 63) slice of the decoder concat): the one or two
  // 16-byte words that hold the group are loaded whole and funnel-shifted. Both words contain at least one requested
  // channel, so they lie inside the pixel (pixel pitches are multiples of 16 bytes) -- no scalar 2-byte loads.
  if (nvalid <= 0) return make_uint4(0, 0, 0, 0);
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const int r = (int)((a & 15) >> 1);
  const uint4* q = reinterpret_cast<const uint4*>(a & ~uintptr_t(15));
  const uint4 lo = q[0];
  if (r == 0) return and4(lo, mask);
  const uint4 hi = (r + nvalid > 8) ? q[1] : make_uint4(0, 0, 0, 0);
  return and4(funnel8(lo, hi, r), mask);
}

__device__ __forceinline__ void store8(bf16* p, int nvalid, const float in[8]) {
  if (nvalid >= 8 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
    bf16x8 t;
#pragma unroll
    for (int i = 0; i < 8; ++i) t.v[i] = __float2bfloat16_rn(in[i]);
    *reinterpret_cast<bf16x8*>(p) = t;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nvalid) p[i] = __float2bfloat16_rn(in[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA, tcgen05 (sm_100a)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)   // suspend-time hint (as CUTLASS's ClusterBarrier::wait): without it a
      : "memory");                                         // waiting warp re-issues try_wait every few cycles and starves the
  return ok != 0;                                          // working warps of its SM sub-partition of issue slots (measured)
}
// Bounded spin: a protocol bug turns into a trap (reported as a CUDA error) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("mimo_b200: mbarrier wait timeout (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}

// Polling wait (mbarrier.test_wait never suspends the thread): for barriers whose arrivals come from ANOTHER SM (remote
// arrives of the peer CTA, multicast tcgen05.commit): a thread suspended in try_wait is not woken promptly by those.
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_poll(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_test_wait(bar, parity)) {
    if (++spins > (1u << 28)) {
      printf("mimo_b200: mbarrier poll timeout (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// L2 prefetch of a 2-D box (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1)
               : "memory");
}

// --- tcgen05 ---
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// The same load without the wait: several loads can be in flight before one tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, version 1 = sm_100).
//   start address >>4 in [0,14), LBO>>4 in [16,30), SBO>>4 in [32,46), version in [46,48), layout in [61,64)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}
constexpr uint32_t kLayoutSW128 = 2;

// The same descriptor as two 32-bit words. The elected MMA thread is a single thread executing a dependent
// instruction stream (~4+ cycles per instruction), so building 64-bit descriptors with shifts/ors per MMA makes the
// ISSUE loop the bottleneck (measured: ~110 cycles per MMA). Instead the high word is a constant and the low word is
// advanced with ONE 32-bit add in 16-byte units (all shared-memory addresses are < 256 KB, so the 14-bit start
// address field never carries into the LBO field).
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | ((layout & 7u) << 29);
}
__device__ __forceinline__ void umma_bf16_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same, with the matrix base offset field (bits [49,52)): (start_address >> 7) & 7 when the start address is not
// aligned to the 1024-byte swizzle-128B repeat (row-shifted windows of a larger tile)
__device__ __forceinline__ uint64_t make_smem_desc_bo(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout,
                                                      uint32_t base_offset) {
  return make_smem_desc(saddr, lbo_bytes, sbo_bytes, layout) | ((uint64_t)(base_offset & 7) << 49);
}

// UMMA instruction descriptor for kind::f16 with bf16 A/B and fp32 D (cute::UMMA::InstrDescriptor).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// One lane of a fully converged warp. The TMA / MMA warps run their loops warp-uniformly and only the issue
// instructions sit under this predicate: with an `if (lane == 0)` around the whole loop ptxas treats every operand as
// divergent and wraps each UTCHMMA / UTMALDG in a vote + R2UR uniformisation loop (~15 instructions per MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): two CTAs of a 2-CTA cluster (one TPC) execute ONE tcgen05.mma with M = 256 (128 accumulator
// rows in each CTA's tensor memory, each CTA supplies its own A rows and HALF of the B rows from its shared memory).
// Measured on B200 (tools/gpu/umma_probe.cu, profiles/r02_umma_probe.txt): a cta_group::1 M128 x N x K16 MMA costs
// 43 + N/2 cycles, the cta_group::2 M256 one max(N/2, ~40) cycles: the pair runs the tensor pipe at its full rate from
// N = 96 on, the single CTA never does (75 % at N = 256, 27 % at N = 32).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) in the CTA with rank `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// Default semantics (release at CTA scope), as CUTLASS's ClusterBarrier::arrive(cta_id) issues it. An explicit
// .release.cluster compiles to MEMBAR.ALL.GPU in front of the arrive: measured ~900 cycles per arrive, and in the epilogue
// warps it also waits for every global store of the previous tile to be acknowledged (profiles/r02_findings.md).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads whose completion is signalled on an mbarrier of EITHER CTA of the pair (shared::cluster address)
__device__ __forceinline__ void tma_load_2d_cg2(const CUtensorMap* m, uint32_t bar_cluster, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_cg2(const CUtensorMap* m, uint32_t bar_cluster, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// issued by ONE thread of the leader CTA (rank 0); descriptors are shared::cta offsets valid in both CTAs
__device__ __forceinline__ void umma2_bf16_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once all previously issued MMAs of this thread have completed) on the mbarrier at the same shared-memory offset in
// every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit2_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------
// host: TMA tensor-map encoding through the driver entry point (no link-time libcuda dependency)
// ---------------------------------------------------------------------------------------------
// dims/strides innermost first; strides[i] is the byte stride of dim i+1 (rank-1 entries).
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle128);

}  // namespace mimo
