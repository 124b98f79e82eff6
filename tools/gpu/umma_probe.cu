// Micro-benchmark of tcgen05.mma issue/execute cost on B200 (sm_100a): cycles per MMA for the instruction shapes the conv
// kernels could use. One CTA (or CTA pair) per SM, operands resident in shared memory (contents irrelevant), one elected
// thread issues `reps` back-to-back MMAs into the same accumulator, commits, waits; clock64() around the whole sequence.
// Also: tcgen05.ld throughput (epilogue cost) for 4 and 8 warps.
//   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/gpu/umma_probe.bin tools/gpu/umma_probe.cu
#include <algorithm>
#include <cstdio>
#include <vector>

#include "../../mimo_unet_b200/csrc/common.cuh"

using namespace mimo;

__device__ __forceinline__ void umma_ts_bf16_w(uint32_t tmem_d, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit2_local(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct ProbeArgs {
  int M, N;        // instruction shape (M = 256 needs group 2)
  int group;       // cta_group 1 or 2
  int a_tmem;      // A operand from tensor memory (group 1 only)
  int reps;
  int b_mn_major;  // B operand MN-major (wgrad-style) instead of K-major
  int a_mn_major;
  int a_row_off;   // A window starts this many 128-byte rows past the 1024-byte swizzle atom (row-shifted conv tap windows)
};

template <int GROUP>
__global__ void __launch_bounds__(128, 1) umma_probe_kernel(ProbeArgs p, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;               // up to 128 rows x 128 B = 16 KB
  uint8_t* smem_b = smem + 32 * 1024;   // up to 256 rows x 128 B = 32 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = GROUP == 2 ? cluster_ctarank() : 0;
  // deterministic finite operand contents
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 0xff);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 0) {
    if constexpr (GROUP == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc(&tmem_ptr, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (GROUP == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  long long t0 = 0, t1 = 0;
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = make_idesc_bf16(p.M, p.N, p.a_mn_major, p.b_mn_major);
    constexpr uint32_t hi = desc_hi(1024, kLayoutSW128);
    const uint32_t a_lo = desc_lo(smem_u32(smem_a) + 128u * (uint32_t)p.a_row_off, p.a_mn_major ? 128 : 16);
    const uint32_t b_lo = desc_lo(smem_u32(smem_b), p.b_mn_major ? 128 : 16);
    const uint32_t d = tmem_base;             // accumulator columns [0, N)
    const uint32_t a_t = tmem_base + 256;     // A operand in TMEM (columns 256..)
    __syncwarp();
    t0 = clock64();
    for (int r = 0; r < p.reps; r += 4) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t ka = p.a_mn_major ? 0u : (uint32_t)(2 * k), kb = p.b_mn_major ? 0u : (uint32_t)(2 * k);
          if (GROUP == 1 && p.a_tmem) umma_ts_bf16_w(d, a_t + 8 * k, b_lo + kb, hi, idesc, 1);
          else if constexpr (GROUP == 2) umma2_bf16_w(d, a_lo + ka, hi, b_lo + kb, hi, idesc, 1);
          else umma_bf16_w(d, a_lo + ka, hi, b_lo + kb, hi, idesc, 1);
        }
      }
      __syncwarp();
    }
    if (elect_one()) {
      if constexpr (GROUP == 2) umma_commit2_local(&bar);
      else umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (GROUP == 2) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if constexpr (GROUP == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else tmem_dealloc(tmem_base, 512);
  }
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
}

// tcgen05.ld throughput: `warps` warps (4 or 8; warp w reads lane quarter w & 3) each load `cols` fp32 columns per
// iteration in x16 / x32 chunks.
template <int X>
__device__ __forceinline__ void tmem_ldx(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void tmem_ldx<16>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
template <>
__device__ __forceinline__ void tmem_ldx<32>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
      "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

template <int X>
__global__ void __launch_bounds__(256, 1) tmem_ld_probe_kernel(int warps, int reps, int wait_each, unsigned long long* out, float* sink) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  long long t0 = 0, t1 = 0;
  float acc = 0.f;
  if (warp < warps) {
    const uint32_t t_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
    __syncwarp();
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      uint32_t v[X];
      tmem_ldx<X>(t_addr + (uint32_t)((r * X) & 127), v);
      if (wait_each) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < X; ++i) acc += __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    t1 = clock64();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (sink && acc == 123.456f) sink[0] = acc;
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
}


// Cost of tcgen05.commit in the issue stream: `reps` iterations of { n_mma MMAs (N = 64); commit -> sink barrier }, then one
// final commit + wait. mode 0: cta_group::1 local, 1: cta_group::2 local barrier, 2: cta_group::2 multicast to both CTAs.
template <int GROUP>
__global__ void __launch_bounds__(128, 1) commit_probe_kernel(int n_mma, int mode, int reps, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, sink[4];
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = GROUP == 2 ? cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 0xff);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&sink[i], 100000);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 0) {
    if constexpr (GROUP == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc(&tmem_ptr, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (GROUP == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  long long t0 = 0, t1 = 0;
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = make_idesc_bf16(GROUP == 2 ? 256 : 128, 64, 0, 0);
    constexpr uint32_t hi = desc_hi(1024, kLayoutSW128);
    const uint32_t a_lo = desc_lo(smem_u32(smem), 16);
    const uint32_t b_lo = desc_lo(smem_u32(smem + 32 * 1024), 16);
    __syncwarp();
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (elect_one()) {
        for (int k = 0; k < n_mma; ++k) {
          if constexpr (GROUP == 2) umma2_bf16_w(tmem_base, a_lo + 2 * (k & 3), hi, b_lo + 2 * (k & 3), hi, idesc, 1);
          else umma_bf16_w(tmem_base, a_lo + 2 * (k & 3), hi, b_lo + 2 * (k & 3), hi, idesc, 1);
        }
        if constexpr (GROUP == 2) {
          if (mode == 2)
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                             smem_u32(&sink[r & 3])), "h"((uint16_t)3) : "memory");
          else umma_commit2_local(&sink[r & 3]);
        } else {
          umma_commit(&sink[r & 3]);
        }
      }
      __syncwarp();
    }
    if (elect_one()) {
      if constexpr (GROUP == 2) umma_commit2_local(&bar);
      else umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (GROUP == 2) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if constexpr (GROUP == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else tmem_dealloc(tmem_base, 512);
  }
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
}

static double median(std::vector<unsigned long long> v) {
  std::sort(v.begin(), v.end());
  return (double)v[v.size() / 2];
}

int main() {
  int dev = 0, sms = 0;
  cudaSetDevice(dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  unsigned long long* out = nullptr;
  cudaMalloc(&out, sizeof(unsigned long long) * sms);
  float* sink = nullptr;
  cudaMalloc(&sink, 4);
  std::vector<unsigned long long> h(sms);
  const size_t smem = 64 * 1024 + 1024;
  cudaFuncSetAttribute(umma_probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(umma_probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int reps = 2048;
  printf("# tcgen05.mma kind::f16 (bf16 in, fp32 acc), K = 16 per instruction, %d back-to-back MMAs, %d SMs busy\n", reps, sms);
  printf("# group M N a_src a_major b_major | cycles/MMA | dense-math cycles (M*N*16*2/8192 per SM) | flops/cycle/SM\n");
  struct Cfg { int M, N, group, a_tmem, amn, bmn, aoff; };
  std::vector<Cfg> cfgs;
  for (int N : {16, 32, 48, 64, 96, 128, 192, 256}) cfgs.push_back({128, N, 1, 0, 0, 0, 0});
  for (int N : {32, 64, 128, 256}) cfgs.push_back({64, N, 1, 0, 0, 0, 0});
  for (int N : {32, 64, 128, 256}) cfgs.push_back({128, N, 1, 1, 0, 0, 0});
  for (int N : {32, 64, 128, 256}) cfgs.push_back({64, N, 1, 1, 0, 0, 0});
  for (int N : {32, 64, 96, 128, 192, 256}) cfgs.push_back({256, N, 2, 0, 0, 0, 0});
  for (int N : {64, 128, 256}) cfgs.push_back({128, N, 2, 0, 0, 0, 0});
  for (int N : {64, 128, 192, 256}) cfgs.push_back({128, N, 1, 0, 1, 1, 0});   // both MN-major (wgrad)
  for (int N : {128, 192, 256}) cfgs.push_back({256, N, 2, 0, 1, 1, 0});
  for (int off : {1, 3, 4}) for (int N : {32, 64, 96, 176}) cfgs.push_back({256, N, 2, 0, 0, 0, off});   // row-shifted A windows
  for (int off : {1, 3}) for (int N : {32, 64}) cfgs.push_back({128, N, 1, 0, 0, 0, off});
  for (const Cfg& c : cfgs) {
    ProbeArgs a{c.M, c.N, c.group, c.a_tmem, reps, c.bmn, c.amn, c.aoff};
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(c.group == 2 ? (sms / 2) * 2 : sms);
    lc.blockDim = dim3(128);
    lc.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = c.group;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    lc.attrs = at;
    lc.numAttrs = c.group == 2 ? 1 : 0;
    cudaMemset(out, 0, sizeof(unsigned long long) * sms);
    cudaError_t e = c.group == 2 ? cudaLaunchKernelEx(&lc, umma_probe_kernel<2>, a, out) : cudaLaunchKernelEx(&lc, umma_probe_kernel<1>, a, out);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%d %3d %3d %s %d %d | FAILED: %s\n", c.group, c.M, c.N, c.a_tmem ? "tmem" : "smem", c.amn, c.bmn, cudaGetErrorString(e));
      cudaGetLastError();
      return 1;
    }
    cudaMemcpy(h.data(), out, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
    std::vector<unsigned long long> v;
    for (int i = 0; i < (int)lc.gridDim.x; ++i)
      if (h[i]) v.push_back(h[i]);
    const double cyc = median(v) / reps;
    const double per_sm_m = c.group == 2 ? c.M / 2.0 : c.M;
    const double math = per_sm_m * c.N * 16 * 2 / 8192.0;
    printf("%d %3d %3d %s %d %d off%d | %7.1f | %6.1f | %7.0f\n", c.group, c.M, c.N, c.a_tmem ? "tmem" : "smem", c.amn, c.bmn, c.aoff, cyc, math,
           per_sm_m * c.N * 32.0 / cyc);
  }
  printf("# commit cost: cycles per iteration of { n MMAs (N=64); tcgen05.commit } ; group/mode: 1/0 local, 2/1 local, 2/2 multicast\n");
  cudaFuncSetAttribute(commit_probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(commit_probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int gm = 0; gm < 3; ++gm) {
    const int group = gm == 0 ? 1 : 2, mode = gm;
    for (int n_mma : {0, 1, 4, 12}) {
      const int r3 = 1024;
      cudaLaunchConfig_t lc = {};
      lc.gridDim = dim3(group == 2 ? (sms / 2) * 2 : sms);
      lc.blockDim = dim3(128);
      lc.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = group; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      lc.attrs = at;
      lc.numAttrs = group == 2 ? 1 : 0;
      cudaMemset(out, 0, sizeof(unsigned long long) * sms);
      cudaError_t e = group == 2 ? cudaLaunchKernelEx(&lc, commit_probe_kernel<2>, n_mma, mode, r3, out)
                                 : cudaLaunchKernelEx(&lc, commit_probe_kernel<1>, n_mma, mode, r3, out);
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("commit probe failed: %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h.data(), out, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
      std::vector<unsigned long long> v;
      for (int i = 0; i < (int)lc.gridDim.x; ++i) if (h[i]) v.push_back(h[i]);
      printf("group %d mode %d n_mma %2d | %7.1f cycles/iteration\n", group, mode, n_mma, median(v) / r3);
    }
  }
  printf("# tcgen05.ld 32x32b: cycles per instruction per warp (all warps concurrently), bytes/cycle/SM\n");
  for (int warps : {4, 8}) {
    for (int wait_each : {1, 0}) {
      for (int X : {16, 32}) {
        const int r2 = 4096;
        if (X == 16) tmem_ld_probe_kernel<16><<<sms, 256>>>(warps, r2, wait_each, out, sink);
        else tmem_ld_probe_kernel<32><<<sms, 256>>>(warps, r2, wait_each, out, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("tmem_ld probe failed: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), out, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
        const double cyc = median(h) / r2;
        printf("warps %d x%d wait_each %d | %6.1f cycles/ld | %7.1f B/cycle/SM\n", warps, X, wait_each, cyc, warps * 32.0 * X * 4 / cyc);
      }
    }
  }
  return 0;
}
