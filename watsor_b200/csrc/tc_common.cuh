// tc_common.cuh -- inline PTX wrappers shared by the tcgen05 kernels (kernels_tc.cu, kernels_fused.cu):
// mbarrier, TMA (cp.async.bulk.tensor load / store), TMEM allocation, tcgen05.mma / commit / ld, and the
// shared-memory / instruction descriptors (bit layouts as in cute/arch/mma_sm100_desc.hpp).  sm_100a only.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// One elected lane of a converged warp.  Unlike `lane == 0`, ptxas knows the guarded region runs in a single
// thread, so the tcgen05.mma operands move to uniform registers without the per-instruction
// ELECT / R2UR.BROADCAST / BRA.U.ANY uniformisation loop (measured: ~100 -> ~30 cycles per issued MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}

template <bool TF32>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (TF32)
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand from tensor memory (lane = row of the 128-row tile, one TF32 per 32-bit column), B from shared memory.
__device__ __forceinline__ void umma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 consecutive columns of this thread's TMEM lane (the mirror image of tmem_ld16's shape)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 16 accumulator columns of this thread's row, summed over the TF32X3 partial accumulators in the fixed
// order ((main0 + main1) + main2) + corr.  All tcgen05.ld are issued before the single wait::ld, so the
// TMEM round trips overlap instead of serialising.
template <bool X3>
__device__ __forceinline__ void load_acc16(uint32_t taddr, int block_n, int n_main, int used, uint32_t* v) {
  if (!X3) {
    tmem_ld16(taddr, v);
    tmem_ld_wait();
    return;
  }
  uint32_t u1[16], u2[16], uc[16];
  tmem_ld16(taddr, v);
  if (used > 1) tmem_ld16(taddr + (uint32_t)block_n, u1);
  if (used > 2) tmem_ld16(taddr + (uint32_t)(2 * block_n), u2);
  tmem_ld16(taddr + (uint32_t)(n_main * block_n), uc);
  tmem_ld_wait();
  if (used > 1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__fadd_rn(__uint_as_float(v[i]), __uint_as_float(u1[i])));
  }
  if (used > 2) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__fadd_rn(__uint_as_float(v[i]), __uint_as_float(u2[i])));
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__fadd_rn(__uint_as_float(v[i]), __uint_as_float(uc[i])));
}

// 32 accumulator columns at once.  With a single main accumulator (or none to add) all four / two tcgen05.ld
// go out before one wait::ld - the epilogue warps are latency bound (profiles/r01_pipeline_trace.md), and two
// dependent ld -> wait round trips per 32 columns were a third of their time.  Same RN additions as load_acc16.
template <bool X3>
__device__ __forceinline__ void load_acc32(uint32_t taddr, int block_n, int n_main, int used, uint32_t* v) {
  if (!X3) {
    tmem_ld16(taddr, v);
    tmem_ld16(taddr + 16u, v + 16);
    tmem_ld_wait();
    return;
  }
  if (used > 1) {
    load_acc16<true>(taddr, block_n, n_main, used, v);
    load_acc16<true>(taddr + 16u, block_n, n_main, used, v + 16);
    return;
  }
  uint32_t uc[32];
  tmem_ld16(taddr, v);
  tmem_ld16(taddr + 16u, v + 16);
  tmem_ld16(taddr + (uint32_t)(n_main * block_n), uc);
  tmem_ld16(taddr + (uint32_t)(n_main * block_n) + 16u, uc + 16);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__fadd_rn(__uint_as_float(v[i]), __uint_as_float(uc[i])));
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout_type=2 (SW128) [61,64).
// Rows are 128 B apart, 8-row swizzle atoms 1024 B apart (SBO).
// Explicit shared-window accesses: the carve-up of the dynamic buffer goes through integer alignment, after
// which nvcc no longer proves the address space and would emit generic LD.E/ST.E in the hottest loops.
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// Read-only tables (written once before the prologue __syncthreads, never again): NOT volatile, so the compiler may
// hoist / pipeline these loads across the FFMA chains that consume them.  Never use for data guarded by an mbarrier.
__device__ __forceinline__ float4 lds128_ro(uint32_t addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// Pipeline tracing (diagnostic builds only: make EXTRA=-DWB_TRACE; see tools/trace_pipeline.py).  CTA 0 stamps
// clock64() per role and iteration into a per-translation-unit buffer that wb_trace_read_*() copies out.
#ifdef WB_TRACE
#define WB_TRACE_SLOTS 12
#define WB_TRACE_ITERS 64
static __device__ long long wb_trace_buf[WB_TRACE_SLOTS * WB_TRACE_ITERS];
#define WB_STAMP(kind, it)                                                                  \
  do {                                                                                      \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (it) < WB_TRACE_ITERS)     \
      wb_trace_buf[(kind) * WB_TRACE_ITERS + (it)] = clock64();                             \
  } while (0)
#else
#define WB_STAMP(kind, it) \
  do {                     \
  } while (0)
#endif

__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A/B format, K-major both, N>>3, M>>4
__host__ __device__ inline uint32_t make_idesc(bool tf32, int m, int n) {
  uint32_t d = 0;
  d |= 1u << 4;                        // c_format = F32
  d |= (tf32 ? 2u : 1u) << 7;          // a_format  (BF16 = 1, TF32 = 2)
  d |= (tf32 ? 2u : 1u) << 10;         // b_format
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}

constexpr int BLOCK_M = 128;
constexpr int ROW_BYTES = 128;                     // one swizzle row = 64 bf16 or 32 fp32 along K
constexpr int A_TILE_BYTES = BLOCK_M * ROW_BYTES;  // 16 KB
constexpr int UMMA_K_BYTES = 32;

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}


}  // namespace
