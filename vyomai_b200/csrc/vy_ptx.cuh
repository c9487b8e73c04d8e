// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM).
// Everything here is device-side and header-only; no CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define VY_DEVINL __device__ __forceinline__

namespace vy {

VY_DEVINL uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

VY_DEVINL uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
VY_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
VY_DEVINL void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
VY_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
VY_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
VY_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      " selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
VY_DEVINL bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {  // local barrier, remote arrivals
  uint32_t ok;
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
      " selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug (wrong descriptor, missing arrive) traps instead of hanging the GPU; the trap surfaces
// as a launch failure on the host (vy_last_error). try_wait suspends the thread in hardware for a bounded time per
// attempt, so the loop is a handful of instructions per microsecond-scale attempt (the single-thread producer / MMA
// warps share their SM sub-partitions with epilogue warps and must not eat their issue slots) and 2^22 failed attempts
// are seconds.
VY_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) __trap();
  }
}

// GEMM flavour of the bounded wait: instead of trapping, a wait that times out raises a flag and returns; every other
// wait of the SAME launch notices the flag within 64 attempts and returns too, so a protocol bug ends that kernel with
// wrong results and a raised flag rather than with a fault or a hang. The flag holds the id of the launch that timed
// out: waits of later launches compare it with their own id, so a stale flag never makes a healthy GEMM bail out.
// `host` is the device address of a pinned host mirror the library reads without synchronising (every vy_gemm call
// and vy_gemm_poison_peek() fail loudly once it is non-zero; vy_gemm_poisoned() lowers both).
struct PoisonRef {
  int* flag;  // device memory, one per device
  int* host;  // mapped pinned host memory, one per device
  int id;     // id of this launch (never 0)
};
template <bool CLUSTER_SCOPE = false>
VY_DEVINL void mbar_wait_soft(uint64_t* bar, uint32_t parity, const PoisonRef& poison) {
  uint32_t spins = 0;
  while (!(CLUSTER_SCOPE ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait(bar, parity))) {
    if ((++spins & 63u) == 0) {
      if (*reinterpret_cast<volatile int*>(poison.flag) == poison.id) return;
      if (spins > (1u << 21)) {
        atomicExch(poison.flag, poison.id);
        *reinterpret_cast<volatile int*>(poison.host) = poison.id;
        __threadfence_system();
        return;
      }
    }
  }
}

// generic-proxy writes to smem -> visible to async proxy (TMA store / tcgen05.mma operand reads)
VY_DEVINL void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------------
VY_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
VY_DEVINL void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
VY_DEVINL void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
VY_DEVINL void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                           int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
VY_DEVINL void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(m)),
      "r"(smem_u32(src)), "r"(c0), "r"(c1)
      : "memory");
}
VY_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
VY_DEVINL void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
VY_DEVINL void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation
// ----------------------------------------------------------------------------------------------
// Called by one full warp. ncols: power of two in [32, 512].
VY_DEVINL void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
VY_DEVINL void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
VY_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
VY_DEVINL void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
VY_DEVINL void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05: descriptors
// ----------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4   bits [46,48) version = 1
//   bits [61,64) layout type (2 = SWIZZLE_128B)
// K-major tile (rows of 128 B, 8-row groups 1024 B apart): LBO unused (1), SBO = 1024.
// MN-major tile (k-rows of 128 B holding 64 contiguous MN elements): LBO = bytes between
// consecutive 128-B MN blocks, SBO = 1024 (8 k-rows).
// layout_type 1 = SWIZZLE_128B_BASE32B: the only layout MN-major tf32 operands may use (32-B
// swizzle atoms, 4 k-rows per 512-B group; pairs with TMA SWIZZLE_128B_ATOM_32B).
VY_DEVINL uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                        uint32_t layout_type = 2) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3fff);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout_type) << 61;
  return d;
}

// Instruction descriptor for kind::f16 / kind::tf32 with fp32 accumulation.
//   fmt: 0 = f16, 1 = bf16, 2 = tf32. a_mn / b_mn: 1 = MN-major operand.
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int M, int N, int a_mn, int b_mn) {
  return (1u << 4)                                  // c_format = F32
         | (static_cast<uint32_t>(fmt) << 7)        // a_format
         | (static_cast<uint32_t>(fmt) << 10)       // b_format
         | (static_cast<uint32_t>(a_mn) << 15)      // a_major
         | (static_cast<uint32_t>(b_mn) << 16)      // b_major
         | (static_cast<uint32_t>(N >> 3) << 17)    // n_dim
         | (static_cast<uint32_t>(M >> 4) << 24);   // m_dim
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
VY_DEVINL void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                        uint32_t accumulate) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
VY_DEVINL void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05 ops of this thread complete.
// (implies tcgen05.fence::before_thread_sync)
VY_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// ----------------------------------------------------------------------------------------------
// thread-block clusters
// ----------------------------------------------------------------------------------------------
VY_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
VY_DEVINL void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// CTA pairs (tcgen05 cta_group::2): two CTAs of a cluster, on the two SMs of one TPC, run ONE M = 256 MMA. Each CTA
// holds its own 128 rows of A and HALF of the B tile in its shared memory and receives its 128 accumulator rows in its
// own TMEM; the even-ranked CTA (the leader) issues the MMAs for both. Every tcgen05 instruction of such a kernel carries
// .cta_group::2. Barriers that gate the leader's MMA thread live in the leader's shared memory, so the peer signals
// them through shared::cluster addresses (mapa).
// ----------------------------------------------------------------------------------------------
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
VY_DEVINL uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// arrive on an mbarrier given by a shared::cluster address (this CTA's own shared::cta addresses are valid ones)
// (default semantics — release at CTA scope — like every barrier arrive of a tcgen05 pipeline: what the waiter consumes
//  are TMEM reads ordered by tcgen05.fence, not global memory. Asking for .release.cluster makes the compiler emit
//  MEMBAR.ALL + ERRBAR in front of the arrive, which waits for the warp's outstanding TMA / global stores: ~10 % of
//  the epilogue warps' time in the ncu stall samples.)
VY_DEVINL void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
VY_DEVINL void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// both CTAs of the pair execute these (one full warp each); the allocation is made in both TMEMs at the same columns
VY_DEVINL void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
VY_DEVINL void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
VY_DEVINL void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes complete on an mbarrier of either CTA of the pair
VY_DEVINL void tma_load_2d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows: 128 from each CTA] * B[N rows: N/2 from each CTA]; issued by ONE thread of the leader
VY_DEVINL void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
VY_DEVINL void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in every CTA of `cta_mask` once the pair's MMAs issued so far are complete
VY_DEVINL void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM -> registers. 32x32b shape: lane i of the warp reads TMEM lane (base_lane + i),
// N consecutive 32-bit columns. A warp may only touch lanes [32*(warp_id%4), +32).
// ----------------------------------------------------------------------------------------------
VY_DEVINL void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
VY_DEVINL void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
VY_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------------------------
// misc
// ----------------------------------------------------------------------------------------------
// One lane of a converged warp. The single-thread roles (TMA producer, tcgen05.mma issuer) run their loops with the
// WHOLE warp and elect a lane only around the issuing instructions: control flow and descriptor arithmetic then stay
// warp-uniform, so the compiler keeps them in uniform registers. Inside an `if (lane == 0)` region it cannot, and wraps
// every UTCHMMA / UTMALDG in an ELECT + R2UR.BROADCAST + BRA.U.ANY waterfall (~25 instructions and ~200 cycles per MMA,
// measured — enough to make the issue loop, not the tensor pipe, the bottleneck of long-K GEMMs).
VY_DEVINL bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " elect.sync _|p, 0xffffffff;\n"
      " selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

VY_DEVINL void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

VY_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
VY_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 128-byte swizzle used by TMA SWIZZLE_128B and UMMA SWIZZLE_128B: XOR byte-address bits [4,7)
// with bits [7,10). `off` is the byte offset inside a 1024-B aligned tile.
VY_DEVINL uint32_t swz128(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }

}  // namespace vy
