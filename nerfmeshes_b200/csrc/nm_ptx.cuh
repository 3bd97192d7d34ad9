// Thin inline-PTX wrappers for sm_100a: mbarrier, bulk async copy (TMA engine, 1-D), tcgen05 (alloc / mma / ld / st /
// commit / fences).  Encodings follow the PTX ISA; descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nm {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes or `ns` elapse — a long hint
// means a waiting warp issues (almost) nothing instead of polling, which matters under the board's power cap
__device__ __forceinline__ uint32_t mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  return ok;
}
#ifndef NM_WAIT_HINT_NS
#define NM_WAIT_HINT_NS 200000u
#endif
// Bounded wait: a protocol bug must become a trap with a code in *err, never a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
#if NM_WAIT_HINT_NS == 0
  while (!mbar_try_wait(bar, parity)) {              // plain polling (the default suspend window): A/B builds only
#else
  while (!mbar_try_wait_hint(bar, parity, NM_WAIT_HINT_NS)) {
#endif
    if (clock64() - t0 > 4000000000LL) {  // ~2 s
      if (err) atomicExch(err, code);
      __threadfence_system();
      __trap();
    }
  }
}

// ---- async-proxy fences / bulk copy ---------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// ---- cluster of two CTAs sharing one weight stream --------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one L2 read, the same shared-memory offset written (and the same mbarrier offset credited) in every CTA of `mask`
__device__ __forceinline__ void bulk_g2s_multicast(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar), "h"(mask)
               : "memory");
}

// shared -> global bulk store (bulk async-group completion): the issuing thread commits a group and, before the shared-memory
// source is rewritten (or the kernel exits), waits for the group's reads (or writes) to finish
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- tcgen05 ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// Warp-converged variants: every lane of the issuing warp executes the statement, one elected lane issues.  Keeping
// the issuing warp converged matters: a lone diverged lane pays ~200 cycles per tcgen05.mma, a converged warp ~100
// (measured, tools/umma_bench.cu), and several issuing warps overlap that cost.
__device__ __forceinline__ void mma_ss_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred pe, p;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred pe, p;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// One schedule block (64x64, K=64) as ONE asm statement: a single elect, operands converted to uniform registers once,
// descriptor increments in PTX.  Issuing the 12 (exact) / 4 (fast) MMAs one statement at a time costs ~85 cycles each
// (five R2UR + elect per MMA); fused they approach the 32-cycle execution time of an N=64 MMA.
// Order (exact): a_hi*b_hi, a_lo*b_hi, a_hi*b_lo; the first MMA overwrites when acc_first == 0.
__device__ __forceinline__ void mma_block_ts3(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                              uint32_t idesc, uint32_t acc_first) {
  asm volatile(
      "{\n\t.reg .pred pe, pacc, pt;\n\t.reg .b32 a;\n\t.reg .b64 b;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pacc, %6, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %3, %5, pacc;\n\t"
      "add.u32 a, %1, 8;\n\tadd.u64 b, %3, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [a], b, %5, pt;\n\t"
      "add.u32 a, %1, 16;\n\tadd.u64 b, %3, 4;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [a], b, %5, pt;\n\t"
      "add.u32 a, %1, 24;\n\tadd.u64 b, %3, 6;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [a], b, %5, pt;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], %3, %5, pt;\n\t"
      "add.u32 a, %2, 8;\n\tadd.u64 b, %3, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [a], b, %5, pt;\n\t"
      "add.u32 a, %2, 16;\n\tadd.u64 b, %3, 4;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [a], b, %5, pt;\n\t"
      "add.u32 a, %2, 24;\n\tadd.u64 b, %3, 6;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [a], b, %5, pt;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %4, %5, pt;\n\t"
      "add.u32 a, %1, 8;\n\tadd.u64 b, %4, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [a], b, %5, pt;\n\t"
      "add.u32 a, %1, 16;\n\tadd.u64 b, %4, 4;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [a], b, %5, pt;\n\t"
      "add.u32 a, %1, 24;\n\tadd.u64 b, %4, 6;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [a], b, %5, pt;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_hi), "r"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(acc_first)
      : "memory");
}
__device__ __forceinline__ void mma_block_ts1(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                              uint32_t idesc, uint32_t acc_first) {
  asm volatile(
      "{\n\t.reg .pred pe, pacc, pt;\n\t.reg .b32 a;\n\t.reg .b64 b;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pacc, %6, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %3, %5, pacc;\n\t"
      "add.u32 a, %1, 8;\n\tadd.u64 b, %3, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [a], b, %5, pt;\n\t"
      "add.u32 a, %1, 16;\n\tadd.u64 b, %3, 4;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [a], b, %5, pt;\n\t"
      "add.u32 a, %1, 24;\n\tadd.u64 b, %3, 6;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [a], b, %5, pt;\n\t"
      "}"
      ::"r"(d_tmem), "r"(a_hi), "r"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(acc_first)
      : "memory");
}
// same for an encoding block (A from shared memory), ksteps (1..4) K=16 steps per pass
__device__ __forceinline__ void mma_block_ss3(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                              uint32_t idesc, uint32_t acc_first, uint32_t ksteps) {
  asm volatile(
      "{\n\t.reg .pred pe, pk, pacc, pt;\n\t.reg .b64 a, b;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pacc, %6, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\t"
      "setp.gt.u32 pk, %7, 0;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, %1, 0;\n\tadd.u64 b, %3, 0;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pacc;\n\t"
      "setp.gt.u32 pk, %7, 1;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, %1, 2;\n\tadd.u64 b, %3, 2;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 2;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, %1, 4;\n\tadd.u64 b, %3, 4;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 3;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, %1, 6;\n\tadd.u64 b, %3, 6;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 0;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, %2, 0;\n\tadd.u64 b, %3, 0;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 1;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, %2, 2;\n\tadd.u64 b, %3, 2;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 2;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, %2, 4;\n\tadd.u64 b, %3, 4;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 3;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, %2, 6;\n\tadd.u64 b, %3, 6;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 0;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, %1, 0;\n\tadd.u64 b, %4, 0;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 1;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, %1, 2;\n\tadd.u64 b, %4, 2;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 2;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, %1, 4;\n\tadd.u64 b, %4, 4;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 3;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, %1, 6;\n\tadd.u64 b, %4, 6;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "}"
      ::"r"(d_tmem), "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(acc_first), "r"(ksteps)
      : "memory");
}
__device__ __forceinline__ void mma_block_ss1(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                              uint32_t idesc, uint32_t acc_first, uint32_t ksteps) {
  asm volatile(
      "{\n\t.reg .pred pe, pk, pacc, pt;\n\t.reg .b64 a, b;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pacc, %6, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\t"
      "setp.gt.u32 pk, %7, 0;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, %1, 0;\n\tadd.u64 b, %3, 0;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pacc;\n\t"
      "setp.gt.u32 pk, %7, 1;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, %1, 2;\n\tadd.u64 b, %3, 2;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 2;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, %1, 4;\n\tadd.u64 b, %3, 4;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 3;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, %1, 6;\n\tadd.u64 b, %3, 6;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "}"
      ::"r"(d_tmem), "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(acc_first), "r"(ksteps)
      : "memory");
}
// The same blocks with the per-k16 descriptor advance given by the caller (>>4 units): K-major SWIZZLE_128B operands step
// by 32 B (2), MN-major ones by two 8-row groups = 2048 B (128).
__device__ __forceinline__ void mma_block_ss3g(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                               uint32_t idesc, uint32_t acc_first, uint32_t ksteps, uint64_t a_step, uint64_t b_step) {
  asm volatile(
      "{\n\t.reg .pred pe, pk, pacc, pt;\n\t.reg .b64 a, b;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pacc, %6, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\t"
      "setp.gt.u32 pk, %7, 0;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, %1, 0;\n\tadd.u64 b, %3, 0;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pacc;\n\t"
      "setp.gt.u32 pk, %7, 1;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, a, %8;\n\tadd.u64 b, b, %9;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 2;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, a, %8;\n\tadd.u64 b, b, %9;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 3;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, a, %8;\n\tadd.u64 b, b, %9;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 0;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, %2, 0;\n\tadd.u64 b, %3, 0;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 1;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, a, %8;\n\tadd.u64 b, b, %9;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 2;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, a, %8;\n\tadd.u64 b, b, %9;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 3;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, a, %8;\n\tadd.u64 b, b, %9;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 0;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, %1, 0;\n\tadd.u64 b, %4, 0;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 1;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, a, %8;\n\tadd.u64 b, b, %9;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 2;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, a, %8;\n\tadd.u64 b, b, %9;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "setp.gt.u32 pk, %7, 3;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, a, %8;\n\tadd.u64 b, b, %9;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %5, pt;\n\t"
      "}"
      ::"r"(d_tmem), "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(acc_first), "r"(ksteps), "l"(a_step), "l"(b_step)
      : "memory");
}
__device__ __forceinline__ void mma_block_ss1g(uint32_t d_tmem, uint64_t a_hi, uint64_t b_hi, uint32_t idesc, uint32_t acc_first,
                                               uint32_t ksteps, uint64_t a_step, uint64_t b_step) {
  asm volatile(
      "{\n\t.reg .pred pe, pk, pacc, pt;\n\t.reg .b64 a, b;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pacc, %4, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\t"
      "setp.gt.u32 pk, %5, 0;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, %1, 0;\n\tadd.u64 b, %2, 0;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, pacc;\n\t"
      "setp.gt.u32 pk, %5, 1;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, a, %6;\n\tadd.u64 b, b, %7;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, pt;\n\t"
      "setp.gt.u32 pk, %5, 2;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, a, %6;\n\tadd.u64 b, b, %7;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, pt;\n\t"
      "setp.gt.u32 pk, %5, 3;\n\tand.pred pk, pk, pe;\n\tadd.u64 a, a, %6;\n\tadd.u64 b, b, %7;\n\t"
      "@pk tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, pt;\n\t"
      "}"
      ::"r"(d_tmem), "l"(a_hi), "l"(b_hi), "r"(idesc), "r"(acc_first), "r"(ksteps), "l"(a_step), "l"(b_step)
      : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
      : "memory");
}
// the same arrival delivered to the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit_elect_multicast(uint32_t bar, uint16_t mask) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(bar), "h"(mask)
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4
// [16,30) (unused for swizzled K-major, set to 1), SBO>>4 [32,46) = 1024 B between 8-row groups, version=1 [46,48),
// layout_type=2 (SWIZZLE_128B) [61,64).  Tile base must be 1024-byte aligned; advancing K by 16 fp16 = +32 B = +2.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major, SWIZZLE_128B descriptor of a 128 (M or N) x 64 (K) tile of 16-bit elements stored as [MN group of 64][K row][128 B]:
// the 64 MN elements of one K index are one 128-byte line (16-byte chunks XOR-swizzled by the K row % 8, as in the K-major
// case), 8 K rows make a 1024-byte group (SBO), the second MN group of 64 follows at LBO = 64 rows * 128 B = 8192 B.
// Canonical form ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units (cute mma_traits_sm100.hpp).  Advancing K by 16 = +2048 B.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 (bits 4-5 = 1), A/B fp16 (0), both K-major,
// N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

#define NM_TMEM_LD32(taddr, r)                                                                                          \
  asm volatile(                                                                                                         \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"  \
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                        \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),     \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),          \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),         \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                       \
      : "r"(taddr)                                                                                                      \
      : "memory")
#define NM_TMEM_LD16(taddr, r)                                                                                          \
  asm volatile(                                                                                                         \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"          \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),     \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                        \
      : "r"(taddr)                                                                                                      \
      : "memory")
#define NM_TMEM_ST16(taddr, r)                                                                                         \
  asm volatile(                                                                                                        \
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(  \
          taddr),                                                                                                      \
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),    \
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])                                           \
      : "memory")
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// {lo16 = fp16(a), hi16 = fp16(b)}, round-to-nearest, saturating to +-65504 (never inf).
__device__ __forceinline__ uint32_t pack_f16x2_sat(float a, float b) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}

}  // namespace ptx
}  // namespace nm
