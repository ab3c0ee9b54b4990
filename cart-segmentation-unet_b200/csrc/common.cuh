// Shared device helpers for the sm_100a kernels: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld) wrappers and UMMA descriptor builders.
// Everything here is inline PTX; no CUTLASS/CuTe dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cs {

#define CS_DEVINL __device__ __forceinline__

CS_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

CS_DEVINL uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred;
}

// ----------------------------------------------------------------------------- mbarrier
CS_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
CS_DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
CS_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

CS_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
CS_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
CS_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a lost arrival turns into a trap (reported as a launch failure) instead of a
// hang that would take the GPU box with it.
CS_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FFFu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > (1ll << 32)) { asm volatile("trap;"); }   // ~2 s at 1.9 GHz
    }
  }
}

// ----------------------------------------------------------------------------- TMA
CS_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
CS_DEVINL void tma_load_2d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
CS_DEVINL void tma_load_4d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
CS_DEVINL void tma_store_4d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
CS_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
CS_DEVINL void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
CS_DEVINL void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------- tcgen05
template <int COLS>
CS_DEVINL void tmem_alloc(uint32_t* dst_smem) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
CS_DEVINL void tmem_dealloc(uint32_t addr) {      // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}
CS_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
CS_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate; issued by ONE thread.
CS_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Whole-warp variants: every lane executes the (convergent) code, one elected lane issues the instruction.  With
// warp-uniform operands ptxas keeps the descriptors on the uniform datapath; the `if (lane == 0)` form costs an
// ELECT / R2UR.BROADCAST / BRA.U.ANY loop per instruction (~180 cycles per MMA measured with ncu on B200).
template <bool PAIR>
CS_DEVINL void umma_bf16_warp(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if (PAIR) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// NK consecutive K-steps of one operand pair in a single block: the descriptor low words advance by STEP (in
// 16-byte units) per step; the first step overwrites the accumulator when accumulate_first == 0.
// The descriptor high word is the same for every operand used here: SBO = 1024 B, version 1, 128B swizzle.
static constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);
template <bool PAIR, int NK, int STEP>
CS_DEVINL void umma_bf16_steps_warp(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                    uint32_t accumulate_first) {
  static_assert(NK == 4 || NK == 8, "NK");
#define CS_MMA_STEP(i, PRED)                                                                     \
  "add.u32 al, %1, " #i "*%7;\n\t"                                                               \
  "add.u32 bl, %2, " #i "*%7;\n\t"                                                               \
  "mov.b64 da, {al, %3};\n\t"                                                                    \
  "mov.b64 db, {bl, %3};\n\t"                                                                    \
  "@q tcgen05.mma.cta_group::%8.kind::f16 [%0], da, db, %5, " PRED ";\n\t"
  if (NK == 4) {
    asm volatile(
        "{\n\t.reg .pred p, q, t;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        CS_MMA_STEP(0, "p") CS_MMA_STEP(1, "t") CS_MMA_STEP(2, "t") CS_MMA_STEP(3, "t")
        "}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "n"(kDescHiSw128), "n"(0), "r"(idesc), "r"(accumulate_first), "n"(STEP),
        "n"(PAIR ? 2 : 1)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p, q, t;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        CS_MMA_STEP(0, "p") CS_MMA_STEP(1, "t") CS_MMA_STEP(2, "t") CS_MMA_STEP(3, "t")
        CS_MMA_STEP(4, "t") CS_MMA_STEP(5, "t") CS_MMA_STEP(6, "t") CS_MMA_STEP(7, "t")
        "}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "n"(kDescHiSw128), "n"(0), "r"(idesc), "r"(accumulate_first), "n"(STEP),
        "n"(PAIR ? 2 : 1)
        : "memory");
  }
#undef CS_MMA_STEP
}

// Four K-steps with separate descriptor high words for A (a register: SBO / base offset vary) and B (SBO = 1024).
template <bool PAIR>
CS_DEVINL void umma_bf16_steps4_warp_hi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t idesc,
                                        uint32_t accumulate_first) {
#define CS_MMA_STEP2(i, PRED)                                                                    \
  "add.u32 al, %1, " #i "*2;\n\t"                                                                \
  "add.u32 bl, %2, " #i "*2;\n\t"                                                                \
  "mov.b64 da, {al, %7};\n\t"                                                                    \
  "mov.b64 db, {bl, %3};\n\t"                                                                    \
  "@q tcgen05.mma.cta_group::%8.kind::f16 [%0], da, db, %5, " PRED ";\n\t"
  asm volatile(
      "{\n\t.reg .pred p, q, t;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.eq.b32 t, 0, 0;\n\t"
      CS_MMA_STEP2(0, "p") CS_MMA_STEP2(1, "t") CS_MMA_STEP2(2, "t") CS_MMA_STEP2(3, "t")
      "}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "n"(kDescHiSw128), "n"(0), "r"(idesc), "r"(accumulate_first), "r"(a_hi),
      "n"(PAIR ? 2 : 1)
      : "memory");
#undef CS_MMA_STEP2
}

// Eight K-steps (cta_group::1) with separate descriptor high words and per-step advances for A and B.
template <int A_STEP, int B_STEP>
CS_DEVINL void umma_bf16_steps8_warp_ab(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                        uint32_t idesc, uint32_t accumulate_first) {
#define CS_MMA_STEP3(i, PRED)                                                                    \
  "add.u32 al, %1, " #i "*%7;\n\t"                                                               \
  "add.u32 bl, %2, " #i "*%8;\n\t"                                                               \
  "mov.b64 da, {al, %3};\n\t"                                                                    \
  "mov.b64 db, {bl, %4};\n\t"                                                                    \
  "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, " PRED ";\n\t"
  asm volatile(
      "{\n\t.reg .pred p, q, t;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.eq.b32 t, 0, 0;\n\t"
      CS_MMA_STEP3(0, "p") CS_MMA_STEP3(1, "t") CS_MMA_STEP3(2, "t") CS_MMA_STEP3(3, "t")
      CS_MMA_STEP3(4, "t") CS_MMA_STEP3(5, "t") CS_MMA_STEP3(6, "t") CS_MMA_STEP3(7, "t")
      "}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "r"(accumulate_first), "n"(A_STEP), "n"(B_STEP)
      : "memory");
#undef CS_MMA_STEP3
}

template <bool PAIR>
CS_DEVINL void umma_commit_warp(uint64_t* bar) {
  if (PAIR) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
        ::"r"(smem_u32(bar)), "h"((uint16_t)3)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar))
        : "memory");
  }
}
// Arrives on `bar` when all previously issued MMAs of this thread have completed.
CS_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (base_lane + t).
CS_DEVINL void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
CS_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------- CTA pairs (cta_group::2)
// Two CTAs of a cluster on one TPC share one MMA: M = 256 (128 rows from each CTA's smem A tile), B is split by
// N between the two CTAs' shared memories, the accumulator rows live in each CTA's own TMEM.  Only the leader
// (cluster rank 0) issues MMAs; TMA loads of BOTH CTAs complete on the leader's mbarrier; tcgen05.commit is
// multicast to the same barrier offset in both CTAs.
CS_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
CS_DEVINL void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
CS_DEVINL uint32_t mapa_cluster(uint32_t cta_smem_addr, uint32_t rank) {   // address of the same offset in CTA `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_smem_addr), "r"(rank));
  return r;
}
CS_DEVINL void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default (.release.cta) semantics, as CUTLASS' ClusterBarrier::arrive(cta_id): the .release.cluster form costs a
  // MEMBAR.ALL.GPU per arrive (it was the top stall of the epilogue warps in ncu)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Release at CLUSTER scope.  NOT used on any hot path: it compiles to a MEMBAR.ALL.GPU per arrive (~1 us), which paced
// the first stem kernel (one arrive per 128-pixel tile).  The operand builders / transform warps publish their
// shared-memory writes with fence.proxy.async + the plain remote arrive above, like CUTLASS' 2-SM transform pipelines.
CS_DEVINL void mbar_arrive_cluster_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
CS_DEVINL void tma_load_2d_pair(void* smem, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
CS_DEVINL void tma_load_4d_pair(void* smem, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
template <int COLS>
CS_DEVINL void tmem_alloc_pair(uint32_t* dst_smem) {   // the same warp of BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
CS_DEVINL void tmem_dealloc_pair(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}
CS_DEVINL void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives on the barrier at this smem offset in BOTH CTAs when all MMAs issued so far have completed.
CS_DEVINL void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}

// ----------------------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (tcgen05), SWIZZLE_128B.  Fields (cf. the PTX ISA "matrix
// descriptor" table): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) |
// layout=2 (128B swizzle) [61,64).
//  * K-major operand  (rows = M/N index, 64 bf16 = 128 B of K per row): SBO = 1024 (8 rows),
//    LBO ignored by the hardware (encoded as 16 B like CUTLASS does).  Advance 16 elements along K inside the swizzle atom: start += 32 B.
//  * MN-major operand (rows = K index, 64 bf16 = 128 B of M/N per row): SBO = 1024 (8 K rows),
//    LBO = byte distance between consecutive 64-element M/N blocks.  Advance 16 along K: start += 2048 B.
CS_DEVINL uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint64_t make_smem_desc_c(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor, kind::f16: D fp32 [4,6)=1, A bf16 [7,10)=1, B bf16 [10,13)=1,
// A major bit 15, B major bit 16 (0 = K-major, 1 = MN-major), N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ----------------------------------------------------------------------------- misc
CS_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
CS_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
CS_DEVINL float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
CS_DEVINL float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// ----------------------------------------------------------------------------- packed fp32 pairs (sm_100: FFMA2)
// Two fp32 values in one 64-bit register; fma.rn.f32x2 does both lanes in one instruction.  Used where a kernel is
// paced by instruction issue rather than by memory (operand transforms, statistics epilogue, pooled BN backward).
CS_DEVINL uint64_t f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
CS_DEVINL uint64_t bf16x2_to_f32x2(uint32_t w) {
  uint64_t r;
  asm("{\n\t.reg .b32 a, b;\n\tshl.b32 a, %1, 16;\n\tand.b32 b, %1, 0xffff0000;\n\tmov.b64 %0, {a, b};\n\t}" : "=l"(r) : "r"(w));
  return r;
}
CS_DEVINL uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
CS_DEVINL uint32_t relu_pack_bf16x2(uint64_t v) {
  uint32_t r;
  asm("{\n\t.reg .f32 lo, hi;\n\tmov.b64 {lo, hi}, %1;\n\tcvt.rn.relu.bf16x2.f32 %0, hi, lo;\n\t}" : "=r"(r) : "l"(v));
  return r;
}
CS_DEVINL void unpack_f32x2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

}  // namespace cs
