// CTA-pair (cta_group::2) helpers: cluster rank / sync, remote mbarrier arrive, 2-CTA TMEM alloc,
// pair TMA loads that complete the LEADER's mbarrier, 2-CTA tcgen05.mma and multicast commit.
#pragma once
#include "tc_common.cuh"

namespace dmf {
namespace tc2 {
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
// address of `local_smem_addr` in the shared memory of CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
// Arrive on a (possibly remote) mbarrier of the cluster.  Deliberately NOT .release.cluster: that form compiles to
// MEMBAR.ALL.GPU + ERRBAR (ncu: 17 % of all stall samples of the InfoNCE backward, and in the GEMM epilogue it
// waits for every outstanding global store).  What these arrivals publish is TMEM state, which is ordered by
// tcgen05.wait::ld/st + tcgen05.fence::before_thread_sync on this side and tcgen05.fence::after_thread_sync on
// the consumer side -- the same protocol CUTLASS uses for its tmem-empty barriers.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(smem_result)),
               "r"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(kCols) : "memory");
}
// TMA load issued by either CTA of the pair; completes the mbarrier of the LEADER (peer bit cleared)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* d, int c0, int c1, uint64_t* bar) {
  const uint32_t bar_leader = tc::smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          tc::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(d)), "r"(bar_leader), "r"(c0), "r"(c1)
      : "memory");
}
// The same load MULTICAST to the CTAs in `mask` (same shared-memory offset in each); every destination CTA's bytes
// complete the mbarrier at this offset in the leader of ITS pair.
__device__ __forceinline__ void tma_load_2d_pair_mc(void* smem_dst, const CUtensorMap* d, int c0, int c1, uint64_t* bar,
                                                    uint16_t mask) {
  const uint32_t bar_leader = tc::smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(tc::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(d)), "r"(bar_leader), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_ss2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_ts2(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior MMAs of this thread retired) on the barrier at the same offset in BOTH CTAs
// (mask = cluster ranks that receive the arrival; 3 = the two CTAs of a 2-CTA cluster)
__device__ __forceinline__ void umma_commit2(uint64_t* bar, uint16_t mask = 3) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          tc::smem_u32(bar)),
      "h"(mask)
      : "memory");
}
}  // namespace tc2
}  // namespace dmf
