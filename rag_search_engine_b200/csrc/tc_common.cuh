// tc_common.cuh — constants and PTX wrappers shared by the tcgen05 probe/filter kernel (knn_tc3.cuh) and the
// refine kernel (knn_refine.cuh): mbarrier / cluster helpers, the UMMA shared-memory descriptor, tcgen05
// fences, commits and TMEM loads.  (r01 kept two superseded TF32 mappings of the same GEMM, knn_tc.cuh and
// knn_tc2.cuh, selectable as tc modes 3-6; they were 2.6-3.5x slower than the fp16-shadow kernel on every
// measured shape and were removed in r02 — their measurements stay in DESIGN.md §5.)
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "knn_scan.cuh"

namespace rse {

constexpr int kTcBN = 256;                                 // queries per tensor-core pass (UMMA M of the CTA pair)
constexpr float kTcEps = 2.5e-3f;                          // bound on |cos~ - cos| (knn_tc3.cuh header)
constexpr int kTcCandCap = 8192;                           // survivors kept per query
constexpr int kTcRefineCap = 4096;                         // rows re-scored exactly per query at most (r01: 2048; a dense
                                                           // neighbourhood of near-duplicates needs the head-room)

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 128 B, 8-row groups of 1024 B)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);          // start address, bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                              // LBO (unused for swizzled K-major) = 1
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                      // SBO = 1024 B
  d |= static_cast<uint64_t>(1) << 46;                              // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                              // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* local_bar, uint32_t target_rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local_bar)), "r"(target_rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void tc2_commit_mc(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

}  // namespace rse
