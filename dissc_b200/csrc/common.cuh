// Shared device helpers: mbarrier + 1-D bulk TMA (cp.async.bulk), error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/dissc_b200.h"

namespace dissc {

// ---- error plumbing (thread-local message behind dissc_last_error) -------
extern thread_local char g_err[512];
int set_err(int code, const char* fmt, ...);

#define DISSC_CUDA(expr)                                                                        \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return ::dissc::set_err(DISSC_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                              \
  } while (0)

#define DISSC_CHECK(cond, code, ...)                       \
  do {                                                     \
    if (!(cond)) return ::dissc::set_err(code, __VA_ARGS__); \
  } while (0)

constexpr int kThreads = 256;

// ---- PTX wrappers -------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// Make mbarrier.init visible to the async (TMA) proxy before the first bulk copy.
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint expires)
// instead of re-issuing the try_wait / branch pair every ~60 cycles -- ncu counted 46 M such spin iterations per launch of
// the C = 64 pair kernel (40 % of all issued instructions), pure power on a power-capped part.
#ifndef DISSC_MBAR_SUSPEND_NS
#define DISSC_MBAR_SUSPEND_NS 2000
#endif
constexpr uint32_t kMbarSuspendNs = DISSC_MBAR_SUSPEND_NS;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(kMbarSuspendNs)
      : "memory");
}

// 1-D bulk TMA: global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- thread-block clusters (2 CTAs sharing one weight stream, conv_tc.cuh) ----------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// all threads of all CTAs of the cluster (also orders shared-memory / mbarrier initialisation before remote accesses)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory offset in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// arrive (release at cluster scope) on an mbarrier of another CTA of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// mbar_wait whose acquire covers writes released by the other CTA of the cluster
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(kMbarSuspendNs)
      : "memory");
}
// 1-D bulk TMA delivered to the SAME shared-memory offset (data and mbarrier) of every CTA in `cta_mask`: one L2 read
// feeds all of them.
__device__ __forceinline__ void tma_load_1d_multicast(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                                      uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}

// One lane of a CONVERGED warp (elect.sync).  Single-thread regions that issue tcgen05.mma / tcgen05.commit / bulk TMA
// must be entered through this and not through `lane == 0`: the compiler then knows exactly one thread runs the region
// and emits the uniform-datapath instructions (UTCHMMA, UTCBAR, UBLKCP) back to back; behind a plain divergent branch
// it wraps EVERY such instruction in an ELECT / BRA.U.ANY waterfall loop (~10 extra instructions per MMA), which made
// the MMA-issue thread the bottleneck of every N <= 128 layer (profiles/README.md, r01_d).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// Programmatic dependent launch (PDL).  A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start
// while its predecessor in the stream is still draining: its prologue (barrier init, TMEM allocation, bias staging) then
// overlaps the predecessor's tail.  pdl_wait() blocks until the predecessor grid has completed and its writes are
// visible -- it must precede EVERY global-memory access that depends on (or could overwrite data read by) the
// predecessor; pdl_launch_dependents() lets the successor start launching.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Embedding-row ids come from the caller (unit ids from a k-means model, speaker ids from a manifest): nn.Embedding
// raises on an id outside the table (device assert on CUDA); here the gather stays inside the table (row 0 is read
// instead) and the handle's error flags -- two ints in mapped pinned host memory, one per kind of id -- are set, which the
// next host-synchronous entry point (dissc_*_status, dissc_gen_forward_host) turns into DISSC_EINDEX.  Plain stores of a
// constant: atomics on host memory need PCIe atomics, which the platform need not provide (compute-sanitizer flags them).
enum { kIdxUnit = 1, kIdxSpeaker = 2 };
__device__ __forceinline__ long long checked_row(long long id, int rows, int* err_flag, int what) {
  if ((unsigned long long)id >= (unsigned long long)rows) {
    if (err_flag) *(reinterpret_cast<volatile int*>(err_flag) + (what == kIdxSpeaker ? 1 : 0)) = 1;
    return 0;
  }
  return id;
}

// leaky-relu for 0 <= slope <= 1 (the reference uses 0.1 and 0.01): max(v, v*slope) is two instructions, bit-identical
// to the select form for every finite v.
__device__ __forceinline__ float leaky(float v, float slope) { return fmaxf(v, v * slope); }

}  // namespace dissc
