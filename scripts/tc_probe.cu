// Bring-up probe for the tcgen05 building blocks used by the tensor-core conv kernel:
// no-swizzle K-major smem descriptors ([k8][row][8] layout), row-shifted A operand,
// TMEM alloc / tcgen05.mma / commit / tcgen05.ld.   nvcc -arch=sm_100a -o tc_probe tc_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
  return d;
}

// D[128 x N] = A[rows shift..shift+127][K] * B[N][K]^T ; A smem layout [K/8][R][8], B smem layout [K/8][N][8]
__global__ void probe_kernel(const __half* A, const __half* B, float* D, int R, int N, int K, int shift, int swap_lbo_sbo) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __half* sA = reinterpret_cast<__half*>(smem);
  __half* sB = sA + (size_t)(K / 8) * R * 8;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;

  for (int i = tid; i < (K / 8) * R * 8; i += blockDim.x) sA[i] = A[i];
  for (int i = tid; i < (K / 8) * N * 8; i += blockDim.x) sB[i] = B[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the MMA (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t lboA = R * 16, sboA = 128, lboB = N * 16, sboB = 128;
    if (swap_lbo_sbo) { uint32_t t = lboA; lboA = sboA; sboA = t; t = lboB; lboB = sboB; sboB = t; }
    for (int kk = 0; kk < K / 16; ++kk) {
      const uint64_t ad = make_desc(smem_u32(sA) + shift * 16 + kk * 2 * R * 16, lboA, sboA);
      const uint64_t bd = make_desc(smem_u32(sB) + kk * 2 * N * 16, lboB, sboB);
      const uint32_t acc = kk > 0;
      asm volatile(
          "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_base),
          "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // everyone waits for the MMAs
  asm volatile(
      "{\n.reg .pred p;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra LAB_DONE;\nbra LAB_WAIT;\nLAB_DONE:\n}" ::"r"(
          smem_u32(&bar)),
      "r"(0)
      : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4) {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t r[8];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 8; ++i) D[(size_t)row * N + c0 + i] = __uint_as_float(r[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
  }
}

static int run(int N, int K, int shift, int swap) {
  const int R = 128 + 56;
  std::vector<__half> hA((size_t)(K / 8) * R * 8), hB((size_t)(K / 8) * N * 8);
  std::vector<float> fA((size_t)R * K), fB((size_t)N * K);
  srand(1);
  for (int r = 0; r < R; ++r)
    for (int k = 0; k < K; ++k) {
      float v = (float)((rand() % 7) - 3);
      fA[(size_t)r * K + k] = v;
      hA[((size_t)(k / 8) * R + r) * 8 + k % 8] = __float2half(v);
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      float v = (float)((rand() % 5) - 2);
      fB[(size_t)n * K + k] = v;
      hB[((size_t)(k / 8) * N + n) * 8 + k % 8] = __float2half(v);
    }
  __half *dA, *dB;
  float* dD;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dD, (size_t)128 * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, (size_t)128 * N * 4);
  size_t smem = hA.size() * 2 + hB.size() * 2;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 128, smem>>>(dA, dB, dD, R, N, K, shift, swap);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("N=%d K=%d shift=%d swap=%d: CUDA error %s\n", N, K, shift, swap, cudaGetErrorString(e));
    return 2;
  }
  std::vector<float> hD((size_t)128 * N);
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)fA[(size_t)(m + shift) * K + k] * fB[(size_t)n * K + k];
      double err = fabs(s - hD[(size_t)m * N + n]);
      if (!(err < 1e-3)) ++bad;
      if (err > maxerr || err != err) maxerr = err;
    }
  printf("N=%3d K=%3d shift=%2d swap=%d: %s (bad %d / %d, maxerr %g)\n", N, K, shift, swap, bad ? "MISMATCH" : "ok", bad,
         128 * N, maxerr);
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
  return bad != 0;
}

int main() {
  int fails = 0;
  for (int N : {16, 32, 64, 128, 256})
    for (int shift : {0, 1, 25, 50}) fails += run(N, N == 16 ? 16 : 32, shift, 0);
  fails += run(256, 64, 7, 0);
  printf("fails=%d\n", fails);
  return 0;
}
