// sm_100a building blocks of the streaming kernels: 1-D bulk asynchronous copies (the TMA engine,
// cp.async.bulk) completing on shared-memory mbarriers, L2 eviction policies, scoped
// acquire/release accesses and a bounded spin (no kernel of this library can hang a GPU: a wait
// that lasts longer than FC_SPIN_TIMEOUT_NS traps, which surfaces as a CUDA error on the host).
#pragma once
#include <cstdint>

__device__ __forceinline__ uint32_t fc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void fc_mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fc_smem_u32(bar)), "r"(count) : "memory");
}
// make freshly initialised barriers visible to the async proxy (the copy engine)
__device__ __forceinline__ void fc_mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fc_mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool fc_mbar_try_wait(unsigned long long *bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(fc_smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// generic-proxy accesses of a staging buffer must be ordered before the copy engine refills it
__device__ __forceinline__ void fc_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ unsigned long long fc_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long fc_policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

__device__ __forceinline__ unsigned long long fc_policy_evict_normal() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// 8-byte loads / stores that carry an L2 eviction policy (the Krylov vectors of a partition that fits the L2 are
// marked evict_last so that the matrix stream, which is evict_first, cannot push them out between two phases)
__device__ __forceinline__ double fc_ld_pol(const double *p, unsigned long long pol) {
  double v;
  asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void fc_st_pol(double *p, double v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}

// global -> shared bulk copy of `bytes` (multiple of 16; both addresses 16-byte aligned), completing
// `bytes` transaction bytes on `bar`
__device__ __forceinline__ void fc_bulk_g2s(void *dst_smem, const void *src, unsigned bytes, unsigned long long *bar,
                                            unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          fc_smem_u32(dst_smem)),
      "l"(src), "r"(bytes), "r"(fc_smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void fc_bulk_g2s(void *dst_smem, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   fc_smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(fc_smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ unsigned long long fc_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

constexpr unsigned long long FC_SPIN_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

// Bounded spin helper: call tick() inside a wait loop.
struct fc_spin_guard {
  unsigned n = 0;
  unsigned long long t0 = 0;
  __device__ __forceinline__ void tick() {
    if ((++n & 0x3fffu) == 0u) {
      const unsigned long long t = fc_globaltimer();
      if (t0 == 0) t0 = t;
      else if (t - t0 > FC_SPIN_TIMEOUT_NS) asm volatile("trap;");
    }
  }
};

__device__ __forceinline__ void fc_mbar_wait(unsigned long long *bar, unsigned parity) {
  fc_spin_guard g;
  while (!fc_mbar_try_wait(bar, parity)) g.tick();
}
