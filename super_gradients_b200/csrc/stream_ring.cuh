// Per-thread shared-memory ring for the streaming (HBM-bound) per-channel passes.
//
// A thread that keeps per-channel coefficients in registers has none left to hold many loads in flight; here every thread owns
// DEPTH private slots in shared memory, fills them with cp.async (16 bytes, L1 bypassed) DEPTH iterations ahead and reads back only
// what it copied itself -- so no barrier is involved, the bytes in flight per SM are DEPTH x U x NIN x threads x 16 whatever the
// register count, and channel slices of wider buffers cost nothing extra because every thread still forms its own addresses.
#pragma once
#include <stdint.h>

#include "common.cuh"

namespace sgb_ring {

__device__ __forceinline__ void cp16(uint32_t smem_addr, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(gmem) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t smem_addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_addr) : "memory");
  return r;
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

template <int NIN, int U, int D, int THREADS>
constexpr size_t bytes() {
  return (size_t)D * U * NIN * THREADS * 16;
}

// Walks `mine` pixels of one thread: pixel q of the thread reads input j from ptr[j] + q * kstep[j] (ptr[j] == nullptr: input absent,
// its raw vector is unspecified).  body(q, raw) is called for q = 0 .. mine-1 in order.  my_ring = ring base + threadIdx.x * 16;
// slot (d, k, j) lives at + ((d * U + k) * NIN + j) * THREADS * 16, so a warp's slots are consecutive 16-byte words (conflict-free).
template <int NIN, int U, int D, int THREADS, class Body>
__device__ __forceinline__ void walk(const uint32_t my_ring, const bf16* (&ptr)[NIN], const int64_t (&kstep)[NIN], const int64_t mine, Body&& body) {
  const int64_t nfull = mine / U;  // ring iterations of U pixels each
  auto issue = [&](int d) {        // requests the next U pixels into slot d (always called in pixel order)
#pragma unroll
    for (int j = 0; j < NIN; ++j) {
      if (ptr[j]) {
#pragma unroll
        for (int k = 0; k < U; ++k) cp16(my_ring + (uint32_t)(((d * U + k) * NIN + j) * THREADS * 16), ptr[j] + k * kstep[j]);
        ptr[j] += U * kstep[j];
      }
    }
  };
#pragma unroll
  for (int d = 0; d < D; ++d) {
    if (d < nfull) issue(d);
    commit();  // one group per slot whether or not it issued anything: wait<D - 1> then always means "the slot of `it` has landed"
  }
  int d = 0;
  for (int64_t it = 0; it < nfull; ++it) {
    wait<D - 1>();
    uint4 raw[U][NIN];
#pragma unroll
    for (int k = 0; k < U; ++k)
#pragma unroll
      for (int j = 0; j < NIN; ++j) raw[k][j] = lds128(my_ring + (uint32_t)(((d * U + k) * NIN + j) * THREADS * 16));
    if (it + D < nfull) issue(d);  // the slot was just read into registers (same thread, program order)
    commit();
#pragma unroll
    for (int k = 0; k < U; ++k) body(it * U + k, raw[k]);
    d = d + 1 == D ? 0 : d + 1;
  }
  wait<0>();
  for (int64_t q = nfull * U; q < mine; ++q) {  // fewer than U pixels left: plain loads
    uint4 raw[NIN];
#pragma unroll
    for (int j = 0; j < NIN; ++j) {
      raw[j] = ptr[j] ? *reinterpret_cast<const uint4*>(ptr[j]) : make_uint4(0, 0, 0, 0);
      if (ptr[j]) ptr[j] += kstep[j];
    }
    body(q, raw);
  }
}

}  // namespace sgb_ring
