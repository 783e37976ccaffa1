// TMA (cp.async.bulk.tensor) + mbarrier helpers of the fused path (sm_100a).
//
// Kernel C reads its [kx][y-pair] tiles out of the work buffer X2 as 4-D boxes of a tensor map
// (dims, fastest first: floats of a (z,kx) row | kx | z | pair*3+volume).  One elected thread arms an
// mbarrier with the box size and issues the copy; the copy engine writes the box into shared memory in
// the 128-byte swizzle (16-byte chunk index XOR row index mod 8), which is exactly the conflict-free
// arrangement the x pencils need: lane t of a pencil owns the rows = t (mod 8).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pfb {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// makes the initialised barriers visible to the async proxy (the copy engine)
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "PFB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra PFB_DONE;\n\t"
        "bra PFB_WAIT;\n\t"
        "PFB_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr uint64_t kL2EvictFirst = 0x12F0000000000000ull;      // createpolicy.fractional.L2::evict_first, fraction 1.0

// one 4-D box, global -> shared, completion counted in bytes on `bar`; X2 is read exactly once: evict first
__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *tmap, uint64_t *bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "l"(kL2EvictFirst)
        : "memory");
}

// Host: tensor map of a work buffer laid out [outer][z][kx][row_floats] float32, box = box_floats x box_kx x 1 x 1,
// 128-byte swizzle for 128-byte box rows, 64-byte swizzle for 64-byte ones.  The driver entry point is looked up at run time so that the library does not link libcuda.
int make_x2_tensor_map(CUtensorMap *out, const void *base, int row_floats, int nkx, int nz, long outer,
                       int box_floats, int box_kx);
// Host: the same 4-D view of the work buffer X1, box = one whole row x 1 kx x box_z consecutive z, no swizzle
// (kernel B's staging copies).
int make_x1_tensor_map(CUtensorMap *out, const void *base, int row_floats, int nkx, int nz, long outer, int box_z);

}  // namespace pfb
