// Tensor memory (TMEM, 256 KB per SM on sm_100a) used as a thread-private stash.
//
// Kernel B needs the forward (y,z) spectrum of a rotated BINARY mask twice: once times FT(map)
// (-> ave) and once times FT(map^2) (-> ave2).  The spectrum plane is as large as the shared
// memory plane it was computed in (131 KB at 128^3), so it cannot wait in shared memory while the
// first product is transformed back in place -- but it fits TMEM, which the search does not use
// otherwise (there is no dense contraction on this path, so no tcgen05.mma).  Every thread parks
// the spectrum values it holds in registers with tcgen05.st and the SAME thread fetches them back
// with tcgen05.ld two block barriers later: no cross-thread layout is involved, the 32x32b shape
// simply maps lane i of a warp to TMEM lane 32 (warp % 4) + i and register j to column base + j.
// Traffic goes over the tensor-memory datapath, not the shared-memory pipe kernel B is bound by.
#pragma once
#include "fft_core.cuh"

namespace pfb {

// one warp allocates ncols (power of two >= 32) columns and publishes the base address through smem
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 8 packed complex pairs (32 registers) of every lane -> 32 consecutive columns at taddr (warp-collective)
__device__ __forceinline__ void tmem_st_c2x8(uint32_t taddr, const C2 (&a)[8]) {
#define PFB_U(k, f) __float_as_uint(a[k].f)
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(PFB_U(0, re.x)), "r"(PFB_U(0, re.y)), "r"(PFB_U(0, im.x)), "r"(PFB_U(0, im.y)),
        "r"(PFB_U(1, re.x)), "r"(PFB_U(1, re.y)), "r"(PFB_U(1, im.x)), "r"(PFB_U(1, im.y)),
        "r"(PFB_U(2, re.x)), "r"(PFB_U(2, re.y)), "r"(PFB_U(2, im.x)), "r"(PFB_U(2, im.y)),
        "r"(PFB_U(3, re.x)), "r"(PFB_U(3, re.y)), "r"(PFB_U(3, im.x)), "r"(PFB_U(3, im.y)),
        "r"(PFB_U(4, re.x)), "r"(PFB_U(4, re.y)), "r"(PFB_U(4, im.x)), "r"(PFB_U(4, im.y)),
        "r"(PFB_U(5, re.x)), "r"(PFB_U(5, re.y)), "r"(PFB_U(5, im.x)), "r"(PFB_U(5, im.y)),
        "r"(PFB_U(6, re.x)), "r"(PFB_U(6, re.y)), "r"(PFB_U(6, im.x)), "r"(PFB_U(6, im.y)),
        "r"(PFB_U(7, re.x)), "r"(PFB_U(7, re.y)), "r"(PFB_U(7, im.x)), "r"(PFB_U(7, im.y))
        : "memory");
#undef PFB_U
}

// the inverse; the wait is part of the same statement so that no use of the registers can be scheduled before it
__device__ __forceinline__ void tmem_ld_c2x8(uint32_t taddr, C2 (&a)[8]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        a[k].re = make_float2(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]));
        a[k].im = make_float2(__uint_as_float(r[4 * k + 2]), __uint_as_float(r[4 * k + 3]));
    }
}

// 16 packed pairs (64 registers, columns taddr .. taddr + 63) with ONE tcgen05.ld and one wait: half the exposed
// latency of two x32 loads each followed by its own wait
__device__ __forceinline__ void tmem_ld_c2x16(uint32_t taddr, C2 (&a)[16]) {
    uint32_t r[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
          "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
          "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
          "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
          "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        a[k].re = make_float2(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]));
        a[k].im = make_float2(__uint_as_float(r[4 * k + 2]), __uint_as_float(r[4 * k + 3]));
    }
}

// fft_pencil2_mul (fft_core.cuh) with the spectrum parked in / fetched from TMEM:
//   FETCH = false  transform; if `park`, store the spectrum chunk by chunk (tcgen05.st); multiply by the factors f
//   FETCH = true   no transform: fetch the parked spectrum (tcgen05.ld), multiply by f
// tcol = this warp's first TMEM column (lane quarter included); chunk q of the pencil uses columns
// [4 LANES q, 4 LANES (q+1)).  The caller issues tmem_wait_st() before the next block barrier.
template <bool FETCH, int LANES, int E, class TW>
__device__ __forceinline__ void fft_pencil2_mul_stash(C2 (&v)[E], float4 *scratch, int stride, int t, const TW &tw,
                                                      const float4 *__restrict__ f, uint32_t tcol, bool park) {
    static_assert(LANES == 8, "the TMEM helpers move 8 packed pairs at a time");
    constexpr int Q = E / LANES;
    C2 f0[LANES];
#pragma unroll
    for (int k0 = 0; k0 < LANES; ++k0) f0[k0] = ldg_c2(f + LANES * (Q * k0));
    if (!FETCH) pencil2_stage1<LANES, E>(v, scratch, stride, t, tw);
    if (FETCH && Q == 2) {
        // the whole parked pencil with one load: chunk q sits in columns [32 q, 32 q + 32)
        C2 fn[LANES];
#pragma unroll
        for (int k0 = 0; k0 < LANES; ++k0) fn[k0] = ldg_c2(f + LANES * (Q * k0 + 1));
        C2 all[16];
        tmem_ld_c2x16(tcol, all);
#pragma unroll
        for (int k0 = 0; k0 < LANES; ++k0) {
            v[Q * k0] = cmul(all[k0], f0[k0]);
            v[Q * k0 + 1] = cmul(all[LANES + k0], fn[k0]);
        }
        __syncwarp();
        return;
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        C2 fn[LANES];
        if (q + 1 < Q) {
#pragma unroll
            for (int k0 = 0; k0 < LANES; ++k0) fn[k0] = ldg_c2(f + LANES * (Q * k0 + q + 1));
        }
        C2 a[LANES];
        if (FETCH) {
            tmem_ld_c2x8(tcol + 4 * LANES * q, a);
        } else {
            pencil2_stage2<LANES>(a, scratch, stride, t, q);
            if (park) tmem_st_c2x8(tcol + 4 * LANES * q, a);
        }
#pragma unroll
        for (int k0 = 0; k0 < LANES; ++k0) v[Q * k0 + q] = cmul(a[k0], f0[k0]);
        if (q + 1 < Q) {
#pragma unroll
            for (int k0 = 0; k0 < LANES; ++k0) f0[k0] = fn[k0];
        }
    }
    __syncwarp();
}

}  // namespace pfb
