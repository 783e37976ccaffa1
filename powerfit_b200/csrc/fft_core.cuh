// In-register / shared-memory FFT building blocks of the fused path (sm_100a).
//
// Every transform uses the kernel exp(+2 pi i n k / N) (see fft_generic.cu).
//
// Two element types share the same butterfly code:
//   float2 : one complex number (x = re, y = im)                       -- scalar FP32 pipe
//   C2     : TWO complex numbers in split form, re = (re0, re1), im = (im0, im1), each half
//            an aligned 64-bit register pair -> every add/mul/fma below is one packed
//            FADD2 / FMUL2 / FFMA2 (add/mul/fma.rn.f32x2) working on both numbers.  The
//            FP32 pipe does the same lane-work either way, but a packed instruction takes
//            one issue slot for two complex operations, which is what the FFT kernels are
//            short of (profiles/r01_ncu_v2_fused_full.txt: issue-bound, FMA pipe 43 %).
//
// pencil<LANES,E>: an N = LANES*E point transform by LANES adjacent lanes of a warp, E
// points per lane:
//   X[k1 + E k0] = sum_{n0<LANES} W_LANES^(n0 k0) W_N^(n0 k1) sum_{n1<E} x[n0 + LANES n1] W_E^(n1 k1)
// lane n0 does the E-point DFT in registers, multiplies by W_N^(n0 k1), the lanes exchange
// through the pencil's own shared-memory storage (XOR swizzle, conflict free) and lane u
// finishes with the LANES-point DFTs of k1 = u, u+LANES, ...; its outputs are
// X[u + LANES m], m < E -- the distribution the inputs had, natural order, no digit reversal.
//
// Row transforms whose two packed numbers must belong to the SAME pencil use one split
// radix-2 step around a packed N/2-point transform:
//   split-in  -> adjacent-out (DIF): in (x[n], x[n+N/2])  -> out (X[2k], X[2k+1])
//   adjacent-in -> split-out  (DIT): in (x[2n], x[2n+1])  -> out (X[k], X[k+N/2])
// so a 2-D plane can keep ONE layout (pairs along y, or pairs (ky, ky+N/2)) through the
// row pass, the column pass (two independent columns per C2) and the row pass back.
#pragma once
#include <cuda_runtime.h>

namespace pfb {

typedef unsigned long long u64_t;

// ------------------------------------------------------------------ packed f32x2 primitives
__device__ __forceinline__ float2 padd(float2 a, float2 b) {
    float2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<u64_t *>(&r))
        : "l"(*reinterpret_cast<const u64_t *>(&a)), "l"(*reinterpret_cast<const u64_t *>(&b)));
    return r;
}
__device__ __forceinline__ float2 psub(float2 a, float2 b) {
    float2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<u64_t *>(&r))
        : "l"(*reinterpret_cast<const u64_t *>(&a)), "l"(*reinterpret_cast<const u64_t *>(&b)));
    return r;
}
__device__ __forceinline__ float2 pmul(float2 a, float2 b) {
    float2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<u64_t *>(&r))
        : "l"(*reinterpret_cast<const u64_t *>(&a)), "l"(*reinterpret_cast<const u64_t *>(&b)));
    return r;
}
__device__ __forceinline__ float2 pfma(float2 a, float2 b, float2 c) {
    float2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*reinterpret_cast<u64_t *>(&r))
        : "l"(*reinterpret_cast<const u64_t *>(&a)), "l"(*reinterpret_cast<const u64_t *>(&b)),
          "l"(*reinterpret_cast<const u64_t *>(&c)));
    return r;
}
__device__ __forceinline__ float2 pneg(float2 a) { return make_float2(-a.x, -a.y); }   // folds into operand modifiers
__device__ __forceinline__ float2 pdup(float s) { return make_float2(s, s); }

struct C2 {
    float2 re, im;
};
__device__ __forceinline__ C2 c2_from(float4 v) { C2 r; r.re = make_float2(v.x, v.y); r.im = make_float2(v.z, v.w); return r; }
__device__ __forceinline__ float4 c2_to(C2 v) { return make_float4(v.re.x, v.re.y, v.im.x, v.im.y); }
// 128-bit moves of a C2 as two 64-bit registers: ptxas then places (re, im) in one aligned
// quad instead of assembling it with MOVs around every access
__device__ __forceinline__ C2 lds_c2(const float4 *p) {
    C2 r;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(*reinterpret_cast<u64_t *>(&r.re)), "=l"(*reinterpret_cast<u64_t *>(&r.im))
                 : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return r;
}
// Stores name their four 32-bit halves: with two 64-bit operands ptxas assembled every stored quad with four MOVs
// in ONE scratch quad -- 281 MOVs in kernel B (11 % of its instructions), each group waiting for the previous store
// to release the quad (7 % of the stall samples, profiles/r02_ncu_v3_fused_full.txt) -- while with four 32-bit
// operands it allocates the producing FADD2/FFMA2 results in place (35 MOVs).
__device__ __forceinline__ void sts_c2(float4 *p, C2 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"((unsigned)__cvta_generic_to_shared(p)),
                 "f"(v.re.x), "f"(v.re.y), "f"(v.im.x), "f"(v.im.y) : "memory");
}
__device__ __forceinline__ C2 ldg_c2(const float4 *p) {
    C2 r;
    asm volatile("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(*reinterpret_cast<u64_t *>(&r.re)), "=l"(*reinterpret_cast<u64_t *>(&r.im))
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_c2(float4 *p, C2 v) {
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.re.x), "f"(v.re.y), "f"(v.im.x), "f"(v.im.y)
                 : "memory");
}
__device__ __forceinline__ C2 c2_zero() { C2 r; r.re = make_float2(0.f, 0.f); r.im = make_float2(0.f, 0.f); return r; }

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int K> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(K)); }

// ------------------------------------------------------------------ complex algebra, both types
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmulf(float2 a, float2 w) {
    return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
}
__device__ __forceinline__ float2 mul_i(float2 a) { return make_float2(-a.y, a.x); }
// a * (c + i s) with compile-time-ish constants
__device__ __forceinline__ float2 rotc(float2 a, float c, float s) {
    return make_float2(a.x * c - a.y * s, a.x * s + a.y * c);
}
// a * h(1+i), a * h(-1+i)
__device__ __forceinline__ float2 rot45(float2 a) {
    const float h = 0.70710678118654752440f;
    return make_float2(h * (a.x - a.y), h * (a.x + a.y));
}
__device__ __forceinline__ float2 rot135(float2 a) {
    const float h = 0.70710678118654752440f;
    return make_float2(h * (-a.x - a.y), h * (a.x - a.y));
}

__device__ __forceinline__ C2 cadd(C2 a, C2 b) { C2 r; r.re = padd(a.re, b.re); r.im = padd(a.im, b.im); return r; }
__device__ __forceinline__ C2 csub(C2 a, C2 b) { C2 r; r.re = psub(a.re, b.re); r.im = psub(a.im, b.im); return r; }
__device__ __forceinline__ C2 mul_i(C2 a) { C2 r; r.re = pneg(a.im); r.im = a.re; return r; }
__device__ __forceinline__ C2 rotc(C2 a, float c, float s) {
    C2 r;
    r.re = pfma(pneg(a.im), pdup(s), pmul(a.re, pdup(c)));
    r.im = pfma(a.re, pdup(s), pmul(a.im, pdup(c)));
    return r;
}
__device__ __forceinline__ C2 rot45(C2 a) {
    const float2 h = pdup(0.70710678118654752440f);
    C2 r; r.re = pmul(psub(a.re, a.im), h); r.im = pmul(padd(a.re, a.im), h); return r;
}
__device__ __forceinline__ C2 rot135(C2 a) {
    const float2 h = pdup(0.70710678118654752440f);
    C2 r; r.re = pmul(padd(a.re, a.im), pneg(h)); r.im = pmul(psub(a.re, a.im), h); return r;
}
// both numbers times their own factor: w = (re0, re1, im0, im1); a "dup" twiddle is (c, c, s, s)
__device__ __forceinline__ C2 cmul(C2 a, C2 w) {
    C2 r;
    r.re = pfma(pneg(a.im), w.im, pmul(a.re, w.re));
    r.im = pfma(a.re, w.im, pmul(a.im, w.re));
    return r;
}
__device__ __forceinline__ C2 cmul4(C2 a, float4 w) {
    const float2 wr = make_float2(w.x, w.y), wi = make_float2(w.z, w.w);
    C2 r;
    r.re = pfma(pneg(a.im), wi, pmul(a.re, wr));
    r.im = pfma(a.re, wi, pmul(a.im, wr));
    return r;
}

// ------------------------------------------------------------------ register DFTs
// 4-point DFT; results X0..X3 land in a, b, c, d
template <class T> __device__ __forceinline__ void dft4(T &a, T &b, T &c, T &d) {
    const T s0 = cadd(a, c), d0 = csub(a, c), s1 = cadd(b, d), d1 = mul_i(csub(b, d));
    a = cadd(s0, s1);
    c = csub(s0, s1);
    b = cadd(d0, d1);
    d = csub(d0, d1);
}

template <class T> __device__ __forceinline__ void dft8(T (&v)[8]) {
    dft4(v[0], v[2], v[4], v[6]);
    dft4(v[1], v[3], v[5], v[7]);
    const T t0 = v[1], t1 = rot45(v[3]), t2 = mul_i(v[5]), t3 = rot135(v[7]);
    const T e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    v[0] = cadd(e0, t0); v[4] = csub(e0, t0);
    v[1] = cadd(e1, t1); v[5] = csub(e1, t1);
    v[2] = cadd(e2, t2); v[6] = csub(e2, t2);
    v[3] = cadd(e3, t3); v[7] = csub(e3, t3);
}

template <class T> __device__ __forceinline__ void dft16(T (&v)[16]) {
#pragma unroll
    for (int lo = 0; lo < 4; ++lo) dft4(v[lo], v[lo + 4], v[lo + 8], v[lo + 12]);
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
    // v[lo + 4 k1] *= W16^(lo k1)
    v[5] = rotc(v[5], c1, s1);        // 1*1
    v[9] = rot45(v[9]);               // 1*2 -> W16^2
    v[13] = rotc(v[13], s1, c1);      // 1*3
    v[6] = rot45(v[6]);               // 2*1
    v[10] = mul_i(v[10]);             // 2*2 -> W16^4
    v[14] = rot135(v[14]);            // 2*3 -> W16^6
    v[7] = rotc(v[7], s1, c1);        // 3*1
    v[11] = rot135(v[11]);            // 3*2 -> W16^6
    v[15] = rotc(v[15], -c1, -s1);    // 3*3 -> W16^9
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
    // X[k1 + 4 k2] sits in v[4 k1 + k2]: transpose the 4x4 register tile
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) { const T tmp = v[4 * a + b]; v[4 * a + b] = v[4 * b + a]; v[4 * b + a] = tmp; }
}

__device__ __forceinline__ float2 cscale(float2 a, float f) { return make_float2(a.x * f, a.y * f); }
__device__ __forceinline__ C2 cscale(C2 a, float f) { C2 r; r.re = pmul(a.re, pdup(f)); r.im = pmul(a.im, pdup(f)); return r; }
__device__ __forceinline__ float2 cneg(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ C2 cneg(C2 a) { C2 r; r.re = pneg(a.re); r.im = pneg(a.im); return r; }

// 3-point DFT, W3 = exp(+2 pi i / 3) = -1/2 + i sqrt(3)/2; results X0..X2 land in a, b, c
template <class T> __device__ __forceinline__ void dft3(T &a, T &b, T &c) {
    const T s = cadd(b, c), d = mul_i(cscale(csub(b, c), 0.86602540378443864676f));
    const T m = csub(a, cscale(s, 0.5f));
    a = cadd(a, s);
    b = cadd(m, d);
    c = csub(m, d);
}

// 24-point DFT, natural order in and out: n = n0 + 3 n1, k = k1 + 8 k0
template <class T> __device__ __forceinline__ void dft24(T (&v)[24]) {
    T t[3][8];
#pragma unroll
    for (int n0 = 0; n0 < 3; ++n0) {
        T u[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) u[n1] = v[n0 + 3 * n1];
        dft8(u);
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) t[n0][k1] = u[k1];
    }
    // t[n0][k1] *= W24^(n0 k1)
    t[1][1] = rotc(t[1][1], 0.96592582628906831f, 0.25881904510252074f);   // W24^1
    t[1][2] = rotc(t[1][2], 0.86602540378443871f, 0.49999999999999994f);   // W24^2
    t[1][3] = rot45(t[1][3]);
    t[1][4] = rotc(t[1][4], 0.50000000000000011f, 0.8660254037844386f);   // W24^4
    t[1][5] = rotc(t[1][5], 0.25881904510252074f, 0.96592582628906831f);   // W24^5
    t[1][6] = mul_i(t[1][6]);
    t[1][7] = rotc(t[1][7], -0.25881904510252063f, 0.96592582628906831f);   // W24^7
    t[2][1] = rotc(t[2][1], 0.86602540378443871f, 0.49999999999999994f);   // W24^2
    t[2][2] = rotc(t[2][2], 0.50000000000000011f, 0.8660254037844386f);   // W24^4
    t[2][3] = mul_i(t[2][3]);
    t[2][4] = rotc(t[2][4], -0.49999999999999978f, 0.86602540378443871f);   // W24^8
    t[2][5] = rotc(t[2][5], -0.86602540378443871f, 0.49999999999999994f);   // W24^10
    t[2][6] = cneg(t[2][6]);
    t[2][7] = rotc(t[2][7], -0.86602540378443882f, -0.49999999999999972f);   // W24^14
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) {
        dft3(t[0][k1], t[1][k1], t[2][k1]);
#pragma unroll
        for (int k0 = 0; k0 < 3; ++k0) v[k1 + 8 * k0] = t[k0][k1];
    }
}

// 12-point DFT, natural order in and out: n = n0 + 3 n1, k = k1 + 4 k0
template <class T> __device__ __forceinline__ void dft12(T (&v)[12]) {
    T t[3][4];
#pragma unroll
    for (int n0 = 0; n0 < 3; ++n0) {
#pragma unroll
        for (int n1 = 0; n1 < 4; ++n1) t[n0][n1] = v[n0 + 3 * n1];
        dft4(t[n0][0], t[n0][1], t[n0][2], t[n0][3]);
    }
    // t[n0][k1] *= W12^(n0 k1)
    t[1][1] = rotc(t[1][1], 0.86602540378443865f, 0.5f);                    // W12^1
    t[1][2] = rotc(t[1][2], 0.5f, 0.86602540378443865f);                    // W12^2
    t[1][3] = mul_i(t[1][3]);                                               // W12^3
    t[2][1] = rotc(t[2][1], 0.5f, 0.86602540378443865f);                    // W12^2
    t[2][2] = rotc(t[2][2], -0.5f, 0.86602540378443865f);                   // W12^4
    t[2][3] = cneg(t[2][3]);                                                // W12^6
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        dft3(t[0][k1], t[1][k1], t[2][k1]);
#pragma unroll
        for (int k0 = 0; k0 < 3; ++k0) v[k1 + 4 * k0] = t[k0][k1];
    }
}

template <int E, class T> struct DftReg;
template <class T> struct DftReg<4, T> { static __device__ __forceinline__ void run(T (&v)[4]) { dft4(v[0], v[1], v[2], v[3]); } };
template <class T> struct DftReg<8, T> { static __device__ __forceinline__ void run(T (&v)[8]) { dft8(v); } };
template <class T> struct DftReg<12, T> { static __device__ __forceinline__ void run(T (&v)[12]) { dft12(v); } };
template <class T> struct DftReg<16, T> { static __device__ __forceinline__ void run(T (&v)[16]) { dft16(v); } };
template <class T> struct DftReg<24, T> { static __device__ __forceinline__ void run(T (&v)[24]) { dft24(v); } };

// ------------------------------------------------------------------ scalar pencil (L lanes x E)
// in : v[n1] = x[t + L n1]       out: v[m] = X[t + L m]
// scratch[p * stride], p < L*E, is the pencil's shared-memory storage (clobbered).
template <int E, int L = 8>
__device__ __forceinline__ void fft_pencil(float2 (&v)[E], float2 *scratch, int stride, int t,
                                           const float2 (&tw)[E], bool active) {
    static_assert(E % L == 0, "E must be a multiple of L");
    DftReg<E, float2>::run(v);
#pragma unroll
    for (int k1 = 1; k1 < E; ++k1) v[k1] = cmulf(v[k1], tw[k1]);
    __syncwarp();
    if (active) {
#pragma unroll
        for (int k1 = 0; k1 < E; ++k1) scratch[(k1 * L + (t ^ (k1 & (L - 1)))) * stride] = v[k1];
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < E / L; ++q) {
        float2 a[L];
        if (active) {
#pragma unroll
            for (int n0 = 0; n0 < L; ++n0) a[n0] = scratch[((t + L * q) * L + (n0 ^ t)) * stride];
        } else {
#pragma unroll
            for (int n0 = 0; n0 < L; ++n0) a[n0] = make_float2(0.f, 0.f);
        }
        DftReg<L, float2>::run(a);
#pragma unroll
        for (int k0 = 0; k0 < L; ++k0) v[(E / L) * k0 + q] = a[k0];
    }
    __syncwarp();
}

template <int E>
__device__ __forceinline__ void load_twiddles(float2 (&tw)[E], const float2 *__restrict__ twN, int t) {
#pragma unroll
    for (int k1 = 0; k1 < E; ++k1) tw[k1] = __ldg(twN + t * k1);      // t*k1 < L*E = N
}

// ------------------------------------------------------------------ packed pencil (LANES x E), two numbers per element
// in : v[n1] = x[t + LANES n1]   out: v[m] = X[t + LANES m]          (both packed numbers alike)
// scratch[p * stride] (float4), p < LANES*E, is the pencil's own shared storage (clobbered);
// lanes without work must be handed a private dummy pencil rather than being predicated off.
// Twiddles W_N^(t k1) are plain (c, s) pairs -- FMUL2/FFMA2 take a 32-bit register as a
// broadcast operand, so nothing is duplicated -- and come from a provider tw(k1): registers
// (TwReg) where they fit, or this lane's column of a [k1][t] table in shared memory (TwSmem:
// the lanes of a pencil read one contiguous 8*LANES-byte run, conflict free).
template <int E> struct TwReg {
    static constexpr bool paired = false;
    const float2 (&w)[E];
    __device__ __forceinline__ float2 operator()(int k1) const { return w[k1]; }
};
// this lane's column of a table stored in PAIRS, [k1 / 2][t] of float4 = (W^(t 2k), W^(t (2k+1))): one 16-byte
// load fetches two twiddles (the shared-memory instruction queue, not its bandwidth, is what kernel B runs out of)
template <int LANES> struct TwSmemPair {
    static constexpr bool paired = true;
    const float4 *p;      // &table[0][t]
    __device__ __forceinline__ float4 pair(int k) const {
        float4 r;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                     : "r"((unsigned)__cvta_generic_to_shared(p + k * LANES)));
        return r;
    }
};
template <int LANES> struct TwSmem {
    static constexpr bool paired = false;
    const float2 *p;      // &table[0][t]
    __device__ __forceinline__ float2 operator()(int k1) const {
        float2 r;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y)
                     : "r"((unsigned)__cvta_generic_to_shared(p + k1 * LANES)));
        return r;
    }
};
// both packed numbers times the same w = c + i s
__device__ __forceinline__ C2 cmulw(C2 a, float2 w) {
    C2 r;
    r.re = pfma(pneg(a.im), pdup(w.y), pmul(a.re, pdup(w.x)));
    r.im = pfma(a.re, pdup(w.y), pmul(a.im, pdup(w.x)));
    return r;
}

// stage 1: E-point DFTs in registers, twiddles, swizzled write to the pencil's storage
template <int LANES, int E, class TW>
__device__ __forceinline__ void pencil2_stage1(C2 (&v)[E], float4 *scratch, int stride, int t, const TW &tw,
                                               bool active = true) {
    static_assert(E % LANES == 0, "E must be a multiple of LANES");
    DftReg<E, C2>::run(v);
    if constexpr (TW::paired) {
#pragma unroll
        for (int k = 0; k < E / 2; ++k) {
            const float4 w = tw.pair(k);
            if (k > 0) v[2 * k] = cmulw(v[2 * k], make_float2(w.x, w.y));
            v[2 * k + 1] = cmulw(v[2 * k + 1], make_float2(w.z, w.w));
        }
    } else {
#pragma unroll
        for (int k1 = 1; k1 < E; ++k1) v[k1] = cmulw(v[k1], tw(k1));
    }
    __syncwarp();
    if (active) {
#pragma unroll
        for (int k1 = 0; k1 < E; ++k1) sts_c2(scratch + (k1 * LANES + (t ^ (k1 & (LANES - 1)))) * stride, v[k1]);
    }
    __syncwarp();
}
// stage 2, chunk q < E/LANES: the LANES-point DFT of k1 = t + LANES q; a[k0] = X[t + LANES (q + (E/LANES) k0)]
template <int LANES>
__device__ __forceinline__ void pencil2_stage2(C2 (&a)[LANES], const float4 *scratch, int stride, int t, int q,
                                               bool active = true) {
    if (active) {
#pragma unroll
        for (int n0 = 0; n0 < LANES; ++n0) a[n0] = lds_c2(scratch + ((t + LANES * q) * LANES + (n0 ^ t)) * stride);
    } else {
#pragma unroll
        for (int n0 = 0; n0 < LANES; ++n0) a[n0] = c2_zero();
    }
    DftReg<LANES, C2>::run(a);
}

// A pencil group without work (a row outside the support box next to rows inside it) passes active = false:
// it takes part in the warp barriers but touches no shared memory.
template <int LANES, int E, class TW>
__device__ __forceinline__ void fft_pencil2(C2 (&v)[E], float4 *scratch, int stride, int t, const TW &tw,
                                            bool active = true) {
    pencil2_stage1<LANES, E>(v, scratch, stride, t, tw, active);
#pragma unroll
    for (int q = 0; q < E / LANES; ++q) {
        C2 a[LANES];
        pencil2_stage2<LANES>(a, scratch, stride, t, q, active);
#pragma unroll
        for (int k0 = 0; k0 < LANES; ++k0) v[(E / LANES) * k0 + q] = a[k0];
    }
    __syncwarp();
}

// the same transform with every output multiplied by a per-element factor (the map spectrum):
// factors of chunk 0 are fetched before the transform starts, those of the later chunks while
// stage 2 runs, so that neither the registers nor the load latency pile up.
// f points at this lane's first factor; factor of output m is f[LANES * m].
template <int LANES, int E, class TW>
__device__ __forceinline__ void fft_pencil2_mul(C2 (&v)[E], float4 *scratch, int stride, int t, const TW &tw,
                                                const float4 *__restrict__ f) {
    constexpr int Q = E / LANES;
    C2 f0[LANES];
#pragma unroll
    for (int k0 = 0; k0 < LANES; ++k0) f0[k0] = ldg_c2(f + LANES * (Q * k0));
    pencil2_stage1<LANES, E>(v, scratch, stride, t, tw);
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        C2 fn[LANES];
        if (q + 1 < Q) {
#pragma unroll
            for (int k0 = 0; k0 < LANES; ++k0) fn[k0] = ldg_c2(f + LANES * (Q * k0 + q + 1));
        }
        C2 a[LANES];
        pencil2_stage2<LANES>(a, scratch, stride, t, q);
#pragma unroll
        for (int k0 = 0; k0 < LANES; ++k0) v[Q * k0 + q] = cmul(a[k0], f0[k0]);
        if (q + 1 < Q) {
#pragma unroll
            for (int k0 = 0; k0 < LANES; ++k0) f0[k0] = fn[k0];
        }
    }
    __syncwarp();
}

// Same, for wide pencils (LANES = 16) where holding every factor across the transform would
// spill: the factors are fetched in two halves around stage 2.
template <int LANES, int E, class TW>
__device__ __forceinline__ void fft_pencil2_mul_late(C2 (&v)[E], float4 *scratch, int stride, int t, const TW &tw,
                                                     const float4 *__restrict__ f) {
    constexpr int Q = E / LANES, HL = LANES / 2;
    pencil2_stage1<LANES, E>(v, scratch, stride, t, tw);
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        C2 fa[HL];
#pragma unroll
        for (int k0 = 0; k0 < HL; ++k0) fa[k0] = ldg_c2(f + LANES * (Q * k0 + q));
        C2 a[LANES];
        pencil2_stage2<LANES>(a, scratch, stride, t, q);
#pragma unroll
        for (int k0 = 0; k0 < HL; ++k0) v[Q * k0 + q] = cmul(a[k0], fa[k0]);
#pragma unroll
        for (int k0 = 0; k0 < HL; ++k0) fa[k0] = ldg_c2(f + LANES * (Q * (k0 + HL) + q));
#pragma unroll
        for (int k0 = 0; k0 < HL; ++k0) v[Q * (k0 + HL) + q] = cmul(a[k0 + HL], fa[k0]);
    }
    __syncwarp();
}

// adjacent-in -> split-out row transform of N = 2*LANES*E points (decimation in time):
// in : v[n1] = (x[2n], x[2n+1]),  n = t + LANES n1        out: v[m] = (X[k], X[k+N/2]),  k = t + LANES m
// twh[m] = W_N^(t + LANES m)
template <int LANES, int E, class TW>
__device__ __forceinline__ void fft_row_adj2split(C2 (&v)[E], float4 *scratch, int stride, int t,
                                                  const TW &tw, const float2 (&twh)[E], bool active = true) {
    fft_pencil2<LANES, E>(v, scratch, stride, t, tw, active);
#pragma unroll
    for (int m = 0; m < E; ++m) {
        const float2 o = cmulf(make_float2(v[m].re.y, v[m].im.y), twh[m]);
        const float er = v[m].re.x, ei = v[m].im.x;
        v[m].re = make_float2(er + o.x, er - o.x);
        v[m].im = make_float2(ei + o.y, ei - o.y);
    }
}

// split-in -> adjacent-out row transform (decimation in frequency):
// in : v[n1] = (x[n], x[n+N/2]),  n = t + LANES n1        out: v[m] = (X[2k], X[2k+1]),  k = t + LANES m
template <int LANES, int E, class TW>
__device__ __forceinline__ void fft_row_split2adj(C2 (&v)[E], float4 *scratch, int stride, int t,
                                                  const TW &tw, const float2 (&twh)[E]) {
#pragma unroll
    for (int n1 = 0; n1 < E; ++n1) {
        const float2 a = make_float2(v[n1].re.x, v[n1].im.x), b = make_float2(v[n1].re.y, v[n1].im.y);
        const float2 d = cmulf(csub(a, b), twh[n1]);
        v[n1].re = make_float2(a.x + b.x, d.x);
        v[n1].im = make_float2(a.y + b.y, d.y);
    }
    fft_pencil2<LANES, E>(v, scratch, stride, t, tw);
}

}  // namespace pfb
