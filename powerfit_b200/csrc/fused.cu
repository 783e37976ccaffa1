// Fused three-kernel search pipeline for cubic grids of N = 64 or 128 voxels.
//
//   A  fused_rotate_fftx  : rotate template+mask (a PAIR of rotations packed as one complex
//                           signal: re = rotation a, im = rotation b), transform along x,
//                           write X1[pair][sig][z][kx][y]            (replaces K1/K2 + 1/3 of K5)
//   B  fused_fftyz_mul    : one (kx, output volume, pair) plane per CTA, 128 KB of shared
//                           memory: forward y, forward z, multiply with FT(map) or FT(map^2),
//                           inverse z, inverse y, write X2[pair][vol][z][kx][y]
//                                                                    (2/3 of K5, K6, 2/3 of K7)
//   C  fused_ifftx_lcc    : inverse x of the gcc/ave/ave2 rows, LCC normalisation and the
//                           running arg-max in registers/shared memory (1/3 of K7, K8-K10)
// (K numbers: SURVEY.md section 2a; reference code powerfitter.py:487-531, kernels.cl:162-226.)
//
// Only two HBM round trips per transform remain, the rotated grids and the gcc/ave/ave2
// grids never exist in memory, and -- because template and mask are compact -- kernels A
// and B skip every (y,z) row outside the template's support box (rows that are exactly
// zero), which removes most of the forward traffic.
//
// All transforms use the kernel exp(+2 pi i k r / N) (see fft_generic.cu).  A pencil of N
// points is handled by 8 threads of one warp holding E = N/8 points each:
//   X[k1 + E k0] = sum_{n0<8} W8^(n0 k0) W_N^(n0 k1) sum_{n1<E} x[n0 + 8 n1] W_E^(n1 k1)
// thread n0 does the E-point DFT in registers, multiplies by W_N^(n0 k1), the eight threads
// exchange through the pencil's own shared-memory storage (XOR-swizzled, conflict free) and
// thread u finishes with the 8-point DFTs of k1 = u, u+8, ...; its outputs are X[u + 8m],
// i.e. the same distribution the inputs had, in natural order -- no digit reversal.
#include "common.cuh"
#include "rotate_device.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace pfb {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmulf(float2 a, float2 w) {
    return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
}
__device__ __forceinline__ float2 mul_i(float2 a) { return make_float2(-a.y, a.x); }

// 4-point DFT, kernel exp(+2 pi i nk/4); results X0..X3 land in a, b, c, d
__device__ __forceinline__ void dft4(float2 &a, float2 &b, float2 &c, float2 &d) {
    const float2 s0 = cadd(a, c), d0 = csub(a, c), s1 = cadd(b, d), d1 = mul_i(csub(b, d));
    a = cadd(s0, s1);
    c = csub(s0, s1);
    b = cadd(d0, d1);
    d = csub(d0, d1);
}

__device__ __forceinline__ void dft8(float2 (&v)[8]) {
    dft4(v[0], v[2], v[4], v[6]);
    dft4(v[1], v[3], v[5], v[7]);
    const float h = 0.70710678118654752440f;
    const float2 t0 = v[1];
    const float2 t1 = make_float2(h * (v[3].x - v[3].y), h * (v[3].x + v[3].y));
    const float2 t2 = mul_i(v[5]);
    const float2 t3 = make_float2(h * (-v[7].x - v[7].y), h * (v[7].x - v[7].y));
    const float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    v[0] = cadd(e0, t0); v[4] = csub(e0, t0);
    v[1] = cadd(e1, t1); v[5] = csub(e1, t1);
    v[2] = cadd(e2, t2); v[6] = csub(e2, t2);
    v[3] = cadd(e3, t3); v[7] = csub(e3, t3);
}

__device__ __forceinline__ void dft16(float2 (&v)[16]) {
#pragma unroll
    for (int lo = 0; lo < 4; ++lo) dft4(v[lo], v[lo + 4], v[lo + 8], v[lo + 12]);
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
    // v[lo + 4 k1] *= W16^(lo k1)
    v[5] = cmulf(v[5], make_float2(c1, s1));                                  // 1*1
    v[9] = make_float2(h * (v[9].x - v[9].y), h * (v[9].x + v[9].y));         // 1*2 -> W16^2
    v[13] = cmulf(v[13], make_float2(s1, c1));                                // 1*3
    v[6] = make_float2(h * (v[6].x - v[6].y), h * (v[6].x + v[6].y));         // 2*1
    v[10] = mul_i(v[10]);                                                     // 2*2 -> W16^4
    v[14] = make_float2(h * (-v[14].x - v[14].y), h * (v[14].x - v[14].y));   // 2*3 -> W16^6
    v[7] = cmulf(v[7], make_float2(s1, c1));                                  // 3*1
    v[11] = make_float2(h * (-v[11].x - v[11].y), h * (v[11].x - v[11].y));   // 3*2 -> W16^6
    v[15] = cmulf(v[15], make_float2(-c1, -s1));                              // 3*3 -> W16^9
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
    // X[k1 + 4 k2] sits in v[4 k1 + k2]: transpose the 4x4 register tile
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) { const float2 tmp = v[4 * a + b]; v[4 * a + b] = v[4 * b + a]; v[4 * b + a] = tmp; }
}

template <int E> __device__ __forceinline__ void dft_reg(float2 (&v)[E]);
template <> __device__ __forceinline__ void dft_reg<8>(float2 (&v)[8]) { dft8(v); }
template <> __device__ __forceinline__ void dft_reg<16>(float2 (&v)[16]) { dft16(v); }

// N = 8E point transform of one pencil by the 8 threads t = 0..7 of a warp octet.
// in : v[n1] = x[t + 8 n1]       out: v[m] = X[t + 8 m]
// scratch[p * stride], p < N, is the pencil's shared-memory storage (clobbered).
template <int E>
__device__ __forceinline__ void fft_pencil(float2 (&v)[E], float2 *scratch, int stride, int t,
                                           const float2 (&tw)[E], bool active) {
    dft_reg<E>(v);
#pragma unroll
    for (int k1 = 1; k1 < E; ++k1) v[k1] = cmulf(v[k1], tw[k1]);
    __syncwarp();
    if (active) {
#pragma unroll
        for (int k1 = 0; k1 < E; ++k1) scratch[(k1 * 8 + (t ^ (k1 & 7))) * stride] = v[k1];
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < E / 8; ++q) {
        float2 a[8];
        if (active) {
#pragma unroll
            for (int n0 = 0; n0 < 8; ++n0) a[n0] = scratch[((t + 8 * q) * 8 + (n0 ^ t)) * stride];
        } else {
#pragma unroll
            for (int n0 = 0; n0 < 8; ++n0) a[n0] = make_float2(0.f, 0.f);
        }
        dft8(a);
#pragma unroll
        for (int k0 = 0; k0 < 8; ++k0) v[(E / 8) * k0 + q] = a[k0];
    }
    __syncwarp();
}

template <int E>
__device__ __forceinline__ void load_twiddles(float2 (&tw)[E], const float2 *__restrict__ twN, int t) {
#pragma unroll
    for (int k1 = 0; k1 < E; ++k1) tw[k1] = __ldg(twN + t * k1);      // t*k1 < 8E = N
}

// ------------------------------------------------------------------------------- kernel A
template <int N>
__global__ void __launch_bounds__(256)
fused_rotate_fftx_kernel(const float *__restrict__ tmpl, const float *__restrict__ mask,
                         const double *__restrict__ rot, int first, int count, int nsig,
                         float2 *__restrict__ X1, const float2 *__restrict__ twN, int rs, int rs2,
                         unsigned ymask, int nzv) {
    constexpr int E = N / 8, TP = 33;
    extern __shared__ float2 smem[];
    float2 *tile_t = smem, *tile_m = smem + N * TP;
    const int pair = blockIdx.y;
    const int j = blockIdx.x % nzv;
    int yt_rank = blockIdx.x / nzv, ytile = 0;
    for (unsigned mbits = ymask;; ++ytile) {
        if (mbits & 1u) { if (yt_rank == 0) break; --yt_rank; }
        mbits >>= 1;
    }
    const int z = (j - rs + N) % N, y0 = 32 * ytile;
    const GridDims d{N, N, N, N / 2, (long)N * N * N};
    const int ra = first + 2 * pair;
    const bool have_b = 2 * pair + 1 < count;
    const double *Ra = rot + (long)ra * 9, *Rb = Ra + 9;
    const int oz = z <= N / 2 ? z : z - N;

    // ---- gather the 32 x N tile of both signals (zeros outside the sphere / support)
    //      zero-fill first, then visit only the x offsets inside the support box [xlo, rs];
    //      offset -N/2 is skipped: it aliases index N/2, which belongs to offset +N/2.
    for (int idx = threadIdx.x; idx < N * TP; idx += 256) {
        tile_t[idx] = make_float2(0.f, 0.f);
        tile_m[idx] = make_float2(0.f, 0.f);
    }
    __syncthreads();
    const int xlo = max(-rs, -(N / 2 - 1)), W = rs - xlo + 1;
    const int lim2 = min(rs2, (N / 2) * (N / 2));
    for (int idx = threadIdx.x; idx < 32 * W; idx += 256) {
        const int r = idx / W, ox = idx % W + xlo;
        const int iy = y0 + r;
        const int oy = iy <= N / 2 ? iy : iy - N;
        if (ox * ox + oy * oy + oz * oz <= lim2) {
            float2 tv = make_float2(0.f, 0.f), mv = make_float2(0.f, 0.f);
            const SrcCoord ca = source_coord(Ra, ox, oy, oz);
            tv.x = sample_trilinear(tmpl, d, ca);
            mv.x = sample_nearest(mask, d, ca);
            if (have_b) {
                const SrcCoord cb = source_coord(Rb, ox, oy, oz);
                tv.y = sample_trilinear(tmpl, d, cb);
                mv.y = sample_nearest(mask, d, cb);
            }
            const int x = ox < 0 ? ox + N : ox;
            tile_t[x * TP + r] = tv;
            tile_m[x * TP + r] = mv;
        }
    }
    __syncthreads();

    // ---- x transforms: thread (row r, t)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & 7, r = warp + 8 * (lane >> 3);
    float2 tw[E];
    load_twiddles<E>(tw, twN, t);
    float2 v[E], v2[E];
#pragma unroll
    for (int n1 = 0; n1 < E; ++n1) v[n1] = tile_t[(t + 8 * n1) * TP + r];
    fft_pencil<E>(v, tile_t + r, TP, t, tw, true);
#pragma unroll
    for (int m = 0; m < E; ++m) tile_t[(t + 8 * m) * TP + r] = v[m];
#pragma unroll
    for (int n1 = 0; n1 < E; ++n1) {
        v[n1] = tile_m[(t + 8 * n1) * TP + r];
        v2[n1] = make_float2(v[n1].x * v[n1].x, v[n1].y * v[n1].y);
    }
    fft_pencil<E>(v, tile_m + r, TP, t, tw, true);
#pragma unroll
    for (int m = 0; m < E; ++m) tile_m[(t + 8 * m) * TP + r] = v[m];
    __syncthreads();

    // ---- coalesced write-out: 32 consecutive y (256 B) per kx
    const size_t plane = (size_t)N * N;
    float2 *o_t = X1 + ((size_t)(pair * nsig + 0) * N + z) * plane + y0;      // X1[pair][sig][z][kx][y]
    float2 *o_m = X1 + ((size_t)(pair * nsig + 1) * N + z) * plane + y0;
    for (int idx = threadIdx.x; idx < 32 * N; idx += 256) {
        const int kx = idx >> 5, rr = idx & 31;
        o_t[(size_t)kx * N + rr] = tile_t[kx * TP + rr];
        o_m[(size_t)kx * N + rr] = tile_m[kx * TP + rr];
    }
    if (nsig == 3) {
        __syncthreads();
        fft_pencil<E>(v2, tile_t + r, TP, t, tw, true);
#pragma unroll
        for (int m = 0; m < E; ++m) tile_t[(t + 8 * m) * TP + r] = v2[m];
        __syncthreads();
        float2 *o_2 = X1 + ((size_t)(pair * nsig + 2) * N + z) * plane + y0;
        for (int idx = threadIdx.x; idx < 32 * N; idx += 256) {
            const int kx = idx >> 5, rr = idx & 31;
            o_2[(size_t)kx * N + rr] = tile_t[kx * TP + rr];
        }
    }
}

// ------------------------------------------------------------------------------- kernel B
template <int N, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
fused_fftyz_mul_kernel(const float2 *__restrict__ X1, float2 *__restrict__ X2, const float2 *__restrict__ Fq,
                       const float2 *__restrict__ F2q, const float2 *__restrict__ twN, int rs, unsigned ymask,
                       int nsig) {
    constexpr int E = N / 8, P = N + 1;
    extern __shared__ float2 plane[];
    const int kx = blockIdx.x, vol = blockIdx.y, pair = blockIdx.z;
    const int sig = vol == 0 ? 0 : (vol == 1 ? 1 : nsig - 1);
    const size_t pl = (size_t)N * N;
    const float2 *src = X1 + (size_t)(pair * nsig + sig) * N * pl + (size_t)kx * N;   // + z*pl + y
    const float2 *Fm = (vol == 2 ? F2q : Fq) + (size_t)kx * pl;
    float2 *dst = X2 + (size_t)(pair * 3 + vol) * N * pl + (size_t)kx * N;            // + z*pl + y
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = THREADS / 32;
    const int t = lane & 7, c = lane >> 3;
    float2 tw[E];
    load_twiddles<E>(tw, twN, t);

    // ---- phase 1: forward y of the rows inside the support box, global -> shared
    const int nzv = min(2 * rs + 1, N);
    const int ntask1 = ((nzv + 31) / 32) * 8;
    for (int w = warp; w < ntask1; w += NW) {
        const int j = 32 * (w >> 3) + (w & 7) + 8 * c;
        const bool act = j < nzv;
        const int z = act ? (j - rs + N) % N : 0;
        float2 v[E];
#pragma unroll
        for (int n1 = 0; n1 < E; ++n1) {
            const int y = t + 8 * n1;
            v[n1] = (act && ((ymask >> (y >> 5)) & 1u)) ? __ldg(src + (size_t)z * pl + y) : make_float2(0.f, 0.f);
        }
        fft_pencil<E>(v, plane + z * P, 1, t, tw, act);
        if (act) {
#pragma unroll
            for (int m = 0; m < E; ++m) plane[z * P + t + 8 * m] = v[m];
        }
    }
    __syncthreads();

    // ---- phase 2: forward z, multiply with the map spectrum, inverse z (columns ky)
    for (int w = warp; w < N / 4; w += NW) {
        const int ky = 32 * (w >> 3) + (w & 7) + 8 * c;
        float2 f[E];
#pragma unroll
        for (int m = 0; m < E; ++m) f[m] = __ldg(Fm + (size_t)ky * N + t + 8 * m);
        float2 v[E];
#pragma unroll
        for (int n1 = 0; n1 < E; ++n1) {
            const int z = t + 8 * n1;
            const int sz = z <= N / 2 ? z : z - N;
            v[n1] = (sz >= -rs && sz <= rs) ? plane[z * P + ky] : make_float2(0.f, 0.f);
        }
        fft_pencil<E>(v, plane + ky, P, t, tw, true);
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = cmulf(v[m], f[m]);
        fft_pencil<E>(v, plane + ky, P, t, tw, true);
#pragma unroll
        for (int m = 0; m < E; ++m) plane[(t + 8 * m) * P + ky] = v[m];
    }
    __syncthreads();

    // ---- phase 3: inverse y of every row, shared -> global
    for (int w = warp; w < N / 4; w += NW) {
        const int z = 32 * (w >> 3) + (w & 7) + 8 * c;
        float2 v[E];
#pragma unroll
        for (int n1 = 0; n1 < E; ++n1) v[n1] = plane[z * P + t + 8 * n1];
        fft_pencil<E>(v, plane + z * P, 1, t, tw, true);
#pragma unroll
        for (int m = 0; m < E; ++m) dst[(size_t)z * pl + t + 8 * m] = v[m];
    }
}

// ------------------------------------------------------------------------------- kernel C
__device__ __forceinline__ void fold_best(int64_t &best, float gcc, float ave, float sd_arg_ave2, float norm,
                                          uint32_t rot) {
    const float var = __fsub_rn(__fmul_rn(sd_arg_ave2, norm), __fmul_rn(ave, ave));
    const float lcc = __fdiv_rn(gcc, __fsqrt_rn(var));
    if (lcc == lcc) {
        const int64_t key = pack_best(__float_as_uint(lcc), rot);
        if (key > best) best = key;
    }
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int K> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(K)); }

// One CTA owns a tile of 32 rows (fixed z, 32 consecutive y) and walks a chunk of rotation
// pairs.  Per pair the three X2 tiles (ave2, ave, gcc) stream through a double-buffered
// cp.async pipeline; thread (row r, t) transforms its row along x and keeps sqrt(var) of
// its 16 voxels in registers until the gcc row arrives.  Tile layout [kx][34]: pitch 34
// keeps 16-byte cp.async destinations aligned and, with the two pencils of a half-warp on
// adjacent rows, every shared access conflict free.
template <int N>
__global__ void __launch_bounds__(256, 2)
fused_ifftx_lcc_kernel(const float2 *__restrict__ X2, const uint8_t *__restrict__ lcc_mask, float norm,
                       int first_index, int count, int pairs_per_chunk, int64_t *__restrict__ best,
                       const float2 *__restrict__ twN) {
    constexpr int E = N / 8, TP = 34, BP = N + 1;
    extern __shared__ float2 smem[];
    float2 *tiles[2] = {smem, smem + N * TP};
    int64_t *lbest = reinterpret_cast<int64_t *>(smem + 2 * N * TP);
    const int y0 = 32 * blockIdx.x, z = blockIdx.y;
    const int npairs = (count + 1) / 2;
    const int p0 = blockIdx.z * pairs_per_chunk, p1 = min(npairs, p0 + pairs_per_chunk);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & 7, r = 4 * warp + (lane >> 3);
    float2 tw[E];
    load_twiddles<E>(tw, twN, t);
    const size_t pl = (size_t)N * N;
    const size_t row = ((size_t)z * N + y0 + r) * N;
    unsigned mbits = 0;
#pragma unroll
    for (int m = 0; m < E; ++m) {
        if (lcc_mask[row + t + 8 * m]) mbits |= 1u << m;
        lbest[r * BP + t + 8 * m] = kBestInit;
    }
    const int nitems = 3 * (p1 - p0);
    auto prefetch = [&](int item) {
        const int p = p0 + item / 3, vol = 2 - item % 3;              // ave2, ave, gcc
        const float2 *src = X2 + ((size_t)(p * 3 + vol) * N + z) * pl + y0;
        float2 *dst = tiles[item & 1];
        for (int idx = threadIdx.x; idx < 16 * N; idx += 256) {
            const int kx = idx >> 4, ch = idx & 15;
            cp_async16(dst + kx * TP + 2 * ch, src + (size_t)kx * N + 2 * ch);
        }
        cp_async_commit();
    };
    if (nitems > 0) prefetch(0);
    float2 sd[E];
    for (int item = 0; item < nitems; ++item) {
        if (item + 1 < nitems) { prefetch(item + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        float2 *tile = tiles[item & 1];
        const int p = p0 + item / 3, vi = item % 3;
        float2 v[E];
#pragma unroll
        for (int n1 = 0; n1 < E; ++n1) v[n1] = tile[(t + 8 * n1) * TP + r];
        fft_pencil<E>(v, tile + r, TP, t, tw, true);
        if (vi == 0) {
#pragma unroll
            for (int m = 0; m < E; ++m) sd[m] = v[m];                  // ave2
        } else if (vi == 1) {
#pragma unroll
            for (int m = 0; m < E; ++m) {                              // 1/sqrt(N ave2 - ave^2)
                // var <= 0 gives inf / NaN exactly where the reference's gcc/sqrt(var) does
                sd[m].x = rsqrtf(__fsub_rn(__fmul_rn(sd[m].x, norm), __fmul_rn(v[m].x, v[m].x)));
                sd[m].y = rsqrtf(__fsub_rn(__fmul_rn(sd[m].y, norm), __fmul_rn(v[m].y, v[m].y)));
            }
        } else {
            const uint32_t ia = (uint32_t)(first_index + 2 * p);
            const bool have_b = 2 * p + 1 < count;
#pragma unroll
            for (int m = 0; m < E; ++m) {
                if ((mbits >> m) & 1u) {
                    int64_t b = lbest[r * BP + t + 8 * m];
                    const float la = __fmul_rn(v[m].x, sd[m].x), lb = __fmul_rn(v[m].y, sd[m].y);
                    if (la == la) { const int64_t k = pack_best(__float_as_uint(la), ia); if (k > b) b = k; }
                    if (have_b && lb == lb) { const int64_t k = pack_best(__float_as_uint(lb), ia + 1); if (k > b) b = k; }
                    lbest[r * BP + t + 8 * m] = b;
                }
            }
        }
        __syncthreads();       // everyone is done with this buffer before it is refilled
    }
#pragma unroll
    for (int m = 0; m < E; ++m) {
        if ((mbits >> m) & 1u) {
            const int64_t b = lbest[r * BP + t + 8 * m];
            if (b > kBestInit) atomicMax(reinterpret_cast<long long *>(best + row + t + 8 * m), (long long)b);
        }
    }
}

// ------------------------------------------------------------------------------- helpers
// Fq[kx][ky][kz] = F[kz][ky][kx]
__global__ void transpose_zx_kernel(const float2 *__restrict__ F, float2 *__restrict__ Fq, int N) {
    __shared__ float2 tl[32][33];
    const int ky = blockIdx.z;
    const int kx0 = blockIdx.x * 32, kz0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
        tl[i][threadIdx.x] = F[((size_t)(kz0 + i) * N + ky) * N + kx0 + threadIdx.x];
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
        Fq[((size_t)(kx0 + i) * N + ky) * N + kz0 + threadIdx.x] = tl[threadIdx.x][i];
}

// max squared distance from voxel 0 (periodic) of any voxel where template or mask is non-zero
__global__ void support_kernel(const float *__restrict__ tmpl, const float *__restrict__ mask, int nz, int ny,
                               int nx, int *out) {
    const long V = (long)nz * ny * nx;
    int best = -1;
    for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long)gridDim.x * blockDim.x) {
        if (tmpl[v] != 0.f || mask[v] != 0.f) {
            int x = (int)(v % nx), y = (int)((v / nx) % ny), z = (int)(v / ((long)nx * ny));
            x = min(x, nx - x); y = min(y, ny - y); z = min(z, nz - z);
            best = max(best, x * x + y * y + z * z);
        }
    }
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0 && best >= 0) atomicMax(out, best);
}

// ------------------------------------------------------------------------------- host side
template <int N> static int fused_init_n() {
    constexpr int TP = 33;
    PFB_CUDA(cudaFuncSetAttribute(fused_rotate_fftx_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(2 * N * TP * sizeof(float2))));
    PFB_CUDA(cudaFuncSetAttribute(fused_fftyz_mul_kernel<N, (N >= 128 ? 512 : 256)>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(N * (N + 1) * sizeof(float2))));
    PFB_CUDA(cudaFuncSetAttribute(fused_ifftx_lcc_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(2 * N * 34 * sizeof(float2) + 32 * (N + 1) * sizeof(int64_t))));
    return PFB_OK;
}

bool fused_supported(int nz, int ny, int nx) { return nz == ny && ny == nx && (nx == 64 || nx == 128); }

int fused_init(Plan *p) {
    if (p->nx == 64) return fused_init_n<64>();
    return fused_init_n<128>();
}

int fused_prepare_target(Plan *p, cudaStream_t s) {
    const int N = p->nx;
    dim3 grid(N / 32, N / 32, N), block(32, 8);
    { LaunchScope ls(p, KC_OTHER, s); transpose_zx_kernel<<<grid, block, 0, s>>>(p->F, p->Fq, N); }
    { LaunchScope ls(p, KC_OTHER, s); transpose_zx_kernel<<<grid, block, 0, s>>>(p->F2, p->F2q, N); }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

// support box of the template/mask pair (synchronises the stream: it needs one int back)
int fused_prepare_template(Plan *p, cudaStream_t s) {
    int *d_r2 = reinterpret_cast<int *>(p->best_scratch);
    int init = -1;
    PFB_CUDA(cudaMemcpyAsync(d_r2, &init, sizeof(int), cudaMemcpyHostToDevice, s));
    { LaunchScope ls(p, KC_OTHER, s);
      support_kernel<<<p->sm_count * 4, 256, 0, s>>>(p->tmpl, p->mask, p->nz, p->ny, p->nx, d_r2); }
    int r2 = -1;
    PFB_CUDA(cudaMemcpyAsync(&r2, d_r2, sizeof(int), cudaMemcpyDeviceToHost, s));
    PFB_CUDA(cudaStreamSynchronize(s));
    const int N = p->nx, rmax = N / 2;
    // a rotated sample at offset r can be non-zero only if |r| <= max_r + sqrt(3)
    double reach = (r2 < 0 ? 0.0 : sqrt((double)r2)) + 1.7320508075688772 + 1e-6;
    int rs = (int)floor(reach);
    if (rs > rmax) rs = rmax;
    double reach2 = reach * reach;
    p->rs = rs;
    p->rs2 = reach2 > (double)rmax * rmax ? rmax * rmax : (int)floor(reach2);
    if (const char *e = getenv("PFB_NO_PRUNE")) { if (atoi(e)) { p->rs = rmax; p->rs2 = rmax * rmax; } }
    unsigned ymask = 0;
    for (int tile = 0; tile < N / 32; ++tile)
        for (int y = 32 * tile; y < 32 * tile + 32; ++y) {
            const int sy = y <= N / 2 ? y : y - N;
            if (sy >= -p->rs && sy <= p->rs) ymask |= 1u << tile;
        }
    p->ymask = ymask;
    return PFB_OK;
}

template <int N>
static int fused_batch_n(Plan *p, int first, int count, int rot_index_offset, int64_t *best, cudaStream_t s) {
    constexpr int TP = 33;
    constexpr int BT = N >= 128 ? 512 : 256;
    const int npairs = (count + 1) / 2;
    const int nzv = std::min(2 * p->rs + 1, N);
    const int nyt = __builtin_popcount(p->ymask);
    {
        LaunchScope ls(p, KC_FUSED_A, s);
        fused_rotate_fftx_kernel<N><<<dim3(nzv * nyt, npairs), 256, 2 * N * TP * sizeof(float2), s>>>(
            p->tmpl, p->mask, p->rot_dev, first, count, p->nsig, p->A, p->tw[0], p->rs, p->rs2, p->ymask, nzv);
    }
    {
        LaunchScope ls(p, KC_FUSED_B, s);
        fused_fftyz_mul_kernel<N, BT><<<dim3(N, 3, npairs), BT, N * (N + 1) * sizeof(float2), s>>>(
            p->A, p->B, p->Fq, p->F2q, p->tw[0], p->rs, p->ymask, p->nsig);
    }
    {
        // enough CTAs for ~4 waves: split the pair loop into chunks
        const int tiles = (N / 32) * N;
        int chunks = std::max(1, std::min(npairs, (4 * p->sm_count * 2 + tiles - 1) / tiles));
        const int ppc = (npairs + chunks - 1) / chunks;
        chunks = (npairs + ppc - 1) / ppc;
        LaunchScope ls(p, KC_FUSED_C, s);
        fused_ifftx_lcc_kernel<N><<<dim3(N / 32, N, chunks), 256,
                                    2 * N * 34 * sizeof(float2) + 32 * (N + 1) * sizeof(int64_t), s>>>(
            p->B, p->lcc_mask, p->norm_factor, rot_index_offset + first, count, ppc, best, p->tw[0]);
    }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

int fused_batch(Plan *p, int first, int count, int rot_index_offset, int64_t *best, cudaStream_t s) {
    if (p->nx == 64) return fused_batch_n<64>(p, first, count, rot_index_offset, best, s);
    return fused_batch_n<128>(p, first, count, rot_index_offset, best, s);
}

}  // namespace pfb
