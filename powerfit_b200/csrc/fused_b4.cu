// Kernel B for 128^3 on class-decimated half planes held entirely in registers ("B4"); selected
// with PFB_B4=1 together with the folding kernel A and the combining kernel C of fused_cls.cu.
//
// fused_fftyz_mul_kernel (fused.cu) keeps 16 complex pairs per thread and a whole 128 x 64 plane
// per CTA: six shared-memory exchanges per plane, one CTA of 16 warps per SM, shared-memory
// bandwidth bound.  Holding the plane in registers needs 32 pairs per
// thread at this size (fused_b3.cu: four exchanges, but only 8 warps per SM, measured slower).
// A HALF plane -- one ky residue class b of fused_cls.cu, 128 z x 32 pairs = 4096 pairs -- fits
// 256 threads x 16 pairs: three register stages each way, four exchanges, 128 registers per
// thread, 69 KB of shared memory, i.e. two CTAs = 16 warps per SM:
//
//   warp w owns the rows z = z_lo + 16 n1 with z_lo in {2w, 2w+1}, n1 < 8 (a 16 x 32 slab)
//   S1  thread (row, c0): 16-point transform over r of the row's pairs c = c0 + 2 r, from HBM
//   X1  exchange inside the warp's slab             -> thread (k'', g) holds (c0, n1) of z_lo = 2w + g
//   S2  radix-2 over c0 (completes the packed 32-point row transform), split step to
//       (G[k], G[k+32]), 8-point transforms over n1, twiddle W_128^(z_lo k1)
//   X2  into the CTA's plane buffer (each warp writes its own slab) + block barrier
//   S3  thread (k1 = warp, column = lane): 16-point transform over z_lo (completes z), times the
//       map spectrum, 16-point transform back, twiddle -- written back IN PLACE
//   block barrier, then the mirror image S2', X4, S1', store of the class's partial inverse G_b.
//
// Same factorisation as fused_b3.cu with the row length halved (tools/b3_model.py).
#include "common.cuh"
#include "fft_core.cuh"

#include <algorithm>
#include <cstdlib>

namespace pfb {

__device__ __forceinline__ C2 b4_split(C2 v, float2 w) {         // (E, O) -> (G[k], G[k + 32]), w = W_64^k
    const float2 o = cmulf(make_float2(v.re.y, v.im.y), w);
    const float er = v.re.x, ei = v.im.x;
    C2 r;
    r.re = make_float2(er + o.x, er - o.x);
    r.im = make_float2(ei + o.y, ei - o.y);
    return r;
}
__device__ __forceinline__ C2 b4_unsplit(C2 v, float2 w) {       // (G[k], G[k + 32]) -> (a + b, (a - b) W_64^k)
    const float2 a = make_float2(v.re.x, v.im.x), b = make_float2(v.re.y, v.im.y);
    const float2 d = cmulf(csub(a, b), w);
    C2 r;
    r.re = make_float2(a.x + b.x, d.x);
    r.im = make_float2(a.y + b.y, d.y);
    return r;
}

constexpr int kB4Slab = 16 * 33;      // pairs per warp slab

__global__ void __launch_bounds__(256, 2)
b4_fftyz_mul_kernel(const float4 *__restrict__ X1, float4 *__restrict__ X2, const float4 *__restrict__ Fb,
                    const float4 *__restrict__ F2b, const float2 *__restrict__ tw128_g, int rs, unsigned nmask,
                    int nsig, int nplanes) {
    constexpr int N = 128, H = 64;
    extern __shared__ float4 buf[];                                   // [8][kB4Slab]
    float2 *tw128 = reinterpret_cast<float2 *>(buf + 8 * kB4Slab);    // [128] W_128^k
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float4 *slab = buf + w * kB4Slab;
    const size_t slab128 = (size_t)N * H;                             // float4 per z of X1 / X2
    const int nzv = min(2 * rs + 1, N);
    const int npairs = nplanes / (3 * N * 2);
    const int b = blockIdx.x & 1;                                     // the grid is even: q keeps its class
    for (int i = threadIdx.x; i < N; i += 256) tw128[i] = tw128_g[i];
    __syncthreads();
    // S1 role: slab row ri = lane >> 1 (z_lo = 2w + (ri >> 3), n1 = ri & 7), c0 = lane & 1
    const int c0 = lane & 1, ri = lane >> 1;
    const int zr = 2 * w + (ri >> 3) + 16 * (ri & 7);
    const bool row_in = (zr + rs) % N < nzv;
    // S2 role: k'' = lane & 15, g = lane >> 4 (z_lo = 2w + g)
    const int kq = lane & 15, g = lane >> 4, zlo = 2 * w + g;
    const float2 tw32 = tw128[4 * kq], twA = tw128[2 * kq], twB = tw128[2 * kq + 32];

    for (int q = blockIdx.x; q < nplanes; q += gridDim.x) {
        const int qq = q >> 1;
        const int pair = qq % npairs, vol = (qq / npairs) % 3, kx = qq / (3 * npairs);
        const int sig = vol == 0 ? 0 : (vol == 1 ? 1 : nsig - 1);
        const float4 *src = X1 + ((size_t)(pair * nsig + sig) * N + zr) * slab128 + (size_t)kx * H + b * 32;
        float4 *dst = X2 + ((size_t)(pair * 3 + vol) * N + zr) * slab128 + (size_t)kx * H;
        const float4 *Fm = (vol == 2 ? F2b : Fb) + ((size_t)((kx * 2 + b) * 8 + w) * 16) * 32 + lane;   // + k0 * 32

        C2 v[16];
        // ---- S1
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int c = c0 + 2 * r;
            v[r] = (row_in && ((nmask >> (c >> 3)) & 1u)) ? ldg_c2(src + c) : c2_zero();
        }
        dft16(v);
        // ---- X1
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; ++k) sts_c2(slab + k * 33 + lane, v[k]);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = lds_c2(slab + kq * 33 + 16 * g + j);
        __syncwarp();
        // ---- S2
        {
            C2 lo[8], hi[8];
#pragma unroll
            for (int n1 = 0; n1 < 8; ++n1) {
                const C2 a = v[2 * n1], bb = cmulw(v[2 * n1 + 1], tw32);
                lo[n1] = b4_split(cadd(a, bb), twA);                  // k = k''
                hi[n1] = b4_split(csub(a, bb), twB);                  // k = k'' + 16
            }
            dft8(lo);
            dft8(hi);
            // ---- X2: slab[g 256 + k1 32 + k], times W_128^(z_lo k1)
#pragma unroll
            for (int k1 = 0; k1 < 8; ++k1) {
                const float2 t = tw128[(zlo * k1) & 127];
                sts_c2(slab + g * 256 + k1 * 32 + kq, k1 ? cmulw(lo[k1], t) : lo[k1]);
                sts_c2(slab + g * 256 + k1 * 32 + 16 + kq, k1 ? cmulw(hi[k1], t) : hi[k1]);
            }
        }
        __syncthreads();
        // ---- S3: k1 = w, column pair = lane; 16-point over z_lo, kz = k1 + 8 k0
        {
            C2 f[8];
#pragma unroll
            for (int k0 = 0; k0 < 8; ++k0) f[k0] = ldg_c2(Fm + k0 * 32);
            float4 *cell = buf + w * 32 + lane;
#pragma unroll
            for (int zl = 0; zl < 16; ++zl) v[zl] = lds_c2(cell + (zl >> 1) * kB4Slab + (zl & 1) * 256);
            dft16(v);
#pragma unroll
            for (int k0 = 0; k0 < 8; ++k0) v[k0] = cmul(v[k0], f[k0]);
#pragma unroll
            for (int k0 = 0; k0 < 8; ++k0) f[k0] = ldg_c2(Fm + (k0 + 8) * 32);
#pragma unroll
            for (int k0 = 0; k0 < 8; ++k0) v[k0 + 8] = cmul(v[k0 + 8], f[k0]);
            dft16(v);
#pragma unroll
            for (int zl = 0; zl < 16; ++zl) {
                const float2 t = tw128[(w * zl) & 127];
                sts_c2(cell + (zl >> 1) * kB4Slab + (zl & 1) * 256, zl ? cmulw(v[zl], t) : v[zl]);
            }
        }
        __syncthreads();
        // ---- S2'
        {
            C2 lo[8], hi[8];
#pragma unroll
            for (int k1 = 0; k1 < 8; ++k1) {
                lo[k1] = lds_c2(slab + g * 256 + k1 * 32 + kq);
                hi[k1] = lds_c2(slab + g * 256 + k1 * 32 + 16 + kq);
            }
            dft8(lo);
            dft8(hi);
#pragma unroll
            for (int n1 = 0; n1 < 8; ++n1) {
                const C2 u1 = b4_unsplit(lo[n1], twA), u2 = b4_unsplit(hi[n1], twB);
                v[2 * n1] = cadd(u1, u2);                             // m0 = 0
                v[2 * n1 + 1] = cmulw(csub(u1, u2), tw32);            // m0 = 1
            }
        }
        // ---- X4
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 16; ++j) sts_c2(slab + kq * 33 + 16 * g + j, v[j]);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = lds_c2(slab + k * 33 + lane);
        __syncwarp();
        // ---- S1': pair m = c0 + 2 m1 of the class, stored in kernel C's tile order [m / 8][b][m % 8]
        dft16(v);
#pragma unroll
        for (int m1 = 0; m1 < 16; ++m1) {
            const int m = c0 + 2 * m1;
            stg_c2(dst + ((m >> 3) * 2 + b) * 8 + (m & 7), v[m1]);
        }
    }
}

// Fb[kx][b][k1][k0][col] = (re F[kz][ky0][kx], re F[kz][ky1][kx], im .., im ..), ky0 = 2 col + b,
// ky1 = 2 (col + 32) + b, kz = k1 + 8 k0: the map spectrum in the order S3's threads consume it
__global__ void b4_spectrum_kernel(const float2 *__restrict__ F, float4 *__restrict__ Fb) {
    constexpr int N = 128;
    const size_t total = (size_t)N * 2 * 8 * 16 * 32;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int col = (int)(i % 32), k0 = (int)((i / 32) % 16), k1 = (int)((i / 512) % 8), b = (int)((i / 4096) % 2);
        const int kx = (int)(i / 8192);
        const int kz = k1 + 8 * k0;
        const float2 a = F[((size_t)kz * N + 2 * col + b) * N + kx], e = F[((size_t)kz * N + 2 * (col + 32) + b) * N + kx];
        Fb[i] = make_float4(a.x, e.x, a.y, e.y);
    }
}

static constexpr size_t kB4Smem = (size_t)8 * kB4Slab * sizeof(float4) + 128 * sizeof(float2);

int b4_init(Plan *p) {
    PFB_CUDA(cudaFuncSetAttribute(b4_fftyz_mul_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kB4Smem));
    return PFB_OK;
}

int b4_prepare_target(Plan *p, cudaStream_t s) {
    { LaunchScope ls(p, KC_OTHER, s);
      b4_spectrum_kernel<<<p->sm_count * 8, 256, 0, s>>>(p->F, reinterpret_cast<float4 *>(p->Fq)); }
    { LaunchScope ls(p, KC_OTHER, s);
      b4_spectrum_kernel<<<p->sm_count * 8, 256, 0, s>>>(p->F2, reinterpret_cast<float4 *>(p->F2q)); }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

int b4_launch(Plan *p, int count, float2 *X2, cudaStream_t s) {
    const int npairs = (count + 1) / 2;
    const int nplanes = 128 * 3 * npairs * 2;
    const int grid = std::min(nplanes, 2 * p->sm_count);
    LaunchScope ls(p, KC_FUSED_B, s);
    b4_fftyz_mul_kernel<<<grid, 256, kB4Smem, s>>>(
        reinterpret_cast<const float4 *>(p->A), reinterpret_cast<float4 *>(X2),
        reinterpret_cast<const float4 *>(p->Fq), reinterpret_cast<const float4 *>(p->F2q), p->tw[0], p->rs, p->nmask,
        p->nsig, nplanes);
    return PFB_OK;
}

}  // namespace pfb
