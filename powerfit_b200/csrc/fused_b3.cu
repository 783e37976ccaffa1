// Kernel B for 128^3 with the whole (y,z) plane resident in registers ("B3"); selected with
// PFB_B3=1, otherwise fused.cu's kernel B runs.  Same inputs (X1) and outputs (X2) as
// fused_fftyz_mul_kernel, so kernels A and C are unchanged.
//
// fused_fftyz_mul_kernel keeps 16 complex pairs per thread, which costs six shared-memory
// exchanges per plane (row, transpose, column, column, transpose, row = 12 accesses per element)
// and makes shared-memory bandwidth the limit of the whole search.  Here
// 256 threads hold 32 pairs each -- the entire 128 x 64 plane -- and four register stages each
// way need only four exchanges (8 accesses per element):
//
//   warp w owns the rows z = w + 8 n1 (n1 < 16), i.e. a 16 x 64 slab of 1024 pairs
//   S1  thread (n1, c0): 32-point transform over r of the row's pairs c = c0 + 2 r, straight from HBM
//   X1  exchange inside the warp's slab               -> thread k' holds (c0, n1)
//   S2  radix-2 over c0 (completes the packed 64-point row transform), the split step to
//       (Y[k], Y[k+64]), 16-point transforms over n1, twiddle W_128^(w k1)
//   X2  exchange through the CTA's plane buffer (each warp writes its own slab) + block barrier
//   S3  thread (w', l'): four 8-point transforms over the slabs (completes z), times the map
//       spectrum, four 8-point transforms back, twiddle -- written back IN PLACE
//   block barrier, then the mirror image: S2' (16-point over k1, unsplit, radix-2), X4, S1'
//   (32-point), store.
//
// Factorisation checked in FP64 against numpy (tools/b3_model.py).  All transforms use the
// kernel exp(+2 pi i n k / N), like the rest of the library.
#include "common.cuh"
#include "fft_core.cuh"

#include <algorithm>
#include <cstdlib>

namespace pfb {

// 32-point DFT in registers, natural order in and out: n = n0 + 4 n1, k = k1 + 8 k0
template <class T> __device__ __forceinline__ void dft32(T (&v)[32]) {
    T a[4][8];
#pragma unroll
    for (int n0 = 0; n0 < 4; ++n0) {
        T t[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) t[n1] = v[n0 + 4 * n1];
        dft8(t);
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) a[n0][k1] = t[k1];
    }
    // a[n0][k1] *= W32^(n0 k1)
    a[1][1] = rotc(a[1][1], 0.98078528040323043f, 0.19509032201612825f);   // W32^1
    a[1][2] = rotc(a[1][2], 0.92387953251128674f, 0.38268343236508978f);   // W32^2
    a[1][3] = rotc(a[1][3], 0.83146961230254524f, 0.55557023301960218f);   // W32^3
    a[1][4] = rot45(a[1][4]);
    a[1][5] = rotc(a[1][5], 0.55557023301960229f, 0.83146961230254524f);   // W32^5
    a[1][6] = rotc(a[1][6], 0.38268343236508984f, 0.92387953251128674f);   // W32^6
    a[1][7] = rotc(a[1][7], 0.19509032201612833f, 0.98078528040323043f);   // W32^7
    a[2][1] = rotc(a[2][1], 0.92387953251128674f, 0.38268343236508978f);   // W32^2
    a[2][2] = rot45(a[2][2]);
    a[2][3] = rotc(a[2][3], 0.38268343236508984f, 0.92387953251128674f);   // W32^6
    a[2][4] = mul_i(a[2][4]);
    a[2][5] = rotc(a[2][5], -0.38268343236508973f, 0.92387953251128674f);   // W32^10
    a[2][6] = rot135(a[2][6]);
    a[2][7] = rotc(a[2][7], -0.92387953251128674f, 0.38268343236508989f);   // W32^14
    a[3][1] = rotc(a[3][1], 0.83146961230254524f, 0.55557023301960218f);   // W32^3
    a[3][2] = rotc(a[3][2], 0.38268343236508984f, 0.92387953251128674f);   // W32^6
    a[3][3] = rotc(a[3][3], -0.19509032201612819f, 0.98078528040323043f);   // W32^9
    a[3][4] = rot135(a[3][4]);
    a[3][5] = rotc(a[3][5], -0.98078528040323043f, 0.19509032201612861f);   // W32^15
    a[3][6] = rotc(a[3][6], -0.92387953251128685f, -0.38268343236508967f);   // W32^18
    a[3][7] = rotc(a[3][7], -0.55557023301960218f, -0.83146961230254524f);   // W32^21
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) {
        dft4(a[0][k1], a[1][k1], a[2][k1], a[3][k1]);
#pragma unroll
        for (int k0 = 0; k0 < 4; ++k0) v[k1 + 8 * k0] = a[k0][k1];
    }
}

// (E, O) packed in one C2 -> (Y[k], Y[k + 64]) with w = W_128^k      (adjacent-in -> split-out)
__device__ __forceinline__ C2 b3_split(C2 v, float2 w) {
    const float2 o = cmulf(make_float2(v.re.y, v.im.y), w);
    const float er = v.re.x, ei = v.im.x;
    C2 r;
    r.re = make_float2(er + o.x, er - o.x);
    r.im = make_float2(ei + o.y, ei - o.y);
    return r;
}
// (Y[k], Y[k + 64]) -> (u[k], v[k]) = (a + b, (a - b) W_128^k)        (split-in -> adjacent-out)
__device__ __forceinline__ C2 b3_unsplit(C2 v, float2 w) {
    const float2 a = make_float2(v.re.x, v.im.x), b = make_float2(v.re.y, v.im.y);
    const float2 d = cmulf(csub(a, b), w);
    C2 r;
    r.re = make_float2(a.x + b.x, d.x);
    r.im = make_float2(a.y + b.y, d.y);
    return r;
}

constexpr int kB3Slab = 32 * 33;      // pairs per warp slab (row pitch 33: conflict-free both ways)

// STAGE: the support rows of the CTA's NEXT plane are copied into a shared-memory staging area with
// cp.async while the current plane is processed (every warp stages its own rows, so a warp-level
// wait suffices); S1 then reads shared memory instead of waiting for HBM.  Needs room for
// maxrows = max support rows per warp (8 maxrows KB next to the 133 KB plane buffer); larger
// templates run the direct-load variant.
template <bool STAGE>
__global__ void __launch_bounds__(256, 1)
b3_fftyz_mul_kernel(const float4 *__restrict__ X1, float4 *__restrict__ X2, const float4 *__restrict__ Fb,
                    const float4 *__restrict__ F2b, const float2 *__restrict__ tw128_g, int rs, unsigned ymask,
                    int nsig, int nplanes, int maxrows) {
    constexpr int N = 128, H = 64;
    extern __shared__ float4 buf[];                                   // [8][kB3Slab]
    float2 *tw128 = reinterpret_cast<float2 *>(buf + 8 * kB3Slab);    // [128] W_128^k
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float4 *slab = buf + w * kB3Slab;
    const size_t slab128 = (size_t)N * H;                             // float4 per z of X1 / X2
    const int nzv = min(2 * rs + 1, N);
    const int npairs = nplanes / (3 * N);
    for (int i = threadIdx.x; i < N; i += 256) tw128[i] = tw128_g[i];
    __syncthreads();
    // thread constants of the S2 role (k' = lane)
    const float2 tw64 = tw128[2 * lane], twA = tw128[lane], twB = tw128[lane + 32];
    // S1 role: row n1 = lane >> 1 of the slab, c0 = lane & 1
    const int n1r = lane >> 1, c0 = lane & 1;
    const int zr = w + 8 * n1r;
    const bool row_in = (zr + rs) % N < nzv;
    // staging: support rows of this warp (bit n1), this thread's row rank among them
    float4 *stage = reinterpret_cast<float4 *>(tw128 + 128) + w * maxrows * H;
    unsigned rowmask = 0;
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1)
        if ((w + 8 * n1 + rs) % N < nzv) rowmask |= 1u << n1;
    const int arank = __popc(rowmask & ((1u << n1r) - 1u));
    // row a keeps its 16-byte chunk c at c ^ (2 (a & 3)): four neighbouring rows read by one quarter warp
    // (chunks c0 + 2 r) then fall into different banks
    auto stage_plane = [&](int qn) {
        if (qn < nplanes) {
            const int pair = qn % npairs, vol = (qn / npairs) % 3, kx = qn / (3 * npairs);
            const int sig = vol == 0 ? 0 : (vol == 1 ? 1 : nsig - 1);
            const float4 *srcq = X1 + (size_t)(pair * nsig + sig) * N * slab128 + (size_t)kx * H;
            int a = 0;
            for (unsigned m = rowmask; m; m &= m - 1, ++a) {
                const int z = w + 8 * (__ffs(m) - 1);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int c = 32 * half + lane;
                    if ((ymask >> (c >> 4)) & 1u)
                        cp_async16(stage + a * H + (c ^ (2 * (a & 3))), srcq + (size_t)z * slab128 + c);
                }
            }
        }
        cp_async_commit();
    };
    if (STAGE) stage_plane(blockIdx.x);

    for (int q = blockIdx.x; q < nplanes; q += gridDim.x) {
        const int pair = q % npairs, vol = (q / npairs) % 3, kx = q / (3 * npairs);
        const int sig = vol == 0 ? 0 : (vol == 1 ? 1 : nsig - 1);
        const float4 *src = X1 + (size_t)(pair * nsig + sig) * N * slab128 + (size_t)kx * H + (size_t)zr * slab128;
        float4 *dst = X2 + (size_t)(pair * 3 + vol) * N * slab128 + (size_t)kx * H + (size_t)zr * slab128;
        const float4 *Fm = (vol == 2 ? F2b : Fb) + ((size_t)kx * 8 + w) * (4 * 8 * 32) + lane;   // + (p*8 + k0)*32

        C2 v[32];
        // ---- S1: the row's pairs c = c0 + 2 r from HBM, 32-point transform over r
        if (STAGE) {
            cp_async_wait<0>();
            __syncwarp();
            const float4 *srow = stage + arank * H;
            const int sw = 2 * (arank & 3);
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const int c = c0 + 2 * r;
                v[r] = (row_in && ((ymask >> (c >> 4)) & 1u)) ? lds_c2(srow + (c ^ sw)) : c2_zero();
            }
            __syncwarp();
            stage_plane(q + gridDim.x);
        } else {
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const int c = c0 + 2 * r;
                v[r] = (row_in && ((ymask >> (c >> 4)) & 1u)) ? ldg_c2(src + c) : c2_zero();
            }
        }
        dft32(v);
        // ---- X1: v[k'] of thread j = 2 n1 + c0  ->  thread k' holds u[j]
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 32; ++k) sts_c2(slab + k * 33 + lane, v[k]);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = lds_c2(slab + lane * 33 + j);
        __syncwarp();
        // ---- S2: radix-2 over c0, split step, 16-point transforms over n1, twiddle
        {
            C2 lo[16], hi[16];
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) {
                const C2 a = v[2 * n1], b = cmulw(v[2 * n1 + 1], tw64);
                lo[n1] = b3_split(cadd(a, b), twA);                   // k = k'
                hi[n1] = b3_split(csub(a, b), twB);                   // k = k' + 32
            }
            dft16(lo);
            dft16(hi);
            // ---- X2: slab[(kk 16 + k1) 32 + k'], times W_128^(w k1)
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) {
                const float2 t = tw128[(w * k1) & 127];
                sts_c2(slab + k1 * 32 + lane, k1 ? cmulw(lo[k1], t) : lo[k1]);
                sts_c2(slab + (16 + k1) * 32 + lane, k1 ? cmulw(hi[k1], t) : hi[k1]);
            }
        }
        __syncthreads();
        // ---- S3: problems p = kk 2 + dk: column pair k' + 32 kk, k1 = 2 w + dk; 8-point over the slabs
        {
            C2 f[8];
#pragma unroll
            for (int k0 = 0; k0 < 8; ++k0) f[k0] = ldg_c2(Fm + k0 * 32);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int k1 = 2 * w + (p & 1);
                float4 *cell = buf + ((p >> 1) * 16 + k1) * 32 + lane;          // + z_lo * kB3Slab
                C2 g[8], fn[8];
#pragma unroll
                for (int zl = 0; zl < 8; ++zl) g[zl] = lds_c2(cell + zl * kB3Slab);
                if (p < 3) {
#pragma unroll
                    for (int k0 = 0; k0 < 8; ++k0) fn[k0] = ldg_c2(Fm + ((p + 1) * 8 + k0) * 32);
                }
                dft8(g);                                              // kz = k1 + 16 k0
#pragma unroll
                for (int k0 = 0; k0 < 8; ++k0) g[k0] = cmul(g[k0], f[k0]);
                dft8(g);                                              // back: index z_lo
#pragma unroll
                for (int zl = 0; zl < 8; ++zl) {
                    const float2 t = tw128[(k1 * zl) & 127];
                    sts_c2(cell + zl * kB3Slab, zl ? cmulw(g[zl], t) : g[zl]);
                }
                if (p < 3) {
#pragma unroll
                    for (int k0 = 0; k0 < 8; ++k0) f[k0] = fn[k0];
                }
            }
        }
        __syncthreads();
        // ---- S2': 16-point transforms over k1, unsplit, radix-2 over kk
        {
            C2 lo[16], hi[16];
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) {
                lo[k1] = lds_c2(slab + k1 * 32 + lane);
                hi[k1] = lds_c2(slab + (16 + k1) * 32 + lane);
            }
            dft16(lo);
            dft16(hi);
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) {
                const C2 u1 = b3_unsplit(lo[n1], twA), u2 = b3_unsplit(hi[n1], twB);
                v[2 * n1] = cadd(u1, u2);                             // m0 = 0
                v[2 * n1 + 1] = cmulw(csub(u1, u2), tw64);            // m0 = 1
            }
        }
        // ---- X4: thread k' holds t[j], j = 2 n1 + m0  ->  thread j holds v[k']
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) sts_c2(slab + lane * 33 + j, v[j]);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = lds_c2(slab + k * 33 + lane);
        __syncwarp();
        // ---- S1': 32-point transform over k' -> m1; y pair m = m0 + 2 m1 of row z
        dft32(v);
#pragma unroll
        for (int m1 = 0; m1 < 32; ++m1) stg_c2(dst + c0 + 2 * m1, v[m1]);
    }
}

// Fb[kx][w][p][k0][l] = (re F[kz][ky][kx], re F[kz][ky+64][kx], im .., im ..), ky = l + 32 (p >> 1),
// kz = 2 w + (p & 1) + 16 k0: the map spectrum in the order S3's threads consume it
__global__ void b3_spectrum_kernel(const float2 *__restrict__ F, float4 *__restrict__ Fb) {
    constexpr int N = 128;
    const size_t total = (size_t)N * 8 * 4 * 8 * 32;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int l = (int)(i % 32), k0 = (int)((i / 32) % 8), p = (int)((i / 256) % 4), w = (int)((i / 1024) % 8);
        const int kx = (int)(i / 8192);
        const int ky = l + 32 * (p >> 1), kz = 2 * w + (p & 1) + 16 * k0;
        const float2 a = F[((size_t)kz * N + ky) * N + kx], e = F[((size_t)kz * N + ky + 64) * N + kx];
        Fb[i] = make_float4(a.x, e.x, a.y, e.y);
    }
}

static constexpr size_t kB3Smem = (size_t)8 * kB3Slab * sizeof(float4) + 128 * sizeof(float2);
static constexpr size_t kSmemMax = 232448;       // 227 KB per CTA

int b3_init(Plan *p) {
    PFB_CUDA(cudaFuncSetAttribute(b3_fftyz_mul_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kB3Smem));
    PFB_CUDA(cudaFuncSetAttribute(b3_fftyz_mul_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
    return PFB_OK;
}

int b3_prepare_target(Plan *p, cudaStream_t s) {
    { LaunchScope ls(p, KC_OTHER, s);
      b3_spectrum_kernel<<<p->sm_count * 8, 256, 0, s>>>(p->F, reinterpret_cast<float4 *>(p->Fq)); }
    { LaunchScope ls(p, KC_OTHER, s);
      b3_spectrum_kernel<<<p->sm_count * 8, 256, 0, s>>>(p->F2, reinterpret_cast<float4 *>(p->F2q)); }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

int b3_launch(Plan *p, int count, float2 *X2, cudaStream_t s) {
    const int npairs = (count + 1) / 2;
    const int nplanes = 128 * 3 * npairs;
    // support rows per warp (rows z = w + 8 n1): at most ceil(nzv / 8) + 1
    const int nzv = std::min(2 * p->rs + 1, 128);
    const int maxrows = std::min(16, (nzv + 7) / 8 + 1);
    const size_t smem_stage = kB3Smem + (size_t)8 * maxrows * 64 * sizeof(float4);
    static const int stage_env = getenv("PFB_B3_STAGE") ? atoi(getenv("PFB_B3_STAGE")) : 1;
    LaunchScope ls(p, KC_FUSED_B, s);
    if (stage_env && smem_stage <= kSmemMax)
        b3_fftyz_mul_kernel<true><<<std::min(nplanes, p->sm_count), 256, smem_stage, s>>>(
            reinterpret_cast<const float4 *>(p->A), reinterpret_cast<float4 *>(X2),
            reinterpret_cast<const float4 *>(p->Fq), reinterpret_cast<const float4 *>(p->F2q), p->tw[0], p->rs,
            p->ymask, p->nsig, nplanes, maxrows);
    else
        b3_fftyz_mul_kernel<false><<<std::min(nplanes, p->sm_count), 256, kB3Smem, s>>>(
            reinterpret_cast<const float4 *>(p->A), reinterpret_cast<float4 *>(X2),
            reinterpret_cast<const float4 *>(p->Fq), reinterpret_cast<const float4 *>(p->F2q), p->tw[0], p->rs,
            p->ymask, p->nsig, nplanes, maxrows);
    return PFB_OK;
}

}  // namespace pfb
