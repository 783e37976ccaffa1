// General mixed-radix complex FFT along one axis of a batch of 3-D volumes.
//
// This is the any-shape path (axis lengths 2^a 3^b 5^c 7^d, the set the reference's
// clFFT backend accepts; /root/reference/src/powerfit_em/powerfit.py:230-233 pads to it).
// It replaces grfftn_builder / clFFT (powerfitter.py:605-638).  One transform direction
// is enough for the whole search: with P(z)(k) = sum_r z(r) exp(+2 pi i k r / n),
//     corr(z, f) = P( P(z) . conj(P f) / V )        (f real),
// so forward and inverse passes are the same kernel ("one direction").
//
// Kernel shape: a CTA owns T neighbouring lines of the axis, stages them in shared
// memory as split re/im planes laid out [element][line] (+1 padding), runs Stockham
// autosort passes between two buffers, and writes the lines back.  For the x axis the T
// lines are contiguous in memory; for y and z the T lines are T neighbouring x
// positions, so every global access is a run of T*8 contiguous bytes.
#include "common.cuh"

#include <cmath>

namespace pfb {

bool factorize(int n, Fft1D *out) {
    Fft1D f;
    f.n = n;
    int m = n;
    if (m < 1) return false;
    auto take = [&](int r) {
        while (m % r == 0 && f.npass < kMaxPasses) { f.radix[f.npass++] = r; m /= r; }
    };
    take(8); take(4); take(2); take(3); take(5); take(7);
    if (m != 1) return false;
    *out = f;
    return true;
}

struct AxisGeom {
    int n;              // transform length
    long stride;        // element stride along the axis
    long inner;         // number of lines that are contiguous neighbours
    long outer_stride;  // distance between groups of `inner` lines
    long nlines;        // lines per volume
    long vol_stride;    // elements per volume
    int T;              // lines per CTA
};

template <int R>
__device__ __forceinline__ void small_dft(float *re, float *im, const float2 *__restrict__ tw, int tw_step) {
    // X[k] = sum_i x[i] W^(i k),  W = exp(+2 pi i / R) = tw[tw_step]
    if (R == 2) {
        const float ar = re[0], ai = im[0];
        re[0] = ar + re[1]; im[0] = ai + im[1];
        re[1] = ar - re[1]; im[1] = ai - im[1];
    } else if (R == 4) {
        const float s0r = re[0] + re[2], s0i = im[0] + im[2];
        const float d0r = re[0] - re[2], d0i = im[0] - im[2];
        const float s1r = re[1] + re[3], s1i = im[1] + im[3];
        const float d1r = re[1] - re[3], d1i = im[1] - im[3];
        re[0] = s0r + s1r; im[0] = s0i + s1i;
        re[2] = s0r - s1r; im[2] = s0i - s1i;
        re[1] = d0r - d1i; im[1] = d0i + d1r;   // d0 + i d1
        re[3] = d0r + d1i; im[3] = d0i - d1r;   // d0 - i d1
    } else if (R == 8) {
        // two radix-4 on even/odd, then combine with W8^k
        float er[4] = {re[0], re[2], re[4], re[6]}, ei[4] = {im[0], im[2], im[4], im[6]};
        float qr[4] = {re[1], re[3], re[5], re[7]}, qi[4] = {im[1], im[3], im[5], im[7]};
        small_dft<4>(er, ei, tw, 0);
        small_dft<4>(qr, qi, tw, 0);
        const float h = 0.70710678118654752440f;
        // W8^1 = h(1+i), W8^2 = i, W8^3 = h(-1+i)
        float t1r = h * (qr[1] - qi[1]), t1i = h * (qr[1] + qi[1]);
        float t2r = -qi[2], t2i = qr[2];
        float t3r = h * (-qr[3] - qi[3]), t3i = h * (qr[3] - qi[3]);
        re[0] = er[0] + qr[0]; im[0] = ei[0] + qi[0];
        re[4] = er[0] - qr[0]; im[4] = ei[0] - qi[0];
        re[1] = er[1] + t1r;   im[1] = ei[1] + t1i;
        re[5] = er[1] - t1r;   im[5] = ei[1] - t1i;
        re[2] = er[2] + t2r;   im[2] = ei[2] + t2i;
        re[6] = er[2] - t2r;   im[6] = ei[2] - t2i;
        re[3] = er[3] + t3r;   im[3] = ei[3] + t3i;
        re[7] = er[3] - t3r;   im[7] = ei[3] - t3i;
    } else {
        float xr[R], xi[R];
#pragma unroll
        for (int i = 0; i < R; ++i) { xr[i] = re[i]; xi[i] = im[i]; }
#pragma unroll
        for (int k = 0; k < R; ++k) {
            float sr = xr[0], si = xi[0];
#pragma unroll
            for (int i = 1; i < R; ++i) {
                const float2 w = __ldg(tw + ((i * k) % R) * tw_step);
                sr += xr[i] * w.x - xi[i] * w.y;
                si += xr[i] * w.y + xi[i] * w.x;
            }
            re[k] = sr; im[k] = si;
        }
    }
}

template <int R>
__device__ __forceinline__ void stockham_pass(const float *__restrict__ sre, const float *__restrict__ sim,
                                              float *__restrict__ dre, float *__restrict__ dim_, int n, int Ns,
                                              int T, int pitch, const float2 *__restrict__ tw) {
    const int nb = n / R;                      // butterflies per line
    const int tw_mul = n / (Ns * R);           // W_{Ns R}^m = tw[m * tw_mul]
    for (int w = threadIdx.x; w < nb * T; w += blockDim.x) {
        const int t = w % T, j = w / T;
        const int k = j % Ns;
        float re[R], im[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int src = (j + i * nb) * pitch + t;
            float xr = sre[src], xi = sim[src];
            if (i > 0 && Ns > 1) {
                const float2 c = __ldg(tw + (long)(i * k) * tw_mul);
                const float yr = xr * c.x - xi * c.y;
                xi = xr * c.y + xi * c.x;
                xr = yr;
            }
            re[i] = xr; im[i] = xi;
        }
        small_dft<R>(re, im, tw, n / R);
        const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int dst = (j0 + i * Ns) * pitch + t;
            dre[dst] = re[i]; dim_[dst] = im[i];
        }
    }
}

__global__ void __launch_bounds__(256)
fft_axis_kernel(float2 *__restrict__ data, AxisGeom g, Fft1D plan, const float2 *__restrict__ tw) {
    extern __shared__ float smem[];
    const int T = g.T, pitch = T + 1, n = g.n;
    float *b0r = smem, *b0i = b0r + n * pitch, *b1r = b0i + n * pitch, *b1i = b1r + n * pitch;
    const long l0 = (long)blockIdx.x * T;
    float2 *vol = data + (long)blockIdx.y * g.vol_stride;
    const int nl = (int)min((long)T, g.nlines - l0);

    // ---- load
    if (g.inner == 1) {             // lines contiguous in memory: element index fastest
        for (int w = threadIdx.x; w < nl * n; w += blockDim.x) {
            const int t = w / n, j = w % n;
            const float2 v = vol[(l0 + t) * g.outer_stride + j];
            b0r[j * pitch + t] = v.x; b0i[j * pitch + t] = v.y;
        }
    } else {
        for (int w = threadIdx.x; w < T * n; w += blockDim.x) {
            const int t = w % T, j = w / T;
            if (t < nl) {
                const long l = l0 + t;
                const float2 v = vol[(l / g.inner) * g.outer_stride + (l % g.inner) + (long)j * g.stride];
                b0r[j * pitch + t] = v.x; b0i[j * pitch + t] = v.y;
            }
        }
    }
    __syncthreads();

    // ---- Stockham passes
    float *sr = b0r, *si = b0i, *dr = b1r, *di = b1i;
    int Ns = 1;
    for (int p = 0; p < plan.npass; ++p) {
        const int R = plan.radix[p];
        switch (R) {
            case 2: stockham_pass<2>(sr, si, dr, di, n, Ns, T, pitch, tw); break;
            case 3: stockham_pass<3>(sr, si, dr, di, n, Ns, T, pitch, tw); break;
            case 4: stockham_pass<4>(sr, si, dr, di, n, Ns, T, pitch, tw); break;
            case 5: stockham_pass<5>(sr, si, dr, di, n, Ns, T, pitch, tw); break;
            case 7: stockham_pass<7>(sr, si, dr, di, n, Ns, T, pitch, tw); break;
            default: stockham_pass<8>(sr, si, dr, di, n, Ns, T, pitch, tw); break;
        }
        Ns *= R;
        __syncthreads();
        float *tr = sr, *ti = si; sr = dr; si = di; dr = tr; di = ti;
    }

    // ---- store
    if (g.inner == 1) {
        for (int w = threadIdx.x; w < nl * n; w += blockDim.x) {
            const int t = w / n, j = w % n;
            vol[(l0 + t) * g.outer_stride + j] = make_float2(sr[j * pitch + t], si[j * pitch + t]);
        }
    } else {
        for (int w = threadIdx.x; w < T * n; w += blockDim.x) {
            const int t = w % T, j = w / T;
            if (t < nl) {
                const long l = l0 + t;
                vol[(l / g.inner) * g.outer_stride + (l % g.inner) + (long)j * g.stride] =
                    make_float2(sr[j * pitch + t], si[j * pitch + t]);
            }
        }
    }
}

int fft_generic_init() {
    PFB_CUDA(cudaFuncSetAttribute(fft_axis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return PFB_OK;
}

// axis: 0 = x (fastest), 1 = y, 2 = z
int launch_fft_axis(Plan *p, float2 *vols, int nvol, int axis, cudaStream_t s) {
    AxisGeom g;
    const Fft1D *f;
    if (axis == 0) {
        f = &p->fx; g.n = p->nx; g.stride = 1; g.inner = 1; g.outer_stride = p->nx;
        g.nlines = (long)p->nz * p->ny;
    } else if (axis == 1) {
        f = &p->fy; g.n = p->ny; g.stride = p->nx; g.inner = p->nx; g.outer_stride = (long)p->nx * p->ny;
        g.nlines = (long)p->nz * p->nx;
    } else {
        f = &p->fz; g.n = p->nz; g.stride = (long)p->nx * p->ny; g.inner = (long)p->nx * p->ny;
        g.outer_stride = 0; g.nlines = (long)p->nx * p->ny;
    }
    if (g.n == 1) return PFB_OK;
    g.vol_stride = p->V;
    // lines per CTA: as many as fit ~64 KB of staging, even, at most 32
    int T = (int)(65536 / (16L * g.n)) - 1;
    if (T > 32) T = 32;
    if (T < 2) T = 2;
    T &= ~1;
    g.T = T;
    const size_t smem = (size_t)4 * g.n * (T + 1) * sizeof(float);
    if (smem > 200 * 1024) {
        set_error("axis length too large for the shared-memory FFT");
        return PFB_ERR_UNSUPPORTED;
    }
    dim3 grid((unsigned)((g.nlines + T - 1) / T), nvol);
    { LaunchScope ls(p, KC_FFT_X + axis, s);
      fft_axis_kernel<<<grid, 256, smem, s>>>(vols, g, *f, p->tw[axis]); }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

}  // namespace pfb
