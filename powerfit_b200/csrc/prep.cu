// One-time input preparation on the device, in FP64 like the reference's host code
// (the steps immediately before the search path):
//
//   target   BaseCorrelator.__init__ / _get_lcc_mask (powerfitter.py:169-180):
//            f = target / target.max(); lcc_mask = f > 0.05 * f.max();
//            GPUCorrelator.__init__ (:410-414): [Laplace], cast to float32
//   template BaseCorrelator.mask.fset / _laplace_filter / _normalize_template (:190-220):
//            N = count(mask != 0); [Laplace]; t *= m; t[ind] -= mean; t[ind] /= std; t *= m;
//            GPUCorrelator.mask.fset (:466-474): cast to float32
//
// scipy.ndimage.laplace(mode='wrap') is reproduced operation for operation: per axis
// d2 = fl(fl(-2 x[i]) + fl(x[i-1] + x[i+1])) (the symmetric branch of NI_Correlate1D), summed
// as (d2_z + d2_y) + d2_x (generic_laplace's axis loop) -- bit-identical in FP64, verified
// against scipy in tests/test_host_logic.py.  Division and max are exact IEEE operations, so
// the float32 target and the lcc_mask equal the host path's bit for bit.  The masked mean and
// standard deviation are FP64 sums in a fixed two-level order; they differ from numpy's
// pairwise sums in the last bits of the FP64 result, which the float32 cast absorbs except
// for isolated one-ulp differences.
#include "common.cuh"

namespace pfb {

__device__ __forceinline__ long long orderable_f64(double x) {
    const long long s = __double_as_longlong(x);
    return s ^ ((s >> 63) & 0x7FFFFFFFFFFFFFFFll);
}
__host__ __device__ inline double unorderable_f64(long long k) {
    const long long s = k ^ ((k >> 63) & 0x7FFFFFFFFFFFFFFFll);
#ifdef __CUDA_ARCH__
    return __longlong_as_double(s);
#else
    double d;
    memcpy(&d, &s, sizeof(d));
    return d;
#endif
}

// max of a (optionally divided by *div) -> atomicMax on the order-preserving integer image
__global__ void max_f64_kernel(const double *__restrict__ a, long V, const long long *div_key, long long *out) {
    const double d = div_key ? unorderable_f64(*div_key) : 1.0;
    long long best = (long long)0x8000000000000000ull;
    for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long)gridDim.x * blockDim.x) {
        const double x = div_key ? __ddiv_rn(a[v], d) : a[v];
        best = max(best, orderable_f64(x));
    }
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, best);
}

__device__ __forceinline__ double d2_wrap(double c, double lo, double hi) {
    return __dadd_rn(__dmul_rn(c, -2.0), __dadd_rn(lo, hi));
}

// scipy.ndimage.laplace(x, mode='wrap') at voxel (z,y,x); `scale` divides the input first
template <bool DIV>
__device__ __forceinline__ double laplace_at(const double *__restrict__ a, int z, int y, int x, int nz, int ny, int nx,
                                             double d) {
    auto at = [&](int zz, int yy, int xx) {
        const double v = a[((long)zz * ny + yy) * nx + xx];
        return DIV ? __ddiv_rn(v, d) : v;
    };
    const double c = at(z, y, x);
    const double dz = d2_wrap(c, at(z == 0 ? nz - 1 : z - 1, y, x), at(z + 1 == nz ? 0 : z + 1, y, x));
    const double dy = d2_wrap(c, at(z, y == 0 ? ny - 1 : y - 1, x), at(z, y + 1 == ny ? 0 : y + 1, x));
    const double dx = d2_wrap(c, at(z, y, x == 0 ? nx - 1 : x - 1), at(z, y, x + 1 == nx ? 0 : x + 1));
    return __dadd_rn(__dadd_rn(dz, dy), dx);
}

// f32 target (normalised, optionally Laplace-filtered) and lcc_mask
__global__ void target_prep_kernel(const double *__restrict__ a, int nz, int ny, int nx, const long long *max_key,
                                   const long long *fmax_key, int laplace, float *__restrict__ f,
                                   uint8_t *__restrict__ lcc_mask) {
    const long V = (long)nz * ny * nx;
    const double d = unorderable_f64(*max_key);
    const double thr = __dmul_rn(unorderable_f64(*fmax_key), 0.05);
    for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long)gridDim.x * blockDim.x) {
        const int x = (int)(v % nx), y = (int)((v / nx) % ny), z = (int)(v / ((long)nx * ny));
        const double t = __ddiv_rn(a[v], d);
        lcc_mask[v] = t > thr ? 1 : 0;
        f[v] = (float)(laplace ? laplace_at<true>(a, z, y, x, nz, ny, nx, d) : t);
    }
}

// ---- masked template statistics: fixed two-level FP64 sums
struct Stat { double s; long long n; long long nb; };

// pass 1: w = [laplace](t) * m -> work; partial sums of w over ind, count of ind, count of mask values != 0,1
__global__ void __launch_bounds__(256)
tmpl_pass1_kernel(const double *__restrict__ t, const double *__restrict__ m, int nz, int ny, int nx, int laplace,
                  double *__restrict__ work, Stat *__restrict__ part) {
    const long V = (long)nz * ny * nx;
    double s = 0.0;
    long long n = 0, nb = 0;
    for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long)gridDim.x * blockDim.x) {
        const int x = (int)(v % nx), y = (int)((v / nx) % ny), z = (int)(v / ((long)nx * ny));
        const double mv = m[v];
        const double w = __dmul_rn(laplace ? laplace_at<false>(t, z, y, x, nz, ny, nx, 1.0) : t[v], mv);
        work[v] = w;
        if (mv != 0.0) { s = __dadd_rn(s, w); ++n; if (mv != 1.0) ++nb; }
    }
    __shared__ double ss[256];
    __shared__ long long sn[256], sb[256];
    ss[threadIdx.x] = s; sn[threadIdx.x] = n; sb[threadIdx.x] = nb;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            ss[threadIdx.x] = __dadd_rn(ss[threadIdx.x], ss[threadIdx.x + o]);
            sn[threadIdx.x] += sn[threadIdx.x + o];
            sb[threadIdx.x] += sb[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part[blockIdx.x].s = ss[0]; part[blockIdx.x].n = sn[0]; part[blockIdx.x].nb = sb[0]; }
}

// single block: total of the per-block partials -> tot[slot]
__global__ void __launch_bounds__(256) stat_final_kernel(const Stat *__restrict__ part, int nparts, Stat *tot, int slot) {
    __shared__ double ss[256];
    __shared__ long long sn[256], sb[256];
    double s = 0.0;
    long long n = 0, nb = 0;
    for (int i = threadIdx.x; i < nparts; i += 256) { s = __dadd_rn(s, part[i].s); n += part[i].n; nb += part[i].nb; }
    ss[threadIdx.x] = s; sn[threadIdx.x] = n; sb[threadIdx.x] = nb;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            ss[threadIdx.x] = __dadd_rn(ss[threadIdx.x], ss[threadIdx.x + o]);
            sn[threadIdx.x] += sn[threadIdx.x + o];
            sb[threadIdx.x] += sb[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { tot[slot].s = ss[0]; tot[slot].n = sn[0]; tot[slot].nb = sb[0]; }
}

// pass 2/3: MODE 0: w[ind] -= mean(tot[0]); partial sums of the centred values
//           MODE 1: partial sums of |w - mean(tot[1])|^2 over ind
template <int MODE>
__global__ void __launch_bounds__(256)
tmpl_pass23_kernel(const double *__restrict__ m, long V, double *__restrict__ work, const Stat *__restrict__ tot,
                   Stat *__restrict__ part) {
    const double mean = __ddiv_rn(tot[MODE].s, (double)tot[0].n);
    double s = 0.0;
    for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long)gridDim.x * blockDim.x) {
        if (m[v] != 0.0) {
            if (MODE == 0) {
                const double w = __dsub_rn(work[v], mean);
                work[v] = w;
                s = __dadd_rn(s, w);
            } else {
                const double dlt = __dsub_rn(work[v], mean);
                s = __dadd_rn(s, __dmul_rn(dlt, dlt));
            }
        }
    }
    __shared__ double ss[256];
    ss[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) ss[threadIdx.x] = __dadd_rn(ss[threadIdx.x], ss[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) { part[blockIdx.x].s = ss[0]; part[blockIdx.x].n = 0; part[blockIdx.x].nb = 0; }
}

// pass 4: t = (w / std) * m on ind, w * m elsewhere (= 0 * 0); cast both to float32
__global__ void tmpl_pass4_kernel(const double *__restrict__ m, long V, const double *__restrict__ work,
                                  const Stat *__restrict__ tot, float *__restrict__ t32, float *__restrict__ m32) {
    const double sd = sqrt(__ddiv_rn(tot[2].s, (double)tot[0].n));
    for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long)gridDim.x * blockDim.x) {
        const double mv = m[v];
        double w = work[v];
        if (mv != 0.0) w = __ddiv_rn(w, sd);
        t32[v] = (float)__dmul_rn(w, mv);
        m32[v] = (float)mv;
    }
}

static constexpr int kPrepBlocks = 592;      // 4 per SM; the sums' order depends on it, so it is fixed

int prep_target(Plan *p, const double *target, int laplace, float *f_out, uint8_t *lcc_mask_out, cudaStream_t s) {
    long long *keys = reinterpret_cast<long long *>(p->best_scratch);        // [0] max(target), [1] max(target / max)
    const long long lowest = (long long)0x8000000000000000ull;
    const long long init[2] = {lowest, lowest};
    PFB_CUDA(cudaMemcpyAsync(keys, init, sizeof(init), cudaMemcpyHostToDevice, s));
    { LaunchScope ls(p, KC_OTHER, s);
      max_f64_kernel<<<kPrepBlocks, 256, 0, s>>>(target, p->V, nullptr, keys); }
    { LaunchScope ls(p, KC_OTHER, s);
      max_f64_kernel<<<kPrepBlocks, 256, 0, s>>>(target, p->V, keys, keys + 1); }
    { LaunchScope ls(p, KC_OTHER, s);
      target_prep_kernel<<<kPrepBlocks, 256, 0, s>>>(target, p->nz, p->ny, p->nx, keys, keys + 1, laplace, f_out,
                                                     lcc_mask_out); }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

int prep_template(Plan *p, const double *tmpl, const double *mask, int laplace, float *t_out, float *m_out,
                  double *norm_factor, int *mask_is_binary, cudaStream_t s) {
    double *work = reinterpret_cast<double *>(p->A);                          // V doubles of the forward work buffer
    static_assert(sizeof(Stat) * (kPrepBlocks + 3) <= kPrepScratchBytes, "prep scratch too small");
    Stat *part = reinterpret_cast<Stat *>(p->prep_scratch);                   // kPrepBlocks partials + 3 totals
    Stat *tot = part + kPrepBlocks;
    { LaunchScope ls(p, KC_OTHER, s);
      tmpl_pass1_kernel<<<kPrepBlocks, 256, 0, s>>>(tmpl, mask, p->nz, p->ny, p->nx, laplace, work, part); }
    { LaunchScope ls(p, KC_OTHER, s); stat_final_kernel<<<1, 256, 0, s>>>(part, kPrepBlocks, tot, 0); }
    { LaunchScope ls(p, KC_OTHER, s);
      tmpl_pass23_kernel<0><<<kPrepBlocks, 256, 0, s>>>(mask, p->V, work, tot, part); }
    { LaunchScope ls(p, KC_OTHER, s); stat_final_kernel<<<1, 256, 0, s>>>(part, kPrepBlocks, tot, 1); }
    { LaunchScope ls(p, KC_OTHER, s);
      tmpl_pass23_kernel<1><<<kPrepBlocks, 256, 0, s>>>(mask, p->V, work, tot, part); }
    { LaunchScope ls(p, KC_OTHER, s); stat_final_kernel<<<1, 256, 0, s>>>(part, kPrepBlocks, tot, 2); }
    Stat h0;
    PFB_CUDA(cudaMemcpyAsync(&h0, tot, sizeof(Stat), cudaMemcpyDeviceToHost, s));
    PFB_CUDA(cudaStreamSynchronize(s));
    *norm_factor = (double)h0.n;
    *mask_is_binary = h0.nb == 0 ? 1 : 0;
    if (h0.n == 0) {
        set_error("Zero-filled mask is not allowed.");
        return PFB_ERR_INVALID;
    }
    { LaunchScope ls(p, KC_OTHER, s);
      tmpl_pass4_kernel<<<kPrepBlocks, 256, 0, s>>>(mask, p->V, work, tot, t_out, m_out); }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

}  // namespace pfb
