// Point-wise stages of the search: spectrum products, the LCC normalisation and the
// running arg-max, plus packing helpers.
//
// Reference operators replaced: conj_multiply (_powerfit.pyx:45-53 / powerfitter.py:568-571),
// calc_lcc (_powerfit.pyx:56-72) + `ave2 *= norm_factor` (powerfitter.py:369), take-best
// (powerfitter.py:327-330) and their fused OpenCL form calc_lcc_and_take_best
// (powerfitter.py:572-584).
#include "common.cuh"

namespace pfb {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// B[pair][0] = A[pair][0] . F ; B[pair][1] = A[pair][1] . F ; B[pair][2] = A[pair][nsig-1] . F2
__global__ void __launch_bounds__(256)
multiply_kernel(const float2 *__restrict__ A, float2 *__restrict__ B, const float2 *__restrict__ F,
                const float2 *__restrict__ F2, long V, int nsig) {
    const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const float2 f = F[v], f2 = F2[v];
    const float2 *a = A + (long)blockIdx.y * nsig * V;
    float2 *b = B + (long)blockIdx.y * 3 * V;
    const float2 zt = a[v], zm = a[V + v];
    const float2 zm2 = nsig == 3 ? a[2 * V + v] : zm;
    b[v] = cmul(zt, f);
    b[V + v] = cmul(zm, f);
    b[2 * V + v] = cmul(zm2, f2);
}

// The LCC of one candidate and the fold into the packed best.  IEEE sqrt and division
// (no fast-math): var <= 0 gives NaN or inf exactly as in the reference; NaN never wins.
__device__ __forceinline__ void fold(int64_t &best, float gcc, float ave, float ave2, float norm, uint32_t rot) {
    const float var = __fsub_rn(__fmul_rn(ave2, norm), __fmul_rn(ave, ave));
    const float lcc = __fdiv_rn(gcc, __fsqrt_rn(var));
    if (lcc == lcc) {
        const int64_t key = pack_best(__float_as_uint(lcc), rot);
        if (key > best) best = key;
    }
}

// One thread per voxel, looping over the rotation pairs of the batch in index order.
__global__ void __launch_bounds__(256)
lcc_best_kernel(const float2 *__restrict__ B, const uint8_t *__restrict__ lcc_mask, float norm, int first_index,
                int count, int64_t *__restrict__ best, long V) {
    const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V || lcc_mask[v] == 0) return;
    int64_t b = best[v];
    const int npairs = (count + 1) / 2;
    for (int p = 0; p < npairs; ++p) {
        const float2 *base = B + (long)p * 3 * V;
        const float2 gcc = base[v], ave = base[V + v], ave2 = base[2 * V + v];
        fold(b, gcc.x, ave.x, ave2.x, norm, (uint32_t)(first_index + 2 * p));
        if (2 * p + 1 < count) fold(b, gcc.y, ave.y, ave2.y, norm, (uint32_t)(first_index + 2 * p + 1));
    }
    best[v] = b;
}

__global__ void __launch_bounds__(256)
lcc_single_kernel(const float *__restrict__ gcc, const float *__restrict__ ave, const float *__restrict__ ave2,
                  const uint8_t *__restrict__ lcc_mask, float norm, int rot_index, int64_t *__restrict__ best, long V) {
    const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V || lcc_mask[v] == 0) return;
    int64_t b = best[v];
    fold(b, gcc[v], ave[v], ave2[v], norm, (uint32_t)rot_index);
    best[v] = b;
}

__global__ void best_init_kernel(int64_t *best, long V) {
    const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < V) best[v] = kBestInit;
}

__global__ void unpack_kernel(const int64_t *__restrict__ best, float *__restrict__ lcc, int32_t *__restrict__ rot, long V) {
    const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int64_t b = best[v];
    lcc[v] = __uint_as_float(unorderable_f32((int32_t)(b >> 32)));
    rot[v] = (int32_t)(0xFFFFFFFFu - (uint32_t)((uint64_t)b & 0xFFFFFFFFull));
}

__global__ void merge_kernel(int64_t *__restrict__ dst, const int64_t *__restrict__ src, long V) {
    const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < V) { const int64_t a = dst[v], b = src[v]; dst[v] = a > b ? a : b; }
}

// target (real) -> complex volumes (f, 0) and (f^2, 0)
__global__ void target_to_complex_kernel(const float *__restrict__ f, float2 *__restrict__ c1, float2 *__restrict__ c2, long V) {
    const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const float x = f[v];
    c1[v] = make_float2(x, 0.f);
    c2[v] = make_float2(x * x, 0.f);
}

// in place: F <- conj(F) / V
__global__ void conj_scale_kernel(float2 *__restrict__ c, long n, float scale) {
    const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const float2 x = c[v];
    c[v] = make_float2(x.x * scale, -x.y * scale);
}

static inline unsigned nblk(long n) { return (unsigned)((n + 255) / 256); }

int launch_multiply(Plan *p, int npairs, cudaStream_t s) {
    { LaunchScope ls(p, KC_MULTIPLY, s);
      multiply_kernel<<<dim3(nblk(p->V), npairs), 256, 0, s>>>(p->A, p->B, p->F, p->F2, p->V, p->nsig); }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

int launch_lcc_best(Plan *p, int first_rot_index, int count, int64_t *best, cudaStream_t s) {
    { LaunchScope ls(p, KC_LCC, s);
      lcc_best_kernel<<<nblk(p->V), 256, 0, s>>>(p->B, p->lcc_mask, p->norm_factor, first_rot_index, count, best, p->V); }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

int launch_lcc_single(Plan *p, const float *gcc, const float *ave, const float *ave2, float norm, int rot_index,
                      int64_t *best, cudaStream_t s) {
    { LaunchScope ls(p, KC_OTHER, s);
      lcc_single_kernel<<<nblk(p->V), 256, 0, s>>>(gcc, ave, ave2, p->lcc_mask, norm, rot_index, best, p->V); }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

int launch_best_init(Plan *p, int64_t *best, cudaStream_t s) {
    { LaunchScope ls(p, KC_OTHER, s);
      best_init_kernel<<<nblk(p->V), 256, 0, s>>>(best, p->V); }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

int launch_unpack(Plan *p, const int64_t *best, float *lcc, int32_t *rot, cudaStream_t s) {
    { LaunchScope ls(p, KC_OTHER, s);
      unpack_kernel<<<nblk(p->V), 256, 0, s>>>(best, lcc, rot, p->V); }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

int launch_merge(Plan *p, int64_t *dst, const int64_t *src, cudaStream_t s) {
    { LaunchScope ls(p, KC_OTHER, s);
      merge_kernel<<<nblk(p->V), 256, 0, s>>>(dst, src, p->V); }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

// FFT(f), FFT(f^2) in the form the search multiplies with: conj(P f)/V (see fft_generic.cu).
int launch_target_spectra(Plan *p, const float *target, cudaStream_t s) {
    { LaunchScope ls(p, KC_OTHER, s);
      target_to_complex_kernel<<<nblk(p->V), 256, 0, s>>>(target, p->F, p->F2, p->V); }
    PFB_CUDA(cudaGetLastError());
    for (int axis = 0; axis < 3; ++axis) {
        int rc = launch_fft_axis(p, p->F, 1, axis, s);
        if (rc) return rc;
        rc = launch_fft_axis(p, p->F2, 1, axis, s);
        if (rc) return rc;
    }
    const float scale = 1.0f / (float)p->V;
    { LaunchScope ls(p, KC_OTHER, s);
      conj_scale_kernel<<<nblk(p->V), 256, 0, s>>>(p->F, p->V, scale); }
    { LaunchScope ls(p, KC_OTHER, s);
      conj_scale_kernel<<<nblk(p->V), 256, 0, s>>>(p->F2, p->V, scale); }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

}  // namespace pfb
