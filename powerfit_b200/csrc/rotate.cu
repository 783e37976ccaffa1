// Rotation of the template / mask into the zero-padded search grid.
//
// Semantics are those of the reference's CPU operator rotate_grid3d
// (/root/reference/src/powerfit_em/_extensions.c:7-196), which is the parity target:
// out(r) = interp(grid, R^T r) for integer offsets r inside the rmax sphere, offsets
// and source indices wrapped periodically, nearest = round-half-away-from-zero,
// trilinear otherwise; nothing outside the sphere (written as 0 here because every
// pass owns its output buffer).  The reference walks the sphere and scatters; these
// kernels walk the OUTPUT grid (one thread per voxel, coalesced stores) and invert the
// wrap, which also makes the +rmax / -rmax alias on even axes deterministic: the
// reference's later write (+rmax) wins, so a voxel index i <= rmax is always read as
// the non-negative offset.
//
// Source coordinates are evaluated in FP64 with the reference's association order
// (z term, + y term, + x term; _extensions.c:60-62,74-76,87-89) and without FMA
// contraction, so round()/floor() decide exactly like the CPU code; interpolation
// weights and values are FP32.
#include "common.cuh"

namespace pfb {

struct GridDims {
    int nz, ny, nx, rmax;
    long V;
};

__device__ __forceinline__ bool signed_offset(int i, int n, int rmax, int &o) {
    if (i <= rmax) { o = i; return true; }
    o = i - n;
    return o >= -rmax;
}

__device__ __forceinline__ int wrap_index(int i, int n) {
    if (i < 0) i += n;
    else if (i >= n) i -= n;
    return i;
}

struct SrcCoord { double x, y, z; };

__device__ __forceinline__ SrcCoord source_coord(const double *__restrict__ R, int x, int y, int z) {
    SrcCoord c;
    c.x = __dadd_rn(__dadd_rn(__dmul_rn(R[6], (double)z), __dmul_rn(R[3], (double)y)), __dmul_rn(R[0], (double)x));
    c.y = __dadd_rn(__dadd_rn(__dmul_rn(R[7], (double)z), __dmul_rn(R[4], (double)y)), __dmul_rn(R[1], (double)x));
    c.z = __dadd_rn(__dadd_rn(__dmul_rn(R[8], (double)z), __dmul_rn(R[5], (double)y)), __dmul_rn(R[2], (double)x));
    return c;
}

__device__ __forceinline__ float sample_nearest(const float *__restrict__ g, const GridDims &d, const SrcCoord &c) {
    const int i = wrap_index((int)round(c.x), d.nx);
    const int j = wrap_index((int)round(c.y), d.ny);
    const int k = wrap_index((int)round(c.z), d.nz);
    return __ldg(g + ((long)k * d.ny + j) * d.nx + i);
}

__device__ __forceinline__ float sample_trilinear(const float *__restrict__ g, const GridDims &d, const SrcCoord &c) {
    const double fx = floor(c.x), fy = floor(c.y), fz = floor(c.z);
    const float wx = (float)(c.x - fx), wy = (float)(c.y - fy), wz = (float)(c.z - fz);
    const float wx1 = 1.f - wx, wy1 = 1.f - wy, wz1 = 1.f - wz;
    const int i0 = wrap_index((int)fx, d.nx), i1 = wrap_index((int)fx + 1, d.nx);
    const int j0 = wrap_index((int)fy, d.ny), j1 = wrap_index((int)fy + 1, d.ny);
    const int k0 = wrap_index((int)fz, d.nz), k1 = wrap_index((int)fz + 1, d.nz);
    const float *r00 = g + ((long)k0 * d.ny + j0) * d.nx;
    const float *r10 = g + ((long)k0 * d.ny + j1) * d.nx;
    const float *r01 = g + ((long)k1 * d.ny + j0) * d.nx;
    const float *r11 = g + ((long)k1 * d.ny + j1) * d.nx;
    const float c00 = __ldg(r00 + i0) * wx1 + __ldg(r00 + i1) * wx;
    const float c10 = __ldg(r10 + i0) * wx1 + __ldg(r10 + i1) * wx;
    const float c01 = __ldg(r01 + i0) * wx1 + __ldg(r01 + i1) * wx;
    const float c11 = __ldg(r11 + i0) * wx1 + __ldg(r11 + i1) * wx;
    const float c0 = c00 * wy1 + c10 * wy;
    const float c1 = c01 * wy1 + c11 * wy;
    return c0 * wz1 + c1 * wz;
}

// One thread per (voxel, rotation pair).  Writes the pair-packed complex volumes
//   A[pair][0][v] = (t_a, t_b)   A[pair][1][v] = (m_a, m_b)   [A[pair][2][v] = (m_a^2, m_b^2)]
// where a = first + 2*pair, b = a + 1 (b beyond `count` contributes zeros).
__global__ void __launch_bounds__(256)
rotate_pack_kernel(const float *__restrict__ tmpl, const float *__restrict__ mask,
                   const double *__restrict__ rot, int first, int count, int nsig,
                   float2 *__restrict__ A, GridDims d) {
    const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= d.V) return;
    const int pair = blockIdx.y;
    const int ix = (int)(v % d.nx);
    const int iy = (int)((v / d.nx) % d.ny);
    const int iz = (int)(v / ((long)d.nx * d.ny));
    float2 t = make_float2(0.f, 0.f), m = make_float2(0.f, 0.f);
    int ox, oy, oz;
    if (signed_offset(ix, d.nx, d.rmax, ox) && signed_offset(iy, d.ny, d.rmax, oy) &&
        signed_offset(iz, d.nz, d.rmax, oz) && ox * ox + oy * oy + oz * oz <= d.rmax * d.rmax) {
        const int a = 2 * pair;
        {
            const SrcCoord c = source_coord(rot + (long)(first + a) * 9, ox, oy, oz);
            t.x = sample_trilinear(tmpl, d, c);
            m.x = sample_nearest(mask, d, c);
        }
        if (a + 1 < count) {
            const SrcCoord c = source_coord(rot + (long)(first + a + 1) * 9, ox, oy, oz);
            t.y = sample_trilinear(tmpl, d, c);
            m.y = sample_nearest(mask, d, c);
        }
    }
    float2 *base = A + (long)pair * nsig * d.V;
    base[v] = t;
    base[d.V + v] = m;
    if (nsig == 3) base[2 * d.V + v] = make_float2(m.x * m.x, m.y * m.y);
}

// Operator-level twin of rotate_grid3d: out[r][v], one rotation per blockIdx.y.
__global__ void __launch_bounds__(256)
rotate_plain_kernel(const float *__restrict__ grid, const double *__restrict__ rot, int nearest,
                    float *__restrict__ out, GridDims d) {
    const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= d.V) return;
    const int ix = (int)(v % d.nx);
    const int iy = (int)((v / d.nx) % d.ny);
    const int iz = (int)(v / ((long)d.nx * d.ny));
    float val = 0.f;
    int ox, oy, oz;
    if (signed_offset(ix, d.nx, d.rmax, ox) && signed_offset(iy, d.ny, d.rmax, oy) &&
        signed_offset(iz, d.nz, d.rmax, oz) && ox * ox + oy * oy + oz * oz <= d.rmax * d.rmax) {
        const SrcCoord c = source_coord(rot + (long)blockIdx.y * 9, ox, oy, oz);
        val = nearest ? sample_nearest(grid, d, c) : sample_trilinear(grid, d, c);
    }
    out[(long)blockIdx.y * d.V + v] = val;
}

static GridDims dims_of(const Plan *p) { return GridDims{p->nz, p->ny, p->nx, p->rmax, p->V}; }

int launch_rotate_pack(Plan *p, const double *rot_dev, int first, int count, cudaStream_t s) {
    const int npairs = (count + 1) / 2;
    dim3 grid((unsigned)((p->V + 255) / 256), npairs);
    { LaunchScope ls(p, KC_ROTATE, s);
      rotate_pack_kernel<<<grid, 256, 0, s>>>(p->tmpl, p->mask, rot_dev, first, count, p->nsig, p->A, dims_of(p)); }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

int launch_rotate_plain(Plan *p, const float *g, const double *rot_dev, int R, int nearest, float *out,
                        cudaStream_t s) {
    dim3 grid((unsigned)((p->V + 255) / 256), R);
    { LaunchScope ls(p, KC_ROTATE, s);
      rotate_plain_kernel<<<grid, 256, 0, s>>>(g, rot_dev, nearest, out, dims_of(p)); }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

}  // namespace pfb
