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
#include "rotate_device.cuh"

namespace pfb {

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
