// Template and mask synthesis from atom coordinates -- the step
// immediately before the search path (/root/reference/src/powerfit_em/powerfit.py:245-267):
//
//   pfb_blur_points    _powerfit.blur_points   (_powerfit.pyx:75-138)  Gaussian splat, 4 sigma cut-off
//   pfb_dilate_points  _powerfit.dilate_points (_powerfit.pyx:141-206) union of balls
//   pfb_core_indices   helpers.determine_core_indices (helpers.py:26-34) erosion depth of the mask
//
// all with the reference's wraparound=True indexing (positions -n+1 .. n-1 map to index mod n).
// The reference scatters atom by atom; here every voxel gathers from the atoms in the same
// order n = 0, 1, ..., so each FP64 sum is accumulated in the reference's order and the result
// is deterministic.  The distance arithmetic repeats the reference's expressions in FP64, so
// the inside/outside decisions (and hence dilate_points and core_indices) are exact; blur
// values differ from the reference only through exp() (CUDA's is within 1 ulp, like libm's).
// No plan is needed: the calls work on any FP64 device grid of the current device.
#include "common.cuh"

#include <algorithm>

namespace pfb {

// candidate positions along one axis that land on index i: p = i and p = i - n, kept if inside
// [lo, hi] (the reference's clipped loop range); the negative one comes first in loop order
__device__ __forceinline__ int axis_candidates(int i, int n, int lo, int hi, int (&p)[2]) {
    int c = 0;
    if (i > 0 && i - n >= lo && i - n <= hi) p[c++] = i - n;
    if (i >= lo && i <= hi) p[c++] = i;
    return c;
}

// MODE 0: out += sum_n w_n exp(-d2 / (2 sigma^2)) for d2 <= (4 sigma)^2     (blur_points)
// MODE 1: out = 1 where any d2 <= r_n^2                                       (dilate_points)
template <int MODE>
__global__ void __launch_bounds__(256)
points_kernel(const double *__restrict__ pts, const double *__restrict__ wr, int npts, double sigma, int nz, int ny,
              int nx, double *__restrict__ out) {
    __shared__ double sx[256], sy[256], sz[256], sw[256];
    const long V = (long)nz * ny * nx;
    const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = v < V;
    const int x = live ? (int)(v % nx) : 0, y = live ? (int)((v / nx) % ny) : 0, z = live ? (int)(v / ((long)nx * ny)) : 0;
    const double extend = 4.0 * sigma, extend2 = extend * extend, dsigma2 = 2.0 * sigma * sigma;
    double acc = live ? out[v] : 0.0;
    bool hit = false;
    for (int base = 0; base < npts; base += 256) {
        __syncthreads();
        const int n = base + threadIdx.x;
        if (n < npts) {
            sx[threadIdx.x] = pts[n];
            sy[threadIdx.x] = pts[(long)npts + n];
            sz[threadIdx.x] = pts[2L * npts + n];
            sw[threadIdx.x] = wr[n];
        }
        __syncthreads();
        if (!live) continue;
        const int cnt = min(256, npts - base);
        for (int k = 0; k < cnt; ++k) {
            const double px = sx[k], py = sy[k], pz = sz[k];
            const double reach = MODE == 0 ? extend : sw[k];
            const double reach2 = MODE == 0 ? extend2 : __dmul_rn(reach, reach);
            // loop ranges of the reference: ceil(p - reach) .. floor(p + reach), clipped to -n+1 .. n-1
            int cz[2], cy[2], cx[2];
            const int nzc = axis_candidates(z, nz, max((int)ceil(pz - reach), -nz + 1), min((int)floor(pz + reach), nz - 1), cz);
            if (nzc == 0) continue;
            const int nyc = axis_candidates(y, ny, max((int)ceil(py - reach), -ny + 1), min((int)floor(py + reach), ny - 1), cy);
            if (nyc == 0) continue;
            const int nxc = axis_candidates(x, nx, max((int)ceil(px - reach), -nx + 1), min((int)floor(px + reach), nx - 1), cx);
            if (nxc == 0) continue;
            for (int a = 0; a < nzc; ++a) {
                const double dz = (double)cz[a] - pz;
                const double z2 = __dmul_rn(dz, dz);
                for (int b = 0; b < nyc; ++b) {
                    const double dy = (double)cy[b] - py;
                    const double y2z2 = __dadd_rn(__dmul_rn(dy, dy), z2);
                    for (int c = 0; c < nxc; ++c) {
                        const double dx = (double)cx[c] - px;
                        const double d2 = __dadd_rn(__dmul_rn(dx, dx), y2z2);
                        if (d2 <= reach2) {
                            if (MODE == 0) acc = __dadd_rn(acc, __dmul_rn(sw[k], exp(-d2 / dsigma2)));
                            else hit = true;
                        }
                    }
                }
            }
        }
    }
    if (live) {
        if (MODE == 0) out[v] = acc;
        else if (hit) out[v] = 1.0;
    }
}

// one erosion step of scipy.ndimage.binary_erosion (6-neighbour cross, border_value = 0: voxels on
// the array faces always erode -- the reference does not wrap here): core += cur; nxt = eroded cur
__global__ void core_step_kernel(const uint8_t *__restrict__ cur, uint8_t *__restrict__ nxt, double *__restrict__ core,
                                 int nz, int ny, int nx, int *__restrict__ remaining) {
    const long V = (long)nz * ny * nx;
    int mine = 0;
    for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long)gridDim.x * blockDim.x) {
        const uint8_t c = cur[v];
        uint8_t e = 0;
        if (c) {
            core[v] += 1.0;
            const int x = (int)(v % nx), y = (int)((v / nx) % ny), z = (int)(v / ((long)nx * ny));
            const long sy = nx, sz = (long)nx * ny;
            e = x > 0 && x < nx - 1 && y > 0 && y < ny - 1 && z > 0 && z < nz - 1 && cur[v - 1] && cur[v + 1] &&
                cur[v - sy] && cur[v + sy] && cur[v - sz] && cur[v + sz];
        }
        nxt[v] = e;
        mine += e;
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(remaining, mine);
}

__global__ void core_init_kernel(const double *__restrict__ mask, uint8_t *__restrict__ cur, double *__restrict__ core,
                                 long V, int *__restrict__ remaining) {
    int mine = 0;
    for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long)gridDim.x * blockDim.x) {
        const uint8_t c = mask[v] > 0.0;
        cur[v] = c;
        core[v] = 0.0;
        mine += c;
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(remaining, mine);
}

}  // namespace pfb

using namespace pfb;

extern "C" {

int pfb_blur_points(const double *points, const double *weights, int n, double sigma, int nz, int ny, int nx,
                    double *out, void *stream) {
    PFB_REQUIRE(points && weights && out && n >= 0 && nz > 0 && ny > 0 && nx > 0 && sigma > 0.0,
                "pfb_blur_points: bad argument");
    const long V = (long)nz * ny * nx;
    points_kernel<0><<<(unsigned)((V + 255) / 256), 256, 0, (cudaStream_t)stream>>>(points, weights, n, sigma, nz, ny,
                                                                                   nx, out);
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

int pfb_dilate_points(const double *points, const double *radii, int n, int nz, int ny, int nx, double *out,
                      void *stream) {
    PFB_REQUIRE(points && radii && out && n >= 0 && nz > 0 && ny > 0 && nx > 0, "pfb_dilate_points: bad argument");
    const long V = (long)nz * ny * nx;
    points_kernel<1><<<(unsigned)((V + 255) / 256), 256, 0, (cudaStream_t)stream>>>(points, radii, n, 1.0, nz, ny, nx,
                                                                                   out);
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

int pfb_core_indices(const double *mask, int nz, int ny, int nx, double *core, uint8_t *scratch, void *stream) {
    PFB_REQUIRE(mask && core && scratch && nz > 0 && ny > 0 && nx > 0, "pfb_core_indices: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    const long V = (long)nz * ny * nx;
    uint8_t *cur = scratch, *nxt = scratch + V;
    int *remaining = reinterpret_cast<int *>(scratch + 2 * V + ((8 - (2 * V) % 8) % 8));
    const int blocks = (int)std::min<long>((V + 255) / 256, 148 * 8);
    int h = 0;
    PFB_CUDA(cudaMemsetAsync(remaining, 0, sizeof(int), s));
    core_init_kernel<<<blocks, 256, 0, s>>>(mask, cur, core, V, remaining);
    PFB_CUDA(cudaMemcpyAsync(&h, remaining, sizeof(int), cudaMemcpyDeviceToHost, s));
    PFB_CUDA(cudaStreamSynchronize(s));
    // while eroded_mask.sum() > 0: core += eroded_mask; eroded_mask = binary_erosion(eroded_mask)
    for (int iter = 0; h > 0; ++iter) {
        PFB_REQUIRE(iter <= nz + ny + nx, "pfb_core_indices: erosion did not terminate");
        PFB_CUDA(cudaMemsetAsync(remaining, 0, sizeof(int), s));
        core_step_kernel<<<blocks, 256, 0, s>>>(cur, nxt, core, nz, ny, nx, remaining);
        PFB_CUDA(cudaMemcpyAsync(&h, remaining, sizeof(int), cudaMemcpyDeviceToHost, s));
        PFB_CUDA(cudaStreamSynchronize(s));
        std::swap(cur, nxt);
    }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

}  // extern "C"
