// Map resampling for the multi-scale image pyramid: the two array
// operations of /root/reference/src/powerfit_em/scripts/__init__.py:93-103,
//
//   pfb_gaussian_filter  volume.lower_resolution (volume.py:129-141) -> scipy.ndimage.gaussian_filter(
//                        array, sigma, mode='constant'): three separable passes (axis 0, 1, 2), each
//                        out[i] = x[i] w[0] + sum_{j = radius..1} (x[i-j] + x[i+j]) w[j] with zeros outside the
//                        array -- scipy's symmetric correlate1d, same order of additions, so FP64-identical
//   pfb_zoom_linear      volume.resample (volume.py:66-72) -> scipy.ndimage.zoom(array, factor, order=1):
//                        trilinear interpolation at x_in = x_out (n_in - 1) / (n_out - 1), zero where that
//                        product exceeds n_in - 1 (scipy's mode='constant' edge behaviour, reproduced)
//
// in FP64 on device grids of the current device.  No plan is needed.
#include "common.cuh"

#include <algorithm>

namespace pfb {

// one separable pass along `axis` (0 = z, 1 = y, 2 = x); w[j], j <= radius, is the kernel's right half
__global__ void gauss_pass_kernel(const double *__restrict__ in, double *__restrict__ out, int nz, int ny, int nx,
                                  int axis, const double *__restrict__ w, int radius) {
    const long V = (long)nz * ny * nx;
    const int n = axis == 0 ? nz : (axis == 1 ? ny : nx);
    const long stride = axis == 0 ? (long)ny * nx : (axis == 1 ? nx : 1);
    for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long)gridDim.x * blockDim.x) {
        const int i = axis == 0 ? (int)(v / ((long)ny * nx)) : (axis == 1 ? (int)((v / nx) % ny) : (int)(v % nx));
        double acc = __dmul_rn(in[v], w[0]);
        for (int j = radius; j >= 1; --j) {
            const double lo = i - j >= 0 ? in[v - j * stride] : 0.0, hi = i + j < n ? in[v + j * stride] : 0.0;
            acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(lo, hi), w[j]));
        }
        out[v] = acc;
    }
}

__global__ void zoom_linear_kernel(const double *__restrict__ in, int nz, int ny, int nx, double *__restrict__ out,
                                   int oz, int oy, int ox) {
    const long V = (long)oz * oy * ox;
    const double sz = oz > 1 ? (double)(nz - 1) / (double)(oz - 1) : 0.0;
    const double sy = oy > 1 ? (double)(ny - 1) / (double)(oy - 1) : 0.0;
    const double sx = ox > 1 ? (double)(nx - 1) / (double)(ox - 1) : 0.0;
    for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long)gridDim.x * blockDim.x) {
        const int x = (int)(v % ox), y = (int)((v / ox) % oy), z = (int)(v / ((long)ox * oy));
        const double cz = z * sz, cy = y * sy, cx = x * sx;
        // scipy's zoom runs in mode='constant': a coordinate that rounding pushes past the last sample
        // (o * (n_in - 1)/(n_out - 1) > n_in - 1 by one ulp) counts as outside and yields 0
        if (cz > (double)(nz - 1) || cy > (double)(ny - 1) || cx > (double)(nx - 1)) { out[v] = 0.0; continue; }
        const int z0 = min((int)floor(cz), nz - 1), y0 = min((int)floor(cy), ny - 1), x0 = min((int)floor(cx), nx - 1);
        const int z1 = min(z0 + 1, nz - 1), y1 = min(y0 + 1, ny - 1), x1 = min(x0 + 1, nx - 1);
        const double tz = cz - z0, ty = cy - y0, tx = cx - x0;
        auto at = [&](int zz, int yy, int xx) { return in[((long)zz * ny + yy) * nx + xx]; };
        const double c00 = at(z0, y0, x0) * (1.0 - tx) + at(z0, y0, x1) * tx;
        const double c01 = at(z0, y1, x0) * (1.0 - tx) + at(z0, y1, x1) * tx;
        const double c10 = at(z1, y0, x0) * (1.0 - tx) + at(z1, y0, x1) * tx;
        const double c11 = at(z1, y1, x0) * (1.0 - tx) + at(z1, y1, x1) * tx;
        const double c0 = c00 * (1.0 - ty) + c01 * ty, c1 = c10 * (1.0 - ty) + c11 * ty;
        out[v] = c0 * (1.0 - tz) + c1 * tz;
    }
}

}  // namespace pfb

using namespace pfb;

extern "C" {

int pfb_gaussian_filter(const double *in, double *out, double *tmp, int nz, int ny, int nx, const double *weights,
                        int radius, void *stream) {
    PFB_REQUIRE(in && out && tmp && weights && nz > 0 && ny > 0 && nx > 0 && radius >= 0,
                "pfb_gaussian_filter: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    const long V = (long)nz * ny * nx;
    const int blocks = (int)std::min<long>((V + 255) / 256, 148 * 16);
    gauss_pass_kernel<<<blocks, 256, 0, s>>>(in, out, nz, ny, nx, 0, weights, radius);
    gauss_pass_kernel<<<blocks, 256, 0, s>>>(out, tmp, nz, ny, nx, 1, weights, radius);
    gauss_pass_kernel<<<blocks, 256, 0, s>>>(tmp, out, nz, ny, nx, 2, weights, radius);
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

int pfb_zoom_linear(const double *in, int nz, int ny, int nx, double *out, int oz, int oy, int ox, void *stream) {
    PFB_REQUIRE(in && out && nz > 0 && ny > 0 && nx > 0 && oz > 0 && oy > 0 && ox > 0, "pfb_zoom_linear: bad argument");
    const long V = (long)oz * oy * ox;
    const int blocks = (int)std::min<long>((V + 255) / 256, 148 * 16);
    zoom_linear_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(in, nz, ny, nx, out, oz, oy, ox);
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

}  // extern "C"
