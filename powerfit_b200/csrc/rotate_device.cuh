// Device helpers shared by the rotation kernels (generic and fused paths).
// Semantics: /root/reference/src/powerfit_em/_extensions.c:7-196 (see rotate.cu).
#pragma once
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

// The same sums with the (z, y) part of a row computed once: row = (R6 z + R3 y, R7 z + R4 y, R8 z + R5 y), then
// + R0..2 x per sample -- the reference's own loop nest (_extensions.c:60-89), identical roundings.
__device__ __forceinline__ SrcCoord source_row(const double *__restrict__ R, int y, int z) {
    SrcCoord c;
    c.x = __dadd_rn(__dmul_rn(R[6], (double)z), __dmul_rn(R[3], (double)y));
    c.y = __dadd_rn(__dmul_rn(R[7], (double)z), __dmul_rn(R[4], (double)y));
    c.z = __dadd_rn(__dmul_rn(R[8], (double)z), __dmul_rn(R[5], (double)y));
    return c;
}
__device__ __forceinline__ SrcCoord source_in_row(const SrcCoord &row, const double *__restrict__ R, int x) {
    SrcCoord c;
    c.x = __dadd_rn(row.x, __dmul_rn(R[0], (double)x));
    c.y = __dadd_rn(row.y, __dmul_rn(R[1], (double)x));
    c.z = __dadd_rn(row.z, __dmul_rn(R[2], (double)x));
    return c;
}
// largest h >= 0 with h*h <= v (v >= 0)
__device__ __forceinline__ int isqrt_floor(int v) {
    int h = (int)sqrtf((float)v);
    while (h * h > v) --h;
    while ((h + 1) * (h + 1) <= v) ++h;
    return h;
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

// Same arithmetic as sample_trilinear from a corner table q[v] = (g[z][y][x], g[z][y][x+1],
// g[z][y+1][x], g[z][y+1][x+1]) (periodic): two 16-byte loads instead of eight scalar gathers.
__device__ __forceinline__ float sample_trilinear_q(const float4 *__restrict__ q, const GridDims &d, const SrcCoord &c) {
    const double fx = floor(c.x), fy = floor(c.y), fz = floor(c.z);
    const float wx = (float)(c.x - fx), wy = (float)(c.y - fy), wz = (float)(c.z - fz);
    const float wx1 = 1.f - wx, wy1 = 1.f - wy, wz1 = 1.f - wz;
    const int i0 = wrap_index((int)fx, d.nx), j0 = wrap_index((int)fy, d.ny);
    const int k0 = wrap_index((int)fz, d.nz), k1 = wrap_index((int)fz + 1, d.nz);
    const float4 a = __ldg(q + ((long)k0 * d.ny + j0) * d.nx + i0);
    const float4 b = __ldg(q + ((long)k1 * d.ny + j0) * d.nx + i0);
    const float c00 = a.x * wx1 + a.y * wx;
    const float c10 = a.z * wx1 + a.w * wx;
    const float c01 = b.x * wx1 + b.y * wx;
    const float c11 = b.z * wx1 + b.w * wx;
    const float c0 = c00 * wy1 + c10 * wy;
    const float c1 = c01 * wy1 + c11 * wy;
    return c0 * wz1 + c1 * wz;
}

}  // namespace pfb
