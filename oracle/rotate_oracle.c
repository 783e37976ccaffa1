/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by powerfit_b200/.
 *
 * Plain-C restatement of the reference's FP64 grid rotation
 *   rotate_grid3d(grid, rotmat, radius, out, nearest)
 * following /root/reference/src/powerfit_em/_extensions.c:7-196 (semantics, loop
 * order and floating-point evaluation order), written from the behaviour and not
 * from the text.  Parity of this file is PINNED: tests/test_oracle.py checks it
 * against the golden vectors generated from the real reference extension
 * (tests/golden/make_golden.py) and, when oracle/_ref exists, against the
 * reference's own compiled _extensions module bit for bit.
 *
 * Semantics restated (file:line are into the reference):
 *  - integer offsets (x,y,z) in [-radius, radius]^3 with x^2+y^2+z^2 <= radius^2
 *    are visited z-outer, y-middle, x-inner (:55-93); the output voxel is the
 *    offset wrapped ONCE by +n when negative (:65-67, :78-80, :91-93).  No
 *    reduction for offsets >= n: when n == 2*radius the offsets +radius and
 *    -radius hit the same voxel and the later (+radius) write wins.
 *  - source coordinate = R^T * offset, accumulated z-term first, then y, then x
 *    (:60-62, :74-76, :87-89).
 *  - nearest: C round() (half away from zero), negative index wrapped once
 *    (:97-110).  trilinear: floor(), weights d and 1-d, neighbour +1 wraps to 0
 *    only when it is exactly 0 (:114-178); combination order x, then y, then z.
 *  - voxels outside the sphere are never written.
 */
#include <math.h>
#include <stddef.h>

typedef struct { long nz, ny, nx; } dims_t;

static inline long wrap_neg(long i, long n) { return i < 0 ? i + n : i; }

/* linear index of source lattice point (k,j,i) under the reference's "add n once
 * if negative" rule */
static inline long src_index(const dims_t *d, long k, long j, long i)
{
    return (wrap_neg(k, d->nz) * d->ny + wrap_neg(j, d->ny)) * d->nx + wrap_neg(i, d->nx);
}

static inline double lerp_ref(double a, double b, double w1, double w)
{
    return a * w1 + b * w;   /* lower * (1-d) + upper * d, as in :140-141 */
}

void pfo_rotate_grid3d(const double *grid, long gnz, long gny, long gnx,
                       const double *rotmat, int radius,
                       double *out, long onz, long ony, long onx, int nearest)
{
    const dims_t g = { gnz, gny, gnx };
    const long r2 = (long)radius * radius;

    for (long z = -radius; z <= radius; ++z) {
        const long dz2 = z * z;
        if (dz2 > r2) continue;
        const double ax_z = rotmat[6] * z, ay_z = rotmat[7] * z, az_z = rotmat[8] * z;
        const long oz = wrap_neg(z, onz);
        for (long y = -radius; y <= radius; ++y) {
            const long dzy2 = dz2 + y * y;
            if (dzy2 > r2) continue;
            const double ax_zy = ax_z + rotmat[3] * y;
            const double ay_zy = ay_z + rotmat[4] * y;
            const double az_zy = az_z + rotmat[5] * y;
            const long oy = wrap_neg(y, ony);
            for (long x = -radius; x <= radius; ++x) {
                if (dzy2 + x * x > r2) continue;
                const double sx = ax_zy + rotmat[0] * x;
                const double sy = ay_zy + rotmat[1] * x;
                const double sz = az_zy + rotmat[2] * x;
                double *dst = out + (oz * ony + oy) * onx + wrap_neg(x, onx);

                if (nearest > 0) {
                    *dst = grid[src_index(&g, (long)round(sz), (long)round(sy), (long)round(sx))];
                    continue;
                }
                const long i0 = (long)floor(sx), j0 = (long)floor(sy), k0 = (long)floor(sz);
                const double wx = sx - i0, wy = sy - j0, wz = sz - k0;
                const double wx1 = 1 - wx, wy1 = 1 - wy, wz1 = 1 - wz;
                /* the +1 neighbour: index i0+1, folded to 0 only when it equals 0
                 * relative to a negative i0 (i0 == -1); src_index does that because
                 * wrap_neg(-1+1) == 0 and wrap_neg(i0+1 < 0) adds n once. */
                const long i1 = i0 + 1, j1 = j0 + 1, k1 = k0 + 1;
                const double c00 = lerp_ref(grid[src_index(&g, k0, j0, i0)], grid[src_index(&g, k0, j0, i1)], wx1, wx);
                const double c10 = lerp_ref(grid[src_index(&g, k0, j1, i0)], grid[src_index(&g, k0, j1, i1)], wx1, wx);
                const double c01 = lerp_ref(grid[src_index(&g, k1, j0, i0)], grid[src_index(&g, k1, j0, i1)], wx1, wx);
                const double c11 = lerp_ref(grid[src_index(&g, k1, j1, i0)], grid[src_index(&g, k1, j1, i1)], wx1, wx);
                const double c0 = c00 * wy1 + c10 * wy;
                const double c1 = c01 * wy1 + c11 * wy;
                *dst = c0 * wz1 + c1 * wz;
            }
        }
    }
}

/* conj(a)*b over interleaved complex128 -- _powerfit.pyx:45-53 */
void pfo_conj_multiply(const double *a, const double *b, double *o, long n)
{
    for (long i = 0; i < n; ++i) {
        const double ar = a[2 * i], ai = a[2 * i + 1], br = b[2 * i], bi = b[2 * i + 1];
        o[2 * i]     = ar * br + ai * bi;
        o[2 * i + 1] = ar * bi - ai * br;
    }
}

/* lcc = gcc / sqrt(ave2 - ave^2) where mask != 0; untouched elsewhere; no guard
 * on a non-positive variance -- _powerfit.pyx:56-72 */
void pfo_calc_lcc(const double *gcc, const double *ave, const double *ave2,
                  const unsigned char *mask, double *lcc, long n)
{
    for (long i = 0; i < n; ++i)
        if (mask[i]) lcc[i] = gcc[i] / sqrt(ave2[i] - ave[i] * ave[i]);
}
