/*
 * powerfit_b200 -- C ABI of the B200-native exhaustive LCC search.
 *
 * This is the drop-in boundary for the reference's `--gpu` correlator
 * (/root/reference/src/powerfit_em/powerfitter.py:396-555, class GPUCorrelator, and the
 * OpenCL/clFFT operators underneath it).  The reference has no FFI of its own for this
 * path: its GPU backend is reached through pyopencl + gpyfft from Python.  The entry
 * points below are what a maintainer binds with ctypes in place of those two packages
 * (see INTEGRATION.md for the stub); each comment names the reference interface the
 * function replaces.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no C++/torch types.
 *  - Every function returns 0 on success and a non-zero code on failure;
 *    pfb_last_error() returns a thread-local description of the last failure.
 *  - Unless a parameter is documented as HOST, pointers are CUDA device pointers on the
 *    plan's device.  `stream` is a cudaStream_t passed as void* (NULL = legacy default
 *    stream).  Calls only enqueue work; the caller synchronises the stream.
 *  - Grids are C-ordered (nz, ny, nx), x fastest, exactly like the reference's numpy
 *    arrays.  Template and mask are centred on voxel (0,0,0) with periodic wrap-around.
 *  - A plan is bound to one device and one grid shape.  It is not thread-safe; distinct
 *    plans may be used concurrently.
 *  - Supported axis lengths: any product of 2, 3, 5 and 7 (the set the reference's clFFT
 *    backend supports and its CLI pads to, powerfit.py:230-233), each >= 2.
 *    Other lengths fail with PFB_ERR_UNSUPPORTED (there is no CPU fallback).
 */
#ifndef POWERFIT_B200_H
#define POWERFIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pfb_plan pfb_plan;

enum {
    PFB_OK = 0,
    PFB_ERR_INVALID = 1,       /* bad argument / call order (ValueError in the reference)  */
    PFB_ERR_UNSUPPORTED = 2,   /* axis length not 2.3.5.7-smooth, or shape too large        */
    PFB_ERR_CUDA = 3,          /* a CUDA runtime call failed; see pfb_last_error()          */
    PFB_ERR_NODEVICE = 4       /* no usable sm_100 device                                   */
};

/* Version / build info: "powerfit_b200 <ver> sm_100a". */
const char *pfb_version(void);
const char *pfb_last_error(void);

/* GPUCorrelator.__init__/_allocate_arrays/_build_ffts (powerfitter.py:399-460):
 * allocates every device buffer and FFT table for one (device, shape).
 * max_batch = rotations in flight per pass (rounded up to even; 0 = choose; halved until the work buffers fit
 * when device memory is short -- pfb_plan_info(plan, 4) reports the batch in use).
 * rmax is derived as min(nz,ny,nx)/2 (powerfitter.py:176).
 * Any 2.3.5.7-smooth shape up to 1024 per axis is accepted.  Shapes whose axes are 32, 64, 96 or 128 voxels long
 * (any mix) and the cubes 192^3 and 256^3 run the fused three-kernel pipeline (pfb_plan_info(plan, 6) == 1); every
 * other shape runs the any-shape pipeline: same results, about ten times slower. */
int pfb_plan_create(int nz, int ny, int nx, int max_batch, int device, pfb_plan **out);
int pfb_plan_destroy(pfb_plan *plan);

/* Query: 0 nz, 1 ny, 2 nx, 3 rmax, 4 batch, 5 device, 6 fused-path-available,
 * 7 kernel launches since plan creation (low 31 bits), 8 support radius of the template in
 * voxels (fused path), 9 class-decimated fused kernels in use (192^3, 256^3).
 * Environment switches read at plan creation (testing / tuning only): PFB_FUSED=0 forces the any-shape
 * pipeline, PFB_NO_PRUNE=1 disables support pruning, PFB_BATCH=n overrides the default batch. */
int pfb_plan_info(const pfb_plan *plan, int what, int64_t *value);

/* GPUCorrelator.__init__ (powerfitter.py:410-420): takes the normalised (and, if the
 * caller wants it, Laplace-filtered) target f and the lcc_mask (uint8, non-zero = score
 * this voxel) and precomputes FFT(f) and FFT(f^2). */
int pfb_set_target(pfb_plan *plan, const float *target, const uint8_t *lcc_mask, void *stream);

/* GPUCorrelator.mask setter (powerfitter.py:466-474): uploads the prepared template
 * (masked, z-scored, masked again -- powerfitter.py:204-220) and mask.  norm_factor is
 * the number of non-zero mask voxels (powerfitter.py:199).  mask_is_binary != 0 promises
 * mask in {0,1}, which lets the search skip the separate mask^2 transform. */
int pfb_set_template(pfb_plan *plan, const float *tmpl, const float *mask, float norm_factor,
                     int mask_is_binary, void *stream);

/* Several templates against ONE map (BASELINE configs[4]: "a batch of 4 distinct templates"; the reference
 * CLI builds one PowerFitter -- and recomputes FT(map), FT(map^2) -- per template, powerfit.py:245-282).
 * A plan has template slots: pfb_template_slots grows their number (slot 0 always exists),
 * pfb_select_template makes one current; pfb_set_template / pfb_prepare_template fill the current slot and
 * pfb_scan / pfb_search_host search with it.  The map spectra, lcc_mask and work buffers are shared. */
int pfb_template_slots(pfb_plan *plan, int nslots);
int pfb_select_template(pfb_plan *plan, int slot);

/* The same two setters with the reference's one-time preparation done on the device in FP64
 * (SURVEY.md 8f rows N3/N2).  All pointers DEVICE unless noted.
 * pfb_prepare_target = BaseCorrelator.__init__ + GPUCorrelator.__init__ (powerfitter.py:169-180,
 * 410-414): f = target / target.max(); lcc_mask = f > 0.05 f.max(); [scipy.ndimage.laplace(f,
 * mode='wrap')]; float32 cast -> f_out, lcc_mask_out (both kept by the caller), then pfb_set_target.
 * Bit-identical to the numpy/scipy formulas.
 * pfb_prepare_template = BaseCorrelator.mask.fset (powerfitter.py:190-220): N = count(mask != 0)
 * -> *norm_factor (HOST); [Laplace]; t *= m; z-score over mask != 0; t *= m; float32 casts ->
 * t_out, m_out; *mask_is_binary (HOST) = all(mask[mask != 0] == 1); then pfb_set_template.
 * Synchronises the stream.  PFB_ERR_INVALID "Zero-filled mask is not allowed." like :201-202. */
int pfb_prepare_target(pfb_plan *plan, const double *target, int laplace, float *f_out,
                       uint8_t *lcc_mask_out, void *stream);
int pfb_prepare_template(pfb_plan *plan, const double *tmpl, const double *mask, int laplace, float *t_out,
                         float *m_out, double *norm_factor, int *mask_is_binary, void *stream);

/* Reset a packed best grid to "LCC 0, rotation 0" (glcc.fill(0), grot.fill(0),
 * powerfitter.py:516-517). best = nz*ny*nx int64 packed keys: (orderable(lcc) << 32) | (0xFFFFFFFF - rot), signed order. */
int pfb_best_init(pfb_plan *plan, int64_t *best, void *stream);

/* GPUCorrelator.scan (powerfitter.py:513-538): for each of the R rotations
 * (rotmats: HOST pointer, R*9 doubles, row-major 3x3 each, powerfitter.py:226-230)
 * rotate template+mask, correlate with the target through FFTs, form the LCC and fold
 * it into `best` with the reference's strict-greater / lowest-index-wins rule.
 * Rotation n is recorded as index rot_index_offset + n (the offset is how a rank that
 * owns a block of the rotation list reports global indices, powerfitter.py:159). */
int pfb_scan(pfb_plan *plan, const double *rotmats_host, int R, int rot_index_offset,
             int64_t *best, void *stream);

/* Split the packed grid into the reference's outputs: lcc float32, rot int32
 * (powerfitter.py:435-436, 536-537). */
int pfb_unpack(pfb_plan *plan, const int64_t *best, float *lcc, int32_t *rot, void *stream);

/* Element-wise max of two packed grids (dst = max(dst, src)); the on-device form of
 * PowerFitter._combine (powerfitter.py:146-163) for partial results already on one GPU. */
int pfb_merge_best(pfb_plan *plan, int64_t *dst, const int64_t *src, void *stream);

/* Optional per-kernel timing for benchmarks: while enabled every launch is bracketed by
 * CUDA events on its stream (adds a little overhead; never on during a timed run).
 * pfb_profile(plan, 1) resets and starts, pfb_profile(plan, 0) stops.  pfb_profile_read
 * returns accumulated milliseconds and launch count of kernel class `cls` (0 <= cls,
 * PFB_ERR_INVALID past the last class) and its name.  (The reference has wall-clock
 * timing only, powerfit.py:281-283.) */
int pfb_profile(pfb_plan *plan, int enable);
int pfb_profile_read(pfb_plan *plan, int cls, double *ms, int64_t *launches, const char **name);

/* ---- operator-level entry points (unit parity with the reference operators) ---- */

/* _extensions.rotate_grid3d (_extensions.c:7-196) / kernels.cl rotate_image3d
 * (kernels.cl:162-226): out[r] (R grids of nz*ny*nx float32) = grid rotated by each
 * matrix (HOST, R*9 doubles); voxels outside the rmax sphere are written as 0. */
int pfb_rotate(pfb_plan *plan, const float *grid, const double *rotmats_host, int R, int nearest,
               float *out, void *stream);

/* grfftn_builder (powerfitter.py:605-638) / numpy.fft: in-place 3-D complex DFT with
 * kernel exp(+2*pi*i*k*r/n), un-normalised, of `nvol` interleaved-complex64 volumes. */
int pfb_fft3_c2c(pfb_plan *plan, float *vols_interleaved, int nvol, void *stream);

/* CLKernels.calc_lcc_and_take_best (powerfitter.py:572-584): gcc/ave/ave2 float32 grids of
 * one rotation; ave2 is multiplied by norm_factor inside, like the reference kernel. */
int pfb_lcc_take_best(pfb_plan *plan, const float *gcc, const float *ave, const float *ave2,
                      float norm_factor, int rot_index, int64_t *best, void *stream);

/* Whole search with HOST buffers (what the Python GPUCorrelator does around scan():
 * uploads, scan, downloads -- powerfitter.py:414-416, 471-474, 536-537).  All pointers
 * HOST.  lcc/rot receive nz*ny*nx values. */
int pfb_search_host(pfb_plan *plan, const float *target, const uint8_t *lcc_mask,
                    const float *tmpl, const float *mask, float norm_factor, int mask_is_binary,
                    const double *rotmats, int R, int rot_index_offset, float *lcc, int32_t *rot);

/* ---- solution extraction: device half of Analyzer._watershed ---- */

/* numpy `corr.max()` of analyzer.py:84 on a float32 device grid of n values (current device):
 * writes the maximum (NaN if any value is NaN, like numpy) to *max_out (device).  scratch = one
 * device int32. */
int pfb_lcc_max(const float *lcc, int64_t n, float *max_out, int32_t *scratch, void *stream);

/* The voxels that can belong to any labelled feature of analyzer.py:90-93: stream compaction of
 * {i : lcc[i] >= cutoff} into idx/val (device, capacity cap, arbitrary order).  *count (device)
 * receives the number found, which may exceed cap -- the caller then retries with a larger
 * buffer. */
int pfb_peak_candidates(const float *lcc, int64_t n, float cutoff, int32_t cap, int32_t *idx, float *val,
                        int32_t *count, void *stream);

/* ---- template / mask synthesis: what powerfit.py:245-267 calls before the search ---- */

/* _powerfit.blur_points(points, weights, sigma, out, wraparound=True) (_powerfit.pyx:75-138):
 * out[z,y,x] += w_n exp(-d^2 / (2 sigma^2)) for every atom n within 4 sigma, positions -n+1..n-1
 * wrapping to index mod n.  points = 3 x n doubles (x row, y row, z row, grid units), all DEVICE;
 * out = nz*ny*nx doubles, accumulated into.  Sums run in atom order like the reference's loop. */
int pfb_blur_points(const double *points, const double *weights, int n, double sigma, int nz, int ny, int nx,
                    double *out, void *stream);

/* _powerfit.dilate_points(points, radii, out, wraparound=True) (_powerfit.pyx:141-206):
 * out = 1 wherever a voxel lies within radii[n] of atom n (out is otherwise left as it is). */
int pfb_dilate_points(const double *points, const double *radii, int n, int nz, int ny, int nx, double *out,
                      void *stream);

/* helpers.determine_core_indices(mask) (helpers.py:26-34): number of binary erosions (scipy's
 * 6-neighbour cross, zero border) each voxel of mask > 0 survives, as doubles.  scratch = 2*nz*ny*nx + 16
 * device bytes.  Synchronises the stream once per erosion level. */
int pfb_core_indices(const double *mask, int nz, int ny, int nx, double *core, uint8_t *scratch, void *stream);

/* ---- image pyramid: the array operations of scripts/__init__.py:93-103 ---- */

/* scipy.ndimage.gaussian_filter(in, sigma, mode='constant') as volume.lower_resolution calls it
 * (volume.py:129-141).  weights = right half of scipy's normalised kernel (radius + 1 doubles, DEVICE);
 * tmp = scratch of the grid's size.  FP64-identical to scipy. */
int pfb_gaussian_filter(const double *in, double *out, double *tmp, int nz, int ny, int nx, const double *weights,
                        int radius, void *stream);

/* scipy.ndimage.zoom(in, factor, order=1) as volume.resample calls it (volume.py:66-72): trilinear
 * interpolation onto an oz*oy*ox grid, x_in = x_out (n_in - 1) / (n_out - 1). */
int pfb_zoom_linear(const double *in, int nz, int ny, int nx, double *out, int oz, int oy, int ox, void *stream);

/* Test hook: one FFT building block of the fused kernels (csrc/fft_core.cuh) on caller data, so that each of them
 * can be pinned against a reference FFT on its own -- the role clFFT's own test-suite plays for the reference's
 * grfftn_builder plans (powerfitter.py:605-638).  in/out: DEVICE float4[count][lanes*e].
 * kind 0: packed pencil (two independent complex sequences per element, lanes*e points);
 * kind 1 / 2: row transforms of ONE sequence of 2*lanes*e points, adjacent-in -> split-out / split-in -> adjacent-out;
 * kind 3: scalar pencil (lanes = 8).  Kernel exp(+2 pi i n k / N), un-normalised. */
int pfb_pencil_fft(int kind, int lanes, int e, const float *in, float *out, int count, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* POWERFIT_B200_H */
