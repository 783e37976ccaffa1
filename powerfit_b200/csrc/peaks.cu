// Device half of the solution extraction that follows a search:
// the reference's Analyzer._watershed (/root/reference/src/powerfit_em/analyzer.py:80-95) takes
// the maximum of the LCC grid and then labels {lcc >= cutoff} for `steps` cutoffs between the
// maximum and half of it -- five full passes of scipy.ndimage.label + maximum_position over
// every voxel on the host.  Only voxels above the LOWEST cutoff can ever be part of a
// labelled feature, and they are a tiny fraction of the grid, so the device does the two
// grid-sized operations (max reduction, threshold + stream compaction) and hands the host a
// short (index, value) list; the connected-component bookkeeping on that list is in
// powerfit_b200/analyzer.py.  No plan is needed: the calls work on any float32 device grid of
// the current device.
#include "common.cuh"

#include <algorithm>

namespace pfb {

__global__ void lcc_max_kernel(const float *__restrict__ lcc, long n, int *__restrict__ key) {
    int best = orderable_f32(0xFF800000u);          // -inf
    bool nan = false;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float v = lcc[i];
        if (v != v) nan = true;
        best = max(best, orderable_f32(__float_as_uint(v)));
    }
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (__any_sync(0xffffffffu, nan)) best = 0x7FFFFFFF;   // numpy's max() propagates NaN
    if ((threadIdx.x & 31) == 0) atomicMax(key, best);
}

__global__ void lcc_max_finish_kernel(const int *__restrict__ key, float *__restrict__ out) {
    const int k = *key;
    *out = k == 0x7FFFFFFF ? __int_as_float(0x7FC00000) : __uint_as_float(unorderable_f32(k));
}

// warp-aggregated stream compaction of {i : lcc[i] >= cutoff}; order of the output is arbitrary
__global__ void peak_candidates_kernel(const float *__restrict__ lcc, long n, float cutoff, int cap,
                                       int *__restrict__ idx, float *__restrict__ val, int *__restrict__ count) {
    const long stride = (long)gridDim.x * blockDim.x;
    const long n_round = (n + 31) / 32 * 32;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const float v = i < n ? lcc[i] : 0.f;
        const bool keep = i < n && v >= cutoff;
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        if (ballot == 0) continue;
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == 0) base = atomicAdd(count, __popc(ballot));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) {
            const int slot = base + __popc(ballot & ((1u << lane) - 1u));
            if (slot < cap) { idx[slot] = (int)i; val[slot] = v; }
        }
    }
}

}  // namespace pfb

using namespace pfb;

extern "C" {

int pfb_lcc_max(const float *lcc, int64_t n, float *max_out, int32_t *scratch, void *stream) {
    PFB_REQUIRE(lcc && max_out && scratch && n > 0, "pfb_lcc_max: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    const int init = orderable_f32(0xFF800000u);
    PFB_CUDA(cudaMemcpyAsync(scratch, &init, sizeof(int), cudaMemcpyHostToDevice, s));
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    lcc_max_kernel<<<blocks, 256, 0, s>>>(lcc, (long)n, scratch);
    lcc_max_finish_kernel<<<1, 1, 0, s>>>(scratch, max_out);
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

int pfb_peak_candidates(const float *lcc, int64_t n, float cutoff, int32_t cap, int32_t *idx, float *val,
                        int32_t *count, void *stream) {
    PFB_REQUIRE(lcc && idx && val && count && n > 0 && cap > 0, "pfb_peak_candidates: bad argument");
    PFB_REQUIRE(n <= 0x7FFFFFFFLL, "pfb_peak_candidates: grid too large for 32-bit voxel indices");
    cudaStream_t s = (cudaStream_t)stream;
    PFB_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t), s));
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    peak_candidates_kernel<<<blocks, 256, 0, s>>>(lcc, (long)n, cutoff, cap, idx, val, count);
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

}  // extern "C"
