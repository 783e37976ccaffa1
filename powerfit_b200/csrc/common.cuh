// Shared declarations of the powerfit_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/powerfit_b200.h"

namespace pfb {

void set_error(const std::string &msg);

#define PFB_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            pfb::set_error(std::string(#call) + ": " + cudaGetErrorString(e_));             \
            return PFB_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

#define PFB_REQUIRE(cond, msg)                                                              \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            pfb::set_error(msg);                                                            \
            return PFB_ERR_INVALID;                                                         \
        }                                                                                   \
    } while (0)

// ---------------------------------------------------------------- packed best key
// key = (orderable(lcc) << 32) | (0xFFFFFFFF - rot) as a SIGNED 64-bit integer:
// orderable() maps the float's bits to an int32 with the same order (-inf .. -0.0 <
// +0.0 .. +inf), so a plain signed max means "greater LCC wins, equal LCC -> lower
// rotation index wins" -- the reference's sequential strict-'>' rule
// (powerfitter.py:327-330, 146-163) -- and the multi-GPU merge is one integer MAX
// all-reduce.  pack(+0.0f, 0) is the initial value ("LCC 0, rotation 0").
__host__ __device__ inline int32_t orderable_f32(uint32_t u) {
    const int32_t s = (int32_t)u;
    return s ^ ((s >> 31) & 0x7FFFFFFF);
}
__host__ __device__ inline uint32_t unorderable_f32(int32_t k) {
    return (uint32_t)(k ^ ((k >> 31) & 0x7FFFFFFF));
}
__host__ __device__ inline int64_t pack_best(uint32_t lcc_bits, uint32_t rot) {
    return (int64_t)(((uint64_t)(uint32_t)orderable_f32(lcc_bits) << 32) | (uint64_t)(0xFFFFFFFFu - rot));
}
constexpr int64_t kBestInit = 0x00000000FFFFFFFFll;  // (+0.0f, rot 0)

// ---------------------------------------------------------------- 1-D FFT description
constexpr int kMaxPasses = 16;
struct Fft1D {
    int n = 0;
    int npass = 0;
    int radix[kMaxPasses] = {0};
};
bool factorize(int n, Fft1D *out);   // radices from {8,4,2,3,5,7}; false if not smooth

struct Plan;
constexpr size_t kPrepScratchBytes = 32768;

// kernel classes for the optional per-kernel timing (pfb_profile*)
enum KernelClass { KC_ROTATE = 0, KC_FFT_X, KC_FFT_Y, KC_FFT_Z, KC_MULTIPLY, KC_LCC, KC_FUSED_A, KC_FUSED_B,
                   KC_FUSED_C, KC_OTHER, KC_COUNT };
const char *kernel_class_name(int cls);
struct ProfRec { cudaEvent_t a, b; int cls; };

// generic path (any 2.3.5.7-smooth shape)
int launch_rotate_pack(Plan *p, const double *rot_dev, int first, int count, cudaStream_t s);
int launch_fft_axis(Plan *p, float2 *vols, int nvol, int axis, cudaStream_t s);
int fft_generic_init();
int launch_multiply(Plan *p, int npairs, cudaStream_t s);
int launch_lcc_best(Plan *p, int first_rot_index, int count, int64_t *best, cudaStream_t s);
int launch_rotate_plain(Plan *p, const float *grid, const double *rot_dev, int R, int nearest,
                        float *out, cudaStream_t s);
int launch_lcc_single(Plan *p, const float *gcc, const float *ave, const float *ave2, float norm,
                      int rot_index, int64_t *best, cudaStream_t s);
int launch_best_init(Plan *p, int64_t *best, cudaStream_t s);
int launch_unpack(Plan *p, const int64_t *best, float *lcc, int32_t *rot, cudaStream_t s);
int launch_merge(Plan *p, int64_t *dst, const int64_t *src, cudaStream_t s);
int launch_target_spectra(Plan *p, const float *target, cudaStream_t s);

// one-time FP64 input preparation (prep.cu)
int prep_target(Plan *p, const double *target, int laplace, float *f_out, uint8_t *lcc_mask_out, cudaStream_t s);
int prep_template(Plan *p, const double *tmpl, const double *mask, int laplace, float *t_out, float *m_out,
                  double *norm_factor, int *mask_is_binary, cudaStream_t s);

// fused path (fused.cu)
bool fused_supported(int nz, int ny, int nx);
int fused_init(Plan *p);
int fused_prepare_target(Plan *p, cudaStream_t s);
int fused_prepare_template(Plan *p, cudaStream_t s);
int fused_scan(Plan *p, int R, int rot_index_offset, int64_t *best, cudaStream_t s);
int fused_a(Plan *p, int first, int count, cudaStream_t s);
int fused_b(Plan *p, int count, float2 *X2, cudaStream_t s);
int fused_c(Plan *p, int first, int count, int rot_index_offset, int64_t *best, const float2 *X2, cudaStream_t s);
int launch_fused_a(Plan *p, int first, int count, cudaStream_t s);

// class-decimated fused path (fused_cls.cu): N = 256, optionally N = 128
int cls_init(Plan *p);
int cls_prepare_target(Plan *p, cudaStream_t s);
int cls_a(Plan *p, int first, int count, cudaStream_t s);
int cls_b(Plan *p, int count, float2 *X2, cudaStream_t s);
int cls_c(Plan *p, int first, int count, int rot_index_offset, int64_t *best, const float2 *X2, cudaStream_t s);

// Per-template state of a plan.  A plan owns the map (FT(f), FT(f^2), lcc_mask) and the work buffers; the
// template-dependent buffers and parameters live in slots so that several templates can be searched against
// the same map (BASELINE configs[4]: a batch of sub-unit templates).  The active slot is mirrored in the Plan
// fields of the same names; pfb_select_template swaps pointers, nothing is copied.
struct TemplateSlot {
    float *tmpl = nullptr, *mask = nullptr;
    float4 *tmplq = nullptr;
    float norm_factor = 0.f;
    int nsig = 2, rs = 0, rs2 = 0;
    unsigned ymask = 0, nmask = 0;
    bool have_template = false;
};

struct Plan {
    int nz = 0, ny = 0, nx = 0, rmax = 0, device = 0;
    long V = 0;
    int batch = 0;            // rotations per pass (even)
    int nsig = 2;             // forward volumes per rotation pair: 2 (binary mask) or 3
    float norm_factor = 0.f;
    bool have_target = false, have_template = false;
    Fft1D fx, fy, fz;
    float2 *tw[3] = {nullptr, nullptr, nullptr};   // exp(+2 pi i k/n) per axis (x,y,z)
    float *tmpl = nullptr, *mask = nullptr;        // prepared template, mask (float32)
    uint8_t *lcc_mask = nullptr;
    float2 *F = nullptr, *F2 = nullptr;            // conj(P f)/V, conj(P f^2)/V, full spectra
    // fused path (axes of 64/128, cubic 192/256): map spectra transposed to [kx][ky][kz], template support box
    bool fused = false;
    float2 *Fq = nullptr, *F2q = nullptr;
    float2 *twdN = nullptr, *twdM = nullptr;       // packed-pencil twiddle tables [k1][t] (fft_core.cuh): z columns, y rows
    float2 *twdX = nullptr;                        // the same for kernel C's x pencils (8 lanes x nx/8)
    float4 *tmplq = nullptr;                       // template corner table for kernel A's gather
    uint32_t *mbits = nullptr;                     // lcc_mask bit-packed in kernel C's lane layout
    CUtensorMap tmapC;                             // kernel C's view of the X2 work buffer (tma.cuh)
    const void *tmapC_base = nullptr;              // buffer the map was encoded for
    CUtensorMap tmapB;                             // kernel B's view of the X1 work buffer (staging copies)
    const void *tmapB_base = nullptr;
    int tmapB_rs = -1, tmapB_nsig = -1;
    int rs = 0, rs2 = 0;
    unsigned ymask = 0;
    // class-decimated variant of kernels B and C (fused_cls.cu)
    bool cls = false;
    unsigned nmask = 0;                            // tiles of 4 folded rows n (n = y mod 64) that meet the support
    float2 *cls_twN = nullptr, *cls_twM = nullptr, *cls_twh = nullptr;
    float4 *cls_fold = nullptr;
    float2 *A = nullptr;                           // forward work: [batch/2][3][V]
    float2 *B = nullptr;                           // product / inverse work: [batch/2][3][V]
    double *rot_dev = nullptr;
    long rot_cap = 0;
    int64_t *best_scratch = nullptr;              // used by pfb_search_host
    void *prep_scratch = nullptr;                 // partial sums of the FP64 template preparation (prep.cu)
    std::vector<TemplateSlot> slots;              // slots[cur] is stale while it is the active one
    int cur = 0;
    // side stream of the fused scan: kernel A of the next batch runs next to kernel C of the current one
    cudaStream_t side = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    unsigned long long launches = 0;
    int sm_count = 148;
    // per-kernel-class timing (off by default; events around every launch when on)
    bool profile = false;
    std::vector<ProfRec> prof;
    double prof_ms[KC_COUNT] = {0};
    long prof_n[KC_COUNT] = {0};
};

// RAII: counts the launch and, when profiling, brackets it with events on stream s.
struct LaunchScope {
    Plan *p; cudaStream_t s; cudaEvent_t b = nullptr;
    LaunchScope(Plan *plan, int cls, cudaStream_t stream) : p(plan), s(stream) {
        p->launches++;
        if (p->profile) {
            ProfRec r; r.cls = cls;
            cudaEventCreate(&r.a); cudaEventCreate(&r.b);
            cudaEventRecord(r.a, s);
            b = r.b;
            p->prof.push_back(r);
        }
    }
    ~LaunchScope() { if (b) cudaEventRecord(b, s); }
};

}  // namespace pfb

struct pfb_plan {
    pfb::Plan p;
};
