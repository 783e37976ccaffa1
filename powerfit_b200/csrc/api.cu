// C ABI of powerfit_b200 (see include/powerfit_b200.h for the contract and the
// reference interfaces each entry point replaces).
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace pfb {

static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }

static int upload_twiddles(int n, float2 **out) {
    std::vector<float2> h(n);
    for (int k = 0; k < n; ++k) {
        const double a = 2.0 * M_PI * (double)k / (double)n;
        h[k] = make_float2((float)cos(a), (float)sin(a));
    }
    PFB_CUDA(cudaMalloc(out, sizeof(float2) * n));
    PFB_CUDA(cudaMemcpy(*out, h.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
    return PFB_OK;
}

static int ensure_rot_capacity(Plan *p, long R) {
    if (R <= p->rot_cap) return PFB_OK;
    if (p->rot_dev) PFB_CUDA(cudaFree(p->rot_dev));
    p->rot_dev = nullptr;
    long cap = std::max<long>(R, 1024);
    PFB_CUDA(cudaMalloc(&p->rot_dev, sizeof(double) * 9 * cap));
    p->rot_cap = cap;
    return PFB_OK;
}

const char *kernel_class_name(int cls) {
    static const char *names[KC_COUNT] = {"rotate", "fft_x", "fft_y", "fft_z", "multiply", "lcc_best",
                                          "fused_rotate_fftx", "fused_fftyz_mul", "fused_ifftx_lcc", "other"};
    return (cls >= 0 && cls < KC_COUNT) ? names[cls] : "?";
}

static void profile_collect(Plan *p) {
    for (ProfRec &r : p->prof) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            p->prof_ms[r.cls] += ms;
            p->prof_n[r.cls] += 1;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    p->prof.clear();
}

static void slot_save(Plan *p, TemplateSlot &t) {
    t.tmpl = p->tmpl; t.mask = p->mask; t.tmplq = p->tmplq; t.norm_factor = p->norm_factor; t.nsig = p->nsig;
    t.rs = p->rs; t.rs2 = p->rs2; t.ymask = p->ymask; t.nmask = p->nmask; t.have_template = p->have_template;
}
static void slot_load(Plan *p, const TemplateSlot &t) {
    p->tmpl = t.tmpl; p->mask = t.mask; p->tmplq = t.tmplq; p->norm_factor = t.norm_factor; p->nsig = t.nsig;
    p->rs = t.rs; p->rs2 = t.rs2; p->ymask = t.ymask; p->nmask = t.nmask; p->have_template = t.have_template;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// one batch of the generic (any-shape) pipeline
static int scan_batch_generic(Plan *p, int first, int count, int rot_index_offset, int64_t *best, cudaStream_t s) {
    const int npairs = (count + 1) / 2;
    int rc;
    if ((rc = launch_rotate_pack(p, p->rot_dev, first, count, s))) return rc;
    for (int axis = 0; axis < 3; ++axis)
        if ((rc = launch_fft_axis(p, p->A, npairs * p->nsig, axis, s))) return rc;
    if ((rc = launch_multiply(p, npairs, s))) return rc;
    for (int axis = 2; axis >= 0; --axis)
        if ((rc = launch_fft_axis(p, p->B, npairs * 3, axis, s))) return rc;
    return launch_lcc_best(p, rot_index_offset + first, count, best, s);
}

}  // namespace pfb

using namespace pfb;

extern "C" {

const char *pfb_version(void) { return "powerfit_b200 0.1.0 sm_100a"; }
const char *pfb_last_error(void) { return g_error.c_str(); }

int pfb_plan_create(int nz, int ny, int nx, int max_batch, int device, pfb_plan **out) {
    PFB_REQUIRE(out != nullptr, "pfb_plan_create: out is NULL");
    *out = nullptr;
    PFB_REQUIRE(nz >= 2 && ny >= 2 && nx >= 2, "pfb_plan_create: every axis must be >= 2");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("pfb_plan_create: no CUDA device visible (this library has no CPU fallback)");
        return PFB_ERR_NODEVICE;
    }
    PFB_REQUIRE(device >= 0 && device < ndev, "pfb_plan_create: bad device ordinal");
    cudaDeviceProp prop;
    PFB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error(std::string("pfb_plan_create: built for sm_100a only, device is ") + prop.name);
        return PFB_ERR_NODEVICE;
    }
    pfb_plan *h = new pfb_plan();
    Plan *p = &h->p;
    p->nz = nz; p->ny = ny; p->nx = nx;
    p->V = (long)nz * ny * nx;
    p->rmax = std::min(nz, std::min(ny, nx)) / 2;
    p->device = device;
    p->sm_count = prop.multiProcessorCount;
    if (!factorize(nx, &p->fx) || !factorize(ny, &p->fy) || !factorize(nz, &p->fz)) {
        delete h;
        set_error("pfb_plan_create: axis lengths must be products of 2, 3, 5 and 7");
        return PFB_ERR_UNSUPPORTED;
    }
    if (std::max(nz, std::max(ny, nx)) > 1024) {
        delete h;
        set_error("pfb_plan_create: axis length above 1024 is not supported");
        return PFB_ERR_UNSUPPORTED;
    }
    DeviceGuard guard(device);
    if (!guard.ok) { delete h; set_error("cudaSetDevice failed"); return PFB_ERR_CUDA; }
    if (max_batch <= 0) {
        // default: as many rotations in flight as fit ~32 GB of work buffers (of 180 GB), at most 512.  Measured at
        // 128^3: 64 -> 47.4k, 128 -> 48.5k, 256 -> 54.3k, 512 -> 54.8k rotations/s; at 64^3: 64 -> 269k, 256 -> 298k,
        // 512 -> 311k (longer launches amortise tails, prologues and the per-CTA start-up of the persistent kernel)
        const long per_pair = 6L * p->V * (long)sizeof(float2);
        long pairs = (32768L << 20) / per_pair;
        pairs = std::max(1L, std::min(256L, pairs));
        max_batch = (int)(2 * pairs);
    }
    if (const char *e = getenv("PFB_BATCH")) max_batch = std::max(1, atoi(e));
    p->batch = (max_batch + 1) & ~1;
    int rc = PFB_OK;
    auto fail = [&](int code) { pfb_plan_destroy(h); return code; };
    if ((rc = fft_generic_init())) return fail(rc);
    if ((rc = upload_twiddles(nx, &p->tw[0]))) return fail(rc);
    if ((rc = upload_twiddles(ny, &p->tw[1]))) return fail(rc);
    if ((rc = upload_twiddles(nz, &p->tw[2]))) return fail(rc);
#define PFB_ALLOC(ptr, bytes)                                                           \
    do {                                                                                \
        cudaError_t e_ = cudaMalloc(&(ptr), (bytes));                                   \
        if (e_ != cudaSuccess) {                                                        \
            set_error(std::string("cudaMalloc " #ptr ": ") + cudaGetErrorString(e_));  \
            return fail(PFB_ERR_CUDA);                                                  \
        }                                                                               \
    } while (0)
    PFB_ALLOC(p->tmpl, sizeof(float) * p->V);
    PFB_ALLOC(p->mask, sizeof(float) * p->V);
    PFB_ALLOC(p->lcc_mask, p->V);
    PFB_ALLOC(p->F, sizeof(float2) * p->V);
    PFB_ALLOC(p->F2, sizeof(float2) * p->V);
    // work buffers: 3 V complex per rotation pair each; if the device is short of memory (other plans, other
    // processes) the batch is halved until they fit
    for (;;) {
        const size_t bytes = sizeof(float2) * (size_t)p->V * 3 * (size_t)(p->batch / 2);
        cudaError_t ea = cudaMalloc(&p->A, bytes), eb = ea == cudaSuccess ? cudaMalloc(&p->B, bytes) : ea;
        if (ea == cudaSuccess && eb == cudaSuccess) break;
        if (p->A) { cudaFree(p->A); p->A = nullptr; }
        if (p->B) { cudaFree(p->B); p->B = nullptr; }
        cudaGetLastError();
        if (p->batch <= 2) {
            set_error(std::string("cudaMalloc work buffers: ") + cudaGetErrorString(ea != cudaSuccess ? ea : eb));
            return fail(PFB_ERR_CUDA);
        }
        p->batch = ((p->batch / 2) + 1) & ~1;
    }
    PFB_ALLOC(p->best_scratch, sizeof(int64_t) * p->V);
    PFB_ALLOC(p->prep_scratch, kPrepScratchBytes);
    p->fused = fused_supported(nz, ny, nx);
    if (const char *e = getenv("PFB_FUSED")) p->fused = p->fused && atoi(e) != 0;
    if (p->fused) {
        p->cls = nx == 256 || nx == 192;
        PFB_ALLOC(p->Fq, sizeof(float2) * p->V);
        PFB_ALLOC(p->F2q, sizeof(float2) * p->V);
        PFB_ALLOC(p->mbits, sizeof(uint32_t) * (size_t)nz * ny * 16);
        PFB_ALLOC(p->tmplq, sizeof(float4) * p->V);
    }
#undef PFB_ALLOC
    p->slots.resize(1);
    if ((rc = ensure_rot_capacity(p, 1024))) return fail(rc);
    if (p->fused && (rc = fused_init(p))) return fail(rc);
    if (p->fused) {
        // the side stream outranks the caller's stream: kernel A's CTAs are placed as soon as kernel C frees slots
        int lo_prio = 0, hi_prio = 0;
        cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio);
        if (cudaStreamCreateWithPriority(&p->side, cudaStreamNonBlocking, hi_prio) != cudaSuccess ||
            cudaEventCreateWithFlags(&p->ev_a, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&p->ev_b, cudaEventDisableTiming) != cudaSuccess) {
            set_error("pfb_plan_create: side stream / events");
            return fail(PFB_ERR_CUDA);
        }
    }
    *out = h;
    return PFB_OK;
}

int pfb_plan_destroy(pfb_plan *h) {
    if (!h) return PFB_OK;
    Plan *p = &h->p;
    DeviceGuard guard(p->device);
    if (!p->slots.empty()) slot_save(p, p->slots[p->cur]);
    for (size_t i = 0; i < p->slots.size(); ++i) {
        if ((int)i == p->cur) continue;           // the active slot's buffers are freed through the Plan fields
        if (p->slots[i].tmpl) cudaFree(p->slots[i].tmpl);
        if (p->slots[i].mask) cudaFree(p->slots[i].mask);
        if (p->slots[i].tmplq) cudaFree(p->slots[i].tmplq);
    }
    void *ptrs[] = {p->tw[0], p->tw[1], p->tw[2], p->tmpl, p->mask, p->lcc_mask, p->F, p->F2,
                    p->A, p->B, p->rot_dev, p->best_scratch, p->prep_scratch, p->Fq, p->F2q, p->twdN, p->twdM, p->twdX, p->mbits, p->tmplq,
                    p->cls_twN, p->cls_twM, p->cls_twh, p->cls_fold};
    for (void *q : ptrs)
        if (q) cudaFree(q);
    if (p->side) { cudaStreamSynchronize(p->side); cudaStreamDestroy(p->side); }
    if (p->ev_a) cudaEventDestroy(p->ev_a);
    if (p->ev_b) cudaEventDestroy(p->ev_b);
    delete h;
    return PFB_OK;
}

int pfb_plan_info(const pfb_plan *h, int what, int64_t *value) {
    PFB_REQUIRE(h && value, "pfb_plan_info: NULL argument");
    const Plan *p = &h->p;
    switch (what) {
        case 0: *value = p->nz; break;
        case 1: *value = p->ny; break;
        case 2: *value = p->nx; break;
        case 3: *value = p->rmax; break;
        case 4: *value = p->batch; break;
        case 5: *value = p->device; break;
        case 6: *value = p->fused ? 1 : 0; break;
        case 8: *value = p->rs; break;
        case 9: *value = p->cls ? 1 : 0; break;
        case 7: *value = (int64_t)(p->launches & 0x7FFFFFFF); break;
        default: set_error("pfb_plan_info: unknown field"); return PFB_ERR_INVALID;
    }
    return PFB_OK;
}

int pfb_template_slots(pfb_plan *h, int nslots) {
    PFB_REQUIRE(h, "pfb_template_slots: NULL plan");
    PFB_REQUIRE(nslots >= 1 && nslots <= 64, "pfb_template_slots: between 1 and 64 slots");
    Plan *p = &h->p;
    DeviceGuard guard(p->device);
    while ((int)p->slots.size() < nslots) {
        TemplateSlot t;
        if (cudaMalloc(&t.tmpl, sizeof(float) * p->V) != cudaSuccess || cudaMalloc(&t.mask, sizeof(float) * p->V) != cudaSuccess ||
            (p->fused && cudaMalloc(&t.tmplq, sizeof(float4) * p->V) != cudaSuccess)) {
            if (t.tmpl) cudaFree(t.tmpl);
            if (t.mask) cudaFree(t.mask);
            if (t.tmplq) cudaFree(t.tmplq);
            cudaGetLastError();
            set_error("pfb_template_slots: out of device memory");
            return PFB_ERR_CUDA;
        }
        p->slots.push_back(t);
    }
    return PFB_OK;
}

int pfb_select_template(pfb_plan *h, int slot) {
    PFB_REQUIRE(h, "pfb_select_template: NULL plan");
    Plan *p = &h->p;
    PFB_REQUIRE(slot >= 0 && slot < (int)p->slots.size(), "pfb_select_template: no such slot (pfb_template_slots first)");
    if (slot == p->cur) return PFB_OK;
    slot_save(p, p->slots[p->cur]);
    slot_load(p, p->slots[slot]);
    p->cur = slot;
    return PFB_OK;
}

int pfb_profile(pfb_plan *h, int enable) {
    PFB_REQUIRE(h, "pfb_profile: NULL plan");
    Plan *p = &h->p;
    DeviceGuard guard(p->device);
    if (enable) {
        profile_collect(p);
        for (int i = 0; i < KC_COUNT; ++i) { p->prof_ms[i] = 0; p->prof_n[i] = 0; }
        p->profile = true;
    } else {
        p->profile = false;
        profile_collect(p);
    }
    return PFB_OK;
}

int pfb_profile_read(pfb_plan *h, int cls, double *ms, int64_t *launches, const char **name) {
    PFB_REQUIRE(h && ms && launches, "pfb_profile_read: NULL argument");
    if (cls < 0 || cls >= KC_COUNT) return PFB_ERR_INVALID;
    Plan *p = &h->p;
    DeviceGuard guard(p->device);
    profile_collect(p);
    *ms = p->prof_ms[cls];
    *launches = p->prof_n[cls];
    if (name) *name = kernel_class_name(cls);
    return PFB_OK;
}

int pfb_set_target(pfb_plan *h, const float *target, const uint8_t *lcc_mask, void *stream) {
    PFB_REQUIRE(h && target && lcc_mask, "pfb_set_target: NULL argument");
    Plan *p = &h->p;
    DeviceGuard guard(p->device);
    cudaStream_t s = (cudaStream_t)stream;
    PFB_CUDA(cudaMemcpyAsync(p->lcc_mask, lcc_mask, p->V, cudaMemcpyDeviceToDevice, s));
    int rc = launch_target_spectra(p, target, s);
    if (rc) return rc;
    if (p->fused && (rc = fused_prepare_target(p, s))) return rc;
    p->have_target = true;
    return PFB_OK;
}

int pfb_set_template(pfb_plan *h, const float *tmpl, const float *mask, float norm_factor, int mask_is_binary,
                     void *stream) {
    PFB_REQUIRE(h && tmpl && mask, "pfb_set_template: NULL argument");
    PFB_REQUIRE(norm_factor > 0.f, "pfb_set_template: norm_factor must be positive (zero-filled mask?)");
    Plan *p = &h->p;
    DeviceGuard guard(p->device);
    cudaStream_t s = (cudaStream_t)stream;
    PFB_CUDA(cudaMemcpyAsync(p->tmpl, tmpl, sizeof(float) * p->V, cudaMemcpyDeviceToDevice, s));
    PFB_CUDA(cudaMemcpyAsync(p->mask, mask, sizeof(float) * p->V, cudaMemcpyDeviceToDevice, s));
    p->norm_factor = norm_factor;
    p->nsig = mask_is_binary ? 2 : 3;
    if (p->fused) {
        int rc = fused_prepare_template(p, s);
        if (rc) return rc;
    }
    p->have_template = true;
    return PFB_OK;
}

int pfb_prepare_target(pfb_plan *h, const double *target, int laplace, float *f_out, uint8_t *lcc_mask_out,
                       void *stream) {
    PFB_REQUIRE(h && target && f_out && lcc_mask_out, "pfb_prepare_target: NULL argument");
    Plan *p = &h->p;
    {
        DeviceGuard guard(p->device);
        int rc = prep_target(p, target, laplace, f_out, lcc_mask_out, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return pfb_set_target(h, f_out, lcc_mask_out, stream);
}

int pfb_prepare_template(pfb_plan *h, const double *tmpl, const double *mask, int laplace, float *t_out, float *m_out,
                         double *norm_factor, int *mask_is_binary, void *stream) {
    PFB_REQUIRE(h && tmpl && mask && t_out && m_out && norm_factor && mask_is_binary,
                "pfb_prepare_template: NULL argument");
    Plan *p = &h->p;
    {
        DeviceGuard guard(p->device);
        int rc = prep_template(p, tmpl, mask, laplace, t_out, m_out, norm_factor, mask_is_binary, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return pfb_set_template(h, t_out, m_out, (float)*norm_factor, *mask_is_binary, stream);
}

int pfb_best_init(pfb_plan *h, int64_t *best, void *stream) {
    PFB_REQUIRE(h && best, "pfb_best_init: NULL argument");
    DeviceGuard guard(h->p.device);
    return launch_best_init(&h->p, best, (cudaStream_t)stream);
}

int pfb_scan(pfb_plan *h, const double *rotmats_host, int R, int rot_index_offset, int64_t *best, void *stream) {
    PFB_REQUIRE(h && best, "pfb_scan: NULL argument");
    Plan *p = &h->p;
    PFB_REQUIRE(p->have_target && p->have_template, "pfb_scan: first set the target, template and mask");
    PFB_REQUIRE(R >= 0 && rot_index_offset >= 0, "pfb_scan: negative rotation count or offset");
    if (R == 0) return PFB_OK;
    PFB_REQUIRE(rotmats_host != nullptr, "pfb_scan: rotmats is NULL");
    PFB_REQUIRE((long)rot_index_offset + R <= 0x7FFFFFFFL, "pfb_scan: rotation index overflow");
    DeviceGuard guard(p->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = ensure_rot_capacity(p, R);
    if (rc) return rc;
    PFB_CUDA(cudaMemcpyAsync(p->rot_dev, rotmats_host, sizeof(double) * 9 * R, cudaMemcpyHostToDevice, s));
    if (p->fused) return fused_scan(p, R, rot_index_offset, best, s);
    for (int first = 0; first < R; first += p->batch) {
        const int count = std::min(p->batch, R - first);
        if ((rc = scan_batch_generic(p, first, count, rot_index_offset, best, s))) return rc;
    }
    return PFB_OK;
}

int pfb_unpack(pfb_plan *h, const int64_t *best, float *lcc, int32_t *rot, void *stream) {
    PFB_REQUIRE(h && best && lcc && rot, "pfb_unpack: NULL argument");
    DeviceGuard guard(h->p.device);
    return launch_unpack(&h->p, best, lcc, rot, (cudaStream_t)stream);
}

int pfb_merge_best(pfb_plan *h, int64_t *dst, const int64_t *src, void *stream) {
    PFB_REQUIRE(h && dst && src, "pfb_merge_best: NULL argument");
    DeviceGuard guard(h->p.device);
    return launch_merge(&h->p, dst, src, (cudaStream_t)stream);
}

int pfb_rotate(pfb_plan *h, const float *grid, const double *rotmats_host, int R, int nearest, float *out,
               void *stream) {
    PFB_REQUIRE(h && grid && rotmats_host && out && R > 0, "pfb_rotate: bad argument");
    Plan *p = &h->p;
    DeviceGuard guard(p->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = ensure_rot_capacity(p, R);
    if (rc) return rc;
    PFB_CUDA(cudaMemcpyAsync(p->rot_dev, rotmats_host, sizeof(double) * 9 * R, cudaMemcpyHostToDevice, s));
    return launch_rotate_plain(p, grid, p->rot_dev, R, nearest, out, s);
}

int pfb_fft3_c2c(pfb_plan *h, float *vols, int nvol, void *stream) {
    PFB_REQUIRE(h && vols && nvol > 0, "pfb_fft3_c2c: bad argument");
    Plan *p = &h->p;
    DeviceGuard guard(p->device);
    for (int axis = 0; axis < 3; ++axis) {
        int rc = launch_fft_axis(p, (float2 *)vols, nvol, axis, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return PFB_OK;
}

int pfb_lcc_take_best(pfb_plan *h, const float *gcc, const float *ave, const float *ave2, float norm_factor,
                      int rot_index, int64_t *best, void *stream) {
    PFB_REQUIRE(h && gcc && ave && ave2 && best, "pfb_lcc_take_best: NULL argument");
    PFB_REQUIRE(h->p.have_target, "pfb_lcc_take_best: set the target (lcc_mask) first");
    DeviceGuard guard(h->p.device);
    return launch_lcc_single(&h->p, gcc, ave, ave2, norm_factor, rot_index, best, (cudaStream_t)stream);
}

int pfb_search_host(pfb_plan *h, const float *target, const uint8_t *lcc_mask, const float *tmpl,
                    const float *mask, float norm_factor, int mask_is_binary, const double *rotmats, int R,
                    int rot_index_offset, float *lcc, int32_t *rot) {
    PFB_REQUIRE(h && target && lcc_mask && tmpl && mask && lcc && rot, "pfb_search_host: NULL argument");
    Plan *p = &h->p;
    DeviceGuard guard(p->device);
    cudaStream_t s = nullptr;
    // stage through the work buffers: A holds target/template/mask, B the outputs
    float *d_target = (float *)p->A, *d_tmpl = d_target + p->V, *d_mask = d_tmpl + p->V;
    uint8_t *d_lm = (uint8_t *)(d_mask + p->V);
    PFB_CUDA(cudaMemcpyAsync(d_target, target, sizeof(float) * p->V, cudaMemcpyHostToDevice, s));
    PFB_CUDA(cudaMemcpyAsync(d_tmpl, tmpl, sizeof(float) * p->V, cudaMemcpyHostToDevice, s));
    PFB_CUDA(cudaMemcpyAsync(d_mask, mask, sizeof(float) * p->V, cudaMemcpyHostToDevice, s));
    PFB_CUDA(cudaMemcpyAsync(d_lm, lcc_mask, p->V, cudaMemcpyHostToDevice, s));
    int rc;
    if ((rc = pfb_set_target(h, d_target, d_lm, s))) return rc;
    if ((rc = pfb_set_template(h, d_tmpl, d_mask, norm_factor, mask_is_binary, s))) return rc;
    if ((rc = pfb_best_init(h, p->best_scratch, s))) return rc;
    if ((rc = pfb_scan(h, rotmats, R, rot_index_offset, p->best_scratch, s))) return rc;
    float *d_lcc = (float *)p->B;
    int32_t *d_rot = (int32_t *)(d_lcc + p->V);
    if ((rc = pfb_unpack(h, p->best_scratch, d_lcc, d_rot, s))) return rc;
    PFB_CUDA(cudaMemcpyAsync(lcc, d_lcc, sizeof(float) * p->V, cudaMemcpyDeviceToHost, s));
    PFB_CUDA(cudaMemcpyAsync(rot, d_rot, sizeof(int32_t) * p->V, cudaMemcpyDeviceToHost, s));
    PFB_CUDA(cudaStreamSynchronize(s));
    return PFB_OK;
}

}  // extern "C"
