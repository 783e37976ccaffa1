// Fused three-kernel search pipeline for grids whose axes are 32, 64, 96 or 128 voxels long (any mix: the x
// pencils of kernels A and C are templated on nx, kernel B's plane on (nz, ny); 64^3 and 128^3 are the tuned cases).
//
//   A  fused_rotate_fftx  : rotate template+mask (a PAIR of rotations packed as one complex
//                           signal: re = rotation a, im = rotation b), transform along x,
//                           write X1[pair][sig][z][kx][y]            (replaces K1/K2 + 1/3 of K5)
//   B  fused_fftyz_mul    : one (kx, output volume, pair) plane per CTA, 128 KB of shared
//                           memory: forward y, forward z, multiply with FT(map) or FT(map^2),
//                           inverse z, inverse y, write X2[pair][vol][z][kx][y]
//                                                                    (2/3 of K5, K6, 2/3 of K7)
//   C  fused_ifftx_lcc    : inverse x of the gcc/ave/ave2 rows, LCC normalisation and the
//                           running arg-max in registers/shared memory (1/3 of K7, K8-K10)
// (K numbers: SURVEY.md section 2a; reference code powerfitter.py:487-531, kernels.cl:162-226.)
//
// Only two HBM round trips per transform remain, the rotated grids and the gcc/ave/ave2
// grids never exist in memory, and -- because template and mask are compact -- kernels A
// and B skip every (y,z) row outside the template's support box (rows that are exactly
// zero), which removes most of the forward traffic.
//
// All transforms use the kernel exp(+2 pi i k r / N) (see fft_generic.cu).  A pencil of N
// points is handled by 8 threads of one warp holding E = N/8 points each:
//   X[k1 + E k0] = sum_{n0<8} W8^(n0 k0) W_N^(n0 k1) sum_{n1<E} x[n0 + 8 n1] W_E^(n1 k1)
// thread n0 does the E-point DFT in registers, multiplies by W_N^(n0 k1), the eight threads
// exchange through the pencil's own shared-memory storage (XOR-swizzled, conflict free) and
// thread u finishes with the 8-point DFTs of k1 = u, u+8, ...; its outputs are X[u + 8m],
// i.e. the same distribution the inputs had, in natural order -- no digit reversal.
#include "common.cuh"
#include "fft_core.cuh"
#include "rotate_device.cuh"
#include "tmem.cuh"
#include "tma.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>

namespace pfb {

// ------------------------------------------------------------------------------- kernel A
// L lanes per x pencil (E = N / L points each); 32 rows per CTA -> 32 L threads
// N = nx (the pencil length); ny and nz are run-time
template <int N, int L>
__global__ void __launch_bounds__(32 * L)
fused_rotate_fftx_kernel(const float4 *__restrict__ tmplq, const float *__restrict__ mask,
                         const double *__restrict__ rot, int first, int count, int nsig,
                         float2 *__restrict__ X1, const float2 *__restrict__ twN, int rs, int rs2,
                         unsigned ymask, int nzv, int ny, int nz) {
    constexpr int E = N / L, TP = 33, THREADS = 32 * L;
    const int rmax = min(N, min(ny, nz)) / 2;
    extern __shared__ float2 smem[];
    float2 *tile_t = smem, *tile_m = smem + N * TP;
    const int pair = blockIdx.y;
    const int j = blockIdx.x % nzv;
    int yt_rank = blockIdx.x / nzv, ytile = 0;
    for (unsigned mbits = ymask;; ++ytile) {
        if (mbits & 1u) { if (yt_rank == 0) break; --yt_rank; }
        mbits >>= 1;
    }
    const int z = (j - rs + nz) % nz, y0 = 32 * ytile;
    const GridDims d{nz, ny, N, rmax, (long)nz * ny * N};
    const int ra = first + 2 * pair;
    const bool have_b = 2 * pair + 1 < count;
    const double *Ra = rot + (long)ra * 9, *Rb = Ra + 9;
    const int oz = z <= nz / 2 ? z : z - nz;

    // ---- gather the 32 x N tile of both signals (zeros outside the sphere / support)
    //      zero-fill first, then visit only the x offsets inside the support box [xlo, rs];
    //      offset -N/2 is skipped: it aliases index N/2, which belongs to offset +N/2.
    // both tiles are one contiguous run of 2 N TP float2 (an even count): 16-byte stores
    for (int idx = threadIdx.x; idx < N * TP; idx += THREADS)
        reinterpret_cast<float4 *>(smem)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int xlo = max(-rs, -(N / 2 - 1));
    const int lim2 = min(rs2, rmax * rmax);
    // both rotations of a position are fetched before either is blended (six gathers in flight); an odd
    // last pair samples rotation a twice and discards the copy.
    // Rows are dealt to groups of LPR lanes; a group walks only the x span of its row that lies inside the
    // sphere (every lane of a busy group has a sample), with the (z, y) part of the source coordinate -- the
    // reference's per-row partial sums -- computed once per row.
    const double *Rb2 = have_b ? Rb : Ra;
    {
        constexpr int LPR = 16, GROUPS = THREADS / LPR;
        const int gl = threadIdx.x % LPR;
        for (int r = threadIdx.x / LPR; r < 32; r += GROUPS) {
            const int iy = y0 + r;
            const int oy = iy <= ny / 2 ? iy : iy - ny;
            const int rem = lim2 - oy * oy - oz * oz;
            if (rem < 0) continue;
            const int hw = isqrt_floor(rem);
            const int lo = max(xlo, -hw), hi = min(rs, hw);
            const SrcCoord rowa = source_row(Ra, oy, oz), rowb = source_row(Rb2, oy, oz);
            for (int ox = lo + gl; ox <= hi; ox += LPR) {
                const SrcCoord ca = source_in_row(rowa, Ra, ox), cb = source_in_row(rowb, Rb2, ox);
                const float ta = sample_trilinear_q(tmplq, d, ca), tb = sample_trilinear_q(tmplq, d, cb);
                const float ma = sample_nearest(mask, d, ca), mb = sample_nearest(mask, d, cb);
                const int x = ox < 0 ? ox + N : ox;
                tile_t[x * TP + r] = make_float2(ta, have_b ? tb : 0.f);
                tile_m[x * TP + r] = make_float2(ma, have_b ? mb : 0.f);
            }
        }
    }
    __syncthreads();

    // ---- x transforms: thread (row r, t)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & (L - 1), r = warp + L * (lane / L);
    float2 tw[E];
    load_twiddles<E>(tw, twN, t);
    float2 v[E], v2[E];
#pragma unroll
    for (int n1 = 0; n1 < E; ++n1) v[n1] = tile_t[(t + L * n1) * TP + r];
    fft_pencil<E, L>(v, tile_t + r, TP, t, tw, true);
#pragma unroll
    for (int m = 0; m < E; ++m) tile_t[(t + L * m) * TP + r] = v[m];
#pragma unroll
    for (int n1 = 0; n1 < E; ++n1) {
        v[n1] = tile_m[(t + L * n1) * TP + r];
        v2[n1] = make_float2(v[n1].x * v[n1].x, v[n1].y * v[n1].y);
    }
    fft_pencil<E, L>(v, tile_m + r, TP, t, tw, true);
#pragma unroll
    for (int m = 0; m < E; ++m) tile_m[(t + L * m) * TP + r] = v[m];
    __syncthreads();

    // ---- coalesced write-out: 32 consecutive y (256 B) per kx, pairs of y interleaved
    //      (re[y], re[y+1], im[y], im[y+1]) -- the layout kernel B's packed passes work in
    const int H = ny / 2;
    const size_t slab = (size_t)N * H;
    float4 *X14 = reinterpret_cast<float4 *>(X1);
    float4 *o_t = X14 + ((size_t)(pair * nsig + 0) * nz + z) * slab + y0 / 2;      // X1[pair][sig][z][kx][y/2]
    float4 *o_m = X14 + ((size_t)(pair * nsig + 1) * nz + z) * slab + y0 / 2;
    for (int idx = threadIdx.x; idx < 16 * N; idx += THREADS) {
        const int kx = idx >> 4, jj = idx & 15;
        const float2 a = tile_t[kx * TP + 2 * jj], b = tile_t[kx * TP + 2 * jj + 1];
        o_t[(size_t)kx * H + jj] = make_float4(a.x, b.x, a.y, b.y);
        const float2 c = tile_m[kx * TP + 2 * jj], e = tile_m[kx * TP + 2 * jj + 1];
        o_m[(size_t)kx * H + jj] = make_float4(c.x, e.x, c.y, e.y);
    }
    if (nsig == 3) {
        __syncthreads();
        fft_pencil<E, L>(v2, tile_t + r, TP, t, tw, true);
#pragma unroll
        for (int m = 0; m < E; ++m) tile_t[(t + L * m) * TP + r] = v2[m];
        __syncthreads();
        float4 *o_2 = X14 + ((size_t)(pair * nsig + 2) * nz + z) * slab + y0 / 2;
        for (int idx = threadIdx.x; idx < 16 * N; idx += THREADS) {
            const int kx = idx >> 4, jj = idx & 15;
            const float2 a = tile_t[kx * TP + 2 * jj], b = tile_t[kx * TP + 2 * jj + 1];
            o_2[(size_t)kx * H + jj] = make_float4(a.x, b.x, a.y, b.y);
        }
    }
}

// ------------------------------------------------------------------------------- kernel B
// Work-buffer layout (X1 and X2 alike): [pair][volume][z][kx][y/2] of float4 =
// (re[y], re[y+1], im[y], im[y+1]) -- two neighbouring y of one complex grid per 16 bytes, so
// that every pass below works on packed pairs (fft_core.cuh) without any re-shuffling:
//   phase 1  rows (fixed z), forward y : adjacent-in -> split-out, plane[z][ky] = (Y[ky], Y[ky+N/2])
//   phase 2  columns ky and ky+N/2 as two independent pencils: forward z, multiply with the
//            map spectrum (stored in the same pairing, Fpk[kx][ky][kz]), inverse z
//   phase 3  rows, inverse y : split-in -> adjacent-out, straight to X2
// Pencil geometry per axis length: column (z) pencils of L x E points, row (y) pencils of LM x EM packed pairs.
template <int N> struct AxisCfg;
template <> struct AxisCfg<32> { static constexpr int L = 4, E = 8, LM = 4, EM = 4; };
template <> struct AxisCfg<64> { static constexpr int L = 8, E = 8, LM = 4, EM = 8; };
template <> struct AxisCfg<96> { static constexpr int L = 4, E = 24, LM = 4, EM = 12; };      // 24 = 3 x 8, 12 = 3 x 4
template <> struct AxisCfg<128> { static constexpr int L = 8, E = 16, LM = 8, EM = 8; };
constexpr int cmin(int a, int b) { return a < b ? a : b; }
constexpr int cmax(int a, int b) { return a > b ? a : b; }
constexpr uint32_t pow2ceil(uint32_t v) { uint32_t r = 32; while (r < v) r *= 2; return r; }
// Kernel B of an (NZ, NY) plane: one warp per group of 32 / LN column pairs -> NY LN / 2 threads (512 at 128 x 128,
// where one CTA fills the register file; 256 at 64 x 64, where two CTAs per SM overlap their phases); as many CTAs
// per SM as registers (128 per thread, 232 with 24-point register DFTs) and shared memory allow, at most four
template <int NZ, int NY> struct FusedCfg {
    static constexpr int LN = AxisCfg<NZ>::L, EN = AxisCfg<NZ>::E, LM = AxisCfg<NY>::LM, EM = AxisCfg<NY>::EM;
    static constexpr int THREADS = NY * LN / 2, REGS = EN >= 24 ? 232 : 128;
    static constexpr size_t PLANE = (size_t)(NZ * (NY / 2 + 1)) * sizeof(float4) + (size_t)(NZ + NY) * sizeof(float2) + 128;
    static constexpr int CTAS = cmax(1, cmin(4, cmin((int)(227 * 1024 / (PLANE + 1024)), 65536 / (THREADS * REGS))));
    // the TMEM stash of a binary mask's spectrum (tmem.cuh moves 8 packed pairs at a time: 8-lane column pencils);
    // without it the mask is transformed forward twice, once per product
    static constexpr bool STASH = LN == 8;
    static constexpr uint32_t TCOLS = pow2ceil((uint32_t)((THREADS / 32 + 3) / 4 * 4 * EN));
};

// Persistent: one CTA per SM walks jobs j = blockIdx.x, + gridDim.x, ...; a job is one (pair, kx) and consists
// of the three output planes gcc, ave, ave2 in that order.  The row loop does phase 3 of the current plane and
// phase 1 of the CTA's next plane row by row (a row's storage is free for the next plane the moment its inverse
// transform has left for HBM), so there are two block barriers per plane and nothing drains between planes.
// Measured alternatives: staging the next plane's rows in shared memory with cp.async is 4 % slower, and
// fetching them into registers before the inverse transform 20 % slower (register pressure) -- the kernel is
// bound by shared-memory bandwidth, not by load latency.
//
// Binary mask (nsig == 2): ave and ave2 are products of the SAME forward spectrum with FT(map) and FT(map^2).
// The spectrum is computed once, in the ave plane; every thread parks the values it holds in tensor memory
// (tmem.cuh) before multiplying, and the ave2 plane has no phase 1 and no forward z: its phase 2 fetches the
// parked spectrum, multiplies by FT(map^2) and transforms back.  That removes one of three forward (y,z)
// transforms from the shared-memory pipe.
//
// STAGED: the support rows of the next forward plane do not come straight from global memory inside the row loop
// (an L2 round trip per row that 16 resident warps cannot hide: 11 % of the stall samples,
// profiles/r02_ncu_v1_fused_full.txt) but are copied by TMA -- two 3-D boxes, rows z = 0..rs and z = N-rs..N-1 of
// the (pair, signal, kx) plane of X1, issued by one thread while phase 2 runs -- into a staging area behind the
// plane, and the row loop reads them from shared memory.  Used when the staging area fits (2 rs + 2 rows).
// NZ x NY: the plane (z rows, y columns); the number of planes per volume is the constant NXT, or run-time (nx_rt)
// when NXT = 0 (the cubes keep it a constant: one more live register spills in the 128^3 kernel).
template <int NZ, int NY, int NXT, bool STAGED>
__global__ void __launch_bounds__(FusedCfg<NZ, NY>::THREADS, FusedCfg<NZ, NY>::CTAS)
fused_fftyz_mul_kernel(const float4 *__restrict__ X1, float4 *__restrict__ X2, const float4 *__restrict__ Fpk,
                       const float4 *__restrict__ F2pk, const float2 *__restrict__ twN_g,
                       const float2 *__restrict__ twM_g, const float2 *__restrict__ twh_g, int rs,
                       unsigned ymask, int nsig, int npairs, int nx_rt, const __grid_constant__ CUtensorMap tmapX1) {
    using Cfg = FusedCfg<NZ, NY>;
    const int nx = NXT ? NXT : nx_rt;
    constexpr int N = NZ, THREADS = Cfg::THREADS;              // N: rows of the plane = length of the column pencils
    constexpr int H = NY / 2, P = H + 1;
    constexpr int LN = Cfg::LN, EN = Cfg::EN, GN = 32 / LN;      // column pencils: N points
    constexpr int LM = Cfg::LM, EM = Cfg::EM, GM = 32 / LM;      // row pencils: packed N/2 points
    constexpr int NW = THREADS / 32;
    constexpr bool STASH = Cfg::STASH;
    constexpr uint32_t TCOLS = Cfg::TCOLS;                       // TMEM columns: 4 EN per warp, 4 warps per lane quarter
    static_assert(H / GN == NW, "one column group per warp: the TMEM stash is indexed by warp");
    static_assert(!STASH || TCOLS * Cfg::CTAS <= 512, "TMEM allocation");
    extern __shared__ float4 smem4[];
    float4 *plane = smem4;                                        // [N][P]
    float2 *twN = reinterpret_cast<float2 *>(plane + N * P);      // [EN][LN] W_N^(t k1)
    float2 *twM = twN + N;                                        // [EM][LM] W_H^(t k1)
    float2 *twh_s = twM + H;                                      // [H] W_N^k of the split radix-2 step
    uint32_t *tslot = reinterpret_cast<uint32_t *>(twh_s + H);
    uint64_t *sbar = reinterpret_cast<uint64_t *>(tslot + 2);                          // staging copies have landed
    // staging rows: zi = z for z <= rs, zi = rs + 1 + (z - (N - rs)) for the rows below zero; 128-byte aligned
    float4 *stage = reinterpret_cast<float4 *>(reinterpret_cast<uint8_t *>(sbar + 1) + 112);
    const size_t slab = (size_t)nx * H;                                                // float4 per z
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // LM = 4: a quarter warp holds rows gM and gM + 4, whose storage is 64 bytes apart modulo the
    // 128-byte bank window (P odd), instead of two neighbouring rows that would collide
    const int tM = lane & (LM - 1), gM = LM == 4 ? (lane >> 3) + 4 * ((lane >> 2) & 1) : lane / LM;
    const int tN = lane & (LN - 1), gN = lane / LN;
    const int nzv = min(2 * rs + 1, N);
    const bool binary = STASH && nsig == 2;
    auto sig_of = [&](int v) { return v < nsig ? v : nsig - 1; };      // ave2 of a binary mask without stash: the mask again
    // job -> (pair, kx), pair fastest: the jobs in flight at any time share their map-spectrum planes (kx)
    const int njobs = npairs * nx;

    int job = blockIdx.x, vol = 0;      // plane whose phase 1 comes next
    int cjob = -1, cvol = 0;            // plane whose phase 2 is done (phase 3 pending)
    // column twiddles in pairs: twN[(k1 / 2) * LN + t] = (W^(t 2k), W^(t (2k+1))) as float4
    for (int i = threadIdx.x; i < N; i += THREADS) {
        const int k1 = i / LN, tt = i % LN;
        twN[2 * ((k1 >> 1) * LN + tt) + (k1 & 1)] = twN_g[i];
    }
    for (int i = threadIdx.x; i < H; i += THREADS) { twM[i] = twM_g[i]; twh_s[i] = twh_g[i]; }
    if (STASH && warp == 0) tmem_alloc(tslot, TCOLS);
    // one thread: both boxes of plane (j, v) of X1 -> staging area
    auto stage_issue = [&](int j, int v) {
        const int pair = j % npairs, kx = j / npairs;
        mbar_arrive_expect_tx(sbar, (uint32_t)(2 * (rs + 1) * H * sizeof(float4)));
        tma_load_4d(stage, &tmapX1, sbar, 0, kx, 0, pair * nsig + sig_of(v));
        tma_load_4d(stage + (rs + 1) * H, &tmapX1, sbar, 0, kx, N - rs, pair * nsig + sig_of(v));   // last row out of bounds: zeros
    };
    if (STAGED && threadIdx.x == 0) {
        mbar_init(sbar, 1);
        mbar_fence_init();
        if (job < njobs) stage_issue(job, 0);
    }
    uint32_t sphase = 0;
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    // this warp's stash: lane quarter warp % 4 (the only one a warp may address), 4 EN columns
    const uint32_t tcol = STASH ? *tslot + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * 4 * EN) : 0u;

    while (true) {
        __syncthreads();          // phase 2 of the current plane is complete (first pass: the tables are in place)
        // the ave2 plane of a binary mask has no phase 1: its spectrum waits in TMEM
        const bool fwd = job < njobs && !(binary && vol == 2);
        {
            // ---- row loop: phase 3 of plane (cjob, cvol) (inverse y, shared -> HBM), then phase 1 of plane
            //      (job, vol) (forward y of the rows inside the support box, HBM -> shared)
            float2 twr[EM], twh[EM];
#pragma unroll
            for (int m = 0; m < EM; ++m) { twr[m] = twM[m * LM + tM]; twh[m] = twh_s[tM + LM * m]; }
            const TwReg<EM> tw{twr};
            const float4 *src = X1;
            if (fwd) {
                const int pair = job % npairs, kx = job / npairs;
                src = X1 + (size_t)(pair * nsig + sig_of(vol)) * N * slab + (size_t)kx * H;    // + z*slab + y/2
            }
            float4 *dst = X2;
            if (cjob >= 0) {
                const int pair = cjob % npairs, kx = cjob / npairs;
                dst = X2 + (size_t)(pair * 3 + cvol) * N * slab + (size_t)kx * H;      // + z*slab + y/2
            }
            if (STAGED && fwd) { mbar_wait(sbar, sphase); sphase ^= 1u; }
            for (int w = warp; w < N / GM; w += NW) {
                const int z = w * GM + gM;
                const bool act = fwd && (z + rs) % N < nzv;
                const bool any = __any_sync(0xffffffffu, act);
                if (cjob >= 0) {
                    C2 v[EM];
#pragma unroll
                    for (int n1 = 0; n1 < EM; ++n1) v[n1] = lds_c2(plane + z * P + tM + LM * n1);
                    fft_row_split2adj<LM, EM>(v, plane + z * P, 1, tM, tw, twh);
#pragma unroll
                    for (int m = 0; m < EM; ++m) stg_c2(dst + (size_t)z * slab + tM + LM * m, v[m]);
                }
                if (any) {
                    C2 vn[EM];                                                         // next plane's row z
#pragma unroll
                    for (int n1 = 0; n1 < EM; ++n1) {
                        const int jj = tM + LM * n1;                                   // y = 2jj, 2jj+1
                        const bool have = act && ((ymask >> (jj >> 4)) & 1u);
                        if (STAGED) {
                            const int zi = z <= rs ? z : z - (N - rs) + rs + 1;
                            vn[n1] = have ? lds_c2(stage + zi * H + jj) : c2_zero();
                        } else {
                            vn[n1] = have ? ldg_c2(src + (size_t)z * slab + jj) : c2_zero();
                        }
                    }
                    // a pencil group outside the support box rides along without touching shared memory
                    fft_row_adj2split<LM, EM>(vn, plane + z * P, 1, tM, tw, twh, act);
                    if (act) {
#pragma unroll
                        for (int m = 0; m < EM; ++m) sts_c2(plane + z * P + tM + LM * m, vn[m]);
                    }
                }
            }
        }
        __syncthreads();
        if (job >= njobs) break;

        // the CTA's next forward plane: copy its support rows into the staging area (STAGED), or pull them towards
        // L2 (one 128-byte line per thread), while phase 2 runs
        {
            int jn = job, vn = vol + 1;
            if (vn == 3 || (binary && vn == 2)) { vn = 0; jn += gridDim.x; }
            if (STAGED) {
                // (a plane without phase 1 -- ave2 of a binary mask -- changes nothing: its successor's rows were
                // staged during the previous phase 2 and are still waiting)
                if (threadIdx.x == 0 && jn < njobs && fwd) {
                    fence_proxy_async_smem();
                    stage_issue(jn, vn);
                }
            } else {
                const int nlines = nzv * 2 * __popc(ymask);
                if (jn < njobs && (int)threadIdx.x < nlines) {
                    const int pair = jn % npairs, kx = jn / npairs;
                    const int j = threadIdx.x / (2 * __popc(ymask)), l = threadIdx.x % (2 * __popc(ymask));
                    const int z = (j - rs + N) % N, tile = __fns(ymask, 0, (l >> 1) + 1);
                    const float4 *a = X1 + (size_t)(pair * nsig + sig_of(vn)) * N * slab + (size_t)kx * H + (size_t)z * slab +
                                      16 * tile + 8 * (l & 1);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
                }
            }
        }

        // ---- phase 2: forward z, multiply with the map spectrum, inverse z (column pairs ky, ky+H)
        {
            const int kx = job / npairs;
            const float4 *Fm = (vol == 2 ? F2pk : Fpk) + (size_t)kx * H * N;           // + ky*N + kz
            const TwSmemPair<LN> tw{reinterpret_cast<const float4 *>(twN) + tN};
            const int ky = warp * GN + gN;
            C2 v[EN];
            if (fwd) {
#pragma unroll
                for (int n1 = 0; n1 < EN; ++n1) {
                    const int z = tN + LN * n1;
                    const int sz = z <= N / 2 ? z : z - N;
                    v[n1] = (sz >= -rs && sz <= rs) ? lds_c2(plane + z * P + ky) : c2_zero();
                }
                if constexpr (STASH)
                    fft_pencil2_mul_stash<false, LN, EN>(v, plane + ky, P, tN, tw, Fm + (size_t)ky * N + tN, tcol,
                                                         binary && vol == 1);
                else
                    fft_pencil2_mul<LN, EN>(v, plane + ky, P, tN, tw, Fm + (size_t)ky * N + tN);
            } else {
                if constexpr (STASH)
                    fft_pencil2_mul_stash<true, LN, EN>(v, plane + ky, P, tN, tw, Fm + (size_t)ky * N + tN, tcol, false);
            }
            fft_pencil2<LN, EN>(v, plane + ky, P, tN, tw);
#pragma unroll
            for (int m = 0; m < EN; ++m) sts_c2(plane + (tN + LN * m) * P + ky, v[m]);
            if (STASH) tmem_wait_st();       // the parked spectrum is in place before this thread passes the next barrier
        }
        cjob = job; cvol = vol;
        if (++vol == 3) { vol = 0; job += gridDim.x; }
    }
    if (STASH && warp == 0) tmem_dealloc(*tslot, TCOLS);
}

// ------------------------------------------------------------------------------- kernel C
// One CTA owns a tile of RT rows (fixed z, RT consecutive y = RT/2 y pairs) and
// walks a chunk of rotation pairs.  Per pair the three X2 tiles (ave2, ave, gcc) are copied
// with cp.async into shared memory in exactly the layout kernel B wrote -- [kx][RT/2 y pairs] of
// float4 -- and eight lanes transform one y PAIR along x as two independent packed pencils.
// A thread keeps 1/sqrt(var) of its 2 rows x 16 x x 2 rotations in registers until the gcc
// tile arrives; the running best lives in shared memory and goes to HBM as one atomicMax per
// voxel per chunk.  The tile is single-buffered: the CTAs of an SM overlap each other's loads,
// and the next tile's copy is issued before the epilogue arithmetic of the current one.
// RT rows per tile (16 by default, 32 with PFB_C_RT=32): 8 lanes per y pair -> 4 RT threads; 96 / RT CTAs per SM.
// Measured at 128^3: 5.28 us/rotation with 16-row tiles (six 64-thread CTAs per SM interleave their load and
// compute phases more finely), 5.61 with 32 rows, 5.85 with 8.
// L lanes per x pencil (8 for the 64- and 128-point pencils, 4 for 32 = 4 x 8 and 96 = 4 x 24).
template <int N, int L, int RT>
__global__ void __launch_bounds__(L * RT / 2, L == 8 ? 96 / RT : (N / L >= 24 ? 4 : 8))
fused_ifftx_lcc_kernel(const float4 *__restrict__ X2, const uint32_t *__restrict__ mbits, float norm,
                       int first_index, int count, int pairs_per_chunk, int64_t *__restrict__ best,
                       const float2 *__restrict__ twN_g, int ny, int nz) {
    constexpr int E = N / L, NP = RT / 2, TP = NP + 1, BP = N + 4, THREADS = L * RT / 2, NBUF = 1;
    const int H = ny / 2;
    extern __shared__ float4 smem4[];
    float4 *tile0 = smem4;                                            // [NBUF][N][TP]
    // running best of the chunk as (LCC, rotation) pairs: rotations arrive in increasing order, so a
    // strict float '>' against a +0.0 start is exactly the packed-key order (common.cuh)
    float2 *lbest = reinterpret_cast<float2 *>(tile0 + NBUF * N * TP);     // [RT][BP] (lcc, rot index bits)
    float2 *tws = reinterpret_cast<float2 *>(lbest + RT * BP);        // [E][L] W_N^(t k1)
    const int y0 = RT * blockIdx.x, z = blockIdx.y;
    const int npairs = (count + 1) / 2;
    const int p0 = blockIdx.z * pairs_per_chunk, p1 = min(npairs, p0 + pairs_per_chunk);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & (L - 1), rp = (32 / L) * warp + lane / L;     // y pair: rows y0+2rp, y0+2rp+1
    const size_t slab = (size_t)N * H;
    const size_t rowa = ((size_t)z * ny + y0 + 2 * rp) * N, rowb = rowa + N;
    float2 *lba = lbest + (2 * rp) * BP + t, *lbb = lba + BP;
    // bit m of a row's word t: lcc_mask at x = t + L m (built once per target by mask_bits_kernel)
    const unsigned ma = mbits[((size_t)z * ny + y0 + 2 * rp) * 8 + t], mb = mbits[((size_t)z * ny + y0 + 2 * rp + 1) * 8 + t];
#pragma unroll
    for (int m = 0; m < E; ++m) {
        lba[L * m] = make_float2(0.f, 0.f);
        lbb[L * m] = make_float2(0.f, 0.f);
    }
    for (int i = threadIdx.x; i < N; i += THREADS) tws[i] = twN_g[i];
    const TwSmem<L> tw{tws + t};
    const int nitems = 3 * (p1 - p0);
    auto prefetch = [&](int item) {
        const int p = p0 + item / 3, vol = 2 - item % 3;              // ave2, ave, gcc
        const float4 *src = X2 + ((size_t)(p * 3 + vol) * nz + z) * slab + y0 / 2;
        float4 *tile = tile0 + (item % NBUF) * N * TP;
        for (int idx = threadIdx.x; idx < NP * N; idx += THREADS) {
            const int kx = idx / NP, c = idx % NP;
            cp_async16(tile + kx * TP + c, src + (size_t)kx * H + c);
        }
        cp_async_commit();
    };
    if (nitems > 0) prefetch(0);
    C2 sd[E];
    constexpr int Q = E / L;
    for (int item = 0; item < nitems; ++item) {
        if (NBUF == 2) {
            // the other buffer was released by the barrier at the end of the previous item
            if (item + 1 < nitems) { prefetch(item + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        float4 *tile = tile0 + (item % NBUF) * N * TP;
        const int p = p0 + item / 3, vi = item % 3;
        const uint32_t ia = (uint32_t)(first_index + 2 * p);
        const bool have_b = 2 * p + 1 < count;
        // one straight-line body per volume kind (VI = 0 ave2, 1 ave, 2 gcc)
        auto run_item = [&](auto vi_tag) {
            constexpr int VI = decltype(vi_tag)::value;
            {
                C2 v[E];
#pragma unroll
                for (int n1 = 0; n1 < E; ++n1) v[n1] = lds_c2(tile + (t + L * n1) * TP + rp);
                pencil2_stage1<L, E>(v, tile + rp, TP, t, tw);
            }
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                // outputs x = t + L m, m = q + Q k0, of this chunk go straight into the epilogue
                C2 a[L];
                pencil2_stage2<L>(a, tile + rp, TP, t, q);
                if (q == Q - 1 && NBUF == 1) {
                    __syncthreads();       // every pencil is out of the tile: refill it while the arithmetic runs
                    if (item + 1 < nitems) prefetch(item + 1);
                }
#pragma unroll
                for (int k0 = 0; k0 < L; ++k0) {
                    const int m = q + Q * k0;
                    if (VI == 0) {
                        sd[m] = a[k0];                                     // ave2
                    } else if (VI == 1) {                                  // 1/sqrt(N ave2 - ave^2)
                        // var <= 0 gives inf / NaN exactly where the reference's gcc/sqrt(var) does
                        const float2 vr = psub(pmul(sd[m].re, pdup(norm)), pmul(a[k0].re, a[k0].re));
                        const float2 vq = psub(pmul(sd[m].im, pdup(norm)), pmul(a[k0].im, a[k0].im));
                        sd[m].re = make_float2(rsqrtf(vr.x), rsqrtf(vr.y));
                        sd[m].im = make_float2(rsqrtf(vq.x), rsqrtf(vq.y));
                    } else {
                        // (row a, row b) of rotation a / rotation b
                        const float2 la = pmul(a[k0].re, sd[m].re), lb = pmul(a[k0].im, sd[m].im);
                        // best of the pair first (ties and NaN: rotation a, the lower index, stays)
                        const bool sa = have_b && (lb.x > la.x || !(la.x == la.x));
                        const bool sb = have_b && (lb.y > la.y || !(la.y == la.y));
                        const float ca = sa ? lb.x : la.x, cb = sb ? lb.y : la.y;          // NaN never passes '>'
                        const uint32_t ja = sa ? ia + 1 : ia, jb = sb ? ia + 1 : ia;
                        if (((ma >> m) & 1u) && ca > lba[L * m].x) lba[L * m] = make_float2(ca, __uint_as_float(ja));
                        if (((mb >> m) & 1u) && cb > lbb[L * m].x) lbb[L * m] = make_float2(cb, __uint_as_float(jb));
                    }
                }
            }
        };
        if (vi == 0) run_item(std::integral_constant<int, 0>{});
        else if (vi == 1) run_item(std::integral_constant<int, 1>{});
        else run_item(std::integral_constant<int, 2>{});
        if (NBUF == 2) __syncthreads();   // this item's tile may be refilled by the next iteration's prefetch
    }
#pragma unroll
    for (int m = 0; m < E; ++m) {
        if ((ma >> m) & 1u) {
            const float2 b = lba[L * m];
            if (b.x > 0.f)
                atomicMax(reinterpret_cast<long long *>(best + rowa + t + L * m),
                          (long long)pack_best(__float_as_uint(b.x), __float_as_uint(b.y)));
        }
        if ((mb >> m) & 1u) {
            const float2 b = lbb[L * m];
            if (b.x > 0.f)
                atomicMax(reinterpret_cast<long long *>(best + rowb + t + L * m),
                          (long long)pack_best(__float_as_uint(b.x), __float_as_uint(b.y)));
        }
    }
}

// ------------------------------------------------------------------------------- kernel C, TMA-fed
// Same work split and epilogue as above (16-row tiles, 64 threads), but the tiles arrive as 4-D tensor-map
// boxes (tma.cuh): one elected thread arms an mbarrier and issues ONE cp.async.bulk.tensor per tile instead
// of every thread issuing 32 16-byte cp.async copies, the copy engine keeps NBUF tiles in flight per CTA,
// and it lays the box out in the 128-byte swizzle -- float4 (row kx, column c) at row * 8 + (c ^ (row & 7)) --
// so the dense 128-byte rows need no padding to be conflict free.  With u = 8 t + (c ^ t) for lane t of the
// pencil in column c, the three access patterns of the in-place pencil become
//   load     x[t + 8 n1]                      at 64 n1 + u
//   exchange row 8 k1 + (t ^ k), k = k1 & 7   at 64 k1 + (u ^ 9 k)
//   gather   row 8 (t + 8 q) + (n0 ^ t)       at 512 q + 64 t + (u ^ 9 n0)
// i.e. one XOR with an immediate per access.
template <int N, int NBUF>
__global__ void __launch_bounds__(64, NBUF == 1 ? 6 : (NBUF == 2 ? 4 : 3))
fused_ifftx_lcc_tma_kernel(const __grid_constant__ CUtensorMap tmap, const uint32_t *__restrict__ mbits, float norm,
                           int first_index, int count, int pairs_per_chunk, int64_t *__restrict__ best,
                           const float2 *__restrict__ twN_g, int ny) {
    constexpr int RT = 16, E = N / 8, BP = N + 4, THREADS = 64, TILE = N * 8;    // TILE: float4 per tile
    extern __shared__ uint8_t smem_raw[];
    // the swizzle is a function of the shared-memory address: tiles start on a 1024-byte boundary
    float4 *tile0 = reinterpret_cast<float4 *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    float2 *lbest = reinterpret_cast<float2 *>(tile0 + NBUF * TILE);  // [RT][BP] (lcc, rot index bits)
    float2 *tws = lbest + RT * BP;                                    // [E][8] W_N^(t k1)
    uint64_t *full = reinterpret_cast<uint64_t *>(tws + N);           // [NBUF]
    const int y0 = RT * blockIdx.x, z = blockIdx.y;
    const int npairs = (count + 1) / 2;
    const int p0 = blockIdx.z * pairs_per_chunk, p1 = min(npairs, p0 + pairs_per_chunk);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & 7, rp = 4 * warp + (lane >> 3);              // y pair: rows y0+2rp, y0+2rp+1
    const int u = 8 * t + (rp ^ t);
    const size_t rowa = ((size_t)z * ny + y0 + 2 * rp) * N, rowb = rowa + N;
    float2 *lba = lbest + (2 * rp) * BP + t, *lbb = lba + BP;
    const unsigned ma = mbits[((size_t)z * ny + y0 + 2 * rp) * 8 + t], mb = mbits[((size_t)z * ny + y0 + 2 * rp + 1) * 8 + t];
    const int nitems = 3 * (p1 - p0);
    auto issue = [&](int item) {                                      // thread 0 only
        const int p = p0 + item / 3, vol = 2 - item % 3;              // ave2, ave, gcc
        uint64_t *bar = full + item % NBUF;
        mbar_arrive_expect_tx(bar, TILE * (uint32_t)sizeof(float4));
        tma_load_4d(tile0 + (item % NBUF) * TILE, &tmap, bar, 2 * y0, 0, z, p * 3 + vol);
    };
    if (threadIdx.x == 0) {
        for (int b = 0; b < NBUF; ++b) mbar_init(full + b, 1);
        mbar_fence_init();
        for (int i = 0; i < NBUF && i < nitems; ++i) issue(i);
    }
#pragma unroll
    for (int m = 0; m < E; ++m) {
        lba[8 * m] = make_float2(0.f, 0.f);
        lbb[8 * m] = make_float2(0.f, 0.f);
    }
    for (int i = threadIdx.x; i < N; i += THREADS) {        // twiddles in pairs, as in kernel B
        const int k1 = i / 8, tt = i % 8;
        tws[2 * ((k1 >> 1) * 8 + tt) + (k1 & 1)] = twN_g[i];
    }
    __syncthreads();
    const TwSmemPair<8> tw{reinterpret_cast<const float4 *>(tws) + t};
    C2 sd[E];
    constexpr int Q = E / 8;
    for (int item = 0; item < nitems; ++item) {
        mbar_wait(full + item % NBUF, (uint32_t)(item / NBUF) & 1u);
        float4 *tile = tile0 + (item % NBUF) * TILE;
        const int p = p0 + item / 3, vi = item % 3;
        const uint32_t ia = (uint32_t)(first_index + 2 * p);
        const bool have_b = 2 * p + 1 < count;
        // one straight-line body per volume kind (VI = 0 ave2, 1 ave, 2 gcc)
        auto run_item = [&](auto vi_tag) {
            constexpr int VI = decltype(vi_tag)::value;
            {
                C2 v[E];
#pragma unroll
                for (int n1 = 0; n1 < E; ++n1) v[n1] = lds_c2(tile + 64 * n1 + u);
                DftReg<E, C2>::run(v);
#pragma unroll
                for (int k = 0; k < E / 2; ++k) {
                    const float4 w = tw.pair(k);
                    if (k > 0) v[2 * k] = cmulw(v[2 * k], make_float2(w.x, w.y));
                    v[2 * k + 1] = cmulw(v[2 * k + 1], make_float2(w.z, w.w));
                }
                __syncwarp();
#pragma unroll
                for (int k1 = 0; k1 < E; ++k1) sts_c2(tile + 64 * k1 + (u ^ (9 * (k1 & 7))), v[k1]);
                __syncwarp();
            }
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                // outputs x = t + 8 m, m = q + Q k0, of this chunk go straight into the epilogue
                C2 a[8];
#pragma unroll
                for (int n0 = 0; n0 < 8; ++n0) a[n0] = lds_c2(tile + 512 * q + 64 * t + (u ^ (9 * n0)));
                DftReg<8, C2>::run(a);
                if (q == Q - 1) {
                    __syncthreads();       // every pencil is out of the tile: hand it back to the copy engine
                    if (threadIdx.x == 0 && item + NBUF < nitems) {
                        fence_proxy_async_smem();
                        issue(item + NBUF);
                    }
                }
#pragma unroll
                for (int k0 = 0; k0 < 8; ++k0) {
                    const int m = q + Q * k0;
                    if (VI == 0) {
                        sd[m] = a[k0];                                     // ave2
                    } else if (VI == 1) {                                  // 1/sqrt(N ave2 - ave^2)
                        // var <= 0 gives inf / NaN exactly where the reference's gcc/sqrt(var) does
                        const float2 vr = psub(pmul(sd[m].re, pdup(norm)), pmul(a[k0].re, a[k0].re));
                        const float2 vq = psub(pmul(sd[m].im, pdup(norm)), pmul(a[k0].im, a[k0].im));
                        sd[m].re = make_float2(rsqrtf(vr.x), rsqrtf(vr.y));
                        sd[m].im = make_float2(rsqrtf(vq.x), rsqrtf(vq.y));
                    } else {
                        // (row a, row b) of rotation a / rotation b
                        const float2 la = pmul(a[k0].re, sd[m].re), lb = pmul(a[k0].im, sd[m].im);
                        // best of the pair first (ties and NaN: rotation a, the lower index, stays)
                        const bool sa = have_b && (lb.x > la.x || !(la.x == la.x));
                        const bool sb = have_b && (lb.y > la.y || !(la.y == la.y));
                        const float ca = sa ? lb.x : la.x, cb = sb ? lb.y : la.y;          // NaN never passes '>'
                        const uint32_t ja = sa ? ia + 1 : ia, jb = sb ? ia + 1 : ia;
                        // the running best is read unconditionally: only a third of a percent of the warps have no
                        // lcc_mask bit in any lane, so a branch around the load only adds BSSY/BSYNC pairs
                        const float cura = lba[8 * m].x, curb = lbb[8 * m].x;
                        const bool ua = ((ma >> m) & 1u) != 0 && ca > cura, ub = ((mb >> m) & 1u) != 0 && cb > curb;
                        if (ua) lba[8 * m] = make_float2(ca, __uint_as_float(ja));
                        if (ub) lbb[8 * m] = make_float2(cb, __uint_as_float(jb));
                    }
                }
            }
        };
        if (vi == 0) run_item(std::integral_constant<int, 0>{});
        else if (vi == 1) run_item(std::integral_constant<int, 1>{});
        else run_item(std::integral_constant<int, 2>{});
    }
#pragma unroll
    for (int m = 0; m < E; ++m) {
        if ((ma >> m) & 1u) {
            const float2 b = lba[8 * m];
            if (b.x > 0.f)
                atomicMax(reinterpret_cast<long long *>(best + rowa + t + 8 * m),
                          (long long)pack_best(__float_as_uint(b.x), __float_as_uint(b.y)));
        }
        if ((mb >> m) & 1u) {
            const float2 b = lbb[8 * m];
            if (b.x > 0.f)
                atomicMax(reinterpret_cast<long long *>(best + rowb + t + 8 * m),
                          (long long)pack_best(__float_as_uint(b.x), __float_as_uint(b.y)));
        }
    }
}

static int encode_tensor_map(CUtensorMap *out, const void *base, int row_floats, int nkx, int nz, long outer,
                             const cuuint32_t (&box)[4], CUtensorMapSwizzle swizzle) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        PFB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !fn) {
            set_error("cuTensorMapEncodeTiled is not available in this driver");
            return PFB_ERR_CUDA;
        }
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    const cuuint64_t dims[4] = {(cuuint64_t)row_floats, (cuuint64_t)nkx, (cuuint64_t)nz, (cuuint64_t)outer};
    const cuuint64_t strides[3] = {(cuuint64_t)row_floats * 4, (cuuint64_t)row_floats * 4 * nkx,
                                   (cuuint64_t)row_floats * 4 * nkx * nz};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void *>(base), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return PFB_ERR_CUDA;
    }
    return PFB_OK;
}

int make_x2_tensor_map(CUtensorMap *out, const void *base, int row_floats, int nkx, int nz, long outer,
                       int box_floats, int box_kx) {
    const cuuint32_t box[4] = {(cuuint32_t)box_floats, (cuuint32_t)box_kx, 1, 1};
    return encode_tensor_map(out, base, row_floats, nkx, nz, outer, box,
                             box_floats * 4 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

int make_x1_tensor_map(CUtensorMap *out, const void *base, int row_floats, int nkx, int nz, long outer, int box_z) {
    const cuuint32_t box[4] = {(cuuint32_t)row_floats, 1, (cuuint32_t)box_z, 1};
    return encode_tensor_map(out, base, row_floats, nkx, nz, outer, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

// ------------------------------------------------------------------------------- helpers
// Fpk[kx][ky][kz] = (re F[kz][ky][kx], re F[kz][ky+ny/2][kx], im ..., im ...), ky < ny/2: the map
// spectrum in the column pairing of kernel B's phase 2
__global__ void pair_transpose_kernel(const float2 *__restrict__ F, float4 *__restrict__ Fpk, int nz, int ny, int nx) {
    __shared__ float4 tl[32][33];
    const int ky = blockIdx.z, H = ny / 2;
    const int kx0 = blockIdx.x * 32, kz0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const float2 a = F[((size_t)(kz0 + i) * ny + ky) * nx + kx0 + threadIdx.x];
        const float2 b = F[((size_t)(kz0 + i) * ny + ky + H) * nx + kx0 + threadIdx.x];
        tl[i][threadIdx.x] = make_float4(a.x, b.x, a.y, b.y);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
        Fpk[((size_t)(kx0 + i) * H + ky) * nz + kz0 + threadIdx.x] = tl[threadIdx.x][i];
}

// mbits[row * 8 + t], row = z*ny + y, N = nx, L lanes per x pencil: bit m of word t < L = (lcc_mask[row][t + L m] != 0)
// -- kernel C's lane layout
__global__ void mask_bits_kernel(const uint8_t *__restrict__ lcc_mask, uint32_t *__restrict__ mbits, int N, int L, long rows) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * 8) return;
    const long row = i >> 3;
    const int t = (int)(i & 7);
    uint32_t w = 0;
    if (t < L)
        for (int m = 0; m < N / L; ++m)
            if (lcc_mask[row * N + t + L * m]) w |= 1u << m;
    mbits[i] = w;
}

// corner table of the template for kernel A's trilinear gather (rotate_device.cuh: sample_trilinear_q)
__global__ void corner_table_kernel(const float *__restrict__ g, float4 *__restrict__ q, int nz, int ny, int nx) {
    const long V = (long)nz * ny * nx;
    for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long)gridDim.x * blockDim.x) {
        const int x = (int)(v % nx), y = (int)((v / nx) % ny);
        const long zrow = v / ((long)nx * ny) * ny;
        const int x1 = x + 1 == nx ? 0 : x + 1, y1 = y + 1 == ny ? 0 : y + 1;
        q[v] = make_float4(g[(zrow + y) * nx + x], g[(zrow + y) * nx + x1], g[(zrow + y1) * nx + x], g[(zrow + y1) * nx + x1]);
    }
}

// max squared distance from voxel 0 (periodic) of any voxel where template or mask is non-zero
__global__ void support_kernel(const float *__restrict__ tmpl, const float *__restrict__ mask, int nz, int ny,
                               int nx, int *out) {
    const long V = (long)nz * ny * nx;
    int best = -1;
    for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long)gridDim.x * blockDim.x) {
        if (tmpl[v] != 0.f || mask[v] != 0.f) {
            int x = (int)(v % nx), y = (int)((v / nx) % ny), z = (int)(v / ((long)nx * ny));
            x = min(x, nx - x); y = min(y, ny - y); z = min(z, nz - z);
            best = max(best, x * x + y * y + z * z);
        }
    }
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0 && best >= 0) atomicMax(out, best);
}

// ------------------------------------------------------------------------------- host side
// plane + twiddle tables + the TMEM base address slot + the staging barrier (padded to a 128-byte boundary) +
// `srows` staging rows of NY/2 float4 (0 = unstaged kernel)
template <int NZ, int NY> static constexpr size_t smem_b(int srows) {
    return FusedCfg<NZ, NY>::PLANE + (size_t)srows * (NY / 2) * sizeof(float4);
}
// staging rows that fit next to the plane(s) of an SM
template <int NZ, int NY> static constexpr int stage_capacity() {
    return (int)((227 * 1024 / FusedCfg<NZ, NY>::CTAS - 1024 - smem_b<NZ, NY>(0)) / ((NY / 2) * sizeof(float4)));
}
template <int N> static constexpr size_t smem_c(int rt) {
    return (size_t)N * (rt / 2 + 1) * sizeof(float4) + (size_t)rt * (N + 4) * sizeof(int64_t) + (size_t)N * sizeof(float2);
}

template <int N> static constexpr size_t smem_c_tma(int nbuf) {
    return 1024 + (size_t)nbuf * N * 8 * sizeof(float4) + (size_t)16 * (N + 4) * sizeof(float2) + (size_t)N * sizeof(float2) +
           (size_t)nbuf * sizeof(uint64_t);
}

// twiddle table of a LANES x E pencil: entry [k1][t] = exp(+2 pi i t k1 / (LANES E))
static int upload_pencil_twiddles(int lanes, int e, float2 **out) {
    std::vector<float2> h((size_t)lanes * e);
    const double n = (double)lanes * e;
    for (int t = 0; t < lanes; ++t)
        for (int k1 = 0; k1 < e; ++k1) {
            const double a = 2.0 * M_PI * (double)(t * k1) / n;
            h[(size_t)k1 * lanes + t] = make_float2((float)cos(a), (float)sin(a));
        }
    PFB_CUDA(cudaMalloc(out, sizeof(float2) * h.size()));
    PFB_CUDA(cudaMemcpy(*out, h.data(), sizeof(float2) * h.size(), cudaMemcpyHostToDevice));
    return PFB_OK;
}

// Axis lengths of the plain fused pipeline.  Every function below that depends on a length goes through one of
// these dispatchers, so a new length is one more case here plus an AxisCfg.
static bool fused_axis(int n) { return n == 32 || n == 64 || n == 96 || n == 128; }
template <int N> using IC = std::integral_constant<int, N>;
template <class F> static int dispatch_axis(int n, F &&f) {
    switch (n) {
        case 32: return f(IC<32>{});
        case 64: return f(IC<64>{});
        case 96: return f(IC<96>{});
        default: return f(IC<128>{});
    }
}
template <class F> static int dispatch_zy(int nz, int ny, F &&f) {
    return dispatch_axis(nz, [&](auto z) { return dispatch_axis(ny, [&](auto y) { return f(z, y); }); });
}
// x pencils: lanes per pencil (kernels A and C), rows per tile of the cp.async kernel C, and whether the TMA-fed
// kernel C (8-lane swizzle arithmetic) exists for the length
template <int N> struct XCfg {
    static constexpr int L = AxisCfg<N>::L, RT = L == 8 ? 16 : 32;
    static constexpr bool TMA = L == 8;
};

template <int N> static int fused_init_x(Plan *p) {
    constexpr int TP = 33, L = XCfg<N>::L, RT = XCfg<N>::RT;
    int rc;
    if ((rc = upload_pencil_twiddles(L, N / L, &p->twdX))) return rc;
    PFB_CUDA(cudaFuncSetAttribute(fused_rotate_fftx_kernel<N, L>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(2 * N * TP * sizeof(float2))));
    PFB_CUDA(cudaFuncSetAttribute(fused_ifftx_lcc_kernel<N, L, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem_c<N>(RT)));
    if constexpr (XCfg<N>::TMA) {
        PFB_CUDA(cudaFuncSetAttribute(fused_ifftx_lcc_tma_kernel<N, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem_c_tma<N>(1)));
        PFB_CUDA(cudaFuncSetAttribute(fused_ifftx_lcc_tma_kernel<N, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem_c_tma<N>(2)));
        PFB_CUDA(cudaFuncSetAttribute(fused_ifftx_lcc_tma_kernel<N, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem_c_tma<N>(3)));
    }
    return PFB_OK;
}

template <int NZ, int NY> static int fused_init_zy(Plan *p) {
    using Cfg = FusedCfg<NZ, NY>;
    int rc;
    if ((rc = upload_pencil_twiddles(Cfg::LN, Cfg::EN, &p->twdN))) return rc;
    if ((rc = upload_pencil_twiddles(Cfg::LM, Cfg::EM, &p->twdM))) return rc;
    constexpr int NXC = NZ == NY ? NZ : 0;
    PFB_CUDA(cudaFuncSetAttribute(fused_fftyz_mul_kernel<NZ, NY, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem_b<NZ, NY>(0)));
    PFB_CUDA(cudaFuncSetAttribute(fused_fftyz_mul_kernel<NZ, NY, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem_b<NZ, NY>(stage_capacity<NZ, NY>())));
    if (NXC) {
        PFB_CUDA(cudaFuncSetAttribute(fused_fftyz_mul_kernel<NZ, NY, NXC, false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b<NZ, NY>(0)));
        PFB_CUDA(cudaFuncSetAttribute(fused_fftyz_mul_kernel<NZ, NY, NXC, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem_b<NZ, NY>(stage_capacity<NZ, NY>())));
    }
    return PFB_OK;
}

bool fused_supported(int nz, int ny, int nx) {
    if (nz == ny && ny == nx && (nx == 192 || nx == 256)) return true;      // class path (fused_cls.cu)
    return fused_axis(nz) && fused_axis(ny) && fused_axis(nx);
}

int fused_init(Plan *p) {
    if (p->cls) return cls_init(p);
    int rc = dispatch_axis(p->nx, [&](auto nx) { return fused_init_x<decltype(nx)::value>(p); });
    if (rc) return rc;
    return dispatch_zy(p->nz, p->ny, [&](auto nz, auto ny) { return fused_init_zy<decltype(nz)::value, decltype(ny)::value>(p); });
}

int fused_prepare_target(Plan *p, cudaStream_t s) {
    if (p->cls) return cls_prepare_target(p, s);
    dim3 grid(p->nx / 32, p->nz / 32, p->ny / 2), block(32, 8);
    { LaunchScope ls(p, KC_OTHER, s);
      pair_transpose_kernel<<<grid, block, 0, s>>>(p->F, reinterpret_cast<float4 *>(p->Fq), p->nz, p->ny, p->nx); }
    { LaunchScope ls(p, KC_OTHER, s);
      pair_transpose_kernel<<<grid, block, 0, s>>>(p->F2, reinterpret_cast<float4 *>(p->F2q), p->nz, p->ny, p->nx); }
    { LaunchScope ls(p, KC_OTHER, s);
      const long rows = (long)p->nz * p->ny;
      const int L = dispatch_axis(p->nx, [](auto nx) { return XCfg<decltype(nx)::value>::L; });
      mask_bits_kernel<<<(unsigned)((rows * 8 + 255) / 256), 256, 0, s>>>(p->lcc_mask, p->mbits, p->nx, L, rows); }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

// support box of the template/mask pair (synchronises the stream: it needs one int back)
int fused_prepare_template(Plan *p, cudaStream_t s) {
    int *d_r2 = reinterpret_cast<int *>(p->best_scratch);
    int init = -1;
    PFB_CUDA(cudaMemcpyAsync(d_r2, &init, sizeof(int), cudaMemcpyHostToDevice, s));
    { LaunchScope ls(p, KC_OTHER, s);
      support_kernel<<<p->sm_count * 4, 256, 0, s>>>(p->tmpl, p->mask, p->nz, p->ny, p->nx, d_r2); }
    { LaunchScope ls(p, KC_OTHER, s);
      corner_table_kernel<<<p->sm_count * 8, 256, 0, s>>>(p->tmpl, p->tmplq, p->nz, p->ny, p->nx); }
    int r2 = -1;
    PFB_CUDA(cudaMemcpyAsync(&r2, d_r2, sizeof(int), cudaMemcpyDeviceToHost, s));
    PFB_CUDA(cudaStreamSynchronize(s));
    const int ny = p->ny, rmax = p->rmax;
    // a rotated sample at offset r can be non-zero only if |r| <= max_r + sqrt(3)
    double reach = (r2 < 0 ? 0.0 : sqrt((double)r2)) + 1.7320508075688772 + 1e-6;
    int rs = (int)floor(reach);
    if (rs > rmax) rs = rmax;
    double reach2 = reach * reach;
    p->rs = rs;
    p->rs2 = reach2 > (double)rmax * rmax ? rmax * rmax : (int)floor(reach2);
    if (const char *e = getenv("PFB_NO_PRUNE")) { if (atoi(e)) { p->rs = rmax; p->rs2 = rmax * rmax; } }
    unsigned ymask = 0;
    for (int tile = 0; tile < ny / 32; ++tile)
        for (int y = 32 * tile; y < 32 * tile + 32; ++y) {
            const int sy = y <= ny / 2 ? y : y - ny;
            if (sy >= -p->rs && sy <= p->rs) ymask |= 1u << tile;
        }
    p->ymask = ymask;
    unsigned nmask = 0;
    for (int y = 0; y < ny; ++y) {
        const int sy = y <= ny / 2 ? y : y - ny;
        if (sy >= -p->rs && sy <= p->rs) nmask |= 1u << ((y % 64) / (ny == 256 ? 4 : 8));      // ClsCfg<N>::RN
    }
    p->nmask = nmask;
    return PFB_OK;
}

// kernel A of a batch: rotate + forward x into the work buffer X1
template <int N, int L>
static int fused_a_n(Plan *p, int first, int count, cudaStream_t s) {
    constexpr int TP = 33;
    const int npairs = (count + 1) / 2;
    const int nzv = std::min(2 * p->rs + 1, p->nz);
    const int nyt = __builtin_popcount(p->ymask);
    LaunchScope ls(p, KC_FUSED_A, s);
    fused_rotate_fftx_kernel<N, L><<<dim3(nzv * nyt, npairs), 32 * L, 2 * N * TP * sizeof(float2), s>>>(
        p->tmplq, p->mask, p->rot_dev, first, count, p->nsig, p->A, p->tw[0], p->rs, p->rs2, p->ymask, nzv, p->ny,
        p->nz);
    return PFB_OK;
}

int launch_fused_a(Plan *p, int first, int count, cudaStream_t s) {
    return dispatch_axis(p->nx, [&](auto nx) {
        return fused_a_n<decltype(nx)::value, XCfg<decltype(nx)::value>::L>(p, first, count, s);
    });
}

// kernel B of a batch: y, z, multiply, z, y out of X1 into the work buffer X2
template <int NZ, int NY>
static int fused_b_n(Plan *p, int count, float2 *X2, cudaStream_t s) {
    using Cfg = FusedCfg<NZ, NY>;
    const int npairs = (count + 1) / 2;
    const int njobs = p->nx * npairs;
    const int grid = std::min(njobs, p->sm_count * Cfg::CTAS);
    // measured (per-kernel events, same box, three alternating repetitions): 128^3 11.96 -> 11.59 us/rotation with
    // staging, 64^3 1.64 -> 1.59
    static const int stage_env = getenv("PFB_B_STAGE") ? atoi(getenv("PFB_B_STAGE")) : 1;
    const int srows = 2 * p->rs + 2;
    const bool staged = stage_env != 0 && 2 * p->rs + 1 < NZ && srows <= stage_capacity<NZ, NY>();
    if (staged && (p->tmapB_base != (const void *)p->A || p->tmapB_rs != p->rs || p->tmapB_nsig != p->nsig)) {
        // X1 as [pair*nsig+sig][z][kx][2 ny floats]; box = one whole (z, kx) row x (rs + 1) consecutive z
        int rc = make_x1_tensor_map(&p->tmapB, p->A, 2 * NY, p->nx, NZ, (long)p->nsig * (p->batch / 2), p->rs + 1);
        if (rc) return rc;
        p->tmapB_base = p->A; p->tmapB_rs = p->rs; p->tmapB_nsig = p->nsig;
    }
    LaunchScope ls(p, KC_FUSED_B, s);
    auto launch = [&](auto kernel, size_t smem) {
        kernel<<<grid, Cfg::THREADS, smem, s>>>(
            reinterpret_cast<const float4 *>(p->A), reinterpret_cast<float4 *>(X2),
            reinterpret_cast<const float4 *>(p->Fq), reinterpret_cast<const float4 *>(p->F2q), p->twdN, p->twdM,
            p->tw[1], p->rs, p->ymask, p->nsig, npairs, p->nx, p->tmapB);
    };
    constexpr int NXC = NZ == NY ? NZ : 0;
    const bool cube = NXC != 0 && p->nx == NXC;
    if (staged) {
        if (cube) launch(fused_fftyz_mul_kernel<NZ, NY, NXC, true>, smem_b<NZ, NY>(srows));
        else launch(fused_fftyz_mul_kernel<NZ, NY, 0, true>, smem_b<NZ, NY>(srows));
    } else {
        if (cube) launch(fused_fftyz_mul_kernel<NZ, NY, NXC, false>, smem_b<NZ, NY>(0));
        else launch(fused_fftyz_mul_kernel<NZ, NY, 0, false>, smem_b<NZ, NY>(0));
    }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

// tiles in flight per kernel-C CTA: 0 = the cp.async kernel (kept as the cross-check), 1..3 = TMA ring depth
static int c_ring_depth() {
    static const int v = getenv("PFB_C_TMA") ? atoi(getenv("PFB_C_TMA")) : 1;
    return std::max(0, std::min(3, v));
}

template <int N, int NBUF>
static int fused_back_tma(Plan *p, int first, int count, int rot_index_offset, int64_t *best, const float2 *X2,
                          cudaStream_t s) {
    const int npairs = (count + 1) / 2;
    if (p->tmapC_base != (const void *)X2) {
        // X2 as [pair*3+vol][z][kx][2 ny floats]; box = 32 floats (8 y pairs) x all kx
        int rc = make_x2_tensor_map(&p->tmapC, X2, 2 * p->ny, N, p->nz, 3L * (p->batch / 2), 32, N);
        if (rc) return rc;
        p->tmapC_base = X2;
    }
    constexpr int per_sm = NBUF == 1 ? 6 : (NBUF == 2 ? 4 : 3);
    const int tiles = (p->ny / 16) * p->nz;
    // split the pair loop into chunks so that the CTAs fill whole waves of resident CTAs (a 4.6-wave launch
    // idles 8 % of the machine in its last wave); at least 8 pairs per CTA amortise its set-up and its atomics.
    // Measured at 128^3, 128 pairs: 32 pairs per chunk 4.59 us/rotation, 22 -> 4.51, 10 -> 4.52, 64 -> 5.08.
    const int slots = p->sm_count * per_sm;
    int ppc = std::max(1, npairs);
    double best_eff = -1.0;
    for (int cand = std::min(npairs, 32); cand >= std::min(npairs, 8); --cand) {
        const double waves = (double)tiles * ((npairs + cand - 1) / cand) / slots;
        const double eff = waves / std::ceil(waves);
        if (waves >= 3.0 && eff > best_eff + 1e-9) { best_eff = eff; ppc = cand; }
    }
    if (best_eff < 0) ppc = std::max(1, std::min(npairs, 8));
    static const int ppc_env = getenv("PFB_C_PPC") ? atoi(getenv("PFB_C_PPC")) : 0;
    if (ppc_env > 0) ppc = ppc_env;
    const int chunks = (npairs + ppc - 1) / ppc;
    LaunchScope ls(p, KC_FUSED_C, s);
    fused_ifftx_lcc_tma_kernel<N, NBUF><<<dim3(p->ny / 16, p->nz, chunks), 64, smem_c_tma<N>(NBUF), s>>>(
        p->tmapC, p->mbits, p->norm_factor, rot_index_offset + first, count, ppc, best, p->twdX, p->ny);
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

// back half: C (inverse x, LCC, running best) out of the work buffer X2
template <int N>
static int fused_back_n(Plan *p, int first, int count, int rot_index_offset, int64_t *best, const float2 *X2,
                        cudaStream_t s) {
    const int npairs = (count + 1) / 2;
    if constexpr (XCfg<N>::TMA) {
        switch (c_ring_depth()) {
            case 1: return fused_back_tma<N, 1>(p, first, count, rot_index_offset, best, X2, s);
            case 2: return fused_back_tma<N, 2>(p, first, count, rot_index_offset, best, X2, s);
            case 3: return fused_back_tma<N, 3>(p, first, count, rot_index_offset, best, X2, s);
            default: break;
        }
    }
    {
        // enough CTAs for ~4 waves of resident CTAs: split the pair loop into chunks
        constexpr int L = XCfg<N>::L, RT = XCfg<N>::RT;
        const int tiles = (p->ny / RT) * p->nz, per_sm = L == 8 ? 96 / RT : (N / L >= 24 ? 4 : 8);
        int chunks = std::max(1, std::min(npairs, (4 * p->sm_count * per_sm + tiles - 1) / tiles));
        int ppc = (npairs + chunks - 1) / chunks;
        static const int ppc_env = getenv("PFB_C_PPC") ? atoi(getenv("PFB_C_PPC")) : 0;
        if (ppc_env > 0) ppc = ppc_env;
        chunks = (npairs + ppc - 1) / ppc;
        LaunchScope ls(p, KC_FUSED_C, s);
        fused_ifftx_lcc_kernel<N, L, RT><<<dim3(p->ny / RT, p->nz, chunks), L * RT / 2, smem_c<N>(RT), s>>>(
            reinterpret_cast<const float4 *>(X2), p->mbits, p->norm_factor, rot_index_offset + first, count, ppc,
            best, p->twdX, p->ny, p->nz);
    }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

int fused_a(Plan *p, int first, int count, cudaStream_t s) {
    if (p->cls) return cls_a(p, first, count, s);
    return launch_fused_a(p, first, count, s);
}

int fused_b(Plan *p, int count, float2 *X2, cudaStream_t s) {
    if (p->cls) return cls_b(p, count, X2, s);
    return dispatch_zy(p->nz, p->ny, [&](auto nz, auto ny) {
        return fused_b_n<decltype(nz)::value, decltype(ny)::value>(p, count, X2, s);
    });
}

int fused_c(Plan *p, int first, int count, int rot_index_offset, int64_t *best, const float2 *X2, cudaStream_t s) {
    if (p->cls) return cls_c(p, first, count, rot_index_offset, best, X2, s);
    return dispatch_axis(p->nx, [&](auto nx) {
        return fused_back_n<decltype(nx)::value>(p, first, count, rot_index_offset, best, X2, s);
    });
}

// The whole rotation list, batch by batch.  Kernel A of batch i+1 (gather-latency bound, writes X1) runs on a
// side stream next to kernel C of batch i (an HBM stream out of X2): A(i+1) waits for B(i), the last reader of
// X1, and B(i+1) waits for A(i+1) and -- by stream order -- for C(i), the last reader of X2.  Per-kernel
// profiling (events around every launch) keeps everything on one stream so that the classes do not overlap.
int fused_scan(Plan *p, int R, int rot_index_offset, int64_t *best, cudaStream_t s) {
    // Off by default: measured at 128^3 the two-stream schedule is 0.6 % SLOWER (53.3k -> 52.9k rotations/s).
    // Kernel C's six CTAs per SM hold every register and 216 KB of shared memory, so kernel A's CTAs only get
    // on an SM when C's last wave drains -- there is no idle resource for the overlap to use.
    static const bool overlap_env = getenv("PFB_OVERLAP") ? atoi(getenv("PFB_OVERLAP")) != 0 : false;
    const bool overlap = overlap_env && !p->profile && p->side != nullptr && R > p->batch;
    int rc;
    // equal-size batches (even, so that rotation pairs never straddle two batches): a 927-rotation shard runs as
    // 4 x 232 rather than 3 x 256 + 159
    const int nbatch = (R + p->batch - 1) / p->batch;
    const int per = std::min(p->batch, (((R + nbatch - 1) / nbatch) + 1) & ~1);
    if (!overlap) {
        for (int first = 0; first < R; first += per) {
            const int count = std::min(per, R - first);
            if ((rc = fused_a(p, first, count, s))) return rc;
            if ((rc = fused_b(p, count, p->B, s))) return rc;
            if ((rc = fused_c(p, first, count, rot_index_offset, best, p->B, s))) return rc;
        }
        return PFB_OK;
    }
    // the side stream joins after everything already queued on s (rotation upload, template set-up)
    PFB_CUDA(cudaEventRecord(p->ev_b, s));
    for (int first = 0; first < R; first += per) {
        const int count = std::min(per, R - first);
        PFB_CUDA(cudaStreamWaitEvent(p->side, p->ev_b, 0));
        if ((rc = fused_a(p, first, count, p->side))) return rc;
        PFB_CUDA(cudaEventRecord(p->ev_a, p->side));
        PFB_CUDA(cudaStreamWaitEvent(s, p->ev_a, 0));
        if ((rc = fused_b(p, count, p->B, s))) return rc;
        PFB_CUDA(cudaEventRecord(p->ev_b, s));
        if ((rc = fused_c(p, first, count, rot_index_offset, best, p->B, s))) return rc;
    }
    return PFB_OK;
}

}  // namespace pfb
