// Class-decimated fused pipeline for cubic grids whose (y,z) plane does not fit one SM's shared
// memory (N = 192, 256).  Same three kernels as fused.cu,
//
//   A  cls_rotate_fftx   rotate + forward x + class fold        -> X1[pair][sig][z][kx][b][n]
//   B  cls_fftyz_mul     forward y,z * FT(map), inverse z,y of ONE ky class
//   C  cls_ifftx_lcc     class combine, inverse x, LCC, running best
//
// but kernel B works on a quarter of a (y,z) plane: with NB = N/64 and ky = NB k' + b the first
// radix-NB step of a decimation-in-frequency y transform splits the outputs into NB residue
// classes b, each a 64-point transform of the folded row
//   g_b[n] = W_N^(n b) sum_j x[n + 64 j] W_NB^(j b),   n < 64.
// The fold acts per (z, kx), so it commutes with the x transform: kernel A, whose tile holds
// the rows n + 64 j of a few n, applies it before it stores.  The z transforms, the
// multiplication with the map spectrum and the inverse z transforms act on every ky column on
// its own, so a CTA that owns class b of plane kx needs only the N x 64 tile of its class
// (132 KB at N = 256).  The inverse y transform of class b yields
//   G_b[n'] = sum_k' X[NB k' + b] W_64^(n' k'),   y[n' + 64 j] = sum_b W_NB^(j b) W_N^(n' b) G_b[n'],
// and that last radix-NB butterfly again acts per (z, kx) and commutes with the inverse x
// transform: kernel B stores G_b, and kernel C applies the butterfly to its tile in shared
// memory before the x pencils run.  HBM traffic stays at the three-kernel minimum (n_f S
// pruned in z + 3 S + 3 S per rotation).  (K numbers / reference lines: see fused.cu.)
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

template <int N> struct ClsCfg;
// LN x EN: column (z) pencils of kernel B; LC: lanes per x pencil of kernel C; PPT: y pairs per
// class in one kernel-C tile (tile = NB * PPT pencils = 2 NB PPT rows); CCTAS: kernel-C CTAs per SM
// LA / RN: lanes per x pencil and values of n per tile in kernel A (tile = NB * RN rows, NB * RN * LA threads);
// THREADS / CTAS: kernel B's CTA size and CTAs per SM
template <> struct ClsCfg<192> { static constexpr int LN = 8, EN = 24, THREADS = 128, CTAS = 2, NB = 3, PPT = 4, LC = 8, LA = 8, RN = 8, CCTAS = 2; };
template <> struct ClsCfg<256> { static constexpr int LN = 16, EN = 16, THREADS = 256, CTAS = 1, NB = 4, PPT = 2, LC = 16, LA = 16, RN = 4, CCTAS = 3; };

// ------------------------------------------------------------------------------- kernel A
// CTA = (z, tile of RN values of n, rotation pair); its NB RN rows are y = n + 64 j.  Gather
// (as fused_rotate_fftx_kernel), x transform with LA lanes per row, then per (kx, n) the
// radix-NB fold over j with the class twiddles, stored as y pairs (g_b[n], g_b[n+1]).
// (two CTAs per SM; budgeting the registers for three -- 80 instead of 126, a few spilled values -- measured 78 ->
// 109 us per rotation at 256^3)
template <int N>
__global__ void __launch_bounds__(ClsCfg<N>::NB * ClsCfg<N>::RN * ClsCfg<N>::LA, 2)
cls_rotate_fftx_kernel(const float4 *__restrict__ tmplq, const float *__restrict__ mask,
                       const double *__restrict__ rot, int first, int count, int nsig,
                       float2 *__restrict__ X1, const float2 *__restrict__ twN, int rs, int rs2,
                       unsigned nmask, int nzv) {
    constexpr int L = ClsCfg<N>::LA, E = N / L, NB = N / 64, RN = ClsCfg<N>::RN, ROWS = NB * RN, TP = ROWS + 1;
    constexpr int THREADS = ROWS * L;
    static_assert(THREADS % 32 == 0 && THREADS % (RN / 2) == 0, "whole warps, fixed pair slot per thread");
    extern __shared__ float2 smem[];
    float2 *tile_t = smem, *tile_m = smem + N * TP;
    const int pair = blockIdx.y;
    const int j = blockIdx.x % nzv;
    int nt_rank = blockIdx.x / nzv, ntile = 0;
    for (unsigned mbits = nmask;; ++ntile) {
        if (mbits & 1u) { if (nt_rank == 0) break; --nt_rank; }
        mbits >>= 1;
    }
    const int z = (j - rs + N) % N, n0 = RN * ntile;
    const GridDims d{N, N, N, N / 2, (long)N * N * N};
    const int ra = first + 2 * pair;
    const bool have_b = 2 * pair + 1 < count;
    const double *Ra = rot + (long)ra * 9, *Rb = Ra + 9;
    const int oz = z <= N / 2 ? z : z - N;

    // both tiles are one contiguous run of 2 N TP float2 (an even count): 16-byte stores
    for (int idx = threadIdx.x; idx < N * TP; idx += THREADS)
        reinterpret_cast<float4 *>(smem)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int xlo = max(-rs, -(N / 2 - 1));
    const int lim2 = min(rs2, (N / 2) * (N / 2));
    // rows are dealt to warps; a warp walks only the x span of its row that lies inside the sphere, with the (z, y)
    // part of the source coordinate -- the reference's per-row partial sums -- computed once per row.
    // (Measured alternative: the spans cut into 32-sample chunks dealt round robin, which balances the warps -- the
    // rows y = n + 64 j of a tile have very different spans -- but pays the row sums and a ballot per chunk:
    // 78.1 -> 83.6 us per rotation at 256^3.)
    {
        const double *Rb2 = have_b ? Rb : Ra;
        constexpr int LPR = 32, GROUPS = THREADS / LPR;
        const int gl = threadIdx.x % LPR;
        for (int rr = threadIdx.x / LPR; rr < ROWS; rr += GROUPS) {
            const int iy = n0 + rr % RN + 64 * (rr / RN);
            const int oy = iy <= N / 2 ? iy : iy - N;
            const int rem = lim2 - oy * oy - oz * oz;
            // offset -N/2 aliases index N/2, which belongs to offset +N/2
            if (rem < 0 || oy <= -(N / 2)) continue;
            const int hw = isqrt_floor(rem);
            const int lo = max(xlo, -hw), hi = min(rs, hw);
            const SrcCoord rowa = source_row(Ra, oy, oz), rowb = source_row(Rb2, oy, oz);
#pragma unroll 2
            for (int ox = lo + gl; ox <= hi; ox += LPR) {
                const SrcCoord ca = source_in_row(rowa, Ra, ox), cb = source_in_row(rowb, Rb2, ox);
                const float ta = sample_trilinear_q(tmplq, d, ca), tb = sample_trilinear_q(tmplq, d, cb);
                const float ma = sample_nearest(mask, d, ca), mb = sample_nearest(mask, d, cb);
                const int x = ox < 0 ? ox + N : ox;
                tile_t[x * TP + rr] = make_float2(ta, have_b ? tb : 0.f);
                tile_m[x * TP + rr] = make_float2(ma, have_b ? mb : 0.f);
            }
        }
    }
    __syncthreads();

    // ---- x transforms: thread (row rr, t)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & (L - 1), rr = (32 / L) * warp + lane / L;
    float2 tw[E];
    load_twiddles<E>(tw, twN, t);
    // A row outside the sphere (y = n + 64 j with j = NB / 2 always, the others away from the central z slabs)
    // was never written by the gather: it is zero, its transforms are zero, and a warp whose rows are all of that
    // kind skips its pencils (warp-uniform; the block barriers stay outside).
    bool live;
    {
        const int iy = n0 + rr % RN + 64 * (rr / RN);
        const int oy = iy <= N / 2 ? iy : iy - N;
        live = __any_sync(0xffffffffu, lim2 - oy * oy - oz * oz >= 0 && oy > -(N / 2));
    }
    float2 v[E];
    if (live) {
#pragma unroll
        for (int n1 = 0; n1 < E; ++n1) v[n1] = tile_t[(t + L * n1) * TP + rr];
        fft_pencil<E, L>(v, tile_t + rr, TP, t, tw, true);
#pragma unroll
        for (int m = 0; m < E; ++m) tile_t[(t + L * m) * TP + rr] = v[m];
    }
    __syncthreads();

    // ---- fold over j and store: thread (kx, pair slot ps): rows r = 2 ps, 2 ps + 1
    constexpr int H = N / 2, PS = RN / 2;
    const size_t slab = (size_t)N * H;
    float4 *X14 = reinterpret_cast<float4 *>(X1);
    const int ps = threadIdx.x % PS;
    float2 wf[2][NB - 1];                                    // W_N^((n0 + r) b), b = 1..NB-1
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int b = 1; b < NB; ++b) wf[e][b - 1] = __ldg(twN + ((n0 + 2 * ps + e) * b) % N);
    auto fold_store = [&](const float2 *tile, int sig) {
        float4 *o = X14 + ((size_t)(pair * nsig + sig) * N + z) * slab + n0 / 2 + ps;     // + kx*H + b*32
        for (int idx = threadIdx.x; idx < PS * N; idx += THREADS) {
            const int kx = idx / PS;
            float2 g[2][NB];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
#pragma unroll
                for (int jj = 0; jj < NB; ++jj) g[e][jj] = tile[kx * TP + 2 * ps + e + RN * jj];
                if (NB == 3) dft3(g[e][0], g[e][1], g[e][NB - 1]);
                else dft4(g[e][0], g[e][1], g[e][2], g[e][NB - 1]);
#pragma unroll
                for (int b = 1; b < NB; ++b) g[e][b] = cmulf(g[e][b], wf[e][b - 1]);
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                __stcs(o + (size_t)kx * H + b * 32, make_float4(g[0][b].x, g[1][b].x, g[0][b].y, g[1][b].y));
            }
        }
    };
    fold_store(tile_t, 0);
    // mask: its squares (core-weighted masks only) are parked in the template tile, which is free now
    __syncthreads();
    if (live) {
#pragma unroll
        for (int n1 = 0; n1 < E; ++n1) {
            v[n1] = tile_m[(t + L * n1) * TP + rr];
            if (nsig == 3) tile_t[(t + L * n1) * TP + rr] = make_float2(v[n1].x * v[n1].x, v[n1].y * v[n1].y);
        }
        fft_pencil<E, L>(v, tile_m + rr, TP, t, tw, true);
#pragma unroll
        for (int m = 0; m < E; ++m) tile_m[(t + L * m) * TP + rr] = v[m];
    }
    __syncthreads();
    fold_store(tile_m, 1);
    if (nsig == 3) {
        if (live) {
#pragma unroll
            for (int n1 = 0; n1 < E; ++n1) v[n1] = tile_t[(t + L * n1) * TP + rr];
            fft_pencil<E, L>(v, tile_t + rr, TP, t, tw, true);
#pragma unroll
            for (int m = 0; m < E; ++m) tile_t[(t + L * m) * TP + rr] = v[m];
        }
        __syncthreads();
        fold_store(tile_t, 2);
    }
}

// ------------------------------------------------------------------------------- kernel B
// Job j = (kx * npairs + pair) * NB + b, the three output planes gcc, ave, ave2 of one (kx, pair, class) in a row;
// the grid is a multiple of NB, so a persistent CTA keeps its class b = blockIdx.x % NB for all its jobs and
// neighbouring CTAs share the input rows (b fastest) and the map-spectrum tile (pair next).
// Binary mask with 8-lane column pencils (N = 192): as in fused_fftyz_mul_kernel, the forward spectrum of the mask
// is computed once, parked in tensor memory (tmem.cuh) and fetched back for the ave2 plane, which then has no
// phase 1 and no forward z.
// Tile layout: plane[z][c], c < 32, float4 = columns (k' = c, c + 32) in split form.
// (256^3: 256 threads at 232 registers; 512 threads at 128 registers measured 138 -> 158 us per rotation)
// STAGED (as in fused_fftyz_mul_kernel): the support rows of the CTA's next forward plane -- 512 bytes per z of its
// class -- are copied by TMA into a staging area behind the tile while phase 2 runs, two 4-D boxes (rows z = 0..rs
// and z = N-rs..N-1) issued by one thread; the row loop then reads them from shared memory instead of waiting for
// an L2 round trip per row with eight resident warps (17 % of the stall samples at 256^3,
// profiles/r02_ncu_v4_cls_256.txt).  Used when the staging area fits next to the tile (2 rs + 2 rows).
template <int N, bool STAGED>
__global__ void __launch_bounds__(ClsCfg<N>::THREADS, ClsCfg<N>::CTAS)
cls_fftyz_mul_kernel(const float4 *__restrict__ X1, float4 *__restrict__ X2, const float4 *__restrict__ Fc,
                     const float4 *__restrict__ F2c, const float2 *__restrict__ twN_g,
                     const float2 *__restrict__ twM_g, const float2 *__restrict__ twh_g,
                     int rs, unsigned nmask, int nsig, int npairs, const __grid_constant__ CUtensorMap tmapX1) {
    using Cfg = ClsCfg<N>;
    constexpr int H = N / 2, HC = 32, P = 33, NB = Cfg::NB, PPT = Cfg::PPT;
    constexpr int LN = Cfg::LN, EN = Cfg::EN, GN = 32 / LN;      // column pencils: N points
    constexpr int LM = 4, EM = 8, GM = 8;                        // row pencils: packed 32 points (64-point rows)
    constexpr int THREADS = Cfg::THREADS, NW = THREADS / 32;
    constexpr bool STASH = LN == 8;                              // the TMEM helpers move 8 packed pairs at a time
    constexpr int CIT = HC / GN / NW;                            // column groups per warp
    constexpr uint32_t TCOLS = 256;                              // CIT * 4 EN columns per lane, rounded up to a power of two
    static_assert(!STASH || (NW <= 4 && CIT * 4 * EN <= (int)TCOLS && TCOLS * Cfg::CTAS <= 512), "TMEM stash geometry");
    extern __shared__ float4 smem4[];
    float4 *plane = smem4;                                        // [N][P]
    float2 *twN = reinterpret_cast<float2 *>(plane + N * P);      // [EN][LN] W_N^(t k1)
    float2 *twM = twN + N;                                        // [EM][LM] W_32^(t k1)
    float2 *twh_s = twM + 32;                                     // [32] W_64^k of the split radix-2 step
    uint32_t *tslot = reinterpret_cast<uint32_t *>(twh_s + 32);
    uint64_t *sbar = reinterpret_cast<uint64_t *>(tslot + 2);     // staging copies have landed
    // staging rows: zi = z for z <= rs, zi = rs + 1 + (z - (N - rs)) for the rows below zero; 128-byte aligned
    float4 *stage = reinterpret_cast<float4 *>(reinterpret_cast<uint8_t *>(sbar + 1) + 112);
    const size_t slab = (size_t)N * H;                            // float4 per z of X1 / X2
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // row pencils: a quarter warp holds rows gM and gM + 4, whose storage is 64 bytes apart modulo
    // the 128-byte bank window (P = 33): conflict-free 16-byte accesses
    const int tM = lane & 3, gM = (lane >> 3) + 4 * ((lane >> 2) & 1);
    const int tN = lane & (LN - 1), gN = lane / LN;
    const int nzv = min(2 * rs + 1, N);
    const int njobs = N * npairs * NB;
    const int b = blockIdx.x % NB;
    const bool stash = STASH && nsig == 2;

    int job = blockIdx.x, vol = 0;      // plane whose phase 1 comes next
    int cjob = -1, cvol = 0;            // plane whose phase 2 is done (phase 3 pending)
    for (int i = threadIdx.x; i < N; i += THREADS) {        // column twiddles in pairs (fft_core.cuh: TwSmemPair)
        const int k1 = i / LN, tt = i % LN;
        twN[2 * ((k1 >> 1) * LN + tt) + (k1 & 1)] = twN_g[i];
    }
    for (int i = threadIdx.x; i < 32; i += THREADS) {
        twM[i] = twM_g[i];
        twh_s[i] = twh_g[i];
    }
    // one thread: both boxes of the class rows of plane (j, v) of X1 -> staging area
    auto stage_issue = [&](int j, int v) {
        const int jj = j / NB;
        const int pair = jj % npairs, kx = jj / npairs;
        const int sig = v == 0 ? 0 : (v == 1 ? 1 : nsig - 1);
        mbar_arrive_expect_tx(sbar, (uint32_t)(2 * (rs + 1) * HC * sizeof(float4)));
        tma_load_4d(stage, &tmapX1, sbar, 0, kx * NB + b, 0, pair * nsig + sig);
        tma_load_4d(stage + (rs + 1) * HC, &tmapX1, sbar, 0, kx * NB + b, N - rs, pair * nsig + sig);   // last row out of bounds: zeros
    };
    if (STAGED && threadIdx.x == 0) {
        mbar_init(sbar, 1);
        mbar_fence_init();
        if (job < njobs) stage_issue(job, 0);
    }
    uint32_t sphase = 0;
    uint32_t tcol = 0;
    if (STASH) {
        if (warp == 0) tmem_alloc(tslot, TCOLS);
        tmem_fence_before_sync();
        __syncthreads();
        tmem_fence_after_sync();
        tcol = *tslot + ((uint32_t)(32 * warp) << 16);           // NW <= 4: every warp its own lane quarter
    }

    while (true) {
        __syncthreads();          // phase 2 of the current plane is complete (first pass: the tables are in place)
        const bool fwd = job < njobs && !(stash && vol == 2);
        {
            // ---- row loop: phase 3 of plane (cjob, cvol) (inverse y of the class, shared -> HBM), then phase 1
            //      of plane (job, vol) (forward y of the folded rows inside the support box, HBM -> shared)
            float2 twr[EM], twh[EM];
#pragma unroll
            for (int m = 0; m < EM; ++m) { twr[m] = twM[m * LM + tM]; twh[m] = twh_s[tM + LM * m]; }
            const TwReg<EM> tw{twr};
            const float4 *src = X1;
            if (fwd) {
                const int jj = job / NB;
                const int pair = jj % npairs, kx = jj / npairs;
                const int sig = vol == 0 ? 0 : (vol == 1 ? 1 : nsig - 1);
                src = X1 + (size_t)(pair * nsig + sig) * N * slab + (size_t)kx * H + b * 32;   // + z*slab + n/2
            }
            float4 *dst = X2;
            if (cjob >= 0) {
                const int jj = cjob / NB;
                const int pair = jj % npairs, kx = jj / npairs;
                dst = X2 + (size_t)(pair * 3 + cvol) * N * slab + (size_t)kx * H;      // + z*slab + tile offset
            }
            if (STAGED && fwd) { mbar_wait(sbar, sphase); sphase ^= 1u; }
            for (int w = warp; w < N / GM; w += NW) {
                const int z = w * GM + gM;
                const bool act = fwd && (z + rs) % N < nzv;
                const bool any = __any_sync(0xffffffffu, act);
                // unstaged: the next plane's row z leaves for the registers before the inverse transform of the
                // current plane's row, which hides most of its L2 round trip (the class path has registers to spare)
                C2 vn[EM];
                if (!STAGED && any) {
#pragma unroll
                    for (int n1 = 0; n1 < EM; ++n1) {
                        const int idx = tM + LM * n1;                                  // n = 2 idx, 2 idx + 1
                        const bool have = act && ((nmask >> (2 * idx / Cfg::RN)) & 1u);
                        vn[n1] = have ? ldg_c2(src + (size_t)z * slab + idx) : c2_zero();
                    }
                }
                if (cjob >= 0) {
                    C2 v[EM];
#pragma unroll
                    for (int n1 = 0; n1 < EM; ++n1) v[n1] = lds_c2(plane + z * P + tM + LM * n1);
                    fft_row_split2adj<LM, EM>(v, plane + z * P, 1, tM, tw, twh);
                    // v[m] = (G_b[2k], G_b[2k+1]), k = tM + LM m: y pair k of class b, tile k / PPT
#pragma unroll
                    for (int m = 0; m < EM; ++m) {
                        const int k = tM + LM * m;
                        stg_c2(dst + (size_t)z * slab + ((k / PPT) * NB + b) * PPT + (k % PPT), v[m]);
                    }
                }
                if (any) {
                    if (STAGED) {
#pragma unroll
                        for (int n1 = 0; n1 < EM; ++n1) {
                            const int idx = tM + LM * n1;
                            const bool have = act && ((nmask >> (2 * idx / Cfg::RN)) & 1u);
                            const int zi = z <= rs ? z : z - (N - rs) + rs + 1;
                            vn[n1] = have ? lds_c2(stage + zi * HC + idx) : c2_zero();
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

        // pull the support rows of the CTA's next forward plane towards L2 while phase 2 runs (512 bytes per row
        // and class): the row loop's direct loads then see L2 instead of HBM latency
        {
            int jn = job, vn = vol + 1;
            if (vn == 3 || (stash && vn == 2)) { vn = 0; jn += gridDim.x; }
            if (STAGED) {
                // (a plane without phase 1 changes nothing: its successor's rows are already waiting)
                if (threadIdx.x == 0 && jn < njobs && fwd) {
                    fence_proxy_async_smem();
                    stage_issue(jn, vn);
                }
            } else if (jn < njobs) {
                const int jj = jn / NB;
                const int pair = jj % npairs, kx = jj / npairs;
                const int sig = vn == 0 ? 0 : (vn == 1 ? 1 : nsig - 1);
                const float4 *nxt = X1 + (size_t)(pair * nsig + sig) * N * slab + (size_t)kx * H + b * 32;
                for (int i = threadIdx.x; i < 4 * nzv; i += THREADS) {
                    const int z = ((i >> 2) - rs + N) % N;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + (size_t)z * slab + 8 * (i & 3)));
                }
            }
        }

        // ---- phase 2: forward z, multiply with the map spectrum, inverse z (column pairs k', k' + 32)
        {
            const int kx = job / NB / npairs;
            const float4 *Fm = (vol == 2 ? F2c : Fc) + (size_t)(kx * NB + b) * HC * N;   // + c*N + kz
            const TwSmemPair<LN> tw{reinterpret_cast<const float4 *>(twN) + tN};
#pragma unroll 1
            for (int it = 0; it < CIT; ++it) {
                const int c = (warp + it * NW) * GN + gN;
                C2 v[EN];
                if (fwd) {
#pragma unroll
                    for (int n1 = 0; n1 < EN; ++n1) {
                        const int z = tN + LN * n1;
                        const int sz = z <= N / 2 ? z : z - N;
                        v[n1] = (sz >= -rs && sz <= rs) ? lds_c2(plane + z * P + c) : c2_zero();
                    }
                }
                if constexpr (STASH) {
                    if (fwd) {
                        fft_pencil2_mul_stash<false, LN, EN>(v, plane + c, P, tN, tw, Fm + (size_t)c * N + tN,
                                                             tcol + it * 4 * EN, stash && vol == 1);
                    } else {
                        fft_pencil2_mul_stash<true, LN, EN>(v, plane + c, P, tN, tw, Fm + (size_t)c * N + tN,
                                                            tcol + it * 4 * EN, false);
                    }
                } else {
                    if (LN >= 16 && THREADS > 256) fft_pencil2_mul_late<LN, EN>(v, plane + c, P, tN, tw, Fm + (size_t)c * N + tN);
                    else fft_pencil2_mul<LN, EN>(v, plane + c, P, tN, tw, Fm + (size_t)c * N + tN);
                }
                fft_pencil2<LN, EN>(v, plane + c, P, tN, tw);
#pragma unroll
                for (int m = 0; m < EN; ++m) sts_c2(plane + (tN + LN * m) * P + c, v[m]);
            }
            if (STASH) tmem_wait_st();      // the parked spectrum is in place before this thread passes the next barrier
        }
        cjob = job; cvol = vol;
        if (++vol == 3) { vol = 0; job += gridDim.x; }
    }
    if (STASH) {
        if (warp == 0) tmem_dealloc(*tslot, TCOLS);
    }
}

// ------------------------------------------------------------------------------- kernel C
// One CTA owns tile u of plane z: for every kx the NB * PPT float4 that kernel B's classes wrote
// for y pairs u PPT .. u PPT + PPT - 1.  Per rotation pair and volume: cp.async the tile, apply
// the radix-NB class butterfly in place (slot (b, p) -> slot (j, p) = rows n' + 64 j, n' + 64 j + 1,
// n' = 2 (u PPT + p)), then L lanes transform one row pair along x and feed the epilogue of
// fused_ifftx_lcc_kernel (fused.cu): ave2 kept, 1/sqrt(var) when ave arrives, LCC and the
// running best when gcc arrives.
template <int N>
__global__ void __launch_bounds__(ClsCfg<N>::NB * ClsCfg<N>::PPT * ClsCfg<N>::LC, ClsCfg<N>::CCTAS)
cls_ifftx_lcc_kernel(const float4 *__restrict__ X2, const uint32_t *__restrict__ mbits,
                     const float4 *__restrict__ fold_g, float norm, int first_index, int count,
                     int pairs_per_chunk, int64_t *__restrict__ best, const float2 *__restrict__ twN_g) {
    using Cfg = ClsCfg<N>;
    constexpr int L = Cfg::LC, E = N / L, H = N / 2, NB = Cfg::NB, PPT = Cfg::PPT, NPC = NB * PPT;
    constexpr int TP = NPC + 1, RT = 2 * NPC, BP = N + 4, THREADS = NPC * L, Q = E / L;
    static_assert(THREADS % PPT == 0, "combine pass needs a fixed pair per thread");
    extern __shared__ float4 smem4[];
    float4 *tile = smem4;                                             // [N][TP]
    float2 *lbest = reinterpret_cast<float2 *>(tile + N * TP);        // [RT][BP] (lcc, rot index bits)
    float2 *tws = lbest + RT * BP;                                    // [E][L] W_N^(t k1)
    const int u = blockIdx.x, z = blockIdx.y;
    const int npairs = (count + 1) / 2;
    const int p0 = blockIdx.z * pairs_per_chunk, p1 = min(npairs, p0 + pairs_per_chunk);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & (L - 1), pc = (32 / L) * warp + lane / L;    // pencil = tile slot pc
    const int ya = 2 * (u * PPT + pc % PPT) + 64 * (pc / PPT);        // rows ya, ya + 1
    const size_t slab = (size_t)N * H;
    const size_t rowa = ((size_t)z * N + ya) * N, rowb = rowa + N;
    float2 *lba = lbest + (2 * pc) * BP + t, *lbb = lba + BP;
    // bit m of a row's word t: lcc_mask at x = t + L m
    const unsigned ma = mbits[((size_t)z * N + ya) * L + t], mb = mbits[((size_t)z * N + ya + 1) * L + t];
#pragma unroll
    for (int m = 0; m < E; ++m) {
        lba[L * m] = make_float2(0.f, 0.f);
        lbb[L * m] = make_float2(0.f, 0.f);
    }
    for (int i = threadIdx.x; i < N; i += THREADS) tws[i] = twN_g[i];
    const TwSmem<L> tw{tws + t};
    // class twiddles of this thread's column of the combine pass: W_N^(n' b), W_N^((n'+1) b)
    const int cpp = threadIdx.x % PPT;
    C2 twc[NB - 1];
#pragma unroll
    for (int bb = 1; bb < NB; ++bb) twc[bb - 1] = c2_from(__ldg(fold_g + bb * 32 + u * PPT + cpp));
    const int nitems = 3 * (p1 - p0);
    auto prefetch = [&](int item) {
        const int p = p0 + item / 3, vol = 2 - item % 3;              // ave2, ave, gcc
        const float4 *src = X2 + ((size_t)(p * 3 + vol) * N + z) * slab + u * NPC;
        for (int idx = threadIdx.x; idx < NPC * N; idx += THREADS) {
            const int kx = idx / NPC, c = idx % NPC;
            cp_async16(tile + kx * TP + c, src + (size_t)kx * H + c);
        }
        cp_async_commit();
    };
    if (nitems > 0) prefetch(0);
    C2 sd[E];
    for (int item = 0; item < nitems; ++item) {
        cp_async_wait<0>();
        __syncthreads();
        // ---- class butterfly, in place
        for (int idx = threadIdx.x; idx < PPT * N; idx += THREADS) {
            float4 *e = tile + (idx / PPT) * TP + cpp;
            C2 g[NB];
#pragma unroll
            for (int bb = 0; bb < NB; ++bb) g[bb] = lds_c2(e + bb * PPT);
#pragma unroll
            for (int bb = 1; bb < NB; ++bb) g[bb] = cmul(g[bb], twc[bb - 1]);
            if (NB == 3) dft3(g[0], g[1], g[NB - 1]);
            else dft4(g[0], g[1], g[2], g[NB - 1]);
#pragma unroll
            for (int bb = 0; bb < NB; ++bb) sts_c2(e + bb * PPT, g[bb]);
        }
        __syncthreads();
        const int p = p0 + item / 3, vi = item % 3;
        const uint32_t ia = (uint32_t)(first_index + 2 * p);
        const bool have_b = 2 * p + 1 < count;
        auto run_item = [&](auto vi_tag) {
            constexpr int VI = decltype(vi_tag)::value;
            {
                C2 v[E];
#pragma unroll
                for (int n1 = 0; n1 < E; ++n1) v[n1] = lds_c2(tile + (t + L * n1) * TP + pc);
                pencil2_stage1<L, E>(v, tile + pc, TP, t, tw);
            }
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                C2 a[L];
                pencil2_stage2<L>(a, tile + pc, TP, t, q);
                if (q == Q - 1) {
                    __syncthreads();       // every pencil is out of the tile: refill it while the arithmetic runs
                    if (item + 1 < nitems) prefetch(item + 1);
                }
#pragma unroll
                for (int k0 = 0; k0 < L; ++k0) {
                    const int m = q + Q * k0;
                    if (VI == 0) {
                        sd[m] = a[k0];                                     // ave2
                    } else if (VI == 1) {                                  // 1/sqrt(N ave2 - ave^2)
                        const float2 vr = psub(pmul(sd[m].re, pdup(norm)), pmul(a[k0].re, a[k0].re));
                        const float2 vq = psub(pmul(sd[m].im, pdup(norm)), pmul(a[k0].im, a[k0].im));
                        sd[m].re = make_float2(rsqrtf(vr.x), rsqrtf(vr.y));
                        sd[m].im = make_float2(rsqrtf(vq.x), rsqrtf(vq.y));
                    } else {
                        const float2 la = pmul(a[k0].re, sd[m].re), lb = pmul(a[k0].im, sd[m].im);
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
    }
#pragma unroll
    for (int m = 0; m < E; ++m) {
        if ((ma >> m) & 1u) {
            const float2 bv = lba[L * m];
            if (bv.x > 0.f)
                atomicMax(reinterpret_cast<long long *>(best + rowa + t + L * m),
                          (long long)pack_best(__float_as_uint(bv.x), __float_as_uint(bv.y)));
        }
        if ((mb >> m) & 1u) {
            const float2 bv = lbb[L * m];
            if (bv.x > 0.f)
                atomicMax(reinterpret_cast<long long *>(best + rowb + t + L * m),
                          (long long)pack_best(__float_as_uint(bv.x), __float_as_uint(bv.y)));
        }
    }
}

// ------------------------------------------------------------------------------- kernel C, TMA-fed
// Same work split as cls_ifftx_lcc_kernel, three changes (measured reasons in DESIGN.md):
//  * the tile arrives as tensor-map boxes (tma.cuh) issued by one thread -- at N = 256 one 128-byte-wide box
//    (NB PPT = 8 float4 per kx) in the 128-byte swizzle, at N = 192 one 64-byte-wide box per class in the
//    64-byte swizzle -- so the dense rows need no padding to be conflict free and no thread spends registers or
//    issue slots on 16-byte cp.async copies;
//  * 1/sqrt(var) of a thread's 2 rows x E x-values x 2 rotations (64 / 96 registers at N = 256 / 192) waits in
//    tensor memory between the ave2, ave and gcc tiles of a rotation pair instead of in registers (tmem.cuh:
//    the same thread stores and fetches, no layout involved): the old kernel needed 255 registers at N = 192;
//  * the running best of the chunk is kept as float LCC + 16-bit rotation offset (6 instead of 8 bytes per
//    voxel), which lets a third CTA fit next to the tiles.
template <int N> struct ClsC {
    using Cfg = ClsCfg<N>;
    static constexpr int NB = Cfg::NB, PPT = Cfg::PPT, NPC = NB * PPT, L = Cfg::LC, E = N / L, Q = E / L;
    static constexpr int W4 = NPC * 16 <= 128 ? NPC : PPT;        // float4 per row of one TMA box
    static constexpr int NBOX = NPC / W4;
    static constexpr int THREADS = NPC * L, RT = 2 * NPC, BP = N + 4, TILE = N * NPC, CTAS = 3;
    static constexpr uint32_t TCOLS = 4 * E <= 64 ? 64 : 128;      // TMEM columns per CTA: 4 E per lane, every warp its own lane quarter
    static_assert(W4 == 8 || W4 == 4, "128-byte or 64-byte swizzle");
    static_assert(THREADS <= 128, "at most four warps: one TMEM lane quarter each");
    // XOR pattern of the swizzle for a row whose low bits are k (plus the row's own offset bits below the box row stride)
    __host__ __device__ static constexpr int xc(int k) { return W4 == 8 ? (8 * (k & (L - 1)) | (k & 7)) : (4 * (k & 7) | ((k >> 1) & 3)); }
    static constexpr size_t smem() {
        return 1024 + (size_t)TILE * sizeof(float4) + (size_t)RT * BP * (sizeof(float) + sizeof(uint16_t)) + (size_t)N * sizeof(float2) +
               sizeof(uint64_t) + 16;
    }
};

template <int N>
__global__ void __launch_bounds__(ClsC<N>::THREADS, ClsC<N>::CTAS)
cls_ifftx_lcc_tma_kernel(const __grid_constant__ CUtensorMap tmap, const uint32_t *__restrict__ mbits,
                         const float4 *__restrict__ fold_g, float norm, int first_index, int count,
                         int pairs_per_chunk, int64_t *__restrict__ best, const float2 *__restrict__ twN_g) {
    using G = ClsC<N>;
    constexpr int L = G::L, E = G::E, NB = G::NB, PPT = G::PPT, NPC = G::NPC, W4 = G::W4, NBOX = G::NBOX;
    constexpr int RT = G::RT, BP = G::BP, THREADS = G::THREADS, Q = G::Q, TILE = G::TILE, RS = W4 * L;   // RS: float4 per L rows of a box
    extern __shared__ uint8_t smem_raw[];
    float4 *tile = reinterpret_cast<float4 *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));   // [NBOX][N][W4]
    float *lb_lcc = reinterpret_cast<float *>(tile + TILE);           // [RT][BP] best LCC of the chunk
    float2 *tws = reinterpret_cast<float2 *>(lb_lcc + RT * BP);       // [E][L] W_N^(t k1)
    uint64_t *full = reinterpret_cast<uint64_t *>(tws + N);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(full + 1);
    uint16_t *lb_rot = reinterpret_cast<uint16_t *>(tslot + 2);       // [RT][BP] its rotation, relative to the chunk's first
    const int u = blockIdx.x, z = blockIdx.y;
    const int npairs = (count + 1) / 2;
    const int p0 = blockIdx.z * pairs_per_chunk, p1 = min(npairs, p0 + pairs_per_chunk);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & (L - 1), pc = (32 / L) * warp + lane / L;    // pencil = tile slot pc = (class, pair)
    const int ya = 2 * (u * PPT + pc % PPT) + 64 * (pc / PPT);        // rows ya, ya + 1
    const size_t rowa = ((size_t)z * N + ya) * N, rowb = rowa + N;
    float *la = lb_lcc + (2 * pc) * BP + t, *lb = la + BP;
    uint16_t *ra = lb_rot + (2 * pc) * BP + t, *rb = ra + BP;
    // bit m of a row's word t: lcc_mask at x = t + L m
    const unsigned ma = mbits[((size_t)z * N + ya) * L + t], mb = mbits[((size_t)z * N + ya + 1) * L + t];
    // swizzled position of this thread's pencil: box, column c inside the box row, u0 = offset of row t
    const int box = pc / W4, c = pc % W4;
    const int u0 = box * (N * W4) + W4 * t + (c ^ (W4 == 8 ? (t & 7) : ((t >> 1) & 3)));
    const int nitems = 3 * (p1 - p0);
    auto issue = [&](int item) {                                      // thread 0 only
        const int p = p0 + item / 3, vol = 2 - item % 3;              // ave2, ave, gcc
        mbar_arrive_expect_tx(full, TILE * (uint32_t)sizeof(float4));
#pragma unroll
        for (int bx = 0; bx < NBOX; ++bx)
            tma_load_4d(tile + bx * (N * W4), &tmap, full, 4 * (u * NPC + bx * W4), 0, z, p * 3 + vol);
    };
    if (threadIdx.x == 0) {
        mbar_init(full, 1);
        mbar_fence_init();
        if (nitems > 0) issue(0);
    }
    if (warp == 0) tmem_alloc(tslot, G::TCOLS);
#pragma unroll
    for (int m = 0; m < E; ++m) {
        la[L * m] = 0.f; lb[L * m] = 0.f;
        ra[L * m] = 0; rb[L * m] = 0;
    }
    for (int i = threadIdx.x; i < N; i += THREADS) {        // twiddles in pairs (fft_core.cuh: TwSmemPair)
        const int k1 = i / L, tt = i % L;
        tws[2 * ((k1 >> 1) * L + tt) + (k1 & 1)] = twN_g[i];
    }
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tcol = *tslot + ((uint32_t)(32 * warp) << 16);
    const TwSmemPair<L> tw{reinterpret_cast<const float4 *>(tws) + t};
    // class twiddles of this thread's column of the combine pass: W_N^(n' b), W_N^((n'+1) b)
    const int cpp = W4 == 8 ? threadIdx.x / (THREADS / PPT) : threadIdx.x % PPT;
    C2 twc[NB - 1];
#pragma unroll
    for (int bb = 1; bb < NB; ++bb) twc[bb - 1] = c2_from(__ldg(fold_g + bb * 32 + u * PPT + cpp));
    const uint32_t rot0 = (uint32_t)(first_index + 2 * p0);
    for (int item = 0; item < nitems; ++item) {
        mbar_wait(full, (uint32_t)item & 1u);
        // ---- class butterfly, in place: slot (b, p) -> slot (j, p) = rows n' + 64 j, n' + 64 j + 1
        // (measured alternative: no separate pass, every pencil combining the NB class values of its elements on
        // the way into its registers -- NB tile reads per element against 1 + 2, one block barrier less: 22.4 ->
        // 23.4 us per rotation at 192^3, 45.4 -> 51.8 at 256^3; the kernel is bound by shared-memory wavefronts)
        if (W4 == 8) {
            // eight consecutive kx per quarter warp, one pair p: the swizzle spreads them over the eight 16-byte lanes
            for (int row = threadIdx.x % (THREADS / PPT); row < N; row += THREADS / PPT) {
                float4 *e = tile + row * 8;
                C2 g[NB];
#pragma unroll
                for (int bb = 0; bb < NB; ++bb) g[bb] = lds_c2(e + ((bb * PPT + cpp) ^ (row & 7)));
#pragma unroll
                for (int bb = 1; bb < NB; ++bb) g[bb] = cmul(g[bb], twc[bb - 1]);
                if (NB == 3) dft3(g[0], g[1], g[NB - 1]);
                else dft4(g[0], g[1], g[2], g[NB - 1]);
#pragma unroll
                for (int bb = 0; bb < NB; ++bb) sts_c2(e + ((bb * PPT + cpp) ^ (row & 7)), g[bb]);
            }
        } else {
            for (int row = threadIdx.x / PPT; row < N; row += THREADS / PPT) {
                float4 *e = tile + row * 4 + (cpp ^ ((row >> 1) & 3));
                C2 g[NB];
#pragma unroll
                for (int bb = 0; bb < NB; ++bb) g[bb] = lds_c2(e + bb * (N * 4));
#pragma unroll
                for (int bb = 1; bb < NB; ++bb) g[bb] = cmul(g[bb], twc[bb - 1]);
                if (NB == 3) dft3(g[0], g[1], g[NB - 1]);
                else dft4(g[0], g[1], g[2], g[NB - 1]);
#pragma unroll
                for (int bb = 0; bb < NB; ++bb) sts_c2(e + bb * (N * 4), g[bb]);
            }
        }
        __syncthreads();
        const int p = p0 + item / 3, vi = item % 3;
        const uint32_t ia = (uint32_t)(first_index + 2 * p);
        const bool have_b = 2 * p + 1 < count;
        // ONE body for the three volume kinds (VI = 0 ave2, 1 ave, 2 gcc; warp-uniform): 1/sqrt(var) waits in TMEM,
        // not in registers, so nothing ties a kind to its own code -- and three unrolled copies of the transform
        // (4 096 instructions at N = 192) showed up as instruction-fetch stalls (no_instruction 0.55)
        {
            const int VI = vi;
            {
                C2 v[E];
#pragma unroll
                for (int n1 = 0; n1 < E; ++n1) v[n1] = lds_c2(tile + RS * n1 + u0);
                DftReg<E, C2>::run(v);
#pragma unroll
                for (int k = 0; k < E / 2; ++k) {
                    const float4 w = tw.pair(k);
                    if (k > 0) v[2 * k] = cmulw(v[2 * k], make_float2(w.x, w.y));
                    v[2 * k + 1] = cmulw(v[2 * k + 1], make_float2(w.z, w.w));
                }
                __syncwarp();
#pragma unroll
                for (int k1 = 0; k1 < E; ++k1) sts_c2(tile + RS * k1 + (u0 ^ G::xc(k1 & (L - 1))), v[k1]);
                __syncwarp();
            }
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                C2 a[L];
#pragma unroll
                for (int n0 = 0; n0 < L; ++n0) a[n0] = lds_c2(tile + RS * L * q + RS * t + (u0 ^ G::xc(n0)));
                DftReg<L, C2>::run(a);
                if (q == Q - 1) {
                    __syncthreads();       // every pencil is out of the tile: hand it back to the copy engine
                    if (threadIdx.x == 0 && item + 1 < nitems) {
                        fence_proxy_async_smem();
                        issue(item + 1);
                    }
                }
                // 1/sqrt(var) of output m = q + Q k0 lives in TMEM columns 4 (L q + k0) .. + 3 of this lane
#pragma unroll
                for (int h = 0; h < L / 8; ++h) {
                    const uint32_t tc = tcol + 4 * (L * q + 8 * h);
                    C2 sd[8];
                    if (VI == 0) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) sd[k] = a[8 * h + k];                  // ave2
                        tmem_st_c2x8(tc, sd);
                    } else {
                        tmem_ld_c2x8(tc, sd);
                        if (VI == 1) {                                                     // 1/sqrt(N ave2 - ave^2)
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const C2 av = a[8 * h + k];
                                const float2 vr = psub(pmul(sd[k].re, pdup(norm)), pmul(av.re, av.re));
                                const float2 vq = psub(pmul(sd[k].im, pdup(norm)), pmul(av.im, av.im));
                                sd[k].re = make_float2(rsqrtf(vr.x), rsqrtf(vr.y));
                                sd[k].im = make_float2(rsqrtf(vq.x), rsqrtf(vq.y));
                            }
                            tmem_st_c2x8(tc, sd);
                        } else {
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const int m = q + Q * (8 * h + k);
                                const C2 gv = a[8 * h + k];
                                const float2 va = pmul(gv.re, sd[k].re), vb = pmul(gv.im, sd[k].im);
                                const bool sa = have_b && (vb.x > va.x || !(va.x == va.x));
                                const bool sb = have_b && (vb.y > va.y || !(va.y == va.y));
                                const float ca = sa ? vb.x : va.x, cb = sb ? vb.y : va.y;          // NaN never passes '>'
                                const uint16_t ja = (uint16_t)(ia - rot0 + (sa ? 1u : 0u)), jb = (uint16_t)(ia - rot0 + (sb ? 1u : 0u));
                                // unconditional read, predicated stores (see fused_ifftx_lcc_tma_kernel)
                                const float cura = la[L * m], curb = lb[L * m];
                                const bool ua = ((ma >> m) & 1u) != 0 && ca > cura, ub = ((mb >> m) & 1u) != 0 && cb > curb;
                                if (ua) { la[L * m] = ca; ra[L * m] = ja; }
                                if (ub) { lb[L * m] = cb; rb[L * m] = jb; }
                            }
                        }
                    }
                }
            }
            if (VI != 2) tmem_wait_st();
        }
    }
#pragma unroll
    for (int m = 0; m < E; ++m) {
        if ((ma >> m) & 1u) {
            const float bv = la[L * m];
            if (bv > 0.f)
                atomicMax(reinterpret_cast<long long *>(best + rowa + t + L * m),
                          (long long)pack_best(__float_as_uint(bv), rot0 + ra[L * m]));
        }
        if ((mb >> m) & 1u) {
            const float bv = lb[L * m];
            if (bv > 0.f)
                atomicMax(reinterpret_cast<long long *>(best + rowb + t + L * m),
                          (long long)pack_best(__float_as_uint(bv), rot0 + rb[L * m]));
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(*tslot, G::TCOLS);
}

// ------------------------------------------------------------------------------- helpers
// Fc[kx][b][c][kz] = (re F[kz][ky0][kx], re F[kz][ky1][kx], im .., im ..), ky0 = NB c + b,
// ky1 = NB (c + 32) + b: the map spectrum in kernel B's class / column-pair order
__global__ void cls_spectrum_kernel(const float2 *__restrict__ F, float4 *__restrict__ Fc, int N, int NB) {
    const size_t total = (size_t)N * NB * 32 * N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int kz = (int)(i % N);
        const int c = (int)((i / N) % 32);
        const int b = (int)((i / ((size_t)N * 32)) % NB);
        const int kx = (int)(i / ((size_t)N * 32 * NB));
        const float2 a = F[((size_t)kz * N + NB * c + b) * N + kx];
        const float2 e = F[((size_t)kz * N + NB * (c + 32) + b) * N + kx];
        Fc[i] = make_float4(a.x, e.x, a.y, e.y);
    }
}

// mbits[row * L + t], row = z*N + y: bit m = (lcc_mask[row][t + L m] != 0) -- kernel C's lane layout
__global__ void cls_mask_bits_kernel(const uint8_t *__restrict__ lcc_mask, uint32_t *__restrict__ mbits, int N, int L,
                                     long rows) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * L) return;
    const long row = i / L;
    const int t = (int)(i % L);
    uint32_t w = 0;
    for (int m = 0; m < N / L; ++m)
        if (lcc_mask[row * N + t + L * m]) w |= 1u << m;
    mbits[i] = w;
}

// ------------------------------------------------------------------------------- host side
template <int N> static constexpr size_t smem_a_cls() {
    return (size_t)2 * N * (ClsCfg<N>::RN * (N / 64) + 1) * sizeof(float2);
}
// tile + twiddle tables + the TMEM base address slot and the staging barrier (padded to 128 bytes) + `srows`
// staging rows of 32 float4 (0 = unstaged kernel)
template <int N> static constexpr size_t smem_b_cls(int srows = 0) {
    return (size_t)N * 33 * sizeof(float4) + (size_t)(N + 64) * sizeof(float2) + 128 + (size_t)srows * 32 * sizeof(float4);
}
template <int N> static constexpr int stage_capacity_cls() {
    return (int)((227 * 1024 / ClsCfg<N>::CTAS - 1024 - smem_b_cls<N>(0)) / (32 * sizeof(float4)));
}
template <int N> static constexpr size_t smem_c_cls() {
    using Cfg = ClsCfg<N>;
    return (size_t)N * (Cfg::NB * Cfg::PPT + 1) * sizeof(float4) + (size_t)2 * Cfg::NB * Cfg::PPT * (N + 4) * sizeof(float2) +
           (size_t)N * sizeof(float2);
}

static int upload_table(const std::vector<float> &h, void **out) {
    PFB_CUDA(cudaMalloc(out, sizeof(float) * h.size()));
    PFB_CUDA(cudaMemcpy(*out, h.data(), sizeof(float) * h.size(), cudaMemcpyHostToDevice));
    return PFB_OK;
}

// [k1][t] = exp(+2 pi i t k1 / (lanes e))
static std::vector<float> pencil_table(int lanes, int e) {
    std::vector<float> h((size_t)2 * lanes * e);
    const double n = (double)lanes * e;
    for (int t = 0; t < lanes; ++t)
        for (int k1 = 0; k1 < e; ++k1) {
            const double a = 2.0 * M_PI * (double)(t * k1) / n;
            h[2 * ((size_t)k1 * lanes + t)] = (float)cos(a);
            h[2 * ((size_t)k1 * lanes + t) + 1] = (float)sin(a);
        }
    return h;
}

template <int N> static int cls_init_n(Plan *p) {
    using Cfg = ClsCfg<N>;
    int rc;
    if ((rc = upload_table(pencil_table(Cfg::LN, Cfg::EN), (void **)&p->cls_twN))) return rc;
    if ((rc = upload_table(pencil_table(4, 8), (void **)&p->cls_twM))) return rc;
    std::vector<float> h(64), f((size_t)Cfg::NB * 32 * 4);
    for (int k = 0; k < 32; ++k) {
        const double a = 2.0 * M_PI * k / 64.0;
        h[2 * k] = (float)cos(a);
        h[2 * k + 1] = (float)sin(a);
    }
    for (int b = 0; b < Cfg::NB; ++b)
        for (int i = 0; i < 32; ++i) {
            const double a0 = 2.0 * M_PI * (double)((2 * i) * b) / N, a1 = 2.0 * M_PI * (double)((2 * i + 1) * b) / N;
            float *o = &f[((size_t)b * 32 + i) * 4];
            o[0] = (float)cos(a0); o[1] = (float)cos(a1); o[2] = (float)sin(a0); o[3] = (float)sin(a1);
        }
    if ((rc = upload_table(h, (void **)&p->cls_twh))) return rc;
    if ((rc = upload_table(f, (void **)&p->cls_fold))) return rc;
    PFB_CUDA(cudaFuncSetAttribute(cls_rotate_fftx_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem_a_cls<N>()));
    PFB_CUDA(cudaFuncSetAttribute(cls_ifftx_lcc_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem_c_cls<N>()));
    PFB_CUDA(cudaFuncSetAttribute(cls_ifftx_lcc_tma_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)ClsC<N>::smem()));
    return PFB_OK;
}

template <int N> static int cls_init_all(Plan *p) {
    int rc = cls_init_n<N>(p);
    if (rc) return rc;
    PFB_CUDA(cudaFuncSetAttribute(cls_fftyz_mul_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem_b_cls<N>(0)));
    PFB_CUDA(cudaFuncSetAttribute(cls_fftyz_mul_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem_b_cls<N>(std::max(0, stage_capacity_cls<N>()))));
    return PFB_OK;
}

int cls_init(Plan *p) { return p->nx == 192 ? cls_init_all<192>(p) : cls_init_all<256>(p); }

int cls_prepare_target(Plan *p, cudaStream_t s) {
    const int N = p->nx, NB = N / 64, L = N == 192 ? ClsCfg<192>::LC : ClsCfg<256>::LC;
    { LaunchScope ls(p, KC_OTHER, s);
      cls_spectrum_kernel<<<p->sm_count * 8, 256, 0, s>>>(p->F, reinterpret_cast<float4 *>(p->Fq), N, NB); }
    { LaunchScope ls(p, KC_OTHER, s);
      cls_spectrum_kernel<<<p->sm_count * 8, 256, 0, s>>>(p->F2, reinterpret_cast<float4 *>(p->F2q), N, NB); }
    { LaunchScope ls(p, KC_OTHER, s);
      const long rows = (long)N * N;
      cls_mask_bits_kernel<<<(unsigned)((rows * L + 255) / 256), 256, 0, s>>>(p->lcc_mask, p->mbits, N, L, rows); }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

template <int N> static int cls_b_n(Plan *p, int count, float2 *X2, cudaStream_t s) {
    using Cfg = ClsCfg<N>;
    const int npairs = (count + 1) / 2;
    const int njobs = N * npairs * Cfg::NB;
    int grid = p->sm_count * Cfg::CTAS;
    grid -= grid % Cfg::NB;
    grid = std::min(grid, njobs);
    static const int stage_env = getenv("PFB_B_STAGE") ? atoi(getenv("PFB_B_STAGE")) : 1;
    const int srows = 2 * p->rs + 2;
    const bool staged = stage_env != 0 && 2 * p->rs + 1 < N && srows <= stage_capacity_cls<N>();
    if (staged && (p->tmapB_base != (const void *)p->A || p->tmapB_rs != p->rs || p->tmapB_nsig != p->nsig)) {
        // X1 as [pair*nsig+sig][z][kx*NB+b][128 floats]; box = one class row x (rs + 1) consecutive z
        int rc = make_x1_tensor_map(&p->tmapB, p->A, 128, N * Cfg::NB, N, (long)p->nsig * (p->batch / 2), p->rs + 1);
        if (rc) return rc;
        p->tmapB_base = p->A; p->tmapB_rs = p->rs; p->tmapB_nsig = p->nsig;
    }
    LaunchScope ls(p, KC_FUSED_B, s);
    auto launch = [&](auto kernel, size_t smem) {
        kernel<<<grid, Cfg::THREADS, smem, s>>>(
            reinterpret_cast<const float4 *>(p->A), reinterpret_cast<float4 *>(X2),
            reinterpret_cast<const float4 *>(p->Fq), reinterpret_cast<const float4 *>(p->F2q), p->cls_twN, p->cls_twM,
            p->cls_twh, p->rs, p->nmask, p->nsig, npairs, p->tmapB);
    };
    if (staged) launch(cls_fftyz_mul_kernel<N, true>, smem_b_cls<N>(srows));
    else launch(cls_fftyz_mul_kernel<N, false>, smem_b_cls<N>(0));
    return PFB_OK;
}

template <int N> static int cls_c_n(Plan *p, int first, int count, int rot_index_offset, int64_t *best,
                                    const float2 *X2, cudaStream_t s) {
    using Cfg = ClsCfg<N>;
    constexpr int NPC = Cfg::NB * Cfg::PPT;
    const int npairs = (count + 1) / 2;
    const int tiles = (N / 2 / NPC) * N;
    static const bool tma_env = getenv("PFB_CLS_C_TMA") ? atoi(getenv("PFB_CLS_C_TMA")) != 0 : true;
    const int per_sm = tma_env ? ClsC<N>::CTAS : Cfg::CCTAS;
    // whole waves of resident CTAs, at least 4 pairs per CTA (see fused_back_tma in fused.cu)
    const int slots = p->sm_count * per_sm;
    int ppc = std::max(1, npairs);
    double best_eff = -1.0;
    for (int cand = std::min(npairs, 32); cand >= std::min(npairs, 4); --cand) {
        const double waves = (double)tiles * ((npairs + cand - 1) / cand) / slots;
        const double eff = waves / std::ceil(waves);
        if (waves >= 3.0 && eff > best_eff + 1e-9) { best_eff = eff; ppc = cand; }
    }
    if (best_eff < 0) ppc = std::max(1, std::min(npairs, 4));
    static const int ppc_env = getenv("PFB_C_PPC") ? atoi(getenv("PFB_C_PPC")) : 0;
    if (ppc_env > 0) ppc = ppc_env;
    ppc = std::min(ppc, 32000);                   // 16-bit rotation offsets inside a chunk
    const int chunks = (npairs + ppc - 1) / ppc;
    if (tma_env) {
        if (p->tmapC_base != (const void *)X2) {
            // X2 as [pair*3+vol][z][kx][2N floats]; box = W4 float4 x all kx (W4 = a tile row, or one class of it)
            int rc = make_x2_tensor_map(&p->tmapC, X2, 2 * N, N, N, 3L * (p->batch / 2), 4 * ClsC<N>::W4, N);
            if (rc) return rc;
            p->tmapC_base = X2;
        }
        LaunchScope ls(p, KC_FUSED_C, s);
        cls_ifftx_lcc_tma_kernel<N><<<dim3(N / 2 / NPC, N, chunks), ClsC<N>::THREADS, ClsC<N>::smem(), s>>>(
            p->tmapC, p->mbits, p->cls_fold, p->norm_factor, rot_index_offset + first, count, ppc, best, p->cls_twN);
        return PFB_OK;
    }
    LaunchScope ls(p, KC_FUSED_C, s);
    cls_ifftx_lcc_kernel<N><<<dim3(N / 2 / NPC, N, chunks), NPC * Cfg::LC, smem_c_cls<N>(), s>>>(
        reinterpret_cast<const float4 *>(X2), p->mbits, p->cls_fold, p->norm_factor, rot_index_offset + first, count,
        ppc, best, p->cls_twN);
    return PFB_OK;
}

template <int N> static int cls_a_n(Plan *p, int first, int count, cudaStream_t s) {
    const int npairs = (count + 1) / 2;
    const int nzv = std::min(2 * p->rs + 1, N);
    const int ntl = __builtin_popcount(p->nmask);
    LaunchScope ls(p, KC_FUSED_A, s);
    cls_rotate_fftx_kernel<N><<<dim3(nzv * ntl, npairs), ClsCfg<N>::NB * ClsCfg<N>::RN * ClsCfg<N>::LA, smem_a_cls<N>(), s>>>(
        p->tmplq, p->mask, p->rot_dev, first, count, p->nsig, p->A, p->tw[0], p->rs, p->rs2, p->nmask, nzv);
    return PFB_OK;
}

int cls_a(Plan *p, int first, int count, cudaStream_t s) {
    int rc = p->nx == 192 ? cls_a_n<192>(p, first, count, s) : cls_a_n<256>(p, first, count, s);
    if (rc) return rc;
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

int cls_b(Plan *p, int count, float2 *X2, cudaStream_t s) {
    int rc = p->nx == 192 ? cls_b_n<192>(p, count, X2, s) : cls_b_n<256>(p, count, X2, s);
    if (rc) return rc;
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

int cls_c(Plan *p, int first, int count, int rot_index_offset, int64_t *best, const float2 *X2, cudaStream_t s) {
    int rc = p->nx == 192 ? cls_c_n<192>(p, first, count, rot_index_offset, best, X2, s)
                          : cls_c_n<256>(p, first, count, rot_index_offset, best, X2, s);
    if (rc) return rc;
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

}  // namespace pfb
