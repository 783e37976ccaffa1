// Operator-level access to the register / shared-memory FFT building blocks of the fused path (fft_core.cuh):
// pfb_pencil_fft runs ONE building block on caller data so that tests can pin each of them against numpy on its own
// (the search kernels only exercise them end to end).  Kernel exp(+2 pi i n k / N), un-normalised, like everything else.
#include "common.cuh"
#include "fft_core.cuh"

#include <cmath>
#include <type_traits>

namespace pfb {

// kind 0  packed pencil (fft_pencil2): element = two independent complex sequences; N = LANES * E points
// kind 1  row transform adjacent-in -> split-out (fft_row_adj2split): one sequence of 2 LANES E points,
//         in element n = (x[2n], x[2n+1]), out element k = (X[k], X[k + LANES E])
// kind 2  row transform split-in -> adjacent-out (fft_row_split2adj): in element n = (x[n], x[n + LANES E]),
//         out element k = (X[2k], X[2k+1])
// kind 3  scalar pencil (fft_pencil, LANES = 8): float2 elements, first half of every float4 is used
template <int LANES, int E, int KIND>
__global__ void pencil_op_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, int count) {
    constexpr int N = LANES * E, G = 32 / LANES;
    extern __shared__ float4 sm[];
    float4 *store = sm;                                            // [G][N]
    float2 *tws = reinterpret_cast<float2 *>(store + G * N);       // [E][LANES] W_N^(t k1)
    float2 *twh_s = tws + N;                                       // [N] W_2N^k
    const int lane = threadIdx.x, t = lane & (LANES - 1), g = lane / LANES;
    for (int i = lane; i < N; i += 32) {
        const int k1 = i / LANES, tt = i % LANES;
        double s, c;
        sincospi(2.0 * (double)(tt * k1) / (double)N, &s, &c);
        tws[i] = make_float2((float)c, (float)s);
        sincospi((double)i / (double)N, &s, &c);                   // exp(2 pi i k / 2N)
        twh_s[i] = make_float2((float)c, (float)s);
    }
    __syncwarp();
    const int p = blockIdx.x * G + g;
    const bool act = p < count;
    const float4 *src = in + (size_t)(act ? p : 0) * N;
    float4 *scratch = store + g * N;
    if constexpr (KIND == 3) {
        float2 v[E], tw[E];
#pragma unroll
        for (int k1 = 0; k1 < E; ++k1) tw[k1] = tws[k1 * LANES + t];
#pragma unroll
        for (int n1 = 0; n1 < E; ++n1) { const float4 x = src[t + LANES * n1]; v[n1] = make_float2(x.x, x.y); }
        fft_pencil<E, LANES>(v, reinterpret_cast<float2 *>(scratch), 1, t, tw, true);
        if (act) {
#pragma unroll
            for (int m = 0; m < E; ++m) out[(size_t)p * N + t + LANES * m] = make_float4(v[m].x, v[m].y, 0.f, 0.f);
        }
    } else {
        C2 v[E];
#pragma unroll
        for (int n1 = 0; n1 < E; ++n1) v[n1] = c2_from(src[t + LANES * n1]);
        const TwSmem<LANES> tw{tws + t};
        if constexpr (KIND == 0) {
            fft_pencil2<LANES, E>(v, scratch, 1, t, tw);
        } else {
            float2 twh[E];
#pragma unroll
            for (int m = 0; m < E; ++m) twh[m] = twh_s[t + LANES * m];
            if constexpr (KIND == 1) fft_row_adj2split<LANES, E>(v, scratch, 1, t, tw, twh);
            else fft_row_split2adj<LANES, E>(v, scratch, 1, t, tw, twh);
        }
        if (act) {
#pragma unroll
            for (int m = 0; m < E; ++m) out[(size_t)p * N + t + LANES * m] = c2_to(v[m]);
        }
    }
}

template <int LANES, int E>
static int launch_pencil_op(int kind, const float4 *in, float4 *out, int count, cudaStream_t s) {
    constexpr int N = LANES * E, G = 32 / LANES;
    const size_t smem = (size_t)G * N * sizeof(float4) + (size_t)2 * N * sizeof(float2);
    const int grid = (count + G - 1) / G;
    switch (kind) {
        case 0: pencil_op_kernel<LANES, E, 0><<<grid, 32, smem, s>>>(in, out, count); break;
        case 1: pencil_op_kernel<LANES, E, 1><<<grid, 32, smem, s>>>(in, out, count); break;
        case 2: pencil_op_kernel<LANES, E, 2><<<grid, 32, smem, s>>>(in, out, count); break;
        default: set_error("pfb_pencil_fft: unknown kind"); return PFB_ERR_INVALID;
    }
    PFB_CUDA(cudaGetLastError());
    return PFB_OK;
}

}  // namespace pfb

using namespace pfb;

extern "C" int pfb_pencil_fft(int kind, int lanes, int e, const float *in, float *out, int count, void *stream) {
    PFB_REQUIRE(in && out && count > 0, "pfb_pencil_fft: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    const float4 *i4 = reinterpret_cast<const float4 *>(in);
    float4 *o4 = reinterpret_cast<float4 *>(out);
    if (kind == 3) {
        // kernel A's scalar pencils: 8 x 8 (64), 8 x 16 (128), 4 x 8 (32), 4 x 24 (96)
        auto scalar = [&](auto lt, auto et) {
            constexpr int L = decltype(lt)::value, E = decltype(et)::value, N = L * E, G = 32 / L;
            pencil_op_kernel<L, E, 3><<<(count + G - 1) / G, 32, G * N * sizeof(float4) + 2 * N * sizeof(float2), s>>>(i4, o4, count);
        };
        if (lanes == 8 && e == 8) scalar(std::integral_constant<int, 8>{}, std::integral_constant<int, 8>{});
        else if (lanes == 8 && e == 16) scalar(std::integral_constant<int, 8>{}, std::integral_constant<int, 16>{});
        else if (lanes == 4 && e == 8) scalar(std::integral_constant<int, 4>{}, std::integral_constant<int, 8>{});
        else if (lanes == 4 && e == 24) scalar(std::integral_constant<int, 4>{}, std::integral_constant<int, 24>{});
        else { set_error("pfb_pencil_fft: scalar pencils are 8 x 8, 8 x 16, 4 x 8 or 4 x 24"); return PFB_ERR_INVALID; }
        PFB_CUDA(cudaGetLastError());
        return PFB_OK;
    }
    // the (lanes, points per lane) pairs the search kernels use
    if (lanes == 4 && e == 4) return launch_pencil_op<4, 4>(kind, i4, o4, count, s);      // 32-point rows (packed 16)
    if (lanes == 4 && e == 8) return launch_pencil_op<4, 8>(kind, i4, o4, count, s);      // 32-point pencils, 64-point rows (packed 32)
    if (lanes == 4 && e == 12) return launch_pencil_op<4, 12>(kind, i4, o4, count, s);    // 96-point rows (packed 48)
    if (lanes == 4 && e == 24) return launch_pencil_op<4, 24>(kind, i4, o4, count, s);    // 96-point pencils
    if (lanes == 8 && e == 8) return launch_pencil_op<8, 8>(kind, i4, o4, count, s);      // 64-point pencils, 128-point rows
    if (lanes == 8 && e == 16) return launch_pencil_op<8, 16>(kind, i4, o4, count, s);    // 128-point pencils
    if (lanes == 8 && e == 24) return launch_pencil_op<8, 24>(kind, i4, o4, count, s);    // 192-point pencils
    if (lanes == 16 && e == 16) return launch_pencil_op<16, 16>(kind, i4, o4, count, s);  // 256-point pencils
    set_error("pfb_pencil_fft: unsupported pencil geometry");
    return PFB_ERR_UNSUPPORTED;
}
