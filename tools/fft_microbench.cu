// Shared-memory-resident pencil FFT throughput (no HBM traffic): how fast can one SM run the
// packed / scalar 128-point pencils of fft_core.cuh at a given number of warps?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/fft_microbench tools/fft_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../powerfit_b200/csrc/fft_core.cuh"
using namespace pfb;

// MODE 0: packed column pencils (8 lanes x 16, float4 plane, stride P)   -- kernel B phase 2 / kernel C
// MODE 1: packed row pencils split2adj (8 lanes x 8)                     -- kernel B phases 1/3
// MODE 2: scalar pencils (8 lanes x 16, float2)                          -- round-1 kernels
template <int MODE, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) bench(float *out, const float2 *tw_g, int iters) {
    constexpr int N = 128, H = 64, P = H + 1;
    extern __shared__ float4 sm[];
    float4 *plane = sm;
    float2 *tws = reinterpret_cast<float2 *>(plane + (MODE == 2 ? 0 : N * P));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, t = lane & 7, g = lane >> 3;
    constexpr int NW = THREADS / 32;
    for (int i = threadIdx.x; i < N * P; i += THREADS) plane[i] = make_float4(i * 1e-4f, 1.f, -i * 1e-4f, 0.5f);
    if (MODE != 2) for (int i = threadIdx.x; i < N; i += THREADS) tws[i] = tw_g[i];
    __syncthreads();
    float acc = 0.f;
    if (MODE == 0) {
        const TwSmem<8> tw{tws + t};
        for (int it = 0; it < iters; ++it)
            for (int w = warp; w < H / 4; w += NW) {
                const int ky = w * 4 + g;
                C2 v[16];
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) v[n1] = lds_c2(plane + (t + 8 * n1) * P + ky);
                fft_pencil2<8, 16>(v, plane + ky, P, t, tw);
#pragma unroll
                for (int m = 0; m < 16; ++m) sts_c2(plane + (t + 8 * m) * P + ky, v[m]);
                acc += v[3].re.x;
            }
    } else if (MODE == 1) {
        float2 twr[8], twh[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) { twr[m] = tws[m * 8 + t]; twh[m] = tws[t + 8 * m]; }
        const TwReg<8> tw{twr};
        for (int it = 0; it < iters; ++it)
            for (int w = warp; w < N / 4; w += NW) {
                const int z = w * 4 + g;
                C2 v[8];
#pragma unroll
                for (int n1 = 0; n1 < 8; ++n1) v[n1] = lds_c2(plane + z * P + t + 8 * n1);
                fft_row_split2adj<8, 8>(v, plane + z * P, 1, t, tw, twh);
#pragma unroll
                for (int m = 0; m < 8; ++m) sts_c2(plane + z * P + t + 8 * m, v[m]);
                acc += v[3].re.x;
            }
    } else {
        float2 *pl2 = reinterpret_cast<float2 *>(sm);     // [128][129]
        float2 tw[16];
        load_twiddles<16>(tw, tw_g, t);
        for (int it = 0; it < iters; ++it)
            for (int w = warp; w < N / 4; w += NW) {
                const int ky = 32 * (w >> 3) + (w & 7) + 8 * g;
                float2 v[16];
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) v[n1] = pl2[(t + 8 * n1) * 129 + ky];
                fft_pencil<16>(v, pl2 + ky, 129, t, tw, true);
#pragma unroll
                for (int m = 0; m < 16; ++m) pl2[(t + 8 * m) * 129 + ky] = v[m];
                acc += v[3].x;
            }
    }
    out[blockIdx.x * THREADS + threadIdx.x] = acc;
}

template <int MODE, int THREADS, int MINB>
void run(const char *name, const float2 *tw, float *out, int ctas_per_sm) {
    const int iters = 200, sms = 148;
    const size_t smem = (MODE == 2 ? 128 * 129 * 8 : 128 * 65 * 16 + 1024);
    cudaFuncSetAttribute(bench<MODE, THREADS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    bench<MODE, THREADS, MINB><<<sms * ctas_per_sm, THREADS, smem>>>(out, tw, 2);
    cudaEventRecord(a);
    bench<MODE, THREADS, MINB><<<sms * ctas_per_sm, THREADS, smem>>>(out, tw, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    // 128-point pencils per CTA per iteration: MODE 0: 64 column pairs = 128; MODE 1: 128 rows; MODE 2: 128
    const double pencils = 128.0 * iters * sms * ctas_per_sm;
    const double lane_ops = pencils * 8 * 344.0;   // scalar-equivalent FP32 lane instructions per pencil
    printf("%-34s threads %4d x %d CTA/SM: %7.3f ms  %7.2f G pencils/s  %5.1f T lane-ops/s (%4.1f%% of 36 T)  err=%s\n", name, THREADS,
           ctas_per_sm, ms, pencils / ms * 1e-6, lane_ops / ms * 1e-9, lane_ops / ms * 1e-9 / 36.0 * 100, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float2 h[128];
    for (int k1 = 0; k1 < 16; ++k1) for (int t = 0; t < 8; ++t) { double a = 2 * M_PI * t * k1 / 128.0; h[k1 * 8 + t] = make_float2(cos(a), sin(a)); }
    float2 *tw; cudaMalloc(&tw, sizeof(h)); cudaMemcpy(tw, h, sizeof(h), cudaMemcpyHostToDevice);
    float *out; cudaMalloc(&out, 148 * 4 * 1024 * 4);
    run<0, 512, 1>("packed column 8x16", tw, out, 1);
    run<0, 256, 1>("packed column 8x16", tw, out, 1);
    run<0, 128, 1>("packed column 8x16", tw, out, 1);
    run<0, 128, 3>("packed column 8x16 (3 CTA)", tw, out, 3);
    run<1, 512, 1>("packed row split2adj 8x8", tw, out, 1);
    run<1, 1024, 1>("packed row split2adj 8x8", tw, out, 1);
    run<1, 256, 1>("packed row split2adj 8x8", tw, out, 1);
    run<2, 512, 1>("scalar column 8x16", tw, out, 1);
    run<2, 256, 1>("scalar column 8x16", tw, out, 1);
    return 0;
}
