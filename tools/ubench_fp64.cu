// ubench_fp64.cu -- micro-benchmarks that pin the sm_100a numbers the stage kernel is designed
// against: FP64 pipe throughput / latency, SHFL and LDS/STS.64 throughput, and how well FP64
// issue overlaps with integer/move traffic.  Development tool (nvcc -arch=sm_100a, run on a B200).
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITER = 4096;

// ILP independent DFMA chains per thread
template <int ILP, int OP>
__global__ void k_fp64(double *out, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) x[i] = __fma_rn(x[i], a, b);
            else if (OP == 1) x[i] = __dadd_rn(x[i], b);
            else x[i] = __dmul_rn(x[i], a);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FP64 + integer ALU mix: per DFMA, NI integer ops (independent)
template <int NI>
__global__ void k_mix(double *out, double a, double b, int c)
{
    double x[8];
    int y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-9 + i; y[i] = threadIdx.x + i; }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i] = __fma_rn(x[i], a, b);
#pragma unroll
            for (int j = 0; j < NI; ++j) y[(i + j) & 7] = (y[(i + j) & 7] ^ c) + it;
        }
    }
    double s = 0;
    int t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += x[i]; t += y[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + t;
}

__global__ void k_shfl(double *out)
{
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __shfl_down_sync(0xffffffffu, x[i], 1);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int WIDTH> // 8 or 16 bytes per lane
__global__ void k_lds(double *out)
{
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    double s = 0;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (WIDTH == 8) {
                s += sm[((w * 8 + i) * 32 + lane + it) & 4095];
            } else {
                const double2 v = *reinterpret_cast<const double2 *>(&sm[(((w * 8 + i) * 32 + lane + it) * 2) & 4095]);
                s += v.x + v.y;
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_div(double *out, double a)
{
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = 1.0 + threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = a / x[i] + 1.0;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x[0] + x[1] + x[2] + x[3];
}

__global__ void k_sqrt(double *out, double a)
{
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = 1.0 + threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = sqrt(x[i]) + a;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x[0] + x[1] + x[2] + x[3];
}

template <typename F>
static float time_it(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, max clock %d MHz\n", p.name, p.multiProcessorCount, clk_khz / 1000);
    double *out; CK(cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024));
    const int sms = p.multiProcessorCount;
    const double ghz = clk_khz * 1e-6;
#define RUN_FP64(ILP, OP, WARPS, name) { \
        float ms = time_it([&] { k_fp64<ILP, OP><<<sms, WARPS * 32>>>(out, 1.0000001, 1e-9); }); \
        double per_sm_clk = (double) ITER * ILP * WARPS * 32 / (ms * 1e-3 * ghz * 1e9); \
        printf("%-6s ILP=%d warps/SM=%2d : %.3f ms  -> %.1f lane-ops/clk/SM (assuming %.3f GHz), chain latency<=%.1f clk\n", \
               name, ILP, WARPS, ms, per_sm_clk, ghz, ms * 1e-3 * ghz * 1e9 / ((double) ITER * ILP) * (WARPS <= 4 ? 1 : 0)); }
    RUN_FP64(1, 0, 4, "DFMA")  RUN_FP64(2, 0, 4, "DFMA")  RUN_FP64(4, 0, 4, "DFMA")  RUN_FP64(8, 0, 4, "DFMA")
    RUN_FP64(1, 0, 8, "DFMA")  RUN_FP64(2, 0, 8, "DFMA")  RUN_FP64(4, 0, 8, "DFMA")  RUN_FP64(8, 0, 8, "DFMA")
    RUN_FP64(1, 0, 12, "DFMA") RUN_FP64(2, 0, 12, "DFMA") RUN_FP64(4, 0, 12, "DFMA") RUN_FP64(8, 0, 12, "DFMA")
    RUN_FP64(1, 0, 16, "DFMA") RUN_FP64(2, 0, 16, "DFMA") RUN_FP64(4, 0, 16, "DFMA") RUN_FP64(4, 0, 32, "DFMA")
    RUN_FP64(4, 1, 12, "DADD") RUN_FP64(8, 1, 16, "DADD") RUN_FP64(4, 2, 12, "DMUL") RUN_FP64(8, 2, 16, "DMUL")
    RUN_FP64(1, 1, 4, "DADD")  RUN_FP64(1, 2, 4, "DMUL")
#define RUN_MIX(NI, WARPS) { \
        float ms = time_it([&] { k_mix<NI><<<sms, WARPS * 32>>>(out, 1.0000001, 1e-9, 12345); }); \
        double per_sm_clk = (double) ITER * 8 * WARPS * 32 / (ms * 1e-3 * ghz * 1e9); \
        printf("MIX    1 DFMA : %d int-pair, warps/SM=%2d : %.3f ms -> %.1f DFMA lanes/clk/SM\n", NI, WARPS, ms, per_sm_clk); }
    RUN_MIX(0, 16) RUN_MIX(1, 16) RUN_MIX(2, 16) RUN_MIX(3, 16) RUN_MIX(1, 12) RUN_MIX(2, 12)
    for (int warps : {4, 8, 16, 32}) {
        float ms = time_it([&] { k_shfl<<<sms, warps * 32>>>(out); });
        printf("SHFL.64 (2 SHFL each) warps/SM=%2d : %.3f ms -> %.2f SHFL warp-instr/clk/SM\n", warps, ms,
               (double) ITER * 8 * 2 * warps / (ms * 1e-3 * ghz * 1e9));
    }
    for (int warps : {4, 8, 16, 32}) {
        float ms = time_it([&] { k_lds<8><<<sms, warps * 32, 32768>>>(out); });
        printf("LDS.64  warps/SM=%2d : %.3f ms -> %.1f B/clk/SM\n", warps, ms, (double) ITER * 8 * warps * 32 * 8 / (ms * 1e-3 * ghz * 1e9));
        ms = time_it([&] { k_lds<16><<<sms, warps * 32, 32768>>>(out); });
        printf("LDS.128 warps/SM=%2d : %.3f ms -> %.1f B/clk/SM\n", warps, ms, (double) ITER * 8 * warps * 32 * 16 / (ms * 1e-3 * ghz * 1e9));
    }
    for (int warps : {4, 12, 16}) {
        float ms = time_it([&] { k_div<<<sms, warps * 32>>>(out, 3.0); });
        printf("DIV  warps/SM=%2d : %.3f ms -> %.2f div lanes/clk/SM\n", warps, ms, (double) (ITER / 8) * 4 * warps * 32 / (ms * 1e-3 * ghz * 1e9));
        ms = time_it([&] { k_sqrt<<<sms, warps * 32>>>(out, 3.0); });
        printf("SQRT warps/SM=%2d : %.3f ms -> %.2f sqrt lanes/clk/SM\n", warps, ms, (double) (ITER / 8) * 4 * warps * 32 / (ms * 1e-3 * ghz * 1e9));
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
