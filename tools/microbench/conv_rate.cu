// Microbenchmark: issue rate of the conversion / XU-pipe instructions that bound the box-filter kernels
// (float32 -> float64, float64 -> float32, float -> int, MUFU) against FP64 adds and an integer-pipe
// float32 -> float64 widening, on sm_100a.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o conv_rate conv_rate.cu && ./conv_rate
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;
constexpr int CH = 8;   // independent chains per thread

// exact float32 -> float64 for normal numbers and zero with integer instructions only
__device__ __forceinline__ double widen_int(float f) {
    const unsigned b = __float_as_uint(f);
    const unsigned a = b & 0x7fffffffu;
    unsigned hi = (a >> 3) + (a >= 0x00800000u ? 0x38000000u : 0u);
    hi |= b & 0x80000000u;
    return __hiloint2double((int)hi, (int)(b << 29));
}

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, float a, double c) {
    float x[CH];
    double acc[CH];
    float facc[CH];
    int iacc[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { x[i] = threadIdx.x * 1e-3f + i + 1.0f; acc[i] = i; facc[i] = 0.0f; iacc[i] = 0; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            x[i] = __fadd_rn(x[i], a);                                            // every mode: one FADD
            if (MODE == 0) acc[i] = __dadd_rn(acc[i], c);                         // + DADD
            if (MODE == 1) acc[i] = __dadd_rn(acc[i], (double)x[i]);              // + F2F.F64.F32 + DADD
            if (MODE == 2) { acc[i] = __dadd_rn(acc[i], c); facc[i] = __fadd_rn(facc[i], (float)acc[i]); }   // + DADD + F2F.F32.F64 + FADD
            if (MODE == 3) iacc[i] += __float2int_ru(x[i]);                       // + F2I + IADD
            if (MODE == 4) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x[i])); facc[i] = __fadd_rn(facc[i], r); }  // + MUFU + FADD
            if (MODE == 5) acc[i] = __dadd_rn(acc[i], widen_int(x[i]));           // + integer widening + DADD
            if (MODE == 6) iacc[i] += __float_as_int(__fadd_ru(x[i], 8388608.0f)) - 0x4B000000;   // + FADD.RP + IADD (ceil without F2I)
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += acc[i] + (double)facc[i] + (double)iacc[i] + (double)x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double* out) {
    int sms, khz;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    dim3 grid(sms * 8), block(256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, block>>>(out, 1.0f, 0.5);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<grid, block>>>(out, 1.0f, 0.5);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    const double groups = (double)grid.x * block.x * ITERS * CH;          // one op group per chain step
    const double per_clk_sm = groups / (ms * 1e-3) / ((double)khz * 1e3) / sms;
    printf("%-52s %8.3f ms   %7.1f groups/clk/SM\n", name, ms, per_clk_sm);
}

int main() {
    double* out; cudaMalloc(&out, sizeof(double) * 148 * 8 * 256 * 2);
    run<0>("FADD + DADD", out);
    run<1>("FADD + F2F.F64.F32 + DADD", out);
    run<2>("FADD + DADD + F2F.F32.F64 + FADD", out);
    run<3>("FADD + F2I.CEIL + IADD", out);
    run<4>("FADD + MUFU.RSQ + FADD", out);
    run<5>("FADD + integer widening (6 ALU ops) + DADD", out);
    run<6>("FADD + FADD.RP magic ceil + IADD", out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s (SM clock from the device attribute; XU-limited rows give the XU rate directly)\n", cudaGetErrorString(e));
    return 0;
}
