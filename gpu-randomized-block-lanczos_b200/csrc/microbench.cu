// Pipe-rate and bandwidth micro-benchmarks (design evidence for DESIGN.md section 5; not on the solve path).
//   which 0: streaming copy GB/s (read+write)      1: streaming read GB/s
//         2: FP32 FFMA TFLOP/s                      3: FP64 DFMA TFLOP/s
//         4: mma.sync m16n8k8 tf32 TFLOP/s          5: mma.sync m8n8k4 f64 TFLOP/s
//         6: mma.sync m16n8k16 f16 TFLOP/s          7: mma.sync m16n8k16 bf16 TFLOP/s
#include "kernels.h"

#include <cstdio>

namespace rbl {

__global__ void mb_copy_kernel(const float4* __restrict__ src, float4* __restrict__ dst, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        float4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < n; i += stride) dst[i] = src[i];
}

__global__ void mb_read_kernel(const float4* __restrict__ src, float* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float s = 0.f;
    for (; i + 3 * stride < n; i += 4 * stride) {
        float4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
        s += a.x + b.y + c.z + d.w;
    }
    for (; i < n; i += stride) s += src[i].x;
    if (s == 123.456f) out[0] = s;
}

template <typename T>
__global__ void mb_fma_kernel(T* out, int iters) {
    T a0 = (T)threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const T x = (T)1.0000001, y = (T)0.5;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = a0 * x + y; a1 = a1 * x + y; a2 = a2 * x + y; a3 = a3 * x + y;
            a4 = a4 * x + y; a5 = a5 * x + y; a6 = a6 * x + y; a7 = a7 * x + y;
        }
    }
    T s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == (T)-1) out[0] = s;
}

__global__ void mb_mma_tf32_kernel(float* out, int iters) {
    float c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 + 4, b1 = a0 + 5;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            asm volatile(
                "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                : "+f"(c[u][0]), "+f"(c[u][1]), "+f"(c[u][2]), "+f"(c[u][3])
                : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == -1.f) out[0] = s;
}

__global__ void mb_mma_f64_kernel(double* out, int iters) {
    double c[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0.0;
    double a = (double)threadIdx.x, b = a + 1.0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[u][0]), "+d"(c[u][1])
                         : "d"(a), "d"(b));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
    if (s == -1.0) out[0] = s;
}

template <int KIND>  // 0: f16, 1: bf16
__global__ void mb_mma_h_kernel(float* out, int iters) {
    float c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 + 4, b1 = a0 + 5;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (KIND == 0)
                asm volatile(
                    "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                    : "+f"(c[u][0]), "+f"(c[u][1]), "+f"(c[u][2]), "+f"(c[u][3])
                    : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile(
                    "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                    : "+f"(c[u][0]), "+f"(c[u][1]), "+f"(c[u][2]), "+f"(c[u][3])
                    : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == -1.f) out[0] = s;
}

double microbench(int which, int64_t size, int iters) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 0.f;
    double result = -1.0;
    if (which == 0 || which == 1) {
        size_t bytes = (size_t)size;
        size_t n4 = bytes / 16;
        float4 *a = nullptr, *b = nullptr;
        if (cudaMalloc(&a, n4 * 16) != cudaSuccess) return -1.0;
        if (cudaMalloc(&b, n4 * 16) != cudaSuccess) { cudaFree(a); return -1.0; }
        cudaMemset(a, 1, n4 * 16);
        cudaMemset(b, 0, n4 * 16);
        int grid = sms * 16;
        for (int w = 0; w < 2; ++w) {
            if (which == 0) mb_copy_kernel<<<grid, 256>>>(a, b, n4);
            else mb_read_kernel<<<grid, 256>>>(a, (float*)b, n4);
        }
        double best = 1e30;
        for (int it = 0; it < iters; ++it) {
            cudaEventRecord(e0);
            if (which == 0) mb_copy_kernel<<<grid, 256>>>(a, b, n4);
            else mb_read_kernel<<<grid, 256>>>(a, (float*)b, n4);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            best = ms < best ? ms : best;
        }
        double moved = (which == 0 ? 2.0 : 1.0) * (double)n4 * 16.0;
        result = moved / (best * 1e-3) / 1e9;
        cudaFree(a);
        cudaFree(b);
    } else {
        void* out = nullptr;
        cudaMalloc(&out, 64);
        const int grid = sms * 8, block = 256;
        const int inner = iters > 0 ? iters : 2000;
        double flops = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (which == 2) { mb_fma_kernel<float><<<grid, block>>>((float*)out, inner); flops = 2.0 * 64 * (double)inner * grid * block; }
            if (which == 3) { mb_fma_kernel<double><<<grid, block>>>((double*)out, inner); flops = 2.0 * 64 * (double)inner * grid * block; }
            if (which == 4) { mb_mma_tf32_kernel<<<grid, block>>>((float*)out, inner); flops = 2.0 * 16 * 8 * 8 * 4 * (double)inner * grid * (block / 32); }
            if (which == 6) { mb_mma_h_kernel<0><<<grid, block>>>((float*)out, inner); flops = 2.0 * 16 * 8 * 16 * 4 * (double)inner * grid * (block / 32); }
            if (which == 7) { mb_mma_h_kernel<1><<<grid, block>>>((float*)out, inner); flops = 2.0 * 16 * 8 * 16 * 4 * (double)inner * grid * (block / 32); }
            if (which == 5) { mb_mma_f64_kernel<<<grid, block>>>((double*)out, inner); flops = 2.0 * 8 * 8 * 4 * 4 * (double)inner * grid * (block / 32); }
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        result = flops / (ms * 1e-3) / 1e12;
        cudaFree(out);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (cudaGetLastError() != cudaSuccess) return -1.0;
    return result;
}

}  // namespace rbl
