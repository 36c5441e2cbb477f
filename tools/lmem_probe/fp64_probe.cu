// Stand-alone probe: does the FP64 sphere quadratic of the renderer (F2F, DMUL/DADD, double sqrt, F2F back) give the same
// answer twice in the same thread while kernels of other streams overlap?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -ftz=true -prec-div=true -prec-sqrt=true -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __noinline__ bool sphere(float ox, float oy, float oz, float dx, float dy, float dz, float radius, float& nearD, float& farD)
{
    const double radiusD = (double)radius;
    const double v = (double)(-(dx * ox + dy * oy + dz * oz));
    const double det = radiusD * radiusD - (double)(ox * ox + oy * oy + oz * oz) + v * v;
    if (det <= 0.0) return false;
    const double s = sqrt(det);
    nearD = (float)(v - s);
    farD = (float)(v + s);
    return farD > nearD;
}

__global__ void __launch_bounds__(128, 8) kernelA(const float* __restrict__ table, unsigned n, unsigned iters, unsigned long long* errors)
{
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned h = tid * 2654435761u + 99u, bad = 0;
    for (unsigned it = 0; it < iters; ++it)
    {
        h = h * 1664525u + 1013904223u;
        const float jitter = table[(h >> 8) % n];
        const float ox = -16.5f + 6.0f, oy = 2.2f - 1.1f, oz = 0.3f;
        const float dx = 0.994f + jitter * 1e-3f, dy = -0.054f + (float)(h & 255u) * 1e-4f, dz = -0.09f;
        float n1 = 0, f1 = 0, n2 = 0, f2 = 0;
        const bool a = sphere(ox, oy, oz, dx, dy, dz, 1.1f, n1, f1);
        const bool b = sphere(ox, oy, oz, dx, dy, dz, 1.1f, n2, f2);
        if (a != b || __float_as_uint(n1) != __float_as_uint(n2) || __float_as_uint(f1) != __float_as_uint(f2)) bad++;
    }
    if (bad) atomicAdd(errors, (unsigned long long)bad);
}

__global__ void __launch_bounds__(128, 6) kernelB(float* out, unsigned n, unsigned iters)
{
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    float x = (float)tid * 1e-6f, y = 0.5f;
    for (unsigned it = 0; it < iters; ++it) { x = x * 1.0001f + y; y = y * 0.9999f - x * 1e-7f; }
    out[tid % n] = x + y;
}

int main()
{
    const unsigned n = 1u << 20;
    float *table, *out; unsigned long long* errors;
    cudaMalloc(&table, n * 4); cudaMalloc(&out, n * 4); cudaMalloc(&errors, 8);
    cudaMemset(table, 0, n * 4);
    cudaStream_t s[3];
    for (auto& x : s) cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking);
    for (int mode = 0; mode < 2; ++mode)
    {
        cudaMemset(errors, 0, 8);
        cudaDeviceSynchronize();
        for (int rep = 0; rep < 30; ++rep)
        {
            kernelA<<<148 * 32, 128, 0, s[0]>>>(table, n, 300, errors);
            if (mode == 1) { kernelA<<<148 * 32, 128, 0, s[1]>>>(table, n, 200, errors); kernelB<<<148 * 24, 128, 0, s[2]>>>(out, n, 4000); }
        }
        cudaDeviceSynchronize();
        unsigned long long e = 0;
        cudaMemcpy(&e, errors, 8, cudaMemcpyDeviceToHost);
        printf("mode %d (%s): %llu sphere evaluations that differ from their own repetition, last error: %s\n", mode, mode ? "overlapping streams" : "one stream", e, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
