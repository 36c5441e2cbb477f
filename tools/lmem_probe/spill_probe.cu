// Stand-alone probe: are REGISTER SPILLS (compiler-generated STL / LDL, incl. LDL.LU) preserved while kernels of other
// streams overlap?  Kernel A holds 24 live values across a long, divergent loop of dependent global loads under a
// 40-register budget (ptxas spills about half of them) and checks them at the end against a recomputation.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Xptxas -v -o spill_probe spill_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

#define NV 24

__global__ void __launch_bounds__(128, 12) kernelA(const unsigned* __restrict__ table, const unsigned* init, unsigned n, unsigned iters, unsigned long long* errors, unsigned* sink)
{
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned v[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = init[((size_t)tid * NV + k) % n];      // 24 values that stay live to the end (loaded, so they cannot be rematerialised)
    unsigned h = tid * 747796405u + 1u, acc = 0;
    const unsigned myIters = iters / 2 + (tid * 2246822519u >> 22) % (iters / 2);   // divergent trip counts inside a warp
    for (unsigned it = 0; it < myIters; ++it)
    {
        h = h * 1664525u + 1013904223u;
        unsigned idx = (h >> 8) % n;
        // a chain of dependent, divergent loads with enough temporaries to keep the register allocator busy
        const unsigned a = table[idx], b = table[(idx * 31u + a) % n], c = table[(idx * 17u + b + 3u) % n], d = table[(idx + c * 5u) % n];
        acc += (a ^ b) + (c ^ d) * 3u + ((a + c) ^ (b + d));
#ifdef PROBE_IEEE_DIV
        // with -prec-div=true -prec-sqrt=true these compile to fast paths + CALL.REL.NOINC slow-path subroutines, as in the renderer
        const float q = (float)(a & 1023u) / (float)((b & 1023u) | 1u), r = sqrtf((float)(c & 4095u));
        const double qd = sqrt((double)(d & 65535u) + 0.5);
        acc += (unsigned)(q * 7.0f) + (unsigned)r + (unsigned)qd;
#endif
        if ((h & 15u) == 0u) { acc ^= table[(acc + it) % n]; }               // divergent side path
    }
    unsigned bad = 0;
#pragma unroll
    for (int k = 0; k < NV; ++k) if (v[k] != __ldcg(init + ((size_t)tid * NV + k) % n)) bad++;
    if (bad) atomicAdd(errors, (unsigned long long)bad);
    if (acc == 0xFFFFFFFFu) sink[0] = acc;
    // keep the values observable so the compiler cannot fold the check away
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < NV; ++k) s += v[k] * (k + 1);
    if (s == 0x12345678u) sink[1] = s;
}

__global__ void __launch_bounds__(128, 6) kernelB(const unsigned* __restrict__ table, unsigned n, unsigned iters, unsigned* out)
{
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned h = tid * 40503u + 977u, acc = 0;
    for (unsigned it = 0; it < iters; ++it) { h = h * 22695477u + 1u; acc += table[(h >> 9) % n]; }
    out[tid % n] = acc;
}

__global__ void kernelC(unsigned* out, unsigned n) { const unsigned i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) out[i] = i; }

int main()
{
    const unsigned n = 1u << 22;
    unsigned *table, *out, *sink, *init; unsigned long long* errors;
    cudaMalloc(&init, n * 4);
    {
        unsigned* h = new unsigned[n];
        for (unsigned i = 0; i < n; ++i) h[i] = i * 2654435761u + 12345u;
        cudaMemcpy(init, h, n * 4, cudaMemcpyHostToDevice);
        delete[] h;
    }
    cudaMalloc(&table, n * 4); cudaMalloc(&out, n * 4); cudaMalloc(&sink, 8); cudaMalloc(&errors, 8);
    cudaMemset(table, 1, n * 4);
    cudaStream_t s[4];
    for (auto& x : s) cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking);
    for (int mode = 0; mode < 3; ++mode)
    {
        cudaMemset(errors, 0, 8);
        cudaDeviceSynchronize();
        for (int rep = 0; rep < 40; ++rep)
        {
            kernelA<<<148 * 32, 128, 0, s[0]>>>(table, init, n, 400, errors, sink);
            if (mode >= 1)
            {
                kernelA<<<148 * 32, 128, 0, s[1]>>>(table, init, n, 300, errors, sink);
                kernelB<<<148 * 24, 128, 0, s[2]>>>(table, n, 600, out);
            }
            if (mode == 2) for (int k = 0; k < 100; ++k) kernelC<<<148 * 4, 256, 0, s[3]>>>(out + n / 2, n / 2);
        }
        cudaDeviceSynchronize();
        unsigned long long e = 0;
        cudaMemcpy(&e, errors, 8, cudaMemcpyDeviceToHost);
        printf("mode %d (%s): %llu spilled values came back wrong, last error: %s\n", mode,
               mode == 0 ? "one stream" : mode == 1 ? "overlapping streams" : "overlapping streams + 100 short launches per repetition", e, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
