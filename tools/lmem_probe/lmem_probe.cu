// Stand-alone probe: is thread-private LOCAL memory preserved while kernels of different streams overlap on the same SMs?
// Kernel A keeps a dynamically indexed local array alive across a long loop of global loads and checks it at the end;
// kernel B (other streams, different local-memory footprint, short, launched back to back) only adds overlap.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lmem_probe lmem_probe.cu && ./lmem_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128, 8) kernelA(const unsigned* __restrict__ table, unsigned n, unsigned iters, unsigned long long* errors, unsigned* sink)
{
    unsigned stack[64];
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned h = tid * 2654435761u + 12345u;
    unsigned sp = 0, bad = 0, acc = 0;
    for (unsigned it = 0; it < iters; ++it)
    {
        h = h * 1664525u + 1013904223u;
        const unsigned v = table[(h >> 8) % n];          // divergent global load (latency)
        acc += v;
        if ((h & 3u) != 0u && sp < 64) { stack[sp] = (tid ^ (sp * 0x9E3779B9u)) + 7u * sp; sp++; }     // push
        else if (sp > 0) { sp--; if (stack[sp] != (tid ^ (sp * 0x9E3779B9u)) + 7u * sp) bad++; }       // pop + verify
    }
    while (sp > 0) { sp--; if (stack[sp] != (tid ^ (sp * 0x9E3779B9u)) + 7u * sp) bad++; }
    if (bad) atomicAdd(errors, (unsigned long long)bad);
    if (acc == 0xFFFFFFFFu) sink[0] = acc;
}

// same idea with a DIFFERENT local-memory footprint per thread (48 words) and a different register budget
__global__ void __launch_bounds__(128, 6) kernelB(const unsigned* __restrict__ table, unsigned n, unsigned iters, unsigned long long* errors, unsigned* out)
{
    unsigned stack[48];
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned h = tid * 40503u + 977u;
    unsigned sp = 0, bad = 0, acc = 0;
    for (unsigned it = 0; it < iters; ++it)
    {
        h = h * 22695477u + 1u;
        acc += table[(h >> 9) % n];
        if ((h & 7u) > 2u && sp < 48) { stack[sp] = (tid * 3u) ^ (sp * 0x85EBCA6Bu); sp++; }
        else if (sp > 0) { sp--; if (stack[sp] != ((tid * 3u) ^ (sp * 0x85EBCA6Bu))) bad++; }
    }
    while (sp > 0) { sp--; if (stack[sp] != ((tid * 3u) ^ (sp * 0x85EBCA6Bu))) bad++; }
    if (bad) atomicAdd(errors + 1, (unsigned long long)bad);
    out[tid % n] = acc;
}

__global__ void kernelC(unsigned* out, unsigned n) { const unsigned i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) out[i] = i; }

// short kernel that needs dynamic shared memory (forces a different L1 / shared carve-out than kernels A and B)
__global__ void __launch_bounds__(128, 8) kernelD(unsigned* out, unsigned n)
{
    extern __shared__ unsigned sm[];
    sm[threadIdx.x] = threadIdx.x * 3u;
    __syncthreads();
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = sm[(threadIdx.x * 7u) & 127u];
}

int main(int argc, char** argv)
{
    const unsigned n = 1u << 22;
    unsigned *table, *out, *sink; unsigned long long* errors;
    cudaMalloc(&table, n * 4); cudaMalloc(&out, n * 4); cudaMalloc(&sink, 4); cudaMalloc(&errors, 16);
    cudaMemset(table, 1, n * 4); cudaMemset(errors, 0, 16);
    cudaStream_t s[4];
    for (auto& x : s) cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking);
    cudaFuncSetAttribute(kernelD, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int mode = 0; mode < 4; ++mode)       // 0: kernel A alone (serial), 1: A on two streams + B and C on two more, 2: + hundreds of short launches, 3: + short launches with 16 KB of dynamic shared memory
    {
        cudaMemset(errors, 0, 16);
        cudaDeviceSynchronize();
        for (int rep = 0; rep < 40; ++rep)
        {
            kernelA<<<148 * 32, 128, 0, s[0]>>>(table, n, 2000, errors, sink);
            if (mode == 1)
            {
                kernelA<<<148 * 32, 128, 0, s[1]>>>(table, n, 1500, errors, sink);
                kernelB<<<148 * 24, 128, 0, s[2]>>>(table, n, 1200, errors, out);
                for (int k = 0; k < 8; ++k) kernelC<<<n / 256, 256, 0, s[3]>>>(out + n / 2, n / 2);
                if (mode >= 2) for (int k = 0; k < 200; ++k)
                {
                    if (mode == 3) kernelD<<<148 * 4, 128, 16 * 1024, s[3]>>>(out + n / 2, n / 2);
                    else kernelC<<<148 * 4, 256, 0, s[3]>>>(out + n / 2, n / 2);
                }
            }
        }
        cudaDeviceSynchronize();
        unsigned long long e[2] = {0, 0};
        cudaMemcpy(e, errors, 16, cudaMemcpyDeviceToHost);
        printf("mode %d (%s): %llu (kernel A, 64-word stack) + %llu (kernel B, 48-word stack) corrupted local-memory words, last error: %s\n", mode, mode == 0 ? "one stream" : mode == 1 ? "overlapping streams" : mode == 2 ? "overlap + 200 short launches per repetition" : "overlap + 200 short launches with 16 KB dynamic shared memory", e[0], e[1], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
