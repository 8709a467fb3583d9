// Shared-memory atomic throughput by address pattern (sm_100a): cycles per warp-wide red.shared.add.u32 [ad], 1
// (ATOMS.POPC.INC) with every SM full of warps.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/atoms_probe tools/atoms_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int WORDS = 12288;   // 48 KB
constexpr int NADDR = 64;      // addresses per lane, cycled

__global__ void __launch_bounds__(256) probe(const unsigned *__restrict__ addr, int iters, unsigned long long *cycles, unsigned *sink)
{
    __shared__ unsigned sh[WORDS];
    for (int i = threadIdx.x; i < WORDS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    unsigned a[8];
    const unsigned base = (unsigned)__cvta_generic_to_shared(sh);
    const unsigned *mine = addr + (size_t)(threadIdx.x & 31) * NADDR + (threadIdx.x >> 5) * 8;   // warps start at different phases
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = base + 4u * mine[u];
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a[u]) : "memory");
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) atomicMax(cycles, (unsigned long long)(t1 - t0));
    unsigned s = 0;
    for (int i = threadIdx.x; i < WORDS; i += blockDim.x) s += sh[i];
    if (s == 12345u) sink[0] = s;
}

int main()
{
    const int iters = 4096, blocks_per_sm = 4;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned *d_addr, *d_sink;
    unsigned long long *d_cyc;
    cudaMalloc(&d_addr, 32 * NADDR * 4);
    cudaMalloc(&d_sink, 4);
    cudaMalloc(&d_cyc, 8);
    struct Pat { const char *name; int id; };
    const Pat pats[] = {{"lane-linear (word = lane)", 0}, {"all lanes one word", 1}, {"random over 12288 words", 2},
                        {"copy-major: (lane & 7) * 356 + k, k in a window of 25", 3}, {"bin-major: k * 8 + (lane & 7), window of 25", 4},
                        {"own bank: k * 32 + lane, window of 25", 5}, {"own bank, window of 200", 6},
                        {"copy-major, window of 200", 7}, {"16 copies: (lane & 15) * 356 + k, window 25", 8},
                        {"lane-linear + row offset per lane pair ((lane>>1)&3)*32", 9}, {"2 lanes per word (word = lane >> 1)", 10},
                        {"4 lanes per word (word = lane >> 2)", 11}, {"own bank, every lane another row: lane * 33", 12},
                        {"half the lanes one word, the rest copy-major window 25", 13}};
    for (const Pat &p : pats) {
        std::vector<unsigned> h(32 * NADDR);
        srand(7);
        for (int l = 0; l < 32; ++l)
            for (int i = 0; i < NADDR; ++i) {
                const int k25 = 100 + rand() % 25, k200 = rand() % 200;
                unsigned w = 0;
                switch (p.id) {
                case 0: w = l; break;
                case 1: w = 5; break;
                case 2: w = rand() % WORDS; break;
                case 3: w = (l & 7) * 356 + k25; break;
                case 4: w = k25 * 8 + (l & 7); break;
                case 5: w = k25 * 32 + l; break;
                case 6: w = k200 * 32 + l; break;
                case 7: w = (l & 7) * 356 + k200; break;
                case 8: w = (l & 15) * 356 + k25; break;
                case 9: w = l + ((l >> 1) & 3) * 32; break;
                case 10: w = l >> 1; break;
                case 11: w = l >> 2; break;
                case 12: w = l * 33; break;
                case 13: w = (l & 1) ? 7000 : (l & 7) * 356 + k25; break;
                }
                h[l * NADDR + i] = w % WORDS;
            }
        cudaMemcpy(d_addr, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(d_cyc, 0, 8);
            probe<<<sms * blocks_per_sm, 256>>>(d_addr, iters, d_cyc, d_sink);
            cudaDeviceSynchronize();
        }
        unsigned long long cyc = 0;
        cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
        const double instr_per_sm = (double)iters * 8 * 8 * blocks_per_sm;   // warp instructions per SM
        printf("%-62s %.2f cycles per warp instruction per SM  (%.1f lanes / clk / SM)\n", p.name, cyc / instr_per_sm, 32.0 * instr_per_sm / cyc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
