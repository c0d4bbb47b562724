// tex_bench.cu -- micro-benchmark behind one design question of the cap path (analysis aid, not part of the product):
// the cap table is read as one random 32-byte bin per lane (LDG.E.256), which costs the LSU data pipe of L1TEX one
// wavefront per lane; does the texture path (tex1Dfetch<uint4>, two fetches per bin) move that traffic to a pipe of its
// own at a useful rate?  Every warp of a persistent grid (148 CTAs x 1024 threads, like the fused kernel) fetches
// `rounds` random bins per lane from a table of `mb` MB and ORs them together; shared-memory traffic of the fused kernel's
// magnitude can be added (smem_lds scattered LDS.128 per fetch) to see the two compete.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tex_bench tools/tex_bench.cu && ./tex_bench [mb] [rounds] [smem_lds]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t rng(uint32_t &s) { s = s * 1664525u + 1013904223u; return s ^ (s >> 15); }

template <int MODE>   // 0: LDG.256, 1: two tex1Dfetch<uint4>, 2: one tex1Dfetch<uint4> (inner mask only)
__global__ void __launch_bounds__(1024, 1) fetch_kernel(const uint4 *__restrict__ tab, cudaTextureObject_t tex, uint32_t nbins, int rounds,
                                                        int smem_lds, uint32_t *out) {
    extern __shared__ float4 s_atoms[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s_atoms[i] = make_float4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    uint32_t s = blockIdx.x * 1024u + threadIdx.x + 1u;
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    float acc = 0.f;
    for (int r = 0; r < rounds; ++r) {
        const uint32_t bin = rng(s) % nbins;
        if (MODE == 0) {
            uint32_t v0, v1, v2, v3, v4, v5, v6, v7;
            asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3), "=r"(v4), "=r"(v5), "=r"(v6), "=r"(v7) : "l"(tab + 2 * (size_t)bin));
            a0 |= v0 | v4; a1 |= v1 | v5; a2 |= v2 | v6; a3 |= v3 | v7;
        } else {
            const uint4 x = tex1Dfetch<uint4>(tex, 2 * (int)bin);
            a0 |= x.x; a1 |= x.y; a2 |= x.z; a3 |= x.w;
            if (MODE == 1) {
                const uint4 y = tex1Dfetch<uint4>(tex, 2 * (int)bin + 1);
                a0 |= y.x; a1 |= y.y; a2 |= y.z; a3 |= y.w;
            }
        }
        for (int k = 0; k < smem_lds; ++k) {
            const float4 b = s_atoms[(rng(s) >> 4) & 4095];
            acc += b.x + b.w;
        }
    }
    if ((a0 ^ a1 ^ a2 ^ a3) == 0x12345u || acc == 1.2345f) out[0] = a0;
}

int main(int argc, char **argv) {
    const size_t mb = argc > 1 ? atoi(argv[1]) : 68;
    const int rounds = argc > 2 ? atoi(argv[2]) : 2000;
    const int smem_lds = argc > 3 ? atoi(argv[3]) : 0;
    const uint32_t nbins = (uint32_t)(mb * 1024 * 1024 / 32);
    uint4 *tab;
    uint32_t *out;
    CK(cudaMalloc(&tab, (size_t)nbins * 32));
    CK(cudaMemset(tab, 1, (size_t)nbins * 32));
    CK(cudaMalloc(&out, 4));
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = tab;
    rd.res.linear.desc = cudaCreateChannelDesc<uint4>();
    rd.res.linear.sizeInBytes = (size_t)nbins * 32;
    cudaTextureDesc td = {};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const size_t smem = 4096 * 16;
    CK(cudaFuncSetAttribute(fetch_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(fetch_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(fetch_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    auto run = [&](int mode, const char *name) {
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            if (mode == 0) fetch_kernel<0><<<148, 1024, smem>>>(tab, tex, nbins, rounds, smem_lds, out);
            else if (mode == 1) fetch_kernel<1><<<148, 1024, smem>>>(tab, tex, nbins, rounds, smem_lds, out);
            else fetch_kernel<2><<<148, 1024, smem>>>(tab, tex, nbins, rounds, smem_lds, out);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
        }
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double lane_fetches = 148.0 * 1024 * rounds;
        printf("%-28s table %zu MB, %d scattered LDS.128 per fetch: %8.3f ms  %.2f lane-fetches/cycle/SM (1.965 GHz)  %.1f G bins/s\n", name, mb, smem_lds,
               ms, lane_fetches / 148.0 / (ms * 1e-3 * 1.965e9), lane_fetches / (ms * 1e-3) / 1e9);
    };
    run(0, "LDG.E.256 per bin");
    run(1, "2 x tex1Dfetch<uint4>");
    run(2, "1 x tex1Dfetch<uint4> (16 B)");
    return 0;
}
