// TMEM read throughput probe (tcgen05.ld 32x32b.x16 / .x64): cycles per load with 1, 4, 8 warps of one CTA per SM.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I instance_nerf_b200/csrc -o scripts/probe/ldtm_probe scripts/probe/ldtm_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
          "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
          "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr) : "memory");
}

// mode 0: x16 loads, wait after every `group` loads; mode 1: x64 loads
template <int MODE, int GROUP>
__global__ void __launch_bounds__(256) k_probe(int iters, int warps, unsigned long long* out, uint32_t* sink) {
    __shared__ uint32_t slot;
    const uint32_t warp = threadIdx.x >> 5;
    if (warp == 0) umma::tmem_alloc<512>(&slot);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t base = slot + (((warp & 3u) * 32u) << 16);
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    if ((int)warp < warps) {
        t0 = clock64();
        for (int i = 0; i < iters; i++) {
            if (MODE == 0) {
                uint32_t v[GROUP][16];
#pragma unroll
                for (int g = 0; g < GROUP; g++) umma::tmem_ld16(base + ((i * GROUP + g) * 16) % 512, v[g]);
                umma::tmem_ld_wait();
#pragma unroll
                for (int g = 0; g < GROUP; g++)
#pragma unroll
                    for (int j = 0; j < 16; j++) acc ^= v[g][j];
            } else {
                uint32_t v[64];
                tmem_ld64(base + (i * 64) % 512, v);
                umma::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 64; j++) acc ^= v[j];
            }
        }
        t1 = clock64();
    }
    if (acc == 0x12345678u) sink[0] = acc;
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (int)warp < warps) out[warp] = (unsigned long long)(t1 - t0);
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc<512>(slot);
}

template <int MODE, int GROUP>
void run(const char* name, int bytes_per_iter_per_warp) {
    unsigned long long* out; uint32_t* sink;
    cudaMalloc(&out, 64); cudaMalloc(&sink, 4);
    const int iters = 4096;
    for (int warps : {1, 4, 8}) {
        cudaMemset(out, 0, 64);
        k_probe<MODE, GROUP><<<148, 256>>>(iters, warps, out, sink);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long h[8];
        cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost);
        unsigned long long mx = 0;
        for (int w = 0; w < warps; w++) mx = h[w] > mx ? h[w] : mx;
        const double cyc = (double)mx / iters;
        printf("%-28s warps %d: %.1f cycles / iteration / warp, %.1f B/clk/SM  (%s)\n", name, warps, cyc,
               (double)bytes_per_iter_per_warp * warps / cyc, cudaGetErrorString(e));
    }
    cudaFree(out); cudaFree(sink);
}

int main() {
    run<0, 1>("x16, wait each", 2048);
    run<0, 2>("x16 x2, one wait", 4096);
    run<0, 4>("x16 x4, one wait", 8192);
    run<1, 1>("x64, wait each", 8192);
    return 0;
}
