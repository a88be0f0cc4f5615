// Stand-alone probe (development tool): tcgen05.mma with the A operand in TENSOR MEMORY (written by tcgen05.st, thread = row),
// B in shared memory (K-major, no swizzle).  Validates the A-in-TMEM layout against a CPU GEMM.
// usage: umma_probe_ts N K mode     mode 0: two fp16 per 32-bit column (k, k+1), 8 columns per K=16 MMA; mode 1: one per column
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../instance_nerf_b200/csrc/umma.cuh"

__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr uint32_t A_COL = 256;   // A operand columns start here; D at column 0

__global__ void __launch_bounds__(128) probe(const __half* A, const __half* B, float* D, int N, int K, int mode, uint32_t idesc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t lbo = 128, sboB = (K / 8) * 128;
    uint8_t* sB = smem;
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
        int r = i / K, k = i % K;
        *(__half*)(sB + umma::tile_off(r, k, lbo, sboB)) = B[r * K + k];
    }
    if (threadIdx.x == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
    if (threadIdx.x < 32) umma::tmem_alloc<512>(&tmem_slot);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t warp = threadIdx.x >> 5, row = threadIdx.x;
    const uint32_t lane_base = (warp * 32u) << 16;
    // A row `row` -> TMEM lane `row`
    const int per_col = mode == 0 ? 2 : 1;
    for (int c0 = 0; c0 < K / per_col; c0 += 8) {
        uint32_t v[8];
        for (int j = 0; j < 8; j++) {
            const int c = c0 + j;
            if (mode == 0) {
                const uint32_t lo = __half_as_ushort(A[row * K + 2 * c]), hi = __half_as_ushort(A[row * K + 2 * c + 1]);
                v[j] = lo | (hi << 16);
            } else {
                v[j] = __half_as_ushort(A[row * K + c]);
            }
        }
        tmem_st8(tmem + lane_base + A_COL + c0, v);
    }
    tmem_st_wait();
    umma::fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
        umma::fence_after_sync();
        for (int k = 0; k < K / 16; k++) {
            uint64_t db = umma::make_desc(umma::smem_u32(sB) + k * 2 * lbo, lbo, sboB);
            mma_f16_ts(tmem, tmem + A_COL + k * (16 / per_col), db, idesc, k > 0);
        }
        umma::commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    for (int c = 0; c < N; c += 16) {
        uint32_t v[16];
        umma::tmem_ld16(tmem + lane_base + c, v);
        umma::tmem_ld_wait();
        for (int j = 0; j < 16; j++) D[threadIdx.x * N + c + j] = __uint_as_float(v[j]);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) umma::tmem_dealloc<512>(tmem);
}

int main(int argc, char** argv) {
    const int M = 128;
    int N = argc > 1 ? atoi(argv[1]) : 64, K = argc > 2 ? atoi(argv[2]) : 64, mode = argc > 3 ? atoi(argv[3]) : 0;
    std::vector<__half> hA(M * K), hB(N * K);
    std::vector<float> fA(M * K), fB(N * K);
    srand(1);
    for (int i = 0; i < M * K; i++) { fA[i] = (rand() % 7 - 3) * 0.25f; hA[i] = __float2half(fA[i]); }
    for (int i = 0; i < N * K; i++) { fB[i] = (rand() % 5 - 2) * 0.5f; hB[i] = __float2half(fB[i]); }
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, 128 * N * 4);
    size_t smem = (size_t)N * K * 2 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 128, smem>>>(dA, dB, dD, N, K, mode, umma::make_idesc_f16(M, N));
    cudaError_t e = cudaDeviceSynchronize();
    printf("TS probe N=%d K=%d mode=%d launch: %s\n", N, K, mode, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> hD(128 * N);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) {
            float s = 0;
            for (int k = 0; k < K; k++) s += fA[m * K + k] * fB[n * K + k];
            if (hD[m * N + n] != s) bad++;
        }
    printf("mismatching elements: %d of %d  (D[0][0..3] = %g %g %g %g)\n", bad, M * N, hD[0], hD[1], hD[2], hD[3]);
    return 0;
}
