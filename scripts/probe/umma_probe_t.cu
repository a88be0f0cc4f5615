// Stand-alone probe (development tool): can a tile written in the K-major no-swizzle layout of umma.cuh
// (row = sample s, 16-byte chunks of 8 columns) be consumed as an MN-major operand, i.e. D[m, n] = sum_s A[s, m] * B[s, n]
// (the dW = dY^T * H reduction over samples of the MLP backward) without writing a transposed copy?
// Tries the descriptor LBO/SBO assignments and prints which one reproduces the CPU result.  Usage: umma_probe_t MA NB
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../instance_nerf_b200/csrc/umma.cuh"

__host__ __device__ constexpr uint32_t idesc_t(uint32_t M, uint32_t N, uint32_t amaj, uint32_t bmaj) {
    return (1u << 4) | (amaj << 15) | (bmaj << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__global__ void __launch_bounds__(128) probe(const __half* A, const __half* B, float* D, int MA, int NB, int variant) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t sboA = (MA / 8) * 128, sboB = (NB / 8) * 128;   // K-major strides of the [128 x MA] / [128 x NB] tiles
    uint8_t* sA = smem;
    uint8_t* sB = smem + 128 * MA * 2;
    for (int i = threadIdx.x; i < 128 * MA; i += blockDim.x) { int s = i / MA, j = i % MA; *(__half*)(sA + umma::tile_off(s, j, 128, sboA)) = A[s * MA + j]; }
    for (int i = threadIdx.x; i < 128 * NB; i += blockDim.x) { int s = i / NB, j = i % NB; *(__half*)(sB + umma::tile_off(s, j, 128, sboB)) = B[s * NB + j]; }
    if (threadIdx.x == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
    if (threadIdx.x < 32) umma::tmem_alloc<256>(&tmem_slot);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = idesc_t(128, NB, 1, 1);
        for (int k = 0; k < 128 / 16; k++) {   // 16 samples per instruction = 2 groups of 8 rows
            // variant 0: LBO = stride between 8-sample groups (K direction), SBO = stride between 8-column chunks (MN direction)
            // variant 1: swapped
            const uint32_t lboA = variant == 0 ? sboA : 128, sbA = variant == 0 ? 128 : sboA;
            const uint32_t lboB = variant == 0 ? sboB : 128, sbB = variant == 0 ? 128 : sboB;
            uint64_t da = umma::make_desc(umma::smem_u32(sA) + k * 2 * sboA, lboA, sbA);
            uint64_t db = umma::make_desc(umma::smem_u32(sB) + k * 2 * sboB, lboB, sbB);
            umma::mma_f16(tmem, da, db, idesc, k > 0);
        }
        umma::commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    const uint32_t warp = threadIdx.x >> 5;
    for (int c = 0; c < NB; c += 16) {
        uint32_t v[16];
        umma::tmem_ld16(tmem + ((warp * 32u) << 16) + c, v);
        umma::tmem_ld_wait();
        for (int j = 0; j < 16; j++) D[threadIdx.x * NB + c + j] = __uint_as_float(v[j]);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) umma::tmem_dealloc<256>(tmem);
}

int main(int argc, char** argv) {
    int MA = argc > 1 ? atoi(argv[1]) : 128, NB = argc > 2 ? atoi(argv[2]) : 64;
    std::vector<__half> hA(128 * MA), hB(128 * NB);
    std::vector<float> fA(128 * MA), fB(128 * NB);
    srand(1);
    for (size_t i = 0; i < fA.size(); i++) { fA[i] = (rand() % 7 - 3) * 0.25f; hA[i] = __float2half(fA[i]); }
    for (size_t i = 0; i < fB.size(); i++) { fB[i] = (rand() % 5 - 2) * 0.5f; hB[i] = __float2half(fB[i]); }
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * NB * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    size_t smem = 128 * (MA + NB) * 2;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int variant = 0; variant < 2; variant++) {
        cudaMemset(dD, 0xff, 128 * NB * 4);
        probe<<<1, 128, smem>>>(dA, dB, dD, MA, NB, variant);
        cudaError_t e = cudaDeviceSynchronize();
        printf("MA=%d NB=%d variant=%d launch: %s\n", MA, NB, variant, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<float> hD(128 * NB);
        cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128 && m < MA; m++)
            for (int n = 0; n < NB; n++) {
                float s = 0;
                for (int k = 0; k < 128; k++) s += fA[k * MA + m] * fB[k * NB + n];
                if (hD[m * NB + n] != s) bad++;
            }
        printf("  mismatches: %d of %d\n", bad, (MA < 128 ? MA : 128) * NB);
    }
    return 0;
}
