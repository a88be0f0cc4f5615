// Stand-alone probe (development tool, not part of the library): validates the UMMA descriptors / TMEM
// layout assumptions of umma.cuh against a CPU GEMM.  Usage: umma_probe M N K  (M in {64,128}).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../instance_nerf_b200/csrc/umma.cuh"

__global__ void __launch_bounds__(128) probe(const __half* A, const __half* B, float* D, int M, int N, int K, uint32_t idesc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t lbo = 128, sboA = (K / 8) * 128, sboB = (K / 8) * 128;
    uint8_t* sA = smem;
    uint8_t* sB = smem + 128 * K * 2;
    for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) {
        int r = i / K, k = i % K;
        *(__half*)(sA + umma::tile_off(r, k, lbo, sboA)) = r < M ? A[r * K + k] : __float2half(0.f);
    }
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
        int r = i / K, k = i % K;
        *(__half*)(sB + umma::tile_off(r, k, lbo, sboB)) = B[r * K + k];
    }
    if (threadIdx.x == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
    if (threadIdx.x < 32) umma::tmem_alloc<256>(&tmem_slot);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        for (int k = 0; k < K / 16; k++) {
            uint64_t da = umma::make_desc(umma::smem_u32(sA) + k * 2 * lbo, lbo, sboA);
            uint64_t db = umma::make_desc(umma::smem_u32(sB) + k * 2 * lbo, lbo, sboB);
            umma::mma_f16(tmem, da, db, idesc, k > 0);
        }
        umma::commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    const uint32_t warp = threadIdx.x >> 5;
    for (int c = 0; c < N; c += 16) {
        uint32_t v[16];
        umma::tmem_ld16(tmem + ((warp * 32u) << 16) + c, v);
        umma::tmem_ld_wait();
        for (int j = 0; j < 16; j++) D[threadIdx.x * N + c + j] = __uint_as_float(v[j]);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) umma::tmem_dealloc<256>(tmem);
}

int main(int argc, char** argv) {
    int M = argc > 1 ? atoi(argv[1]) : 128, N = argc > 2 ? atoi(argv[2]) : 64, K = argc > 3 ? atoi(argv[3]) : 32;
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
    size_t smem = (128 + N) * K * 2;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 128, smem>>>(dA, dB, dD, M, N, K, umma::make_idesc_f16(M, N));
    cudaError_t e = cudaDeviceSynchronize();
    printf("M=%d N=%d K=%d launch: %s\n", M, N, K, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> hD(128 * N);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    // expected
    int bad = 0;
    std::vector<int> lane_of_row(M, -1);
    for (int m = 0; m < M; m++) {
        std::vector<float> ref(N);
        for (int n = 0; n < N; n++) { float s = 0; for (int k = 0; k < K; k++) s += fA[m * K + k] * fB[n * K + k]; ref[n] = s; }
        for (int lane = 0; lane < 128; lane++) {
            bool ok = true;
            for (int n = 0; n < N; n++) if (hD[lane * N + n] != ref[n]) { ok = false; break; }
            if (ok) { lane_of_row[m] = lane; break; }
        }
        if (lane_of_row[m] != m) bad++;
    }
    printf("rows whose TMEM lane != row: %d\n", bad);
    printf("row->lane:");
    for (int m = 0; m < M; m += (M >= 64 ? 8 : 1)) printf(" %d:%d", m, lane_of_row[m]);
    printf("\n");
    return 0;
}
