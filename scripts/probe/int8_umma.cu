// Probe (groundwork for DESIGN.md section 10 item 6): an int8 x int8 -> int32 GEMM on tcgen05.mma.kind::i8 with the
// CTA-pair pipeline of csrc/gemm_tf32x3.cu (TMA SW128 K-major operands, 256 x 256 tile per SM pair, two int32
// accumulator stages in TMEM).  Checks C = A * B^T against the host on a small case, then times 8192^3.
//   C[m][n] = sum_k A[m][k] * B[n][k]      A: M x K int8 (k contiguous), B: N x K int8 (k contiguous), C: int32
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I easy-ml_b200/csrc -I include \
//        scripts/probe/int8_umma.cu -o scripts/probe/int8_umma -L easy-ml_b200/lib -leasyml_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../easy-ml_b200/lib'
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.h"
#include "sm100.cuh"

using namespace eml;
using namespace eml::sm100;

namespace {
constexpr int PM = 256, PN = 256, HALF = 128, BKB = 128;  // BKB: bytes (= int8 elements) of k per stage row
constexpr int TILE_BYTES = HALF * BKB;                    // 16 KB
constexpr int STAGE_BYTES = 2 * TILE_BYTES;               // A rows of this CTA, B rows of this CTA
constexpr int STAGES = 6, ACC_STAGES = 2;
constexpr uint32_t TMEM_COLS = 512;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 + 256;
constexpr int THREADS = 12 * 32;

__host__ __device__ constexpr uint32_t umma_idesc_i8(int M, int N) {
    return (2u << 4)        // c_format = S32
           | (1u << 7)      // a_format = signed 8 bit
           | (1u << 10)     // b_format = signed 8 bit
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

struct Maps {
    CUtensorMap a, b;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
int8_pair_kernel(const __grid_constant__ Maps maps, int32_t* __restrict__ C, int M, int N, int kblocks, int tiles_m,
                 int tiles_n) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* acc_full = empty + STAGES;
    uint64_t* acc_empty = acc_full + ACC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + ACC_STAGES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster = blockIdx.x >> 1, clusters = gridDim.x >> 1;
    const int tiles = tiles_m * tiles_n;
    if (threadIdx.x == 0) {
        prefetch_tensormap(&maps.a);
        prefetch_tensormap(&maps.b);
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < ACC_STAGES; a++) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 16);  // 8 epilogue warps of each CTA
        }
        fence_barrier_init();
    }
    __syncwarp();
    if (warp == 1) tmem_alloc_pair(tmem_slot, TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int n = 0;
            for (int t = cluster; t < tiles; t += clusters) {
                const int tm = t / tiles_n, tn = t % tiles_n;
                const int row_a = tm * PM + (int)rank * HALF, row_b = tn * PN + (int)rank * HALF;
                for (int kb = 0; kb < kblocks; kb++, n++) {
                    const int s = n % STAGES;
                    if (n >= STAGES) mbar_wait(&empty[s], ((n / STAGES) - 1) & 1);
                    uint8_t* st = base + (size_t)s * STAGE_BYTES;
                    if (rank == 0) mbar_expect_tx(&full[s], 2 * STAGE_BYTES);
                    tma_load_3d_pair(st, &maps.a, &full[s], kb * BKB, row_a, 0);
                    tma_load_3d_pair(st + TILE_BYTES, &maps.b, &full[s], kb * BKB, row_b, 0);
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_i8(PM, PN);
            int n = 0, it = 0;
            for (int t = cluster; t < tiles; t += clusters, it++) {
                const int as = it & 1;
                if (it >= ACC_STAGES) mbar_wait(&acc_empty[as], ((it >> 1) - 1) & 1);
                tc_fence_after();
                const uint32_t d = tmem_acc + (uint32_t)as * PN;
                for (int kb = 0; kb < kblocks; kb++, n++) {
                    const int s = n % STAGES;
                    mbar_wait(&full[s], (n / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t st = smem_u32(base + (size_t)s * STAGE_BYTES);
                    const uint64_t da = umma_desc_k_sw128(st), db = umma_desc_k_sw128(st + TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BKB / 32; k++)  // 32 int8 = 32 bytes of k per MMA: descriptor + 2 (16-byte units)
                        umma_i8_pair(d, da + 2 * k, db + 2 * k, idesc, (uint32_t)(kb | k));
                    umma_commit_pair(&empty[s], 0b11);
                }
                umma_commit_pair(&acc_full[as], 0b11);
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3, half = (warp - 4) >> 2;
        int it = 0;
        for (int t = cluster; t < tiles; t += clusters, it++) {
            const int tm = t / tiles_n, tn = t % tiles_n;
            const int as = it & 1;
            mbar_wait(&acc_full[as], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t t0 = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * PN + half * 128);
            const int64_t row = (int64_t)tm * PM + (int64_t)rank * HALF + q * 32 + lane;
            const int64_t n0 = (int64_t)tn * PN + half * 128;
#pragma unroll 1
            for (int c = 0; c < 4; c++) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(t0 + (uint32_t)(c * 32), v);
                tmem_ld_wait();
                if (row < M) {
                    int32_t* p = C + row * N + n0 + c * 32;
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if (n0 + c * 32 + 4 * j + 3 < N)
                            *reinterpret_cast<int4*>(p + 4 * j) = make_int4((int)v[4 * j], (int)v[4 * j + 1], (int)v[4 * j + 2], (int)v[4 * j + 3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&acc_empty[as], 0);
        }
    }
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair(tmem_acc, TMEM_COLS);
}

int plane_map(CUtensorMap* map, const int8_t* base, int64_t rows, int64_t K) {
    uint64_t dims[3] = {(uint64_t)K, (uint64_t)rows, 1};
    uint64_t strides[2] = {(uint64_t)K, (uint64_t)rows * (uint64_t)K};
    uint32_t box[3] = {(uint32_t)BKB, (uint32_t)HALF, 1};
    return make_tensor_map(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

float run(const int8_t* dA, const int8_t* dB, int32_t* dC, int M, int N, int K, int reps) {
    Maps maps;
    if (plane_map(&maps.a, dA, M, K) || plane_map(&maps.b, dB, N, K)) {
        printf("tensor map failed: %s\n", last_error());
        exit(1);
    }
    cudaFuncSetAttribute(int8_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    const int tiles_m = (M + PM - 1) / PM, tiles_n = (N + PN - 1) / PN;
    int clusters = 74;
    if (tiles_m * tiles_n < clusters) clusters = tiles_m * tiles_n;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int8_pair_kernel<<<2 * clusters, THREADS, SMEM_BYTES>>>(maps, dC, M, N, K / BKB, tiles_m, tiles_n);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < reps; r++)
        int8_pair_kernel<<<2 * clusters, THREADS, SMEM_BYTES>>>(maps, dC, M, N, K / BKB, tiles_m, tiles_n);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("launch status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return ms / reps;
}
}  // namespace

int main() {
    {  // correctness: 512 x 768 x 640, digits in [-64, 64]
        const int M = 512, N = 768, K = 640;
        std::vector<int8_t> a((size_t)M * K), b((size_t)N * K);
        uint32_t s = 12345;
        auto next = [&]() { s = s * 1664525u + 1013904223u; return (int8_t)((int)((s >> 16) % 129) - 64); };
        for (auto& x : a) x = next();
        for (auto& x : b) x = next();
        int8_t *dA, *dB;
        int32_t* dC;
        cudaMalloc(&dA, a.size());
        cudaMalloc(&dB, b.size());
        cudaMalloc(&dC, sizeof(int32_t) * M * N);
        cudaMemcpy(dA, a.data(), a.size(), cudaMemcpyHostToDevice);
        cudaMemcpy(dB, b.data(), b.size(), cudaMemcpyHostToDevice);
        run(dA, dB, dC, M, N, K, 1);
        std::vector<int32_t> c((size_t)M * N);
        cudaMemcpy(c.data(), dC, sizeof(int32_t) * M * N, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int m = 0; m < M; m += 7)
            for (int n = 0; n < N; n += 5) {
                int32_t ref = 0;
                for (int k = 0; k < K; k++) ref += (int)a[(size_t)m * K + k] * (int)b[(size_t)n * K + k];
                if (ref != c[(size_t)m * N + n]) bad++;
            }
        printf("correctness 512x768x640: %ld mismatches (sampled)\n", bad);
        cudaFree(dA); cudaFree(dB); cudaFree(dC);
    }
    {  // throughput: 8192^3
        const int n = 8192;
        int8_t *dA, *dB;
        int32_t* dC;
        cudaMalloc(&dA, (size_t)n * n);
        cudaMalloc(&dB, (size_t)n * n);
        cudaMalloc(&dC, sizeof(int32_t) * (size_t)n * n);
        cudaMemset(dA, 3, (size_t)n * n);
        cudaMemset(dB, 5, (size_t)n * n);
        float ms = run(dA, dB, dC, n, n, n, 10);
        printf("int8 8192^3: %.3f ms = %.1f TOP/s (dense int8 nominal 4500)\n", ms, 2.0 * n * n * (double)n / (ms * 1e-3) / 1e12);
    }
    return 0;
}
