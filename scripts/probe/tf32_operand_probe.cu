// Probe (groundwork for DESIGN.md section 10 item 1): two facts about tcgen05.mma.kind::tf32 that decide how the f32
// operands of the 3xTF32 GEMM can be fed.
//   1. What does the tensor core do with the low 13 mantissa bits of a 32-bit operand word?  D = A * e0 with A holding
//      full-precision f32 values returns f(A) exactly: compared with truncation and with round-to-nearest-away.
//   2. MN-major operands: B stored [k][n] (n contiguous), 128-byte swizzled, as TMA would write a box
//      {32 n, 32 k, n-groups} -- canonical layout ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)) with T = 4 for 32-bit elements,
//      LBO = 4096 B (next group of 32 n), SBO = 1024 B (next group of 8 k); instruction descriptor bit 16 (b_major).
//      A is K-major as in csrc/gemm_tf32x3.cu.  Small integers: the product is exact, compared with the host.
// One CTA, cta_group::1, M = 128, N = 64, K = 32 (four MMAs of K = 8).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I easy-ml_b200/csrc -I include \
//        scripts/probe/tf32_operand_probe.cu -o scripts/probe/tf32_operand_probe
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "sm100.cuh"

using namespace eml::sm100;

namespace {
constexpr int M = 128, N = 64, K = 32;

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_type << 61;  // 2: SWIZZLE_128B
    return d;
}

// mode 0: B K-major ([n][32 k], 128-byte rows, swizzled);  mode 1: B MN-major ([n-group][32 k][32 n], swizzled)
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                    int mode, uint32_t layout_type = 2, uint32_t lbo = 4096, uint32_t sbo = 1024) {
    __shared__ __align__(1024) uint8_t a_tile[M * 128];
    __shared__ __align__(1024) uint8_t b_tile[N * 128];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    // A(r, k): byte r * 128 + (((k / 4) ^ (r & 7)) << 4) + (k % 4) * 4
    for (int e = t; e < M * K; e += 128) {
        const int r = e / K, k = e % K;
        *reinterpret_cast<float*>(a_tile + r * 128 + ((((k >> 2) ^ (r & 7)) << 4) | ((k & 3) << 2))) = A[r * K + k];
    }
    if (mode == 2) {  // the tile holds its own word indexes: D tells which word the tensor core reads for (k, n)
        for (int w = t; w < N * 32; w += 128) reinterpret_cast<float*>(b_tile)[w] = (float)w;
    }
    for (int e = t; e < N * K && mode != 2; e += 128) {
        const int k = e / N, n = e % N;  // B given as [k][n]
        const float v = B[k * N + n];
        if (mode == 0) {
            *reinterpret_cast<float*>(b_tile + n * 128 + ((((k >> 2) ^ (n & 7)) << 4) | ((k & 3) << 2))) = v;
        } else {
            const int g = n >> 5, nn = n & 31;
            *reinterpret_cast<float*>(b_tile + g * 4096 + k * 128 + ((((nn >> 2) ^ (k & 7)) << 4) | ((nn & 3) << 2))) = v;
        }
    }
    if (t == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) tmem_alloc(&tmem_slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (t == 0) {
        uint32_t idesc = umma_idesc_tf32(M, N);
        if (mode >= 1) idesc |= 1u << 16;  // b_major = MN
        const uint32_t a0 = smem_u32(a_tile), b0 = smem_u32(b_tile);
        for (int kg = 0; kg < (mode == 2 ? 1 : K / 8); kg++) {
            const uint64_t da = umma_desc_k_sw128(a0) + 2 * kg;  // +32 bytes per group of 8 k
            const uint64_t db = mode == 0 ? umma_desc_k_sw128(b0) + 2 * kg : desc_mn_sw128(b0 + kg * 1024, lbo, sbo, layout_type);
            if (mode == 2 && kg == 0) printf("desc_b = 0x%016llx idesc = 0x%08x\n", (unsigned long long)db, idesc);
            umma_tf32(tmem, da, db, idesc, kg > 0);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    // warp w reads TMEM lanes 32 w .. 32 w + 31 (rows), 64 columns
    for (int c = 0; c < N; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; j++) D[(warp * 32 + lane) * N + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

float trunc13(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xffffe000u;
    memcpy(&x, &u, 4);
    return x;
}
float rna13(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xffffe000u;
    memcpy(&x, &u, 4);
    return x;
}
}  // namespace

int main() {
    std::vector<float> A(M * K), B(K * N), D(M * N);
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4);
    cudaMalloc(&dB, B.size() * 4);
    cudaMalloc(&dD, D.size() * 4);
    // ---- 1. operand rounding: A(r, 0) = full-precision values, B = e0 (B(0, 0) = 1) ----
    srand(1);
    std::fill(A.begin(), A.end(), 0.f);
    std::fill(B.begin(), B.end(), 0.f);
    for (int r = 0; r < M; r++) {
        uint32_t u = 0x3f800000u | ((uint32_t)rand() & 0x7fffffu);  // [1, 2) with random mantissa
        if (r & 1) u |= 0x80000000u;
        if (r == 2) u = 0x3f800000u | 0x1000u;   // exactly half an ulp of tf32 above 1
        if (r == 3) u = 0x3f800000u | 0x0fffu;   // just below half
        if (r == 4) u = 0x3f800000u | 0x1fffu;   // just below one ulp
        memcpy(&A[r * K], &u, 4);
    }
    B[0] = 1.0f;
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    probe_kernel<<<1, 128>>>(dA, dB, dD, 0);
    if (cudaDeviceSynchronize() != cudaSuccess) {
        printf("probe 1 failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int is_trunc = 0, is_rna = 0, is_exact = 0;
    for (int r = 0; r < M; r++) {
        const float x = A[r * K], got = D[r * N];
        is_trunc += got == trunc13(x);
        is_rna += got == rna13(x);
        is_exact += got == x;
    }
    printf("operand low bits over %d values: truncated %d, rounded-to-nearest-away %d, kept %d\n", M, is_trunc, is_rna, is_exact);
    for (int r = 2; r <= 4; r++) {
        uint32_t ux, ug;
        memcpy(&ux, &A[r * K], 4);
        memcpy(&ug, &D[r * N], 4);
        printf("  x = 0x%08x -> 0x%08x\n", ux, ug);
    }
    // ---- 2. MN-major B ----
    for (auto& v : A) v = (float)(rand() % 15 - 7);
    for (auto& v : B) v = (float)(rand() % 15 - 7);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 2; mode++) {
        probe_kernel<<<1, 128>>>(dA, dB, dD, mode);
        if (cudaDeviceSynchronize() != cudaSuccess) {
            printf("probe 2 mode %d failed: %s\n", mode, cudaGetErrorString(cudaGetLastError()));
            return 1;
        }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r = 0; r < M; r++)
            for (int n = 0; n < N; n++) {
                float ref = 0;
                for (int k = 0; k < K; k++) ref += A[r * K + k] * B[k * N + n];
                bad += ref != D[r * N + n];
            }
        printf("B %s: %d of %d elements differ from the host product\n", mode ? "MN-major (n contiguous)" : "K-major", bad, M * N);
    }
    // ---- 3. which shared-memory word does the tensor core read for B(k, n) with the MN-major descriptor? ----
    std::fill(A.begin(), A.end(), 0.f);
    for (int r = 0; r < 8; r++) A[r * K + r] = 1.0f;  // row r selects k = r
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    const uint32_t types[5] = {0, 1, 2, 4, 6};
    const char* names[5] = {"NONE / interleave", "128B_BASE32B", "128B", "64B", "32B"};
    for (int ti = 0; ti < 5; ti++) {
        cudaMemset(dD, 0xff, D.size() * 4);
        probe_kernel<<<1, 128>>>(dA, dB, dD, 2, types[ti], 4096, 1024);
        if (cudaDeviceSynchronize() != cudaSuccess) {
            printf("probe 3 (%s) failed: %s\n", names[ti], cudaGetErrorString(cudaGetLastError()));
            return 1;
        }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        printf("layout type %u (%s), LBO 4096, SBO 1024: word read for B(k, n)\n", types[ti], names[ti]);
        for (int k = 0; k < 8; k++) {
            printf("  k=%d:", k);
            for (int n = 0; n < N; n++) printf(" %d", (int)D[k * N + n]);
            printf("\n");
        }
    }
    return 0;
}
