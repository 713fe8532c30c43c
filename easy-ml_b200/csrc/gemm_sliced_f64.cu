// gemm_sliced_f64.cu -- f64 GEMM beyond the FP64 pipe: error-free slicing onto the INT8 tensor cores.
//
// The DMMA kernel (gemm_dmma_f64.cu) runs at 97.7 % of a ceiling that no FP64 instruction mix can raise
// (scripts/probe/fp64_pipes.cu).  B200's INT8 tensor cores are ~100x faster (this file's kernel: 3.5 POP/s), and a
// product of two f64 matrices can be assembled from exact integer products (the Ozaki scheme):
//
//   row i of A is scaled by R_i = 2^(e_i + 1), e_i the largest exponent of the row, so x = a / R_i lies in (-1, 1);
//   x = sum_{p < s} dA_p 2^(-6 - 7p) + rho,  dA_p in [-64, 64] (int8),  |rho| <= 2^(-7s)   -- exact in f64 arithmetic;
//   the same for the columns of B (S_j, dB_q).  Then
//   a_ik b_kj = R_i S_j ( sum_{p + q < s} dA_p dB_q 2^(-12 - 7 (p + q))  +  err ),   |err| <= (s + 2) 2^(-7s).
//   The slice products  sum_k dA_p[i,k] dB_q[k,j]  are int8 GEMMs whose int32 sums are exact (|d d'| <= 2^12,
//   at most s K terms per accumulator: K <= 32768 for s = 8); the diagonals p + q = d are accumulated in ONE int32
//   TMEM accumulator each and added into C in f64, smallest weight first.
//
// Parity.  The contract is |C - ref| <= 1e-12 sum_k |a_ik b_kj| per element.  The truncation error above is relative to
// the row / column maxima, not to the elements, so it only meets the contract when sum_k |x_ik y_kj| is not tiny:
//   K (s + 2) 2^(-7s) <= 0.9e-12 sum_k |x_ik| |y_kj|.
// A GUARD pass proves that before anything is trusted: it multiplies the planes floor(64 |x|), floor(64 |y|) (lower
// bounds of the magnitudes) with the same int8 kernel and checks every element's sum against the threshold
// K (s + 2) 2^(12 - 7s) / 0.9e-12 (rows / columns that are entirely zero are exact and exempt).  If any element fails,
// or an operand holds Inf / NaN / extreme exponents, the call runs the DMMA kernel instead -- the contract never depends
// on the data.  Opt-in entry point (eml_gemm_f64_sliced_dev); the default f64 path stays on the DMMA kernel.
#include <climits>
#include <cmath>
#include <cstdlib>

#include "common.h"
#include "sm100.cuh"

namespace eml {

namespace {
using namespace sm100;

namespace sl {
constexpr int PM = 256, PN = 256, HALF = 128, BKB = 128;  // BKB: int8 elements (= bytes) of k per stage row
constexpr int TILE_BYTES = HALF * BKB;                    // 16 KB
constexpr int STAGE_BYTES = 2 * TILE_BYTES;
constexpr int STAGES = 6, ACC_STAGES = 2;
constexpr uint32_t TMEM_COLS = 512;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 + 256;
constexpr int THREADS = 12 * 32;
constexpr int MAX_SLICES = 10;
constexpr int ZERO_ROW = INT_MIN;  // exponent sentinel of an all-zero row / column

__host__ __device__ constexpr uint32_t idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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

// unbiased exponent from the bit pattern (the library ilogb costs ~20 instructions per element): zero -> ZERO_ROW,
// subnormals report -1023 (the scan flags |exponent| > 900 anyway), Inf / NaN report 1024
__device__ __forceinline__ int exponent_of(double v) {
    const int hi = __double2hiint(v) & 0x7fffffff;
    if ((hi | __double2loint(v)) == 0) return ZERO_ROW;
    return (hi >> 20) - 1023;
}

// ---------------------------------------------------------------- exponents of the rows of A / columns of B
// out[r] = largest ilogb over the row, ZERO_ROW for an all-zero row; *bad |= 1 for Inf / NaN / |exponent| > 900
__global__ void __launch_bounds__(256)
row_exponent_kernel(const double* __restrict__ X, int64_t rows, int64_t cols, int64_t ld, int* __restrict__ out,
                    int* __restrict__ bad) {
    pdl_prologue();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + warp;
    if (r >= rows) return;
    int e = ZERO_ROW, flag = 0;
    const double* row = X + r * ld;
    int64_t c = lane;
    for (; c + 96 < cols; c += 128) {  // four loads in flight per lane
        const double v0 = row[c], v1 = row[c + 32], v2 = row[c + 64], v3 = row[c + 96];
        e = max(max(e, exponent_of(v0)), max(max(exponent_of(v1), exponent_of(v2)), exponent_of(v3)));
    }
    for (; c < cols; c += 32) e = max(e, exponent_of(row[c]));
    if (e == 1024) flag = 1;  // Inf or NaN
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        e = max(e, __shfl_xor_sync(0xffffffffu, e, o));
        flag |= __shfl_xor_sync(0xffffffffu, flag, o);
    }
    if (lane == 0) {
        out[r] = e;
        if (flag || (e != ZERO_ROW && (e > 900 || e < -900))) atomicOr(bad, 1);
    }
}
// columns: out[c] = largest ilogb over the column.  One thread per column and row chunk, combined with atomicMax
// (out pre-filled with ZERO_ROW); reads are coalesced along the row.
__global__ void __launch_bounds__(256)
col_exponent_kernel(const double* __restrict__ X, int64_t rows, int64_t cols, int64_t ld, int64_t rows_per_block,
                    int* __restrict__ out, int* __restrict__ bad) {
    pdl_prologue();
    const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    if (c >= cols) return;
    int e = ZERO_ROW, flag = 0;
    int64_t r = r0;
    for (; r + 3 < r1; r += 4) {
        const double v0 = X[r * ld + c], v1 = X[(r + 1) * ld + c], v2 = X[(r + 2) * ld + c], v3 = X[(r + 3) * ld + c];
        e = max(max(e, exponent_of(v0)), max(max(exponent_of(v1), exponent_of(v2)), exponent_of(v3)));
    }
    for (; r < r1; r++) e = max(e, exponent_of(X[r * ld + c]));
    if (e == 1024) flag = 1;
    if (e != ZERO_ROW) atomicMax(out + c, e);
    if (flag || (e != ZERO_ROW && (e > 900 || e < -900))) atomicOr(bad, 1);
}
__global__ void fill_int_kernel(int* p, int v, int64_t n) {
    pdl_prologue();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// scale[r] = 2^(e_r + 1) (1 for a zero row), inv[r] = 2^-(e_r + 1): both exact powers of two
__global__ void scale_kernel(const int* __restrict__ e, double* __restrict__ scale, double* __restrict__ inv, int64_t n) {
    pdl_prologue();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int x = e[i] == ZERO_ROW || e[i] > 1000 || e[i] < -1000 ? -1 : e[i];
        scale[i] = ldexp(1.0, x + 1);
        inv[i] = ldexp(1.0, -(x + 1));
    }
}

// ---------------------------------------------------------------- slicing
// digits of x in (-1, 1): x = sum_p d[p] 2^(-6 - 7p) + rho; plane `slices` = floor(64 |x|) (the guard's lower bound)
__device__ __forceinline__ void digits(double x, int slices, int8_t (&d)[MAX_SLICES], int8_t& guard) {
    guard = (int8_t)__double2int_rz(fabs(x) * 64.0);  // floor: 0..63
    if (slices == 7 || slices == 8) {
        // Fixed point in two 32-bit halves, digits on the integer pipe (the FP64-arithmetic form below costs ~160
        // instructions per element and made this kernel issue-bound):
        //   x 2^27 = H + r,  H = round(x 2^27) (four digits, weights 2^-6 .. 2^-27),  |r| <= 1/2 exactly representable;
        //   L = round(r 2^(7 (s - 4)))  (the remaining s - 4 digits).
        // Balanced base-128 digits from the least significant end: d in [-64, 63], the leading one of each half in
        // [-64, 64].  Dropping trailing digits leaves |rest| <= 2^(-7 s') (1 + 2^-7) for the s' kept: within the
        // (s + 2) 2^(-7s) budget of the error analysis.
        const double xs = x * 134217728.0;  // 2^27
        int H = __double2int_rn(xs);
        int L = __double2int_rn((xs - (double)H) * (slices == 8 ? 268435456.0 : 2097152.0));  // 2^28 / 2^21
#pragma unroll
        for (int p = 7; p >= 5; p--) {
            d[p] = 0;
            if (p < slices) {
                const int low = ((L + 64) & 127) - 64;
                d[p] = (int8_t)low;
                L = (L - low) >> 7;
            }
        }
        d[4] = (int8_t)L;
#pragma unroll
        for (int p = 3; p >= 1; p--) {
            const int low = ((H + 64) & 127) - 64;
            d[p] = (int8_t)low;
            H = (H - low) >> 7;
        }
        d[0] = (int8_t)H;
#pragma unroll
        for (int p = 8; p < MAX_SLICES; p++) d[p] = 0;
        return;
    }
    double t = x * 64.0;
    constexpr double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: (t + MAGIC) - MAGIC rounds to the nearest integer
#pragma unroll
    for (int p = 0; p < MAX_SLICES; p++) {
        d[p] = 0;
        if (p < slices) {
            const double r = __dsub_rn(__dadd_rn(t, MAGIC), MAGIC);
            d[p] = (int8_t)__double2int_rz(r);
            t = (t - r) * 128.0;
        }
    }
}
// k-contiguous source (A, row-major M x K): planes[p][row][k], row stride Kp bytes.  One thread = 4 consecutive k.
__global__ void __launch_bounds__(256)
slice_rows_kernel(const double* __restrict__ X, int64_t rows, int64_t K, int64_t ld, const double* __restrict__ inv,
                  int8_t* __restrict__ planes, int64_t Kp, int slices) {
    pdl_prologue();
    const int64_t k4 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    const int64_t r = blockIdx.y;
    if (k4 >= Kp) return;
    const double is = inv[r];
    int8_t d[4][MAX_SLICES], g[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const double v = (k4 + j < K) ? X[r * ld + k4 + j] : 0.0;
        digits(v * is, slices, d[j], g[j]);
    }
    const int64_t plane = rows * Kp;
    int8_t* out = planes + r * Kp + k4;
#pragma unroll
    for (int p = 0; p < MAX_SLICES; p++)
        if (p < slices) *reinterpret_cast<char4*>(out + p * plane) = make_char4(d[0][p], d[1][p], d[2][p], d[3][p]);
    *reinterpret_cast<char4*>(out + slices * plane) = make_char4(g[0], g[1], g[2], g[3]);
}
// n-contiguous source (B, row-major K x N): planes[p][n][k] -- transposed through shared memory.
// Block tile: 128 k x 32 n.
__global__ void __launch_bounds__(256)
slice_cols_kernel(const double* __restrict__ X, int64_t K, int64_t N, int64_t ld, const double* __restrict__ inv,
                  int8_t* __restrict__ planes, int64_t Kp, int slices) {
    pdl_prologue();
    __shared__ int8_t tile[MAX_SLICES + 1][32][132];  // [plane][n][k], row padded to 132 bytes
    const int64_t k0 = (int64_t)blockIdx.x * 128, n0 = (int64_t)blockIdx.y * 32;
    const int tn = threadIdx.x & 31, tk = threadIdx.x >> 5;  // 32 n x 8 k per pass
    const int64_t n = n0 + tn;
    const double is = n < N ? inv[n] : 1.0;
    for (int kk = tk; kk < 128; kk += 8) {
        const int64_t k = k0 + kk;
        const double v = (n < N && k < K) ? X[k * ld + n] : 0.0;
        int8_t d[MAX_SLICES], g;
        digits(v * is, slices, d, g);
#pragma unroll
        for (int p = 0; p < MAX_SLICES; p++)
            if (p < slices) tile[p][tn][kk] = d[p];
        tile[slices][tn][kk] = g;
    }
    __syncthreads();
    // store: each thread writes 4 consecutive k of one n row of one plane
    const int64_t plane = N * Kp;
    for (int idx = threadIdx.x; idx < (slices + 1) * 32 * 32; idx += 256) {
        const int p = idx / (32 * 32), rem = idx % (32 * 32), nn = rem / 32, k4 = (rem % 32) * 4;
        if (n0 + nn < N && k0 + k4 < Kp)
            *reinterpret_cast<char4*>(planes + p * plane + (n0 + nn) * Kp + k0 + k4) =
                make_char4(tile[p][nn][k4], tile[p][nn][k4 + 1], tile[p][nn][k4 + 2], tile[p][nn][k4 + 3]);
    }
}

// ---------------------------------------------------------------- the int8 GEMM over slice planes
// tile raster: groups of 8 tile rows walked column by column, so that the 74 tiles in flight form a ~8 x 9 block and
// share 17 operand panels per slice pair through L2 (row-major order touched 35: twice the DRAM traffic, and this
// kernel is power-bound)
__device__ __forceinline__ void tile_of(int t, int tiles_m, int tiles_n, int& tm, int& tn) {
    constexpr int GROUP = 8;
    const int per_group = GROUP * tiles_n;
    const int first_m = (t / per_group) * GROUP;
    const int rows = tiles_m - first_m < GROUP ? tiles_m - first_m : GROUP;
    const int in_group = t % per_group;
    tm = first_m + in_group % rows;
    tn = in_group / rows;
}
struct Maps {
    CUtensorMap a, b;  // dims {k, rows, plane}
};

// GUARD = false: for every tile, the diagonals d = s-1 .. 0; per diagonal the pairs (p, d - p) accumulate in one int32
//                TMEM accumulator; the epilogue adds  acc * R_i S_j 2^(-12 - 7d)  into C (the first diagonal writes).
// GUARD = true : one product, planes (s, s) = floor(64 |x|), floor(64 |y|); the epilogue leaves the smallest sum in
//                flag[1] (zero rows / columns exempt): the host compares it with the threshold of each slice count.
template <bool GUARD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
sliced_kernel(const __grid_constant__ Maps maps, double* __restrict__ C, int64_t M, int64_t N, int64_t ldc, int kblocks,
              int tiles_m, int tiles_n, int slices, const double* __restrict__ row_scale,
              const double* __restrict__ col_scale, const int* __restrict__ row_exp, const int* __restrict__ col_exp,
              int threshold, int* __restrict__ flag, unsigned int* __restrict__ step_counters) {
    pdl_prologue();
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
    const int ndiag = GUARD ? 1 : slices;
    if (threadIdx.x == 0) {
        prefetch_tensormap(&maps.a);
        prefetch_tensormap(&maps.b);
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < ACC_STAGES; a++) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 16);
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
        // ---------------- TMA producer (both CTAs) ----------------
        if (lane == 0) {
            int n = 0;
            for (int t = cluster; t < tiles; t += clusters) {
                int tm, tn;
                tile_of(t, tiles_m, tiles_n, tm, tn);
                const int row_a = tm * PM + (int)rank * HALF, row_b = tn * PN + (int)rank * HALF;
                for (int di = 0; di < ndiag; di++) {
                    const int d = GUARD ? 0 : slices - 1 - di;
                    for (int p = 0; p <= d; p++) {
                        const int pa = GUARD ? slices : p, pb = GUARD ? slices : d - p;
                        const int wave = (t - cluster) / clusters;
                        const int in_wave = tiles - wave * clusters < clusters ? tiles - wave * clusters : clusters;
                        for (int kb = 0; kb < kblocks; kb++, n++) {
                            if (step_counters && kb == 0) {
                                // keep the producers of all CTAs within one slice pair of each other: the tiles in
                                // flight share their operand panels through L2 only while they stream the same planes
                                // (measured: 14.2 -> 12.5 ms at 8192^3; meeting four times per pair gains nothing
                                // more).  Every CTA of this persistent grid is resident, so the wait cannot deadlock.
                                unsigned int* ctr = step_counters + (wave * (MAX_SLICES * MAX_SLICES) + (GUARD ? 0 : (slices - 1 - d) * MAX_SLICES + p));
                                atomicAdd(ctr, 1u);
                                // (bounded: the rendezvous is an optimisation -- if the others do not show up within
                                //  ~1 s, e.g. because another context holds some SMs, carry on alone)
                                unsigned int seen, spins = 0;
                                do {
                                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
                                    if (seen < 2u * in_wave) __nanosleep(200);
                                } while (seen < 2u * in_wave && ++spins < 5000000u);
                            }
                            const int s = n % STAGES;
                            if (n >= STAGES) mbar_wait(&empty[s], ((n / STAGES) - 1) & 1);
                            uint8_t* st = base + (size_t)s * STAGE_BYTES;
                            if (rank == 0) mbar_expect_tx(&full[s], 2 * STAGE_BYTES);
                            tma_load_3d_pair(st, &maps.a, &full[s], kb * BKB, row_a, pa);
                            tma_load_3d_pair(st + TILE_BYTES, &maps.b, &full[s], kb * BKB, row_b, pb);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (leader CTA) ----------------
        if (rank == 0 && lane == 0) {
            constexpr uint32_t idesc = idesc_i8(PM, PN);
            int n = 0, it = 0;
            for (int t = cluster; t < tiles; t += clusters) {
                for (int di = 0; di < ndiag; di++, it++) {
                    const int d = GUARD ? 0 : slices - 1 - di;
                    const int as = it & 1;
                    if (it >= ACC_STAGES) mbar_wait(&acc_empty[as], ((it >> 1) - 1) & 1);
                    tc_fence_after();
                    const uint32_t dst = tmem_acc + (uint32_t)as * PN;
                    uint32_t started = 0;
                    for (int p = 0; p <= d; p++) {
                        for (int kb = 0; kb < kblocks; kb++, n++) {
                            const int s = n % STAGES;
                            mbar_wait(&full[s], (n / STAGES) & 1);
                            tc_fence_after();
                            const uint32_t st = smem_u32(base + (size_t)s * STAGE_BYTES);
                            const uint64_t da = umma_desc_k_sw128(st), db = umma_desc_k_sw128(st + TILE_BYTES);
#pragma unroll
                            for (int k = 0; k < BKB / 32; k++) {
                                umma_i8_pair(dst, da + 2 * k, db + 2 * k, idesc, started);
                                started = 1;
                            }
                            umma_commit_pair(&empty[s], 0b11);
                        }
                    }
                    umma_commit_pair(&acc_full[as], 0b11);
                }
            }
        }
    } else if (warp >= 4) {
        // ---------------- epilogue (both CTAs): int32 diagonal sum -> f64 contribution ----------------
        const int q = warp & 3, half = (warp - 4) >> 2;
        int it = 0;
        for (int t = cluster; t < tiles; t += clusters) {
            int tm, tn;
                tile_of(t, tiles_m, tiles_n, tm, tn);
            const int64_t row = (int64_t)tm * PM + (int64_t)rank * HALF + q * 32 + lane;
            const int64_t n0 = (int64_t)tn * PN + half * 128;
            const bool row_ok = row < M;
            const double rs = row_ok ? row_scale[row] : 0.0;
            const bool row_zero = row_ok && GUARD && row_exp[row] == ZERO_ROW;
            for (int di = 0; di < ndiag; di++, it++) {
                const int d = GUARD ? 0 : slices - 1 - di;
                const int as = it & 1;
                mbar_wait(&acc_full[as], (it >> 1) & 1);
                tc_fence_after();
                const uint32_t t0 = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * PN + half * 128);
                const double w = ldexp(rs, -12 - 7 * d);  // R_i 2^(-12 - 7d): a power of two, exact
                int lowest = INT_MAX;
#pragma unroll 1
                for (int c = 0; c < 4; c++) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t0 + (uint32_t)(c * 32), v);
                    tmem_ld_wait();
                    const int64_t col0 = n0 + c * 32;
                    if (!row_ok || col0 >= N) continue;
                    if (GUARD) {
                        if (!row_zero) {
#pragma unroll
                            for (int j = 0; j < 32; j++)
                                if (col0 + j < N && col_exp[col0 + j] != ZERO_ROW) lowest = min(lowest, (int)v[j]);
                        }
                    } else {
                        double* p = C + row * ldc + col0;
                        if (col0 + 32 <= N && (ldc & 1) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0) {
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                const double2 cs = *reinterpret_cast<const double2*>(col_scale + col0 + j);
                                double2 o = di == 0 ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2*>(p + j);
                                o.x += (double)(int)v[j] * w * cs.x;
                                o.y += (double)(int)v[j + 1] * w * cs.y;
                                *reinterpret_cast<double2*>(p + j) = o;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j++)
                                if (col0 + j < N) {
                                    const double o = di == 0 ? 0.0 : p[j];
                                    p[j] = o + (double)(int)v[j] * w * col_scale[col0 + j];
                                }
                        }
                    }
                }
                if (GUARD) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) lowest = min(lowest, __shfl_xor_sync(0xffffffffu, lowest, o));
                    if (lane == 0 && lowest != INT_MAX) atomicMin(flag + 1, lowest);
                }
                // (the next diagonal of this tile is read back by the same thread: program order is enough)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&acc_empty[as], 0);
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair(tmem_acc, TMEM_COLS);
}

int plane_map(CUtensorMap* map, const int8_t* base, int64_t rows, int64_t K, int64_t Kp, int planes) {
    uint64_t dims[3] = {(uint64_t)K, (uint64_t)rows, (uint64_t)planes};
    uint64_t strides[2] = {(uint64_t)Kp, (uint64_t)rows * (uint64_t)Kp};
    uint32_t box[3] = {(uint32_t)BKB, (uint32_t)HALF, 1};
    return make_tensor_map(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

template <bool GUARD>
int launch(const Maps& maps, double* C, int64_t M, int64_t N, int64_t K, int64_t ldc, int slices, const double* rs,
           const double* cs, const int* re, const int* ce, int threshold, int* flag, int sms, cudaStream_t stream,
           unsigned int* step_counters = nullptr) {
    auto kern = sliced_kernel<GUARD>;
    static thread_local int configured_dev = -1;
    int dev = 0;
    EML_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        EML_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        configured_dev = dev;
    }
    const int tiles_m = (int)((M + PM - 1) / PM), tiles_n = (int)((N + PN - 1) / PN);
    int clusters = sms / 2;
    if ((int64_t)tiles_m * tiles_n < clusters) clusters = tiles_m * tiles_n;
    EML_CUDA(launch_k(kern, dim3(2 * clusters), dim3(THREADS), SMEM_BYTES, stream, maps, C, M, N, ldc,
                      (int)((K + BKB - 1) / BKB), tiles_m, tiles_n, slices, rs, cs, re, ce, threshold, flag, step_counters));
    count_launch();
    return check_launch(GUARD ? "sliced f64 GEMM (guard)" : "sliced f64 GEMM");
}
}  // namespace sl
}  // namespace

// One operand cut into slice planes: A's rows (k contiguous in the source) or B's columns (n contiguous in the source).
// `bad` (device int) is raised by the exponent scan for Inf / NaN / extreme exponents.
int sliced_prepare(const double* X, bool rows_of_a, int64_t mn, int64_t K, int64_t ld, int slices, cudaStream_t stream,
                   SlicedSide* out) {
    using namespace sl;
    out->mn = mn;
    out->K = K;
    out->Kp = (K + BKB - 1) / BKB * BKB;
    out->slices = slices;
    EML_TRY(out->planes.alloc((size_t)(slices + 1) * mn * out->Kp, stream));
    EML_TRY(out->ints.alloc(sizeof(int) * (size_t)(mn + 8), stream));
    const int64_t mne = (mn + 1) & ~int64_t(1);
    EML_TRY(out->scales.alloc(sizeof(double) * (size_t)(2 * mne), stream));
    double* inv = out->scales.as<double>() + mne;
    int* expo = out->ints.as<int>();
    int* bad = expo + mn;
    EML_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), stream));
    if (rows_of_a) {
        if (mn > 65535) EML_BAD_ARG("sliced gemm: more than 65535 rows in one call");
        EML_CUDA(launch_k(row_exponent_kernel, dim3((unsigned)((mn + 7) / 8)), dim3(256), 0, stream, X, mn, K, ld, expo, bad));
        EML_CUDA(launch_k(scale_kernel, dim3((unsigned)((mn + 255) / 256)), dim3(256), 0, stream, expo, out->scales.as<double>(), inv, mn));
        EML_CUDA(launch_k(slice_rows_kernel, dim3((unsigned)((out->Kp / 4 + 255) / 256), (unsigned)mn), dim3(256), 0, stream, X,
                          mn, K, ld, inv, out->planes.as<int8_t>(), out->Kp, slices));
        count_launch(3);
    } else {
        const int64_t rpb = 256;
        EML_CUDA(launch_k(fill_int_kernel, dim3((unsigned)((mn + 255) / 256)), dim3(256), 0, stream, expo, ZERO_ROW, mn));
        EML_CUDA(launch_k(col_exponent_kernel, dim3((unsigned)((mn + 255) / 256), (unsigned)((K + rpb - 1) / rpb)), dim3(256), 0,
                          stream, X, K, mn, ld, rpb, expo, bad));
        EML_CUDA(launch_k(scale_kernel, dim3((unsigned)((mn + 255) / 256)), dim3(256), 0, stream, expo, out->scales.as<double>(), inv, mn));
        EML_CUDA(launch_k(slice_cols_kernel, dim3((unsigned)(out->Kp / 128), (unsigned)((mn + 31) / 32)), dim3(256), 0, stream, X,
                          K, mn, ld, inv, out->planes.as<int8_t>(), out->Kp, slices));
        count_launch(4);
    }
    return check_launch("sliced f64 GEMM (slicing)");
}

// C = A * B from prepared operands, using `use_slices` of the prepared slices (0 = as few as the guard can prove,
// at least 7).  *path = number of slices used when the guard proved the contract and the sliced kernel wrote C;
// *path = 0: nothing was written (the caller runs the DMMA kernel).  Synchronises `stream` once (the guard's verdict).
int sliced_multiply(DeviceCtx* ctx, const SlicedSide& a, const SlicedSide& b, double* C, int64_t ldc, int use_slices,
                    int* path, cudaStream_t stream) {
    using namespace sl;
    *path = 0;
    const int64_t M = a.mn, N = b.mn, K = a.K;
    const int prepared = a.slices;
    Maps maps;
    EML_TRY(plane_map(&maps.a, a.planes.as<int8_t>(), M, K, a.Kp, prepared + 1));
    EML_TRY(plane_map(&maps.b, b.planes.as<int8_t>(), N, K, b.Kp, prepared + 1));
    const int sms = ctx ? ctx->sm_count : 148;
    int* a_exp = a.ints.as<int>();
    int* b_exp = b.ints.as<int>();
    int* flag = a_exp + M;  // [0]: A's `bad` word, [1]: the guard's minimum; B's `bad` word is read separately
    EML_CUDA(launch_k(fill_int_kernel, dim3(1), dim3(32), 0, stream, flag + 1, INT_MAX, (int64_t)1));
    // guard: the smallest sum_k floor(64|x|) floor(64|y|) over the elements of C.  The guard kernel reads plane
    // `prepared` (the floor planes) whatever the slice count that will be used.
    EML_TRY(launch<true>(maps, C, M, N, K, ldc, prepared, a.scales.as<double>(), b.scales.as<double>(), a_exp, b_exp, 0,
                         flag, sms, stream));
    int host_flags[3] = {0, 0, 0};
    EML_CUDA(cudaMemcpyAsync(&host_flags[0], flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
    EML_CUDA(cudaMemcpyAsync(&host_flags[2], b_exp + N, sizeof(int), cudaMemcpyDeviceToHost, stream));
    EML_CUDA(cudaStreamSynchronize(stream));
    if (host_flags[0] != 0 || host_flags[2] != 0) return EML_OK;  // Inf / NaN / extreme exponents
    // s slices are proven when every sum >= K (s + 2) 2^(12 - 7s) / 0.9e-12   (and the int32 sums cannot overflow)
    auto proven = [&](int s) {
        return (double)host_flags[1] >= ceil((double)K * (s + 2) * ldexp(1.0, 12 - 7 * s) / 0.9e-12) &&
               (int64_t)s * K * 4096 < (int64_t(1) << 31);
    };
    int s = 0;
    if (use_slices > 0) {
        if (use_slices <= prepared && proven(use_slices)) s = use_slices;
    } else {
        for (int c = 7; c <= prepared && !s; c++)
            if (proven(c)) s = c;
    }
    if (!s) return EML_OK;
    // one counter per (wave of tiles, slice pair): see the producer
    TempBuf counters;
    unsigned int* ctr = nullptr;
    if (!getenv("EML_SLICED_NO_LOCKSTEP")) {
        const int64_t tiles = ((M + PM - 1) / PM) * ((N + PN - 1) / PN);
        const int64_t clusters = tiles < sms / 2 ? tiles : sms / 2;
        const size_t bytes = sizeof(unsigned int) * MAX_SLICES * MAX_SLICES * (size_t)((tiles + clusters - 1) / clusters);
        EML_TRY(counters.alloc(bytes, stream));
        EML_CUDA(cudaMemsetAsync(counters.p, 0, bytes, stream));
        ctr = counters.as<unsigned int>();
    }
    // the digits of an s-slice expansion are the first s digits of the prepared one; plane `prepared` is not touched
    EML_TRY(launch<false>(maps, C, M, N, K, ldc, s, a.scales.as<double>(), b.scales.as<double>(), a_exp, b_exp, 0, flag, sms,
                          stream, ctr));
    set_last_kernel("sliced_int8_f64");
    *path = s;
    return EML_OK;
}

bool sliced_shape_ok(int64_t M, int64_t N, int64_t K, int slices) {
    // int32 accumulators hold `slices` pair products of K terms of at most 2^12 each
    return (int64_t)slices * K * 4096 < (int64_t(1) << 31) && M >= 256 && N >= 256 && K >= 256 && M <= 65535;
}

// C = A * B, row-major f64 on the device.  *path = 1 when the int8-sliced kernel produced C, 0 when the operands did
// not pass the guard (or the shape is outside the kernel's range) and the DMMA kernel ran instead.
int gemm_f64_sliced_dev(DeviceCtx* ctx, const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K,
                        int64_t lda, int64_t ldb, int64_t ldc, int slices, int* path, cudaStream_t stream) {
    if (path) *path = 0;
    if (slices != 0 && (slices < 6 || slices > sl::MAX_SLICES))
        EML_BAD_ARG("sliced gemm: 6..%d slices, or 0 for as few as the guard can prove (got %d)", sl::MAX_SLICES, slices);
    if (M < 1 || N < 1 || K < 1 || lda < K || ldb < N || ldc < N) EML_BAD_ARG("sliced gemm: bad shape");
    const int prepare = slices ? slices : 8;
    int took = 0;
    // the guard's verdict is read on the host: a stream that is being captured into a graph cannot be synchronised
    cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
    EML_CUDA(cudaStreamIsCapturing(stream, &capturing));
    if (capturing == cudaStreamCaptureStatusNone && sliced_shape_ok(M, N, K, slices ? slices : 7)) {
        SlicedSide a, b;
        EML_TRY(sliced_prepare(A, true, M, K, lda, prepare, stream, &a));
        EML_TRY(sliced_prepare(B, false, N, K, ldb, prepare, stream, &b));
        EML_TRY(sliced_multiply(ctx, a, b, C, ldc, slices, &took, stream));
    }
    if (!took)
        return contract_dev<double>(ctx, A, B, C, 1, M, N, K, Strides3{0, lda, 1}, Strides3{0, ldb, 1}, Strides3{0, ldc, 1},
                                    false, false, stream);
    if (path) *path = took;
    return EML_OK;
}

}  // namespace eml
