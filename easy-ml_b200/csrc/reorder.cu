// reorder.cu -- Tensor::reorder / transpose (src/tensors/mod.rs:1473-1600) and Matrix::transpose
// (src/matrices/mod.rs:1277-1330) on device-resident buffers: out (row-major over out_lens) [i_0 .. i_{r-1}] =
// in[sum_d i_d * in_strides[d]].  HBM-bound: every element is read once and written once.
//
// A permutation of a contiguous tensor always leaves one output dimension whose INPUT stride is 1.  When that is
// the last output dimension, consecutive threads read and write consecutive addresses already.  Otherwise the
// kernel moves 64 x 64 tiles spanned by (that dimension, the last output dimension) through shared memory, so the
// reads run along the input's contiguous index and the writes along the output's: both coalesced.
#include "common.h"

namespace eml {

namespace {
constexpr int MAX_RANK = 6;

struct Dims {
    int rank;
    int64_t len[MAX_RANK];   // output lengths
    int64_t istr[MAX_RANK];  // input stride of each output dimension
    int64_t ostr[MAX_RANK];  // output (row-major) strides
};

template <class T>
__global__ void __launch_bounds__(256)
reorder_gather_kernel(const T* __restrict__ in, T* __restrict__ out, Dims d, int64_t total) {
    pdl_prologue();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t rest = i, src = 0;
#pragma unroll
        for (int k = MAX_RANK - 1; k >= 0; k--) {
            if (k < d.rank) {
                src += (rest % d.len[k]) * d.istr[k];
                rest /= d.len[k];
            }
        }
        out[i] = in[src];
    }
}

// tile over (dimension `a`: input stride 1, the last dimension `z`: output stride 1); blockIdx.z walks the rest.
// 64 x 64 tiles, 16 elements per thread, every load of a tile issued before the first store to shared memory:
// 16 KB (f32) / 32 KB (f64) in flight per CTA, 256-byte (f32) runs on both sides.
constexpr int TILE = 64, TILE_ROWS_PER_PASS = 4;  // 256 threads = 64 x 4
template <class T>
__global__ void __launch_bounds__(256, 4)
reorder_tiled_kernel(const T* __restrict__ in, T* __restrict__ out, Dims d, int a, int64_t rest_total, int64_t len_z,
                     int64_t len_a, int64_t istr_z, int64_t ostr_a) {
    // (the four scalars are d.len[z], d.len[a], d.istr[z], d.ostr[a]: indexing the parameter struct with a run-time
    //  index would make the compiler copy it to local memory)
    pdl_prologue();
    __shared__ T tile[TILE][TILE + 1];
    const int z = d.rank - 1;
    const int tx = threadIdx.x & (TILE - 1), ty = threadIdx.x / TILE;
    const int64_t z0 = (int64_t)blockIdx.x * TILE, a0 = (int64_t)blockIdx.y * TILE;
    for (int64_t rest = blockIdx.z; rest < rest_total; rest += gridDim.z) {
        int64_t r = rest, src = 0, dst = 0;
#pragma unroll
        for (int k = MAX_RANK - 1; k >= 0; k--) {
            if (k < d.rank && k != a && k != z) {
                const int64_t ix = r % d.len[k];
                r /= d.len[k];
                src += ix * d.istr[k];
                dst += ix * d.ostr[k];
            }
        }
        T v[TILE / TILE_ROWS_PER_PASS];
        const bool a_ok = a0 + tx < len_a;
        const T* ip = in + src + (z0 + ty) * istr_z + a0 + tx;  // istr[a] == 1
#pragma unroll
        for (int p = 0; p < TILE / TILE_ROWS_PER_PASS; p++) {
            v[p] = (a_ok && z0 + ty + p * TILE_ROWS_PER_PASS < len_z) ? *ip : T(0);
            ip += TILE_ROWS_PER_PASS * istr_z;
        }
#pragma unroll
        for (int p = 0; p < TILE / TILE_ROWS_PER_PASS; p++) tile[ty + p * TILE_ROWS_PER_PASS][tx] = v[p];
        __syncthreads();
        const bool z_ok = z0 + tx < len_z;
        T* op = out + dst + (a0 + ty) * ostr_a + z0 + tx;
#pragma unroll
        for (int p = 0; p < TILE / TILE_ROWS_PER_PASS; p++) {
            if (z_ok && a0 + ty + p * TILE_ROWS_PER_PASS < len_a) *op = tile[tx][ty + p * TILE_ROWS_PER_PASS];
            op += TILE_ROWS_PER_PASS * ostr_a;
        }
        __syncthreads();
    }
}
}  // namespace

template <class T>
int launch_reorder(const T* in, T* out, int rank, const int64_t* out_lens, const int64_t* in_strides,
                   cudaStream_t stream) {
    if (rank < 1 || rank > MAX_RANK) EML_BAD_ARG("reorder: rank %d outside 1..%d", rank, MAX_RANK);
    if (!in || !out || !out_lens || !in_strides) EML_BAD_ARG("reorder: null argument");
    Dims d{};
    d.rank = rank;
    int64_t total = 1;
    for (int k = rank - 1; k >= 0; k--) {
        if (out_lens[k] < 1) EML_BAD_ARG("reorder: dimension %d has length %lld (every length is >= 1)", k, (long long)out_lens[k]);
        if (in_strides[k] < 0) EML_BAD_ARG("reorder: negative stride");
        d.len[k] = out_lens[k];
        d.istr[k] = in_strides[k];
        d.ostr[k] = total;
        total *= out_lens[k];
    }
    int a = -1;  // an output dimension (not the last) that the input walks with stride 1
    if (rank >= 2 && d.istr[rank - 1] != 1)
        for (int k = 0; k < rank - 1; k++)
            if (d.istr[k] == 1 && d.len[k] > 1) a = k;
    if (a >= 0 && d.len[rank - 1] > 1) {
        const int64_t rest = total / (d.len[a] * d.len[rank - 1]);
        const int64_t gx = (d.len[rank - 1] + TILE - 1) / TILE, gy = (d.len[a] + TILE - 1) / TILE;
        if (gx <= 0x7fffffff && gy <= 65535) {
            const unsigned gz = (unsigned)(rest < 65535 ? rest : 65535);
            EML_CUDA(launch_k(reorder_tiled_kernel<T>, dim3((unsigned)gx, (unsigned)gy, gz), dim3(256), 0, stream, in, out, d, a, rest,
                              d.len[rank - 1], d.len[a], d.istr[rank - 1], d.ostr[a]));
            count_launch();
            set_last_kernel("reorder_tiled");
            return check_launch("reorder (tiled)");
        }
    }
    const int64_t want = (total + 255) / 256;
    const unsigned blocks = (unsigned)(want < 148 * 16 ? want : 148 * 16);
    EML_CUDA(launch_k(reorder_gather_kernel<T>, dim3(blocks), dim3(256), 0, stream, in, out, d, total));
    count_launch();
    set_last_kernel("reorder_gather");
    return check_launch("reorder (gather)");
}

namespace {
// C[j][i] = C[i][j] for every i < j: 32 x 32 tiles on and above the diagonal are read along their rows and written,
// transposed through shared memory, along the rows of the mirrored tile
template <class T>
__global__ void __launch_bounds__(256)
mirror_upper_kernel(T* __restrict__ C, int64_t n, int64_t ldc) {
    pdl_prologue();
    if (blockIdx.x < blockIdx.y) return;  // tile column < tile row: below the diagonal
    __shared__ T tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t i0 = (int64_t)blockIdx.y * 32, j0 = (int64_t)blockIdx.x * 32;
#pragma unroll
    for (int r = ty; r < 32; r += 8)
        if (i0 + r < n && j0 + tx < n) tile[r][tx] = C[(i0 + r) * ldc + j0 + tx];
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int64_t row = j0 + r, col = i0 + tx;  // element (row, col) of the lower triangle <- tile[tx][r] = C[col][row]
        if (row < n && col < row) C[row * ldc + col] = tile[tx][r];
    }
}
}  // namespace

template <class T>
int launch_mirror_upper(T* C, int64_t n, int64_t ldc, cudaStream_t stream) {
    const int64_t t = (n + 31) / 32;
    if (t > 65535) EML_BAD_ARG("mirror: matrix too large");
    EML_CUDA(launch_k(mirror_upper_kernel<T>, dim3((unsigned)t, (unsigned)t), dim3(256), 0, stream, C, n, ldc));
    count_launch();
    return check_launch("mirror upper triangle");
}
template int launch_mirror_upper<float>(float*, int64_t, int64_t, cudaStream_t);
template int launch_mirror_upper<double>(double*, int64_t, int64_t, cudaStream_t);

template int launch_reorder<float>(const float*, float*, int, const int64_t*, const int64_t*, cudaStream_t);
template int launch_reorder<double>(const double*, double*, int, const int64_t*, const int64_t*, cudaStream_t);

}  // namespace eml
