// reorder.cu -- Tensor::reorder / transpose (src/tensors/mod.rs:1473-1600) and Matrix::transpose
// (src/matrices/mod.rs:1277-1330) on device-resident buffers: out (row-major over out_lens) [i_0 .. i_{r-1}] =
// in[sum_d i_d * in_strides[d]].  HBM-bound: every element is read once and written once.
//
// A permutation of a contiguous tensor always leaves one output dimension whose INPUT stride is 1.  When that is
// the last output dimension, consecutive threads read and write consecutive addresses already.  Otherwise the
// kernel moves 32 x 32 tiles spanned by (that dimension, the last output dimension) through shared memory, so the
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

// tile over (dimension `a`: input stride 1, the last dimension `z`: output stride 1); blockIdx.z walks the rest
template <class T>
__global__ void __launch_bounds__(256)
reorder_tiled_kernel(const T* __restrict__ in, T* __restrict__ out, Dims d, int a, int64_t rest_total) {
    pdl_prologue();
    __shared__ T tile[32][33];
    const int z = d.rank - 1;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int64_t z0 = (int64_t)blockIdx.x * 32, a0 = (int64_t)blockIdx.y * 32;
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
#pragma unroll
        for (int j = ty; j < 32; j += 8) {
            const int64_t iz = z0 + j, ia = a0 + tx;
            if (iz < d.len[z] && ia < d.len[a]) tile[j][tx] = in[src + iz * d.istr[z] + ia];  // istr[a] == 1
        }
        __syncthreads();
#pragma unroll
        for (int j = ty; j < 32; j += 8) {
            const int64_t ia = a0 + j, iz = z0 + tx;
            if (iz < d.len[z] && ia < d.len[a]) out[dst + ia * d.ostr[a] + iz] = tile[tx][j];
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
        const int64_t gx = (d.len[rank - 1] + 31) / 32, gy = (d.len[a] + 31) / 32;
        if (gx <= 0x7fffffff && gy <= 65535) {
            const unsigned gz = (unsigned)(rest < 65535 ? rest : 65535);
            EML_CUDA(launch_k(reorder_tiled_kernel<T>, dim3((unsigned)gx, (unsigned)gy, gz), dim3(256), 0, stream, in, out, d, a, rest));
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

template int launch_reorder<float>(const float*, float*, int, const int64_t*, const int64_t*, cudaStream_t);
template int launch_reorder<double>(const double*, double*, int, const int64_t*, const int64_t*, cudaStream_t);

}  // namespace eml
