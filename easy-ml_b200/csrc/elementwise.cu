// elementwise.cu -- HBM-bound kernels around the GEMMs: container unary/binary ops with their tape
// derivatives in one pass, the backward chain rule, SGD, deterministic reductions, generic einsum.
//
// Compiled with -fmad=false: +,-,*,/ stay separately rounded IEEE operations, so these kernels are
// bit-exact with the reference's rules (src/differentiation/functions.rs) for the arithmetic ops;
// exp/ln/sin/cos/pow differ from Rust's libm by the usual ulp or two (tolerance stated in the tests).
#include <type_traits>

#include <map>
#include <mutex>
#include <utility>

#include <cstdlib>

#include "common.h"
#include "ew_rules.cuh"
#include "tf32_split.cuh"

namespace eml {

namespace {

constexpr int EW_THREADS = 256;

inline int ew_grid(int64_t n, int per_thread) {
    int64_t blocks = (n + (int64_t)EW_THREADS * per_thread - 1) / ((int64_t)EW_THREADS * per_thread);
    const int64_t cap = 148 * 16;  // 16 resident 256-thread CTAs per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <class T, int OP, bool VEC>
__global__ void __launch_bounds__(EW_THREADS)
unary_kernel(T c, const T* __restrict__ x, T* __restrict__ y, T* __restrict__ dydx, int64_t n) {
    pdl_prologue();
    constexpr int V = VEC ? Vec<T>::N : 1;
    int64_t stride = (int64_t)gridDim.x * blockDim.x * V;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V; i < n; i += stride) {
        if (VEC && i + V <= n) {
            T xv[Vec<T>::N], yv[Vec<T>::N], dv[Vec<T>::N];
            vload<T>(x + i, xv);
#pragma unroll
            for (int j = 0; j < Vec<T>::N; j++) unary_rule<T, OP>(c, xv[j], yv[j], dv[j]);
            if (y) vstore<T>(y + i, yv);
            if (dydx) vstore<T>(dydx + i, dv);
        } else {
            for (int j = 0; j < V && i + j < n; j++) {
                T yy, dd;
                unary_rule<T, OP>(c, x[i + j], yy, dd);
                if (y) y[i + j] = yy;
                if (dydx) dydx[i + j] = dd;
            }
        }
    }
}

template <class T, int OP, bool VEC>
__global__ void __launch_bounds__(EW_THREADS)
binary_kernel(const T* __restrict__ x, const T* __restrict__ y, T* __restrict__ z, T* __restrict__ dzdx,
              T* __restrict__ dzdy, int64_t n) {
    pdl_prologue();
    constexpr int V = VEC ? Vec<T>::N : 1;
    int64_t stride = (int64_t)gridDim.x * blockDim.x * V;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V; i < n; i += stride) {
        if (VEC && i + V <= n) {
            T xv[Vec<T>::N], yv[Vec<T>::N], zv[Vec<T>::N], ax[Vec<T>::N], ay[Vec<T>::N];
            vload<T>(x + i, xv);
            vload<T>(y + i, yv);
#pragma unroll
            for (int j = 0; j < Vec<T>::N; j++) binary_rule<T, OP>(xv[j], yv[j], zv[j], ax[j], ay[j]);
            if (z) vstore<T>(z + i, zv);
            if (dzdx) vstore<T>(dzdx + i, ax);
            if (dzdy) vstore<T>(dzdy + i, ay);
        } else {
            for (int j = 0; j < V && i + j < n; j++) {
                T zz, a, b;
                binary_rule<T, OP>(x[i + j], y[i + j], zz, a, b);
                if (z) z[i + j] = zz;
                if (dzdx) dzdx[i + j] = a;
                if (dzdy) dzdy[i + j] = b;
            }
        }
    }
}

// MODE 0: dparent += dout*deriv   1: w -= lr*d  (lr in c)   2: out = c   3: dst += src   4: out = b - lr*a
//      5: dparent = dout*deriv + 0   6: dparent += dout*c   7: dparent = dout*c + 0
// (5 and 7 are the FIRST contribution to a derivative that the reference holds as 0 at that point: 0 + x, so a
//  product of -0 still ends up as +0; the buffer needs no zero fill and is not read)
template <class T, int MODE, bool VEC>
__global__ void __launch_bounds__(EW_THREADS)
axpy_kernel(T c, const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int64_t n) {
    pdl_prologue();
    constexpr int V = VEC ? Vec<T>::N : 1;
    int64_t stride = (int64_t)gridDim.x * blockDim.x * V;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V; i < n; i += stride) {
        if (VEC && i + V <= n) {
            T av[Vec<T>::N], bv[Vec<T>::N], ov[Vec<T>::N];
            if (MODE != 2) vload<T>(a + i, av);
            if (MODE == 0 || MODE == 4 || MODE == 5) vload<T>(b + i, bv);
            if (MODE == 0 || MODE == 1 || MODE == 3 || MODE == 6) vload<T>(out + i, ov);
#pragma unroll
            for (int j = 0; j < Vec<T>::N; j++) {
                if (MODE == 0) ov[j] = ov[j] + av[j] * bv[j];
                if (MODE == 1) ov[j] = ov[j] - av[j] * c;
                if (MODE == 2) ov[j] = c;
                if (MODE == 3) ov[j] = ov[j] + av[j];
                if (MODE == 4) ov[j] = bv[j] - av[j] * c;
                if (MODE == 5) ov[j] = av[j] * bv[j] + T(0);
                if (MODE == 6) ov[j] = ov[j] + av[j] * c;
                if (MODE == 7) ov[j] = av[j] * c + T(0);
            }
            vstore<T>(out + i, ov);
        } else {
            for (int j = 0; j < V && i + j < n; j++) {
                int64_t q = i + j;
                if (MODE == 0) out[q] = out[q] + a[q] * b[q];
                if (MODE == 1) out[q] = out[q] - a[q] * c;
                if (MODE == 2) out[q] = c;
                if (MODE == 3) out[q] = out[q] + a[q];
                if (MODE == 4) out[q] = b[q] - a[q] * c;
                if (MODE == 5) out[q] = a[q] * b[q] + T(0);
                if (MODE == 6) out[q] = out[q] + a[q] * c;
                if (MODE == 7) out[q] = a[q] * c + T(0);
            }
        }
    }
}

// ---- deterministic reductions ----
constexpr int RED_THREADS = 256;

template <class T>
__device__ __forceinline__ T block_reduce(T v, T* smem) {
    // fixed tree: warp shuffles then the 8 warp sums added in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) smem[w] = v;
    __syncthreads();
    T total = T(0);
    if (threadIdx.x == 0)
        for (int i = 0; i < RED_THREADS / 32; i++) total += smem[i];
    __syncthreads();
    return total;  // valid in thread 0
}

// Single-pass deterministic reductions.  Every block writes its partial, takes a ticket from a per-stream counter
// (atomicInc wraps to zero on the last ticket, so the counter resets itself), and the block that drew the last
// ticket adds the partials in index order.  The atomic only elects that block; the sum order is fixed, so the
// result is bit-identical run to run.  One launch instead of two.
constexpr int MAX_TICKETS = 4096;

// block b sums elements [b*chunk, min(n,(b+1)*chunk)) of x[i*incx] * (y ? y[i*incy] : 1); last block: out (+)= total
template <class T>
__global__ void __launch_bounds__(RED_THREADS)
reduce_kernel(const T* __restrict__ x, int64_t incx, const T* __restrict__ y, int64_t incy, int64_t n, int64_t chunk,
              T* partials, unsigned* ticket, T* out, int accumulate) {
    pdl_prologue();
    __shared__ T smem[RED_THREADS / 32];
    __shared__ bool last;
    int64_t lo = (int64_t)blockIdx.x * chunk, hi = lo + chunk < n ? lo + chunk : n;
    T acc = T(0);
    for (int64_t i = lo + threadIdx.x; i < hi; i += RED_THREADS) {
        T v = x[i * incx];
        if (y) v = v * y[i * incy];
        acc += v;
    }
    T total = block_reduce<T>(acc, smem);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = total;
        __threadfence();
        last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    acc = T(0);
    for (int i = threadIdx.x; i < (int)gridDim.x; i += RED_THREADS) acc += __ldcg(partials + i);
    total = block_reduce<T>(acc, smem);
    if (threadIdx.x == 0) out[0] = accumulate ? out[0] + total : total;
}

// zero-initialised, self-resetting ticket counters, one array per (device, stream): reductions on one stream are
// ordered, so they can share it
int ticket_counters(cudaStream_t stream, unsigned** out) {
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, unsigned*> table;
    int dev = 0;
    EML_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(dev, stream);
    auto it = table.find(key);
    if (it == table.end()) {
        unsigned* p = nullptr;
        EML_CUDA(cudaMalloc(&p, sizeof(unsigned) * MAX_TICKETS));
        EML_CUDA(cudaMemset(p, 0, sizeof(unsigned) * MAX_TICKETS));
        it = table.emplace(key, p).first;
    }
    *out = it->second;
    return EML_OK;
}

template <class T>
int reduce_impl(const T* x, int64_t incx, const T* y, int64_t incy, int64_t n, T* out, bool accumulate,
                cudaStream_t stream) {
    int blocks = (int)((n + 4095) / 4096);
    if (blocks > 1024) blocks = 1024;
    if (blocks < 1) blocks = 1;
    int64_t chunk = (n + blocks - 1) / blocks;
    TempBuf partials;
    EML_TRY(partials.alloc(sizeof(T) * blocks, stream));
    unsigned* ticket = nullptr;
    EML_TRY(ticket_counters(stream, &ticket));
    launch_k(reduce_kernel<T>, dim3(blocks), dim3(RED_THREADS), 0, stream, x, incx, y, incy, n, chunk, partials.as<T>(), ticket, out,
                                                         accumulate ? 1 : 0);
    count_launch();
    return check_launch("reduce");
}

// column sums: partial[chunk][c] = sum over the chunk's rows; the last chunk-block of a column block adds the
// chunks in index order
template <class T>
__global__ void __launch_bounds__(256)
col_sums_kernel(const T* __restrict__ x, int64_t rows, int64_t cols, int64_t rows_per_chunk, T* partial,
                unsigned* tickets, T* __restrict__ out) {
    pdl_prologue();
    __shared__ T smem[8][33];
    __shared__ bool last;
    int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    int64_t c = (int64_t)blockIdx.x * 32 + tx;
    int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk, r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
    T acc = T(0);
    if (c < cols)
        for (int64_t r = r0 + ty; r < r1; r += 8) acc += x[r * cols + c];
    smem[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && c < cols) {
        T t = T(0);
        for (int i = 0; i < 8; i++) t += smem[i][tx];
        partial[(int64_t)blockIdx.y * cols + c] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicInc(tickets + blockIdx.x, gridDim.y - 1) == gridDim.y - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (ty == 0 && c < cols) {
        T t = T(0);
        for (int i = 0; i < (int)gridDim.y; i++) t += __ldcg(partial + (int64_t)i * cols + c);
        out[c] = t;
    }
}

template <class T>
__global__ void __launch_bounds__(256) row_sums_kernel(const T* __restrict__ x, int64_t rows, int64_t cols, T* __restrict__ out) {
    pdl_prologue();
    int lane = threadIdx.x & 31;
    int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    T acc = T(0);
    for (int64_t c = lane; c < cols; c += 32) acc += x[r * cols + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) out[r] = acc;
}

// ---- generic two-operand einsum (bit-exact with the reference's loop nest) ----
constexpr int MAX_OUT = 6, MAX_SUM = 12;
struct EinsumParams {
    int n_out, n_sum;
    int64_t total;
    int64_t out_len[MAX_OUT], a_out[MAX_OUT], b_out[MAX_OUT];
    int64_t sum_len[MAX_SUM], a_sum[MAX_SUM], b_sum[MAX_SUM];
    int64_t sum_total;
};

template <class T>
__global__ void __launch_bounds__(256) einsum2_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ out, EinsumParams p) {
    pdl_prologue();
    for (int64_t flat = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; flat < p.total;
         flat += (int64_t)gridDim.x * blockDim.x) {
        int64_t rem = flat, a_off = 0, b_off = 0;
        for (int d = p.n_out - 1; d >= 0; d--) {
            int64_t i = rem % p.out_len[d];
            rem /= p.out_len[d];
            a_off += i * p.a_out[d];
            b_off += i * p.b_out[d];
        }
        T sum = T(0);  // reference src/tensors/einsum.rs:932
        if (p.n_sum == 0) {
            sum = sum + A[a_off] * B[b_off];  // :945 (fmad off: separate roundings)
        } else {
            int64_t idx[MAX_SUM];
            for (int d = 0; d < p.n_sum; d++) idx[d] = 0;
            for (int64_t s = 0; s < p.sum_total; s++) {
                sum = sum + A[a_off] * B[b_off];  // :968
                // odometer, last dimension fastest (reference src/tensors/indexing.rs:905-925)
                for (int d = p.n_sum - 1; d >= 0; d--) {
                    a_off += p.a_sum[d];
                    b_off += p.b_sum[d];
                    if (++idx[d] < p.sum_len[d]) break;
                    a_off -= p.a_sum[d] * p.sum_len[d];
                    b_off -= p.b_sum[d] * p.sum_len[d];
                    idx[d] = 0;
                }
            }
        }
        out[flat] = sum;
    }
}

// ---- generic N-operand einsum, N = 1..6 (Einsum1 / Einsum3..6, reference src/tensors/einsum.rs:829-883,
// :1019-1666): out[o] = sum_s ((x1 * x2) * x3) ... in the reference's order -- zero-seeded sum, summation
// dimensions in order of first appearance with the last fastest, left-associated product, every operation
// separately rounded (this file is compiled -fmad=false): bit-exact with the reference.
constexpr int MAX_OPS = 6, MAX_SUM_N = 36;  // six rank-6 operands can carry 36 distinct summed names
struct EinsumNParams {
    int n_ops, n_out, n_sum;
    int64_t total, sum_total;
    const void* x[MAX_OPS];
    int64_t out_len[MAX_OUT], sum_len[MAX_SUM_N];
    int64_t out_stride[MAX_OPS][MAX_OUT];
    int64_t sum_stride[MAX_OPS][MAX_SUM_N];
};

template <class T>
__global__ void __launch_bounds__(256) einsum_n_kernel(T* __restrict__ out, const __grid_constant__ EinsumNParams p) {
    pdl_prologue();
    for (int64_t flat = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; flat < p.total;
         flat += (int64_t)gridDim.x * blockDim.x) {
        int64_t off[MAX_OPS];
#pragma unroll
        for (int i = 0; i < MAX_OPS; i++) off[i] = 0;
        int64_t rem = flat;
        for (int d = p.n_out - 1; d >= 0; d--) {
            int64_t ix = rem % p.out_len[d];
            rem /= p.out_len[d];
#pragma unroll
            for (int i = 0; i < MAX_OPS; i++)
                if (i < p.n_ops) off[i] += ix * p.out_stride[i][d];
        }
        auto product = [&]() {
            T v = static_cast<const T*>(p.x[0])[off[0]];
#pragma unroll
            for (int i = 1; i < MAX_OPS; i++)
                if (i < p.n_ops) v = v * static_cast<const T*>(p.x[i])[off[i]];  // ((p1 * p2) * p3) ...
            return v;
        };
        T sum = T(0);
        if (p.n_sum == 0) {
            sum = sum + product();
        } else {
            int idx[MAX_SUM_N];
            for (int d = 0; d < p.n_sum; d++) idx[d] = 0;
            for (int64_t s = 0; s < p.sum_total; s++) {
                sum = sum + product();
                for (int d = p.n_sum - 1; d >= 0; d--) {  // odometer, last dimension fastest
#pragma unroll
                    for (int i = 0; i < MAX_OPS; i++)
                        if (i < p.n_ops) off[i] += p.sum_stride[i][d];
                    if (++idx[d] < p.sum_len[d]) break;
#pragma unroll
                    for (int i = 0; i < MAX_OPS; i++)
                        if (i < p.n_ops) off[i] -= p.sum_stride[i][d] * p.sum_len[d];
                    idx[d] = 0;
                }
            }
        }
        out[flat] = sum;
    }
}

template <class T, int OP>
int launch_unary_op(T c, const T* x, T* y, T* dydx, int64_t n, cudaStream_t stream) {
    bool vec = aligned16(x) && (!y || aligned16(y)) && (!dydx || aligned16(dydx));
    if (vec)
        launch_k(unary_kernel<T, OP, true>, dim3(ew_grid(n, Vec<T>::N)), dim3(EW_THREADS), 0, stream, c, x, y, dydx, n);
    else
        launch_k(unary_kernel<T, OP, false>, dim3(ew_grid(n, 1)), dim3(EW_THREADS), 0, stream, c, x, y, dydx, n);
    count_launch();
    return check_launch("unary");
}

template <class T, int OP>
int launch_binary_op(const T* x, const T* y, T* z, T* dzdx, T* dzdy, int64_t n, cudaStream_t stream) {
    bool vec = aligned16(x) && aligned16(y) && (!z || aligned16(z)) && (!dzdx || aligned16(dzdx)) &&
               (!dzdy || aligned16(dzdy));
    if (vec)
        launch_k(binary_kernel<T, OP, true>, dim3(ew_grid(n, Vec<T>::N)), dim3(EW_THREADS), 0, stream, x, y, z, dzdx, dzdy, n);
    else
        launch_k(binary_kernel<T, OP, false>, dim3(ew_grid(n, 1)), dim3(EW_THREADS), 0, stream, x, y, z, dzdx, dzdy, n);
    count_launch();
    return check_launch("binary");
}

template <class T, int MODE>
int launch_axpy(T c, const T* a, const T* b, T* out, int64_t n, cudaStream_t stream) {
    if (n <= 0) return EML_OK;
    bool vec = aligned16(out) && (!a || aligned16(a)) && (!b || aligned16(b));
    if (vec)
        launch_k(axpy_kernel<T, MODE, true>, dim3(ew_grid(n, Vec<T>::N)), dim3(EW_THREADS), 0, stream, c, a, b, out, n);
    else
        launch_k(axpy_kernel<T, MODE, false>, dim3(ew_grid(n, 1)), dim3(EW_THREADS), 0, stream, c, a, b, out, n);
    count_launch();
    return check_launch("axpy");
}

}  // namespace

int stream_tickets(cudaStream_t stream, unsigned** out) { return ticket_counters(stream, out); }

namespace {
// out[i] = w[i] - (d[i] [+ slot0 for i == 0]) * lr for up to SGD_BATCH_MAX containers in one launch (blockIdx.y picks
// the container); the same separately rounded operations as the single-container kernel (axpy_kernel MODE 1 / 4)
template <class T>
__global__ void __launch_bounds__(EW_THREADS) sgd_batch_kernel(const __grid_constant__ SgdBatch b) {
    pdl_prologue();
    const int y = blockIdx.y;
    if (y == b.count) {  // the rider (see SgdBatch)
        if (blockIdx.x == 0) {
            const unsigned char* src = static_cast<const unsigned char*>(b.copy_src);
            unsigned char* dst = static_cast<unsigned char*>(b.copy_dst);
            for (size_t i = threadIdx.x; i < b.copy_bytes; i += blockDim.x) dst[i] = src[i];
            __threadfence_system();
        }
        return;
    }
    T* out = static_cast<T*>(b.out[y]);
    const T* w = static_cast<const T*>(b.w[y]);
    const T* d = static_cast<const T*>(b.d[y]);
    const T lr = (T)b.lr[y];
    const int64_t n = b.n[y];
    float* small = sizeof(T) == 4 ? b.small[y] : nullptr;
    int unusual = 0;
    const T* ws = static_cast<const T*>(b.ws[y]);
    const int splits = b.ws_splits[y];
    const int64_t stride = b.ws_stride[y];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        T g;
        if (ws) {  // the deferred split-K reduction: 0 + partials in split order (splitk_reduce_kernel's chain)
            g = T(0);
            for (int s0 = 0; s0 < splits; s0 += 8) {
                T v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) v[u] = s0 + u < splits ? ws[(int64_t)(s0 + u) * stride + i] : T(0);
#pragma unroll
                for (int u = 0; u < 8; u++)
                    if (s0 + u < splits) g += v[u];
            }
        } else {
            g = d[i];
        }
        if (i == 0 && b.slot0[y]) {  // the parked constant-operand sums, folded in sweep order
            const T* slots = static_cast<const T*>(b.slot0[y]);
            T s = T(0);
            for (int k = 0; k < b.slot0_n[y]; k++) s = s + slots[k];
            g = g + s;
        }
        const T v = w[i] - g * lr;
        out[i] = v;
        if (small) {
            float hi, lo;
            tf32::split((float)v, hi, lo);
            small[i] = lo;
            unusual |= !tf32::ordinary((float)v);
        }
    }
    if (small) {  // the last CTA of this job writes the plane's "non-finite value seen" word
        unusual = __syncthreads_or(unusual);
        if (threadIdx.x == 0) {
            const unsigned add = unusual ? 0x10001u : 1u;
            const unsigned before = atomicAdd(b.tickets + y, add);
            if ((before & 0xffffu) + 1u == gridDim.x) {
                *b.small_flag[y] = ((before + add) >> 16) ? 1u : 0u;
                b.tickets[y] = 0u;
            }
        }
    }
}
}  // namespace

template <class T>
int launch_sgd_batch(const SgdBatch& b, cudaStream_t stream) {
    if (b.count < 1) return EML_OK;
    int64_t most = 0;
    for (int y = 0; y < b.count; y++) most = b.n[y] > most ? b.n[y] : most;
    int64_t blocks = (most + EW_THREADS - 1) / EW_THREADS;
    // few, longer-lived CTAs: every CTA of a job that writes a small-part plane ends with an atomic on the job's ticket
    // word, and same-address atomics are served one after the other (EML_SGD_BLOCKS overrides: measurement aid)
    static const int64_t cap = [] {
        const char* e = getenv("EML_SGD_BLOCKS");
        const int v = e ? atoi(e) : 0;
        return (int64_t)(v > 0 ? v : 148 * 2);
    }();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    launch_k(sgd_batch_kernel<T>, dim3((unsigned)blocks, (unsigned)(b.count + (b.copy_bytes ? 1 : 0))), dim3(EW_THREADS), 0, stream, b);
    count_launch();
    return check_launch("sgd");
}
template int launch_sgd_batch<float>(const SgdBatch&, cudaStream_t);
template int launch_sgd_batch<double>(const SgdBatch&, cudaStream_t);

template <class T>
int launch_unary(int op, T c, const T* x, T* y, T* dydx, int64_t n, cudaStream_t stream) {
    if (n <= 0) return EML_OK;
    switch (op) {
#define EML_U(OP) \
    case OP:      \
        return launch_unary_op<T, OP>(c, x, y, dydx, n, stream);
        EML_U(EML_OP_NEG) EML_U(EML_OP_SIN) EML_U(EML_OP_COS) EML_U(EML_OP_EXP) EML_U(EML_OP_LN)
        EML_U(EML_OP_SQRT) EML_U(EML_OP_ADD_C) EML_U(EML_OP_SUB_C) EML_U(EML_OP_MUL_C) EML_U(EML_OP_DIV_C)
        EML_U(EML_OP_POW_C) EML_U(EML_OP_C_SUB) EML_U(EML_OP_C_DIV) EML_U(EML_OP_C_POW)
#undef EML_U
        default:
            EML_BAD_ARG("unknown unary op %d", op);
    }
}

template <class T>
int launch_binary(int op, const T* x, const T* y, T* z, T* dzdx, T* dzdy, int64_t n, cudaStream_t stream) {
    if (n <= 0) return EML_OK;
    switch (op) {
#define EML_B(OP) \
    case OP:      \
        return launch_binary_op<T, OP>(x, y, z, dzdx, dzdy, n, stream);
        EML_B(EML_OP_ADD) EML_B(EML_OP_SUB) EML_B(EML_OP_MUL) EML_B(EML_OP_DIV) EML_B(EML_OP_POW)
#undef EML_B
        default:
            EML_BAD_ARG("unknown binary op %d", op);
    }
}

template <class T>
int launch_chain(const T* dout, const T* deriv, T* dparent, int64_t n, cudaStream_t stream) {
    return launch_axpy<T, 0>(T(0), dout, deriv, dparent, n, stream);
}
// dparent (+)= dout * deriv with deriv an array or (deriv == nullptr) the constant c; overwrite = first contribution
template <class T>
int launch_chain_ex(const T* dout, const T* deriv, T c, T* dparent, bool overwrite, int64_t n, cudaStream_t stream) {
    if (deriv)
        return overwrite ? launch_axpy<T, 5>(T(0), dout, deriv, dparent, n, stream)
                         : launch_axpy<T, 0>(T(0), dout, deriv, dparent, n, stream);
    return overwrite ? launch_axpy<T, 7>(c, dout, nullptr, dparent, n, stream)
                     : launch_axpy<T, 6>(c, dout, nullptr, dparent, n, stream);
}
template <class T>
int launch_sgd(T* w, const T* d, T lr, int64_t n, cudaStream_t stream) {
    return launch_axpy<T, 1>(lr, d, nullptr, w, n, stream);
}
template <class T>
int launch_sgd_out(T* out, const T* w, const T* d, T lr, int64_t n, cudaStream_t stream) {
    return launch_axpy<T, 4>(lr, d, w, out, n, stream);
}
template <class T>
int launch_fill(T* out, T value, int64_t n, cudaStream_t stream) {
    return launch_axpy<T, 2>(value, nullptr, nullptr, out, n, stream);
}
template <class T>
int launch_add_inplace(T* dst, const T* src, int64_t n, cudaStream_t stream) {
    return launch_axpy<T, 3>(T(0), src, nullptr, dst, n, stream);
}

namespace {
__global__ void __launch_bounds__(256) zero_ranges_kernel(const __grid_constant__ ZeroRanges z) {
    pdl_prologue();
    const int y = blockIdx.y;
    uint4* p = static_cast<uint4*>(z.p[y]);
    const size_t n = z.bytes[y] / 16;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = zero;
}
__global__ void __launch_bounds__(256) copy_to_host_kernel(const unsigned char* __restrict__ src, unsigned char* dst, size_t bytes) {
    pdl_prologue();
    for (size_t i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
}
}  // namespace
int launch_zero_ranges(const ZeroRanges& z, cudaStream_t stream) {
    if (z.count < 1) return EML_OK;
    size_t most = 0;
    for (int y = 0; y < z.count; y++) {
        if ((reinterpret_cast<uintptr_t>(z.p[y]) & 15) || (z.bytes[y] & 15)) EML_BAD_ARG("zero fill: range not 16-byte aligned");
        most = z.bytes[y] > most ? z.bytes[y] : most;
    }
    size_t blocks = (most / 16 + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    launch_k(zero_ranges_kernel, dim3((unsigned)blocks, (unsigned)z.count), dim3(256), 0, stream, z);
    count_launch();
    return check_launch("zero fill");
}
int launch_copy_to_host(const void* src, void* dst_host_devptr, size_t bytes, cudaStream_t stream) {
    if (bytes < 1) return EML_OK;
    launch_k(copy_to_host_kernel, dim3(1), dim3(256), 0, stream, static_cast<const unsigned char*>(src),
             static_cast<unsigned char*>(dst_host_devptr), bytes);
    count_launch();
    return check_launch("copy to page-locked host memory");
}

namespace {
template <class T>
__global__ void fold_slots_kernel(T* dst, const T* slots, int n, int compact) {
    pdl_prologue();
    if (blockIdx.x || threadIdx.x) return;
    T s = T(0);
    for (int k = 0; k < n; k++) s = s + slots[k];
    if (compact) {
        dst[0] = s;
    } else {
        dst[0] = dst[0] + s;
    }
}
}  // namespace
template <class T>
int launch_fold_slots(T* dst, const T* slots, int n, bool compact, cudaStream_t stream) {
    if (n < 1) return EML_OK;
    launch_k(fold_slots_kernel<T>, dim3(1), dim3(32), 0, stream, dst, slots, n, compact ? 1 : 0);
    count_launch();
    return check_launch("fold of the constant-operand sums");
}
template int launch_fold_slots<float>(float*, const float*, int, bool, cudaStream_t);
template int launch_fold_slots<double>(double*, const double*, int, bool, cudaStream_t);

template <class T>
int launch_dot(const T* x, int64_t incx, const T* y, int64_t incy, int64_t n, T* out_dev, cudaStream_t stream) {
    return reduce_impl<T>(x, incx, y, incy, n, out_dev, false, stream);
}
template <class T>
int launch_reduce_sum(DeviceCtx*, const T* x, int64_t n, T* out, bool accumulate, cudaStream_t stream) {
    return reduce_impl<T>(x, 1, nullptr, 0, n, out, accumulate, stream);
}
template <class T>
int launch_dot_accumulate(DeviceCtx*, const T* x, const T* y, int64_t n, T* out, bool accumulate,
                          cudaStream_t stream) {
    return reduce_impl<T>(x, 1, y, 1, n, out, accumulate, stream);
}

template <class T>
int launch_col_sums(const T* x, int64_t rows, int64_t cols, T* out, cudaStream_t stream) {
    int chunks = (int)((rows + 255) / 256);
    if (chunks > 128) chunks = 128;
    if (chunks < 1) chunks = 1;
    int64_t rpc = (rows + chunks - 1) / chunks;
    TempBuf partial;
    EML_TRY(partial.alloc(sizeof(T) * chunks * cols, stream));
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)chunks);
    if (grid.x > (unsigned)MAX_TICKETS) EML_BAD_ARG("col_sums: more than %d column blocks", MAX_TICKETS);
    unsigned* tickets = nullptr;
    EML_TRY(ticket_counters(stream, &tickets));
    launch_k(col_sums_kernel<T>, dim3(grid), dim3(256), 0, stream, x, rows, cols, rpc, partial.as<T>(), tickets, out);
    count_launch();
    return check_launch("col_sums");
}

template <class T>
int launch_row_sums(const T* x, int64_t rows, int64_t cols, T* out, cudaStream_t stream) {
    launch_k(row_sums_kernel<T>, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, stream, x, rows, cols, out);
    count_launch();
    return check_launch("row_sums");
}

// out[r, c] = x[r, c] - sums[f] / samples, f = c (axis 1: columns are features) or r (axis 0): the centring
// pass of the covariance (x - mean with mean = sum / S, the reference's own two roundings)
namespace {
template <class T>
__global__ void __launch_bounds__(256)
center_kernel(const T* __restrict__ x, const T* __restrict__ sums, T samples, int64_t rows, int64_t cols, int axis,
              T* __restrict__ out) {
    pdl_prologue();
    const int64_t n = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = axis == 1 ? i % cols : i / cols;
        out[i] = x[i] - sums[f] / samples;
    }
}
}  // namespace

template <class T>
int launch_center(const T* x, const T* sums, int64_t samples, int64_t rows, int64_t cols, int axis, T* out,
                  cudaStream_t stream) {
    int64_t want = (rows * cols + 255) / 256;
    int blocks = (int)(want < 148 * 16 ? want : 148 * 16);
    if (blocks < 1) blocks = 1;
    launch_k(center_kernel<T>, dim3(blocks), dim3(256), 0, stream, x, sums, (T)samples, rows, cols, axis, out);
    count_launch();
    return check_launch("center");
}

template <class T>
int launch_einsum2(const T* A, const T* B, T* out, int n_out, const int64_t* out_len, const int64_t* a_out_stride,
                   const int64_t* b_out_stride, int n_sum, const int64_t* sum_len, const int64_t* a_sum_stride,
                   const int64_t* b_sum_stride, cudaStream_t stream) {
    if (n_out < 0 || n_out > MAX_OUT || n_sum < 0 || n_sum > MAX_SUM)
        EML_BAD_ARG("einsum2: n_out=%d (max %d) n_sum=%d (max %d)", n_out, MAX_OUT, n_sum, MAX_SUM);
    EinsumParams p{};
    p.n_out = n_out;
    p.n_sum = n_sum;
    p.total = 1;
    for (int d = 0; d < n_out; d++) {
        if (out_len[d] < 0) EML_BAD_ARG("einsum2: negative length");
        p.out_len[d] = out_len[d];
        p.a_out[d] = a_out_stride[d];
        p.b_out[d] = b_out_stride[d];
        p.total *= out_len[d];
    }
    p.sum_total = 1;
    for (int d = 0; d < n_sum; d++) {
        if (sum_len[d] < 0) EML_BAD_ARG("einsum2: negative length");
        p.sum_len[d] = sum_len[d];
        p.a_sum[d] = a_sum_stride[d];
        p.b_sum[d] = b_sum_stride[d];
        p.sum_total *= sum_len[d];
    }
    if (p.total == 0) return EML_OK;
    int64_t blocks = (p.total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    launch_k(einsum2_kernel<T>, dim3((unsigned)blocks), dim3(256), 0, stream, A, B, out, p);
    count_launch();
    set_last_kernel("einsum_generic");
    return check_launch("einsum2");
}

template <class T>
int launch_einsum_n(int n_ops, const T* const* x, T* out, int n_out, const int64_t* out_len, const int64_t* out_strides,
                    int n_sum, const int64_t* sum_len, const int64_t* sum_strides, cudaStream_t stream) {
    if (n_ops < 1 || n_ops > MAX_OPS) EML_BAD_ARG("einsum: 1..%d operands (got %d)", MAX_OPS, n_ops);
    if (n_out < 0 || n_out > MAX_OUT || n_sum < 0 || n_sum > MAX_SUM_N)
        EML_BAD_ARG("einsum: n_out=%d (max %d) n_sum=%d (max %d)", n_out, MAX_OUT, n_sum, MAX_SUM_N);
    EinsumNParams p{};
    p.n_ops = n_ops;
    p.n_out = n_out;
    p.n_sum = n_sum;
    p.total = p.sum_total = 1;
    for (int i = 0; i < n_ops; i++) p.x[i] = x[i];
    for (int d = 0; d < n_out; d++) {
        if (out_len[d] < 0) EML_BAD_ARG("einsum: negative length");
        p.out_len[d] = out_len[d];
        p.total *= out_len[d];
        for (int i = 0; i < n_ops; i++) p.out_stride[i][d] = out_strides[i * n_out + d];
    }
    for (int d = 0; d < n_sum; d++) {
        if (sum_len[d] < 0) EML_BAD_ARG("einsum: negative length");
        p.sum_len[d] = sum_len[d];
        p.sum_total *= sum_len[d];
        for (int i = 0; i < n_ops; i++) p.sum_stride[i][d] = sum_strides[i * n_sum + d];
    }
    if (p.total == 0) return EML_OK;
    int64_t blocks = (p.total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    launch_k(einsum_n_kernel<T>, dim3((unsigned)blocks), dim3(256), 0, stream, out, p);
    count_launch();
    set_last_kernel("einsum_generic");
    return check_launch("einsum_n");
}

#define EML_INST(T)                                                                                              \
    template int launch_einsum_n<T>(int, const T* const*, T*, int, const int64_t*, const int64_t*, int,          \
                                    const int64_t*, const int64_t*, cudaStream_t);                               \
    template int launch_unary<T>(int, T, const T*, T*, T*, int64_t, cudaStream_t);                               \
    template int launch_binary<T>(int, const T*, const T*, T*, T*, T*, int64_t, cudaStream_t);                   \
    template int launch_chain<T>(const T*, const T*, T*, int64_t, cudaStream_t);                                 \
    template int launch_chain_ex<T>(const T*, const T*, T, T*, bool, int64_t, cudaStream_t);                     \
    template int launch_sgd<T>(T*, const T*, T, int64_t, cudaStream_t);                                          \
    template int launch_sgd_out<T>(T*, const T*, const T*, T, int64_t, cudaStream_t);                            \
    template int launch_fill<T>(T*, T, int64_t, cudaStream_t);                                                   \
    template int launch_add_inplace<T>(T*, const T*, int64_t, cudaStream_t);                                     \
    template int launch_dot<T>(const T*, int64_t, const T*, int64_t, int64_t, T*, cudaStream_t);                 \
    template int launch_reduce_sum<T>(DeviceCtx*, const T*, int64_t, T*, bool, cudaStream_t);                    \
    template int launch_dot_accumulate<T>(DeviceCtx*, const T*, const T*, int64_t, T*, bool, cudaStream_t);      \
    template int launch_col_sums<T>(const T*, int64_t, int64_t, T*, cudaStream_t);                               \
    template int launch_row_sums<T>(const T*, int64_t, int64_t, T*, cudaStream_t);                               \
    template int launch_center<T>(const T*, const T*, int64_t, int64_t, int64_t, int, T*, cudaStream_t);         \
    template int launch_einsum2<T>(const T*, const T*, T*, int, const int64_t*, const int64_t*, const int64_t*,  \
                                   int, const int64_t*, const int64_t*, const int64_t*, cudaStream_t);
EML_INST(float)
EML_INST(double)
#undef EML_INST

}  // namespace eml
