// common.h -- shared host-side plumbing of libeasyml_b200 (errors, per-device context, launch counting).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <mutex>

#include "easyml_b200.h"

namespace eml {

// ---- errors (thread-local message behind eml_last_error) ----
void set_error(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
const char* last_error();

#define EML_CUDA(expr)                                                              \
    do {                                                                            \
        cudaError_t eml_e_ = (expr);                                                \
        if (eml_e_ != cudaSuccess) return ::eml::cuda_fail(eml_e_, #expr, __FILE__, __LINE__); \
    } while (0)

#define EML_TRY(expr)                \
    do {                             \
        int eml_rc_ = (expr);        \
        if (eml_rc_ != EML_OK) return eml_rc_; \
    } while (0)

#define EML_BAD_ARG(...)             \
    do {                             \
        ::eml::set_error(__VA_ARGS__); \
        return EML_ERR_BAD_ARG;      \
    } while (0)

// Programmatic dependent launch (PDL).  Most of an autodiff step is a chain of short dependent kernels on one stream;
// launched normally, each pays the full launch latency after its predecessor has drained.  Every kernel of this
// library starts with pdl_prologue(): it lets the NEXT kernel of the stream be scheduled right away
// (griddepcontrol.launch_dependents) and then waits until the PREVIOUS kernel has completed and flushed its memory
// (griddepcontrol.wait) before touching any global data -- so kernels still execute in order, only their launch and
// prologue overlap the predecessor.  launch_k() sets the matching launch attribute; EML_PDL=0 disables it, and it
// is off while a step is being captured into a graph.  Launched without the attribute the prologue is a no-op.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
// The two halves, for kernels with a set-up worth overlapping with the predecessor's tail (barrier initialisation, TMEM
// allocation, descriptor prefetch): nothing the predecessor may have written is read before pdl_wait().
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
bool pdl_enabled();
void pdl_suppress(bool suppress);  // thread-local (graph capture)
// EML_TRACE=1 (development aid, off by default): check_launch() records a timing event behind every kernel on the stream
// it was launched on, and eml_sync() prints the completion times of the most recent ones -- a coarse timeline that shows
// which kernels of the two streams of a sweep overlap.  (The event records get in the way of programmatic dependent
// launch, so a traced run is slower than a normal one.)
void trace_note_stream(cudaStream_t stream);
void trace_dump();

template <class... P, class... A>
inline cudaError_t launch_k(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    trace_note_stream(stream);
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

// every kernel launch of this library goes through here (bench.py reports the count)
void count_launch(int n = 1);
int64_t launches();
void set_last_kernel(const char* name);
const char* last_kernel();
// call after a launch: picks up launch-configuration errors
int check_launch(const char* what);

// ---- per-device context ----
struct DeviceCtx {
    int device = -1;
    int sm_count = 148;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;  // library stream for host entry points / NULL-stream _dev calls
    cudaStream_t copy_in = nullptr, copy_out = nullptr;  // H2D / D2H streams of the pipelined host GEMM
    cudaEvent_t events[64] = {};    // reusable (timing disabled) events of the pipelined host GEMM
    std::mutex mu;  // guards the staging buffers of the host entry points
    // staging for host entry points (device side), grow-only, one slot per operand
    void* stage[3] = {nullptr, nullptr, nullptr};
    size_t stage_bytes[3] = {0, 0, 0};
};

// context of the calling thread's current device; EML_ERR_NO_DEVICE when there is no GPU
int get_ctx(DeviceCtx** out);
// Scratch (pack buffers, split-K partials, reduction partials) and the tape's containers are stream-ordered:
// a block is only ever reused by work enqueued LATER on the SAME stream, so no synchronisation is needed.
// Blocks come from CUDA's stream-ordered allocator the first time and are then kept in per-(device, stream,
// size) free lists: a training step repeats the same sizes, so steady state makes no allocator call at all
// (an NN step went from ~120 cudaMallocAsync/cudaFreeAsync calls to none).
int stream_alloc(void** p, size_t bytes, cudaStream_t stream);
// While a step is being captured into a CUDA graph every block must come from the cache (a fresh cudaMallocAsync
// would become graph-owned memory that the cache may not recycle): a miss is then an error.
void stream_alloc_forbid_new(bool forbid);
void stream_free(void* p, size_t bytes, cudaStream_t stream);
void stream_cache_release_all();  // returns every cached block to CUDA (eml_shutdown)

struct TempBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaStream_t s = nullptr;
    TempBuf() = default;
    TempBuf(const TempBuf&) = delete;
    TempBuf& operator=(const TempBuf&) = delete;
    ~TempBuf() {
        if (p) stream_free(p, bytes, s);
    }
    int alloc(size_t bytes, cudaStream_t stream);
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};
int stage_buffer(DeviceCtx* ctx, int slot, size_t bytes, void** out);
void shutdown_all();

struct Strides3 {
    int64_t b, r, c;  // batch, row, column strides in elements
};

// A run of container-level unary operators applied to a product's output INSIDE the GEMM epilogue (f32 CTA-pair kernel):
// the tape hands over the fused elementwise group that follows a matmul (RecordTensor::unary chains, reference
// container_operations.rs:1644-1804 / swapped.rs:36-45) when it is one of the whitelisted straight-line chains -- the
// reference examples' sigmoid, 1 / (1 + e^-x), as neg, exp, + c, c / x.  C itself is always stored (the sweep recomputes
// the chain from it); out[k] != null: op k's container is held by somebody and is stored too.  Same rules, same
// -fmad=false arithmetic as the elementwise kernels: bit-identical to running the group as its own kernel.
struct EpilogueChain {
    int n_ops = 0;
    int op[4] = {0, 0, 0, 0};      // EML_OP_*
    float c[4] = {0, 0, 0, 0};     // scalar operands
    float* out[4] = {nullptr, nullptr, nullptr, nullptr};  // dense, row stride N
};

// Per-call options of the contraction dispatcher and what the kernel it chose reports back.
struct GemmOptions {
    // ---- in ----
    // Split/pack cache of the f32 tensor-core GEMM: an operand whose contents never change (a constant record
    // container: the NN's inputs) is identified by a non-zero id; its packed TF32 planes are kept and reused by later
    // calls with the same id, shape and strides.  pack_cache_drop forgets an id when its buffer dies.
    uint64_t a_id = 0, b_id = 0;
    // The result's bits must not depend on how the caller partitioned the problem (row panels of the pipelined host
    // GEMM, slabs of the sharded GEMM): no split-K reformulation.
    bool no_split_k = false;
    // Fused all-gather: up to 7 peer-mapped copies of C; the f64 TMA kernel stores every finished tile to them too.
    double* const* remote_c = nullptr;
    int n_remote = 0;
    // Set by contract_dev itself for C = G * G^T (A and B the same buffer read through mirrored strides): compute the
    // tiles on and above the diagonal only (contract_dev mirrors the upper triangle afterwards).
    bool upper_only = false;
    // With accumulate: C holds NO defined numbers yet -- the result must be what accumulating into zeros would give
    // (0 + product: the same bits, including +0 where the product is -0), without anybody having written the zeros.
    // The reverse sweep's first contribution to a derivative buffer.  Paths that end in a split-K reduction write
    // `0 + sum` there (ACC_ZERO_PLUS); the others zero C themselves first.  C must be dense (row stride N, batch 1).
    bool zero_plus = false;
    // Elementwise chain for the epilogue (see EpilogueChain); honoured only by the kernel that stores C itself
    const EpilogueChain* epilogue = nullptr;
    // The caller can take the partial sums of a deterministic split-K launch instead of C (a first contribution only:
    // zero_plus, or no accumulation): when the launch would end in the sequential reduction kernel, that kernel is NOT
    // launched, C is not written, and deferred_* describe what C is:  0 + ws[0] + ws[1] + ...  in that order, partial s
    // of element i at ws[s * stride + i].  The consumer adds them itself (the weight update of a training step:
    // sgd_batch_kernel), or launch_splitk_reduce(..., ACC_ZERO_PLUS) materialises C later -- the same bits either way.
    // The workspace becomes the caller's: stream_free(deferred_ws, deferred_bytes, <the launch's stream>).
    bool defer_reduce = false;
    // ---- out ----
    void* deferred_ws = nullptr;
    size_t deferred_bytes = 0;
    int deferred_splits = 0;
    int64_t deferred_stride = 0;
    bool epilogue_done = false;  // the chain's containers were written by the GEMM
    bool remote_done = false;  // the kernel that ran stored the tiles to the peers
    bool upper_done = false;   // the kernel that ran computed the upper triangle only
};

// how a kernel's result meets what C holds
constexpr int ACC_WRITE = 0, ACC_ADD = 1, ACC_ZERO_PLUS = 2;

// ---- kernels (one translation unit each) ----
// general strided batched contraction C[b,m,n] (+)= sum_k A[b,m,k] B[b,k,n]; dispatches between the
// tensor-core kernels and the SIMT kernel.  exact = reference operation order, bit-exact.
template <class T>
int contract_dev(DeviceCtx* ctx, const T* A, const T* B, T* C, int64_t batch, int64_t M, int64_t N, int64_t K,
                 Strides3 sA, Strides3 sB, Strides3 sC, bool accumulate, bool exact, cudaStream_t stream,
                 GemmOptions* opt = nullptr);

template <class T>
int launch_gemm_simt(const T* A, const T* B, T* C, int64_t batch, int64_t M, int64_t N, int64_t K, Strides3 sA,
                     Strides3 sB, Strides3 sC, bool accumulate, bool exact, cudaStream_t stream);

// N <= 16, A contiguous along k, M large: HBM-bound streaming kernel (deterministic shuffle-tree reduction)
template <class T>
int launch_gemm_skinny(const T* A, const T* B, T* C, int64_t batch, int64_t M, int64_t N, int64_t K, Strides3 sA,
                       Strides3 sB, Strides3 sC, bool accumulate, cudaStream_t stream);

// N <= 16, A contiguous along m (A^T stored row-major), long K: per-k-chunk partials + fixed-order reduction
template <class T>
int launch_gemm_skinny_tn(DeviceCtx* ctx, const T* A, const T* B, T* C, int64_t M, int64_t N, int64_t K, Strides3 sA,
                          Strides3 sB, Strides3 sC, int accumulate /* ACC_* */, cudaStream_t stream);

// f64 tensor-core (DMMA) kernel.  Returns EML_OK and sets *handled=false when the layout is not one it
// supports (caller falls through to another GPU kernel).
int launch_gemm_dmma_f64(DeviceCtx* ctx, const double* A, const double* B, double* C, int64_t batch, int64_t M,
                         int64_t N, int64_t K, Strides3 sA, Strides3 sB, Strides3 sC, bool accumulate,
                         cudaStream_t stream, bool* handled, GemmOptions* opt);

// f64 GEMM through error-free int8 slicing (gemm_sliced_f64.cu).  An operand is prepared once (exponent scan + slice
// planes) and can serve several products (the pipelined host GEMM slices B once for all of its row panels).
struct SlicedSide {
    TempBuf planes, ints, scales;  // [slices + 1][mn][Kp] int8; exponents (+ the `bad` word); 2^(e + 1) per row / column
    int64_t mn = 0, K = 0, Kp = 0;
    int slices = 0;
};
bool sliced_shape_ok(int64_t M, int64_t N, int64_t K, int slices);
int sliced_prepare(const double* X, bool rows_of_a, int64_t mn, int64_t K, int64_t ld, int slices, cudaStream_t stream,
                   SlicedSide* out);
int sliced_multiply(DeviceCtx* ctx, const SlicedSide& a, const SlicedSide& b, double* C, int64_t ldc, int use_slices,
                    int* path, cudaStream_t stream);
// *path = number of slices the sliced kernel used, 0 = fell back to the DMMA kernel
int gemm_f64_sliced_dev(DeviceCtx* ctx, const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K,
                        int64_t lda, int64_t ldb, int64_t ldc, int slices, int* path, cudaStream_t stream);

// measured FP64 tensor-pipe (DMMA) peak of the current device, TFLOP/s (register-resident chains)
int measure_dmma_peak(DeviceCtx* ctx, double* tflops);

// C[j][i] = C[i][j] for i < j (after a GEMM that computed the tiles on and above the diagonal only; it also makes the
// result exactly symmetric, as the reference's element-by-element covariance is)
template <class T>
int launch_mirror_upper(T* C, int64_t n, int64_t ldc, cudaStream_t stream);  // C[j][i] = C[i][j] for i < j

// Split/pack cache of the f32 tensor-core GEMM (see GemmOptions::a_id / b_id)
// (gemm_tf32x3.cu) the plane a producer kernel fills for a later product's B operand; *plane = nullptr: not possible
int tf32x3_reserve_small_plane_b(uint64_t id, const float* src, int64_t rows, int64_t cols, cudaStream_t stream, float** plane,
                                 unsigned** flag);
uint64_t new_buffer_identity();
void pack_cache_drop(uint64_t id);
void pack_cache_release_all();

// f32 3xTF32 on tcgen05 + TMEM, operands staged by TMA (pack kernel splits into big/small parts).
int launch_gemm_tf32x3_f32(DeviceCtx* ctx, const float* A, const float* B, float* C, int64_t batch, int64_t M,
                           int64_t N, int64_t K, Strides3 sA, Strides3 sB, Strides3 sC, bool accumulate,
                           cudaStream_t stream, bool* handled, GemmOptions* opt);

// C[b][m][n] (+)= sum over splits (ascending) of ws[s][b][m][n]: the fixed-order reduction behind every
// split-K path (deterministic: no atomics).  ws is dense; C is strided.
// does launch_splitk_reduce add the partials of an element one after the other (the plain kernel)?
bool splitk_reduce_is_sequential(int64_t total, int splits);
template <class T>
int launch_splitk_reduce(const T* ws, int splits, T* C, int64_t batch, int64_t M, int64_t N, Strides3 sC,
                         int accumulate /* ACC_* */, cudaStream_t stream);

// out (row-major over out_lens) [i] = in[sum_d i_d * in_strides[d]]  (reorder.cu)
template <class T>
int launch_reorder(const T* in, T* out, int rank, const int64_t* out_lens, const int64_t* in_strides,
                   cudaStream_t stream);

// in-kernel split variant of the f32 tensor-core GEMM (no pack pass); *handled = false when TMA cannot read the
// operands in place (alignment), in which case the packed path runs
int launch_gemm_tf32x3_fused(DeviceCtx* ctx, const float* A, const float* B, float* C, int64_t batch, int64_t M,
                             int64_t N, int64_t K, Strides3 sA, Strides3 sB, Strides3 sC, int splits, bool accumulate,
                             cudaStream_t stream, bool* handled, unsigned* nonfinite);

template <class T>
int launch_einsum2(const T* A, const T* B, T* out, int n_out, const int64_t* out_len, const int64_t* a_out_stride,
                   const int64_t* b_out_stride, int n_sum, const int64_t* sum_len, const int64_t* a_sum_stride,
                   const int64_t* b_sum_stride, cudaStream_t stream);

// N-operand generic einsum (1..6 operands); out_strides is [n_ops][n_out], sum_strides [n_ops][n_sum]
template <class T>
int launch_einsum_n(int n_ops, const T* const* x, T* out, int n_out, const int64_t* out_len, const int64_t* out_strides,
                    int n_sum, const int64_t* sum_len, const int64_t* sum_strides, cudaStream_t stream);

template <class T>
int launch_dot(const T* x, int64_t incx, const T* y, int64_t incy, int64_t n, T* out_dev, cudaStream_t stream);

template <class T>
int launch_unary(int op, T c, const T* x, T* y, T* dydx, int64_t n, cudaStream_t stream);
template <class T>
int launch_binary(int op, const T* x, const T* y, T* z, T* dzdx, T* dzdy, int64_t n, cudaStream_t stream);
template <class T>
int launch_chain(const T* dout, const T* deriv, T* dparent, int64_t n, cudaStream_t stream);
template <class T>
int launch_chain_ex(const T* dout, const T* deriv, T c, T* dparent, bool overwrite, int64_t n, cudaStream_t stream);
template <class T>
int launch_sgd(T* w, const T* d, T lr, int64_t n, cudaStream_t stream);
// out[i] = w[i] - lr * d[i] (the same update into a fresh container, one pass)
template <class T>
int launch_sgd_out(T* out, const T* w, const T* d, T lr, int64_t n, cudaStream_t stream);
// several weight updates in one launch (a training step updates every layer's weights back to back)
constexpr int SGD_BATCH_MAX = 8;
struct SgdBatch {
    int count = 0;
    void* out[SGD_BATCH_MAX];
    const void* w[SGD_BATCH_MAX];
    const void* d[SGD_BATCH_MAX];
    const void* slot0[SGD_BATCH_MAX];  // non-null: element 0's derivative still lacks the sums parked beside the arena
    int slot0_n[SGD_BATCH_MAX];        // ... the in-order fold of this many slots
    int64_t n[SGD_BATCH_MAX];
    double lr[SGD_BATCH_MAX];
    // f32 only: also write the small-part plane of the updated weights (tf32x3_reserve_small_plane_b): the next forward
    // product then finds it and needs no pass of its own.  The job's CTAs count themselves on ticket[y] (and whether
    // they met a non-finite value); the last one writes the plane's flag word.
    float* small[SGD_BATCH_MAX] = {};
    unsigned* small_flag[SGD_BATCH_MAX] = {};
    // non-null: d[y] has not been formed -- it is the deferred split-K reduction 0 + ws[0][i] + ws[1][i] + ... of the
    // product that computed it (GemmOptions::defer_reduce); the update adds the partials itself, in that order
    const void* ws[SGD_BATCH_MAX] = {};
    int ws_splits[SGD_BATCH_MAX] = {};
    int64_t ws_stride[SGD_BATCH_MAX] = {};
    unsigned* tickets = nullptr;
    // rider: a few bytes (the step's loss) stored straight into page-locked host memory by one extra block row of the
    // same launch -- the read-back of a training step then costs no launch of its own
    const void* copy_src = nullptr;
    void* copy_dst = nullptr;
    size_t copy_bytes = 0;
};
template <class T>
int launch_sgd_batch(const SgdBatch& b, cudaStream_t stream);
// out[i] = value (fill), used by the tape (seeding)
template <class T>
int launch_fill(T* out, T value, int64_t n, cudaStream_t stream);
// Zero fill of several ranges in ONE kernel (pointers and byte counts multiples of 16).  The sweep's vec![0; len] used
// to be cudaMemsetAsync calls: between kernels of one stream a memset breaks the programmatic-dependent-launch chain,
// and as a node of a recorded step it costs far more than a kernel node (measured with eml_graph_profile).
constexpr int ZERO_RANGES_MAX = 8;
struct ZeroRanges {
    int count = 0;
    void* p[ZERO_RANGES_MAX];
    size_t bytes[ZERO_RANGES_MAX];
};
int launch_zero_ranges(const ZeroRanges& z, cudaStream_t stream);
// dst_host[0..bytes) = src[0..bytes) by a kernel that stores straight into page-locked host memory through its device
// mapping (a few KB at most: a loss value).  A copy-engine transfer of a few bytes costs several microseconds of
// stream time and, like a memset, sits badly between kernels.
int launch_copy_to_host(const void* src, void* dst_host_devptr, size_t bytes, cudaStream_t stream);
// deterministic full reduction: out[0] (+)= sum_i x[i] (fixed tree); and the column / row sums used by the
// constant-operand path of the matmul backward.
template <class T>
int launch_reduce_sum(DeviceCtx* ctx, const T* x, int64_t n, T* out, bool accumulate, cudaStream_t stream);
template <class T>
int launch_col_sums(const T* x, int64_t rows, int64_t cols, T* out, cudaStream_t stream);  // out[c] = sum_r x[r,c]
template <class T>
int launch_row_sums(const T* x, int64_t rows, int64_t cols, T* out, cudaStream_t stream);  // out[r] = sum_c x[r,c]
// out = x - sums[feature] / samples (covariance centring); axis 1: columns are features, 0: rows
template <class T>
int launch_center(const T* x, const T* sums, int64_t samples, int64_t rows, int64_t cols, int axis, T* out,
                  cudaStream_t stream);
// out[0] (+)= sum_i x[i]*y[i], fixed order
template <class T>
int launch_dot_accumulate(DeviceCtx* ctx, const T* x, const T* y, int64_t n, T* out, bool accumulate,
                          cudaStream_t stream);
// zero-initialised, self-resetting ticket counters of a stream (single-pass deterministic reductions)
int stream_tickets(cudaStream_t stream, unsigned** out);
// two words of that array have fixed jobs: the second-level ticket of two-level reductions, and the "a split/pack pass
// of this stream met an Inf / NaN / >= 2^127 input" flag of the f32 tensor-core GEMM (set by the pack kernels and the
// in-kernel converters, read and cleared by the repair kernel that follows the product)
constexpr int TICKET_SECOND_LEVEL = 4095, TICKET_NONFINITE = 4094;
// After a 3xTF32 product C = A * B (no accumulate): when a non-finite input was seen, every NaN of C is recomputed as
// the f32 fma chain over k -- an infinite operand meets the other operand's zero small part in the cross terms
// (0 * inf) where the reference's plain products give +-inf.  One early-exit launch when nothing was seen.
int launch_tf32_repair(const float* A, const float* B, float* C, int64_t batch, int64_t M, int64_t N, int64_t K, Strides3 sA,
                       Strides3 sB, Strides3 sC, const unsigned* flag_a, const unsigned* flag_b, unsigned* stream_flag,
                       cudaStream_t stream);
// tape_kernels.cu: the reference's constant-operand sums, thin (M <= 4) products, short-contraction products
template <class T>
int launch_quirk_cols(const T* G, int64_t gm, const T* B, int64_t bm, int64_t N, T* out, cudaStream_t stream);
template <class T>
int launch_quirk_rows(const T* A, int64_t K, const T* G, int64_t N, int64_t M, T* out, cudaStream_t stream);
template <class T>
int launch_scatter_segments(const T* g, const int* perm, const int* seg_start, int n_seg, T* arena, const int64_t* seg_off,
                            cudaStream_t stream);
template <class T>
int launch_scalar_run(T* arena, int64_t self_off, int64_t n, const int64_t* lp, const int64_t* rp, const T* ld, const T* rd,
                      cudaStream_t stream);
bool thin_shape(int64_t M, int64_t N, int64_t K);
template <class T>
int launch_thin_fwd(const T* A, int64_t sAr, int64_t sAc, const T* B, int64_t M, int64_t N, int64_t K, T* C, int64_t sCr,
                    int64_t sCc, bool accumulate, cudaStream_t stream);
template <class T>
int launch_thin_bwd(const T* G, const T* A, const T* B, int64_t M, int64_t N, int64_t K, T* dA, bool over_a, T* dB,
                    bool over_b, bool quirk_a, bool quirk_b, T* slot0, cudaStream_t stream);
bool smallk_shape(int64_t M, int64_t N, int64_t K, size_t elem);
template <class T>
int launch_gemm_smallk(const T* A, int64_t lda, const T* B, int64_t sBr, int64_t sBc, T* C, int64_t ldc, int64_t M, int64_t N,
                       int64_t K, bool accumulate, cudaStream_t stream);
// dst[i] = src[idx(i)] strided gather/scatter helpers for matmul-output derivative vectors
template <class T>
int launch_add_inplace(T* dst, const T* src, int64_t n, cudaStream_t stream);
// s = ((+0 + slots[0]) + slots[1]) + ... + slots[n-1];  compact: slots[0] = s (dst == slots),
// else dst[0] = dst[0] + s.  One thread: the order of the additions is the point.
template <class T>
int launch_fold_slots(T* dst, const T* slots, int n, bool compact, cudaStream_t stream);

}  // namespace eml
