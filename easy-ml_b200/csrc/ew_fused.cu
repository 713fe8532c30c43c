// ew_fused.cu -- the fused elementwise groups of the reverse-mode tape (see ew_fused.h): one kernel evaluates a whole
// chain of container-level unary / binary rules per element, one kernel applies the chain rule of the whole chain.
//
// Compiled with -fmad=false like elementwise.cu and built from the same rule functions (ew_rules.cuh), so every value
// and every derivative carries the bits the one-kernel-per-node path produces (tests/test_gpu_fused_tape.py compares
// the two paths bit for bit; EML_FUSE=0 selects the one-kernel-per-node path).
//
// HBM-bound: a thread handles one 16-byte vector per grid-stride step; intermediate results live in registers (the
// previous op's result, the carried derivative of a chain) or in per-thread shared-memory slots (values read by a later
// op, derivatives kept from the recomputation for the backward phase, derivative accumulators of nodes with several
// consumers) -- indexed by program constants, which registers cannot be.
#include <cstdlib>

#include "common.h"
#include "ew_fused.h"
#include "ew_rules.cuh"
#include "tf32_split.cuh"

namespace eml {

namespace {

template <class T, int V>
__device__ __forceinline__ void ld(const T* p, T (&v)[V]) {
    if constexpr (V == 1) {
        v[0] = p[0];
    } else {
        vload<T>(p, v);
    }
}
template <class T, int V>
__device__ __forceinline__ void st(T* p, const T (&v)[V]) {
    if constexpr (V == 1) {
        p[0] = v[0];
    } else {
        vstore<T>(p, v);
    }
}

// One thread handles U vectors of V elements per step: the program is interpreted once for all of them (the dispatch
// overhead of an op -- descriptor loads, source selection, the switch -- is ~50 instructions, more than the arithmetic of
// most rules on one vector).  Vector u of thread t in block b starts at element ((b * U + u) * THREADS + t) * V, so every
// vector access of a warp is one contiguous 512-byte request.
template <class T, int V, int U>
struct Lanes {
    T v[U][V];
};

template <class T, int V, int U>
__device__ __forceinline__ void ld_all(const T* p, const bool (&ok)[U], Lanes<T, V, U>& x) {
#pragma unroll
    for (int u = 0; u < U; u++) {
        if (ok[u]) {
            ld<T, V>(p + u * (EW_THREADS_FUSED * V), x.v[u]);
        } else {
#pragma unroll
            for (int j = 0; j < V; j++) x.v[u][j] = T(0);
        }
    }
}
template <class T, int V, int U>
__device__ __forceinline__ void st_all(T* p, const bool (&ok)[U], const Lanes<T, V, U>& x) {
#pragma unroll
    for (int u = 0; u < U; u++)
        if (ok[u]) st<T, V>(p + u * (EW_THREADS_FUSED * V), x.v[u]);
}
// shared-memory slots need no bounds (every thread owns its U vectors of every slot)
template <class T, int V, int U>
__device__ __forceinline__ void ld_slot(const T* p, Lanes<T, V, U>& x) {
#pragma unroll
    for (int u = 0; u < U; u++) ld<T, V>(p + u * (EW_THREADS_FUSED * V), x.v[u]);
}
template <class T, int V, int U>
__device__ __forceinline__ void st_slot(T* p, const Lanes<T, V, U>& x) {
#pragma unroll
    for (int u = 0; u < U; u++) st<T, V>(p + u * (EW_THREADS_FUSED * V), x.v[u]);
}

// The interpreter is an accumulator machine: `x` (U x V registers per thread) holds the result of the previous op and is
// updated IN PLACE by the next one, so the common case -- a chain -- moves nothing between registers at the joins of the
// switch; other operands come from a leaf or a slot.  The rule bodies are the functions of ew_rules.cuh (the same
// expressions as unary_kernel / binary_kernel of elementwise.cu).  The value-only variant lets the compiler drop the
// derivative arithmetic; the backward kernel takes the variant with derivatives only for the ops whose derivative the
// sweep multiplies by.
template <class T, int V, int U, int OP, bool DERIV>
__device__ __forceinline__ void unary_case(T c, Lanes<T, V, U>& x, T* dslot) {
    Lanes<T, V, U> d;
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int j = 0; j < V; j++) {
            T yy, dd;
            unary_rule<T, OP>(c, x.v[u][j], yy, dd);
            x.v[u][j] = yy;
            d.v[u][j] = dd;
        }
    if constexpr (DERIV) st_slot<T, V, U>(dslot, d);
}
template <class T, int V, int U, int OP, bool DERIV>
__device__ __forceinline__ void binary_case(Lanes<T, V, U>& x, const Lanes<T, V, U>& b, T* daslot, T* dbslot) {
    Lanes<T, V, U> da, db;
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int j = 0; j < V; j++) {
            T zz, dx, dy;
            binary_rule<T, OP>(x.v[u][j], b.v[u][j], zz, dx, dy);
            x.v[u][j] = zz;
            da.v[u][j] = dx;
            db.v[u][j] = dy;
        }
    if constexpr (DERIV) {
        if (daslot) st_slot<T, V, U>(daslot, da);
        if (dbslot) st_slot<T, V, U>(dbslot, db);
    }
}

template <class T, int V, int U, bool DERIV>
__device__ __forceinline__ void apply_rule(int op, T c, Lanes<T, V, U>& x, const Lanes<T, V, U>& b, T* daslot, T* dbslot) {
    switch (op) {
#define EML_U(OP) \
    case (OP):    \
        unary_case<T, V, U, OP, DERIV>(c, x, daslot); \
        break;
        EML_U(EML_OP_NEG) EML_U(EML_OP_SIN) EML_U(EML_OP_COS) EML_U(EML_OP_EXP) EML_U(EML_OP_LN) EML_U(EML_OP_SQRT)
        EML_U(EML_OP_ADD_C) EML_U(EML_OP_SUB_C) EML_U(EML_OP_MUL_C) EML_U(EML_OP_DIV_C) EML_U(EML_OP_POW_C)
        EML_U(EML_OP_C_SUB) EML_U(EML_OP_C_DIV) EML_U(EML_OP_C_POW)
#undef EML_U
#define EML_B(OP)                     \
    case 14 + ((OP) - EML_OP_ADD):    \
        binary_case<T, V, U, OP, DERIV>(x, b, daslot, dbslot); \
        break;
        EML_B(EML_OP_ADD) EML_B(EML_OP_SUB) EML_B(EML_OP_MUL) EML_B(EML_OP_DIV) EML_B(EML_OP_POW)
#undef EML_B
        default:
            break;
    }
}

// Evaluates the group's ops in tape order for the U vectors starting at element i.  Forward kernel: stores the
// containers that are still held.  Backward kernel (BWD): keeps the non-constant derivatives of the rules in their slots.
template <class T, int V, int U, bool BWD>
__device__ __forceinline__ void evaluate(const EwProgram& p, int64_t i, const bool (&ok)[U], T* slots) {
    constexpr int STRIDE = EW_THREADS_FUSED * V * U;  // elements between two slots of one thread
    Lanes<T, V, U> x, b;
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int j = 0; j < V; j++) x.v[u][j] = b.v[u][j] = T(0);
    for (int k = 0; k < p.n_ops; k++) {
        const EwOp& o = p.ops[k];
        const int a_src = o.a_src, b_src = o.b_src;
        if (b_src == EW_SRC_PREV) b = x;  // (before x is replaced by operand a)
        if (a_src == EW_SRC_SLOT) ld_slot<T, V, U>(slots + o.a_idx * STRIDE, x);
        else if (a_src == EW_SRC_LEAF) ld_all<T, V, U>(static_cast<const T*>(p.in[o.a_idx]) + i, ok, x);
        if (b_src == EW_SRC_SLOT) ld_slot<T, V, U>(slots + o.b_idx * STRIDE, b);
        else if (b_src == EW_SRC_LEAF) ld_all<T, V, U>(static_cast<const T*>(p.in[o.b_idx]) + i, ok, b);
        if constexpr (BWD) {
            const int da = o.da_slot, db = o.db_slot;
            if (da >= 0 || db >= 0)
                apply_rule<T, V, U, true>(o.op, (T)o.c, x, b, da >= 0 ? slots + da * STRIDE : nullptr,
                                          db >= 0 ? slots + db * STRIDE : nullptr);
            else
                apply_rule<T, V, U, false>(o.op, (T)o.c, x, b, nullptr, nullptr);
        } else {
            apply_rule<T, V, U, false>(o.op, (T)o.c, x, b, nullptr, nullptr);
        }
        if (o.save >= 0) st_slot<T, V, U>(slots + o.save * STRIDE, x);
        if constexpr (!BWD) {
            if (o.out) st_all<T, V, U>(static_cast<T*>(p.out[k]) + i, ok, x);
        }
    }
}

template <class T, int V, int U>
__global__ void __launch_bounds__(EW_THREADS_FUSED) ew_forward_kernel(const __grid_constant__ EwProgram p) {
    pdl_prologue();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* slots = reinterpret_cast<T*>(smem_raw) + threadIdx.x * V;
    constexpr int64_t PER_BLOCK = (int64_t)EW_THREADS_FUSED * V * U;
    for (int64_t base = (int64_t)blockIdx.x * PER_BLOCK; base < p.n; base += (int64_t)gridDim.x * PER_BLOCK) {
        const int64_t i = base + threadIdx.x * V;
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; u++) ok[u] = i + u * (EW_THREADS_FUSED * V) < p.n;
        evaluate<T, V, U, false>(p, i, ok, slots);
    }
}

template <class T, int V, int U>
__device__ __forceinline__ void initial_derivative(const EwProgram& p, int k, int64_t i, const bool (&ok)[U],
                                                   Lanes<T, V, U>& g) {
    const int init = p.ops[k].g_init;
    if (init == EW_INIT_LOAD) {
        ld_all<T, V, U>(static_cast<const T*>(p.gptr[k]) + i, ok, g);
    } else {
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int j = 0; j < V; j++)
                g.v[u][j] = (init == EW_INIT_SEED && i + u * (EW_THREADS_FUSED * V) + j == p.seed_elem) ? T(1) : T(0);
    }
}

// accumulator slot += g * d  (d: the constant cd, or the derivative kept in dslot)
template <class T, int V, int U>
__device__ __forceinline__ void add_to_slot(T* acc_slot, bool is_const, T cd, const T* dslot, const Lanes<T, V, U>& g) {
    Lanes<T, V, U> d, acc;
    if (is_const) {
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int j = 0; j < V; j++) d.v[u][j] = cd;
    } else {
        ld_slot<T, V, U>(dslot, d);
    }
    ld_slot<T, V, U>(acc_slot, acc);
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int j = 0; j < V; j++) acc.v[u][j] = acc.v[u][j] + g.v[u][j] * d.v[u][j];
    st_slot<T, V, U>(acc_slot, acc);
}
// g = g * d + 0, in place: the derivative carried to the previous op of a chain (the first contribution to a
// derivative the reference holds as 0: 0 + x)
template <class T, int V, int U>
__device__ __forceinline__ void carry_on(bool is_const, T cd, const T* dslot, Lanes<T, V, U>& g) {
    if (is_const) {
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int j = 0; j < V; j++) g.v[u][j] = g.v[u][j] * cd + T(0);
    } else {
        Lanes<T, V, U> d;
        ld_slot<T, V, U>(dslot, d);
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int j = 0; j < V; j++) g.v[u][j] = g.v[u][j] * d.v[u][j] + T(0);
    }
}

// The reverse sweep (reference src/differentiation.rs:623-648) restricted to the group's entries: nodes in descending
// order; for each, derivatives[left parent] += derivatives[node] * d/dleft, then the right parent (:641-644).
template <class T, int V, int U>
__global__ void __launch_bounds__(EW_THREADS_FUSED) ew_backward_kernel(const __grid_constant__ EwProgram p) {
    pdl_prologue();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int STRIDE = EW_THREADS_FUSED * V * U;
    T* slots = reinterpret_cast<T*>(smem_raw) + threadIdx.x * V;
    constexpr int64_t PER_BLOCK = (int64_t)EW_THREADS_FUSED * V * U;
    for (int64_t base = (int64_t)blockIdx.x * PER_BLOCK; base < p.n; base += (int64_t)gridDim.x * PER_BLOCK) {
        const int64_t i = base + threadIdx.x * V;
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; u++) ok[u] = i + u * (EW_THREADS_FUSED * V) < p.n;
        evaluate<T, V, U, true>(p, i, ok, slots);
        Lanes<T, V, U> g;  // the derivative of the node being visited; carried in place along a chain
        // accumulators start from what the sweep has already put there (or from zero / the seed)
        for (int k = 0; k < p.n_ops; k++)
            if (p.ops[k].g_src == EW_G_SLOT) {
                initial_derivative<T, V, U>(p, k, i, ok, g);
                st_slot<T, V, U>(slots + p.ops[k].g_idx * STRIDE, g);
            }
        for (int l = 0; l < p.n_leafgrad; l++) {
            if (p.lg[l].overwrite) {
#pragma unroll
                for (int u = 0; u < U; u++)
#pragma unroll
                    for (int j = 0; j < V; j++) g.v[u][j] = T(0);
            } else {
                ld_all<T, V, U>(static_cast<const T*>(p.lg[l].ptr) + i, ok, g);
            }
            st_slot<T, V, U>(slots + p.lg[l].slot * STRIDE, g);
        }
        for (int k = p.n_ops - 1; k >= 0; k--) {
            const EwOp& o = p.ops[k];
            const int g_src = o.g_src, a_tgt = o.a_tgt, b_tgt = o.b_tgt;
            if (g_src == EW_G_SLOT) ld_slot<T, V, U>(slots + o.g_idx * STRIDE, g);
            else if (g_src == EW_G_DIRECT) initial_derivative<T, V, U>(p, k, i, ok, g);
            // (EW_G_CARRY: g already holds it)
            if (o.g_store) st_all<T, V, U>(static_cast<T*>(p.gptr[k]) + i, ok, g);
            // left parent, then right parent; a contribution carried in registers goes last (it replaces g, and no
            // other contribution of this node goes to the same place)
            if (a_tgt == EW_TGT_SLOT)
                add_to_slot<T, V, U>(slots + o.a_tgt_idx * STRIDE, o.a_const != 0, (T)o.ca, slots + o.da_slot * STRIDE, g);
            if (b_tgt == EW_TGT_SLOT)
                add_to_slot<T, V, U>(slots + o.b_tgt_idx * STRIDE, o.b_const != 0, (T)o.cb, slots + o.db_slot * STRIDE, g);
            if (a_tgt == EW_TGT_CARRY) carry_on<T, V, U>(o.a_const != 0, (T)o.ca, slots + o.da_slot * STRIDE, g);
            else if (b_tgt == EW_TGT_CARRY) carry_on<T, V, U>(o.b_const != 0, (T)o.cb, slots + o.db_slot * STRIDE, g);
        }
        for (int l = 0; l < p.n_leafgrad; l++) {
            ld_slot<T, V, U>(slots + p.lg[l].slot * STRIDE, g);
            st_all<T, V, U>(static_cast<T*>(p.lg[l].ptr) + i, ok, g);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Straight-line kernels for unary chains over one input (an activation function written with container operators,
// e.g. the reference examples' sigmoid 1 / (1 + e^-x) = div_swapped(1, exp(-x) + 1), src/neural_networks.rs:336):
// the op sequence is a template parameter pack, so there is nothing to interpret -- every intermediate and every kept
// derivative is a register.  Same rule functions, same operation order as the interpreter; which of the two runs is a
// performance matter only (the whitelist is chain_dispatch() below; everything else goes through the interpreter).
// ---------------------------------------------------------------------------------------------------------------
template <int OP>
constexpr bool rule_has_constant_derivative() {
    return OP == EML_OP_NEG || OP == EML_OP_ADD_C || OP == EML_OP_SUB_C || OP == EML_OP_MUL_C || OP == EML_OP_DIV_C ||
           OP == EML_OP_C_SUB;
}

template <class T, int V, int U, int K, int OP0, int... REST>
__device__ __forceinline__ void chain_forward(const EwProgram& p, Lanes<T, V, U>& x, int64_t i, const bool (&ok)[U]) {
    const T c = (T)p.ops[K].c;
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int j = 0; j < V; j++) {
            T yy, dd;
            unary_rule<T, OP0>(c, x.v[u][j], yy, dd);
            x.v[u][j] = yy;
        }
    if (p.ops[K].out) st_all<T, V, U>(static_cast<T*>(p.out[K]) + i, ok, x);
    if constexpr (sizeof...(REST) > 0) chain_forward<T, V, U, K + 1, REST...>(p, x, i, ok);
}

template <class T, int V, int U, int... OPS>
__global__ void __launch_bounds__(EW_THREADS_FUSED) chain_forward_kernel(const __grid_constant__ EwProgram p) {
    pdl_prologue();
    constexpr int64_t PER_BLOCK = (int64_t)EW_THREADS_FUSED * V * U;
    for (int64_t base = (int64_t)blockIdx.x * PER_BLOCK; base < p.n; base += (int64_t)gridDim.x * PER_BLOCK) {
        const int64_t i = base + threadIdx.x * V;
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; u++) ok[u] = i + u * (EW_THREADS_FUSED * V) < p.n;
        Lanes<T, V, U> x;
        ld_all<T, V, U>(static_cast<const T*>(p.in[0]) + i, ok, x);
        chain_forward<T, V, U, 0, OPS...>(p, x, i, ok);
    }
}

// Recomputes op K (keeping its derivative in registers), descends to the later ops, and on the way back applies the
// chain rule of op K: on return `g` holds the complete derivative of op K's INPUT (node K - 1, or the leaf).
// EW_INIT_PRODUCT: g[r][c..c+V) = sum_k G[r][k] * B[c..c+V)[k], one fma chain per element from +0, k ascending (the
// arithmetic of gemm_smallk_reg_kernel / gemm_smallk_kernel).  sBt = B transposed in shared memory ([kc][cols]: a thread
// reads its V columns of row k as one vector, a warp 32 consecutive vectors); bcol = this thread's first column, fixed for
// the whole grid-stride loop because a block's stride is a whole number of rows.  G's row is the same for the whole warp
// (a warp's vectors lie in one row): broadcast loads.
template <class T, int V, int U>
__device__ __forceinline__ void product_derivative(const EwProgram& p, int64_t i, const bool (&ok)[U], const T* sBt, int bcol,
                                                   Lanes<T, V, U>& g) {
    const T* G = static_cast<const T*>(p.prod_g);
    const int kc = p.prod_kc;
    const int64_t cols = p.prod_cols;
    const T* grow[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        grow[u] = G + (ok[u] ? (i + u * (EW_THREADS_FUSED * V)) / cols : 0) * kc;
#pragma unroll
        for (int j = 0; j < V; j++) g.v[u][j] = T(0);
    }
    for (int k = 0; k < kc; k++) {
        T b[V];
        ld<T, V>(sBt + k * cols + bcol, b);
#pragma unroll
        for (int u = 0; u < U; u++) {
            const T gv = __ldg(grow[u] + k);
#pragma unroll
            for (int j = 0; j < V; j++) g.v[u][j] = fma(gv, b[j], g.v[u][j]);
        }
    }
#pragma unroll
    for (int u = 0; u < U; u++)
        if (!ok[u]) {
#pragma unroll
            for (int j = 0; j < V; j++) g.v[u][j] = T(0);
        }
}

template <class T, int V, int U, bool PROD, int K, int N, int OP0, int... REST>
__device__ __forceinline__ void chain_backward(const EwProgram& p, Lanes<T, V, U>& x, Lanes<T, V, U>& g, int64_t i,
                                               const bool (&ok)[U], const T* sBt, int bcol) {
    const T c = (T)p.ops[K].c;
    Lanes<T, V, U> d;
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int j = 0; j < V; j++) {
            T yy, dd;
            unary_rule<T, OP0>(c, x.v[u][j], yy, dd);
            x.v[u][j] = yy;
            d.v[u][j] = dd;
        }
    if constexpr (sizeof...(REST) > 0) {
        chain_backward<T, V, U, PROD, K + 1, N, REST...>(p, x, g, i, ok, sBt, bcol);  // g: derivative of node K, complete
    } else if constexpr (PROD) {
        product_derivative<T, V, U>(p, i, ok, sBt, bcol, g);  // the last node: what its consumer, a short product, owes it
    } else {
        initial_derivative<T, V, U>(p, K, i, ok, g);  // the last node: nothing inside the chain contributes to it
    }
    if (p.ops[K].g_store) st_all<T, V, U>(static_cast<T*>(p.gptr[K]) + i, ok, g);
    // derivatives[input] = what the sweep already holds for it + g * d
    Lanes<T, V, U> base;
    if constexpr (K > 0) {
        initial_derivative<T, V, U>(p, K - 1, i, ok, base);
    } else {
        if (p.n_leafgrad > 0 && !p.lg[0].overwrite) {
            ld_all<T, V, U>(static_cast<const T*>(p.lg[0].ptr) + i, ok, base);
        } else {
#pragma unroll
            for (int u = 0; u < U; u++)
#pragma unroll
                for (int j = 0; j < V; j++) base.v[u][j] = T(0);
        }
    }
    if (p.ops[K].a_tgt == EW_TGT_NONE) {  // (only after the sweep, when node K - 1 already holds its final derivative)
        g = base;
        return;
    }
    if constexpr (rule_has_constant_derivative<OP0>()) {
        const T cd = (T)p.ops[K].ca;
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int j = 0; j < V; j++) g.v[u][j] = base.v[u][j] + g.v[u][j] * cd;
    } else {
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int j = 0; j < V; j++) g.v[u][j] = base.v[u][j] + g.v[u][j] * d.v[u][j];
    }
}

template <class T, int V, int U, bool PROD, int... OPS>
__global__ void __launch_bounds__(EW_THREADS_FUSED) chain_backward_kernel(const __grid_constant__ EwProgram p) {
    pdl_prologue();
    constexpr int64_t PER_BLOCK = (int64_t)EW_THREADS_FUSED * V * U;
    const T* sBt = nullptr;
    int bcol = 0;
    if constexpr (PROD) {
        // B [cols x kc] row-major -> shared memory [kc][cols] (read coalesced; once per persistent CTA)
        extern __shared__ __align__(16) unsigned char prod_smem[];
        T* sb = reinterpret_cast<T*>(prod_smem);
        const int kc = p.prod_kc, cols = (int)p.prod_cols;
        const T* B = static_cast<const T*>(p.prod_b);
        for (int e = threadIdx.x; e < cols * kc; e += EW_THREADS_FUSED) sb[(e % kc) * cols + e / kc] = B[e];
        __syncthreads();
        sBt = sb;
        bcol = (int)((int64_t)(threadIdx.x * V) % p.prod_cols);  // (EW_THREADS_FUSED * V) % prod_cols == 0
    }
    for (int64_t base = (int64_t)blockIdx.x * PER_BLOCK; base < p.n; base += (int64_t)gridDim.x * PER_BLOCK) {
        const int64_t i = base + threadIdx.x * V;
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; u++) ok[u] = i + u * (EW_THREADS_FUSED * V) < p.n;
        Lanes<T, V, U> x, g;
        ld_all<T, V, U>(static_cast<const T*>(p.in[0]) + i, ok, x);
        chain_backward<T, V, U, PROD, 0, (int)sizeof...(OPS), OPS...>(p, x, g, i, ok, sBt, bcol);
        if (p.n_leafgrad > 0) st_all<T, V, U>(static_cast<T*>(p.lg[0].ptr) + i, ok, g);
        if constexpr (sizeof(T) == 4) {
            if (p.n_leafgrad > 0 && p.lg_small) {  // the TF32 small parts of what was just stored (tf32_split.cuh)
#pragma unroll
                for (int u = 0; u < U; u++)
#pragma unroll
                    for (int j = 0; j < V; j++) {
                        float hi, lo;
                        tf32::split(g.v[u][j], hi, lo);
                        g.v[u][j] = lo;
                    }
                st_all<T, V, U>(p.lg_small + i, ok, g);
            }
        }
    }
}

template <class T>
size_t ew_product_smem(const EwProgram& p) { return sizeof(T) * (size_t)p.prod_cols * (size_t)p.prod_kc; }

template <class K>
int launch_chain(K kernel, const EwProgram& p, int v, int u, cudaStream_t stream, const char* what) {
    const int64_t per_block = (int64_t)EW_THREADS_FUSED * v * u;
    int64_t blocks = (p.n + per_block - 1) / per_block;
    if (blocks > 148 * 8) blocks = 148 * 8;
    launch_k(kernel, dim3((unsigned)blocks), dim3(EW_THREADS_FUSED), 0, stream, p);
    count_launch();
    return check_launch(what);
}

// a unary chain over one leaf: ops[0] reads the leaf, every later op reads its predecessor
bool is_unary_chain(const EwProgram& p) {
    if (p.n_in != 1 || p.n_ops < 1) return false;
    for (int k = 0; k < p.n_ops; k++) {
        const EwOp& o = p.ops[k];
        if (o.op >= 14 || o.a_src != (k == 0 ? EW_SRC_LEAF : EW_SRC_PREV)) return false;
    }
    return true;
}

template <class T, bool BWD, int... OPS>
int chain_go(const EwProgram& p, cudaStream_t stream) {
    constexpr int VN = Vec<T>::N;
    const char* what = BWD ? "elementwise chain backward" : "elementwise chain forward";
    if (p.n % VN != 0) {
        if constexpr (BWD) return launch_chain(chain_backward_kernel<T, 1, 1, false, OPS...>, p, 1, 1, stream, what);
        else return launch_chain(chain_forward_kernel<T, 1, 1, OPS...>, p, 1, 1, stream, what);
    }
    if constexpr (BWD) {
        if (p.ops[p.n_ops - 1].g_init == EW_INIT_PRODUCT) {
            // persistent: B is staged once per CTA
            const int64_t per_block = (int64_t)EW_THREADS_FUSED * VN * 2;
            int64_t blocks = (p.n + per_block - 1) / per_block;
            if (blocks > 148 * 4) blocks = 148 * 4;
            launch_k(chain_backward_kernel<T, VN, 2, true, OPS...>, dim3((unsigned)blocks), dim3(EW_THREADS_FUSED),
                     ew_product_smem<T>(p), stream, p);
            count_launch();
            return check_launch("elementwise chain backward (product-fed)");
        }
        return launch_chain(chain_backward_kernel<T, VN, 2, false, OPS...>, p, VN, 2, stream, what);
    } else {
        return launch_chain(chain_forward_kernel<T, VN, 2, OPS...>, p, VN, 2, stream, what);
    }
}

// The whitelist: every single rule, and the sigmoid of the reference's network examples.  *handled = false: interpret.
template <class T, bool BWD>
int chain_dispatch(const EwProgram& p, cudaStream_t stream, bool* handled) {
    *handled = false;
    static const bool off = [] {
        const char* e = getenv("EML_EW_CHAINS");
        return e && e[0] == '0';
    }();
    if (off || !is_unary_chain(p)) return EML_OK;
    *handled = true;
    if (p.n_ops == 1) {
        switch ((int)p.ops[0].op) {
#define EML_C(OP) \
    case (OP):    \
        return chain_go<T, BWD, OP>(p, stream);
            EML_C(EML_OP_NEG) EML_C(EML_OP_SIN) EML_C(EML_OP_COS) EML_C(EML_OP_EXP) EML_C(EML_OP_LN) EML_C(EML_OP_SQRT)
            EML_C(EML_OP_ADD_C) EML_C(EML_OP_SUB_C) EML_C(EML_OP_MUL_C) EML_C(EML_OP_DIV_C) EML_C(EML_OP_POW_C)
            EML_C(EML_OP_C_SUB) EML_C(EML_OP_C_DIV) EML_C(EML_OP_C_POW)
#undef EML_C
        }
    }
    if (p.n_ops == 4 && p.ops[0].op == EML_OP_NEG && p.ops[1].op == EML_OP_EXP && p.ops[2].op == EML_OP_ADD_C &&
        p.ops[3].op == EML_OP_C_DIV)
        return chain_go<T, BWD, EML_OP_NEG, EML_OP_EXP, EML_OP_ADD_C, EML_OP_C_DIV>(p, stream);
    *handled = false;
    return EML_OK;
}

bool chain_whitelisted(const EwProgram& p) {
    static const bool off = [] {
        const char* e = getenv("EML_EW_CHAINS");
        return e && e[0] == '0';
    }();
    if (off || !is_unary_chain(p)) return false;
    if (p.n_ops == 1) return true;
    return p.n_ops == 4 && p.ops[0].op == EML_OP_NEG && p.ops[1].op == EML_OP_EXP && p.ops[2].op == EML_OP_ADD_C &&
           p.ops[3].op == EML_OP_C_DIV;
}

template <class K>
int launch_program(K kernel, const EwProgram& p, int v, int u, size_t elem_bytes, cudaStream_t stream, const char* what) {
    const size_t smem = (size_t)p.n_slots * EW_THREADS_FUSED * v * u * elem_bytes;
    if (smem > (size_t(48) << 10))
        EML_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 8;  // 2048 threads
    if (smem) {
        const int fit = (int)((size_t(220) << 10) / smem);
        per_sm = fit < per_sm ? (fit < 1 ? 1 : fit) : per_sm;
    }
    const int64_t per_block = (int64_t)EW_THREADS_FUSED * v * u;
    int64_t blocks = (p.n + per_block - 1) / per_block;
    if (blocks > 148 * per_sm) blocks = 148 * per_sm;
    launch_k(kernel, dim3((unsigned)blocks), dim3(EW_THREADS_FUSED), smem, stream, p);
    count_launch();
    return check_launch(what);
}

// vectors per thread: as many as the slot budget allows once the container is large enough to fill the machine
int unroll_for(const EwProgram& p, int v, size_t elem_bytes) {
    static const int forced = [] {
        const char* e = getenv("EML_EW_UNROLL");  // 1, 2 or 4: measurement aid
        return e ? atoi(e) : 0;
    }();
    if (forced == 1 || ((forced == 2 || forced == 4) && (size_t)p.n_slots * EW_THREADS_FUSED * v * elem_bytes * forced <= (size_t(200) << 10)))
        return forced;
    const size_t slot_bytes = (size_t)EW_THREADS_FUSED * v * elem_bytes;  // one slot of one vector per thread
    for (int u = 4; u > 1; u >>= 1) {
        if ((size_t)p.n_slots * slot_bytes * u > (size_t(96) << 10)) continue;       // >= 2 CTAs per SM
        if (p.n < (int64_t)EW_THREADS_FUSED * v * u * 148 * 2) continue;               // >= 2 CTAs of work per SM
        return u;
    }
    return 1;
}

template <class T, bool BWD>
int launch_any(const EwProgram& p, cudaStream_t stream) {
    if (p.n <= 0 || p.n_ops <= 0) return EML_OK;
    const char* what = BWD ? "fused elementwise backward" : "fused elementwise forward";
    if (p.n_ops > EW_MAX_OPS || p.n_in > EW_MAX_IN || p.n_slots > EW_MAX_SLOTS)
        EML_BAD_ARG("%s: program exceeds the group limits (%d ops, %d inputs, %d slots)", what, p.n_ops, p.n_in, p.n_slots);
    {
        bool handled = false;
        EML_TRY((chain_dispatch<T, BWD>(p, stream, &handled)));
        if (handled) return EML_OK;
    }
    constexpr int VN = Vec<T>::N;
#define EML_GO(V, U)                                                                                              \
    return BWD ? launch_program(ew_backward_kernel<T, V, U>, p, V, U, sizeof(T), stream, what)                     \
               : launch_program(ew_forward_kernel<T, V, U>, p, V, U, sizeof(T), stream, what)
    // every pointer of a program comes from the stream-ordered allocator or the sweep's arena (256-byte aligned), so the
    // vector path only needs the element count to be a whole number of 16-byte vectors
    if (p.n % VN != 0) EML_GO(1, 1);
    // a container too small to give every SM a CTA of 16-byte vectors is latency-bound: one element per thread spreads
    // it over four times as many threads (NN step: the 8192 x 10 output-layer group, 3 us faster per step)
    static const bool small_scalar = [] {
        const char* e = getenv("EML_EW_SMALL_SCALAR");
        return !(e && e[0] == '0');
    }();
    if (small_scalar && p.n < (int64_t)148 * EW_THREADS_FUSED * VN) EML_GO(1, 1);
    switch (unroll_for(p, VN, sizeof(T))) {
        case 4: EML_GO(VN, 4);
        case 2: EML_GO(VN, 2);
        default: EML_GO(VN, 1);
    }
#undef EML_GO
}

}  // namespace

bool ew_backward_writes_small(const EwProgram& p) { return chain_whitelisted(p) && p.n_leafgrad == 1; }

// the straight-line chain kernels with 16-byte vectors: every vector lies in one row, and a block's stride is a whole
// number of rows (so a thread keeps its columns of B in registers)
template <class T>
bool ew_backward_takes_product(const EwProgram& p) {
    constexpr int VN = Vec<T>::N;
    return chain_whitelisted(p) && p.n % VN == 0 && p.prod_g && p.prod_b && p.prod_kc >= 1 && p.prod_kc <= EW_PROD_MAX_K &&
           p.prod_cols > 0 && p.prod_cols % VN == 0 && (EW_THREADS_FUSED * VN) % p.prod_cols == 0 &&
           ew_product_smem<T>(p) <= (size_t(32) << 10) && p.ops[p.n_ops - 1].g_src == EW_G_DIRECT;
}
template bool ew_backward_takes_product<float>(const EwProgram&);
template bool ew_backward_takes_product<double>(const EwProgram&);

template <class T>
int launch_ew_forward(const EwProgram& p, cudaStream_t stream) { return launch_any<T, false>(p, stream); }
template <class T>
int launch_ew_backward(const EwProgram& p, cudaStream_t stream) { return launch_any<T, true>(p, stream); }

template int launch_ew_forward<float>(const EwProgram&, cudaStream_t);
template int launch_ew_forward<double>(const EwProgram&, cudaStream_t);
template int launch_ew_backward<float>(const EwProgram&, cudaStream_t);
template int launch_ew_backward<double>(const EwProgram&, cudaStream_t);

}  // namespace eml
