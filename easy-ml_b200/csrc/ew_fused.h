// ew_fused.h -- program format of the fused elementwise groups (host builds it in tape.cu, ew_fused.cu runs it).
//
// A run of consecutive container-level elementwise tape nodes (RecordTensor::unary / binary and their operator
// front-ends, reference src/differentiation/container_record/container_operations.rs:1644-1804, swapped.rs:25-45,
// mod.rs:1224-1247) over containers of one shape is ONE program: the forward kernel evaluates the whole chain per
// element in registers and stores only the containers somebody still holds; the backward kernel recomputes the chain
// from its inputs, and applies the chain rule of every node in descending tape order -- the same separately rounded
// multiplies and adds, in the same order, as one kernel per node would perform, so the derivatives are bit-identical.
#pragma once
#include <cstdint>

namespace eml {

constexpr int EW_MAX_OPS = 12;    // nodes per group
constexpr int EW_MAX_IN = 8;      // distinct input buffers (leaves) per group
constexpr int EW_MAX_SLOTS = 24;  // per-thread 16-byte shared-memory slots (values / derivatives / accumulators)
constexpr int EW_THREADS_FUSED = 256;

enum : uint8_t { EW_SRC_NONE = 0, EW_SRC_PREV = 1, EW_SRC_SLOT = 2, EW_SRC_LEAF = 3 };
enum : uint8_t { EW_TGT_NONE = 0, EW_TGT_SLOT = 1, EW_TGT_CARRY = 2 };
enum : uint8_t { EW_G_DIRECT = 0, EW_G_SLOT = 1, EW_G_CARRY = 2 };  // where a node's own derivative is held
enum : uint8_t { EW_INIT_ZERO = 0, EW_INIT_LOAD = 1, EW_INIT_SEED = 2, EW_INIT_PRODUCT = 3 };  // what it starts from
constexpr int EW_PROD_MAX_K = 16;  // EW_INIT_PRODUCT: longest contraction the backward kernels form on the fly

// One op of a program.  The small fields are bit-fields of two 64-bit words: the kernels index ops[k] with a loop
// variable, and a constant-bank load per field (the address-divergence pipe) was the top pipe of the first version.
struct EwOp {
    // word 0: evaluation
    uint64_t op : 8;     // dense rule number (ew_dense_op): 0..13 the unary rules, 14..18 the binary ones
    uint64_t a_src : 4;  // operand a: previous op's result / value slot / leaf
    uint64_t a_idx : 8;
    uint64_t b_src : 4;  // operand b (binary rules)
    uint64_t b_idx : 8;
    int64_t save : 8;    // value slot receiving this op's result (-1: only the next op reads it)
    uint64_t out : 1;    // forward: store the result to out[k]
    int64_t da_slot : 8; // backward: slots receiving d/da, d/db during the recomputation (-1: constant or unused)
    int64_t db_slot : 8;
    uint64_t pad0 : 7;
    // word 1: the sweep
    uint64_t a_tgt : 4;  // where g * d/da goes: nowhere / accumulator slot / the carry register
    uint64_t a_tgt_idx : 8;
    uint64_t b_tgt : 4;
    uint64_t b_tgt_idx : 8;
    uint64_t a_const : 1;  // the derivative is the number ca / cb for every element
    uint64_t b_const : 1;
    uint64_t g_src : 4;    // where this node's own derivative is held while its consumers add to it
    uint64_t g_idx : 8;
    uint64_t g_init : 4;   // ... and what it starts from: 0, the sweep's arena (gptr[k]), or the seed
    uint64_t g_store : 1;  // store the node's final derivative to gptr[k]
    uint64_t pad1 : 21;
    double c, ca, cb;      // scalar operand; constant derivatives
};
static_assert(sizeof(EwOp) == 40, "EwOp layout");

// EML_OP_* -> dense rule number (a dense switch compiles to one indexed branch instead of a compare chain)
inline int ew_dense_op(int op) { return op < 32 ? op : 14 + (op - 32); }

struct EwLeafGrad {
    void* ptr;          // derivatives of an input container that lives outside the group
    int8_t slot;        // accumulator slot
    uint8_t overwrite;  // 1: first contribution to that container in the whole sweep (starts from 0, no read)
};

struct EwProgram {
    int n_ops, n_in, n_leafgrad, n_slots;
    int64_t n;          // elements per container
    int64_t seed_elem;  // EW_G_SEED: the element whose derivative starts as 1
    const void* in[EW_MAX_IN];
    void* out[EW_MAX_OPS];   // forward outputs (per op)
    void* gptr[EW_MAX_OPS];  // backward: per op, its derivatives in the sweep's arena (EW_G_LOAD / g_store)
    EwLeafGrad lg[EW_MAX_IN];
    // f32 straight-line chains only (ew_backward_writes_small): next to the derivatives of the (single) input container
    // also store their TF32 small parts, element for element -- the plane the backward GEMM that reads these derivatives
    // as its B operand would otherwise compute in a pass of its own (tf32x3_reserve_small_plane_b)
    float* lg_small;
    // EW_INIT_PRODUCT (the group's last node only): the node's output was the left operand of a product C = H * B with a
    // short second extent (an output layer: H[8192 x 512] * W2[512 x 10]) and that product is its only consumer, so the
    // derivatives the sweep owes it are dH = G * B^T with a contraction of prod_kc <= EW_PROD_MAX_K numbers.  The backward
    // kernel forms them on the fly -- element (r, c) = fma chain over k ascending of G[r][k] * B[c][k] from +0, the
    // arithmetic of the short-contraction kernel the sweep would otherwise launch -- instead of reading them from the
    // arena: dH is neither written nor read, and the launch is saved.  prod_cols = row length of the group's containers.
    const void* prod_g;
    const void* prod_b;
    int prod_kc;
    int64_t prod_cols;
    EwOp ops[EW_MAX_OPS];
};
// true when launch_ew_backward<T>(p) runs a kernel that can form the last node's derivatives as that product
// (p.prod_* filled in, ops[n_ops - 1].g_init not yet switched to EW_INIT_PRODUCT)
template <class T>
bool ew_backward_takes_product(const EwProgram& p);
// true when launch_ew_backward<float>(p) runs a kernel that honours p.lg_small
bool ew_backward_writes_small(const EwProgram& p);

template <class T>
int launch_ew_forward(const EwProgram& p, cudaStream_t stream);
template <class T>
int launch_ew_backward(const EwProgram& p, cudaStream_t stream);

}  // namespace eml
