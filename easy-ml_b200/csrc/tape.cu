// tape.cu -- device-resident reverse-mode tape behind RecordTensor / RecordMatrix / WengertList.
//
// The reference tape is scalar-granular: a tracked M x K by K x N matmul appends M*N*(2K-1) `Operation`s
// (record_scalar_product, reference src/differentiation/container_record/container_operations.rs:1263-1315);
// at the benchmark's NN shape that is 158 GB.  Here a matmul, a container unary op and a container
// binary op are ONE macro node each.  The node remembers which range of reference tape indexes it stands
// for, so tape length, every element's Index and every derivative the reverse sweep
// (Record::try_derivatives, reference src/differentiation.rs:623-648) would produce are reproduced,
// without the entries ever existing.  Sweep of a matmul node = two GEMMs; of an elementwise node = one
// fused multiply-accumulate pass.
#include <algorithm>
#include <cstring>
#include <numeric>
#include <memory>
#include <string>
#include <vector>

#include <cstdlib>

#include "common.h"
#include "ew_fused.h"

namespace eml {

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaStream_t s = nullptr;
    std::shared_ptr<DevBuf> parent;  // set for a view into another buffer (the view owns nothing)
    uint64_t identity = 0;           // non-zero: immutable constant whose packed TF32 planes may be cached
    // non-zero: the weight update that produced these numbers also wrote their TF32 small-part plane for the NEXT forward
    // product (cached under this id, looked up only by a forward product that reads the container as its right operand)
    uint64_t plane_id = 0;
    ~DevBuf() {
        if (plane_id) pack_cache_drop(plane_id);
        if (identity) pack_cache_drop(identity);
        if (p && !parent) stream_free(p, bytes, s);
    }
};
using Buf = std::shared_ptr<DevBuf>;

Buf view_of(const Buf& arena, size_t offset, size_t bytes) {
    auto b = std::make_shared<DevBuf>();
    b->p = static_cast<char*>(arena->p) + offset;
    b->bytes = bytes;
    b->s = arena->s;
    b->parent = arena;
    return b;
}

int new_buf(size_t bytes, cudaStream_t s, Buf* out) {
    auto b = std::make_shared<DevBuf>();
    b->s = s;
    b->bytes = bytes;
    EML_TRY(stream_alloc(&b->p, bytes, s));
    *out = b;
    return EML_OK;
}

enum class Kind { VARIABLES, UNARY, BINARY, MATMUL, VIEW, SCALARS };

// A run of consecutive SCALAR tape entries -- what the scalar Record operators append one at a time
// (WengertList::append_nullary / append_unary / append_binary, reference src/differentiation.rs:811-878) -- so that
// scalar Records and containers share one list and one index space (the reference's network examples fold container
// outputs into scalar Records, src/neural_networks.rs:105-131).  Entries are kept on the host; the sweep uploads a run
// and walks it with one thread in descending order: exactly the reference's loop (:632-645), bit for bit.
struct ScalarRun {
    std::vector<int64_t> lp, rp;
    std::vector<double> ld, rd;
};

// The (T, Index) pairs of a container built by RecordTensor::from_existing / RecordMatrix::from_existing (reference
// container_record/mod.rs:219-224, :508): any index per element -- a range or mask of another container, a row of a
// constants matrix re-wrapped with a history (every Index 0: src/neural_networks.rs:105-120), duplicates allowed.
// In the sweep the container's derivatives are collected densely (like a node's) and then scattered:
// derivatives[index[e]] += g[e], all elements carrying one index summed in ascending element order by one thread.
struct IndexMap {
    std::vector<int64_t> idx;      // per element, row-major
    bool all_equal = false;        // every element carries idx[0] (the re-wrapped constants of the NN examples)
    int n_seg = 0;                 // distinct indexes
    std::vector<int64_t> seg_idx;  // the distinct indexes, ascending
    Buf perm, seg_start;           // device: elements grouped by index (int32), segment boundaries (int32, n_seg + 1)
};

// A record container: numbers on the device + how its elements map to tape indexes.
struct Value {
    Buf numbers;
    int64_t rows = 0, cols = 0;
    bool tracked = false;  // history.is_some()
    int node = -1;         // producing node in the CURRENT epoch; -1: every Index is 0 (constants)
    uint64_t epoch = 0;    // tape epoch the node id belongs to
    std::shared_ptr<IndexMap> map;  // from_existing: the indexes of the elements (node = the VIEW node when tracked)
    bool b_of_tc_product = false;   // was the right operand of a product the f32 CTA-pair kernel computes (see sgd_update)
    bool alive = false;
    int64_t n() const { return rows * cols; }
};

struct Node {
    Kind kind;
    int64_t base = 0;   // first reference tape index this node stands for
    int64_t count = 0;  // how many reference entries
    int64_t n = 0;      // elements of the produced container
    int64_t M = 0, N = 0, K = 0;
    // parents: node ids (-1 = Index 0 everywhere) and whether that side takes part in the sweep
    int p0 = -1, p1 = -1;
    bool p0_tracked = false, p1_tracked = false;
    Buf a, b;  // matmul: parent numbers A, B.  unary: a = dy/dx.  binary: a = dz/dx, b = dz/dy.
    // elementwise rules whose derivative is the same number for every element (-x, x +- c, x * c, x / c, c - x,
    // x + y, x - y): no derivative array is written in the forward pass or read in the sweep
    bool a_const = false, b_const = false;
    double a_c = 0, b_c = 0;
    // ---- member of a fused elementwise group (evaluated lazily, see Group) ----
    int group = -1;        // index into the tape's groups of this epoch; -1: evaluated eagerly, one kernel per node
    int op = -1;           // EML_OP_*
    double c = 0;          // scalar operand
    int xn = -1, yn = -1;  // operand produced by a node of the SAME group (its node id); else the numbers below
    Buf xin, yin;          // operand numbers living outside the group (kept alive: the backward pass recomputes)
    Buf value;             // this node's numbers once somebody needed them (null: interior, never stored)
    std::shared_ptr<IndexMap> map;  // Kind::VIEW: where the container's derivatives go (stands for NO tape entries)
    std::shared_ptr<ScalarRun> run;  // Kind::SCALARS: the entries (the first `n` of the run belong to this copy of the node)
    // index of element e of the produced container on the reference tape
    int64_t index_of(int64_t e) const {
        if (kind == Kind::MATMUL && K > 1) return base + e * (2 * K - 1) + 2 * K - 2;
        return base + e;
    }
    // produced element whose derivative reference entry `index` carries (interior matmul entries carry
    // the derivative of their output element: the adds propagate it unchanged to every entry of the chain)
    int64_t element_of(int64_t index) const {
        if (kind == Kind::MATMUL && K > 1) return (index - base) / (2 * K - 1);
        return index - base;
    }
};

// A fused elementwise group: a run of CONSECUTIVE tape nodes, all container-level unary / binary rules over
// containers of one element count.  While the group is open its containers have no numbers yet; the first use that
// needs numbers (a matmul operand, a read, the sweep, ...) flushes it: ONE kernel evaluates the whole run and stores
// only the containers that are still held.  In the sweep the group is again ONE kernel (ew_fused.cu).  Consecutive
// node ids make the fused backward equivalent to visiting the nodes one by one in descending order: nothing else can
// contribute to a derivative between two nodes of the run.
struct Group {
    int first = 0, last = -1;
    int64_t n = 0;
    bool flushed = false;
    std::vector<eml_val> owner;  // per node: the container handle waiting for its numbers (-1: released meanwhile)
};

}  // namespace

}  // namespace eml

using namespace eml;

struct eml_derivs {
    eml_tape* tape = nullptr;  // the tape must outlive its derivatives objects
    int dtype = EML_F32;
    int device = 0;
    cudaStream_t stream = nullptr;
    int64_t len = 0;
    uint64_t epoch = 0;
    std::vector<Node> nodes;  // copies of the metadata (buffers shared)
    std::vector<Group> groups;
    std::vector<Buf> grads;   // per node, n elements
    // per node: its derivatives are in `grads`.  Interior nodes of a fused group are not stored by the sweep (nobody
    // holds a container for them); Derivatives::at(index) / Vec::from(Derivatives) materialise them on demand.
    std::vector<char> have;
    // The reference's constant-operand sums belong to derivatives[0] (element 0 of the first node).  The sweep leaves
    // them in the arena's first slot; they are added on first use (a host read) or on the fly by the weight update,
    // which saves a launch per step.
    Buf arena;
    bool slot0_pending = false;
    int slot0_n = 0;  // the parked sum is the in-order fold of this many slots at the head of the arena (see sweep_from)
    std::vector<std::vector<int64_t>> scatter_offsets;  // host copies of uploaded scatter targets (kept until freed)
    // per node: the matmul node whose product G * B^T the sweep folded into this node's group kernel instead of storing
    // it (ProductFeed), or -1; ensure_have launches that product when somebody asks for these derivatives after all
    std::vector<int> fed_by;
    // per node: its derivatives are still the partial sums of a split-K product (GemmOptions::defer_reduce) -- a leaf's
    // weight gradient that nothing else contributes to.  The weight update adds them itself; any other reader gets them
    // through ensure_have, which launches the reduction (0 + partials in split order: the same bits).
    struct PendingReduce {
        Buf ws;
        int splits = 0;
        int64_t stride = 0;
    };
    std::vector<PendingReduce> reduce;
};

struct eml_graph {
    eml_tape* tape = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
};

struct eml_tape {
    bool capturing = false;       // between eml_tape_capture_begin and _end
    int64_t capture_len = 0;      // tape state at capture_begin: a recorded step must come back to it
    size_t capture_nodes = 0;
    int dtype = EML_F32;
    int device = 0;
    DeviceCtx* ctx = nullptr;
    cudaStream_t stream = nullptr;
    int64_t len = 0;
    uint64_t epoch = 1;
    std::vector<Node> nodes;
    std::vector<Group> groups;
    int open = -1;  // the group still accepting nodes (its containers have no numbers yet), or -1
    // weight updates requested but not launched yet: consecutive eml_tape_sgd_update calls go out as ONE kernel
    struct PendingSgd {
        Buf out, w, arena;
        const void* d;
        const void* slot0;
        int slot0_n;
        int64_t n;
        double lr;
        float* small = nullptr;          // f32: the small-part plane of the updated weights, and its flag word
        unsigned* small_flag = nullptr;
        Buf ws;                          // d is still a deferred split-K reduction (eml_derivs::PendingReduce)
        int ws_splits = 0;
        int64_t ws_stride = 0;
    };
    std::vector<PendingSgd> sgd;
    // A product whose launch is deferred to the next flush: when the operators that follow it form a fused elementwise
    // group over its output (the activation of a layer), group and product go out as ONE kernel -- the GEMM's epilogue
    // applies the chain (GemmOptions::epilogue).  The output container has its buffer already; every path that reads
    // numbers flushes first, so nobody can see it before the product has run.
    // OPT-IN (EML_EPILOGUE=1), because it is SLOWER on this hardware: measured on the NN step (8192 x 784 x 512 + sigmoid),
    // 0.285 ms per step with the activation in the epilogue against 0.194 ms with the group as its own kernel.  The
    // epilogue of the CTA-pair kernel runs on 8 warps per SM with one tile per CTA pair (nothing to overlap with), so
    // 128 exp + IEEE divisions per thread are latency-bound there, while the separate kernel runs at full occupancy
    // and HBM speed (6.9 us for the 16 MB container); the traffic saved (one re-read of the product) is 2.5 us.
    struct PendingMatmul {
        bool active = false;
        Buf a, b, c;
        uint64_t a_id = 0, b_id = 0;
        int64_t M = 0, N = 0, K = 0;
    } mm;
    bool defer_matmul = false;
    // EML_DEFER_REDUCE=1: a split-K weight gradient that only the weight update reads is left as partial sums, and the
    // update adds them itself (eml_derivs::PendingReduce): one launch fewer, the same bits.  OPT-IN: measured 2 us SLOWER
    // per NN step (0.1866 against 0.1844 ms) -- the launch that follows the CTA-pair GEMM pays the exposed transition
    // either way (the GEMM's CTAs hold the whole register file, so nothing of the next kernel is resident early), and
    // the update then carries the reduction's 14 MB of reads on top.
    bool defer_reduce = false;
    // EML_FEED_PRODUCTS=1: a short product of the sweep (dH = dY * W2^T of an output layer) is formed inside the backward
    // kernel of the group it feeds instead of being launched (ProductFeed).  Built, bit-identical, tested -- and OPT-IN,
    // because it measured slower on the NN step (0.188 ms against 0.184 ms): with the step's working set resident in the
    // 126 MB L2 the group kernel is bound by instruction issue, not by the 32 MB of traffic the fusion saves, and the
    // ten fma per element land on top of the recomputed activation.
    bool feed_products = false;
    bool fuse_packs = true;  // EML_FUSE_PACKS=0: producers do not write small-part planes for the products that follow
    bool zero_plus = true;  // EML_ZERO_PLUS=0: products accumulate into zero-filled buffers (see matmul_overwrites)
    // a small asynchronous read-back requested while updates are queued rides in their launch (SgdBatch rider)
    Buf rider_src;
    void* rider_dst = nullptr;
    size_t rider_bytes = 0;
    bool fuse = true;  // EML_FUSE=0: one kernel per elementwise node (the comparison path of the parity tests)
    // Second stream of the reverse sweep: the two products of a matmul node (dA, dB) and the constant-operand sums are
    // independent of each other, so the one nothing downstream waits for (a leaf's weight gradient, a sum parked for
    // derivatives[0]) runs here while the main stream carries on with the chain towards the inputs.  Forked and joined
    // with events inside one sweep, so a recorded step captures it as parallel graph branches.  EML_SIDE=0: one stream.
    // Two of them: [0] takes products, [1] the sums -- a sum queued behind a product would start late, and once a
    // tensor-core GEMM owns the SMs (its CTAs take the whole register file) nothing else gets in until it is done.
    cudaStream_t side[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    std::vector<Value> values;
    std::vector<eml_val> free_slots;
    std::mutex mu;
};

namespace eml {
namespace {

size_t esize(int dtype) { return dtype == EML_F32 ? 4 : 8; }

int not_while_capturing(eml_tape* t, const char* what) {
    if (t->capturing) {
        set_error("%s blocks on the device and cannot be part of a recorded step (between eml_tape_capture_begin and "
                  "_end); use eml_val_numbers_async", what);
        return EML_ERR_STATE;
    }
    return EML_OK;
}

int check_tape(eml_tape* t) {
    if (!t) EML_BAD_ARG("null tape");
    int dev = -1;
    EML_CUDA(cudaGetDevice(&dev));
    if (dev != t->device) EML_CUDA(cudaSetDevice(t->device));
    return EML_OK;
}

int get_value(eml_tape* t, eml_val v, Value** out) {
    if (v < 0 || v >= (eml_val)t->values.size() || !t->values[v].alive) EML_BAD_ARG("invalid container handle %lld", (long long)v);
    *out = &t->values[v];
    return EML_OK;
}

eml_val put_value(eml_tape* t, Value&& val) {
    val.alive = true;
    if (!t->free_slots.empty()) {
        eml_val h = t->free_slots.back();
        t->free_slots.pop_back();
        t->values[h] = std::move(val);
        return h;
    }
    t->values.push_back(std::move(val));
    return (eml_val)t->values.size() - 1;
}

// node id usable in the current epoch, or -1.  A tracked container from before the last clear() that was
// not reset is the reference's "stale Record": its indexes point at entries that no longer exist.
int live_node(eml_tape* t, const Value& v, bool* stale) {
    *stale = false;
    if (v.node < 0) return -1;
    if (v.epoch != t->epoch) {
        *stale = true;
        return -1;
    }
    return v.node;
}


// ---------------------------------------------------------------------------------------------------------------
// Fused elementwise groups: program construction (the kernels are in ew_fused.cu)
// ---------------------------------------------------------------------------------------------------------------
inline bool is_unary_op(int op) { return op < EML_OP_ADD; }

// Forward structure of a group: leaves (distinct outside buffers), operand sources, value slots.
// Returns false when the group does not fit the program format.
bool program_structure(const std::vector<Node>& nodes, const Group& g, EwProgram* p) {
    *p = EwProgram{};
    const int n_ops = g.last - g.first + 1;
    if (n_ops > EW_MAX_OPS) return false;
    p->n_ops = n_ops;
    p->n = g.n;
    p->seed_elem = -1;
    bool needs_slot[EW_MAX_OPS] = {};
    auto leaf = [&](const Buf& b) -> int {  // index of the leaf, -1: too many
        for (int l = 0; l < p->n_in; l++)
            if (p->in[l] == b->p) return l;
        if (p->n_in == EW_MAX_IN) return -1;
        p->in[p->n_in] = b->p;
        return p->n_in++;
    };
    for (int k = 0; k < n_ops; k++) {
        const Node& nd = nodes[g.first + k];
        EwOp& o = p->ops[k];
        o.op = (uint64_t)ew_dense_op(nd.op);
        o.c = nd.c;
        o.save = o.da_slot = o.db_slot = -1;
        if (nd.xn >= 0) {
            const int j = nd.xn - g.first;
            if (j == k - 1) o.a_src = EW_SRC_PREV;
            else o.a_src = EW_SRC_SLOT, o.a_idx = (uint64_t)j, needs_slot[j] = true;  // a_idx: node for now
        } else {
            const int l = leaf(nd.xin);
            if (l < 0) return false;
            o.a_src = EW_SRC_LEAF, o.a_idx = (uint64_t)l;
        }
        if (is_unary_op(nd.op)) continue;
        if (nd.yn >= 0) {
            const int j = nd.yn - g.first;
            if (j == k - 1) o.b_src = EW_SRC_PREV;
            else o.b_src = EW_SRC_SLOT, o.b_idx = (uint64_t)j, needs_slot[j] = true;
        } else {
            const int l = leaf(nd.yin);
            if (l < 0) return false;
            o.b_src = EW_SRC_LEAF, o.b_idx = (uint64_t)l;
        }
    }
    for (int k = 0; k < n_ops; k++)
        if (needs_slot[k]) p->ops[k].save = p->n_slots++;
    for (int k = 0; k < n_ops; k++) {  // node -> slot
        EwOp& o = p->ops[k];
        if (o.a_src == EW_SRC_SLOT) o.a_idx = (uint64_t)p->ops[o.a_idx].save;
        if (o.b_src == EW_SRC_SLOT) o.b_idx = (uint64_t)p->ops[o.b_idx].save;
    }
    return p->n_slots <= EW_MAX_SLOTS;
}

// which sides of node nd take part in the sweep, and whether their derivative is one number for every element
inline void sweep_sides(const Node& nd, bool* a, bool* b) {
    *a = nd.kind == Kind::UNARY ? true : nd.p0_tracked;
    *b = nd.kind == Kind::BINARY && nd.p1_tracked;
}

// worst case of the slots the backward program of this group may need, whatever the sweep looks like
bool group_fits(const std::vector<Node>& nodes, const Group& g) {
    EwProgram p;
    if (!program_structure(nodes, g, &p)) return false;
    int slots = p.n_slots + p.n_ops + p.n_in;  // values + one accumulator per node + one per leaf
    for (int i = g.first; i <= g.last; i++) {
        bool a, b;
        sweep_sides(nodes[i], &a, &b);
        if (a && !nodes[i].a_const) slots++;
        if (b && !nodes[i].b_const) slots++;
    }
    return slots <= EW_MAX_SLOTS;
}

// Backward program of group g inside one sweep.
//   written[i]: the arena already holds defined numbers for node i (zero fill or earlier contributions)
//   have[i]   : node i's FINAL derivatives are in the arena
//   remat     : after the sweep, store the derivatives of the nodes the sweep kept in registers only; nothing outside
//               the group is touched and nodes whose derivatives are final are read, not accumulated again
// The derivatives a short product owes the group's last node, formed inside the group's backward kernel instead of by a
// launch of their own (EwProgram::prod_*, EW_INIT_PRODUCT): G = derivatives of the product's output [rows x kc],
// B = its right operand [cols x kc] row-major, cols = row length of the group's containers.
struct ProductFeed {
    const void* g = nullptr;
    const void* b = nullptr;
    int kc = 0;
    int64_t cols = 0;
    bool keep = false;  // somebody still holds the last node's container: its derivatives are stored as well
};

// can the backward kernel of group g take that product?  (decided before the product's launch is skipped)
template <class T>
bool group_takes_product(const std::vector<Node>& nodes, const Group& g, const ProductFeed& f, const std::vector<char>& written,
                         int seed_node) {
    if (written[g.last] || g.last == seed_node) return false;
    EwProgram p;
    if (!program_structure(nodes, g, &p)) return false;
    p.prod_g = f.g, p.prod_b = f.b, p.prod_kc = f.kc, p.prod_cols = f.cols;
    return ew_backward_takes_product<T>(p);
}

template <class T>
int group_backward(const std::vector<Node>& nodes, const Group& g, const std::vector<Buf>& grads,
                   std::vector<char>& written, std::vector<char>& have, int seed_node, int64_t seed_elem, bool remat,
                   cudaStream_t st, float* lg_small = nullptr, bool* small_written = nullptr, const ProductFeed* feed = nullptr) {
    EwProgram p;
    if (!program_structure(nodes, g, &p)) {
        set_error("internal: a fused elementwise group does not fit its program");
        return EML_ERR_STATE;
    }
    const int n_ops = p.n_ops;
    p.seed_elem = seed_elem;
    int refs[EW_MAX_OPS] = {};           // in-group references to node k from sides that contribute
    bool next_only[EW_MAX_OPS];           // ... all of them from op k + 1
    for (int k = 0; k < n_ops; k++) next_only[k] = true;
    bool frozen[EW_MAX_OPS];
    for (int k = 0; k < n_ops; k++) frozen[k] = remat && have[g.first + k];
    auto contributes_to = [&](int k, int target_node, bool tracked) {
        if (!tracked || target_node < 0) return false;
        return !frozen[target_node - g.first];
    };
    for (int k = 0; k < n_ops; k++) {
        const Node& nd = nodes[g.first + k];
        bool a, b;
        sweep_sides(nd, &a, &b);
        if (contributes_to(k, nd.xn, a)) refs[nd.xn - g.first]++, next_only[nd.xn - g.first] &= (nd.xn - g.first == k - 1);
        if (contributes_to(k, nd.yn, b)) refs[nd.yn - g.first]++, next_only[nd.yn - g.first] &= (nd.yn - g.first == k - 1);
    }
    // where each node's own derivative starts and lives
    for (int k = 0; k < n_ops; k++) {
        const int id = g.first + k;
        EwOp& o = p.ops[k];
        p.gptr[k] = grads[id]->p;
        if (frozen[k]) o.g_init = EW_INIT_LOAD;
        else if (!remat && id == seed_node && !written[id]) o.g_init = EW_INIT_SEED;
        else o.g_init = written[id] ? EW_INIT_LOAD : EW_INIT_ZERO;
        if (refs[k] == 0) o.g_src = EW_G_DIRECT;
        else if (refs[k] == 1 && next_only[k] && o.g_init == EW_INIT_ZERO) o.g_src = EW_G_CARRY;
        else o.g_src = EW_G_SLOT, o.g_idx = (uint64_t)p.n_slots++;
        const bool changed = refs[k] > 0 || o.g_init != EW_INIT_LOAD;
        if (remat) o.g_store = !have[id];
        else o.g_store = nodes[id].value && changed;
    }
    // contributions of each side
    auto leaf_grad = [&](int parent) -> int {
        for (int l = 0; l < p.n_leafgrad; l++)
            if (p.lg[l].ptr == grads[parent]->p) return p.lg[l].slot;
        EwLeafGrad& lg = p.lg[p.n_leafgrad++];
        lg.ptr = grads[parent]->p;
        lg.slot = (int8_t)p.n_slots++;
        lg.overwrite = !written[parent];
        written[parent] = 1;
        return lg.slot;
    };
    for (int k = 0; k < n_ops; k++) {
        const Node& nd = nodes[g.first + k];
        EwOp& o = p.ops[k];
        bool a, b;
        sweep_sides(nd, &a, &b);
        o.a_const = nd.a_const, o.ca = nd.a_c;
        o.b_const = nd.b_const, o.cb = nd.b_c;
        for (int side = 0; side < 2; side++) {
            const bool tracked = side ? b : a;
            const int in_group = side ? nd.yn : nd.xn, parent = side ? nd.p1 : nd.p0;
            uint8_t tgt = EW_TGT_NONE, idx = 0;
            if (tracked && in_group >= 0) {
                const int j = in_group - g.first;
                if (!frozen[j]) {
                    if (p.ops[j].g_src == EW_G_CARRY) tgt = EW_TGT_CARRY;
                    else tgt = EW_TGT_SLOT, idx = p.ops[j].g_idx;
                }
            } else if (tracked && parent >= 0 && !remat) {
                bool known = false;
                for (int l = 0; l < p.n_leafgrad; l++) known |= p.lg[l].ptr == grads[parent]->p;
                if (!known && p.n_leafgrad == EW_MAX_IN) {
                    set_error("internal: too many tracked inputs in a fused elementwise group");
                    return EML_ERR_STATE;
                }
                tgt = EW_TGT_SLOT, idx = (uint8_t)leaf_grad(parent);
            }
            const bool is_const = side ? nd.b_const : nd.a_const;
            int dslot = -1;
            if (tgt != EW_TGT_NONE && !is_const) dslot = p.n_slots++;
            if (side) o.b_tgt = tgt, o.b_tgt_idx = idx, o.db_slot = dslot;
            else o.a_tgt = tgt, o.a_tgt_idx = idx, o.da_slot = dslot;
        }
    }
    if (p.n_slots > EW_MAX_SLOTS) {
        set_error("internal: a fused elementwise group needs %d slots", p.n_slots);
        return EML_ERR_STATE;
    }
    if (feed) {  // (group_takes_product said yes: the last node starts from nothing and a chain kernel runs)
        p.prod_g = feed->g, p.prod_b = feed->b, p.prod_kc = feed->kc, p.prod_cols = feed->cols;
        EwOp& last = p.ops[n_ops - 1];
        if (remat || last.g_init != EW_INIT_ZERO || !ew_backward_takes_product<T>(p)) {
            set_error("internal: a fused elementwise group cannot take the product it was promised");
            return EML_ERR_STATE;
        }
        last.g_init = EW_INIT_PRODUCT;
        last.g_store = feed->keep ? 1 : 0;
    }
    bool anything = p.n_leafgrad > 0;
    for (int k = 0; k < n_ops; k++) anything |= p.ops[k].g_store != 0;
    if (lg_small && sizeof(T) == 4 && ew_backward_writes_small(p)) {
        p.lg_small = lg_small;
        if (small_written) *small_written = true;
    }
    if (anything) EML_TRY(launch_ew_backward<T>(p, st));
    for (int k = 0; k < n_ops; k++) {
        const int id = g.first + k;
        if (p.ops[k].g_store) have[id] = 1, written[id] = 1;
        else if (!remat) have[id] = nodes[id].value && written[id];  // stored containers whose arena entry is final
    }
    return EML_OK;
}

// Gives every still-held container of the open group its numbers: one kernel for the whole run.
template <class T>
int flush_sgd(eml_tape* t) {
    for (size_t i = 0; i < t->sgd.size(); i += SGD_BATCH_MAX) {
        SgdBatch b;
        for (size_t j = i; j < t->sgd.size() && j < i + SGD_BATCH_MAX; j++) {
            const auto& u = t->sgd[j];
            const int y = b.count++;
            b.out[y] = u.out->p, b.w[y] = u.w->p, b.d[y] = u.d, b.slot0[y] = u.slot0, b.slot0_n[y] = u.slot0_n, b.n[y] = u.n, b.lr[y] = u.lr;
            b.ws[y] = u.ws ? u.ws->p : nullptr, b.ws_splits[y] = u.ws_splits, b.ws_stride[y] = u.ws_stride;
            b.small[y] = u.small, b.small_flag[y] = u.small_flag;
        }
        EML_TRY(stream_tickets(t->stream, &b.tickets));
        if (t->rider_src && i + SGD_BATCH_MAX >= t->sgd.size()) {  // the last launch carries the rider
            b.copy_src = t->rider_src->p, b.copy_dst = t->rider_dst, b.copy_bytes = t->rider_bytes;
            t->rider_src.reset();
        }
        EML_TRY(launch_sgd_batch<T>(b, t->stream));
    }
    t->sgd.clear();
    return EML_OK;
}

template <class T>
int flush_groups(eml_tape* t, bool keep_pending = false);

template <class T>
int flush_open(eml_tape* t, bool keep_pending = false) {
    // (updates requested earlier come first: a group opened after them reads the updated weights)
    if (!t->sgd.empty()) EML_TRY(flush_sgd<T>(t));
    return flush_groups<T>(t, keep_pending);
}

// the deferred product, with or without an elementwise chain for its epilogue (*fused: the chain went with it)
template <class T>
int run_pending_matmul(eml_tape* t, const EpilogueChain* epi, bool* fused) {
    eml_tape::PendingMatmul mm = std::move(t->mm);
    t->mm = eml_tape::PendingMatmul();
    if (fused) *fused = false;
    if (!mm.active) return EML_OK;
    GemmOptions ids;  // constant operands keep their packed planes across steps
    ids.a_id = mm.a_id;
    ids.b_id = mm.b_id;
    ids.epilogue = epi;
    EML_TRY(contract_dev<T>(t->ctx, (const T*)mm.a->p, (const T*)mm.b->p, (T*)mm.c->p, 1, mm.M, mm.N, mm.K,
                            Strides3{0, mm.K, 1}, Strides3{0, mm.N, 1}, Strides3{0, mm.N, 1}, false, false, t->stream, &ids));
    if (fused) *fused = ids.epilogue_done;
    return EML_OK;
}

// keep_pending: called on the way to opening a new group -- a deferred product with no group open yet stays deferred
// (the new group may be the chain over its output)
template <class T>
int flush_groups(eml_tape* t, bool keep_pending) {
    if (t->open < 0) {
        if (keep_pending) return EML_OK;
        return run_pending_matmul<T>(t, nullptr, nullptr);
    }
    Group& g = t->groups[t->open];
    t->open = -1;
    g.flushed = true;
    EwProgram p;
    if (!program_structure(t->nodes, g, &p)) {
        set_error("internal: a fused elementwise group does not fit its program");
        return EML_ERR_STATE;
    }
    bool any = false;
    for (int k = 0; k < p.n_ops; k++) {
        const eml_val h = g.owner[k];
        if (h < 0) continue;  // released before anybody needed its numbers: never stored
        Buf b;
        EML_TRY(new_buf(sizeof(T) * (size_t)g.n, t->stream, &b));
        t->values[h].numbers = b;
        t->nodes[g.first + k].value = b;
        p.ops[k].out = 1;
        p.out[k] = b->p;
        any = true;
    }
    bool fused = false;
    if (t->mm.active) {
        // the group is a chain of unary rules over the deferred product's output: offer it to the GEMM's epilogue
        EpilogueChain epi;
        bool offer = any && sizeof(T) == 4 && p.n_ops <= 4 && p.n_in == 1 && p.in[0] == t->mm.c->p && g.n == t->mm.M * t->mm.N;
        for (int k = 0; k < p.n_ops && offer; k++) {
            const Node& nd = t->nodes[g.first + k];
            offer = is_unary_op(nd.op) && (k == 0 ? nd.xn < 0 : nd.xn == g.first + k - 1);
            epi.op[k] = nd.op;
            epi.c[k] = (float)nd.c;
            epi.out[k] = p.ops[k].out ? (float*)p.out[k] : nullptr;
        }
        epi.n_ops = p.n_ops;
        EML_TRY(run_pending_matmul<T>(t, offer ? &epi : nullptr, &fused));
    }
    if (any && !fused) EML_TRY(launch_ew_forward<T>(p, t->stream));
    return EML_OK;
}

int flush(eml_tape* t) { return t->dtype == EML_F32 ? flush_open<float>(t) : flush_open<double>(t); }

// Appends one lazily evaluated elementwise node (already filled in except for the group bookkeeping) and the container
// that will receive its numbers.  x / y: the operand containers (y null for unary rules).
template <class T>
int append_lazy(eml_tape* t, Node nd, Value* x, Value* y, Value&& result, eml_val* out) {
    const int64_t n = nd.n;
    auto operand_in_open_group = [&](Value* v) { return v && !v->numbers; };
    // try to extend the open group
    bool extended = false;
    if (t->open >= 0) {
        Group& g = t->groups[t->open];
        if (g.last == (int)t->nodes.size() - 1 && g.n == n) {
            nd.group = t->open;
            nd.xn = operand_in_open_group(x) ? x->node : -1;
            nd.yn = operand_in_open_group(y) ? y->node : -1;
            nd.xin = nd.xn < 0 ? x->numbers : nullptr;
            nd.yin = (y && nd.yn < 0) ? y->numbers : nullptr;
            t->nodes.push_back(nd);
            g.last++;
            if (group_fits(t->nodes, g)) {
                extended = true;
            } else {
                g.last--;
                t->nodes.pop_back();
            }
        }
    }
    if (!extended) {
        EML_TRY(flush_open<T>(t, true));  // the operands (held by the caller) now have numbers
        Group g;
        g.first = g.last = (int)t->nodes.size();
        g.n = n;
        t->groups.push_back(g);
        t->open = (int)t->groups.size() - 1;
        nd.group = t->open;
        nd.xn = nd.yn = -1;
        nd.xin = x->numbers;
        nd.yin = y ? y->numbers : nullptr;
        t->nodes.push_back(nd);
    }
    t->len += nd.count;
    result.tracked = true;
    result.node = (int)t->nodes.size() - 1;
    result.epoch = t->epoch;
    const eml_val h = put_value(t, std::move(result));
    t->groups[t->open].owner.push_back(h);
    *out = h;
    return EML_OK;
}

template <class T>
int new_container(eml_tape* t, const void* src, bool src_on_device, int64_t rows, int64_t cols, bool variables,
                  eml_val* out) {
    if (rows < 1 || cols < 1) EML_BAD_ARG("container shape %lldx%lld: Easy ML containers have at least one element",
                                          (long long)rows, (long long)cols);
    if (!src || !out) EML_BAD_ARG("null argument");
    if (!src_on_device) EML_TRY(not_while_capturing(t, "creating a container from host memory"));
    EML_TRY(flush_open<T>(t));
    Value v;
    v.rows = rows;
    v.cols = cols;
    EML_TRY(new_buf(sizeof(T) * rows * cols, t->stream, &v.numbers));
    EML_CUDA(cudaMemcpyAsync(v.numbers->p, src, sizeof(T) * rows * cols,
                             src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, t->stream));
    if (!src_on_device) EML_CUDA(cudaStreamSynchronize(t->stream));  // the host buffer is the caller's
    // numbers of a container are never modified in place (operators and the SGD update produce fresh containers);
    // constants typically live across many steps, so their split/packed planes are worth keeping
    if (!variables) v.numbers->identity = new_buffer_identity();
    if (variables) {
        Node nd;
        nd.kind = Kind::VARIABLES;
        nd.base = t->len;  // append_nullary_repeating: starting_index = operations.len()
        nd.count = nd.n = rows * cols;
        t->len += nd.count;
        t->nodes.push_back(nd);
        v.tracked = true;
        v.node = (int)t->nodes.size() - 1;
        v.epoch = t->epoch;
    }
    *out = put_value(t, std::move(v));
    return EML_OK;
}

int node_of_index(const std::vector<Node>& nodes, int64_t index);

// Classifies the indexes of a re-wrapped container and, when it takes part in sweeps, prepares the grouping of its
// elements by index on the device.
int build_index_map(eml_tape* t, std::vector<int64_t>&& idx, bool tracked, std::shared_ptr<IndexMap>* out) {
    auto map = std::make_shared<IndexMap>();
    map->idx = std::move(idx);
    const int64_t n = (int64_t)map->idx.size();
    map->all_equal = true;
    for (int64_t e = 1; e < n && map->all_equal; e++) map->all_equal = map->idx[e] == map->idx[0];
    if (tracked) {
        if (n > 0x7fffffffLL) EML_BAD_ARG("from_existing: too many elements");
        std::vector<int> perm((size_t)n);
        std::iota(perm.begin(), perm.end(), 0);
        if (!map->all_equal)
            std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return map->idx[a] < map->idx[b]; });
        std::vector<int> seg_start;
        for (int64_t j = 0; j < n; j++)
            if (j == 0 || map->idx[perm[j]] != map->idx[perm[j - 1]]) {
                seg_start.push_back((int)j);
                map->seg_idx.push_back(map->idx[perm[j]]);
            }
        seg_start.push_back((int)n);
        map->n_seg = (int)map->seg_idx.size();
        // every index must be the Index of a container element that exists on the list right now (the reference does
        // not check and panics with an out-of-bounds index when derivatives are taken)
        for (int64_t index : map->seg_idx) {
            const int nd = node_of_index(t->nodes, index);
            if (nd < 0) EML_BAD_ARG("from_existing: index %lld is not on the WengertList (its length is %lld)", (long long)index, (long long)t->len);
            if (t->nodes[nd].index_of(t->nodes[nd].element_of(index)) != index)
                EML_BAD_ARG("from_existing: index %lld is an entry inside a matrix product, not the Index of a container element", (long long)index);
        }
        if (!map->all_equal) {
            EML_TRY(not_while_capturing(t, "re-wrapping a container with arbitrary indexes"));
            EML_TRY(new_buf(sizeof(int) * (size_t)n, t->stream, &map->perm));
            EML_TRY(new_buf(sizeof(int) * seg_start.size(), t->stream, &map->seg_start));
            EML_CUDA(cudaMemcpyAsync(map->perm->p, perm.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, t->stream));
            EML_CUDA(cudaMemcpyAsync(map->seg_start->p, seg_start.data(), sizeof(int) * seg_start.size(), cudaMemcpyHostToDevice, t->stream));
            EML_CUDA(cudaStreamSynchronize(t->stream));  // the host vectors die with this call
        }
    }
    *out = map;
    return EML_OK;
}

// the container of a from_existing call: numbers already on the device
int put_view(eml_tape* t, Buf numbers, int64_t rows, int64_t cols, std::shared_ptr<IndexMap> map, bool tracked, eml_val* out) {
    Value v;
    v.rows = rows;
    v.cols = cols;
    v.numbers = numbers;
    v.map = map;
    v.tracked = tracked;
    if (tracked) {
        Node nd;
        nd.kind = Kind::VIEW;  // no tape entries: from_existing appends nothing to the list
        nd.base = t->len;
        nd.count = 0;
        nd.n = rows * cols;
        nd.map = map;
        t->nodes.push_back(nd);
        v.node = (int)t->nodes.size() - 1;
        v.epoch = t->epoch;
    } else {
        v.numbers->identity = new_buffer_identity();  // history None: constants
    }
    *out = put_value(t, std::move(v));
    return EML_OK;
}

template <class T>
int tape_matmul(eml_tape* t, eml_val ha, eml_val hb, eml_val* out) {
    Value *a, *b;
    EML_TRY(get_value(t, ha, &a));
    EML_TRY(get_value(t, hb, &b));
    if (a->cols != b->rows)
        EML_BAD_ARG("Mismatched Matrices, left is %lldx%lld, right is %lldx%lld, * is only defined for MxN * NxL",
                    (long long)a->rows, (long long)a->cols, (long long)b->rows, (long long)b->cols);
    bool sa, sb;
    int na = live_node(t, *a, &sa), nb = live_node(t, *b, &sb);
    if (sa || sb) {
        set_error("container was not reset after WengertList::clear(); its indexes are stale");
        return EML_ERR_STATE;
    }
    const int64_t M = a->rows, K = a->cols, N = b->cols;
    EML_TRY(flush_open<T>(t));  // the operands may be containers of the open elementwise group
    Value c;
    c.rows = M;
    c.cols = N;
    EML_TRY(new_buf(sizeof(T) * M * N, t->stream, &c.numbers));
    t->mm.active = true;
    t->mm.a = a->numbers;
    t->mm.b = b->numbers;
    t->mm.c = c.numbers;
    t->mm.a_id = a->numbers->identity;  // constant operands keep their packed planes across steps
    t->mm.b_id = b->numbers->identity ? b->numbers->identity : b->numbers->plane_id;
    b->b_of_tc_product = sizeof(T) == 4 && M > 128 && N > 128 && K >= 32 && M * N * K >= (int64_t(1) << 24);
    t->mm.M = M, t->mm.N = N, t->mm.K = K;
    // deferred only where the epilogue could take the chain (f32, tracked, a shape of the CTA-pair kernel) and the
    // tape fuses at all; everything else goes out now
    const bool defer = t->defer_matmul && t->fuse && sizeof(T) == 4 && (a->tracked || b->tracked) && M > 128 && N > 128 && K >= 32;
    if (!defer) EML_TRY(run_pending_matmul<T>(t, nullptr, nullptr));
    // history = lhs's else rhs's (reference container_operations.rs:1357-1361); with a history every
    // product is recorded with BOTH operands as parents (:1291-1296), whatever their own history
    c.tracked = a->tracked || b->tracked;
    if (c.tracked) {
        Node nd;
        nd.kind = Kind::MATMUL;
        nd.base = t->len;
        nd.count = M * N * (2 * K - 1);
        nd.n = M * N;
        nd.M = M;
        nd.N = N;
        nd.K = K;
        nd.p0 = na;
        nd.p1 = nb;
        nd.p0_tracked = nd.p1_tracked = true;
        nd.a = a->numbers;
        nd.b = b->numbers;
        t->len += nd.count;
        t->nodes.push_back(nd);
        c.node = (int)t->nodes.size() - 1;
        c.epoch = t->epoch;
    }
    *out = put_value(t, std::move(c));
    return EML_OK;
}

template <class T>
int tape_unary(eml_tape* t, int op, double cst, eml_val hx, eml_val* out) {
    Value* x;
    EML_TRY(get_value(t, hx, &x));
    if (op < EML_OP_NEG || op > EML_OP_C_POW) EML_BAD_ARG("unknown unary op %d", op);
    bool stale;
    int nx = live_node(t, *x, &stale);
    if (stale) {
        set_error("container was not reset after WengertList::clear(); its indexes are stale");
        return EML_ERR_STATE;
    }
    Value y;
    y.rows = x->rows;
    y.cols = x->cols;
    const int64_t n = x->n();
    if (!x->tracked) {  // history None -> constants (reference container_record/mod.rs:852)
        EML_TRY(new_buf(sizeof(T) * n, t->stream, &y.numbers));
        EML_TRY(launch_unary<T>(op, (T)cst, (const T*)x->numbers->p, (T*)y.numbers->p, (T*)nullptr, n, t->stream));
        *out = put_value(t, std::move(y));
        return EML_OK;
    }
    Node nd;
    nd.kind = Kind::UNARY;
    nd.base = t->len;
    nd.count = nd.n = n;
    nd.p0 = nx;
    nd.p0_tracked = true;
    nd.op = op;
    nd.c = cst;
    // the same expressions as unary_rule (csrc/ew_rules.cuh), evaluated in T
    const T c = (T)cst;
    switch (nx >= 0 ? op : -1) {  // (a parent carrying Index 0 keeps the array: its sweep is a dot product)
        case EML_OP_NEG: case EML_OP_C_SUB: nd.a_const = true; nd.a_c = (double)(-T(1)); break;
        case EML_OP_ADD_C: case EML_OP_SUB_C: nd.a_const = true; nd.a_c = (double)T(1); break;
        case EML_OP_MUL_C: nd.a_const = true; nd.a_c = (double)c; break;
        case EML_OP_DIV_C: nd.a_const = true; nd.a_c = (double)(T(1) / c); break;
        default: break;
    }
    if (t->fuse && nx >= 0) return append_lazy<T>(t, nd, x, nullptr, std::move(y), out);
    // one kernel per node: value and derivative array now
    EML_TRY(flush_open<T>(t));
    EML_TRY(new_buf(sizeof(T) * n, t->stream, &y.numbers));
    if (!nd.a_const) EML_TRY(new_buf(sizeof(T) * n, t->stream, &nd.a));
    EML_TRY(launch_unary<T>(op, c, (const T*)x->numbers->p, (T*)y.numbers->p, nd.a ? (T*)nd.a->p : nullptr, n, t->stream));
    t->len += n;
    t->nodes.push_back(nd);
    y.tracked = true;
    y.node = (int)t->nodes.size() - 1;
    y.epoch = t->epoch;
    *out = put_value(t, std::move(y));
    return EML_OK;
}

template <class T>
int tape_binary(eml_tape* t, int op, eml_val hx, eml_val hy, eml_val* out) {
    Value *x, *y;
    EML_TRY(get_value(t, hx, &x));
    EML_TRY(get_value(t, hy, &y));
    if (op < EML_OP_ADD || op > EML_OP_POW) EML_BAD_ARG("unknown binary op %d", op);
    if (x->rows != y->rows || x->cols != y->cols)
        EML_BAD_ARG("Record containers must have the same shape for a binary operation: (left: %lldx%lld, right: %lldx%lld)",
                    (long long)x->rows, (long long)x->cols, (long long)y->rows, (long long)y->cols);
    bool sx, sy;
    int nx = live_node(t, *x, &sx), ny = live_node(t, *y, &sy);
    if (sx || sy) {
        set_error("container was not reset after WengertList::clear(); its indexes are stale");
        return EML_ERR_STATE;
    }
    Value z;
    z.rows = x->rows;
    z.cols = x->cols;
    const int64_t n = x->n();
    if (!x->tracked && !y->tracked) {  // (None, None) reference mod.rs:984
        EML_TRY(new_buf(sizeof(T) * n, t->stream, &z.numbers));
        EML_TRY(launch_binary<T>(op, (const T*)x->numbers->p, (const T*)y->numbers->p, (T*)z.numbers->p, (T*)nullptr,
                                 (T*)nullptr, n, t->stream));
        *out = put_value(t, std::move(z));
        return EML_OK;
    }
    Node nd;
    nd.kind = Kind::BINARY;
    nd.base = t->len;
    nd.count = nd.n = n;
    nd.p0 = nx;
    nd.p1 = ny;
    nd.p0_tracked = x->tracked;  // binary_x_history / binary_y_history / binary_both_history
    nd.p1_tracked = y->tracked;
    nd.op = op;
    if ((op == EML_OP_ADD || op == EML_OP_SUB) && nd.p0_tracked && nx >= 0) nd.a_const = true, nd.a_c = 1.0;
    if ((op == EML_OP_ADD || op == EML_OP_SUB) && nd.p1_tracked && ny >= 0)
        nd.b_const = true, nd.b_c = op == EML_OP_ADD ? 1.0 : -1.0;
    if (t->fuse && (!nd.p0_tracked || nx >= 0) && (!nd.p1_tracked || ny >= 0))
        return append_lazy<T>(t, nd, x, y, std::move(z), out);
    EML_TRY(flush_open<T>(t));
    EML_TRY(new_buf(sizeof(T) * n, t->stream, &z.numbers));
    if (nd.p0_tracked && !nd.a_const) EML_TRY(new_buf(sizeof(T) * n, t->stream, &nd.a));
    if (nd.p1_tracked && !nd.b_const) EML_TRY(new_buf(sizeof(T) * n, t->stream, &nd.b));
    EML_TRY(launch_binary<T>(op, (const T*)x->numbers->p, (const T*)y->numbers->p, (T*)z.numbers->p,
                             nd.a ? (T*)nd.a->p : nullptr, nd.b ? (T*)nd.b->p : nullptr, n, t->stream));
    t->len += n;
    t->nodes.push_back(nd);
    z.tracked = true;
    z.node = (int)t->nodes.size() - 1;
    z.epoch = t->epoch;
    *out = put_value(t, std::move(z));
    return EML_OK;
}

// node whose entries contain reference tape index `index` (-1: out of range)
int node_of_index(const std::vector<Node>& nodes, int64_t index) {
    if (nodes.empty() || index < 0) return -1;
    int lo = 0, hi = (int)nodes.size() - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) / 2;
        if (nodes[mid].base <= index) lo = mid;
        else hi = mid - 1;
    }
    while (lo > 0 && nodes[lo].count == 0) lo--;  // a view stands for no entries
    if (index < nodes[lo].base || index >= nodes[lo].base + nodes[lo].count) return -1;
    return lo;
}

// The sweep of a re-wrapped container: its densely collected derivatives g are added to derivatives[index[e]].
template <class T>
int scatter_view(eml_tape* t, eml_derivs* d, const Node& nd, const T* g, const Buf& arena, const std::vector<size_t>& offs,
                 cudaStream_t st) {
    const IndexMap& map = *nd.map;
    auto target = [&](int64_t index, int64_t* elem_off) -> int {  // offset (elements of T) inside the arena
        const int n = node_of_index(d->nodes, index);
        if (n < 0) {
            set_error("index out of bounds: the len is %lld but the index is %lld", (long long)d->len, (long long)index);
            return EML_ERR_STATE;
        }
        *elem_off = (int64_t)(offs[n] / sizeof(T)) + d->nodes[n].element_of(index);
        return EML_OK;
    };
    if (map.all_equal) {
        int64_t off;
        EML_TRY(target(map.idx[0], &off));
        return launch_reduce_sum<T>(t->ctx, g, nd.n, static_cast<T*>(arena->p) + off, true, st);
    }
    if (t->capturing) {
        set_error("a container re-wrapped with arbitrary indexes cannot be swept inside a recorded step (its scatter targets "
                  "are uploaded per sweep)");
        return EML_ERR_STATE;
    }
    d->scatter_offsets.emplace_back(map.n_seg);
    std::vector<int64_t>& off = d->scatter_offsets.back();
    for (int sgm = 0; sgm < map.n_seg; sgm++) EML_TRY(target(map.seg_idx[sgm], &off[sgm]));
    TempBuf dev_off;
    EML_TRY(dev_off.alloc(sizeof(int64_t) * (size_t)map.n_seg, st));
    EML_CUDA(cudaMemcpyAsync(dev_off.p, off.data(), sizeof(int64_t) * (size_t)map.n_seg, cudaMemcpyHostToDevice, st));
    return launch_scatter_segments<T>(g, (const int*)map.perm->p, (const int*)map.seg_start->p, map.n_seg,
                                      static_cast<T*>(arena->p), dev_off.as<int64_t>(), st);
}

// The sweep of a run of scalar entries: parents resolved to arena offsets on the host, one thread walks the run.
template <class T>
int sweep_scalars(eml_tape* t, eml_derivs* d, const Node& nd, int64_t self_off, const Buf& arena, const std::vector<size_t>& offs,
                  cudaStream_t st) {
    if (t->capturing) {
        set_error("scalar Record entries cannot be swept inside a recorded step (the run is uploaded per sweep)");
        return EML_ERR_STATE;
    }
    const size_t n = (size_t)nd.n;
    d->scatter_offsets.emplace_back(2 * n);
    std::vector<int64_t>& off = d->scatter_offsets.back();
    int cached = -1;
    auto resolve = [&](int64_t index, int64_t* out) -> int {
        if (cached < 0 || index < d->nodes[cached].base || index >= d->nodes[cached].base + d->nodes[cached].count)
            cached = node_of_index(d->nodes, index);
        if (cached < 0) {
            set_error("index out of bounds: the len is %lld but the index is %lld", (long long)d->len, (long long)index);
            return EML_ERR_STATE;
        }
        *out = (int64_t)(offs[cached] / sizeof(T)) + d->nodes[cached].element_of(index);
        return EML_OK;
    };
    std::vector<T> der(2 * n);
    for (size_t e = 0; e < n; e++) {
        EML_TRY(resolve(nd.run->lp[e], &off[e]));
        EML_TRY(resolve(nd.run->rp[e], &off[n + e]));
        der[e] = (T)nd.run->ld[e];
        der[n + e] = (T)nd.run->rd[e];
    }
    TempBuf dev_off, dev_der;
    EML_TRY(dev_off.alloc(sizeof(int64_t) * 2 * n, st));
    EML_TRY(dev_der.alloc(sizeof(T) * 2 * n, st));
    EML_CUDA(cudaMemcpyAsync(dev_off.p, off.data(), sizeof(int64_t) * 2 * n, cudaMemcpyHostToDevice, st));
    EML_CUDA(cudaMemcpyAsync(dev_der.p, der.data(), sizeof(T) * 2 * n, cudaMemcpyHostToDevice, st));
    EML_CUDA(cudaStreamSynchronize(st));  // `der` dies with this call
    return launch_scalar_run<T>(static_cast<T*>(arena->p), self_off, (int64_t)n, dev_off.as<int64_t>(), dev_off.as<int64_t>() + n,
                                dev_der.as<T>(), dev_der.as<T>() + n, st);
}

// Does the sweep of matmul node c WRITE (0 + sum, not +=) the derivatives of its parent `parent` when it is the first
// to contribute to them?  True for the kernels that start their fma chain from +0 themselves: the thin (M <= 4)
// backward, and dA = G * B^T through the short-contraction kernel.  (Such a parent needs no zero fill and is not read.)
bool matmul_overwrites(const Node& c, int parent, size_t elem, bool zero_plus) {
    // every product of the sweep can be the first contribution: the thin kernels and the short-contraction kernel start
    // their fma chains from +0, the others are asked for 0 + product (GemmOptions::zero_plus).  EML_ZERO_PLUS=0 (the
    // comparison path of the tests): only the former, everything else accumulates into a zero-filled buffer.
    if (zero_plus) return true;
    if (c.p0 == c.p1) return false;
    if (thin_shape(c.M, c.N, c.K)) return true;
    return parent == c.p0 && smallk_shape(c.M, c.K, c.N, elem);
}

// The reverse sweep over macro nodes.
template <class T>
int sweep_from(eml_tape* t, int nv, int64_t seed_elem, eml_derivs** out);

template <class T>
int tape_sweep(eml_tape* t, eml_val hv, int64_t row, int64_t col, eml_derivs** out) {
    Value* v;
    EML_TRY(get_value(t, hv, &v));
    if (!v->tracked) {
        set_error("Record has no WengertList to find derivatives from");
        return EML_ERR_STATE;
    }
    if (row < 0 || row >= v->rows || col < 0 || col >= v->cols) EML_BAD_ARG("element (%lld,%lld) out of range", (long long)row, (long long)col);
    bool stale;
    int nv = live_node(t, *v, &stale);
    if (stale) {
        set_error("container was not reset after WengertList::clear(); its indexes are stale");
        return EML_ERR_STATE;
    }
    if (t->nodes.empty()) {
        set_error("the WengertList is empty: index out of bounds in the reference");
        return EML_ERR_STATE;
    }
    EML_TRY(flush_open<T>(t));
    return sweep_from<T>(t, nv, row * v->cols + col, out);
}

// Record::try_derivatives (reference src/differentiation.rs:623-648) seeded at element `seed_elem` of node `nv`
// (nv < 0: a container whose every Index is 0).
template <class T>
int sweep_from(eml_tape* t, int nv, int64_t seed_elem, eml_derivs** out) {
    cudaStream_t st = t->stream;
    auto d = std::make_unique<eml_derivs>();
    d->tape = t;
    d->dtype = t->dtype;
    d->device = t->device;
    d->stream = st;
    d->len = t->len;
    d->epoch = t->epoch;
    d->nodes = t->nodes;
    d->groups = t->groups;
    d->have.assign(t->nodes.size(), 1);
    // vec![0; len] (:627): one arena for every node's derivatives (+ the Index-0 slot), one memset
    d->grads.resize(t->nodes.size());
    std::vector<size_t> offs(t->nodes.size() + 1);
    // [0, SLOT_BYTES): the sums that belong to derivatives[0] (contributions of parents carrying Index 0), one slot each
    constexpr size_t SLOT_BYTES = 1024;
    size_t total = SLOT_BYTES;
    for (size_t i = 0; i < t->nodes.size(); i++) {
        offs[i] = total;
        total += (sizeof(T) * (size_t)t->nodes[i].n + 255) & ~size_t(255);
    }
    Buf arena;
    EML_TRY(new_buf(total, st, &arena));
    // vec![0; len]: a node's derivatives only need the zero fill when something ADDS into them first.  The sweep
    // visits consumers in descending order, so the first contribution to node p comes from its highest consumer; when
    // that is an elementwise rule, it writes  0 + dout * deriv  instead (same bits) and p is neither cleared nor read.
    const int n_nodes = (int)t->nodes.size();
    std::vector<int> first_consumer(n_nodes, -1);
    std::vector<int> lowest_consumer(n_nodes, -1);  // the consumer the sweep visits LAST (-2: not a plain node / group)
    auto consumed_by = [&](int target, int consumer, bool plain) {
        first_consumer[target] = consumer;
        if (!plain) lowest_consumer[target] = -2;
        else if (lowest_consumer[target] == -1) lowest_consumer[target] = consumer;  // nodes are visited in ascending order here
    };
    for (int i = 0; i < n_nodes; i++) {
        const Node& nd = t->nodes[i];
        if (nd.kind == Kind::VARIABLES) continue;
        if (nd.kind == Kind::VIEW) {  // its derivatives are ADDED to the nodes its indexes point into
            for (int64_t index : nd.map->seg_idx) {
                const int target = node_of_index(t->nodes, index);
                if (target >= 0) consumed_by(target, i, false);
            }
            continue;
        }
        if (nd.kind == Kind::SCALARS) {  // every entry adds to its parents' entries
            int last = -1;
            for (int64_t e = 0; e < nd.n; e++)
                for (int64_t index : {nd.run->lp[(size_t)e], nd.run->rp[(size_t)e]}) {
                    if (last >= 0 && index >= t->nodes[last].base && index < t->nodes[last].base + t->nodes[last].count) continue;
                    last = node_of_index(t->nodes, index);
                    if (last >= 0) consumed_by(last, i, false);
                }
            continue;
        }
        const bool uses0 = nd.kind == Kind::BINARY ? nd.p0_tracked : true, uses1 = nd.kind == Kind::BINARY ? nd.p1_tracked : nd.kind == Kind::MATMUL;
        if (uses0 && nd.p0 >= 0) consumed_by(nd.p0, i, true);
        if (uses1 && nd.p1 >= 0) consumed_by(nd.p1, i, true);
    }
    // small-part planes written by the sweep kernel of an elementwise group for the backward GEMM that reads the same
    // derivatives as its right operand (per matmul node: the cache identity, 0 = none)
    std::vector<uint64_t> presmall(n_nodes, 0);
    std::vector<char> written(n_nodes, 0);
    d->fed_by.assign(n_nodes, -1);
    d->reduce.assign(n_nodes, eml_derivs::PendingReduce{});
    std::vector<ProductFeed> feeds(n_nodes);  // per group-last node: the product its kernel forms itself (see ProductFeed)
    bool seed_in_kernel = false;
    for (int i = 0; i < n_nodes; i++) {
        const int c = first_consumer[i];
        const int grp = t->nodes[i].group;
        if (grp >= 0 && (c < 0 || t->nodes[c].group == grp)) {
            // every contribution comes from inside its own fused group: that kernel starts it from zero (or the seed)
            written[i] = 0;
            if (i == nv) seed_in_kernel = true;
            continue;
        }
        const bool overwritten_first = c >= 0 && i != nv && t->nodes[c].kind != Kind::VIEW && t->nodes[c].kind != Kind::SCALARS &&
                                       (t->nodes[c].kind != Kind::MATMUL || matmul_overwrites(t->nodes[c], i, sizeof(T), t->zero_plus));
        written[i] = !overwritten_first;  // "written" = holds defined numbers (zeros) before the sweep
    }
    {   // (the slots at the head of the arena are written, not added to: no zeros needed there)
        ZeroRanges z;
        for (int i = 0; i < n_nodes;) {
            if (!written[i]) {
                i++;
                continue;
            }
            int j = i;
            while (j + 1 < n_nodes && written[j + 1]) j++;
            if (z.count == ZERO_RANGES_MAX) {
                EML_TRY(launch_zero_ranges(z, st));
                z.count = 0;
            }
            z.p[z.count] = (char*)arena->p + offs[i];
            z.bytes[z.count++] = (j + 1 < n_nodes ? offs[j + 1] : total) - offs[i];
            i = j + 1;
        }
        EML_TRY(launch_zero_ranges(z, st));
    }
    for (size_t i = 0; i < t->nodes.size(); i++) d->grads[i] = view_of(arena, offs[i], sizeof(T) * (size_t)t->nodes[i].n);
    // Contributions to reference index 0 from parents that carry Index 0 (constants).  The reference adds them to
    // derivatives[0] one after the other in sweep order; here each one is formed in a slot of its own (0 + sum) and the
    // slots are folded in that order when derivatives[0] is needed -- the same sequence of additions, but a sum may be
    // computed on the second stream while the main one moves on.
    T* const slot0_ptr = static_cast<T*>(arena->p);
    const int max_slots = (int)(SLOT_BYTES / sizeof(T));
    int n_slots = 0;
    // ---- second stream (see eml_tape::side) ----
    const bool side = t->side[0] != nullptr;
    bool forked[2] = {false, false};
    std::vector<char> side_busy(n_nodes, 0);  // derivatives of this node have work pending on a second stream
    auto fork = [&](int which) -> int {  // that stream continues from everything enqueued on the main one so far
        EML_CUDA(cudaEventRecord(t->ev_fork, st));
        EML_CUDA(cudaStreamWaitEvent(t->side[which], t->ev_fork, 0));
        forked[which] = true;
        return EML_OK;
    };
    auto join = [&]() -> int {
        for (int which = 0; which < 2; which++) {
            if (!forked[which]) continue;
            EML_CUDA(cudaEventRecord(t->ev_join[which], t->side[which]));
            EML_CUDA(cudaStreamWaitEvent(st, t->ev_join[which], 0));
            forked[which] = false;
        }
        std::fill(side_busy.begin(), side_busy.end(), 0);
        return EML_OK;
    };
    auto busy = [&](int n) { return n >= 0 && side_busy[n]; };
    auto next_slot = [&](T** out) -> int {
        if (n_slots == max_slots) {  // fold what there is into the first slot and go on from the second
            EML_TRY(join());
            EML_TRY(launch_fold_slots<T>(slot0_ptr, slot0_ptr, n_slots, true, st));
            n_slots = 1;
        }
        *out = slot0_ptr + n_slots++;
        return EML_OK;
    };
    // derivatives[self.index] = 1  (:630)
    if (nv >= 0 && seed_in_kernel) {
        // the seed is an argument of the fused group's kernel
    } else if (nv >= 0) {
        EML_TRY(launch_fill<T>((T*)d->grads[nv]->p + seed_elem, T(1), 1, st));
    } else {
        T* slot;
        EML_TRY(next_slot(&slot));
        EML_TRY(launch_fill<T>(slot, T(1), 1, st));
    }
    // dparent (+)= g * (deriv array | constant): the first contribution to an uncleared buffer overwrites
    auto contribute = [&](int p, const T* g, const Buf& deriv, bool is_const, double c, int64_t n) -> int {
        const bool overwrite = !written[p];
        written[p] = 1;
        return launch_chain_ex<T>(g, is_const ? nullptr : (const T*)deriv->p, (T)c, (T*)d->grads[p]->p, overwrite, n, st);
    };

    for (int i = (int)t->nodes.size() - 1; i >= 0; i--) {
        const Node& nd = t->nodes[i];
        const T* g = (const T*)d->grads[i]->p;
        if (nd.kind == Kind::VARIABLES) continue;
        if (forked[0] || forked[1]) {  // anything this node reads or adds to that a second stream is still working on?
            bool wait = nd.kind == Kind::VIEW || nd.kind == Kind::SCALARS;
            if (nd.group >= 0) {
                const Group& grp = t->groups[nd.group];
                for (int k = grp.first; k <= grp.last && !wait; k++)
                    wait = busy(k) || busy(t->nodes[k].p0) || busy(t->nodes[k].p1);
            } else {
                wait = wait || busy(i) || busy(nd.p0) || busy(nd.p1);
            }
            if (wait) EML_TRY(join());
        }
        if (nd.kind == Kind::VIEW) {
            EML_TRY(scatter_view<T>(t, d.get(), nd, g, arena, offs, st));
            continue;
        }
        if (nd.kind == Kind::SCALARS) {
            EML_TRY(sweep_scalars<T>(t, d.get(), nd, (int64_t)(offs[i] / sizeof(T)), arena, offs, st));
            continue;
        }
        if (nd.group >= 0) {  // the whole run [first, last] in one kernel
            const Group& grp = t->groups[nd.group];
            // The group's input is the output of a product whose sweep comes next and reads these derivatives as the
            // right operand of dB = A^T G on THIS stream, and the group is the last to contribute to them: its kernel
            // writes their TF32 small parts too, and that GEMM finds the plane instead of making it (f32 only).
            float* lg_small = nullptr;
            uint64_t gid = 0;
            const int par = t->nodes[grp.first].p0;
            if (sizeof(T) == 4 && t->fuse_packs && par >= 0 && t->nodes[par].kind == Kind::MATMUL && t->nodes[par].p1 >= 0 &&
                t->nodes[par].p0 != t->nodes[par].p1 && lowest_consumer[par] >= grp.first && lowest_consumer[par] <= grp.last &&
                grp.n == t->nodes[par].n) {
                const Node& mm = t->nodes[par];
                const bool leaf1 = t->nodes[mm.p1].kind == Kind::VARIABLES;
                const bool db_on_main = !side || mm.p0 < 0 || !leaf1;  // (the rule of the matmul branch below)
                // dB = A^T[K x M] * G[M x N]: the CTA-pair kernel takes it when both output extents exceed a tile
                const bool tc = mm.K > 128 && mm.N > 128 && mm.M >= 32 && mm.M * mm.N * mm.K >= (int64_t(1) << 24);
                if (db_on_main && tc && !thin_shape(mm.M, mm.N, mm.K)) {
                    gid = new_buffer_identity();
                    EML_TRY(tf32x3_reserve_small_plane_b(gid, (const float*)d->grads[par]->p, mm.M, mm.N, st, &lg_small, nullptr));
                }
            }
            bool small_written = false;
            EML_TRY(group_backward<T>(t->nodes, grp, d->grads, written, d->have, nv, seed_elem, false, st, lg_small, &small_written,
                                      d->fed_by[grp.last] >= 0 ? &feeds[grp.last] : nullptr));
            if (gid) {
                if (small_written) presmall[par] = gid;
                else pack_cache_drop(gid);
            }
            i = grp.first;
            continue;
        }
        if (nd.kind == Kind::UNARY) {
            const bool unit = !nd.a && !nd.a_const;  // the in-place SGD node of a recorded step: derivative 1
            if (nd.p0 >= 0) {
                EML_TRY(contribute(nd.p0, g, nd.a, nd.a_const || unit, unit ? 1.0 : nd.a_c, nd.n));
            } else {
                T* slot;
                EML_TRY(next_slot(&slot));
                if (!nd.a) EML_TRY(launch_reduce_sum<T>(t->ctx, g, nd.n, slot, false, st));
                else EML_TRY(launch_dot_accumulate<T>(t->ctx, g, (const T*)nd.a->p, nd.n, slot, false, st));
            }
        } else if (nd.kind == Kind::BINARY) {
            if (nd.p0_tracked) {
                if (nd.p0 >= 0) {
                    EML_TRY(contribute(nd.p0, g, nd.a, nd.a_const, nd.a_c, nd.n));
                } else {
                    T* slot;
                    EML_TRY(next_slot(&slot));
                    EML_TRY(launch_dot_accumulate<T>(t->ctx, g, (const T*)nd.a->p, nd.n, slot, false, st));
                }
            }
            if (nd.p1_tracked) {
                if (nd.p1 >= 0) {
                    EML_TRY(contribute(nd.p1, g, nd.b, nd.b_const, nd.b_c, nd.n));
                } else {
                    T* slot;
                    EML_TRY(next_slot(&slot));
                    EML_TRY(launch_dot_accumulate<T>(t->ctx, g, (const T*)nd.b->p, nd.n, slot, false, st));
                }
            }
        } else {  // MATMUL: dA += G * B^T, dB += A^T * G
            const int64_t M = nd.M, N = nd.N, K = nd.K;
            const T* A = (const T*)nd.a->p;
            const T* B = (const T*)nd.b->p;
            if (thin_shape(M, N, K) && nd.p0 != nd.p1) {
                // the whole node in one launch: both products and / or the constant-operand sums
                T* dA = nd.p0 >= 0 ? (T*)d->grads[nd.p0]->p : nullptr;
                T* dB = nd.p1 >= 0 ? (T*)d->grads[nd.p1]->p : nullptr;
                const bool over_a = dA && !written[nd.p0], over_b = dB && !written[nd.p1];
                if (dA) written[nd.p0] = 1;
                if (dB) written[nd.p1] = 1;
                T* slot = nullptr;
                if (nd.p0 < 0 || nd.p1 < 0) EML_TRY(next_slot(&slot));
                EML_TRY(launch_thin_bwd<T>(g, A, B, M, N, K, dA, over_a, dB, over_b, nd.p0 < 0, nd.p1 < 0, slot, st));
                continue;
            }
            // The two halves of the node are independent.  One of them moves to the second stream when nothing of this
            // sweep waits for it soon: a constant-operand sum (parked in its slot until derivatives[0] is needed), or
            // the product that ends in a leaf (a variables container: no node below reads it) while the other one
            // continues the chain.
            const bool leaf0 = nd.p0 >= 0 && t->nodes[nd.p0].kind == Kind::VARIABLES;
            const bool leaf1 = nd.p1 >= 0 && t->nodes[nd.p1].kind == Kind::VARIABLES;
            bool a_on_side = false, b_on_side = false;
            if (side && nd.p0 != nd.p1) {
                if (nd.p0 < 0) a_on_side = true;
                else if (nd.p1 < 0) b_on_side = true;
                else if (leaf1) b_on_side = true;
                else if (leaf0) a_on_side = true;
            }
            T* slot = nullptr;
            if (nd.p0 < 0 || nd.p1 < 0) EML_TRY(next_slot(&slot));  // (may join: before the fork)
            // products to side[0], sums to side[1]
            const int which = (a_on_side && nd.p0 < 0) || (b_on_side && nd.p1 < 0) ? 1 : 0;
            if (a_on_side || b_on_side) EML_TRY(fork(which));
            const cudaStream_t sa = a_on_side ? t->side[which] : st, sb = b_on_side ? t->side[which] : st;
            // dA = G * B^T with a short contraction, owed to the last node of a fused group that nothing else contributes
            // to: the group's backward kernel forms it on the fly (no launch, dA neither written nor read)
            bool a_fed = false;
            if (t->feed_products && nd.p0 >= 0 && nd.p0 != nd.p1 && !a_on_side && !written[nd.p0] && nd.p0 != nv &&
                first_consumer[nd.p0] == i && lowest_consumer[nd.p0] == i && N <= EW_PROD_MAX_K && smallk_shape(M, K, N, sizeof(T))) {
                const Node& pn = t->nodes[nd.p0];
                if (pn.group >= 0 && t->groups[pn.group].last == nd.p0 && t->groups[pn.group].n == M * K) {
                    ProductFeed f;
                    f.g = g, f.b = B, f.kc = (int)N, f.cols = K;
                    for (const Value& v : t->values) f.keep |= v.alive && v.node == nd.p0 && v.epoch == t->epoch;
                    if (group_takes_product<T>(t->nodes, t->groups[pn.group], f, written, nv)) {
                        feeds[nd.p0] = f;
                        d->fed_by[nd.p0] = i;
                        a_fed = true;
                    }
                }
            }
            if (a_fed) {
                // (nothing to launch)
            } else if (nd.p0 >= 0) {
                GemmOptions ids;
                ids.b_id = nd.b->identity;
                ids.zero_plus = !written[nd.p0];  // the first contribution: 0 + product, nobody cleared the buffer
                written[nd.p0] = 1;
                EML_TRY(contract_dev<T>(t->ctx, g, B, (T*)d->grads[nd.p0]->p, 1, M, K, N, Strides3{0, N, 1},
                                        Strides3{0, 1, N}, Strides3{0, K, 1}, true, false, sa, &ids));
                if (a_on_side) side_busy[nd.p0] = 1;
            } else {
                // every element of A carries Index 0: derivatives[0] += sum_ik (G B^T)[i,k]
                //   = sum_j colsum(G)[j] * colsum(B)[j]
                EML_TRY(launch_quirk_cols<T>(g, M, B, K, N, slot, sa));
            }
            if (nd.p1 >= 0) {
                GemmOptions ids;
                ids.a_id = nd.a->identity;
                ids.b_id = sb == st ? presmall[i] : 0;  // the small parts of g, written by the group that produced g
                ids.zero_plus = !written[nd.p1];
                // a leaf's gradient nothing else contributes to, on this stream: a split-K launch may leave its partial
                // sums to the weight update (PendingReduce)
                ids.defer_reduce = t->defer_reduce && leaf1 && ids.zero_plus && sb == st && first_consumer[nd.p1] == i &&
                                   lowest_consumer[nd.p1] == i && nd.p1 != nv;
                written[nd.p1] = 1;
                EML_TRY(contract_dev<T>(t->ctx, A, g, (T*)d->grads[nd.p1]->p, 1, K, N, M, Strides3{0, 1, K},
                                        Strides3{0, N, 1}, Strides3{0, N, 1}, true, false, sb, &ids));
                if (ids.deferred_ws) {
                    auto wsb = std::make_shared<DevBuf>();
                    wsb->p = ids.deferred_ws, wsb->bytes = ids.deferred_bytes, wsb->s = sb;
                    d->reduce[nd.p1].ws = wsb;
                    d->reduce[nd.p1].splits = ids.deferred_splits;
                    d->reduce[nd.p1].stride = ids.deferred_stride;
                }
                if (b_on_side) side_busy[nd.p1] = 1;
            } else {
                // derivatives[0] += sum_kj (A^T G)[k,j] = sum_i rowsum(A)[i] * rowsum(G)[i]
                EML_TRY(launch_quirk_rows<T>(A, K, g, N, M, slot, sb));
            }
        }
    }
    EML_TRY(join());
    for (uint64_t id : presmall) pack_cache_drop(id);  // (stream-ordered: the planes outlive the products enqueued above)
    // reference index 0 is element 0 of the first node: added on first use (fold_slot0) or by the weight update
    d->arena = arena;
    d->slot0_pending = n_slots > 0;
    d->slot0_n = n_slots;
    *out = d.release();
    return EML_OK;
}

int find_node(const eml_derivs* d, int64_t index) {
    if (index < 0 || index >= d->len) return -1;
    return node_of_index(d->nodes, index);
}

}  // namespace
}  // namespace eml

#define TAPE_DISPATCH(t, expr_f32, expr_f64) ((t)->dtype == EML_F32 ? (expr_f32) : (expr_f64))

extern "C" {

int eml_tape_create(int dtype, eml_tape** out) {
    if (!out || (dtype != EML_F32 && dtype != EML_F64)) EML_BAD_ARG("tape_create: dtype must be EML_F32 or EML_F64");
    DeviceCtx* ctx = nullptr;
    EML_TRY(get_ctx(&ctx));
    auto* t = new eml_tape();
    t->dtype = dtype;
    t->device = ctx->device;
    t->ctx = ctx;
    t->stream = ctx->stream;
    if (const char* e = getenv("EML_FUSE")) t->fuse = !(e[0] == '0');
    if (const char* e = getenv("EML_EPILOGUE")) t->defer_matmul = e[0] == '1';
    if (const char* e = getenv("EML_ZERO_PLUS")) t->zero_plus = !(e[0] == '0');
    if (const char* e = getenv("EML_FEED_PRODUCTS")) t->feed_products = e[0] == '1';
    if (const char* e = getenv("EML_DEFER_REDUCE")) t->defer_reduce = e[0] == '1';
    if (const char* e = getenv("EML_FUSE_PACKS")) t->fuse_packs = !(e[0] == '0');
    const char* se = getenv("EML_SIDE");
    if (!(se && se[0] == '0')) {
        cudaError_t e = cudaEventCreateWithFlags(&t->ev_fork, cudaEventDisableTiming);
        for (int which = 0; which < 2 && e == cudaSuccess; which++) {
            e = cudaStreamCreateWithFlags(&t->side[which], cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_join[which], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) {
            delete t;
            return cuda_fail(e, "second stream of the tape", __FILE__, __LINE__);
        }
    }
    *out = t;
    return EML_OK;
}

int eml_tape_destroy(eml_tape* t) {
    if (!t) return EML_OK;
    check_tape(t);
    cudaStreamSynchronize(t->stream);
    for (int which = 0; which < 2; which++) {
        if (t->side[which]) {
            cudaStreamSynchronize(t->side[which]);
            cudaStreamDestroy(t->side[which]);
        }
        if (t->ev_join[which]) cudaEventDestroy(t->ev_join[which]);
    }
    if (t->ev_fork) cudaEventDestroy(t->ev_fork);
    delete t;
    return EML_OK;
}

int eml_tape_len(eml_tape* t, int64_t* len) {
    if (!t || !len) EML_BAD_ARG("null argument");
    *len = t->len;
    return EML_OK;
}

int eml_tape_clear(eml_tape* t) {
    EML_TRY(check_tape(t));
    std::lock_guard<std::mutex> lock(t->mu);
    EML_TRY(flush(t));  // containers created before the clear stay usable as numbers
    t->nodes.clear();
    t->groups.clear();
    t->len = 0;
    t->epoch++;
    return EML_OK;
}

int eml_tape_variables(eml_tape* t, const void* host, int64_t rows, int64_t cols, eml_val* out) {
    EML_TRY(check_tape(t));
    std::lock_guard<std::mutex> lock(t->mu);
    return TAPE_DISPATCH(t, new_container<float>(t, host, false, rows, cols, true, out),
                         new_container<double>(t, host, false, rows, cols, true, out));
}
int eml_tape_constants(eml_tape* t, const void* host, int64_t rows, int64_t cols, eml_val* out) {
    EML_TRY(check_tape(t));
    std::lock_guard<std::mutex> lock(t->mu);
    return TAPE_DISPATCH(t, new_container<float>(t, host, false, rows, cols, false, out),
                         new_container<double>(t, host, false, rows, cols, false, out));
}
int eml_tape_variables_dev(eml_tape* t, const void* dev, int64_t rows, int64_t cols, eml_val* out) {
    EML_TRY(check_tape(t));
    std::lock_guard<std::mutex> lock(t->mu);
    return TAPE_DISPATCH(t, new_container<float>(t, dev, true, rows, cols, true, out),
                         new_container<double>(t, dev, true, rows, cols, true, out));
}
int eml_tape_constants_dev(eml_tape* t, const void* dev, int64_t rows, int64_t cols, eml_val* out) {
    EML_TRY(check_tape(t));
    std::lock_guard<std::mutex> lock(t->mu);
    return TAPE_DISPATCH(t, new_container<float>(t, dev, true, rows, cols, false, out),
                         new_container<double>(t, dev, true, rows, cols, false, out));
}

int eml_tape_from_existing(eml_tape* t, const void* host_numbers, const int64_t* host_indexes, int64_t rows, int64_t cols,
                           int has_history, eml_val* out) {
    EML_TRY(check_tape(t));
    if (!host_numbers || !host_indexes || !out) EML_BAD_ARG("null argument");
    if (rows < 1 || cols < 1) EML_BAD_ARG("container shape %lldx%lld: Easy ML containers have at least one element",
                                          (long long)rows, (long long)cols);
    std::lock_guard<std::mutex> lock(t->mu);
    EML_TRY(not_while_capturing(t, "creating a container from host memory"));
    EML_TRY(flush(t));
    std::shared_ptr<IndexMap> map;
    EML_TRY(build_index_map(t, std::vector<int64_t>(host_indexes, host_indexes + rows * cols), has_history != 0, &map));
    Buf numbers;
    const size_t bytes = esize(t->dtype) * (size_t)(rows * cols);
    EML_TRY(new_buf(bytes, t->stream, &numbers));
    EML_CUDA(cudaMemcpyAsync(numbers->p, host_numbers, bytes, cudaMemcpyHostToDevice, t->stream));
    EML_CUDA(cudaStreamSynchronize(t->stream));
    return put_view(t, numbers, rows, cols, map, has_history != 0, out);
}

int eml_tape_range(eml_tape* t, eml_val src, int64_t row0, int64_t n_rows, int64_t col0, int64_t n_cols, int has_history,
                   eml_val* out) {
    EML_TRY(check_tape(t));
    if (!out) EML_BAD_ARG("null argument");
    std::lock_guard<std::mutex> lock(t->mu);
    Value* v;
    EML_TRY(get_value(t, src, &v));
    if (row0 < 0 || col0 < 0 || n_rows < 1 || n_cols < 1 || row0 + n_rows > v->rows || col0 + n_cols > v->cols)
        EML_BAD_ARG("range [%lld, +%lld) x [%lld, +%lld) does not fit a %lldx%lld container", (long long)row0, (long long)n_rows,
                    (long long)col0, (long long)n_cols, (long long)v->rows, (long long)v->cols);
    bool stale;
    const int n = live_node(t, *v, &stale);
    if (stale) {
        set_error("container was not reset after WengertList::clear(); its indexes are stale");
        return EML_ERR_STATE;
    }
    EML_TRY(flush(t));
    std::vector<int64_t> idx((size_t)(n_rows * n_cols));
    for (int64_t r = 0; r < n_rows; r++)
        for (int64_t c = 0; c < n_cols; c++) {
            const int64_t e = (row0 + r) * v->cols + col0 + c;
            idx[(size_t)(r * n_cols + c)] = v->map ? v->map->idx[(size_t)e] : (n < 0 ? 0 : t->nodes[n].index_of(e));
        }
    std::shared_ptr<IndexMap> map;
    EML_TRY(build_index_map(t, std::move(idx), has_history != 0, &map));
    const size_t es = esize(t->dtype);
    Buf numbers;
    EML_TRY(new_buf(es * (size_t)(n_rows * n_cols), t->stream, &numbers));
    EML_CUDA(cudaMemcpy2DAsync(numbers->p, n_cols * es, (const char*)v->numbers->p + (row0 * v->cols + col0) * es, v->cols * es,
                               n_cols * es, n_rows, cudaMemcpyDeviceToDevice, t->stream));
    return put_view(t, numbers, n_rows, n_cols, map, has_history != 0, out);
}

static int append_scalar(eml_tape* t, int64_t lp, double ld, int64_t rp, double rd, int64_t* index) {
    EML_TRY(check_tape(t));
    if (!index) EML_BAD_ARG("null argument");
    std::lock_guard<std::mutex> lock(t->mu);
    // a missing parent is the entry's OWN index with derivative 0, as in the reference (:811-846): "won't be needed but
    // will always be valid"
    if (lp == -1) lp = t->len;
    if (rp == -1) rp = t->len;
    if (lp < 0 || rp < 0 || lp > t->len || rp > t->len)
        EML_BAD_ARG("scalar entry: parent index out of bounds (the len is %lld)", (long long)t->len);
    if (t->nodes.empty() || t->nodes.back().kind != Kind::SCALARS || t->open >= 0) {
        EML_TRY(flush(t));
        Node nd;
        nd.kind = Kind::SCALARS;
        nd.base = t->len;
        nd.run = std::make_shared<ScalarRun>();
        t->nodes.push_back(nd);
    }
    Node& nd = t->nodes.back();
    nd.run->lp.push_back(lp);
    nd.run->rp.push_back(rp);
    nd.run->ld.push_back(ld);
    nd.run->rd.push_back(rd);
    nd.n = ++nd.count;
    *index = t->len++;
    return EML_OK;
}

int eml_tape_append_nullary(eml_tape* t, int64_t* index) { return append_scalar(t, -1, 0.0, -1, 0.0, index); }
int eml_tape_append_unary(eml_tape* t, int64_t parent, double derivative, int64_t* index) {
    if (parent < 0) EML_BAD_ARG("scalar entry: negative parent index");
    return append_scalar(t, parent, derivative, -1, 0.0, index);
}
int eml_tape_append_binary(eml_tape* t, int64_t left_parent, double left_derivative, int64_t right_parent,
                           double right_derivative, int64_t* index) {
    if (left_parent < 0 || right_parent < 0) EML_BAD_ARG("scalar entry: negative parent index");
    return append_scalar(t, left_parent, left_derivative, right_parent, right_derivative, index);
}

int eml_tape_derivatives_at(eml_tape* t, int64_t index, eml_derivs** out) {
    EML_TRY(check_tape(t));
    if (!out) EML_BAD_ARG("null argument");
    std::lock_guard<std::mutex> lock(t->mu);
    EML_TRY(flush(t));
    const int nv = node_of_index(t->nodes, index);
    if (nv < 0) EML_BAD_ARG("index out of bounds: the len is %lld but the index is %lld", (long long)t->len, (long long)index);
    const int64_t elem = t->nodes[nv].element_of(index);
    if (t->nodes[nv].index_of(elem) != index)
        EML_BAD_ARG("index %lld is an entry inside a matrix product, not a Record's", (long long)index);
    return TAPE_DISPATCH(t, sweep_from<float>(t, nv, elem, out), sweep_from<double>(t, nv, elem, out));
}

int eml_tape_reset(eml_tape* t, eml_val h) {
    EML_TRY(check_tape(t));
    std::lock_guard<std::mutex> lock(t->mu);
    Value* v;
    EML_TRY(get_value(t, h, &v));
    if (!v->tracked) return EML_OK;  // history None => noop (reference mod.rs:351)
    EML_TRY(flush(t));
    Node nd;
    nd.kind = Kind::VARIABLES;
    nd.base = t->len;
    nd.count = nd.n = v->n();
    t->len += nd.count;
    t->nodes.push_back(nd);
    v->node = (int)t->nodes.size() - 1;
    v->epoch = t->epoch;
    v->map.reset();  // renumbered start + i: an ordinary variables container from here on
    return EML_OK;
}

int eml_val_release(eml_tape* t, eml_val h) {
    EML_TRY(check_tape(t));
    std::lock_guard<std::mutex> lock(t->mu);
    Value* v;
    EML_TRY(get_value(t, h, &v));
    if (!v->numbers && t->open >= 0) {
        // a container of the open group dropped before anybody needed its numbers: they will never be stored
        Group& g = t->groups[t->open];
        if (v->epoch == t->epoch && v->node >= g.first && v->node <= g.last && g.owner[v->node - g.first] == h)
            g.owner[v->node - g.first] = -1;
    }
    *v = Value();
    t->free_slots.push_back(h);
    return EML_OK;
}

int eml_val_shape(eml_tape* t, eml_val h, int64_t* rows, int64_t* cols, int* tracked) {
    if (!t) EML_BAD_ARG("null tape");
    std::lock_guard<std::mutex> lock(t->mu);
    Value* v;
    EML_TRY(get_value(t, h, &v));
    if (rows) *rows = v->rows;
    if (cols) *cols = v->cols;
    if (tracked) *tracked = v->tracked ? 1 : 0;
    return EML_OK;
}

int eml_val_numbers(eml_tape* t, eml_val h, void* host_out) {
    EML_TRY(check_tape(t));
    std::lock_guard<std::mutex> lock(t->mu);
    Value* v;
    EML_TRY(get_value(t, h, &v));
    EML_TRY(not_while_capturing(t, "eml_val_numbers"));
    EML_TRY(flush(t));
    EML_CUDA(cudaMemcpyAsync(host_out, v->numbers->p, esize(t->dtype) * v->n(), cudaMemcpyDeviceToHost, t->stream));
    EML_CUDA(cudaStreamSynchronize(t->stream));
    return EML_OK;
}

int eml_val_numbers_async(eml_tape* t, eml_val h, void* pinned_host_out, void** arrival) {
    EML_TRY(check_tape(t));
    if (!pinned_host_out) EML_BAD_ARG("null argument");
    std::lock_guard<std::mutex> lock(t->mu);
    Value* v;
    EML_TRY(get_value(t, h, &v));
    {
        // a few KB (a loss value): stored by a kernel through the buffer's device mapping -- by the queued weight
        // updates' own launch when there is one and the numbers already exist; anything else, or memory that is not
        // page-locked / mapped, takes the copy engine
        const size_t bytes = esize(t->dtype) * (size_t)v->n();
        void* mapped = nullptr;
        const bool small = bytes <= 4096 && cudaHostGetDevicePointer(&mapped, pinned_host_out, 0) == cudaSuccess && mapped;
        const bool riding = small && v->numbers && !t->sgd.empty() && !t->rider_src;
        if (riding) {
            t->rider_src = v->numbers;
            t->rider_dst = mapped;
            t->rider_bytes = bytes;
        }
        const int flushed = flush(t);
        if (flushed != EML_OK) {
            t->rider_src.reset();
            return flushed;
        }
        if (riding) {
            // went out with the updates
        } else if (small) {
            EML_TRY(launch_copy_to_host(v->numbers->p, mapped, bytes, t->stream));
        } else {
            cudaGetLastError();
            EML_CUDA(cudaMemcpyAsync(pinned_host_out, v->numbers->p, bytes, cudaMemcpyDeviceToHost, t->stream));
        }
    }
    if (arrival && t->capturing) {
        *arrival = nullptr;  // inside a recorded step the arrival comes from eml_graph_launch
    } else if (arrival) {
        cudaEvent_t ev;
        EML_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        EML_CUDA(cudaEventRecord(ev, t->stream));
        *arrival = ev;
    }
    return EML_OK;
}

int eml_arrival_wait(void* arrival) {
    if (!arrival) EML_BAD_ARG("null arrival");
    cudaEvent_t ev = (cudaEvent_t)arrival;
    cudaError_t e = cudaEventSynchronize(ev);
    cudaEventDestroy(ev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaEventSynchronize", __FILE__, __LINE__);
    return EML_OK;
}

int eml_tape_sync(eml_tape* t) {
    EML_TRY(check_tape(t));
    EML_TRY(not_while_capturing(t, "eml_tape_sync"));
    {
        std::lock_guard<std::mutex> lock(t->mu);
        EML_TRY(flush(t));
    }
    EML_CUDA(cudaStreamSynchronize(t->stream));
    return EML_OK;
}

int eml_tape_capture_begin(eml_tape* t) {
    EML_TRY(check_tape(t));
    std::lock_guard<std::mutex> lock(t->mu);
    if (t->capturing) {
        set_error("a step is already being captured on this tape");
        return EML_ERR_STATE;
    }
    EML_TRY(flush(t));  // work requested before the recording is not part of it
    EML_CUDA(cudaStreamBeginCapture(t->stream, cudaStreamCaptureModeThreadLocal));
    t->capturing = true;
    t->capture_len = t->len;
    t->capture_nodes = t->nodes.size();
    stream_alloc_forbid_new(true);
    pdl_suppress(true);  // recorded kernels take plain stream-order edges
    return EML_OK;
}

int eml_tape_capture_end(eml_tape* t, eml_graph** out) {
    EML_TRY(check_tape(t));
    if (!out) EML_BAD_ARG("null argument");
    std::lock_guard<std::mutex> lock(t->mu);
    if (!t->capturing) {
        set_error("no capture in progress on this tape");
        return EML_ERR_STATE;
    }
    const int flushed = flush(t);  // work requested inside the recording belongs to it
    t->capturing = false;
    stream_alloc_forbid_new(false);
    pdl_suppress(false);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(t->stream, &graph);
    if (flushed != EML_OK) {
        if (graph) cudaGraphDestroy(graph);
        return flushed;
    }
    if (e != cudaSuccess || !graph) {
        cudaGetLastError();
        set_error("the step could not be captured: %s", cudaGetErrorString(e));
        return EML_ERR_CUDA;
    }
    if (t->len != t->capture_len || t->nodes.size() != t->capture_nodes) {
        cudaGraphDestroy(graph);
        set_error("a recorded step must end in the tape state it started in (length %lld -> %lld): clear() the list and "
                  "reset() the variables before eml_tape_capture_end", (long long)t->capture_len, (long long)t->len);
        return EML_ERR_STATE;
    }
    auto g = std::make_unique<eml_graph>();
    g->tape = t;
    g->graph = graph;
    e = cudaGraphInstantiate(&g->exec, graph, 0);
    if (e != cudaSuccess) {
        cudaGraphDestroy(graph);
        return cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__);
    }
    *out = g.release();
    return EML_OK;
}

int eml_graph_launch(eml_graph* g, void** arrival) {
    if (!g || !g->exec) EML_BAD_ARG("null graph");
    EML_TRY(check_tape(g->tape));
    EML_CUDA(cudaGraphLaunch(g->exec, g->tape->stream));
    count_launch();
    if (arrival) {
        cudaEvent_t ev;
        EML_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        EML_CUDA(cudaEventRecord(ev, g->tape->stream));
        *arrival = ev;
    }
    return EML_OK;
}

int eml_graph_profile(eml_graph* g, int replays, char* report, int64_t capacity) {
    if (!g || !g->exec || !report || capacity < 1 || replays < 1) EML_BAD_ARG("graph_profile: bad argument");
    EML_TRY(check_tape(g->tape));
    cudaStream_t st = g->tape->stream;
    size_t n = 0;
    EML_CUDA(cudaGraphGetNodes(g->graph, nullptr, &n));
    std::vector<cudaGraphNode_t> nodes(n);
    if (n) EML_CUDA(cudaGraphGetNodes(g->graph, nodes.data(), &n));
    cudaEvent_t e0, e1;
    EML_CUDA(cudaEventCreate(&e0));
    EML_CUDA(cudaEventCreate(&e1));
    auto time_us = [&](double* us) -> int {
        for (int i = 0; i < 3; i++) EML_CUDA(cudaGraphLaunch(g->exec, st));
        EML_CUDA(cudaEventRecord(e0, st));
        for (int i = 0; i < replays; i++) EML_CUDA(cudaGraphLaunch(g->exec, st));
        EML_CUDA(cudaEventRecord(e1, st));
        EML_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        EML_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        *us = (double)ms * 1e3 / replays;
        return EML_OK;
    };
    std::string out;
    char line[512];
    struct K {
        cudaGraphNode_t node;
        std::string name;
    };
    std::vector<K> kernels;
    for (cudaGraphNode_t nd : nodes) {
        cudaGraphNodeType type;
        EML_CUDA(cudaGraphNodeGetType(nd, &type));
        if (type == cudaGraphNodeTypeMemcpy || type == cudaGraphNodeTypeMemset) {
            kernels.push_back(K{nd, type == cudaGraphNodeTypeMemcpy ? "(memcpy node)" : "(memset node)"});
            continue;
        }
        if (type != cudaGraphNodeTypeKernel) continue;
        cudaKernelNodeParams kp{};
        EML_CUDA(cudaGraphKernelNodeGetParams(nd, &kp));
        const char* name = nullptr;
        if (cudaFuncGetName(&name, kp.func) != cudaSuccess || !name) name = "?", cudaGetLastError();
        kernels.push_back(K{nd, name});
    }
    double whole = 0, base = 0;
    EML_TRY(time_us(&whole));
    // the weight updates stay off from here on: with a kernel missing the step computes garbage, and garbage must not
    // accumulate in the weights (non-finite weights would send the f32 products down their repair path)
    for (const K& k : kernels)
        if (k.name.find("sgd") != std::string::npos) EML_CUDA(cudaGraphNodeSetEnabled(g->exec, k.node, 0));
    EML_TRY(time_us(&base));
    snprintf(line, sizeof line, "%.2f\tus per replay, %zu nodes (%zu kernels)\n%.2f\tus per replay without the weight updates\n", whole,
             n, kernels.size(), base);
    out += line;
    for (const K& k : kernels) {
        if (k.name.find("sgd") != std::string::npos) continue;
        EML_CUDA(cudaGraphNodeSetEnabled(g->exec, k.node, 0));
        double us = 0;
        EML_TRY(time_us(&us));
        EML_CUDA(cudaGraphNodeSetEnabled(g->exec, k.node, 1));
        snprintf(line, sizeof line, "%+.2f\t%s\n", us - base, k.name.substr(0, 160).c_str());
        out += line;
    }
    for (const K& k : kernels)
        if (k.name.find("sgd") != std::string::npos) EML_CUDA(cudaGraphNodeSetEnabled(g->exec, k.node, 1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const size_t len = out.size() < (size_t)capacity - 1 ? out.size() : (size_t)capacity - 1;
    memcpy(report, out.data(), len);
    report[len] = 0;
    return EML_OK;
}

int eml_graph_destroy(eml_graph* g) {
    if (!g) return EML_OK;
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
    return EML_OK;
}

int eml_val_indexes(eml_tape* t, eml_val h, int64_t* host_out) {
    if (!t || !host_out) EML_BAD_ARG("null argument");
    std::lock_guard<std::mutex> lock(t->mu);
    Value* v;
    EML_TRY(get_value(t, h, &v));
    bool stale;
    int n = live_node(t, *v, &stale);
    if (stale) {
        set_error("container was not reset after WengertList::clear(); its indexes are stale");
        return EML_ERR_STATE;
    }
    if (v->map) {
        memcpy(host_out, v->map->idx.data(), sizeof(int64_t) * (size_t)v->n());
        return EML_OK;
    }
    for (int64_t e = 0; e < v->n(); e++) host_out[e] = n < 0 ? 0 : t->nodes[n].index_of(e);
    return EML_OK;
}

int eml_val_device_ptr(eml_tape* t, eml_val h, const void** dev) {
    if (!t || !dev) EML_BAD_ARG("null argument");
    std::lock_guard<std::mutex> lock(t->mu);
    Value* v;
    EML_TRY(get_value(t, h, &v));
    EML_TRY(check_tape(t));
    EML_TRY(flush(t));
    *dev = v->numbers->p;
    return EML_OK;
}

int eml_tape_matmul(eml_tape* t, eml_val a, eml_val b, eml_val* c) {
    EML_TRY(check_tape(t));
    if (!c) EML_BAD_ARG("null argument");
    std::lock_guard<std::mutex> lock(t->mu);
    return TAPE_DISPATCH(t, tape_matmul<float>(t, a, b, c), tape_matmul<double>(t, a, b, c));
}
int eml_tape_unary(eml_tape* t, int op, double cst, eml_val x, eml_val* y) {
    EML_TRY(check_tape(t));
    if (!y) EML_BAD_ARG("null argument");
    std::lock_guard<std::mutex> lock(t->mu);
    return TAPE_DISPATCH(t, tape_unary<float>(t, op, cst, x, y), tape_unary<double>(t, op, cst, x, y));
}
int eml_tape_binary(eml_tape* t, int op, eml_val x, eml_val y, eml_val* z) {
    EML_TRY(check_tape(t));
    if (!z) EML_BAD_ARG("null argument");
    std::lock_guard<std::mutex> lock(t->mu);
    return TAPE_DISPATCH(t, tape_binary<float>(t, op, x, y, z), tape_binary<double>(t, op, x, y, z));
}

int eml_tape_derivatives(eml_tape* t, eml_val v, int64_t row, int64_t col, eml_derivs** out) {
    EML_TRY(check_tape(t));
    if (!out) EML_BAD_ARG("null argument");
    std::lock_guard<std::mutex> lock(t->mu);
    return TAPE_DISPATCH(t, tape_sweep<float>(t, v, row, col, out), tape_sweep<double>(t, v, row, col, out));
}

int eml_derivs_free(eml_derivs* d) {
    if (!d) return EML_OK;
    cudaSetDevice(d->device);
    delete d;
    return EML_OK;
}

int eml_derivs_len(eml_derivs* d, int64_t* len) {
    if (!d || !len) EML_BAD_ARG("null argument");
    *len = d->len;
    return EML_OK;
}

// Derivatives of a node the sweep kept in registers only (interior of a fused elementwise group, no container held):
// run the group's backward program once more, storing them (nothing else is touched).
// the deferred split-K reduction of a weight gradient (eml_derivs::PendingReduce), launched for a reader other than the
// weight update
static int materialise_reduce(eml_derivs* d, int n) {
    if (n < 0 || n >= (int)d->reduce.size() || !d->reduce[n].ws) return EML_OK;
    eml_derivs::PendingReduce& r = d->reduce[n];
    const int64_t cnt = d->nodes[n].n;
    if (d->dtype == EML_F32)
        EML_TRY(launch_splitk_reduce<float>((const float*)r.ws->p, r.splits, (float*)d->grads[n]->p, 1, 1, cnt, Strides3{cnt, cnt, 1},
                                            ACC_ZERO_PLUS, d->stream));
    else
        EML_TRY(launch_splitk_reduce<double>((const double*)r.ws->p, r.splits, (double*)d->grads[n]->p, 1, 1, cnt, Strides3{cnt, cnt, 1},
                                             ACC_ZERO_PLUS, d->stream));
    r.ws.reset();
    return EML_OK;
}

static int ensure_have(eml_derivs* d, int n) {
    EML_TRY(materialise_reduce(d, n));
    if (d->have[n]) return EML_OK;
    const Group& g = d->groups[d->nodes[n].group];
    if (!d->have[g.last] && g.last < (int)d->fed_by.size() && d->fed_by[g.last] >= 0) {
        // the product the sweep folded into this group's kernel (ProductFeed): launched now, the same fma chains from +0
        const Node& mm = d->nodes[d->fed_by[g.last]];
        const void* G = d->grads[d->fed_by[g.last]]->p;
        if (d->dtype == EML_F32)
            EML_TRY(contract_dev<float>(d->tape->ctx, (const float*)G, (const float*)mm.b->p, (float*)d->grads[g.last]->p, 1, mm.M, mm.K, mm.N,
                                        Strides3{0, mm.N, 1}, Strides3{0, 1, mm.N}, Strides3{0, mm.K, 1}, false, false, d->stream));
        else
            EML_TRY(contract_dev<double>(d->tape->ctx, (const double*)G, (const double*)mm.b->p, (double*)d->grads[g.last]->p, 1, mm.M, mm.K,
                                         mm.N, Strides3{0, mm.N, 1}, Strides3{0, 1, mm.N}, Strides3{0, mm.K, 1}, false, false, d->stream));
        d->have[g.last] = 1;
        if (d->have[n]) return EML_OK;
    }
    std::vector<char> written(d->have);
    if (d->dtype == EML_F32) return group_backward<float>(d->nodes, g, d->grads, written, d->have, -1, -1, true, d->stream);
    return group_backward<double>(d->nodes, g, d->grads, written, d->have, -1, -1, true, d->stream);
}

static int fold_slot0(eml_derivs* d) {
    if (!d->slot0_pending) return EML_OK;
    EML_TRY(materialise_reduce(d, 0));  // (the parked sums are added to element 0 of node 0: its numbers come first)
    EML_TRY(flush(d->tape));  // a queued weight update reads derivatives[0] and the parked sum separately
    d->slot0_pending = false;
    if (d->dtype == EML_F32) return launch_fold_slots<float>((float*)d->grads[0]->p, (const float*)d->arena->p, d->slot0_n, false, d->stream);
    return launch_fold_slots<double>((double*)d->grads[0]->p, (const double*)d->arena->p, d->slot0_n, false, d->stream);
}

static int derivs_not_while_capturing(eml_derivs* d) {
    if (d && d->tape && d->tape->capturing) {
        set_error("reading derivatives on the host blocks on the device and cannot be part of a recorded step");
        return EML_ERR_STATE;
    }
    return EML_OK;
}

int eml_derivs_at(eml_derivs* d, int64_t index, double* out) {
    EML_TRY(derivs_not_while_capturing(d));
    if (!d || !out) EML_BAD_ARG("null argument");
    int n = find_node(d, index);
    if (n < 0) EML_BAD_ARG("index out of bounds: the len is %lld but the index is %lld", (long long)d->len, (long long)index);
    int64_t e = d->nodes[n].element_of(index);
    EML_CUDA(cudaSetDevice(d->device));
    EML_TRY(fold_slot0(d));
    EML_TRY(ensure_have(d, n));
    if (d->dtype == EML_F32) {
        float f;
        EML_CUDA(cudaMemcpyAsync(&f, (const float*)d->grads[n]->p + e, 4, cudaMemcpyDeviceToHost, d->stream));
        EML_CUDA(cudaStreamSynchronize(d->stream));
        *out = f;
    } else {
        EML_CUDA(cudaMemcpyAsync(out, (const double*)d->grads[n]->p + e, 8, cudaMemcpyDeviceToHost, d->stream));
        EML_CUDA(cudaStreamSynchronize(d->stream));
    }
    return EML_OK;
}

static int derivs_at_val(eml_derivs* d, eml_val h, void* dst, bool dst_on_device) {
    if (!d || !dst) EML_BAD_ARG("null argument");
    eml_tape* t = d->tape;
    EML_TRY(check_tape(t));
    std::lock_guard<std::mutex> lock(t->mu);
    Value* v;
    EML_TRY(get_value(t, h, &v));
    size_t es = esize(d->dtype);
    EML_TRY(fold_slot0(d));
    int n = v->node;
    if (n >= 0 && (v->epoch != d->epoch || n >= (int)d->nodes.size())) {
        set_error("container does not belong to the computation these derivatives were taken from");
        return EML_ERR_STATE;
    }
    cudaMemcpyKind kind = dst_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (v->map) {
        // a re-wrapped container: element e reads derivatives[index[e]] (Derivatives::at_tensor looks every Index up)
        const IndexMap& map = *v->map;
        std::vector<int64_t> distinct = map.seg_idx;
        if (distinct.empty()) {  // history None: no grouping was prepared
            distinct = map.idx;
            std::sort(distinct.begin(), distinct.end());
            distinct.erase(std::unique(distinct.begin(), distinct.end()), distinct.end());
        }
        std::vector<char> got(es * distinct.size());
        for (size_t sgm = 0; sgm < distinct.size(); sgm++) {
            const int tn = find_node(d, distinct[sgm]);
            if (tn < 0) EML_BAD_ARG("index out of bounds: the len is %lld but the index is %lld", (long long)d->len, (long long)distinct[sgm]);
            EML_TRY(ensure_have(d, tn));
            EML_CUDA(cudaMemcpyAsync(got.data() + es * sgm, (const char*)d->grads[tn]->p + es * d->nodes[tn].element_of(distinct[sgm]), es,
                                     cudaMemcpyDeviceToHost, d->stream));
        }
        EML_CUDA(cudaStreamSynchronize(d->stream));
        std::vector<char> host(es * (size_t)v->n());
        for (int64_t e = 0; e < v->n(); e++) {
            const size_t sgm = std::lower_bound(distinct.begin(), distinct.end(), map.idx[(size_t)e]) - distinct.begin();
            memcpy(host.data() + es * e, got.data() + es * sgm, es);
        }
        if (dst_on_device) {
            EML_CUDA(cudaMemcpyAsync(dst, host.data(), host.size(), cudaMemcpyHostToDevice, d->stream));
            EML_CUDA(cudaStreamSynchronize(d->stream));
        } else {
            memcpy(dst, host.data(), host.size());
        }
        return EML_OK;
    }
    if (n >= 0) {
        EML_TRY(ensure_have(d, n));
        EML_CUDA(cudaMemcpyAsync(dst, d->grads[n]->p, es * v->n(), kind, d->stream));
    } else {
        // every Index is 0: each element reads derivatives[0]
        for (int64_t e = 0; e < v->n(); e++)
            EML_CUDA(cudaMemcpyAsync((char*)dst + es * e, d->grads[0]->p, es, kind, d->stream));
    }
    if (!dst_on_device) EML_CUDA(cudaStreamSynchronize(d->stream));
    return EML_OK;
}

int eml_derivs_at_val(eml_derivs* d, eml_val v, void* host_out) {
    EML_TRY(derivs_not_while_capturing(d));
    return derivs_at_val(d, v, host_out, false);
}
int eml_derivs_at_val_dev(eml_derivs* d, eml_val v, void* dev_out) { return derivs_at_val(d, v, dev_out, true); }

int eml_derivs_to_host(eml_derivs* d, void* host_out) {
    if (!d || !host_out) EML_BAD_ARG("null argument");
    EML_TRY(derivs_not_while_capturing(d));
    if (d->len > (int64_t(1) << 28))
        EML_BAD_ARG("derivs_to_host: the reference tape would have %lld entries; materialising it is refused", (long long)d->len);
    EML_CUDA(cudaSetDevice(d->device));
    {
        std::lock_guard<std::mutex> lock(d->tape->mu);
        EML_TRY(fold_slot0(d));
    }
    size_t es = esize(d->dtype);
    std::vector<char> tmp;
    for (size_t i = 0; i < d->nodes.size(); i++) {
        const Node& nd = d->nodes[i];
        if (nd.kind == Kind::VIEW) continue;  // stands for no entries
        EML_TRY(ensure_have(d, (int)i));
        tmp.resize(es * nd.n);
        EML_CUDA(cudaMemcpyAsync(tmp.data(), d->grads[i]->p, es * nd.n, cudaMemcpyDeviceToHost, d->stream));
        EML_CUDA(cudaStreamSynchronize(d->stream));
        char* dst = (char*)host_out + es * nd.base;
        if (nd.kind == Kind::MATMUL && nd.K > 1) {
            int64_t seg = 2 * nd.K - 1;
            for (int64_t e = 0; e < nd.n; e++)
                for (int64_t s = 0; s < seg; s++) memcpy(dst + es * (e * seg + s), tmp.data() + es * e, es);
        } else {
            memcpy(dst, tmp.data(), es * nd.n);
        }
    }
    return EML_OK;
}

int eml_tape_sgd_update(eml_tape* t, eml_val hw, eml_derivs* d, double lr) {
    EML_TRY(check_tape(t));
    if (!d) EML_BAD_ARG("null argument");
    std::lock_guard<std::mutex> lock(t->mu);
    Value* w;
    EML_TRY(get_value(t, hw, &w));
    // containers of the open group get their numbers now; updates already queued stay queued (they go out together)
    // unless this one reads what one of them writes
    bool chained = false;
    for (const auto& u : t->sgd) chained |= w->numbers && u.out == w->numbers;
    if (chained) EML_TRY(flush(t));
    else EML_TRY(t->dtype == EML_F32 ? flush_groups<float>(t) : flush_groups<double>(t));
    if (!w->tracked || w->node < 0 || w->map || w->epoch != t->epoch || d->epoch != t->epoch || w->node >= (int)d->nodes.size()) {
        set_error("sgd_update: container is not a tracked variable of the computation these derivatives belong to");
        return EML_ERR_STATE;
    }
    const bool deferred = w->node < (int)d->reduce.size() && d->reduce[w->node].ws != nullptr;
    if (!deferred) EML_TRY(ensure_have(d, w->node));
    const int64_t n = w->n();
    size_t es = esize(t->dtype);
    // x - (derivatives[&x] * lr): Record - T => append_unary(x.index, 1) per element (nd.a stays null = derivative 1)
    Node nd;
    nd.kind = Kind::UNARY;
    nd.base = t->len;
    nd.count = nd.n = n;
    nd.p0 = w->node;
    nd.p0_tracked = true;
    eml_tape::PendingSgd u;
    u.w = w->numbers;
    if (t->capturing) {
        // recorded step: the update happens in place, so a replay finds the weights where the last one left them
        u.out = w->numbers;
    } else {
        // eagerly: a new container (the nodes that captured the old numbers keep them)
        EML_TRY(new_buf(es * n, t->stream, &u.out));
    }
    if (t->dtype == EML_F32 && w->b_of_tc_product && t->fuse_packs) {
        if (!u.out->plane_id) u.out->plane_id = new_buffer_identity();
        EML_TRY(tf32x3_reserve_small_plane_b(u.out->plane_id, (const float*)u.out->p, w->rows, w->cols, t->stream, &u.small,
                                             &u.small_flag));
    }
    u.arena = d->arena;
    if (deferred) u.ws = d->reduce[w->node].ws, u.ws_splits = d->reduce[w->node].splits, u.ws_stride = d->reduce[w->node].stride;
    u.d = d->grads[w->node]->p;
    u.slot0 = (d->slot0_pending && w->node == 0) ? d->arena->p : nullptr;  // derivatives[0] + the parked sums, on the fly
    u.slot0_n = d->slot0_n;
    u.n = n;
    u.lr = lr;
    w->numbers = u.out;
    t->sgd.push_back(std::move(u));
    t->len += n;
    t->nodes.push_back(nd);
    w->node = (int)t->nodes.size() - 1;
    return EML_OK;
}

}  // extern "C"
