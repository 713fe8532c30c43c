"""Row-block sharded GEMM over the GPUs of one box (BASELINE config 5).

Rows of C = A * B are independent units (matrix_view_multiplication computes each C[i, j] from row i of A
and column j of B only, reference src/matrices/operations.rs:187-194), so A and C are partitioned by
contiguous row blocks, B is replicated:

    rank r owns rows [row_range(M, world, r))            no cross-rank reduction -> same bits as one GPU
    B travels from its owner in K panels (broadcast)     panel p+1 is in flight while panel p is multiplied
    C slabs are delivered either by one NCCL all-gather (in place: every slab already sits at its offset;
    the correctness baseline) or -- gather="fused" -- by the GEMM kernel itself: the launch of the last K
    panel stores every finished tile into all peers' copies of C over NVLink (peer-mapped pointers), so no
    separate collective runs; a stream-ordered 1-element all-reduce is the only synchronisation

One process per GPU; `torch.distributed` (NCCL on GPUs, gloo on CPU for the host-logic tests) is the
plumbing.  The slab product itself is injected (`slab_gemm`): on the GPU it is the library's
eml_contract_f64_dev with `accumulate`, which continues one k-ordered accumulation chain per element
across panels, so the panelled product is bit-identical to a single launch.
"""
from dataclasses import dataclass


def rows_per_rank(m, world):
    """Slab height: ceil(M / world).  Every rank holds a slab of this height (the last ones may be partly or
    wholly padding), so the all-gather moves equal-sized pieces."""
    return (m + world - 1) // world


def row_range(m, world, rank):
    """[first, last) rows of C (and A) owned by `rank`; empty when the rank is past the end."""
    h = rows_per_rank(m, world)
    first = min(m, rank * h)
    return first, min(m, first + h)


def k_panels(k, panels):
    """[first, last) k ranges of the broadcast panels.

    Geometric schedule: the first two panels are k / 2^(panels-1) deep and every later one doubles, so the only
    exposed communication is the arrival of a thin first panel, while most of the work runs in a few deep
    launches (each launch re-reads C once and pays one prologue/epilogue per tile).  Boundaries are multiples
    of 32 (the kernels' k tile) whenever k allows it."""
    panels = max(1, min(panels, k))
    if panels == 1:
        return [(0, k)]
    unit = max(1, k >> (panels - 1))
    if k >= 64:
        unit = max(32, unit // 32 * 32)
    edges, depth = [0], unit
    while edges[-1] + depth < k and len(edges) < panels:
        edges.append(edges[-1] + depth)
        if len(edges) > 2:
            depth *= 2
    edges.append(k)
    return [(a, b) for a, b in zip(edges, edges[1:]) if b > a]


@dataclass
class ShardPlan:
    m: int
    n: int
    k: int
    world: int
    rank: int
    panels: int

    @property
    def slab_rows(self):
        return rows_per_rank(self.m, self.world)

    @property
    def my_rows(self):
        return row_range(self.m, self.world, self.rank)

    def comm_bytes(self, itemsize):
        """Bytes this rank receives per step: broadcast of B (non-owners) + the other ranks' C slabs."""
        bcast = self.k * self.n * itemsize
        gather = (self.world - 1) * self.slab_rows * self.n * itemsize
        return {"broadcast_in": bcast, "allgather_in": gather}


class ShardedGemm:
    """C = A * B with A row-sharded.  Buffers are torch tensors on this rank's device.

    a_slab : [slab_rows, K]  this rank's rows of A (rows past M are ignored)
    b      : [K, N]          replicated destination of the broadcast (b_src on the owner holds the data)
    c_full : [world * slab_rows, N]  every rank ends up with all of C in c_full[:M]
    slab_gemm(a_view, b_view, c_view, accumulate, scatter) computes c_view (+)= a_view @ b_view on the device;
    scatter=True (only ever on the last K panel, only with gather="fused") asks it to also deliver the
    finished rows to every peer.
    """

    def __init__(self, plan, dist, slab_gemm, b_owner=0, gather="nccl", flag=None, pull_b=None):
        assert gather in ("nccl", "fused")
        self.plan, self.dist, self.slab_gemm, self.b_owner = plan, dist, slab_gemm, b_owner
        self.gather, self.flag = gather, flag  # flag: 1-element tensor on this rank's device (fused barrier)
        # pull_b(k0, k1) -> object with .wait(): fetches rows [k0, k1) of B from the owner into this rank's b
        # without NCCL (copy engines over NVLink on a side stream: no SMs are taken from the GEMM).
        # None = dist.broadcast.
        self.pull_b = pull_b

    def step(self, a_slab, b, c_full, b_src=None):
        p, dist = self.plan, self.dist
        first, last = p.my_rows
        rows = last - first
        h = p.slab_rows
        c_slab = c_full[p.rank * h:(p.rank + 1) * h]
        panels = k_panels(p.k, p.panels)
        works = []
        for (k0, k1) in panels:
            blk = b[k0:k1]
            if p.rank == self.b_owner and b_src is not None and b_src.data_ptr() != b.data_ptr():
                blk.copy_(b_src[k0:k1], non_blocking=True)
            if p.world == 1:
                works.append(None)
            elif self.pull_b is not None:
                works.append(self.pull_b(k0, k1) if p.rank != self.b_owner else None)
            else:
                works.append(dist.broadcast(blk, src=self.b_owner, async_op=True))
        for i, (k0, k1) in enumerate(panels):
            if works[i] is not None:
                works[i].wait()
            if rows > 0:
                scatter = self.gather == "fused" and p.world > 1 and i == len(panels) - 1
                self.slab_gemm(a_slab[:rows, k0:k1], b[k0:k1], c_slab[:rows], i > 0, scatter)
        if p.world > 1:
            if self.gather == "fused":
                dist.all_reduce(self.flag)  # stream-ordered barrier: every peer's stores have been issued and drained
            else:
                dist.all_gather_into_tensor(c_full, c_slab)  # in place: the slab is a view into c_full
        return c_full[:p.m]


class ShardedBatchContraction:
    """out[b] = x[b] * y[b] with the batch dimension sharded (BASELINE config 3 over several GPUs, SURVEY.md
    section 8e: "the natural split is the batch dimension, no collective but the gather").

    Einsum2::to computes every output element from one batch index of each input
    (src/tensors/einsum.rs:908-980), so batches are independent units: rank r owns the contiguous batches
    row_range(batch, world, r), both operands are generated / resident per shard, and nothing is exchanged while
    computing.  gather="none" leaves the result sharded (what a caller that keeps working per shard wants);
    gather="nccl" delivers every shard to every rank with one in-place all-gather.

    x_shard : [slab, M, K], y_shard : [slab, K, N]   this rank's batches (entries past the end are ignored)
    out_full: [world * slab, M, N]                   rank r's results land in out_full[r * slab : (r + 1) * slab]
    contract(x_view, y_view, out_view) computes the batched product of the views on the device.
    """

    def __init__(self, batch, world, rank, dist, contract, gather="none"):
        assert gather in ("none", "nccl")
        self.batch, self.world, self.rank, self.dist, self.contract, self.gather = batch, world, rank, dist, contract, gather

    @property
    def slab(self):
        return rows_per_rank(self.batch, self.world)

    @property
    def my_batches(self):
        return row_range(self.batch, self.world, self.rank)

    def step(self, x_shard, y_shard, out_full):
        first, last = self.my_batches
        n = last - first
        h = self.slab
        mine = out_full[self.rank * h:(self.rank + 1) * h]
        if n > 0:
            self.contract(x_shard[:n], y_shard[:n], mine[:n])
        if self.world > 1 and self.gather == "nccl":
            self.dist.all_gather_into_tensor(out_full, mine)  # in place: the shard is a view into out_full
        return out_full[:self.batch] if self.gather == "nccl" or self.world == 1 else mine[:n]
