"""Row-block sharded GEMM over the GPUs of one box (BASELINE config 5).

Rows of C = A * B are independent units (matrix_view_multiplication computes each C[i, j] from row i of A
and column j of B only, reference src/matrices/operations.rs:187-194), so A and C are partitioned by
contiguous row blocks, B is replicated:

    rank r owns rows [row_range(M, world, r))            no cross-rank reduction -> same bits as one GPU
    B travels from its owner in K panels (broadcast)     panel p+1 is in flight while panel p is multiplied
    C slabs are delivered with one all-gather             in place: every slab already sits at its offset

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
    """[first, last) k ranges of the broadcast panels (the last one takes the remainder)."""
    panels = max(1, min(panels, k))
    step = (k + panels - 1) // panels
    return [(p * step, min(k, (p + 1) * step)) for p in range(panels) if p * step < k]


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
    slab_gemm(a_view, b_view, c_view, accumulate) computes c_view (+)= a_view @ b_view on the device.
    """

    def __init__(self, plan, dist, slab_gemm, b_owner=0):
        self.plan, self.dist, self.slab_gemm, self.b_owner = plan, dist, slab_gemm, b_owner

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
            works.append(dist.broadcast(blk, src=self.b_owner, async_op=True) if p.world > 1 else None)
        for i, (k0, k1) in enumerate(panels):
            if works[i] is not None:
                works[i].wait()
            if rows > 0:
                self.slab_gemm(a_slab[:rows, k0:k1], b[k0:k1], c_slab[:rows], i > 0)
        if p.world > 1:
            dist.all_gather_into_tensor(c_full, c_slab)  # in place: the slab is a view into c_full
        return c_full[:p.m]
