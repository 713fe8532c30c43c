"""Host logic of the row-sharded GEMM (easy-ml_b200/sharded.py) on CPU: world_size 2 and 3 over gloo.

The slab product is a stand-in (torch CPU matmul with an accumulate flag); what is tested here is the
partition, the panelled broadcast of B, the in-place all-gather and the ragged cases (M not divisible by
the world size, more ranks than rows) -- the parts that are identical on NCCL.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from easy_ml_b200.sharded import ShardedGemm, ShardPlan, k_panels, row_range, rows_per_rank


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, m, n, k, panels, out, gather="nccl"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(1234)
        a_full = torch.randint(-8, 9, (m, k), generator=g).double()
        b_full = torch.randint(-8, 9, (k, n), generator=g).double()
        plan = ShardPlan(m=m, n=n, k=k, world=world, rank=rank, panels=panels)
        h = plan.slab_rows
        first, last = plan.my_rows
        a_slab = torch.zeros((h, k), dtype=torch.float64)
        a_slab[:last - first] = a_full[first:last]
        b = torch.zeros((k, n), dtype=torch.float64)  # only the owner has the data before the step
        b_src = b_full.clone() if rank == 0 else None
        c_full = torch.full((world * h, n), -1.0, dtype=torch.float64)
        calls = []

        scatters = []

        def slab_gemm(a, bb, c, accumulate, scatter):
            calls.append((tuple(a.shape), tuple(bb.shape), accumulate))
            scatters.append(scatter)
            if accumulate:
                c += a @ bb
            else:
                c.copy_(a @ bb)

        flag = torch.zeros(1)
        op = ShardedGemm(plan, dist, slab_gemm, b_owner=0, gather=gather, flag=flag)
        c = op.step(a_slab, b, c_full, b_src)
        if gather == "fused":
            # on the GPU the last launch stores the rows into the peers; emulate those stores here, after
            # checking that exactly the last panel was asked to scatter and that no collective moved C
            assert scatters == [False] * (len(scatters) - 1) + [True] * (1 if scatters else 0)
            mine = c_full[rank * h:(rank + 1) * h].clone()
            pieces = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(pieces, mine)
            for r, piece in enumerate(pieces):
                if r != rank:
                    assert torch.all(c_full[r * h:(r + 1) * h] == -1.0)  # untouched by the step itself
                    c_full[r * h:(r + 1) * h] = piece
            c = c_full[:m]
        else:
            assert not any(scatters)
        ok = torch.equal(c, a_full @ b_full) and torch.equal(b, b_full)
        # first panel overwrites, later ones accumulate; one call per panel on ranks that own rows
        expect_calls = len(k_panels(k, panels)) if last > first else 0
        ok = ok and len(calls) == expect_calls and all(acc == (i > 0) for i, (_, _, acc) in enumerate(calls))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,m,n,k,panels", [(2, 64, 48, 40, 4), (2, 33, 20, 17, 3), (3, 10, 7, 9, 8), (3, 2, 5, 6, 2)])
def test_sharded_gemm_over_gloo(world, m, n, k, panels):
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, m, n, k, panels, out), nprocs=world, join=True)
        assert all(out.get(r) for r in range(world)), dict(out)


def test_sharded_gemm_fused_gather_protocol_over_gloo():
    """gather="fused": only the last K panel is asked to scatter, no collective moves C, one barrier."""
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, 24, 16, 20, 4, out, "fused"), nprocs=2, join=True)
        assert all(out.get(r) for r in range(2)), dict(out)


def test_partition_covers_rows_once():
    for m in (1, 2, 7, 64, 1000, 32768):
        for world in (1, 2, 3, 4, 8):
            h = rows_per_rank(m, world)
            ranges = [row_range(m, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and max(e for _, e in ranges) == m
            assert all(0 <= e - s <= h for s, e in ranges)
            covered = sum(e - s for s, e in ranges)
            assert covered == m
            for (s0, e0), (s1, e1) in zip(ranges, ranges[1:]):
                assert e0 == s1 or (s1 == m and e1 == m)


def test_k_panels_cover_k_once():
    for k in (1, 5, 17, 4096):
        for p in (1, 2, 8, 100):
            ps = k_panels(k, p)
            assert ps[0][0] == 0 and ps[-1][1] == k
            assert all(a[1] == b[0] for a, b in zip(ps, ps[1:]))
            assert all(b > a for a, b in ps)


def _batch_worker(rank, world, port, batch, gather, out):
    from easy_ml_b200.sharded import ShardedBatchContraction

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(99)
        x_full = torch.randint(-8, 9, (batch, 5, 7), generator=g).double()
        y_full = torch.randint(-8, 9, (batch, 7, 3), generator=g).double()
        calls = []

        def contract(x, y, o):
            calls.append(x.shape[0])
            o.copy_(torch.bmm(x, y))

        op = ShardedBatchContraction(batch, world, rank, dist, contract, gather=gather)
        first, last = op.my_batches
        h = op.slab
        x = torch.zeros((h, 5, 7), dtype=torch.float64)
        y = torch.zeros((h, 7, 3), dtype=torch.float64)
        x[:last - first] = x_full[first:last]
        y[:last - first] = y_full[first:last]
        out_full = torch.full((world * h, 5, 3), -1.0, dtype=torch.float64)
        got = op.step(x, y, out_full)
        want = torch.bmm(x_full, y_full)
        if gather == "nccl":
            ok = torch.equal(got, want)
        else:  # result stays sharded: this rank's batches only, nobody else's slot touched
            ok = torch.equal(got, want[first:last])
            others = torch.cat([out_full[:rank * h], out_full[(rank + 1) * h:]])
            ok = ok and bool(torch.all(others == -1.0))
        ok = ok and calls == ([last - first] if last > first else [])
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,batch,gather", [(2, 8, "nccl"), (2, 5, "nccl"), (3, 2, "nccl"), (2, 7, "none")])
def test_batch_sharded_contraction_over_gloo(world, batch, gather):
    """BASELINE config 3 across ranks: contiguous batch shards, no exchange while computing, optional gather;
    ragged (batch not divisible by the world size) and more ranks than batches."""
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_batch_worker, args=(world, port, batch, gather, out), nprocs=world, join=True)
        assert all(out.get(r) for r in range(world)), dict(out)
