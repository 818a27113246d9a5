"""Row-sharded hot path (halo2-svd041_b200/workload.py) at world_size 2 on CPU: two processes, gloo backend,
the `*_dev` backend replaced by a stand-in that computes values with the C oracle on CPU tensors.  Checks
that the sharded witnesses, gathered, are byte-identical to the single-rank ones (field addition is exact;
the k row totals of B.v are either recomputed on every rank -- the default, no collective -- or all-gathered), including uneven splits (n, k not divisible
by the world size, which exercises the padded all-gather + compaction)."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import corac
from oracle import pyoracle as po

wl = importlib.import_module("halo2-svd041_b200.workload")


def _np(t):
    return np.ascontiguousarray(t.numpy()).view(np.uint64)


def _put(t, arr):
    t.copy_(torch.from_numpy(np.ascontiguousarray(arr).view(np.int64).reshape(t.shape)))


class OracleBackend:
    """CPU stand-in with the gpu.Handle `*_dev` method names (test infrastructure)."""

    def fr_matmul_dev(self, a, b, c, b_transposed=False):
        _put(c, corac.field_mat_mul(_np(a), _np(b)))

    def rescale_witness_dev(self, c_s, count, P, lb, out_q, out_wit, shift_bits=-1, a_num_bits=-1):
        q, _, w = corac.rescale_witness(_np(c_s).reshape(-1, 4)[:count], P, lb, shift_bits, a_num_bits)
        _put(out_q, q)
        _put(out_wit, w)

    def gamma_powers_dev(self, gamma, d, out):
        _put(out, corac.gamma_powers(_np(gamma).reshape(1, 4), d))

    def mat_vec_prefix_dev(self, a, v, out, totals=None):
        if a.shape[0] == 0:
            return
        pre = corac.mat_vec_prefix(_np(a), _np(v)[: a.shape[1]])
        _put(out, pre)
        if totals is not None:
            _put(totals, pre[:, -1])

    def mat_vec_totals_dev(self, a, v, totals):
        _put(totals, corac.mat_vec_prefix(_np(a), _np(v)[: a.shape[1]])[:, -1])

    def mat_vec_prefix_pair_dev(self, a0, out0, tot0, a1, out1, tot1, v):
        self.mat_vec_prefix_dev(a0, v, out0, tot0)
        self.mat_vec_prefix_dev(a1, v, out1, tot1)

    def is_equal_witness_dev(self, x, y, diff, is_zero, inv):
        d = corac.zkvec_sub(_np(x), _np(y))
        vals = po.unpack_mont(d)
        _put(diff, d)
        _put(is_zero, po.pack_mont([1 if v == 0 else 0 for v in vals]))
        _put(inv, po.pack_mont([1 if v == 0 else pow(v, -1, po.R_MOD) for v in vals]))


def _inputs(n, k, m, P):
    rng = np.random.default_rng(7)
    a = corac.quantize(rng.uniform(-5, 5, size=(n, k)), P)
    b = corac.quantize(rng.uniform(-5, 5, size=(k, m)), P)
    gamma = po.pack_mont([0xABCDEF0123456789ABCDEF % po.R_MOD])
    return a, b, gamma


def _run_rank(rank, world, port, n, k, m, P, lb, outdir, exchange="redundant"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    comm = None
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        comm = dist.group.WORLD
    plan = wl.ShardPlan(n, k, m, world, rank)
    W = corac.rescale_witness_count(P, lb)
    bufs = wl.alloc_buffers(torch, plan, W, torch.device("cpu"))
    a, b, gamma = _inputs(n, k, m, P)
    r0, r1 = plan.rows
    _put(bufs.a_slab, a[r0:r1])
    _put(bufs.b, b)
    _put(bufs.gamma, gamma)
    wl.run_step(OracleBackend(), plan, bufs, P, lb, dist if world > 1 else None, comm, exchange=exchange)
    b0, b1 = plan.brows
    np.savez(os.path.join(outdir, f"rank{rank}of{world}.npz"), c=_np(bufs.c_slab), q=_np(bufs.q_slab), wit=_np(bufs.wit_slab),
             powers=_np(bufs.powers), pcv=_np(bufs.prefix_cv), pbv=_np(bufs.prefix_bv)[: b1 - b0], pabv=_np(bufs.prefix_abv),
             diff=_np(bufs.diff), is_zero=_np(bufs.is_zero), inv=_np(bufs.inv))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("exchange", ["redundant", "allgather"])
@pytest.mark.parametrize("n,k,m", [(10, 7, 9), (8, 8, 8)])
def test_row_sharded_step_matches_single_rank_and_oracle(tmp_path, n, k, m, exchange):
    P, lb = 42, 19
    outdir = str(tmp_path)
    _run_rank(0, 1, _free_port(), n, k, m, P, lb, outdir)                       # single rank, in-process
    mp.spawn(_run_rank, args=(2, _free_port(), n, k, m, P, lb, outdir, exchange), nprocs=2, join=True)   # world_size 2, gloo
    one = np.load(os.path.join(outdir, "rank0of1.npz"))
    two = [np.load(os.path.join(outdir, f"rank{r}of2.npz")) for r in range(2)]
    for key in ("c", "q", "wit", "pcv", "pbv", "pabv", "diff", "is_zero", "inv"):
        cat = np.concatenate([t[key] for t in two], axis=0)
        assert cat.shape == one[key].shape and (cat == one[key]).all(), key
    assert (two[0]["powers"] == one["powers"]).all() and (two[1]["powers"] == one["powers"]).all()
    # and the single-rank result is the oracle's full witness
    a, b, gamma = _inputs(n, k, m, P)
    c = corac.field_mat_mul(a, b)
    fw = corac.freivalds_witness(a, b, c, gamma)
    assert (one["c"] == c).all() and (one["pcv"] == fw["prefix_cv"]).all() and (one["pbv"] == fw["prefix_bv"]).all()
    assert (one["pabv"] == fw["prefix_abv"]).all() and not one["diff"].any()


def test_split_range_covers_everything():
    for total in (1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            parts = [wl.split_range(total, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
