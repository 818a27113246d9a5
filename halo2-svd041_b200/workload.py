"""The hot path as one "step": honest_prover_mat_mul -> rescale_matrix -> verify_mul witnesses for
one (A, B) pair, row-sharded over `world` GPUs (one process per GPU), spelled out in building-block calls.

Sharding: rank g owns rows [r0, r1) of A and of C (mat-mul, rescale, C.v and A.(Bv)
witnesses are row-local); B is replicated; the k rows of B are split for the B.v running-sum witnesses.
Every rank needs all k row totals (B v) as the second operand of A.(Bv); two ways to get them:
  exchange="redundant" (default, what h2svd_zkmatrix_mul_witness[_dev] and bench.py do): every rank
      recomputes the totals itself (h2svd_mat_vec_totals_dev: lazily accumulated, no running sums) --
      NO collective on the data path; measured in round 1, the all-gather below cost a 66 us tail at 8 ranks;
  exchange="allgather": the totals of the local rows are exchanged with ONE all-gather (NCCL over NVLink).
Field addition is exact, so either way the sharded witnesses are byte-identical to the single-GPU ones.
The C library's own h2svd_zkmatrix_mul_witness_dev runs the same schedule inside one call (and is what
bench.py records into a CUDA graph); this module is the readable, testable statement of the sharding.

`backend` is any object with the `*_dev` methods of gpu.Handle (the tests inject a CPU stand-in to
cover the sharding logic under gloo); `comm` is a torch.distributed process group or None.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple


def split_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous near-equal split: the first (total % world) ranks get one extra row."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


@dataclass(frozen=True)
class ShardPlan:
    n: int
    k: int
    m: int
    world: int = 1
    rank: int = 0

    @property
    def rows(self) -> Tuple[int, int]:      # rows of A / C owned by this rank
        return split_range(self.n, self.world, self.rank)

    @property
    def brows(self) -> Tuple[int, int]:     # rows of B whose B.v running sums this rank produces
        return split_range(self.k, self.world, self.rank)

    @property
    def kmax(self) -> int:                  # padded per-rank row count of the all-gather
        return -(-self.k // self.world)

    def mul_adds(self) -> int:
        """Fr multiply-adds of the whole job (all ranks): mat-mul + Freivalds."""
        return self.n * self.k * self.m + (self.m - 1) + self.n * self.m + self.k * self.m + self.n * self.k


@dataclass
class StepBuffers:
    """Device tensors of one rank (int64[..., 4] == bn256::Fr limbs).  Inputs: a_slab, b, gamma."""
    a_slab: object
    b: object
    gamma: object
    c_slab: object
    q_slab: object
    wit_slab: object
    powers: object
    prefix_cv: object
    prefix_bv: object      # local rows of B only
    prefix_abv: object
    bv_local: object       # [kmax, 4] zero-padded row totals of the local B rows
    bv_all: object         # [world * kmax, 4] all-gather target
    bv: object             # [k, 4] compacted (B v)
    csv: object
    abv: object
    diff: object
    is_zero: object
    inv: object
    bv_index: Optional[object] = None   # compaction gather indices when k % world != 0


def alloc_buffers(torch, plan: ShardPlan, W: int, device) -> StepBuffers:
    r0, r1 = plan.rows
    b0, b1 = plan.brows
    rows, brows = r1 - r0, b1 - b0

    def fr(*shape):
        return torch.zeros(shape + (4,), dtype=torch.int64, device=device)

    idx = None
    if plan.world > 1 and plan.k % plan.world != 0:
        pos = []
        for g in range(plan.world):
            lo, hi = split_range(plan.k, plan.world, g)
            pos += [g * plan.kmax + i for i in range(hi - lo)]
        idx = torch.tensor(pos, dtype=torch.int64, device=device)
    return StepBuffers(
        a_slab=fr(rows, plan.k), b=fr(plan.k, plan.m), gamma=fr(1), c_slab=fr(rows, plan.m),
        q_slab=fr(rows, plan.m), wit_slab=fr(rows * plan.m, W), powers=fr(plan.m),
        prefix_cv=fr(rows, plan.m), prefix_bv=fr(max(brows, 1), plan.m), prefix_abv=fr(rows, plan.k),
        bv_local=fr(plan.kmax), bv_all=fr(plan.world * plan.kmax), bv=fr(plan.k), csv=fr(rows), abv=fr(rows),
        diff=fr(rows), is_zero=fr(rows), inv=fr(rows), bv_index=idx)


def step_matmul(backend, plan: ShardPlan, bufs: StepBuffers) -> None:
    """honest_prover_mat_mul (reference src/matrix/mod.rs:546): this rank's rows of C = A.B"""
    backend.fr_matmul_dev(bufs.a_slab, bufs.b, bufs.c_slab)


def step_rescale(backend, plan: ShardPlan, bufs: StepBuffers, precision_bits: int, lookup_bits: int) -> None:
    """ZkMatrix::rescale_matrix (:354) witnesses for this rank's rows of C"""
    r0, r1 = plan.rows
    backend.rescale_witness_dev(bufs.c_slab, (r1 - r0) * plan.m, precision_bits, lookup_bits, bufs.q_slab,
                                bufs.wit_slab)


def _bv_running_sums(backend, plan: ShardPlan, bufs: StepBuffers, dist=None, comm=None, exchange: str = "redundant"):
    """gamma powers (reference src/matrix/mod.rs:316-326), the B.v running sums of this rank's rows of B (:336) and
    the k row totals (B v): recomputed on every rank, or all-gathered.  Returns the tensor holding (B v)."""
    b0, b1 = plan.brows
    brows = b1 - b0
    backend.gamma_powers_dev(bufs.gamma, plan.m, bufs.powers)
    if brows > 0:
        backend.mat_vec_prefix_dev(bufs.b[b0:b1], bufs.powers, bufs.prefix_bv[:brows], bufs.bv_local[:brows])
    if plan.world > 1 and exchange == "redundant":
        backend.mat_vec_totals_dev(bufs.b, bufs.powers, bufs.bv)                      # all k totals, no exchange step
        return bufs.bv
    if plan.world > 1:
        dist.all_gather_into_tensor(bufs.bv_all, bufs.bv_local, group=comm)          # the one exchange step
        if bufs.bv_index is not None:
            bufs.bv.copy_(bufs.bv_all.index_select(0, bufs.bv_index))
            return bufs.bv
        return bufs.bv_all
    return bufs.bv_local


def step_matmul_rescale(backend, plan: ShardPlan, bufs: StepBuffers, precision_bits: int, lookup_bits: int) -> None:
    """honest_prover_mat_mul (:546) + rescale_matrix (:354) of this rank's rows of C through the one-call entry point
    h2svd_fr_matmul_rescale_dev (two kernels back to back, or the experimental fused launch when the tuning hook is on)."""
    fused = getattr(backend, "fr_matmul_rescale_dev", None)
    if fused is None:   # CPU stand-in of the sharding tests
        step_matmul(backend, plan, bufs)
        step_rescale(backend, plan, bufs, precision_bits, lookup_bits)
        return
    fused(bufs.a_slab, bufs.b, bufs.c_slab, precision_bits, lookup_bits, bufs.q_slab, bufs.wit_slab)


def step_freivalds_pre(backend, plan: ShardPlan, bufs: StepBuffers, dist=None, comm=None, exchange: str = "redundant"):
    """The part of ZkMatrix::verify_mul (:299) that does not need C: gamma powers (:316-326), the B.v
    running sums of this rank's rows of B (:336), the one exchange step (the all-gather of the k row
    totals) and A.(Bv) for this rank's rows (:337).  Returns the tensor holding (B v)."""
    bv = _bv_running_sums(backend, plan, bufs, dist, comm, exchange)
    backend.mat_vec_prefix_dev(bufs.a_slab, bv, bufs.prefix_abv, bufs.abv)
    return bv


def step_freivalds_post(backend, plan: ShardPlan, bufs: StepBuffers, bv) -> None:
    """The part of verify_mul that needs C: C.v (:335) for this rank's rows, is_equal (:339-341)."""
    backend.mat_vec_prefix_dev(bufs.c_slab, bufs.powers, bufs.prefix_cv, bufs.csv)
    backend.is_equal_witness_dev(bufs.csv, bufs.abv, bufs.diff, bufs.is_zero, bufs.inv)


class SideStream:
    """A second (backend, CUDA stream) pair on the same GPU: the C-independent part of verify_mul (two of its three
    mat-vecs: integer-pipe work) runs there under the mat-mul (tensor-pipe work), C.v next to the rescale (HBM-bound)."""

    def __init__(self, torch, backend, stream, main_stream) -> None:
        self.torch, self.backend, self.stream, self.main = torch, backend, stream, main_stream
        self.ev_fork = torch.cuda.Event()
        self.ev_join = torch.cuda.Event()

    def run(self, fn):
        self.ev_fork.record(self.main)           # inputs (a_slab, b, gamma) are ready on the main stream
        self.stream.wait_event(self.ev_fork)
        with self.torch.cuda.stream(self.stream):
            out = fn(self.backend)
            self.ev_join.record(self.stream)
        return out

    def join(self) -> None:
        self.main.wait_event(self.ev_join)


def step_freivalds(backend, plan: ShardPlan, bufs: StepBuffers, dist=None, comm=None, exchange: str = "redundant") -> None:
    """ZkMatrix::verify_mul (:299) witnesses, row-sharded."""
    bv = step_freivalds_pre(backend, plan, bufs, dist, comm, exchange)
    step_freivalds_post(backend, plan, bufs, bv)


def run_step(backend, plan: ShardPlan, bufs: StepBuffers, precision_bits: int, lookup_bits: int, dist=None,
             comm=None, side: Optional[SideStream] = None, exchange: str = "redundant") -> None:
    if side is None:
        step_matmul(backend, plan, bufs)
        step_rescale(backend, plan, bufs, precision_bits, lookup_bits)
        step_freivalds(backend, plan, bufs, dist, comm, exchange)
        return
    bv = side.run(lambda be: step_freivalds_pre(be, plan, bufs, dist, comm, exchange))
    step_matmul(backend, plan, bufs)
    side.run(lambda be: step_freivalds_post(be, plan, bufs, bv))   # stream order on the side stream: after the pre part
    step_rescale(backend, plan, bufs, precision_bits, lookup_bits)
    side.join()
