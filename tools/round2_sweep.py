"""Tuning sweep on one B200: rescale burst sizes, mat-vec prefix kernels, mat-mul call -- isolated, L2 flushed, CUDA events."""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("halo2-svd041_b200")
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
h = pkg.Handle(0, stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gen = torch.Generator(device=dev)
gen.manual_seed(3)


def fr(*shape):
    return torch.zeros(shape + (4,), dtype=torch.int64, device=dev)


def timed(fn, reps=7):
    ts = []
    for i in range(reps + 2):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts)), float(min(ts))


def quant(rows, cols, P=63):
    x = (torch.rand((rows, cols), dtype=torch.float64, device=dev, generator=gen) - 0.5) * 4
    o = fr(rows, cols)
    h.quantize_dev(x, P, o)
    return o


n = 1024
a, b = quant(n, n), quant(n, n)
c = fr(n, n)
h.fr_matmul_dev(a, b, c)
h.sync()
print("mat-mul call (split + mm), quantized:", timed(lambda: h.fr_matmul_dev(a, b, c)), h.last_matmul_engine())
for rows in (128, 256, 512, 1024):
    ar, cr = a[:rows].contiguous(), fr(rows, n)
    res = []
    for width in (0, 8, 16, 24):
        h.tune("matmul_small_width", width)
        res.append((width, round(timed(lambda: h.fr_matmul_dev(ar, b, cr))[0], 1)))
    h.tune("matmul_small_width", 0)
    print(f"mat-mul call {rows}-row slab, quantized, by tile width (0 = cost model): {res}")
for kind, name in ((0, "copy"), (1, "write 16B streaming"), (2, "write 256B bulk from smem"), (3, "read")):
    print(f"HBM {name}: {h.microbench_hbm(kind):.0f} GB/s")
W = h.rescale_witness_count(63, 19)
q, wit = fr(n, n), fr(n * n, W)
ref = None
for store, name in ((0, "bulk copies"), (1, "TMA tensor stores"), (2, "coalesced STG.128")):
    h.tune("rescale_store", store)
    wit.fill_(-1)
    t = timed(lambda: h.rescale_witness_dev(c, n * n, 63, 19, q, wit))
    if ref is None:
        ref = wit.clone()
    same = bool((wit == ref).all().item())
    print(f"rescale {name}: median {t[0]:.1f} us  min {t[1]:.1f} us  -> {n * n * 32 * (2 + W) / t[0] / 1e3:.0f} GB/s  same_bytes={same}")
h.tune("rescale_store", 0)
for ch in (16, 12, 8):
    h.tune("rescale_ch", ch)
    wit.fill_(-1)
    t = timed(lambda: h.rescale_witness_dev(c, n * n, 63, 19, q, wit))
    print(f"rescale bulk copies, {ch} witnesses per burst: median {t[0]:.1f} us  same_bytes={bool((wit == ref).all().item())}")
# one rank's share of the 8-GPU job: 128 x 1024 elements
cnt = 128 * n
for store, ch in ((0, 8), (0, 6), (0, 4), (0, 12), (0, 16), (2, 8)):
    h.tune("rescale_store", store)
    h.tune("rescale_ch", ch)
    t = timed(lambda: h.rescale_witness_dev(c, cnt, 63, 19, q, wit))
    print(f"rescale of a 128-row slab, store={store} burst={ch}: median {t[0]:.1f} us")
h.tune("rescale_store", 0)
h.tune("rescale_ch", 8)
del ref
powers = fr(n)
g = torch.tensor([[0x1234567, 0x89ABCDEF, 0x13579BDF, 0x2468ACE]], dtype=torch.int64, device=dev)
h.gamma_powers_dev(g, n, powers)
out, tot = fr(n, n), fr(n)
for rows in (1024, 128):
    base = None
    for seg, x2 in ((0, 0), (1, 0), (0, 1), (1, 1)):
        h.tune("matvec_seg", seg)
        h.tune("matvec_x2", x2)
        o = out[:rows]
        o.fill_(-1)
        t = timed(lambda: h.mat_vec_prefix_dev(c[:rows], powers, o, tot[:rows]))
        if base is None:
            base = o.clone()
        print(f"mat_vec_prefix {rows}x1024 seg={seg} x2={x2}: median {t[0]:.1f} us min {t[1]:.1f} us same={bool((o == base).all().item())}")
x, s, o2 = quant(4096, 1024, 32), quant(4096, 1024, 32), fr(4096, 1024)
base = None
for seg, x2 in ((0, 0), (1, 0), (0, 1), (1, 1)):
    h.tune("matvec_seg", seg)
    h.tune("matvec_x2", x2)
    o2.fill_(-1)
    t = timed(lambda: h.zkvec_inner_prefix_dev(x, s, o2))
    if base is None:
        base = o2.clone()
    print(f"zkvec inner 4096x1024 seg={seg} x2={x2}: median {t[0]:.1f} us min {t[1]:.1f} us -> {4096 * 1024 * 96 / t[0] / 1e3:.0f} GB/s same={bool((o2 == base).all().item())}")
h.tune("matvec_seg", -1)
h.tune("matvec_x2", 0)
tt = fr(n)
print("mat_vec_totals 1024x1024:", timed(lambda: h.mat_vec_totals_dev(b, powers, tt)))
print("gamma_powers 1024:", timed(lambda: h.gamma_powers_dev(g, n, powers)))
h.close()
