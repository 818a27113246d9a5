"""Edge cases of the conditional subtraction fast path (top limb equal to the modulus's top limb) and of the
error paths, through the CUDA library."""
import numpy as np
import pytest

from oracle import corac
from oracle import pyoracle as po
from tests.util import random_fr, raw_limbs

pytestmark = pytest.mark.gpu


def _eq(a, b):
    return a.shape == b.shape and bool((a == b).all())


def test_modular_reduction_near_the_modulus(handle):
    """Sums / products whose pre-reduction value has the same top 32-bit limb as r: the rare exact path of
    fr::cond_sub_r.  Inputs are chosen so that x + y and x - y land within 2^224 of r and of 2r."""
    R = po.R_MOD
    top = R >> 224
    vals = []
    for d in (0, 1, 2, (1 << 200) + 12345, (1 << 223) - 1):
        for base in (R - 1 - d, (top << 224) + d, (top << 224) - 1 - d if d < (top << 224) else 0):
            vals.append(base % R)
    xs = [v for v in vals for _ in vals]
    ys = [w for _ in vals for w in vals]
    x, y = raw_limbs(xs), raw_limbs(ys)          # raw limbs ARE the Montgomery-form elements here
    # field subtraction (sub_kernel) and addition through x - (0 - y)
    assert _eq(handle.zkvec_sub(x, y), corac.zkvec_sub(x, y))
    neg_y = corac.zkvec_sub(np.zeros_like(y), y)
    assert _eq(handle.zkvec_sub(x, neg_y), corac.zkvec_sub(x, neg_y))
    # running sums (add_fast + mont_mul_fast inside the mat-vec kernels), one pair per row and long rows
    n = len(xs)
    assert _eq(handle.zkvec_inner_prefix(x.reshape(1, n, 4), y.reshape(1, n, 4)), corac.zkvec_inner_prefix(x.reshape(1, n, 4), y.reshape(1, n, 4)))
    ones = np.broadcast_to(po.pack_mont([1]), (n, 4)).copy()
    assert _eq(handle.zkvec_inner_prefix(x.reshape(1, n, 4), ones.reshape(1, n, 4)), corac.zkvec_inner_prefix(x.reshape(1, n, 4), ones.reshape(1, n, 4)))
    # rescale / range-check witnesses of the same values (mont_mul_small + add_fast chains)
    q, wit = handle.rescale_witness(x, 42, 19)
    eq, _, ewit = corac.rescale_witness(x, 42, 19)
    assert _eq(q.reshape(-1, 4), eq) and _eq(wit, ewit)
    assert _eq(handle.range_check_witness(x, 93, 19), corac.range_check_witness(x, 93, 19))


def test_non_canonical_inputs_are_rejected(handle, pkg):
    rng = np.random.default_rng(0)
    a, b = random_fr(rng, 4, 4), random_fr(rng, 4, 4)
    bad = a.copy()
    bad[1, 2] = raw_limbs([po.R_MOD])[0]           # == r: not canonical
    for call in (lambda: handle.fr_matmul(bad, b), lambda: handle.rescale_witness(bad, 32, 19),
                 lambda: handle.freivalds_witness(bad, b, a, b[0, 0]),
                 lambda: handle.zkmatrix_mul_witness(bad, b, b[0, 0], 32, 19)):
        with pytest.raises(pkg.H2svdError) as ei:
            call()
        assert ei.value.code == -5                 # H2SVD_ERANGE
    assert _eq(handle.fr_matmul(a, b), corac.field_mat_mul(a, b))   # the handle stays usable


def test_empty_and_single_element_inputs(handle):
    rng = np.random.default_rng(1)
    e = np.zeros((0, 4), dtype=np.uint64)
    assert handle.zkvec_sub(e, e).shape == (0, 4)
    assert handle.rescale_witness(e, 32, 19)[1].shape[0] == 0
    assert handle.range_check_witness(e, 40, 19).shape[0] == 0
    x = random_fr(rng, 1)
    q, wit = handle.rescale_witness(x, 63, 19)
    eq, _, ewit = corac.rescale_witness(x, 63, 19)
    assert _eq(q.reshape(-1, 4), eq) and _eq(wit, ewit)


def test_two_handles_on_two_devices_in_one_process(pkg):
    """One process, one handle per GPU (the C++/Rust shim's multi-GPU mode): kernels that opt in to large dynamic
    shared memory must be configured on every device they run on."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(5)
    a, b = random_fr(rng, 40, 70), random_fr(rng, 70, 33)
    gamma = random_fr(rng, 1)
    want = corac.field_mat_mul(a, b)
    ew = corac.rescale_witness(want.reshape(-1, 4), 63, 19)[2]
    a2, b2 = random_fr(rng, 96, 80), random_fr(rng, 80, 72)     # large enough for the tensor-core mat-mul engine
    want2 = corac.field_mat_mul(a2, b2)
    for dev in (1, 0, 1):
        with pkg.Handle(dev) as h:
            res = h.zkmatrix_mul_witness(a, b, gamma, 63, 19)
            assert _eq(res["c_s"], want) and _eq(res["wit"], ew) and not res["diff"].any()
            assert _eq(h.fr_matmul(a2, b2), want2) and h.last_matmul_engine() == "tensor"
            x, s = random_fr(rng, 700, 300), random_fr(rng, 700, 300)
            assert _eq(h.zkvec_inner_prefix(x, s), corac.zkvec_inner_prefix(x, s))
