// Bulk hand-off of GPU witnesses into halo2-base: host-side expansion of the Witness-only arrays the
// kernels return into the COMPLETE ordered advice-cell stream of the reference's operations, plus the static per-unit
// structure (cell kinds, selector offsets, lookup cells, copy constraints, constants) that goes with it.
//
// The reference pushes ~100 cells per rescaled element one `assign_region` at a time (src/matrix/mod.rs:364-373 ->
// FixedPointChip041::signed_div_scale -> RangeChip), each push building QuantumCell vectors on the heap.  Every unit
// (element / row) of one operation has the SAME structure, so the structure is built once (`h2svd_cells_layout`) and the
// values of all units are produced in one pass of 32-byte copies (`h2svd_expand_cells`, memcpy speed, multi-threaded).
// The Rust side then appends the value stream to `Context::advice` with one `extend` and replays the layout's offsets
// unit by unit for selectors / lookups / copy constraints; `raw_synthesize_phase0/1`
// (src/utils/executor.rs:100,116) see exactly the cells the per-cell path would have produced.
//
// Host code only (no kernels): compiled with the rest of the library so that it ships in the same .so.
#include <string.h>

#include <functional>
#include <new>
#include <thread>
#include <vector>

#include "common.cuh"

namespace h2svd {
int rescale_params(int P, int lb, int S, int A, int* n_d, int* n_r);  // rescale.cu
}

struct h2svd_cells_layout_impl {
    h2svd_cells_layout pub;   // must stay the first member: the public pointer IS this object
    std::vector<uint8_t> kind;
    std::vector<int32_t> source;
    std::vector<uint32_t> gates;
    std::vector<int32_t> lookups;
    std::vector<int32_t> copies;
    std::vector<h2svd_fr> constants;
};

namespace h2svd {
namespace {

using fr::Fr;

h2svd_fr wire(const Fr& f) {
    h2svd_fr o;
    for (int i = 0; i < 4; i++) o.l[i] = (uint64_t)f.l[2 * i] | ((uint64_t)f.l[2 * i + 1] << 32);
    return o;
}
Fr int_one() {
    Fr o = fr::zero();
    o.l[0] = 1u;
    return o;
}

// Records the cells of one unit in assignment order (the halo2-base 0.4.1 layouts of SURVEY.md A.2 / A.4 / A.5).
struct Builder {
    h2svd_cells_layout_impl* L;
    uint32_t nwit = 0, ninputs = 0;
    int here() const { return (int)L->kind.size(); }
    int W() {  // QuantumCell::Witness: next value of the unit's witness stripe
        L->kind.push_back(H2SVD_CELL_WITNESS);
        L->source.push_back((int32_t)nwit++);
        return here() - 1;
    }
    int Wagain(int cell) {  // a Witness cell carrying the same value as an earlier Witness cell (no copy constraint)
        L->kind.push_back(H2SVD_CELL_WITNESS);
        L->source.push_back(L->source[cell]);
        return here() - 1;
    }
    int C(const Fr& canonical_int) {  // QuantumCell::Constant
        const h2svd_fr v = wire(fr::to_mont(canonical_int));
        int idx = -1;
        for (size_t i = 0; i < L->constants.size(); i++)
            if (memcmp(&L->constants[i], &v, sizeof v) == 0) idx = (int)i;
        if (idx < 0) {
            L->constants.push_back(v);
            idx = (int)L->constants.size() - 1;
        }
        L->kind.push_back(H2SVD_CELL_CONSTANT);
        L->source.push_back(idx);
        return here() - 1;
    }
    int E(int src) {  // QuantumCell::Existing: an earlier cell of the unit (>= 0) or input i of the unit (-1 - i)
        L->kind.push_back(H2SVD_CELL_EXISTING);
        L->source.push_back(src);
        if (src < 0 && (uint32_t)(-src) > ninputs) ninputs = (uint32_t)(-src);
        return here() - 1;
    }
    void gate(int off) { L->gates.push_back((uint32_t)off); }
    void lookup(int cell) {
        L->lookups.push_back(cell);
        if (cell < 0 && (uint32_t)(-cell) > ninputs) ninputs = (uint32_t)(-cell);
    }
    void copy(int a, int b) {
        L->copies.push_back(a);
        L->copies.push_back(b);
    }
    // -Fr as a canonical integer
    static Fr neg(const Fr& x) {
        Fr m, o;
        for (int i = 0; i < 8; i++) m.l[i] = fr::modulus(i);
        fr::sub_n<8>(o.l, m.l, x.l);
        return o;
    }

    // RangeChip::range_check(a, n * lb): witnesses l0, l1, s1, l2, s2, ...
    void range_check(int a, int n, int lb) {
        if (n == 1) {
            lookup(a);
            return;
        }
        const int row = here();
        W();
        for (int i = 1; i < n; i++) {
            W();
            C(fr::pow2(lb * i));
            W();
            gate(row + 3 * (i - 1));
        }
        copy(a, here() - 1);
        lookup(row);
        for (int i = 0; i < n - 1; i++) lookup(row + 1 + 3 * i);
    }
    // RangeChip::range_check(a, range_bits) for any width (last limb handling)
    void range_check_bits(int a, int range_bits, int lb) {
        const int n = (range_bits + lb - 1) / lb, rem = range_bits % lb;
        int last = a;
        if (n == 1) {
            lookup(a);
        } else {
            const int row = here();
            range_check(a, n, lb);
            last = row + 1 + 3 * (n - 2);
        }
        if (rem == 1) {
            const int row = here();
            C(fr::zero());
            E(last);
            E(last);
            E(last);
            gate(row);
        } else if (rem > 1) {  // gate.mul(last, 2^(lb - rem)) -> looked up
            const int row = here();
            C(fr::zero());
            E(last);
            C(fr::pow2(lb - rem));
            W();
            gate(row);
            lookup(here() - 1);
        }
    }
    // RangeChip::check_big_less_than_safe(a, bound), n = ceil(bound.bits() / lb)
    void cbls(int a, const Fr& bound, int n, int lb) {
        range_check(a, n, lb);
        const int row = here();
        W();                                 // a + 2^bits - bound
        C(bound);
        C(int_one());
        W();                                 // a + 2^bits
        C(neg(fr::pow2(n * lb)));
        C(int_one());
        E(a);
        gate(row);
        gate(row + 3);
        range_check(row, n, lb);
    }
};

int bit_length(const Fr& x) {
    for (int i = 7; i >= 0; i--)
        if (x.l[i]) return 32 * i + (32 - __builtin_clz(x.l[i]));
    return 0;
}

h2svd_cells_layout* finish(h2svd_cells_layout_impl* L, const Builder& b) {
    L->pub.cells = (uint32_t)L->kind.size();
    L->pub.witnesses = b.nwit;
    L->pub.inputs = b.ninputs;
    L->pub.n_gates = (uint32_t)L->gates.size();
    L->pub.n_lookups = (uint32_t)L->lookups.size();
    L->pub.n_copies = (uint32_t)(L->copies.size() / 2);
    L->pub.n_constants = (uint32_t)L->constants.size();
    L->pub.kind = L->kind.data();
    L->pub.source = L->source.data();
    L->pub.gates = L->gates.data();
    L->pub.lookups = L->lookups.data();
    L->pub.copies = L->copies.data();
    L->pub.constants = L->constants.data();
    return &L->pub;
}

}  // namespace
}  // namespace h2svd

using namespace h2svd;

#define REQ(cond, msg)                  \
    do {                                \
        if (!(cond)) {                  \
            h2svd::set_error("%s", msg); \
            return H2SVD_EINVAL;        \
        }                               \
    } while (0)

extern "C" {

// FixedPointChip041::signed_div_scale(a) (reference src/matrix/mod.rs:369, :104), input 0 = a
int h2svd_rescale_cells_layout(int precision_bits, int lookup_bits, int shift_bits, int a_num_bits, h2svd_cells_layout** out) {
    REQ(out, "rescale_cells_layout: out is null");
    *out = nullptr;
    int n_d = 0, n_r = 0;
    const int P = precision_bits, lb = lookup_bits;
    const int S = shift_bits < 0 ? 3 * P : shift_bits, A = a_num_bits < 0 ? 4 * P : a_num_bits;
    const int W = rescale_params(P, lb, S, A, &n_d, &n_r);
    REQ(W > 0, "rescale_cells_layout: parameters out of range");
    auto* L = new (std::nothrow) h2svd_cells_layout_impl();
    if (!L) return H2SVD_ENOMEM;
    Builder b{L};
    // gate.add(a, Constant(2^S)): [a, 2^S, 1, a + 2^S]
    int row = b.here();
    b.E(-1);
    b.C(fr::pow2(S));
    b.C(int_one());
    const int a_shift = b.W();
    b.gate(row);
    // range.div_mod(a_shift, 2^P, A): [rem, 2^P, div, a_shift]
    row = b.here();
    const int rem = b.W();
    b.C(fr::pow2(P));
    const int div = b.W();
    b.E(a_shift);
    b.gate(row);
    Fr bound_d = fr::pow2(A - P);
    bound_d.l[0] |= 1u;                      // 2^A / 2^P + 1
    b.cbls(div, bound_d, n_d, lb);
    b.cbls(rem, fr::pow2(P), n_r, lb);
    // gate.sub(div, Constant(2^(S-P))): [q, 2^(S-P), 1, div]
    row = b.here();
    b.W();
    b.C(fr::pow2(S - P));
    b.C(int_one());
    b.E(div);
    b.gate(row);
    *out = finish(L, b);
    REQ((int)(*out)->witnesses == W, "rescale_cells_layout: internal error (witness count)");
    return H2SVD_OK;
}

// check_abs_less_than(x [- y], bnd) (reference src/matrix/mod.rs:425-459), input 0 = x, input 1 = y (with_diff)
int h2svd_abs_less_than_cells_layout(const uint64_t bnd[4], int lookup_bits, int with_diff, h2svd_cells_layout** out) {
    REQ(out && bnd, "abs_less_than_cells_layout: null argument");
    *out = nullptr;
    Fr bv;
    for (int i = 0; i < 4; i++) {
        bv.l[2 * i] = (uint32_t)bnd[i];
        bv.l[2 * i + 1] = (uint32_t)(bnd[i] >> 32);
    }
    int n = 0;
    Fr bound;
    const int W = abs_less_than_params(bv, lookup_bits, with_diff, &n, &bound);
    REQ(W > 0, "abs_less_than_cells_layout: parameters out of range");
    auto* L = new (std::nothrow) h2svd_cells_layout_impl();
    if (!L) return H2SVD_ENOMEM;
    Builder b{L};
    int d = -1;
    if (with_diff) {  // gate.sub(x, y): [x - y, y, 1, x]
        const int row = b.here();
        d = b.W();
        b.E(-2);
        b.C(int_one());
        b.E(-1);
        b.gate(row);
    }
    // gate.add(d, Constant(bnd - 1)): [d, bnd - 1, 1, t]
    Fr bm1;
    const Fr one = int_one();
    fr::sub_n<8>(bm1.l, bv.l, one.l);
    const int row = b.here();
    b.E(d);
    b.C(bm1);
    b.C(one);
    const int t = b.W();
    b.gate(row);
    b.cbls(t, bound, n, lookup_bits);
    *out = finish(L, b);
    REQ((int)(*out)->witnesses == W, "abs_less_than_cells_layout: internal error (witness count)");
    return H2SVD_OK;
}

// RangeChip::range_check(x, range_bits) (ZkVector::entries_less_than, src/matrix/mod.rs:185-197), input 0 = x
int h2svd_range_check_cells_layout(int range_bits, int lookup_bits, h2svd_cells_layout** out) {
    REQ(out, "range_check_cells_layout: out is null");
    *out = nullptr;
    const int W = range_check_params(range_bits, lookup_bits, nullptr, nullptr);
    REQ(W >= 0, "range_check_cells_layout: parameters out of range");
    auto* L = new (std::nothrow) h2svd_cells_layout_impl();
    if (!L) return H2SVD_ENOMEM;
    Builder b{L};
    b.range_check_bits(-1, range_bits, lookup_bits);
    *out = finish(L, b);
    REQ((int)(*out)->witnesses == W, "range_check_cells_layout: internal error (witness count)");
    return H2SVD_OK;
}

// gate.is_equal(a, b) (verify_mul, src/matrix/mod.rs:339-341): sub [a - b, b, 1, a] + is_zero [z, d, inv, 1, 0, d, z, 0];
// input 0 = a (c_s . v total), input 1 = b (a . (b v) total); witness stripe = diff, is_zero, inv
int h2svd_is_equal_cells_layout(h2svd_cells_layout** out) {
    REQ(out, "is_equal_cells_layout: out is null");
    auto* L = new (std::nothrow) h2svd_cells_layout_impl();
    if (!L) return H2SVD_ENOMEM;
    Builder b{L};
    int row = b.here();
    const int d = b.W();
    b.E(-2);
    b.C(int_one());
    b.E(-1);
    b.gate(row);
    row = b.here();
    const int z = b.W();
    b.E(d);
    b.W();            // inv
    b.C(int_one());
    b.C(fr::zero());
    b.E(d);
    b.Wagain(z);      // is_zero once more: a second Witness cell with the same value
    b.C(fr::zero());
    b.gate(row);
    b.gate(row + 4);
    *out = finish(L, b);
    return H2SVD_OK;
}

void h2svd_cells_layout_destroy(h2svd_cells_layout* layout) {
    delete reinterpret_cast<h2svd_cells_layout_impl*>(layout);
}

static void run_threads(size_t units, int threads, const std::function<void(size_t, size_t)>& fn);

// out_values[u * cells + c] for every unit u: Witness -> wit[u * witnesses + source], Constant -> constants[source],
// Existing -> an earlier cell of the same unit or inputs[u * inputs + i].
int h2svd_expand_cells(const h2svd_cells_layout* layout, const h2svd_fr* inputs, const h2svd_fr* wit, size_t units,
                       h2svd_fr* out_values, int threads) {
    REQ(layout && out_values, "expand_cells: null argument");
    REQ(layout->witnesses == 0 || wit, "expand_cells: null witness array");
    REQ(layout->inputs == 0 || inputs, "expand_cells: null input array");
    const uint32_t cells = layout->cells, nw = layout->witnesses, ni = layout->inputs;
    for (uint32_t c = 0; c < cells; c++)
        REQ(layout->kind[c] != H2SVD_CELL_EXISTING || layout->source[c] < (int32_t)c, "expand_cells: malformed layout");
    run_threads(units, threads, [=](size_t u0, size_t u1) {
        for (size_t u = u0; u < u1; u++) {
            h2svd_fr* o = out_values + u * cells;
            const h2svd_fr* w = wit + u * nw;
            const h2svd_fr* in = inputs + u * ni;
            for (uint32_t c = 0; c < cells; c++) {
                const int32_t s = layout->source[c];
                switch (layout->kind[c]) {
                    case H2SVD_CELL_WITNESS: o[c] = w[s]; break;
                    case H2SVD_CELL_CONSTANT: o[c] = layout->constants[s]; break;
                    default: o[c] = s >= 0 ? o[s] : in[-1 - s]; break;
                }
            }
        }
    });
    return H2SVD_OK;
}

// field_mat_vec_mul (src/matrix/mod.rs:574-599) -> GateChip::inner_product per row: [0, a_0, v_0, s_0, a_1, v_1, s_1, ...]
// (1 + 3 * len cells per row, gate at every 3 * j); v_row_stride = 0 for a shared vector, len for per-row vectors
// (ZkVector::inner_product with u = x, v = self, :100).
int h2svd_expand_inner_product_cells(const h2svd_fr* a, const h2svd_fr* v, size_t v_row_stride, const h2svd_fr* prefix,
                                     size_t rows, size_t len, h2svd_fr* out_values, int threads) {
    REQ(a && v && prefix && out_values, "expand_inner_product_cells: null argument");
    const size_t per = 1 + 3 * len;
    run_threads(rows, threads, [=](size_t r0, size_t r1) {
        for (size_t r = r0; r < r1; r++) {
            h2svd_fr* o = out_values + r * per;
            memset(o, 0, sizeof(h2svd_fr));   // Constant(0): Montgomery form of 0 is 0
            const h2svd_fr* ar = a + r * len;
            const h2svd_fr* vr = v + r * v_row_stride;
            const h2svd_fr* pr = prefix + r * len;
            for (size_t j = 0; j < len; j++) {
                o[1 + 3 * j] = ar[j];
                o[2 + 3 * j] = vr[j];
                o[3 + 3 * j] = pr[j];
            }
        }
    });
    return H2SVD_OK;
}

// The challenge powers of verify_mul (:316-326): [one] then (m - 1) x gate.mul(v_{i-1}, gamma) = [0, v_{i-1}, gamma, v_i];
// 1 + 4 * (m - 1) cells, gates at 1 + 4 * i; `one` is additionally constrained to the constant 1 (assert_is_const).
int h2svd_expand_gamma_power_cells(const h2svd_fr* gamma, const h2svd_fr* powers, size_t m, h2svd_fr* out_values) {
    REQ(gamma && powers && out_values && m >= 1, "expand_gamma_power_cells: bad argument");
    out_values[0] = powers[0];
    for (size_t i = 1; i < m; i++) {
        h2svd_fr* o = out_values + 1 + 4 * (i - 1);
        memset(o, 0, sizeof(h2svd_fr));
        o[1] = powers[i - 1];
        o[2] = *gamma;
        o[3] = powers[i];
    }
    return H2SVD_OK;
}

// gate.is_equal per row from the arrays h2svd_freivalds_witness returns (the is_zero value appears twice in the cells)
int h2svd_expand_is_equal_cells(const h2svd_fr* x, const h2svd_fr* y, const h2svd_fr* diff, const h2svd_fr* is_zero,
                                const h2svd_fr* inv, size_t count, h2svd_fr* out_values) {
    REQ(x && y && diff && is_zero && inv && out_values, "expand_is_equal_cells: null argument");
    h2svd_fr one;
    const uint64_t c1[4] = {1, 0, 0, 0};
    h2svd_host_fr_from_canonical(c1, &one);
    for (size_t i = 0; i < count; i++) {
        h2svd_fr* o = out_values + 12 * i;
        memset(o, 0, 12 * sizeof(h2svd_fr));
        o[0] = diff[i];
        o[1] = y[i];
        o[2] = one;
        o[3] = x[i];
        o[4] = is_zero[i];
        o[5] = diff[i];
        o[6] = inv[i];
        o[7] = one;
        o[9] = diff[i];
        o[10] = is_zero[i];
    }
    return H2SVD_OK;
}

}  // extern "C"

static void run_threads(size_t units, int threads, const std::function<void(size_t, size_t)>& fn) {
    size_t nt = threads > 0 ? (size_t)threads : std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > units / 64 + 1) nt = units / 64 + 1;   // not worth a thread below 64 units
    if (nt <= 1) {
        fn(0, units);
        return;
    }
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; t++) {
        const size_t u0 = units * t / nt, u1 = units * (t + 1) / nt;
        th.emplace_back([&fn, u0, u1]() { fn(u0, u1); });
    }
    for (auto& t : th) t.join();
}
