// h2svd_zk.hpp -- C++17 host mirror of the reference's ZkMatrix / ZkVector API over the C ABI.
//
// The reference (neilcouture/halo2-svd041) is Rust; its toolchain is not available in this build
// environment, so the host side above include/h2svd_b200.h is written in C++ with the reference's own
// names, argument order and error behaviour (src/matrix/mod.rs), the way the Rust shim of
// INTEGRATION.md / rust/h2svd-b200 does it:
//   * every VALUE is produced by the GPU library (one bulk C-ABI call per operation);
//   * this layer only does what halo2-base's Context does with those values: push advice cells in the
//     reference's order, mark gate selectors, record copy constraints / constants / lookup cells.
// `Context` records exactly what halo2-base 0.4.1's virtual column records (cell layouts: SURVEY.md
// A.2-A.5), so a test can diff it cell by cell against the oracle's model of the reference, and
// `mock_verify` checks the gates / copies / constants / lookups the way MockProver would.
//
// Rust -> C++ naming: `T::new(...)` is `T::create(...)` (`new` is a keyword); a Rust `assert!`/panic is a
// thrown std::logic_error; `&Vec<T>` is `const std::vector<T>&`.  Everything else keeps its name.
//
// There is no CPU fallback: every operation that needs values calls libh2svd_b200 and throws if the
// library reports an error (e.g. H2SVD_ENODEV without a B200).
#ifndef H2SVD_ZK_HPP
#define H2SVD_ZK_HPP

#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "h2svd_b200.h"

namespace h2svd {
namespace zk {

using Fr = h2svd_fr;
inline bool operator==(const Fr& a, const Fr& b) {
    return a.l[0] == b.l[0] && a.l[1] == b.l[1] && a.l[2] == b.l[2] && a.l[3] == b.l[3];
}
inline bool operator!=(const Fr& a, const Fr& b) { return !(a == b); }

// ---- scalar field helpers (host, for Constant cells and gate checks) ---------------------------------
namespace field {
inline Fr from_u64(uint64_t x) {
    const uint64_t w[4] = {x, 0, 0, 0};
    Fr o;
    h2svd_host_fr_from_canonical(w, &o);
    return o;
}
inline Fr zero() { return Fr{{0, 0, 0, 0}}; }
inline Fr one() { return from_u64(1); }
inline Fr pow2(int bits) {  // 2^bits as a field element, 0 <= bits <= 253
    if (bits < 0 || bits > 253) throw std::logic_error("field::pow2: exponent out of range");
    uint64_t w[4] = {0, 0, 0, 0};
    w[bits >> 6] = 1ull << (bits & 63);
    Fr o;
    h2svd_host_fr_from_canonical(w, &o);
    return o;
}
inline Fr add(const Fr& a, const Fr& b) { Fr o; h2svd_host_fr_add(&a, &b, &o); return o; }
inline Fr sub(const Fr& a, const Fr& b) { Fr o; h2svd_host_fr_sub(&a, &b, &o); return o; }
inline Fr mul(const Fr& a, const Fr& b) { Fr o; h2svd_host_fr_mul(&a, &b, &o); return o; }
inline Fr neg(const Fr& a) { return sub(zero(), a); }
inline std::array<uint64_t, 4> canonical(const Fr& a) {
    std::array<uint64_t, 4> w;
    h2svd_host_fr_to_canonical(&a, w.data());
    return w;
}
}  // namespace field

inline void check(int rc, const char* what) {
    if (rc != H2SVD_OK) throw std::runtime_error(std::string(what) + ": " + h2svd_last_error());
}
inline void require(bool cond, const char* what) {  // the reference's assert!/assert_eq!
    if (!cond) throw std::logic_error(std::string("assertion failed: ") + what);
}

// ---- GPU handle --------------------------------------------------------------------------------------
class Gpu {
  public:
    explicit Gpu(int device = -1) { check(h2svd_create(&h_, device, nullptr), "h2svd_create"); }
    ~Gpu() { h2svd_destroy(h_); }
    Gpu(const Gpu&) = delete;
    Gpu& operator=(const Gpu&) = delete;
    h2svd_ctx* raw() const { return h_; }
    // The reference's free functions (honest_prover_mat_mul, field_mat_vec_mul, ...) take no chip or
    // handle; they use this per-thread default handle (the current CUDA device).
    static Gpu& current() {
        thread_local Gpu g(-1);
        return g;
    }

  private:
    h2svd_ctx* h_ = nullptr;
};

// ---- halo2-base Context model -------------------------------------------------------------------------
struct AssignedValue {
    Fr v{};
    uint32_t ctx_id = 0;
    size_t index = 0;
    const Fr& value() const { return v; }
};

enum class CellKind : uint8_t { Witness = 0, Existing = 1, Constant = 2 };
struct QuantumCell {
    CellKind kind;
    Fr v;
    AssignedValue src;  // for Existing
};
inline QuantumCell Existing(const AssignedValue& a) { return QuantumCell{CellKind::Existing, a.v, a}; }
inline QuantumCell Witness(const Fr& v) { return QuantumCell{CellKind::Witness, v, {}}; }
inline QuantumCell Constant(const Fr& v) { return QuantumCell{CellKind::Constant, v, {}}; }

struct CellRef {
    uint32_t ctx_id;
    size_t index;
};

class Context {
  public:
    explicit Context(uint32_t id = 0) : ctx_id(id) {}
    uint32_t ctx_id;
    std::vector<Fr> advice;
    std::vector<uint8_t> kind;      // CellKind per cell
    std::vector<uint8_t> selector;  // gate a + b*c - d == 0 on cells [i, i+4)
    std::vector<std::pair<CellRef, CellRef>> copies;
    std::vector<std::pair<size_t, Fr>> constants;
    std::vector<size_t> lookups;    // cells constrained to [0, 2^lookup_bits)

    AssignedValue get(ptrdiff_t i) const {
        const size_t idx = i < 0 ? advice.size() + i : (size_t)i;
        return AssignedValue{advice[idx], ctx_id, idx};
    }
    AssignedValue last() const { return get(-1); }
    // Pushes the cells in order, switches the gate on at the given offsets, returns the first row.
    size_t assign_region(const std::vector<QuantumCell>& cells, const std::vector<size_t>& gate_offsets) {
        const size_t row = advice.size();
        for (const QuantumCell& c : cells) push(c);
        for (size_t off : gate_offsets) selector[row + off] = 1;
        return row;
    }
    AssignedValue load_witness(const Fr& v) {
        push(Witness(v));
        return last();
    }
    AssignedValue load_constant(const Fr& v) {
        push(Constant(v));
        return last();
    }
    // bulk load_witness (the shim's replacement for the reference's per-cell loop, e.g. :558-565)
    std::vector<AssignedValue> assign_witnesses(const Fr* v, size_t n) {
        std::vector<AssignedValue> out;
        out.reserve(n);
        const size_t row = advice.size();
        advice.insert(advice.end(), v, v + n);
        kind.insert(kind.end(), n, (uint8_t)CellKind::Witness);
        selector.insert(selector.end(), n, 0);
        for (size_t i = 0; i < n; i++) out.push_back(AssignedValue{v[i], ctx_id, row + i});
        return out;
    }
    void constrain_equal(const AssignedValue& a, const AssignedValue& b) {
        copies.push_back({CellRef{a.ctx_id, a.index}, CellRef{b.ctx_id, b.index}});
    }
    // Bulk form of `units` identical assign_region sequences: `values` holds the complete cell stream
    // (units * layout->cells values, from h2svd_expand_cells) and is appended with ONE insert; the layout's offsets are
    // then replayed unit by unit for kinds / selectors / lookups / copy constraints / constants.  inputs[u * layout->inputs
    // + i] is the i-th input cell of unit u.  Leaves the context exactly as the per-cell path would (same vectors, same order
    // within every unit).  Returns the index of the first appended cell.
    size_t append_units(const h2svd_cells_layout* layout, const Fr* values, size_t units, const AssignedValue* inputs) {
        const size_t base = advice.size(), cells = layout->cells;
        advice.insert(advice.end(), values, values + units * cells);
        kind.resize(base + units * cells);
        selector.resize(base + units * cells, 0);
        for (size_t u = 0; u < units; u++) {
            const size_t row = base + u * cells;
            auto ref = [&](int32_t off) -> CellRef {
                if (off >= 0) return CellRef{ctx_id, row + (size_t)off};
                const AssignedValue& in = inputs[u * layout->inputs + (size_t)(-1 - off)];
                return CellRef{in.ctx_id, in.index};
            };
            // the per-cell path records a copy / constant as each cell is pushed and an extra constrain_equal(a, acc) right
            // after the region that ends in `acc`: replaying the extra pairs (stored in order) after their second cell
            // reproduces the very same vectors
            uint32_t next_copy = 0;
            for (uint32_t c = 0; c < cells; c++) {
                const uint8_t k = layout->kind[c];
                kind[row + c] = k == H2SVD_CELL_WITNESS ? (uint8_t)CellKind::Witness
                                : k == H2SVD_CELL_CONSTANT ? (uint8_t)CellKind::Constant : (uint8_t)CellKind::Existing;
                if (k == H2SVD_CELL_EXISTING) copies.push_back({ref(layout->source[c]), CellRef{ctx_id, row + c}});
                if (k == H2SVD_CELL_CONSTANT) constants.push_back({row + c, layout->constants[layout->source[c]]});
                while (next_copy < layout->n_copies && layout->copies[2 * next_copy + 1] == (int32_t)c) {
                    copies.push_back({ref(layout->copies[2 * next_copy]), ref(layout->copies[2 * next_copy + 1])});
                    next_copy++;
                }
            }
            for (uint32_t g = 0; g < layout->n_gates; g++) selector[row + layout->gates[g]] = 1;
            for (uint32_t l = 0; l < layout->n_lookups; l++) lookups.push_back(ref(layout->lookups[l]).index);
        }
        return base;
    }

  private:
    void push(const QuantumCell& c) {
        const size_t idx = advice.size();
        advice.push_back(c.v);
        kind.push_back((uint8_t)c.kind);
        selector.push_back(0);
        if (c.kind == CellKind::Existing) copies.push_back({CellRef{c.src.ctx_id, c.src.index}, CellRef{ctx_id, idx}});
        if (c.kind == CellKind::Constant) constants.push_back({idx, c.v});
    }
};

// ---- GateChip: cell layouts of halo2-base 0.4.1; values supplied by the caller -----------
class GateChip {
  public:
    // [a, b, 1, a+b] -> last
    AssignedValue add(Context& ctx, const QuantumCell& a, const QuantumCell& b, const Fr& sum) const {
        ctx.assign_region({a, b, Constant(field::one()), Witness(sum)}, {0});
        return ctx.get(-1);
    }
    // [a-b, b, 1, a] -> first
    AssignedValue sub(Context& ctx, const QuantumCell& a, const QuantumCell& b, const Fr& diff) const {
        ctx.assign_region({Witness(diff), b, Constant(field::one()), a}, {0});
        return ctx.get(-4);
    }
    // [0, a, b, a*b] -> last
    AssignedValue mul(Context& ctx, const QuantumCell& a, const QuantumCell& b, const Fr& prod) const {
        ctx.assign_region({Constant(field::zero()), a, b, Witness(prod)}, {0});
        return ctx.get(-1);
    }
    // [0, a0, b0, s0, a1, b1, s1, ...] with the running sums s_j supplied (GPU-produced) -> last
    AssignedValue inner_product(Context& ctx, const std::vector<AssignedValue>& a, const std::vector<AssignedValue>& b,
                                const Fr* prefix) const {
        require(a.size() == b.size(), "inner_product: a.len() == b.len()");
        std::vector<QuantumCell> cells;
        cells.reserve(1 + 3 * a.size());
        cells.push_back(Constant(field::zero()));
        std::vector<size_t> gates(a.size());
        for (size_t j = 0; j < a.size(); j++) {
            cells.push_back(Existing(a[j]));
            cells.push_back(Existing(b[j]));
            cells.push_back(Witness(prefix[j]));
            gates[j] = 3 * j;
        }
        ctx.assign_region(cells, gates);
        return ctx.get(-1);
    }
    // range_check's limb recomposition: inner_product(limbs, [1, 2^lb, 2^2lb, ...]) where b starts with
    // Constant(1): cells [l0, l1, 2^lb, s1, l2, 2^2lb, s2, ...]; wit = l0, l1, s1, l2, s2, ...
    // Returns the row of l0.
    size_t limb_inner_product(Context& ctx, const Fr* wit, int n, int lb) const {
        std::vector<QuantumCell> cells;
        cells.reserve(3 * n - 2);
        cells.push_back(Witness(wit[0]));
        std::vector<size_t> gates;
        for (int i = 1; i < n; i++) {
            cells.push_back(Witness(wit[2 * i - 1]));
            cells.push_back(Constant(field::pow2(lb * i)));
            cells.push_back(Witness(wit[2 * i]));
            gates.push_back(3 * (i - 1));
        }
        return ctx.assign_region(cells, gates);
    }
    // is_zero: [z, a, inv, 1, 0, a, z, 0], gates at 0 and 4 -> cell 6
    AssignedValue is_zero(Context& ctx, const AssignedValue& a, const Fr& z, const Fr& inv) const {
        ctx.assign_region({Witness(z), Existing(a), Witness(inv), Constant(field::one()), Constant(field::zero()),
                           Existing(a), Witness(z), Constant(field::zero())},
                          {0, 4});
        return ctx.get(-2);
    }
    void assert_is_const(Context& ctx, const AssignedValue& a, const Fr& c) const {
        // halo2-base: constrain_equal(a, constant) -- recorded as a constant on the existing cell, no new cell
        ctx.constants.push_back({a.index, c});
        (void)a;
    }
};

// ---- RangeChip -------------------------------------------------------------------------
class RangeChip {
  public:
    explicit RangeChip(int lookup_bits) : lookup_bits(lookup_bits) {}
    int lookup_bits;
    GateChip gate;

    // range_check(a, n*lb) given wit = l0, l1, s1, l2, s2, ... (2n-1 values; none when n == 1).
    // Returns the number of witness values consumed.
    int range_check(Context& ctx, const AssignedValue& a, int n, const Fr* wit) const {
        if (n == 1) {
            ctx.lookups.push_back(a.index);
            return 0;
        }
        const size_t row = gate.limb_inner_product(ctx, wit, n, lookup_bits);
        ctx.constrain_equal(a, ctx.get(-1));
        ctx.lookups.push_back(row);
        for (int i = 0; i < n - 1; i++) ctx.lookups.push_back(row + 1 + 3 * i);
        return 2 * n - 1;
    }
    // range_check(a, range_bits) for any width (halo2-base RangeChip::range_check): n = ceil(range_bits/lb) limbs;
    // when range_bits % lb = rem: rem == 1 -> assert_bit-style gate on the last limb, rem > 1 -> the last limb
    // times 2^(lb-rem) (one more witness) is looked up too.  Returns the number of witness values consumed.
    int range_check_bits(Context& ctx, const AssignedValue& a, int range_bits, const Fr* wit) const {
        const int n = (range_bits + lookup_bits - 1) / lookup_bits, rem = range_bits % lookup_bits;
        int used = 0;
        AssignedValue last = a;
        if (n == 1) {
            ctx.lookups.push_back(a.index);
        } else {
            const size_t row = gate.limb_inner_product(ctx, wit, n, lookup_bits);
            ctx.constrain_equal(a, ctx.get(-1));
            ctx.lookups.push_back(row);
            for (int i = 0; i < n - 1; i++) ctx.lookups.push_back(row + 1 + 3 * i);
            last = ctx.get((ptrdiff_t)(row + 1 + 3 * (n - 2)));
            used = 2 * n - 1;
        }
        if (rem == 1) {
            ctx.assign_region({Constant(field::zero()), Existing(last), Existing(last), Existing(last)}, {0});
        } else if (rem > 1) {
            const AssignedValue chk = gate.mul(ctx, Existing(last), Constant(field::pow2(lookup_bits - rem)), wit[used]);
            ctx.lookups.push_back(chk.index);
            used++;
        }
        return used;
    }
    // check_big_less_than_safe(a, bound) with n = ceil(bound.bits()/lb):
    //   wit = range_check(a) | chk, xp | range_check(chk)        (4n values, 4 when n == 1 -> just chk, xp)
    int check_big_less_than_safe(Context& ctx, const AssignedValue& a, const Fr& bound, int n, const Fr* wit) const {
        int used = range_check(ctx, a, n, wit);
        const int bits = n * lookup_bits;
        // check_less_than: [a + 2^bits - b, b, 1, a + 2^bits, -2^bits, 1, a], gates 0 and 3
        ctx.assign_region({Witness(wit[used]), Constant(bound), Constant(field::one()), Witness(wit[used + 1]),
                           Constant(field::neg(field::pow2(bits))), Constant(field::one()), Existing(a)},
                          {0, 3});
        const AssignedValue chk = ctx.get(-7);
        used += 2;
        used += range_check(ctx, chk, n, wit + used);
        return used;
    }
};

// ---- FixedPointChip041<PRECISION_BITS> (third-party in the reference; model: SURVEY.md A.5) ----------------
template <uint32_t PRECISION_BITS>
class FixedPointChip041 {
  public:
    explicit FixedPointChip041(int lookup_bits, int shift_bits = -1, int a_num_bits = -1, Gpu* gpu = nullptr)
        : lookup_bits(lookup_bits),
          S(shift_bits < 0 ? 3 * (int)PRECISION_BITS : shift_bits),
          A(a_num_bits < 0 ? 4 * (int)PRECISION_BITS : a_num_bits),
          range_(lookup_bits),
          gpu_(gpu ? gpu : &Gpu::current()) {
        static_assert(PRECISION_BITS >= 1 && PRECISION_BITS <= 63, "PRECISION_BITS out of range");
        W = h2svd_rescale_witness_count((int)PRECISION_BITS, lookup_bits, S, A);
        if (W < 0) throw std::logic_error(std::string("FixedPointChip041: ") + h2svd_last_error());
        n_d = (A - (int)PRECISION_BITS + 1 + lookup_bits - 1) / lookup_bits;
        n_r = ((int)PRECISION_BITS + 1 + lookup_bits - 1) / lookup_bits;
    }
    int lookup_bits, S, A, W = 0, n_d = 0, n_r = 0;
    const GateChip& gate() const { return range_.gate; }
    const RangeChip& range_gate() const { return range_; }
    Gpu& gpu() const { return *gpu_; }

    Fr quantization(double x) const {
        Fr o;
        check(h2svd_quantize(gpu_->raw(), &x, 1, (int)PRECISION_BITS, &o), "quantization");
        return o;
    }
    std::vector<Fr> quantization(const std::vector<double>& x) const {
        std::vector<Fr> o(x.size());
        if (!x.empty()) check(h2svd_quantize(gpu_->raw(), x.data(), x.size(), (int)PRECISION_BITS, o.data()), "quantization");
        return o;
    }
    // values above (r-1)/2 are negative; result = signed integer / 2^P
    double dequantization(const Fr& v) const {
        static const std::array<uint64_t, 4> half = field::canonical(field::neg(field::one()));  // r - 1
        std::array<uint64_t, 4> w = field::canonical(v);
        bool negative = false;
        // compare w with (r-1)/2
        std::array<uint64_t, 4> h;
        for (int i = 0; i < 4; i++) h[i] = (half[i] >> 1) | (i < 3 ? half[i + 1] << 63 : 0);
        for (int i = 3; i >= 0; i--) {
            if (w[i] != h[i]) {
                negative = w[i] > h[i];
                break;
            }
        }
        if (negative) w = field::canonical(field::neg(v));
        long double mag = 0;
        for (int i = 3; i >= 0; i--) mag = mag * 18446744073709551616.0L + (long double)w[i];
        const double r = (double)std::ldexp(mag, -(int)PRECISION_BITS);
        return negative ? -r : r;
    }
    // qsub = gate.sub (src/matrix/mod.rs:76): [a-b, b, 1, a] -> first
    AssignedValue qsub(Context& ctx, const AssignedValue& a, const AssignedValue& b) const {
        Fr d;
        check(h2svd_zkvec_sub(gpu_->raw(), &a.v, &b.v, 1, &d), "qsub");
        return gate().sub(ctx, Existing(a), Existing(b), d);
    }
    // Cells of ONE signed_div_scale(a) given its W GPU-produced Witness values.
    std::pair<AssignedValue, AssignedValue> assign_signed_div_scale(Context& ctx, const AssignedValue& a,
                                                                     const Fr* w) const {
        const int P = (int)PRECISION_BITS;
        const AssignedValue a_shift = gate().add(ctx, Existing(a), Constant(field::pow2(S)), w[0]);
        // div_mod: [rem, 2^P, div, a_shift], gate 0
        ctx.assign_region({Witness(w[1]), Constant(field::pow2(P)), Witness(w[2]), Existing(a_shift)}, {0});
        const AssignedValue rem = ctx.get(-4), div = ctx.get(-2);
        int used = 3;
        const Fr bound_d = field::add(field::pow2(A - P), field::one());
        used += range_.check_big_less_than_safe(ctx, div, bound_d, n_d, w + used);
        used += range_.check_big_less_than_safe(ctx, rem, field::pow2(P), n_r, w + used);
        const AssignedValue q = gate().sub(ctx, Existing(div), Constant(field::pow2(S - P)), w[used]);
        require(used + 1 == W, "signed_div_scale consumed W witnesses");
        return {q, rem};
    }
    std::pair<AssignedValue, AssignedValue> signed_div_scale(Context& ctx, const AssignedValue& a) const {
        std::vector<Fr> w((size_t)W);
        Fr q;
        check(h2svd_rescale_witness(gpu_->raw(), &a.v, 1, (int)PRECISION_BITS, lookup_bits, S, A, &q, w.data()),
              "signed_div_scale");
        return assign_signed_div_scale(ctx, a, w.data());
    }
    // qsqrt: value model only (parity unpinned): one witness cell floor(sqrt(a * 2^P))
    AssignedValue qsqrt(Context& ctx, const AssignedValue& a) const {
        Fr o;
        check(h2svd_isqrt_fixed(gpu_->raw(), &a.v, 1, (int)PRECISION_BITS, &o), "qsqrt");
        return ctx.load_witness(o);
    }

  private:
    RangeChip range_;
    Gpu* gpu_;
};

using AssignedMatrix = std::vector<std::vector<AssignedValue>>;

inline std::vector<Fr> gather_values(const AssignedMatrix& m) {
    std::vector<Fr> out;
    out.reserve(m.size() * (m.empty() ? 0 : m[0].size()));
    for (const auto& row : m)
        for (const auto& c : row) out.push_back(c.v);
    return out;
}
inline std::vector<Fr> gather_values(const std::vector<AssignedValue>& v) {
    std::vector<Fr> out;
    out.reserve(v.size());
    for (const auto& c : v) out.push_back(c.v);
    return out;
}

// ---- free functions of src/matrix/mod.rs ------------------------------------------------------------------
// field_mat_mul (:510-537): c = a*b over Fr, outside the circuit
inline std::vector<std::vector<Fr>> field_mat_mul(const AssignedMatrix& a, const AssignedMatrix& b) {
    require(!a.empty() && !b.empty() && a[0].size() == b.size(), "a[0].len() == b.len()");  // :515
    const size_t n = a.size(), k = b.size(), m = b[0].size();
    const std::vector<Fr> fa = gather_values(a), fb = gather_values(b);
    std::vector<Fr> c(n * m);
    check(h2svd_fr_matmul(Gpu::current().raw(), fa.data(), fb.data(), c.data(), n, k, m, 0), "field_mat_mul");
    std::vector<std::vector<Fr>> out(n);
    for (size_t i = 0; i < n; i++) out[i].assign(c.begin() + i * m, c.begin() + (i + 1) * m);
    return out;
}
// honest_prover_mat_mul (:546-568): c_s loaded row-major as unconstrained witnesses
inline AssignedMatrix honest_prover_mat_mul(Context& ctx, const AssignedMatrix& a, const AssignedMatrix& b) {
    const std::vector<std::vector<Fr>> c_s = field_mat_mul(a, b);
    AssignedMatrix out;
    out.reserve(c_s.size());
    for (const auto& row : c_s) out.push_back(ctx.assign_witnesses(row.data(), row.size()));
    return out;
}
// field_mat_vec_mul (:574-599): y_i = <a_i, v> with every running sum assigned
inline std::vector<AssignedValue> field_mat_vec_mul(Context& ctx, const GateChip& gate, const AssignedMatrix& a,
                                                    const std::vector<AssignedValue>& v) {
    require(!a.empty() && a[0].size() == v.size(), "a[0].len() == v.len()");  // :580
    const size_t rows = a.size(), len = v.size();
    const std::vector<Fr> fa = gather_values(a), fv = gather_values(v);
    std::vector<Fr> prefix(rows * len);
    check(h2svd_mat_vec_prefix(Gpu::current().raw(), fa.data(), fv.data(), rows, len, prefix.data()), "field_mat_vec_mul");
    std::vector<AssignedValue> y;
    y.reserve(rows);
    for (size_t i = 0; i < rows; i++) y.push_back(gate.inner_product(ctx, a[i], v, prefix.data() + i * len));
    return y;
}

template <uint32_t PRECISION_BITS>
class ZkMatrix;

// ---- ZkVector (:21-216) --------------------------------------------------------------------------------
template <uint32_t PRECISION_BITS>
class ZkVector {
  public:
    using Chip = FixedPointChip041<PRECISION_BITS>;
    std::vector<AssignedValue> v;

    // ZkVector::new (:29-40)
    static ZkVector create(Context& ctx, const Chip& fpchip, const std::vector<double>& x) {
        const std::vector<Fr> q = fpchip.quantization(x);
        return ZkVector{ctx.assign_witnesses(q.data(), q.size())};
    }
    size_t size() const { return v.size(); }  // :43
    std::vector<double> dequantize(const Chip& fpchip) const {  // :50
        std::vector<double> out;
        for (const auto& e : v) out.push_back(fpchip.dequantization(e.v));
        return out;
    }
    // inner_product (:79-106): gate.inner_product(u = x, v = self) then one signed_div_scale
    AssignedValue inner_product(Context& ctx, const Chip& fpchip, const std::vector<AssignedValue>& x) const {
        require(size() == x.size(), "self.size() == x.len()");  // :86
        const std::vector<Fr> fx = gather_values(x), fs = gather_values(v);
        std::vector<Fr> prefix(x.size());
        check(h2svd_zkvec_inner_prefix(fpchip.gpu().raw(), fx.data(), fs.data(), 1, x.size(), prefix.data()),
              "ZkVector::inner_product");
        const AssignedValue res_s = fpchip.gate().inner_product(ctx, x, v, prefix.data());
        return fpchip.signed_div_scale(ctx, res_s).first;
    }
    AssignedValue _norm_square(Context& ctx, const Chip& fpchip) const { return inner_product(ctx, fpchip, v); }  // :111
    AssignedValue norm(Context& ctx, const Chip& fpchip) const {  // :124
        return fpchip.qsqrt(ctx, _norm_square(ctx, fpchip));
    }
    // _dist_square (:136-149): n x qsub (one bulk call), then the norm square of the differences
    AssignedValue _dist_square(Context& ctx, const Chip& fpchip, const std::vector<AssignedValue>& x) const {
        require(size() == x.size(), "self.size() == x.len()");  // :142
        const std::vector<Fr> fs = gather_values(v), fx = gather_values(x);
        std::vector<Fr> d(x.size());
        check(h2svd_zkvec_sub(fpchip.gpu().raw(), fs.data(), fx.data(), x.size(), d.data()), "ZkVector::_dist_square");
        ZkVector diff;
        for (size_t i = 0; i < x.size(); i++)
            diff.v.push_back(fpchip.gate().sub(ctx, Existing(v[i]), Existing(x[i]), d[i]));
        return diff._norm_square(ctx, fpchip);
    }
    AssignedValue dist(Context& ctx, const Chip& fpchip, const std::vector<AssignedValue>& x) const {  // :156
        return fpchip.qsqrt(ctx, _dist_square(ctx, fpchip, x));
    }
    // mul (:169-182): y_i = inner_product(row_i) for every row of a -- two bulk calls for all rows,
    // cells emitted per row in the reference's order (inner product, then its signed_div_scale)
    ZkVector mul(Context& ctx, const Chip& fpchip, const ZkMatrix<PRECISION_BITS>& a) const;
    // entries_less_than (:185-197): range_check(elem, max_bits) for every entry (one bulk call)
    void entries_less_than(Context& ctx, const Chip& fpchip, size_t max_bits) const {
        range_check_all(ctx, fpchip, v, max_bits);
    }
    // entries_in_desc_order (:199-216): all the qsub differences first, then their range checks
    void entries_in_desc_order(Context& ctx, const Chip& fpchip, size_t max_bits) const {
        if (v.size() < 2) return;
        const size_t n = v.size() - 1;
        const std::vector<Fr> fs = gather_values(v);
        std::vector<Fr> d(n);
        check(h2svd_zkvec_sub(fpchip.gpu().raw(), fs.data(), fs.data() + 1, n, d.data()), "entries_in_desc_order");
        std::vector<AssignedValue> vec_diff;
        for (size_t i = 0; i < n; i++) vec_diff.push_back(fpchip.gate().sub(ctx, Existing(v[i]), Existing(v[i + 1]), d[i]));
        range_check_all(ctx, fpchip, vec_diff, max_bits);
    }

  private:
    static void range_check_all(Context& ctx, const Chip& fpchip, const std::vector<AssignedValue>& xs, size_t max_bits) {
        const int W = h2svd_range_check_witness_count((int)max_bits, fpchip.lookup_bits);
        if (W < 0) throw std::logic_error(std::string("range_check: ") + h2svd_last_error());
        const std::vector<Fr> fx = gather_values(xs);
        std::vector<Fr> w(xs.size() * (size_t)W + 1);
        if (W > 0 && !xs.empty())
            check(h2svd_range_check_witness(fpchip.gpu().raw(), fx.data(), xs.size(), (int)max_bits, fpchip.lookup_bits,
                                            w.data()),
                  "range_check");
        for (size_t i = 0; i < xs.size(); i++)
            fpchip.range_gate().range_check_bits(ctx, xs[i], (int)max_bits, w.data() + i * (size_t)W);
    }

  public:
};

// ---- ZkMatrix (:219-420) ---------------------------------------------------------------------------------
template <uint32_t PRECISION_BITS>
class ZkMatrix {
  public:
    using Chip = FixedPointChip041<PRECISION_BITS>;
    AssignedMatrix matrix;
    size_t num_rows = 0, num_col = 0;

    // ZkMatrix::new (:230-252)
    static ZkMatrix create(Context& ctx, const Chip& fpchip, const std::vector<std::vector<double>>& m) {
        require(!m.empty(), "matrix is not empty");
        ZkMatrix out;
        out.num_rows = m.size();
        out.num_col = m[0].size();
        std::vector<double> flat;
        for (const auto& row : m) {
            require(row.size() == out.num_col, "row.len() == num_col");  // :239
            flat.insert(flat.end(), row.begin(), row.end());
        }
        const std::vector<Fr> q = fpchip.quantization(flat);
        for (size_t i = 0; i < out.num_rows; i++)
            out.matrix.push_back(ctx.assign_witnesses(q.data() + i * out.num_col, out.num_col));
        return out;
    }
    std::vector<std::vector<double>> dequantize(const Chip& fpchip) const {  // :257
        std::vector<std::vector<double>> out(num_rows);
        for (size_t i = 0; i < num_rows; i++)
            for (size_t j = 0; j < num_col; j++) out[i].push_back(fpchip.dequantization(matrix[i][j].v));
        return out;
    }
    // verify_mul (:299-342): Freivalds with v = (1, gamma, ..., gamma^(d-1))
    static void verify_mul(Context& ctx, const Chip& fpchip, const ZkMatrix& a, const ZkMatrix& b,
                           const AssignedMatrix& c_s, const AssignedValue& init_rand, bool strict = false) {
        require(a.num_col == b.num_rows, "a.num_col == b.num_rows");            // :307
        require(c_s.size() == a.num_rows, "c_s.len() == a.num_rows");            // :308
        require(!c_s.empty() && c_s[0].size() == b.num_col, "c_s[0].len() == b.num_col");  // :309
        require(c_s[0].size() >= 1, "c_s[0].len() >= 1");                        // :310
        const size_t n = a.num_rows, k = a.num_col, m = b.num_col;
        const std::vector<Fr> fa = gather_values(a.matrix), fb = gather_values(b.matrix), fc = gather_values(c_s);
        std::vector<Fr> pw(m), pcv(n * m), pbv(k * m), pabv(n * k), diff(n), isz(n), inv(n);
        check(h2svd_freivalds_witness(fpchip.gpu().raw(), fa.data(), fb.data(), fc.data(), &init_rand.v, n, k, m,
                                      pw.data(), pcv.data(), pbv.data(), pabv.data(), diff.data(), isz.data(),
                                      inv.data()),
              "ZkMatrix::verify_mul");
        const GateChip& gate = fpchip.gate();
        std::vector<AssignedValue> v;
        const AssignedValue one = ctx.load_witness(pw[0]);  // :318
        gate.assert_is_const(ctx, one, field::one());        // :319
        v.push_back(one);
        for (size_t i = 1; i < m; i++) v.push_back(gate.mul(ctx, Existing(v[i - 1]), Existing(init_rand), pw[i]));  // :322-326
        std::vector<AssignedValue> cs_v, b_v, ab_v;
        for (size_t i = 0; i < n; i++) cs_v.push_back(gate.inner_product(ctx, c_s[i], v, pcv.data() + i * m));         // :335
        for (size_t i = 0; i < k; i++) b_v.push_back(gate.inner_product(ctx, b.matrix[i], v, pbv.data() + i * m));    // :336
        for (size_t i = 0; i < n; i++) ab_v.push_back(gate.inner_product(ctx, a.matrix[i], b_v, pabv.data() + i * k)); // :337
        for (size_t i = 0; i < n; i++) {                                                                               // :339-341
            const AssignedValue d = gate.sub(ctx, Existing(cs_v[i]), Existing(ab_v[i]), diff[i]);
            const AssignedValue z = gate.is_zero(ctx, d, isz[i], inv[i]);
            // The reference DISCARDS this boolean (gate.is_equal's result is never constrained, :339-341), so a wrong
            // c_s still satisfies its circuit.  `strict` is this library's opt-in fix: z == 1.
            if (strict) gate.assert_is_const(ctx, z, field::one());
        }
    }
    // verify_mul with the Freivalds result actually constrained (NOT the reference's constraint system)
    static void verify_mul_strict(Context& ctx, const Chip& fpchip, const ZkMatrix& a, const ZkMatrix& b,
                                  const AssignedMatrix& c_s, const AssignedValue& init_rand) {
        verify_mul(ctx, fpchip, a, b, c_s, init_rand, true);
    }
    // rescale_matrix (:354-375): one signed_div_scale per element, row-major
    static ZkMatrix rescale_matrix(Context& ctx, const Chip& fpchip, const AssignedMatrix& c_s) {
        require(!c_s.empty(), "c_s is not empty");
        const size_t rows = c_s.size(), cols = c_s[0].size();
        const std::vector<Fr> fc = gather_values(c_s);
        std::vector<Fr> q(rows * cols), wit(rows * cols * (size_t)fpchip.W);
        check(h2svd_rescale_witness(fpchip.gpu().raw(), fc.data(), rows * cols, (int)PRECISION_BITS, fpchip.lookup_bits,
                                    fpchip.S, fpchip.A, q.data(), wit.data()),
              "ZkMatrix::rescale_matrix");
        ZkMatrix out;
        out.num_rows = rows;
        out.num_col = cols;
        for (size_t i = 0; i < rows; i++) {
            std::vector<AssignedValue> new_row;
            for (size_t j = 0; j < cols; j++)
                new_row.push_back(
                    fpchip.assign_signed_div_scale(ctx, c_s[i][j], wit.data() + (i * cols + j) * (size_t)fpchip.W).first);
            out.matrix.push_back(std::move(new_row));
        }
        return out;
    }
    // The same rescale_matrix with the bulk hand-off: one GPU call, one h2svd_expand_cells pass, ONE append
    // of the complete cell stream instead of ~100 assign_region pushes per element.
    static ZkMatrix rescale_matrix_bulk(Context& ctx, const Chip& fpchip, const AssignedMatrix& c_s) {
        require(!c_s.empty(), "c_s is not empty");
        const size_t rows = c_s.size(), cols = c_s[0].size(), count = rows * cols;
        const std::vector<Fr> fc = gather_values(c_s);
        std::vector<Fr> q(count), wit(count * (size_t)fpchip.W);
        check(h2svd_rescale_witness(fpchip.gpu().raw(), fc.data(), count, (int)PRECISION_BITS, fpchip.lookup_bits, fpchip.S,
                                    fpchip.A, q.data(), wit.data()),
              "ZkMatrix::rescale_matrix_bulk");
        h2svd_cells_layout* layout = nullptr;
        check(h2svd_rescale_cells_layout((int)PRECISION_BITS, fpchip.lookup_bits, fpchip.S, fpchip.A, &layout),
              "h2svd_rescale_cells_layout");
        std::vector<Fr> values(count * (size_t)layout->cells);
        check(h2svd_expand_cells(layout, fc.data(), wit.data(), count, values.data(), 0), "h2svd_expand_cells");
        std::vector<AssignedValue> inputs;
        inputs.reserve(count);
        for (const auto& row : c_s) inputs.insert(inputs.end(), row.begin(), row.end());
        const size_t base = ctx.append_units(layout, values.data(), count, inputs.data());
        ZkMatrix out;
        out.num_rows = rows;
        out.num_col = cols;
        // the quotient cell is the first cell of the final gate.sub region: 4 cells before the end of the unit
        const size_t q_off = layout->cells - 4;
        for (size_t i = 0; i < rows; i++) {
            std::vector<AssignedValue> new_row;
            for (size_t j = 0; j < cols; j++) new_row.push_back(ctx.get((ptrdiff_t)(base + (i * cols + j) * layout->cells + q_off)));
            out.matrix.push_back(std::move(new_row));
        }
        h2svd_cells_layout_destroy(layout);
        return out;
    }
    // transpose_matrix (:408-419): copies cells, no constraints
    static ZkMatrix transpose_matrix(const ZkMatrix& a) {
        ZkMatrix out;
        out.num_rows = a.num_col;
        out.num_col = a.num_rows;
        out.matrix.assign(a.num_col, std::vector<AssignedValue>(a.num_rows));
        for (size_t i = 0; i < a.num_rows; i++)
            for (size_t j = 0; j < a.num_col; j++) out.matrix[j][i] = a.matrix[i][j];
        return out;
    }
};

template <uint32_t PRECISION_BITS>
ZkVector<PRECISION_BITS> ZkVector<PRECISION_BITS>::mul(Context& ctx, const Chip& fpchip,
                                                        const ZkMatrix<PRECISION_BITS>& a) const {
    require(a.num_col == size(), "a.num_col == self.size()");  // :175
    const size_t rows = a.num_rows, len = size();
    const std::vector<Fr> fa = gather_values(a.matrix), fs = gather_values(v);
    std::vector<Fr> prefix(rows * len), totals(rows), q(rows), wit(rows * (size_t)fpchip.W);
    // gate.inner_product(u = row, v = self): the row is the first operand (:100)
    check(h2svd_mat_vec_prefix(fpchip.gpu().raw(), fa.data(), fs.data(), rows, len, prefix.data()), "ZkVector::mul");
    for (size_t i = 0; i < rows; i++) totals[i] = prefix[i * len + len - 1];
    check(h2svd_rescale_witness(fpchip.gpu().raw(), totals.data(), rows, (int)PRECISION_BITS, fpchip.lookup_bits,
                                fpchip.S, fpchip.A, q.data(), wit.data()),
          "ZkVector::mul");
    ZkVector y;
    for (size_t i = 0; i < rows; i++) {
        const AssignedValue res_s = fpchip.gate().inner_product(ctx, a.matrix[i], v, prefix.data() + i * len);
        y.v.push_back(fpchip.assign_signed_div_scale(ctx, res_s, wit.data() + i * (size_t)fpchip.W).first);
    }
    return y;
}

// ---- range-check helpers of src/matrix/mod.rs (:425-501, :610-627) ----------------------------------------------
using BigUint = std::array<uint64_t, 4>;  // canonical integer, little-endian limbs (the reference passes &BigUint)
inline BigUint biguint(unsigned __int128 x) { return BigUint{(uint64_t)x, (uint64_t)(x >> 64), 0, 0}; }
inline Fr biguint_to_fe(const BigUint& x) {
    Fr o;
    check(h2svd_host_fr_from_canonical(x.data(), &o), "biguint_to_fe");
    return o;
}
inline BigUint biguint_sub1(const BigUint& x) {
    BigUint o = x;
    for (int i = 0; i < 4; i++)
        if (o[i]-- != 0) break;
    return o;
}
inline BigUint biguint_2x_minus1(const BigUint& x) {
    BigUint o;
    uint64_t c = 0;
    for (int i = 0; i < 4; i++) {
        o[i] = (x[i] << 1) | c;
        c = x[i] >> 63;
    }
    return biguint_sub1(o);
}
inline int biguint_bits(const BigUint& x) {
    for (int i = 3; i >= 0; i--)
        if (x[i]) return 64 * i + (64 - __builtin_clzll(x[i]));
    return 0;
}

// Cells of ONE check_abs_less_than(x, bnd) (:425-437) given its GPU-produced witnesses w = t | cbls(t, 2*bnd-1).
inline void assign_check_abs_less_than(Context& ctx, const RangeChip& range, const AssignedValue& x, const BigUint& bnd,
                                       const Fr* w) {
    const BigUint new_bnd = biguint_2x_minus1(bnd);
    const int n = (biguint_bits(new_bnd) + range.lookup_bits - 1) / range.lookup_bits;
    const AssignedValue translated_x = range.gate.add(ctx, Existing(x), Constant(biguint_to_fe(biguint_sub1(bnd))), w[0]);
    range.check_big_less_than_safe(ctx, translated_x, biguint_to_fe(new_bnd), n, w + 1);
}
inline void check_abs_less_than(Context& ctx, const RangeChip& range, const AssignedValue& x, const BigUint& bnd) {
    const int W = h2svd_abs_less_than_witness_count(bnd.data(), range.lookup_bits, 0);
    if (W < 0) throw std::logic_error(std::string("check_abs_less_than: ") + h2svd_last_error());
    std::vector<Fr> w((size_t)W);
    check(h2svd_abs_less_than_witness(Gpu::current().raw(), &x.v, nullptr, 1, bnd.data(), range.lookup_bits, w.data()),
          "check_abs_less_than");
    assign_check_abs_less_than(ctx, range, x, bnd, w.data());
}
// check_mat_diff (:441-459): |a[i][j] - b[i][j]| < tol for every entry -- one bulk call, cells per entry in order
inline void check_mat_diff(Context& ctx, const RangeChip& range, const AssignedMatrix& a, const AssignedMatrix& b,
                           const BigUint& tol) {
    require(a.size() == b.size(), "a.len() == b.len()");                             // :448
    require(!a.empty() && a[0].size() == b[0].size(), "a[0].len() == b[0].len()");  // :449
    const size_t rows = a.size(), cols = a[0].size();
    const int W = h2svd_abs_less_than_witness_count(tol.data(), range.lookup_bits, 1);
    if (W < 0) throw std::logic_error(std::string("check_mat_diff: ") + h2svd_last_error());
    const std::vector<Fr> fa = gather_values(a), fb = gather_values(b);
    std::vector<Fr> w(rows * cols * (size_t)W);
    check(h2svd_abs_less_than_witness(Gpu::current().raw(), fa.data(), fb.data(), rows * cols, tol.data(),
                                      range.lookup_bits, w.data()),
          "check_mat_diff");
    for (size_t i = 0; i < rows; i++)
        for (size_t j = 0; j < cols; j++) {
            const Fr* we = w.data() + (i * cols + j) * (size_t)W;
            const AssignedValue diff = range.gate.sub(ctx, Existing(a[i][j]), Existing(b[i][j]), we[0]);  // :453
            assign_check_abs_less_than(ctx, range, diff, tol, we + 1);                                    // :454
        }
}
// check_mat_id (:461-483)
inline void check_mat_id(Context& ctx, const RangeChip& range, const AssignedMatrix& a, const AssignedValue& scalar_id,
                         const BigUint& tol) {
    const AssignedValue zero = ctx.load_constant(field::zero());
    AssignedMatrix b(a.size());
    for (size_t i = 0; i < a.size(); i++)
        for (size_t j = 0; j < a[0].size(); j++) b[i].push_back(i == j ? scalar_id : zero);
    check_mat_diff(ctx, range, a, b, tol);
}
// check_mat_entries_bounded (:490-501)
inline void check_mat_entries_bounded(Context& ctx, const RangeChip& range, const AssignedMatrix& a, const BigUint& bnd) {
    if (a.empty()) return;
    const size_t rows = a.size(), cols = a[0].size();
    const int W = h2svd_abs_less_than_witness_count(bnd.data(), range.lookup_bits, 0);
    if (W < 0) throw std::logic_error(std::string("check_mat_entries_bounded: ") + h2svd_last_error());
    const std::vector<Fr> fa = gather_values(a);
    std::vector<Fr> w(rows * cols * (size_t)W);
    check(h2svd_abs_less_than_witness(Gpu::current().raw(), fa.data(), nullptr, rows * cols, bnd.data(),
                                      range.lookup_bits, w.data()),
          "check_mat_entries_bounded");
    for (size_t i = 0; i < rows; i++)
        for (size_t j = 0; j < cols; j++)
            assign_check_abs_less_than(ctx, range, a[i][j], bnd, w.data() + (i * cols + j) * (size_t)W);
}
// mat_times_diag_mat (:610-627): a * [Diag(v) 0]^T, one gate.mul per entry
inline AssignedMatrix mat_times_diag_mat(Context& ctx, const GateChip& gate, const AssignedMatrix& a,
                                         const std::vector<AssignedValue>& v) {
    require(!a.empty() && v.size() <= a[0].size(), "v.len() <= a[0].len()");  // :616
    const size_t rows = a.size(), lda = a[0].size(), cols = v.size();
    const std::vector<Fr> fa = gather_values(a), fv = gather_values(v);
    std::vector<Fr> prod(rows * cols);
    check(h2svd_mat_times_diag(Gpu::current().raw(), fa.data(), fv.data(), rows, lda, cols, prod.data()), "mat_times_diag_mat");
    AssignedMatrix m(rows);
    for (size_t i = 0; i < rows; i++)
        for (size_t j = 0; j < cols; j++) m[i].push_back(gate.mul(ctx, Existing(a[i][j]), Existing(v[j]), prod[i * cols + j]));
    return m;
}

// ---- the caller of the path: src/svd/mod.rs (check_svd_phase0 :32-116, check_svd_phase1 :127-144, err_calc :155-163) ----
namespace svd {
inline std::pair<double, double> err_calc(uint32_t p, size_t size, double max_norm, double eps_svd, double eps_u) {
    const double precision = std::pow(2.0, -1.0 * ((double)p + 1.0));
    const double err_svd = precision * (double)size * (1.0 + max_norm + eps_svd + precision) +
                           (double)size * max_norm * precision + std::pow(1.0 + eps_u, 0.5) * (max_norm + eps_svd) * eps_u +
                           std::pow(1.0 + eps_u, 0.5) * eps_svd;
    const double err_u = eps_u + precision * (double)size * (2.0 * (1.0 + eps_u) + precision);
    return {err_svd, err_u};
}
template <uint32_t P>
struct Phase0Out {
    ZkMatrix<P> u_t, v_t;
    AssignedMatrix m_times_vt, u_times_ut, v_times_vt;
};
template <uint32_t P>
Phase0Out<P> check_svd_phase0(Context& ctx, const FixedPointChip041<P>& fpchip, const ZkMatrix<P>& m, const ZkMatrix<P>& u,
                              const ZkMatrix<P>& v, const ZkVector<P>& d, double err_svd, double err_u, uint32_t max_bits_d) {
    require(m.num_rows == u.num_rows, "m.num_rows == u.num_rows");
    require(m.num_col == v.num_rows, "m.num_col == v.num_rows");
    const size_t N = m.num_rows, M = m.num_col, minNM = N < M ? N : M;
    require(u.num_rows == u.num_col && v.num_rows == v.num_col, "unitaries are square");
    require(minNM == d.v.size(), "min(N, M) == d.len()");
    const RangeChip& range = fpchip.range_gate();
    const GateChip& gate = fpchip.gate();
    const size_t max_bits = (size_t)max_bits_d + P;
    d.entries_less_than(ctx, fpchip, max_bits);
    d.entries_in_desc_order(ctx, fpchip, max_bits);
    const BigUint unit_bnd_q = biguint(((unsigned __int128)1 << P) + 1);
    check_mat_entries_bounded(ctx, range, u.matrix, unit_bnd_q);
    check_mat_entries_bounded(ctx, range, v.matrix, unit_bnd_q);
    Phase0Out<P> out;
    out.u_t = ZkMatrix<P>::transpose_matrix(u);
    out.v_t = ZkMatrix<P>::transpose_matrix(v);
    AssignedMatrix u_times_d;
    if (minNM == M) {
        u_times_d = mat_times_diag_mat(ctx, gate, u.matrix, d.v);
    } else {
        const AssignedValue zero = ctx.load_constant(field::zero());
        u_times_d = mat_times_diag_mat(ctx, gate, u.matrix, d.v);
        for (auto& row : u_times_d)
            for (size_t j = N; j < M; j++) row.push_back(zero);
    }
    out.m_times_vt = honest_prover_mat_mul(ctx, m.matrix, out.v_t.matrix);
    const double scale = std::ldexp(1.0, 2 * (int)P);
    const BigUint err_svd_scale = biguint((unsigned __int128)std::round(err_svd * scale));
    const BigUint err_u_scale = biguint((unsigned __int128)std::round(err_u * scale));
    check_mat_diff(ctx, range, u_times_d, out.m_times_vt, err_svd_scale);
    const AssignedValue quant_square = ctx.load_constant(field::pow2(2 * (int)P));
    out.u_times_ut = honest_prover_mat_mul(ctx, u.matrix, out.u_t.matrix);
    check_mat_id(ctx, range, out.u_times_ut, quant_square, err_u_scale);
    out.v_times_vt = honest_prover_mat_mul(ctx, v.matrix, out.v_t.matrix);
    check_mat_id(ctx, range, out.v_times_vt, quant_square, err_u_scale);
    return out;
}
template <uint32_t P>
void check_svd_phase1(Context& ctx, const FixedPointChip041<P>& fpchip, const ZkMatrix<P>& m, const ZkMatrix<P>& u,
                      const ZkMatrix<P>& v, const Phase0Out<P>& p0, const AssignedValue& init_rand) {
    ZkMatrix<P>::verify_mul(ctx, fpchip, m, p0.v_t, p0.m_times_vt, init_rand);
    ZkMatrix<P>::verify_mul(ctx, fpchip, u, p0.u_t, p0.u_times_ut, init_rand);
    ZkMatrix<P>::verify_mul(ctx, fpchip, v, p0.v_t, p0.v_times_vt, init_rand);
}
}  // namespace svd

// ---- MockProver-style check of recorded contexts ------------------------------------------------------------
// Gates a + b*c - d == 0 at every selected row, copy constraints, constants, lookups (< 2^lookup_bits).
// Returns human-readable failures (empty == satisfied).
inline std::vector<std::string> mock_verify(const std::vector<const Context*>& ctxs, int lookup_bits,
                                            size_t max_failures = 16) {
    std::vector<std::string> fails;
    auto fail = [&](const std::string& s) {
        if (fails.size() < max_failures) fails.push_back(s);
    };
    auto find = [&](uint32_t id) -> const Context* {
        for (const Context* c : ctxs)
            if (c->ctx_id == id) return c;
        return nullptr;
    };
    for (const Context* c : ctxs) {
        for (size_t i = 0; i < c->advice.size(); i++) {
            if (!c->selector[i]) continue;
            if (i + 3 >= c->advice.size()) {
                fail("ctx " + std::to_string(c->ctx_id) + ": gate at " + std::to_string(i) + " runs off the column");
                continue;
            }
            const Fr lhs = field::add(c->advice[i], field::mul(c->advice[i + 1], c->advice[i + 2]));
            if (lhs != c->advice[i + 3]) fail("ctx " + std::to_string(c->ctx_id) + ": gate violated at " + std::to_string(i));
        }
        for (const auto& cp : c->copies) {
            const Context* ca = find(cp.first.ctx_id);
            const Context* cb = find(cp.second.ctx_id);
            if (!ca || !cb || cp.first.index >= ca->advice.size() || cp.second.index >= cb->advice.size() ||
                ca->advice[cp.first.index] != cb->advice[cp.second.index])
                fail("ctx " + std::to_string(c->ctx_id) + ": copy constraint violated at " + std::to_string(cp.second.index));
        }
        for (const auto& k : c->constants)
            if (c->advice[k.first] != k.second)
                fail("ctx " + std::to_string(c->ctx_id) + ": constant violated at " + std::to_string(k.first));
        for (size_t cell : c->lookups) {
            const std::array<uint64_t, 4> w = field::canonical(c->advice[cell]);
            const bool ok = w[1] == 0 && w[2] == 0 && w[3] == 0 && (lookup_bits >= 64 || (w[0] >> lookup_bits) == 0);
            if (!ok) fail("ctx " + std::to_string(c->ctx_id) + ": lookup violated at " + std::to_string(cell));
        }
    }
    return fails;
}

}  // namespace zk
}  // namespace h2svd

#endif  // H2SVD_ZK_HPP
