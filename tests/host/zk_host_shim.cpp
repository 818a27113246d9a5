// Test driver for include/h2svd_zk.hpp: runs the reference's usage scenarios through the C++ host
// mirror and exports the recorded halo2-base contexts so pytest can diff them cell by cell against
// the oracle's model (oracle/pyoracle.py).  Linked against libh2svd_b200.so for the GPU tests and
// against tests/host/abi_over_oracle.c for the CPU-only tests of the bookkeeping.
#include <cstring>
#include <memory>

#include "../../include/h2svd_zk.hpp"

using namespace h2svd::zk;

static std::vector<Context> g_ctx;
static std::vector<std::string> g_fail;
static std::string g_error;
static std::vector<double> g_scalars;  // dequantized results a scenario wants to report

template <uint32_t P>
static void run_zkmatrix(int lb, const double* a, const double* b, size_t n, size_t k, size_t m, const Fr& gamma) {
    // README.md:34-47 usage: phase 0 = honest_prover_mat_mul + rescale_matrix, phase 1 = verify_mul
    FixedPointChip041<P> fpchip(lb);
    g_ctx.emplace_back(0);
    g_ctx.emplace_back(1);
    Context& ctx = g_ctx[0];
    std::vector<std::vector<double>> am(n, std::vector<double>(k)), bm(k, std::vector<double>(m));
    for (size_t i = 0; i < n; i++) for (size_t j = 0; j < k; j++) am[i][j] = a[i * k + j];
    for (size_t i = 0; i < k; i++) for (size_t j = 0; j < m; j++) bm[i][j] = b[i * m + j];
    const ZkMatrix<P> za = ZkMatrix<P>::create(ctx, fpchip, am);
    const ZkMatrix<P> zb = ZkMatrix<P>::create(ctx, fpchip, bm);
    const AssignedMatrix c_s = honest_prover_mat_mul(ctx, za.matrix, zb.matrix);
    const ZkMatrix<P> c = ZkMatrix<P>::rescale_matrix(ctx, fpchip, c_s);
    Context& ctx1 = g_ctx[1];
    const AssignedValue init_rand = ctx1.load_witness(gamma);
    ZkMatrix<P>::verify_mul(ctx1, fpchip, za, zb, c_s, init_rand);
    for (const auto& row : c.dequantize(fpchip)) g_scalars.insert(g_scalars.end(), row.begin(), row.end());
}

// rescale_matrix through the per-cell path (ctx 0) and through the bulk hand-off (ctx 1) on the same quantized product
// (needs the real library: the CPU-only build links the C ABI over the oracle, which has no cell-layout functions)
#ifdef ZKH_WITH_EXPAND
template <uint32_t P>
static void run_rescale_bulk(int lb, const double* a, const double* b, size_t n, size_t k, size_t m) {
    FixedPointChip041<P> fpchip(lb);
    g_ctx.emplace_back(0);
    g_ctx.emplace_back(0);
    std::vector<std::vector<double>> am(n, std::vector<double>(k)), bm(k, std::vector<double>(m));
    for (size_t i = 0; i < n; i++) for (size_t j = 0; j < k; j++) am[i][j] = a[i * k + j];
    for (size_t i = 0; i < k; i++) for (size_t j = 0; j < m; j++) bm[i][j] = b[i * m + j];
    for (int which = 0; which < 2; which++) {
        Context& ctx = g_ctx[which];
        const ZkMatrix<P> za = ZkMatrix<P>::create(ctx, fpchip, am);
        const ZkMatrix<P> zb = ZkMatrix<P>::create(ctx, fpchip, bm);
        const AssignedMatrix c_s = honest_prover_mat_mul(ctx, za.matrix, zb.matrix);
        const ZkMatrix<P> c = which == 0 ? ZkMatrix<P>::rescale_matrix(ctx, fpchip, c_s)
                                         : ZkMatrix<P>::rescale_matrix_bulk(ctx, fpchip, c_s);
        for (const auto& row : c.matrix)
            for (const AssignedValue& x : row) g_scalars.push_back((double)x.index);   // the returned cells must be the same cells
    }
}
#endif

// src/matrix/test_matrix.rs:39-198 (test_zkvector), same inputs and call order, P = 32
static void run_zkvector(int lb) {
    constexpr uint32_t P = 32;
    FixedPointChip041<P> fpchip(lb);
    g_ctx.emplace_back(0);
    Context& ctx = g_ctx[0];
    const size_t N = 5, M = 4;
    std::vector<std::vector<double>> matrix(N, std::vector<double>(M));
    for (size_t i = 0; i < N; i++) for (size_t j = 0; j < M; j++) matrix[i][j] = (double)i + (double)j / 10.0;
    const ZkMatrix<P> zkmatrix = ZkMatrix<P>::create(ctx, fpchip, matrix);
    std::vector<double> v1, v2;
    for (size_t i = 0; i < M; i++) v1.push_back((i % 2 == 0 ? (double)i : -(double)i) + (double)(i * i + 1) / 10.0);
    for (size_t i = 0; i < M; i++) v2.push_back((i % 2 == 0 ? 1.0 : -1.0) * (1.0 + (double)(i * i * i)) / 10.0);
    const ZkVector<P> zkvec1 = ZkVector<P>::create(ctx, fpchip, v1);
    const ZkVector<P> zkvec2 = ZkVector<P>::create(ctx, fpchip, v2);
    auto dq = [&](const AssignedValue& x) { g_scalars.push_back(fpchip.dequantization(x.v)); };
    dq(zkvec1.inner_product(ctx, fpchip, zkvec2.v));
    dq(zkvec1.norm(ctx, fpchip));
    dq(zkvec2.norm(ctx, fpchip));
    dq(zkvec1.dist(ctx, fpchip, zkvec2.v));
    dq(zkvec1._norm_square(ctx, fpchip));
    dq(zkvec2._norm_square(ctx, fpchip));
    dq(zkvec1._dist_square(ctx, fpchip, zkvec2.v));
    for (double x : zkvec1.mul(ctx, fpchip, zkmatrix).dequantize(fpchip)) g_scalars.push_back(x);
    for (double x : zkvec2.mul(ctx, fpchip, zkmatrix).dequantize(fpchip)) g_scalars.push_back(x);
}

// src/matrix/test_matrix.rs:201-265 (test_field_mat_times_vec) with caller-supplied inputs, P = 32
static void run_mat_times_vec(int lb, const double* mat, const double* vec, size_t n, size_t m) {
    constexpr uint32_t P = 32;
    FixedPointChip041<P> fpchip(lb);
    g_ctx.emplace_back(0);
    Context& ctx = g_ctx[0];
    std::vector<std::vector<double>> matrix(n, std::vector<double>(m));
    for (size_t i = 0; i < n; i++) for (size_t j = 0; j < m; j++) matrix[i][j] = mat[i * m + j];
    const ZkMatrix<P> zkmatrix = ZkMatrix<P>::create(ctx, fpchip, matrix);
    const ZkVector<P> zkvec1 = ZkVector<P>::create(ctx, fpchip, std::vector<double>(vec, vec + m));
    const std::vector<AssignedValue> zku1_s = field_mat_vec_mul(ctx, fpchip.gate(), zkmatrix.matrix, zkvec1.v);
    for (const AssignedValue& x : zku1_s) g_scalars.push_back(fpchip.dequantization(fpchip.signed_div_scale(ctx, x).first.v));
}

// do_zk_svd's circuit (examples/svd_example.rs:98-201 -> src/svd/mod.rs): phase 0 on ctx 0, phase 1 on ctx 1
template <uint32_t P>
static void run_svd(int lb, const double* m, const double* u, const double* v, const double* d, size_t n, size_t mm,
                    size_t err_size, const Fr& gamma) {
    FixedPointChip041<P> fpchip(lb);
    g_ctx.emplace_back(0);
    g_ctx.emplace_back(1);
    Context& ctx = g_ctx[0];
    auto mat = [](const double* p, size_t r, size_t c) {
        std::vector<std::vector<double>> o(r, std::vector<double>(c));
        for (size_t i = 0; i < r; i++) for (size_t j = 0; j < c; j++) o[i][j] = p[i * c + j];
        return o;
    };
    const ZkMatrix<P> zm = ZkMatrix<P>::create(ctx, fpchip, mat(m, n, mm));
    const ZkMatrix<P> zu = ZkMatrix<P>::create(ctx, fpchip, mat(u, n, n));
    const ZkMatrix<P> zv = ZkMatrix<P>::create(ctx, fpchip, mat(v, mm, mm));
    const ZkVector<P> zd = ZkVector<P>::create(ctx, fpchip, std::vector<double>(d, d + (n < mm ? n : mm)));
    const auto errs = svd::err_calc(P, err_size, 100.0, 1e-10, 1e-10);
    const auto p0 = svd::check_svd_phase0(ctx, fpchip, zm, zu, zv, zd, errs.first, errs.second, 30);
    Context& ctx1 = g_ctx[1];
    const AssignedValue init_rand = ctx1.load_witness(gamma);
    svd::check_svd_phase1(ctx1, fpchip, zm, zu, zv, p0, init_rand);
}

template <typename F>
static int guarded(int lb, F&& f) {
    g_ctx.clear();
    g_fail.clear();
    g_scalars.clear();
    g_error.clear();
    try {
        f();
        std::vector<const Context*> ptrs;
        for (const Context& c : g_ctx) ptrs.push_back(&c);
        g_fail = mock_verify(ptrs, lb);
        return (int)g_fail.size();
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1;
    }
}

extern "C" {
int zkh_run_zkmatrix(int P, int lb, const double* a, const double* b, size_t n, size_t k, size_t m, const uint64_t* gamma) {
    Fr g;
    std::memcpy(g.l, gamma, 32);
    return guarded(lb, [&] {
        switch (P) {
            case 32: run_zkmatrix<32>(lb, a, b, n, k, m, g); break;
            case 42: run_zkmatrix<42>(lb, a, b, n, k, m, g); break;
            case 63: run_zkmatrix<63>(lb, a, b, n, k, m, g); break;
            default: throw std::logic_error("unsupported PRECISION_BITS in the test driver");
        }
    });
}
int zkh_run_svd(int P, int lb, const double* m, const double* u, const double* v, const double* d, size_t n, size_t mm,
                size_t err_size, const uint64_t* gamma) {
    Fr g;
    std::memcpy(g.l, gamma, 32);
    return guarded(lb, [&] {
        switch (P) {
            case 32: run_svd<32>(lb, m, u, v, d, n, mm, err_size, g); break;
            case 42: run_svd<42>(lb, m, u, v, d, n, mm, err_size, g); break;
            case 63: run_svd<63>(lb, m, u, v, d, n, mm, err_size, g); break;
            default: throw std::logic_error("unsupported PRECISION_BITS in the test driver");
        }
    });
}
#ifdef ZKH_WITH_EXPAND
int zkh_run_rescale_bulk(int P, int lb, const double* a, const double* b, size_t n, size_t k, size_t m) {
    return guarded(lb, [&] {
        switch (P) {
            case 32: run_rescale_bulk<32>(lb, a, b, n, k, m); break;
            case 42: run_rescale_bulk<42>(lb, a, b, n, k, m); break;
            case 63: run_rescale_bulk<63>(lb, a, b, n, k, m); break;
            default: throw std::logic_error("unsupported PRECISION_BITS in the test driver");
        }
    });
}
#endif
int zkh_run_zkvector(int lb) { return guarded(lb, [&] { run_zkvector(lb); }); }
int zkh_run_mat_times_vec(int lb, const double* mat, const double* vec, size_t n, size_t m) {
    return guarded(lb, [&] { run_mat_times_vec(lb, mat, vec, n, m); });
}
// shape-assert behaviour: verify_mul with mismatching shapes must throw (the reference panics, :307-310)
int zkh_run_bad_shapes(int lb) {
    return guarded(lb, [&] {
        FixedPointChip041<32> fpchip(lb);
        g_ctx.emplace_back(0);
        Context& ctx = g_ctx[0];
        const ZkMatrix<32> a = ZkMatrix<32>::create(ctx, fpchip, {{1.0, 2.0}, {3.0, 4.0}});
        const ZkMatrix<32> b = ZkMatrix<32>::create(ctx, fpchip, {{1.0, 2.0, 3.0}});
        honest_prover_mat_mul(ctx, a.matrix, b.matrix);
    });
}
// A dishonest prover: c_s[0][0] is off by one.  The reference's verify_mul accepts it (the is_equal result is
// discarded, src/matrix/mod.rs:339-341); verify_mul_strict (this library's opt-in fix) rejects it.
int zkh_run_dishonest_product(int lb, int strict) {
    return guarded(lb, [&] {
        FixedPointChip041<32> fpchip(lb);
        g_ctx.emplace_back(0);
        g_ctx.emplace_back(1);
        Context& ctx = g_ctx[0];
        const ZkMatrix<32> a = ZkMatrix<32>::create(ctx, fpchip, {{1.5, -2.0, 0.25}, {3.0, 4.5, -1.0}});
        const ZkMatrix<32> b = ZkMatrix<32>::create(ctx, fpchip, {{0.5, 1.0}, {-1.5, 2.0}, {2.5, -0.75}});
        std::vector<std::vector<Fr>> c = field_mat_mul(a.matrix, b.matrix);
        c[0][0] = field::add(c[0][0], field::one());
        AssignedMatrix c_s;
        for (const auto& row : c) c_s.push_back(ctx.assign_witnesses(row.data(), row.size()));
        Context& ctx1 = g_ctx[1];
        const AssignedValue init_rand = ctx1.load_witness(field::from_u64(0x9e3779b97f4a7c15ull));
        if (strict) ZkMatrix<32>::verify_mul_strict(ctx1, fpchip, a, b, c_s, init_rand);
        else ZkMatrix<32>::verify_mul(ctx1, fpchip, a, b, c_s, init_rand);
    });
}
const char* zkh_error() { return g_error.c_str(); }
const char* zkh_failure(int i) { return i < (int)g_fail.size() ? g_fail[i].c_str() : ""; }
size_t zkh_ctx_count() { return g_ctx.size(); }
size_t zkh_ctx_len(size_t c) { return g_ctx[c].advice.size(); }
size_t zkh_ctx_ncopies(size_t c) { return g_ctx[c].copies.size(); }
size_t zkh_ctx_nconstants(size_t c) { return g_ctx[c].constants.size(); }
size_t zkh_ctx_nlookups(size_t c) { return g_ctx[c].lookups.size(); }
size_t zkh_nscalars() { return g_scalars.size(); }
void zkh_scalars(double* out) { std::memcpy(out, g_scalars.data(), g_scalars.size() * sizeof(double)); }
void zkh_ctx_export(size_t c, uint64_t* advice, uint8_t* kind, uint8_t* selector, uint64_t* copies, uint64_t* const_idx,
                    uint64_t* const_val, uint64_t* lookups) {
    const Context& x = g_ctx[c];
    std::memcpy(advice, x.advice.data(), x.advice.size() * sizeof(Fr));
    std::memcpy(kind, x.kind.data(), x.kind.size());
    std::memcpy(selector, x.selector.data(), x.selector.size());
    for (size_t i = 0; i < x.copies.size(); i++) {
        copies[4 * i] = x.copies[i].first.ctx_id;
        copies[4 * i + 1] = x.copies[i].first.index;
        copies[4 * i + 2] = x.copies[i].second.ctx_id;
        copies[4 * i + 3] = x.copies[i].second.index;
    }
    for (size_t i = 0; i < x.constants.size(); i++) {
        const_idx[i] = x.constants[i].first;
        std::memcpy(const_val + 4 * i, x.constants[i].second.l, 32);
    }
    for (size_t i = 0; i < x.lookups.size(); i++) lookups[i] = x.lookups[i];
}
}
