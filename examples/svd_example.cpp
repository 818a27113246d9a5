// svd_example -- C++ counterpart of the reference's examples/svd_example.rs on the B200 witness path.
//
//   ./svd_example <name> [--data-dir DIR] [--task 1|2|3] [--gamma HEX]
//
// Reads DIR/<name>.in (the JSON {m,u,d,v} written by the reference's input-creator.py, examples/svd_example.rs:316-330),
// builds the do_zk_svd circuit (examples/svd_example.rs:98-201 -> src/svd/mod.rs) with every witness value computed by
// libh2svd_b200 (K = 20, LOOKUP_BITS = 19, PRECISION_BITS = 42 as in :68-69, :319), checks it the way MockProver would and
// prints the verdict: `matrix` verifies, `matrix-wrong` fails (README.md:93).  --task 1 / 2 run the reference's two smoke
// drivers instead (test_zkvector, test_field_mat_times_vec; src/matrix/test_matrix.rs:39, :201) and print the circuit
// values next to the f64 ground truth, like the reference.
//
// Unlike the reference this harness asserts: exit code 0 = satisfied, 1 = constraint system violated, 2 = usage / IO.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>

#include "../include/h2svd_zk.hpp"

using namespace h2svd::zk;
constexpr uint32_t PRECISION_BITS = 42;  // examples/svd_example.rs:69
constexpr int K = 20;                    // :68
constexpr int LOOKUP_BITS = K - 1;       // :319

// ---- minimal parser for the {"m": [[..]], "u": [[..]], "d": [..], "v": [[..]]} files -------------------------------
struct Json {
    std::string s;
    size_t p = 0;
    void ws() { while (p < s.size() && strchr(" \t\r\n", s[p])) p++; }
    bool eat(char c) { ws(); if (p < s.size() && s[p] == c) { p++; return true; } return false; }
    void expect(char c) { if (!eat(c)) throw std::runtime_error(std::string("JSON: expected '") + c + "' at offset " + std::to_string(p)); }
    double number() {
        ws();
        char* end = nullptr;
        const double v = std::strtod(s.c_str() + p, &end);
        if (end == s.c_str() + p) throw std::runtime_error("JSON: number expected at offset " + std::to_string(p));
        p = (size_t)(end - s.c_str());
        return v;
    }
    std::string key() {
        expect('"');
        const size_t q = s.find('"', p);
        if (q == std::string::npos) throw std::runtime_error("JSON: unterminated string");
        std::string k = s.substr(p, q - p);
        p = q + 1;
        return k;
    }
    std::vector<double> vec() {
        std::vector<double> v;
        expect('[');
        if (eat(']')) return v;
        do v.push_back(number()); while (eat(','));
        expect(']');
        return v;
    }
    std::vector<std::vector<double>> mat() {
        std::vector<std::vector<double>> m;
        expect('[');
        if (eat(']')) return m;
        do m.push_back(vec()); while (eat(','));
        expect(']');
        return m;
    }
};
struct CircuitInput {  // examples/svd_example.rs:60-66
    std::vector<double> d;
    std::vector<std::vector<double>> m, u, v;
};
static CircuitInput load_input(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("Unable to read file " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    Json j{ss.str()};
    CircuitInput in;
    j.expect('{');
    do {
        const std::string k = j.key();
        j.expect(':');
        if (k == "d") in.d = j.vec();
        else if (k == "m") in.m = j.mat();
        else if (k == "u") in.u = j.mat();
        else if (k == "v") in.v = j.mat();
        else throw std::runtime_error("JSON: unexpected key " + k);
    } while (j.eat(','));
    j.expect('}');
    if (in.m.empty() || in.u.empty() || in.v.empty() || in.d.empty()) throw std::runtime_error("JSON was not well-formatted");
    return in;
}

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static int report(const std::vector<const Context*>& ctxs) {
    size_t cells = 0, lookups = 0;
    for (const Context* c : ctxs) {
        cells += c->advice.size();
        lookups += c->lookups.size();
    }
    const std::vector<std::string> fails = mock_verify(ctxs, LOOKUP_BITS);
    std::printf("advice cells: %zu, lookup cells: %zu\n", cells, lookups);
    if (fails.empty()) {
        std::printf("constraint system satisfied: VERIFIED\n");
        return 0;
    }
    std::printf("constraint system VIOLATED (%zu shown):\n", fails.size());
    for (const std::string& f : fails) std::printf("  %s\n", f.c_str());
    return 1;
}

// do_zk_svd (examples/svd_example.rs:232): phase 0 on context 0, challenge, phase 1 on context 1
static int do_zk_svd(const CircuitInput& in, const Fr& gamma) {
    const double t0 = now_s();
    FixedPointChip041<PRECISION_BITS> fpchip(LOOKUP_BITS);
    Context ctx(0), ctx1(1);
    const ZkMatrix<PRECISION_BITS> m = ZkMatrix<PRECISION_BITS>::create(ctx, fpchip, in.m);
    const ZkMatrix<PRECISION_BITS> u = ZkMatrix<PRECISION_BITS>::create(ctx, fpchip, in.u);
    const ZkMatrix<PRECISION_BITS> v = ZkMatrix<PRECISION_BITS>::create(ctx, fpchip, in.v);
    const ZkVector<PRECISION_BITS> d = ZkVector<PRECISION_BITS>::create(ctx, fpchip, in.d);
    const size_t size = std::max(in.m.size(), in.m[0].size());
    const auto errs = svd::err_calc(PRECISION_BITS, size, 100.0, 1e-10, 1e-10);  // :115-121, :149
    std::printf("err_svd = %.3e, err_u = %.3e\n", errs.first, errs.second);
    const auto p0 = svd::check_svd_phase0(ctx, fpchip, m, u, v, d, errs.first, errs.second, 30);
    const double t1 = now_s();
    const AssignedValue init_rand = ctx1.load_witness(gamma);  // rlc.gamma_pow_cached()[0] in the reference (:183-184)
    svd::check_svd_phase1(ctx1, fpchip, m, u, v, p0, init_rand);
    const double t2 = now_s();
    std::printf("witness generation: phase 0 %.1f ms, phase 1 %.1f ms (%zu x %zu)\n", (t1 - t0) * 1e3, (t2 - t1) * 1e3,
                in.m.size(), in.m[0].size());
    return report({&ctx, &ctx1});
}

// test_zkvector (src/matrix/test_matrix.rs:39-198)
static int test_zkvector() {
    constexpr uint32_t P = 32;
    FixedPointChip041<P> fpchip(LOOKUP_BITS);
    Context ctx(0);
    const size_t N = 5, M = 4;
    std::vector<std::vector<double>> matrix(N, std::vector<double>(M));
    for (size_t i = 0; i < N; i++) for (size_t j = 0; j < M; j++) matrix[i][j] = (double)i + (double)j / 10.0;
    std::vector<double> v1, v2;
    for (size_t i = 0; i < M; i++) v1.push_back((i % 2 == 0 ? (double)i : -(double)i) + (double)(i * i + 1) / 10.0);
    for (size_t i = 0; i < M; i++) v2.push_back((i % 2 == 0 ? 1.0 : -1.0) * (1.0 + (double)(i * i * i)) / 10.0);
    const ZkMatrix<P> zkmatrix = ZkMatrix<P>::create(ctx, fpchip, matrix);
    const ZkVector<P> zkvec1 = ZkVector<P>::create(ctx, fpchip, v1), zkvec2 = ZkVector<P>::create(ctx, fpchip, v2);
    double ip = 0, n1 = 0, n2 = 0, dist = 0;
    for (size_t i = 0; i < M; i++) { ip += v1[i] * v2[i]; n1 += v1[i] * v1[i]; n2 += v2[i] * v2[i]; dist += (v1[i] - v2[i]) * (v1[i] - v2[i]); }
    auto dq = [&](const AssignedValue& x) { return fpchip.dequantization(x.v); };
    std::printf("Inner product:  f64 %.10f   zk ckt %.10f\n", ip, dq(zkvec1.inner_product(ctx, fpchip, zkvec2.v)));
    std::printf("Norm v1:        f64 %.10f   zk ckt %.10f\n", std::sqrt(n1), dq(zkvec1.norm(ctx, fpchip)));
    std::printf("Norm v2:        f64 %.10f   zk ckt %.10f\n", std::sqrt(n2), dq(zkvec2.norm(ctx, fpchip)));
    std::printf("dist:           f64 %.10f   zk ckt %.10f\n", std::sqrt(dist), dq(zkvec1.dist(ctx, fpchip, zkvec2.v)));
    std::printf("Norm-squared:   f64 %.10f %.10f   zk ckt %.10f %.10f\n", n1, n2, dq(zkvec1._norm_square(ctx, fpchip)),
                dq(zkvec2._norm_square(ctx, fpchip)));
    std::printf("dist-squared:   f64 %.10f   zk ckt %.10f\n", dist, dq(zkvec1._dist_square(ctx, fpchip, zkvec2.v)));
    const std::vector<double> u1 = zkvec1.mul(ctx, fpchip, zkmatrix).dequantize(fpchip);
    std::printf("Matrix transform of v1: zk ckt [");
    for (double x : u1) std::printf(" %.8f", x);
    std::printf(" ]\n");
    return report({&ctx});
}

// test_field_mat_times_vec (src/matrix/test_matrix.rs:201-265)
static int test_field_mat_times_vec() {
    constexpr uint32_t P = 32;
    FixedPointChip041<P> fpchip(LOOKUP_BITS);
    Context ctx(0);
    const size_t N = 5, M = 5;
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> uni(-100.0, 100.0);
    std::vector<std::vector<double>> matrix(N, std::vector<double>(M));
    std::vector<double> v1(M);
    for (auto& row : matrix) for (double& x : row) x = uni(rng);
    for (double& x : v1) x = uni(rng);
    const ZkMatrix<P> zkmatrix = ZkMatrix<P>::create(ctx, fpchip, matrix);
    const ZkVector<P> zkvec1 = ZkVector<P>::create(ctx, fpchip, v1);
    const std::vector<AssignedValue> zku1_s = field_mat_vec_mul(ctx, fpchip.gate(), zkmatrix.matrix, zkvec1.v);
    for (size_t i = 0; i < N; i++) {
        double f = 0;
        for (size_t j = 0; j < M; j++) f += matrix[i][j] * v1[j];
        std::printf("row %zu: f64 %.8f   zk ckt %.8f\n", i, f, fpchip.dequantization(fpchip.signed_div_scale(ctx, zku1_s[i]).first.v));
    }
    return report({&ctx});
}

int main(int argc, char** argv) {
    std::printf("svd_example started...\n");
    std::string name, dir = "./data";
    int task = 3;
    BigUint g = {0x0123456789abcdefull, 0x1234567890abcdefull, 0x1234567890abcdefull, 0};
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "--data-dir" && i + 1 < argc) dir = argv[++i];
        else if (a == "--task" && i + 1 < argc) task = std::atoi(argv[++i]);
        else if (a == "--gamma" && i + 1 < argc) { g = {std::strtoull(argv[++i], nullptr, 16), 0, 0, 0}; }
        else name = a;
    }
    try {
        if (task == 1) return test_zkvector();
        if (task == 2) return test_field_mat_times_vec();
        if (name.empty()) {
            std::fprintf(stderr, "Incorrect usage; use: svd_example <filename> [--data-dir DIR] [--task 1|2|3]\n");
            return 2;
        }
        const CircuitInput in = load_input(dir + "/" + name + ".in");
        std::printf("data loaded...\n");
        return do_zk_svd(in, biguint_to_fe(g));
    } catch (const std::exception& e) {
        std::fprintf(stderr, "svd_example: %s\n", e.what());
        return 2;
    }
}
