// Host-side derivative check of lm_block.cuh: the analytic eval_block against the dual-number eval_block_dual on
// random blocks and lines (values must agree to the last bits, Jacobians to 1e-9 relative). No GPU needed:
//   nvcc -std=c++17 -O2 -o /tmp/lm_block_check scripts/lm_block_check.cu && /tmp/lm_block_check
#include "../limap_b200/csrc/lm_block.cuh"
#include <cstdio>
#include <random>

using namespace lm;

int main() {
  std::mt19937_64 rng(7);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  double worst_r = 0, worst_j = 0, worst_v = 0;
  int n_clamped = 0;
  for (int it = 0; it < 200000; ++it) {
    // a line in minimal form
    double x[6];
    double qn = 0;
    for (int i = 0; i < 4; ++i) { x[i] = U(rng); qn += x[i] * x[i]; }
    for (int i = 0; i < 4; ++i) x[i] /= std::sqrt(qn);
    const double ang = U(rng) * 1.5;
    x[4] = std::cos(ang); x[5] = std::sin(ang);
    LineLocal LL;
    line_from_minimal(x, true, LL);
    LineShared L;
    for (int i = 0; i < 3; ++i) {
      L.d[i] = LL.d[i].a; L.m[i] = LL.m[i].a;
      for (int c = 0; c < 4; ++c) { L.dv[i][c] = LL.d[i].v[c]; L.mv[i][c] = LL.m[i].v[c]; }
    }
    for (int col = 0; col < 4; ++col) { // the one-column form is the 4-wide form, bit for bit
      Dual<1> d1[3], m1[3];
      line_from_minimal_col(x, col, true, d1, m1);
      for (int i = 0; i < 3; ++i)
        if (d1[i].a != LL.d[i].a || m1[i].a != LL.m[i].a || d1[i].v[0] != LL.d[i].v[col] || m1[i].v[0] != LL.m[i].v[col]) {
          std::printf("line_from_minimal_col differs (col %d)\n", col);
          return 1;
        }
    }
    LMBlockDev B;
    B.k[0] = 600 + 200 * U(rng); B.k[1] = B.k[0] * (1 + 0.1 * U(rng)); B.k[2] = 400 + 30 * U(rng); B.k[3] = 300 + 30 * U(rng);
    double q[4], n2 = 0;
    for (int i = 0; i < 4; ++i) { q[i] = U(rng); n2 += q[i] * q[i]; }
    const double a = q[0], b = q[1], c = q[2], d = q[3], nr = 1.0 / n2;
    B.R[0] = (a * a + b * b - c * c - d * d) * nr; B.R[1] = 2 * (b * c - a * d) * nr; B.R[2] = 2 * (a * c + b * d) * nr;
    B.R[3] = 2 * (a * d + b * c) * nr; B.R[4] = (a * a - b * b + c * c - d * d) * nr; B.R[5] = 2 * (c * d - a * b) * nr;
    B.R[6] = 2 * (b * d - a * c) * nr; B.R[7] = 2 * (a * b + c * d) * nr; B.R[8] = (a * a - b * b - c * c + d * d) * nr;
    for (int i = 0; i < 3; ++i) B.t[i] = 5 * U(rng);
    for (int i = 0; i < 4; ++i) B.p[i] = (i % 2 ? 300 : 400) * (1 + U(rng));
    B.w = 1.0;
    for (int i = 0; i < 3; ++i) B.vdir[i] = U(rng);
    B.wvp = (it % 3 == 0) ? 0.5 : 0.0;
    BlockEval ea, ed;
    eval_block(B, L, 10.0, true, ea);
    eval_block_dual(B, L, 10.0, true, ed);
    for (int k = 0; k < 2; ++k) worst_r = std::fmax(worst_r, std::fabs(ea.r[k] - ed.r[k]) / (1e-300 + std::fabs(ed.r[k])));
    worst_v = std::fmax(worst_v, std::fabs(ea.rv - ed.rv));
    double sc = 0;
    for (int k = 0; k < 8; ++k) sc = std::fmax(sc, std::fabs(ed.J[k]));
    for (int k = 0; k < 8; ++k) worst_j = std::fmax(worst_j, std::fabs(ea.J[k] - ed.J[k]) / (1e-300 + sc));
    if (B.wvp > 0) {
      double sv = 1e-12;
      for (int k = 0; k < 4; ++k) sv = std::fmax(sv, std::fabs(ed.Jv[k]));
      for (int k = 0; k < 4; ++k) worst_j = std::fmax(worst_j, std::fabs(ea.Jv[k] - ed.Jv[k]) / sv);
    }
  }
  std::printf("max rel residual diff %.3e, max rel Jacobian diff %.3e, max VP residual diff %.3e\n", worst_r, worst_j, worst_v);
  const bool ok = worst_r <= 1e-13 && worst_j <= 1e-9 && worst_v <= 1e-15;
  std::printf(ok ? "OK\n" : "FAILED\n");
  return ok ? 0 : 1;
}
