// Self-test of the stand-in Eigen (oracle/refbuild/standin/Eigen/StandIn.h): the semantics the reference's hot-path code
// relies on, checked against values worked out by hand.  Built and run by tests/test_standin_eigen.py.
//   g++ -std=c++17 -I oracle/refbuild/standin oracle/refbuild/standin_selftest.cpp -o /tmp/standin_selftest && /tmp/standin_selftest
#include <cmath>
#include <cstdio>
#include <vector>

#include <Eigen/Dense>
#include <Eigen/IterativeLinearSolvers>
#include <Eigen/Sparse>
#include <Eigen/SparseLU>

static int failures = 0;
#define CHECK(cond)                                                   \
    do {                                                              \
        if (!(cond)) {                                                \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); \
            ++failures;                                               \
        }                                                             \
    } while (0)

int main() {
    using namespace Eigen;
    // ---- comma initialiser: scalars row by row, then blocks (PSPG.inl:42  Ae << A, B, C, D) -----------------------------
    Matrix<double, 2, 3> m;
    m << 1, 2, 3,
         4, 5, 6;
    CHECK(m(0, 0) == 1 && m(0, 2) == 3 && m(1, 0) == 4 && m(1, 2) == 6);
    CHECK(m.data()[1] == 4);  // column-major storage
    Matrix<double, 2, 2> A;  A << 1, 2, 3, 4;
    Matrix<double, 2, 1> B;  B << 5, 6;
    Matrix<double, 1, 2> C;  C << 7, 8;
    Matrix<double, 1, 1> D;  D << 9;
    Matrix<double, 3, 3> blk;
    blk << A, B, C, D;
    const double want[3][3] = {{1, 2, 5}, {3, 4, 6}, {7, 8, 9}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) CHECK(blk(i, j) == want[i][j]);
    Matrix<double, 3, 1> stacked;
    stacked << B, D;  // be << [12 rows], [4 rows] in the reference
    CHECK(stacked[0] == 5 && stacked[1] == 6 && stacked[2] == 9);

    // ---- products, transpose, scalar factors, 1x1 value / implicit conversion -------------------------------------------
    Matrix<double, 3, 2> mt = m.transpose();
    CHECK(mt(2, 1) == 6 && mt(0, 1) == 4);
    Matrix<double, 2, 2> mm = m * mt;  // [[14, 32], [32, 77]]
    CHECK(mm(0, 0) == 14 && mm(0, 1) == 32 && mm(1, 0) == 32 && mm(1, 1) == 77);
    Matrix<double, 2, 2> sc = 2.0 * A * 0.5 + A - A;
    CHECK(sc(1, 0) == 3 && sc(0, 1) == 2);
    Matrix<double, 2, 2> neg = -A.transpose();
    CHECK(neg(0, 1) == -3);
    CHECK((C * B).value() == 7 * 5 + 8 * 6);
    const double asDouble = std::sqrt(C * B);  // MomContEquation.inl:111: sqrt of a 1x1 product
    CHECK(std::abs(asDouble - std::sqrt(83.0)) < 1e-15);
    Matrix<double, 2, 2> acc;  acc.setZero();
    acc += A * 2.0;
    acc -= A;
    acc *= 3.0;
    CHECK(acc(1, 1) == 12);
    Matrix<double, 4, 4> big;  big.setZero();
    big.block(2, 2, 2, 2) = A;  // diagBlock, MatricesBuilder.hpp:120-129
    CHECK(big(2, 2) == 1 && big(3, 2) == 3 && big(0, 0) == 0);
    std::vector<double> raw = {1.5, 2.5, 3.5};
    Matrix<double, 3, 1> mapped = Map<Matrix<double, 3, 1>>(raw.data(), raw.size());
    CHECK(mapped[2] == 3.5);
    DiagonalMatrix<double, Dynamic> dg;  dg.resize(3);  dg.setZero();
    dg.diagonal()[0] = 2;  dg.diagonal()[1] = 4;  dg.diagonal()[2] = 8;
    VectorXd xv(3);  xv[0] = 1;  xv[1] = 1;  xv[2] = 0.5;
    VectorXd dx = dg * xv;
    CHECK(dx[0] == 2 && dx[1] == 4 && dx[2] == 4 && std::abs(dx.norm() - 6.0) < 1e-15);

    // ---- Triplet default, setFromTriplets: duplicates summed in triplet order, sorted rows, explicit zeros kept -----------
    Triplet<double> def;
    CHECK(def.row() == 0 && def.col() == 0 && def.value() == 0.0);
    std::vector<Triplet<double>> T;
    T.emplace_back(2, 1, 1e16);   // (2,1): 1e16 + 1 - 1e16 in THIS order = 0 in fp64, any other order gives 1 or 0 differently
    T.emplace_back(0, 1, 3.0);
    T.emplace_back(2, 1, 1.0);
    T.emplace_back(2, 1, -1e16);
    T.emplace_back(1, 0, 0.0);    // an explicit zero
    T.push_back(def);             // the (0,0,0.0) filler of masked slots (PSPG.inl:16)
    T.emplace_back(0, 0, 5.0);
    SparseMatrix<double> S(3, 3);
    S.setFromTriplets(T.begin(), T.end());
    CHECK(S.nonZeros() == 4);
    CHECK(S.outerIndexPtr()[0] == 0 && S.outerIndexPtr()[1] == 2 && S.outerIndexPtr()[2] == 4 && S.outerIndexPtr()[3] == 4);
    CHECK(S.innerIndexPtr()[0] == 0 && S.innerIndexPtr()[1] == 1 && S.innerIndexPtr()[2] == 0 && S.innerIndexPtr()[3] == 2);
    CHECK(S.valuePtr()[0] == 5.0 && S.valuePtr()[1] == 0.0 && S.valuePtr()[2] == 3.0);
    CHECK(S.valuePtr()[3] == (1e16 + 1.0) - 1e16);  // summed left to right in triplet order
    int visited = 0;
    for (SparseMatrix<double>::InnerIterator it(S, 1); it; ++it) {  // column walk of m_applyBCPSPG (PSPG.inl:219-228)
        CHECK(it.col() == 1);
        if (it.row() == 0) it.valueRef() = 0;  // zeroed, not removed
        ++visited;
    }
    S.makeCompressed();
    CHECK(visited == 2 && S.nonZeros() == 4 && S.valuePtr()[2] == 0.0);
    VectorXd ones(3);  ones[0] = ones[1] = ones[2] = 1;
    VectorXd Sy = S * ones;
    CHECK(Sy[0] == 5.0 && Sy[1] == 0.0);

    // ---- SparseLU / ConjugateGradient on a small SPD system -----------------------------------------------------------------
    std::vector<Triplet<double>> L;
    const int n = 6;
    for (int i = 0; i < n; ++i) {
        L.emplace_back(i, i, 2.5);
        if (i + 1 < n) {
            L.emplace_back(i, i + 1, -1.0);
            L.emplace_back(i + 1, i, -1.0);
        }
    }
    SparseMatrix<double> K(n, n);
    K.setFromTriplets(L.begin(), L.end());
    VectorXd xs(n), rhs;
    for (int i = 0; i < n; ++i) xs[i] = 1.0 + 0.25 * i;
    rhs = K * xs;
    SparseLU<SparseMatrix<double>, COLAMDOrdering<int>> lu;
    lu.analyzePattern(K);
    lu.factorize(K);
    CHECK(lu.info() == Success);
    VectorXd x1 = lu.solve(rhs);
    ConjugateGradient<SparseMatrix<double>, Lower | Upper> cg;
    cg.compute(K);
    VectorXd x2 = cg.solve(rhs), x3 = cg.solveWithGuess(rhs, xs);
    for (int i = 0; i < n; ++i) CHECK(std::abs(x1[i] - xs[i]) < 1e-13 && std::abs(x2[i] - xs[i]) < 1e-12 && std::abs(x3[i] - xs[i]) < 1e-13);
    CHECK(cg.info() == Success && cg.iterations() <= 2 * n);

    std::printf(failures ? "standin selftest: %d failure(s)\n" : "standin selftest: ok\n", failures);
    return failures ? 1 : 0;
}
