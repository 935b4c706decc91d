// ORACLE (test infrastructure only) -- dense column-major linear algebra used by the CPU restatement.
//
// Nothing under oracle/ is product code: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
// legs may build or call it.  The reference uses Eigen (absent from this image); the few Eigen routines the
// hot path relies on are restated here:
//   * Eigen::FullPivLU kernel()/solve()  as used by ocs2 LinearAlgebra::luConstraintProjection
//     (lib/ocs2_ros2/ocs2_core/src/misc/LinearAlgebra.cpp:183-199)
//   * lower Cholesky / triangular solves as used by HPIPM's Riccati recursion (see riccati.hpp)
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <stdexcept>
#include <vector>

namespace orc {

using Vec = std::vector<double>;

struct Mat {
  int r = 0, c = 0;
  std::vector<double> a;  // column-major (mirrors Eigen's default)
  Mat() = default;
  Mat(int rows, int cols, double v = 0.0) : r(rows), c(cols), a(static_cast<size_t>(rows) * cols, v) {}
  double& operator()(int i, int j) { return a[static_cast<size_t>(j) * r + i]; }
  double operator()(int i, int j) const { return a[static_cast<size_t>(j) * r + i]; }
  static Mat identity(int n) {
    Mat m(n, n);
    for (int i = 0; i < n; ++i) m(i, i) = 1.0;
    return m;
  }
  int size() const { return r * c; }
};

inline Vec vzero(int n) { return Vec(static_cast<size_t>(n), 0.0); }
inline Vec operator+(const Vec& a, const Vec& b) {
  assert(a.size() == b.size());
  Vec o(a);
  for (size_t i = 0; i < a.size(); ++i) o[i] += b[i];
  return o;
}
inline Vec operator-(const Vec& a, const Vec& b) {
  assert(a.size() == b.size());
  Vec o(a);
  for (size_t i = 0; i < a.size(); ++i) o[i] -= b[i];
  return o;
}
inline Vec operator*(double s, const Vec& a) {
  Vec o(a);
  for (auto& v : o) v *= s;
  return o;
}
inline void axpy(double s, const Vec& x, Vec& y) {
  assert(x.size() == y.size());
  for (size_t i = 0; i < x.size(); ++i) y[i] += s * x[i];
}
inline double dot(const Vec& a, const Vec& b) {
  assert(a.size() == b.size());
  double s = 0;
  for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
  return s;
}
inline double sqnorm(const Vec& a) { return dot(a, a); }

inline Mat transpose(const Mat& A) {
  Mat T(A.c, A.r);
  for (int j = 0; j < A.c; ++j)
    for (int i = 0; i < A.r; ++i) T(j, i) = A(i, j);
  return T;
}
// C = op(A) * op(B)
inline Mat mul(const Mat& A, const Mat& B, bool tA = false, bool tB = false) {
  const int m = tA ? A.c : A.r, k = tA ? A.r : A.c, k2 = tB ? B.c : B.r, n = tB ? B.r : B.c;
  assert(k == k2);
  (void)k2;
  Mat C(m, n);
  for (int j = 0; j < n; ++j)
    for (int p = 0; p < k; ++p) {
      const double b = tB ? B(j, p) : B(p, j);
      if (b == 0.0) continue;
      for (int i = 0; i < m; ++i) C(i, j) += (tA ? A(p, i) : A(i, p)) * b;
    }
  return C;
}
inline Vec mul(const Mat& A, const Vec& x, bool tA = false) {
  const int m = tA ? A.c : A.r, k = tA ? A.r : A.c;
  assert(static_cast<int>(x.size()) == k);
  Vec y(m, 0.0);
  for (int p = 0; p < k; ++p)
    for (int i = 0; i < m; ++i) y[i] += (tA ? A(p, i) : A(i, p)) * x[p];
  return y;
}
inline Mat operator+(const Mat& A, const Mat& B) {
  assert(A.r == B.r && A.c == B.c);
  Mat C(A);
  for (size_t i = 0; i < C.a.size(); ++i) C.a[i] += B.a[i];
  return C;
}
inline Mat operator-(const Mat& A, const Mat& B) {
  assert(A.r == B.r && A.c == B.c);
  Mat C(A);
  for (size_t i = 0; i < C.a.size(); ++i) C.a[i] -= B.a[i];
  return C;
}
inline Mat operator*(double s, const Mat& A) {
  Mat C(A);
  for (auto& v : C.a) v *= s;
  return C;
}
inline void addTo(Mat& A, const Mat& B, double s = 1.0) {
  assert(A.r == B.r && A.c == B.c);
  for (size_t i = 0; i < A.a.size(); ++i) A.a[i] += s * B.a[i];
}

// ---- Eigen::FullPivLU restatement (rank-revealing, complete pivoting) ---------------------------------
struct FullPivLU {
  Mat lu;                  // packed L (unit lower) and U
  std::vector<int> rowT;   // row transpositions
  std::vector<int> colT;   // column transpositions
  std::vector<int> P, Q;   // row permutation (P*A*Q = L*U): P[i] = original row placed at i ; Q[j] = original col at j
  int rank = 0;
  int rows, cols;
  double maxPivot = 0.0;

  explicit FullPivLU(const Mat& A) : lu(A), rows(A.r), cols(A.c) {
    const int size = std::min(rows, cols);
    P.resize(rows);
    Q.resize(cols);
    for (int i = 0; i < rows; ++i) P[i] = i;
    for (int j = 0; j < cols; ++j) Q[j] = j;
    int nonzeroPivots = size;
    for (int k = 0; k < size; ++k) {
      // biggest coefficient in the bottom-right corner
      int pi = k, pj = k;
      double best = -1.0;
      for (int j = k; j < cols; ++j)
        for (int i = k; i < rows; ++i)
          if (std::fabs(lu(i, j)) > best) {
            best = std::fabs(lu(i, j));
            pi = i;
            pj = j;
          }
      if (best == 0.0) {
        nonzeroPivots = k;
        break;
      }
      maxPivot = std::max(maxPivot, best);
      if (pi != k) {
        for (int j = 0; j < cols; ++j) std::swap(lu(k, j), lu(pi, j));
        std::swap(P[k], P[pi]);
      }
      if (pj != k) {
        for (int i = 0; i < rows; ++i) std::swap(lu(i, k), lu(i, pj));
        std::swap(Q[k], Q[pj]);
      }
      if (k < rows - 1)
        for (int i = k + 1; i < rows; ++i) lu(i, k) /= lu(k, k);
      if (k < size - 1 || cols > size)
        for (int j = k + 1; j < cols; ++j) {
          const double ukj = lu(k, j);
          if (ukj == 0.0) continue;
          for (int i = k + 1; i < rows; ++i) lu(i, j) -= lu(i, k) * ukj;
        }
    }
    // rank with Eigen's default threshold: eps * diagonalSize
    const double thr = 2.220446049250313e-16 * size * maxPivot;
    rank = 0;
    for (int i = 0; i < nonzeroPivots; ++i) rank += (std::fabs(lu(i, i)) > thr);
  }

  // Particular solution of A x = B with free variables set to zero (Eigen's FullPivLU::solve semantics).
  Mat solve(const Mat& B) const {
    assert(B.r == rows);
    const int smalldim = std::min(rows, cols);
    Mat c(rows, B.c);
    for (int i = 0; i < rows; ++i)
      for (int j = 0; j < B.c; ++j) c(i, j) = B(P[i], j);
    // L^-1 (unit lower, leading smalldim block; extra rows eliminated too)
    for (int j = 0; j < B.c; ++j) {
      for (int k = 0; k < smalldim; ++k)
        for (int i = k + 1; i < rows; ++i) c(i, j) -= lu(i, k) * c(k, j);
      // U^-1 on the leading rank x rank block
      for (int k = rank - 1; k >= 0; --k) {
        c(k, j) /= lu(k, k);
        for (int i = 0; i < k; ++i) c(i, j) -= lu(i, k) * c(k, j);
      }
    }
    Mat X(cols, B.c);
    for (int i = 0; i < rank; ++i)
      for (int j = 0; j < B.c; ++j) X(Q[i], j) = c(i, j);
    return X;
  }
  Vec solve(const Vec& b) const {
    Mat B(static_cast<int>(b.size()), 1);
    B.a = b;
    return solve(B).a;
  }

  // Null-space basis (cols x (cols-rank)), Eigen's FullPivLU::kernel() construction for a full-row-rank matrix:
  // for every non-pivot column k: x_pivots = -U_rr^-1 U_rk, x_k = 1, then undo the column permutation.
  Mat kernel() const {
    const int dimker = cols - rank;
    Mat K(cols, dimker);
    if (dimker == 0) return K;
    for (int kk = 0; kk < dimker; ++kk) {
      Vec y(rank);
      for (int i = 0; i < rank; ++i) y[i] = -lu(i, rank + kk);
      for (int k = rank - 1; k >= 0; --k) {
        y[k] /= lu(k, k);
        for (int i = 0; i < k; ++i) y[i] -= lu(i, k) * y[k];
      }
      for (int i = 0; i < rank; ++i) K(Q[i], kk) = y[i];
      K(Q[rank + kk], kk) = 1.0;
    }
    return K;
  }
};

// lower Cholesky, in place on the lower triangle; returns false if not positive definite
inline bool choleskyLower(Mat& A) {
  const int n = A.r;
  for (int j = 0; j < n; ++j) {
    double d = A(j, j);
    for (int k = 0; k < j; ++k) d -= A(j, k) * A(j, k);
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    A(j, j) = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A(i, j);
      for (int k = 0; k < j; ++k) s -= A(i, k) * A(j, k);
      A(i, j) = s / d;
    }
    for (int i = 0; i < j; ++i) A(i, j) = 0.0;
  }
  return true;
}
// solve L X = B (L lower) in place
inline void solveLower(const Mat& L, Mat& B) {
  for (int j = 0; j < B.c; ++j)
    for (int i = 0; i < L.r; ++i) {
      double s = B(i, j);
      for (int k = 0; k < i; ++k) s -= L(i, k) * B(k, j);
      B(i, j) = s / L(i, i);
    }
}
// solve L^T X = B in place
inline void solveLowerT(const Mat& L, Mat& B) {
  for (int j = 0; j < B.c; ++j)
    for (int i = L.r - 1; i >= 0; --i) {
      double s = B(i, j);
      for (int k = i + 1; k < L.r; ++k) s -= L(k, i) * B(k, j);
      B(i, j) = s / L(i, i);
    }
}
inline Mat asCol(const Vec& v) {
  Mat m(static_cast<int>(v.size()), 1);
  m.a = v;
  return m;
}

// dense inverse through partial-pivot Gauss-Jordan (tests only: textbook Riccati recipe uses .inverse())
inline Mat inverse(const Mat& A) {
  const int n = A.r;
  assert(A.c == n);
  Mat M(A), I = Mat::identity(n);
  for (int k = 0; k < n; ++k) {
    int p = k;
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(M(i, k)) > std::fabs(M(p, k))) p = i;
    if (M(p, k) == 0.0) throw std::runtime_error("singular matrix");
    if (p != k)
      for (int j = 0; j < n; ++j) {
        std::swap(M(k, j), M(p, j));
        std::swap(I(k, j), I(p, j));
      }
    const double d = M(k, k);
    for (int j = 0; j < n; ++j) {
      M(k, j) /= d;
      I(k, j) /= d;
    }
    for (int i = 0; i < n; ++i) {
      if (i == k) continue;
      const double f = M(i, k);
      if (f == 0.0) continue;
      for (int j = 0; j < n; ++j) {
        M(i, j) -= f * M(k, j);
        I(i, j) -= f * I(k, j);
      }
    }
  }
  return I;
}

}  // namespace orc
