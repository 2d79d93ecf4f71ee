// TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
//
// mini_eigen: the smallest dense-matrix vocabulary that lets the UNMODIFIED reference
// translation units
//     controller/src/controller/{mppi,rk4}.cpp, rigid2d/src/rigid2d/utilities.cpp,
//     bmapping/src/bmapping/particle_filter.cpp
// compile in a container that has no Eigen3.  It implements only the
// API surface those files touch (mppi.cpp:59-60,75-79,88-93,115-121,134; rk4.cpp:57-65,99-114;
// particle_filter.cpp:27-59,109-118,572-598) with eager evaluation.  Element-wise expressions
// are arithmetic-identical to Eigen's lazy ones; reductions (sum, dot, minCoeff, products) run
// in plain index order where real Eigen may use packet-wise partial sums - a <=1e-15 relative
// difference that every parity tolerance in tests/ absorbs.
//
// This is our own code, not a copy of Eigen: there is one view type over strided storage, one
// owning type, and free operators that always return owning temporaries.
#ifndef B2N_ORACLE_MINI_EIGEN_HPP
#define B2N_ORACLE_MINI_EIGEN_HPP

#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <vector>

namespace Eigen
{
typedef std::ptrdiff_t Index;

class Mat;

// Non-owning strided window onto doubles: element (i,j) lives at p[i*rs + j*cs].
class View
{
public:
  View() : p_(nullptr), r_(0), c_(0), rs_(1), cs_(0) {}
  View(double *p, Index r, Index c, Index rs, Index cs) : p_(p), r_(r), c_(c), rs_(rs), cs_(cs) {}

  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Index size() const { return r_ * c_; }

  double &operator()(Index i, Index j) const { return p_[i * rs_ + j * cs_]; }
  // flat access for vectors (row or column shaped) and 1x1 results
  double &operator()(Index k) const { return (c_ == 1) ? p_[k * rs_] : p_[k * cs_]; }

  View row(Index i) const { return View(p_ + i * rs_, 1, c_, rs_, cs_); }
  View col(Index j) const { return View(p_ + j * cs_, r_, 1, rs_, cs_); }
  View leftCols(Index n) const { return View(p_, r_, n, rs_, cs_); }
  View rightCols(Index n) const { return View(p_ + (c_ - n) * cs_, r_, n, rs_, cs_); }
  View transpose() const { return View(p_, c_, r_, cs_, rs_); }
  View array() const { return *this; }

  // element-wise copy; column-major traversal, ascending, so a left shift of columns onto
  // themselves (mppi.cpp:134) behaves as a true shift.
  View &operator=(const View &o) { assignFrom(o); return *this; }
  void assignFrom(const View &o) const
  {
    if (o.r_ == r_ && o.c_ == c_) {
      for (Index j = 0; j < c_; j++) for (Index i = 0; i < r_; i++) (*this)(i, j) = o(i, j);
    } else if (o.size() == size() && (r_ == 1 || c_ == 1) && (o.r_ == 1 || o.c_ == 1)) {
      for (Index k = 0; k < size(); k++) (*this)(k) = o(k);   // vector <- transposed vector
    } else {
      throw std::invalid_argument("mini_eigen: shape mismatch in assignment");
    }
  }

  const View &operator+=(const View &o) const { for (Index j = 0; j < c_; j++) for (Index i = 0; i < r_; i++) (*this)(i, j) += o(i, j); return *this; }
  const View &operator-=(const View &o) const { for (Index j = 0; j < c_; j++) for (Index i = 0; i < r_; i++) (*this)(i, j) -= o(i, j); return *this; }
  const View &operator-=(double s) const { for (Index j = 0; j < c_; j++) for (Index i = 0; i < r_; i++) (*this)(i, j) -= s; return *this; }
  const View &operator+=(double s) const { for (Index j = 0; j < c_; j++) for (Index i = 0; i < r_; i++) (*this)(i, j) += s; return *this; }
  const View &operator/=(double s) const { for (Index j = 0; j < c_; j++) for (Index i = 0; i < r_; i++) (*this)(i, j) /= s; return *this; }
  const View &operator*=(double s) const { for (Index j = 0; j < c_; j++) for (Index i = 0; i < r_; i++) (*this)(i, j) *= s; return *this; }

  double minCoeff() const
  {
    double m = (*this)(0, 0);
    for (Index j = 0; j < c_; j++) for (Index i = 0; i < r_; i++) { const double v = (*this)(i, j); if (v < m) m = v; }
    return m;
  }
  double sum() const
  {
    double s = 0.0;
    for (Index j = 0; j < c_; j++) for (Index i = 0; i < r_; i++) s += (*this)(i, j);
    return s;
  }
  double dot(const View &o) const
  {
    double s = 0.0;
    for (Index k = 0; k < size(); k++) s += (*this)(k) * o(k);
    return s;
  }

  // comma initialiser: v << a, b, c;
  struct Comma
  {
    const View *v; Index k;
    Comma &operator,(double x) { (*v)(k / v->c_, k % v->c_) = x; k++; return *this; }
  };
  Comma operator<<(double x) const { Comma c{this, 0}; c, x; return c; }

  struct LLT;
  LLT llt() const;

protected:
  double *p_;
  Index r_, c_, rs_, cs_;
};

// Owning column-major matrix; MatrixXd, VectorXd and the fixed-size vectors are all this type.
class Mat : public View
{
public:
  Mat() {}
  explicit Mat(Index n) { alloc(n, 1); }
  Mat(Index r, Index c) { alloc(r, c); }
  Mat(const Mat &o) : View() { alloc(o.rows(), o.cols()); View::assignFrom(o); }
  Mat(const View &o) { alloc(o.rows(), o.cols()); View::assignFrom(o); }
  Mat(Mat &&o) noexcept : View() { steal(o); }
  Mat &operator=(const Mat &o) { if (this != &o) { Mat t(static_cast<const View &>(o)); steal(t); } return *this; }
  Mat &operator=(Mat &&o) noexcept { if (this != &o) steal(o); return *this; }
  Mat &operator=(const View &o) { Mat t(o); steal(t); return *this; }

  static Mat Zero(Index r, Index c) { Mat m(r, c); return m; }
  static Mat Zero(Index n) { Mat m(n, 1); return m; }
  static Mat Constant(Index r, Index c, double v) { Mat m(r, c); for (auto &x : m.d_) x = v; return m; }

private:
  void alloc(Index r, Index c) { d_.assign(static_cast<size_t>(r * c), 0.0); seat(r, c); }
  void seat(Index r, Index c) { p_ = d_.data(); r_ = r; c_ = c; rs_ = 1; cs_ = r; }
  void steal(Mat &o) { const Index r = o.r_, c = o.c_; d_ = std::move(o.d_); seat(r, c); o.seat(0, 0); }
  std::vector<double> d_;
};

typedef Mat MatrixXd;
typedef Mat VectorXd;

template <int N>
struct FixedVec : public Mat
{
  FixedVec() : Mat(N, 1) {}
  FixedVec(double a, double b) : Mat(N, 1) { static_assert(N == 2, "2 coefficients"); (*this)(0) = a; (*this)(1) = b; }
  FixedVec(double a, double b, double c) : Mat(N, 1) { static_assert(N == 3, "3 coefficients"); (*this)(0) = a; (*this)(1) = b; (*this)(2) = c; }
  FixedVec(const View &o) : Mat(o) {}
  FixedVec &operator=(const View &o) { Mat::operator=(o); return *this; }
};
typedef FixedVec<2> Vector2d;
typedef FixedVec<3> Vector3d;

// Eigen::Ref<T>: a shallow window; constructible from anything viewable.
template <class T>
class Ref : public View
{
public:
  Ref(const View &v) : View(v) {}
  Ref &operator=(const View &o) { View::assignFrom(o); return *this; }
  Ref &operator=(const Ref &o) { View::assignFrom(o); return *this; }
  Ref(const Ref &) = default;
};

// ---- free operators: always evaluate into an owning temporary -------------------------------
inline Mat operator+(const View &a, const View &b) { Mat m(a); m += b; return m; }
inline Mat operator-(const View &a, const View &b) { Mat m(a); m -= b; return m; }
inline Mat operator*(const View &a, double s) { Mat m(a); m *= s; return m; }
inline Mat operator*(double s, const View &a)
{
  Mat m(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); j++) for (Index i = 0; i < a.rows(); i++) m(i, j) = s * a(i, j);
  return m;
}
inline Mat operator/(const View &a, double s) { Mat m(a); m /= s; return m; }
inline Mat operator+(const View &a, double s) { Mat m(a); m += s; return m; }
// matrix product, inner index ascending
inline Mat operator*(const View &a, const View &b)
{
  if (a.cols() != b.rows()) throw std::invalid_argument("mini_eigen: product shape mismatch");
  Mat m(a.rows(), b.cols());
  for (Index j = 0; j < b.cols(); j++)
    for (Index i = 0; i < a.rows(); i++) {
      double s = 0.0;
      for (Index k = 0; k < a.cols(); k++) s += a(i, k) * b(k, j);
      m(i, j) = s;
    }
  return m;
}
inline Mat exp(const View &a)
{
  Mat m(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); j++) for (Index i = 0; i < a.rows(); i++) m(i, j) = std::exp(a(i, j));
  return m;
}

// Unblocked lower Cholesky, column by column: diagonal = sqrt(a_kk - |row k of L|^2), the
// sub-column is reduced by the already-known columns and then DIVIDED by the diagonal.
struct View::LLT
{
  Mat l;
  Mat matrixL() const { return l; }
};
inline View::LLT View::llt() const
{
  const Index n = r_;
  Mat a(*this);
  for (Index k = 0; k < n; k++) {
    double x = a(k, k);
    for (Index q = 0; q < k; q++) x -= a(k, q) * a(k, q);
    if (x <= 0.0) break;            // not positive definite: leave the rest as is (unguarded upstream)
    x = std::sqrt(x);
    a(k, k) = x;
    for (Index i = k + 1; i < n; i++) {
      double s = 0.0;
      for (Index q = 0; q < k; q++) s += a(i, q) * a(k, q);
      a(i, k) = (a(i, k) - s) / x;
    }
  }
  for (Index j = 0; j < n; j++) for (Index i = 0; i < j; i++) a(i, j) = 0.0;
  LLT f; f.l = a; return f;
}

} // namespace Eigen
#endif
