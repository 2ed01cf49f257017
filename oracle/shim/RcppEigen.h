/*
 * TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * Minimal stand-in for <RcppEigen.h>, sufficient to compile the reference's
 * src/coreLoop.cpp UNMODIFIED in an image that has neither R, Rcpp nor Eigen.
 * It supplies only the container/view types that file names
 * (src/atlasqtl_types.h:8-13: Eigen::Map<MatrixXd|ArrayXXd|ArrayXd|VectorXd>,
 * Eigen::ArrayXd, Eigen::VectorXi; Rcpp::List / Rcpp::as) with the handful of
 * operations src/coreLoop.cpp:56-85 and :108-136 apply to them: element access,
 * `.col(k) += s * col` / `s * (col - col)`, and the q-vector expression
 * `-(a + s + log(b)) / 2`.  All arithmetic is plain IEEE double, evaluated
 * element by element in the written order, which is what Eigen's lazy
 * expressions do for these coefficient-wise operations.  Column-major storage,
 * zero-copy views, exactly like Eigen::Map over R memory.
 */
#ifndef ORACLE_SHIM_RCPPEIGEN_H_
#define ORACLE_SHIM_RCPPEIGEN_H_

#include <cmath>
#include <cstddef>
#include <vector>

namespace Eigen {

struct MatrixXd {};
struct ArrayXXd {};
struct VectorXd {};

template <class T> class Map;

/* ---- dense owning 1-D array ---- */
class ArrayXd {
 public:
  ArrayXd() {}
  explicit ArrayXd(std::size_t n) : v_(n) {}
  std::size_t size() const { return v_.size(); }
  double& operator[](std::size_t i) { return v_[i]; }
  const double& operator[](std::size_t i) const { return v_[i]; }
  double& operator()(std::size_t i) { return v_[i]; }
  const double& operator()(std::size_t i) const { return v_[i]; }
 private:
  std::vector<double> v_;
};

/* ---- 1-D view ---- */
template <> class Map<ArrayXd> {
 public:
  Map(double* p, std::size_t n) : p_(p), n_(n) {}
  std::size_t size() const { return n_; }
  double& operator[](std::size_t i) { return p_[i]; }
  const double& operator[](std::size_t i) const { return p_[i]; }
  double& operator()(std::size_t i) { return p_[i]; }
  const double& operator()(std::size_t i) const { return p_[i]; }
 private:
  double* p_; std::size_t n_;
};
template <> class Map<VectorXd> : public Map<ArrayXd> {
 public: Map(double* p, std::size_t n) : Map<ArrayXd>(p, n) {}
};

inline ArrayXd operator+(const Map<ArrayXd>& a, double s) {
  ArrayXd r(a.size()); for (std::size_t i = 0; i < a.size(); ++i) r[i] = a[i] + s; return r;
}
inline ArrayXd log(const Map<ArrayXd>& a) {
  ArrayXd r(a.size()); for (std::size_t i = 0; i < a.size(); ++i) r[i] = std::log(a[i]); return r;
}
inline ArrayXd operator+(const ArrayXd& a, const ArrayXd& b) {
  ArrayXd r(a.size()); for (std::size_t i = 0; i < a.size(); ++i) r[i] = a[i] + b[i]; return r;
}
inline ArrayXd operator-(const ArrayXd& a) {
  ArrayXd r(a.size()); for (std::size_t i = 0; i < a.size(); ++i) r[i] = -a[i]; return r;
}
inline ArrayXd operator/(const ArrayXd& a, double s) {
  ArrayXd r(a.size()); for (std::size_t i = 0; i < a.size(); ++i) r[i] = a[i] / s; return r;
}

/* ---- column views and the two column expressions the sweep uses ---- */
struct ConstCol { const double* p; std::size_t n; };
struct DiffCol { const double* a; const double* b; std::size_t n; };
struct ScaledCol { double s; const double* p; std::size_t n; };
struct ScaledDiffCol { double s; const double* a; const double* b; std::size_t n; };
inline DiffCol operator-(const ConstCol& x, const ConstCol& y) { return DiffCol{x.p, y.p, x.n}; }
struct Col;
inline DiffCol operator-(const ConstCol& x, const Col& y);
inline ScaledCol operator*(double s, const ConstCol& x) { return ScaledCol{s, x.p, x.n}; }
inline ScaledDiffCol operator*(double s, const DiffCol& x) { return ScaledDiffCol{s, x.a, x.b, x.n}; }
struct Col {
  double* p; std::size_t n;
  operator ConstCol() const { return ConstCol{p, n}; }
  Col& operator+=(const ScaledCol& e) { for (std::size_t i = 0; i < n; ++i) p[i] += e.s * e.p[i]; return *this; }
  Col& operator+=(const ScaledDiffCol& e) { for (std::size_t i = 0; i < n; ++i) p[i] += e.s * (e.a[i] - e.b[i]); return *this; }
};

inline DiffCol operator-(const ConstCol& x, const Col& y) { return DiffCol{x.p, y.p, x.n}; }

/* ---- 2-D column-major views ---- */
class Map2DBase {
 public:
  Map2DBase(double* p, std::size_t r, std::size_t c) : p_(p), r_(r), c_(c) {}
  std::size_t rows() const { return r_; }
  std::size_t cols() const { return c_; }
  double& operator()(std::size_t i, std::size_t j) { return p_[i + j * r_]; }
  const double& operator()(std::size_t i, std::size_t j) const { return p_[i + j * r_]; }
  Col col(std::size_t j) { return Col{p_ + j * r_, r_}; }
  ConstCol col(std::size_t j) const { return ConstCol{p_ + j * r_, r_}; }
 protected:
  double* p_; std::size_t r_, c_;
};
template <> class Map<MatrixXd> : public Map2DBase {
 public: Map(double* p, std::size_t r, std::size_t c) : Map2DBase(p, r, c) {}
};
template <> class Map<ArrayXXd> : public Map2DBase {
 public: Map(double* p, std::size_t r, std::size_t c) : Map2DBase(p, r, c) {}
};

/* ---- integer index vector (copied, 0-based) ---- */
class VectorXi {
 public:
  VectorXi(const int* p, std::size_t n) : v_(p, p + n) {}
  int size() const { return (int)v_.size(); }
  int operator[](std::size_t i) const { return v_[i]; }
 private:
  std::vector<int> v_;
};

}  // namespace Eigen

namespace Rcpp {
/* A `List` of q p-by-p matrices (cp_X_rm in coreDualMisLoop). */
struct ListElem { double* p; std::size_t r, c; };
class List {
 public:
  List() {}
  void push_back(double* p, std::size_t r, std::size_t c) { v_.push_back(ListElem{p, r, c}); }
  const ListElem& operator[](std::size_t k) const { return v_[k]; }
 private:
  std::vector<ListElem> v_;
};
template <class T> inline T as(const ListElem& e) { return T(e.p, e.r, e.c); }
}  // namespace Rcpp

#endif
