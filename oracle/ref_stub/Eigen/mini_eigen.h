// mini_eigen.h -- a small, eager-evaluation stand-in for the subset of Eigen 3 that the reference's g2o + vertex/edge
// types use.  TEST INFRASTRUCTURE ONLY: it exists so that the UNMODIFIED reference sources under /root/reference
// (Thirdparty/g2o/g2o/{core,types,solvers,stuff}, include/G2O_Plane3D.h, include/g2o_cuboid.h, src/g2o_cuboid.cc,
// src/matrix_utils.cc) can be compiled here, where Eigen itself is not installed, into oracle/_ref/ (see oracle/Makefile.ref).
// Nothing in the product links it.  Algorithms that influence results follow Eigen's published ones (quaternion <-> matrix,
// AngleAxis, closed-form 2x2/3x3/4x4 inverses, pivoted LDLT with sign tracking); everything is column-major double/float
// arithmetic without expression templates (every operator returns a plain Matrix).
#ifndef PPO_MINI_EIGEN_H
#define PPO_MINI_EIGEN_H

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <sstream>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_WORLD_VERSION 3
#define EIGEN_MAJOR_VERSION 2
#define EIGEN_MINOR_VERSION 0
#define EIGEN_VERSION_AT_LEAST(x, y, z) (EIGEN_WORLD_VERSION > x || (EIGEN_WORLD_VERSION >= x && (EIGEN_MAJOR_VERSION > y || (EIGEN_MAJOR_VERSION >= y && EIGEN_MINOR_VERSION >= z))))
#define EIGEN_DEFINE_STL_VECTOR_SPECIALIZATION(...)

namespace Eigen {

typedef std::ptrdiff_t DenseIndex;
const int Dynamic = -1;
enum { ColMajor = 0, RowMajor = 0x1, AutoAlign = 0, DontAlign = 0x2 };
enum { Unaligned = 0, Aligned = 1 };
const unsigned int AlignedBit = 0x80;
enum { Lower = 0x1, Upper = 0x2 };
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };
enum TransformTraits { Isometry = 0x1, Affine = 0x2, AffineCompact = 0x10 | Affine, Projective = 0x20 };
enum DecompositionOptions { EigenvaluesOnly = 0x40, ComputeEigenvectors = 0x80 };
inline void initParallel() {}

template <typename T>
struct NumTraits {
  typedef T Real;
  static T epsilon() { return std::numeric_limits<T>::epsilon(); }
  static T dummy_precision() { return sizeof(T) == 4 ? T(1e-5) : T(1e-12); }
  static T highest() { return (std::numeric_limits<T>::max)(); }
  static T lowest() { return -(std::numeric_limits<T>::max)(); }
};

template <typename T>
class aligned_allocator : public std::allocator<T> {
 public:
  template <class U>
  struct rebind {
    typedef aligned_allocator<U> other;
  };
  aligned_allocator() {}
  aligned_allocator(const aligned_allocator &o) : std::allocator<T>(o) {}
  template <class U>
  aligned_allocator(const aligned_allocator<U> &) {}
};

template <typename Scalar, int Rows, int Cols, int Options = 0, int MaxRows = Rows, int MaxCols = Cols>
class Matrix;
template <typename Xpr, int BR, int BC>
class Block;
template <typename Xpr>
class Transpose;
template <typename Xpr>
class ArrayWrapper;
template <typename Xpr>
class DiagonalWrapper;
template <typename Xpr>
class Diagonal;
template <typename Plain, int MapOptions = Unaligned>
class Map;
template <typename Xpr, int Dir>
class VectorwiseOp;
template <typename Derived>
struct CommaInitializer;
template <typename MatrixType, int UpLo = Lower>
class LDLT;
template <typename MatrixType, int UpLo = Lower>
class LLT;

template <typename T>
struct traits;
template <typename T>
struct traits<const T> : traits<T> {};
template <typename S, int R, int C, int O, int MR, int MC>
struct traits<Matrix<S, R, C, O, MR, MC> > {
  typedef S Scalar;
  enum { Rows = R, Cols = C, Options = O };
};
template <typename X, int BR, int BC>
struct traits<Block<X, BR, BC> > {
  typedef typename traits<X>::Scalar Scalar;
  enum { Rows = BR, Cols = BC, Options = 0 };
};
template <typename X>
struct traits<Transpose<X> > {
  typedef typename traits<X>::Scalar Scalar;
  enum { Rows = traits<X>::Cols, Cols = traits<X>::Rows, Options = 0 };
};
template <typename X>
struct traits<ArrayWrapper<X> > : traits<X> {};
template <typename X>
struct traits<Diagonal<X> > {
  typedef typename traits<X>::Scalar Scalar;
  enum { Rows = (traits<X>::Rows == Dynamic || traits<X>::Cols == Dynamic) ? Dynamic : (traits<X>::Rows < traits<X>::Cols ? traits<X>::Rows : traits<X>::Cols), Cols = 1, Options = 0 };
};
template <typename P, int MO>
struct traits<Map<P, MO> > : traits<P> {};

namespace internal {
template <typename T>
struct remove_const {
  typedef T type;
};
template <typename T>
struct remove_const<const T> {
  typedef T type;
};
template <bool C, typename A, typename B>
struct conditional {
  typedef A type;
};
template <typename A, typename B>
struct conditional<false, A, B> {
  typedef B type;
};
template <int A, int B>
struct pick_size {  // size of an element-wise binary result: a fixed size wins over Dynamic
  enum { value = (A == Dynamic) ? B : A };
};
}  // namespace internal

template <typename Derived>
class MatrixBase {
 public:
  typedef typename traits<Derived>::Scalar Scalar;
  typedef Scalar RealScalar;
  typedef DenseIndex Index;
  enum {
    RowsAtCompileTime = traits<Derived>::Rows,
    ColsAtCompileTime = traits<Derived>::Cols,
    SizeAtCompileTime = (RowsAtCompileTime == Dynamic || ColsAtCompileTime == Dynamic) ? Dynamic : RowsAtCompileTime * ColsAtCompileTime,
    IsVectorAtCompileTime = (RowsAtCompileTime == 1 || ColsAtCompileTime == 1) ? 1 : 0,
    Flags = 0
  };
  typedef Matrix<Scalar, RowsAtCompileTime, ColsAtCompileTime> PlainObject;
  typedef PlainObject PlainMatrix;

  Derived &derived() { return *static_cast<Derived *>(this); }
  const Derived &derived() const { return *static_cast<const Derived *>(this); }
  Derived &const_cast_derived() const { return *const_cast<Derived *>(static_cast<const Derived *>(this)); }

  int rows() const { return derived().rows_(); }
  int cols() const { return derived().cols_(); }
  int size() const { return rows() * cols(); }
  Scalar coeff(int i, int j) const { return derived().get_(i, j); }
  Scalar &coeffRef(int i, int j) { return derived().ref_(i, j); }
  Scalar operator()(int i, int j) const { return derived().get_(i, j); }
  Scalar &operator()(int i, int j) { return derived().ref_(i, j); }
  // vector access
  Scalar coeff(int i) const { return cols() == 1 ? derived().get_(i, 0) : derived().get_(0, i); }
  Scalar &coeffRef(int i) { return cols() == 1 ? derived().ref_(i, 0) : derived().ref_(0, i); }
  Scalar operator()(int i) const { return coeff(i); }
  Scalar &operator()(int i) { return coeffRef(i); }
  Scalar operator[](int i) const { return coeff(i); }
  Scalar &operator[](int i) { return coeffRef(i); }
  Scalar x() const { return coeff(0); }
  Scalar y() const { return coeff(1); }
  Scalar z() const { return coeff(2); }
  Scalar w() const { return coeff(3); }
  Scalar &x() { return coeffRef(0); }
  Scalar &y() { return coeffRef(1); }
  Scalar &z() { return coeffRef(2); }
  Scalar &w() { return coeffRef(3); }

  PlainObject eval() const { return PlainObject(derived()); }
  template <typename NewScalar>
  Matrix<NewScalar, RowsAtCompileTime, ColsAtCompileTime> cast() const {
    Matrix<NewScalar, RowsAtCompileTime, ColsAtCompileTime> r;
    r.resize(rows(), cols());
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) r(i, j) = NewScalar(coeff(i, j));
    return r;
  }

  // ---- views (const-correctness is relaxed: a view obtained from a const object is never written through) --------
  template <int BR, int BC>
  Block<Derived, BR, BC> block(int i, int j) const { return Block<Derived, BR, BC>(const_cast_derived(), i, j, BR, BC); }
  Block<Derived, Dynamic, Dynamic> block(int i, int j, int r, int c) const { return Block<Derived, Dynamic, Dynamic>(const_cast_derived(), i, j, r, c); }
  template <int BR, int BC>
  Block<Derived, BR, BC> block(int i, int j, int r, int c) const { return Block<Derived, BR, BC>(const_cast_derived(), i, j, r, c); }
  Block<Derived, RowsAtCompileTime, 1> col(int j) const { return Block<Derived, RowsAtCompileTime, 1>(const_cast_derived(), 0, j, rows(), 1); }
  Block<Derived, 1, ColsAtCompileTime> row(int i) const { return Block<Derived, 1, ColsAtCompileTime>(const_cast_derived(), i, 0, 1, cols()); }
  template <int N>
  Block<Derived, (ColsAtCompileTime == 1 ? N : 1), (ColsAtCompileTime == 1 ? 1 : N)> segment(int start) const {
    typedef Block<Derived, (ColsAtCompileTime == 1 ? N : 1), (ColsAtCompileTime == 1 ? 1 : N)> B;
    return cols() == 1 ? B(const_cast_derived(), start, 0, N, 1) : B(const_cast_derived(), 0, start, 1, N);
  }
  Block<Derived, (ColsAtCompileTime == 1 ? Dynamic : 1), (ColsAtCompileTime == 1 ? 1 : Dynamic)> segment(int start, int n) const {
    typedef Block<Derived, (ColsAtCompileTime == 1 ? Dynamic : 1), (ColsAtCompileTime == 1 ? 1 : Dynamic)> B;
    return cols() == 1 ? B(const_cast_derived(), start, 0, n, 1) : B(const_cast_derived(), 0, start, 1, n);
  }
  template <int N>
  Block<Derived, (ColsAtCompileTime == 1 ? N : 1), (ColsAtCompileTime == 1 ? 1 : N)> head() const { return segment<N>(0); }
  template <int N>
  Block<Derived, (ColsAtCompileTime == 1 ? N : 1), (ColsAtCompileTime == 1 ? 1 : N)> tail() const { return segment<N>(size() - N); }
  Block<Derived, (ColsAtCompileTime == 1 ? Dynamic : 1), (ColsAtCompileTime == 1 ? 1 : Dynamic)> head(int n) const { return segment(0, n); }
  Block<Derived, (ColsAtCompileTime == 1 ? Dynamic : 1), (ColsAtCompileTime == 1 ? 1 : Dynamic)> tail(int n) const { return segment(size() - n, n); }
  template <int R, int C>
  Block<Derived, R, C> topLeftCorner() const { return block<R, C>(0, 0); }
  template <int R, int C>
  Block<Derived, R, C> topRightCorner() const { return block<R, C>(0, cols() - C); }
  template <int R, int C>
  Block<Derived, R, C> bottomLeftCorner() const { return block<R, C>(rows() - R, 0); }
  template <int R, int C>
  Block<Derived, R, C> bottomRightCorner() const { return block<R, C>(rows() - R, cols() - C); }
  Block<Derived, Dynamic, Dynamic> topLeftCorner(int r, int c) const { return block(0, 0, r, c); }
  Block<Derived, Dynamic, Dynamic> topRightCorner(int r, int c) const { return block(0, cols() - c, r, c); }
  Block<Derived, Dynamic, Dynamic> bottomLeftCorner(int r, int c) const { return block(rows() - r, 0, r, c); }
  Block<Derived, Dynamic, Dynamic> bottomRightCorner(int r, int c) const { return block(rows() - r, cols() - c, r, c); }
  template <int N>
  Block<Derived, N, ColsAtCompileTime> topRows() const { return Block<Derived, N, ColsAtCompileTime>(const_cast_derived(), 0, 0, N, cols()); }
  template <int N>
  Block<Derived, N, ColsAtCompileTime> bottomRows() const { return Block<Derived, N, ColsAtCompileTime>(const_cast_derived(), rows() - N, 0, N, cols()); }
  template <int N>
  Block<Derived, RowsAtCompileTime, N> leftCols() const { return Block<Derived, RowsAtCompileTime, N>(const_cast_derived(), 0, 0, rows(), N); }
  template <int N>
  Block<Derived, RowsAtCompileTime, N> rightCols() const { return Block<Derived, RowsAtCompileTime, N>(const_cast_derived(), 0, cols() - N, rows(), N); }
  Block<Derived, Dynamic, ColsAtCompileTime> topRows(int n) const { return Block<Derived, Dynamic, ColsAtCompileTime>(const_cast_derived(), 0, 0, n, cols()); }
  Block<Derived, Dynamic, ColsAtCompileTime> bottomRows(int n) const { return Block<Derived, Dynamic, ColsAtCompileTime>(const_cast_derived(), rows() - n, 0, n, cols()); }
  Block<Derived, Dynamic, ColsAtCompileTime> middleRows(int s, int n) const { return Block<Derived, Dynamic, ColsAtCompileTime>(const_cast_derived(), s, 0, n, cols()); }
  Block<Derived, RowsAtCompileTime, Dynamic> leftCols(int n) const { return Block<Derived, RowsAtCompileTime, Dynamic>(const_cast_derived(), 0, 0, rows(), n); }
  Block<Derived, RowsAtCompileTime, Dynamic> rightCols(int n) const { return Block<Derived, RowsAtCompileTime, Dynamic>(const_cast_derived(), 0, cols() - n, rows(), n); }
  Block<Derived, RowsAtCompileTime, Dynamic> middleCols(int s, int n) const { return Block<Derived, RowsAtCompileTime, Dynamic>(const_cast_derived(), 0, s, rows(), n); }
  Transpose<Derived> transpose() const { return Transpose<Derived>(const_cast_derived()); }
  Transpose<Derived> adjoint() const { return Transpose<Derived>(const_cast_derived()); }
  ArrayWrapper<Derived> array() const { return ArrayWrapper<Derived>(const_cast_derived()); }
  Derived &matrix() { return derived(); }
  const Derived &matrix() const { return derived(); }
  Diagonal<Derived> diagonal() const { return Diagonal<Derived>(const_cast_derived()); }
  DiagonalWrapper<Derived> asDiagonal() const { return DiagonalWrapper<Derived>(derived()); }
  VectorwiseOp<Derived, 1> rowwise() const { return VectorwiseOp<Derived, 1>(derived()); }
  VectorwiseOp<Derived, 0> colwise() const { return VectorwiseOp<Derived, 0>(derived()); }
  Derived &noalias() { return derived(); }
  Derived &lazyAssign(const PlainObject &o) { return derived() = o; }

  // ---- reductions ------------------------------------------------------------------------------------------
  Scalar squaredNorm() const {
    Scalar s = 0;
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) s += coeff(i, j) * coeff(i, j);
    return s;
  }
  Scalar norm() const { return std::sqrt(squaredNorm()); }
  Scalar sum() const {
    Scalar s = 0;
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) s += coeff(i, j);
    return s;
  }
  Scalar mean() const { return sum() / Scalar(size()); }
  Scalar trace() const {
    Scalar s = 0;
    for (int i = 0; i < rows() && i < cols(); i++) s += coeff(i, i);
    return s;
  }
  template <typename O>
  Scalar dot(const MatrixBase<O> &o) const {
    Scalar s = 0;
    for (int i = 0; i < size(); i++) s += coeff(i) * o.coeff(i);
    return s;
  }
  Scalar maxCoeff() const {
    Scalar m = coeff(0, 0);
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) m = (std::max)(m, coeff(i, j));
    return m;
  }
  Scalar minCoeff() const {
    Scalar m = coeff(0, 0);
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) m = (std::min)(m, coeff(i, j));
    return m;
  }
  template <typename I>
  Scalar maxCoeff(I *idx) const {
    Scalar m = coeff(0);
    *idx = 0;
    for (int i = 1; i < size(); i++)
      if (coeff(i) > m) m = coeff(i), *idx = I(i);
    return m;
  }
  template <typename I>
  Scalar minCoeff(I *idx) const {
    Scalar m = coeff(0);
    *idx = 0;
    for (int i = 1; i < size(); i++)
      if (coeff(i) < m) m = coeff(i), *idx = I(i);
    return m;
  }
  template <typename I>
  Scalar maxCoeff(I *ri, I *ci) const {
    Scalar m = coeff(0, 0);
    *ri = 0, *ci = 0;
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++)
        if (coeff(i, j) > m) m = coeff(i, j), *ri = I(i), *ci = I(j);
    return m;
  }
  bool allFinite() const {
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++)
        if (!std::isfinite(coeff(i, j))) return false;
    return true;
  }
  bool hasNaN() const {
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++)
        if (coeff(i, j) != coeff(i, j)) return true;
    return false;
  }
  template <typename O>
  bool isApprox(const MatrixBase<O> &o, Scalar prec = NumTraits<Scalar>::dummy_precision()) const {
    Scalar d = 0;
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) d += (coeff(i, j) - o.coeff(i, j)) * (coeff(i, j) - o.coeff(i, j));
    return d <= prec * prec * (std::min)(squaredNorm(), o.squaredNorm());
  }

  // ---- element-wise / algebra returning plain objects ----------------------------------------------------------
  PlainObject normalized() const {
    PlainObject r(derived());
    const Scalar n = norm();
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) r(i, j) = coeff(i, j) / n;
    return r;
  }
  void normalize() {
    const Scalar n = norm();
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) coeffRef(i, j) /= n;
  }
  PlainObject cwiseAbs() const {
    PlainObject r(derived());
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) r(i, j) = std::abs(coeff(i, j));
    return r;
  }
  PlainObject cwiseSqrt() const {
    PlainObject r(derived());
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) r(i, j) = std::sqrt(coeff(i, j));
    return r;
  }
  PlainObject cwiseInverse() const {
    PlainObject r(derived());
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) r(i, j) = Scalar(1) / coeff(i, j);
    return r;
  }
  template <typename O>
  PlainObject cwiseProduct(const MatrixBase<O> &o) const {
    PlainObject r(derived());
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) r(i, j) = coeff(i, j) * o.coeff(i, j);
    return r;
  }
  template <typename O>
  PlainObject cwiseQuotient(const MatrixBase<O> &o) const {
    PlainObject r(derived());
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) r(i, j) = coeff(i, j) / o.coeff(i, j);
    return r;
  }
  template <typename O>
  PlainObject cwiseMax(const MatrixBase<O> &o) const {
    PlainObject r(derived());
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) r(i, j) = (std::max)(coeff(i, j), o.coeff(i, j));
    return r;
  }
  template <typename O>
  PlainObject cwiseMin(const MatrixBase<O> &o) const {
    PlainObject r(derived());
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) r(i, j) = (std::min)(coeff(i, j), o.coeff(i, j));
    return r;
  }
  template <typename O>
  Matrix<Scalar, 3, 1> cross(const MatrixBase<O> &o) const {
    Matrix<Scalar, 3, 1> r;
    r(0) = coeff(1) * o.coeff(2) - coeff(2) * o.coeff(1);
    r(1) = coeff(2) * o.coeff(0) - coeff(0) * o.coeff(2);
    r(2) = coeff(0) * o.coeff(1) - coeff(1) * o.coeff(0);
    return r;
  }
  Matrix<Scalar, (RowsAtCompileTime == Dynamic ? Dynamic : RowsAtCompileTime + 1), 1> homogeneous() const {
    Matrix<Scalar, (RowsAtCompileTime == Dynamic ? Dynamic : RowsAtCompileTime + 1), 1> r;
    r.resize(size() + 1, 1);
    for (int i = 0; i < size(); i++) r(i) = coeff(i);
    r(size()) = Scalar(1);
    return r;
  }
  Matrix<Scalar, (RowsAtCompileTime == Dynamic ? Dynamic : RowsAtCompileTime - 1), 1> hnormalized() const {
    Matrix<Scalar, (RowsAtCompileTime == Dynamic ? Dynamic : RowsAtCompileTime - 1), 1> r;
    r.resize(size() - 1, 1);
    for (int i = 0; i < size() - 1; i++) r(i) = coeff(i) / coeff(size() - 1);
    return r;
  }
  PlainObject inverse() const;
  Scalar determinant() const;
  LDLT<PlainObject> ldlt() const;
  LLT<PlainObject> llt() const;

  // ---- in-place ------------------------------------------------------------------------------------------------
  Derived &setZero() { return setConstant(Scalar(0)); }
  Derived &setOnes() { return setConstant(Scalar(1)); }
  Derived &setConstant(const Scalar &v) {
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) coeffRef(i, j) = v;
    return derived();
  }
  void fill(const Scalar &v) { setConstant(v); }
  Derived &setIdentity() {
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) coeffRef(i, j) = (i == j) ? Scalar(1) : Scalar(0);
    return derived();
  }
  Derived &setRandom() {
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) coeffRef(i, j) = Scalar(2.0 * std::rand() / RAND_MAX - 1.0);
    return derived();
  }
  template <typename O>
  Derived &assign_(const MatrixBase<O> &o) {
    // sources may alias the destination (e.g. x = x.transpose() is not used by the reference, but m.block() = m2 * m.block() is
    // evaluated into a plain temporary by operator* already), so a straight copy is enough
    derived().resize_like_(o.rows(), o.cols());
    if (rows() == o.rows() && cols() == o.cols()) {
      for (int j = 0; j < cols(); j++)
        for (int i = 0; i < rows(); i++) coeffRef(i, j) = o.coeff(i, j);
    } else {  // vector <-> transposed vector
      assert(size() == o.size());
      for (int i = 0; i < size(); i++) coeffRef(i) = o.coeff(i);
    }
    return derived();
  }
  template <typename O>
  Derived &operator+=(const MatrixBase<O> &o) {
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) coeffRef(i, j) += o.coeff(i, j);
    return derived();
  }
  template <typename O>
  Derived &operator-=(const MatrixBase<O> &o) {
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) coeffRef(i, j) -= o.coeff(i, j);
    return derived();
  }
  Derived &operator*=(const Scalar &s) {
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) coeffRef(i, j) *= s;
    return derived();
  }
  Derived &operator/=(const Scalar &s) {
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) coeffRef(i, j) /= s;
    return derived();
  }
  template <typename O>
  Derived &operator*=(const MatrixBase<O> &o) {
    PlainObject t = derived() * o;
    return assign_(t);
  }
  PlainObject operator-() const {
    PlainObject r(derived());
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) r(i, j) = -coeff(i, j);
    return r;
  }
  CommaInitializer<Derived> operator<<(const Scalar &s);
  template <typename O>
  CommaInitializer<Derived> operator<<(const MatrixBase<O> &o);
};

// ---- storage --------------------------------------------------------------------------------------------------------
template <typename S, int R, int C, bool Dyn = (R == Dynamic || C == Dynamic)>
struct DenseStorage {
  S d[R * C > 0 ? R * C : 1];
  DenseStorage() {
    for (int i = 0; i < R * C; i++) d[i] = S();
  }
  int rows() const { return R; }
  int cols() const { return C; }
  void resize(int, int) {}
  S *data() { return d; }
  const S *data() const { return d; }
};
template <typename S, int R, int C>
struct DenseStorage<S, R, C, true> {
  std::vector<S> d;
  int r, c;
  DenseStorage() : r(R == Dynamic ? 0 : R), c(C == Dynamic ? 0 : C) {}
  int rows() const { return r; }
  int cols() const { return c; }
  void resize(int nr, int nc) {
    if (nr != r || nc != c || (int)d.size() != nr * nc) {
      r = nr, c = nc;
      d.assign((size_t)nr * nc, S());
    }
  }
  S *data() { return d.empty() ? 0 : &d[0]; }
  const S *data() const { return d.empty() ? 0 : &d[0]; }
};

template <typename S, int R, int C, int O, int MR, int MC>
class Matrix : public MatrixBase<Matrix<S, R, C, O, MR, MC> > {
  DenseStorage<S, R, C> m;
  enum { IsRowMajor = (O & RowMajor) ? 1 : 0, IsDyn = (R == Dynamic || C == Dynamic) ? 1 : 0 };

 public:
  typedef MatrixBase<Matrix> Base;
  typedef S Scalar;
  typedef Map<Matrix, Unaligned> MapType;
  typedef Map<Matrix, Aligned> AlignedMapType;
  typedef Map<const Matrix, Unaligned> ConstMapType;
  typedef Map<const Matrix, Aligned> ConstAlignedMapType;
  Matrix() {}
  Matrix(const Matrix &o) : Base(), m(o.m) {}
  // one argument: size of a dynamic vector, or (fixed size) a pointer to coefficients / a scalar for 1x1
  explicit Matrix(int n) {
    if (IsDyn) {
      if (C == 1) m.resize(n, 1);
      else if (R == 1) m.resize(1, n);
      else m.resize(n, n);
    } else if (R * C == 1) {
      m.data()[0] = S(n);
    }
  }
  explicit Matrix(const S *p) {
    for (int i = 0; i < R * C; i++) m.data()[i] = p[i];
  }
  // two arguments: (rows, cols) for dynamic, (x, y) for a fixed 2-vector
  template <typename A, typename B>
  Matrix(const A &a, const B &b) {
    init2_(a, b);
  }
  Matrix(const S &x, const S &y, const S &z) {
    m.resize(3, 1);
    m.data()[0] = x, m.data()[1] = y, m.data()[2] = z;
  }
  Matrix(const S &x, const S &y, const S &z, const S &w) {
    m.resize(4, 1);
    m.data()[0] = x, m.data()[1] = y, m.data()[2] = z, m.data()[3] = w;
  }
  template <typename Od>
  Matrix(const MatrixBase<Od> &o) {
    this->assign_(o);
  }
  template <typename Od>
  Matrix(const DiagonalWrapper<Od> &dw) {
    assign_diag_(dw);
  }
  Matrix &operator=(const Matrix &o) {
    m = o.m;
    return *this;
  }
  template <typename Od>
  Matrix &operator=(const MatrixBase<Od> &o) {
    return this->assign_(o);
  }
  template <typename Od>
  Matrix &operator=(const DiagonalWrapper<Od> &dw) {
    assign_diag_(dw);
    return *this;
  }
  int rows_() const { return m.rows(); }
  int cols_() const { return m.cols(); }
  S get_(int i, int j) const { return m.data()[IsRowMajor ? i * m.cols() + j : j * m.rows() + i]; }
  S &ref_(int i, int j) { return m.data()[IsRowMajor ? i * m.cols() + j : j * m.rows() + i]; }
  void resize_like_(int r, int c) {
    if (!IsDyn) return;
    if (R != Dynamic && C == Dynamic && r != R && c == R) std::swap(r, c);
    if (R == Dynamic && C != Dynamic && c != C && r == C) std::swap(r, c);
    if ((R == 1 || C == 1) && r * c > 0) {  // vectors accept either orientation
      if (C == 1) r = r * c, c = 1;
      else c = r * c, r = 1;
    }
    m.resize(r, c);
  }
  void resize(int r, int c) { m.resize(R == Dynamic ? r : R, C == Dynamic ? c : C); }
  void resize(int n) {
    if (C == 1) m.resize(n, 1);
    else if (R == 1) m.resize(1, n);
    else m.resize(n, n);
  }
  void conservativeResize(int r, int c) {
    Matrix t(*this);
    m.resize(r, c);
    for (int j = 0; j < c && j < t.cols(); j++)
      for (int i = 0; i < r && i < t.rows(); i++) ref_(i, j) = t.get_(i, j);
  }
  void conservativeResize(int n) {
    if (C == 1) conservativeResize(n, 1);
    else conservativeResize(1, n);
  }
  S *data() { return m.data(); }
  const S *data() const { return m.data(); }
  void swap(Matrix &o) { std::swap(m, o.m); }

  static Matrix Constant(int r, int c, const S &v) {
    Matrix t;
    t.resize(r, c);
    t.setConstant(v);
    return t;
  }
  static Matrix Constant(int n, const S &v) {
    Matrix t;
    t.resize(n);
    t.setConstant(v);
    return t;
  }
  static Matrix Constant(const S &v) {
    Matrix t;
    t.setConstant(v);
    return t;
  }
  static Matrix Zero() { return Constant(S(0)); }
  static Matrix Zero(int n) { return Constant(n, S(0)); }
  static Matrix Zero(int r, int c) { return Constant(r, c, S(0)); }
  static Matrix Ones() { return Constant(S(1)); }
  static Matrix Ones(int n) { return Constant(n, S(1)); }
  static Matrix Ones(int r, int c) { return Constant(r, c, S(1)); }
  static Matrix Identity() {
    Matrix t;
    t.setIdentity();
    return t;
  }
  static Matrix Identity(int r, int c) {
    Matrix t;
    t.resize(r, c);
    t.setIdentity();
    return t;
  }
  static Matrix Random() {
    Matrix t;
    t.setRandom();
    return t;
  }
  static Matrix Random(int r, int c) {
    Matrix t;
    t.resize(r, c);
    t.setRandom();
    return t;
  }
  static Matrix Unit(int k) {
    Matrix t;
    t.setZero();
    t(k) = S(1);
    return t;
  }
  static Matrix UnitX() { return Unit(0); }
  static Matrix UnitY() { return Unit(1); }
  static Matrix UnitZ() { return Unit(2); }
  static Matrix UnitW() { return Unit(3); }

 private:
  template <typename A, typename B>
  void init2_(const A &a, const B &b) {
    if (IsDyn) {
      m.resize(R == Dynamic ? int(a) : R, C == Dynamic ? int(b) : C);
    } else {
      m.data()[0] = S(a);
      m.data()[1] = S(b);
    }
  }
  template <typename Od>
  void assign_diag_(const DiagonalWrapper<Od> &dw) {
    const int n = dw.v.size();
    m.resize(n, n);
    this->setZero();
    for (int i = 0; i < n; i++) ref_(i, i) = dw.v.coeff(i);
  }
};

// ---- views ----------------------------------------------------------------------------------------------------------
template <typename Xpr, int BR, int BC>
class Block : public MatrixBase<Block<Xpr, BR, BC> > {
  typedef typename internal::remove_const<Xpr>::type X;
  X *x;
  int i0, j0, r, c;

 public:
  typedef MatrixBase<Block> Base;
  typedef typename traits<Xpr>::Scalar Scalar;
  Block(const Xpr &xpr, int i, int j, int rows, int cols) : x(const_cast<X *>(&xpr)), i0(i), j0(j), r(rows), c(cols) {}
  Block(const Block &o) : Base(), x(o.x), i0(o.i0), j0(o.j0), r(o.r), c(o.c) {}
  int rows_() const { return r; }
  int cols_() const { return c; }
  Scalar get_(int i, int j) const { return static_cast<const X *>(x)->get_(i0 + i, j0 + j); }
  Scalar &ref_(int i, int j) { return x->ref_(i0 + i, j0 + j); }
  void resize_like_(int, int) {}
  Block &operator=(const Block &o) {
    typename Base::PlainObject t(o);
    return this->assign_(t);
  }
  template <typename Od>
  Block &operator=(const MatrixBase<Od> &o) {
    typename MatrixBase<Od>::PlainObject t(o.derived());  // the source may alias this block
    return this->assign_(t);
  }
};
template <typename Xpr>
class Transpose : public MatrixBase<Transpose<Xpr> > {
  typedef typename internal::remove_const<Xpr>::type X;
  X *x;

 public:
  typedef MatrixBase<Transpose> Base;
  typedef typename traits<Xpr>::Scalar Scalar;
  explicit Transpose(const Xpr &xpr) : x(const_cast<X *>(&xpr)) {}
  Transpose(const Transpose &o) : Base(), x(o.x) {}
  int rows_() const { return static_cast<const X *>(x)->cols_(); }
  int cols_() const { return static_cast<const X *>(x)->rows_(); }
  Scalar get_(int i, int j) const { return static_cast<const X *>(x)->get_(j, i); }
  Scalar &ref_(int i, int j) { return x->ref_(j, i); }
  void resize_like_(int, int) {}
  Transpose &operator=(const Transpose &o) {
    typename Base::PlainObject t(o);
    return this->assign_(t);
  }
  template <typename Od>
  Transpose &operator=(const MatrixBase<Od> &o) {
    typename MatrixBase<Od>::PlainObject t(o.derived());
    return this->assign_(t);
  }
};
template <typename Xpr>
class Diagonal : public MatrixBase<Diagonal<Xpr> > {
  typedef typename internal::remove_const<Xpr>::type X;
  X *x;

 public:
  typedef MatrixBase<Diagonal> Base;
  typedef typename traits<Xpr>::Scalar Scalar;
  explicit Diagonal(const Xpr &xpr) : x(const_cast<X *>(&xpr)) {}
  Diagonal(const Diagonal &o) : Base(), x(o.x) {}
  int rows_() const { return (std::min)(static_cast<const X *>(x)->rows_(), static_cast<const X *>(x)->cols_()); }
  int cols_() const { return 1; }
  Scalar get_(int i, int) const { return static_cast<const X *>(x)->get_(i, i); }
  Scalar &ref_(int i, int) { return x->ref_(i, i); }
  void resize_like_(int, int) {}
  Diagonal &operator=(const Diagonal &o) {
    typename Base::PlainObject t(o);
    return this->assign_(t);
  }
  template <typename Od>
  Diagonal &operator=(const MatrixBase<Od> &o) {
    return this->assign_(o);
  }
};
template <typename Xpr>
class DiagonalWrapper {
 public:
  typename MatrixBase<Xpr>::PlainObject v;
  explicit DiagonalWrapper(const Xpr &xpr) : v(xpr) {}
  Matrix<typename traits<Xpr>::Scalar, Dynamic, Dynamic> toDenseMatrix() const { return Matrix<typename traits<Xpr>::Scalar, Dynamic, Dynamic>(*this); }
};
// array(): element-wise * and / (and comparison-free math used by the reference)
template <typename Xpr>
class ArrayWrapper : public MatrixBase<ArrayWrapper<Xpr> > {
  typedef typename internal::remove_const<Xpr>::type X;
  X *x;

 public:
  typedef MatrixBase<ArrayWrapper> Base;
  typedef typename traits<Xpr>::Scalar Scalar;
  explicit ArrayWrapper(const Xpr &xpr) : x(const_cast<X *>(&xpr)) {}
  ArrayWrapper(const ArrayWrapper &o) : Base(), x(o.x) {}
  int rows_() const { return static_cast<const X *>(x)->rows_(); }
  int cols_() const { return static_cast<const X *>(x)->cols_(); }
  Scalar get_(int i, int j) const { return static_cast<const X *>(x)->get_(i, j); }
  Scalar &ref_(int i, int j) { return x->ref_(i, j); }
  void resize_like_(int, int) {}
  using Base::operator+=;
  using Base::operator-=;
  using Base::operator*=;
  using Base::operator/=;
  ArrayWrapper &operator+=(const Scalar &v) {
    for (int j = 0; j < cols_(); j++)
      for (int i = 0; i < rows_(); i++) ref_(i, j) += v;
    return *this;
  }
  ArrayWrapper &operator-=(const Scalar &v) {
    for (int j = 0; j < cols_(); j++)
      for (int i = 0; i < rows_(); i++) ref_(i, j) -= v;
    return *this;
  }
  template <typename Od>
  ArrayWrapper &operator*=(const ArrayWrapper<Od> &o) {
    for (int j = 0; j < cols_(); j++)
      for (int i = 0; i < rows_(); i++) ref_(i, j) *= o.get_(i, j);
    return *this;
  }
  template <typename Od>
  ArrayWrapper &operator/=(const ArrayWrapper<Od> &o) {
    for (int j = 0; j < cols_(); j++)
      for (int i = 0; i < rows_(); i++) ref_(i, j) /= o.get_(i, j);
    return *this;
  }
  typename Base::PlainObject abs() const { return this->cwiseAbs(); }
  typename Base::PlainObject sqrt() const { return this->cwiseSqrt(); }
  typename Base::PlainObject square() const { return this->cwiseProduct(*this); }
  typename Base::PlainObject inverse() const { return this->cwiseInverse(); }
  template <typename Od>
  ArrayWrapper &operator=(const MatrixBase<Od> &o) {
    return this->assign_(o);
  }
};
template <typename A, typename B>
typename MatrixBase<A>::PlainObject operator*(const ArrayWrapper<A> &a, const ArrayWrapper<B> &b) {
  return a.cwiseProduct(b);
}
template <typename A, typename B>
typename MatrixBase<A>::PlainObject operator/(const ArrayWrapper<A> &a, const ArrayWrapper<B> &b) {
  return a.cwiseQuotient(b);
}
template <typename A, typename B>
typename MatrixBase<A>::PlainObject operator/(const ArrayWrapper<A> &a, const MatrixBase<B> &b) {
  return a.cwiseQuotient(b);
}
template <typename A>
typename MatrixBase<A>::PlainObject operator+(const ArrayWrapper<A> &a, const typename traits<A>::Scalar &v) {
  typename MatrixBase<A>::PlainObject r(a);
  for (int j = 0; j < r.cols(); j++)
    for (int i = 0; i < r.rows(); i++) r(i, j) += v;
  return r;
}
template <typename A>
typename MatrixBase<A>::PlainObject operator-(const ArrayWrapper<A> &a, const typename traits<A>::Scalar &v) {
  return a + (-v);
}

template <typename Plain, int MapOptions>
class Map : public MatrixBase<Map<Plain, MapOptions> > {
  typedef typename traits<Plain>::Scalar S;
  enum { R = traits<Plain>::Rows, C = traits<Plain>::Cols };
  S *p;
  int r, c;

 public:
  typedef MatrixBase<Map> Base;
  typedef S Scalar;
  Map(const S *ptr) : p(const_cast<S *>(ptr)), r(R), c(C) {}
  Map(const S *ptr, int n) : p(const_cast<S *>(ptr)), r(C == 1 ? n : (R == Dynamic ? n : R)), c(C == 1 ? 1 : (R == 1 ? n : (C == Dynamic ? n : C))) {}
  Map(const S *ptr, int rows, int cols) : p(const_cast<S *>(ptr)), r(rows), c(cols) {}
  Map(const Map &o) : Base(), p(o.p), r(o.r), c(o.c) {}
  int rows_() const { return r; }
  int cols_() const { return c; }
  S get_(int i, int j) const { return p[(size_t)j * r + i]; }
  S &ref_(int i, int j) { return p[(size_t)j * r + i]; }
  void resize_like_(int, int) {}
  void resize(int rows, int cols) { assert(rows == r && cols == c); (void)rows; (void)cols; }  // a Map cannot change size (Eigen asserts the same)
  S *data() { return p; }
  const S *data() const { return p; }
  Map &operator=(const Map &o) {
    typename Base::PlainObject t(o);
    return this->assign_(t);
  }
  template <typename Od>
  Map &operator=(const MatrixBase<Od> &o) {
    typename MatrixBase<Od>::PlainObject t(o.derived());
    return this->assign_(t);
  }
};

template <typename Xpr, int Dir>
class VectorwiseOp {
  typedef typename traits<Xpr>::Scalar S;
  typename MatrixBase<Xpr>::PlainObject m;
  typedef Matrix<S, (Dir == 1 ? traits<Xpr>::Rows : 1), (Dir == 1 ? 1 : traits<Xpr>::Cols)> Res;
  template <typename F>
  Res reduce(F f) const {
    Res res;
    if (Dir == 1) {
      res.resize(m.rows(), 1);
      for (int i = 0; i < m.rows(); i++) {
        S a = m(i, 0);
        for (int j = 1; j < m.cols(); j++) a = f(a, m(i, j));
        res(i, 0) = a;
      }
    } else {
      res.resize(1, m.cols());
      for (int j = 0; j < m.cols(); j++) {
        S a = m(0, j);
        for (int i = 1; i < m.rows(); i++) a = f(a, m(i, j));
        res(0, j) = a;
      }
    }
    return res;
  }
  static S fmax_(S a, S b) { return a < b ? b : a; }
  static S fmin_(S a, S b) { return b < a ? b : a; }
  static S fsum_(S a, S b) { return a + b; }

 public:
  explicit VectorwiseOp(const Xpr &x) : m(x) {}
  Res maxCoeff() const { return reduce(&fmax_); }
  Res minCoeff() const { return reduce(&fmin_); }
  Res sum() const { return reduce(&fsum_); }
  Res mean() const {
    Res r = sum();
    r /= S(Dir == 1 ? m.cols() : m.rows());
    return r;
  }
  Res norm() const {
    typename MatrixBase<Xpr>::PlainObject sq = m.cwiseProduct(m);
    Res r = VectorwiseOp<typename MatrixBase<Xpr>::PlainObject, Dir>(sq).sum();
    for (int i = 0; i < r.size(); i++) r(i) = std::sqrt(r(i));
    return r;
  }
};

template <typename Derived>
struct CommaInitializer {
  Derived &m;
  int row, col, blockRows;
  CommaInitializer(Derived &mat, const typename traits<Derived>::Scalar &s) : m(mat), row(0), col(1), blockRows(1) { m.coeffRef(0, 0) = s; }
  template <typename O>
  CommaInitializer(Derived &mat, const MatrixBase<O> &o) : m(mat), row(0), col(o.cols()), blockRows(o.rows()) {
    for (int j = 0; j < o.cols(); j++)
      for (int i = 0; i < o.rows(); i++) m.coeffRef(i, j) = o.coeff(i, j);
  }
  CommaInitializer &operator,(const typename traits<Derived>::Scalar &s) {
    if (col == m.cols()) row += blockRows, col = 0, blockRows = 1;
    m.coeffRef(row, col++) = s;
    return *this;
  }
  template <typename O>
  CommaInitializer &operator,(const MatrixBase<O> &o) {
    if (col == m.cols()) row += blockRows, col = 0, blockRows = o.rows();
    for (int j = 0; j < o.cols(); j++)
      for (int i = 0; i < o.rows(); i++) m.coeffRef(row + i, col + j) = o.coeff(i, j);
    col += o.cols();
    return *this;
  }
  Derived &finished() { return m; }
};
template <typename D>
CommaInitializer<D> MatrixBase<D>::operator<<(const Scalar &s) {
  return CommaInitializer<D>(derived(), s);
}
template <typename D>
template <typename O>
CommaInitializer<D> MatrixBase<D>::operator<<(const MatrixBase<O> &o) {
  return CommaInitializer<D>(derived(), o);
}

// ---- arithmetic -----------------------------------------------------------------------------------------------------
template <typename A, typename B>
struct SumType {
  typedef Matrix<typename traits<A>::Scalar, internal::pick_size<traits<A>::Rows, traits<B>::Rows>::value, internal::pick_size<traits<A>::Cols, traits<B>::Cols>::value> type;
};
template <typename A, typename B>
typename SumType<A, B>::type operator+(const MatrixBase<A> &a, const MatrixBase<B> &b) {
  typename SumType<A, B>::type r;
  r.resize(a.rows(), a.cols());
  for (int j = 0; j < a.cols(); j++)
    for (int i = 0; i < a.rows(); i++) r(i, j) = a.coeff(i, j) + b.coeff(i, j);
  return r;
}
template <typename A, typename B>
typename SumType<A, B>::type operator-(const MatrixBase<A> &a, const MatrixBase<B> &b) {
  typename SumType<A, B>::type r;
  r.resize(a.rows(), a.cols());
  for (int j = 0; j < a.cols(); j++)
    for (int i = 0; i < a.rows(); i++) r(i, j) = a.coeff(i, j) - b.coeff(i, j);
  return r;
}
template <typename A, typename B>
struct ProductType {
  typedef Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> type;
};
template <typename A, typename B>
typename ProductType<A, B>::type operator*(const MatrixBase<A> &a, const MatrixBase<B> &b) {
  typename ProductType<A, B>::type r;
  r.resize(a.rows(), b.cols());
  assert(a.cols() == b.rows());
  const int K = a.cols();
  for (int j = 0; j < b.cols(); j++)
    for (int i = 0; i < a.rows(); i++) {
      typename traits<A>::Scalar s = 0;
      for (int k = 0; k < K; k++) s += a.coeff(i, k) * b.coeff(k, j);
      r(i, j) = s;
    }
  return r;
}
template <typename A>
typename MatrixBase<A>::PlainObject operator*(const MatrixBase<A> &a, const typename traits<A>::Scalar &s) {
  typename MatrixBase<A>::PlainObject r(a.derived());
  r *= s;
  return r;
}
template <typename A>
typename MatrixBase<A>::PlainObject operator*(const typename traits<A>::Scalar &s, const MatrixBase<A> &a) {
  typename MatrixBase<A>::PlainObject r(a.derived());
  for (int j = 0; j < r.cols(); j++)
    for (int i = 0; i < r.rows(); i++) r(i, j) = s * a.coeff(i, j);
  return r;
}
template <typename A>
typename MatrixBase<A>::PlainObject operator/(const MatrixBase<A> &a, const typename traits<A>::Scalar &s) {
  typename MatrixBase<A>::PlainObject r(a.derived());
  r /= s;
  return r;
}
template <typename A, typename B>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, Dynamic> operator*(const MatrixBase<A> &a, const DiagonalWrapper<B> &d) {
  Matrix<typename traits<A>::Scalar, traits<A>::Rows, Dynamic> r;
  r.resize(a.rows(), a.cols());
  for (int j = 0; j < a.cols(); j++)
    for (int i = 0; i < a.rows(); i++) r(i, j) = a.coeff(i, j) * d.v.coeff(j);
  return r;
}
template <typename A, typename B>
Matrix<typename traits<B>::Scalar, Dynamic, traits<B>::Cols> operator*(const DiagonalWrapper<A> &d, const MatrixBase<B> &b) {
  Matrix<typename traits<B>::Scalar, Dynamic, traits<B>::Cols> r;
  r.resize(b.rows(), b.cols());
  for (int j = 0; j < b.cols(); j++)
    for (int i = 0; i < b.rows(); i++) r(i, j) = d.v.coeff(i) * b.coeff(i, j);
  return r;
}
template <typename A, typename B>
bool operator==(const MatrixBase<A> &a, const MatrixBase<B> &b) {
  if (a.rows() != b.rows() || a.cols() != b.cols()) return false;
  for (int j = 0; j < a.cols(); j++)
    for (int i = 0; i < a.rows(); i++)
      if (a.coeff(i, j) != b.coeff(i, j)) return false;
  return true;
}
template <typename A, typename B>
bool operator!=(const MatrixBase<A> &a, const MatrixBase<B> &b) {
  return !(a == b);
}
template <typename D>
std::ostream &operator<<(std::ostream &os, const MatrixBase<D> &m) {
  for (int i = 0; i < m.rows(); i++) {
    for (int j = 0; j < m.cols(); j++) os << (j ? " " : "") << m.coeff(i, j);
    if (i + 1 < m.rows()) os << "\n";
  }
  return os;
}

// ---- inverse / determinant (closed forms like Eigen for sizes <= 4 (cofactors), partial-pivoting LU otherwise) -----------
namespace internal {
template <typename M>
typename M::Scalar det_lu(M a) {
  typedef typename M::Scalar S;
  const int n = a.rows();
  S det = 1;
  for (int k = 0; k < n; k++) {
    int p = k;
    for (int i = k + 1; i < n; i++)
      if (std::abs(a(i, k)) > std::abs(a(p, k))) p = i;
    if (a(p, k) == S(0)) return S(0);
    if (p != k) {
      for (int j = 0; j < n; j++) std::swap(a(k, j), a(p, j));
      det = -det;
    }
    det *= a(k, k);
    for (int i = k + 1; i < n; i++) {
      const S f = a(i, k) / a(k, k);
      for (int j = k + 1; j < n; j++) a(i, j) -= f * a(k, j);
    }
  }
  return det;
}
template <typename M>
M inverse_lu(M a) {
  typedef typename M::Scalar S;
  const int n = a.rows();
  M inv;
  inv.resize(n, n);
  inv.setIdentity();
  for (int k = 0; k < n; k++) {
    int p = k;
    for (int i = k + 1; i < n; i++)
      if (std::abs(a(i, k)) > std::abs(a(p, k))) p = i;
    if (p != k)
      for (int j = 0; j < n; j++) std::swap(a(k, j), a(p, j)), std::swap(inv(k, j), inv(p, j));
    const S d = S(1) / a(k, k);
    for (int j = 0; j < n; j++) a(k, j) *= d, inv(k, j) *= d;
    for (int i = 0; i < n; i++)
      if (i != k) {
        const S f = a(i, k);
        if (f != S(0))
          for (int j = 0; j < n; j++) a(i, j) -= f * a(k, j), inv(i, j) -= f * inv(k, j);
      }
  }
  return inv;
}
template <typename M>
typename M::Scalar cofactor3(const M &m, int i, int j) {
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
}
}  // namespace internal
template <typename D>
typename MatrixBase<D>::Scalar MatrixBase<D>::determinant() const {
  PlainObject m(derived());
  const int n = rows();
  if (n == 1) return m(0, 0);
  if (n == 2) return m(0, 0) * m(1, 1) - m(1, 0) * m(0, 1);
  if (n == 3) return m(0, 0) * internal::cofactor3(m, 0, 0) + m(1, 0) * internal::cofactor3(m, 1, 0) + m(2, 0) * internal::cofactor3(m, 2, 0);
  return internal::det_lu(m);
}
template <typename D>
typename MatrixBase<D>::PlainObject MatrixBase<D>::inverse() const {
  PlainObject m(derived()), r;
  const int n = rows();
  r.resize(n, n);
  if (n == 1) {
    r(0, 0) = Scalar(1) / m(0, 0);
  } else if (n == 2) {
    const Scalar invdet = Scalar(1) / (m(0, 0) * m(1, 1) - m(1, 0) * m(0, 1));
    r(0, 0) = m(1, 1) * invdet;
    r(1, 0) = -m(1, 0) * invdet;
    r(0, 1) = -m(0, 1) * invdet;
    r(1, 1) = m(0, 0) * invdet;
  } else if (n == 3) {  // Eigen: compute_inverse_size3_helper (cofactors of the first column give the determinant)
    Scalar c0[3] = {internal::cofactor3(m, 0, 0), internal::cofactor3(m, 1, 0), internal::cofactor3(m, 2, 0)};
    const Scalar det = c0[0] * m(0, 0) + c0[1] * m(1, 0) + c0[2] * m(2, 0);
    const Scalar invdet = Scalar(1) / det;
    r(0, 0) = c0[0] * invdet, r(0, 1) = c0[1] * invdet, r(0, 2) = c0[2] * invdet;
    r(1, 0) = internal::cofactor3(m, 0, 1) * invdet, r(1, 1) = internal::cofactor3(m, 1, 1) * invdet, r(1, 2) = internal::cofactor3(m, 2, 1) * invdet;
    r(2, 0) = internal::cofactor3(m, 0, 2) * invdet, r(2, 1) = internal::cofactor3(m, 1, 2) * invdet, r(2, 2) = internal::cofactor3(m, 2, 2) * invdet;
  } else {
    r = internal::inverse_lu(m);
  }
  return r;
}

// ---- Cholesky -------------------------------------------------------------------------------------------------------
// LDLT with diagonal pivoting and sign tracking, following Eigen 3.2's ldlt_inplace<Lower>::unblocked and LDLT::solve.
template <typename MatrixType, int UpLo>
class LDLT {
  typedef typename MatrixType::Scalar S;
  MatrixType m;
  std::vector<int> tr;  // transpositions
  int sign;             // 1 positive semi-definite, -1 negative, 0 zero, 2 indefinite
  bool ok;

 public:
  LDLT() : sign(0), ok(false) {}
  explicit LDLT(const MatrixType &a) { compute(a); }
  LDLT &compute(const MatrixType &a) {
    m = a;
    const int n = m.rows();
    tr.assign(n, 0);
    sign = 0;
    ok = true;
    if (n <= 1) {
      if (n == 1) {
        tr[0] = 0;
        sign = m(0, 0) > 0 ? 1 : (m(0, 0) < 0 ? -1 : 0);
      }
      return *this;
    }
    bool found_zero_pivot = false;
    for (int k = 0; k < n; k++) {
      int p = k;
      S big = std::abs(m(k, k));
      for (int i = k + 1; i < n; i++)
        if (std::abs(m(i, i)) > big) big = std::abs(m(i, i)), p = i;
      tr[k] = p;
      if (p != k) {  // symmetric swap of rows/columns k and p of the lower triangle
        const int s = n - p - 1;
        for (int j = 0; j < k; j++) std::swap(m(k, j), m(p, j));
        for (int i = 0; i < s; i++) std::swap(m(p + 1 + i, k), m(p + 1 + i, p));
        std::swap(m(k, k), m(p, p));
        for (int i = k + 1; i < p; i++) std::swap(m(i, k), m(p, i));
      }
      const int rs = n - k - 1;
      if (k > 0) {
        // temp = A10 * D, diagonal update, column update
        std::vector<S> temp(k);
        for (int j = 0; j < k; j++) temp[j] = m(j, j) * m(k, j);
        S d = m(k, k);
        for (int j = 0; j < k; j++) d -= m(k, j) * temp[j];
        m(k, k) = d;
        for (int i = 0; i < rs; i++) {
          S v = m(k + 1 + i, k);
          for (int j = 0; j < k; j++) v -= m(k + 1 + i, j) * temp[j];
          m(k + 1 + i, k) = v;
        }
      }
      const S piv = m(k, k);
      const bool pivot_is_valid = std::abs(piv) > S(0);
      if (k == 0 && !pivot_is_valid) {  // the whole matrix is zero
        sign = 0;
        for (int j = 0; j < n; j++) tr[j] = j;
        return *this;
      }
      if (rs > 0 && pivot_is_valid)
        for (int i = 0; i < rs; i++) m(k + 1 + i, k) /= piv;
      else if (rs > 0)
        for (int i = 0; i < rs; i++)
          if (m(k + 1 + i, k) != S(0)) ok = false;
      if (found_zero_pivot && pivot_is_valid) sign = 2;
      else if (!pivot_is_valid) found_zero_pivot = true;
      if (sign == 1) {
        if (piv < 0) sign = 2;
      } else if (sign == -1) {
        if (piv > 0) sign = 2;
      } else if (sign == 0) {
        if (piv > 0) sign = 1;
        else if (piv < 0) sign = -1;
      }
    }
    return *this;
  }
  bool isPositive() const { return sign == 1 || sign == 0; }
  bool isNegative() const { return sign == -1 || sign == 0; }
  ComputationInfo info() const { return ok ? Success : NumericalIssue; }
  Matrix<S, MatrixType::RowsAtCompileTime, 1> vectorD() const {
    Matrix<S, MatrixType::RowsAtCompileTime, 1> d;
    d.resize(m.rows(), 1);
    for (int i = 0; i < m.rows(); i++) d(i) = m(i, i);
    return d;
  }
  const MatrixType &matrixLDLT() const { return m; }
  template <typename Rhs>
  typename MatrixBase<Rhs>::PlainObject solve(const MatrixBase<Rhs> &b) const {
    typename MatrixBase<Rhs>::PlainObject x(b.derived());
    const int n = m.rows(), nc = x.cols();
    for (int k = 0; k < n; k++)
      if (tr[k] != k)
        for (int j = 0; j < nc; j++) std::swap(x(k, j), x(tr[k], j));
    for (int j = 0; j < nc; j++)
      for (int i = 0; i < n; i++) {  // L y = P b (unit lower)
        S v = x(i, j);
        for (int k = 0; k < i; k++) v -= m(i, k) * x(k, j);
        x(i, j) = v;
      }
    S dmax = 0;
    for (int i = 0; i < n; i++) dmax = (std::max)(dmax, std::abs(m(i, i)));
    const S tol = (std::max)(dmax * NumTraits<S>::epsilon(), S(1) / NumTraits<S>::highest());
    for (int i = 0; i < n; i++)
      for (int j = 0; j < nc; j++) {
        if (std::abs(m(i, i)) > tol) x(i, j) /= m(i, i);
        else x(i, j) = 0;
      }
    for (int j = 0; j < nc; j++)
      for (int i = n - 1; i >= 0; i--) {  // L^T z = y
        S v = x(i, j);
        for (int k = i + 1; k < n; k++) v -= m(k, i) * x(k, j);
        x(i, j) = v;
      }
    for (int k = n - 1; k >= 0; k--)
      if (tr[k] != k)
        for (int j = 0; j < nc; j++) std::swap(x(k, j), x(tr[k], j));
    return x;
  }
};
template <typename MatrixType, int UpLo>
class LLT {
  typedef typename MatrixType::Scalar S;
  MatrixType m;
  bool ok;

 public:
  LLT() : ok(false) {}
  explicit LLT(const MatrixType &a) { compute(a); }
  LLT &compute(const MatrixType &a) {
    m = a;
    ok = true;
    const int n = m.rows();
    for (int k = 0; k < n; k++) {
      S d = m(k, k);
      for (int j = 0; j < k; j++) d -= m(k, j) * m(k, j);
      if (!(d > S(0))) {
        ok = false;
        return *this;
      }
      d = std::sqrt(d);
      m(k, k) = d;
      for (int i = k + 1; i < n; i++) {
        S v = m(i, k);
        for (int j = 0; j < k; j++) v -= m(i, j) * m(k, j);
        m(i, k) = v / d;
      }
    }
    return *this;
  }
  ComputationInfo info() const { return ok ? Success : NumericalIssue; }
  MatrixType matrixL() const {
    MatrixType l(m);
    for (int j = 0; j < l.cols(); j++)
      for (int i = 0; i < j; i++) l(i, j) = 0;
    return l;
  }
  template <typename Rhs>
  typename MatrixBase<Rhs>::PlainObject solve(const MatrixBase<Rhs> &b) const {
    typename MatrixBase<Rhs>::PlainObject x(b.derived());
    const int n = m.rows();
    for (int j = 0; j < x.cols(); j++) {
      for (int i = 0; i < n; i++) {
        S v = x(i, j);
        for (int k = 0; k < i; k++) v -= m(i, k) * x(k, j);
        x(i, j) = v / m(i, i);
      }
      for (int i = n - 1; i >= 0; i--) {
        S v = x(i, j);
        for (int k = i + 1; k < n; k++) v -= m(k, i) * x(k, j);
        x(i, j) = v / m(i, i);
      }
    }
    return x;
  }
};
template <typename D>
LDLT<typename MatrixBase<D>::PlainObject> MatrixBase<D>::ldlt() const {
  return LDLT<PlainObject>(PlainObject(derived()));
}
template <typename D>
LLT<typename MatrixBase<D>::PlainObject> MatrixBase<D>::llt() const {
  return LLT<PlainObject>(PlainObject(derived()));
}
// symmetric eigenvalues (Jacobi) -- only eigenvalues() is used (g2o's batch statistics / debugging helpers)
template <typename MatrixType>
class SelfAdjointEigenSolver {
  typedef typename MatrixType::Scalar S;
  Matrix<S, MatrixType::RowsAtCompileTime, 1> ev;

 public:
  SelfAdjointEigenSolver() {}
  explicit SelfAdjointEigenSolver(const MatrixType &a, int = ComputeEigenvectors) { compute(a); }
  SelfAdjointEigenSolver &compute(const MatrixType &a0, int = ComputeEigenvectors) {
    MatrixType a(a0);
    const int n = a.rows();
    for (int j = 0; j < n; j++)
      for (int i = 0; i < j; i++) a(i, j) = a(j, i);
    for (int sweep = 0; sweep < 64; sweep++) {
      S off = 0;
      for (int p = 0; p < n; p++)
        for (int q = p + 1; q < n; q++) off += a(p, q) * a(p, q);
      if (off < S(1e-300)) break;
      for (int p = 0; p < n; p++)
        for (int q = p + 1; q < n; q++) {
          if (a(p, q) == S(0)) continue;
          const S theta = (a(q, q) - a(p, p)) / (2 * a(p, q));
          const S t = (theta >= 0 ? S(1) : S(-1)) / (std::abs(theta) + std::sqrt(theta * theta + 1));
          const S c = S(1) / std::sqrt(t * t + 1), s = t * c;
          for (int k = 0; k < n; k++) {
            const S akp = a(k, p), akq = a(k, q);
            a(k, p) = c * akp - s * akq, a(k, q) = s * akp + c * akq;
          }
          for (int k = 0; k < n; k++) {
            const S apk = a(p, k), aqk = a(q, k);
            a(p, k) = c * apk - s * aqk, a(q, k) = s * apk + c * aqk;
          }
        }
    }
    ev.resize(n, 1);
    std::vector<S> v(n);
    for (int i = 0; i < n; i++) v[i] = a(i, i);
    std::sort(v.begin(), v.end());
    for (int i = 0; i < n; i++) ev(i) = v[i];
    return *this;
  }
  const Matrix<S, MatrixType::RowsAtCompileTime, 1> &eigenvalues() const { return ev; }
};

// ---- typedefs -------------------------------------------------------------------------------------------------------
#define PPO_ME_TYPEDEFS(T, S)                   \
  typedef Matrix<T, 2, 2> Matrix2##S;           \
  typedef Matrix<T, 3, 3> Matrix3##S;           \
  typedef Matrix<T, 4, 4> Matrix4##S;           \
  typedef Matrix<T, Dynamic, Dynamic> MatrixX##S; \
  typedef Matrix<T, 2, Dynamic> Matrix2X##S;    \
  typedef Matrix<T, 3, Dynamic> Matrix3X##S;    \
  typedef Matrix<T, 4, Dynamic> Matrix4X##S;    \
  typedef Matrix<T, Dynamic, 2> MatrixX2##S;    \
  typedef Matrix<T, Dynamic, 3> MatrixX3##S;    \
  typedef Matrix<T, 2, 1> Vector2##S;           \
  typedef Matrix<T, 3, 1> Vector3##S;           \
  typedef Matrix<T, 4, 1> Vector4##S;           \
  typedef Matrix<T, Dynamic, 1> VectorX##S;     \
  typedef Matrix<T, 1, 2> RowVector2##S;        \
  typedef Matrix<T, 1, 3> RowVector3##S;        \
  typedef Matrix<T, 1, 4> RowVector4##S;        \
  typedef Matrix<T, 1, Dynamic> RowVectorX##S;
PPO_ME_TYPEDEFS(double, d)
PPO_ME_TYPEDEFS(float, f)
PPO_ME_TYPEDEFS(int, i)
#undef PPO_ME_TYPEDEFS

// ---- geometry -------------------------------------------------------------------------------------------------------
template <typename S>
class AngleAxis;
template <typename S, int Options = 0>
class Quaternion {
  Matrix<S, 4, 1> c;  // x y z w

 public:
  typedef S Scalar;
  typedef Matrix<S, 3, 1> Vector3;
  typedef Matrix<S, 3, 3> Matrix3;
  typedef Matrix<S, 4, 1> Coefficients;
  Quaternion() {}
  Quaternion(const S &w, const S &x, const S &y, const S &z) { c(0) = x, c(1) = y, c(2) = z, c(3) = w; }
  explicit Quaternion(const S *data) { c = Coefficients(data); }
  Quaternion(const Quaternion &o) : c(o.c) {}
  template <typename O>
  explicit Quaternion(const Quaternion<O> &o) {
    c(0) = S(o.x()), c(1) = S(o.y()), c(2) = S(o.z()), c(3) = S(o.w());
  }
  explicit Quaternion(const AngleAxis<S> &aa) { *this = aa; }
  template <typename D>
  explicit Quaternion(const MatrixBase<D> &m) {
    *this = m;
  }
  Quaternion &operator=(const Quaternion &o) {
    c = o.c;
    return *this;
  }
  Quaternion &operator=(const AngleAxis<S> &aa);
  template <typename D>
  Quaternion &operator=(const MatrixBase<D> &m) {
    if (m.rows() == 4 && m.cols() == 1) {
      for (int i = 0; i < 4; i++) c(i) = m.coeff(i);
      return *this;
    }
    // rotation matrix -> quaternion ("Quaternion Calculus and Fast Animation", Shoemake; Eigen's quaternionbase_assign_impl)
    S t = m.coeff(0, 0) + m.coeff(1, 1) + m.coeff(2, 2);
    if (t > S(0)) {
      t = std::sqrt(t + S(1.0));
      w() = S(0.5) * t;
      t = S(0.5) / t;
      x() = (m.coeff(2, 1) - m.coeff(1, 2)) * t;
      y() = (m.coeff(0, 2) - m.coeff(2, 0)) * t;
      z() = (m.coeff(1, 0) - m.coeff(0, 1)) * t;
    } else {
      int i = 0;
      if (m.coeff(1, 1) > m.coeff(0, 0)) i = 1;
      if (m.coeff(2, 2) > m.coeff(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(m.coeff(i, i) - m.coeff(j, j) - m.coeff(k, k) + S(1.0));
      c(i) = S(0.5) * t;
      t = S(0.5) / t;
      w() = (m.coeff(k, j) - m.coeff(j, k)) * t;
      c(j) = (m.coeff(j, i) + m.coeff(i, j)) * t;
      c(k) = (m.coeff(k, i) + m.coeff(i, k)) * t;
    }
    return *this;
  }
  static Quaternion Identity() { return Quaternion(1, 0, 0, 0); }
  Quaternion &setIdentity() {
    c(0) = c(1) = c(2) = 0, c(3) = 1;
    return *this;
  }
  S x() const { return c(0); }
  S y() const { return c(1); }
  S z() const { return c(2); }
  S w() const { return c(3); }
  S &x() { return c(0); }
  S &y() { return c(1); }
  S &z() { return c(2); }
  S &w() { return c(3); }
  const Coefficients &coeffs() const { return c; }
  Coefficients &coeffs() { return c; }
  Block<Coefficients, 3, 1> vec() const { return c.template head<3>(); }
  S squaredNorm() const { return c.squaredNorm(); }
  S norm() const { return c.norm(); }
  void normalize() { c.normalize(); }
  Quaternion normalized() const {
    Quaternion q(*this);
    q.normalize();
    return q;
  }
  S dot(const Quaternion &o) const { return c.dot(o.c); }
  Quaternion conjugate() const { return Quaternion(w(), -x(), -y(), -z()); }
  Quaternion inverse() const {
    const S n2 = squaredNorm();
    if (n2 > S(0)) return Quaternion(w() / n2, -x() / n2, -y() / n2, -z() / n2);
    return Quaternion(0, 0, 0, 0);
  }
  Quaternion operator*(const Quaternion &b) const {
    const Quaternion &a = *this;
    return Quaternion(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(), a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                      a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(), a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
  }
  Quaternion &operator*=(const Quaternion &b) {
    *this = *this * b;
    return *this;
  }
  template <typename D>
  Vector3 operator*(const MatrixBase<D> &v) const { return _transformVector(Vector3(v)); }
  Vector3 _transformVector(const Vector3 &v) const {  // Eigen: 30 flops variant
    Vector3 uv = vec().cross(v);
    uv += uv;
    return v + w() * uv + vec().cross(uv);
  }
  Matrix3 toRotationMatrix() const {
    Matrix3 res;
    const S tx = S(2) * x(), ty = S(2) * y(), tz = S(2) * z();
    const S twx = tx * w(), twy = ty * w(), twz = tz * w();
    const S txx = tx * x(), txy = ty * x(), txz = tz * x();
    const S tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
    res(0, 0) = S(1) - (tyy + tzz);
    res(0, 1) = txy - twz;
    res(0, 2) = txz + twy;
    res(1, 0) = txy + twz;
    res(1, 1) = S(1) - (txx + tzz);
    res(1, 2) = tyz - twx;
    res(2, 0) = txz - twy;
    res(2, 1) = tyz + twx;
    res(2, 2) = S(1) - (txx + tyy);
    return res;
  }
  Matrix3 matrix() const { return toRotationMatrix(); }
  S angularDistance(const Quaternion &o) const {
    S d = std::abs(dot(o));
    if (d >= S(1.0)) return S(0);
    return S(2) * std::acos(d);
  }
  bool isApprox(const Quaternion &o, S prec = NumTraits<S>::dummy_precision()) const { return c.isApprox(o.c, prec); }
  template <typename N>
  Quaternion<N> cast() const { return Quaternion<N>(N(w()), N(x()), N(y()), N(z())); }
};
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

template <typename S>
class AngleAxis {
  Matrix<S, 3, 1> ax;
  S ang;

 public:
  typedef Matrix<S, 3, 1> Vector3;
  typedef Matrix<S, 3, 3> Matrix3;
  AngleAxis() : ang(0) {}
  template <typename D>
  AngleAxis(const S &angle, const MatrixBase<D> &axis) : ax(axis), ang(angle) {}
  explicit AngleAxis(const Quaternion<S> &q) { *this = q; }
  template <typename D>
  explicit AngleAxis(const MatrixBase<D> &m) {
    *this = Quaternion<S>(m);
  }
  S angle() const { return ang; }
  S &angle() { return ang; }
  const Vector3 &axis() const { return ax; }
  Vector3 &axis() { return ax; }
  AngleAxis &operator=(const Quaternion<S> &q) {  // Eigen 3.2
    const S n2 = q.vec().squaredNorm();
    if (n2 < NumTraits<S>::dummy_precision() * NumTraits<S>::dummy_precision()) {
      ang = 0;
      ax = Vector3(1, 0, 0);
    } else {
      ang = S(2) * std::acos((std::min)((std::max)(S(-1), q.w()), S(1)));
      ax = q.vec() / std::sqrt(n2);
    }
    return *this;
  }
  template <typename D>
  AngleAxis &operator=(const MatrixBase<D> &m) {
    return *this = Quaternion<S>(m);
  }
  Matrix3 toRotationMatrix() const {
    Matrix3 res;
    const Vector3 sin_axis = std::sin(ang) * ax;
    const S c = std::cos(ang);
    const Vector3 cos1_axis = (S(1) - c) * ax;
    S tmp;
    tmp = cos1_axis.x() * ax.y();
    res(0, 1) = tmp - sin_axis.z();
    res(1, 0) = tmp + sin_axis.z();
    tmp = cos1_axis.x() * ax.z();
    res(0, 2) = tmp + sin_axis.y();
    res(2, 0) = tmp - sin_axis.y();
    tmp = cos1_axis.y() * ax.z();
    res(1, 2) = tmp - sin_axis.x();
    res(2, 1) = tmp + sin_axis.x();
    Vector3 d = cos1_axis.cwiseProduct(ax);
    res(0, 0) = d(0) + c, res(1, 1) = d(1) + c, res(2, 2) = d(2) + c;
    return res;
  }
  Matrix3 matrix() const { return toRotationMatrix(); }
  AngleAxis inverse() const { return AngleAxis(-ang, ax); }
  Quaternion<S> operator*(const AngleAxis &o) const { return Quaternion<S>(*this) * Quaternion<S>(o); }
  Quaternion<S> operator*(const Quaternion<S> &o) const { return Quaternion<S>(*this) * o; }
  template <typename D>
  Vector3 operator*(const MatrixBase<D> &v) const { return toRotationMatrix() * Vector3(v); }
};
typedef AngleAxis<double> AngleAxisd;
typedef AngleAxis<float> AngleAxisf;
template <typename S, int O>
Quaternion<S, O> &Quaternion<S, O>::operator=(const AngleAxis<S> &aa) {
  const S ha = S(0.5) * aa.angle();
  w() = std::cos(ha);
  const Matrix<S, 3, 1> v = std::sin(ha) * aa.axis();
  x() = v(0), y() = v(1), z() = v(2);
  return *this;
}
template <typename S>
Quaternion<S> operator*(const Quaternion<S> &q, const AngleAxis<S> &a) {
  return q * Quaternion<S>(a);
}
template <typename D, typename S>
Matrix<S, 3, 3> operator*(const MatrixBase<D> &m, const AngleAxis<S> &a) {
  return Matrix<S, 3, 3>(m) * a.toRotationMatrix();
}

template <typename S, int Dim, int Mode, int Options = 0>
class Transform {
  Matrix<S, Dim + 1, Dim + 1> m;

 public:
  typedef Matrix<S, Dim + 1, Dim + 1> MatrixType;
  typedef Matrix<S, Dim, Dim> LinearMatrixType;
  typedef Matrix<S, Dim, 1> VectorType;
  Transform() { m.setIdentity(); }
  template <typename D>
  explicit Transform(const MatrixBase<D> &o) {
    *this = o;
  }
  explicit Transform(const Quaternion<S> &q) {
    m.setIdentity();
    linear() = q.toRotationMatrix();
  }
  template <typename D>
  Transform &operator=(const MatrixBase<D> &o) {
    if (o.rows() == Dim + 1) {
      m = o;
    } else {
      m.setIdentity();
      linear() = o;
    }
    return *this;
  }
  Transform &operator=(const Quaternion<S> &q) {
    m.setIdentity();
    linear() = q.toRotationMatrix();
    return *this;
  }
  static Transform Identity() { return Transform(); }
  void setIdentity() { m.setIdentity(); }
  const MatrixType &matrix() const { return m; }
  MatrixType &matrix() { return m; }
  Block<MatrixType, Dim, Dim> linear() const { return m.template block<Dim, Dim>(0, 0); }
  Block<MatrixType, Dim, Dim> rotation() const { return linear(); }
  Block<MatrixType, Dim, 1> translation() const { return m.template block<Dim, 1>(0, Dim); }
  S operator()(int i, int j) const { return m(i, j); }
  S &operator()(int i, int j) { return m(i, j); }
  Transform operator*(const Transform &o) const {
    Transform r;
    r.m = m * o.m;
    return r;
  }
  template <typename D>
  VectorType operator*(const MatrixBase<D> &v) const {
    return LinearMatrixType(linear()) * VectorType(v) + VectorType(translation());
  }
  Transform inverse(int = Mode) const {
    Transform r;
    if (Mode == Isometry) {
      r.linear() = LinearMatrixType(linear()).transpose();
      r.translation() = -(LinearMatrixType(r.linear()) * VectorType(translation()));
    } else {
      r.m = m.inverse();
    }
    return r;
  }
  template <typename D>
  Transform &translate(const MatrixBase<D> &t) {
    translation() = VectorType(translation()) + LinearMatrixType(linear()) * VectorType(t);
    return *this;
  }
  template <typename D>
  Transform &pretranslate(const MatrixBase<D> &t) {
    translation() = VectorType(translation()) + VectorType(t);
    return *this;
  }
  Transform &rotate(const Quaternion<S> &q) {
    linear() = LinearMatrixType(linear()) * q.toRotationMatrix();
    return *this;
  }
};
typedef Transform<double, 3, Isometry> Isometry3d;
typedef Transform<double, 2, Isometry> Isometry2d;
typedef Transform<double, 3, Affine> Affine3d;
typedef Transform<double, 2, Affine> Affine2d;
typedef Transform<float, 3, Isometry> Isometry3f;
typedef Transform<float, 3, Affine> Affine3f;

}  // namespace Eigen

#endif  // PPO_MINI_EIGEN_H
