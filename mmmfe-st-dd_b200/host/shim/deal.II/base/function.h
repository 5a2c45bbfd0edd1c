// Minimal stand-in for the deal.II classes that an inc/data.h-style problem-data header uses, so that such a
// header compiles unmodified without deal.II.  Only what those headers touch is provided:
// Point, Tensor, Vector, FunctionTime, Function, TensorFunction, Assert and the three exception makers.
#ifndef MMMFE_SHIM_DEALII_FUNCTION_H
#define MMMFE_SHIM_DEALII_FUNCTION_H

#include <cassert>
#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

// Under nvcc every class below is usable in device code as well (host/data_device.cu compiles an unmodified
// inc/data.h-style header for the GPU: the right-hand-side, boundary and exact-solution tables are then evaluated
// by CUDA kernels instead of host loops).  g++ sees plain host classes.
#if defined(__CUDACC__)
#define MMMFE_HD __host__ __device__
#else
#define MMMFE_HD
#endif

namespace dealii {

template <int dim, typename Number = double>
class Point {
 public:
  MMMFE_HD Point() { for (int d = 0; d < dim; ++d) v_[d] = Number(); }
  MMMFE_HD Point(Number x, Number y) { static_assert(dim == 2, "2-d constructor"); v_[0] = x; v_[1] = y; }
  MMMFE_HD Point(Number x, Number y, Number z) { static_assert(dim == 3, "3-d constructor"); v_[0] = x; v_[1] = y; v_[2] = z; }
  MMMFE_HD Number &operator[](unsigned i) { return v_[i]; }
  MMMFE_HD const Number &operator[](unsigned i) const { return v_[i]; }
  MMMFE_HD Number &operator()(unsigned i) { return v_[i]; }
  MMMFE_HD const Number &operator()(unsigned i) const { return v_[i]; }

 private:
  Number v_[dim];
};

template <int rank, int dim, typename Number = double>
class Tensor;

template <int dim, typename Number>
class Tensor<1, dim, Number> {
 public:
  MMMFE_HD Tensor() { clear(); }
  MMMFE_HD void clear() { for (int d = 0; d < dim; ++d) v_[d] = Number(); }
  MMMFE_HD Number &operator[](unsigned i) { return v_[i]; }
  MMMFE_HD const Number &operator[](unsigned i) const { return v_[i]; }

 private:
  Number v_[dim];
};

template <int dim, typename Number>
class Tensor<2, dim, Number> {
 public:
  MMMFE_HD void clear() { for (int d = 0; d < dim; ++d) v_[d].clear(); }
  MMMFE_HD Tensor<1, dim, Number> &operator[](unsigned i) { return v_[i]; }
  MMMFE_HD const Tensor<1, dim, Number> &operator[](unsigned i) const { return v_[i]; }

 private:
  Tensor<1, dim, Number> v_[dim];
};

// The data headers only ever hold the components of one function value in a Vector (dim + 1 at most): inline
// storage, so that the class works in device code too.
template <typename Number = double>
class Vector {
 public:
  static constexpr std::size_t capacity = 8;
  MMMFE_HD Vector() : n_(0) { for (std::size_t i = 0; i < capacity; ++i) v_[i] = Number(); }
  MMMFE_HD explicit Vector(std::size_t n) : n_(n) {
    assert(n <= capacity);
    for (std::size_t i = 0; i < capacity; ++i) v_[i] = Number();
  }
  MMMFE_HD std::size_t size() const { return n_; }
  MMMFE_HD Number &operator()(std::size_t i) { return v_[i]; }
  MMMFE_HD const Number &operator()(std::size_t i) const { return v_[i]; }
  MMMFE_HD Number &operator[](std::size_t i) { return v_[i]; }
  MMMFE_HD const Number &operator[](std::size_t i) const { return v_[i]; }

 private:
  std::size_t n_;
  Number v_[capacity];
};

template <typename Number = double>
class FunctionTime {
 public:
  MMMFE_HD explicit FunctionTime(Number t0 = Number()) : time_(t0) {}
  MMMFE_HD virtual ~FunctionTime() {}
  MMMFE_HD Number get_time() const { return time_; }
  MMMFE_HD virtual void set_time(const Number t) { time_ = t; }
  MMMFE_HD virtual void advance_time(const Number dt) { set_time(time_ + dt); }

 private:
  Number time_;
};

template <int dim, typename Number = double>
class Function : public FunctionTime<Number> {
 public:
  const unsigned int n_components;
  MMMFE_HD explicit Function(unsigned int n_comp = 1, Number t0 = Number()) : FunctionTime<Number>(t0), n_components(n_comp) {}
  MMMFE_HD virtual ~Function() {}
  MMMFE_HD virtual Number value(const Point<dim> &, const unsigned int = 0) const { return Number(); }
  MMMFE_HD virtual void vector_value(const Point<dim> &p, Vector<Number> &values) const {
    for (unsigned c = 0; c < n_components; ++c) values(c) = value(p, c);
  }
  virtual void value_list(const std::vector<Point<dim>> &pts, std::vector<Number> &vals, const unsigned int c = 0) const {
    for (std::size_t i = 0; i < pts.size(); ++i) vals[i] = value(pts[i], c);
  }
  virtual void vector_value_list(const std::vector<Point<dim>> &pts, std::vector<Vector<Number>> &vals) const {
    for (std::size_t i = 0; i < pts.size(); ++i) vector_value(pts[i], vals[i]);
  }
  virtual void vector_gradient(const Point<dim> &, std::vector<Tensor<1, dim, Number>> &) const {}
};

// exception helpers: the data headers only construct them inside Assert(...)
struct ExcBase_ { std::string what; };
inline ExcBase_ ExcMessage(const std::string &m) { return {m}; }
inline ExcBase_ ExcNotImplemented() { return {"not implemented"}; }
template <typename A, typename B>
inline ExcBase_ ExcDimensionMismatch(const A &a, const B &b) {
  return {"dimension mismatch: " + std::to_string(a) + " != " + std::to_string(b)};
}

}  // namespace dealii

// deal.II's Assert is active in debug builds only; the reference is run as a release build (README "make release").
#ifndef Assert
#ifdef DEBUG
#define Assert(cond, exc) do { if (!(cond)) throw std::runtime_error((exc).what); } while (0)
#else
#define Assert(cond, exc) do { } while (0)
#endif
#endif
#ifndef AssertThrow
#define AssertThrow(cond, exc) do { if (!(cond)) throw std::runtime_error((exc).what); } while (0)
#endif

#endif
