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

namespace dealii {

template <int dim, typename Number = double>
class Point {
 public:
  Point() { for (int d = 0; d < dim; ++d) v_[d] = Number(); }
  Point(Number x, Number y) { static_assert(dim == 2, "2-d constructor"); v_[0] = x; v_[1] = y; }
  Point(Number x, Number y, Number z) { static_assert(dim == 3, "3-d constructor"); v_[0] = x; v_[1] = y; v_[2] = z; }
  Number &operator[](unsigned i) { return v_[i]; }
  const Number &operator[](unsigned i) const { return v_[i]; }
  Number &operator()(unsigned i) { return v_[i]; }
  const Number &operator()(unsigned i) const { return v_[i]; }

 private:
  Number v_[dim];
};

template <int rank, int dim, typename Number = double>
class Tensor;

template <int dim, typename Number>
class Tensor<1, dim, Number> {
 public:
  Tensor() { clear(); }
  void clear() { for (int d = 0; d < dim; ++d) v_[d] = Number(); }
  Number &operator[](unsigned i) { return v_[i]; }
  const Number &operator[](unsigned i) const { return v_[i]; }

 private:
  Number v_[dim];
};

template <int dim, typename Number>
class Tensor<2, dim, Number> {
 public:
  void clear() { for (int d = 0; d < dim; ++d) v_[d].clear(); }
  Tensor<1, dim, Number> &operator[](unsigned i) { return v_[i]; }
  const Tensor<1, dim, Number> &operator[](unsigned i) const { return v_[i]; }

 private:
  Tensor<1, dim, Number> v_[dim];
};

template <typename Number = double>
class Vector {
 public:
  Vector() = default;
  explicit Vector(std::size_t n) : v_(n, Number()) {}
  std::size_t size() const { return v_.size(); }
  Number &operator()(std::size_t i) { return v_[i]; }
  const Number &operator()(std::size_t i) const { return v_[i]; }
  Number &operator[](std::size_t i) { return v_[i]; }
  const Number &operator[](std::size_t i) const { return v_[i]; }

 private:
  std::vector<Number> v_;
};

template <typename Number = double>
class FunctionTime {
 public:
  explicit FunctionTime(Number t0 = Number()) : time_(t0) {}
  virtual ~FunctionTime() = default;
  Number get_time() const { return time_; }
  virtual void set_time(const Number t) { time_ = t; }
  virtual void advance_time(const Number dt) { set_time(time_ + dt); }

 private:
  Number time_;
};

template <int dim, typename Number = double>
class Function : public FunctionTime<Number> {
 public:
  const unsigned int n_components;
  explicit Function(unsigned int n_comp = 1, Number t0 = Number()) : FunctionTime<Number>(t0), n_components(n_comp) {}
  virtual ~Function() = default;
  virtual Number value(const Point<dim> &, const unsigned int = 0) const { return Number(); }
  virtual void vector_value(const Point<dim> &p, Vector<Number> &values) const {
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
