// Default problem data of the DarcyVT front-end, in the interface of the reference's inc/data.h (namespace
// vt_darcy, classes KInverse, RightHandSidePressure, PressureBoundaryValues, ExactSolution, InitialCondition), so
// that any of the reference's inc/data*.h files can be dropped in instead (build with
// MMMFE_DATA_HEADER=/path/to/data.h).  Manufactured solution of the reference's default test case:
//   p(x,y,t) = sin(8t) sin(11x) cos(11y - pi'/4),  u = -grad p,  K = I,  pi' = 3.1415 (the reference's 4-digit pi,
//   inc/data.h:101 -- kept on purpose: results must match the reference, not the textbook).
#ifndef MMMFE_DATA_DEFAULT_H
#define MMMFE_DATA_DEFAULT_H

#include <deal.II/base/function.h>
#include <deal.II/base/tensor_function.h>

namespace vt_darcy {
using namespace dealii;

// (MMMFE_HD: usable in device code too -- the data classes are also evaluated by CUDA kernels, host/data_device.cu)
namespace manufactured {
constexpr double kQuarterPi = 3.1415 / 4;
MMMFE_HD inline double sx(double x) { return std::sin(11 * x); }
MMMFE_HD inline double cx(double x) { return std::cos(11 * x); }
MMMFE_HD inline double cy(double y) { return std::cos(11 * y - kQuarterPi); }
MMMFE_HD inline double sy(double y) { return std::sin(11 * y - kQuarterPi); }
MMMFE_HD inline double pressure(double x, double y, double t) { return std::sin(8 * t) * sx(x) * cy(y); }
}  // namespace manufactured

template <int dim>
class KInverse : public TensorFunction<2, dim> {
 public:
  virtual void value_list(const std::vector<Point<dim>> &points, std::vector<Tensor<2, dim>> &values) const override {
    for (std::size_t q = 0; q < points.size(); ++q) {
      values[q].clear();
      for (int d = 0; d < dim; ++d) values[q][d][d] = 1.0;
    }
  }
};

template <int dim>
class RightHandSidePressure : public Function<dim> {
 public:
  RightHandSidePressure(const double c0 = 1.0, const double alpha = 1.0) : Function<dim>(1), c0_(c0), alpha_(alpha) {}
  virtual double value(const Point<dim> &p, const unsigned int = 0) const override {
    // c0 p_t - laplace p with c0 = 1 hard-wired, as in the reference data file
    const double t = this->get_time();
    const double shape = manufactured::sx(p[0]) * manufactured::cy(p[1]);
    return 8 * std::cos(8 * t) * shape + std::sin(8 * t) * 242 * shape;
  }

 private:
  double c0_, alpha_;
};

template <int dim>
class PressureBoundaryValues : public Function<dim> {
 public:
  PressureBoundaryValues() : Function<dim>(1) {}
  virtual double value(const Point<dim> &p, const unsigned int = 0) const override {
    // in 3-d (space-time mortar faces) the third coordinate is the time
    const double t = (dim == 3) ? p[dim - 1] : this->get_time();
    return manufactured::pressure(p[0], p[1], t);
  }
};

template <int dim>
class ExactSolution : public Function<dim> {
 public:
  ExactSolution() : Function<dim>(dim + 1) {}
  virtual void vector_value(const Point<dim> &p, Vector<double> &values) const override {
    using namespace manufactured;
    const double st = std::sin(8 * this->get_time());
    values(0) = -st * 11 * cx(p[0]) * cy(p[1]);
    values(1) = st * 11 * sx(p[0]) * sy(p[1]);
    values(2) = st * sx(p[0]) * cy(p[1]);
  }
};

template <int dim>
class InitialCondition : public Function<dim> {
 public:
  InitialCondition() : Function<dim>(dim + 1) {}
  virtual void vector_value(const Point<dim> &p, Vector<double> &values) const override {
    values(0) = 0;
    values(1) = 0;
    values(2) = manufactured::pressure(p[0], p[1], 0.0);
  }
};

}  // namespace vt_darcy
#endif
