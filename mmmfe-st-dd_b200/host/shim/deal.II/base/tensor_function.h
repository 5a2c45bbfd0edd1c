#ifndef MMMFE_SHIM_DEALII_TENSOR_FUNCTION_H
#define MMMFE_SHIM_DEALII_TENSOR_FUNCTION_H
#include <deal.II/base/function.h>

namespace dealii {

template <int rank, int dim, typename Number = double>
class TensorFunction : public FunctionTime<Number> {
 public:
  MMMFE_HD explicit TensorFunction(Number t0 = Number()) : FunctionTime<Number>(t0) {}
  MMMFE_HD virtual ~TensorFunction() {}
  MMMFE_HD virtual Tensor<rank, dim, Number> value(const Point<dim> &) const { return Tensor<rank, dim, Number>(); }
  virtual void value_list(const std::vector<Point<dim>> &pts, std::vector<Tensor<rank, dim, Number>> &vals) const {
    for (std::size_t i = 0; i < pts.size(); ++i) vals[i] = value(pts[i]);
  }
};

}  // namespace dealii
#endif
