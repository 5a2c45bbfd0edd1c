#ifndef MMMFE_SHIM_DEALII_EXCEPTIONS_H
#define MMMFE_SHIM_DEALII_EXCEPTIONS_H
#include <deal.II/base/function.h>
#endif
