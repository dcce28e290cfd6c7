// STUB (test infrastructure): deal.II's 32-bit index configuration (the reference's GPU builds use it).
#pragma once
namespace dealii {
namespace types {
typedef unsigned int global_dof_index;
}
}  // namespace dealii
