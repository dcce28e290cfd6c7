// STUB (test infrastructure, see oracle/README_ref.md): stands in for the reference's include/headers.h, which pulls
// in all of deal.II.  Provides ONLY what the reference translation units compiled into oracle/_ref use from it:
// a single-rank MPI query, and minimal data-only versions of dealii::IndexSet, Utilities::MPI::Partitioner,
// AffineConstraints<double> and of the reference's distributedDeviceVec alias with the member functions
// utils/constraintMatrixInfoDevice.cc calls.  No arithmetic of the reference path lives here.
#pragma once
#include <mpi.h>

#include <cassert>
#include <cmath>
#include <complex>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include <deal.II/base/types.h>

#define Assert(cond, exc) assert(cond)
#define AssertThrow(cond, exc) assert(cond)

namespace dealii {
inline int ExcMessage(const char *) { return 0; }

class IndexSet {
public:
  using ElementIterator = std::vector<types::global_dof_index>::const_iterator;
  std::vector<types::global_dof_index> idx;  // ascending
  ElementIterator begin() const { return idx.begin(); }
  ElementIterator end() const { return idx.end(); }
};

namespace Utilities {
namespace MPI {
inline unsigned int this_mpi_process(MPI_Comm) { return 0; }
inline unsigned int n_mpi_processes(MPI_Comm) { return 1; }

// owned range [start, end) + ascending ghost list; local index = owned offset, then ghosts in list order
class Partitioner {
public:
  IndexSet owned, ghosts;
  types::global_dof_index start = 0, nGlobal = 0;
  const IndexSet &locally_owned_range() const { return owned; }
  const IndexSet &ghost_indices() const { return ghosts; }
  types::global_dof_index size() const { return nGlobal; }
  unsigned int global_to_local(types::global_dof_index g) const {
    if (g >= start && g < start + owned.idx.size()) return (unsigned int)(g - start);
    auto it = std::lower_bound(ghosts.idx.begin(), ghosts.idx.end(), g);
    assert(it != ghosts.idx.end() && *it == g);
    return (unsigned int)(owned.idx.size() + (it - ghosts.idx.begin()));
  }
};
}  // namespace MPI
}  // namespace Utilities

template <typename T>
class AffineConstraints {
public:
  std::map<types::global_dof_index, std::vector<std::pair<types::global_dof_index, T>>> lines;
  std::map<types::global_dof_index, T> inhom;
  bool is_constrained(types::global_dof_index g) const { return lines.count(g) != 0; }
  T get_inhomogeneity(types::global_dof_index g) const {
    auto it = inhom.find(g);
    return it == inhom.end() ? T(0) : it->second;
  }
  const std::vector<std::pair<types::global_dof_index, T>> *get_constraint_entries(types::global_dof_index g) const {
    auto it = lines.find(g);
    return it == lines.end() ? nullptr : &it->second;
  }
};
}  // namespace dealii

namespace dftfe {
// data-only stand-in for linearAlgebra::MultiVector<T, DEVICE>: a borrowed device pointer
template <typename T>
class distributedDeviceVec {
public:
  T *ptr = nullptr;
  unsigned int nLocal = 0, nVec = 0;
  T *begin() { return ptr; }
  const T *begin() const { return ptr; }
  unsigned int localSize() const { return nLocal; }
  unsigned int numVectors() const { return nVec; }
};
}  // namespace dftfe
