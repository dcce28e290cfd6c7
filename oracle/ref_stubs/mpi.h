// STUB (test infrastructure): single-process stand-in for <mpi.h>, enough for the type-id helpers of the reference's
// include/dftfeDataTypes.h and the one MPI_Allreduce in utils/DeviceKernelsGeneric.cc (one rank: a copy).
#pragma once
#include <cstring>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
#define MPI_COMM_WORLD 0
#define MPI_INT 4
#define MPI_LONG 8
#define MPI_UNSIGNED 4
#define MPI_UNSIGNED_LONG 8
#define MPI_UNSIGNED_LONG_LONG 8
#define MPI_DOUBLE 8
#define MPI_FLOAT 4
#define MPI_LONG_DOUBLE 16
#define MPI_C_DOUBLE_COMPLEX 16
#define MPI_C_FLOAT_COMPLEX 8
#define MPI_DOUBLE_COMPLEX 16
#define MPI_COMPLEX 8
#define MPI_SUM 0
inline int MPI_Allreduce(const void *in, void *out, int n, MPI_Datatype dt, MPI_Op, MPI_Comm) {
  std::memcpy(out, in, (size_t)n * dt);
  return 0;
}
