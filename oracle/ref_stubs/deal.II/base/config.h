// STUB (test infrastructure): deal.II is not installed in this image.
#pragma once
#include <complex>
#include <mpi.h>
