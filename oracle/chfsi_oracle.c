/*
 * C restatement of the reference's CPU Chebyshev-filter path - TEST INFRASTRUCTURE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this file's shared object; the product
 * (dftfe_b200/) never does.  PARITY UNPINNED at kernel granularity (see the
 * header of oracle/chfsi_oracle.py): the reference ships no golden vectors for
 * this path and cannot be built in this image, so this is a line-by-line port of
 * its CPU twin, cross-checked against the independent numpy restatement and
 * against closed-form answers.
 *
 * Follows (paths relative to dftfeDevelopers/dftfe):
 *   src/dftOperator/matrixVectorProductImplementations.cc:97-169  per-cell dcopy -> dgemm -> daxpy
 *   src/dftOperator/kohnShamDFTOperator.cc:950-1044               HX with M^-1/2 scalings
 *   utils/constraintMatrixInfo.cc:247-293, 338-375, 393-411       distribute / slave->master / set_zero
 *   src/linAlg/linearAlgebraOperationsOpt.cc:276-352              three-term recurrence
 *
 * Parallelism: the reference runs one MPI rank per core over a domain
 * decomposition; here OpenMP threads work through the cells colour by colour
 * (cells of one colour share no DoF, so the daxpy assembly needs no locks), each
 * thread calling a single-threaded BLAS dgemm - the same per-cell arithmetic on
 * the same number of cores, without the ghost exchange.
 *
 * BLAS: cblas_dgemm is taken from the OpenBLAS bundled with numpy (ILP64 symbol
 * scipy_cblas_dgemm64_) via dlopen; a plain blocked loop is the fallback.
 */
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef void (*cblas_dgemm64_fn)(int order, int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha,
                                 const double *a, int64_t lda, const double *b, int64_t ldb, double beta, double *c,
                                 int64_t ldc);
typedef void (*set_threads_fn)(int);

static cblas_dgemm64_fn g_dgemm = NULL;

int oracle_init_blas(const char *path) {
  void *h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) return -1;
  g_dgemm = (cblas_dgemm64_fn)dlsym(h, "scipy_cblas_dgemm64_");
  set_threads_fn st = (set_threads_fn)dlsym(h, "scipy_openblas_set_num_threads64_");
  if (st) st(1); /* one BLAS thread per OpenMP thread, like one MPI rank per core */
  return g_dgemm ? 0 : -2;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* Y(n x B) = H(n x n, row-major as stored: mem[I*n+J]) * X(n x B); column-major
 * view: cellY(B x n) = cellX(B x n) * H_cm(n x n)  == dgemm('N','N',B,n,n) of the reference. */
static void cell_gemm(int n, int B, double alpha, const double *H, const double *X, double *Y) {
  if (g_dgemm) {
    /* CblasColMajor=102, CblasNoTrans=111 */
    g_dgemm(102, 111, 111, B, n, n, alpha, X, B, H, n, 0.0, Y, B);
    return;
  }
  for (int i = 0; i < n; ++i) {
    double *y = Y + (size_t)i * B;
    for (int j = 0; j < B; ++j) y[j] = 0.0;
    for (int k = 0; k < n; ++k) {
      const double h = alpha * H[(size_t)i * n + k];
      const double *x = X + (size_t)k * B;
      for (int j = 0; j < B; ++j) y[j] += h * x[j];
    }
  }
}

/* matrixVectorProductImplementations.cc:97-169.  cells are visited colour by colour. */
void oracle_local_hx(int n, int B, int64_t nCells, const double *H, const uint32_t *cellRows, int nColours,
                     const int32_t *colourStart, const int32_t *colourCells, const double *src, double *dst,
                     double scalar) {
#pragma omp parallel
  {
    double *cx = (double *)malloc(sizeof(double) * (size_t)n * B);
    double *cy = (double *)malloc(sizeof(double) * (size_t)n * B);
    for (int col = 0; col < nColours; ++col) {
#pragma omp for schedule(dynamic, 1)
      for (int32_t q = colourStart[col]; q < colourStart[col + 1]; ++q) {
        const int64_t c = colourCells[q];
        const uint32_t *rows = cellRows + c * n;
        for (int i = 0; i < n; ++i) memcpy(cx + (size_t)i * B, src + (size_t)rows[i] * B, sizeof(double) * B);
        cell_gemm(n, B, scalar, H + (size_t)c * n * n, cx, cy);
        for (int i = 0; i < n; ++i) {
          double *d = dst + (size_t)rows[i] * B;
          const double *y = cy + (size_t)i * B;
          for (int j = 0; j < B; ++j) d[j] += y[j];
        }
      }
    }
    free(cx);
    free(cy);
  }
  (void)nCells;
}

/* acc[i] = fma(w, x[i], acc[i]): exact fused multiply-add for the numpy oracle (numpy has none) */
void oracle_fma_axpy(int64_t n, double w, const double *x, double *acc) {
  for (int64_t i = 0; i < n; ++i) acc[i] = fma(w, x[i], acc[i]);
}

/* utils/constraintMatrixInfo.cc:247-293 */
void oracle_distribute(int B, int64_t nCon, const uint32_t *rows, const uint32_t *sizes, const uint32_t *starts,
                       const uint32_t *cols, const double *vals, const double *inhom, double *x) {
  double *tmp = (double *)malloc(sizeof(double) * B);
  for (int64_t i = 0; i < nCon; ++i) {
    for (int j = 0; j < B; ++j) tmp[j] = inhom[i];
    for (uint32_t k = 0; k < sizes[i]; ++k) {
      const double w = vals[starts[i] + k];
      const double *xc = x + (size_t)cols[starts[i] + k] * B;
      /* one rounding per term: the device kernel's `x[row] += w * x[col]` is a DFMA under nvcc's default
       * -fmad=true (utils/constraintMatrixInfoDevice.cc:71-74); the same as the numpy oracle (oracle_fma_axpy) */
      for (int j = 0; j < B; ++j) tmp[j] = fma(w, xc[j], tmp[j]);
    }
    memcpy(x + (size_t)rows[i] * B, tmp, sizeof(double) * B);
  }
  free(tmp);
}

/* utils/constraintMatrixInfo.cc:338-375 */
void oracle_slave_to_master(int B, int64_t nCon, const uint32_t *rows, const uint32_t *sizes, const uint32_t *starts,
                            const uint32_t *cols, const double *vals, double *x) {
  for (int64_t i = 0; i < nCon; ++i) {
    double *xr = x + (size_t)rows[i] * B;
    for (uint32_t k = 0; k < sizes[i]; ++k) {
      const double w = vals[starts[i] + k];
      double *xc = x + (size_t)cols[starts[i] + k] * B;
      for (int j = 0; j < B; ++j) {
        volatile double prod = w * xr[j];
        xc[j] = xc[j] + prod;
      }
    }
    for (int j = 0; j < B; ++j) xr[j] = 0.0;
  }
}

static void row_scale(int B, int64_t rows, const double *s, double alpha, double *x) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; ++r) {
    const double f = alpha * s[r];
    double *xr = x + (size_t)r * B;
    for (int j = 0; j < B; ++j) xr[j] *= f;
  }
}

typedef struct {
  int n, B;
  int64_t nCells, M, G, nCon;
  const double *H;
  const uint32_t *cellRows;
  int nColours;
  const int32_t *colourStart, *colourCells;
  const uint32_t *conRows, *conSizes, *conStarts, *conCols;
  const double *conVals, *conInhom;
  const double *sqrtM, *invSqrtM;
} oracle_problem;

/* kohnShamDFTOperator.cc:950-1044, single rank (no ghosts to exchange) */
void oracle_hx(const oracle_problem *p, double *src, double *dst, int scaleFlag, double scalar) {
  row_scale(p->B, p->M, p->invSqrtM, scalar, src);
  if (scaleFlag) row_scale(p->B, p->M, p->sqrtM, 1.0, dst);
  oracle_distribute(p->B, p->nCon, p->conRows, p->conSizes, p->conStarts, p->conCols, p->conVals, p->conInhom, src);
  oracle_local_hx(p->n, p->B, p->nCells, p->H, p->cellRows, p->nColours, p->colourStart, p->colourCells, src, dst,
                  1.0);
  oracle_slave_to_master(p->B, p->nCon, p->conRows, p->conSizes, p->conStarts, p->conCols, p->conVals, dst);
  row_scale(p->B, p->M, p->invSqrtM, 1.0, dst);
  row_scale(p->B, p->M, p->sqrtM, 1.0 / scalar, src);
}

/* linearAlgebraOperationsOpt.cc:276-352; X in/out, Y scratch, both (M+G) x B */
void oracle_cheb_filter(const oracle_problem *p, double *X, double *Y, int m, double a, double b, double a0) {
  const int64_t tot = (p->M + p->G) * (int64_t)p->B;
  double e = (b - a) / 2.0, c = (b + a) / 2.0;
  double sigma = e / (a0 - c), sigma1 = sigma, gamma = 2.0 / sigma1, sigma2;
  double *x = X, *y = Y;
  memset(y, 0, sizeof(double) * tot);
  oracle_hx(p, x, y, 0, 1.0);
  double alpha1 = sigma1 / e, alpha2 = -c;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < tot; ++i) y[i] = alpha1 * (y[i] + alpha2 * x[i]);
  for (int degree = 2; degree <= m; ++degree) {
    sigma2 = 1.0 / (gamma - sigma);
    alpha1 = 2.0 * sigma2 / e;
    alpha2 = -(sigma * sigma2);
    const double coeff = -c * alpha1;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < tot; ++i) x[i] = alpha2 * x[i] + coeff * y[i];
    oracle_hx(p, y, x, 1, alpha1);
    double *t = x;
    x = y;
    y = t;
    sigma = sigma2;
  }
  if (y != X) memcpy(X, y, sizeof(double) * tot);
}
