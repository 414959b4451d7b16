/* TEST INFRASTRUCTURE — plain-C restatement of the reference CP-PPD inner loop.
 *
 * Oracle / CPU baseline only (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline and
 * --impl reference legs).  The product never links or loads this file.
 *
 * Follows pysparselp/ChambollePockPPD.py of the reference:
 *   cpppd_c_primal : :198-228   d = c + A_eq^T y_eq + A_ineq^T y_ineq ; x2 = clip(x - T d) ;
 *                               xbar = (1+theta) x2 - theta x ; x = x2
 *   cpppd_c_dual   : :231-240, :333-341   r = A xbar - b ; y += Sigma r ; y_ineq = max(y_ineq, 0)
 *   cpppd_c_precond: :122-179   diag_t, diag_sigma
 * Per column (row) the sum is accumulated sequentially in row (stored) order from 0.0, the
 * order scipy's csc_matvec (csr_matvec) uses, and the file is compiled with
 * -ffp-contract=off, so results are bit-identical to the numpy/scipy reference; columns / rows
 * are independent, so OpenMP threads do not change any result.
 *
 * Matrices: A = [A_eq; A_ineq] stacked, given twice: CSR (rowptr, colidx, val) and CSC with
 * entries of a column sorted by row (colptr, rowidx, cval).
 */
#include <math.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Launchers such as torchrun export OMP_NUM_THREADS=1 to their children; the CPU baseline must not
 * silently inherit that.  threads <= 0 leaves the runtime's setting alone.  Returns the thread count the
 * parallel loops below will use. */
int cpppd_c_set_threads(int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
  return omp_get_max_threads();
#else
  (void)threads;
  return 1;
#endif
}

void cpppd_c_primal(int64_t n, int64_t m_eq, int has_eq, int has_ineq, const int64_t *colptr,
                    const int32_t *rowidx, const double *cval, const double *y, const double *c,
                    const double *T, const double *lb, const double *ub, double theta, double one_plus_theta,
                    double *x, double *xbar) {
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < n; ++j) {
    double s_eq = 0.0, s_in = 0.0;
    for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
      double t = cval[p] * y[rowidx[p]];
      if (rowidx[p] < m_eq) s_eq += t; else s_in += t;
    }
    double d = c[j];
    if (has_eq) d = d + s_eq;
    if (has_ineq) d = d + s_in;
    double xo = x[j];
    double x2 = xo - T[j] * d;
    if (lb[j] > x2) x2 = lb[j];
    if (ub[j] < x2) x2 = ub[j];
    xbar[j] = one_plus_theta * x2 - theta * xo;
    x[j] = x2;
  }
}

void cpppd_c_dual(int64_t m, int64_t m_eq, const int64_t *rowptr, const int32_t *colidx, const double *val,
                  const double *xbar, const double *b, const double *sigma, double *y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < m; ++i) {
    double acc = 0.0;
    for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) acc += val[p] * xbar[colidx[p]];
    double r = acc - b[i];
    double yn = y[i] + sigma[i] * r;
    if (i >= m_eq && yn < 0.0) yn = 0.0;
    y[i] = yn;
  }
}

static double abs_pow(double a, double p) {
  double v = fabs(a);
  if (p == 1.0) return v;
  if (p == 2.0) return v * v;
  if (p == 0.0) return 1.0;
  return pow(v, p);
}

void cpppd_c_precond(int64_t n, int64_t m, int64_t m_eq, int has_eq, int has_ineq, const int64_t *rowptr,
                     const double *val, const int64_t *colptr, const int32_t *rowidx, const double *cval,
                     double alpha, double *T, double *sigma) {
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < n; ++j) {
    double s_eq = 0.0, s_in = 0.0;
    for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
      double t = abs_pow(cval[p], 2.0 - alpha) * 1.0;
      if (rowidx[p] < m_eq) s_eq += t; else s_in += t;
    }
    double tmp = 0.0;
    if (has_eq) tmp = tmp + s_eq;
    if (has_ineq) tmp = tmp + s_in;
    if (tmp == 0.0) tmp = 1.0;
    T[j] = 1.0 / tmp;
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < m; ++i) {
    double acc = 0.0;
    for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) acc += abs_pow(val[p], alpha) * 1.0;
    if (acc == 0.0) acc = 1.0;
    sigma[i] = 1.0 / acc;
  }
}

void cpppd_c_iterate(int64_t iters, int64_t n, int64_t m, int64_t m_eq, int has_eq, int has_ineq,
                     const int64_t *rowptr, const int32_t *colidx, const double *val, const int64_t *colptr,
                     const int32_t *rowidx, const double *cval, const double *c, const double *b,
                     const double *lb, const double *ub, const double *T, const double *sigma, double theta,
                     double one_plus_theta, double *x, double *xbar, double *y) {
  for (int64_t k = 0; k < iters; ++k) {
    cpppd_c_primal(n, m_eq, has_eq, has_ineq, colptr, rowidx, cval, y, c, T, lb, ub, theta, one_plus_theta, x, xbar);
    cpppd_c_dual(m, m_eq, rowptr, colidx, val, xbar, b, sigma, y);
  }
}
