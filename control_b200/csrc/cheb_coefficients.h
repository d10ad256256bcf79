/* Chebyshev recurrence scalars: the same arithmetic, in the same order, as the oracle's
 * restatement (oracle/cheb.py::chebyshev_coefficients) of PETSc KSPCHEBYSHEV's first-kind recurrence (
 * SURVEY.md Appendix A.1; configured by the reference at control/control.py:1973-1982). */
#ifndef CHEB_COEFFICIENTS_H
#define CHEB_COEFFICIENTS_H
#ifdef __cplusplus
#include <vector>
static inline void cheb_coefficients(double e_min, double e_max, int steps, double *scale,
                                     std::vector<double> &omegas)
{
    *scale = 2.0 / (e_max + e_min);
    const double alpha = 1.0 - *scale * e_min;
    const double mu = 1.0 / alpha;
    const double omegaprod = 2.0 / alpha;
    double c_prev = 1.0, c_cur = mu;
    omegas.clear();
    for (int k = 2; k <= steps; ++k) {
        const double c_next = 2.0 * mu * c_cur - c_prev;
        omegas.push_back(omegaprod * c_cur / c_next);
        c_prev = c_cur;
        c_cur = c_next;
    }
}
#endif
#endif
