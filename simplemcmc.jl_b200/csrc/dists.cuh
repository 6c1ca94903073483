// Device log-densities and their adjoints (FP64).
// Formulas: /root/reference/src/distribs.jl:12-33 (Bernoulli, Normal, Uniform, defined locally there) and the
// nine libRmath densities the reference reaches by ccall (src/distribs.jl:40-48,65-67), in the scale
// parametrisation with R's edge cases (outside support -> -inf, invalid parameter -> NaN).  dbinom, dpois, dgamma,
// dbeta (a, b > 2) and dt follow nmath's published ALGORITHM -- Loader's saddle-point form built from stirlerr()
// and bd0() -- not the textbook closed forms: those cancel catastrophically for large counts / shapes (1e-8 relative
// at lambda = 1e8, oracle/rmath.py, tests/test_oracle_rmath.py) where libRmath stays at 1e-15.  dweibull, dcauchy,
// dlnorm, dexp are single expressions in nmath and are written as such.  Adjoints: the @dfunc table, src/diff.jl:94-174.
#pragma once
#include <math.h>
#include "../../include/simplemcmc_tape.h"

#define SMC_LOG_SQRT_2PI 0.91893853320467274178
#define SMC_PI 3.14159265358979323846

// digamma for x > 0 by upward recurrence + asymptotic series (|err| < 1e-15 for x >= 10);
// reflection for x <= 0 (not a pole): psi(1-x) - pi/tan(pi x)
__device__ __forceinline__ double smc_digamma(double x) {
  if (isnan(x)) return x;
  double refl = 0.0;
  if (x <= 0.0) {
    if (x == floor(x)) return nan("");
    refl = -SMC_PI / tan(SMC_PI * x);
    x = 1.0 - x;
  }
  double acc = 0.0;
  while (x < 10.0) { acc -= 1.0 / x; x += 1.0; }
  double xi = 1.0 / x, x2 = xi * xi;
  // 1/12, 1/120, 1/252, 1/240, 1/132, 691/32760, 1/12
  double s = x2 * (1.0 / 12.0 - x2 * (1.0 / 120.0 - x2 * (1.0 / 252.0 - x2 * (1.0 / 240.0 - x2 * (1.0 / 132.0
             - x2 * (691.0 / 32760.0 - x2 * (1.0 / 12.0)))))));
  return acc + log(x) - 0.5 * xi - s + refl;
}

__device__ __forceinline__ double smc_xlogy(double x, double y) { return x == 0.0 ? 0.0 : x * log(y); }
__device__ __forceinline__ double smc_xlog1py(double x, double y) { return x == 0.0 ? 0.0 : x * log1p(y); }

#define SMC_LN_2PI 1.837877066409345483560659472811
#define SMC_2PI 6.283185307179586476925286766559
#define SMC_DBL_MIN 2.2250738585072014e-308
#define SMC_DBL_EPS 2.220446049250313e-16

// stirlerr(n) = log n! - log(sqrt(2 pi n) (n/e)^n)  (nmath/stirlerr.c): table for 2n integer, n <= 15; series beyond
static __device__ const double smc_sferr_halves[31] = {
#include "stirlerr_table.inc"
};
static __device__ __noinline__ double smc_stirlerr(double n) {
  if (n <= 15.0) {
    double nn = n + n;
    if (nn == (double)(int)nn) return smc_sferr_halves[(int)nn];
    return lgamma(n + 1.0) - (n + 0.5) * log(n) + n - SMC_LOG_SQRT_2PI;
  }
  const double S0 = 1.0 / 12.0, S1 = 1.0 / 360.0, S2 = 1.0 / 1260.0, S3 = 1.0 / 1680.0, S4 = 1.0 / 1188.0;
  double nn = n * n;
  if (n > 500.0) return (S0 - S1 / nn) / n;
  if (n > 80.0) return (S0 - (S1 - S2 / nn) / nn) / n;
  if (n > 35.0) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
  return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}
// bd0(x, M) = x log(x/M) + M - x  (nmath/bd0.c): series in v = (x-M)/(x+M) near x = M
static __device__ __noinline__ double smc_bd0(double x, double M) {
  if (!isfinite(x) || !isfinite(M) || M == 0.0) return nan("");
  if (fabs(x - M) < 0.1 * (x + M)) {
    double v = (x - M) / (x + M), s = (x - M) * v;
    if (fabs(s) < SMC_DBL_MIN) return s;
    double ej = 2.0 * x * v;
    v = v * v;
    for (int j = 1; j < 1000; j++) {
      ej *= v;
      double s1 = s + ej / (double)((j << 1) + 1);
      if (s1 == s) return s1;
      s = s1;
    }
  }
  return x * log(x / M) + M - x;
}
// log dpois_raw(x, lambda)  (nmath/dpois.c)
static __device__ __noinline__ double smc_dpois_raw(double x, double lam) {
  if (lam == 0.0) return x == 0.0 ? 0.0 : -INFINITY;
  if (!isfinite(lam) || x < 0.0) return -INFINITY;
  if (x <= lam * SMC_DBL_MIN) return -lam;
  if (lam < x * SMC_DBL_MIN) return isfinite(x) ? -lam + x * log(lam) - lgamma(x + 1.0) : -INFINITY;
  return -0.5 * log(SMC_2PI * x) + (-smc_stirlerr(x) - smc_bd0(x, lam));
}
// log dbinom_raw(x, n, p, q)  (nmath/dbinom.c)
static __device__ __noinline__ double smc_dbinom_raw(double x, double n, double p, double q) {
  if (p == 0.0) return x == 0.0 ? 0.0 : -INFINITY;
  if (q == 0.0) return x == n ? 0.0 : -INFINITY;
  if (x == 0.0) {
    if (n == 0.0) return 0.0;
    return p < 0.1 ? -smc_bd0(n, n * q) - n * p : n * log(q);
  }
  if (x == n) return q < 0.1 ? -smc_bd0(n, n * p) - n * q : n * log(p);
  if (x < 0.0 || x > n) return -INFINITY;
  double lc = smc_stirlerr(n) - smc_stirlerr(x) - smc_stirlerr(n - x) - smc_bd0(x, n * p) - smc_bd0(n - x, n * q);
  double lf = SMC_LN_2PI + log(x) + log1p(-x / n);
  return lc - 0.5 * lf;
}
// log B(a, b)  (nmath/lbeta.c; its lgammacor(x), x >= 10, is the same function as stirlerr(x))
static __device__ __noinline__ double smc_lbeta(double a, double b) {
  double p = fmin(a, b), q = fmax(a, b);
  if (p < 0.0) return nan("");
  if (p == 0.0) return INFINITY;
  if (!isfinite(q)) return -INFINITY;
  if (p >= 10.0) {
    double corr = smc_stirlerr(p) + smc_stirlerr(q) - smc_stirlerr(p + q);
    return log(q) * -0.5 + SMC_LOG_SQRT_2PI + corr + (p - 0.5) * log(p / (p + q)) + q * log1p(-p / (p + q));
  }
  if (q >= 10.0) {
    double corr = smc_stirlerr(q) - smc_stirlerr(p + q);
    return lgamma(p) + corr + p - p * log(p + q) + (q - 0.5) * log1p(-p / (p + q));
  }
  return log(tgamma(p) * (tgamma(q) / tgamma(p + q)));
}
// R_nonint
__device__ __forceinline__ bool smc_nonint(double x) { return fabs(x - rint(x)) > 1e-7 * fmax(1.0, fabs(x)); }

// per-element log density; argument order = the reference's logpdfD(params..., variate)
__device__ __forceinline__ double smc_logpdf(int dist, double a0, double a1, double a2) {
  const double ninf = -INFINITY, qnan = nan("");
  switch (dist) {
    case SMC_DIST_NORMAL: {  // (mu, sigma, x)  distribs.jl:23-28 ; sigma<=0 is an assert there, NaN here
      if (!(a1 > 0.0)) return qnan;
      double r = (a2 - a0) / a1;
      return -r * r * 0.5 - log(a1) - SMC_LOG_SQRT_2PI;
    }
    case SMC_DIST_UNIFORM:  // (a, b, x)  distribs.jl:30-33
      if (!(a0 < a1)) return qnan;
      return (a0 <= a2 && a2 <= a1) ? -log(a1 - a0) : ninf;
    case SMC_DIST_BERNOULLI:  // (p, x)  distribs.jl:12-21
      if (!(0.0 <= a0 && a0 <= 1.0)) return qnan;
      if (a1 == 0.0) return log(1.0 - a0);
      if (a1 == 1.0) return log(a0);
      return qnan;
    case SMC_DIST_WEIBULL: {  // (shape, scale, x)  dweibull
      if (!(a0 > 0.0) || !(a1 > 0.0)) return qnan;
      if (a2 < 0.0 || isinf(a2)) return ninf;
      if (a2 == 0.0 && a0 < 1.0) return INFINITY;
      double t1 = pow(a2 / a1, a0 - 1.0), t2 = t1 * (a2 / a1);
      return -t2 + log(a0 * t1 / a1);
    }
    case SMC_DIST_BINOMIAL: {  // (n, p, x)  dbinom
      if (isnan(a0) || isnan(a1) || isnan(a2)) return qnan;
      if (!(a1 >= 0.0 && a1 <= 1.0) || a0 < 0.0 || smc_nonint(a0)) return qnan;
      if (a2 < 0.0 || smc_nonint(a2) || isinf(a2)) return ninf;
      return smc_dbinom_raw(rint(a2), rint(a0), a1, 1.0 - a1);
    }
    case SMC_DIST_GAMMA: {  // (shape, scale, x)  dgamma
      if (isnan(a0) || isnan(a1) || isnan(a2)) return qnan;
      if (a0 < 0.0 || !(a1 > 0.0)) return qnan;
      if (a2 < 0.0) return ninf;
      if (a0 == 0.0) return a2 == 0.0 ? INFINITY : ninf;
      if (a2 == 0.0) return a0 < 1.0 ? INFINITY : (a0 > 1.0 ? ninf : -log(a1));
      if (a0 < 1.0) {  // dpois_raw(shape, x/scale) * shape / x
        double pr = smc_dpois_raw(a0, a2 / a1), r = a0 / a2;
        return pr + (isfinite(r) ? log(r) : log(a0) - log(a2));
      }
      return smc_dpois_raw(a0 - 1.0, a2 / a1) - log(a1);
    }
    case SMC_DIST_CAUCHY: {  // (location, scale, x)  dcauchy
      if (!(a1 > 0.0)) return qnan;
      double y = (a2 - a0) / a1;
      return -log(SMC_PI * a1 * (1.0 + y * y));
    }
    case SMC_DIST_LOGNORMAL: {  // (meanlog, sdlog, x)  dlnorm
      if (!(a1 > 0.0)) return qnan;
      if (a2 <= 0.0) return ninf;
      double y = (log(a2) - a0) / a1;
      return -(SMC_LOG_SQRT_2PI + 0.5 * y * y + log(a2 * a1));
    }
    case SMC_DIST_BETA: {  // (a, b, x)  dbeta
      if (isnan(a0) || isnan(a1) || isnan(a2)) return qnan;
      if (!(a0 > 0.0) || !(a1 > 0.0)) return qnan;
      if (a2 < 0.0 || a2 > 1.0) return ninf;
      if (a2 == 0.0) return a0 > 1.0 ? ninf : (a0 < 1.0 ? INFINITY : log(a1));
      if (a2 == 1.0) return a1 > 1.0 ? ninf : (a1 < 1.0 ? INFINITY : log(a0));
      if (a0 <= 2.0 || a1 <= 2.0)
        return (a0 - 1.0) * log(a2) + (a1 - 1.0) * log1p(-a2) - smc_lbeta(a0, a1);
      return log(a0 + a1 - 1.0) + smc_dbinom_raw(a0 - 1.0, a0 + a1 - 2.0, a2, 1.0 - a2);
    }
    case SMC_DIST_POISSON:  // (lambda, x)  dpois
      if (isnan(a0) || isnan(a1) || a0 < 0.0) return qnan;
      if (a1 < 0.0 || smc_nonint(a1) || isinf(a1)) return ninf;
      return smc_dpois_raw(rint(a1), a0);
    case SMC_DIST_TDIST: {  // (df, x)  dt
      if (isnan(a1) || !(a0 > 0.0)) return qnan;
      if (isinf(a1)) return ninf;
      if (isinf(a0)) return -(SMC_LOG_SQRT_2PI + 0.5 * a1 * a1);
      double t = -smc_bd0(a0 * 0.5, (a0 + 1.0) * 0.5) + smc_stirlerr((a0 + 1.0) * 0.5) - smc_stirlerr(a0 * 0.5);
      double x2n = a1 * a1 / a0, l_x2n, u;
      if (x2n > 1.0 / SMC_DBL_EPS) {
        l_x2n = log(fabs(a1)) - log(a0) * 0.5;
        u = a0 * l_x2n;
      } else if (x2n > 0.2) {
        l_x2n = log(1.0 + x2n) * 0.5;
        u = a0 * l_x2n;
      } else {
        l_x2n = log1p(x2n) * 0.5;
        u = -smc_bd0(a0 * 0.5, (a0 + a1 * a1) * 0.5) + a1 * a1 * 0.5;
      }
      return t - u - (SMC_LOG_SQRT_2PI + l_x2n);
    }
    case SMC_DIST_EXPONENTIAL:  // (scale, x)  dexp
      if (!(a0 > 0.0)) return qnan;
      if (a1 < 0.0) return ninf;
      return -a1 / a0 - log(a0);
  }
  return qnan;
}

// d logpdf / d argument k, per element (src/diff.jl:94-174; the ::Real variants are these summed)
__device__ __forceinline__ double smc_dlogpdf(int dist, int k, double a0, double a1, double a2) {
  switch (dist) {
    case SMC_DIST_NORMAL: {  // (mu, sigma, x)
      double d = a2 - a0, s2 = a1 * a1;
      if (k == 0) return d / s2;
      if (k == 1) return (d * d / s2 - 1.0) / a1;
      return (a0 - a2) / s2;
    }
    case SMC_DIST_UNIFORM: {  // (a, b, x)
      double in = (a0 <= a2 && a2 <= a1) ? 1.0 : 0.0;
      if (k == 0) return in / (a1 - a0);
      if (k == 1) return in / (a0 - a1);
      return 0.0;
    }
    case SMC_DIST_WEIBULL: {  // (sh, sc, x)
      double r = a2 / a1, rs = pow(r, a0);
      if (k == 0) return (1.0 - rs) * log(r) + 1.0 / a0;
      if (k == 1) return (rs - 1.0) * a0 / a1;
      return ((1.0 - rs) * a0 - 1.0) / a2;
    }
    case SMC_DIST_BETA:  // (a, b, x)
      if (k == 2) return (a0 - 1.0) / a2 - (a1 - 1.0) / (1.0 - a2);
      if (k == 0) return smc_digamma(a0 + a1) - smc_digamma(a0) + log(a2);
      return smc_digamma(a0 + a1) - smc_digamma(a1) + log(1.0 - a2);
    case SMC_DIST_TDIST: {  // (df, x)
      if (k == 1) return -(a0 + 1.0) * a1 / (a0 + a1 * a1);
      double t = a1 * a1 + a0;
      return ((a1 * a1 - 1.0) / t + log(a0 / t) + smc_digamma((a0 + 1.0) * 0.5) - smc_digamma(a0 * 0.5)) * 0.5;
    }
    case SMC_DIST_EXPONENTIAL:  // (sc, x)
      if (k == 1) return -1.0 / a0;
      return (a1 - a0) / (a0 * a0);
    case SMC_DIST_GAMMA:  // (sh, sc, x)
      if (k == 2) return -(a1 + a2 - a0 * a1) / (a1 * a2);
      if (k == 0) return log(a2) - log(a1) - smc_digamma(a0);
      return (a2 - a1 * a0) / (a1 * a1);
    case SMC_DIST_CAUCHY: {  // (mu, sc, x)
      double d = a2 - a0, den = a1 * a1 + d * d;
      if (k == 2) return 2.0 * (a0 - a2) / den;
      if (k == 0) return 2.0 * d / den;
      return (d * d - a1 * a1) / (a1 * den);
    }
    case SMC_DIST_LOGNORMAL: {  // (lmu, lsc, x)
      double t = a1 * a1, lx = log(a2);
      if (k == 2) return (a0 - t - lx) / (t * a2);
      if (k == 0) return (lx - a0) / t;
      return (a0 * a0 - t - lx * (2.0 * a0 - lx)) / (a1 * t);
    }
    case SMC_DIST_BERNOULLI:  // (p, x)
      return 1.0 / (a0 - (1.0 - a1));
    case SMC_DIST_BINOMIAL:  // (n, p, x)
      return a2 / a1 - (a0 - a2) / (1.0 - a1);
    case SMC_DIST_POISSON:  // (lambda, x)
      return a1 / a0 - 1.0;
    case SMC_DIST_TESTDIFF:
      return 1.0;
  }
  return nan("");
}
