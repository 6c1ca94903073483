// Device log-densities and their adjoints (FP64).
// Formulas: /root/reference/src/distribs.jl:12-33 (Bernoulli, Normal, Uniform, defined locally there) and the
// nine libRmath densities the reference reaches by ccall (src/distribs.jl:40-48,65-67): closed-form
// log-densities in the scale parametrisation with R's edge cases (outside support -> -inf, invalid
// parameter -> NaN).  Adjoints: the @dfunc table, src/diff.jl:94-174.
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
      if (!(a1 >= 0.0 && a1 <= 1.0) || a0 < 0.0 || a0 != rint(a0)) return qnan;
      if (a2 < 0.0 || a2 > a0 || a2 != rint(a2)) return ninf;
      double lc = lgamma(a0 + 1.0) - lgamma(a2 + 1.0) - lgamma(a0 - a2 + 1.0);
      return lc + smc_xlogy(a2, a1) + smc_xlog1py(a0 - a2, -a1);
    }
    case SMC_DIST_GAMMA: {  // (shape, scale, x)  dgamma
      if (a0 < 0.0 || !(a1 > 0.0)) return qnan;
      if (a2 < 0.0) return ninf;
      if (a0 == 0.0) return a2 == 0.0 ? INFINITY : ninf;
      if (a2 == 0.0) return a0 < 1.0 ? INFINITY : (a0 > 1.0 ? ninf : -log(a1));
      return (a0 - 1.0) * log(a2) - a2 / a1 - lgamma(a0) - a0 * log(a1);
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
      if (!(a0 > 0.0) || !(a1 > 0.0)) return qnan;
      if (a2 < 0.0 || a2 > 1.0) return ninf;
      if (a2 == 0.0) return a0 > 1.0 ? ninf : (a0 < 1.0 ? INFINITY : log(a1));
      if (a2 == 1.0) return a1 > 1.0 ? ninf : (a1 < 1.0 ? INFINITY : log(a0));
      return lgamma(a0 + a1) - lgamma(a0) - lgamma(a1) + smc_xlogy(a0 - 1.0, a2) + smc_xlog1py(a1 - 1.0, -a2);
    }
    case SMC_DIST_POISSON:  // (lambda, x)  dpois
      if (a0 < 0.0) return qnan;
      if (a1 < 0.0 || a1 != rint(a1)) return ninf;
      return smc_xlogy(a1, a0) - a0 - lgamma(a1 + 1.0);
    case SMC_DIST_TDIST:  // (df, x)  dt
      if (!(a0 > 0.0)) return qnan;
      return lgamma((a0 + 1.0) * 0.5) - lgamma(a0 * 0.5) - 0.5 * log(a0 * SMC_PI) - (a0 + 1.0) * 0.5 * log1p(a1 * a1 / a0);
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
