// Shared by the statement executor (smc_core.cu) and the register-program VM (vm_exec.cu): operand descriptors of tape
// registers and the elementwise operation table.
#pragma once
#include "smc_internal.h"
#include "dists.cuh"

struct Opd {
  const double* p;
  long long se;  // element stride
  int sc;        // chain stride (0 = shared by all chains)
  int is_imm;
  double imm;
};
struct Dst {
  double* p;
  long long se;
  int sc;
};

Opd smc_make_opd(const smc_model_s* m, uint32_t r);
Dst smc_make_dst(const smc_model_s* m, uint32_t r);
int smc_scratch_reserve(smc_model_s* m, int64_t doubles);

__device__ __forceinline__ double opd_ld(const Opd& o, long long e, long long c) {
  return o.is_imm ? o.imm : __ldg(o.p + e * o.se + c * o.sc);
}

__device__ __forceinline__ double smc_apply(int op, int dist, int k, double a, double b, double c) {
  switch (op) {
    case SMC_OP_COPY: return a;
    case SMC_OP_NEG: return -a;
    case SMC_OP_EXP: return exp(a);
    case SMC_OP_LOG: return log(a);
    case SMC_OP_SIN: return sin(a);
    case SMC_OP_COS: return cos(a);
    case SMC_OP_ABS: return fabs(a);
    case SMC_OP_SIGN: return a > 0.0 ? 1.0 : (a < 0.0 ? -1.0 : a);
    case SMC_OP_DIGAMMA: return smc_digamma(a);
    case SMC_OP_ZERO: return 0.0;
    case SMC_OP_SQRT: return sqrt(a);
    case SMC_OP_ADD: return a + b;
    case SMC_OP_SUB: return a - b;
    case SMC_OP_MUL: return a * b;
    case SMC_OP_DIV: return a / b;
    case SMC_OP_POW: return pow(a, b);
    case SMC_OP_MAX: return fmax(a, b);
    case SMC_OP_MIN: return fmin(a, b);
    case SMC_OP_GT: return a > b ? 1.0 : 0.0;
    case SMC_OP_LT: return a < b ? 1.0 : 0.0;
    case SMC_OP_GE: return a >= b ? 1.0 : 0.0;
    case SMC_OP_LE: return a <= b ? 1.0 : 0.0;
    case SMC_OP_EQ: return a == b ? 1.0 : 0.0;
    case SMC_OP_LOGPDF: return smc_logpdf(dist, a, b, c);
    case SMC_OP_DLOGPDF: return smc_dlogpdf(dist, k, a, b, c);
  }
  return nan("");
}
