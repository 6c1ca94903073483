/* Device op tape of libsimplemcmc_cuda.so -- binary layout (little endian) and opcode numbers.
 *
 * The tape is what the host front end lowers the reference's derivation output to: the statement
 * list `m.exprs ++ m.dexprs` that generateModelFunction (/root/reference/src/parsing.jl:409-467)
 * would have eval'ed as Julia code, with the register shapes that preCalculate's `vhint`
 * (/root/reference/src/parsing.jl:354-384) used to supply.  One statement = one 3-address operation
 * (/root/reference/src/parsing.jl:123-183) or one fused distribution adjoint
 * (/root/reference/src/diff.jl:94-174).
 *
 * blob := smc_tape_header, smc_tape_reg[n_regs], smc_tape_stmt[n_stmts], smc_tape_data[n_data],
 *         index tables: { uint32 n; uint32 pad; int32 idx[n rounded up to even] } * n_idx,
 *         programs (header.reserved[1] of them): see smc_prog_header
 */
#ifndef SIMPLEMCMC_TAPE_H
#define SIMPLEMCMC_TAPE_H
#include <stdint.h>

#define SMC_TAPE_MAGIC 0x54434D53u /* "SMCT" */
#define SMC_TAPE_VERSION 4u

/* register kinds */
enum {
  SMC_REG_IMM = 0,   /* scalar literal, value in .imm */
  SMC_REG_DATA = 1,  /* external array bound by name (the globals of Main, parsing.jl:484-485); aux = data slot */
  SMC_REG_HDR = 2,   /* computed, parameter independent: evaluated once (the let-header, parsing.jl:432-442) */
  SMC_REG_CHAIN = 3, /* computed per chain, stored [numel][C] */
  SMC_REG_PARAM = 4, /* view of the parameter vector __beta, aux = first row (betaAssign, parsing.jl:338-349) */
  SMC_REG_GRAD = 5   /* view of the returned gradient vector, aux = first row */
};

/* statement flags */
#define SMC_F_ACC 1u     /* dst += result (the `dv = dv + ...` of diff.jl:214) instead of dst = result */
#define SMC_F_TRANS_A 2u /* MATMUL: use transpose(A) */
#define SMC_F_TRANS_B 4u /* MATMUL: use transpose(B) */
#define SMC_F_FWD_ONLY 8u /* forward statement that runs only when the gradient is NOT asked for: a later statement of the
                             gradient part recomputes its values and takes its reductions over (so that an evaluation with
                             gradient makes one pass over the index space instead of two) */

/* opcodes.  Elementwise ops broadcast scalar operands; when dst is a scalar and an operand has n>1
 * elements the statement sums its n results (the `sum(...)` of the ::Real rules, diff.jl:29-178). */
enum {
  SMC_OP_COPY = 0, SMC_OP_NEG = 1, SMC_OP_EXP = 2, SMC_OP_LOG = 3, SMC_OP_SIN = 4, SMC_OP_COS = 5,
  SMC_OP_ABS = 6, SMC_OP_SIGN = 7, SMC_OP_DIGAMMA = 8, SMC_OP_ZERO = 9, SMC_OP_SQRT = 10,
  SMC_OP_ADD = 16, SMC_OP_SUB = 17, SMC_OP_MUL = 18, SMC_OP_DIV = 19, SMC_OP_POW = 20, SMC_OP_MAX = 21,
  SMC_OP_MIN = 22, SMC_OP_GT = 23, SMC_OP_LT = 24, SMC_OP_GE = 25, SMC_OP_LE = 26, SMC_OP_EQ = 27,
  SMC_OP_LOGPDF = 32,    /* aux0 = distribution id; sources = density arguments, variate LAST (parsing.jl:107-108) */
  SMC_OP_DLOGPDF = 33,   /* aux0 = distribution id, aux1 = argument differentiated; elementwise d logpdf / d arg */
  SMC_OP_MATMUL = 40,    /* dst[r x n] = A[r x k] * B[k x n], column major; aux0 = r, aux1 = k, imm = n (as double) */
  SMC_OP_TRANSPOSE = 41, /* aux0 = rows of src, aux1 = cols of src */
  SMC_OP_GATHER = 42,    /* dst[e] = src[idx[e]], aux0 = index table */
  SMC_OP_SCATTER_SET = 43, /* dst[idx[e]] = src[e]  -- the reference's ref adjoint (parsing.jl:277-281) */
  SMC_OP_SCATTER_ADD = 44, /* dst[idx[e]] += src[e] -- the corrected adjoint (TODO.md:11) */
  SMC_OP_GLM = 48,       /* fused  eta = X*beta ; L = sum_i l(eta_i, y_i) ; G = X' * dl/deta  (see smc_tape_stmt) */
  SMC_OP_FUSED = 49,     /* a run of elementwise / reducing statements over one index space, executed as one register
                            program (smc_prog_instr) without materialising its temporaries; aux0 = program, aux1 = n */
  SMC_OP_AFFNORM = 50    /* affine residual into a Normal density (the pattern of examples/ornstein.jl:24-26):
                              dst (+)= sum_e logpdfNormal(mu, sigma, r_e),  r_e = ((+-t_0) +- t_1) +- ...,  t_j = D_j[e] * s_j | s_j
                            src = {mu, sigma} (scalars); aux1 = n; aux0 = index table holding the descriptor
                              { T, reg(dL/dmu) | -1, reg(dL/dsigma) | -1,
                                T x { kind (0: array * scalar, 1: scalar), sign (+1 | -1), array register | -1, element offset
                                      into that array (a contiguous view), scalar register, reg(dL/ds_j) | -1 } }
                            the adjoint registers are per-chain scalars written by the same launch */
};

/* program-only opcodes (besides the elementwise opcodes above, which then act on program-local registers) */
enum {
  SMC_P_LOADV = 100,  /* local[dst] = tape register `reg` at the current element (and chain) */
  SMC_P_LOADS = 101,  /* local[dst] = scalar tape register `reg` (prologue: once per thread) */
  SMC_P_STOREV = 102, /* tape register `reg` at the current element = local[src0] */
  SMC_P_REDUCE = 103  /* reduction slot aux0 += local[src0] */
};
#define SMC_PROG_MAX_LOCALS 48
#define SMC_PROG_MAX_REDUCE 8

/* distribution ids (src/distribs.jl:40-51) */
enum {
  SMC_DIST_NORMAL = 0, SMC_DIST_UNIFORM = 1, SMC_DIST_WEIBULL = 2, SMC_DIST_BETA = 3, SMC_DIST_TDIST = 4,
  SMC_DIST_EXPONENTIAL = 5, SMC_DIST_GAMMA = 6, SMC_DIST_CAUCHY = 7, SMC_DIST_LOGNORMAL = 8,
  SMC_DIST_BERNOULLI = 9, SMC_DIST_BINOMIAL = 10, SMC_DIST_POISSON = 11, SMC_DIST_TESTDIFF = 12
};

/* GLM families for SMC_OP_GLM (aux0) */
enum {
  SMC_GLM_NORMAL = 0,    /* l = logpdfNormal(0, sigma, a*eta + b_i);      imm = sigma, aux1 = sign bits of a, b
                            (+ SMC_GLM_F_GRAM)                                                                     */
  SMC_GLM_BERNOULLI = 1, /* l = logpdfBernoulli(1/(1+exp(s*eta)), y_i);   aux1 = 1 when s = -1                  */
  SMC_GLM_BINOMIAL = 2,  /* l = logpdfBinomial(n, 1/(1+exp(s*eta)), y_i); imm = n, aux1 = 1 when s = -1                */
  SMC_GLM_POISSON = 3    /* l = logpdfPoisson(exp(s*eta), y_i)   (reserved: not emitted by the host yet)               */
};

/* aux1 flag of a NORMAL GLM statement: evaluate from the sufficient statistics X'X, X'y, y'y (computed once when the model
 * is finalized) instead of streaming X -- an opt-in algebraic rewrite in the spirit of the reference's hoisting of
 * parameter-independent terms (src/parsing.jl:432-442, TODO.md:4-9); it changes the arithmetic, so hosts must ask for it */
#define SMC_GLM_F_GRAM 4u

typedef struct {
  uint32_t magic, version;
  uint32_t n_regs, n_stmts;
  uint32_t n_params; /* length of __beta */
  uint32_t n_data, n_idx;
  uint32_t result_reg; /* scalar register holding the final accumulator */
  uint32_t flags;
  uint32_t reserved[3]; /* reserved[0] = number of forward statements (the rest only serve the gradient), reserved[1] = number of programs */
} smc_tape_header;

typedef struct {
  uint32_t kind, rows, cols, ndim;
  int64_t aux;
  double imm;
} smc_tape_reg;

typedef struct {
  uint32_t op, flags, dst, nsrc;
  uint32_t src[4]; /* GLM: src = {X, beta, y};  dst = L (scalar), aux2 = register of G */
  uint32_t aux0, aux1, aux2, pad;
  double imm;
} smc_tape_stmt;

typedef struct {
  char name[56];
  uint32_t rows, cols;
} smc_tape_data;

/* program := smc_prog_header, smc_prog_instr[n_instr] (the first n_prologue run once per thread), smc_prog_red[n_red] */
typedef struct {
  uint32_t n_instr, n_prologue, n_red, n;
} smc_prog_header;

typedef struct {
  uint16_t op;
  uint8_t dst, nsrc;
  uint8_t src[3];
  uint8_t flags; /* SMC_F_ACC: local[dst] += value */
  uint16_t aux0, aux1;
  uint32_t reg;
} smc_prog_instr;

typedef struct {
  uint32_t reg, acc; /* destination scalar register of a reduction slot; acc: dst += total */
} smc_prog_red;

#endif
