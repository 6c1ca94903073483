// Register-program VM: executes one SMC_OP_FUSED statement -- a run of elementwise / reducing tape statements over one
// index space (include/simplemcmc_tape.h, smc_prog_instr) -- in a single launch for all chains in lockstep.
//
// The reference's generated function materialises one temporary array per statement
// (/root/reference/src/parsing.jl:409-467); here every intermediate stays on chip:
//   * plan time (smc_vm_plan): the instruction list becomes micro-ops whose operands are addressed directly -- a
//     program-local vector slot, a per-chain scalar staged once per block, or a tape array in HBM -- so LOADV / LOADS /
//     STOREV / REDUCE instructions disappear into operand modes and side effects of the computing instruction; values
//     that are only handed from one micro-op to the next stay in registers (chain micro-ops, vm_form_chains);
//   * run time (k_vm): thread = (chain lane, element lane); each thread owns V elements per pass, so one decode of
//     a micro-op is amortised over V elements.  Locals live in shared memory as [slot][v][thread] (conflict free),
//     chains are the fastest thread index (coalesced per-chain operands, broadcast data operands).
// Reductions (`sum`, `dot`, vectorised logpdf: src/distribs.jl:97-105) accumulate per thread in element order, then
// over the element lanes by a fixed tree, then over element chunks in k_vm_final: bit-reproducible for a given chain
// count.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include "exec_common.cuh"

namespace {

constexpr int NT = 256;  // threads per block

// device forms (built by smc_vm_prepare from the plan)
struct VmDevUop {   // 16 bytes
  uint8_t op, dist, darg, flags;
  uint8_t o[4];     // operand-table entries of the sources 0..2 and of the accumulate operand
  uint8_t dst, store, red, code3;   // operand-table entries of the local / array destinations, reduction slot; chain stage 3
  uint32_t pad1;
};
struct VmDevOpd {   // 32 bytes: one distinct operand of the program
  int kind, idx;    // VM_LOCAL: slot; VM_SCAL: scalar slot; VM_ARRAY: p / se / sc
  const double* p;
  long long se, sc;
};
struct VmOt {       // operand table entry resolved for this block (kept in shared memory), 32 bytes
  const double* gbase;  // tape array in HBM: element v of a thread's pass = gbase + (flags&2 ? cx : 0) + eb*Cc + v*S; else NULL
  long long Cc, S;
  int off;              // shared-memory operand: index (doubles) = off + (flags&1 ? tid : 0) + (flags&2 ? cx : 0) + v*(flags&1 ? NT : 0)
  int flags;
};

struct VmDec {       // a micro-op pre-decoded for this block (shared memory), 32 bytes: all the shared-memory fast path reads
  uint32_t w0;       // op | dist << 8 | darg << 16 | flags << 24
  uint32_t w1;       // reduction slot | gen << 8 (bit k: operand k is a tape array in HBM) | sel << 16 (2 bits per operand:
                     //   1 = "+ tid, stride NT" (local), 2 = "+ chain lane" (per-chain scalar)) | third chain stage << 24
  int b0, b1, b2, bd;   // shared-memory indices of operands 0..2 and of the local destination, before the per-thread term
  uint32_t ot;       // operand-table entries o[0] | o[1] << 8 | o[2] << 16 | o[3] << 24 (generic path)
  uint32_t st;       // operand-table entry of the array destination
};

struct VmArgs {
  const VmDevUop* uops;
  const VmDevOpd* ot;
  const VmDesc* scal;
  int n_uops, n_ot, n_scal, n_loc, n_red, tcc, S, direct, V;
  long long n, C, chunk;
  int* bad;
  double* partial;  // [n_red][S][C]
  Dst rd[SMC_PROG_MAX_REDUCE];   // destinations of the reduction slots (direct mode: S == 1)
  int racc[SMC_PROG_MAX_REDUCE];
};

// the heavy operations (densities, their adjoints, pow, trigonometry, ...) are called, not inlined
__device__ __noinline__ double vm_apply_slow(int op, int dist, int k, double a, double b, double c) {
  return smc_apply(op, dist, k, a, b, c);
}

// One pass of a micro-op over the vc elements a thread owns (element v at pointer + v * stride).  The operands of four
// elements are read before anything is written, so a destination slot may coincide with a source slot and the loads of a
// batch are independent; the side effects (accumulate operand, local write, array store, reduction) are uniform
// per-micro-op flags.
// EXPR sees p, q (and w when W3) = operands 0, 1, 2 of the element.  The macros are expanded twice: once where every
// operand lives in shared memory (32-bit indices into the block's shared array: LDS / STS with cheap addressing),
// once with generic 64-bit pointers (a tape array in HBM among the operands).
#define VM_LOAD4(W3, EXPR)                                                                                       \
  const double p_0 = p0[(v + 0) * s0], p_1 = p0[(v + 1) * s0], p_2 = p0[(v + 2) * s0], p_3 = p0[(v + 3) * s0];   \
  const double q_0 = p1[(v + 0) * s1], q_1 = p1[(v + 1) * s1], q_2 = p1[(v + 2) * s1], q_3 = p1[(v + 3) * s1];   \
  const double w_0 = W3 ? p2[(v + 0) * s2] : 0.0, w_1 = W3 ? p2[(v + 1) * s2] : 0.0,                             \
               w_2 = W3 ? p2[(v + 2) * s2] : 0.0, w_3 = W3 ? p2[(v + 3) * s2] : 0.0;                             \
  double r_0, r_1, r_2, r_3;                                                                                     \
  { const double p = p_0, q = q_0, w = w_0; (void)p; (void)q; (void)w; r_0 = (EXPR); }                           \
  { const double p = p_1, q = q_1, w = w_1; (void)p; (void)q; (void)w; r_1 = (EXPR); }                           \
  { const double p = p_2, q = q_2, w = w_2; (void)p; (void)q; (void)w; r_2 = (EXPR); }                           \
  { const double p = p_3, q = q_3, w = w_3; (void)p; (void)q; (void)w; r_3 = (EXPR); }
#define VM_LOAD1(W3)                                                             \
  const double p = p0[v * s0], q = p1[v * s1], w = W3 ? p2[v * s2] : 0.0;        \
  (void)p; (void)q; (void)w;
#define VM_RUN(W3, EXPR)                                                                                \
  {                                                                                                     \
    int v = 0;                                                                                          \
    double rsum = 0.0;                                                                                  \
    for (; v + 4 <= vc; v += 4) {                                                                       \
      VM_LOAD4(W3, EXPR)                                                                                \
      if (fl & VMF_ACC) {                                                                               \
        const double a_0 = pa[(v + 0) * sa], a_1 = pa[(v + 1) * sa], a_2 = pa[(v + 2) * sa], a_3 = pa[(v + 3) * sa]; \
        r_0 = a_0 + r_0; r_1 = a_1 + r_1; r_2 = a_2 + r_2; r_3 = a_3 + r_3;                             \
      }                                                                                                 \
      if (fl & VMF_WRITE) { pd[(v + 0) * NT] = r_0; pd[(v + 1) * NT] = r_1; pd[(v + 2) * NT] = r_2; pd[(v + 3) * NT] = r_3; } \
      if (fl & VMF_STORE) { ps[(v + 0) * ss] = r_0; ps[(v + 1) * ss] = r_1; ps[(v + 2) * ss] = r_2; ps[(v + 3) * ss] = r_3; } \
      if (fl & VMF_REDUCE) { rsum += r_0; rsum += r_1; rsum += r_2; rsum += r_3; }                      \
    }                                                                                                   \
    for (; v < vc; v++) {                                                                               \
      VM_LOAD1(W3)                                                                                      \
      double r = (EXPR);                                                                                \
      if (fl & VMF_ACC) r = pa[v * sa] + r;                                                             \
      if (fl & VMF_WRITE) pd[v * NT] = r;                                                               \
      if (fl & VMF_STORE) ps[v * ss] = r;                                                               \
      rsum += r;                                                                                        \
    }                                                                                                   \
    if (fl & VMF_REDUCE) Rd[uo.red * NT + tid] += rsum;                                                 \
  }

// the operation table; p0/p1/p2 (sources), pd (local destination), pa (accumulate operand), ps (array destination) and
// their strides are in scope
#define VM_SWITCH                                                                                                   \
  switch (uo.op) {                                                                                                  \
    case SMC_OP_COPY: VM_RUN(0, p) break;                                                                           \
    case SMC_OP_NEG: VM_RUN(0, -p) break;                                                                           \
    case SMC_OP_ADD: VM_RUN(0, p + q) break;                                                                        \
    case SMC_OP_SUB: VM_RUN(0, p - q) break;                                                                        \
    case SMC_OP_MUL: VM_RUN(0, p * q) break;                                                                        \
    case SMC_OP_DIV: VM_RUN(0, p / q) break;                                                                        \
    case SMC_OP_EXP: VM_RUN(0, exp(p)) break;                                                                       \
    case SMC_OP_LOG: VM_RUN(0, log(p)) break;                                                                       \
    case SMC_OP_ZERO: VM_RUN(0, 0.0) break;                                                                         \
    default: {                                                                                                      \
      /* Normal density and its adjoints with one sigma for the whole pass (a per-chain scalar, the usual case): */ \
      /* 1/sigma, 1/sigma^2 and log(sigma) are formed once per pass, the element loop is multiply-add only       */ \
      /* (src/distribs.jl:23-28, src/diff.jl:94-99)                                                              */ \
      if (uo.dist == SMC_DIST_NORMAL && s1 == 0 && (uo.op == SMC_OP_LOGPDF || uo.op == SMC_OP_DLOGPDF) && p1[0] > 0.0) { \
        const double sg = p1[0], inv = 1.0 / sg, inv2 = 1.0 / (sg * sg);                                            \
        const double lconst = -log(sg) - SMC_LOG_SQRT_2PI;                                                          \
        if (uo.op == SMC_OP_LOGPDF) VM_RUN(1, fma(-0.5 * ((w - p) * inv), (w - p) * inv, lconst))                   \
        else if (uo.darg == 0) VM_RUN(1, (w - p) * inv2)                                                            \
        else if (uo.darg == 1) VM_RUN(1, ((w - p) * (w - p) * inv2 - 1.0) * inv)                                    \
        else VM_RUN(1, (p - w) * inv2)                                                                              \
        break;                                                                                                      \
      }                                                                                                             \
      double rsum = 0.0;                                                                                            \
      bool gaveup = false;                                                                                          \
      for (int v = 0; v < vc; v++) {                                                                                \
        double r = vm_apply_slow(uo.op, uo.dist, uo.darg, p0[v * s0], p1[v * s1], p2[v * s2]);                      \
        if (uo.op == SMC_OP_LOGPDF && (r == -INFINITY || isnan(r))) { /* "give up eval", src/distribs.jl:69-70 */   \
          gaveup = true;                                                                                            \
          r = 0.0;                                                                                                  \
        }                                                                                                           \
        if (fl & VMF_ACC) r = pa[v * sa] + r;                                                                       \
        if (fl & VMF_WRITE) pd[v * NT] = r;                                                                         \
        if (fl & VMF_STORE) ps[v * ss] = r;                                                                         \
        rsum += r;                                                                                                  \
      }                                                                                                             \
      if (gaveup) a.bad[c] = 1;                                                                                     \
      if (fl & VMF_REDUCE) Rd[uo.red * NT + tid] += rsum;                                                           \
    }                                                                                                               \
  }


// Chain micro-ops (VM_OP_CHAIN, see smc_internal.h): r starts as operand 0 and goes through up to three stages in
// registers; an operand that is the same for every element of the pass (a per-chain scalar) is read once.  Every stage is
// one rounded operation, exactly what the separate micro-ops computed (no contraction across stages).
#define VM_CH_STAGE(CODE, O)                                                                                   \
  switch (CODE) {                                                                                              \
    case VM_CH_ADD: r_0 = __dadd_rn(r_0, O##_0); r_1 = __dadd_rn(r_1, O##_1); r_2 = __dadd_rn(r_2, O##_2); r_3 = __dadd_rn(r_3, O##_3); break; \
    case VM_CH_SUB: r_0 = __dsub_rn(r_0, O##_0); r_1 = __dsub_rn(r_1, O##_1); r_2 = __dsub_rn(r_2, O##_2); r_3 = __dsub_rn(r_3, O##_3); break; \
    case VM_CH_RSUB: r_0 = __dsub_rn(O##_0, r_0); r_1 = __dsub_rn(O##_1, r_1); r_2 = __dsub_rn(O##_2, r_2); r_3 = __dsub_rn(O##_3, r_3); break; \
    case VM_CH_MUL: r_0 = __dmul_rn(r_0, O##_0); r_1 = __dmul_rn(r_1, O##_1); r_2 = __dmul_rn(r_2, O##_2); r_3 = __dmul_rn(r_3, O##_3); break; \
    case VM_CH_DIV: r_0 = r_0 / O##_0; r_1 = r_1 / O##_1; r_2 = r_2 / O##_2; r_3 = r_3 / O##_3; break;        \
    case VM_CH_RDIV: r_0 = O##_0 / r_0; r_1 = O##_1 / r_1; r_2 = O##_2 / r_2; r_3 = O##_3 / r_3; break;       \
    case VM_CH_NEG: r_0 = -r_0; r_1 = -r_1; r_2 = -r_2; r_3 = -r_3; break;                                    \
    case VM_CH_NLOG:                                                                                           \
      r_0 = fma(-0.5 * ((r_0 - O##_0) * inv), (r_0 - O##_0) * inv, lconst);                                    \
      r_1 = fma(-0.5 * ((r_1 - O##_1) * inv), (r_1 - O##_1) * inv, lconst);                                    \
      r_2 = fma(-0.5 * ((r_2 - O##_2) * inv), (r_2 - O##_2) * inv, lconst);                                    \
      r_3 = fma(-0.5 * ((r_3 - O##_3) * inv), (r_3 - O##_3) * inv, lconst);                                    \
      break;                                                                                                   \
    case VM_CH_ND_RMO: r_0 = (r_0 - O##_0) * inv2; r_1 = (r_1 - O##_1) * inv2; r_2 = (r_2 - O##_2) * inv2; r_3 = (r_3 - O##_3) * inv2; break; \
    case VM_CH_ND_OMR: r_0 = (O##_0 - r_0) * inv2; r_1 = (O##_1 - r_1) * inv2; r_2 = (O##_2 - r_2) * inv2; r_3 = (O##_3 - r_3) * inv2; break; \
    case VM_CH_ND_SIG:                                                                                         \
      r_0 = ((r_0 - O##_0) * (r_0 - O##_0) * inv2 - 1.0) * inv;                                                \
      r_1 = ((r_1 - O##_1) * (r_1 - O##_1) * inv2 - 1.0) * inv;                                                \
      r_2 = ((r_2 - O##_2) * (r_2 - O##_2) * inv2 - 1.0) * inv;                                                \
      r_3 = ((r_3 - O##_3) * (r_3 - O##_3) * inv2 - 1.0) * inv;                                                \
      break;                                                                                                   \
    default: break;                                                                                            \
  }
#define VM_CH_OPND(O, P, S, K)                                                                                 \
  const double O##_0 = (S) ? (P)[(v + 0) * (S)] : (K), O##_1 = (S) ? (P)[(v + 1) * (S)] : (K),                 \
               O##_2 = (S) ? (P)[(v + 2) * (S)] : (K), O##_3 = (S) ? (P)[(v + 3) * (S)] : (K);

__global__ void __launch_bounds__(NT) k_vm(VmArgs a) {
  extern __shared__ __align__(16) double vsmd[];
  unsigned char* vsm = (unsigned char*)vsmd;
  VmDevUop* U = (VmDevUop*)vsm;                                   // [n_uops]
  VmDec* DEC = (VmDec*)(vsm + (size_t)((a.n_uops + 1) & ~1) * 16);   // [n_uops]
  VmOt* OT = (VmOt*)(DEC + a.n_uops);                             // [n_ot]
  const int sc_off = (((a.n_uops + 1) & ~1) * 16 + a.n_uops * 32 + a.n_ot * 32) / 8;   // (indices in doubles from vsmd)
  const int lc_off = sc_off + max(1, a.n_scal) * a.tcc;
  double* Sc = vsmd + sc_off;                                     // [max(1, n_scal)][tcc]
  double* Lc = vsmd + lc_off;                                     // [n_loc][V][NT]
  double* Rd = Lc + (size_t)a.n_loc * a.V * NT;                   // [n_red][NT]
  const int tid = threadIdx.x, V = a.V;
  const int tcc = a.tcc, eyn = NT / tcc;
  const int cx = tid % tcc, ey = tid / tcc;
  const long long cbase = (long long)blockIdx.x * tcc;
  const long long c = cbase + cx;
  const bool active = c < a.C;
  for (int i = tid; i < a.n_uops * 4; i += NT) ((uint32_t*)U)[i] = ((const uint32_t*)a.uops)[i];
  for (int i = tid; i < a.n_ot; i += NT) {
    const VmDevOpd d = a.ot[i];
    VmOt t;
    t.gbase = nullptr; t.Cc = 0; t.S = 0; t.off = sc_off; t.flags = 0;   // "no operand": a valid dummy
    if (d.kind == VM_LOCAL) { t.off = lc_off + d.idx * V * NT; t.flags = 1; }
    else if (d.kind == VM_SCAL) { t.off = sc_off + d.idx * tcc; t.flags = 2; }
    else if (d.kind == VM_ARRAY) { t.gbase = d.p + cbase * d.sc; t.flags = d.sc ? 2 : 0; t.Cc = d.se; t.S = (long long)eyn * d.se; }
    OT[i] = t;
  }
  if (tid == 0 && a.n_scal == 0) Sc[0] = 0.0;
  for (int i = tid; i < a.n_scal * tcc; i += NT) {
    const int k = i / tcc, cl = i - k * tcc;
    const VmDesc d = a.scal[k];
    const long long cc = cbase + cl;
    Sc[i] = d.p ? (cc < a.C ? __ldg(d.p + cc * d.sc) : 0.0) : d.imm;
  }
  for (int k = 0; k < a.n_red; k++) Rd[k * NT + tid] = 0.0;
  __syncthreads();
  for (int u = tid; u < a.n_uops; u += NT) {
    const VmDevUop uo = U[u];
    const VmOt t0 = OT[uo.o[0]], t1 = OT[uo.o[1]], t2 = OT[uo.o[2]];
    VmDec d;
    d.w0 = (uint32_t)uo.op | ((uint32_t)uo.dist << 8) | ((uint32_t)uo.darg << 16) | ((uint32_t)uo.flags << 24);
    const uint32_t gen = (t0.gbase ? 1u : 0u) | (t1.gbase ? 2u : 0u) | (t2.gbase ? 4u : 0u);
    const uint32_t sel = (uint32_t)(t0.flags & 3) | ((uint32_t)(t1.flags & 3) << 2) | ((uint32_t)(t2.flags & 3) << 4);
    d.w1 = (uint32_t)uo.red | (gen << 8) | (sel << 16) | ((uint32_t)uo.code3 << 24);
    d.b0 = t0.off; d.b1 = t1.off; d.b2 = t2.off; d.bd = OT[uo.dst].off;
    d.ot = (uint32_t)uo.o[0] | ((uint32_t)uo.o[1] << 8) | ((uint32_t)uo.o[2] << 16) | ((uint32_t)uo.o[3] << 24);
    d.st = uo.store;
    DEC[u] = d;
  }
  __syncthreads();

  if (active) {
    const long long e0 = (long long)blockIdx.y * a.chunk, e1 = min(a.n, e0 + a.chunk);
    for (long long eb = e0 + ey; eb < e1; eb += (long long)eyn * V) {
      const int vc = (int)min((long long)V, (e1 - eb + eyn - 1) / eyn);   // elements eb + v*eyn, v < vc
      for (int u = 0; u < a.n_uops; u++) {
        const uint4 wa = ((const uint4*)DEC)[2 * u], wb = ((const uint4*)DEC)[2 * u + 1];
        struct { int op, dist, darg, red; } uo;
        uo.op = wa.x & 0xff; uo.dist = (wa.x >> 8) & 0xff; uo.darg = (wa.x >> 16) & 0xff; uo.red = wa.y & 0xff;
        const int fl = wa.x >> 24;
        const int gen = (wa.y >> 8) & 0xff, sel = (wa.y >> 16) & 0xff;
        // shared-memory operands: index of element 0 and stride (0 for a per-chain scalar, NT for a local)
        const int i0 = (int)wa.z + ((sel & 1) ? tid : 0) + ((sel & 2) ? cx : 0);
        const int i1 = (int)wa.w + ((sel & 4) ? tid : 0) + ((sel & 8) ? cx : 0);
        const int i2 = (int)wb.x + ((sel & 16) ? tid : 0) + ((sel & 32) ? cx : 0);
        double* const pd = vsmd + (int)wb.y + tid;
        if (uo.op == VM_OP_CHAIN) {
          const VmOt t0 = OT[wb.z & 0xff], t1 = OT[(wb.z >> 8) & 0xff], t2 = OT[(wb.z >> 16) & 0xff], t3 = OT[(wb.z >> 24) & 0xff];
          const VmOt ts = OT[wb.w & 0xff];
#define VM_OT_PTR(t) ((t).gbase ? (t).gbase + (((t).flags & 2) ? cx : 0) + eb * (t).Cc                         \
                               : vsmd + (t).off + (((t).flags & 1) ? tid : 0) + (((t).flags & 2) ? cx : 0))
#define VM_OT_STRIDE(t) ((t).gbase ? (t).S : (((t).flags & 1) ? (long long)NT : 0LL))
          const double *p0 = VM_OT_PTR(t0), *p1 = VM_OT_PTR(t1), *p2 = VM_OT_PTR(t2), *p3 = VM_OT_PTR(t3);
          const long long s0 = VM_OT_STRIDE(t0), s1 = VM_OT_STRIDE(t1), s2 = VM_OT_STRIDE(t2), s3 = VM_OT_STRIDE(t3);
#undef VM_OT_PTR
#undef VM_OT_STRIDE
          double* ps = ts.gbase ? const_cast<double*>(ts.gbase) + ((ts.flags & 2) ? cx : 0) + eb * ts.Cc : vsmd + sc_off;
          const long long ss = ts.S;
          const int c1 = uo.dist, c2 = uo.darg, c3 = (int)(wa.y >> 24);
          const double k1 = p1[0], k2 = p2[0], k3 = p3[0];   // (entry 0 of the operand table is a readable dummy)
          // a Normal stage: sigma is the operand after the stage's own; 1/sigma, 1/sigma^2, log(sigma) once per pass
          double inv = 0.0, inv2 = 0.0, lconst = 0.0;
          bool give_up = false;
          if (c1 >= VM_CH_NLOG || c2 >= VM_CH_NLOG) {
            const double sg = (c1 >= VM_CH_NLOG) ? k2 : k3;
            inv = 1.0 / sg; inv2 = 1.0 / (sg * sg);
            lconst = -log(sg) - SMC_LOG_SQRT_2PI;
            give_up = !(sg > 0.0) && (c1 == VM_CH_NLOG || c2 == VM_CH_NLOG);   // src/distribs.jl:25: sigma <= 0
          }
          double rsum = 0.0;
          int v = 0;
          for (; v + 4 <= vc; v += 4) {
            double r_0 = p0[(v + 0) * s0], r_1 = p0[(v + 1) * s0], r_2 = p0[(v + 2) * s0], r_3 = p0[(v + 3) * s0];
            VM_CH_OPND(q, p1, s1, k1)
            VM_CH_OPND(w, p2, s2, k2)
            VM_CH_OPND(x, p3, s3, k3)
            VM_CH_STAGE(c1, q)
            VM_CH_STAGE(c2, w)
            VM_CH_STAGE(c3, x)
            if (give_up) { r_0 = 0.0; r_1 = 0.0; r_2 = 0.0; r_3 = 0.0; }
            if (fl & VMF_WRITE) { pd[(v + 0) * NT] = r_0; pd[(v + 1) * NT] = r_1; pd[(v + 2) * NT] = r_2; pd[(v + 3) * NT] = r_3; }
            if (fl & VMF_STORE) { ps[(v + 0) * ss] = r_0; ps[(v + 1) * ss] = r_1; ps[(v + 2) * ss] = r_2; ps[(v + 3) * ss] = r_3; }
            if (fl & VMF_REDUCE) { rsum += r_0; rsum += r_1; rsum += r_2; rsum += r_3; }
          }
          for (; v < vc; v++) {
            double r_0 = p0[v * s0], r_1 = 0.0, r_2 = 0.0, r_3 = 0.0;
            const double q_0 = p1[v * s1], w_0 = p2[v * s2], x_0 = p3[v * s3];
            const double q_1 = 1.0, q_2 = 1.0, q_3 = 1.0, w_1 = 1.0, w_2 = 1.0, w_3 = 1.0, x_1 = 1.0, x_2 = 1.0, x_3 = 1.0;
            VM_CH_STAGE(c1, q)
            VM_CH_STAGE(c2, w)
            VM_CH_STAGE(c3, x)
            (void)r_1; (void)r_2; (void)r_3;
            if (give_up) r_0 = 0.0;
            if (fl & VMF_WRITE) pd[v * NT] = r_0;
            if (fl & VMF_STORE) ps[v * ss] = r_0;
            rsum += r_0;
          }
          if (give_up) a.bad[c] = 1;
          if (fl & VMF_REDUCE) Rd[uo.red * NT + tid] += rsum;
        } else if (gen == 0 && (fl == VMF_WRITE || fl == VMF_REDUCE)) {
          const double *p0 = vsmd + i0, *p1 = vsmd + i1, *p2 = vsmd + i2;
          const int s0 = (sel & 1) ? NT : 0, s1 = (sel & 4) ? NT : 0, s2 = (sel & 16) ? NT : 0;
          const double *pa = nullptr;
          double* ps = nullptr;
          const int sa = 0, ss = 0;
          VM_SWITCH
        } else {
          const VmOt t0 = OT[wb.z & 0xff], t1 = OT[(wb.z >> 8) & 0xff], t2 = OT[(wb.z >> 16) & 0xff];
          const double* p0 = t0.gbase ? t0.gbase + ((t0.flags & 2) ? cx : 0) + eb * t0.Cc : vsmd + i0;
          const double* p1 = t1.gbase ? t1.gbase + ((t1.flags & 2) ? cx : 0) + eb * t1.Cc : vsmd + i1;
          const double* p2 = t2.gbase ? t2.gbase + ((t2.flags & 2) ? cx : 0) + eb * t2.Cc : vsmd + i2;
          const long long s0 = t0.gbase ? t0.S : ((t0.flags & 1) ? NT : 0);
          const long long s1 = t1.gbase ? t1.S : ((t1.flags & 1) ? NT : 0);
          const long long s2 = t2.gbase ? t2.S : ((t2.flags & 1) ? NT : 0);
          const VmOt ta = OT[(wb.z >> 24) & 0xff], ts = OT[wb.w & 0xff];
          const double* pa = ta.gbase ? ta.gbase + ((ta.flags & 2) ? cx : 0) + eb * ta.Cc
                                      : vsmd + ta.off + ((ta.flags & 1) ? tid : 0) + ((ta.flags & 2) ? cx : 0);
          const long long sa = ta.gbase ? ta.S : ((ta.flags & 1) ? NT : 0);
          double* ps = ts.gbase ? const_cast<double*>(ts.gbase) + ((ts.flags & 2) ? cx : 0) + eb * ts.Cc : vsmd + sc_off;
          const long long ss = ts.S;
          VM_SWITCH
        }
      }
    }
  }
  if (a.n_red > 0) {
    __syncthreads();
    for (int h = eyn >> 1; h >= 1; h >>= 1) {   // fixed tree over the element lanes
      if (ey < h)
        for (int k = 0; k < a.n_red; k++) Rd[k * NT + tid] += Rd[k * NT + tid + h * tcc];
      __syncthreads();
    }
    if (ey == 0 && active) {
      for (int k = 0; k < a.n_red; k++) {
        const double t = Rd[k * NT + tid];
        if (a.direct) {
          double* q = a.rd[k].p + c * a.rd[k].sc;
          *q = a.racc[k] ? (*q + t) : t;
        } else {
          a.partial[((long long)k * a.S + blockIdx.y) * a.C + c] = t;
        }
      }
    }
  }
}
#undef VM_CH_STAGE
#undef VM_CH_OPND
#undef VM_SWITCH
#undef VM_RUN
#undef VM_LOAD4
#undef VM_LOAD1

struct VmFinalArgs {
  const double* partial;
  int n_red, S;
  long long C;
  Dst d[SMC_PROG_MAX_REDUCE];
  int acc[SMC_PROG_MAX_REDUCE];
};

// dst[k][c] (=|+=) sum_s partial[k][s][c]: thread per output, chunks in order
__global__ void k_vm_final(VmFinalArgs a) {
  long long total = (long long)a.n_red * a.C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int k = (int)(i / a.C);
    long long c = i - (long long)k * a.C;
    double t = 0.0;
    for (int s = 0; s < a.S; s++) t += a.partial[((long long)k * a.S + s) * a.C + c];
    double* q = a.d[k].p + c * a.d[k].sc;
    *q = a.acc[k] ? (*q + t) : t;
  }
}

// few chains, many chunks: one warp per output, lanes stride the chunks, fixed shuffle tree
__global__ void k_vm_final_warp(VmFinalArgs a) {
  const int lane = threadIdx.x & 31;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= (long long)a.n_red * a.C) return;
  int k = (int)(w / a.C);
  long long c = w - (long long)k * a.C;
  double t = 0.0;
  for (int s = lane; s < a.S; s += 32) t += a.partial[((long long)k * a.S + s) * a.C + c];
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (lane == 0) {
    double* q = a.d[k].p + c * a.d[k].sc;
    *q = a.acc[k] ? (*q + t) : t;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------ planning
// Chains: a micro-op whose only effect is to write a local that the NEXT micro-op reads once, and that nobody reads
// afterwards, hands its value over in registers.  Typical model code is chain shaped (resid = y - (a + b .* x), then the
// density of resid): of the locals' shared-memory traffic, which bounds the interpreter, most is such hand-overs.
static void vm_form_chains(std::vector<VmUop>& U) {
  const int n = (int)U.size();
  auto reads = [](const VmUop& u, int d) {
    int c = 0;
    for (int k = 0; k < 3; k++) c += (k < u.nsrc && u.mode[k] == VM_LOCAL && u.idx[k] == d) ? 1 : 0;
    if ((u.flags & VMF_ACC) && u.mode[3] == VM_LOCAL && u.idx[3] == d) c += 100;   // as accumulate operand: never a hand-over
    return c;
  };
  auto dead_after = [&](int j, int d) {
    for (int t = j + 1; t < n; t++) {
      if (reads(U[t], d)) return false;
      if ((U[t].flags & VMF_WRITE) && U[t].dst == d) return true;
    }
    return true;
  };
  auto simple = [](const VmUop& u) {
    if (u.flags & VMF_ACC) return false;
    if (u.op == SMC_OP_NEG) return u.nsrc == 1;
    return (u.op == SMC_OP_ADD || u.op == SMC_OP_SUB || u.op == SMC_OP_MUL || u.op == SMC_OP_DIV) && u.nsrc == 2;
  };
  auto normal = [](const VmUop& u) {
    return !(u.flags & VMF_ACC) && (u.op == SMC_OP_LOGPDF || u.op == SMC_OP_DLOGPDF) && u.dist == SMC_DIST_NORMAL && u.nsrc == 3 &&
           u.mode[1] == VM_SCAL;   // one sigma per chain for the whole pass
  };
  auto code_of = [](int op, bool carry_first) {
    switch (op) {
      case SMC_OP_ADD: return (int)VM_CH_ADD;
      case SMC_OP_SUB: return (int)(carry_first ? VM_CH_SUB : VM_CH_RSUB);
      case SMC_OP_MUL: return (int)VM_CH_MUL;
      case SMC_OP_DIV: return (int)(carry_first ? VM_CH_DIV : VM_CH_RDIV);
      default: return (int)VM_CH_NEG;
    }
  };
  std::vector<VmUop> out;
  for (int i = 0; i < n; i++) {
    const VmUop& h = U[i];
    if (!simple(h) || h.flags != VMF_WRITE) { out.push_back(h); continue; }
    VmUop c;                       // the chain being built
    memset(&c, 0, sizeof(c));
    c.op = VM_OP_CHAIN;
    c.mode[0] = h.mode[0]; c.idx[0] = h.idx[0];
    int codes[3] = {code_of(h.op, true), 0, 0};
    if (h.nsrc == 2) { c.mode[1] = h.mode[1]; c.idx[1] = h.idx[1]; }
    int ns = 1, j = i + 1;
    uint16_t carry = h.dst;
    uint8_t flags = h.flags;
    uint16_t dst = h.dst, store = 0, red = 0;
    bool closed = false;
    while (!closed && j < n && ns < 3 && flags == VMF_WRITE) {
      const VmUop& u = U[j];
      if (reads(u, carry) != 1 || !dead_after(j, carry)) break;
      int pos = 0;
      for (int k = 0; k < u.nsrc; k++)
        if (u.mode[k] == VM_LOCAL && u.idx[k] == carry) pos = k;
      const int slot = ns + 1;       // operand slot of this stage
      if (simple(u)) {
        codes[ns] = code_of(u.op, pos == 0);
        if (u.nsrc == 2) { c.mode[slot] = u.mode[1 - pos]; c.idx[slot] = u.idx[1 - pos]; }
      } else if (normal(u) && pos != 1 && slot + 1 <= 3) {
        const int other = 2 - pos;   // (mu, sigma, x): the carried value is the mean or the variate
        c.mode[slot] = u.mode[other]; c.idx[slot] = u.idx[other];
        c.mode[slot + 1] = u.mode[1]; c.idx[slot + 1] = u.idx[1];
        const bool carry_is_x = (pos == 2);
        if (u.op == SMC_OP_LOGPDF) codes[ns] = VM_CH_NLOG;
        else if (u.darg == 1) codes[ns] = VM_CH_ND_SIG;
        else if (u.darg == 0) codes[ns] = carry_is_x ? VM_CH_ND_RMO : VM_CH_ND_OMR;   // (x - mu) / sigma^2
        else codes[ns] = carry_is_x ? VM_CH_ND_OMR : VM_CH_ND_RMO;                    // (mu - x) / sigma^2
        closed = true;
      } else {
        break;
      }
      ns++;
      flags = u.flags; dst = u.dst; store = u.store; red = u.red;
      carry = u.dst;
      j++;
    }
    if (ns == 1) { out.push_back(h); continue; }
    c.dist = (uint8_t)codes[0]; c.darg = (uint8_t)codes[1]; c.code3 = (uint32_t)codes[2];
    c.flags = flags; c.dst = dst; c.store = store; c.red = red; c.nsrc = 4;
    out.push_back(c);
    i = j - 1;
  }
  U.swap(out);
}

int smc_vm_plan(smc_model_s* m, SmcProgram& pr) {
  (void)m;
  struct Src { uint8_t mode = VM_NONE; uint16_t idx = 0; };
  Src cur[SMC_PROG_MAX_LOCALS];
  int last_writer[SMC_PROG_MAX_LOCALS];
  for (int i = 0; i < SMC_PROG_MAX_LOCALS; i++) last_writer[i] = -1;
  std::vector<VmUop>& U = pr.uops;
  U.clear();
  pr.desc_regs.clear();
  pr.scal_regs.clear();
  std::vector<uint32_t> stored;
  auto index_of = [](std::vector<uint32_t>& v, uint32_t reg) {
    for (size_t i = 0; i < v.size(); i++)
      if (v[i] == reg) return (uint16_t)i;
    v.push_back(reg);
    return (uint16_t)(v.size() - 1);
  };
  auto was_stored = [&](uint32_t reg) { return std::find(stored.begin(), stored.end(), reg) != stored.end(); };
  // slot d is about to be redefined: every other slot that is just a name for "local d" gets its own copy first
  auto materialise_aliases = [&](int d) {
    for (int t = 0; t < SMC_PROG_MAX_LOCALS; t++) {
      if (t == d || cur[t].mode != VM_LOCAL || cur[t].idx != d) continue;
      VmUop u;
      memset(&u, 0, sizeof(u));
      u.op = SMC_OP_COPY; u.nsrc = 1; u.mode[0] = VM_LOCAL; u.idx[0] = (uint16_t)d; u.dst = (uint16_t)t; u.flags = VMF_WRITE;
      U.push_back(u);
      cur[t].idx = (uint16_t)t;
      last_writer[t] = (int)U.size() - 1;
    }
  };
  const int NI = (int)pr.instr.size();
  for (int i = 0; i < NI; i++) {
    const smc_prog_instr& in = pr.instr[i];
    if (in.op == SMC_P_LOADS || in.op == SMC_P_LOADV) materialise_aliases(in.dst);
    if (in.op == SMC_P_LOADS || (in.op == SMC_P_LOADV && pr.h.n == 1)) {
      if (was_stored(in.reg)) { smc_set_error("program: register %u is reloaded after being stored", in.reg); return SMC_ERR_INVALID; }
      cur[in.dst].mode = VM_SCAL;
      cur[in.dst].idx = index_of(pr.scal_regs, in.reg);
      last_writer[in.dst] = -1;
    } else if (in.op == SMC_P_LOADV) {
      if (was_stored(in.reg)) { smc_set_error("program: register %u is reloaded after being stored", in.reg); return SMC_ERR_INVALID; }
      bool stored_later = false;
      for (int j = i + 1; j < NI; j++)
        if (pr.instr[j].op == SMC_P_STOREV && pr.instr[j].reg == in.reg) stored_later = true;
      uint16_t d = index_of(pr.desc_regs, in.reg);
      if (stored_later) {   // read-modify-write of one register: take the value now
        VmUop u;
        memset(&u, 0, sizeof(u));
        u.op = SMC_OP_COPY; u.nsrc = 1; u.mode[0] = VM_ARRAY; u.idx[0] = d; u.dst = in.dst; u.flags = VMF_WRITE;
        U.push_back(u);
        cur[in.dst].mode = VM_LOCAL; cur[in.dst].idx = in.dst;
        last_writer[in.dst] = (int)U.size() - 1;
      } else {
        cur[in.dst].mode = VM_ARRAY; cur[in.dst].idx = d;
        last_writer[in.dst] = -1;
      }
    } else if (in.op == SMC_P_STOREV || in.op == SMC_P_REDUCE) {
      const Src s = cur[in.src[0]];
      if (s.mode == VM_NONE) { smc_set_error("program: instruction %d reads an undefined local", i); return SMC_ERR_INVALID; }
      const uint8_t flag = in.op == SMC_P_STOREV ? VMF_STORE : VMF_REDUCE;
      VmUop* tgt = nullptr;
      if (s.mode == VM_LOCAL && !U.empty() && last_writer[s.idx] == (int)U.size() - 1 && !(U.back().flags & flag)) tgt = &U.back();
      if (!tgt) {
        VmUop u;
        memset(&u, 0, sizeof(u));
        u.op = SMC_OP_COPY; u.nsrc = 1; u.mode[0] = s.mode; u.idx[0] = s.idx;
        U.push_back(u);
        tgt = &U.back();
      }
      tgt->flags |= flag;
      if (in.op == SMC_P_STOREV) {
        tgt->store = index_of(pr.desc_regs, in.reg);
        stored.push_back(in.reg);
      } else {
        tgt->red = in.aux0;
      }
    } else if (in.op == SMC_OP_COPY && !(in.flags & SMC_F_ACC) && in.nsrc == 1 && cur[in.src[0]].mode == VM_LOCAL &&
               cur[in.src[0]].idx == in.dst && cur[in.dst].mode == VM_LOCAL && cur[in.dst].idx == in.dst) {
      // copying a name of slot d back into slot d: nothing to do
    } else if (in.op == SMC_OP_COPY && !(in.flags & SMC_F_ACC) && in.nsrc == 1 && cur[in.src[0]].mode != VM_NONE &&
               !(cur[in.src[0]].mode == VM_LOCAL && cur[in.src[0]].idx == in.dst)) {
      // copy propagation (the pass-through adjoints `dx = dx + ds` of src/diff.jl:29-38 are plain copies): the
      // destination becomes another name of the source value; it is materialised only if that slot is overwritten
      // while the name is still in use (see `materialise_aliases`)
      materialise_aliases(in.dst);
      cur[in.dst] = cur[in.src[0]];
      last_writer[in.dst] = -1;
    } else {
      VmUop u;
      memset(&u, 0, sizeof(u));
      u.op = (uint8_t)in.op; u.dist = (uint8_t)in.aux0; u.darg = (uint8_t)in.aux1; u.nsrc = in.nsrc;
      if (in.op > 255 || in.aux0 > 255 || in.aux1 > 255) { smc_set_error("program: instruction %d has an operand out of range", i); return SMC_ERR_INVALID; }
      // operands are resolved before other names of the destination slot get their own copy: the micro-op reads its
      // operands before it writes, so it may keep reading the slot it is about to overwrite
      for (int k = 0; k < in.nsrc; k++) {
        const Src s = cur[in.src[k]];
        if (s.mode == VM_NONE) { smc_set_error("program: instruction %d reads an undefined local", i); return SMC_ERR_INVALID; }
        u.mode[k] = s.mode; u.idx[k] = s.idx;
      }
      u.flags = VMF_WRITE;
      if (in.flags & SMC_F_ACC) {
        const Src s = cur[in.dst];
        if (s.mode == VM_NONE) { smc_set_error("program: instruction %d accumulates into an undefined local", i); return SMC_ERR_INVALID; }
        u.flags |= VMF_ACC; u.mode[3] = s.mode; u.idx[3] = s.idx;
      }
      u.dst = in.dst;
      materialise_aliases(in.dst);
      U.push_back(u);
      cur[in.dst].mode = VM_LOCAL; cur[in.dst].idx = in.dst;
      last_writer[in.dst] = (int)U.size() - 1;
    }
  }
  // dead local writes (values consumed only through a fused store / reduction) and dense renumbering of the slots
  {
    bool live[SMC_PROG_MAX_LOCALS] = {false};
    for (int i = (int)U.size() - 1; i >= 0; i--) {
      VmUop& u = U[i];
      if (u.flags & VMF_WRITE) {
        if (!live[u.dst]) u.flags &= ~VMF_WRITE;
        live[u.dst] = false;
      }
      for (int k = 0; k < 4; k++)
        if (u.mode[k] == VM_LOCAL) live[u.idx[k]] = true;
    }
    std::vector<VmUop> kept;
    for (auto& u : U)
      if (u.flags & (VMF_WRITE | VMF_STORE | VMF_REDUCE)) kept.push_back(u);
    U.swap(kept);
    if (!getenv("SMC_VM_NO_CHAINS")) vm_form_chains(U);
    int remap[SMC_PROG_MAX_LOCALS];
    for (int i = 0; i < SMC_PROG_MAX_LOCALS; i++) remap[i] = -1;
    int n = 0;
    auto slot = [&](uint16_t s) { if (remap[s] < 0) remap[s] = n++; return (uint16_t)remap[s]; };
    for (auto& u : U) {
      for (int k = 0; k < 4; k++)
        if (u.mode[k] == VM_LOCAL) u.idx[k] = slot(u.idx[k]);
      if (u.flags & VMF_WRITE) u.dst = slot(u.dst);
    }
    pr.n_loc = n;
  }
  if (const char* dump = getenv("SMC_VM_DUMP")) {
    if (dump[0] == '1') {
      static const char* mode_name[] = {"-", "L", "S", "A"};
      fprintf(stderr, "[vm] program n=%u: %zu instructions -> %zu micro-ops, %d locals, %zu scalars, %zu arrays\n", pr.h.n,
              pr.instr.size(), U.size(), pr.n_loc, pr.scal_regs.size(), pr.desc_regs.size());
      for (auto& u : U) {
        fprintf(stderr, "[vm]   op %2u", u.op);
        if (u.op == SMC_OP_LOGPDF || u.op == SMC_OP_DLOGPDF) fprintf(stderr, "(dist %u arg %u)", u.dist, u.darg);
        if (u.op == VM_OP_CHAIN) fprintf(stderr, "(chain stages %u %u %u)", u.dist, u.darg, u.code3);
        for (int k = 0; k < (u.op == VM_OP_CHAIN ? 4 : 3); k++)
          if (u.mode[k]) fprintf(stderr, " %s%u", mode_name[u.mode[k]], u.idx[k]);
        if (u.flags & VMF_ACC) fprintf(stderr, " +=%s%u", mode_name[u.mode[3]], u.idx[3]);
        if (u.flags & VMF_WRITE) fprintf(stderr, " -> L%u", u.dst);
        if (u.flags & VMF_STORE) fprintf(stderr, " -> A%u", u.store);
        if (u.flags & VMF_REDUCE) fprintf(stderr, " -> R%u", u.red);
        fprintf(stderr, "\n");
      }
    }
  }
  return SMC_OK;
}

extern "C" int smc_vm_plan_describe(const void* program, size_t nbytes, char* out, size_t cap) {
  if (!program || !out || cap == 0) { smc_set_error("smc_vm_plan_describe: NULL argument"); return SMC_ERR_INVALID; }
  const unsigned char* p = (const unsigned char*)program;
  SmcProgram pr;
  if (nbytes < sizeof(pr.h)) { smc_set_error("smc_vm_plan_describe: truncated program"); return SMC_ERR_INVALID; }
  memcpy(&pr.h, p, sizeof(pr.h));
  const size_t need = sizeof(pr.h) + (size_t)pr.h.n_instr * sizeof(smc_prog_instr) + (size_t)pr.h.n_red * sizeof(smc_prog_red);
  if (nbytes < need) { smc_set_error("smc_vm_plan_describe: truncated program"); return SMC_ERR_INVALID; }
  pr.instr.resize(pr.h.n_instr);
  if (pr.h.n_instr) memcpy(pr.instr.data(), p + sizeof(pr.h), (size_t)pr.h.n_instr * sizeof(smc_prog_instr));
  pr.red.resize(pr.h.n_red);
  if (pr.h.n_red) memcpy(pr.red.data(), p + sizeof(pr.h) + (size_t)pr.h.n_instr * sizeof(smc_prog_instr), (size_t)pr.h.n_red * sizeof(smc_prog_red));
  SMC_TRY(smc_vm_plan(nullptr, pr));
  static const char* mode_name[] = {"-", "L", "S", "A"};
  size_t w = 0;
  out[0] = 0;
  auto put = [&](const char* fmt, ...) {
    if (w + 1 >= cap) return;
    va_list ap;
    va_start(ap, fmt);
    int k = vsnprintf(out + w, cap - w, fmt, ap);
    va_end(ap);
    if (k > 0) w = std::min(cap - 1, w + (size_t)k);
  };
  for (auto& u : pr.uops) {
    if (u.op == VM_OP_CHAIN) put("chain(%u,%u,%u)", u.dist, u.darg, u.code3);
    else if (u.op == SMC_OP_LOGPDF || u.op == SMC_OP_DLOGPDF) put("op%u(dist%u,arg%u)", u.op, u.dist, u.darg);
    else put("op%u", u.op);
    for (int k = 0; k < (u.op == VM_OP_CHAIN ? 4 : 3); k++) put(" %s%u", mode_name[u.mode[k]], u.mode[k] ? u.idx[k] : 0);
    if (u.flags & VMF_ACC) put(" +=%s%u", mode_name[u.mode[3]], u.idx[3]);
    if (u.flags & VMF_WRITE) put(" ->L%u", u.dst);
    if (u.flags & VMF_STORE) put(" ->A%u", u.store);
    if (u.flags & VMF_REDUCE) put(" ->R%u", u.red);
    put("\n");
  }
  put("locals=%d scalars=%zu arrays=%zu\n", pr.n_loc, pr.scal_regs.size(), pr.desc_regs.size());
  return (int)pr.uops.size();
}

static VmDesc desc_of(const smc_model_s* m, uint32_t reg) {
  Opd o = smc_make_opd(m, reg);
  VmDesc d;
  d.p = o.is_imm ? nullptr : o.p;
  d.se = o.se; d.sc = o.sc; d.imm = o.imm;
  return d;
}

// device tables: VmDevUop[n_uops (padded to even)] | VmDevOpd[n_ot] | VmDesc[n_scal]
int smc_vm_prepare(smc_model_s* m, SmcProgram& pr) {
  const size_t nu = pr.uops.size(), ns = pr.scal_regs.size();
  std::vector<VmDevOpd> ot(1);
  memset(&ot[0], 0, sizeof(VmDevOpd));   // entry 0: "no operand" (reads a valid dummy)
  auto entry = [&](int kind, int idx) -> int {
    if (kind == VM_NONE) return 0;
    for (size_t i = 1; i < ot.size(); i++)
      if (ot[i].kind == kind && ot[i].idx == idx) return (int)i;
    VmDevOpd d;
    memset(&d, 0, sizeof(d));
    d.kind = kind; d.idx = idx;
    if (kind == VM_ARRAY) {
      VmDesc a = desc_of(m, pr.desc_regs[idx]);
      d.p = a.p; d.se = a.se; d.sc = a.sc;
    }
    ot.push_back(d);
    return (int)ot.size() - 1;
  };
  std::vector<VmDevUop> du(nu);
  for (size_t i = 0; i < nu; i++) {
    const VmUop& u = pr.uops[i];
    VmDevUop d;
    memset(&d, 0, sizeof(d));
    d.op = u.op; d.dist = u.dist; d.darg = u.darg; d.flags = u.flags;
    for (int k = 0; k < 4; k++) {
      if (u.mode[k] == VM_ARRAY && !desc_of(m, pr.desc_regs[u.idx[k]]).p) { smc_set_error("program: an immediate is used as an array operand"); return SMC_ERR_INVALID; }
      d.o[k] = (uint8_t)entry(u.mode[k], u.idx[k]);
    }
    d.dst = (uint8_t)((u.flags & VMF_WRITE) ? entry(VM_LOCAL, u.dst) : 0);
    d.store = (uint8_t)((u.flags & VMF_STORE) ? entry(VM_ARRAY, u.store) : 0);
    d.red = (uint8_t)u.red;
    d.code3 = (uint8_t)(u.op == VM_OP_CHAIN ? u.code3 : 0);
    du[i] = d;
    if (ot.size() > 255) { smc_set_error("program: more than 255 distinct operands"); return SMC_ERR_INVALID; }
  }
  pr.n_ot = (int)ot.size();
  const size_t nu2 = (nu + 1) & ~(size_t)1;
  const size_t bytes = nu2 * sizeof(VmDevUop) + ot.size() * sizeof(VmDevOpd) + ns * sizeof(VmDesc);
  std::vector<unsigned char> host(std::max<size_t>(bytes, 32), 0);
  if (nu) memcpy(host.data(), du.data(), nu * sizeof(VmDevUop));
  memcpy(host.data() + nu2 * sizeof(VmDevUop), ot.data(), ot.size() * sizeof(VmDevOpd));
  VmDesc* D = (VmDesc*)(host.data() + nu2 * sizeof(VmDevUop) + ot.size() * sizeof(VmDevOpd));
  for (size_t i = 0; i < ns; i++) D[i] = desc_of(m, pr.scal_regs[i]);
  if (!pr.dev_tables || pr.dev_tables_bytes < host.size()) {
    if (pr.dev_tables) smc_dev_free(pr.dev_tables, m->ctx->stream);
    pr.dev_tables = nullptr;
    SMC_CUDA(smc_dev_alloc(&pr.dev_tables, host.size(), m->ctx->stream));
    pr.dev_tables_bytes = host.size();
  }
  SMC_CUDA(cudaMemcpyAsync(pr.dev_tables, host.data(), host.size(), cudaMemcpyHostToDevice, m->ctx->stream));
  SMC_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return SMC_OK;
}

int smc_vm_launch(smc_model_s* m, SmcProgram& pr, int64_t C, int* bad) {
  cudaStream_t st = m->ctx->stream;
  static SmcOncePerDevice once;
  const size_t smem_max = 200 * 1024;
  if (once.first()) SMC_CUDA(cudaFuncSetAttribute(k_vm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
  VmArgs a;
  memset(&a, 0, sizeof(a));
  const size_t nu = pr.uops.size(), ns = pr.scal_regs.size();
  const size_t nu2 = (nu + 1) & ~(size_t)1;
  a.uops = (const VmDevUop*)pr.dev_tables;
  a.ot = (const VmDevOpd*)((const unsigned char*)pr.dev_tables + nu2 * sizeof(VmDevUop));
  a.scal = (const VmDesc*)(a.ot + pr.n_ot);
  a.n_uops = (int)nu; a.n_ot = pr.n_ot; a.n_scal = (int)ns; a.n_loc = pr.n_loc; a.n_red = (int)pr.h.n_red;
  a.n = pr.h.n; a.C = C; a.bad = bad;
  // Thread geometry.  Chains fill the block first (coalesced per-chain operands); the index space is split over
  // element lanes of the block and over grid chunks only as far as needed to put ~2 blocks on every SM, so that a
  // thread keeps as many elements as possible (a micro-op is decoded once per thread and pass).
  const long long want_threads = 2LL * m->ctx->sm_count * NT;
  // (rounded down: one full wave of blocks beats one and a fraction)
  long long lanes = std::max<long long>(1, std::min<long long>(want_threads / C, std::max<long long>(1, a.n / 4)));
  int tcc = 1;
  while (tcc < C && tcc < NT) tcc <<= 1;                 // every chain of a small run in one block ...
  while (tcc > 32 && NT / tcc < lanes) tcc >>= 1;        // ... but give up chain width for element lanes down to a warp
  a.tcc = tcc;
  const int eyn = NT / tcc;
  const long long chain_tiles = (C + tcc - 1) / tcc;
  long long S = std::max<long long>(1, std::min<long long>(lanes / eyn, (a.n + (long long)eyn * 4 - 1) / ((long long)eyn * 4)));
  a.chunk = (a.n + S - 1) / S;
  S = (a.n + a.chunk - 1) / a.chunk;
  a.S = (int)S;
  // V = elements a thread processes per decode of a micro-op: as many as the locals' shared-memory budget allows
  const long long per_thread = (a.chunk + eyn - 1) / eyn;
  const size_t fixed = nu2 * 16 + nu * 32 + (size_t)pr.n_ot * 32 + std::max<size_t>(1, ns) * tcc * 8 + (size_t)a.n_red * NT * 8;
  const size_t budget = 96 * 1024;   // two blocks per SM (three blocks at 80 registers measured slower: spills)
  long long vmax = pr.n_loc > 0 ? (long long)((budget > fixed ? budget - fixed : 0) / ((size_t)pr.n_loc * NT * 8)) : 32;
  vmax = std::max<long long>(1, std::min<long long>(vmax, 32));
  long long passes = (per_thread + vmax - 1) / vmax;
  a.V = (int)std::max<long long>(1, (per_thread + passes - 1) / passes);
  a.direct = (S == 1) ? 1 : 0;
  for (int k = 0; k < a.n_red; k++) { a.rd[k] = smc_make_dst(m, pr.red[k].reg); a.racc[k] = (int)pr.red[k].acc; }
  if (a.n_red > 0 && !a.direct) SMC_TRY(smc_scratch_reserve(m, (int64_t)a.n_red * S * C));
  a.partial = m->scratch;
  const size_t smem = fixed + (size_t)pr.n_loc * a.V * NT * 8;
  if (smem > smem_max) { smc_set_error("program needs %zu bytes of shared memory", smem); return SMC_ERR_MEMORY; }
  dim3 grid((unsigned)chain_tiles, (unsigned)S);
  k_vm<<<grid, NT, smem, st>>>(a);
  m->ctx->launches++;
  if (a.n_red > 0 && !a.direct) {
    VmFinalArgs f;
    f.partial = m->scratch; f.n_red = a.n_red; f.S = (int)S; f.C = C;
    for (int k = 0; k < a.n_red; k++) { f.d[k] = a.rd[k]; f.acc[k] = a.racc[k]; }
    const long long outs = (long long)a.n_red * C;
    if (outs < 4096 && S >= 8) k_vm_final_warp<<<(unsigned)((outs * 32 + 255) / 256), 256, 0, st>>>(f);
    else k_vm_final<<<(unsigned)std::max<long long>(1, std::min<long long>((outs + 255) / 256, 148LL * 8)), 256, 0, st>>>(f);
    m->ctx->launches++;
  }
  return SMC_OK;
}
