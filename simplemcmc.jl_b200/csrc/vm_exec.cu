// Register-program VM: executes one SMC_OP_FUSED statement -- a run of elementwise / reducing tape statements over one
// index space (include/simplemcmc_tape.h, smc_prog_instr) -- in a single launch for all chains in lockstep.
//
// The reference's generated function materialises one temporary array per statement
// (/root/reference/src/parsing.jl:409-467); here every intermediate stays on chip:
//   * plan time (smc_vm_plan): the instruction list becomes micro-ops whose operands are addressed directly -- a
//     program-local vector slot, a per-chain scalar staged once per block, or a tape array in HBM -- so LOADV / LOADS /
//     STOREV / REDUCE instructions disappear into operand modes and side effects of the computing instruction;
//   * run time (k_vm<V>): thread = (chain lane, element lane); each thread owns V elements per pass, so one decode of
//     a micro-op is amortised over V elements.  Locals live in shared memory as [slot][v][thread] (conflict free),
//     chains are the fastest thread index (coalesced per-chain operands, broadcast data operands).
// Reductions (`sum`, `dot`, vectorised logpdf: src/distribs.jl:97-105) accumulate per thread in element order, then
// over the element lanes by a fixed tree, then over element chunks in k_vm_final: bit-reproducible for a given chain
// count.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include "exec_common.cuh"

namespace {

constexpr int NT = 256;  // threads per block

struct VmArgs {
  const VmUop* uops;
  const VmDesc* descs;
  const VmDesc* scal;
  int n_uops, n_descs, n_scal, n_loc, n_red, tcc, S, direct;
  long long n, C, chunk;
  int* bad;
  double* partial;  // [n_red][S][C]
  Dst rd[SMC_PROG_MAX_REDUCE];   // destinations of the reduction slots (direct mode: S == 1)
  int racc[SMC_PROG_MAX_REDUCE];
};

// the heavy operations (densities, their adjoints, pow, trigonometry, ...) are called, not inlined V times
__device__ __noinline__ double vm_apply_slow(int op, int dist, int k, double a, double b, double c) {
  return smc_apply(op, dist, k, a, b, c);
}

template <int V>
__global__ void __launch_bounds__(NT) k_vm(VmArgs a) {
  extern __shared__ __align__(16) unsigned char vsm[];
  VmUop* U = (VmUop*)vsm;
  VmDesc* D = (VmDesc*)(U + a.n_uops);
  double* Sc = (double*)(D + a.n_descs);   // [n_scal][tcc]
  double* Lc = Sc + a.n_scal * a.tcc;      // [n_loc][V][NT]
  double* Rd = Lc + (size_t)a.n_loc * V * NT;  // [n_red][NT]
  const int tid = threadIdx.x;
  {
    const uint4* src = (const uint4*)a.uops;   // uops and descs are contiguous in the device table
    uint4* dst = (uint4*)vsm;
    const int words = (a.n_uops + a.n_descs) * 2;
    for (int i = tid; i < words; i += NT) dst[i] = src[i];
  }
  const int tcc = a.tcc, eyn = NT / tcc;
  const int cx = tid % tcc, ey = tid / tcc;
  const long long cbase = (long long)blockIdx.x * tcc;
  const long long c = cbase + cx;
  const bool active = c < a.C;
  for (int i = tid; i < a.n_scal * tcc; i += NT) {
    const int k = i / tcc, cl = i - k * tcc;
    const VmDesc d = a.scal[k];
    const long long cc = cbase + cl;
    Sc[i] = d.p ? (cc < a.C ? __ldg(d.p + cc * d.sc) : 0.0) : d.imm;
  }
  for (int k = 0; k < a.n_red; k++) Rd[k * NT + tid] = 0.0;
  __syncthreads();

  if (active) {
    const long long e0 = (long long)blockIdx.y * a.chunk, e1 = min(a.n, e0 + a.chunk);
    for (long long eb = e0 + ey; eb < e1; eb += (long long)eyn * V) {
      long long ev[V];
      bool ok[V];
#pragma unroll
      for (int v = 0; v < V; v++) {
        long long e = eb + (long long)v * eyn;
        ok[v] = e < e1;
        ev[v] = ok[v] ? e : e1 - 1;   // tail lanes recompute the last element; their side effects are masked
      }
      auto fetch = [&](int mode, int idx, double (&x)[V]) {
        if (mode == VM_LOCAL) {
          const double* p = Lc + (size_t)idx * V * NT + tid;
#pragma unroll
          for (int v = 0; v < V; v++) x[v] = p[v * NT];
        } else if (mode == VM_SCAL) {
          const double t = Sc[idx * tcc + cx];
#pragma unroll
          for (int v = 0; v < V; v++) x[v] = t;
        } else if (mode == VM_ARRAY) {
          const VmDesc d = D[idx];
          const double* p = d.p + c * d.sc;
#pragma unroll
          for (int v = 0; v < V; v++) x[v] = __ldg(p + ev[v] * d.se);
        } else {
#pragma unroll
          for (int v = 0; v < V; v++) x[v] = 0.0;
        }
      };
      for (int u = 0; u < a.n_uops; u++) {
        const VmUop uo = U[u];
        double x0[V], x1[V], r[V];
        fetch(uo.mode[0], uo.idx[0], x0);
        fetch(uo.mode[1], uo.idx[1], x1);
#define VM_EACH(expr)                        \
  _Pragma("unroll") for (int v = 0; v < V; v++) { \
    const double p = x0[v], q = x1[v];       \
    (void)q;                                 \
    r[v] = (expr);                           \
  }                                          \
  break;
        switch (uo.op) {
          case SMC_OP_COPY: VM_EACH(p)
          case SMC_OP_NEG: VM_EACH(-p)
          case SMC_OP_ADD: VM_EACH(p + q)
          case SMC_OP_SUB: VM_EACH(p - q)
          case SMC_OP_MUL: VM_EACH(p * q)
          case SMC_OP_DIV: VM_EACH(p / q)
          case SMC_OP_EXP: VM_EACH(exp(p))
          case SMC_OP_LOG: VM_EACH(log(p))
          case SMC_OP_ZERO: VM_EACH(0.0)
          default: {
            double x2[V];
            fetch(uo.mode[2], uo.idx[2], x2);
#pragma unroll
            for (int v = 0; v < V; v++) {
              double t = vm_apply_slow(uo.op, uo.dist, uo.darg, x0[v], x1[v], x2[v]);
              if (uo.op == SMC_OP_LOGPDF && (t == -INFINITY || isnan(t))) {   // "give up eval", src/distribs.jl:69-70
                if (ok[v]) a.bad[c] = 1;
                t = 0.0;
              }
              r[v] = t;
            }
          }
        }
#undef VM_EACH
        if (uo.flags & VMF_ACC) {
          double xa[V];
          fetch(uo.mode[3], uo.idx[3], xa);
#pragma unroll
          for (int v = 0; v < V; v++) r[v] = xa[v] + r[v];
        }
        if (uo.flags & VMF_WRITE) {
          double* p = Lc + (size_t)uo.dst * V * NT + tid;
#pragma unroll
          for (int v = 0; v < V; v++) p[v * NT] = r[v];
        }
        if (uo.flags & VMF_STORE) {
          const VmDesc d = D[uo.store];
          double* p = const_cast<double*>(d.p) + c * d.sc;
#pragma unroll
          for (int v = 0; v < V; v++)
            if (ok[v]) p[ev[v] * d.se] = r[v];
        }
        if (uo.flags & VMF_REDUCE) {
          double t = Rd[uo.red * NT + tid];
#pragma unroll
          for (int v = 0; v < V; v++)
            if (ok[v]) t += r[v];
          Rd[uo.red * NT + tid] = t;
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
  const int NI = (int)pr.instr.size();
  for (int i = 0; i < NI; i++) {
    const smc_prog_instr& in = pr.instr[i];
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
    } else {
      VmUop u;
      memset(&u, 0, sizeof(u));
      u.op = (uint8_t)in.op; u.dist = (uint8_t)in.aux0; u.darg = (uint8_t)in.aux1; u.nsrc = in.nsrc;
      if (in.op > 255 || in.aux0 > 255 || in.aux1 > 255) { smc_set_error("program: instruction %d has an operand out of range", i); return SMC_ERR_INVALID; }
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
  return SMC_OK;
}

static VmDesc desc_of(const smc_model_s* m, uint32_t reg) {
  Opd o = smc_make_opd(m, reg);
  VmDesc d;
  d.p = o.is_imm ? nullptr : o.p;
  d.se = o.se; d.sc = o.sc; d.imm = o.imm;
  return d;
}

int smc_vm_prepare(smc_model_s* m, SmcProgram& pr) {
  const size_t nu = pr.uops.size(), nd = pr.desc_regs.size(), ns = pr.scal_regs.size();
  const size_t bytes = nu * sizeof(VmUop) + (nd + ns) * sizeof(VmDesc);
  std::vector<unsigned char> host(std::max<size_t>(bytes, 32));
  if (nu) memcpy(host.data(), pr.uops.data(), nu * sizeof(VmUop));
  VmDesc* D = (VmDesc*)(host.data() + nu * sizeof(VmUop));
  for (size_t i = 0; i < nd; i++) {
    D[i] = desc_of(m, pr.desc_regs[i]);
    if (!D[i].p) { smc_set_error("program: an immediate is used as an array operand"); return SMC_ERR_INVALID; }
  }
  for (size_t i = 0; i < ns; i++) D[nd + i] = desc_of(m, pr.scal_regs[i]);
  if (!pr.dev_tables || pr.dev_tables_bytes < host.size()) {
    if (pr.dev_tables) cudaFree(pr.dev_tables);
    pr.dev_tables = nullptr;
    SMC_CUDA(cudaMalloc(&pr.dev_tables, host.size()));
    pr.dev_tables_bytes = host.size();
  }
  SMC_CUDA(cudaMemcpyAsync(pr.dev_tables, host.data(), host.size(), cudaMemcpyHostToDevice, m->ctx->stream));
  SMC_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return SMC_OK;
}

int smc_vm_launch(smc_model_s* m, SmcProgram& pr, int64_t C, int* bad) {
  cudaStream_t st = m->ctx->stream;
  static bool attr_done = false;
  if (!attr_done) {
    SMC_CUDA(cudaFuncSetAttribute(k_vm<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SMC_CUDA(cudaFuncSetAttribute(k_vm<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SMC_CUDA(cudaFuncSetAttribute(k_vm<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  VmArgs a;
  memset(&a, 0, sizeof(a));
  const size_t nu = pr.uops.size(), nd = pr.desc_regs.size(), ns = pr.scal_regs.size();
  a.uops = (const VmUop*)pr.dev_tables;
  a.descs = (const VmDesc*)((const unsigned char*)pr.dev_tables + nu * sizeof(VmUop));
  a.scal = a.descs + nd;
  a.n_uops = (int)nu; a.n_descs = (int)nd; a.n_scal = (int)ns; a.n_loc = pr.n_loc; a.n_red = (int)pr.h.n_red;
  a.n = pr.h.n; a.C = C; a.bad = bad;
  int tcc = 1;
  while (tcc < C && tcc < (a.n == 1 ? NT : 32)) tcc <<= 1;
  a.tcc = tcc;
  const int eyn = NT / tcc;
  int V = 1;
  const size_t loc_budget = 64 * 1024;
  if (a.n >= (long long)eyn * 4 && (size_t)pr.n_loc * 4 * NT * 8 <= loc_budget) V = 4;
  else if (a.n >= (long long)eyn * 2 && (size_t)pr.n_loc * 2 * NT * 8 <= loc_budget) V = 2;
  const long long chain_tiles = (C + tcc - 1) / tcc;
  const long long per_pass = (long long)eyn * V;
  long long S = std::max<long long>(1, std::min<long long>((2LL * m->ctx->sm_count + chain_tiles - 1) / chain_tiles,
                                                           (a.n + per_pass * 2 - 1) / (per_pass * 2)));
  a.chunk = (a.n + S - 1) / S;
  S = (a.n + a.chunk - 1) / a.chunk;
  a.S = (int)S;
  a.direct = (S == 1) ? 1 : 0;
  for (int k = 0; k < a.n_red; k++) { a.rd[k] = smc_make_dst(m, pr.red[k].reg); a.racc[k] = (int)pr.red[k].acc; }
  if (a.n_red > 0 && !a.direct) SMC_TRY(smc_scratch_reserve(m, (int64_t)a.n_red * S * C));
  a.partial = m->scratch;
  const size_t smem = (nu + nd) * 32 + (size_t)ns * tcc * 8 + (size_t)pr.n_loc * V * NT * 8 + (size_t)a.n_red * NT * 8;
  if (smem > 200 * 1024) { smc_set_error("program needs %zu bytes of shared memory", smem); return SMC_ERR_MEMORY; }
  dim3 grid((unsigned)chain_tiles, (unsigned)S);
  if (V == 4) k_vm<4><<<grid, NT, smem, st>>>(a);
  else if (V == 2) k_vm<2><<<grid, NT, smem, st>>>(a);
  else k_vm<1><<<grid, NT, smem, st>>>(a);
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
