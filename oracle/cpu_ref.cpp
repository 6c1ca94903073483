// ORACLE / CPU BASELINE (test infrastructure only; never linked into the product).
//
// C++ restatement of the closure SimpleMCMC.jl generates for the two GLM examples, statement by statement
// (one temporary vector per statement, sequential `+=` sums as in /root/reference/src/distribs.jl:100-102),
// and of the simpleHMC loop around it (/root/reference/src/samplers.jl:50-59,146-167).  Statement lists:
// SURVEY.md Appendix A.1 (examples/linear.jl) and A.2 (examples/binomial.jl).  One chain per core (std::thread):
// this is the "whole box of host cores" figure reported next to the GPU numbers.
//
// Build: make -C oracle   (g++ -O3 -march=x86-64-v3 -pthread -shared -fPIC)
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>
#include <atomic>
#include <thread>

namespace {
const double kFac = -std::log(std::sqrt(2.0 * M_PI));

struct Work {  // the temporaries the generated function allocates on every call
  std::vector<double> t2, resid, t4, t6, dresid, dt2, t3v, t4v, prob, dprob, dt4v, dt3v;
  explicit Work(int64_t N)
      : t2(N), resid(N), t4(N), t6(N), dresid(N), dt2(N), t3v(N), t4v(N), prob(N), dprob(N), dt4v(N), dt3v(N) {}
};

// logpdfNormal(mu::Real, sigma::Real, x::Array): sequential sum (src/distribs.jl:23-28,97-105)
inline double logpdf_normal_vec(double mu, double sigma, const double* x, int64_t n) {
  double res = 0.0;
  for (int64_t i = 0; i < n; i++) {
    double r = (x[i] - mu) / sigma;
    res += -r * r * 0.5 - std::log(sigma) + kFac;
  }
  return res;
}

// X column major [N x M]
inline void gemv_n(const double* X, int64_t N, int M, const double* b, double* out) {
  std::memset(out, 0, sizeof(double) * N);
  for (int j = 0; j < M; j++) {
    const double* col = X + (int64_t)j * N;
    double bj = b[j];
    for (int64_t i = 0; i < N; i++) out[i] += col[i] * bj;
  }
}
inline void gemv_t(const double* X, int64_t N, int M, const double* d, double* out) {
  for (int j = 0; j < M; j++) {
    const double* col = X + (int64_t)j * N;
    double s = 0.0;
    for (int64_t i = 0; i < N; i++) s += col[i] * d[i];
    out[j] = s;
  }
}

// SURVEY.md Appendix A.1
double linear_ll(const double* X, const double* Y, int64_t N, int M, const double* beta, double* dbeta, Work& w) {
  double acc = 0.0;
  double t1 = logpdf_normal_vec(0.0, 1.0, beta, M);
  double a1 = acc + t1;
  gemv_n(X, N, M, beta, w.t2.data());                                  // t2 = X * beta
  for (int64_t i = 0; i < N; i++) w.resid[i] = Y[i] - w.t2[i];        // resid = Y - t2
  double t3 = logpdf_normal_vec(0.0, 1.0, w.resid.data(), N);
  double a2 = a1 + t3;
  if (!dbeta) return a2;
  double da2 = 1.0, da1 = 0.0, dt3 = 0.0, dt1 = 0.0;
  da1 = da1 + da2;
  dt3 = dt3 + da2;
  const double t5 = 1.0 * 1.0;                                          // hoisted: 1.0 .* 1.0
  for (int64_t i = 0; i < N; i++) w.t4[i] = 0.0 - w.resid[i];
  for (int64_t i = 0; i < N; i++) w.t6[i] = w.t4[i] / t5;
  for (int64_t i = 0; i < N; i++) w.dresid[i] = 0.0 + w.t6[i] * dt3;
  for (int64_t i = 0; i < N; i++) w.dt2[i] = 0.0 + (-w.dresid[i]);
  gemv_t(X, N, M, w.dt2.data(), dbeta);                                // dbeta = dbeta + transpose(X) * dt2
  dt1 = dt1 + da1;
  for (int j = 0; j < M; j++) dbeta[j] = dbeta[j] + ((0.0 - beta[j]) / t5) * dt1;
  return a2;
}

// SURVEY.md Appendix A.2 ; returns -inf when an observation has zero probability ("give up eval")
double logistic_ll(const double* X, const double* Y, int64_t N, int M, const double* beta, double* dbeta, Work& w) {
  double t1 = logpdf_normal_vec(0.0, 1.0, beta, M);
  double a1 = 0.0 + t1;
  gemv_n(X, N, M, beta, w.t2.data());
  for (int64_t i = 0; i < N; i++) w.t3v[i] = std::exp(w.t2[i]);
  for (int64_t i = 0; i < N; i++) w.t4v[i] = 1.0 + w.t3v[i];
  for (int64_t i = 0; i < N; i++) w.prob[i] = 1.0 / w.t4v[i];
  double t5 = 0.0;
  for (int64_t i = 0; i < N; i++) {                                     // logpdfBernoulli, src/distribs.jl:12-21
    double p = w.prob[i];
    if (Y[i] == 0.0) { if (p == 1.0) return -INFINITY; t5 += std::log(1.0 - p); }
    else { if (p == 0.0) return -INFINITY; t5 += std::log(p); }
  }
  double a2 = a1 + t5;
  if (!dbeta) return a2;
  double dt5 = 1.0, dt1 = 1.0;
  for (int64_t i = 0; i < N; i++) w.dprob[i] = (1.0 / (w.prob[i] - (1.0 - Y[i]))) * dt5;
  for (int64_t i = 0; i < N; i++) w.dt4v[i] = (-1.0 / (w.t4v[i] * w.t4v[i])) * w.dprob[i];
  for (int64_t i = 0; i < N; i++) w.dt3v[i] = w.dt4v[i];
  for (int64_t i = 0; i < N; i++) w.dt2[i] = std::exp(w.t2[i]) * w.dt3v[i];
  gemv_t(X, N, M, w.dt2.data(), dbeta);
  for (int j = 0; j < M; j++) dbeta[j] = dbeta[j] + ((0.0 - beta[j]) / 1.0) * dt1;
  return a2;
}

typedef double (*ll_fn)(const double*, const double*, int64_t, int, const double*, double*, Work&);
}  // namespace

extern "C" {

// one evaluation: model 0 = linear, 1 = logistic; grad may be NULL
double cpu_ref_eval(int model, const double* X, const double* Y, int64_t N, int M, const double* beta, double* grad) {
  Work w(N);
  return (model == 0 ? linear_ll : logistic_ll)(X, Y, N, M, beta, grad, w);
}

// simpleHMC for `chains` independent chains, one per OpenMP thread at a time.  Returns the number of
// leapfrogs taken; final[chains][M] receives the last states, accept_rate[chains] the acceptance rates.
int64_t cpu_ref_hmc(int model, const double* X, const double* Y, int64_t N, int M, const double* init /*[chains][M]*/,
                    int chains, int64_t steps, int isteps, double stepsize, uint64_t seed, double* final_state,
                    double* accept_rate, int threads) {
  ll_fn ll = model == 0 ? linear_ll : logistic_ll;
  int nthreads = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (nthreads < 1) nthreads = 1;
  if (nthreads > chains) nthreads = chains;
  std::atomic<int> next(0);
  std::atomic<long long> total(0);
  auto worker = [&]() {
   for (;;) {
    int c = next.fetch_add(1);
    if (c >= chains) break;
    Work w(N);
    std::mt19937_64 gen(seed + 7919ull * (uint64_t)c);
    std::normal_distribution<double> randn(0.0, 1.0);
    std::uniform_real_distribution<double> rand01(0.0, 1.0);
    std::vector<double> b0(init + (int64_t)c * M, init + (int64_t)(c + 1) * M), g0(M), b(M), g(M), v(M);
    double l0 = ll(X, Y, N, M, b0.data(), g0.data(), w);   // calc!(state0, ll_func)
    int64_t acc = 0;
    for (int64_t i = 0; i < steps; i++) {
      double ss = 0.0;
      for (int j = 0; j < M; j++) { v[j] = randn(gen); ss += v[j] * v[j]; }
      double H0 = l0 - ss / 2.0;
      b = b0; g = g0;
      double l = l0, H = H0;
      for (int j = 1; j <= isteps && std::isfinite(l); j++) {   // leapFrog, src/samplers.jl:50-59
        for (int k = 0; k < M; k++) { v[k] += g[k] * stepsize / 2.0; b[k] += stepsize * v[k]; }
        l = ll(X, Y, N, M, b.data(), g.data(), w);
        if (!std::isfinite(l)) std::fill(g.begin(), g.end(), 0.0);
        double s2 = 0.0;
        for (int k = 0; k < M; k++) { v[k] += g[k] * stepsize / 2.0; s2 += v[k] * v[k]; }
        H = l - s2 / 2.0;
        total.fetch_add(1, std::memory_order_relaxed);
      }
      if (rand01(gen) < std::exp(H - H0)) { b0 = b; g0 = g; l0 = l; acc++; }
    }
    if (final_state) std::memcpy(final_state + (int64_t)c * M, b0.data(), sizeof(double) * M);
    if (accept_rate) accept_rate[c] = (double)acc / (double)steps;
   }
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; t++) pool.emplace_back(worker);
  for (auto& t : pool) t.join();
  return (int64_t)total.load();
}

int cpu_ref_max_threads(void) {
  int n = (int)std::thread::hardware_concurrency();
  return n < 1 ? 1 : n;
}
}
