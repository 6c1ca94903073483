// Counter-based Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11).
// The reference draws from Julia's global MersenneTwister (randn(nparams) then rand(), src/samplers.jl:151,162);
// with many chains in lockstep each (chain, step, purpose, index) gets its own counter instead, so a chain's
// stream does not depend on how chains are sharded over GPUs.
#pragma once
#include <stdint.h>

struct smc_u4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ void smc_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
  uint64_t p = (uint64_t)a * (uint64_t)b;
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
}

__host__ __device__ __forceinline__ smc_u4 smc_philox4x32_10(smc_u4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0, lo0, hi1, lo1;
    smc_mulhilo(0xD2511F53u, c.x, hi0, lo0);
    smc_mulhilo(0xCD9E8D57u, c.z, hi1, lo1);
    smc_u4 n;
    n.x = hi1 ^ c.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ c.w ^ k1;
    n.w = lo0;
    c = n;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// purpose streams
#define SMC_STREAM_NORMAL 0u   // momentum / proposal normals
#define SMC_STREAM_UNIFORM 1u  // accept / slice / direction uniforms

__host__ __device__ __forceinline__ smc_u4 smc_rng_block(uint64_t seed, uint64_t chain, uint64_t step, uint32_t stream,
                                                        uint32_t block) {
  smc_u4 c;
  c.x = block;
  c.y = stream | ((uint32_t)(chain >> 32) << 8);
  c.z = (uint32_t)step;
  c.w = (uint32_t)chain;
  return smc_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32));
}

// 53-bit uniform strictly inside (0,1)
__host__ __device__ __forceinline__ double smc_u53(uint32_t hi, uint32_t lo) {
  uint64_t x = ((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6);
  return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
}

// k-th uniform of (chain, step)
__device__ __forceinline__ double smc_uniform(uint64_t seed, uint64_t chain, uint64_t step, uint32_t k) {
  smc_u4 r = smc_rng_block(seed, chain, step, SMC_STREAM_UNIFORM, k >> 1);
  return (k & 1) ? smc_u53(r.z, r.w) : smc_u53(r.x, r.y);
}

// two standard normals (Box-Muller) from block b of (chain, step): normals 2b and 2b+1
__device__ __forceinline__ void smc_normal_pair(uint64_t seed, uint64_t chain, uint64_t step, uint32_t b, double& z0,
                                                double& z1) {
  smc_u4 r = smc_rng_block(seed, chain, step, SMC_STREAM_NORMAL, b);
  double u1 = smc_u53(r.x, r.y), u2 = smc_u53(r.z, r.w);
  double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  z0 = rad * c;
  z1 = rad * s;
}
