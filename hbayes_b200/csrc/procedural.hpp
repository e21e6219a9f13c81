// procedural.hpp -- CPT values defined by a formula instead of a stored array, for tables too large to
// build on the host (BASELINE config C4: one 2^30-entry CPT, 8 GiB).  The same function runs on the host
// (small twins of the network, parity tests) and in the device kernel that fills a GPU's slice.
//
// CPT of child X (d states, first/slowest variable) with npar joint parent configurations:
//   u(k, pa) = 0.05 + 0.95 * uniform(hash(seed, k * npar + pa)),   value(x, pa) = u(x, pa) / sum_k u(k, pa)
// i.e. the U(0.05, 1)-then-normalise recipe of the synthetic generators, made stateless.
// Every operation is a single correctly rounded IEEE-754 binary64 operation in a fixed order, so host and
// device agree bit for bit.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define HB_HD __host__ __device__ __forceinline__
#else
#define HB_HD inline
#endif

namespace hb {

HB_HD uint64_t proc_hash(uint64_t seed, uint64_t idx) {  // SplitMix64 finaliser over seed + (idx + 1) * golden
  uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

HB_HD double proc_mul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  volatile double r = a * b;  // volatile: no contraction with the following add on the host
  return r;
#endif
}
HB_HD double proc_add(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  volatile double r = a + b;
  return r;
#endif
}
HB_HD double proc_div(double a, double b) {
#ifdef __CUDA_ARCH__
  return __ddiv_rn(a, b);
#else
  return a / b;
#endif
}

HB_HD double proc_u(uint64_t seed, uint64_t idx) {
  const double unif = (double)(proc_hash(seed, idx) >> 11) * (1.0 / 9007199254740992.0);
  return proc_add(0.05, proc_mul(0.95, unif));
}

// value at flat index `idx` of a CPT with `d` child states and `npar` parent configurations
HB_HD double proc_cpt_value(uint64_t seed, int d, uint64_t npar, uint64_t idx) {
  const uint64_t pa = idx % npar;
  double sum = 0.0;
  for (int k = 0; k < d; ++k) sum = proc_add(sum, proc_u(seed, (uint64_t)k * npar + pa));
  return proc_div(proc_u(seed, idx), sum);
}

}  // namespace hb
