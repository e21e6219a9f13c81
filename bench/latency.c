/* bench/latency.c -- BASELINE config C1 from plain C: latency of one junction-tree query on cancer.net through the
 * stateful C ABI, i.e. what a Haskell `changeEvidence` + `posterior` for every variable costs behind the FFI
 * (Bayes/FactorElimination/JTree.hs:454-466, Bayes/FactorElimination.hs:314-327), without the Python binding in the way.
 *
 *   gcc -std=c99 -O2 bench/latency.c -Iinclude -Lhbayes_b200 -lhbayes_cuda -Wl,-rpath,$PWD/hbayes_b200 -o latency
 *   ./latency tests/golden/cancer.net [iterations]        -> one JSON line
 */
#define _POSIX_C_SOURCE 199309L
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "../include/hbayes_cuda.h"

#define CHECK(call)                                                        \
  do {                                                                     \
    int rc_ = (call);                                                      \
    if (rc_ != HB_OK) {                                                    \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, hb_last_error()); \
      return 1;                                                            \
    }                                                                      \
  } while (0)

static double now_us(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec * 1e6 + t.tv_nsec / 1e3;
}

int main(int argc, char** argv) {
  hb_net* net = NULL;
  hb_jt* jt = NULL;
  int32_t dev = 0, dims[256];
  if (argc < 2) {
    fprintf(stderr, "usage: %s file.net [iterations]\n", argv[0]);
    return 2;
  }
  const int iters = argc > 2 ? atoi(argv[2]) : 2000;
  CHECK(hb_net_load_hugin(argv[1], &net));
  const int n = hb_net_n_vars(net);
  if (n > 256) return 2;
  CHECK(hb_net_dims(net, dims));
  CHECK(hb_jt_compile(net, 0, &dev, 1, &jt));
  /* the three evidence sets of the reference's golden cases: none, variable 0 in state 0, variable 0 in state 1 */
  const int32_t var0[1] = {0}, s0[1] = {0}, s1[1] = {1};
  int32_t kernels = 0;
  CHECK(hb_jt_specialise(jt, 0, NULL, &kernels)); /* generated kernels built here, not inside the timed loop */
  CHECK(hb_jt_specialise(jt, 1, var0, &kernels));
  double sum = 0.0;
  for (int pass = 0; pass < 2; ++pass) { /* pass 0 warms up */
    const int m = pass ? iters : 50;
    const double t0 = now_us();
    for (int it = 0; it < m; ++it) {
      const int which = it % 3;
      CHECK(hb_jt_set_evidence(jt, which ? 1 : 0, var0, which == 2 ? s1 : s0));
      for (int32_t v = 0; v < n; ++v) {
        double p[16];
        int32_t scope[1];
        int64_t k = 0;
        CHECK(hb_jt_posterior(jt, 1, &v, scope, p, 16, &k));
        sum += p[0];
      }
    }
    const double us = (now_us() - t0) / m;
    if (pass) {
      int64_t info[12];
      CHECK(hb_jt_kernel_info(jt, info));
      printf("{\"c_abi_single_query_latency_us\": %.2f, \"iterations\": %d, \"on_chip_kernel\": %d, \"checksum\": %.6f}\n", us, m, (int)info[10], sum);
    }
  }
  hb_jt_free(jt);
  hb_net_free(net);
  return 0;
}
