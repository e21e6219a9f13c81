/* examples/cancer_query.c -- the reference's tutorial query (Bayes/Examples/Tutorial.hs, cancer network) from plain C:
 * load the Hugin file, build + calibrate the junction tree on GPU 0, print the prior of every variable, set evidence,
 * print the posteriors, then run a small batch.
 *
 *   gcc -std=c99 examples/cancer_query.c -Iinclude -Lhbayes_b200 -lhbayes_cuda -Wl,-rpath,$PWD/hbayes_b200 -o cancer_query
 *   ./cancer_query tests/golden/cancer.net
 */
#include <stdio.h>
#include <stdlib.h>

#include "../include/hbayes_cuda.h"

#define CHECK(call)                                                        \
  do {                                                                     \
    int rc_ = (call);                                                      \
    if (rc_ != HB_OK) {                                                    \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, hb_last_error()); \
      return 1;                                                            \
    }                                                                      \
  } while (0)

static int print_all(hb_jt* jt, int n, const int32_t* dims, const char* title) {
  printf("%s\n", title);
  for (int32_t v = 0; v < n; ++v) {
    double p[16];
    int32_t scope[1];
    int64_t k = 0;
    CHECK(hb_jt_posterior(jt, 1, &v, scope, p, 16, &k));
    printf("  var %d:", v);
    for (int i = 0; i < dims[v]; ++i) printf(" %.6f", p[i]);
    printf("\n");
  }
  return 0;
}

int main(int argc, char** argv) {
  hb_net* net = NULL;
  hb_jt* jt = NULL;
  int32_t dev = 0, dims[64];
  if (argc < 2) {
    fprintf(stderr, "usage: %s file.net\n", argv[0]);
    return 2;
  }
  CHECK(hb_net_load_hugin(argv[1], &net));
  const int n = hb_net_n_vars(net);
  CHECK(hb_net_dims(net, dims));
  CHECK(hb_jt_compile(net, 0 /* nodeComparisonForTriangulation */, &dev, 1, &jt));
  if (print_all(jt, n, dims, "priors")) return 1;
  {
    const int32_t vars[1] = {0}, states[1] = {0}; /* changeEvidence [D =: Present] */
    CHECK(hb_jt_set_evidence(jt, 1, vars, states));
  }
  if (print_all(jt, n, dims, "posteriors given variable 0 in state 0")) return 1;
  {
    /* a batch: row r observes variable 0 in state r % 2 */
    enum { ROWS = 4 };
    int8_t ev[ROWS * 64];
    double out[ROWS * 64];
    int width = 0;
    for (int v = 0; v < n; ++v) width += dims[v];
    for (int r = 0; r < ROWS; ++r)
      for (int v = 0; v < n; ++v) ev[r * n + v] = (int8_t)(v == 0 ? r % 2 : -1);
    CHECK(hb_jt_posteriors_batch(jt, ROWS, ev, out, HB_HOST_PTRS));
    printf("batch of %d rows, %d doubles each; row 1:", ROWS, width);
    for (int i = 0; i < width; ++i) printf(" %.4f", out[width + i]);
    printf("\n");
  }
  hb_jt_free(jt);
  hb_net_free(net);
  return 0;
}
