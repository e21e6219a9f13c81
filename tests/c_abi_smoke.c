/* Plain-C user of include/hbayes_cuda.h: proves the header is C (no C++/torch types) and that a structure-only
 * handle answers the parity hooks without a GPU.  Built and run by tests/test_host_structure.py. */
#include <stdio.h>
#include <string.h>

#include "../include/hbayes_cuda.h"

int main(int argc, char** argv) {
  hb_net* net = NULL;
  hb_jt* jt = NULL;
  char* json = NULL;
  int32_t order[64], n = 0;
  if (argc < 2) return 2;
  if (hb_net_load_hugin(argv[1], &net) != HB_OK) {
    fprintf(stderr, "%s\n", hb_last_error());
    return 1;
  }
  if (hb_jt_compile(net, 0, NULL, HB_STRUCTURE_ONLY, &jt) != HB_OK) return 1;
  if (hb_jt_describe_json(jt, &json) != HB_OK || !strstr(json, "\"clusters\"")) return 1;
  printf("%s\n", json);
  hb_string_free(json);
  if (hb_ve_order(net, 1, order, &n) != HB_OK || n <= 0) return 1;
  {
    int8_t ev[5] = {-1, -1, -1, -1, -1};
    double out[10];
    /* no device behind this handle: a compute call must fail loudly, never fall back */
    if (hb_jt_posteriors_batch(jt, 1, ev, out, HB_HOST_PTRS) != HB_E_CUDA) return 1;
  }
  hb_jt_free(jt);
  hb_net_free(net);
  return 0;
}
