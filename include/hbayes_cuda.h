/* hbayes_cuda.h -- C ABI of libhbayes_cuda, the B200 engine behind hbayes' exact-inference hot path.
 *
 * The reference (alpheccar/hbayes, pure Haskell) has no FFI of its own; these entry points are what a
 * Haskell `foreign import ccall` module would bind to replace, for this path only, the functions cited
 * beside each declaration (paths relative to the hbayes source tree).  INTEGRATION.md shows that
 * binding.  Plain pointers and sizes only; no C++ or torch types; no callbacks.
 *
 * Conventions
 *   - every function returns an int status: 0 = ok, negative = error class below; hb_last_error()
 *     returns a thread-local message valid until the next call on the same thread.  Nothing throws or
 *     aborts across this boundary.
 *   - the caller owns every input and output buffer; the library copies what it keeps.
 *   - handles are opaque and released by hb_*_free (usable as Haskell ForeignPtr finalizers).
 *   - vertices are dense ids 0..n_vars-1 (the reference's Vertex ids); a CPT's scope is
 *     [child, parents in `ingoing` order] and its values are laid out with the FIRST variable slowest
 *     (Bayes/Network.hs:69-79, Bayes/Factor/PrivateCPT.hs:315-320).
 *   - evidence matrices are int8 [n_rows][n_vars], -1 = unobserved; a row stands for
 *     changeEvidence [v =: s | v <- ascending vertex id, observed].
 *   - zero-probability evidence yields NaN rows exactly as the reference does (1.0 / 0 * 0); not an error.
 *   - there is no CPU fallback: without a CUDA device every compile call fails with HB_E_CUDA.
 */
#ifndef HBAYES_CUDA_H
#define HBAYES_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_OK 0
#define HB_E_ARG (-1)         /* bad argument, inconsistent CPT, evidence row not matching */
#define HB_E_NO_CLUSTER (-2)  /* query variables not inside one cluster: Haskell `Nothing` (FactorElimination.hs:319) */
#define HB_E_CUDA (-3)
#define HB_E_OOM (-4)
#define HB_E_IO (-5)
#define HB_E_UNSUPPORTED (-6)
#define HB_E_NCCL (-7)

/* pass devices = NULL, n_devices = HB_STRUCTURE_ONLY to a compile call to get a handle that only answers the
 * describe/program_json parity hooks (host graph work); every compute call on it fails with HB_E_CUDA. */
#define HB_STRUCTURE_ONLY (-1)

#define HB_HOST_PTRS 0    /* evidence / out are host buffers: copies are part of the call */
#define HB_DEVICE_PTRS 1  /* evidence / out are device buffers on the handle's GPU */

/* hb_*_set_precision: EXACT (default) = every multiply and add rounded separately, in the reference's association order:
 * results are bit-identical to the reference's Double arithmetic.  CONTRACT = the kernels generated with NVRTC may
 * contract `acc + a * b` into one fused multiply-add (one rounding instead of two, half the fp64 instructions):
 * results agree with the reference to ~1e-15 relative (the north star's bar is 1e-12), not bit for bit.  The
 * precompiled kernels are always exact. */
#define HB_PRECISION_EXACT 0
#define HB_PRECISION_CONTRACT 1

typedef struct hb_net hb_net;
typedef struct hb_jt hb_jt;
typedef struct hb_ve hb_ve;

const char* hb_last_error(void);
void hb_string_free(char* s);
/* number of visible CUDA devices (0 when there is none); never fails */
int hb_device_count(void);

/* Page-locked host memory for evidence / result matrices.  The batch calls accept any host pointer; pageable buffers are
 * staged through a page-locked bounce ring inside the library, page-locked ones are moved by DMA directly (measured on C2,
 * 2^20 rows: 17 ms per call from page-locked buffers, 19-30 ms from pageable ones).  A Haskell caller allocates its matrices here (mallocForeignPtrBytes memory is pageable) and frees
 * them with hb_host_free as the ForeignPtr finalizer; hb_host_register page-locks a buffer the caller already owns for as
 * long as it stays allocated (unregister BEFORE freeing it). */
int hb_host_alloc(int64_t bytes, void** out);
void hb_host_free(void* p);
int hb_host_register(void* p, int64_t bytes);
int hb_host_unregister(void* p);

/* ---- network ------------------------------------------------------------------------------- */
/* runBN of a BNMonad (Bayes/BayesianNetwork.hs:211-212): CPT v has scope cpt_scope[cpt_off[v]..cpt_off[v+1])
 * (child first) and values cpt_vals[val_off[v]..val_off[v+1]). */
int hb_net_create(int32_t n_vars, const int32_t* dims, const int64_t* cpt_off, const int32_t* cpt_scope,
                  const int64_t* val_off, const double* cpt_vals, hb_net** out);
/* same, but CPT v with proc_seed[v] != 0 is PROCEDURAL: it has no stored values (val_off[v+1] == val_off[v]); entry
 * idx is u(idx) / sum_k u(k * npar + idx % npar), u = 0.05 + 0.95 * uniform(splitmix64(seed, idx)) -- the synthetic
 * generators' recipe made stateless (hbayes_b200/csrc/procedural.hpp), for tables too large to build on the host
 * (BASELINE config C4).  Each GPU generates only the slice it owns. */
int hb_net_create_procedural(int32_t n_vars, const int32_t* dims, const int64_t* cpt_off, const int32_t* cpt_scope,
                             const int64_t* val_off, const double* cpt_vals, const uint64_t* proc_seed, hb_net** out);
/* entries [offset, offset + count) of CPT v in reference layout, stored or procedural */
int hb_net_cpt_values(const hb_net* net, int32_t v, int64_t offset, int64_t count, double* out);
/* importBayesianGraph + runBN (Bayes/ImportExport/HuginNet.hs:160-195), dialect of SURVEY Appendix C */
int hb_net_load_hugin(const char* path, hb_net** out);
int hb_net_n_vars(const hb_net* net);
int hb_net_dims(const hb_net* net, int32_t* dims_out /* n_vars */);
/* {"names","dims","scopes","values"}; free with hb_string_free */
int hb_net_describe_json(const hb_net* net, char** out_json);
void hb_net_free(hb_net* net);

/* minDegreeOrder (metric 0) / minFillOrder (metric 1) (Bayes/VariableElimination.hs:210-236);
 * order_out has room for n_vars ids, *n_out receives how many were written */
int hb_ve_order(const hb_net* net, int32_t metric, int32_t* order_out, int32_t* n_out);
/* degreeOrder (Bayes/VariableElimination.hs:186-207): the width an elimination order induces (largest neighbour count
 * at elimination time, with fill-in) */
int hb_ve_degree_order(const hb_net* net, const int32_t* order, int32_t n, int32_t* width);

/* ---- junction tree ------------------------------------------------------------------------- */
/* createJunctionTree cmp g (Bayes/FactorElimination.hs:298-307).  cmp: 0 = nodeComparisonForTriangulation
 * (weightedEdges, :71-111), 1 = numberOfAddedEdges.  devices: distinct CUDA ordinals.  The handle lives on devices[0]
 * (stateful single-row calls, device-pointer batches); with n_devices > 1 every further GPU gets a replica of the
 * compiled tree and the HOST-buffer hb_jt_posteriors_batch shards its rows over all of them -- contiguous slices, one
 * host thread and one GPU each, no communication (evidence rows are independent).  The tree is calibrated without
 * evidence. */
int hb_jt_compile(const hb_net* net, int32_t cmp, const int32_t* devices, int32_t n_devices, hb_jt** out);
/* same, with the vertex elimination order of the triangulation given explicitly (any other `cmp` closure
 * is evaluated on the Haskell side: the metric is static) */
int hb_jt_compile_with_order(const hb_net* net, const int32_t* elim_order, const int32_t* devices, int32_t n_devices,
                             hb_jt** out);
/* createJunctionTree for ONE slice of a split clique (see hb_jt_set_split): the tree is built, the split is set, and
 * nothing is calibrated (the unsplit tables may not fit one GPU). */
int hb_jt_compile_split(const hb_net* net, int32_t cmp, const int32_t* devices, int32_t n_devices, int32_t n_fixed,
                        const int32_t* vars, const int32_t* states, hb_jt** out);
/* Split clique with the exchange step INSIDE the library (BASELINE config C4): one process per GPU, `world`
 * of them, each owning the slice vars[i] = states[i] of the tables that mention the split variables (the state
 * combinations are dealt one per rank: world = prod dims(vars)).  rank 0 obtains a 128-byte id with hb_nccl_unique_id and
 * hands it to the others by whatever channel the host program has (MPI, torch.distributed, a file); the compile call is
 * collective (ncclCommInitRank).  Programs compiled on the handle sum over the split variables where the reference's
 * elimination does: every message that leaves the split clique and every posterior inside it is computed from the
 * rank's slice and summed over the ranks with ncclAllReduce(ncclDouble, ncclSum) on the handle's stream -- |separator|
 * (or |query|) doubles per row; everything downstream of such a message is no longer conditioned on the slice and is
 * computed once, not once per slice.  Batch calls are collective too (same rows on every rank, split variables
 * unobserved = -1) and return complete, normalised posteriors on every rank.  Agreement with the unsplit engine:
 * ~1e-15 relative (the sum over the split variables is reassociated), tests use 1e-12.  HB_E_NCCL: libnccl missing, or
 * an NCCL call failed.  HB_E_UNSUPPORTED: a separator holds some split variables and eliminates others.
 * device = HB_STRUCTURE_ONLY gives a handle that only answers hb_jt_program_json (no GPU, no communicator). */
int hb_nccl_unique_id(void* out128);
int hb_jt_compile_split_nccl(const hb_net* net, int32_t cmp, int32_t device, int32_t n_fixed, const int32_t* vars,
                             const int32_t* states, int32_t rank, int32_t world, const void* unique_id128, hb_jt** out);
/* Persistence (the counterpart of the Binary save/load of a junction tree, Bayes/ImportExport.hs:36-46): the network
 * and the tree structure in one binary file.  Loading skips moralisation / triangulation / spanning tree; schedules
 * are recompiled on demand (milliseconds) and the tree is re-calibrated on the GPU. */
int hb_jt_save(const hb_jt* jt, const char* path);
int hb_jt_load(const char* path, const int32_t* devices, int32_t n_devices, hb_jt** out);
/* parity hook: {"elim_order","clusters","root","parent_sep","sep_parent","sep_child","sep_vars","children","factors"} */
int hb_jt_describe_json(const hb_jt* jt, char** out_json);
/* parity hook: the operation trace (every message and every single-variable posterior, with product scopes)
 * and the physical step list compiled for changeEvidence [vars[0], vars[1], ...] */
int hb_jt_program_json(hb_jt* jt, int32_t n, const int32_t* vars, char** out_json);
/* Inspection hook: the CUDA source of the kernels specialised per schedule step for this observed list (what the
 * library hands to NVRTC on the first large batch; csrc/jit.cpp).  Free with hb_string_free. */
int hb_jt_jit_source(hb_jt* jt, int32_t n_observed, const int32_t* observed, char** out_source);
/* Inspection hook: the CUDA source of the whole-schedule on-chip kernel `hb_tree` for this observed list (csrc/jit.cpp,
 * SURVEY 2.2 K4): changeEvidence + collect + distribute + every posterior for a tile of rows in one launch, per-row
 * tables in shared memory.  HB_E_UNSUPPORTED when the program does not qualify (tables too large for shared memory). */
int hb_jt_tree_source(hb_jt* jt, int32_t n_observed, const int32_t* observed, char** out_source);
/* Builds the specialised kernels of the program for this observed list NOW (upload + NVRTC compilation, seconds) instead
 * of inside the first large batch call, e.g. right after hb_jt_compile in a service.  *n_kernels (optional) = steps
 * covered; 0 when NVRTC is unavailable or HB_JIT=0 -- not an error, the precompiled kernels are used. */
int hb_jt_specialise(hb_jt* jt, int32_t n_observed, const int32_t* observed, int32_t* n_kernels);
/* how many schedule steps of the last batch call's program run on a kernel specialised for them (0: none yet, NVRTC
 * unavailable, or HB_JIT=0) */
int hb_jt_specialised_kernels(const hb_jt* jt, int32_t* n);
/* What the last batch call's (or hb_jt_specialise's) program runs on: [0] 1 if the program qualifies for the on-chip kernel,
 * [1] evidence rows per CTA tile, [2] threads per CTA, [3] dependency levels, [4] shared-memory arena entries per row,
 * [5] shared-memory bytes per CTA, [6] fp64 multiplies per row, [7] fp64 adds per row, [8] / [9] shared-memory table
 * entries read / written per row by the generated kernel, [10] 1 if the on-chip kernel is installed and used,
 * [11] origin of the installed generated kernels: -1 none, 0 this process (incl. hb_jt_load), 1 the disk cache,
 * 2 compiled by NVRTC now. */
int hb_jt_kernel_info(const hb_jt* jt, int64_t* out12);
/* Tuning hook (HB_TREE_PROFILE=1 in the environment when the on-chip kernel is generated): SM cycles per phase of the
 * on-chip kernel since the previous call, summed over all tiles: out64[0] evidence preparation, out64[1 + l] dependency
 * level l, out64[1 + levels] normalise + write-out; zeros when the instrumentation is off. */
int hb_jt_tree_profile(hb_jt* jt, int64_t* out64);
/* changeEvidence (Bayes/FactorElimination/JTree.hs:454-466); list order is preserved */
int hb_jt_set_evidence(hb_jt* jt, int32_t n, const int32_t* vars, const int32_t* states);
/* posterior jt q_vars (Bayes/FactorElimination.hs:314-327).  out_scope receives the variable order of the
 * result (n_q ids), out its values (first variable of out_scope slowest); *n_written the number of values. */
int hb_jt_posterior(hb_jt* jt, int32_t n_q, const int32_t* q_vars, int32_t* out_scope, double* out, int64_t out_len,
                    int64_t* n_written);
/* one junction-tree query per row: changeEvidence row, then posterior [v] for every v.
 * out: double [n_rows][sum_v dims[v]], variables in ascending id.  Rows of one call may have different
 * observed sets with HB_HOST_PTRS (grouped internally); with HB_DEVICE_PTRS all rows must share one. */
int hb_jt_posteriors_batch(hb_jt* jt, int64_t n_rows, const int8_t* evidence, double* out, int32_t flags);
/* GPUs the handle spreads host-buffer batches over (the n_devices of its compile call) */
int hb_jt_n_devices(const hb_jt* jt);
/* Split clique (BASELINE config C4): fixes vars[i] to states[i] for every later batch call on this
 * handle -- the slice of the oversized table this GPU owns.  Potentials mentioning the variables are stored
 * restricted to those states (1/prod dims of the memory), evidence rows must carry the same states, and batch
 * results are left UNNORMALISED: the caller sums the result matrices of all slices (ncclAllReduce(sum) over the
 * ranks; the sum over the split variable is the exchange step) and calls hb_jt_normalise_batch.  Results agree
 * with the unsplit engine to ~1e-15 relative, not bit for bit (the sum over the split variable is reassociated). */
int hb_jt_set_split(hb_jt* jt, int32_t n, const int32_t* vars, const int32_t* states);
/* factorNorm + factorDivide of a summed, unnormalised result matrix [n_rows][sum dims] in place */
int hb_jt_normalise_batch(hb_jt* jt, int64_t n_rows, double* out, int32_t flags);
/* Batched `posterior jt q` for a LIST of queries (each a set of variables inside one cluster, e.g. a CPT family and
 * its parents: the E-step of Bayes/EM.hs:47-66 for all samples at once).  Query i = query_vars[query_off[i]..
 * query_off[i+1]).  out: [n_rows][sum_i prod dims(query i)] (hb_jt_queries_width); each block is laid out in the
 * result variable order of the reference, written to out_scopes (same shape as query_vars) when not NULL.
 * Mixed observed sets need one call per observed set (the result order may depend on it). */
int hb_jt_queries_batch(hb_jt* jt, int32_t n_queries, const int32_t* query_off, const int32_t* query_vars, int64_t n_rows,
                        const int8_t* evidence, double* out, int32_t flags, int32_t* out_scopes);
int hb_jt_queries_width(hb_jt* jt, int32_t n_queries, const int32_t* query_off, const int32_t* query_vars, int64_t* width);
/* changeFactor f jt (Bayes/FactorElimination/JTree.hs:107-115, Bayes/Factor.hs:66-81): replaces the values of the
 * CPT whose variable LIST equals scope (same variables in the same order: isUsingSameVariablesAs is list equality,
 * Bayes/Factor/PrivateCPT.hs:135-136) and recalibrates under the current evidence.  values: prod dims(scope) doubles,
 * first variable slowest.  *replaced = 1 if a CPT matched, 0 if none did (the reference then returns the tree
 * unchanged; so does this call).  The compiled schedules are kept: on the device this is a potential upload. */
int hb_jt_change_factor(hb_jt* jt, int32_t n_scope, const int32_t* scope, const double* values, int32_t* replaced);
/* learnEM samples g (Bayes/EM.hs:36-66), one expectation/maximisation step on the handle's network: per sample
 * (row) changeEvidence + posterior of every CPT family and of its parents, summed sample after sample (cptSum) and
 * divided (cptDivide, divide _ 0 = 0).  A row means changeEvidence [v =: s | v ascending, observed], as in the other
 * batch calls; rows may have different observed sets.  out_values: sum_v cpt entries
 * doubles, vertex after vertex, each table laid out as the reference returns it -- variable order written to
 * out_scopes (sum_v |scope v| ids): the parents in the order of their joint posterior, then the child, NOT the
 * CPT's own [child, parents] order (cptDivide's union [b, a], Bayes/Factor/CPT.hs:159-160).  evidence is a host or
 * device pointer according to flags; out_values and out_scopes are host pointers. */
int hb_jt_em_step(hb_jt* jt, int64_t n_rows, const int8_t* evidence, double* out_values, int32_t* out_scopes, int32_t flags);
int hb_jt_set_precision(hb_jt* jt, int32_t mode);
/* CUDA stream (a cudaStream_t) the handle launches on; NULL = the default stream */
int hb_jt_set_stream(hb_jt* jt, void* cuda_stream);
/* bytes of device workspace the row arena may use (0 = default) */
int hb_jt_set_workspace(hb_jt* jt, int64_t bytes);
/* counters of the last batch call: [0] kernels launched, [1] row tiles, [2] rows per tile, [3] arena bytes,
 * [4] per-row arena entries, [5] product entries per row, [6] per-row table entries read per row,
 * [7] per-row table entries written per row */
int hb_jt_last_stats(const hb_jt* jt, int64_t* out8);
/* CUDA-event timing of the kernel groups on the handle's stream (off by default).  hb_jt_last_timing
 * synchronises the stream; out4 = milliseconds of the last batch call summed over its row tiles:
 * [0] evidence preparation kernel, [1] all product/sum-out kernels, [2] normalise/expand kernel, [3] whole pass */
int hb_jt_set_timing(hb_jt* jt, int32_t on);
int hb_jt_last_timing(hb_jt* jt, double* out4);
void hb_jt_free(hb_jt* jt);

/* ---- factor algebra on free-standing factors ------------------------------------------------- */
/* factorProduct (Bayes/Factor/CPT.hs:138 = _factorProduct, PrivateCPT.hs:215-238) of k factors and, when
 * sum_var >= 0, factorProjectOut [sum_var] of the product (CPT.hs:140-141): together marginalizeOneVariable
 * (Bayes/VariableElimination/Buckets.hs:105-111), the unit the GPU kernels fuse.  Factor i has scope
 * scopes[scope_off[i]..scope_off[i+1]) (vertex ids, reference order; empty = Scalar) and values
 * vals[val_off[i]..val_off[i+1]) with the first variable slowest; dims[v] is the dimension of vertex v.
 * out_scope receives the result's variable order (foldr1 union of the scopes, minus sum_var), out its values.
 * One call = one compile + one launch: meant for tests and for checking the algebra, not for throughput. */
int hb_factor_marginalize(int32_t n_vars, const int32_t* dims, int32_t k, const int32_t* scope_off, const int32_t* scopes,
                          const int64_t* val_off, const double* vals, int32_t sum_var, int32_t device, int32_t* out_scope,
                          int32_t* n_scope, double* out, int64_t out_len, int64_t* n_written);

/* ---- variable elimination ------------------------------------------------------------------ */
/* posteriorMarginal g p r (Bayes/VariableElimination.hs:116-130): p = variables to eliminate, in order,
 * r = remaining (query) variables.  priorMarginal is the same with no evidence. */
int hb_ve_compile(const hb_net* net, const int32_t* p, int32_t np, const int32_t* r, int32_t nr,
                  const int32_t* devices, int32_t n_devices, hb_ve** out);
/* result variable order (depends on the observed set: rule A.5); scope_out has room for nr ids */
int hb_ve_result_scope(hb_ve* ve, int32_t n_obs, const int32_t* observed, int32_t* scope_out, int32_t* n_out);
int hb_ve_program_json(hb_ve* ve, int32_t n_obs, const int32_t* observed, char** out_json);
/* assignment in the caller's order (posteriorMarginal's [DVI] argument); out has prod_r dims[r] doubles */
int hb_ve_posterior(hb_ve* ve, int32_t n, const int32_t* vars, const int32_t* states, int32_t* out_scope, double* out,
                    int64_t out_len, int64_t* n_written);
/* one posteriorMarginal per row (assignment = observed entries in ascending vertex id); out: [n_rows][prod_r dims[r]] */
int hb_ve_posterior_batch(hb_ve* ve, int64_t n_rows, const int8_t* evidence, double* out, int32_t flags);
/* mpe g p r assignment (Bayes/VariableElimination.hs:79-114): the variables of p are summed out (the observed
 * ones belong there), those of r maximised in turn with MAXCPT semantics (Bayes/Factor/MaxCPT.hs:25-71: the LAST
 * maximal state wins a tie).  p and r must cover every variable.  The handle answers only the hb_ve_mpe* calls. */
int hb_ve_compile_mpe(const hb_net* net, const int32_t* p, int32_t np, const int32_t* r, int32_t nr, const int32_t* devices,
                      int32_t n_devices, hb_ve** out);
/* One query, assignment in the caller's order.  out_states: nr int8, the maximising state of r[i] (-1: the variable
 * got no instantiation); *out_value = P(that instantiation, evidence), unnormalised, bit-identical to the
 * reference's MAXCPT scalar; out_order / *n_order (optional): the variables in the order of the reference's
 * instantiation list [DVI] (mpeInstantiations, MaxCPT.hs:73-75). */
int hb_ve_mpe(hb_ve* ve, int32_t n, const int32_t* vars, const int32_t* states, int8_t* out_states, double* out_value,
              int32_t* out_order, int32_t* n_order);
/* one mpe per evidence row (assignment = observed entries in ascending vertex id).  out_states: int8 [n_rows][nr];
 * out_value: double [n_rows] or NULL.  Mixed observed sets are grouped with HB_HOST_PTRS. */
int hb_ve_mpe_batch(hb_ve* ve, int64_t n_rows, const int8_t* evidence, int8_t* out_states, double* out_value, int32_t flags);
/* the instantiation-list variable order for a given observed list (structural) */
int hb_ve_mpe_order(hb_ve* ve, int32_t n_obs, const int32_t* observed, int32_t* order_out, int32_t* n_out);
/* changeFactor f g on the network behind a variable-elimination handle (Bayes/Factor.hs:66-81); see hb_jt_change_factor */
int hb_ve_change_factor(hb_ve* ve, int32_t n_scope, const int32_t* scope, const double* values, int32_t* replaced);
/* builds the kernels generated for this observed list now (see hb_jt_specialise) */
int hb_ve_specialise(hb_ve* ve, int32_t n_observed, const int32_t* observed, int32_t* n_kernels);
int hb_ve_set_precision(hb_ve* ve, int32_t mode);
int hb_ve_set_stream(hb_ve* ve, void* cuda_stream);
int hb_ve_set_workspace(hb_ve* ve, int64_t bytes);
int hb_ve_last_stats(const hb_ve* ve, int64_t* out8);
int hb_ve_set_timing(hb_ve* ve, int32_t on);
int hb_ve_last_timing(hb_ve* ve, double* out4);
void hb_ve_free(hb_ve* ve);

#ifdef __cplusplus
}
#endif
#endif /* HBAYES_CUDA_H */
