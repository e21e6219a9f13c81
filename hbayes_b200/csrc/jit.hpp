// jit.hpp -- per-step kernel specialisation (NVRTC), see jit.cpp.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "model.hpp"

namespace hb {

struct JitModule;

// May this step be turned into a specialised kernel?  (a per-row sum step without scalars, small enough to unroll)
bool jit_eligible(const Program& P, const Step& st);
// Size of the generated body of a step (chain evaluations of one group; the outermost jit_loop_levels(st) levels stay
// run-time loops so that the unrolled rest stays small).
int jit_loop_levels(const Step& st);
int64_t jit_unrolled_ops(const Step& st);
// Loads per row the specialised kernel of the step would issue, and what the templated kernels issue: a step is only
// worth a specialised kernel when the first is clearly smaller.
void jit_load_counts(const Program& P, const Step& st, int64_t* specialised, int64_t* templated);
// How the kernel of a step walks its groups (part of the generated source, so the launch has to agree with it).
struct JitShape {
  bool rereads = false;     // the step reads its per-row inputs at least three times over
  bool interleave = false;  // groups dealt round robin to the lanes instead of one contiguous strip per lane
  int unroll = 0;           // groups of the loop evaluated together (0: left to the compiler)
  // Register block: the output variables whose states one thread evaluates in lock step (a group = all their state
  // combinations, U entries).  Default: the step's blocking variable w alone.  A re-reading step gets the set of output
  // variables that leaves the fewest loads per output entry (`custom`): an outer-product block over variables of DIFFERENT
  // inputs uses every loaded value several times.  The arithmetic of an entry does not depend on the block.
  std::vector<int> block;
  bool custom = false;
  int V = 2;                // evidence rows per thread (2: adjacent rows, 16-byte accesses)
  int U = 1;                // entries per group
  int64_t groups = 1;       // n_out / U
  int64_t inner = 1;        // > 1: a lane's strip is cut in units of `inner` consecutive groups (groups / inner units in all)
  int n_loop = 0;           // outermost levels that stay run-time loops in the generated body
  int64_t body_budget = 600;  // ... so that the unrolled rest stays below this many chain evaluations
  int64_t ops = 0;          // chain evaluations in the unrolled part of a group's body (-1: does not fit)
  int min_blocks = 2;       // __launch_bounds__(256, min_blocks)
  bool eligible = false;    // the body fits the limits
};
JitShape jit_step_shape(const Program& P, const Step& st);
// May the generated kernels leave out the `0 +` every reference sum starts with?  Yes when no potential is negative, -0.0
// or NaN: then 0 + x == x bit for bit for every intermediate value.  (HB_JIT_ZERO_ADD=1 keeps the adds.)
bool jit_zero_add_elidable(const Program& P);
// The CUDA source of the specialised kernels `hb_step_<si>` of the given steps (for tests and inspection).
std::string jit_source(const Program& P, const std::vector<int>& steps);
void jit_destroy(JitModule* m);
// kernel parameters: (const double* arena, double* arena_out (the same arena), const double* pots, const double* shared,
//                     const int* rowoff, int nr, long long RT)
bool jit_launch(void* fn, unsigned gx, unsigned gy, unsigned bx, unsigned by, void** params, void* stream);

// ---- cubin cache: this process (jit_registry_*; hb_jt_save / hb_jt_load carry it across processes), a directory on disk
// (HB_CACHE_DIR, default ~/.cache/hbayes_b200; HB_JIT_CACHE=0 disables it), then NVRTC ----
std::string jit_cache_key(const std::string& src, bool fma);
void jit_registry_put(const std::string& key, const std::string& cubin);
std::shared_ptr<const std::string> jit_registry_get(const std::string& key);
// Source -> cubin.  Host-only work (no CUDA context): safe on a background thread.  *origin: 0 = this process' registry,
// 1 = disk cache, 2 = compiled now.  fma = true compiles with --fmad=true (the 1e-12 mode; never bit-exact).
std::shared_ptr<const std::string> jit_compile(const std::string& src, bool fma, std::string* why, std::string* key_out, int* origin);
JitModule* jit_load(const std::string& cubin, const std::string& key, std::string* why);  // on the current device
void* jit_function(JitModule* m, const std::string& name);
const std::string& jit_module_key(const JitModule* m);

// ---- whole-schedule on-chip kernel `hb_tree`: one launch per batch, per-row tables in shared memory ----
struct TreePlan {
  int R = 0;                   // evidence rows per CTA tile (even: a thread works on two adjacent rows)
  int T = 256;                 // threads per CTA
  int min_blocks = 1;          // CTAs per SM the kernel is bounded for
  int n_levels = 0;            // dependency levels (one __syncthreads each)
  std::vector<int> level;      // per program step: level of a per-row product step, -1 otherwise
  std::vector<int64_t> off;    // per table: entry offset of a per-row table in the shared-memory arena, -1 otherwise
  int64_t arena_entries = 0;   // per row, with reuse of dead tables from one level to the next
  int64_t invariant_entries = 0;  // potential pool + shared arena entries staged in shared memory (0: read through L1/L2)
  int64_t stage_pitch = 0;     // doubles per row of the output staging tile (0: results are written straight to HBM)
  size_t smem_bytes = 0;
  int64_t ops_mul = 0, ops_add = 0;  // fp64 multiplies / adds per row (what the reference does, no more)
};
// false (reason in *why) when the program does not qualify: maximisation steps, tables too large for shared memory, ...
bool tree_plan(const Program& P, TreePlan& plan, std::string* why);
// loads_per_row / stores_per_row (optional): shared-memory table entries one evidence row reads / writes in the kernel
std::string tree_source(const Program& P, const TreePlan& plan, int64_t* loads_per_row = nullptr, int64_t* stores_per_row = nullptr);
// kernel parameters of hb_tree: (const double* pots, const double* shared, const int8* evidence, double* out,
//                                long long n_rows, int* flag, int normalise)
bool jit_launch_smem(void* fn, unsigned grid, unsigned block, unsigned smem_bytes, void** params, void* stream);
bool jit_set_dynamic_smem(void* fn, int bytes);
int jit_occupancy(void* fn, int block, size_t smem_bytes);  // resident CTAs per SM, 0 = unknown

}  // namespace hb
