// jit.hpp -- per-step kernel specialisation (NVRTC), see jit.cpp.
#pragma once
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
// The CUDA source of the specialised kernels `hb_step_<si>` of the given steps (for tests and inspection).
std::string jit_source(const Program& P, const std::vector<int>& steps);
// Compiles and loads the kernels on the current device.  fn_per_step[si] = function handle or nullptr.  Returns nullptr
// (with a reason in *why) when NVRTC / the driver API are not available or the compilation fails: not an error.
JitModule* jit_build(const Program& P, const std::vector<int>& steps, std::vector<void*>& fn_per_step, std::string* why);
void jit_destroy(JitModule* m);
// kernel parameters: (const double* arena, double* arena_out (the same arena), const double* pots, const double* shared,
//                     const int* rowoff, int nr, long long RT)
bool jit_launch(void* fn, unsigned gx, unsigned gy, unsigned bx, unsigned by, void** params, void* stream);

}  // namespace hb
