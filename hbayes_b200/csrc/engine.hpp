// engine.hpp -- device side of libhbayes_cuda: a compiled Program resident on one GPU.
#pragma once
#include <cstdint>

#include "model.hpp"

namespace hb {

struct Executor;

struct RunStats {          // filled by executor_run (host-side counts, no device sync)
  int64_t launches = 0;    // kernels launched
  int64_t tiles = 0;       // row tiles
  int64_t rows_per_tile = 0;
  int64_t arena_bytes = 0;
};

// Uploads potentials, index maps and step descriptors to `device` and evaluates the batch-invariant
// (shared) steps once.  `workspace_bytes` caps the per-row arena (0 = default: 60 % of free memory, at most 24 GiB).
Executor* executor_create(const Program& P, int device, int64_t workspace_bytes);
void executor_destroy(Executor* ex);

// changeFactor (Bayes/FactorElimination/JTree.hs:107-115): the network's CPT values changed, the schedule did
// not.  Re-uploads the potential pool from P.net and re-evaluates the batch-invariant steps; synchronises
// the device.  Captured graphs stay valid (same addresses).
void executor_refresh_potentials(Executor* ex);

// One pass of the program over n_rows evidence rows.  d_evidence: int8 [n_rows][n_vars] (-1 = unobserved),
// d_out: double [n_rows][out_width]; both DEVICE pointers.  Asynchronous on `stream` (a cudaStream_t).
// A row whose observed set differs from the program's, or whose state is out of range, raises the
// executor's error flag (see executor_check).
// d_mpe_states (mpe programs only, else nullptr): int8 [n_rows][|r|], the maximising state of every variable of r.
void executor_run(Executor* ex, int64_t n_rows, const int8_t* d_evidence, double* d_out, void* stream, RunStats* stats,
                  bool normalise = true, int8_t* d_mpe_states = nullptr);

// factorNorm + factorDivide of an expanded, UNNORMALISED result matrix (device pointer), per row and per
// query: the last step of a split-clique query after the partial results of the slices have been summed.
void executor_normalise(Executor* ex, int64_t n_rows, double* d_out, void* stream);

// Synchronises `stream` and throws HB_E_ARG if a row of an earlier run was inconsistent with the program.
void executor_check(Executor* ex, void* stream);

int executor_device(const Executor* ex);
// Generates, compiles (NVRTC) and loads the kernels specialised for the heaviest steps of the program now, instead of
// at the first large batch; returns how many steps they cover (0: disabled or unavailable).  Idempotent.
int executor_specialise(Executor* ex);
// how many steps of the program currently run on a kernel specialised for them (jit.cpp); 0 = templated kernels only
int executor_specialised_kernels(const Executor* ex);
// true when the program runs on the whole-schedule on-chip kernel (one launch per batch, per-row tables in shared memory)
bool executor_on_chip(const Executor* ex);
// the NCCL communicator (an ncclComm_t, owned by the handle) a split-clique program exchanges its partial sums over
void executor_set_comm(Executor* ex, void* comm);
// writes the states of this rank's slice into the split variables' columns of a (device) evidence matrix; a row that
// observes a split variable at another state raises the executor's error flag
void executor_fill_fixed(Executor* ex, int8_t* d_evidence, int64_t n_rows, void* stream);
// HB_TREE_PROFILE=1 (read when the kernel is generated): SM cycles per phase of the on-chip kernel since the last call,
// summed over all tiles -- [0] evidence preparation, [1 + l] dependency level l, [1 + levels] normalise + write-out
void executor_tree_profile(Executor* ex, void* stream, int64_t out64[64]);
// where the installed generated kernels came from: -1 none, 0 this process' registry, 1 disk cache, 2 compiled now
int executor_jit_origin(const Executor* ex);
const char* executor_jit_key(const Executor* ex);  // cache key of the installed cubin ("" when none)
// 1e-12 mode: kernels generated AFTER this call are compiled with multiply-add contraction (not bit-exact)
void executor_set_fma(Executor* ex, bool on);
// on-chip plan: [0] eligible, [1] rows per tile, [2] threads per CTA, [3] dependency levels, [4] arena entries per row,
// [5] shared-memory bytes per CTA, [6] fp64 multiplies per row, [7] fp64 adds per row, [8] / [9] shared-memory table
// entries read / written per row by the generated kernel (0 until it has been generated)
void executor_tree_info(const Executor* ex, int64_t out10[10]);

// CUDA-event timing of the kernel groups on the launching stream (off by default).  After a run,
// executor_last_timing synchronises the stream and returns milliseconds summed over the row tiles:
// [0] prep_rows_kernel, [1] all prodsum kernels, [2] output_kernel, [3] whole pass.
void executor_set_timing(Executor* ex, bool on);
void executor_last_timing(Executor* ex, void* stream, double out4[4]);

// ---- learnEM accumulation (em.cu); plain device buffers, no Executor ----
// d_acc[c] = (...((d_acc[c] + rows[0][c]) + rows[1][c]) + ...): the reference's sample-by-sample cptSum.
void em_accumulate(double* d_acc, const double* d_rows, int64_t n_rows, int64_t width, void* stream);
// d_dst[d_row_of[i]][d_col_of[c]] = d_src[i][c]
void em_scatter(double* d_dst, const double* d_src, const int64_t* d_row_of, const uint32_t* d_col_of, int64_t n_rows, int64_t width,
                void* stream);
// cptDivide of the accumulated family / parent tables, entry e of the concatenated learnt tables reading
// family column d_a_col[e] and parent column d_b_col[e] (-1: no parents, scale by 1/n_samples).
void em_divide(double* d_out, const double* d_acc, const uint32_t* d_a_col, const int32_t* d_b_col, int64_t n_entries, double n_samples,
               void* stream);

}  // namespace hb
