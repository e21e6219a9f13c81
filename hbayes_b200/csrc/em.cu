// em.cu -- the accumulation side of learnEM (Bayes/EM.hs:16-66) on the device.
//
// The junction-tree program of hb_jt_em_step leaves, per sample, the posterior of every CPT family and of
// every parent set in one row of a [rows][width] matrix.  What the reference then does is
//   foldl1 sumG results            -- cptSum [acc, sample]: acc + sample, sample after sample (EM.hs:18-27,44)
//   fmapWithVertex divideG ...     -- cptDivide family parents (EM.hs:29-32, Bayes/Factor/CPT.hs:144-160)
// Both are restated here with the reference's association so that the learnt tables are bit-identical:
//   * em_colsum_kernel: one thread per column, a strict left fold over the rows in sample order (the rows are
//     staged through shared memory with cp.async, the additions stay one dependent chain per column);
//   * em_divide_kernel: `_factorProduct (Op _ 1.0 divide) [b, a]` = divide 1.0 (divide b a) with
//     `divide _ 0 = 0` (CPT.hs:113-114), or factorScale (1.0 / n) a for a vertex without parents.
#include <cuda_runtime.h>

#include <algorithm>
#include <string>

#include "engine.hpp"

namespace hb {

#define HB_CUDA(call)                                                                                      \
  do {                                                                                                     \
    cudaError_t e_ = (call);                                                                               \
    if (e_ != cudaSuccess)                                                                                 \
      throw Error(e_ == cudaErrorMemoryAllocation ? HB_E_OOM : HB_E_CUDA,                                  \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                                     \
  } while (0)

namespace {

// One CTA owns 32 adjacent columns.  All 8 warps stream tiles of EM_TILE rows x 32 columns into shared memory with
// cp.async (EM_STAGES tiles in flight, so the loads are not latency-bound); warp 0 then adds the tile's rows to its
// running sums in row order -- the dependent chain of additions IS the reference's fold and is not reassociated.
constexpr int EM_TILE = 256, EM_STAGES = 3, EM_THREADS = 256;
constexpr int EM_SMEM = EM_STAGES * EM_TILE * 32 * (int)sizeof(double);

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

__global__ void __launch_bounds__(EM_THREADS) em_colsum_kernel(double* __restrict__ acc, const double* __restrict__ rows, int64_t n_rows,
                                                               int64_t width) {
  extern __shared__ double tile[];  // [EM_STAGES][EM_TILE][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + lane;
  const bool live = c < width;
  const int64_t n_tiles = (n_rows + EM_TILE - 1) / EM_TILE;
  auto issue = [&](int64_t t) {
    if (t < n_tiles && live) {
      double* dst = tile + (t % EM_STAGES) * (EM_TILE * 32) + lane;
      const int64_t r0 = t * EM_TILE;
      const int nr = (int)min((int64_t)EM_TILE, n_rows - r0);
      const double* src = rows + r0 * width + c;
      for (int rr = warp; rr < nr; rr += EM_THREADS / 32) cp_async8(dst + rr * 32, src + rr * width);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int t = 0; t < EM_STAGES - 1; ++t) issue(t);
  double a = (warp == 0 && live) ? acc[c] : 0.0;
  for (int64_t t = 0; t < n_tiles; ++t) {
    issue(t + EM_STAGES - 1);
    asm volatile("cp.async.wait_group %0;" ::"n"(EM_STAGES - 1) : "memory");
    __syncthreads();
    if (warp == 0 && live) {
      const double* src = tile + (t % EM_STAGES) * (EM_TILE * 32) + lane;
      const int nr = (int)min((int64_t)EM_TILE, n_rows - t * EM_TILE);
      int rr = 0;
      for (; rr + 8 <= nr; rr += 8) {
        double x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = src[(rr + j) * 32];
#pragma unroll
        for (int j = 0; j < 8; ++j) a = __dadd_rn(a, x[j]);
      }
      for (; rr < nr; ++rr) a = __dadd_rn(a, src[rr * 32]);
    }
    __syncthreads();  // the stage is refilled by the next iteration's issue
  }
  if (warp == 0 && live) acc[c] = a;
}

// dst[row_of[i]][col_of[c]] = src[i][c]: puts the rows of one observed-set group back at their sample
// positions, in the column layout of the first sample's program
__global__ void __launch_bounds__(256) em_scatter_kernel(double* __restrict__ dst, const double* __restrict__ src,
                                                         const int64_t* __restrict__ row_of, const uint32_t* __restrict__ col_of,
                                                         int64_t n_rows, int64_t width) {
  const int64_t total = n_rows * width;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / width, c = t - i * width;
    dst[row_of[i] * width + col_of[c]] = src[t];
  }
}

__global__ void __launch_bounds__(256) em_divide_kernel(double* __restrict__ out, const double* __restrict__ acc,
                                                        const uint32_t* __restrict__ a_col, const int32_t* __restrict__ b_col,
                                                        int64_t n_entries, double n_samples) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_entries) return;
  const double va = acc[a_col[e]];
  double r;
  if (b_col[e] < 0) {  // no parents: b = Scalar n  ->  factorScale (1.0 / n) a
    r = __dmul_rn(__ddiv_rn(1.0, n_samples), va);
  } else {
    const double vb = acc[b_col[e]];
    const double inner = va == 0.0 ? 0.0 : __ddiv_rn(vb, va);
    r = inner == 0.0 ? 0.0 : __ddiv_rn(1.0, inner);
  }
  out[e] = r;
}

}  // namespace

void em_accumulate(double* d_acc, const double* d_rows, int64_t n_rows, int64_t width, void* stream) {
  if (n_rows <= 0 || width <= 0) return;
  HB_CUDA(cudaFuncSetAttribute(em_colsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, EM_SMEM));
  em_colsum_kernel<<<(unsigned)((width + 31) / 32), EM_THREADS, EM_SMEM, (cudaStream_t)stream>>>(d_acc, d_rows, n_rows, width);
  HB_CUDA(cudaGetLastError());
}

void em_scatter(double* d_dst, const double* d_src, const int64_t* d_row_of, const uint32_t* d_col_of, int64_t n_rows, int64_t width,
                void* stream) {
  if (n_rows <= 0 || width <= 0) return;
  const int64_t blocks = std::min<int64_t>((n_rows * width + 255) / 256, 148 * 8);
  em_scatter_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_dst, d_src, d_row_of, d_col_of, n_rows, width);
  HB_CUDA(cudaGetLastError());
}

void em_divide(double* d_out, const double* d_acc, const uint32_t* d_a_col, const int32_t* d_b_col, int64_t n_entries, double n_samples,
               void* stream) {
  if (n_entries <= 0) return;
  em_divide_kernel<<<(unsigned)((n_entries + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_out, d_acc, d_a_col, d_b_col, n_entries,
                                                                                         n_samples);
  HB_CUDA(cudaGetLastError());
}

}  // namespace hb
