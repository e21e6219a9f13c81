// engine.cu -- sm_100a kernels and the row-tiled replay of a compiled Program.
//
// Data layout in HBM
//   potential pool   double[pot_entries]            all CPTs (restricted to the fixed states of a split clique),
//                                                   reference layout, batch-invariant; procedural CPTs are
//                                                   generated here on the device
//   shared arena     double[shared_entries]         evidence-independent derived tables (computed once)
//   row arena        double[row_entries][RT]        per-row derived tables, ROW-FASTEST: entry e of row r
//                                                   of a tile sits at (off + e) * RT + r, so the 32 lanes
//                                                   of a warp (2 rows each, same entry) read and write 512
//                                                   contiguous bytes
//   slice offsets    int32[n_slices][RT]            entry offset of each sliced CPT for each row
//   evidence         int8[n_rows][n_vars]           caller's matrix, -1 = unobserved
//   output           double[n_rows][out_width]      caller's matrix
//
// Kernels
//   prep_rows_kernel            evidence -> slice offsets + consistency check        (Bayes/Factor.hs:90-94)
//   prodsum_kernel<K,V,D,MASK>  fused factor product + sum-out of one variable + evidence slicing
//                               (marginalizeOneVariable = _factorProduct + cptFactorProjectOutWith,
//                                Buckets.hs:105-111, PrivateCPT.hs:215-266)
//   fused_kernel<KO,KI,V,DI,MASK>  a chain of up to four such operations whose intermediate tables stay in
//                               registers (Step::levels)
//   prodsum_generic_kernel      the same for buckets with more tables than the templates take
//   output_kernel               factorNorm + factorDivide + expansion to the query scope
//                               (PrivateCPT.hs:167-187, Factor.hs:171-172)
//   normalise_kernel            the same normalisation after the slices of a split clique have been summed
//   proc_fill_kernel            procedural CPT values (procedural.hpp)
// All arithmetic is binary64 with explicit round-to-nearest multiply/add (no FMA contraction), in the
// reference's association order: ps * (v1 * (v2 * ...)) and 0 + t0 + t1 + ...
// The per-row steps of a tile are captured into a CUDA graph in which independent steps overlap (executor_run).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <future>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "engine.hpp"
#include "jit.hpp"
#include "nccl_dyn.hpp"
#include "procedural.hpp"

namespace hb {

#define HB_CUDA(call)                                                                                      \
  do {                                                                                                     \
    cudaError_t e_ = (call);                                                                               \
    if (e_ != cudaSuccess)                                                                                 \
      throw Error(e_ == cudaErrorMemoryAllocation ? HB_E_OOM : HB_E_CUDA,                                  \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                                     \
  } while (0)

namespace {

constexpr int MAXK = 6;       // chain inputs handled by the templated kernel
constexpr int MAXS = 4;       // structural scalars per step
constexpr int MAXL = 4;       // levels of a fused chain
constexpr int MAXOBS = 8;     // observed variables inside one query
#ifndef HB_BLOCK
#define HB_BLOCK 256
#endif
constexpr int BLOCK = HB_BLOCK;

// A table as a kernel sees it: element (e, row) lives at base[e * mult + rowterm(row)].
struct TabRef {
  const double* base;     // pool + table offset (row tables: arena + off * RT)
  const int32_t* rowoff;  // sliced potential: per-row entry offset for this tile, else nullptr
  int is_row;             // 1: per-row table (mult = RT, rowterm = row)
  int is_one;             // 1: the constant 1.0
};

// Index maps are stored PRE-MULTIPLIED by the table's element stride (RT for row tables, 1 otherwise),
// so the offset of output entry o in an input is just hi[o / C] + lo[o % C] (uint32 elements).
struct InArg {
  TabRef t;
  const uint32_t* hi;
  const uint32_t* lo;
  unsigned sstride;       // stride of the summed variable, in elements (pre-multiplied like the maps)
  unsigned wstride;       // stride of the step's blocking variable (pre-multiplied)
};

template <int K>
struct StepArgs {
  InArg in[K > 0 ? K : 1];
  TabRef sc[MAXS];
  int n_sc;
  int d;                  // terms per output entry
  unsigned C;             // reduced output index o' = hiI * C + loI
  unsigned n_out;
  unsigned U, n_red;      // blocking variable: U states, n_red = n_out / U reduced entries; output entry = o' * U + u
  unsigned wst[K > 0 ? K : 1];  // stride of the blocking variable in each input (pre-multiplied; 0: independent of it)
  double* out;            // arena + off * RT (row) or shared + off
  int out_is_row;
  int nr;                 // rows in this tile
  long long RT;           // row stride of the arena
  signed char* arg;       // maximisation step: argmax table [entry][RT]; nullptr otherwise
  int scaled_in, n_scale; // input whose loaded entries are multiplied by the one-entry tables scale[0..n_scale) (-1: none)
  TabRef scale[MAX_SCALES];
};

struct GenericArgs {      // any number of inputs / scalars, descriptors in global memory
  const InArg* in;        // InArg::wstride holds the blocking-variable stride
  const TabRef* sc;
  int k, n_sc, d;
  unsigned C, n_out, U;
  double* out;
  int out_is_row, nr;
  long long RT;
  signed char* arg;       // max step (MAXCPT): argmax table, [entry][RT] like the arena; nullptr for a sum step
  int scaled_in, n_scale;
  TabRef scale[MAX_SCALES];
};

__device__ __forceinline__ const double* row_ptr(const TabRef& t, int row) {
  return t.base + (t.is_row ? (long long)row : (t.rowoff ? (long long)__ldg(t.rowoff + row) : 0ll));
}

// ps = s1 * (s2 * (... * sn))   (elementProduct over the scalars, PrivateCPT.hs:225-227)
__device__ __forceinline__ double fold_scalars(const TabRef* sc, int n, int row) {
  double ps = *row_ptr(sc[n - 1], row);
  for (int i = n - 2; i >= 0; --i) ps = __dmul_rn(*row_ptr(sc[i], row), ps);
  return ps;
}

// V rows per thread: a pair of adjacent rows of a row table is one 16-byte access.
template <int V>
struct Rows;

// the scale factors of a step's scaled input for this thread's rows (1.0 where there is none: never multiplied in)
template <class R, int V>
__device__ __forceinline__ void load_scales(const TabRef* scale, int n_scale, int row, R* sfac) {
#pragma unroll
  for (int j = 0; j < MAX_SCALES; ++j)
    sfac[j] = j < n_scale ? R::splat(*row_ptr(scale[j], row), V == 2 ? *row_ptr(scale[j], row + 1) : 0.0) : R::splat(1.0, 1.0);
}
// x = ((x * s0) * s1) * ...
template <class R>
__device__ __forceinline__ R apply_scales(R x, const R* sfac, int n_scale) {
#pragma unroll
  for (int j = 0; j < MAX_SCALES; ++j)
    if (j < n_scale) x = x.mul(sfac[j]);
  return x;
}
template <>
struct Rows<1> {
  double x;
  __device__ __forceinline__ static Rows load(const double* p0, const double*, bool, unsigned off) { return {p0[off]}; }
  __device__ __forceinline__ static Rows splat(double a, double) { return {a}; }
  __device__ __forceinline__ Rows mul(const Rows& o) const { return {__dmul_rn(x, o.x)}; }
  __device__ __forceinline__ Rows add(const Rows& o) const { return {__dadd_rn(x, o.x)}; }
  // maximumBy (compare `on` fst) (MaxCPT.hs:41): the later term replaces the kept one unless the kept one is GT
  __device__ __forceinline__ void take_max(const Rows& c, int s, int* arg) {
    if (s == 0 || x <= c.x) x = c.x, arg[0] = s;
  }
  __device__ __forceinline__ static void store_arg(signed char* p, const int* arg) { *p = (signed char)arg[0]; }
  __device__ __forceinline__ void store(double* p) const {
#ifndef HB_PLAIN_STORES
    __stcs(p, x);
#else
    *p = x;
#endif
  }
};
template <>
struct Rows<2> {
  double x, y;
  __device__ __forceinline__ static Rows load(const double* p0, const double* p1, bool is_row, unsigned off) {
    if (is_row) {
      const double2 v = *reinterpret_cast<const double2*>(p0 + off);
      return {v.x, v.y};
    }
    return {p0[off], p1[off]};
  }
  __device__ __forceinline__ static Rows splat(double a, double b) { return {a, b}; }
  __device__ __forceinline__ Rows mul(const Rows& o) const { return {__dmul_rn(x, o.x), __dmul_rn(y, o.y)}; }
  __device__ __forceinline__ Rows add(const Rows& o) const { return {__dadd_rn(x, o.x), __dadd_rn(y, o.y)}; }
  __device__ __forceinline__ void take_max(const Rows& c, int s, int* arg) {
    if (s == 0 || x <= c.x) x = c.x, arg[0] = s;
    if (s == 0 || y <= c.y) y = c.y, arg[1] = s;
  }
  __device__ __forceinline__ static void store_arg(signed char* p, const int* arg) {  // rows (2t, 2t+1): one aligned 2-byte store
    *reinterpret_cast<char2*>(p) = make_char2((signed char)arg[0], (signed char)arg[1]);
  }
  __device__ __forceinline__ void store(double* p) const {
#ifndef HB_PLAIN_STORES
    __stcs(reinterpret_cast<double2*>(p), make_double2(x, y));  // evict-first: a derived table is read again only many launches later
#else
    *reinterpret_cast<double2*>(p) = make_double2(x, y);
#endif
  }
};

// Walks a strip of (reduced output entry o', pair of blocking-variable states) items.
struct Strip {
  unsigned it, it_end, op, ch, nch, hiI, loI;
  __device__ __forceinline__ bool init(unsigned n_red, unsigned U, unsigned C) {
    nch = (U + 1) >> 1;
    const unsigned n_items = n_red * nch;
    const unsigned lanes = gridDim.x * blockDim.y, lane = blockIdx.x * blockDim.y + threadIdx.y;  // strips: the fast grid axis
    const unsigned per = (n_items + lanes - 1) / lanes;
    it = lane * per;
    it_end = min(n_items, it + per);
    if (it >= it_end) return false;
    op = it / nch;
    ch = it - op * nch;
    hiI = op / C;
    loI = op - hiI * C;
    return true;
  }
  __device__ __forceinline__ void next(unsigned C) {
    ++it;
    if (++ch == nch) {
      ch = 0;
      ++op;
      if (++loI == C) loI = 0, ++hiI;
    }
  }
};

// The fused clique-message kernel: out[o] = 0 + sum_s ps * (in1[o,s] * (in2[o,s] * (... * inK[o,s]))).
// One thread owns V adjacent evidence rows (threadIdx.x) and walks a contiguous strip of output entries
// (threadIdx.y / blockIdx.y), U entries per iteration so that U*K*D independent loads are in flight.
// All lanes of a warp share the output entry: index-map loads are broadcasts, and every row-table access
// of the warp is one contiguous 256*V-byte segment.  D = terms per entry (0: run-time count).
// MASK: bit i set = input i depends on the blocking variable (two loads per item); clear = one load serves both
// blocked states.  The mask is a template parameter so that the loads stay unconditional, back to back.
// SC: one input is read through scale factors (compile.cpp "scale steps"); a template parameter so that every other
// instantiation is the code it was without them (as run-time tests they cost registers and spills in all 400 kernels:
// C3 4.7 -> 7.1 ms, C5 24.2 -> 28.7 ms).  Scaled steps without a generated kernel run the <D = 0, full MASK> variants.
template <int K, int V, int D, int MASK, bool MAXM = false, bool SC = false>
__global__ void __launch_bounds__(BLOCK) prodsum_kernel(const __grid_constant__ StepArgs<K> a) {
  typedef Rows<V> R;
  const int row = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  if (row >= a.nr) return;  // V == 2: the pad row after an odd tile exists in the arena and in rowoff
  const double* p0[K];
  const double* p1[K];
  bool isrow[K];
  unsigned ss[K];
#pragma unroll
  for (int i = 0; i < K; ++i) {
    isrow[i] = a.in[i].t.is_row != 0;
    p0[i] = row_ptr(a.in[i].t, row);
    p1[i] = (V == 2 && !isrow[i]) ? row_ptr(a.in[i].t, row + 1) : p0[i];
    ss[i] = a.in[i].sstride;
  }
  const bool has_ps = a.n_sc > 0;
  R ps = R::splat(1.0, 1.0);
  if (has_ps) ps = R::splat(fold_scalars(a.sc, a.n_sc, row), V == 2 ? fold_scalars(a.sc, a.n_sc, row + 1) : 0.0);
  double* outp = a.out + (a.out_is_row ? row : 0);
  const size_t om = a.out_is_row ? (size_t)a.RT : 1;
  const int d = D > 0 ? D : a.d;
  R sfac[MAX_SCALES];
  if constexpr (SC) load_scales<R, V>(a.scale, a.n_scale, row, sfac);
  Strip sp;
  if (!sp.init(a.n_red, a.U, a.C)) return;
  for (; sp.it < sp.it_end; sp.next(a.C)) {
    const unsigned u0 = 2 * sp.ch;
    const bool has1 = u0 + 1 < a.U;
    unsigned off0[K], off1[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
      off0[i] = __ldg(a.in[i].hi + sp.hiI) + __ldg(a.in[i].lo + sp.loI) + u0 * a.wst[i];
      off1[i] = off0[i] + (has1 ? a.wst[i] : 0);  // odd U: the last pair repeats its first state, not stored
    }
    R acc0 = R::splat(0.0, 0.0), acc1 = R::splat(0.0, 0.0);
    int arg0[V], arg1[V];  // MAXM: state of the kept term per row
    auto term = [&](int s) {
      R x0[K], x1[K];
#pragma unroll
      for (int i = 0; i < K; ++i) {
        x0[i] = R::load(p0[i], p1[i], isrow[i], off0[i] + s * ss[i]);
        if constexpr (SC)
          if (i == a.scaled_in) x0[i] = apply_scales(x0[i], sfac, a.n_scale);
        if ((MASK >> i) & 1) {
          x1[i] = R::load(p0[i], p1[i], isrow[i], off1[i] + s * ss[i]);
          if constexpr (SC)
            if (i == a.scaled_in) x1[i] = apply_scales(x1[i], sfac, a.n_scale);
        } else
          x1[i] = x0[i];
      }
      R c0 = x0[K - 1], c1 = x1[K - 1];
#pragma unroll
      for (int i = K - 2; i >= 0; --i) c0 = x0[i].mul(c0), c1 = x1[i].mul(c1);
      if (has_ps) c0 = ps.mul(c0), c1 = ps.mul(c1);
      if constexpr (MAXM)
        acc0.take_max(c0, s, arg0), acc1.take_max(c1, s, arg1);
      else
        acc0 = acc0.add(c0), acc1 = acc1.add(c1);
    };
    if (D > 0) {
#pragma unroll
      for (int s = 0; s < (D > 0 ? D : 1); ++s) term(s);
    } else {
      for (int s = 0; s < d; ++s) term(s);
    }
    const size_t o = (size_t)sp.op * a.U + u0;
    acc0.store(outp + o * om);
    if (has1) acc1.store(outp + (o + 1) * om);
    if constexpr (MAXM) {
      R::store_arg(a.arg + o * (size_t)a.RT + row, arg0);
      if (has1) R::store_arg(a.arg + (o + 1) * (size_t)a.RT + row, arg1);
    }
  }
}

// ---- fused chains (Step::levels.size() > 1) ----------------------------------------------------
// value_l = 0 + sum_{s_l} value_{l+1} * (in_1 * (in_2 * ...)); the inner value is a register, the table it
// stands for never exists.  Inputs are flattened level 0 first; input i belongs to level lvl[i] and moves
// by st[i][m] entries per state of the summed variable of level m <= lvl[i].
template <int K>
struct FusedArgs {
  InArg in[K];
  unsigned st[K][MAXL];
  int lvl[K];
  int L;
  int d[MAXL];
  unsigned C, n_out;
  unsigned U, n_red;      // blocking variable (see StepArgs)
  unsigned wst[K];
  double* out;
  int out_is_row, nr;
  long long RT;
  int scaled_in, n_scale;
  TabRef scale[MAX_SCALES];
};

constexpr int FU = 2;  // output entries evaluated in lock-step by one thread of the fused kernel

// Fused chain kernel.  Template: KO = stored inputs of the outer levels (flattened first, level from
// lvl[]), KI = stored inputs of the innermost level (the last KI inputs), DI = terms of the innermost
// sum (0: run-time count).  The innermost level is where the loads are: it is fully static, so all its
// FU * DI * KI loads are issued back to back.  The outer levels are walked as an odometer (run-time
// uniform control flow, executed once per finished inner sum).
#ifndef HB_FUSED_MINB
#define HB_FUSED_MINB 2
#endif
template <int KO, int KI, int V, int DI, int MASK, bool SC = false>
__global__ void __launch_bounds__(BLOCK, HB_FUSED_MINB) fused_kernel(const __grid_constant__ FusedArgs<KO + KI> a) {
  constexpr int K = KO + KI;
  typedef Rows<V> R;
  const int row = (blockIdx.y * blockDim.x + threadIdx.x) * V;
  if (row >= a.nr) return;
  const double* p0[K];
  const double* p1[K];
  bool isrow[K];
#pragma unroll
  for (int i = 0; i < K; ++i) {
    isrow[i] = a.in[i].t.is_row != 0;
    p0[i] = row_ptr(a.in[i].t, row);
    p1[i] = (V == 2 && !isrow[i]) ? row_ptr(a.in[i].t, row + 1) : p0[i];
  }
  double* outp = a.out + (a.out_is_row ? row : 0);
  const size_t om = a.out_is_row ? (size_t)a.RT : 1;
  const int inner = a.L - 1;
  const int d_inner = DI > 0 ? DI : a.d[inner];
  R sfac[MAX_SCALES];
  if constexpr (SC) load_scales<R, V>(a.scale, a.n_scale, row, sfac);
  unsigned sti[KI];  // stride of the innermost summed variable in the innermost inputs
#pragma unroll
  for (int i = 0; i < KI; ++i) sti[i] = a.st[KO + i][inner];
  Strip sp;
  if (!sp.init(a.n_red, a.U, a.C)) return;
  for (; sp.it < sp.it_end; sp.next(a.C)) {
    const unsigned u0 = 2 * sp.ch;
    const bool has1 = u0 + 1 < a.U;
    unsigned cur0[K], cur1[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
      cur0[i] = __ldg(a.in[i].hi + sp.hiI) + __ldg(a.in[i].lo + sp.loI) + u0 * a.wst[i];
      cur1[i] = cur0[i] + (has1 ? a.wst[i] : 0);  // odd U: the last pair repeats its first state, not stored
    }
    R acc0[MAXL - 1], acc1[MAXL - 1], v0, v1;
    int digit[MAXL - 1];
#pragma unroll
    for (int l = 0; l < MAXL - 1; ++l) digit[l] = 0, acc0[l] = R::splat(0.0, 0.0), acc1[l] = R::splat(0.0, 0.0);
    for (;;) {
      // ---- innermost level: v = 0 + sum_s in_1 * (in_2 * ...), static shape; an input whose MASK bit is clear
      // does not depend on the blocking variable and is loaded once for both blocked states ----
      v0 = R::splat(0.0, 0.0), v1 = R::splat(0.0, 0.0);
      auto term = [&](int q) {
        R x0[KI], x1[KI];
#pragma unroll
        for (int i = 0; i < KI; ++i) {
          x0[i] = R::load(p0[KO + i], p1[KO + i], isrow[KO + i], cur0[KO + i] + q * sti[i]);
          if constexpr (SC)
            if (KO + i == a.scaled_in) x0[i] = apply_scales(x0[i], sfac, a.n_scale);
          if ((MASK >> i) & 1) {
            x1[i] = R::load(p0[KO + i], p1[KO + i], isrow[KO + i], cur1[KO + i] + q * sti[i]);
            if constexpr (SC)
              if (KO + i == a.scaled_in) x1[i] = apply_scales(x1[i], sfac, a.n_scale);
          } else
            x1[i] = x0[i];
        }
        R c0 = x0[KI - 1], c1 = x1[KI - 1];
#pragma unroll
        for (int i = KI - 2; i >= 0; --i) c0 = x0[i].mul(c0), c1 = x1[i].mul(c1);
        v0 = v0.add(c0), v1 = v1.add(c1);
      };
      if (DI > 0) {
#pragma unroll
        for (int q = 0; q < (DI > 0 ? DI : 1); ++q) term(q);
      } else {
        for (int q = 0; q < d_inner; ++q) term(q);
      }
      // ---- carry the finished value up through the outer levels ----
      int l = inner - 1;
#pragma unroll
      for (int ll = MAXL - 2; ll >= 0; --ll) {
        if (ll == l) {
          R r0, r1;
          bool have = false;
#pragma unroll
          for (int i = KO - 1; i >= 0; --i) {
            if (a.lvl[i] == ll) {
              R x0 = R::load(p0[i], p1[i], isrow[i], cur0[i]);
              R x1 = R::load(p0[i], p1[i], isrow[i], cur1[i]);
              if constexpr (SC)
                if (i == a.scaled_in) x0 = apply_scales(x0, sfac, a.n_scale), x1 = apply_scales(x1, sfac, a.n_scale);
              r0 = have ? x0.mul(r0) : x0;
              r1 = have ? x1.mul(r1) : x1;
              have = true;
            }
          }
          acc0[ll] = acc0[ll].add(have ? v0.mul(r0) : v0);
          acc1[ll] = acc1[ll].add(have ? v1.mul(r1) : v1);
          if (++digit[ll] < a.d[ll]) {
#pragma unroll
            for (int i = 0; i < K; ++i) cur0[i] += a.st[i][ll], cur1[i] += a.st[i][ll];
            l = -2;  // next term of this level: back to the innermost sum
          } else {
            const unsigned back = (unsigned)(a.d[ll] - 1);
#pragma unroll
            for (int i = 0; i < K; ++i) cur0[i] -= back * a.st[i][ll], cur1[i] -= back * a.st[i][ll];
            v0 = acc0[ll], v1 = acc1[ll];
            acc0[ll] = R::splat(0.0, 0.0), acc1[ll] = R::splat(0.0, 0.0);
            digit[ll] = 0;
            l = ll - 1;
          }
        }
      }
      if (l == -1) break;  // level 0 finished (or there is no outer level): v is the output entry
    }
    const size_t o = (size_t)sp.op * a.U + u0;
    v0.store(outp + o * om);
    if (has1) v1.store(outp + (o + 1) * om);
  }
}

// Any number of inputs and scalars (descriptors in global memory), one row per thread.  Rare: buckets
// with more than MAXK tables, or steps made of scalars only.
__global__ void __launch_bounds__(BLOCK) prodsum_generic_kernel(const GenericArgs a) {
  const int row = blockIdx.y * blockDim.x + threadIdx.x;
  if (row >= a.nr) return;
  const double ps = a.n_sc > 0 ? fold_scalars(a.sc, a.n_sc, row) : 1.0;
  double* outp = a.out + (a.out_is_row ? row : 0);
  const size_t om = a.out_is_row ? (size_t)a.RT : 1;
  const unsigned step = gridDim.x * blockDim.y;
  for (unsigned o = blockIdx.x * blockDim.y + threadIdx.y; o < a.n_out; o += step) {
    const unsigned op = o / a.U, u = o - op * a.U;
    const unsigned hiI = op / a.C, loI = op - hiI * a.C;
    double acc = 0.0;
    int arg = 0;
    for (int s = 0; s < a.d; ++s) {
      double c = ps;
      for (int i = a.k - 1; i >= 0; --i) {
        const InArg& in = a.in[i];
        const unsigned off = __ldg(in.hi + hiI) + __ldg(in.lo + loI) + u * in.wstride + s * in.sstride;
        double v = row_ptr(in.t, row)[off];
        if (i == a.scaled_in)
          for (int j = 0; j < a.n_scale; ++j) v = __dmul_rn(v, *row_ptr(a.scale[j], row));
        c = (i == a.k - 1) ? v : __dmul_rn(v, c);
      }
      if (a.k > 0 && a.n_sc > 0) c = __dmul_rn(ps, c);
      if (a.arg) {  // maximumBy (compare `on` fst): the later term wins unless the kept one compares GT (MaxCPT.hs:41)
        if (s == 0 || acc <= c) acc = c, arg = s;
      } else
        acc = __dadd_rn(acc, c);
    }
    outp[(size_t)o * om] = acc;
    if (a.arg) a.arg[(size_t)o * a.RT + row] = (signed char)arg;
  }
}

// ---- potential pool: procedural CPTs are generated in place (restricted to the fixed states) ------
struct ProcFillArgs {
  double* dst;          // pool + pot_off[v]
  long long entries;    // stored entries
  unsigned long long seed, npar, full_base;
  int d, n_free;        // child states, number of non-fixed scope variables (slowest first)
  int dim[40];
  unsigned long long full_stride[40];
};

__global__ void __launch_bounds__(BLOCK) proc_fill_kernel(const ProcFillArgs a) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < a.entries; e += (long long)gridDim.x * blockDim.x) {
    unsigned long long idx = a.full_base, rest = (unsigned long long)e;
    for (int j = a.n_free - 1; j >= 0; --j) {
      const unsigned long long dj = (unsigned long long)a.dim[j];
      idx += (rest % dj) * a.full_stride[j];
      rest /= dj;
    }
    a.dst[e] = proc_cpt_value(a.seed, a.d, a.npar, idx);
  }
}

// ---- evidence -> slice offsets, and the consistency check -------------------------------------
struct PrepArgs {
  const int8_t* ev;        // [nr][n_vars] for this tile
  int nr, n_vars, n_slices;
  long long RT;
  const int8_t* observed;  // [n_vars] 0 = unobserved, 1 = observed per row, 2 + s = fixed at state s
  const int32_t* dims;     // [n_vars]
  const int32_t* sl_begin; // [n_slices + 1]
  const int32_t* sl_var;
  const int32_t* sl_stride;
  int32_t* rowoff;         // [n_slices][RT]
  int* flag;
};

// thread = row; blockIdx.y = a chunk of PREP_CHUNK variables (consistency check) and of PREP_CHUNK slices: a network with
// hundreds of sliced potentials and few thousand rows would otherwise run this on a handful of CTAs (C3: 168 us of a 4.7 ms pass)
constexpr int PREP_CHUNK = 32;
__global__ void __launch_bounds__(BLOCK) prep_rows_kernel(const PrepArgs a) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  const int j0 = blockIdx.y * PREP_CHUNK, j1 = min(a.n_slices, j0 + PREP_CHUNK);
  if (row >= a.nr) {  // pad rows up to the next multiple of 64 read slice 0 (their results are never used)
    if (row < ((a.nr + 63) & ~63) && row < a.RT)
      for (int j = j0; j < j1; ++j) a.rowoff[(long long)j * a.RT + row] = 0;
    return;
  }
  const int8_t* e = a.ev + (long long)row * a.n_vars;
  bool bad = false;
  for (int v = blockIdx.y * PREP_CHUNK; v < a.n_vars; v += gridDim.y * PREP_CHUNK)
    for (int k = v; k < min(a.n_vars, v + PREP_CHUNK); ++k) {
      const int s = e[k];
      const int o = a.observed[k];
      bad |= o == 0 ? (s >= 0) : (o == 1 ? (s < 0 || s >= a.dims[k]) : (s != o - 2));
    }
  if (bad) atomicOr(a.flag, 1);
  for (int j = j0; j < j1; ++j) {
    int off = 0;
    for (int t = a.sl_begin[j]; t < a.sl_begin[j + 1]; ++t) {
      int s = e[a.sl_var[t]];
      s = s < 0 ? 0 : (s >= a.dims[a.sl_var[t]] ? 0 : s);  // keep reads in bounds for flagged rows
      off += a.sl_stride[t] * s;
    }
    a.rowoff[(long long)j * a.RT + row] = off;
  }
}

// ---- normalise + expand ------------------------------------------------------------------------
struct OutDesc {
  TabRef t;                // rowoff is rebased per tile at launch
  int slice;               // slice slot or -1
  int n_phys, n_full;
  long long out_col;
  int map_off, order_off;  // into the int32 map pool
  int n_obs;
  int obs_var[MAXOBS], obs_stride[MAXOBS], obs_dim[MAXOBS];
  int gather;              // split clique: left unnormalised here, summed over the ranks and normalised afterwards
};

struct OutChunk {           // a run of consecutive queries whose columns fit one shared-memory tile
  int desc_begin, desc_end;
  int col_begin, width;
  int direct;               // a single very wide query: written straight to the caller's matrix, no tile
};

struct OutArgs {
  const OutDesc* desc;
  const OutChunk* chunks;
  const uint32_t* maps;
  const int32_t* rowoff;   // [n_slices][RT]
  const int8_t* ev;        // this tile
  double* out;             // this tile: [nr][width]
  int nr, n_vars;
  long long RT, width;
  int normalise;           // 0: leave the query results unnormalised (partial sums of a split clique)
};

#ifndef HB_OUT_ROWS
#define HB_OUT_ROWS 32
#endif
constexpr int OUT_ROWS = HB_OUT_ROWS;  // evidence rows per CTA
constexpr int OUT_COLS = 128;    // at most this many output columns per chunk (a single wider query gets its own chunk)

// CTA = OUT_ROWS (32) rows x one column chunk.  Phase 1: thread = (row, query) -- norm = 0 + x0 + x1 + ... over the
// result's layout order, every entry of the full table is (1.0 / norm) * value, where an entry that
// contradicts the evidence has the value 0.0 (zero-probability evidence gives NaN exactly as 0 * (1/0)
// does in the reference); arena reads are coalesced along the rows, the tile goes to shared memory.
// Phase 2: the tile is written out row by row, 256-byte coalesced segments of the caller's matrix.
__global__ void __launch_bounds__(BLOCK) output_kernel(const OutArgs a) {
  extern __shared__ double tile[];  // [OUT_ROWS][width + 1]
  const OutChunk ch = a.chunks[blockIdx.y];
  const int row0 = blockIdx.x * OUT_ROWS;
  const int pitch = ch.width + 1;
  const int r = threadIdx.x % OUT_ROWS, row = row0 + r;
  if (row < a.nr) {
    const int8_t* e = a.ev + (long long)row * a.n_vars;
    for (int q = ch.desc_begin + threadIdx.x / OUT_ROWS; q < ch.desc_end; q += BLOCK / OUT_ROWS) {
      const OutDesc& d = a.desc[q];
      const double* p = nullptr;
      long long m = 1;
      if (!d.t.is_one) {
        if (d.t.is_row) {
          p = d.t.base + row;
          m = a.RT;
        } else
          p = d.t.base + (d.slice >= 0 ? (long long)a.rowoff[(long long)d.slice * a.RT + row] : 0ll);
      }
      const uint32_t* order = a.maps + d.order_off;
      const uint32_t* fmap = a.maps + d.map_off;
      double norm = 0.0;
      for (int j = 0; j < d.n_phys; ++j) norm = __dadd_rn(norm, p ? p[(long long)order[j] * m] : 1.0);
      const double inv = (a.normalise && !d.gather) ? __ddiv_rn(1.0, norm) : 1.0;
      int st[MAXOBS];
      for (int j = 0; j < d.n_obs; ++j) st[j] = e[d.obs_var[j]];
      double* o = ch.direct ? a.out + (long long)row * a.width + d.out_col : tile + r * pitch + (int)(d.out_col - ch.col_begin);
      for (int f = 0; f < d.n_full; ++f) {
        bool match = true;
        for (int j = 0; j < d.n_obs; ++j) match &= ((f / d.obs_stride[j]) % d.obs_dim[j]) == st[j];
        const double v = match ? (p ? p[(long long)fmap[f] * m] : 1.0) : 0.0;
        o[f] = __dmul_rn(inv, v);
      }
    }
  }
  if (ch.direct) return;
  __syncthreads();
  const int rows = min(OUT_ROWS, a.nr - row0);
  for (int idx = threadIdx.x; idx < rows * ch.width; idx += BLOCK) {
    const int rr = idx / ch.width, c = idx - rr * ch.width;
    a.out[(long long)(row0 + rr) * a.width + ch.col_begin + c] = tile[rr * pitch + c];
  }
}

// factorNorm + factorDivide of an already expanded, unnormalised result matrix (the sum over the slices of
// a split clique): per row and per query block, norm = 0 + x0 + x1 + ..., x_i *= 1.0 / norm.
struct NormArgs {
  double* out;
  const int* col_begin;  // [n_blocks + 1]
  int n_blocks;
  long long n_rows, width;
};
__global__ void __launch_bounds__(BLOCK) normalise_kernel(const NormArgs a) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= a.n_rows * a.n_blocks) return;
  const long long row = gid / a.n_blocks;
  const int b = (int)(gid % a.n_blocks);
  double* p = a.out + row * a.width;
  double norm = 0.0;
  for (int c = a.col_begin[b]; c < a.col_begin[b + 1]; ++c) norm = __dadd_rn(norm, p[c]);
  const double inv = __ddiv_rn(1.0, norm);
  for (int c = a.col_begin[b]; c < a.col_begin[b + 1]; ++c) p[c] = __dmul_rn(inv, p[c]);
}

// Split clique over NCCL, queries that mention a split variable: every rank has written its own slice of the unnormalised
// result (zeros elsewhere).  forward: the blocks of those queries are packed into a contiguous matrix [rows][gw] for the
// all-reduce; backward: the summed blocks are normalised (norm = 0 + x0 + x1 + ..., x * (1.0 / norm)) and written back.
struct GatherArgs {
  double* out;           // [n_rows][width]
  double* tmp;           // [n_rows][gw]
  const int* out_begin;  // [n_blocks + 1]: first column of each block in `out` (last entry unused)
  const int* tmp_begin;  // [n_blocks + 1]: first column of each block in `tmp`
  int n_blocks, backward, normalise;
  long long n_rows, width, gw;
};
__global__ void __launch_bounds__(BLOCK) gather_blocks_kernel(const GatherArgs a) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= a.n_rows * a.n_blocks) return;
  const long long row = gid / a.n_blocks;
  const int b = (int)(gid % a.n_blocks);
  double* o = a.out + row * a.width + a.out_begin[b];
  double* t = a.tmp + row * a.gw + a.tmp_begin[b];
  const int n = a.tmp_begin[b + 1] - a.tmp_begin[b];
  if (!a.backward) {
    for (int c = 0; c < n; ++c) t[c] = o[c];
    return;
  }
  double norm = 0.0;
  for (int c = 0; c < n; ++c) norm = __dadd_rn(norm, t[c]);
  const double inv = a.normalise ? __ddiv_rn(1.0, norm) : 1.0;
  for (int c = 0; c < n; ++c) o[c] = __dmul_rn(inv, t[c]);
}

// Split clique over NCCL: the caller's rows leave the split variables unobserved (-1); every rank fills in the states of
// ITS slice.  A row that observes a split variable at another state contradicts the slice: flagged.
struct FillArgs {
  int8_t* ev;               // [n_rows][n_vars]
  const int8_t* observed;   // as PrepArgs::observed: 2 + s = fixed at state s
  long long n_rows;
  int n_vars;
  int* flag;
};
__global__ void __launch_bounds__(BLOCK) fill_fixed_kernel(const FillArgs a) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= a.n_rows * a.n_vars) return;
  const int o = a.observed[gid % a.n_vars];
  if (o < 2) return;
  const int s = a.ev[gid];
  if (s < 0)
    a.ev[gid] = (int8_t)(o - 2);
  else if (s != o - 2)
    atomicOr(a.flag, 1);
}

// mpe traceback (mpeInstantiations, Bayes/Factor/MaxCPT.hs:73-75): the reference carries the maximising
// instantiation inside every MAXCPT entry; here each maximisation step left an argmax table, and the states are read
// back from the last eliminated variable to the first -- a step's table is indexed by variables maximised after it.
// desc, per step in REVERSE elimination order: slot, n_dep, arg_off lo, arg_off hi, then n_dep x (slot, stride).
struct TraceArgs {
  const int32_t* desc;
  int n_steps, n_r, nr;
  const signed char* args;
  long long RT;
  signed char* states;  // [nr][n_r] of this tile; -1 = the variable got no instantiation
};
__global__ void __launch_bounds__(BLOCK) mpe_trace_kernel(const TraceArgs a) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= a.nr) return;
  signed char* st = a.states + (size_t)row * a.n_r;
  for (int i = 0; i < a.n_r; ++i) st[i] = -1;
  const int32_t* d = a.desc;
  for (int j = 0; j < a.n_steps; ++j) {
    const int slot = d[0], n = d[1];
    long long idx = (long long)(unsigned)d[2] | ((long long)d[3] << 32);
    for (int k = 0; k < n; ++k) idx += (long long)d[4 + 2 * k + 1] * st[d[4 + 2 * k]];
    st[slot] = a.args[(size_t)idx * a.RT + row];
    d += 4 + 2 * n;
  }
}

}  // namespace

// ================================ host side =====================================================
struct Executor {
  const Program* P = nullptr;
  int device = 0;
  int64_t workspace = 0;
  double *d_pots = nullptr, *d_shared = nullptr, *d_arena = nullptr;
  signed char* d_args = nullptr;  // argmax arena of the maximisation steps: int8 [arg_entries][RT]
  int32_t* d_trace = nullptr;     // traceback descriptors (mpe), see mpe_trace_kernel
  int n_trace = 0;
  uint32_t *d_maps = nullptr;     // index maps, element stride 1 (potentials, shared tables, output maps)
  uint32_t *d_maps_rt = nullptr;  // the same maps multiplied by RT (row tables)
  int32_t* d_rowoff = nullptr;
  int8_t* d_observed = nullptr;
  int32_t *d_dims = nullptr, *d_sl_begin = nullptr, *d_sl_var = nullptr, *d_sl_stride = nullptr;
  OutDesc* d_out_desc = nullptr;
  OutChunk* d_out_chunks = nullptr;
  int* d_norm_cols = nullptr;
  std::vector<OutChunk> out_chunks;
  int out_tile_bytes = 0;
  InArg* d_gen_in = nullptr;     // descriptors of the generic-path steps (rebuilt when RT changes)
  TabRef* d_gen_sc = nullptr;
  int* d_flag = nullptr;
  int64_t RT = 0;                // current arena row stride
  int64_t max_row_table = 1;     // largest per-row table (entries): bounds RT so that offsets fit 32 bits
  std::vector<uint32_t> maps;    // host copy of the stride-1 maps
  std::vector<int64_t> hi_off, lo_off;  // per (step, input): offsets into the map pool
  std::vector<int64_t> in_base;         // first (step,input) index of each step
  std::vector<OutDesc> out_desc;
  std::vector<int> out_steps;
  // The product/sum-out launches of one tile only depend on (RT, rows in the tile): they are captured once
  // into a CUDA graph and replayed (programs with hundreds of small steps are launch-bound otherwise).
  std::map<int, cudaGraphExec_t> graphs;  // rows in tile -> instantiated graph, valid for the current RT
  cudaStream_t capture_stream = nullptr;
  int n_row_steps = 0;
  // Inside the graph independent steps may overlap: during capture they are spread over several streams,
  // ordered only by their true dependencies (computed once from the arena regions they read and write:
  // read-after-write, and write-after-read/write because the arena reuses memory by liveness).
  std::vector<cudaStream_t> side_streams;
  std::vector<cudaEvent_t> step_events;
  std::vector<std::vector<int>> step_deps;  // per row step (in launch order): earlier row steps it must follow
  std::vector<int> step_lane;               // per row step: which stream it is captured on
  // Launch nodes: the row steps ordered by dependency level (a topological order of step_deps): the steps that
  // only depend on finished work are issued first, whatever their position in the reference's elimination order.
  struct Node {
    int si = -1;            // program step index
    std::vector<int> deps;  // earlier nodes
    int lane = 0;
  };
  std::vector<Node> nodes;
  std::vector<int64_t> gen_in_at, gen_sc_at;  // per program step: descriptor offsets in d_gen_in / d_gen_sc, -1 = none
  // Kernels specialised per step (jit.cpp), built on the first pass over a large batch; jit_fn[si] = nullptr: templated kernel
  JitModule* jit = nullptr;
  std::vector<void*> jit_fn;
  bool jit_elided = false;          // the generated kernels were built on the premise of jit_zero_add_elidable()
  int jit_blocked = 0;              // how many of them use a register block over several output variables
  std::vector<JitShape> jit_shape;  // of the steps that have a kernel: how it walks its groups (the launch has to agree with the source)
  bool jit_tried = false;
  double work_seen = 0;  // joint entries evaluated so far (rows x product entries per row)
  // Whole-schedule on-chip kernel (jit.cpp, hb_tree): when the per-row tables of a tile fit shared memory the whole pass
  // is ONE launch and the HBM row arena is not used at all.
  // split clique over NCCL (Program::n_reduce / n_gather > 0): the communicator is the handle's, not the executor's
  void* comm = nullptr;
  double* d_gather = nullptr;  // [rows of a tile][gather width]
  size_t gather_cap = 0;
  int *d_gather_out = nullptr, *d_gather_tmp = nullptr;
  int gather_width = 0, gather_blocks = 0;
  TreePlan tree;
  bool tree_eligible = false;
  void* tree_fn = nullptr;
  int tree_grid = 0;
  int64_t tree_loads = 0, tree_stores = 0;  // shared-memory table entries read / written per evidence row
  bool fma = false;  // 1e-12 mode: generated kernels are compiled with multiply-add contraction (never bit-exact)
  // NVRTC runs on a background thread (host-only work); the cubin is installed by the next call that finds it ready,
  // so a stream of batches never stalls on a compilation.
  struct Pending {
    std::future<std::shared_ptr<const std::string>> cubin;
    std::shared_ptr<std::string> key = std::make_shared<std::string>();
    std::shared_ptr<std::string> why = std::make_shared<std::string>();
    bool is_tree = false;
    std::vector<int> steps;
  };
  std::unique_ptr<Pending> pending;
  int jit_origin = -1;  // where the installed cubin came from: 0 this process, 1 disk cache, 2 compiled now
  int n_lanes = 8;
  void drop_graphs() {
    for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
    graphs.clear();
  }
  bool timing = false;
  std::vector<cudaEvent_t> events;  // 4 per tile of the last run when timing is on
  size_t events_used = 0;
  cudaEvent_t next_event() {
    if (events_used == events.size()) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) throw Error(HB_E_CUDA, "cudaEventCreate failed");
      events.push_back(e);
    }
    return events[events_used++];
  }
  ~Executor() {
    if (pending && pending->cubin.valid()) pending->cubin.wait();
    cudaSetDevice(device);
    drop_graphs();
    jit_destroy(jit);
    if (capture_stream) cudaStreamDestroy(capture_stream);
    for (cudaStream_t st : side_streams) cudaStreamDestroy(st);
    for (cudaEvent_t e : step_events) cudaEventDestroy(e);
    for (cudaEvent_t e : events) cudaEventDestroy(e);
    for (void* p : {(void*)d_gather, (void*)d_gather_out, (void*)d_gather_tmp, (void*)d_args, (void*)d_trace, (void*)d_pots, (void*)d_shared, (void*)d_arena, (void*)d_maps, (void*)d_maps_rt, (void*)d_rowoff, (void*)d_observed,
                    (void*)d_dims, (void*)d_sl_begin, (void*)d_sl_var, (void*)d_sl_stride, (void*)d_out_desc, (void*)d_out_chunks, (void*)d_norm_cols, (void*)d_gen_in,
                    (void*)d_gen_sc, (void*)d_flag})
      if (p) cudaFree(p);
  }
};

namespace {

template <class T>
T* upload(const std::vector<T>& v) {
  T* d = nullptr;
  HB_CUDA(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(T)));
  if (!v.empty()) HB_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

template <class T>
void free_dev(T*& p) {
  if (p) HB_CUDA(cudaFree(p));
  p = nullptr;
}

TabRef tab_ref(const Executor& ex, const Table& t, int64_t RT) {
  TabRef r{nullptr, nullptr, 0, 0};
  switch (t.kind) {
    case T_ONE: r.is_one = 1; break;
    case T_POT:
      r.base = ex.d_pots + t.off;
      if (t.slice_slot >= 0) r.rowoff = ex.d_rowoff + (int64_t)t.slice_slot * RT;
      break;
    case T_SHARED: r.base = ex.d_shared + t.off; break;
    case T_ROW: r.base = ex.d_arena + t.off * RT, r.is_row = 1; break;
  }
  return r;
}

// input number `i` of a step, counting level 0 first
const StepInput& flat_input(const Step& st, int i, int* level = nullptr) {
  for (size_t l = 0; l < st.levels.size(); ++l) {
    if (i < (int)st.levels[l].in.size()) {
      if (level) *level = (int)l;
      return st.levels[l].in[i];
    }
    i -= (int)st.levels[l].in.size();
  }
  throw Error(HB_E_ARG, "input index out of range");
}

// the scaled input of a step (flat index, -1 if none) and its one-entry tables as the kernels see them
static int scaled_input(const Executor& ex, const Step& st, int64_t RT, TabRef* scale, int* n_scale) {
  int flat = 0;
  *n_scale = 0;
  for (const Level& l : st.levels)
    for (const StepInput& in : l.in) {
      if (!in.scales.empty()) {
        if ((int)in.scales.size() > MAX_SCALES) throw Error(HB_E_UNSUPPORTED, "too many scale factors on one input");
        *n_scale = (int)in.scales.size();
        for (size_t j = 0; j < in.scales.size(); ++j) scale[j] = tab_ref(ex, ex.P->tables[in.scales[j]], RT);
        return flat;
      }
      ++flat;
    }
  return -1;
}

InArg in_arg(const Executor& ex, const Step& st, int si, int i, int64_t RT) {
  int level = 0;
  const StepInput& in = flat_input(st, i, &level);
  const Table& t = ex.P->tables[in.table];
  InArg a;
  a.t = tab_ref(ex, t, RT);
  const uint32_t* pool = t.kind == T_ROW ? ex.d_maps_rt : ex.d_maps;
  a.hi = pool + ex.hi_off[ex.in_base[si] + i];
  a.lo = pool + ex.lo_off[ex.in_base[si] + i];
  a.sstride = (unsigned)(in.lstride[level] * (t.kind == T_ROW ? RT : 1));
  a.wstride = (unsigned)(in.wstride * (t.kind == T_ROW ? RT : 1));
  return a;
}

struct LaunchShape {
  dim3 grid, block;
};
// Launch shapes.  Rows run along x (blocks of up to BLOCK row-threads), output entries are cut into
// contiguous strips along threadIdx.y / blockIdx.x.
//  * streaming shape: a block is as wide in rows as possible; strips only to fill the machine.
//  * reuse shape: when a step enumerates small per-row inputs many times, a block owns just 32 row-threads
//    (one warp wide) and ALL output entries, so the rows of its inputs (`tile_bytes`) stay in L1 while the
//    block sweeps the entries -- the re-reads stop going to L2.
struct Reuse {
  int64_t row_entries = 0;  // entries of all per-row inputs
  int64_t row_loads = 0;    // loads of per-row inputs per row
};
LaunchShape launch_shape(int row_threads, int rows_per_thread, int64_t n_out, const Reuse& reuse = Reuse()) {
  LaunchShape s;
  constexpr bool reuse_on = true;
  constexpr int reuse_bx = 32;   // row-threads per block
  constexpr int reuse_kb = 96;   // tile budget, KB
  constexpr int reuse_min = 3;   // minimum re-read factor
  const int64_t tile_bytes = reuse.row_entries * reuse_bx * rows_per_thread * 8;
  if (reuse_on && row_threads >= 64 && n_out >= 32 && reuse.row_entries > 0 && reuse.row_loads >= reuse_min * reuse.row_entries &&
      tile_bytes <= (int64_t)reuse_kb * 1024) {
    s.block = dim3(reuse_bx, BLOCK / reuse_bx);
    s.grid = dim3(1, (row_threads + reuse_bx - 1) / reuse_bx);
    return s;
  }
  // grid.x = strips of output entries (the fast axis of the block scheduler), grid.y = row blocks: CTAs that are resident
  // together work on the same rows.  A step that re-reads its per-row inputs (too large for the L1 shape above) gets row
  // blocks one warp wide so that the strips of a row block find those rows in L2.
  int cap = reuse_on && reuse.row_entries > 0 && reuse.row_loads >= reuse_min * reuse.row_entries ? 32 : BLOCK;
  while ((row_threads + cap - 1) / cap > 65535) cap <<= 1;  // gridDim.y limit
  int bx = 1;
  while (bx < row_threads && bx < cap) bx <<= 1;
  const int by = std::max(1, BLOCK / bx);
  s.block = dim3(bx, by);
  s.grid.y = (row_threads + bx - 1) / bx;
  const int64_t max_x = (n_out + 2 * by - 1) / (2 * by);  // at least two entries per thread (U = 2)
  int64_t want = (148 * 8 * 2 + s.grid.y - 1) / s.grid.y;
  if (cap == 32) {  // a re-reading step: so many strips per row block that one wave of CTAs touches about 64 MB of rows (they stay in L2)
    const int64_t slab = (reuse.row_entries + n_out) * bx * rows_per_thread * 8;
    const int64_t blocks_per_wave = std::max<int64_t>(1, ((int64_t)64 << 20) / std::max<int64_t>(slab, 1));  // 16 .. 96 MB measured the same on C3
    want = std::max(want, (296 + blocks_per_wave - 1) / blocks_per_wave);
  }
  s.grid.x = (unsigned)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(max_x, want), 65535));
  return s;
}

// A thread takes two adjacent rows (16-byte accesses) in big tiles, and in tiles of 2..8 rows, where the row
// axis is too short to coalesce along and wider accesses are what helps (measured: C4, 4 rows, 15.7 -> 12.2 ms;
// at 16 rows one row per thread is the faster shape).
bool two_rows_per_thread(int nr) { return nr >= 64 || (nr >= 2 && nr <= 8); }

Reuse reuse_of(const Program& P, const Step& st) {
  Reuse r;
  int64_t terms = 1;
  for (const Level& l : st.levels) {
    terms *= l.d;
    for (const StepInput& in : l.in)
      if (P.tables[in.table].kind == T_ROW) r.row_entries += P.tables[in.table].size, r.row_loads += st.n_out * terms;
  }
  return r;
}

// picks the kernel instantiation whose MASK template argument equals the run-time mask
template <int K, int V, int D, int M>
struct ProdsumMask {
  static void go(int mask, const StepArgs<K>& a, const LaunchShape& sh, cudaStream_t stream) {
    if (mask == M)
      prodsum_kernel<K, V, D, M><<<sh.grid, sh.block, 0, stream>>>(a);
    else
      ProdsumMask<K, V, D, M - 1>::go(mask, a, sh, stream);
  }
};
template <int K, int V, int D>
struct ProdsumMask<K, V, D, -1> {
  static void go(int, const StepArgs<K>&, const LaunchShape&, cudaStream_t) { throw Error(HB_E_UNSUPPORTED, "bad blocking mask"); }
};

template <int K, int V, int D>
void launch_kvd(const StepArgs<K>& a, int mask, const LaunchShape& sh, cudaStream_t stream) {
  constexpr int FULL = (1 << K) - 1;
  if constexpr (V == 2 && K <= 3)
    ProdsumMask<K, V, D, FULL>::go(mask, a, sh, stream);
  else
    prodsum_kernel<K, V, D, FULL><<<sh.grid, sh.block, 0, stream>>>(a);  // every input loaded for both states
}

template <int K, int V>
void launch_kv(const StepArgs<K>& a, int d, int mask, const LaunchShape& sh, cudaStream_t stream) {
  switch (d) {
    case 1: launch_kvd<K, V, 1>(a, mask, sh, stream); break;
    case 2: launch_kvd<K, V, 2>(a, mask, sh, stream); break;
    case 3: launch_kvd<K, V, 3>(a, mask, sh, stream); break;
    case 4: launch_kvd<K, V, 4>(a, mask, sh, stream); break;
    default: launch_kvd<K, V, 0>(a, mask, sh, stream); break;
  }
}

template <int K>
void launch_k(const Executor& ex, const Step& st, int si, int64_t RT, int nr, cudaStream_t stream) {
  const Program& P = *ex.P;
  StepArgs<K> a;
  memset(&a, 0, sizeof a);
  int mask = 0;
  for (int i = 0; i < K; ++i) {
    a.in[i] = in_arg(ex, st, si, i, RT);
    a.wst[i] = a.in[i].wstride;
    if (a.wst[i] != 0 || K > 3) mask |= 1 << i;
  }
  a.U = (unsigned)st.U;
  a.n_red = (unsigned)st.n_red;
  a.n_sc = (int)st.scalars.size();
  for (int i = 0; i < a.n_sc; ++i) a.sc[i] = tab_ref(ex, P.tables[st.scalars[i]], RT);
  a.d = st.levels[0].d;
  a.C = (unsigned)st.C;
  a.n_out = (unsigned)st.n_out;
  const Table& out = P.tables[st.out];
  a.out_is_row = out.kind == T_ROW;
  a.out = a.out_is_row ? ex.d_arena + out.off * RT : ex.d_shared + out.off;
  a.nr = nr;
  a.RT = RT;
  a.scaled_in = scaled_input(ex, st, RT, a.scale, &a.n_scale);
  const int64_t items2 = st.n_red * ((st.U + 1) / 2) * 2;
  if (st.levels[0].is_max) {  // maximisation step (mpe): run-time term count, every input loaded for both blocked states
    constexpr int FULL = (1 << K) - 1;
    a.arg = ex.d_args + st.arg_off * RT;
    if (two_rows_per_thread(nr)) {
      const LaunchShape sh = launch_shape((nr + 1) / 2, 2, items2, reuse_of(P, st));
      prodsum_kernel<K, 2, 0, FULL, true><<<sh.grid, sh.block, 0, stream>>>(a);
    } else {
      const LaunchShape sh = launch_shape(nr, 1, items2);
      prodsum_kernel<K, 1, 0, FULL, true><<<sh.grid, sh.block, 0, stream>>>(a);
    }
    return;
  }
  if (a.scaled_in >= 0) {  // scaled input, no generated kernel: the general variant (run-time term count, every input loaded twice)
    constexpr int FULL = (1 << K) - 1;
    if (two_rows_per_thread(nr)) {
      const LaunchShape sh = launch_shape((nr + 1) / 2, 2, items2, reuse_of(P, st));
      prodsum_kernel<K, 2, 0, FULL, false, true><<<sh.grid, sh.block, 0, stream>>>(a);
    } else {
      const LaunchShape sh = launch_shape(nr, 1, items2);
      prodsum_kernel<K, 1, 0, FULL, false, true><<<sh.grid, sh.block, 0, stream>>>(a);
    }
    return;
  }
  if (two_rows_per_thread(nr))
    launch_kv<K, 2>(a, a.d, mask, launch_shape((nr + 1) / 2, 2, items2, reuse_of(P, st)), stream);
  else
    launch_kv<K, 1>(a, a.d, (1 << K) - 1, launch_shape(nr, 1, items2), stream);
}

template <int KO, int KI, int V, int DI, int M>
struct FusedMask {
  static void go(int mask, const FusedArgs<KO + KI>& a, const LaunchShape& sh, cudaStream_t stream) {
    if (mask == M)
      fused_kernel<KO, KI, V, DI, M><<<sh.grid, sh.block, 0, stream>>>(a);
    else
      FusedMask<KO, KI, V, DI, M - 1>::go(mask, a, sh, stream);
  }
};
template <int KO, int KI, int V, int DI>
struct FusedMask<KO, KI, V, DI, -1> {
  static void go(int, const FusedArgs<KO + KI>&, const LaunchShape&, cudaStream_t) { throw Error(HB_E_UNSUPPORTED, "bad blocking mask"); }
};

template <int KO, int KI, int V>
void launch_fused_d(const FusedArgs<KO + KI>& a, int di, int mask, const LaunchShape& sh, cudaStream_t stream) {
  constexpr int FULL = (1 << KI) - 1;
  if constexpr (V == 2) {
    switch (di) {
      case 2: FusedMask<KO, KI, V, 2, FULL>::go(mask, a, sh, stream); break;
      case 3: FusedMask<KO, KI, V, 3, FULL>::go(mask, a, sh, stream); break;
      case 4: FusedMask<KO, KI, V, 4, FULL>::go(mask, a, sh, stream); break;
      default: FusedMask<KO, KI, V, 0, FULL>::go(mask, a, sh, stream); break;
    }
  } else {
    fused_kernel<KO, KI, V, 0, FULL><<<sh.grid, sh.block, 0, stream>>>(a);  // tiny batches: run-time count, no reuse
  }
}

template <int KO, int KI>
void launch_fused(const Executor& ex, const Step& st, int si, int64_t RT, int nr, cudaStream_t stream) {
  constexpr int K = KO + KI;
  const Program& P = *ex.P;
  FusedArgs<K> a;
  memset(&a, 0, sizeof a);
  a.L = (int)st.levels.size();
  for (int l = 0; l < a.L; ++l) a.d[l] = st.levels[l].d;
  int mask = 0;
  for (int i = 0; i < K; ++i) {
    int level = 0;
    const StepInput& in = flat_input(st, i, &level);
    a.in[i] = in_arg(ex, st, si, i, RT);
    a.wst[i] = a.in[i].wstride;
    if (i >= KO && a.wst[i] != 0) mask |= 1 << (i - KO);
    a.lvl[i] = level;
    const int64_t mult = P.tables[in.table].kind == T_ROW ? RT : 1;
    for (int m = 0; m <= level; ++m) a.st[i][m] = (unsigned)(in.lstride[m] * mult);
  }
  a.C = (unsigned)st.C;
  a.n_out = (unsigned)st.n_out;
  a.U = (unsigned)st.U;
  a.n_red = (unsigned)st.n_red;
  const Table& out = P.tables[st.out];
  a.out_is_row = out.kind == T_ROW;
  a.out = a.out_is_row ? ex.d_arena + out.off * RT : ex.d_shared + out.off;
  a.nr = nr;
  a.RT = RT;
  a.scaled_in = scaled_input(ex, st, RT, a.scale, &a.n_scale);
  const int di = st.levels.back().d;
  const int64_t items2 = st.n_red * ((st.U + 1) / 2) * 2;
  if (a.scaled_in >= 0) {  // as in launch_k
    constexpr int FULL = (1 << KI) - 1;
    if (two_rows_per_thread(nr)) {
      const LaunchShape sh = launch_shape((nr + 1) / 2, 2, items2, reuse_of(P, st));
      fused_kernel<KO, KI, 2, 0, FULL, true><<<sh.grid, sh.block, 0, stream>>>(a);
    } else {
      const LaunchShape sh = launch_shape(nr, 1, items2);
      fused_kernel<KO, KI, 1, 0, FULL, true><<<sh.grid, sh.block, 0, stream>>>(a);
    }
    return;
  }
  if (two_rows_per_thread(nr))
    launch_fused_d<KO, KI, 2>(a, di, mask, launch_shape((nr + 1) / 2, 2, items2, reuse_of(P, st)), stream);
  else
    launch_fused_d<KO, KI, 1>(a, 0, (1 << KI) - 1, launch_shape(nr, 1, items2), stream);
}

template <int KO>
void launch_fused_ki(const Executor& ex, const Step& st, int si, int64_t RT, int nr, cudaStream_t stream) {
  switch ((int)st.levels.back().in.size()) {
    case 1: launch_fused<KO, 1>(ex, st, si, RT, nr, stream); break;
    case 2: launch_fused<KO, 2>(ex, st, si, RT, nr, stream); break;
    case 3: launch_fused<KO, 3>(ex, st, si, RT, nr, stream); break;
    default: throw Error(HB_E_UNSUPPORTED, "fused step: innermost level has 0 or more than 3 stored inputs");
  }
}

}  // namespace

// Strips of groups per row block for a generated step kernel: enough CTAs to fill the machine twice, and for a step that
// re-reads its per-row inputs so many that ONE wave of resident CTAs (296) covers only as many row blocks as fit ~64 MB of
// tables -- those rows then stay in L2 while their strips sweep them (C3's largest step read 9.4 GB for 1.2 GB of tables).
static unsigned strips_for(const Reuse& reuse, int64_t out_entries, unsigned bx, unsigned by, unsigned row_blocks, int64_t n_red) {
  // CTAs per SM a streaming step is cut into (HB_JIT_WAVES): its kernels are light in registers, more resident CTAs = more loads in flight
  static const int64_t stream_ctas = 148 * (int64_t)std::max(1, getenv("HB_JIT_WAVES") ? atoi(getenv("HB_JIT_WAVES")) : 8);
  const bool rereads = reuse.row_entries > 0 && reuse.row_loads >= 3 * reuse.row_entries;
  int64_t want = ((rereads ? 296 : stream_ctas) + row_blocks - 1) / row_blocks;
  if (rereads) {
    const int64_t slab = (reuse.row_entries + out_entries) * bx * 2 * 8;
    const int64_t blocks_per_wave = std::max<int64_t>(1, ((int64_t)64 << 20) / std::max<int64_t>(slab, 1));  // 16 .. 96 MB measured the same on C3
    want = std::max(want, (296 + blocks_per_wave - 1) / blocks_per_wave);
  }
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, (n_red + by - 1) / by), 65535));
}

static bool is_generic(const Step& st) {
  const int k = st.n_inputs(), s = (int)st.scalars.size();
  if (st.levels.size() > 1) return false;  // the compiler only fuses within MAXL levels / MAXK inputs, without scalars
  return !(k >= 1 && k <= MAXK && s <= MAXS);
}

// the descriptor-driven form of a single-level step (prodsum_generic_kernel)
static GenericArgs generic_args(const Executor& ex, const Step& st, int si, int64_t RT, int nr) {
  const Program& P = *ex.P;
  GenericArgs a;
  a.in = ex.d_gen_in + ex.gen_in_at[si];
  a.sc = ex.d_gen_sc + ex.gen_sc_at[si];
  a.k = st.n_inputs();
  a.n_sc = (int)st.scalars.size();
  a.d = st.levels[0].d;
  a.C = (unsigned)st.C;
  a.n_out = (unsigned)st.n_out;
  a.U = (unsigned)st.U;
  const Table& out = P.tables[st.out];
  a.out_is_row = out.kind == T_ROW;
  a.out = a.out_is_row ? ex.d_arena + out.off * RT : ex.d_shared + out.off;
  a.nr = nr;
  a.RT = RT;
  a.arg = st.levels[0].is_max ? ex.d_args + st.arg_off * RT : nullptr;
  a.scaled_in = scaled_input(ex, st, RT, a.scale, &a.n_scale);
  return a;
}

static void launch_prodsum(Executor& ex, const Step& st, int si, int64_t RT, int nr, cudaStream_t stream) {
  const int k = st.n_inputs(), s = (int)st.scalars.size();
  // the step has a specialised kernel: rows along x, so with one strip of groups per lane not for few-row tiles (a warp would
  // touch 16 or 32 distant strips); with the groups dealt round robin the lanes of a warp sit on adjacent groups
  if (!ex.jit_fn.empty() && ex.jit_fn[si] && (nr >= 64 || ex.jit_shape[si].interleave)) {
    const JitShape& shape = ex.jit_shape[si];
    // Row threads (two rows each) along threadIdx.x, strips of groups along threadIdx.y / blockIdx.x, row blocks along
    // blockIdx.y; enough of both to fill 148 SMs twice.  A step that reads its per-row inputs several times over gets narrow
    // row blocks (one warp of row threads): the strips of one row block are scheduled together and share those rows in L2
    // (C3's largest step: 9.4 GB of DRAM traffic for 1.2 GB of tables with 128-thread row blocks along the fast grid axis).
    const unsigned row_threads = (unsigned)(shape.V == 1 ? nr : (nr + 1) / 2);
    const Reuse reuse = reuse_of(*ex.P, st);
    unsigned bx = reuse.row_entries > 0 && reuse.row_loads >= 3 * reuse.row_entries ? 32 : 128;
    while (bx > 1 && bx / 2 >= row_threads) bx /= 2;
    while ((row_threads + bx - 1) / bx > 65535u) bx *= 2;  // gridDim.y limit
    const int64_t units = shape.groups / shape.inner;  // what the lanes share out: groups, or runs of `inner` groups
    const unsigned by = (unsigned)std::max<int64_t>(1, std::min<int64_t>(256 / bx, units));
    const unsigned gy = (row_threads + bx - 1) / bx;  // row blocks
    const unsigned gx = strips_for(reuse, ex.P->tables[st.out].size, bx, by, gy, units);
    const double *arena = ex.d_arena, *pots = ex.d_pots, *shared = ex.d_shared;
    double* arena_out = ex.d_arena;
    const int32_t* rowoff = ex.d_rowoff;
    long long rt = RT;
    void* params[] = {&arena, &arena_out, &pots, &shared, &rowoff, &nr, &rt};
    if (!jit_launch(ex.jit_fn[si], gx, gy, bx, by, params, stream)) throw Error(HB_E_CUDA, "launch of a specialised step kernel failed");
    return;
  }
  if (st.levels.size() > 1) {
    const int ki = (int)st.levels.back().in.size();
    if ((int)st.levels.size() > MAXL || s) throw Error(HB_E_UNSUPPORTED, "fused step exceeds the kernel limits");
    switch (k - ki) {
      case 0: launch_fused_ki<0>(ex, st, si, RT, nr, stream); break;
      case 1: launch_fused_ki<1>(ex, st, si, RT, nr, stream); break;
      case 2: launch_fused_ki<2>(ex, st, si, RT, nr, stream); break;
      case 3: launch_fused_ki<3>(ex, st, si, RT, nr, stream); break;
      case 4: launch_fused_ki<4>(ex, st, si, RT, nr, stream); break;
      default: throw Error(HB_E_UNSUPPORTED, "fused step: more than 4 stored inputs in the outer levels");
    }
  } else if (!is_generic(st)) {
    switch (k) {
      case 1: launch_k<1>(ex, st, si, RT, nr, stream); break;
      case 2: launch_k<2>(ex, st, si, RT, nr, stream); break;
      case 3: launch_k<3>(ex, st, si, RT, nr, stream); break;
      case 4: launch_k<4>(ex, st, si, RT, nr, stream); break;
      case 5: launch_k<5>(ex, st, si, RT, nr, stream); break;
      default: launch_k<6>(ex, st, si, RT, nr, stream); break;
    }
  } else {
    const GenericArgs a = generic_args(ex, st, si, RT, nr);
    LaunchShape sh = launch_shape(nr, 1, st.n_out);
    prodsum_generic_kernel<<<sh.grid, sh.block, 0, stream>>>(a);
  }
}

// (Re)builds everything that depends on the arena row stride RT: the arena itself, the slice-offset
// table, the RT-multiplied index maps, the generic-path descriptors and the output descriptors.
static void set_row_stride(Executor& ex, int64_t RT) {
  if (RT == ex.RT) return;
  const Program& P = *ex.P;
  ex.drop_graphs();
  free_dev(ex.d_arena);
  free_dev(ex.d_rowoff);
  free_dev(ex.d_maps_rt);
  free_dev(ex.d_args);
  if (P.arg_entries > 0) HB_CUDA(cudaMalloc(&ex.d_args, (size_t)(P.arg_entries * RT)));
  HB_CUDA(cudaMalloc(&ex.d_arena, std::max<int64_t>(P.row_entries, 1) * RT * sizeof(double)));
  HB_CUDA(cudaMalloc(&ex.d_rowoff, std::max<int64_t>((int64_t)P.slices.size(), 1) * RT * sizeof(int32_t)));
  ex.RT = RT;
  std::vector<uint32_t> scaled(ex.maps.size());
  for (size_t i = 0; i < scaled.size(); ++i) scaled[i] = (uint32_t)((uint64_t)ex.maps[i] * (uint64_t)RT);  // only row-table maps are read from this copy
  ex.d_maps_rt = upload(scaled);
  std::vector<InArg> gin;
  std::vector<TabRef> gsc;
  for (size_t si = 0; si < P.steps.size(); ++si) {
    const Step& st = P.steps[si];
    if (st.kind != STEP_PRODSUM || ex.gen_in_at[si] < 0) continue;  // descriptor-driven steps
    if (ex.gen_in_at[si] != (int64_t)gin.size() || ex.gen_sc_at[si] != (int64_t)gsc.size()) throw Error(HB_E_ARG, "internal: descriptor offsets");
    for (int i = 0; i < st.n_inputs(); ++i) gin.push_back(in_arg(ex, st, (int)si, i, RT));
    for (int id : st.scalars) gsc.push_back(tab_ref(ex, P.tables[id], RT));
  }
  free_dev(ex.d_gen_in);
  free_dev(ex.d_gen_sc);
  ex.d_gen_in = upload(gin);
  ex.d_gen_sc = upload(gsc);
  for (size_t j = 0; j < ex.out_steps.size(); ++j) {
    const Step& st = P.steps[ex.out_steps[j]];
    const Table& t = P.tables[st.result];
    ex.out_desc[j].t = tab_ref(ex, t, RT);
    ex.out_desc[j].slice = t.kind == T_POT ? t.slice_slot : -1;
  }
  free_dev(ex.d_out_desc);
  ex.d_out_desc = upload(ex.out_desc);
}

// Fills the potential pool: every CPT restricted to the fixed (split) states; stored ones are gathered on the
// host, procedural ones are generated on the device.  Also used by changeFactor (executor_refresh_potentials).
static void upload_potentials(Executor& ex) {
  const Program& P = *ex.P;
  const Network& net = *P.net;
  // potential pool: every CPT restricted to the fixed (split) states; stored ones are gathered on the host,
  // procedural ones are generated on the device
  {
    std::vector<double> staged;
    for (int v = 0; v < net.n(); ++v) {
      const std::vector<int>& sc = net.scope[v];
      std::vector<int> dim;
      std::vector<int64_t> fstride;
      int64_t full = 1, base = 0, entries = 1;
      for (int i = (int)sc.size() - 1; i >= 0; --i) {
        if (P.fixed_state[sc[i]] >= 0)
          base += full * P.fixed_state[sc[i]];
        else
          dim.insert(dim.begin(), net.dims[sc[i]]), fstride.insert(fstride.begin(), full), entries *= net.dims[sc[i]];
        full *= net.dims[sc[i]];
      }
      if (net.procedural(v)) {
        if (dim.size() > 40) throw Error(HB_E_UNSUPPORTED, "procedural CPT with more than 40 free variables");
        ProcFillArgs a;
        memset(&a, 0, sizeof a);
        a.dst = ex.d_pots + P.pot_off[v];
        a.entries = entries;
        a.seed = net.proc_seed[v];
        a.npar = (unsigned long long)(net.cpt_entries(v) / net.dims[v]);
        a.full_base = (unsigned long long)base;
        a.d = net.dims[v];
        a.n_free = (int)dim.size();
        for (size_t j = 0; j < dim.size(); ++j) a.dim[j] = dim[j], a.full_stride[j] = (unsigned long long)fstride[j];
        proc_fill_kernel<<<(unsigned)std::min<int64_t>((entries + BLOCK - 1) / BLOCK, 148 * 16), BLOCK>>>(a);
      } else {
        staged.resize((size_t)entries);
        std::vector<int> digit(dim.size(), 0);
        int64_t src = base;
        for (int64_t e = 0; e < entries; ++e) {
          staged[(size_t)e] = net.values[v][(size_t)src];
          for (int j = (int)dim.size() - 1; j >= 0; --j) {
            if (++digit[j] < dim[j]) {
              src += fstride[j];
              break;
            }
            src -= fstride[j] * (digit[j] - 1);
            digit[j] = 0;
          }
        }
        HB_CUDA(cudaMemcpy(ex.d_pots + P.pot_off[v], staged.data(), (size_t)entries * sizeof(double), cudaMemcpyHostToDevice));
      }
    }
    if (!P.extra_values.empty())
      HB_CUDA(cudaMemcpy(ex.d_pots + P.extra_off, P.extra_values.data(), P.extra_values.size() * sizeof(double), cudaMemcpyHostToDevice));
    HB_CUDA(cudaGetLastError());
  }
}

// batch-invariant steps: evaluated once per set of potentials, with a single pseudo-row
static void run_shared_steps(Executor& ex);

Executor* executor_create(const Program& P, int device, int64_t workspace_bytes) {
  HB_CUDA(cudaSetDevice(device));
  auto ex = std::make_unique<Executor>();
  ex->P = &P;
  ex->device = device;
  const Network& net = *P.net;
  HB_CUDA(cudaMalloc(&ex->d_pots, std::max<int64_t>(P.pot_entries, 1) * sizeof(double)));
  upload_potentials(*ex);
  HB_CUDA(cudaMalloc(&ex->d_shared, std::max<int64_t>(P.shared_entries, 1) * sizeof(double)));
  // index-map pool
  std::vector<uint32_t>& maps = ex->maps;
  for (size_t si = 0; si < P.steps.size(); ++si) {
    const Step& st = P.steps[si];
    ex->in_base.push_back((int64_t)ex->hi_off.size());
    if (st.kind == STEP_PRODSUM) {
      if (st.n_out >= ((int64_t)1 << 31)) throw Error(HB_E_UNSUPPORTED, "a table with 2^31 or more entries");
      for (int i = 0; i < st.n_inputs(); ++i) {
        const StepInput& in = flat_input(st, i);
        ex->hi_off.push_back((int64_t)maps.size());
        maps.insert(maps.end(), in.hi.begin(), in.hi.end());
        ex->lo_off.push_back((int64_t)maps.size());
        maps.insert(maps.end(), in.lo.begin(), in.lo.end());
        const Table& t = P.tables[in.table];
        if (t.kind == T_ROW) ex->max_row_table = std::max(ex->max_row_table, t.size);
      }
    } else {
      if ((int)st.obs_var.size() > MAXOBS) throw Error(HB_E_UNSUPPORTED, "more than 8 observed variables in one query");
      OutDesc d;
      memset(&d, 0, sizeof d);
      d.n_full = (int)st.n_full;
      d.out_col = st.out_col;
      d.map_off = (int)maps.size();
      maps.insert(maps.end(), st.full_map.begin(), st.full_map.end());
      // the stored entries in the order the full layout visits them (norm is a left fold over that order)
      d.order_off = (int)maps.size();
      int n_phys = 0;
      for (int64_t f = 0; f < st.n_full; ++f) {
        bool first_state = true;
        for (size_t j = 0; j < st.obs_var.size(); ++j) first_state &= ((f / st.obs_stride[j]) % st.obs_dim[j]) == 0;
        if (first_state) maps.push_back((uint32_t)st.full_map[f]), ++n_phys;
      }
      d.n_phys = n_phys;
      d.n_obs = (int)st.obs_var.size();
      d.gather = st.gather ? 1 : 0;
      for (int j = 0; j < d.n_obs; ++j) d.obs_var[j] = st.obs_var[j], d.obs_stride[j] = (int)st.obs_stride[j], d.obs_dim[j] = (int)st.obs_dim[j];
      ex->out_desc.push_back(d);
      ex->out_steps.push_back((int)si);
    }
  }
  ex->d_maps = upload(maps);
  if (!P.max_steps.empty()) {  // traceback descriptors, last maximised variable first
    std::vector<int32_t> desc;
    for (size_t j = P.max_steps.size(); j-- > 0;) {
      const Program::MaxStep& ms = P.max_steps[j];
      desc.push_back(ms.slot);
      desc.push_back((int32_t)ms.dep_slot.size());
      desc.push_back((int32_t)(uint32_t)(ms.arg_off & 0xffffffffll));
      desc.push_back((int32_t)(ms.arg_off >> 32));
      for (size_t k = 0; k < ms.dep_slot.size(); ++k) desc.push_back(ms.dep_slot[k]), desc.push_back((int32_t)ms.dep_stride[k]);
    }
    ex->d_trace = upload(desc);
    ex->n_trace = (int)P.max_steps.size();
  }
  {
    const int LANES = std::max(1, std::min(64, getenv("HB_LANES") ? atoi(getenv("HB_LANES")) : 16));
    ex->n_lanes = LANES;
    struct Access {
      std::vector<std::pair<int64_t, int64_t>> reads;
      std::pair<int64_t, int64_t> write;
    };
    std::vector<Access> acc;
    auto overlap = [](const std::pair<int64_t, int64_t>& a, const std::pair<int64_t, int64_t>& b) {
      return a.first < b.first + b.second && b.first < a.first + a.second;
    };
    std::vector<int> lane_last(LANES, -1);
    int rr = 0;
    for (const Step& st : P.steps) {
      if (st.kind != STEP_PRODSUM || st.shared) continue;
      Access a;
      a.write = {P.tables[st.out].off, P.tables[st.out].size};
      for (int i = 0; i < st.n_inputs(); ++i) {
        const Table& t = P.tables[flat_input(st, i).table];
        if (t.kind == T_ROW) a.reads.push_back({t.off, t.size});
      }
      for (int id : st.scalars)
        if (P.tables[id].kind == T_ROW) a.reads.push_back({P.tables[id].off, P.tables[id].size});
      const int s = (int)acc.size();
      std::vector<int> deps;
      int producer = -1;
      for (int t = s - 1; t >= 0; --t) {
        bool raw = false, war = overlap(acc[t].write, a.write);
        for (auto& r : a.reads) raw = raw || overlap(acc[t].write, r);
        for (auto& r : acc[t].reads) war = war || overlap(r, a.write);
        if (raw || war) deps.push_back(t);
        if (raw && producer < 0) producer = t;
      }
      // the steps of one reference operation group (a message, a posterior) form a chain: same stream
      (void)producer;
      (void)rr;
      const int lane = st.group >= 0 ? st.group % LANES : 0;
      ex->step_deps.push_back(deps);
      ex->step_lane.push_back(lane);
      acc.push_back(std::move(a));
      ++ex->n_row_steps;
    }
  }
  {  // launch nodes
    const int n = ex->n_row_steps;
    std::vector<int> row_si;
    for (size_t si = 0; si < P.steps.size(); ++si)
      if (P.steps[si].kind == STEP_PRODSUM && !P.steps[si].shared) row_si.push_back((int)si);
    std::vector<int> level(n, 0);
    int n_levels = 0;
    for (int s = 0; s < n; ++s) {
      for (int dep : ex->step_deps[s]) level[s] = std::max(level[s], level[dep] + 1);
      n_levels = std::max(n_levels, level[s] + 1);
    }
    std::vector<int> node_of(n, -1);
    for (int l = 0; l < n_levels; ++l)
      for (int s = 0; s < n; ++s) {
        if (level[s] != l) continue;
        Executor::Node one;
        one.si = row_si[s];
        one.lane = ex->step_lane[s];
        node_of[s] = (int)ex->nodes.size();
        ex->nodes.push_back(std::move(one));
      }
    for (int s = 0; s < n; ++s) {
      std::vector<int>& deps = ex->nodes[node_of[s]].deps;
      for (int dep : ex->step_deps[s]) deps.push_back(node_of[dep]);
      std::sort(deps.begin(), deps.end());
    }
    ex->gen_in_at.assign(P.steps.size(), -1);
    ex->gen_sc_at.assign(P.steps.size(), -1);
    int64_t gi = 0, gs = 0;
    for (size_t si = 0; si < P.steps.size(); ++si) {
      const Step& st = P.steps[si];
      if (st.kind != STEP_PRODSUM || !is_generic(st)) continue;
      ex->gen_in_at[si] = gi, ex->gen_sc_at[si] = gs;
      gi += (int64_t)st.n_inputs(), gs += (int64_t)st.scalars.size();
    }
  }
  // column chunks of the normalise/expand kernel
  {
    int max_w = 1;
    constexpr int MAX_TILE_COLS = 200 * 1024 / (OUT_ROWS * 8) - 1;
    for (size_t j = 0; j < ex->out_desc.size(); ++j) {
      const OutDesc& d = ex->out_desc[j];
      const bool wide = d.n_full > MAX_TILE_COLS;
      if (wide || ex->out_chunks.empty() || ex->out_chunks.back().direct || ex->out_chunks.back().width + d.n_full > OUT_COLS)
        ex->out_chunks.push_back(OutChunk{(int)j, (int)j, (int)d.out_col, 0, wide ? 1 : 0});
      ex->out_chunks.back().desc_end = (int)j + 1;
      ex->out_chunks.back().width += d.n_full;
      if (!wide) max_w = std::max(max_w, ex->out_chunks.back().width);
    }
    ex->out_tile_bytes = OUT_ROWS * (max_w + 1) * (int)sizeof(double);
    HB_CUDA(cudaFuncSetAttribute(output_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    ex->d_out_chunks = upload(ex->out_chunks);
  }
  // evidence preparation tables
  std::vector<int8_t> obs(P.is_observed.begin(), P.is_observed.end());
  for (int v = 0; v < net.n(); ++v)
    if (P.fixed_state[v] >= 0) obs[v] = (int8_t)(2 + P.fixed_state[v]);
  std::vector<int32_t> dims(net.dims.begin(), net.dims.end()), sl_begin = {0}, sl_var, sl_stride;
  for (const Slice& s : P.slices) {
    for (size_t j = 0; j < s.vars.size(); ++j) sl_var.push_back(s.vars[j]), sl_stride.push_back((int32_t)s.strides[j]);
    sl_begin.push_back((int32_t)sl_var.size());
  }
  ex->d_observed = upload(obs);
  ex->d_dims = upload(dims);
  ex->d_sl_begin = upload(sl_begin);
  ex->d_sl_var = upload(sl_var);
  ex->d_sl_stride = upload(sl_stride);
  // the error flag, followed (8 bytes in) by 64 cycle counters of the on-chip kernel's HB_TREE_PROFILE instrumentation
  HB_CUDA(cudaMalloc(&ex->d_flag, 8 + 64 * sizeof(unsigned long long)));
  HB_CUDA(cudaMemset(ex->d_flag, 0, 8 + 64 * sizeof(unsigned long long)));
  if (workspace_bytes <= 0) {
    size_t free_b = 0, total_b = 0;
    HB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    workspace_bytes = std::min<int64_t>((int64_t)(free_b * 0.6), (int64_t)24 << 30);
    // the ranks of a split clique all-reduce whole tiles: they must cut a batch into the same tiles whatever memory each has free
    // (half of the device's TOTAL memory, which is the same on every rank, not a share of what happens to be free)
    if (P.n_reduce > 0 || P.n_gather > 0) workspace_bytes = std::min<int64_t>((int64_t)(total_b / 2), (int64_t)64 << 30);
  }
  ex->workspace = workspace_bytes;
  if (P.n_gather > 0) {
    std::vector<int> ob, tb = {0};
    for (const OutDesc& d : ex->out_desc)
      if (d.gather) ob.push_back((int)d.out_col), tb.push_back(tb.back() + d.n_full);
    ob.push_back(0);
    ex->gather_blocks = (int)tb.size() - 1;
    ex->gather_width = tb.back();
    ex->d_gather_out = upload(ob);
    ex->d_gather_tmp = upload(tb);
  }
  run_shared_steps(*ex);
  {
    std::string why;
    ex->tree_eligible = tree_plan(P, ex->tree, &why);
    if (!ex->tree_eligible && getenv("HB_JIT_VERBOSE") && atoi(getenv("HB_JIT_VERBOSE"))) fprintf(stderr, "[hbayes jit] no on-chip kernel: %s\n", why.c_str());
  }
  return ex.release();
}

static void run_shared_steps(Executor& ex) {
  const Program& P = *ex.P;
  if (!ex.RT) set_row_stride(ex, 2);  // shared tables do not depend on the row stride; the descriptors need one (two rows: always fits)
  for (size_t si = 0; si < P.steps.size(); ++si) {
    const Step& st = P.steps[si];
    if (st.kind == STEP_PRODSUM && st.shared) launch_prodsum(ex, st, (int)si, ex.RT, 1, 0);
  }
  HB_CUDA(cudaGetLastError());
  HB_CUDA(cudaDeviceSynchronize());
}

void executor_refresh_potentials(Executor* ex) {
  HB_CUDA(cudaSetDevice(ex->device));
  HB_CUDA(cudaDeviceSynchronize());
  if (ex->jit_elided && !jit_zero_add_elidable(*ex->P)) {
    // the generated kernels left out the `0 +` of every sum on the premise that no potential is negative, -0.0 or NaN
    // (jit.cpp): the new factor breaks it, so they go and the next specialisation emits the adds
    ex->pending.reset();
    ex->tree_fn = nullptr;
    ex->jit_fn.clear();
    jit_destroy(ex->jit);
    ex->jit = nullptr;
    ex->drop_graphs();
    ex->jit_tried = false;
    ex->jit_elided = false;
  }
  upload_potentials(*ex);  // same addresses: captured graphs stay valid
  run_shared_steps(*ex);
}

void executor_destroy(Executor* ex) { delete ex; }
int executor_device(const Executor* ex) { return ex->device; }

static bool jit_enabled() { return !(getenv("HB_JIT") && atoi(getenv("HB_JIT")) == 0); }
static bool tree_enabled() { return jit_enabled() && !(getenv("HB_TREE") && atoi(getenv("HB_TREE")) == 0); }

// the steps worth a specialised kernel of their own (when the program has no on-chip kernel)
static std::vector<int> steps_to_specialise(const Program& P) {
  std::vector<std::pair<int64_t, int>> cand;  // (loads per row on the templated kernels, step)
  for (size_t si = 0; si < P.steps.size(); ++si) {
    const Step& st = P.steps[si];
    if (!jit_eligible(P, st)) continue;
    int64_t spec = 0, templ = 0;
    jit_load_counts(P, st, &spec, &templ);
    // measured on C2: also steps without fewer loads gain (no index maps, immediate strides); tiny steps do not matter
    if (templ >= 32) cand.push_back({templ, (int)si});
  }
  std::sort(cand.rbegin(), cand.rend());
  std::vector<int> steps;
  int64_t budget = getenv("HB_JIT_BUDGET") ? atoll(getenv("HB_JIT_BUDGET")) : 8000;  // bounds the compilation time: C2 needs 4.6 K (49 kernels, 5.5 s)
  // C5 (340 light steps): 96 kernels 20.8 ms per pass, all 334 eligible steps 19.3 ms (12 s of NVRTC once, then the cubin cache)
  const size_t max_kernels = getenv("HB_JIT_MAX_KERNELS") ? (size_t)atoll(getenv("HB_JIT_MAX_KERNELS")) : 384;
  // the steps with a register block over several output variables (re-reading steps, jit_step_shape) are the heaviest of a
  // program and few: they do not draw on the budget of the others
  int blocked_left = getenv("HB_JIT_MAX_BLOCKED") ? atoi(getenv("HB_JIT_MAX_BLOCKED")) : 24;
  for (auto& c : cand) {
    const JitShape sh = jit_step_shape(P, P.steps[c.second]);
    if (steps.size() >= max_kernels) continue;
    if (sh.custom) {
      if (blocked_left <= 0 && sh.n_loop > 0) continue;
      if (blocked_left > 0) {
        --blocked_left;
        steps.push_back(c.second);
        continue;
      }
    }
    const int64_t ops = sh.ops;
    if (ops < 0 || ops > budget) continue;
    budget -= ops;
    steps.push_back(c.second);
  }
  std::sort(steps.begin(), steps.end());
  return steps;
}

// starts the compilation of the program's kernels on a background thread (registry / disk cache hits return at once)
static void start_specialise(Executor* ex) {
  if (ex->jit_tried || !jit_enabled()) return;
  ex->jit_tried = true;
  const Program& P = *ex->P;
  auto pend = std::make_unique<Executor::Pending>();
  ex->jit_elided = jit_zero_add_elidable(P);
  std::string src;
  if (ex->tree_eligible && tree_enabled()) {
    pend->is_tree = true;
    src = tree_source(P, ex->tree, &ex->tree_loads, &ex->tree_stores);
  } else {
    pend->steps = steps_to_specialise(P);
    if (pend->steps.empty()) return;
    src = jit_source(P, pend->steps);
  }
  auto key = pend->key;
  auto why = pend->why;
  const bool fma = ex->fma;
  pend->cubin = std::async(std::launch::async, [src = std::move(src), key, why, fma]() {
    int origin = -1;
    auto cubin = jit_compile(src, fma, why.get(), key.get(), &origin);
    if (cubin) *why = std::to_string(origin);
    return cubin;
  });
  ex->pending = std::move(pend);
}

// installs a finished compilation (wait = false: only if it is ready)
static void finish_specialise(Executor* ex, bool wait) {
  if (!ex->pending) return;
  if (!wait && ex->pending->cubin.wait_for(std::chrono::seconds(0)) != std::future_status::ready) return;
  std::unique_ptr<Executor::Pending> pend = std::move(ex->pending);
  std::shared_ptr<const std::string> cubin = pend->cubin.get();
  if (!cubin) return;  // NVRTC unavailable or the compilation failed: the precompiled kernels stay in charge
  const Program& P = *ex->P;
  HB_CUDA(cudaSetDevice(ex->device));
  HB_CUDA(cudaDeviceSynchronize());  // no launch of this executor is in flight while its kernels are swapped
  std::string why;
  JitModule* mod = jit_load(*cubin, *pend->key, &why);
  if (!mod) return;
  ex->jit_origin = atoi(pend->why->c_str());
  if (pend->is_tree) {
    void* fn = jit_function(mod, "hb_tree");
    if (!fn || !jit_set_dynamic_smem(fn, (int)ex->tree.smem_bytes)) {
      jit_destroy(mod);
      return;
    }
    int per_sm = jit_occupancy(fn, ex->tree.T, ex->tree.smem_bytes);
    if (per_sm <= 0) per_sm = 1;
    cudaDeviceProp prop;
    HB_CUDA(cudaGetDeviceProperties(&prop, ex->device));
    ex->tree_grid = per_sm * prop.multiProcessorCount;
    ex->tree_fn = fn;
    if (getenv("HB_JIT_VERBOSE") && atoi(getenv("HB_JIT_VERBOSE")))
      fprintf(stderr, "[hbayes jit] on-chip kernel hb_tree: %d rows x %d threads per CTA, %d CTAs per SM, %d levels, %lld arena entries per row, %zu bytes of shared memory\n",
              ex->tree.R, ex->tree.T, per_sm, ex->tree.n_levels, (long long)ex->tree.arena_entries, ex->tree.smem_bytes);
  } else {
    ex->jit_fn.assign(P.steps.size(), nullptr);
    ex->jit_shape.assign(P.steps.size(), JitShape());
    int n_blocked = 0;
    for (int si : pend->steps) {
      void* fn = jit_function(mod, "hb_step_" + std::to_string(si));
      if (!fn) {
        ex->jit_fn.clear();
        jit_destroy(mod);
        return;
      }
      ex->jit_fn[si] = fn;
      ex->jit_shape[si] = jit_step_shape(P, P.steps[si]);
      n_blocked += ex->jit_shape[si].custom;
    }
    ex->jit_blocked = n_blocked;
    if (getenv("HB_JIT_VERBOSE") && atoi(getenv("HB_JIT_VERBOSE")))
      fprintf(stderr, "[hbayes jit] %zu specialised kernels, %d of them with a register block over several output variables\n", pend->steps.size(), n_blocked);
  }
  jit_destroy(ex->jit);
  ex->jit = mod;
  ex->drop_graphs();
}

int executor_specialise(Executor* ex) {
  start_specialise(ex);
  finish_specialise(ex, true);
  return executor_specialised_kernels(ex);
}

void executor_run(Executor* ex, int64_t n_rows, const int8_t* d_evidence, double* d_out, void* stream_, RunStats* stats, bool normalise,
                  int8_t* d_mpe_states_) {
  signed char* d_mpe_states = (signed char*)d_mpe_states_;
  if (n_rows <= 0) return;
  cudaStream_t stream = (cudaStream_t)stream_;
  const Program& P = *ex->P;
  HB_CUDA(cudaSetDevice(ex->device));
  const int n_vars = P.net->n();
  if ((P.n_reduce > 0 || P.n_gather > 0) && !ex->comm)
    throw Error(HB_E_NCCL, "this program sums a split clique over the ranks: the handle has no communicator (hb_jt_compile_split_nccl)");
  {  // Kernels generated for this program (jit.cpp): the on-chip kernel, or per-step kernels for its heaviest steps.  NVRTC
     // takes seconds, so the compilation starts in the background once the program has seen >= 2^31 joint entries in one
     // call (C2: 140 K rows) or 2^35 in all, and is installed by the first later call that finds it finished; cached
     // cubins (this process, disk) are found at once.  executor_specialise() does the same synchronously.
    static const double jit_min_work = getenv("HB_JIT_MIN_WORK") ? atof(getenv("HB_JIT_MIN_WORK")) : 2147483648.0;
    static const bool jit_sync = getenv("HB_JIT_SYNC") && atoi(getenv("HB_JIT_SYNC")) != 0;
    const double work = (double)n_rows * (double)std::max<int64_t>(P.stat_prod, 1);
    ex->work_seen += work;
    if (work >= jit_min_work || ex->work_seen >= 16.0 * jit_min_work) start_specialise(ex);
    finish_specialise(ex, jit_sync);
  }
  if (ex->tree_fn) {  // the whole pass in one launch: no row arena, no tiles
    ex->events_used = 0;
    if (ex->timing)
      for (int i = 0; i < 2; ++i) HB_CUDA(cudaEventRecord(ex->next_event(), stream));
    const double *pots = ex->d_pots, *shared = ex->d_shared;
    long long nrows = n_rows;
    int norm = normalise ? 1 : 0;
    int* flag = ex->d_flag;
    void* params[] = {&pots, &shared, &d_evidence, &d_out, &nrows, &flag, &norm};
    const int64_t n_tiles = (n_rows + ex->tree.R - 1) / ex->tree.R;
    const unsigned grid = (unsigned)std::min<int64_t>(n_tiles, ex->tree_grid);
    if (!jit_launch_smem(ex->tree_fn, grid, (unsigned)ex->tree.T, (unsigned)ex->tree.smem_bytes, params, stream))
      throw Error(HB_E_CUDA, "launch of the on-chip kernel failed");
    if (ex->timing)
      for (int i = 0; i < 2; ++i) HB_CUDA(cudaEventRecord(ex->next_event(), stream));
    if (stats) {
      stats->launches = 1;
      stats->tiles = n_tiles;
      stats->rows_per_tile = ex->tree.R;
      stats->arena_bytes = 0;
    }
    return;
  }
  // rows per tile: as many as the workspace allows, in multiples of 32 (one warp of rows), and few enough
  // that every row-table offset (entry * RT + row) fits 32 bits
  const int64_t per_row = std::max<int64_t>(P.row_entries, 1) * 8 + (int64_t)P.slices.size() * 4 + P.arg_entries;
  // tiles are multiples of 64 rows; a batch smaller than that (huge per-row tables, few rows) gets an even stride
  auto round_rows = [](int64_t r) { return r >= 64 ? r / 64 * 64 : r / 2 * 2; };
  int64_t cap = round_rows(ex->workspace / per_row);
  cap = std::min<int64_t>(cap, round_rows((((int64_t)1 << 32) - 1) / (ex->max_row_table + 1)));
  cap = std::min<int64_t>(cap, (int64_t)1 << 22);
  if (cap < 2) throw Error(HB_E_OOM, "the row arena of two rows (" + std::to_string(per_row * 2) + " bytes) exceeds the workspace or 32-bit offsets");
  const int64_t RT = std::min<int64_t>(cap, n_rows >= 64 ? (n_rows + 63) / 64 * 64 : (n_rows + 1) / 2 * 2);
  if (RT > ex->RT || ex->RT > 4 * RT || ex->RT > cap) {
    HB_CUDA(cudaStreamSynchronize(stream));
    set_row_stride(*ex, RT);
  }
  const int64_t stride = ex->RT;
  int64_t launches = 0, tiles = 0;
  ex->events_used = 0;
  auto mark = [&]() {
    if (ex->timing) HB_CUDA(cudaEventRecord(ex->next_event(), stream));
  };
  for (int64_t r0 = 0; r0 < n_rows; r0 += stride, ++tiles) {
    const int nr = (int)std::min<int64_t>(stride, n_rows - r0);
    mark();
    PrepArgs pa{d_evidence + r0 * n_vars, nr, n_vars, (int)P.slices.size(), stride, ex->d_observed, ex->d_dims,
                ex->d_sl_begin, ex->d_sl_var, ex->d_sl_stride, ex->d_rowoff, ex->d_flag};
    {
      const dim3 pgrid((((nr + 63) & ~63) + BLOCK - 1) / BLOCK, (unsigned)std::max<size_t>(1, (P.slices.size() + PREP_CHUNK - 1) / PREP_CHUNK));
      prep_rows_kernel<<<pgrid, BLOCK, 0, stream>>>(pa);
    }
    ++launches;
    mark();
    const std::vector<Executor::Node>& nodes = ex->nodes;
    auto launch_steps = [&](cudaStream_t on) {
      for (const Executor::Node& n : nodes) launch_prodsum(*ex, P.steps[n.si], n.si, stride, nr, on);
    };
    static const bool graphs_on = !(getenv("HB_GRAPHS") && atoi(getenv("HB_GRAPHS")) == 0);
    // Split clique: a table that holds this GPU's partial sum over the split variables is all-reduced right after the step
    // that produces it -- one exchange of |table| doubles per row and per message -- unless only the final normalisation
    // reads it: those (the posteriors' own results) are exchanged together after the last step, as one NCCL group.
    std::vector<const Table*> late;
    bool early_reduce = false;
    for (const Executor::Node& n : nodes) {
      const Table& t = P.tables[P.steps[n.si].out];
      if (!t.reduce) continue;
      bool only_output = true;
      for (const Step& other : P.steps)
        if (other.kind == STEP_PRODSUM && !other.shared)
          for (const Level& l : other.levels)
            for (const StepInput& in : l.in) only_output = only_output && in.table != P.steps[n.si].out;
      if (only_output)
        late.push_back(&t);
      else
        early_reduce = true;
    }
    if (early_reduce) {  // exchanges between the steps: plain launches in dependency order, the same on every rank
      for (const Executor::Node& n : nodes) {
        const Step& st = P.steps[n.si];
        launch_prodsum(*ex, st, n.si, stride, nr, stream);
        const Table& t = P.tables[st.out];
        if (t.reduce && std::find(late.begin(), late.end(), &t) == late.end())
          nccl_allreduce_sum(ex->comm, ex->d_arena + t.off * stride, (size_t)(t.size * stride), stream);
      }
    } else if (graphs_on && nodes.size() >= 16) {
      auto it = ex->graphs.find(nr);
      if (it == ex->graphs.end()) {
        if (!ex->capture_stream) HB_CUDA(cudaStreamCreateWithFlags(&ex->capture_stream, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        while ((int)ex->side_streams.size() < ex->n_lanes - 1) {
          cudaStream_t st;
          HB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
          ex->side_streams.push_back(st);
        }
        while ((int)ex->step_events.size() < ex->n_row_steps + ex->n_lanes) {
          cudaEvent_t e;
          HB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
          ex->step_events.push_back(e);
        }
        static const bool overlap_on = !(getenv("HB_GRAPH_LANES") && atoi(getenv("HB_GRAPH_LANES")) == 0);
        auto lane_stream = [&](int lane) { return lane == 0 || !overlap_on ? ex->capture_stream : ex->side_streams[lane - 1]; };
        HB_CUDA(cudaStreamBeginCapture(ex->capture_stream, cudaStreamCaptureModeThreadLocal));
        try {
          // fork: every side stream joins the capture
          cudaEvent_t fork = ex->step_events[ex->n_row_steps];
          HB_CUDA(cudaEventRecord(fork, ex->capture_stream));
          if (overlap_on)
            for (cudaStream_t st : ex->side_streams) HB_CUDA(cudaStreamWaitEvent(st, fork, 0));
          for (size_t ni = 0; ni < nodes.size(); ++ni) {
            const Executor::Node& n = nodes[ni];
            cudaStream_t on = lane_stream(n.lane);
            if (overlap_on)
              for (int dep : n.deps)
                if (nodes[dep].lane != n.lane) HB_CUDA(cudaStreamWaitEvent(on, ex->step_events[dep], 0));
            launch_prodsum(*ex, P.steps[n.si], n.si, stride, nr, on);
            if (overlap_on) HB_CUDA(cudaEventRecord(ex->step_events[ni], on));
          }
          // join
          if (overlap_on)
            for (size_t l = 0; l < ex->side_streams.size(); ++l) {
              cudaEvent_t e = ex->step_events[ex->n_row_steps + 1 + l];
              HB_CUDA(cudaEventRecord(e, ex->side_streams[l]));
              HB_CUDA(cudaStreamWaitEvent(ex->capture_stream, e, 0));
            }
        } catch (...) {
          cudaStreamEndCapture(ex->capture_stream, &graph);
          if (graph) cudaGraphDestroy(graph);
          throw;
        }
        HB_CUDA(cudaStreamEndCapture(ex->capture_stream, &graph));
        cudaGraphExec_t exec = nullptr;
        const cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) throw Error(HB_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
        if (ex->graphs.size() >= 8) ex->drop_graphs();  // ragged tails: keep the cache small
        it = ex->graphs.emplace(nr, exec).first;
      }
      HB_CUDA(cudaGraphLaunch(it->second, stream));
    } else {
      launch_steps(stream);
    }
    if (!late.empty()) {
      nccl_group_start();
      for (const Table* t : late) nccl_allreduce_sum(ex->comm, ex->d_arena + t->off * stride, (size_t)(t->size * stride), stream);
      nccl_group_end();
    }
    launches += (int64_t)nodes.size();
    mark();
    if (!ex->out_desc.empty()) {
      OutArgs oa{ex->d_out_desc, ex->d_out_chunks, ex->d_maps, ex->d_rowoff, d_evidence + r0 * n_vars,
                 d_out + r0 * P.out_width, nr, n_vars, stride, P.out_width, normalise ? 1 : 0};
      dim3 grid((nr + OUT_ROWS - 1) / OUT_ROWS, (unsigned)ex->out_chunks.size());
      output_kernel<<<grid, BLOCK, ex->out_tile_bytes, stream>>>(oa);
      ++launches;
      if (P.n_gather > 0) {  // queries that mention a split variable: sum the ranks' slices, then normalise
        const size_t need = (size_t)nr * ex->gather_width;
        if (need > ex->gather_cap) {
          HB_CUDA(cudaStreamSynchronize(stream));
          free_dev(ex->d_gather);
          HB_CUDA(cudaMalloc(&ex->d_gather, need * sizeof(double)));
          ex->gather_cap = need;
        }
        GatherArgs ga{d_out + r0 * P.out_width, ex->d_gather, ex->d_gather_out, ex->d_gather_tmp, ex->gather_blocks, 0, normalise ? 1 : 0,
                      nr, P.out_width, ex->gather_width};
        const unsigned gb = (unsigned)(((int64_t)nr * ex->gather_blocks + BLOCK - 1) / BLOCK);
        gather_blocks_kernel<<<gb, BLOCK, 0, stream>>>(ga);
        nccl_allreduce_sum(ex->comm, ex->d_gather, need, stream);
        ga.backward = 1;
        gather_blocks_kernel<<<gb, BLOCK, 0, stream>>>(ga);
        launches += 2;
      }
    }
    if (d_mpe_states && ex->n_trace > 0) {
      TraceArgs ta{ex->d_trace, ex->n_trace, (int)P.mpe_r.size(), nr, ex->d_args, stride, d_mpe_states + r0 * (int64_t)P.mpe_r.size()};
      mpe_trace_kernel<<<(nr + BLOCK - 1) / BLOCK, BLOCK, 0, stream>>>(ta);
      ++launches;
    }
    mark();
  }
  HB_CUDA(cudaGetLastError());
  if (stats) {
    stats->launches = launches;
    stats->tiles = tiles;
    stats->rows_per_tile = stride;
    stats->arena_bytes = std::max<int64_t>(P.row_entries, 1) * stride * 8;
  }
}

void executor_normalise(Executor* ex, int64_t n_rows, double* d_out, void* stream_) {
  if (n_rows <= 0) return;
  HB_CUDA(cudaSetDevice(ex->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!ex->d_norm_cols) {
    std::vector<int> cols;
    for (const OutDesc& d : ex->out_desc) cols.push_back((int)d.out_col);
    cols.push_back((int)ex->P->out_width);
    ex->d_norm_cols = upload(cols);
  }
  NormArgs a{d_out, ex->d_norm_cols, (int)ex->out_desc.size(), n_rows, ex->P->out_width};
  const int64_t threads = n_rows * a.n_blocks;
  normalise_kernel<<<(unsigned)((threads + BLOCK - 1) / BLOCK), BLOCK, 0, stream>>>(a);
  HB_CUDA(cudaGetLastError());
}

void executor_set_timing(Executor* ex, bool on) { ex->timing = on; }

int executor_specialised_kernels(const Executor* ex) {
  if (ex->tree_fn) {  // the on-chip kernel covers every per-row step (a program without evidence has none: everything is
    int n = 0;        // batch-invariant and the kernel only normalises and writes -- still one generated kernel)
    for (int lv : ex->tree.level) n += lv >= 0;
    return std::max(n, 1);
  }
  int n = 0;
  for (void* f : ex->jit_fn) n += f != nullptr;
  return n;
}
bool executor_on_chip(const Executor* ex) { return ex->tree_fn != nullptr; }
void executor_set_comm(Executor* ex, void* comm) { ex->comm = comm; }
void executor_fill_fixed(Executor* ex, int8_t* d_evidence, int64_t n_rows, void* stream) {
  if (n_rows <= 0) return;
  HB_CUDA(cudaSetDevice(ex->device));
  const int n_vars = ex->P->net->n();
  FillArgs a{d_evidence, ex->d_observed, n_rows, n_vars, ex->d_flag};
  fill_fixed_kernel<<<(unsigned)((n_rows * n_vars + BLOCK - 1) / BLOCK), BLOCK, 0, (cudaStream_t)stream>>>(a);
  HB_CUDA(cudaGetLastError());
}
void executor_tree_profile(Executor* ex, void* stream, int64_t out64[64]) {
  HB_CUDA(cudaSetDevice(ex->device));
  HB_CUDA(cudaMemcpyAsync(out64, (const char*)ex->d_flag + 8, 64 * sizeof(int64_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  HB_CUDA(cudaMemsetAsync((char*)ex->d_flag + 8, 0, 64 * sizeof(int64_t), (cudaStream_t)stream));
  HB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
}
int executor_jit_origin(const Executor* ex) { return ex->jit_origin; }
const char* executor_jit_key(const Executor* ex) { return ex->jit ? jit_module_key(ex->jit).c_str() : ""; }
void executor_set_fma(Executor* ex, bool on) { ex->fma = on; }
void executor_tree_info(const Executor* ex, int64_t out10[10]) {
  const TreePlan& t = ex->tree;
  const int64_t v[10] = {ex->tree_eligible ? 1 : 0, t.R, t.T, t.n_levels, t.arena_entries, (int64_t)t.smem_bytes, t.ops_mul, t.ops_add,
                         ex->tree_loads, ex->tree_stores};
  for (int i = 0; i < 10; ++i) out10[i] = v[i];
}

void executor_last_timing(Executor* ex, void* stream, double out4[4]) {
  HB_CUDA(cudaSetDevice(ex->device));
  HB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  out4[0] = out4[1] = out4[2] = out4[3] = 0.0;
  for (size_t i = 0; i + 3 < ex->events_used; i += 4) {
    float ms = 0;
    for (int j = 0; j < 3; ++j) {
      HB_CUDA(cudaEventElapsedTime(&ms, ex->events[i + j], ex->events[i + j + 1]));
      out4[j] += ms;
    }
    HB_CUDA(cudaEventElapsedTime(&ms, ex->events[i], ex->events[i + 3]));
    out4[3] += ms;
  }
}

void executor_check(Executor* ex, void* stream) {
  HB_CUDA(cudaSetDevice(ex->device));
  int flag = 0;
  HB_CUDA(cudaMemcpyAsync(&flag, ex->d_flag, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  HB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (flag) {
    HB_CUDA(cudaMemsetAsync(ex->d_flag, 0, sizeof(int), (cudaStream_t)stream));
    throw Error(HB_E_ARG, "evidence row inconsistent with the compiled observed set (or state out of range)");
  }
}

}  // namespace hb
