// jit.cpp -- per-step kernel specialisation at schedule-compile time.
//
// The templated kernels of engine.cu walk a step through index maps and a run-time odometer; for the steps that
// enumerate a large joint from small tables (the messages of a big clique) that costs several loads per joint entry
// and makes the L1TEX data pipe, not HBM, the bound (profiles/README.md).  Here the step is turned into CUDA source
// with every stride, dimension and table offset as an immediate: one thread owns two adjacent rows and a strip of
// groups of output entries (a group = all states of the step's blocking variable), the loop nest over the summed
// variables is fully unrolled for a group and every load is emitted ONCE per distinct address, so an input that does
// not depend on an inner summed variable or on the blocking variable is read once and kept in a register.
// The floating-point expression of every entry is the one the templated kernels (and the reference) evaluate:
//   value_l = 0 + sum_{s_l} value_{l+1} * (in_1 * (in_2 * ...))        (__dmul_rn / __dadd_rn, no contraction)
// The source is compiled with NVRTC for sm_100a and loaded through the driver API; both libraries are dlopen'ed on
// first use, so the library itself only depends on the CUDA runtime.  Any failure leaves the templated kernels in
// charge (HB_JIT=0 forces that; HB_JIT_VERBOSE=1 reports what happened on stderr).
#include "jit.hpp"

#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <initializer_list>
#include <map>
#include <cmath>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

namespace hb {

namespace {

bool verbose() {
  static const bool v = getenv("HB_JIT_VERBOSE") && atoi(getenv("HB_JIT_VERBOSE")) != 0;
  return v;
}

// ---- dynamic bindings -------------------------------------------------------------------------------
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  bool ok = false;
};
struct Driver {
  int (*ModuleLoadData)(void**, const void*) = nullptr;
  int (*ModuleGetFunction)(void**, void*, const char*) = nullptr;
  int (*ModuleUnload)(void*) = nullptr;
  int (*LaunchKernel)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**, void**) = nullptr;
  int (*FuncSetAttribute)(void*, int, int) = nullptr;
  int (*Occupancy)(int*, void*, int, size_t) = nullptr;
  bool ok = false;
};

void* open_first(std::initializer_list<const char*> names) {
  for (const char* n : names)
    if (void* h = dlopen(n, RTLD_NOW | RTLD_GLOBAL)) return h;
  return nullptr;
}

const Nvrtc& nvrtc() {
  static Nvrtc n = [] {
    Nvrtc r;
    void* h = open_first({"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"});
    if (!h) return r;
#define HB_SYM(field, name) *(void**)(&r.field) = dlsym(h, name)
    HB_SYM(CreateProgram, "nvrtcCreateProgram");
    HB_SYM(CompileProgram, "nvrtcCompileProgram");
    HB_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
    HB_SYM(GetCUBIN, "nvrtcGetCUBIN");
    HB_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
    HB_SYM(GetProgramLog, "nvrtcGetProgramLog");
    HB_SYM(DestroyProgram, "nvrtcDestroyProgram");
#undef HB_SYM
    r.ok = r.CreateProgram && r.CompileProgram && r.GetCUBINSize && r.GetCUBIN && r.GetProgramLogSize && r.GetProgramLog && r.DestroyProgram;
    return r;
  }();
  return n;
}

const Driver& driver() {
  static Driver d = [] {
    Driver r;
    void* h = open_first({"libcuda.so.1", "libcuda.so"});
    if (!h) return r;
    *(void**)(&r.ModuleLoadData) = dlsym(h, "cuModuleLoadData");
    *(void**)(&r.ModuleGetFunction) = dlsym(h, "cuModuleGetFunction");
    *(void**)(&r.ModuleUnload) = dlsym(h, "cuModuleUnload");
    *(void**)(&r.LaunchKernel) = dlsym(h, "cuLaunchKernel");
    *(void**)(&r.FuncSetAttribute) = dlsym(h, "cuFuncSetAttribute");
    *(void**)(&r.Occupancy) = dlsym(h, "cuOccupancyMaxActiveBlocksPerMultiprocessor");
    r.ok = r.ModuleLoadData && r.ModuleGetFunction && r.ModuleUnload && r.LaunchKernel;
    return r;
  }();
  return d;
}

// ---- code generation --------------------------------------------------------------------------------
struct In {
  int id, level;
  const Table* t;
  const StepInput* in;
};

int64_t stride_in(const Table& t, int v) {
  for (size_t i = 0; i < t.pscope.size(); ++i)
    if (t.pscope[i] == v) return t.pstride[i];
  return 0;
}

int loop_levels_for(const Step& st, int64_t U, int64_t max_ops);  // defined below

struct Gen {
  const Program& P;
  const Step& st;
  std::ostringstream o;
  std::vector<In> ins;
  std::map<std::pair<int, int64_t>, std::string> loaded;
  int tmp = 0;
  int U = 1;
  // The n_loop outermost levels stay run-time loops (a step whose fully unrolled body would be too large); the levels
  // below them are unrolled.  ptr[i] = name of input i's pointer in the current scope (it absorbs the loop variables).
  int n_loop = 0;
  int scope = 0;
  std::vector<std::string> ptr;
  // V = evidence rows per thread: 2 (adjacent rows, one 16-byte access per row table) or 1.  no_block: one output entry
  // per item instead of a group of all states of the blocking variable (finer items for the narrow steps of hb_tree).
  int V = 2;
  bool no_block = false;
  // Register block: the output variables a thread enumerates itself (U = product of their dimensions entries evaluated in
  // lock-step, every load emitted once per distinct address); the others index the item.  Default: the step's blocking
  // variable w alone.  `custom_block` = a caller-chosen set (hb_tree picks it per step by the loads it saves).
  std::vector<int> blocked;
  bool custom_block = false;
  int64_t body_budget = 600;           // chain evaluations unrolled per item; outer levels beyond it become run-time loops
  // hb_step_<i>: groups dealt round robin to the lanes (adjacent lanes, adjacent groups) instead of one contiguous strip per
  // lane; `unroll` = groups of the strip loop the compiler may evaluate together (more loads in flight per thread)
  bool interleave = false;
  int unroll = 0;
  // The reference starts every sum at 0 (`sum` = foldl (+) 0): 0 + x.  When no potential of the network is negative, -0.0 or
  // NaN, no intermediate value can be -0.0 or a signalling NaN, so 0 + x == x bit for bit and the add is not emitted
  // (jit_zero_add_elidable; a changeFactor that breaks the premise drops the kernels, engine.cu).
  bool elide_zero_add = false;
  // Re-reading steps: the digits of the output variables that the largest per-row input does not mention vary fastest along
  // a lane's strip of groups, so consecutive groups re-read that input at the same addresses (L1 hits instead of L2 traffic).
  bool reorder_digits = false;
  // ... and the strip of a lane becomes two loops: `go` over the digits the largest input depends on, `gi` over the others,
  // so that input's addresses (and, registers permitting, its values) are invariant in the inner loop.
  bool split_loop = false;
  size_t n_slow = 0;     // leading digits of `red` that index the largest per-row input (split_loop)
  int64_t n_inner = 1;   // groups of the inner loop (split_loop), 1 otherwise
  int min_blocks = 2;    // hb_step_<i>: CTAs per SM the kernel is bounded for (a large register block needs the registers of one)
  int64_t n_groups = 1;  // n_out / U, set by prepare()
  std::vector<std::vector<int>> udig;  // per lock-step entry: state of every blocked variable
  std::vector<int64_t> uoff;           // per lock-step entry: offset in the stored output layout
  Gen(const Program& p, const Step& s) : P(p), st(s) {}
  const char* DV() const { return V == 2 ? "double2" : "double"; }
  const char* zero() const { return V == 2 ? "zero" : "0.0"; }

  std::string load(const In& x, int64_t off) {
    auto key = std::make_pair(x.id, off);
    auto it = loaded.find(key);
    if (it != loaded.end()) return it->second;
    std::string name = "x" + std::to_string(x.id) + "_" + std::to_string(off) + (scope ? "_" + std::to_string(scope) : "");
    const Table& t = *x.t;
    const std::string& q = ptr[x.id];
    const char* ld = tree_off ? "*(" : "__ldg(";  // in hb_tree the potentials are staged in shared memory
    if (V == 1) {
      if (t.kind == T_ROW)
        o << "    const double " << name << " = " << q << "[(size_t)" << off << " * RT];\n";
      else if (t.kind == T_POT && t.slice_slot >= 0)
        o << "    const double " << name << " = " << q << "a[" << off << "];\n";
      else
        o << "    const double " << name << " = " << ld << q << " + " << off << ");\n";
    } else if (t.kind == T_ROW)
      o << "    const double2 " << name << " = *reinterpret_cast<const double2*>(" << q << " + (size_t)" << off << " * RT);\n";
    else if (t.kind == T_POT && t.slice_slot >= 0)
      o << "    const double2 " << name << " = make_double2(" << q << "a[" << off << "], " << q << "b[" << off << "]);\n";
    else
      o << "    const double " << name << "s = " << ld << q << " + " << off << "); const double2 " << name << " = make_double2(" << name << "s, "
        << name << "s);\n";
    ++n_loads;
    for (size_t j = 0; j < x.in->scales.size(); ++j) {  // x = ((t[i] * s0) * s1) * ...: the scale step evaluated where it is read
      const std::string scaled = name + "_s" + std::to_string(j);
      o << "    const " << DV() << " " << scaled << " = hb_mul(" << name << ", scl" << j << ");\n";
      name = scaled;
    }
    loaded[key] = name;
    return name;
  }
  int64_t n_loads = 0;  // load instructions emitted so far (one per distinct address of a group body)
  std::string mul(const std::string& a, const std::string& b) {
    std::string n = "m" + std::to_string(++tmp);
    o << "    const " << DV() << " " << n << " = hb_mul(" << a << ", " << b << ");\n";
    return n;
  }
  std::string add(const std::string& a, const std::string& b) {
    std::string n = "s" + std::to_string(++tmp);
    o << "    const " << DV() << " " << n << " = hb_add(" << a << ", " << b << ");\n";
    return n;
  }
  // chain of the inputs of level li for entry u (offsets of the unrolled levels from `bound`), times the inner value
  std::string term(size_t li, int u, const std::vector<int>& bound, const std::vector<std::string>& inner) {
    std::string chain;
    for (size_t k = ins.size(); k-- > 0;) {
      const In& x = ins[k];
      if (x.level != (int)li) continue;
      int64_t off = 0;
      for (size_t j = 0; j < blocked.size(); ++j) off += (int64_t)udig[u][j] * stride_in(*x.t, blocked[j]);
      for (size_t m = (size_t)n_loop; m <= li; ++m) off += (int64_t)bound[m] * x.in->lstride[m];  // looped levels live in the pointer
      const std::string v = load(x, off);
      chain = chain.empty() ? v : mul(v, chain);
    }
    if (!inner.empty()) chain = chain.empty() ? inner[u] : mul(inner[u], chain);
    if (li == 0 && !scalar_product.empty()) chain = chain.empty() ? scalar_product : mul(scalar_product, chain);  // ps * chain
    return chain;
  }
  // value_l of the U lock-step entries, with the summed states of the outer levels bound
  std::vector<std::string> level(size_t li, std::vector<int>& bound) {
    const Level& lv = st.levels[li];
    std::vector<std::string> acc(U);
    if ((int)li < n_loop) {  // run-time loop over the states of this level's summed variable
      for (int u = 0; u < U; ++u) {
        acc[u] = "a" + std::to_string(li) + "_" + std::to_string(u);
        o << "    " << DV() << " " << acc[u] << " = " << zero() << ";\n";
      }
      o << "    for (unsigned k" << li << " = 0; k" << li << " < " << lv.d << "u; ++k" << li << ") {\n";
      const auto saved_loaded = loaded;
      const auto saved_ptr = ptr;
      const int saved_scope = scope;
      scope = (int)li + 1;
      for (const In& x : ins) {  // inputs of this level and of the levels below move with the loop variable
        if (x.level < (int)li || x.in->lstride[li] == 0) continue;
        const Table& t = *x.t;
        const std::string q = "q" + std::to_string(x.id) + "_" + std::to_string(li);
        const std::string step = "k" + std::to_string(li) + " * " + std::to_string(x.in->lstride[li]) + "u";
        if (t.kind == T_ROW)
          o << "    const double* const " << q << " = " << ptr[x.id] << " + (size_t)(" << step << ") * RT;\n";
        else if (t.kind == T_POT && t.slice_slot >= 0) {
          o << "    const double* const " << q << "a = " << ptr[x.id] << "a + (" << step << ");\n";
          if (V == 2) o << "    const double* const " << q << "b = " << ptr[x.id] << "b + (" << step << ");\n";
        } else
          o << "    const double* const " << q << " = " << ptr[x.id] << " + (" << step << ");\n";
        ptr[x.id] = q;
      }
      std::vector<std::string> inner;
      if (li + 1 < st.levels.size()) inner = level(li + 1, bound);
      for (int u = 0; u < U; ++u) {
        const std::string c = term(li, u, bound, inner);
        o << "    " << acc[u] << " = hb_add(" << acc[u] << ", " << c << ");\n";
      }
      o << "    }\n";
      loaded = saved_loaded;
      ptr = saved_ptr;
      scope = saved_scope;
      return acc;
    }
    for (int s = 0; s < lv.d; ++s) {
      bound[li] = s;
      std::vector<std::string> inner;
      if (li + 1 < st.levels.size()) inner = level(li + 1, bound);
      for (int u = 0; u < U; ++u) {
        const std::string chain = term(li, u, bound, inner);
        acc[u] = acc[u].empty() ? (elide_zero_add ? chain : add(zero(), chain)) : add(acc[u], chain);
      }
    }
    return acc;
  }

  // ---- pieces shared by the per-step kernels (hb_step_<i>) and the whole-schedule on-chip kernel (hb_tree) ----
  // In the on-chip kernel the row arena is shared memory: `tree_off` gives every per-row table its place there and the
  // row stride RT is the compile-time tile height, so every address below folds into an immediate.
  const std::vector<int64_t>* tree_off = nullptr;
  std::vector<int> red;  // output variables other than the blocking one, stored row-major (last fastest), then w
  std::vector<int64_t> rdim;  // states of each digit of `red` (a digit may stand for a run of merged variables)
  bool merge_digits = !(getenv("HB_JIT_MERGE") && atoi(getenv("HB_JIT_MERGE")) == 0);
  int64_t row_off(const Table& t) const { return tree_off ? (*tree_off)[(size_t)(&t - &P.tables[0])] : t.off; }

  void prepare() {
    const Table& out = P.tables[st.out];
    U = st.U;
    int id = 0;
    ins.clear();
    for (size_t l = 0; l < st.levels.size(); ++l)
      for (const StepInput& in : st.levels[l].in) ins.push_back(In{id++, (int)l, &P.tables[in.table], &in});
    ptr.clear();
    for (const In& x : ins) ptr.push_back("p" + std::to_string(x.id));
    if (!custom_block) {
      blocked.clear();
      if (!no_block && st.w_var >= 0) blocked.push_back(st.w_var);
    }
    red.clear();
    rdim.clear();
    auto pstride_of = [](const Table& t, int v) {
      for (size_t i = 0; i < t.pscope.size(); ++i)
        if (t.pscope[i] == v) return t.pstride[i];
      return (int64_t)0;
    };
    for (int v : out.pscope) {
      if (std::find(blocked.begin(), blocked.end(), v) != blocked.end()) continue;
      // Adjacent digits a (slower), v (faster) that every table of the step walks as ONE mixed-radix digit -- stride(a) =
      // stride(v) * dim(v) in the output and in every input (0 = 0 counts) -- are merged: one digit of dim(a) * dim(v) states
      // with v's strides.  The 2^23 groups of a step over 23 binary variables then cost no index arithmetic at all.
      const int64_t dv = P.net->dims[v];
      bool merge = merge_digits && !red.empty() && rdim.back() * dv < ((int64_t)1 << 31);
      if (merge) {
        const int a = red.back();
        merge = pstride_of(out, a) == pstride_of(out, v) * dv;
        for (const In& x : ins)
          if (pstride_of(*x.t, a) != pstride_of(*x.t, v) * dv) merge = false;
      }
      if (merge) {
        red.back() = v;  // the run is named after its fastest variable
        rdim.back() *= dv;
      } else {
        red.push_back(v);
        rdim.push_back(dv);
      }
    }
    if (reorder_digits && red.size() > 1) {
      const Table* big = nullptr;
      for (const In& x : ins)
        if (x.t->kind == T_ROW && (!big || x.t->size > big->size)) big = x.t;
      if (big) {
        std::vector<int> r2;
        std::vector<int64_t> d2;
        for (int pass = 0; pass < 2; ++pass)
          for (size_t j = 0; j < red.size(); ++j)
            if ((pstride_of(*big, red[j]) != 0) == (pass == 0)) r2.push_back(red[j]), d2.push_back(rdim[j]);
        red = r2;
        rdim = d2;
        n_slow = 0;
        for (int v : red) n_slow += pstride_of(*big, v) != 0;
        n_inner = 1;
        if (split_loop && n_slow > 0 && n_slow < red.size())
          for (size_t j = n_slow; j < red.size(); ++j) n_inner *= rdim[j];
      }
    }
    U = 1;
    for (int v : blocked) U *= P.net->dims[v];
    n_groups = st.n_out / U;
    auto out_stride = [&](int v) {
      for (size_t i = 0; i < out.pscope.size(); ++i)
        if (out.pscope[i] == v) return out.pstride[i];
      return (int64_t)0;
    };
    udig.assign(U, std::vector<int>(blocked.size(), 0));
    uoff.assign(U, 0);
    for (int u = 0; u < U; ++u) {
      int rem = u;
      for (size_t j = blocked.size(); j-- > 0;) udig[u][j] = rem % P.net->dims[blocked[j]], rem /= P.net->dims[blocked[j]];
      for (size_t j = 0; j < blocked.size(); ++j) uoff[u] += udig[u][j] * out_stride(blocked[j]);
    }
    n_loop = std::max(0, loop_levels_for(st, U, body_budget));
  }
  // offset of the item's first entry in the stored output layout, as an expression of the digits d<v> of `red`
  std::string out_base() const {
    const Table& out = P.tables[st.out];
    std::string base;
    for (int v : red)
      for (size_t i = 0; i < out.pscope.size(); ++i)
        if (out.pscope[i] == v && out.pstride[i])
          base += (base.empty() ? "" : " + ") + ("d" + std::to_string(v)) + " * " + std::to_string(out.pstride[i]) + "u";
    return base.empty() ? "0u" : base;
  }

  // row pointers that do not depend on the group (needs `row`, `arena`, `arena_out`, `pots`, `shared`, `rowoff`, `RT` in scope)
  void emit_bases(const char* ind, bool with_outp = true) {
    const Table& out = P.tables[st.out];
    for (const In& x : ins) {
      const Table& t = *x.t;
      if (t.kind == T_ROW)
        o << ind << "const double* const b" << x.id << " = arena + (size_t)" << row_off(t) << " * RT + row;\n";
      else if (t.kind == T_POT && t.slice_slot >= 0) {
        o << ind << "const double* const b" << x.id << "a = pots + " << t.off << " + rowoff[(size_t)" << t.slice_slot << " * RT + row];\n";
        if (V == 2) o << ind << "const double* const b" << x.id << "b = pots + " << t.off << " + rowoff[(size_t)" << t.slice_slot << " * RT + row + 1];\n";
      } else
        o << ind << "const double* const b" << x.id << " = " << (t.kind == T_POT ? "pots" : "shared") << " + " << t.off << ";\n";
    }
    // the one-entry tables the scaled input's entries are multiplied by
    for (const In& x : ins)
      for (size_t j = 0; j < x.in->scales.size(); ++j) {
        const Table& t = P.tables[x.in->scales[j]];
        const char* pool = t.kind == T_POT ? "pots" : "shared";
        const std::string n = "scl" + std::to_string(j);
        if (t.kind == T_POT && t.slice_slot >= 0) {
          if (V == 2)
            o << ind << "const double2 " << n << " = make_double2(pots[" << t.off << " + rowoff[(size_t)" << t.slice_slot << " * RT + row]], pots[" << t.off
              << " + rowoff[(size_t)" << t.slice_slot << " * RT + row + 1]]);\n";
          else
            o << ind << "const double " << n << " = pots[" << t.off << " + rowoff[(size_t)" << t.slice_slot << " * RT + row]];\n";
        } else if (V == 2)
          o << ind << "const double2 " << n << " = make_double2(" << pool << "[" << t.off << "], " << pool << "[" << t.off << "]);\n";
        else
          o << ind << "const double " << n << " = " << pool << "[" << t.off << "];\n";
      }
    // structural scalars of level 0: ps = s1 * (s2 * ...), applied as ps * chain (PrivateCPT.hs:225-227)
    std::string ps;
    for (size_t k = st.scalars.size(); k-- > 0;) {
      const Table& t = P.tables[st.scalars[k]];
      const std::string n = "sc" + std::to_string(k);
      const char* pool = t.kind == T_POT ? "pots" : "shared";
      if (V == 1) {
        if (t.kind == T_ROW)
          o << ind << "const double " << n << " = arena[(size_t)" << row_off(t) << " * RT + row];\n";
        else if (t.kind == T_POT && t.slice_slot >= 0)
          o << ind << "const double " << n << " = pots[" << t.off << " + rowoff[(size_t)" << t.slice_slot << " * RT + row]];\n";
        else
          o << ind << "const double " << n << " = " << pool << "[" << t.off << "];\n";
      } else if (t.kind == T_ROW)
        o << ind << "const double2 " << n << " = *reinterpret_cast<const double2*>(arena + (size_t)" << row_off(t) << " * RT + row);\n";
      else if (t.kind == T_POT && t.slice_slot >= 0)
        o << ind << "const double2 " << n << " = make_double2(pots[" << t.off << " + rowoff[(size_t)" << t.slice_slot << " * RT + row]], pots[" << t.off
          << " + rowoff[(size_t)" << t.slice_slot << " * RT + row + 1]]);\n";
      else
        o << ind << "const double2 " << n << " = make_double2(" << pool << "[" << t.off << "], " << pool << "[" << t.off << "]);\n";
      if (ps.empty())
        ps = n;
      else {
        const std::string m = "scp" + std::to_string(k);
        o << ind << "const " << DV() << " " << m << " = hb_mul(" << n << ", " << ps << ");\n";
        ps = m;
      }
    }
    scalar_product = ps;
    if (with_outp) o << ind << "double* const outp = arena_out + (size_t)" << row_off(out) << " * RT + row;\n";  // same arena: the output region is never read here
  }
  std::string scalar_product;  // name of ps, empty when the step has no structural scalars

  // the body of one group `g` of output entries: digits, per-input pointers, the unrolled nest, the stores
  enum { STORE_STREAMING = 0, STORE_PLAIN = 1, STORE_RESULT = 2 };  // __stcs to HBM / plain store / res[u] = value
  void emit_group(int store) {
    loaded.clear();
    if (n_inner > 1) {  // the digits of `go` (slow) and of `gi` (fast) separately
      o << "    unsigned rem = go;\n";
      for (size_t j = n_slow; j-- > 0;)
        o << "    const unsigned d" << red[j] << " = " << (j ? "rem % " + std::to_string(rdim[j]) + "u;" : "rem;") << (j ? " rem /= " + std::to_string(rdim[j]) + "u;" : "")
          << "\n";
      o << "    unsigned remi = gi;\n";
      for (size_t j = red.size(); j-- > n_slow;)
        o << "    const unsigned d" << red[j] << " = " << (j > n_slow ? "remi % " + std::to_string(rdim[j]) + "u;" : "remi;")
          << (j > n_slow ? " remi /= " + std::to_string(rdim[j]) + "u;" : "") << "\n";
    } else {
      if (!red.empty()) o << "    unsigned rem = g;\n";
      for (size_t j = red.size(); j-- > 0;)
        o << "    const unsigned d" << red[j] << " = " << (j ? "rem % " + std::to_string(rdim[j]) + "u;" : "rem;") << (j ? " rem /= " + std::to_string(rdim[j]) + "u;" : "")
          << "\n";
    }
    for (const In& x : ins) {
      std::string base;
      for (int v : red) {
        const int64_t s = stride_in(*x.t, v);
        if (s) base += (base.empty() ? "" : " + ") + ("d" + std::to_string(v)) + " * " + std::to_string(s) + "u";
      }
      if (base.empty()) base = "0u";
      const Table& t = *x.t;
      if (t.kind == T_ROW)
        o << "    const double* const p" << x.id << " = b" << x.id << " + (size_t)(" << base << ") * RT;\n";
      else if (t.kind == T_POT && t.slice_slot >= 0) {
        o << "    const double* const p" << x.id << "a = b" << x.id << "a + (" << base << ");\n";
        if (V == 2) o << "    const double* const p" << x.id << "b = b" << x.id << "b + (" << base << ");\n";
      } else
        o << "    const double* const p" << x.id << " = b" << x.id << " + (" << base << ");\n";
    }
    o << "    const unsigned ob = " << out_base() << ";\n";
    std::vector<int> bound(st.levels.size(), 0);
    const std::vector<std::string> res = level(0, bound);
    if (store == STORE_RESULT) o << "    *ob_out = ob;\n";
    for (int u = 0; u < U; ++u) {
      if (store == STORE_RESULT)
        o << "    res[" << u << "] = " << res[u] << ";\n";
      else if (store == STORE_STREAMING && V == 2)
        o << "    __stcs(reinterpret_cast<double2*>(outp + (size_t)(ob + " << uoff[u] << "u) * RT), " << res[u] << ");\n";
      else if (store == STORE_STREAMING)
        o << "    __stcs(outp + (size_t)(ob + " << uoff[u] << "u) * RT, " << res[u] << ");\n";
      else if (V == 2)
        o << "    *reinterpret_cast<double2*>(outp + (size_t)(ob + " << uoff[u] << "u) * RT) = " << res[u] << ";\n";
      else
        o << "    outp[(size_t)(ob + " << uoff[u] << "u) * RT] = " << res[u] << ";\n";
    }
  }

  std::string kernel(const std::string& name) {
    prepare();
    // two CTAs per SM (<= 128 registers): measured on C2, 1 -> 86.2, 2 -> 89.7, 3 -> 75.9 M queries/s
    o << "extern \"C\" __global__ void __launch_bounds__(256, " << (getenv("HB_JIT_MINBLOCKS") ? atoi(getenv("HB_JIT_MINBLOCKS")) : min_blocks) << ") " << name
      << "(const double* __restrict__ arena, double* __restrict__ arena_out, const double* __restrict__ pots,\n"
         "    const double* __restrict__ shared, const int* __restrict__ rowoff, int nr, long long RT) {\n";
    // strips of groups along blockIdx.x (the fast axis of the block scheduler), row blocks along blockIdx.y: the CTAs that are
    // resident together work on the SAME rows, so a step that enumerates its inputs many times finds them in L2
    o << "  const int row = (blockIdx.y * blockDim.x + threadIdx.x)" << (V == 2 ? " * 2" : "") << ";\n  if (row >= nr) return;\n";
    o << "  const unsigned lanes = gridDim.x * blockDim.y, lane = blockIdx.x * blockDim.y + threadIdx.y;\n";
    const int64_t n_outer = n_groups / n_inner;
    if (!interleave) {
      o << "  const unsigned per = (" << n_outer << "u + lanes - 1) / lanes;\n";
      o << "  const unsigned g_end = min(" << n_outer << "u, lane * per + per);\n";
    }
    if (V == 2) o << "  const double2 zero = make_double2(0.0, 0.0);\n";
    emit_bases("  ");
    if (interleave && unroll > 1) {
      // `unroll` groups at a time, all their loads ahead of the first store (the compiler does not move a load of the next
      // group above the stores of this one): a light streaming step keeps several times the bytes in flight per thread
      o << "  auto body = [&](const unsigned g, double2* __restrict__ res, unsigned* __restrict__ ob_out) {\n";
      emit_group(STORE_RESULT);
      o << "  };\n";
      o << "  for (unsigned g = lane; g < " << n_groups << "u; g += lanes * " << unroll << "u) {\n";
      for (int k = 0; k < unroll; ++k) {
        if (k) o << "    const bool has" << k << " = g + lanes * " << k << "u < " << n_groups << "u;\n";
        o << "    double2 r" << k << "[" << U << "]; unsigned ob" << k << ";\n";
        o << "    body(" << (k ? "has" + std::to_string(k) + " ? g + lanes * " + std::to_string(k) + "u : g" : std::string("g")) << ", r" << k << ", &ob" << k << ");\n";
      }
      for (int k = 0; k < unroll; ++k) {
        if (k) o << "    if (has" << k << ") {\n";
        for (int u = 0; u < U; ++u)
          o << "    __stcs(reinterpret_cast<double2*>(outp + (size_t)(ob" << k << " + " << uoff[u] << "u) * RT), r" << k << "[" << u << "]);\n";
        if (k) o << "    }\n";
      }
      o << "  }\n}\n\n";
      return o.str();
    }
    if (interleave)  // the lanes that run together take adjacent groups: a streaming step touches HBM like a copy does
      o << "  for (unsigned g = lane; g < " << n_groups << "u; g += lanes) {\n";
    else if (n_inner > 1)
      o << "  for (unsigned go = lane * per; go < g_end; ++go) {\n#pragma unroll 1\n  for (unsigned gi = 0; gi < " << n_inner << "u; ++gi) {\n";
    else
      o << "  for (unsigned g = lane * per; g < g_end; ++g) {\n";
    emit_group(STORE_STREAMING);
    o << (n_inner > 1 && !interleave ? "  }\n" : "") << "  }\n}\n\n";
    return o.str();
  }

  // The same step inside the on-chip kernel: its (group, row pair) items are dealt to the threads of the CTA, thread
  // `rot` first, so that the small steps of one dependency level land on different warps.
  // items the step is cut into for a tile of R rows and T threads: groups of all states of the blocking variable x row
  // pairs when that already feeds every thread, else single output entries, else single entries x single rows
  // The step as a device function of one item (inlined at its call sites) and the loop that deals the items to the
  // threads.  A thread with several items evaluates `ni` of them together: the bodies are independent and no store
  // separates them, so their shared-memory loads and fp64 chains overlap.
  std::string call;  // filled by tree_step: the loop inside hb_tree
  int64_t tree_items = 0;  // filled by tree_step
  // `block` = the register block chosen for this step, v = rows per thread
  void tree_step(int si, int R, int T, int rot, const std::vector<int>& block, int v) {
    V = v;
    custom_block = true;
    blocked = block;
    merge_digits = false;  // the blocks of a clique-sized step have two or three digits: nothing to gain (C2: same instruction count)
    prepare();
    const int64_t rows = V == 2 ? R / 2 : R;
    const int64_t items = st.n_out / U * rows;
    tree_items = items;
    const std::string fn = "hb_s" + std::to_string(si);
    o << "__device__ __forceinline__ void " << fn << "(const double* __restrict__ arena, const double* __restrict__ pots,\n"
         "    const double* __restrict__ shared, const int* __restrict__ rowoff, const unsigned g, const unsigned row, " << DV()
      << "* __restrict__ res, unsigned* __restrict__ ob_out) {\n";
    o << "    constexpr unsigned RT = " << R << "u;\n";
    if (V == 2) o << "    const double2 zero = make_double2(0.0, 0.0);\n";
    emit_bases("    ", false);
    emit_group(STORE_RESULT);
    o << "}\n\n";
    // A thread keeps its row (pair) for the whole kernel: rowA / slotA (two rows per thread) or rowB / slotB (one row) are
    // computed once at kernel start; the blocks of a step go to the slots round robin, slot `rot` first.
    const int64_t n_blocks = items / rows, n_slots = T / rows;
    const char* rw = V == 2 ? "rowA" : "rowB";
    const char* sl = V == 2 ? "slotA" : "slotB";
    const int64_t first = (n_slots - (rot / rows) % n_slots) % n_slots;
    // one item per thread at a time: evaluating two or four together (no store between their bodies) measured 6 % slower on
    // C2 -- larger bodies cost more in registers and instruction fetch than the overlap gains
    const int ni = 1;
    std::ostringstream c;
    c << "    {  // step " << si << ": " << n_blocks << " blocks of " << U << " x " << rows << (V == 2 ? " row pairs" : " rows")
      << (ni > 1 ? ", " + std::to_string(ni) + " items at a time\n" : "\n");
    c << "    double* const outb = arena_out + " << row_off(P.tables[st.out]) * R << "u + " << rw << ";\n";
    const std::string g_first = first ? "(" + std::string(sl) + " + " + std::to_string(first) + "u) % " + std::to_string(n_slots) + "u" : std::string(sl);
    if (n_blocks <= n_slots)
      c << "    { const unsigned g = " << g_first << ";\n    if (g < " << n_blocks << "u) {\n";
    else
      c << "    for (unsigned g = " << g_first << "; g < " << n_blocks << "u; g += " << ni * n_slots << "u) {{\n";
    for (int k = 0; k < ni; ++k) {
      if (k == 0)
        c << "      const unsigned g0 = g;\n";
      else
        c << "      const bool has" << k << " = g + " << k * n_slots << "u < " << n_blocks << "u;\n      const unsigned g" << k << " = has" << k << " ? g + " << k * n_slots
          << "u : g;\n";
      c << "      " << DV() << " r" << k << "[" << U << "];\n      unsigned ob" << k << ";\n";
      c << "      " << fn << "(arena, pots, shared, rowoff, g" << k << ", " << rw << ", r" << k << ", &ob" << k << ");\n";
    }
    for (int k = 0; k < ni; ++k) {
      if (k > 0) c << "      if (has" << k << ") {\n";
      for (int u = 0; u < U; ++u) {
        const std::string at = "outb + (ob" + std::to_string(k) + " + " + std::to_string(uoff[u]) + "u) * RT";
        if (V == 2)
          c << "      *reinterpret_cast<double2*>(" << at << ") = r" << k << "[" << u << "];\n";
        else
          c << "      *(" << at << ") = r" << k << "[" << u << "];\n";
      }
      if (k > 0) c << "      }\n";
    }
    c << "    }}\n    }\n";
    call = c.str();
  }
};

}  // namespace

// chain evaluations in the generated body of a group when the `n_loop` outermost levels stay run-time loops
// (about 150 bytes of source each)
static int64_t body_ops(const Step& st, int n_loop, int64_t U) {
  int64_t ops = 0, terms = 1;
  for (size_t l = 0; l < st.levels.size(); ++l) {
    if ((int)l >= n_loop) terms *= st.levels[l].d;
    ops += terms * (int64_t)std::max<size_t>(1, st.levels[l].in.size());
  }
  return ops * U;
}
// compile time grows faster than linearly with the size of a straight-line kernel; 100 .. 600 measured the same on C2
constexpr int64_t MAX_BODY_OPS = 600;

// how many outermost levels stay run-time loops so that the unrolled rest fits MAX_BODY_OPS (-1: not even the
// innermost level alone fits)
namespace {
int loop_levels_for(const Step& st, int64_t U, int64_t max_ops) {
  for (int n = 0; n < (int)st.levels.size(); ++n)
    if (body_ops(st, n, U) <= max_ops) return n;
  // even the innermost level alone exceeds the budget: acceptable as long as it stays below the hard limit
  return body_ops(st, (int)st.levels.size() - 1, U) <= MAX_BODY_OPS ? (int)st.levels.size() - 1 : -1;
}
}  // namespace
int jit_loop_levels(const Step& st) { return loop_levels_for(st, st.U, MAX_BODY_OPS); }
int64_t jit_unrolled_ops(const Step& st) {
  const int n = jit_loop_levels(st);
  return n < 0 ? -1 : body_ops(st, n, st.U);
}

static bool jit_basic(const Program& P, const Step& st) {
  if (st.kind != STEP_PRODSUM || st.shared || !st.scalars.empty() || st.levels.empty() || st.levels[0].is_max) return false;
  if (P.tables[st.out].kind != T_ROW || st.U > 8 || st.n_inputs() < 1) return false;
  if (st.n_red >= ((int64_t)1 << 31)) return false;
  for (const Level& l : st.levels)
    for (const StepInput& in : l.in)
      if (P.tables[in.table].kind == T_ONE) return false;
  return true;
}

bool jit_eligible(const Program& P, const Step& st) {
  if (!jit_basic(P, st)) return false;
  // A step whose body only fits with run-time loops around it stays on the precompiled kernels: as a kernel of its own it
  // measured no faster (C3 with every such step generated: 6.6 vs 6.2 ms; with only one looped level allowed: 4.91 vs 4.72 ms).
  // (Inside hb_tree such steps do use the loops.)  The exception: a re-reading step with a register block over several
  // output variables (jit_step_shape) -- there the loops buy a body with a fraction of the loads per entry.
  if (jit_loop_levels(st) == 0) return true;
  return jit_step_shape(P, st).eligible;
}

// loads per row (8-byte slots, one row) the specialised kernel issues for the step / the templated kernels issue
void jit_load_counts(const Program& P, const Step& st, int64_t* specialised, int64_t* templated) {
  Gen g(P, st);
  g.kernel("probe");
  *specialised = (int64_t)g.loaded.size() * st.n_red;
  const int64_t nch = (st.U + 1) / 2;  // the templated kernels walk the blocking variable two states at a time
  int64_t terms = 1, t = 0;
  for (const Level& l : st.levels) {
    terms *= l.d;
    for (const StepInput& in : l.in) t += st.n_red * nch * terms * (in.wstride != 0 ? 2 : 1);
  }
  *templated = t;
}

struct JitModule {
  void* module = nullptr;
  std::string key;  // cache key of the cubin it was loaded from
  ~JitModule() {
    if (module && driver().ok) driver().ModuleUnload(module);
  }
};

// arithmetic of the generated kernels: explicit round-to-nearest multiply / add, never contracted (the 1e-12 mode
// compiles the same source with --fmad=true and HB_FMA defined: plain operators, contraction allowed)
static const char* const kPreamble =
    "#ifdef HB_FMA\n"
    "__device__ __forceinline__ double hb_mul(double a, double b) { return a * b; }\n"
    "__device__ __forceinline__ double hb_add(double a, double b) { return a + b; }\n"
    "#else\n"
    "__device__ __forceinline__ double hb_mul(double a, double b) { return __dmul_rn(a, b); }\n"
    "__device__ __forceinline__ double hb_add(double a, double b) { return __dadd_rn(a, b); }\n"
    "#endif\n"
    "__device__ __forceinline__ double2 hb_mul(double2 a, double2 b) { return make_double2(hb_mul(a.x, b.x), hb_mul(a.y, b.y)); }\n"
    "__device__ __forceinline__ double2 hb_add(double2 a, double2 b) { return make_double2(hb_add(a.x, b.x), hb_add(a.y, b.y)); }\n\n";

// A step that enumerates its per-row inputs several times over (the message of a big clique built from small tables) keeps
// one contiguous strip of groups per lane: the re-reads of a lane stay in L1 / L2 (launch_prodsum, engine.cu).  A streaming
// step -- every per-row entry read about once -- deals its groups round robin.
namespace {
// loads one group executes when the output variables `block` are evaluated in lock step and the n_loop outermost levels are
// run-time loops: per input one load per distinct address of the unrolled part (states of the block variables it mentions
// x states of the summed variables it mentions), repeated by every looped level around it
int64_t block_loads(const Program& P, const Step& st, const std::vector<int>& block, int n_loop) {
  int64_t total = 0;
  for (size_t l = 0; l < st.levels.size(); ++l)
    for (const StepInput& in : st.levels[l].in) {
      const Table& t = P.tables[in.table];
      int64_t n = 1;
      for (int v : block)
        if (stride_in(t, v) != 0) n *= P.net->dims[v];
      for (size_t m = 0; m <= l; ++m)
        if ((int)m < n_loop || in.lstride[m] != 0) n *= st.levels[m].d;
      total += n;
    }
  return total;
}
}  // namespace

JitShape jit_step_shape(const Program& P, const Step& st) {
  static const int mode = getenv("HB_JIT_INTERLEAVE") ? atoi(getenv("HB_JIT_INTERLEAVE")) : 1;  // 0 never, 1 streaming steps, 2 always
  // measured (profiles/r2_step_probe.jsonl): C4, 4 rows: 2.64 ms per pass with one group per iteration, 2.34 ms with four; C5: 23.5 / 20.8 ms
  static const int unroll = getenv("HB_JIT_UNROLL") ? atoi(getenv("HB_JIT_UNROLL")) : 4;
  // register blocks over several output variables for re-reading steps: at most this many entries in lock step (0 = off)
  // measured on C3 (profiles/r2_step_block.jsonl, ms per pass): off 4.41 | 8 entries, two CTAs per SM, 320-evaluation bodies 3.85 |
  // 16 entries, one CTA per SM (255 registers), 600-evaluation bodies 3.77 | 8, one, 600: 3.78 | 16, two, 320: 4.10
  static const int block_u = getenv("HB_JIT_BLOCK_U") ? atoi(getenv("HB_JIT_BLOCK_U")) : 16;
  static const int block_loops = getenv("HB_JIT_BLOCK_LOOPS") ? atoi(getenv("HB_JIT_BLOCK_LOOPS")) : 2;
  static const int block_minb = getenv("HB_JIT_BLOCK_MINB") ? atoi(getenv("HB_JIT_BLOCK_MINB")) : 1;
  static const int64_t block_body = getenv("HB_JIT_BLOCK_BODY") ? atoll(getenv("HB_JIT_BLOCK_BODY")) : MAX_BODY_OPS;  // unrolled chain evaluations per group
  // one row per thread: half the registers of a row pair (measured on C3: 4.07 ms per pass against 4.31 with row pairs)
  static const int block_v = getenv("HB_JIT_BLOCK_V") && atoi(getenv("HB_JIT_BLOCK_V")) == 2 ? 2 : 1;
  static const int64_t block_min_joint = getenv("HB_JIT_BLOCK_MIN_JOINT") ? atoll(getenv("HB_JIT_BLOCK_MIN_JOINT")) : 8192;
  int64_t terms = 1, row_entries = 0, row_loads = 0;
  for (const Level& l : st.levels) {
    terms *= l.d;
    for (const StepInput& in : l.in)
      if (P.tables[in.table].kind == T_ROW) row_entries += P.tables[in.table].size, row_loads += st.n_out * terms;
  }
  JitShape sh;
  sh.rereads = row_entries > 0 && row_loads >= 3 * row_entries;
  sh.interleave = mode == 2 || (mode == 1 && !sh.rereads);
  if (st.w_var >= 0) sh.block.push_back(st.w_var);
  sh.U = st.U;
  sh.groups = st.n_red;
  sh.n_loop = jit_loop_levels(st);
  sh.ops = jit_unrolled_ops(st);
  sh.eligible = jit_basic(P, st) && sh.n_loop == 0;
  if (sh.interleave && !sh.rereads && sh.eligible)  // only light bodies: the registers of `unroll` groups are live together
    for (int n = unroll; n > 1 && !sh.unroll; n /= 2)
      if (sh.ops * n <= 96) sh.unroll = n;
  // a re-reading step walks its groups as (outer, inner): the inner groups leave the largest per-row input where it is
  const auto set_inner = [&]() {
    static const bool split = !(getenv("HB_JIT_SPLIT_LOOP") && atoi(getenv("HB_JIT_SPLIT_LOOP")) == 0);
    static const bool reorder = !(getenv("HB_JIT_DIGIT_ORDER") && atoi(getenv("HB_JIT_DIGIT_ORDER")) == 0);
    sh.inner = 1;
    if (!split || !reorder || !sh.rereads || sh.interleave) return;
    const Table* big = nullptr;
    for (const Level& l : st.levels)
      for (const StepInput& in : l.in)
        if (P.tables[in.table].kind == T_ROW && (!big || P.tables[in.table].size > big->size)) big = &P.tables[in.table];
    if (!big) return;
    int64_t inner = 1, outer = 1;
    for (int v : P.tables[st.out].pscope) {
      if (std::find(sh.block.begin(), sh.block.end(), v) != sh.block.end()) continue;
      (stride_in(*big, v) != 0 ? outer : inner) *= P.net->dims[v];
    }
    // the compiler hoists the loads of `big` out of the inner loop: only when they fit the registers (it spills otherwise)
    int64_t big_loads = sh.V;
    for (int v : sh.block)
      if (stride_in(*big, v) != 0) big_loads *= P.net->dims[v];
    for (size_t l = 0; l < st.levels.size(); ++l)
      for (const StepInput& in : st.levels[l].in)
        if (&P.tables[in.table] == big)
          for (size_t m = 0; m <= l; ++m)
            if (in.lstride[m] != 0) big_loads *= st.levels[m].d;
    static const int64_t max_hoist = getenv("HB_JIT_SPLIT_MAX") ? atoll(getenv("HB_JIT_SPLIT_MAX")) : 32;  // doubles
    if (inner > 1 && outer > 1 && big_loads <= max_hoist) sh.inner = inner;
  };
  set_inner();
  if (!sh.rereads || sh.interleave || block_u <= st.U || !jit_basic(P, st) || st.n_out * terms < block_min_joint) return sh;
  // the block with the fewest loads per output entry among the sets of output variables with at most block_u state
  // combinations; ties go to the smaller block, then to the one that holds w
  const Table& out = P.tables[st.out];
  const std::vector<int>& cand = out.pscope;
  if (cand.size() > 20) return sh;
  const int base_loop = std::max(0, sh.n_loop);
  const double base = (double)block_loads(P, st, sh.block, base_loop) / std::max(1, st.U);
  std::vector<int> best;
  double best_cost = base;
  int64_t best_u = st.U;
  int best_loop = base_loop;
  std::vector<int> cur;
  const auto consider = [&](int64_t u) {
    const int n_loop = loop_levels_for(st, u, block_body);
    if (n_loop < 0 || n_loop > block_loops || body_ops(st, n_loop, u) > MAX_BODY_OPS) return;
    const double cost = (double)block_loads(P, st, cur, n_loop) / (double)u;
    const bool has_w = std::find(cur.begin(), cur.end(), st.w_var) != cur.end();
    const bool best_has_w = std::find(best.begin(), best.end(), st.w_var) != best.end();
    if (cost < best_cost * (1 - 1e-9) || (cost <= best_cost * (1 + 1e-9) && !best.empty() && (u < best_u || (u == best_u && has_w && !best_has_w)))) {
      best = cur;
      best_cost = cost;
      best_u = u;
      best_loop = n_loop;
    }
  };
  const std::function<void(size_t, int64_t)> walk = [&](size_t i, int64_t u) {
    if (i == cand.size()) {
      if (!cur.empty()) consider(u);
      return;
    }
    walk(i + 1, u);
    const int64_t d = P.net->dims[cand[i]];
    if (u * d <= block_u) {
      cur.push_back(cand[i]);
      walk(i + 1, u * d);
      cur.pop_back();
    }
  };
  walk(0, 1);
  // worth it from two thirds of the loads down (a larger block costs registers: one CTA per SM instead of two)
  if (best.empty() || best_cost > base * 0.67) return sh;
  sh.block = best;
  sh.custom = true;
  sh.U = (int)best_u;
  sh.groups = st.n_out / best_u;
  sh.n_loop = best_loop;
  sh.ops = body_ops(st, best_loop, best_u);
  sh.min_blocks = block_minb;
  sh.body_budget = block_body;
  sh.V = block_v;
  sh.eligible = true;
  set_inner();
  return sh;
}

bool jit_zero_add_elidable(const Program& P) {
  static const bool off = getenv("HB_JIT_ZERO_ADD") && atoi(getenv("HB_JIT_ZERO_ADD")) != 0;  // 1: always emit 0 + x
  if (off) return false;
  const auto ok = [](double x) { return x >= 0.0 && !std::signbit(x); };  // false for NaN, negative values and -0.0
  for (int v = 0; v < P.net->n(); ++v)
    if (!P.net->procedural(v))  // a procedural CPT is a quotient of sums of values in [0.05, 1) (procedural.hpp): positive
      for (double x : P.net->values[v])
        if (!ok(x)) return false;
  for (double x : P.extra_values)
    if (!ok(x)) return false;
  return true;
}

std::string jit_source(const Program& P, const std::vector<int>& steps) {
  static const bool reorder = !(getenv("HB_JIT_DIGIT_ORDER") && atoi(getenv("HB_JIT_DIGIT_ORDER")) == 0);
  std::string src = kPreamble;
  // measured on C3 (profiles/r2_step_block.jsonl): without the adds the pass is SLOWER (4.20 against 3.84 ms; the largest looped
  // kernel, step 567, goes from 0.78 to 0.96 ms), while the on-chip kernel gains 4 % on C2 -- so only hb_tree leaves them out
  static const bool step_elide = getenv("HB_JIT_STEP_ELIDE") && atoi(getenv("HB_JIT_STEP_ELIDE")) != 0;
  const bool elide = step_elide && jit_zero_add_elidable(P);
  for (int si : steps) {
    Gen g(P, P.steps[si]);
    const JitShape sh = jit_step_shape(P, P.steps[si]);
    g.interleave = sh.interleave;
    g.unroll = sh.unroll;
    g.elide_zero_add = elide;
    g.reorder_digits = reorder && sh.rereads && !sh.interleave;
    g.split_loop = sh.inner > 1;
    if (sh.custom) {
      g.custom_block = true;
      g.blocked = sh.block;
      g.body_budget = sh.body_budget;
      g.min_blocks = sh.min_blocks;
      g.V = sh.V;
    }
    src += g.kernel("hb_step_" + std::to_string(si));
  }
  return src;
}

// ---- cubin cache --------------------------------------------------------------------------------------
// NVRTC takes seconds for the kernels of one program; the cubin only depends on the generated source and the compiler.
// Three layers: this process (also filled by hb_jt_load from a saved handle), a directory on disk
// (HB_CACHE_DIR, else $XDG_CACHE_HOME/hbayes_b200, else ~/.cache/hbayes_b200; HB_JIT_CACHE=0 turns it off), NVRTC.
namespace {

const char* const kNvrtcOptions[] = {"--gpu-architecture=sm_100a", "--fmad=false", "-default-device", "--std=c++17", "--generate-line-info"};
const char* const kNvrtcOptionsFma[] = {"--gpu-architecture=sm_100a", "--fmad=true", "-default-device", "--std=c++17", "--generate-line-info", "-DHB_FMA"};

std::mutex g_cubin_mutex;
std::map<std::string, std::shared_ptr<const std::string>>& cubin_registry() {
  static std::map<std::string, std::shared_ptr<const std::string>> r;
  return r;
}

std::string cache_dir() {
  if (const char* off = getenv("HB_JIT_CACHE"))
    if (atoi(off) == 0) return "";
  if (const char* d = getenv("HB_CACHE_DIR")) return d;
  if (const char* x = getenv("XDG_CACHE_HOME")) return std::string(x) + "/hbayes_b200";
  if (const char* h = getenv("HOME")) return std::string(h) + "/.cache/hbayes_b200";
  return "";
}

void make_dirs(const std::string& path) {
  for (size_t i = 1; i <= path.size(); ++i)
    if (i == path.size() || path[i] == '/') mkdir(path.substr(0, i).c_str(), 0755);
}

std::string hex64(uint64_t v) {
  char buf[17];
  snprintf(buf, sizeof buf, "%016llx", (unsigned long long)v);
  return buf;
}

}  // namespace

std::string jit_cache_key(const std::string& src, bool fma) {
  uint64_t a = 1469598103934665603ull, b = 0x9E3779B97F4A7C15ull;  // two independent 64-bit FNV-style hashes
  auto feed = [&](const char* p, size_t n) {
    for (size_t i = 0; i < n; ++i) {
      a = (a ^ (unsigned char)p[i]) * 1099511628211ull;
      b = (b + (unsigned char)p[i]) * 0xD6E8FEB86659FD93ull;
      b ^= b >> 29;
    }
  };
  feed(src.data(), src.size());
  if (fma)
    for (const char* o : kNvrtcOptionsFma) feed(o, strlen(o) + 1);
  else
    for (const char* o : kNvrtcOptions) feed(o, strlen(o) + 1);
  static const std::string ver = [] {
    int major = 0, minor = 0;
    void* h = open_first({"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"});
    if (h)
      if (auto f = (int (*)(int*, int*))dlsym(h, "nvrtcVersion")) f(&major, &minor);
    return "nvrtc " + std::to_string(major) + "." + std::to_string(minor) + " hbjit 2";
  }();
  feed(ver.data(), ver.size());
  return hex64(a) + hex64(b);
}

void jit_registry_put(const std::string& key, const std::string& cubin) {
  std::lock_guard<std::mutex> lock(g_cubin_mutex);
  cubin_registry()[key] = std::make_shared<const std::string>(cubin);
}

std::shared_ptr<const std::string> jit_registry_get(const std::string& key) {
  std::lock_guard<std::mutex> lock(g_cubin_mutex);
  auto it = cubin_registry().find(key);
  return it == cubin_registry().end() ? nullptr : it->second;
}

// source -> cubin: registry, then disk, then NVRTC (host-only work: may run on any thread, no CUDA context needed)
std::shared_ptr<const std::string> jit_compile(const std::string& src, bool fma, std::string* why, std::string* key_out, int* origin) {
  auto fail = [&](const std::string& m) -> std::shared_ptr<const std::string> {
    if (why) *why = m;
    if (verbose()) fprintf(stderr, "[hbayes jit] %s\n", m.c_str());
    return nullptr;
  };
  const std::string key = jit_cache_key(src, fma);
  if (key_out) *key_out = key;
  if (auto hit = jit_registry_get(key)) {
    if (origin) *origin = 0;
    return hit;
  }
  const std::string dir = cache_dir();
  const std::string path = dir.empty() ? "" : dir + "/" + key + ".cubin";
  if (!path.empty()) {
    if (FILE* f = fopen(path.c_str(), "rb")) {
      std::string bytes;
      char buf[1 << 16];
      size_t n;
      while ((n = fread(buf, 1, sizeof buf, f)) > 0) bytes.append(buf, n);
      fclose(f);
      if (bytes.size() > 64 && bytes.compare(0, 4, "\x7f" "ELF") == 0) {
        jit_registry_put(key, bytes);
        if (origin) *origin = 1;
        if (verbose()) fprintf(stderr, "[hbayes jit] cubin %s from the disk cache\n", key.c_str());
        return jit_registry_get(key);
      }
    }
  }
  const Nvrtc& rt = nvrtc();
  if (!rt.ok) return fail("libnvrtc not available");
  nvrtcProgram prog = nullptr;
  if (rt.CreateProgram(&prog, src.c_str(), "hbayes_steps.cu", 0, nullptr, nullptr) != 0) return fail("nvrtcCreateProgram failed");
  const int rc = fma ? rt.CompileProgram(prog, 6, kNvrtcOptionsFma) : rt.CompileProgram(prog, 5, kNvrtcOptions);
  if (rc != 0) {
    size_t n = 0;
    rt.GetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) rt.GetProgramLog(prog, &log[0]);
    rt.DestroyProgram(&prog);
    return fail("nvrtc compile error: " + log.substr(0, 2000));
  }
  size_t size = 0;
  if (rt.GetCUBINSize(prog, &size) != 0 || size == 0) {
    rt.DestroyProgram(&prog);
    return fail("nvrtc produced no cubin");
  }
  std::string cubin(size, '\0');
  rt.GetCUBIN(prog, &cubin[0]);
  rt.DestroyProgram(&prog);
  jit_registry_put(key, cubin);
  if (!path.empty()) {  // atomically: another process may be writing the same key
    make_dirs(dir);
    const std::string tmp = path + ".tmp" + std::to_string((long long)getpid());
    if (FILE* f = fopen(tmp.c_str(), "wb")) {
      const bool ok = fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
      fclose(f);
      if (!ok || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());
    }
  }
  if (origin) *origin = 2;
  if (verbose()) fprintf(stderr, "[hbayes jit] compiled %zu bytes of source into %zu bytes of cubin (%s)\n", src.size(), size, key.c_str());
  return jit_registry_get(key);
}

// loads a cubin on the current device
JitModule* jit_load(const std::string& cubin, const std::string& key, std::string* why) {
  const Driver& drv = driver();
  if (!drv.ok) {
    if (why) *why = "libcuda not available";
    return nullptr;
  }
  auto mod = std::make_unique<JitModule>();
  mod->key = key;
  if (drv.ModuleLoadData(&mod->module, cubin.data()) != 0) {
    if (why) *why = "cuModuleLoadData failed";
    mod->module = nullptr;
    return nullptr;
  }
  return mod.release();
}

void* jit_function(JitModule* m, const std::string& name) {
  void* fn = nullptr;
  if (!m || driver().ModuleGetFunction(&fn, m->module, name.c_str()) != 0) return nullptr;
  return fn;
}

const std::string& jit_module_key(const JitModule* m) { return m->key; }

// ================= whole-schedule on-chip kernel ========================================
// changeEvidence -> collect -> distribute -> posterior for every query, for a tile of R evidence rows, in ONE launch:
// the per-row tables live in shared memory ([entry][R], row-fastest like the HBM arena), the steps of one dependency
// level run back to back and levels are separated by __syncthreads; potentials come through L1/L2; only the evidence
// rows and the posteriors touch HBM.  Every step body is the one the per-step kernels use (same Gen), so the
// floating-point sequence of every entry is unchanged.
namespace {
constexpr int64_t TREE_SMEM_MAX = 227 * 1024;
constexpr int64_t TREE_INVARIANT_MAX = 40 * 1024;  // potentials + shared tables staged in shared memory up to this size
constexpr int64_t TREE_STAGE_MAX = 32 * 1024;  // output staging tile; wider results are written straight to HBM
constexpr int64_t TREE_MAX_BODY = 60000;       // chain evaluations in the whole kernel (bounds the compilation time)

int64_t round8(int64_t x) { return (x + 7) / 8 * 8; }
}  // namespace

bool tree_plan(const Program& P, TreePlan& plan, std::string* why) {
  const bool elide = jit_zero_add_elidable(P);
  auto fail = [&](const std::string& m) {
    if (why) *why = m;
    return false;
  };
  plan = TreePlan();
  if (!P.max_steps.empty() || !P.mpe_r.empty()) return fail("maximisation steps");
  if (P.n_reduce > 0 || P.n_gather > 0) return fail("split clique with exchange steps");
  const size_t ns = P.steps.size();
  plan.level.assign(ns, -1);
  plan.off.assign(P.tables.size(), -1);
  std::vector<int> producer(P.tables.size(), -1), last(P.tables.size(), -1);
  int64_t body = 0;
  int n_out_steps = 0;
  for (size_t si = 0; si < ns; ++si) {
    const Step& st = P.steps[si];
    if (st.kind == STEP_OUTPUT) {
      ++n_out_steps;
      if ((int)st.obs_var.size() > 8) return fail("more than 8 observed variables in one query");
      continue;
    }
    if (st.shared) continue;
    if (st.levels.empty() || st.levels[0].is_max || P.tables[st.out].kind != T_ROW) return fail("unsupported step");
    if (st.U > 32 || st.n_red >= ((int64_t)1 << 24)) return fail("step too large");
    const int64_t ops = jit_unrolled_ops(st);
    if (ops < 0) return fail("step body too large to unroll");
    body += ops;
    int lv = 0;
    auto reads = [&](int id) {
      const Table& t = P.tables[id];
      if (t.kind == T_ONE) return false;
      if (t.kind == T_ROW) {
        if (producer[id] < 0) return false;
        lv = std::max(lv, plan.level[producer[id]] + 1);
      }
      return true;
    };
    for (const Level& l : st.levels)
      for (const StepInput& in : l.in)
        if (!reads(in.table)) return fail("unsupported input table");
    for (int id : st.scalars)
      if (!reads(id)) return fail("unsupported scalar table");
    plan.level[si] = lv;
    producer[st.out] = (int)si;
    plan.n_levels = std::max(plan.n_levels, lv + 1);
  }
  if (n_out_steps == 0) return fail("no query");
  // Results wider than this per row are not worth the on-chip kernel: they cannot be staged for coalesced rows (measured on
  // the learnEM family-query program of C2, 1226 doubles per sample: 11.7 ms per 2^18 samples against 10.2 ms on the
  // precompiled and 8.9 ms on the per-step generated kernels -- that pass is bound by writing its 2.6 GB result matrix)
  if (P.out_width > 512) return fail("result rows too wide");
  if (body > TREE_MAX_BODY) return fail("program too large for one kernel");
  // liveness at level granularity: a table may be overwritten from the level AFTER its last reader's
  for (size_t si = 0; si < ns; ++si) {
    const Step& st = P.steps[si];
    if (st.kind == STEP_OUTPUT) {
      if (P.tables[st.result].kind == T_ROW) last[st.result] = plan.n_levels;  // read by the output phase
      continue;
    }
    if (plan.level[si] < 0) continue;
    for (const Level& l : st.levels)
      for (const StepInput& in : l.in) last[in.table] = std::max(last[in.table], plan.level[si]);
    for (int id : st.scalars) last[id] = std::max(last[id], plan.level[si]);
  }
  {
    std::vector<std::pair<int64_t, int64_t>> free_list;  // (offset, size) sorted by offset
    int64_t top = 0;
    auto release = [&](int64_t off, int64_t size) {
      auto it = std::lower_bound(free_list.begin(), free_list.end(), std::make_pair(off, (int64_t)0));
      it = free_list.insert(it, {off, size});
      if (it + 1 != free_list.end() && it->first + it->second == (it + 1)->first) it->second += (it + 1)->second, free_list.erase(it + 1);
      if (it != free_list.begin() && (it - 1)->first + (it - 1)->second == it->first) (it - 1)->second += it->second, it = free_list.erase(it), --it;
      if (it->first + it->second == top) top = it->first, free_list.erase(it);
    };
    std::vector<int> live;
    for (int lv = 0; lv < plan.n_levels; ++lv) {
      for (size_t i = 0; i < live.size();) {
        if (last[live[i]] < lv) {
          release(plan.off[live[i]], P.tables[live[i]].size);
          live[i] = live.back(), live.pop_back();
        } else
          ++i;
      }
      std::vector<int> born;
      for (size_t si = 0; si < ns; ++si)
        if (plan.level[si] == lv) born.push_back(P.steps[si].out);
      std::stable_sort(born.begin(), born.end(), [&](int a, int b) { return P.tables[a].size > P.tables[b].size; });
      for (int id : born) {
        const int64_t size = P.tables[id].size;
        size_t best = free_list.size();  // best fit
        for (size_t i = 0; i < free_list.size(); ++i)
          if (free_list[i].second >= size && (best == free_list.size() || free_list[i].second < free_list[best].second)) best = i;
        if (best < free_list.size()) {
          plan.off[id] = free_list[best].first;
          free_list[best].first += size, free_list[best].second -= size;
          if (free_list[best].second == 0) free_list.erase(free_list.begin() + best);
        } else
          plan.off[id] = top, top += size;
        plan.arena_entries = std::max(plan.arena_entries, top);
        live.push_back(id);
      }
    }
    plan.arena_entries = std::max<int64_t>(plan.arena_entries, 1);
  }
  const int64_t NV = P.net->n(), NS = (int64_t)P.slices.size(), W = P.out_width;
  // the batch-invariant tables (potential pool, shared arena) are copied to shared memory once per CTA when they are small
  const int64_t inv_entries = P.pot_entries + P.shared_entries;
  plan.invariant_entries = inv_entries * 8 <= TREE_INVARIANT_MAX ? inv_entries : 0;
  auto bytes = [&](int64_t R, bool staged) {
    return plan.arena_entries * R * 8 + plan.invariant_entries * 8 + round8(NS * R * 4) + (staged ? R * (W | 1) * 8 : 0) + round8(R * NV);
  };
  int R = 0;
  if (const char* env = getenv("HB_TREE_ROWS")) R = std::max(2, atoi(env) / 2 * 2);
  bool staged = true;
  if (!R)
    for (int cand : {32, 16, 8, 4, 2}) {
      staged = cand * (W | 1) * 8 <= TREE_STAGE_MAX;
      if (bytes(cand, staged) <= TREE_SMEM_MAX) {
        R = cand;
        break;
      }
    }
  if (!R) return fail("the per-row tables of two rows do not fit shared memory");
  staged = R * (W | 1) * 8 <= TREE_STAGE_MAX;
  if (bytes(R, staged) > TREE_SMEM_MAX) return fail("HB_TREE_ROWS does not fit shared memory");
  plan.R = R;
  plan.stage_pitch = staged ? (W | 1) : 0;
  plan.smem_bytes = (size_t)bytes(R, staged);
  // threads: enough for the widest level (output entries x row pairs), between 64 and 256
  int64_t widest = 1;
  {
    std::vector<int64_t> items(plan.n_levels, 0);
    for (size_t si = 0; si < ns; ++si)
      if (plan.level[si] >= 0) items[plan.level[si]] += P.steps[si].n_out * (R / 2);
    for (int64_t x : items) widest = std::max(widest, x);
  }
  int T = 64;  // measured on C2 (profiles/README.md): 256, 384 and 512 threads are within 3 %; 384 is the best
  while (T < 384 && T < widest) T = T < 256 ? T * 2 : 384;
  if (const char* env = getenv("HB_TREE_THREADS")) T = std::max(32, std::min(1024, atoi(env) / 32 * 32));
  plan.T = T;
  int minb = (int)std::min<int64_t>(TREE_SMEM_MAX / (int64_t)plan.smem_bytes, 2048 / T);
  minb = std::max(1, std::min(minb, 65536 / (T * 128)));  // leave at least 128 registers per thread
  if (const char* env = getenv("HB_TREE_MINBLOCKS")) minb = std::max(1, atoi(env));
  plan.min_blocks = minb;
  for (size_t si = 0; si < ns; ++si) {
    const Step& st = P.steps[si];
    if (plan.level[si] < 0) continue;
    int64_t terms = st.n_out;
    for (size_t l = 0; l < st.levels.size(); ++l) {
      terms *= st.levels[l].d;
      const int64_t k = (int64_t)st.levels[l].in.size() + (l + 1 < st.levels.size() ? 1 : 0) + (l == 0 && !st.scalars.empty() ? 1 : 0);
      plan.ops_mul += terms * std::max<int64_t>(0, k - 1);
      // the reference adds every term to a sum that starts at 0; without the `0 +` (jit_zero_add_elidable) a sum of d terms
      // costs d - 1 adds -- counted as executed, so that the fp64 rate of the bench line is the kernel's
      plan.ops_add += elide ? terms - terms / st.levels[l].d : terms;
    }
    if (st.scalars.size() > 1) plan.ops_mul += (int64_t)st.scalars.size() - 1;
  }
  return true;
}

std::string tree_source(const Program& P, const TreePlan& plan, int64_t* loads_per_row, int64_t* stores_per_row) {
  const bool elide = jit_zero_add_elidable(P);
  const Network& net = *P.net;
  const int R = plan.R, T = plan.T;
  const int64_t NV = net.n(), NS = (int64_t)P.slices.size(), W = P.out_width, WP = plan.stage_pitch;
  std::ostringstream o;
  o << kPreamble;
  auto table = [&](const char* type, const char* name, const std::vector<int64_t>& v) {
    o << "__device__ const " << type << " " << name << "[" << std::max<size_t>(v.size(), 1) << "] = {";  // read with divergent indices: L1, not the constant bank
    for (size_t i = 0; i < v.size(); ++i) o << (i ? "," : "") << v[i];
    if (v.empty()) o << "0";
    o << "};\n";
  };
  std::vector<int64_t> obs, dims, sl_begin = {0}, sl_var, sl_stride;
  for (int v = 0; v < NV; ++v) obs.push_back(P.fixed_state[v] >= 0 ? 2 + P.fixed_state[v] : (P.is_observed[v] ? 1 : 0)), dims.push_back(net.dims[v]);
  for (const Slice& sl : P.slices) {
    for (size_t j = 0; j < sl.vars.size(); ++j) sl_var.push_back(sl.vars[j]), sl_stride.push_back(sl.strides[j]);
    sl_begin.push_back((int64_t)sl_var.size());
  }
  {  // valid states of every variable in a row of this program: unobserved -> must be negative, observed -> 0..dim-1,
     // fixed (split clique) -> exactly that state
    std::vector<int64_t> chk;
    for (int v = 0; v < NV; ++v) chk.push_back(obs[v] == 0 ? -128 : obs[v] == 1 ? 0 : obs[v] - 2);
    for (int v = 0; v < NV; ++v) chk.push_back(obs[v] == 0 ? -1 : obs[v] == 1 ? dims[v] - 1 : obs[v] - 2);
    table("signed char", "CHK", chk);
  }
  o << "__device__ __forceinline__ int hb_state(int s, int dim) { return (unsigned)s < (unsigned)dim ? s : 0; }\n";
  // queries: Q = {kind (0 one, 1 potential pool, 2 shared arena, 3 per-row table), base, slice, n_phys, n_full, column,
  // map offset, order offset, n_obs}, then per observed query variable (var, stride, dim)
  std::vector<int64_t> q, qobs, qmap;
  int nq = 0;
  for (const Step& st : P.steps) {
    if (st.kind != STEP_OUTPUT) continue;
    const Table& t = P.tables[st.result];
    const int64_t kind = t.kind == T_ONE ? 0 : t.kind == T_POT ? 1 : t.kind == T_SHARED ? 2 : 3;
    const int64_t map_off = (int64_t)qmap.size();
    for (int64_t f = 0; f < st.n_full; ++f) qmap.push_back(st.full_map[(size_t)f]);
    const int64_t order_off = (int64_t)qmap.size();
    int64_t n_phys = 0;  // the stored entries in the order the full layout visits them (norm is a left fold over that order)
    for (int64_t f = 0; f < st.n_full; ++f) {
      bool first_state = true;
      for (size_t j = 0; j < st.obs_var.size(); ++j) first_state &= ((f / st.obs_stride[j]) % st.obs_dim[j]) == 0;
      if (first_state) qmap.push_back(st.full_map[(size_t)f]), ++n_phys;
    }
    for (int64_t x : {kind, kind == 3 ? plan.off[st.result] : t.off, (int64_t)(t.kind == T_POT ? t.slice_slot : -1), n_phys, st.n_full, st.out_col, map_off,
                      order_off, (int64_t)st.obs_var.size()})
      q.push_back(x);
    for (int j = 0; j < 8; ++j) {
      const bool has = j < (int)st.obs_var.size();
      qobs.push_back(has ? st.obs_var[j] : 0), qobs.push_back(has ? st.obs_stride[j] : 1), qobs.push_back(has ? st.obs_dim[j] : 1);
    }
    ++nq;
  }
  table("int", "Q", q);
  table("int", "QOBS", qobs);
  table("unsigned", "QMAP", qmap);
  std::ostringstream fns;  // one device function per step, defined before the kernel
  o << "@@STEP_FUNCTIONS@@";
  const int64_t inv_at = plan.arena_entries * R * 8, rowoff_at = inv_at + plan.invariant_entries * 8, stage_at = rowoff_at + round8(NS * R * 4),
                ev_at = stage_at + (WP ? R * WP * 8 : 0);
  o << "extern \"C\" __global__ void __launch_bounds__(" << T << ", " << plan.min_blocks << ") hb_tree(const double* __restrict__ g_pots,\n"
       "    const double* __restrict__ g_shared, const signed char* __restrict__ ev, double* __restrict__ out, long long n_rows,\n"
       "    int* __restrict__ flag, int normalise) {\n";
  o << "  extern __shared__ __align__(16) unsigned char smem_raw[];\n";
  o << "  double* const sm = reinterpret_cast<double*>(smem_raw);\n";
  o << "  const double* const arena = sm;\n  double* const arena_out = sm;\n";
  if (plan.invariant_entries > 0) {  // potentials and shared tables: HBM -> shared memory once per CTA (the CTA is persistent)
    o << "  double* const inv = reinterpret_cast<double*>(smem_raw + " << inv_at << ");\n";
    o << "  for (unsigned i = threadIdx.x; i < " << P.pot_entries << "u; i += " << T << "u) inv[i] = g_pots[i];\n";
    o << "  for (unsigned i = threadIdx.x; i < " << P.shared_entries << "u; i += " << T << "u) inv[" << P.pot_entries << "u + i] = g_shared[i];\n";
    o << "  const double* const pots = inv;\n  const double* const shared = inv + " << P.pot_entries << ";\n";
  } else
    o << "  const double* const pots = g_pots;\n  const double* const shared = g_shared;\n";
  o << "  int* const rowoff = reinterpret_cast<int*>(smem_raw + " << rowoff_at << ");\n";
  if (WP) o << "  double* const stage = reinterpret_cast<double*>(smem_raw + " << stage_at << ");\n";
  o << "  signed char* const evs = reinterpret_cast<signed char*>(smem_raw + " << ev_at << ");\n";
  o << "  constexpr unsigned RT = " << R << "u;\n";
  o << "  const unsigned tid = threadIdx.x;\n";
  o << "  const unsigned rowA = (tid % " << R / 2 << "u) * 2u, slotA = tid / " << R / 2 << "u, rowB = tid % " << R << "u, slotB = tid / " << R << "u;\n";
  const int64_t ept = (R * NV + T - 1) / T;  // evidence bytes per thread and tile
  for (int64_t k = 0; k < ept; ++k)
    o << "  signed char pre" << k << " = (tid + " << k * T << "u < " << R * NV << "u && (long long)blockIdx.x * " << R * NV << "ll + tid + " << k * T << "u < n_rows * " << NV
      << "ll) ? ev[(long long)blockIdx.x * " << R * NV << "ll + tid + " << k * T << "u] : (signed char)-1;\n";
  o << "  for (long long tile = blockIdx.x; tile * " << R << " < n_rows; tile += gridDim.x) {\n";
  o << "    const long long row0 = tile * " << R << ";\n";
  if (getenv("HB_TREE_PROFILE") && atoi(getenv("HB_TREE_PROFILE")) != 0) o << "    long long t0 = clock64();\n";
  o << "    const int nr = (int)(n_rows - row0 < " << R << " ? n_rows - row0 : " << R << ");\n";
  // Evidence of this tile -> shared memory.  It was loaded into registers while the previous tile was computed (the CTA is
  // persistent), so its HBM latency is hidden.  Pad rows of the last tile: unobserved everywhere, slice 0, never written.
  for (int64_t k = 0; k < ept; ++k)
    o << "    if (tid + " << k * T << "u < " << R * NV << "u) evs[tid + " << k * T << "u] = pre" << k << ";\n";
  o << "    __syncthreads();\n";
  o << "    {\n      const long long nt = (tile + gridDim.x) * " << R * NV << "ll;\n";
  for (int64_t k = 0; k < ept; ++k)
    o << "      pre" << k << " = (tid + " << k * T << "u < " << R * NV << "u && nt + tid + " << k * T << "u < n_rows * " << NV << "ll) ? ev[nt + tid + " << k * T
      << "u] : (signed char)-1;\n";
  o << "    }\n";
  // consistency of every row with the compiled observed set: state in [CHK[v], CHK[NV + v]]
  o << "    for (unsigned i = tid; i < (unsigned)nr * " << NV << "u; i += " << T << "u) {\n"
       "      const int s = evs[i];\n"
       "      if (s < CHK[i % " << NV << "u] || s > CHK[" << NV << "u + i % " << NV << "u]) atomicOr(flag, 1);\n    }\n";
  if (NS > 0) {  // per-row entry offset of every sliced potential (factorFromInstantiation folded into the addressing)
    o << "    for (unsigned i = tid; i < " << NS * R << "u; i += " << T << "u) {\n"
         "      const unsigned j = i / " << R << "u, r = i % " << R << "u;\n"
         "      const signed char* const e = evs + r * " << NV << "u;\n"
         "      int off = 0;\n"
         "      switch (j) {\n";
    for (int64_t j = 0; j < NS; ++j) {
      const Slice& sl = P.slices[(size_t)j];
      o << "        case " << j << ": off = ";
      for (size_t t = 0; t < sl.vars.size(); ++t)
        o << (t ? " + " : "") << sl.strides[t] << " * hb_state(e[" << sl.vars[t] << "], " << net.dims[sl.vars[t]] << ")";
      o << "; break;\n";
    }
    o << "      }\n      rowoff[i] = off;\n    }\n";
  }
  o << "    __syncthreads();\n";
  int64_t stat_loads = 0, stat_items = 0;  // shared-memory loads per evidence row; work items per tile
  static const bool profile = getenv("HB_TREE_PROFILE") && atoi(getenv("HB_TREE_PROFILE")) != 0;
  auto lap = [&](int slot) {  // HB_TREE_PROFILE=1: cycles per phase, summed over the CTAs' tiles, behind the error flag
    if (profile) o << "    if (tid == 0) { const long long t1 = clock64(); atomicAdd(reinterpret_cast<unsigned long long*>(flag + 2) + " << slot
                   << ", (unsigned long long)(t1 - t0)); t0 = t1; }\n";
  };
  lap(0);
  static const double grain = getenv("HB_TREE_GRAIN") ? atof(getenv("HB_TREE_GRAIN")) : 1.0;  // tuning: items wanted per thread
  for (int lv = 0; lv < plan.n_levels; ++lv) {
    o << "    // ---- dependency level " << lv << " ----\n";
    int64_t rot = 0;
    for (size_t si = 0; si < P.steps.size(); ++si) {
      if (plan.level[si] != lv) continue;
      const Step& st = P.steps[si];
      // the coarsest cut of the step that still keeps the threads busy: blocks of all states of w x row pairs when they give
      // at least half the threads an item (measured on C2: threshold 1.0 T -> 6.40 ms, 0.5 T -> 6.32 ms, 0.25 T -> 9.66 ms
      // per 2^20 rows: few long items cost more than the loads they save), else single entries x row pairs, else single
      // entries x single rows (narrow steps: latency is the bound and every thread should have something to do)
      static const double group_grain = getenv("HB_TREE_GROUP_GRAIN") ? atof(getenv("HB_TREE_GROUP_GRAIN")) : 0.5;
      int v = 2;
      std::vector<int> block;
      if (st.w_var >= 0 && st.n_red * (R / 2) >= group_grain * T)
        block = {st.w_var};
      else if (st.n_out * (R / 2) < grain * T)
        v = 1;
      Gen g(P, st);
      g.tree_off = &plan.off;
      g.elide_zero_add = elide;
      g.tree_step((int)si, R, T, (int)(rot % T), block, v);
      fns << g.o.str();
      o << g.call;
      rot += g.tree_items;
      stat_loads += g.n_loads * (g.tree_items / (v == 2 ? R / 2 : R));
      stat_items += g.tree_items;
    }
    o << "    __syncthreads();\n";
    lap(1 + lv);
  }
  // factorNorm + factorDivide + expansion to the query scope (what output_kernel of engine.cu does, reading shared
  // memory): thread = (query, row), specialised per query -- norm = 0 + x0 + x1 + ... over the layout order, every entry
  // (1.0 / norm) * value, 0.0 where the entry contradicts the evidence.
  o << "    for (unsigned i = tid; i < " << (int64_t)nq * R << "u; i += " << T << "u) {\n"
       "      const unsigned r = i % " << R << "u, qi = i / " << R << "u;\n"
       "      const signed char* const e = evs + r * " << NV << "u;\n";
  if (WP)
    o << "      double* const op = stage + r * " << WP << "u;\n";
  else
    o << "      if ((int)r >= nr) continue;\n      double* const op = out + (row0 + r) * " << W << "ll;\n";
  o << "      switch (qi) {\n";
  {
    int qi = 0;
    bool any_generic = false;
    for (const Step& st : P.steps) {
      if (st.kind != STEP_OUTPUT) continue;
      const int this_q = qi++;
      if (st.n_full > 64) {
        any_generic = true;
        continue;
      }
      const Table& t = P.tables[st.result];
      o << "        case " << this_q << ": {\n";
      auto value = [&](int64_t k) -> std::string {  // stored entry k of the result
        if (t.kind == T_ONE) return "1.0";
        if (t.kind == T_ROW) return "sm[(size_t)" + std::to_string(plan.off[st.result] + k) + " * RT + r]";
        if (t.kind == T_SHARED) return "shared[" + std::to_string(t.off + k) + "]";
        return "gp[" + std::to_string(k) + "]";
      };
      if (t.kind == T_POT) {
        o << "          const double* const gp = pots + " << t.off;
        if (t.slice_slot >= 0) o << " + rowoff[" << t.slice_slot * R << "u + r]";
        o << ";\n";
      }
      std::map<int64_t, std::string> have;
      auto use = [&](int64_t k) {
        auto it = have.find(k);
        if (it != have.end()) return it->second;
        const std::string n = "v" + std::to_string(k);
        o << "          const double " << n << " = " << value(k) << ";\n";
        return have[k] = n;
      };
      std::string norm = "0.0";
      int n_terms = 0;
      for (int64_t f = 0; f < st.n_full; ++f) {
        bool first_state = true;
        for (size_t j = 0; j < st.obs_var.size(); ++j) first_state &= ((f / st.obs_stride[j]) % st.obs_dim[j]) == 0;
        if (!first_state) continue;
        const std::string v = use(st.full_map[(size_t)f]);
        if (elide && n_terms == 0) {  // 0 + x == x (see Gen::elide_zero_add)
          ++n_terms;
          norm = v;
          continue;
        }
        const std::string n = "n" + std::to_string(n_terms++);
        o << "          const double " << n << " = __dadd_rn(" << norm << ", " << v << ");\n";
        norm = n;
      }
      // 1.0 / norm, correctly rounded: __drcp_rn gives the same bits as __ddiv_rn(1.0, norm) without the division's slow path
      o << "          const double inv = normalise ? __drcp_rn(" << norm << ") : 1.0;\n";
      for (int64_t f = 0; f < st.n_full; ++f) {
        std::string match;
        for (size_t j = 0; j < st.obs_var.size(); ++j)
          match += (match.empty() ? "" : " && ") + ("e[" + std::to_string(st.obs_var[j]) + "] == " + std::to_string((f / st.obs_stride[j]) % st.obs_dim[j]));
        const std::string v = use(st.full_map[(size_t)f]);
        o << "          op[" << st.out_col + f << "] = __dmul_rn(inv, " << (match.empty() ? v : "(" + match + ") ? " + v + " : 0.0") << ");\n";
      }
      o << "        } break;\n";
    }
    if (any_generic) {  // very wide queries (joint family posteriors): descriptor-driven
      o << "        default: {\n"
           "          const int* const d = Q + qi * 9;\n"
           "          const int kind = d[0], n_phys = d[3], n_full = d[4], n_obs = d[8];\n"
           "          const double* gp = nullptr;\n"
           "          if (kind == 1) gp = pots + d[1] + (d[2] >= 0 ? rowoff[d[2] * " << R << " + r] : 0);\n"
           "          else if (kind == 2) gp = shared + d[1];\n"
           "          const double* const sp = sm + (size_t)d[1] * RT + r;\n"
           "          const unsigned* const order = QMAP + d[7];\n"
           "          const unsigned* const fmap = QMAP + d[6];\n"
           "          double norm = 0.0;\n"
           "          for (int j = 0; j < n_phys; ++j) norm = __dadd_rn(norm, kind == 3 ? sp[(size_t)order[j] * RT] : (kind == 0 ? 1.0 : gp[order[j]]));\n"
           "          const double inv = normalise ? __drcp_rn(norm) : 1.0;\n"
           "          int st[8];\n"
           "          for (int j = 0; j < n_obs; ++j) st[j] = e[QOBS[qi * 24 + j * 3]];\n"
           "          for (int f = 0; f < n_full; ++f) {\n"
           "            bool match = true;\n"
           "            for (int j = 0; j < n_obs; ++j) match &= ((f / QOBS[qi * 24 + j * 3 + 1]) % QOBS[qi * 24 + j * 3 + 2]) == st[j];\n"
           "            const double v = match ? (kind == 3 ? sp[(size_t)fmap[f] * RT] : (kind == 0 ? 1.0 : gp[fmap[f]])) : 0.0;\n"
           "            op[d[5] + f] = __dmul_rn(inv, v);\n          }\n"
           "        } break;\n";
    }
  }
  o << "      }\n    }\n";
  if (WP) {
    o << "    __syncthreads();\n";
    o << "    for (unsigned i = tid; i < (unsigned)nr * " << W << "u; i += " << T << "u) out[row0 * " << W << "ll + i] = stage[(i / " << W << "u) * " << WP << "u + i % " << W
      << "u];\n";
  }
  lap(1 + plan.n_levels);
  o << "  }\n}\n";
  o << "// per evidence row: " << stat_loads << " table loads, " << plan.ops_mul << " fp64 multiplies, " << plan.ops_add << " fp64 adds; " << stat_items
    << " work items per tile of " << R << " rows\n";
  if (loads_per_row) *loads_per_row = stat_loads;
  if (stores_per_row) {
    *stores_per_row = 0;
    for (size_t si = 0; si < P.steps.size(); ++si)
      if (plan.level[si] >= 0) *stores_per_row += P.steps[si].n_out;
  }
  std::string src = o.str();
  const size_t at = src.find("@@STEP_FUNCTIONS@@");
  src.replace(at, strlen("@@STEP_FUNCTIONS@@"), fns.str());
  return src;
}

void jit_destroy(JitModule* m) { delete m; }

bool jit_launch(void* fn, unsigned gx, unsigned gy, unsigned bx, unsigned by, void** params, void* stream) {
  return driver().LaunchKernel(fn, gx, gy, 1, bx, by, 1, 0, stream, params, nullptr) == 0;
}

bool jit_launch_smem(void* fn, unsigned grid, unsigned block, unsigned smem_bytes, void** params, void* stream) {
  return driver().LaunchKernel(fn, grid, 1, 1, block, 1, 1, smem_bytes, stream, params, nullptr) == 0;
}

bool jit_set_dynamic_smem(void* fn, int bytes) {
  return driver().FuncSetAttribute && driver().FuncSetAttribute(fn, 8 /* CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES */, bytes) == 0;
}

int jit_occupancy(void* fn, int block, size_t smem_bytes) {
  int n = 0;
  if (!driver().Occupancy || driver().Occupancy(&n, fn, block, smem_bytes) != 0) return 0;
  return n;
}

}  // namespace hb
