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

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <initializer_list>
#include <map>
#include <memory>
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
  Gen(const Program& p, const Step& s) : P(p), st(s) {}

  std::string load(const In& x, int64_t off) {
    auto key = std::make_pair(x.id, off);
    auto it = loaded.find(key);
    if (it != loaded.end()) return it->second;
    std::string name = "x" + std::to_string(x.id) + "_" + std::to_string(off) + (scope ? "_" + std::to_string(scope) : "");
    const Table& t = *x.t;
    const std::string& q = ptr[x.id];
    if (t.kind == T_ROW)
      o << "    const double2 " << name << " = *reinterpret_cast<const double2*>(" << q << " + (size_t)" << off << " * RT);\n";
    else if (t.kind == T_POT && t.slice_slot >= 0)
      o << "    const double2 " << name << " = make_double2(" << q << "a[" << off << "], " << q << "b[" << off << "]);\n";
    else
      o << "    const double " << name << "s = __ldg(" << q << " + " << off << "); const double2 " << name << " = make_double2(" << name << "s, "
        << name << "s);\n";
    loaded[key] = name;
    return name;
  }
  std::string mul(const std::string& a, const std::string& b) {
    std::string n = "m" + std::to_string(++tmp);
    o << "    const double2 " << n << " = make_double2(__dmul_rn(" << a << ".x, " << b << ".x), __dmul_rn(" << a << ".y, " << b << ".y));\n";
    return n;
  }
  std::string add(const std::string& a, const std::string& b) {
    std::string n = "s" + std::to_string(++tmp);
    o << "    const double2 " << n << " = make_double2(__dadd_rn(" << a << ".x, " << b << ".x), __dadd_rn(" << a << ".y, " << b << ".y));\n";
    return n;
  }
  // chain of the inputs of level li for entry u (offsets of the unrolled levels from `bound`), times the inner value
  std::string term(size_t li, int u, const std::vector<int>& bound, const std::vector<std::string>& inner) {
    std::string chain;
    for (size_t k = ins.size(); k-- > 0;) {
      const In& x = ins[k];
      if (x.level != (int)li) continue;
      int64_t off = (int64_t)u * x.in->wstride;
      for (size_t m = (size_t)n_loop; m <= li; ++m) off += (int64_t)bound[m] * x.in->lstride[m];  // looped levels live in the pointer
      const std::string v = load(x, off);
      chain = chain.empty() ? v : mul(v, chain);
    }
    if (!inner.empty()) chain = chain.empty() ? inner[u] : mul(inner[u], chain);
    return chain;
  }
  // value_l of the U lock-step entries, with the summed states of the outer levels bound
  std::vector<std::string> level(size_t li, std::vector<int>& bound) {
    const Level& lv = st.levels[li];
    std::vector<std::string> acc(U);
    if ((int)li < n_loop) {  // run-time loop over the states of this level's summed variable
      for (int u = 0; u < U; ++u) {
        acc[u] = "a" + std::to_string(li) + "_" + std::to_string(u);
        o << "    double2 " << acc[u] << " = zero;\n";
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
        else if (t.kind == T_POT && t.slice_slot >= 0)
          o << "    const double* const " << q << "a = " << ptr[x.id] << "a + (" << step << ");\n"
            << "    const double* const " << q << "b = " << ptr[x.id] << "b + (" << step << ");\n";
        else
          o << "    const double* const " << q << " = " << ptr[x.id] << " + (" << step << ");\n";
        ptr[x.id] = q;
      }
      std::vector<std::string> inner;
      if (li + 1 < st.levels.size()) inner = level(li + 1, bound);
      for (int u = 0; u < U; ++u) {
        const std::string c = term(li, u, bound, inner);
        o << "    " << acc[u] << " = make_double2(__dadd_rn(" << acc[u] << ".x, " << c << ".x), __dadd_rn(" << acc[u] << ".y, " << c << ".y));\n";
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
        acc[u] = acc[u].empty() ? add("zero", chain) : add(acc[u], chain);
      }
    }
    return acc;
  }

  std::string kernel(const std::string& name) {
    const Network& net = *P.net;
    const Table& out = P.tables[st.out];
    U = st.U;
    int id = 0;
    for (size_t l = 0; l < st.levels.size(); ++l)
      for (const StepInput& in : st.levels[l].in) ins.push_back(In{id++, (int)l, &P.tables[in.table], &in});
    n_loop = jit_loop_levels(st);
    ptr.clear();
    for (const In& x : ins) ptr.push_back("p" + std::to_string(x.id));
    std::vector<int> red;  // output variables other than the blocking one, stored row-major (last fastest), then w
    for (int v : out.pscope)
      if (v != st.w_var) red.push_back(v);
    // two CTAs per SM (<= 128 registers): measured on C2, 1 -> 86.2, 2 -> 89.7, 3 -> 75.9 M queries/s
    o << "extern \"C\" __global__ void __launch_bounds__(256, " << (getenv("HB_JIT_MINBLOCKS") ? atoi(getenv("HB_JIT_MINBLOCKS")) : 2) << ") " << name
      << "(const double* __restrict__ arena, double* __restrict__ arena_out, const double* __restrict__ pots,\n"
         "    const double* __restrict__ shared, const int* __restrict__ rowoff, int nr, long long RT) {\n";
    o << "  const int row = (blockIdx.x * blockDim.x + threadIdx.x) * 2;\n  if (row >= nr) return;\n";
    o << "  const unsigned lanes = gridDim.y * blockDim.y, lane = blockIdx.y * blockDim.y + threadIdx.y;\n";
    o << "  const unsigned per = (" << st.n_red << "u + lanes - 1) / lanes;\n";
    o << "  const unsigned g_end = min(" << st.n_red << "u, lane * per + per);\n";
    o << "  const double2 zero = make_double2(0.0, 0.0);\n";
    for (const In& x : ins) {  // row pointers that do not depend on the group
      const Table& t = *x.t;
      if (t.kind == T_ROW)
        o << "  const double* const b" << x.id << " = arena + (size_t)" << t.off << " * RT + row;\n";
      else if (t.kind == T_POT && t.slice_slot >= 0)
        o << "  const double* const b" << x.id << "a = pots + " << t.off << " + rowoff[(size_t)" << t.slice_slot << " * RT + row];\n"
          << "  const double* const b" << x.id << "b = pots + " << t.off << " + rowoff[(size_t)" << t.slice_slot << " * RT + row + 1];\n";
      else
        o << "  const double* const b" << x.id << " = " << (t.kind == T_POT ? "pots" : "shared") << " + " << t.off << ";\n";
    }
    o << "  double* const outp = arena_out + (size_t)" << out.off << " * RT + row;\n";  // same arena: the output region is never read here
    o << "  for (unsigned g = lane * per; g < g_end; ++g) {\n";
    if (!red.empty()) o << "    unsigned rem = g;\n";
    for (size_t j = red.size(); j-- > 0;)
      o << "    const unsigned d" << red[j] << " = rem % " << net.dims[red[j]] << "u;" << (j ? " rem /= " + std::to_string(net.dims[red[j]]) + "u;" : "")
        << "\n";
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
      else if (t.kind == T_POT && t.slice_slot >= 0)
        o << "    const double* const p" << x.id << "a = b" << x.id << "a + (" << base << ");\n"
          << "    const double* const p" << x.id << "b = b" << x.id << "b + (" << base << ");\n";
      else
        o << "    const double* const p" << x.id << " = b" << x.id << " + (" << base << ");\n";
    }
    std::vector<int> bound(st.levels.size(), 0);
    const std::vector<std::string> res = level(0, bound);
    for (int u = 0; u < U; ++u)
      o << "    __stcs(reinterpret_cast<double2*>(outp + (size_t)(g * " << U << "u + " << u << "u) * RT), " << res[u] << ");\n";
    o << "  }\n}\n\n";
    return o.str();
  }
};

}  // namespace

// chain evaluations in the generated body of a group when the `n_loop` outermost levels stay run-time loops
// (about 150 bytes of source each)
static int64_t body_ops(const Step& st, int n_loop) {
  int64_t ops = 0, terms = 1;
  for (size_t l = 0; l < st.levels.size(); ++l) {
    if ((int)l >= n_loop) terms *= st.levels[l].d;
    ops += terms * (int64_t)std::max<size_t>(1, st.levels[l].in.size());
  }
  return ops * st.U;
}
constexpr int64_t MAX_BODY_OPS = 600;  // compile time grows faster than linearly with the size of a straight-line kernel

// how many outermost levels stay run-time loops so that the unrolled rest fits MAX_BODY_OPS (-1: not even the
// innermost level alone fits)
int jit_loop_levels(const Step& st) {
  for (int n = 0; n < (int)st.levels.size(); ++n)
    if (body_ops(st, n) <= MAX_BODY_OPS) return n;
  return -1;
}
int64_t jit_unrolled_ops(const Step& st) {
  const int n = jit_loop_levels(st);
  return n < 0 ? -1 : body_ops(st, n);
}

bool jit_eligible(const Program& P, const Step& st) {
  if (st.kind != STEP_PRODSUM || st.shared || !st.scalars.empty() || st.levels.empty() || st.levels[0].is_max) return false;
  if (P.tables[st.out].kind != T_ROW || st.U > 8 || st.n_inputs() < 1) return false;
  if (st.n_red >= ((int64_t)1 << 31)) return false;
  for (const Level& l : st.levels)
    for (const StepInput& in : l.in)
      if (P.tables[in.table].kind == T_ONE) return false;
  // Steps whose body only fits with run-time loops around it can be generated too (HB_JIT_LOOPS=1; parity-checked on
  // C3), but at two CTAs per SM they spill and measured no faster than the templated kernels (C3: 6.6 vs 6.2 ms), so
  // by default only fully unrolled steps are specialised.
  static const bool loops_on = getenv("HB_JIT_LOOPS") && atoi(getenv("HB_JIT_LOOPS")) != 0;
  const int n_loop = jit_loop_levels(st);
  return n_loop == 0 || (loops_on && n_loop > 0);
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
  ~JitModule() {
    if (module && driver().ok) driver().ModuleUnload(module);
  }
};

std::string jit_source(const Program& P, const std::vector<int>& steps) {
  std::string src;
  for (int si : steps) {
    Gen g(P, P.steps[si]);
    src += g.kernel("hb_step_" + std::to_string(si));
  }
  return src;
}

JitModule* jit_build(const Program& P, const std::vector<int>& steps, std::vector<void*>& fn_per_step, std::string* why) {
  auto fail = [&](const std::string& m) -> JitModule* {
    if (why) *why = m;
    if (verbose()) fprintf(stderr, "[hbayes jit] %s\n", m.c_str());
    return nullptr;
  };
  if (steps.empty()) return fail("no eligible step");
  const Nvrtc& rt = nvrtc();
  const Driver& drv = driver();
  if (!rt.ok) return fail("libnvrtc not available");
  if (!drv.ok) return fail("libcuda not available");
  const std::string src = jit_source(P, steps);
  nvrtcProgram prog = nullptr;
  if (rt.CreateProgram(&prog, src.c_str(), "hbayes_steps.cu", 0, nullptr, nullptr) != 0) return fail("nvrtcCreateProgram failed");
  const char* opts[] = {"--gpu-architecture=sm_100a", "--fmad=false", "-default-device", "--std=c++17", "--generate-line-info"};
  const int rc = rt.CompileProgram(prog, 5, opts);
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
  auto mod = std::make_unique<JitModule>();
  if (drv.ModuleLoadData(&mod->module, cubin.data()) != 0) return fail("cuModuleLoadData failed");
  fn_per_step.assign(P.steps.size(), nullptr);
  for (int si : steps) {
    void* fn = nullptr;
    const std::string name = "hb_step_" + std::to_string(si);
    if (drv.ModuleGetFunction(&fn, mod->module, name.c_str()) != 0) return fail("cuModuleGetFunction(" + name + ") failed");
    fn_per_step[si] = fn;
  }
  if (verbose()) fprintf(stderr, "[hbayes jit] %zu specialised kernels, %zu bytes of source, %zu bytes of cubin\n", steps.size(), src.size(), size);
  return mod.release();
}

void jit_destroy(JitModule* m) { delete m; }

bool jit_launch(void* fn, unsigned gx, unsigned gy, unsigned bx, unsigned by, void** params, void* stream) {
  return driver().LaunchKernel(fn, gx, gy, 1, bx, by, 1, 0, stream, params, nullptr) == 0;
}

}  // namespace hb
