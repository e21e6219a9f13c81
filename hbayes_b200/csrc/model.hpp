// model.hpp -- host-side structures of libhbayes_cuda: network, junction tree, compiled program.
//
// Everything here is integer/graph work done once per (network, observed-variable list)
// (SURVEY 8a rows a5,a6,a9-a15).  The floating-point work it schedules runs in engine.cu.
// Reference semantics are cited per function in the .cpp files (paths relative to the hbayes
// tree, e.g. Bayes/FactorElimination.hs:130-143).
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/hbayes_cuda.h"  // status codes HB_OK, HB_E_*

namespace hb {


struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// ---- Bayesian network: what runBN / importBayesianGraph yield (Bayes/Network.hs:69-79) ----------
struct Network {
  std::vector<int> dims;                    // vertex id = index
  std::vector<std::vector<int>> scope;      // scope[v] = [v, parents in `ingoing` order]
  std::vector<std::vector<double>> values;  // reference layout: first variable slowest (empty for a procedural CPT)
  std::vector<uint64_t> proc_seed;          // != 0: CPT v is procedural (procedural.hpp), values[v] is empty
  std::vector<std::string> names;
  int n() const { return (int)dims.size(); }
  bool procedural(int v) const { return !proc_seed.empty() && proc_seed[v] != 0; }
  int64_t cpt_entries(int v) const {  // factorDimension of CPT v
    int64_t p = 1;
    for (int x : scope[v]) p *= dims[x];
    return p;
  }
  double cpt_value(int v, int64_t idx) const;  // entry idx of CPT v, stored or procedural
  int64_t width() const {  // sum_v dims[v]: doubles per junction-tree query
    int64_t w = 0;
    for (int d : dims) w += d;
    return w;
  }
};
Network* network_create(int n_vars, const int32_t* dims, const int64_t* cpt_off, const int32_t* cpt_scope,
                        const int64_t* val_off, const double* vals, const uint64_t* proc_seed = nullptr);
Network* network_load_hugin(const char* path);

// ---- Junction tree structure (Bayes/FactorElimination.hs:283-307, JTree.hs:89-105) -------------
struct Tree {
  std::vector<int> elim_order;              // triangulation order (vertex ids)
  std::vector<std::vector<int>> clusters;   // maximal clusters, ascending vertex ids; index = cluster-graph vertex
  int root = 0;
  std::vector<int> parent_sep;              // cluster -> separator number (1-based) or -1 for the root
  std::vector<int> sep_parent, sep_child;   // separator (index sep-1) -> cluster
  std::vector<std::vector<int>> sep_vars;   // separator (index sep-1) -> ascending vertex ids
  std::vector<std::vector<int>> children;   // cluster -> separator numbers, newest first
  std::vector<std::vector<int>> factors;    // cluster -> CPT vertex ids in list order (last assigned first)
  std::vector<int> by_rank;                 // cluster ids in ascending `Cluster` (lexicographic) order
  int n_seps() const { return (int)sep_parent.size(); }
};
// cmp: 0 = nodeComparisonForTriangulation (weightedEdges), 1 = numberOfAddedEdges; `order` overrides cmp.
Tree* tree_build(const Network& net, int cmp, const int32_t* order);
// first minimal-state-space cluster covering `vars`, scanning clusters in ascending Cluster order
int tree_covering_cluster(const Network& net, const Tree& t, const std::vector<int>& vars);
// last top-most cluster (pre-order) holding all q vars (traverseTree findClusterFor), -1 if none
int tree_find_cluster(const Tree& t, const std::vector<int>& q);

std::vector<int> ve_elimination_order(const Network& net, int metric /*0 minDegree, 1 minFill*/);
int ve_degree_order(const Network& net, const std::vector<int>& order);

// ---- Compiled program: flat step list replayed per evidence row --------------------------------
//
// A *table* is a symbolic factor.  Its structural scope is the reference's variable list; its
// physical scope drops the observed variables: a table is only ever stored at the observed states
// of the row (evidence slicing -- value-identical to multiplying by the 0/1 indicator).
enum TableKind : int32_t {
  T_ONE = 0,     // the constant 1.0: an indicator read at its observed state, the product of an empty
                 // bucket, posterior's `Scalar 1.0`.  Never stored, never multiplied (x*1.0 is exact).
  T_POT = 1,     // a CPT of the network, in the potential pool (batch-invariant; sliced per row when it
                 // mentions an observed variable)
  T_SHARED = 2,  // derived table that does not depend on the evidence states: computed once per program
  T_ROW = 3,     // derived per-row table, lives in the row arena [entry][row]
};

struct Table {
  int32_t kind = T_ROW;
  bool scalar = false;           // structural `Scalar` (PrivateCPT.hs:67-72): empty structural scope
  int pot = -1;                  // vertex id for T_POT
  std::vector<int> scope;        // structural scope, reference order
  std::vector<int> pscope;       // physical scope = scope minus observed variables (same order)
  std::vector<int64_t> pstride;  // stride of each pscope variable in the stored layout
  int64_t size = 1;              // physical entries
  int64_t off = 0;               // entry offset in its pool (potential pool / shared arena / row arena)
  int slice_slot = -1;           // T_POT mentioning observed variables: index into Program::slices
  bool reduce = false;           // split clique over NCCL: the stored table is this GPU's partial sum over the split variables;
                                 // it is all-reduced (sum over the ranks) right after the step that produces it
  int alias_of = -1;             // a final product of a single table is that table
  int def_step = -1, last_use = -1;
};

struct StepInput {
  int table = -1;
  std::vector<int64_t> lstride;  // stride (entries) along the summed variable of level 0..own level (0 if absent)
  int64_t wstride = 0;           // stride along the step's blocking variable (0: the input does not depend on it)
  std::vector<uint32_t> hi, lo;  // entry offset of reduced output index o': hi[o' / C] + lo[o' % C]
  // One-entry tables every loaded entry is multiplied by, in this order: x = ((t[i] * s0) * s1) * ...  (compile.cpp "scale
  // steps": a step that only multiplies a table by the restricted prior of an observed / fixed variable is evaluated where
  // its result is read instead of being stored).  At most one input of a step carries scales, at most MAX_SCALES of them.
  std::vector<int> scales;
};
constexpr int MAX_SCALES = 3;

// One reference operation (itemProduct of a bucket + itemProjectOut of one variable).  A step holds a
// chain of them, level 0 outermost: the table a level produces is consumed only by the level above it,
// as the FIRST factor of that level's product, so it never has to exist in memory:
//   value_l = 0 + sum_{s_l} value_{l+1} * (in_1 * (in_2 * ...))      (value_{L} absent for the innermost level)
struct Level {
  int sum_var = -1;  // eliminated vertex, -1 for a final product
  int d = 1;         // physical terms (1 when nothing is summed or the variable is observed)
  bool is_max = false;  // MAXCPT (Bayes/Factor/MaxCPT.hs:38-47): keep the LAST maximal term and record its state
  std::vector<StepInput> in;  // stored inputs in reference chain order (after the fused inner value)
};

enum StepKind : int32_t { STEP_PRODSUM = 0, STEP_OUTPUT = 1 };

struct Step {
  int32_t kind = STEP_PRODSUM;
  bool shared = false;          // all inputs batch-invariant: runs once at upload time, not per row
  std::vector<Level> levels;    // STEP_PRODSUM: >= 1 level
  std::vector<int> scalars;     // structural scalars of level 0, ps = s1 * (s2 * ...), applied as ps * chain (never fused)
  int out = -1;                 // output table
  int64_t n_out = 1, C = 1, n_hi = 1;
  // Register blocking: one output variable w (the fastest one of the output's stored layout) is walked by
  // the thread itself, two states at a time; an input of the innermost level that does not mention w is
  // loaded once for both.  n_red = n_out / U entries of the other output variables go through hi/lo.
  int w_var = -1;
  int U = 1;
  int64_t n_red = 1;
  std::vector<int> full_scope;  // STEP_OUTPUT: the result's structural scope
  // STEP_OUTPUT: normalise table `result` and expand it to the full query scope
  int result = -1;
  int64_t out_col = 0, n_full = 1;
  std::vector<int32_t> full_map;               // full entry -> physical entry of `result` (ignoring observed vars)
  std::vector<int> obs_var;                    // observed query variables ...
  std::vector<int64_t> obs_stride, obs_dim;    // ... their stride and dimension in the full layout
  bool gather = false;                         // STEP_OUTPUT, split clique over NCCL: the query holds a split variable -- every rank fills
                                               // its own slice of the (unnormalised) result, the ranks' matrices are summed, then normalised
  int group = -1;                              // provenance: index into Program::groups (first emitter)
  int64_t arg_off = -1;                        // max step: entry offset of its argmax table in the argmax arena
  int n_inputs() const {
    int n = 0;
    for (const Level& l : levels) n += (int)l.in.size();
    return n;
  }
};

struct Group {  // one reference-level `marginal` call: a message or a posterior
  int kind;     // 0 up, 1 down, 2 posterior, 3 variable elimination
  int cluster = -1, sep = -1;
  std::vector<int> query;
  std::vector<int> out_scope;  // result variable order (rule A.5 / A.10)
  int result = -1;             // table id
  // the reference-level trace (what the oracle's interpreter does for the same evidence list)
  struct TraceStep {
    int var;  // eliminated vertex, -1 = final product
    std::vector<std::vector<int>> inputs;
    std::vector<int> prod, out;
  };
  std::vector<TraceStep> trace;
};

struct Slice {  // per-row entry offset of a sliced potential: sum_j stride_j * state(var_j)
  int pot;
  std::vector<int> vars;              // observed (per-row) variables of the CPT
  std::vector<int64_t> strides;       // their strides in the stored layout
};

struct Program {
  const Network* net = nullptr;
  std::vector<int> observed;     // in the caller's order (changeEvidence list order)
  std::vector<char> is_observed; // per vertex (fixed variables included)
  std::vector<int> fixed_state;  // per vertex: compile-time state or -1
  std::vector<Table> tables;
  std::vector<Step> steps;       // shared steps and per-row steps interleaved in dependency order
  std::vector<Group> groups;
  std::vector<Slice> slices;      // potentials that are sliced per row
  std::vector<int64_t> pot_off;  // vertex -> entry offset in the potential pool
  std::vector<double> extra_values;  // free-standing factors (hb_factor_*), appended to the pool at extra_off
  int64_t extra_off = 0;
  int64_t pot_entries = 0;
  int64_t shared_entries = 0;    // shared arena size
  int64_t row_entries = 0;       // per-row arena size
  int64_t row_entries_no_reuse = 0;  // what it would be without reusing freed space
  int64_t out_width = 0;         // doubles written per row
  // statistics (per row, physical)
  int64_t stat_prod = 0, stat_reads = 0, stat_row_read = 0, stat_row_write = 0, stat_max_table = 0;
  int64_t stat_ops = 0;          // reference-level product/sum-out operations behind the per-row steps
  // ---- mpe (Bayes/VariableElimination.hs:79-114): maximisation steps in elimination order ----
  struct MaxStep {
    int var = -1;               // maximised vertex
    int slot = -1;              // its position in mpe_r
    int64_t arg_off = 0;        // argmax table: int8 state per entry of the step's output, same stored layout
    std::vector<int> dep_slot;  // the output's variables (all maximised later) as positions in mpe_r ...
    std::vector<int64_t> dep_stride;  // ... and their strides in the stored layout
  };
  std::vector<MaxStep> max_steps;
  std::vector<int> mpe_r;        // variables to maximise, caller's order
  std::vector<int> mpe_order;    // variable order of the reference's instantiation list (structural)
  int64_t arg_entries = 0;       // per-row argmax arena size (int8 entries)
  int n_reduce = 0, n_gather = 0;  // tables all-reduced / queries gathered across the ranks of a split clique (0: no communication)
};

struct Query {
  std::vector<int> vars;
};
struct CompileOptions {
  // variables with a state known at compile time (the slice of a split clique this GPU owns): treated like
  // observed variables, and potentials mentioning them are stored restricted to that state
  std::vector<std::pair<int, int>> fixed;
  // Split clique with the sum over the split variables taken where it belongs: every message / posterior
  // that eliminates the split variables is computed from this GPU's slice and all-reduced over the ranks (|result| doubles
  // per row); what follows is no longer conditioned on the slice.  false: the whole program stays conditioned and the
  // caller sums the result matrices (cutset conditioning, hb_jt_set_split).
  bool reduce_groups = false;
  bool fuse = true;    // chain a step into its only consumer (the intermediate table stays in registers)
  bool block = true;   // choose the blocking variable of each step by the loads it saves (false: the last variable)
  int max_levels = 4;  // levels per fused step
  int max_outer = 4;   // stored inputs of the outer levels of a fused step
  int max_inner = 3;   // stored inputs of its innermost level
};
Program* compile_jt(const Network& net, const Tree& tree, const std::vector<int>& observed,
                    const std::vector<Query>& queries, const CompileOptions& opt = CompileOptions());
Program* compile_ve(const Network& net, const std::vector<int>& p, const std::vector<int>& r,
                    const std::vector<int>& observed, const CompileOptions& opt = CompileOptions());
// mpe g p r assignment (VariableElimination.hs:79-114): sum out p, maximise over r.  The program's single output is
// the UNNORMALISED scalar P(mpe, e); the maximising states come from the argmax tables (Program::max_steps).
Program* compile_mpe(const Network& net, const std::vector<int>& p, const std::vector<int>& r,
                     const std::vector<int>& observed, const CompileOptions& opt = CompileOptions());

// One factor-algebra operation on free-standing factors: factorProduct of `factors` and, when sum_var >= 0,
// factorProjectOut [sum_var] of the product (= marginalizeOneVariable, Buckets.hs:105-111).  `dims` gives the
// dimension of every vertex id that occurs.  The program has a single group whose out_scope is the result scope.
struct FreeFactor {
  std::vector<int> scope;
  std::vector<double> values;
};
struct FactorOp {
  std::unique_ptr<Network> net;  // carrier of the vertex dimensions
  std::unique_ptr<Program> prog;
};
FactorOp* compile_factor_op(const std::vector<int>& dims, const std::vector<FreeFactor>& factors, int sum_var);

std::string describe_network(const Network& net);
std::string describe_tree(const Tree& t);
std::string describe_program(const Program& p);

}  // namespace hb
