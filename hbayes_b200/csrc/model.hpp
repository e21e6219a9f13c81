// model.hpp -- host-side structures of libhbayes_cuda: network, junction tree, compiled program.
//
// Everything here is integer/graph work done once per network (SURVEY 8a rows a5,a6,a9-a15); the
// floating-point work it schedules runs in engine.cu.  Reference semantics are cited per function
// in the .cpp files (paths relative to the hbayes tree, e.g. Bayes/FactorElimination.hs:130-143).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace hb {

struct Error {
  int code;
  std::string msg;
};

// ---- Bayesian network: what runBN / importBayesianGraph yield (Bayes/Network.hs:69-79) ----------
struct Network {
  std::vector<int> dims;                    // vertex id = index
  std::vector<std::vector<int>> scope;      // scope[v] = [v, parents in `ingoing` order]
  std::vector<std::vector<double>> values;  // reference layout: first variable slowest
  std::vector<std::string> names;
  int n() const { return (int)dims.size(); }
  int64_t width() const {  // sum_v dims[v]: doubles per JT query
    int64_t w = 0;
    for (int d : dims) w += d;
    return w;
  }
};
Network* network_create(int n_vars, const int32_t* dims, const int64_t* cpt_off, const int32_t* cpt_scope,
                        const int64_t* val_off, const double* vals);
Network* network_load_hugin(const char* path);

// ---- Junction tree structure (Bayes/FactorElimination.hs:283-307, JTree.hs:89-105) -------------
struct Tree {
  std::vector<int> elim_order;              // triangulation order (vertex ids)
  std::vector<std::vector<int>> clusters;   // maximal clusters, ascending vertex ids; index = cluster-graph vertex
  int root = 0;
  std::vector<int> parent_sep;              // cluster -> separator number (1-based) or -1 for the root
  std::vector<int> sep_parent, sep_child;   // separator (index sep-1) -> cluster
  std::vector<std::vector<int>> sep_vars;   // separator (index sep-1) -> ascending vertex ids
  std::vector<std::vector<int>> children;   // cluster -> separators, newest first
  std::vector<std::vector<int>> factors;    // cluster -> CPT vertex ids in list order (last assigned first)
  std::vector<int> by_rank;                 // cluster ids in ascending `Cluster` (lexicographic) order
};
// cmp: 0 = nodeComparisonForTriangulation (weightedEdges), 1 = numberOfAddedEdges; order overrides cmp.
Tree* tree_build(const Network& net, int cmp, const int32_t* order);
int tree_covering_cluster(const Network& net, const Tree& t, const std::vector<int>& vars);
// top-most cluster holding all q vars (traverseTree findClusterFor), -1 if none
int tree_find_cluster(const Tree& t, const std::vector<int>& q);

std::vector<int> ve_elimination_order(const Network& net, int metric /*0 minDegree, 1 minFill*/);

// ---- Compiled program: flat step list replayed per evidence row --------------------------------
enum Loc : int32_t { LOC_ARENA = 0, LOC_POT = 1, LOC_ONE = 2 };
enum StepKind : int32_t { STEP_PRODSUM = 0, STEP_OUTPUT = 1 };

struct Table {           // a symbolic factor
  bool scalar = false;   // structural Scalar (PrivateCPT.hs:67-72)
  bool is_one = false;   // physically the constant 1.0 (evidence indicator at its observed state, empty bucket)
  bool is_evidence = false;
  int pot = -1;          // vertex id when this is a CPT (LOC_POT) or an evidence indicator
  std::vector<int> scope;   // structural scope, reference order
  std::vector<int> pscope;  // physical scope: scope minus sliced (observed) variables
  std::vector<int64_t> pstride;  // stride of each pscope variable in the physical layout
  int64_t size = 1;         // physical entries
  int32_t loc = LOC_ARENA;
  int64_t off = 0;          // arena entry offset / offset into the potential pool
  int alias_of = -1;
  int def_step = -1, last_use = -1;
};

struct StepInput {
  int table;
  int64_t sstride;                 // stride of the summed variable in this input (0 if absent / sliced)
  std::vector<int32_t> hi, lo;     // off(o) = hi[o / C] + lo[o % C]
};

struct Step {
  int32_t kind = STEP_PRODSUM;
  int sum_var = -1;                // eliminated vertex, -1 for a final product
  int d = 1;                       // physical terms per output (1 if no sum / summed variable observed)
  std::vector<StepInput> in;       // physical table inputs, chain order v1*(v2*(...*vk))
  std::vector<int> scalars;        // structural scalars s1*(s2*...), multiplied as ps * chain
  int out = -1;                    // output table
  int64_t n_out = 1, C = 1, n_hi = 1;
  // STEP_OUTPUT: normalise `in[0]` and expand to the full query scope
  int64_t out_col = 0, n_full = 1;
  std::vector<int32_t> full_map;               // full entry -> physical entry
  std::vector<int> obs_slot;                   // observed query variables: evidence slot,
  std::vector<int64_t> obs_stride, obs_dim;    //   stride and dim in the full layout
  // provenance, for hb_*_describe_json (the parity hook)
  int group = -1;
  std::vector<std::vector<int>> trace_inputs;  // structural scopes of ALL reference inputs (incl. indicators)
  std::vector<int> trace_prod, trace_out;
};

struct Group {  // one reference-level `marginal` call: a message or a posterior
  int kind;     // 0 up, 1 down, 2 posterior, 3 ve
  int cluster = -1, sep = -1;
  std::vector<int> query;
  std::vector<int> out_scope;  // result variable order (rule A.5 / A.10)
  int first_step = 0, n_steps = 0;
};

struct Program {
  const Network* net = nullptr;
  std::vector<int> observed;        // ordered as given by the caller (changeEvidence list order)
  std::vector<int> slot_of_var;     // vertex -> evidence slot or -1
  std::vector<Table> tables;
  std::vector<Step> steps;
  std::vector<Group> groups;
  int64_t arena_entries = 0;        // per-row arena size
  int64_t out_width = 0;            // doubles written per row
  // sliced potentials: evoff[slot][row] = sum stride * state
  std::vector<int> sliced_pots;                        // vertex ids with >=1 observed variable in scope
  std::vector<int> pot_slice_slot;                     // vertex -> index into sliced_pots or -1
  std::vector<std::vector<std::pair<int, int64_t>>> pot_slice_terms;  // per sliced pot: (evidence slot, stride)
  std::vector<int64_t> pot_off;                        // vertex -> offset into the potential pool
  int64_t pot_pool = 0;
  int64_t max_table = 0, total_prod = 0, total_reads = 0, total_out = 0;
};

struct Query {
  std::vector<int> vars;
};
Program* compile_jt(const Network& net, const Tree& tree, const std::vector<int>& observed,
                    const std::vector<Query>& queries);
Program* compile_ve(const Network& net, const std::vector<int>& p, const std::vector<int>& r,
                    const std::vector<int>& observed);

std::string describe_network(const Network& net);
std::string describe_tree(const Tree& t);
std::string describe_program(const Program& p);

}  // namespace hb
