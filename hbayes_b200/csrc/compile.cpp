// compile.cpp -- the schedule compiler: runs the reference's bucket elimination SYMBOLICALLY, once per
// (network, observed-variable list), and emits a flat list of fused product/sum-out steps with index
// maps.  The GPU replays that list for every evidence row (engine.cu).
//
// What is reproduced from the reference, operation for operation:
//   marginal / createBuckets / addBucket / updateBucket / marginalizeOneVariable
//       (Bayes/VariableElimination.hs:56-73, Bayes/VariableElimination/Buckets.hs:47-111)
//   product scope = foldr1 union, value = ps * (v1 * (v2 * ...)), scalars folded into ps
//       (Bayes/Factor/PrivateCPT.hs:215-238)
//   sum-out keeps the remaining variables in their order, adds terms in ascending state from 0
//       (Bayes/Factor/PrivateCPT.hs:243-266, Bayes/Factor/CPT.hs:140-141)
//   newMessage input order f ++ e ++ input, collect / distribute (Bayes/FactorElimination/JTree.hs:288-369,485-491)
//   changeEvidence: indicators in the caller's order, each prepended to its cluster (JTree.hs:424-466)
//   posterior input order d : u ++ f ++ e (Bayes/FactorElimination.hs:314-327)
//
// What is different (and why it is value-identical): an observed variable is never
// enumerated.  Its indicator is the constant 1.0 at the observed state and every table is only
// stored at that state ("evidence slicing"): x * 1.0 == x and 0 + x + 0 == x exactly, and entries at
// the other states are exactly 0 in the reference and never reach a result.  Requires finite table
// values (0 * inf would be NaN in the reference).
#include <algorithm>
#include <map>
#include <memory>
#include <sstream>

#include "model.hpp"

namespace hb {
namespace {

bool contains(const std::vector<int>& xs, int x) { return std::find(xs.begin(), xs.end(), x) != xs.end(); }

// Data.List.union on duplicate-free lists (Bayes/PrivateTypes.hs:118-128)
std::vector<int> list_union(const std::vector<int>& xs, const std::vector<int>& ys) {
  std::vector<int> out = xs;
  for (int y : ys)
    if (!contains(out, y)) out.push_back(y);
  return out;
}

struct Compiler {
  const Network& net;
  Program& P;
  int one = -1;  // the `Scalar 1.0` table
  bool reduce_groups = false;
  // split clique: the split (fixed) variables a table is still conditioned on -- it holds the term of THIS GPU's states
  std::map<int, std::vector<int>> dep;
  std::vector<int> dep_of(int id) const {
    auto it = dep.find(id);
    return it == dep.end() ? std::vector<int>{} : it->second;
  }

  Compiler(const Network& n, Program& p, const std::vector<int>& observed, const CompileOptions& opt) : net(n), P(p) {
    reduce_groups = opt.reduce_groups;
    P.net = &net;
    P.observed = observed;
    P.is_observed.assign(net.n(), 0);
    P.fixed_state.assign(net.n(), -1);
    for (auto& f : opt.fixed) {
      if (f.first < 0 || f.first >= net.n() || f.second < 0 || f.second >= net.dims[f.first]) throw Error(HB_E_ARG, "bad fixed (split) variable or state");
      P.fixed_state[f.first] = f.second;
    }
    for (int v : observed) {
      if (v < 0 || v >= net.n()) throw Error(HB_E_ARG, "evidence on unknown vertex " + std::to_string(v));
      if (P.is_observed[v]) throw Error(HB_E_ARG, "vertex " + std::to_string(v) + " observed twice");
      P.is_observed[v] = 1;
    }
    for (int v = 0; v < net.n(); ++v)
      if (P.fixed_state[v] >= 0 && !P.is_observed[v]) throw Error(HB_E_ARG, "a fixed (split) variable must be listed as observed");
    // potential pool: a CPT is stored whole, except that variables with a compile-time state are restricted away
    P.pot_off.assign(net.n(), 0);
    for (int v = 0; v < net.n(); ++v) {
      int64_t entries = 1;
      for (int x : net.scope[v])
        if (P.fixed_state[x] < 0) entries *= net.dims[x];
      P.pot_off[v] = P.pot_entries, P.pot_entries += entries;
    }
    Table t;
    t.kind = T_ONE;
    t.scalar = true;
    one = add(t);
  }

  int add(const Table& t) {
    P.tables.push_back(t);
    return (int)P.tables.size() - 1;
  }
  const Table& T(int id) const { return P.tables[id]; }
  int resolve(int id) const {
    while (P.tables[id].alias_of >= 0) id = P.tables[id].alias_of;
    return id;
  }

  void set_physical(Table& t) {  // row-major over the unobserved variables of the structural scope
    t.pscope.clear();
    for (int v : t.scope)
      if (!P.is_observed[v]) t.pscope.push_back(v);
    t.pstride.assign(t.pscope.size(), 1);
    t.size = 1;
    for (int i = (int)t.pscope.size() - 1; i >= 0; --i) t.pstride[i] = t.size, t.size *= net.dims[t.pscope[i]];
  }

  // CPT of vertex v as it sits in the potential pool (Bayes/Network.hs:69-79): strides of the FULL table,
  // observed variables selected per row through a slice offset.
  int potential(int v) {
    Table t;
    t.kind = T_POT;
    t.pot = v;
    t.scope = net.scope[v];
    t.off = P.pot_off[v];
    const int ns = (int)t.scope.size();
    std::vector<int64_t> stored(ns, 0);
    int64_t acc = 1;  // stored layout: the reference layout with the fixed variables removed
    for (int i = ns - 1; i >= 0; --i)
      if (P.fixed_state[t.scope[i]] < 0) stored[i] = acc, acc *= net.dims[t.scope[i]];
    Slice sl;
    sl.pot = v;
    t.size = 1;
    for (int i = 0; i < ns; ++i) {
      const int x = t.scope[i];
      if (P.fixed_state[x] >= 0)
        continue;  // restricted away when the pool is built (engine.cu)
      else if (P.is_observed[x])
        sl.vars.push_back(x), sl.strides.push_back(stored[i]);
      else
        t.pscope.push_back(x), t.pstride.push_back(stored[i]), t.size *= net.dims[x];
    }
    if (!sl.vars.empty()) {
      t.slice_slot = (int)P.slices.size();
      P.slices.push_back(sl);
    }
    const int id = add(t);
    for (int x : P.tables[id].scope)
      if (P.fixed_state[x] >= 0) dep[id].push_back(x);
    return id;
  }

  // a free-standing factor (hb_factor_*): values appended to the potential pool, reference layout
  int free_factor(const std::vector<int>& scope, const std::vector<double>& values) {
    Table t;
    t.kind = T_POT;
    t.scope = scope;
    t.scalar = scope.empty();  // createCPTWithDims [] = Scalar
    set_physical(t);
    if ((int64_t)values.size() != t.size) throw Error(HB_E_ARG, "factor value count does not match its scope");
    t.off = P.pot_entries;
    if (P.extra_values.empty()) P.extra_off = P.pot_entries;
    P.extra_values.insert(P.extra_values.end(), values.begin(), values.end());
    P.pot_entries += t.size;
    return add(t);
  }

  // factorFromInstantiation (Bayes/Factor.hs:90-94) of an observed variable: 1.0 at the row's state.
  int indicator(int v) {
    Table t;
    t.kind = T_ONE;
    t.scope = {v};
    return add(t);
  }

  bool batch_invariant(int id) const {
    const Table& t = T(resolve(id));
    return t.kind == T_ONE || t.kind == T_SHARED || (t.kind == T_POT && t.slice_slot < 0);
  }

  // One reference operation: itemProduct of `inputs` followed (sum_var >= 0) by itemProjectOut sum_var.
  // MAXCPT bookkeeping: table id -> variables of the instantiation every entry carries, in list order.  The order
  // is structural (multiply concatenates, maximization prepends: MaxCPT.hs:29-47), so it is known at compile time.
  std::map<int, std::vector<int>> inst;
  const std::vector<int>& inst_of(int id) {
    static const std::vector<int> none;
    auto it = inst.find(id);
    return it == inst.end() ? none : it->second;
  }

  int product_sumout(const std::vector<int>& inputs, int sum_var, Group& g, bool is_max = false) {
    const int result = product_sumout_traced(inputs, sum_var, g, is_max);
    if (is_max || !inst.empty()) {  // op ps (v1 `op` (v2 ...)): the scalars' instantiations first, then the tables'
      std::vector<int> vars;
      for (int pass = 0; pass < 2; ++pass)
        for (int id : inputs)
          if ((T(id).scalar ? 0 : 1) == pass)
            for (int v : inst_of(id)) vars.push_back(v);
      bool any_table = false, mentions = false;
      for (int id : inputs)
        if (!T(id).scalar) any_table = true, mentions = mentions || contains(T(id).scope, sum_var);
      if (is_max && any_table && mentions && sum_var >= 0) vars.insert(vars.begin(), sum_var);  // addTo newInst possible
      if (!vars.empty() && result != one) inst[result] = vars;
    }
    return result;
  }

  int product_sumout_traced(const std::vector<int>& inputs, int sum_var, Group& g, bool is_max) {
    Group::TraceStep tr;
    tr.var = sum_var;
    for (int id : inputs) tr.inputs.push_back(T(id).scope);
    if (inputs.empty()) {  // _factorProduct [] = Scalar 1.0; projectOut of a Scalar is the Scalar
      g.trace.push_back(tr);
      return one;
    }
    // Common subexpressions: the same operation on the same tables (e.g. the first eliminations of two
    // posteriors of one cluster) is the same sequence of floating-point operations; emit it once.
    std::vector<int> key;
    for (int id : inputs) {  // an alias stands for its target when both have the same structural scope
      const int r = resolve(id);
      key.push_back(T(id).scope == T(r).scope ? r : id);
    }
    key.push_back(-2 - sum_var);
    if (is_max) key.push_back(-1);
    std::vector<int> conditioned;  // union of the inputs' split variables
    for (int id : inputs)
      for (int x : dep_of(id))
        if (!contains(conditioned, x)) conditioned.push_back(x);
    auto memo = cse.find(key);
    if (memo != cse.end()) {
      tr.prod = memo->second.prod;
      tr.out = T(memo->second.table).scope;
      g.trace.push_back(tr);
      return memo->second.table;
    }
    const int result = product_sumout_new(inputs, sum_var, tr, is_max, reduce_groups && !conditioned.empty());
    g.trace.push_back(tr);
    cse[key] = Memo{result, tr.prod};
    if (!conditioned.empty() && result != one) dep[result] = conditioned;
    return result;
  }

  struct Memo {
    int table;
    std::vector<int> prod;
  };
  std::map<std::vector<int>, Memo> cse;

  int product_sumout_new(const std::vector<int>& inputs, int sum_var, Group::TraceStep& tr, bool is_max = false, bool per_row = false) {
    std::vector<int> naked = T(inputs.back()).scope;
    for (int i = (int)inputs.size() - 2; i >= 0; --i) naked = list_union(T(inputs[i]).scope, naked);
    std::vector<int> chain, scalars;  // structural partition isScalarFactor (PrivateCPT.hs:223)
    for (int id : inputs) (T(id).scalar ? scalars : chain).push_back(id);
    Table out;
    out.scalar = true;
    if (!chain.empty()) {
      for (int v : naked)
        if (v != sum_var) out.scope.push_back(v);
      out.scalar = out.scope.empty();  // createCPTWithDims [] = Scalar
    }
    tr.prod = chain.empty() ? std::vector<int>{} : naked;
    tr.out = out.scope;

    // ---- physical lowering ----
    Step st;
    st.group = (int)P.groups.size();
    bool shared = true;
    for (int id : scalars)
      if (T(resolve(id)).kind != T_ONE) st.scalars.push_back(resolve(id)), shared = shared && batch_invariant(id);
    std::vector<int> real;
    for (int id : chain)
      if (T(resolve(id)).kind != T_ONE) real.push_back(resolve(id)), shared = shared && batch_invariant(id);
    if (real.empty() && st.scalars.empty()) {  // only ones: the constant 1.0
      out.kind = T_ONE;
      return add(out);
    }
    const bool sums = sum_var >= 0 && !P.is_observed[sum_var] && contains(naked, sum_var) && !chain.empty();
    if (!sums && st.scalars.empty() && real.size() == 1) {  // 1.0 * v (and 0 + v): the table itself
      out.alias_of = real[0];
      out.kind = T(real[0]).kind;
      set_physical(out);
      return add(out);
    }
    if (real.empty() && st.scalars.size() == 1) {
      out.alias_of = st.scalars[0];
      out.kind = T(st.scalars[0]).kind;
      return add(out);
    }
    if (is_max && sums) shared = false;  // the argmax tables are per row
    if (per_row) shared = false;         // conditioned on this GPU's slice and possibly all-reduced per batch: never hoisted
    out.kind = shared ? T_SHARED : T_ROW;
    set_physical(out);
    st.shared = shared;
    Level lv;
    lv.sum_var = sums ? sum_var : -1;
    lv.d = sums ? net.dims[sum_var] : 1;
    lv.is_max = is_max && sums;
    for (int id : real) {
      StepInput in;
      in.table = id;
      lv.in.push_back(in);
    }
    st.levels.push_back(std::move(lv));
    st.n_out = out.size;
    int id = add(out);
    st.out = id;
    P.steps.push_back(std::move(st));
    return id;
  }

  // marginal (VariableElimination.hs:56-73) over symbolic factors; `assignment` = indicator tables.
  // mpe = true: mpemarginal (VariableElimination.hs:79-97) -- after the sums over p the buckets become MAXCPTs and
  // the variables of r are maximised out in turn; the final product is then a scalar.
  int marginal(const std::vector<int>& lf, const std::vector<int>& p, const std::vector<int>& r,
               const std::vector<int>& assignment, Group& g, bool mpe = false) {
    std::vector<int> order = p;
    order.insert(order.end(), r.begin(), r.end());
    std::map<int, std::vector<int>> buckets;  // Data.Map DV [f]; DV order = vertex id
    std::vector<int> rest = lf;
    for (int dv : order) {  // createBuckets (Buckets.hs:47-65)
      std::vector<int> fk, others;
      for (int id : rest) (contains(T(id).scope, dv) ? fk : others).push_back(id);
      rest.swap(others);
      buckets[dv] = fk;
    }
    size_t head = 0;  // order[head..] = variables still to process
    auto add_bucket = [&](int id) {  // addBucket (Buckets.hs:92-98): prepend
      for (size_t i = head; i < order.size(); ++i)
        if (contains(T(id).scope, order[i])) {
          auto& b = buckets[order[i]];
          b.insert(b.begin(), id);
          return;
        }
    };
    for (int id : assignment) add_bucket(id);
    for (int dv : p) {  // marginalizeOneVariable + updateBucket (Buckets.hs:74-89,105-111)
      std::vector<int> fk = buckets.at(dv);
      int f2 = product_sumout(fk, dv, g);
      if (head < order.size()) ++head;
      if (T(f2).scalar)
        buckets[dv] = {f2};
      else {
        buckets.erase(dv);
        add_bucket(f2);
      }
    }
    if (mpe)
      for (int dv : r) {
        std::vector<int> fk = buckets.at(dv);
        int f2 = product_sumout(fk, dv, g, true);
        if (head < order.size()) ++head;
        if (T(f2).scalar)
          buckets[dv] = {f2};
        else {
          buckets.erase(dv);
          add_bucket(f2);
        }
      }
    std::vector<int> fin;
    for (auto& kv : buckets) fin.insert(fin.end(), kv.second.begin(), kv.second.end());
    return product_sumout(fin, -1, g, mpe);
  }

  // Split clique, end of a message / posterior: the split variables the result is conditioned on but no longer mentions
  // were eliminated inside this group -- from this GPU's slice only.  The result is linear in that partial sum, so the
  // sum over the ranks is taken here, on |result| doubles per row.  Returns true when the result still mentions split
  // variables (it stays one slice per GPU).
  // Returns the table that now stands for the group's result; *sliced = the result still mentions split variables (it
  // stays one slice per GPU).  The all-reduce runs in place on a private copy of the result (|result| entries per row):
  // the result itself may be shared, through common-subexpression elimination, with groups that still use it as this
  // GPU's partial sum.
  int close_group(int res, bool is_query, bool* sliced) {
    *sliced = false;
    if (!reduce_groups) return res;
    const std::vector<int> d = dep_of(res);
    if (d.empty()) return res;
    std::vector<int> gone, kept;
    for (int x : d) (contains(T(res).scope, x) ? kept : gone).push_back(x);
    *sliced = !kept.empty();
    if (gone.empty()) return res;
    // a query that keeps some split variables: every rank writes its term at its own states of the kept ones, zero
    // elsewhere, and the sum of the expanded results over ALL ranks is also the sum over the eliminated ones
    if (is_query && !kept.empty()) return res;
    if (!kept.empty())
      throw Error(HB_E_UNSUPPORTED, "split clique: a separator holds some of the split variables and eliminates others (needs a sub-communicator)");
    const int rid = resolve(res);
    if (P.tables[rid].kind == T_ONE) return res;
    Table out;  // the copy: 0 + 1.0 * x, exact
    out.scalar = T(res).scalar;
    out.scope = T(res).scope;
    out.kind = T_ROW;
    set_physical(out);
    Step st;
    st.group = (int)P.groups.size();
    st.shared = false;
    Level lv;
    StepInput in;
    in.table = rid;
    lv.in.push_back(in);
    st.levels.push_back(std::move(lv));
    st.n_out = out.size;
    out.reduce = true;
    const int id = add(out);
    st.out = id;
    P.steps.push_back(std::move(st));
    return id;
  }

  std::vector<int> first_appearance(const std::vector<int>& factors) const {  // nub . concatMap factorVariables
    std::vector<int> seen;
    for (int id : factors)
      for (int v : T(id).scope)
        if (!contains(seen, v)) seen.push_back(v);
    return seen;
  }

  // Normalisation + expansion of a query result (factorNorm / factorDivide, Factor.hs:171-172,
  // PrivateCPT.hs:167-187): norm = left fold over the layout order, every entry times 1.0 / norm.
  void emit_output(int result, Group& g) {
    Step st;
    st.kind = STEP_OUTPUT;
    st.group = (int)P.groups.size();
    const int rid = resolve(result);
    const Table& t = T(result);
    st.result = rid;
    g.out_scope = t.scope;
    g.result = result;
    const int nq = (int)t.scope.size();
    std::vector<int64_t> fstride(nq, 1);
    st.n_full = 1;
    for (int i = nq - 1; i >= 0; --i) fstride[i] = st.n_full, st.n_full *= net.dims[t.scope[i]];
    st.full_scope = t.scope;  // full_map is built in finish(), once the stored layouts are final
    const Table& stored = T(rid);
    for (int i = 0; i < nq; ++i)
      if (P.is_observed[t.scope[i]]) {
        st.obs_var.push_back(t.scope[i]);
        st.obs_stride.push_back(fstride[i]);
        st.obs_dim.push_back(net.dims[t.scope[i]]);
      }
    st.n_out = stored.size;
    st.out_col = P.out_width;
    P.out_width += st.n_full;
    P.steps.push_back(std::move(st));
  }

  template <class F>
  static void for_each_input(const Step& st, F&& f) {
    if (st.kind == STEP_OUTPUT) {
      f(st.result);
      return;
    }
    for (const Level& l : st.levels)
      for (const StepInput& in : l.in) {
        f(in.table);
        for (int id : in.scales) f(id);
      }
    for (int id : st.scalars) f(id);
  }

  // After the symbolic run: drop dead steps (messages no requested query needs), fuse producer/consumer
  // chains, build the index maps, place the per-row tables in the row arena by liveness.
  void finish(const CompileOptions& opt) {
    // ---- 1. dead steps ----
    {
      const int ns = (int)P.steps.size();
      std::vector<char> live_t(P.tables.size(), 0), live_s(ns, 0);
      for (int s = ns - 1; s >= 0; --s) {
        const Step& st = P.steps[s];
        if (st.kind == STEP_OUTPUT || live_t[st.out]) {
          live_s[s] = 1;
          for_each_input(st, [&](int id) { live_t[id] = 1; });
        }
      }
      std::vector<Step> kept;
      for (int s = 0; s < ns; ++s)
        if (live_s[s]) kept.push_back(std::move(P.steps[s]));
      P.steps.swap(kept);
    }
    for (const Step& st : P.steps)
      if (st.kind == STEP_PRODSUM && !st.shared) ++P.stat_ops;
    // ---- 1b. scale steps ----
    // A per-row step that only multiplies a table by a ONE-ENTRY table and sums nothing -- T * c, c the prior of an observed
    // or fixed root restricted to its state -- is not materialised for the product steps that read it: they read T and
    // multiply every loaded entry by c themselves.  Same operands, same single rounding (IEEE multiplication commutes), so
    // the values are those the stored table would have held; what disappears is one write and one read of the whole table
    // per reader (C4: three 2^26-entry copies per row, 40 % of the pass' traffic).
    if (opt.fuse) {
      auto unit = [&](int id) {
        const Table& t = P.tables[id];
        return t.size == 1 && (t.kind == T_POT || t.kind == T_SHARED) && !t.scalar;
      };
      for (size_t q = 0; q < P.steps.size(); ++q) {
        const Step& Q = P.steps[q];
        if (Q.kind != STEP_PRODSUM || Q.shared || !Q.scalars.empty() || Q.levels.size() != 1) continue;
        const Level& lq = Q.levels[0];
        if (lq.is_max || lq.sum_var >= 0 || lq.d != 1 || lq.in.size() != 2) continue;
        const int big_at = unit(lq.in[1].table) ? 0 : (unit(lq.in[0].table) ? 1 : -1);
        if (big_at < 0) continue;
        const StepInput& big = lq.in[(size_t)big_at];
        const int c = lq.in[(size_t)(1 - big_at)].table;
        if (!lq.in[(size_t)(1 - big_at)].scales.empty() || (int)big.scales.size() >= MAX_SCALES) continue;
        const Table& T = P.tables[big.table];
        std::vector<int> a = T.pscope, b = P.tables[Q.out].pscope;
        std::sort(a.begin(), a.end());
        std::sort(b.begin(), b.end());
        if (a != b) continue;  // the product must be entry for entry over T's scope
        std::vector<int> scales = big.scales;
        scales.push_back(c);
        for (size_t r = q + 1; r < P.steps.size(); ++r) {
          Step& S = P.steps[r];
          if (S.kind != STEP_PRODSUM || S.levels.empty() || S.levels[0].is_max) continue;
          bool other_scaled = false, reads = false;
          for (const Level& l : S.levels)
            for (const StepInput& in : l.in) {
              if (in.table == Q.out) reads = true;
              else if (!in.scales.empty()) other_scaled = true;
            }
          if (!reads || other_scaled) continue;
          int hits = 0;
          for (const Level& l : S.levels)
            for (const StepInput& in : l.in) hits += in.table == Q.out;
          if (hits != 1) continue;
          for (Level& l : S.levels)
            for (StepInput& in : l.in)
              if (in.table == Q.out) in.table = big.table, in.scales = scales;
        }
      }
      // the scale steps nobody reads any more
      const int ns = (int)P.steps.size();
      std::vector<char> live_t(P.tables.size(), 0), live_s(ns, 0);
      for (int s2 = ns - 1; s2 >= 0; --s2) {
        const Step& st = P.steps[s2];
        if (st.kind == STEP_OUTPUT || live_t[st.out]) {
          live_s[s2] = 1;
          for_each_input(st, [&](int id) { live_t[id] = 1; });
        }
      }
      std::vector<Step> kept;
      for (int s2 = 0; s2 < ns; ++s2)
        if (live_s[s2]) kept.push_back(std::move(P.steps[s2]));
      P.steps.swap(kept);
    }
    // ---- 2. fusion ----
    // Consumer S whose FIRST chain input T is produced by step Q, read by nobody else, and whose scope is
    // exactly S's product scope (every entry of T is used once): Q's levels are appended below S's and T
    // disappears.  The inner value is complete before it is multiplied, so the FP sequence is unchanged.
    if (opt.fuse) {
      std::vector<int> uses(P.tables.size(), 0), producer(P.tables.size(), -1);
      for (const Step& st : P.steps) for_each_input(st, [&](int id) { ++uses[id]; });
      for (int s = 0; s < (int)P.steps.size(); ++s)
        if (P.steps[s].kind == STEP_PRODSUM) producer[P.steps[s].out] = s;
      std::vector<char> dead(P.steps.size(), 0);
      for (int s = 0; s < (int)P.steps.size(); ++s) {
        Step& S = P.steps[s];
        if (S.kind != STEP_PRODSUM || !S.scalars.empty() || S.levels[0].in.empty() || S.levels[0].is_max) continue;
        const int tid = S.levels[0].in[0].table;
        const int q = producer[tid];
        if (q < 0 || uses[tid] != 1 || P.tables[tid].reduce) continue;
        const Step& Q = P.steps[q];
        if (Q.shared != S.shared || !Q.scalars.empty() || Q.levels[0].is_max) continue;
        {  // the table read through scale factors must exist; and a step keeps at most one scaled input
          if (!S.levels[0].in[0].scales.empty()) continue;
          int scaled = 0;
          const Step* both[2] = {&S, &Q};
          for (const Step* x : both)
            for (const Level& l : x->levels)
              for (const StepInput& in : l.in) scaled += !in.scales.empty();
          if (scaled > 1) continue;
        }
        const int q_inner = (int)Q.levels.back().in.size();
        if ((int)(S.levels.size() + Q.levels.size()) > opt.max_levels || q_inner < 1 || q_inner > opt.max_inner ||
            S.n_inputs() - 1 + Q.n_inputs() - q_inner > opt.max_outer)
          continue;
        std::vector<int> want = P.tables[S.out].pscope;  // S's physical product scope
        if (S.levels[0].sum_var >= 0) want.push_back(S.levels[0].sum_var);
        std::vector<int> have = P.tables[tid].pscope;
        std::sort(want.begin(), want.end());
        std::sort(have.begin(), have.end());
        if (want != have) continue;
        S.levels[0].in.erase(S.levels[0].in.begin());
        for (const Level& l : Q.levels) S.levels.push_back(l);
        dead[q] = 1;
        producer[tid] = -1;
      }
      std::vector<Step> kept;
      for (size_t s = 0; s < P.steps.size(); ++s)
        if (!dead[s]) kept.push_back(std::move(P.steps[s]));
      P.steps.swap(kept);
    }
    // ---- 3. stored layouts, index maps and statistics ----
    for (Step& st : P.steps) {
      if (st.kind == STEP_OUTPUT) {  // full entry -> stored entry of the result (layouts are final by now)
        const Table& stored = P.tables[st.result];
        const int nq = (int)st.full_scope.size();
        st.full_map.assign((size_t)st.n_full, 0);
        std::vector<int> digit(nq, 0);
        for (int64_t e = 0; e < st.n_full; ++e) {
          int64_t off = 0;
          for (int i = 0; i < nq; ++i)
            for (size_t j = 0; j < stored.pscope.size(); ++j)
              if (stored.pscope[j] == st.full_scope[i]) off += stored.pstride[j] * digit[i];
          st.full_map[e] = (int32_t)off;
          for (int i = nq - 1; i >= 0; --i) {
            if (++digit[i] < net.dims[st.full_scope[i]]) break;
            digit[i] = 0;
          }
        }
        continue;
      }
      Table& out = P.tables[st.out];
      const int nk = (int)out.pscope.size();
      auto stride_in = [&](const Table& t, int v) -> int64_t {
        for (size_t i = 0; i < t.pscope.size(); ++i)
          if (t.pscope[i] == v) return t.pstride[i];
        return 0;
      };
      // blocking variable: the output variable whose absence from the innermost-level inputs saves the most
      // loads when a thread walks two of its states together; the output is stored with that variable fastest
      int wpos = -1;
      if (opt.block) {
        const Level& inner = st.levels.back();
        double terms = 1;
        for (const Level& l : st.levels) terms *= l.d;
        double best = -1;
        for (int j = 0; j < nk; ++j) {
          const int U = net.dims[out.pscope[j]];
          const double keep = (double)((U + 1) / 2) / U;  // loads left for an input that does not depend on it
          double saved = 0;
          if ((int)inner.in.size() <= 3)
            for (const StepInput& in : inner.in)
              if (stride_in(P.tables[in.table], out.pscope[j]) == 0) saved += terms * (1.0 - keep);
          if (saved >= best) best = saved, wpos = j;  // ties: the last variable (the natural row-major layout)
        }
      } else if (nk > 0)
        wpos = nk - 1;
      st.w_var = wpos >= 0 ? out.pscope[wpos] : -1;
      st.U = wpos >= 0 ? net.dims[st.w_var] : 1;
      st.n_red = st.n_out / st.U;
      std::vector<int> red;  // the other output variables, in scope order
      for (int j = 0; j < nk; ++j)
        if (j != wpos) red.push_back(out.pscope[j]);
      {  // stored layout of the output: row-major over `red`, then w fastest
        int64_t acc = st.U;
        std::vector<int64_t> ps(nk, 1);
        for (int j = nk - 1; j >= 0; --j)
          if (j != wpos) ps[j] = acc, acc *= net.dims[out.pscope[j]];
        out.pstride = ps;
      }
      const int nr = (int)red.size();
      int split = nr;
      int64_t C = 1;
      while (split > 0 && C * C < st.n_red && C * net.dims[red[split - 1]] <= 4096) C *= net.dims[red[--split]];
      st.C = C;
      st.n_hi = st.n_red / C;
      int64_t terms = 1;
      for (size_t l = 0; l < st.levels.size(); ++l) {
        terms *= st.levels[l].d;
        for (StepInput& in : st.levels[l].in) {
          const Table& t = P.tables[in.table];
          in.lstride.clear();
          for (size_t m = 0; m <= l; ++m) in.lstride.push_back(st.levels[m].sum_var >= 0 ? stride_in(t, st.levels[m].sum_var) : 0);
          in.wstride = st.w_var >= 0 ? stride_in(t, st.w_var) : 0;
          auto fill = [&](std::vector<uint32_t>& dst, size_t count, int from, int to) {
            dst.assign(count, 0);
            std::vector<int> digit(nr, 0);
            int64_t off = 0;
            for (size_t e = 0; e < count; ++e) {
              dst[e] = (uint32_t)off;
              for (int j = to - 1; j >= from; --j) {
                const int64_t sj = stride_in(t, red[j]);
                if (++digit[j] < net.dims[red[j]]) {
                  off += sj;
                  break;
                }
                off -= sj * (digit[j] - 1);
                digit[j] = 0;
              }
            }
          };
          fill(in.hi, (size_t)st.n_hi, 0, split);
          fill(in.lo, (size_t)C, split, nr);
          if (!st.shared) {
            P.stat_reads += st.n_out * terms;
            if (t.kind == T_ROW) P.stat_row_read += std::min<int64_t>(t.size, st.n_out * terms);
          }
        }
        if (!st.shared) P.stat_prod += st.n_out * terms;
      }
      if (!st.shared) {
        P.stat_row_write += st.n_out;
        P.stat_max_table = std::max(P.stat_max_table, st.n_out * terms);
      }
    }
    // ---- 3b. argmax tables of the maximisation steps (mpe): one int8 per output entry, never reused ----
    for (Step& st : P.steps) {
      if (st.kind != STEP_PRODSUM || !st.levels[0].is_max) continue;
      const Table& out = P.tables[st.out];
      Program::MaxStep ms;
      ms.var = st.levels[0].sum_var;
      ms.arg_off = st.arg_off = P.arg_entries;
      P.arg_entries += out.size;
      for (size_t j = 0; j < out.pscope.size(); ++j) ms.dep_slot.push_back(out.pscope[j]), ms.dep_stride.push_back(out.pstride[j]);  // vertex ids; compile_mpe turns them into slots
      P.max_steps.push_back(std::move(ms));
    }
    // ---- 4. liveness and arena placement ----
    for (auto& t : P.tables) t.def_step = t.last_use = -1;
    const int last = (int)P.steps.size() - 1;
    for (int s = 0; s <= last; ++s) {
      const Step& st = P.steps[s];
      if (st.kind == STEP_PRODSUM) P.tables[st.out].def_step = s;
      // query results are read by ONE normalise/expand kernel after the last step: they live to the end
      for_each_input(st, [&](int id) { P.tables[id].last_use = st.kind == STEP_OUTPUT ? last : std::max(P.tables[id].last_use, s); });
    }
    // shared arena: no reuse (computed once, read by every row)
    for (auto& t : P.tables)
      if (t.kind == T_SHARED && t.alias_of < 0 && t.def_step >= 0) t.off = P.shared_entries, P.shared_entries += t.size;
    // row arena.  When giving every per-row table its own place costs at most twice the memory of reusing freed
    // space, do that: reuse orders otherwise independent steps (write-after-read) and the engine overlaps
    // independent steps inside its CUDA graph.  Otherwise (long elimination chains over huge tables): first-fit
    // over a free list, tables released after their last reader.
    {
      int64_t no_reuse = 0;
      for (const Step& st : P.steps)
        if (st.kind == STEP_PRODSUM && !st.shared) no_reuse += P.tables[st.out].size;
      P.row_entries_no_reuse = no_reuse;
    }
    std::vector<std::pair<int64_t, int64_t>> free_list;  // (offset, size), sorted by offset
    int64_t top = 0;
    std::vector<std::vector<int>> dies_at(P.steps.size());
    for (size_t id = 0; id < P.tables.size(); ++id) {
      const Table& t = P.tables[id];
      if (t.kind == T_ROW && t.alias_of < 0 && t.def_step >= 0 && t.last_use >= 0) dies_at[t.last_use].push_back((int)id);
    }
    auto release = [&](int64_t off, int64_t size) {
      auto it = std::lower_bound(free_list.begin(), free_list.end(), std::make_pair(off, (int64_t)0));
      it = free_list.insert(it, {off, size});
      if (it + 1 != free_list.end() && it->first + it->second == (it + 1)->first) it->second += (it + 1)->second, free_list.erase(it + 1);
      if (it != free_list.begin() && (it - 1)->first + (it - 1)->second == it->first) (it - 1)->second += it->second, it = free_list.erase(it), --it;
      if (it->first + it->second == top) top = it->first, free_list.erase(it);
    };
    for (int s = 0; s <= last; ++s) {
      Step& st = P.steps[s];
      if (st.kind == STEP_PRODSUM && !st.shared) {
        Table& t = P.tables[st.out];
        bool placed = false;
        for (size_t i = 0; i < free_list.size(); ++i)
          if (free_list[i].second >= t.size) {
            t.off = free_list[i].first;
            free_list[i].first += t.size, free_list[i].second -= t.size;
            if (free_list[i].second == 0) free_list.erase(free_list.begin() + i);
            placed = true;
            break;
          }
        if (!placed) t.off = top, top += t.size;
        P.row_entries = std::max(P.row_entries, top);
      }
      for (int id : dies_at[s]) release(P.tables[id].off, P.tables[id].size);
    }
    if (P.row_entries_no_reuse <= 2 * P.row_entries) {
      int64_t at = 0;
      for (const Step& st : P.steps)
        if (st.kind == STEP_PRODSUM && !st.shared) P.tables[st.out].off = at, at += P.tables[st.out].size;
      P.row_entries = at;
    }
  }
};

void js_list(std::ostringstream& o, const std::vector<int>& v) {
  o << "[";
  for (size_t i = 0; i < v.size(); ++i) o << (i ? "," : "") << v[i];
  o << "]";
}
template <class V>
void js_nums(std::ostringstream& o, const V& v) {
  o << "[";
  for (size_t i = 0; i < v.size(); ++i) o << (i ? "," : "") << (long long)v[i];
  o << "]";
}

}  // namespace

// createJunctionTree's calibration / changeEvidence (JTree.hs:454-466) followed by `posterior q` for
// every requested query, as one program.
Program* compile_jt(const Network& net, const Tree& tree, const std::vector<int>& observed,
                    const std::vector<Query>& queries, const CompileOptions& opt) {
  auto prog = std::make_unique<Program>();
  Compiler C(net, *prog, observed, opt);
  const int nc = (int)tree.clusters.size(), nsep = tree.n_seps();
  std::vector<int> pot(net.n());
  for (int v = 0; v < net.n(); ++v) pot[v] = C.potential(v);
  std::vector<std::vector<int>> f(nc), e(nc);
  for (int c = 0; c < nc; ++c)
    for (int v : tree.factors[c]) f[c].push_back(pot[v]);
  // updateTreeValues changeEvidence (JTree.hs:424-440,459-460): caller's order, each PREPENDED
  for (int v : observed) {
    int c = tree_covering_cluster(net, tree, {v});
    e[c].insert(e[c].begin(), C.indicator(v));
  }
  std::vector<int> up(nsep + 1, -1), down(nsep + 1, -1);
  auto message = [&](int kind, int cluster, int sep, const std::vector<int>& incoming) {  // newMessage (JTree.hs:485-491)
    Group g;
    g.kind = kind;
    g.cluster = cluster;
    g.sep = sep;
    std::vector<int> all = f[cluster];
    all.insert(all.end(), e[cluster].begin(), e[cluster].end());
    all.insert(all.end(), incoming.begin(), incoming.end());
    const std::vector<int>& keep = tree.sep_vars[sep - 1];
    std::vector<int> remove;
    for (int v : C.first_appearance(all))
      if (!contains(keep, v)) remove.push_back(v);
    int res = C.marginal(all, remove, keep, {}, g);
    bool sliced = false;
    res = C.close_group(res, false, &sliced);
    g.out_scope = C.T(res).scope;
    g.result = res;
    prog->groups.push_back(std::move(g));
    return res;
  };
  // collect (JTree.hs:329-348): level-synchronous sweep from the leaves; an up message is a pure function
  // of the subtree below it, so it is emitted once, the first time the sweep finds its inputs ready.
  {
    std::vector<int> level;
    for (int c : tree.by_rank)
      if (tree.children[c].empty()) level.push_back(c);
    std::vector<int> rank(nc);
    for (int i = 0; i < nc; ++i) rank[tree.by_rank[i]] = i;
    while (!level.empty()) {
      for (int h : level) {
        int ps = tree.parent_sep[h];
        if (ps < 0 || up[ps] >= 0) continue;
        bool ready = true;
        std::vector<int> inc;
        for (int s : tree.children[h]) {
          if (up[s] < 0) ready = false;
          inc.push_back(up[s]);
        }
        if (ready) up[ps] = message(0, h, ps, inc);
      }
      std::vector<int> next;
      for (int h : level)
        if (tree.parent_sep[h] >= 0) next.push_back(tree.sep_parent[tree.parent_sep[h] - 1]);
      std::sort(next.begin(), next.end(), [&](int a, int b) { return rank[a] < rank[b]; });
      next.erase(std::unique(next.begin(), next.end()), next.end());
      level.swap(next);
    }
  }
  // distribute (JTree.hs:350-369): all down messages of a node (children newest first), then recurse.
  {
    std::vector<int> stack = {tree.root};
    while (!stack.empty()) {
      int node = stack.back();
      stack.pop_back();
      for (int child : tree.children[node]) {
        std::vector<int> inc;  // [down from parent] ++ [up of the other children]  (JTree.hs:307-322)
        if (tree.parent_sep[node] >= 0) inc.push_back(down[tree.parent_sep[node]]);
        for (int s : tree.children[node])
          if (s != child) inc.push_back(up[s]);
        down[child] = message(1, node, child, inc);
      }
      for (auto it = tree.children[node].rbegin(); it != tree.children[node].rend(); ++it) stack.push_back(tree.sep_child[*it - 1]);
    }
  }
  // posterior (FactorElimination.hs:314-327)
  for (const Query& q : queries) {
    if (q.vars.empty()) throw Error(HB_E_ARG, "empty posterior query");
    for (size_t i = 0; i < q.vars.size(); ++i) {
      if (q.vars[i] < 0 || q.vars[i] >= net.n()) throw Error(HB_E_ARG, "query on unknown vertex");
      for (size_t j = 0; j < i; ++j)
        if (q.vars[i] == q.vars[j]) throw Error(HB_E_ARG, "query repeats a vertex");
    }
    int c = tree_find_cluster(tree, q.vars);
    if (c < 0) throw Error(HB_E_NO_CLUSTER, "no cluster holds all query variables");
    Group g;
    g.kind = 2;
    g.cluster = c;
    g.query = q.vars;
    std::vector<int> all;
    all.push_back(tree.parent_sep[c] >= 0 ? down[tree.parent_sep[c]] : C.one);
    for (int s : tree.children[c]) all.push_back(up[s]);
    all.insert(all.end(), f[c].begin(), f[c].end());
    all.insert(all.end(), e[c].begin(), e[c].end());
    std::vector<int> remove;
    for (int v : C.first_appearance(all))
      if (!contains(q.vars, v)) remove.push_back(v);
    int res = C.marginal(all, remove, q.vars, {}, g);
    bool sliced = false;
    res = C.close_group(res, true, &sliced);
    C.emit_output(res, g);
    prog->steps.back().gather = sliced;
    prog->groups.push_back(std::move(g));
  }
  C.finish(opt);
  for (const Step& st : prog->steps) {
    if (st.kind == STEP_PRODSUM && prog->tables[st.out].reduce) ++prog->n_reduce;
    if (st.kind == STEP_OUTPUT && st.gather) ++prog->n_gather;
  }
  return prog.release();
}

// posteriorMarginal g p r assignment (VariableElimination.hs:116-130): all CPTs in ascending vertex id,
// indicators added with addBucket in the caller's order.
Program* compile_ve(const Network& net, const std::vector<int>& p, const std::vector<int>& r,
                    const std::vector<int>& observed, const CompileOptions& opt) {
  auto prog = std::make_unique<Program>();
  Compiler C(net, *prog, observed, opt);
  for (int v : p)
    if (v < 0 || v >= net.n()) throw Error(HB_E_ARG, "elimination order mentions an unknown vertex");
  for (int v : r)
    if (v < 0 || v >= net.n()) throw Error(HB_E_ARG, "query mentions an unknown vertex");
  std::vector<int> lf, asg;
  for (int v = 0; v < net.n(); ++v) lf.push_back(C.potential(v));
  for (int v : observed) asg.push_back(C.indicator(v));
  Group g;
  g.kind = 3;
  g.query = r;
  int res = C.marginal(lf, p, r, asg, g);
  C.emit_output(res, g);
  prog->groups.push_back(std::move(g));
  C.finish(opt);
  return prog.release();
}

// mpe g p r assignment (VariableElimination.hs:101-114) -> mpemarginal (:79-97).
Program* compile_mpe(const Network& net, const std::vector<int>& p, const std::vector<int>& r,
                     const std::vector<int>& observed, const CompileOptions& opt) {
  auto prog = std::make_unique<Program>();
  Compiler C(net, *prog, observed, opt);
  for (int v : p)
    if (v < 0 || v >= net.n()) throw Error(HB_E_ARG, "elimination order mentions an unknown vertex");
  for (size_t i = 0; i < r.size(); ++i) {
    const int v = r[i];
    if (v < 0 || v >= net.n()) throw Error(HB_E_ARG, "maximisation order mentions an unknown vertex");
    if (contains(p, v) || std::find(r.begin(), r.begin() + i, v) != r.begin() + i) throw Error(HB_E_ARG, "a variable occurs twice in the elimination orders");
    // the reference multiplies by the indicator and then maximises (a tie among zeros picks the last state); the
    // engine never enumerates an observed variable, so evidence belongs to the summed list, as the reference's
    // documentation asks ("should contain evidence variables", VariableElimination.hs:104)
    if (prog->is_observed[v]) throw Error(HB_E_UNSUPPORTED, "mpe: an observed variable must be in the summed list p, not in r");
  }
  std::vector<int> lf, asg;
  for (int v = 0; v < net.n(); ++v) lf.push_back(C.potential(v));
  for (int v : observed) asg.push_back(C.indicator(v));
  Group g;
  g.kind = 3;
  g.query = r;
  int res = C.marginal(lf, p, r, asg, g, true);
  if (!C.T(res).scalar) throw Error(HB_E_ARG, "mpe: p and r must cover every variable (the final MAXCPT factor should be a scalar one)");
  prog->mpe_r = r;
  prog->mpe_order = C.inst_of(res);
  C.emit_output(res, g);
  prog->groups.push_back(std::move(g));
  C.finish(opt);
  for (Program::MaxStep& ms : prog->max_steps) {
    auto slot_of = [&](int v) {
      auto it = std::find(r.begin(), r.end(), v);
      if (it == r.end()) throw Error(HB_E_ARG, "internal: a maximised table depends on a variable outside r");
      return (int)(it - r.begin());
    };
    ms.slot = slot_of(ms.var);
    for (int& d : ms.dep_slot) d = slot_of(d);
  }
  return prog.release();
}

FactorOp* compile_factor_op(const std::vector<int>& dims, const std::vector<FreeFactor>& factors, int sum_var) {
  auto op = std::make_unique<FactorOp>();
  op->net = std::make_unique<Network>();
  Network& net = *op->net;  // every vertex gets a trivial uniform prior: only the dimensions matter here
  net.dims = dims;
  for (int v = 0; v < (int)dims.size(); ++v) {
    if (dims[v] < 1 || dims[v] > 127) throw Error(HB_E_ARG, "vertex dimension must be in 1..127");
    net.scope.push_back({v});
    net.values.push_back(std::vector<double>((size_t)dims[v], 1.0 / dims[v]));
    net.names.push_back("n" + std::to_string(v));
  }
  if (sum_var >= (int)dims.size()) throw Error(HB_E_ARG, "unknown variable to sum out");
  op->prog = std::make_unique<Program>();
  CompileOptions opt;
  opt.fuse = false;
  Compiler C(net, *op->prog, {}, opt);
  std::vector<int> ids;
  for (const FreeFactor& f : factors) {
    for (size_t i = 0; i < f.scope.size(); ++i) {
      if (f.scope[i] < 0 || f.scope[i] >= (int)dims.size()) throw Error(HB_E_ARG, "factor mentions an unknown vertex");
      for (size_t j = 0; j < i; ++j)
        if (f.scope[i] == f.scope[j]) throw Error(HB_E_ARG, "factor repeats a vertex");
    }
    ids.push_back(C.free_factor(f.scope, f.values));
  }
  Group g;
  g.kind = 3;
  int res = C.product_sumout(ids, -1, g);                 // factorProduct
  if (sum_var >= 0) res = C.product_sumout({res}, sum_var, g);  // factorProjectOut [sum_var] (1.0 * p, then the sum)
  C.emit_output(res, g);
  op->prog->groups.push_back(std::move(g));
  C.finish(opt);
  return op.release();
}

// Parity hook (hb_jt_describe_json / hb_ve_describe_json): the reference-level operation trace of every
// message and query, plus the physical step list with its index maps.
std::string describe_program(const Program& P) {
  std::ostringstream o;
  o << "{\"observed\":";
  js_list(o, P.observed);
  o << ",\"row_entries\":" << P.row_entries << ",\"shared_entries\":" << P.shared_entries << ",\"out_width\":" << P.out_width
    << ",\"stats\":{\"prod\":" << P.stat_prod << ",\"reads\":" << P.stat_reads << ",\"row_read\":" << P.stat_row_read
    << ",\"row_write\":" << P.stat_row_write << ",\"max_table\":" << P.stat_max_table << ",\"steps\":" << P.steps.size() << ",\"ops\":" << P.stat_ops << "}";
  o << ",\"groups\":[";
  const char* kinds[] = {"up", "down", "posterior", "ve"};
  for (size_t i = 0; i < P.groups.size(); ++i) {
    const Group& g = P.groups[i];
    o << (i ? "," : "") << "{\"kind\":\"" << kinds[g.kind] << "\",\"cluster\":" << g.cluster << ",\"sep\":" << g.sep << ",\"query\":";
    js_list(o, g.query);
    o << ",\"out_scope\":";
    js_list(o, g.out_scope);
    o << ",\"steps\":[";
    for (size_t j = 0; j < g.trace.size(); ++j) {
      const auto& t = g.trace[j];
      o << (j ? "," : "") << "{\"kind\":\"" << (t.var >= 0 ? "elim" : "final") << "\",\"var\":" << t.var << ",\"inputs\":[";
      for (size_t k = 0; k < t.inputs.size(); ++k) {
        o << (k ? "," : "");
        js_list(o, t.inputs[k]);
      }
      o << "],\"prod\":";
      js_list(o, t.prod);
      o << ",\"out\":";
      js_list(o, t.out);
      o << "}";
    }
    o << "]}";
  }
  o << "],\"physical\":[";
  for (size_t i = 0; i < P.steps.size(); ++i) {
    const Step& s = P.steps[i];
    o << (i ? "," : "") << "{\"kind\":" << s.kind << ",\"shared\":" << (s.shared ? 1 : 0) << ",\"group\":" << s.group << ",\"n_out\":" << s.n_out
      << ",\"C\":" << s.C << ",\"out\":" << s.out << ",\"w_var\":" << s.w_var << ",\"U\":" << s.U;
    if (s.kind == STEP_PRODSUM) {
      o << ",\"out_scope\":";
      js_list(o, P.tables[s.out].pscope);
      o << ",\"out_off\":" << P.tables[s.out].off << ",\"reduce\":" << (P.tables[s.out].reduce ? 1 : 0) << ",\"scalars\":";
      js_list(o, s.scalars);
      o << ",\"levels\":[";
      for (size_t l = 0; l < s.levels.size(); ++l) {
        const Level& lv = s.levels[l];
        o << (l ? "," : "") << "{\"sum_var\":" << lv.sum_var << ",\"d\":" << lv.d << ",\"is_max\":" << (lv.is_max ? 1 : 0) << ",\"in\":[";
        for (size_t k = 0; k < lv.in.size(); ++k) {
          const Table& t = P.tables[lv.in[k].table];
          o << (k ? "," : "") << "{\"table\":" << lv.in[k].table << ",\"kind\":" << t.kind << ",\"pot\":" << t.pot << ",\"off\":" << t.off
            << ",\"size\":" << t.size << ",\"lstride\":";
          js_nums(o, lv.in[k].lstride);
          o << ",\"wstride\":" << lv.in[k].wstride << ",\"scales\":";
          js_list(o, lv.in[k].scales);
          o << ",\"scope\":";
          js_list(o, t.pscope);
          o << ",\"strides\":";
          js_nums(o, t.pstride);
          o << "}";
        }
        o << "]}";
      }
      o << "]";
    } else
      o << ",\"result\":" << s.result << ",\"out_col\":" << s.out_col << ",\"n_full\":" << s.n_full << ",\"gather\":" << (s.gather ? 1 : 0);
    o << "}";
  }
  o << "],\"mpe_order\":";
  js_list(o, P.mpe_order);
  o << ",\"arg_entries\":" << P.arg_entries << ",\"max_steps\":[";
  for (size_t i = 0; i < P.max_steps.size(); ++i) {
    const Program::MaxStep& ms = P.max_steps[i];
    o << (i ? "," : "") << "{\"var\":" << ms.var << ",\"slot\":" << ms.slot << ",\"arg_off\":" << ms.arg_off << ",\"dep_slot\":";
    js_list(o, ms.dep_slot);
    o << ",\"dep_stride\":";
    js_nums(o, ms.dep_stride);
    o << "}";
  }
  o << "]}";
  return o.str();
}

}  // namespace hb
