// hbayes_oracle.cpp -- CPU restatement of the hbayes hot path.  TEST INFRASTRUCTURE ONLY.
//
// This file is the parity ORACLE (and the "port" CPU baseline).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load the
// library built from it (oracle/liboracle.so).  Nothing under hbayes_b200/ links or calls it.
//
// It is an *interpreter*: like the reference it re-runs bucket elimination with concrete
// factor objects for every evidence set (the product, by contrast, compiles a step list once
// and replays it on the GPU).  Arithmetic is IEEE-754 binary64 `*` and `+` in exactly the
// reference's association order; build with -ffp-contract=off (see oracle/Makefile).
//
// Reference files followed (under /root/reference/), cited again at each function:
//   Bayes/PrivateTypes.hs, Bayes/Factor/PrivateCPT.hs, Bayes/Factor/CPT.hs, Bayes/Factor.hs,
//   Bayes/VariableElimination/Buckets.hs, Bayes/VariableElimination.hs,
//   Bayes/FactorElimination/JTree.hs, Bayes/FactorElimination.hs, Bayes.hs,
//   Bayes/ImportExport/HuginNet.hs
//
// Parity pin: checked in tests/test_oracle.py against the reference's golden vectors
// (Bayes/Test/ReferencePatterns.hs:119-161), its VE==JT test (Test/CompareEliminations.hs:67-84),
// the independent pure-Python twin (oracle/twin.py, bit-compared) and brute-force enumeration.
// GHC is not available in this image: digits beyond the reference's 4-decimal goldens are pinned
// by the restatement only.

#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

namespace {

// DV = DV !Vertex !Int, derived Ord (PrivateTypes.hs:202): vertex id, then dimension.
struct DV {
  int v, d;
  bool operator==(const DV& o) const { return v == o.v && d == o.d; }
  bool operator!=(const DV& o) const { return !(*this == o); }
  bool operator<(const DV& o) const { return v != o.v ? v < o.v : d < o.d; }
};
using Vars = std::vector<DV>;

bool has(const Vars& xs, const DV& x) { return std::find(xs.begin(), xs.end(), x) != xs.end(); }

// Data.List.union for duplicate-free lists: xs ++ [y | y <- ys, y notElem xs]  (PrivateTypes.hs:118-128)
Vars l_union(const Vars& xs, const Vars& ys) {
  Vars out = xs;
  for (const DV& y : ys)
    if (!has(out, y)) out.push_back(y);
  return out;
}
// Data.List.(\\)
Vars l_diff(const Vars& xs, const Vars& ys) {
  Vars out = xs;
  for (const DV& y : ys) {
    auto it = std::find(out.begin(), out.end(), y);
    if (it != out.end()) out.erase(it);
  }
  return out;
}

// PrivateCPT (PrivateCPT.hs:67-72): Table dims values | Scalar value.
struct Factor {
  bool scalar = false;
  Vars vars;
  std::vector<double> vals;
  bool contains(const DV& dv) const {  // _containsVariable: by vertex id (PrivateCPT.hs:156-158)
    if (scalar) return false;
    for (const DV& x : vars)
      if (x.v == dv.v) return true;
    return false;
  }
};
using FP = std::shared_ptr<const Factor>;

FP mk_scalar(double v) {
  auto f = std::make_shared<Factor>();
  f->scalar = true;
  f->vals = {v};
  return f;
}
// createCPTWithDims (PrivateCPT.hs:270-281)
FP mk_table(const Vars& vars, std::vector<double>&& vals) {
  if (vars.empty()) return mk_scalar(vals.at(0));
  auto f = std::make_shared<Factor>();
  f->vars = vars;
  f->vals = std::move(vals);
  return f;
}

// indexStrides (PrivateCPT.hs:315-320): first variable slowest.
std::vector<int64_t> index_strides(const Vars& d) {
  std::vector<int64_t> s(d.size(), 1);
  for (int i = (int)d.size() - 2; i >= 0; --i) s[i] = s[i + 1] * d[i + 1].d;
  return s;
}
// indexStridesFor (PrivateCPT.hs:302-310): read `dr` with a multi-index over `di`; absent => 0.
std::vector<int64_t> index_strides_for(const Vars& dr, const Vars& di) {
  std::vector<int64_t> own = index_strides(dr), out(di.size(), 0);
  for (size_t j = 0; j < di.size(); ++j)
    for (size_t i = 0; i < dr.size(); ++i)
      if (dr[i] == di[j]) {
        out[j] = own[i];
        break;
      }
  return out;
}
int64_t table_size(const Vars& d) {
  int64_t p = 1;
  for (const DV& x : d) p *= x.d;
  return p;
}

// _factorProduct (Op 1.0 (*)) (PrivateCPT.hs:215-238, CPT.hs:138).
// scope = foldr1 union; value = ps * (v1 * (v2 * (... * vk))); enumeration lexicographic
// (PrivateTypes.hs:256-260), realised here as an odometer carrying one offset per input.
FP factor_product(const std::vector<FP>& l) {
  if (l.empty()) return mk_scalar(1.0);
  Vars naked = l.back()->vars;
  for (int i = (int)l.size() - 2; i >= 0; --i) naked = l_union(l[i]->vars, naked);
  std::vector<const Factor*> cpts;
  std::vector<double> sc;
  for (const FP& f : l) (f->scalar ? (void)sc.push_back(f->vals[0]) : (void)cpts.push_back(f.get()));
  double ps = 1.0;
  if (!sc.empty()) {
    ps = sc.back();
    for (int i = (int)sc.size() - 2; i >= 0; --i) ps = sc[i] * ps;
  }
  if (cpts.empty()) return mk_scalar(ps);
  const int k = (int)cpts.size(), n = (int)naked.size();
  std::vector<std::vector<int64_t>> st(k);
  for (int i = 0; i < k; ++i) st[i] = index_strides_for(cpts[i]->vars, naked);
  const int64_t total = table_size(naked);
  std::vector<double> out((size_t)total);
  std::vector<int> idx(n, 0);
  std::vector<int64_t> off(k, 0);
  for (int64_t e = 0; e < total; ++e) {
    double acc = cpts[k - 1]->vals[off[k - 1]];
    for (int i = k - 2; i >= 0; --i) acc = cpts[i]->vals[off[i]] * acc;
    out[e] = ps * acc;
    for (int j = n - 1; j >= 0; --j) {  // odometer, last variable fastest
      if (++idx[j] < naked[j].d) {
        for (int i = 0; i < k; ++i) off[i] += st[i][j];
        break;
      }
      idx[j] = 0;
      for (int i = 0; i < k; ++i) off[i] -= st[i][j] * (naked[j].d - 1);
    }
  }
  return mk_table(naked, std::move(out));
}

// factorProjectOut (CPT.hs:140-141) = cptFactorProjectOutWith (sum . map fst) (PrivateCPT.hs:243-266).
// kept scope = d \\ s in original order; terms added left to right from 0 in ascending state.
FP factor_project_out(const Vars& s, const FP& f) {
  if (f->scalar) return f;
  const Vars& d = f->vars;
  Vars keep = l_diff(d, s);
  Vars ks = keep;
  ks.insert(ks.end(), s.begin(), s.end());
  std::vector<int64_t> st = index_strides_for(d, ks);
  const int nk = (int)keep.size(), ns = (int)s.size();
  const int64_t nkeep = table_size(keep), nsum = table_size(s);
  std::vector<int64_t> soff((size_t)nsum);
  {
    std::vector<int> idx(ns, 0);
    for (int64_t e = 0; e < nsum; ++e) {
      int64_t o = 0;
      for (int j = 0; j < ns; ++j) o += st[nk + j] * idx[j];
      soff[e] = o;
      for (int j = ns - 1; j >= 0; --j) {
        if (++idx[j] < s[j].d) break;
        idx[j] = 0;
      }
    }
  }
  std::vector<double> out((size_t)nkeep);
  std::vector<int> idx(nk, 0);
  int64_t base = 0;
  for (int64_t e = 0; e < nkeep; ++e) {
    double acc = 0.0;
    for (int64_t t = 0; t < nsum; ++t) acc = acc + f->vals[base + soff[t]];
    out[e] = acc;
    for (int j = nk - 1; j >= 0; --j) {
      if (++idx[j] < keep[j].d) {
        base += st[j];
        break;
      }
      idx[j] = 0;
      base -= st[j] * (keep[j].d - 1);
    }
  }
  return mk_table(keep, std::move(out));
}

// _factorNorm (PrivateCPT.hs:167-173)
double factor_norm(const FP& f) {
  if (f->scalar) return f->vals[0];
  double acc = 0.0;
  for (double v : f->vals) acc = acc + v;
  return acc;
}
// _factorScale (PrivateCPT.hs:180-187) with scale = (*) (CPT.hs:112)
FP factor_scale(double s, const FP& f) {
  if (f->scalar) return mk_scalar(s * f->vals[0]);
  std::vector<double> v(f->vals.size());
  for (size_t i = 0; i < v.size(); ++i) v[i] = s * f->vals[i];
  return mk_table(f->vars, std::move(v));
}
// factorDivide (Factor.hs:171-172)
FP factor_divide(const FP& f, double d) { return factor_scale(1.0 / d, f); }
// factorFromInstantiation (Factor.hs:90-94)
FP factor_from_instantiation(DV dv, int a) {
  std::vector<double> v(dv.d, 0.0);
  for (int i = 0; i < dv.d; ++i) v[i] = (i == a) ? 1.0 : 0.0;
  return mk_table({dv}, std::move(v));
}

// ---- Buckets (Buckets.hs) -------------------------------------------------------------
struct TraceStep {
  int kind;  // 0 = elim, 1 = final product
  int var;   // eliminated vertex (-1 for final)
  std::vector<Vars> inputs;
  Vars prod, out;
};
using Trace = std::vector<TraceStep>;

struct Buckets {
  Vars e;                                // remaining order
  std::map<DV, std::vector<FP>> m;       // Data.Map DV [f]
};

// createBuckets (Buckets.hs:53-65)
Buckets create_buckets(const std::vector<FP>& s, const Vars& e, const Vars& r) {
  Buckets b;
  b.e = e;
  b.e.insert(b.e.end(), r.begin(), r.end());
  std::vector<FP> rf = s;
  for (const DV& dv : b.e) {
    std::vector<FP> fk, rest;
    for (const FP& f : rf) (f->contains(dv) ? fk : rest).push_back(f);
    rf.swap(rest);
    b.m[dv] = fk;
  }
  return b;
}
// addBucket (Buckets.hs:92-98): prepend to the bucket of the first remaining variable it mentions.
void add_bucket(Buckets& b, const FP& f) {
  for (const DV& dv : b.e)
    if (f->contains(dv)) {
      auto& lst = b.m[dv];
      lst.insert(lst.begin(), f);
      return;
    }
}
// marginalizeOneVariable + updateBucket (Buckets.hs:74-89,105-111)
void marginalize_one_variable(Buckets& b, const DV& dv, Trace* tr) {
  std::vector<FP> fk = b.m.at(dv);
  FP p = factor_product(fk);
  FP f2 = factor_project_out({dv}, p);
  if (tr) {
    TraceStep t;
    t.kind = 0;
    t.var = dv.v;
    for (const FP& f : fk) t.inputs.push_back(f->vars);
    t.prod = p->vars;
    t.out = f2->vars;
    tr->push_back(t);
  }
  if (!b.e.empty()) b.e.erase(b.e.begin());
  if (f2->scalar) {
    b.m[dv] = {f2};
  } else {
    b.m.erase(dv);
    add_bucket(b, f2);
  }
}
// marginal (VariableElimination.hs:56-73)
FP marginal(const std::vector<FP>& lf, const Vars& p, const Vars& r,
            const std::vector<std::pair<DV, int>>& assignment, Trace* tr) {
  Buckets b = create_buckets(lf, p, r);
  for (auto& a : assignment) add_bucket(b, factor_from_instantiation(a.first, a.second));
  for (const DV& dv : p) marginalize_one_variable(b, dv, tr);
  std::vector<FP> rest;
  for (auto& kv : b.m) rest.insert(rest.end(), kv.second.begin(), kv.second.end());
  FP res = factor_product(rest);
  if (tr) {
    TraceStep t;
    t.kind = 1;
    t.var = -1;
    for (const FP& f : rest) t.inputs.push_back(f->vars);
    t.prod = res->vars;
    t.out = res->vars;
    tr->push_back(t);
  }
  return res;
}

// ---- Network -------------------------------------------------------------------------
struct Net {
  std::vector<int> dims;
  std::vector<std::vector<int>> parents;  // `ingoing` order == CPT scope order (Network.hs:69-79)
  std::vector<FP> cpts;                   // allVertexValues: ascending vertex id
  std::vector<std::string> names;
  int n() const { return (int)dims.size(); }
  DV dv(int v) const { return DV{v, dims[v]}; }
};

// ---- Hugin loader (HuginNet.hs; dialect: SURVEY Appendix C) ------------------------------
std::vector<std::string> split_any(const std::string& s, const char* delims) {
  std::vector<std::string> out;
  std::string cur;
  for (char c : s) {
    if (strchr(delims, c)) {
      if (!cur.empty()) out.push_back(cur), cur.clear();
    } else
      cur.push_back(c);
  }
  if (!cur.empty()) out.push_back(cur);
  return out;
}

Net* load_hugin(const char* path, std::string& err) {
  std::ifstream in(path);
  if (!in) {
    err = std::string("cannot open ") + path;
    return nullptr;
  }
  std::stringstream ss;
  ss << in.rdbuf();
  std::string raw = ss.str(), text;
  bool com = false;  // filterComment (HuginNet.hs:180-186)
  for (char c : raw) {
    if (com) {
      if (c == '\n') com = false, text.push_back(c);
    } else if (c == '%')
      com = true;
    else
      text.push_back(c);
  }
  struct Pot {
    std::vector<std::string> conds;
    std::vector<double> vals;
  };
  std::vector<std::pair<std::string, int>> nodes;
  std::vector<Pot> pots;
  size_t pos = 0;
  auto skip_ws = [&]() {
    while (pos < text.size() && isspace((unsigned char)text[pos])) ++pos;
  };
  while (true) {
    skip_ws();
    if (pos >= text.size()) break;
    size_t eol = text.find('\n', pos);
    if (eol == std::string::npos) {
      err = "unexpected end of file in section header";
      return nullptr;
    }
    std::string head = text.substr(pos, eol - pos);
    pos = eol + 1;
    if (text.compare(pos, 2, "{\n") != 0) {
      err = "expected '{' after '" + head + "'";
      return nullptr;
    }
    pos += 2;
    size_t close = text.find('}', pos);
    if (close == std::string::npos) {
      err = "missing '}'";
      return nullptr;
    }
    std::string body = text.substr(pos, close - pos);
    pos = close + 1;
    if (head.compare(0, 3, "net") == 0 && head.size() == 3) {
      continue;
    } else if (head.compare(0, 4, "node") == 0) {
      std::vector<std::string> tk = split_any(head.substr(4), " \t");
      if (tk.size() != 1) {
        err = "bad node header '" + head + "'";
        return nullptr;
      }
      int nstates = 0;
      std::stringstream bl(body);
      std::string line;
      while (std::getline(bl, line)) {
        size_t p0 = line.find_first_not_of(" \t");
        if (p0 == std::string::npos || line.compare(p0, 6, "states") != 0) continue;
        bool inq = false;
        for (char c : line)
          if (c == '"') {
            if (!inq) ++nstates;
            inq = !inq;
          }
      }
      nodes.push_back({tk[0], nstates});
    } else if (head.compare(0, 9, "potential") == 0) {
      Pot p;
      p.conds = split_any(head.substr(9), "()| ");  // splitCPT
      size_t dpos = body.find("data");
      if (dpos == std::string::npos) {
        err = "potential without data";
        return nullptr;
      }
      size_t eq = body.find('=', dpos), semi = body.find(';', dpos);
      for (auto& t : split_any(body.substr(eq + 1, semi - eq - 1), "() \n\t"))  // splitValues
        p.vals.push_back(strtod(t.c_str(), nullptr));
      pots.push_back(p);
    } else {
      err = "unknown section '" + head + "'";
      return nullptr;
    }
  }
  auto net = std::make_unique<Net>();
  std::map<std::string, int> ids;
  for (size_t i = 0; i < nodes.size(); ++i) {
    ids[nodes[i].first] = (int)i;
    net->names.push_back(nodes[i].first);
    net->dims.push_back(nodes[i].second);
  }
  net->parents.resize(nodes.size());
  net->cpts.resize(nodes.size());
  for (auto& p : pots) {
    Vars dvs;
    for (auto& c : p.conds) {
      if (!ids.count(c)) {
        err = "unknown variable " + c;
        return nullptr;
      }
      dvs.push_back(net->dv(ids[c]));
    }
    // addSection (HuginNet.hs:139-147): file layout = parents slowest .. child fastest;
    // changeVariableOrder (CPT.hs:44-51) re-lays it out as [child, parents...].
    Vars old(dvs.begin() + 1, dvs.end());
    old.push_back(dvs[0]);
    if ((int64_t)p.vals.size() != table_size(dvs)) {
      err = "potential size mismatch for " + p.conds[0];
      return nullptr;
    }
    std::vector<int64_t> st = index_strides_for(old, dvs);
    std::vector<double> nv((size_t)table_size(dvs));
    std::vector<int> idx(dvs.size(), 0);
    for (size_t e = 0; e < nv.size(); ++e) {
      int64_t o = 0;
      for (size_t j = 0; j < dvs.size(); ++j) o += st[j] * idx[j];
      nv[e] = p.vals[o];
      for (int j = (int)dvs.size() - 1; j >= 0; --j) {
        if (++idx[j] < dvs[j].d) break;
        idx[j] = 0;
      }
    }
    int child = dvs[0].v;
    for (size_t j = 1; j < dvs.size(); ++j) net->parents[child].push_back(dvs[j].v);
    net->cpts[child] = mk_table(dvs, std::move(nv));
  }
  for (size_t i = 0; i < nodes.size(); ++i)
    if (!net->cpts[i]) {
      err = "variable without potential: " + nodes[i].first;
      return nullptr;
    }
  return net.release();
}

// ---- Moral graph, triangulation, cluster tree (FactorElimination.hs) ------------------------
using Adj = std::map<int, std::set<int>>;

// moralGraph (FactorElimination.hs:384-427) incl. the accumulator quirk: `parents` is
// read from the graph being mutated, so marriage edges (added in BOTH directions) make co-parents
// each other's "parents" for later (higher-id) vertices.
Adj moral_graph(const Net& net) {
  std::set<std::pair<int, int>> edges;
  std::vector<std::vector<int>> ingoing(net.n());
  for (int v = 0; v < net.n(); ++v)
    for (int p : net.parents[v])
      if (edges.insert({p, v}).second) ingoing[v].push_back(p);
  for (int v = 0; v < net.n(); ++v) {
    std::vector<int> ps = ingoing[v];
    std::set<std::pair<int, int>> snap = edges;
    for (int x : ps)
      for (int y : ps)
        if (x != y && !snap.count({x, y}) && !snap.count({y, x}))
          if (edges.insert({x, y}).second) ingoing[y].push_back(x);
  }
  Adj adj;
  for (int v = 0; v < net.n(); ++v) adj[v];
  for (auto& e : edges) adj[e.first].insert(e.second), adj[e.second].insert(e.first);
  return adj;
}

// weightedEdges / numberOfAddedEdges (FactorElimination.hs:71-95): ordered pairs of non-adjacent neighbours.
// Integer in the reference; __int128 here is ample for the test networks (documented limit).
__int128 tri_metric(const Net& net, const Adj& adj, int v, int cmp) {
  __int128 tot = 0;
  const std::set<int>& nb = adj.at(v);
  for (int x : nb)
    for (int y : nb)
      if (x != y && !adj.at(x).count(y)) {
        if (cmp == 0)
          tot += (__int128)net.cpts[x]->vals.size() * (__int128)net.cpts[y]->vals.size();
        else
          tot += 1;
      }
  return tot;
}

struct JT {
  const Net* net = nullptr;
  std::vector<int> elim_order;
  std::vector<Vars> clusters;             // index = cluster-graph vertex id (maximal, renumbered)
  // tree, keyed by cluster id; "ascending Cluster order" lists are materialised where needed
  int root = 0;
  std::vector<std::vector<int>> children;  // cluster -> [sep] newest first
  std::vector<int> parent_sep;             // cluster -> sep or -1
  std::vector<int> sep_parent, sep_child;  // sep (1-based) -> cluster
  std::vector<Vars> sep_vars;
  std::vector<std::vector<FP>> factors, evidence;
  std::vector<std::vector<int>> factor_ids;  // vertex ids of the CPTs, same order as `factors`
  std::vector<FP> up, down;                  // per sep
  std::vector<int> by_order;                 // cluster ids in ascending Cluster order (Map.keys)
};

// keepMaximalClusters (FactorElimination.hs:162-186)
std::vector<std::set<int>> keep_maximal(const std::vector<std::set<int>>& l) {
  std::vector<std::set<int>> prefix;
  for (auto& cur : l) {
    int hit = -1;
    for (size_t i = 0; i < prefix.size(); ++i)
      if (std::includes(prefix[i].begin(), prefix[i].end(), cur.begin(), cur.end())) {
        hit = (int)i;
        break;
      }
    if (hit < 0)
      prefix.push_back(cur);
    else {
      auto r = prefix[hit];
      prefix.erase(prefix.begin() + hit);
      prefix.push_back(r);
    }
  }
  return prefix;
}

void refresh_order(JT& t, const std::vector<int>& members) {
  t.by_order = members;
  std::sort(t.by_order.begin(), t.by_order.end(),
            [&](int a, int b) { return t.clusters[a] < t.clusters[b]; });  // lexicographic on ascending DV lists
}

// createUninitializedJunctionTree (FactorElimination.hs:283-295): triangulate (:130-143) with the
// STATIC metric, createClusterGraph (:201-213), maximumSpanningTree (:223-276).
JT* build_tree(const Net& net, int cmp, const int* order) {
  Adj moral = moral_graph(net);
  std::map<int, __int128> key;
  if (order)
    for (int i = 0; i < net.n(); ++i) key[order[i]] = i;
  else
    for (int v = 0; v < net.n(); ++v) key[v] = tri_metric(net, moral, v, cmp);
  Adj g = moral;
  std::vector<std::set<int>> raw;
  auto t = std::make_unique<JT>();
  t->net = &net;
  while (!g.empty()) {
    int sel = -1;
    for (auto& kv : g)  // minimumBy over allVertices (ascending): first minimal
      if (sel < 0 || key[kv.first] < key[sel]) sel = kv.first;
    std::set<int> nb = g[sel];
    for (int x : nb)
      for (int y : nb)
        if (x != y) g[x].insert(y);
    for (int x : nb) g[x].erase(sel);
    g.erase(sel);
    std::set<int> c = nb;
    c.insert(sel);
    raw.push_back(c);
    t->elim_order.push_back(sel);
  }
  std::vector<std::set<int>> vc = keep_maximal(raw);
  const int nc = (int)vc.size();
  for (auto& c : vc) {
    Vars vs;
    for (int v : c) vs.push_back(net.dv(v));
    t->clusters.push_back(vs);
  }
  t->children.resize(nc);
  t->parent_sep.assign(nc, -1);
  t->factors.resize(nc);
  t->evidence.resize(nc);
  t->factor_ids.resize(nc);
  t->sep_parent = {-1};
  t->sep_child = {-1};
  t->sep_vars = {Vars{}};
  std::vector<int> members = {0}, remaining;
  for (int i = 1; i < nc; ++i) remaining.push_back(i);
  refresh_order(*t, members);
  while (!remaining.empty()) {
    // findMax (:238-250): maximumBy => LAST maximal over rv <- remaining, lv <- treeNodes
    int brv = -1, blv = -1, bev = -1;
    for (int rv : remaining)
      for (int lv : t->by_order) {
        std::vector<int> inter;
        std::set_intersection(vc[rv].begin(), vc[rv].end(), vc[lv].begin(), vc[lv].end(),
                              std::back_inserter(inter));
        int ev = (int)inter.size();
        if (brv < 0 || !(bev > ev)) brv = rv, blv = lv, bev = ev;
      }
    remaining.erase(std::find(remaining.begin(), remaining.end(), brv));
    Vars sep;
    for (const DV& x : t->clusters[brv])
      if (has(t->clusters[blv], x)) sep.push_back(x);
    int ns = (int)t->sep_parent.size();
    t->sep_parent.push_back(blv);
    t->sep_child.push_back(brv);
    t->sep_vars.push_back(sep);
    t->children[blv].insert(t->children[blv].begin(), ns);  // newest first (JTree.hs:212)
    t->parent_sep[brv] = ns;
    members.push_back(brv);
    refresh_order(*t, members);
  }
  t->up.assign(t->sep_parent.size(), nullptr);
  t->down.assign(t->sep_parent.size(), nullptr);
  return t.release();
}

// updateTreeValues (JTree.hs:424-440): first minimal-size covering cluster in ascending Cluster order.
int covering_cluster(const JT& t, const Vars& fv) {
  int best = -1;
  __int128 bsz = 0;
  for (int c : t.by_order) {
    bool ok = true;
    for (const DV& v : fv)
      if (!has(t.clusters[c], v)) {
        ok = false;
        break;
      }
    if (!ok) continue;
    __int128 sz = 1;
    for (const DV& v : t.clusters[c]) sz *= v.d;
    if (best < 0 || sz < bsz) best = c, bsz = sz;
  }
  return best;
}

// newMessage (JTree.hs:485-491)
FP new_message(const std::vector<FP>& input, const std::vector<FP>& f, const std::vector<FP>& e,
               const Vars& sepc, Trace* tr) {
  std::vector<FP> all = f;
  all.insert(all.end(), e.begin(), e.end());
  all.insert(all.end(), input.begin(), input.end());
  Vars seen;
  for (const FP& x : all)
    for (const DV& v : x->vars)
      if (!has(seen, v)) seen.push_back(v);  // nub . concatMap factorVariables
  return marginal(all, l_diff(seen, sepc), sepc, {}, tr);
}

struct MsgTrace {
  int kind;  // 0 up, 1 down, 2 posterior
  int cluster, sep, qvar;
  Trace steps;
};

// collect (JTree.hs:329-348) + updateUpSeparator (:288-305)
void collect(JT& t, std::vector<MsgTrace>* tr) {
  std::vector<int> nodes;
  for (int c : t.by_order)
    if (t.children[c].empty()) nodes.push_back(c);
  while (!nodes.empty()) {
    for (int h : nodes) {
      bool ready = true;
      for (int s : t.children[h])
        if (!t.up[s]) ready = false;
      if (!ready || t.parent_sep[h] < 0) continue;
      std::vector<FP> inc;
      for (int s : t.children[h]) inc.push_back(t.up[s]);
      int p = t.parent_sep[h];
      MsgTrace mt{0, h, p, -1, {}};
      t.up[p] = new_message(inc, t.factors[h], t.evidence[h], t.sep_vars[p], tr ? &mt.steps : nullptr);
      if (tr) tr->push_back(mt);
    }
    std::set<int> ps;
    for (int h : nodes)
      if (t.parent_sep[h] >= 0) ps.insert(t.sep_parent[t.parent_sep[h]]);
    nodes.assign(ps.begin(), ps.end());
    std::sort(nodes.begin(), nodes.end(), [&](int a, int b) { return t.clusters[a] < t.clusters[b]; });
  }
}

// distribute (JTree.hs:350-369) + updateDownSeparator (:307-322)
void distribute_node(JT& t, int node, std::vector<MsgTrace>* tr) {
  for (int child : t.children[node]) {
    std::vector<FP> inc;
    if (t.parent_sep[node] >= 0 && t.down[t.parent_sep[node]]) inc.push_back(t.down[t.parent_sep[node]]);
    for (int s : t.children[node])
      if (s != child) inc.push_back(t.up[s]);
    MsgTrace mt{1, node, child, -1, {}};
    t.down[child] = new_message(inc, t.factors[node], t.evidence[node], t.sep_vars[child], tr ? &mt.steps : nullptr);
    if (tr) tr->push_back(mt);
  }
  for (int child : t.children[node]) distribute_node(t, t.sep_child[child], tr);
}

void calibrate(JT& t, std::vector<MsgTrace>* tr) {
  std::fill(t.up.begin(), t.up.end(), nullptr);
  std::fill(t.down.begin(), t.down.end(), nullptr);
  collect(t, tr);
  distribute_node(t, t.root, tr);
}

// createJunctionTree (FactorElimination.hs:298-307) + setFactors (JTree.hs:443-451)
JT* create_junction_tree(const Net& net, int cmp, const int* order) {
  JT* t = build_tree(net, cmp, order);
  for (int v = 0; v < net.n(); ++v) {
    int c = covering_cluster(*t, net.cpts[v]->vars);
    t->factors[c].insert(t->factors[c].begin(), net.cpts[v]);  // f : oldf
    t->factor_ids[c].insert(t->factor_ids[c].begin(), v);
  }
  calibrate(*t, nullptr);
  return t;
}

// changeEvidence (JTree.hs:454-466): caller's list order kept; each indicator prepended.
void change_evidence(JT& t, int n, const int* vars, const int* states, std::vector<MsgTrace>* tr) {
  for (auto& e : t.evidence) e.clear();
  for (int i = 0; i < n; ++i) {
    DV dv = t.net->dv(vars[i]);
    int c = covering_cluster(t, {dv});
    t.evidence[c].insert(t.evidence[c].begin(), factor_from_instantiation(dv, states[i]));
  }
  calibrate(t, tr);
}

// traverseTree (findClusterFor v) (JTree.hs:398-415, FactorElimination.hs:330-338): a matching
// cluster stops the descent below it, but siblings are still visited and overwrite the state.
void find_cluster(const JT& t, int cur, const Vars& v, int& state) {
  bool sub = true;
  for (const DV& x : v)
    if (!has(t.clusters[cur], x)) sub = false;
  if (sub) {
    state = cur;
    return;
  }
  for (int s : t.children[cur]) find_cluster(t, t.sep_child[s], v, state);
}

// posterior (FactorElimination.hs:314-327)
FP posterior(const JT& t, const Vars& v, MsgTrace* mt) {
  int c = -1;
  find_cluster(t, t.root, v, c);
  if (c < 0) return nullptr;
  std::vector<FP> all;
  FP d = (t.parent_sep[c] >= 0 && t.down[t.parent_sep[c]]) ? t.down[t.parent_sep[c]] : mk_scalar(1.0);
  all.push_back(d);
  for (int s : t.children[c]) all.push_back(t.up[s]);
  all.insert(all.end(), t.factors[c].begin(), t.factors[c].end());
  all.insert(all.end(), t.evidence[c].begin(), t.evidence[c].end());
  Vars seen;
  for (const FP& x : all)
    for (const DV& y : x->vars)
      if (!has(seen, y)) seen.push_back(y);
  if (mt) mt->cluster = c;
  FP un = marginal(all, l_diff(seen, v), v, {}, mt ? &mt->steps : nullptr);
  return factor_divide(un, factor_norm(un));
}

// ---- VE orders (VariableElimination.hs:150-236) ------------------------------------------
// interactionGraph: only variables occurring in a factor with >= 2 variables get a vertex;
// eliminationOrderForMetric: greedy, first minimal over ascending ids, vertex removed WITHOUT fill-in.
std::vector<int> ve_order(const Net& net, int metric /*0 minDegree, 1 minFill*/) {
  Adj adj;
  for (const FP& f : net.cpts)
    for (const DV& x : f->vars)
      for (const DV& y : f->vars)
        if (x.v != y.v) adj[x.v].insert(y.v), adj[y.v].insert(x.v);
  std::vector<int> out;
  while (!adj.empty()) {
    int best = -1;
    long bm = 0;
    for (auto& kv : adj) {
      long m = 0;
      if (metric == 0)
        m = (long)kv.second.size();
      else
        for (int x : kv.second)
          for (int y : kv.second)
            if (x != y && !adj[x].count(y)) ++m;
      if (best < 0 || m < bm) best = kv.first, bm = m;
    }
    for (int x : adj[best]) adj[x].erase(best);
    adj.erase(best);
    out.push_back(best);
  }
  return out;
}

// ---- JSON helpers --------------------------------------------------------------------------
void js_vars(std::ostringstream& o, const Vars& v) {
  o << "[";
  for (size_t i = 0; i < v.size(); ++i) o << (i ? "," : "") << v[i].v;
  o << "]";
}
void js_ints(std::ostringstream& o, const std::vector<int>& v) {
  o << "[";
  for (size_t i = 0; i < v.size(); ++i) o << (i ? "," : "") << v[i];
  o << "]";
}
void js_trace(std::ostringstream& o, const Trace& tr) {
  o << "[";
  for (size_t i = 0; i < tr.size(); ++i) {
    const TraceStep& s = tr[i];
    o << (i ? "," : "") << "{\"kind\":\"" << (s.kind == 0 ? "elim" : "final") << "\",\"var\":" << s.var
      << ",\"inputs\":[";
    for (size_t j = 0; j < s.inputs.size(); ++j) {
      o << (j ? "," : "");
      js_vars(o, s.inputs[j]);
    }
    o << "],\"prod\":";
    js_vars(o, s.prod);
    o << ",\"out\":";
    js_vars(o, s.out);
    o << "}";
  }
  o << "]";
}

thread_local std::string g_err;
char* dup_str(const std::string& s) {
  char* p = (char*)malloc(s.size() + 1);
  memcpy(p, s.c_str(), s.size() + 1);
  return p;
}

}  // namespace

// ================================ C API (ctypes) ============================================
extern "C" {

const char* orc_last_error() { return g_err.c_str(); }
void orc_string_free(char* p) { free(p); }

void* orc_net_create(int n_vars, const int* dims, const int64_t* cpt_off, const int* cpt_scope,
                     const int64_t* val_off, const double* vals) {
  auto net = std::make_unique<Net>();
  net->dims.assign(dims, dims + n_vars);
  net->parents.resize(n_vars);
  net->cpts.resize(n_vars);
  for (int v = 0; v < n_vars; ++v) {
    net->names.push_back("n" + std::to_string(v));
    Vars sc;
    for (int64_t i = cpt_off[v]; i < cpt_off[v + 1]; ++i) sc.push_back(net->dv(cpt_scope[i]));
    if (sc.empty() || sc[0].v != v || table_size(sc) != val_off[v + 1] - val_off[v]) {
      g_err = "inconsistent CPT for vertex " + std::to_string(v);
      return nullptr;
    }
    for (size_t j = 1; j < sc.size(); ++j) net->parents[v].push_back(sc[j].v);
    net->cpts[v] = mk_table(sc, std::vector<double>(vals + val_off[v], vals + val_off[v + 1]));
  }
  return net.release();
}
void* orc_net_load_hugin(const char* path) {
  std::string err;
  Net* n = load_hugin(path, err);
  if (!n) g_err = err;
  return n;
}
void orc_net_free(void* n) { delete (Net*)n; }
int orc_net_nvars(void* n) { return ((Net*)n)->n(); }

char* orc_net_describe_json(void* np) {
  Net& n = *(Net*)np;
  std::ostringstream o;
  o.precision(17);
  o << "{\"names\":[";
  for (int v = 0; v < n.n(); ++v) o << (v ? "," : "") << "\"" << n.names[v] << "\"";
  o << "],\"dims\":";
  js_ints(o, n.dims);
  o << ",\"scopes\":[";
  for (int v = 0; v < n.n(); ++v) {
    o << (v ? "," : "");
    js_vars(o, n.cpts[v]->vars);
  }
  o << "],\"values\":[";
  for (int v = 0; v < n.n(); ++v) {
    o << (v ? "," : "") << "[";
    for (size_t i = 0; i < n.cpts[v]->vals.size(); ++i) o << (i ? "," : "") << n.cpts[v]->vals[i];
    o << "]";
  }
  o << "]}";
  return dup_str(o.str());
}

// cmp: 0 = nodeComparisonForTriangulation (weightedEdges), 1 = numberOfAddedEdges
void* orc_jt_create(void* net, int cmp) { return create_junction_tree(*(Net*)net, cmp, nullptr); }
void* orc_jt_create_with_order(void* net, const int* order) { return create_junction_tree(*(Net*)net, 0, order); }
void orc_jt_free(void* t) { delete (JT*)t; }

// changeEvidence; when trace_json != NULL also returns the op trace of this recalibration plus the
// traces of posterior [v] for every v (what the product's compiled step list must equal).
int orc_jt_change_evidence(void* tp, int n, const int* vars, const int* states, char** trace_json) {
  JT& t = *(JT*)tp;
  std::vector<MsgTrace> tr;
  change_evidence(t, n, vars, states, trace_json ? &tr : nullptr);
  if (trace_json) {
    for (int v = 0; v < t.net->n(); ++v) {
      MsgTrace mt{2, -1, -1, v, {}};
      posterior(t, {t.net->dv(v)}, &mt);
      tr.push_back(mt);
    }
    std::ostringstream o;
    o << "[";
    for (size_t i = 0; i < tr.size(); ++i) {
      const char* kn[] = {"up", "down", "posterior"};
      o << (i ? "," : "") << "{\"kind\":\"" << kn[tr[i].kind] << "\",\"cluster\":" << tr[i].cluster
        << ",\"sep\":" << tr[i].sep << ",\"qvar\":" << tr[i].qvar << ",\"steps\":";
      js_trace(o, tr[i].steps);
      o << "}";
    }
    o << "]";
    *trace_json = dup_str(o.str());
  }
  return 0;
}

// posterior: returns number of values written, -1 when no cluster covers the query (Haskell Nothing)
int64_t orc_jt_posterior(void* tp, int nq, const int* qvars, int* out_scope, double* out, int64_t out_len) {
  JT& t = *(JT*)tp;
  Vars v;
  for (int i = 0; i < nq; ++i) v.push_back(t.net->dv(qvars[i]));
  FP r = posterior(t, v, nullptr);
  if (!r) return -1;
  if ((int64_t)r->vals.size() > out_len) return -2;
  for (size_t i = 0; i < r->vars.size() && out_scope; ++i) out_scope[i] = r->vars[i].v;
  memcpy(out, r->vals.data(), r->vals.size() * sizeof(double));
  return (int64_t)r->vals.size();
}

char* orc_jt_describe_json(void* tp) {
  JT& t = *(JT*)tp;
  std::ostringstream o;
  o << "{\"elim_order\":";
  js_ints(o, t.elim_order);
  o << ",\"clusters\":[";
  for (size_t i = 0; i < t.clusters.size(); ++i) {
    o << (i ? "," : "");
    js_vars(o, t.clusters[i]);
  }
  o << "],\"root\":" << t.root << ",\"parent_sep\":";
  js_ints(o, t.parent_sep);
  o << ",\"sep_parent\":";
  js_ints(o, std::vector<int>(t.sep_parent.begin() + 1, t.sep_parent.end()));
  o << ",\"sep_child\":";
  js_ints(o, std::vector<int>(t.sep_child.begin() + 1, t.sep_child.end()));
  o << ",\"sep_vars\":[";
  for (size_t i = 1; i < t.sep_vars.size(); ++i) {
    o << (i > 1 ? "," : "");
    js_vars(o, t.sep_vars[i]);
  }
  o << "],\"children\":[";
  for (size_t i = 0; i < t.children.size(); ++i) {
    o << (i ? "," : "");
    js_ints(o, t.children[i]);
  }
  o << "],\"factors\":[";
  for (size_t i = 0; i < t.factor_ids.size(); ++i) {
    o << (i ? "," : "");
    js_ints(o, t.factor_ids[i]);
  }
  o << "]}";
  return dup_str(o.str());
}

// One JT query per row: changeEvidence [v =: s | v ascending, observed] then posterior [v]
// for every v; out = n_rows x sum_v dims[v].  Rows are split contiguously over n_threads, each thread
// on its own copy of the tree (the reference itself is single-threaded; this is the harness's row split).
int orc_jt_posteriors_batch(void* tp, int64_t n_rows, const int8_t* evidence, double* out, int n_threads) {
  JT& proto = *(JT*)tp;
  const Net& net = *proto.net;
  const int nv = net.n();
  int64_t width = 0;
  for (int d : net.dims) width += d;
  if (n_threads < 1) n_threads = 1;
  auto work = [&](int64_t lo, int64_t hi) {
    JT t = proto;
    std::vector<int> vars, states;
    for (int64_t r = lo; r < hi; ++r) {
      vars.clear();
      states.clear();
      for (int v = 0; v < nv; ++v)
        if (evidence[r * nv + v] >= 0) vars.push_back(v), states.push_back(evidence[r * nv + v]);
      change_evidence(t, (int)vars.size(), vars.data(), states.data(), nullptr);
      double* o = out + r * width;
      for (int v = 0; v < nv; ++v) {
        FP p = posterior(t, {net.dv(v)}, nullptr);
        for (int i = 0; i < net.dims[v]; ++i) o[i] = p->vals[i];
        o += net.dims[v];
      }
    }
  };
  std::vector<std::thread> th;
  for (int i = 0; i < n_threads; ++i) {
    int64_t lo = n_rows * i / n_threads, hi = n_rows * (i + 1) / n_threads;
    if (lo < hi) th.emplace_back(work, lo, hi);
  }
  for (auto& x : th) x.join();
  return 0;
}

// posteriorMarginal g p r assignment (VariableElimination.hs:116-130); assignment in the caller's order.
int64_t orc_ve_posterior(void* np, const int* p, int n_p, const int* r, int n_r, int n_ev, const int* ev_vars,
                         const int* ev_states, int* out_scope, double* out, int64_t out_len, char** trace_json) {
  Net& net = *(Net*)np;
  Vars pv, rv;
  for (int i = 0; i < n_p; ++i) pv.push_back(net.dv(p[i]));
  for (int i = 0; i < n_r; ++i) rv.push_back(net.dv(r[i]));
  std::vector<std::pair<DV, int>> asg;
  for (int i = 0; i < n_ev; ++i) asg.push_back({net.dv(ev_vars[i]), ev_states[i]});
  Trace tr;
  FP res = marginal(net.cpts, pv, rv, asg, trace_json ? &tr : nullptr);
  FP nr = factor_divide(res, factor_norm(res));
  if (trace_json) {
    std::ostringstream o;
    js_trace(o, tr);
    *trace_json = dup_str(o.str());
  }
  if ((int64_t)nr->vals.size() > out_len) return -2;
  for (size_t i = 0; i < nr->vars.size() && out_scope; ++i) out_scope[i] = nr->vars[i].v;
  memcpy(out, nr->vals.data(), nr->vals.size() * sizeof(double));
  return (int64_t)nr->vals.size();
}

// Batched VE: one posteriorMarginal per evidence row (ascending-id assignment), rows over threads.
int orc_ve_posterior_batch(void* np, const int* p, int n_p, const int* r, int n_r, int64_t n_rows,
                           const int8_t* evidence, double* out, int64_t out_width, int n_threads) {
  Net& net = *(Net*)np;
  const int nv = net.n();
  if (n_threads < 1) n_threads = 1;
  auto work = [&](int64_t lo, int64_t hi) {
    std::vector<int> ev, es;
    for (int64_t row = lo; row < hi; ++row) {
      ev.clear();
      es.clear();
      for (int v = 0; v < nv; ++v)
        if (evidence[row * nv + v] >= 0) ev.push_back(v), es.push_back(evidence[row * nv + v]);
      orc_ve_posterior(np, p, n_p, r, n_r, (int)ev.size(), ev.data(), es.data(), nullptr,
                       out + row * out_width, out_width, nullptr);
    }
  };
  std::vector<std::thread> th;
  for (int i = 0; i < n_threads; ++i) {
    int64_t lo = n_rows * i / n_threads, hi = n_rows * (i + 1) / n_threads;
    if (lo < hi) th.emplace_back(work, lo, hi);
  }
  for (auto& x : th) x.join();
  return 0;
}

// minDegreeOrder (metric 0) / minFillOrder (metric 1); returns count written (interaction-graph vertices only)
int orc_ve_order(void* np, int metric, int* out) {
  std::vector<int> o = ve_order(*(Net*)np, metric);
  for (size_t i = 0; i < o.size(); ++i) out[i] = o[i];
  return (int)o.size();
}

// Factor-algebra entry points for the property tests (CPT.hs:68-98).
// product of k tables: scopes concatenated in `scopes` with offsets `soff`; returns out size, fills out_scope.
int64_t orc_factor_product(int k, const int* soff, const int* scopes, const int* dims_of_scope, const int64_t* voff,
                           const double* vals, int* out_scope, int* out_nvars, double* out, int64_t out_len) {
  std::vector<FP> l;
  for (int i = 0; i < k; ++i) {
    Vars sc;
    for (int j = soff[i]; j < soff[i + 1]; ++j) sc.push_back(DV{scopes[j], dims_of_scope[j]});
    l.push_back(mk_table(sc, std::vector<double>(vals + voff[i], vals + voff[i + 1])));
  }
  FP r = factor_product(l);
  if ((int64_t)r->vals.size() > out_len) return -2;
  *out_nvars = (int)r->vars.size();
  for (size_t i = 0; i < r->vars.size(); ++i) out_scope[i] = r->vars[i].v;
  memcpy(out, r->vals.data(), r->vals.size() * sizeof(double));
  return (int64_t)r->vals.size();
}
int64_t orc_factor_project_out(int nvars, const int* scope, const int* dims, const double* vals, int n_sum,
                               const int* sum_vars, int* out_scope, int* out_nvars, double* out, int64_t out_len) {
  Vars sc, s;
  for (int i = 0; i < nvars; ++i) sc.push_back(DV{scope[i], dims[i]});
  for (int i = 0; i < n_sum; ++i)
    for (const DV& x : sc)
      if (x.v == sum_vars[i]) s.push_back(x);
  FP f = mk_table(sc, std::vector<double>(vals, vals + table_size(sc)));
  FP r = factor_project_out(s, f);
  if ((int64_t)r->vals.size() > out_len) return -2;
  *out_nvars = (int)r->vars.size();
  for (size_t i = 0; i < r->vars.size(); ++i) out_scope[i] = r->vars[i].v;
  memcpy(out, r->vals.data(), r->vals.size() * sizeof(double));
  return (int64_t)r->vals.size();
}

}  // extern "C"
