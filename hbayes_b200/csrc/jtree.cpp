// jtree.cpp -- junction-tree construction and VE elimination orders (host, once per network).
//
// Integer/graph work only.  Every ordering rule that decides the shape of the tree in the
// reference is reproduced; the data structures are not: the
// reference's IntMap/Map graphs become bit matrices here.
#include <algorithm>
#include <functional>
#include <memory>
#include <sstream>

#include "model.hpp"

namespace hb {
namespace {

typedef unsigned __int128 u128;

// Square bit matrix over the vertices.
struct BitMatrix {
  int n = 0, w = 0;
  std::vector<uint64_t> bits;
  explicit BitMatrix(int n_) : n(n_), w((n_ + 63) / 64), bits((size_t)n_ * ((n_ + 63) / 64), 0) {}
  uint64_t* row(int r) { return bits.data() + (size_t)r * w; }
  const uint64_t* row(int r) const { return bits.data() + (size_t)r * w; }
  bool get(int r, int c) const { return (row(r)[c >> 6] >> (c & 63)) & 1; }
  void set(int r, int c) { row(r)[c >> 6] |= uint64_t(1) << (c & 63); }
  void clear(int r, int c) { row(r)[c >> 6] &= ~(uint64_t(1) << (c & 63)); }
  std::vector<int> members(int r) const {
    std::vector<int> out;
    const uint64_t* p = row(r);
    for (int i = 0; i < w; ++i)
      for (uint64_t x = p[i]; x; x &= x - 1) out.push_back(i * 64 + __builtin_ctzll(x));
    return out;
  }
};

// moralGraph (Bayes/FactorElimination.hs:398-427).  The fold visits vertices in ascending id and
// reads `parents` from the graph it is mutating; marriages add BOTH directed edges, so co-parents of
// a lower-id child count as each other's parents when a higher-id vertex is visited (A.6).
BitMatrix moral_graph(const Network& net) {
  const int n = net.n();
  BitMatrix into(n);  // into.get(v, x): directed edge x -> v exists
  for (int v = 0; v < n; ++v)
    for (size_t j = 1; j < net.scope[v].size(); ++j) into.set(v, net.scope[v][j]);
  for (int v = 0; v < n; ++v) {
    std::vector<int> ps = into.members(v);
    std::vector<std::pair<int, int>> marry;
    for (size_t a = 0; a < ps.size(); ++a)
      for (size_t b = a + 1; b < ps.size(); ++b)
        if (!into.get(ps[a], ps[b]) && !into.get(ps[b], ps[a])) marry.push_back({ps[a], ps[b]});
    for (auto& m : marry) into.set(m.first, m.second), into.set(m.second, m.first);
  }
  BitMatrix und(n);  // convertToUndirected (:409-420)
  for (int v = 0; v < n; ++v)
    for (int x : into.members(v)) und.set(v, x), und.set(x, v);
  return und;
}

// weightedEdges / numberOfAddedEdges (FactorElimination.hs:71-95) on the ORIGINAL moral graph:
// sum over ordered pairs of distinct, non-adjacent neighbours.  Integer in the reference.
u128 triangulation_metric(const Network& net, const BitMatrix& moral, int v, int cmp) {
  std::vector<int> nb = moral.members(v);
  u128 tot = 0;
  for (int x : nb)
    for (int y : nb)
      if (x != y && !moral.get(x, y)) tot += cmp == 0 ? (u128)net.cpt_entries(x) * (u128)net.cpt_entries(y) : (u128)1;
  return tot;
}

bool subset_of(const std::vector<uint64_t>& a, const std::vector<uint64_t>& b) {
  for (size_t i = 0; i < a.size(); ++i)
    if (a[i] & ~b[i]) return false;
  return true;
}

u128 state_space(const Network& net, const std::vector<int>& vars) {
  u128 s = 1;
  const u128 cap = (u128)1 << 120;
  for (int v : vars) {
    s *= (u128)net.dims[v];
    if (s > cap) s = cap;  // saturate: such a cluster is unusable anyway
  }
  return s;
}

}  // namespace

Tree* tree_build(const Network& net, int cmp, const int32_t* order) {
  const int n = net.n();
  if (cmp != 0 && cmp != 1) throw Error(HB_E_ARG, "unknown triangulation comparator");
  BitMatrix moral = moral_graph(net);

  // triangulate (FactorElimination.hs:130-143).  `cmp` is closed over the original moral graph
  // (:291-292), so minimumBy over the remaining vertices (ascending id, first minimal) amounts to a
  // stable sort by a static key.
  std::vector<int> seq(n);
  if (order) {
    std::vector<char> seen(n, 0);
    for (int i = 0; i < n; ++i) {
      if (order[i] < 0 || order[i] >= n || seen[order[i]]) throw Error(HB_E_ARG, "elimination order is not a permutation");
      seen[order[i]] = 1;
      seq[i] = order[i];
    }
  } else {
    std::vector<u128> key(n);
    for (int v = 0; v < n; ++v) key[v] = triangulation_metric(net, moral, v, cmp), seq[v] = v;
    std::stable_sort(seq.begin(), seq.end(), [&](int a, int b) { return key[a] < key[b]; });
  }
  auto t = std::make_unique<Tree>();
  t->elim_order = seq;
  const int W = moral.w;
  BitMatrix g = moral;
  std::vector<std::vector<uint64_t>> raw;  // one cluster per eliminated vertex
  for (int v : seq) {
    std::vector<uint64_t> nb(g.row(v), g.row(v) + W);
    for (int x : g.members(v)) {
      uint64_t* rx = g.row(x);
      for (int i = 0; i < W; ++i) rx[i] |= nb[i];
      g.clear(x, x);
      g.clear(x, v);
    }
    std::fill(g.row(v), g.row(v) + W, 0);
    nb[v >> 6] |= uint64_t(1) << (v & 63);
    raw.push_back(nb);
  }
  // keepMaximalClusters (:162-186): a cluster contained in an earlier kept one is dropped and the
  // (first) containing cluster moves to the end of the kept list.
  std::vector<std::vector<uint64_t>> kept;
  for (auto& cur : raw) {
    size_t hit = kept.size();
    for (size_t i = 0; i < kept.size(); ++i)
      if (subset_of(cur, kept[i])) {
        hit = i;
        break;
      }
    if (hit == kept.size())
      kept.push_back(cur);
    else
      std::rotate(kept.begin() + hit, kept.begin() + hit + 1, kept.end());
  }
  const int nc = (int)kept.size();
  t->clusters.resize(nc);
  for (int c = 0; c < nc; ++c)
    for (int i = 0; i < W; ++i)
      for (uint64_t x = kept[c][i]; x; x &= x - 1) t->clusters[c].push_back(i * 64 + __builtin_ctzll(x));

  // createClusterGraph (:201-213): complete graph, weight = number of shared vertices.
  std::vector<int> share((size_t)nc * nc, 0);
  for (int a = 0; a < nc; ++a)
    for (int b = a + 1; b < nc; ++b) {
      int s = 0;
      for (int i = 0; i < W; ++i) s += __builtin_popcountll(kept[a][i] & kept[b][i]);
      share[(size_t)a * nc + b] = share[(size_t)b * nc + a] = s;
    }
  // rank of each cluster in the derived Ord of `Cluster` = lexicographic on ascending variable lists
  std::vector<int> by_rank(nc), rank(nc);
  for (int c = 0; c < nc; ++c) by_rank[c] = c;
  std::sort(by_rank.begin(), by_rank.end(), [&](int a, int b) { return t->clusters[a] < t->clusters[b]; });
  for (int i = 0; i < nc; ++i) rank[by_rank[i]] = i;
  t->by_rank = by_rank;

  // maximumSpanningTree / buildTree / findMax (:223-276): root = cluster-graph vertex 0; each round takes
  // the LAST maximal (remaining vertex ascending) x (tree node ascending Cluster) pair (maximumBy).
  t->root = 0;
  t->parent_sep.assign(nc, -1);
  t->children.assign(nc, {});
  t->factors.assign(nc, {});
  std::vector<char> in_tree(nc, 0);
  in_tree[0] = 1;
  for (int added = 1; added < nc; ++added) {
    int best_rv = -1, best_lv = -1, best_w = -1;
    for (int rv = 0; rv < nc; ++rv) {
      if (in_tree[rv]) continue;
      for (int lv = 0; lv < nc; ++lv) {
        if (!in_tree[lv]) continue;
        int w = share[(size_t)rv * nc + lv];
        // later pairs in the enumeration win ties: larger rv, then larger Cluster rank of lv
        if (w > best_w || (w == best_w && (rv > best_rv || rank[lv] > rank[best_lv]))) best_rv = rv, best_lv = lv, best_w = w;
      }
    }
    std::vector<int> sep;
    std::set_intersection(t->clusters[best_rv].begin(), t->clusters[best_rv].end(), t->clusters[best_lv].begin(),
                          t->clusters[best_lv].end(), std::back_inserter(sep));
    t->sep_parent.push_back(best_lv);
    t->sep_child.push_back(best_rv);
    t->sep_vars.push_back(sep);
    const int sep_no = (int)t->sep_parent.size();  // separator keys start at 1 (JTree.hs:203-214)
    t->children[best_lv].insert(t->children[best_lv].begin(), sep_no);
    t->parent_sep[best_rv] = sep_no;
    in_tree[best_rv] = 1;
  }
  // setFactors (JTree.hs:443-451): CPTs in ascending vertex id, each PREPENDED to its cluster's list.
  for (int v = 0; v < n; ++v) {
    int c = tree_covering_cluster(net, *t, net.scope[v]);
    if (c < 0) throw Error(HB_E_ARG, "no cluster covers the family of vertex " + std::to_string(v));
    t->factors[c].insert(t->factors[c].begin(), v);
  }
  return t.release();
}

// updateTreeValues (JTree.hs:424-440): minimumBy clusterSize over the covering clusters in treeNodes
// (= ascending Cluster) order, i.e. the first minimal one.
int tree_covering_cluster(const Network& net, const Tree& t, const std::vector<int>& vars) {
  int best = -1;
  u128 best_size = 0;
  for (int c : t.by_rank) {
    const auto& cl = t.clusters[c];
    bool covers = true;
    for (int v : vars)
      if (!std::binary_search(cl.begin(), cl.end(), v)) {
        covers = false;
        break;
      }
    if (!covers) continue;
    u128 s = state_space(net, cl);
    if (best < 0 || s < best_size) best = c, best_size = s;
  }
  return best;
}

// traverseTree (findClusterFor v) (JTree.hs:398-415, FactorElimination.hs:330-338): pre-order over the
// newest-first children; a covering cluster is not descended into, but later siblings still run and
// overwrite the answer.
int tree_find_cluster(const Tree& t, const std::vector<int>& q) {
  int found = -1;
  std::function<void(int)> visit = [&](int c) {
    const auto& cl = t.clusters[c];
    bool covers = true;
    for (int v : q)
      if (!std::binary_search(cl.begin(), cl.end(), v)) covers = false;
    if (covers) {
      found = c;
      return;
    }
    for (int s : t.children[c]) visit(t.sep_child[s - 1]);
  };
  visit(t.root);
  return found;
}

// minDegreeOrder / minFillOrder (VariableElimination.hs:150-236).  The interaction graph only has the
// vertices that share a factor with another variable; the greedy loop removes the chosen vertex
// WITHOUT adding fill-in edges (A.11).
std::vector<int> ve_elimination_order(const Network& net, int metric) {
  const int n = net.n();
  if (metric != 0 && metric != 1) throw Error(HB_E_ARG, "unknown elimination metric");
  BitMatrix adj(n);
  std::vector<char> present(n, 0);
  for (int v = 0; v < n; ++v)
    for (int x : net.scope[v])
      for (int y : net.scope[v])
        if (x != y) adj.set(x, y), present[x] = present[y] = 1;
  std::vector<int> out;
  int left = 0;
  for (int v = 0; v < n; ++v) left += present[v];
  while (left--) {
    int best = -1;
    long best_m = 0;
    for (int v = 0; v < n; ++v) {
      if (!present[v]) continue;
      std::vector<int> nb = adj.members(v);
      long m = 0;
      if (metric == 0)
        m = (long)nb.size();
      else
        for (int x : nb)
          for (int y : nb)
            if (x != y && !adj.get(x, y)) ++m;
      if (best < 0 || m < best_m) best = v, best_m = m;
    }
    for (int x : adj.members(best)) adj.clear(x, best);
    std::fill(adj.row(best), adj.row(best) + adj.w, 0);
    present[best] = 0;
    out.push_back(best);
  }
  return out;
}

// degreeOrder (VariableElimination.hs:186-207): the largest number of neighbours a variable has at the moment it is
// eliminated, on the interaction graph WITH fill-in edges between its neighbours (the width the order induces).
int ve_degree_order(const Network& net, const std::vector<int>& order) {
  const int n = net.n();
  BitMatrix adj(n);
  for (int v = 0; v < n; ++v)
    for (int x : net.scope[v])
      for (int y : net.scope[v])
        if (x != y) adj.set(x, y);
  int width = 0;
  for (int v : order) {
    if (v < 0 || v >= n) throw Error(HB_E_ARG, "elimination order mentions an unknown vertex");
    std::vector<int> nb = adj.members(v);
    width = std::max(width, (int)nb.size());
    for (int x : nb) {
      for (int y : nb)
        if (x != y) adj.set(x, y);
      adj.clear(x, v);
    }
    std::fill(adj.row(v), adj.row(v) + adj.w, 0);
  }
  return width;
}

static void js_list(std::ostringstream& o, const std::vector<int>& v) {
  o << "[";
  for (size_t i = 0; i < v.size(); ++i) o << (i ? "," : "") << v[i];
  o << "]";
}
static void js_lists(std::ostringstream& o, const std::vector<std::vector<int>>& v) {
  o << "[";
  for (size_t i = 0; i < v.size(); ++i) {
    o << (i ? "," : "");
    js_list(o, v[i]);
  }
  o << "]";
}

std::string describe_tree(const Tree& t) {
  std::ostringstream o;
  o << "{\"elim_order\":";
  js_list(o, t.elim_order);
  o << ",\"clusters\":";
  js_lists(o, t.clusters);
  o << ",\"root\":" << t.root << ",\"parent_sep\":";
  js_list(o, t.parent_sep);
  o << ",\"sep_parent\":";
  js_list(o, t.sep_parent);
  o << ",\"sep_child\":";
  js_list(o, t.sep_child);
  o << ",\"sep_vars\":";
  js_lists(o, t.sep_vars);
  o << ",\"children\":";
  js_lists(o, t.children);
  o << ",\"factors\":";
  js_lists(o, t.factors);
  o << "}";
  return o.str();
}

}  // namespace hb
