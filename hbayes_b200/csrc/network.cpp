// network.cpp -- network construction and the Hugin `.net` loader of libhbayes_cuda.
//
// Mirrors what the reference's model layer hands to the inference layer:
//   * a CPT per vertex with scope [child, parents in `ingoing` order] and values laid out with the
//     first variable slowest (Bayes/Network.hs:69-79, Bayes/Factor/PrivateCPT.hs:315-320);
//   * importBayesianGraph (Bayes/ImportExport/HuginNet.hs:160-195): vertex ids follow the order of the
//     `node` sections, `potential ( X | P1 P2 )` gives scope [X,P1,P2] and its `data` is stored in the
//     file as [P1,P2,X] (child fastest) and re-laid out child-slowest (HuginNet.hs:139-147,
//     Bayes/Factor/CPT.hs:44-51).
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <sstream>

#include "model.hpp"
#include "procedural.hpp"

namespace hb {

static int64_t scope_entries(const Network& net, const std::vector<int>& sc) {
  int64_t p = 1;
  for (int v : sc) p *= net.dims[v];
  return p;
}

static void check_network(const Network& net) {
  const int n = net.n();
  for (int v = 0; v < n; ++v) {
    if (net.dims[v] < 1 || net.dims[v] > 127)
      throw Error(HB_E_ARG, "vertex " + std::to_string(v) + ": dimension must be in 1..127");
    const auto& sc = net.scope[v];
    if (sc.empty() || sc[0] != v) throw Error(HB_E_ARG, "CPT " + std::to_string(v) + ": scope must start with its own vertex");
    for (size_t i = 0; i < sc.size(); ++i) {
      if (sc[i] < 0 || sc[i] >= n) throw Error(HB_E_ARG, "CPT " + std::to_string(v) + ": unknown vertex in scope");
      for (size_t j = 0; j < i; ++j)
        if (sc[i] == sc[j]) throw Error(HB_E_ARG, "CPT " + std::to_string(v) + ": repeated vertex in scope");
    }
    if (net.procedural(v) ? !net.values[v].empty() : (int64_t)net.values[v].size() != scope_entries(net, sc))
      throw Error(HB_E_ARG, "CPT " + std::to_string(v) + ": value count does not match its scope");
  }
}

double Network::cpt_value(int v, int64_t idx) const {
  if (!procedural(v)) return values[v][(size_t)idx];
  return proc_cpt_value(proc_seed[v], dims[v], (uint64_t)(cpt_entries(v) / dims[v]), (uint64_t)idx);
}

Network* network_create(int n_vars, const int32_t* dims, const int64_t* cpt_off, const int32_t* cpt_scope,
                        const int64_t* val_off, const double* vals, const uint64_t* proc_seed) {
  if (n_vars <= 0 || !dims || !cpt_off || !cpt_scope || !val_off || !vals) throw Error(HB_E_ARG, "hb_net_create: null or empty argument");
  auto net = std::make_unique<Network>();
  net->dims.assign(dims, dims + n_vars);
  net->scope.resize(n_vars);
  net->values.resize(n_vars);
  if (proc_seed) net->proc_seed.assign(proc_seed, proc_seed + n_vars);
  for (int v = 0; v < n_vars; ++v) {
    if (cpt_off[v + 1] < cpt_off[v] || val_off[v + 1] < val_off[v]) throw Error(HB_E_ARG, "hb_net_create: offsets must be non-decreasing");
    net->scope[v].assign(cpt_scope + cpt_off[v], cpt_scope + cpt_off[v + 1]);
    net->values[v].assign(vals + val_off[v], vals + val_off[v + 1]);
    net->names.push_back("n" + std::to_string(v));
  }
  check_network(*net);
  return net.release();
}

// ---- Hugin text ------------------------------------------------------------------------------
namespace {

struct Cursor {
  const std::string& s;
  size_t i = 0;
  explicit Cursor(const std::string& t) : s(t) {}
  bool eof() const { return i >= s.size(); }
  void blanks() {
    while (!eof() && isspace((unsigned char)s[i])) ++i;
  }
  std::string line() {  // up to (not including) the next LF; the LF is consumed
    size_t e = s.find('\n', i);
    if (e == std::string::npos) throw Error(HB_E_IO, "hugin: header line without line feed");
    std::string r = s.substr(i, e - i);
    i = e + 1;
    return r;
  }
  std::string body() {  // "{\n" ... "}"
    if (s.compare(i, 2, "{\n") != 0) throw Error(HB_E_IO, "hugin: expected '{' on its own line");
    i += 2;
    size_t e = s.find('}', i);
    if (e == std::string::npos) throw Error(HB_E_IO, "hugin: unterminated section");
    std::string r = s.substr(i, e - i);
    i = e + 1;
    return r;
  }
};

std::vector<std::string> fields(const std::string& s, const char* delims) {  // split, dropping delimiters and blanks
  std::vector<std::string> out;
  size_t i = 0;
  while (i < s.size()) {
    while (i < s.size() && strchr(delims, s[i])) ++i;
    size_t b = i;
    while (i < s.size() && !strchr(delims, s[i])) ++i;
    if (i > b) out.push_back(s.substr(b, i - b));
  }
  return out;
}

bool starts_with(const std::string& s, const char* p) { return s.compare(0, strlen(p), p) == 0; }

}  // namespace

Network* network_load_hugin(const char* path) {
  if (!path) throw Error(HB_E_ARG, "hb_net_load_hugin: null path");
  FILE* f = fopen(path, "rb");
  if (!f) throw Error(HB_E_IO, std::string("cannot open ") + path);
  std::string raw;
  char buf[1 << 16];
  size_t got;
  while ((got = fread(buf, 1, sizeof buf, f)) > 0) raw.append(buf, got);
  fclose(f);
  // filterComment (HuginNet.hs:180-186): '%' to end of line is dropped, the LF stays.
  std::string text;
  text.reserve(raw.size());
  for (size_t i = 0; i < raw.size(); ++i) {
    if (raw[i] == '%') {
      while (i < raw.size() && raw[i] != '\n') ++i;
      if (i < raw.size()) text.push_back('\n');
    } else
      text.push_back(raw[i]);
  }

  struct Pot {
    std::vector<std::string> vars;  // child first
    std::vector<double> data;       // file layout: parents slowest, child fastest
  };
  std::vector<std::string> node_names;
  std::vector<int> node_dims;
  std::vector<Pot> pots;
  Cursor c(text);
  for (c.blanks(); !c.eof(); c.blanks()) {
    std::string head = c.line();
    std::string body = c.body();
    if (head == "net") continue;
    if (starts_with(head, "node")) {
      auto nm = fields(head.substr(4), " \t");
      if (nm.size() != 1) throw Error(HB_E_IO, "hugin: bad node header '" + head + "'");
      // the only recognised command is  states = ( "a" "b" ... );  (HuginNet.hs:54-69)
      int nstates = 0;
      std::istringstream lines(body);
      std::string ln;
      while (std::getline(lines, ln)) {
        size_t b = ln.find_first_not_of(" \t");
        if (b == std::string::npos || ln.compare(b, 6, "states") != 0) continue;
        size_t q = 0;
        for (char ch : ln) q += (ch == '"');
        nstates += (int)(q / 2);
      }
      if (nstates == 0) throw Error(HB_E_IO, "hugin: node '" + nm[0] + "' has no states");
      node_names.push_back(nm[0]);
      node_dims.push_back(nstates);
    } else if (starts_with(head, "potential")) {
      Pot p;
      p.vars = fields(head.substr(9), "() |");  // splitCPT
      size_t d = body.find("data");
      size_t eq = d == std::string::npos ? d : body.find('=', d);
      if (p.vars.empty() || eq == std::string::npos) throw Error(HB_E_IO, "hugin: potential without variables or data");
      size_t semi = body.find(';', eq);
      std::string nums = body.substr(eq + 1, semi == std::string::npos ? std::string::npos : semi - eq - 1);
      for (const std::string& tok : fields(nums, "() \n\t")) {  // splitValues
        char* end = nullptr;
        double x = strtod(tok.c_str(), &end);
        if (end == tok.c_str() || *end) throw Error(HB_E_IO, "hugin: bad number '" + tok + "'");
        p.data.push_back(x);
      }
      pots.push_back(std::move(p));
    } else
      throw Error(HB_E_IO, "hugin: unknown section '" + head + "'");
  }
  if (node_names.empty()) throw Error(HB_E_IO, "hugin: no node sections");

  auto net = std::make_unique<Network>();
  net->names = node_names;
  net->dims = node_dims;
  const int n = net->n();
  net->scope.assign(n, {});
  net->values.assign(n, {});
  std::map<std::string, int> id;
  for (int v = 0; v < n; ++v) id[node_names[v]] = v;
  for (const Pot& p : pots) {
    std::vector<int> sc;
    for (const std::string& s : p.vars) {
      auto it = id.find(s);
      if (it == id.end()) throw Error(HB_E_IO, "hugin: potential mentions unknown node '" + s + "'");
      sc.push_back(it->second);
    }
    const int child = sc[0], k = (int)sc.size();
    int64_t total = 1;
    for (int v : sc) total *= net->dims[v];
    if ((int64_t)p.data.size() != total) throw Error(HB_E_IO, "hugin: wrong number of values for '" + p.vars[0] + "'");
    // file strides of [P1..Pm, X] (last fastest), seen in scope order [X, P1..Pm]
    std::vector<int64_t> fstride(k);
    fstride[0] = 1;
    int64_t acc = net->dims[child];
    for (int j = k - 1; j >= 1; --j) fstride[j] = acc, acc *= net->dims[sc[j]];
    std::vector<double> vals((size_t)total);
    std::vector<int> digit(k, 0);
    int64_t src = 0;
    for (int64_t e = 0; e < total; ++e) {
      vals[e] = p.data[src];
      for (int j = k - 1; j >= 0; --j) {
        if (++digit[j] < net->dims[sc[j]]) {
          src += fstride[j];
          break;
        }
        src -= fstride[j] * (digit[j] - 1);
        digit[j] = 0;
      }
    }
    net->scope[child] = sc;
    net->values[child] = std::move(vals);
  }
  for (int v = 0; v < n; ++v)
    if (net->scope[v].empty()) throw Error(HB_E_IO, "hugin: node '" + node_names[v] + "' has no potential");
  check_network(*net);
  return net.release();
}

std::string describe_network(const Network& net) {
  std::ostringstream o;
  o.precision(17);
  o << "{\"names\":[";
  for (int v = 0; v < net.n(); ++v) o << (v ? "," : "") << "\"" << net.names[v] << "\"";
  o << "],\"dims\":[";
  for (int v = 0; v < net.n(); ++v) o << (v ? "," : "") << net.dims[v];
  o << "],\"scopes\":[";
  for (int v = 0; v < net.n(); ++v) {
    o << (v ? "," : "") << "[";
    for (size_t i = 0; i < net.scope[v].size(); ++i) o << (i ? "," : "") << net.scope[v][i];
    o << "]";
  }
  o << "],\"values\":[";
  for (int v = 0; v < net.n(); ++v) {
    o << (v ? "," : "") << "[";
    for (size_t i = 0; i < net.values[v].size(); ++i) {  // procedural CPTs print as []
      const double x = net.values[v][i];
      o << (i ? "," : "");
      if (x != x)
        o << "NaN";  // what Python's json module reads (and writes) for the non-finite doubles
      else if (x - x != 0.0)
        o << (x > 0 ? "Infinity" : "-Infinity");
      else
        o << x;
    }
    o << "]";
  }
  o << "]}";
  return o.str();
}

}  // namespace hb
