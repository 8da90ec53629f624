// Host-side quorum-descriptor builder: the step BEFORE the tally.
// Restates, for descriptor construction only,
//   node/graph/graph.go:46-75     AddNodes (an edge signer -> signee per certification)
//   node/graph/graph.go:77-88     SetSelfNodes
//   node/graph/graph.go:90-108    RemoveNodes,  :131-140 Revoke
//   node/graph/graph.go:117-125   GetPeers
//   node/graph/graph.go:279-393   GetReachableNodes, GetCliques, findMaximalClique, bidirect, putWeight
//   node/graph/graph.go:420-438   bfs
//   quorum/wotqs/wotqs.go:36-127  newQC, complement, getQuorumFrom, ChooseQuorum
// The reference recomputes this on every call (client.go:64,101,141,238; server.go:182,211,237,300,473);
// graphs are tiny (<= ~30 vertices) and pointer-chasing, so this stays on the CPU and its output (a
// bftq_qc_ids_t list) is what the GPU tally consumes.  Go iterates maps in random order; this code
// iterates in insertion order, one of the legal orders (order-independent for disjoint cliques).
#pragma once
#include <cstdint>
#include <unordered_map>
#include <vector>

namespace bftq { namespace wot {

enum : int { READ = 0x01, WRITE = 0x02, AUTH = 0x04, CERT = 0x08, PEER = 0x10 };   // quorum/quorum.go:10-16

struct Vertex {
  uint64_t id = 0;
  bool has_instance = false;
  std::vector<uint64_t> edges;             // out-edges in insertion order (ids of signees)
  bool has_edge(uint64_t to) const { for (uint64_t e : edges) if (e == to) return true; return false; }
};

struct QC { std::vector<uint64_t> nodes; int f = 0, min = 0, threshold = 0, suff = 0; };

class Graph {
 public:
  // graph.AddNodes for one node: `signers` = the issuer ids of the certifications on it.
  void add_node(uint64_t id, const uint64_t* signers, uint32_t n) {
    if (revoked(id)) return;
    Vertex& me = vertex(id);
    me.has_instance = true;
    for (uint32_t i = 0; i < n; i++) {
      if (revoked(signers[i])) continue;
      Vertex& v = vertex(signers[i]);
      if (!v.has_edge(id)) v.edges.push_back(id);
    }
  }
  void set_self(uint64_t id) {
    int i = find(id);
    if (i < 0 || !order_[i].has_instance) add_node(id, nullptr, 0);
    self_.push_back(id);
  }
  void remove_node(uint64_t id) {
    for (auto& v : order_) {
      for (size_t k = 0; k < v.edges.size(); k++) if (v.edges[k] == id) { v.edges.erase(v.edges.begin() + k); break; }
    }
    int i = find(id);
    if (i >= 0) { order_.erase(order_.begin() + i); reindex(); }
    for (size_t k = 0; k < self_.size(); k++) if (self_[k] == id) { self_.erase(self_.begin() + k); break; }
  }
  void revoke(uint64_t id) {
    int i = find(id);
    if (i >= 0 && order_[i].has_instance) remove_node(id);
    revoked_.push_back(id);
  }
  uint64_t self_id() const {
    if (self_.empty()) return 0;
    int i = find(self_[0]);
    return (i >= 0 && order_[i].has_instance) ? self_[0] : 0;
  }

  void choose_quorum(int rw, std::vector<QC>& out) const {          // wotqs.go:117-127
    const int distance = (rw & CERT) ? 0 : ((rw & AUTH) ? 1 : 2);
    quorum_from(rw, self_id(), distance, out);
  }

 private:
  std::vector<Vertex> order_;
  std::unordered_map<uint64_t, int> index_;
  std::vector<uint64_t> self_, revoked_;

  bool revoked(uint64_t id) const { for (uint64_t r : revoked_) if (r == id) return true; return false; }
  int find(uint64_t id) const { auto it = index_.find(id); return it == index_.end() ? -1 : it->second; }
  void reindex() { index_.clear(); for (size_t i = 0; i < order_.size(); i++) index_[order_[i].id] = (int)i; }
  Vertex& vertex(uint64_t id) {
    int i = find(id);
    if (i >= 0) return order_[i];
    Vertex v; v.id = id;
    order_.push_back(v);
    index_[id] = (int)order_.size() - 1;
    return order_.back();
  }

  struct VD { int v; int d; };
  template <typename F> void bfs(int start, F proc) const {          // graph.go:420-438
    std::vector<uint64_t> seen{order_[start].id};
    std::vector<VD> q{{start, 0}};
    size_t head = 0;
    while (head < q.size()) {
      const VD vd = q[head++];
      if (proc(vd)) return;
      for (uint64_t to : order_[vd.v].edges) {
        bool s = false;
        for (uint64_t x : seen) s = s || x == to;
        if (s) continue;
        const int ti = find(to);
        if (ti < 0) continue;                  // edge to a vertex that no longer exists
        q.push_back({ti, vd.d + 1});
        seen.push_back(to);
      }
    }
  }
  bool bidirect(int v, const std::vector<int>& clique) const {       // graph.go:370-380
    for (int c : clique) {
      if (!order_[c].has_edge(order_[v].id)) return false;
      if (!order_[v].has_edge(order_[c].id)) return false;
    }
    return true;
  }
  bool find_maximal_clique(int s, std::vector<uint64_t>& nodes) const {   // graph.go:333-368
    std::vector<int> clique{s};
    for (int v = 0; v < (int)order_.size(); v++) {
      if (!order_[v].has_instance || v == s) continue;
      if (bidirect(v, clique)) clique.push_back(v);
    }
    for (int v = 0; v < (int)order_.size(); v++) {
      if (!order_[v].has_instance || v == s) continue;
      bool in = false;
      for (int c : clique) in = in || c == v;
      if (!in && bidirect(v, std::vector<int>{s})) return false;    // "found more than one maximal cliques"
    }
    nodes.clear();
    for (int c : clique) nodes.push_back(order_[c].id);
    return true;
  }
  struct Clique { std::vector<uint64_t> nodes; int weight = 0; };
  void get_cliques(uint64_t sid, int distance, std::vector<Clique>& cliques) const {   // graph.go:297-319
    const int s = find(sid);
    if (s < 0 || !order_[s].has_instance) return;
    bfs(s, [&](const VD& vd) {
      if (distance >= 0 && vd.d > distance) return true;
      if (order_[vd.v].has_instance) {
        bool in = false;
        for (auto& c : cliques) for (uint64_t n : c.nodes) in = in || n == order_[vd.v].id;
        if (!in) {
          Clique c;
          if (find_maximal_clique(vd.v, c.nodes)) {
            for (uint64_t e : order_[s].edges) for (uint64_t n : c.nodes) if (n == e) c.weight++;   // putWeight
            cliques.push_back(c);
          }
        }
      }
      return false;
    });
  }
  void reachable(uint64_t sid, int distance, std::vector<uint64_t>& nodes) const {     // graph.go:279-295
    const int s = find(sid);
    if (s < 0) return;
    bfs(s, [&](const VD& vd) {
      if (distance >= 0 && vd.d > distance) return true;
      if (order_[vd.v].has_instance) nodes.push_back(order_[vd.v].id);
      return false;
    });
  }
  void peers(std::vector<uint64_t>& nodes) const {                                      // graph.go:117-125
    const uint64_t me = self_id();
    for (auto& v : order_) if (v.has_instance && v.id != me) nodes.push_back(v.id);
  }
  bool new_qc(const std::vector<uint64_t>& clique_nodes, int weight, int rw, QC& qc) const {   // wotqs.go:36-70
    qc = QC();
    const uint64_t me = self_id();
    for (uint64_t n : clique_nodes) if (!(rw & PEER) || n != me) qc.nodes.push_back(n);
    const int n = (int)qc.nodes.size();
    if (n == 0) return false;
    if (rw == WRITE) return true;                    // all-zero thresholds
    const int f = (n - 1) / 3;
    if (f < 1) return false;
    qc.f = f; qc.min = 3 * f + 1; qc.threshold = 2 * f + 1; qc.suff = f + (n - f) / 2 + 1;
    if (rw & (CERT | READ)) qc.threshold = f + 1;
    if (weight <= n - qc.suff) qc.suff = 0;
    return true;
  }
  void complement(const std::vector<uint64_t>& u, const std::vector<QC>& c, std::vector<QC>& e, int rw) const {   // wotqs.go:72-93
    std::vector<uint64_t> nodes;
    for (uint64_t n1 : u) {
      bool found = false;
      for (auto& qc : c) for (uint64_t n2 : qc.nodes) found = found || n1 == n2;
      if (!found) nodes.push_back(n1);
    }
    QC q;
    if (new_qc(nodes, 0, rw, q)) e.push_back(q);
  }
  void quorum_from(int rw, uint64_t s, int distance, std::vector<QC>& out) const {     // wotqs.go:95-115
    std::vector<Clique> cliques;
    get_cliques(s, distance, cliques);
    std::vector<QC> qcs;
    for (auto& c : cliques) { QC q; if (new_qc(c.nodes, c.weight, rw | AUTH, q)) qcs.push_back(q); }
    if (rw & (READ | WRITE)) {
      std::vector<QC> res;
      if (rw & AUTH) res = qcs;
      std::vector<uint64_t> r;
      reachable(s, distance, r);
      complement(r, qcs, res, READ);
      if (rw & WRITE) {
        std::vector<QC> both = qcs;
        both.insert(both.end(), res.begin(), res.end());
        std::vector<uint64_t> p;
        peers(p);
        complement(p, both, res, WRITE);
      }
      out = res;
    } else {
      out = qcs;
    }
  }
};

}}  // namespace bftq::wot
