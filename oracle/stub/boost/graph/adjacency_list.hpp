// Minimal stand-in for the slice of Boost.Graph that the reference's assembler uses
// (assembler/graph_wrapper.hpp:38-61,77-129,152-198,302-309).  Boost is not installed in this image and
// there is no network; this file is written for this repo (TEST INFRASTRUCTURE, oracle side only) so
// that the reference's WHOLE driver (haplotypecaller.hpp: do_work) can be compiled around either
// likelihood engine for the chrM-style end-to-end comparison.  It is not a
// copy of Boost: only the interface the reference calls, with the observable behaviour it relies on:
//   adjacency_list<vecS, vecS, bidirectionalS, VP, EP>   vertices 0..n-1, out/in edge lists in
//                                                        insertion order, parallel edges allowed
//   add_vertex, add_edge, source, target, out_edges, in_edges, out_degree, in_degree, edge, edges,
//   vertices, make_iterator_range, g[v], g[e]
//   dfs_visitor<>, visitor(vis), depth_first_search(filtered_graph, ...) calling back_edge on an edge
//   to a vertex that is on the current DFS stack (grey), filtered_graph<G, EdgePredicate>
#pragma once
#include <cstddef>
#include <deque>
#include <utility>
#include <vector>

namespace boost
{

struct vecS {};
struct bidirectionalS {};

template <class It> struct iterator_range_shim {
    It b, e;
    It begin() const { return b; }
    It end() const { return e; }
};
template <class It> iterator_range_shim<It> make_iterator_range(std::pair<It, It> p) { return {p.first, p.second}; }

struct edge_desc {
    std::size_t src = 0, dst = 0, idx = 0;
    friend bool operator==(const edge_desc& a, const edge_desc& b) { return a.idx == b.idx; }
    friend bool operator!=(const edge_desc& a, const edge_desc& b) { return a.idx != b.idx; }
    friend bool operator<(const edge_desc& a, const edge_desc& b) { return a.idx < b.idx; }
};

template <class OutS, class VertS, class DirS, class VP, class EP>
class adjacency_list
{
public:
    using vertex_descriptor = std::size_t;
    using edge_descriptor = edge_desc;
    using vertex_property_type = VP;
    using edge_property_type = EP;

    std::deque<VP> vprop;
    std::deque<EP> eprop;
    std::vector<std::vector<edge_desc>> out_, in_;
    std::vector<edge_desc> all_edges;
    std::vector<std::size_t> all_vertices;

    VP& operator[](vertex_descriptor v) { return vprop[v]; }
    const VP& operator[](vertex_descriptor v) const { return vprop[v]; }
    EP& operator[](const edge_descriptor& e) { return eprop[e.idx]; }
    const EP& operator[](const edge_descriptor& e) const { return eprop[e.idx]; }
};

template <class G> struct graph_traits {
    using vertex_descriptor = typename G::vertex_descriptor;
    using edge_descriptor = typename G::edge_descriptor;
};

#define BGS_TPL template <class O, class V, class D, class VP, class EP>
#define BGS_G adjacency_list<O, V, D, VP, EP>

BGS_TPL std::size_t add_vertex(BGS_G& g)
{
    g.vprop.emplace_back(); g.out_.emplace_back(); g.in_.emplace_back();
    g.all_vertices.push_back(g.vprop.size() - 1);
    return g.vprop.size() - 1;
}
BGS_TPL std::pair<edge_desc, bool> add_edge(std::size_t u, std::size_t v, BGS_G& g)
{
    edge_desc e{u, v, g.eprop.size()};
    g.eprop.emplace_back();
    g.out_[u].push_back(e); g.in_[v].push_back(e); g.all_edges.push_back(e);
    return {e, true};
}
BGS_TPL std::size_t source(const edge_desc& e, const BGS_G&) { return e.src; }
BGS_TPL std::size_t target(const edge_desc& e, const BGS_G&) { return e.dst; }
BGS_TPL auto out_edges(std::size_t u, const BGS_G& g) { return std::make_pair(g.out_[u].begin(), g.out_[u].end()); }
BGS_TPL auto in_edges(std::size_t v, const BGS_G& g) { return std::make_pair(g.in_[v].begin(), g.in_[v].end()); }
BGS_TPL std::size_t out_degree(std::size_t u, const BGS_G& g) { return g.out_[u].size(); }
BGS_TPL std::size_t in_degree(std::size_t v, const BGS_G& g) { return g.in_[v].size(); }
BGS_TPL std::pair<edge_desc, bool> edge(std::size_t u, std::size_t v, const BGS_G& g)
{
    for (const auto& e : g.out_[u]) if (e.dst == v) return {e, true};
    return {edge_desc{}, false};
}
BGS_TPL auto edges(const BGS_G& g) { return std::make_pair(g.all_edges.begin(), g.all_edges.end()); }
BGS_TPL auto vertices(const BGS_G& g) { return std::make_pair(g.all_vertices.begin(), g.all_vertices.end()); }

#undef BGS_TPL
#undef BGS_G

// ---- depth_first_search.hpp -----------------------------------------------------------------
struct null_visitor {};
template <class V = null_visitor> struct dfs_visitor {
    template <class E, class G> void back_edge(E, G&) {}
};
template <class Vis> struct visitor_param { Vis& vis; };
template <class Vis> visitor_param<Vis> visitor(Vis& v) { return {v}; }

// ---- filtered_graph.hpp ---------------------------------------------------------------------
template <class G, class EdgePred> struct filtered_graph {
    const G& g; EdgePred pred;
    filtered_graph(const G& g_, EdgePred p) : g(g_), pred(p) {}
};

// DFS over every vertex (white/grey/black); an edge to a grey vertex is a back edge.
template <class G, class EdgePred, class Vis>
void depth_first_search(const filtered_graph<G, EdgePred>& fg, visitor_param<Vis> vp)
{
    const G& g = fg.g;
    const std::size_t n = g.vprop.size();
    std::vector<char> color(n, 0);
    struct Frame { std::size_t v, next; };
    std::vector<Frame> stack;
    for (std::size_t s = 0; s < n; s++) {
        if (color[s]) continue;
        color[s] = 1; stack.push_back({s, 0});
        while (!stack.empty()) {
            Frame& f = stack.back();
            if (f.next < g.out_[f.v].size()) {
                const auto e = g.out_[f.v][f.next++];
                if (!fg.pred(e)) continue;
                if (color[e.dst] == 0) { color[e.dst] = 1; stack.push_back({e.dst, 0}); }
                else if (color[e.dst] == 1) vp.vis.back_edge(e, fg);
            } else { color[f.v] = 2; stack.pop_back(); }
        }
    }
}

} // namespace boost
