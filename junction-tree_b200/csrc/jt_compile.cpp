// Host compile phase of libjt_b200 in C++ (no CUDA calls): triangulation, maximal cliques,
// junction tree and the emission of the level-ordered message schedule (the plan blob).
//
// What it replaces in the reference (paths relative to the reference checkout):
//   jt_triangulate    find_triangulation        junctiontree/construction.py:176-353
//   jt_junction_tree  construct_junction_tree   junctiontree/construction.py:522-601
//   jt_plan_build     everything the reference recomputes per call in Python: per-edge einsum
//                     subscripts and orders (computation.py:47-96, 140-224), clique <- factor maps
//                     (junctiontree.py:203-226), marginalisation subscripts (:229-274), evidence
//                     slicing (computation.py:11-34) -- here compiled once into index tables.
//
// The algorithms are those of junctiontree/construction.py and junctiontree/schedule.py of this
// package (not the reference's, whose construction is invalid on most inputs -- SURVEY.md
// section 9).  The Python implementations remain as the cross-check: for the same input this
// file produces the same cliques, the same tree and a byte-identical plan blob
// (tests/test_native_compile.py).
//
// Variables are integers 0..n_vars-1; the integer is also the tie-break rank (the Python side
// numbers the labels in sorted order).

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <new>
#include <queue>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "jt_host.h"

namespace {

typedef unsigned __int128 u128;

// products of sizes as exact integers (Python ints on the other side); saturates at 2^127
u128 sat_mul(u128 a, u128 b) {
    const u128 cap = (u128)1 << 127;
    if (a == 0 || b == 0) return 0;
    if (a >= cap / b) return cap;
    return a * b;
}

u128 sat_add(u128 a, u128 b) {
    const u128 cap = (u128)1 << 127;
    return (a >= cap || b >= cap || a + b >= cap) ? cap : a + b;
}

int bad(const char* what) { return jt_fail(JT_ERR_INVALID, "host compile: %s", what); }

bool csr_ok(int32_t n, const int32_t* ptr, const int32_t* data, int32_t limit) {
    if (n < 0 || (n > 0 && (!ptr || ptr[0] != 0))) return false;
    for (int32_t i = 0; i < n; ++i) {
        if (ptr[i + 1] < ptr[i]) return false;
        for (int32_t k = ptr[i]; k < ptr[i + 1]; ++k)
            if (!data || data[k] < 0 || data[k] >= limit) return false;
    }
    return true;
}

}  // namespace

struct jt_ibuf {
    std::vector<std::vector<int32_t>> arrays;
};

namespace {

// ------------------------------------------------------------------------------------------
// triangulation: min-fill on the current graph, ties by cluster weight, then by rank
// (junctiontree/construction.py: elimination_clusters, find_triangulation)

struct Graph {
    std::vector<std::vector<int32_t>> adj;   // sorted neighbour lists
    std::vector<char> alive;

    bool has(int32_t a, int32_t b) const { return std::binary_search(adj[a].begin(), adj[a].end(), b); }
    void add(int32_t a, int32_t b) {
        auto& v = adj[a];
        auto it = std::lower_bound(v.begin(), v.end(), b);
        if (it == v.end() || *it != b) v.insert(it, b);
    }
    void remove(int32_t a, int32_t b) {
        auto& v = adj[a];
        auto it = std::lower_bound(v.begin(), v.end(), b);
        if (it != v.end() && *it == b) v.erase(it);
    }
};

void fill_and_weight(const Graph& g, const int64_t* sizes, int32_t var, int64_t& fill, u128& weight) {
    const auto& nb = g.adj[var];
    fill = 0;
    for (size_t i = 0; i < nb.size(); ++i)
        for (size_t j = i + 1; j < nb.size(); ++j)
            if (!g.has(nb[i], nb[j])) ++fill;
    weight = (u128)sizes[var];
    for (int32_t n : nb) weight = sat_mul(weight, (u128)sizes[n]);
}

struct HeapEntry {
    int64_t fill;
    u128 weight;
    int32_t var;
    int32_t version;
    bool operator>(const HeapEntry& o) const {
        if (fill != o.fill) return fill > o.fill;
        if (weight != o.weight) return weight > o.weight;
        if (var != o.var) return var > o.var;
        return version > o.version;
    }
};

}  // namespace

extern "C" {

int jt_ibuf_count(const jt_ibuf* b) { return b ? (int)b->arrays.size() : 0; }

int64_t jt_ibuf_size(const jt_ibuf* b, int k) {
    return (b && k >= 0 && k < (int)b->arrays.size()) ? (int64_t)b->arrays[k].size() : -1;
}

const int32_t* jt_ibuf_data(const jt_ibuf* b, int k) {
    return (b && k >= 0 && k < (int)b->arrays.size()) ? b->arrays[k].data() : nullptr;
}

void jt_ibuf_destroy(jt_ibuf* b) { delete b; }

void jt_free(void* p) { free(p); }

int jt_triangulate(int32_t n_vars, const int64_t* var_sizes, int32_t n_factors, const int32_t* factor_ptr,
                   const int32_t* factor_vars, const int32_t* order, int32_t n_order, jt_ibuf** out) {
    if (!out) return bad("null output");
    *out = nullptr;
    if (n_vars < 0 || (n_vars > 0 && !var_sizes)) return bad("variable sizes");
    if (!csr_ok(n_factors, factor_ptr, factor_vars, n_vars)) return bad("factor lists");
    for (int32_t v = 0; v < n_vars; ++v)
        if (var_sizes[v] <= 0) return bad("variable size must be positive");

    // variables that occur in a factor take part; the others are ignored
    std::vector<char> used(n_vars, 0);
    for (int32_t f = 0; f < n_factors; ++f)
        for (int32_t k = factor_ptr[f]; k < factor_ptr[f + 1]; ++k) used[factor_vars[k]] = 1;
    int32_t n_used = 0;
    for (int32_t v = 0; v < n_vars; ++v) n_used += used[v];

    Graph g;
    g.adj.resize(n_vars);
    g.alive.assign(used.begin(), used.end());
    for (int32_t f = 0; f < n_factors; ++f)
        for (int32_t i = factor_ptr[f]; i < factor_ptr[f + 1]; ++i)
            for (int32_t j = i + 1; j < factor_ptr[f + 1]; ++j)
                if (factor_vars[i] != factor_vars[j]) {
                    g.add(factor_vars[i], factor_vars[j]);
                    g.add(factor_vars[j], factor_vars[i]);
                }

    std::vector<int32_t> elim_order, fill_edges;
    std::vector<std::vector<int32_t>> clusters;
    auto eliminate = [&](int32_t var) {
        const std::vector<int32_t> nb = g.adj[var];            // sorted by rank
        for (size_t i = 0; i < nb.size(); ++i)
            for (size_t j = i + 1; j < nb.size(); ++j)
                if (!g.has(nb[i], nb[j])) {
                    g.add(nb[i], nb[j]);
                    g.add(nb[j], nb[i]);
                    fill_edges.push_back(nb[i]);
                    fill_edges.push_back(nb[j]);
                }
        for (int32_t n : nb) g.remove(n, var);
        g.adj[var].clear();
        g.alive[var] = 0;
        elim_order.push_back(var);
        std::vector<int32_t> cluster;
        cluster.push_back(var);
        cluster.insert(cluster.end(), nb.begin(), nb.end());
        clusters.push_back(std::move(cluster));
        return nb;
    };

    if (order) {
        if (n_order != n_used) return bad("order must be a permutation of the variables used by the factors");
        std::vector<char> seen(n_vars, 0);
        for (int32_t i = 0; i < n_order; ++i) {
            const int32_t v = order[i];
            if (v < 0 || v >= n_vars || !used[v] || seen[v])
                return bad("order must be a permutation of the variables used by the factors");
            seen[v] = 1;
        }
        for (int32_t i = 0; i < n_order; ++i) eliminate(order[i]);
    } else {
        std::vector<int32_t> version(n_vars, 0);
        std::priority_queue<HeapEntry, std::vector<HeapEntry>, std::greater<HeapEntry>> heap;
        for (int32_t v = 0; v < n_vars; ++v) {
            if (!used[v]) continue;
            HeapEntry e;
            fill_and_weight(g, var_sizes, v, e.fill, e.weight);
            e.var = v;
            e.version = 0;
            heap.push(e);
        }
        std::vector<char> mark(n_vars, 0);
        std::vector<int32_t> touched;
        int32_t remaining = n_used;
        while (remaining > 0) {
            const HeapEntry top = heap.top();
            heap.pop();
            if (!g.alive[top.var] || top.version != version[top.var]) continue;   // stale entry
            const std::vector<int32_t> nb = eliminate(top.var);
            --remaining;
            // scores can only change for the neighbours and for their neighbours
            touched.clear();
            for (int32_t n : nb)
                if (!mark[n]) { mark[n] = 1; touched.push_back(n); }
            for (int32_t n : nb)
                for (int32_t m : g.adj[n])
                    if (!mark[m]) { mark[m] = 1; touched.push_back(m); }
            for (int32_t t : touched) {
                mark[t] = 0;
                HeapEntry e;
                fill_and_weight(g, var_sizes, t, e.fill, e.weight);
                e.var = t;
                e.version = ++version[t];
                heap.push(e);
            }
        }
    }

    // maximal cliques: a cluster is dropped iff an earlier kept clique that holds its variable
    // contains it (only those can)
    std::vector<std::vector<int32_t>> cliques;                 // sorted by rank
    std::vector<std::vector<int32_t>> cliques_of(n_vars);
    for (size_t i = 0; i < clusters.size(); ++i) {
        std::vector<int32_t> sorted_cluster = clusters[i];
        std::sort(sorted_cluster.begin(), sorted_cluster.end());
        bool contained = false;
        for (int32_t ix : cliques_of[elim_order[i]])
            if (std::includes(cliques[ix].begin(), cliques[ix].end(), sorted_cluster.begin(), sorted_cluster.end())) {
                contained = true;
                break;
            }
        if (contained) continue;
        const int32_t ix = (int32_t)cliques.size();
        for (int32_t v : sorted_cluster) cliques_of[v].push_back(ix);
        cliques.push_back(std::move(sorted_cluster));
    }
    if (cliques.empty()) cliques.push_back({});                // only scalar factors: one empty clique

    jt_ibuf* res = new (std::nothrow) jt_ibuf;
    if (!res) return jt_fail(JT_ERR_NOMEM, "out of host memory");
    res->arrays.resize(5);
    auto& cptr = res->arrays[0];
    auto& cvars = res->arrays[1];
    cptr.push_back(0);
    for (const auto& c : cliques) {
        cvars.insert(cvars.end(), c.begin(), c.end());
        cptr.push_back((int32_t)cvars.size());
    }
    auto& f2c = res->arrays[2];
    for (int32_t f = 0; f < n_factors; ++f) {
        const int32_t b = factor_ptr[f], e = factor_ptr[f + 1];
        if (b == e) {
            f2c.push_back(0);
            continue;
        }
        std::vector<int32_t> fs(factor_vars + b, factor_vars + e);
        std::sort(fs.begin(), fs.end());
        fs.erase(std::unique(fs.begin(), fs.end()), fs.end());
        int32_t home = -1;
        for (int32_t ix : cliques_of[factor_vars[b]])
            if (std::includes(cliques[ix].begin(), cliques[ix].end(), fs.begin(), fs.end())) {
                home = ix;
                break;
            }
        if (home < 0) {
            delete res;
            return bad("internal error: factor without a containing clique");
        }
        f2c.push_back(home);
    }
    res->arrays[3] = std::move(fill_edges);
    res->arrays[4] = std::move(elim_order);
    *out = res;
    return JT_OK;
}

int jt_junction_tree(int32_t n_vars, const int64_t* var_sizes, int32_t n_cliques, const int32_t* clique_ptr,
                     const int32_t* clique_vars, int32_t root, jt_ibuf** out) {
    if (!out) return bad("null output");
    *out = nullptr;
    if (n_vars < 0 || (n_vars > 0 && !var_sizes)) return bad("variable sizes");
    if (!csr_ok(n_cliques, clique_ptr, clique_vars, n_vars)) return bad("clique lists");
    if (root >= n_cliques) return bad("root is not a clique index");
    jt_ibuf* res = new (std::nothrow) jt_ibuf;
    if (!res) return jt_fail(JT_ERR_NOMEM, "out of host memory");
    res->arrays.resize(5);
    *out = res;
    const int32_t n = n_cliques;
    if (n == 0) {
        res->arrays[0].push_back(0);
        return JT_OK;
    }
    std::vector<std::vector<int32_t>> csets(n);               // sorted variable sets
    std::vector<u128> weights(n);
    std::vector<std::vector<int32_t>> members(n_vars);
    for (int32_t c = 0; c < n; ++c) {
        csets[c].assign(clique_vars + clique_ptr[c], clique_vars + clique_ptr[c + 1]);
        std::sort(csets[c].begin(), csets[c].end());
        csets[c].erase(std::unique(csets[c].begin(), csets[c].end()), csets[c].end());
        u128 w = 1;
        for (int32_t v : csets[c]) {
            w = sat_mul(w, (u128)var_sizes[v]);
            members[v].push_back(c);
        }
        weights[c] = w;
    }
    // candidate edges: more shared variables first, then lighter clique pairs, then pair index
    std::map<std::pair<int32_t, int32_t>, int32_t> shared;
    for (int32_t v = 0; v < n_vars; ++v)
        for (size_t i = 0; i < members[v].size(); ++i)
            for (size_t j = i + 1; j < members[v].size(); ++j) ++shared[{members[v][i], members[v][j]}];
    struct Cand {
        int32_t k;
        u128 w;
        int32_t a, b;
    };
    std::vector<Cand> cands;
    cands.reserve(shared.size());
    for (const auto& kv : shared)
        cands.push_back({kv.second, sat_add(weights[kv.first.first], weights[kv.first.second]), kv.first.first,
                         kv.first.second});
    std::sort(cands.begin(), cands.end(), [](const Cand& x, const Cand& y) {
        if (x.k != y.k) return x.k > y.k;
        if (x.w != y.w) return x.w < y.w;
        if (x.a != y.a) return x.a < y.a;
        return x.b < y.b;
    });
    std::vector<int32_t> uf(n);
    for (int32_t i = 0; i < n; ++i) uf[i] = i;
    std::function<int32_t(int32_t)> find = [&](int32_t x) {
        while (uf[x] != x) {
            uf[x] = uf[uf[x]];
            x = uf[x];
        }
        return x;
    };
    std::vector<std::vector<int32_t>> nbrs(n);
    int32_t n_edges = 0;
    for (const Cand& c : cands) {
        const int32_t ra = find(c.a), rb = find(c.b);
        if (ra != rb) {
            uf[ra] = rb;
            nbrs[c.a].push_back(c.b);
            nbrs[c.b].push_back(c.a);
            ++n_edges;
        }
    }
    if (n_edges < n - 1) {
        // unconnected components: join them through empty separators
        std::map<int32_t, int32_t> comps;
        for (int32_t ix = 0; ix < n; ++ix) comps.insert({find(ix), ix});
        std::vector<int32_t> reps;
        for (const auto& kv : comps) reps.push_back(kv.second);
        std::sort(reps.begin(), reps.end());
        for (size_t i = 0; i + 1 < reps.size(); ++i) {
            uf[find(reps[i])] = find(reps[i + 1]);
            nbrs[reps[i]].push_back(reps[i + 1]);
            nbrs[reps[i + 1]].push_back(reps[i]);
        }
    }
    for (auto& v : nbrs) std::sort(v.begin(), v.end());

    auto bfs = [&](int32_t src, std::vector<int32_t>& prev) {
        std::vector<char> seen(n, 0);
        prev.assign(n, -1);
        std::deque<int32_t> q;
        q.push_back(src);
        seen[src] = 1;
        int32_t last = src;
        while (!q.empty()) {
            const int32_t u = q.front();
            q.pop_front();
            last = u;
            for (int32_t w : nbrs[u])
                if (!seen[w]) {
                    seen[w] = 1;
                    prev[w] = u;
                    q.push_back(w);
                }
        }
        return last;
    };
    if (root < 0) {
        // centre of the tree: middle of a longest path; of the two centres the heavier clique
        std::vector<int32_t> prev;
        const int32_t end_a = bfs(0, prev);
        const int32_t end_b = bfs(end_a, prev);
        std::vector<int32_t> path;
        for (int32_t u = end_b; u >= 0; u = prev[u]) path.push_back(u);
        const size_t mid = (path.size() - 1) / 2;
        int32_t c1 = path[mid], c2 = path[path.size() - 1 - mid];
        if (c1 > c2) std::swap(c1, c2);
        root = weights[c2] > weights[c1] ? c2 : c1;
    }
    // orient away from the root; separators are numbered in breadth-first order
    auto& sep_ptr = res->arrays[0];
    auto& sep_vars = res->arrays[1];
    auto& parent = res->arrays[2];
    auto& parent_sep = res->arrays[3];
    auto& order = res->arrays[4];
    parent.assign(n, -1);
    parent_sep.assign(n, -1);
    sep_ptr.push_back(0);
    std::vector<char> visited(n, 0);
    visited[root] = 1;
    std::deque<int32_t> q;
    q.push_back(root);
    int32_t n_seps = 0;
    while (!q.empty()) {
        const int32_t u = q.front();
        q.pop_front();
        order.push_back(u);
        for (int32_t w : nbrs[u])
            if (!visited[w]) {
                visited[w] = 1;
                std::set_intersection(csets[u].begin(), csets[u].end(), csets[w].begin(), csets[w].end(),
                                      std::back_inserter(sep_vars));
                sep_ptr.push_back((int32_t)sep_vars.size());
                parent[w] = u;
                parent_sep[w] = n + n_seps++;
                q.push_back(w);
            }
    }
    return JT_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// schedule emission (junctiontree/schedule.py: Plan._compile and the _build_* methods)

namespace {

constexpr int64_t kLoTableMax = 1024;

struct Space {
    std::vector<int32_t> vars, hi_vars, lo_vars;
    std::vector<int64_t> shape, hi_shape, lo_shape;
    int64_t n = 1, n_lo = 1, n_hi = 1;
};

int64_t prod(const std::vector<int64_t>& xs) {
    int64_t p = 1;
    for (int64_t x : xs) p *= x;
    return p;
}

std::vector<int64_t> row_major_strides(const std::vector<int64_t>& shape) {
    std::vector<int64_t> st(shape.size(), 0);
    int64_t acc = 1;
    for (size_t i = shape.size(); i-- > 0;) {
        st[i] = acc;
        acc *= shape[i];
    }
    return st;
}

Space make_space(const std::vector<int32_t>& vars, const std::vector<int64_t>& sizes) {
    Space s;
    s.vars = vars;
    for (int32_t v : vars) s.shape.push_back(sizes[v]);
    s.n = prod(s.shape);
    size_t k = 0;
    int64_t acc = 1;
    for (size_t i = s.shape.size(); i-- > 0;) {      // trailing axes that go into the lo table
        if (k > 0 && acc * s.shape[i] > kLoTableMax) break;
        acc *= s.shape[i];
        ++k;
    }
    const size_t cut = vars.size() - k;
    s.hi_vars.assign(vars.begin(), vars.begin() + cut);
    s.lo_vars.assign(vars.begin() + cut, vars.end());
    s.hi_shape.assign(s.shape.begin(), s.shape.begin() + cut);
    s.lo_shape.assign(s.shape.begin() + cut, s.shape.end());
    s.n_lo = prod(s.lo_shape);
    s.n_hi = prod(s.hi_shape);
    return s;
}

struct Emitter {
    // inputs
    int32_t n_vars = 0, n_cliques = 0, n_seps = 0, n_nodes = 0;
    bool has_tree = false, has_factors = false;
    std::vector<std::vector<int32_t>> node_vars, factors, out_scopes;
    std::vector<int64_t> sizes, full_sizes;
    std::vector<int32_t> order, parent, parent_sep, depth, f2c, evidence_vars, out_clique, lik_vars, lik_clique;
    std::vector<int64_t> lik_off;
    int64_t lik_base = 0, lik_entries = 0;
    std::vector<std::vector<std::pair<int32_t, int32_t>>> children;   // (sep node, child clique)
    int32_t root = -1, max_depth = 0;
    // derived
    std::vector<std::vector<int64_t>> node_shape, fin_shape;
    std::vector<int64_t> node_size, node_off, fin_off, fin_size, fout_off, fout_size;
    std::vector<int64_t> ev_card, evf_ptr, evf_var, evf_stride;
    int64_t clique_entries = 0, sep_entries = 0, up_base = 0, down_base = 0, fin_entries = 0, fout_entries = 0,
            uni_entries = 0;
    std::vector<char> uniform, uniform_up, uniform_down;   // uniform_down: per separator (node id - n_cliques)
    std::vector<std::vector<int32_t>> by_depth;
    // outputs
    std::vector<int32_t> tab;
    std::unordered_map<std::string, int64_t> tab_index;
    std::vector<std::vector<int64_t>> tasks, msgs, launches;
    std::vector<int64_t> stride_of;          // scratch: var -> stride (0 = absent)
    bool overflow = false;

    int64_t tab_add(const std::vector<int64_t>& arr) {
        std::vector<int32_t> a32(arr.size());
        for (size_t i = 0; i < arr.size(); ++i) {
            if (arr[i] < 0 || arr[i] >= ((int64_t)1 << 31)) overflow = true;
            a32[i] = (int32_t)arr[i];
        }
        std::string key(reinterpret_cast<const char*>(a32.data()), a32.size() * sizeof(int32_t));
        auto hit = tab_index.find(key);
        if (hit != tab_index.end()) return hit->second;
        const int64_t off = (int64_t)tab.size();
        tab_index.emplace(std::move(key), off);
        tab.insert(tab.end(), a32.begin(), a32.end());
        return off;
    }

    // table of the additive map x -> sum_v digit_v(x) * stride_of[v] over (vars, shape)
    std::vector<int64_t> table(const std::vector<int32_t>& vars, const std::vector<int64_t>& shape) const {
        // built axis by axis (row-major: a later axis varies faster), no divisions
        std::vector<int64_t> t(1, 0), next;
        for (size_t ax = 0; ax < vars.size(); ++ax) {
            const int64_t sz = shape[ax], st = stride_of[vars[ax]];
            next.resize(t.size() * (size_t)sz);
            size_t o = 0;
            for (int64_t base : t)
                for (int64_t d = 0; d < sz; ++d) next[o++] = base + d * st;
            t.swap(next);
        }
        return t;
    }

    void set_strides(const std::vector<int32_t>& vars, const std::vector<int64_t>& strides) {
        for (size_t i = 0; i < vars.size(); ++i) stride_of[vars[i]] = strides[i];
    }
    void clear_strides(const std::vector<int32_t>& vars) {
        for (int32_t v : vars) stride_of[v] = 0;
    }
    void node_strides(int32_t node, bool set) {
        if (set) set_strides(node_vars[node], row_major_strides(node_shape[node]));
        else clear_strides(node_vars[node]);
    }

    int64_t bel_off(int32_t sep) const { return node_off[sep]; }
    int64_t up_off(int32_t sep) const { return up_base + node_off[sep] - clique_entries; }
    int64_t down_off(int32_t sep) const { return down_base + node_off[sep] - clique_entries; }

    bool touches(const Space& sp, const std::vector<int32_t>& vars) const {
        for (int32_t v : sp.vars)
            if (std::find(vars.begin(), vars.end(), v) != vars.end()) return true;
        return false;
    }

    // stride_of must be set by the caller
    void add_msg(int64_t off, const Space& s_space, const Space* r_space, int32_t fid, bool uni) {
        std::vector<int64_t> row(JT_MSG_WORDS, 0);
        row[JT_M_OFF] = off;
        row[JT_M_UNI] = uni ? 1 : 0;
        row[JT_M_AHI] = tab_add(table(s_space.hi_vars, s_space.hi_shape));
        row[JT_M_ALO] = tab_add(table(s_space.lo_vars, s_space.lo_shape));
        if (r_space) {
            row[JT_M_BHI] = tab_add(table(r_space->hi_vars, r_space->hi_shape));
            row[JT_M_BLO] = tab_add(table(r_space->lo_vars, r_space->lo_shape));
        }
        row[JT_M_FID] = fid;
        msgs.push_back(std::move(row));
    }

    std::vector<int64_t> new_task(int kind, const Space& s_space, const Space* r_space, int32_t node, int32_t src_node,
                                  bool src_is_psi) {
        std::vector<int64_t> row(JT_TASK_WORDS, 0);
        row[JT_T_KIND] = kind;
        row[JT_T_SRC] = row[JT_T_OUT] = row[JT_T_BETA] = row[JT_T_BEL] = row[JT_T_OWN] = -1;
        if (src_is_psi && src_node >= 0 && uniform[src_node]) row[JT_T_FLAGS] |= JT_TF_SRC_UNIFORM;
        row[JT_T_NS] = s_space.n;
        row[JT_T_NSLO] = s_space.n_lo;
        row[JT_T_NR] = r_space ? r_space->n : 1;
        row[JT_T_NRLO] = r_space ? r_space->n_lo : 1;
        row[JT_T_NODE] = node;
        if (src_node >= 0) {
            node_strides(src_node, true);
            row[JT_T_SRC] = node_off[src_node];
            row[JT_T_SRC_SHI] = tab_add(table(s_space.hi_vars, s_space.hi_shape));
            row[JT_T_SRC_SLO] = tab_add(table(s_space.lo_vars, s_space.lo_shape));
            row[JT_T_SRC_RHI] = tab_add(table(r_space->hi_vars, r_space->hi_shape));
            row[JT_T_SRC_RLO] = tab_add(table(r_space->lo_vars, r_space->lo_shape));
            node_strides(src_node, false);
        }
        return row;
    }

    struct Incoming {
        int64_t off;
        int32_t sep;
        bool uni;
    };

    // r-dependent messages first, then the ones that depend on s only
    void attach_msgs(std::vector<int64_t>& row, const std::vector<Incoming>& in, const Space& s_space,
                     const Space& r_space) {
        row[JT_T_RMSG_BEGIN] = (int64_t)msgs.size();
        for (const Incoming& m : in)
            if (touches(r_space, node_vars[m.sep])) {
                node_strides(m.sep, true);
                add_msg(m.off, s_space, &r_space, -1, m.uni);
                node_strides(m.sep, false);
            }
        row[JT_T_RMSG_END] = row[JT_T_SMSG_BEGIN] = (int64_t)msgs.size();
        for (const Incoming& m : in)
            if (!touches(r_space, node_vars[m.sep])) {
                node_strides(m.sep, true);
                add_msg(m.off, s_space, nullptr, -1, m.uni);
                node_strides(m.sep, false);
            }
        row[JT_T_SMSG_END] = (int64_t)msgs.size();
    }

    void launch(int phase, size_t begin, int level) {
        if (tasks.size() > begin) launches.push_back({phase, (int64_t)begin, (int64_t)tasks.size(), level});
    }
    void launch_split(int phase_all, int phase_uniform, int phase_instance, size_t begin, size_t middle, int level) {
        const size_t end = tasks.size();
        if (end > begin) launches.push_back({phase_all, (int64_t)begin, (int64_t)end, level});
        if (middle > begin) launches.push_back({phase_uniform, (int64_t)begin, (int64_t)middle, level});
        if (end > middle) launches.push_back({phase_instance, (int64_t)middle, (int64_t)end, level});
    }

    std::vector<int32_t> minus(const std::vector<int32_t>& a, const std::vector<int32_t>& b) const {
        std::vector<int32_t> r;
        for (int32_t v : a)
            if (std::find(b.begin(), b.end(), v) == b.end()) r.push_back(v);
        return r;
    }

    int prepare() {
        stride_of.assign(n_vars, 0);
        node_shape.resize(n_nodes);
        node_size.resize(n_nodes);
        node_off.resize(n_nodes);
        int64_t acc = 0;
        for (int32_t k = 0; k < n_nodes; ++k) {
            for (int32_t v : node_vars[k]) node_shape[k].push_back(sizes[v]);
            node_size[k] = prod(node_shape[k]);
            node_off[k] = acc;
            acc += node_size[k];
            if (k < n_cliques) clique_entries += node_size[k];
            else sep_entries += node_size[k];
        }
        up_base = clique_entries + sep_entries;
        down_base = up_base + sep_entries;
        // soft evidence: likelihood tables after the down-messages, each multiplied into the
        // smallest clique containing its variable
        lik_base = down_base + sep_entries;
        for (int32_t v : lik_vars) {
            int32_t best = -1;
            for (int32_t c = 0; c < n_cliques; ++c)
                if (std::find(node_vars[c].begin(), node_vars[c].end(), v) != node_vars[c].end() &&
                    (best < 0 || node_size[c] < node_size[best]))
                    best = c;
            if (best < 0) return jt_fail(JT_ERR_INVALID, "host compile: no clique contains a likelihood variable");
            lik_clique.push_back(best);
            lik_off.push_back(lik_entries);
            lik_entries += sizes[v];
        }

        // factor tables, evidence strides, output scopes
        evf_ptr.push_back(0);
        for (int32_t v : evidence_vars) ev_card.push_back(full_sizes[v]);
        if (has_factors) {
            std::vector<int32_t> ev_index(n_vars, -1);
            for (size_t i = 0; i < evidence_vars.size(); ++i) ev_index[evidence_vars[i]] = (int32_t)i;
            for (size_t f = 0; f < factors.size(); ++f) {
                std::vector<int64_t> full;
                for (int32_t v : factors[f]) full.push_back(ev_index[v] >= 0 ? full_sizes[v] : sizes[v]);
                fin_shape.push_back(full);
                fin_off.push_back(fin_entries);
                fin_size.push_back(prod(full));
                fin_entries += prod(full);
                const std::vector<int64_t> st = row_major_strides(full);
                for (size_t i = 0; i < factors[f].size(); ++i)
                    if (ev_index[factors[f][i]] >= 0) {
                        evf_var.push_back(ev_index[factors[f][i]]);
                        evf_stride.push_back(st[i]);
                    }
                evf_ptr.push_back((int64_t)evf_var.size());
            }
            for (size_t k = 0; k < out_scopes.size(); ++k) {
                std::vector<int64_t> eff;
                for (int32_t v : out_scopes[k]) eff.push_back(sizes[v]);
                fout_off.push_back(fout_entries);
                fout_size.push_back(prod(eff));
                fout_entries += prod(eff);
            }
        }

        // uniform cliques / subtrees (schedule.py: _find_uniform_cliques)
        uniform.assign(n_cliques, 0);
        uniform_up.assign(n_cliques, 0);
        uniform_down.assign(n_seps, 0);
        if (has_factors && has_tree) {
            std::vector<char> observed(n_vars, 0), touched(n_cliques, 0);
            for (int32_t v : evidence_vars) observed[v] = 1;
            for (size_t f = 0; f < factors.size(); ++f)
                for (int32_t v : factors[f])
                    if (observed[v]) touched[f2c[f]] = 1;
            for (int32_t c : lik_clique) touched[c] = 1;      // a likelihood makes the potential per-instance
            if (children[root].empty()) touched[root] = 1;
            for (size_t i = order.size(); i-- > 0;) {
                const int32_t c = order[i];
                uniform[c] = !touched[c];
                bool up = uniform[c];
                for (const auto& kid : children[c]) up = up && uniform_up[kid.second];
                uniform_up[c] = up;
            }
            for (int32_t c = 0; c < n_cliques; ++c)
                if (uniform[c]) uni_entries += node_size[c];
            // a down-message is uniform when everything on its source side is
            for (int32_t c : order) {
                const bool above = parent[c] < 0 || uniform_down[parent_sep[c] - n_cliques];
                for (const auto& kid : children[c]) {
                    bool u = uniform[c] && above;
                    for (const auto& other : children[c])
                        if (other.first != kid.first) u = u && uniform_up[other.second];
                    uniform_down[kid.first - n_cliques] = u;
                }
            }
        }
        by_depth.assign(max_depth + 1, {});
        for (int32_t c : order) by_depth[depth[c]].push_back(c);
        return JT_OK;
    }

    void build_init() {
        if (!has_factors) return;
        std::vector<std::vector<int32_t>> by_clique(n_cliques);
        for (size_t f = 0; f < f2c.size(); ++f) by_clique[f2c[f]].push_back((int32_t)f);
        const size_t begin = tasks.size();
        std::vector<int32_t> ordered;
        size_t n_uniform = 0;
        for (int32_t c = 0; c < n_cliques; ++c)
            if (uniform[c]) { ordered.push_back(c); ++n_uniform; }
        for (int32_t c = 0; c < n_cliques; ++c)
            if (!uniform[c]) ordered.push_back(c);
        for (int32_t c : ordered) {
            const Space s_space = make_space(node_vars[c], sizes);
            std::vector<int64_t> row = new_task(JT_KIND_INIT, s_space, nullptr, c, -1, false);
            row[JT_T_OUT] = node_off[c];
            if (uniform[c]) row[JT_T_FLAGS] |= JT_TF_TASK_UNIFORM;
            row[JT_T_SMSG_BEGIN] = row[JT_T_RMSG_BEGIN] = row[JT_T_RMSG_END] = (int64_t)msgs.size();
            for (int32_t f : by_clique[c]) {
                // observed axes contribute through the per-instance base offset only
                set_strides(factors[f], row_major_strides(fin_shape[f]));
                clear_strides(evidence_vars);
                add_msg(fin_off[f], s_space, nullptr, f, false);
                clear_strides(factors[f]);
            }
            for (size_t k = 0; k < lik_vars.size(); ++k)
                if (lik_clique[k] == c) {                     // fid -2: operand read from the workspace
                    stride_of[lik_vars[k]] = 1;
                    add_msg(lik_base + lik_off[k], s_space, nullptr, -2, false);
                    stride_of[lik_vars[k]] = 0;
                }
            row[JT_T_SMSG_END] = (int64_t)msgs.size();
            tasks.push_back(std::move(row));
        }
        launch_split(JT_PHASE_INIT, JT_PHASE_INIT_UNIFORM, JT_PHASE_INIT_INSTANCE, begin, begin + n_uniform, 0);
    }

    std::vector<Incoming> child_ups(int32_t c) const {
        std::vector<Incoming> in;
        for (const auto& kid : children[c]) in.push_back({up_off(kid.first), kid.first, (bool)uniform_up[kid.second]});
        return in;
    }

    void build_collect() {
        for (int32_t d = max_depth; d >= 1; --d) {
            const size_t begin = tasks.size();
            std::vector<int32_t> ordered;
            size_t n_uniform = 0;
            for (int32_t c : by_depth[d])
                if (uniform_up[c]) { ordered.push_back(c); ++n_uniform; }
            for (int32_t c : by_depth[d])
                if (!uniform_up[c]) ordered.push_back(c);
            for (int32_t c : ordered) {
                const int32_t psep = parent_sep[c];
                const Space s_space = make_space(node_vars[psep], sizes);
                const Space r_space = make_space(minus(node_vars[c], node_vars[psep]), sizes);
                std::vector<int64_t> row = new_task(JT_KIND_PROJECT, s_space, &r_space, c, c, true);
                row[JT_T_OUT] = up_off(psep);
                if (uniform_up[c]) row[JT_T_FLAGS] |= JT_TF_TASK_UNIFORM;
                attach_msgs(row, child_ups(c), s_space, r_space);
                tasks.push_back(std::move(row));
            }
            launch_split(JT_PHASE_COLLECT, JT_PHASE_COLLECT_UNIFORM, JT_PHASE_COLLECT_INSTANCE, begin,
                         begin + n_uniform, d);
        }
    }

    struct Deferred {
        std::vector<int64_t> row;
        std::vector<Incoming> others;
        Space s_space, r_space;
        bool sending;      // messages still to be attached (false: a finished belief-only task)
    };

    Incoming down_msg(int32_t c) const {
        const int32_t psep = parent_sep[c];
        return {down_off(psep), psep, (bool)uniform_down[psep - n_cliques]};
    }

    // Uniform mode: a down-message whose source side is evidence-free is computed once in the
    // uniform workspace (JT_PHASE_DIST_UNIFORM) and read by its consumers as a broadcast scalar;
    // a non-writer task with such a message shrinks to an elementwise task in the instance
    // launch.  Task order of the first launch of a level: [full form, uniform down | full form,
    // others | elementwise forms]; JT_PHASE_DIST_PRE is the first two groups,
    // JT_PHASE_DIST_PRE_INSTANCE the last two.  (schedule.py: _build_distribute)
    void build_distribute() {
        for (int32_t d = 0; d <= max_depth; ++d) {
            std::vector<Deferred> pre_ud, pre_other, main_sending, main_leaves, inst, unis;
            for (int32_t c : by_depth[d]) {
                const auto& kids = children[c];
                std::vector<Incoming> incoming;
                if (parent[c] >= 0) incoming.push_back(down_msg(c));
                for (const Incoming& m : child_ups(c)) incoming.push_back(m);
                if (kids.empty()) {
                    if (parent[c] < 0) continue;              // single-clique tree: belief = potential
                    Deferred t;
                    t.s_space = make_space(node_vars[c], sizes);
                    t.r_space = make_space({}, sizes);
                    t.row = new_task(JT_KIND_PROJECT, t.s_space, &t.r_space, c, c, true);
                    t.row[JT_T_BETA] = node_off[c];
                    attach_msgs(t.row, incoming, t.s_space, t.r_space);
                    t.sending = false;
                    main_leaves.push_back(std::move(t));
                    continue;
                }
                for (size_t i = 0; i < kids.size(); ++i) {
                    const int32_t sep = kids[i].first, kid = kids[i].second;
                    const bool ud = uniform_down[sep - n_cliques];
                    const bool writer = i + 1 == kids.size();
                    Deferred t;
                    t.s_space = make_space(node_vars[sep], sizes);
                    t.r_space = make_space(minus(node_vars[c], node_vars[sep]), sizes);
                    t.row = new_task(JT_KIND_PROJECT, t.s_space, &t.r_space, c, c, true);
                    t.row[JT_T_OUT] = down_off(sep);
                    t.row[JT_T_BEL] = bel_off(sep);
                    t.row[JT_T_OWN] = up_off(sep);
                    if (uniform_up[kid]) t.row[JT_T_FLAGS] |= JT_TF_OWN_UNIFORM;
                    for (const Incoming& m : incoming)
                        if (m.sep != sep) t.others.push_back(m);
                    t.sending = true;
                    if (writer) t.row[JT_T_BETA] = node_off[c];
                    Deferred u, e;
                    if (ud) {
                        u.s_space = t.s_space;
                        u.r_space = t.r_space;
                        u.others = t.others;
                        u.row = new_task(JT_KIND_PROJECT, u.s_space, &u.r_space, c, c, true);
                        u.row[JT_T_OUT] = down_off(sep);
                        u.row[JT_T_FLAGS] |= JT_TF_TASK_UNIFORM;
                        u.sending = true;
                        if (!writer) {
                            e.s_space = t.s_space;
                            e.r_space = make_space({}, sizes);
                            e.row = new_task(JT_KIND_PROJECT, e.s_space, &e.r_space, c, sep, false);
                            e.row[JT_T_SRC] = down_off(sep);
                            e.row[JT_T_FLAGS] |= JT_TF_SRC_UNIFORM;
                            e.row[JT_T_OUT] = down_off(sep);
                            e.row[JT_T_BEL] = bel_off(sep);
                            e.row[JT_T_OWN] = up_off(sep);
                            if (uniform_up[kid]) e.row[JT_T_FLAGS] |= JT_TF_OWN_UNIFORM;
                            e.sending = true;
                        }
                    }
                    if (writer) main_sending.push_back(std::move(t));
                    else if (ud) pre_ud.push_back(std::move(t));
                    else pre_other.push_back(std::move(t));
                    if (ud) {
                        unis.push_back(std::move(u));
                        if (!writer) inst.push_back(std::move(e));
                    }
                }
            }
            size_t n_sending = 0;
            auto place = [&](std::vector<Deferred>& group) {
                for (Deferred& t : group) {
                    if (t.sending) {
                        attach_msgs(t.row, t.others, t.s_space, t.r_space);
                        ++n_sending;
                    }
                    tasks.push_back(std::move(t.row));
                }
            };
            size_t begin = tasks.size();
            const size_t n_ud = pre_ud.size(), n_full = pre_ud.size() + pre_other.size();
            place(pre_ud);
            place(pre_other);
            place(inst);
            if (n_full) launches.push_back({JT_PHASE_DIST_PRE, (int64_t)begin, (int64_t)(begin + n_full), d});
            if (tasks.size() > begin + n_ud)
                launches.push_back({JT_PHASE_DIST_PRE_INSTANCE, (int64_t)(begin + n_ud), (int64_t)tasks.size(), d});
            // tasks that send a message first, the belief-only ones (leaves) last
            begin = tasks.size();
            n_sending = 0;
            place(main_sending);
            place(main_leaves);
            launch(JT_PHASE_DIST_MAIN, begin, d);
            if (n_sending)
                launches.push_back({JT_PHASE_DIST_MAIN_MESSAGES, (int64_t)begin, (int64_t)(begin + n_sending), d});
            begin = tasks.size();
            place(unis);
            launch(JT_PHASE_DIST_UNIFORM, begin, d);
        }
    }

    void build_marginal() {
        if (!has_factors) return;
        size_t begin = tasks.size();
        for (size_t k = 0; k < out_scopes.size(); ++k) {
            const int32_t c = out_clique[k];
            const Space s_space = make_space(out_scopes[k], sizes);
            const Space r_space = make_space(minus(node_vars[c], out_scopes[k]), sizes);
            std::vector<int64_t> row = new_task(JT_KIND_PROJECT, s_space, &r_space, c, c, false);
            row[JT_T_OUT] = fout_off[k];
            row[JT_T_OUT_SPACE] = 1;
            row[JT_T_AUX] = (int64_t)k;
            row[JT_T_RMSG_BEGIN] = row[JT_T_RMSG_END] = row[JT_T_SMSG_BEGIN] = row[JT_T_SMSG_END] = (int64_t)msgs.size();
            tasks.push_back(std::move(row));
        }
        launch(JT_PHASE_MARGINAL, begin, 0);
        if (!has_tree) return;
        begin = tasks.size();
        for (size_t k = 0; k < out_scopes.size(); ++k) {
            const int32_t c = out_clique[k];
            const Space s_space = make_space(out_scopes[k], sizes);
            const Space r_space = make_space(minus(node_vars[c], out_scopes[k]), sizes);
            std::vector<int64_t> row = new_task(JT_KIND_PROJECT, s_space, &r_space, c, c, true);
            row[JT_T_OUT] = fout_off[k];
            row[JT_T_OUT_SPACE] = 1;
            row[JT_T_AUX] = (int64_t)k;
            std::vector<Incoming> incoming;
            if (parent[c] >= 0) incoming.push_back(down_msg(c));
            for (const Incoming& m : child_ups(c)) incoming.push_back(m);
            attach_msgs(row, incoming, s_space, r_space);
            tasks.push_back(std::move(row));
        }
        launch(JT_PHASE_MARGINAL_DIRECT, begin, 0);
    }

    std::vector<int64_t> blob() const {
        std::vector<int64_t> w(JT_H_WORDS, 0);
        w[JT_H_MAGIC] = JT_MAGIC;
        w[JT_H_VERSION] = JT_ABI_VERSION;
        w[JT_H_NCLIQUES] = n_cliques;
        w[JT_H_NSEPS] = n_seps;
        w[JT_H_NFACTORS] = has_factors ? (int64_t)factors.size() : 0;
        w[JT_H_NEVID] = (int64_t)evidence_vars.size();
        w[JT_H_CLIQUE_ENTRIES] = clique_entries;
        w[JT_H_SEP_ENTRIES] = sep_entries;
        w[JT_H_FIN_ENTRIES] = fin_entries;
        w[JT_H_FOUT_ENTRIES] = fout_entries;
        w[JT_H_NTAB] = (int64_t)tab.size();
        w[JT_H_NTASKS] = (int64_t)tasks.size();
        w[JT_H_NMSGS] = (int64_t)msgs.size();
        w[JT_H_NLAUNCHES] = (int64_t)launches.size();
        w[JT_H_MAXDEPTH] = max_depth;
        w[JT_H_NEVF] = (int64_t)evf_var.size();
        w[JT_H_ROOT_ENTRIES] = root >= 0 ? node_size[root] : 0;
        w[JT_H_UNI_ENTRIES] = uni_entries;
        w[JT_H_NOUT] = (int64_t)fout_off.size();
        w[JT_H_LIK_ENTRIES] = lik_entries;
        auto put = [&](const std::vector<int64_t>& v) { w.insert(w.end(), v.begin(), v.end()); };
        put(node_off);
        put(node_size);
        put(fin_off);
        put(fin_size);
        put(fout_off);
        put(fout_size);
        put(ev_card);
        if (has_factors) put(evf_ptr);
        put(evf_var);
        put(evf_stride);
        for (const auto& t : tasks) put(t);
        for (const auto& m : msgs) put(m);
        for (const auto& l : launches) put(l);
        std::vector<int32_t> t32 = tab;
        if (t32.size() % 2) t32.push_back(0);
        const size_t words = w.size();
        w.resize(words + t32.size() / 2);
        memcpy(w.data() + words, t32.data(), t32.size() * sizeof(int32_t));
        return w;
    }
};

void read_csr(std::vector<std::vector<int32_t>>& dst, int32_t n, const int32_t* ptr, const int32_t* data) {
    dst.resize(n);
    for (int32_t i = 0; i < n; ++i) dst[i].assign(data + ptr[i], data + ptr[i + 1]);
}

bool has_duplicates(std::vector<int32_t> v) {
    std::sort(v.begin(), v.end());
    return std::adjacent_find(v.begin(), v.end()) != v.end();
}

bool subset(const std::vector<int32_t>& a, const std::vector<int32_t>& b) {
    for (int32_t v : a)
        if (std::find(b.begin(), b.end(), v) == b.end()) return false;
    return true;
}

}  // namespace

extern "C" int jt_plan_build(int32_t n_vars, const int64_t* sizes, const int64_t* full_sizes, int32_t n_cliques,
                             int32_t n_seps, const int32_t* node_ptr, const int32_t* node_vars, int32_t has_tree,
                             const int32_t* order, const int32_t* parent, const int32_t* parent_sep,
                             int32_t n_factors, const int32_t* factor_ptr, const int32_t* factor_vars,
                             const int32_t* factor_to_clique, int32_t n_evidence, const int32_t* evidence_vars,
                             int32_t n_outputs, const int32_t* output_ptr, const int32_t* output_vars,
                             int32_t n_likelihood, const int32_t* likelihood_vars, void** blob, size_t* nbytes) {
    if (!blob || !nbytes) return bad("null output");
    *blob = nullptr;
    *nbytes = 0;
    if (n_vars < 0 || (n_vars > 0 && !sizes) || n_cliques < 0 || n_seps < 0) return bad("sizes");
    if (!full_sizes) full_sizes = sizes;
    Emitter e;
    e.n_vars = n_vars;
    e.n_cliques = n_cliques;
    e.n_seps = has_tree ? n_seps : 0;
    e.n_nodes = n_cliques + e.n_seps;
    e.has_tree = has_tree != 0;
    e.has_factors = n_factors >= 0;
    if (!csr_ok(e.n_nodes, node_ptr, node_vars, n_vars)) return bad("node variable lists");
    read_csr(e.node_vars, e.n_nodes, node_ptr, node_vars);
    e.sizes.assign(sizes, sizes + n_vars);
    e.full_sizes.assign(full_sizes, full_sizes + n_vars);
    for (int32_t v = 0; v < n_vars; ++v)
        if (e.sizes[v] <= 0 || e.full_sizes[v] <= 0) return bad("variable size must be positive");
    for (const auto& vs : e.node_vars)
        if (has_duplicates(vs)) return bad("duplicate variable in a clique or separator");

    // tree arrays: breadth-first order, parent and parent separator per clique
    e.parent.assign(n_cliques, -1);
    e.parent_sep.assign(n_cliques, -1);
    e.depth.assign(n_cliques, 0);
    e.children.assign(n_cliques, {});
    if (e.has_tree) {
        if (n_cliques == 0 || !order || !parent || !parent_sep) return bad("tree arrays");
        if (e.n_seps != n_cliques - 1) return bad("a tree over N cliques has N - 1 separators");
        std::vector<char> seen(n_cliques, 0), sep_used(e.n_seps, 0);
        e.order.assign(order, order + n_cliques);
        for (int32_t i = 0; i < n_cliques; ++i) {
            const int32_t c = order[i];
            if (c < 0 || c >= n_cliques || seen[c]) return bad("tree order is not a permutation of the cliques");
            seen[c] = 1;
            const int32_t p = parent[c], s = parent_sep[c];
            if (i == 0) {
                if (p != -1 || s != -1) return bad("the first clique of the order must be the root");
                continue;
            }
            if (p < 0 || p >= n_cliques || p == c || !seen[p]) return bad("tree order must list parents before children");
            if (s < n_cliques || s >= e.n_nodes || sep_used[s - n_cliques]) return bad("separator ids must be N..N+S-1, each used once");
            sep_used[s - n_cliques] = 1;
            if (!subset(e.node_vars[s], e.node_vars[c]) || !subset(e.node_vars[s], e.node_vars[p]))
                return bad("separator is not contained in both of its cliques");
            e.parent[c] = p;
            e.parent_sep[c] = s;
            e.depth[c] = e.depth[p] + 1;
            e.max_depth = std::max(e.max_depth, e.depth[c]);
            e.children[p].push_back({s, c});
        }
        e.root = order[0];
    } else {
        for (int32_t c = 0; c < n_cliques; ++c) e.order.push_back(c);
    }

    if (n_evidence < 0 || (n_evidence > 0 && !evidence_vars)) return bad("evidence variables");
    for (int32_t i = 0; i < n_evidence; ++i) {
        const int32_t v = evidence_vars[i];
        if (v < 0 || v >= n_vars) return bad("evidence variable index");
        if (e.sizes[v] != 1) return bad("observed variable must have effective size 1");
        e.evidence_vars.push_back(v);
    }
    if (e.has_factors) {
        if (!csr_ok(n_factors, factor_ptr, factor_vars, n_vars) || (n_factors > 0 && !factor_to_clique))
            return bad("factor lists");
        read_csr(e.factors, n_factors, factor_ptr, factor_vars);
        for (int32_t f = 0; f < n_factors; ++f) {
            const int32_t home = factor_to_clique[f];
            if (has_duplicates(e.factors[f])) return bad("duplicate variable in a factor");
            if (home < 0 || home >= n_cliques || !subset(e.factors[f], e.node_vars[home]))
                return bad("factor is not contained in its clique");
            e.f2c.push_back(home);
        }
        if (n_outputs < 0) {                                  // default: the factor scopes
            e.out_scopes = e.factors;
            e.out_clique = e.f2c;
        } else {
            if (!csr_ok(n_outputs, output_ptr, output_vars, n_vars)) return bad("output scopes");
            read_csr(e.out_scopes, n_outputs, output_ptr, output_vars);
        }
    }
    if (n_likelihood < 0 || (n_likelihood > 0 && !likelihood_vars)) return bad("likelihood variables");
    if (n_likelihood > 0 && (!e.has_factors || !e.has_tree)) return bad("soft evidence needs the factor graph and the tree");
    for (int32_t i = 0; i < n_likelihood; ++i) {
        const int32_t v = likelihood_vars[i];
        if (v < 0 || v >= n_vars) return bad("likelihood variable index");
        if (std::find(e.evidence_vars.begin(), e.evidence_vars.end(), v) != e.evidence_vars.end())
            return bad("an observed variable cannot also carry a likelihood");
        if (std::find(e.lik_vars.begin(), e.lik_vars.end(), v) != e.lik_vars.end())
            return bad("duplicate likelihood variable");
        e.lik_vars.push_back(v);
    }
    int rc = e.prepare();
    if (rc != JT_OK) return rc;
    if (e.has_factors && n_outputs >= 0) {
        // home of an explicit output scope: the smallest clique containing it
        for (const auto& scope : e.out_scopes) {
            if (has_duplicates(scope)) return bad("duplicate variable in an output scope");
            int32_t best = -1;
            for (int32_t c = 0; c < n_cliques; ++c)
                if (subset(scope, e.node_vars[c]) && (best < 0 || e.node_size[c] < e.node_size[best])) best = c;
            if (best < 0) return bad("no clique contains an output scope");
            e.out_clique.push_back(best);
        }
    }
    e.build_init();
    e.build_collect();
    e.build_distribute();
    e.build_marginal();
    if (e.overflow) return bad("index table entry does not fit int32");
    const std::vector<int64_t> words = e.blob();
    void* mem = malloc(words.size() * sizeof(int64_t));
    if (!mem) return jt_fail(JT_ERR_NOMEM, "out of host memory");
    memcpy(mem, words.data(), words.size() * sizeof(int64_t));
    *blob = mem;
    *nbytes = words.size() * sizeof(int64_t);
    return JT_OK;
}
