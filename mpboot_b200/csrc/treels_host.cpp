// A minimal host-side stand-in for the pieces of class IQTree that mpgpu_optimize_spr_bb calls
// back into: treels_logl (iqtree.h, vector<double>), the treels map keyed by the tree's
// canonical form (iqtree.cpp:3299-3312, 3701-3708) and the random_double() stream.  The real
// host wires the hooks to IQTree's own members instead; tests and bench.py
// use this container so that no Python sits between the device and the bookkeeping.
// The canonical form here is a 64-bit order-independent hash of the unrooted topology (the
// reference uses the sorted-taxa newick string; both identify the topology).
#include "mpgpu_internal.h"

#include <algorithm>
#include <cstring>
#include <set>
#include <unordered_map>

struct mpgpu_treels {
    int n = 0;
    std::vector<double> logl;                              // treels_logl
    std::unordered_map<uint64_t, int32_t> index;           // treels
    std::vector<int64_t> mats;                             // 4 per materialised tree: remove_ref, insert_ref, tree_index, fingerprint
    std::vector<std::set<int32_t>> mulhits;                // boot_trees_parsimony (-mulhits)
    std::vector<std::vector<std::pair<int32_t, int32_t>>> top;   // boot_trees_parsimony_top (-mulhits -topboot, -distinct_iter_top_boot)
    std::vector<std::vector<int32_t>> top_iter;            // boot_trees_parsimony_top_iter (-distinct_iter_top_boot)
    mpgpu_rng_fn rng = nullptr; void *rng_user = nullptr;
};

namespace {

uint64_t fin64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

uint64_t subtree_hash(const mpgpu::HostTree &t, int ref)
{
    // iterative post-order: hash of the subtree that holds ref's node, entered through ref
    struct Frame { int ref; int state; uint64_t a; };
    std::vector<Frame> st;
    st.push_back({ref, 0, 0});
    uint64_t ret = 0;
    while (!st.empty()) {
        Frame &f = st.back();
        if (t.is_tip(f.ref)) { ret = fin64((uint64_t)(f.ref / 3)); st.pop_back(); continue; }
        if (f.state == 0) { f.state = 1; st.push_back({t.back(t.next(f.ref)), 0, 0}); continue; }
        if (f.state == 1) { f.a = ret; f.state = 2; st.push_back({t.back(t.next(t.next(f.ref))), 0, 0}); continue; }
        uint64_t a = f.a, b = ret;
        if (a > b) std::swap(a, b);
        ret = fin64(a * 0x9E3779B97F4A7C15ULL + b + 0x632BE59BD9B4E019ULL);
        st.pop_back();
    }
    return ret;
}

double hook_rng(void *user)
{
    mpgpu_treels *h = (mpgpu_treels *)user;
    return h->rng(h->rng_user);
}

int32_t hook_push(void *user, double cur_logl)
{
    mpgpu_treels *h = (mpgpu_treels *)user;
    h->logl.push_back(cur_logl);                           // iqtree.cpp:3345-3348
    return (int32_t)h->logl.size() - 1;
}

int32_t hook_materialize(void *user, const int32_t *bn, const int32_t *bs, int32_t remove_ref, int32_t insert_ref, int32_t tree_index)
{
    mpgpu_treels *h = (mpgpu_treels *)user;
    mpgpu::HostTree t;
    t.n = h->n;
    const int len = 3 * (2 * h->n - 1);
    t.bn.assign(bn, bn + len); t.bs.assign(bs, bs + len);
    if (remove_ref) mpgpu::apply_spr_move(t, remove_ref, insert_ref);
    const uint64_t fp = fin64(subtree_hash(t, t.back(3)) ^ 0x1234567ULL);
    auto it = h->index.find(fp);                           // treels.find(tree_str), iqtree.cpp:3701-3706
    if (it != h->index.end()) tree_index = it->second;
    else h->index[fp] = tree_index;
    const int64_t rec[4] = {remove_ref, insert_ref, tree_index, (int64_t)fp};
    h->mats.insert(h->mats.end(), rec, rec + 4);
    return tree_index;
}

void hook_mulhit(void *user, int32_t sample, int32_t tree_index, int32_t clear_first)
{
    mpgpu_treels *h = (mpgpu_treels *)user;
    if ((size_t)sample >= h->mulhits.size()) h->mulhits.resize((size_t)sample + 1);
    std::set<int32_t> &s = h->mulhits[sample];
    if (clear_first) s.clear();                            // iqtree.cpp:3517-3520
    s.insert(tree_index);                                  // :3531-3534
}

int32_t hook_tophit(void *user, int32_t sample, int32_t tree_index, int32_t rell, int32_t pop_worst)
{
    mpgpu_treels *h = (mpgpu_treels *)user;
    if ((size_t)sample >= h->top.size()) h->top.resize((size_t)sample + 1);
    std::vector<std::pair<int32_t, int32_t>> &t = h->top[sample];
    if (pop_worst && !t.empty()) t.pop_back();             // iqtree.cpp:3571
    size_t pos = 0;
    while (pos < t.size() && !(t[pos].second < rell)) pos++;   // :3563-3566
    t.insert(t.begin() + pos, std::make_pair(tree_index, rell));
    return t.back().second;
}

// -distinct_iter_top_boot: iqtree.cpp:3624-3677
int32_t hook_disthit(void *user, int32_t sample, int32_t tree_index, int32_t rell, int32_t cur_it, int32_t top_n, int32_t threshold)
{
    mpgpu_treels *h = (mpgpu_treels *)user;
    if ((size_t)sample >= h->top.size()) h->top.resize((size_t)sample + 1);
    if ((size_t)sample >= h->top_iter.size()) h->top_iter.resize((size_t)sample + 1);
    std::vector<std::pair<int32_t, int32_t>> &top = h->top[sample];
    std::vector<int32_t> &iter = h->top_iter[sample];
    const int t = std::min((int)top_n, (int)iter.size());
    int c;
    for (c = 0; c < t; c++) if (top[c].first == tree_index) return threshold;          // the tree is in the list: nothing (:3627-3634)
    for (c = 0; c < t; c++)                                                            // this iteration has an entry: keep the better (:3637-3645)
        if (iter[c] == cur_it) {
            if (rell > top[c].second) { top[c].second = rell; top[c].first = tree_index; }
            break;
        }
    if (c == t && t < top_n) { iter.push_back(cur_it); top.push_back(std::make_pair(tree_index, rell)); }   // :3648-3651
    else if (c == t && t == top_n) {                                                   // full: the worst entry goes (:3654-3668)
        int worst = 0;
        for (int d = 1; d < t; d++) if (top[d].second < top[worst].second) worst = d;
        top[worst] = std::make_pair(tree_index, rell);
        iter[worst] = cur_it;
    }
    int32_t thr = top[0].second;                                                       // :3671-3677
    for (size_t d = 1; d < top.size(); d++) if (top[d].second < thr) thr = top[d].second;
    return thr;
}

}  // namespace

extern "C" {

// A self-contained random_double() for hosts without their own stream (tests, bench.py):
// splitmix64 on the uint64_t state `user` points to.  MPBoot proper passes its SPRNG stream.
double mpgpu_splitmix64_double(void *user)
{
    uint64_t *state = (uint64_t *)user;
    uint64_t z = (*state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

mpgpu_treels *mpgpu_treels_create(int ntaxa)
{
    mpgpu_treels *h = new mpgpu_treels();
    h->n = ntaxa;
    return h;
}
void mpgpu_treels_destroy(mpgpu_treels *h) { delete h; }
int64_t mpgpu_treels_size(const mpgpu_treels *h) { return h ? (int64_t)h->logl.size() : 0; }
void mpgpu_treels_logl(const mpgpu_treels *h, double *out) { if (h && !h->logl.empty()) memcpy(out, h->logl.data(), h->logl.size() * sizeof(double)); }
int64_t mpgpu_treels_num_materialized(const mpgpu_treels *h) { return h ? (int64_t)h->mats.size() / 4 : 0; }
void mpgpu_treels_materialized(const mpgpu_treels *h, int64_t *out) { if (h && !h->mats.empty()) memcpy(out, h->mats.data(), h->mats.size() * sizeof(int64_t)); }
void mpgpu_treels_hooks(mpgpu_treels *h, mpgpu_rng_fn rng, void *rng_user, mpgpu_bb_hooks *out)
{
    h->rng = rng; h->rng_user = rng_user;
    out->user = h;
    out->random_double = hook_rng;
    out->push_tree_logl = hook_push;
    out->materialize = hook_materialize;
    out->mulhit = hook_mulhit;
    out->tophit = hook_tophit;
    out->disthit = hook_disthit;
}
// boot_trees_parsimony_top_iter: the iteration of every entry mpgpu_treels_toplists reports, in the same order
int64_t mpgpu_treels_topiters(const mpgpu_treels *h, int32_t nsamples, int32_t *flat, int64_t capacity)
{
    int64_t tot = 0;
    for (int32_t s = 0; s < nsamples; s++) {
        if (!h || (size_t)s >= h->top_iter.size()) continue;
        for (int32_t v : h->top_iter[s]) { if (flat && tot < capacity) flat[tot] = v; tot++; }
    }
    return tot;
}
int64_t mpgpu_treels_toplists(const mpgpu_treels *h, int32_t nsamples, int32_t *sizes, int32_t *flat, int64_t capacity)
{
    int64_t tot = 0;
    for (int32_t s = 0; s < nsamples; s++) {
        const bool have = h && (size_t)s < h->top.size();
        if (sizes) sizes[s] = have ? (int32_t)h->top[s].size() : 0;
        if (!have) continue;
        for (const auto &pr : h->top[s]) { if (flat && tot < capacity) { flat[2 * tot] = pr.first; flat[2 * tot + 1] = pr.second; } tot++; }
    }
    return tot;
}
int64_t mpgpu_treels_mulhits(const mpgpu_treels *h, int32_t nsamples, int32_t *sizes, int32_t *flat, int64_t capacity)
{
    int64_t tot = 0;
    for (int32_t s = 0; s < nsamples; s++) {
        const bool have = h && (size_t)s < h->mulhits.size();
        if (sizes) sizes[s] = have ? (int32_t)h->mulhits[s].size() : 0;
        if (!have) continue;
        for (int32_t v : h->mulhits[s]) { if (flat && tot < capacity) flat[tot] = v; tot++; }
    }
    return tot;
}

}  // extern "C"
