// C-ABI layer of libmpgpu.so (include/mpgpu.h): context management, data movement and the
// host-side orchestration of the kernels in fitch_kernels.cu.  No CPU fallback anywhere: every
// compute entry point needs a CUDA device and fails loudly otherwise.
#include "mpgpu_internal.h"

#include <chrono>
#include <cstdio>

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace mpgpu {

static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
int cuda_fail(cudaError_t e, const char *what)
{
    g_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    return 2;
}

static int states_of(int datatype)
{
    switch (datatype) {
    case MPGPU_BINARY_DATA: return 2;
    case MPGPU_DNA_DATA:    return 4;
    case MPGPU_AA_DATA:     return 20;
    case MPGPU_GENERIC_32:  return 32;
    default: return -1;
    }
}

// isInformative (sprparsimony.cpp:2460-2499): at least two distinct codes below `undetermined`
static void find_informative(Ctx *c, const uint8_t *yvec)
{
    int und = 0;
    const uint32_t *mt = state_mask_table(c->datatype, nullptr, &und);
    c->informative.assign(c->P, 1);
    c->present.assign(c->P, 0);
    for (int t = 0; t < c->n; t++) {                       // unambiguous states per pattern (ParsTree::findMstScore)
        const uint8_t *row = yvec + (size_t)t * c->P;
        for (int i = 0; i < c->P; i++) {
            const uint32_t m = mt[row[i]];
            if (m && !(m & (m - 1))) c->present[i] |= m;
        }
    }
    if (c->sort_alignment) {
        std::vector<int32_t> first(c->P, -1);
        std::fill(c->informative.begin(), c->informative.end(), 0);
        for (int t = 0; t < c->n; t++) {
            const uint8_t *row = yvec + (size_t)t * c->P;
            for (int i = 0; i < c->P; i++) {
                const int code = row[i];
                if (code >= und) continue;
                if (first[i] < 0) first[i] = code;
                else if (first[i] != code) c->informative[i] = 1;
            }
        }
    }
    c->n_inf = 0;
    for (int i = 0; i < c->P; i++) c->n_inf += c->informative[i];
}

static void free_alignment(Ctx *c)
{
    if (c->d_codes) cudaFree(c->d_codes);
    if (c->d_site_start) cudaFree(c->d_site_start);
    if (c->d_inf_ptn) cudaFree(c->d_inf_ptn);
    if (c->d_views) cudaFree(c->d_views);
    if (c->d_vcount) cudaFree(c->d_vcount);
    c->d_codes = nullptr; c->d_site_start = nullptr; c->d_inf_ptn = nullptr; c->d_views = nullptr; c->d_vcount = nullptr;
    c->tree_set = false; c->lens_valid = false;
}

int shard_sum(Ctx *c, void *dev_i32, int64_t count)
{
    if (c->shard_count == 1) return 0;         // (replicate shards hold complete planes: nothing of the Fitch / Sankoff path is exchanged)
    return group_sum(c, dev_i32, count);
}

int group_sum(Ctx *c, void *dev_i32, int64_t count)
{
    if (c->xcount() == 1 || count <= 0 || c->exchange_off) return 0;
    // small vectors (every count vector of the search): one kernel on the stream over NVLink peer memory, latency-bound;
    // large ones (-bb: rows x replicates of s32) are bandwidth-bound and go to the host's collective (NCCL ring / NVLS)
    // when one is installed next to the peer exchange -- the one-shot kernel moves shard_count x the vector
    if (c->peer.ready && (!c->allreduce || (size_t)count <= 4 * c->peer.cap)) return peer_allreduce(c, dev_i32, count);
    if (!c->allreduce) { set_error("sharded context: this call needs mpgpu_set_allreduce (or use the *_partial calls)"); return 1; }
    if (c->allreduce(c->allreduce_user, dev_i32, count, (void *)c->stream)) { set_error("the all-reduce callback failed"); return 1; }
    return 0;
}

// layout + tip planes for the current weights (compressDNA, sprparsimony.cpp:2864-2961)
static int build_planes(Ctx *c, bool realloc_views)
{
    std::vector<int64_t> site_start(c->n_inf + 1);
    std::vector<int32_t> inf_ptn(c->n_inf > 0 ? c->n_inf : 1);
    int64_t sites = 0; int k = 0;
    for (int i = 0; i < c->P; i++) {
        if (!c->informative[i]) continue;
        site_start[k] = sites; inf_ptn[k] = i; k++;
        sites += c->weights[i];
    }
    site_start[k] = sites;
    c->n_sites = sites;
    int64_t words = (sites + 31) / 32;                       // compressedEntries :2870
    int64_t refw = words % 8 ? words + (8 - words % 8) : words;   // padded to INTS_PER_VECTOR (AVX) :2876
    c->ref_words = (int)refw;
    const int64_t quantum = (int64_t)kWordPad * c->shard_count;
    int64_t glob = (words + quantum - 1) / quantum * quantum;
    if (glob == 0) glob = quantum;
    const int newWl = (int)(glob / c->shard_count);
    c->glob_words = (int)glob; c->Wl = newWl; c->w0 = (int64_t)c->shard_rank * newWl;
    {
        const int SG = c->S < 4 ? c->S : 4, G = (c->S + SG - 1) / SG;
        c->view_stride = (size_t)G * SG * c->Wl;      // state-interleaved groups, see fitch_kernels.cu
    }

    // the staging buffer is page-locked and owned by the context: wait for the previous upload from it (normally long
    // finished), then nothing below has to be waited for -- re-weighting (ratchet, replicates) is a copy and a launch
    MPGPU_CUDA(cudaStreamSynchronize(c->stream));
    if (!c->site_pin.reserve((size_t)c->n_inf + 1)) { set_error("pinned allocation failed"); return 1; }
    memcpy(c->site_pin.data(), site_start.data(), sizeof(int64_t) * (c->n_inf + 1));
    if (realloc_views || !c->d_site_start || !c->d_inf_ptn) {
        if (c->d_site_start) { cudaFree(c->d_site_start); c->d_site_start = nullptr; }
        if (c->d_inf_ptn) { cudaFree(c->d_inf_ptn); c->d_inf_ptn = nullptr; }
        MPGPU_CUDA(cudaMalloc((void **)&c->d_site_start, sizeof(int64_t) * (c->n_inf + 1)));
        MPGPU_CUDA(cudaMalloc((void **)&c->d_inf_ptn, sizeof(int32_t) * inf_ptn.size()));
        MPGPU_CUDA(cudaMemcpy(c->d_inf_ptn, inf_ptn.data(), sizeof(int32_t) * inf_ptn.size(), cudaMemcpyHostToDevice));   // informative patterns depend on the codes only
    }
    MPGPU_CUDA(cudaMemcpyAsync(c->d_site_start, c->site_pin.data(), sizeof(int64_t) * (c->n_inf + 1), cudaMemcpyHostToDevice, c->stream));

    const size_t nviews = (size_t)(4 * c->n - 6);
    const size_t need = nviews * c->view_stride;
    if (!c->d_views || need > c->views_alloc) {              // re-weighting (replicates, ratchet) keeps the allocation
        if (c->d_views) cudaFree(c->d_views);
        if (c->d_vcount) cudaFree(c->d_vcount);
        c->d_views = nullptr; c->d_vcount = nullptr; c->views_alloc = 0;
        MPGPU_CUDA(cudaMalloc((void **)&c->d_views, need * sizeof(uint32_t)));
        MPGPU_CUDA(cudaMalloc((void **)&c->d_vcount, (nviews + 4) * sizeof(uint32_t)));     // + slack: the peer exchange moves whole int4
        c->views_alloc = need;
    }
    if (c->n_inf > 0) { if (int rc = launch_compress(c)) return rc; }
    else MPGPU_CUDA(cudaMemsetAsync(c->d_views, 0xff, (size_t)c->n * c->view_stride * sizeof(uint32_t), c->stream));
    c->lens_valid = false;
    c->kids_valid = false;
    c->ptn_site_valid = false;
    c->reps.tree_valid = false;
    if (c->sk.on) return sk_build(c);
    return 0;
}

// dependency schedule of all directed views of c->tree -> c->sched (post-order: children first)
static void build_schedule(Ctx *c)
{
    const HostTree &t = c->tree;
    const int n = t.n;
    const int nviews = 4 * n - 6;
    std::vector<int32_t> &level = c->sc_level, &stack = c->sc_stack;
    level.assign(nviews, -1);
    for (int i = 0; i < n; i++) level[i] = 0;
    c->sched.clear();
    c->sched_levels = 0;
    stack.clear();
    for (int node = n + 1; node <= 2 * n - 2; node++) for (int s = 0; s < 3; s++) {
        const int root = 3 * node + s;
        if (!t.has_back(root)) continue;
        if (level[t.vid(root)] >= 0) continue;
        stack.push_back(root);
        while (!stack.empty()) {
            const int r = stack.back();
            const int v = t.vid(r);
            if (level[v] >= 0) { stack.pop_back(); continue; }
            const int a = t.back(t.next(r)), b = t.back(t.next(t.next(r)));
            const int la = level[t.vid(a)], lb = level[t.vid(b)];
            if (la >= 0 && lb >= 0) {
                const int l = std::max(la, lb) + 1;
                level[v] = l;
                if (l > c->sched_levels) c->sched_levels = l;
                Triple tr; tr.dst = v; tr.a = t.vid(a); tr.b = t.vid(b); tr.pad = l;
                c->sched.push_back(tr);
                stack.pop_back();
            } else {
                if (la < 0) stack.push_back(a);
                if (lb < 0) stack.push_back(b);
            }
        }
    }
}

static inline int2 kid_pair(const Triple &tr) { return tr.a < tr.b ? make_int2(tr.a, tr.b) : make_int2(tr.b, tr.a); }

// all directed views of c->tree: level schedule + one launch per level (throughput path)
int compute_views(Ctx *c, bool want_start_edge)
{
    const int nviews = 4 * c->n - 6;
    if (c->wave_pending) c->wcount_zeroed = false;      // lists were in flight and are dropped here: their counters are not zero
    c->wave_pending = 0; c->wave_lists.clear(); c->wave_used = 0; c->wc_used = 0; c->wave_fetched = false;
    c->vstale.assign(nviews, 0); c->n_stale = 0;
    c->views_stale = false;
    c->start_edge_valid = false;
    build_schedule(c);
    const size_t total = c->sched.size();
    const int nl = c->sched_levels;
    // One k_fitch_wave launch for the whole schedule when its list fits in shared memory (up to ~3000 taxa) instead of one launch
    // per dependency level: every search, every refinement replicate and every computeParsimony of the drop-in starts here, and
    // ~20-60 launches of 3.5 us each were most of it (C2: ~200 us -> ~50 us; 100 x 5000: 137 us per set_tree, 1251 of them in a
    // -bb 1000 run).  MPGPU_LEVEL_VIEWS=1 keeps the per-level launches.
    static const bool level_views = getenv("MPGPU_LEVEL_VIEWS") != nullptr;
    if (!level_views && !c->sk.on && c->reduces() && total > 0 &&
        wave_smem_bytes(c->S, (nl + 3) / 4 + (int)total) <= 200 * 1024) {
        std::vector<int32_t> &dl = c->sc_dl;
        std::vector<Triple> &all = c->sc_stale;
        dl.assign(nviews, 0);
        all.clear();
        c->vcount.assign(nviews, 0);
        c->view_kids.assign(nviews, make_int2(-1, -1));
        for (const Triple &tr0 : c->sched) {              // children first; pad = level
            Triple tr = tr0;
            const int l = tr.pad;
            const bool fa = dl[tr.a] != 0, fb = dl[tr.b] != 0;    // b = the operand that is not of this list (a tip) when there is one; a = the one of level l - 1
            if (!fa && fb) std::swap(tr.a, tr.b);
            else if (fa && fb && dl[tr.a] != l - 1) std::swap(tr.a, tr.b);
            dl[tr.dst] = l;
            all.push_back(tr);
            c->view_kids[tr.dst] = kid_pair(tr0);
        }
        c->dl_dirty = true;
        c->kids_valid = true;
        if (int rc = submit_stale(c, all, nl, true, false)) return rc;
        const bool edge_w = want_start_edge && c->reduces();
        if (edge_w) {
            MPGPU_CUDA(cudaMemsetAsync(c->d_scalar, 0, sizeof(uint32_t), c->stream));
            if (int rc = launch_edge_mismatch(c, c->tree.vid(3), c->tree.vid(c->tree.back(3)), c->d_scalar)) return rc;
            if (int rc = shard_sum(c, c->d_scalar, 1)) return rc;
            if (!c->vcount_pin.reserve((size_t)nviews + 1)) { set_error("pinned allocation failed"); return 1; }
            MPGPU_CUDA(cudaMemcpyAsync(c->vcount_pin.data() + nviews, c->d_scalar, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        }
        if (c->wave_pending) {                              // (0 when submit_stale fell back: then everything is settled already)
            if (int rc = fetch_wave_counts(c)) return rc;
            MPGPU_CUDA(cudaStreamSynchronize(c->stream));
            settle_views(c, false);
        } else MPGPU_CUDA(cudaStreamSynchronize(c->stream));
        if (edge_w) { c->start_edge_mis = c->vcount_pin.data()[nviews]; c->start_edge_valid = true; }
        c->reps.tree_valid = false;
        return 0;
    }
    if (int rc = ensure(c->d_triples, c->triples_cap, total)) return rc;
    std::vector<int32_t> start(nl + 2, 0);
    for (const Triple &tr : c->sched) start[tr.pad + 1]++;
    for (int l = 1; l <= nl + 1; l++) start[l] += start[l - 1];          // level l occupies [start[l], start[l+1])
    std::vector<Triple> flat(total);
    {
        std::vector<int32_t> fill(start.begin(), start.end());
        for (const Triple &tr : c->sched) { Triple x = tr; x.pad = 0; flat[fill[tr.pad]++] = x; }
    }
    MPGPU_CUDA(cudaMemcpyAsync(c->d_triples, flat.data(), total * sizeof(Triple), cudaMemcpyHostToDevice, c->stream));
    MPGPU_CUDA(cudaMemsetAsync(c->d_vcount, 0, nviews * sizeof(uint32_t), c->stream));
    if (c->sk.on) { if (int rc = sk_compute_levels(c, start, nl)) return rc; }
    else for (int l = 1; l <= nl; l++)
        if (int rc = launch_level(c, c->d_triples + start[l], start[l + 1] - start[l])) return rc;
    c->reps.tree_valid = false;
    if (c->shard_count > 1 && c->reduces()) { if (int rc = shard_sum(c, c->d_vcount, nviews)) return rc; }
    // the mismatch count across the edge at tip 1 (evaluateParsimony at tr->start, :3277) rides along with the counts
    const bool edge = want_start_edge && !c->sk.on && c->reduces();
    if (edge) {
        MPGPU_CUDA(cudaMemsetAsync(c->d_scalar, 0, sizeof(uint32_t), c->stream));
        if (int rc = launch_edge_mismatch(c, c->tree.vid(3), c->tree.vid(c->tree.back(3)), c->d_scalar)) return rc;
        if (int rc = shard_sum(c, c->d_scalar, 1)) return rc;
    }
    if (!c->vcount_pin.reserve((size_t)nviews + 1)) { set_error("pinned allocation failed"); return 1; }
    MPGPU_CUDA(cudaMemcpyAsync(c->vcount_pin.data(), c->d_vcount, nviews * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    if (edge) MPGPU_CUDA(cudaMemcpyAsync(c->vcount_pin.data() + nviews, c->d_scalar, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    MPGPU_CUDA(cudaStreamSynchronize(c->stream));
    c->vcount.assign(c->vcount_pin.data(), c->vcount_pin.data() + nviews);
    if (edge) { c->start_edge_mis = c->vcount_pin.data()[nviews]; c->start_edge_valid = true; }
    c->view_kids.assign(nviews, make_int2(-1, -1));
    for (const Triple &tr : c->sched) c->view_kids[tr.dst] = kid_pair(tr);
    c->kids_valid = true;
    return 0;
}

// After a move on c->tree (latency path): a view is stale iff its children changed or one of
// them is stale -- about half of the views, in a forest as deep as the tree.  They are
// recomputed by ONE k_fitch_wave launch; the mismatch counts of the stale views come back
// compact.  Needs complete counts on every shard (all-reduce callback) like compute_lengths.
// defer = true: nothing is waited for; the counts are scattered by settle_views() after the caller's next
// stream synchronize (the search loop plans and launches the next scan batch in the meantime).
int update_views(Ctx *c, bool defer)
{
    const int nviews = 4 * c->n - 6;
    if (!c->kids_valid || (int)c->view_kids.size() != nviews || (int)c->vcount.size() != nviews ||
        (c->shard_count > 1 && !c->reduces()) || getenv("MPGPU_NO_WAVE"))
        return compute_views(c);
    static const bool prof = getenv("MPGPU_PROFILE") != nullptr;
    static double t_host = 0, t_dev = 0; static long n_calls = 0, n_triples = 0, n_levels = 0;
    std::chrono::steady_clock::time_point p0, p1;
    if (c->n_stale) { if (int rc = ensure_all_views(c)) return rc; }      // the eager scheme starts from current views
    if (prof) p0 = std::chrono::steady_clock::now();
    build_schedule(c);
    c->dl_dirty = true;
    std::vector<int32_t> &dl = c->sc_dl;
    std::vector<Triple> &stale = c->sc_stale;
    dl.assign(nviews, 0);                               // stale level (0 = clean)
    stale.clear();
    int nlevels = 0;
    for (const Triple &tr0 : c->sched) {
        const int2 k = kid_pair(tr0), o = c->view_kids[tr0.dst];
        if (dl[tr0.a] == 0 && dl[tr0.b] == 0 && k.x == o.x && k.y == o.y) continue;
        Triple tr = tr0;
        const int l = std::max(dl[tr.a], dl[tr.b]) + 1;
        // b = the clean operand when there is one (prefetched); a = the stale one of the previous level (cached)
        const bool sa = dl[tr.a] != 0, sb = dl[tr.b] != 0;
        if (!sa && sb) std::swap(tr.a, tr.b);
        else if (sa && sb && dl[tr.a] != l - 1) std::swap(tr.a, tr.b);
        tr.pad = l;
        dl[tr.dst] = l;
        if (l > nlevels) nlevels = l;
        stale.push_back(tr);
        c->view_kids[tr.dst] = k;
    }
    if (prof) p1 = std::chrono::steady_clock::now();
    if (int rc = submit_stale(c, stale, nlevels, defer, false)) return rc;
    if (prof && !defer) {
        const auto p2 = std::chrono::steady_clock::now();
        t_host += std::chrono::duration<double>(p1 - p0).count(); t_dev += std::chrono::duration<double>(p2 - p1).count();
        n_calls++; n_triples += (long)stale.size(); n_levels += nlevels;
        if (n_calls % 256 == 0)
            fprintf(stderr, "[mpgpu profile] update_views x%ld: host %.1f us, launch..sync %.1f us per call; %.0f stale views in %.0f levels\n",
                    n_calls, 1e6 * t_host / n_calls, 1e6 * t_dev / n_calls, (double)n_triples / n_calls, (double)n_levels / n_calls);
    }
    return 0;
}

// The stale list (children before parents, pad = stale level) goes to the device as one k_fitch_wave launch (Fitch) or one
// launch per level (Sankoff); the mismatch counts of the recomputed views come back compact.  Several lists can be in
// flight (one per plan piece of a scan batch): they share the pinned / device staging arrays at increasing offsets and
// are landed together by settle_views() after the caller's next stream synchronize.  incremental: the lengths of the
// listed views are updated from their children's (every other view's length is current: the lazy scheme of the SPR
// search); otherwise settle_views recomputes all lengths from the schedule.
int submit_stale(Ctx *c, std::vector<Triple> &stale, int nlevels, bool defer, bool incremental)
{
    const int nviews = 4 * c->n - 6;
    const size_t total = stale.size();
    if (total == 0) return 0;
    if (!incremental) c->reps.tree_valid = false;      // (the lazy scheme dropped the tree rows when the move was marked)
    if (c->sk.on) {
        if (c->wave_pending) { if (int rc = fetch_wave_counts(c)) return rc; MPGPU_CUDA(cudaStreamSynchronize(c->stream)); settle_views(c, false); }
        if (int rc = sk_update_stale(c, stale, nlevels)) return rc;
        if (incremental) for (const Triple &tr : stale) c->vlen[tr.dst] = c->vcount[tr.dst] & c->sk.sum_mask();
        return 0;
    }
    const int hdr = (nlevels + 3) / 4;
    if (wave_smem_bytes(c->S, hdr + (int)total) > 200 * 1024) {            // list does not fit in shared memory: everything, level by level
        if (c->wave_pending) { if (int rc = fetch_wave_counts(c)) return rc; MPGPU_CUDA(cudaStreamSynchronize(c->stream)); settle_views(c, false); }
        c->kids_valid = false;
        if (int rc = compute_views(c)) return rc;
        if (c->reduces()) compute_lengths(c);
        return 0;
    }
    // staging for every list that can be in flight before the next settle: each view is listed at most once per settle
    // in the lazy scheme, the eager one has a single list
    const size_t cap_tr = (size_t)2 * nviews + 1024, cap_wc = (size_t)nviews + 1024;
    if (c->wave_pin.size() < cap_tr || c->wcount_pin.size() < cap_wc || c->wave_cap < cap_tr || c->wcount_cap < cap_wc ||
        c->wave_used + hdr + total > cap_tr || c->wc_used + total > cap_wc) {
        if (c->wave_pending) { if (int rc = fetch_wave_counts(c)) return rc; MPGPU_CUDA(cudaStreamSynchronize(c->stream)); settle_views(c, false); }
        if (!c->wave_pin.reserve(cap_tr) || !c->wcount_pin.reserve(cap_wc)) { set_error("pinned allocation failed"); return 1; }
        if (int rc = ensure(c->d_wave, c->wave_cap, cap_tr)) return rc;
        if (int rc = ensure(c->d_wcount, c->wcount_cap, cap_wc)) return rc;
        MPGPU_CUDA(cudaMemsetAsync(c->d_wcount, 0, c->wcount_cap * sizeof(uint32_t), c->stream));
        c->wcount_zeroed = true;
    }
    if (!c->wcount_zeroed) {
        MPGPU_CUDA(cudaMemsetAsync(c->d_wcount, 0, c->wcount_cap * sizeof(uint32_t), c->stream));
        c->wcount_zeroed = true;
    }
    std::vector<int32_t> &dl = c->sc_dl, &slot = c->sc_slot, &fill = c->sc_fill;     // dl[view] = stale level of the views of THIS list
    // counting sort by stale level straight into the pinned list; slots = position within the level
    const size_t off = c->wave_used, wc_off = c->wc_used;
    int32_t *level_end = reinterpret_cast<int32_t *>(c->wave_pin.data() + off);
    Triple *dst = c->wave_pin.data() + off + hdr;
    fill.assign(nlevels + 2, 0);
    for (const Triple &tr : stale) fill[tr.pad + 1]++;
    for (int l = 1; l <= nlevels + 1; l++) fill[l] += fill[l - 1];
    for (int l = 1; l <= nlevels; l++) level_end[l - 1] = fill[l + 1];
    for (int l = nlevels; l < 4 * hdr; l++) level_end[l] = (int32_t)total;
    const int cap = wave_slot_cap(c->S);
    slot.resize(nviews);
    for (const Triple &tr : stale) {                     // children precede parents in `stale`, so slot[tr.a] is final
        const int l = tr.pad;
        const int at = fill[l]++;
        const int pos = at - (l >= 2 ? level_end[l - 2] : 0);
        const int a_slot = dl[tr.a] == l - 1 && l > 1 ? slot[tr.a] : 0xFF;
        const int d_slot = pos < cap ? pos : 0xFF;
        slot[tr.dst] = d_slot;
        Triple x = tr;
        x.pad = a_slot | d_slot << 8 | (dl[tr.b] == 0 ? 0x10000 : 0);
        dst[at] = x;
    }
    // a short list is read by the kernel straight from the mapped staging array (every CTA reads it once: ~1 KB each over
    // PCIe costs less than a copy-engine hop); a long one (the eager scheme: ~2n views) goes through device memory
    const bool zero_copy = hdr + total <= 96;
    if (!zero_copy)
        MPGPU_CUDA(cudaMemcpyAsync(c->d_wave + off, c->wave_pin.data() + off, (hdr + total) * sizeof(Triple), cudaMemcpyHostToDevice, c->stream));
    if (int rc = launch_wave(c, zero_copy ? c->wave_pin.data() + off : c->d_wave + off, nlevels, hdr, (int)total, c->d_wcount + wc_off)) return rc;
    if (c->shard_count > 1) { if (int rc = shard_sum(c, c->d_wcount + wc_off, (int64_t)total)) return rc; }
    PendingWave pw; pw.list_off = (int)(off + hdr); pw.total = (int)total; pw.wc_off = (int)wc_off; pw.incremental = incremental;
    c->wave_lists.push_back(pw);
    c->wave_used += hdr + total; c->wc_used += total;
    c->wave_pending += (int)total;
    if (defer) return 0;
    if (int rc = fetch_wave_counts(c)) return rc;
    MPGPU_CUDA(cudaStreamSynchronize(c->stream));
    settle_views(c, false);
    return 0;
}

// the counts of every list in flight -> the pinned copy, and the device counters back to zero (stream-ordered; the caller
// synchronizes, then settle_views lands them)
int fetch_wave_counts(Ctx *c)
{
    if (!c->wave_pending || c->wave_fetched) return 0;
    MPGPU_CUDA(cudaMemcpyAsync(c->wcount_pin.data(), c->d_wcount, (size_t)c->wc_used * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    MPGPU_CUDA(cudaMemsetAsync(c->d_wcount, 0, (size_t)c->wc_used * sizeof(uint32_t), c->stream));
    c->wave_fetched = true;
    return 0;
}

// after a stream synchronize that followed fetch_wave_counts: the counts of the lists in flight land in vcount (and vlen)
void settle_views(Ctx *c, bool lengths)
{
    if (!c->wave_pending) return;
    const uint32_t *wcp = c->wcount_pin.data();
    bool full = false;
    for (const PendingWave &pw : c->wave_lists) {
        const Triple *list = c->wave_pin.data() + pw.list_off;
        const uint32_t *wc = wcp + pw.wc_off;
        if (pw.incremental) {
            if (c->sk.on) for (int i = 0; i < pw.total; i++) { c->vcount[list[i].dst] = wc[i]; c->vlen[list[i].dst] = wc[i] & c->sk.sum_mask(); }
            else for (int i = 0; i < pw.total; i++) {             // sorted by level: children first
                const Triple &tr = list[i];
                c->vcount[tr.dst] = wc[i];
                c->vlen[tr.dst] = c->vlen[tr.a] + c->vlen[tr.b] + wc[i];
            }
        } else {
            for (int i = 0; i < pw.total; i++) c->vcount[list[i].dst] = wc[i];
            full = true;
        }
    }
    c->wave_lists.clear();
    c->wave_pending = 0; c->wave_used = 0; c->wc_used = 0; c->wave_fetched = false;
    if (lengths && full) compute_lengths(c);
}

// ---- lazy views (the SPR search) -----------------------------------------------------------------------------------
// After a move about half of the directed views are out of date (every view whose subtree contains a changed edge), but
// the next scan batch reads only the views around the next few pruning points.  The search therefore only MARKS views
// after a move and recomputes, right before a batch is launched, the stale ones among the views its plan reads plus
// what those depend on -- typically the path from the last moves to the pruning point, a few levels deep, instead of
// ~2n views in a forest as deep as the tree.  Views nobody reads stay stale until ensure_all_views (need_tree).
static inline int2 kid_pair_of(const HostTree &t, int r)
{
    const int a = t.vid(t.back(t.next(r))), b = t.vid(t.back(t.next(t.next(r))));
    return a < b ? make_int2(a, b) : make_int2(b, a);
}

// the adjacency of these nodes changed: their views whose children differ from the last computation are stale, and so
// is every view that contains a stale one (walk towards the views that read it)
void mark_stale_nodes(Ctx *c, const int *nodes, int k)
{
    const HostTree &t = c->tree;
    const int n = t.n, nviews = 4 * n - 6;
    if ((int)c->vstale.size() != nviews) { c->vstale.assign(nviews, 0); c->n_stale = 0; }
    std::vector<int32_t> &stack = c->sc_stack;
    stack.clear();
    for (int i = 0; i < k; i++) {
        const int node = nodes[i];
        if (node <= n) continue;
        for (int s = 0; s < 3; s++) {
            const int r = 3 * node + s, v = t.vid(r);
            if (c->vstale[v]) continue;
            const int2 now = kid_pair_of(t, r), was = c->view_kids[v];
            if (now.x != was.x || now.y != was.y) { c->vstale[v] = 1; c->n_stale++; stack.push_back(r); }
        }
    }
    while (!stack.empty()) {
        const int r = stack.back(); stack.pop_back();
        const int x = t.back(r);                             // view(r) is a child of the views behind the other two slots of x's node
        if (t.is_tip(x)) continue;
        for (int r2 = t.next(x); r2 != x; r2 = t.next(r2)) {
            const int v = t.vid(r2);
            if (!c->vstale[v]) { c->vstale[v] = 1; c->n_stale++; stack.push_back(r2); }
        }
    }
    c->reps.tree_valid = false;
}

// Recompute the stale ones among the views behind the ring slots refs[0..count) and, first, the stale views they are
// built from.  defer: the counts land with settle_views after the caller's next fetch_wave_counts + synchronize.
int ensure_views(Ctx *c, const int32_t *refs, int count, bool defer)
{
    if (c->n_stale == 0 || count == 0) return 0;
    const HostTree &t = c->tree;
    const int nviews = 4 * t.n - 6;
    std::vector<int32_t> &dl = c->sc_dl, &stack = c->sc_stack;
    std::vector<Triple> &stale = c->sc_stale;
    if ((int)dl.size() != nviews || c->dl_dirty) { dl.assign(nviews, 0); c->dl_dirty = false; }
    stale.clear();
    stack.clear();
    int nlevels = 0;
    for (int i = 0; i < count; i++) {
        if (!c->vstale[t.vid(refs[i])]) continue;
        stack.push_back(refs[i]);
        while (!stack.empty()) {
            const int r = stack.back();
            const int v = t.vid(r);
            if (!c->vstale[v]) { stack.pop_back(); continue; }
            const int a = t.back(t.next(r)), b = t.back(t.next(t.next(r)));
            const int va = t.vid(a), vb = t.vid(b);
            const bool sa = c->vstale[va] != 0, sb = c->vstale[vb] != 0;
            if (sa || sb) { if (sa) stack.push_back(a); if (sb) stack.push_back(b); continue; }
            Triple tr; tr.dst = v; tr.a = va; tr.b = vb;
            const int l = std::max(dl[va], dl[vb]) + 1;
            // b = the clean operand when there is one (prefetched); a = the fresh one of the previous level (cached)
            const bool fa = dl[va] != 0, fb = dl[vb] != 0;
            if (!fa && fb) std::swap(tr.a, tr.b);
            else if (fa && fb && dl[tr.a] != l - 1) std::swap(tr.a, tr.b);
            tr.pad = l;
            dl[v] = l;
            if (l > nlevels) nlevels = l;
            stale.push_back(tr);
            c->view_kids[v] = va < vb ? make_int2(va, vb) : make_int2(vb, va);
            c->vstale[v] = 0; c->n_stale--;
            stack.pop_back();
        }
    }
    if (stale.empty()) return 0;
    c->dl_dirty = true;                       // an error return below leaves dl dirty
    c->lazy_lists++; c->lazy_views += (int64_t)stale.size(); c->lazy_levels += nlevels;
    const int rc = submit_stale(c, stale, nlevels, defer, true);
    for (const Triple &tr : stale) dl[tr.dst] = 0;
    c->dl_dirty = false;
    return rc;
}

int ensure_all_views(Ctx *c)
{
    if (c->n_stale == 0) return 0;
    const int n = c->n;
    std::vector<int32_t> refs;
    refs.reserve((size_t)c->n_stale);
    for (int node = n + 1; node <= 2 * n - 2; node++)
        for (int s = 0; s < 3; s++) if (c->vstale[c->tree.vid(3 * node + s)]) refs.push_back(3 * node + s);
    return ensure_views(c, refs.data(), (int)refs.size(), false);
}

// subtree lengths from (all-reduced) mismatch counts, children before parents
void compute_lengths(Ctx *c)
{
    const int nviews = 4 * c->n - 6;
    c->vlen.assign(nviews, 0);
    if (c->sk.on) {      // parsimonyScore[node] of the Sankoff kernel: unweighted u16 sum of per-pattern minima (:491, :547), tips 0
        for (const Triple &tr : c->sched) c->vlen[tr.dst] = c->vcount[tr.dst] & c->sk.sum_mask();
        c->lens_valid = true;
        return;
    }
    for (const Triple &tr : c->sched) c->vlen[tr.dst] = c->vlen[tr.a] + c->vlen[tr.b] + c->vcount[tr.dst];
    c->lens_valid = true;
}

int need_tree(Ctx *c, bool lens)
{
    if (!c) { set_error("null context"); return 1; }
    if (!c->d_views) { set_error("no alignment loaded"); return 1; }
    if (!c->tree_set) { set_error("no tree set"); return 1; }
    if (c->views_stale) {                   // the planes were re-weighted under this tree (mpgpu_set_weights / mpgpu_set_cost_matrix)
        if (cudaSetDevice(c->device) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
        if (int rc = compute_views(c)) return rc;
        if (c->reduces()) compute_lengths(c);
    }
    if (c->n_stale) {                       // an SPR search left views it did not need out of date
        if (cudaSetDevice(c->device) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaSetDevice");
        if (int rc = ensure_all_views(c)) return rc;
    }
    if (lens && !c->lens_valid) { set_error("view lengths not set (sharded context: call mpgpu_set_view_counts)"); return 1; }
    return 0;
}

// device copies of the plan ranges [ops0, n_ops) and [task0, ntasks): the whole plan, or one more piece
static int upload_plan_range(Ctx *c, int ops0, int task0)
{
    ScanPlan &pl = c->plan;
    const int nops = pl.n_ops - ops0, nt = (int)pl.tasks.size() - task0;
    if (nops > 0) {
        MPGPU_CUDA(cudaMemcpyAsync(c->d_offs + ops0, pl.offs.data() + ops0, (size_t)nops * sizeof(ScanOffs), cudaMemcpyHostToDevice, c->stream));
        MPGPU_CUDA(cudaMemcpyAsync(c->d_ctl + ops0, pl.ctl.data() + ops0, (size_t)nops * sizeof(ScanCtl), cudaMemcpyHostToDevice, c->stream));
    }
    const int nsub = (int)pl.sub_tasks.size();           // a split plan (single piece): the sub-tasks sit behind the tasks
    if (nt + nsub > 0) {
        memcpy(pl.tasks_pin.data() + task0, pl.tasks.data() + task0, (size_t)nt * sizeof(ScanTask));
        if (nsub) memcpy(pl.tasks_pin.data() + task0 + nt, pl.sub_tasks.data(), (size_t)nsub * sizeof(ScanTask));
        MPGPU_CUDA(cudaMemcpyAsync(c->d_tasks + task0, pl.tasks_pin.data() + task0, (size_t)(nt + nsub) * sizeof(ScanTask), cudaMemcpyHostToDevice, c->stream));
    }
    return 0;
}

// device buffers for a plan whose arrays were sized by ScanPlanner::begin (upper bounds)
static int reserve_plan(Ctx *c)
{
    ScanPlan &pl = c->plan;
    if (int rc = ensure(c->d_offs, c->offs_cap, pl.offs.size() + 1)) return rc;
    if (int rc = ensure(c->d_ctl, c->ctl_cap, pl.ctl.size() + 1)) return rc;
    if (int rc = ensure(c->d_tasks, c->tasks_cap, std::max(pl.item_cap, (size_t)pl.task_cap) + 1)) return rc;
    const int32_t *before = c->d_counts;
    if (int rc = ensure(c->d_counts, c->counts_cap, (size_t)pl.task_cap + pl.cand_ref.size() + 1)) return rc;
    if (c->d_counts != before) c->counts_dirty = c->counts_cap;         // a fresh allocation is not zero
    return 0;
}

int upload_plan(Ctx *c)
{
    if (int rc = reserve_plan(c)) return rc;
    return upload_plan_range(c, 0, 0);
}

// MPGPU_PROFILE=3: device-side timeline of the single-piece batches of the search (CUDA events between the stream ops)
// next to the host's: where a move's ~70 us go.  Printed at exit.
struct BatchProf {
    bool on = getenv("MPGPU_PROFILE") && atoi(getenv("MPGPU_PROFILE")) >= 3;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool armed = false;
    std::chrono::steady_clock::time_point h0, h1, h2;
    double dev[3] = {0, 0, 0}, host[3] = {0, 0, 0}; long n = 0;
    void rec(int k, cudaStream_t s) { if (!ev[0]) for (auto &e : ev) cudaEventCreate(&e); cudaEventRecord(ev[k], s); }
    ~BatchProf() {
        if (on && n) fprintf(stderr, "[mpgpu profile] %ld single-piece batches: device uploads+wave %.1f us, scan %.1f us, publish %.1f us | host plan+enqueue %.1f us, "
                             "enqueue..flag %.1f us, flag..done(settle+mp) %.1f us\n", n, 1e3 * dev[0] / n, 1e3 * dev[1] / n, 1e3 * dev[2] / n,
                             1e6 * host[0] / n, 1e6 * host[1] / n, 1e6 * host[2] / n);
    }
};
static BatchProf g_bp;

// MPGPU_PROFILE=4: host timeline of mpgpu_scan_visits (the e2e sweep), averaged over the calls and printed at exit: when each
// stage of the call was reached, in us since the call began
struct SweepProf {
    bool on = getenv("MPGPU_PROFILE") && atoi(getenv("MPGPU_PROFILE")) == 4;
    static const int K = 16;
    const char *name[K] = {nullptr};
    double sum[K] = {0}; long cnt[K] = {0};
    std::chrono::steady_clock::time_point t0;
    cudaEvent_t ev[8] = {nullptr}; const char *ename[8] = {nullptr}; double esum[8] = {0}; long ecnt[8] = {0}; int nev = 0;
    void begin() { if (on) { t0 = std::chrono::steady_clock::now(); nev = 0; } }
    void rec(cudaStream_t st, const char *what) {          // a device-side timestamp in stream order
        if (!on || nev >= 8) return;
        if (!ev[nev]) cudaEventCreate(&ev[nev]);
        ename[nev] = what; cudaEventRecord(ev[nev], st); nev++;
    }
    void collect() {
        if (!on || nev < 2) return;
        cudaEventSynchronize(ev[nev - 1]);
        for (int k = 1; k < nev; k++) { float ms = 0; if (cudaEventElapsedTime(&ms, ev[0], ev[k]) == cudaSuccess) { esum[k] += ms; ecnt[k]++; } }
        nev = 0;
    }
    void mark(int k, const char *what) {
        if (!on || k >= K) return;
        name[k] = what; sum[k] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); cnt[k]++;
    }
    ~SweepProf() {
        if (!on) return;
        for (int k = 0; k < K; k++) if (cnt[k]) fprintf(stderr, "[mpgpu sweep timeline] %-34s %8.1f us (n = %ld)\n", name[k], 1e6 * sum[k] / cnt[k], cnt[k]);
        for (int k = 1; k < 8; k++) if (ecnt[k]) fprintf(stderr, "[mpgpu sweep timeline] device: %-26s %8.1f us after the first stream op (n = %ld)\n", ename[k], 1e3 * esum[k] / ecnt[k], ecnt[k]);
    }
};
static SweepProf g_sw;

// d_counts[0, n) = 0 before a scan.  k_publish leaves the range it read zeroed, so in the search loop this is usually
// nothing; any other reader leaves the counters dirty and the next scan pays one memset.
int zero_counts(Ctx *c, size_t n)
{
    if (c->counts_dirty) {
        const size_t m = std::min(c->counts_dirty, c->counts_cap);
        MPGPU_CUDA(cudaMemsetAsync(c->d_counts, 0, m * sizeof(int32_t), c->stream));
    }
    c->counts_dirty = n;                    // the scan about to be launched writes [0, n)
    return 0;
}

// counts layout: [0, task_cap) joined-edge counts per task slot, then one count per candidate
int run_scan(Ctx *c)
{
    ScanPlan &pl = c->plan;
    if (c->sk.on) return sk_run_scan(c);
    const size_t nout = (size_t)pl.task_cap + pl.n_cand;
    if (int rc = zero_counts(c, nout + 1)) return rc;
    if (int rc = launch_scan(c, 0, (int)pl.tasks.size(), pl.max_slot)) return rc;
    if (c->shard_count > 1 && c->reduces()) return shard_sum(c, c->d_counts, (int64_t)nout);
    return 0;
}

// The read-back of a small batch without a copy engine and without a stream synchronize: one block copies the scan
// counts and the counts of the view updates in flight to mapped page-locked memory, puts the device counters back to
// zero (the next batch needs no memset) and writes the flag word last; the host spins on the flag.
__global__ void __launch_bounds__(512) k_publish(int32_t *__restrict__ counts, int nout, uint32_t *__restrict__ wcount, int nwc,
                                                 int32_t *__restrict__ host_counts, uint32_t *__restrict__ host_wc,
                                                 volatile uint32_t *host_flag, uint32_t epoch, unsigned int *ticket)
{
    // a large read-back (a whole sweep: ~15 k counters) is spread over a few blocks -- one block's posted writes over PCIe took 19 us
    // for 61 KB -- and the block that draws the last ticket raises the flag (every block fences its writes before it draws)
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int i = tid; i < nout; i += nth) { host_counts[i] = __ldcg(counts + i); counts[i] = 0; }
    for (int i = tid; i < nwc; i += nth) { host_wc[i] = __ldcg(wcount + i); wcount[i] = 0; }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (gridDim.x == 1) { *host_flag = epoch; return; }
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
            *ticket = 0;
            __threadfence_system();
            *host_flag = epoch;
        }
    }
}

static int ensure_h_counts(Ctx *c, size_t nout)
{
    if (nout * sizeof(int32_t) > c->h_counts_cap) {
        if (c->h_counts) cudaFreeHost(c->h_counts);
        c->h_counts = nullptr; c->h_counts_cap = 0;
        const size_t want = (nout + nout / 2 + 1024) * sizeof(int32_t);
        MPGPU_CUDA(cudaHostAlloc((void **)&c->h_counts, want, cudaHostAllocMapped));        // pinned: the read-back is on the e2e path
        c->h_counts_cap = want;
    }
    if (!c->h_flag) {
        MPGPU_CUDA(cudaHostAlloc((void **)&c->h_flag, 64, cudaHostAllocMapped));
        *c->h_flag = 0;
    }
    return 0;
}

static const int kPublishMax = 16384;     // counters one k_publish block moves; larger read-backs take the copy engine

// spin on the flag word; the stream is queried now and then so that a failed launch cannot hang the host
static int wait_flag(Ctx *c)
{
    if (g_bp.armed) g_bp.h1 = std::chrono::steady_clock::now();
    if (c->wave_pending) c->wave_fetched = true;
    c->counts_dirty = 0;
    volatile uint32_t *flag = c->h_flag;
    for (uint32_t spins = 1; *flag != c->flag_epoch; spins++) {
        if ((spins & 0xFFFF) == 0) {
            const cudaError_t e = cudaStreamQuery(c->stream);
            if (e != cudaSuccess && e != cudaErrorNotReady) return cuda_fail(e, "scan batch");
            if (e == cudaSuccess && *flag != c->flag_epoch) { set_error("k_publish finished without raising its flag"); return 1; }
        }
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
    if (g_bp.armed) g_bp.h2 = std::chrono::steady_clock::now();
    return 0;
}

int launch_publish(Ctx *c, int nout)
{
    if (int rc = ensure_h_counts(c, (size_t)nout)) return rc;
    c->flag_epoch++;
    if (c->flag_epoch == 0) c->flag_epoch = 1;
    if (!c->d_done) { MPGPU_CUDA(cudaMalloc((void **)&c->d_done, 64)); MPGPU_CUDA(cudaMemsetAsync(c->d_done, 0, 64, c->stream)); }
    static const int max_blocks = getenv("MPGPU_PUBLISH_BLOCKS") ? std::max(1, atoi(getenv("MPGPU_PUBLISH_BLOCKS"))) : 16;   // tuning knob
    const int blocks = std::max(1, std::min(max_blocks, (nout + (int)c->wc_used + 1023) / 1024));
    k_publish<<<blocks, 512, 0, c->stream>>>(c->d_counts, nout, c->d_wcount, (int)c->wc_used, c->h_counts, c->wcount_pin.data(),
                                             c->h_flag, c->flag_epoch, c->d_done + 1);      // (word 0 is the fused publish's ticket)
    c->launches++;
    MPGPU_CUDA(cudaGetLastError());
    if (g_bp.armed) g_bp.rec(3, c->stream);
    return wait_flag(c);
}

int finish_scan(Ctx *c, int32_t *visit_begin, uint32_t *mp, int32_t *cand_ref, int32_t *cand_prune, int capacity)
{
    ScanPlan &pl = c->plan;
    if (c->sk.on) return sk_finish_scan(c, visit_begin, mp, cand_ref, cand_prune, capacity);
    if (pl.n_cand > capacity) { set_error("candidate capacity too small"); return 1; }
    const size_t nout = (size_t)pl.task_cap + pl.n_cand;
    static const bool no_publish = getenv("MPGPU_NO_PUBLISH") != nullptr;
    // what does not depend on the counts goes to the caller while the device is still scoring
    if (visit_begin) memcpy(visit_begin, pl.visit_begin.data(), pl.visit_begin.size() * sizeof(int32_t));
    if (cand_ref) memcpy(cand_ref, pl.cand_ref.data(), pl.n_cand * sizeof(int32_t));
    if (cand_prune) memcpy(cand_prune, pl.cand_prune.data(), pl.n_cand * sizeof(int32_t));
    g_sw.mark(10, "finish: index arrays copied");
    if (c->pub_inflight) {                 // the scan's last block publishes (latency path)
        c->pub_inflight = false;
        if (g_bp.armed) g_bp.rec(3, c->stream);
        if (int rc = wait_flag(c)) return rc;
    } else if (!no_publish && c->shard_count == 1 && nout + c->wc_used <= (size_t)kPublishMax && (c->wc_used == 0 || c->wcount_pin.data())) {
        if (int rc = launch_publish(c, (int)nout)) return rc;
    } else {
        if (int rc = ensure_h_counts(c, nout)) return rc;
        MPGPU_CUDA(cudaMemcpyAsync(c->h_counts, c->d_counts, nout * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        if (int rc = fetch_wave_counts(c)) return rc;
        MPGPU_CUDA(cudaStreamSynchronize(c->stream));
        c->counts_dirty = nout + 1;
    }
    g_sw.mark(11, "finish: counts on the host");
    g_sw.rec(c->stream, "publish done"); g_sw.collect();
    settle_views(c, true);
    if (!c->lens_valid) { set_error("view lengths not set"); return 1; }
    const size_t ntasks = pl.tasks.size();
    pl.task_const.resize(ntasks);
    for (size_t k = 0; k < ntasks; k++)                 // len(S) + len(D1) + len(D2)
        pl.task_const[k] = c->vlen[pl.task_vids[3 * k]] + c->vlen[pl.task_vids[3 * k + 1]] + c->vlen[pl.task_vids[3 * k + 2]];
    const int32_t *base = c->h_counts, *cnt = c->h_counts + pl.task_cap;
    const int32_t *ctask = pl.cand_task.data();
    const uint32_t *tconst = pl.task_const.data();
    for (int j = 0; j < pl.n_cand; j++) {
        const int ti = ctask[j];
        mp[j] = tconst[ti] + (uint32_t)base[ti] + (uint32_t)cnt[j];
    }
    g_sw.mark(12, "finish: scores assembled");
    if (g_bp.armed && g_bp.ev[3] && cudaEventSynchronize(g_bp.ev[3]) == cudaSuccess) {
        const auto h3 = std::chrono::steady_clock::now();
        float ms;
        for (int k = 0; k < 3; k++) if (cudaEventElapsedTime(&ms, g_bp.ev[k], g_bp.ev[k + 1]) == cudaSuccess) g_bp.dev[k] += ms;
        g_bp.host[0] += std::chrono::duration<double>(g_bp.h1 - g_bp.h0).count();
        g_bp.host[1] += std::chrono::duration<double>(g_bp.h2 - g_bp.h1).count();
        g_bp.host[2] += std::chrono::duration<double>(h3 - g_bp.h2).count();
        g_bp.n++;
    }
    g_bp.armed = false;
    return 0;
}

// Plan, upload and launch a batch of visits in pieces: while the device scores the first visits
// the host enumerates the next ones (the enumeration is the largest host term of the e2e path).
int scan_batch_pipelined(Ctx *c, const int32_t *order, int first, int count, int mintrav, int maxtrav)
{
    ScanPlan &pl = c->plan;
    const uint8_t *vstale = c->n_stale ? c->vstale.data() : nullptr;      // lazy views: the planner notes the stale views it reads
    // the planner's slot table lives in the context: the SPR search patches the five nodes a move touches instead of paying O(n)
    // per batch (set_tree and the stepwise phase drop it)
    const uint32_t tab_stride = c->sk.on ? (uint32_t)(c->sk.vstride / 4) : (uint32_t)(c->view_stride / (c->S < 4 ? c->S : 4));
    if (!c->ref_valid || c->ref_vstride != tab_stride || c->ref_table.size() != (size_t)3 * (2 * c->n - 1)) {
        scan_ref_build(c->tree, tab_stride, c->ref_table);
        c->ref_vstride = tab_stride; c->ref_valid = true;
    }
    if (c->sk.on) {          // one piece: the per-(candidate, segment) output is sized from the finished plan
        ScanPlanner planner;
        if (int rc = planner.begin(c->tree, order, first, count, mintrav, maxtrav, (uint32_t)(c->sk.vstride / 4), pl, false, vstale, 0, c->ref_table.data())) return rc;
        planner.add(0, count);
        planner.finish();
        if (int rc = ensure_views(c, pl.need_refs.data(), (int)pl.need_refs.size(), false)) return rc;
        if (int rc = upload_plan(c)) return rc;
        return sk_run_scan(c);
    }
    const uint32_t vstride_vec = (uint32_t)(c->view_stride / (c->S < 4 ? c->S : 4));
    int pieces = count >= 32 ? 2 : 1;        // measured on B200 (C2 sweep, staged pieces, r02): 2: 0.248 ms, 3: 0.249, 4: 0.264, 6: 0.304, 8: 0.344 e2e (copy-engine pieces: 2: 0.267)
    if (const char *e = getenv("MPGPU_SCAN_PIECES")) { int v = atoi(e); if (v >= 1) pieces = v; }
    // tuning knob MPGPU_SCAN_SPLITS="p1,p2,...": piece boundaries in percent of the visits (pieces = boundaries + 1), large batches only
    static const std::vector<int> split_pct = []() {
        std::vector<int> v;
        if (const char *e = getenv("MPGPU_SCAN_SPLITS")) { for (const char *q = e; *q;) { v.push_back(atoi(q)); while (*q && *q != ',') q++; if (*q == ',') q++; } }
        return v;
    }();
    if (!split_pct.empty() && count >= 32) pieces = (int)split_pct.size() + 1;
    // a small batch is latency-bound (one warp per task and chunk walks ~100 ops): cut its tasks into sub-tasks
    static const int split_depth = getenv("MPGPU_SPLIT_DEPTH") ? atoi(getenv("MPGPU_SPLIT_DEPTH")) : 3;
    static const int split_max = getenv("MPGPU_SPLIT_MAXCOUNT") ? atoi(getenv("MPGPU_SPLIT_MAXCOUNT")) : 16;
    const int sd = pieces == 1 && count <= split_max && split_depth > 0 ? split_depth : 0;
    g_bp.armed = g_bp.on && pieces == 1;
    if (g_bp.armed) { g_bp.h0 = std::chrono::steady_clock::now(); g_bp.rec(0, c->stream); }
    ScanPlanner planner;
    if (int rc = planner.begin(c->tree, order, first, count, mintrav, maxtrav, vstride_vec, pl, false, vstale, sd, c->ref_table.data())) return rc;
    if (int rc = reserve_plan(c)) return rc;
    if (int rc = zero_counts(c, (size_t)pl.task_cap + pl.cand_ref.size() + 1)) return rc;
    g_sw.mark(1, "planner begun, counters zeroed");
    g_sw.rec(c->stream, "start");
    int v0 = 0;
    for (int k = 0; k < pieces; k++) {
        // early pieces are smaller: the device should get going as soon as possible
        // (tuning knob MPGPU_SCAN_FIRST_PCT: share of the visits in the first of two pieces; MPGPU_SCAN_EQUAL: equal pieces)
        static const int first_pct = getenv("MPGPU_SCAN_FIRST_PCT") ? atoi(getenv("MPGPU_SCAN_FIRST_PCT")) : 0;
        static const bool equal_pieces = getenv("MPGPU_SCAN_EQUAL") != nullptr;
        int v1 = k == pieces - 1 ? count : std::min(count, v0 + std::max(1, (int)((long long)count * (k + 1) / (pieces * (pieces + 1) / 2))));
        if (k < pieces - 1 && equal_pieces) v1 = std::min(count, std::max(v0 + 1, (int)((long long)count * (k + 1) / pieces)));
        if (k == 0 && pieces == 2 && first_pct > 0 && first_pct < 100) v1 = std::min(count - 1, std::max(1, (int)((long long)count * first_pct / 100)));
        if (k < pieces - 1 && (int)split_pct.size() == pieces - 1) v1 = std::min(count, std::max(v0 + 1, (int)((long long)count * split_pct[k] / 100)));
        const int ops0 = pl.n_ops, task0 = (int)pl.tasks.size();
        if (!vstale && !sd && v1 - v0 >= 96) planner.add_parallel(v0, v1, plan_threads());     // a big piece of a big batch: several host threads
        else planner.add(v0, v1);
        if (pieces <= 4 && g_sw.on) g_sw.mark(2 + 2 * k, k == 0 ? "piece 0 enumerated" : (k == 1 ? "piece 1 enumerated" : "piece k enumerated"));
        if (sd) planner.split();
        // latency path (one piece, one shard, a small plan): the plan rides to the device with the wave launch (or a k_stage
        // launch) and the scan's last block publishes the counts -- no copy engine and no stream synchronize in the step
        static const bool no_lean = getenv("MPGPU_NO_LEAN") != nullptr || getenv("MPGPU_NO_PUBLISH") != nullptr;
        const int nops_piece = pl.n_ops - ops0;
        const size_t nt = pl.tasks.size(), nsub = pl.sub_tasks.size();
        const size_t plan_bytes = (size_t)nops_piece * 16 + (nt - task0 + nsub) * sizeof(ScanTask);
        // staged pieces: the streams of this piece go from mapped host memory to their device arrays through a kernel (riding on
        // the wave launch when there is one) instead of three copy-engine hops; any number of pieces, one shard or many
        const bool staged = !no_lean && plan_bytes <= 512 * 1024;
        const bool lean = staged && pieces == 1 && c->shard_count == 1 && plan_bytes <= 96 * 1024;
        if (staged) {
            memcpy(pl.tasks_pin.data() + task0, pl.tasks.data() + task0, (nt - task0) * sizeof(ScanTask));
            if (nsub) memcpy(pl.tasks_pin.data() + nt, pl.sub_tasks.data(), nsub * sizeof(ScanTask));
            const int o0 = ops0 & ~1;                     // 16-byte units: an odd first op re-copies its (identical) predecessor
            const int n16 = (pl.n_ops - o0 + 1) / 2;
            StageArgs &st = c->stage_req;
            st.src[0] = reinterpret_cast<const uint4 *>(pl.tasks_pin.data() + task0); st.dst[0] = reinterpret_cast<uint4 *>(c->d_tasks + task0); st.n[0] = (int)(nt - task0 + nsub) * 2;
            st.src[1] = reinterpret_cast<const uint4 *>(pl.offs.data() + o0); st.dst[1] = reinterpret_cast<uint4 *>(c->d_offs + o0); st.n[1] = n16;
            st.src[2] = reinterpret_cast<const uint4 *>(pl.ctl.data() + o0); st.dst[2] = reinterpret_cast<uint4 *>(c->d_ctl + o0); st.n[2] = n16;
            c->stage_pending = true;
            if (plan_bytes > 64 * 1024) { if (int rc = launch_stage(c)) return rc; }     // too much for the wave launch's single extra CTA
        }
        if (!pl.need_refs.empty()) {
            if (int rc = ensure_views(c, pl.need_refs.data(), (int)pl.need_refs.size(), true)) return rc;
            pl.need_refs.clear();
        }
        if (staged) {
            if (int rc = launch_stage(c)) return rc;            // no wave was launched: the plan goes alone
            const size_t nout = (size_t)pl.task_cap + pl.n_cand;
            if (lean && nout + c->wc_used <= (size_t)kPublishMax) {
                if (int rc = ensure_h_counts(c, nout)) return rc;
                if (!c->d_done) { MPGPU_CUDA(cudaMalloc((void **)&c->d_done, 64)); MPGPU_CUDA(cudaMemsetAsync(c->d_done, 0, 64, c->stream)); }
                c->pub_request = true; c->pub_nout = (int)nout;
            }
        } else if (int rc = upload_plan_range(c, ops0, task0)) return rc;
        if (g_bp.armed) g_bp.rec(1, c->stream);
        if (!pl.sub_tasks.empty()) { if (int rc = launch_scan(c, (int)pl.tasks.size(), (int)pl.sub_tasks.size(), pl.max_slot)) return rc; }
        else if (int rc = launch_scan(c, task0, (int)pl.tasks.size() - task0, pl.max_slot)) return rc;
        c->pub_request = false;                 // (nothing was launched: finish_scan publishes on its own)
        if (pieces <= 4) g_sw.rec(c->stream, k == 0 ? "scan of piece 0 done" : (k == 1 ? "scan of piece 1 done" : "scan of piece k done"));
        if (pieces <= 4) g_sw.mark(3 + 2 * k, k == 0 ? "piece 0 staged + launched" : (k == 1 ? "piece 1 staged + launched" : "piece k staged + launched"));
        v0 = v1;
        if (v0 >= count) break;
    }
    planner.finish();
    if (g_bp.armed) g_bp.rec(2, c->stream);
    if (c->shard_count > 1 && c->reduces()) return shard_sum(c, c->d_counts, (int64_t)pl.task_cap + pl.n_cand);
    return 0;
}

// Per-site mismatch counters of the current tree, bit-sliced, in d_bitcnt[nbits][Wl]: child-view
// pairs of the n-2 inner views facing tr->start, plus the start edge (storePerSiteNodeScores :294).
int compute_site_counters(Ctx *c, int nbits)
{
    const HostTree &t = c->tree;
    const int n = c->n;
    std::vector<int32_t> &order = c->sc_order;
    visit_order(t, order);
    // staging in a page-locked buffer of the context: the upload is asynchronous and nothing has to be waited for here
    // (two halves used alternately: the previous upload may still be in flight when the next tree arrives)
    const size_t need = (size_t)2 * (n - 1);
    if (!c->pairs_pin.reserve(2 * need + 16)) { set_error("pinned allocation failed"); return 1; }
    int32_t *pairs = c->pairs_pin.data() + (c->pairs_flip ? need : 0);
    c->pairs_flip ^= 1;
    size_t k = 0;
    for (int i = n + 1; i <= 2 * n - 2; i++) {
        const int r = order[i];
        pairs[k++] = t.vid(t.back(t.next(r)));
        pairs[k++] = t.vid(t.back(t.next(t.next(r))));
    }
    pairs[k++] = t.vid(3); pairs[k++] = t.vid(t.back(3));
    const int npairs = (int)k / 2;
    if (c->n_stale) {                          // inside an SPR search (lazy views): the views facing tr->start must be current
        std::vector<int32_t> &refs = c->sc_refs;
        refs.clear();
        for (int i = n + 1; i <= 2 * n - 2; i++) { const int r = order[i]; refs.push_back(t.back(t.next(r))); refs.push_back(t.back(t.next(t.next(r)))); }
        refs.push_back(t.back(3));
        if (int rc = ensure_views(c, refs.data(), (int)refs.size(), true)) return rc;
    }
    if (int rc = ensure(c->d_pairs, c->pairs_cap, k)) return rc;
    if (int rc = ensure(c->d_bitcnt, c->bitcnt_cap, (size_t)16 * c->Wl)) return rc;
    MPGPU_CUDA(cudaMemcpyAsync(c->d_pairs, pairs, k * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    if (int rc = launch_site_counters(c, npairs, nbits)) return rc;
    return 0;
}

// Site of each reported pattern: the reference walks site += aliaswgt[ptn] over the prefix
// (pllComputePatternParsimony :3380-3390); -1 = beyond the reference's padded plane.
int ensure_ptn_site(Ctx *c)
{
    if (c->ptn_site_valid) return 0;
    const int upper = c->sort_alignment ? c->n_inf : c->P;
    std::vector<int64_t> ptn_site(upper > 0 ? upper : 1);
    int64_t site = 0;
    bool ident = true;
    for (int i = 0; i < upper; i++) {
        ptn_site[i] = site < (int64_t)c->ref_words * 32 ? site : -1;
        if (ptn_site[i] != i) ident = false;
        site += c->weights[i];
    }
    c->ptn_identity = ident;
    if (int rc = ensure(c->d_ptn_site, c->ptn_site_cap, ptn_site.size())) return rc;
    MPGPU_CUDA(cudaMemcpyAsync(c->d_ptn_site, ptn_site.data(), ptn_site.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    MPGPU_CUDA(cudaStreamSynchronize(c->stream));
    c->ptn_site_valid = true;
    return 0;
}

}  // namespace mpgpu

using namespace mpgpu;


extern "C" {

const char *mpgpu_last_error(void) { return g_error.c_str(); }

int mpgpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int mpgpu_create(mpgpu_ctx **out, int device, void *stream, int shard_rank, int shard_count)
{
    if (!out) { set_error("null out pointer"); return 1; }
    *out = nullptr;
    if (shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count) { set_error("bad shard arguments"); return 1; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device available: libmpgpu has no CPU fallback");
        return 3;
    }
    if (device < 0 || device >= ndev) { set_error("bad device index"); return 1; }
    MPGPU_CUDA(cudaSetDevice(device));
    mpgpu_ctx *c = new mpgpu_ctx();
    c->device = device; c->shard_rank = shard_rank; c->shard_count = shard_count;
    if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
    else { MPGPU_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    MPGPU_CUDA(cudaMalloc((void **)&c->d_scalar, 64 * sizeof(uint32_t)));
    *out = c;
    return 0;
}

int mpgpu_destroy(mpgpu_ctx *c)
{
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    free_alignment(c);
    sk_free(c);
    if (c->d_triples) cudaFree(c->d_triples);
    if (c->d_wave) cudaFree(c->d_wave);
    if (c->d_wcount) cudaFree(c->d_wcount);
    if (c->d_scalar) cudaFree(c->d_scalar);
    if (c->d_offs) cudaFree(c->d_offs);
    if (c->d_ctl) cudaFree(c->d_ctl);
    if (c->d_tasks) cudaFree(c->d_tasks);
    if (c->d_counts) cudaFree(c->d_counts);
    if (c->h_counts) cudaFreeHost(c->h_counts);
    if (c->h_flag) cudaFreeHost(c->h_flag);
    if (c->d_done) cudaFree(c->d_done);
    if (c->d_bitcnt) cudaFree(c->d_bitcnt);
    if (c->d_pairs) cudaFree(c->d_pairs);
    if (c->d_ptn) cudaFree(c->d_ptn);
    if (c->d_ptn_site) cudaFree(c->d_ptn_site);
    free_reps(c);
    peer_free(c);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int mpgpu_set_allreduce(mpgpu_ctx *c, mpgpu_allreduce_fn fn, void *user)
{
    if (!c) { set_error("null context"); return 1; }
    c->allreduce = fn; c->allreduce_user = user;
    if (fn && c->tree_set && !c->lens_valid) {            // the partial counts of the resident tree can be completed now
        MPGPU_CUDA(cudaSetDevice(c->device));
        if (int rc = compute_views(c)) return rc;
        compute_lengths(c);
    }
    return 0;
}

void *mpgpu_stream(mpgpu_ctx *c) { return c ? (void *)c->stream : nullptr; }
int mpgpu_synchronize(mpgpu_ctx *c) { if (!c) return 1; MPGPU_CUDA(cudaStreamSynchronize(c->stream)); return 0; }
int64_t mpgpu_launch_count(mpgpu_ctx *c) { return c ? c->launches : 0; }

int mpgpu_load_alignment(mpgpu_ctx *c, int ntaxa, int npatterns, int datatype,
                         const uint8_t *yvector, const int32_t *aliaswgt, int sort_alignment)
{
    if (!c || !yvector || !aliaswgt) { set_error("null argument"); return 1; }
    const int S = states_of(datatype);
    if (S < 0) { set_error("unsupported data type"); return 1; }
    if (ntaxa < 4 || npatterns < 1) { set_error("need at least 4 taxa and 1 pattern"); return 1; }
    MPGPU_CUDA(cudaSetDevice(c->device));
    free_alignment(c);
    free_reps(c);
    c->sk.on = false;                       // a new alignment starts in Fitch mode (resetGlobalParamOnNewAln, sprparsimony.cpp:143): the
                                            // cost matrix's segment bounds belong to the old pattern set
    c->n = ntaxa; c->P = npatterns; c->datatype = datatype; c->S = S; c->sort_alignment = sort_alignment;
    c->weights.assign(aliaswgt, aliaswgt + npatterns);
    for (int i = 0; i < npatterns; i++) if (aliaswgt[i] < 0) { set_error("negative pattern weight"); return 1; }
    find_informative(c, yvector);
    MPGPU_CUDA(cudaMalloc((void **)&c->d_codes, (size_t)ntaxa * npatterns));
    MPGPU_CUDA(cudaMemcpyAsync(c->d_codes, yvector, (size_t)ntaxa * npatterns, cudaMemcpyHostToDevice, c->stream));
    c->Wl = 0; c->glob_words = 0;
    return build_planes(c, true);
}

int mpgpu_set_weights(mpgpu_ctx *c, const int32_t *aliaswgt)
{
    if (!c || !aliaswgt) { set_error("null argument"); return 1; }
    if (!c->d_codes) { set_error("no alignment loaded"); return 1; }
    MPGPU_CUDA(cudaSetDevice(c->device));
    for (int i = 0; i < c->P; i++) if (aliaswgt[i] < 0) { set_error("negative pattern weight"); return 1; }
    c->weights.assign(aliaswgt, aliaswgt + c->P);
    if (int rc = build_planes(c, false)) return rc;
    if (c->tree_set) { c->views_stale = true; c->lens_valid = false; c->kids_valid = false; }   // views depend on the planes: recomputed on first use
    return 0;
}

int mpgpu_get_layout(mpgpu_ctx *c, int *states, int *ref_words, int *shard_words, int *n_informative, int64_t *n_sites)
{
    if (!c || !c->d_views) { set_error("no alignment loaded"); return 1; }
    if (states) *states = c->S;
    if (ref_words) *ref_words = c->ref_words;
    if (shard_words) *shard_words = c->Wl;
    if (n_informative) *n_informative = c->n_inf;
    if (n_sites) *n_sites = c->n_sites;
    return 0;
}

static int copy_view_ref_layout(mpgpu_ctx *c, int vid, uint32_t *out)
{
    if (c->shard_count != 1) { set_error("plane read-back is single-shard only"); return 1; }
    if (c->sk.on && vid >= c->n) { set_error("Fitch planes of inner views do not exist under -cost: use mpgpu_sankoff_view"); return 1; }
    std::vector<uint32_t> tmp(c->view_stride);
    MPGPU_CUDA(cudaMemcpyAsync(tmp.data(), c->d_views + (size_t)vid * c->view_stride, c->view_stride * sizeof(uint32_t),
                               cudaMemcpyDeviceToHost, c->stream));
    MPGPU_CUDA(cudaStreamSynchronize(c->stream));
    const int SG = c->S < 4 ? c->S : 4;
    for (int s = 0; s < c->S; s++)
        for (int w = 0; w < c->ref_words; w++)
            out[(size_t)s * c->ref_words + w] = w < c->Wl ? tmp[(size_t)(s / SG) * c->Wl * SG + (size_t)w * SG + (s % SG)] : 0xFFFFFFFFu;
    return 0;
}

int mpgpu_get_tip_planes(mpgpu_ctx *c, int tip, uint32_t *out)
{
    if (!c || !c->d_views || !out) { set_error("no alignment loaded"); return 1; }
    if (tip < 1 || tip > c->n) { set_error("tip out of range"); return 1; }
    MPGPU_CUDA(cudaSetDevice(c->device));
    return copy_view_ref_layout(c, tip - 1, out);
}

int mpgpu_set_tree(mpgpu_ctx *c, const int32_t *back_node, const int32_t *back_slot)
{
    return mpgpu::set_tree_impl(c, back_node, back_slot, false);
}

}  // extern "C"

namespace mpgpu {
// mpgpu_set_tree; want_start_edge = also bring back the mismatch count across the edge at tip 1 with the view counts, so
// that the score of the tree (evaluateParsimony at tr->start) costs no second round trip (start_edge_mis / start_edge_valid)
int set_tree_impl(Ctx *c, const int32_t *back_node, const int32_t *back_slot, bool want_start_edge)
{
    if (!c || !back_node || !back_slot) { set_error("null argument"); return 1; }
    if (!c->d_views) { set_error("no alignment loaded"); return 1; }
    MPGPU_CUDA(cudaSetDevice(c->device));
    const int n = c->n, len = 3 * (2 * n - 1);
    // validate into a temporary and commit only on success: every slot of a complete unrooted binary tree is hooked
    // symmetrically, and the table is ONE tree (a symmetric table can still be cyclic or disconnected, which would send
    // the schedule builder's traversal round in circles): a walk from tip 1 must reach all 2n-2 nodes, each once.
    HostTree t;
    t.n = n; t.bn.assign(back_node, back_node + len); t.bs.assign(back_slot, back_slot + len);
    const char *bad = nullptr;
    for (int node = 1; node <= 2 * n - 2 && !bad; node++) {
        const int ns = node <= n ? 1 : 3;
        for (int s = 0; s < ns; s++) {
            const int r = 3 * node + s;
            const int bnode = t.bn[r], bslot = t.bs[r];
            if (bnode < 1 || bnode > 2 * n - 2 || bslot < 0 || bslot > (bnode <= n ? 0 : 2)) { bad = "ring table: dangling or out-of-range back pointer"; break; }
            const int b = 3 * bnode + bslot;
            if (t.bn[b] != node || t.bs[b] != s) { bad = "ring table: back pointers are not symmetric"; break; }
        }
    }
    if (!bad) {
        std::vector<uint8_t> seen((size_t)2 * n - 1, 0);
        std::vector<int> stack;
        int reached = 1;
        seen[1] = 1;
        stack.push_back(t.back(3));
        while (!stack.empty() && !bad) {
            const int x = stack.back(); stack.pop_back();       // ref by which the walk enters node x/3
            const int node = x / 3;
            if (seen[node]) { bad = "ring table: not a tree (a node is reachable along two paths)"; break; }
            seen[node] = 1; reached++;
            if (node > n) { stack.push_back(t.back(t.next(x))); stack.push_back(t.back(t.next(t.next(x)))); }
        }
        if (!bad && reached != 2 * n - 2) bad = "ring table: not connected (tip 1 does not reach every node)";
    }
    if (bad) { c->tree_set = false; c->lens_valid = false; c->kids_valid = false; set_error(bad); return 1; }
    c->tree.n = n; c->tree.bn.swap(t.bn); c->tree.bs.swap(t.bs);
    c->tree_set = true; c->lens_valid = false; c->ref_valid = false;
    if (int rc = compute_views(c, want_start_edge)) return rc;
    if (c->reduces()) compute_lengths(c);
    return 0;
}
}  // namespace mpgpu

extern "C" {

int mpgpu_get_view_counts_partial(mpgpu_ctx *c, uint32_t *counts)
{
    if (int rc = need_tree(c, false)) return rc;
    memcpy(counts, c->vcount.data(), c->vcount.size() * sizeof(uint32_t));
    return 0;
}

int mpgpu_set_view_counts(mpgpu_ctx *c, const uint32_t *counts)
{
    if (int rc = need_tree(c, false)) return rc;
    c->vcount.assign(counts, counts + (4 * c->n - 6));
    compute_lengths(c);
    return 0;
}

static int check_ref(mpgpu_ctx *c, int node, int slot)
{
    if (node < 1 || node > 2 * c->n - 2 || slot < 0 || slot > (node <= c->n ? 0 : 2)) { set_error("bad (node,slot)"); return 1; }
    return 0;
}

int mpgpu_view_length(mpgpu_ctx *c, int node, int slot, uint32_t *length)
{
    if (int rc = need_tree(c, true)) return rc;
    if (int rc = check_ref(c, node, slot)) return rc;
    *length = c->vlen[c->tree.vid(3 * node + slot)];
    return 0;
}

int mpgpu_get_view_planes(mpgpu_ctx *c, int node, int slot, uint32_t *out)
{
    if (int rc = need_tree(c, false)) return rc;
    if (int rc = check_ref(c, node, slot)) return rc;
    MPGPU_CUDA(cudaSetDevice(c->device));
    return copy_view_ref_layout(c, c->tree.vid(3 * node + slot), out);
}

static int edge_mismatch(mpgpu_ctx *c, int node, int slot, uint32_t *count, bool reduce)
{
    if (int rc = need_tree(c, false)) return rc;
    if (c->sk.on) { set_error("edge mismatch counts are a Fitch quantity (a cost matrix is set)"); return 1; }
    if (int rc = check_ref(c, node, slot)) return rc;
    MPGPU_CUDA(cudaSetDevice(c->device));
    const int r = 3 * node + slot;
    MPGPU_CUDA(cudaMemsetAsync(c->d_scalar, 0, sizeof(uint32_t), c->stream));
    if (int rc = launch_edge_mismatch(c, c->tree.vid(r), c->tree.vid(c->tree.back(r)), c->d_scalar)) return rc;
    if (reduce) { if (int rc = shard_sum(c, c->d_scalar, 1)) return rc; }
    MPGPU_CUDA(cudaMemcpyAsync(count, c->d_scalar, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    MPGPU_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int mpgpu_edge_mismatch_partial(mpgpu_ctx *c, int node, int slot, uint32_t *count)
{
    return edge_mismatch(c, node, slot, count, false);
}

int mpgpu_tree_score(mpgpu_ctx *c, uint32_t *score)
{
    if (int rc = need_tree(c, c && c->reduces())) return rc;
    if (c->sk.on) { MPGPU_CUDA(cudaSetDevice(c->device)); return sk_tree_score(c, 3, score); }
    uint32_t mis = 0;
    if (int rc = edge_mismatch(c, 1, 0, &mis, c->shard_count > 1 && c->reduces())) return rc;
    if (c->reduces()) *score = mis + c->vlen[c->tree.vid(c->tree.back(3))];
    else *score = mis;
    return 0;
}

int mpgpu_pattern_parsimony(mpgpu_ctx *c, uint16_t *ptn_pars, int32_t *sum)
{
    if (int rc = need_tree(c, false)) return rc;
    if (!ptn_pars) { set_error("null argument"); return 1; }
    MPGPU_CUDA(cudaSetDevice(c->device));
    if (c->sk.on) return sk_pattern_parsimony(c, ptn_pars, c->sort_alignment ? c->n_inf : c->P, sum);
    const int nbits = 16;
    if (int rc = compute_site_counters(c, nbits)) return rc;
    if (int rc = ensure_ptn_site(c)) return rc;
    const int upper = c->sort_alignment ? c->n_inf : c->P;                // :3380-3381
    if (int rc = ensure(c->d_ptn, c->ptn_cap, (size_t)upper + 2)) return rc;
    if (int rc = launch_gather_patterns(c, nbits, upper)) return rc;
    if (c->shard_count > 1 && c->reduces()) {
        // every pattern is non-zero on exactly one shard: summing u16 pairs as int32 cannot carry
        MPGPU_CUDA(cudaMemsetAsync(c->d_ptn + upper, 0, 2 * sizeof(uint16_t), c->stream));
        if (int rc = shard_sum(c, c->d_ptn, (upper + 1) / 2)) return rc;
    }
    if (upper > 0)
        MPGPU_CUDA(cudaMemcpyAsync(ptn_pars, c->d_ptn, (size_t)upper * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->stream));
    MPGPU_CUDA(cudaStreamSynchronize(c->stream));
    if (sum) {
        int s = 0;
        for (int i = 0; i < upper; i++) s += (int)ptn_pars[i] * c->weights[i];
        *sum = s;
    }
    return 0;
}

// ---- R11: -cost (Sankoff) ----------------------------------------------------------------------
int mpgpu_set_cost_matrix(mpgpu_ctx *c, const uint32_t *cost, int nstates, const int32_t *segment_upper, int nseg, uint32_t *highest)
{
    if (!c) { set_error("null context"); return 1; }
    if (!c->d_views) { set_error("no alignment loaded"); return 1; }
    MPGPU_CUDA(cudaSetDevice(c->device));
    Sankoff &k = c->sk;
    if (!cost) {                                            // back to Fitch
        if (c->reps.loaded && k.on) { set_error("replicates are loaded for -cost scoring: reload them after leaving -cost"); return 1; }
        k.on = false;
        c->lens_valid = false; c->kids_valid = false;
        if (c->tree_set) { if (int rc = compute_views(c)) return rc; if (c->reduces()) compute_lengths(c); }
        return 0;
    }
    if (nstates != c->S) { set_error("cost matrix size does not match the alignment's state count"); return 1; }
    if (!segment_upper || nseg < 1) { set_error("segment_upper is required (IQTree::doSegmenting, iqtree.cpp:3793)"); return 1; }
    if (c->reps.loaded) { set_error("set the cost matrix before mpgpu_load_replicates (the replicate tables are built per scoring mode)"); return 1; }
    uint32_t mx = 0;
    bool asym = false;
    for (int i = 0; i < nstates; i++)
        for (int j = 0; j < nstates; j++) {
            // an asymmetric matrix is legal in the reference (ParsTree::initCostMatrix only repairs the triangle inequality,
            // parstree.cpp:31-90): scores then depend on where the tree is rooted, and the kernels take the reference's rooted forms
            if (cost[i * nstates + j] != cost[j * nstates + i]) asym = true;
            mx = std::max(mx, cost[i * nstates + j]);
        }
    k.asym = asym;
    if (mx >= 65535) { set_error("cost matrix entry too large"); return 1; }
    k.cost.assign(cost, cost + nstates * nstates);
    k.highest = mx + 1;                                     // initializeCostMatrix :159-163
    k.seg_upper.assign(segment_upper, segment_upper + nseg);
    k.nseg = nseg;
    k.cost_dirty = true;
    k.on = true;
    if (int rc = sk_build(c)) { k.on = false; return rc; }
    c->lens_valid = false; c->kids_valid = false;
    if (c->tree_set) { if (int rc = compute_views(c)) return rc; compute_lengths(c); }
    if (highest) *highest = k.highest;
    return 0;
}

int mpgpu_sankoff_layout(mpgpu_ctx *c, int *vector_length, int *n_bounds, uint32_t *remainder_bounds, int capacity)
{
    if (!c || !c->sk.on) { set_error("no cost matrix set"); return 1; }
    if (vector_length) *vector_length = c->sk.Lref;
    const int nb = c->sk.nseg > 1 ? c->sk.nseg - 1 : 0;
    if (n_bounds) *n_bounds = nb;
    if (remainder_bounds) {
        if (capacity < nb) { set_error("capacity too small"); return 1; }
        for (int i = 0; i < nb; i++) remainder_bounds[i] = c->sk.lb[i];
    }
    return 0;
}

int mpgpu_sankoff_view(mpgpu_ctx *c, int node, int slot, uint16_t *out)
{
    if (int rc = need_tree(c, false)) return rc;
    if (!c->sk.on) { set_error("no cost matrix set"); return 1; }
    if (int rc = check_ref(c, node, slot)) return rc;
    if (!out) { set_error("null argument"); return 1; }
    MPGPU_CUDA(cudaSetDevice(c->device));
    return sk_raw_view(c, 3 * node + slot, out);
}

int mpgpu_scan_bounds(mpgpu_ctx *c, uint32_t *est_max, int capacity)
{
    if (!c || !c->sk.on) { set_error("no cost matrix set"); return 1; }
    if (!est_max || capacity < (int)c->sk.h_est.size()) { set_error("capacity too small"); return 1; }
    memcpy(est_max, c->sk.h_est.data(), c->sk.h_est.size() * sizeof(uint32_t));
    return 0;
}

int mpgpu_visit_order(mpgpu_ctx *c, int32_t *order)
{
    if (int rc = need_tree(c, false)) return rc;
    std::vector<int32_t> o;
    visit_order(c->tree, o);
    memcpy(order, o.data(), o.size() * sizeof(int32_t));
    return 0;
}

int mpgpu_scan_plan(mpgpu_ctx *c, const int32_t *order, int first, int count, int mintrav, int maxtrav,
                    int *n_cand, int *n_tasks)
{
    if (int rc = need_tree(c, true)) return rc;
    if (!order || first < 1 || count < 0 || first + count > 2 * c->n - 1) { set_error("bad visit range"); return 1; }
    MPGPU_CUDA(cudaSetDevice(c->device));
    const uint32_t vstride_vec = c->sk.on ? (uint32_t)(c->sk.vstride / 4) : (uint32_t)(c->view_stride / (c->S < 4 ? c->S : 4));
    if (int rc = build_scan_plan(c->tree, order, first, count, mintrav, maxtrav, vstride_vec, c->plan)) return rc;
    if (int rc = upload_plan(c)) return rc;
    if (n_cand) *n_cand = c->plan.n_cand;
    if (n_tasks) *n_tasks = (int)c->plan.tasks.size();
    return 0;
}

int64_t mpgpu_scan_plan_bytes(mpgpu_ctx *c)
{
    if (!c) return 0;
    return (int64_t)((size_t)c->plan.n_ops * (sizeof(ScanOffs) + sizeof(ScanCtl)) + c->plan.tasks.size() * sizeof(ScanTask));
}

int mpgpu_scan_launch(mpgpu_ctx *c, void **dev_counts)
{
    if (int rc = need_tree(c, true)) return rc;
    MPGPU_CUDA(cudaSetDevice(c->device));
    if (int rc = run_scan(c)) return rc;
    if (dev_counts) *dev_counts = (void *)c->d_counts;
    return 0;
}

int mpgpu_scan_finish(mpgpu_ctx *c, int32_t *visit_begin, uint32_t *mp, int32_t *cand_ref, int32_t *cand_prune, int capacity)
{
    if (int rc = need_tree(c, true)) return rc;
    if (!mp) { set_error("null argument"); return 1; }
    MPGPU_CUDA(cudaSetDevice(c->device));
    return finish_scan(c, visit_begin, mp, cand_ref, cand_prune, capacity);
}

int mpgpu_scan_visits(mpgpu_ctx *c, const int32_t *order, int first, int count, int mintrav, int maxtrav,
                      int32_t *visit_begin, uint32_t *mp, int32_t *cand_ref, int32_t *cand_prune,
                      int capacity, int *n_cand)
{
    if (c && !c->reduces()) { set_error("mpgpu_scan_visits on a sharded context needs mpgpu_set_allreduce (or use plan/launch/finish)"); return 1; }
    if (int rc = need_tree(c, true)) return rc;
    if (!order || !mp || first < 1 || count < 0 || first + count > 2 * c->n - 1) { set_error("bad visit range"); return 1; }
    g_sw.begin();
    MPGPU_CUDA(cudaSetDevice(c->device));
    g_sw.mark(0, "device set");
    if (int rc = scan_batch_pipelined(c, order, first, count, mintrav, maxtrav)) return rc;
    if (n_cand) *n_cand = c->plan.n_cand;
    if (c->plan.n_cand > capacity) { set_error("candidate capacity too small"); return 1; }
    return finish_scan(c, visit_begin, mp, cand_ref, cand_prune, capacity);
}

}  // extern "C"
