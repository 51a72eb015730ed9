// C-ABI layer, part 2: replicate scoring (R8) and pllOptimizeSprParsimony under -bb.
//
//   mpgpu_load_replicates      boot_samples_pars + segments -> resident tensor operand, exception lists
//   mpgpu_reps_current_tree    REPS of the current tree
//   mpgpu_reps_candidates      REPS of candidates of the last scan batch
//   mpgpu_optimize_spr(_bb)    the SPR hill-climb; with hooks/state = IQTree::saveCurrentTree's
//                              default policy replayed on the host in the reference's order
// No CPU fallback: every number a decision is based on comes from the device kernels.
#include "mpgpu_internal.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace mpgpu {

// wall-clock breakdown of the search loop, printed when MPGPU_PROFILE=1 (development aid)
struct Prof {
    bool on = getenv("MPGPU_PROFILE") != nullptr;
    double t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long n[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::chrono::steady_clock::time_point t0;
    void start() { if (on) t0 = std::chrono::steady_clock::now(); }
    void stop(int k) { if (on) { t[k] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); n[k]++; } }
    void report(const char *const *names, int cnt) {
        if (!on) return;
        for (int k = 0; k < cnt; k++) fprintf(stderr, "[mpgpu profile] %-12s %9.3f ms  (%ld)\n", names[k], t[k] * 1e3, n[k]);
    }
};

static Prof g_rp;      // REPS internals (sections are closed with a stream synchronize when profiling is on)
static const char *const g_rp_names[] = {"tree:counters", "tree:check", "tree:contract", "chunk:build", "chunk:rows", "chunk:contract", "chunk:combine", "chunk:readback"};
static inline void rp_stop(Ctx *c, int k) { if (g_rp.on) { cudaStreamSynchronize(c->stream); g_rp.stop(k); g_rp.start(); } }

void free_reps(Ctx *c)
{
    Reps &r = c->reps;
    void *ptrs[] = {r.d_w8, r.d_w16T, r.d_seg_upper, r.d_segmax, r.d_exc_ptn, r.d_exc_group, r.d_rows_site, r.d_rows_ptn, r.d_X, r.d_row_of,
                    r.d_row_tasks, r.d_edges, r.d_calls, r.d_res, r.d_thr, r.d_call_hit, r.d_hit_list, r.d_res_hit, r.d_full,
                    r.d_keep, r.d_keep_list, r.d_tree_ptn, r.d_remain, r.d_blist, r.d_segsum, r.d_pmax};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (r.ev0) { cudaEventDestroy(r.ev0); cudaEventDestroy(r.ev1); }
    if (r.h_pin) cudaFreeHost(r.h_pin);
    const bool use_tensor = r.use_tensor, timing = r.timing, nowrap = r.nowrap;
    r = Reps();
    r.use_tensor = use_tensor; r.timing = timing; r.nowrap = nowrap;
    c->d_row_of = nullptr; c->d_row_tasks = nullptr; c->d_rows_site = nullptr;
}

// Patterns [p_lo, p_hi) whose first expanded site falls into this shard's word slice, and the
// 128-pattern K-blocks [kb_lo, kb_hi) the tensor kernel has to visit for them.
static void update_kb_range(Ctx *c)
{
    Reps &r = c->reps;
    if (c->shard_count == 1) { r.p_lo = 0; r.p_hi = r.upper; r.kb_lo = 0; r.kb_hi = r.Kpad / 128; return; }
    const int64_t s_lo = c->w0 * 32, s_hi = (c->w0 + c->Wl) * 32;
    int p_lo = r.upper, p_hi = 0;
    if (c->ptn_identity) {
        p_lo = (int)std::min<int64_t>(s_lo, r.upper); p_hi = (int)std::min<int64_t>(s_hi, r.upper);
    } else {
        int64_t site = 0;
        for (int i = 0; i < r.upper; i++) {
            if (site >= s_lo && site < s_hi) { if (i < p_lo) p_lo = i; p_hi = i + 1; }
            site += c->weights[i];
        }
    }
    if (p_hi <= p_lo) { r.p_lo = r.p_hi = 0; r.kb_lo = r.kb_hi = 0; return; }
    r.p_lo = p_lo; r.p_hi = p_hi;
    r.kb_lo = p_lo / 128; r.kb_hi = (p_hi + 127) / 128;
}

static int ensure_rows(Ctx *c, int rows)
{
    Reps &r = c->reps;
    const bool ident = c->ptn_identity;                    // site rows double as pattern rows: no gather, no second buffer
    if (!ident && !r.d_rows_ptn) r.row_cap = 0;            // weights changed under us: the pattern-space buffer is needed now
    if (rows <= r.row_cap) return 0;
    const size_t per_row = (size_t)c->Wl * 4 + (ident ? 0 : (size_t)r.Pw * 4) + (size_t)r.G * r.Bpad * 4;
    size_t budget = (size_t)4 << 30;
    if (const char *e = getenv("MPGPU_REPS_ROW_BYTES")) { long long v = atoll(e); if (v > 0) budget = (size_t)v; }
    int limit = (int)std::min<size_t>(budget / per_row, (size_t)1 << 20);
    if (limit < kTreeRows + 64) limit = kTreeRows + 64;
    int want = std::max(rows + rows / 2, 1024);
    if (want > limit) want = limit;
    if (want <= r.row_cap) return 0;                       // already at the limit: callers chunk
    if (r.d_rows_site) cudaFree(r.d_rows_site);
    if (r.d_rows_ptn) cudaFree(r.d_rows_ptn);
    if (r.d_X) cudaFree(r.d_X);
    r.d_rows_site = nullptr; r.d_rows_ptn = nullptr; r.d_X = nullptr; r.row_cap = 0; r.tree_valid = false;
    MPGPU_CUDA(cudaMalloc((void **)&r.d_rows_site, (size_t)want * c->Wl * 4));
    if (!ident) MPGPU_CUDA(cudaMalloc((void **)&r.d_rows_ptn, (size_t)want * r.Pw * 4));
    MPGPU_CUDA(cudaMalloc((void **)&r.d_X, (size_t)want * r.G * r.Bpad * 4));
    r.row_cap = want;
    c->d_rows_site = r.d_rows_site;
    return 0;
}

// pattern-indexed bit rows a_base[0..nrows) -> X[x_row0 ..) (all groups)
// (bit i of a row = pattern 32 * a_word0 + i: site rows of a shard start at its first word)
static int contract_rows(Ctx *c, const uint32_t *a_base, int a_pitch, int a_word0, int x_row0, int nrows)
{
    Reps &r = c->reps;
    if (nrows == 0) return 0;
    MPGPU_CUDA(cudaMemsetAsync(r.d_X + (size_t)x_row0 * r.G * r.Bpad, 0, (size_t)nrows * r.G * r.Bpad * 4, c->stream));
    if (int rc = launch_reps_exc(c, a_base, a_pitch, a_word0, x_row0, nrows)) return rc;
    if (r.use_tensor) { if (int rc = launch_reps_tc(c, a_base, a_pitch, a_word0, x_row0, nrows)) return rc; }
    r.rows_scored += nrows;
    return 0;
}

// Exception lists + tensor operand for the current classification (seg_flagged, heavy):
// group 0 = everything wrap-free, group g >= 1 = the g-th flagged segment.  Exceptions = all
// patterns of flagged segments plus, in group 0, the patterns with a weight above 255 (or every
// pattern when the tensor path is switched off).
static int build_classification(Ctx *c)
{
    Reps &r = c->reps;
    const int nseg = (int)r.seg_upper.size();
    std::vector<int> group_of_seg(nseg, 0);
    int G = 1;
    for (int g = 0; g < nseg; g++) if (r.seg_flagged[g]) group_of_seg[g] = G++;
    std::vector<int32_t> exc_ptn, exc_group;
    std::vector<uint8_t> is_exc(r.Kpad, 0);
    r.n_heavy = 0;
    for (int pass = 0; pass < 2; pass++) {                         // group 0 first, then the flagged segments in order
        int s = 0;
        for (int p = 0; p < r.upper; p++) {
            while (p >= r.seg_upper[s]) s++;
            const int g = group_of_seg[s];
            if (pass == 0 ? (g == 0 && (r.heavy[p] || !r.use_tensor)) : g > 0) {
                exc_ptn.push_back(p); exc_group.push_back(g); is_exc[p] = 1;
                if (g == 0 && r.heavy[p]) r.n_heavy++;
            }
        }
    }
    if (G != r.G) {                                                 // X changes shape: drop the row buffers
        if (r.d_rows_site) cudaFree(r.d_rows_site);
        if (r.d_rows_ptn) cudaFree(r.d_rows_ptn);
        if (r.d_X) cudaFree(r.d_X);
        r.d_rows_site = nullptr; r.d_rows_ptn = nullptr; r.d_X = nullptr; r.row_cap = 0;
        c->d_rows_site = nullptr;
    }
    r.G = G;
    r.n_exc = (int)exc_ptn.size();
    r.tree_valid = false;
    if (int rc = ensure(r.d_exc_ptn, r.exc_cap, exc_ptn.size() + 1)) return rc;
    if (int rc = ensure(r.d_exc_group, r.exc_group_cap, exc_group.size() + 1)) return rc;
    uint8_t *d_is_exc = nullptr;
    MPGPU_CUDA(cudaMalloc((void **)&d_is_exc, (size_t)r.Kpad));
    if (r.n_exc) {
        MPGPU_CUDA(cudaMemcpyAsync(r.d_exc_ptn, exc_ptn.data(), (size_t)r.n_exc * 4, cudaMemcpyHostToDevice, c->stream));
        MPGPU_CUDA(cudaMemcpyAsync(r.d_exc_group, exc_group.data(), (size_t)r.n_exc * 4, cudaMemcpyHostToDevice, c->stream));
    }
    MPGPU_CUDA(cudaMemcpyAsync(d_is_exc, is_exc.data(), (size_t)r.Kpad, cudaMemcpyHostToDevice, c->stream));
    int rc = launch_build_w8(c, d_is_exc);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    cudaFree(d_is_exc);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "building the tensor operand");
    r.reclassifications++;
    return 0;
}

// rows 0..15 = bit planes of the per-site counters of the current tree, row 16 = sum_bit 2^bit * plane.
// Also the wrap check of this tree (k_seg_check): segments that might wrap move to a group of
// their own before anything is contracted.
static int refresh_tree_rows(Ctx *c)
{
    Reps &r = c->reps;
    if (r.tree_valid && r.row_cap > 0) return 0;
    const int nbits = 16;
    const int nseg = (int)r.seg_upper.size();
    g_rp.start();
    if (int rc = compute_site_counters(c, nbits)) return rc;
    if (int rc = ensure_ptn_site(c)) return rc;
    update_kb_range(c);
    rp_stop(c, 0);
    // per-pattern scores of the tree -> wrap check
    const int upper0 = c->sort_alignment ? c->n_inf : c->P;
    if (int rc = ensure(c->d_ptn, c->ptn_cap, (size_t)(upper0 > 0 ? upper0 : 1))) return rc;
    if (int rc = launch_gather_patterns(c, nbits, upper0)) return rc;
    if (r.keep_on && r.upper > 0) {                      // the exact skip test reads them after the batch (c->d_ptn is scratch for other calls)
        if (int rc = ensure(r.d_tree_ptn, r.tree_ptn_cap, (size_t)r.upper)) return rc;
        MPGPU_CUDA(cudaMemcpyAsync(r.d_tree_ptn, c->d_ptn, (size_t)r.upper * sizeof(uint16_t), cudaMemcpyDeviceToDevice, c->stream));
    }
    std::vector<int32_t> segmax(nseg);
    bool exact_check = true;
    if (c->shard_count == 1 && (int)r.seg_wmax.size() == nseg) {
        // cheap sufficient test first: (max score in the segment + 1) * (largest weight sum of the segment over the replicates)
        MPGPU_CUDA(cudaMemsetAsync(r.d_segmax, 0, (size_t)nseg * 4, c->stream));
        if (int rc = launch_seg_cmax(c, r.d_segmax)) return rc;
        MPGPU_CUDA(cudaMemcpyAsync(segmax.data(), r.d_segmax, (size_t)nseg * 4, cudaMemcpyDeviceToHost, c->stream));
        MPGPU_CUDA(cudaStreamSynchronize(c->stream));
        exact_check = false;
        for (int g = 0; g < nseg; g++)
            if (!r.seg_flagged[g] && (int64_t)(segmax[g] + 1) * r.seg_wmax[g] >= 65536) { exact_check = true; break; }
    }
    if (exact_check) {
        MPGPU_CUDA(cudaMemsetAsync(r.d_segmax, 0, (size_t)nseg * 4, c->stream));
        if (int rc = launch_seg_check(c, r.d_segmax)) return rc;
        if (int rc = shard_sum(c, r.d_segmax, nseg)) return rc;
        MPGPU_CUDA(cudaMemcpyAsync(segmax.data(), r.d_segmax, (size_t)nseg * 4, cudaMemcpyDeviceToHost, c->stream));
        MPGPU_CUDA(cudaStreamSynchronize(c->stream));
        bool grew = false;
        for (int g = 0; g < nseg; g++) if (segmax[g] >= 65536 && !r.seg_flagged[g]) { r.seg_flagged[g] = 1; grew = true; }
        if (grew) { if (int rc = build_classification(c)) return rc; }
    }
    rp_stop(c, 1);
    if (int rc = ensure_rows(c, kTreeRows + 64)) return rc;
    if (c->ptn_identity) {
        if (int rc = contract_rows(c, c->d_bitcnt, c->Wl, (int)c->w0, 0, nbits)) return rc;
    } else {
        if (int rc = launch_gather_rows(c, c->d_bitcnt, r.d_rows_ptn, nbits)) return rc;
        if (int rc = contract_rows(c, r.d_rows_ptn, r.Pw, 0, 0, nbits)) return rc;
    }
    if (int rc = launch_reps_tree_row(c, 0, nbits, kTreeRows - 1)) return rc;
    if (int rc = shard_sum(c, r.d_X + (size_t)(kTreeRows - 1) * r.G * r.Bpad, (int64_t)r.G * r.Bpad)) return rc;
    rp_stop(c, 2);
    r.tree_valid = true;
    return 0;
}

// Result of one REPS batch on the host: per call its row of B results, or nothing when no
// replicate of the call can pass its threshold
struct RepsOut {
    int Bpad = 0;
    std::vector<int32_t> dense;                   // concatenated rows [Bpad]
    std::vector<int64_t> dense_off;               // per call: offset into dense, -1 = no replicate can be affected
    std::vector<int32_t> orig;                    // per call: score on original_sample (column Buser), when loaded
    std::vector<int32_t> keep_slot;               // per call (keep_on): slot of its two bit rows in reps.d_keep, -1 = not kept
};

// keep_on: the bit rows (edge row, delta row; pattern space) of the calls of this chunk that were read back move to the side buffer
// before the next chunk overwrites the row buffers
static int keep_chunk_rows(Ctx *c, int ncalls, int done, RepsOut &out)
{
    Reps &r = c->reps;
    std::vector<int32_t> list;
    for (int i = 0; i < ncalls; i++) if (out.dense_off[done + i] >= 0) list.push_back(i);
    const int nl = (int)list.size();
    if (!nl) return 0;
    const size_t need = (size_t)(r.keep_used + nl) * 2 * r.Pw;
    if (need > r.keep_cap || !r.d_keep) {                  // grow, keeping the slots of the batch's earlier chunks
        uint32_t *nb = nullptr;
        const size_t want = need + need / 2 + 1024;
        MPGPU_CUDA(cudaMalloc((void **)&nb, want * sizeof(uint32_t)));
        if (r.d_keep && r.keep_used) MPGPU_CUDA(cudaMemcpyAsync(nb, r.d_keep, (size_t)r.keep_used * 2 * r.Pw * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
        MPGPU_CUDA(cudaStreamSynchronize(c->stream));
        if (r.d_keep) cudaFree(r.d_keep);
        r.d_keep = nb; r.keep_cap = want;
    }
    if (int rc = ensure(r.d_keep_list, r.keep_list_cap, (size_t)nl)) return rc;
    MPGPU_CUDA(cudaMemcpyAsync(r.d_keep_list, list.data(), (size_t)nl * 4, cudaMemcpyHostToDevice, c->stream));
    const uint32_t *src = c->ptn_identity ? r.d_rows_site : r.d_rows_ptn;
    const int pitch = c->ptn_identity ? c->Wl : r.Pw;
    if (int rc = launch_keep_rows(c, src, pitch, r.d_calls, r.d_keep_list, nl, r.d_keep + (size_t)r.keep_used * 2 * r.Pw)) return rc;
    MPGPU_CUDA(cudaStreamSynchronize(c->stream));         // `list` is pageable and goes out of scope
    for (int i = 0; i < nl; i++) out.keep_slot[done + list[i]] = r.keep_used + i;
    r.keep_used += nl;
    return 0;
}

// out[i] = max over the segments nseg/4 < s < nseg-1 of (prefix of 16-bit segment sums + remain bound) for the kept call in `slot`
// and replicate blist[i] (INT_MIN when no segment is in range): what the skip test of iqtree.cpp:3433-3445 compares with boot_logl
static int prefix_max(Ctx *c, int slot, const int32_t *blist, int nl, int32_t *out)
{
    Reps &r = c->reps;
    if (nl == 0) return 0;
    if (slot < 0 || slot >= r.keep_used) { set_error("internal: the rows of this call were not kept"); return 2; }
    const int nseg = (int)r.seg_upper.size();
    if (int rc = ensure(r.d_blist, r.blist_cap, (size_t)nl)) return rc;
    if (int rc = ensure(r.d_pmax, r.pmax_cap, (size_t)nl)) return rc;
    if (int rc = ensure(r.d_segsum, r.segsum_cap, (size_t)nl * nseg)) return rc;
    MPGPU_CUDA(cudaMemcpyAsync(r.d_blist, blist, (size_t)nl * 4, cudaMemcpyHostToDevice, c->stream));
    if (int rc = launch_prefix_max(c, r.d_tree_ptn, r.d_keep + (size_t)slot * 2 * r.Pw, r.d_remain, r.d_blist, nl, r.d_segsum, r.d_pmax)) return rc;
    MPGPU_CUDA(cudaMemcpyAsync(out, r.d_pmax, (size_t)nl * 4, cudaMemcpyDeviceToHost, c->stream));
    MPGPU_CUDA(cudaStreamSynchronize(c->stream));
    r.prefix_calls++; r.prefix_pairs += nl;
    return 0;
}

// Read-back of one chunk whose results sit in reps.d_res[ncalls][Bpad] (and hit flags in d_call_hit when thr):
// the original-frequency column of every call, and only the rows of calls that can change a replicate.
// replicate shards: rows [nrows][Bpad] of this rank's replicates (src, or the rows src[list[i]]) -> full-width rows
// [nrows][Btot_pad] in d_full, every rank's columns in place, summed over the group (the other ranks' columns are zero here)
static int assemble_full_rows(Ctx *c, const int32_t *src, int nrows, int Btot_pad)
{
    Reps &r = c->reps;
    if (int rc = ensure(r.d_full, r.full_cap, (size_t)nrows * Btot_pad + 4)) return rc;
    MPGPU_CUDA(cudaMemsetAsync(r.d_full, 0, (size_t)nrows * Btot_pad * 4, c->stream));
    MPGPU_CUDA(cudaMemcpy2DAsync(r.d_full + r.rep_lo, (size_t)Btot_pad * 4, src, (size_t)r.Bpad * 4, (size_t)r.Buser * 4, (size_t)nrows,
                                 cudaMemcpyDeviceToDevice, c->stream));
    return group_sum(c, r.d_full, (int64_t)nrows * Btot_pad);
}

static int reps_read_back(Ctx *c, int ncalls, int done, const int32_t *thr, RepsOut &out, std::vector<int32_t> &hit_list)
{
    Reps &r = c->reps;
    const bool reps_sharded = c->rep_count > 1;
    const int W = out.Bpad;                                // row width on the host: Bpad, or the padded total over the replicate shards
    // ---- read back: the original-frequency score of every call (ratchet iterations) ----
    if (r.has_orig) {
        MPGPU_CUDA(cudaMemcpy2DAsync(out.orig.data() + done, 4, r.d_res + r.Buser, (size_t)r.Bpad * 4, 4, (size_t)ncalls,
                                     cudaMemcpyDeviceToHost, c->stream));
        MPGPU_CUDA(cudaStreamSynchronize(c->stream));
    }
    // ---- read back: only the rows of calls that can change a replicate ----
    bool all_rows = true;
    if (thr) {
        int32_t *fl = (int32_t *)r.pinned((size_t)ncalls * 4);
        if (!fl) { set_error("pinned host allocation failed"); return 2; }
        if (reps_sharded) { if (int rc = group_sum(c, r.d_call_hit, ncalls)) return rc; }     // a call is read back when ANY shard's replicates can be hit
        MPGPU_CUDA(cudaMemcpyAsync(fl, r.d_call_hit, (size_t)ncalls * 4, cudaMemcpyDeviceToHost, c->stream));
        MPGPU_CUDA(cudaStreamSynchronize(c->stream));
        hit_list.clear();
        for (int i = 0; i < ncalls; i++) if (fl[i]) hit_list.push_back(i);
        if ((int)hit_list.size() * 2 < ncalls) {
            all_rows = false;
            const int nl = (int)hit_list.size();
            if (nl) {
                if (int rc = ensure(r.d_hit_list, r.hit_list_cap, (size_t)nl)) return rc;
                if (int rc = ensure(r.d_res_hit, r.res_hit_cap, (size_t)nl * r.Bpad)) return rc;
                MPGPU_CUDA(cudaMemcpyAsync(r.d_hit_list, hit_list.data(), (size_t)nl * 4, cudaMemcpyHostToDevice, c->stream));
                if (int rc = launch_gather_res_rows(c, r.d_res, r.d_hit_list, nl, r.d_res_hit)) return rc;
                const int32_t *rows = r.d_res_hit;
                if (reps_sharded) { if (int rc = assemble_full_rows(c, r.d_res_hit, nl, W)) return rc; rows = r.d_full; }
                const size_t off = out.dense.size(), bytes = (size_t)nl * W * 4;
                void *hp = r.pinned(bytes);
                if (!hp) { set_error("pinned host allocation failed"); return 2; }
                MPGPU_CUDA(cudaMemcpyAsync(hp, rows, bytes, cudaMemcpyDeviceToHost, c->stream));
                MPGPU_CUDA(cudaStreamSynchronize(c->stream));
                out.dense.resize(off + (size_t)nl * W);
                memcpy(out.dense.data() + off, hp, bytes);
                for (int i = 0; i < nl; i++) out.dense_off[done + hit_list[i]] = (int64_t)(off + (size_t)i * W);
            }
        }
    }
    if (all_rows) {
        const int32_t *rows = r.d_res;
        if (reps_sharded) { if (int rc = assemble_full_rows(c, r.d_res, ncalls, W)) return rc; rows = r.d_full; }
        const size_t off = out.dense.size(), bytes = (size_t)ncalls * W * 4;
        void *hp = r.pinned(bytes);
        if (!hp) { set_error("pinned host allocation failed"); return 2; }
        MPGPU_CUDA(cudaMemcpyAsync(hp, rows, bytes, cudaMemcpyDeviceToHost, c->stream));
        MPGPU_CUDA(cudaStreamSynchronize(c->stream));
        out.dense.resize(off + (size_t)ncalls * W);
        memcpy(out.dense.data() + off, hp, bytes);
        for (int i = 0; i < ncalls; i++) out.dense_off[done + i] = (int64_t)(off + (size_t)i * W);
    }
    return 0;
}

// -cost: the calls' vectors are per-pattern cost rows (sankoff.cu), one exact contraction per chunk
static int sk_reps_run(Ctx *c, const int32_t *cands, int m, const int32_t *thr, RepsOut &out, bool device_only)
{
    Reps &r = c->reps;
    const ScanPlan &pl = c->plan;
    if (thr) {
        if (!r.d_thr) MPGPU_CUDA(cudaMalloc((void **)&r.d_thr, (size_t)r.Bpad * 4));
        MPGPU_CUDA(cudaMemcpyAsync(r.d_thr, thr + r.rep_lo, (size_t)r.Buser * 4, cudaMemcpyHostToDevice, c->stream));   // (replicate shards: this rank's thresholds)
    }
    std::vector<int32_t> &row_of = r.h_row_of, &call_row = r.h_row_tasks;
    MPGPU_CUDA(cudaStreamSynchronize(c->stream));
    const int cap = sk_reps_rows_capacity(c) - 1;
    std::vector<int32_t> hit_list, visits, vrow;
    const bool asym = c->sk.asym;
    int done = 0;
    while (done < m) {
        row_of.assign(pl.n_cand > 0 ? pl.n_cand : 1, -1);
        call_row.clear(); visits.clear();
        vrow.assign(pl.visit_ref.size() + 1, -1);
        int nsel = 0, k = done;
        // call_row: 0 = the current tree at tr->start; -(1 + i) = the i-th visit row of the chunk; 1 + j = the j-th selected candidate
        // (visit rows sit in front of the candidates' rows: the final row numbers are known once the chunk is carved)
        for (; k < m; k++) {
            const int j = cands[k];
            if (j < 0) {
                // -(2 + v): the current tree as the v-th planned visit evaluates it (rooted at that visit's edge, :2286); the same
                // vector as -1 unless the cost matrix is asymmetric
                const int v = -j - 2;
                if (!asym || j == -1) { call_row.push_back(0); continue; }
                if (v >= (int)pl.visit_ref.size()) { set_error("visit index out of range"); return 1; }
                if (vrow[v] < 0) { if (nsel + (int)visits.size() + 1 > cap) break; vrow[v] = (int)visits.size(); visits.push_back(v); }
                call_row.push_back(-(1 + vrow[v]));
                continue;
            }
            if (j >= pl.n_cand) { set_error("candidate index out of range"); return 1; }
            if (row_of[j] < 0) { if (nsel + (int)visits.size() + 1 > cap) break; row_of[j] = nsel++; }
            call_row.push_back(1 + row_of[j]);
        }
        if (k == done) { set_error("REPS row buffers too small for a single candidate"); return 1; }
        if (device_only && k < m) { set_error("REPS batch does not fit the row buffers in one piece (raise MPGPU_REPS_ROW_BYTES)"); return 1; }
        const int ncalls = k - done, nvis = (int)visits.size();
        for (int &cr : call_row) cr = cr < 0 ? -cr : (cr > 0 ? cr + nvis : 0);      // visit row i -> 1 + i, candidate j -> 1 + nvis + j
        if (int rc = sk_reps_chunk(c, row_of.data(), nsel, call_row.data(), ncalls, thr != nullptr, visits.data(), nvis)) return rc;
        if (device_only) return 0;
        if (int rc = reps_read_back(c, ncalls, done, thr, out, hit_list)) return rc;
        done = k;
    }
    return 0;
}

// REPS vectors for the calls cands[0..m) of the last planned scan batch (-1 = the current tree),
// in order.  thr (host, [B], nullable): only entries with res <= thr[b] are needed by the caller.
// device_only: enqueue the work of a single chunk and return without reading anything back (d_res holds
// [m][Bpad] when the stream reaches this point).
static int reps_run(Ctx *c, const int32_t *cands, int m, const int32_t *thr, RepsOut &out, bool device_only = false)
{
    Reps &r = c->reps;
    const ScanPlan &pl = c->plan;
    const HostTree &t = c->tree;
    out.Bpad = c->rep_count > 1 ? (r.B_total + 255) / 256 * 256 : r.Bpad;
    out.dense.clear();
    out.dense_off.assign(m, -1);
    out.orig.assign(r.has_orig ? m : 0, 0);
    out.keep_slot.assign(r.keep_on ? m : 0, -1);
    r.keep_used = 0;
    if (m == 0) return 0;
    if (c->sk.on && r.nowrap) { set_error("option reps_nowrap (-autovec) is not available under -cost"); return 1; }
    if (c->sk.on) return sk_reps_run(c, cands, m, thr, out, device_only);
    if (int rc = refresh_tree_rows(c)) return rc;
    // rows the whole list would like to have; ensure_rows clamps to the memory budget
    {
        int want = kTreeRows;
        for (int i = 0; i < m; i++) if (cands[i] >= 0) want += 2;
        if (int rc = ensure_rows(c, want)) return rc;
        if (int rc = refresh_tree_rows(c)) return rc;        // a reallocation drops the tree rows
    }
    const int max_rows = r.row_cap - kTreeRows;
    if (thr) {
        if (!r.d_thr) MPGPU_CUDA(cudaMalloc((void **)&r.d_thr, (size_t)r.Bpad * 4));
        MPGPU_CUDA(cudaMemcpyAsync(r.d_thr, thr + r.rep_lo, (size_t)r.Buser * 4, cudaMemcpyHostToDevice, c->stream));   // (replicate shards: this rank's thresholds)
    }
    // staging vectors live in the context: asynchronous uploads may still read them after we return
    std::vector<int32_t> &row_of = r.h_row_of, &row_tasks = r.h_row_tasks;
    std::vector<int4> &edges = r.h_edges;
    std::vector<int2> &calls = r.h_calls;
    MPGPU_CUDA(cudaStreamSynchronize(c->stream));             // previous batch's uploads are done with them
    row_of.assign(pl.n_cand > 0 ? pl.n_cand : 1, -1);
    std::vector<int32_t> task_row(pl.tasks.size());
    std::vector<int32_t> hit_list;
    int done = 0;
    while (done < m) {
        g_rp.start();
        // ---- carve a chunk that fits the row buffers ----
        std::fill(row_of.begin(), row_of.end(), -1);
        std::fill(task_row.begin(), task_row.end(), -1);
        row_tasks.clear(); edges.clear(); calls.clear();
        int nrows = 0, k = done;
        for (; k < m; k++) {
            const int j = cands[k];
            if (j < 0) { calls.push_back(make_int2(-1, -1)); continue; }
            if (j >= pl.n_cand) { set_error("candidate index out of range"); return 1; }
            const int ti = pl.cand_task[j];
            const int need = (task_row[ti] < 0 ? 1 : 0) + (row_of[j] < 0 ? 1 : 0);
            if (nrows + need > max_rows) break;
            if (task_row[ti] < 0) {
                const int pr = pl.cand_prune[j];
                task_row[ti] = kTreeRows + nrows++;
                edges.push_back(make_int4(t.vid(pr), t.vid(t.back(pr)), task_row[ti], 0));
                row_tasks.push_back(ti);
            }
            if (row_of[j] < 0) row_of[j] = nrows++;          // relative to the first batch row
            calls.push_back(make_int2(task_row[ti], kTreeRows + row_of[j]));
        }
        if (k == done) { set_error("REPS row buffers too small for a single candidate"); return 1; }
        if (device_only && k < m) { set_error("REPS batch does not fit the row buffers in one piece (raise MPGPU_REPS_ROW_BYTES)"); return 1; }
        const int ncalls = k - done;
        rp_stop(c, 3);
        // ---- upload the chunk ----
        if (int rc = ensure(r.d_row_of, r.row_of_cap, row_of.size())) return rc;
        if (int rc = ensure(r.d_row_tasks, r.row_tasks_cap, row_tasks.size() + 1)) return rc;
        if (int rc = ensure(r.d_edges, r.edges_cap, edges.size() + 1)) return rc;
        if (int rc = ensure(r.d_calls, r.calls_cap, calls.size())) return rc;
        if (int rc = ensure(r.d_res, r.res_cap, (size_t)ncalls * r.Bpad)) return rc;
        c->d_row_of = r.d_row_of; c->d_row_tasks = r.d_row_tasks; c->d_rows_site = r.d_rows_site + (size_t)kTreeRows * c->Wl;
        MPGPU_CUDA(cudaMemcpyAsync(r.d_calls, calls.data(), calls.size() * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
        if (nrows > 0) {
            MPGPU_CUDA(cudaMemcpyAsync(r.d_row_of, row_of.data(), row_of.size() * 4, cudaMemcpyHostToDevice, c->stream));
            MPGPU_CUDA(cudaMemcpyAsync(r.d_row_tasks, row_tasks.data(), row_tasks.size() * 4, cudaMemcpyHostToDevice, c->stream));
            MPGPU_CUDA(cudaMemcpyAsync(r.d_edges, edges.data(), edges.size() * sizeof(int4), cudaMemcpyHostToDevice, c->stream));
            // ---- rows: edge rows, delta rows (second pass of the scan), site -> pattern space ----
            if (int rc = launch_edge_rows(c, r.d_edges, (int)edges.size(), r.d_rows_site)) return rc;
            if (int rc = launch_scan_rows(c, (int)row_tasks.size(), pl.max_slot)) return rc;
            rp_stop(c, 4);
            if (c->ptn_identity) {
                if (int rc = contract_rows(c, r.d_rows_site + (size_t)kTreeRows * c->Wl, c->Wl, (int)c->w0, kTreeRows, nrows)) return rc;
            } else {
                if (int rc = launch_gather_rows(c, r.d_rows_site + (size_t)kTreeRows * c->Wl,
                                                r.d_rows_ptn + (size_t)kTreeRows * r.Pw, nrows)) return rc;
                if (int rc = contract_rows(c, r.d_rows_ptn + (size_t)kTreeRows * r.Pw, r.Pw, 0, kTreeRows, nrows)) return rc;
            }
        }
        if (nrows > 0) { if (int rc = shard_sum(c, r.d_X + (size_t)kTreeRows * r.G * r.Bpad, (int64_t)nrows * r.G * r.Bpad)) return rc; }
        rp_stop(c, 5);
        // ---- combine ----
        if (thr) {
            if (int rc = ensure(r.d_call_hit, r.call_hit_cap, (size_t)ncalls)) return rc;
            MPGPU_CUDA(cudaMemsetAsync(r.d_call_hit, 0, (size_t)ncalls * 4, c->stream));
        }
        if (int rc = launch_reps_combine(c, kTreeRows - 1, r.d_calls, ncalls, r.d_res, thr ? r.d_thr : nullptr, r.d_call_hit)) return rc;
        rp_stop(c, 6);
        if (device_only) return 0;
        if (int rc = reps_read_back(c, ncalls, done, thr, out, hit_list)) return rc;
        if (r.keep_on) { if (int rc = keep_chunk_rows(c, ncalls, done, out)) return rc; }
        rp_stop(c, 7);
        done = k;
    }
    return 0;
}

// ---- the search --------------------------------------------------------------------------------
struct BBRun {
    const mpgpu_bb_hooks *hooks;
    mpgpu_bb_state *st;
    // host screen: replicate b can only be touched by a call with res <= screen[b] = floor(-boot_logl[b] + eps)
    // (rell > boot_logl - eps, rell >= boot_logl and rell == boot_logl all imply it); kept current as boot_logl moves
    std::vector<int32_t> screen;
    // MPGPU_BB_DISTINCT_ITER: replicates of the call at hand that reach their threshold, and their skip maxima
    std::vector<int32_t> dist_list, dist_max;
    int rc = 0;                      // first error of a bb_save (the replay lambdas return nothing)
};

static inline int32_t bb_screen_of(double boot_logl, double eps)
{
    const double lim = std::floor(-boot_logl + eps);
    return lim >= 2147483647.0 ? 2147483647 : (lim <= -2147483648.0 ? (int32_t)-2147483647 - 1 : (int32_t)lim);
}

static inline bool bb_passes_logl(const mpgpu_bb_state *st, double cur_logl)
{
    return !(st->logl_cutoff != 0.0 && cur_logl <= st->logl_cutoff - 1e-4);          // iqtree.cpp:3343
}
static inline bool bb_passes(const mpgpu_bb_state *st, uint32_t mp) { return bb_passes_logl(st, -(double)mp); }

// Sum over segments of the 16-bit-wrapped lane sums of a * b (iqtree.cpp:3285-3292)
static int32_t segmented_u16_dot(const uint16_t *a, const uint16_t *b, const std::vector<int32_t> &seg_upper, int upper)
{
    int32_t total = 0;
    int p = 0;
    for (size_t s = 0; s < seg_upper.size(); s++) {
        uint32_t acc = 0;
        const int hi = std::min(upper, (int)seg_upper[s]);
        for (; p < hi; p++) acc += (uint32_t)a[p] * b[p];
        total += (int32_t)(acc & 0xFFFFu);
    }
    return total;
}

// IQTree::saveCurrentTree for one call that passed the cutoff (iqtree.cpp:3345-3348, 3687-3731)
static void bb_save(BBRun *bb, Ctx *c, double cur_logl, const RepsOut &ro, int call, int remove_ref, int insert_ref)
{
    mpgpu_bb_state *st = bb->st;
    const mpgpu_bb_hooks *hk = bb->hooks;
    const double eps = st->ufboot_epsilon;
    int32_t tree_index = hk->push_tree_logl(hk->user, cur_logl);
    const int32_t tree_index_pushed = tree_index;         // treels_logl.size() - 1 of this call
    if (st->updates_off) { st->n_reps++; return; }        // -min_iter1_cand, iteration 1 (iqtree.cpp:3404)
    int32_t *const orig_logl = st->boot_tree_orig_logl;   // -cutoff_from_btrees
    bool have = false;
    const bool mulhits = st->policy == MPGPU_BB_MULHITS;
    auto one = [&](int b, int32_t res) {
        const double rell = -(double)res;
        const double bl = st->boot_logl[b];
        if (st->policy == MPGPU_BB_MULHITS_TOP) {        // iqtree.cpp:3536-3583
            const int32_t N = st->top_n;
            if (st->top_count[b] < N || rell > (double)st->boot_threshold[b]) {
                const int32_t pushed = tree_index_pushed;
                if (!have) {
                    have = true;
                    tree_index = hk->materialize(hk->user, c->tree.bn.data(), c->tree.bs.data(), remove_ref, insert_ref, tree_index);
                }
                if (tree_index == pushed) {              // a newly added tree (:3556)
                    const int32_t r = (int32_t)rell;
                    if (st->top_count[b] < N) {
                        hk->tophit(hk->user, b, tree_index, r, 0);
                        st->top_count[b]++;
                        if (!(st->boot_threshold[b] < r)) st->boot_threshold[b] = r;          // :3569
                    } else {
                        st->boot_threshold[b] = hk->tophit(hk->user, b, tree_index, r, 1);    // :3571-3578
                    }
                }
            }
            return;
        }
        if (mulhits) {                                   // iqtree.cpp:3498-3531
            if (rell >= bl) {
                if (!have) {
                    have = true;
                    tree_index = hk->materialize(hk->user, c->tree.bn.data(), c->tree.bs.data(), remove_ref, insert_ref, tree_index);
                }
                if (rell > bl) st->boot_logl[b] = rell;
                if (orig_logl && cur_logl > (double)orig_logl[b]) orig_logl[b] = (int32_t)cur_logl;      // :3524-3527
                hk->mulhit(hk->user, b, tree_index, rell > bl);
            }
            return;
        }
        if (rell > bl + eps || (rell > bl - eps && hk->random_double(hk->user) <= 1.0 / (st->boot_counts[b] + 1))) {
            if (!have) {
                have = true;
                tree_index = hk->materialize(hk->user, c->tree.bn.data(), c->tree.bs.data(), remove_ref, insert_ref, tree_index);
            }
            if (rell > bl) st->boot_counts[b] = 1;
            if (orig_logl) orig_logl[b] = (int32_t)cur_logl;                                              // :3717-3718
            st->boot_logl[b] = std::max(bl, rell);
            st->boot_trees[b] = tree_index;
        }
        if (rell == st->boot_logl[b]) st->boot_counts[b]++;
    };
    if (ro.dense_off[call] >= 0 && st->policy == MPGPU_BB_DISTINCT_ITER) {          // iqtree.cpp:3587-3685
        const int32_t *row = ro.dense.data() + ro.dense_off[call];
        Reps &r = c->reps;
        std::vector<int32_t> &L = bb->dist_list, &M = bb->dist_max;
        L.clear();
        for (int b = 0; b < st->B; b++) if (-(double)row[b] >= (double)st->boot_threshold[b]) L.push_back(b);     // :3588 can fire
        // the remain-bound skip (:3433-3445, `continue` at :3484) compares with boot_logl, acceptance with boot_threshold <= boot_logl:
        // decided exactly from the segment prefixes of the call's own pattern vector
        const bool skip_test = r.remain_loaded && r.seg_upper.size() > 1;
        if (skip_test && !L.empty()) {
            M.resize(L.size());
            if (int rc = prefix_max(c, ro.keep_slot[call], L.data(), (int)L.size(), M.data())) { if (!bb->rc) bb->rc = rc; st->n_reps++; return; }
        }
        for (size_t i = 0; i < L.size(); i++) {
            const int b = L[i];
            if (skip_test && M[i] != (int32_t)0x80000000 && (double)(-(int64_t)M[i]) < st->boot_logl[b] - eps) continue;
            const double rell = -(double)row[b];
            const int32_t thr = st->boot_threshold[b];
            st->boot_counts[b]++;                                                              // :3588-3590
            if (rell > (double)thr || hk->random_double(hk->user) <= st->top_n * 1.0 / st->boot_counts[b]) {      // :3592-3594 (else: rell == thr)
                if (rell > st->boot_logl[b]) st->boot_counts[b] = 1;                           // :3597
                if (!have) {
                    have = true;
                    tree_index = hk->materialize(hk->user, c->tree.bn.data(), c->tree.bs.data(), remove_ref, insert_ref, tree_index);
                }
                if (orig_logl) orig_logl[b] = (int32_t)cur_logl;                               // :3618-3619
                st->boot_trees[b] = tree_index;                                                // :3621
                st->boot_logl[b] = std::max(st->boot_logl[b], rell);
                st->boot_threshold[b] = hk->disthit(hk->user, b, tree_index, (int32_t)rell, st->cur_it, st->top_n, thr);   // :3624-3678
            }
        }
    } else
    if (ro.dense_off[call] >= 0) {                      // else: no replicate of this call reaches its threshold
        const int32_t *row = ro.dense.data() + ro.dense_off[call];
        if (st->policy == MPGPU_BB_MULHITS_TOP) {
            for (int b = 0; b < st->B; b++) one(b, row[b]);
        } else {
            const int32_t *scr = bb->screen.data();
            for (int b = 0; b < st->B; b++) {
                if (row[b] > scr[b]) continue;          // integer screen: nothing of :3498-3531 / :3687-3731 can fire
                const double before = st->boot_logl[b];
                one(b, row[b]);
                if (st->boot_logl[b] != before) bb->screen[b] = bb_screen_of(st->boot_logl[b], eps);
            }
        }
    }
    st->n_reps++;
}

}  // namespace mpgpu

using namespace mpgpu;

// pllOptimizeSprParsimony (sprparsimony.cpp:3244-3319) with the node loop's scoring batched on
// the device.  Speculation: the candidates of the next K visits are scored against the current
// tree; the host replays testInsertParsimony's bookkeeping (:2168-2176), the saveCurrentTree
// up-calls (:2163-2166, :2286-2289) and the node loop's acceptance test (:3306-3314) strictly in
// order, and throws the rest of a batch away as soon as a move is applied (the only event that
// changes any score).
static int optimize_impl(mpgpu_ctx *c, int32_t *back_node, int32_t *back_slot, int mintrav, int maxtrav,
                         mpgpu_rng_fn rng, void *rng_user, BBRun *bb, uint32_t *best, int64_t *n_insertions,
                         bool stepwise_on = false, const uint32_t *start_score = nullptr)
{
    if (!c->reduces()) { set_error("the SPR search on a sharded context needs mpgpu_set_allreduce"); return 1; }
    if (mintrav != 1) { set_error("mintrav must be 1 (assert at sprparsimony.cpp:2278)"); return 1; }
    if (bb && c->sk.on != c->reps.loaded_sankoff) { set_error("the replicates were loaded for the other scoring mode: call mpgpu_load_replicates after mpgpu_set_cost_matrix"); return 1; }
    // MPGPU_PROFILE=1: host-side section times per search; MPGPU_PROFILE=2: accumulated over the process, printed at exit
    static const bool cumulative = getenv("MPGPU_PROFILE") && atoi(getenv("MPGPU_PROFILE")) >= 2;
    static Prof g_sp;
    Prof local_prof;
    Prof &prof = cumulative ? g_sp : local_prof;
    static const char *const prof_names[] = {"plan+launch", "scan wait", "reps", "replay", "views", "set_tree"};
    prof.start();
    if (int rc = set_tree_impl(c, back_node, back_slot, true)) return rc;
    prof.stop(5);
    // -cost, plain mode: evaluateSankoff... leaves early when a prefix of segment sums plus the remainder bound
    // exceeds tr->bestParsimony (:951-956); its return value is then > best, i.e. the insertion changes nothing
    // (-bb runs with perSiteScores: no early exit; neither do the SPR rounds of the stepwise-addition tree, where the
    // reference keeps doing_stepwise_addition set for the whole of _pllMakeParsimonyTreeFast, :3226-3233 / :951)
    const bool sk_early = c->sk.on && !c->sk.exact && c->sk.nseg > 1 && !bb && !stepwise_on;
    const int n = c->n, nvisit = 2 * n - 2;
    uint32_t score = 0;
    // the SPR rounds of a stepwise-addition tree start from the last insertion's score (randomMP = tr->bestParsimony, :3171: no
    // evaluation at tr->start) -- the same number unless the cost matrix is asymmetric, where it is rooted at the last tip
    if (start_score) score = *start_score;
    else if (c->start_edge_valid) score = c->start_edge_mis + c->vlen[c->tree.vid(c->tree.back(3))];   // :3277, read back with the view counts
    else if (int rc = mpgpu_tree_score(c, &score)) return rc;
    uint32_t bestParsimony = score;
    c->search_start_score = score; c->search_moves = 0; c->search_batches = 0;
    uint32_t randomMP = bestParsimony, startMP = 0;
    uint32_t cur_score = score;                                     // score of the tree in c->tree
    unsigned int bestIterationScoreHits = 1;
    int64_t scored = 0;
    std::vector<int32_t> order, vbegin, cref, cprune, pass_cands, call_of, vis_idx;
    std::vector<uint32_t> vis_score;
    std::vector<uint32_t> mp;
    std::vector<int32_t> thr;
    RepsOut ro;
    int32_t ratchet_stale = 0;
    if (bb && bb->st->ratchet) {
        if (!c->reps.has_orig) { set_error("ratchet iteration: load the replicates with mpgpu_load_replicates2 (original_sample)"); return 1; }
        if (!bb->st->ratchet_pattern_pars) { set_error("ratchet iteration: ratchet_pattern_pars is null"); return 1; }
        ratchet_stale = segmented_u16_dot(bb->st->ratchet_pattern_pars, c->reps.original_sample.data(), c->reps.seg_upper, c->reps.upper);
    }
    // Speculation depth: visits planned and scored per batch.  Everything after the first accepted move of a batch is
    // thrown away, so right after a move the batch follows the recent distance between moves (x2); while no move happens
    // it grows fourfold.  The distance is remembered across searches (refinement and ratchet searches start near an
    // optimum: whole passes in one batch).  The ceiling keeps the thrown-away kernel time of one batch near 25 us
    // (~9 G scored chunk-insertions/s, ~36 insertions per visit).  Only wasted work depends on any of this, never a result.
    double &move_gap = c->move_gap;
    const int spec_cap = [&]() {
        const double units_per_visit = 36.0 * (double)(c->Wl / kChunkWords) * (c->S <= 4 ? 1.0 : c->S / 4.0);
        double v = 25e-6 * 9e9 / units_per_visit;
        if (bb) {                               // under -bb the batch's candidates also go through the replicate contraction
            const double vb = 25e-6 * 2.5e15 / (2.0 * 36.0 * (double)std::max(1, c->reps.upper) * (double)std::max(1, c->reps.B));
            if (vb < v) v = vb;
        }
        return v > (double)nvisit ? nvisit : (v < 16.0 ? 16 : (int)v);
    }();
    int since_move = 0;
    static const bool eager_views = getenv("MPGPU_EAGER_VIEWS") != nullptr;
    const bool lazy = !eager_views && c->kids_valid && c->lens_valid;
    auto after_move_batch = [spec_cap](double gap) {
        static const int fixed = getenv("MPGPU_SEARCH_BATCH") ? atoi(getenv("MPGPU_SEARCH_BATCH")) : 0;   // tuning knob
        if (fixed > 0) return fixed;
        const int b = (int)(2.0 * gap) + 1;
        return b < 2 ? 2 : (b > spec_cap ? spec_cap : b);
    };
    do {
        startMP = randomMP;
        visit_order(c->tree, order);                              // nodeRectifierPars :3297
        int i = 1;
        int batch = after_move_batch(move_gap);
        while (i <= nvisit) {
            int count = std::min(batch, nvisit - i + 1);
            prof.start();
            if (int rc = scan_batch_pipelined(c, order.data(), i, count, mintrav, maxtrav)) return rc;
            c->search_batches++;
            prof.stop(0); prof.start();
            const int nc = c->plan.n_cand;
            vbegin.resize(count + 1); mp.resize(nc + 1); cref.resize(nc + 1); cprune.resize(nc + 1);
            if (int rc = finish_scan(c, vbegin.data(), mp.data(), cref.data(), cprune.data(), nc + 1)) return rc;
            prof.stop(1); prof.start();
            // -cost with an asymmetric matrix: rearrangeParsimony evaluates the current tree at the visited edge before it saves it
            // (:2286-2289), so the score (and the vector, see sk_reps_run) of a visit's current-tree call is rooted there
            const bool visit_rooted = bb && c->sk.on && c->sk.asym;
            if (visit_rooted) {
                vis_idx.resize(count); vis_score.resize(count);
                for (int v = 0; v < count; v++) vis_idx[v] = v;
                if (int rc = sk_visit_edges(c, vis_idx.data(), count, nullptr)) return rc;
                for (int v = 0; v < count; v++) vis_score[v] = c->sk.h_tot[v].x;
            }
            if (bb) {
                // every saveCurrentTree call of the batch, in order; call_of[] = index into the REPS results
                mpgpu_bb_state *st = bb->st;
                pass_cands.clear();
                call_of.assign((size_t)count + nc, -1);
                // ratchet iteration: whether a call passes depends on the previous passing call's vector, so
                // every call of the batch is scored as long as the chain is alive (and none once it broke)
                const bool all = st->ratchet && bb_passes_logl(st, -(double)ratchet_stale);
                for (int v = 0; v < count; v++) {
                    if (st->ratchet ? all : bb_passes(st, visit_rooted ? vis_score[v] : cur_score)) {
                        call_of[(size_t)v + vbegin[v]] = (int32_t)pass_cands.size();
                        pass_cands.push_back(visit_rooted ? -(2 + v) : -1);
                    }
                    for (int j = vbegin[v]; j < vbegin[v + 1]; j++)
                        if (st->ratchet ? all : bb_passes(st, mp[j])) { call_of[(size_t)v + 1 + j] = (int32_t)pass_cands.size(); pass_cands.push_back(j); }
                }
                thr.resize(st->B);
                if (st->policy == MPGPU_BB_MULHITS_TOP) {
                    // :3540 acts on top_count < N || rell > boot_threshold and never looks at boot_logl; lists only fill up and
                    // thresholds only rise within a batch, so the values at batch start give a safe superset
                    for (int b = 0; b < st->B; b++)
                        thr[b] = st->top_count[b] < st->top_n ? 2147483647
                               : (st->boot_threshold[b] <= -2147483647 ? 2147483647 : -st->boot_threshold[b] - 1);
                } else if (st->policy == MPGPU_BB_DISTINCT_ITER) {
                    // :3588 acts on rell >= boot_threshold; a threshold is the minimum of a list whose entries are only ever replaced by
                    // better ones or joined by entries that reached it: it only rises within a batch
                    for (int b = 0; b < st->B; b++) thr[b] = st->boot_threshold[b] <= -2147483647 ? 2147483647 : -st->boot_threshold[b];
                } else
                for (int b = 0; b < st->B; b++) thr[b] = bb_screen_of(st->boot_logl[b], st->ufboot_epsilon);
                if (st->updates_off && !st->ratchet) {     // -min_iter1_cand, iteration 1: the calls only extend treels_logl, no REPS vector is used
                    ro.dense.clear(); ro.dense_off.assign(pass_cands.size(), -1); ro.orig.clear(); ro.keep_slot.clear();
                } else
                if (int rc = reps_run(c, pass_cands.data(), (int)pass_cands.size(), thr.data(), ro)) return rc;
            }
            prof.stop(2); prof.start();
            bool moved = false;
            int v = 0;
            for (; v < count && !moved; v++) {
                int insertNode = 0, removeNode = 0;
                unsigned long bestTreeScoreHits = 1;              // :3303
                auto save_call = [&](int k, uint32_t m, int remove_ref, int insert_ref) {
                    bb->st->n_calls++;
                    if (!bb->st->ratchet) { if (k >= 0) bb_save(bb, c, -(double)m, ro, k, remove_ref, insert_ref); return; }
                    const double cur_logl = -(double)ratchet_stale;               // iqtree.cpp:3283-3294
                    if (k < 0 || !bb_passes_logl(bb->st, cur_logl)) return;
                    bb_save(bb, c, cur_logl, ro, k, remove_ref, insert_ref);
                    ratchet_stale = ro.orig[k];                                   // _pattern_pars now holds this tree's vector
                };
                if (bb) save_call(call_of[(size_t)v + vbegin[v]], visit_rooted ? vis_score[v] : cur_score, 0, 0);   // rearrangeParsimony :2286-2289
                for (int j = vbegin[v]; j < vbegin[v + 1]; j++) {
                    const uint32_t m = mp[j];
                    scored++;
                    if (sk_early && c->sk.h_est[j] > bestParsimony) continue;
                    if (bb) save_call(call_of[(size_t)v + 1 + j], m, cprune[j], cref[j]);        // testInsertParsimony :2163-2166
                    if (m < bestParsimony) bestTreeScoreHits = 1;                 // :2168
                    else if (m == bestParsimony) bestTreeScoreHits++;
                    if (m < bestParsimony || (m == bestParsimony && rng(rng_user) <= 1.0 / bestTreeScoreHits)) {
                        bestParsimony = m; insertNode = cref[j]; removeNode = cprune[j];
                    }
                }
                if (bb && bb->rc) return bb->rc;
                if (bestParsimony == randomMP) bestIterationScoreHits++;          // :3306
                if (bestParsimony < randomMP) bestIterationScoreHits = 1;
                if ((bestParsimony < randomMP ||
                     (bestParsimony == randomMP && rng(rng_user) <= 1.0 / bestIterationScoreHits)) &&
                    removeNode && insertNode) {
                    const int pa = c->tree.back(c->tree.next(removeNode)), pb = c->tree.back(c->tree.next(c->tree.next(removeNode)));
                    const int touched[5] = {removeNode / 3, pa / 3, pb / 3, insertNode / 3, c->tree.back(insertNode) / 3};
                    apply_spr_move(c->tree, removeNode, insertNode);              // :3312
                    if (lazy) mark_stale_nodes(c, touched, 5);
                    if (c->ref_valid) for (int k5 = 0; k5 < 5; k5++) scan_ref_fill(c->tree, c->ref_vstride, c->ref_table.data(), touched[k5]);
                    randomMP = bestParsimony;
                    cur_score = bestParsimony;
                    moved = true;
                    c->search_moves++;
                }
            }
            i += v;
            prof.stop(3);
            if (moved) {
                prof.start();
                // lazy views: the move only marked the views it invalidated; the next batch recomputes the ones it reads.
                // Eager scheme (MPGPU_EAGER_VIEWS): every stale view now, behind the host's back -- the next batch is
                // planned and launched while k_fitch_wave runs, its counts land with that batch's read-back (finish_scan)
                c->tree_set = true;
                if (!lazy) {
                    c->lens_valid = false;
                    if (int rc = update_views(c, true)) return rc;
                    if (!c->wave_pending) compute_lengths(c);
                }
                move_gap = 0.75 * move_gap + 0.25 * (double)(since_move + v);   // visits since the previous move
                since_move = 0;
                batch = after_move_batch(move_gap);
                prof.stop(4);
            } else {
                since_move += v;
                batch = std::min(batch * 4, nvisit);
            }
        }
    } while (randomMP < startMP);
    if ((double)since_move > move_gap) move_gap = 0.75 * move_gap + 0.25 * (double)since_move;   // a quiet search: speculate deeper next time
    if (c->wave_pending) { if (int rc = fetch_wave_counts(c)) return rc; MPGPU_CUDA(cudaStreamSynchronize(c->stream)); settle_views(c, true); }
    if (cumulative) {
        static bool registered = false;
        if (!registered) { registered = true; atexit([]() { g_sp.report(prof_names, 6); g_rp.report(g_rp_names, 8); }); }
    } else {
        prof.report(prof_names, 6);
        if (prof.on && c->lazy_lists)
            fprintf(stderr, "[mpgpu profile] lazy views: %lld lists, %.1f views in %.1f levels each (context totals)\n", (long long)c->lazy_lists,
                    (double)c->lazy_views / c->lazy_lists, (double)c->lazy_levels / c->lazy_lists);
        g_rp.report(g_rp_names, 8);
        for (int k = 0; k < 8; k++) { g_rp.t[k] = 0; g_rp.n[k] = 0; }
    }
    memcpy(back_node, c->tree.bn.data(), c->tree.bn.size() * sizeof(int32_t));
    memcpy(back_slot, c->tree.bs.data(), c->tree.bs.size() * sizeof(int32_t));
    *best = startMP;
    if (n_insertions) *n_insertions = scored;
    if (bb) bb->st->ratchet_last_score = ratchet_stale;
    return 0;
}

// PLL's private generator (pllrepo/src/utils.c:335-357): a 28-bit multiplicative congruential
// generator worked in 12-bit limbs; returns a double in [0,1) and advances *seed.
static double pll_randum(int64_t *seed)
{
    const int64_t s0 = *seed & 4095, s1 = (*seed >> 12) & 4095, s2 = (*seed >> 24) & 255;
    int64_t acc = 1549 * s0;
    const int64_t n0 = acc & 4095;
    acc = (acc >> 12) + 1549 * s1 + 406 * s0;
    const int64_t n1 = acc & 4095;
    acc = (acc >> 12) + 1549 * s2 + 406 * s1;
    const int64_t n2 = acc & 255;
    *seed = (n2 << 24) | (n1 << 12) | n0;
    return 0.00390625 * ((double)n2 + 0.000244140625 * ((double)n1 + 0.000244140625 * (double)n0));
}

// _pllMakeParsimonyTreeFast (sprparsimony.cpp:3107-3209), stepwise phase on the device:
// for each new taxon every branch of the partial tree is scored in one launch (k_tip_insert),
// the host walks stepwiseAddition's DFS (:2977-3019, recursion below q only while the subtree
// behind q has a positive length) over those numbers with the reference's tie-break draws.
static int stepwise_phase(mpgpu_ctx *c, int64_t *seed, mpgpu_rng_fn rng, void *rng_user, uint32_t *best_out, int64_t *scored)
{
    const int n = c->n;
    HostTree &t = c->tree;
    t.n = n;
    c->ref_valid = false;
    t.bn.assign(3 * (2 * n - 1), 0); t.bs.assign(3 * (2 * n - 1), 0);
    std::vector<int> perm(n + 2);
    for (int i = 1; i <= n; i++) perm[i] = i;                                   // makePermutationFast :2221
    for (int i = 1; i <= n; i++) {
        const int k = (int)((double)(n + 1 - i) * pll_randum(seed));
        std::swap(perm[i], perm[i + k]);
    }
    int nextnode = n + 1;
    // buildSimpleTree :1968: tips ip, iq joined, the first inner node hangs ir and is inserted between them
    const int ip = perm[1], iq = perm[2], ir = perm[3];
    const int f = 3 * std::min(ip, std::min(iq, ir));                            // tr->start
    {
        const int s = 3 * nextnode++;
        t.hookup(3 * ir, s);
        t.hookup(s + 1, 3 * ip);
        t.hookup(s + 2, 3 * iq);
    }
    c->tree_set = true; c->lens_valid = false;
    if (int rc = compute_views(c)) return rc;
    if (!c->reduces()) { set_error("stepwise addition on a sharded context needs mpgpu_set_allreduce"); return 1; }
    compute_lengths(c);
    uint32_t treelen = 0;
    const bool sk = c->sk.on;
    if (sk) { if (int rc = sk_tree_score(c, f, &treelen)) return rc; }
    else {
        uint32_t mis = 0;
        MPGPU_CUDA(cudaMemsetAsync(c->d_scalar, 0, sizeof(uint32_t), c->stream));
        if (int rc = launch_edge_mismatch(c, t.vid(f), t.vid(t.back(f)), c->d_scalar)) return rc;
        if (int rc = shard_sum(c, c->d_scalar, 1)) return rc;
        MPGPU_CUDA(cudaMemcpyAsync(&mis, c->d_scalar, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        MPGPU_CUDA(cudaStreamSynchronize(c->stream));
        treelen = mis + c->vlen[t.vid(f)] + c->vlen[t.vid(t.back(f))];
    }
    std::vector<int4> edges;
    std::vector<int32_t> edge_of_ref(3 * (2 * n - 1)), ins;
    std::vector<int> stack;
    int4 *d_edges = nullptr; size_t edges_cap = 0;
    int32_t *d_ins = nullptr; size_t ins_cap = 0;
    unsigned long hits = 1;
    uint32_t best = treelen;
    int rc = 0;
    for (int ntips = 4; ntips <= n && !rc; ntips++) {
        const int p = 3 * perm[ntips];                                           // the new tip
        const int q = 3 * nextnode++;                                            // its inner node, slot 0 on the tip
        // every branch of the current tree, keyed by the ref on the far side from tr->start
        edges.clear(); stack.clear();
        stack.push_back(t.back(f));
        while (!stack.empty()) {
            const int x = stack.back(); stack.pop_back();
            edge_of_ref[x] = (int)edges.size();
            edges.push_back(make_int4(t.vid(x), t.vid(t.back(x)), t.vid(p), 0));
            if (!t.is_tip(x)) { stack.push_back(t.back(t.next(t.next(x)))); stack.push_back(t.back(t.next(x))); }
        }
        const int ne = (int)edges.size();
        if (sk) {                                        // junction(x, back(x), new tip) = the whole tree's score (no early exit: :951 `!stepwiseAddition_on`)
            if ((rc = sk_junctions(c, edges.data(), ne, nullptr, true))) break;    // (an asymmetric matrix: rooted at the new tip, :2994-2998)
            ins.resize(ne);
            for (int e2 = 0; e2 < ne; e2++) ins[e2] = (int32_t)(c->sk.h_tot[e2].x - treelen);
        } else {
        if ((rc = ensure(d_edges, edges_cap, (size_t)ne))) break;
        if ((rc = ensure(d_ins, ins_cap, (size_t)ne))) break;
        ins.resize(ne);
        cudaError_t e = cudaMemcpyAsync(d_edges, edges.data(), (size_t)ne * sizeof(int4), cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_ins, 0, (size_t)ne * 4, c->stream);
        if (e != cudaSuccess) { rc = cuda_fail(e, "stepwise addition upload"); break; }
        if ((rc = launch_tip_insert(c, d_edges, ne, d_ins))) break;
        if ((rc = shard_sum(c, d_ins, ne))) break;
        e = cudaMemcpyAsync(ins.data(), d_ins, (size_t)ne * 4, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess && fetch_wave_counts(c)) { rc = 1; break; }
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { rc = cuda_fail(e, "stepwise addition read-back"); break; }
        }
        settle_views(c, true);                                                    // counts of the previous step's view update
        // stepwiseAddition(tr, pr, q, f->back): pre-order, children only below a subtree of positive length
        best = 2147483647u;                                                       // tr->bestParsimony = INT_MAX :3141
        int insert_ref = 0;
        stack.clear(); stack.push_back(t.back(f));
        while (!stack.empty()) {
            const int x = stack.back(); stack.pop_back();
            const uint32_t mp = treelen + (uint32_t)ins[edge_of_ref[x]];
            (*scored)++;
            if (mp < best) hits = 1;
            else if (mp == best) hits++;
            if (mp < best || (mp == best && rng(rng_user) <= 1.0 / hits)) { best = mp; insert_ref = x; }
            if (!t.is_tip(x) && c->vlen[t.vid(x)] > 0) {                          // :3014
                stack.push_back(t.back(t.next(t.next(x))));
                stack.push_back(t.back(t.next(x)));
            }
        }
        const int r = t.back(insert_ref);                                         // :3156-3162
        t.hookup(p, q);
        t.hookup(q + 1, insert_ref);
        t.hookup(q + 2, r);
        treelen = best;
        c->lens_valid = false;
        if ((rc = update_views(c, true))) break;                                  // settled after the next step's read-back
        if (!c->wave_pending) compute_lengths(c);
    }
    if (!rc && c->wave_pending) {
        if (fetch_wave_counts(c)) rc = 1;
        else if (cudaStreamSynchronize(c->stream) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "stepwise addition");
        else settle_views(c, true);
    }
    if (d_edges) cudaFree(d_edges);
    if (d_ins) cudaFree(d_ins);
    *best_out = best;
    return rc;
}

extern "C" {

int mpgpu_stepwise_addition(mpgpu_ctx *c, int64_t *random_seed, int spr_dist, mpgpu_rng_fn rng, void *rng_user,
                            int32_t *back_node, int32_t *back_slot, uint32_t *best, int64_t *n_insertions)
{
    if (!c || !random_seed || !rng || !back_node || !back_slot || !best) { set_error("null argument"); return 1; }
    if (!c->d_views) { set_error("no alignment loaded"); return 1; }
    MPGPU_CUDA(cudaSetDevice(c->device));
    int64_t scored = 0;
    uint32_t b0 = 0;
    if (int rc = stepwise_phase(c, random_seed, rng, rng_user, &b0, &scored)) { c->tree_set = false; return rc; }
    memcpy(back_node, c->tree.bn.data(), c->tree.bn.size() * sizeof(int32_t));
    memcpy(back_slot, c->tree.bs.data(), c->tree.bs.size() * sizeof(int32_t));
    int64_t more = 0;
    if (int rc = optimize_impl(c, back_node, back_slot, 1, spr_dist, rng, rng_user, nullptr, best, &more, true, &b0)) return rc;
    if (n_insertions) *n_insertions = scored + more;
    return 0;
}

int mpgpu_search_info(mpgpu_ctx *c, uint32_t *start_score, int64_t *moves, int64_t *batches)
{
    if (!c) { set_error("null argument"); return 1; }
    if (start_score) *start_score = c->search_start_score;
    if (moves) *moves = c->search_moves;
    if (batches) *batches = c->search_batches;
    return 0;
}

int mpgpu_optimize_spr(mpgpu_ctx *c, int32_t *back_node, int32_t *back_slot, int mintrav, int maxtrav,
                       mpgpu_rng_fn rng, void *rng_user, uint32_t *best, int64_t *n_insertions)
{
    if (!c || !back_node || !back_slot || !rng || !best) { set_error("null argument"); return 1; }
    return optimize_impl(c, back_node, back_slot, mintrav, maxtrav, rng, rng_user, nullptr, best, n_insertions);
}

int mpgpu_refine_replicates(mpgpu_ctx *c, int B, const uint16_t *boot_samples, int stride,
                            int32_t *trees_bn, int32_t *trees_bs, int mintrav, int maxtrav,
                            mpgpu_rng_fn rng, void *rng_user, uint32_t *scores, int64_t *n_insertions)
{
    if (!c || !boot_samples || !trees_bn || !trees_bs || !rng || !scores) { set_error("null argument"); return 1; }
    if (!c->d_views) { set_error("no alignment loaded"); return 1; }
    if (B < 0 || stride < c->P) { set_error("boot_samples stride is smaller than the number of patterns"); return 1; }
    const std::vector<int32_t> saved = c->weights;
    const size_t ring = (size_t)3 * (2 * c->n - 1);
    std::vector<int32_t> w(c->P);
    int64_t total = 0;
    int rc = 0;
    for (int b = 0; b < B && !rc; b++) {
        const uint16_t *row = boot_samples + (size_t)b * stride;
        for (int i = 0; i < c->P; i++) w[i] = row[i];
        c->tree_set = false;                               // the planes change: no view of the previous sample survives
        if ((rc = mpgpu_set_weights(c, w.data()))) break;
        int64_t ins = 0;
        rc = optimize_impl(c, trees_bn + b * ring, trees_bs + b * ring, mintrav, maxtrav, rng, rng_user, nullptr, scores + b, &ins);
        total += ins;
    }
    c->tree_set = false;
    if (int rc2 = mpgpu_set_weights(c, saved.data())) { if (!rc) rc = rc2; }
    if (n_insertions) *n_insertions = total;
    return rc;
}

int mpgpu_optimize_spr_bb(mpgpu_ctx *c, int32_t *back_node, int32_t *back_slot, int mintrav, int maxtrav,
                          const mpgpu_bb_hooks *hooks, mpgpu_bb_state *state, uint32_t *best, int64_t *n_insertions)
{
    if (!c || !back_node || !back_slot || !hooks || !state || !best) { set_error("null argument"); return 1; }
    if (!hooks->random_double || !hooks->push_tree_logl || !hooks->materialize) { set_error("incomplete -bb hooks"); return 1; }
    if (!c->reps.loaded) { set_error("no replicates loaded (mpgpu_load_replicates)"); return 1; }
    if (state->B != (c->rep_count > 1 ? c->reps.B_total : c->reps.Buser) || !state->boot_logl || !state->boot_counts || !state->boot_trees) { set_error("bad -bb state"); return 1; }
    if (state->policy != MPGPU_BB_DEFAULT && state->policy != MPGPU_BB_MULHITS && state->policy != MPGPU_BB_MULHITS_TOP &&
        state->policy != MPGPU_BB_DISTINCT_ITER) { set_error("unknown -bb policy"); return 1; }
    const bool distinct = state->policy == MPGPU_BB_DISTINCT_ITER;
    if (distinct) {
        if (!hooks->disthit || state->top_n < 1 || !state->boot_threshold) { set_error("policy MPGPU_BB_DISTINCT_ITER needs the disthit hook, top_n >= 1 and boot_threshold"); return 1; }
        if (c->sk.on) { set_error("policy MPGPU_BB_DISTINCT_ITER is not available under -cost"); return 1; }
        if (c->shard_count > 1 || c->rep_count > 1) { set_error("policy MPGPU_BB_DISTINCT_ITER needs an unsharded context"); return 1; }
    }
    if (c->reps.keep_on != distinct) { c->reps.keep_on = distinct; c->reps.tree_valid = false; }
    if (state->policy == MPGPU_BB_MULHITS_TOP && (!hooks->tophit || state->top_n < 1 || !state->top_count || !state->boot_threshold)) {
        set_error("policy MPGPU_BB_MULHITS_TOP needs the tophit hook, top_n >= 1, top_count and boot_threshold"); return 1;
    }
    if (state->policy == MPGPU_BB_MULHITS && !hooks->mulhit) { set_error("policy MPGPU_BB_MULHITS needs the mulhit hook"); return 1; }
    state->n_calls = 0; state->n_reps = 0;
    BBRun bb{hooks, state, {}, {}, {}, 0};
    bb.screen.resize((size_t)state->B);
    for (int b = 0; b < state->B; b++) bb.screen[b] = bb_screen_of(state->boot_logl[b], state->ufboot_epsilon);
    return optimize_impl(c, back_node, back_slot, mintrav, maxtrav, hooks->random_double, hooks->user, &bb, best, n_insertions);
}

int mpgpu_set_option(mpgpu_ctx *c, const char *name, int value)
{
    if (!c || !name) { set_error("null argument"); return 1; }
    if (!strcmp(name, "reps_tensor")) {
        if (c->reps.loaded) { set_error("reps_tensor must be set before mpgpu_load_replicates"); return 1; }
        c->reps.use_tensor = value != 0;
        return 0;
    }
    if (!strcmp(name, "reps_nowrap")) { c->reps.nowrap = value != 0; return 0; }
    if (!strcmp(name, "sankoff_exact")) { c->sk.exact = value != 0; return 0; }
    if (!strcmp(name, "sankoff_u32")) {
        if (c->sk.on) { set_error("sankoff_u32 must be set before mpgpu_set_cost_matrix"); return 1; }
        c->sk.wide = value != 0;
        return 0;
    }
    if (!strcmp(name, "exchange")) { c->exchange_off = value == 0; return 0; }
    if (!strcmp(name, "reps_timing")) { c->reps.timing = value != 0; c->reps.timed_rows = 0; return 0; }
    set_error(std::string("unknown option: ") + name);
    return 1;
}

int mpgpu_reps_timing(mpgpu_ctx *c, float *tc_ms, int *rows, int *patterns, int *splits)
{
    if (!c || !c->reps.loaded || !c->reps.ev1 || c->reps.timed_rows == 0) { set_error("no timed k_reps_tc launch (option reps_timing)"); return 1; }
    MPGPU_CUDA(cudaEventSynchronize(c->reps.ev1));
    float ms = 0;
    MPGPU_CUDA(cudaEventElapsedTime(&ms, c->reps.ev0, c->reps.ev1));
    if (tc_ms) *tc_ms = ms;
    if (rows) *rows = c->reps.timed_rows;
    if (patterns) *patterns = c->reps.timed_kblocks * 128;
    if (splits) *splits = c->reps.timed_splits;
    c->reps.timed_rows = 0;
    return 0;
}

int mpgpu_int8_peak(mpgpu_ctx *c, int iters, double *tops)
{
    if (!c || !tops || iters < 1) { set_error("bad argument"); return 1; }
    MPGPU_CUDA(cudaSetDevice(c->device));
    return measure_int8_peak(c, iters, tops);
}

int mpgpu_reps_info(mpgpu_ctx *c, int *groups, int *exceptions, int *tensor)
{
    if (!c || !c->reps.loaded) { set_error("no replicates loaded"); return 1; }
    if (groups) *groups = c->reps.G;
    if (exceptions) *exceptions = c->reps.n_exc;
    if (tensor) *tensor = c->reps.use_tensor ? 1 : 0;
    return 0;
}

int mpgpu_sankoff_reps_stats(mpgpu_ctx *c, int64_t *tensor_chunks, int64_t *exact_chunks)
{
    if (!c) { set_error("null context"); return 1; }
    if (tensor_chunks) *tensor_chunks = c->sk.tensor_chunks;
    if (exact_chunks) *exact_chunks = c->sk.exact_chunks;
    return 0;
}

int mpgpu_load_replicates(mpgpu_ctx *c, int B, const uint16_t *boot, int stride, const int32_t *segment_upper, int nseg)
{
    return mpgpu_load_replicates2(c, B, boot, stride, segment_upper, nseg, nullptr);
}

int mpgpu_load_replicates2(mpgpu_ctx *c, int B, const uint16_t *boot, int stride, const int32_t *segment_upper, int nseg,
                           const uint16_t *original_sample)
{
    if (!c || !boot || !segment_upper) { set_error("null argument"); return 1; }
    if (!c->d_codes) { set_error("no alignment loaded"); return 1; }
    if (B < 1 || nseg < 1) { set_error("need at least one replicate and one segment"); return 1; }
    if (c->sk.on && !c->sort_alignment) { set_error("-cost with -bb needs sort_alignment (informative patterns first)"); return 1; }
    MPGPU_CUDA(cudaSetDevice(c->device));
    free_reps(c);
    Reps &r = c->reps;
    const int B_all = B;
    int rep_lo = 0;
    if (c->rep_count > 1) {                  // replicate shards: this context keeps replicates [lo, hi) (rows of `boot`)
        rep_lo = (int)((int64_t)B * c->rep_rank / c->rep_count);
        const int rep_hi = (int)((int64_t)B * (c->rep_rank + 1) / c->rep_count);
        if (rep_hi <= rep_lo) { set_error("fewer replicates than replicate shards"); return 1; }
        boot += (size_t)rep_lo * stride;
        B = rep_hi - rep_lo;
    }
    const int upper0 = c->sort_alignment ? c->n_inf : c->P;
    if (stride < upper0) { set_error("boot_samples stride is smaller than the number of reported patterns"); return 1; }
    for (int s = 0; s < nseg; s++) {
        if (segment_upper[s] <= (s ? segment_upper[s - 1] : 0)) { set_error("segment_upper must be increasing"); return 1; }
        if (s < nseg - 1 && segment_upper[s] % 16) { set_error("inner segment bounds must be multiples of 16 (iqtree.cpp:3806)"); return 1; }
    }
    r.Buser = B; r.has_orig = original_sample != nullptr;
    r.B = B + (r.has_orig ? 1 : 0); r.Bpad = (r.B + 255) / 256 * 256;
    r.upper = std::min(upper0, (int)segment_upper[nseg - 1]);       // the loop at :3424 stops at the last bound
    r.Kpad = (std::max(r.upper, 1) + 127) / 128 * 128; r.Pw = r.Kpad / 32;
    r.seg_upper.assign(segment_upper, segment_upper + nseg);

    // ---- device copies: exact weights pattern-major, heavy flags, segment bounds ----
    const int nseg_ = nseg;
    r.seg_flagged.assign(nseg_, 0);
    r.heavy.assign(std::max(r.upper, 1), 0);
    uint16_t *d_boot16 = nullptr; uint8_t *d_heavy = nullptr;
    MPGPU_CUDA(cudaMalloc((void **)&r.d_w8, (size_t)r.Bpad * r.Kpad));
    MPGPU_CUDA(cudaMalloc((void **)&r.d_w16T, (size_t)std::max(r.upper, 1) * r.Bpad * sizeof(uint16_t)));
    MPGPU_CUDA(cudaMalloc((void **)&r.d_seg_upper, (size_t)nseg_ * 4));
    MPGPU_CUDA(cudaMalloc((void **)&r.d_segmax, ((size_t)nseg_ + 4) * 4));      // + slack: the peer exchange moves whole int4
    MPGPU_CUDA(cudaMalloc((void **)&d_boot16, (size_t)r.B * stride * sizeof(uint16_t)));
    MPGPU_CUDA(cudaMalloc((void **)&d_heavy, r.heavy.size()));
    MPGPU_CUDA(cudaMemcpyAsync(r.d_seg_upper, segment_upper, (size_t)nseg_ * 4, cudaMemcpyHostToDevice, c->stream));
    MPGPU_CUDA(cudaMemcpyAsync(d_boot16, boot, (size_t)B * stride * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream));
    if (r.has_orig) {                                     // one more "replicate": the original frequencies
        r.original_sample.assign(original_sample, original_sample + std::max(r.upper, 1));
        MPGPU_CUDA(cudaMemsetAsync(d_boot16 + (size_t)B * stride, 0, (size_t)stride * sizeof(uint16_t), c->stream));
        MPGPU_CUDA(cudaMemcpyAsync(d_boot16 + (size_t)B * stride, original_sample, (size_t)r.upper * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream));
    }
    MPGPU_CUDA(cudaMemsetAsync(d_heavy, 0, r.heavy.size(), c->stream));
    int rc = launch_transpose_boot(c, d_boot16, stride, d_heavy);
    cudaError_t e = cudaMemcpyAsync(r.heavy.data(), d_heavy, r.heavy.size(), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_boot16); cudaFree(d_heavy);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "uploading replicate weights");
    r.G = 1;
    r.loaded_sankoff = c->sk.on;
    if (!c->sk.on) {
        if (int rc2 = build_classification(c)) return rc2;
        r.reclassifications = 0;
        if (c->shard_count == 1 && r.upper > 0) {           // seg_wmax: the wrap check against an all-zero score vector = max_b sum w
            const int upper1 = c->sort_alignment ? c->n_inf : c->P;
            if (int rc2 = ensure(c->d_ptn, c->ptn_cap, (size_t)(upper1 > 0 ? upper1 : 1))) return rc2;
            MPGPU_CUDA(cudaMemsetAsync(c->d_ptn, 0, (size_t)(upper1 > 0 ? upper1 : 1) * sizeof(uint16_t), c->stream));
            MPGPU_CUDA(cudaMemsetAsync(r.d_segmax, 0, (size_t)nseg_ * 4, c->stream));
            r.p_lo = 0; r.p_hi = r.upper;
            if (int rc2 = launch_seg_check(c, r.d_segmax)) return rc2;
            r.seg_wmax.assign(nseg_, 0);
            MPGPU_CUDA(cudaMemcpyAsync(r.seg_wmax.data(), r.d_segmax, (size_t)nseg_ * 4, cudaMemcpyDeviceToHost, c->stream));
            MPGPU_CUDA(cudaStreamSynchronize(c->stream));
        } else r.seg_wmax.clear();
        if (r.use_tensor) { if (int rc2 = make_w8_tensor_map(c)) return rc2; }
    } else {
        // -cost: the rows are per-pattern costs (sankoff.cu).  The u8 operand is only usable when no replicate weight
        // exceeds 255; whether a chunk may take the tensor path is decided per chunk (costs <= 255, no segment can wrap).
        r.n_heavy = 0;
        for (int p = 0; p < r.upper; p++) r.n_heavy += r.heavy[p] ? 1 : 0;
        r.n_exc = 0;
        if (r.use_tensor && r.n_heavy == 0) {
            uint8_t *d_is_exc = nullptr;
            MPGPU_CUDA(cudaMalloc((void **)&d_is_exc, (size_t)r.Kpad));
            MPGPU_CUDA(cudaMemsetAsync(d_is_exc, 0, (size_t)r.Kpad, c->stream));
            int rc2 = launch_build_w8(c, d_is_exc);
            cudaError_t e2 = cudaStreamSynchronize(c->stream);
            cudaFree(d_is_exc);
            if (rc2) return rc2;
            if (e2 != cudaSuccess) return cuda_fail(e2, "building the tensor operand");
            if (int rc3 = make_w8_tensor_map(c)) return rc3;
        }
    }
    r.rep_lo = rep_lo; r.B_total = B_all;
    r.loaded = true;
    r.tree_valid = false;
    return 0;
}

int mpgpu_set_remain_bounds(mpgpu_ctx *c, const int32_t *bounds, int per_replicate)
{
    if (!c) { set_error("null context"); return 1; }
    Reps &r = c->reps;
    if (!r.loaded) { set_error("no replicates loaded (mpgpu_load_replicates)"); return 1; }
    if (c->shard_count > 1 || c->rep_count > 1) { set_error("mpgpu_set_remain_bounds needs an unsharded context"); return 1; }
    MPGPU_CUDA(cudaSetDevice(c->device));
    if (!bounds) { r.remain_loaded = false; return 0; }
    const int nseg = (int)r.seg_upper.size();
    if (per_replicate != nseg - 1) { set_error("remain bounds: need nseg - 1 values per replicate (iqtree.cpp:3842)"); return 1; }
    r.remain_loaded = false;
    if (nseg < 2) return 0;                                   // a single segment is never tested (:3433)
    if (r.d_remain) { cudaFree(r.d_remain); r.d_remain = nullptr; }
    MPGPU_CUDA(cudaMalloc((void **)&r.d_remain, (size_t)r.Buser * (nseg - 1) * 4));
    MPGPU_CUDA(cudaMemcpyAsync(r.d_remain, bounds, (size_t)r.Buser * (nseg - 1) * 4, cudaMemcpyHostToDevice, c->stream));
    MPGPU_CUDA(cudaStreamSynchronize(c->stream));
    r.remain_loaded = true;
    return 0;
}

int mpgpu_reps_prefix_max(mpgpu_ctx *c, int32_t cand_idx, const int32_t *samples, int m, int32_t *out)
{
    if (int rc = need_tree(c, true)) return rc;
    if (!samples || !out || m < 0) { set_error("bad argument"); return 1; }
    Reps &r = c->reps;
    if (!r.loaded || !r.remain_loaded) { set_error("mpgpu_reps_prefix_max needs replicates and remain bounds (mpgpu_set_remain_bounds)"); return 1; }
    if (c->sk.on || c->shard_count > 1 || c->rep_count > 1) { set_error("mpgpu_reps_prefix_max: Fitch scoring on an unsharded context only"); return 1; }
    for (int i = 0; i < m; i++) if (samples[i] < 0 || samples[i] >= r.Buser) { set_error("replicate index out of range"); return 1; }
    MPGPU_CUDA(cudaSetDevice(c->device));
    const bool was = r.keep_on;
    if (!was) { r.keep_on = true; r.tree_valid = false; }
    RepsOut ro;
    int rc = reps_run(c, &cand_idx, 1, nullptr, ro);
    if (!rc) rc = prefix_max(c, ro.keep_slot[0], samples, m, out);
    if (!was) { r.keep_on = false; }
    return rc;
}

int mpgpu_set_replicate_shards(mpgpu_ctx *c, int rank, int count)
{
    if (!c) { set_error("null context"); return 1; }
    if (count < 1 || rank < 0 || rank >= count) { set_error("bad shard arguments"); return 1; }
    if (c->shard_count != 1) { set_error("replicate shards are for contexts that hold the whole alignment (shard_count = 1 at mpgpu_create)"); return 1; }
    if (c->reps.loaded) { set_error("mpgpu_set_replicate_shards must precede mpgpu_load_replicates"); return 1; }
    if (c->peer.ready || c->peer.region) { set_error("mpgpu_set_replicate_shards must precede mpgpu_peer_prepare"); return 1; }
    c->rep_rank = rank; c->rep_count = count;
    return 0;
}

int mpgpu_reps_current_tree(mpgpu_ctx *c, int32_t *res)
{
    if (int rc = need_tree(c, true)) return rc;
    if (!res) { set_error("null argument"); return 1; }
    if (!c->reps.loaded) { set_error("no replicates loaded (mpgpu_load_replicates)"); return 1; }
    if (!c->reduces()) { set_error("mpgpu_reps_current_tree on a sharded context needs mpgpu_set_allreduce"); return 1; }
    MPGPU_CUDA(cudaSetDevice(c->device));
    // the current tree needs no scan plan: a one-call batch with an empty plan
    const int32_t cand = -1;
    RepsOut ro;
    if (int rc = reps_run(c, &cand, 1, nullptr, ro)) return rc;
    memcpy(res, ro.dense.data() + ro.dense_off[0], (size_t)(c->rep_count > 1 ? c->reps.B_total : c->reps.Buser) * 4);
    return 0;
}

int mpgpu_reps_candidates_device(mpgpu_ctx *c, const int32_t *cand_idx, int m, void **dev_res, int *pitch)
{
    if (int rc = need_tree(c, true)) return rc;
    if (!cand_idx || m < 0) { set_error("bad argument"); return 1; }
    if (!c->reps.loaded) { set_error("no replicates loaded (mpgpu_load_replicates)"); return 1; }
    if (!c->reduces()) { set_error("mpgpu_reps_candidates_device on a sharded context needs mpgpu_set_allreduce"); return 1; }
    MPGPU_CUDA(cudaSetDevice(c->device));
    RepsOut ro;
    if (int rc = reps_run(c, cand_idx, m, nullptr, ro, true)) return rc;
    if (dev_res) *dev_res = (void *)c->reps.d_res;
    if (pitch) *pitch = c->reps.Bpad;
    return 0;
}

int mpgpu_reps_candidates(mpgpu_ctx *c, const int32_t *cand_idx, int m, int32_t *res)
{
    if (int rc = need_tree(c, true)) return rc;
    if (!cand_idx || !res || m < 0) { set_error("bad argument"); return 1; }
    if (!c->reps.loaded) { set_error("no replicates loaded (mpgpu_load_replicates)"); return 1; }
    if (!c->reduces()) { set_error("mpgpu_reps_candidates on a sharded context needs mpgpu_set_allreduce"); return 1; }
    MPGPU_CUDA(cudaSetDevice(c->device));
    RepsOut ro;
    if (int rc = reps_run(c, cand_idx, m, nullptr, ro)) return rc;
    const size_t Bu = (size_t)(c->rep_count > 1 ? c->reps.B_total : c->reps.Buser);
    for (int i = 0; i < m; i++) memcpy(res + (size_t)i * Bu, ro.dense.data() + ro.dense_off[i], Bu * 4);
    return 0;
}

}  // extern "C"
