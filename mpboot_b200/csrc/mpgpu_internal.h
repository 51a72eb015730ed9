// Internal declarations shared by the CUDA kernels, the C-ABI layer and the host-side SPR
// driver of libmpgpu.so.  Nothing here is part of the public boundary (include/mpgpu.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/mpgpu.h"

namespace mpgpu {

// Device words per plane are padded to a multiple of this (one warp-iteration of 128-bit lanes).
static const int kWordPad = 128;
// One unit of work of the scan kernel: 32 words (1024 expanded sites) of every plane.
static const int kChunkWords = 32;
static const int kMaxStates = 32;
static const int kMaxTrav = 12;          // deepest SPR radius the scan kernel's stack supports

// ---- host mirror of the PLL node rings (pllrepo/src/pll.h:687-702) ---------------------
struct HostTree {
    int n = 0;                            // tips
    std::vector<int32_t> bn, bs;          // ring tables, 3*(2n-1) entries
    int back(int ref) const { return 3 * bn[ref] + bs[ref]; }
    bool has_back(int ref) const { return bn[ref] != 0; }
    bool is_tip(int ref) const { return ref / 3 <= n; }
    int next(int ref) const {
        int node = ref / 3;
        if (node <= n) return ref;
        return 3 * node + (ref % 3 + 1) % 3;
    }
    int vid(int ref) const {              // directed-view id of the subtree behind `ref`
        int node = ref / 3;
        return node <= n ? node - 1 : n + 3 * (node - n - 1) + ref % 3;
    }
    int num_views() const { return 4 * n - 6 + 0 * n; }
    void hookup(int a, int b) {           // hookupDefault (pllrepo/src/utils.c:456)
        bn[a] = b / 3; bs[a] = b % 3;
        bn[b] = a / 3; bs[b] = a % 3;
    }
};

// One Fitch combine dst = fitch(a, b) of the level schedule
struct Triple { int32_t dst, a, b, pad; };

// ---- scan program (built on the host, interpreted by k_spr_scan) ------------------------
// One "expand" op per expanded node, as two parallel streams (see k_spr_scan):
struct ScanOffs { int32_t c1, c2; };   // views of the two children (pointing at the expanded node) as
                                       // offsets in vector units (vid * view_stride / SG): the kernel adds
struct ScanCtl {
    uint32_t outs;      // out1 | out2 << 16: candidate index relative to the task's first (0xFFFF: not scored)
    uint32_t meta;      // src | dst1 << 8 | dst2 << 16: stack slots (0xFF none); src 0xFF / 0xFE = the
                        // task's D2 / D1 view (top-level expansions)
};
struct ScanTask {
    int32_t s_vid;      // pruned subtree (view offset in vector units)
    int32_t d1, d2;     // the two views that become neighbours when the node is removed (offsets)
    int32_t op_begin, op_end;
    int32_t base_out;   // output slot of popc(~any(D1&D2)) (length of the joined edge) = the task's index; -1: a sub-task
                        // that leaves the count to the task's first sub-task
    int32_t cand_base;  // output slot of the task's first candidate = task_cap + its candidate index
    int32_t pad;
};

// One record per ring slot for the enumeration of a scan plan: the two slots behind an inner slot, the view offset (vector units)
// and tip = view id << 2 | noted-by-the-current-plan << 1 | is-tip
struct ScanRef { int32_t c1, c2, voff, tip; };

// Page-locked (and device-mapped: k_publish writes wave counts straight into one) host array: the plan streams are uploaded piece by piece while the host keeps
// appending, so the copies must be truly asynchronous (pageable memory is staged synchronously).
template <typename T> struct PinnedArray {
    T *p = nullptr; size_t cap = 0;
    PinnedArray() {}
    ~PinnedArray() { if (p) cudaFreeHost(p); }
    size_t size() const { return cap; }
    T *data() { return p; }
    bool reserve(size_t n) {
        if (n <= cap) return true;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        if (cudaHostAlloc((void **)&p, n * sizeof(T), cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); p = nullptr; return false; }
        cap = n;
        return true;
    }
private:
    PinnedArray(const PinnedArray &);
    PinnedArray &operator=(const PinnedArray &);
};

struct ScanPlan {
    PinnedArray<ScanOffs> offs;
    PinnedArray<ScanCtl> ctl;
    PinnedArray<ScanTask> tasks_pin;      // upload copy of `tasks`
    std::vector<ScanOffs> offs_host;      // host-only plans (mpgpu_host_enumerate: no device, no page-locked memory)
    std::vector<ScanCtl> ctl_host;
    std::vector<ScanTask> tasks;
    std::vector<int32_t> visit_begin;     // count+1
    std::vector<int32_t> visit_ref;       // the ring slot each planned visit prunes at (tr->nodep[i])
    std::vector<int32_t> cand_ref, cand_prune, cand_task;
    std::vector<int32_t> task_vids;       // view ids of S, D1, D2 per task
    std::vector<ScanTask> sub_tasks;      // latency path: the tasks cut into independent sub-tasks (ScanPlanner::split); device copy behind the tasks
    std::vector<int32_t> task_tops;       // [2 * task] the task's top-level ops (-1: that side is a tip), recorded for split plans
    std::vector<int32_t> kid_op;          // scratch of split plans: op tree
    size_t item_cap = 0;                  // ScanTask slots of the device array (tasks + sub-tasks)
    std::vector<int32_t> need_refs;       // ring slots whose (stale) views the plan reads, duplicates possible (lazy views)
    std::vector<uint32_t> task_const;     // len(S)+len(D1)+len(D2) per task (filled by finish_scan from the view lengths)
    int n_cand = 0;                      // candidates / ops actually used (the vectors are sized to an upper bound)
    int n_ops = 0;
    int max_slot = 0;
    int task_cap = 0;                    // base slots in front of the candidate counts (2 per planned visit)
};

// ---- replicate scoring state (R8; reps_kernels.cu, mpgpu_bb.cu) -------------------------------
static const int kTreeRows = 17;          // rows 0..15: bit planes of the tree's per-site counters, row 16: their weighted sum
struct Reps {
    bool loaded = false;
    bool loaded_sankoff = false;          // loaded while a cost matrix was set: REPS runs on per-pattern cost rows (sankoff.cu)
    int rep_lo = 0, B_total = 0;          // replicate shards: first replicate held here, replicates over all shards (= Buser otherwise)
    int32_t *d_full = nullptr; size_t full_cap = 0;   // replicate shards: full-width rows [calls][Btot_pad] assembled by the exchange
    int B = 0, Bpad = 0;                  // weight columns (replicates, + 1 for original_sample), padded to the tensor tile (256)
    int Buser = 0;                        // replicates proper: columns [0, Buser); column Buser = original_sample when has_orig
    bool has_orig = false;
    std::vector<uint16_t> original_sample; // host copy [upper] (ratchet: score of the host's initial _pattern_pars)
    int upper = 0;                        // patterns that take part: min(numInformative or P, last segment_upper)
    int Kpad = 0, Pw = 0;                 // patterns padded to 128; words per pattern row
    std::vector<int32_t> seg_upper;
    int G = 1;                            // column groups: 0 = wrap-free bulk, g >= 1 = one wrap-prone segment each
    int n_exc = 0, n_heavy = 0;           // exception patterns (all of groups >= 1, plus weights > 255 of group 0)
    int kb_lo = 0, kb_hi = 0;             // 128-pattern K-blocks that can have bits on this shard
    int p_lo = 0, p_hi = 0;               // patterns whose first expanded site lies in this shard's word slice
    int32_t *d_segmax = nullptr;          // [nseg] wrap check: max over replicates of the segment bound
    bool use_tensor = true;
    bool nowrap = false;                  // option "reps_nowrap" (-autovec): the replicates' sums are plain ints, only original_sample's column keeps the 16-bit segment sums
    uint8_t *d_w8 = nullptr;              // [Bpad][Kpad] u8, K-major: the tensor kernel's B operand
    uint16_t *d_w16T = nullptr;           // [upper][Bpad] u16, pattern-major: the exact weights
    int32_t *d_seg_upper = nullptr;       // [nseg]
    std::vector<uint8_t> heavy;           // [upper] some replicate weight > 255
    std::vector<uint8_t> seg_flagged;     // [nseg] segment is in a group of its own (can wrap)
    std::vector<int32_t> seg_wmax;        // [nseg] max over replicates of the segment's weight sum (the wrap check at score 0)
    int32_t *d_exc_ptn = nullptr, *d_exc_group = nullptr; size_t exc_cap = 0, exc_group_cap = 0;
    alignas(64) unsigned char tmap_w8[128];
    bool tmap_valid = false;
    // rows (one index space for the three buffers)
    int row_cap = 0;
    uint32_t *d_rows_site = nullptr;      // [row_cap][Wl]   mismatch rows in expanded-site space
    uint32_t *d_rows_ptn = nullptr;       // [row_cap][Pw]   the same rows in pattern space
    int32_t *d_X = nullptr;               // [row_cap][G][Bpad]
    bool tree_valid = false;              // rows 0..16 describe the current tree and weights
    // per batch (host staging + device copies)
    std::vector<int32_t> h_row_of, h_row_tasks;
    std::vector<int4> h_edges;
    std::vector<int2> h_calls;
    int32_t *d_row_of = nullptr; size_t row_of_cap = 0;
    int32_t *d_row_tasks = nullptr; size_t row_tasks_cap = 0;
    int4 *d_edges = nullptr; size_t edges_cap = 0;
    int2 *d_calls = nullptr; size_t calls_cap = 0;
    int32_t *d_res = nullptr; size_t res_cap = 0;          // [calls][Bpad]
    int32_t *d_thr = nullptr;                              // [Bpad]
    int32_t *d_call_hit = nullptr; size_t call_hit_cap = 0;   // [calls] some replicate of the call passes its threshold
    int32_t *d_hit_list = nullptr; size_t hit_list_cap = 0;   // calls to read back
    int32_t *d_res_hit = nullptr; size_t res_hit_cap = 0;     // their rows, compacted
    void *h_pin = nullptr; size_t pin_cap = 0;              // pinned staging for the read-backs
    void *pinned(size_t bytes) {
        if (bytes <= pin_cap && h_pin) return h_pin;
        if (h_pin) cudaFreeHost(h_pin);
        h_pin = nullptr; pin_cap = 0;
        const size_t want = bytes + bytes / 2 + 4096;
        if (cudaHostAlloc(&h_pin, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); h_pin = nullptr; return nullptr; }
        pin_cap = want;
        return h_pin;
    }
    int64_t rows_scored = 0;              // statistics: rows pushed through the contraction
    bool timing = false;                  // option "reps_timing": CUDA events around the largest k_reps_tc launch
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int timed_rows = 0, timed_kblocks = 0, timed_splits = 0;
    int reclassifications = 0;            // times a wrap check moved a segment out of the tensor path's bulk group
    // -distinct_iter_top_boot (MPGPU_BB_DISTINCT_ITER): what the exact remain-bound skip needs after a batch's chunks are gone
    bool keep_on = false;                 // keep the two bit rows of every call that is read back, and the tree's per-pattern scores
    uint32_t *d_keep = nullptr; size_t keep_cap = 0;          // [slots][2][Pw]: edge row, delta row (pattern space)
    int keep_used = 0;                                         // slots in use by the current batch
    int32_t *d_keep_list = nullptr; size_t keep_list_cap = 0;
    uint16_t *d_tree_ptn = nullptr; size_t tree_ptn_cap = 0;   // [upper] per-pattern scores of the tree the kept rows belong to
    int32_t *d_remain = nullptr;                               // [Buser][nseg-1] boot_samples_pars_remain_bounds (mpgpu_set_remain_bounds)
    bool remain_loaded = false;
    int32_t *d_blist = nullptr; size_t blist_cap = 0;
    int32_t *d_segsum = nullptr; size_t segsum_cap = 0;
    int32_t *d_pmax = nullptr; size_t pmax_cap = 0;
    int64_t prefix_calls = 0, prefix_pairs = 0;                // statistics
};

// ---- Sankoff (-cost) state (R11; sankoff.cu) ------------------------------------------------------
struct Sankoff {
    bool on = false;                      // a cost matrix is set: every score is weighted parsimony
    bool exact = false;                   // option "sankoff_exact": perSiteScores mode, no lower-bound early exit (:951)
    bool cost_dirty = true;
    bool wide = false;                    // option "sankoff_u32" (-short_off): weights and segment sums are 32-bit, no 16-bit wrap
    uint32_t sum_mask() const { return wide ? 0xFFFFFFFFu : 0xFFFFu; }
    bool asym = false;                    // cost[i][j] != cost[j][i] somewhere: insertion and stepwise scores take the reference's rooted form
    std::vector<uint32_t> cost;           // pllCostMatrix [S][S]
    uint32_t highest = 0;                 // highest_cost = max + 1 (:160)
    std::vector<int32_t> seg_upper; int nseg = 0;
    std::vector<uint32_t> lb;             // pllRemainderLowerBounds [nseg-1]
    int Lref = 0;                         // the reference's vector length (informative patterns padded to 16)
    int Lp = 0, Lh = 0;                   // this shard's patterns per state row (whole chunks) and 32-bit words (pattern pairs)
    int Lp_glob = 0, pair0 = 0;           // patterns over all shards; first pattern pair of this shard
    size_t vstride = 0;                   // 32-bit words per view: S * Lh
    uint32_t *d_views = nullptr; size_t views_cap = 0;     // [4n-6][S][Lh] transformed cost vectors, u16x2
    uint2 *d_w = nullptr;                 // [Lh] weights of the pair's two patterns
    int32_t *d_seg = nullptr;             // [Lh] segment of the pair
    uint32_t *d_lb = nullptr;             // [nseg]
    uint32_t *d_mask = nullptr;           // [256] code -> state mask
    uint32_t *d_segout = nullptr; size_t segout_cap = 0;   // [rows][nseg] exact weighted sums per segment (mod 2^32)
    uint2 *d_tot = nullptr; size_t tot_cap = 0;            // [rows] (total, max_seg(prefix + lb))
    uint2 *h_tot = nullptr; size_t h_tot_cap = 0;          // pinned copy
    int4 *d_list = nullptr; size_t list_cap = 0;
    uint32_t *d_tmp = nullptr; size_t tmp_cap = 0;
    // -bb under -cost: per-pattern cost rows (row 0 = the current tree), their REPS, row bookkeeping
    uint32_t *d_rows = nullptr; size_t rows_cap = 0;       // [rows][Lh]
    int32_t *d_X = nullptr; size_t X_cap = 0;              // [rows][Bpad]
    int32_t *d_row_of = nullptr; size_t row_of_cap = 0;    // [n_cand]
    int32_t *d_call_row = nullptr; size_t call_row_cap = 0;
    uint32_t *d_stack = nullptr; size_t stack_cap = 0;     // S > 4: the scan kernel's per-warp stacks (resident warps only)
    uint8_t *d_rows8 = nullptr; size_t rows8_cap = 0;      // [rows][Kpad] the same rows as u8 (tensor path)
    uint32_t *d_colmax = nullptr; size_t colmax_cap = 0;   // [Lh] per-pattern maximum over the chunk's rows (u16x2) + [2] flags
    int64_t tensor_chunks = 0, exact_chunks = 0;           // statistics: which contraction the chunks took
    std::vector<uint32_t> h_est;          // per candidate of the last scan: max_seg(prefix + lb); > bestParsimony <=> the reference exits early
};

// ---- exchange step of sharded contexts over NVLink peer memory (peer_exchange.cu) -------------------
static const int kMaxPeers = 8;           // shards of one NVSwitch box
static const int kPeerBlocks = 32;        // CTAs of the one-shot all-reduce (one flag per block, rank and parity)
struct PeerExchange {
    bool ready = false;
    void *region = nullptr; size_t bytes = 0;      // this rank's slots[2][R][cap] int32 + flags[2][R][kPeerBlocks] u32
    size_t cap = 0;                                // ints per slot
    void *mapped[kMaxPeers] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // IPC mappings of the peers' regions
    int32_t *slots[kMaxPeers] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint32_t *flags[kMaxPeers] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint32_t epoch = 0;
    int *d_err = nullptr;                          // set by a block whose peer never arrived (bounded spin)
    int32_t *d_tmp = nullptr; size_t tmp_cap = 0;  // aligned scratch for vectors that are not whole int4
    int64_t calls = 0, elements = 0;
};

// Latency path of the search (single-piece batches on one shard): no copy engine in the step.
// StageArgs: up to three host-mapped ranges (in 16-byte units) the next wave launch -- or a k_stage launch when no view is
// stale -- copies to their device arrays before the scan starts: the plan's tasks, offs and ctl streams.
struct StageArgs { const uint4 *src[3]; uint4 *dst[3]; int n[3]; };
// PubArgs: the scan kernel's last block publishes the counts (flag == nullptr: nothing to do), see k_spr_scan
struct PubArgs { int32_t *host_counts; uint32_t *host_wc; uint32_t *wcount; uint32_t *flag; uint32_t *done; int nout, nwc; uint32_t epoch; };

struct PendingWave { int list_off, total, wc_off; bool incremental; };

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int shard_rank = 0, shard_count = 1;  // PATTERN shards: this context holds a word slice of every plane
    // REPLICATE shards (mpgpu_set_replicate_shards; pattern-unsharded contexts only): every context holds the whole alignment and
    // scores all candidates, but only columns [rep_lo, rep_lo + Buser) of the replicate-weight matrix: the -bb contraction, the
    // dominant cost, divides by the number of GPUs, and only the rows of calls that can change a replicate are exchanged
    int rep_rank = 0, rep_count = 1;
    int xrank() const { return shard_count > 1 ? shard_rank : rep_rank; }     // rank / size of the exchange group
    int xcount() const { return shard_count > 1 ? shard_count : rep_count; }
    mpgpu_allreduce_fn allreduce = nullptr; void *allreduce_user = nullptr;
    PeerExchange peer;
    bool exchange_off = false;            // option "exchange" 0: shard_sum is skipped (timing the kernels alone; results stay partial)
    bool reduces() const { return shard_count == 1 || allreduce != nullptr || peer.ready; }   // results are complete on this shard
    int64_t launches = 0;
    // the last SPR search on this context (mpgpu_search_info)
    uint32_t search_start_score = 0; int64_t search_moves = 0, search_batches = 0;

    // alignment
    int n = 0, P = 0, datatype = 0, S = 0, sort_alignment = 1;
    std::vector<int32_t> weights;
    std::vector<uint8_t> informative;
    std::vector<uint32_t> present;        // [P] unambiguous states present among the tips (findMstScore, parstree.cpp:606)
    int n_inf = 0;
    int64_t n_sites = 0;
    int ref_words = 0;                    // reference parsimonyLength (padded to 8)
    int glob_words = 0;                   // padded words per plane over all shards
    int Wl = 0;                           // words per plane of this shard
    int64_t w0 = 0;                       // first global word of this shard
    uint8_t *d_codes = nullptr;           // [n][P]
    int64_t *d_site_start = nullptr;      // [n_inf+1] first expanded site of informative pattern k
    int32_t *d_inf_ptn = nullptr;         // [n_inf] pattern index of informative pattern k
    PinnedArray<int64_t> site_pin;        // upload staging of d_site_start (re-weighting: no allocation, no wait)

    // views
    uint32_t *d_views = nullptr;          // [4n-6][S][Wl]
    size_t view_stride = 0;               // S*Wl
    size_t views_alloc = 0;               // words allocated behind d_views
    uint32_t *d_vcount = nullptr;         // [4n-6] mismatch count of each view (this shard)
    std::vector<uint32_t> vcount;         // host copy (all-reduced when sharded)
    std::vector<uint32_t> vlen;           // subtree length of each view
    bool tree_set = false, lens_valid = false;
    bool views_stale = false;             // the planes changed under a resident tree: views are recomputed on first use (need_tree)
    uint32_t start_edge_mis = 0; bool start_edge_valid = false;   // mismatch count across the edge at tip 1, read back with the view counts
    PinnedArray<uint32_t> vcount_pin;     // read-back staging of compute_views
    double move_gap = 8.0;                // SPR search: recent distance (node visits) between applied moves, kept across searches
    HostTree tree;
    std::vector<Triple> sched;            // every inner view once, children before parents; pad = dependency level (1-based)
    int sched_levels = 0;
    std::vector<int32_t> sc_level, sc_stack, sc_dl, sc_slot, sc_fill;   // scratch of build_schedule / update_views (no allocation per move)
    std::vector<Triple> sc_stale;
    Triple *d_triples = nullptr; size_t triples_cap = 0;
    // incremental update after a move (update_views): children of every view at its last compute
    std::vector<int2> view_kids; bool kids_valid = false;
    PinnedArray<Triple> wave_pin; PinnedArray<uint32_t> wcount_pin;
    int wave_pending = 0;                 // views whose recomputation is in flight (settle_views after the next synchronize)
    std::vector<PendingWave> wave_lists;  // the lists in flight, in launch order
    size_t wave_used = 0, wc_used = 0;    // staging consumed by them (Triples of wave_pin / d_wave, counters of wcount_pin / d_wcount)
    bool wave_fetched = false;            // their counts are on the way to wcount_pin (fetch_wave_counts / k_publish)
    bool wcount_zeroed = false;           // d_wcount is all zero outside the lists in flight
    // lazy views of the SPR search: after a move only the views the next scan batch reads are recomputed (ensure_views)
    std::vector<uint8_t> vstale;          // [4n-6] the view's content is out of date
    int n_stale = 0;
    int64_t lazy_lists = 0, lazy_views = 0, lazy_levels = 0;   // statistics of the lazy lists (MPGPU_PROFILE)
    std::vector<int32_t> sc_refs;         // scratch: ring slots handed to ensure_views
    bool dl_dirty = false;                // sc_dl holds levels of an abandoned list
    Triple *d_wave = nullptr; size_t wave_cap = 0;
    uint32_t *d_wcount = nullptr; size_t wcount_cap = 0;
    uint32_t *d_scalar = nullptr;         // small scratch for scalar results

    // scan
    ScanPlan plan;
    std::vector<ScanRef> ref_table; uint32_t ref_vstride = 0; bool ref_valid = false;   // the planner's slot table, kept across the batches of a search
    ScanOffs *d_offs = nullptr; size_t offs_cap = 0;
    ScanCtl *d_ctl = nullptr; size_t ctl_cap = 0;
    ScanTask *d_tasks = nullptr; size_t tasks_cap = 0;
    int32_t *d_counts = nullptr; size_t counts_cap = 0;
    int32_t *h_counts = nullptr; size_t h_counts_cap = 0;   // pinned read-back buffer (bytes)
    StageArgs stage_req = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}, {0, 0, 0}}; bool stage_pending = false;
    bool pub_request = false, pub_inflight = false;       // the next launch_scan publishes / a fused publish is in flight (finish_scan spins)
    int pub_nout = 0;
    uint32_t *d_done = nullptr;           // block ticket of the fused publish
    size_t counts_dirty = 0;              // leading ints of d_counts that may be non-zero (zero_counts clears them before a scan)
    uint32_t *h_flag = nullptr;           // mapped page-locked word k_publish writes last: the host spins on it instead of a stream synchronize
    uint32_t flag_epoch = 0;

    // pattern scores
    uint32_t *d_bitcnt = nullptr; size_t bitcnt_cap = 0;   // bit-sliced per-site counters
    int32_t *d_pairs = nullptr; size_t pairs_cap = 0;
    PinnedArray<int32_t> pairs_pin; int pairs_flip = 0;    // staging of the (child view, child view) pairs of compute_site_counters
    std::vector<int32_t> sc_order;
    uint16_t *d_ptn = nullptr; size_t ptn_cap = 0;
    int64_t *d_ptn_site = nullptr; size_t ptn_site_cap = 0;
    bool ptn_site_valid = false;
    bool ptn_identity = false;            // pattern i sits at expanded site i (all reported weights are 1, single shard)

    // replicate scoring
    Reps reps;
    // weighted parsimony
    Sankoff sk;
    // aliases the scan kernel's ROWS mode reads (owned by reps)
    int32_t *d_row_of = nullptr, *d_row_tasks = nullptr;
    uint32_t *d_rows_site = nullptr;
};

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);
#define MPGPU_STR2(x) #x
#define MPGPU_STR(x) MPGPU_STR2(x)
#define MPGPU_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return ::mpgpu::cuda_fail(e__, #expr " (" __FILE__ ":" MPGPU_STR(__LINE__) ")"); } while (0)

template <typename T>
static inline int ensure(T *&ptr, size_t &cap, size_t need)
{
    if (need <= cap && ptr) return 0;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; cap = 0;
    size_t want = need + need / 4 + 16;
    MPGPU_CUDA(cudaMalloc((void **)&ptr, want * sizeof(T)));
    cap = want;
    return 0;
}

// ---- shared host helpers (mpgpu_api.cu) -----------------------------------------------------
void peer_free(Ctx *c);
int peer_allreduce(Ctx *c, void *dev_i32, int64_t count);
int shard_sum(Ctx *c, void *dev_i32, int64_t count);
int group_sum(Ctx *c, void *dev_i32, int64_t count);   // the same exchange over whatever group the context belongs to (pattern or replicate shards)   // in-place int32 all-reduce over the shards (no-op for one shard)
int compute_views(Ctx *c, bool want_start_edge = false);
int set_tree_impl(Ctx *c, const int32_t *back_node, const int32_t *back_slot, bool want_start_edge);
int update_views(Ctx *c, bool defer = false);   // after apply_spr_move on c->tree: recompute only the stale views, one launch
int submit_stale(Ctx *c, std::vector<Triple> &stale, int nlevels, bool defer, bool incremental);
int fetch_wave_counts(Ctx *c);                  // before the stream synchronize that precedes settle_views
void settle_views(Ctx *c, bool lengths);        // after that synchronize: land the counts of the deferred view updates
void mark_stale_nodes(Ctx *c, const int *nodes, int k);   // lazy scheme: the adjacency of these nodes changed
int ensure_views(Ctx *c, const int32_t *refs, int count, bool defer);   // recompute the stale ones among the views behind these ring slots (and what they depend on)
int ensure_all_views(Ctx *c);
int zero_counts(Ctx *c, size_t n);              // d_counts[0, n) = 0 before a scan
int launch_publish(Ctx *c, int nout);           // counts (and the wave counts in flight) -> mapped host memory, device counters back to zero, flag
void compute_lengths(Ctx *c);
int need_tree(Ctx *c, bool lens);
int run_scan(Ctx *c);
int scan_batch_pipelined(Ctx *c, const int32_t *order, int first, int count, int mintrav, int maxtrav);
int upload_plan(Ctx *c);
int finish_scan(Ctx *c, int32_t *visit_begin, uint32_t *mp, int32_t *cand_ref, int32_t *cand_prune, int capacity);
int compute_site_counters(Ctx *c, int nbits);       // bit-sliced per-site counters of the current tree -> d_bitcnt
int ensure_ptn_site(Ctx *c);                        // first expanded site of every reported pattern -> d_ptn_site
void free_reps(Ctx *c);

// ---- Sankoff (-cost) path (sankoff.cu) ------------------------------------------------------------
void sk_free(Ctx *c);
int sk_build(Ctx *c);
int sk_compute_levels(Ctx *c, const std::vector<int32_t> &start, int nl);
int sk_update_stale(Ctx *c, std::vector<Triple> &stale, int nlevels);
int sk_junctions(Ctx *c, const int4 *list, int count, uint16_t *ptn, bool tip_rooted = false);
int sk_tree_score(Ctx *c, int start_ref, uint32_t *score);
int sk_pattern_parsimony(Ctx *c, uint16_t *ptn_pars, int count, int32_t *sum);
int sk_raw_view(Ctx *c, int ref, uint16_t *out);
int sk_run_scan(Ctx *c);
int sk_reps_rows_capacity(Ctx *c);
int sk_reps_chunk(Ctx *c, const int32_t *h_row_of, int nsel, const int32_t *h_call_row, int ncalls, bool use_thr,
                  const int32_t *h_visits = nullptr, int nvis = 0);
int sk_visit_edges(Ctx *c, const int32_t *visits, int nv, uint32_t *d_rows_out);   // the current tree evaluated at the edges of planned visits -> h_tot (and rows)
int sk_finish_scan(Ctx *c, int32_t *visit_begin, uint32_t *mp, int32_t *cand_ref, int32_t *cand_prune, int capacity);

// ---- kernel launchers (fitch_kernels.cu) -------------------------------------------------
int launch_compress(Ctx *c);
int launch_level(Ctx *c, const Triple *d_triples, int ntriples);
int launch_wave(Ctx *c, const Triple *d_list, int nlevels, int hdr, int total, uint32_t *d_wcount);
int launch_stage(Ctx *c);                        // the pending StageArgs alone (no wave to ride on)
int wave_slot_cap(int S);
size_t wave_smem_bytes(int S, int entries);      // dynamic shared memory of k_fitch_wave for a list of `entries` Triples
int launch_edge_mismatch(Ctx *c, int vidA, int vidB, uint32_t *d_out);
int launch_scan(Ctx *c, int task0, int ntasks, int nslots);
int launch_scan_rows(Ctx *c, int ntasks, int nslots);
int launch_tip_insert(Ctx *c, const int4 *d_edges, int nedges, int32_t *d_out);
int launch_site_counters(Ctx *c, int npairs, int nbits);
int launch_gather_patterns(Ctx *c, int nbits, int count);
const uint32_t *state_mask_table(int datatype, int *ncodes, int *undetermined);

// ---- kernel launchers (reps_kernels.cu) ----------------------------------------------------
int launch_transpose_boot(Ctx *c, const uint16_t *d_boot16, int stride, uint8_t *d_heavy);
int launch_build_w8(Ctx *c, const uint8_t *d_is_exc);
int launch_seg_check(Ctx *c, int32_t *d_segmax);
int launch_seg_cmax(Ctx *c, int32_t *d_cmax);
int launch_edge_rows(Ctx *c, const int4 *d_edges, int nedges, uint32_t *d_rows);
int launch_gather_rows(Ctx *c, const uint32_t *d_src, uint32_t *d_dst, int nrows);
int launch_reps_exc(Ctx *c, const uint32_t *a_base, int a_pitch, int a_word0, int x_row0, int nrows);
int make_w8_tensor_map(Ctx *c);
int launch_reps_tc(Ctx *c, const uint32_t *a_base, int a_pitch, int a_word0, int x_row0, int nrows);
int launch_reps_tc_bytes(Ctx *c, const uint8_t *rows8, int pitch_bytes, int nrows, int32_t *X, int x_pitch);
int measure_int8_peak(Ctx *c, int iters, double *tops);
int launch_reps_tree_row(Ctx *c, int plane_row0, int nbits, int t_row);
int launch_reps_combine(Ctx *c, int t_row, const int2 *d_calls, int ncalls, int32_t *d_res, const int32_t *d_thr, int32_t *d_call_hit);
int launch_gather_res_rows(Ctx *c, const int32_t *d_res, const int32_t *d_list, int nlist, int32_t *d_out);
int launch_keep_rows(Ctx *c, const uint32_t *d_src, int pitch, const int2 *d_calls, const int32_t *d_list, int nl, uint32_t *d_dst);
int launch_prefix_max(Ctx *c, const uint16_t *d_tree_ptn, const uint32_t *d_rows2, const int32_t *d_remain, const int32_t *d_blist, int nl,
                      int32_t *d_segsum, int32_t *d_out);

// ---- host SPR logic (spr_host.cpp) ---------------------------------------------------------
void visit_order(const HostTree &t, std::vector<int32_t> &order);
class ScanPlanner {                      // incremental form of build_scan_plan (pieces of consecutive visits)
public:
    ScanPlanner();
    ~ScanPlanner();
    int begin(const HostTree &t, const int32_t *order, int first, int count,
              int mintrav, int maxtrav, uint32_t vstride, ScanPlan &plan, bool host_only = false,
              const uint8_t *vstale = nullptr, int split_depth = 0, ScanRef *table = nullptr);
    void add(int v0, int v1);
    // the same result as add(v0, v1) -- byte for byte -- with the enumeration spread over `nthreads` host threads (the caller's
    // included): every thread enumerates a range of visits into buffers of its own, then copies its share into place.  Falls
    // back to add() for small ranges, lazy plans and plans that are split into sub-tasks.
    void add_parallel(int v0, int v1, int nthreads);
    void finish();
    void split();
private:
    struct Impl;
    Impl *impl;
    ScanPlanner(const ScanPlanner &);
    ScanPlanner &operator=(const ScanPlanner &);
};
int plan_threads();                     // host threads ScanPlanner::add_parallel may use (MPGPU_PLAN_THREADS; 1 = the caller alone)
int build_scan_plan(const HostTree &t, const int32_t *order,
                    int first, int count, int mintrav, int maxtrav, uint32_t vstride_vec, ScanPlan &plan);
void apply_spr_move(HostTree &t, int remove_ref, int insert_ref);
void scan_ref_build(const HostTree &t, uint32_t vstride, std::vector<ScanRef> &tab);
void scan_ref_fill(const HostTree &t, uint32_t vstride, ScanRef *tab, int node);

}  // namespace mpgpu

struct mpgpu_ctx : public mpgpu::Ctx {};
