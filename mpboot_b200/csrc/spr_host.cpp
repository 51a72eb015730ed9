// Host side of the SPR scan: the reference's tree walking, restated on ring tables so that the
// device can score whole prune neighbourhoods in one launch while the host keeps the exact
// visit order, candidate order and RNG discipline of the reference.
//
//   visit_order        nodeRectifierPars / reorderNodes      (sprparsimony.cpp:2046-2101)
//   build_scan_plan    rearrangeParsimony + addTraverseParsimony enumeration
//                                                           (sprparsimony.cpp:2208-2218, 2259-2376)
//   apply_spr_move     restoreTreeRearrangeParsimony         (sprparsimony.cpp:2379-2384, 2191-2205)
#include "mpgpu_internal.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <unistd.h>

namespace mpgpu {

// tr->nodep[1..2n-2] after nodeRectifierPars: tips in number order, then the inner nodes in
// pre-order from tr->start->back (tr->start = tip 1), each as the ring slot that was entered.
void visit_order(const HostTree &t, std::vector<int32_t> &order)
{
    const int n = t.n;
    order.assign(2 * n - 1, 0);
    for (int i = 1; i <= n; i++) order[i] = 3 * i;
    int count = 0;
    std::vector<int> stack;
    stack.push_back(t.back(3 * 1));
    while (!stack.empty()) {
        int p = stack.back(); stack.pop_back();
        if (t.is_tip(p)) continue;
        order[n + 1 + count++] = p;
        // recursion order: p->next->back first, then p->next->next->back
        stack.push_back(t.back(t.next(t.next(p))));
        stack.push_back(t.back(t.next(p)));
    }
}

// the records of all ring slots of `node` (see ScanRef)
void scan_ref_fill(const HostTree &t, uint32_t vstride, ScanRef *tab, int node)
{
    const int n = t.n;
    const int ns = node <= n ? 1 : 3;
    for (int sl = 0; sl < ns; sl++) {
        const int r = 3 * node + sl;
        const int v = t.vid(r);
        ScanRef &e = tab[r];
        e.voff = (int32_t)((uint32_t)v * vstride);
        e.tip = v << 2 | (node <= n ? 1 : 0);
        e.c1 = 0; e.c2 = 0;
        if (node > n) {
            const int a = 3 * node + (sl + 1) % 3, b = 3 * node + (sl + 2) % 3;
            e.c1 = t.back(a); e.c2 = t.back(b);
        }
    }
}

void scan_ref_build(const HostTree &t, uint32_t vstride, std::vector<ScanRef> &tab)
{
    const int n = t.n;
    tab.assign((size_t)3 * (2 * n - 1), ScanRef{0, 0, 0, 1});
    for (int node = 1; node <= 2 * n - 2; node++) scan_ref_fill(t, vstride, tab.data(), node);
}

namespace {

// Flat per-ref tables of the tree for the enumeration (built once per plan): the two refs
// behind an inner ref, tip flags and view offsets, so that the hot recursion touches no
// division and no ring arithmetic.
struct Builder {
    const int n;
    ScanPlan &plan;
    // per ref, one record (one cache line touch per visited node): back(next(ref)), back(next(next(ref))), view offset and
    // tip = view id << 2 | noted << 1 | is-tip (ScanRef, mpgpu_internal.h).  The table is the caller's (the SPR search keeps
    // one per context and patches the five nodes a move touches: building it is O(n), most of the host cost of a small batch)
    // or, without one, built here.
    typedef ScanRef Ref;
    std::vector<Ref> own_;
    Ref *ref_ = nullptr;
    // views of the same data for the callers that index by field
    struct Field1 { Ref *const *r; int32_t operator[](int i) const { return (*r)[i].c1; } } c1{&ref_};
    struct Field2 { Ref *const *r; int32_t operator[](int i) const { return (*r)[i].c2; } } c2{&ref_};
    struct FieldT { Ref *const *r; bool operator[](int i) const { return ((*r)[i].tip & 1) != 0; } } tip{&ref_};
    const uint8_t *vstale = nullptr;     // some views are stale: every view the plan reads is checked (need)
    bool lazy = false;
    std::vector<int32_t> noted;          // refs whose "noted" bit this plan set (cleared again by the destructor)
    int prune_ref = 0, task_index = 0, cand_base = 0;
    // raw cursors into the plan's arrays (sized up front)
    ScanOffs *offs = nullptr; ScanCtl *ctl = nullptr;
    int32_t *cand_ref = nullptr, *cand_prune = nullptr, *cand_task = nullptr;
    int nops = 0, ncand = 0, max_slot = 0;
    // op tree (recorded when the plan may be split into sub-tasks): the ops that expand an op's first / second child
    int32_t *kid_op = nullptr;           // [2 * op], -1 = that child is not expanded

    Builder(const HostTree &t, ScanPlan &pp, uint32_t vstride, const uint8_t *vs, ScanRef *table)
        : n(t.n), plan(pp), vstale(vs), lazy(vs != nullptr)
    {
        if (table) ref_ = table;
        else { scan_ref_build(t, vstride, own_); ref_ = own_.data(); }
    }
    ~Builder() { for (int32_t r : noted) ref_[r].tip &= ~2; }
    int32_t voff(int ref) const { return ref_[ref].voff; }
    // the plan reads the view behind `ref`: noted when it is stale (once per plan)
    void need(int ref)
    {
        int32_t &f = ref_[ref].tip;
        if (!(f & 2) && vstale[f >> 2]) { f |= 2; noted.push_back(ref); plan.need_refs.push_back(ref); }
    }

    // One expand op for the node whose children (seen from it) are a and b; src = where its up-view comes from.
    // The op's control words are built in registers while the children are walked and stored once.
    int expand(int a, int b, uint32_t src, int mintrav, int maxtrav, int depth)
    {
        const int me = nops++;
        offs[me].c1 = ref_[a].voff; offs[me].c2 = ref_[b].voff;
        if (lazy) { need(a); need(b); }
        uint32_t outs = 0xFFFFFFFFu, meta = src | 0xFF00u | 0xFF0000u;
        const int k1 = child(a, 0, mintrav, maxtrav, depth, outs, meta);
        const int k2 = child(b, 1, mintrav, maxtrav, depth, outs, meta);
        ctl[me].outs = outs; ctl[me].meta = meta;
        if (kid_op) { kid_op[2 * me] = k1; kid_op[2 * me + 1] = k2; }
        return me;
    }

    // addTraverseParsimony(tr, pr, p, q, mintrav, maxtrav, doAll = FALSE) for q = x, the `which`-th child of the op being
    // built (outs / meta); depth is x's distance from the removed node (1-based).
    int child(int x, int which, int mintrav, int maxtrav, int depth, uint32_t &outs, uint32_t &meta)
    {
        if (--mintrav <= 0) {                                   // testInsertParsimony(p, x)
            const int idx = ncand++;
            cand_ref[idx] = x;                                  // cand_prune / cand_task: constant per task, filled by end_task()
            const uint32_t rel = (uint32_t)(idx - cand_base);
            outs = which == 0 ? ((outs & 0xFFFF0000u) | rel) : ((outs & 0x0000FFFFu) | (rel << 16));
        }
        const Ref &rx = ref_[x];
        if (!(rx.tip & 1) && (--maxtrav > 0)) {
            // U_x goes here.  The first child is expanded by the very next op, which reads its up-view before it writes
            // anything, so first children only need two slots, alternating with the depth (never the slot the op itself
            // reads); a second child waits for the first child's whole subtree and gets the slot of its depth.
            const int slot = which == 0 ? (depth & 1) : 1 + depth;
            meta = which == 0 ? ((meta & ~0xFF00u) | ((uint32_t)slot << 8)) : ((meta & ~0xFF0000u) | ((uint32_t)slot << 16));
            if (slot + 1 > max_slot) max_slot = slot + 1;
            return expand(rx.c1, rx.c2, (uint32_t)slot, mintrav, maxtrav, depth + 1);
        }
        return -1;
    }

    // candidates [cand_base, ncand) belong to the task that just ended
    void end_task()
    {
        std::fill(cand_prune + cand_base, cand_prune + ncand, prune_ref);
        std::fill(cand_task + cand_base, cand_task + ncand, task_index);
    }

    // the two addTraverseParsimony calls made for one inner neighbour `nb` of the removed node:
    // candidates are the branches to nb's children; the far side of nb is the task's D2 (when nb
    // is the D1 neighbour, src code 0xFF) or D1 (src code 0xFE).
    int expand_top(int nb, uint32_t src_code, int mintrav, int maxtrav)
    {
        return expand(ref_[nb].c1, ref_[nb].c2, src_code, mintrav, maxtrav, 1);
    }

    // ---- sub-tasks (latency path) ----------------------------------------------------------------------------------
    // A small batch is latency-bound: one warp walks a task's ~100 ops one after the other while most of the device
    // idles.  split() re-emits a task as independent sub-tasks, one per op at depth `sdepth` of the op tree (and per
    // childless op above it): the ops on the path from the task's top-level op down to that op are replayed without their
    // scores (an op above the split depth is scored by the first sub-task that passes through it) and with only the
    // up-view of the followed child, then the op's whole subtree follows verbatim.  The stack slots of the original
    // program stay valid because a sub-task is a prefix-closed slice of it.
    struct PathStep { int op, which; };
    std::vector<PathStep> path;
    std::vector<uint8_t> scored;         // [op] an internal op above the split depth whose scores were already given to a sub-task

    int subtree_end(int op) const        // ops are emitted in DFS order: [op, subtree_end) is the op's subtree
    {
        int last = op;
        for (;;) {
            const int k2 = kid_op[2 * last + 1], k1 = kid_op[2 * last];
            if (k2 >= 0) last = k2; else if (k1 >= 0) last = k1; else break;
        }
        return last + 1;
    }
    void emit_sub(const ScanTask &lt, int head, bool &first)
    {
        ScanTask st = lt;
        st.base_out = first ? lt.base_out : -1;
        first = false;
        st.op_begin = nops;
        for (const PathStep &ps : path) {
            const int me = nops++;
            offs[me] = offs[ps.op];
            uint32_t outs = 0xFFFFFFFFu, meta = ctl[ps.op].meta;
            if (!scored[ps.op]) { outs = ctl[ps.op].outs; scored[ps.op] = 1; }
            meta = ps.which == 0 ? (meta | 0xFF0000u) : (meta | 0xFF00u);       // only the followed child's up-view
            ctl[me].outs = outs; ctl[me].meta = meta;
        }
        const int e = subtree_end(head);
        memcpy(offs + nops, offs + head, (size_t)(e - head) * sizeof(ScanOffs));
        memcpy(ctl + nops, ctl + head, (size_t)(e - head) * sizeof(ScanCtl));
        nops += e - head;
        st.op_end = nops;
        plan.sub_tasks.push_back(st);
    }
    void split_op(const ScanTask &lt, int op, int depth, int sdepth, bool &first)
    {
        const int k1 = kid_op[2 * op], k2 = kid_op[2 * op + 1];
        if (depth >= sdepth || (k1 < 0 && k2 < 0)) { emit_sub(lt, op, first); return; }
        if (k1 >= 0) { path.push_back(PathStep{op, 0}); split_op(lt, k1, depth + 1, sdepth, first); path.pop_back(); }
        if (k2 >= 0) { path.push_back(PathStep{op, 1}); split_op(lt, k2, depth + 1, sdepth, first); path.pop_back(); }
    }
};

}  // namespace

// Enumerates what rearrangeParsimony(tr, pr, tr->nodep[i], mintrav, maxtrav, doAll=FALSE) tests
// for i in [first, first+count).  The plan can be built in pieces (scan_plan_begin, then
// scan_plan_add for consecutive visit ranges) so that the device can start on the first piece
// while the host enumerates the next.  Output layout of the device count vector: slots
// [0, task_cap) hold the joined-edge count of each task, candidate j sits at task_cap + j.
struct ScanPlanner::Impl {
    Builder b;
    const HostTree &t;
    const int32_t *order;
    int first, mintrav, maxtrav;
    int split_depth = 0;
    uint32_t vstride;
    Impl(const HostTree &tt, ScanPlan &plan, uint32_t vs, const int32_t *ord,
         int f, int mi, int ma, const uint8_t *vstale, ScanRef *table) : b(tt, plan, vs, vstale, table), t(tt), order(ord), first(f), mintrav(mi), maxtrav(ma), vstride(vs) {}
};

// ---- host threads for the enumeration of large plans ----------------------------------------------------------------------
// A whole sweep (C2: 398 visits, 14 476 candidates) takes one thread ~105 us to enumerate -- as long as the device needs to
// score it -- and sits in front of the last piece's launch on the e2e path.  The visits are independent, so a few detached
// workers enumerate ranges of them side by side.  They spin for a short while after a job (a sweep loop keeps them hot) and
// sleep on a condition variable otherwise; the pool is created on first use, per process (a forked child gets its own), and
// never torn down (no join at exit: the threads own nothing).
namespace {
inline void cpu_relax()
{
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
}
// A job is a number of independent items; every thread (the caller included) draws items from a shared counter until none is
// left, so the caller never waits for a worker that is asleep or was descheduled -- it only waits for items already drawn.
struct PlanJob {
    std::atomic<int> next{0}, done{0};
    int total = 0;
    std::function<void(int, int)> fn;                    // (thread, item); only ever called while the caller is still inside run()
};
class PlanPool {
public:
    explicit PlanPool(int workers) : n_(workers)
    {
        for (int k = 0; k <= workers; k++) parts.emplace_back(new ScanPlan());
        for (int k = 1; k <= workers; k++) std::thread([this, k]() { loop(k); }).detach();
    }
    int workers() const { return n_; }
    void run(int total, const std::function<void(int, int)> &fn)
    {
        std::shared_ptr<PlanJob> job = std::make_shared<PlanJob>();
        job->total = total; job->fn = fn;
        { std::lock_guard<std::mutex> lk(m_); cur_ = job; gen_.fetch_add(1, std::memory_order_release); }
        if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
        drain(*job, 0);
        while (job->done.load(std::memory_order_acquire) != total) cpu_relax();
    }
    std::vector<std::unique_ptr<ScanPlan>> parts;        // per thread: the buffers its items are enumerated into (kept across calls)
    std::mutex busy;                                     // one plan at a time uses the pool (and its buffers); others enumerate alone
private:
    static void drain(PlanJob &job, int k)
    {
        for (;;) {
            const int i = job.next.fetch_add(1, std::memory_order_relaxed);
            if (i >= job.total) return;                  // (a late thread gets here long after run() returned: it touches the job only)
            job.fn(k, i);
            job.done.fetch_add(1, std::memory_order_release);
        }
    }
    void loop(int k)
    {
        unsigned seen = 0;
        for (;;) {
            const auto t0 = std::chrono::steady_clock::now();
            unsigned spins = 0;
            while (gen_.load(std::memory_order_acquire) == seen) {
                cpu_relax();
                if ((++spins & 255) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(500)) {
                    std::unique_lock<std::mutex> lk(m_);
                    sleepers_.fetch_add(1, std::memory_order_release);
                    cv_.wait(lk, [&]() { return gen_.load(std::memory_order_acquire) != seen; });
                    sleepers_.fetch_sub(1, std::memory_order_release);
                }
            }
            std::shared_ptr<PlanJob> job;
            { std::lock_guard<std::mutex> lk(m_); job = cur_; seen = gen_.load(std::memory_order_acquire); }
            if (job) drain(*job, k);
        }
    }
    const int n_;
    std::shared_ptr<PlanJob> cur_;
    std::atomic<unsigned> gen_{0};
    std::atomic<int> sleepers_{0};
    std::mutex m_;
    std::condition_variable cv_;
};
}  // namespace
// host threads for the enumeration of a large batch: MPGPU_PLAN_THREADS, else up to 4 when the machine has the cores to spare
int plan_threads()
{
    static const int n = []() {
        if (const char *e = getenv("MPGPU_PLAN_THREADS")) { const int v = atoi(e); return v < 1 ? 1 : (v > 8 ? 8 : v); }
        // one process per GPU: the processes of a node share its cores (torchrun exports LOCAL_WORLD_SIZE), and a spinning helper
        // must not take a core another rank's calling thread needs -- at most half the hardware threads are ever used
        unsigned hc = std::thread::hardware_concurrency();
        if (const char *w = getenv("LOCAL_WORLD_SIZE")) { const int lw = atoi(w); if (lw > 1) hc /= (unsigned)lw; }
        return hc >= 8 ? 4 : (hc >= 4 ? 2 : 1);
    }();
    return n;
}
namespace {
PlanPool *plan_pool(int workers)
{
    static std::mutex mk;
    static PlanPool *pools[8] = {nullptr};               // one per size (in practice one size per process), never torn down
    static pid_t owner = 0;
    std::lock_guard<std::mutex> lk(mk);
    if (owner != getpid()) { for (PlanPool *&p : pools) p = nullptr; owner = getpid(); }      // a forked child: the parent's helpers are not here
    if (workers < 1) workers = 1;
    if (workers > 7) workers = 7;
    if (!pools[workers]) pools[workers] = new PlanPool(workers);
    return pools[workers];
}
}  // namespace

ScanPlanner::ScanPlanner() : impl(nullptr) {}
ScanPlanner::~ScanPlanner() { delete impl; }

int ScanPlanner::begin(const HostTree &t, const int32_t *order, int first, int count,
                       int mintrav, int maxtrav_in, uint32_t vstride, ScanPlan &plan, bool host_only, const uint8_t *vstale,
                       int split_depth, ScanRef *table)
{
    delete impl; impl = nullptr;
    plan.tasks.clear(); plan.visit_begin.clear(); plan.visit_ref.clear(); plan.task_vids.clear(); plan.need_refs.clear();
    plan.sub_tasks.clear(); plan.task_tops.clear();
    plan.n_cand = 0; plan.n_ops = 0; plan.max_slot = 0;
    plan.task_cap = 2 * count;
    const int n = t.n;
    int maxtrav = maxtrav_in;
    if (maxtrav > n - 3) maxtrav = n - 3;                       // :2275 (tr->ntips == mxtips during the search)
    if (maxtrav > kMaxTrav) { set_error("maxtrav exceeds the scan kernel's stack depth"); return 1; }   // slots < 0xFE
    if ((uint64_t)(4 * n - 6) * vstride >= 0x7fffffffULL) { set_error("view array too large for 32-bit scan offsets"); return 1; }
    impl = new Impl(t, plan, vstride, order, first, mintrav, maxtrav, vstale, table);
    Builder &b = impl->b;
    // upper bounds: one side of a visit reaches at most 4 * (2^maxtrav - 1) branches, never more than the tree has
    const size_t per_side = std::min<size_t>((size_t)4 << std::max(maxtrav, 0), (size_t)2 * n);
    const size_t cap1 = (size_t)count * 2 * per_side + 16;
    // split plans: every op once more (the sub-tasks' copies) plus the replayed paths (up to 2^(d-1) heads per side, d - 1 ops each)
    impl->split_depth = host_only ? 0 : split_depth;
    const size_t cap = impl->split_depth > 0 ? 2 * cap1 + (size_t)count * 4 * ((size_t)split_depth << split_depth) : cap1;
    plan.item_cap = impl->split_depth > 0 ? (size_t)plan.task_cap * (1 + ((size_t)2 << split_depth)) : (size_t)plan.task_cap;
    if (impl->split_depth > 0) { if (plan.kid_op.size() < 2 * cap1) plan.kid_op.resize(2 * cap1); b.kid_op = plan.kid_op.data(); }
    if (host_only) {
        if (plan.offs_host.size() < cap + cap / 4) { plan.offs_host.resize(cap + cap / 4); plan.ctl_host.resize(cap + cap / 4); }   // (grow only: no refill)
    } else if (!plan.offs.reserve(cap + cap / 4) || !plan.ctl.reserve(cap + cap / 4) || !plan.tasks_pin.reserve(plan.item_cap + 16)) {
        set_error("page-locked host allocation for the scan plan failed"); return 1;
    }
    if (plan.cand_ref.size() < cap1) { plan.cand_ref.resize(cap1); plan.cand_prune.resize(cap1); plan.cand_task.resize(cap1); }
    b.offs = host_only ? plan.offs_host.data() : plan.offs.data();
    b.ctl = host_only ? plan.ctl_host.data() : plan.ctl.data();
    b.cand_ref = plan.cand_ref.data(); b.cand_prune = plan.cand_prune.data(); b.cand_task = plan.cand_task.data();
    plan.tasks.reserve((size_t)plan.task_cap);
    return 0;
}

// visits [v0, v1) relative to `first`; must be called with consecutive ranges
void ScanPlanner::add(int v0, int v1)
{
    Builder &b = impl->b;
    ScanPlan &plan = b.plan;
    const HostTree &t = impl->t;
    const int mintrav = impl->mintrav, maxtrav = impl->maxtrav;
    for (int v = v0; v < v1; v++) {
        plan.visit_begin.push_back(b.ncand);
        plan.visit_ref.push_back(impl->order[impl->first + v]);
        if (maxtrav < mintrav) continue;                        // :2280
        const int p = impl->order[impl->first + v];
        const int q = t.back(p);

        if (!b.tip[p]) {                                        // :2303
            const int p1 = b.c1[p], p2 = b.c2[p];
            if (!b.tip[p1] || !b.tip[p2]) {
                ScanTask task;
                task.s_vid = b.voff(q); task.d1 = b.voff(p1); task.d2 = b.voff(p2);
                task.op_begin = b.nops; task.base_out = (int)plan.tasks.size(); task.cand_base = plan.task_cap + b.ncand; task.pad = 0;
                b.cand_base = b.ncand;
                b.prune_ref = p; b.task_index = (int)plan.tasks.size();
                if (b.lazy) { b.need(q); b.need(p1); b.need(p2); b.need(p); }     // p itself: the -bb edge row of the task reads both sides of (p, q)
                int top1 = -1, top2 = -1;
                if (!b.tip[p1]) top1 = b.expand_top(p1, 0xFFu, mintrav, maxtrav);
                if (!b.tip[p2]) top2 = b.expand_top(p2, 0xFEu, mintrav, maxtrav);
                if (b.kid_op) { plan.task_tops.push_back(top1); plan.task_tops.push_back(top2); }
                b.end_task();
                task.op_end = b.nops;
                plan.tasks.push_back(task);
                plan.task_vids.push_back(t.vid(q)); plan.task_vids.push_back(t.vid(p1)); plan.task_vids.push_back(t.vid(p2));
            }
        }
        if (!b.tip[q] && maxtrav > 0) {                         // :2333
            const int q1 = b.c1[q], q2 = b.c2[q];
            const bool ok1 = !b.tip[q1] && (!b.tip[b.c1[q1]] || !b.tip[b.c2[q1]]);
            const bool ok2 = !b.tip[q2] && (!b.tip[b.c1[q2]] || !b.tip[b.c2[q2]]);
            if (ok1 || ok2) {
                const int mintrav2 = mintrav > 2 ? mintrav : 2;
                ScanTask task;
                task.s_vid = b.voff(p); task.d1 = b.voff(q1); task.d2 = b.voff(q2);
                task.op_begin = b.nops; task.base_out = (int)plan.tasks.size(); task.cand_base = plan.task_cap + b.ncand; task.pad = 0;
                b.cand_base = b.ncand;
                b.prune_ref = q; b.task_index = (int)plan.tasks.size();
                if (b.lazy) { b.need(p); b.need(q1); b.need(q2); b.need(q); }
                int top1 = -1, top2 = -1;
                if (!b.tip[q1]) top1 = b.expand_top(q1, 0xFFu, mintrav2, maxtrav);
                if (!b.tip[q2]) top2 = b.expand_top(q2, 0xFEu, mintrav2, maxtrav);
                if (b.kid_op) { plan.task_tops.push_back(top1); plan.task_tops.push_back(top2); }
                b.end_task();
                task.op_end = b.nops;
                plan.tasks.push_back(task);
                plan.task_vids.push_back(t.vid(p)); plan.task_vids.push_back(t.vid(q1)); plan.task_vids.push_back(t.vid(q2));
            }
        }
    }
    plan.n_cand = b.ncand; plan.n_ops = b.nops; plan.max_slot = b.max_slot;
}

void ScanPlanner::add_parallel(int v0, int v1, int nthreads)
{
    Builder &b = impl->b;
    ScanPlan &plan = b.plan;
    const int nv = v1 - v0;
    if (nthreads > 8) nthreads = 8;
    if (nthreads < 2 || nv < 16 * nthreads || b.lazy || b.kid_op || impl->maxtrav < impl->mintrav) { add(v0, v1); return; }
    PlanPool *pool = plan_pool(nthreads - 1);
    // contexts driven from different host threads (one per GPU) share the pool: whoever finds it taken enumerates on its own
    std::unique_lock<std::mutex> taken(pool->busy, std::try_to_lock);
    if (!taken.owns_lock()) { add(v0, v1); return; }
    const int W = pool->workers() + 1;
    const Impl *me = impl;
    // the range in chunks of 16 visits; a chunk is enumerated by whichever thread draws it, into that thread's own buffers
    // (local op, candidate and task numbers), and remembers where
    const int CH = 16, C = (nv + CH - 1) / CH;
    struct Chunk { int owner, ops, nops, cand, ncand, task, ntask, vis, nvis; };
    std::vector<Chunk> chunks((size_t)C);
    std::vector<std::unique_ptr<ScanPlanner>> sub((size_t)W);
    pool->run(C, [&](int k, int c) {
        ScanPlan &pp = *pool->parts[k];
        if (!sub[k]) {
            sub[k].reset(new ScanPlanner());
            if (sub[k]->begin(me->t, me->order, me->first + v0, nv, me->mintrav, me->maxtrav, me->vstride, pp, true, nullptr, 0, me->b.ref_)) return;
        }
        Chunk &ck = chunks[c];
        ck.owner = k; ck.ops = pp.n_ops; ck.cand = pp.n_cand; ck.task = (int)pp.tasks.size(); ck.vis = (int)pp.visit_ref.size();
        sub[k]->add(c * CH, std::min(nv, c * CH + CH));
        ck.nops = pp.n_ops - ck.ops; ck.ncand = pp.n_cand - ck.cand; ck.ntask = (int)pp.tasks.size() - ck.task; ck.nvis = (int)pp.visit_ref.size() - ck.vis;
    });
    std::vector<int> ops0((size_t)C + 1), cand0((size_t)C + 1), task0((size_t)C + 1);
    ops0[0] = b.nops; cand0[0] = b.ncand; task0[0] = (int)plan.tasks.size();
    for (int c = 0; c < C; c++) { ops0[c + 1] = ops0[c] + chunks[c].nops; cand0[c + 1] = cand0[c] + chunks[c].ncand; task0[c + 1] = task0[c] + chunks[c].ntask; }
    // every chunk's streams move to their place in the plan, in visit order (control words are relative to their task: unchanged)
    pool->run(C, [&](int, int c) {
        const Chunk &ck = chunks[c];
        const ScanPlan &pp = *pool->parts[ck.owner];
        memcpy(b.offs + ops0[c], pp.offs_host.data() + ck.ops, (size_t)ck.nops * sizeof(ScanOffs));
        memcpy(b.ctl + ops0[c], pp.ctl_host.data() + ck.ops, (size_t)ck.nops * sizeof(ScanCtl));
        memcpy(b.cand_ref + cand0[c], pp.cand_ref.data() + ck.cand, (size_t)ck.ncand * sizeof(int32_t));
        memcpy(b.cand_prune + cand0[c], pp.cand_prune.data() + ck.cand, (size_t)ck.ncand * sizeof(int32_t));
        const int32_t *ct = pp.cand_task.data() + ck.cand;
        int32_t *dst = b.cand_task + cand0[c];
        const int shift = task0[c] - ck.task;
        for (int j = 0; j < ck.ncand; j++) dst[j] = ct[j] + shift;
    });
    for (int c = 0; c < C; c++) {
        const Chunk &ck = chunks[c];
        const ScanPlan &pp = *pool->parts[ck.owner];
        for (int i = 0; i < ck.ntask; i++) {
            ScanTask tk = pp.tasks[(size_t)ck.task + i];
            tk.op_begin += ops0[c] - ck.ops; tk.op_end += ops0[c] - ck.ops;
            tk.base_out = task0[c] + i;
            tk.cand_base = plan.task_cap + cand0[c] + (tk.cand_base - pp.task_cap - ck.cand);
            plan.tasks.push_back(tk);
        }
        for (int i = 0; i < ck.nvis; i++) {
            plan.visit_begin.push_back(pp.visit_begin[(size_t)ck.vis + i] - ck.cand + cand0[c]);
            plan.visit_ref.push_back(pp.visit_ref[(size_t)ck.vis + i]);
        }
        plan.task_vids.insert(plan.task_vids.end(), pp.task_vids.begin() + 3 * (size_t)ck.task, pp.task_vids.begin() + 3 * (size_t)(ck.task + ck.ntask));
    }
    for (int k = 0; k < W; k++) if (sub[k] && pool->parts[k]->max_slot > b.max_slot) b.max_slot = pool->parts[k]->max_slot;
    b.nops = ops0[C]; b.ncand = cand0[C];
    plan.n_cand = b.ncand; plan.n_ops = b.nops; plan.max_slot = b.max_slot;
}

void ScanPlanner::finish()
{
    ScanPlan &plan = impl->b.plan;
    plan.visit_begin.push_back(plan.n_cand);
}

// after the last add() of a plan begun with split_depth > 0: plan.sub_tasks = the tasks cut into sub-tasks at that depth of
// their op trees (ops appended behind the plan's own; n_ops grows, the tasks and their op ranges stay as they are)
void ScanPlanner::split()
{
    Builder &b = impl->b;
    ScanPlan &plan = b.plan;
    plan.sub_tasks.clear();
    if (!b.kid_op || impl->split_depth <= 0) return;
    b.scored.assign((size_t)b.nops, 0);
    const int nlogical_ops = b.nops;
    (void)nlogical_ops;
    for (size_t k = 0; k < plan.tasks.size(); k++) {
        bool first = true;
        for (int side = 0; side < 2; side++) {
            const int top = plan.task_tops[2 * k + side];
            if (top >= 0) { b.path.clear(); b.split_op(plan.tasks[k], top, 1, impl->split_depth, first); }
        }
    }
    plan.n_ops = b.nops;
}

int build_scan_plan(const HostTree &t, const int32_t *order,
                    int first, int count, int mintrav, int maxtrav, uint32_t vstride, ScanPlan &plan)
{
    ScanPlanner pl;
    if (int rc = pl.begin(t, order, first, count, mintrav, maxtrav, vstride, plan)) return rc;
    pl.add(0, count);
    pl.finish();
    return 0;
}

// removeNodeParsimony(removeNode) followed by restoreTreeParsimony(removeNode, insertNode)
void apply_spr_move(HostTree &t, int p, int q)
{
    const int a = t.back(t.next(p)), b = t.back(t.next(t.next(p)));
    t.hookup(a, b);
    const int r = t.back(q);
    t.hookup(t.next(p), q);
    t.hookup(t.next(t.next(p)), r);
}

}  // namespace mpgpu

// ---- host-only entry points (no device, no CUDA call): the tree-walking half of the path for hosts that keep their own
// search loop and for the CPU tests of the host logic ---------------------------------------------------------------
using namespace mpgpu;

static int host_tree_from(int ntaxa, const int32_t *back_node, const int32_t *back_slot, HostTree &t)
{
    if (ntaxa < 4 || !back_node || !back_slot) { set_error("bad tree argument"); return 1; }
    const int len = 3 * (2 * ntaxa - 1);
    t.n = ntaxa; t.bn.assign(back_node, back_node + len); t.bs.assign(back_slot, back_slot + len);
    return 0;
}

extern "C" {

int mpgpu_host_visit_order(int ntaxa, const int32_t *back_node, const int32_t *back_slot, int32_t *order)
{
    HostTree t;
    if (int rc = host_tree_from(ntaxa, back_node, back_slot, t)) return rc;
    if (!order) { set_error("null argument"); return 1; }
    std::vector<int32_t> o;
    visit_order(t, o);
    memcpy(order, o.data(), o.size() * sizeof(int32_t));
    return 0;
}

int mpgpu_host_enumerate(int ntaxa, const int32_t *back_node, const int32_t *back_slot, const int32_t *order, int first, int count,
                         int mintrav, int maxtrav, int32_t *visit_begin, int32_t *cand_ref, int32_t *cand_prune, int capacity, int *n_cand)
{
    HostTree t;
    if (int rc = host_tree_from(ntaxa, back_node, back_slot, t)) return rc;
    if (!order || !visit_begin || first < 1 || count < 0 || first + count > 2 * ntaxa - 1) { set_error("bad visit range"); return 1; }
    static thread_local ScanPlan plan;                     // its arrays are sized to an upper bound: keep them across calls
    ScanPlanner pl;
    if (int rc = pl.begin(t, order, first, count, mintrav, maxtrav, 1u, plan, true)) return rc;
    pl.add_parallel(0, count, plan_threads());
    pl.finish();
    if (n_cand) *n_cand = plan.n_cand;
    if (plan.n_cand > capacity) { set_error("candidate capacity too small"); return 1; }
    memcpy(visit_begin, plan.visit_begin.data(), plan.visit_begin.size() * sizeof(int32_t));
    if (cand_ref) memcpy(cand_ref, plan.cand_ref.data(), (size_t)plan.n_cand * sizeof(int32_t));
    if (cand_prune) memcpy(cand_prune, plan.cand_prune.data(), (size_t)plan.n_cand * sizeof(int32_t));
    return 0;
}

int mpgpu_host_plan_threads(void) { return plan_threads(); }

int mpgpu_host_plan_selftest(int ntaxa, const int32_t *back_node, const int32_t *back_slot, const int32_t *order, int first, int count,
                             int mintrav, int maxtrav, int nthreads, int pieces)
{
    HostTree t;
    if (int rc = host_tree_from(ntaxa, back_node, back_slot, t)) return rc;
    if (!order || first < 1 || count < 0 || first + count > 2 * ntaxa - 1 || pieces < 1) { set_error("bad visit range"); return 1; }
    ScanPlan a, b;
    {
        ScanPlanner pl;
        if (int rc = pl.begin(t, order, first, count, mintrav, maxtrav, 7u, a, true)) return rc;
        pl.add(0, count);
        pl.finish();
    }
    {
        ScanPlanner pl;
        if (int rc = pl.begin(t, order, first, count, mintrav, maxtrav, 7u, b, true)) return rc;
        int v0 = 0;
        for (int k = 0; k < pieces; k++) {
            const int v1 = k == pieces - 1 ? count : std::min(count, (int)((long long)count * (k + 1) / pieces));
            pl.add_parallel(v0, v1, nthreads);
            v0 = v1;
        }
        pl.finish();
    }
    const char *bad = nullptr;
    if (a.n_cand != b.n_cand || a.n_ops != b.n_ops || a.max_slot != b.max_slot || a.task_cap != b.task_cap) bad = "counts";
    else if (a.tasks.size() != b.tasks.size() || (a.tasks.size() && memcmp(a.tasks.data(), b.tasks.data(), a.tasks.size() * sizeof(ScanTask)))) bad = "tasks";
    else if (memcmp(a.offs_host.data(), b.offs_host.data(), (size_t)a.n_ops * sizeof(ScanOffs))) bad = "view offsets";
    else if (memcmp(a.ctl_host.data(), b.ctl_host.data(), (size_t)a.n_ops * sizeof(ScanCtl))) bad = "control words";
    else if (memcmp(a.cand_ref.data(), b.cand_ref.data(), (size_t)a.n_cand * 4) || memcmp(a.cand_prune.data(), b.cand_prune.data(), (size_t)a.n_cand * 4) ||
             memcmp(a.cand_task.data(), b.cand_task.data(), (size_t)a.n_cand * 4)) bad = "candidate tables";
    else if (a.visit_begin != b.visit_begin || a.visit_ref != b.visit_ref || a.task_vids != b.task_vids) bad = "visit tables";
    if (bad) { set_error(std::string("parallel enumeration differs from the sequential one: ") + bad); return 2; }
    return 0;
}

int mpgpu_host_apply_spr(int ntaxa, int32_t *back_node, int32_t *back_slot, int32_t remove_ref, int32_t insert_ref)
{
    HostTree t;
    if (int rc = host_tree_from(ntaxa, back_node, back_slot, t)) return rc;
    const int len = 3 * (2 * ntaxa - 1);
    if (remove_ref < 3 * (ntaxa + 1) || remove_ref >= len || insert_ref < 3 || insert_ref >= len) { set_error("bad move"); return 1; }
    apply_spr_move(t, remove_ref, insert_ref);
    memcpy(back_node, t.bn.data(), (size_t)len * sizeof(int32_t));
    memcpy(back_slot, t.bs.data(), (size_t)len * sizeof(int32_t));
    return 0;
}

}  // extern "C"
