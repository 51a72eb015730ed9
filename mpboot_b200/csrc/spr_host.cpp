// Host side of the SPR scan: the reference's tree walking, restated on ring tables so that the
// device can score whole prune neighbourhoods in one launch while the host keeps the exact
// visit order, candidate order and RNG discipline of the reference.
//
//   visit_order        nodeRectifierPars / reorderNodes      (sprparsimony.cpp:2046-2101)
//   build_scan_plan    rearrangeParsimony + addTraverseParsimony enumeration
//                                                           (sprparsimony.cpp:2208-2218, 2259-2376)
//   apply_spr_move     restoreTreeRearrangeParsimony         (sprparsimony.cpp:2379-2384, 2191-2205)
#include "mpgpu_internal.h"

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

namespace {

struct Builder {
    const HostTree &t;
    ScanPlan &plan;
    int prune_ref;
    int task_index;
    uint32_t vstride;                                               // view stride in vector units
    Builder(const HostTree &tt, ScanPlan &pp, uint32_t vs) : t(tt), plan(pp), prune_ref(0), task_index(0), vstride(vs) {}
    int32_t voff(int ref) const { return (int32_t)((uint32_t)t.vid(ref) * vstride); }

    int cand_base = 0;

    int new_op(uint32_t src, int c1, int c2)
    {
        ScanOffs o; o.c1 = voff(c1); o.c2 = voff(c2);
        ScanCtl c; c.outs = 0xFFFFFFFFu; c.meta = src | 0xFF00u | 0xFF0000u;
        plan.offs.push_back(o); plan.ctl.push_back(c);
        return (int)plan.offs.size() - 1;
    }
    void set_out(int op, int which, int idx)
    {
        const uint32_t rel = (uint32_t)(idx - cand_base);
        uint32_t &w = plan.ctl[op].outs;
        w = which == 0 ? ((w & 0xFFFF0000u) | rel) : ((w & 0x0000FFFFu) | (rel << 16));
    }
    void set_dst(int op, int which, int slot)
    {
        uint32_t &w = plan.ctl[op].meta;
        w = which == 0 ? ((w & ~0xFF00u) | ((uint32_t)slot << 8)) : ((w & ~0xFF0000u) | ((uint32_t)slot << 16));
    }

    // addTraverseParsimony(tr, pr, p, q, mintrav, maxtrav, doAll = FALSE) for q = x.
    // parent_op/which identify the expand op that scores x; depth is x's distance (1-based).
    void traverse(int x, int mintrav, int maxtrav, int parent_op, int which, int depth)
    {
        if (--mintrav <= 0) {                                   // testInsertParsimony(p, x)
            int idx = plan.n_cand++;
            plan.cand_ref.push_back(x);
            plan.cand_prune.push_back(prune_ref);
            plan.cand_task.push_back(task_index);
            set_out(parent_op, which, idx);
        }
        if (!t.is_tip(x) && (--maxtrav > 0)) {
            const int slot = 2 * (depth - 1) + which;           // U_x goes here
            set_dst(parent_op, which, slot);
            if (slot + 1 > plan.max_slot) plan.max_slot = slot + 1;
            const int c1 = t.back(t.next(x)), c2 = t.back(t.next(t.next(x)));
            const int me = new_op((uint32_t)slot, c1, c2);
            traverse(c1, mintrav, maxtrav, me, 0, depth + 1);
            traverse(c2, mintrav, maxtrav, me, 1, depth + 1);
        }
    }

    // the two addTraverseParsimony calls made for one inner neighbour `nb` of the removed node:
    // candidates are the branches to nb's children; the far side of nb is the task's D2 (when nb
    // is the D1 neighbour, src code 0xFF) or D1 (src code 0xFE).
    void expand_top(int nb, uint32_t src_code, int mintrav, int maxtrav)
    {
        const int c1 = t.back(t.next(nb)), c2 = t.back(t.next(t.next(nb)));
        const int me = new_op(src_code, c1, c2);
        traverse(c1, mintrav, maxtrav, me, 0, 1);
        traverse(c2, mintrav, maxtrav, me, 1, 1);
    }
};

}  // namespace

// Enumerates what rearrangeParsimony(tr, pr, tr->nodep[i], mintrav, maxtrav, doAll=FALSE) tests
// for i in [first, first+count).  Returns 0, or 1 if maxtrav exceeds the kernel's stack.
int build_scan_plan(const HostTree &t, const std::vector<uint32_t> &vlen, const int32_t *order,
                    int first, int count, int mintrav, int maxtrav_in, uint32_t vstride, ScanPlan &plan)
{
    plan.offs.clear(); plan.ctl.clear(); plan.tasks.clear(); plan.visit_begin.clear();
    plan.cand_ref.clear(); plan.cand_prune.clear(); plan.cand_task.clear(); plan.task_const.clear();
    plan.n_cand = 0; plan.max_slot = 0;
    const int n = t.n;
    int maxtrav = maxtrav_in;
    if (maxtrav > n - 3) maxtrav = n - 3;                       // :2275 (tr->ntips == mxtips during the search)
    if (maxtrav > kMaxTrav) { set_error("maxtrav exceeds the scan kernel's stack depth"); return 1; }   // slots < 0xFE
    if ((uint64_t)(4 * n - 6) * vstride >= 0x7fffffffULL) { set_error("view array too large for 32-bit scan offsets"); return 1; }
    Builder b(t, plan, vstride);

    for (int v = 0; v < count; v++) {
        plan.visit_begin.push_back(plan.n_cand);
        if (maxtrav < mintrav) continue;                        // :2280
        const int p = order[first + v];
        const int q = t.back(p);

        if (!t.is_tip(p)) {                                     // :2303
            const int p1 = t.back(t.next(p)), p2 = t.back(t.next(t.next(p)));
            if (!t.is_tip(p1) || !t.is_tip(p2)) {
                ScanTask task;
                task.s_vid = b.voff(q); task.d1 = b.voff(p1); task.d2 = b.voff(p2);
                task.op_begin = (int)plan.offs.size(); task.base_out = 0; task.cand_base = plan.n_cand; task.pad = 0; b.cand_base = plan.n_cand;
                b.prune_ref = p; b.task_index = (int)plan.tasks.size();
                if (!t.is_tip(p1)) b.expand_top(p1, 0xFFu, mintrav, maxtrav);
                if (!t.is_tip(p2)) b.expand_top(p2, 0xFEu, mintrav, maxtrav);
                task.op_end = (int)plan.offs.size();
                plan.tasks.push_back(task);
                plan.task_const.push_back(vlen[t.vid(q)] + vlen[t.vid(p1)] + vlen[t.vid(p2)]);
            }
        }
        if (!t.is_tip(q) && maxtrav > 0) {                      // :2333
            const int q1 = t.back(t.next(q)), q2 = t.back(t.next(t.next(q)));
            const bool ok1 = !t.is_tip(q1) && (!t.is_tip(t.back(t.next(q1))) || !t.is_tip(t.back(t.next(t.next(q1)))));
            const bool ok2 = !t.is_tip(q2) && (!t.is_tip(t.back(t.next(q2))) || !t.is_tip(t.back(t.next(t.next(q2)))));
            if (ok1 || ok2) {
                const int mintrav2 = mintrav > 2 ? mintrav : 2;
                ScanTask task;
                task.s_vid = b.voff(p); task.d1 = b.voff(q1); task.d2 = b.voff(q2);
                task.op_begin = (int)plan.offs.size(); task.base_out = 0; task.cand_base = plan.n_cand; task.pad = 0; b.cand_base = plan.n_cand;
                b.prune_ref = q; b.task_index = (int)plan.tasks.size();
                if (!t.is_tip(q1)) b.expand_top(q1, 0xFFu, mintrav2, maxtrav);
                if (!t.is_tip(q2)) b.expand_top(q2, 0xFEu, mintrav2, maxtrav);
                task.op_end = (int)plan.offs.size();
                plan.tasks.push_back(task);
                plan.task_const.push_back(vlen[t.vid(p)] + vlen[t.vid(q1)] + vlen[t.vid(q2)]);
            }
        }
    }
    plan.visit_begin.push_back(plan.n_cand);
    // base counters live after the candidate counters in the device output vector
    for (size_t i = 0; i < plan.tasks.size(); i++) plan.tasks[i].base_out = plan.n_cand + (int)i;
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
