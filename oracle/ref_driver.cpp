/* TEST INFRASTRUCTURE -- not product code.  Built only into oracle/_ref/libmpref.so.
 *
 * Drives the UNMODIFIED reference parsimony engine: this translation unit textually
 * includes /root/reference/sprparsimony.cpp (compiled where it lies, never copied), which
 * gives the driver access to its file-static kernels (newviewParsimonyIterativeFast :554,
 * evaluateParsimonyIterativeFast :965, rearrangeParsimony :2259, testInsertParsimony :2106,
 * compressDNA :2828 ...) as well as the exported entry points (pllOptimizeSprParsimony
 * :3244, _pllComputeRandomizedStepwiseAdditionParsimonyTree :3224,
 * pllComputePatternParsimony :3363).  The PLL helpers it needs (hookupDefault, randum,
 * getBitVector, pllBaseSubstitute, mask32 ...) come from /root/reference/pllrepo/src/utils.c,
 * compiled separately by oracle/Makefile.
 *
 * What is NOT the reference here (and therefore stated explicitly):
 *   - class IQTree is the stand-in of oracle/shim/mp_iqtree_shim.h; its saveCurrentTree()
 *     forwards to a recorder instead of running iqtree.cpp:3271.
 *   - random_double() (tools.cpp:3362, SPRNG) is replaced by a splitmix64 stream shared
 *     with the product tests: parity is about *which* draws are made in *which* order.
 *   - the REPS loop (iqtree.cpp:3411-3449) is re-typed below on the reference's own
 *     vectorclass Vec16us so the 16-bit wrap semantics are the library's, not ours.
 */
#include "mp_iqtree_shim.h"
#include "sprparsimony.cpp"          /* -I /root/reference */

#include <stdint.h>
#include <vector>

extern "C" void pllBaseSubstitute(pllInstance *tr, partitionList *partitions);   /* utils.c:2526 */

/* Sankoff-only helper referenced from compressSankoffDNA (sprparsimony.cpp:2809): the weight of a
 * minimum spanning tree over the unambiguous states present in a pattern, under the cost matrix.
 * ParsTree lives in parstree.cpp, which cannot be compiled without the whole program, so this is
 * NOT the reference: it is re-typed from parstree.cpp:606-677 (Prim from the first present state,
 * ties to the lowest state index) on the PLL codes of the pattern.  An MST's weight does not depend
 * on the tie-breaking, so any correct MST gives the reference's number. */
static struct mpref *g_mst_handle = NULL;
static int mst_score_of_pattern(struct mpref *h, int ptn);
ParsTree::ParsTree(Alignment *a) { aln = a; }
int ParsTree::findMstScore(int ptn) { return mst_score_of_pattern(g_mst_handle, ptn); }

/* ---- globals the engine expects from iqtree.cpp / tools.cpp ------------------------- */
Params *globalParam = NULL;                 /* iqtree.cpp:601 */
parsimonyNumber *pllCostMatrix = NULL;      /* iqtree.cpp:  Sankoff cost matrix (-cost), NULL = Fitch */
int pllCostNstates = 0;
parsimonyNumber *vectorCostMatrix = NULL;
int pllRepsSegments = -1;
int *pllSegmentUpper = NULL;

static uint64_t g_rng_state = 0x9E3779B97F4A7C15ULL;
static uint64_t g_rng_draws = 0;

#define MPREF_API extern "C" __attribute__((visibility("default")))

/* splitmix64 -> [0,1) ; the same generator is exported by the C port oracle */
MPREF_API double mpref_random_double(void *unused)
{
    (void)unused;
    uint64_t z = (g_rng_state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    g_rng_draws++;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
double random_double() { return mpref_random_double(NULL); }
MPREF_API void mpref_seed_rng(uint64_t seed) { g_rng_state = seed; g_rng_draws = 0; }
MPREF_API uint64_t mpref_rng_draws(void) { return g_rng_draws; }

void outError(const char *error, bool quit) { fprintf(stderr, "mpref outError: %s\n", error); if (quit) exit(2); }

/* ---- handle ------------------------------------------------------------------------- */
struct mpref {
    int n, P, datatype, states;
    pllInstance *tr;
    partitionList *pr;
    pInfo *pinfo;
    node *pool;
    std::vector<nodeptr> base;            /* base[i] = ring slot 0 of node i (1..2n-2) */
    std::vector<unsigned char> ybuf;
    ParsTree iq;                          /* compressSankoffDNA dynamic_casts iqtree to ParsTree (:2809) */
    Alignment aln;
    std::vector<parsimonyNumber> cost;    /* Sankoff: this handle's cost matrix and segments (globals point here) */
    std::vector<int> seg_upper;
    mpref() : iq((Alignment *)NULL) {}
    Params params;
    /* recorder for the saveCurrentTree up-call */
    std::vector<int> saved_mp;
    std::vector<unsigned short> saved_ptn;   /* P per call when record_ptn */
    int record_ptn;
    bool allocated;
    struct BootSim *boot;                 /* -bb bookkeeping of IQTree::saveCurrentTree (re-typed below) */
};

static void boot_save_current_tree(mpref *h, double cur_logl);


static void save_hook(void *user, double cur_logl)
{
    mpref *h = (mpref *)user;
    if (h->boot) boot_save_current_tree(h, cur_logl);
    h->saved_mp.push_back((int)(-cur_logl));
    if (h->record_ptn) {
        size_t off = h->saved_ptn.size();
        h->saved_ptn.resize(off + h->P, 0);
        int sum = 0;
        /* what IQTree::saveCurrentTree does at iqtree.cpp:3365 */
        pllComputePatternParsimony(h->tr, h->pr, &h->saved_ptn[off], &sum);
    }
}

static int states_of(int datatype)
{
    switch (datatype) {
    case PLL_BINARY_DATA: return 2;
    case PLL_DNA_DATA:    return 4;
    case PLL_AA_DATA:     return 20;
    case PLL_GENERIC_32:  return 32;
    default: return -1;
    }
}

/* chars: [n][P] raw alignment characters (ASCII), mapped with the reference's own tables
 * through pllBaseSubstitute (utils.c:2526). weights: pattern frequencies (tr->aliaswgt). */
MPREF_API mpref *mpref_create(int n, int P, int datatype, const unsigned char *chars, const int *weights,
                             int sort_alignment, int n_informative)
{
    int S = states_of(datatype);
    if (S < 0 || n < 4 || P < 1) return NULL;
    mpref *h = new mpref();
    h->n = n; h->P = P; h->datatype = datatype; h->states = S;
    h->record_ptn = 0; h->allocated = false; h->boot = NULL;

    pllInstance *tr = (pllInstance *)calloc(1, sizeof(pllInstance));
    partitionList *pr = (partitionList *)calloc(1, sizeof(partitionList));
    pInfo *pi = (pInfo *)calloc(1, sizeof(pInfo));
    h->tr = tr; h->pr = pr; h->pinfo = pi;

    pr->numberOfPartitions = 1;
    pr->perGeneBranchLengths = PLL_FALSE;
    pr->partitionData = (pInfo **)calloc(1, sizeof(pInfo *));
    pr->partitionData[0] = pi;
    pi->dataType = datatype;
    pi->states = S;
    pi->lower = 0; pi->upper = P; pi->width = P;
    pi->parsVect = NULL; pi->perSitePartialPars = NULL;
    pi->informativePtnWgt = NULL; pi->informativePtnScore = NULL;

    tr->mxtips = n;
    tr->originalCrunchedLength = P;
    tr->aliaswgt = (int *)malloc(sizeof(int) * P);
    memcpy(tr->aliaswgt, weights, sizeof(int) * P);
    h->ybuf.assign(chars, chars + (size_t)n * P);
    tr->yVector = (unsigned char **)calloc(n + 1, sizeof(unsigned char *));
    for (int i = 1; i <= n; i++) tr->yVector[i] = &h->ybuf[(size_t)(i - 1) * P];
    pllBaseSubstitute(tr, pr);

    /* node rings laid out as pllTreeInitDefaults does (utils.c:1963-2041) */
    int inner = n - 1;
    h->pool = (node *)calloc(n + 3 * inner, sizeof(node));
    tr->nodep = (nodeptr *)calloc(2 * n, sizeof(nodeptr));
    tr->constraintVector = (int *)calloc(2 * n, sizeof(int));
    h->base.assign(2 * n, (nodeptr)NULL);
    node *p0 = h->pool;
    for (int i = 1; i <= n; i++) {
        node *p = p0++;
        p->number = i; p->next = p; p->back = NULL;
        tr->nodep[i] = p; h->base[i] = p;
    }
    for (int i = n + 1; i <= n + inner; i++) {
        node *a = p0++, *b = p0++, *c = p0++;      /* slot 0, 1, 2 */
        a->number = b->number = c->number = i;
        a->next = b; b->next = c; c->next = a;
        a->xPars = 1; a->x = 1; a->xBips = 1;
        tr->nodep[i] = a; h->base[i] = a;
    }
    tr->start = tr->nodep[1];
    tr->ntips = n;
    tr->nextnode = 2 * n - 1;
    tr->grouped = PLL_FALSE; tr->constrained = PLL_FALSE;
    tr->parsimonyScore = NULL; tr->ti = NULL;
    tr->bestParsimony = UINT_MAX;

    /* Params: only the fields the engine reads */
    memset((void *)&h->params, 0, sizeof(Params));
    h->params.gbo_replicates = 0;
    h->params.ratchet_iter = -1;
    h->params.sort_alignment = sort_alignment ? true : false;
    h->params.sankoff_short_int = true;
    globalParam = &h->params;

    h->aln.resize(P);
    for (int i = 0; i < P; i++) { h->aln[i].frequency = weights[i]; h->aln[i].is_const = false; h->aln[i].ras_pars_score = 0; }
    h->aln.n_informative_patterns = n_informative;
    h->iq.aln = &h->aln; h->iq.pllInst = tr; h->iq.pllPartitions = pr;
    h->iq.hook = save_hook; h->iq.hook_user = h;
    iqtree = &h->iq;
    first_call = true;
    return h;
}

static void make_current(mpref *h)
{
    globalParam = &h->params; iqtree = &h->iq; g_mst_handle = h;
    if (!h->cost.empty()) {
        pllCostMatrix = &h->cost[0]; pllCostNstates = h->states;
        pllRepsSegments = (int)h->seg_upper.size(); pllSegmentUpper = &h->seg_upper[0];
    } else {
        pllCostMatrix = NULL; pllCostNstates = 0;
    }
}

/* -cost: what ParsTree::initParsData / IQTree::doSegmenting leave in the globals (pllCostMatrix,
 * pllCostNstates, pllRepsSegments, pllSegmentUpper) before the engine runs, then initializeCostMatrix
 * (sprparsimony.cpp:159).  cost[i*S+j] = cost from state i to j.  nseg/segment_upper as doSegmenting. */
MPREF_API int mpref_set_cost_matrix(mpref *h, const unsigned int *cost, const int *segment_upper, int nseg)
{
    if (h->allocated) { make_current(h); _pllFreeParsimonyDataStructures(h->tr, h->pr); h->allocated = false; }
    if (vectorCostMatrix) { rax_free(vectorCostMatrix); vectorCostMatrix = NULL; }
    if (!cost) { h->cost.clear(); h->seg_upper.clear(); make_current(h); return 0; }
    h->cost.assign(cost, cost + (size_t)h->states * h->states);
    h->seg_upper.assign(segment_upper, segment_upper + nseg);
    make_current(h);
    initializeCostMatrix();
    first_call = true;
    return (int)highest_cost;
}

/* -short_off (tools.cpp:2365): 32-bit Sankoff vectors and sums (Vec8ui instead of Vec16us); call before mpref_set_cost_matrix */
MPREF_API void mpref_set_sankoff_short(mpref *h, int on) { h->params.sankoff_short_int = on != 0; }

/* Sankoff parsVect of one node as the engine holds it: u16 [parsimonyLength/16][S][16] (:2725-2733) */
MPREF_API int mpref_get_sankoff_vect(mpref *h, int node, unsigned short *out)
{
    size_t L = h->pinfo->parsimonyLength, S = h->states;
    const unsigned short *v = (const unsigned short *)&h->pinfo->parsVect[L * S * (size_t)node];
    memcpy(out, v, L * S * sizeof(unsigned short));
    return (int)L;
}
MPREF_API void mpref_set_best(mpref *h, unsigned int best) { h->tr->bestParsimony = best; }
MPREF_API int mpref_remainder_bounds(mpref *h, unsigned int *out)
{
    (void)h;
    if (!pllRemainderLowerBounds) return 0;
    for (int i = 0; i < pllRepsSegments - 1; i++) out[i] = pllRemainderLowerBounds[i];
    return pllRepsSegments - 1;
}

static int mst_score_of_pattern(mpref *h, int ptn)
{
    const int S = h->states;
    const unsigned int *bv = getBitVector(h->datatype);
    std::vector<char> present(S, 0);
    for (int t = 1; t <= h->n; t++) {
        unsigned int m = bv[h->tr->yVector[t][ptn]];
        if (m && !(m & (m - 1))) present[__builtin_ctz(m)] = 1;      /* pat[j] < num_states: unambiguous only */
    }
    int cnt = 0;
    for (int i = 0; i < S; i++) cnt += present[i];
    if (cnt <= 1) return 0;
    std::vector<unsigned int> label(S, UINT_MAX);
    std::vector<char> added(S, 0);
    for (int c = 0; c < S; c++) if (present[c]) { label[c] = 0; break; }
    unsigned int score = 0;
    for (int it = 0; it < cnt; it++) {
        int add = -1; unsigned int best = UINT_MAX;
        for (int c = 0; c < S; c++) if (present[c] && !added[c] && label[c] < best) { best = label[c]; add = c; }
        if (add < 0) break;
        added[add] = 1; score += label[add];
        for (int c = 0; c < S; c++)
            if (present[c] && !added[c] && label[c] > h->cost[(size_t)add * S + c]) label[c] = h->cost[(size_t)add * S + c];
    }
    return (int)score;
}

MPREF_API void mpref_boot_free(mpref *h);
MPREF_API void mpref_destroy(mpref *h)
{
    if (!h) return;
    make_current(h);
    if (h->allocated) _pllFreeParsimonyDataStructures(h->tr, h->pr);
    free(h->tr->aliaswgt); free(h->tr->yVector); free(h->tr->nodep); free(h->tr->constraintVector);
    free(h->pool); free(h->pr->partitionData); free(h->pinfo); free(h->pr); free(h->tr);
    if (globalParam == &h->params) globalParam = NULL;
    if (iqtree == &h->iq) iqtree = NULL;
    mpref_boot_free(h);
    delete h;
}

/* PLL code of every input byte for a data type (pins the char->code tables utils.c:98-157) */
MPREF_API int mpref_char_map(int datatype, unsigned char out[256])
{
    if (states_of(datatype) < 0) return -1;
    pllInstance tr; partitionList pr; pInfo pi; pInfo *pip = &pi;
    memset(&tr, 0, sizeof tr); memset(&pr, 0, sizeof pr); memset(&pi, 0, sizeof pi);
    unsigned char buf[256]; unsigned char *yv[2] = {NULL, buf};
    for (int i = 0; i < 256; i++) buf[i] = (unsigned char)i;
    tr.mxtips = 1; tr.yVector = yv;
    pr.numberOfPartitions = 1; pr.partitionData = &pip;
    pi.dataType = datatype; pi.lower = 0; pi.upper = 256;
    pllBaseSubstitute(&tr, &pr);
    memcpy(out, buf, 256);
    return 0;
}

/* PLL state-set bit mask of every code (globalVariables.h:60-104) */
MPREF_API int mpref_bitvector(int datatype, unsigned int *out, int ncodes)
{
    const unsigned int *bv = getBitVector(datatype);
    for (int i = 0; i < ncodes; i++) out[i] = bv[i];
    return getUndetermined(datatype);
}

static nodeptr slot_ptr(mpref *h, int number, int slot)
{
    nodeptr p = h->base[number];
    for (int s = 0; s < slot; s++) p = p->next;
    return p;
}
static int slot_of(mpref *h, nodeptr p)
{
    nodeptr b = h->base[p->number];
    if (p == b) return 0;
    if (p == b->next) return 1;
    return 2;
}

/* ring tables: index (node*3 + slot), node in 1..2n-2; tips use slot 0 only.
 * back_node = 0 means NULL. Resets orientation flags like _allocateParsimonyDataStructures. */
MPREF_API void mpref_set_ring(mpref *h, const int *back_node, const int *back_slot)
{
    int n = h->n;
    for (int i = 1; i <= 2 * n - 2; i++) {
        int ns = (i <= n) ? 1 : 3;
        for (int s = 0; s < ns; s++) {
            nodeptr p = slot_ptr(h, i, s);
            int bn = back_node[i * 3 + s];
            p->back = bn ? slot_ptr(h, bn, back_slot[i * 3 + s]) : (nodeptr)NULL;
            p->z[0] = PLL_DEFAULTZ;
        }
    }
    for (int i = 1; i <= 2 * n - 2; i++) h->tr->nodep[i] = h->base[i];
    for (int i = n + 1; i <= 2 * n - 2; i++) {
        nodeptr p = h->base[i];
        p->xPars = 1; p->next->xPars = 0; p->next->next->xPars = 0;
    }
    h->tr->start = h->tr->nodep[1];
    h->tr->ntips = n;
    h->tr->nextnode = 2 * n - 1;
}

MPREF_API void mpref_get_ring(mpref *h, int *back_node, int *back_slot)
{
    int n = h->n;
    for (int i = 1; i <= 2 * n - 2; i++) {
        int ns = (i <= n) ? 1 : 3;
        for (int s = 0; s < 3; s++) { back_node[i * 3 + s] = 0; back_slot[i * 3 + s] = 0; }
        for (int s = 0; s < ns; s++) {
            nodeptr p = slot_ptr(h, i, s);
            if (p->back) { back_node[i * 3 + s] = p->back->number; back_slot[i * 3 + s] = slot_of(h, p->back); }
        }
    }
}

/* current tr->nodep[] visit order as (node, slot) pairs, index 1..2n-2 */
MPREF_API void mpref_get_nodep(mpref *h, int *node_out, int *slot_out)
{
    for (int i = 1; i <= 2 * h->n - 2; i++) {
        node_out[i] = h->tr->nodep[i]->number;
        slot_out[i] = slot_of(h, h->tr->nodep[i]);
    }
}

MPREF_API void mpref_set_weights(mpref *h, const int *weights)
{
    memcpy(h->tr->aliaswgt, weights, sizeof(int) * h->P);
    for (int i = 0; i < h->P; i++) h->aln[i].frequency = weights[i];
}

/* _allocateParsimonyDataStructures (sprparsimony.cpp:3032): returns parsimonyLength W */
MPREF_API int mpref_allocate(mpref *h, int perSiteScores)
{
    make_current(h);
    if (h->allocated) _pllFreeParsimonyDataStructures(h->tr, h->pr);
    _allocateParsimonyDataStructures(h->tr, h->pr, perSiteScores);
    h->allocated = true;
    h->params.gbo_replicates = perSiteScores ? 1000 : 0;
    return (int)h->pinfo->parsimonyLength;
}

MPREF_API int mpref_num_informative(mpref *h) { return h->pinfo->numInformativePatterns; }

MPREF_API void mpref_get_parsvect(mpref *h, int node, unsigned int *out)
{
    size_t W = h->pinfo->parsimonyLength, S = h->states;
    memcpy(out, &h->pinfo->parsVect[W * S * (size_t)node], W * S * sizeof(unsigned int));
}
MPREF_API unsigned int mpref_node_score(mpref *h, int node) { return h->tr->parsimonyScore[node]; }

/* nodeRectifierPars + evaluateParsimony(start, full) as pllOptimizeSprParsimony :3275-3277 */
MPREF_API unsigned int mpref_evaluate_full(mpref *h, int perSiteScores)
{
    make_current(h);
    nodeRectifierPars(h->tr);
    h->tr->bestParsimony = UINT_MAX;
    return evaluateParsimony(h->tr, h->pr, h->tr->start, PLL_TRUE, perSiteScores);
}

/* evaluateParsimony at an arbitrary ring slot, lazily (full = FALSE) */
MPREF_API unsigned int mpref_evaluate_at(mpref *h, int node, int slot, int full, int perSiteScores)
{
    make_current(h);
    return evaluateParsimony(h->tr, h->pr, slot_ptr(h, node, slot), full ? PLL_TRUE : PLL_FALSE, perSiteScores);
}

MPREF_API void mpref_pattern_parsimony(mpref *h, unsigned short *out, int *sum)
{
    make_current(h);
    pllComputePatternParsimony(h->tr, h->pr, out, sum);
}

MPREF_API int mpref_min_pars_pattern(mpref *h, int site)
{
    return pllCalcMinParsScorePattern(h->tr, h->datatype, site);
}

MPREF_API void mpref_record(mpref *h, int record_ptn) { h->record_ptn = record_ptn; h->saved_mp.clear(); h->saved_ptn.clear(); }
MPREF_API int mpref_saved_count(mpref *h) { return (int)h->saved_mp.size(); }
MPREF_API void mpref_saved_mp(mpref *h, int *out) { if (!h->saved_mp.empty()) memcpy(out, &h->saved_mp[0], h->saved_mp.size() * sizeof(int)); }
MPREF_API void mpref_saved_ptn(mpref *h, unsigned short *out) { if (!h->saved_ptn.empty()) memcpy(out, &h->saved_ptn[0], h->saved_ptn.size() * sizeof(unsigned short)); }

/* One node visit of the SPR sweep (sprparsimony.cpp:3301-3305) on visit index i of the
 * current tr->nodep[] order.  best_in: value of tr->bestParsimony on entry.
 * out[0]=bestParsimony after, out[1]=removeNode number (0 none), out[2]=its slot,
 * out[3]=insertNode number, out[4]=its slot, out[5]=bestTreeScoreHits.
 * With perSiteScores=1 every scored insertion is reported through the recorder. */
MPREF_API int mpref_rearrange(mpref *h, int i, int mintrav, int maxtrav, int perSiteScores,
                             unsigned int best_in, unsigned int *out)
{
    make_current(h);
    pllInstance *tr = h->tr;
    tr->insertNode = NULL; tr->removeNode = NULL;
    bestTreeScoreHits = 1;
    tr->bestParsimony = best_in;
    tr->ntips = tr->mxtips;
    int rc = rearrangeParsimony(tr, h->pr, tr->nodep[i], mintrav, maxtrav, PLL_FALSE, perSiteScores);
    out[0] = tr->bestParsimony;
    out[1] = tr->removeNode ? tr->removeNode->number : 0;
    out[2] = tr->removeNode ? slot_of(h, tr->removeNode) : 0;
    out[3] = tr->insertNode ? tr->insertNode->number : 0;
    out[4] = tr->insertNode ? slot_of(h, tr->insertNode) : 0;
    out[5] = (unsigned int)bestTreeScoreHits;
    return rc;
}

/* restoreTreeRearrangeParsimony (sprparsimony.cpp:2379): apply the recorded move */
MPREF_API void mpref_apply_move(mpref *h, int perSiteScores)
{
    make_current(h);
    restoreTreeRearrangeParsimony(h->tr, h->pr, perSiteScores);
}
MPREF_API void mpref_node_rectifier(mpref *h) { nodeRectifierPars(h->tr); }

/* The real search: pllOptimizeSprParsimony (sprparsimony.cpp:3244).
 * cur_score must equal the tree's score (assert :3279). Returns startMP. */
MPREF_API int mpref_optimize_spr(mpref *h, int mintrav, int maxtrav, int bb, int ratchet_realloc)
{
    make_current(h);
    h->params.gbo_replicates = bb ? 1000 : 0;
    h->params.ratchet_iter = ratchet_realloc ? 1 : -1;
    h->iq.on_ratchet_hclimb1 = ratchet_realloc ? true : false;
    h->iq.on_ratchet_hclimb2 = false;
    h->iq.on_opt_btree = false;
    if (!h->allocated || ratchet_realloc) {
        first_call = true;
    } else {
        first_call = false;
    }
    /* the assert at :3279 needs iqtree->curScore; compute it with the same engine first */
    if (first_call) {
        if (h->allocated) { _pllFreeParsimonyDataStructures(h->tr, h->pr); h->allocated = false; }
        _allocateParsimonyDataStructures(h->tr, h->pr, bb);
        h->allocated = true;
        first_call = false;
        h->iq.on_ratchet_hclimb1 = false;
        h->params.ratchet_iter = -1;
    }
    nodeRectifierPars(h->tr);
    h->tr->bestParsimony = UINT_MAX;
    size_t keep = h->saved_mp.size(), keepp = h->saved_ptn.size();
    unsigned int s0 = evaluateParsimony(h->tr, h->pr, h->tr->start, PLL_TRUE, bb);
    h->saved_mp.resize(keep); h->saved_ptn.resize(keepp);
    h->iq.curScore = -(double)s0;
    return pllOptimizeSprParsimony(h->tr, h->pr, mintrav, maxtrav, &h->iq);
}

/* _pllComputeRandomizedStepwiseAdditionParsimonyTree (sprparsimony.cpp:3224) */
MPREF_API unsigned int mpref_ras(mpref *h, long seed, int sprDist)
{
    make_current(h);
    if (h->allocated) { _pllFreeParsimonyDataStructures(h->tr, h->pr); h->allocated = false; }
    for (int i = 1; i <= 2 * h->n - 2; i++) h->tr->nodep[i] = h->base[i];
    for (int i = 1; i <= 2 * h->n - 2; i++) {
        nodeptr p = h->base[i];
        int ns = (i <= h->n) ? 1 : 3;
        for (int s = 0; s < ns; s++, p = p->next) p->back = NULL;
    }
    h->tr->randomNumberSeed = seed;
    h->params.gbo_replicates = 0;
    _pllComputeRandomizedStepwiseAdditionParsimonyTree(h->tr, h->pr, sprDist, &h->iq);
    return h->tr->bestParsimony;
}

/* ---- REPS (iqtree.cpp:3411-3449) on the reference's own Vec16us ------------------- */
/* pars, w: u16 arrays padded to a multiple of 16 (zeros) and 32-byte aligned by the caller.
 * res[b] = sum over segments of horizontal_add(sum_lanes u16(pars*w)) */
MPREF_API void mpref_reps(const unsigned short *pars, const unsigned short *boot, int B, int stride,
                         const int *segment_upper, int nseg, int *res_out)
{
    for (int b = 0; b < B; b++) {
        const unsigned short *w = boot + (size_t)b * stride;
        int ptn = 0, res = 0;
        Vec16us vc_rell = 0;
        for (int seg = 0; seg < nseg; seg++) {
            for (; ptn < segment_upper[seg]; ptn += 16)
                vc_rell = Vec16us().load(&pars[ptn]) * Vec16us().load(&w[ptn]) + vc_rell;
            res += horizontal_add(vc_rell);
            vc_rell = 0;
        }
        res_out[b] = res;
    }
}

/* ---- timing helper for bench.py --impl reference ----------------------------------- */
/* Runs the SPR sweep body (rearrangeParsimony over every node of the current tree, moves
 * NOT applied) `reps` times and returns the number of insertions scored. */
static unsigned long g_insert_counter = 0;
MPREF_API unsigned long mpref_sweep_count_insertions(mpref *h, int mintrav, int maxtrav, int perSiteScores, int reps)
{
    make_current(h);
    pllInstance *tr = h->tr;
    unsigned long total = 0;
    for (int r = 0; r < reps; r++) {
        nodeRectifierPars(tr);
        tr->bestParsimony = UINT_MAX;
        unsigned int s0 = evaluateParsimony(tr, h->pr, tr->start, PLL_TRUE, perSiteScores);
        tr->ntips = tr->mxtips;
        size_t before = h->saved_mp.size();
        for (int i = 1; i <= 2 * h->n - 2; i++) {
            tr->insertNode = NULL; tr->removeNode = NULL;
            bestTreeScoreHits = 1;
            tr->bestParsimony = s0;
            rearrangeParsimony(tr, h->pr, tr->nodep[i], mintrav, maxtrav, PLL_FALSE, perSiteScores);
        }
        total += (unsigned long)(h->saved_mp.size() - before);
    }
    (void)g_insert_counter;
    return total;
}

/* ---- -bb bookkeeping: IQTree::saveCurrentTree (iqtree.cpp:3271-3760), default policy ------
 * Re-typed (class IQTree cannot be linked, see the header of this file) for maximum_parsimony,
 * spr_parsimony, !store_candidate_trees, !multiple_hits, distinct_iter_top_boot < 1,
 * !auto_vectorize, !do_first_rell, outside ratchet iterations.  The REPS loop runs on the
 * reference's own Vec16us (load_a on 32-byte aligned buffers, as iqtree.cpp:3428); pattern
 * scores come from the reference's pllComputePatternParsimony; the skip bound is
 * pllComputeRellRemainBound (:3821-3858) over pllCalcMinParsScorePattern and ras_pars_score.
 * mpref_boot_set_mulhits / _topboot / _distinct / _ratchet switch on the other branches of the same function. */
#include <map>
#include <set>
struct BootSim {
    bool multiple_hits;                                       /* params->multiple_hits (-mulhits), :3498-3531 */
    std::vector<std::set<int> > boot_trees_parsimony;
    int store_top_boot_trees;                                 /* params->store_top_boot_trees (-mulhits -topboot N), :3536-3583 */
    std::vector<std::vector<std::pair<int, int> > > boot_trees_parsimony_top;   /* (tree_index, rell), decreasing */
    std::vector<int> boot_threshold;
    int distinct_iter_top_boot;                               /* params->distinct_iter_top_boot (-distinct_iter_top_boot K), :3587-3685 */
    int curIt;                                                /* IQTree::curIt of the search */
    std::vector<std::vector<int> > boot_trees_parsimony_top_iter;
    int B, stride, nseg;
    std::vector<unsigned short *> boot_samples_pars;          /* aligned, P+16, zero padded (:220-233) */
    std::vector<int> segment_upper;
    std::vector<std::vector<int> > remain;                   /* boot_samples_pars_remain_bounds */
    bool use_skip;
    std::vector<double> boot_logl, treels_logl;
    std::vector<int> boot_counts, boot_trees;
    double logl_cutoff, eps;
    unsigned short *pattern_pars;                            /* _pattern_pars */
    long calls, reps_rows, skipped, bad_sum;
    std::map<unsigned long long, int> treels;                /* topology fingerprint -> tree_index */
    std::vector<long long> mat;                              /* 5 per materialised tree */
    bool on_ratchet_hclimb1;                                 /* iqtree.cpp:3283: cur_logl from original_sample */
    unsigned short *original_sample;
};

static void *aligned32(size_t bytes) { void *p = NULL; if (posix_memalign(&p, 32, bytes)) return NULL; memset(p, 0, bytes); return p; }

static unsigned long long fin64(unsigned long long z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static unsigned long long topo_hash(pllInstance *tr, nodeptr p)
{
    if (p->number <= tr->mxtips) return fin64((unsigned long long)p->number);
    unsigned long long a = topo_hash(tr, p->next->back), b = topo_hash(tr, p->next->next->back);
    if (a > b) std::swap(a, b);
    return fin64(a * 0x9E3779B97F4A7C15ULL + b + 0x632BE59BD9B4E019ULL);
}
MPREF_API unsigned long long mpref_tree_fingerprint(mpref *h) { return fin64(topo_hash(h->tr, h->base[1]->back) ^ 0x1234567ULL); }

MPREF_API void mpref_boot_free(mpref *h)
{
    BootSim *b = h->boot;
    if (!b) return;
    for (size_t i = 0; i < b->boot_samples_pars.size(); i++) free(b->boot_samples_pars[i]);
    free(b->pattern_pars); free(b->original_sample);
    delete b;
    h->boot = NULL;
}

MPREF_API void mpref_boot_init(mpref *h, int B, const unsigned short *boot, int stride, const int *seg_upper, int nseg,
                              double logl_cutoff, double eps, const int *ras_pars_score)
{
    mpref_boot_free(h);
    BootSim *b = new BootSim();
    b->B = B; b->nseg = nseg; b->stride = h->P + 16;
    b->boot_samples_pars.resize(B);
    for (int s = 0; s < B; s++) {
        b->boot_samples_pars[s] = (unsigned short *)aligned32(sizeof(unsigned short) * (b->stride + 16));
        memcpy(b->boot_samples_pars[s], boot + (size_t)s * stride, sizeof(unsigned short) * (stride < h->P ? stride : h->P));
    }
    b->segment_upper.assign(seg_upper, seg_upper + nseg);
    b->use_skip = ras_pars_score != NULL && nseg > 1;
    if (b->use_skip) {                                   /* pllComputeRellRemainBound(nptn), iqtree.cpp:3821 */
        int nunit = h->params.sort_alignment ? h->aln.n_informative_patterns : h->P;
        std::vector<int> min_unit_pars(nunit);
        for (int i = 0; i < nunit; i++) {
            int pll_min = pllCalcMinParsScorePattern(h->tr, h->datatype, i);
            int cur_min = ras_pars_score[i];
            min_unit_pars[i] = (pll_min < cur_min) ? pll_min : cur_min;
        }
        b->remain.resize(B);
        for (int s = 0; s < B; s++) {
            b->remain[s].resize(nseg - 1);
            for (int g = 0; g < nseg - 1; g++) {
                int remain = 0;
                for (int pos = seg_upper[g]; pos < nunit; pos++) remain += min_unit_pars[pos] * b->boot_samples_pars[s][pos];
                b->remain[s][g] = remain;
            }
        }
    }
    b->boot_logl.assign(B, -(double)LONG_MAX);           /* iqtree.cpp:248 */
    b->boot_trees.assign(B, -1); b->boot_counts.assign(B, 0);
    b->logl_cutoff = logl_cutoff; b->eps = eps;
    b->pattern_pars = (unsigned short *)aligned32(sizeof(unsigned short) * (b->stride + 16));
    b->calls = b->reps_rows = b->skipped = b->bad_sum = 0;
    b->on_ratchet_hclimb1 = false; b->original_sample = NULL;
    b->multiple_hits = false; b->boot_trees_parsimony.assign(B, std::set<int>());
    b->store_top_boot_trees = 0; b->boot_trees_parsimony_top.assign(B, std::vector<std::pair<int, int> >());
    b->boot_threshold.assign(B, -INT_MAX);                    /* iqtree.cpp:267 */
    b->distinct_iter_top_boot = 0; b->curIt = 1;
    h->boot = b;
}

MPREF_API void mpref_boot_set_ratchet(mpref *h, const unsigned short *original_sample, const unsigned short *initial_ptn)
{
    BootSim *b = h->boot;
    b->on_ratchet_hclimb1 = true;
    free(b->original_sample);
    b->original_sample = (unsigned short *)aligned32(sizeof(unsigned short) * (b->stride + 16));
    memcpy(b->original_sample, original_sample, sizeof(unsigned short) * h->P);
    memset(b->pattern_pars, 0, sizeof(unsigned short) * (b->stride + 16));
    memcpy(b->pattern_pars, initial_ptn, sizeof(unsigned short) * h->P);
}

MPREF_API void mpref_boot_set_mulhits(mpref *h, int on) { h->boot->multiple_hits = on != 0; }
MPREF_API void mpref_boot_set_topboot(mpref *h, int n) { h->boot->multiple_hits = n > 0 || h->boot->multiple_hits; h->boot->store_top_boot_trees = n; }
MPREF_API void mpref_boot_set_distinct(mpref *h, int k, int cur_it)
{
    BootSim *b = h->boot;
    if (b->distinct_iter_top_boot != k) {                     /* iqtree.cpp:270-277 */
        b->distinct_iter_top_boot = k;
        b->boot_trees_parsimony_top.assign(b->B, std::vector<std::pair<int, int> >());
        b->boot_trees_parsimony_top_iter.assign(b->B, std::vector<int>());
        b->boot_threshold.assign(b->B, -INT_MAX);
    }
    b->curIt = cur_it;
}
MPREF_API int mpref_boot_topiters(mpref *h, int *flat, int cap)
{
    BootSim *b = h->boot;
    int tot = 0;
    for (size_t s = 0; s < b->boot_trees_parsimony_top_iter.size(); s++)
        for (size_t k = 0; k < b->boot_trees_parsimony_top_iter[s].size(); k++) { if (tot < cap) flat[tot] = b->boot_trees_parsimony_top_iter[s][k]; tot++; }
    return tot;
}
/* boot_trees_parsimony_top: sizes[B], thresholds[B], then (tree_index, rell) pairs in list order, concatenated; returns the pair count */
MPREF_API int mpref_boot_toplists(mpref *h, int *sizes, int *thresholds, int *flat, int cap)
{
    BootSim *b = h->boot;
    int tot = 0;
    for (int s = 0; s < b->B; s++) {
        sizes[s] = (int)b->boot_trees_parsimony_top[s].size();
        thresholds[s] = b->boot_threshold[s];
        for (size_t k = 0; k < b->boot_trees_parsimony_top[s].size(); k++) {
            if (tot < cap) { flat[2 * tot] = b->boot_trees_parsimony_top[s][k].first; flat[2 * tot + 1] = b->boot_trees_parsimony_top[s][k].second; }
            tot++;
        }
    }
    return tot;
}
/* boot_trees_parsimony: sizes[B] and the sets' members, ascending, concatenated; returns the total */
MPREF_API int mpref_boot_mulhits(mpref *h, int *sizes, int *flat, int cap)
{
    BootSim *b = h->boot;
    int tot = 0;
    for (int s = 0; s < b->B; s++) {
        sizes[s] = (int)b->boot_trees_parsimony[s].size();
        for (std::set<int>::iterator it = b->boot_trees_parsimony[s].begin(); it != b->boot_trees_parsimony[s].end(); ++it) {
            if (tot < cap) flat[tot] = *it;
            tot++;
        }
    }
    return tot;
}
MPREF_API void mpref_boot_set_cutoff(mpref *h, double c) { if (h->boot) h->boot->logl_cutoff = c; }
MPREF_API void mpref_boot_set_state(mpref *h, const double *bl, const int *bc, const int *bt)
{
    BootSim *b = h->boot;
    b->boot_logl.assign(bl, bl + b->B); b->boot_counts.assign(bc, bc + b->B); b->boot_trees.assign(bt, bt + b->B);
}
MPREF_API void mpref_boot_get_state(mpref *h, double *bl, int *bc, int *bt)
{
    BootSim *b = h->boot;
    memcpy(bl, &b->boot_logl[0], sizeof(double) * b->B);
    memcpy(bc, &b->boot_counts[0], sizeof(int) * b->B);
    memcpy(bt, &b->boot_trees[0], sizeof(int) * b->B);
}
MPREF_API long mpref_boot_counters(mpref *h, long *out4)
{
    BootSim *b = h->boot;
    out4[0] = b->calls; out4[1] = (long)b->treels_logl.size(); out4[2] = b->reps_rows; out4[3] = b->skipped;
    return b->bad_sum;
}
MPREF_API void mpref_boot_treels(mpref *h, double *out) { if (!h->boot->treels_logl.empty()) memcpy(out, &h->boot->treels_logl[0], sizeof(double) * h->boot->treels_logl.size()); }
MPREF_API int mpref_boot_nmat(mpref *h) { return (int)(h->boot->mat.size() / 5); }
MPREF_API void mpref_boot_mats(mpref *h, long long *out) { if (!h->boot->mat.empty()) memcpy(out, &h->boot->mat[0], sizeof(long long) * h->boot->mat.size()); }

static void boot_save_current_tree(mpref *h, double cur_logl)
{
    BootSim *b = h->boot;
    long call = b->calls++;
    if (b->on_ratchet_hclimb1) {                                                           /* :3283-3294 */
        int ptn = 0, segment_id = 0, score = 0;
        Vec16us vc_score = 0;
        for (; segment_id < b->nseg; segment_id++) {
            for (; ptn < b->segment_upper[segment_id]; ptn += 16)
                vc_score = Vec16us().load_a(&b->pattern_pars[ptn]) * Vec16us().load_a(&b->original_sample[ptn]) + vc_score;
            score += horizontal_add(vc_score);
            vc_score = 0;
        }
        cur_logl = -score;
    }
    if (b->logl_cutoff != 0.0 && cur_logl <= b->logl_cutoff - 1e-4) return;              /* :3343 */
    int tree_index = (int)b->treels_logl.size();                                          /* :3345 */
    b->treels_logl.push_back(cur_logl);
    int test_pars = 0;
    pllComputePatternParsimony(h->tr, h->pr, b->pattern_pars, &test_pars);                 /* :3365 */
    if (!b->on_ratchet_hclimb1 && test_pars != -int(cur_logl)) b->bad_sum++;               /* :3366 */
    b->reps_rows++;
    bool have_str = false;
    unsigned short *_pattern_pars = b->pattern_pars;
    int reps_segments = b->nseg;
    for (int sample = 0; sample < b->B; sample++) {                                        /* :3405 */
        double rell = 0.0;
        bool skipped = false;
        unsigned short *boot_sample = b->boot_samples_pars[sample];
        int ptn = 0, segment_id = 0, res = 0;
        Vec16us vc_rell = 0;
        for (; segment_id < reps_segments; segment_id++) {                                 /* :3424 */
            for (; ptn < b->segment_upper[segment_id]; ptn += 16)
                vc_rell = Vec16us().load_a(&_pattern_pars[ptn]) * Vec16us().load_a(&boot_sample[ptn]) + vc_rell;
            res += horizontal_add(vc_rell);
            vc_rell = 0;
            if ((!skipped) && b->use_skip && (reps_segments > 1) && (segment_id > reps_segments / 4) && (segment_id < reps_segments - 1)) {
                int reps_total = res + b->remain[sample][segment_id];
                if ((double)(-reps_total) < b->boot_logl[sample] - b->eps) { skipped = true; break; }
            }
        }
        rell = -(double)res;
        if (skipped) { b->skipped++; continue; }                                           /* :3484 */
        if (b->multiple_hits && b->store_top_boot_trees) {                                 /* :3536-3583 */
            std::vector<std::pair<int, int> > &top = b->boot_trees_parsimony_top[sample];
            const int N = b->store_top_boot_trees;
            if ((int)top.size() < N || rell > b->boot_threshold[sample]) {
                if (!have_str) {
                    have_str = true;
                    unsigned long long fp = mpref_tree_fingerprint(h);
                    std::map<unsigned long long, int>::iterator it = b->treels.find(fp);
                    if (it != b->treels.end()) tree_index = it->second;
                    else { tree_index = (int)b->treels_logl.size() - 1; b->treels[fp] = tree_index; }
                    long long m[5] = { call, 0, 0, tree_index, (long long)fp };
                    b->mat.insert(b->mat.end(), m, m + 5);
                }
                if (tree_index == (int)b->treels_logl.size() - 1) {                        /* newly added :3556 */
                    int count = (int)top.size();
                    if (count < N) {
                        std::vector<std::pair<int, int> >::iterator it;
                        for (it = top.begin(); it < top.end(); it++) if (it->second < rell) break;
                        top.insert(it, std::make_pair(tree_index, (int)rell));
                        b->boot_threshold[sample] = (b->boot_threshold[sample] < rell) ? b->boot_threshold[sample] : (int)rell;
                    } else if (count == N && rell > b->boot_threshold[sample]) {
                        top.pop_back();
                        std::vector<std::pair<int, int> >::iterator it;
                        for (it = top.begin(); it < top.end(); it++) if (it->second < rell) break;
                        top.insert(it, std::make_pair(tree_index, (int)rell));
                        b->boot_threshold[sample] = top[N - 1].second;
                    }
                }
            }
            continue;
        }
        if (b->multiple_hits) {                                                            /* :3498-3531 (no -topboot) */
            if (rell >= b->boot_logl[sample]) {
                if (!have_str) {
                    have_str = true;
                    unsigned long long fp = mpref_tree_fingerprint(h);
                    std::map<unsigned long long, int>::iterator it = b->treels.find(fp);
                    if (it != b->treels.end()) tree_index = it->second;
                    else { tree_index = (int)b->treels_logl.size() - 1; b->treels[fp] = tree_index; }
                    long long m[5] = { call, 0, 0, tree_index, (long long)fp };
                    b->mat.insert(b->mat.end(), m, m + 5);
                }
                if (rell > b->boot_logl[sample]) { b->boot_trees_parsimony[sample].clear(); b->boot_logl[sample] = rell; }
                if (b->boot_trees_parsimony[sample].find(tree_index) == b->boot_trees_parsimony[sample].end())
                    b->boot_trees_parsimony[sample].insert(tree_index);
            }
            continue;                                                                      /* neither :3587 nor :3687 applies */
        }
        if (b->distinct_iter_top_boot >= 1) {                                              /* :3587-3685 (!multiple_hits) */
            const int K = b->distinct_iter_top_boot;
            std::vector<std::pair<int, int> > &top = b->boot_trees_parsimony_top[sample];
            std::vector<int> &top_iter = b->boot_trees_parsimony_top_iter[sample];
            if (rell >= b->boot_threshold[sample]) b->boot_counts[sample]++;
            if (rell > b->boot_threshold[sample]
                || (rell == b->boot_threshold[sample] && random_double() <= (K * 1.0 / (b->boot_counts[sample])))) {
                if (rell > b->boot_logl[sample]) b->boot_counts[sample] = 1;
                if (!have_str) {
                    have_str = true;
                    unsigned long long fp = mpref_tree_fingerprint(h);
                    std::map<unsigned long long, int>::iterator it = b->treels.find(fp);
                    if (it != b->treels.end()) tree_index = it->second;
                    else { tree_index = (int)b->treels_logl.size() - 1; b->treels[fp] = tree_index; }
                    long long m[5] = { call, 0, 0, tree_index, (long long)fp };
                    b->mat.insert(b->mat.end(), m, m + 5);
                }
                b->boot_trees[sample] = tree_index;
                b->boot_logl[sample] = std::max(b->boot_logl[sample], rell);
                int t = std::min(K, (int)top_iter.size());
                int c;
                bool tree_exists = false;
                for (c = 0; t > 0 && c < t; c++) if (top[c].first == tree_index) { tree_exists = true; break; }
                if (tree_exists) continue;
                for (c = 0; t > 0 && c < t; c++) {                                         /* the iteration's representative */
                    if (top_iter[c] == b->curIt) {
                        if (rell > top[c].second) { top[c].second = rell; top[c].first = tree_index; }
                        break;
                    }
                }
                if (c == t & t < K) {                                                      /* room left: add */
                    top_iter.push_back(b->curIt);
                    top.push_back(std::make_pair(tree_index, rell));
                }
                if (c == t & t == K) {                                                     /* full: replace the worst */
                    int worst_id = 0;
                    int worst_score = top[worst_id].second;
                    for (int d = 1; d < t; d++) if (top[d].second < worst_score) { worst_score = top[d].second; worst_id = d; }
                    top[worst_id].first = tree_index; top[worst_id].second = rell; top_iter[worst_id] = b->curIt;
                }
                b->boot_threshold[sample] = top[0].second;
                for (size_t d = 1; d < top.size(); d++) if (top[d].second < b->boot_threshold[sample]) b->boot_threshold[sample] = top[d].second;
            }
            continue;
        }
        if (rell > b->boot_logl[sample] + b->eps
            || (rell > b->boot_logl[sample] - b->eps && random_double() <= 1.0 / (b->boot_counts[sample] + 1))) {   /* :3689 */
            if (!have_str) {                                                               /* :3692-3708 */
                have_str = true;
                unsigned long long fp = mpref_tree_fingerprint(h);
                std::map<unsigned long long, int>::iterator it = b->treels.find(fp);
                if (it != b->treels.end()) tree_index = it->second;
                else { tree_index = (int)b->treels_logl.size() - 1; b->treels[fp] = tree_index; }
                long long m[5] = { call, 0, 0, tree_index, (long long)fp };
                b->mat.insert(b->mat.end(), m, m + 5);
            }
            if (rell > b->boot_logl[sample]) b->boot_counts[sample] = 1;                   /* :3712 */
            b->boot_logl[sample] = std::max(b->boot_logl[sample], rell);                   /* :3720 */
            b->boot_trees[sample] = tree_index;
        }
        if (rell == b->boot_logl[sample]) b->boot_counts[sample]++;                        /* :3729 */
    }
}
