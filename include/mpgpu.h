/* mpgpu.h -- C-ABI of the B200-native parsimony hot path (libmpgpu.so).
 *
 * This is the drop-in boundary for MPBoot's data-parallel parsimony path.  The reference has
 * no FFI; its "operator API" is the set of free functions of sprparsimony.h:13-54 that work on
 * a pllInstance/partitionList pair plus the REPS block of IQTree::saveCurrentTree.  Every entry
 * point below names the reference interface it replaces (file:line under the reference tree).
 * INTEGRATION.md shows the shim a maintainer adds to sprparsimony.cpp / iqtree.cpp.
 *
 * Conventions
 *   - all functions return 0 on success, non-zero on error; mpgpu_last_error() gives the text
 *     (the host shim turns it into outError(), tools.cpp:91).  There is NO CPU fallback: if
 *     no CUDA device is usable every compute call fails.
 *   - the caller owns all host buffers; the context owns all device memory.  Calls are
 *     synchronous on return unless stated otherwise.
 *   - threading: like the reference's engine (file-scope globals, sprparsimony.cpp:127-141) a context is not re-entrant,
 *     and contexts on the SAME device must be driven from one host thread (under -cost they share one constant bank for
 *     the cost matrix, re-bound on alternating use).  One process (or thread) per GPU is the intended layout.
 *     The library itself starts up to three detached helper threads the first time a large batch (>= 96 node visits) is
 *     enumerated (mpgpu_host_plan_threads below; MPGPU_PLAN_THREADS=1: none); they touch host memory of the call only,
 *     never CUDA, and sleep between batches.  A forked child gets helpers of its own.
 *   - trees travel as PLL "ring tables": nodes 1..n are tips, n+1..2n-2 inner nodes; an inner
 *     node has ring slots 0,1,2 (slot s+1 = ->next of slot s, pllrepo/src/pll.h:687-702), a
 *     tip only slot 0.  back_node[3*i+s] / back_slot[3*i+s] = node number and slot hooked to
 *     slot s of node i (->back), 0 = NULL.  Both arrays have 3*(2n-1) entries.  A "ref"
 *     (pointer to one ring slot, i.e. a nodeptr) is encoded as 3*node+slot.
 *   - alignment data are PLL codes exactly as in tr->yVector (after pllBaseSubstitute,
 *     pllrepo/src/utils.c:2526) and pattern frequencies exactly as in tr->aliaswgt.
 */
#ifndef MPGPU_H
#define MPGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mpgpu_ctx mpgpu_ctx;

/* PLL data types (pllrepo/src/pll.h:238-245) accepted by mpgpu_load_alignment */
#define MPGPU_BINARY_DATA 0
#define MPGPU_DNA_DATA    1
#define MPGPU_AA_DATA     2
#define MPGPU_GENERIC_32  6

const char *mpgpu_last_error(void);

/* Number of usable CUDA devices (0 when none / driver missing). */
int mpgpu_device_count(void);

/* Create a context on CUDA device `device`.  `stream` is a cudaStream_t to launch on
 * (e.g. torch's current stream) or NULL for a private stream.  shard_rank/shard_count
 * select the contiguous word slice of every bit plane this context holds (pattern sharding,
 * SURVEY 8e); use 0/1 for a single GPU.  When shard_count > 1 the per-shard partial counts
 * returned by the *_partial calls must be summed over shards by the caller (NCCL
 * all-reduce of the int32 vector) before they are scores. */
int mpgpu_create(mpgpu_ctx **out, int device, void *stream, int shard_rank, int shard_count);
int mpgpu_destroy(mpgpu_ctx *ctx);
/* Pattern-sharded contexts (shard_count > 1): the one exchange step of the path is an
 * in-place int32 SUM all-reduce over the shards of a small device vector.  The host supplies it
 * (ncclAllReduce on its communicator, or torch.distributed): `fn` must enqueue the reduction of
 * dev_buf[0..count) on `stream` (the context's stream) and return 0.  With a callback installed
 * every call of this header returns complete results on every shard -- view lengths, scores,
 * per-insertion scores, pattern scores, replicate scores, and the SPR searches run replicated
 * (same host decisions on every rank, same RNG stream required).  Without one, sharded contexts
 * only offer the *_partial calls and the caller reduces. */
typedef int (*mpgpu_allreduce_fn)(void *user, void *dev_buf, int64_t count, void *stream);
int mpgpu_set_allreduce(mpgpu_ctx *ctx, mpgpu_allreduce_fn fn, void *user);
/* The same exchange step inside the library, on the device, without NCCL and without the host: a one-shot all-reduce
 * over NVLink peer memory (every shard stores its partial vector into every peer's exchange region, raises a flag, waits
 * for the peers' flags and sums; one small kernel on the context's stream, peer_exchange.cu).  For the shards of one
 * NVSwitch box (shard_count <= 8), one process per GPU:
 *   mpgpu_peer_prepare  allocates this shard's region for vectors of up to `capacity` int32 (longer ones go in pieces)
 *                       and returns its CUDA IPC handle (64 bytes) in handle_out;
 *   the host gathers the handles of all shards in shard order (any transport: torch.distributed all_gather, MPI, a file);
 *   mpgpu_peer_connect  maps the peers' regions.  From then on every call behaves as with an all-reduce callback
 *                       installed (complete results on every shard), and the callback, if any, is no longer used.
 * All shards must issue the same sequence of calls (they run the same replicated search anyway).  A shard whose peer
 * does not show up within 2 s records an error (mpgpu_peer_stats) instead of hanging the device. */
int mpgpu_peer_prepare(mpgpu_ctx *ctx, int64_t capacity, void *handle_out);
int mpgpu_peer_connect(mpgpu_ctx *ctx, const void *handles);
/* exchange steps issued / int32 elements reduced so far, and the device-side error flag (0 = none, 1 + q = shard q
 * did not arrive).  Any pointer may be NULL. */
int mpgpu_peer_stats(mpgpu_ctx *ctx, int64_t *calls, int64_t *elements, int *error);
/* The stream all kernels are launched on (cudaStream_t). */
void *mpgpu_stream(mpgpu_ctx *ctx);
int mpgpu_synchronize(mpgpu_ctx *ctx);
/* Kernels launched by this context since creation (bench.py's gpu_launches). */
int64_t mpgpu_launch_count(mpgpu_ctx *ctx);

/* ---- R1: _allocateParsimonyDataStructures / compressDNA (sprparsimony.cpp:3032, 2828) ----
 * yvector: [ntaxa][npatterns] PLL codes (tr->yVector[1..n]); aliaswgt: [npatterns]
 * (tr->aliaswgt).  sort_alignment mirrors Params::sort_alignment (isInformative returns TRUE
 * for every site when it is 0, sprparsimony.cpp:2462).  Builds the tip bit planes on the
 * device: pattern i expanded aliaswgt[i] times, padding bits set in every state. */
int mpgpu_load_alignment(mpgpu_ctx *ctx, int ntaxa, int npatterns, int datatype,
                         const uint8_t *yvector, const int32_t *aliaswgt, int sort_alignment);
/* _updateInternalPllOnRatchet + re-compress (sprparsimony.cpp:3022, 3249-3252): new pattern
 * frequencies over the same resident codes (ratchet / bootstrap-replicate re-weighting). */
int mpgpu_set_weights(mpgpu_ctx *ctx, const int32_t *aliaswgt);
/* states, reference parsimonyLength (words per plane padded to 8 as the AVX build,
 * sprparsimony.cpp:2870-2879), device words per plane of THIS shard, informative patterns
 * (numInformativePatterns :2862), expanded sites.  Any pointer may be NULL. */
int mpgpu_get_layout(mpgpu_ctx *ctx, int *states, int *ref_words, int *shard_words,
                     int *n_informative, int64_t *n_sites);
/* Tip planes in the reference layout parsVect[tip][state][ref_words] (tip in 1..n);
 * single-shard contexts only.  For parity tests of R1. */
int mpgpu_get_tip_planes(mpgpu_ctx *ctx, int tip, uint32_t *out);

/* ---- R2/R3: directed Fitch views of a tree (newviewParsimonyIterativeFast :554) ----
 * Uploads the topology and recomputes every directed view: for each ring slot (node,slot)
 * the Fitch state sets of the subtree that contains `node` when the edge at that slot is cut
 * (what parsVect[node] holds when xPars sits on that slot, pll.h:622-640). */
int mpgpu_set_tree(mpgpu_ctx *ctx, const int32_t *back_node, const int32_t *back_slot);
/* Mismatch count of every directed view of this shard (popcount of t_N, :773), indexed by
 * view id: tips 0..n-1 (always 0), inner (node,slot) -> n + 3*(node-n-1) + slot.
 * 4n-6 entries.  Sum over shards, then hand back with mpgpu_set_view_counts. */
int mpgpu_get_view_counts_partial(mpgpu_ctx *ctx, uint32_t *counts);
int mpgpu_set_view_counts(mpgpu_ctx *ctx, const uint32_t *counts);
/* tr->parsimonyScore[] equivalent: Fitch length of the subtree behind ref (node,slot). */
int mpgpu_view_length(mpgpu_ctx *ctx, int node, int slot, uint32_t *length);
/* The planes of one directed view in the reference layout [state][ref_words] (parity). */
int mpgpu_get_view_planes(mpgpu_ctx *ctx, int node, int slot, uint32_t *out);

/* ---- R4: evaluateParsimony(tr, pr, tr->start, full) (:1889, :965) ----
 * Tree length evaluated across the edge at tip 1 (tr->start).  Single shard: *score is the
 * score.  Sharded: *score is this shard's partial mismatch count of that edge and the caller
 * adds the all-reduced value to mpgpu_view_length of tip 1's neighbour. */
int mpgpu_tree_score(mpgpu_ctx *ctx, uint32_t *score);
/* Mismatch count across an arbitrary edge given by one of its refs (partial if sharded). */
int mpgpu_edge_mismatch_partial(mpgpu_ctx *ctx, int node, int slot, uint32_t *count);

/* ---- R5: pllComputePatternParsimony(ushort) (:3363) ----
 * Per-pattern Fitch score of the current tree, ptn_pars[0..upper) with upper =
 * numInformativePatterns (sort_alignment) or npatterns, and *sum = sum ptn_pars*aliaswgt.
 * Same site arithmetic as the reference (site += aliaswgt[ptn] over the prefix). */
int mpgpu_pattern_parsimony(mpgpu_ctx *ctx, uint16_t *ptn_pars, int32_t *sum);

/* ---- R6: SPR scan (rearrangeParsimony :2259, addTraverseParsimony :2208,
 *          testInsertParsimony :2106) ----
 * nodeRectifierPars (:2083): visit order of the sweep as refs, order[1..2n-2] (order[0]
 * unused), computed for the tree last given to mpgpu_set_tree. */
int mpgpu_visit_order(mpgpu_ctx *ctx, int32_t *order);
/* Scores, in one batched launch, every insertion that rearrangeParsimony would test for the
 * visits order[first..first+count) on the CURRENT tree.  Results are laid out per visit in
 * the reference's visit order: visit_begin[k]..visit_begin[k+1] index mp[]; within a visit
 * first the candidates of the p side then of the q side, each in addTraverseParsimony's
 * pre-order.  cand_ref[j] = ref of the insertion branch q (tr->insertNode if chosen),
 * cand_prune[j] = ref of the pruned node (tr->removeNode).  mp[j] is the full tree score
 * the reference computes at :2160.  visit_begin needs count+1 entries; the candidate arrays
 * `capacity` entries; *n_cand returns the number produced (error if capacity is too small).
 * Single shard only (sharded callers use the _partial variant below). */
int mpgpu_scan_visits(mpgpu_ctx *ctx, const int32_t *order, int first, int count,
                      int mintrav, int maxtrav,
                      int32_t *visit_begin, uint32_t *mp, int32_t *cand_ref, int32_t *cand_prune,
                      int capacity, int *n_cand);
/* Split form: plan on the host (identical on every rank), count on the device, finish after the
 * caller's all-reduce.  The device count vector has 2*count + n_cand int32 entries (one slot per
 * possible prune task of the planned visits, then one per candidate); n_tasks returns the tasks
 * actually planned. */
int mpgpu_scan_plan(mpgpu_ctx *ctx, const int32_t *order, int first, int count, int mintrav, int maxtrav,
                    int *n_cand, int *n_tasks);
/* Bytes of scan program (ops + tasks) the last mpgpu_scan_plan uploaded host->device. */
int64_t mpgpu_scan_plan_bytes(mpgpu_ctx *ctx);
/* Launches the scan of the planned batch; the int32 partial counts stay on the device at
 * *dev_counts (2*count + n_cand entries) for an in-place NCCL all-reduce.  Asynchronous. */
int mpgpu_scan_launch(mpgpu_ctx *ctx, void **dev_counts);
int mpgpu_scan_finish(mpgpu_ctx *ctx, int32_t *visit_begin, uint32_t *mp, int32_t *cand_ref,
                      int32_t *cand_prune, int capacity);

/* ---- pllOptimizeSprParsimony (:3244) ----
 * Hill-climbing SPR search on the tree given as ring tables (modified in place, like `tr`).
 * rng is the host's random_double() (tools.cpp:3362); it is called exactly where and as often
 * as the reference calls it.  Returns the reference's return value (startMP) in *best.
 * n_insertions (nullable) returns how many insertions were scored on the device. */
typedef double (*mpgpu_rng_fn)(void *user);
int mpgpu_optimize_spr(mpgpu_ctx *ctx, int32_t *back_node, int32_t *back_slot,
                       int mintrav, int maxtrav, mpgpu_rng_fn rng, void *rng_user,
                       uint32_t *best, int64_t *n_insertions);

/* About the last mpgpu_optimize_spr / _bb / stepwise search on this context: the score of the tree it started
 * from -- what the reference computes at sprparsimony.cpp:3277 and asserts equal to IQ-TREE's own kernel
 * (-iqtree->curScore, :3279), so the host can keep that cross-check --, the moves applied (:3312) and the
 * scan batches launched.  Any pointer may be NULL. */
int mpgpu_search_info(mpgpu_ctx *ctx, uint32_t *start_score, int64_t *moves, int64_t *batches);

/* ---- R12 / N1: the refinement loop of IQTree::optimizeBootTrees, default policy (iqtree.cpp:2795-2862) ----
 * For sample = 0 .. B-1, in order: the alignment is re-weighted with boot_samples[sample]
 * (Alignment::modifyPatternFreq, alignment.cpp:117 -- here: new frequencies over the codes that are
 * already resident, no alignment re-creation, no PHYLIP/newick round trip), the replicate's tree
 * (ring tables, trees_bn/trees_bs[sample][3*(2n-1)], modified in place) is hill-climbed with
 * pllOptimizeSprParsimony (doNNISearch -> :3244, radius maxtrav = params->spr_maxtrav or
 * opt_btree_spr), and scores[sample] receives its return value.  rng is the host's random_double();
 * the draws of consecutive samples are consumed in order, exactly as the reference's sequential
 * loop does, so the refined trees and scores are the reference's.  The caller maps the trees
 * to treels indices (:2847-2858).  The original frequencies are restored before returning. */
int mpgpu_refine_replicates(mpgpu_ctx *ctx, int B, const uint16_t *boot_samples, int stride,
                            int32_t *trees_bn, int32_t *trees_bs, int mintrav, int maxtrav,
                            mpgpu_rng_fn rng, void *rng_user, uint32_t *scores, int64_t *n_insertions);

/* ---- R7: _pllComputeRandomizedStepwiseAdditionParsimonyTree (sprparsimony.cpp:3224, 3107, 2977) ----
 * Builds a randomized-stepwise-addition tree over the loaded alignment and runs the SPR rounds
 * of _pllMakeParsimonyTreeFast on it (radius spr_dist).  *random_seed is tr->randomNumberSeed
 * (PLL's private generator for the taxon order, pllrepo/src/utils.c:335; updated like the
 * reference does); rng is the host's random_double() for the tie-breaks (:3004, :3199).  The tree
 * comes back as ring tables with the reference's node numbering (inner nodes n+1.. in insertion
 * order), *best = tr->bestParsimony. */
int mpgpu_stepwise_addition(mpgpu_ctx *ctx, int64_t *random_seed, int spr_dist, mpgpu_rng_fn rng, void *rng_user,
                            int32_t *back_node, int32_t *back_slot, uint32_t *best, int64_t *n_insertions);

/* Replicate shards (multi-GPU -bb): every context holds the WHOLE alignment (shard_count = 1 at mpgpu_create) and scores all
 * candidates itself, but keeps only its share [B*rank/count, B*(rank+1)/count) of the replicates, so the replicate contraction --
 * the dominant cost under -bb -- divides by the number of GPUs.  Call before mpgpu_peer_prepare / mpgpu_set_allreduce and before
 * mpgpu_load_replicates (which is still given the full B x stride table on every rank).  mpgpu_reps_* and the -bb searches keep
 * their single-GPU meaning on every rank: the per-call hit flags are summed over the group and only the rows of calls that can
 * change some replicate are assembled to full width through the exchange step, so every rank replays the same bookkeeping
 * (iqtree.cpp:3405-3449 sees B replicates).  Also the way to run -cost -bb on several GPUs (pattern shards refuse it). */
int mpgpu_set_replicate_shards(mpgpu_ctx *ctx, int rank, int count);

/* ---- R8: replicate scoring, the REPS block of IQTree::saveCurrentTree (iqtree.cpp:3356-3449) ----
 * boot_samples: [B][stride] u16, boot_samples_pars exactly as IQTree::setParams fills it
 * (iqtree.cpp:220-233, 285-313; stride >= number of reported patterns); segment_upper/nseg
 * from IQTree::doSegmenting (iqtree.cpp:3793).  The weights stay resident on the device.
 * Results are the reference's integers, including the 16-bit wrap of each segment sum
 * (vectorclass/vectori256.h:1726-1740): segments that provably cannot reach 2^16 on any tree
 * and weights <= 255 run on the int8 tensor cores, everything else through an exact
 * CUDA-core path (mpgpu_reps_info reports the split). */
int mpgpu_load_replicates(mpgpu_ctx *ctx, int B, const uint16_t *boot_samples, int stride,
                          const int32_t *segment_upper, int nseg);
/* The same plus IQTree::original_sample (u16, the ORIGINAL pattern frequencies, at least as many
 * entries as reported patterns): needed for ratchet iterations, where saveCurrentTree re-scores
 * cur_logl on the original frequencies (iqtree.cpp:3283-3294).  It rides along as one more
 * column of the replicate contraction. */
int mpgpu_load_replicates2(mpgpu_ctx *ctx, int B, const uint16_t *boot_samples, int stride,
                           const int32_t *segment_upper, int nseg, const uint16_t *original_sample);
/* groups = 1 + wrap-prone segments, exceptions = patterns on the exact CUDA-core path,
 * tensor = 1 when the tcgen05 path is enabled.  Any pointer may be NULL. */
int mpgpu_reps_info(mpgpu_ctx *ctx, int *groups, int *exceptions, int *tensor);
/* ---- R11: -cost, Sankoff weighted parsimony --------------------------------------------------------
 * Replaces the Sankoff half of the engine: compressSankoffDNA (sprparsimony.cpp:2637-2826),
 * newviewSankoffParsimonyIterativeFastSIMD (:477-551), evaluateSankoffParsimonyIterativeFastSIMD
 * (:880-961), pllComputeSankoffPatternParsimony (:3346-3360), the remainder lower bounds built from
 * ParsTree::findMstScore (parstree.cpp:606; sprparsimony.cpp:2801-2823).  The reference switches on the
 * global pllCostMatrix (:556, :967, :2830; set in iqtree.cpp:605); here the switch is this call.
 *
 * cost = [nstates][nstates] u32 (ParsTree::cost_matrix after loadCostMatrixFile, parstree.cpp:31-90),
 * segment_upper/nseg = pllSegmentUpper / pllRepsSegments (IQTree::doSegmenting, iqtree.cpp:3793): interior
 * bounds multiples of 16, the last one = the number of informative patterns.  cost = NULL returns to Fitch.
 * After this call every entry point above (tree score, view lengths = tr->parsimonyScore[], pattern
 * parsimony, scan, SPR search, stepwise addition) computes weighted parsimony with the reference's
 * integers, including the per-segment 16-bit wrap of the weighted sums (:944-948).
 * Precondition checked here (error otherwise, there is no fallback): (ntaxa+1)*(max cost+1) <= 65535 -- then no u16 of
 * the reference wraps inside a vector.  The matrix may be asymmetric (the reference only repairs the triangle inequality,
 * parstree.cpp:31-90): scores then depend on where the tree is rooted, and every entry point takes the reference's own rooting
 * (an insertion at the node above the insertion point, :2160; a stepwise insertion at the new tip, :2994-2998; the tree at
 * tr->start's neighbour; under -bb the current tree's vector at the edge of every node visit in turn, :2286-2289) -- one more
 * min-plus per scored insertion than for a symmetric matrix.  Sharded contexts: install
 * mpgpu_set_allreduce first (the shards hold ranges of pattern pairs; per-segment sums are reduced before the
 * 16-bit masks).  -cost with -bb: see mpgpu_sankoff_reps_stats below (unsharded contexts only). */
int mpgpu_set_cost_matrix(mpgpu_ctx *ctx, const uint32_t *cost, int nstates, const int32_t *segment_upper, int nseg,
                          uint32_t *highest_cost);
/* vector_length = the reference's per-node vector length in patterns (informative patterns padded to 16,
 * :2664-2667); remainder_bounds = pllRemainderLowerBounds[nseg-1].  Pointers may be NULL. */
int mpgpu_sankoff_layout(mpgpu_ctx *ctx, int *vector_length, int *n_bounds, uint32_t *remainder_bounds, int capacity);
/* The cost vector of the subtree behind ring slot (node, slot) in the reference's layout u16
 * [vector_length][nstates] (parsVect of that node when xPars sits on the slot). */
int mpgpu_sankoff_view(mpgpu_ctx *ctx, int node, int slot, uint16_t *out);
/* Early termination (:951-956).  mpgpu_scan_* always return the exact score of every insertion (what the
 * reference computes with perSiteScores = 1).  In plain mode the reference leaves evaluateSankoff... as
 * soon as a prefix of segment sums plus the remainder bound exceeds tr->bestParsimony and returns that
 * estimate instead; est_max[j] = max over segments of (prefix + bound) for candidate j of the last scan, so
 * "est_max[j] > bestParsimony" <=> the reference exits early for j (its return value is then > best and
 * changes nothing).  mpgpu_optimize_spr / mpgpu_stepwise_addition replay exactly this; option
 * "sankoff_exact" = 1 switches the replay off (the reference's perSiteScores mode). */
int mpgpu_scan_bounds(mpgpu_ctx *ctx, uint32_t *est_max, int capacity);
/* -cost together with -bb: set the cost matrix first, then mpgpu_load_replicates; mpgpu_reps_* and
 * mpgpu_optimize_spr_bb (every policy) then score the Sankoff pattern vectors (pllComputeSankoffPatternParsimony :3346)
 * of the current tree / of every insertion with the u16 semantics of iqtree.cpp:3424-3449.  Per chunk of calls the
 * contraction runs on the int8 tensor cores when every cost and replicate weight fits in a byte and no 16-bit segment
 * sum of the chunk can wrap (proved from the chunk's column maxima), and on an exact CUDA-core kernel otherwise;
 * the counters say how many chunks took which. */
int mpgpu_sankoff_reps_stats(mpgpu_ctx *ctx, int64_t *tensor_chunks, int64_t *exact_chunks);

/* Options: "sankoff_exact" 0/1 (see mpgpu_scan_bounds);
 * "reps_tensor" 0/1 (0 = everything through the exact CUDA-core kernel; for tests);
 * "reps_nowrap" 0/1 (-autovec, iqtree.cpp:3418-3423: the replicate scores are plain int dot products, no 16-bit segment sums and no
 * skip test; the original_sample column of ratchet iterations keeps the segmented u16 sums of :3283-3294; Fitch scoring only);
 * "reps_timing" 0/1 (CUDA events around the largest tensor-kernel launch, read by mpgpu_reps_timing);
 * "sankoff_u32" 0/1 (-short_off, tools.cpp:2365: the reference's 32-bit Sankoff vectors -- pattern weights are not cut to 16 bits and
 * the per-segment weighted sums and node scores do not wrap at 2^16; set before mpgpu_set_cost_matrix; the u16 bound on
 * (ntaxa+1)*(max cost+1) still applies, within it the vectors themselves are the same numbers);
 * "exchange" 1/0 (sharded contexts: 0 skips the exchange step, so every result stays this shard's partial -- for timing the
 * kernels without it). */
int mpgpu_set_option(mpgpu_ctx *ctx, const char *name, int value);
/* Device time (ms, CUDA events on the context's stream) of the largest tensor-kernel launch since the
 * last call, with its shape: rows x patterns (K, padded to 128) x replicates, and the K splits used. */
int mpgpu_reps_timing(mpgpu_ctx *ctx, float *tc_ms, int *rows, int *patterns, int *splits);
/* Measurement aid (no reference counterpart): the issue rate of tcgen05.mma kind::i8 on this device, in int8 TOP/s -- one
 * CTA per SM issuing M = 128, N = 256, K = 32 MMAs back to back from resident shared-memory tiles, `iters` x 64 per CTA.
 * It is the measured denominator bench.py reports k_reps_tc's roofline against (MEASURED_PEAKS.json has no int8 figure). */
int mpgpu_int8_peak(mpgpu_ctx *ctx, int iters, double *tops);
/* res[b] = -rell[b] of the CURRENT tree for b < B (what the loop at :3424-3449 leaves in `res`
 * when no replicate is skipped).  Single shard. */
int mpgpu_reps_current_tree(mpgpu_ctx *ctx, int32_t *res);
/* The same for candidates of the last mpgpu_scan_visits / mpgpu_scan_plan batch: cand_idx[i]
 * indexes mp[] of that batch (-1 = the current tree as evaluated at tr->start; -(2 + v) = the current tree as the batch's v-th
 * node visit evaluates it, at its own edge, sprparsimony.cpp:2286 -- the same vector unless the cost matrix is asymmetric);
 * res is [m][B].  This is the REPS vector
 * saveCurrentTree would compute inside testInsertParsimony (:2163-2166) for that insertion.
 * Single shard. */
int mpgpu_reps_candidates(mpgpu_ctx *ctx, const int32_t *cand_idx, int m, int32_t *res);
/* Asynchronous form: the vectors stay on the device, *dev_res = int32 [m][*pitch] (valid until
 * the next REPS call on this context, ordered on the context's stream).  The batch must fit the
 * row buffers in one piece. */
int mpgpu_reps_candidates_device(mpgpu_ctx *ctx, const int32_t *cand_idx, int m, void **dev_res, int *pitch);

/* ---- pllOptimizeSprParsimony under -bb: search + saveCurrentTree ----
 * Replaces the pair pllOptimizeSprParsimony (sprparsimony.cpp:3244) / IQTree::saveCurrentTree
 * (iqtree.cpp:3271-3760) for maximum_parsimony && spr_parsimony with !store_candidate_trees.  Described here for the
 * default policy (!multiple_hits, distinct_iter_top_boot < 1, outside ratchet iterations); the other policies and
 * ratchet iterations are selected through mpgpu_bb_state below.  Every scored
 * insertion and, once per node visit, the current tree (:2286-2289) goes through the cutoff
 * filter (:3343), is appended to treels_logl (push_tree_logl), gets its REPS vector from the
 * device and updates boot_logl / boot_counts / boot_trees exactly as :3687-3731, drawing
 * random_double() for ties in the reference's order.  When a candidate wins its first
 * replicate the host materialises it (`materialize`, :3692-3708): it receives the current
 * ring tables and the move (remove_ref regrafted on insert_ref; 0,0 = the current tree)
 * and returns the tree index to store in boot_trees (the treels lookup is the host's). */
typedef struct mpgpu_bb_hooks {
    void *user;
    double (*random_double)(void *user);
    int32_t (*push_tree_logl)(void *user, double cur_logl);
    int32_t (*materialize)(void *user, const int32_t *back_node, const int32_t *back_slot,
                           int32_t remove_ref, int32_t insert_ref, int32_t tree_index);
    /* policy MPGPU_BB_MULHITS only (may be NULL otherwise): replicate `sample` scored rell >= boot_logl[sample]
     * with tree `tree_index`; clear_first != 0 <=> rell > boot_logl[sample], i.e. boot_trees_parsimony[sample]
     * .clear() comes first (iqtree.cpp:3517-3520), then insert-if-absent (:3531-3534). */
    void (*mulhit)(void *user, int32_t sample, int32_t tree_index, int32_t clear_first);
    /* policy MPGPU_BB_MULHITS_TOP only (may be NULL otherwise): a newly seen tree enters boot_trees_parsimony_top[sample]
     * (kept in decreasing rell order: insert before the first entry with a smaller rell, iqtree.cpp:3563-3568);
     * pop_worst != 0 <=> the list is full and its last entry is dropped first (:3571).  Returns the rell of the list's
     * last entry after the update (the new boot_threshold when the list was full, :3578). */
    int32_t (*tophit)(void *user, int32_t sample, int32_t tree_index, int32_t rell, int32_t pop_worst);
    /* policy MPGPU_BB_DISTINCT_ITER only (may be NULL otherwise): replicate `sample` accepted tree `tree_index` with score rell in
     * iteration cur_it; the host updates boot_trees_parsimony_top[sample] / boot_trees_parsimony_top_iter[sample] as
     * iqtree.cpp:3624-3669 (nothing if the tree is in the list; else the entry of this iteration is replaced when rell is better;
     * else appended while the list has fewer than top_n entries; else the worst entry is replaced) and returns the new
     * boot_threshold[sample]: the smallest rell of the list (:3671-3677), or `threshold` unchanged when the tree was in the list
     * (the `continue` at :3634). */
    int32_t (*disthit)(void *user, int32_t sample, int32_t tree_index, int32_t rell, int32_t cur_it, int32_t top_n, int32_t threshold);
} mpgpu_bb_hooks;
#define MPGPU_BB_DEFAULT 0     /* iqtree.cpp:3687-3731 */
#define MPGPU_BB_MULHITS_TOP 2 /* -mulhits -topboot N (store_top_boot_trees, iqtree.cpp:3536-3583): per replicate the N best newly
                                * seen trees; state->top_n = N, state->top_count / state->boot_threshold [B] in/out; boot_logl,
                                * boot_counts, boot_trees untouched, no tie-break draw */
#define MPGPU_BB_MULHITS 1     /* params->multiple_hits without -topboot, iqtree.cpp:3498-3531: every tree that ties a
                                * replicate's best score is kept; no tie-break draw, boot_counts / boot_trees untouched */
#define MPGPU_BB_DISTINCT_ITER 3 /* -distinct_iter_top_boot K (iqtree.cpp:3587-3685): per replicate up to K trees from distinct
                                 * iterations; a tree is accepted when rell > boot_threshold, or rell == boot_threshold and
                                 * random_double() <= K / boot_counts (boot_counts counts the calls with rell >= boot_threshold,
                                 * reset to 1 by a new best); state->top_n = K, state->cur_it = IQTree::curIt, boot_threshold,
                                 * boot_logl, boot_counts, boot_trees in/out.  Under this policy the remain-bound skip of the REPS
                                 * loop (:3433-3445) changes decisions (it compares with boot_logl, acceptance with boot_threshold):
                                 * it is replayed exactly when the bounds were given with mpgpu_set_remain_bounds.  Fitch scoring,
                                 * unsharded contexts. */
typedef struct mpgpu_bb_state {
    int32_t B;
    double *boot_logl;        /* [B] in/out */
    int32_t *boot_counts;     /* [B] in/out */
    int32_t *boot_trees;      /* [B] in/out */
    double logl_cutoff;       /* IQTree::logl_cutoff (0.0 = no filter) */
    double ufboot_epsilon;    /* Params::ufboot_epsilon (0.5) */
    int64_t n_calls;          /* out: saveCurrentTree calls */
    int64_t n_reps;           /* out: calls that passed the cutoff (REPS vectors computed and used) */
    /* Ratchet iteration (IQTree::on_ratchet_hclimb1, iqtree.cpp:3283-3294): the search runs on the
     * perturbed frequencies (mpgpu_set_weights) and every call's cur_logl is minus the score, on the
     * ORIGINAL frequencies (mpgpu_load_replicates2), of whatever _pattern_pars holds on entry -- the
     * vector of the previous call that reached pllComputePatternParsimony.  ratchet_pattern_pars =
     * the host's _pattern_pars before the search (u16, reported patterns); ratchet_last_score
     * returns the score the next call would see. */
    int32_t ratchet;                          /* 0 = normal iteration */
    const uint16_t *ratchet_pattern_pars;     /* in, when ratchet */
    int32_t ratchet_last_score;               /* out */
    int32_t policy;                           /* MPGPU_BB_DEFAULT / MPGPU_BB_MULHITS / MPGPU_BB_MULHITS_TOP */
    int32_t top_n;                            /* MPGPU_BB_MULHITS_TOP: params->store_top_boot_trees */
    int32_t *top_count;                       /* [B] in/out: boot_trees_parsimony_top[sample].size() */
    int32_t *boot_threshold;                  /* [B] in/out: IQTree::boot_threshold (vector<int>, starts at -INT_MAX, iqtree.cpp:267) */
    int32_t cur_it;                           /* MPGPU_BB_DISTINCT_ITER: IQTree::curIt of this search */
    int32_t updates_off;                      /* -min_iter1_cand in iteration 1 (params->minimize_iter1_candidates && curIt == 1,
                                               * iqtree.cpp:3404): a call that passes the cutoff is appended to treels_logl and
                                               * nothing else happens -- no replicate is touched, no tree materialised, no draw */
    int32_t *boot_tree_orig_logl;             /* -cutoff_from_btrees (IQTree::boot_tree_orig_logl, vector<int> [B]) or NULL: the
                                               * call's cur_logl is stored when a replicate accepts its tree (default :3717-3718,
                                               * distinct-iteration :3618-3619) or, under -mulhits, raised to it (:3524-3527);
                                               * the host derives logl_cutoff from it between searches (:1657-1661) */
} mpgpu_bb_state;
int mpgpu_optimize_spr_bb(mpgpu_ctx *ctx, int32_t *back_node, int32_t *back_slot, int mintrav, int maxtrav,
                          const mpgpu_bb_hooks *hooks, mpgpu_bb_state *state,
                          uint32_t *best, int64_t *n_insertions);

/* boot_samples_pars_remain_bounds (iqtree.cpp:3821-3858, IQTree::pllComputeRellRemainBound): bounds[b][s], s < nseg - 1 = a lower
 * bound of replicate b's score over the patterns from segment_upper[s] on; per_replicate must be nseg - 1.  After
 * mpgpu_load_replicates (which drops earlier bounds); NULL clears them.  Only policy MPGPU_BB_DISTINCT_ITER reads them: under the
 * other policies the skip they allow is decision-neutral and the device scores every replicate anyway. */
int mpgpu_set_remain_bounds(mpgpu_ctx *ctx, const int32_t *bounds, int per_replicate);
/* What the skip test of iqtree.cpp:3433-3445 compares with boot_logl, for candidate cand_idx of the last scan batch (-1 = the
 * current tree) and the replicates samples[0..m): out[i] = max over the segments nseg/4 < s < nseg-1 of (sum of the 16-bit
 * segment sums up to s + bounds[samples[i]][s]); INT32_MIN when no segment is in that range.  The reference skips the replicate
 * <=> -(out[i]) < boot_logl - ufboot_epsilon. */
int mpgpu_reps_prefix_max(mpgpu_ctx *ctx, int32_t cand_idx, const int32_t *samples, int m, int32_t *out);

/* ---- N2: support summarisation, the split table of a weighted collection of trees -----------------------------------
 * Replaces the arithmetic of MTreeSet::convertSplits (mtreeset.cpp:362-440) as IQTree::summarizeBootstrap drives it
 * (iqtree.cpp:3872-3989: SW_COUNT, no threshold): every tree's bipartitions in the order MTree::convertSplits pushes them
 * (mtree.cpp:917-939, post-order by edge), normalised like Split::shouldInvert (split.cpp:100-107), merged over the
 * trees with their weights summed, distinct splits in first-seen order.
 * A tree is the reverse-Polish token stream of that traversal from its root leaf's neighbour: token t >= 0 = leaf with
 * taxon id t (pushes {t}, emits it), token -k (k >= 2) = inner node joining the k topmost subtrees (emits their union);
 * the last token emits the edge to the root leaf.  tokens of tree i: [token_begin[i], token_begin[i+1]); tree_weight[i] as
 * MTreeSet::tree_weights (trees of weight 0 are scanned too -- leave them out, or put a tree whose supports are wanted
 * last with weight 0 and read its splits' rows through emit_unique).
 * Out: *n_unique distinct splits; split_bits [n_unique][(ntaxa+31)/32] (taxon i = bit i%32 of word i/32, as class Split),
 * split_weight [n_unique] = summed tree weights, both in first-seen order; emit_unique [number of tokens] = row of every
 * emitted split.  Any of the three may be NULL; capacity = rows available in split_bits / split_weight.
 * Needs a context only for its device and stream (no alignment). */
int mpgpu_split_table(mpgpu_ctx *ctx, int ntaxa, int ntrees, const int32_t *tokens, const int64_t *token_begin,
                      const int32_t *tree_weight, int32_t *n_unique, uint32_t *split_bits, int32_t *split_weight,
                      int32_t *emit_unique, int capacity);

/* ---- host-only entry points: no device, no CUDA call ----
 * The tree-walking half of the path on ring tables, for hosts that keep their own search loop and for tests of the host
 * logic: nodeRectifierPars' visit order (sprparsimony.cpp:2046-2101), the candidates rearrangeParsimony /
 * addTraverseParsimony test for visits order[first .. first+count) in the reference's order (:2208-2376) -- the same
 * enumeration mpgpu_scan_* score -- and restoreTreeRearrangeParsimony's move (:2379). */
int mpgpu_host_visit_order(int ntaxa, const int32_t *back_node, const int32_t *back_slot, int32_t *order);
int mpgpu_host_enumerate(int ntaxa, const int32_t *back_node, const int32_t *back_slot, const int32_t *order, int first, int count,
                         int mintrav, int maxtrav, int32_t *visit_begin, int32_t *cand_ref, int32_t *cand_prune, int capacity,
                         int *n_cand);
int mpgpu_host_apply_spr(int ntaxa, int32_t *back_node, int32_t *back_slot, int32_t remove_ref, int32_t insert_ref);
/* Host threads the library uses to enumerate a large batch (>= 96 node visits in one piece, outside SPR searches): the caller's plus
 * up to three detached helpers, started on first use, that sleep between batches; environment MPGPU_PLAN_THREADS=<n> (1 = none),
 * default 4 on machines with 8 or more hardware threads per process of the node (LOCAL_WORLD_SIZE, as torchrun exports it), 2 with 4
 * or more, else 1.  The result does not depend on it. */
int mpgpu_host_plan_threads(void);
/* Test aid: the scan program of those visits built once by one thread and once by `nthreads` host threads in `pieces` pieces (the
 * way mpgpu_scan_visits builds a large batch); 0 <=> tasks, view offsets, control words, candidate and visit tables are the same
 * bytes. */
int mpgpu_host_plan_selftest(int ntaxa, const int32_t *back_node, const int32_t *back_slot, const int32_t *order, int first, int count,
                             int mintrav, int maxtrav, int nthreads, int pieces);

/* ---- host-side helper: a minimal treels / treels_logl container (iqtree.h) ----
 * For hosts that do not bring their own (tests, bench.py): collects treels_logl, keys
 * materialised trees by a canonical topology hash like the treels map (iqtree.cpp:3299-3312,
 * 3701-3708), forwards random_double to `rng`.  mpgpu_treels_materialized returns
 * [k][4] = remove_ref, insert_ref, tree_index, topology hash per materialised tree. */
/* random_double() for hosts without their own stream: splitmix64 on the uint64_t `user` points to
 * (an mpgpu_rng_fn). */
double mpgpu_splitmix64_double(void *user);
typedef struct mpgpu_treels mpgpu_treels;
mpgpu_treels *mpgpu_treels_create(int ntaxa);
void mpgpu_treels_destroy(mpgpu_treels *t);
int64_t mpgpu_treels_size(const mpgpu_treels *t);
void mpgpu_treels_logl(const mpgpu_treels *t, double *out);
int64_t mpgpu_treels_num_materialized(const mpgpu_treels *t);
void mpgpu_treels_materialized(const mpgpu_treels *t, int64_t *out);
void mpgpu_treels_hooks(mpgpu_treels *t, mpgpu_rng_fn rng, void *rng_user, mpgpu_bb_hooks *out);
/* -mulhits: boot_trees_parsimony as collected through the mulhit hook: sizes[nsamples], then the members of
 * every set in ascending order, concatenated into flat (up to capacity); returns the total count. */
int64_t mpgpu_treels_mulhits(const mpgpu_treels *t, int32_t nsamples, int32_t *sizes, int32_t *flat, int64_t capacity);
/* -mulhits -topboot: boot_trees_parsimony_top as collected through the tophit hook: sizes[nsamples], then (tree_index, rell)
 * pairs in list order, concatenated into flat (2 ints per pair, up to capacity pairs); returns the pair count. */
int64_t mpgpu_treels_toplists(const mpgpu_treels *t, int32_t nsamples, int32_t *sizes, int32_t *flat, int64_t capacity);
/* -distinct_iter_top_boot: the lists above are then filled by the disthit hook (list order = insertion order); this returns
 * boot_trees_parsimony_top_iter, the iteration of every entry in the same order (up to capacity ints); returns the count. */
int64_t mpgpu_treels_topiters(const mpgpu_treels *t, int32_t nsamples, int32_t *flat, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* MPGPU_H */
