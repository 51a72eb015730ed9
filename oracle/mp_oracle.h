/* TEST INFRASTRUCTURE -- the checker, never the product.
 * Plain-C restatement of the reference's parsimony hot path (see mp_oracle.c).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it. */
#ifndef MP_ORACLE_H
#define MP_ORACLE_H
#include <stdint.h>

typedef struct mporacle mporacle;

mporacle *mporacle_create(int n, int P, int datatype, const uint8_t *codes, const int *weights, int sort_alignment);
void mporacle_destroy(mporacle *o);
void mporacle_set_weights(mporacle *o, const int *weights);
int  mporacle_allocate(mporacle *o, int per_site);
int  mporacle_num_informative(mporacle *o);
void mporacle_get_parsvect(mporacle *o, int node, uint32_t *out);
unsigned mporacle_node_score(mporacle *o, int node);
void mporacle_set_ring(mporacle *o, const int *back_node, const int *back_slot);
void mporacle_get_ring(mporacle *o, int *back_node, int *back_slot);
void mporacle_get_nodep(mporacle *o, int *refs);
void mporacle_node_rectifier(mporacle *o);
unsigned mporacle_evaluate_full(mporacle *o, int per_site);
void mporacle_pattern_parsimony(mporacle *o, uint16_t *out, int *sum);
int  mporacle_min_pars_pattern(mporacle *o, int site);
void mporacle_record(mporacle *o, int record_ptn);
int  mporacle_saved_count(mporacle *o);
void mporacle_saved_mp(mporacle *o, int *out);
void mporacle_saved_ptn(mporacle *o, uint16_t *out);
int  mporacle_rearrange(mporacle *o, int i, int mintrav, int maxtrav, int per_site, unsigned best_in, unsigned *out6);
void mporacle_apply_move(mporacle *o, int per_site);
int  mporacle_optimize_spr(mporacle *o, int mintrav, int maxtrav, int bb);
unsigned mporacle_ras(mporacle *o, long seed, int spr_dist);
unsigned long mporacle_sweep_count_insertions(mporacle *o, int mintrav, int maxtrav, int per_site, int reps);

void mporacle_saved_refs(mporacle *o, int *out);      /* (pruned ref, insertion ref) of every recorded saveCurrentTree call */
/* -mulhits */
void mporacle_boot_set_mulhits(mporacle *o, int on);
int  mporacle_boot_mulhits(mporacle *o, int *sizes, int *flat, int cap);
void mporacle_boot_set_topboot(mporacle *o, int n);
int  mporacle_boot_toplists(mporacle *o, int *sizes, int *thresholds, int *flat, int cap);
void mporacle_boot_set_distinct(mporacle *o, int k, int cur_it);   /* -distinct_iter_top_boot k, iqtree.cpp:3587-3685 */
int  mporacle_boot_topiters(mporacle *o, int *flat, int cap);
/* Sankoff (-cost) */
int  mporacle_set_cost_matrix(mporacle *o, const unsigned *cost, const int *segment_upper, int nseg);
int  mporacle_get_sankoff_vect(mporacle *o, int node, uint16_t *out);
void mporacle_set_best(mporacle *o, unsigned best);
int  mporacle_remainder_bounds(mporacle *o, unsigned *out);

double mporacle_random_double(void *unused);
void mporacle_seed_rng(uint64_t seed);
uint64_t mporacle_rng_draws(void);

uint32_t mporacle_code_mask(int datatype, int code);
int  mporacle_undetermined(int datatype);
void mporacle_reps(const uint16_t *pars, const uint16_t *boot, int B, int stride,
                   const int *segment_upper, int nseg, int *res_out);
int  mporacle_segments(const int *ras_score, const int *freq, int nptn, int n_informative, int *segment_upper);
#endif
