/* TEST INFRASTRUCTURE (oracle/_ref build only) -- not product code.
 *
 * The reference's parsimony engine (/root/reference/sprparsimony.cpp) reaches
 * into class IQTree for a handful of fields and one virtual up-call
 * (sprparsimony.cpp:438, 2807, 3026, 3249-3253, 3279, 3323, 3396-3400).
 * Linking the real IQTree would pull in the whole ~150-file program, so the
 * reference driver pre-defines the include guards of iqtree.h / parstree.h
 * (IQPTREE_H, PARSTREE_H_) and supplies this minimal stand-in with the same
 * member names.  sprparsimony.cpp itself is compiled UNMODIFIED from where it
 * lies under /root/reference.  The members that only the diagnostic tail of
 * that file touches (computeUserTreeParsimomy, convertNewickTo*) are declared
 * but never defined; those functions are garbage-collected at link time.
 */
#ifndef MPB200_IQTREE_SHIM_H
#define MPB200_IQTREE_SHIM_H

#define IQPTREE_H      /* guard of /root/reference/iqtree.h   */
#define PARSTREE_H_    /* guard of /root/reference/parstree.h */

#include <vector>
#include <string>
#include <sstream>
#include <fstream>
#include <iostream>
#include <algorithm>
#include <cassert>
#include <climits>
#include <cstring>

#include "tools.h"               /* real reference header: Params, random_double(), outError() */
#include "pllrepo/src/pll.h"     /* real reference header */
extern "C" {
#include "pllrepo/src/pllInternal.h"   /* upstream reaches it through an extern "C" wrapper (phylolib.h:17) */
}

using namespace std;

/* pattern.h:24-64 -- only the fields the parsimony engine reads */
struct Pattern {
    int  frequency;
    bool is_const;
    int  ras_pars_score;
};

/* alignment.h:45, :609 */
class Alignment : public std::vector<Pattern> {
public:
    Alignment() : n_informative_patterns(0) {}
    Alignment(char *aln_file, char *sequence_type, InputType &intype); /* never defined */
    int    getNSeq();                 /* never defined */
    string getSeqName(int i);         /* never defined */
    int n_informative_patterns;
};

typedef void (*mpref_save_hook_t)(void *user, double cur_logl);

/* iqtree.h -- members used by sprparsimony.cpp */
class IQTree {
public:
    IQTree() : aln(NULL), pllInst(NULL), pllPartitions(NULL), curScore(0.0), logl_cutoff(0.0),
               on_ratchet_hclimb1(false), on_ratchet_hclimb2(false), on_opt_btree(false),
               hook(NULL), hook_user(NULL) {}
    IQTree(Alignment *a);             /* never defined */
    virtual ~IQTree() {}

    /* iqtree.h:516, iqtree.cpp:3271 -- the -bb up-call; here it forwards to the driver */
    virtual void saveCurrentTree(double cur_logl) { if (hook) hook(hook_user, cur_logl); }

    /* only reached from pllComputePatternParsimonySlow / the diagnostic tail */
    virtual void initializeAllPartialPars() {}
    virtual void clearAllPartialLH() {}
    virtual int  computeParsimony() { return 0; }
    void   readTree(const char *file, bool &is_rooted);      /* never defined */
    void   setAlignment(Alignment *a);                        /* never defined */
    void   initializePLL(Params &params);                     /* never defined */
    string getTreeString();                                   /* never defined */
    void   printTree(ostream &out, int brtype);               /* never defined */

    Alignment     *aln;
    pllInstance   *pllInst;
    partitionList *pllPartitions;
    double curScore;
    double logl_cutoff;
    bool   on_ratchet_hclimb1, on_ratchet_hclimb2, on_opt_btree;

    mpref_save_hook_t hook;
    void *hook_user;
};

class ParsTree : public IQTree {
public:
    ParsTree(Alignment *a);           /* never defined */
    void initParsData(Params *p);     /* never defined */
    int  findMstScore(int ptn);       /* parstree.cpp:606, Sankoff only, never defined */
};

void optimizeAlignment(IQTree *&tree, Params &params);        /* phyloanalysis.h:118, never defined */

#endif
