// Hand-written sm_100a kernels of the bit-plane Fitch path.
//
// Data layout in HBM: one "view" = the Fitch state sets of a directed
// subtree: S bit planes of Wl 32-bit words; bit j of word i of plane k = "expanded site
// 32*(w0+i)+j may be in state k" (same meaning as the reference's parsVect,
// sprparsimony.cpp:2870-2960; padding bits are 1 in every state so they never score).
// The planes are STATE-INTERLEAVED in groups of SG = min(S,4) states:
//     views[vid][group g][word w][j]  holds plane k = 4*g + j,
// so the 4 (or 2) state words of one site word are one aligned 128-bit (64-bit) vector: a lane
// fetches all DNA states of its word with ONE LDG.128 and a warp moves 512 contiguous bytes per
// group.  Wl is a multiple of 128 words.
//
// Kernels
//   k_compress_tips   R1  compressDNA                    (sprparsimony.cpp:2898-2961)
//   k_fitch_level     R3  newviewParsimonyIterativeFast  (sprparsimony.cpp:643-878)
//   k_edge_mismatch   R4  evaluateParsimonyIterativeFast (sprparsimony.cpp:1032-1205)
//   k_spr_scan        R6  testInsertParsimony batched over whole prune neighbourhoods
//                         (sprparsimony.cpp:2106-2188, 2208-2218, 2259-2376)
//   k_site_counters / k_gather_patterns
//                     R5  storePerSiteNodeScores + pllComputePatternParsimony
//                         (sprparsimony.cpp:294-343, 3363-3392)
#include "mpgpu_internal.h"

#include <cstdlib>

namespace mpgpu {

// ------------------------------------------------------------------------------------------
// state-set masks of the PLL codes (pllrepo/src/globalVariables.h:60-104, restated)
// ------------------------------------------------------------------------------------------
__host__ __device__ inline uint32_t code_mask(int datatype, uint32_t code)
{
    switch (datatype) {
    case MPGPU_AA_DATA:                         // bitVectorAA: 20 states, B = N|D, Z = Q|E, 22 = any
        if (code < 20) return 1u << code;
        if (code == 20) return 12u;
        if (code == 21) return 96u;
        return 0xFFFFFu;
    case MPGPU_GENERIC_32:                      // bitVector32: 32 states, 32 = any
        return code < 32 ? (1u << code) : 0xFFFFFFFFu;
    default:                                    // bitVectorIdentity (binary, DNA): code is the mask
        return code;
    }
}

// one immutable table per data type, built once (callers keep the pointer; contexts on several host threads share them)
const uint32_t *state_mask_table(int datatype, int *ncodes, int *undetermined)
{
    struct Tables {
        uint32_t t[4][256];                      // indexed by a code byte; codes the data type does not define map to the empty set
        Tables()
        {
            const int dts[4] = {MPGPU_BINARY_DATA, MPGPU_DNA_DATA, MPGPU_AA_DATA, MPGPU_GENERIC_32};
            const int ncs[4] = {4, 16, 23, 33};
            for (int k = 0; k < 4; k++) for (int i = 0; i < 256; i++) t[k][i] = i < ncs[k] ? code_mask(dts[k], (uint32_t)i) : 0u;
        }
    };
    static const Tables tables;                  // thread-safe initialisation (C++11 magic static)
    int nc = 0, und = 0, k = 0;
    switch (datatype) {
    case MPGPU_BINARY_DATA: nc = 4;  und = 3;  k = 0; break;
    case MPGPU_DNA_DATA:    nc = 16; und = 15; k = 1; break;
    case MPGPU_AA_DATA:     nc = 23; und = 22; k = 2; break;
    case MPGPU_GENERIC_32:  nc = 33; und = 32; k = 3; break;
    default: return nullptr;
    }
    if (ncodes) *ncodes = nc;
    if (undetermined) *undetermined = und;
    return tables.t[k];
}

// ------------------------------------------------------------------------------------------
// layout helpers
// ------------------------------------------------------------------------------------------
template <int S> struct Lay {
    static const int SG = S < 4 ? S : 4;          // states per interleave group
    static const int G = (S + SG - 1) / SG;       // groups per view
};

// all S state words of site word `w` of a view: p = view base + w*SG, group stride gs = Wl*SG
template <int S>
__device__ __forceinline__ void load_states(const uint32_t *__restrict__ p, size_t gs, uint32_t (&r)[S])
{
    if (Lay<S>::SG == 4) {
#pragma unroll
        for (int g = 0; g < Lay<S>::G; g++) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p + g * gs));
            r[4 * g] = v.x; r[4 * g + 1] = v.y; r[4 * g + 2] = v.z; r[4 * g + 3] = v.w;
        }
    } else {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
        r[0] = v.x; r[S > 1 ? 1 : 0] = v.y;
    }
}
template <int S>
__device__ __forceinline__ void load_states_rw(const uint32_t *p, size_t gs, uint32_t (&r)[S])
{
    if (Lay<S>::SG == 4) {
#pragma unroll
        for (int g = 0; g < Lay<S>::G; g++) {
            const uint4 v = *reinterpret_cast<const uint4 *>(p + g * gs);
            r[4 * g] = v.x; r[4 * g + 1] = v.y; r[4 * g + 2] = v.z; r[4 * g + 3] = v.w;
        }
    } else {
        const uint2 v = *reinterpret_cast<const uint2 *>(p);
        r[0] = v.x; r[S > 1 ? 1 : 0] = v.y;
    }
}
template <int S>
__device__ __forceinline__ void store_states(uint32_t *p, size_t gs, const uint32_t (&r)[S])
{
    if (Lay<S>::SG == 4) {
#pragma unroll
        for (int g = 0; g < Lay<S>::G; g++)
            *reinterpret_cast<uint4 *>(p + g * gs) = make_uint4(r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]);
    } else {
        *reinterpret_cast<uint2 *>(p) = make_uint2(r[0], r[S > 1 ? 1 : 0]);
    }
}
// (a & b) | (~n & (a | b)) : the Fitch set for one state given the "some state intersects" mask n
__device__ __forceinline__ uint32_t fitch1(uint32_t a, uint32_t b, uint32_t n) { return (a & b) | (~n & (a | b)); }
template <int S>
__device__ __forceinline__ uint32_t any_and(const uint32_t (&a)[S], const uint32_t (&b)[S])
{
    uint32_t n = 0;
#pragma unroll
    for (int k = 0; k < S; k++) n |= a[k] & b[k];
    return n;
}

// ------------------------------------------------------------------------------------------
// R1: tip planes from codes + pattern frequencies
// ------------------------------------------------------------------------------------------
// One thread per (tip, word of this shard).  site_start[k] = first expanded site of the k-th
// informative pattern (exclusive prefix sum of its aliaswgt), inf_ptn[k] = its pattern index.
__global__ void k_compress_tips(const uint8_t *__restrict__ codes, int P, int ntaxa,
                                const int64_t *__restrict__ site_start, const int32_t *__restrict__ inf_ptn,
                                int n_inf, int64_t n_sites, int datatype, int S, int Wl, int64_t w0,
                                uint32_t *__restrict__ views, size_t view_stride)
{
    int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)ntaxa * Wl) return;
    int tip = (int)(gid / Wl);
    int w = (int)(gid % Wl);
    int64_t site0 = (w0 + w) * 32;

    uint32_t mask[32];
    if (site0 >= n_sites) {
#pragma unroll
        for (int j = 0; j < 32; j++) mask[j] = 0xFFFFFFFFu;
    } else {
        // binary search: largest k with site_start[k] <= site0
        int lo = 0, hi = n_inf - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (site_start[mid] <= site0) lo = mid; else hi = mid - 1;
        }
        int k = lo;
        int64_t next_start = site_start[k + 1];
        uint32_t cur = code_mask(datatype, codes[(size_t)tip * P + inf_ptn[k]]);
#pragma unroll
        for (int j = 0; j < 32; j++) {
            int64_t site = site0 + j;
            if (site >= n_sites) { mask[j] = 0xFFFFFFFFu; continue; }
            while (site >= next_start) {
                k++;
                next_start = site_start[k + 1];
                cur = code_mask(datatype, codes[(size_t)tip * P + inf_ptn[k]]);
            }
            mask[j] = cur;
        }
    }
    const int SG = S < 4 ? S : 4;
    uint32_t *dst = views + (size_t)tip * view_stride;            // tips are views 0..n-1
    for (int s = 0; s < S; s++) {
        uint32_t word = 0;
#pragma unroll
        for (int j = 0; j < 32; j++) word |= ((mask[j] >> s) & 1u) << j;
        dst[(size_t)(s / SG) * Wl * SG + (size_t)w * SG + (s % SG)] = word;
    }
}

int launch_compress(Ctx *c)
{
    int64_t total = (int64_t)c->n * c->Wl;
    int threads = 128;
    int blocks = (int)((total + threads - 1) / threads);
    k_compress_tips<<<blocks, threads, 0, c->stream>>>(c->d_codes, c->P, c->n, c->d_site_start, c->d_inf_ptn,
                                                       c->n_inf, c->n_sites, c->datatype, c->S, c->Wl, c->w0,
                                                       c->d_views, c->view_stride);
    c->launches++;
    MPGPU_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// R3: one level of the directed-view schedule.  One thread per (triple, site word): all states
// of the word in registers (one LDG.128 per group and operand).  dst = fitch(a, b),
// count[dst] += popc(t_N).
// ------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(128) k_fitch_level(uint32_t *views, size_t view_stride, int Wl,
                                                     const Triple *__restrict__ triples, uint32_t *__restrict__ vcount)
{
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;          // Wl is a multiple of 128
    const Triple t = triples[blockIdx.y];
    const size_t gs = (size_t)Wl * Lay<S>::SG;
    const size_t off = (size_t)w * Lay<S>::SG;
    uint32_t a[S], b[S];
    load_states_rw<S>(views + (size_t)t.a * view_stride + off, gs, a);
    load_states_rw<S>(views + (size_t)t.b * view_stride + off, gs, b);
    const uint32_t n = any_and<S>(a, b);
#pragma unroll
    for (int k = 0; k < S; k++) a[k] = fitch1(a[k], b[k], n);
    store_states<S>(views + (size_t)t.dst * view_stride + off, gs, a);
    int cnt = __reduce_add_sync(0xffffffffu, __popc(~n));
    if (lane == 0 && cnt) atomicAdd(&vcount[t.dst], (uint32_t)cnt);
}

int launch_level(Ctx *c, const Triple *d_triples, int ntriples)
{
    if (ntriples == 0) return 0;
    dim3 grid(c->Wl / 128, ntriples);
    switch (c->S) {
    case 2:  k_fitch_level<2><<<grid, 128, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, d_triples, c->d_vcount); break;
    case 4:  k_fitch_level<4><<<grid, 128, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, d_triples, c->d_vcount); break;
    case 20: k_fitch_level<20><<<grid, 128, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, d_triples, c->d_vcount); break;
    case 32: k_fitch_level<32><<<grid, 128, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, d_triples, c->d_vcount); break;
    default: set_error("unsupported state count"); return 1;
    }
    c->launches++;
    MPGPU_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// R3, latency path: every stale view after an SPR move in ONE launch.  The stale views form a
// dependency forest as deep as the tree (each needs one stale and, mostly, one clean child): far
// too little work per level for a launch each, and a chain of L2 round trips if done naively.
// Fitch is word-local, so a CTA owns a 32-word column of ALL views and walks the levels alone:
//   * the triple list is staged in shared memory once,
//   * NW warps share the triples of a level; each warp prefetches the CLEAN operand of its next
//     triple (any level ahead) into registers while it works on the current one,
//   * a fresh view is written to global memory and to a two-generation shared-memory cache, so
//     the stale operand of the next level comes from shared memory, not from L2,
//   * __syncthreads() between levels; no CTA ever reads a word another CTA writes.
// Triple.pad = a_slot | dst_slot << 8 | b_clean << 16 (slots 0xFF: not cached).  list = hdr
// Triples reinterpreted as int32 level ends, then the triples; wcount[k] += popc(t_N) of triple k.
// ------------------------------------------------------------------------------------------
// host-mapped plan ranges -> their device arrays, spread over the grid (see StageArgs)
__device__ __forceinline__ void stage_copy(const StageArgs &st, int first, int stride)
{
#pragma unroll
    for (int k = 0; k < 3; k++)
        for (int i = first; i < st.n[k]; i += stride) st.dst[k][i] = st.src[k][i];
}
__global__ void __launch_bounds__(256) k_stage(const StageArgs st) { stage_copy(st, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x); }

int launch_stage(Ctx *c)
{
    if (!c->stage_pending) return 0;
    c->stage_pending = false;
    const int n16 = c->stage_req.n[0] + c->stage_req.n[1] + c->stage_req.n[2];
    int blocks = (n16 + 255) / 256;                 // one 16-byte unit per thread where the plan is large: a single round trip to host memory
    if (blocks < 8) blocks = 8;
    if (blocks > 64) blocks = 64;
    k_stage<<<blocks, 256, 0, c->stream>>>(c->stage_req);
    c->launches++;
    MPGPU_CUDA(cudaGetLastError());
    return 0;
}

template <int S> struct WaveCfg { static const int CAP = S <= 4 ? 32 : (S <= 20 ? 16 : 12); };

template <int S, int NW>
__global__ void __launch_bounds__(32 * NW) k_fitch_wave(uint32_t *views, size_t view_stride, int Wl,
                                                        const Triple *__restrict__ list, int nlevels, int hdr, int total,
                                                        uint32_t *__restrict__ wcount, const StageArgs st)
{
    extern __shared__ uint4 wave_smem[];
    // the next scan's plan rides along (latency path of the search): one extra CTA copies it while the others update views
    if (blockIdx.x * 32 >= Wl) { stage_copy(st, threadIdx.x, blockDim.x); return; }
    const int CAP = WaveCfg<S>::CAP;
    int4 *sl = reinterpret_cast<int4 *>(wave_smem);
    uint32_t *cache = reinterpret_cast<uint32_t *>(sl + hdr + total);
    for (int i = threadIdx.x; i < hdr + total; i += 32 * NW) sl[i] = __ldg(reinterpret_cast<const int4 *>(list) + i);
    __syncthreads();
    const int32_t *level_end = reinterpret_cast<const int32_t *>(sl);
    const int4 *tri = sl + hdr;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = blockIdx.x * 32 + lane;
    const size_t gs = (size_t)Wl * Lay<S>::SG;
    const size_t off = (size_t)w * Lay<S>::SG;

    // this warp's next triple: index nt, in level nl.  The triples are dealt to the warps round robin over the WHOLE list, not per
    // level: a lazy list of the search is a path with a view or two per level, and a warp that owns triples 8 levels apart has the
    // clean operand of its next one in flight long before that level is reached (dealt per level, warps 0 and 1 did every level
    // and waited a full L2 round trip in each).
    int nt = warp < total ? warp : -1, nl = 0;
    if (nt >= 0) while (nt >= level_end[nl]) nl++;
    uint32_t bp[S];
    if (nt >= 0) { const int4 d = tri[nt]; if (d.w & 0x10000) load_states_rw<S>(views + (size_t)d.z * view_stride + off, gs, bp); }

    for (int l = 0; l < nlevels; l++) {
        while (nt >= 0 && nl == l) {
            const int cur = nt;
            const int4 d = tri[cur];
            uint32_t a[S], b[S];
#pragma unroll
            for (int k = 0; k < S; k++) b[k] = bp[k];
            // next triple of this warp and its clean operand
            nt = cur + NW < total ? cur + NW : -1;
            if (nt >= 0) while (nt >= level_end[nl]) nl++;
            if (nt >= 0) { const int4 dn = tri[nt]; if (dn.w & 0x10000) load_states_rw<S>(views + (size_t)dn.z * view_stride + off, gs, bp); }
            if (!(d.w & 0x10000)) load_states_rw<S>(views + (size_t)d.z * view_stride + off, gs, b);
            const int a_slot = d.w & 0xFF, d_slot = (d.w >> 8) & 0xFF;
            if (a_slot != 0xFF) {
                const uint32_t *src = cache + ((size_t)(((l + 1) & 1) * CAP + a_slot) * S) * 32 + lane;
#pragma unroll
                for (int k = 0; k < S; k++) a[k] = src[k * 32];
            } else {
                load_states_rw<S>(views + (size_t)d.y * view_stride + off, gs, a);
            }
            const uint32_t n = any_and<S>(a, b);
#pragma unroll
            for (int k = 0; k < S; k++) a[k] = fitch1(a[k], b[k], n);
            store_states<S>(views + (size_t)d.x * view_stride + off, gs, a);
            if (d_slot != 0xFF) {
                uint32_t *dstc = cache + ((size_t)((l & 1) * CAP + d_slot) * S) * 32 + lane;
#pragma unroll
                for (int k = 0; k < S; k++) dstc[k * 32] = a[k];
            }
            const int cnt = __reduce_add_sync(0xffffffffu, __popc(~n));
            if (lane == 0 && cnt) atomicAdd(&wcount[cur], (uint32_t)cnt);
        }
        __syncthreads();
    }
}

int wave_slot_cap(int S) { return S <= 4 ? 32 : (S <= 20 ? 16 : 12); }
size_t wave_smem_bytes(int S, int entries) { return (size_t)entries * sizeof(Triple) + (size_t)2 * wave_slot_cap(S) * S * 128; }

template <int S, int NW>
static int launch_wave_t(Ctx *c, const Triple *d_list, int nlevels, int hdr, int total, uint32_t *d_wcount)
{
    const size_t smem = wave_smem_bytes(S, hdr + total);
    static size_t opted_dev[64] = {0}; size_t &opted = opted_dev[c->device & 63];   /* the attribute is per device */
    if (smem > opted) {
        MPGPU_CUDA(cudaFuncSetAttribute(k_fitch_wave<S, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        opted = smem;
    }
    StageArgs st = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}, {0, 0, 0}};
    if (c->stage_pending) { st = c->stage_req; c->stage_pending = false; }
    const bool staging = st.n[0] + st.n[1] + st.n[2] > 0;
    k_fitch_wave<S, NW><<<c->Wl / 32 + (staging ? 1 : 0), 32 * NW, smem, c->stream>>>(c->d_views, c->view_stride, c->Wl, d_list, nlevels, hdr, total, d_wcount, st);
    return 0;
}

int launch_wave(Ctx *c, const Triple *d_list, int nlevels, int hdr, int total, uint32_t *d_wcount)
{
    if (nlevels == 0) return 0;
    // 16 warps per CTA: with the triples dealt round robin over the whole list a warp's next clean operand is in flight ~8 levels
    // ahead on the lazy lists of the search (a view or two per level); 4 warps measured the same before that change
    static const int forced = getenv("MPGPU_WAVE_WARPS") ? atoi(getenv("MPGPU_WAVE_WARPS")) : 0;      // tuning knob
    const bool narrow = forced == 4;
    int rc = 0;
    if (narrow) {
        switch (c->S) {
        case 2:  rc = launch_wave_t<2, 4>(c, d_list, nlevels, hdr, total, d_wcount); break;
        case 4:  rc = launch_wave_t<4, 4>(c, d_list, nlevels, hdr, total, d_wcount); break;
        case 20: rc = launch_wave_t<20, 4>(c, d_list, nlevels, hdr, total, d_wcount); break;
        case 32: rc = launch_wave_t<32, 4>(c, d_list, nlevels, hdr, total, d_wcount); break;
        default: set_error("unsupported state count"); return 1;
        }
    } else {
        const int NW = 16;
        switch (c->S) {
        case 2:  rc = launch_wave_t<2, NW>(c, d_list, nlevels, hdr, total, d_wcount); break;
        case 4:  rc = launch_wave_t<4, NW>(c, d_list, nlevels, hdr, total, d_wcount); break;
        case 20: rc = launch_wave_t<20, NW>(c, d_list, nlevels, hdr, total, d_wcount); break;
        case 32: rc = launch_wave_t<32, NW>(c, d_list, nlevels, hdr, total, d_wcount); break;
        default: set_error("unsupported state count"); return 1;
        }
    }
    if (rc) return rc;
    c->launches++;
    MPGPU_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// R4: mismatch count across one edge: popc(~OR_k(A_k & B_k))
// ------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(128) k_edge_mismatch(const uint32_t *__restrict__ views, size_t view_stride, int Wl,
                                                       int vidA, int vidB, uint32_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t gs = (size_t)Wl * Lay<S>::SG;
    const size_t off = (size_t)w * Lay<S>::SG;
    uint32_t a[S], b[S];
    load_states<S>(views + (size_t)vidA * view_stride + off, gs, a);
    load_states<S>(views + (size_t)vidB * view_stride + off, gs, b);
    int cnt = __reduce_add_sync(0xffffffffu, __popc(~any_and<S>(a, b)));
    if (lane == 0 && cnt) atomicAdd(out, (uint32_t)cnt);
}

int launch_edge_mismatch(Ctx *c, int vidA, int vidB, uint32_t *d_out)
{
    dim3 grid(c->Wl / 128);
    switch (c->S) {
    case 2:  k_edge_mismatch<2><<<grid, 128, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, vidA, vidB, d_out); break;
    case 4:  k_edge_mismatch<4><<<grid, 128, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, vidA, vidB, d_out); break;
    case 20: k_edge_mismatch<20><<<grid, 128, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, vidA, vidB, d_out); break;
    case 32: k_edge_mismatch<32><<<grid, 128, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, vidA, vidB, d_out); break;
    default: set_error("unsupported state count"); return 1;
    }
    c->launches++;
    MPGPU_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// R7: stepwise addition (stepwiseAddition, sprparsimony.cpp:2977-3019): the extra steps of
// hanging tip T on the branch between the views A and B: popc(~OR_k(fitch(A,B)_k & T_k)).
// One thread per (branch, site word); the tree's own length is added on the host.
// ------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(128) k_tip_insert(const uint32_t *__restrict__ views, size_t view_stride, int Wl,
                                                    const int4 *__restrict__ edges, int32_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int4 e = edges[blockIdx.y];                              // vidA, vidB, vidT
    const size_t gs = (size_t)Wl * Lay<S>::SG;
    const size_t off = (size_t)w * Lay<S>::SG;
    uint32_t a[S], b[S];
    load_states<S>(views + (size_t)e.x * view_stride + off, gs, a);
    load_states<S>(views + (size_t)e.y * view_stride + off, gs, b);
    const uint32_t n = any_and<S>(a, b);
#pragma unroll
    for (int k = 0; k < S; k++) a[k] = fitch1(a[k], b[k], n);
    load_states<S>(views + (size_t)e.z * view_stride + off, gs, b);
    const int cnt = __reduce_add_sync(0xffffffffu, __popc(~any_and<S>(a, b)));
    if (lane == 0 && cnt) atomicAdd(&out[blockIdx.y], cnt);
}

int launch_tip_insert(Ctx *c, const int4 *d_edges, int nedges, int32_t *d_out)
{
    if (nedges == 0) return 0;
    dim3 grid(c->Wl / 128, nedges);
    switch (c->S) {
    case 2:  k_tip_insert<2><<<grid, 128, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, d_edges, d_out); break;
    case 4:  k_tip_insert<4><<<grid, 128, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, d_edges, d_out); break;
    case 20: k_tip_insert<20><<<grid, 128, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, d_edges, d_out); break;
    case 32: k_tip_insert<32><<<grid, 128, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, d_edges, d_out); break;
    default: set_error("unsupported state count"); return 1;
    }
    c->launches++;
    MPGPU_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// R6: the SPR scan.
//
// A task is one pruned subtree S with the two views D1, D2 that become neighbours once the
// pruned node is taken out (removeNodeParsimony, sprparsimony.cpp:2245).  Its program is a
// list of "expand" ops, parent before child: expanding node y (whose up-view U_y -- everything
// on the far side of y in the pruned tree -- is in a stack slot, or for the two top-level
// nodes is simply D2/D1) reads the views A, B of y's two children once and produces for each
// child c with sibling view X:   U_c = fitch(U_y, X)          (the newview of insertParsimony)
//                                F   = fitch(U_c, view(c))    (the re-oriented p, :1951)
//                                cnt = popc(~OR_k(F_k & S_k)) (the evaluate at :2160)
// Because Fitch length is root-invariant the full score of the insertion is
// len(S) + len(pruned tree) + cnt, and len(pruned tree) = len(D1)+len(D2)+popc(~any(D1&D2)),
// which the task accumulates once into base_out.  Each child view is read once per task
// (4*S*W bytes per insertion instead of the canonical 8*S*W), the up-views never leave the SM.
//
// One warp owns (task, chunk of 32 words); lane l owns word l of the chunk for every state
// and every stack slot, so the stack needs no synchronisation at all.
// Work order: all tasks of chunk 0, then chunk 1, ... so that concurrently resident warps
// touch the same few word columns of every view and the views are served from L1/L2.
// ------------------------------------------------------------------------------------------
// shared-memory stack access by 32-bit shared address (keeps address arithmetic to one IMAD)
__device__ __forceinline__ void lds_vec(uint32_t addr, uint4 &v)
{
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
}
__device__ __forceinline__ void sts_vec(uint32_t addr, const uint4 &v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void lds_vec(uint32_t addr, uint2 &v)
{
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
}
__device__ __forceinline__ void sts_vec(uint32_t addr, const uint2 &v)
{
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(addr), "r"(v.x), "r"(v.y) : "memory");
}

template <int S> struct VecOf { typedef uint4 T; };
template <> struct VecOf<2> { typedef uint2 T; };

// A view operand held as G vectors of SG states
template <int S> struct StateVec {
    typedef typename VecOf<S>::T V;
    V g[Lay<S>::G];
    __device__ __forceinline__ void load(const V *__restrict__ p, uint32_t gstride)
    {
#pragma unroll
        for (int i = 0; i < Lay<S>::G; i++) g[i] = __ldg(p + (size_t)i * gstride);
    }
    __device__ __forceinline__ void load_shared(uint32_t addr)
    {
#pragma unroll
        for (int i = 0; i < Lay<S>::G; i++) lds_vec(addr + i * 32 * (int)sizeof(V), g[i]);
    }
    __device__ __forceinline__ void store_shared(uint32_t addr) const
    {
#pragma unroll
        for (int i = 0; i < Lay<S>::G; i++) sts_vec(addr + i * 32 * (int)sizeof(V), g[i]);
    }
    __device__ __forceinline__ void unpack(uint32_t (&r)[S]) const
    {
#pragma unroll
        for (int i = 0; i < Lay<S>::G; i++) {
            const uint32_t *w = reinterpret_cast<const uint32_t *>(&g[i]);
#pragma unroll
            for (int j = 0; j < Lay<S>::SG; j++) r[Lay<S>::SG * i + j] = w[j];
        }
    }
    __device__ __forceinline__ void pack(const uint32_t (&r)[S])
    {
#pragma unroll
        for (int i = 0; i < Lay<S>::G; i++) {
            uint32_t *w = reinterpret_cast<uint32_t *>(&g[i]);
#pragma unroll
            for (int j = 0; j < Lay<S>::SG; j++) w[j] = r[Lay<S>::SG * i + j];
        }
    }
};

// Scores one child on the VW words this lane owns: Uc = fitch(U, X) (X = sibling view), optional store of Uc, optional
// count.  ROWS: instead of counting, the mismatch bits of the insertion (the candidate's per-site delta row under -bb,
// DESIGN.md section 5) go to rowp[32*j] when rowp is not null.  One REDUX and one RED per child whatever VW is.
template <int S, bool ROWS, int VW>
__device__ __forceinline__ void scan_child(uint32_t (&U)[VW][S], const uint32_t (&X)[VW][S], const uint32_t (&C)[VW][S],
                                           const uint32_t (&Sv)[VW][S], bool do_out, int32_t *__restrict__ outp,
                                           uint32_t *__restrict__ rowp,
                                           bool do_dst, uint32_t dst_addr, bool lane0, bool in_place)
{
    uint32_t Uc[VW][S];
#pragma unroll
    for (int j = 0; j < VW; j++) {
        const uint32_t n = any_and<S>(U[j], X[j]);
#pragma unroll
        for (int k = 0; k < S; k++) Uc[j][k] = fitch1(U[j][k], X[j][k], n);
    }
    if (do_dst) {
#pragma unroll
        for (int j = 0; j < VW; j++) {
            StateVec<S> sv; sv.pack(Uc[j]);
            sv.store_shared(dst_addr + j * (Lay<S>::G * 32 * (int)sizeof(typename VecOf<S>::T)));
        }
    }
    if (do_out) {
        int pc = 0;
#pragma unroll
        for (int j = 0; j < VW; j++) {
            const uint32_t m = any_and<S>(Uc[j], C[j]);
            uint32_t z = 0;
#pragma unroll
            for (int k = 0; k < S; k++) z |= fitch1(Uc[j][k], C[j][k], m) & Sv[j][k];
            if (ROWS) { if (rowp) rowp[32 * j] = ~z; }
            else pc += __popc(~z);
        }
        if (!ROWS) {
            const int cnt = __reduce_add_sync(0xffffffffu, pc);
            if (lane0) atomicAdd(outp, cnt);
        }
    }
    if (in_place) {                 // the first child's up-view stays in registers: the very next op expands that child
#pragma unroll
        for (int j = 0; j < VW; j++)
#pragma unroll
            for (int k = 0; k < S; k++) U[j][k] = Uc[j][k];
    }
}

// keep a value in its register: stops ptxas from re-deriving it from special registers /
// kernel parameters at every use (it did, ~8 instructions per shared or global address)
__device__ __forceinline__ void pin(uint32_t &x) { asm volatile("" : "+r"(x)); }
__device__ __forceinline__ void pin(const void *&x) { asm volatile("" : "+l"(x)); }

// VW site words of a view operand, one StateVec per word; word j of this lane sits 32 vectors after word j-1
template <int S, int VW> struct WideVec {
    typedef typename VecOf<S>::T V;
    StateVec<S> w[VW];
    __device__ __forceinline__ void load(const V *__restrict__ p, uint32_t gstride)
    {
#pragma unroll
        for (int j = 0; j < VW; j++) w[j].load(p + 32 * j, gstride);
    }
    __device__ __forceinline__ void load_shared(uint32_t addr)
    {
#pragma unroll
        for (int j = 0; j < VW; j++) w[j].load_shared(addr + j * (Lay<S>::G * 32 * (int)sizeof(V)));
    }
    __device__ __forceinline__ void unpack(uint32_t (&r)[VW][S]) const
    {
#pragma unroll
        for (int j = 0; j < VW; j++) w[j].unpack(r[j]);
    }
};

// The program arrives as two streams so that the loads of op i+1's child views can be issued
// while op i computes without rotating whole op records through registers:
//   offs[i] = (c1, c2) view offsets in vector units
//   ctl[i]  = x: out1 | out2 << 16, each relative to the task's first candidate, 0xFFFF = none
//             y: src | dst1 << 8 | dst2 << 16   (stack slots; 0xFF = none;
//                src 0xFF / 0xFE = the task's D2 / D1 view for the two top-level expansions)
// PF = also software-prefetch op i+1's child views into a second register set (small S only).
// ROWS = the -bb second pass: for the candidates with row_of[candidate] >= 0 the per-site
// mismatch row is written to rows[row_of][Wl] instead of counting (tasks = only those that
// have such a candidate; task_ids maps the dense task index to the plan's task).
// VW = 32-word chunks per warp (1, 2 or 4; Wl is a multiple of 128 words): a lane owns word l of each of them, so the
// control-word decode, the offs / ctl loads, the branches and the address arithmetic of an op are paid once for VW
// chunks and a scored child costs one REDUX + one RED (r02: 66 -> ~38 warp instructions per insertion and chunk).
template <int S, bool PF, bool ROWS, int VW>
__device__ __forceinline__ void scan_body(const typename VecOf<S>::T *__restrict__ views, int Wl,
                           const ScanTask *__restrict__ tasks, int ntasks,
                           const int2 *__restrict__ offs, const int2 *__restrict__ ctl,
                           int nslots, int32_t *__restrict__ out,
                           const int32_t *__restrict__ task_ids, const int32_t *__restrict__ row_of, int row_bias,
                           uint32_t *__restrict__ rows)
{
    typedef typename VecOf<S>::T V;
    extern __shared__ uint4 smem4[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned gw = blockIdx.x * (blockDim.x >> 5) + warp;
    const unsigned ngroups = Wl / (kChunkWords * VW);
    if (gw >= (unsigned)ntasks * ngroups) return;
    const unsigned cgroup = gw / (unsigned)ntasks;
    unsigned ti = gw - cgroup * (unsigned)ntasks;
    if (ROWS) ti = (unsigned)__ldg(task_ids + ti);
    const int4 t0 = __ldg(reinterpret_cast<const int4 *>(tasks + ti));       // s_vid, d1, d2, op_begin
    const int4 t1 = __ldg(reinterpret_cast<const int4 *>(tasks + ti) + 1);   // op_end, base_out, cand_base
    const uint32_t gsv = (uint32_t)Wl;                                       // group stride in vectors
    const void *vb_ = views + (size_t)cgroup * (kChunkWords * VW) + lane;
    pin(vb_);
    const V *vbase = static_cast<const V *>(vb_);
    constexpr uint32_t kSlotBytes = VW * Lay<S>::G * 32 * sizeof(V);         // [word j][group][lane] vectors
    // stack slots 0 and 1 (first children) never reach shared memory: such an up-view is consumed by the very next op and
    // stays in registers; the stack holds slots 2 .. nslots + 1 (second children, one per depth), addressed as slot * bytes
    uint32_t sstack = (uint32_t)__cvta_generic_to_shared(smem4) +
                      (uint32_t)warp * (uint32_t)nslots * kSlotBytes + lane * (uint32_t)sizeof(V) - 2u * kSlotBytes;
    pin(sstack);
    const bool lane0 = lane == 0;
    int32_t *outc = out + t1.z;                                              // candidates of this task
    const int32_t *rowc = ROWS ? row_of + (t1.z - row_bias) : nullptr;   // cand_base counts from the base slots
    uint32_t *rowbase = ROWS ? rows + (size_t)cgroup * (kChunkWords * VW) + lane : nullptr;

    uint32_t Sv[VW][S];
    {
        WideVec<S, VW> a, b, c;
        a.load(vbase + (uint32_t)t0.x, gsv);
        b.load(vbase + (uint32_t)t0.y, gsv);
        c.load(vbase + (uint32_t)t0.z, gsv);
        a.unpack(Sv);
        if (!ROWS && t1.y >= 0) {                   // (a sub-task with base_out < 0 leaves the joined-edge count to its first sibling)
            uint32_t d1[VW][S], d2[VW][S];
            b.unpack(d1); c.unpack(d2);
            int pc = 0;
#pragma unroll
            for (int j = 0; j < VW; j++) pc += __popc(~any_and<S>(d1[j], d2[j]));
            const int cnt = __reduce_add_sync(0xffffffffu, pc);
            if (lane0) atomicAdd(&out[t1.y], cnt);
        }
    }

    int oi = t0.w;
    const int oe = t1.x;
    if (oi >= oe) return;

    int2 f0 = __ldg(offs + oi);
    int2 f1 = f0;
    if (oi + 1 < oe) f1 = __ldg(offs + oi + 1);
    WideVec<S, VW> A0, B0, A1, B1;
    A0.load(vbase + (uint32_t)f0.x, gsv);
    B0.load(vbase + (uint32_t)f0.y, gsv);

#define MPGPU_SCAN_STEP(AC, BC, AN, BN, FCUR, FNEXT, FLOAD, CWCUR, CWNEXT)                             \
    {                                                                                                  \
        if (PF) {                                                                                      \
            if (oi + 1 < oe) { AN.load(vbase + (uint32_t)FNEXT.x, gsv); BN.load(vbase + (uint32_t)FNEXT.y, gsv); } \
        } else {                                                                                       \
            AC.load(vbase + (uint32_t)FCUR.x, gsv); BC.load(vbase + (uint32_t)FCUR.y, gsv);            \
        }                                                                                              \
        if (oi + 2 < oe) FLOAD = __ldg(offs + oi + 2);                                                 \
        const int2 cw = CWCUR;                                                                         \
        if (oi + 1 < oe) CWNEXT = __ldg(ctl + oi + 1);      /* the control word of the next op, one op ahead */ \
        const uint32_t src = cw.y & 0xff, dst1 = (cw.y >> 8) & 0xff, dst2 = (cw.y >> 16) & 0xff;       \
        const uint32_t o1 = cw.x & 0xffff, o2 = (uint32_t)cw.x >> 16;                                  \
        /* src 0 / 1: the up-view the previous op left in U (its first child, expanded right away) */  \
        if (src >= 2) {                                                                                \
            WideVec<S, VW> Uv;                                                                         \
            if (src < 0xfe) Uv.load_shared(sstack + src * kSlotBytes);                                 \
            else Uv.load(vbase + (uint32_t)(src == 0xff ? t0.z : t0.y), gsv);                          \
            Uv.unpack(U);                                                                              \
        }                                                                                              \
        uint32_t A[VW][S], B[VW][S];                                                                   \
        AC.unpack(A); BC.unpack(B);                                                                    \
        uint32_t *rp1 = nullptr, *rp2 = nullptr;                                                       \
        if (ROWS) {                                                                                    \
            if (o1 != 0xffff) { const int r = __ldg(rowc + o1); if (r >= 0) rp1 = rowbase + (size_t)r * Wl; } \
            if (o2 != 0xffff) { const int r = __ldg(rowc + o2); if (r >= 0) rp2 = rowbase + (size_t)r * Wl; } \
        }                                                                                              \
        /* second child first: its up-view goes to the stack; then the first child's replaces U */     \
        if (o2 != 0xffff || dst2 != 0xff)                                                              \
            scan_child<S, ROWS, VW>(U, A, B, Sv, o2 != 0xffff, outc + o2, rp2, dst2 != 0xff, sstack + dst2 * kSlotBytes, lane0, false); \
        if (o1 != 0xffff || dst1 != 0xff)                                                              \
            scan_child<S, ROWS, VW>(U, B, A, Sv, o1 != 0xffff, outc + o1, rp1, false, 0u, lane0, dst1 != 0xff); \
    }

    uint32_t U[VW][S];
#pragma unroll
    for (int j = 0; j < VW; j++)
#pragma unroll
        for (int k = 0; k < S; k++) U[j][k] = 0;

    int2 cw0 = __ldg(ctl + oi), cw1 = cw0;
    for (;;) {
        MPGPU_SCAN_STEP(A0, B0, A1, B1, f0, f1, f0, cw0, cw1)
        if (++oi >= oe) break;
        MPGPU_SCAN_STEP(A1, B1, A0, B0, f1, f0, f1, cw1, cw0)
        if (++oi >= oe) break;
    }
#undef MPGPU_SCAN_STEP
}

// pub.flag != nullptr (latency path of the search): the block that finishes last copies the counts -- and the counts of the
// view updates in flight -- to mapped page-locked memory, puts the device counters back to zero and raises the flag the
// host spins on: no copy engine, no second launch, no stream synchronize in the step.
template <int S, bool PF, bool ROWS, int VW, bool PUB>
__global__ void k_spr_scan(const typename VecOf<S>::T *__restrict__ views, int Wl,
                           const ScanTask *__restrict__ tasks, int ntasks,
                           const int2 *__restrict__ offs, const int2 *__restrict__ ctl,
                           int nslots, int32_t *__restrict__ out,
                           const int32_t *__restrict__ task_ids, const int32_t *__restrict__ row_of, int row_bias,
                           uint32_t *__restrict__ rows, const PubArgs pub)
{
    scan_body<S, PF, ROWS, VW>(views, Wl, tasks, ntasks, offs, ctl, nslots, out, task_ids, row_of, row_bias, rows);
    if (!PUB) return;              // (a template parameter: the tail costs the wide variants 14 registers, i.e. a resident CTA)
    __shared__ int s_last;
    __syncthreads();                                 // every RED of this block is issued
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(pub.done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int i = threadIdx.x; i < pub.nout; i += blockDim.x) { pub.host_counts[i] = __ldcg(out + i); out[i] = 0; }
    for (int i = threadIdx.x; i < pub.nwc; i += blockDim.x) { pub.host_wc[i] = __ldcg(pub.wcount + i); pub.wcount[i] = 0; }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {          // (every thread fenced its own writes at system scope before the barrier)
        *pub.done = 0;
        *reinterpret_cast<volatile uint32_t *>(pub.flag) = pub.epoch;
    }
}

template <int S, bool ROWS, int VW, bool PF4 = false, bool PUB = false>
static int launch_scan_v(Ctx *c, int task0, int ntasks, int nslots)
{
    if (!PUB && !ROWS && VW == 1 && !PF4 && c->pub_request) return launch_scan_v<S, ROWS, VW, PF4, !ROWS && VW == 1>(c, task0, ntasks, nslots);
    typedef typename VecOf<S>::T V;
    constexpr bool PF = S <= 4 && (VW <= 2 || PF4);
    nslots = nslots > 2 ? nslots - 2 : 1;       // the planner's slots 0 / 1 live in registers (see k_spr_scan)
    const size_t per_warp = (size_t)nslots * Lay<S>::G * 32 * sizeof(V) * VW;
    int wpb = 4;           // small CTAs: ragged task lengths retire early, measured best on B200 (profiles/)
    if (const char *e = getenv("MPGPU_SCAN_WPB")) { int v = atoi(e); if (v >= 1 && v <= 32) wpb = v; }   // tuning knob
    const size_t budget = 96 * 1024;
    while (wpb > 1 && per_warp * wpb > budget) wpb >>= 1;
    const size_t smem = per_warp * wpb;
    if (smem > 200 * 1024) { set_error("scan stack does not fit in shared memory"); return 1; }
    static size_t configured_dev[64] = {0}; size_t &configured = configured_dev[c->device & 63];   /* the attribute is per device */
    if (smem > 48 * 1024 && smem > configured) {
        MPGPU_CUDA(cudaFuncSetAttribute(k_spr_scan<S, PF, ROWS, VW, PUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
        configured = 200 * 1024;
    }
    const long long warps = (long long)ntasks * (c->Wl / (kChunkWords * VW));
    const long long blocks = (warps + wpb - 1) / wpb;
    if (blocks > 0x7fffffffLL || warps > 0xffffffffLL) { set_error("scan grid too large"); return 1; }
    PubArgs pub = {nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0u};
    if (PUB && c->pub_request) {
        c->pub_request = false;
        c->flag_epoch++;
        if (c->flag_epoch == 0) c->flag_epoch = 1;
        pub.host_counts = c->h_counts; pub.host_wc = c->wcount_pin.data(); pub.wcount = c->d_wcount; pub.flag = c->h_flag; pub.done = c->d_done;
        pub.nout = c->pub_nout; pub.nwc = (int)c->wc_used; pub.epoch = c->flag_epoch;
        c->pub_inflight = true;
    }
    k_spr_scan<S, PF, ROWS, VW, PUB><<<(unsigned)blocks, wpb * 32, smem, c->stream>>>(
        reinterpret_cast<const V *>(c->d_views), c->Wl, c->d_tasks + (ROWS ? 0 : task0), ntasks, reinterpret_cast<const int2 *>(c->d_offs),
        reinterpret_cast<const int2 *>(c->d_ctl), nslots, c->d_counts,
        ROWS ? c->d_row_tasks : nullptr, ROWS ? c->d_row_of : nullptr, c->plan.task_cap, ROWS ? c->d_rows_site : nullptr, pub);
    c->launches++;
    MPGPU_CUDA(cudaGetLastError());
    return 0;
}

// Chunks per warp: as wide as the grid allows while it still fills the device a few times over (a small batch of a small
// alignment keeps one chunk per warp: there the launch is latency-bound and more warps finish sooner).
template <int S, bool ROWS>
static int launch_scan_t(Ctx *c, int task0, int ntasks, int nslots)
{
    int vw = 1;
    if (S <= 4) {
        static const int forced = getenv("MPGPU_SCAN_VW") ? atoi(getenv("MPGPU_SCAN_VW")) : 0;     // tuning knob
        static int sms_dev[64] = {0};
        int &sms = sms_dev[c->device & 63];
        if (!sms) { cudaDeviceProp prop; MPGPU_CUDA(cudaGetDeviceProperties(&prop, c->device)); sms = prop.multiProcessorCount; }
        const long long warps1 = (long long)ntasks * (c->Wl / kChunkWords);
        const long long fill = (long long)sms * 32;                     // ~2 waves of 16 resident warps per SM
        static const int auto_max = getenv("MPGPU_SCAN_VW_MAX") ? atoi(getenv("MPGPU_SCAN_VW_MAX")) : 4;   // measured on B200 (profiles/r02b): C2 sweep 0.1605 / 0.1603 / 0.1476 ms at 1 / 2 / 4 chunks per warp
        if (auto_max >= 4 && warps1 / 4 >= fill) vw = 4;
        else if (auto_max >= 2 && warps1 / 2 >= fill) vw = 2;
        if (forced == 1 || forced == 2 || forced == 4) vw = forced;
    }
    switch (vw) {
    case 4: {
        static const bool pf4 = getenv("MPGPU_SCAN_PF4") && atoi(getenv("MPGPU_SCAN_PF4")) != 0;    // tuning knob
        if (pf4) return launch_scan_v<S, ROWS, (S <= 4 ? 4 : 1), true>(c, task0, ntasks, nslots);
        return launch_scan_v<S, ROWS, (S <= 4 ? 4 : 1)>(c, task0, ntasks, nslots);
    }
    case 2:  return launch_scan_v<S, ROWS, (S <= 4 ? 2 : 1)>(c, task0, ntasks, nslots);
    default: return launch_scan_v<S, ROWS, 1>(c, task0, ntasks, nslots);
    }
}

// tasks [task0, task0 + ntasks) of the uploaded plan
int launch_scan(Ctx *c, int task0, int ntasks, int nslots)
{
    if (ntasks == 0) return 0;
    switch (c->S) {
    case 2:  return launch_scan_t<2, false>(c, task0, ntasks, nslots);
    case 4:  return launch_scan_t<4, false>(c, task0, ntasks, nslots);
    case 20: return launch_scan_t<20, false>(c, task0, ntasks, nslots);
    case 32: return launch_scan_t<32, false>(c, task0, ntasks, nslots);
    default: set_error("unsupported state count"); return 1;
    }
}

// -bb second pass over the tasks listed in d_row_tasks (see k_spr_scan<ROWS>)
int launch_scan_rows(Ctx *c, int ntasks, int nslots)
{
    if (ntasks == 0) return 0;
    switch (c->S) {
    case 2:  return launch_scan_t<2, true>(c, 0, ntasks, nslots);
    case 4:  return launch_scan_t<4, true>(c, 0, ntasks, nslots);
    case 20: return launch_scan_t<20, true>(c, 0, ntasks, nslots);
    case 32: return launch_scan_t<32, true>(c, 0, ntasks, nslots);
    default: set_error("unsupported state count"); return 1;
    }
}

// ------------------------------------------------------------------------------------------
// R5: per-site mismatch counters of the current tree, kept bit-sliced: plane b of the counter
// holds bit b of the count of every site, so adding a 1-bit mismatch row is a ripple-carry of
// word-wide AND/XOR instead of the reference's 32 scalar increments per word
// (storePerSiteNodeScores, sprparsimony.cpp:294-319).  pairs = the (a,b) child views of the
// n-2 inner views facing tr->start plus the start edge itself.
// ------------------------------------------------------------------------------------------
// A CTA owns 32 site words; its NW warps share the pairs (U at a time each, loads in flight together), every warp folds
// its share into bit-sliced counters in registers, and the partial counters are added plane by plane (a ripple-carry
// adder over the 16 planes) in a shared-memory tree.  (r01: one thread per word walked all ~n pairs alone, 122 us per
// tree on C2 with 3200 threads on the whole device.)
template <int S, int NW>
__global__ void __launch_bounds__(32 * NW) k_site_counters(const uint32_t *__restrict__ views, size_t view_stride, int Wl,
                                                           const int32_t *__restrict__ pairs, int npairs, int nbits,
                                                           uint32_t *__restrict__ bitcnt)
{
    __shared__ uint32_t part[NW][16][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = blockIdx.x * 32 + lane;                 // Wl is a multiple of 128
    const size_t gs = (size_t)Wl * Lay<S>::SG;
    const size_t off = (size_t)w * Lay<S>::SG;
    uint32_t cnt[16];
#pragma unroll
    for (int b = 0; b < 16; b++) cnt[b] = 0;
    constexpr int U = S <= 4 ? 8 : 2;
    for (int i0 = warp * U; i0 < npairs; i0 += NW * U) {
        uint32_t mis[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            mis[u] = 0;
            if (i0 + u < npairs) {
                uint32_t a[S], bb[S];
                load_states<S>(views + (size_t)__ldg(pairs + 2 * (i0 + u)) * view_stride + off, gs, a);
                load_states<S>(views + (size_t)__ldg(pairs + 2 * (i0 + u) + 1) * view_stride + off, gs, bb);
                mis[u] = ~any_and<S>(a, bb);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            uint32_t carry = mis[u];
#pragma unroll
            for (int b = 0; b < 16; b++) {
                const uint32_t t = cnt[b] & carry;
                cnt[b] ^= carry;
                carry = t;
            }
        }
    }
#pragma unroll
    for (int stride = NW / 2; stride >= 1; stride >>= 1) {
        if (warp >= stride && warp < 2 * stride) {
#pragma unroll
            for (int b = 0; b < 16; b++) part[warp][b][lane] = cnt[b];
        }
        __syncthreads();
        if (warp < stride) {
            uint32_t carry = 0;
#pragma unroll
            for (int b = 0; b < 16; b++) {
                const uint32_t x = part[warp + stride][b][lane], a = cnt[b];
                cnt[b] = a ^ x ^ carry;
                carry = (a & x) | (carry & (a ^ x));
            }
        }
        __syncthreads();
    }
    if (warp == 0) {
#pragma unroll
        for (int b = 0; b < 16; b++) if (b < nbits) bitcnt[(size_t)b * Wl + w] = cnt[b];
    }
}

int launch_site_counters(Ctx *c, int npairs, int nbits)
{
    constexpr int NW = 8;
    const int threads = 32 * NW, blocks = c->Wl / 32;
    switch (c->S) {
    case 2:  k_site_counters<2, NW><<<blocks, threads, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, c->d_pairs, npairs, nbits, c->d_bitcnt); break;
    case 4:  k_site_counters<4, NW><<<blocks, threads, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, c->d_pairs, npairs, nbits, c->d_bitcnt); break;
    case 20: k_site_counters<20, NW><<<blocks, threads, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, c->d_pairs, npairs, nbits, c->d_bitcnt); break;
    case 32: k_site_counters<32, NW><<<blocks, threads, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, c->d_pairs, npairs, nbits, c->d_bitcnt); break;
    default: set_error("unsupported state count"); return 1;
    }
    c->launches++;
    MPGPU_CUDA(cudaGetLastError());
    return 0;
}

// ptn_pars[i] = counter at expanded site ptn_site[i] (pllComputePatternParsimony :3384);
// sites outside this shard's word range contribute 0 (the caller sums over shards).
__global__ void k_gather_patterns(const uint32_t *__restrict__ bitcnt, int Wl, int64_t w0, int nbits,
                                  const int64_t *__restrict__ ptn_site, int count, uint16_t *__restrict__ ptn_pars)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int64_t site = ptn_site[i];
    const int64_t w = (site >> 5) - w0;
    uint32_t v = 0;
    if (site >= 0 && w >= 0 && w < Wl) {
        const int bit = (int)(site & 31);
        for (int b = 0; b < nbits; b++) v |= ((bitcnt[(size_t)b * Wl + w] >> bit) & 1u) << b;
    }
    ptn_pars[i] = (uint16_t)v;
}

int launch_gather_patterns(Ctx *c, int nbits, int count)
{
    if (count == 0) return 0;
    int threads = 128, blocks = (count + threads - 1) / threads;
    k_gather_patterns<<<blocks, threads, 0, c->stream>>>(c->d_bitcnt, c->Wl, c->w0, nbits, c->d_ptn_site, count, c->d_ptn);
    c->launches++;
    MPGPU_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mpgpu
