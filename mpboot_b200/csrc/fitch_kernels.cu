// Hand-written sm_100a kernels of the bit-plane Fitch path.
//
// Data layout in HBM: one "view" = the Fitch state sets of a directed
// subtree, stored as S bit planes of Wl 32-bit words, views[vid][state][word]; bit j of word i
// = "expanded site 32*(w0+i)+j may be in that state" (same meaning as the reference's
// parsVect, sprparsimony.cpp:2870-2960; padding bits are 1 in every state so they never score).
// Wl is a multiple of 128 words, so every plane row starts 512-byte aligned and a warp always
// moves whole 128-byte lines (32-bit lanes) or 512-byte runs (128-bit lanes).
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

static uint32_t g_mask_table[256];
const uint32_t *state_mask_table(int datatype, int *ncodes, int *undetermined)
{
    int nc = 0, und = 0;
    switch (datatype) {
    case MPGPU_BINARY_DATA: nc = 4;  und = 3;  break;
    case MPGPU_DNA_DATA:    nc = 16; und = 15; break;
    case MPGPU_AA_DATA:     nc = 23; und = 22; break;
    case MPGPU_GENERIC_32:  nc = 33; und = 32; break;
    default: return nullptr;
    }
    for (int i = 0; i < nc; i++) g_mask_table[i] = code_mask(datatype, (uint32_t)i);
    if (ncodes) *ncodes = nc;
    if (undetermined) *undetermined = und;
    return g_mask_table;
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
    uint32_t *dst = views + (size_t)tip * view_stride + w;       // tips are views 0..n-1
    for (int s = 0; s < S; s++) {
        uint32_t word = 0;
#pragma unroll
        for (int j = 0; j < 32; j++) word |= ((mask[j] >> s) & 1u) << j;
        dst[(size_t)s * Wl] = word;
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
// 128-bit helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld4(const uint32_t *p) { return *reinterpret_cast<const uint4 *>(p); }
__device__ __forceinline__ void st4(uint32_t *p, uint4 v) { *reinterpret_cast<uint4 *>(p) = v; }
__device__ __forceinline__ uint4 and4(uint4 a, uint4 b) { return make_uint4(a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w); }
__device__ __forceinline__ uint4 or4(uint4 a, uint4 b) { return make_uint4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w); }
// (a & b) | (~n & (a | b)) : the Fitch set for one state given the "some state intersects" mask n
__device__ __forceinline__ uint32_t fitch1(uint32_t a, uint32_t b, uint32_t n) { return (a & b) | (~n & (a | b)); }
__device__ __forceinline__ uint4 fitch4(uint4 a, uint4 b, uint4 n)
{
    return make_uint4(fitch1(a.x, b.x, n.x), fitch1(a.y, b.y, n.y), fitch1(a.z, b.z, n.z), fitch1(a.w, b.w, n.w));
}
__device__ __forceinline__ int popc_not4(uint4 n) { return __popc(~n.x) + __popc(~n.y) + __popc(~n.z) + __popc(~n.w); }

// ------------------------------------------------------------------------------------------
// R3: one level of the directed-view schedule.  One warp per (triple, 128-word chunk);
// lanes hold 128 bits of every plane.  dst = fitch(a, b), count[dst] += popc(t_N).
// ------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(128) k_fitch_level(uint32_t *__restrict__ views, size_t view_stride, int Wl,
                                                     const Triple *__restrict__ triples, uint32_t *__restrict__ vcount)
{
    const int lane = threadIdx.x & 31;
    const int chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (chunk * kWordPad >= Wl) return;
    const Triple t = triples[blockIdx.y];
    const size_t off = (size_t)chunk * kWordPad + lane * 4;
    const uint32_t *A = views + (size_t)t.a * view_stride + off;
    const uint32_t *B = views + (size_t)t.b * view_stride + off;
    uint32_t *D = views + (size_t)t.dst * view_stride + off;

    uint4 any = make_uint4(0, 0, 0, 0);
    if (S <= 4) {
        uint4 a[S <= 4 ? S : 1], b[S <= 4 ? S : 1];
#pragma unroll
        for (int k = 0; k < S; k++) { a[k] = ld4(A + (size_t)k * Wl); b[k] = ld4(B + (size_t)k * Wl); }
#pragma unroll
        for (int k = 0; k < S; k++) any = or4(any, and4(a[k], b[k]));
#pragma unroll
        for (int k = 0; k < S; k++) st4(D + (size_t)k * Wl, fitch4(a[k], b[k], any));
    } else {
#pragma unroll 4
        for (int k = 0; k < S; k++) any = or4(any, and4(ld4(A + (size_t)k * Wl), ld4(B + (size_t)k * Wl)));
#pragma unroll 4
        for (int k = 0; k < S; k++)
            st4(D + (size_t)k * Wl, fitch4(ld4(A + (size_t)k * Wl), ld4(B + (size_t)k * Wl), any));
    }
    int cnt = popc_not4(any);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0 && cnt) atomicAdd(&vcount[t.dst], (uint32_t)cnt);
}

int launch_level(Ctx *c, const Triple *d_triples, int ntriples)
{
    if (ntriples == 0) return 0;
    int chunks = c->Wl / kWordPad;
    dim3 grid((chunks + 3) / 4, ntriples);
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
// R4: mismatch count across one edge: popc(~OR_k(A_k & B_k))
// ------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(128) k_edge_mismatch(const uint32_t *__restrict__ views, size_t view_stride, int Wl,
                                                       int vidA, int vidB, uint32_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int chunk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (chunk * kWordPad >= Wl) return;
    const size_t off = (size_t)chunk * kWordPad + lane * 4;
    const uint32_t *A = views + (size_t)vidA * view_stride + off;
    const uint32_t *B = views + (size_t)vidB * view_stride + off;
    uint4 any = make_uint4(0, 0, 0, 0);
#pragma unroll 4
    for (int k = 0; k < S; k++) any = or4(any, and4(ld4(A + (size_t)k * Wl), ld4(B + (size_t)k * Wl)));
    int cnt = popc_not4(any);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0 && cnt) atomicAdd(out, (uint32_t)cnt);
}

int launch_edge_mismatch(Ctx *c, int vidA, int vidB, uint32_t *d_out)
{
    int chunks = c->Wl / kWordPad;
    dim3 grid((chunks + 3) / 4);
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
// One warp owns (task, chunk of 32 words); lane l owns word l of the chunk for every plane
// and every stack slot, so the stack needs no synchronisation at all.
// Work order: all tasks of chunk 0, then chunk 1, ... so that concurrently resident warps
// touch the same few word columns of every view and the views are served from L2.
// ------------------------------------------------------------------------------------------
template <int S>
__device__ __forceinline__ uint32_t any_and(const uint32_t (&a)[S], const uint32_t (&b)[S])
{
    uint32_t n = 0;
#pragma unroll
    for (int k = 0; k < S; k++) n |= a[k] & b[k];
    return n;
}

template <int S>
__global__ void k_spr_scan(const uint32_t *__restrict__ views, size_t view_stride, int Wl,
                           const ScanTask *__restrict__ tasks, int ntasks,
                           const ScanOp *__restrict__ ops, int nslots, int32_t *__restrict__ out)
{
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const long long gw = (long long)blockIdx.x * wpb + warp;
    const int nchunks = Wl / kChunkWords;
    if (gw >= (long long)ntasks * nchunks) return;
    const int chunk = (int)(gw / ntasks);
    const int ti = (int)(gw % ntasks);
    const ScanTask task = tasks[ti];
    const size_t off = (size_t)chunk * kChunkWords + lane;
    uint32_t *stack = smem + (size_t)warp * nslots * S * 32 + lane;     // [slot][k][lane]

    uint32_t Sv[S];
    {
        const uint32_t *p = views + (size_t)task.s_vid * view_stride + off;
#pragma unroll
        for (int k = 0; k < S; k++) Sv[k] = __ldg(p + (size_t)k * Wl);
    }
    if (task.base_out >= 0) {
        const uint32_t *p1 = views + (size_t)task.d1 * view_stride + off;
        const uint32_t *p2 = views + (size_t)task.d2 * view_stride + off;
        uint32_t n = 0;
#pragma unroll
        for (int k = 0; k < S; k++) n |= __ldg(p1 + (size_t)k * Wl) & __ldg(p2 + (size_t)k * Wl);
        int cnt = __reduce_add_sync(0xffffffffu, __popc(~n));
        if (lane == 0 && cnt) atomicAdd(&out[task.base_out], cnt);
    }

    for (int oi = task.op_begin; oi < task.op_end; oi++) {
        const int4 o0 = __ldg(reinterpret_cast<const int4 *>(ops + oi));
        const int4 o1 = __ldg(reinterpret_cast<const int4 *>(ops + oi) + 1);
        const int src = o0.x, c1 = o0.y, c2 = o0.z, out1 = o0.w;
        const int out2 = o1.x, dst1 = o1.y, dst2 = o1.z;

        uint32_t U[S], A[S], B[S];
        {
            const uint32_t *pa = views + (size_t)c1 * view_stride + off;
            const uint32_t *pb = views + (size_t)c2 * view_stride + off;
#pragma unroll
            for (int k = 0; k < S; k++) { A[k] = __ldg(pa + (size_t)k * Wl); B[k] = __ldg(pb + (size_t)k * Wl); }
            if (src >= 0) {
                const uint32_t *ps = stack + (size_t)src * S * 32;
#pragma unroll
                for (int k = 0; k < S; k++) U[k] = ps[k * 32];
            } else {
                const uint32_t *pu = views + (size_t)(~src) * view_stride + off;
#pragma unroll
                for (int k = 0; k < S; k++) U[k] = __ldg(pu + (size_t)k * Wl);
            }
        }
        // child 1 (sibling view B)
        if (out1 >= 0 || dst1 >= 0) {
            const uint32_t n = any_and<S>(U, B);
            uint32_t U1[S];
#pragma unroll
            for (int k = 0; k < S; k++) U1[k] = fitch1(U[k], B[k], n);
            if (dst1 >= 0) {
                uint32_t *pd = stack + (size_t)dst1 * S * 32;
#pragma unroll
                for (int k = 0; k < S; k++) pd[k * 32] = U1[k];
            }
            if (out1 >= 0) {
                const uint32_t m = any_and<S>(U1, A);
                uint32_t z = 0;
#pragma unroll
                for (int k = 0; k < S; k++) z |= fitch1(U1[k], A[k], m) & Sv[k];
                int cnt = __reduce_add_sync(0xffffffffu, __popc(~z));
                if (lane == 0 && cnt) atomicAdd(&out[out1], cnt);
            }
        }
        // child 2 (sibling view A)
        if (out2 >= 0 || dst2 >= 0) {
            const uint32_t n = any_and<S>(U, A);
            uint32_t U2[S];
#pragma unroll
            for (int k = 0; k < S; k++) U2[k] = fitch1(U[k], A[k], n);
            if (dst2 >= 0) {
                uint32_t *pd = stack + (size_t)dst2 * S * 32;
#pragma unroll
                for (int k = 0; k < S; k++) pd[k * 32] = U2[k];
            }
            if (out2 >= 0) {
                const uint32_t m = any_and<S>(U2, B);
                uint32_t z = 0;
#pragma unroll
                for (int k = 0; k < S; k++) z |= fitch1(U2[k], B[k], m) & Sv[k];
                int cnt = __reduce_add_sync(0xffffffffu, __popc(~z));
                if (lane == 0 && cnt) atomicAdd(&out[out2], cnt);
            }
        }
    }
}

template <int S>
static int launch_scan_t(Ctx *c, int ntasks, int nslots)
{
    const size_t per_warp = (size_t)(nslots > 0 ? nslots : 1) * S * 32 * sizeof(uint32_t);
    int wpb = 8;
    const size_t budget = 96 * 1024;
    while (wpb > 1 && per_warp * wpb > budget) wpb >>= 1;
    const size_t smem = per_warp * wpb;
    if (smem > 200 * 1024) { set_error("scan stack does not fit in shared memory"); return 1; }
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        MPGPU_CUDA(cudaFuncSetAttribute(k_spr_scan<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
        configured = 200 * 1024;
    }
    const long long warps = (long long)ntasks * (c->Wl / kChunkWords);
    const long long blocks = (warps + wpb - 1) / wpb;
    if (blocks > 0x7fffffffLL) { set_error("scan grid too large"); return 1; }
    k_spr_scan<S><<<(unsigned)blocks, wpb * 32, smem, c->stream>>>(c->d_views, c->view_stride, c->Wl, c->d_tasks, ntasks,
                                                                   c->d_ops, nslots > 0 ? nslots : 1, c->d_counts);
    c->launches++;
    MPGPU_CUDA(cudaGetLastError());
    return 0;
}

int launch_scan(Ctx *c, int ntasks, int nslots)
{
    if (ntasks == 0) return 0;
    switch (c->S) {
    case 2:  return launch_scan_t<2>(c, ntasks, nslots);
    case 4:  return launch_scan_t<4>(c, ntasks, nslots);
    case 20: return launch_scan_t<20>(c, ntasks, nslots);
    case 32: return launch_scan_t<32>(c, ntasks, nslots);
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
template <int S>
__global__ void k_site_counters(const uint32_t *__restrict__ views, size_t view_stride, int Wl,
                                const int32_t *__restrict__ pairs, int npairs, int nbits,
                                uint32_t *__restrict__ bitcnt)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= Wl) return;
    uint32_t cnt[16];
#pragma unroll
    for (int b = 0; b < 16; b++) cnt[b] = 0;
    for (int i = 0; i < npairs; i++) {
        const uint32_t *A = views + (size_t)pairs[2 * i] * view_stride + w;
        const uint32_t *B = views + (size_t)pairs[2 * i + 1] * view_stride + w;
        uint32_t n = 0;
#pragma unroll 4
        for (int k = 0; k < S; k++) n |= __ldg(A + (size_t)k * Wl) & __ldg(B + (size_t)k * Wl);
        uint32_t carry = ~n;
#pragma unroll
        for (int b = 0; b < 16; b++) {
            const uint32_t t = cnt[b] & carry;
            cnt[b] ^= carry;
            carry = t;
        }
    }
    for (int b = 0; b < nbits; b++) bitcnt[(size_t)b * Wl + w] = cnt[b];
}

int launch_site_counters(Ctx *c, int npairs, int nbits)
{
    int threads = 128, blocks = (c->Wl + threads - 1) / threads;
    switch (c->S) {
    case 2:  k_site_counters<2><<<blocks, threads, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, c->d_pairs, npairs, nbits, c->d_bitcnt); break;
    case 4:  k_site_counters<4><<<blocks, threads, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, c->d_pairs, npairs, nbits, c->d_bitcnt); break;
    case 20: k_site_counters<20><<<blocks, threads, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, c->d_pairs, npairs, nbits, c->d_bitcnt); break;
    case 32: k_site_counters<32><<<blocks, threads, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, c->d_pairs, npairs, nbits, c->d_bitcnt); break;
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
