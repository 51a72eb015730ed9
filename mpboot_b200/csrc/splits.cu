// N2: support summarisation -- the split table of a weighted collection of trees.
//
// Replaces the arithmetic of MTreeSet::convertSplits (mtreeset.cpp:362-440) driven by IQTree::summarizeBootstrap
// (iqtree.cpp:3872-3989, 4020-4040): for every tree of the collection, in order, the bipartition ("split") behind every
// edge in the post-order MTree::convertSplits pushes them (mtree.cpp:917-939: an edge's split follows every split of the
// subtree below it), normalised like Split::shouldInvert (split.cpp:100-107: the side with fewer taxa; at a tie the side
// containing taxon 0), looked up in a hash set, its weight summed over the trees, new splits appended in first-seen order.
//
// On the device a tree arrives as the reverse-Polish token stream of that traversal: token t >= 0 pushes the leaf {t} and
// emits it, token -k (k >= 2) joins the k topmost sets and emits the union.  Every emit gets its taxon bit set (taxon i =
// bit i % 32 of word i / 32, like Split), is normalised and hashed; a lock-free table keyed by the 64-bit hash keeps, per
// distinct split, the smallest emit index (= first seen in the reference's order) and the weight sum; every emit is then
// compared word by word with its table entry's first emit, so a hash collision can never merge two splits (the call
// reports it and the host retries with another seed); the distinct splits come back in first-seen order.
#include "mpgpu_internal.h"

#include <cub/device/device_scan.cuh>

namespace mpgpu {

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// One block per tree.  bits[e][W] for the tree's emits e (global emit index = emit_begin[tree] + local index); the stack of
// the reverse-Polish machine holds emit indices (a set on the stack IS an earlier emit), kept in shared memory.
__global__ void k_split_emit(const int32_t *__restrict__ tokens, const int64_t *__restrict__ token_begin,
                             const int64_t *__restrict__ emit_begin, int ntaxa, int W,
                             uint32_t *__restrict__ bits, uint64_t *__restrict__ hash, uint64_t seed, int *__restrict__ err)
{
    extern __shared__ int32_t stack[];                 // up to ntaxa entries
    const int tree = blockIdx.x;
    const int64_t t0 = token_begin[tree], t1 = token_begin[tree + 1];
    const int64_t e0 = emit_begin[tree];
    __shared__ int sp_sh, bad;
    if (threadIdx.x == 0) { sp_sh = 0; bad = 0; }
    __syncthreads();
    for (int64_t t = t0; t < t1; t++) {
        const int tok = tokens[t];
        const int64_t e = e0 + (t - t0);
        uint32_t *dst = bits + (size_t)e * W;
        const int sp = sp_sh;
        if (tok >= 0) {
            if (tok >= ntaxa) { if (threadIdx.x == 0) bad = 1; }
            for (int w = threadIdx.x; w < W; w += blockDim.x) dst[w] = (w == (tok >> 5) && tok < ntaxa) ? (1u << (tok & 31)) : 0u;
            __syncthreads();
            if (threadIdx.x == 0) { if (sp < ntaxa) { stack[sp] = (int32_t)(t - t0); sp_sh = sp + 1; } else bad = 1; }
        } else {
            const int k = -tok;
            if (k < 2 || k > sp) { if (threadIdx.x == 0) bad = 1; __syncthreads(); break; }
            for (int w = threadIdx.x; w < W; w += blockDim.x) {
                uint32_t v = 0;
                for (int j = 0; j < k; j++) v |= bits[(size_t)(e0 + stack[sp - 1 - j]) * W + w];
                dst[w] = v;
            }
            __syncthreads();
            if (threadIdx.x == 0) { stack[sp - k] = (int32_t)(t - t0); sp_sh = sp - k + 1; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && bad) atomicExch(err, 1);
    __syncthreads();
    // normalise (Split::shouldInvert / invert) and hash: one warp per emit
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t tail = (ntaxa & 31) ? ((1u << (ntaxa & 31)) - 1u) : 0xFFFFFFFFu;
    for (int64_t e = e0 + warp; e < e0 + (t1 - t0); e += nwarps) {
        uint32_t *b = bits + (size_t)e * W;
        int cnt = 0;
        for (int w = lane; w < W; w += 32) cnt += __popc(b[w]);
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        const bool has0 = (b[0] & 1u) != 0;
        const bool inv = 2 * cnt > ntaxa || (2 * cnt == ntaxa && !has0);
        uint64_t h = 0;
        for (int w = lane; w < W; w += 32) {
            uint32_t v = b[w];
            if (inv) { v = ~v; if (w == W - 1) v &= tail; b[w] = v; }
            h ^= mix64(((uint64_t)(w + 1) << 32 | v) + seed);
        }
        for (int o = 16; o > 0; o >>= 1) h ^= __shfl_xor_sync(0xffffffffu, h, o);
        h = mix64(h ^ seed);
        if (h == 0) h = 1;                              // 0 marks an empty table slot
        if (lane == 0) hash[e] = h;
    }
}

// table insert: slot of every emit, first-seen emit per slot
__global__ void k_split_insert(const uint64_t *__restrict__ hash, int64_t nemit, unsigned long long *__restrict__ keys,
                               unsigned long long *__restrict__ first, uint32_t mask, int32_t *__restrict__ slot_of)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nemit) return;
    const unsigned long long h = hash[e];
    uint32_t s = (uint32_t)h & mask;
    for (;;) {
        const unsigned long long old = atomicCAS(&keys[s], 0ull, h);
        if (old == 0ull || old == h) break;
        s = (s + 1) & mask;
    }
    slot_of[e] = (int32_t)s;
    atomicMin(&first[s], (unsigned long long)e);
}

// exact comparison with the slot's first emit, weight sums, representative flags
__global__ void k_split_verify(const uint32_t *__restrict__ bits, int W, int64_t nemit, const int32_t *__restrict__ slot_of,
                               const unsigned long long *__restrict__ first, const int32_t *__restrict__ emit_weight,
                               int32_t *__restrict__ wsum, int32_t *__restrict__ is_rep, int *__restrict__ err)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nemit) return;
    const int s = slot_of[e];
    const int64_t f = (int64_t)first[s];
    if (f != e) {
        const uint32_t *a = bits + (size_t)e * W, *b = bits + (size_t)f * W;
        bool same = true;
        for (int w = 0; w < W; w++) same &= a[w] == b[w];
        if (!same) atomicExch(err, 2);                  // two different splits share a 64-bit hash: retry with another seed
    }
    is_rep[e] = f == e ? 1 : 0;
    const int wgt = emit_weight[e];
    if (wgt) atomicAdd(&wsum[s], wgt);
}

// distinct splits in first-seen order
__global__ void k_split_compact(const uint32_t *__restrict__ bits, int W, int64_t nemit, const int32_t *__restrict__ slot_of,
                                const unsigned long long *__restrict__ first, const int32_t *__restrict__ is_rep,
                                const int32_t *__restrict__ pos, const int32_t *__restrict__ wsum,
                                uint32_t *__restrict__ out_bits, int32_t *__restrict__ out_weight, int32_t *__restrict__ emit_unique)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nemit) return;
    const int s = slot_of[e];
    const int u = pos[first[s]];
    emit_unique[e] = u;
    if (is_rep[e]) {
        for (int w = 0; w < W; w++) out_bits[(size_t)u * W + w] = bits[(size_t)e * W + w];
        out_weight[u] = wsum[s];
    }
}

}  // namespace mpgpu

using namespace mpgpu;

extern "C" int mpgpu_split_table(mpgpu_ctx *c, int ntaxa, int ntrees, const int32_t *tokens, const int64_t *token_begin,
                                 const int32_t *tree_weight, int32_t *n_unique, uint32_t *split_bits, int32_t *split_weight,
                                 int32_t *emit_unique, int capacity)
{
    if (!c || !tokens || !token_begin || !tree_weight || !n_unique) { set_error("null argument"); return 1; }
    if (ntaxa < 3 || ntrees < 1) { set_error("need at least 3 taxa and one tree"); return 1; }
    MPGPU_CUDA(cudaSetDevice(c->device));
    const int W = (ntaxa + 31) / 32;
    const int64_t nemit = token_begin[ntrees] - token_begin[0];
    if (nemit < 1 || nemit > (int64_t)1 << 30) { set_error("bad token stream"); return 1; }
    for (int t = 0; t < ntrees; t++)
        if (token_begin[t + 1] < token_begin[t]) { set_error("token_begin must be non-decreasing"); return 1; }
    std::vector<int64_t> ebegin(ntrees + 1);
    std::vector<int32_t> eweight((size_t)nemit);
    for (int t = 0; t <= ntrees; t++) ebegin[t] = token_begin[t] - token_begin[0];
    for (int t = 0; t < ntrees; t++) for (int64_t e = ebegin[t]; e < ebegin[t + 1]; e++) eweight[e] = tree_weight[t];
    uint32_t tsize = 1024;
    while ((int64_t)tsize < 2 * nemit) tsize <<= 1;
    int32_t *d_tok = nullptr, *d_ew = nullptr, *d_slot = nullptr, *d_wsum = nullptr, *d_rep = nullptr, *d_pos = nullptr, *d_ow = nullptr, *d_eu = nullptr;
    int64_t *d_tb = nullptr, *d_eb = nullptr;
    uint32_t *d_bits = nullptr, *d_ob = nullptr;
    uint64_t *d_hash = nullptr;
    unsigned long long *d_keys = nullptr, *d_first = nullptr;
    int *d_err = nullptr;
    void *d_scan = nullptr; size_t scan_bytes = 0;
    int rc = 0;
    auto freeall = [&]() {
        void *ptrs[] = {d_tok, d_ew, d_slot, d_wsum, d_rep, d_pos, d_ow, d_eu, d_tb, d_eb, d_bits, d_ob, d_hash, d_keys, d_first, d_err, d_scan};
        for (void *p : ptrs) if (p) cudaFree(p);
    };
#define SPL_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { rc = cuda_fail(e__, #expr); freeall(); return rc; } } while (0)
    SPL_CUDA(cudaMalloc((void **)&d_tok, (size_t)nemit * 4));
    SPL_CUDA(cudaMalloc((void **)&d_ew, (size_t)nemit * 4));
    SPL_CUDA(cudaMalloc((void **)&d_slot, (size_t)nemit * 4));
    SPL_CUDA(cudaMalloc((void **)&d_rep, (size_t)nemit * 4));
    SPL_CUDA(cudaMalloc((void **)&d_pos, (size_t)nemit * 4));
    SPL_CUDA(cudaMalloc((void **)&d_eu, (size_t)nemit * 4));
    SPL_CUDA(cudaMalloc((void **)&d_tb, (size_t)(ntrees + 1) * 8));
    SPL_CUDA(cudaMalloc((void **)&d_eb, (size_t)(ntrees + 1) * 8));
    SPL_CUDA(cudaMalloc((void **)&d_bits, (size_t)nemit * W * 4));
    SPL_CUDA(cudaMalloc((void **)&d_hash, (size_t)nemit * 8));
    SPL_CUDA(cudaMalloc((void **)&d_keys, (size_t)tsize * 8));
    SPL_CUDA(cudaMalloc((void **)&d_first, (size_t)tsize * 8));
    SPL_CUDA(cudaMalloc((void **)&d_wsum, (size_t)tsize * 4));
    SPL_CUDA(cudaMalloc((void **)&d_err, sizeof(int)));
    SPL_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_rep, d_pos, (int)nemit, c->stream));
    SPL_CUDA(cudaMalloc(&d_scan, scan_bytes ? scan_bytes : 16));
    SPL_CUDA(cudaMemcpyAsync(d_tok, tokens + token_begin[0], (size_t)nemit * 4, cudaMemcpyHostToDevice, c->stream));
    SPL_CUDA(cudaMemcpyAsync(d_ew, eweight.data(), (size_t)nemit * 4, cudaMemcpyHostToDevice, c->stream));
    SPL_CUDA(cudaMemcpyAsync(d_tb, ebegin.data(), (size_t)(ntrees + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    SPL_CUDA(cudaMemcpyAsync(d_eb, ebegin.data(), (size_t)(ntrees + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    const int threads = 128;
    const int blocks = (int)((nemit + threads - 1) / threads);
    int err = 0;
    int nu = 0;
    for (int attempt = 0; attempt < 4; attempt++) {
        const uint64_t seed = 0x9E3779B97F4A7C15ULL * (uint64_t)(attempt + 1);
        SPL_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), c->stream));
        SPL_CUDA(cudaMemsetAsync(d_keys, 0, (size_t)tsize * 8, c->stream));
        SPL_CUDA(cudaMemsetAsync(d_first, 0xff, (size_t)tsize * 8, c->stream));
        SPL_CUDA(cudaMemsetAsync(d_wsum, 0, (size_t)tsize * 4, c->stream));
        k_split_emit<<<ntrees, 128, (size_t)ntaxa * sizeof(int32_t), c->stream>>>(d_tok, d_tb, d_eb, ntaxa, W, d_bits, d_hash, seed, d_err);
        k_split_insert<<<blocks, threads, 0, c->stream>>>(d_hash, nemit, d_keys, d_first, tsize - 1, d_slot);
        k_split_verify<<<blocks, threads, 0, c->stream>>>(d_bits, W, nemit, d_slot, d_first, d_ew, d_wsum, d_rep, d_err);
        c->launches += 3;
        SPL_CUDA(cudaGetLastError());
        SPL_CUDA(cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        SPL_CUDA(cudaStreamSynchronize(c->stream));
        if (err != 2) break;                            // 2 = hash collision between different splits: another seed
    }
    if (err == 1) { freeall(); set_error("malformed token stream (stack underflow, bad taxon or bad join count)"); return 1; }
    if (err == 2) { freeall(); set_error("split hashing kept colliding"); return 1; }
    SPL_CUDA(cub::DeviceScan::ExclusiveSum(d_scan, scan_bytes, d_rep, d_pos, (int)nemit, c->stream));
    {
        int last_pos = 0, last_rep = 0;
        SPL_CUDA(cudaMemcpyAsync(&last_pos, d_pos + (nemit - 1), 4, cudaMemcpyDeviceToHost, c->stream));
        SPL_CUDA(cudaMemcpyAsync(&last_rep, d_rep + (nemit - 1), 4, cudaMemcpyDeviceToHost, c->stream));
        SPL_CUDA(cudaStreamSynchronize(c->stream));
        nu = last_pos + last_rep;
    }
    *n_unique = nu;
    if (split_bits || split_weight) {
        if (nu > capacity) { freeall(); set_error("split capacity too small"); return 1; }
    }
    SPL_CUDA(cudaMalloc((void **)&d_ob, (size_t)(nu > 0 ? nu : 1) * W * 4));
    SPL_CUDA(cudaMalloc((void **)&d_ow, (size_t)(nu > 0 ? nu : 1) * 4));
    k_split_compact<<<blocks, threads, 0, c->stream>>>(d_bits, W, nemit, d_slot, d_first, d_rep, d_pos, d_wsum, d_ob, d_ow, d_eu);
    c->launches++;
    SPL_CUDA(cudaGetLastError());
    if (split_bits) SPL_CUDA(cudaMemcpyAsync(split_bits, d_ob, (size_t)nu * W * 4, cudaMemcpyDeviceToHost, c->stream));
    if (split_weight) SPL_CUDA(cudaMemcpyAsync(split_weight, d_ow, (size_t)nu * 4, cudaMemcpyDeviceToHost, c->stream));
    if (emit_unique) SPL_CUDA(cudaMemcpyAsync(emit_unique, d_eu, (size_t)nemit * 4, cudaMemcpyDeviceToHost, c->stream));
    SPL_CUDA(cudaStreamSynchronize(c->stream));
#undef SPL_CUDA
    freeall();
    return 0;
}
