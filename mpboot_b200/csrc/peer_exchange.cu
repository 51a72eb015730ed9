// The exchange step of the pattern-sharded path, inside the library and on the device.
//
// Every score of the path is a sum over site words, so the shards only ever exchange small int32 vectors (per-insertion
// counts of a scan batch, view counts after a tree change, per-branch counts of a stepwise-addition step, REPS rows)
// that have to be summed over the shards in place.  NCCL does this in ~20-40 us per call for a few tens of KB, launched
// from the host; at 160 us per C2 sweep that was 12-18 % of the step at 4-8 GPUs (SCALE_r01).  Here it is ONE small
// kernel on the context's stream, a one-shot all-reduce over NVLink peer memory:
//
//   every rank owns an exchange region  slots[2][R][cap] int32 + flags[2][R][kPeerBlocks] u32  (R = shards),
//   mapped into every other rank through CUDA IPC (mpgpu_peer_prepare / mpgpu_peer_connect);
//   block b of rank r  1. stores its slice of the vector into slots[e & 1][r] of EVERY rank (P2P stores through NVSwitch),
//                      2. fences (system scope) and writes the epoch e into flags[e & 1][r][b] of every rank,
//                      3. waits until its own flags[e & 1][q][b] == e for every q,
//                      4. sums slots[e & 1][q][slice] over q into the vector.
//   Two parities: a rank can run at most one epoch ahead of a peer (it needs the peer's flag of epoch e to leave e, and
//   the peer raises it only after it finished summing e - 1), so epoch e + 1 never overwrites data a peer still reads.
//   Blocks never wait for another block of their own GPU, only for the same block of a peer, so the kernel cannot
//   deadlock on residency; a bounded spin (2 s) turns a missing peer into an error instead of a hung device.
#include "mpgpu_internal.h"

#include <cstring>

namespace mpgpu {

struct PeerArgs {
    int32_t *slots[kMaxPeers];
    uint32_t *flags[kMaxPeers];
    int rank, nranks;
    uint32_t epoch;
    unsigned long long cap;        // ints per slot
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// count is a multiple of 4 (the caller pads: the buffers are sized in whole int4), slices are whole int4
__global__ void __launch_bounds__(256) k_peer_allreduce(int32_t *__restrict__ buf, int count4, PeerArgs a, int *__restrict__ err)
{
    const int b = blockIdx.x, nb = gridDim.x;
    const int per = (count4 + nb - 1) / nb;
    const int lo = b * per, hi = min(count4, lo + per);
    const unsigned par = a.epoch & 1u;
    const size_t slot_off = ((size_t)par * a.nranks + a.rank) * a.cap;          // my slot in everyone's region
    const int4 *src = reinterpret_cast<const int4 *>(buf);
    // 1. push
    for (int p = 0; p < a.nranks; p++) {
        int4 *dst = reinterpret_cast<int4 *>(a.slots[(a.rank + p) % a.nranks] + slot_off);     // start with myself, then round the ring
        for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    // 2. signal, 3. wait
    if (threadIdx.x < a.nranks) {
        const int q = threadIdx.x;
        st_release_sys(a.flags[q] + ((size_t)par * a.nranks + a.rank) * kPeerBlocks + b, a.epoch);
        const uint32_t *mine = a.flags[a.rank] + ((size_t)par * a.nranks + q) * kPeerBlocks + b;
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(mine) != a.epoch) {
            if (global_timer_ns() - t0 > 2000000000ull) { atomicExch(err, 1 + q); break; }
        }
    }
    __syncthreads();
    // 4. sum
    const int32_t *local = a.slots[a.rank] + (size_t)par * a.nranks * a.cap;
    int4 *out = reinterpret_cast<int4 *>(buf);
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        // .cg: the slots are written by the peers over NVLink; never take them from this SM's L1 (a line of epoch e - 2)
        int4 s = __ldcg(reinterpret_cast<const int4 *>(local + (size_t)4 * i));
        for (int q = 1; q < a.nranks; q++) {
            const int4 v = __ldcg(reinterpret_cast<const int4 *>(local + (size_t)q * a.cap + (size_t)4 * i));
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        out[i] = s;
    }
}

void peer_free(Ctx *c)
{
    PeerExchange &px = c->peer;
    for (int q = 0; q < kMaxPeers; q++)
        if (px.mapped[q]) { cudaIpcCloseMemHandle(px.mapped[q]); px.mapped[q] = nullptr; }
    if (px.region) cudaFree(px.region);
    if (px.d_err) cudaFree(px.d_err);
    if (px.d_tmp) cudaFree(px.d_tmp);
    px = PeerExchange();
}

// in-place sum over the shards of dev_i32[0 .. count), on the context's stream; nothing is waited for
int peer_allreduce(Ctx *c, void *dev_i32, int64_t count)
{
    PeerExchange &px = c->peer;
    if (!px.ready) { set_error("peer exchange not connected"); return 1; }
    int32_t *buf = static_cast<int32_t *>(dev_i32);
    PeerArgs a;
    for (int q = 0; q < kMaxPeers; q++) { a.slots[q] = px.slots[q]; a.flags[q] = px.flags[q]; }
    a.rank = c->xrank(); a.nranks = c->xcount(); a.cap = (unsigned long long)px.cap;
    int64_t done = 0;
    while (done < count) {
        const int64_t piece = std::min<int64_t>(count - done, (int64_t)px.cap);
        int32_t *p = buf + done;
        // the library's exchanged vectors start on a 16-byte boundary and have at least 3 int32 of slack behind them
        // (ensure() over-allocates; d_vcount / d_scalar are padded), so a count that is not a multiple of 4 simply
        // carries up to 3 meaningless elements along; only an unaligned start goes through the scratch vector
        const bool aligned = ((uintptr_t)p & 15) == 0 && ((piece & 3) == 0 || done + piece == count);
        int32_t *work = p;
        if (!aligned) {                           // odd tail or offset: through an aligned scratch vector (rare: the library's own vectors are padded)
            const size_t need = (size_t)((piece + 3) & ~(int64_t)3);
            if (int rc = ensure(px.d_tmp, px.tmp_cap, need)) return rc;
            MPGPU_CUDA(cudaMemsetAsync(px.d_tmp, 0, need * sizeof(int32_t), c->stream));
            MPGPU_CUDA(cudaMemcpyAsync(px.d_tmp, p, (size_t)piece * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
            work = px.d_tmp;
        }
        const int count4 = (int)((piece + 3) / 4);
        int blocks = (count4 + 511) / 512;         // >= 2 int4 per thread
        if (blocks < 1) blocks = 1;
        if (blocks > kPeerBlocks) blocks = kPeerBlocks;
        a.epoch = ++px.epoch;
        k_peer_allreduce<<<blocks, 256, 0, c->stream>>>(work, count4, a, px.d_err);
        c->launches++;
        MPGPU_CUDA(cudaGetLastError());
        if (!aligned) MPGPU_CUDA(cudaMemcpyAsync(p, px.d_tmp, (size_t)piece * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
        done += piece;
        px.calls++; px.elements += piece;
    }
    return 0;
}

}  // namespace mpgpu

using namespace mpgpu;

extern "C" {

int mpgpu_peer_prepare(mpgpu_ctx *c, int64_t capacity, void *handle_out)
{
    if (!c || !handle_out) { set_error("null argument"); return 1; }
    if (c->xcount() < 2) { set_error("peer exchange is for sharded contexts (pattern shards at mpgpu_create, or mpgpu_set_replicate_shards)"); return 1; }
    if (c->xcount() > kMaxPeers) { set_error("peer exchange supports up to 8 shards (one NVSwitch box)"); return 1; }
    MPGPU_CUDA(cudaSetDevice(c->device));
    peer_free(c);
    PeerExchange &px = c->peer;
    if (capacity < 1024) capacity = 1024;
    px.cap = (size_t)((capacity + 3) & ~(int64_t)3);
    const size_t R = (size_t)c->xcount();
    const size_t slot_bytes = 2 * R * px.cap * sizeof(int32_t);
    const size_t flag_bytes = 2 * R * kPeerBlocks * sizeof(uint32_t);
    px.bytes = slot_bytes + flag_bytes;
    MPGPU_CUDA(cudaMalloc(&px.region, px.bytes));
    MPGPU_CUDA(cudaMemset(px.region, 0, px.bytes));
    MPGPU_CUDA(cudaMalloc((void **)&px.d_err, sizeof(int)));
    MPGPU_CUDA(cudaMemset(px.d_err, 0, sizeof(int)));
    MPGPU_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    MPGPU_CUDA(cudaIpcGetMemHandle(&h, px.region));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}

int mpgpu_peer_connect(mpgpu_ctx *c, const void *handles)
{
    if (!c || !handles) { set_error("null argument"); return 1; }
    PeerExchange &px = c->peer;
    if (!px.region) { set_error("mpgpu_peer_prepare first"); return 1; }
    MPGPU_CUDA(cudaSetDevice(c->device));
    const size_t R = (size_t)c->xcount();
    const size_t slot_bytes = 2 * R * px.cap * sizeof(int32_t);
    for (int q = 0; q < c->xcount(); q++) {
        void *base = nullptr;
        if (q == c->xrank()) base = px.region;
        else {
            cudaIpcMemHandle_t h;
            memcpy(&h, static_cast<const char *>(handles) + (size_t)q * sizeof(h), sizeof(h));
            MPGPU_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
            px.mapped[q] = base;
        }
        px.slots[q] = static_cast<int32_t *>(base);
        px.flags[q] = reinterpret_cast<uint32_t *>(static_cast<char *>(base) + slot_bytes);
    }
    px.epoch = 0;
    px.ready = true;
    if (c->tree_set && !c->lens_valid) c->views_stale = true;      // partial counts of a resident tree get completed on first use
    return 0;
}

int mpgpu_peer_stats(mpgpu_ctx *c, int64_t *calls, int64_t *elements, int *error)
{
    if (!c) { set_error("null argument"); return 1; }
    if (calls) *calls = c->peer.calls;
    if (elements) *elements = c->peer.elements;
    if (error) {
        *error = 0;
        if (c->peer.d_err) {
            MPGPU_CUDA(cudaSetDevice(c->device));
            MPGPU_CUDA(cudaMemcpy(error, c->peer.d_err, sizeof(int), cudaMemcpyDeviceToHost));
        }
    }
    return 0;
}

}  // extern "C"
