// Replicate scoring (REPS) kernels of the -bb path: R8 of DESIGN.md.
//
// Reference: IQTree::saveCurrentTree, iqtree.cpp:3411-3449 --
//     rell[b] = - sum_seg ( ( sum_{ptn in seg} pattern_pars[ptn] * boot_sample[b][ptn] ) mod 2^16 )
// on u16 SIMD lanes (vectorclass/vectori256.h:1726-1740: lane products, lane sums and the
// horizontal add all wrap at 16 bits).
//
// Here every per-pattern score vector is a signed sum of 1-bit rows:
//   current tree      c_T   = sum_bit 2^bit * plane_bit            (bit-sliced site counters)
//   candidate (p->e)  c     = c_T - mis(edge p|q) + delta_e        (Fitch length is root-invariant per site)
// so REPS is  X[row][b] = sum_ptn bit[row][ptn] * w[b][ptn]  -- a dense {0,1} x u8 -> s32 contraction
// (rows x patterns x replicates) -- followed by a tiny linear combine.  The contraction runs on
// the int8 tensor cores (tcgen05.mma kind::i8, accumulators in TMEM, replicate weights staged by
// TMA); patterns whose arithmetic the u8 tensor path cannot represent exactly -- a replicate
// weight above 255, or a segment whose 16-bit wrap cannot be ruled out -- are "exceptions" and
// go through an exact CUDA-core kernel into their own column group, where the mod-2^16 is
// applied after the combine.
#include <type_traits>

#include "mpgpu_internal.h"

#include <cuda.h>

namespace mpgpu {

// ------------------------------------------------------------------------------------------
// Replicate weights.  boot16 [B][stride] u16 (boot_samples_pars, iqtree.cpp:220-233) is kept on
// the device pattern-major:   w16T[ptn][Bpad]   (exact weights; exception path, wrap check)
// and, for the tensor kernel: w8[Bpad][Kpad] u8, K-major, exception patterns and padding = 0.
// ------------------------------------------------------------------------------------------
// 32x32 tile transpose; heavy[ptn] = 1 when some replicate weight of the pattern exceeds 255
__global__ void k_transpose_boot(const uint16_t *__restrict__ boot16, int B, int stride, int upper, int Bpad,
                                 uint16_t *__restrict__ w16T, uint8_t *__restrict__ heavy)
{
    __shared__ uint16_t tile[32][33];
    const int p0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int b = b0 + i, p = p0 + threadIdx.x;
        tile[i][threadIdx.x] = (b < B && p < upper) ? boot16[(size_t)b * stride + p] : (uint16_t)0;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int p = p0 + i, b = b0 + threadIdx.x;
        if (p < upper && b < Bpad) {
            const uint16_t v = tile[threadIdx.x][i];
            w16T[(size_t)p * Bpad + b] = v;
            if (v > 255) heavy[p] = 1;
        }
    }
}

__global__ void k_build_w8(const uint16_t *__restrict__ w16T, int upper, const uint8_t *__restrict__ is_exc,
                           int Bpad, int Kpad, uint8_t *__restrict__ w8)
{
    __shared__ uint8_t tile[32][33];
    const int p0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int p = p0 + i, b = b0 + threadIdx.x;
        uint32_t v = 0;
        if (p < upper && !is_exc[p]) v = w16T[(size_t)p * Bpad + b];
        tile[i][threadIdx.x] = (uint8_t)v;              // <= 255 by construction: heavier patterns are exceptions
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int b = b0 + i, p = p0 + threadIdx.x;
        if (p < Kpad) w8[(size_t)b * Kpad + p] = tile[threadIdx.x][i];
    }
}

int launch_transpose_boot(Ctx *c, const uint16_t *d_boot16, int stride, uint8_t *d_heavy)
{
    Reps &r = c->reps;
    dim3 grid((r.upper + 31) / 32, r.Bpad / 32), block(32, 8);
    if (r.upper == 0) return 0;
    k_transpose_boot<<<grid, block, 0, c->stream>>>(d_boot16, r.B, stride, r.upper, r.Bpad, r.d_w16T, d_heavy);
    c->launches++;
    MPGPU_CUDA(cudaGetLastError());
    return 0;
}

int launch_build_w8(Ctx *c, const uint8_t *d_is_exc)
{
    Reps &r = c->reps;
    dim3 grid(r.Kpad / 32, r.Bpad / 32), block(32, 8);
    k_build_w8<<<grid, block, 0, c->stream>>>(r.d_w16T, r.upper, d_is_exc, r.Bpad, r.Kpad, r.d_w8);
    c->launches++;
    MPGPU_CUDA(cudaGetLastError());
    return 0;
}

// Wrap check for the tree the rows are built from.  Every candidate scored against tree T has
// per-pattern score <= c_T + 1 (one SPR adds at most one step per site), so a segment whose
//     sum_{ptn in seg} (c_T[ptn] + 1) * w[b][ptn]  <  2^16   for every replicate b
// cannot wrap in the reference's u16 lanes for T or any of its candidates: its mod-2^16 is the
// identity and it may stay in the tensor path's bulk group.  segmax[seg] = max over b of that
// sum over THIS shard's patterns [p_lo, p_hi) (clipped); shards add their maxima (an upper bound
// of the true maximum) and a segment is wrap-prone when the total reaches 2^16.
__global__ void __launch_bounds__(256) k_seg_check(const uint16_t *__restrict__ ptn_pars, const uint16_t *__restrict__ w16T,
                                                   const int32_t *__restrict__ seg_upper, int seg0, int p_lo, int p_hi, int B, int Bpad,
                                                   int32_t *__restrict__ segmax)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int seg = seg0 + blockIdx.y;
    int lo = seg ? seg_upper[seg - 1] : 0;
    int hi = seg_upper[seg];
    if (lo < p_lo) lo = p_lo;
    if (hi > p_hi) hi = p_hi;
    unsigned long long sum = 0;
    if (b < B)
        for (int p = lo; p < hi; p++) sum += (unsigned long long)(__ldg(ptn_pars + p) + 1u) * __ldg(w16T + (size_t)p * Bpad + b);
    int v = sum > (1ull << 28) ? (1 << 28) : (int)sum;
    v = __reduce_max_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(&segmax[seg], v);
}

// cmax[seg] = max over the segment's patterns of the tree's per-pattern score: (cmax + 1) * max_b sum_{ptn in seg} w_b[ptn]
// bounds the sum k_seg_check computes, without touching the weights
__global__ void __launch_bounds__(128) k_seg_cmax(const uint16_t *__restrict__ ptn_pars, const int32_t *__restrict__ seg_upper, int upper,
                                                  int32_t *__restrict__ cmax)
{
    const int seg = blockIdx.x;
    const int lo = seg ? seg_upper[seg - 1] : 0;
    const int hi = min(seg_upper[seg], upper);
    int m = 0;
    for (int p = lo + threadIdx.x; p < hi; p += blockDim.x) m = max(m, (int)__ldg(ptn_pars + p));
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(&cmax[seg], m);
}

int launch_seg_cmax(Ctx *c, int32_t *d_cmax)
{
    Reps &r = c->reps;
    const int nseg = (int)r.seg_upper.size();
    if (r.upper == 0) return 0;
    k_seg_cmax<<<nseg, 128, 0, c->stream>>>(c->d_ptn, r.d_seg_upper, r.upper, d_cmax);
    c->launches++;
    MPGPU_CUDA(cudaGetLastError());
    return 0;
}

int launch_seg_check(Ctx *c, int32_t *d_segmax)
{
    Reps &r = c->reps;
    const int nseg = (int)r.seg_upper.size();
    if (r.upper == 0 || r.p_hi <= r.p_lo) return 0;
    for (int done = 0; done < nseg; done += 65535) {
        const int chunk = nseg - done < 65535 ? nseg - done : 65535;
        dim3 grid((r.B + 255) / 256, chunk);
        k_seg_check<<<grid, 256, 0, c->stream>>>(c->d_ptn, r.d_w16T, r.d_seg_upper, done, r.p_lo, r.p_hi, r.B, r.Bpad, d_segmax);
        c->launches++;
        MPGPU_CUDA(cudaGetLastError());
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// Rows.  Site space: one bit per expanded site of this shard (Wl words per row).
// ------------------------------------------------------------------------------------------
// mismatch row across an edge: ~OR_k(A_k & B_k)   (the bits evaluateParsimony counts, :1096-1125)
template <int S>
__global__ void __launch_bounds__(128) k_edge_rows(const uint32_t *__restrict__ views, size_t view_stride, int Wl,
                                                   const int4 *__restrict__ edges, uint32_t *__restrict__ rows)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int4 e = edges[blockIdx.y];                            // vidA, vidB, row
    constexpr int SG = S < 4 ? S : 4;
    const size_t gs = (size_t)Wl * SG, off = (size_t)w * SG;
    uint32_t n = 0;
#pragma unroll
    for (int g = 0; g < (S + SG - 1) / SG; g++) {
        if (SG == 4) {
            const uint4 a = __ldg(reinterpret_cast<const uint4 *>(views + (size_t)e.x * view_stride + off + g * gs));
            const uint4 b = __ldg(reinterpret_cast<const uint4 *>(views + (size_t)e.y * view_stride + off + g * gs));
            n |= (a.x & b.x) | (a.y & b.y) | (a.z & b.z) | (a.w & b.w);
        } else {
            const uint2 a = __ldg(reinterpret_cast<const uint2 *>(views + (size_t)e.x * view_stride + off));
            const uint2 b = __ldg(reinterpret_cast<const uint2 *>(views + (size_t)e.y * view_stride + off));
            n |= (a.x & b.x) | (a.y & b.y);
        }
    }
    rows[(size_t)e.z * Wl + w] = ~n;
}

int launch_edge_rows(Ctx *c, const int4 *d_edges, int nedges, uint32_t *d_rows)
{
    if (nedges == 0) return 0;
    dim3 grid(c->Wl / 128, nedges);
    switch (c->S) {
    case 2:  k_edge_rows<2><<<grid, 128, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, d_edges, d_rows); break;
    case 4:  k_edge_rows<4><<<grid, 128, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, d_edges, d_rows); break;
    case 20: k_edge_rows<20><<<grid, 128, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, d_edges, d_rows); break;
    case 32: k_edge_rows<32><<<grid, 128, 0, c->stream>>>(c->d_views, c->view_stride, c->Wl, d_edges, d_rows); break;
    default: set_error("unsupported state count"); return 1;
    }
    c->launches++;
    MPGPU_CUDA(cudaGetLastError());
    return 0;
}

// Site rows -> pattern rows: bit ptn of the output = bit at the FIRST expanded site of pattern
// ptn (what pllComputePatternParsimony reads, :3384).  One warp per 32 patterns x kGatherRows
// rows: the site of each pattern is looked up once and reused for every row of the block.
static const int kGatherRows = 32;
__global__ void __launch_bounds__(256) k_gather_rows(const uint32_t *__restrict__ src, int Wl, int64_t w0,
                                                     const int64_t *__restrict__ ptn_site, int upper, int Pw,
                                                     uint32_t *__restrict__ dst, int nrows)
{
    const int lane = threadIdx.x & 31;
    const int pw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pw >= Pw) return;
    const int ptn = 32 * pw + lane;
    int word = -1, sh = 0;
    if (ptn < upper) {
        const int64_t site = ptn_site[ptn];
        const int64_t w = (site >> 5) - w0;
        if (site >= 0 && w >= 0 && w < Wl) { word = (int)w; sh = (int)(site & 31); }
    }
    const int r0 = blockIdx.y * kGatherRows;
    const int r1 = r0 + kGatherRows < nrows ? r0 + kGatherRows : nrows;
    for (int row = r0; row < r1; row++) {
        uint32_t bit = 0;
        if (word >= 0) bit = (__ldg(src + (size_t)row * Wl + word) >> sh) & 1u;
        const uint32_t packed = __ballot_sync(0xffffffffu, bit);
        if (lane == 0) dst[(size_t)row * Pw + pw] = packed;
    }
}

int launch_gather_rows(Ctx *c, const uint32_t *d_src, uint32_t *d_dst, int nrows)
{
    if (nrows == 0) return 0;
    Reps &r = c->reps;
    for (int done = 0; done < nrows; done += 65535 * kGatherRows) {
        const int chunk = nrows - done < 65535 * kGatherRows ? nrows - done : 65535 * kGatherRows;
        dim3 grid((r.Pw + 7) / 8, (chunk + kGatherRows - 1) / kGatherRows);
        k_gather_rows<<<grid, 256, 0, c->stream>>>(d_src + (size_t)done * c->Wl, c->Wl, c->w0, c->d_ptn_site, r.upper, r.Pw,
                                                    d_dst + (size_t)done * r.Pw, chunk);
        c->launches++;
        MPGPU_CUDA(cudaGetLastError());
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// Exact CUDA-core contraction over the exception patterns.
//   X[row][group(e)][b] += bit[row][ptn(e)] * w16e[e][b]
// exceptions are sorted by group; one thread per (row, b).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_reps_exc(const uint32_t *__restrict__ rows_a, int a_pitch, int a_word0, int a_words, int row0, int x_row0,
                                                  const int32_t *__restrict__ exc_ptn, const int32_t *__restrict__ exc_group,
                                                  int n_exc, const uint16_t *__restrict__ w16T, int Bpad, int G,
                                                  int32_t *__restrict__ X)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = row0 + blockIdx.y;
    if (b >= Bpad) return;
    const uint32_t *bits = rows_a + (size_t)row * a_pitch;
    int32_t *xrow = X + (size_t)(x_row0 + row) * G * Bpad;
    int acc = 0, cur = -1;
    for (int e = 0; e < n_exc; e++) {
        const int g = __ldg(exc_group + e);
        if (g != cur) {
            if (cur >= 0 && acc) atomicAdd(&xrow[(size_t)cur * Bpad + b], acc);
            cur = g; acc = 0;
        }
        const int p = __ldg(exc_ptn + e);
        const int w = (p >> 5) - a_word0;                      // the row starts at pattern 32 * a_word0
        if (w >= 0 && w < a_words && ((__ldg(bits + w) >> (p & 31)) & 1u)) acc += (int)__ldg(w16T + (size_t)p * Bpad + b);
    }
    if (cur >= 0 && acc) atomicAdd(&xrow[(size_t)cur * Bpad + b], acc);
}

// rows a_base[0..nrows) (pitch a_pitch words; bit i of a row = pattern 32 * a_word0 + i) -> X[x_row0 ..][group][b]
int launch_reps_exc(Ctx *c, const uint32_t *a_base, int a_pitch, int a_word0, int x_row0, int nrows)
{
    Reps &r = c->reps;
    if (nrows == 0 || r.n_exc == 0) return 0;
    for (int done = 0; done < nrows; done += 65535) {
        const int chunk = nrows - done < 65535 ? nrows - done : 65535;
        dim3 grid((r.Bpad + 255) / 256, chunk);
        k_reps_exc<<<grid, 256, 0, c->stream>>>(a_base, a_pitch, a_word0, a_pitch, done, x_row0, r.d_exc_ptn, r.d_exc_group, r.n_exc,
                                                 r.d_w16T, r.Bpad, r.G, r.d_X);
        c->launches++;
        MPGPU_CUDA(cudaGetLastError());
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// Tensor-core contraction:  X[row][0][b] += sum_k bit[row][k] * w8[b][k]
//
// tcgen05.mma.cta_group::1.kind::i8, M = 128 rows, N = 256 replicates, K = 32 per instruction.
// CTA tile 128 x 256, K-block = 128 patterns (= one 128-byte swizzle row of u8).
//   warps 0-3  A producers: read 128 bits per row, expand to 0/1 bytes, write the K-major
//              SWIZZLE_128B tile by hand (generic proxy -> fence.proxy.async), then epilogue
//   warp 4     B producer: one TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) per K-block
//   warp 5     TMEM owner + MMA issuer (one elected lane)
// 4-stage mbarrier ring (full/empty), accumulator 128 lanes x 256 columns of TMEM (s32).
// Split-K over blockIdx.z; partial tiles are added with integer atomics (exact, order-free).
// ------------------------------------------------------------------------------------------
namespace tc {

constexpr int M = 128, N = 256, KB = 128, STAGES = 4;
constexpr int A_BYTES = M * KB, B_BYTES = N * KB, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
constexpr int threads_for(int npw) { return (npw + 2) * 32; }     // NPW producer warps + the TMA warp + the MMA warp

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tmap, int x, int y, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(bar) : "memory");
}
// K-major, SWIZZLE_128B operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);           // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: 8 rows x 128 B, bits [32,46)
    d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                            // layout type: SWIZZLE_128B
    return d;
}
// instruction descriptor: D = s32, A = B = u8, both K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc()
{
    return (2u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// 4 bits -> 4 bytes of 0/1 (bit j -> byte j)
__device__ __forceinline__ uint32_t spread4(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }

// BYTES: the A operand is already u8, K-major, [row][Kpad] (per-pattern Sankoff costs, -cost with -bb): the TMA warp
// loads its 128 x 128 tile through tmap_a next to the weights (one more cp.async.bulk.tensor per stage, rows beyond
// nrows zero-filled) and warps 0-3 only run the epilogue.
// NPW = A-producer warps (4: a thread expands the 128 bits of its row per K-block; 8: two threads share a row, 64 bits each --
// r02: four producers kept the tensor pipe at ~70 %).  Tiles are rasterised N-fastest (tile = m_tile * ntn + n_tile): the CTAs
// resident together share their A rows through L2 instead of re-reading them from DRAM once per replicate tile.
template <bool BYTES, int NPW>
__global__ void __launch_bounds__(threads_for(NPW), 1)
k_reps_tc(const __grid_constant__ CUtensorMap tmap_w8, const __grid_constant__ CUtensorMap tmap_a,
          const uint32_t *__restrict__ rows_a, int a_pitch, int a_kb0,
          int x_row0, int nrows, int kb_lo, int kb_hi, int kb_per_split, int B, int x_pitch, int32_t *__restrict__ X,
          int mtn, int ntn, int n_fastest)
{
    constexpr int TMA_WARP = NPW, MMA_WARP = NPW + 1;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t bars = base + STAGES * STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    const uint32_t accum_bar = bars + 8u * (2 * STAGES);
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 1);
    uint8_t *smem_gen = smem_raw + (base - smem_u32(smem_raw));            // generic pointer to the aligned base

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    const int m0 = (n_fastest ? tile / ntn : tile % mtn) * M, n0 = (n_fastest ? tile % ntn : tile / mtn) * N;
    const int kb_begin = kb_lo + blockIdx.z * kb_per_split;
    int kb_end = kb_begin + kb_per_split;
    if (kb_end > kb_hi) kb_end = kb_hi;
    const int nkb = kb_end - kb_begin;
    if (nkb <= 0) return;                                                  // uniform over the CTA

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full_bar(s), BYTES ? 1 : NPW + 1); mbar_init(empty_bar(s), 1); }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "n"(N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp < NPW) {
        // ---- A producers: thread t owns tile row t % 128; with 8 producer warps the two threads of a row expand 64 bits each ----
        constexpr int SHARE = NPW / 4;                       // threads per row
        constexpr int CH = 8 / SHARE;                        // 16-byte chunks (16 bits each) per thread and K-block
        const int r = threadIdx.x & 127, half = threadIdx.x >> 7;
        const bool live = (m0 + r) < nrows;
        if (!BYTES) {
        typedef typename std::conditional<SHARE == 1, uint4, uint2>::type BitsT;
        const BitsT *src = reinterpret_cast<const BitsT *>(reinterpret_cast<const uint4 *>(rows_a + (size_t)(m0 + r) * a_pitch) - a_kb0) + half;   // row starts at K-block a_kb0
        const uint32_t sw = (uint32_t)(r & 7);
        // the bits of K-block it+PF are requested while K-block it is expanded: the L2 latency of
        // the row loads is off the critical path of the MMA pipeline
        constexpr int PF = 4;
        BitsT pre[PF];
#pragma unroll
        for (int d = 0; d < PF; d++) {
            pre[d] = BitsT();
            if (live && d < nkb) pre[d] = __ldg(src + (size_t)(kb_begin + d) * SHARE);
        }
        for (int it0 = 0; it0 < nkb; it0 += PF) {
#pragma unroll
            for (int d = 0; d < PF; d++) {
                const int it = it0 + d;
                if (it >= nkb) break;
                const int s = it % STAGES;
                const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                const BitsT bits = pre[d];
                if (live && it + PF < nkb) pre[d] = __ldg(src + (size_t)(kb_begin + it + PF) * SHARE);
                mbar_wait(empty_bar(s), ph ^ 1u);
                uint8_t *arow = smem_gen + (size_t)s * STAGE_BYTES + (size_t)r * 128;
                const uint32_t *wv = reinterpret_cast<const uint32_t *>(&bits);
#pragma unroll
                for (int cc = 0; cc < CH; cc++) {
                    const int cidx = half * CH + cc;
                    const uint32_t h = (wv[cc >> 1] >> ((cc & 1) * 16)) & 0xFFFFu;
                    uint4 v;
                    v.x = spread4(h & 0xF); v.y = spread4((h >> 4) & 0xF); v.z = spread4((h >> 8) & 0xF); v.w = spread4(h >> 12);
                    *reinterpret_cast<uint4 *>(arow + ((cidx ^ sw) << 4)) = v;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(full_bar(s));
            }
        }
        }
    } else if (warp == TMA_WARP) {
        // ---- B producer (TMA) ----
        if (lane == 0) {
            for (int it = 0; it < nkb; it++) {
                const int s = it % STAGES;
                const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1u);
                mbar_arrive_expect_tx(full_bar(s), BYTES ? STAGE_BYTES : B_BYTES);
                if (BYTES) tma_load_2d(base + s * STAGE_BYTES, &tmap_a, (kb_begin + it) * KB, m0, full_bar(s));
                tma_load_2d(base + s * STAGE_BYTES + A_BYTES, &tmap_w8, (kb_begin + it) * KB, n0, full_bar(s));
            }
        }
    } else {
        // ---- MMA issuer ----
        if (lane == 0) {
            const uint32_t idesc = umma_idesc();
            for (int it = 0; it < nkb; it++) {
                const int s = it % STAGES;
                const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                mbar_wait(full_bar(s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t ad = umma_desc(base + s * STAGE_BYTES);
                const uint64_t bd = umma_desc(base + s * STAGE_BYTES + A_BYTES);
#pragma unroll
                for (int k = 0; k < KB / 32; k++)                           // +32 bytes along K inside the swizzle row
                    umma_i8(tmem_base, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (it | k) != 0);
                umma_commit(empty_bar(s));                                   // frees the stage when these MMAs retire
            }
            umma_commit(accum_bar);
        }
        __syncwarp();
    }

    if (warp < 4) {
        // ---- epilogue: TMEM -> registers -> (transpose through smem) -> coalesced integer atomics ----
        mbar_wait(accum_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t *tile = reinterpret_cast<uint32_t *>(smem_gen) + warp * (32 * 33);   // stage memory is free now
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int cb = 0; cb < N / 32; cb++) {
            uint32_t v[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr + (uint32_t)(cb * 32)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; j++) tile[lane * 33 + j] = v[j];
            __syncwarp();
            const int col = n0 + cb * 32 + lane;
            for (int i = 0; i < 32; i++) {
                const int mrow = m0 + warp * 32 + i;
                const int val = (int)tile[i * 33 + lane];
                if (mrow < nrows && col < B && val != 0)
                    atomicAdd(&X[(size_t)(x_row0 + mrow) * x_pitch + col], val);
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == MMA_WARP)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(N) : "memory");
}

// ---- A operand in tensor memory (r02) -------------------------------------------------------------------------------
// With both operands in shared memory the MMA reads 12 KB per instruction (A 4 KB + B 8 KB every 128 cycles = 96 B/clk)
// while the producers write the A tile (32 B/clk) and TMA writes the B tile (64 B/clk): 192 B/clk against the SM's
// 128 B/clk of shared-memory bandwidth -- the tensor pipe cannot exceed ~2/3 (measured 66-70 %).  The {0,1} A tile does
// not have to exist in shared memory at all: tcgen05.mma takes A from TENSOR memory (lane = tile row, 32-bit column =
// 4 consecutive K bytes), and the producers can put it there straight from registers with tcgen05.st.  Producer thread r
// (row r, TMEM lane r) expands its 128 bits of a K-block into 32 words and stores them to the stage's 32 columns; the
// MMA issuer passes [tmem_a + 8 k] for the k-th K = 32 slice.  Shared memory then carries the weights only
// (64 B/clk read + 64 B/clk written), the stages are 32 KB and six of them fit.
// TMEM: columns [0, 256) accumulator, [256, 256 + 32 * TA_STAGES) A stages (512 allocated; one CTA per SM).
constexpr int TA_STAGES = 6;
constexpr int TA_SMEM_BYTES = TA_STAGES * B_BYTES + 1024 + 256;
constexpr int TA_ACOL = 256;

__device__ __forceinline__ void umma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t"
        "}" :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__global__ void __launch_bounds__(192, 1)
k_reps_tc_ta(const __grid_constant__ CUtensorMap tmap_w8,
             const uint32_t *__restrict__ rows_a, int a_pitch, int a_kb0,
             int x_row0, int nrows, int kb_lo, int kb_hi, int kb_per_split, int B, int x_pitch, int32_t *__restrict__ X,
             int mtn, int ntn, int n_fastest)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + TA_STAGES * B_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (TA_STAGES + s); };
    const uint32_t accum_bar = bars + 8u * (2 * TA_STAGES);
    const uint32_t tmem_slot = bars + 8u * (2 * TA_STAGES + 1);
    uint8_t *smem_gen = smem_raw + (base - smem_u32(smem_raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    const int m0 = (n_fastest ? tile / ntn : tile % mtn) * M, n0 = (n_fastest ? tile % ntn : tile / mtn) * N;
    const int kb_begin = kb_lo + blockIdx.z * kb_per_split;
    int kb_end = kb_begin + kb_per_split;
    if (kb_end > kb_hi) kb_end = kb_hi;
    const int nkb = kb_end - kb_begin;
    if (nkb <= 0) return;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TA_STAGES; s++) { mbar_init(full_bar(s), 5); mbar_init(empty_bar(s), 1); }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp < 4) {
        // ---- A producers: thread r owns tile row r = TMEM lane r ----
        const int r = threadIdx.x;
        const bool live = (m0 + r) < nrows;
        const uint4 *src = reinterpret_cast<const uint4 *>(rows_a + (size_t)(m0 + r) * a_pitch) - a_kb0;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)TA_ACOL;
        constexpr int PF = 4;
        uint4 pre[PF];
#pragma unroll
        for (int d = 0; d < PF; d++) {
            pre[d] = make_uint4(0, 0, 0, 0);
            if (live && d < nkb) pre[d] = __ldg(src + (kb_begin + d));
        }
        for (int it0 = 0; it0 < nkb; it0 += PF) {
#pragma unroll
            for (int d = 0; d < PF; d++) {
                const int it = it0 + d;
                if (it >= nkb) break;
                const int s = it % TA_STAGES;
                const uint32_t ph = (uint32_t)(it / TA_STAGES) & 1u;
                const uint4 bits = pre[d];
                if (live && it + PF < nkb) pre[d] = __ldg(src + (kb_begin + it + PF));
                const uint32_t wv[4] = {bits.x, bits.y, bits.z, bits.w};
                uint32_t v[32];
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = spread4((wv[j >> 3] >> ((j & 7) * 4)) & 0xFu);
                mbar_wait(empty_bar(s), ph ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                    "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                    "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                    :: "r"(lane_addr + (uint32_t)(s * 32)),
                       "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                       "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
                       "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
                       "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
                    : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(full_bar(s));
            }
        }
    } else if (warp == 4) {
        if (lane == 0) {
            for (int it = 0; it < nkb; it++) {
                const int s = it % TA_STAGES;
                const uint32_t ph = (uint32_t)(it / TA_STAGES) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1u);
                mbar_arrive_expect_tx(full_bar(s), B_BYTES);
                tma_load_2d(base + s * B_BYTES, &tmap_w8, (kb_begin + it) * KB, n0, full_bar(s));
            }
        }
    } else {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc();
            for (int it = 0; it < nkb; it++) {
                const int s = it % TA_STAGES;
                const uint32_t ph = (uint32_t)(it / TA_STAGES) & 1u;
                mbar_wait(full_bar(s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t bd = umma_desc(base + s * B_BYTES);
                const uint32_t at = tmem_base + (uint32_t)(TA_ACOL + s * 32);
#pragma unroll
                for (int k = 0; k < KB / 32; k++)
                    umma_i8_ts(tmem_base, at + (uint32_t)(8 * k), bd + (uint64_t)(2 * k), idesc, (it | k) != 0);
                umma_commit(empty_bar(s));
            }
            umma_commit(accum_bar);
        }
        __syncwarp();
    }

    if (warp < 4) {
        mbar_wait(accum_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t *tile_s = reinterpret_cast<uint32_t *>(smem_gen) + warp * (32 * 33);
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int cb = 0; cb < N / 32; cb++) {
            uint32_t v[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr + (uint32_t)(cb * 32)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; j++) tile_s[lane * 33 + j] = v[j];
            __syncwarp();
            const int col = n0 + cb * 32 + lane;
            for (int i = 0; i < 32; i++) {
                const int mrow = m0 + warp * 32 + i;
                const int val = (int)tile_s[i * 33 + lane];
                if (mrow < nrows && col < B && val != 0)
                    atomicAdd(&X[(size_t)(x_row0 + mrow) * x_pitch + col], val);
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(512) : "memory");
}

// ---- the denominator of k_reps_tc's roofline, measured: tcgen05.mma kind::i8 issued back to back ----------------------
// One CTA per SM, one warp: a 128 x 128 A tile and a 256 x 128 B tile sit in shared memory (SWIZZLE_128B, the layout the
// contraction uses), and the elected lane issues M = 128, N = 256, K = 32 MMAs into one TMEM accumulator with nothing else
// in the way -- no loads, no producers, no epilogue; a commit every 64 MMAs keeps the queue bounded.  ops = 2 M N K per MMA.
__global__ void __launch_bounds__(32, 1) k_i8_peak(int iters)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar = base + STAGE_BYTES, tmem_slot = bar + 8;
    const int lane = threadIdx.x;
    for (int i = lane; i < STAGE_BYTES / 4; i += 32) reinterpret_cast<uint32_t *>(gen)[i] = 0x01000101u * (uint32_t)((i * 2654435761u) >> 31);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "n"(N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (lane == 0) {
        const uint32_t idesc = umma_idesc();
        const uint64_t ad = umma_desc(base), bd = umma_desc(base + A_BYTES);
        uint32_t phase = 0;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int j = 0; j < 16; j++)
#pragma unroll
                for (int k = 0; k < KB / 32; k++) umma_i8(tmem_base, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (it | j | k) != 0);
            umma_commit(bar);
            mbar_wait(bar, phase);
            phase ^= 1u;
        }
    }
    __syncwarp();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(N) : "memory");
}

}  // namespace tc

// measured issue rate of tcgen05.mma kind::i8 on this device, in int8 TOP/s (2 ops per multiply-accumulate)
int measure_int8_peak(Ctx *c, int iters, double *tops)
{
    cudaDeviceProp prop;
    MPGPU_CUDA(cudaGetDeviceProperties(&prop, c->device));
    const int smem = tc::STAGE_BYTES + 1024 + 64;
    MPGPU_CUDA(cudaFuncSetAttribute(tc::k_i8_peak, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    MPGPU_CUDA(cudaEventCreate(&e0)); MPGPU_CUDA(cudaEventCreate(&e1));
    tc::k_i8_peak<<<prop.multiProcessorCount, 32, smem, c->stream>>>(iters / 8 + 1);          // warm-up
    MPGPU_CUDA(cudaEventRecord(e0, c->stream));
    tc::k_i8_peak<<<prop.multiProcessorCount, 32, smem, c->stream>>>(iters);
    MPGPU_CUDA(cudaEventRecord(e1, c->stream));
    MPGPU_CUDA(cudaGetLastError());
    MPGPU_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    MPGPU_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    c->launches += 2;
    const double ops = 2.0 * tc::M * tc::N * 32.0 * 64.0 * (double)iters * prop.multiProcessorCount;
    *tops = ops / (ms * 1e-3) / 1e12;
    return 0;
}

// cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode(EncodeTiledFn *out)
{
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        MPGPU_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available in this driver"); return 1; }
        encode = (EncodeTiledFn)fn;
    }
    *out = encode;
    return 0;
}

int make_w8_tensor_map(Ctx *c)
{
    Reps &r = c->reps;
    EncodeTiledFn encode = nullptr;
    if (int rc = get_encode(&encode)) return rc;
    static_assert(sizeof(CUtensorMap) <= sizeof(r.tmap_w8), "tensor map storage too small");
    const cuuint64_t dims[2] = {(cuuint64_t)r.Kpad, (cuuint64_t)r.Bpad};
    const cuuint64_t strides[1] = {(cuuint64_t)r.Kpad};                   // bytes between replicate rows
    const cuuint32_t box[2] = {(cuuint32_t)tc::KB, (cuuint32_t)tc::N};
    const cuuint32_t estr[2] = {1, 1};
    CUresult res = encode(reinterpret_cast<CUtensorMap *>(r.tmap_w8), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, r.d_w8, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (res != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed for the replicate-weight matrix"); return 1; }
    r.tmap_valid = true;
    return 0;
}

// X[x_row0 .. x_row0+nrows)[group 0] += rows x w8   (this shard's K-blocks [kb_lo, kb_hi));
// rows = a_base[0..nrows), pitch a_pitch words (a multiple of 4); bit i of a row = pattern 32 * a_word0 + i
// (a_word0 a multiple of 4: rows start on a K-block boundary)
int launch_reps_tc(Ctx *c, const uint32_t *a_base, int a_pitch, int a_word0, int x_row0, int nrows)
{
    Reps &r = c->reps;
    if (nrows == 0) return 0;
    if (!r.tmap_valid) { set_error("replicate weights not loaded"); return 1; }
    const int kb_lo = r.kb_lo, kb_hi = r.kb_hi;
    if (kb_hi <= kb_lo) return 0;
    static bool configured_dev[64] = {false}; bool &configured = configured_dev[c->device & 63];   /* the attribute is per device */
    if (!configured) {
        MPGPU_CUDA(cudaFuncSetAttribute(tc::k_reps_tc<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
        MPGPU_CUDA(cudaFuncSetAttribute(tc::k_reps_tc<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
        configured = true;
    }
    static const int producers = getenv("MPGPU_REPS_PRODUCERS") ? atoi(getenv("MPGPU_REPS_PRODUCERS")) : 4;   // (8 measured slower: 1.10 vs 1.02 ms)     // tuning knobs
    static const int n_fastest = getenv("MPGPU_REPS_RASTER") ? atoi(getenv("MPGPU_REPS_RASTER")) : 1;
    const int mt = (nrows + tc::M - 1) / tc::M, nt = r.Bpad / tc::N;
    const int nkb = kb_hi - kb_lo;
    // Split K so that the grid comes out in full waves of 148 CTAs (one per SM): cost model
    // waves x (K-blocks per CTA + epilogue), the epilogue (TMEM -> atomics) ~ 24 K-blocks of MMA time.
    int splits = 1;
    {
        const int ctas = mt * nt, sms = 148, epi = 24;
        long best = -1;
        for (int sp = 1; sp <= 32 && sp * 8 <= nkb; sp++) {
            const long waves = ((long)ctas * sp + sms - 1) / sms;
            const long cost = waves * ((nkb + sp - 1) / sp + epi);
            if (best < 0 || cost < best) { best = cost; splits = sp; }
        }
    }
    if (const char *e = getenv("MPGPU_REPS_SPLITS")) { int v = atoi(e); if (v >= 1) splits = v; }
    const int per = (nkb + splits - 1) / splits;
    splits = (nkb + per - 1) / per;
    dim3 grid(mt * nt, 1, splits);
    const bool timed = r.timing && nrows >= r.timed_rows;      // keep the events of the largest launch
    if (timed) {
        if (!r.ev0) { MPGPU_CUDA(cudaEventCreate(&r.ev0)); MPGPU_CUDA(cudaEventCreate(&r.ev1)); }
        MPGPU_CUDA(cudaEventRecord(r.ev0, c->stream));
    }
    static const int tmem_a = getenv("MPGPU_REPS_TMEM_A") ? atoi(getenv("MPGPU_REPS_TMEM_A")) : 1;
    if (tmem_a) {
        static bool ta_configured_dev[64] = {false}; bool &tac = ta_configured_dev[c->device & 63];
        if (!tac) { MPGPU_CUDA(cudaFuncSetAttribute(tc::k_reps_tc_ta, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::TA_SMEM_BYTES)); tac = true; }
        tc::k_reps_tc_ta<<<grid, 192, tc::TA_SMEM_BYTES, c->stream>>>(*reinterpret_cast<const CUtensorMap *>(r.tmap_w8), a_base, a_pitch, a_word0 / 4,
                                                                     x_row0, nrows, kb_lo, kb_hi, per, r.B, r.G * r.Bpad, r.d_X, mt, nt, n_fastest);
    } else if (producers == 4)
        tc::k_reps_tc<false, 4><<<grid, tc::threads_for(4), tc::SMEM_BYTES, c->stream>>>(*reinterpret_cast<const CUtensorMap *>(r.tmap_w8),
                                                                           *reinterpret_cast<const CUtensorMap *>(r.tmap_w8), a_base, a_pitch, a_word0 / 4,
                                                                           x_row0, nrows, kb_lo, kb_hi, per, r.B, r.G * r.Bpad, r.d_X, mt, nt, n_fastest);
    else
        tc::k_reps_tc<false, 8><<<grid, tc::threads_for(8), tc::SMEM_BYTES, c->stream>>>(*reinterpret_cast<const CUtensorMap *>(r.tmap_w8),
                                                                           *reinterpret_cast<const CUtensorMap *>(r.tmap_w8), a_base, a_pitch, a_word0 / 4,
                                                                           x_row0, nrows, kb_lo, kb_hi, per, r.B, r.G * r.Bpad, r.d_X, mt, nt, n_fastest);
    if (timed) { MPGPU_CUDA(cudaEventRecord(r.ev1, c->stream)); r.timed_rows = nrows; r.timed_kblocks = nkb; r.timed_splits = splits; }
    c->launches++;
    MPGPU_CUDA(cudaGetLastError());
    return 0;
}

// X[0..nrows)[b] += rows8 x w8 over all K-blocks: rows8 = u8 [nrows][pitch_bytes] (pitch = Kpad), X pitch x_pitch (zeroed by
// the caller).  The -cost / -bb contraction (per-pattern Sankoff costs that fit in u8, wrap-free segments).
int launch_reps_tc_bytes(Ctx *c, const uint8_t *rows8, int pitch_bytes, int nrows, int32_t *X, int x_pitch)
{
    Reps &r = c->reps;
    if (nrows == 0) return 0;
    if (!r.tmap_valid) { set_error("replicate weights not loaded"); return 1; }
    static bool configured_dev[64] = {false}; bool &configured = configured_dev[c->device & 63];   /* the attribute is per device */
    if (!configured) {
        MPGPU_CUDA(cudaFuncSetAttribute(tc::k_reps_tc<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
        configured = true;
    }
    const int mt = (nrows + tc::M - 1) / tc::M, nt = r.Bpad / tc::N;
    const int nkb = r.Kpad / tc::KB;
    int splits = 1;
    {
        const int ctas = mt * nt, sms = 148, epi = 24;
        long best = -1;
        for (int sp = 1; sp <= 32 && sp * 8 <= nkb; sp++) {
            const long waves = ((long)ctas * sp + sms - 1) / sms;
            const long cost = waves * ((nkb + sp - 1) / sp + epi);
            if (best < 0 || cost < best) { best = cost; splits = sp; }
        }
    }
    const int per = (nkb + splits - 1) / splits;
    splits = (nkb + per - 1) / per;
    dim3 grid(mt * nt, 1, splits);
    static const int n_fastest = getenv("MPGPU_REPS_RASTER") ? atoi(getenv("MPGPU_REPS_RASTER")) : 1;
    const bool timed = r.timing && nrows >= r.timed_rows;
    if (timed) {
        if (!r.ev0) { MPGPU_CUDA(cudaEventCreate(&r.ev0)); MPGPU_CUDA(cudaEventCreate(&r.ev1)); }
        MPGPU_CUDA(cudaEventRecord(r.ev0, c->stream));
    }
    alignas(64) unsigned char tmap_a[128];
    {
        EncodeTiledFn encode = nullptr;
        if (int rc = get_encode(&encode)) return rc;
        const cuuint64_t dims[2] = {(cuuint64_t)pitch_bytes, (cuuint64_t)nrows};
        const cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
        const cuuint32_t box[2] = {(cuuint32_t)tc::KB, (cuuint32_t)tc::M};
        const cuuint32_t estr[2] = {1, 1};
        CUresult res = encode(reinterpret_cast<CUtensorMap *>(tmap_a), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)rows8, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (res != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed for the cost rows"); return 1; }
    }
    tc::k_reps_tc<true, 4><<<grid, tc::threads_for(4), tc::SMEM_BYTES, c->stream>>>(*reinterpret_cast<const CUtensorMap *>(r.tmap_w8),
                                                                          *reinterpret_cast<const CUtensorMap *>(tmap_a),
                                                                          reinterpret_cast<const uint32_t *>(rows8), pitch_bytes, 0,
                                                                          0, nrows, 0, nkb, per, r.B, x_pitch, X, mt, nt, n_fastest);
    if (timed) { MPGPU_CUDA(cudaEventRecord(r.ev1, c->stream)); r.timed_rows = nrows; r.timed_kblocks = nkb; r.timed_splits = splits; }
    c->launches++;
    MPGPU_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// Combine.
//   tree row:  X[t][g][b] = sum_bit X[plane_bit][g][b] << bit
//   call j:    res[j][b]  = sum_g mask_g( X[t][g][b] - X[e_j][g][b] + X[d_j][g][b] ),  mask_0 = id, else & 0xFFFF
// A call with a hit (res <= thr[b] for some b, i.e. rell >= boot_logl[b] at the start of the batch)
// is flagged so that the host only reads back the rows that can change a replicate.
// ------------------------------------------------------------------------------------------
__global__ void k_reps_tree_row(int32_t *__restrict__ X, int plane_row0, int nbits, int t_row, int pitch)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pitch) return;
    int acc = 0;
    for (int b = 0; b < nbits; b++) acc += X[(size_t)(plane_row0 + b) * pitch + i] << b;
    X[(size_t)t_row * pitch + i] = acc;
}

__global__ void __launch_bounds__(256) k_reps_combine(const int32_t *__restrict__ X, int G, int Bpad, int B, int t_row,
                                                      const int2 *__restrict__ calls, int call0, int ncalls, int32_t *__restrict__ res,
                                                      const int32_t *__restrict__ thr, int32_t *__restrict__ call_hit, int nowrap_cols)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = call0 + blockIdx.y;                           // uniform over the block
    if (j >= ncalls) return;
    int hit = 0;
    if (b < Bpad) {
        const int2 cd = calls[j];                               // x = edge row (-1 none), y = delta row (-1 none)
        const size_t pitch = (size_t)G * Bpad;
        int total = 0;
        for (int g = 0; g < G; g++) {
            int v = X[(size_t)t_row * pitch + (size_t)g * Bpad + b];
            if (cd.x >= 0) v -= X[(size_t)cd.x * pitch + (size_t)g * Bpad + b];
            if (cd.y >= 0) v += X[(size_t)cd.y * pitch + (size_t)g * Bpad + b];
            total += (g == 0 || b < nowrap_cols) ? v : (v & 0xFFFF);     // columns below nowrap_cols: the plain-int loop of -autovec (iqtree.cpp:3418-3423)
        }
        res[(size_t)j * Bpad + b] = total;
        hit = thr && b < B && total <= thr[b];
    }
    if (thr && __syncthreads_or(hit) && threadIdx.x == 0) call_hit[j] = 1;
}

// rows of res listed in `list` -> contiguous rows of `out`
__global__ void k_gather_res_rows(const int32_t *__restrict__ res, int Bpad, const int32_t *__restrict__ list, int32_t *__restrict__ out)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < Bpad) out[(size_t)blockIdx.y * Bpad + b] = res[(size_t)list[blockIdx.y] * Bpad + b];
}

int launch_reps_tree_row(Ctx *c, int plane_row0, int nbits, int t_row)
{
    Reps &r = c->reps;
    const int pitch = r.G * r.Bpad;
    k_reps_tree_row<<<(pitch + 255) / 256, 256, 0, c->stream>>>(r.d_X, plane_row0, nbits, t_row, pitch);
    c->launches++;
    MPGPU_CUDA(cudaGetLastError());
    return 0;
}

int launch_reps_combine(Ctx *c, int t_row, const int2 *d_calls, int ncalls, int32_t *d_res, const int32_t *d_thr, int32_t *d_call_hit)
{
    Reps &r = c->reps;
    if (ncalls == 0) return 0;
    for (int done = 0; done < ncalls; done += 65535) {
        const int chunk = ncalls - done < 65535 ? ncalls - done : 65535;
        dim3 grid((r.Bpad + 255) / 256, chunk);
        k_reps_combine<<<grid, 256, 0, c->stream>>>(r.d_X, r.G, r.Bpad, r.Buser, t_row, d_calls, done, ncalls, d_res, d_thr, d_call_hit,
                                                    r.nowrap ? r.Buser : 0);
        c->launches++;
        MPGPU_CUDA(cudaGetLastError());
    }
    return 0;
}

int launch_gather_res_rows(Ctx *c, const int32_t *d_res, const int32_t *d_list, int nlist, int32_t *d_out)
{
    Reps &r = c->reps;
    for (int done = 0; done < nlist; done += 65535) {
        const int chunk = nlist - done < 65535 ? nlist - done : 65535;
        dim3 grid((r.Bpad + 255) / 256, chunk);
        k_gather_res_rows<<<grid, 256, 0, c->stream>>>(d_res, r.Bpad, d_list + done, d_out + (size_t)done * r.Bpad);
        c->launches++;
        MPGPU_CUDA(cudaGetLastError());
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// -distinct_iter_top_boot (iqtree.cpp:3587-3685).  Under that policy a replicate is accepted against boot_threshold
// (the worst of its top list) while the remain-bound skip of the REPS loop (:3433-3445) compares against boot_logl
// (its best), so the skip changes decisions and has to be decided exactly:
//     skipped(call, b)  <=>  exists s, nseg/4 < s < nseg-1 :  -(prefix_s + remain[b][s]) < boot_logl[b] - eps,
//     prefix_s = sum_{j <= s} ( sum_{ptn in segment j} c[ptn] * w_b[ptn]  mod 2^16 )
//  <=>  -(max_s (prefix_s + remain[b][s])) < boot_logl[b] - eps.
// The maximum does not depend on the bookkeeping state, so the device computes it for the (call, replicate) pairs the
// replay asks for: c = c_T - mis + delta from the tree's per-pattern scores and the call's two kept bit rows.
// ------------------------------------------------------------------------------------------
// rows (edge row, delta row) of the calls list[0..nl) -> dst[i][2][Pw]; a call without rows (the current tree) keeps zeros
__global__ void __launch_bounds__(256) k_keep_rows(const uint32_t *__restrict__ src, int pitch, const int2 *__restrict__ calls,
                                                   const int32_t *__restrict__ list, int Pw, uint32_t *__restrict__ dst)
{
    const int i = blockIdx.y;
    const int2 cd = calls[list[i]];
    uint32_t *d = dst + (size_t)i * 2 * Pw;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < Pw; w += gridDim.x * blockDim.x) {
        d[w] = cd.x >= 0 ? __ldg(src + (size_t)cd.x * pitch + w) : 0u;
        d[Pw + w] = cd.y >= 0 ? __ldg(src + (size_t)cd.y * pitch + w) : 0u;
    }
}

int launch_keep_rows(Ctx *c, const uint32_t *d_src, int pitch, const int2 *d_calls, const int32_t *d_list, int nl, uint32_t *d_dst)
{
    Reps &r = c->reps;
    for (int done = 0; done < nl; done += 65535) {
        const int chunk = nl - done < 65535 ? nl - done : 65535;
        dim3 grid((r.Pw + 255) / 256 < 8 ? (r.Pw + 255) / 256 : 8, chunk);
        k_keep_rows<<<grid, 256, 0, c->stream>>>(d_src, pitch, d_calls, d_list + done, r.Pw, d_dst + (size_t)done * 2 * r.Pw);
        c->launches++;
        MPGPU_CUDA(cudaGetLastError());
    }
    return 0;
}

// segsum[e][s] = ( sum_{ptn in segment s} (c_T[ptn] - mis[ptn] + delta[ptn]) * w[blist[e]][ptn] ) mod 2^16.
// lane = list entry (consecutive replicates of a dense list read w16T coalesced), the warps of the grid split the segments;
// the candidate's score of a pattern is uniform over the warp.
__global__ void __launch_bounds__(256) k_seg_sums(const uint16_t *__restrict__ ptn_pars, const uint32_t *__restrict__ rows2, int Pw,
                                                  const uint16_t *__restrict__ w16T, int Bpad, const int32_t *__restrict__ seg_upper,
                                                  int nseg, int upper, const int32_t *__restrict__ blist, int nl, int32_t *__restrict__ segsum)
{
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * 32 + lane;
    const int b = e < nl ? blist[e] : -1;
    const int nwarps = gridDim.y * (blockDim.x >> 5);
    for (int s = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5); s < nseg; s += nwarps) {
        const int lo = s ? __ldg(seg_upper + s - 1) : 0;
        int hi = __ldg(seg_upper + s);
        if (hi > upper) hi = upper;
        uint32_t acc = 0;
        for (int p = lo; p < hi; p++) {
            const uint32_t mis = (__ldg(rows2 + (p >> 5)) >> (p & 31)) & 1u, dlt = (__ldg(rows2 + Pw + (p >> 5)) >> (p & 31)) & 1u;
            const uint32_t cand = (uint32_t)__ldg(ptn_pars + p) - mis + dlt;
            if (b >= 0) acc += cand * (uint32_t)__ldg(w16T + (size_t)p * Bpad + b);      // mod 2^32 keeps the low 16 bits exact
        }
        if (b >= 0) segsum[(size_t)e * nseg + s] = (int32_t)(acc & 0xFFFFu);
    }
}

// out[e] = max over nseg/4 < s < nseg-1 of (sum_{j<=s} segsum[e][j] + remain[blist[e]][s]); INT_MIN when no segment is in range.
// One warp per entry: 32 segments per step, warp scan with a carry.
__global__ void __launch_bounds__(256) k_prefix_max(const int32_t *__restrict__ segsum, int nseg, const int32_t *__restrict__ remain,
                                                    const int32_t *__restrict__ blist, int nl, int32_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= nl) return;
    const int b = blist[e];
    int carry = 0, best = (int)0x80000000;
    for (int s0 = 0; s0 < nseg - 1; s0 += 32) {
        const int s = s0 + lane;
        int v = s < nseg ? segsum[(size_t)e * nseg + s] : 0;
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += t; }
        const int prefix = carry + v;
        if (s > nseg / 4 && s < nseg - 1) {
            const int tot = prefix + __ldg(remain + (size_t)b * (nseg - 1) + s);
            if (tot > best) best = tot;
        }
        carry = __shfl_sync(0xffffffffu, prefix, 31);
    }
    best = __reduce_max_sync(0xffffffffu, best);
    if (lane == 0) out[e] = best;
}

int launch_prefix_max(Ctx *c, const uint16_t *d_tree_ptn, const uint32_t *d_rows2, const int32_t *d_remain, const int32_t *d_blist, int nl,
                      int32_t *d_segsum, int32_t *d_out)
{
    Reps &r = c->reps;
    const int nseg = (int)r.seg_upper.size();
    if (nl == 0) return 0;
    const int gx = (nl + 31) / 32;
    int gy = (nseg + 7) / 8;                                   // 8 warps per block, one segment per warp and round
    const int want = (148 * 8 + gx - 1) / gx;                  // enough blocks for every SM when the list is short
    if (gy > want) gy = want;
    if (gy < 1) gy = 1;
    k_seg_sums<<<dim3(gx, gy), 256, 0, c->stream>>>(d_tree_ptn, d_rows2, r.Pw, r.d_w16T, r.Bpad, r.d_seg_upper, nseg, r.upper, d_blist, nl, d_segsum);
    c->launches++;
    MPGPU_CUDA(cudaGetLastError());
    k_prefix_max<<<(nl + 7) / 8, 256, 0, c->stream>>>(d_segsum, nseg, d_remain, d_blist, nl, d_out);
    c->launches++;
    MPGPU_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mpgpu
