// sgr_backward.cu -- backward path of the B200 rasterizer.
//
// Replaces (reference: gaussian_splatting/submodules/diff-gaussian-rasterization/cuda_rasterizer/)
//   renderCUDA (backward)     backward.cu:399-557  -> blend_backward_kernel
//   computeCov2DCUDA          backward.cu:144-274  \
//   preprocessCUDA (backward) backward.cu:346-396   > preprocess_backward_kernel (one fused pass)
//   computeColorFromSH / computeCov3D backward :20-139, :278-341 /
//   9x torch::zeros           rasterize_points.cu:151-159 -> outputs fully written by the kernel;
//                              only the 48 B/Gaussian accumulator record is memset.
//
// The reference issues 9 global atomicAdd per contributing (pixel, Gaussian) pair, up to 256-way
// contended.  Here a warp parks the two per-pair scalars of its 32 pixels in shared memory, reduces 16
// splats at a time in a transposed pass and issues two 16-byte vector reductions
// (red.global.add.v4.f32 -> SASS REDG.E.ADD.F32x4) plus one scalar per (8x4 block, Gaussian): 3 L2
// operations instead of 9 per (pixel, Gaussian).
#include <mutex>

#include "sgr_internal.cuh"

namespace sgr {

// Record pipeline of the backward blend.  SGR_BWD_PIPE3 = 1 (default): SGR_BWD_NS stages of SGR_BWD_B records with
// full / empty mbarriers -- a warp waits only for the loader warps' data, never for its sibling warps, so the
// tile's eight warps may drift NS - 2 batches apart and the statistical imbalance between the blocks of one batch
// (the busiest block of a 128-record batch does 1.32x the mean) is averaged over more records: 1.029 -> 0.973 ms
// at the headline size against SGR_BWD_PIPE3 = 0 (two stages of 128 records, two CTA barriers per batch).
#ifndef SGR_BWD_PIPE3
#define SGR_BWD_PIPE3 1
#endif
#ifndef SGR_BWD_NS
#define SGR_BWD_NS 3
#endif
#ifndef SGR_BWD_B
#define SGR_BWD_B (SGR_BWD_PIPE3 ? 96 : 128)
#endif
constexpr int BWD_B = SGR_BWD_B;  // Gaussians per shared-memory batch (loader threads: the first BWD_B of the CTA)
constexpr int BWD_NW = 8;   // warps per CTA (16x16 pixels)

// Blend accumulators (zeroed by a memset before the blend pass):
//   gacc   f32[P][8]  (dmean2D.x, dmean2D.y, dconic.x, dconic.y | dconic.w, dopacity, -, -)   in the scratch buffer
//   dcol   f32[P][3]  dL/dRGB, accumulated straight into the caller's dL_dcolors output.  With SH colours the
//          channels the forward clamped (forward.cu:63-70) are masked here, so this array IS the per-view
//          SH factor the view-parallel step exchanges, complete as soon as the blend pass ends.
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float *addr, float a, float b)
{
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// ------------------------------------------------------------------------------------------------
// blend backward.  One CTA per 16x16 tile, one thread per pixel, one warp per 8x4 pixel block (the
// forward's geometry); the tile's list is walked back to front over its first max(n_contrib)
// positions only, a batch of BWD_B records at a time, with the footprint masks deciding which warp
// visits which splat (sgr_internal.cuh).  Splat records move through a three-stage cp.async (LDGSTS) pipeline
// with full / empty mbarriers: the copies of batch b+1 are in flight while batch b is processed, and no warp
// ever waits for a sibling warp, only for data.
//
// Reduction.  The 9 per-pair values are NOT reduced across the warp with shuffles.  Phase 1 (lane =
// pixel) parks S = dL/dG * G and the colour weight alpha*T of every contributing splat in a per-warp
// [pixel][slot] shared-memory tile; after CH splats the warp turns around (phase 2, lane = slot x half
// block): each lane walks the 16 pixels of its half, accumulates RAW moments of S about the half's
// centre -- the pixel offsets are compile-time constants, so a pixel costs 3 FMAs for the moments and 3
// for the colours -- shifts them to the splat's centre once, combines the two halves with one xor-16
// shuffle per value and adds the result to the per-Gaussian accumulators with two red.v4 + one scalar
// reduction (chunks are ~94 % full on the headline scene).
//
// alpha is recomputed as in the forward; G = exp(power) comes from ex2.approx (MUFU.EX2) instead of
// libdevice's expf -- 2 instructions instead of 8 -- except within 1e-5 (relative) of the 1/255 threshold,
// where the exact expf decides, so the set of contributing pairs is the forward's.
// ------------------------------------------------------------------------------------------------
#ifdef SGR_BLEND_STATS
__device__ unsigned long long g_bwd_stats[8];
#define BWD_STAT(k, v) atomicAdd(&g_bwd_stats[k], (unsigned long long)(v))
int read_bwd_stats(unsigned long long *out, int reset)
{
    SGR_CUDA(cudaMemcpyFromSymbol(out, g_bwd_stats, sizeof(unsigned long long) * 8));
    if (reset) {
        unsigned long long z[8] = {};
        SGR_CUDA(cudaMemcpyToSymbol(g_bwd_stats, z, sizeof(z)));
    }
    return SGR_OK;
}
#endif
constexpr int CH = 16;             // splats per chunk
constexpr int CH_PITCH = CH + 1;   // float2 elements per pixel row: odd pitch = conflict-free transpose
// per-warp chunk scratch (bytes from its base): pair[32][CH_PITCH] float2 | meta[CH] x 32 B | dp[32] float4
//   meta: (x, y, conic a, conic b) (conic c, opacity, clamp bits, id) -- what phase 2 needs of a parked splat
//   dp:   (dL/dpixel r g b, -T_final <bg, dL/dpixel>) of the block's pixels
// (57.2 KB per CTA in total with three record stages: four CTAs per SM fit, with 64 registers per thread)
constexpr uint32_t CB_META = 32 * CH_PITCH * 8, CB_DP = CB_META + CH * 32, CB_BYTES = CB_DP + 32 * 16;
// CTA shared-memory map (dynamic): the record stages, their membership words, the mbarriers, chunk scratch
constexpr uint32_t BWD_NS = SGR_BWD_PIPE3 ? SGR_BWD_NS : 2;  // record stages
constexpr uint32_t SM_A = 0, SM_B = SM_A + BWD_NS * BWD_B * 16, SM_C = SM_B + BWD_NS * BWD_B * 16,
                   SM_MEMBER = SM_C + BWD_NS * BWD_B * 16,
                   SM_BAR = SM_MEMBER + (SGR_BWD_PIPE3 ? BWD_NS : 1) * BWD_NW * (BWD_B / 32) * 4,  // PIPE3: full[3], empty[3]
                   SM_LAST = SM_BAR + (SGR_BWD_PIPE3 ? 2 * BWD_NS * 8 : 0), SM_CHUNK = SM_LAST + 16,
                   BWD_SMEM_BYTES = SM_CHUNK + BWD_NW * CB_BYTES;
static_assert(SM_BAR % 8 == 0 && BWD_B % 32 == 0, "mbarriers are 8-byte aligned; a batch is whole membership words");
static_assert(4 * (BWD_SMEM_BYTES + 1024) <= 228 * 1024, "four CTAs of the backward blend must fit one SM's shared memory");

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrives on `bar` when all cp.async copies this thread has issued so far have landed (does not change the
// barrier's pending count: the expected count must include it)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float ex2_approx(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// phase 2: reduce the parked chunk (nf splats) over the block's pixels and add it to the Gaussians'
// accumulators.  Lane = (chunk slot, half): half h holds block rows 2h, 2h+1 = phase-1 lanes 16h..16h+15.
__device__ __forceinline__ void chunk_flush(uint32_t cb, int nf, float blk_x0, float blk_y0, float ddelx_dx,
                                            float ddely_dy, float *__restrict__ gacc, float *__restrict__ dcol)
{
    const unsigned lane = threadIdx.x & 31u;
    const int slot = (int)(lane & (CH - 1)), half = (int)(lane >> 4);
    __syncwarp();
    const float4 M0 = lds128(cb + CB_META + slot * 32), M1 = lds128(cb + CB_META + slot * 32 + 16);
    const uint32_t clamp = __float_as_uint(M1.z), id = __float_as_uint(M1.w);
    // d = mean - pixel = (ax - cx_i, ay - cy_i) with the half's centre as origin: cx_i = (i & 7) - 3.5,
    // cy_i = (i >> 3) - 0.5
    const float ax = M0.x - (blk_x0 + 3.5f), ay = M0.y - (blk_y0 + 2.0f * (float)half + 0.5f);
    float s_lo = 0.f, s_hi = 0.f, t_lo = 0.f, t_hi = 0.f, u = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
    const uint32_t pa = cb + (half * 16 * CH_PITCH + slot) * 8, da = cb + CB_DP + half * 16 * 16;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const float2 pr = lds64(pa + i * CH_PITCH * 8);
        const float4 d = lds128(da + i * 16);
        const float cx = (float)(i & 7) - 3.5f;
        if (i < 8) {
            s_lo += pr.x;
            t_lo = fmaf(pr.x, cx, t_lo);
        } else {
            s_hi += pr.x;
            t_hi = fmaf(pr.x, cx, t_hi);
        }
        u = fmaf(pr.x, cx * cx, u);
        c0 = fmaf(pr.y, d.x, c0);
        c1 = fmaf(pr.y, d.y, c1);
        c2 = fmaf(pr.y, d.z, c2);
    }
    const float s0 = s_lo + s_hi, sx = t_lo + t_hi, sy = 0.5f * (s_hi - s_lo), sxy = 0.5f * (t_hi - t_lo);
    float m0 = s0;
    float m1 = fmaf(ax, s0, -sx);                                    // sum S dx
    float m2 = fmaf(ay, s0, -sy);                                    // sum S dy
    float m3 = fmaf(ax, fmaf(ax, s0, -2.0f * sx), u);                // sum S dx^2
    float m4 = fmaf(ax, m2, fmaf(-ay, sx, sxy));                     // sum S dx dy
    float m5 = fmaf(ay, fmaf(ay, s0, -2.0f * sy), 0.25f * s0);       // sum S dy^2
    m0 += __shfl_xor_sync(0xffffffffu, m0, 16);
    m1 += __shfl_xor_sync(0xffffffffu, m1, 16);
    m2 += __shfl_xor_sync(0xffffffffu, m2, 16);
    m3 += __shfl_xor_sync(0xffffffffu, m3, 16);
    m4 += __shfl_xor_sync(0xffffffffu, m4, 16);
    m5 += __shfl_xor_sync(0xffffffffu, m5, 16);
    c0 += __shfl_xor_sync(0xffffffffu, c0, 16);
    c1 += __shfl_xor_sync(0xffffffffu, c1, 16);
    c2 += __shfl_xor_sync(0xffffffffu, c2, 16);
    if (slot < nf) {
        // moments -> gradients (backward.cu:537-554): dG/ddel = -G (Q d)
        float *g = gacc + (size_t)id * 8, *dc = dcol + (size_t)id * 3;
        if (half == 0) {
            const float gmx = -(M0.z * m1 + M0.w * m2) * ddelx_dx;
            const float gmy = -(M1.x * m2 + M0.w * m1) * ddely_dy;
            red_add_v4(g, gmx, gmy, -0.5f * m3, -0.5f * m4);
            if (!(clamp & 1u)) atomicAdd(dc, c0);
        } else {
            red_add_v2(g + 4, -0.5f * m5, __fdividef(m0, M1.y));
            if (!(clamp & 2u)) atomicAdd(dc + 1, c1);
            if (!(clamp & 4u)) atomicAdd(dc + 2, c2);
        }
    }
    __syncwarp();
}

template <bool packed>
__global__ void __launch_bounds__(256, 4) blend_backward_kernel(
    const uint32_t *__restrict__ tile_order, const uint32_t *__restrict__ tile_start, const uint32_t *__restrict__ plist,
    const float4 *__restrict__ rec, int W, int H, int gx, const float *__restrict__ bg,
    const float *__restrict__ final_Ts, const uint32_t *__restrict__ n_contrib, const float *__restrict__ dL_dpixels,
    float *__restrict__ gacc, float *__restrict__ dcol)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int tile = (int)tile_order[blockIdx.x];
    const uint32_t lo = tile_start[tile], hi = tile_start[tile + 1];
    if (hi == lo) return;
    const int tile_y = tile / gx, tile_x = tile - tile_y * gx;
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31, wid = tid >> 5;
    const int tx0 = tile_x * SGR_TILE, ty0 = tile_y * SGR_TILE;
    const int bx0 = tx0 + (int)(wid & 1u) * 8, by0 = ty0 + (int)(wid >> 1) * 4;
    const uint32_t pxi = bx0 + (lane & 7u), pyi = by0 + (lane >> 3);
    const bool inside = pxi < (uint32_t)W && pyi < (uint32_t)H;
    const uint32_t sm = smem_addr_pinned(s_raw);
    int *s_tile_last = (int *)(s_raw + SM_LAST);
    uint32_t(*s_member)[BWD_B / 32] = (uint32_t(*)[BWD_B / 32])(s_raw + SM_MEMBER);
    if (tid == 0) *s_tile_last = 0;

    const size_t pix = (size_t)pyi * W + pxi, plane = (size_t)H * W;
    float T = inside ? final_Ts[pix] : 0.0f;
    const int last_contributor = inside ? (int)n_contrib[pix] : 0;
    const uint32_t cb = sm + SM_CHUNK + wid * CB_BYTES;
    {
        float dp0 = 0.f, dp1 = 0.f, dp2 = 0.f;
        if (inside) {
            dp0 = dL_dpixels[pix];
            dp1 = dL_dpixels[plane + pix];
            dp2 = dL_dpixels[2 * plane + pix];
        }
        // (dL/dpixel, -T_final * <bg, dL/dpixel>) of this pixel: phase 2 reads the block's table; the
        // hit path of phase 1 reloads its own entry instead of pinning four registers across the loop
        sts128(cb + CB_DP + lane * 16, dp0, dp1, dp2, -T * fmaf(bg[2], dp2, fmaf(bg[1], dp1, bg[0] * dp0)));
    }
    // the pixel centre, pinned: under register pressure nvcc otherwise re-derives it from special
    // registers inside the splat loop
    float pxf = (float)pxi, pyf = (float)pyi;
    asm volatile("mov.f32 %0, %0;" : "+f"(pxf));
    asm volatile("mov.f32 %0, %0;" : "+f"(pyf));
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
    // Splats behind the deepest last contributor of the warp (of the tile) cannot touch any of its
    // pixels: the tile only walks list positions [0, tile_last), each warp only [0, warp_last).
    const int warp_last = __reduce_max_sync(0xffffffffu, last_contributor);
    __syncthreads();
    if (lane == 0 && warp_last > 0) atomicMax(s_tile_last, warp_last);
    __syncthreads();
    const int n = *s_tile_last;
    if (n == 0) return;
    int nfill = 0;
    const float blk_x0 = (float)bx0, blk_y0 = (float)by0;

    // record pipeline (threads 0..BWD_B-1 own one slot of every batch): issue() starts the async
    // copies of one splat record into buffer `buf`; list words are fetched one batch further ahead
    const bool loader = tid < BWD_B;
    auto fetch = [&](int b0) -> uint32_t {
        // back to front; dead instances (packed list, empty footprint) are dropped here: 0xffffffff
        if (!(loader && b0 + tid < n)) return 0xffffffffu;
        const uint32_t w = plist[lo + (uint32_t)(n - 1 - (b0 + tid))];
        return (packed && (w & 0xffu) == 0u) ? 0xffffffffu : w;
    };
    auto issue = [&](uint32_t w, int buf) {
        if (w != 0xffffffffu) {
            const uint32_t id = packed ? (w >> 8) : w;
            const float4 *r = rec + (size_t)id * 3;
            const uint32_t e = buf * BWD_B + tid;
            cp_async16_a(sm + SM_A + e * 16, r);
            cp_async16_a(sm + SM_B + e * 16, r + 1);
            cp_async16_a(sm + SM_C + e * 16, r + 2);
        }
        cp_async_commit();
    };
    // membership words of the batch in stage `st` from the loader threads' masks
    auto publish_members = [&](uint32_t mask, uint32_t(*set)[BWD_B / 32]) {
#pragma unroll
        for (int blk = 0; blk < BWD_NW; blk++) {
            const uint32_t word = __ballot_sync(0xffffffffu, (mask >> blk) & 1u);
            if (lane == 0) set[blk][wid] = word;
        }
    };
#if SGR_BWD_PIPE3
    constexpr int NS = (int)BWD_NS;
    uint64_t *bars = (uint64_t *)(s_raw + SM_BAR);  // full[0..NS) then empty[0..NS)
    if (tid == 0) {
        for (int st = 0; st < NS; st++) {
            mbar_init(&bars[st], 2 * BWD_B);      // per loader thread: one async arrive (copies landed) + one plain
            mbar_init(&bars[NS + st], BWD_NW);    // one arrive per warp when it is done with the stage
        }
        mbar_fence_init();
    }
    __syncthreads();
    const int nb = (n + BWD_B - 1) / BWD_B;
    // loader warps: fill stage b % NS with batch b (waits until every warp has released the stage's previous batch)
    auto produce = [&](int b) {
        const int st = b % NS;
        if (b >= NS) mbar_wait(&bars[NS + st], (uint32_t)((b / NS - 1) & 1));
        const uint32_t w = fetch(b * BWD_B);
        uint32_t mask = 0;
        if (w != 0xffffffffu) {
            const uint32_t id = packed ? (w >> 8) : w;
            const float4 *r = rec + (size_t)id * 3;
            const uint32_t e = st * BWD_B + tid;
            cp_async16_a(sm + SM_A + e * 16, r);
            cp_async16_a(sm + SM_B + e * 16, r + 1);
            cp_async16_a(sm + SM_C + e * 16, r + 2);
            if (packed) mask = w & 0xffu;
        }
        if (!packed) {  // the mask needs the landed record
            cp_async_commit();
            cp_async_wait<0>();
            if (w != 0xffffffffu) {
                const float4 r0 = lds128(sm + SM_A + (st * BWD_B + tid) * 16), r1 = lds128(sm + SM_B + (st * BWD_B + tid) * 16);
                mask = block_mask_of_record(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, tx0, ty0);
            }
        }
        cp_async_mbar_arrive_noinc(&bars[st]);
        publish_members(mask, s_member + st * BWD_NW);
        mbar_arrive(&bars[st]);
    };
    if (loader) produce(0);
    for (int b = 0; b < nb; b++) {
        const int b0 = b * BWD_B, buf = b % NS;
        const uint32_t sa = sm + SM_A + buf * (BWD_B * 16), sb = sm + SM_B + buf * (BWD_B * 16),
                       sc = sm + SM_C + buf * (BWD_B * 16);
        if (loader && b + 1 < nb) produce(b + 1);
        mbar_wait(&bars[buf], (uint32_t)((b / NS) & 1));  // batch b has landed and its membership words are visible
        const uint32_t(*member)[BWD_B / 32] = s_member + buf * BWD_NW;
#else
    uint32_t w_cur = fetch(0);
    issue(w_cur, 0);
    uint32_t w_next = fetch(BWD_B);

    for (int b0 = 0, buf = 0; b0 < n; b0 += BWD_B, buf ^= 1) {
        const uint32_t sa = sm + SM_A + buf * (BWD_B * 16), sb = sm + SM_B + buf * (BWD_B * 16),
                       sc = sm + SM_C + buf * (BWD_B * 16);
        cp_async_wait<0>();
        __syncthreads();  // batch b0 has landed in `buf`; every warp is done with the other buffer
        if (loader) {
            uint32_t mask = 0;
            if (w_cur != 0xffffffffu) {
                if (packed) {
                    mask = w_cur & 0xffu;
                } else {
                    const float4 r0 = lds128(sa + tid * 16), r1 = lds128(sb + tid * 16);
                    mask = block_mask_of_record(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, tx0, ty0);
                }
            }
            publish_members(mask, s_member);
            // next batch's records start moving now; they are not needed before the next barrier
            issue(w_next, buf ^ 1);
            w_cur = w_next;
            w_next = fetch(b0 + 2 * BWD_B);
        }
        __syncthreads();
        const uint32_t(*member)[BWD_B / 32] = s_member;
#endif
        const int m = min(BWD_B, n - b0);
#ifdef SGR_BLEND_STATS
        if (tid == 0) BWD_STAT(6, m);
#endif
        const int jmin = n - b0 - warp_last;  // first batch slot whose list position is < warp_last
        // list position of slot j is n-1-(b0+j); it precedes this pixel's last contributor iff j > jlim
        const int jlim = n - 1 - b0 - last_contributor;
#pragma unroll 1
        for (int c = 0; c * 32 < m; c++) {
            uint32_t mw = member[wid][c];
            const int cut = jmin - c * 32;
            if (cut >= 32) mw = 0;
            else if (cut > 0) mw &= ~((1u << cut) - 1u);
#pragma unroll 1
            while (mw) {
                const int j = c * 32 + (__ffs(mw) - 1);
                mw &= mw - 1;
                const float4 A = lds128(sa + j * 16);
                const float4 B = lds128(sb + j * 16);
                const float dx = __fsub_rn(A.x, pxf), dy = __fsub_rn(A.y, pyf);
                const float power = splat_power(dx, dy, A.z, A.w, B.x);
                const bool valid = j > jlim && !(power > 0.0f) && !(power < B.y);
#ifdef SGR_BLEND_STATS
                {
                    const bool hit = valid && !(fminf(0.99f, __fmul_rn(B.z, expf(power))) < 1.0f / 255.0f);
                    const unsigned mc = __ballot_sync(0xffffffffu, valid), mh = __ballot_sync(0xffffffffu, hit);
                    const unsigned ml = __ballot_sync(0xffffffffu, j > jlim);
                    if (lane == 0) {
                        BWD_STAT(0, 1);
                        BWD_STAT(1, mc != 0);
                        BWD_STAT(2, mh != 0);
                        BWD_STAT(3, __popc(mc));
                        BWD_STAT(4, __popc(mh));
                        BWD_STAT(5, __popc(ml));
                    }
                }
#endif
                if (!__any_sync(0xffffffffu, valid)) continue;
                // S = dL/dG * G and the colour weight alpha*T of this pair; zero for lanes without one.
                // (A candidate warp almost always keeps a contributing lane, so there is no second vote:
                // an all-zero slot is harmless.)
                float S = 0.f, dchannel = 0.f;
                if (valid) {
                    float G = ex2_approx(power * 1.4426950408889634f);
                    float alpha = fminf(0.99f, B.z * G);
                    if (fabsf(alpha - 1.0f / 255.0f) < 4.0e-8f) {  // on the threshold: decide as the forward did
                        G = expf(power);
                        alpha = fminf(0.99f, __fmul_rn(B.z, G));
                    }
                    if (!(alpha < 1.0f / 255.0f)) {
                        const float2 Cc = lds64(sc + j * 16);
                        const float4 dp = lds128(cb + CB_DP + lane * 16);
                        const float om = 1.0f - alpha;  // in [0.01, 1): a bare MUFU.RCP suffices
                        float inv;
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(om));
                        T = T * inv;
                        dchannel = alpha * T;
                        // acc holds the colour blended behind this splat (backward.cu:493-499 keeps
                        // last_alpha / last_color and folds them in one iteration later; same
                        // operations on the same values, evaluated eagerly below)
                        float dL_dalpha = (B.w - acc0) * dp.x;
                        dL_dalpha = fmaf(Cc.x - acc1, dp.y, dL_dalpha);
                        dL_dalpha = fmaf(Cc.y - acc2, dp.z, dL_dalpha);
                        dL_dalpha *= T;
                        dL_dalpha = fmaf(dp.w, inv, dL_dalpha);
                        S = B.z * dL_dalpha * G;
                        acc0 = fmaf(alpha, B.w, om * acc0);
                        acc1 = fmaf(alpha, Cc.x, om * acc1);
                        acc2 = fmaf(alpha, Cc.y, om * acc2);
                    }
                }
                // park: every lane its pair; the record + id by all lanes alike (same address, same
                // value: one wavefront, no branch)
                sts64(cb + (lane * CH_PITCH + nfill) * 8, S, dchannel);
                const uint32_t ma = cb + CB_META + nfill * 32;
                sts128(ma, A.x, A.y, A.z, A.w);
                {
                    const float2 Cm = lds64(sc + j * 16 + 8);  // (clamp bits, id)
                    sts128(ma + 16, B.x, B.z, Cm.x, Cm.y);
                }
                if (++nfill == CH) {
                    chunk_flush(cb, CH, blk_x0, blk_y0, 0.5f * W, 0.5f * H, gacc, dcol);
                    nfill = 0;
                }
            }
        }
#if SGR_BWD_PIPE3
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[NS + buf]);  // this warp no longer reads the stage
#endif
    }
    if (nfill) chunk_flush(cb, nfill, blk_x0, blk_y0, 0.5f * W, 0.5f * H, gacc, dcol);
}

// ------------------------------------------------------------------------------------------------
// per-Gaussian backward (cov2D inverse, projection, SH, scale/rotation), one fused pass.
// HBM-bound: every input array is staged into shared memory with coalesced transfers (1-D bulk
// copies through the TMA engine for the contiguous arrays, 16-byte cp.async into padded rows for
// the SH block), each thread works on its own Gaussian out of shared memory, the SH gradient
// overwrites the SH row in place, and all outputs leave through coalesced stores.
// ------------------------------------------------------------------------------------------------
struct PreBwdArgs {
    int p0, p1;  // Gaussian range of this launch (one chunk of the per-Gaussian pass); p0 % PB_T == 0
    const float *means, *scales, *rots, *shs, *cov_pre;
    ViewConsts v;
    const int32_t *radii;
    const float *gacc;  // [P][8] blend accumulators
    const float *dcol;  // [P][3] accumulated dL/dRGB (= the dL_dcolors output; clamp-masked with SH)
    float *dmeans2D, *dopacity, *dmeans3D, *dcov3D, *dsh, *dscales, *drots;
    // optional f32[P][11]: (dL_dmeans3D 3 | dL_dopacity 1 | dL_dscales 3 | dL_drotations 4) of a Gaussian as ONE
    // 44-byte record instead of the four arrays (the view-parallel step all-reduces these records chunk by chunk)
    float *rec11;
    int in_bulk_ok, out_bulk_ok, sh_stride, sh_vec;
    // raw-parameter mode (SgrGaussians.activations): scales / rots are the model's raw parameters, shs its DC
    // array and sh_rest the others; dopacity / dscales / drots / dsh / dsh_rest are gradients of the RAW parameters
    const float *sh_rest;
    const float4 *rec;  // splat records of the forward: rec[3i+1].z = the activated opacity
    float *dsh_rest;
    // view-parallel peer mode (csrc/sgr_peer.cu): dsh rows leave as the sum over ALL views' SH gradients.  view_tab is a
    // device table of nviews pointers to the views' factor blocks (f32[3P] dL/dRGB + the view's camera position at
    // campos_off), entries of other ranks pointing into THEIR memory over NVLink; view my_view is this launch's own
    // (its factor is `dcol`).  dsh_scale multiplies every dsh row (1/num_views for a mean over the batch).
    const float *const *view_tab;
    int nviews, my_view;
    size_t campos_off;
    float dsh_scale;
    // peer mode, records: CTA b of this launch (a chunk of the pass) stores its 64 records into the staging array of the
    // rank that OWNS them -- owner = b * nviews / gridDim.x, stage_tab[owner] = that rank's array for records coming
    // from this rank -- with one TMA bulk store (a posted write over NVLink): the reduce-scatter half of the records'
    // all-reduce is fused into this kernel's write-out.  NULL: records go to rec11.
    float *const *stage_tab;
    int pf_off;  // floats: where the other views' factor blocks land in shared memory (MULTI)
};

#define SH_C0 0.28209479177387814f
#define SH_C1 0.4886025119029199f
__constant__ float b_SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                 -1.0925484305920792f, 0.5462742152960396f};
__constant__ float b_SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                 0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                 -0.5900435899266435f};

struct f3 {
    float x, y, z;
};

// SH backward (backward.cu:20-139) on one shared-memory row [M][3], in place: row k is read once
// (t_k = <sh_k, dL/dRGB> feeds the view-direction gradient) and then overwritten by dL/dsh_k.
// `dc` is coefficient 0's row of three, `rest` coefficient 1's (one combined row: rest = dc + 3; raw-parameter
// mode: two arrays).
template <bool OTHERS = false>
__device__ __forceinline__ void sh_backward_inplace(int deg, int M, float *dc, float *rest, f3 dir_orig, f3 dRGB,
                                                    f3 &dmean, float out_scale = 1.0f, const float *oacc = nullptr)
{
    const f3 dS = {dRGB.x * out_scale, dRGB.y * out_scale, dRGB.z * out_scale};  // x 1.0f is exact
    const float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
    const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
    float ddx = 0.f, ddy = 0.f, ddz = 0.f;
#define SHB(k, w, cx, cy, cz)                                                          \
    {                                                                                  \
        float *q_ = (k) == 0 ? dc : rest + ((k) - 1) * 3;                              \
        const float t_ = q_[0] * dRGB.x + q_[1] * dRGB.y + q_[2] * dRGB.z;             \
        ddx += (cx) * t_;                                                              \
        ddy += (cy) * t_;                                                              \
        ddz += (cz) * t_;                                                              \
        const float w_ = (w);                                                          \
        if (OTHERS) { /* the other views' SH gradients of this Gaussian (registers) join the row here */ \
            q_[0] = fmaf(w_, dS.x, oacc[(k) * 3]);                                     \
            q_[1] = fmaf(w_, dS.y, oacc[(k) * 3 + 1]);                                 \
            q_[2] = fmaf(w_, dS.z, oacc[(k) * 3 + 2]);                                 \
        } else {                                                                       \
            q_[0] = w_ * dS.x;                                                         \
            q_[1] = w_ * dS.y;                                                         \
            q_[2] = w_ * dS.z;                                                         \
        }                                                                              \
    }
    SHB(0, SH_C0, 0.f, 0.f, 0.f);
    if (deg > 0) {
        SHB(1, -SH_C1 * y, 0.f, -SH_C1, 0.f);
        SHB(2, SH_C1 * z, 0.f, 0.f, SH_C1);
        SHB(3, -SH_C1 * x, -SH_C1, 0.f, 0.f);
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            SHB(4, b_SH_C2[0] * xy, b_SH_C2[0] * y, b_SH_C2[0] * x, 0.f);
            SHB(5, b_SH_C2[1] * yz, 0.f, b_SH_C2[1] * z, b_SH_C2[1] * y);
            SHB(6, b_SH_C2[2] * (2.f * zz - xx - yy), b_SH_C2[2] * 2.f * -x, b_SH_C2[2] * 2.f * -y, b_SH_C2[2] * 2.f * 2.f * z);
            SHB(7, b_SH_C2[3] * xz, b_SH_C2[3] * z, 0.f, b_SH_C2[3] * x);
            SHB(8, b_SH_C2[4] * (xx - yy), b_SH_C2[4] * 2.f * x, b_SH_C2[4] * 2.f * -y, 0.f);
            if (deg > 2) {
                SHB(9, b_SH_C3[0] * y * (3.f * xx - yy), b_SH_C3[0] * 3.f * 2.f * xy, b_SH_C3[0] * 3.f * (xx - yy), 0.f);
                SHB(10, b_SH_C3[1] * xy * z, b_SH_C3[1] * yz, b_SH_C3[1] * xz, b_SH_C3[1] * xy);
                SHB(11, b_SH_C3[2] * y * (4.f * zz - xx - yy), b_SH_C3[2] * -2.f * xy,
                    b_SH_C3[2] * (-3.f * yy + 4.f * zz - xx), b_SH_C3[2] * 4.f * 2.f * yz);
                SHB(12, b_SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy), b_SH_C3[3] * -3.f * 2.f * xz,
                    b_SH_C3[3] * -3.f * 2.f * yz, b_SH_C3[3] * 3.f * (2.f * zz - xx - yy));
                SHB(13, b_SH_C3[4] * x * (4.f * zz - xx - yy), b_SH_C3[4] * (-3.f * xx + 4.f * zz - yy),
                    b_SH_C3[4] * -2.f * xy, b_SH_C3[4] * 4.f * 2.f * xz);
                SHB(14, b_SH_C3[5] * z * (xx - yy), b_SH_C3[5] * 2.f * xz, b_SH_C3[5] * -2.f * yz, b_SH_C3[5] * (xx - yy));
                SHB(15, b_SH_C3[6] * x * (xx - 3.f * yy), b_SH_C3[6] * 3.f * (xx - yy), b_SH_C3[6] * -3.f * 2.f * xy, 0.f);
            }
        }
    }
#undef SHB
    // rows above the active degree stay zero (the reference returns zero-filled dL_dsh)
    for (int k = (deg + 1) * (deg + 1) * 3 - 3; k < (M - 1) * 3; k++) rest[k] = 0.f;
    // dnormvdv (auxiliary.h:107-117)
    const f3 o = dir_orig;
    const float sum2 = o.x * o.x + o.y * o.y + o.z * o.z;
    const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    dmean.x += ((+sum2 - o.x * o.x) * ddx - o.y * o.x * ddy - o.z * o.x * ddz) * invsum32;
    dmean.y += (-o.x * o.y * ddx + (sum2 - o.y * o.y) * ddy - o.z * o.y * ddz) * invsum32;
    dmean.z += (-o.x * o.z * ddx - o.y * o.z * ddy + (sum2 - o.z * o.z) * ddz) * invsum32;
}

// Another view's SH gradient of one Gaussian, accumulated in registers: acc[3k + c] += basis_k(normalize(dir)) * g_c
// (the dL_dsh part of backward.cu:20-139 is this outer product per view).
__device__ __forceinline__ void sh_accumulate(int deg, float *acc, f3 dir_orig, f3 g)
{
    const float inv = 1.0f / sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
    const float x = dir_orig.x * inv, y = dir_orig.y * inv, z = dir_orig.z * inv;
#define SHA(k, w)                                           \
    {                                                       \
        const float w_ = (w);                               \
        acc[(k) * 3] = fmaf(w_, g.x, acc[(k) * 3]);         \
        acc[(k) * 3 + 1] = fmaf(w_, g.y, acc[(k) * 3 + 1]); \
        acc[(k) * 3 + 2] = fmaf(w_, g.z, acc[(k) * 3 + 2]); \
    }
    SHA(0, SH_C0);
    if (deg > 0) {
        SHA(1, -SH_C1 * y);
        SHA(2, SH_C1 * z);
        SHA(3, -SH_C1 * x);
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            SHA(4, b_SH_C2[0] * xy);
            SHA(5, b_SH_C2[1] * yz);
            SHA(6, b_SH_C2[2] * (2.f * zz - xx - yy));
            SHA(7, b_SH_C2[3] * xz);
            SHA(8, b_SH_C2[4] * (xx - yy));
            if (deg > 2) {
                SHA(9, b_SH_C3[0] * y * (3.f * xx - yy));
                SHA(10, b_SH_C3[1] * xy * z);
                SHA(11, b_SH_C3[2] * y * (4.f * zz - xx - yy));
                SHA(12, b_SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
                SHA(13, b_SH_C3[4] * x * (4.f * zz - xx - yy));
                SHA(14, b_SH_C3[5] * z * (xx - yy));
                SHA(15, b_SH_C3[6] * x * (xx - 3.f * yy));
            }
        }
    }
#undef SHA
}

#ifndef SGR_PB_T
#define SGR_PB_T 64
#endif
constexpr int PB_VB = 7;  // other views whose factors a thread fetches at the top of the kernel (8 ranks: all of them)
constexpr int PB_T = SGR_PB_T;
// dynamic shared memory carve-up (floats): inputs, outputs, then the SH block
constexpr int PB_GACC = 0;                     // PB_T*8
constexpr int PB_DCOL = PB_GACC + PB_T * 8;    // PB_T*3
constexpr int PB_MEANS = PB_DCOL + PB_T * 3;   // PB_T*3
constexpr int PB_SCALES = PB_MEANS + PB_T * 3;
constexpr int PB_ROTS = PB_SCALES + PB_T * 3;  // PB_T*4 (16B aligned: offset is a multiple of 4 floats)
constexpr int PB_COV = PB_ROTS + PB_T * 4;     // PB_T*6
constexpr int PB_O_M2D = PB_COV + PB_T * 6;    // outputs
constexpr int PB_O_COV = PB_O_M2D + PB_T * 3;
// the 11 all-reduce-bound floats per Gaussian: four arrays (OPA | M3D | SCL | ROT) or PB_T records of 11
constexpr int PB_O_R11 = PB_O_COV + PB_T * 6;
constexpr int PB_O_OPA = PB_O_R11;
constexpr int PB_O_M3D = PB_O_OPA + PB_T;
constexpr int PB_O_SCL = PB_O_M3D + PB_T * 3;
constexpr int PB_O_ROT = PB_O_SCL + PB_T * 3;  // multiple of 4 floats
constexpr int PB_SHDC = PB_O_R11 + PB_T * 11;  // raw mode: the DC coefficients / their gradient, PB_T*3
constexpr int PB_SH = PB_SHDC + PB_T * 3 + (PB_T & 3 ? 4 - (PB_T & 3) : 0);  // PB_T * sh_stride
static_assert(PB_T % 4 == 0, "every region starts on a 16-byte boundary");
static_assert(PB_ROTS % 4 == 0 && PB_O_ROT % 4 == 0 && PB_SH % 4 == 0, "16-byte alignment of float4 regions");

// 1-D bulk async copy shared -> global (TMA store, no tensor map); bytes % 16 == 0, both sides 16-byte aligned
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void pb_stage(float *dst, const float *__restrict__ src, int n)
{
    for (int i = threadIdx.x; i < n; i += PB_T) dst[i] = __ldg(src + i);
}
__device__ __forceinline__ void pb_flush(float *__restrict__ dst, const float *src, int n)
{
    for (int i = threadIdx.x; i < n; i += PB_T) dst[i] = src[i];
}

// MULTI (view-parallel peer mode, never with RAW): dsh rows leave as the sum over all views, see PreBwdArgs::view_tab
template <bool RAW, bool MULTI = false>
__global__ void __launch_bounds__(PB_T) preprocess_backward_kernel(const PreBwdArgs a)
{
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x;
    const int base = a.p0 + blockIdx.x * PB_T;
    const int n = min(PB_T, a.p1 - base);
    const int i = base + tid;
    const int M = a.v.M;
    const bool full = (n == PB_T);

    if (a.in_bulk_ok && full) {
        if (tid == 0) {
            mbar_init(&s_bar, 1);
            mbar_fence_init();
            uint32_t bytes = PB_T * 32 + PB_T * 12;
            if (a.shs) bytes += PB_T * 12;  // dL/dRGB is only an input of the SH backward
            if (RAW && a.shs) bytes += PB_T * 12 + PB_T * (uint32_t)(a.v.M - 1) * 12;
            if (a.scales) bytes += PB_T * 12 + PB_T * 16;
            if (a.cov_pre) bytes += PB_T * 24;
            // peer mode: the other views' factor blocks of these 64 Gaussians, one bulk copy per view straight out of
            // its owner GPU's memory (TMA over NVLink): the factors' all-gather is fused into this kernel's staging.
            // (A separate, later-waited barrier for them measured slower: 0.308 vs 0.257 ms for 1.6 M Gaussians.)
            const int pulled = MULTI ? min(a.nviews - 1, PB_VB) : 0;
            bytes += (uint32_t)pulled * PB_T * 12;
            mbar_expect_tx(&s_bar, bytes);
            for (int u = 0; u < pulled; u++) {
                const int v = u + (u >= a.my_view ? 1 : 0);
                bulk_g2s(sm + a.pf_off + u * PB_T * 3, a.view_tab[v] + (size_t)base * 3, PB_T * 12, &s_bar);
            }
            if (RAW && a.shs) {
                bulk_g2s(sm + PB_SHDC, a.shs + (size_t)base * 3, PB_T * 12, &s_bar);
                if (a.v.M > 1)
                    bulk_g2s(sm + PB_SH, a.sh_rest + (size_t)base * (a.v.M - 1) * 3, PB_T * (uint32_t)(a.v.M - 1) * 12, &s_bar);
            }
            bulk_g2s(sm + PB_GACC, a.gacc + (size_t)base * 8, PB_T * 32, &s_bar);
            if (a.shs) bulk_g2s(sm + PB_DCOL, a.dcol + (size_t)base * 3, PB_T * 12, &s_bar);
            bulk_g2s(sm + PB_MEANS, a.means + (size_t)base * 3, PB_T * 12, &s_bar);
            if (a.scales) {
                bulk_g2s(sm + PB_SCALES, a.scales + (size_t)base * 3, PB_T * 12, &s_bar);
                bulk_g2s(sm + PB_ROTS, a.rots + (size_t)base * 4, PB_T * 16, &s_bar);
            }
            if (a.cov_pre) bulk_g2s(sm + PB_COV, a.cov_pre + (size_t)base * 6, PB_T * 24, &s_bar);
        }
    } else {
        pb_stage(sm + PB_GACC, a.gacc + (size_t)base * 8, n * 8);
        if (a.shs) pb_stage(sm + PB_DCOL, a.dcol + (size_t)base * 3, n * 3);
        pb_stage(sm + PB_MEANS, a.means + (size_t)base * 3, n * 3);
        if (a.scales) {
            pb_stage(sm + PB_SCALES, a.scales + (size_t)base * 3, n * 3);
            pb_stage(sm + PB_ROTS, a.rots + (size_t)base * 4, n * 4);
        }
        if (a.cov_pre) pb_stage(sm + PB_COV, a.cov_pre + (size_t)base * 6, n * 6);
        if (RAW && a.shs) {
            pb_stage(sm + PB_SHDC, a.shs + (size_t)base * 3, n * 3);
            pb_stage(sm + PB_SH, a.sh_rest + (size_t)base * (M - 1) * 3, n * (M - 1) * 3);
        }
        if (MULTI) {
            const int pulled = min(a.nviews - 1, PB_VB);
            for (int u = 0; u < pulled; u++) {
                const int v = u + (u >= a.my_view ? 1 : 0);
                pb_stage(sm + a.pf_off + u * PB_T * 3, a.view_tab[v] + (size_t)base * 3, n * 3);
            }
        }
    }
    float *s_sh = sm + PB_SH;
    const int row_f = RAW ? (M - 1) * 3 : M * 3;  // floats per row of the SH block in shared memory
    if (!RAW && a.shs) {
        const float *src = a.shs + (size_t)base * row_f;
        if (a.sh_vec) {
            const int row_v = row_f >> 2, total = n * row_v;
            int r = tid / row_v, c = tid - r * row_v;
            const int dr = PB_T / row_v, dc = PB_T - dr * row_v;
            for (int k = tid; k < total; k += PB_T) {
                // (skipping the rows of culled Gaussians saves 8 % of this kernel's DRAM bytes but the per-chunk
                // radius test costs more than that: 0.384 vs 0.333 ms at the headline size, round 2)
                cp_async16(s_sh + r * a.sh_stride + c * 4, src + (size_t)k * 4);
                r += dr;
                c += dc;
                if (c >= row_v) {
                    c -= row_v;
                    r++;
                }
            }
        } else {
            const int total = n * row_f;
            for (int k = tid; k < total; k += PB_T) {
                const int r = k / row_f, c = k - r * row_f;
                cp_async4(s_sh + r * a.sh_stride + c, src + k);
            }
        }
        cp_async_commit();
    }
    int radius = 0;
    if (tid < n) radius = a.radii[i];
    const int others = MULTI ? a.nviews - 1 : 0;
    __syncthreads();
    if (a.in_bulk_ok && full) mbar_wait(&s_bar, 0);
    if (!RAW && a.shs) {
        cp_async_wait<0>();
        __syncthreads();
    }

    // peer mode: the other views' SH gradients of this Gaussian, summed in registers (48 accumulators); they join the
    // row when this view's gradient is written into it (or make up the whole row if the Gaussian is culled here)
    float oacc[MULTI ? 48 : 1];
    auto sum_other_views = [&]() {
        if (!MULTI) return;
#pragma unroll
        for (int k = 0; k < (MULTI ? 48 : 1); k++) oacc[k] = 0.f;
        if (others <= 0) return;
        const float mx = sm[PB_MEANS + tid * 3], my = sm[PB_MEANS + tid * 3 + 1], mz = sm[PB_MEANS + tid * 3 + 2];
        for (int u0 = 0; u0 < others; u0 += PB_VB) {
            float pf[PB_VB][3];
#pragma unroll
            for (int u = 0; u < PB_VB; u++) {
                const int uu = u0 + u, v = uu + (uu >= a.my_view ? 1 : 0);
                const bool on = uu < others;
                if (u0 == 0) {  // staged in shared memory by the bulk copies above
                    const float *g = sm + a.pf_off + u * PB_T * 3 + tid * 3;
                    pf[u][0] = on ? g[0] : 0.f;
                    pf[u][1] = on ? g[1] : 0.f;
                    pf[u][2] = on ? g[2] : 0.f;
                } else {        // more than PB_VB + 1 views: the rest straight from global / peer memory
                    const float *g = a.view_tab[on ? v : a.my_view] + (size_t)i * 3;
                    pf[u][0] = on ? __ldcg(g) : 0.f;
                    pf[u][1] = on ? __ldcg(g + 1) : 0.f;
                    pf[u][2] = on ? __ldcg(g + 2) : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < PB_VB; u++) {
                const f3 g = {pf[u][0] * a.dsh_scale, pf[u][1] * a.dsh_scale, pf[u][2] * a.dsh_scale};
                if (g.x == 0.f && g.y == 0.f && g.z == 0.f) continue;  // not visible in that view (or past the last view)
                const int uu = u0 + u, v = uu + (uu >= a.my_view ? 1 : 0);
                const float *cp = a.view_tab[v] + a.campos_off;
                sh_accumulate(a.v.D, oacc, {mx - __ldg(cp), my - __ldg(cp + 1), mz - __ldg(cp + 2)}, g);
            }
        }
    };
    if (MULTI && tid < n) sum_other_views();
    if (tid < n) {
        const bool vis = radius > 0;
        // the 11 all-reduce-bound outputs: four arrays, or one 44-byte record per Gaussian (stride 11: conflict-free)
        const bool aos = a.rec11 != nullptr;
        float *o_m2d = sm + PB_O_M2D + tid * 3, *o_cov = sm + PB_O_COV + tid * 6;
        float *o_m3d = aos ? sm + PB_O_R11 + tid * 11 : sm + PB_O_M3D + tid * 3;
        float *o_opa = aos ? o_m3d + 3 : sm + PB_O_OPA + tid;
        float *o_scl = aos ? o_m3d + 4 : sm + PB_O_SCL + tid * 3;
        float *o_rot = aos ? o_m3d + 7 : sm + PB_O_ROT + tid * 4;
        float *row = s_sh + tid * (RAW ? row_f : a.sh_stride);
        if (!vis) {
            o_m2d[0] = o_m2d[1] = o_m2d[2] = 0.f;
            *o_opa = 0.f;
            o_m3d[0] = o_m3d[1] = o_m3d[2] = 0.f;
#pragma unroll
            for (int k = 0; k < 6; k++) o_cov[k] = 0.f;
            o_scl[0] = o_scl[1] = o_scl[2] = 0.f;
            o_rot[0] = o_rot[1] = o_rot[2] = o_rot[3] = 0.f;
            if (a.shs && a.dsh) {
                if (MULTI) {
#pragma unroll
                    for (int k = 0; k < 48; k++)
                        if (k < row_f) row[k] = oacc[k];
                    for (int k = 48; k < row_f; k++) row[k] = 0.f;
                } else {
                    for (int k = 0; k < row_f; k++) row[k] = 0.f;
                }
                if (RAW) sm[PB_SHDC + tid * 3] = sm[PB_SHDC + tid * 3 + 1] = sm[PB_SHDC + tid * 3 + 2] = 0.f;
            }
        } else {
            const float4 *gr = (const float4 *)(sm + PB_GACC) + tid * 2;
            const float4 g0 = gr[0], g1 = gr[1];
            const float dmx = g0.x, dmy = g0.y;
            const float dcx = g0.z, dcy = g0.w, dcz = g1.x;
            o_m2d[0] = dmx;
            o_m2d[1] = dmy;
            o_m2d[2] = 0.f;
            // raw mode: d sigmoid = o (1 - o), with the activated opacity the forward stored in the splat record
            float act_o = 1.0f;
            if (RAW) {
                const float o = __ldg(&a.rec[(size_t)i * 3 + 1]).z;
                act_o = o * (1.0f - o);
            }
            *o_opa = g1.y * act_o;

            const float *vm = a.v.viewmatrix, *proj = a.v.projmatrix;
            const float mx = sm[PB_MEANS + tid * 3], my = sm[PB_MEANS + tid * 3 + 1], mz = sm[PB_MEANS + tid * 3 + 2];
            float c3[6];
            float4 q = make_float4(1, 0, 0, 0);
            float sc0 = 0, sc1 = 0, sc2 = 0, q_len = 1.0f;
            if (a.cov_pre) {
#pragma unroll
                for (int k = 0; k < 6; k++) c3[k] = sm[PB_COV + tid * 6 + k];
            } else {
                sc0 = sm[PB_SCALES + tid * 3], sc1 = sm[PB_SCALES + tid * 3 + 1], sc2 = sm[PB_SCALES + tid * 3 + 2];
                q = ((const float4 *)(sm + PB_ROTS))[tid];
                if (RAW) {  // the forward's activations, recomputed (sugar_model.py:417, 479)
                    sc0 = expf(sc0), sc1 = expf(sc1), sc2 = expf(sc2);
                    q_len = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
                    const float inv = __fdiv_rn(1.0f, q_len);
                    q = make_float4(q.x * inv, q.y * inv, q.z * inv, q.w * inv);
                }
                cov3d_from_scale_rot(sc0, sc1, sc2, a.v.scale_modifier, q, c3);
            }
            // ---- computeCov2DCUDA (backward.cu:144-274) ----
            float tx = xf_row(vm, 0, mx, my, mz), ty = xf_row(vm, 1, mx, my, mz);
            const float tzv = xf_row(vm, 2, mx, my, mz);
            const float h_x = a.v.focal_x, h_y = a.v.focal_y;
            const float limx = 1.3f * a.v.tanfovx, limy = 1.3f * a.v.tanfovy;
            const float txtz = tx / tzv, tytz = ty / tzv;
            tx = fminf(limx, fmaxf(-limx, txtz)) * tzv;
            ty = fminf(limy, fmaxf(-limy, tytz)) * tzv;
            const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
            const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
            const float J00 = h_x / tzv, J02 = -(h_x * tx) / (tzv * tzv), J11 = h_y / tzv, J12 = -(h_y * ty) / (tzv * tzv);
            float T0[3], T1[3];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                T0[r] = vm[4 * r] * J00 + vm[2 + 4 * r] * J02;
                T1[r] = vm[1 + 4 * r] * J11 + vm[2 + 4 * r] * J12;
            }
            const float V[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
            float p[3], qv[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                p[k] = T0[0] * V[k][0] + T0[1] * V[k][1] + T0[2] * V[k][2];
                qv[k] = T1[0] * V[k][0] + T1[1] * V[k][1] + T1[2] * V[k][2];
            }
            const float ca = T0[0] * p[0] + T0[1] * p[1] + T0[2] * p[2] + 0.3f;
            const float cb = T0[0] * qv[0] + T0[1] * qv[1] + T0[2] * qv[2];
            const float cc = T1[0] * qv[0] + T1[1] * qv[1] + T1[2] * qv[2] + 0.3f;
            const float denom = ca * cc - cb * cb;
            float dL_da = 0, dL_db = 0, dL_dc = 0;
            const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
            float dcov[6];
            if (denom2inv != 0) {
                dL_da = denom2inv * (-cc * cc * dcx + 2 * cb * cc * dcy + (denom - ca * cc) * dcz);
                dL_dc = denom2inv * (-ca * ca * dcz + 2 * ca * cb * dcy + (denom - ca * cc) * dcx);
                dL_db = denom2inv * 2 * (cb * cc * dcx - (denom + 2 * cb * cb) * dcy + ca * cb * dcz);
                dcov[0] = (T0[0] * T0[0] * dL_da + T0[0] * T1[0] * dL_db + T1[0] * T1[0] * dL_dc);
                dcov[3] = (T0[1] * T0[1] * dL_da + T0[1] * T1[1] * dL_db + T1[1] * T1[1] * dL_dc);
                dcov[5] = (T0[2] * T0[2] * dL_da + T0[2] * T1[2] * dL_db + T1[2] * T1[2] * dL_dc);
                dcov[1] = 2 * T0[0] * T0[1] * dL_da + (T0[0] * T1[1] + T0[1] * T1[0]) * dL_db + 2 * T1[0] * T1[1] * dL_dc;
                dcov[2] = 2 * T0[0] * T0[2] * dL_da + (T0[0] * T1[2] + T0[2] * T1[0]) * dL_db + 2 * T1[0] * T1[2] * dL_dc;
                dcov[4] = 2 * T0[2] * T0[1] * dL_da + (T0[1] * T1[2] + T0[2] * T1[1]) * dL_db + 2 * T1[1] * T1[2] * dL_dc;
            } else {
#pragma unroll
                for (int k = 0; k < 6; k++) dcov[k] = 0.f;
            }
#pragma unroll
            for (int k = 0; k < 6; k++) o_cov[k] = dcov[k];
            float dT0[3], dT1[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                dT0[k] = 2 * p[k] * dL_da + qv[k] * dL_db;
                dT1[k] = 2 * qv[k] * dL_dc + p[k] * dL_db;
            }
            const float dJ00 = vm[0] * dT0[0] + vm[4] * dT0[1] + vm[8] * dT0[2];
            const float dJ02 = vm[2] * dT0[0] + vm[6] * dT0[1] + vm[10] * dT0[2];
            const float dJ11 = vm[1] * dT1[0] + vm[5] * dT1[1] + vm[9] * dT1[2];
            const float dJ12 = vm[2] * dT1[0] + vm[6] * dT1[1] + vm[10] * dT1[2];
            const float tz = 1.f / tzv, tz2 = tz * tz, tz3 = tz2 * tz;
            const float dtx = x_grad_mul * -h_x * tz2 * dJ02;
            const float dty = y_grad_mul * -h_y * tz2 * dJ12;
            const float dtz = -h_x * tz2 * dJ00 - h_y * tz2 * dJ11 + (2 * h_x * tx) * tz3 * dJ02 + (2 * h_y * ty) * tz3 * dJ12;
            f3 dmean = {vm[0] * dtx + vm[1] * dty + vm[2] * dtz, vm[4] * dtx + vm[5] * dty + vm[6] * dtz,
                        vm[8] * dtx + vm[9] * dty + vm[10] * dtz};
            // ---- preprocessCUDA backward (backward.cu:346-396) ----
            const float hw = xf_row(proj, 3, mx, my, mz);
            const float m_w = 1.0f / (hw + 0.0000001f);
            const float mul1 = (proj[0] * mx + proj[4] * my + proj[8] * mz + proj[12]) * m_w * m_w;
            const float mul2 = (proj[1] * mx + proj[5] * my + proj[9] * mz + proj[13]) * m_w * m_w;
            dmean.x += (proj[0] * m_w - proj[3] * mul1) * dmx + (proj[1] * m_w - proj[3] * mul2) * dmy;
            dmean.y += (proj[4] * m_w - proj[7] * mul1) * dmx + (proj[5] * m_w - proj[7] * mul2) * dmy;
            dmean.z += (proj[8] * m_w - proj[11] * mul1) * dmx + (proj[9] * m_w - proj[11] * mul2) * dmy;
            if (a.shs) {
                // accumulated dL/dRGB, already masked where the forward clamped the colour
                const f3 dRGB = {sm[PB_DCOL + tid * 3], sm[PB_DCOL + tid * 3 + 1], sm[PB_DCOL + tid * 3 + 2]};
                const float *cp = a.v.campos;
                sh_backward_inplace<MULTI>(a.v.D, M, RAW ? sm + PB_SHDC + tid * 3 : row, RAW ? row : row + 3,
                                           {mx - cp[0], my - cp[1], mz - cp[2]}, dRGB, dmean,
                                           MULTI ? a.dsh_scale : 1.0f, oacc);
            }
            o_m3d[0] = dmean.x;
            o_m3d[1] = dmean.y;
            o_m3d[2] = dmean.z;
            // ---- computeCov3D backward (backward.cu:278-341) ----
            if (a.scales) {
                const float r = q.x, x = q.y, y = q.z, z = q.w;
                const float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                                       {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                                       {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
                const float s[3] = {a.v.scale_modifier * sc0, a.v.scale_modifier * sc1, a.v.scale_modifier * sc2};
                const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                                        {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                                        {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
                float dMt[3][3];
#pragma unroll
                for (int c = 0; c < 3; c++)
#pragma unroll
                    for (int w = 0; w < 3; w++)
                        dMt[w][c] = 2.0f * s[w] * R[0][w] * dS[c][0] + 2.0f * s[w] * R[1][w] * dS[c][1] +
                                    2.0f * s[w] * R[2][w] * dS[c][2];
#pragma unroll
                for (int k = 0; k < 3; k++) o_scl[k] = R[0][k] * dMt[k][0] + R[1][k] * dMt[k][1] + R[2][k] * dMt[k][2];
#pragma unroll
                for (int k = 0; k < 3; k++)
#pragma unroll
                    for (int w = 0; w < 3; w++) dMt[k][w] *= s[k];
                float4 dq;
                dq.x = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
                dq.y = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) -
                       4 * x * (dMt[2][2] + dMt[1][1]);
                dq.z = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) -
                       4 * y * (dMt[2][2] + dMt[0][0]);
                dq.w = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) -
                       4 * z * (dMt[1][1] + dMt[0][0]);
                if (RAW) {
                    // chain rule of the activations: exp -> x s;  normalize -> (g - q <q, g>) / |q_raw|
                    o_scl[0] *= sc0;
                    o_scl[1] *= sc1;
                    o_scl[2] *= sc2;
                    const float qd = q.x * dq.x + q.y * dq.y + q.z * dq.z + q.w * dq.w, inv = __fdiv_rn(1.0f, q_len);
                    dq = make_float4((dq.x - q.x * qd) * inv, (dq.y - q.y * qd) * inv, (dq.z - q.z * qd) * inv,
                                     (dq.w - q.w * qd) * inv);
                }
                o_rot[0] = dq.x;
                o_rot[1] = dq.y;
                o_rot[2] = dq.z;
                o_rot[3] = dq.w;
            } else {
                o_scl[0] = o_scl[1] = o_scl[2] = 0.f;
                o_rot[0] = o_rot[1] = o_rot[2] = o_rot[3] = 0.f;
            }
        }
    }
    // ---- write-out: the small arrays leave as 1-D bulk copies (TMA stores issued by one thread) when the
    // block is full and everything is 16-byte aligned, else through coalesced loops
    const bool out_bulk = a.out_bulk_ok && full;
    float *rec_out = a.rec11;
    if (a.stage_tab) rec_out = a.stage_tab[(int)((long long)blockIdx.x * a.nviews / gridDim.x)];
    if (out_bulk) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (out_bulk) {
        if (tid == 0) {
            bulk_s2g(a.dmeans2D + (size_t)base * 3, sm + PB_O_M2D, PB_T * 12);
            bulk_s2g(a.dcov3D + (size_t)base * 6, sm + PB_O_COV, PB_T * 24);
            if (a.rec11) {
                bulk_s2g(rec_out + (size_t)base * 11, sm + PB_O_R11, PB_T * 44);
            } else {
                bulk_s2g(a.dopacity + base, sm + PB_O_OPA, PB_T * 4);
                bulk_s2g(a.dmeans3D + (size_t)base * 3, sm + PB_O_M3D, PB_T * 12);
                bulk_s2g(a.dscales + (size_t)base * 3, sm + PB_O_SCL, PB_T * 12);
                bulk_s2g(a.drots + (size_t)base * 4, sm + PB_O_ROT, PB_T * 16);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    } else {
        pb_flush(a.dmeans2D + (size_t)base * 3, sm + PB_O_M2D, n * 3);
        pb_flush(a.dcov3D + (size_t)base * 6, sm + PB_O_COV, n * 6);
        if (a.rec11) {
            pb_flush(rec_out + (size_t)base * 11, sm + PB_O_R11, n * 11);
        } else {
            pb_flush(a.dopacity + base, sm + PB_O_OPA, n);
            pb_flush(a.dmeans3D + (size_t)base * 3, sm + PB_O_M3D, n * 3);
            pb_flush(a.dscales + (size_t)base * 3, sm + PB_O_SCL, n * 3);
            pb_flush(a.drots + (size_t)base * 4, sm + PB_O_ROT, n * 4);
        }
    }
    if (RAW && a.dsh && M > 0) {
        // the gradient blocks of a CTA are contiguous in both arrays: two bulk stores (or two plain loops)
        if (out_bulk) {
            if (tid == 0) {
                bulk_s2g(a.dsh + (size_t)base * 3, sm + PB_SHDC, PB_T * 12);
                if (M > 1) bulk_s2g(a.dsh_rest + (size_t)base * row_f, sm + PB_SH, PB_T * (uint32_t)row_f * 4);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            pb_flush(a.dsh + (size_t)base * 3, sm + PB_SHDC, n * 3);
            pb_flush(a.dsh_rest + (size_t)base * row_f, sm + PB_SH, n * row_f);
        }
    } else if (a.dsh && M > 0) {
        float *dst = a.dsh + (size_t)base * row_f;
        if (a.shs) {
            if (a.sh_vec) {
                const int row_v = row_f >> 2, total = n * row_v;
                int r = tid / row_v, c = tid - r * row_v;
                const int dr = PB_T / row_v, dc = PB_T - dr * row_v;
                for (int k = tid; k < total; k += PB_T) {
                    ((float4 *)dst)[k] = *(const float4 *)(s_sh + r * a.sh_stride + c * 4);
                    r += dr;
                    c += dc;
                    if (c >= row_v) {
                        c -= row_v;
                        r++;
                    }
                }
            } else {
                const int total = n * row_f;
                for (int k = tid; k < total; k += PB_T) {
                    const int r = k / row_f, c = k - r * row_f;
                    dst[k] = s_sh[r * a.sh_stride + c];
                }
            }
        } else {
            for (int k = tid; k < n * row_f; k += PB_T) dst[k] = 0.f;
        }
    }
    // the bulk stores read shared memory asynchronously: it must stay valid until they have
    if (out_bulk && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// View-parallel epilogue (one thread per Gaussian of a chunk), after the exchange:
//  * dL_dsh from the views' factors: dL_dsh[i] = sum_v basis(normalize(mean_i - campos_v)) (x) dRGB_v[i]
//    (the SH part of backward.cu:20-139 is an outer product per view).  The 3*M accumulators stay in
//    registers over the views and the rows leave through shared memory as coalesced stores.
//  * the all-reduced 44-byte records (means3D 3 | opacity 1 | scales 3 | rotations 4) are split into
//    the four gradient arrays autograd expects.
// Either half is optional.
// ------------------------------------------------------------------------------------------------
// Where view v's factor block [P,3] starts: a slice of one gathered buffer (base + v * stride), or -- peer mode -- an
// entry of a device-resident pointer table whose entries may point into OTHER GPUs' memory (CUDA IPC mappings over
// NVLink): the all-gather is then fused into this kernel's loads and no gathered copy exists.
__device__ __forceinline__ const float *view_block(const float *base, size_t stride, const float *const *tab, int v)
{
    return tab ? tab[v] : base + (size_t)v * stride;
}

constexpr int SF_T = 256;  // 64 Gaussians per CTA, four threads each
#ifndef SGR_SF_VB
#define SGR_SF_VB 4
#endif
constexpr int SF_VB = SGR_SF_VB;  // views whose factors a thread fetches before it uses any

// Four threads per Gaussian: thread q owns SH coefficients 4q .. 4q+3, i.e. 48 contiguous bytes of the row
// (three 16-byte stores; the four threads of a Gaussian cover its 192-byte row, a warp 1.5 KB without gaps), and
// evaluates only those four basis functions per view -- 12 accumulators instead of 48 per thread, no shared-memory
// transpose, 4x the threads in flight.  Rows with M != 16 use the generic path below (one thread per Gaussian).
// `view_stride` is the distance in floats between two views' factor blocks (>= 3P: the exchange appends each view's
// camera position to its block); campos[v] may live there too.
__global__ void __launch_bounds__(SF_T) view_grad_finalize_m16_kernel(int p0, int p1, int deg, int nviews,
                                                                       const float *__restrict__ means,
                                                                       const float *__restrict__ campos, int campos_stride,
                                                                       const float *__restrict__ dRGB, size_t view_stride,
                                                                       const float *const *__restrict__ view_tab,
                                                                       size_t campos_off, float *__restrict__ dsh,
                                                                       const float *__restrict__ rec11, float scale,
                                                                       float *__restrict__ dmeans3D,
                                                                       float *__restrict__ dopacity,
                                                                       float *__restrict__ dscales, float *__restrict__ drots)
{
    const int t = blockIdx.x * SF_T + threadIdx.x;
    const int i = p0 + (t >> 2), q = t & 3;
    if (i >= p1) return;
    // Every load of the thread is issued before its first store or use: the kernel is a latency chain per thread
    // (records -> stores, mean + factors -> rows), and what hides DRAM latency here is bytes in flight per thread.
    float rv[4] = {0.f, 0.f, 0.f, 0.f};
    if (rec11 && q < 3) {
        // the 11 reduced floats of a Gaussian, split over its four threads: means3D | opacity + scales | rotations
        const float *r = rec11 + (size_t)i * 11 + (q == 0 ? 0 : q == 1 ? 3 : 7);
        rv[0] = __ldcs(r);
        rv[1] = __ldcs(r + 1);
        rv[2] = __ldcs(r + 2);
        if (q) rv[3] = __ldcs(r + 3);
    }
    float mx = 0.f, my = 0.f, mz = 0.f;
    float gv[SF_VB][3];
    if (dsh) {
        mx = __ldg(means + 3 * (size_t)i);
        my = __ldg(means + 3 * (size_t)i + 1);
        mz = __ldg(means + 3 * (size_t)i + 2);
#pragma unroll
        for (int u = 0; u < SF_VB; u++) {
            const bool on = u < nviews;
            const float *g = view_block(dRGB, view_stride, view_tab, on ? u : 0) + (size_t)i * 3;
            gv[u][0] = on ? __ldcs(g) : 0.f;
            gv[u][1] = on ? __ldcs(g + 1) : 0.f;
            gv[u][2] = on ? __ldcs(g + 2) : 0.f;
        }
    }
    if (rec11) {
        if (q == 0) {
            __stcs(dmeans3D + 3 * (size_t)i, rv[0] * scale);
            __stcs(dmeans3D + 3 * (size_t)i + 1, rv[1] * scale);
            __stcs(dmeans3D + 3 * (size_t)i + 2, rv[2] * scale);
        } else if (q == 1) {
            __stcs(dopacity + i, rv[0] * scale);
            __stcs(dscales + 3 * (size_t)i, rv[1] * scale);
            __stcs(dscales + 3 * (size_t)i + 1, rv[2] * scale);
            __stcs(dscales + 3 * (size_t)i + 2, rv[3] * scale);
        } else if (q == 2) {
            __stcs((float4 *)(drots + 4 * (size_t)i), make_float4(rv[0] * scale, rv[1] * scale, rv[2] * scale, rv[3] * scale));
        }
    }
    if (!dsh) return;
    float acc[4][3];
#pragma unroll
    for (int k = 0; k < 4; k++) acc[k][0] = acc[k][1] = acc[k][2] = 0.f;
    // The views' factors are independent 12-byte reads 3P floats apart: a thread fetches SF_VB views' worth at a time
    // (the first batch above, with everything else), so 8 views cost 2 dependent DRAM round trips, not 8.
    for (int v0 = 0; v0 < nviews; v0 += SF_VB) {
        if (v0) {
#pragma unroll
            for (int u = 0; u < SF_VB; u++) {
                const bool on = v0 + u < nviews;
                const float *g = view_block(dRGB, view_stride, view_tab, on ? v0 + u : v0) + (size_t)i * 3;
                gv[u][0] = on ? __ldcs(g) : 0.f;
                gv[u][1] = on ? __ldcs(g + 1) : 0.f;
                gv[u][2] = on ? __ldcs(g + 2) : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < SF_VB; u++) {
            const float gr = gv[u][0] * scale, gg = gv[u][1] * scale, gb = gv[u][2] * scale;
            if (gr == 0.f && gg == 0.f && gb == 0.f) continue;  // not visible in this view (or past the last view)
            const float *cp = view_tab ? view_tab[v0 + u] + campos_off : campos + (size_t)(v0 + u) * campos_stride;
            const float ox = mx - __ldg(cp), oy = my - __ldg(cp + 1), oz = mz - __ldg(cp + 2);
            const float inv = 1.0f / sqrtf(ox * ox + oy * oy + oz * oz);
            const float x = ox * inv, y = oy * inv, z = oz * inv;
            const float xx = x * x, yy = y * y, zz = z * z;
            float w[4] = {0.f, 0.f, 0.f, 0.f};
            if (q == 0) {
                w[0] = SH_C0;
                if (deg > 0) {
                    w[1] = -SH_C1 * y;
                    w[2] = SH_C1 * z;
                    w[3] = -SH_C1 * x;
                }
            } else if (q == 1) {
                if (deg > 1) {
                    w[0] = b_SH_C2[0] * x * y;
                    w[1] = b_SH_C2[1] * y * z;
                    w[2] = b_SH_C2[2] * (2.f * zz - xx - yy);
                    w[3] = b_SH_C2[3] * x * z;
                }
            } else if (q == 2) {
                if (deg > 1) w[0] = b_SH_C2[4] * (xx - yy);
                if (deg > 2) {
                    w[1] = b_SH_C3[0] * y * (3.f * xx - yy);
                    w[2] = b_SH_C3[1] * x * y * z;
                    w[3] = b_SH_C3[2] * y * (4.f * zz - xx - yy);
                }
            } else if (deg > 2) {
                w[0] = b_SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                w[1] = b_SH_C3[4] * x * (4.f * zz - xx - yy);
                w[2] = b_SH_C3[5] * z * (xx - yy);
                w[3] = b_SH_C3[6] * x * (xx - 3.f * yy);
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                acc[k][0] = fmaf(w[k], gr, acc[k][0]);
                acc[k][1] = fmaf(w[k], gg, acc[k][1]);
                acc[k][2] = fmaf(w[k], gb, acc[k][2]);
            }
        }
    }
    // written once, 576 MB per backward at 3M Gaussians: streaming stores keep them from displacing the L2's contents
    float4 *dst = (float4 *)(dsh + (size_t)i * 48 + q * 12);
    __stcs(dst, make_float4(acc[0][0], acc[0][1], acc[0][2], acc[1][0]));
    __stcs(dst + 1, make_float4(acc[1][1], acc[1][2], acc[2][0], acc[2][1]));
    __stcs(dst + 2, make_float4(acc[2][2], acc[3][0], acc[3][1], acc[3][2]));
}

constexpr int SFG_T = 128;

__global__ void __launch_bounds__(SFG_T) view_grad_finalize_kernel(int p0, int p1, int P, int M, int deg, int nviews,
                                                                   const float *__restrict__ means,
                                                                   const float *__restrict__ campos, int campos_stride,
                                                                   const float *__restrict__ dRGB, size_t view_stride,
                                                                   const float *const *__restrict__ view_tab,
                                                                   size_t campos_off, float *__restrict__ dsh,
                                                                   const float *__restrict__ rec11, float scale,
                                                                   float *__restrict__ dmeans3D, float *__restrict__ dopacity,
                                                                   float *__restrict__ dscales, float *__restrict__ drots)
{
    extern __shared__ __align__(16) float s_rows[];  // SFG_T rows x stride floats (SH half only)
    const int tid = threadIdx.x, base = p0 + blockIdx.x * SFG_T, i = base + tid;
    const int n = min(SFG_T, p1 - base);
    if (rec11 && tid < n) {
        const float *r = rec11 + (size_t)i * 11;
        float v[11];
#pragma unroll
        for (int k = 0; k < 11; k++) v[k] = r[k] * scale;
        dmeans3D[3 * (size_t)i] = v[0];
        dmeans3D[3 * (size_t)i + 1] = v[1];
        dmeans3D[3 * (size_t)i + 2] = v[2];
        dopacity[i] = v[3];
        dscales[3 * (size_t)i] = v[4];
        dscales[3 * (size_t)i + 1] = v[5];
        dscales[3 * (size_t)i + 2] = v[6];
        drots[4 * (size_t)i] = v[7];
        drots[4 * (size_t)i + 1] = v[8];
        drots[4 * (size_t)i + 2] = v[9];
        drots[4 * (size_t)i + 3] = v[10];
    }
    if (!dsh) return;
    const int row_f = M * 3;
    const int stride = (row_f & 1) ? row_f : row_f + 1;  // odd stride: conflict-free scalar access
    float acc[16][3];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k][0] = acc[k][1] = acc[k][2] = 0.f;
    if (tid < n) {
        const float mx = means[3 * (size_t)i], my = means[3 * (size_t)i + 1], mz = means[3 * (size_t)i + 2];
        for (int v = 0; v < nviews; v++) {
            const float *g = view_block(dRGB, view_stride, view_tab, v) + (size_t)i * 3;
            const float gr = g[0] * scale, gg = g[1] * scale, gb = g[2] * scale;
            if (gr == 0.f && gg == 0.f && gb == 0.f) continue;  // not visible in this view
            const float *cp = view_tab ? view_tab[v] + campos_off : campos + (size_t)v * campos_stride;
            const float ox = mx - cp[0], oy = my - cp[1], oz = mz - cp[2];
            const float inv = 1.0f / sqrtf(ox * ox + oy * oy + oz * oz);
            const float x = ox * inv, y = oy * inv, z = oz * inv;
            float w[16];
            w[0] = SH_C0;
            if (deg > 0) {
                w[1] = -SH_C1 * y;
                w[2] = SH_C1 * z;
                w[3] = -SH_C1 * x;
                if (deg > 1) {
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    w[4] = b_SH_C2[0] * xy;
                    w[5] = b_SH_C2[1] * yz;
                    w[6] = b_SH_C2[2] * (2.f * zz - xx - yy);
                    w[7] = b_SH_C2[3] * xz;
                    w[8] = b_SH_C2[4] * (xx - yy);
                    if (deg > 2) {
                        w[9] = b_SH_C3[0] * y * (3.f * xx - yy);
                        w[10] = b_SH_C3[1] * xy * z;
                        w[11] = b_SH_C3[2] * y * (4.f * zz - xx - yy);
                        w[12] = b_SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
                        w[13] = b_SH_C3[4] * x * (4.f * zz - xx - yy);
                        w[14] = b_SH_C3[5] * z * (xx - yy);
                        w[15] = b_SH_C3[6] * x * (xx - 3.f * yy);
                    }
                }
            }
            const int used = (deg + 1) * (deg + 1);
#pragma unroll
            for (int k = 0; k < 16; k++)
                if (k < used) {
                    acc[k][0] = fmaf(w[k], gr, acc[k][0]);
                    acc[k][1] = fmaf(w[k], gg, acc[k][1]);
                    acc[k][2] = fmaf(w[k], gb, acc[k][2]);
                }
        }
        float *row = s_rows + tid * stride;
#pragma unroll
        for (int k = 0; k < 16; k++)
            if (k < M) {
                row[k * 3] = acc[k][0];
                row[k * 3 + 1] = acc[k][1];
                row[k * 3 + 2] = acc[k][2];
            }
        for (int k = 16; k < M; k++) row[k * 3] = row[k * 3 + 1] = row[k * 3 + 2] = 0.f;
    }
    __syncthreads();
    float *dst = dsh + (size_t)base * row_f;
    for (int k = tid; k < n * row_f; k += SFG_T) {
        const int r = k / row_f, c = k - r * row_f;
        dst[k] = s_rows[r * stride + c];
    }
}

static int finalize_impl(int32_t P, int32_t p0, int32_t p1, int32_t M, int32_t sh_degree, int32_t num_views,
                         const float *means3D, const float *campos, const float *dRGB, int64_t view_stride,
                         int32_t campos_stride, const float *const *view_tab, float *dL_dsh, const float *reduced_records,
                         float scale, float *dL_dmeans3D, float *dL_dopacity, float *dL_dscales, float *dL_drotations,
                         void *stream, const char *who);

// CTA range [b0, b1) of chunk c of the per-Gaussian pass.  taper == 0: equal chunks.  taper != 0: every chunk half the size
// of the one before it (weights 2^(n-1-c)), so that what a caller does with a finished chunk while the next one is
// computed (the peer exchange: reduce + split) is matched by that next chunk's shorter run, and little is left after the
// last, smallest chunk.
static void chunk_blocks(int blocks, int nchunks, int c, int taper, int *b0, int *b1)
{
    if (!taper || nchunks > 30) {
        *b0 = (int)((int64_t)blocks * c / nchunks);
        *b1 = (int)((int64_t)blocks * (c + 1) / nchunks);
        return;
    }
    const int64_t full = ((int64_t)1 << nchunks) - 1;
    auto edge = [&](int k) { return (int)((int64_t)blocks * ((((int64_t)1 << nchunks) - ((int64_t)1 << (nchunks - k)))) / full); };
    *b0 = edge(c);
    *b1 = c + 1 == nchunks ? blocks : edge(c + 1);
}

int launch_backward(const SgrView *view, const SgrGaussians *g, const int32_t *radii, const void *geom_buffer,
                    const void *binning_buffer, const void *image_buffer, int64_t num_rendered,
                    const float *dL_dout_color, float *dL_dmeans2D, float *dL_dcolors, float *dL_dopacity,
                    float *dL_dmeans3D, float *dL_dcov3D, float *dL_dsh, float *dL_dscales, float *dL_drotations,
                    void *grad_scratch, cudaStream_t st, const SgrBackwardPlan *plan)
{
    const int P = g->P, W = view->image_width, H = view->image_height;
    const int gx = (W + SGR_TILE - 1) / SGR_TILE, gy = (H + SGR_TILE - 1) / SGR_TILE;
    GeomState geom = GeomState::carve((void *)geom_buffer, P);
    ImageState img = ImageState::carve((void *)image_buffer, W, H);
    BinState bin = BinState::carve((void *)binning_buffer, (size_t)num_rendered);
    float *gacc = (float *)align_up((size_t)grad_scratch);
    const SgrStageHook hook = plan ? plan->hook : nullptr;
    void *hook_ctx = plan ? plan->hook_ctx : nullptr;
    float *rec11 = plan ? plan->reduce_records : nullptr;
    {
        // opt in to > 48 KB of dynamic shared memory once per device (the call is not free); calls
        // may come from several threads (autograd engine threads of different devices)
        static std::mutex mu;
        static bool attr_set[64] = {};
        int dev = 0;
        SGR_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lock(mu);
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            SGR_CUDA(cudaFuncSetAttribute(blend_backward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)BWD_SMEM_BYTES));
            SGR_CUDA(cudaFuncSetAttribute(blend_backward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)BWD_SMEM_BYTES));
            SGR_CUDA(cudaFuncSetAttribute(preprocess_backward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          200 * 1024));
            SGR_CUDA(cudaFuncSetAttribute(preprocess_backward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          200 * 1024));
            SGR_CUDA(cudaFuncSetAttribute(preprocess_backward_kernel<false, true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
    }
    // the blend accumulators: 32 B per Gaussian of scratch + the dL_dcolors output itself
    SGR_CUDA(cudaMemsetAsync(gacc, 0, (size_t)P * 32, st));
    SGR_CUDA(cudaMemsetAsync(dL_dcolors, 0, (size_t)P * 12, st));
    if (num_rendered > 0) {
        SGR_LAUNCH(K_BLEND_BWD, st,
                   if (ids_packed(P))
                       blend_backward_kernel<true><<<gx * gy, 256, BWD_SMEM_BYTES, st>>>(
                           img.tile_order, img.tile_start, bin.plist, geom.rec, W, H, gx, view->bg, img.final_T,
                           img.n_contrib, dL_dout_color, gacc, dL_dcolors);
                   else
                       blend_backward_kernel<false><<<gx * gy, 256, BWD_SMEM_BYTES, st>>>(
                           img.tile_order, img.tile_start, bin.plist, geom.rec, W, H, gx, view->bg, img.final_T,
                           img.n_contrib, dL_dout_color, gacc, dL_dcolors));
    }
    // dL_dcolors is final here (with SH colours: the clamp-masked dL/dRGB, i.e. this view's SH factor)
    if (hook) hook(hook_ctx, SGR_STAGE_BLEND_DONE);
    const bool peer = plan && plan->peer_flag_tab;
    // this backward's own signals: on `st`, or on the plan's signal stream behind an event recorded on `st` (so that the
    // next kernel on `st` does not queue behind the signal's launch)
    auto signal = [&](int slot, int ev_index, bool on_main = false) -> int {
        cudaStream_t sig = st;
        if (plan->peer_signal_stream && !on_main) {
            static std::mutex mu;
            static cudaEvent_t evs[64][20] = {};
            int dev = 0;
            SGR_CUDA(cudaGetDevice(&dev));
            if (dev < 0 || dev >= 64 || ev_index < 0 || ev_index >= 20) {
                set_error("peer_signal_stream: device / chunk index out of range");
                return SGR_EINVAL;
            }
            std::lock_guard<std::mutex> lock(mu);
            if (!evs[dev][ev_index]) SGR_CUDA(cudaEventCreateWithFlags(&evs[dev][ev_index], cudaEventDisableTiming));
            sig = (cudaStream_t)plan->peer_signal_stream;
            SGR_CUDA(cudaEventRecord(evs[dev][ev_index], st));
            SGR_CUDA(cudaStreamWaitEvent(sig, evs[dev][ev_index], 0));
        }
        return sgr_peer_signal(plan->peer_flag_tab, plan->peer_nranks, slot, plan->peer_rank, plan->peer_seq, sig);
    };
    if (peer) {
        // BLEND goes out on `st` itself: the wait right behind it spins, and a spinning kernel must never be dispatched
        // ahead of a signal of its own rank that the peers need in order to answer
        int rc = signal(plan->peer_slot_blend, 0, true);
        if (rc) return rc;
        if (plan->peer_view_blocks && dL_dsh && g->shs) {
            // the per-Gaussian pass sums every view's SH gradient into dL_dsh: the peers' factor blocks must be final
            if (g->activations != 0 || !plan->peer_flags) {
                set_error("peer_view_blocks: needs peer_flags and activated (not raw) parameters");
                return SGR_EINVAL;
            }
            rc = sgr_peer_wait(plan->peer_flags, plan->peer_nranks, plan->peer_slot_blend, 1, plan->peer_seq,
                               plan->peer_timeout_s, st);
            if (rc) return rc;
        }
    }
    PreBwdArgs a;
    a.means = g->means3D;
    a.scales = g->scales;
    a.rots = g->rotations;
    a.shs = g->shs;
    a.cov_pre = g->cov3D_precomp;
    a.v.viewmatrix = view->viewmatrix;
    a.v.projmatrix = view->projmatrix;
    a.v.campos = view->campos;
    a.v.bg = view->bg;
    a.v.W = W;
    a.v.H = H;
    a.v.gx = gx;
    a.v.gy = gy;
    a.v.tanfovx = view->tanfovx;
    a.v.tanfovy = view->tanfovy;
    a.v.focal_y = H / (2.0f * view->tanfovy);
    a.v.focal_x = W / (2.0f * view->tanfovx);
    a.v.scale_modifier = view->scale_modifier;
    a.v.D = view->sh_degree;
    a.v.M = g->M;
    a.v.prefiltered = view->prefiltered;
    a.radii = radii;
    a.gacc = gacc;
    a.dcol = dL_dcolors;
    a.dmeans2D = dL_dmeans2D;
    a.dopacity = dL_dopacity;
    a.dmeans3D = dL_dmeans3D;
    a.dcov3D = dL_dcov3D;
    a.dsh = dL_dsh;
    a.dscales = dL_dscales;
    a.drots = dL_drotations;
    a.rec11 = rec11;
    a.view_tab = nullptr;
    a.stage_tab = nullptr;
    a.pf_off = 0;
    a.nviews = 1;
    a.my_view = 0;
    a.campos_off = (size_t)3 * P;
    a.dsh_scale = 1.0f;
    const bool raw = g->activations != 0;
    if (peer && plan->peer_view_blocks && dL_dsh && g->shs) {
        a.view_tab = (const float *const *)plan->peer_view_blocks;
        a.nviews = plan->peer_nranks;
        a.my_view = plan->peer_rank;
        a.dsh_scale = plan->peer_dsh_scale;
        a.stage_tab = (float *const *)plan->peer_record_stages;
        if (a.stage_tab && (!rec11 || PB_T != 64)) {
            set_error("peer_record_stages needs reduce_records (the local staging array) and 64-thread CTAs");
            return SGR_EINVAL;
        }
    }
    a.sh_rest = g->sh_rest;
    a.rec = geom.rec;
    a.dsh_rest = plan ? plan->dL_dsh_rest : nullptr;
    if (raw && g->shs && dL_dsh && g->M > 1 && !a.dsh_rest) {
        set_error("raw-parameter mode needs SgrBackwardPlan.dL_dsh_rest");
        return SGR_EINVAL;
    }
    auto al16 = [](const void *p) { return ((uintptr_t)p & 15u) == 0; };
    a.in_bulk_ok = al16(gacc) && al16(dL_dcolors) && al16(g->means3D) && (!g->scales || al16(g->scales)) &&
                   (!g->rotations || al16(g->rotations)) && (!g->cov3D_precomp || al16(g->cov3D_precomp)) &&
                   (!raw || !g->shs || (al16(g->shs) && (g->M == 1 || al16(g->sh_rest))));
    a.out_bulk_ok = al16(dL_dmeans2D) && al16(dL_dcov3D) &&
                    (rec11 ? al16(rec11) : (al16(dL_dopacity) && al16(dL_dmeans3D) && al16(dL_dscales) && al16(dL_drotations)));
    if (raw && g->shs && dL_dsh) a.out_bulk_ok = a.out_bulk_ok && al16(dL_dsh) && (g->M == 1 || al16(a.dsh_rest));
#ifdef SGR_PB_NO_BULK_STORE
    a.out_bulk_ok = 0;
#endif
    a.sh_stride = 0;
    a.sh_vec = 0;
    if (g->shs && !raw) {
        const int row_f = g->M * 3;
        if ((row_f % 4) == 0 && al16(g->shs) && al16(dL_dsh)) {
            int s4 = row_f / 4;
            if ((s4 & 1) == 0) s4 += 1;
            a.sh_stride = s4 * 4;
            a.sh_vec = 1;
        } else {
            a.sh_stride = (row_f & 1) ? row_f : row_f + 1;
        }
    }
    a.pf_off = PB_SH + PB_T * (raw ? (g->M > 0 ? (g->M - 1) * 3 : 0) : a.sh_stride);
    const int pulled = a.view_tab ? (a.nviews - 1 < PB_VB ? a.nviews - 1 : PB_VB) : 0;
    const size_t dyn = (size_t)(a.pf_off + pulled * PB_T * 3) * sizeof(float) + 16;
    // the per-Gaussian pass, in `num_chunks` Gaussian ranges (multiples of the CTA size) so that a caller
    // can start reducing a finished range while the next one is computed
    int nchunks = plan && plan->num_chunks > 1 ? plan->num_chunks : 1;
    const int blocks = (P + PB_T - 1) / PB_T;
    if (nchunks > blocks) nchunks = blocks;
    int split_c = -1, split_p0 = 0, split_p1 = 0;
    auto side_event = [&](int index, cudaEvent_t *out) -> int {
        static std::mutex mu;
        static cudaEvent_t evs[64][20] = {};
        int dev = 0;
        SGR_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || index < 0 || index >= 20) {
            set_error("peer_side_stream: device / chunk index out of range");
            return SGR_EINVAL;
        }
        std::lock_guard<std::mutex> lock(mu);
        if (!evs[dev][index]) SGR_CUDA(cudaEventCreateWithFlags(&evs[dev][index], cudaEventDisableTiming));
        *out = evs[dev][index];
        return SGR_OK;
    };
    auto split_chunk = [&](cudaStream_t B, int c, int p0, int p1) -> int {
        int rc = sgr_peer_wait(plan->peer_flags, plan->peer_nranks, plan->peer_slot_reduced0 + c, 1, plan->peer_seq,
                               plan->peer_timeout_s, B);
        if (rc) return rc;
        return finalize_impl(P, p0, p1, g->M, view->sh_degree, plan->peer_nranks, g->means3D, nullptr, nullptr, 0, 0, nullptr,
                             nullptr, plan->peer_sums, plan->peer_dsh_scale, dL_dmeans3D, dL_dopacity, dL_dscales,
                             dL_drotations, B, "peer exchange: bad arguments to the record split");
    };
    if (peer && plan->peer_side_stream &&
        (!plan->peer_rec_tab || !plan->peer_sum_tab || !plan->peer_sums || !plan->peer_flags || !rec11 || !dL_dmeans3D ||
         !dL_dopacity || !dL_dscales || !dL_drotations || nchunks > 16)) {
        set_error("peer_side_stream needs peer_rec_tab / peer_sum_tab / peer_sums / peer_flags, reduce_records, the four "
                  "record-bound gradient arrays and at most 16 chunks");
        return SGR_EINVAL;
    }
    for (int c = 0; c < nchunks; c++) {
        int b0, b1;
        chunk_blocks(blocks, nchunks, c, plan ? plan->chunk_taper : 0, &b0, &b1);
        a.p0 = b0 * PB_T;
        a.p1 = min(P, b1 * PB_T);
        if (b1 > b0)
            SGR_LAUNCH(K_PRE_BWD, st,
                       if (raw) preprocess_backward_kernel<true><<<b1 - b0, PB_T, dyn, st>>>(a);
                       else if (a.view_tab) preprocess_backward_kernel<false, true><<<b1 - b0, PB_T, dyn, st>>>(a);
                       else preprocess_backward_kernel<false><<<b1 - b0, PB_T, dyn, st>>>(a));
        if (hook) hook(hook_ctx, SGR_STAGE_CHUNK_DONE + c);
        if (peer) {
            int rc = signal(plan->peer_slot_chunk0 + c, 1 + c);
            if (rc) return rc;
            if (plan->peer_side_stream) {
                // The records' exchange of this chunk, on the side stream.  Its kernels spin on flags: they may start
                // only once THIS rank's own signal for the chunk is out (an event behind the signal kernel) -- a
                // spinning kernel must never sit in front of a signal some rank is waiting for.
                cudaStream_t B = (cudaStream_t)plan->peer_side_stream;
                cudaStream_t sig = plan->peer_signal_stream ? (cudaStream_t)plan->peer_signal_stream : st;
                cudaEvent_t e2 = nullptr;
                rc = side_event(c, &e2);
                if (rc) return rc;
                SGR_CUDA(cudaEventRecord(e2, sig));
                SGR_CUDA(cudaStreamWaitEvent(B, e2, 0));
                unsigned int *counters = (unsigned int *)plan->peer_flags + (size_t)(SGR_PEER_MAX_SLOTS - 1) * SGR_PEER_MAX_RANKS;
                for (int r = 0; r < plan->peer_nranks; r++) {
                    const bool me = r == plan->peer_rank;
                    if (!me && !((plan->peer_emulate_ranks >> r) & 1)) continue;
                    // one kernel: wait(CHUNK c, every rank) -> sum the owned slice -> its last CTA signals REDUCED c
                    rc = sgr_peer_reduce_records_synced(plan->peer_rec_tab, plan->peer_sum_tab, plan->peer_nranks, r, a.p0, a.p1,
                                                        plan->peer_flags, plan->peer_slot_chunk0 + c,
                                                        me ? plan->peer_flag_tab : nullptr, plan->peer_slot_reduced0 + c,
                                                        plan->peer_seq, me ? (void *)(counters + c) : nullptr,
                                                        plan->peer_timeout_s, B);
                    if (rc) return rc;
                }
                // the split of the chunk before runs behind this chunk's reduce: the other owners' slices of it have
                // landed by then, and only the last (smallest) chunk's reduce + split is exposed
                if (split_c >= 0) {
                    rc = split_chunk(B, split_c, split_p0, split_p1);
                    if (rc) return rc;
                }
                split_c = c;
                split_p0 = a.p0;
                split_p1 = a.p1;
            }
        }
    }
    if (peer && plan->peer_side_stream && split_c >= 0) {
        const int rc = split_chunk((cudaStream_t)plan->peer_side_stream, split_c, split_p0, split_p1);
        if (rc) return rc;
    }
    SGR_CUDA(cudaGetLastError());
    if (view->debug) SGR_CUDA(cudaStreamSynchronize(st));
    return SGR_OK;
}

}  // namespace sgr

extern "C" int sgr_backward_chunk_range_tapered(int32_t P, int32_t num_chunks, int32_t chunk, int32_t taper, int32_t *p0,
                                                int32_t *p1)
{
    using namespace sgr;
    if (P < 0 || num_chunks < 1 || chunk < 0 || chunk >= num_chunks || !p0 || !p1) {
        set_error("bad arguments to sgr_backward_chunk_range");
        return SGR_EINVAL;
    }
    const int blocks = (P + PB_T - 1) / PB_T;
    const int nchunks = num_chunks > blocks ? (blocks > 0 ? blocks : 1) : num_chunks;
    if (chunk >= nchunks) {
        *p0 = *p1 = P;
        return SGR_OK;
    }
    int b0, b1;
    chunk_blocks(blocks, nchunks, chunk, taper, &b0, &b1);
    *p0 = b0 * PB_T;
    *p1 = b1 * PB_T;
    if (*p1 > P) *p1 = P;
    if (*p0 > P) *p0 = P;
    return SGR_OK;
}

extern "C" int sgr_backward_chunk_range(int32_t P, int32_t num_chunks, int32_t chunk, int32_t *p0, int32_t *p1)
{
    return sgr_backward_chunk_range_tapered(P, num_chunks, chunk, 0, p0, p1);
}

namespace sgr {
static int finalize_impl(int32_t P, int32_t p0, int32_t p1, int32_t M, int32_t sh_degree, int32_t num_views,
                         const float *means3D, const float *campos, const float *dRGB, int64_t view_stride,
                         int32_t campos_stride, const float *const *view_tab, float *dL_dsh, const float *reduced_records,
                         float scale, float *dL_dmeans3D, float *dL_dopacity, float *dL_dscales, float *dL_drotations,
                         void *stream, const char *who)
{
    const bool sh = dL_dsh != nullptr, rec = reduced_records != nullptr;
    const bool src_ok = view_tab ? true : (campos && dRGB && view_stride >= (int64_t)3 * P && campos_stride >= 3);
    if (P < 0 || p0 < 0 || p1 < p0 || p1 > P || (!sh && !rec) ||
        (sh && (M <= 0 || sh_degree < 0 || sh_degree > 3 || (sh_degree + 1) * (sh_degree + 1) > M || num_views <= 0 ||
                !means3D || !src_ok)) ||
        (rec && (!dL_dmeans3D || !dL_dopacity || !dL_dscales || !dL_drotations))) {
        set_error(who);
        return SGR_EINVAL;
    }
    if (p1 == p0) return SGR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t campos_off = (size_t)3 * P;  // peer mode: a view's camera position follows its [P,3] factors
    const bool al16 = (((uintptr_t)dL_dsh | (uintptr_t)dL_drotations) & 15u) == 0;
    if ((!sh || M == 16) && al16) {
        const int threads = (p1 - p0) * 4;
        SGR_LAUNCH(K_FINALIZE, st,
                   view_grad_finalize_m16_kernel<<<(threads + SF_T - 1) / SF_T, SF_T, 0, st>>>(
                       p0, p1, sh_degree, num_views, means3D, campos, campos_stride, dRGB, (size_t)view_stride, view_tab,
                       campos_off, dL_dsh, reduced_records, scale, dL_dmeans3D, dL_dopacity, dL_dscales, dL_drotations));
    } else {
        const int row_f = M * 3, stride = (row_f & 1) ? row_f : row_f + 1;
        const size_t dyn = sh ? (size_t)SFG_T * stride * sizeof(float) : 0;
        SGR_LAUNCH(K_FINALIZE, st,
                   view_grad_finalize_kernel<<<(p1 - p0 + SFG_T - 1) / SFG_T, SFG_T, dyn, st>>>(
                       p0, p1, P, M, sh_degree, num_views, means3D, campos, campos_stride, dRGB, (size_t)view_stride,
                       view_tab, campos_off, dL_dsh, reduced_records, scale, dL_dmeans3D, dL_dopacity, dL_dscales,
                       dL_drotations));
    }
    SGR_CUDA(cudaGetLastError());
    return SGR_OK;
}
}  // namespace sgr

extern "C" int sgr_view_grad_finalize(int32_t P, int32_t p0, int32_t p1, int32_t M, int32_t sh_degree, int32_t num_views,
                                      const float *means3D, const float *campos, const float *dRGB, int64_t view_stride,
                                      int32_t campos_stride, float *dL_dsh, const float *reduced_records, float scale,
                                      float *dL_dmeans3D, float *dL_dopacity, float *dL_dscales, float *dL_drotations,
                                      void *stream)
{
    return sgr::finalize_impl(P, p0, p1, M, sh_degree, num_views, means3D, campos, dRGB, view_stride, campos_stride, nullptr,
                              dL_dsh, reduced_records, scale, dL_dmeans3D, dL_dopacity, dL_dscales, dL_drotations, stream,
                              "bad arguments to sgr_view_grad_finalize");
}

extern "C" int sgr_view_grad_finalize_peers(int32_t P, int32_t p0, int32_t p1, int32_t M, int32_t sh_degree,
                                            int32_t num_views, const float *means3D, const float *const *view_blocks,
                                            float *dL_dsh, const float *reduced_records, float scale, float *dL_dmeans3D,
                                            float *dL_dopacity, float *dL_dscales, float *dL_drotations, void *stream)
{
    if (dL_dsh && !view_blocks) {
        sgr::set_error("bad arguments to sgr_view_grad_finalize_peers");
        return SGR_EINVAL;
    }
    return sgr::finalize_impl(P, p0, p1, M, sh_degree, num_views, means3D, nullptr, nullptr, 0, 0, view_blocks, dL_dsh,
                              reduced_records, scale, dL_dmeans3D, dL_dopacity, dL_dscales, dL_drotations, stream,
                              "bad arguments to sgr_view_grad_finalize_peers");
}

extern "C" int sgr_sh_grad_from_factors(int32_t P, int32_t M, int32_t sh_degree, int32_t num_views, const float *means3D,
                                        const float *campos, const float *dRGB, float *dL_dsh, void *stream)
{
    if (!dL_dsh && P > 0) {
        sgr::set_error("bad arguments to sgr_sh_grad_from_factors");
        return SGR_EINVAL;
    }
    return sgr_view_grad_finalize(P, 0, P, M, sh_degree, num_views, means3D, campos, dRGB, (int64_t)3 * P, 3, dL_dsh, nullptr,
                                  1.0f, nullptr, nullptr, nullptr, nullptr, stream);
}
