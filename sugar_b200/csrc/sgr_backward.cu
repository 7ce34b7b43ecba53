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
// contended.  Here every warp reduces its 32 pixels with a 14-shuffle multi-value butterfly,
// parks the 9 partial sums in its private shared-memory row, and after each batch one thread per
// Gaussian folds the (at most 8) warp rows and issues two 16-byte vector reductions
// (red.global.add.v4.f32 -> SASS REDG.E.ADD.F32x4) plus one scalar: 3 L2 operations per
// (tile, Gaussian) instead of 9 per (pixel, Gaussian).
#include "sgr_internal.cuh"

namespace sgr {

constexpr int BWD_B = 128;  // Gaussians per shared-memory batch
constexpr int BWD_NW = 8;   // warps per CTA (16x16 pixels)

// accumulator record, 12 floats per Gaussian:
//   [0] dmean2D.x [1] dmean2D.y [2] dconic.x [3] dconic.y | [4] dconic.w [5] dopacity [6] dR [7] dG | [8] dB
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Sum 8 values across the warp with 9 shuffles (recursive halving), result k lands in lane 4k.
__device__ __forceinline__ float warp_reduce8(const float (&v)[8], unsigned lane)
{
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float w[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float send = b4 ? v[k] : v[k + 4];
        const float keep = b4 ? v[k + 4] : v[k];
        w[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    float u[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const float send = b3 ? w[k] : w[k + 2];
        const float keep = b3 ? w[k + 2] : w[k];
        u[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    float t;
    {
        const float send = b2 ? u[0] : u[1];
        const float keep = b2 ? u[1] : u[0];
        t = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    return t;  // lane L holds value index (b4?4:0)+(b3?2:0)+(b2?1:0)
}

__global__ void __launch_bounds__(256) blend_backward_kernel(const uint32_t *__restrict__ tile_start,
                                                             const uint32_t *__restrict__ plist,
                                                             const float4 *__restrict__ rec, int W, int H, int gx,
                                                             const float *__restrict__ bg,
                                                             const float *__restrict__ final_Ts,
                                                             const uint32_t *__restrict__ n_contrib,
                                                             const float *__restrict__ dL_dpixels, float *__restrict__ gacc)
{
    __shared__ float4 s_a[BWD_B];
    __shared__ float4 s_b[BWD_B];
    __shared__ float2 s_c[BWD_B];
    __shared__ uint32_t s_id[BWD_B];
    __shared__ float s_acc[BWD_NW][BWD_B][9];
    __shared__ uint32_t s_mask[BWD_NW][BWD_B / 32];

    const int tile = blockIdx.y * gx + blockIdx.x;
    const int tid = threadIdx.y * SGR_TILE + threadIdx.x;
    const unsigned lane = tid & 31, wid = tid >> 5;
    const uint32_t pxi = blockIdx.x * SGR_TILE + threadIdx.x, pyi = blockIdx.y * SGR_TILE + threadIdx.y;
    const bool inside = pxi < (uint32_t)W && pyi < (uint32_t)H;
    const float pxf = (float)pxi, pyf = (float)pyi;
    const uint32_t lo = tile_start[tile], hi = tile_start[tile + 1];
    const int n = (int)(hi - lo);
    if (n == 0) return;

    const size_t pix = (size_t)pyi * W + pxi, plane = (size_t)H * W;
    const float T_final = inside ? final_Ts[pix] : 0.0f;
    float T = T_final;
    const uint32_t last_contributor = inside ? n_contrib[pix] : 0u;
    float dp0 = 0.f, dp1 = 0.f, dp2 = 0.f;
    if (inside) {
        dp0 = dL_dpixels[pix];
        dp1 = dL_dpixels[plane + pix];
        dp2 = dL_dpixels[2 * plane + pix];
    }
    const float bg_dot = fmaf(bg[2], dp2, fmaf(bg[1], dp1, bg[0] * dp0));
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_alpha = 0.f;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    uint32_t contributor = (uint32_t)n;

    for (int b0 = 0; b0 < n; b0 += BWD_B) {
        __syncthreads();
        if (tid < BWD_B && b0 + tid < n) {
            const uint32_t id = plist[hi - 1 - (uint32_t)(b0 + tid)];
            const float4 *r = rec + (size_t)id * 3;
            const float4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2);
            s_id[tid] = id;
            s_a[tid] = r0;
            s_b[tid] = r1;
            s_c[tid] = make_float2(r2.x, r2.y);
        }
        __syncthreads();
        const int m = min(BWD_B, n - b0);
#pragma unroll 1
        for (int c = 0; c * 32 < m; c++) {
            uint32_t touched = 0;
            const int jend = min(32, m - c * 32);
#pragma unroll 1
            for (int jj = 0; jj < jend; jj++) {
                const int j = c * 32 + jj;
                contributor--;
                const float4 A = s_a[j];
                const float4 B = s_b[j];
                const float dx = __fsub_rn(A.x, pxf), dy = __fsub_rn(A.y, pyf);
                const float power = splat_power(dx, dy, A.z, A.w, B.x);
                bool valid = inside && contributor < last_contributor && !(power > 0.0f) && !(power < B.y);
                if (!__any_sync(0xffffffffu, valid)) continue;
                float v[8], v8 = 0.f;
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = 0.f;
                if (valid) {
                    const float G = expf(power);
                    const float alpha = fminf(0.99f, __fmul_rn(B.z, G));
                    valid = !(alpha < 1.0f / 255.0f);
                    if (valid) {
                        const float2 Cc = s_c[j];
                        const float one_m = 1.0f - alpha;
                        const float inv = __frcp_rn(one_m);
                        T = T * inv;
                        const float dchannel = alpha * T;
                        const float la = last_alpha, om_la = 1.0f - last_alpha;
                        acc0 = fmaf(la, lc0, om_la * acc0);
                        acc1 = fmaf(la, lc1, om_la * acc1);
                        acc2 = fmaf(la, lc2, om_la * acc2);
                        lc0 = B.w;
                        lc1 = Cc.x;
                        lc2 = Cc.y;
                        float dL_dalpha = (lc0 - acc0) * dp0;
                        dL_dalpha = fmaf(lc1 - acc1, dp1, dL_dalpha);
                        dL_dalpha = fmaf(lc2 - acc2, dp2, dL_dalpha);
                        dL_dalpha *= T;
                        last_alpha = alpha;
                        dL_dalpha = fmaf(-T_final * inv, bg_dot, dL_dalpha);
                        const float dL_dG = B.z * dL_dalpha;
                        const float gdx = G * dx, gdy = G * dy;
                        const float dG_ddelx = -gdx * A.z - gdy * A.w;
                        const float dG_ddely = -gdy * B.x - gdx * A.w;
                        v[0] = dL_dG * dG_ddelx * ddelx_dx;
                        v[1] = dL_dG * dG_ddely * ddely_dy;
                        v[2] = -0.5f * gdx * dx * dL_dG;
                        v[3] = -0.5f * gdx * dy * dL_dG;
                        v[4] = -0.5f * gdy * dy * dL_dG;
                        v[5] = G * dL_dalpha;
                        v[6] = dchannel * dp0;
                        v[7] = dchannel * dp1;
                        v8 = dchannel * dp2;
                    }
                }
                if (!__any_sync(0xffffffffu, valid)) continue;
                const float r8 = warp_reduce8(v, lane);
                v8 = warp_sum(v8);
                if ((lane & 3u) == 0) s_acc[wid][j][lane >> 2] = r8;
                if (lane == 1) s_acc[wid][j][8] = v8;
                touched |= 1u << jj;
            }
            if (lane == 0) s_mask[wid][c] = touched;
        }
        __syncthreads();
        if (tid < m) {
            float s[9];
#pragma unroll
            for (int k = 0; k < 9; k++) s[k] = 0.f;
            bool any = false;
#pragma unroll
            for (int w = 0; w < BWD_NW; w++) {
                if ((s_mask[w][tid >> 5] >> (tid & 31)) & 1u) {
                    any = true;
#pragma unroll
                    for (int k = 0; k < 9; k++) s[k] += s_acc[w][tid][k];
                }
            }
            if (any) {
                float *g = gacc + (size_t)s_id[tid] * 12;
                red_add_v4(g, s[0], s[1], s[2], s[3]);
                red_add_v4(g + 4, s[4], s[5], s[6], s[7]);
                atomicAdd(g + 8, s[8]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// per-Gaussian backward (cov2D inverse, projection, SH, scale/rotation), one fused pass
// ------------------------------------------------------------------------------------------------
struct PreBwdArgs {
    int P;
    const float *means, *scales, *rots, *shs, *cov_pre;
    ViewConsts v;
    const int32_t *radii;
    const uint32_t *aux;  // clamp bits
    const float *gacc;
    float *dmeans2D, *dcolors, *dopacity, *dmeans3D, *dcov3D, *dsh, *dscales, *drots;
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
__device__ __forceinline__ f3 operator*(float s, f3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// SH backward (backward.cu:20-139).  sh/dsh rows are [M][3] floats in global memory.
__device__ __forceinline__ void sh_backward(int deg, const float *__restrict__ sh, float *__restrict__ dsh, int M,
                                            f3 dir_orig, f3 dRGB, f3 &dmean)
{
    const float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
    const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
    auto S = [&](int k) -> f3 { return {sh[k * 3], sh[k * 3 + 1], sh[k * 3 + 2]}; };
    auto D = [&](int k, float w) {
        dsh[k * 3] = w * dRGB.x;
        dsh[k * 3 + 1] = w * dRGB.y;
        dsh[k * 3 + 2] = w * dRGB.z;
    };
    f3 dx = {0, 0, 0}, dy = {0, 0, 0}, dz = {0, 0, 0};
    D(0, SH_C0);
    if (deg > 0) {
        D(1, -SH_C1 * y);
        D(2, SH_C1 * z);
        D(3, -SH_C1 * x);
        dx = -SH_C1 * S(3);
        dy = -SH_C1 * S(1);
        dz = SH_C1 * S(2);
        if (deg > 1) {
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            D(4, b_SH_C2[0] * xy);
            D(5, b_SH_C2[1] * yz);
            D(6, b_SH_C2[2] * (2.f * zz - xx - yy));
            D(7, b_SH_C2[3] * xz);
            D(8, b_SH_C2[4] * (xx - yy));
            const f3 s4 = S(4), s5 = S(5), s6 = S(6), s7 = S(7), s8 = S(8);
            dx = dx + (b_SH_C2[0] * y) * s4 + (b_SH_C2[2] * 2.f * -x) * s6 + (b_SH_C2[3] * z) * s7 + (b_SH_C2[4] * 2.f * x) * s8;
            dy = dy + (b_SH_C2[0] * x) * s4 + (b_SH_C2[1] * z) * s5 + (b_SH_C2[2] * 2.f * -y) * s6 + (b_SH_C2[4] * 2.f * -y) * s8;
            dz = dz + (b_SH_C2[1] * y) * s5 + (b_SH_C2[2] * 2.f * 2.f * z) * s6 + (b_SH_C2[3] * x) * s7;
            if (deg > 2) {
                D(9, b_SH_C3[0] * y * (3.f * xx - yy));
                D(10, b_SH_C3[1] * xy * z);
                D(11, b_SH_C3[2] * y * (4.f * zz - xx - yy));
                D(12, b_SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
                D(13, b_SH_C3[4] * x * (4.f * zz - xx - yy));
                D(14, b_SH_C3[5] * z * (xx - yy));
                D(15, b_SH_C3[6] * x * (xx - 3.f * yy));
                const f3 s9 = S(9), s10 = S(10), s11 = S(11), s12 = S(12), s13 = S(13), s14 = S(14), s15 = S(15);
                dx = dx + (b_SH_C3[0] * 3.f * 2.f * xy) * s9 + (b_SH_C3[1] * yz) * s10 + (b_SH_C3[2] * -2.f * xy) * s11 +
                     (b_SH_C3[3] * -3.f * 2.f * xz) * s12 + (b_SH_C3[4] * (-3.f * xx + 4.f * zz - yy)) * s13 +
                     (b_SH_C3[5] * 2.f * xz) * s14 + (b_SH_C3[6] * 3.f * (xx - yy)) * s15;
                dy = dy + (b_SH_C3[0] * 3.f * (xx - yy)) * s9 + (b_SH_C3[1] * xz) * s10 +
                     (b_SH_C3[2] * (-3.f * yy + 4.f * zz - xx)) * s11 + (b_SH_C3[3] * -3.f * 2.f * yz) * s12 +
                     (b_SH_C3[4] * -2.f * xy) * s13 + (b_SH_C3[5] * -2.f * yz) * s14 + (b_SH_C3[6] * -3.f * 2.f * xy) * s15;
                dz = dz + (b_SH_C3[1] * xy) * s10 + (b_SH_C3[2] * 4.f * 2.f * yz) * s11 +
                     (b_SH_C3[3] * 3.f * (2.f * zz - xx - yy)) * s12 + (b_SH_C3[4] * 4.f * 2.f * xz) * s13 +
                     (b_SH_C3[5] * (xx - yy)) * s14;
            }
        }
    }
    // rows above the active degree stay zero (the reference returns zero-filled dL_dsh)
    const int used = (deg + 1) * (deg + 1);
    for (int k = used; k < M; k++) {
        dsh[k * 3] = 0.f;
        dsh[k * 3 + 1] = 0.f;
        dsh[k * 3 + 2] = 0.f;
    }
    const f3 ddir = {dot3(dx, dRGB), dot3(dy, dRGB), dot3(dz, dRGB)};
    // dnormvdv (auxiliary.h:107-117)
    const f3 o = dir_orig;
    const float sum2 = o.x * o.x + o.y * o.y + o.z * o.z;
    const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    dmean.x += ((+sum2 - o.x * o.x) * ddir.x - o.y * o.x * ddir.y - o.z * o.x * ddir.z) * invsum32;
    dmean.y += (-o.x * o.y * ddir.x + (sum2 - o.y * o.y) * ddir.y - o.z * o.y * ddir.z) * invsum32;
    dmean.z += (-o.x * o.z * ddir.x - o.y * o.z * ddir.y + (sum2 - o.z * o.z) * ddir.z) * invsum32;
}

__global__ void __launch_bounds__(256) preprocess_backward_kernel(const PreBwdArgs a)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= a.P) return;
    const int M = a.v.M;
    const bool vis = a.radii[i] > 0;
    if (!vis) {
        a.dmeans2D[3 * i] = a.dmeans2D[3 * i + 1] = a.dmeans2D[3 * i + 2] = 0.f;
        a.dcolors[3 * i] = a.dcolors[3 * i + 1] = a.dcolors[3 * i + 2] = 0.f;
        a.dopacity[i] = 0.f;
        a.dmeans3D[3 * i] = a.dmeans3D[3 * i + 1] = a.dmeans3D[3 * i + 2] = 0.f;
#pragma unroll
        for (int k = 0; k < 6; k++) a.dcov3D[6 * i + k] = 0.f;
        a.dscales[3 * i] = a.dscales[3 * i + 1] = a.dscales[3 * i + 2] = 0.f;
        a.drots[4 * i] = a.drots[4 * i + 1] = a.drots[4 * i + 2] = a.drots[4 * i + 3] = 0.f;
        if (a.dsh)
            for (int k = 0; k < M * 3; k++) a.dsh[(size_t)i * M * 3 + k] = 0.f;
        return;
    }
    const float4 *gr = (const float4 *)(a.gacc + (size_t)i * 12);
    const float4 g0 = gr[0], g1 = gr[1];
    const float g2 = a.gacc[(size_t)i * 12 + 8];
    const float dmx = g0.x, dmy = g0.y;
    const float dcx = g0.z, dcy = g0.w, dcz = g1.x;
    a.dmeans2D[3 * i] = dmx;
    a.dmeans2D[3 * i + 1] = dmy;
    a.dmeans2D[3 * i + 2] = 0.f;
    a.dopacity[i] = g1.y;
    a.dcolors[3 * i] = g1.z;
    a.dcolors[3 * i + 1] = g1.w;
    a.dcolors[3 * i + 2] = g2;

    const float *vm = a.v.viewmatrix, *proj = a.v.projmatrix;
    const float mx = a.means[3 * i], my = a.means[3 * i + 1], mz = a.means[3 * i + 2];
    float c3[6];
    float4 q = make_float4(1, 0, 0, 0);
    float sc0 = 0, sc1 = 0, sc2 = 0;
    if (a.cov_pre) {
#pragma unroll
        for (int k = 0; k < 6; k++) c3[k] = a.cov_pre[6 * i + k];
    } else {
        sc0 = a.scales[3 * i], sc1 = a.scales[3 * i + 1], sc2 = a.scales[3 * i + 2];
        q = ((const float4 *)a.rots)[i];
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
    for (int k = 0; k < 6; k++) a.dcov3D[6 * i + k] = dcov[k];
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
        const uint32_t cl = a.aux[i];
        const f3 dRGB = {(cl & 1u) ? 0.f : g1.z, (cl & 2u) ? 0.f : g1.w, (cl & 4u) ? 0.f : g2};
        const float *cp = a.v.campos;
        sh_backward(a.v.D, a.shs + (size_t)i * M * 3, a.dsh + (size_t)i * M * 3, M, {mx - cp[0], my - cp[1], mz - cp[2]},
                    dRGB, dmean);
    }
    a.dmeans3D[3 * i] = dmean.x;
    a.dmeans3D[3 * i + 1] = dmean.y;
    a.dmeans3D[3 * i + 2] = dmean.z;
    // ---- computeCov3D backward (backward.cu:278-341) ----
    if (a.scales) {
        const float r = q.x, x = q.y, y = q.z, z = q.w;
        // R[col][row], glm layout
        const float R[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                               {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                               {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
        const float s[3] = {a.v.scale_modifier * sc0, a.v.scale_modifier * sc1, a.v.scale_modifier * sc2};
        const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                                {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                                {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
        // dL_dM[c][w] = sum_k 2 M[k][w] dS[c][k], M[k][w] = s[w] R[k][w];  dMt[c][w] = dM[w][c]
        float dMt[3][3];
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int w = 0; w < 3; w++)
                dMt[w][c] = 2.0f * s[w] * R[0][w] * dS[c][0] + 2.0f * s[w] * R[1][w] * dS[c][1] + 2.0f * s[w] * R[2][w] * dS[c][2];
#pragma unroll
        for (int k = 0; k < 3; k++)
            a.dscales[3 * i + k] = R[0][k] * dMt[k][0] + R[1][k] * dMt[k][1] + R[2][k] * dMt[k][2];
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
        ((float4 *)a.drots)[i] = dq;
    } else {
        a.dscales[3 * i] = a.dscales[3 * i + 1] = a.dscales[3 * i + 2] = 0.f;
        ((float4 *)a.drots)[i] = make_float4(0, 0, 0, 0);
    }
    if (!a.shs && a.dsh)
        for (int k = 0; k < M * 3; k++) a.dsh[(size_t)i * M * 3 + k] = 0.f;
}

int launch_backward(const SgrView *view, const SgrGaussians *g, const int32_t *radii, const void *geom_buffer,
                    const void *binning_buffer, const void *image_buffer, int64_t num_rendered,
                    const float *dL_dout_color, float *dL_dmeans2D, float *dL_dcolors, float *dL_dopacity,
                    float *dL_dmeans3D, float *dL_dcov3D, float *dL_dsh, float *dL_dscales, float *dL_drotations,
                    void *grad_scratch, cudaStream_t st)
{
    const int P = g->P, W = view->image_width, H = view->image_height;
    const int gx = (W + SGR_TILE - 1) / SGR_TILE, gy = (H + SGR_TILE - 1) / SGR_TILE;
    GeomState geom = GeomState::carve((void *)geom_buffer, P);
    ImageState img = ImageState::carve((void *)image_buffer, W, H);
    BinState bin = BinState::carve((void *)binning_buffer, (size_t)num_rendered);
    float *gacc = (float *)align_up((size_t)grad_scratch);
    SGR_CUDA(cudaMemsetAsync(gacc, 0, (size_t)P * 48, st));
    if (num_rendered > 0) {
        SGR_LAUNCH(K_BLEND_BWD, st,
                   blend_backward_kernel<<<dim3(gx, gy), dim3(SGR_TILE, SGR_TILE), 0, st>>>(
                       img.tile_start, bin.plist, geom.rec, W, H, gx, view->bg, img.final_T, img.n_contrib,
                       dL_dout_color, gacc));
    }
    PreBwdArgs a;
    a.P = P;
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
    a.aux = geom.aux;
    a.gacc = gacc;
    a.dmeans2D = dL_dmeans2D;
    a.dcolors = dL_dcolors;
    a.dopacity = dL_dopacity;
    a.dmeans3D = dL_dmeans3D;
    a.dcov3D = dL_dcov3D;
    a.dsh = dL_dsh;
    a.dscales = dL_dscales;
    a.drots = dL_drotations;
    SGR_LAUNCH(K_PRE_BWD, st, preprocess_backward_kernel<<<(P + 255) / 256, 256, 0, st>>>(a));
    SGR_CUDA(cudaGetLastError());
    if (view->debug) SGR_CUDA(cudaStreamSynchronize(st));
    return SGR_OK;
}

}  // namespace sgr
