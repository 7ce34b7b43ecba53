// sgr_field.cu -- SuGaR surface-regularisation field as fused kernels.
//
// Replaces the PyTorch op chain of SuGaR.get_field_values / compute_density
// (sugar_scene/sugar_model.py:1247-1316, 1345-1368) with its inputs
//   get_covariance(return_sqrt, inverse_scales)   sugar_model.py:730-750   (L^-1 = R(q) diag(1/s))
//   get_beta, beta_mode 'average'                 sugar_model.py:1172-1195
//   pytorch3d.transforms.quaternion_to_matrix     (pytorch3d 0.7.4, real-first, two_s = 2/|q|^2)
// The reference materialises N x K x 3 x 3 inverse-scaled rotations (576 MB at N=1M, K=16) plus
// several N x K x 3 temporaries and their autograd copies; here one lane owns one
// (sample, neighbour) pair, gathers a 48-byte packed Gaussian record, and the K partial
// opacities are summed with shuffles.  Backward recomputes the pair and scatters with
// 16-byte vector reductions into a packed per-Gaussian gradient record.
#include "sgr_internal.cuh"

namespace sgr {

// packed per-Gaussian record, 12 floats:  mu.xyz sig | q.rijk | inv_s.xyz s_min
__global__ void __launch_bounds__(256) field_pack_kernel(int P, const float *__restrict__ points,
                                                         const float *__restrict__ scaling,
                                                         const float *__restrict__ quats,
                                                         const float *__restrict__ strengths, float4 *__restrict__ rec)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    const float s0 = scaling[3 * i], s1 = scaling[3 * i + 1], s2 = scaling[3 * i + 2];
    rec[(size_t)i * 3] = make_float4(points[3 * i], points[3 * i + 1], points[3 * i + 2], strengths[i]);
    rec[(size_t)i * 3 + 1] = make_float4(quats[4 * i], quats[4 * i + 1], quats[4 * i + 2], quats[4 * i + 3]);
    // 1 / scaling.clamp(min=1e-8)   (sugar_model.py:732-733)
    rec[(size_t)i * 3 + 2] =
        make_float4(1.0f / fmaxf(s0, 1e-8f), 1.0f / fmaxf(s1, 1e-8f), 1.0f / fmaxf(s2, 1e-8f), fminf(s0, fminf(s1, s2)));
}

struct Rot {
    float m[3][3];  // row-major R[row][col]
    float ts;       // two_s = 2 / |q|^2
};
__device__ __forceinline__ Rot quat_to_rot(float4 q)
{
    const float r = q.x, i = q.y, j = q.z, k = q.w;
    Rot o;
    o.ts = 2.0f / (r * r + i * i + j * j + k * k);
    const float ts = o.ts;
    o.m[0][0] = 1 - ts * (j * j + k * k);
    o.m[0][1] = ts * (i * j - k * r);
    o.m[0][2] = ts * (i * k + j * r);
    o.m[1][0] = ts * (i * j + k * r);
    o.m[1][1] = 1 - ts * (i * i + k * k);
    o.m[1][2] = ts * (j * k - i * r);
    o.m[2][0] = ts * (i * k - j * r);
    o.m[2][1] = ts * (j * k + i * r);
    o.m[2][2] = 1 - ts * (i * i + j * j);
    return o;
}

template <int G>
__device__ __forceinline__ float group_sum(float v)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct FieldArgs {
    int N, K, P, group;
    float density_factor, sdf_offset, opacity_min_clamp;
    const float *x;
    const int64_t *idx;
    const float4 *rec;
};

// straight-through clamp of sugar_model.py:1280-1281 followed by the SDF of :1303-1306
__device__ __forceinline__ void sdf_from_density(float density, float beta, float opacity_min_clamp, float sdf_offset,
                                                 float &sdf, float &root, float &cl, float &dcl_ddens)
{
    float d = density;
    dcl_ddens = 1.0f;
    if (density >= 1.0f) {
        d = density / (density + 1e-12f);
        dcl_ddens = 1.0f / (density + 1e-12f);
    }
    cl = fmaxf(d, opacity_min_clamp);
    if (d < opacity_min_clamp) dcl_ddens = 0.0f;
    root = sqrtf(-2.0f * logf(cl));
    sdf = beta * (root - sdf_offset);
}

template <int G>
__global__ void __launch_bounds__(256) field_forward_kernel(const FieldArgs a, float *__restrict__ density,
                                                            float *__restrict__ nbr_opacity, float *__restrict__ beta,
                                                            float *__restrict__ sdf)
{
    constexpr int SPB = 256 / G;  // samples per block
    const int g = threadIdx.x / G, l = threadIdx.x % G;
    const int n = blockIdx.x * SPB + g;
    const bool live = n < a.N;
    const int nrow = a.group > 1 ? n / a.group : n;  // neighbour row shared by `group` consecutive samples
    float dens = 0.f, bsum = 0.f;
    if (live) {
        const float x0 = a.x[3 * n], x1 = a.x[3 * n + 1], x2 = a.x[3 * n + 2];
        for (int k = l; k < a.K; k += G) {
            const int64_t id = a.idx[(size_t)nrow * a.K + k];
            const float4 r0 = __ldg(a.rec + id * 3), r1 = __ldg(a.rec + id * 3 + 1), r2 = __ldg(a.rec + id * 3 + 2);
            const Rot R = quat_to_rot(r1);
            const float sh0 = x0 - r0.x, sh1 = x1 - r0.y, sh2 = x2 - r0.z;
            const float w0 = (R.m[0][0] * sh0 + R.m[1][0] * sh1 + R.m[2][0] * sh2) * r2.x;
            const float w1 = (R.m[0][1] * sh0 + R.m[1][1] * sh1 + R.m[2][1] * sh2) * r2.y;
            const float w2 = (R.m[0][2] * sh0 + R.m[1][2] * sh1 + R.m[2][2] * sh2) * r2.z;
            const float d2 = fminf(fmaxf(w0 * w0 + w1 * w1 + w2 * w2, 0.f), 1e8f);
            const float o = a.density_factor * r0.w * expf(-0.5f * d2);
            if (nbr_opacity) nbr_opacity[(size_t)n * a.K + k] = o;
            dens += o;
            bsum += r2.w;
        }
    }
    dens = group_sum<G>(dens);
    bsum = group_sum<G>(bsum);
    if (live && l == 0) {
        const float b = bsum / (float)a.K;
        if (density) density[n] = dens;
        if (beta) beta[n] = b;
        if (sdf) {
            float s, root, cl, dd;
            sdf_from_density(dens, b, a.opacity_min_clamp, a.sdf_offset, s, root, cl, dd);
            sdf[n] = s;
        }
    }
}

__device__ __forceinline__ void red_v4(float *addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// gradient record per Gaussian, 12 floats: dmu.xyz dsig | dq.rijk | ds.xyz -
template <int G>
__global__ void __launch_bounds__(256) field_backward_kernel(const FieldArgs a, const float *__restrict__ scaling,
                                                             const float *__restrict__ g_density,
                                                             const float *__restrict__ g_nbr, const float *__restrict__ g_beta,
                                                             const float *__restrict__ g_sdf, float *__restrict__ g_x,
                                                             float *__restrict__ grec)
{
    constexpr int SPB = 256 / G;
    const int g = threadIdx.x / G, l = threadIdx.x % G;
    const int n = blockIdx.x * SPB + g;
    const bool live = n < a.N;
    const int nrow = a.group > 1 ? n / a.group : n;
    float x0 = 0, x1 = 0, x2 = 0;
    if (live) x0 = a.x[3 * n], x1 = a.x[3 * n + 1], x2 = a.x[3 * n + 2];
    // pass 1: recompute density and beta for the sample (needed for d sdf / d density)
    float dens = 0.f, bsum = 0.f;
    if (live)
        for (int k = l; k < a.K; k += G) {
            const int64_t id = a.idx[(size_t)nrow * a.K + k];
            const float4 r0 = __ldg(a.rec + id * 3), r1 = __ldg(a.rec + id * 3 + 1), r2 = __ldg(a.rec + id * 3 + 2);
            const Rot R = quat_to_rot(r1);
            const float sh0 = x0 - r0.x, sh1 = x1 - r0.y, sh2 = x2 - r0.z;
            const float w0 = (R.m[0][0] * sh0 + R.m[1][0] * sh1 + R.m[2][0] * sh2) * r2.x;
            const float w1 = (R.m[0][1] * sh0 + R.m[1][1] * sh1 + R.m[2][1] * sh2) * r2.y;
            const float w2 = (R.m[0][2] * sh0 + R.m[1][2] * sh1 + R.m[2][2] * sh2) * r2.z;
            const float d2 = fminf(fmaxf(w0 * w0 + w1 * w1 + w2 * w2, 0.f), 1e8f);
            dens += a.density_factor * r0.w * expf(-0.5f * d2);
            bsum += r2.w;
        }
    dens = group_sum<G>(dens);
    bsum = group_sum<G>(bsum);
    float gd = 0.f, gb = 0.f;  // dL/d density (total), dL/d beta (total)
    if (live) {
        const float b = bsum / (float)a.K;
        gd = g_density ? g_density[n] : 0.f;
        gb = g_beta ? g_beta[n] : 0.f;
        if (g_sdf) {
            float s, root, cl, dd;
            sdf_from_density(dens, b, a.opacity_min_clamp, a.sdf_offset, s, root, cl, dd);
            const float gs = g_sdf[n];
            gb += gs * (root - a.sdf_offset);
            if (dd != 0.f) gd += gs * (-b / (cl * root)) * dd;
        }
    }
    // pass 2: per-pair gradients
    float gx0 = 0.f, gx1 = 0.f, gx2 = 0.f;
    if (live)
        for (int k = l; k < a.K; k += G) {
            const int64_t id = a.idx[(size_t)nrow * a.K + k];
            const float4 r0 = __ldg(a.rec + id * 3), r1 = __ldg(a.rec + id * 3 + 1), r2 = __ldg(a.rec + id * 3 + 2);
            const Rot R = quat_to_rot(r1);
            const float sh[3] = {x0 - r0.x, x1 - r0.y, x2 - r0.z};
            float u[3], w[3];
#pragma unroll
            for (int j = 0; j < 3; j++) u[j] = R.m[0][j] * sh[0] + R.m[1][j] * sh[1] + R.m[2][j] * sh[2];
            w[0] = u[0] * r2.x;
            w[1] = u[1] * r2.y;
            w[2] = u[2] * r2.z;
            const float d2raw = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
            const float d2 = fminf(fmaxf(d2raw, 0.f), 1e8f);
            const float e = a.density_factor * expf(-0.5f * d2);
            const float o = r0.w * e;
            const float Gk = gd + (g_nbr ? g_nbr[(size_t)n * a.K + k] : 0.f);
            const float dsig = Gk * e;
            const float dd2 = (d2raw > 1e8f || d2raw < 0.f) ? 0.f : Gk * o * -0.5f;
            const float invs[3] = {r2.x, r2.y, r2.z};
            float du[3], ds[3];
            const float *sc = scaling + id * 3;
            int amin = 0;
            float smin = sc[0];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const float dw = dd2 * 2.0f * w[j];
                du[j] = dw * invs[j];
                const float sj = sc[j];
                ds[j] = (sj >= 1e-8f) ? -(dw * u[j]) * invs[j] * invs[j] : 0.f;
                if (j > 0 && sj < smin) {
                    smin = sj;
                    amin = j;
                }
            }
            ds[amin] += gb / (float)a.K;
            float dsh[3];
#pragma unroll
            for (int i = 0; i < 3; i++) dsh[i] = R.m[i][0] * du[0] + R.m[i][1] * du[1] + R.m[i][2] * du[2];
            gx0 += dsh[0];
            gx1 += dsh[1];
            gx2 += dsh[2];
            // dL/dR[i][j] = sh[i] * du[j];  R = I + ts * E(q)
            float A[3][3];
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) A[i][j] = sh[i] * du[j];
            const float qr = r1.x, qi = r1.y, qj = r1.z, qk = r1.w, ts = R.ts;
            const float dts = A[0][0] * -(qj * qj + qk * qk) + A[0][1] * (qi * qj - qk * qr) + A[0][2] * (qi * qk + qj * qr) +
                              A[1][0] * (qi * qj + qk * qr) + A[1][1] * -(qi * qi + qk * qk) + A[1][2] * (qj * qk - qi * qr) +
                              A[2][0] * (qi * qk - qj * qr) + A[2][1] * (qj * qk + qi * qr) + A[2][2] * -(qi * qi + qj * qj);
            const float c = -ts * ts * dts;  // d ts / d q_c = -ts^2 q_c
            const float dqr = ts * (-qk * A[0][1] + qj * A[0][2] + qk * A[1][0] - qi * A[1][2] - qj * A[2][0] + qi * A[2][1]) + c * qr;
            const float dqi = ts * (qj * (A[0][1] + A[1][0]) + qk * (A[0][2] + A[2][0]) + qr * (A[2][1] - A[1][2]) -
                                    2.f * qi * (A[1][1] + A[2][2])) + c * qi;
            const float dqj = ts * (qi * (A[0][1] + A[1][0]) + qr * (A[0][2] - A[2][0]) + qk * (A[1][2] + A[2][1]) -
                                    2.f * qj * (A[0][0] + A[2][2])) + c * qj;
            const float dqk = ts * (qr * (A[1][0] - A[0][1]) + qi * (A[0][2] + A[2][0]) + qj * (A[1][2] + A[2][1]) -
                                    2.f * qk * (A[0][0] + A[1][1])) + c * qk;
            float *gr = grec + id * 12;
            red_v4(gr, -dsh[0], -dsh[1], -dsh[2], dsig);
            red_v4(gr + 4, dqr, dqi, dqj, dqk);
            red_v4(gr + 8, ds[0], ds[1], ds[2], 0.f);
        }
    gx0 = group_sum<G>(gx0);
    gx1 = group_sum<G>(gx1);
    gx2 = group_sum<G>(gx2);
    if (live && l == 0 && g_x) {
        g_x[3 * n] = gx0;
        g_x[3 * n + 1] = gx1;
        g_x[3 * n + 2] = gx2;
    }
}

__global__ void __launch_bounds__(256) field_unpack_kernel(int P, const float4 *__restrict__ grec, float *g_points,
                                                           float *g_scaling, float *g_quats, float *g_strengths)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    const float4 a = grec[(size_t)i * 3], b = grec[(size_t)i * 3 + 1], c = grec[(size_t)i * 3 + 2];
    if (g_points) g_points[3 * i] = a.x, g_points[3 * i + 1] = a.y, g_points[3 * i + 2] = a.z;
    if (g_strengths) g_strengths[i] = a.w;
    if (g_quats) ((float4 *)g_quats)[i] = b;
    if (g_scaling) g_scaling[3 * i] = c.x, g_scaling[3 * i + 1] = c.y, g_scaling[3 * i + 2] = c.z;
}

static int group_width(int K)
{
    int G = 1;
    while (G < K && G < 32) G <<= 1;
    return G;
}

}  // namespace sgr

using namespace sgr;

extern "C" {

size_t sgr_field_scratch_bytes(int32_t P) { return align_up((size_t)(P < 0 ? 0 : P) * 48) * 2 + SGR_ALIGN; }

static int field_check(const SgrFieldParams *p, const void *x, const void *idx, const void *a, const void *b,
                       const void *c, const void *d, const void *scratch)
{
    if (!p || p->N < 0 || p->K <= 0 || p->P <= 0) {
        set_error("bad field sizes");
        return SGR_EINVAL;
    }
    if (p->N > 0 && (!x || !idx || !a || !b || !c || !d || !scratch)) {
        set_error("null pointer passed to sgr_field_*");
        return SGR_EINVAL;
    }
    return SGR_OK;
}

static FieldArgs make_args(const SgrFieldParams *p, const float *x, const int64_t *idx, const float4 *rec)
{
    FieldArgs a;
    a.N = p->N;
    a.K = p->K;
    a.P = p->P;
    a.group = p->samples_per_idx_row > 1 ? p->samples_per_idx_row : 1;
    a.density_factor = p->density_factor;
    // np.sqrt(-2. * np.log(min(density_threshold, 1.)))   (sugar_model.py:1305), evaluated in double like numpy
    const double thr = p->density_threshold < 1.0f ? (double)p->density_threshold : 1.0;
    a.sdf_offset = (float)sqrt(-2.0 * log(thr));
    a.opacity_min_clamp = p->opacity_min_clamp;
    a.x = x;
    a.idx = idx;
    a.rec = rec;
    return a;
}

int sgr_field_forward(const SgrFieldParams *p, const float *x, const int64_t *nbr_idx, const float *points,
                      const float *scaling, const float *quaternions, const float *strengths, float *density,
                      float *nbr_opacity, float *beta, float *sdf, void *scratch, void *stream)
{
    int rc = field_check(p, x, nbr_idx, points, scaling, quaternions, strengths, scratch);
    if (rc) return rc;
    if (p->N == 0) return SGR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    float4 *rec = (float4 *)align_up((size_t)scratch);
    SGR_LAUNCH(K_FIELD_PACK, st,
               field_pack_kernel<<<(p->P + 255) / 256, 256, 0, st>>>(p->P, points, scaling, quaternions, strengths, rec));
    const FieldArgs a = make_args(p, x, nbr_idx, rec);
    const int G = group_width(p->K);
    const int blocks = (p->N + 256 / G - 1) / (256 / G);
    sgr::prof_begin(K_FIELD_FWD, st);
    switch (G) {
        case 1: field_forward_kernel<1><<<blocks, 256, 0, st>>>(a, density, nbr_opacity, beta, sdf); break;
        case 2: field_forward_kernel<2><<<blocks, 256, 0, st>>>(a, density, nbr_opacity, beta, sdf); break;
        case 4: field_forward_kernel<4><<<blocks, 256, 0, st>>>(a, density, nbr_opacity, beta, sdf); break;
        case 8: field_forward_kernel<8><<<blocks, 256, 0, st>>>(a, density, nbr_opacity, beta, sdf); break;
        case 16: field_forward_kernel<16><<<blocks, 256, 0, st>>>(a, density, nbr_opacity, beta, sdf); break;
        default: field_forward_kernel<32><<<blocks, 256, 0, st>>>(a, density, nbr_opacity, beta, sdf); break;
    }
    sgr::prof_end(st);
    SGR_CUDA(cudaGetLastError());
    return SGR_OK;
}

int sgr_field_backward(const SgrFieldParams *p, const float *x, const int64_t *nbr_idx, const float *points,
                       const float *scaling, const float *quaternions, const float *strengths,
                       const float *g_density, const float *g_nbr_opacity, const float *g_beta, const float *g_sdf,
                       float *g_x, float *g_points, float *g_scaling, float *g_quaternions, float *g_strengths,
                       void *scratch, void *stream)
{
    int rc = field_check(p, x, nbr_idx, points, scaling, quaternions, strengths, scratch);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    float4 *rec = (float4 *)align_up((size_t)scratch);
    float *grec = (float *)((char *)rec + align_up((size_t)p->P * 48));
    SGR_CUDA(cudaMemsetAsync(grec, 0, (size_t)p->P * 48, st));
    if (p->N > 0) {
        SGR_LAUNCH(K_FIELD_PACK, st,
                   field_pack_kernel<<<(p->P + 255) / 256, 256, 0, st>>>(p->P, points, scaling, quaternions, strengths, rec));
        const FieldArgs a = make_args(p, x, nbr_idx, rec);
        const int G = group_width(p->K);
        const int blocks = (p->N + 256 / G - 1) / (256 / G);
#define SGR_FB(GW)                                                                                                   \
    field_backward_kernel<GW><<<blocks, 256, 0, st>>>(a, scaling, g_density, g_nbr_opacity, g_beta, g_sdf, g_x, grec)
        sgr::prof_begin(K_FIELD_BWD, st);
        switch (G) {
            case 1: SGR_FB(1); break;
            case 2: SGR_FB(2); break;
            case 4: SGR_FB(4); break;
            case 8: SGR_FB(8); break;
            case 16: SGR_FB(16); break;
            default: SGR_FB(32); break;
        }
        sgr::prof_end(st);
#undef SGR_FB
    }
    SGR_LAUNCH(K_FIELD_UNPACK, st,
               field_unpack_kernel<<<(p->P + 255) / 256, 256, 0, st>>>(p->P, (const float4 *)grec, g_points, g_scaling,
                                                                       g_quaternions, g_strengths));
    SGR_CUDA(cudaGetLastError());
    return SGR_OK;
}

}  // extern "C"
