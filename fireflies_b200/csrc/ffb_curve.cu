// NURBS-curve camera paths and the small batched geometry helpers, sm_100a.
//
// Replaces (reference paths relative to the Fireflies tree):
//   fireflies/entity/curve.py:48-96 (Curve.sample_rotation / sample_translation / randomize) together with
//   geomdl==5.3.1 NURBS.Curve.evaluate_single (requirements.txt:16; utils/io.py:77-108 builds the curve)
//                                                                -> ffb_nurbs_curve_eval, ffb_curve_pose
//   fireflies/utils/intersections.py:5-33                        -> ffb_ray_plane, ffb_sphere_sphere
//
// geomdl evaluates in Python floats: the curve point is formed in fp64 with the operation order of the published
// algorithm (Piegl & Tiller A2.2 / A4.1: linear span walk, triangular basis table, homogeneous sum, divide) and only
// then rounded to fp32, exactly what `torch.tensor(curve.evaluate_single(t))` does.  Explicit round-to-nearest
// intrinsics keep the compiler from contracting the products into FMAs, so the fp64 result is reproducible bit for bit.
// The reference walks its path one Python call per sample (two evaluations + ~40 aten ops); here a batch of B path
// parameters is one launch, one thread per sample (the work per sample is ~degree^2 flops on a few cached knots).
#include "ffb_common.cuh"

namespace ffb {
namespace curve {

constexpr int kMaxDegree = FFB_NURBS_MAX_DEGREE;

__device__ __forceinline__ void curve_point(const double* __restrict__ ctrlw, const double* __restrict__ knots, int n, int p,
                                            double t, double (&out)[3]) {
    int span = p + 1;                                   // linear walk: first knot beyond t, clamped at the last span
    while (span < n && knots[span] <= t) ++span;
    --span;
    double left[kMaxDegree + 1], right[kMaxDegree + 1], N[kMaxDegree + 1];
#pragma unroll
    for (int j = 0; j <= kMaxDegree; ++j) { left[j] = 0.0; right[j] = 0.0; N[j] = 1.0; }
    for (int j = 1; j <= p; ++j) {
        left[j] = __dsub_rn(t, knots[span + 1 - j]);
        right[j] = __dsub_rn(knots[span + j], t);
        double saved = 0.0;
        for (int r = 0; r < j; ++r) {
            const double temp = __ddiv_rn(N[r], __dadd_rn(right[r + 1], left[j - r]));
            N[r] = __dadd_rn(saved, __dmul_rn(right[r + 1], temp));
            saved = __dmul_rn(left[j - r], temp);
        }
        N[j] = saved;
    }
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = 0; i <= p; ++i) {
        const double* c = ctrlw + 4 * (span - p + i);
#pragma unroll
        for (int d = 0; d < 4; ++d) acc[d] = __dadd_rn(acc[d], __dmul_rn(N[i], c[d]));
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) out[d] = __ddiv_rn(acc[d], acc[3]);
}

__global__ void __launch_bounds__(128) nurbs_eval_kernel(const double* __restrict__ ctrlw, const double* __restrict__ knots, int n,
                                                         int p, const double* __restrict__ t, int B, double* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double c[3];
    curve_point(ctrlw, knots, n, p, t[b], c);
    out[3 * b + 0] = c[0];
    out[3 * b + 1] = c[1];
    out[3 * b + 2] = c[2];
}

// T(C(t)) @ toMat4x4(Rodrigues([0,1,0] -> d)) @ W, d = fp32(C(t+dt)) - fp32(C(t)) with x and z negated (curve.py:48-96).
__global__ void __launch_bounds__(128) curve_pose_kernel(const double* __restrict__ ctrlw, const double* __restrict__ knots, int n,
                                                         int p, const double* __restrict__ t, int B, double dt,
                                                         const float* __restrict__ world, float* __restrict__ out_world,
                                                         float* __restrict__ out_rot, float* __restrict__ out_trans) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double c0[3], c1[3];
    const double tb = t[b];
    curve_point(ctrlw, knots, n, p, __dadd_rn(tb, dt), c1);
    curve_point(ctrlw, knots, n, p, tb, c0);
    const float p0[3] = {__double2float_rn(c0[0]), __double2float_rn(c0[1]), __double2float_rn(c0[2])};
    float d[3] = {__fsub_rn(__double2float_rn(c1[0]), p0[0]), __fsub_rn(__double2float_rn(c1[1]), p0[1]),
                  __fsub_rn(__double2float_rn(c1[2]), p0[2])};
    d[0] = -d[0];
    d[2] = -d[2];
    // F.normalize(d, dim=0): d / max(|d|, 1e-12); the source direction [0,1,0] is already unit length
    const float len = fmaxf(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2]))), 1e-12f);
    const float v[3] = {__fdiv_rn(d[0], len), __fdiv_rn(d[1], len), __fdiv_rn(d[2], len)};
    // cross([0,1,0], v) = (v.z, 0, -v.x); dot = v.y
    const float cx = v[2], cy = 0.0f, cz = -v[0];
    const float dot = v[1];
    const float K[3][3] = {{0.0f, -cz, cy}, {cz, 0.0f, -cx}, {-cy, cx, 0.0f}};
    const float cn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz)));
    const float n2 = __fmul_rn(cn, cn);
    const float one_m_dot = __fsub_rn(1.0f, dot);
    float R[4][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float k2 = __fadd_rn(__fadd_rn(__fmul_rn(K[i][0], K[0][j]), __fmul_rn(K[i][1], K[1][j])), __fmul_rn(K[i][2], K[2][j]));
            const float corr = __fdiv_rn(__fmul_rn(k2, one_m_dot), n2);          // 0/0 = NaN for d parallel to [0,1,0], as in the reference
            R[i][j] = __fadd_rn(__fadd_rn(i == j ? 1.0f : 0.0f, K[i][j]), corr);
        }
#pragma unroll
    for (int i = 0; i < 3; ++i) { R[i][3] = 0.0f; R[3][i] = 0.0f; }
    R[3][3] = 1.0f;
    if (out_rot) {
#pragma unroll
        for (int i = 0; i < 16; ++i) out_rot[16 * b + i] = R[i >> 2][i & 3];
    }
    if (out_trans) {
#pragma unroll
        for (int i = 0; i < 16; ++i) out_trans[16 * b + i] = ((i >> 2) == (i & 3)) ? 1.0f : ((i & 3) == 3 ? p0[i >> 2] : 0.0f);
    }
    if (out_world) {
        // T @ R: R's last column is (0,0,0,1), so the product is R with the translation written into it (exact)
#pragma unroll
        for (int i = 0; i < 3; ++i) R[i][3] = p0[i];
        float W[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) W[i] = world[i];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float s = __fmul_rn(R[i][0], W[j]);
                s = fmaf(R[i][1], W[4 + j], s);
                s = fmaf(R[i][2], W[8 + j], s);
                s = fmaf(R[i][3], W[12 + j], s);
                out_world[16 * b + 4 * i + j] = s;
            }
    }
}

// intersections.py:5-12: t = ((po - o) . n) / (n . d); |n . d| < 1e-6 -> denom/denom (1, or NaN when exactly 0)
__global__ void __launch_bounds__(128) ray_plane_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                                        const float* __restrict__ po, const float* __restrict__ pn, int N,
                                                        float* __restrict__ t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float* n = pn + 3 * i;
    float den = __fadd_rn(__fadd_rn(__fmul_rn(n[0], d[3 * i]), __fmul_rn(n[1], d[3 * i + 1])), __fmul_rn(n[2], d[3 * i + 2]));
    if (fabsf(den) < 0.000001f) den = __fdiv_rn(den, den);
    const float num = __fadd_rn(__fadd_rn(__fmul_rn(__fsub_rn(po[3 * i], o[3 * i]), n[0]),
                                          __fmul_rn(__fsub_rn(po[3 * i + 1], o[3 * i + 1]), n[1])),
                                __fmul_rn(__fsub_rn(po[3 * i + 2], o[3 * i + 2]), n[2]));
    t[i] = __fdiv_rn(num, den);
}

// intersections.py:26-33: |a - b|^2 <= (ra + rb)^2, D coordinates per centre
__global__ void __launch_bounds__(128) sphere_sphere_kernel(const float* __restrict__ a, const float* __restrict__ ra,
                                                            const float* __restrict__ b, const float* __restrict__ rb, int N, int D,
                                                            uint8_t* __restrict__ hit) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float s = 0.0f;
    for (int k = 0; k < D; ++k) {
        const float e = __fsub_rn(a[(size_t)i * D + k], b[(size_t)i * D + k]);
        s = __fadd_rn(s, __fmul_rn(e, e));
    }
    const float r = __fadd_rn(ra[i], rb[i]);
    hit[i] = s <= __fmul_rn(r, r) ? 1 : 0;
}

static int check_curve(const double* ctrlw, const double* knots, int32_t n, int32_t p, const double* t, int32_t B, const char* who) {
    if (!ctrlw || !knots || (!t && B > 0) || B < 0) return fail_arg(FFB_E_ARG, who);
    if (p < 1 || p > kMaxDegree) return fail_arg(FFB_E_ARG, "nurbs: degree must be in [1, FFB_NURBS_MAX_DEGREE]");
    if (n < p + 1) return fail_arg(FFB_E_ARG, "nurbs: a curve of degree p needs at least p+1 control points");
    return 0;
}

}  // namespace curve
}  // namespace ffb

using namespace ffb;
using namespace ffb::curve;

extern "C" int ffb_nurbs_curve_eval(const double* ctrlw, const double* knots, int32_t n_ctrl, int32_t degree, const double* t,
                                    int32_t B, double* out, void* stream) {
    if (int rc = check_curve(ctrlw, knots, n_ctrl, degree, t, B, "nurbs_curve_eval: bad argument")) return rc;
    if (B == 0) return 0;
    if (!out) return fail_arg(FFB_E_ARG, "nurbs_curve_eval: null out");
    nurbs_eval_kernel<<<(B + 127) / 128, 128, 0, as_stream(stream)>>>(ctrlw, knots, n_ctrl, degree, t, B, out);
    FFB_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ffb_curve_pose(const double* ctrlw, const double* knots, int32_t n_ctrl, int32_t degree, const double* t, int32_t B,
                              double dt, const float* world, float* out_world, float* out_rot, float* out_trans, void* stream) {
    if (int rc = check_curve(ctrlw, knots, n_ctrl, degree, t, B, "curve_pose: bad argument")) return rc;
    if (B == 0) return 0;
    if (out_world && !world) return fail_arg(FFB_E_ARG, "curve_pose: out_world needs world");
    if (!out_world && !out_rot && !out_trans) return fail_arg(FFB_E_ARG, "curve_pose: no output");
    curve_pose_kernel<<<(B + 127) / 128, 128, 0, as_stream(stream)>>>(ctrlw, knots, n_ctrl, degree, t, B, dt, world, out_world, out_rot,
                                                                     out_trans);
    FFB_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ffb_ray_plane(const float* origin, const float* direction, const float* plane_origin, const float* plane_normal,
                             int32_t N, float* t_out, void* stream) {
    if (!origin || !direction || !plane_origin || !plane_normal || !t_out || N < 0) return fail_arg(FFB_E_ARG, "ray_plane: bad argument");
    if (N == 0) return 0;
    ray_plane_kernel<<<(N + 127) / 128, 128, 0, as_stream(stream)>>>(origin, direction, plane_origin, plane_normal, N, t_out);
    FFB_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int ffb_sphere_sphere(const float* a, const float* ra, const float* b, const float* rb, int32_t N, int32_t D, uint8_t* hit,
                                 void* stream) {
    if (!a || !ra || !b || !rb || !hit || N < 0 || D < 1) return fail_arg(FFB_E_ARG, "sphere_sphere: bad argument");
    if (N == 0) return 0;
    sphere_sphere_kernel<<<(N + 127) / 128, 128, 0, as_stream(stream)>>>(a, ra, b, rb, N, D, hit);
    FFB_CUDA(cudaGetLastError());
    return 0;
}
