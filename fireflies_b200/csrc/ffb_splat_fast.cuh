// Fast splat kernels (included by ffb_splat.cu inside namespace ffb::splat).
//
// Same tiling as the general kernels (CTA = 64x32 texels, warp = 8 rows x 32 columns, lane = column) but the
// inner loop is rebuilt around what the first ncu capture showed: the general kernels are instruction-issue
// bound (fwd 163, bwd 386 thread-instructions per texel at 80 % issue utilisation, DRAM at 17 % / 7 %).
//   * windows are symmetric Chebyshev masks around floor(P): |c - floor(P0)| <= H and |r - floor(P1)| <= H,
//     evaluated arithmetically (FFMA.SAT) instead of four integer compares + selects per texel; identical to
//     the reference's clipped footprint whenever the texture is larger than the footprint (the dispatcher
//     falls back to the general kernels otherwise);
//   * everything that only depends on (candidate, row) -- (r - P1)^2, r - P1, the row masks -- is computed once per
//     CTA while staging the candidates and read back with broadcast LDS.128;
//   * per-warp candidate culling is one ballot over 32 staged candidates instead of a test per candidate per lane;
//   * two rows are processed per instruction with the sm_100 packed fp32 pipe (FADD2 / FMUL2 / FFMA2);
//   * g = 2^(d2^2 * K), K = -log2(e)/sigma^2: two multiplies and ONE MUFU.EX2.  d2 is bit-identical to the
//     reference's; the exponent carries <= 1.5e-7 relative error, i.e. <= w * 1.5e-7 on g (w = (d2/sigma)^2;
//     2.4e-6 at g = 1e-7), inside the 1e-5 forward budget.
#pragma once

constexpr int FCH = 32;                   // candidates staged per chunk (one ballot)

struct FastConsts {
    float K2;                             // -log2(e) / sigma^2
    float thr_s, thr_o;                   // 4*H + 2 for the sum / soft-OR Chebyshev masks
};

__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float cheb_mask(float e, float thr) {      // 1 if |e| <= H else 0 (e integer valued)
    return __saturatef(fmaf(fabsf(e), -4.f, thr));
}
__device__ __forceinline__ float2 bc(float x) { return make_float2(x, x); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }

// Per-chunk shared staging: candidate records + the per-(candidate,row) tables for the CTA's 32 rows.
template <bool NEED_DY, bool MASK_O>
struct Stage {
    float4 pf[FCH];                       // p0, p1, floor(p0), floor(p1)
    uint2 u[FCH];                         // union windows (columns, rows)
    int idx[FCH];                         // point index
    float dy2[FCH][TH];
    float ms[FCH][TH];
    float dy[NEED_DY ? FCH : 1][TH];
    float mo[MASK_O ? FCH : 1][TH];
};

template <bool NEED_DY, bool MASK_O>
__device__ __forceinline__ void stage_chunk(Stage<NEED_DY, MASK_O>& s, const PointRec* __restrict__ recs, const int* __restrict__ list,
                                            int base, int n, int row0, const FastConsts& fc, int tid) {
    __syncthreads();                      // previous chunk fully consumed
    if (tid < n) {
        const int id = list[base + tid];
        const PointRec r = recs[id];
        s.pf[tid] = make_float4(r.p0, r.p1, floorf(r.p0), floorf(r.p1));
        s.u[tid] = make_uint2(r.uc, r.ur);
        s.idx[tid] = id;
    }
    __syncthreads();
    const int row = tid & 31;
    const float rf = (float)(row0 + row);
    for (int c = tid >> 5; c < n; c += CTA / 32) {
        const float4 pf = s.pf[c];
        const float d = rf - pf.y;
        s.dy2[c][row] = __fmul_rn(d, d);
        const float e = rf - pf.w;
        s.ms[c][row] = cheb_mask(e, fc.thr_s);
        if (NEED_DY) s.dy[c][row] = d;
        if (MASK_O) s.mo[c][row] = cheb_mask(e, fc.thr_o);
    }
    __syncthreads();
}

template <bool NEED_DY, bool MASK_O>
__device__ __forceinline__ unsigned cull_chunk(const Stage<NEED_DY, MASK_O>& s, int n, int lane, int wc0, int wr0) {
    bool hit = false;
    if (lane < n) {
        const uint2 u = s.u[lane];
        hit = !((int)(u.x >> 16) <= wc0 || (int)(u.x & 0xffff) >= wc0 + 32 || (int)(u.y >> 16) <= wr0 || (int)(u.y & 0xffff) >= wr0 + WROWS);
    }
    return __ballot_sync(0xffffffffu, hit);
}

template <bool SUM, bool SOFTOR, bool SUM_T, bool MASK_O>
__global__ void __launch_bounds__(CTA) splat_fwd_fast(RasterParams q, FastConsts fc) {
    __shared__ Stage<false, MASK_O> st;
    const int tile = blockIdx.x % q.T, b = blockIdx.x / q.T;
    const int bin = q.shared_pattern ? 0 : b;
    const int tx = tile % q.tgx, ty = tile / q.tgx;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wrow = (warp >> 1) * WROWS;
    const int wc0 = tx * TW + (warp & 1) * 32, wr0 = ty * TH + wrow;
    const int c = wc0 + lane;
    const float cf = (float)c;
    const int* toff = q.tile_off + (size_t)bin * (q.T + 1);
    const int beg = toff[tile], end = toff[tile + 1];
    const int* list = q.list + (size_t)bin * q.cap;
    const PointRec* recs = q.recs + (size_t)bin * q.N;

    float2 acc_s[WROWS / 2], acc_p[WROWS / 2];
#pragma unroll
    for (int j = 0; j < WROWS / 2; ++j) { acc_s[j] = bc(0.f); acc_p[j] = bc(1.f); }

    for (int base = beg; base < end; base += FCH) {
        const int n = min(FCH, end - base);
        stage_chunk(st, recs, list, base, n, ty * TH, fc, tid);
        unsigned m = cull_chunk(st, n, lane, wc0, wr0);
        while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            const float4 pf = st.pf[k];
            const float dx = cf - pf.x;
            const float dx2 = __fmul_rn(dx, dx);
            const float ex = cf - pf.z;
            const float mcs = cheb_mask(ex, fc.thr_s);
            const float mco = MASK_O ? -cheb_mask(ex, fc.thr_o) : 0.f;
            const float4* d2p = reinterpret_cast<const float4*>(&st.dy2[k][wrow]);
            const float4* msp = reinterpret_cast<const float4*>(&st.ms[k][wrow]);
            const float4* mop = reinterpret_cast<const float4*>(&st.mo[MASK_O ? k : 0][MASK_O ? wrow : 0]);
            float4 D[2] = {d2p[0], d2p[1]}, Ms[2], Mo[2];
            if (SUM) { Ms[0] = msp[0]; Ms[1] = msp[1]; }
            if (SOFTOR && MASK_O) { Mo[0] = mop[0]; Mo[1] = mop[1]; }
#pragma unroll
            for (int j = 0; j < WROWS / 2; ++j) {
                const float2 dy2 = (j & 1) ? make_float2(D[j >> 1].z, D[j >> 1].w) : make_float2(D[j >> 1].x, D[j >> 1].y);
                const float2 d2 = __fadd2_rn(bc(dx2), dy2);                       // dc*dc + dr*dr, as the reference
                const float2 t = __fmul2_rn(__fmul2_rn(d2, d2), bc(fc.K2));
                const float2 g = make_float2(ex2_approx(t.x), ex2_approx(t.y));
                if (SUM) {
                    const float2 mr = (j & 1) ? make_float2(Ms[j >> 1].z, Ms[j >> 1].w) : make_float2(Ms[j >> 1].x, Ms[j >> 1].y);
                    acc_s[j] = __ffma2_rn(g, __fmul2_rn(mr, bc(mcs)), acc_s[j]);
                }
                if (SOFTOR) {
                    if (MASK_O) {
                        const float2 mr = (j & 1) ? make_float2(Mo[j >> 1].z, Mo[j >> 1].w) : make_float2(Mo[j >> 1].x, Mo[j >> 1].y);
                        acc_p[j] = __ffma2_rn(__fmul2_rn(g, __fmul2_rn(mr, bc(mco))), acc_p[j], acc_p[j]);   // p -= p * g * m
                    } else {
                        acc_p[j] = __ffma2_rn(neg2(g), acc_p[j], acc_p[j]);                                  // p *= (1 - g)
                    }
                }
            }
        }
    }
    // epilogue: every texel of the tile is written exactly once
    float rs[WROWS], ro[WROWS];
#pragma unroll
    for (int j = 0; j < WROWS / 2; ++j) {
        rs[2 * j] = acc_s[j].x; rs[2 * j + 1] = acc_s[j].y;
        ro[2 * j] = 1.f - acc_p[j].x; ro[2 * j + 1] = 1.f - acc_p[j].y;
    }
    const size_t frame = (size_t)q.ts0 * q.ts1;
    if (SOFTOR && c < q.ts0) {
        float* o = q.out_softor + (size_t)b * frame + c;
#pragma unroll
        for (int i = 0; i < WROWS; ++i)
            if (wr0 + i < q.ts1) o[(size_t)(wr0 + i) * q.ts0] = ro[i];
    }
    if (SUM && c < q.ts0) {
        if (!SUM_T) {
            float* o = q.out_sum + (size_t)b * frame + c;
#pragma unroll
            for (int i = 0; i < WROWS; ++i)
                if (wr0 + i < q.ts1) o[(size_t)(wr0 + i) * q.ts0] = rs[i];
        } else {
            float* o = q.out_sum + (size_t)b * frame + (size_t)c * q.ts1 + wr0;
            if ((q.ts1 & 3) == 0 && wr0 + WROWS <= q.ts1) {
                reinterpret_cast<float4*>(o)[0] = make_float4(rs[0], rs[1], rs[2], rs[3]);
                reinterpret_cast<float4*>(o)[1] = make_float4(rs[4], rs[5], rs[6], rs[7]);
            } else {
#pragma unroll
                for (int i = 0; i < WROWS; ++i)
                    if (wr0 + i < q.ts1) o[i] = rs[i];
            }
        }
    }
}

// Backward.  Pass 1 rebuilds the soft-OR product per texel (factors clamped at 2^-24 so an exact zero -- a point
// sitting on a texel centre -- keeps the quotient below finite: prod/om_n then equals the exclusive product for
// that point and ~6e-8 (instead of 0) for the others), caching g in shared memory; pass 2 forms
// dL/dg = gS*m_s + gO*m_o*prod/om, weights it with dg/dP = 4 g d2 (c-P)/sigma^2 and reduces per point.
template <bool SUM, bool SOFTOR, bool SUM_T, bool MASK_O>
__global__ void __launch_bounds__(CTA) splat_bwd_fast(RasterParams q, FastConsts fc) {
    __shared__ Stage<true, MASK_O> st;
    __shared__ float dp_s[FCH][2];
    extern __shared__ float2 gcache2[];     // SOFTOR: [8 warps][KCACHE][4 pairs][32 lanes]
    const int tile = blockIdx.x % q.T, b = blockIdx.x / q.T;
    const int bin = q.shared_pattern ? 0 : b;
    const int tx = tile % q.tgx, ty = tile / q.tgx;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wrow = (warp >> 1) * WROWS;
    const int wc0 = tx * TW + (warp & 1) * 32, wr0 = ty * TH + wrow;
    const int c = wc0 + lane;
    const float cf = (float)c;
    const int* toff = q.tile_off + (size_t)bin * (q.T + 1);
    const int beg = toff[tile], end = toff[tile + 1];
    if (beg == end) return;
    const int* list = q.list + (size_t)bin * q.cap;
    const PointRec* recs = q.recs + (size_t)bin * q.N;
    const size_t frame = (size_t)q.ts0 * q.ts1;
    float2* gc = gcache2 + (SOFTOR ? (size_t)warp * KCACHE * (WROWS / 2) * 32 + lane : 0);
    constexpr float OM_MIN = 5.9604645e-8f;      // 2^-24

    // upstream gradients of this lane's 8 texels, as row pairs
    float2 gs[WROWS / 2], go[WROWS / 2];
    {
        float a[WROWS], o[WROWS];
#pragma unroll
        for (int i = 0; i < WROWS; ++i) { a[i] = 0.f; o[i] = 0.f; }
        if (c < q.ts0) {
            if (SOFTOR) {
                const float* p = q.g_softor + (size_t)b * frame + c;
#pragma unroll
                for (int i = 0; i < WROWS; ++i)
                    if (wr0 + i < q.ts1) o[i] = __ldg(p + (size_t)(wr0 + i) * q.ts0);
            }
            if (SUM) {
                if (!SUM_T) {
                    const float* p = q.g_sum + (size_t)b * frame + c;
#pragma unroll
                    for (int i = 0; i < WROWS; ++i)
                        if (wr0 + i < q.ts1) a[i] = __ldg(p + (size_t)(wr0 + i) * q.ts0);
                } else {
                    const float* p = q.g_sum + (size_t)b * frame + (size_t)c * q.ts1 + wr0;
                    if ((q.ts1 & 3) == 0 && wr0 + WROWS <= q.ts1) {
                        const float4 x = __ldg(reinterpret_cast<const float4*>(p));
                        const float4 y = __ldg(reinterpret_cast<const float4*>(p) + 1);
                        a[0] = x.x; a[1] = x.y; a[2] = x.z; a[3] = x.w; a[4] = y.x; a[5] = y.y; a[6] = y.z; a[7] = y.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < WROWS; ++i)
                            if (wr0 + i < q.ts1) a[i] = __ldg(p + i);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < WROWS / 2; ++j) { gs[j] = make_float2(a[2 * j], a[2 * j + 1]); go[j] = make_float2(o[2 * j], o[2 * j + 1]); }
    }

    float2 prod[WROWS / 2];
#pragma unroll
    for (int j = 0; j < WROWS / 2; ++j) prod[j] = bc(1.f);
    const bool single = end - beg <= FCH;

    // ---- pass 1 ----
    if (SOFTOR) {
        int kk = 0;
        for (int base = beg; base < end; base += FCH) {
            const int n = min(FCH, end - base);
            stage_chunk(st, recs, list, base, n, ty * TH, fc, tid);
            unsigned m = cull_chunk(st, n, lane, wc0, wr0);
            while (m) {
                const int k = __ffs(m) - 1;
                m &= m - 1;
                const float4 pf = st.pf[k];
                const float dx = cf - pf.x;
                const float dx2 = __fmul_rn(dx, dx);
                const float mco = MASK_O ? -cheb_mask(cf - pf.z, fc.thr_o) : 0.f;
                const float4* d2p = reinterpret_cast<const float4*>(&st.dy2[k][wrow]);
                const float4* mop = reinterpret_cast<const float4*>(&st.mo[MASK_O ? k : 0][MASK_O ? wrow : 0]);
                float4 D[2] = {d2p[0], d2p[1]}, Mo[2];
                if (MASK_O) { Mo[0] = mop[0]; Mo[1] = mop[1]; }
#pragma unroll
                for (int j = 0; j < WROWS / 2; ++j) {
                    const float2 dy2 = (j & 1) ? make_float2(D[j >> 1].z, D[j >> 1].w) : make_float2(D[j >> 1].x, D[j >> 1].y);
                    const float2 d2 = __fadd2_rn(bc(dx2), dy2);
                    const float2 t = __fmul2_rn(__fmul2_rn(d2, d2), bc(fc.K2));
                    const float2 g = make_float2(ex2_approx(t.x), ex2_approx(t.y));
                    if (kk < KCACHE) gc[(kk * (WROWS / 2) + j) * 32] = g;
                    float2 om;
                    if (MASK_O) {
                        const float2 mr = (j & 1) ? make_float2(Mo[j >> 1].z, Mo[j >> 1].w) : make_float2(Mo[j >> 1].x, Mo[j >> 1].y);
                        om = __ffma2_rn(g, __fmul2_rn(mr, bc(mco)), bc(1.f));
                    } else {
                        om = __fadd2_rn(bc(1.f), neg2(g));
                    }
                    om.x = fmaxf(om.x, OM_MIN); om.y = fmaxf(om.y, OM_MIN);
                    prod[j] = __fmul2_rn(prod[j], om);
                }
                ++kk;
            }
        }
    }

    // ---- pass 2 ----
    const float inv_s2 = q.rcp_sigma * q.rcp_sigma;
    const float k0 = 4.f * (float)q.ts0 * inv_s2, k1 = 4.f * (float)q.ts1 * inv_s2;
    int kk = 0;
    for (int base = beg; base < end; base += FCH) {
        const int n = min(FCH, end - base);
        if (!(SOFTOR && single)) stage_chunk(st, recs, list, base, n, ty * TH, fc, tid);
        if (tid < n) { dp_s[tid][0] = 0.f; dp_s[tid][1] = 0.f; }
        __syncthreads();
        unsigned m = cull_chunk(st, n, lane, wc0, wr0);
        while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            const float4 pf = st.pf[k];
            const float dx = cf - pf.x;
            const float dx2 = __fmul_rn(dx, dx);
            const float ex = cf - pf.z;
            const float mcs = cheb_mask(ex, fc.thr_s);
            const float mco = MASK_O ? cheb_mask(ex, fc.thr_o) : 1.f;
            const float4* d2p = reinterpret_cast<const float4*>(&st.dy2[k][wrow]);
            const float4* dyp = reinterpret_cast<const float4*>(&st.dy[k][wrow]);
            const float4* msp = reinterpret_cast<const float4*>(&st.ms[k][wrow]);
            const float4* mop = reinterpret_cast<const float4*>(&st.mo[MASK_O ? k : 0][MASK_O ? wrow : 0]);
            float4 D[2] = {d2p[0], d2p[1]}, Y[2] = {dyp[0], dyp[1]}, Ms[2], Mo[2];
            if (SUM) { Ms[0] = msp[0]; Ms[1] = msp[1]; }
            if (SOFTOR && MASK_O) { Mo[0] = mop[0]; Mo[1] = mop[1]; }
            float2 a0 = bc(0.f), a1 = bc(0.f);
#pragma unroll
            for (int j = 0; j < WROWS / 2; ++j) {
                const float2 dy2 = (j & 1) ? make_float2(D[j >> 1].z, D[j >> 1].w) : make_float2(D[j >> 1].x, D[j >> 1].y);
                const float2 dy = (j & 1) ? make_float2(Y[j >> 1].z, Y[j >> 1].w) : make_float2(Y[j >> 1].x, Y[j >> 1].y);
                const float2 d2 = __fadd2_rn(bc(dx2), dy2);
                float2 g;
                if (SOFTOR && kk < KCACHE) g = gc[(kk * (WROWS / 2) + j) * 32];
                else {
                    const float2 t = __fmul2_rn(__fmul2_rn(d2, d2), bc(fc.K2));
                    g = make_float2(ex2_approx(t.x), ex2_approx(t.y));
                }
                float2 coef = bc(0.f);
                if (SOFTOR) {
                    float2 om, mm = bc(1.f);
                    if (MASK_O) {
                        const float2 mr = (j & 1) ? make_float2(Mo[j >> 1].z, Mo[j >> 1].w) : make_float2(Mo[j >> 1].x, Mo[j >> 1].y);
                        mm = __fmul2_rn(mr, bc(mco));
                        om = __ffma2_rn(neg2(g), mm, bc(1.f));
                    } else {
                        om = __fadd2_rn(bc(1.f), neg2(g));
                    }
                    const float2 r = make_float2(rcp_approx(fmaxf(om.x, OM_MIN)), rcp_approx(fmaxf(om.y, OM_MIN)));
                    coef = __fmul2_rn(go[j], __fmul2_rn(prod[j], r));          // gO * prod_{m != n}(1 - g_m)
                    if (MASK_O) coef = __fmul2_rn(coef, mm);
                }
                if (SUM) {
                    const float2 mr = (j & 1) ? make_float2(Ms[j >> 1].z, Ms[j >> 1].w) : make_float2(Ms[j >> 1].x, Ms[j >> 1].y);
                    coef = __ffma2_rn(gs[j], __fmul2_rn(mr, bc(mcs)), coef);
                }
                const float2 w = __fmul2_rn(__fmul2_rn(coef, g), d2);
                a0 = __ffma2_rn(w, bc(dx), a0);
                a1 = __ffma2_rn(w, dy, a1);
            }
            const float s0 = warp_sum(a0.x + a0.y), s1 = warp_sum(a1.x + a1.y);
            if (lane == 0) { atomicAdd(&dp_s[k][0], s0); atomicAdd(&dp_s[k][1], s1); }
            ++kk;
        }
        __syncthreads();
        if (tid < n) {
            float* o = q.d_pts + ((size_t)b * q.N + st.idx[tid]) * 2;
            const float v0 = dp_s[tid][0] * k0, v1 = dp_s[tid][1] * k1;
            if (v0 != 0.f) atomicAdd(o, v0);
            if (v1 != 0.f) atomicAdd(o + 1, v1);
        }
    }
}
