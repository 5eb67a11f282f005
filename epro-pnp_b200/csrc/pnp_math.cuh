// pnp_math.cuh -- scalar building blocks of the EPro-PnP hot path, shared by the sm_100a kernels
// (pnp_kernels.cu) and by a host-side build of the same functions that the CPU tests drive
// (tests/host_emul.cpp) so formulas can be checked against the golden vectors without a GPU.
//
// Everything is fp32 and allocation-free; matrices are tiny fixed-size arrays with static indexing
// so they live in registers.  Reference behaviour (file:line under /root/reference/epropnp/) is
// cited per function; the code itself is written from the math, not translated.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "epropnp_b200.h"

#if defined(__CUDACC__)
#define PNP_HD __host__ __device__ __forceinline__
// once-per-object fp64 routines: kept out of line so they do not inflate the register allocation of
// the per-point loops they are called next to
#define PNP_HD_COLD __host__ __device__ __noinline__
#else
#define PNP_HD inline
#define PNP_HD_COLD inline
#endif

namespace pnp {

// ------------------------------------------------------------------------------------------------
// Solver hyper-parameters: the C ABI's POD (include/epropnp_b200.h), passed by value to the kernels.
using Params = ::EpnpParams;

struct Cam {
    float k[9];                 // row-major 3x3
    float lbx, lby, ubx, uby;
    float z_min;
    int   bounded;
};

template <int DOF> struct Dim {
    static constexpr int POSE = (DOF == 6) ? 7 : 4;
    static constexpr int NA = DOF * (DOF + 1) / 2;     // packed upper triangle of J^T J
    static constexpr int NV = NA + DOF + 1;            // + J^T r + cost
};

// ------------------------------------------------------------------------------------------------
// Two-lane fp32 value.  On sm_100a each op below is ONE packed instruction (FFMA2 / FMUL2 / FADD2); the host
// build evaluates the lanes with scalar fmaf so the packed formulas can be checked on the CPU.
#if defined(__CUDACC__) || defined(EPNP_SIMT_EMUL)
typedef float2 V2;
#else
struct alignas(8) V2 { float x, y; };
#endif
PNP_HD V2 v2(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
PNP_HD V2 v2splat(float x) { return v2(x, x); }
PNP_HD V2 v2fma(V2 a, V2 b, V2 c) {
#if defined(__CUDA_ARCH__)
    return __ffma2_rn(a, b, c);
#else
    return v2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
PNP_HD V2 v2mul(V2 a, V2 b) {
#if defined(__CUDA_ARCH__)
    return __fmul2_rn(a, b);
#else
    return v2(a.x * b.x, a.y * b.y);
#endif
}
PNP_HD V2 v2add(V2 a, V2 b) {
#if defined(__CUDA_ARCH__)
    return __fadd2_rn(a, b);
#else
    return v2(a.x + b.x, a.y + b.y);
#endif
}
PNP_HD V2 v2neg(V2 a) { return v2(-a.x, -a.y); }

// packed upper-triangular index, row-major: (i, j>=i)
PNP_HD constexpr int tri(int i, int j, int n) { return i * n - i * (i - 1) / 2 + (j - i); }

// ------------------------------------------------------------------------------------------------
// rotations  (common.py:22-64)
PNP_HD void quat_to_rot(const float* q, float* R) {
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    const float ww = w * w, xx = x * x, yy = y * y, zz = z * z;
    R[0] = ww + xx - yy - zz; R[1] = 2.f * (x * y - w * z); R[2] = 2.f * (x * z + w * y);
    R[3] = 2.f * (x * y + w * z); R[4] = ww - xx + yy - zz; R[5] = 2.f * (y * z - w * x);
    R[6] = 2.f * (x * z - w * y); R[7] = 2.f * (y * z + w * x); R[8] = ww - xx - yy + zz;
}

PNP_HD void yaw_to_rot(float yaw, float* R) {
    float s, c;
#if defined(__CUDA_ARCH__)
    sincosf(yaw, &s, &c);
#else
    s = sinf(yaw); c = cosf(yaw);
#endif
    R[0] = c;  R[1] = 0.f; R[2] = s;
    R[3] = 0.f; R[4] = 1.f; R[5] = 0.f;
    R[6] = -s; R[7] = 0.f; R[8] = c;
}

template <int DOF> PNP_HD void pose_to_rot(const float* pose, float* R) {
    if (DOF == 6) quat_to_rot(pose + 3, R); else yaw_to_rot(pose[3], R);
}

// P = K [R | t] as 3 rows of 4  (the pre-multiplied form of camera.py:21-30 `project_b`)
PNP_HD void make_proj(const float* K, const float* R, const float* t, float* P) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
            P[r * 4 + c] = K[r * 3 + 0] * R[0 + c] + K[r * 3 + 1] * R[3 + c] + K[r * 3 + 2] * R[6 + c];
        P[r * 4 + 3] = K[r * 3 + 0] * t[0] + K[r * 3 + 1] * t[1] + K[r * 3 + 2] * t[2];
    }
}

// pose (+) step  (levenberg_marquardt.py:255-265, camera.py:145-165): translation adds; the
// rotation increment d lives in the tangent space, q <- normalize(q + T(q) d).
template <int DOF> PNP_HD void pose_add(const float* pose, const float* step, float* out) {
    out[0] = pose[0] + step[0]; out[1] = pose[1] + step[1]; out[2] = pose[2] + step[2];
    if (DOF == 4) { out[3] = pose[3] + step[3]; return; }
    const float w = pose[3], x = pose[4], y = pose[5], z = pose[6];
    const float a = step[3], b = step[4], c = step[5];
    float qw = w + (x * a + y * b + z * c);
    float qx = x + (-w * a - z * b + y * c);
    float qy = y + (z * a - w * b - x * c);
    float qz = z + (-y * a + x * b - w * c);
    const float n = fmaxf(sqrtf(qw * qw + qx * qx + qy * qy + qz * qz), 1e-12f);
    const float inv = 1.0f / n;
    out[3] = qw * inv; out[4] = qx * inv; out[5] = qy * inv; out[6] = qz * inv;
}

// ------------------------------------------------------------------------------------------------
// Cost of one correspondence under a pre-multiplied projection (AMIS inner loop; cost-only
// evaluate_pnp).  camera.py:21-30 + :81-93 clamp, cost_fun.py:8-12,52-59.
// `rcp`: 1/x functor so the device build can use the approximate MUFU reciprocal.
struct ExactRcp { PNP_HD float operator()(float x) const { return 1.0f / x; } };
struct ExactSqrt { PNP_HD float operator()(float x) const { return sqrtf(x); } };

template <bool BOUNDED, class Rcp, class Sqrt>
PNP_HD float point_cost(const float* P, const Cam& c, float delta, float half_d2,
                        float X, float Y, float Z, float u, float v, float wu, float wv,
                        Rcp rcp, Sqrt sq) {
    float xh = fmaf(P[0], X, fmaf(P[1], Y, fmaf(P[2], Z, P[3])));
    float yh = fmaf(P[4], X, fmaf(P[5], Y, fmaf(P[6], Z, P[7])));
    float zh = fmaf(P[8], X, fmaf(P[9], Y, fmaf(P[10], Z, P[11])));
    float iz = rcp(fmaxf(zh, c.z_min));
    float px = xh * iz, py = yh * iz;
    if (BOUNDED) {
        px = fminf(fmaxf(px, c.lbx), c.ubx);
        py = fminf(fmaxf(py, c.lby), c.uby);
    }
    float rx = (px - u) * wu, ry = (py - v) * wv;
    float s2 = fmaf(rx, rx, ry * ry);
    float s = sq(s2);
    return (s <= delta) ? 0.5f * s2 : fmaf(delta, s, -half_d2);
}

// ------------------------------------------------------------------------------------------------
// Reverse mode of point_cost: given g = dL/d(cost of this pose), accumulate dL/d(X,Y,Z,u,v,wu,wv) of one
// correspondence and dL/d(delta).  Differentiates exactly what the reference's autograd sees on the cost
// path (camera.py:21-30 project_b, :81-93 clamps -> zero gradient where clamped, z.clamp(min) -> zero where
// zh < z_min; cost_fun.py:8-12 Huber):
//   d cost / d r = r * min(1, delta / s)          d cost / d delta = (s - delta) for outliers, 0 for inliers
// grad[0..2] += dL/dX, grad[3..4] += dL/d(u,v), grad[5..6] += dL/d(wu,wv); returns dL/d delta contribution.
template <bool BOUNDED, class Rcp>
PNP_HD float point_cost_backward(const float* P, const Cam& c, float delta, float g,
                                 float X, float Y, float Z, float u, float v, float wu, float wv,
                                 float* grad, Rcp rcp) {
    const float xh = fmaf(P[0], X, fmaf(P[1], Y, fmaf(P[2], Z, P[3])));
    const float yh = fmaf(P[4], X, fmaf(P[5], Y, fmaf(P[6], Z, P[7])));
    const float zh = fmaf(P[8], X, fmaf(P[9], Y, fmaf(P[10], Z, P[11])));
    const float iz = rcp(fmaxf(zh, c.z_min));
    const float px = xh * iz, py = yh * iz;
    float pxc = px, pyc = py;
    if (BOUNDED) {
        pxc = fminf(fmaxf(px, c.lbx), c.ubx);
        pyc = fminf(fmaxf(py, c.lby), c.uby);
    }
    const float ex = pxc - u, ey = pyc - v;
    const float rx = ex * wu, ry = ey * wv;
    const float s2 = fmaf(rx, rx, ry * ry);
    const float s = sqrtf(s2);
    const bool inlier = (s <= delta);
    const float k = inlier ? g : g * (delta / s);
    const float grx = k * rx, gry = k * ry;
    grad[5] = fmaf(grx, ex, grad[5]);
    grad[6] = fmaf(gry, ey, grad[6]);
    const float gex = grx * wu, gey = gry * wv;          // dL/d(pxc - u), dL/d(pyc - v)
    grad[3] -= gex;
    grad[4] -= gey;
    const float gpx = (!BOUNDED || pxc == px) ? gex : 0.f;
    const float gpy = (!BOUNDED || pyc == py) ? gey : 0.f;
    const float gxh = gpx * iz, gyh = gpy * iz;
    const float gzh = (zh >= c.z_min) ? -(gxh * px + gyh * py) : 0.f;
    grad[0] = fmaf(P[0], gxh, fmaf(P[4], gyh, fmaf(P[8], gzh, grad[0])));
    grad[1] = fmaf(P[1], gxh, fmaf(P[5], gyh, fmaf(P[9], gzh, grad[1])));
    grad[2] = fmaf(P[2], gxh, fmaf(P[6], gyh, fmaf(P[10], gzh, grad[2])));
    return inlier ? 0.f : g * (s - delta);
}

// ------------------------------------------------------------------------------------------------
// One correspondence's contribution to the Gauss-Newton normal equations:
//   acc[0..NA)       upper triangle of J~^T J~
//   acc[NA..NA+DOF)  J~^T r~
//   acc[NA+DOF]      sum of Huber costs
// camera.py:10-18 (project_a), :81-105 (clamp + clip mask), :111-143 (2xDOF Jacobian);
// cost_fun.py:52-84 (weighted residual, Huber, robust rescale of residual and Jacobian).
template <int DOF, bool CLIP>
PNP_HD void point_normal_eq(const float* R, const float* t, const Cam& c, float delta, float huber_eps,
                            float X, float Y, float Z, float u, float v, float wu, float wv,
                            float* acc) {
    constexpr int NA = Dim<DOF>::NA;
    const float xr = fmaf(R[0], X, fmaf(R[1], Y, R[2] * Z));
    const float yr = fmaf(R[3], X, fmaf(R[4], Y, R[5] * Z));
    const float zr = fmaf(R[6], X, fmaf(R[7], Y, R[8] * Z));
    const float xc = xr + t[0], yc = yr + t[1], zc = zr + t[2];
    const float* K = c.k;
    const float xh = fmaf(K[0], xc, fmaf(K[1], yc, K[2] * zc));
    const float yh = fmaf(K[3], xc, fmaf(K[4], yc, K[5] * zc));
    const float zh = fmaf(K[6], xc, fmaf(K[7], yc, K[8] * zc));
    const float z = fmaxf(zh, c.z_min);
    const float iz = 1.0f / z;
    float px = xh * iz, py = yh * iz;
    if (c.bounded) {
        px = fminf(fmaxf(px, c.lbx), c.ubx);
        py = fminf(fmaxf(py, c.lby), c.uby);
    }
    float rx = (px - u) * wu, ry = (py - v) * wv;
    const float s2 = fmaf(rx, rx, ry * ry);
    const float s = sqrtf(s2);
    acc[NA + DOF] += (s <= delta) ? 0.5f * s2 : fmaf(delta, s, -0.5f * delta * delta);
    // sqrt(rho') = sqrt(min(delta / max(s, eps), 1)): exactly 1 for inliers (s <= delta), which is most points
    // near the optimum -- skip the divide + sqrt there (warp-uniform most of the time)
    const float sc = (s <= delta && delta >= huber_eps) ? 1.0f : sqrtf(fminf(delta / fmaxf(s, huber_eps), 1.0f));
    rx *= sc; ry *= sc;

    float jx[DOF], jy[DOF];
    jx[0] = K[0] * iz; jx[1] = K[1] * iz; jx[2] = (K[2] - px) * iz;
    jy[0] = K[3] * iz; jy[1] = K[4] * iz; jy[2] = (K[5] - py) * iz;
    if (DOF == 6) {
        const float ax = 2.f * xr, ay = 2.f * yr, az = 2.f * zr;          // J3 * skew(2 x_rot)
        jx[3] = jx[1] * az - jx[2] * ay; jx[4] = jx[2] * ax - jx[0] * az; jx[5] = jx[0] * ay - jx[1] * ax;
        jy[3] = jy[1] * az - jy[2] * ay; jy[4] = jy[2] * ax - jy[0] * az; jy[5] = jy[0] * ay - jy[1] * ax;
    } else {
        jx[3] = jx[0] * zr - jx[2] * xr;                                   // d/d yaw
        jy[3] = jy[0] * zr - jy[2] * xr;
    }
    float sx = wu * sc, sy = wv * sc;
    if (CLIP) {           // rows whose projection sits on a clamp carry no gradient
        const bool cz = (z == c.z_min);
        const bool cx = cz || (c.bounded && (px == c.lbx || px == c.ubx));
        const bool cy = cz || (c.bounded && (py == c.lby || py == c.uby));
        sx = cx ? 0.f : sx;
        sy = cy ? 0.f : sy;
    }
#pragma unroll
    for (int i = 0; i < DOF; ++i) { jx[i] *= sx; jy[i] *= sy; }
    int idx = 0;
#pragma unroll
    for (int i = 0; i < DOF; ++i) {
#pragma unroll
        for (int j = i; j < DOF; ++j) { acc[idx] = fmaf(jx[i], jx[j], fmaf(jy[i], jy[j], acc[idx])); ++idx; }
    }
#pragma unroll
    for (int i = 0; i < DOF; ++i) acc[NA + i] = fmaf(jx[i], rx, fmaf(jy[i], ry, acc[NA + i]));
}

// 1/x and sqrt(x) without the slow-path branches of the IEEE forms: special-function seed plus one Newton step
// (x is a clamped depth >= z_min, resp. a sum of squares; neither needs denormal or negative handling).
PNP_HD float rcp_newton(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
#else
    return 1.0f / x;
#endif
}
PNP_HD float sqrt_newton(float x) {
#if defined(__CUDA_ARCH__)
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(fmaxf(x, 1e-30f)));
    const float s = x * y;
    return fmaf(0.5f * y, fmaf(-s, s, x), s);
#else
    return sqrtf(x);
#endif
}

// Row-packed form of point_normal_eq (the LM kernel's evaluation): lane x carries the u-row of the 2xDOF
// Jacobian, lane y the v-row, so the DOF(DOF+1)/2 + DOF accumulations are one FFMA2 each instead of two FFMA.
//   acc2[k].x + acc2[k].y == acc[k] of point_normal_eq (k < NA + DOF), `cost` == acc[NA + DOF]
// kuv[i] = (K[0][i], K[1][i]); `nu`, `nv` are the NEGATED observations as the staged pair records hold them.
template <int DOF, bool CLIP>
PNP_HD void point_normal_eq_rows(const float* R, const float* t, const Cam& c, const V2* kuv, float delta,
                                 float huber_eps, float X, float Y, float Z, float nu, float nv, float wu,
                                 float wv, V2* acc2, float& cost) {
    constexpr int NA = Dim<DOF>::NA;
    const float xr = fmaf(R[0], X, fmaf(R[1], Y, R[2] * Z));
    const float yr = fmaf(R[3], X, fmaf(R[4], Y, R[5] * Z));
    const float zr = fmaf(R[6], X, fmaf(R[7], Y, R[8] * Z));
    const float xc = xr + t[0], yc = yr + t[1], zc = zr + t[2];
    const V2 h = v2fma(kuv[0], v2splat(xc), v2fma(kuv[1], v2splat(yc), v2mul(kuv[2], v2splat(zc))));
    const float zh = fmaf(c.k[6], xc, fmaf(c.k[7], yc, c.k[8] * zc));
    const float z = fmaxf(zh, c.z_min);
    const V2 iz = v2splat(rcp_newton(z));
    V2 p = v2mul(h, iz);
    if (c.bounded) {
        p.x = fminf(fmaxf(p.x, c.lbx), c.ubx);
        p.y = fminf(fmaxf(p.y, c.lby), c.uby);
    }
    const V2 w = v2(wu, wv);
    V2 r = v2mul(v2add(p, v2(nu, nv)), w);
    const float s2 = fmaf(r.x, r.x, r.y * r.y);
    const float s = sqrt_newton(s2);
    cost += (s <= delta) ? 0.5f * s2 : fmaf(delta, s, -0.5f * delta * delta);
    const float sc = (s <= delta && delta >= huber_eps) ? 1.0f : sqrt_newton(fminf(delta * rcp_newton(fmaxf(s, huber_eps)), 1.0f));
    const V2 sc2 = v2splat(sc);
    r = v2mul(r, sc2);

    V2 j[DOF];
    j[0] = v2mul(kuv[0], iz); j[1] = v2mul(kuv[1], iz); j[2] = v2mul(v2add(kuv[2], v2neg(p)), iz);
    if (DOF == 6) {
        const V2 ax = v2splat(2.f * xr), ay = v2splat(2.f * yr), az = v2splat(2.f * zr);
        j[3] = v2fma(j[1], az, v2neg(v2mul(j[2], ay)));
        j[4] = v2fma(j[2], ax, v2neg(v2mul(j[0], az)));
        j[5] = v2fma(j[0], ay, v2neg(v2mul(j[1], ax)));
    } else {
        j[3] = v2fma(j[0], v2splat(zr), v2neg(v2mul(j[2], v2splat(xr))));
    }
    V2 sw = v2mul(w, sc2);
    if (CLIP) {
        const bool cz = (z == c.z_min);
        const bool cx = cz || (c.bounded && (p.x == c.lbx || p.x == c.ubx));
        const bool cy = cz || (c.bounded && (p.y == c.lby || p.y == c.uby));
        sw.x = cx ? 0.f : sw.x;
        sw.y = cy ? 0.f : sw.y;
    }
#pragma unroll
    for (int i = 0; i < DOF; ++i) j[i] = v2mul(j[i], sw);
    int idx = 0;
#pragma unroll
    for (int i = 0; i < DOF; ++i) {
#pragma unroll
        for (int k = i; k < DOF; ++k) { acc2[idx] = v2fma(j[i], j[k], acc2[idx]); ++idx; }
    }
#pragma unroll
    for (int i = 0; i < DOF; ++i) acc2[NA + i] = v2fma(j[i], r, acc2[NA + i]);
}

// The same per-point quantities written out instead of reduced (evaluate_pnp with out_residual /
// out_jacobian, common.py:67-100): res[2], jac[2*DOF] (row x then row y), returns the Huber cost.
template <int DOF>
PNP_HD float point_residual_jac(const float* R, const float* t, const Cam& c, float delta, float huber_eps,
                                bool clip, float X, float Y, float Z, float u, float v, float wu, float wv,
                                float* res, float* jac) {
    const float xr = fmaf(R[0], X, fmaf(R[1], Y, R[2] * Z));
    const float yr = fmaf(R[3], X, fmaf(R[4], Y, R[5] * Z));
    const float zr = fmaf(R[6], X, fmaf(R[7], Y, R[8] * Z));
    const float xc = xr + t[0], yc = yr + t[1], zc = zr + t[2];
    const float* K = c.k;
    const float xh = fmaf(K[0], xc, fmaf(K[1], yc, K[2] * zc));
    const float yh = fmaf(K[3], xc, fmaf(K[4], yc, K[5] * zc));
    const float zh = fmaf(K[6], xc, fmaf(K[7], yc, K[8] * zc));
    const float z = fmaxf(zh, c.z_min);
    const float iz = 1.0f / z;
    float px = xh * iz, py = yh * iz;
    if (c.bounded) {
        px = fminf(fmaxf(px, c.lbx), c.ubx);
        py = fminf(fmaxf(py, c.lby), c.uby);
    }
    float rx = (px - u) * wu, ry = (py - v) * wv;
    const float s2 = fmaf(rx, rx, ry * ry);
    const float s = sqrtf(s2);
    const float cost = (s <= delta) ? 0.5f * s2 : fmaf(delta, s, -0.5f * delta * delta);
    const float sc = (s <= delta && delta >= huber_eps) ? 1.0f : sqrtf(fminf(delta / fmaxf(s, huber_eps), 1.0f));
    res[0] = rx * sc; res[1] = ry * sc;
    float* jx = jac; float* jy = jac + DOF;
    jx[0] = K[0] * iz; jx[1] = K[1] * iz; jx[2] = (K[2] - px) * iz;
    jy[0] = K[3] * iz; jy[1] = K[4] * iz; jy[2] = (K[5] - py) * iz;
    if (DOF == 6) {
        const float ax = 2.f * xr, ay = 2.f * yr, az = 2.f * zr;
        jx[3] = jx[1] * az - jx[2] * ay; jx[4] = jx[2] * ax - jx[0] * az; jx[5] = jx[0] * ay - jx[1] * ax;
        jy[3] = jy[1] * az - jy[2] * ay; jy[4] = jy[2] * ax - jy[0] * az; jy[5] = jy[0] * ay - jy[1] * ax;
    } else {
        jx[3] = jx[0] * zr - jx[2] * xr;
        jy[3] = jy[0] * zr - jy[2] * xr;
    }
    float sx = wu * sc, sy = wv * sc;
    if (clip) {
        const bool cz = (z == c.z_min);
        if (cz || (c.bounded && (px == c.lbx || px == c.ubx))) sx = 0.f;
        if (cz || (c.bounded && (py == c.lby || py == c.uby))) sy = 0.f;
    }
#pragma unroll
    for (int i = 0; i < DOF; ++i) { jx[i] *= sx; jy[i] *= sy; }
    return cost;
}

// ------------------------------------------------------------------------------------------------
// Backward of the differentiable Gauss-Newton step (LMSolver.gn_step evaluated with autograd on,
// levenberg_marquardt.py:243-253 through camera.py:119-129 and cost_fun.py:52-84).
//   step = -(H + eps I)^-1 g,   H = sum_n J~_n^T J~_n,   g = sum_n J~_n^T r~_n         (pose detached)
// With sbar = dL/dstep and v = -(H + eps I)^-1 sbar:   dL = sum_n d[ b_n . (J~_n v) + a_n . (r~_n + J~_n step) ]
// where a_n = J~_n v and b_n = r~_n + J~_n step are held constant.  The per-point scalar is differentiated in
// forward mode: Dual<8> carries the tangents w.r.t. (X, Y, Z, u, v, wu, wv, delta) through the same formulas as
// point_residual_jac, so clamps, the Huber rescale and the clip mask differentiate exactly as torch's autograd does.
template <int ND> struct Dual {
    float v;
    float d[ND];
};
template <int ND> PNP_HD Dual<ND> dual_const(float c) {
    Dual<ND> r; r.v = c;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = 0.f;
    return r;
}
template <int ND> PNP_HD Dual<ND> dual_seed(float c, int k) { Dual<ND> r = dual_const<ND>(c); r.d[k] = 1.f; return r; }
template <int ND> PNP_HD Dual<ND> operator+(const Dual<ND>& a, const Dual<ND>& b) {
    Dual<ND> r; r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
template <int ND> PNP_HD Dual<ND> operator-(const Dual<ND>& a, const Dual<ND>& b) {
    Dual<ND> r; r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
template <int ND> PNP_HD Dual<ND> operator*(const Dual<ND>& a, const Dual<ND>& b) {
    Dual<ND> r; r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = fmaf(a.v, b.d[i], a.d[i] * b.v);
    return r;
}
template <int ND> PNP_HD Dual<ND> operator*(const Dual<ND>& a, float c) {
    Dual<ND> r; r.v = a.v * c;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] * c;
    return r;
}
template <int ND> PNP_HD Dual<ND> operator+(const Dual<ND>& a, float c) { Dual<ND> r = a; r.v += c; return r; }
template <int ND> PNP_HD Dual<ND> dual_recip(const Dual<ND>& a) {
    Dual<ND> r; r.v = 1.0f / a.v;
    const float k = -r.v * r.v;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] * k;
    return r;
}
template <int ND> PNP_HD Dual<ND> dual_sqrt(const Dual<ND>& a) {            // a.v > 0
    Dual<ND> r; r.v = sqrtf(a.v);
    const float k = 0.5f / r.v;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] * k;
    return r;
}
template <int ND> PNP_HD Dual<ND> dual_gate(const Dual<ND>& a, float value, bool pass) {   // clamp: value replaced, tangent gated
    Dual<ND> r; r.v = value;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = pass ? a.d[i] : 0.f;
    return r;
}

// One correspondence's contribution to dL/d(X, Y, Z, u, v, wu, wv, delta) (grad[0..8), accumulated with +=).
// R, t: rotation / translation of the (detached) pose; vv = v, sv = step (DOF-vectors).
template <int DOF>
PNP_HD void gn_step_point_backward(const float* R, const float* t, const Cam& c, float delta, float huber_eps,
                                   float X, float Y, float Z, float u, float v, float wu, float wv,
                                   const float* vv, const float* sv, float* grad) {
    typedef Dual<8> D;
    const D dX = dual_seed<8>(X, 0), dY = dual_seed<8>(Y, 1), dZ = dual_seed<8>(Z, 2);
    const D du = dual_seed<8>(u, 3), dv = dual_seed<8>(v, 4), dwu = dual_seed<8>(wu, 5), dwv = dual_seed<8>(wv, 6);
    const D dd = dual_seed<8>(delta, 7);
    const D xr = dX * R[0] + dY * R[1] + dZ * R[2];
    const D yr = dX * R[3] + dY * R[4] + dZ * R[5];
    const D zr = dX * R[6] + dY * R[7] + dZ * R[8];
    const D xc = xr + t[0], yc = yr + t[1], zc = zr + t[2];
    const float* K = c.k;
    const D xh = xc * K[0] + yc * K[1] + zc * K[2];
    const D yh = xc * K[3] + yc * K[4] + zc * K[5];
    const D zh = xc * K[6] + yc * K[7] + zc * K[8];
    const bool z_free = zh.v >= c.z_min;                                   // z.clamp(min=z_min): tangent passes iff zh >= z_min
    const D z = dual_gate(zh, fmaxf(zh.v, c.z_min), z_free);
    const D iz = dual_recip(z);
    D px = xh * iz, py = yh * iz;
    if (c.bounded) {
        const float cx = fminf(fmaxf(px.v, c.lbx), c.ubx), cy = fminf(fmaxf(py.v, c.lby), c.uby);
        px = dual_gate(px, cx, cx == px.v);
        py = dual_gate(py, cy, cy == py.v);
    }
    const D ex = px - du, ey = py - dv;
    D rx = ex * dwu, ry = ey * dwv;
    const D s2 = rx * rx + ry * ry;
    // sqrt(rho') = sqrt(min(delta / max(s, eps), 1)): constant 1 for inliers
    D sc = dual_const<8>(1.0f);
    const float s = sqrtf(s2.v);
    if (!(s <= delta && delta >= huber_eps)) {
        const float sm = fmaxf(s, huber_eps);
        const float ratio = delta / sm;
        if (ratio < 1.0f) {
            const D sd = (s >= huber_eps && s > 0.f) ? dual_sqrt(s2) : dual_const<8>(sm);
            sc = dual_sqrt(dd * dual_recip(sd));
        }
    }
    rx = rx * sc; ry = ry * sc;
    // clip mask (constant): rows whose projection sits on a clamp carry no Jacobian
    const bool cz = (z.v == c.z_min);
    const bool clip_x = cz || (c.bounded && (px.v == c.lbx || px.v == c.ubx));
    const bool clip_y = cz || (c.bounded && (py.v == c.lby || py.v == c.uby));
    const D sx = clip_x ? dual_const<8>(0.f) : dwu * sc;
    const D sy = clip_y ? dual_const<8>(0.f) : dwv * sc;
    // (J_cam w)_row = (K_row0 m_x + K_row1 m_y + (K_row2 - p_row) m_z) / z,  m(w) = w_t + d(x_rot)/d(rot) w_r
    D Au, Av, Bu, Bv;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
        const float* w = which == 0 ? vv : sv;
        D mx, my, mz;
        if (DOF == 6) {
            mx = (yr * w[5] - zr * w[4]) * 2.0f + w[0];
            my = (zr * w[3] - xr * w[5]) * 2.0f + w[1];
            mz = (xr * w[4] - yr * w[3]) * 2.0f + w[2];
        } else {
            mx = zr * w[3] + w[0];
            my = dual_const<8>(w[1]);
            mz = xr * (-w[3]) + w[2];
        }
        const D ju = (mx * K[0] + my * K[1] + (dual_const<8>(K[2]) - px) * mz) * iz * sx;
        const D jv = (mx * K[3] + my * K[4] + (dual_const<8>(K[5]) - py) * mz) * iz * sy;
        if (which == 0) { Au = ju; Av = jv; } else { Bu = ju; Bv = jv; }
    }
    const float au = Au.v, av = Av.v;                      // a = J~ v
    const float bu = rx.v + Bu.v, bv = ry.v + Bv.v;        // b = r~ + J~ step
#pragma unroll
    for (int k = 0; k < 8; ++k)
        grad[k] += bu * Au.d[k] + au * (rx.d[k] + Bu.d[k]) + bv * Av.d[k] + av * (ry.d[k] + Bv.d[k]);
}

// dL/dstep from dL/d(pose (+) step) (pose_add: translation plain, quaternion q' = normalize(q + T(q) step_r))
template <int DOF> PNP_HD void pose_add_backward(const float* pose, const float* step, const float* gout, float* gstep) {
    gstep[0] = gout[0]; gstep[1] = gout[1]; gstep[2] = gout[2];
    if (DOF == 4) { gstep[3] = gout[3]; return; }
    const float w = pose[3], x = pose[4], y = pose[5], z = pose[6];
    const float a = step[3], b = step[4], c = step[5];
    const float q[4] = {w + (x * a + y * b + z * c), x + (-w * a - z * b + y * c), y + (z * a - w * b - x * c),
                        z + (-y * a + x * b - w * c)};
    const float n = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), 1e-12f);
    const float inv = 1.0f / n;
    const float dot = (q[0] * gout[3] + q[1] * gout[4] + q[2] * gout[5] + q[3] * gout[6]) * inv * inv;
    float gq[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) gq[i] = (gout[3 + i] - q[i] * dot) * inv;
    // q~ = q + T step_r with T = [[x, y, z], [-w, -z, y], [z, -w, -x], [-y, x, -w]]
    gstep[3] = x * gq[0] - w * gq[1] + z * gq[2] - y * gq[3];
    gstep[4] = y * gq[0] - z * gq[1] - w * gq[2] + x * gq[3];
    gstep[5] = z * gq[0] + y * gq[1] - x * gq[2] - w * gq[3];
}

// ------------------------------------------------------------------------------------------------
// Small symmetric-positive-definite linear algebra, N <= 6, static indexing only, templated on the
// working precision.  The per-point loops are fp32; these once-per-object factorizations run in fp64
// (`Hi`): the matrices are ill-conditioned by construction (information ~1e5 next to an identity
// block, epropnp.py:297-298) and fp64 here costs nothing measurable while it removes the largest
// rounding term of the fp32 reference from our side of the parity budget.
typedef double Hi;

PNP_HD float inv_sqrt(float x) {
#if defined(__CUDA_ARCH__)
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}
PNP_HD double inv_sqrt(double x) {
#if defined(__CUDA_ARCH__)
    // MUFU.RSQ seed (2 ulp in fp32) + two Newton steps in fp64: relative error ~1e-16 with six DP
    // multiply-adds instead of the ~40-instruction IEEE rsqrt(double) sequence.  Arguments here are
    // Cholesky pivots (1e-12 .. 1e8), well inside fp32 range; <= 0 / NaN still yield NaN / inf.
    const double y0 = (double)rsqrtf((float)x);
    const double hx = 0.5 * x;
    double y = y0 * (1.5 - hx * y0 * y0);
    y = y * (1.5 - hx * y * y);
    return y;
#else
    return 1.0 / sqrt(x);
#endif
}
PNP_HD double fast_recip(double x) {
#if defined(__CUDA_ARCH__)
    // same idea for 1/x: MUFU.RCP seed + two Newton steps
    const double y0 = (double)(1.0f / (float)x);
    double y = y0 * (2.0 - x * y0);
    y = y * (2.0 - x * y);
    return y;
#else
    return 1.0 / x;
#endif
}
PNP_HD float fast_recip(float x) { return 1.0f / x; }

// a: packed upper triangle (row-major).  L: n*n row-major storage of which the LOWER triangle is written (the strict
// upper part is never read by any caller and stays unset), with the RECIPROCAL of the diagonal stored in Dinv (so the
// substitutions below multiply instead of divide).  Returns false when a pivot is not positive (LAPACK potrf's
// failure test: pivot <= 0 or NaN).
template <int N, class T> PNP_HD bool chol_packed(const T* a, T* L, T* Dinv) {
    // every loop has the constant trip count N with a compile-time-resolvable guard: loops whose bounds depend on an
    // outer index are not reliably unrolled, and then L / Dinv are indexed dynamically and live in local memory
    bool ok = true;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        T d = a[tri(j, j, N)];
#pragma unroll
        for (int k = 0; k < N; ++k)
            if (k < j) d -= L[j * N + k] * L[j * N + k];
        ok = ok && (d > T(0));
        const T inv = inv_sqrt(d);
        Dinv[j] = inv;
        L[j * N + j] = d * inv;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (i > j) {
                T v = a[tri(j, i, N)];
#pragma unroll
                for (int k = 0; k < N; ++k)
                    if (k < j) v -= L[i * N + k] * L[j * N + k];
                L[i * N + j] = v * inv;
            }
        }
    }
    return ok;
}

// x = A^-1 b through the Cholesky factor (A SPD).  NaN when A is not PD.
template <int N, class T> PNP_HD void chol_solve(const T* L, const T* Dinv, const T* b, T* x) {
    T y[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        T v = b[i];
#pragma unroll
        for (int k = 0; k < N; ++k)
            if (k < i) v -= L[i * N + k] * y[k];
        y[i] = v * Dinv[i];
    }
#pragma unroll
    for (int ii = 0; ii < N; ++ii) {
        const int i = N - 1 - ii;
        T v = y[i];
#pragma unroll
        for (int k = 0; k < N; ++k)
            if (k > i) v -= L[k * N + i] * x[k];
        x[i] = v * Dinv[i];
    }
}

// ------------------------------------------------------------------------------------------------
// Levenberg-Marquardt state machine for ONE object (levenberg_marquardt.py:154-179, 192-241).
// The Jacobian is never stored: "current" and "candidate" are the reduced normal equations.
template <int DOF> struct LMState {
    float pose[Dim<DOF>::POSE];
    float pose_new[Dim<DOF>::POSE];
    float a[Dim<DOF>::NA];      // J^T J at pose (upper triangle)
    float g[DOF];               // J^T r at pose
    float cost;
    float radius, shrink;
    float model_change;         // predicted decrease of the pending step
};

// Adopt a freshly reduced evaluation (acc layout of point_normal_eq) as the current linearisation.
template <int DOF> PNP_HD void lm_adopt(LMState<DOF>& s, const float* acc) {
    constexpr int NA = Dim<DOF>::NA;
#pragma unroll
    for (int i = 0; i < NA; ++i) s.a[i] = acc[i];
#pragma unroll
    for (int i = 0; i < DOF; ++i) s.g[i] = acc[NA + i];
    s.cost = acc[NA + DOF];
}

// step = -(A + diag(add))^-1 g and the model cost change -step^T (A step / 2 + g)
// (levenberg_marquardt.py:205-216, 225).  T = float inside the LM / GN iterations (a step only has to
// decrease the cost; the iteration corrects its rounding), so the per-iteration serial section is a
// short fp32 chain with MUFU rsqrt and no divisions.
template <int DOF, class T> PNP_HD float damped_step(const float* a, const float* g, const float* add, float* step) {
    constexpr int NA = Dim<DOF>::NA;
    T al[NA], gh[DOF], L[DOF * DOF], Dinv[DOF], st[DOF];
#pragma unroll
    for (int i = 0; i < NA; ++i) al[i] = (T)a[i];
#pragma unroll
    for (int i = 0; i < DOF; ++i) { al[tri(i, i, DOF)] += (T)add[i]; gh[i] = (T)g[i]; }
    chol_packed<DOF, T>(al, L, Dinv);
    chol_solve<DOF, T>(L, Dinv, gh, st);
    T m = 0;
#pragma unroll
    for (int i = 0; i < DOF; ++i) {
        T v = 0;
#pragma unroll
        for (int j = 0; j < DOF; ++j) v += (T)a[(i <= j) ? tri(i, j, DOF) : tri(j, i, DOF)] * (-st[j]);
        m += (-st[i]) * (v * T(0.5) + gh[i]);
    }
#pragma unroll
    for (int i = 0; i < DOF; ++i) step[i] = (float)(-st[i]);
    return (float)(-m);
}

// Same result to ~fp64 accuracy at fp32 latency: fp32 Cholesky solve, then ONE step of iterative
// refinement whose residual g - (A + D) x is formed in fp64 (6 independent length-DOF dot products, no
// fp64 sqrt / divide chain) and solved again with the fp32 factor.
template <int DOF> PNP_HD float damped_step_refined(const float* a, const float* g, const float* add, float* step) {
    constexpr int NA = Dim<DOF>::NA;
    float al[NA], L[DOF * DOF], Dinv[DOF], x[DOF], r[DOF], dx[DOF];
#pragma unroll
    for (int i = 0; i < NA; ++i) al[i] = a[i];
#pragma unroll
    for (int i = 0; i < DOF; ++i) al[tri(i, i, DOF)] += add[i];
    chol_packed<DOF, float>(al, L, Dinv);
    chol_solve<DOF, float>(L, Dinv, g, x);
    Hi xh[DOF];
#pragma unroll
    for (int i = 0; i < DOF; ++i) {
        Hi v = (Hi)g[i];
#pragma unroll
        for (int j = 0; j < DOF; ++j) {
            const Hi aij = (Hi)a[(i <= j) ? tri(i, j, DOF) : tri(j, i, DOF)] + ((i == j) ? (Hi)add[i] : Hi(0));
            v -= aij * (Hi)x[j];
        }
        r[i] = (float)v;
    }
    chol_solve<DOF, float>(L, Dinv, r, dx);
#pragma unroll
    for (int i = 0; i < DOF; ++i) { xh[i] = (Hi)x[i] + (Hi)dx[i]; step[i] = (float)(-xh[i]); }
    Hi m = 0;
#pragma unroll
    for (int i = 0; i < DOF; ++i) {
        Hi v = 0;
#pragma unroll
        for (int j = 0; j < DOF; ++j) v += (Hi)a[(i <= j) ? tri(i, j, DOF) : tri(j, i, DOF)] * (-xh[j]);
        m += (-xh[i]) * (v * 0.5 + (Hi)g[i]);
    }
    return (float)(-m);
}

// Damped step from the current linearisation -> s.pose_new, s.model_change  (:205-225)
template <int DOF> PNP_HD void lm_propose(LMState<DOF>& s, const Params& p) {
    float add[DOF], step[DOF];
    const float inv_radius = 1.0f / s.radius;
#pragma unroll
    for (int i = 0; i < DOF; ++i) {
        const float d = s.a[tri(i, i, DOF)];
        add[i] = fmaf(fminf(fmaxf(d, p.min_lm_diagonal), p.max_lm_diagonal), inv_radius, p.eps);
    }
    // plain fp32 Cholesky step: inside the iteration a step only has to decrease the cost (the next evaluation corrects
    // its rounding); measured on B200 +2.9 % on the fused solve against the fp64-refined step, parity unchanged
    s.model_change = damped_step<DOF, float>(s.a, s.g, add, step);
    pose_add<DOF>(s.pose, step, s.pose_new);
}

// Accept / reject the pending step given the candidate evaluation `acc`  (:227-240)
template <int DOF> PNP_HD bool lm_update(LMState<DOF>& s, const float* acc, const Params& p) {
    constexpr int NA = Dim<DOF>::NA;
    const float cost_new = acc[NA + DOF];
    const float rho = (s.cost - cost_new) / s.model_change;
    const bool ok = (rho >= p.min_relative_decrease) && (s.model_change > 0.0f);
    if (ok) {
#pragma unroll
        for (int i = 0; i < Dim<DOF>::POSE; ++i) s.pose[i] = s.pose_new[i];
        const float q = 2.0f * rho - 1.0f;
        s.radius = s.radius / fmaxf(1.0f - q * q * q, 1.0f / 3.0f);
    }
    s.radius = fmaxf(fminf(s.radius, p.max_radius), p.eps);
    if (ok) {
        s.shrink = 2.0f;
        lm_adopt<DOF>(s, acc);
    } else {
        s.radius = s.radius / s.shrink;
        s.shrink *= 2.0f;
    }
    return ok;
}

// Gauss-Newton step (fast mode :136-152 and gn_step :243-253): pose_out = pose (+) -(A+eps I)^-1 g
template <int DOF> PNP_HD void gn_advance(const float* pose, const float* acc, float eps, float* pose_out) {
    float add[DOF], step[DOF];
#pragma unroll
    for (int i = 0; i < DOF; ++i) add[i] = eps;
    damped_step_refined<DOF>(acc, acc + Dim<DOF>::NA, add, step);
    pose_add<DOF>(pose, step, pose_out);
}

// pose covariance = (J^T J + eps I)^-1  (:170-179), one COLUMN per caller: the kernel spreads the DOF
// columns over DOF lanes (each repeats the cheap Cholesky and does one pair of substitutions) instead of
// inverting on a single thread.  fp64.
template <int DOF> PNP_HD void pose_covariance_column(const float* a_packed, float eps, int col, float* out) {
    constexpr int NA = Dim<DOF>::NA;
    Hi al[NA], L[DOF * DOF], Dinv[DOF], e[DOF], x[DOF];
#pragma unroll
    for (int i = 0; i < NA; ++i) al[i] = (Hi)a_packed[i];
#pragma unroll
    for (int i = 0; i < DOF; ++i) { al[tri(i, i, DOF)] += (Hi)eps; e[i] = (i == col) ? Hi(1) : Hi(0); }
    chol_packed<DOF, Hi>(al, L, Dinv);
    chol_solve<DOF, Hi>(L, Dinv, e, x);
#pragma unroll
    for (int i = 0; i < DOF; ++i) out[i] = (float)x[i];
}

template <int DOF> PNP_HD void pose_covariance(const float* a_packed, float eps, float* cov_full) {
    float c[DOF];
    for (int j = 0; j < DOF; ++j) {
        pose_covariance_column<DOF>(a_packed, eps, j, c);
        for (int i = 0; i < DOF; ++i) cov_full[i * DOF + j] = c[i];
    }
}

// ------------------------------------------------------------------------------------------------
// AMIS proposal for 6DoF: translation ~ multivariate t (df=3), rotation ~ angular central Gaussian.
// Stored with everything log_prob needs pre-computed.
struct Proposal6 {
    float mu[3];
    float lt[6];     // L_t lower: l00 l10 l11 l20 l21 l22
    float lr[10];    // L_r lower: l00 l10 l11 l20 l21 l22 l30 l31 l32 l33
    float ct, cr;    // additive log-normalisers (already negated)
    float ilt[3], ilr[4];   // reciprocals of the diagonals of L_t / L_r (densities multiply, never divide)
};

// lgamma(1.5) - lgamma(3) + 1.5 log(3 pi):  log-normaliser of the df=3, n=3 Student-t
#define PNP_MVT3_LOGNORM 2.5510838435810745f
// log(2 pi^2): area of S^3
#define PNP_LOG_S3_AREA 2.9826069522587457f

PNP_HD void proposal_finish(Proposal6& p) {
    p.ct = -(logf(p.lt[0]) + logf(p.lt[2]) + logf(p.lt[5]) + PNP_MVT3_LOGNORM);
    p.cr = -(logf(p.lr[0]) + logf(p.lr[2]) + logf(p.lr[5]) + logf(p.lr[9]) + PNP_LOG_S3_AREA);
    p.ilt[0] = 1.0f / p.lt[0]; p.ilt[1] = 1.0f / p.lt[2]; p.ilt[2] = 1.0f / p.lt[5];
    p.ilr[0] = 1.0f / p.lr[0]; p.ilr[1] = 1.0f / p.lr[2]; p.ilr[2] = 1.0f / p.lr[5]; p.ilr[3] = 1.0f / p.lr[9];
}

// chol of a symmetric 3x3 given by its packed upper triangle, with the reference's fallback:
// not PD -> identity (cholesky_wrapper, epropnp.py:16-33; default_diag is None on the 6DoF path)
template <class T> PNP_HD void chol3_or_identity(const T* a6, float* l) {
    T L[9], Dinv[3];
    if (chol_packed<3, T>(a6, L, Dinv)) {
        l[0] = (float)L[0]; l[1] = (float)L[3]; l[2] = (float)L[4]; l[3] = (float)L[6]; l[4] = (float)L[7]; l[5] = (float)L[8];
    } else { l[0] = 1.f; l[1] = 0.f; l[2] = 1.f; l[3] = 0.f; l[4] = 0.f; l[5] = 1.f; }
}

// L_r = chol(C + det(C)^(1/4) * dispersion * I), identity when not PD  (epropnp.py:301-302, 341-342)
template <class T> PNP_HD void acg_dispersed_chol(const T* c10, float dispersion, float* lr) {
    T L[16], a[10], Dinv[4];
    chol_packed<4, T>(c10, L, Dinv);
    const T d = L[0] * L[5] * L[10] * L[15];              // sqrt(det C); NaN if C is not PD
    const T add = (d * inv_sqrt(d)) * (T)dispersion;      // det^(1/4) * dispersion  (sqrt(d) = d * rsqrt(d))
#pragma unroll
    for (int i = 0; i < 10; ++i) a[i] = c10[i];
    a[0] += add; a[4] += add; a[7] += add; a[9] += add;
    if (chol_packed<4, T>(a, L, Dinv)) {
        lr[0] = (float)L[0]; lr[1] = (float)L[4]; lr[2] = (float)L[5]; lr[3] = (float)L[8]; lr[4] = (float)L[9];
        lr[5] = (float)L[10]; lr[6] = (float)L[12]; lr[7] = (float)L[13]; lr[8] = (float)L[14]; lr[9] = (float)L[15];
    } else {
        lr[0] = 1.f; lr[1] = 0.f; lr[2] = 1.f; lr[3] = 0.f; lr[4] = 0.f; lr[5] = 1.f;
        lr[6] = 0.f; lr[7] = 0.f; lr[8] = 0.f; lr[9] = 1.f;
    }
}

// First proposal from the LM solution and its covariance (EProPnP6DoF.initial_fit, epropnp.py:288-302).
// The reference forms (T S^-1 T^T + I_4)^-1 with S = cov[3:,3:] and T = T(q) the 4x3 tangent map.  For a
// unit quaternion [T | q] is orthogonal (columns of T are orthonormal and orthogonal to q), hence
//     (T S^-1 T^T + I)^-1 = T (S^-1 + I)^-1 T^T + q q^T = T [S (I + S)^-1] T^T + q q^T
// which needs one WELL-conditioned 3x3 SPD solve instead of two ill-conditioned inverses, and
// det of that matrix = det(S (I+S)^-1).  Same quantity, evaluated stably; fp64.
PNP_HD_COLD void initial_fit6(const float* pose, const float* cov /*6x6 full*/, float dispersion, Proposal6& p) {
    p.mu[0] = pose[0]; p.mu[1] = pose[1]; p.mu[2] = pose[2];
    const Hi ctt[6] = {(Hi)cov[0], (Hi)cov[1], (Hi)cov[2], (Hi)cov[7], (Hi)cov[8], (Hi)cov[14]};
    chol3_or_identity<Hi>(ctt, p.lt);
    const Hi S[9] = {(Hi)cov[21], (Hi)cov[22], (Hi)cov[23], (Hi)cov[22], (Hi)cov[28], (Hi)cov[29],
                     (Hi)cov[23], (Hi)cov[29], (Hi)cov[35]};
    const Hi ips[6] = {S[0] + 1.0, S[1], S[2], S[4] + 1.0, S[5], S[8] + 1.0};        // I + S (packed upper)
    Hi L[9], Dinv[3], C[9];
    chol_packed<3, Hi>(ips, L, Dinv);
#pragma unroll
    for (int c = 0; c < 3; ++c) {              // C[:, c] = (I + S)^-1 S[:, c]   (C = S (I+S)^-1 is symmetric)
        const Hi b[3] = {S[c], S[3 + c], S[6 + c]};
        Hi x[3];
        chol_solve<3, Hi>(L, Dinv, b, x);
        C[c] = x[0]; C[3 + c] = x[1]; C[6 + c] = x[2];
    }
    const Hi w = pose[3], x = pose[4], y = pose[5], z = pose[6];
    const Hi T[12] = {x, y, z, -w, -z, y, z, -w, -x, -y, x, -w};    // 4x3
    const Hi q[4] = {w, x, y, z};
    Hi TC[12];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            TC[i * 3 + j] = T[i * 3 + 0] * C[0 + j] + T[i * 3 + 1] * C[3 + j] + T[i * 3 + 2] * C[6 + j];
    Hi rc[10], tr = 0;
    int idx = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i; j < 4; ++j) {
            rc[idx] = TC[i * 3 + 0] * T[j * 3 + 0] + TC[i * 3 + 1] * T[j * 3 + 1] + TC[i * 3 + 2] * T[j * 3 + 2] + q[i] * q[j];
            if (i == j) tr += rc[idx];
            ++idx;
        }
    const Hi itr = fast_recip(tr);
#pragma unroll
    for (int i = 0; i < 10; ++i) rc[i] *= itr;
    acg_dispersed_chol<Hi>(rc, dispersion, p.lr);
    proposal_finish(p);
}

// The refits in the middle of the AMIS loop work from statistics that were accumulated in fp32 over
// softmax weights carrying ~1e-4 of cost noise, exactly like the reference's fp32
// tensors.  Measured on the golden cases: the thread-0 factorization that closes a refit (`Refit`, on
// the serial critical path between two AMIS iterations) is as accurate in fp32 as in fp64, so it is a
// short MUFU.RSQ chain; the scatter-matrix inverse inside the ACG fixed point (`RefitInv`, evaluated by
// all threads in parallel, not serial-critical) does lose accuracy in fp32 and stays fp64.
typedef float Refit;
typedef double RefitInv;

// Lambda^-1 (fp32 out) of the ACG scatter matrix given as packed fp32 upper triangle (a00 a01 a02 a03 a11 a12 a13 a22
// a23 a33).  The reference calls rot_cov.inverse() (epropnp.py:336).  Closed form through the twelve 2x2 minors of the
// row pairs (0,1) and (2,3), in fp64: ~80 multiply-adds of depth 4 and ONE reciprocal, instead of the ~300-instruction
// dependent chain of a Cholesky factorisation + triangular inverse + product (every thread of the CTA evaluates this
// twice per refit, on the critical path between two AMIS iterations).  Lambda = sum w q q^T + eps I is positive
// definite; the minors' cancellation costs cond(Lambda) * 1e-16 of relative accuracy, far below the fp32 output.
PNP_HD void acg_scatter_inverse(const float* lam10, float* inv16) {
    typedef RefitInv T;
    const T a00 = lam10[0], a01 = lam10[1], a02 = lam10[2], a03 = lam10[3], a11 = lam10[4], a12 = lam10[5], a13 = lam10[6],
            a22 = lam10[7], a23 = lam10[8], a33 = lam10[9];
    // minors of rows 0, 1 (columns i < j) and of rows 2, 3; symmetric input: a10 = a01, a20 = a02, ...
    const T s0 = a00 * a11 - a01 * a01, s1 = a00 * a12 - a01 * a02, s2 = a00 * a13 - a01 * a03;
    const T s3 = a01 * a12 - a11 * a02, s4 = a01 * a13 - a11 * a03, s5 = a02 * a13 - a12 * a03;
    const T c5 = a22 * a33 - a23 * a23, c4 = a12 * a33 - a13 * a23, c3 = a12 * a23 - a13 * a22;
    const T c2 = a02 * a33 - a03 * a23, c1 = a02 * a23 - a03 * a22, c0 = a02 * a13 - a03 * a12;
    const T det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
    const T id = fast_recip(det);
    const T i00 = (a11 * c5 - a12 * c4 + a13 * c3) * id;
    const T i01 = (-a01 * c5 + a02 * c4 - a03 * c3) * id;
    const T i02 = (a13 * s5 - a23 * s4 + a33 * s3) * id;
    const T i03 = (-a12 * s5 + a22 * s4 - a23 * s3) * id;
    const T i11 = (a00 * c5 - a02 * c2 + a03 * c1) * id;
    const T i12 = (-a03 * s5 + a23 * s2 - a33 * s1) * id;
    const T i13 = (a02 * s5 - a22 * s2 + a23 * s1) * id;
    const T i22 = (a03 * s4 - a13 * s2 + a33 * s0) * id;
    const T i23 = (-a02 * s4 + a12 * s2 - a23 * s0) * id;
    const T i33 = (a02 * s3 - a12 * s1 + a22 * s0) * id;
    inv16[0] = (float)i00; inv16[1] = (float)i01; inv16[2] = (float)i02; inv16[3] = (float)i03;
    inv16[4] = (float)i01; inv16[5] = (float)i11; inv16[6] = (float)i12; inv16[7] = (float)i13;
    inv16[8] = (float)i02; inv16[9] = (float)i12; inv16[10] = (float)i22; inv16[11] = (float)i23;
    inv16[12] = (float)i03; inv16[13] = (float)i13; inv16[14] = (float)i23; inv16[15] = (float)i33;
}

// Next proposal from the refit statistics (EProPnP6DoF.estimate_params tail, epropnp.py:325, 341-342)
PNP_HD void refit_finish6(const float* mean, const float* tc6, const float* lam10, float dispersion, Proposal6& np) {
    np.mu[0] = mean[0]; np.mu[1] = mean[1]; np.mu[2] = mean[2];
    Refit a6[6], c10[10];
#pragma unroll
    for (int i = 0; i < 6; ++i) a6[i] = (Refit)tc6[i];
#pragma unroll
    for (int i = 0; i < 10; ++i) c10[i] = (Refit)lam10[i];
    chol3_or_identity<Refit>(a6, np.lt);
    acg_dispersed_chol<Refit>(c10, dispersion, np.lr);
    proposal_finish(np);
}

// one sample from the proposal given base noise  (pyro MultivariateStudentT.rsample;
// distributions.py:42-52 AngularCentralGaussian.rsample)
PNP_HD void proposal_draw6(const Proposal6& p, const float* n3, float chi2, const float* n4, float* smp) {
    const float sc = 1.0f / sqrtf(chi2 / 3.0f);
    const float y0 = n3[0] * sc, y1 = n3[1] * sc, y2 = n3[2] * sc;
    smp[0] = p.mu[0] + p.lt[0] * y0;
    smp[1] = p.mu[1] + (p.lt[1] * y0 + p.lt[2] * y1);
    smp[2] = p.mu[2] + (p.lt[3] * y0 + p.lt[4] * y1 + p.lt[5] * y2);
    const float g0 = p.lr[0] * n4[0];
    const float g1 = p.lr[1] * n4[0] + p.lr[2] * n4[1];
    const float g2 = p.lr[3] * n4[0] + p.lr[4] * n4[1] + p.lr[5] * n4[2];
    const float g3 = p.lr[6] * n4[0] + p.lr[7] * n4[1] + p.lr[8] * n4[2] + p.lr[9] * n4[3];
    const float n = sqrtf(g0 * g0 + g1 * g1 + g2 * g2 + g3 * g3);
    if (n < 1e-6f) { smp[3] = 1.f; smp[4] = 0.f; smp[5] = 0.f; smp[6] = 0.f; }
    else { smp[3] = g0 / n; smp[4] = g1 / n; smp[5] = g2 / n; smp[6] = g3 / n; }
}

// log / exp of the proposal densities and of the weight bookkeeping: the special-function unit on the device (lg2.approx /
// ex2.approx: absolute error ~4e-7 of the logarithm, <= ~30 ulp of an exponential of |x| < 20 -- the log-weights these
// feed are compared at 1e-4 of a scale of several hundred), libm in the host build.
PNP_HD float fast_log(float x) {
#if defined(__CUDA_ARCH__)
    return __logf(x);
#else
    return logf(x);
#endif
}
PNP_HD float fast_log1p(float x) {
#if defined(__CUDA_ARCH__)
    return __logf(1.0f + x);
#else
    return log1pf(x);
#endif
}
PNP_HD float fast_exp(float x) {
#if defined(__CUDA_ARCH__)
    return __expf(x);
#else
    return expf(x);
#endif
}

// log q(sample) under one proposal  (pyro MultivariateStudentT.log_prob + distributions.py:32-40)
PNP_HD float proposal_logpdf6(const Proposal6& p, const float* smp) {
    const float d0 = smp[0] - p.mu[0], d1 = smp[1] - p.mu[1], d2 = smp[2] - p.mu[2];
    const float y0 = d0 * p.ilt[0];
    const float y1 = (d1 - p.lt[1] * y0) * p.ilt[1];
    const float y2 = (d2 - p.lt[3] * y0 - p.lt[4] * y1) * p.ilt[2];
    const float mt = y0 * y0 + y1 * y1 + y2 * y2;
    const float z0 = smp[3] * p.ilr[0];
    const float z1 = (smp[4] - p.lr[1] * z0) * p.ilr[1];
    const float z2 = (smp[5] - p.lr[3] * z0 - p.lr[4] * z1) * p.ilr[2];
    const float z3 = (smp[6] - p.lr[6] * z0 - p.lr[7] * z1 - p.lr[8] * z2) * p.ilr[3];
    const float mr = z0 * z0 + z1 * z1 + z2 * z2 + z3 * z3;
    return (-3.0f * fast_log1p(mt * (1.0f / 3.0f)) + p.ct) + (-2.0f * fast_log(mr) + p.cr);
}

// q^T Lambda^-1 q for the ACG fixed-point iteration (epropnp.py:335-337)
PNP_HD float quad4(const float* inv16, const float* q) {
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) v = fmaf(inv16[i * 4 + j], q[j], v);
        r = fmaf(q[i], v, r);
    }
    return r;
}

// ------------------------------------------------------------------------------------------------
// AMIS proposal for 4DoF (EProPnP4DoF, epropnp.py:199-260): translation ~ multivariate t (df=3), yaw ~
// 0.75 von Mises(mode, kappa) + 0.25 uniform on the circle (distributions.py:55-79).
struct Proposal4 {
    float mu[3];
    float lt[6];        // L_t lower: l00 l10 l11 l20 l21 l22
    float ilt[3];       // reciprocal diagonal
    float ct;           // -(sum log diag L_t + Student-t log-normaliser)
    float mode, kappa;  // von Mises location / concentration
    float cvm;          // log(0.75) - log(2 pi) - log I0(kappa)
};

#define PNP_LOG_UNIFORM_MIX (-3.2241714810970856f)      // log(0.25 / (2 pi))
#define PNP_LOG_VM_MIX (-0.2876820724517809f)           // log(0.75)
#define PNP_LOG_2PI 1.8378770664093453f

// log I0(x) with the polynomial approximations torch.distributions.von_mises uses
// (_log_modified_bessel_fn, order 0: Abramowitz-Stegun 9.8.1 / 9.8.2), so densities match the reference's.
PNP_HD float log_bessel_i0(float x) {
    if (x < 3.75f) {
        float y = x / 3.75f;
        y = y * y;
        float r = 0.45813e-2f;
        r = 0.360768e-1f + y * r; r = 0.2659732f + y * r; r = 1.2067492f + y * r;
        r = 3.0899424f + y * r; r = 3.5156229f + y * r; r = 1.0f + y * r;
        return logf(r);
    }
    const float y = 3.75f / x;
    float r = 0.392377e-2f;
    r = -0.1647633e-1f + y * r; r = 0.2635537e-1f + y * r; r = -0.2057706e-1f + y * r; r = 0.916281e-2f + y * r;
    r = -0.157565e-2f + y * r; r = 0.225319e-2f + y * r; r = 0.1328592e-1f + y * r; r = 0.39894228f + y * r;
    return x - 0.5f * logf(x) + logf(r);
}

PNP_HD void proposal4_finish(Proposal4& p) {
    p.ct = -(logf(p.lt[0]) + logf(p.lt[2]) + logf(p.lt[5]) + PNP_MVT3_LOGNORM);
    p.ilt[0] = 1.0f / p.lt[0]; p.ilt[1] = 1.0f / p.lt[2]; p.ilt[2] = 1.0f / p.lt[5];
    p.cvm = PNP_LOG_VM_MIX - PNP_LOG_2PI - log_bessel_i0(p.kappa);
}

// chol of a symmetric 3x3 (packed upper), not PD -> diag(1, 1, 4): cholesky_wrapper(..., [1.0, 1.0, 4.0])
template <class T> PNP_HD void chol3_or_default4(const T* a6, float* l) {
    T L[9], Dinv[3];
    if (chol_packed<3, T>(a6, L, Dinv)) {
        l[0] = (float)L[0]; l[1] = (float)L[3]; l[2] = (float)L[4]; l[3] = (float)L[6]; l[4] = (float)L[7]; l[5] = (float)L[8];
    } else { l[0] = 1.f; l[1] = 0.f; l[2] = 1.f; l[3] = 0.f; l[4] = 0.f; l[5] = 4.f; }
}

// EProPnP4DoF.initial_fit (epropnp.py:216-220): kappa_0 = 0.33 / max(var(yaw), eps)
PNP_HD_COLD void initial_fit4(const float* pose, const float* cov /*4x4 full*/, float eps, Proposal4& p) {
    p.mu[0] = pose[0]; p.mu[1] = pose[1]; p.mu[2] = pose[2];
    const Hi ctt[6] = {(Hi)cov[0], (Hi)cov[1], (Hi)cov[2], (Hi)cov[5], (Hi)cov[6], (Hi)cov[10]};
    chol3_or_default4<Hi>(ctt, p.lt);
    p.mode = pose[3];
    p.kappa = 0.33f / fmaxf(cov[15], eps);
    proposal4_finish(p);
}

// EProPnP4DoF.estimate_params tail (epropnp.py:246-260) from the weighted statistics
PNP_HD void refit_finish4(const float* mean, const float* tc6, float sum_sin, float sum_cos, float eps, Proposal4& np) {
    np.mu[0] = mean[0]; np.mu[1] = mean[1]; np.mu[2] = mean[2];
    Refit a6[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) a6[i] = (Refit)tc6[i];
    chol3_or_default4<Refit>(a6, np.lt);
    np.mode = atan2f(sum_sin, sum_cos);
    const float r_sq = sum_sin * sum_sin + sum_cos * sum_cos;
    np.kappa = 0.33f * fmaxf(sqrtf(r_sq), eps) * (2.0f - r_sq) / fmaxf(1.0f - r_sq, eps);
    proposal4_finish(np);
}

PNP_HD float wrap_pi(float a) {
    const float two_pi = 6.283185307179586f;
    a = a - two_pi * floorf((a + 3.14159265358979f) / two_pi);
    return a;
}

// translation part of a draw (shared with 6DoF): t = mu + L (n3 * sqrt(3 / chi2))
PNP_HD void draw_translation(const float* mu, const float* lt, const float* n3, float chi2, float* t) {
    const float sc = 1.0f / sqrtf(chi2 / 3.0f);
    const float y0 = n3[0] * sc, y1 = n3[1] * sc, y2 = n3[2] * sc;
    t[0] = mu[0] + lt[0] * y0;
    t[1] = mu[1] + (lt[1] * y0 + lt[2] * y1);
    t[2] = mu[2] + (lt[3] * y0 + lt[4] * y1 + lt[5] * y2);
}

// log q(sample) under one 4DoF proposal: Student-t + logaddexp(von Mises part, uniform part)
PNP_HD float proposal_logpdf4(const Proposal4& p, const float* smp) {
    const float d0 = smp[0] - p.mu[0], d1 = smp[1] - p.mu[1], d2 = smp[2] - p.mu[2];
    const float y0 = d0 * p.ilt[0];
    const float y1 = (d1 - p.lt[1] * y0) * p.ilt[1];
    const float y2 = (d2 - p.lt[3] * y0 - p.lt[4] * y1) * p.ilt[2];
    const float mt = y0 * y0 + y1 * y1 + y2 * y2;
    const float vm = p.kappa * cosf(smp[3] - p.mode) + p.cvm;
    const float hi = fmaxf(vm, PNP_LOG_UNIFORM_MIX), lo = fminf(vm, PNP_LOG_UNIFORM_MIX);
    const float rot = hi + fast_log1p(fast_exp(lo - hi));
    return (-3.0f * fast_log1p(mt * (1.0f / 3.0f)) + p.ct) + rot;
}

// ------------------------------------------------------------------------------------------------
// Counter-based RNG for the production (non-injected) AMIS draws: Philox-4x32-10.
struct Philox {
    uint32_t k0, k1;
    PNP_HD static void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
        const uint64_t p = (uint64_t)a * b;
        hi = (uint32_t)(p >> 32); lo = (uint32_t)p;
    }
    PNP_HD void operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t* out) const {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t h0, l0, h1, l1;
            mulhilo(0xD2511F53u, c0, h0, l0);
            mulhilo(0xCD9E8D57u, c2, h1, l1);
            const uint32_t n0 = h1 ^ c1 ^ a, n2 = h0 ^ c3 ^ b;
            c0 = n0; c1 = l1; c2 = n2; c3 = l0;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};

// two uniforms -> two standard normals (Box-Muller)
PNP_HD void box_muller(uint32_t u0, uint32_t u1, float& n0, float& n1) {
    const float a = ((float)(u0 >> 8) + 0.5f) * (1.0f / 16777216.0f);     // (0,1)
    const float b = ((float)(u1 >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float r, s, c;
#if defined(__CUDA_ARCH__)
    // MUFU-based log / sin / cos: ~1e-6 relative error, irrelevant for random draws, ~4x fewer instructions
    r = sqrtf(-2.0f * __logf(a));
    __sincosf(6.283185307179586f * b - 3.14159265358979f, &s, &c);
#else
    r = sqrtf(-2.0f * logf(a));
    s = sinf(6.283185307179586f * b - 3.14159265358979f); c = cosf(6.283185307179586f * b - 3.14159265358979f);
#endif
    n0 = r * c; n1 = r * s;
}

// base noise of sample `m` of global object `obj`: n3[3], chi2 (3 dof), n4[4]
PNP_HD void draw_base_noise(uint64_t seed, uint32_t obj, uint32_t m, float* n3, float& chi2, float* n4) {
    Philox ph{(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t r[12];
    ph(obj, m, 0u, 0x45505250u, r);
    ph(obj, m, 1u, 0x45505250u, r + 4);
    ph(obj, m, 2u, 0x45505250u, r + 8);
    float g[12];
#pragma unroll
    for (int i = 0; i < 6; ++i) box_muller(r[2 * i], r[2 * i + 1], g[2 * i], g[2 * i + 1]);
    n3[0] = g[0]; n3[1] = g[1]; n3[2] = g[2];
    chi2 = g[3] * g[3] + g[4] * g[4] + g[5] * g[5];
    n4[0] = g[6]; n4[1] = g[7]; n4[2] = g[8]; n4[3] = g[9];
}

// translation-only base noise (4DoF): n3[3], chi2 -- the same Philox blocks 0 and 1 as draw_base_noise
PNP_HD void draw_base_noise_t(uint64_t seed, uint32_t obj, uint32_t m, float* n3, float& chi2) {
    Philox ph{(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t r[8];
    ph(obj, m, 0u, 0x45505250u, r);
    ph(obj, m, 1u, 0x45505250u, r + 4);
    float g[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) box_muller(r[2 * i], r[2 * i + 1], g[2 * i], g[2 * i + 1]);
    n3[0] = g[0]; n3[1] = g[1]; n3[2] = g[2];
    chi2 = g[3] * g[3] + g[4] * g[4] + g[5] * g[5];
}

// uniform in (0, 1) from 24 random bits
PNP_HD float u01(uint32_t r) { return ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// Production 4DoF draw of sample `m` (index `s_in_iter` within its AMIS iteration of `S` samples):
// n3 / chi2 as in the 6DoF case; yaw: the first round(0.25 S) samples of an iteration are uniform on
// [-pi, pi), the rest von Mises(mode, kappa) by Best-Fisher rejection (distributions.py:61-72 uses numpy's
// sampler of the same family).  Returns the yaw.
PNP_HD float draw_yaw(uint64_t seed, uint32_t obj, uint32_t m, int s_in_iter, int S, float mode, float kappa) {
    Philox ph{(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t r[4];
    const int n_uniform = (int)floorf(0.25f * (float)S + 0.5f);
    ph(obj, m, 2u, 0x45505250u, r);
    if (s_in_iter < n_uniform) return (2.0f * u01(r[3]) - 1.0f) * 3.14159265358979f;
    if (!(kappa > 1e-12f)) return (2.0f * u01(r[3]) - 1.0f) * 3.14159265358979f;    // flat (or NaN) concentration
    // Best-Fisher constants.  rho = (tau - sqrt(2 tau)) / (2 kappa) cancels catastrophically in fp32 below kappa ~ 1e-3
    // (tau rounds to 2, rho to 0, rr to inf); tau (tau - 2) = 4 kappa^2 gives the cancellation-free form used here, which
    // tends to kappa / 2 (rr -> 1 / kappa, the uniform limit) as kappa -> 0.
    const float tau = 1.0f + sqrtf(1.0f + 4.0f * kappa * kappa);
    const float rho = 2.0f * kappa / (tau + sqrtf(2.0f * tau));
    const float rr = (1.0f + rho * rho) / (2.0f * rho);
    float f = 1.0f;
    for (uint32_t attempt = 0; attempt < 64u; ++attempt) {
        ph(obj, m, 3u + attempt, 0x45505250u, r);
        const float z = cosf(3.14159265358979f * u01(r[0]));
        f = (1.0f + rr * z) / (rr + z);
        const float c = kappa * (rr - f);
        const float u2 = u01(r[1]);
        if (c * (2.0f - c) - u2 > 0.0f || logf(c / u2) + 1.0f - c >= 0.0f) {
            const float a = acosf(fminf(fmaxf(f, -1.0f), 1.0f));
            return wrap_pi(mode + ((u01(r[2]) < 0.5f) ? -a : a));
        }
    }
    return wrap_pi(mode);
}

}  // namespace pnp
