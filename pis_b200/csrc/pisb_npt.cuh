// pisb_npt.cuh -- the NPT wrapper around the hot path: kernels + host 3x3 arithmetic for
//   PotentialManager::verlet_step_npt_mtk   src/potentials/potential.rs:112-135
//   MTKBarostat                             src/ensemble/npt.rs:11-88
//   Atoms::scale_box                        src/atoms/transformations.rs:6-15
//   Atoms::pressure_tensor                  src/atoms/properties.rs:45-59
//
// The barostat itself is nine numbers; it lives on the HOST (one small device->host read of the reduced tensors per
// step, which the barostat's data dependence forces anyway: the box of step k+1 follows from the pressure of step k).
// What runs on the device are the two per-atom passes either side of the NVT step:
//   k_npt_pre  : v = S v (potential.rs:120);  scale_box: s = h_inv x, x = h' s (transformations.rs:7-14); the
//                list-build positions xb take the same affine map, so the skin trigger keeps measuring the
//                NON-affine displacement (the affine part is bounded on the host from |h' h_build^-1 - I|);
//                KE of the scaled velocities (the thermostat's first half step starts from it, potential.rs:41)
//   k_npt_post : v = S v (potential.rs:129);  V V^T (mass-less, as in the reference), X F^T, KE, all in one pass
//
// Third-party arithmetic restated on the host (source not under /root/reference): nalgebra 0.34.1 Matrix3::exp =
// Al-Mohy & Higham (2009) scaling-and-squaring Pade approximant; the barostat's arguments are tiny, so it is the
// order-3 approximant U = A (A^2 + 60 I), V = 12 A^2 + 120 I, exp(A) = (V - U)^-1 (V + U) by LU with partial pivoting
// (higher norms fall through to orders 5..13).  try_inverse = adjugate / determinant.
#pragma once
#include "pisb_kernels.cuh"

namespace pisb {

struct Mat3Dev {
    double m[9];  // column-major
};

struct NptPreArgs {
    int n;
    double4 *xt;
    float4 *xf;
    double *vx, *vy, *vz;
    double *xbx, *xby, *xbz;
    const double *mass;
    Mat3Dev scale;      // velocity scaling matrix S
    Mat3Dev hinv_old;   // box inverse before scale_box
    Mat3Dev h_new;      // box after scale_box
    int new_ortho;
    int pbc[3];
    double *partials;
    unsigned int *ticket;
    pisb_thermo *ke_out;  // ke_out->ke = KE of the scaled velocities
};

__device__ __forceinline__ void mat3_apply(const Mat3Dev &a, double &x, double &y, double &z) {
    double r0, r1, r2;
    matvec<false>(a.m, x, y, z, r0, r1, r2);
    x = r0;
    y = r1;
    z = r2;
}

__global__ void __launch_bounds__(TPB) k_npt_pre(NptPreArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double red[1] = {0.0};
    if (i < a.n) {
        double4 x = a.xt[i];
        double vx = a.vx[i], vy = a.vy[i], vz = a.vz[i];
        double bx = a.xbx[i], by = a.xby[i], bz = a.xbz[i];
        const double m = a.mass[type_of(x.w) - 1];
        mat3_apply(a.scale, vx, vy, vz);
        a.vx[i] = vx;
        a.vy[i] = vy;
        a.vz[i] = vz;
        red[0] = __dmul_rn(__dmul_rn(0.5, m), norm2(vx, vy, vz));
        mat3_apply(a.hinv_old, x.x, x.y, x.z);
        mat3_apply(a.h_new, x.x, x.y, x.z);
        a.xt[i] = x;
        mat3_apply(a.hinv_old, bx, by, bz);
        mat3_apply(a.h_new, bx, by, bz);
        a.xbx[i] = bx;
        a.xby[i] = by;
        a.xbz[i] = bz;
        a.xf[i] = make_float4((float)x.x, (float)x.y, (float)x.z, __int_as_float(type_of(x.w)));
    }
    pisb_thermo *out = a.ke_out;
    block_reduce_finalize<1, TPB>(red, a.partials, a.ticket, [&](int, double s) { out->ke = s; });
}

struct NptPostArgs {
    int n;
    const double4 *xt;
    double *vx, *vy, *vz;
    const double *fx, *fy, *fz;
    const double *mass;
    Mat3Dev scale;
    int apply_scale;
    double *partials;
    unsigned int *ticket;
    double *tensors;      // [0..8] V V^T, [9..17] X F^T (column-major), [18] KE
    pisb_thermo *thermo;  // may be null: ke and virial_ref of the record are overwritten with the post-scaling values
};

__global__ void __launch_bounds__(TPB) k_npt_post(NptPostArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double red[19];
#pragma unroll
    for (int q = 0; q < 19; ++q) red[q] = 0.0;
    if (i < a.n) {
        const double4 x = a.xt[i];
        double v[3] = {a.vx[i], a.vy[i], a.vz[i]};
        const double f[3] = {a.fx[i], a.fy[i], a.fz[i]};
        const double m = a.mass[type_of(x.w) - 1];
        if (a.apply_scale) {
            mat3_apply(a.scale, v[0], v[1], v[2]);
            a.vx[i] = v[0];
            a.vy[i] = v[1];
            a.vz[i] = v[2];
        }
        const double r[3] = {x.x, x.y, x.z};
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
                red[c * 3 + rr] = __dmul_rn(v[rr], v[c]);
                red[9 + c * 3 + rr] = __dmul_rn(r[rr], f[c]);
            }
        red[18] = __dmul_rn(__dmul_rn(0.5, m), norm2(v[0], v[1], v[2]));
    }
    double *t = a.tensors;
    pisb_thermo *th = a.thermo;
    block_reduce_finalize<19, TPB>(red, a.partials, a.ticket, [&](int q, double s) {
        t[q] = s;
        if (th && q == 18) {
            th->ke = s;
            th->virial_ref = (t[9] + t[13]) + t[17];
        }
    });
}

// ---- host 3x3 arithmetic in nalgebra's evaluation order (column-major) ----
namespace m3 {

inline void matvec(const double *m, const double *x, double *y) {
    double y0 = m[0] * x[0], y1 = m[1] * x[0], y2 = m[2] * x[0];
    y0 = m[3] * x[1] + y0;
    y1 = m[4] * x[1] + y1;
    y2 = m[5] * x[1] + y2;
    y0 = m[6] * x[2] + y0;
    y1 = m[7] * x[2] + y1;
    y2 = m[8] * x[2] + y2;
    y[0] = y0;
    y[1] = y1;
    y[2] = y2;
}
inline void mul(const double *a, const double *b, double *c) {
    double t[9];
    for (int j = 0; j < 3; ++j) matvec(a, b + 3 * j, t + 3 * j);
    for (int k = 0; k < 9; ++k) c[k] = t[k];
}
inline double det(const double *m) {
    const double c0 = m[4] * m[8] - m[5] * m[7];
    const double c1 = m[1] * m[8] - m[2] * m[7];
    const double c2 = m[1] * m[5] - m[2] * m[4];
    return m[0] * c0 - m[3] * c1 + m[6] * c2;
}
// Matrix3::try_inverse: adjugate / determinant.  false if singular.
inline bool inverse(const double *m, double *o) {
    const double c0 = m[4] * m[8] - m[5] * m[7];
    const double c1 = m[1] * m[8] - m[2] * m[7];
    const double c2 = m[1] * m[5] - m[2] * m[4];
    const double d = m[0] * c0 - m[3] * c1 + m[6] * c2;
    if (d == 0.0) return false;
    o[0] = c0 / d;
    o[3] = (m[6] * m[5] - m[8] * m[3]) / d;
    o[6] = (m[3] * m[7] - m[4] * m[6]) / d;
    o[1] = -c1 / d;
    o[4] = (m[0] * m[8] - m[2] * m[6]) / d;
    o[7] = (m[6] * m[1] - m[7] * m[0]) / d;
    o[2] = c2 / d;
    o[5] = (m[3] * m[2] - m[5] * m[0]) / d;
    o[8] = (m[0] * m[4] - m[1] * m[3]) / d;
    return true;
}
inline void symmetrize(const double *a, double *o) {  // src/math.rs:41-43
    double t[9];
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) t[c * 3 + r] = (a[c * 3 + r] + a[r * 3 + c]) * 0.5;
    for (int k = 0; k < 9; ++k) o[k] = t[k];
}
inline double onenorm(const double *a) {
    double best = 0.0;
    for (int j = 0; j < 3; ++j) best = std::max(best, (std::fabs(a[3 * j]) + std::fabs(a[3 * j + 1])) + std::fabs(a[3 * j + 2]));
    return best;
}
inline void solve_lu(const double *q, const double *p, double *x) {
    double lu[9];
    int perm[3] = {0, 1, 2};
    for (int k = 0; k < 9; ++k) lu[k] = q[k];
    auto L = [&](int r, int c) -> double & { return lu[c * 3 + r]; };
    for (int k = 0; k < 3; ++k) {
        int piv = k;
        for (int r = k + 1; r < 3; ++r)
            if (std::fabs(L(r, k)) > std::fabs(L(piv, k))) piv = r;
        if (piv != k) {
            for (int c = 0; c < 3; ++c) std::swap(L(k, c), L(piv, c));
            std::swap(perm[k], perm[piv]);
        }
        for (int r = k + 1; r < 3; ++r) {
            L(r, k) = L(r, k) / L(k, k);
            for (int c = k + 1; c < 3; ++c) L(r, c) -= L(r, k) * L(k, c);
        }
    }
    for (int j = 0; j < 3; ++j) {
        double y[3];
        for (int r = 0; r < 3; ++r) y[r] = p[3 * j + perm[r]];
        for (int r = 1; r < 3; ++r)
            for (int c = 0; c < r; ++c) y[r] -= L(r, c) * y[c];
        for (int r = 2; r >= 0; --r) {
            for (int c = r + 1; c < 3; ++c) y[r] -= L(r, c) * y[c];
            y[r] = y[r] / L(r, r);
        }
        for (int r = 0; r < 3; ++r) x[3 * j + r] = y[r];
    }
}
inline int ell(const double *a, int m) {
    // C(2p, p) (2p + 1)! for p = 2m + 1
    const double c = m == 3 ? 4487938430976000.0 : m == 5 ? 1.8236839872145106e+28 : m == 7 ? 1.275506339396217e+42
                   : m == 9 ? 7.209685231212166e+56 : 2.4719128253168207e+88;
    double aa[9], pw[9];
    for (int k = 0; k < 9; ++k) pw[k] = aa[k] = std::fabs(a[k]);
    for (int k = 1; k < 2 * m + 1; ++k) mul(pw, aa, pw);
    const double a1 = onenorm(a);
    if (a1 == 0.0) return 0;
    const double alpha = onenorm(pw) / (a1 * c);
    if (alpha == 0.0) return 0;
    const double v = std::ceil(std::log2(alpha / std::ldexp(1.0, -53)) / (2.0 * m));
    return v > 0.0 ? (int)v : 0;
}
// Pade numerator/denominator halves: U = A sum b[2k+1] A^2k, V = sum b[2k] A^2k
inline void pade_uv(const double *a, int m, const double *b, double *u, double *v) {
    double a2[9], pw[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, su[9], sv[9];
    mul(a, a, a2);
    for (int k = 0; k < 9; ++k) {
        su[k] = pw[k] * b[1];
        sv[k] = pw[k] * b[0];
    }
    for (int k = 1; 2 * k <= m; ++k) {
        mul(pw, a2, pw);
        for (int e = 0; e < 9; ++e) {
            su[e] = pw[e] * b[2 * k + 1] + su[e];
            sv[e] = pw[e] * b[2 * k] + sv[e];
        }
    }
    mul(a, su, u);
    for (int k = 0; k < 9; ++k) v[k] = sv[k];
}
inline void expm(const double *a, double *out) {
    static const double b3[] = {120., 60., 12., 1.};
    static const double b5[] = {30240., 15120., 3360., 420., 30., 1.};
    static const double b7[] = {17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.};
    static const double b9[] = {17643225600., 8821612800., 2075673600., 302702400., 30270240., 2162160., 110880., 3960., 90., 1.};
    static const double b13[] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
                                 129060195264000., 10559470521600., 670442572800., 33522128640.,
                                 1323241920., 40840800., 960960., 16380., 182., 1.};
    double a2[9], a4[9], a6[9], u[9], v[9], p[9], q[9];
    mul(a, a, a2);
    mul(a2, a2, a4);
    mul(a4, a2, a6);
    const double d4 = std::pow(onenorm(a4), 0.25), d6 = std::pow(onenorm(a6), 1.0 / 6.0);
    const double eta1 = std::max(d4, d6);
    int order = 0, s = 0;
    const double *b = nullptr;
    double as[9];
    for (int k = 0; k < 9; ++k) as[k] = a[k];
    if (eta1 < 1.495585217958292e-2 && ell(a, 3) == 0) order = 3, b = b3;
    else if (eta1 < 2.539398330063230e-1 && ell(a, 5) == 0) order = 5, b = b5;
    else {
        double a8[9], a10[9];
        mul(a4, a4, a8);
        const double d8 = std::pow(onenorm(a8), 0.125), eta3 = std::max(d6, d8);
        if (eta3 < 9.504178996162932e-1 && ell(a, 7) == 0) order = 7, b = b7;
        else if (eta3 < 2.097847961257068 && ell(a, 9) == 0) order = 9, b = b9;
        else {
            mul(a4, a6, a10);
            const double d10 = std::pow(onenorm(a10), 0.1), eta5 = std::min(eta3, std::max(d8, d10));
            if (eta5 > 0.0) s = std::max(0, (int)std::ceil(std::log2(eta5 / 4.25)));
            for (int k = 0; k < 9; ++k) as[k] = std::ldexp(a[k], -s);
            s += ell(as, 13);
            for (int k = 0; k < 9; ++k) as[k] = std::ldexp(a[k], -s);
            order = 13;
            b = b13;
        }
    }
    if (order == 13) {
        double b2[9], b4[9], b6[9], t1[9], t2[9];
        const double eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        mul(as, as, b2);
        mul(b2, b2, b4);
        mul(b4, b2, b6);
        for (int k = 0; k < 9; ++k) t1[k] = b[13] * b6[k] + b[11] * b4[k] + b[9] * b2[k];
        mul(b6, t1, t2);
        for (int k = 0; k < 9; ++k) t2[k] = t2[k] + b[7] * b6[k] + b[5] * b4[k] + b[3] * b2[k] + b[1] * eye[k];
        mul(as, t2, u);
        for (int k = 0; k < 9; ++k) t1[k] = b[12] * b6[k] + b[10] * b4[k] + b[8] * b2[k];
        mul(b6, t1, t2);
        for (int k = 0; k < 9; ++k) v[k] = t2[k] + b[6] * b6[k] + b[4] * b4[k] + b[2] * b2[k] + b[0] * eye[k];
    } else {
        pade_uv(a, order, b, u, v);
    }
    for (int k = 0; k < 9; ++k) {
        p[k] = u[k] + v[k];
        q[k] = v[k] - u[k];
    }
    solve_lu(q, p, out);
    for (int k = 0; k < s; ++k) mul(out, out, out);
}

}  // namespace m3
}  // namespace pisb
