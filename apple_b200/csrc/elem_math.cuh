// Per-tetrahedron hyperelastic math: energy, first Piola-Kirchhoff force, Hessian diagonal,
// Hessian-vector product and Hessian quadratic form, for the three energies the reference ships.
//
// Written from the closed forms in SURVEY.md appendix A, i.e. what these reference files compute:
//   warp/fem/func/_deformation.py:15-31   F = u^T dhdX + I, jvp = p^T dhdX, vjp = dhdX M^T
//   warp/fem/func/_gradient.py:30-40      cof(F) = [f1 x f2 | f2 x f0 | f0 x f1]
//   warp/fem/_stable_neo_hookean.py:17-103
//   warp/fem/_stable_neo_hookean_muscle.py:18-114
//   warp/fem/_arap.py:17-76  (hess_prod is the mathematically correct product, not the
//                             reference's swapped-argument variant)
//   warp/fem/func/_misc.py:31-70          lambdas (clamped), twist matrices Q0..Q2
//   warp/fem/_base.py:317-320,379-380     per-entry clamp of the diagonal, per-cell clamp of p^T H p
//
// Layout choice (differs from the reference): only the 3x3 block D = rows 1..3 of dhdX is kept
// (row 0 of dhdX is minus their sum), so F = I + sum_a (u_a - u_0) (x) D_a and the nodal force of
// corner 0 is minus the sum of the other three.  Everything is a template on the scalar type and
// is __host__ __device__ so that the same code is checked on the CPU against the oracle by
// tests/native (test-only harness; the product only ever runs it on the GPU).
#pragma once

#include <cmath>

#if defined(__CUDACC__)
#define APL_HD __host__ __device__ __forceinline__
#else
#define APL_HD inline
#endif

#define APL_KIND_SNH 0
#define APL_KIND_ARAP 1
#define APL_KIND_SNH_MUSCLE 2
#define APL_KIND_SNH_ARAP 3  // two potentials on the same cells fused into one pass

#define APL_OP_FUN 1
#define APL_OP_GRAD 2
#define APL_OP_HESS_DIAG 4
#define APL_OP_HESS_PROD 8
#define APL_OP_HESS_QUAD 16
#define APL_OP_HESS_OFFD 32  // off-diagonal entries (xy, xz, yz) of the 3x3 vertex blocks of the Hessian (block Jacobi)
#define APL_OP_PSD 64        // modifier: Hessian terms of the Stable Neo-Hookean kinds use the eigenvalue-clamped d2Psi/dF2

namespace apl {

using std::fabs;
using std::sqrt;

// Number of scalars in the per-tet static record: D[9], vol, mu, lambda, (activation[6]).
template <int KIND>
struct RecSize {
    static constexpr int value = (KIND == APL_KIND_SNH_MUSCLE) ? 18 : (KIND == APL_KIND_SNH_ARAP ? 14 : 12);
};

template <typename T>
APL_HD T apl_max(T a, T b) {
    return a > b ? a : b;
}

// Raw hardware approximations (one MUFU each, ~1 ulp, flush-to-zero) for the fp32 Jacobi rotation.
APL_HD float apl_rsqrt_raw(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / std::sqrt(x);
#endif
}
APL_HD float apl_rcp_raw(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}

template <typename T>
APL_HD T apl_rsqrt(T x) {
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(T) == 4) {
        // one Newton step on the hardware approximation: full fp32 accuracy
        float r = apl_rsqrt_raw((float)x);
        r = r * (1.5f - 0.5f * (float)x * r * r);
        return (T)r;
    } else {
        return (T)1 / sqrt(x);
    }
#else
    return (T)1 / std::sqrt(x);
#endif
}

// 1/x: hardware approximation plus one Newton step in fp32 on the device (~1 ulp, no IEEE-division
// slow path), a true division otherwise.
template <typename T>
APL_HD T apl_rcp(T x) {
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(T) == 4) {
        float r = apl_rcp_raw((float)x);
        const float e = fmaf(-(float)x, r, 1.0f);
        return (T)fmaf(r, e, r);
    } else {
        return (T)1 / x;
    }
#else
    return (T)1 / x;
#endif
}

// sqrt(x) for x >= 0 through the reciprocal square root (0 for x = 0)
template <typename T>
APL_HD T apl_sqrt_pos(T x) {
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(T) == 4) return x * apl_rsqrt(apl_max(x, (T)1e-36));
    else return sqrt(x);
#else
    return std::sqrt(x);
#endif
}

// e[a][i] = w[a+1][i] - w[0][i]   (edge differences of a per-corner field)
template <typename T>
APL_HD void edge_diff(const T (*w)[3], T e[3][3]) {
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int i = 0; i < 3; ++i) e[a][i] = w[a + 1][i] - w[0][i];
}

// M[3*i+J] = sum_a e[a][i] * D[3*a+J]      (this is  w^T dhdX  for a per-corner field w)
template <typename T>
APL_HD void edge_outer(const T e[3][3], const T* D, T* M) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int J = 0; J < 3; ++J)
            M[3 * i + J] = e[0][i] * D[J] + e[1][i] * D[3 + J] + e[2][i] * D[6 + J];
}

// out[a][i] (+)= s * (dhdX M^T)[a][i]  with dhdX rows 1..3 = D and row 0 = -(sum of rows 1..3)
template <typename T>
APL_HD void vjp_rows(const T* D, const T* M, T s, T out[4][3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        T r1 = D[0] * M[3 * i] + D[1] * M[3 * i + 1] + D[2] * M[3 * i + 2];
        T r2 = D[3] * M[3 * i] + D[4] * M[3 * i + 1] + D[5] * M[3 * i + 2];
        T r3 = D[6] * M[3 * i] + D[7] * M[3 * i + 1] + D[8] * M[3 * i + 2];
        out[1][i] = s * r1;
        out[2][i] = s * r2;
        out[3][i] = s * r3;
        out[0][i] = -s * (r1 + r2 + r3);
    }
}

// out[a][i] = (dhdX M^T)[a][i] without a scale factor (the caller folds it into M)
template <typename T>
APL_HD void vjp_rows1(const T* D, const T* M, T out[4][3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const T r1 = D[0] * M[3 * i] + D[1] * M[3 * i + 1] + D[2] * M[3 * i + 2];
        const T r2 = D[3] * M[3 * i] + D[4] * M[3 * i + 1] + D[5] * M[3 * i + 2];
        const T r3 = D[6] * M[3 * i] + D[7] * M[3 * i + 1] + D[8] * M[3 * i + 2];
        out[1][i] = r1;
        out[2][i] = r2;
        out[3][i] = r3;
        out[0][i] = -r1 - r2 - r3;
    }
}

// cof(F), J = det F.  C[3*i+j] = (column j of cof)_i, columns f1xf2, f2xf0, f0xf1.
template <typename T>
APL_HD T cofactor(const T* F, T* C) {
    C[0] = F[4] * F[8] - F[7] * F[5];
    C[3] = F[7] * F[2] - F[1] * F[8];
    C[6] = F[1] * F[5] - F[4] * F[2];
    C[1] = F[5] * F[6] - F[8] * F[3];
    C[4] = F[8] * F[0] - F[2] * F[6];
    C[7] = F[2] * F[3] - F[5] * F[0];
    C[2] = F[3] * F[7] - F[6] * F[4];
    C[5] = F[6] * F[1] - F[0] * F[7];
    C[8] = F[0] * F[4] - F[3] * F[1];
    return F[0] * C[0] + F[3] * C[3] + F[6] * C[6];
}

// X(F, dF) = [f1 x p2 - f2 x p1 | f2 x p0 - f0 x p2 | f0 x p1 - f1 x p0]  (columns), the
// directional derivative of cof(F) along dF (func/_hess_prod.py:62-77).
template <typename T>
APL_HD void dcofactor(const T* F, const T* P, T* X) {
#define APL_CR(ax, ay, az, bx, by, bz, ox, oy, oz) \
    ox = (ay) * (bz) - (az) * (by);                \
    oy = (az) * (bx) - (ax) * (bz);                \
    oz = (ax) * (by) - (ay) * (bx);
    T ax, ay, az, bx, by, bz;
    // column 0: f1 x p2 - f2 x p1
    APL_CR(F[1], F[4], F[7], P[2], P[5], P[8], ax, ay, az)
    APL_CR(F[2], F[5], F[8], P[1], P[4], P[7], bx, by, bz)
    X[0] = ax - bx; X[3] = ay - by; X[6] = az - bz;
    // column 1: f2 x p0 - f0 x p2
    APL_CR(F[2], F[5], F[8], P[0], P[3], P[6], ax, ay, az)
    APL_CR(F[0], F[3], F[6], P[2], P[5], P[8], bx, by, bz)
    X[1] = ax - bx; X[4] = ay - by; X[7] = az - bz;
    // column 2: f0 x p1 - f1 x p0
    APL_CR(F[0], F[3], F[6], P[1], P[4], P[7], ax, ay, az)
    APL_CR(F[1], F[4], F[7], P[0], P[3], P[6], bx, by, bz)
    X[2] = ax - bx; X[5] = ay - by; X[8] = az - bz;
#undef APL_CR
}

template <typename T>
APL_HD T ddot9(const T* A, const T* B) {
    T s = A[0] * B[0];
#pragma unroll
    for (int k = 1; k < 9; ++k) s += A[k] * B[k];
    return s;
}

// |dhdX_a|^2 for a = 0..3 (row 0 = -(D0+D1+D2))
template <typename T>
APL_HD void row_norms(const T* D, T n[4]) {
    T r0 = D[0] + D[3] + D[6], r1 = D[1] + D[4] + D[7], r2 = D[2] + D[5] + D[8];
    n[0] = r0 * r0 + r1 * r1 + r2 * r2;
    n[1] = D[0] * D[0] + D[1] * D[1] + D[2] * D[2];
    n[2] = D[3] * D[3] + D[4] * D[4] + D[5] * D[5];
    n[3] = D[6] * D[6] + D[7] * D[7] + D[8] * D[8];
}

// ------------------------------------------------------------------------------------------
// 3x3 SVD, rotation-variant convention of warp/math/_rotation.py:9-13: F = U diag(s) V^T with
// U, V proper rotations, s0 >= s1 >= |s2|, s2 carrying the sign of det F.
// Cyclic Jacobi on F^T F for V, then modified Gram-Schmidt on F V for U and the singular values
// (column norms of F V are more accurate than square roots of the eigenvalues).
// ------------------------------------------------------------------------------------------
template <typename T>
APL_HD void jacobi_rotate(T* A, T* V, int p, int q) {
    // A symmetric, stored fully (row-major 3x3).  Annihilates A[p][q] with the rotation
    // tan(theta) = t = sgn(d) 2 a_pq / (|d| + sqrt(d^2 + 4 a_pq^2)),  d = a_qq - a_pp  (|theta| <= pi/4).
    const T apq = A[3 * p + q];
    const T app = A[3 * p + p], aqq = A[3 * q + q];
    const int r = 3 - p - q;  // the untouched index
    T c, s, t;
    if constexpr (sizeof(T) == 4) {
        // Branch-free: an off-diagonal entry below rounding level relative to the diagonal (or so small
        // that b*b underflows) gets t = 0, i.e. the identity rotation.  The angle only steers
        // convergence, so approximate reciprocals are fine; c gets one Newton step so that
        // c^2 + s^2 = 1 to fp32 accuracy and V stays orthonormal.
        const float d = aqq - app, b = 2.0f * apq;
        const float h2 = d * d + b * b;
        const float h = h2 * apl_rsqrt_raw(fmaxf(h2, 1e-36f));
        float tt = b * apl_rcp_raw(fabsf(d) + h + 1e-30f);
        tt = d < 0.0f ? -tt : tt;
        const bool negligible = !(apq * apq > 1e-16f * fabsf(app * aqq)) || !(fabsf(apq) > 1e-18f);
        tt = negligible ? 0.0f : tt;
        const float x = tt * tt + 1.0f;
        float cc = apl_rsqrt_raw(x);
        cc = cc * (1.5f - 0.5f * x * cc * cc);
        t = tt; c = cc; s = tt * cc;
        A[3 * p + q] = A[3 * q + p] = 0.0f;
    } else {
        // Skip when the off-diagonal entry is below rounding level relative to the diagonal (or
        // absolutely tiny): rotating further cannot improve the result, and b*b would underflow.
        if (!(fabs(apq) > (T)1e-150) || !(apq * apq > (T)1e-34 * fabs(app * aqq))) {
            A[3 * p + q] = A[3 * q + p] = (T)0;
            return;
        }
        const T d = aqq - app, b = (T)2 * apq;
        const T h = sqrt(d * d + b * b);
        t = b / (fabs(d) + h);
        if (d < (T)0) t = -t;
        c = (T)1 / sqrt(t * t + (T)1);
        s = t * c;
        A[3 * p + q] = A[3 * q + p] = (T)0;
    }
    const T arp = A[3 * r + p], arq = A[3 * r + q];
    A[3 * p + p] = app - t * apq;
    A[3 * q + q] = aqq + t * apq;
    A[3 * r + p] = A[3 * p + r] = c * arp - s * arq;
    A[3 * r + q] = A[3 * q + r] = s * arp + c * arq;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const T vp = V[3 * k + p], vq = V[3 * k + q];
        V[3 * k + p] = c * vp - s * vq;
        V[3 * k + q] = s * vp + c * vq;
    }
}

// true when every lane of the warp has converged (host: this element has converged)
APL_HD bool apl_all_done(bool done) {
#if defined(__CUDA_ARCH__)
    return __all_sync(__activemask(), done);
#else
    return done;
#endif
}

template <typename T>
APL_HD void swap_cols_neg(T* A, T* V, int a, int b, bool doit) {
    // conditional (select-based, branch-free) swap of eigenpairs a <-> b, negating one column to keep
    // det V = +1
    const T ea = A[3 * a + a], eb = A[3 * b + b];
    A[3 * a + a] = doit ? eb : ea;
    A[3 * b + b] = doit ? ea : eb;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const T va = V[3 * k + a], vb = V[3 * k + b];
        V[3 * k + a] = doit ? vb : va;
        V[3 * k + b] = doit ? -va : vb;
    }
}

template <typename T>
APL_HD void svd3_rv(const T* F, T* U, T* sig, T* V) {
    T A[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            A[3 * i + j] = F[i] * F[j] + F[3 + i] * F[3 + j] + F[6 + i] * F[6 + j];
    V[0] = 1; V[1] = 0; V[2] = 0;
    V[3] = 0; V[4] = 1; V[5] = 0;
    V[6] = 0; V[7] = 0; V[8] = 1;
    // Cyclic Jacobi converges quadratically; stop when the off-diagonal mass is below rounding level
    // relative to the trace (checked per warp so that the loop stays convergent).
    constexpr int kMaxSweeps = (sizeof(T) == 4) ? 6 : 10;
    const T tr = A[0] + A[4] + A[8];
    const T eps = (sizeof(T) == 4) ? (T)2.4e-7 : (T)8.9e-16;
    const T tol = eps * eps * tr * tr;
#pragma unroll 1
    for (int sweep = 0; sweep < kMaxSweeps; ++sweep) {
        jacobi_rotate(A, V, 0, 1);
        jacobi_rotate(A, V, 0, 2);
        jacobi_rotate(A, V, 1, 2);
        const T off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
        if (apl_all_done(off <= tol)) break;
    }
    // sort eigenvalues descending (det V stays +1)
    swap_cols_neg(A, V, 0, 1, A[0] < A[4]);
    swap_cols_neg(A, V, 0, 2, A[0] < A[8]);
    swap_cols_neg(A, V, 1, 2, A[4] < A[8]);
    // B = F V
    T B[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            B[3 * i + j] = F[3 * i] * V[j] + F[3 * i + 1] * V[3 + j] + F[3 * i + 2] * V[6 + j];
    const T tiny = (sizeof(T) == 4) ? (T)1e-18 : (T)1e-150;
    // u0
    T n0 = B[0] * B[0] + B[3] * B[3] + B[6] * B[6];
    T u0x, u0y, u0z;
    if (n0 > tiny) {
        T r = apl_rsqrt(n0);
        u0x = B[0] * r; u0y = B[3] * r; u0z = B[6] * r;
        sig[0] = n0 * r;
    } else {
        u0x = 1; u0y = 0; u0z = 0;
        sig[0] = 0;
    }
    // u1 = normalize(b1 - (u0.b1) u0)
    T d = u0x * B[1] + u0y * B[4] + u0z * B[7];
    T b1x = B[1] - d * u0x, b1y = B[4] - d * u0y, b1z = B[7] - d * u0z;
    T n1 = b1x * b1x + b1y * b1y + b1z * b1z;
    T u1x, u1y, u1z;
    if (n1 > tiny) {
        T r = apl_rsqrt(n1);
        u1x = b1x * r; u1y = b1y * r; u1z = b1z * r;
        sig[1] = n1 * r;
    } else {
        // any unit vector orthogonal to u0
        if (fabs(u0x) < (T)0.6) { u1x = 0; u1y = -u0z; u1z = u0y; }
        else { u1x = -u0y; u1y = u0x; u1z = 0; }
        T r = apl_rsqrt(u1x * u1x + u1y * u1y + u1z * u1z);
        u1x *= r; u1y *= r; u1z *= r;
        sig[1] = 0;
    }
    // u2 = u0 x u1, s2 = u2 . b2 (signed)
    T u2x = u0y * u1z - u0z * u1y, u2y = u0z * u1x - u0x * u1z, u2z = u0x * u1y - u0y * u1x;
    sig[2] = u2x * B[2] + u2y * B[5] + u2z * B[8];
    U[0] = u0x; U[3] = u0y; U[6] = u0z;
    U[1] = u1x; U[4] = u1y; U[7] = u1z;
    U[2] = u2x; U[5] = u2y; U[8] = u2z;
}

// ------------------------------------------------------------------------------------------
// Energy terms.  Each writes (ACC = false) or adds (ACC = true) its contribution, already multiplied
// by the cell volume and clamped exactly like the reference kernels, to
//   psi   : Psi * vol                                  (OP_FUN)
//   P     : vol * first Piola-Kirchhoff stress (3x3)   (OP_GRAD;      nodal forces = dhdX P^T)
//   dg    : max(diag * vol, 0) per entry               (OP_HESS_DIAG)
//   M     : vol * dP[dF]  (3x3)                        (OP_HESS_PROD; H p = dhdX M^T)
//   quad  : max(p^T H p * vol, 0)                      (OP_HESS_QUAD)
// Keeping P and M as 3x3 matrices lets several energies on the same cell share one dhdX-product.
// ------------------------------------------------------------------------------------------
template <bool ACC, typename T>
APL_HD void put(T& dst, T v) {
    if constexpr (ACC) dst += v;
    else dst = v;
}

// Negative part of the spectrum of d2Psi/dF2 of the Stable Neo-Hookean energy, analytically (opt-in APL_OP_PSD; a
// superset of the reference, whose only safeguards are the clamps of warp/fem/_base.py:317-320,379-380).
//   H_F = mu 1 + lambda g g^T + c3 H_J,   g = vec cof F,  H_J dF = X(F, dF),  c3 = -mu + lambda (J - 1).
// With F = U diag(s) V^T (U, V rotations) and dF = U A V^T, H_J acts on A pair-wise: on the off-diagonal pair (i, j)
// with third index k it maps (A_ij, A_ji) to -s_k (A_ji, A_ij), so
//   twist  modes  U (E_ij - E_ji)/sqrt2 V^T   have eigenvalue  mu + c3 s_k,
//   flip   modes  U (E_ij + E_ji)/sqrt2 V^T   have eigenvalue  mu - c3 s_k      (both orthogonal to g),
// and the three scaling modes U diag(e) V^T are the eigenvectors of the 3x3 matrix
//   A_s = mu 1 + c3 [[0, s2, s1], [s2, 0, s0], [s1, s0, 0]] + lambda gh gh^T,   gh = (s1 s2, s0 s2, s0 s1).
// SnhNeg holds the frame and min(eigenvalue, 0) of all nine modes; H_F^+ = H_F - sum_i neg_i q_i q_i^T.
template <typename T>
struct SnhNeg {
    T U[9], V[9];
    T tw[3], fl[3];   // min(eigenvalue, 0) of the twist / flip mode about axis k
    T sc[3], S[9];    // min(eigenvalue, 0) of the scaling modes, eigenvectors = columns of S
};

template <typename T>
APL_HD void snh_negative_modes(const T* F, T mu, T la, T c3, SnhNeg<T>& n) {
    T sg[3];
    svd3_rv(F, n.U, sg, n.V);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const T a = mu + c3 * sg[k], b = mu - c3 * sg[k];
        n.tw[k] = a < (T)0 ? a : (T)0;
        n.fl[k] = b < (T)0 ? b : (T)0;
    }
    const T gh[3] = {sg[1] * sg[2], sg[0] * sg[2], sg[0] * sg[1]};
    T A[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) A[3 * i + j] = la * gh[i] * gh[j] + (i == j ? mu : c3 * sg[3 - i - j]);
    n.S[0] = 1; n.S[1] = 0; n.S[2] = 0; n.S[3] = 0; n.S[4] = 1; n.S[5] = 0; n.S[6] = 0; n.S[7] = 0; n.S[8] = 1;
    constexpr int kMaxSweeps = (sizeof(T) == 4) ? 6 : 10;
    const T scale = fabs(A[0]) + fabs(A[4]) + fabs(A[8]) + fabs(A[1]) + fabs(A[2]) + fabs(A[5]);
    const T eps = (sizeof(T) == 4) ? (T)2.4e-7 : (T)8.9e-16;
    const T tol = eps * eps * scale * scale;
#pragma unroll 1
    for (int sweep = 0; sweep < kMaxSweeps; ++sweep) {
        jacobi_rotate(A, n.S, 0, 1);
        jacobi_rotate(A, n.S, 0, 2);
        jacobi_rotate(A, n.S, 1, 2);
        const T off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
        if (apl_all_done(off <= tol)) break;
    }
#pragma unroll
    for (int m = 0; m < 3; ++m) n.sc[m] = A[4 * m] < (T)0 ? A[4 * m] : (T)0;
}

// C = sum_i neg_i <q_i, dF> q_i (3x3, row-major) and r = sum_i neg_i <q_i, dF>^2 (<= 0)
template <typename T>
APL_HD T snh_negative_apply(const SnhNeg<T>& n, const T* dF, T* C) {
    T A[9], W[9];   // A = U^T dF V
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) W[3 * i + j] = n.U[i] * dF[j] + n.U[3 + i] * dF[3 + j] + n.U[6 + i] * dF[6 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) A[3 * i + j] = W[3 * i] * n.V[j] + W[3 * i + 1] * n.V[3 + j] + W[3 * i + 2] * n.V[6 + j];
    T Ch[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    T r = (T)0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int i = (k + 1) % 3, j = (k + 2) % 3;
        const T t = (T)0.5 * (A[3 * i + j] - A[3 * j + i]), f = (T)0.5 * (A[3 * i + j] + A[3 * j + i]);
        // <q,dF> q = (A_ij -+ A_ji)/2 (E_ij -+ E_ji);  <q,dF>^2 = 2 t^2 resp. 2 f^2
        Ch[3 * i + j] = n.tw[k] * t + n.fl[k] * f;
        Ch[3 * j + i] = -n.tw[k] * t + n.fl[k] * f;
        r += (T)2 * (n.tw[k] * t * t + n.fl[k] * f * f);
    }
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        const T c = n.S[m] * A[0] + n.S[3 + m] * A[4] + n.S[6 + m] * A[8];
        const T w = n.sc[m] * c;
        Ch[0] += w * n.S[m]; Ch[4] += w * n.S[3 + m]; Ch[8] += w * n.S[6 + m];
        r += w * c;
    }
    // C = U Ch V^T
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) W[3 * i + j] = n.U[3 * i] * Ch[j] + n.U[3 * i + 1] * Ch[3 + j] + n.U[3 * i + 2] * Ch[6 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[3 * i + j] = W[3 * i] * n.V[3 * j] + W[3 * i + 1] * n.V[3 * j + 1] + W[3 * i + 2] * n.V[3 * j + 2];
    return r;
}

// K = sum_i neg_i (q_i d)(q_i d)^T for a row d of dhdX, packed [xx, yy, zz, xy, xz, yz] (negative semi-definite)
template <typename T>
APL_HD void snh_negative_block(const SnhNeg<T>& n, const T* d, T* K) {
    const T dh[3] = {n.V[0] * d[0] + n.V[3] * d[1] + n.V[6] * d[2], n.V[1] * d[0] + n.V[4] * d[1] + n.V[7] * d[2],
                     n.V[2] * d[0] + n.V[5] * d[1] + n.V[8] * d[2]};   // V^T d
    T Kh[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int i = (k + 1) % 3, j = (k + 2) % 3;
        // twist: z_i = d_j / sqrt2, z_j = -d_i / sqrt2;  flip: z_i = d_j / sqrt2, z_j = d_i / sqrt2
        const T a = (T)0.5 * (n.tw[k] + n.fl[k]), b = (T)0.5 * (n.fl[k] - n.tw[k]);
        Kh[3 * i + i] += a * dh[j] * dh[j];
        Kh[3 * j + j] += a * dh[i] * dh[i];
        Kh[3 * i + j] += b * dh[i] * dh[j];
        Kh[3 * j + i] += b * dh[i] * dh[j];
    }
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        const T z[3] = {n.S[m] * dh[0], n.S[3 + m] * dh[1], n.S[6 + m] * dh[2]};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) Kh[3 * i + j] += n.sc[m] * z[i] * z[j];
    }
    // K = U Kh U^T
    T W[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) W[3 * i + j] = n.U[3 * i] * Kh[j] + n.U[3 * i + 1] * Kh[3 + j] + n.U[3 * i + 2] * Kh[6 + j];
    auto kij = [&](int i, int j) { return W[3 * i] * n.U[3 * j] + W[3 * i + 1] * n.U[3 * j + 1] + W[3 * i + 2] * n.U[3 * j + 2]; };
    K[0] = kij(0, 0); K[1] = kij(1, 1); K[2] = kij(2, 2); K[3] = kij(0, 1); K[4] = kij(0, 2); K[5] = kij(1, 2);
}

// Stable Neo-Hookean (warp/fem/_stable_neo_hookean.py:17-103) on F with dhdX block D.
// od (OP_HESS_OFFD): off-diagonal entries (xy, xz, yz) of the vertex blocks  vol (mu |D_a|^2 1 + lambda w_a w_a^T),
// w_a = row a of dhdX cof(F)^T (the h6 term has no block-diagonal part: func/_hess_diag.py:69-72).
template <typename T, int OPS, bool ACC>
APL_HD void snh_terms(const T* F, const T* dF, const T* D, T vol, T mu, T la, T& psi, T& quad, T* P, T* M,
                      T dg[4][3], T od[4][3]) {
    constexpr bool kFun = (OPS & APL_OP_FUN) != 0, kGrad = (OPS & APL_OP_GRAD) != 0;
    constexpr bool kDiag = (OPS & APL_OP_HESS_DIAG) != 0, kProd = (OPS & APL_OP_HESS_PROD) != 0;
    constexpr bool kQuad = (OPS & APL_OP_HESS_QUAD) != 0, kOffd = (OPS & APL_OP_HESS_OFFD) != 0;
    constexpr bool kPsd = (OPS & APL_OP_PSD) != 0 && (kDiag || kOffd || kProd || kQuad);
    T C[9];
    const T J = cofactor(F, C);
    const T Jm1 = J - (T)1;
    const T c3 = -mu + la * Jm1;
    SnhNeg<T> neg;
    if constexpr (kPsd) snh_negative_modes(F, mu, la, c3, neg);
    if constexpr (kFun) {
        const T I2 = ddot9(F, F);
        put<ACC>(psi, vol * ((T)0.5 * mu * (I2 - (T)3) - mu * Jm1 + (T)0.5 * la * Jm1 * Jm1));
    }
    if constexpr (kGrad) {
        const T a = vol * mu, b = vol * c3;
#pragma unroll
        for (int k = 0; k < 9; ++k) put<ACC>(P[k], a * F[k] + b * C[k]);
    }
    if constexpr (kDiag || kOffd) {
        T W[4][3], n[4];
        vjp_rows(D, C, (T)1, W);
        row_norms(D, n);
        const T Dr[4][3] = {{-(D[0] + D[3] + D[6]), -(D[1] + D[4] + D[7]), -(D[2] + D[5] + D[8])},
                            {D[0], D[1], D[2]}, {D[3], D[4], D[5]}, {D[6], D[7], D[8]}};
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            T K[6] = {0, 0, 0, 0, 0, 0};
            if constexpr (kPsd) snh_negative_block(neg, Dr[a], K);
            if constexpr (kDiag) {
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    put<ACC>(dg[a][i], apl_max(vol * (la * W[a][i] * W[a][i] + mu * n[a] - K[i]), (T)0));
            }
            if constexpr (kOffd) {
                put<ACC>(od[a][0], vol * (la * W[a][0] * W[a][1] - K[3]));
                put<ACC>(od[a][1], vol * (la * W[a][0] * W[a][2] - K[4]));
                put<ACC>(od[a][2], vol * (la * W[a][1] * W[a][2] - K[5]));
            }
        }
    }
    if constexpr (kProd || kQuad) {
        T X[9];
        dcofactor(F, dF, X);
        const T s = ddot9(C, dF);
        T Cn[9];
        T rn = (T)0;
        if constexpr (kPsd) rn = snh_negative_apply(neg, dF, Cn);
        if constexpr (kProd) {
            const T a = vol * la * s, b = vol * mu, c = vol * c3;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                T m = a * C[k] + b * dF[k] + c * X[k];
                if constexpr (kPsd) m -= vol * Cn[k];
                put<ACC>(M[k], m);
            }
        }
        if constexpr (kQuad) {
            const T q = la * s * s + mu * ddot9(dF, dF) + c3 * ddot9(dF, X) - rn;
            put<ACC>(quad, apl_max(vol * q, (T)0));
        }
    }
}

// ------------------------------------------------------------------------------------------
// ARAP kinematics without an iterative SVD.
//
// Everything the ARAP terms need from F = U diag(s) V^T (rotation-variant convention) is
//   R   = U V^T                                   the rotation of the polar decomposition F = R S,
//   s   = the three singular values               (energy),
//   Lam = sum_k lambda_k w_k w_k^T                with the twist axes w = (v2, v0, v1) of the modes
//         Q0, Q1, Q2 (func/_misc.py:56-70: Q_k = R [w_k]_x / sqrt2) and their clamped rates lambda_k
//         (func/_misc.py:31-43), i.e. Lam = f(tr(S) 1 - S) with f(m) = 2 / max(m, 2).
// With S^2 = C = F^T F, Cayley-Hamilton gives S and R from the invariants of S alone:
//   S = (I1 I3 1 + (I1^2 - I2) C - C^2) / (I1 I2 - I3),   R = (I1 F - F S + cof F) / I2,
//   I1 = s0 + s1 + s2, I2 = s0 s1 + s1 s2 + s2 s0, I3 = s0 s1 s2 = det F,
// the two larger singular values come from the trigonometric eigenvalues of C and the smallest one
// (signed) from det F / (s0 s1), and f(M) is the Newton interpolation polynomial of f on the three
// eigenvalues of M = I1 1 - S (exact for a 3x3 matrix function), with the isolated eigenvalue taken
// first so that the divided differences of the two close ones are never amplified.  Errors stay at a few
// ulps of F while I2 is not small; strongly compressed / inverted elements (I2 <= 0.5 s0^2) take the
// Jacobi SVD above instead (rare, warp-divergent).  Lam is returned as [xx, yy, zz, xy, xz, yz].
// ------------------------------------------------------------------------------------------
#ifndef APL_ARAP_CLOSED_FORM
#define APL_ARAP_CLOSED_FORM 1
#endif

template <typename T>
APL_HD T apl_acos(T x) {
#if defined(__CUDA_ARCH__)
    if constexpr (sizeof(T) == 4) return acosf(x);
    else return acos(x);
#else
    return std::acos(x);
#endif
}
// sin and cos on [0, pi/3] (the eigenvalue angle).  fp32: Taylor polynomials in x^2 (truncation error
// < 4e-9 on the interval, no range reduction, no slow path -- the same arithmetic on host and device);
// fp64: the library functions.
template <typename T>
APL_HD void apl_sincos(T x, T& sn, T& cs) {
    if constexpr (sizeof(T) == 4) {
        const float x2 = x * x;
        float c = -2.7557319e-7f;                 // -1/10!
        c = c * x2 + 2.4801587e-5f;               //  1/8!
        c = c * x2 - 1.3888889e-3f;               // -1/6!
        c = c * x2 + 4.1666668e-2f;               //  1/4!
        c = c * x2 - 0.5f;
        cs = c * x2 + 1.0f;
        float q = -2.5052108e-8f;                 // -1/11!
        q = q * x2 + 2.7557319e-6f;               //  1/9!
        q = q * x2 - 1.9841270e-4f;               // -1/7!
        q = q * x2 + 8.3333338e-3f;               //  1/5!
        q = q * x2 - 1.6666667e-1f;               // -1/3!
        sn = x + x * x2 * q;
    } else {
#if defined(__CUDA_ARCH__)
        sincos(x, &sn, &cs);
#else
        sn = std::sin(x);
        cs = std::cos(x);
#endif
    }
}

template <typename T>
APL_HD void polar_twist_svd(const T* F, T* R, T* L, T* sg) {
    T U[9], V[9];
    svd3_rv(F, U, sg, V);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            R[3 * i + j] = U[3 * i] * V[3 * j] + U[3 * i + 1] * V[3 * j + 1] + U[3 * i + 2] * V[3 * j + 2];
    const T two = (T)2;
    const T l2 = two / apl_max(sg[0] + sg[1], two);   // axis v2
    const T l0 = two / apl_max(sg[1] + sg[2], two);   // axis v0
    const T l1 = two / apl_max(sg[2] + sg[0], two);   // axis v1
    L[0] = l0 * V[0] * V[0] + l1 * V[1] * V[1] + l2 * V[2] * V[2];
    L[1] = l0 * V[3] * V[3] + l1 * V[4] * V[4] + l2 * V[5] * V[5];
    L[2] = l0 * V[6] * V[6] + l1 * V[7] * V[7] + l2 * V[8] * V[8];
    L[3] = l0 * V[0] * V[3] + l1 * V[1] * V[4] + l2 * V[2] * V[5];
    L[4] = l0 * V[0] * V[6] + l1 * V[1] * V[7] + l2 * V[2] * V[8];
    L[5] = l0 * V[3] * V[6] + l1 * V[4] * V[7] + l2 * V[5] * V[8];
}

template <typename T>
APL_HD void polar_twist(const T* F, T* R, T* L, T* sg) {
#if !APL_ARAP_CLOSED_FORM
    polar_twist_svd(F, R, L, sg);
#else
    // C = F^T F as [xx, yy, zz, xy, xz, yz]
    T C[6];
    C[0] = F[0] * F[0] + F[3] * F[3] + F[6] * F[6];
    C[1] = F[1] * F[1] + F[4] * F[4] + F[7] * F[7];
    C[2] = F[2] * F[2] + F[5] * F[5] + F[8] * F[8];
    C[3] = F[0] * F[1] + F[3] * F[4] + F[6] * F[7];
    C[4] = F[0] * F[2] + F[3] * F[5] + F[6] * F[8];
    C[5] = F[1] * F[2] + F[4] * F[5] + F[7] * F[8];
    T cof[9];
    const T J = cofactor(F, cof);
    // two largest eigenvalues of C (trigonometric form on the deviator)
    const T q = (C[0] + C[1] + C[2]) * (T)(1.0 / 3.0);
    const T b0 = C[0] - q, b1 = C[1] - q, b2 = C[2] - q;
    const T p2 = (b0 * b0 + b1 * b1 + b2 * b2 + (T)2 * (C[3] * C[3] + C[4] * C[4] + C[5] * C[5])) * (T)(1.0 / 6.0);
    const T detB = b0 * (b1 * b2 - C[5] * C[5]) - C[3] * (C[3] * b2 - C[5] * C[4]) + C[4] * (C[3] * C[5] - b1 * C[4]);
    const T tiny = (sizeof(T) == 4) ? (T)1e-30 : (T)1e-280;
    const T rs = apl_rsqrt(apl_max(p2, tiny));   // 1 / p
    const T p = p2 * rs;
    T r = (T)0.5 * detB * rs * rs * rs;          // det(B / p) / 2
    r = r > (T)1 ? (T)1 : (r < (T)-1 ? (T)-1 : r);   // (a NaN from an underflow maps to -1: p is ~0 then)
    const T phi = apl_acos(r) * (T)(1.0 / 3.0);
    T sn, cs;
    apl_sincos(phi, sn, cs);                     // phi in [0, pi/3]
    const T e0 = q + (T)2 * p * cs;
    const T e1 = q + p * ((T)1.7320508075688772935 * sn - cs);   // q + 2 p cos(phi - 2 pi / 3)
    const T s0 = apl_sqrt_pos(apl_max(e0, (T)0)), s1 = apl_sqrt_pos(apl_max(e1, (T)0));
    const T s01 = s0 * s1;
    const T s2 = s01 > tiny ? J * apl_rcp(s01) : (T)0;
    const T I1 = s0 + s1 + s2, I2 = s01 + s2 * (s0 + s1);
    if (!(I2 > (T)0.5 * e0)) {   // strongly compressed / inverted / degenerate (also NaN): robust path
        polar_twist_svd(F, R, L, sg);
        return;
    }
    sg[0] = s0; sg[1] = s1; sg[2] = s2;
    // S = (I1 I3 + (I1^2 - I2) C - C^2) / (I1 I2 - I3)
    const T inv_den = apl_rcp(I1 * I2 - J);
    const T kc = (I1 * I1 - I2) * inv_den, k0 = I1 * J * inv_den;
    T S[6];
    S[0] = k0 + kc * C[0] - inv_den * (C[0] * C[0] + C[3] * C[3] + C[4] * C[4]);
    S[1] = k0 + kc * C[1] - inv_den * (C[3] * C[3] + C[1] * C[1] + C[5] * C[5]);
    S[2] = k0 + kc * C[2] - inv_den * (C[4] * C[4] + C[5] * C[5] + C[2] * C[2]);
    S[3] = kc * C[3] - inv_den * (C[0] * C[3] + C[3] * C[1] + C[4] * C[5]);
    S[4] = kc * C[4] - inv_den * (C[0] * C[4] + C[3] * C[5] + C[4] * C[2]);
    S[5] = kc * C[5] - inv_den * (C[3] * C[4] + C[1] * C[5] + C[5] * C[2]);
    // R = (I1 F - F S + cof F) / I2 = F (M / I2) + cof F / I2 with M = I1 - S
    const T inv_I2 = apl_rcp(I2);
    const T N0 = (I1 - S[0]) * inv_I2, N1 = (I1 - S[1]) * inv_I2, N2 = (I1 - S[2]) * inv_I2;
    const T N3 = -S[3] * inv_I2, N4 = -S[4] * inv_I2, N5 = -S[5] * inv_I2;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const T f0 = F[3 * i], f1 = F[3 * i + 1], f2 = F[3 * i + 2];
        R[3 * i + 0] = cof[3 * i + 0] * inv_I2 + f0 * N0 + f1 * N3 + f2 * N4;
        R[3 * i + 1] = cof[3 * i + 1] * inv_I2 + f0 * N3 + f1 * N1 + f2 * N5;
        R[3 * i + 2] = cof[3 * i + 2] * inv_I2 + f0 * N4 + f1 * N5 + f2 * N2;
    }
    // Lam = f(M), M = I1 - S, eigenvalues m = (s1 + s2) <= (s0 + s2) <= (s0 + s1), f(m) = 2 / max(m, 2).
    // Newton form on the nodes (a, b, c) = (isolated end, other end, middle):
    //   f(M) = f_a + f[a,b] (M - a) + f[a,b,c] (M - a)(M - b)
    const T two = (T)2;
    const T m_lo = s1 + s2, m_mid = s0 + s2, m_hi = s0 + s1;
    const bool hi_first = (m_hi - m_mid) >= (m_mid - m_lo);
    const T ma = hi_first ? m_hi : m_lo, mb = hi_first ? m_lo : m_hi, mc = m_mid;
    const T fa = two * apl_rcp(apl_max(ma, two)), fb = two * apl_rcp(apl_max(mb, two));
    const T fc = two * apl_rcp(apl_max(mc, two));
    const T dab = mb - ma, dbc = mc - mb, dac = mc - ma;
    const T fab = fabs(dab) > tiny ? (fb - fa) * apl_rcp(dab) : (T)0;
    const T fbc = fabs(dbc) > tiny ? (fc - fb) * apl_rcp(dbc) : (T)0;
    const T fabc = fabs(dac) > tiny ? (fbc - fab) * apl_rcp(dac) : (T)0;
    // A = M - a, B = M - b (symmetric, same off-diagonals = -S_offdiag)
    const T A0 = I1 - S[0] - ma, A1 = I1 - S[1] - ma, A2 = I1 - S[2] - ma;
    const T B0 = I1 - S[0] - mb, B1 = I1 - S[1] - mb, B2 = I1 - S[2] - mb;
    const T o3 = -S[3], o4 = -S[4], o5 = -S[5];
    L[0] = fa + fab * A0 + fabc * (A0 * B0 + o3 * o3 + o4 * o4);
    L[1] = fa + fab * A1 + fabc * (o3 * o3 + A1 * B1 + o5 * o5);
    L[2] = fa + fab * A2 + fabc * (o4 * o4 + o5 * o5 + A2 * B2);
    L[3] = fab * o3 + fabc * (A0 * o3 + o3 * B1 + o4 * o5);
    L[4] = fab * o4 + fabc * (A0 * o4 + o3 * o5 + o4 * B2);
    L[5] = fab * o5 + fabc * (o3 * o4 + A1 * o5 + o5 * B2);
#endif
}

// y = Lam x for the packed symmetric Lam
template <typename T>
APL_HD void sym_mul(const T* L, T x0, T x1, T x2, T& y0, T& y1, T& y2) {
    y0 = L[0] * x0 + L[3] * x1 + L[4] * x2;
    y1 = L[3] * x0 + L[1] * x1 + L[5] * x2;
    y2 = L[4] * x0 + L[5] * x1 + L[2] * x2;
}

// ARAP (warp/fem/_arap.py:17-76) on F with dhdX block D, in terms of (R, Lam, s):
//   Psi  = mu/2 |F - R|^2,   P = mu (F - R),
//   <Q_k, dF> = w_k . a / sqrt2 with a = axial vector of (R^T dF) - (R^T dF)^T, hence
//   p^T H p = mu (|dF|^2 - a^T Lam a / 2),   H p = mu (dF - R [b]_x),  b = Lam a / 2  (row i of R [b]_x = r_i x b),
//   (dhdX Q_k^T)[a][i] = w_k . (D_a x r_i) / sqrt2, hence diag[a][i] = mu (|D_a|^2 - z^T Lam z / 2), z = D_a x r_i.
template <typename T, int OPS, bool ACC>
APL_HD void arap_terms(const T* F, const T* dF, const T* D, T vol, T mu, T& psi, T& quad, T* P, T* M, T dg[4][3],
                       T od[4][3]) {
    constexpr bool kFun = (OPS & APL_OP_FUN) != 0, kGrad = (OPS & APL_OP_GRAD) != 0;
    constexpr bool kDiag = (OPS & APL_OP_HESS_DIAG) != 0, kProd = (OPS & APL_OP_HESS_PROD) != 0;
    constexpr bool kQuad = (OPS & APL_OP_HESS_QUAD) != 0, kOffd = (OPS & APL_OP_HESS_OFFD) != 0;
    T R[9], L[6], sg[3];
    polar_twist(F, R, L, sg);
    if constexpr (kFun || kGrad) {
        // Psi = mu/2 |F - R|^2 as the reference writes it (_arap.py:17-24); more accurate than sum (s_i - 1)^2
        // from the trigonometric singular values, whose individual errors only cancel in symmetric functions
        T E[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) E[k] = F[k] - R[k];
        if constexpr (kFun) put<ACC>(psi, vol * (T)0.5 * mu * ddot9(E, E));
        if constexpr (kGrad) {
            const T a = vol * mu;
#pragma unroll
            for (int k = 0; k < 9; ++k) put<ACC>(P[k], a * E[k]);
        }
    }
    if constexpr (kDiag || kOffd) {
        // vertex block (i, j) = mu vol (|D_a|^2 delta_ij - z_i^T Lam z_j / 2),  z_i = D_a x r_i
        T n[4];
        row_norms(D, n);
        const T Dr[4][3] = {{-(D[0] + D[3] + D[6]), -(D[1] + D[4] + D[7]), -(D[2] + D[5] + D[8])},
                            {D[0], D[1], D[2]}, {D[3], D[4], D[5]}, {D[6], D[7], D[8]}};
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            T z[3][3], y[3][3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                z[i][0] = Dr[a][1] * R[3 * i + 2] - Dr[a][2] * R[3 * i + 1];
                z[i][1] = Dr[a][2] * R[3 * i + 0] - Dr[a][0] * R[3 * i + 2];
                z[i][2] = Dr[a][0] * R[3 * i + 1] - Dr[a][1] * R[3 * i + 0];
                sym_mul(L, z[i][0], z[i][1], z[i][2], y[i][0], y[i][1], y[i][2]);
            }
            if constexpr (kDiag) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const T h4 = (T)0.5 * (z[i][0] * y[i][0] + z[i][1] * y[i][1] + z[i][2] * y[i][2]);
                    put<ACC>(dg[a][i], apl_max(vol * mu * (n[a] - h4), (T)0));
                }
            }
            if constexpr (kOffd) {
                const T s = (T)-0.5 * vol * mu;
                put<ACC>(od[a][0], s * (z[0][0] * y[1][0] + z[0][1] * y[1][1] + z[0][2] * y[1][2]));
                put<ACC>(od[a][1], s * (z[0][0] * y[2][0] + z[0][1] * y[2][1] + z[0][2] * y[2][2]));
                put<ACC>(od[a][2], s * (z[1][0] * y[2][0] + z[1][1] * y[2][1] + z[1][2] * y[2][2]));
            }
        }
    }
    if constexpr (kProd || kQuad) {
        // A = R^T dF; a = (A21 - A12, A02 - A20, A10 - A01)
        T A[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) A[3 * i + j] = R[i] * dF[j] + R[3 + i] * dF[3 + j] + R[6 + i] * dF[6 + j];
        const T a0 = A[7] - A[5], a1 = A[2] - A[6], a2 = A[3] - A[1];
        T b0, b1, b2;
        sym_mul(L, a0, a1, a2, b0, b1, b2);
        b0 *= (T)0.5; b1 *= (T)0.5; b2 *= (T)0.5;
        if constexpr (kQuad) {
            const T q = ddot9(dF, dF) - (a0 * b0 + a1 * b1 + a2 * b2);
            put<ACC>(quad, apl_max(vol * mu * q, (T)0));
        }
        if constexpr (kProd) {
            const T s = vol * mu;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const T r0 = R[3 * i], r1 = R[3 * i + 1], r2 = R[3 * i + 2];
                put<ACC>(M[3 * i + 0], s * (dF[3 * i + 0] - (r1 * b2 - r2 * b1)));
                put<ACC>(M[3 * i + 1], s * (dF[3 * i + 1] - (r2 * b0 - r0 * b2)));
                put<ACC>(M[3 * i + 2], s * (dF[3 * i + 2] - (r0 * b1 - r1 * b0)));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// One tetrahedron, every requested operator, every energy of the cell's record:
//   SNH / ARAP / SNH_MUSCLE : rec = [D(9), vol, mu, lambda, (activation 6)]
//   SNH_ARAP (two potentials sharing the cell, WarpModel sum fused into one pass):
//                             rec = [D(9), vol_snh, mu_snh, lambda_snh, vol_arap, mu_arap]
// Outputs: psi, quad (scalars), g = dhdX P^T, dg, hp = dhdX M^T, all volume-weighted and clamped as in
// warp/fem/_base.py:243-383 (diag per entry, quad per cell -- per potential for the fused kind).
// OP_HESS_OFFD (never together with OP_HESS_PROD) returns the off-diagonal entries (xy, xz, yz) of the 3x3 vertex
// blocks in hp; with dg they are the block-Jacobi preconditioner.
// ------------------------------------------------------------------------------------------
template <typename T, int KIND, int OPS>
APL_HD void elem_eval(const T* rec, const T (*uc)[3], const T (*pc)[3], T& psi, T& quad, T g[4][3], T dg[4][3],
                      T hp[4][3]) {
    constexpr bool kGrad = (OPS & APL_OP_GRAD) != 0;
    constexpr bool kProd = (OPS & APL_OP_HESS_PROD) != 0;
    constexpr bool kQuad = (OPS & APL_OP_HESS_QUAD) != 0;
    constexpr bool kNeedP = kProd || kQuad;

    T D[9];
    T F[9];
    {
        T e[3][3];
        edge_diff(uc, e);
        if constexpr (KIND == APL_KIND_SNH_MUSCLE) {
            // A = I + sym(a); D <- D A; F <- G = F A = A + sum_a e_a (x) (D A)_a
            const T A[9] = {(T)1 + rec[12], rec[15], rec[16],
                            rec[15], (T)1 + rec[13], rec[17],
                            rec[16], rec[17], (T)1 + rec[14]};
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int J = 0; J < 3; ++J)
                    D[3 * a + J] = rec[3 * a] * A[J] + rec[3 * a + 1] * A[3 + J] + rec[3 * a + 2] * A[6 + J];
            edge_outer(e, D, F);
#pragma unroll
            for (int k = 0; k < 9; ++k) F[k] += A[k];
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) D[k] = rec[k];
            edge_outer(e, D, F);
            F[0] += (T)1; F[4] += (T)1; F[8] += (T)1;
        }
    }
    T dF[9];
    if constexpr (kNeedP) {
        T e[3][3];
        edge_diff(pc, e);
        edge_outer(e, D, dF);
    }
    static_assert(!((OPS & APL_OP_HESS_OFFD) && kProd), "hess_offd and hess_prod share an output");
    T P[9], M[9];
    if constexpr (KIND == APL_KIND_SNH || KIND == APL_KIND_SNH_MUSCLE) {
        snh_terms<T, OPS, false>(F, dF, D, rec[9], rec[10], rec[11], psi, quad, P, M, dg, hp);
    } else if constexpr (KIND == APL_KIND_ARAP) {
        arap_terms<T, OPS, false>(F, dF, D, rec[9], rec[10], psi, quad, P, M, dg, hp);
    } else {
        static_assert(KIND == APL_KIND_SNH_ARAP, "unknown energy kind");
        snh_terms<T, OPS, false>(F, dF, D, rec[9], rec[10], rec[11], psi, quad, P, M, dg, hp);
        arap_terms<T, OPS, true>(F, dF, D, rec[12], rec[13], psi, quad, P, M, dg, hp);
    }
    if constexpr (kGrad) vjp_rows1(D, P, g);
    if constexpr (kProd) vjp_rows1(D, M, hp);
}

// ------------------------------------------------------------------------------------------
// Packed fp32 pairs.  Blackwell issues two fp32 lanes per instruction (SASS FFMA2 / FADD2 / FMUL2, PTX *.f32x2) with an
// optional SCALAR BROADCAST operand; the fp32 FLOP rate is the same as with scalar FFMA (measured: profiles/
// r2i_ubench_f32x2.txt, 72.4 vs 69.7 TFLOP/s) but every pair costs ONE issue slot -- and these kernels are bound by
// instruction issue and the shared-memory pipe, not by the FMA pipe (26-41 % busy).  The two lanes carry the SAME
// operation on two quantities of ONE tet that share the dhdX factor:
//   {F, dF} = {u, p}^T dhdX          (deformation gradient of the state and of the direction)
//   {g, Hp} = dhdX {P, M}^T          (nodal forces of the stress and of its tangent)
// so no data has to be shuffled between threads or registers: the gathered vertex rows are stored INTERLEAVED in shared
// memory ([ux uy px py | uz pz - -]) and the slots hold [gx gy hx hy | gz hz], both produced / consumed as they come.
// The 3x3 stress / tangent algebra in between stays scalar.
// ------------------------------------------------------------------------------------------
struct f32x2 {
    float lo, hi;
};
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ unsigned long long p2_bits(f32x2 a) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.lo), "f"(a.hi));
    return r;
}
__device__ __forceinline__ f32x2 p2_from(unsigned long long r) {
    f32x2 a;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a.lo), "=f"(a.hi) : "l"(r));
    return a;
}
#endif
APL_HD f32x2 p2_make(float a, float b) { f32x2 r; r.lo = a; r.hi = b; return r; }
APL_HD f32x2 p2_add(f32x2 a, f32x2 b) {
#if defined(__CUDA_ARCH__)
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(p2_bits(a)), "l"(p2_bits(b)));
    return p2_from(r);
#else
    return p2_make(a.lo + b.lo, a.hi + b.hi);
#endif
}
APL_HD f32x2 p2_sub(f32x2 a, f32x2 b) {
#if defined(__CUDA_ARCH__)
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(p2_bits(a)), "l"(p2_bits(b)));
    return p2_from(r);
#else
    return p2_make(a.lo - b.lo, a.hi - b.hi);
#endif
}
// a * s (s broadcast to both lanes)
APL_HD f32x2 p2_mul_s(f32x2 a, float s) {
#if defined(__CUDA_ARCH__)
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(p2_bits(a)), "l"(p2_bits(p2_make(s, s))));
    return p2_from(r);
#else
    return p2_make(a.lo * s, a.hi * s);
#endif
}
// a * s + c (s broadcast to both lanes)
APL_HD f32x2 p2_fma_s(f32x2 a, float s, f32x2 c) {
#if defined(__CUDA_ARCH__)
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(p2_bits(a)), "l"(p2_bits(p2_make(s, s))), "l"(p2_bits(c)));
    return p2_from(r);
#else
    return p2_make(std::fma(a.lo, s, c.lo), std::fma(a.hi, s, c.hi));
#endif
}

// One tetrahedron, fp32, operator set that reads BOTH nodal fields (hess_prod / hess_quad), packed:
//   u01[c] = (u_x, u_y), p01[c] = (p_x, p_y), upz[c] = (u_z, p_z) of corner c  ->  {F, dF} in 27 packed instructions;
// PACK_OUT (operator sets {grad, hess_prod} (+ fun)): g01[c] = (g_x, g_y), h01[c] = (Hp_x, Hp_y), gzhz[c] = (g_z, Hp_z)
// from {P, M} in 36 packed instructions; otherwise the scalar outputs g, dg, hp of elem_eval.
// Same arithmetic as elem_eval, element for element.
template <int KIND, int OPS, bool PACK_OUT>
APL_HD void elem_eval_packed(const float* rec, const f32x2 u01[4], const f32x2 p01[4], const f32x2 upz[4], float& psi,
                             float& quad, f32x2 g01[4], f32x2 h01[4], f32x2 gzhz[4], float g[4][3], float dg[4][3],
                             float hp[4][3]) {
    static_assert((OPS & (APL_OP_HESS_PROD | APL_OP_HESS_QUAD)) != 0, "the packed form pairs the state with the direction");
    static_assert(!PACK_OUT || ((OPS & APL_OP_GRAD) && (OPS & APL_OP_HESS_PROD) &&
                                !(OPS & (APL_OP_HESS_DIAG | APL_OP_HESS_OFFD))), "packed outputs: grad + hess_prod");
    float D[9], A[9];
    if constexpr (KIND == APL_KIND_SNH_MUSCLE) {
        const float Am[9] = {1.0f + rec[12], rec[15], rec[16], rec[15], 1.0f + rec[13], rec[17], rec[16], rec[17], 1.0f + rec[14]};
#pragma unroll
        for (int k = 0; k < 9; ++k) A[k] = Am[k];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int J = 0; J < 3; ++J) D[3 * a + J] = rec[3 * a] * A[J] + rec[3 * a + 1] * A[3 + J] + rec[3 * a + 2] * A[6 + J];
    } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) { D[k] = rec[k]; A[k] = (k % 4 == 0) ? 1.0f : 0.0f; }
    }
    // edge differences of (u_x, u_y), (p_x, p_y), (u_z, p_z)
    f32x2 eu[3], ep[3], ez[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        eu[a] = p2_sub(u01[a + 1], u01[0]);
        ep[a] = p2_sub(p01[a + 1], p01[0]);
        ez[a] = p2_sub(upz[a + 1], upz[0]);
    }
    // {F[J], F[3+J]}, {dF[J], dF[3+J]}, {F[6+J], dF[6+J]} = sum_a e[a] * D[3a+J]
    float F[9], dF[9];
#pragma unroll
    for (int J = 0; J < 3; ++J) {
        const f32x2 f = p2_fma_s(eu[2], D[6 + J], p2_fma_s(eu[1], D[3 + J], p2_mul_s(eu[0], D[J])));
        const f32x2 d = p2_fma_s(ep[2], D[6 + J], p2_fma_s(ep[1], D[3 + J], p2_mul_s(ep[0], D[J])));
        const f32x2 z = p2_fma_s(ez[2], D[6 + J], p2_fma_s(ez[1], D[3 + J], p2_mul_s(ez[0], D[J])));
        F[J] = f.lo + A[J]; F[3 + J] = f.hi + A[3 + J]; F[6 + J] = z.lo + A[6 + J];
        dF[J] = d.lo; dF[3 + J] = d.hi; dF[6 + J] = z.hi;
    }
    float P[9], M[9];
    float ods[4][3];   // (no packed operator set has hess_offd)
    if constexpr (KIND == APL_KIND_SNH || KIND == APL_KIND_SNH_MUSCLE) {
        snh_terms<float, OPS, false>(F, dF, D, rec[9], rec[10], rec[11], psi, quad, P, M, dg, ods);
    } else if constexpr (KIND == APL_KIND_ARAP) {
        arap_terms<float, OPS, false>(F, dF, D, rec[9], rec[10], psi, quad, P, M, dg, ods);
    } else {
        snh_terms<float, OPS, false>(F, dF, D, rec[9], rec[10], rec[11], psi, quad, P, M, dg, ods);
        arap_terms<float, OPS, true>(F, dF, D, rec[12], rec[13], psi, quad, P, M, dg, ods);
    }
    if constexpr (PACK_OUT) {
        // nodal forces: out[a][i] = sum_J D[3(a-1)+J] X[3i+J]; rows (x, y) of P and of M, and row z of both, as pairs
        f32x2 Pp[3], Mp[3], Zp[3];
#pragma unroll
        for (int J = 0; J < 3; ++J) {
            Pp[J] = p2_make(P[J], P[3 + J]);
            Mp[J] = p2_make(M[J], M[3 + J]);
            Zp[J] = p2_make(P[6 + J], M[6 + J]);
        }
#pragma unroll
        for (int a = 1; a < 4; ++a) {
            const float* d = D + 3 * (a - 1);
            g01[a] = p2_fma_s(Pp[2], d[2], p2_fma_s(Pp[1], d[1], p2_mul_s(Pp[0], d[0])));
            h01[a] = p2_fma_s(Mp[2], d[2], p2_fma_s(Mp[1], d[1], p2_mul_s(Mp[0], d[0])));
            gzhz[a] = p2_fma_s(Zp[2], d[2], p2_fma_s(Zp[1], d[1], p2_mul_s(Zp[0], d[0])));
        }
        const f32x2 zero = p2_make(0.0f, 0.0f);
        g01[0] = p2_sub(p2_sub(p2_sub(zero, g01[1]), g01[2]), g01[3]);
        h01[0] = p2_sub(p2_sub(p2_sub(zero, h01[1]), h01[2]), h01[3]);
        gzhz[0] = p2_sub(p2_sub(p2_sub(zero, gzhz[1]), gzhz[2]), gzhz[3]);
    } else {
        if constexpr ((OPS & APL_OP_GRAD) != 0) vjp_rows1(D, P, g);
        if constexpr ((OPS & APL_OP_HESS_PROD) != 0) vjp_rows1(D, M, hp);
    }
}

// ------------------------------------------------------------------------------------------
// Mixed derivative product: d/dq [ grad_u E . p ] per cell, for each material parameter q of the cell's energy
// -- what the reference's inverse problems call `mixed_derivative_prod(state, p)` after the adjoint solve
// (exp/2026/01/28/smas/src/31-inverse-activation-stable-neo-hookean.py:472-487; the method is absent from the
// current src/, so the restatement below is the analytic derivative of the energies of this file).
// With phi = vol <P(F; q), dF> (dF = p^T dhdX):
//   SNH (on G = F A, dG = dF A; A = I for the plain energy):
//     d phi / d mu     = vol (<G, dG> - <cof G, dG>)
//     d phi / d lambda = vol (J_G - 1) <cof G, dG>
//     d phi / d A      = vol (F^T M_G + dF^T P_G),  P_G = mu G + c3 cof G,  M_G = lambda <cof G, dG> cof G + mu dG + c3 X(G, dG)
//     activation a = (a0, a1, a2 | a3 = xy, a4 = xz, a5 = yz):  d/da_k = N_kk (k < 3),  N_01 + N_10,  N_02 + N_20,  N_12 + N_21
//   ARAP:  d phi / d mu = vol <F - R, dF>.
// Outputs: d_mu, d_la (0 for ARAP), d_act[6] (muscle only).  The fused SNH+ARAP record has two sets of materials
// and is not supported here (evaluate the two potentials separately).
// ------------------------------------------------------------------------------------------
template <typename T, int KIND>
APL_HD void elem_mixed(const T* rec, const T (*uc)[3], const T (*pc)[3], T& d_mu, T& d_la, T* d_act) {
    static_assert(KIND != APL_KIND_SNH_ARAP, "mixed derivatives are evaluated per potential");
    T F[9], dF[9];
    {
        T e[3][3];
        edge_diff(uc, e);
        edge_outer(e, rec, F);
        F[0] += (T)1; F[4] += (T)1; F[8] += (T)1;
        edge_diff(pc, e);
        edge_outer(e, rec, dF);
    }
    const T vol = rec[9], mu = rec[10];
    d_mu = d_la = (T)0;
    if constexpr (KIND == APL_KIND_ARAP) {
        T R[9], L[6], sg[3];
        polar_twist(F, R, L, sg);
        T s = (T)0;
#pragma unroll
        for (int k = 0; k < 9; ++k) s += (F[k] - R[k]) * dF[k];
        d_mu = vol * s;
    } else {
        const T la = rec[11];
        T G[9], dG[9];
        if constexpr (KIND == APL_KIND_SNH_MUSCLE) {
            const T A[9] = {(T)1 + rec[12], rec[15], rec[16],
                            rec[15], (T)1 + rec[13], rec[17],
                            rec[16], rec[17], (T)1 + rec[14]};
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    G[3 * i + j] = F[3 * i] * A[j] + F[3 * i + 1] * A[3 + j] + F[3 * i + 2] * A[6 + j];
                    dG[3 * i + j] = dF[3 * i] * A[j] + dF[3 * i + 1] * A[3 + j] + dF[3 * i + 2] * A[6 + j];
                }
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) { G[k] = F[k]; dG[k] = dF[k]; }
        }
        T C[9];
        const T J = cofactor(G, C);
        const T Jm1 = J - (T)1;
        const T cdg = ddot9(C, dG);
        d_mu = vol * (ddot9(G, dG) - cdg);
        d_la = vol * Jm1 * cdg;
        if constexpr (KIND == APL_KIND_SNH_MUSCLE) {
            const T c3 = -mu + la * Jm1;
            T X[9], PG[9], MG[9];
            dcofactor(G, dG, X);
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                PG[k] = mu * G[k] + c3 * C[k];
                MG[k] = la * cdg * C[k] + mu * dG[k] + c3 * X[k];
            }
            T N[9];   // N = F^T M_G + dF^T P_G
#pragma unroll
            for (int m = 0; m < 3; ++m)
#pragma unroll
                for (int n = 0; n < 3; ++n)
                    N[3 * m + n] = F[m] * MG[n] + F[3 + m] * MG[3 + n] + F[6 + m] * MG[6 + n] +
                                   dF[m] * PG[n] + dF[3 + m] * PG[3 + n] + dF[6 + m] * PG[6 + n];
            d_act[0] = vol * N[0]; d_act[1] = vol * N[4]; d_act[2] = vol * N[8];
            d_act[3] = vol * (N[1] + N[3]); d_act[4] = vol * (N[2] + N[6]); d_act[5] = vol * (N[5] + N[7]);
        }
    }
}

}  // namespace apl
