// TEST-ONLY harness: compiles apple_b200/csrc/elem_math.cuh for the HOST so that the per-tet
// closed forms can be checked against the oracle on a machine without a GPU.  It is built by
// tests/test_elem_math_host.py into a scratch directory and is never part of the product library.
#include "../../apple_b200/csrc/elem_math.cuh"

template <typename T, int KIND>
static void run(int n, const T* rec, const T* uc, const T* pc, T* psi, T* quad, T* g, T* dg, T* hp) {
    constexpr int NR = apl::RecSize<KIND>::value;
    constexpr int ALL = APL_OP_FUN | APL_OP_GRAD | APL_OP_HESS_DIAG | APL_OP_HESS_PROD | APL_OP_HESS_QUAD;
    for (int t = 0; t < n; ++t) {
        T G[4][3], DG[4][3], HP[4][3];
        apl::elem_eval<T, KIND, ALL>(rec + (size_t)t * NR, (const T(*)[3])(uc + (size_t)t * 12),
                                     (const T(*)[3])(pc + (size_t)t * 12), psi[t], quad[t], G, DG, HP);
        for (int a = 0; a < 4; ++a)
            for (int i = 0; i < 3; ++i) {
                g[t * 12 + a * 3 + i] = G[a][i];
                dg[t * 12 + a * 3 + i] = DG[a][i];
                hp[t * 12 + a * 3 + i] = HP[a][i];
            }
    }
}

template <typename T>
static void dispatch(int kind, int n, const void* rec, const void* uc, const void* pc, void* psi, void* quad,
                     void* g, void* dg, void* hp) {
#define CALL(K) run<T, K>(n, (const T*)rec, (const T*)uc, (const T*)pc, (T*)psi, (T*)quad, (T*)g, (T*)dg, (T*)hp)
    if (kind == APL_KIND_SNH) CALL(APL_KIND_SNH);
    else if (kind == APL_KIND_ARAP) CALL(APL_KIND_ARAP);
    else if (kind == APL_KIND_SNH_ARAP) CALL(APL_KIND_SNH_ARAP);
    else CALL(APL_KIND_SNH_MUSCLE);
#undef CALL
}

extern "C" void elem_eval_host(int kind, int is_f64, int n, const void* rec, const void* uc, const void* pc,
                               void* psi, void* quad, void* g, void* dg, void* hp) {
    if (is_f64) dispatch<double>(kind, n, rec, uc, pc, psi, quad, g, dg, hp);
    else dispatch<float>(kind, n, rec, uc, pc, psi, quad, g, dg, hp);
}

extern "C" void svd3_host(int is_f64, int n, const void* F, void* U, void* s, void* V) {
    for (int t = 0; t < n; ++t) {
        if (is_f64) apl::svd3_rv<double>((const double*)F + 9 * t, (double*)U + 9 * t, (double*)s + 3 * t, (double*)V + 9 * t);
        else apl::svd3_rv<float>((const float*)F + 9 * t, (float*)U + 9 * t, (float*)s + 3 * t, (float*)V + 9 * t);
    }
}

// R (rotation-variant polar factor), Lam = sum_k lambda_k w_k w_k^T packed [xx,yy,zz,xy,xz,yz], singular values;
// path 0 = product path (closed form with SVD fallback), 1 = Jacobi SVD only
extern "C" void polar_twist_host(int is_f64, int path, int n, const void* F, void* R, void* L, void* s) {
    for (int t = 0; t < n; ++t) {
        if (is_f64) {
            const double* f = (const double*)F + 9 * t;
            if (path == 0) apl::polar_twist<double>(f, (double*)R + 9 * t, (double*)L + 6 * t, (double*)s + 3 * t);
            else apl::polar_twist_svd<double>(f, (double*)R + 9 * t, (double*)L + 6 * t, (double*)s + 3 * t);
        } else {
            const float* f = (const float*)F + 9 * t;
            if (path == 0) apl::polar_twist<float>(f, (float*)R + 9 * t, (float*)L + 6 * t, (float*)s + 3 * t);
            else apl::polar_twist_svd<float>(f, (float*)R + 9 * t, (float*)L + 6 * t, (float*)s + 3 * t);
        }
    }
}

// per-cell mixed derivative products d(grad E . p)/d(mu, lambda, activation); kind 0 SNH, 1 ARAP, 2 muscle
extern "C" void elem_mixed_host(int kind, int is_f64, int n, const void* rec, const void* uc, const void* pc, void* d_mu,
                                void* d_la, void* d_act) {
    auto go = [&](auto zero) {
        using T = decltype(zero);
        const int NR = kind == 2 ? 18 : 12;
        for (int t = 0; t < n; ++t) {
            const T* r = (const T*)rec + (size_t)t * NR;
            const T(*u)[3] = (const T(*)[3])((const T*)uc + (size_t)t * 12);
            const T(*p)[3] = (const T(*)[3])((const T*)pc + (size_t)t * 12);
            T a[6] = {0, 0, 0, 0, 0, 0};
            T& m = ((T*)d_mu)[t];
            T& l = ((T*)d_la)[t];
            if (kind == 0) apl::elem_mixed<T, APL_KIND_SNH>(r, u, p, m, l, a);
            else if (kind == 1) apl::elem_mixed<T, APL_KIND_ARAP>(r, u, p, m, l, a);
            else apl::elem_mixed<T, APL_KIND_SNH_MUSCLE>(r, u, p, m, l, a);
            for (int k = 0; k < 6; ++k) ((T*)d_act)[6 * t + k] = a[k];
        }
    };
    if (is_f64) go(0.0); else go(0.0f);
}

// Opt-in supersets (block Jacobi, PSD projection).  sel 0: DIAG|OFFD, 1: DIAG|OFFD|PSD, 2: PROD|QUAD|PSD.
// dg (n,4,3) = block diagonals, hp (n,4,3) = block off-diagonals (xy, xz, yz) for sel 0/1; hp = H+ p, quad = max(p H+ p, 0) for sel 2.
template <typename T, int KIND, int OPS>
static void run_ops(int n, const T* rec, const T* uc, const T* pc, T* quad, T* dg, T* hp) {
    constexpr int NR = apl::RecSize<KIND>::value;
    for (int t = 0; t < n; ++t) {
        T G[4][3], DG[4][3] = {}, HP[4][3] = {};
        T psi = 0, q = 0;
        apl::elem_eval<T, KIND, OPS>(rec + (size_t)t * NR, (const T(*)[3])(uc + (size_t)t * 12),
                                     (const T(*)[3])(pc + (size_t)t * 12), psi, q, G, DG, HP);
        quad[t] = q;
        for (int a = 0; a < 4; ++a)
            for (int i = 0; i < 3; ++i) {
                dg[t * 12 + a * 3 + i] = DG[a][i];
                hp[t * 12 + a * 3 + i] = HP[a][i];
            }
    }
}

extern "C" void elem_eval_ops_host(int kind, int is_f64, int sel, int n, const void* rec, const void* uc, const void* pc,
                                   void* quad, void* dg, void* hp) {
    auto go = [&](auto zero) {
        using T = decltype(zero);
#define CALL(K, O) run_ops<T, K, O>(n, (const T*)rec, (const T*)uc, (const T*)pc, (T*)quad, (T*)dg, (T*)hp)
#define SEL(K)                                                                     \
    if (sel == 0) CALL(K, APL_OP_HESS_DIAG | APL_OP_HESS_OFFD);                    \
    else if (sel == 1) CALL(K, APL_OP_HESS_DIAG | APL_OP_HESS_OFFD | APL_OP_PSD);  \
    else CALL(K, APL_OP_HESS_PROD | APL_OP_HESS_QUAD | APL_OP_PSD)
        if (kind == APL_KIND_SNH) { SEL(APL_KIND_SNH); }
        else if (kind == APL_KIND_ARAP) { SEL(APL_KIND_ARAP); }
        else if (kind == APL_KIND_SNH_ARAP) { SEL(APL_KIND_SNH_ARAP); }
        else { SEL(APL_KIND_SNH_MUSCLE); }
#undef SEL
#undef CALL
    };
    if (is_f64) go(0.0); else go(0.0f);
}
