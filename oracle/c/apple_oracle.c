/* CPU oracle in plain C (TEST / BASELINE INFRASTRUCTURE ONLY -- nothing under apple_b200/ links it).
 *
 * A literal restatement, one C function per Warp kernel, of the reference's FEM operators for linear
 * tetrahedra (paths relative to /root/reference/src/liblaf/apple):
 *   fun_kernel        warp/fem/_base.py:243-263      grad_kernel       warp/fem/_base.py:265-291
 *   hess_diag_kernel  warp/fem/_base.py:293-324      hess_prod_kernel  warp/fem/_base.py:326-351
 *   hess_quad_kernel  warp/fem/_base.py:353-383
 * with the device functions of warp/fem/func/*.py and the energies of
 *   warp/fem/_stable_neo_hookean.py:17-103, warp/fem/_arap.py:17-76,
 *   warp/fem/_stable_neo_hookean_muscle.py:18-114.
 * Like the reference, every operator is its own pass over the cells that recomputes F, and nodal
 * results are scattered with atomic adds (wp.atomic_add -> a compare-and-swap loop).  The reference's own CPU
 * path (Warp device="cpu") runs kernels on ONE thread; here the cell loop is split over all host
 * threads (pthreads; this image's gcc has no libgomp), so this is the stronger baseline.  wp.svd3 is replaced by a cyclic-Jacobi SVD in the same
 * rotation-variant convention (U, V proper rotations; warp/math/_rotation.py:9-13).
 * The ARAP hess_prod is the mathematically correct product (see oracle/fem.py).
 * It is validated against oracle/fem.py (numpy) by tests/test_oracle_c.py.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* -DORACLE_F32 builds the same restatement in single precision (liboracle_f32.so): the reference is dtype-generic
 * (JAX x64 flag, warp/fem/_base.py:100-104), and bench.py's reference arm runs in the dtype of the GPU arm. */
#ifdef ORACLE_F32
typedef float real;
typedef uint32_t real_bits;
#define REAL_TINY 1e-30f      /* |a_pq| below this: skip the Jacobi rotation */
#define REAL_SMALL 1e-18f     /* column norm below this: degenerate direction */
#define REAL_OFF_TOL 1e-13f   /* off-diagonal mass relative to trace^2 at convergence */
#else
typedef double real;
typedef uint64_t real_bits;
#define REAL_TINY 1e-300
#define REAL_SMALL 1e-150
#define REAL_OFF_TOL 1e-34
#endif

enum { KIND_SNH = 0, KIND_ARAP = 1, KIND_MUSCLE = 2 };

typedef struct {
    int kind;
    int64_t n_cells;
    const int32_t* cells; /* (T,4) */
    const real* dhdX;     /* (T,4,3) */
    const real* dV;       /* (T) */
    const real* mu;
    const real* la;
    const real* act; /* (T,6) */
} pot_t;

/* ---- func/_deformation.py ---- */
static void gather4(const real* u, const int32_t* c, real uc[4][3]) { /* func/_misc.py:22-28 */
    for (int a = 0; a < 4; ++a)
        for (int i = 0; i < 3; ++i) uc[a][i] = u[3 * (int64_t)c[a] + i];
}
static void defgrad(real uc[4][3], const real* D, real F[3][3]) { /* :15-19  F = u^T dhdX + I */
    for (int i = 0; i < 3; ++i)
        for (int J = 0; J < 3; ++J) {
            real s = (i == J) ? 1.0 : 0.0;
            for (int a = 0; a < 4; ++a) s += uc[a][i] * D[3 * a + J];
            F[i][J] = s;
        }
}
static void jvp(real pc[4][3], const real* D, real dF[3][3]) { /* :22-25 */
    for (int i = 0; i < 3; ++i)
        for (int J = 0; J < 3; ++J) {
            real s = 0;
            for (int a = 0; a < 4; ++a) s += pc[a][i] * D[3 * a + J];
            dF[i][J] = s;
        }
}
static void vjp(const real* D, real M[3][3], real out[4][3]) { /* :28-31  dhdX M^T */
    for (int a = 0; a < 4; ++a)
        for (int i = 0; i < 3; ++i) {
            real s = 0;
            for (int J = 0; J < 3; ++J) s += D[3 * a + J] * M[i][J];
            out[a][i] = s;
        }
}
static real ddot33(real A[3][3], real B[3][3]) {
    real s = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) s += A[i][j] * B[i][j];
    return s;
}
static real ddot43(real A[4][3], real B[4][3]) {
    real s = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 3; ++j) s += A[i][j] * B[i][j];
    return s;
}
static void cross(const real a[3], const real b[3], real o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
static void col(real F[3][3], int j, real o[3]) { o[0] = F[0][j]; o[1] = F[1][j]; o[2] = F[2][j]; }
static real det3(real F[3][3]) { /* func/_identity.py:31-39 */
    return F[0][0] * (F[1][1] * F[2][2] - F[1][2] * F[2][1]) - F[0][1] * (F[1][0] * F[2][2] - F[1][2] * F[2][0]) +
           F[0][2] * (F[1][0] * F[2][1] - F[1][1] * F[2][0]);
}
static void g3(real F[3][3], real C[3][3]) { /* func/_gradient.py:30-40 */
    real f0[3], f1[3], f2[3], c[3];
    col(F, 0, f0); col(F, 1, f1); col(F, 2, f2);
    cross(f1, f2, c); for (int i = 0; i < 3; ++i) C[i][0] = c[i];
    cross(f2, f0, c); for (int i = 0; i < 3; ++i) C[i][1] = c[i];
    cross(f0, f1, c); for (int i = 0; i < 3; ++i) C[i][2] = c[i];
}
static void h6mat(real F[3][3], real P[3][3], real X[3][3]) { /* func/_hess_prod.py:62-77 */
    real f0[3], f1[3], f2[3], p0[3], p1[3], p2[3], a[3], b[3];
    col(F, 0, f0); col(F, 1, f1); col(F, 2, f2); col(P, 0, p0); col(P, 1, p1); col(P, 2, p2);
    cross(f1, p2, a); cross(f2, p1, b); for (int i = 0; i < 3; ++i) X[i][0] = a[i] - b[i];
    cross(f2, p0, a); cross(f0, p2, b); for (int i = 0; i < 3; ++i) X[i][1] = a[i] - b[i];
    cross(f0, p1, a); cross(f1, p0, b); for (int i = 0; i < 3; ++i) X[i][2] = a[i] - b[i];
}

/* ---- rotation-variant SVD (stands in for wp.svd3, math/_rotation.py:9-13) ---- */
static void svd3_rv(real F[3][3], real U[3][3], real s[3], real V[3][3]) {
    real A[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            A[i][j] = 0;
            for (int k = 0; k < 3; ++k) A[i][j] += F[k][i] * F[k][j];
            V[i][j] = (i == j);
        }
    for (int sweep = 0; sweep < 30; ++sweep) {
        real off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
        real tr = A[0][0] + A[1][1] + A[2][2];
        if (off <= REAL_OFF_TOL * tr * tr) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabs(A[p][q]) < REAL_TINY) continue;
                real theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                real t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                real c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 3; ++k) { /* A <- A J */
                    real akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - sn * akq; A[k][q] = sn * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) { /* A <- J^T A */
                    real apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - sn * aqk; A[q][k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    real vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - sn * vkq; V[k][q] = sn * vkp + c * vkq;
                }
            }
    }
    /* sort descending, keep det V = +1 */
    for (int a = 0; a < 2; ++a)
        for (int b = a + 1; b < 3; ++b)
            if (A[a][a] < A[b][b]) {
                real t = A[a][a]; A[a][a] = A[b][b]; A[b][b] = t;
                for (int k = 0; k < 3; ++k) { real va = V[k][a]; V[k][a] = V[k][b]; V[k][b] = -va; }
            }
    real B[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { B[i][j] = 0; for (int k = 0; k < 3; ++k) B[i][j] += F[i][k] * V[k][j]; }
    real b0[3], b1[3], b2[3], u0[3], u1[3], u2[3];
    col(B, 0, b0); col(B, 1, b1); col(B, 2, b2);
    real n0 = sqrt(b0[0] * b0[0] + b0[1] * b0[1] + b0[2] * b0[2]);
    if (n0 > REAL_SMALL) { for (int i = 0; i < 3; ++i) u0[i] = b0[i] / n0; } else { u0[0] = 1; u0[1] = u0[2] = 0; n0 = 0; }
    real d = u0[0] * b1[0] + u0[1] * b1[1] + u0[2] * b1[2];
    for (int i = 0; i < 3; ++i) b1[i] -= d * u0[i];
    real n1 = sqrt(b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2]);
    if (n1 > REAL_SMALL) { for (int i = 0; i < 3; ++i) u1[i] = b1[i] / n1; }
    else {
        real e[3] = {0, 0, 0}; e[fabs(u0[0]) < 0.6 ? 0 : 1] = 1; cross(u0, e, u1);
        real n = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
        for (int i = 0; i < 3; ++i) u1[i] /= n;
        n1 = 0;
    }
    cross(u0, u1, u2);
    s[0] = n0; s[1] = n1; s[2] = u2[0] * b2[0] + u2[1] * b2[1] + u2[2] * b2[2];
    for (int i = 0; i < 3; ++i) { U[i][0] = u0[i]; U[i][1] = u1[i]; U[i][2] = u2[i]; }
}
static void lambdas(const real s[3], real l[3]) { /* func/_misc.py:31-43, clamp=True */
    l[0] = 2.0 / fmax(s[0] + s[1], 2.0); l[1] = 2.0 / fmax(s[1] + s[2], 2.0); l[2] = 2.0 / fmax(s[2] + s[0], 2.0);
}
static void Qs(real U[3][3], real V[3][3], real Q[3][3][3]) { /* func/_misc.py:56-70 */
    static const int um[3] = {1, 1, 0}, vn[3] = {0, 2, 2}, un[3] = {0, 2, 2}, vm[3] = {1, 1, 0};
    const real r = 1.0 / sqrt(2.0);
    for (int k = 0; k < 3; ++k)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Q[k][i][j] = (U[i][um[k]] * V[j][vn[k]] - U[i][un[k]] * V[j][vm[k]]) * r;
}

/* per-cell context: F (or G = F A), effective dhdX (dhdX or dhdX A), coefficients */
typedef struct { real F[3][3]; real D[12]; real mu, la, vol; real A[3][3]; } ctx_t;

static void setup(const pot_t* P, int64_t c, const real* u, ctx_t* x) {
    real uc[4][3];
    gather4(u, P->cells + 4 * c, uc);
    const real* D = P->dhdX + 12 * c;
    real F[3][3];
    defgrad(uc, D, F);
    x->mu = P->mu[c]; x->la = P->la ? P->la[c] : 0; x->vol = P->dV[c];
    if (P->kind == KIND_MUSCLE) { /* func/_misc.py:46-53; _stable_neo_hookean_muscle.py: G = F A, dhdX A */
        const real* a = P->act + 6 * c;
        real A[3][3] = {{1 + a[0], a[3], a[4]}, {a[3], 1 + a[1], a[5]}, {a[4], a[5], 1 + a[2]}};
        memcpy(x->A, A, sizeof(A));
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) { x->F[i][j] = 0; for (int k = 0; k < 3; ++k) x->F[i][j] += F[i][k] * A[k][j]; }
        for (int r = 0; r < 4; ++r)
            for (int j = 0; j < 3; ++j) { x->D[3 * r + j] = 0; for (int k = 0; k < 3; ++k) x->D[3 * r + j] += D[3 * r + k] * A[k][j]; }
    } else {
        memcpy(x->F, F, sizeof(F));
        memcpy(x->D, D, 12 * sizeof(real));
    }
}

static real energy_density(const pot_t* P, ctx_t* x) {
    if (P->kind == KIND_ARAP) { /* _arap.py:17-21 */
        real U[3][3], s[3], V[3][3], e = 0;
        svd3_rv(x->F, U, s, V);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                real R = U[i][0] * V[j][0] + U[i][1] * V[j][1] + U[i][2] * V[j][2];
                e += (x->F[i][j] - R) * (x->F[i][j] - R);
            }
        return 0.5 * x->mu * e;
    }
    real J = det3(x->F), I2 = ddot33(x->F, x->F); /* _stable_neo_hookean.py:17-27 */
    return 0.5 * x->mu * (I2 - 3.0) - x->mu * (J - 1.0) + 0.5 * x->la * (J - 1.0) * (J - 1.0);
}

static void first_pk_force(const pot_t* P, ctx_t* x, real out[4][3]) { /* vjp(dhdX, P) */
    real Pk[3][3];
    if (P->kind == KIND_ARAP) { /* _arap.py:24-28 */
        real U[3][3], s[3], V[3][3];
        svd3_rv(x->F, U, s, V);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                Pk[i][j] = x->mu * (x->F[i][j] - (U[i][0] * V[j][0] + U[i][1] * V[j][1] + U[i][2] * V[j][2]));
    } else { /* _stable_neo_hookean.py:30-39 (muscle: times A^T is absorbed by using dhdX A in the vjp) */
        real C[3][3], J = det3(x->F);
        g3(x->F, C);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Pk[i][j] = 0.5 * x->mu * 2.0 * x->F[i][j] + (-x->mu + x->la * (J - 1.0)) * C[i][j];
    }
    vjp(x->D, Pk, out);
}

static void hess_diag_cell(const pot_t* P, ctx_t* x, real out[4][3]) {
    real n[4];
    for (int a = 0; a < 4; ++a) n[a] = x->D[3 * a] * x->D[3 * a] + x->D[3 * a + 1] * x->D[3 * a + 1] + x->D[3 * a + 2] * x->D[3 * a + 2];
    if (P->kind == KIND_ARAP) { /* _arap.py:31-40 */
        real U[3][3], s[3], V[3][3], l[3], Q[3][3][3], W[4][3];
        svd3_rv(x->F, U, s, V); lambdas(s, l); Qs(U, V, Q);
        for (int a = 0; a < 4; ++a) for (int i = 0; i < 3; ++i) out[a][i] = 2.0 * n[a];
        for (int k = 0; k < 3; ++k) {
            vjp(x->D, Q[k], W);
            for (int a = 0; a < 4; ++a) for (int i = 0; i < 3; ++i) out[a][i] += -2.0 * l[k] * W[a][i] * W[a][i];
        }
        for (int a = 0; a < 4; ++a) for (int i = 0; i < 3; ++i) out[a][i] *= 0.5 * x->mu;
        return;
    }
    real C[3][3], W[4][3]; /* _stable_neo_hookean.py:42-65; h6_diag == 0 */
    g3(x->F, C); vjp(x->D, C, W);
    for (int a = 0; a < 4; ++a) for (int i = 0; i < 3; ++i) out[a][i] = x->la * W[a][i] * W[a][i] + 0.5 * x->mu * 2.0 * n[a];
}

static void hess_prod_cell(const pot_t* P, ctx_t* x, real pc[4][3], real out[4][3], real* quad) {
    real dF[3][3];
    jvp(pc, x->D, dF);
    real M[3][3];
    if (P->kind == KIND_ARAP) { /* _arap.py:43-58 (correct argument order), :61-76 */
        real U[3][3], s[3], V[3][3], l[3], Q[3][3][3];
        svd3_rv(x->F, U, s, V); lambdas(s, l); Qs(U, V, Q);
        real q = 2.0 * ddot33(dF, dF);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M[i][j] = 2.0 * dF[i][j];
        for (int k = 0; k < 3; ++k) {
            real c = ddot33(Q[k], dF);
            q += -2.0 * l[k] * c * c;
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M[i][j] += -2.0 * l[k] * c * Q[k][i][j];
        }
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M[i][j] *= 0.5 * x->mu;
        *quad = 0.5 * x->mu * q;
    } else { /* _stable_neo_hookean.py:68-103 */
        real C[3][3], X[3][3], J = det3(x->F);
        g3(x->F, C); h6mat(x->F, dF, X);
        real c3 = -x->mu + x->la * (J - 1.0), sdot = ddot33(C, dF);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) M[i][j] = x->la * sdot * C[i][j] + 0.5 * x->mu * 2.0 * dF[i][j] + c3 * X[i][j];
        *quad = x->la * sdot * sdot + 0.5 * x->mu * 2.0 * ddot33(dF, dF) + c3 * ddot33(dF, X);
    }
    vjp(x->D, M, out);
    (void)ddot43;
}

static void atomic_add(real* addr, real val) { /* wp.atomic_add on the CPU */
    real_bits* p = (real_bits*)addr;
    real_bits old = __atomic_load_n(p, __ATOMIC_RELAXED), neu;
    do {
        real f;
        memcpy(&f, &old, sizeof(real));
        f += val;
        memcpy(&neu, &f, sizeof(real));
    } while (!__atomic_compare_exchange_n(p, &old, neu, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}

static void scatter(real* out, const int32_t* c, real v[4][3], real scale, int clamp) {
    for (int a = 0; a < 4; ++a)
        for (int i = 0; i < 3; ++i) {
            real val = v[a][i] * scale;
            if (clamp && val < 0) val = 0;
            atomic_add(out + 3 * (int64_t)c[a] + i, val);
        }
}

/* ---- the five kernels (accumulate into caller-zeroed outputs) ---- */
static pot_t make(int kind, int64_t T, const int32_t* cells, const real* dhdX, const real* dV, const real* mu,
                  const real* la, const real* act) {
    pot_t P = {kind, T, cells, dhdX, dV, mu, la, act};
    return P;
}

/* ---- minimal parallel-for over cells ---- */
typedef struct {
    const pot_t* P; const real* u; const real* p; real* out; int op; int64_t lo, hi; double acc;
} job_t;
static int g_threads = 0;
static int n_threads(void) {
    if (g_threads <= 0) {
        const char* e = getenv("ORACLE_NUM_THREADS");
        long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
        g_threads = n < 1 ? 1 : (n > 256 ? 256 : (int)n);
    }
    return g_threads;
}
static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    const pot_t* P = j->P;
    double acc = 0;
    for (int64_t c = j->lo; c < j->hi; ++c) {
        ctx_t x; real pc[4][3], g[4][3], q;
        setup(P, c, j->u, &x);
        switch (j->op) {
            case 0: acc += energy_density(P, &x) * x.vol; break;
            case 1: first_pk_force(P, &x, g); scatter(j->out, P->cells + 4 * c, g, x.vol, 0); break;
            case 2: hess_diag_cell(P, &x, g); scatter(j->out, P->cells + 4 * c, g, x.vol, 1); break;
            case 3: gather4(j->p, P->cells + 4 * c, pc); hess_prod_cell(P, &x, pc, g, &q);
                    scatter(j->out, P->cells + 4 * c, g, x.vol, 0); break;
            default: gather4(j->p, P->cells + 4 * c, pc); hess_prod_cell(P, &x, pc, g, &q);
                     q *= x.vol; acc += q > 0 ? q : 0; /* _base.py:379-380 */ break;
        }
    }
    j->acc = acc;
    return 0;
}
static double run(const pot_t* P, int op, const real* u, const real* p, real* out) {
    int nt = n_threads();
    if (P->n_cells < 4096) nt = 1;
    pthread_t th[256];
    job_t jobs[256];
    for (int t = 0; t < nt; ++t) {
        jobs[t] = (job_t){P, u, p, out, op, P->n_cells * t / nt, P->n_cells * (t + 1) / nt, 0.0};
        if (t > 0) pthread_create(&th[t], 0, worker, &jobs[t]);
    }
    worker(&jobs[0]);
    double acc = jobs[0].acc;
    for (int t = 1; t < nt; ++t) { pthread_join(th[t], 0); acc += jobs[t].acc; }
    return acc;
}

double oracle_fun(int kind, int64_t T, const int32_t* cells, const real* dhdX, const real* dV, const real* mu,
                  const real* la, const real* act, const real* u) {
    pot_t P = make(kind, T, cells, dhdX, dV, mu, la, act);
    return run(&P, 0, u, 0, 0);
}
void oracle_grad(int kind, int64_t T, const int32_t* cells, const real* dhdX, const real* dV, const real* mu,
                 const real* la, const real* act, const real* u, real* out) {
    pot_t P = make(kind, T, cells, dhdX, dV, mu, la, act);
    run(&P, 1, u, 0, out);
}
void oracle_hess_diag(int kind, int64_t T, const int32_t* cells, const real* dhdX, const real* dV, const real* mu,
                      const real* la, const real* act, const real* u, real* out) {
    pot_t P = make(kind, T, cells, dhdX, dV, mu, la, act);
    run(&P, 2, u, 0, out);
}
void oracle_hess_prod(int kind, int64_t T, const int32_t* cells, const real* dhdX, const real* dV, const real* mu,
                      const real* la, const real* act, const real* u, const real* p, real* out) {
    pot_t P = make(kind, T, cells, dhdX, dV, mu, la, act);
    run(&P, 3, u, p, out);
}
double oracle_hess_quad(int kind, int64_t T, const int32_t* cells, const real* dhdX, const real* dV, const real* mu,
                        const real* la, const real* act, const real* u, const real* p) {
    pot_t P = make(kind, T, cells, dhdX, dV, mu, la, act);
    return run(&P, 4, u, p, 0);
}
int oracle_num_threads(void) { return n_threads(); }
