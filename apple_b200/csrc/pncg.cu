// Fused PNCG iteration: vector kernels with device-resident scalars + the phase driver.
//
// What is restated (the optimizer, liblaf.peach.optim.PNCG, is external to the reference; these are
// the recurrences of the reference's own PNCG-like benchmark, benches/bench_pncg_branching_backends.py):
//   preconditioner fix-up   :407-410     Dai-Kou beta          :663-679    beta reset     :288-289
//   direction/descent guard :290-303     initial step          :606-610    Armijo search  :413-456
// and the problem glue of forward/_problem.py:24-59 folded into the kernels: fixed DOFs are masked
// instead of gathered/scattered between "free" and "full" vectors (forward/dof_map/_dof_map.py:30-49).
//
// One iteration (all on one stream, no host round trip, every kernel a no-op once scal[DONE] != 0):
//   REDUCE     11 masked sums over (g, g_prev, diag, p_prev)                      1 pass over 4 vectors
//   FINALIZE   1 thread: |g|, termination tests, preconditioner mean, beta, descent guard
//   DIRECTION  p = -P g + beta p_prev, g.p; zeroes the trial buffers g', diag'    1 pass
//   PASS_B     pHp = hess_quad(x, p)                    element kernel (OP_HESS_QUAD)
//   ALPHA      1 thread: alpha_0 = -(g.p)/pHp sanitised * overstep, capped by max_step
//   TRIAL j    f', g', diag' at x + alpha_j p           element kernel (FUN|GRAD|DIAG); the trial
//              point is formed inside the gather, never written to memory; skipped once accepted
//   LS j       Armijo test on trial j; on failure alpha_{j+1} = alpha_j / 2 and re-zero g', diag'
//   COMMIT     accepted: x += alpha p, f = f'; otherwise g' = g, diag' = diag.  k += 1.
// The host flips the (g, g'), (diag, diag'), (p, p_prev) roles after every iteration.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>

#include "common.h"
#include "fem_kernels.cuh"
#include "vec_ops.cuh"

struct apl_xchg;

namespace apl {

int fem_eval_pncg(apl_fem* f, int ops, const void* x, const void* p, const void* axpy_p, double* scal,
                  int alpha_idx, int skip_a, int skip_b, double* fun_d, double* quad_d, void* grad, void* diag,
                  int scatter, cudaStream_t stream, int dyn_j = 0, void* third = nullptr);
int ext_force_pncg(int dtype, int ops, int64_t k, const void* force, const int32_t* indices, const void* x,
                   const void* axpy_p, double* scal, int alpha_idx, int skip_a, int skip_b, double* fun_d,
                   void* grad, cudaStream_t stream, int dyn_j = 0);

int xchg_push_ex(apl_xchg* x, int dtype, int nf, const void* f0, const void* f1, const void* f2, int ld, const void* scal,
                 int n_scal, int scal_f64, const double* skip_scal, int skip_a, int skip_b, int dyn_j, cudaStream_t s);
int xchg_pull_ex(apl_xchg* x, int dtype, int nf, void* f0, void* f1, void* f2, int ld, void* scal, int n_scal,
                 int scal_f64, const double* skip_scal, int skip_a, int skip_b, int dyn_j, cudaStream_t s);

constexpr int kNSums = 11;

// Block-Jacobi (opt-in, apl_pncg_set_block_jacobi): the preconditioner is the inverse of the 3x3 vertex blocks of the
// Hessian, diag = (xx, yy, zz) and offd = (xy, xz, yz), restricted to the vertex's FREE components (rows / columns of
// fixed components are replaced by the identity).  inv = [xx, yy, zz, xy, xz, yz] of the inverse; false when the
// restricted block is not positive definite (the caller then applies the reference's scalar rule to that vertex).
__device__ __forceinline__ bool block_inverse(const double d[3], const double o[3], const bool fr[3], double inv[6]) {
    const double a00 = fr[0] ? d[0] : 1.0, a11 = fr[1] ? d[1] : 1.0, a22 = fr[2] ? d[2] : 1.0;
    const double a01 = (fr[0] && fr[1]) ? o[0] : 0.0, a02 = (fr[0] && fr[2]) ? o[1] : 0.0, a12 = (fr[1] && fr[2]) ? o[2] : 0.0;
    const double c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
    const double m2 = a00 * a11 - a01 * a01;
    const double det = a00 * c00 + a01 * c01 + a02 * c02;
    if (!(a00 > 0.0 && m2 > 0.0 && det > 0.0) || !isfinite(det)) return false;
    const double r = 1.0 / det;
    inv[0] = c00 * r; inv[1] = (a00 * a22 - a02 * a02) * r; inv[2] = m2 * r;
    inv[3] = c01 * r; inv[4] = c02 * r; inv[5] = (a01 * a02 - a00 * a12) * r;
    return true;
}
__device__ __forceinline__ void sym_apply(const double inv[6], const double x[3], double y[3]) {
    y[0] = inv[0] * x[0] + inv[3] * x[1] + inv[4] * x[2];
    y[1] = inv[3] * x[0] + inv[1] * x[1] + inv[5] * x[2];
    y[2] = inv[4] * x[0] + inv[5] * x[1] + inv[2] * x[2];
}

template <typename T>
__global__ void __launch_bounds__(kVecThreads) pncg_reduce_kernel(long long rows, const T* __restrict__ g,
                                                                 const T* __restrict__ gprev,
                                                                 const T* __restrict__ diag,
                                                                 const T* __restrict__ offd,
                                                                 const T* __restrict__ pprev,
                                                                 const uchar4* __restrict__ mask, double* scal,
                                                                 double* partials, unsigned int* counter) {
    if (__ldcg(scal + APL_S_DONE) != 0.0) return;
    double s[kNSums];
#pragma unroll
    for (int i = 0; i < kNSums; ++i) s[i] = 0.0;
    for (long long r = blockIdx.x * (long long)kVecThreads + threadIdx.x; r < rows;
         r += (long long)gridDim.x * kVecThreads) {
        const uchar4 m4 = mask[r];
        const unsigned char m[4] = {m4.x, m4.y, m4.z, m4.w};
        if (((m4.x | m4.y | m4.z | m4.w) & APL_M_COUNT) == 0) continue;
        T gv[4], gp[4], dv[4], pp[4];
        ld4(g, r, gv); ld4(gprev, r, gp); ld4(diag, r, dv); ld4(pprev, r, pp);
        bool blocked = false;
        if (offd) {
            T ov[4];
            ld4(offd, r, ov);
            const bool fr[3] = {(m[0] & APL_M_FREE) != 0, (m[1] & APL_M_FREE) != 0, (m[2] & APL_M_FREE) != 0};
            const double d3[3] = {(double)dv[0], (double)dv[1], (double)dv[2]};
            const double o3[3] = {(double)ov[0], (double)ov[1], (double)ov[2]};
            double inv[6];
            if (block_inverse(d3, o3, fr, inv)) {
                blocked = true;
                double gi[3], yi[3], Pg[3], Py[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    gi[c] = fr[c] ? (double)gv[c] : 0.0;
                    yi[c] = fr[c] ? (double)gv[c] - (double)gp[c] : 0.0;
                }
                sym_apply(inv, gi, Pg);
                sym_apply(inv, yi, Py);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (!fr[c]) continue;
                    s[0] += 1.0; s[1] += fabs(d3[c]);
                    s[2] += gi[c] * Py[c]; s[4] += yi[c] * Py[c]; s[6] += gi[c] * Pg[c];
                    s[8] += gi[c] * (double)pp[c];
                    s[9] += yi[c] * (double)pp[c];
                    s[10] += gi[c] * gi[c];
                }
            }
        }
        if (blocked) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if ((m[c] & (APL_M_FREE | APL_M_COUNT)) != (APL_M_FREE | APL_M_COUNT)) continue;
            const double gi = (double)gv[c], yi = (double)gv[c] - (double)gp[c], pi = (double)pp[c];
            const double d = fabs((double)dv[c]);
            if (d > 0.0) {
                const double w = 1.0 / d;
                s[0] += 1.0; s[1] += d;
                s[2] += gi * yi * w; s[4] += yi * yi * w; s[6] += gi * gi * w;
            } else {
                s[3] += gi * yi; s[5] += yi * yi; s[7] += gi * gi;
            }
            s[8] += gi * pi;
            s[9] += yi * pi;
            s[10] += gi * gi;
        }
    }
    grid_reduce<kNSums>(s, partials, counter, scal + APL_S_SUMS);
}

struct PncgParams {
    double max_steps;
    double rtol_g;      // |g| <= rtol_g * |g_first|  -> DONE = 1
    double atol_g;      // |g| <= atol_g              -> DONE = 1
    double max_fails;   // consecutive failed line searches -> DONE = 3
    double overstep;
    double max_step;
    double c1;
    int max_halvings;
};

__global__ void pncg_finalize_kernel(double* scal, PncgParams prm) {
    if (scal[APL_S_DONE] != 0.0) return;
    const double* S = scal + APL_S_SUMS;
    const double k = scal[APL_S_K];
    const double cnt = S[0], dsum = S[1];
    const double mean = cnt > 0.0 ? dsum / cnt : 1.0;
    const double inv = 1.0 / mean;
    const double gPy = S[2] + S[3] * inv, yPy = S[4] + S[5] * inv, gPg = S[6] + S[7] * inv;
    const double gpp = S[8], yp = S[9], gnorm2 = S[10];
    if (k == 0.0) scal[APL_S_GNORM2_FIRST] = gnorm2;
    scal[APL_S_GNORM2] = gnorm2;
    scal[APL_S_DIAG_MEAN] = mean;
    scal[APL_S_GPG] = gPg;
    const double g0 = scal[APL_S_GNORM2_FIRST];
    if (!(gnorm2 == gnorm2) || isinf(gnorm2)) { scal[APL_S_DONE] = 4.0; return; }
    if (gnorm2 <= prm.atol_g * prm.atol_g || gnorm2 <= prm.rtol_g * prm.rtol_g * g0) { scal[APL_S_DONE] = 1.0; return; }
    if (k >= prm.max_steps) { scal[APL_S_DONE] = 2.0; return; }
    if (scal[APL_S_FAILS] >= prm.max_fails) { scal[APL_S_DONE] = 3.0; return; }
    // Dai-Kou beta (bench :663-679) with the reset rules (:288-289)
    double beta = 0.0;
    if (k != 0.0) {
        if (fabs(yp) > 1.0e-12) beta = gPy / yp - (yPy / yp) * (gpp / yp);
        else beta = INFINITY;
        if (!isfinite(beta) || fabs(beta) > 10.0) beta = 0.0;
    }
    // descent guard (:290-303): g.p = -g.Pg + beta g.p_prev must be finite and negative
    const double gp_pred = -gPg + beta * gpp;
    if (!(isfinite(gp_pred) && gp_pred < 0.0)) beta = 0.0;
    scal[APL_S_BETA] = beta;
    scal[APL_S_PHP] = 0.0;
    scal[APL_S_F_NEW] = 0.0;
    scal[APL_S_J] = 0.0;
    for (int j = 0; j < 16; ++j) {
        scal[APL_S_ALPHA_J + j] = 0.0;
        scal[APL_S_ACC_J + j] = 0.0;
        scal[APL_S_FT_J + j] = 0.0;
    }
}

template <typename T>
__global__ void __launch_bounds__(kVecThreads) pncg_direction_kernel(long long rows, const T* __restrict__ g,
                                                                    const T* __restrict__ diag,
                                                                    const T* __restrict__ offd,
                                                                    const T* __restrict__ pprev, T* __restrict__ p,
                                                                    T* __restrict__ gz, T* __restrict__ dz,
                                                                    T* __restrict__ oz,
                                                                    const uchar4* __restrict__ mask, double* scal,
                                                                    double* partials, unsigned int* counter) {
    if (__ldcg(scal + APL_S_DONE) != 0.0) return;
    const double beta = __ldcg(scal + APL_S_BETA);
    const double mean = __ldcg(scal + APL_S_DIAG_MEAN);
    double s[1] = {0.0};
    const T zero[4] = {(T)0, (T)0, (T)0, (T)0};
    for (long long r = blockIdx.x * (long long)kVecThreads + threadIdx.x; r < rows;
         r += (long long)gridDim.x * kVecThreads) {
        const uchar4 m4 = mask[r];
        const unsigned char m[4] = {m4.x, m4.y, m4.z, m4.w};
        T gv[4], dv[4], pp[4], pv[4];
        ld4(g, r, gv); ld4(diag, r, dv); ld4(pprev, r, pp);
        bool blocked = false;
        if (offd) {
            T ov[4];
            ld4(offd, r, ov);
            const bool fr[3] = {(m[0] & APL_M_FREE) != 0, (m[1] & APL_M_FREE) != 0, (m[2] & APL_M_FREE) != 0};
            const double d3[3] = {(double)dv[0], (double)dv[1], (double)dv[2]};
            const double o3[3] = {(double)ov[0], (double)ov[1], (double)ov[2]};
            double inv[6];
            if (block_inverse(d3, o3, fr, inv)) {
                blocked = true;
                const double gi[3] = {fr[0] ? (double)gv[0] : 0.0, fr[1] ? (double)gv[1] : 0.0, fr[2] ? (double)gv[2] : 0.0};
                double Pg[3];
                sym_apply(inv, gi, Pg);
                const bool counted = ((m4.x | m4.y | m4.z) & APL_M_COUNT) != 0;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const double pi = fr[c] ? -Pg[c] + beta * (double)pp[c] : 0.0;
                    if (counted) s[0] += gi[c] * pi;
                    pv[c] = (T)pi;
                }
                pv[3] = (T)0;
            }
        }
        if (!blocked) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                double pi = 0.0;
                if (m[c] & APL_M_FREE) {
                    double d = fabs((double)dv[c]);
                    if (!(d > 0.0)) d = mean;
                    pi = -(double)gv[c] / d + beta * (double)pp[c];
                    if (m[c] & APL_M_COUNT) s[0] += (double)gv[c] * pi;
                }
                pv[c] = (T)pi;
            }
        }
        st4(p, r, pv);
        st4(gz, r, zero);
        st4(dz, r, zero);
        if (oz) st4(oz, r, zero);
    }
    grid_reduce<1>(s, partials, counter, scal + APL_S_GP);
}

__global__ void pncg_alpha_kernel(double* scal, PncgParams prm) {
    if (scal[APL_S_DONE] != 0.0) return;
    // bench :606-610
    double alpha = -scal[APL_S_GP] / scal[APL_S_PHP];
    if (alpha != alpha || alpha == -INFINITY) alpha = 0.0;
    else if (alpha == INFINITY) alpha = 1.0;
    if (!(alpha > 0.0 && isfinite(alpha))) alpha = 1.0;
    alpha *= prm.overstep;
    if (alpha > prm.max_step) alpha = prm.max_step;  // forward/_problem.py:29-34
    scal[APL_S_ALPHA_J] = alpha;
}

// Armijo test of trial j (bench :413-456).  State for trial j+1 goes to slot j+1, so every thread
// of the grid reads slot j while thread 0 writes slot j+1: no intra-kernel race.
__global__ void pncg_bump_j_kernel(double* scal) {
    if (scal[APL_S_DONE] != 0.0) return;
    scal[APL_S_J] += 1.0;
}

// j < 0: the trial index is the device-side counter scal[APL_S_J] - 1 (the trial that just ran) and the
// outcome drives a CUDA-graph WHILE node: condition 1 = run another trial.
template <typename T>
__global__ void __launch_bounds__(kVecThreads) pncg_ls_kernel(long long rows, int j, int last, T* __restrict__ gz,
                                                             T* __restrict__ dz, T* __restrict__ oz, double* scal, PncgParams prm,
                                                             cudaGraphConditionalHandle cond, int use_cond) {
    if (__ldcg(scal + APL_S_DONE) != 0.0) return;
    if (j < 0) {
        j = (int)__ldcg(scal + APL_S_J) - 1;
        last = (j >= prm.max_halvings) ? 1 : 0;
    }
    const double acc = __ldcg(scal + APL_S_ACC_J + j);
    const double alpha = __ldcg(scal + APL_S_ALPHA_J + j);
    const double ft = __ldcg(scal + APL_S_FT_J + j);
    const double f = __ldcg(scal + APL_S_F);
    const double gp = __ldcg(scal + APL_S_GP);
    // slot value: 1 = accepted, 0 = this trial is live, -1 = gave up (no further trials)
    const bool gave_up = acc < 0.0;
    bool accepted = acc > 0.0;
    bool newly = false;
    if (!accepted && !gave_up) {
        newly = isfinite(ft) && (ft <= f + prm.c1 * alpha * gp);
        accepted = newly;
    }
    // alpha > 0 is part of the loop condition (:448): a zero step cannot be halved further
    const bool retry = !accepted && !gave_up && !last && alpha > 0.0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        scal[APL_S_ACC_J + j + 1] = accepted ? 1.0 : (retry ? 0.0 : -1.0);
        scal[APL_S_ALPHA_J + j + 1] = retry ? alpha * 0.5 : alpha;
        if (newly) scal[APL_S_F_NEW] = ft;
        if (use_cond) cudaGraphSetConditional(cond, retry ? 1u : 0u);
    }
    if (!retry) return;
    const T zero[4] = {(T)0, (T)0, (T)0, (T)0};
    for (long long r = blockIdx.x * (long long)kVecThreads + threadIdx.x; r < rows;
         r += (long long)gridDim.x * kVecThreads) {
        st4(gz, r, zero);
        st4(dz, r, zero);
        if (oz) st4(oz, r, zero);
    }
}

template <typename T>
__global__ void __launch_bounds__(kVecThreads) pncg_commit_kernel(long long rows, int jfinal, T* __restrict__ x,
                                                                 const T* __restrict__ p, const T* __restrict__ g,
                                                                 T* __restrict__ gz, const T* __restrict__ diag,
                                                                 T* __restrict__ dz, const T* __restrict__ offd,
                                                                 T* __restrict__ oz, double* scal) {
    if (__ldcg(scal + APL_S_DONE) != 0.0) return;
    if (jfinal < 0) jfinal = (int)__ldcg(scal + APL_S_J);  // conditional-graph line search: trials run so far
    const bool accepted = __ldcg(scal + APL_S_ACC_J + jfinal) > 0.0;
    const double alpha = __ldcg(scal + APL_S_ALPHA_J + jfinal);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int halvings = 0;
        for (int j = 1; j < jfinal; ++j) halvings += (scal[APL_S_ACC_J + j] == 0.0) ? 1 : 0;
        scal[APL_S_LS_STEPS] = (double)halvings;
        scal[APL_S_ACCEPTED] = accepted ? 1.0 : 0.0;
        scal[APL_S_ALPHA] = alpha;
        if (accepted) {
            scal[APL_S_F_PREV] = scal[APL_S_F];
            scal[APL_S_F] = scal[APL_S_F_NEW];
            scal[APL_S_N_ACCEPTED] += 1.0;
            scal[APL_S_FAILS] = 0.0;
        } else {
            scal[APL_S_FAILS] += 1.0;
        }
        scal[APL_S_K] += 1.0;
    }
    const T a = (T)alpha;
    for (long long r = blockIdx.x * (long long)kVecThreads + threadIdx.x; r < rows;
         r += (long long)gridDim.x * kVecThreads) {
        if (accepted) {
            T xv[4], pv[4];
            ld4(x, r, xv); ld4(p, r, pv);
#pragma unroll
            for (int c = 0; c < 4; ++c) xv[c] += a * pv[c];
            st4(x, r, xv);
        } else {
            T v[4];
            ld4(g, r, v); st4(gz, r, v);
            ld4(diag, r, v); st4(dz, r, v);
            if (oz) { ld4(offd, r, v); st4(oz, r, v); }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kVecThreads) pncg_zero2_kernel(long long rows, T* __restrict__ a, T* __restrict__ b) {
    const T zero[4] = {(T)0, (T)0, (T)0, (T)0};
    for (long long r = blockIdx.x * (long long)kVecThreads + threadIdx.x; r < rows;
         r += (long long)gridDim.x * kVecThreads) {
        st4(a, r, zero);
        st4(b, r, zero);
    }
}

}  // namespace apl

using namespace apl;

struct apl_pncg {
    int dtype = 0, device = 0;
    int64_t rows = 0;  // n_points; every vector is (rows, 4)
    void* x = nullptr;
    void* p[2] = {nullptr, nullptr};
    void* g[2] = {nullptr, nullptr};
    void* d[2] = {nullptr, nullptr};
    void* o[2] = {nullptr, nullptr};   // vertex-block off-diagonals (block Jacobi, opt-in); nullptr = scalar Jacobi
    int psd = 0;                       // APL_OP_PSD in the Hessian passes
    const uint8_t* mask = nullptr;
    double* scal = nullptr;
    double* partials = nullptr;
    unsigned int* counter = nullptr;
    int cur = 0;  // index of the current g / diag / p buffers
    int grid = 1;
    int scatter = APL_SCATTER_TILE;
    PncgParams prm;
    std::vector<apl_fem*> fems;
    struct Ext { const void* force; const int32_t* idx; int64_t k; };
    std::vector<Ext> exts;
    cudaGraphExec_t graph[2] = {nullptr, nullptr};
    cudaGraphConditionalHandle cond[2] = {0, 0};
    int graph_mode = 0;  // 0 none, 1 static trials, 2 conditional WHILE
    cudaStream_t capture_stream = nullptr;  // graphs are captured here (the legacy stream cannot capture)
    bool use_graph = false;
    // sharded meshes: partial reductions and nodal sums of this rank are completed through peer memory right after
    // the phase that produced them (apl_pncg_set_exchange); nullptr = single GPU
    apl_xchg* xchg = nullptr;
};

namespace {

// Completes this rank's partial results of a phase over all ranks: halo sum of up to two nodal fields (ld = 4) and
// the global sum of n_scal workspace scalars starting at scal[idx (+ trial counter)], guarded by the same skip flags
// as the launches that produced them.
int exchange(apl_pncg* w, void* f0, void* f1, int idx, int n_scal, int skip_a, int skip_b, int dyn_j, cudaStream_t s,
             void* f2 = nullptr) {
    if (!w->xchg) return APL_OK;
    const int nf = f0 ? (f1 ? (f2 ? 3 : 2) : 1) : 0;
    int rc = xchg_push_ex(w->xchg, w->dtype, nf, f0, f1, f2, 4, w->scal + idx, n_scal, 1, w->scal, skip_a, skip_b,
                          dyn_j, s);
    if (rc != APL_OK) return rc;
    return xchg_pull_ex(w->xchg, w->dtype, nf, f0, f1, f2, 4, w->scal + idx, n_scal, 1, w->scal, skip_a, skip_b,
                        dyn_j, s);
}

template <typename T>
int phase_typed(apl_pncg* w, int phase, int j, cudaStream_t s) {
    const int c = w->cur, o = 1 - c;
    T* x = (T*)w->x;
    T* g = (T*)w->g[c];
    T* gz = (T*)w->g[o];
    T* dg = (T*)w->d[c];
    T* dz = (T*)w->d[o];
    T* p = (T*)w->p[c];
    T* pprev = (T*)w->p[o];
    T* og = (T*)w->o[c];      // block off-diagonals at the current iterate / of the trial (nullptr: scalar Jacobi)
    T* oz = (T*)w->o[o];
    const int ops_a = APL_OP_FUN | APL_OP_GRAD | APL_OP_HESS_DIAG | (og ? APL_OP_HESS_OFFD : 0) | (w->psd ? APL_OP_PSD : 0);
    const int ops_b = APL_OP_HESS_QUAD | (w->psd ? APL_OP_PSD : 0);
    const uchar4* mask = (const uchar4*)w->mask;
    const int J = w->prm.max_halvings;
    switch (phase) {
        case APL_PHASE_REDUCE:
            pncg_reduce_kernel<T><<<w->grid, kVecThreads, 0, s>>>(w->rows, g, gz, dg, og, pprev, mask, w->scal,
                                                                  w->partials, w->counter);
            if (int rc = exchange(w, nullptr, nullptr, APL_S_SUMS, kNSums, APL_S_DONE, -1, 0, s)) return rc;
            break;
        case APL_PHASE_FINALIZE:
            pncg_finalize_kernel<<<1, 1, 0, s>>>(w->scal, w->prm);
            break;
        case APL_PHASE_DIRECTION:
            pncg_direction_kernel<T><<<w->grid, kVecThreads, 0, s>>>(w->rows, g, dg, og, pprev, p, gz, dz, oz, mask,
                                                                     w->scal, w->partials, w->counter);
            break;
        case APL_PHASE_PASS_B:
            for (apl_fem* f : w->fems) {
                int rc = fem_eval_pncg(f, ops_b, x, p, nullptr, w->scal, 0, APL_S_DONE, -1, nullptr,
                                       w->scal + APL_S_PHP, nullptr, nullptr, w->scatter, s);
                if (rc != APL_OK) return rc;
            }
            // g.p (DIRECTION) and p.Hp (this pass) are adjacent partial sums: one exchange completes both
            if (int rc = exchange(w, nullptr, nullptr, APL_S_GP, 2, APL_S_DONE, -1, 0, s)) return rc;
            break;
        case APL_PHASE_ALPHA:
            pncg_alpha_kernel<<<1, 1, 0, s>>>(w->scal, w->prm);
            break;
        case APL_PHASE_TRIAL: {
            // j == -1: trial index taken from the device-side counter (conditional-graph line search)
            if (j < -1 || j > J) { set_error("apl_pncg_phase: trial index out of range"); return APL_ERR_INVALID; }
            const int dyn = j < 0 ? 1 : 0, jj = dyn ? 0 : j;
            for (apl_fem* f : w->fems) {
                int rc = fem_eval_pncg(f, ops_a, x, nullptr, p, w->scal,
                                       APL_S_ALPHA_J + jj, APL_S_DONE, APL_S_ACC_J + jj, w->scal + APL_S_FT_J + jj,
                                       nullptr, gz, dz, w->scatter, s, dyn, oz);
                if (rc != APL_OK) return rc;
            }
            for (const auto& e : w->exts) {
                int rc = ext_force_pncg(w->dtype, APL_OP_FUN | APL_OP_GRAD, e.k, e.force, e.idx, x, p, w->scal,
                                        APL_S_ALPHA_J + jj, APL_S_DONE, APL_S_ACC_J + jj, w->scal + APL_S_FT_J + jj, gz, s,
                                        dyn);
                if (rc != APL_OK) return rc;
            }
            if (int rc = exchange(w, gz, dz, APL_S_FT_J + jj, 1, APL_S_DONE, APL_S_ACC_J + jj, dyn, s, oz)) return rc;
            if (dyn) pncg_bump_j_kernel<<<1, 1, 0, s>>>(w->scal);
            break;
        }
        case APL_PHASE_LS: {
            if (j < -1 || j > J) { set_error("apl_pncg_phase: trial index out of range"); return APL_ERR_INVALID; }
            if (j < 0) {
                pncg_ls_kernel<T><<<w->grid, kVecThreads, 0, s>>>(w->rows, -1, 0, gz, dz, oz, w->scal, w->prm,
                                                                   w->cond[w->cur], 1);
            } else {
                const int last = (j == J) ? 1 : 0;
                pncg_ls_kernel<T><<<last ? 1 : w->grid, last ? 32 : kVecThreads, 0, s>>>(w->rows, j, last, gz, dz, oz,
                                                                                       w->scal, w->prm, 0, 0);
            }
            break;
        }
        case APL_PHASE_COMMIT:
            pncg_commit_kernel<T><<<w->grid, kVecThreads, 0, s>>>(w->rows, j < 0 ? -1 : J + 1, x, p, g, gz, dg, dz, og, oz,
                                                                  w->scal);
            break;
        case APL_PHASE_INIT: {
            // f, g, diag at x into the CURRENT buffers; resets every scalar
            APL_CUDA_CHECK(cudaMemsetAsync(w->scal, 0, sizeof(double) * APL_PNCG_NSCAL, s));
            pncg_zero2_kernel<T><<<w->grid, kVecThreads, 0, s>>>(w->rows, g, dg);
            pncg_zero2_kernel<T><<<w->grid, kVecThreads, 0, s>>>(w->rows, gz, pprev);
            if (og) pncg_zero2_kernel<T><<<w->grid, kVecThreads, 0, s>>>(w->rows, og, oz);
            for (apl_fem* f : w->fems) {
                int rc = fem_eval_pncg(f, ops_a, x, nullptr, nullptr, w->scal, 0,
                                       -1, -1, w->scal + APL_S_F, nullptr, g, dg, w->scatter, s, 0, og);
                if (rc != APL_OK) return rc;
            }
            for (const auto& e : w->exts) {
                int rc = ext_force_pncg(w->dtype, APL_OP_FUN | APL_OP_GRAD, e.k, e.force, e.idx, x, nullptr, w->scal, 0,
                                        -1, -1, w->scal + APL_S_F, g, s);
                if (rc != APL_OK) return rc;
            }
            if (int rc = exchange(w, g, dg, APL_S_F, 1, -1, -1, 0, s, og)) return rc;
            break;
        }
        default:
            set_error("apl_pncg_phase: unknown phase");
            return APL_ERR_INVALID;
    }
    APL_CUDA_CHECK(cudaGetLastError());
    return APL_OK;
}

int run_phase(apl_pncg* w, int phase, int j, cudaStream_t s) {
    return w->dtype == APL_F32 ? phase_typed<float>(w, phase, j, s) : phase_typed<double>(w, phase, j, s);
}

// One iteration as a CUDA graph whose backtracking is a conditional WHILE node:
//   head: REDUCE .. ALPHA, TRIAL(0), LS            (LS sets the loop condition: 1 = Armijo failed, retry)
//   WHILE (condition) { TRIAL(j), LS }              (device-side loop, at most max_halvings times)
//   tail: COMMIT
int build_conditional_graph(apl_pncg* w, cudaStream_t cs) {
    const int c = w->cur;
    cudaGraph_t g = nullptr;
    APL_CUDA_CHECK(cudaGraphCreate(&g, 0));
    auto fail = [&](int rc) { cudaGraphDestroy(g); return rc; };
    cudaError_t e = cudaGraphConditionalHandleCreate(&w->cond[c], g, 0, cudaGraphCondAssignDefault);
    if (e != cudaSuccess) { set_error(std::string("cudaGraphConditionalHandleCreate: ") + cudaGetErrorString(e)); return fail(APL_ERR_CUDA); }
    static const int head[] = {APL_PHASE_REDUCE, APL_PHASE_FINALIZE, APL_PHASE_DIRECTION, APL_PHASE_PASS_B,
                               APL_PHASE_ALPHA, APL_PHASE_TRIAL, APL_PHASE_LS};
    e = cudaStreamBeginCaptureToGraph(cs, g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { set_error(std::string("cudaStreamBeginCaptureToGraph: ") + cudaGetErrorString(e)); return fail(APL_ERR_CUDA); }
    int rc = APL_OK;
    for (int ph : head) {
        rc = run_phase(w, ph, -1, cs);
        if (rc != APL_OK) break;
    }
    std::vector<cudaGraphNode_t> deps;
    if (rc == APL_OK) {
        cudaStreamCaptureStatus st;
        const cudaGraphNode_t* d = nullptr;
        size_t nd = 0;
        e = cudaStreamGetCaptureInfo_v2(cs, &st, nullptr, nullptr, &d, &nd);
        if (e == cudaSuccess) deps.assign(d, d + nd);
    }
    cudaGraph_t tmp = nullptr;
    cudaError_t e2 = cudaStreamEndCapture(cs, &tmp);
    if (rc != APL_OK) return fail(rc);
    if (e != cudaSuccess || e2 != cudaSuccess) { set_error("conditional graph: capturing the head failed"); return fail(APL_ERR_CUDA); }
    cudaGraphNodeParams prm = {cudaGraphNodeTypeConditional};
    prm.conditional.handle = w->cond[c];
    prm.conditional.type = cudaGraphCondTypeWhile;
    prm.conditional.size = 1;
    cudaGraphNode_t cond_node = nullptr;
    e = cudaGraphAddNode(&cond_node, g, deps.data(), deps.size(), &prm);
    if (e != cudaSuccess) { set_error(std::string("cudaGraphAddNode(conditional): ") + cudaGetErrorString(e)); return fail(APL_ERR_CUDA); }
    cudaGraph_t body = prm.conditional.phGraph_out[0];
    e = cudaStreamBeginCaptureToGraph(cs, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { set_error(std::string("capture of the WHILE body: ") + cudaGetErrorString(e)); return fail(APL_ERR_CUDA); }
    rc = run_phase(w, APL_PHASE_TRIAL, -1, cs);
    if (rc == APL_OK) rc = run_phase(w, APL_PHASE_LS, -1, cs);
    e = cudaStreamEndCapture(cs, &tmp);
    if (rc != APL_OK) return fail(rc);
    if (e != cudaSuccess) { set_error(std::string("end capture of the WHILE body: ") + cudaGetErrorString(e)); return fail(APL_ERR_CUDA); }
    e = cudaStreamBeginCaptureToGraph(cs, g, &cond_node, nullptr, 1, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { set_error(std::string("capture of the tail: ") + cudaGetErrorString(e)); return fail(APL_ERR_CUDA); }
    rc = run_phase(w, APL_PHASE_COMMIT, -1, cs);
    e = cudaStreamEndCapture(cs, &tmp);
    if (rc != APL_OK) return fail(rc);
    if (e != cudaSuccess) { set_error(std::string("end capture of the tail: ") + cudaGetErrorString(e)); return fail(APL_ERR_CUDA); }
    e = cudaGraphInstantiate(&w->graph[c], g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { set_error(std::string("cudaGraphInstantiate(conditional): ") + cudaGetErrorString(e)); return APL_ERR_CUDA; }
    return APL_OK;
}

int one_iteration(apl_pncg* w, cudaStream_t s) {
    static const int head[] = {APL_PHASE_REDUCE, APL_PHASE_FINALIZE, APL_PHASE_DIRECTION, APL_PHASE_PASS_B,
                               APL_PHASE_ALPHA};
    for (int ph : head) {
        int rc = run_phase(w, ph, 0, s);
        if (rc != APL_OK) return rc;
    }
    for (int j = 0; j <= w->prm.max_halvings; ++j) {
        int rc = run_phase(w, APL_PHASE_TRIAL, j, s);
        if (rc != APL_OK) return rc;
        rc = run_phase(w, APL_PHASE_LS, j, s);
        if (rc != APL_OK) return rc;
    }
    return run_phase(w, APL_PHASE_COMMIT, 0, s);
}

void drop_graphs(apl_pncg* w) {
    for (int i = 0; i < 2; ++i)
        if (w->graph[i]) {
            cudaGraphExecDestroy(w->graph[i]);
            w->graph[i] = nullptr;
        }
}

}  // namespace

extern "C" {

int apl_pncg_create(int dtype, int64_t n_points, int device, void* x, void* p0, void* p1, void* g0, void* g1,
                    void* d0, void* d1, const uint8_t* mask, double* scal, apl_pncg_t** out) {
    if (!out) { set_error("apl_pncg_create: out is NULL"); return APL_ERR_INVALID; }
    *out = nullptr;
    if ((dtype != APL_F32 && dtype != APL_F64) || n_points <= 0 || !x || !p0 || !p1 || !g0 || !g1 || !d0 || !d1 ||
        !mask || !scal) {
        set_error("apl_pncg_create: bad arguments");
        return APL_ERR_INVALID;
    }
    APL_CUDA_CHECK(cudaSetDevice(device));
    apl_pncg* w = new apl_pncg();
    w->dtype = dtype; w->device = device; w->rows = n_points;
    w->x = x; w->p[0] = p0; w->p[1] = p1; w->g[0] = g0; w->g[1] = g1; w->d[0] = d0; w->d[1] = d1;
    w->mask = mask; w->scal = scal;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete w; set_error(cudaGetErrorString(e)); return APL_ERR_CUDA; }
    long long need = (n_points + kVecThreads - 1) / kVecThreads;
    long long cap = (long long)prop.multiProcessorCount * 4;
    w->grid = (int)(need < 1 ? 1 : (need > cap ? cap : need));
    e = cudaMalloc((void**)&w->partials, sizeof(double) * 16 * (size_t)cap);
    if (e == cudaSuccess) e = cudaMalloc((void**)&w->counter, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(w->counter, 0, sizeof(unsigned int));
    if (e != cudaSuccess) { cudaFree(w->partials); cudaFree(w->counter); delete w; set_error(cudaGetErrorString(e)); return APL_ERR_CUDA; }
    w->prm.max_steps = 1500; w->prm.rtol_g = 0; w->prm.atol_g = 0; w->prm.max_fails = 1e300;
    w->prm.overstep = 1.0; w->prm.max_step = 1.0; w->prm.c1 = 1e-4; w->prm.max_halvings = 8;
    *out = w;
    return APL_OK;
}

void apl_pncg_destroy(apl_pncg_t* w) {
    if (!w) return;
    drop_graphs(w);
    if (w->capture_stream) cudaStreamDestroy(w->capture_stream);
    cudaFree(w->partials);
    cudaFree(w->counter);
    delete w;
}

int apl_pncg_add_fem(apl_pncg_t* w, apl_fem_t* fem) {
    if (!w || !fem) { set_error("apl_pncg_add_fem: NULL argument"); return APL_ERR_INVALID; }
    if (fem->dtype != w->dtype || fem->device != w->device || fem->host.n_points > w->rows) {
        set_error("apl_pncg_add_fem: potential does not match the workspace (dtype, device or n_points)");
        return APL_ERR_INVALID;
    }
    w->fems.push_back(fem);
    drop_graphs(w);
    return APL_OK;
}

int apl_pncg_add_ext_force(apl_pncg_t* w, int64_t k, const void* force, const int32_t* indices) {
    if (!w || k < 0 || (k > 0 && (!force || !indices))) { set_error("apl_pncg_add_ext_force: bad arguments"); return APL_ERR_INVALID; }
    if (k > 0) w->exts.push_back({force, indices, k});
    drop_graphs(w);
    return APL_OK;
}

int apl_pncg_set_params(apl_pncg_t* w, double max_steps, double rtol_g, double atol_g, double max_fails,
                        double overstep, double max_step, double c1, int max_halvings, int scatter, int use_graph) {
    if (!w) { set_error("apl_pncg_set_params: NULL workspace"); return APL_ERR_INVALID; }
    if (max_halvings < 0 || max_halvings > 14) { set_error("apl_pncg_set_params: max_halvings must be in [0, 14]"); return APL_ERR_INVALID; }
    w->prm.max_steps = max_steps; w->prm.rtol_g = rtol_g; w->prm.atol_g = atol_g; w->prm.max_fails = max_fails;
    w->prm.overstep = overstep; w->prm.max_step = max_step; w->prm.c1 = c1; w->prm.max_halvings = max_halvings;
    w->scatter = scatter;
    w->use_graph = use_graph != 0;
    w->graph_mode = use_graph;
    drop_graphs(w);
    return APL_OK;
}

int apl_pncg_set_block_jacobi(apl_pncg_t* w, void* o0, void* o1, int psd) {
    if (!w) { set_error("apl_pncg_set_block_jacobi: NULL workspace"); return APL_ERR_INVALID; }
    if ((o0 == nullptr) != (o1 == nullptr)) { set_error("apl_pncg_set_block_jacobi: give both buffers or none"); return APL_ERR_INVALID; }
    if ((o0 || psd) && w->scatter != APL_SCATTER_TILE) {
        set_error("apl_pncg_set_block_jacobi: block Jacobi / PSD projection need the TILE assembly");
        return APL_ERR_INVALID;
    }
    w->o[0] = o0; w->o[1] = o1; w->psd = psd ? 1 : 0;
    drop_graphs(w);
    return APL_OK;
}

int apl_pncg_set_exchange(apl_pncg_t* w, apl_xchg_t* x) {
    if (!w) { set_error("apl_pncg_set_exchange: NULL workspace"); return APL_ERR_INVALID; }
    w->xchg = x;
    drop_graphs(w);
    return APL_OK;
}

int apl_pncg_current(const apl_pncg_t* w) { return w ? w->cur : APL_ERR_INVALID; }

int apl_pncg_flip(apl_pncg_t* w) {
    if (!w) { set_error("apl_pncg_flip: NULL workspace"); return APL_ERR_INVALID; }
    w->cur = 1 - w->cur;
    return APL_OK;
}

int apl_pncg_phase(apl_pncg_t* w, int phase, int j, void* stream) {
    if (!w) { set_error("apl_pncg_phase: NULL workspace"); return APL_ERR_INVALID; }
    if (phase == APL_PHASE_INIT) w->cur = 0;
    return run_phase(w, phase, j, (cudaStream_t)stream);
}

int apl_pncg_iterate(apl_pncg_t* w, int n_iters, void* stream) {
    if (!w || n_iters < 0) { set_error("apl_pncg_iterate: bad arguments"); return APL_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    for (int it = 0; it < n_iters; ++it) {
        if (w->use_graph) {
            if (!w->capture_stream)
                APL_CUDA_CHECK(cudaStreamCreateWithFlags(&w->capture_stream, cudaStreamNonBlocking));
            if (!w->graph[w->cur] && w->graph_mode == 2) {
                int rc = build_conditional_graph(w, w->capture_stream);
                if (rc != APL_OK) return rc;
            }
            if (!w->graph[w->cur]) {
                // capture one iteration (all max_halvings + 1 flag-guarded trials) for this buffer parity
                cudaGraph_t graph = nullptr;
                APL_CUDA_CHECK(cudaStreamBeginCapture(w->capture_stream, cudaStreamCaptureModeThreadLocal));
                int rc = one_iteration(w, w->capture_stream);
                cudaError_t e = cudaStreamEndCapture(w->capture_stream, &graph);
                if (rc != APL_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
                if (e != cudaSuccess) { set_error(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e)); return APL_ERR_CUDA; }
                e = cudaGraphInstantiate(&w->graph[w->cur], graph, 0);
                cudaGraphDestroy(graph);
                if (e != cudaSuccess) { set_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); return APL_ERR_CUDA; }
            }
            APL_CUDA_CHECK(cudaGraphLaunch(w->graph[w->cur], s));
        } else {
            int rc = one_iteration(w, s);
            if (rc != APL_OK) return rc;
        }
        w->cur = 1 - w->cur;
    }
    return APL_OK;
}

}  // extern "C"
