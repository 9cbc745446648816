// Fused Jacobi-preconditioned conjugate gradients on hess_prod: the adjoint / sensitivity solve H(u) x = b of the
// reference's inverse problems,
//   jax.scipy.sparse.linalg.cg(lambda p: model.hess_prod(u, p), -dLdu, tol=1e-5, atol=1e-15, maxiter=n // 10,
//                              M=lambda x: P * x)             exp/2025/09/24/inverse-grin/src/35-inverse-small-reg.py:223-260
// with P = 1 / hess_diag(u) (:223).  The reference runs it as XLA ops around an FFI callback per matvec; here one
// iteration is five stream-ordered launches with every scalar on the device (replayed as a CUDA graph):
//   MATVEC   Ap += H(u) p                         element kernel (OP_HESS_PROD), one pass over the tets
//   DOT      pAp = p . Ap
//   UPDATE   alpha = rz / pAp;  x += alpha p;  r -= alpha Ap;  rz' = r . M r;  rr = r . r
//   STEP     1 thread: convergence (|r| <= max(tol |b|, atol): jax semantics), beta = rz' / rz, k += 1
//   DIRECT   p = M r + beta p;  Ap = 0
// Fixed DOFs are masked (bit 0 of the per-DOF byte, as in pncg.cu) instead of gathered / scattered
// (forward/dof_map/_dof_map.py:30-49).  M = 1 / fix(|hess_diag|) with the PNCG fix-up for non-positive entries.
#include <cuda_runtime.h>

#include <cmath>

#include "common.h"
#include "fem_kernels.cuh"
#include "vec_ops.cuh"

namespace apl {

int fem_eval_pncg(apl_fem* f, int ops, const void* x, const void* p, const void* axpy_p, double* scal,
                  int alpha_idx, int skip_a, int skip_b, double* fun_d, double* quad_d, void* grad, void* diag,
                  int scatter, cudaStream_t stream, int dyn_j = 0, void* third = nullptr);

// scalars of the workspace (device doubles)
enum { C_RZ = 0, C_PAP = 1, C_RZ_NEW = 2, C_RR = 3, C_BB = 4, C_DONE = 5, C_K = 6, C_BETA = 7, C_TARGET2 = 8,
       C_DCNT = 9, C_DSUM = 10, C_DMEAN = 11, C_SUMS = 12 /* 3 scratch sums */ };

struct PcgParams { double tol, atol, max_iters; };

template <typename T>
__device__ __forceinline__ double fix_diag(T d, double mean) {
    const double a = fabs((double)d);
    return a > 0.0 ? a : mean;
}

// r = b - Ap on free DOFs (Ap = H x0 accumulated by the caller's matvec, or 0); sums: #positive |d|, their sum, b.b
template <typename T>
__global__ void __launch_bounds__(kVecThreads) pcg_init1_kernel(long long rows, const T* __restrict__ b, const T* __restrict__ Ap,
                                                               T* __restrict__ r, const T* __restrict__ diag,
                                                               const uchar4* __restrict__ mask, double* scal,
                                                               double* partials, unsigned int* counter) {
    double s[3] = {0.0, 0.0, 0.0};
    for (long long i = blockIdx.x * (long long)kVecThreads + threadIdx.x; i < rows; i += (long long)gridDim.x * kVecThreads) {
        const uchar4 m4 = mask[i];
        const unsigned char m[4] = {m4.x, m4.y, m4.z, m4.w};
        T bv[4], av[4], dv[4], rv[4];
        ld4(b, i, bv); ld4(Ap, i, av); ld4(diag, i, dv);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double ri = 0.0;
            if (c < 3 && (m[c] & APL_M_FREE)) {
                ri = (double)bv[c] - (double)av[c];
                const double d = fabs((double)dv[c]);
                if (d > 0.0) { s[0] += 1.0; s[1] += d; }
                s[2] += (double)bv[c] * (double)bv[c];
            }
            rv[c] = (T)ri;
        }
        st4(r, i, rv);
    }
    grid_reduce<3>(s, partials, counter, scal + C_SUMS);
}

__global__ void pcg_init_scalars_kernel(double* scal, PcgParams prm) {
    const double cnt = scal[C_SUMS], dsum = scal[C_SUMS + 1], bb = scal[C_SUMS + 2];
    scal[C_DCNT] = cnt; scal[C_DSUM] = dsum;
    scal[C_DMEAN] = cnt > 0.0 ? dsum / cnt : 1.0;
    scal[C_BB] = bb;
    const double t = prm.tol * prm.tol * bb, a = prm.atol * prm.atol;
    scal[C_TARGET2] = t > a ? t : a;
    scal[C_K] = 0.0; scal[C_DONE] = 0.0; scal[C_BETA] = 0.0;
}

// p = M r; rz = r . M r; rr = r . r; Ap = 0
template <typename T>
__global__ void __launch_bounds__(kVecThreads) pcg_init2_kernel(long long rows, const T* __restrict__ r, T* __restrict__ p,
                                                               T* __restrict__ Ap, const T* __restrict__ diag,
                                                               const uchar4* __restrict__ mask, double* scal,
                                                               double* partials, unsigned int* counter) {
    const double mean = __ldcg(scal + C_DMEAN);
    double s[2] = {0.0, 0.0};
    const T zero[4] = {(T)0, (T)0, (T)0, (T)0};
    for (long long i = blockIdx.x * (long long)kVecThreads + threadIdx.x; i < rows; i += (long long)gridDim.x * kVecThreads) {
        const uchar4 m4 = mask[i];
        const unsigned char m[4] = {m4.x, m4.y, m4.z, m4.w};
        T rv[4], dv[4], pv[4];
        ld4(r, i, rv); ld4(diag, i, dv);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double z = 0.0;
            if (c < 3 && (m[c] & APL_M_FREE)) {
                z = (double)rv[c] / fix_diag(dv[c], mean);
                s[0] += (double)rv[c] * z;
                s[1] += (double)rv[c] * (double)rv[c];
            }
            pv[c] = (T)z;
        }
        st4(p, i, pv);
        st4(Ap, i, zero);
    }
    grid_reduce<2>(s, partials, counter, scal + C_SUMS);
}

__global__ void pcg_init_done_kernel(double* scal) {
    scal[C_RZ] = scal[C_SUMS];
    scal[C_RR] = scal[C_SUMS + 1];
    if (!(scal[C_RR] > scal[C_TARGET2])) scal[C_DONE] = 1.0;    // already converged (e.g. b = 0)
    if (!isfinite(scal[C_RR])) scal[C_DONE] = 3.0;
}

template <typename T>
__global__ void __launch_bounds__(kVecThreads) pcg_dot_kernel(long long rows, const T* __restrict__ p, const T* __restrict__ Ap,
                                                             const uchar4* __restrict__ mask, double* scal,
                                                             double* partials, unsigned int* counter) {
    if (__ldcg(scal + C_DONE) != 0.0) return;
    double s[1] = {0.0};
    for (long long i = blockIdx.x * (long long)kVecThreads + threadIdx.x; i < rows; i += (long long)gridDim.x * kVecThreads) {
        const uchar4 m4 = mask[i];
        const unsigned char m[4] = {m4.x, m4.y, m4.z, m4.w};
        T pv[4], av[4];
        ld4(p, i, pv); ld4(Ap, i, av);
#pragma unroll
        for (int c = 0; c < 3; ++c)
            if (m[c] & APL_M_FREE) s[0] += (double)pv[c] * (double)av[c];
    }
    grid_reduce<1>(s, partials, counter, scal + C_PAP);
}

template <typename T>
__global__ void __launch_bounds__(kVecThreads) pcg_update_kernel(long long rows, T* __restrict__ x, T* __restrict__ r,
                                                                const T* __restrict__ p, const T* __restrict__ Ap,
                                                                const T* __restrict__ diag, const uchar4* __restrict__ mask,
                                                                double* scal, double* partials, unsigned int* counter) {
    if (__ldcg(scal + C_DONE) != 0.0) return;
    const double pAp = __ldcg(scal + C_PAP), rz = __ldcg(scal + C_RZ), mean = __ldcg(scal + C_DMEAN);
    // a direction of non-positive curvature (indefinite H) or a non-finite product ends the solve (STEP sets DONE = 3)
    const bool ok = pAp > 0.0 && isfinite(pAp) && isfinite(rz);
    const double alpha = ok ? rz / pAp : 0.0;
    double s[2] = {0.0, 0.0};
    for (long long i = blockIdx.x * (long long)kVecThreads + threadIdx.x; i < rows; i += (long long)gridDim.x * kVecThreads) {
        const uchar4 m4 = mask[i];
        const unsigned char m[4] = {m4.x, m4.y, m4.z, m4.w};
        T xv[4], rv[4], pv[4], av[4], dv[4];
        ld4(x, i, xv); ld4(r, i, rv); ld4(p, i, pv); ld4(Ap, i, av); ld4(diag, i, dv);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (!(m[c] & APL_M_FREE)) continue;
            const double xi = (double)xv[c] + alpha * (double)pv[c];
            const double ri = (double)rv[c] - alpha * (double)av[c];
            xv[c] = (T)xi; rv[c] = (T)ri;
            const double rr = (double)rv[c];            // the rounded residual is what the next iteration reads
            s[0] += rr * rr / fix_diag(dv[c], mean);
            s[1] += rr * rr;
        }
        st4(x, i, xv);
        st4(r, i, rv);
    }
    grid_reduce<2>(s, partials, counter, scal + C_SUMS);
}

__global__ void pcg_step_kernel(double* scal, PcgParams prm) {
    if (scal[C_DONE] != 0.0) return;
    const double pAp = scal[C_PAP], rz = scal[C_RZ];
    if (!(pAp > 0.0 && isfinite(pAp) && isfinite(rz))) { scal[C_DONE] = 3.0; return; }
    const double rz_new = scal[C_SUMS], rr = scal[C_SUMS + 1];
    scal[C_RZ_NEW] = rz_new;
    scal[C_RR] = rr;
    scal[C_K] += 1.0;
    scal[C_BETA] = rz > 0.0 ? rz_new / rz : 0.0;
    scal[C_RZ] = rz_new;
    if (!isfinite(rr)) scal[C_DONE] = 3.0;
    else if (!(rr > scal[C_TARGET2])) scal[C_DONE] = 1.0;
    else if (scal[C_K] >= prm.max_iters) scal[C_DONE] = 2.0;
}

template <typename T>
__global__ void __launch_bounds__(kVecThreads) pcg_direction_kernel(long long rows, const T* __restrict__ r, T* __restrict__ p,
                                                                   T* __restrict__ Ap, const T* __restrict__ diag,
                                                                   const uchar4* __restrict__ mask, const double* scal) {
    if (__ldcg(scal + C_DONE) != 0.0) return;
    const double beta = __ldcg(scal + C_BETA), mean = __ldcg(scal + C_DMEAN);
    const T zero[4] = {(T)0, (T)0, (T)0, (T)0};
    for (long long i = blockIdx.x * (long long)kVecThreads + threadIdx.x; i < rows; i += (long long)gridDim.x * kVecThreads) {
        const uchar4 m4 = mask[i];
        const unsigned char m[4] = {m4.x, m4.y, m4.z, m4.w};
        T rv[4], pv[4], dv[4];
        ld4(r, i, rv); ld4(p, i, pv); ld4(diag, i, dv);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double pi = 0.0;
            if (c < 3 && (m[c] & APL_M_FREE)) pi = (double)rv[c] / fix_diag(dv[c], mean) + beta * (double)pv[c];
            pv[c] = (T)pi;
        }
        st4(p, i, pv);
        st4(Ap, i, zero);
    }
}

}  // namespace apl

using namespace apl;

struct apl_pcg {
    int dtype = 0, device = 0;
    int64_t rows = 0;
    const void* u = nullptr;     // the state the Hessian is evaluated at, (rows, 4)
    void* x = nullptr;
    const void* b = nullptr;
    void* r = nullptr;
    void* p = nullptr;
    void* Ap = nullptr;
    const void* diag = nullptr;
    const uint8_t* mask = nullptr;
    double* scal = nullptr;
    double* partials = nullptr;
    unsigned int* counter = nullptr;
    int grid = 1, scatter = APL_SCATTER_TILE, psd = 0, use_graph = 1;
    PcgParams prm{1e-5, 1e-15, 1000.0};
    std::vector<apl_fem*> fems;
    cudaGraphExec_t graph = nullptr;
    cudaStream_t capture_stream = nullptr;
};

namespace {

int matvec(apl_pcg* w, cudaStream_t s, bool guarded) {
    const int ops = APL_OP_HESS_PROD | (w->psd ? APL_OP_PSD : 0);
    for (apl_fem* f : w->fems) {
        int rc = fem_eval_pncg(f, ops, w->u, w->p, nullptr, w->scal, 0, guarded ? C_DONE : -1, -1, nullptr, nullptr, nullptr,
                               nullptr, w->scatter, s, 0, w->Ap);
        if (rc != APL_OK) return rc;
    }
    return APL_OK;
}

template <typename T>
int iteration(apl_pcg* w, cudaStream_t s) {
    const uchar4* mask = (const uchar4*)w->mask;
    int rc = matvec(w, s, true);
    if (rc != APL_OK) return rc;
    pcg_dot_kernel<T><<<w->grid, kVecThreads, 0, s>>>(w->rows, (const T*)w->p, (const T*)w->Ap, mask, w->scal, w->partials, w->counter);
    pcg_update_kernel<T><<<w->grid, kVecThreads, 0, s>>>(w->rows, (T*)w->x, (T*)w->r, (const T*)w->p, (const T*)w->Ap,
                                                        (const T*)w->diag, mask, w->scal, w->partials, w->counter);
    pcg_step_kernel<<<1, 1, 0, s>>>(w->scal, w->prm);
    pcg_direction_kernel<T><<<w->grid, kVecThreads, 0, s>>>(w->rows, (const T*)w->r, (T*)w->p, (T*)w->Ap, (const T*)w->diag, mask,
                                                           w->scal);
    APL_CUDA_CHECK(cudaGetLastError());
    return APL_OK;
}

template <typename T>
int init(apl_pcg* w, int x_is_zero, cudaStream_t s) {
    const uchar4* mask = (const uchar4*)w->mask;
    APL_CUDA_CHECK(cudaMemsetAsync(w->scal, 0, sizeof(double) * APL_PCG_NSCAL, s));
    APL_CUDA_CHECK(cudaMemsetAsync(w->Ap, 0, (size_t)w->rows * 4 * sizeof(T), s));
    if (x_is_zero) {
        APL_CUDA_CHECK(cudaMemsetAsync(w->x, 0, (size_t)w->rows * 4 * sizeof(T), s));
    } else {
        // Ap = H x0: the matvec reads its direction from w->p, so x0 is staged there (p is rebuilt below)
        APL_CUDA_CHECK(cudaMemcpyAsync(w->p, w->x, (size_t)w->rows * 4 * sizeof(T), cudaMemcpyDeviceToDevice, s));
        int rc = matvec(w, s, false);
        if (rc != APL_OK) return rc;
    }
    pcg_init1_kernel<T><<<w->grid, kVecThreads, 0, s>>>(w->rows, (const T*)w->b, (const T*)w->Ap, (T*)w->r, (const T*)w->diag, mask,
                                                       w->scal, w->partials, w->counter);
    pcg_init_scalars_kernel<<<1, 1, 0, s>>>(w->scal, w->prm);
    pcg_init2_kernel<T><<<w->grid, kVecThreads, 0, s>>>(w->rows, (const T*)w->r, (T*)w->p, (T*)w->Ap, (const T*)w->diag, mask,
                                                       w->scal, w->partials, w->counter);
    pcg_init_done_kernel<<<1, 1, 0, s>>>(w->scal);
    APL_CUDA_CHECK(cudaGetLastError());
    return APL_OK;
}

void drop_graph(apl_pcg* w) {
    if (w->graph) { cudaGraphExecDestroy(w->graph); w->graph = nullptr; }
}

}  // namespace

extern "C" {

int apl_pcg_create(int dtype, int64_t n_points, int device, const void* u, void* x, const void* b, void* r, void* p, void* Ap,
                   const void* diag, const uint8_t* mask, double* scal, apl_pcg_t** out) {
    if (!out) { set_error("apl_pcg_create: out is NULL"); return APL_ERR_INVALID; }
    *out = nullptr;
    if ((dtype != APL_F32 && dtype != APL_F64) || n_points <= 0 || !u || !x || !b || !r || !p || !Ap || !diag || !mask || !scal) {
        set_error("apl_pcg_create: bad arguments");
        return APL_ERR_INVALID;
    }
    APL_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    APL_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    apl_pcg* w = new apl_pcg();
    w->dtype = dtype; w->device = device; w->rows = n_points;
    w->u = u; w->x = x; w->b = b; w->r = r; w->p = p; w->Ap = Ap; w->diag = diag; w->mask = mask; w->scal = scal;
    const long long need = (n_points + kVecThreads - 1) / kVecThreads, cap = (long long)prop.multiProcessorCount * 4;
    w->grid = (int)(need < 1 ? 1 : (need > cap ? cap : need));
    cudaError_t e = cudaMalloc((void**)&w->partials, sizeof(double) * 4 * (size_t)cap);
    if (e == cudaSuccess) e = cudaMalloc((void**)&w->counter, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(w->counter, 0, sizeof(unsigned int));
    if (e != cudaSuccess) { cudaFree(w->partials); cudaFree(w->counter); delete w; set_error(cudaGetErrorString(e)); return APL_ERR_CUDA; }
    *out = w;
    return APL_OK;
}

void apl_pcg_destroy(apl_pcg_t* w) {
    if (!w) return;
    drop_graph(w);
    if (w->capture_stream) cudaStreamDestroy(w->capture_stream);
    cudaFree(w->partials);
    cudaFree(w->counter);
    delete w;
}

int apl_pcg_add_fem(apl_pcg_t* w, apl_fem_t* fem) {
    if (!w || !fem) { set_error("apl_pcg_add_fem: NULL argument"); return APL_ERR_INVALID; }
    if (fem->dtype != w->dtype || fem->device != w->device || fem->host.n_points > w->rows) {
        set_error("apl_pcg_add_fem: potential does not match the workspace (dtype, device or n_points)");
        return APL_ERR_INVALID;
    }
    w->fems.push_back(fem);
    drop_graph(w);
    return APL_OK;
}

int apl_pcg_set_params(apl_pcg_t* w, double tol, double atol, int64_t max_iters, int psd, int scatter, int use_graph) {
    if (!w || max_iters < 0 || tol < 0 || atol < 0) { set_error("apl_pcg_set_params: bad arguments"); return APL_ERR_INVALID; }
    if (psd && scatter != APL_SCATTER_TILE) { set_error("apl_pcg_set_params: the PSD projection needs the TILE assembly"); return APL_ERR_INVALID; }
    w->prm.tol = tol; w->prm.atol = atol; w->prm.max_iters = (double)max_iters;
    w->psd = psd ? 1 : 0; w->scatter = scatter; w->use_graph = use_graph ? 1 : 0;
    drop_graph(w);
    return APL_OK;
}

int apl_pcg_init(apl_pcg_t* w, int x_is_zero, void* stream) {
    if (!w) { set_error("apl_pcg_init: NULL workspace"); return APL_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    return w->dtype == APL_F32 ? init<float>(w, x_is_zero, s) : init<double>(w, x_is_zero, s);
}

int apl_pcg_iterate(apl_pcg_t* w, int n_iters, void* stream) {
    if (!w || n_iters < 0) { set_error("apl_pcg_iterate: bad arguments"); return APL_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    auto one = [&](cudaStream_t st) { return w->dtype == APL_F32 ? iteration<float>(w, st) : iteration<double>(w, st); };
    if (w->use_graph && !w->graph) {
        if (!w->capture_stream) APL_CUDA_CHECK(cudaStreamCreateWithFlags(&w->capture_stream, cudaStreamNonBlocking));
        cudaGraph_t g = nullptr;
        APL_CUDA_CHECK(cudaStreamBeginCapture(w->capture_stream, cudaStreamCaptureModeThreadLocal));
        int rc = one(w->capture_stream);
        cudaError_t e = cudaStreamEndCapture(w->capture_stream, &g);
        if (rc != APL_OK) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) { set_error(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e)); return APL_ERR_CUDA; }
        e = cudaGraphInstantiate(&w->graph, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { set_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); return APL_ERR_CUDA; }
    }
    for (int it = 0; it < n_iters; ++it) {
        if (w->use_graph) APL_CUDA_CHECK(cudaGraphLaunch(w->graph, s));
        else if (int rc = one(s)) return rc;
    }
    return APL_OK;
}

}  // extern "C"
