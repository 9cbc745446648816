// C ABI of apple_b200: handle management, static-data packing and upload, operator dispatch.
// Interface documentation lives in include/apple_b200.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>

#include <cmath>
#include <cstring>

#include "common.h"
#include "fem_kernels.cuh"

namespace apl {
const char* last_error_cstr();

template <typename T, int KIND>
int launch_fem(const apl_fem* fem, int ops, const FemArgs<T>& args, int scatter, cudaStream_t stream);

static int rec_size(int kind) { return kind == APL_KIND_SNH_MUSCLE ? 18 : (kind == APL_KIND_SNH_ARAP ? 14 : 12); }

// Packs the caller-order reference arrays into [nplanes][plane_stride] 16-byte vectors in packed
// (tile) order.  Record = D (rows 1..3 of dhdX), vol, mu, lambda, activation[6].
template <typename T>
static int pack_planes(apl_fem* f, const T* dhdX, const T* dV, const T* mu, const T* la, const T* act,
                       const T* dV2, const T* mu2, std::vector<T>& planes) {
    constexpr int VEC = 16 / (int)sizeof(T);
    const int64_t n = f->host.n_packed();
    const int nrec = f->nrec;
    planes.assign((size_t)f->nplanes * f->plane_stride * VEC, (T)0);
    for (int64_t pos = 0; pos < n; ++pos) {
        const int64_t c = f->host.order[(size_t)pos];
        T rec[20] = {0};
        if (dhdX) {
            const T* d = dhdX + 12 * c;
            double mx = 0, dev = 0;
            for (int J = 0; J < 3; ++J) {
                dev = std::max(dev, std::fabs((double)d[J] + d[3 + J] + d[6 + J] + d[9 + J]));
                for (int a = 0; a < 4; ++a) mx = std::max(mx, std::fabs((double)d[3 * a + J]));
            }
            const double tol = (sizeof(T) == 4 ? 1e-3 : 1e-9) * mx;
            if (!(dev <= tol)) {
                set_error("dhdX rows of cell " + std::to_string(c) +
                          " do not sum to zero (only linear tetrahedra are supported)");
                return APL_ERR_MESH;
            }
            // rows 1..3 of dhdX (row 0 is minus their sum)
            for (int k = 0; k < 9; ++k) rec[k] = d[3 + k];
        }
        rec[9] = dV ? dV[c] : (T)0;
        rec[10] = mu ? mu[c] : (T)0;
        rec[11] = la ? la[c] : (T)0;
        if (act)
            for (int k = 0; k < 6; ++k) rec[12 + k] = act[6 * c + k];
        if (f->kind == APL_KIND_SNH_ARAP) {  // second potential on the same cells
            rec[12] = dV2 ? dV2[c] : (T)0;
            rec[13] = mu2 ? mu2[c] : (T)0;
        }
        for (int k = 0; k < nrec; ++k) {
            const int plane = k / VEC, lane = k % VEC;
            planes[((size_t)plane * f->plane_stride + pos) * VEC + lane] = rec[k];
        }
    }
    return APL_OK;
}

template <typename T>
static int upload_planes(apl_fem* f, const void* dhdX, const void* dV, const void* mu, const void* la,
                         const void* act, const void* dV2 = nullptr, const void* mu2 = nullptr) {
    std::vector<T> planes;
    int rc = pack_planes<T>(f, (const T*)dhdX, (const T*)dV, (const T*)mu, (const T*)la, (const T*)act,
                            (const T*)dV2, (const T*)mu2, planes);
    if (rc != APL_OK) return rc;
    if (f->device >= 0) {
        APL_CUDA_CHECK(cudaMemcpy(f->d_planes, planes.data(), planes.size() * sizeof(T), cudaMemcpyHostToDevice));
    } else {  // host-only handle: keep the packed planes for inspection (apl_fem_host_planes)
        const unsigned char* b = reinterpret_cast<const unsigned char*>(planes.data());
        f->host_planes.assign(b, b + planes.size() * sizeof(T));
    }
    return APL_OK;
}

struct PncgExtras {
    const void* axpy_p = nullptr;
    const double* scal = nullptr;
    int alpha_idx = 0, skip_a = -1, skip_b = -1, dyn_j = 0;
    double* fun_d = nullptr;
    double* quad_d = nullptr;
};

template <typename T>
static int eval_typed(apl_fem* f, int ops, const void* u, const void* p, int ld_in, void* fun, void* quad,
                      void* grad, void* diag, void* prod, int ld_out, int scatter, cudaStream_t stream,
                      const PncgExtras* ex = nullptr, int part = APL_PART_ALL) {
    FemArgs<T> a;
    if (ex) {
        a.axpy_p = (const T*)ex->axpy_p;
        a.scal = ex->scal;
        a.alpha_idx = ex->alpha_idx;
        a.skip_a = ex->skip_a;
        a.skip_b = ex->skip_b;
        a.dyn_j = ex->dyn_j;
        a.fun_d = (ops & APL_OP_FUN) ? ex->fun_d : nullptr;
        a.quad_d = (ops & APL_OP_HESS_QUAD) ? ex->quad_d : nullptr;
    }
    {
        const int64_t nt = f->host.n_tiles(), nb = f->n_boundary_tiles;
        const int64_t begin = part == APL_PART_INTERIOR ? nb : 0;
        const int64_t count = part == APL_PART_ALL ? nt : (part == APL_PART_BOUNDARY ? nb : nt - nb);
        a.tiles = (const int4*)f->d_tiles + begin;
        a.n_tiles = (int)count;
    }
    a.conn = (const uchar4*)f->d_conn;
    a.slots = (const ushort4*)f->d_slots;
    a.tile_verts = (const int*)f->d_tile_verts;
    a.tile_voff = (const unsigned short*)f->d_tile_voff;
    a.tile_vperm = (const unsigned char*)f->d_tile_vperm;
    a.planes = (const uint4*)f->d_planes;
    a.plane_stride = f->plane_stride;
    a.u = (const T*)u;
    a.p = (const T*)p;
    a.ld_in = ld_in;
    a.grad = (ops & APL_OP_GRAD) ? (T*)grad : nullptr;
    a.diag = (ops & APL_OP_HESS_DIAG) ? (T*)diag : nullptr;
    a.prod = (ops & (APL_OP_HESS_PROD | APL_OP_HESS_OFFD)) ? (T*)prod : nullptr;
    a.ld_out = ld_out;
    a.fun = (ops & APL_OP_FUN) ? (T*)fun : nullptr;
    a.quad = (ops & APL_OP_HESS_QUAD) ? (T*)quad : nullptr;
    a.partials = f->d_partials;
    a.counter = f->d_counter;
    switch (f->kind) {
        case APL_KIND_SNH: return launch_fem<T, APL_KIND_SNH>(f, ops, a, scatter, stream);
        case APL_KIND_ARAP: return launch_fem<T, APL_KIND_ARAP>(f, ops, a, scatter, stream);
        case APL_KIND_SNH_ARAP: return launch_fem<T, APL_KIND_SNH_ARAP>(f, ops, a, scatter, stream);
        default: return launch_fem<T, APL_KIND_SNH_MUSCLE>(f, ops, a, scatter, stream);
    }
}

// ---- small kernels: external force, field copy --------------------------------------------------

template <typename T>
__global__ void ext_force_kernel(int ops, long long k, const T* __restrict__ force,
                                 const int* __restrict__ indices, const T* __restrict__ u, int ld_in, T* fun,
                                 T* grad, int ld_out, const T* __restrict__ axpy_p, const double* scal,
                                 int alpha_idx, int skip_a, int skip_b, double* fun_d, int dyn_j) {
    // warp/potential/_ext_force.py:17-39: W = -f.u[vid] (atomic to out[0]); grad[vid] -= f
    const int joff = (scal && dyn_j) ? (int)__ldcg(scal + APL_S_J) : 0;
    if (scal) {
        if (skip_a >= 0 && __ldcg(scal + skip_a) != 0.0) return;
        if (skip_b >= 0 && __ldcg(scal + skip_b + joff) != 0.0) return;
    }
    const T alpha = axpy_p ? (T)__ldcg(scal + alpha_idx + joff) : (T)0;
    if (fun_d) fun_d += joff;
    double w = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < k;
         i += (long long)gridDim.x * blockDim.x) {
        const int vid = indices[i];
        const T fx = force[3 * i], fy = force[3 * i + 1], fz = force[3 * i + 2];
        if (ops & APL_OP_FUN) {
            const T* r = u + (long long)ld_in * vid;
            T ux = r[0], uy = r[1], uz = r[2];
            if (axpy_p) {
                const T* d = axpy_p + (long long)ld_in * vid;
                ux += alpha * d[0]; uy += alpha * d[1]; uz += alpha * d[2];
            }
            w -= (double)(fx * ux + fy * uy + fz * uz);
        }
        if (ops & APL_OP_GRAD) {
            T* g = grad + (long long)ld_out * vid;
            atomicAdd(g, -fx);
            atomicAdd(g + 1, -fy);
            atomicAdd(g + 2, -fz);
        }
    }
    if (ops & APL_OP_FUN) {
        w = warp_sum(w);
        __shared__ double red[32];
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) red[wid] = w;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
            if (fun) atomicAdd(fun, (T)s);
            if (fun_d) atomicAdd(fun_d, s);
        }
    }
}

template <typename T>
__global__ void field_copy_kernel(long long n, const T* __restrict__ src, int ld_src, T* __restrict__ dst,
                                  int ld_dst) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const T x = src[i * ld_src], y = src[i * ld_src + 1], z = src[i * ld_src + 2];
        T* d = dst + i * ld_dst;
        d[0] = x; d[1] = y; d[2] = z;
        if (ld_dst == 4) d[3] = (T)0;
    }
}

// ---- halo exchange helpers (sharded meshes) ---------------------------------------------------------
// pack:   send[i, f*3 + c] = field_f[send_index[i], c]                       (rows shared with other ranks)
// unpack: for every shared vertex, sum its partials in ascending RANK order (own partial in its own
//         position), so that all replicas end up bit-identical, and write the total back.
template <typename T>
__global__ void halo_pack_kernel(long long n, const long long* __restrict__ index, int nf, const T* f0, const T* f1,
                                 const T* f2, int ld, T* __restrict__ send) {
    const T* f[3] = {f0, f1, f2};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const long long v = index[i];
        for (int k = 0; k < nf; ++k) {
            const T* r = f[k] + v * ld;
            T* d = send + (i * nf + k) * 3;
            d[0] = r[0]; d[1] = r[1]; d[2] = r[2];
        }
    }
}

template <typename T>
__global__ void halo_unpack_kernel(long long n_shared, const long long* __restrict__ shared,
                                   const int* __restrict__ row_ptr, const long long* __restrict__ src, int nf, T* f0,
                                   T* f1, T* f2, int ld, const T* __restrict__ recv) {
    T* f[3] = {f0, f1, f2};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_shared;
         i += (long long)gridDim.x * blockDim.x) {
        const long long v = shared[i];
        for (int k = 0; k < nf; ++k) {
            T* own = f[k] + v * ld;
            T a0 = 0, a1 = 0, a2 = 0;
            for (int e = row_ptr[i]; e < row_ptr[i + 1]; ++e) {
                const long long j = src[e];  // -1: this rank's own partial, else a row of the receive buffer
                const T* r = j < 0 ? own : recv + (j * nf + k) * 3;
                a0 += r[0]; a1 += r[1]; a2 += r[2];
            }
            own[0] = a0; own[1] = a1; own[2] = a2;
        }
    }
}

// ---- mixed derivative product (per cell, no assembly) ------------------------------------------------
// One thread per packed tet position; gathers straight from global memory (a setup / adjoint-time operator, not
// the hot path).
template <typename T, int KIND>
__global__ void __launch_bounds__(256) fem_mixed_kernel(const int4* __restrict__ tiles, int n_tiles,
                                                        const unsigned char* __restrict__ conn,
                                                        const int* __restrict__ tile_verts, const uint4* __restrict__ planes,
                                                        long long plane_stride, const int* __restrict__ order,
                                                        const T* __restrict__ u, const T* __restrict__ p, int ld, T* d_mu,
                                                        T* d_la, T* d_act) {
    constexpr int NREC = RecSize<KIND>::value;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int4 h = __ldg(tiles + tile);
        const int n_tets = h.y & 0xffff;
        const int j = threadIdx.x;
        if (j >= n_tets) continue;
        const long long pos = (long long)h.x + j;
        const int cell = __ldg(order + pos);
        const unsigned char* cc = conn + 4ll * pos;
        const int l[4] = {cc[0], cc[1], cc[2], cc[3]};
        Rec<T, NREC> rec;
#pragma unroll
        for (int k = 0; k < Rec<T, NREC>::NPL; ++k) rec.q[k] = __ldg(planes + k * plane_stride + pos);
        T uc[4][3], pc[4][3];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const long long gv = __ldg(tile_verts + h.z + l[c]);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                uc[c][i] = __ldg(u + gv * ld + i);
                pc[c][i] = __ldg(p + gv * ld + i);
            }
        }
        T m, la, act[6];
        elem_mixed<T, KIND>(rec.s, uc, pc, m, la, act);
        if (d_mu) d_mu[cell] = m;
        if constexpr (KIND != APL_KIND_ARAP) { if (d_la) d_la[cell] = la; }
        if constexpr (KIND == APL_KIND_SNH_MUSCLE) {
            if (d_act)
#pragma unroll
                for (int k = 0; k < 6; ++k) d_act[6ll * cell + k] = act[k];
        }
    }
}

template <typename T>
static int mixed_typed(apl_fem* f, const void* u, const void* p, int ld, void* d_mu, void* d_la, void* d_act,
                       cudaStream_t stream) {
    const int n_tiles = (int)f->host.n_tiles();
    if (n_tiles == 0) return APL_OK;
    int grid = f->num_sms * 8;
    if (grid > n_tiles) grid = n_tiles;
#define APL_MIXED(K)                                                                                              \
    fem_mixed_kernel<T, K><<<grid, 256, 0, stream>>>((const int4*)f->d_tiles, n_tiles,                             \
                                                     (const unsigned char*)f->d_conn, (const int*)f->d_tile_verts, \
                                                     (const uint4*)f->d_planes, f->plane_stride, f->d_order,       \
                                                     (const T*)u, (const T*)p, ld, (T*)d_mu, (T*)d_la, (T*)d_act)
    switch (f->kind) {
        case APL_KIND_SNH: APL_MIXED(APL_KIND_SNH); break;
        case APL_KIND_ARAP: APL_MIXED(APL_KIND_ARAP); break;
        default: APL_MIXED(APL_KIND_SNH_MUSCLE); break;
    }
#undef APL_MIXED
    APL_CUDA_CHECK(cudaGetLastError());
    return APL_OK;
}

static int grid_for(long long n, int block) {
    long long g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > 148 * 16) g = 148 * 16;
    return (int)g;
}

// ---- entry points used by the native PNCG driver (pncg.cu) ------------------------------------------

int fem_eval_pncg(apl_fem* f, int ops, const void* x, const void* p, const void* axpy_p, double* scal,
                  int alpha_idx, int skip_a, int skip_b, double* fun_d, double* quad_d, void* grad, void* diag,
                  int scatter, cudaStream_t stream, int dyn_j, void* third) {
    PncgExtras ex;
    ex.axpy_p = axpy_p; ex.scal = scal; ex.alpha_idx = alpha_idx; ex.skip_a = skip_a; ex.skip_b = skip_b;
    ex.dyn_j = dyn_j;
    ex.fun_d = fun_d; ex.quad_d = quad_d;
    if (f->kind == APL_KIND_ARAP) ops &= ~APL_OP_PSD;
    // `third`: the hess_prod output or the vertex-block off-diagonals (APL_OP_HESS_OFFD), ld = 4 like the others
    return f->dtype == APL_F32
               ? eval_typed<float>(f, ops, x, p, 4, nullptr, nullptr, grad, diag, third, 4, scatter, stream, &ex)
               : eval_typed<double>(f, ops, x, p, 4, nullptr, nullptr, grad, diag, third, 4, scatter, stream, &ex);
}

int ext_force_pncg(int dtype, int ops, int64_t k, const void* force, const int32_t* indices, const void* x,
                   const void* axpy_p, double* scal, int alpha_idx, int skip_a, int skip_b, double* fun_d,
                   void* grad, cudaStream_t stream, int dyn_j) {
    if (k == 0) return APL_OK;
    const int block = 256, grid = grid_for(k, block);
    if (dtype == APL_F32)
        ext_force_kernel<float><<<grid, block, 0, stream>>>(ops, k, (const float*)force, indices, (const float*)x, 4,
                                                            nullptr, (float*)grad, 4, (const float*)axpy_p, scal,
                                                            alpha_idx, skip_a, skip_b, fun_d, dyn_j);
    else
        ext_force_kernel<double><<<grid, block, 0, stream>>>(ops, k, (const double*)force, indices, (const double*)x,
                                                             4, nullptr, (double*)grad, 4, (const double*)axpy_p, scal,
                                                             alpha_idx, skip_a, skip_b, fun_d, dyn_j);
    APL_CUDA_CHECK(cudaGetLastError());
    return APL_OK;
}

}  // namespace apl

namespace apl {
// Device copies of a handle's packed tables (tile headers, connectivity, slots, vertex tables) and the
// allocations that do not depend on the static record (planes, scalar partials).  f->host, f->nplanes,
// f->plane_stride and f->device are set by the caller.
int fem_upload_tables(apl_fem* f) {
    const HostTables& h = f->host;
    const int device = f->device;
    const size_t plane_bytes = (size_t)f->nplanes * f->plane_stride * 16;
    f->static_bytes = (int64_t)(plane_bytes + h.n_tiles() * 16 + h.conn.size() + h.slots.size() * 2 +
                                h.tile_verts.size() * 5 + h.tile_voff.size() * 2);
    APL_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    APL_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    f->num_sms = prop.multiProcessorCount;
    f->max_grid = f->num_sms * 16;
    auto up = [&](void** dst, const void* src, size_t bytes, size_t slack) -> cudaError_t {
        // `slack` zeroed bytes after the data: 16-byte granular bulk copies of the last tile stay in bounds
        cudaError_t e = cudaMalloc(dst, bytes + slack + 16);
        if (e != cudaSuccess) return e;
        if (slack) {
            e = cudaMemset((char*)*dst + bytes, 0, slack + 16);
            if (e != cudaSuccess) return e;
        }
        return bytes ? cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) : cudaSuccess;
    };
    {
        // device tile header: tet_start, n_tets | n_verts << 16, vert_start, voff_start
        std::vector<int32_t> hdr((size_t)h.n_tiles() * 4);
        for (int64_t t = 0; t < h.n_tiles(); ++t) {
            const int32_t* src = h.tiles.data() + 6 * t;
            hdr[4 * t] = src[0];
            hdr[4 * t + 1] = src[1] | (src[3] << 16);
            hdr[4 * t + 2] = src[2];
            hdr[4 * t + 3] = src[4];
        }
        APL_CUDA_CHECK(up(&f->d_tiles, hdr.data(), hdr.size() * 4, 0));
    }
    APL_CUDA_CHECK(up(&f->d_conn, h.conn.data(), h.conn.size(), 64));
    APL_CUDA_CHECK(up(&f->d_slots, h.slots.data(), h.slots.size() * 2, 64));
    APL_CUDA_CHECK(up(&f->d_tile_verts, h.tile_verts.data(), h.tile_verts.size() * 4, 0));
    APL_CUDA_CHECK(up(&f->d_tile_voff, h.tile_voff.data(), h.tile_voff.size() * 2, 0));
    APL_CUDA_CHECK(up(&f->d_tile_vperm, h.tile_vperm.data(), h.tile_vperm.size(), 0));
    APL_CUDA_CHECK(cudaMalloc(&f->d_planes, plane_bytes));
    APL_CUDA_CHECK(cudaMalloc((void**)&f->d_partials, sizeof(double) * 2 * f->max_grid));
    APL_CUDA_CHECK(cudaMalloc((void**)&f->d_counter, sizeof(unsigned int)));
    APL_CUDA_CHECK(cudaMemset(f->d_counter, 0, sizeof(unsigned int)));
    return APL_OK;
}
}  // namespace apl

using namespace apl;

extern "C" {

int apl_version(void) { return 100; }

const char* apl_last_error(void) { return last_error_cstr(); }

int apl_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        set_error(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
        return APL_ERR_CUDA;
    }
    return n;
}

void apl_fem_destroy(apl_fem_t* f) {
    if (!f) return;
    if (f->device >= 0) {
        cudaFree(f->d_tiles);
        cudaFree(f->d_conn);
        cudaFree(f->d_slots);
        cudaFree(f->d_tile_verts);
        cudaFree(f->d_tile_voff);
        cudaFree(f->d_tile_vperm);
        cudaFree(f->d_planes);
        cudaFree(f->d_order);
        cudaFree(f->d_partials);
        cudaFree(f->d_counter);
    }
    delete f;
}

static int fem_create_impl(int kind, int dtype, int64_t n_cells, int64_t n_points, const int32_t* cells,
                           const void* dhdX, const void* dV, const void* mu, const void* lambda_,
                           const void* activation, const void* dV2, const void* mu2, const double* points, int device,
                           apl_fem_t** out) {
    if (!out) { set_error("apl_fem_create: out is NULL"); return APL_ERR_INVALID; }
    *out = nullptr;
    if (kind < 0 || kind > 3 || (dtype != APL_F32 && dtype != APL_F64)) {
        set_error("apl_fem_create: unknown kind or dtype");
        return APL_ERR_INVALID;
    }
    if (!dhdX || !dV || !mu || (kind != APL_KIND_ARAP && !lambda_) || (kind == APL_KIND_SNH_MUSCLE && !activation) ||
        (kind == APL_KIND_SNH_ARAP && (!dV2 || !mu2))) {
        set_error("apl_fem_create: a required array (dhdX, dV, mu, lambda_, activation) is NULL");
        return APL_ERR_INVALID;
    }
    apl_fem* f = new apl_fem();
    f->kind = kind;
    f->dtype = dtype;
    f->device = device;
    f->nrec = rec_size(kind);
    const int vec = dtype == APL_F32 ? 4 : 2;
    f->nplanes = (f->nrec + vec - 1) / vec;
    int rc = build_tiles(n_cells, n_points, cells, points, dtype == APL_F32 ? 4 : 8, f->host);
    if (rc != APL_OK) { delete f; return rc; }
    f->plane_stride = (f->host.n_packed() + 31) / 32 * 32;
    if (f->plane_stride == 0) f->plane_stride = 32;
    const HostTables& h = f->host;
    const size_t plane_bytes = (size_t)f->nplanes * f->plane_stride * 16;
    f->static_bytes = (int64_t)(plane_bytes + h.n_tiles() * 16 + h.conn.size() + h.slots.size() * 2 +
                                h.tile_verts.size() * 5 + h.tile_voff.size() * 2);
    if (device >= 0) {
        rc = apl::fem_upload_tables(f);
        if (rc != APL_OK) { apl_fem_destroy(f); return rc; }
    }
    rc = (dtype == APL_F32) ? upload_planes<float>(f, dhdX, dV, mu, lambda_, activation, dV2, mu2)
                            : upload_planes<double>(f, dhdX, dV, mu, lambda_, activation, dV2, mu2);
    if (rc != APL_OK) { apl_fem_destroy(f); return rc; }
    *out = f;
    return APL_OK;
}

int apl_fem_create(int kind, int dtype, int64_t n_cells, int64_t n_points, const int32_t* cells,
                   const void* dhdX, const void* dV, const void* mu, const void* lambda_,
                   const void* activation, const double* points, int device, apl_fem_t** out) {
    if (kind == APL_KIND_SNH_ARAP) {
        set_error("apl_fem_create: use apl_fem_create_snh_arap for the fused kind");
        return APL_ERR_INVALID;
    }
    return fem_create_impl(kind, dtype, n_cells, n_points, cells, dhdX, dV, mu, lambda_, activation, nullptr, nullptr,
                           points, device, out);
}

int apl_fem_create_snh_arap(int dtype, int64_t n_cells, int64_t n_points, const int32_t* cells, const void* dhdX,
                            const void* dV_snh, const void* mu_snh, const void* lambda_snh, const void* dV_arap,
                            const void* mu_arap, const double* points, int device, apl_fem_t** out) {
    return fem_create_impl(APL_KIND_SNH_ARAP, dtype, n_cells, n_points, cells, dhdX, dV_snh, mu_snh, lambda_snh,
                           nullptr, dV_arap, mu_arap, points, device, out);
}

int apl_fem_info(const apl_fem_t* f, int64_t info[10]) {
    if (!f || !info) { set_error("apl_fem_info: NULL argument"); return APL_ERR_INVALID; }
    info[0] = f->host.n_cells;
    info[1] = f->host.n_points;
    info[2] = f->host.n_tiles();
    info[3] = (int64_t)f->host.tile_verts.size();
    info[4] = f->static_bytes;
    info[5] = f->kind;
    info[6] = f->dtype;
    info[7] = f->device;
    info[8] = (int64_t)f->host.tile_voff.size();
    info[9] = f->host.n_packed();
    return APL_OK;
}

int apl_fem_host_tables(const apl_fem_t* f, int32_t* tiles, int64_t* order, uint8_t* conn, uint16_t* slots,
                        int32_t* tile_verts, uint16_t* tile_voff, uint8_t* tile_vperm) {
    if (!f) { set_error("apl_fem_host_tables: NULL handle"); return APL_ERR_INVALID; }
    const HostTables& h = f->host;
    if (tiles) memcpy(tiles, h.tiles.data(), h.tiles.size() * 4);
    if (order) memcpy(order, h.order.data(), h.order.size() * 8);
    if (conn) memcpy(conn, h.conn.data(), h.conn.size());
    if (slots) memcpy(slots, h.slots.data(), h.slots.size() * 2);
    if (tile_verts) memcpy(tile_verts, h.tile_verts.data(), h.tile_verts.size() * 4);
    if (tile_voff) memcpy(tile_voff, h.tile_voff.data(), h.tile_voff.size() * 2);
    if (tile_vperm) memcpy(tile_vperm, h.tile_vperm.data(), h.tile_vperm.size());
    return APL_OK;
}

int apl_fem_host_planes(const apl_fem_t* f, void* planes, int64_t* n_planes, int64_t* plane_stride) {
    if (!f) { set_error("apl_fem_host_planes: NULL handle"); return APL_ERR_INVALID; }
    if (f->device >= 0) { set_error("apl_fem_host_planes: only for host-only handles (device = -1)"); return APL_ERR_STATE; }
    if (n_planes) *n_planes = f->nplanes;
    if (plane_stride) *plane_stride = f->plane_stride;
    if (planes) memcpy(planes, f->host_planes.data(), f->host_planes.size());
    return APL_OK;
}

int apl_fem_set_materials(apl_fem_t* f, const void* dV, const void* mu, const void* lambda_,
                          const void* activation) {
    if (!f) { set_error("apl_fem_set_materials: NULL handle"); return APL_ERR_INVALID; }
    if (f->kind == APL_KIND_SNH_ARAP) { set_error("apl_fem_set_materials: not supported for fused potentials"); return APL_ERR_STATE; }
    // Read back (device handles) or take the kept copy (host-only handles), patch the requested columns,
    // upload / keep.  Setup-time path, not hot.
    const int vec = f->dtype == APL_F32 ? 4 : 2;
    const size_t esz = f->dtype == APL_F32 ? 4 : 8;
    const size_t bytes = (size_t)f->nplanes * f->plane_stride * 16;
    std::vector<unsigned char> tmp;
    std::vector<unsigned char>& buf = f->device >= 0 ? tmp : f->host_planes;
    if (f->device >= 0) {
        tmp.resize(bytes);
        APL_CUDA_CHECK(cudaSetDevice(f->device));
        APL_CUDA_CHECK(cudaMemcpy(buf.data(), f->d_planes, bytes, cudaMemcpyDeviceToHost));
    } else if (buf.size() != bytes) {
        set_error("apl_fem_set_materials: host-only handle without packed planes");
        return APL_ERR_STATE;
    }
    auto put = [&](int k, int64_t pos, const void* src, int64_t idx) {
        const int plane = k / vec, lane = k % vec;
        memcpy(buf.data() + (((size_t)plane * f->plane_stride + pos) * vec + lane) * esz,
               (const unsigned char*)src + (size_t)idx * esz, esz);
    };
    for (int64_t pos = 0; pos < f->host.n_packed(); ++pos) {
        const int64_t c = f->host.order[(size_t)pos];
        if (dV) put(9, pos, dV, c);
        if (mu) put(10, pos, mu, c);
        if (lambda_ && f->kind != APL_KIND_ARAP) put(11, pos, lambda_, c);
        if (activation && f->kind == APL_KIND_SNH_MUSCLE)
            for (int k = 0; k < 6; ++k) put(12 + k, pos, activation, 6 * c + k);
    }
    if (f->device >= 0) APL_CUDA_CHECK(cudaMemcpy(f->d_planes, buf.data(), bytes, cudaMemcpyHostToDevice));
    return APL_OK;
}

int apl_fem_eval_part(apl_fem_t* f, int part, int ops, const void* u, const void* p, int ld_in, void* fun, void* quad,
                      void* grad, void* diag, void* prod, int ld_out, int scatter, void* stream) {
    if (!f) { set_error("apl_fem_eval: NULL handle"); return APL_ERR_INVALID; }
    if (f->device < 0) { set_error("apl_fem_eval: handle was created host-only (device = -1)"); return APL_ERR_STATE; }
    if (part != APL_PART_ALL && part != APL_PART_BOUNDARY && part != APL_PART_INTERIOR) {
        set_error("apl_fem_eval_part: unknown part");
        return APL_ERR_INVALID;
    }
    if (ops <= 0 || ops > 127 || !(ops & 63)) { set_error("apl_fem_eval: ops must be a non-empty OR of APL_OP_*"); return APL_ERR_INVALID; }
    if ((ops & APL_OP_HESS_OFFD) && (ops & (APL_OP_HESS_PROD | APL_OP_HESS_QUAD))) {
        set_error("apl_fem_eval: APL_OP_HESS_OFFD shares the `prod` output and combines with FUN, GRAD and HESS_DIAG only");
        return APL_ERR_INVALID;
    }
    if ((ops & APL_OP_PSD) && f->kind == APL_KIND_ARAP) ops &= ~APL_OP_PSD;   // ARAP: the clamped twist rates are the projection
    if ((ld_in != 3 && ld_in != 4) || (ld_out != 3 && ld_out != 4)) {
        set_error("apl_fem_eval: leading dimensions must be 3 or 4");
        return APL_ERR_INVALID;
    }
    if (scatter != APL_SCATTER_TILE && scatter != APL_SCATTER_ATOMIC && scatter != APL_SCATTER_TILE_SIMPLE) {
        set_error("apl_fem_eval: unknown scatter mode");
        return APL_ERR_INVALID;
    }
    if (!u || ((ops & (APL_OP_HESS_PROD | APL_OP_HESS_QUAD)) && !p) || ((ops & APL_OP_FUN) && !fun) ||
        ((ops & APL_OP_HESS_QUAD) && !quad) || ((ops & APL_OP_GRAD) && !grad) ||
        ((ops & APL_OP_HESS_DIAG) && !diag) || ((ops & (APL_OP_HESS_PROD | APL_OP_HESS_OFFD)) && !prod)) {
        set_error("apl_fem_eval: an array required by `ops` is NULL");
        return APL_ERR_INVALID;
    }
    {   // vector accesses of the kernels: 16-byte rows for ld = 4, 8-byte vector REDs for fp32 ld = 3
#ifdef APL_GATHER_LDG
        const uintptr_t in_mask = ld_in == 4 ? 15u : 7u;   // the register-staged gather uses 8-byte loads
#else
        const uintptr_t in_mask = ld_in == 4 ? 15u : 0u;
#endif
        const uintptr_t out_mask = ld_out == 4 ? 15u : 7u;
        auto bad = [](const void* ptr, uintptr_t mask) { return ptr && ((uintptr_t)ptr & mask) != 0; };
        if (bad(u, in_mask) || ((ops & (APL_OP_HESS_PROD | APL_OP_HESS_QUAD)) && bad(p, in_mask)) ||
            ((ops & APL_OP_GRAD) && bad(grad, out_mask)) || ((ops & APL_OP_HESS_DIAG) && bad(diag, out_mask)) ||
            ((ops & (APL_OP_HESS_PROD | APL_OP_HESS_OFFD)) && bad(prod, out_mask))) {
            set_error("apl_fem_eval: nodal fields must be 16-byte aligned for ld = 4 and outputs 8-byte aligned for ld = 3");
            return APL_ERR_INVALID;
        }
    }
    {   // launches go to the CURRENT device: it must be the one the handle's tables live on
        int cur = -1;
        APL_CUDA_CHECK(cudaGetDevice(&cur));
        if (cur != f->device) {
            set_error("apl_fem_eval: the handle lives on device " + std::to_string(f->device) +
                      " but the current device is " + std::to_string(cur) + " (call cudaSetDevice first)");
            return APL_ERR_STATE;
        }
    }
    cudaStream_t s = (cudaStream_t)stream;
    return f->dtype == APL_F32
               ? eval_typed<float>(f, ops, u, p, ld_in, fun, quad, grad, diag, prod, ld_out, scatter, s, nullptr, part)
               : eval_typed<double>(f, ops, u, p, ld_in, fun, quad, grad, diag, prod, ld_out, scatter, s, nullptr, part);
}

int apl_fem_eval(apl_fem_t* f, int ops, const void* u, const void* p, int ld_in, void* fun, void* quad,
                 void* grad, void* diag, void* prod, int ld_out, int scatter, void* stream) {
    return apl_fem_eval_part(f, APL_PART_ALL, ops, u, p, ld_in, fun, quad, grad, diag, prod, ld_out, scatter, stream);
}

int apl_fem_mark_boundary(apl_fem_t* f, const uint8_t* vertex_flags, int64_t* n_boundary) {
    if (!f) { set_error("apl_fem_mark_boundary: NULL handle"); return APL_ERR_INVALID; }
    HostTables& h = f->host;
    const int64_t nt = h.n_tiles();
    // stable partition of the 6-int tile headers: tiles touching a flagged vertex first
    // (in packed order, whatever an earlier call left: the result only depends on the flags)
    std::vector<int64_t> by_start((size_t)nt);
    for (int64_t t = 0; t < nt; ++t) by_start[(size_t)t] = t;
    std::sort(by_start.begin(), by_start.end(),
              [&](int64_t a, int64_t b) { return h.tiles[(size_t)(6 * a)] < h.tiles[(size_t)(6 * b)]; });
    std::vector<int32_t> front, back;
    front.reserve(h.tiles.size());
    back.reserve(h.tiles.size());
    for (int64_t i = 0; i < nt; ++i) {
        const int32_t* hdr = h.tiles.data() + 6 * by_start[(size_t)i];
        bool touches = false;
        if (vertex_flags)
            for (int32_t k = 0; k < hdr[3] && !touches; ++k) touches = vertex_flags[h.tile_verts[(size_t)hdr[2] + k]] != 0;
        std::vector<int32_t>& dst = touches ? front : back;
        dst.insert(dst.end(), hdr, hdr + 6);
    }
    f->n_boundary_tiles = (int64_t)front.size() / 6;
    front.insert(front.end(), back.begin(), back.end());
    h.tiles.swap(front);
    if (n_boundary) *n_boundary = f->n_boundary_tiles;
    if (f->device >= 0 && nt > 0) {
        std::vector<int32_t> hdr((size_t)nt * 4);
        for (int64_t t = 0; t < nt; ++t) {
            const int32_t* src = h.tiles.data() + 6 * t;
            hdr[4 * t] = src[0];
            hdr[4 * t + 1] = src[1] | (src[3] << 16);
            hdr[4 * t + 2] = src[2];
            hdr[4 * t + 3] = src[4];
        }
        APL_CUDA_CHECK(cudaSetDevice(f->device));
        APL_CUDA_CHECK(cudaDeviceSynchronize());   // setup-time call: no launch may still read the old order
        APL_CUDA_CHECK(cudaMemcpy(f->d_tiles, hdr.data(), hdr.size() * 4, cudaMemcpyHostToDevice));
    }
    return APL_OK;
}

int apl_fem_mixed_derivative_prod(apl_fem_t* f, const void* u, const void* p, int ld_in, void* d_mu, void* d_lambda,
                                  void* d_activation, void* stream) {
    if (!f) { set_error("apl_fem_mixed_derivative_prod: NULL handle"); return APL_ERR_INVALID; }
    if (f->device < 0) { set_error("apl_fem_mixed_derivative_prod: handle was created host-only (device = -1)"); return APL_ERR_STATE; }
    if (f->kind == APL_KIND_SNH_ARAP) {
        set_error("apl_fem_mixed_derivative_prod: not available for fused potentials (evaluate the two parts separately)");
        return APL_ERR_STATE;
    }
    if (!u || !p || (ld_in != 3 && ld_in != 4)) { set_error("apl_fem_mixed_derivative_prod: bad arguments"); return APL_ERR_INVALID; }
    if (!f->d_order) {   // first use: packed tet position -> caller's cell
        const int64_t n = f->host.n_packed();
        std::vector<int32_t> ord((size_t)n + 1);
        for (int64_t i = 0; i < n; ++i) ord[(size_t)i] = (int32_t)f->host.order[(size_t)i];
        APL_CUDA_CHECK(cudaSetDevice(f->device));
        APL_CUDA_CHECK(cudaMalloc((void**)&f->d_order, ((size_t)n + 1) * sizeof(int32_t)));
        APL_CUDA_CHECK(cudaMemcpy(f->d_order, ord.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    cudaStream_t s = (cudaStream_t)stream;
    return f->dtype == APL_F32 ? mixed_typed<float>(f, u, p, ld_in, d_mu, d_lambda, d_activation, s)
                               : mixed_typed<double>(f, u, p, ld_in, d_mu, d_lambda, d_activation, s);
}

int apl_ext_force_eval(int dtype, int ops, int64_t k, const void* force, const int32_t* indices, const void* u,
                       int ld_in, void* fun, void* grad, int ld_out, void* stream) {
    if (k < 0 || (k > 0 && (!force || !indices))) { set_error("apl_ext_force_eval: bad arguments"); return APL_ERR_INVALID; }
    ops &= (APL_OP_FUN | APL_OP_GRAD);
    if (k == 0 || ops == 0) return APL_OK;
    if (((ops & APL_OP_FUN) && (!u || !fun)) || ((ops & APL_OP_GRAD) && !grad)) {
        set_error("apl_ext_force_eval: an array required by `ops` is NULL");
        return APL_ERR_INVALID;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int block = 256, grid = grid_for(k, block);
    if (dtype == APL_F32)
        ext_force_kernel<float><<<grid, block, 0, s>>>(ops, k, (const float*)force, indices, (const float*)u, ld_in,
                                                       (float*)fun, (float*)grad, ld_out, nullptr, nullptr, 0, -1, -1,
                                                       nullptr, 0);
    else if (dtype == APL_F64)
        ext_force_kernel<double><<<grid, block, 0, s>>>(ops, k, (const double*)force, indices, (const double*)u,
                                                        ld_in, (double*)fun, (double*)grad, ld_out, nullptr, nullptr, 0,
                                                        -1, -1, nullptr, 0);
    else { set_error("apl_ext_force_eval: unknown dtype"); return APL_ERR_INVALID; }
    APL_CUDA_CHECK(cudaGetLastError());
    return APL_OK;
}

int apl_halo_pack(int dtype, int64_t n, const int64_t* index, int nf, const void* f0, const void* f1, const void* f2,
                  int ld, void* send, void* stream) {
    if (n < 0 || nf < 1 || nf > 3 || (n > 0 && (!index || !f0 || !send)) || (ld != 3 && ld != 4)) {
        set_error("apl_halo_pack: bad arguments");
        return APL_ERR_INVALID;
    }
    if (n == 0) return APL_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int block = 256, grid = grid_for(n, block);
    if (dtype == APL_F32)
        halo_pack_kernel<float><<<grid, block, 0, s>>>(n, (const long long*)index, nf, (const float*)f0, (const float*)f1,
                                                       (const float*)f2, ld, (float*)send);
    else
        halo_pack_kernel<double><<<grid, block, 0, s>>>(n, (const long long*)index, nf, (const double*)f0,
                                                        (const double*)f1, (const double*)f2, ld, (double*)send);
    APL_CUDA_CHECK(cudaGetLastError());
    return APL_OK;
}

int apl_halo_unpack(int dtype, int64_t n_shared, const int64_t* shared, const int32_t* row_ptr, const int64_t* src,
                    int nf, void* f0, void* f1, void* f2, int ld, const void* recv, void* stream) {
    if (n_shared < 0 || nf < 1 || nf > 3 || (n_shared > 0 && (!shared || !row_ptr || !src || !f0 || !recv)) ||
        (ld != 3 && ld != 4)) {
        set_error("apl_halo_unpack: bad arguments");
        return APL_ERR_INVALID;
    }
    if (n_shared == 0) return APL_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int block = 256, grid = grid_for(n_shared, block);
    if (dtype == APL_F32)
        halo_unpack_kernel<float><<<grid, block, 0, s>>>(n_shared, (const long long*)shared, row_ptr,
                                                         (const long long*)src, nf, (float*)f0, (float*)f1, (float*)f2, ld,
                                                         (const float*)recv);
    else
        halo_unpack_kernel<double><<<grid, block, 0, s>>>(n_shared, (const long long*)shared, row_ptr,
                                                          (const long long*)src, nf, (double*)f0, (double*)f1, (double*)f2,
                                                          ld, (const double*)recv);
    APL_CUDA_CHECK(cudaGetLastError());
    return APL_OK;
}

int apl_field_copy(int dtype, int64_t n, const void* src, int ld_src, void* dst, int ld_dst, void* stream) {
    if (n < 0 || (n > 0 && (!src || !dst)) || (ld_src != 3 && ld_src != 4) || (ld_dst != 3 && ld_dst != 4)) {
        set_error("apl_field_copy: bad arguments");
        return APL_ERR_INVALID;
    }
    if (n == 0) return APL_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int block = 256, grid = grid_for(n, block);
    if (dtype == APL_F32)
        field_copy_kernel<float><<<grid, block, 0, s>>>(n, (const float*)src, ld_src, (float*)dst, ld_dst);
    else if (dtype == APL_F64)
        field_copy_kernel<double><<<grid, block, 0, s>>>(n, (const double*)src, ld_src, (double*)dst, ld_dst);
    else { set_error("apl_field_copy: unknown dtype"); return APL_ERR_INVALID; }
    APL_CUDA_CHECK(cudaGetLastError());
    return APL_OK;
}

}  // extern "C"
