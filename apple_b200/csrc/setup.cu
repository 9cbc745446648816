// Device-side setup of a FEM potential: everything the reference does once per mesh on the host in JAX --
// Region.compute_grad (jax/fem/region/_region.py:84-108: dXdr, drdX, dV = det / 6, dhdX = dhdr . drdX), the
// linear-tet element (jax/fem/element/_tetra.py:36-45), the one-point rule (jax/fem/quadrature/_tetra.py:12-15)
// and WarpPotentialFem.from_region (warp/fem/_base.py:93-111: dV *= Fraction, materials) -- starting from mesh
// arrays that already live in HBM:
//   1. bounding box of the rest positions, 63-bit Morton key of every tet centroid        (kernels)
//   2. stable radix sort of (key, cell) pairs                                             (cub::DeviceRadixSort)
//   3. connectivity permuted into packed order -> host                                    (kernel + one D2H copy)
//   4. tile cut + conflict-aware tile tables, all host threads                            (tiling.cpp)
//   5. the static planes [Dm^-1 (9), vol * Fraction, materials] computed from the rest positions in packed order,
//      written straight into the handle's device planes                                   (kernel)
// The host never sees dhdX / dV / materials; at 64 M tets the whole setup is seconds instead of the minutes the
// numpy path (apple_b200.fem.Region + apl_fem_create) needs.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.h"

namespace apl {
int fem_upload_tables(apl_fem* f);

namespace {

__device__ __forceinline__ unsigned long long spread21_dev(unsigned long long x) {
    x &= 0x1fffffULL;
    x = (x | (x << 32)) & 0x1f00000000ffffULL;
    x = (x | (x << 16)) & 0x1f0000ff0000ffULL;
    x = (x | (x << 8)) & 0x100f00f00f00f00fULL;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ULL;
    x = (x | (x << 2)) & 0x1249249249249249ULL;
    return x;
}

// per-block min / max of the three coordinates -> partial[6 * block .. ]
__global__ void bbox_kernel(long long n_points, const double* __restrict__ points, double* __restrict__ partial) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < n_points;
         v += (long long)gridDim.x * blockDim.x)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double x = points[3 * v + k];
            lo[k] = fmin(lo[k], x);
            hi[k] = fmax(hi[k], x);
        }
    __shared__ double red[6][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        if (lane == 0) {
            red[k][wid] = lo[k];
            red[3 + k][wid] = hi[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        const int k = threadIdx.x;
        double r = red[k][0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = k < 3 ? fmin(r, red[k][w]) : fmax(r, red[k][w]);
        partial[6 * blockIdx.x + k] = r;
    }
}

// Morton key of the tet centroid, the very arithmetic of tiling.cpp's morton_order (no FMA contraction, so
// that host and device quantise identically); also flags connectivity outside [0, n_points).
__global__ void morton_kernel(long long n_cells, long long n_points, const int4* __restrict__ cells,
                              const double* __restrict__ points, double lo0, double lo1, double lo2, double scale,
                              unsigned long long* __restrict__ keys, int* __restrict__ vals, int* __restrict__ bad) {
    const double lo[3] = {lo0, lo1, lo2};
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < n_cells;
         c += (long long)gridDim.x * blockDim.x) {
        const int4 q4 = cells[c];
        const int v[4] = {q4.x, q4.y, q4.z, q4.w};
        bool ok = true;
#pragma unroll
        for (int a = 0; a < 4; ++a) ok &= v[a] >= 0 && v[a] < n_points;
        unsigned long long code = 0;
        if (ok) {
            unsigned long long q[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double s = 0.0;
#pragma unroll
                for (int a = 0; a < 4; ++a) s = __dadd_rn(s, points[3ll * v[a] + k]);
                q[k] = (unsigned long long)__dmul_rn(__dsub_rn(__dmul_rn(0.25, s), lo[k]), scale);
            }
            code = spread21_dev(q[0]) | (spread21_dev(q[1]) << 1) | (spread21_dev(q[2]) << 2);
        } else {
            atomicExch(bad, (int)(c < 0x7fffffff ? c : 0x7ffffffe) + 1);
        }
        keys[c] = code;
        vals[c] = (int)c;
    }
}

__global__ void permute_cells_kernel(long long n_cells, const int4* __restrict__ cells, const int* __restrict__ order,
                                     int4* __restrict__ packed) {
    for (long long pos = blockIdx.x * (long long)blockDim.x + threadIdx.x; pos < n_cells;
         pos += (long long)gridDim.x * blockDim.x)
        packed[pos] = cells[order[pos]];
}

__global__ void iota_kernel(long long n, int* __restrict__ v) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        v[i] = (int)i;
}

// One thread per packed tet position: rest shape in fp64 from the four rest positions (the closed form of
// Region.compute_grad for the linear tetrahedron: the rows of (dXdr)^-1 are cross products of the edge vectors
// over the determinant; rows 1..3 of dhdX are those rows, row 0 minus their sum is never stored), then the
// record [D(9), vol * Fraction, mu, lambda, activation(6) | vol * Fraction2, mu2] is rounded to T and written
// as 16-byte plane vectors.
template <typename T, int NREC>
__global__ void pack_planes_kernel(long long n_cells, long long plane_stride, const int4* __restrict__ packed,
                                   const int* __restrict__ order, const double* __restrict__ points,
                                   const T* __restrict__ fraction, const T* __restrict__ mu, const T* __restrict__ la,
                                   const T* __restrict__ act, const T* __restrict__ fraction2, const T* __restrict__ mu2,
                                   int kind, uint4* __restrict__ planes, int* __restrict__ flags) {
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int NPL = (NREC + VEC - 1) / VEC;
    for (long long pos = blockIdx.x * (long long)blockDim.x + threadIdx.x; pos < n_cells;
         pos += (long long)gridDim.x * blockDim.x) {
        const int4 q4 = packed[pos];
        const long long c = order[pos];
        const int v[4] = {q4.x, q4.y, q4.z, q4.w};
        double X[4][3];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int k = 0; k < 3; ++k) X[a][k] = points[3ll * v[a] + k];
        double e1[3], e2[3], e3[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            e1[k] = X[1][k] - X[0][k];
            e2[k] = X[2][k] - X[0][k];
            e3[k] = X[3][k] - X[0][k];
        }
        const double c23[3] = {e2[1] * e3[2] - e2[2] * e3[1], e2[2] * e3[0] - e2[0] * e3[2], e2[0] * e3[1] - e2[1] * e3[0]};
        const double c31[3] = {e3[1] * e1[2] - e3[2] * e1[1], e3[2] * e1[0] - e3[0] * e1[2], e3[0] * e1[1] - e3[1] * e1[0]};
        const double c12[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        const double det = e1[0] * c23[0] + e1[1] * c23[1] + e1[2] * c23[2];
        if (det == 0.0) atomicOr(flags, 1);        // degenerate tetrahedron: an error, as in Region.compute_grad
        else if (det < 0.0) atomicOr(flags, 2);    // dV <= 0 only warns (jax/fem/region/_region.py:98-99)
        const double inv = 1.0 / det;
        union {
            uint4 q[NPL];
            T s[NPL * VEC];
        } rec;
#pragma unroll
        for (int k = 0; k < NPL * VEC; ++k) rec.s[k] = (T)0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            rec.s[k] = (T)(c23[k] * inv);
            rec.s[3 + k] = (T)(c31[k] * inv);
            rec.s[6 + k] = (T)(c12[k] * inv);
        }
        const double dV = det * (1.0 / 6.0);
        // the product is formed like the host path forms it: Fraction * dV in fp64, then rounded to T
        rec.s[9] = (T)(fraction ? (double)fraction[c] * dV : dV);
        rec.s[10] = mu[c];
        if (la) rec.s[11] = la[c];
        if constexpr (NREC == 18) {
#pragma unroll
            for (int k = 0; k < 6; ++k) rec.s[12 + k] = act[6 * c + k];
        }
        if constexpr (NREC == 14) {
            rec.s[12] = (T)(fraction2 ? (double)fraction2[c] * dV : dV);
            rec.s[13] = mu2[c];
        }
#pragma unroll
        for (int k = 0; k < NPL; ++k) planes[(long long)k * plane_stride + pos] = rec.q[k];
    }
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
    template <typename U> U* as() { return reinterpret_cast<U*>(p); }
};
struct PinnedBuf {
    void* p = nullptr;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
};

int grid_of(long long n, int block, int sms) {
    long long g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > (long long)sms * 16) g = (long long)sms * 16;
    return (int)g;
}

template <typename T>
void launch_pack(apl_fem* f, long long n_cells, const int4* packed, const int* order, const double* points,
                 const void* fraction, const void* mu, const void* la, const void* act, const void* fraction2,
                 const void* mu2, int* flags, int grid) {
#define APL_PACK(NREC)                                                                                          \
    pack_planes_kernel<T, NREC><<<grid, 256>>>(n_cells, f->plane_stride, packed, order, points, (const T*)fraction, \
                                               (const T*)mu, (const T*)la, (const T*)act, (const T*)fraction2,    \
                                               (const T*)mu2, f->kind, (uint4*)f->d_planes, flags)
    if (f->nrec == 18) APL_PACK(18);
    else if (f->nrec == 14) APL_PACK(14);
    else APL_PACK(12);
#undef APL_PACK
}

}  // namespace
}  // namespace apl

using namespace apl;

extern "C" int apl_fem_create_from_mesh(int kind, int dtype, int64_t n_cells, int64_t n_points, const int32_t* cells,
                                        const double* points, const void* fraction, const void* mu, const void* lambda_,
                                        const void* activation, const void* fraction2, const void* mu2, int morton,
                                        int device, apl_fem_t** out) {
    if (!out) { set_error("apl_fem_create_from_mesh: out is NULL"); return APL_ERR_INVALID; }
    *out = nullptr;
    if (kind < 0 || kind > 3 || (dtype != APL_F32 && dtype != APL_F64)) {
        set_error("apl_fem_create_from_mesh: unknown kind or dtype");
        return APL_ERR_INVALID;
    }
    if (device < 0) { set_error("apl_fem_create_from_mesh: needs a CUDA device"); return APL_ERR_INVALID; }
    if (n_cells < 0 || n_points <= 0 || n_cells > (int64_t)INT32_MAX / 2 || n_points > (int64_t)INT32_MAX / 4) {
        set_error("apl_fem_create_from_mesh: sizes outside the int32 range of the packed tables");
        return APL_ERR_INVALID;
    }
    if ((n_cells > 0 && !cells) || !points || !mu || (kind != APL_KIND_ARAP && !lambda_) ||
        (kind == APL_KIND_SNH_MUSCLE && !activation) || (kind == APL_KIND_SNH_ARAP && !mu2)) {
        set_error("apl_fem_create_from_mesh: a required array (cells, points, mu, lambda_, activation, mu2) is NULL");
        return APL_ERR_INVALID;
    }
    APL_CUDA_CHECK(cudaSetDevice(device));
    int sms = 0;
    APL_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));

    // 1. bounding box -> Morton keys
    DevBuf d_keys, d_keys2, d_vals, d_order, d_packed, d_tmp, d_flags, d_part;
    APL_CUDA_CHECK(d_flags.alloc(2 * sizeof(int)));
    APL_CUDA_CHECK(cudaMemset(d_flags.p, 0, 2 * sizeof(int)));
    APL_CUDA_CHECK(d_order.alloc((size_t)(n_cells + 1) * sizeof(int)));
    APL_CUDA_CHECK(d_packed.alloc((size_t)n_cells * sizeof(int4)));
    const int gc = grid_of(n_cells, 256, sms);
    if (morton) {
        const int gp = std::min(grid_of(n_points, 256, sms), 1024);
        APL_CUDA_CHECK(d_part.alloc((size_t)gp * 6 * sizeof(double)));
        bbox_kernel<<<gp, 256>>>(n_points, points, d_part.as<double>());
        std::vector<double> part((size_t)gp * 6);
        APL_CUDA_CHECK(cudaMemcpy(part.data(), d_part.p, part.size() * sizeof(double), cudaMemcpyDeviceToHost));
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int b = 0; b < gp; ++b)
            for (int k = 0; k < 3; ++k) {
                lo[k] = std::min(lo[k], part[(size_t)6 * b + k]);
                hi[k] = std::max(hi[k], part[(size_t)6 * b + 3 + k]);
            }
        double ext = 0;
        for (int k = 0; k < 3; ++k) ext = std::max(ext, hi[k] - lo[k]);
        const double scale = ext > 0 ? (double)((1 << 21) - 1) / ext : 0.0;
        APL_CUDA_CHECK(d_keys.alloc((size_t)n_cells * 8));
        APL_CUDA_CHECK(d_keys2.alloc((size_t)n_cells * 8));
        APL_CUDA_CHECK(d_vals.alloc((size_t)n_cells * sizeof(int)));
        morton_kernel<<<gc, 256>>>(n_cells, n_points, (const int4*)cells, points, lo[0], lo[1], lo[2], scale,
                                   d_keys.as<unsigned long long>(), d_vals.as<int>(), d_flags.as<int>() + 1);
        // 2. stable sort by key (radix sort keeps equal keys in cell order, like the host's stable_sort)
        size_t tmp_bytes = 0;
        APL_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys.as<unsigned long long>(),
                                                       d_keys2.as<unsigned long long>(), d_vals.as<int>(),
                                                       d_order.as<int>(), (int)n_cells, 0, 63));
        APL_CUDA_CHECK(d_tmp.alloc(tmp_bytes));
        APL_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, d_keys.as<unsigned long long>(),
                                                       d_keys2.as<unsigned long long>(), d_vals.as<int>(),
                                                       d_order.as<int>(), (int)n_cells, 0, 63));
    } else {
        // keep the caller's order; the range check of the connectivity still runs
        APL_CUDA_CHECK(d_keys.alloc((size_t)n_cells * 8));
        morton_kernel<<<gc, 256>>>(n_cells, n_points, (const int4*)cells, points, 0.0, 0.0, 0.0, 0.0,
                                   d_keys.as<unsigned long long>(), d_order.as<int>(), d_flags.as<int>() + 1);
    }
    // 3. connectivity in packed order -> host
    permute_cells_kernel<<<gc, 256>>>(n_cells, (const int4*)cells, d_order.as<int>(), d_packed.as<int4>());
    APL_CUDA_CHECK(cudaGetLastError());
    int flags[2] = {0, 0};
    APL_CUDA_CHECK(cudaMemcpy(flags, d_flags.p, sizeof(flags), cudaMemcpyDeviceToHost));
    if (flags[1]) {
        set_error("cells[" + std::to_string(flags[1] - 1) + "] references a vertex outside [0, n_points)");
        return APL_ERR_MESH;
    }
    // the sort scratch is no longer needed: release it before the host tables are uploaded
    { DevBuf a, b, c, d; std::swap(a.p, d_keys.p); std::swap(b.p, d_keys2.p); std::swap(c.p, d_vals.p); std::swap(d.p, d_tmp.p); }
    std::vector<int32_t> h_packed((size_t)n_cells * 4);
    std::vector<int64_t> h_order((size_t)n_cells);
    {
        std::vector<int32_t> o32((size_t)n_cells);
        if (n_cells) {
            APL_CUDA_CHECK(cudaMemcpy(h_packed.data(), d_packed.p, h_packed.size() * 4, cudaMemcpyDeviceToHost));
            APL_CUDA_CHECK(cudaMemcpy(o32.data(), d_order.p, o32.size() * 4, cudaMemcpyDeviceToHost));
        }
        for (int64_t i = 0; i < n_cells; ++i) h_order[(size_t)i] = o32[(size_t)i];
    }

    // 4. tiles and tile tables on the host
    apl_fem* f = new apl_fem();
    f->kind = kind;
    f->dtype = dtype;
    f->device = device;
    f->nrec = kind == APL_KIND_SNH_MUSCLE ? 18 : (kind == APL_KIND_SNH_ARAP ? 14 : 12);
    const int vec = dtype == APL_F32 ? 4 : 2;
    f->nplanes = (f->nrec + vec - 1) / vec;
    int rc = build_tiles_packed(n_cells, n_points, h_packed.data(), std::move(h_order), dtype == APL_F32 ? 4 : 8, f->host);
    if (rc != APL_OK) { delete f; return rc; }
    std::vector<int32_t>().swap(h_packed);
    f->plane_stride = (f->host.n_packed() + 31) / 32 * 32;
    if (f->plane_stride == 0) f->plane_stride = 32;
    rc = fem_upload_tables(f);
    if (rc != APL_OK) { apl_fem_destroy(f); return rc; }

    // 5. static planes from the rest positions
    cudaError_t e = cudaMemset(f->d_planes, 0, (size_t)f->nplanes * f->plane_stride * 16);
    if (e == cudaSuccess && n_cells > 0) {
        if (dtype == APL_F32)
            launch_pack<float>(f, n_cells, d_packed.as<int4>(), d_order.as<int>(), points, fraction, mu, lambda_,
                               activation, fraction2, mu2, d_flags.as<int>(), gc);
        else
            launch_pack<double>(f, n_cells, d_packed.as<int4>(), d_order.as<int>(), points, fraction, mu, lambda_,
                                activation, fraction2, mu2, d_flags.as<int>(), gc);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpy(flags, d_flags.p, sizeof(int), cudaMemcpyDeviceToHost);
    }
    if (e != cudaSuccess) {
        set_error(std::string("apl_fem_create_from_mesh: ") + cudaGetErrorString(e));
        apl_fem_destroy(f);
        return APL_ERR_CUDA;
    }
    if (flags[0] & 1) {
        set_error("degenerate tetrahedron (zero rest volume)");
        apl_fem_destroy(f);
        return APL_ERR_MESH;
    }
    // packed position -> caller's cell stays on the device (mixed derivative products, material updates)
    f->d_order = d_order.as<int32_t>();
    d_order.p = nullptr;
    *out = f;
    return APL_OK;
}
