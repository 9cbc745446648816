// Vector-kernel helpers shared by the fused optimizers (pncg.cu, pcg.cu): 16-byte row access of (n, 4) nodal
// arrays, the deterministic grid-wide sum and the DOF mask bits.
#pragma once

#include <cuda_runtime.h>

#include "fem_kernels.cuh"

namespace apl {

constexpr int kVecThreads = 256;

template <typename T>
__device__ __forceinline__ void ld4(const T* __restrict__ base, long long row, T v[4]) {
    if constexpr (sizeof(T) == 4) {
        const float4 q = reinterpret_cast<const float4*>(base)[row];
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
        const double2 a = reinterpret_cast<const double2*>(base)[2 * row];
        const double2 b = reinterpret_cast<const double2*>(base)[2 * row + 1];
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
}

template <typename T>
__device__ __forceinline__ void st4(T* __restrict__ base, long long row, const T v[4]) {
    if constexpr (sizeof(T) == 4) {
        reinterpret_cast<float4*>(base)[row] = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        reinterpret_cast<double2*>(base)[2 * row] = make_double2(v[0], v[1]);
        reinterpret_cast<double2*>(base)[2 * row + 1] = make_double2(v[2], v[3]);
    }
}

// Grid-wide deterministic sum of NS doubles per thread -> out[0..NS) (overwritten by the last CTA).
template <int NS>
__device__ __forceinline__ void grid_reduce(double (&v)[NS], double* partials, unsigned int* counter, double* out) {
    constexpr int NW = kVecThreads / 32;
    __shared__ double red[NS][NW];
    __shared__ bool is_last;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const double w = warp_sum(v[s]);
        if (lane == 0) red[s][wid] = w;
    }
    __syncthreads();
    if (tid < NS) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += red[tid][w];
        partials[(size_t)blockIdx.x * NS + tid] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int done = atomicAdd(counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int s = 0; s < NS; ++s) {
        double acc = 0;
        for (int b = tid; b < (int)gridDim.x; b += kVecThreads) acc += __ldcg(partials + (size_t)b * NS + s);
        acc = warp_sum(acc);
        __syncthreads();
        if (lane == 0) red[0][wid] = acc;
        __syncthreads();
        if (tid == 0) {
            double t = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) t += red[0][w];
            out[s] = t;
        }
    }
    if (tid == 0) *counter = 0u;
}

// mask bits: 1 = free DOF (updated), 2 = counted in reductions (owned by this rank)
#define APL_M_FREE 1
#define APL_M_COUNT 2


}  // namespace apl
