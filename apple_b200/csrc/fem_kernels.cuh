// Element kernels for sm_100a: one pass over the tets evaluates any subset of
// {fun, grad, hess_diag, hess_prod, hess_quad} (the five Warp kernels of warp/fem/_base.py:243-383).
//
// TILE variant (the product path)
//   persistent CTAs, one 256-thread CTA per tile at a time:
//   1. coalesced 16-byte loads of the tile's static planes (Dm^-1, volume, materials), byte-wide
//      local connectivity and reduction slots                                   -> registers
//   2. each distinct tile vertex is gathered ONCE from global memory           -> shared memory
//   3. one thread per tet computes F, P, H-terms from shared memory and writes each corner's
//      contribution to its private shared-memory slot (no atomics)
//   4. one thread per tile vertex sums its contiguous slot range and issues ONE vector RED per
//      field to global memory (REDG.ADD.F32x4 / F32x2, or F64)
//   Energies / quadratic forms are reduced warp -> CTA -> per-CTA partial, and the last CTA to finish
//   adds the partials in a fixed order (deterministic for a fixed grid).
//
// ATOMIC variant (measurement baseline, same arithmetic)
//   one thread per tet, direct global gathers and 12 scalar REDs per field per tet, i.e. the
//   reference's strategy (wp.atomic_add at warp/fem/_base.py:288-289).
#pragma once

#include <cuda_runtime.h>

#include "common.h"
#include "elem_math.cuh"

namespace apl {

template <typename T>
struct FemArgs {
    const int4* tiles;
    int n_tiles;
    const uchar4* conn;
    const ushort4* slots;
    const int* tile_verts;
    const unsigned short* tile_voff;
    const uint4* planes;
    long long plane_stride;
    const T* u;
    const T* p;
    int ld_in;
    T* grad;
    T* diag;
    T* prod;
    int ld_out;
    T* fun;
    T* quad;
    double* partials;
    unsigned int* counter;
    // --- fused PNCG path (all optional) ---
    const T* axpy_p = nullptr;       // if set, the field evaluated is u + scal[alpha_idx] * axpy_p
    const double* scal = nullptr;    // device scalars of the PNCG workspace
    int alpha_idx = 0;
    int skip_a = -1, skip_b = -1;    // the launch is a no-op if scal[skip_a] != 0 or scal[skip_b] != 0
    double* fun_d = nullptr;         // double-precision sinks for the energy / quadratic form
    double* quad_d = nullptr;
};

template <typename T>
__device__ __forceinline__ bool fem_skip(const FemArgs<T>& a) {
    if (a.scal == nullptr) return false;
    if (a.skip_a >= 0 && __ldcg(a.scal + a.skip_a) != 0.0) return true;
    if (a.skip_b >= 0 && __ldcg(a.scal + a.skip_b) != 0.0) return true;
    return false;
}

// ---- small device helpers --------------------------------------------------------------------

template <typename T>
__device__ __forceinline__ void load_row(const T* __restrict__ base, int v, int ld, T* dst4);

template <>
__device__ __forceinline__ void load_row<float>(const float* __restrict__ base, int v, int ld, float* dst4) {
    if (ld == 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(base) + v);
        *reinterpret_cast<float4*>(dst4) = q;
    } else {
        const float* r = base + 3ll * v;
        const float x = __ldg(r), y = __ldg(r + 1), z = __ldg(r + 2);
        *reinterpret_cast<float4*>(dst4) = make_float4(x, y, z, 0.f);
    }
}

template <>
__device__ __forceinline__ void load_row<double>(const double* __restrict__ base, int v, int ld, double* dst4) {
    if (ld == 4) {
        const double2 a = __ldg(reinterpret_cast<const double2*>(base) + 2ll * v);
        const double2 b = __ldg(reinterpret_cast<const double2*>(base) + 2ll * v + 1);
        *reinterpret_cast<double2*>(dst4) = a;
        *reinterpret_cast<double2*>(dst4 + 2) = b;
    } else {
        const double* r = base + 3ll * v;
        const double x = __ldg(r), y = __ldg(r + 1), z = __ldg(r + 2);
        *reinterpret_cast<double2*>(dst4) = make_double2(x, y);
        *reinterpret_cast<double2*>(dst4 + 2) = make_double2(z, 0.0);
    }
}

// out[v, 0:3] += val[0:3]  -- one vector RED where the layout allows it
__device__ __forceinline__ void red_row(float* base, int v, int ld, const float* val) {
    if (ld == 4) {
        float* q = base + 4ll * v;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q), "f"(val[0]), "f"(val[1]),
                     "f"(val[2]), "f"(0.f)
                     : "memory");
    } else {
        float* q = base + 3ll * v;
        if ((v & 1) == 0) {  // 3v even -> (x,y) is 8-byte aligned
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(q), "f"(val[0]), "f"(val[1]) : "memory");
            atomicAdd(q + 2, val[2]);
        } else {  // 3v+1 even -> (y,z) is 8-byte aligned
            atomicAdd(q, val[0]);
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(q + 1), "f"(val[1]), "f"(val[2]) : "memory");
        }
    }
}

__device__ __forceinline__ void red_row(double* base, int v, int ld, const double* val) {
    double* q = base + (long long)ld * v;
    atomicAdd(q, val[0]);
    atomicAdd(q + 1, val[1]);
    atomicAdd(q + 2, val[2]);
}

template <typename T, int NREC>
struct Rec {
    static constexpr int VEC = 16 / (int)sizeof(T);
    static constexpr int NPL = (NREC + VEC - 1) / VEC;
    union {
        uint4 q[NPL];
        T s[NPL * VEC];
    };
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// CTA-wide sum of (e, q) -> partials[2*bid..]; the last CTA adds all partials in index order and
// accumulates into *fun / *quad.  Deterministic for a fixed grid size.
template <typename T, int NT>
__device__ __forceinline__ void finish_scalars(double e, double q, double* partials, unsigned int* counter,
                                               T* fun, T* quad, double* fun_d, double* quad_d) {
    __shared__ double red[2 * (NT / 32)];
    __shared__ bool is_last;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    e = warp_sum(e);
    q = warp_sum(q);
    if (lane == 0) {
        red[2 * wid] = e;
        red[2 * wid + 1] = q;
    }
    __syncthreads();
    if (tid == 0) {
        double se = 0, sq = 0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) {
            se += red[2 * w];
            sq += red[2 * w + 1];
        }
        partials[2 * blockIdx.x] = se;
        partials[2 * blockIdx.x + 1] = sq;
        __threadfence();
        const unsigned int done = atomicAdd(counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double se = 0, sq = 0;
        for (int b = tid; b < (int)gridDim.x; b += NT) {
            se += __ldcg(partials + 2 * b);
            sq += __ldcg(partials + 2 * b + 1);
        }
        se = warp_sum(se);
        sq = warp_sum(sq);
        __syncthreads();
        if (lane == 0) {
            red[2 * wid] = se;
            red[2 * wid + 1] = sq;
        }
        __syncthreads();
        if (tid == 0) {
            double te = 0, tq = 0;
#pragma unroll
            for (int w = 0; w < NT / 32; ++w) {
                te += red[2 * w];
                tq += red[2 * w + 1];
            }
            if (fun) *fun += (T)te;
            if (quad) *quad += (T)tq;
            if (fun_d) *fun_d += te;
            if (quad_d) *quad_d += tq;
            *counter = 0u;
        }
    }
}

// ---- compile-time layout of one instantiation ---------------------------------------------------

template <typename T, int OPS>
struct TileCfg {
    static constexpr bool kFun = (OPS & APL_OP_FUN) != 0;
    static constexpr bool kGrad = (OPS & APL_OP_GRAD) != 0;
    static constexpr bool kDiag = (OPS & APL_OP_HESS_DIAG) != 0;
    static constexpr bool kProd = (OPS & APL_OP_HESS_PROD) != 0;
    static constexpr bool kQuad = (OPS & APL_OP_HESS_QUAD) != 0;
    static constexpr bool kNeedP = kProd || kQuad;
    static constexpr int NOUT = (kGrad ? 1 : 0) + (kDiag ? 1 : 0) + (kProd ? 1 : 0);
    // slot stride in scalars: 3*NOUT rounded up so that a slot is a whole number of 16-byte vectors
    static constexpr int SS = (NOUT == 0) ? 0
                              : (NOUT == 1) ? 4
                              : (sizeof(T) == 4) ? (NOUT == 2 ? 8 : 12) : (NOUT == 2 ? 6 : 10);
    static constexpr int kNSlots = 4 * kTileTets;
    static constexpr size_t kUsBytes = (size_t)4 * kTileVerts * sizeof(T);
    static constexpr size_t kPsBytes = kNeedP ? kUsBytes : 0;
    static constexpr size_t kSlotBytes = (size_t)kNSlots * SS * sizeof(T);
    static constexpr size_t kVoffBytes = NOUT ? (size_t)((kTileVerts + 1) * 2 + 14) / 16 * 16 : 0;
    static constexpr size_t kSmemBytes = kUsBytes + kPsBytes + kSlotBytes + kVoffBytes;
};

template <typename T, int SS>
__device__ __forceinline__ void store_slot(T* dst, const T* v) {
    if constexpr (sizeof(T) == 4) {
        if constexpr (SS % 4 == 0) {
#pragma unroll
            for (int k = 0; k < SS / 4; ++k)
                reinterpret_cast<float4*>(dst)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        } else {
#pragma unroll
            for (int k = 0; k < SS / 2; ++k) reinterpret_cast<float2*>(dst)[k] = make_float2(v[2 * k], v[2 * k + 1]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < SS / 2; ++k) reinterpret_cast<double2*>(dst)[k] = make_double2(v[2 * k], v[2 * k + 1]);
    }
}

template <typename T, int SS>
__device__ __forceinline__ void load_slot(const T* src, T* v) {
    if constexpr (sizeof(T) == 4) {
        if constexpr (SS % 4 == 0) {
#pragma unroll
            for (int k = 0; k < SS / 4; ++k) {
                const float4 q = reinterpret_cast<const float4*>(src)[k];
                v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < SS / 2; ++k) {
                const float2 q = reinterpret_cast<const float2*>(src)[k];
                v[2 * k] = q.x; v[2 * k + 1] = q.y;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < SS / 2; ++k) {
            const double2 q = reinterpret_cast<const double2*>(src)[k];
            v[2 * k] = q.x; v[2 * k + 1] = q.y;
        }
    }
}

// ---- TILE kernel ----------------------------------------------------------------------------------

template <typename T, int KIND, int OPS>
__global__ void __launch_bounds__(kTileTets, (sizeof(T) == 4 ? 2 : 1)) fem_tile_kernel(const FemArgs<T> a) {
    using Cfg = TileCfg<T, OPS>;
    constexpr int NOUT = Cfg::NOUT;
    constexpr int SS = Cfg::SS;
    constexpr int NREC = RecSize<KIND>::value;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* us = reinterpret_cast<T*>(smem_raw);
    T* ps = reinterpret_cast<T*>(smem_raw + Cfg::kUsBytes);
    T* sl = reinterpret_cast<T*>(smem_raw + Cfg::kUsBytes + Cfg::kPsBytes);
    unsigned short* voff =
        reinterpret_cast<unsigned short*>(smem_raw + Cfg::kUsBytes + Cfg::kPsBytes + Cfg::kSlotBytes);

    const int tid = threadIdx.x;
    double e_acc = 0.0, q_acc = 0.0;
    if (fem_skip(a)) return;
    const bool axpy = a.axpy_p != nullptr;
    const T alpha = axpy ? (T)__ldcg(a.scal + a.alpha_idx) : (T)0;

    // Software prefetch of the INDICES one tile ahead (tile header, then this thread's vertex id), so
    // that the per-tile critical path is one memory round trip (static planes + vertex gather issued
    // together) instead of three dependent ones.
    int tile = blockIdx.x;
    int4 h = tile < a.n_tiles ? __ldg(a.tiles + tile) : make_int4(0, 0, 0, 0);
    int gv = (tid < (h.y >> 16)) ? __ldg(a.tile_verts + h.z + tid) : 0;
    for (; tile < a.n_tiles; tile += gridDim.x) {
        // header: tet_start, n_tets | n_verts << 16, vert_start, voff_start
        const int n_tets = h.y & 0xffff, n_verts = h.y >> 16;
        const int next = tile + gridDim.x;
        const int4 hn = next < a.n_tiles ? __ldg(a.tiles + next) : make_int4(0, 0, 0, 0);
        const bool active = tid < n_tets;
        const long long t = (long long)h.x + tid;
        uchar4 lc = make_uchar4(0, 0, 0, 0);
        ushort4 s4 = make_ushort4(0, 0, 0, 0);
        Rec<T, NREC> rec;
        if (active) {
            lc = __ldg(a.conn + t);
            if constexpr (NOUT > 0) s4 = __ldg(a.slots + t);
#pragma unroll
            for (int k = 0; k < Rec<T, NREC>::NPL; ++k) rec.q[k] = __ldg(a.planes + k * a.plane_stride + t);
        }
        if (tid < n_verts) {
            load_row<T>(a.u, gv, a.ld_in, us + 4 * tid);
            if constexpr (Cfg::kNeedP) load_row<T>(a.p, gv, a.ld_in, ps + 4 * tid);
            if (axpy) {  // trial point of the line search, never materialised in global memory
                T d[4];
                load_row<T>(a.axpy_p, gv, a.ld_in, d);
                us[4 * tid] += alpha * d[0];
                us[4 * tid + 1] += alpha * d[1];
                us[4 * tid + 2] += alpha * d[2];
            }
        }
        if constexpr (NOUT > 0) {
            for (int i = tid; i <= n_verts; i += kTileTets) voff[i] = __ldg(a.tile_voff + h.w + i);
        }
        __syncthreads();
        const int gvn = (tid < (hn.y >> 16)) ? __ldg(a.tile_verts + hn.z + tid) : 0;

        if (active) {
            T uc[4][3], pc[4][3];
            const int l[4] = {lc.x, lc.y, lc.z, lc.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                T tmp[4];
                load_slot<T, 4>(us + 4 * l[c], tmp);
                uc[c][0] = tmp[0]; uc[c][1] = tmp[1]; uc[c][2] = tmp[2];
                if constexpr (Cfg::kNeedP) {
                    load_slot<T, 4>(ps + 4 * l[c], tmp);
                    pc[c][0] = tmp[0]; pc[c][1] = tmp[1]; pc[c][2] = tmp[2];
                }
            }
            T psi = 0, quad = 0;
            T g[4][3], dg[4][3], hp[4][3];
            elem_eval<T, KIND, OPS>(rec.s, uc, pc, psi, quad, g, dg, hp);
            if constexpr (Cfg::kFun) e_acc += (double)psi;
            if constexpr (Cfg::kQuad) q_acc += (double)quad;
            if constexpr (NOUT > 0) {
                const int sidx[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    T v[SS];
                    int k = 0;
                    if constexpr (Cfg::kGrad) { v[k] = g[c][0]; v[k + 1] = g[c][1]; v[k + 2] = g[c][2]; k += 3; }
                    if constexpr (Cfg::kDiag) { v[k] = dg[c][0]; v[k + 1] = dg[c][1]; v[k + 2] = dg[c][2]; k += 3; }
                    if constexpr (Cfg::kProd) { v[k] = hp[c][0]; v[k + 1] = hp[c][1]; v[k + 2] = hp[c][2]; k += 3; }
#pragma unroll
                    for (int j = 3 * NOUT; j < SS; ++j) v[j] = (T)0;
                    store_slot<T, SS>(sl + sidx[c] * SS, v);
                }
            }
        }
        if constexpr (NOUT > 0) {
            __syncthreads();
            if (tid < n_verts) {
                const int s0 = voff[tid], s1 = voff[tid + 1];
                T acc[3 * NOUT];
#pragma unroll
                for (int j = 0; j < 3 * NOUT; ++j) acc[j] = (T)0;
                for (int s = s0; s < s1; ++s) {
                    T v[SS];
                    load_slot<T, SS>(sl + s * SS, v);
#pragma unroll
                    for (int j = 0; j < 3 * NOUT; ++j) acc[j] += v[j];
                }
                int k = 0;
                if constexpr (Cfg::kGrad) { if (a.grad) red_row(a.grad, gv, a.ld_out, acc + k); k += 3; }
                if constexpr (Cfg::kDiag) { if (a.diag) red_row(a.diag, gv, a.ld_out, acc + k); k += 3; }
                if constexpr (Cfg::kProd) { if (a.prod) red_row(a.prod, gv, a.ld_out, acc + k); k += 3; }
            }
        }
        __syncthreads();
        h = hn;
        gv = gvn;
    }
    if constexpr (Cfg::kFun || Cfg::kQuad)
        finish_scalars<T, kTileTets>(e_acc, q_acc, a.partials, a.counter, Cfg::kFun ? a.fun : nullptr,
                                     Cfg::kQuad ? a.quad : nullptr, Cfg::kFun ? a.fun_d : nullptr,
                                     Cfg::kQuad ? a.quad_d : nullptr);
}

// ---- ATOMIC kernel (baseline) -----------------------------------------------------------------------

template <typename T, int KIND, int OPS>
__global__ void __launch_bounds__(kTileTets) fem_atomic_kernel(const FemArgs<T> a) {
    using Cfg = TileCfg<T, OPS>;
    constexpr int NREC = RecSize<KIND>::value;
    const int tid = threadIdx.x;
    double e_acc = 0.0, q_acc = 0.0;
    if (fem_skip(a)) return;
    const bool axpy = a.axpy_p != nullptr;
    const T alpha = axpy ? (T)__ldcg(a.scal + a.alpha_idx) : (T)0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int4 h = __ldg(a.tiles + tile);
        if (tid < (h.y & 0xffff)) {
            const long long t = (long long)h.x + tid;
            const uchar4 lc = __ldg(a.conn + t);
            Rec<T, NREC> rec;
#pragma unroll
            for (int k = 0; k < Rec<T, NREC>::NPL; ++k) rec.q[k] = __ldg(a.planes + k * a.plane_stride + t);
            const int l[4] = {lc.x, lc.y, lc.z, lc.w};
            int gv[4];
            T uc[4][3], pc[4][3];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                gv[c] = __ldg(a.tile_verts + h.z + l[c]);
                const T* r = a.u + (long long)a.ld_in * gv[c];
                uc[c][0] = __ldg(r); uc[c][1] = __ldg(r + 1); uc[c][2] = __ldg(r + 2);
                if (axpy) {
                    const T* rd = a.axpy_p + (long long)a.ld_in * gv[c];
                    uc[c][0] += alpha * __ldg(rd); uc[c][1] += alpha * __ldg(rd + 1); uc[c][2] += alpha * __ldg(rd + 2);
                }
                if constexpr (Cfg::kNeedP) {
                    const T* rp = a.p + (long long)a.ld_in * gv[c];
                    pc[c][0] = __ldg(rp); pc[c][1] = __ldg(rp + 1); pc[c][2] = __ldg(rp + 2);
                }
            }
            T psi = 0, quad = 0;
            T g[4][3], dg[4][3], hp[4][3];
            elem_eval<T, KIND, OPS>(rec.s, uc, pc, psi, quad, g, dg, hp);
            if constexpr (Cfg::kFun) e_acc += (double)psi;
            if constexpr (Cfg::kQuad) q_acc += (double)quad;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const long long o = (long long)a.ld_out * gv[c];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    if constexpr (Cfg::kGrad) { if (a.grad) atomicAdd(a.grad + o + i, g[c][i]); }
                    if constexpr (Cfg::kDiag) { if (a.diag) atomicAdd(a.diag + o + i, dg[c][i]); }
                    if constexpr (Cfg::kProd) { if (a.prod) atomicAdd(a.prod + o + i, hp[c][i]); }
                }
            }
        }
    }
    if constexpr (Cfg::kFun || Cfg::kQuad)
        finish_scalars<T, kTileTets>(e_acc, q_acc, a.partials, a.counter, Cfg::kFun ? a.fun : nullptr,
                                     Cfg::kQuad ? a.quad : nullptr, Cfg::kFun ? a.fun_d : nullptr,
                                     Cfg::kQuad ? a.quad_d : nullptr);
}

// ---- launcher ---------------------------------------------------------------------------------------

// Declared per (T, KIND); defined in fem_inst.cu, one translation unit per pair.
template <typename T, int KIND>
int launch_fem(const apl_fem* fem, int ops, const FemArgs<T>& args, int scatter, cudaStream_t stream);

template <typename T, int KIND, int OPS>
int launch_one(const apl_fem* fem, const FemArgs<T>& args, int scatter, cudaStream_t stream) {
    using Cfg = TileCfg<T, OPS>;
    if (args.n_tiles == 0) return APL_OK;
    if (scatter == APL_SCATTER_TILE) {
        static int blocks_per_sm = -1;
        auto kern = fem_tile_kernel<T, KIND, OPS>;
        if (blocks_per_sm < 0) {
            APL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)Cfg::kSmemBytes));
            int b = 0;
            APL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, kTileTets, Cfg::kSmemBytes));
            blocks_per_sm = b > 0 ? b : 1;
        }
        int grid = fem->num_sms * blocks_per_sm;
        if (grid > args.n_tiles) grid = args.n_tiles;
        if (grid > fem->max_grid) grid = fem->max_grid;
        kern<<<grid, kTileTets, Cfg::kSmemBytes, stream>>>(args);
    } else {
        static int blocks_per_sm = -1;
        auto kern = fem_atomic_kernel<T, KIND, OPS>;
        if (blocks_per_sm < 0) {
            int b = 0;
            APL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, kTileTets, 0));
            blocks_per_sm = b > 0 ? b : 1;
        }
        int grid = fem->num_sms * blocks_per_sm;
        if (grid > args.n_tiles) grid = args.n_tiles;
        if (grid > fem->max_grid) grid = fem->max_grid;
        kern<<<grid, kTileTets, 0, stream>>>(args);
    }
    APL_CUDA_CHECK(cudaGetLastError());
    return APL_OK;
}

template <typename T, int KIND>
int launch_fem_impl(const apl_fem* fem, int ops, const FemArgs<T>& args, int scatter, cudaStream_t stream) {
    // hess_quad never shares a pass with vector outputs here: PNCG needs it alone (pass B)
    if ((ops & APL_OP_HESS_QUAD) && ops != APL_OP_HESS_QUAD) {
        int rc = launch_one<T, KIND, APL_OP_HESS_QUAD>(fem, args, scatter, stream);
        if (rc != APL_OK) return rc;
        ops &= ~APL_OP_HESS_QUAD;
    }
    // smallest instantiated superset; outputs that were not requested are NULL and skipped
    switch (ops) {
        case 0: return APL_OK;
        case 16: return launch_one<T, KIND, 16>(fem, args, scatter, stream);
        case 1: return launch_one<T, KIND, 1>(fem, args, scatter, stream);
        case 2: return launch_one<T, KIND, 2>(fem, args, scatter, stream);
        case 4: return launch_one<T, KIND, 4>(fem, args, scatter, stream);
        case 8: return launch_one<T, KIND, 8>(fem, args, scatter, stream);
        case 3: return launch_one<T, KIND, 3>(fem, args, scatter, stream);
        case 5: case 6: case 7: return launch_one<T, KIND, 7>(fem, args, scatter, stream);
        case 9: case 10: case 11: return launch_one<T, KIND, 11>(fem, args, scatter, stream);
        default: return launch_one<T, KIND, 15>(fem, args, scatter, stream);
    }
}

}  // namespace apl
