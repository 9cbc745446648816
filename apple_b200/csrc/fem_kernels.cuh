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

#include <cstdlib>
#include <mutex>

#include "common.h"
#include "elem_math.cuh"

#include "tile_logic.cuh"

namespace apl {

template <typename T>
struct FemArgs {
    const int4* tiles;
    int n_tiles;
    const uchar4* conn;
    const ushort4* slots;
    const int* tile_verts;
    const unsigned short* tile_voff;
    const unsigned char* tile_vperm;
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
    int dyn_j = 0;                   // add the device-side trial counter scal[APL_S_J] to alpha_idx, skip_b, fun_d
    double* fun_d = nullptr;         // double-precision sinks for the energy / quadratic form
    double* quad_d = nullptr;
};

template <typename T>
__device__ __forceinline__ int fem_joff(const FemArgs<T>& a) {
    return (a.scal != nullptr && a.dyn_j) ? (int)__ldcg(a.scal + APL_S_J) : 0;
}

template <typename T>
__device__ __forceinline__ bool fem_skip(const FemArgs<T>& a, int joff) {
    if (a.scal == nullptr) return false;
    if (a.skip_a >= 0 && __ldcg(a.scal + a.skip_a) != 0.0) return true;
    if (a.skip_b >= 0 && __ldcg(a.scal + a.skip_b + joff) != 0.0) return true;
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

// Scalar REDs as explicit PTX: `atomicAdd` with an unused result compiles to ATOMG ... PT, RZ here, which still
// allocates a WRITE scoreboard that clears only when L2 answers -- every consumer warp then stalled for the whole
// atomic round trip at the back edge of the tile loop (ncu r2d: 23 % of all stall samples on that branch).  `red` has
// no destination: only the operand-read scoreboard, released when the LSU has taken the request.
__device__ __forceinline__ void red_add(float* q, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(q), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add(double* q, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(q), "d"(v) : "memory");
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
            red_add(q + 2, val[2]);
        } else {  // 3v+1 even -> (y,z) is 8-byte aligned
            red_add(q, val[0]);
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(q + 1), "f"(val[1]), "f"(val[2]) : "memory");
        }
    }
}

__device__ __forceinline__ void red_row(double* base, int v, int ld, const double* val) {
    double* q = base + (long long)ld * v;
    red_add(q, val[0]);
    red_add(q + 1, val[1]);
    red_add(q + 2, val[2]);
}

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

// ---- the three per-tile phases shared by the simple and the pipelined kernel -----------------------

// (tile_compute / tile_compute_pair and the per-lane part of the reduction live in tile_logic.cuh, which is
// also compiled for the host by the test harness)

// Lanes l and l+16 of a warp sum the two halves of a vertex's slot range (tile_reduce_lane), exchange with
// one shuffle, and the lower lane parks the sums at the vertex's local id.
template <typename T, int OPS, int NT = kTileTets, int NSLOTS = kSlotsAlloc>
__device__ __forceinline__ void tile_reduce(int tid, int n_verts, const unsigned char* vperm,
                                            const unsigned short* voff, const T* sl, T* vbuf) {
    using Cfg = TileCfg<T, OPS>;
    constexpr int NOUT = Cfg::NOUT;
    const int half = (tid >> 4) & 1;
    for (int t = (tid >> 5) * 16 + (tid & 15); t < ((n_verts + 15) & ~15); t += NT / 2) {   // NT consumer threads
        // (the bound is rounded up to 16 so that whole warps stay together for the shuffle below;
        //  out-of-range lanes have cnt = 0)
        int v;
        T acc[3 * NOUT];
        tile_reduce_lane<T, OPS, NSLOTS>(half, t, n_verts, vperm, voff, sl, v, acc);
#pragma unroll
        for (int j = 0; j < 3 * NOUT; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
        if (half == 0 && t < n_verts) tile_reduce_park<T, OPS>(vbuf, v, acc);
    }
}

// Thread `tid` adds the sums of local vertex tid (ascending global id: neighbouring lanes hit
// neighbouring addresses) to global memory with one vector RED per field.
template <typename T, int OPS>
__device__ __forceinline__ void tile_flush(int tid, int n_verts, int gv, const T* vbuf, const FemArgs<T>& a) {
    using Cfg = TileCfg<T, OPS>;
    constexpr int NOUT = Cfg::NOUT;
    if (tid < n_verts) {
        T acc[3 * NOUT];
        tile_flush_read<T, OPS>(vbuf, tid, acc);
        int k = 0;
        if constexpr (Cfg::kGrad) { if (a.grad) red_row(a.grad, gv, a.ld_out, acc + k); k += 3; }
        if constexpr (Cfg::kDiag) { if (a.diag) red_row(a.diag, gv, a.ld_out, acc + k); k += 3; }
        if constexpr (Cfg::kProd) { if (a.prod) red_row(a.prod, gv, a.ld_out, acc + k); k += 3; }
    }
}

template <int N>
__device__ __forceinline__ void consumer_sync_n() { asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory"); }

// ---- TILE kernel, unpipelined (APL_SCATTER_TILE_SIMPLE) ---------------------------------------------

template <typename T, int KIND, int OPS>
__global__ void __launch_bounds__(kTileTets, (sizeof(T) == 4 ? 2 : 1)) fem_tile_kernel(const FemArgs<T> a) {
    using Cfg = TileCfg<T, OPS>;
    constexpr int NOUT = Cfg::NOUT;
    constexpr int NREC = RecSize<KIND>::value;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* vbuf = reinterpret_cast<T*>(smem_raw);
    T* us = vbuf;
    T* ps = vbuf + 4 * kTileVerts;
    T* sl = reinterpret_cast<T*>(smem_raw + Cfg::kVbufBytes);
    unsigned short* voff = reinterpret_cast<unsigned short*>(smem_raw + Cfg::kVbufBytes + Cfg::kSlotBytes);
    unsigned char* vperm = smem_raw + Cfg::kVbufBytes + Cfg::kSlotBytes + Cfg::kVoffBytes;

    const int tid = threadIdx.x;
    double e_acc = 0.0, q_acc = 0.0;
    const int joff = fem_joff(a);
    if (fem_skip(a, joff)) return;
    const bool axpy = a.axpy_p != nullptr;
    const T alpha = axpy ? (T)__ldcg(a.scal + a.alpha_idx + joff) : (T)0;

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
            else if (axpy) load_row<T>(a.axpy_p, gv, a.ld_in, ps + 4 * tid);
            if constexpr (NOUT > 0) vperm[tid] = __ldg(a.tile_vperm + h.z + tid);
        }
        if constexpr (NOUT > 0) {
            for (int i = tid; i <= n_verts; i += kTileTets) voff[i] = __ldg(a.tile_voff + h.w + i);
        }
        __syncthreads();
        const int gvn = (tid < (hn.y >> 16)) ? __ldg(a.tile_verts + hn.z + tid) : 0;

        if (active) tile_compute<T, KIND, OPS>(rec.s, lc, s4, us, ps, axpy, alpha, sl, e_acc, q_acc);
        if constexpr (NOUT > 0) {
            __syncthreads();
            tile_reduce<T, OPS>(tid, n_verts, vperm, voff, sl, vbuf);
            __syncthreads();
            tile_flush<T, OPS>(tid, n_verts, gv, vbuf, a);
        }
        __syncthreads();
        h = hn;
        gv = gvn;
    }
    if constexpr (Cfg::kFun || Cfg::kQuad)
        finish_scalars<T, kTileTets>(e_acc, q_acc, a.partials, a.counter, Cfg::kFun ? a.fun : nullptr,
                                     Cfg::kQuad ? a.quad : nullptr, (Cfg::kFun && a.fun_d) ? a.fun_d + joff : nullptr,
                                     Cfg::kQuad ? a.quad_d : nullptr);
}

// ---- PIPELINED kernel (APL_SCATTER_TILE, the product path) ------------------------------------------
//
// 9 warps per CTA, persistent, no CTA-wide barrier in the tile loop.
//
// Warp 8 is the PRODUCER: for every tile it issues TMA bulk copies (cp.async.bulk, completion on an mbarrier)
// of the tile's static planes, connectivity, slots and vertex tables into the next free shared-memory stage,
// waits for the vertex table, and gathers the tile's vertices into the same stage (16-byte rows: cp.async /
// LDGSTS; 12-byte rows: one 8-byte and one 4-byte load per row through registers, one 16-byte store).
//
// Warps 0-7 are CONSUMERS (one thread per tet) and run two phases per tile:
//   COMPUTE(i): wait for the stage, evaluate the tet, write each corner's contribution to its slot of slot
//               buffer i % NB, release the stage to the producer;
//   REDUCE(i):  two lanes per vertex sum the vertex's slot range of buffer i % NB and issue the vector REDs
//               to global memory straight from the reducing lanes.
// With NB = 2 slot buffers the phases are software-pipelined -- COMPUTE(i+1) runs BEFORE REDUCE(i) -- and
// hand-offs are mbarriers (slots_full[b]: every consumer arrived after its slot stores; slots_free[b]: every
// consumer arrived after its reduction), so a warp that finishes a phase early starts the next phase of the
// next tile instead of idling at a barrier; warps drift apart by up to one phase.  Instantiations whose two
// slot buffers would not fit the target CTAs per SM run the phases in order with NB = 1 (same mbarriers).

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
#ifndef APL_WAIT_SLEEP_NS
#define APL_WAIT_SLEEP_NS 0
#endif
#ifndef APL_WAIT_HINT_NS
#define APL_WAIT_HINT_NS 0   // > 0: suspend-time hint of mbarrier.try_wait (fewer spin instructions competing for issue slots)
#endif
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    do {
#if APL_WAIT_HINT_NS > 0
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"((unsigned)APL_WAIT_HINT_NS)
            : "memory");
#else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
#endif
#if APL_WAIT_SLEEP_NS > 0
        if (!ok) __nanosleep(APL_WAIT_SLEEP_NS);   // back off: a spinning warp competes with computing warps for issue slots
#endif
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// DRAM -> L2 only (no shared memory needed): lets the producer run the static stream several tiles ahead of the stages
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
template <int BYTES>
__device__ __forceinline__ void cp_async(unsigned dst, const void* src) {
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(unsigned bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

template <typename T>
__device__ __forceinline__ void gather_row_async(unsigned dst, const T* base, int v, int ld) {
    if (ld == 4) {
        const char* src = reinterpret_cast<const char*>(base) + (size_t)v * 4 * sizeof(T);
        cp_async<16>(dst, src);
        if constexpr (sizeof(T) == 8) cp_async<16>(dst + 16, src + 16);
    } else {
        const T* src = base + 3ll * v;
        cp_async<sizeof(T)>(dst, src);
        cp_async<sizeof(T)>(dst + sizeof(T), src + 1);
        cp_async<sizeof(T)>(dst + 2 * sizeof(T), src + 2);
    }
}

// Packed layout (TileCfg::kPackedIn): row v of the first array is [ux uy px py] (16 bytes), row v of the second array,
// kTileVerts * 16 bytes further, is [uz pz] (8 bytes).  dst = address of the first array.
__device__ __forceinline__ void gather_row_packed(unsigned dst, const float* u, const float* p, int v, int lv, int ld) {
    const unsigned a = dst + 16u * (unsigned)lv, b = dst + 16u * (unsigned)kTileVerts + 8u * (unsigned)lv;
    if (ld == 4) {
        const float* su = u + 4ll * v;
        const float* sp = p + 4ll * v;
        cp_async<8>(a, su);
        cp_async<8>(a + 8, sp);
        cp_async<4>(b, su + 2);
        cp_async<4>(b + 4, sp + 2);
    } else {
        const float* su = u + 3ll * v;
        const float* sp = p + 3ll * v;
        cp_async<4>(a, su);
        cp_async<4>(a + 4, su + 1);
        cp_async<4>(a + 8, sp);
        cp_async<4>(a + 12, sp + 1);
        cp_async<4>(b, su + 2);
        cp_async<4>(b + 4, sp + 2);
    }
}

// ld = 3 row through registers.  fp32: the 12-byte row of vertex gv starts at 12 gv, so either its first or its
// last 8 bytes are 8-byte aligned when the base pointer is (the caller checks); otherwise three scalar loads.
template <typename T>
__device__ __forceinline__ void load_row3_ldg(const T* __restrict__ base, int gv, bool base_aligned8, T* r) {
    if constexpr (sizeof(T) == 4) {
        const char* row = reinterpret_cast<const char*>(base) + 12ll * gv;
        if (base_aligned8) {
            const bool odd = gv & 1;
            const float2 pr = __ldg(reinterpret_cast<const float2*>(row + (odd ? 4 : 0)));
            const float one = __ldg(reinterpret_cast<const float*>(row + (odd ? 0 : 8)));
            r[0] = odd ? one : pr.x;
            r[1] = odd ? pr.x : pr.y;
            r[2] = odd ? pr.y : one;
        } else {
            const float* q = reinterpret_cast<const float*>(row);
            r[0] = __ldg(q); r[1] = __ldg(q + 1); r[2] = __ldg(q + 2);
        }
    } else {
        const T* row = base + 3ll * gv;
        r[0] = __ldg(row); r[1] = __ldg(row + 1); r[2] = __ldg(row + 2);
    }
}
template <typename T>
__device__ __forceinline__ void store_row4(T* dst, const T* r) {
    if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(dst) = make_float4(r[0], r[1], r[2], 0.f);
    } else {
        *reinterpret_cast<double2*>(dst) = make_double2(r[0], r[1]);
        *reinterpret_cast<double2*>(dst + 2) = make_double2(r[2], 0.0);
    }
}

constexpr int kPipeThreads = kTileTets + 32;

#ifndef APL_GATHER_LDG
#define APL_GATHER_LDG 0   // 1: 12-byte rows through registers (one 8-byte + one 4-byte load per row, issued before the
                           //    wait for the stage); 0: three 4-byte cp.async per row
#endif
#ifndef APL_SLOT_BUFS
#define APL_SLOT_BUFS 2    // slot buffers wanted (2: software-pipelined phases; 1: phases in order)
#endif
#ifndef APL_PRODUCER_WARPS
#define APL_PRODUCER_WARPS 1   // producer warps of the kernels that do not spill
#endif
#ifndef APL_PRODUCER_WARPS_FUSED
#define APL_PRODUCER_WARPS_FUSED 4   // fp32 SNH+ARAP kernels: a whole producer WARPGROUP (the gather is split four ways)
                                     // and register re-allocation with setmaxnreg: 32 registers for the producers, 104 for
                                     // the consumers (the sum must not exceed the launch allocation, 80 x 384).  A 9-warp
                                     // CTA caps every thread at 96 registers (five warps share one scheduler's file) and
                                     // the fused kernels spill there; measured +3.7 % (run r2n).  The other kernels do
                                     // not spill and lose 2-5 % to the extra warps.
#endif
#ifndef APL_L2_PREFETCH
#define APL_L2_PREFETCH 0   // 1: bulk L2 prefetch of tile it + 2's static stream; 2: also the vertex rows of tile it + 1.
                            // NEGATIVE result, kept as a knob: interleaved A/B on one box (run r2q, G tets/s, 0 / 1 / 2):
                            // SNH 8 M 37.0 / 35.6 / 31.8, SNH 64 M 39.2 / 37.6 / 33.2, SNH+ARAP 64 M 25.5 / 25.2 / 24.5, HVP
                            // alone 45.1 / 40.1 / 35.3 -- the bulk copies of the stages already keep enough bytes in flight.
#endif
#ifndef APL_PRODUCER_PARK
#define APL_PRODUCER_PARK 1
#endif
#ifndef APL_WANT_CTAS
#define APL_WANT_CTAS 2    // 3: fp32 kernels with a small enough tile state run three CTAs per SM (one slot buffer)
#endif

template <typename T, int KIND, int OPS>
struct PipeCfg {
    using Cfg = TileCfg<T, OPS>;
    static constexpr int kConsumers = kTileTets;   // one consumer thread per tet
    static constexpr int kProducerWarps =
        (KIND == APL_KIND_SNH_ARAP && sizeof(T) == 4 && kTileTets == 256) ? APL_PRODUCER_WARPS_FUSED : APL_PRODUCER_WARPS;
    static constexpr int kThreads = kConsumers + 32 * kProducerWarps;
    // register re-allocation between the producer warpgroup and the two consumer warpgroups (fp32, two CTAs per SM)
    static constexpr bool kRegRealloc = kProducerWarps == 4 && sizeof(T) == 4 && kTileTets == 256;
    static constexpr int kNSlots = kSlotsAlloc;
    static constexpr size_t kSlotBytes = (size_t)kNSlots * Cfg::SS * sizeof(T);
    static constexpr int NREC = RecSize<KIND>::value;
    static constexpr int NPL = Rec<T, NREC>::NPL;
    // one stage: per-tet static data + the gathered vertex fields u, p (all offsets multiples of 16 bytes)
    static constexpr size_t kVbufBytes = (size_t)kTileVerts * 8 * sizeof(T);
    static constexpr size_t oPlanes = 0;
    static constexpr size_t oConn = oPlanes + (size_t)NPL * kTileTets * 16;
    static constexpr size_t oSlots = oConn + (size_t)kTileTets * 4;
    static constexpr size_t oVbuf = oSlots + (size_t)kTileTets * 8;
    static constexpr size_t oHdr = oVbuf + kVbufBytes;
    static constexpr size_t kStageBytes = oHdr + 16;
    // one vertex-table slot (requested one tile ahead of its stage): global ids, reduce order, slot offsets
    static constexpr size_t oVerts = 0;
    static constexpr size_t oVperm = oVerts + (size_t)kTileVerts * 4;
    static constexpr size_t oVoff = oVperm + (size_t)kTileVerts;
    static constexpr size_t kVtabBytes = oVoff + Cfg::kVoffRaw;
    // mbarriers: full[S], empty[S], vfull[S + NB + 1], slots_full[NB], slots_free[NB]
    static constexpr size_t bar_bytes(int nb, int stages) { return ((size_t)8 * (3 * stages + 3 * nb + 1) + 15) / 16 * 16; }
    // Shared memory per CTA for N CTAs per SM (227 KB per SM, 1 KB reserved per CTA).  fp32 kernels whose tile state
    // is small enough run THREE CTAs per SM with one slot buffer (phases in order); the others TWO CTAs per SM (fp64:
    // one) with two slot buffers (software-pipelined phases) when those fit next to two stages.
    static constexpr size_t budget(int ctas) { return (size_t)227 * 1024 / ctas - 1024; }
    static constexpr size_t total(int nb, int stages) {
        return bar_bytes(nb, stages) + (size_t)nb * kSlotBytes + (size_t)(stages + nb + 1) * kVtabBytes +
               (size_t)stages * kStageBytes;
    }
#ifdef APL_SMEM_BUDGET_KB
    static constexpr bool kThree = false;
    static constexpr int kWantCtas = 2;
    static constexpr size_t kBudget = (size_t)APL_SMEM_BUDGET_KB * 1024;
#else
    static constexpr bool kThree = sizeof(T) == 4 && kTileTets == 256 && APL_WANT_CTAS >= 3 && total(1, 2) <= budget(3);
    static constexpr int kWantCtas = kThree ? 3 : (sizeof(T) == 4 ? 2 : 1) * (256 / kTileTets);
    static constexpr size_t kBudget = budget(kWantCtas);
#endif
    static constexpr int kSlotBufs =
        (Cfg::NOUT == 0 || kThree) ? 1 : ((APL_SLOT_BUFS >= 2 && total(2, 2) <= kBudget) ? 2 : 1);
    static constexpr int kStages = total(kSlotBufs, 4) <= kBudget ? 4 : (total(kSlotBufs, 3) <= kBudget ? 3 : 2);
    // a vertex table lives from its request (one tile ahead of the stage) until the tile's REDUCE phase has
    // finished in every consumer, which trails the stage release by up to kSlotBufs tiles
    static constexpr int kVring = kStages + kSlotBufs + 1;
    static constexpr size_t kBarBytes = bar_bytes(kSlotBufs, kStages);
    static constexpr size_t oSlotBuf = kBarBytes;
    static constexpr size_t oVring = oSlotBuf + (size_t)kSlotBufs * kSlotBytes;
    static constexpr size_t oStages = oVring + (size_t)kVring * kVtabBytes;
    static constexpr size_t kSmemBytes = oStages + (size_t)kStages * kStageBytes;
    // CTAs per SM the shared memory allows: the register budget of __launch_bounds__ follows it, so that a
    // kernel that is shared-memory-limited does not spill to fit one more CTA
    static constexpr int kSmemCtas = (int)((size_t)227 * 1024 / (kSmemBytes + 1024));
    static constexpr int kMinCtas = kSmemCtas < 1 ? 1 : (kSmemCtas < kWantCtas ? kSmemCtas : kWantCtas);
};

// REDUCE phase of one tile: lanes l and l+16 of a warp sum the two halves of a vertex's slot range
// (tile_reduce_lane), exchange with one shuffle, and issue the REDs from both lanes (field f from the lane of
// half f & 1).  Groups of 16 vertices (reduce order: about decreasing valence) are dealt to the warps round
// robin, rotated by `rot` from tile to tile so that no warp always owns the heaviest group.
// ALL slot reads of the thread (at most two groups: kTileVerts <= NT) come first, then `release` (the arrival on the
// slots-free barrier), then the REDs: an mbarrier arrival has release semantics and would otherwise wait for
// the thread's outstanding REDs to be performed -- a round trip to L2 on the critical path of every tile.
template <typename T, int OPS, int NT, int NSLOTS, bool PACKED_SLOTS, typename Release>
__device__ __forceinline__ void tile_reduce_flush(int tid, int rot, int n_verts, const unsigned char* vperm,
                                                  const unsigned short* voff, const int* verts, const T* sl,
                                                  const FemArgs<T>& a, Release&& release) {
    using Cfg = TileCfg<T, OPS>;
    constexpr int NOUT = Cfg::NOUT;
    static_assert(kTileVerts <= NT, "two reduce trips per thread cover a tile");
    const int half = (tid >> 4) & 1;
    const int w = ((tid >> 5) + rot) & (NT / 32 - 1);
    T* outs[3] = {nullptr, nullptr, nullptr};
    {
        int k = 0;
        if constexpr (Cfg::kGrad) outs[k++] = a.grad;
        if constexpr (Cfg::kDiag) outs[k++] = a.diag;
        if constexpr (Cfg::kProd) outs[k++] = a.prod;
    }
    const int bound = (n_verts + 15) & ~15;   // rounded up to 16 so that whole warps stay together for the shuffle
    T acc[2][3 * NOUT];
    int gv[2] = {-1, -1};
#pragma unroll
    for (int trip = 0; trip < 2; ++trip) {
        const int t = w * 16 + (tid & 15) + trip * (NT / 2);
        if (t < bound) {                      // warp-uniform
            int v;
            tile_reduce_lane<T, OPS, NSLOTS>(half, t, n_verts, vperm, voff, sl, v, acc[trip]);
#pragma unroll
            for (int j = 0; j < 3 * NOUT; ++j) acc[trip][j] += __shfl_xor_sync(0xffffffffu, acc[trip][j], 16);
            if (t < n_verts) gv[trip] = verts[v];
        }
    }
    release();
#pragma unroll
    for (int trip = 0; trip < 2; ++trip) {
        if (gv[trip] < 0) continue;
        const T* s = acc[trip];
        if constexpr (NOUT == 1) {
            if (half == 0 && outs[0]) red_row(outs[0], gv[trip], a.ld_out, s);
        } else {
            T* base = half ? outs[1] : outs[0];
            // slot order: [g(3), second field(3)], or packed [gx gy hx hy gz hz] (tile_compute_packed)
            constexpr bool PK = Cfg::kPacked && PACKED_SLOTS;
            const T val[3] = {half ? s[PK ? 2 : 3] : s[0], half ? s[PK ? 3 : 4] : s[PK ? 1 : 1],
                              half ? s[5] : s[PK ? 4 : 2]};
            if (base) red_row(base, gv[trip], a.ld_out, val);
            if constexpr (NOUT == 3) {
                if (half == 0 && outs[2]) red_row(outs[2], gv[trip], a.ld_out, s + 6);
            }
        }
    }
}

template <typename T, int KIND, int OPS>
__global__ void __launch_bounds__(PipeCfg<T, KIND, OPS>::kThreads, PipeCfg<T, KIND, OPS>::kMinCtas)
    fem_pipe_kernel(const FemArgs<T> a) {
    using Cfg = TileCfg<T, OPS>;
    using PC = PipeCfg<T, KIND, OPS>;
    constexpr int NOUT = Cfg::NOUT;
    constexpr int NC = PC::kConsumers;             // consumer threads (warps 0 .. NC/32 - 1), producer = warp NC/32
    constexpr int S = PC::kStages, SV = PC::kVring, NB = PC::kSlotBufs;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // mbarriers full[S], empty[S], vfull[SV], slots_full[NB], slots_free[NB];  NB slot buffers;
    // SV vertex-table slots;  S stages
    static_assert(8 * (2 * S + SV + 2 * NB) <= (int)PC::kBarBytes, "mbarrier area too small");
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw);
    unsigned char* vring = smem_raw + PC::oVring;
    unsigned char* stages = smem_raw + PC::oStages;
    const unsigned bar0 = smem_u32(bars);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (S + s); };
    auto vfull = [&](int s) { return bar0 + 8u * (2 * S + s); };
    auto sfull = [&](int b) { return bar0 + 8u * (2 * S + SV + b); };
    auto sfree = [&](int b) { return bar0 + 8u * (2 * S + SV + NB + b); };

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    double e_acc = 0.0, q_acc = 0.0;
    const int joff = fem_joff(a);
    if (fem_skip(a, joff)) return;
    const bool axpy = a.axpy_p != nullptr;
    const T alpha = axpy ? (T)__ldcg(a.scal + a.alpha_idx + joff) : (T)0;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full(s), 32 * PC::kProducerWarps + 1);   // every gather lane + the expect_tx arrival
            mbar_init(empty(s), NC);         // every consumer thread releases the stage
        }
        for (int s = 0; s < SV; ++s) mbar_init(vfull(s), 1);
        for (int b = 0; b < NB; ++b) {
            mbar_init(sfull(b), NC);
            mbar_init(sfree(b), NC);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int my_tiles = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (PC::kProducerWarps == 1 ? warp == NC / 32 : warp >= NC / 32) {
        // ================================= producer warp(s) ==============================
        if constexpr (PC::kRegRealloc) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        constexpr int PW = PC::kProducerWarps;
        const int pw = PW == 1 ? 0 : warp - NC / 32;      // producer warp index; warp 0 issues the bulk copies
        // Iteration `it`: wait until stage it % S is free, issue the bulk copies of tile `it`'s static
        // data, request the vertex tables of tile `it + 1`, then gather tile `it`'s vertices (its
        // tables were requested one iteration ago).  Headers are prefetched two iterations ahead, so
        // the producer never waits on a fresh global load in steady state.
        auto load_hdr = [&](int it) {
            int4 h = make_int4(0, 0, 0, 0);
            if (lane == 0 && it < my_tiles) h = __ldg(a.tiles + (blockIdx.x + it * gridDim.x));
            return h;
        };
        auto request_vtab = [&](int it, const int4& h) {  // lane 0 only
            const int sv = it % SV;
            const unsigned vt32 = smem_u32(vring + (size_t)sv * PC::kVtabBytes);
            const unsigned n_verts = (unsigned)(h.y >> 16);
            const unsigned bv = (n_verts * 4u + 15u) & ~15u;
            const unsigned bp = (n_verts + 15u) & ~15u, bo = ((n_verts + 1u) * 2u + 15u) & ~15u;
            mbar_expect_tx(vfull(sv), bv + (NOUT > 0 ? bp + bo : 0u));
            bulk_g2s(vt32 + (unsigned)PC::oVerts, a.tile_verts + h.z, bv, vfull(sv));
            if constexpr (NOUT > 0) {
                bulk_g2s(vt32 + (unsigned)PC::oVperm, a.tile_vperm + h.z, bp, vfull(sv));
                bulk_g2s(vt32 + (unsigned)PC::oVoff, a.tile_voff + h.w, bo, vfull(sv));
            }
        };
        const T* pf = Cfg::kNeedP ? a.p : a.axpy_p;     // second gathered field (nullptr: none)
        const bool want_p = Cfg::kNeedP || axpy;
        const bool aligned8 = (((size_t)a.u | (size_t)(want_p ? pf : a.u)) & 7u) == 0;
        int4 h_cur = load_hdr(0), h_nxt = load_hdr(1), h_nn = load_hdr(2);
        if (lane == 0 && pw == 0) request_vtab(0, h_cur);
        for (int it = 0; it < my_tiles; ++it) {
            const int s = it % S;
            const unsigned ph = (unsigned)(it / S) & 1u;
            unsigned char* st = stages + (size_t)s * PC::kStageBytes;
            const unsigned st32 = smem_u32(st);
            const int4 h_n3 = load_hdr(it + 3);   // three tiles ahead: consumed (L2 prefetch) one iteration later
            int4 h;
            h.x = __shfl_sync(0xffffffffu, h_cur.x, 0);
            h.y = __shfl_sync(0xffffffffu, h_cur.y, 0);
            h.z = __shfl_sync(0xffffffffu, h_cur.z, 0);
            h.w = __shfl_sync(0xffffffffu, h_cur.w, 0);
            const int n_tets = h.y & 0xffff, n_verts = h.y >> 16;
            // the tables of tile `it` were requested one iteration ago
            const int sv = it % SV;
            // With a producer WARPGROUP only warp 0 polls the mbarriers; the other three park at a hardware barrier (a
            // parked warp issues nothing, a polling one competes with the consumers for issue slots: the fused kernels'
            // four polling producer warps accounted for a third of all issued instructions, ncu r2z).
            constexpr bool kPark = PW > 1 && APL_PRODUCER_PARK;
            static_assert(!(kPark && APL_GATHER_LDG), "the register-staged gather reads the vertex table before the hand-shake");
            if (!kPark || pw == 0) mbar_wait(vfull(sv), (unsigned)(it / SV) & 1u);
            const int* verts = reinterpret_cast<const int*>(vring + (size_t)sv * PC::kVtabBytes + PC::oVerts);
            const bool via_regs = APL_GATHER_LDG && a.ld_in == 3 && !Cfg::kPackedIn;
            constexpr int R = APL_GATHER_LDG ? (kTileVerts / 32 + PW - 1) / PW : 1;
            T ru[R][3], rp[R][3];
            if (via_regs) {
                // 12-byte rows through registers: the loads are issued BEFORE the wait for the stage, so their
                // latency overlaps with the time the consumers still hold it
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    const int v = lane + 32 * (pw + PW * k);
                    if (v < n_verts) {
                        const int gv = verts[v];
                        load_row3_ldg<T>(a.u, gv, aligned8, ru[k]);
                        if (want_p) load_row3_ldg<T>(pf, gv, aligned8, rp[k]);
                    }
                }
            }
            // stage s free <=> every consumer is past COMPUTE(it - S); its REDUCE phases up to tile
            // it - S - NB are then finished too, which frees vertex slot (it + 1) % SV
            if (!kPark || pw == 0) mbar_wait(empty(s), ph ^ 1u);
            if constexpr (kPark) asm volatile("bar.sync 2, %0;" ::"n"(32 * PW) : "memory");
            if (lane == 0 && pw == 0) {
                *reinterpret_cast<int4*>(st + PC::oHdr) = h;
                const unsigned bp = (unsigned)n_tets * 16u;
                const unsigned bc = ((unsigned)n_tets * 4u + 15u) & ~15u;
                const unsigned bs = ((unsigned)n_tets * 8u + 15u) & ~15u;
                mbar_expect_tx(full(s), (unsigned)PC::NPL * bp + bc + (NOUT > 0 ? bs : 0u));
#pragma unroll
                for (int k = 0; k < PC::NPL; ++k)
                    bulk_g2s(st32 + (unsigned)(PC::oPlanes + (size_t)k * kTileTets * 16),
                             a.planes + k * a.plane_stride + h.x, bp, full(s));
                bulk_g2s(st32 + (unsigned)PC::oConn, a.conn + h.x, bc, full(s));
                if constexpr (NOUT > 0) bulk_g2s(st32 + (unsigned)PC::oSlots, a.slots + h.x, bs, full(s));
                if (it + 1 < my_tiles) request_vtab(it + 1, h_nxt);
#if APL_L2_PREFETCH
                // experiment (off by default, see APL_L2_PREFETCH): tile it + 2's static stream pulled into L2 two tile
                // periods before its bulk copies are issued
                if (it + 2 < my_tiles) {
                    const unsigned nt2 = (unsigned)(h_nn.y & 0xffff), nv2 = (unsigned)(h_nn.y >> 16);
                    const unsigned pb = nt2 * 16u;
#pragma unroll
                    for (int k = 0; k < PC::NPL; ++k) bulk_prefetch_l2(a.planes + k * a.plane_stride + h_nn.x, pb);
                    bulk_prefetch_l2(a.conn + h_nn.x, (nt2 * 4u + 15u) & ~15u);
                    if constexpr (NOUT > 0) bulk_prefetch_l2(a.slots + h_nn.x, (nt2 * 8u + 15u) & ~15u);
                    bulk_prefetch_l2(a.tile_verts + h_nn.z, (nv2 * 4u + 15u) & ~15u);
                }
#endif
            }
            const unsigned us32 = st32 + (unsigned)PC::oVbuf;
            const unsigned ps32 = us32 + 4u * kTileVerts * (unsigned)sizeof(T);
            if (via_regs) {
                T* vb = reinterpret_cast<T*>(st + PC::oVbuf);
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    const int v = lane + 32 * (pw + PW * k);
                    if (v < n_verts) {
                        store_row4<T>(vb + 4 * v, ru[k]);
                        if (want_p) store_row4<T>(vb + 4 * kTileVerts + 4 * v, rp[k]);
                    }
                }
                mbar_arrive(full(s));   // release: the stores above are visible to the consumers that acquire the phase
            } else {
                for (int v = lane + 32 * pw; v < n_verts; v += 32 * PW) {
                    const int gv = verts[v];
                    if constexpr (Cfg::kPackedIn) {
                        gather_row_packed(us32, (const float*)a.u, (const float*)pf, gv, v, a.ld_in);
                    } else {
                        gather_row_async<T>(us32 + (unsigned)v * 4u * (unsigned)sizeof(T), a.u, gv, a.ld_in);
                        if (want_p) gather_row_async<T>(ps32 + (unsigned)v * 4u * (unsigned)sizeof(T), pf, gv, a.ld_in);
                    }
                }
                cp_async_arrive_noinc(full(s));
            }
#if APL_L2_PREFETCH >= 2
            // The vertex rows of tile it + 1 into L2 now (its table was requested above and the producer has nothing else
            // to do until the next stage frees up): the gather of the next iteration then hits L2 instead of DRAM.  Matters
            // when the nodal fields do not fit the 126 MB L2 (64 M tets: 2 x 156 MB).
            if (it + 1 < my_tiles) {
                const int sv1 = (it + 1) % SV;
                mbar_wait(vfull(sv1), (unsigned)((it + 1) / SV) & 1u);
                const int* verts1 = reinterpret_cast<const int*>(vring + (size_t)sv1 * PC::kVtabBytes + PC::oVerts);
                const int nv1 = __shfl_sync(0xffffffffu, h_nxt.y, 0) >> 16;
                const size_t row = (size_t)a.ld_in * sizeof(T);
                for (int v = lane + 32 * pw; v < nv1; v += 32 * PW) {
                    const size_t off = (size_t)verts1[v] * row;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(a.u) + off));
                    if (want_p) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(pf) + off));
                }
            }
#endif
            h_cur = h_nxt;
            h_nxt = h_nn;
            h_nn = h_n3;
        }
    } else {
        // ================================= consumer warps ================================
        if constexpr (PC::kRegRealloc) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // COMPUTE(it): returns the tile's vertex count (the header lives in the stage, which is released here)
        auto compute = [&](int it) -> int {
            const int s = it % S;
            unsigned char* st = stages + (size_t)s * PC::kStageBytes;
            mbar_wait(full(s), (unsigned)(it / S) & 1u);
            const int4 h = *reinterpret_cast<const int4*>(st + PC::oHdr);
            const int n_tets = h.y & 0xffff;
            const T* us = reinterpret_cast<const T*>(st + PC::oVbuf);
            const T* ps = us + 4 * kTileVerts;
            T* sl = reinterpret_cast<T*>(smem_raw + PC::oSlotBuf + (size_t)(it % NB) * PC::kSlotBytes);
            if constexpr (NOUT > 0) {
                // slot buffer it % NB is free once every consumer has finished REDUCE(it - NB)
                if (it >= NB) mbar_wait(sfree(it % NB), (unsigned)(it / NB - 1) & 1u);
            }
            if (tid < n_tets) {
                Rec<T, PC::NREC> rec;
#pragma unroll
                for (int k = 0; k < PC::NPL; ++k)
                    rec.q[k] = reinterpret_cast<const uint4*>(st + PC::oPlanes + (size_t)k * kTileTets * 16)[tid];
                const uchar4 lc = reinterpret_cast<const uchar4*>(st + PC::oConn)[tid];
                ushort4 s4 = make_ushort4(0, 0, 0, 0);
                if constexpr (NOUT > 0) s4 = reinterpret_cast<const ushort4*>(st + PC::oSlots)[tid];
                if constexpr (Cfg::kPackedIn)
                    tile_compute_packed<KIND, OPS>(rec.s, lc, s4, (const float*)us, (float*)sl, e_acc, q_acc);
                else
                    tile_compute<T, KIND, OPS>(rec.s, lc, s4, us, ps, axpy, alpha, sl, e_acc, q_acc);
            }
            if constexpr (NOUT > 0) mbar_arrive(sfull(it % NB));
            mbar_arrive(empty(s));
            return h.y >> 16;
        };
        if constexpr (NOUT == 0) {
            for (int it = 0; it < my_tiles; ++it) compute(it);
        } else {
            auto reduce = [&](int it, int n_verts) {
                const unsigned char* vt = vring + (size_t)(it % SV) * PC::kVtabBytes;
                const T* sl = reinterpret_cast<const T*>(smem_raw + PC::oSlotBuf + (size_t)(it % NB) * PC::kSlotBytes);
                mbar_wait(sfull(it % NB), (unsigned)(it / NB) & 1u);
                tile_reduce_flush<T, OPS, NC, PC::kNSlots, true>(tid, it, n_verts, vt + PC::oVperm,
                                                           reinterpret_cast<const unsigned short*>(vt + PC::oVoff),
                                                           reinterpret_cast<const int*>(vt + PC::oVerts), sl, a,
                                                           [&] { mbar_arrive(sfree(it % NB)); });
            };
            if constexpr (NB == 2) {
                int nv_cur = my_tiles > 0 ? compute(0) : 0;
                for (int it = 0; it < my_tiles; ++it) {
                    const int nv_next = it + 1 < my_tiles ? compute(it + 1) : 0;
                    reduce(it, nv_cur);
                    nv_cur = nv_next;
                }
            } else {
                for (int it = 0; it < my_tiles; ++it) {
                    const int nv = compute(it);
                    reduce(it, nv);
                }
            }
        }
    }
    if constexpr (Cfg::kFun || Cfg::kQuad)
        finish_scalars<T, PC::kThreads>(e_acc, q_acc, a.partials, a.counter, Cfg::kFun ? a.fun : nullptr,
                                        Cfg::kQuad ? a.quad : nullptr, (Cfg::kFun && a.fun_d) ? a.fun_d + joff : nullptr,
                                        Cfg::kQuad ? a.quad_d : nullptr);
}

// ---- ATOMIC kernel (baseline) -----------------------------------------------------------------------

template <typename T, int KIND, int OPS>
__global__ void __launch_bounds__(kTileTets) fem_atomic_kernel(const FemArgs<T> a) {
    using Cfg = TileCfg<T, OPS>;
    constexpr int NREC = RecSize<KIND>::value;
    const int tid = threadIdx.x;
    double e_acc = 0.0, q_acc = 0.0;
    const int joff = fem_joff(a);
    if (fem_skip(a, joff)) return;
    const bool axpy = a.axpy_p != nullptr;
    const T alpha = axpy ? (T)__ldcg(a.scal + a.alpha_idx + joff) : (T)0;
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
                    if constexpr (Cfg::kGrad) { if (a.grad) red_add(a.grad + o + i, g[c][i]); }
                    if constexpr (Cfg::kDiag) { if (a.diag) red_add(a.diag + o + i, dg[c][i]); }
                    if constexpr (Cfg::kProd) { if (a.prod) red_add(a.prod + o + i, hp[c][i]); }
                }
            }
        }
    }
    if constexpr (Cfg::kFun || Cfg::kQuad)
        finish_scalars<T, kTileTets>(e_acc, q_acc, a.partials, a.counter, Cfg::kFun ? a.fun : nullptr,
                                     Cfg::kQuad ? a.quad : nullptr, (Cfg::kFun && a.fun_d) ? a.fun_d + joff : nullptr,
                                     Cfg::kQuad ? a.quad_d : nullptr);
}

// ---- launcher ---------------------------------------------------------------------------------------

// Declared per (T, KIND); defined in fem_inst.cu, one translation unit per pair.
template <typename T, int KIND>
int launch_fem(const apl_fem* fem, int ops, const FemArgs<T>& args, int scatter, cudaStream_t stream);

// Grid of a persistent kernel: resident CTAs per SM x SMs.  The opt-in to more than 48 KB of dynamic shared
// memory and the occupancy are per DEVICE (a process may hold handles on several GPUs), so they are cached per
// (instantiation, device) under a mutex.
constexpr int kMaxDevices = 64;
template <typename K>
int resident_blocks(K kern, int device, int threads, size_t smem, int* cache, std::mutex& mu, int& out) {
    if (device < 0 || device >= kMaxDevices) {
        set_error("device index outside the range the launch cache supports");
        return APL_ERR_INVALID;
    }
    std::lock_guard<std::mutex> lock(mu);
    if (cache[device] == 0) {
        if (smem > 0)
            APL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int b = 0;
        APL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, threads, smem));
        cache[device] = b > 0 ? b : 1;
    }
    out = cache[device];
    return APL_OK;
}

template <typename T, int KIND, int OPS>
int launch_one(const apl_fem* fem, const FemArgs<T>& args, int scatter, cudaStream_t stream) {
    using Cfg = TileCfg<T, OPS>;
    if (args.n_tiles == 0) return APL_OK;
    static std::mutex mu;
    static int cache[3][kMaxDevices] = {};
    int blocks_per_sm = 1;
    auto grid_of = [&](int per_sm) {
        int grid = fem->num_sms * per_sm;
        if (grid > args.n_tiles) grid = args.n_tiles;
        if (grid > fem->max_grid) grid = fem->max_grid;
        return grid;
    };
    if (scatter == APL_SCATTER_TILE) {
        using PC = PipeCfg<T, KIND, OPS>;
        auto kern = fem_pipe_kernel<T, KIND, OPS>;
        int rc = resident_blocks(kern, fem->device, PC::kThreads, PC::kSmemBytes, cache[0], mu, blocks_per_sm);
        if (rc != APL_OK) return rc;
        kern<<<grid_of(blocks_per_sm), PC::kThreads, PC::kSmemBytes, stream>>>(args);
    } else if constexpr ((OPS & (APL_OP_HESS_OFFD | APL_OP_PSD)) != 0) {
        set_error("apl_fem_eval: APL_OP_HESS_OFFD / APL_OP_PSD are implemented by the TILE assembly only");
        return APL_ERR_INVALID;
    } else if (scatter == APL_SCATTER_TILE_SIMPLE) {
        auto kern = fem_tile_kernel<T, KIND, OPS>;
        int rc = resident_blocks(kern, fem->device, kTileTets, Cfg::kSmemBytes, cache[1], mu, blocks_per_sm);
        if (rc != APL_OK) return rc;
        kern<<<grid_of(blocks_per_sm), kTileTets, Cfg::kSmemBytes, stream>>>(args);
    } else {
        auto kern = fem_atomic_kernel<T, KIND, OPS>;
        int rc = resident_blocks(kern, fem->device, kTileTets, 0, cache[2], mu, blocks_per_sm);
        if (rc != APL_OK) return rc;
        kern<<<grid_of(blocks_per_sm), kTileTets, 0, stream>>>(args);
    }
    APL_CUDA_CHECK(cudaGetLastError());
    return APL_OK;
}

// The opt-in supersets (vertex-block off-diagonals, eigenvalue-clamped Hessian) are instantiated in their OWN
// translation units (fem_inst.cu with -DAPL_INST_SUPERSET): instantiating them next to the classic operator sets changed
// nvcc's inlining decisions for the classic kernels (fused SNH+ARAP E+g+Hp: 40 -> 200 bytes of spill stack, 23 -> 13.7
// G tets/s at 8 M tets, round-2 run r2m).  Declared per (T, KIND) like launch_fem; defined in fem_inst.cu.
template <typename T, int KIND>
int launch_fem_superset(const apl_fem* fem, int ops, const FemArgs<T>& args, int scatter, cudaStream_t stream);

template <typename T, int KIND>
int launch_fem_superset_impl(const apl_fem* fem, int ops, const FemArgs<T>& args, int scatter, cudaStream_t stream) {
    // ops carries HESS_OFFD and / or PSD (the caller has already dropped PSD where it does not apply)
    if (scatter != APL_SCATTER_TILE) {
        set_error("apl_fem_eval: APL_OP_HESS_OFFD / APL_OP_PSD are implemented by the TILE assembly only");
        return APL_ERR_INVALID;
    }
    constexpr int BLK = APL_OP_FUN | APL_OP_GRAD | APL_OP_HESS_DIAG | APL_OP_HESS_OFFD;
    const int psd = ops & APL_OP_PSD, base = ops & ~APL_OP_PSD;
    if (base & APL_OP_HESS_OFFD) {
        if (base & ~BLK) {
            set_error("apl_fem_eval: APL_OP_HESS_OFFD combines with FUN, GRAD and HESS_DIAG only");
            return APL_ERR_INVALID;
        }
        return psd ? launch_one<T, KIND, BLK | APL_OP_PSD>(fem, args, scatter, stream)
                   : launch_one<T, KIND, BLK>(fem, args, scatter, stream);
    }
    // PSD without OFFD: the quadratic form alone (PNCG pass B), and / or fun / grad / diag / prod in one pass
    if (base & APL_OP_HESS_QUAD) {
        int rc = launch_one<T, KIND, APL_OP_HESS_QUAD | APL_OP_PSD>(fem, args, scatter, stream);
        if (rc != APL_OK) return rc;
        if (!(base & (APL_OP_HESS_DIAG | APL_OP_HESS_PROD))) {
            const int rest = base & ~APL_OP_HESS_QUAD;     // fun / grad only: no Hessian term left to project
            return rest ? launch_fem<T, KIND>(fem, rest, args, scatter, stream) : APL_OK;
        }
    }
    // no hess_prod left (fun / grad / diag only, e.g. PNCG pass A with the PSD option but the scalar preconditioner): the
    // block instantiation, whose operator set does not read p (its off-diagonal output is NULL and skipped)
    if (!(base & APL_OP_HESS_PROD)) return launch_one<T, KIND, BLK | APL_OP_PSD>(fem, args, scatter, stream);
    return launch_one<T, KIND, 15 | APL_OP_PSD>(fem, args, scatter, stream);
}

template <typename T, int KIND>
int launch_fem_impl(const apl_fem* fem, int ops, const FemArgs<T>& args, int scatter, cudaStream_t stream) {
    if (ops & (APL_OP_HESS_OFFD | APL_OP_PSD)) {
        if (KIND == APL_KIND_ARAP) ops &= ~APL_OP_PSD;   // the clamped twist rates already are the projection
        if (!(ops & (APL_OP_HESS_DIAG | APL_OP_HESS_OFFD | APL_OP_HESS_PROD | APL_OP_HESS_QUAD))) ops &= ~APL_OP_PSD;
    }
    if (ops & (APL_OP_HESS_OFFD | APL_OP_PSD)) return launch_fem_superset<T, KIND>(fem, ops, args, scatter, stream);
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
