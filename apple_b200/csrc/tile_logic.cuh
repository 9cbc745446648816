// Per-tile CONSUMER logic of the element kernels -- everything between "the stage is in shared memory" and "the
// per-vertex sums are ready for the global RED": record / connectivity decoding, the corner gather, the slot
// stores, the per-lane part of the slot reduction and the parked sums.  No CUDA-only
// construct is used here (the shuffle, the barriers and the REDs stay in fem_kernels.cuh), so that the very
// same code is compiled for the HOST by tests/native/tile_host.cpp and replayed thread by thread on the packed
// tables against the oracle (the product only ever runs it on the GPU).
#pragma once

#include "common.h"
#include "elem_math.cuh"

#if defined(__CUDACC__)
#define APL_TL __device__ __forceinline__
#else
#define APL_TL inline
#include <cstring>
// minimal stand-ins for the CUDA vector types / bit casts used below (host test build only)
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
struct ushort4 { unsigned short x, y, z, w; };
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline double2 make_double2(double x, double y) { return {x, y}; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline long long __double_as_longlong(double d) { long long l; std::memcpy(&l, &d, 8); return l; }
inline double __longlong_as_double(long long l) { double d; std::memcpy(&d, &l, 8); return d; }
#endif

namespace apl {

template <typename T, int NREC>
struct Rec {
    static constexpr int VEC = 16 / (int)sizeof(T);
    static constexpr int NPL = (NREC + VEC - 1) / VEC;
    union {
        uint4 q[NPL];
        T s[NPL * VEC];
    };
};

// ---- compile-time layout of one instantiation ---------------------------------------------------

template <typename T, int OPS>
struct TileCfg {
    static constexpr bool kFun = (OPS & APL_OP_FUN) != 0;
    static constexpr bool kGrad = (OPS & APL_OP_GRAD) != 0;
    static constexpr bool kDiag = (OPS & APL_OP_HESS_DIAG) != 0;
    // third nodal output: H p (HESS_PROD) or the off-diagonals of the vertex blocks (HESS_OFFD) -- never both
    static constexpr bool kProd = (OPS & (APL_OP_HESS_PROD | APL_OP_HESS_OFFD)) != 0;
    static constexpr bool kQuad = (OPS & APL_OP_HESS_QUAD) != 0;
    static constexpr bool kNeedP = (OPS & (APL_OP_HESS_PROD | APL_OP_HESS_QUAD)) != 0;
    static constexpr int NOUT = (kGrad ? 1 : 0) + (kDiag ? 1 : 0) + (kProd ? 1 : 0);
    // packed fp32 forms (elem_math.cuh: elem_eval_packed).  kPackedIn: the operator set reads u AND p -> interleaved
    // vertex rows [ux uy px py | uz pz - -] and {F, dF} in packed instructions; kPacked: in addition the nodal outputs are
    // exactly {grad, hess_prod} -> slots [gx gy hx hy | gz hz] from packed {g, Hp}.
#ifndef APL_PACKED
#define APL_PACKED 1
#endif
    static constexpr bool kPackedIn = APL_PACKED && sizeof(T) == 4 && kNeedP && (OPS & APL_OP_PSD) == 0;
    static constexpr bool kPacked = kPackedIn && kGrad && (OPS & APL_OP_HESS_PROD) != 0 &&
                                    (OPS & (APL_OP_HESS_DIAG | APL_OP_HESS_OFFD)) == 0;
    // scalars per slot: 3*NOUT rounded up to whole 16-byte planes plus, for fp32, one 8-byte tail plane
    static constexpr int SS = (NOUT == 0) ? 0
                              : (NOUT == 1) ? 4
                              : (sizeof(T) == 4) ? (NOUT == 2 ? 6 : 10) : (NOUT == 2 ? 6 : 10);
    static constexpr int kNSlots = kSlotsAlloc;
    // per-vertex buffer: u (4 scalars) and p (4 scalars) during compute, then reused for the 3*NOUT
    // reduced sums of each vertex between the reduce and the flush phase
    static constexpr int VB = (3 * NOUT > 8) ? 12 : 8;
    static constexpr size_t kVbufBytes = (size_t)kTileVerts * VB * sizeof(T);
    static constexpr size_t kSlotBytes = (size_t)kNSlots * SS * sizeof(T);
    static constexpr size_t kVoffRaw = ((size_t)(kTileVerts + 1) * 2 + 15) / 16 * 16;  // n_verts+1 uint16
    static constexpr size_t kVoffBytes = NOUT ? kVoffRaw : 0;
    static constexpr size_t kVpermBytes = NOUT ? (size_t)kTileVerts : 0;
    // simple (unpipelined) kernel
    static constexpr size_t kSmemBytes = kVbufBytes + kSlotBytes + kVoffBytes + kVpermBytes;
};

// The slot buffer is split into 16-byte planes: vector q of slot s lives at sl + (q * kNSlots + s) * 16 B.
// One thread reading consecutive slots of "its" vertex and a warp of such threads then touch
// neighbouring 16-byte words (see tile_reduce), instead of words a whole slot stride apart.

template <typename T, int SS, int NSLOTS = kSlotsAlloc>
APL_TL void store_slot_planes(T* sl, int s, const T* v) {
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int NQ = SS / VEC;  // full 16-byte planes; fp32 may add one 8-byte tail plane
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        uint4 w;
        if constexpr (sizeof(T) == 4) {
            w = make_uint4(__float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]), __float_as_uint(v[4 * q + 2]),
                           __float_as_uint(v[4 * q + 3]));
        } else {
            const unsigned long long a = (unsigned long long)__double_as_longlong(v[2 * q]);
            const unsigned long long b = (unsigned long long)__double_as_longlong(v[2 * q + 1]);
            w = make_uint4((unsigned)a, (unsigned)(a >> 32), (unsigned)b, (unsigned)(b >> 32));
        }
        reinterpret_cast<uint4*>(sl)[q * NSLOTS + s] = w;
    }
    if constexpr (sizeof(T) == 4 && SS % 4 == 2) {
        float2* tail = reinterpret_cast<float2*>(reinterpret_cast<uint4*>(sl) + NQ * NSLOTS);
        tail[s] = make_float2((float)v[4 * NQ], (float)v[4 * NQ + 1]);
    }
}

template <typename T, int SS, int NSLOTS = kSlotsAlloc>
APL_TL void load_slot_planes(const T* sl, int s, T* v) {
    constexpr int VEC = 16 / (int)sizeof(T);
    constexpr int NQ = SS / VEC;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const uint4 w = reinterpret_cast<const uint4*>(sl)[q * NSLOTS + s];
        if constexpr (sizeof(T) == 4) {
            v[4 * q] = __uint_as_float(w.x); v[4 * q + 1] = __uint_as_float(w.y);
            v[4 * q + 2] = __uint_as_float(w.z); v[4 * q + 3] = __uint_as_float(w.w);
        } else {
            v[2 * q] = __longlong_as_double((long long)(((unsigned long long)w.y << 32) | w.x));
            v[2 * q + 1] = __longlong_as_double((long long)(((unsigned long long)w.w << 32) | w.z));
        }
    }
    if constexpr (sizeof(T) == 4 && SS % 4 == 2) {
        const float2* tail = reinterpret_cast<const float2*>(reinterpret_cast<const uint4*>(sl) + NQ * NSLOTS);
        const float2 w = tail[s];
        v[4 * NQ] = (T)w.x; v[4 * NQ + 1] = (T)w.y;
    }
}

template <typename T, int SS>
APL_TL void store_slot(T* dst, const T* v) {
    if constexpr (sizeof(T) == 4) {
        static_assert(SS % 4 == 0, "fp32 slots are whole float4s");
#pragma unroll
        for (int k = 0; k < SS / 4; ++k)
            reinterpret_cast<float4*>(dst)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    } else {
        static_assert(SS % 2 == 0, "fp64 slots are whole double2s");
#pragma unroll
        for (int k = 0; k < SS / 2; ++k) reinterpret_cast<double2*>(dst)[k] = make_double2(v[2 * k], v[2 * k + 1]);
    }
}

template <typename T, int SS>
APL_TL void load_slot(const T* src, T* v) {
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int k = 0; k < SS / 4; ++k) {
            const float4 q = reinterpret_cast<const float4*>(src)[k];
            v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SS / 2; ++k) {
            const double2 q = reinterpret_cast<const double2*>(src)[k];
            v[2 * k] = q.x; v[2 * k + 1] = q.y;
        }
    }
}

// One tet: gather its corners from the shared vertex buffer, evaluate, write one slot per corner.
template <typename T, int KIND, int OPS>
APL_TL void tile_compute(const T* rec, uchar4 lc, ushort4 s4, const T* us, const T* ps, bool axpy,
                                             T alpha, T* sl, double& e_acc, double& q_acc) {
    using Cfg = TileCfg<T, OPS>;
    constexpr int NOUT = Cfg::NOUT, SS = Cfg::SS;
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
        } else {
            if (axpy) {  // line-search trial point x + alpha p, never materialised in global memory
                load_slot<T, 4>(ps + 4 * l[c], tmp);
                uc[c][0] += alpha * tmp[0]; uc[c][1] += alpha * tmp[1]; uc[c][2] += alpha * tmp[2];
            }
        }
    }
    T psi = 0, quad = 0;
    T g[4][3], dg[4][3], hp[4][3];
    elem_eval<T, KIND, OPS>(rec, uc, pc, psi, quad, g, dg, hp);
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
            store_slot_planes<T, SS>(sl, sidx[c], v);
        }
    }
}

// Packed form of tile_compute (TileCfg::kPackedIn): vb holds the tile's vertices as kTileVerts rows [ux uy px py]
// followed by kTileVerts rows [uz pz].
template <int KIND, int OPS>
APL_TL void tile_compute_packed(const float* rec, uchar4 lc, ushort4 s4, const float* vb, float* sl, double& e_acc,
                                double& q_acc) {
    using Cfg = TileCfg<float, OPS>;
    constexpr int NOUT = Cfg::NOUT, SS = Cfg::SS;
    static_assert(Cfg::kPackedIn && (!Cfg::kPacked || SS == 6), "packed slots are one 16-byte plane plus the 8-byte tail");
    const int l[4] = {lc.x, lc.y, lc.z, lc.w};
    f32x2 u01[4], p01[4], upz[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        // 16-byte rows [ux uy px py] (same bank pattern as the scalar form's u rows, which the tile tables make
        // conflict-free) followed by 8-byte rows [uz pz]; 32-byte interleaved rows doubled the bank conflicts (run r2o)
        const float4 q = *reinterpret_cast<const float4*>(vb + 4 * l[c]);
        const float2 r = *reinterpret_cast<const float2*>(vb + 4 * kTileVerts + 2 * l[c]);
        u01[c] = p2_make(q.x, q.y);
        p01[c] = p2_make(q.z, q.w);
        upz[c] = p2_make(r.x, r.y);
    }
    float psi = 0, quad = 0;
    f32x2 g01[4], h01[4], gzhz[4];
    float g[4][3], dg[4][3], hp[4][3];
    elem_eval_packed<KIND, OPS, Cfg::kPacked>(rec, u01, p01, upz, psi, quad, g01, h01, gzhz, g, dg, hp);
    if constexpr (Cfg::kFun) e_acc += (double)psi;
    if constexpr (Cfg::kQuad) q_acc += (double)quad;
    if constexpr (NOUT > 0) {
        const int sidx[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float v[SS];
            if constexpr (Cfg::kPacked) {
                v[0] = g01[c].lo; v[1] = g01[c].hi; v[2] = h01[c].lo; v[3] = h01[c].hi; v[4] = gzhz[c].lo; v[5] = gzhz[c].hi;
            } else {
                int k = 0;
                if constexpr (Cfg::kGrad) { v[k] = g[c][0]; v[k + 1] = g[c][1]; v[k + 2] = g[c][2]; k += 3; }
                if constexpr (Cfg::kDiag) { v[k] = dg[c][0]; v[k + 1] = dg[c][1]; v[k + 2] = dg[c][2]; k += 3; }
                if constexpr (Cfg::kProd) { v[k] = hp[c][0]; v[k + 1] = hp[c][1]; v[k + 2] = hp[c][2]; k += 3; }
#pragma unroll
                for (int j = 3 * NOUT; j < SS; ++j) v[j] = 0.0f;
            }
            store_slot_planes<float, SS>(sl, sidx[c], v);
        }
    }
}

// Two threads (lanes l and l+16 of a warp) together sum the slot range of one vertex, taken in valence
// order (balanced trip counts, all consumer warps busy), combine with one shuffle, and the lower lane
// parks the 3*NOUT sums in the vertex buffer at the vertex's local id.  The 16 lanes of a half warp
// handle 16 consecutive vertices of the reduce order whose ranges start in distinct bank groups
// (tiling.cpp), so their 16- and 8-byte reads do not conflict.
//
// tile_reduce_lane: what ONE lane does for reduce position t before the exchange with its partner lane.
template <typename T, int OPS, int NSLOTS>
APL_TL void tile_reduce_lane(int half, int t, int n_verts, const unsigned char* vperm, const unsigned short* voff,
                             const T* sl, int& v, T* acc) {
    using Cfg = TileCfg<T, OPS>;
    constexpr int NOUT = Cfg::NOUT, SS = Cfg::SS;
    v = 0;
    int s0 = 0, cnt = 0;
    if (t < n_verts) {
        v = vperm[t];
        // voff is in reduce order: bits 0..11 = first slot of the range, bits 12..15 = unused pad slots
        // after it (tiling.cpp chooses order and pads so that the 16 lanes of a group start in
        // distinct bank groups: conflict-free 16/8-byte loads).
        const unsigned a0 = voff[t], a1 = voff[t + 1];
        s0 = (int)(a0 & 0x0fffu);
        cnt = (int)(a1 & 0x0fffu) - s0 - (int)(a0 >> 12);
    }
#pragma unroll
    for (int j = 0; j < 3 * NOUT; ++j) acc[j] = (T)0;
    // APL_REDUCE_UNROLL slots per trip with all loads issued before the first add (a lane past its range re-reads its
    // first slot with weight 0, so the trip is branch-free).  Measured (run r2n, 8 M tets, G tets/s, U = 1 / 2 / 4):
    // SNH 33.7 / 32.9 / 32.3, SNH+ARAP 23.3 / 22.6 / 23.0 -- the phase holds 30 % of the SNH kernel's stall samples, but
    // batching its loads does not shorten it (other warps already cover the latency), so the default stays serial.
#ifndef APL_REDUCE_UNROLL
#define APL_REDUCE_UNROLL 1
#endif
    constexpr int U = APL_REDUCE_UNROLL;
    for (int i = half; i < cnt; i += 2 * U) {
        T val[U][SS];
        T wgt[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const bool ok = i + 2 * k < cnt;
            wgt[k] = ok ? (T)1 : (T)0;
            load_slot_planes<T, SS, NSLOTS>(sl, s0 + (ok ? i + 2 * k : i), val[k]);
        }
#pragma unroll
        for (int k = 0; k < U; ++k)
#pragma unroll
            for (int j = 0; j < 3 * NOUT; ++j) acc[j] += wgt[k] * val[k][j];
    }
}

// the lower lane parks the sums of local vertex v; the flush phase reads them back by local id
template <typename T, int OPS>
APL_TL void tile_reduce_park(T* vbuf, int v, const T* acc) {
    constexpr int NOUT = TileCfg<T, OPS>::NOUT;
#pragma unroll
    for (int j = 0; j < 3 * NOUT; ++j) vbuf[v * (3 * NOUT) + j] = acc[j];
}
template <typename T, int OPS>
APL_TL void tile_flush_read(const T* vbuf, int v, T* acc) {
    constexpr int NOUT = TileCfg<T, OPS>::NOUT;
#pragma unroll
    for (int j = 0; j < 3 * NOUT; ++j) acc[j] = vbuf[v * (3 * NOUT) + j];
}

}  // namespace apl
