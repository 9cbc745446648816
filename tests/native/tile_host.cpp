// TEST-ONLY harness: compiles apple_b200/csrc/tile_logic.cuh (the consumer-side logic of the element kernels:
// record / connectivity decoding, corner gather, slot stores, per-lane slot reduction) for the HOST and replays it thread by thread, tile by tile, on the packed tables of a
// host-only handle.  What the GPU adds on top -- the producer's bulk copies, the mbarriers, the half-warp
// shuffle and the global REDs -- is emulated here in the obvious way.  Never part of the product library.
#include <cstdint>
#include <vector>

#include "../../apple_b200/csrc/tile_logic.cuh"

using namespace apl;

template <typename T, int KIND, int OPS>
static void run(int64_t n_tiles, const int32_t* tiles, const uint8_t* conn, const uint16_t* slots,
                const int32_t* tile_verts, const uint16_t* tile_voff, const uint8_t* tile_vperm, const T* planes,
                int64_t plane_stride, const T* u, const T* p, double alpha, T* grad, T* diag, T* prod, double* fun, double* quad) {
    using Cfg = TileCfg<T, OPS>;
    const bool axpy = alpha != 0.0;   // line-search trial point u + alpha p formed in the gather (PNCG pass A)
    constexpr int NOUT = Cfg::NOUT, SS = Cfg::SS;
    constexpr int NREC = RecSize<KIND>::value;
    constexpr int NC = kTileTets;                 // consumer threads
    constexpr int NSLOTS = kSlotsAlloc;
    constexpr int VEC = 16 / (int)sizeof(T);
    std::vector<T> vbuf((size_t)kTileVerts * (Cfg::VB > 8 ? Cfg::VB : 8) + 16), sl((size_t)NSLOTS * (SS > 0 ? SS : 1) + 16);
    double e_acc = 0, q_acc = 0;
    for (int64_t tile = 0; tile < n_tiles; ++tile) {
        const int32_t* h = tiles + 6 * tile;
        const int ts = h[0], n_tets = h[1], vs = h[2], n_verts = h[3], vo = h[4];
        // producer: gather the tile's vertices into 4-scalar rows (u, then p)
        T* us = vbuf.data();
        T* ps = vbuf.data() + 4 * kTileVerts;
        for (int v = 0; v < n_verts; ++v) {
            const int gv = tile_verts[vs + v];
            for (int i = 0; i < 3; ++i) {
                us[4 * v + i] = u[3 * (int64_t)gv + i];
                ps[4 * v + i] = p[3 * (int64_t)gv + i];
            }
            us[4 * v + 3] = ps[4 * v + 3] = (T)0;
        }
        if constexpr (Cfg::kPackedIn) {   // rows [ux uy px py], then rows [uz pz] (fem_kernels.cuh: gather_row_packed)
            for (int v = 0; v < n_verts; ++v) {
                const int gv = tile_verts[vs + v];
                T* ra = vbuf.data() + 4 * v;
                T* rb = vbuf.data() + 4 * kTileVerts + 2 * v;
                ra[0] = u[3 * (int64_t)gv]; ra[1] = u[3 * (int64_t)gv + 1]; ra[2] = p[3 * (int64_t)gv]; ra[3] = p[3 * (int64_t)gv + 1];
                rb[0] = u[3 * (int64_t)gv + 2]; rb[1] = p[3 * (int64_t)gv + 2];
            }
        }
        for (auto& x : sl) x = (T)(0.0 / 0.0);   // NaN: a slot that is read without having been written shows up
        // compute phase: one consumer thread per tet
        auto load_rec = [&](int64_t pos, Rec<T, NREC>& r) {
            for (int k = 0; k < Rec<T, NREC>::NPL; ++k)
                for (int j = 0; j < VEC; ++j) r.s[k * VEC + j] = planes[((size_t)k * plane_stride + pos) * VEC + j];
        };
        {
            for (int tid = 0; tid < NC; ++tid) {
                if (tid >= n_tets) continue;
                Rec<T, NREC> r;
                load_rec(ts + tid, r);
                uchar4 lc;
                ushort4 s4;
                std::memcpy(&lc, conn + 4 * ((size_t)ts + tid), 4);
                std::memcpy(&s4, slots + 4 * ((size_t)ts + tid), 8);
                if constexpr (Cfg::kPackedIn)
                    tile_compute_packed<KIND, OPS>((const float*)r.s, lc, s4, (const float*)vbuf.data(), (float*)sl.data(), e_acc, q_acc);
                else
                    tile_compute<T, KIND, OPS>(r.s, lc, s4, us, ps, axpy, (T)alpha, sl.data(), e_acc, q_acc);
            }
        }
        if constexpr (NOUT > 0) {
            // reduce phase: warp by warp, iteration by iteration; lanes l and l + 16 exchange through the shuffle and
            // the sums go straight to global memory (fem_kernels.cuh: tile_reduce_flush)
            const unsigned char* vperm = tile_vperm + vs;
            const unsigned short* voff = tile_voff + vo;
            for (int warp = 0; warp < NC / 32; ++warp)
                for (int t0 = warp * 16; t0 < ((n_verts + 15) & ~15); t0 += NC / 2) {
                    int v[32];
                    T acc[32][3 * NOUT];
                    for (int lane = 0; lane < 32; ++lane) {
                        const int tid = 32 * warp + lane, half = (tid >> 4) & 1;
                        tile_reduce_lane<T, OPS, NSLOTS>(half, t0 + (tid & 15), n_verts, vperm, voff, sl.data(), v[lane], acc[lane]);
                    }
                    for (int lane = 0; lane < 16; ++lane) {
                        if (t0 + lane >= n_verts) continue;
                        T sum[3 * NOUT];
                        for (int j = 0; j < 3 * NOUT; ++j) sum[j] = acc[lane][j] + acc[lane + 16][j];
                        const int64_t gv = tile_verts[vs + v[lane]];
                        if constexpr (Cfg::kPacked) {   // slots hold [gx gy hx hy | gz hz] (tile_reduce_flush maps them back)
                            const T g3[3] = {sum[0], sum[1], sum[4]}, h3[3] = {sum[2], sum[3], sum[5]};
                            for (int i = 0; i < 3; ++i) { grad[3 * gv + i] += g3[i]; prod[3 * gv + i] += h3[i]; }
                            continue;
                        }
                        int k = 0;
                        if constexpr (Cfg::kGrad) { for (int i = 0; i < 3; ++i) grad[3 * gv + i] += sum[k + i]; k += 3; }
                        if constexpr (Cfg::kDiag) { for (int i = 0; i < 3; ++i) diag[3 * gv + i] += sum[k + i]; k += 3; }
                        if constexpr (Cfg::kProd) { for (int i = 0; i < 3; ++i) prod[3 * gv + i] += sum[k + i]; k += 3; }
                    }
                }
        }
    }
    *fun += e_acc;
    *quad += q_acc;
}

template <typename T, int KIND>
static int by_ops(int ops, int64_t n_tiles, const int32_t* tiles, const uint8_t* conn, const uint16_t* slots,
                  const int32_t* tv, const uint16_t* voff, const uint8_t* vperm, const void* planes, int64_t stride,
                  const void* u, const void* p, double alpha, void* grad, void* diag, void* prod, double* fun, double* quad) {
#define GO(O)                                                                                                    \
    run<T, KIND, O>(n_tiles, tiles, conn, slots, tv, voff, vperm, (const T*)planes, stride, (const T*)u,  \
                    (const T*)p, alpha, (T*)grad, (T*)diag, (T*)prod, fun, quad)
    switch (ops) {
        case 11: GO(11); return 0;
        case 7: GO(7); return 0;
        case 15: GO(15); return 0;
        case 16: GO(16); return 0;
        case 2: GO(2); return 0;
        default: return -1;
    }
#undef GO
}

// kind: APL_KIND_*; ops in {2, 7, 11, 15, 16}; alpha != 0 evaluates at u + alpha p (ops without
// hess_prod / hess_quad only, as in PNCG's trial passes).  Outputs are accumulated (caller zeroes).
extern "C" int tile_emulate(int kind, int is_f64, int ops, int64_t n_tiles, const int32_t* tiles,
                            const uint8_t* conn, const uint16_t* slots, const int32_t* tv, const uint16_t* voff,
                            const uint8_t* vperm, const void* planes, int64_t stride, const void* u, const void* p,
                            double alpha, void* grad, void* diag, void* prod, double* fun, double* quad) {
#define ARGS ops, n_tiles, tiles, conn, slots, tv, voff, vperm, planes, stride, u, p, alpha, grad, diag, prod, fun, quad
#define KINDS(T)                                                           \
    switch (kind) {                                                        \
        case APL_KIND_SNH: return by_ops<T, APL_KIND_SNH>(ARGS);           \
        case APL_KIND_ARAP: return by_ops<T, APL_KIND_ARAP>(ARGS);         \
        case APL_KIND_SNH_ARAP: return by_ops<T, APL_KIND_SNH_ARAP>(ARGS); \
        default: return by_ops<T, APL_KIND_SNH_MUSCLE>(ARGS);              \
    }
    if (is_f64) { KINDS(double) } else { KINDS(float) }
    return -1;
}
