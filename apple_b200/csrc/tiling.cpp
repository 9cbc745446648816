// Host-side mesh packing: Morton ordering of the tets and greedy tiling.
//
// A tile is a run of <= 256 consecutive tets (in packed order) that together touch <= 192 distinct
// vertices.  For every tile we store
//   * the global ids of the vertices it touches, indexed by tile-local id   (tile_verts)
//   * the local ids in REDUCE order (about decreasing valence)             (tile_vperm)
//   * per tet, the tile-local id of each corner as one byte                (conn)
//   * per tet corner, a slot: slots are grouped by local vertex, so the element kernel can write
//     every corner contribution to its own shared-memory slot (no atomics) and a second phase sums
//     each vertex's contiguous slot range                                     (slots, tile_voff)
// This replaces the reference's per-tet global connectivity (`cells`, warp/fem/_base.py:59-74) and
// its 12 global atomics per tet per field (warp/fem/_base.py:288-289).
//
// Everything that is free in this encoding is chosen to keep the kernels' warp-wide shared-memory
// accesses free of bank conflicts (a warp-wide 16-byte access is served per quarter warp, an 8-byte
// access per half warp; lanes that touch different addresses in the same bank group serialise):
//   * LOCAL VERTEX IDS.  The corner gather of 8 consecutive tets reads row `16 B * local id`: the
//     ids are a balanced colouring (id mod 8; mod 4 for 32-byte fp64 rows) of the graph "vertices
//     gathered by the same quarter warp for the same corner", so that those rows fall into distinct
//     bank groups.  Within a colour class ids ascend with the global id, which keeps the tile's
//     vertex list almost ascending (neighbouring lanes gather / RED neighbouring rows).
//   * REDUCE ORDER AND PADDING.  16 consecutive vertices of the reduce order are summed by one warp,
//     lane j reading slot start[j] + i: within every group of 16 the order and the optional one-slot
//     pads are searched so that the starts are distinct mod 16 (8-byte plane) and, within each half
//     of the group, distinct mod 8 (16-byte planes).
//   * SLOT POSITIONS inside a vertex's range: greedy + local search so that the 8 (16) stores of a
//     quarter (half) warp for one corner hit distinct slots mod 8 (mod 16).
// tools/smem_model.py replays the kernels' access pattern on these tables and counts wavefronts.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <thread>
#include <unordered_map>
#include <cstring>
#include <numeric>

#include "common.h"

namespace apl {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
const char* last_error_cstr() { return g_last_error.c_str(); }

static inline uint64_t spread21(uint64_t x) {
    x &= 0x1fffffULL;
    x = (x | (x << 32)) & 0x1f00000000ffffULL;
    x = (x | (x << 16)) & 0x1f0000ff0000ffULL;
    x = (x | (x << 8)) & 0x100f00f00f00f00fULL;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ULL;
    x = (x | (x << 2)) & 0x1249249249249249ULL;
    return x;
}

static void morton_order(int64_t n_cells, int64_t n_points, const int32_t* cells, const double* points,
                         std::vector<int64_t>& order) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int64_t v = 0; v < n_points; ++v)
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::min(lo[k], points[3 * v + k]);
            hi[k] = std::max(hi[k], points[3 * v + k]);
        }
    double ext = 0;
    for (int k = 0; k < 3; ++k) ext = std::max(ext, hi[k] - lo[k]);
    const double scale = ext > 0 ? (double)((1 << 21) - 1) / ext : 0.0;
    std::vector<std::pair<uint64_t, int64_t>> keyed((size_t)n_cells);
    bool sorted = true;
    uint64_t prev = 0;
    for (int64_t c = 0; c < n_cells; ++c) {
        uint64_t q[3];
        for (int k = 0; k < 3; ++k) {
            double s = 0;
            for (int a = 0; a < 4; ++a) s += points[3 * (int64_t)cells[4 * c + a] + k];
            q[k] = (uint64_t)((0.25 * s - lo[k]) * scale);
        }
        const uint64_t code = spread21(q[0]) | (spread21(q[1]) << 1) | (spread21(q[2]) << 2);
        keyed[(size_t)c] = {code, c};
        if (code < prev) sorted = false;
        prev = code;
    }
    if (!sorted) std::stable_sort(keyed.begin(), keyed.end());
    for (int64_t c = 0; c < n_cells; ++c) order[(size_t)c] = keyed[(size_t)c].second;
}

namespace {

// Deterministic per-tile random numbers for the local searches (the tables must not depend on the
// machine or the run).
struct Lcg {
    uint32_t s;
    explicit Lcg(uint32_t seed) : s(seed * 2654435761u + 12345u) {}
    uint32_t next() {
        s = s * 1664525u + 1013904223u;
        return s >> 8;
    }
    int below(int n) { return (int)(next() % (uint32_t)n); }
};

constexpr int kGroupsQ = (kTileTets / 8) * 4;    // (quarter warp, corner) groups of a tile
constexpr int kGroupsH = (kTileTets / 16) * 4;   // (half warp, corner) groups
constexpr int kSlotCap = kSlotsAlloc;
static_assert(kSlotCap < 4096, "tile_voff keeps the slot offset in 12 bits");

// Occupancy of the bank groups by one warp-wide access: c[r] lanes (distinct addresses) fall into bank
// group r, the access takes m = max c[r] wavefronts, nmax bank groups are that full.
struct Banks {
    uint8_t c[16];
    uint8_t m, nmax;
    void clear() { memset(this, 0, sizeof(*this)); }
    void rescan(int R) {
        m = 0; nmax = 0;
        for (int r = 0; r < R; ++r) {
            if (c[r] > m) { m = c[r]; nmax = 1; }
            else if (c[r] == m) ++nmax;
        }
    }
    // change of m when one lane moves from bank group r1 to r2
    int delta(int r1, int r2) const {
        if (r1 == r2) return 0;
        if (c[r2] + 1 > m) return 1;
        if (c[r1] == m && nmax == 1 && c[r2] + 1 < m) return -1;
        return 0;
    }
    void apply(int r1, int r2, int R) {
        --c[r1]; ++c[r2];
        rescan(R);
    }
};

// Scratch of one tile (allocated once per build_tiles call).
struct TileScratch {
    // global id -> provisional local id
    static constexpr int kHash = 512;
    static_assert(kHash >= 2 * kTileVerts, "hash table too small");
    int32_t hkey[kHash];
    uint8_t hval[kHash];
    // colouring
    uint8_t gmem[kGroupsQ][8];
    uint8_t gsz[kGroupsQ];
    Banks gb[kGroupsQ];
    int vptr[kTileVerts + 2];
    int vgrp[4 * kTileTets];
    int mark[kGroupsQ];
    // slots
    Banks bq[kGroupsQ], bh[kGroupsH];
    int16_t owner[kSlotCap];
    int16_t slot_of[4 * kTileTets];
};

// ---- 1. local vertex ids: balanced colouring of the gather-conflict graph -----------------------
// tl[t * nc + a]: provisional local id (first-touch order) of corner a of the tile's t-th tet (nc = 4).  Vertices
// read by the same quarter warp for the same corner should differ in (id mod M).  new_id[provisional] is a
// bijection onto [0, nv) with id = 32 w + M k + colour: w = the vertex's window of 32 in ascending global id,
// k ascending with the global id inside a (window, colour) class.
void color_local_ids(TileScratch& ws, int nt, int nc, int nv, const uint8_t* tl, const int32_t* gid, int M,
                     uint32_t seed, int* new_id) {
    const int nq = (nt + 7) / 8, ng = nq * nc;
    auto& gmem = ws.gmem;
    auto& gsz = ws.gsz;
    auto& gb = ws.gb;
    int* vptr = ws.vptr;
    int* vgrp = ws.vgrp;
    int col[kTileVerts];
    for (int v = 0; v <= nv + 1; ++v) vptr[v] = 0;
    for (int q = 0; q < nq; ++q)
        for (int a = 0; a < nc; ++a) {
            const int g = nc * q + a;
            gb[g].clear();
            ws.mark[g] = -1;
            int n = 0;
            for (int t = 8 * q; t < std::min(8 * q + 8, nt); ++t) {
                const uint8_t v = tl[t * nc + a];
                bool seen = false;
                for (int k = 0; k < n; ++k) seen |= (gmem[g][k] == v);
                if (!seen) gmem[g][n++] = v;
            }
            gsz[g] = (uint8_t)n;
            for (int k = 0; k < n; ++k) ++vptr[gmem[g][k] + 2];
        }
    for (int v = 0; v < nv; ++v) vptr[v + 2] += vptr[v + 1];
    for (int g = 0; g < ng; ++g)
        for (int k = 0; k < gsz[g]; ++k) vgrp[vptr[gmem[g][k] + 1]++] = g;
    // the groups of v are now vgrp[vptr[v] .. vptr[v+1])
    // Windows of 32 vertices in ascending global id: lanes 32 w .. 32 w + 31 of the gather / flush loops get
    // exactly the vertices of window w (the same global-memory lines per warp-wide access as an ascending
    // list), and the colours are balanced inside every window: id = 32 w + M k + colour.
    constexpr int kWin = 32, kMaxWin = (kTileVerts + kWin - 1) / kWin;
    int win[kTileVerts];
    {
        int by_gid[kTileVerts];
        for (int v = 0; v < nv; ++v) by_gid[v] = v;
        std::sort(by_gid, by_gid + nv, [&](int a, int b) { return gid[a] < gid[b]; });
        for (int r = 0; r < nv; ++r) win[by_gid[r]] = r / kWin;
    }
    int cap[kMaxWin][8], used[kMaxWin][8];
    for (int w = 0; w < kMaxWin; ++w) {
        const int m = std::max(0, std::min(kWin, nv - w * kWin));   // vertices in this window
        for (int c = 0; c < M; ++c) {
            cap[w][c] = c < m ? (m - c + M - 1) / M : 0;
            used[w][c] = 0;
        }
    }
    int order[kTileVerts];
    for (int v = 0; v < nv; ++v) order[v] = v;
    std::stable_sort(order, order + nv,
                     [&](int a, int b) { return vptr[a + 1] - vptr[a] > vptr[b + 1] - vptr[b]; });
    for (int i = 0; i < nv; ++i) {
        const int v = order[i];
        int best = -1, best_inc = 1 << 30, best_sec = 1 << 30, best_fill = 1 << 30;
        for (int c = 0; c < M; ++c) {
            if (used[win[v]][c] >= cap[win[v]][c]) continue;
            int inc = 0, sec = 0;
            for (int e = vptr[v]; e < vptr[v + 1]; ++e) {
                const Banks& b = gb[vgrp[e]];
                inc += (b.c[c] + 1 > b.m) && b.m >= 1;   // one more wavefront for this group
                sec += b.c[c];
            }
            const int fill = used[win[v]][c] - cap[win[v]][c];
            if (inc < best_inc || (inc == best_inc && (sec < best_sec || (sec == best_sec && fill < best_fill)))) {
                best = c; best_inc = inc; best_sec = sec; best_fill = fill;
            }
        }
        col[v] = best;
        ++used[win[v]][best];
        for (int e = vptr[v]; e < vptr[v + 1]; ++e) {
            Banks& b = gb[vgrp[e]];
            ++b.c[best];
            b.rescan(M);
        }
    }
    // local search: swap the colours of two vertices of one window (class sizes are preserved) when the
    // number of wavefronts does not grow
    Lcg rng(seed);
    int stamp = 0;
    auto try_swap = [&](int v, int w) {
        const int a = col[v], b = col[w];
        ++stamp;
        for (int e = vptr[w]; e < vptr[w + 1]; ++e) ws.mark[vgrp[e]] = stamp;
        int d = 0;
        for (int e = vptr[v]; e < vptr[v + 1]; ++e) {
            const int g = vgrp[e];
            if (ws.mark[g] == stamp) ws.mark[g] = -stamp - 1;   // contains both: unchanged
            else d += gb[g].delta(a, b);
        }
        for (int e = vptr[w]; e < vptr[w + 1]; ++e) {
            const int g = vgrp[e];
            if (ws.mark[g] == stamp) d += gb[g].delta(b, a);
        }
        if (d > 0) return false;
        for (int e = vptr[w]; e < vptr[w + 1]; ++e) {
            const int g = vgrp[e];
            if (ws.mark[g] == stamp) gb[g].apply(b, a, M);
        }
        for (int e = vptr[v]; e < vptr[v + 1]; ++e) {
            const int g = vgrp[e];
            if (ws.mark[g] != -stamp - 1) gb[g].apply(a, b, M);
        }
        col[v] = b; col[w] = a;
        return true;
    };
    for (int pass = 0; pass < 1; ++pass) {
        int nbad = 0;
        for (int g = 0; g < ng; ++g) {
            if (gb[g].m <= 1) continue;
            ++nbad;
            for (int i = 0; i < gsz[g] && gb[g].m > 1; ++i) {
                const int v = gmem[g][i];
                if (gb[g].c[col[v]] < gb[g].m) continue;
                for (int tries = 0; tries < 24; ++tries) {
                    const int w = rng.below(nv);
                    if (win[w] == win[v] && col[w] != col[v] && try_swap(v, w)) break;
                }
            }
        }
        if (nbad == 0) break;
    }
    // ids: inside a window the colour classes are interleaved, ascending global id inside a class
    for (int w = 0; w * kWin < nv; ++w)
        for (int c = 0; c < M; ++c) {
            int members[kWin], n = 0;
            for (int v = 0; v < nv; ++v)
                if (win[v] == w && col[v] == c) members[n++] = v;
            std::sort(members, members + n, [&](int a, int b) { return gid[a] < gid[b]; });
            for (int k = 0; k < n; ++k) new_id[members[k]] = w * kWin + M * k + c;
        }
}

// ---- 2. reduce order: conflict-free slot-range starts within every group of 16 vertices ------------
// cnt[0..n): valences of the group's vertices (about descending).  Finds an order and per-vertex pads
// (0..kMaxPad unused slots after the range, at most `pads_left` in total) such that the n range starts are
// distinct mod 16 (8-byte plane; skipped when !mod16) and distinct mod 8 within positions 0..7 and 8..15
// (16-byte planes).  Returns false when the bounded search fails.
constexpr int kMaxPad = 3;
struct GroupSearch {
    int n, budget, target;
    bool mod16;
    int cnt[16];
    bool taken[16];
    int order[16], pad[16];
    bool dfs(int j, int s, unsigned used16, unsigned used8, int pads_left) {
        if (j == n) return true;
        if (--budget < 0) return false;
        if (j == 8) used8 = 0;
        if (mod16 && (used16 >> (s & 15) & 1u)) return false;
        if (used8 >> (s & 7) & 1u) return false;
        used16 |= 1u << (s & 15);
        used8 |= 1u << (s & 7);
        int last = -1;
        for (int i = 0; i < n; ++i) {
            if (taken[i] || cnt[i] == last) continue;   // equal valences are interchangeable
            last = cnt[i];
            taken[i] = true;
            order[j] = i;
            // a constant odd stride never collides: first the pad that reaches the group's target stride
            const int want = std::min(std::max(target - cnt[i], 0), kMaxPad);
            for (int k = -1; k <= kMaxPad; ++k) {
                const int pd = k < 0 ? want : k;
                if ((k >= 0 && pd == want) || pd > pads_left) continue;
                pad[j] = pd;
                if (dfs(j + 1, s + cnt[i] + pd, used16, used8, pads_left - pd)) return true;
                if (budget < 0) break;
            }
            taken[i] = false;
            if (budget < 0) return false;
        }
        return false;
    }
};

// ---- 3. slot positions -------------------------------------------------------------------------------
// Bank group of a slot in a 16-byte plane = slot mod 8 per quarter warp; mod 16 per half warp for the
// 8-byte tail plane of fp32 slots.  Greedy, then local search over swaps of two slots of the same vertex
// (all slots of a vertex are equivalent for the reduction; pad slots are never used).
void place_slots(TileScratch& ws, int nt, int nc, const uint8_t* tl, const int* cnt, const int* off, int n_slots,
                 bool f64, uint32_t seed) {
    auto& bq = ws.bq;
    auto& bh = ws.bh;
    int16_t* owner = ws.owner;
    int16_t* slot_of = ws.slot_of;
    const int nq = ((nt + 7) / 8) * nc, nh = ((nt + 15) / 16) * nc, nx = nc * nt;
    const bool half = !f64;
    for (int g = 0; g < nq; ++g) bq[g].clear();
    for (int g = 0; g < nh; ++g) bh[g].clear();
    for (int s = 0; s < n_slots; ++s) owner[s] = -1;
    // most constrained first: corners of low-valence vertices have the fewest slots to choose from
    int16_t corder[4 * kTileTets];
    {
        int bucket[4 * kTileTets + 2];
        const int nb = nx + 2;
        for (int i = 0; i < nb; ++i) bucket[i] = 0;
        for (int x = 0; x < nx; ++x) ++bucket[cnt[tl[x]] + 1];
        for (int i = 1; i < nb; ++i) bucket[i] += bucket[i - 1];
        for (int x = 0; x < nx; ++x) corder[bucket[cnt[tl[x]]]++] = (int16_t)x;
    }
    for (int i = 0; i < nx; ++i) {
        const int x = corder[i], t = x / nc, a = x % nc;
        const int l = tl[x];
        Banks& q = bq[(t >> 3) * nc + a];
        Banks& h = bh[(t >> 4) * nc + a];
        int best = -1, best_cost = 1 << 30;
        for (int k = 0; k < cnt[l]; ++k) {
            const int s = off[l] + k;
            if (owner[s] >= 0) continue;
            const int cost = 4 * q.c[s & 7] + (half ? h.c[s & 15] : 0);
            if (cost < best_cost) {
                best_cost = cost;
                best = s;
                if (cost == 0) break;
            }
        }
        owner[best] = (int16_t)x;
        slot_of[x] = (int16_t)best;
        ++q.c[best & 7];
        ++h.c[best & 15];
    }
    for (int g = 0; g < nq; ++g) bq[g].rescan(8);
    for (int g = 0; g < nh; ++g) bh[g].rescan(16);
    auto gq_of = [nc](int x) { return ((x / nc) >> 3) * nc + (x % nc); };
    auto gh_of = [nc](int x) { return ((x / nc) >> 4) * nc + (x % nc); };
    // wavefront change when corner x moves s -> s2 and the owner y of s2 (if any) moves s2 -> s
    auto swap_delta = [&](int x, int y, int s, int s2) {
        int d = 0;
        const int qx = gq_of(x), hx = gh_of(x);
        if (y < 0) {
            d += bq[qx].delta(s & 7, s2 & 7);
            if (half) d += bh[hx].delta(s & 15, s2 & 15);
            return d;
        }
        const int qy = gq_of(y), hy = gh_of(y);
        if (qx != qy) d += bq[qx].delta(s & 7, s2 & 7) + bq[qy].delta(s2 & 7, s & 7);
        if (half && hx != hy) d += bh[hx].delta(s & 15, s2 & 15) + bh[hy].delta(s2 & 15, s & 15);
        return d;
    };
    auto do_swap = [&](int x, int y, int s, int s2) {
        bq[gq_of(x)].apply(s & 7, s2 & 7, 8);
        bh[gh_of(x)].apply(s & 15, s2 & 15, 16);
        if (y >= 0) {
            bq[gq_of(y)].apply(s2 & 7, s & 7, 8);
            bh[gh_of(y)].apply(s2 & 15, s & 15, 16);
            slot_of[y] = (int16_t)s;
        }
        owner[s2] = (int16_t)x;
        owner[s] = (int16_t)y;
        slot_of[x] = (int16_t)s2;
    };
    Lcg rng(seed ^ 0x9e3779b9u);
    auto improve = [&](int x) {
        const int l = tl[x];
        const int len = cnt[l];
        if (len < 2) return;
        const int s = slot_of[x];
        int best_s2 = -1, best_delta = 1;
        const int r0 = rng.below(len);
        for (int k = 0; k < len; ++k) {
            const int s2 = off[l] + (r0 + k) % len;
            if (s2 == s) continue;
            const int d = swap_delta(x, owner[s2], s, s2);
            if (d < best_delta) {
                best_delta = d;
                best_s2 = s2;
                if (d < 0) break;
            }
        }
        if (best_s2 >= 0) do_swap(x, owner[best_s2], s, best_s2);
    };
    for (int pass = 0; pass < 2; ++pass) {
        int nbad = 0;
        for (int g = 0; g < nq; ++g) {
            if (bq[g].m <= 1) continue;
            ++nbad;
            const int a = g % nc, t0 = (g / nc) * 8, t1 = std::min(t0 + 8, nt);
            for (int t = t0; t < t1; ++t)
                if (bq[g].m > 1 && bq[g].c[slot_of[nc * t + a] & 7] == bq[g].m) improve(nc * t + a);
        }
        for (int g = 0; g < nh && half; ++g) {
            if (bh[g].m <= 1) continue;
            ++nbad;
            const int a = g % nc, t0 = (g / nc) * 16, t1 = std::min(t0 + 16, nt);
            for (int t = t0; t < t1; ++t)
                if (bh[g].m > 1 && bh[g].c[slot_of[nc * t + a] & 15] == bh[g].m) improve(nc * t + a);
        }
        if (nbad == 0) break;
    }
}

// ---- a closed tile: local ids, reduce order, slots -> tables ------------------------------------------
// One tet per consumer thread (nc = 4 corners, conn / slots rows of 4 entries).
struct TileRange {
    int64_t item_start;    // first item (tet, or pair slot) of the tile
    int32_t ni, nv;        // items, distinct vertices
    int64_t first_touch;   // into `touched`: the tile's vertices in first-touch order
    int64_t vert_start, voff_start;
};

void finish_tile(TileScratch& ws, HostTables& out, int64_t tile, const TileRange& r, int nc, int row, int tets_per_item,
                 uint8_t* tl /* ni x nc provisional ids, rewritten to final ids */, const int32_t* verts, bool f64) {
    const int ni = r.ni, nv = r.nv;
    const int gather_mod = f64 ? 4 : 8;   // nodal rows in shared memory are 4 scalars: 16 B (fp32) / 32 B (fp64)
    // 1. final local ids
    int new_id[kTileVerts];
    color_local_ids(ws, ni, nc, nv, tl, verts, gather_mod, (uint32_t)tile, new_id);
    int cnt[kTileVerts];
    for (int l = 0; l < nv; ++l) {
        out.tile_verts[(size_t)r.vert_start + new_id[l]] = verts[l];
        cnt[l] = 0;
    }
    for (int x = 0; x < ni * nc; ++x) {
        tl[x] = (uint8_t)new_id[tl[x]];
        ++cnt[tl[x]];
    }
    // 2. reduce order: decreasing valence (balanced trip counts within a warp), then per group of 16
    //    the order / pads that make the range starts conflict-free
    int perm[kTileVerts];
    for (int l = 0; l < nv; ++l) perm[l] = l;
    std::stable_sort(perm, perm + nv, [&](int a, int b) { return cnt[a] > cnt[b]; });
    int start[kTileVerts + 1], padv[kTileVerts], off[kTileVerts];
    start[0] = 0;
    // slot capacity of the kernels: every corner of a full tile plus kSlotPads pad slots (common.h: kSlotsAlloc)
    const int slot_cap = kSlotsAlloc;
    const int pads_total = slot_cap - nc * ni;   // >= kSlotPads
    int pads_left = pads_total;
    for (int g0 = 0; g0 < nv; g0 += 16) {
        GroupSearch gs;
        gs.n = std::min(16, nv - g0);
        // the pad budget is shared about in proportion: every later vertex keeps 70 % of its share of the tile's
        // pads (the first groups have the longest ranges and gain most from conflict-free starts)
        const int reserve = (int)((int64_t)(nv - g0 - gs.n) * pads_total * 7 / (10 * std::max(nv, 1)));
        const int allowed = std::max(0, pads_left - reserve);
        int grp[16];
        for (int i = 0; i < gs.n; ++i) {
            grp[i] = perm[g0 + i];
            gs.cnt[i] = cnt[grp[i]];
        }
        gs.target = gs.cnt[0] | 1;
        bool found = false;
        for (int attempt = 0; attempt < 2 && !found; ++attempt) {   // all planes, then the 16-byte planes only
            if (attempt == 1 && f64) break;
            gs.mod16 = !f64 && attempt == 0;
            gs.budget = 1500;
            for (int i = 0; i < gs.n; ++i) gs.taken[i] = false;
            found = gs.dfs(0, start[g0], 0u, 0u, allowed);
        }
        if (found) {
            for (int j = 0; j < gs.n; ++j) {
                perm[g0 + j] = grp[gs.order[j]];
                padv[g0 + j] = gs.pad[j];
            }
        } else {
            int left = allowed;
            for (int j = 0; j < gs.n; ++j) {   // odd strides while the budget lasts
                padv[g0 + j] = ((cnt[grp[j]] & 1) || left == 0) ? 0 : 1;
                left -= padv[g0 + j];
            }
        }
        for (int j = 0; j < gs.n; ++j) {
            start[g0 + j + 1] = start[g0 + j] + cnt[perm[g0 + j]] + padv[g0 + j];
            pads_left -= padv[g0 + j];
        }
    }
    for (int t = 0; t < nv; ++t) {
        off[perm[t]] = start[t];
        out.tile_vperm[(size_t)r.vert_start + t] = (uint8_t)perm[t];
        out.tile_voff[(size_t)r.voff_start + t] = (uint16_t)(start[t] | (padv[t] << 12));
    }
    out.tile_voff[(size_t)r.voff_start + nv] = (uint16_t)start[nv];
    int32_t* hdr = out.tiles.data() + 6 * tile;
    hdr[0] = (int32_t)(r.item_start * tets_per_item);
    hdr[1] = ni * tets_per_item;
    hdr[2] = (int32_t)r.vert_start;
    hdr[3] = nv;
    hdr[4] = (int32_t)r.voff_start;
    hdr[5] = start[nv];
    // 3. slot positions
    place_slots(ws, ni, nc, tl, cnt, off, start[nv], f64, (uint32_t)tile);
    for (int t = 0; t < ni; ++t)
        for (int a = 0; a < nc; ++a) {
            out.conn[(size_t)(r.item_start + t) * row + a] = tl[t * nc + a];
            out.slots[(size_t)(r.item_start + t) * row + a] = (uint16_t)ws.slot_of[nc * t + a];
        }
}

// global id -> provisional local id of the tile's vertices (hash table in the scratch)
void hash_tile_verts(TileScratch& ws, const int32_t* verts, int nv) {
    for (int i = 0; i < TileScratch::kHash; ++i) ws.hkey[i] = -1;
    for (int l = 0; l < nv; ++l) {
        int h = (int)(((uint32_t)verts[l] * 2654435761u) >> 23) & (TileScratch::kHash - 1);
        while (ws.hkey[h] >= 0) h = (h + 1) & (TileScratch::kHash - 1);
        ws.hkey[h] = verts[l];
        ws.hval[h] = (uint8_t)l;
    }
}
inline uint8_t lookup_tile_vert(const TileScratch& ws, int32_t v) {
    int h = (int)(((uint32_t)v * 2654435761u) >> 23) & (TileScratch::kHash - 1);
    while (ws.hkey[h] != v) h = (h + 1) & (TileScratch::kHash - 1);
    return ws.hval[h];
}

template <typename F>
void for_tiles_parallel(int64_t n_tiles, F&& body) {
    int n_threads = (int)std::thread::hardware_concurrency();
    if (const char* e = getenv("APL_TILING_THREADS")) n_threads = atoi(e);
    n_threads = std::max(1, std::min(n_threads, 32));
    if (n_tiles < 64) n_threads = 1;
    if (n_threads == 1) {
        std::vector<TileScratch> ws(1);
        for (int64_t t = 0; t < n_tiles; ++t) body(ws[0], t);
        return;
    }
    std::atomic<int64_t> next{0};
    std::vector<std::thread> pool;
    for (int i = 0; i < n_threads; ++i)
        pool.emplace_back([&] {
            std::vector<TileScratch> ws(1);
            for (;;) {
                const int64_t t0 = next.fetch_add(16);
                if (t0 >= n_tiles) break;
                for (int64_t t = t0; t < std::min(t0 + 16, n_tiles); ++t) body(ws[0], t);
            }
        });
    for (auto& th : pool) th.join();
}

}  // namespace


// ---- pass 1: tile boundaries ---------------------------------------------------------------------------
// The packed tets are cut into tiles chunk by chunk (chunks of kCutChunk packed positions, cut independently
// and in parallel; a tile never spans two chunks).  A tile closes when it is full.  Vertex budget: tiles must
// start at multiples of 4 tets (16-byte aligned byte-wide connectivity), so the budget is checked every 4 tets
// with room for the worst case of 16 new vertices in the next 4.
namespace {
constexpr int64_t kCutChunk = 1 << 17;
struct ChunkCut {
    std::vector<TileRange> ranges;   // first_touch relative to this chunk's `touched`; vert_start / voff_start unset
    std::vector<int32_t> touched;
};

template <typename CellAt>
void cut_chunk(int64_t begin, int64_t end, CellAt&& cell_at, ChunkCut& out) {
    constexpr int kH = 512;   // open-addressing set of the open tile's vertices (<= kTileVerts entries)
    static_assert(kH >= 2 * kTileVerts, "hash set too small");
    int32_t keys[kH];
    auto reset = [&] { for (int i = 0; i < kH; ++i) keys[i] = -1; };
    reset();
    out.ranges.reserve((size_t)((end - begin) / 200 + 4));
    out.touched.reserve((size_t)((end - begin) / 2 + 16));
    int64_t tile_start = begin, first = 0;
    auto close_tile = [&](int64_t tile_end) {
        const int nt = (int)(tile_end - tile_start);
        if (nt == 0) return;
        const int nv = (int)((int64_t)out.touched.size() - first);
        out.ranges.push_back({tile_start, nt, nv, first, 0, 0});
        first = (int64_t)out.touched.size();
        tile_start = tile_end;
        reset();
    };
    for (int64_t pos = begin; pos < end; ++pos) {
        const int32_t* c = cell_at(pos);
        const int64_t in_tile = pos - tile_start;
        const int nv_open = (int)((int64_t)out.touched.size() - first);
        if (in_tile == kTileTets || (in_tile % 4 == 0 && in_tile > 0 && nv_open + 16 > kTileVerts)) close_tile(pos);
        for (int a = 0; a < 4; ++a) {
            const int32_t v = c[a];
            int h = (int)(((uint32_t)v * 2654435761u) >> 23) & (kH - 1);
            while (keys[h] >= 0 && keys[h] != v) h = (h + 1) & (kH - 1);
            if (keys[h] < 0) {
                keys[h] = v;
                out.touched.push_back(v);
            }
        }
    }
    close_tile(end);
}

template <typename F>
void parallel_for_chunks(int64_t n, F&& body) {
    int n_threads = (int)std::thread::hardware_concurrency();
    if (const char* e = getenv("APL_TILING_THREADS")) n_threads = atoi(e);
    n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(n_threads, 32), n));
    if (n_threads == 1) {
        for (int64_t i = 0; i < n; ++i) body(i);
        return;
    }
    std::atomic<int64_t> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t)
        pool.emplace_back([&] {
            for (;;) {
                const int64_t i = next.fetch_add(1);
                if (i >= n) break;
                body(i);
            }
        });
    for (auto& th : pool) th.join();
}
}  // namespace

// out.order must hold the packed order (packed position -> caller's cell).  `cells` is either the caller's
// array (cells_packed = false: the tet at packed position pos is cells[order[pos]]) or already permuted into
// packed order (cells_packed = true: the device-side setup sorts the connectivity on the GPU).
static int build_tiles_ordered(int64_t n_cells, int64_t n_points, const int32_t* cells, bool cells_packed, bool f64,
                               HostTables& out) {
    auto cell_at = [&](int64_t pos) { return cells + 4 * (cells_packed ? pos : out.order[(size_t)pos]); };
    out.conn.resize((size_t)n_cells * 4);
    out.slots.resize((size_t)n_cells * 4);

    // ---- pass 1 (parallel over chunks): tile boundaries and the distinct vertices of every tile in first-touch
    //      order; then (serial, per tile) where each tile's tables start (vertex lists at multiples of 16 entries,
    //      offset lists at multiples of 8 entries: every per-tile table is 16-byte aligned for bulk copies)
    const int64_t n_chunks = (n_cells + kCutChunk - 1) / kCutChunk;
    std::vector<ChunkCut> cuts((size_t)n_chunks);
    parallel_for_chunks(n_chunks, [&](int64_t i) {
        cut_chunk(i * kCutChunk, std::min(n_cells, (i + 1) * kCutChunk), cell_at, cuts[(size_t)i]);
    });
    std::vector<TileRange> ranges;
    std::vector<int64_t> chunk_touch_start((size_t)n_chunks + 1, 0);
    {
        size_t n_ranges = 0;
        for (int64_t i = 0; i < n_chunks; ++i) {
            n_ranges += cuts[(size_t)i].ranges.size();
            chunk_touch_start[(size_t)i + 1] = chunk_touch_start[(size_t)i] + (int64_t)cuts[(size_t)i].touched.size();
        }
        ranges.reserve(n_ranges);
    }
    std::vector<int32_t> touched((size_t)chunk_touch_start[(size_t)n_chunks]);   // concatenated first-touch vertex lists
    parallel_for_chunks(n_chunks, [&](int64_t i) {
        const auto& t = cuts[(size_t)i].touched;
        if (!t.empty()) memcpy(touched.data() + chunk_touch_start[(size_t)i], t.data(), t.size() * sizeof(int32_t));
    });
    int64_t vert_end = 0, voff_end = 0;
    for (int64_t i = 0; i < n_chunks; ++i) {
        for (TileRange r : cuts[(size_t)i].ranges) {
            vert_end = (vert_end + 15) / 16 * 16;
            voff_end = (voff_end + 7) / 8 * 8;
            r.first_touch += chunk_touch_start[(size_t)i];
            r.vert_start = vert_end;
            r.voff_start = voff_end;
            vert_end += r.nv;
            voff_end += r.nv + 1;
            ranges.push_back(r);
        }
        ChunkCut().ranges.swap(cuts[(size_t)i].ranges);
        std::vector<int32_t>().swap(cuts[(size_t)i].touched);
    }
    // + padding so that 16-byte granular bulk copies of the last tile stay in bounds
    if (vert_end + 16 > (int64_t)INT32_MAX || voff_end + 16 > (int64_t)INT32_MAX) {
        set_error("tile vertex table exceeds int32 range");
        return APL_ERR_INVALID;
    }
    out.tile_verts.assign((size_t)vert_end + 16, 0);
    out.tile_vperm.assign((size_t)vert_end + 16, 0);
    out.tile_voff.assign((size_t)voff_end + 16, 0);
    const int64_t n_tiles = (int64_t)ranges.size();
    out.tiles.assign((size_t)n_tiles * 6, 0);

    // ---- pass 2 (parallel over tiles): local ids, reduce order, slots
    for_tiles_parallel(n_tiles, [&](TileScratch& ws, int64_t tile) {
        const TileRange& r = ranges[(size_t)tile];
        const int32_t* verts = touched.data() + r.first_touch;
        uint8_t tl[kTileTets * 4];
        hash_tile_verts(ws, verts, r.nv);
        for (int t = 0; t < r.ni; ++t) {
            const int32_t* c = cell_at(r.item_start + t);
            for (int a = 0; a < 4; ++a) tl[4 * t + a] = lookup_tile_vert(ws, c[a]);
        }
        finish_tile(ws, out, tile, r, 4, 4, 1, tl, verts, f64);
    });
    return APL_OK;
}

int build_tiles(int64_t n_cells, int64_t n_points, const int32_t* cells, const double* points, int elem_bytes,
                HostTables& out) {
    if (n_cells < 0 || n_points <= 0 || (!cells && n_cells > 0)) {
        set_error("build_tiles: bad sizes");
        return APL_ERR_INVALID;
    }
    if (n_cells > (int64_t)INT32_MAX / 2) {
        set_error("n_cells exceeds the int32 range of the packed tables");
        return APL_ERR_INVALID;
    }
    for (int64_t i = 0; i < 4 * n_cells; ++i)
        if (cells[i] < 0 || cells[i] >= n_points) {
            set_error("cells[" + std::to_string(i / 4) + "] references vertex " + std::to_string(cells[i]) +
                      " outside [0, n_points)");
            return APL_ERR_MESH;
        }
    const bool f64 = elem_bytes == 8;
    out = HostTables();
    out.n_cells = n_cells;
    out.n_points = n_points;
    // Morton order of the cells
    out.order.resize((size_t)n_cells);
    if (points) morton_order(n_cells, n_points, cells, points, out.order);
    else std::iota(out.order.begin(), out.order.end(), (int64_t)0);
    return build_tiles_ordered(n_cells, n_points, cells, false, f64, out);
}

// Tiling of connectivity that is ALREADY in packed order (sorted on the device, setup.cu): out.order holds the
// permutation that produced it (packed position -> caller's cell).
int build_tiles_packed(int64_t n_cells, int64_t n_points, const int32_t* packed_cells, std::vector<int64_t>&& order,
                       int elem_bytes, HostTables& out) {
    if (n_cells < 0 || n_points <= 0 || (int64_t)order.size() != n_cells) {
        set_error("build_tiles_packed: bad sizes");
        return APL_ERR_INVALID;
    }
    if (n_cells > (int64_t)INT32_MAX / 2) {
        set_error("n_cells exceeds the int32 range of the packed tables");
        return APL_ERR_INVALID;
    }
    out = HostTables();
    out.n_cells = n_cells;
    out.n_points = n_points;
    out.order = std::move(order);
    return build_tiles_ordered(n_cells, n_points, packed_cells, true, elem_bytes == 8, out);
}

}  // namespace apl
