// Host-side mesh packing: Morton ordering of the tets and greedy tiling.
//
// A tile is a run of <= 256 consecutive tets (in packed order) that together touch <= 256 distinct
// vertices.  For every tile we store
//   * the ascending list of the global vertex ids it touches           (tile_verts)
//   * the same local ids ordered by decreasing valence                 (tile_vperm)
//   * per tet, the tile-local id of each corner as one byte          (conn)
//   * per tet corner, a slot in [0, 4*n_tets): slots are grouped by local vertex, so the element
//     kernel can write every corner contribution to its own shared-memory slot (no atomics) and a
//     second phase sums each vertex's contiguous slot range                     (slots, tile_voff)
// This replaces the reference's per-tet global connectivity (`cells`, warp/fem/_base.py:59-74) and
// its 12 global atomics per tet per field (warp/fem/_base.py:288-289).
#include <algorithm>
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

int build_tiles(int64_t n_cells, int64_t n_points, const int32_t* cells, const double* points,
                HostTables& out) {
    if (n_cells < 0 || n_points <= 0 || (!cells && n_cells > 0)) {
        set_error("build_tiles: bad sizes");
        return APL_ERR_INVALID;
    }
    for (int64_t i = 0; i < 4 * n_cells; ++i)
        if (cells[i] < 0 || cells[i] >= n_points) {
            set_error("cells[" + std::to_string(i / 4) + "] references vertex " + std::to_string(cells[i]) +
                      " outside [0, n_points)");
            return APL_ERR_MESH;
        }
    out = HostTables();
    out.n_cells = n_cells;
    out.n_points = n_points;
    out.order.resize((size_t)n_cells);
    if (points) morton_order(n_cells, n_points, cells, points, out.order);
    else std::iota(out.order.begin(), out.order.end(), (int64_t)0);
    out.conn.resize((size_t)n_cells * 4);
    out.slots.resize((size_t)n_cells * 4);
    out.tiles.reserve((size_t)(n_cells / kTileTets + 1) * 6);
    out.tile_verts.reserve((size_t)(n_cells / 2 + 16));

    std::vector<int32_t> stamp((size_t)n_points, -1);  // tile that last touched the vertex
    std::vector<int32_t> lid((size_t)n_points, 0);     // its local id in that tile
    std::vector<int32_t> verts;                        // distinct vertices of the open tile (first-touch order)
    verts.reserve(kTileVerts);
    int32_t cnt[kTileVerts];
    int32_t off[kTileVerts + 1];
    int32_t perm[kTileVerts];
    int64_t tile_start = 0;
    int32_t tile_id = 0;

    // Closes the tile [tile_start, tile_end).  Local vertex ids ascend with the global id (coalesced
    // gathers and REDs); tile_vperm lists the local ids by DECREASING valence so that the per-vertex
    // slot reduction of the kernel has nearly uniform trip counts within a warp.  The vertex list
    // starts at a multiple of 16 entries and the offset list at a multiple of 8 entries, i.e. every
    // per-tile table is 16-byte aligned for bulk copies.
    auto close_tile = [&](int64_t tile_end) {
        const int nt = (int)(tile_end - tile_start);
        if (nt == 0) return;
        const int nv = (int)verts.size();
        for (int l = 0; l < nv; ++l) {
            lid[(size_t)verts[l]] = l;
            cnt[l] = 0;
            perm[l] = l;
        }
        for (int64_t pos = tile_start; pos < tile_end; ++pos) {
            const int32_t* c = cells + 4 * out.order[(size_t)pos];
            for (int a = 0; a < 4; ++a) ++cnt[lid[(size_t)c[a]]];
        }
        std::sort(perm, perm + nv, [&](int a, int b) { return verts[a] < verts[b]; });
        while (out.tile_verts.size() % 16) {
            out.tile_verts.push_back(0);
            out.tile_vperm.push_back(0);
        }
        while (out.tile_voff.size() % 8) out.tile_voff.push_back(0);
        const int32_t vert_start = (int32_t)out.tile_verts.size();
        const int32_t voff_start = (int32_t)out.tile_voff.size();
        int32_t cnt_new[kTileVerts];
        for (int l = 0; l < nv; ++l) {
            const int old = perm[l];
            lid[(size_t)verts[old]] = l;
            out.tile_verts.push_back(verts[old]);
            cnt_new[l] = cnt[old];
        }
        // reduce order: local ids by decreasing valence
        for (int l = 0; l < nv; ++l) perm[l] = l;
        std::stable_sort(perm, perm + nv, [&](int a, int b) { return cnt_new[a] > cnt_new[b]; });
        for (int l = 0; l < nv; ++l) out.tile_vperm.push_back((uint8_t)perm[l]);
        // Slot ranges are laid out in REDUCE order and padded to odd lengths: reduce thread t reads
        // start[t] + i, so neighbouring lanes are an odd number of slots apart and their 16-byte
        // (and 8-byte) reads fall into distinct banks.  Bit 15 of an entry flags a padded range.
        int32_t start[kTileVerts + 1];
        start[0] = 0;
        for (int t = 0; t < nv; ++t) {
            const int c = cnt_new[perm[t]];
            const int padded = c | 1;
            start[t + 1] = start[t] + padded;
            off[perm[t]] = start[t];
            out.tile_voff.push_back((uint16_t)(start[t] | ((padded != c) ? 0x8000 : 0)));
        }
        out.tile_voff.push_back((uint16_t)start[nv]);
        out.tiles.push_back((int32_t)tile_start);
        out.tiles.push_back(nt);
        out.tiles.push_back(vert_start);
        out.tiles.push_back(nv);
        out.tiles.push_back(voff_start);
        out.tiles.push_back(start[nv]);
        // Positions inside a vertex's range are free: choose them so that the 32 stores of one
        // warp-wide STS (same corner of 32 consecutive tets) spread over the banks.  Bank of a slot
        // in a 16-byte plane = slot mod 8 per quarter warp; mod 16 per half warp for the 8-byte plane.
        uint8_t taken[4 * kTileTets + kTileVerts];
        memset(taken, 0, sizeof(taken));
        for (int64_t w0 = tile_start; w0 < tile_end; w0 += 32) {
            const int64_t w1 = std::min(w0 + 32, tile_end);
            for (int a = 0; a < 4; ++a) {
                int used_h[2][16], used_q[4][8];
                memset(used_h, 0, sizeof(used_h));
                memset(used_q, 0, sizeof(used_q));
                for (int64_t pos = w0; pos < w1; ++pos) {
                    const int32_t* c = cells + 4 * out.order[(size_t)pos];
                    const int l = lid[(size_t)c[a]];
                    const int half = (int)((pos - w0) >> 4), quarter = (int)((pos - w0) >> 3);
                    const int base = off[l], n = cnt_new[l];
                    int best = -1, best_cost = 1 << 30;
                    for (int k = 0; k < n; ++k) {
                        if (taken[base + k]) continue;
                        const int cost = 4 * used_q[quarter][(base + k) & 7] + used_h[half][(base + k) & 15];
                        if (cost < best_cost) {
                            best_cost = cost;
                            best = base + k;
                            if (cost == 0) break;
                        }
                    }
                    taken[best] = 1;
                    ++used_h[half][best & 15];
                    ++used_q[quarter][best & 7];
                    out.conn[(size_t)pos * 4 + a] = (uint8_t)l;
                    out.slots[(size_t)pos * 4 + a] = (uint16_t)best;
                }
            }
        }
        verts.clear();
        tile_start = tile_end;
        ++tile_id;
    };

    if (n_cells > (int64_t)INT32_MAX) {
        set_error("n_cells exceeds int32 range");
        return APL_ERR_INVALID;
    }
    for (int64_t pos = 0; pos < n_cells; ++pos) {
        const int32_t* c = cells + 4 * out.order[(size_t)pos];
        // A tile closes when it is full.  Vertex budget: tiles must start at multiples of 4 tets
        // (16-byte aligned byte-wide connectivity), so the budget is checked every 4 tets with room
        // for the worst case of 16 new vertices in the next 4.
        const int64_t in_tile = pos - tile_start;
        if (in_tile == kTileTets || (in_tile % 4 == 0 && in_tile > 0 && (int)verts.size() + 16 > kTileVerts)) {
            close_tile(pos);
        }
        for (int a = 0; a < 4; ++a)
            if (stamp[(size_t)c[a]] != tile_id) {
                stamp[(size_t)c[a]] = tile_id;
                verts.push_back(c[a]);
            }
    }
    close_tile(n_cells);
    // pad the tables so that 16-byte granular bulk copies of the last tile stay in bounds
    for (int k = 0; k < 16; ++k) {
        out.tile_verts.push_back(0);
        out.tile_vperm.push_back(0);
    }
    for (int k = 0; k < 16; ++k) out.tile_voff.push_back(0);
    if (out.tile_verts.size() > (size_t)INT32_MAX || out.tile_voff.size() > (size_t)INT32_MAX) {
        set_error("tile vertex table exceeds int32 range");
        return APL_ERR_INVALID;
    }
    return APL_OK;
}

}  // namespace apl
