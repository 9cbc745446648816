// Shared declarations of the apple_b200 native library (not part of the public ABI).
#pragma once

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/apple_b200.h"

namespace apl {

#ifndef APL_TILE_TETS
#define APL_TILE_TETS 256
#endif
constexpr int kTileTets = APL_TILE_TETS;   // tets per tile == consumer threads per CTA
constexpr int kTileVerts = APL_TILE_TETS * 3 / 4;  // max distinct vertices per tile (uint8 local ids, <= 256)
// Reduction slots of a tile: one per tet corner plus up to kSlotPads unused pad slots that keep the slot-range
// starts of a reduce group in distinct bank groups (tiling.cpp); the kernels allocate kSlotsAlloc slots per buffer.
#ifndef APL_SLOT_PADS
#define APL_SLOT_PADS (APL_TILE_TETS * 5 / 8)
#endif
constexpr int kSlotPads = APL_SLOT_PADS;
constexpr int kSlotsAlloc = 4 * APL_TILE_TETS + kSlotPads;

void set_error(const std::string& msg);

// Host-side packed mesh (built by tiling.cpp, uploaded by capi.cu).
struct HostTables {
    int64_t n_cells = 0, n_points = 0;
    std::vector<int32_t> tiles;        // (n_tiles,6): tet_start, n_tets, vert_start, n_verts, voff_start, n_slots
    std::vector<int64_t> order;        // packed tet position -> caller's cell index
    std::vector<uint8_t> conn;         // tile-local vertex ids, (n_cells,4)
    std::vector<uint16_t> slots;       // reduction slots, same shape
    std::vector<int32_t> tile_verts;   // global vertex id of every tile-local id (tile start padded to x16)
    std::vector<uint8_t> tile_vperm;   // same indexing: local ids in reduce order (about decreasing valence)
    std::vector<uint16_t> tile_voff;   // per tile n_verts+1 slot offsets, starting at voff_start (x8)
    int64_t n_tiles() const { return (int64_t)tiles.size() / 6; }
    int64_t n_packed() const { return (int64_t)order.size(); }   // packed tet positions
};

// cells: (n_cells,4) int32.  points: (n_points,3) double or nullptr.  elem_bytes: 4 (fp32) or 8 (fp64), the
// scalar size of the nodal rows the kernels keep in shared memory (decides which local ids collide).
// Returns APL_OK or error code.
int build_tiles(int64_t n_cells, int64_t n_points, const int32_t* cells, const double* points, int elem_bytes,
                HostTables& out);

// Tiling of connectivity ALREADY in packed (Morton) order -- the device-side setup (setup.cu) sorts the cells on
// the GPU: packed_cells[pos] is the tet at packed position pos, order[pos] the caller's index of that tet.
int build_tiles_packed(int64_t n_cells, int64_t n_points, const int32_t* packed_cells, std::vector<int64_t>&& order,
                       int elem_bytes, HostTables& out);

}  // namespace apl

struct apl_fem {
    int kind = 0, dtype = 0, device = -1;
    int nrec = 0;      // scalars per tet record
    int nplanes = 0;   // 16-byte planes per tet
    apl::HostTables host;
    std::vector<unsigned char> host_planes;  // packed static planes, kept only by host-only handles
    int64_t plane_stride = 0;  // in 16-byte vectors (n_cells rounded up)
    int64_t static_bytes = 0;
    // device tables
    void* d_tiles = nullptr;
    void* d_conn = nullptr;
    void* d_slots = nullptr;
    void* d_tile_verts = nullptr;
    void* d_tile_voff = nullptr;
    void* d_tile_vperm = nullptr;
    void* d_planes = nullptr;
    int32_t* d_order = nullptr;     // packed tet position -> caller's cell; uploaded on first use
    double* d_partials = nullptr;   // per-CTA scalar partials (2 per CTA)
    unsigned int* d_counter = nullptr;
    int max_grid = 0;
    int num_sms = 0;
    int64_t n_boundary_tiles = 0;   // tile headers [0, n_boundary_tiles) touch a flagged (shared) vertex
};

#define APL_CUDA_CHECK(expr)                                                                       \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            apl::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" +     \
                           __FILE__ + ":" + std::to_string(__LINE__) + ")");                       \
            return APL_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)
