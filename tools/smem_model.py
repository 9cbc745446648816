"""Shared-memory wavefront model of fem_pipe_kernel (fp32, grad+prod slots: NOUT=2, SS=6).

Reads the packed tables of a host-only handle (no GPU) and counts, for every warp-wide LDS/STS of the
consumer warps, the number of wavefronts = max over the 32 banks of distinct 4-byte words touched
(identical words are broadcast).  Prints wavefronts per 32 tets, split by access class, so that the
table layout (tiling.cpp) can be tuned offline and compared with ncu's
l1tex__data_pipe_lsu_wavefronts_mem_shared_op_{ld,st}.sum.

usage: python tools/smem_model.py [--n 58] [--tiles 400] [--quarter] [--layout pair]
"""
import argparse, ctypes, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))

KSLOTS = 4 * 256 + 192

def wavefronts(addr, nbytes, active=None, quarter=False):
    """addr: (..., 32) byte addresses per lane; nbytes in {1,2,4,8,16}.  Returns (...) wavefront counts."""
    addr = np.asarray(addr, dtype=np.int64)
    if active is None:
        active = np.ones(addr.shape, bool)
    nw = max(nbytes // 4, 1)
    words = (addr // 4)[..., None] + np.arange(nw)           # (..., 32, nw)
    act = np.broadcast_to(active[..., None], words.shape)
    def count(words, act):
        lead = words.shape[:-2]
        w = words.reshape(*lead, -1); a = act.reshape(*lead, -1)
        w = np.where(a, w, -1)
        w = np.sort(w, axis=-1)
        first = np.ones(w.shape, bool); first[..., 1:] = w[..., 1:] != w[..., :-1]
        first &= w >= 0
        bank = np.where(first, w % 32, 32)
        out = np.zeros(lead, np.int64)
        for b in range(32):
            out = np.maximum(out, (bank == b).sum(-1))
        return out
    if not quarter or nbytes < 8:
        return count(words, act)
    g = 32 // (nbytes // 4)     # lanes per group: 8 for 16 B, 16 for 8 B
    tot = 0
    for k in range(32 // g):
        tot = tot + count(words[..., k * g:(k + 1) * g, :], act[..., k * g:(k + 1) * g, :])
    return tot

def host_tables(n, layout=0):
    from apple_b200 import _lib, build
    from bench import build_mesh
    from apple_b200.fem import Region
    build.build(); L = _lib.lib()
    mesh, _, _ = build_mesh(n)
    reg = Region.from_pyvista(mesh, grad=True)          # the product's own rest-shape precompute (no oracle here)
    dhdX, dV = reg.dhdX.reshape(-1, 4, 3), reg.dV.reshape(-1)
    T, V = mesh.n_cells, mesh.n_points
    one = np.ones(T); P = _lib.host_ptr; h = ctypes.c_void_p()
    cells = np.ascontiguousarray(mesh.cells, dtype=np.int32); pts = np.ascontiguousarray(mesh.points)
    rc = L.apl_fem_create(0, _lib.F32, T, V, P(cells), P(dhdX.astype(np.float32)), P(dV.astype(np.float32)),
                          P(one.astype(np.float32)), P(one.astype(np.float32)), None, P(pts), -1, ctypes.byref(h))
    assert rc == 0, L.apl_last_error()
    info = (ctypes.c_int64 * 10)(); L.apl_fem_info(h, info)
    nt, nv, nvo, npk = info[2], info[3], info[8], info[9]
    tiles = np.zeros((nt, 6), np.int32); order = np.zeros(npk, np.int64)
    rows, width = (npk // 2, 8) if layout == 1 else (T, 4)
    conn = np.zeros((rows, width), np.uint8); slots = np.zeros((rows, width), np.uint16)
    tv = np.zeros(nv, np.int32); voff = np.zeros(nvo, np.uint16); vperm = np.zeros(nv, np.uint8)
    L.apl_fem_host_tables(h, P(tiles), P(order), P(conn), P(slots), P(tv), P(voff), P(vperm))
    L.apl_fem_destroy(h)
    return tiles, conn, slots, tv, voff, vperm

def global_lines(tiles, tv, max_tiles=None, row_bytes=12):
    """Distinct 128-byte lines per warp-wide LDGSTS.32 of the producer's vertex gather (ld = 3 rows, three
    4-byte copies per row): the global-memory side of the same LSU pipe.  Returns (lines per instruction with
    the tables' local-id order, with an ascending vertex list)."""
    sel = tiles if max_tiles is None else tiles[np.linspace(0, len(tiles) - 1, max_tiles).astype(int)]
    got = asc = n = 0
    for (ts, nt, vs, nv, vo, ns) in sel:
        gl = tv[vs:vs + nv].astype(np.int64)
        for order, acc in ((gl, 0), (np.sort(gl), 1)):
            tot = 0
            for g in range(0, nv, 32):
                for off in (0, 4, 8):
                    tot += len(np.unique((row_bytes * order[g:g + 32] + off) // 128))
            if acc == 0:
                got += tot
            else:
                asc += tot
        n += 3 * ((nv + 31) // 32)
    return got / n, asc / n


def model(tiles, conn, slots, voff, vperm, quarter=False, max_tiles=None, layout=0, n_real_tets=None):
    acc = dict(static=0, gather=0, slot_st=0, red_small=0, red_ld=0, vbuf_st=0, flush_ld=0)
    ideal = dict(acc)
    ntets = 0
    sel = tiles if max_tiles is None else tiles[np.linspace(0, len(tiles) - 1, max_tiles).astype(int)]
    nc = 5 if layout == 1 else 4                 # corners per item (a tet, or a pair of tets)
    tpi = 2 if layout == 1 else 1                # tets per item
    n_warps = 4 if layout == 1 else 8            # consumer warps per CTA
    for (ts, n_t, vs, nv, vo, nslots) in sel:
        ntets += n_t
        n = n_t // tpi                           # items = consumer threads with work
        it0 = ts // tpi
        nw = (n + 31) // 32
        lane_t = np.arange(nw * 32).reshape(nw, 32)
        act = lane_t < n
        c = np.zeros((nw * 32, nc), np.int64); c[:n] = conn[it0:it0 + n, :nc]
        s = np.zeros((nw * 32, nc), np.int64); s[:n] = slots[it0:it0 + n, :nc]
        c = c.reshape(nw, 32, nc); s = s.reshape(nw, 32, nc)
        # header + connectivity + record planes (3 per tet) + slot ids; the pair item reads 8 B / 16 B rows
        static = (1 + 1 + 12 + 2) if layout == 0 else (1 + 2 + 24 + 4)
        acc["static"] += nw * static; ideal["static"] += nw * static
        for k in range(nc):
            g = wavefronts(c[:, :, k] * 16, 16, act, quarter).sum()
            acc["gather"] += 2 * g; ideal["gather"] += 2 * 4 * nw
            acc["slot_st"] += wavefronts(s[:, :, k] * 16, 16, act, quarter).sum() + wavefronts(s[:, :, k] * 8, 8, act, quarter).sum()
            ideal["slot_st"] += 6 * nw
        # reduce: warp w, lanes j (half 0) and j+16 (half 1) share vertex t = 16 w + j (+128)
        raw = voff[vo:vo + nv + 1].astype(np.int64)
        start = raw & 0x0fff; padded = raw[:-1] >> 12
        cnt = np.diff(start) - padded
        perm = vperm[vs:vs + nv].astype(np.int64)
        for base in range(0, 192, 16 * n_warps):
            for w in range(n_warps):
                t = base + 16 * w + np.arange(16)
                if t[0] >= ((nv + 15) & ~15): continue
                ok = t < nv
                tt = np.where(ok, t, 0)
                s0 = np.where(ok, start[tt], 0); cn = np.where(ok, cnt[tt], 0)
                acc["red_small"] += 3; ideal["red_small"] += 3
                trips = int(np.ceil(cn.max() / 2)) if cn.max() > 0 else 0
                for m in range(trips):
                    i0 = 2 * m
                    a = np.concatenate([s0 + i0, s0 + i0 + 1]); on = np.concatenate([i0 < cn, i0 + 1 < cn])
                    acc["red_ld"] += wavefronts(a[None] * 16, 16, on[None], quarter)[0] + wavefronts(a[None] * 8, 8, on[None], quarter)[0]
                for m in range(trips):      # conflict-free bound: one wavefront per quarter / half warp with an active lane
                    on = np.concatenate([2 * m < cn, 2 * m + 1 < cn])
                    ideal["red_ld"] += on.reshape(4, 8).any(1).sum() + on.reshape(2, 16).any(1).sum()
                v = np.where(ok, perm[tt], 0)
                on = np.concatenate([ok, np.zeros(16, bool)]); a = np.concatenate([v * 24, np.zeros(16, np.int64)])
                for k in range(3):
                    acc["vbuf_st"] += wavefronts((a + 8 * k)[None], 8, on[None], quarter)[0]
                ideal["vbuf_st"] += 3
        for w in range((nv + 31) // 32):
            lanes = 32 * w + np.arange(32); on = lanes < nv
            for k in range(3):
                acc["flush_ld"] += wavefronts((lanes * 24 + 8 * k)[None], 8, on[None], quarter)[0]
            acc["flush_ld"] += 1; ideal["flush_ld"] += 7
    per = {k: 32.0 * v / ntets for k, v in acc.items()}
    per_i = {k: 32.0 * v / ntets for k, v in ideal.items()}
    return per, per_i

if __name__ == "__main__":
    ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=58)
    ap.add_argument("--tiles", type=int, default=400); ap.add_argument("--quarter", action="store_true")
    ap.add_argument("--layout", default="tet", choices=["tet", "pair"])
    a = ap.parse_args()
    layout = 1 if a.layout == "pair" else 0
    tiles, conn, slots, tv, voff, vperm = host_tables(a.n, layout)
    print("tiles", len(tiles), "mean tets", tiles[:, 1].mean(), "mean verts", tiles[:, 3].mean(), "mean slots", tiles[:, 5].mean())
    per, ideal = model(tiles, conn, slots, voff, vperm, a.quarter, a.tiles, layout)
    if layout:
        print("(pair layout: per 32 PACKED tets; clones of unpaired tets are counted as work)")
    ld = per["static"] + per["gather"] + per["red_small"] + per["red_ld"] + per["flush_ld"]
    st = per["slot_st"] + per["vbuf_st"]
    for k in per: print(f"{k:10s} {per[k]:7.2f}   ideal {ideal[k]:7.2f}")
    print(f"ld {ld:.1f} (ncu 143)  st {st:.1f} (ncu 56)   per 32 tets")
    got, asc = global_lines(tiles, tv, a.tiles)
    print(f"global side: {got:.2f} lines per warp-wide LDGSTS.32 of the vertex gather (ascending list: {asc:.2f})")
