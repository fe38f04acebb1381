// hcz3d_edges.cuh -- where the partial moment sums of a tile's perimeter live (single-sweep HCZ D3Q19 step, hcz3d_sweep.cu).
//
// The sweep kernel produces the moments of the NEXT step (phi = sum f, P_term = sum g, j = sum c g of the populations that will
// have arrived) while it pushes: a CTA owns a TY x TZ (y,z) tile and can only add up what its own nodes push.  For a node
// inside the tile that is everything; a node on the tile's border also receives from the neighbouring tiles, and the tile
// itself pushes into the one-cell ring around it.  Those ring sums are stored in EDGE arrays with exactly one writer per
// slot (plain stores, no atomics, deterministic), and whoever reads a moment of a border node adds, in a fixed order,
//
//     M[node]  +  EY[...]  +  EZ[...]  +  EC[...]
//
//   EY  bottom / top : what the tile BELOW / ABOVE (same tile column) pushed into the node's row        [2][nTY][nz]
//   EZ  left / right : what the tile to the LEFT / RIGHT (same tile row) pushed into the node's column  [2][ny][nTZ]
//   EC  4 corners    : what the DIAGONAL tile pushed into a corner node (one direction of D3Q19)        [4][nTY][nTZ]
//
// per x-plane and moment (5 moments).  Both sides -- the producer's ring cell and the consumer's border node -- go through
// edge_offsets() below, so they cannot disagree; tests/test_hcz3d_edges.py emulates the whole scheme on the CPU with it.
#pragma once
#include "lattice.cuh"

namespace clbm {

struct EdgeGeom {
    int ny, nz, nTY, nTZ;
    int off_eyb, off_eyt, off_ezl, off_ezr, off_ec;   // starts of the five blocks inside one plane of an edge array
    int eplane;                                       // doubles per plane
};

template <int TY, int TZ>
CLBM_HD EdgeGeom make_edge_geom(int ny, int nz)
{
    EdgeGeom e;
    e.ny = ny; e.nz = nz; e.nTY = ny / TY; e.nTZ = nz / TZ;
    e.off_eyb = 0;
    e.off_eyt = e.off_eyb + e.nTY * nz;
    e.off_ezl = e.off_eyt + e.nTY * nz;
    e.off_ezr = e.off_ezl + ny * e.nTZ;
    e.off_ec = e.off_ezr + ny * e.nTZ;
    e.eplane = e.off_ec + 4 * e.nTY * e.nTZ;
    return e;
}

// edge slots of the node (yy, zz) (global, already wrapped into [0,ny) x [0,nz)):
//   e[0] = its EY slot (node in the first / last row of its tile), e[1] = its EZ slot (first / last column),
//   e[2] = its EC slot (tile corner); -1 where the node has none
template <int TY, int TZ>
CLBM_HD void edge_offsets(const EdgeGeom &g, int yy, int zz, int e[3])
{
    const int ly = yy % TY, lz = zz % TZ, R = yy / TY, C = zz / TZ;
    const bool yb = ly == 0, yt = ly == TY - 1, zl = lz == 0, zr = lz == TZ - 1;
    e[0] = yb ? g.off_eyb + R * g.nz + zz : (yt ? g.off_eyt + R * g.nz + zz : -1);
    e[1] = zl ? g.off_ezl + yy * g.nTZ + C : (zr ? g.off_ezr + yy * g.nTZ + C : -1);
    e[2] = ((yb || yt) && (zl || zr)) ? g.off_ec + (((yt ? 2 : 0) + (zr ? 1 : 0)) * g.nTY + R) * g.nTZ + C : -1;
}

// the slot a tile writes for the cell (dy, dz) of its ring (tile coordinates, dy in {-1..TY}, dz in {-1..TZ}, on the ring):
// the ring cell is a border node of a neighbouring tile; it arrives from below / above (EY), from the side (EZ) or
// diagonally (EC) according to which side of OUR tile it lies on
template <int TY, int TZ>
CLBM_HD int ring_slot(const EdgeGeom &g, int y0, int z0, int dy, int dz)
{
    int yy = y0 + dy, zz = z0 + dz;
    yy = yy < 0 ? yy + g.ny : (yy >= g.ny ? yy - g.ny : yy);
    zz = zz < 0 ? zz + g.nz : (zz >= g.nz ? zz - g.nz : zz);
    int e[3];
    edge_offsets<TY, TZ>(g, yy, zz, e);
    const bool ys = dy < 0 || dy >= TY, zs = dz < 0 || dz >= TZ;
    return (ys && zs) ? e[2] : (ys ? e[0] : e[1]);
}

// ---- the push of a tile's post-collision plane as TMA box stores (hcz3d_sweep.cu, VAR bit 2) ----
// A box store needs a start that is 16-byte aligned in global memory and has no negative coordinate (the device traps otherwise,
// tools/probe/tma_store_probe.cu); z is the fastest index, so only the directions with c_z = 0 qualify (z0 is a multiple of TZ,
// TZ is even), and only in tiles whose shifted box stays inside [0, ny): not the first / last tile row.  Everything else is
// pushed by the threads, with the periodic wrap.  Kernel and CPU emulation (tests/test_hcz3d_edges.py) share this rule.
template <int TY, int TZ>
CLBM_HD bool push_by_box(int y0, int ny, int cz) { return cz == 0 && y0 > 0 && y0 + TY < ny; }
// start of that box in (y, z) for a direction with c_z = 0
struct PushBox { int y, z; };
template <int TY, int TZ>
CLBM_HD PushBox push_box_start(int y0, int z0, int cy) { return PushBox{y0 + cy, z0}; }

// what the nodes of ONE plane of the tile push into the cell (dy, dz) (tile coordinates; ring cells have dy = -1 / TY or
// dz = -1 / TZ), grouped by c_x: index 0 = A (c_x = +1), 1 = B (c_x = 0), 2 = C (c_x = -1).  S = the plane's post-collision
// populations [38][TY][TZ].  jx needs no sums of its own: it is +P_term(A) - P_term(C).
struct PushSums { double ph[3], pt[3], jy[3], jz[3]; };

// FY / FZ: the only c_y / c_z whose source node can lie inside the tile (2 = any).  A cell of the ring row below the tile
// (dy = -1) is reached from the tile's first row by c_y = -1 only, and so on: the filter drops the other directions at
// compile time (they would be predicated off at run time: same sums, a quarter of the instructions).
template <int TY, int TZ, int FY = 2, int FZ = 2>
CLBM_D void gather_pushes(const double *S, int dy, int dz, PushSums &o)
{
    constexpr int NT = TY * TZ;
    // source node of direction k: (dy - c_y, dz - c_z); inside the tile?
    const bool vy[3] = {dy + 1 >= 0 && dy + 1 < TY, dy >= 0 && dy < TY, dy - 1 >= 0 && dy - 1 < TY};   // index c_y + 1
    const bool vz[3] = {dz + 1 >= 0 && dz + 1 < TZ, dz >= 0 && dz < TZ, dz - 1 >= 0 && dz - 1 < TZ};   // index c_z + 1
#pragma unroll
    for (int j = 0; j < 3; ++j) o.ph[j] = o.pt[j] = o.jy[j] = o.jz[j] = 0.0;
    const int base = dy * TZ + dz;
#pragma unroll
    for (int k = 0; k < 19; ++k) {
        const int cx = D3Q19::cx(k), cy = D3Q19::cy(k), cz = D3Q19::cz(k);
        if ((FY != 2 && cy != FY) || (FZ != 2 && cz != FZ)) continue;
        const int grp = cx > 0 ? 0 : (cx == 0 ? 1 : 2);
        const bool ok = vy[cy + 1] && vz[cz + 1];
        const int src = base - cy * TZ - cz;
        const double vf = ok ? S[k * NT + src] : 0.0;
        const double vg = ok ? S[(19 + k) * NT + src] : 0.0;
        o.ph[grp] += vf;
        o.pt[grp] += vg;
        if (cy > 0) o.jy[grp] += vg; else if (cy < 0) o.jy[grp] -= vg;
        if (cz > 0) o.jz[grp] += vg; else if (cz < 0) o.jz[grp] -= vg;
    }
}


// One plane's pushes (s, of source plane xsrc) folded into a cell's running sums:
//   T[5] = (A + B) of plane xsrc-1 on entry, of plane xsrc on exit: phi, P_term, P_term(A) (for jx = P_term(A) - P_term(C)), jy, jz
//   A[4] = group A of plane xsrc on entry, of plane xsrc+1 on exit: phi, P_term, jy, jz
//   v[5] <- the completed moments (phi, P_term, jx, jy, jz) of plane xsrc-1 = (A + B) + C; for the FIRST plane (xsrc = 0) only
//           the C group of plane -1 exists: v is then what gets parked in plane nx-1's slot until the march has come round
CLBM_D void fold_pushes(double T[5], double A[4], const PushSums &s, bool first, double v[5])
{
    if (!first) {
        v[0] = T[0] + s.ph[2]; v[1] = T[1] + s.pt[2]; v[2] = T[2] - s.pt[2]; v[3] = T[3] + s.jy[2]; v[4] = T[4] + s.jz[2];
    } else {
        v[0] = s.ph[2]; v[1] = s.pt[2]; v[2] = -s.pt[2]; v[3] = s.jy[2]; v[4] = s.jz[2];
    }
    T[0] = A[0] + s.ph[1]; T[1] = A[1] + s.pt[1]; T[2] = A[1]; T[3] = A[2] + s.jy[1]; T[4] = A[3] + s.jz[1];
    A[0] = s.ph[0]; A[1] = s.pt[0]; A[2] = s.jy[0]; A[3] = s.jz[0];
}

// after the last plane (periodic x): plane nx-1 = its (A + B) + the parked C group; plane 0 = what was stored for it (B + C)
// + the A group the last plane pushed.  m = moment index 0..4.
CLBM_D double finish_last(const double T[5], int m, double parked) { return T[m] + parked; }
CLBM_D double finish_first(const double A[4], int m, double stored) { return stored + (m == 0 ? A[0] : (m <= 2 ? A[1] : A[m - 1])); }

}  // namespace clbm
