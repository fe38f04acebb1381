// mrt.cuh -- w = M^-1 S M v for one D2Q9 population vector in the k-ordering of the SC / PF case headers
// (c_k = (-1,0),(0,-1),(-1,-1),(-1,1),(0,0),(1,0),(0,1),(1,1),(1,-1)), moment basis and rates of
// CooLBM_MRT_combustion.cpp:313-323 / :339: rows (rho, e, eps, jx, qx, jy, qy, pxx, pxy),
//   e = -4 + 3 c^2,  eps = 4 - 21/2 c^2 + 9/2 c^4,  q = (-5 + 3 c^2) c,  pxx = cx^2 - cy^2,  pxy = cx cy,
// S = (s_c, s_e, s_eps, s_c, s_q, s_c, s_q, s_nu, s_nu).  The rows are orthogonal with squared norms
// 9, 36, 36, 6, 12, 6, 12, 4, 4, so M^-1 = M^T diag(1/norm^2) (the M_inv table of :326-336) and the whole operator is
// 9 sums / differences, 9 scalings and the mirrored back-substitution: ~70 FP64 operations instead of two 9x9 products.
// Used by the HCZ D2Q9 kernels when clbm_params.collision = CLBM_COLLISION_MRT (s_c = s_nu = omega there).
#pragma once
#include "lattice.cuh"

namespace clbm {

struct MrtRates { double s_c, s_e, s_eps, s_q, s_nu; };

CLBM_D void mrt9_relax(const double *v, const MrtRates &S, double *w)
{
    const double r = v[4];
    const double ax = v[0] + v[5], ay = v[1] + v[6];
    const double d1 = v[7] + v[2], d2 = v[3] + v[8];             // the two diagonals
    const double a = ax + ay, ad = d1 + d2;
    const double dxa = v[5] - v[0], dya = v[6] - v[1];
    const double e72 = v[7] - v[2], e83 = v[8] - v[3];
    const double dxd = e72 + e83, dyd = e72 - e83;
    // moments, scaled by rate / squared row norm
    const double n0 = S.s_c * (1. / 9.) * (r + a + ad);
    const double n1 = S.s_e * (1. / 36.) * (2.0 * ad - a - 4.0 * r);
    const double n2 = S.s_eps * (1. / 36.) * (4.0 * r - 2.0 * a + ad);
    const double n3 = S.s_c * (1. / 6.) * (dxa + dxd);
    const double n4 = S.s_q * (1. / 12.) * (dxd - 2.0 * dxa);
    const double n5 = S.s_c * (1. / 6.) * (dya + dyd);
    const double n6 = S.s_q * (1. / 12.) * (dyd - 2.0 * dya);
    const double n7 = S.s_nu * (1. / 4.) * (ax - ay);
    const double n8 = S.s_nu * (1. / 4.) * (d1 - d2);
    // back to population space: w_k = sum_j M[j][k] n_j
    w[4] = n0 - 4.0 * n1 + 4.0 * n2;
    const double ea = n0 - n1 - 2.0 * n2;
    const double ox = n3 - 2.0 * n4, oy = n5 - 2.0 * n6;
    w[5] = ea + n7 + ox;
    w[0] = ea + n7 - ox;
    w[6] = ea - n7 + oy;
    w[1] = ea - n7 - oy;
    const double ed = n0 + 2.0 * n1 + n2;
    const double px = n3 + n4, py = n5 + n6;
    w[7] = ed + n8 + (px + py);
    w[2] = ed + n8 - (px + py);
    w[8] = ed - n8 + (px - py);
    w[3] = ed - n8 - (px - py);
}

// ---- D3Q19: w = M^-1 S M v in the orthogonal basis of d'Humieres et al. (Phil. Trans. R. Soc. A 360, 2002) evaluated at the c_k of
// the PF laplace3D.h ordering (oracle/clbm_oracle.c: mrt19_rows, same rows, same order):
//   rho | e = 19 c^2 - 30 | eps = (21 c^4 - 53 c^2 + 24)/2 | j_a = c_a, q_a = (5 c^2 - 9) c_a | 3 p_xx, 3 pi_xx = (3 c^2 - 5) 3 p_xx |
//   p_ww = c_y^2 - c_z^2, pi_ww | p_xy, p_yz, p_xz | m_x, m_y, m_z.
// Rates by moment order, carried over from the D2Q9 assignment: conserved + stress moments omega, e: s_e, eps and pi: s_eps,
// q and m: s_q.  The rows are compile-time constants: after unrolling only the non-zero products remain (about 230 of 361 per
// direction of the transform), in the oracle's k = 0..18 / j = 0..18 summation order.  The reference has no D3Q19 MRT operator:
// parity unpinned, pinned at S = omega I to the BGK kernels.
struct Mrt19 {
    CLBM_HD static constexpr double row(int j, int k)
    {
        const double cx = D3Q19::cx(k), cy = D3Q19::cy(k), cz = D3Q19::cz(k), c2 = cx * cx + cy * cy + cz * cz;
        switch (j) {
        case 0: return 1.0;
        case 1: return 19.0 * c2 - 30.0;
        case 2: return (21.0 * c2 * c2 - 53.0 * c2 + 24.0) / 2.0;
        case 3: return cx;
        case 4: return (5.0 * c2 - 9.0) * cx;
        case 5: return cy;
        case 6: return (5.0 * c2 - 9.0) * cy;
        case 7: return cz;
        case 8: return (5.0 * c2 - 9.0) * cz;
        case 9: return 3.0 * cx * cx - c2;
        case 10: return (3.0 * c2 - 5.0) * (3.0 * cx * cx - c2);
        case 11: return cy * cy - cz * cz;
        case 12: return (3.0 * c2 - 5.0) * (cy * cy - cz * cz);
        case 13: return cx * cy;
        case 14: return cy * cz;
        case 15: return cx * cz;
        case 16: return (cy * cy - cz * cz) * cx;
        case 17: return (cz * cz - cx * cx) * cy;
        default: return (cx * cx - cy * cy) * cz;
        }
    }
    CLBM_HD static constexpr double norm2(int j)
    {
        double s = 0.0;
        for (int k = 0; k < 19; ++k) s += row(j, k) * row(j, k);
        return s;
    }
    // which of the five rates a moment relaxes with: 0 omega, 1 s_e, 2 s_eps, 3 s_q
    CLBM_HD static constexpr int rate(int j)
    {
        constexpr int r[19] = {0, 1, 2, 0, 3, 0, 3, 0, 3, 0, 2, 0, 2, 0, 0, 0, 3, 3, 3};
        return r[j];
    }
};

CLBM_D void mrt19_relax(const double *v, const MrtRates &S, double *w)
{
    double m[19];
#pragma unroll
    for (int j = 0; j < 19; ++j) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < 19; ++k) {
            const double c = Mrt19::row(j, k);
            if (c != 0.0) a += c * v[k];
        }
        const int r = Mrt19::rate(j);
        m[j] = (r == 0 ? S.s_c : (r == 1 ? S.s_e : (r == 2 ? S.s_eps : S.s_q))) * a;
    }
#pragma unroll
    for (int k = 0; k < 19; ++k) {
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < 19; ++j) {
            const double c = Mrt19::row(j, k) / Mrt19::norm2(j);
            if (c != 0.0) a += c * m[j];
        }
        w[k] = a;
    }
}

}  // namespace clbm
