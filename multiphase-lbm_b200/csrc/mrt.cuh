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

}  // namespace clbm
