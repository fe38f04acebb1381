"""Host-side initial states for the compliant-vessel path beyond the reference's own Initialize_P_U_g.

The reference's hard-coded set-up (AB/apps/PulsatileBloodFlow2D.h:740-757) starts from a vessel that is closed at
the inlet and at most two rows wide, independent of N; for N >= 128 that start diverges within a few iterations in
the reference itself (the untouched header segfaults).  `open_vessel_at_rest` is the state the
reference's N = 64 run relaxes to: walls at their zero-over-pressure position, fluid at rest at the tissue pressure.
It is handed to the device through clbm_pulsatile_upload, exactly as a reference-side driver would hand over its
own arrays.
"""
import numpy as np

T9 = np.array([1 / 9., 1 / 9., 1 / 36., 1 / 36., 4 / 9., 1 / 9., 1 / 9., 1 / 36., 1 / 36.])


def open_vessel_at_rest(N, p_tissue=0.02, alpha=0.01, margin=0.0):
    """-> dict(lattice, flag, P, Ux, Uy, yr1, yr2) in the reference layout (i = y + ny x, lattice[p*9*nelem + k*nelem + i]).
    Walls at yr1 = 0.5 + margin, yr2 = ny - 1.5 - margin (the wall ODE's rest position for P = p_tissue - margin*alpha,
    AB:243-272); fluid nodes at rest with g_k = t_k P (Equilibrium_g with U = 0, AB:501-507)."""
    nx, ny = 1 + 10 * (N - 2), N
    ne = nx * ny
    Y0 = (ny - 1) // 2
    c = Y0 + 0.5
    yr1 = np.full(nx, 0.5 + margin)
    yr2 = np.full(nx, ny - 1.5 - margin)
    Y = np.arange(ny)
    F = np.where(Y <= Y0, (yr1[0] - c) / (Y - c), (yr2[0] - c) / (Y - c))
    col_flag = np.where(F < 1.0, 0, 1).astype(np.uint8)
    flag = np.tile(col_flag, nx)
    Pval = p_tissue - margin * alpha
    P = np.where(flag == 1, Pval, 0.0)
    lattice = np.zeros(2 * 9 * ne)
    for k in range(9):
        lattice[k * ne:(k + 1) * ne] = T9[k] * P
    return {"lattice": lattice, "flag": flag, "P": P, "Ux": np.zeros(ne), "Uy": np.zeros(ne), "yr1": yr1, "yr2": yr2}
