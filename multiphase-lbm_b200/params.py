"""clbm_params mirror (include/clbm.h) and the reference's derived lattice-unit parameters.

The struct carries the scalar members of the reference LBM_* aggregates
(SC/apps/laplace2D.h:104-114, PF/apps/rayleighTaylor2D.h:113-123).
"""
import ctypes

ABI_VERSION = 3

MODEL_SC_D2Q9, MODEL_SC_D3Q19, MODEL_HCZ_D2Q9, MODEL_HCZ_D3Q19, MODEL_PULSATILE = range(5)
SC_FORCE_LAPLACE, SC_FORCE_CONTACT, SC_FORCE_CONSTG, SC_FORCE_EXPGUO = 0, 1, 2, 3
HCZ_FORCE_GRAVITY, HCZ_FORCE_LAYERED = 0, 1
COLLISION_BGK, COLLISION_MRT = 0, 1
REDUCE_MASS, REDUCE_ENERGY, REDUCE_UMAX = 0, 1, 2
(CASE_SC_LAPLACE2D, CASE_SC_CONTACT2D, CASE_SC_DROPLET3D, CASE_SC_DROPLET3D_PER,
 CASE_HCZ_RT2D, CASE_HCZ_LAPLACE3D, CASE_SC_LAYERED2D, CASE_HCZ_LAYERED2D, CASE_SC_RT2D) = range(9)

MODEL_Q = {MODEL_SC_D2Q9: 9, MODEL_SC_D3Q19: 19, MODEL_HCZ_D2Q9: 9, MODEL_HCZ_D3Q19: 19, MODEL_PULSATILE: 9}
MODEL_SETS = {MODEL_SC_D2Q9: 1, MODEL_SC_D3Q19: 1, MODEL_HCZ_D2Q9: 2, MODEL_HCZ_D3Q19: 2, MODEL_PULSATILE: 1}
# algorithmic bytes per lattice update: every population read once + written once + 1 mask byte
MODEL_BYTES_PER_LU = {m: 2 * MODEL_SETS[m] * MODEL_Q[m] * 8 + 1 for m in MODEL_Q}


class Params(ctypes.Structure):
    _fields_ = [
        ("abi_version", ctypes.c_int32), ("model", ctypes.c_int32),
        ("nx", ctypes.c_int32), ("ny", ctypes.c_int32), ("nz", ctypes.c_int32),
        ("nx_global", ctypes.c_int32), ("x_offset", ctypes.c_int32),
        ("sc_force", ctypes.c_int32), ("device", ctypes.c_int32), ("fused", ctypes.c_int32),
        ("omega", ctypes.c_double), ("gravity", ctypes.c_double),
        ("rho_w", ctypes.c_double), ("a", ctypes.c_double), ("b", ctypes.c_double),
        ("R", ctypes.c_double), ("TT", ctypes.c_double),
        ("phi_l", ctypes.c_double), ("phi_g", ctypes.c_double),
        ("rho_l", ctypes.c_double), ("rho_g", ctypes.c_double), ("kappa", ctypes.c_double),
        ("gx", ctypes.c_double), ("gy", ctypes.c_double), ("G", ctypes.c_double), ("p_shift", ctypes.c_double),
        ("gx_const", ctypes.c_double),
        ("s_e", ctypes.c_double), ("s_eps", ctypes.c_double), ("s_q", ctypes.c_double),
        ("collision", ctypes.c_int32), ("reserved0", ctypes.c_int32),
    ]

    @property
    def nelem(self):
        return self.nx * self.ny * self.nz

    @property
    def Q(self):
        return MODEL_Q[self.model]

    @property
    def sets(self):
        return MODEL_SETS[self.model]

    @property
    def lattice_size(self):
        """sizeOfLattice(): 2 buffers x sets x Q x nelem doubles (SC/apps/laplace2D.h:93)."""
        return 2 * self.sets * self.Q * self.nelem

    def copy(self, **kw):
        q = Params.from_buffer_copy(bytes(self))
        for k, v in kw.items():
            setattr(q, k, v)
        return q


def lb_parameters(ulb, lref, Re):
    """lbParameters_*: nu, omega, dx, dt (SC/apps/laplace2D.h:52-58)."""
    nu = ulb * lref / Re
    omega = 1.0 / (3.0 * nu + 0.5)
    dx = 1.0 / lref
    dt = dx * ulb
    return nu, omega, dx, dt


def make_params(model, nx, ny, nz=1, **kw):
    p = Params()
    p.abi_version = ABI_VERSION
    p.model = model
    p.nx, p.ny, p.nz = nx, ny, nz
    p.nx_global, p.x_offset = nx, 0
    p.device = -1
    p.fused = 1
    p.omega = 1.0
    p.a, p.b, p.R = 1.0, 4.0, 1.0
    p.G = -1.0
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def sc_params(model, nx, ny, nz=1, *, omega=None, tau=None, ulb=0.01, N=None, Re=6.0, rho_w=0.12,
              a=1.0, b=4.0, R=1.0, TT0=0.875, gravity=0.0, sc_force=SC_FORCE_LAPLACE, **kw):
    """Shan-Chen parameter set; defaults = SC/apps/Config_Files/config_Laplace2D.txt.
    TT = TT0 * Tc, Tc = 0.3773 a / (b R)  (SC/apps/laplace2D.h:468-470)."""
    if omega is None:
        omega = 1.0 / tau if tau is not None else lb_parameters(ulb, N if N else nx, Re)[1]
    Tc = 0.3773 * a / (b * R)
    return make_params(model, nx, ny, nz, omega=omega, rho_w=rho_w, a=a, b=b, R=R, TT=TT0 * Tc,
                       gravity=gravity, sc_force=sc_force, **kw)


def hcz_params(model, nx, ny, nz=1, *, omega=None, ulb=0.04, N=None, Re=3000.0, phi_l=0.251, phi_g=0.024,
               rho_l=0.12, rho_g=0.04, a=4.0, b=4.0, kappa=0.01, gravity=-6.25e-6, **kw):
    """HCZ parameter set; defaults = PF/apps/Config_Files/config_rayleighTaylor2D.txt."""
    if omega is None:
        omega = lb_parameters(ulb, N if N else nx, Re)[1]
    return make_params(model, nx, ny, nz, omega=omega, phi_l=phi_l, phi_g=phi_g, rho_l=rho_l, rho_g=rho_g,
                       a=a, b=b, kappa=kappa, gravity=gravity, **kw)


def hcz_mrt_params(nx, ny, *, s_e=None, s_eps=None, s_q=None, **kw):
    """HCZ D2Q9 with the MRT collision operator (include/clbm.h, CLBM_COLLISION_MRT); a rate left at None equals omega,
    so hcz_mrt_params(nx, ny) is the BGK operator evaluated in moment space"""
    p = hcz_params(MODEL_HCZ_D2Q9, nx, ny, **kw)
    p.collision = COLLISION_MRT
    p.s_e = p.omega if s_e is None else s_e
    p.s_eps = p.omega if s_eps is None else s_eps
    p.s_q = p.omega if s_q is None else s_q
    return p


def sc_mrt_params(nx, ny, *, s_e=None, s_eps=None, s_q=None, **kw):
    """Yuan-CS Shan-Chen D2Q9 with the MRT collision operator (CLBM_COLLISION_MRT); a rate left at None equals omega"""
    p = sc_params(MODEL_SC_D2Q9, nx, ny, **kw)
    p.collision = COLLISION_MRT
    p.s_e = p.omega if s_e is None else s_e
    p.s_eps = p.omega if s_eps is None else s_eps
    p.s_q = p.omega if s_q is None else s_q
    return p


class PulsatileParams(ctypes.Structure):
    """clbm_pulsatile_params mirror (include/clbm.h): the user-set members of LBM_PulsatileBloodFlow2D
    (AB/apps/PulsatileBloodFlow2D.h:740-749)."""
    _fields_ = [
        ("abi_version", ctypes.c_int32), ("N", ctypes.c_int32), ("device", ctypes.c_int32),
        ("is_severed", ctypes.c_int32), ("deformable", ctypes.c_int32), ("t_beat", ctypes.c_int32),
        ("tau", ctypes.c_double), ("alpha", ctypes.c_double), ("p0_in", ctypes.c_double), ("p0_out", ctypes.c_double),
    ]


def pulsatile_params(N=64, tau=0.75, alpha=0.01, p0_in=0.20, p0_out=0.19, is_severed=1, deformable=1, device=-1):
    """defaults = the values hard-coded in PulsatileBloodFlow2D() (AB/apps/PulsatileBloodFlow2D.h:721, :740-749)"""
    p = PulsatileParams()
    p.abi_version = ABI_VERSION
    p.N, p.device, p.is_severed, p.deformable, p.t_beat = N, device, int(is_severed), int(deformable), 0
    p.tau, p.alpha, p.p0_in, p.p0_out = tau, alpha, p0_in, p0_out
    return p


# Pulsatile: 145 B/LU as the other single-set D2Q9 paths; the reference also stores P, Ux, Uy every step (169)
PULSATILE_BYTES_PER_LU = 169


def sc_p_shift(rhog, rhol, a, b, R, TT):
    """p_shift of the constant-G variant: the smallest shift that keeps rho/3 - (P_eos + p_shift) >= 0 on [rhog, rhol],
    sampled at 601 points, plus 1e-12 (SC/apps/twoLayeredFlow2D.h:535-546)."""
    worst = -1e30
    Ns = 600
    for s in range(Ns + 1):
        r = rhog + (rhol - rhog) * (float(s) / Ns)
        d = 1.0 - r
        Z = 1.0 + (4.0 * r - 2.0 * r * r) / (d * d * d)
        S = (1.0 / 3.0) * r - (r * R * TT * Z - a * r * r)
        worst = max(worst, -S)
    return max(0.0, worst) + 1e-12


def sc_layered_params(nx, ny, *, omega=None, tau=None, ulb=0.1, N=None, Re=60.0, rhol=0.21, rhog=0.067, rho_w=0.067, a=1.0, b=4.0,
                      R=1.0, TT0=0.95, gx=1e-8, gy=0.0, G=-1.0, **kw):
    """constant-G Shan-Chen parameter set; defaults = SC/apps/Config_Files/config_twoLayeredFlow2D.txt"""
    p = sc_params(MODEL_SC_D2Q9, nx, ny, omega=omega, tau=tau, ulb=ulb, N=N if N else ny - 1, Re=Re, rho_w=rho_w, a=a, b=b, R=R,
                  TT0=TT0, sc_force=SC_FORCE_CONSTG, **kw)
    p.gx, p.gy, p.G = gx, gy, G
    p.p_shift = sc_p_shift(rhog, rhol, a, b, R, p.TT)
    return p


def sc_rt_params(nx, ny=None, *, omega=None, tau=None, ulb=0.04, N=None, Re=30.72, g=-5.0, gravity=-1.25e-5, rho_w=0.2, a=1.0, b=4.0, **kw):
    """Shan-Chen Rayleigh-Taylor parameter set (psi = 1 - exp(-rho), Guo forcing); defaults =
    SC/apps/Config_Files/config_RayleighTaylor2D.txt; the lattice is N x (4N + 2) (SC/apps/RayleighTaylor2D.h:603)"""
    if ny is None:
        ny = 4 * nx + 2
    p = sc_params(MODEL_SC_D2Q9, nx, ny, omega=omega, tau=tau, ulb=ulb, N=N if N else nx, Re=Re, rho_w=rho_w, a=a, b=b,
                  gravity=gravity, sc_force=SC_FORCE_EXPGUO, **kw)
    p.G = g
    return p


class YL2DParams(ctypes.Structure):
    """clbm_yl2d_params mirror (include/clbm.h): the config keys of Young_Laplace2D() (AB/apps/Young_Laplace2D.h:465-494)."""
    _fields_ = [
        ("abi_version", ctypes.c_int32), ("nx", ctypes.c_int32), ("ny", ctypes.c_int32), ("device", ctypes.c_int32),
        ("Sigma", ctypes.c_double), ("W", ctypes.c_double), ("M", ctypes.c_double),
        ("RhoL", ctypes.c_double), ("RhoH", ctypes.c_double), ("tau", ctypes.c_double),
    ]


def yl2d_params(nx=128, ny=128, Sigma=0.01, W=4.0, M=0.02, RhoL=0.001, RhoH=1.0, tau=0.8, device=-1):
    """defaults = AB/apps/Config_Files/config_laplace2D.txt"""
    p = YL2DParams()
    p.abi_version, p.nx, p.ny, p.device = ABI_VERSION, nx, ny, device
    p.Sigma, p.W, p.M, p.RhoL, p.RhoH, p.tau = Sigma, W, M, RhoL, RhoH, tau
    return p


# Young-Laplace: two population sets read + written once, plus the stored velocity (2 doubles read + written)
YL2D_BYTES_PER_LU = 4 * 9 * 8 + 4 * 8


def hcz_layered_params(nx, ny, *, omega=None, tau=None, ulb=0.1, N=None, Re=60.0, phi_l=0.251, phi_g=0.024, rho_l=0.12, rho_g=0.04,
                       a=4.0, b=4.0, kappa=0.001, gx=0.0, gx_const=1e-8, **kw):
    """HCZ two-layered channel flow; defaults = PF/apps/Config_Files/config_twoLayeredFlow2D.txt"""
    if omega is None:
        omega = 1.0 / tau if tau is not None else lb_parameters(ulb, N if N else ny - 1, Re)[1]
    p = make_params(MODEL_HCZ_D2Q9, nx, ny, omega=omega, phi_l=phi_l, phi_g=phi_g, rho_l=rho_l, rho_g=rho_g, a=a, b=b, kappa=kappa,
                    gravity=0.0, sc_force=HCZ_FORCE_LAYERED, **kw)
    p.gx, p.gx_const = gx, gx_const
    return p
