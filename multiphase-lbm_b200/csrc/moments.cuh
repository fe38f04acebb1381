// moments.cuh -- zeroth/first moments of a population vector held in registers, and the EOS.
// Summation groups follow the reference so that the device result differs from the CPU
// functor only by FMA contraction:
//   D2Q9 : SC/apps/laplace2D.h:148-170, PF/apps/rayleighTaylor2D.h:197-230
//   D3Q19: PF/apps/laplace3D.h:216-258
#pragma once
#include "lattice.cuh"

namespace clbm {

template <class L> struct Mom;

template <> struct Mom<D2Q9> {
    CLBM_D static double sum(const double *f)
    {
        double m = f[0] + f[2] + f[3], p = f[5] + f[7] + f[8], z = f[6] + f[1] + f[4];
        return m + p + z;
    }
    // raw first moments (no division)
    CLBM_D static void first(const double *f, double &jx, double &jy, double &jz)
    {
        double xm = f[0] + f[2] + f[3], xp = f[5] + f[7] + f[8];
        double ym = f[1] + f[2] + f[8], yp = f[3] + f[6] + f[7];
        jx = xp - xm;
        jy = yp - ym;
        jz = 0.0;
    }
};

template <> struct Mom<D3Q19> {
    CLBM_D static double sum(const double *f)
    {
        double m = f[0] + f[3] + f[4] + f[5] + f[6];
        double p = f[10] + f[13] + f[14] + f[15] + f[16];
        double z = f[9] + f[1] + f[2] + f[7] + f[8] + f[11] + f[12] + f[17] + f[18];
        return m + p + z;
    }
    CLBM_D static void first(const double *f, double &jx, double &jy, double &jz)
    {
        double xm = f[0] + f[3] + f[4] + f[5] + f[6], xp = f[10] + f[13] + f[14] + f[15] + f[16];
        double ym = f[1] + f[3] + f[7] + f[8] + f[14], yp = f[4] + f[11] + f[13] + f[17] + f[18];
        double zm = f[2] + f[5] + f[7] + f[16] + f[18], zp = f[6] + f[8] + f[12] + f[15] + f[17];
        jx = xp - xm;
        jy = yp - ym;
        jz = zp - zm;
    }
};

// ---- Yuan / Carnahan-Starling pseudopotential (SC/apps/laplace2D.h:173-195) ----
struct ScEos {
    double R, TT, a;
    CLBM_D double Z(double rho) const
    {
        const double d = 1.0 - rho;
        return 1.0 + (4.0 * rho - 2.0 * rho * rho) / (d * d * d);
    }
    // s = R T Z - a rho - cs2 ; G1 = sign(s)/3 ; P - cs2 rho = rho * s
    CLBM_D double G1_of_Z(double rho, double Zr) const
    {
        const double s = R * TT * Zr - a * rho - (1.0 / 3.0);
        return (s > 0.0) ? (1.0 / 3.0) : -(1.0 / 3.0);
    }
    CLBM_D double psi_of_Z(double rho, double Zr, double G1) const
    {
        const double P = rho * R * TT * Zr - a * rho * rho;
        const double val = 6.0 * (P - (1.0 / 3.0) * rho) / G1;
        return (val > 0.0) ? sqrt(val) : 0.0;
    }
    CLBM_D double psi(double rho) const
    {
        const double Zr = Z(rho);
        return psi_of_Z(rho, Zr, G1_of_Z(rho, Zr));
    }
};

// ---- fast FP64 reciprocal / square root for the kernels that are checked at 1e-10 (NOT for the bit-exact paths) --------
// An IEEE FP64 division or square root is a 25-35 instruction sequence with a slow-path branch.  The hardware seeds
// (rcp.approx / rsqrt.approx, ~2^-22 relative) refined by two Newton steps in FMA arithmetic are accurate to 1-2 ulp in
// about 8 instructions.  Arguments must be normal and positive where a square root is taken (callers guard).
CLBM_D double fast_rcp(double x)
{
#if defined(CLBM_HOST_CHECK) && !defined(__CUDA_ARCH__)
    return 1.0 / x;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
#endif
}
CLBM_D double fast_sqrt(double x)      // x > 1e-290
{
#if defined(CLBM_HOST_CHECK) && !defined(__CUDA_ARCH__)
    return sqrt(x);
#else
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = y * fma(-0.5 * x * y, y, 1.5);           // y ~ x^-1/2
    y = y * fma(-0.5 * x * y, y, 1.5);
    const double s = x * y;
    return fma(fma(-s, s, x), 0.5 * y, s);       // one correction of s ~ x^1/2
#endif
}

// ---- HCZ Carnahan-Starling "psi" = p_th(x) - x/3 (PF/apps/rayleighTaylor2D.h:237-242, 374-379) ----
CLBM_D double hcz_psi(double x, double a, double b)
{
    const double rt = b * x * 0.25;
    const double d = 1.0 - rt;
    const double pth = (x / 3.0) * (1.0 + rt + rt * rt - rt * rt * rt) / (d * d * d) - a * x * x;
    return pth - x / 3.0;
}

// Same function with ONE division (an IEEE FP64 division is a ~30-instruction sequence on the SM): x/3 becomes
// x * (1/3) and the quotient N/D is formed once.  Differs from hcz_psi by O(1 ulp); used by the fused kernels.
CLBM_D double hcz_psi1(double x, double a, double b)
{
    const double rt = b * x * 0.25;
    const double d = 1.0 - rt;
    const double x3 = x * (1.0 / 3.0);
    return x3 * ((1.0 + rt + rt * rt - rt * rt * rt) * fast_rcp(d * d * d)) - a * x * x - x3;
}

}  // namespace clbm
