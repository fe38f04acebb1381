// dropin_common.h -- shared helpers of the DROP-IN harnesses (test infrastructure).
//
// A drop-in harness is the test of INTEGRATION.md's patch: it #includes ONE untouched case header of the reference where it
// lies under /root/reference, lets the REFERENCE build the state (its own iniLattice + inigeom_*), and then advances two
// copies of that state:
//   (A) with the reference's own hot line  for_each(par_unseq, lattice, lattice + nelem, lbm); *parity = 1 - *parity;
//   (B) through the C ABI: clbm_create / clbm_upload / clbm_step(n) / clbm_download_lattice into the second host array
// and finally evaluates the REFERENCE's accessors (density, u_actual, macro_phi_P, velocity, ...) on both arrays.  The binary
// prints one JSON line of relative L-inf errors and exits 0 iff all are below tol (default 1e-10) and the masks are equal.
// Built by `make -C oracle ref` (only where /root/reference exists) into oracle/_ref/, linked against libclbm.so.
#pragma once
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "clbm.h"
#include "harness_common.h"

struct ErrList {
    std::vector<std::pair<std::string, double>> e;
    double tol;
    bool ok = true;
    explicit ErrList(double t) : tol(t) {}
    // relative L-inf with global-max normalisation
    void field(const char* name, const std::vector<double>& got, const std::vector<double>& ref)
    {
        double num = 0, den = 0;
        for (size_t i = 0; i < ref.size(); ++i) { num = std::fmax(num, std::fabs(got[i] - ref[i])); den = std::fmax(den, std::fabs(ref[i])); }
        const double r = den > 0 ? num / den : num;
        e.emplace_back(name, r);
        if (!(r < tol)) ok = false;
    }
    void exact(const char* name, bool same) { e.emplace_back(name, same ? 0.0 : 1.0); if (!same) ok = false; }
    int finish(const char* what, size_t nelem, int steps)
    {
        std::printf("{\"case\": \"%s\", \"nelem\": %zu, \"steps\": %d, \"tol\": %.1e, \"ok\": %s", what, nelem, steps, tol, ok ? "true" : "false");
        for (auto& p : e) std::printf(", \"%s\": %.3e", p.first.c_str(), p.second);
        std::printf("}\n");
        return ok ? 0 : 1;
    }
};

#define DROPIN_CLBM(call)                                                                          \
    do {                                                                                           \
        if ((call) != CLBM_OK) { std::fprintf(stderr, "%s failed: %s\n", #call, clbm_last_error()); return 2; } \
    } while (0)
