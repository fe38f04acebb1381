// Harness around AB/apps/PulsatileBloodFlow2D.h (untouched).  Setup mirrors PulsatileBloodFlow2D() :719-757
// with N a parameter (the reference hard-codes N = 64); the loop body is the reference's :764-790.
//   N=64 steps=2755 vtk=1 dir=/tmp/x  -> writes the reference's own sol_%07d.vtk files every tf/100 steps into dir
//   N=.. steps=.. out=dump.bin dump_at=a,b,c -> binary fp64 P,Ux,Uy,yr1,yr2 + u8 flag + the full lattice at those steps
//   state=FILE -> start from a caller-made state (open vessel at rest of the large-N tests / bench.py), see below
#include "harness_common.h"
#include <unistd.h>
#include <set>
#include <sstream>
#include "PulsatileBloodFlow2D.h"

int main(int argc, char** argv)
{
    Args A(argc, argv);
    int N = A.i("N", 64), steps = A.i("steps", -1), vtk = A.i("vtk", 0);
    std::string dir = A.s("dir", "");
    std::set<int> dump_at;
    { std::stringstream ss(A.s("dump_at", "")); std::string tok; while (std::getline(ss, tok, ',')) if (!tok.empty()) dump_at.insert(std::stoi(tok)); }
    if (!dir.empty() && chdir(dir.c_str()) != 0) { std::perror("chdir"); return 2; }

    Dim_PulsatileBloodFlow2D dim{1 + 10 * (N - 2), N};
    vector<CellData> lattice_vect(LBM_PulsatileBloodFlow2D::sizeOfLattice(dim.nelem));
    CellData* lattice = lattice_vect.data();
    vector<CellType_PulsatileBloodFlow2D> flag_vect(dim.nelem, CellType_PulsatileBloodFlow2D::bulk);
    vector<int> parity_vect{0};
    int* parity = parity_vect.data();
    auto [c_vect, opp_vect, t_vect] = d2q9_constants_PulsatileBloodFlow2D();
    LBM_PulsatileBloodFlow2D lbm{lattice, flag_vect.data(), parity, c_vect.data(), opp_vect.data(), t_vect.data(), dim};
    lbm.tau = A.d("tau", 0.75);
    lbm.s8 = 1.0 / lbm.tau;
    lbm.s5 = 1.0;
    { double Svec[9] = {1, 1, 1, 1, lbm.s5, 1, lbm.s5, lbm.s8, lbm.s8}; std::copy(std::begin(Svec), std::end(Svec), lbm.S); }
    lbm.deformable = A.i("deformable", 1) != 0;
    lbm.is_severed = A.i("is_severed", 1) != 0;
    lbm.alpha = A.d("alpha", 0.01);
    lbm.p0_in = A.d("p0_in", 0.20);
    lbm.p0_out = A.d("p0_out", 0.19);
    lbm.t_beat = max(1, dim.nx);
    lbm.Setup_Simulation_Parameters();
    lbm.Initialize_Yr_and_Vw_and_p();
    lbm.Initialize_Fobj_for_Vessel_Walls();
    lbm.Find_or_Update_Boundary_Nodes();
    lbm.Initialize_P_U_g();

    // state=FILE: start from a caller-made state instead of Initialize_P_U_g (what clbm_pulsatile_upload / the oracle's
    // pulsatile_set_state do): lattice[2*9*nelem], P, Ux, Uy [nelem], yr1, yr2 [nx] as raw doubles; the mask, Fobj and the
    // border lists are re-derived from the wall positions by the REFERENCE's own functions (:275-285, :294-382)
    std::string state = A.s("state", "");
    if (!state.empty()) {
        FILE* f = std::fopen(state.c_str(), "rb");
        if (!f) { std::perror("state"); return 2; }
        auto rd = [&](double* dst, size_t n) { if (std::fread(dst, sizeof(double), n, f) != n) { std::fprintf(stderr, "short state file\n"); std::exit(2); } };
        rd(lattice, lattice_vect.size());
        rd(lbm.P.data(), dim.nelem); rd(lbm.Ux.data(), dim.nelem); rd(lbm.Uy.data(), dim.nelem);
        rd(lbm.yr1.data(), dim.nx); rd(lbm.yr2.data(), dim.nx);
        std::fclose(f);
        lbm.y1new = lbm.yr1; lbm.y2new = lbm.yr2;
        std::fill(lbm.Vw1.begin(), lbm.Vw1.end(), 0.0); std::fill(lbm.Vw2.begin(), lbm.Vw2.end(), 0.0);
        lbm.Initialize_Fobj_for_Vessel_Walls();
        lbm.Find_or_Update_Boundary_Nodes();
        *parity = 0;
    }

    int tf = lbm.t_beat + 2 * lbm.t_propagation;
    int step = max(1, tf / 100);
    if (steps < 0) steps = tf + 1;
    Dump D(A.s("out", ""));
    auto dump = [&]() {
        D.put(lbm.P); D.put(lbm.Ux); D.put(lbm.Uy); D.put(lbm.yr1); D.put(lbm.yr2);
        D.put_u8((uint8_t*)flag_vect.data(), dim.nelem);
        D.put(lattice, lattice_vect.size());
        int par = *parity; if (D.f) std::fwrite(&par, sizeof(int), 1, D.f);
    };
    auto t0 = std::chrono::high_resolution_clock::now();
    for (int t = 0; t < steps; ++t) {
        for_each(std::execution::par_unseq, lattice, lattice + dim.nelem, lbm);
        lbm.Boundary_Conditions();
        lbm.Streaming();
        lbm.Inlet_ZouHe(t);
        lbm.Outlet_ZouHe(t);
        lbm.Macroscopic_Properties_g();
        if (lbm.deformable) lbm.Calculate_Pressure_and_Move_Walls(t);
        if (vtk && t % step == 0) saveVtkFields_PulsatileBloodFlow2D(lbm, t);
        *parity = 1 - *parity;
        if (dump_at.count(t + 1)) dump();
    }
    auto t1 = std::chrono::high_resolution_clock::now();
    report("pulsatile", dim.nelem, steps, 1, std::chrono::duration<double>(t1 - t0).count());
    return 0;
}
