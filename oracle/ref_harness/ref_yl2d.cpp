// Harness around AB/apps/Young_Laplace2D.h (untouched).  Setup mirrors Young_Laplace2D() :499-526, loop body :555-565.
// Dumps C, P, Rho, Ux, Uy, the whole lattice (4*9*nelem) and the parity after `steps` iterations.
#include "harness_common.h"
#include "Young_Laplace2D.h"
int main(int argc, char** argv)
{
    Args A(argc, argv);
    int nx = A.i("nx", 32), ny = A.i("ny", 32), steps = A.i("steps", 10);
    Dim_YL2D dim{nx, ny};
    vector<CellData> lattice_vect(LBM_Young_Laplace2D::sizeOfLattice(dim.nelem), 0.0);
    CellData* lattice = lattice_vect.data();
    vector<CellType_YL2D> flag_vect(dim.nelem, CellType_YL2D::bulk);
    int p_store = 0;
    int* parity = &p_store;
    auto [c_vect, opp_vect, t_vect] = d2q9_constants_YL2D();
    LBM_Young_Laplace2D lbm{lattice, flag_vect.data(), parity, c_vect.data(), opp_vect.data(), t_vect.data(), dim};
    lbm.Sigma = A.d("Sigma", 0.01); lbm.W = A.d("W", 4.0); lbm.M = A.d("M", 0.02);
    lbm.Rhol = A.d("RhoL", 0.001); lbm.Rhoh = A.d("RhoH", 1.0);
    lbm.tau = A.d("tau", 0.8); lbm.s8 = 1.0 / lbm.tau;
    lbm.Beta = 12.0 * lbm.Sigma / lbm.W;
    lbm.kappa = 1.5 * lbm.Sigma * lbm.W;
    lbm.dRho3 = (lbm.Rhoh - lbm.Rhol) / 3.0;
    for (size_t i = 0; i < dim.nelem; ++i) lbm.iniCell((int)i);
    inigeom_Young_Laplace2D(lbm);
    lbm.update_fields();
    vector<int> cell_index(dim.nelem);
    std::iota(cell_index.begin(), cell_index.end(), 0);
    auto t0 = std::chrono::high_resolution_clock::now();
    for (int it = 0; it < steps; ++it) {
        std::for_each(std::execution::par_unseq, cell_index.begin(), cell_index.end(), [&](int i) { lbm.collide_stream_at(i); });
        *parity = 1 - *parity;
        lbm.update_fields();
    }
    auto t1 = std::chrono::high_resolution_clock::now();
    report("yl2d", dim.nelem, steps, 1, std::chrono::duration<double>(t1 - t0).count());
    Dump D(A.s("out", ""));
    if (D.f) {
        D.put(lbm.C); D.put(lbm.P); D.put(lbm.Rho); D.put(lbm.Ux); D.put(lbm.Uy);
        D.put(lattice, lattice_vect.size());
        int par = *parity; std::fwrite(&par, sizeof(int), 1, D.f);
    }
    return 0;
}
