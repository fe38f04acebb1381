// vtk_writer_check.cpp -- test program (tests/test_vtk_writer.py): drives coolbm::VtkWriter (multiphase-lbm_b200/apps/case_common.h)
// with analytic fields, once per output format, without touching a device.
#include <array>

#include "../../multiphase-lbm_b200/apps/case_common.h"

int main(int argc, char **argv)
{
    const int nx = argc > 1 ? std::atoi(argv[1]) : 5, ny = argc > 2 ? std::atoi(argv[2]) : 4, nz = argc > 3 ? std::atoi(argv[3]) : 1;
    coolbm::VtkWriter w(42, nx, ny, nz, 0.125);
    auto xyz = [=](size_t i, int &x, int &y, int &z) { z = (int)(i % nz); y = (int)((i / nz) % ny); x = (int)(i / ((size_t)nz * ny)); };
    w.scalars("Density", "float", [&](size_t i) { int x, y, z; xyz(i, x, y, z); return 1.0 + 0.5 * x + 0.25 * y + 0.125 * z; });
    w.scalars("Flag", "int", [&](size_t i) { int x, y, z; xyz(i, x, y, z); return (y == 0 || y == ny - 1) ? "1" : "0"; }, true);
    w.vectors("Force", [&](size_t i) { int x, y, z; xyz(i, x, y, z); return std::array<double, 3>{1.0 * x, -1.0 * y, 0.5 * z}; });
    w.vectors_rows("Velocity", [&](size_t i) { int x, y, z; xyz(i, x, y, z); return std::array<double, 3>{0.1 * x, 0.2 * y, 9.0}; });
    return 0;
}
