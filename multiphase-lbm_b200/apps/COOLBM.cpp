// COOLBM.cpp -- case selector of the B200 drivers.  The reference hard-codes `string problem = "...";` in
// apps/COOLBM.cpp and is rebuilt per case (SC/apps/COOLBM.cpp:66-84, PF/apps/COOLBM.cpp, AB/apps/COOLBM.cpp:118-150);
// here the same names select the case at run time:
//     COOLBM <problem> [config_dir]        problem in {laplace2D, contactAngle2D, twoLayeredFlow2D, droplet3D, rayleighTaylor2D, laplace3D, twoLayeredPF2D, Young_Laplace2D,
//                                          PulsatileBloodFlow2D [N] [max_iter] [vtk 0/1]}
// config_dir defaults to ../apps/Config_Files, the path the reference drivers open relative to their build directory.
#include <array>

#include "PulsatileBloodFlow2D.h"
#include "RayleighTaylor2D.h"
#include "Young_Laplace2D.h"
#include "contactAngle2D.h"
#include "laplace2D.h"
#include "laplace3D.h"
#include "rayleighTaylor2D.h"
#include "twoLayeredFlow2D.h"
#include "twoLayeredPF2D.h"

std::string problem = "laplace2D";

int main(int argc, char **argv)
{
    if (argc > 1) problem = argv[1];
    try {
        if (problem == "PulsatileBloodFlow2D") {
            const int N = argc > 2 ? std::stoi(argv[2]) : 64, max_iter = argc > 3 ? std::stoi(argv[3]) : -1;
            coolbm::PulsatileBloodFlow2D(N, 0.75, 0.01, true, true, max_iter, argc > 4 ? std::stoi(argv[4]) != 0 : true);
            return 0;
        }
        const std::string dir = argc > 2 ? argv[2] : "../apps/Config_Files";
        if (problem == "laplace2D") coolbm::Laplace2D(dir);
        else if (problem == "contactAngle2D") coolbm::contactAngle2D(dir);
        else if (problem == "droplet3D") coolbm::droplet3D(dir);
        else if (problem == "twoLayeredFlow2D") coolbm::twoLayeredFlow2D(dir);
        else if (problem == "Young_Laplace2D") coolbm::Young_Laplace2D(dir);
        else if (problem == "twoLayeredPF2D") coolbm::twoLayeredPF2D(dir);
        else if (problem == "rayleighTaylor2D") coolbm::rayleighTaylor2D(dir);
        else if (problem == "RayleighTaylor2D") coolbm::RayleighTaylor2D(dir);   // the Shan-Chen one (capital R, as in SC/apps)
        else if (problem == "laplace3D") coolbm::laplace3D(dir);
        else { std::cerr << "unknown problem \"" << problem << "\"\n"; return 2; }
    } catch (const std::exception &e) {
        std::cerr << "terminate: " << e.what() << "\n";
        return 1;
    }
    return 0;
}
