// Compile check of the shim against the REFERENCE'S OWN headers (not the mock): run where /root/reference exists, with
// the stand-in third-party headers of oracle/refbuild/standin:
//   g++ -std=c++17 -fsyntax-only -fopenmp -I include -I oracle/refbuild/standin -I oracle/refbuild \
//       -I /root/reference/srcs shim/compile_check_reference.cpp
#include "simulation/Equation.hpp"
#include "simulation/Problem.hpp"
#include "simulation/Solver.hpp"
#include "mesh/Mesh.hpp"

#include "pfem_b200_equations.hpp"

template class MomContEqIncompNewtonB200<2>;
template class MomContEqIncompNewtonB200<3>;
template class WCompNewtonStepB200<2>;
template class WCompNewtonStepB200<3>;
int main() { return 0; }
