// g++ -std=c++17 -fsyntax-only -I include -I shim/mock shim/compile_check.cpp
#include "pfem3d_mock.hpp"
#include "../pfem_b200_equations.hpp"
template class MomContEqIncompNewtonB200<2>;
template class MomContEqIncompNewtonB200<3>;
template class WCompNewtonStepB200<2>;
template class WCompNewtonStepB200<3>;
int main() { return 0; }
