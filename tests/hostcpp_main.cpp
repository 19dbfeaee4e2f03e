// Exercises the C++ host mirror end to end (compiled and run by tests/test_hostcpp.py on the GPU box).
// Builds H2 / STO-3G by hand (the geometry of the reference's test_integrals, scf/src/lib.rs:273-285:
// R = 1.2 bohr along z), prints (00|00), (10|10), (11|00), G for D = identity, and two orbital values + two voxels.
#include <cstdio>

#include "../fock_rs_b200/hostcpp/fock_b200.hpp"

int main() {
    using namespace fock;
    const double exps[3] = {0.3425250914e+01, 0.6239137298e+00, 0.1688554040e+00};
    const double coefs[3] = {0.1543289673e+00, 0.5353281423e+00, 0.4446345422e+00};
    std::vector<ContractedGaussian> basis(2);
    basis[0].position = {0.0, 0.0, 0.0};
    basis[1].position = {0.0, 0.0, 1.2};
    for (auto &f : basis)
        for (int i = 0; i < 3; ++i) f.primitives.push_back(GaussianPrimitive::make({0, 0, 0}, exps[i], coefs[i]));
    try {
        ElectronTensor t = ElectronTensor::from_basis(basis);
        std::printf("%.15e %.15e %.15e\n", t(0, 0, 0, 0), t(1, 0, 1, 0), t(1, 1, 0, 0));
        const std::vector<double> g = t.builder().build_g({1.0, 0.0, 0.0, 1.0});
        std::printf("%.15e %.15e %.15e %.15e\n", g[0], g[1], g[2], g[3]);
        // sigma_g-like orbital (0.55, 0.55) at the bond midpoint and at (0.3, -0.2, 0.1); 2^3 voxel grid, box 4
        const std::vector<double> psi = t.builder().orbital_eval({0.55, 0.55}, {0.0, 0.0, 0.6, 0.3, -0.2, 0.1});
        const std::vector<float> vox = t.builder().orbital_voxelize({0.55, 0.55}, 2, 4.0);
        std::printf("%.15e %.15e %.9e %.9e\n", psi[0], psi[1], (double)vox[0], (double)vox[7]);
    } catch (const FockError &e) {
        std::printf("ERROR %d %s\n", e.code, e.what());
        return 2;
    }
    return 0;
}
