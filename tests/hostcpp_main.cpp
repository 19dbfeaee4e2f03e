// Exercises the C++ host mirror end to end (compiled and run by tests/test_hostcpp.py on the GPU box).
// Builds H2 / STO-3G by hand (the geometry of the reference's test_integrals, scf/src/lib.rs:273-285:
// R = 1.2 bohr along z), prints (00|00), (10|10), (11|00), G for D = identity, and two orbital values + two voxels.
//
// `hostcpp_main multi <n>`: the multi-GPU path from compiled host code alone (no Python, no torch): n threads, one
// context per GPU (rank = device ordinal), one NCCL communicator attached through fock_comm_init; every rank's
// fock_build_g must return the COMPLETE G (what scf/src/lib.rs:177-185 computes), equal to the single-GPU one.
// Prints "MULTI_OK <max |G_n - G_1|> <max |G_rank - G_rank0|>".
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "../fock_rs_b200/hostcpp/fock_b200.hpp"

// a small water-chain-like s/p basis built by hand: enough shells for every s/p class and a few hundred quartets
static std::vector<fock::ContractedGaussian> chain_basis(int n_centres) {
    using namespace fock;
    const double es[3] = {5.0331513, 1.1695961, 0.3803890}, cs[3] = {-0.09996723, 0.39951283, 0.70011547};
    const double cp[3] = {0.15591627, 0.60768372, 0.39195739};
    std::vector<ContractedGaussian> basis;
    for (int a = 0; a < n_centres; ++a) {
        const std::array<double, 3> pos{1.7 * a, 0.3 * (a % 2), 0.1 * a};
        ContractedGaussian s;
        s.position = pos;
        for (int i = 0; i < 3; ++i) s.primitives.push_back(GaussianPrimitive::make({0, 0, 0}, es[i], cs[i]));
        basis.push_back(s);
        const int ang[3][3] = {{0, 0, 1}, {0, 1, 0}, {1, 0, 0}};  // the reference's p order (basis_set/src/lib.rs:87-99)
        for (auto &l : ang) {
            ContractedGaussian p;
            p.position = pos;
            for (int i = 0; i < 3; ++i) p.primitives.push_back(GaussianPrimitive::make({l[0], l[1], l[2]}, es[i], cp[i]));
            basis.push_back(p);
        }
    }
    return basis;
}

static int run_multi(int world) {
    using namespace fock;
    const auto basis = chain_basis(6);
    const int n = (int)basis.size();
    std::vector<double> D((size_t)n * n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) D[i + (size_t)j * n] = D[j + (size_t)i * n] = std::sin(1.0 + 0.37 * i + 0.11 * j);
    std::vector<double> g1;
    {
        fock_opts o;
        fock_default_opts(&o);
        o.storage = FOCK_STORAGE_DIRECT;
        o.device = 0;
        FockBuilder b(basis, &o);
        b.eri_build();
        g1 = b.build_g(D);
    }
    const auto id = FockBuilder::comm_unique_id();
    std::vector<std::vector<double>> g(world);
    std::vector<std::string> errs(world);
    std::vector<std::thread> th;
    for (int r = 0; r < world; ++r)
        th.emplace_back([&, r] {
            try {
                fock_opts o;
                fock_default_opts(&o);
                o.storage = FOCK_STORAGE_DIRECT;
                o.device = r;
                o.rank = r;
                o.world = world;
                FockBuilder b(basis, &o);
                b.comm_init(id);
                b.eri_build();
                g[r] = b.build_g(D);
                g[r] = b.build_g(D);  // a second iteration through the same communicator
            } catch (const FockError &e) {
                errs[r] = e.what();
            }
        });
    for (auto &t : th) t.join();
    for (int r = 0; r < world; ++r)
        if (!errs[r].empty()) {
            std::printf("ERROR rank %d: %s\n", r, errs[r].c_str());
            return 2;
        }
    double d1 = 0.0, dr = 0.0;
    for (int r = 0; r < world; ++r)
        for (size_t i = 0; i < g1.size(); ++i) {
            d1 = std::fmax(d1, std::fabs(g[r][i] - g1[i]));
            dr = std::fmax(dr, std::fabs(g[r][i] - g[0][i]));
        }
    std::printf("MULTI_OK %.3e %.3e\n", d1, dr);
    return 0;
}

int main(int argc, char **argv) {
    using namespace fock;
    if (argc >= 3 && !std::strcmp(argv[1], "multi")) return run_multi(std::atoi(argv[2]));
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
