// fock_b200.hpp -- header-only C++ mirror of the reference's two-electron interface over the C ABI
// (include/fock_b200.h).  The reference's host language is Rust, which this image lacks; this is the
// compiled-language host side a maintainer would otherwise write in Rust.
//
//   fock::GaussianPrimitive / ContractedGaussian   <-> basis/src/primitives.rs:7-11, contracted_gaussian.rs:9-15
//   fock::ElectronTensor::from_basis, operator()    <-> basis/src/electron_tensor.rs:81, :179
//   fock::FockBuilder::build_g                      <-> scf/src/lib.rs:177-185
//
// Errors become exceptions here (the reference panics); there is no CPU path.
#pragma once
#include <array>
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fock_b200.h"

namespace fock {

struct GaussianPrimitive {
    std::array<int, 3> angular{};
    double exponent = 0.0;
    double coefficient = 0.0;  // norm folded in by make()

    // basis/src/primitives.rs:19-27
    static double norm(const std::array<int, 3> &a, double alpha) {
        long prod = 1;
        for (int l : a)
            for (int m = l + 1; m <= 2 * l; ++m) prod *= m;
        const double two_over_pi = 0.63661977236758134308;
        return std::sqrt(std::sqrt(std::pow(two_over_pi * alpha, 3))) *
               std::sqrt(std::pow(8.0 * alpha, a[0] + a[1] + a[2]) / (double)prod);
    }
    // GaussianPrimitive::new (primitives.rs:41-48)
    static GaussianPrimitive make(const std::array<int, 3> &a, double exponent, double coefficient) {
        return GaussianPrimitive{a, exponent, coefficient * norm(a, exponent)};
    }
};

struct ContractedGaussian {
    std::array<double, 3> position{};
    std::vector<GaussianPrimitive> primitives;
};

class FockError : public std::runtime_error {
   public:
    FockError(int code, const std::string &msg) : std::runtime_error(msg), code(code) {}
    int code;
};

class FockBuilder {
   public:
    explicit FockBuilder(const std::vector<ContractedGaussian> &basis, const fock_opts *opts = nullptr) : n_((int)basis.size()) {
        std::vector<double> center, exps, coefs;
        std::vector<int32_t> angular, off{0};
        for (const auto &f : basis) {
            if (f.primitives.empty()) throw FockError(FOCK_ERR_BAD_ARG, "contracted Gaussian without primitives");
            center.insert(center.end(), f.position.begin(), f.position.end());
            for (int k = 0; k < 3; ++k) angular.push_back(f.primitives[0].angular[k]);
            for (const auto &p : f.primitives) {
                exps.push_back(p.exponent);
                coefs.push_back(p.coefficient);
            }
            off.push_back((int32_t)exps.size());
        }
        fock_basis b{n_, center.data(), angular.data(), off.data(), exps.data(), coefs.data()};
        const int st = fock_create(&b, opts, &ctx_);
        if (st != FOCK_OK) throw FockError(st, fock_last_error(nullptr));
    }
    FockBuilder(const FockBuilder &) = delete;
    FockBuilder &operator=(const FockBuilder &) = delete;
    FockBuilder(FockBuilder &&o) noexcept : ctx_(o.ctx_), n_(o.n_) { o.ctx_ = nullptr; }
    ~FockBuilder() { fock_destroy(ctx_); }

    int n_basis() const { return n_; }
    void eri_build() { check(fock_eri_build(ctx_)); }
    double eri(int i, int j, int k, int l) const {
        double v = 0.0;
        check(fock_eri_get(ctx_, i, j, k, l, &v));
        return v;
    }
    std::vector<double> export_packed() const {
        std::vector<double> out(fock_packed_len(n_));
        check(fock_eri_export_packed(ctx_, out.data(), out.size()));
        return out;
    }
    // G = J - K/2, column-major N x N in and out (nalgebra DMatrix layout)
    std::vector<double> build_g(const std::vector<double> &density) const {
        if ((int)density.size() != n_ * n_) throw FockError(FOCK_ERR_BAD_ARG, "density must be N x N");
        std::vector<double> g((size_t)n_ * n_);
        check(fock_build_g(ctx_, density.data(), g.data()));
        return g;
    }
    // MolecularOrbital::evaluate (scf/src/molecular_orbitals.rs:37-43) for one coefficient column at `points`
    // ([n][3], bohr); |coeff| <= min_coeff dropped as MolecularOrbitals::new does (:70-79).
    std::vector<double> orbital_eval(const std::vector<double> &coeff, const std::vector<double> &points, double min_coeff = 0.05) const {
        if ((int)coeff.size() != n_ || points.size() % 3) throw FockError(FOCK_ERR_BAD_ARG, "coeff must be N, points [n][3]");
        std::vector<double> v(points.size() / 3);
        check(fock_orbital_eval(ctx_, coeff.data(), min_coeff, points.data(), v.size(), v.data()));
        return v;
    }
    // State::create_wave_texture_data (volumetric/src/main.rs:100-125): n^3 f32 voxels, x fastest
    std::vector<float> orbital_voxelize(const std::vector<double> &coeff, int n_voxels = 200, double box_size = 20.0,
                                        double min_coeff = 0.05) const {
        if ((int)coeff.size() != n_) throw FockError(FOCK_ERR_BAD_ARG, "coeff must be N");
        std::vector<float> g((size_t)n_voxels * n_voxels * n_voxels);
        check(fock_orbital_voxelize(ctx_, coeff.data(), min_coeff, n_voxels, box_size, g.data()));
        return g;
    }
    // The exchange step (fock_comm_*): rank 0 makes an id, every rank attaches; build_g then returns the complete G.
    static std::array<char, FOCK_COMM_ID_BYTES> comm_unique_id() {
        std::array<char, FOCK_COMM_ID_BYTES> id{};
        const int st = fock_comm_unique_id(id.data());
        if (st != FOCK_OK) throw FockError(st, fock_comm_last_error());
        return id;
    }
    void comm_init(const std::array<char, FOCK_COMM_ID_BYTES> &id) { check(fock_comm_init(ctx_, id.data())); }
    bool comm_attached() const { return fock_comm_attached(ctx_) != 0; }
    std::vector<double> diagonal() const {
        std::vector<double> d((size_t)n_ * (n_ + 1) / 2);
        check(fock_diagonal_export(ctx_, d.data(), d.size()));
        return d;
    }
    fock_ctx *raw() const { return ctx_; }

   private:
    void check(int st) const {
        if (st != FOCK_OK) throw FockError(st, fock_last_error(ctx_));
    }
    fock_ctx *ctx_ = nullptr;
    int n_ = 0;
};

// ElectronTensor::from_basis(&basis) and tensor[(i,j,k,l)]
class ElectronTensor {
   public:
    static ElectronTensor from_basis(const std::vector<ContractedGaussian> &basis, const fock_opts *opts = nullptr) {
        ElectronTensor t(basis, opts);
        t.builder_.eri_build();
        return t;
    }
    double operator()(int i, int j, int k, int l) const { return builder_.eri(i, j, k, l); }
    int size() const { return builder_.n_basis(); }
    FockBuilder &builder() { return builder_; }

   private:
    ElectronTensor(const std::vector<ContractedGaussian> &basis, const fock_opts *opts) : builder_(basis, opts) {}
    FockBuilder builder_;
};

}  // namespace fock
