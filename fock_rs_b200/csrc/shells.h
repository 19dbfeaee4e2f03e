// shells.h -- host-side regrouping of the flat basis-function list into shells, shell pairs and
// per-class pair sublists, plus the primitive-pair data the class kernels read.
//
// The reference discards shell structure when it expands BSE shells into functions
// (basis_set/src/lib.rs:84 TODO); it is reconstructed here from what the C ABI receives:
// consecutive functions with the same centre, the same exponent list, the same total angular
// momentum and proportional coefficient lists form one shell.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace fock {

constexpr int kMaxClassL = 2;  // class kernels cover s, p and d shells; the rest goes to the generic kernel
constexpr int kNumPairClasses = 6;  // ss, ps, pp, ds, dp, dd
// pair class of shells with angular momenta la >= lb
inline int pair_class(int la, int lb) { return la * (la + 1) / 2 + lb; }
constexpr int kPairClassLA[kNumPairClasses] = {0, 1, 1, 2, 2, 2}, kPairClassLB[kNumPairClasses] = {0, 0, 1, 0, 1, 2};

inline int ncart(int L) { return (L + 1) * (L + 2) / 2; }

// Component index of Cartesian powers (x,y,z) within a shell of L = x+y+z, in the reference's
// enumeration order (x outer .. z inner, basis_set/src/lib.rs:87-99): p = (z,y,x),
// d = (zz,yz,yy,xz,xy,xx).
inline int cart_index(int x, int y, int z) {
    const int L = x + y + z;
    return x * (L + 1) - x * (x - 1) / 2 + y;
}

struct Shell {
    int L = 0;
    double c[3] = {0, 0, 0};
    int nprim = 0;
    int prim_begin = 0;  // into ShellData::exps / coefs (radial coefficients, norm of component 0 folded in)
    int bf[6] = {-1, -1, -1, -1, -1, -1};  // basis-function index of each Cartesian component (-1 = absent)
    double scale[6] = {1, 1, 1, 1, 1, 1};  // coefficient list of component c = scale[c] * radial list
    bool class_ok = false;  // eligible for the class kernels (L <= kMaxClassL; s and p need unit component scales)
};

struct ShellPair {
    int A = 0, B = 0;    // shell ids, (L_A, A) >= (L_B, B)
    int cls = 0;         // 0 ss, 1 ps, 2 pp, 3 ds, 4 dp, 5 dd
    int K = 0;           // nprim_A * nprim_B
    double maxdiag = 0;  // max over components of the basis-function diagonal (xy|xy)
};

// Pairs of one class with one primitive-pair count, sorted by maxdiag (descending) once the
// diagonal is known.  `nf` doubles per primitive pair (see pair_fields()).
struct Sublist {
    int cls = 0, K = 0, nf = 0;
    std::vector<int> pair_ids;  // into ShellData::pairs
    // device arrays (owned by PairTablesDev)
    size_t bra_off = 0;  // doubles offset into d_bra: [pair][K][nf]
    size_t ket_off = 0;  // doubles offset into d_ket: [K][nf][npad]
    size_t meta_off = 0; // ints offset into d_meta: [pair] -> {shellA, shellB}
    int npad = 0;
};

struct ShellData {
    std::vector<Shell> shells;
    std::vector<double> exps, coefs;
    std::vector<int> bf_shell, bf_comp;  // per basis function
    std::vector<ShellPair> pairs;        // class-eligible pairs only
    std::vector<Sublist> sublists;
    bool all_class_ok = false;           // every function sits in a class-eligible shell
    int n_class_shells = 0;
    int n_partial_shells = 0;  // s/p/d shells missing Cartesian components (fragmented by the regrouping)
    std::string describe() const;
};

struct PairTablesDev {
    double *d_bra = nullptr;
    double *d_ket = nullptr;
    int *d_meta = nullptr;      // [npairs_total][2]
    int *d_shell_bf = nullptr;  // [nshell][6]
    double *d_maxdiag = nullptr;
    size_t bra_doubles = 0, ket_doubles = 0, meta_ints = 0;
};

// Doubles per primitive pair for pair class cls; layouts in shells.cpp (fill_record) / eri_class.cuh.
int pair_fields(int cls);

// Build shells from the flat basis.  Returns false with `err` set on malformed input.
bool build_shells(int n_basis, const double *center, const int32_t *angular, const int32_t *prim_offset,
                  const double *exponent, const double *coefficient, ShellData &out, std::string &err);

// Enumerate class-eligible shell pairs and split them into (class, K) sublists in natural order.
void build_pairs(ShellData &sd);

// Sort every sublist by maxdiag (descending); pair.maxdiag must be set.
void sort_sublists(ShellData &sd);

// Fill host images of the bra (AoS) and ket (SoA) primitive-pair tables and the meta table for
// the current sublist order.  Sets the offsets in each Sublist.
void build_pair_tables(ShellData &sd, std::vector<double> &bra, std::vector<double> &ket, std::vector<int> &meta);

// Model FP64 flops of one primitive quartet of class (la lb|lc ld) -- SURVEY.md section 8d.
double model_flops_per_prim_quartet(int la, int lb, int lc, int ld);

}  // namespace fock
