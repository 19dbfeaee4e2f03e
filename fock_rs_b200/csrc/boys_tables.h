// boys_tables.h -- per-interval polynomial form of the Boys function for the class kernels.
//
// The reference's boys(T, n) for |T| < 30 (boys/src/lib.rs:58-81) is, on each table interval
// [k*0.15, (k+1)*0.15), an analytic function of T: its F_10 estimate is linear in T there and the
// downward recursion adds polynomial multiples of exp(-T).  Re-expanding that SAME function as a
// degree-7 polynomial about the interval midpoint reproduces it to ~1e-16 while replacing one
// exp() and a ten-step recursion by seven FMAs.  This is a change of evaluation scheme, not of the
// function: the table is generated from the reference's own LUT and recursion (in long double) at
// context creation, and tests compare it against the oracle's boys() on a dense grid.
#pragma once
#include <vector>

namespace fock {

constexpr int kBoysPolyIntervals = 200;  // T in [0, 30), width 0.15 (the reference's DT)
constexpr int kBoysPolyDegree = 7;
constexpr int kBoysPolyCoefs = kBoysPolyDegree + 1;
constexpr int kBoysPolyOrders = 9;  // F_0 .. F_8 (shell classes up to (dd|dd))
// accurate mode: exact F_n on [0, 40.05) with the same interval width
constexpr int kBoysAccIntervals = 267;

// Layout: [order][coef][interval] doubles, followed (faithful table only) by exp(-mid_k) [interval].
// `lut` = the reference's 200-entry F_10 table.
void make_boys_poly_faithful(const double *lut, std::vector<double> &out);
void make_boys_poly_accurate(std::vector<double> &out);

inline size_t boys_poly_faithful_len() { return (size_t)kBoysPolyOrders * kBoysPolyCoefs * kBoysPolyIntervals + kBoysPolyIntervals; }
inline size_t boys_poly_accurate_len() { return (size_t)kBoysPolyOrders * kBoysPolyCoefs * kBoysAccIntervals + kBoysAccIntervals; }

}  // namespace fock
