// Baseline instrumentation (test infrastructure): exposes the time the reference's CPU solve
// (here: the dense stand-in for Eigen::SimplicialLLT, see eigen_standin/Eigen/SparseCore) has
// consumed, so bench.py can report `ba` of the reference with and without its CPU solve.
#include <Eigen/SparseCore>
extern "C" double droid_ref_solve_seconds() { return Eigen::standin_solve_seconds(); }
extern "C" long droid_ref_solve_calls() { return Eigen::standin_solve_calls(); }
