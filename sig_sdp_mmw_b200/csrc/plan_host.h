// Host-side graph plan (native C++): what mmw._process_state (mmw.py:26-41) and the
// edge-list set-up (mmw.py:52-57) produce, in the flat layout the kernels walk.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace sigsdp {

struct HostPlan {
    int64_t n = 0;
    int64_t E_g = 0, E_a = 0;   // undirected gain / association edges
    int64_t nnz = 0;            // union pattern incl. diagonal = n + 2 (E_g + E_a)
    int64_t nnzT = 0;           // stored entries of T
    int max_row = 0;
    // union pattern in INTERNAL numbering, columns ascending inside a row
    std::vector<int32_t> rowptr, col, eid;   // eid: -1 diag, [0,E_g) gain, E_g + a asso
    std::vector<double> tfwd, tbwd;          // T[row,col], T[col,row] (0 off the gain pattern)
    // edge lists in the CALLER's numbering and the reference's order
    std::vector<int32_t> gi, gj, ai, aj;
    std::vector<double> tij, tji;
    // node vectors, INTERNAL numbering
    std::vector<double> S_sum, tnorm, h_max;
    // perm[new] = old, iperm[old] = new
    std::vector<int32_t> perm, iperm;
    int order = 0;
};

// returns 0 or a negative SIGSDP_E* code; err gets the message
int build_host_plan(int64_t n, const int32_t* Sp, const int32_t* Si, const double* Sx,
                    const int32_t* Qp, const int32_t* Qi, const double* Qx, const double* h_max,
                    int order, HostPlan& out, std::string& err);

// sequential greedy pass of sdp_solver.rounding_one_attempt (sdp_solver.py:70-101)
int round_greedy_host(int64_t n, int Z, const int32_t* Sp, const int32_t* Si, const double* Sx,
                      const int32_t* Qp, const int32_t* Qi, const double* Qx, const double* h_max,
                      const int32_t* rank, const int32_t* pref, int32_t* z_vec, int64_t* remainder,
                      std::string& err);

}  // namespace sigsdp
