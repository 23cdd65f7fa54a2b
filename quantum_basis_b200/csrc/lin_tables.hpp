// quantum_basis_b200/csrc/lin_tables.hpp -- host-side sector tables and model parameters shared by the on-device
// generators (builders.cu: the reference's Lin-table order, src/basis.cc:1144-1190) and the species-order layouts of the
// Hubbard model (species.cu).
#pragma once
#include "internal.hpp"
#include <vector>

namespace qb {

constexpr int kMaxBonds = 256;

struct Bond { int i, j; int w; };    // site pair with multiplicity (duplicates in the caller's list are merged)

struct ModelParams {
    int kind;                         // 0 heisenberg, 1 hubbard
    double J, t, U;
    int nbonds;
    Bond bonds[kMaxBonds];
};

// Lin tables of a sector (see builders.cu: SectorTables is the device view of these arrays)
struct HostTables {
    int nsites = 0, bps = 1, nA = 0, nB = 0, t0 = 0, t1 = 0;
    int64_t dim = 0;
    std::vector<int64_t> Jb;
    std::vector<int32_t> rankA, class_off;
    std::vector<uint32_t> alist;
};

int make_tables(int nsites, int bps, int t0, int t1, HostTables &T);                 // builders.cu
int merge_bonds(int nsites, int nbonds, const int32_t *bonds, ModelParams &M);       // builders.cu

// builders.cu: perm[r] = rank_in_class[up word of row r] * Dd + rank_in_class[down word of row r] for every row r of the
// reference's Lin order of the electron sector T (bps == 2); d_rank_in_class and d_perm are device arrays.
int species_perm_build(const HostTables &T, const int32_t *d_rank_in_class, int64_t Dd, int32_t *d_perm);
// out[p - lo] = the reference's row of species index p for lo <= p < hi (device arrays); the start vectors of sharded runs
int species_ref_rows_build(const HostTables &T, const int32_t *d_rank_in_class, int64_t Dd, int64_t lo, int64_t hi, int32_t *d_out);
void species_perm_host(const HostTables &T, const int32_t *rank_in_class, int64_t Dd, int32_t *perm);   // same row function, host arrays

// species.cu: the species-order handles behind qbgpu_build_hubbard / qbgpu_create_matfree_hubbard with QBGPU_SPECIES_ORDER
int species_build_stored(qbgpu_matrix_t *out, const HostTables &T, const ModelParams &M, int api_complex, int flags,
                         int64_t row_lo = 0, int64_t row_hi = -1);
int species_build_matfree(qbgpu_matrix_t *out, const HostTables &T, const ModelParams &M, int api_complex, int flags,
                          int64_t row_lo = 0, int64_t row_hi = -1);

}  // namespace qb
