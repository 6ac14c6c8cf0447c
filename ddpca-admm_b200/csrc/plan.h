// plan.h -- host-side planning of one multigrid level's device layout.
//
// The reference smooths with lexicographic symmetric Gauss-Seidel
// (MGPIS.h:65-77): row i of the forward sweep needs every coupled row j < i.
// On the GPU a level is therefore split into STAGES: sets of row GROUPS that
// have no coupling between them and can be relaxed in parallel; stages run in
// sequence.  A GROUP is a run of <= 3 consecutive rows sharing one column
// pattern (the DOFs of one mesh node, MULTIGRID.h:1170-1176) and is relaxed
// sequentially by one warp, in row order.
//
//   LEX : stage = wavefront of the dependency DAG of the reference ordering;
//         the arithmetic is the reference's, row for row.
//   MC  : stage = colour of a greedy colouring of the group graph; this is the
//         same algorithm applied to a symmetric permutation of the level.
//
// In both modes the level is then permuted symmetrically so that every stage
// is a contiguous row range (perm[new] = old); within a stage groups keep
// their original relative order, so strictly-lower / strictly-upper parts of
// the permuted operator are exactly the coupled-earlier / coupled-later sets.
#pragma once
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace ddpca {

// std::vector whose resize() leaves new elements uninitialised: the multi-GB index / value arrays of the set-up are
// written once, in full, by parallel loops -- a serial zero-fill (and its page faults) before that is pure overhead.
template <class T>
struct DefaultInitAlloc : std::allocator<T> {
    template <class U> struct rebind { using other = DefaultInitAlloc<U>; };
    DefaultInitAlloc() = default;
    template <class U> DefaultInitAlloc(const DefaultInitAlloc<U> &) noexcept {}
    template <class U> void construct(U *p) noexcept { ::new (static_cast<void *>(p)) U; }
    template <class U, class... A> void construct(U *p, A &&...a) { ::new (static_cast<void *>(p)) U(std::forward<A>(a)...); }
};
template <class T> using RawVec = std::vector<T, DefaultInitAlloc<T>>;

struct CsrHost {
    int rows = 0, cols = 0;
    std::vector<int> rp;
    RawVec<int> ci;
    RawVec<double> v;
    long nnz() const { return (long)ci.size(); }
};

// A block-diagonal operator given by its blocks (the caller's CSR arrays, not owned): block s holds the rows
// [row_off[s], row_off[s+1]) and the columns [col_off[s], col_off[s+1]) of the whole, with block-local column indices.
struct CsrBlocks {
    int nsub = 0;
    std::vector<const int *> rp, ci;
    std::vector<const double *> v;
    std::vector<int> row_off, col_off;   // [nsub + 1]
    int rows() const { return row_off[nsub]; }
    int cols() const { return col_off[nsub]; }
    long nnz() const { long t = 0; for (int s = 0; s < nsub; s++) t += rp[s][row_off[s + 1] - row_off[s]]; return t; }
    static CsrBlocks single(int rows, int cols, const int *rp_, const int *ci_, const double *v_)
    {
        CsrBlocks b;
        b.nsub = 1; b.rp = {rp_}; b.ci = {ci_}; b.v = {v_}; b.row_off = {0, rows}; b.col_off = {0, cols};
        return b;
    }
};

struct LevelPlan {
    int n = 0;
    std::vector<int> perm;         // perm[new] = old
    std::vector<int> iperm;        // iperm[old] = new
    std::vector<int> group_start;  // [ngroups+1] first NEW row of each group
    std::vector<int> stage_group;  // [nstages+1] first group of each stage
    int ngroups() const { return (int)group_start.size() - 1; }
    int nstages() const { return (int)stage_group.size() - 1; }
};

// Build the plan from the sparsity pattern alone.  mode: 0 LEX, 1 MC, -1 identity
// (no smoothing on this level: one stage, groups of one row, perm = identity).
// Returns false and fills err on malformed input (unsorted columns, missing diagonal).
bool build_level_plan(int n, const int *rp, const int *ci, int mode, LevelPlan &out, std::string &err);

// The same for a block-diagonal level whose blocks are the row ranges [sub_off[s], sub_off[s+1]): blocks are planned
// in parallel and merged; identical to build_level_plan on the whole level.
bool build_level_plan_blocks(int n, const int *rp, const int *ci, int mode, const std::vector<int> &sub_off, LevelPlan &out, std::string &err);

// The same with the blocks given separately (square blocks): nothing is concatenated or copied.
bool build_level_plan_subs(const CsrBlocks &B, int mode, LevelPlan &out, std::string &err);

// LEX plan of the unit triangular factors I + L and I + L^T of a sparse LDL^T (L strictly lower, CSR): the
// dependency wavefronts, single-row groups; equal to build_level_plan(..., LEX, ...) on either operator.
void build_tri_plan(int n, const int *Lrp, const int *Lci, LevelPlan &out);

// B = Pr * A * Pc^T with sorted columns: row `i` of B is row prow[i] of A, column j of A
// becomes icol[j].  prow has B.rows entries, icol has A.cols entries.
void permute_csr(int rows, int cols, const int *rp, const int *ci, const double *v,
                 const std::vector<int> &prow, const std::vector<int> &icol, CsrHost &out);

// The same for a block-diagonal A given by its blocks, without forming A.
void permute_csr_blocks(const CsrBlocks &B, const std::vector<int> &prow, const std::vector<int> &icol, CsrHost &out);

// out = A^T (sorted columns)
void transpose_csr(const CsrHost &A, CsrHost &out);
// Removes stored entries whose value is exactly 0.0 (in place); returns how many were dropped.
long drop_zeros_csr(CsrHost &A);

}  // namespace ddpca
