// plan.cpp -- see plan.h.  Pure host C++, no CUDA.
#include "plan.h"

#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <cstring>
#include <numeric>

namespace ddpca {

static const int kMaxGroup = 3;

bool build_level_plan(int n, const int *rp, const int *ci, int mode, LevelPlan &out, std::string &err)
{
    out = LevelPlan();
    out.n = n;
    if (mode < 0) {
        out.perm.resize(n);
        std::iota(out.perm.begin(), out.perm.end(), 0);
        out.iperm = out.perm;
        out.group_start.resize(n + 1);
        std::iota(out.group_start.begin(), out.group_start.end(), 0);
        out.stage_group = {0, n};
        return true;
    }
    // ---- 0. sanity: sorted columns, diagonal present -------------------------------
    for (int i = 0; i < n; i++) {
        bool diag = false;
        for (int p = rp[i]; p < rp[i + 1]; p++) {
            if (p > rp[i] && ci[p] <= ci[p - 1]) { err = "columns not strictly increasing in row " + std::to_string(i); return false; }
            if (ci[p] < 0 || ci[p] >= n) { err = "column index out of range in row " + std::to_string(i); return false; }
            if (ci[p] == i) diag = true;
        }
        if (!diag) { err = "no diagonal entry in row " + std::to_string(i) + " (Gauss-Seidel needs one)"; return false; }
    }
    // ---- 1. groups: consecutive rows with one column pattern, at most 3 ------------
    std::vector<int> gfirst;   // first OLD row of each group
    std::vector<int> gid(n);   // group of each OLD row
    for (int i = 0; i < n;) {
        int len = rp[i + 1] - rp[i];
        int j = i + 1;
        while (j < n && j - i < kMaxGroup && rp[j + 1] - rp[j] == len &&
               std::memcmp(ci + rp[i], ci + rp[j], sizeof(int) * len) == 0)
            j++;
        // every row of the group holds its own diagonal, hence (identical patterns) all
        // in-group columns i..j-1 are present in each row.
        for (int k = i; k < j; k++) gid[k] = (int)gfirst.size();
        gfirst.push_back(i);
        i = j;
    }
    int ng = (int)gfirst.size();
    gfirst.push_back(n);
    // ---- 2. group graph (from the first row of each group), symmetrised ------------
    std::vector<int> adj_ptr(ng + 1, 0);
    std::vector<int> adj;
    // direct adjacency: columns are sorted and groups are contiguous, so the group ids met
    // along a row are non-decreasing and de-duplicate by comparing with the previous one.
    for (int g = 0; g < ng; g++) {
        int i = gfirst[g];
        int last = -1, cnt = 0;
        for (int p = rp[i]; p < rp[i + 1]; p++) {
            int h = gid[ci[p]];
            if (h != last) { last = h; if (h != g) cnt++; }
        }
        adj_ptr[g + 1] = adj_ptr[g] + cnt;
    }
    adj.resize(adj_ptr[ng]);
    for (int g = 0; g < ng; g++) {
        int i = gfirst[g];
        int last = -1, q = adj_ptr[g];
        for (int p = rp[i]; p < rp[i + 1]; p++) {
            int h = gid[ci[p]];
            if (h != last) { last = h; if (h != g) adj[q++] = h; }
        }
    }
    // FE stiffness patterns are structurally symmetric; verify, and only symmetrise if not.
    bool symmetric = true;
    for (int g = 0; g < ng && symmetric; g++)
        for (int p = adj_ptr[g]; p < adj_ptr[g + 1]; p++) {
            int h = adj[p];
            if (!std::binary_search(adj.begin() + adj_ptr[h], adj.begin() + adj_ptr[h + 1], g)) { symmetric = false; break; }
        }
    if (!symmetric) {
        std::vector<std::pair<int, int>> edges;  // (g,h), g != h
        edges.reserve(adj.size() * 2);
        for (int g = 0; g < ng; g++)
            for (int p = adj_ptr[g]; p < adj_ptr[g + 1]; p++) { edges.emplace_back(g, adj[p]); edges.emplace_back(adj[p], g); }
        std::fill(adj_ptr.begin(), adj_ptr.end(), 0);
        std::sort(edges.begin(), edges.end());
        edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
        for (auto &e : edges) adj_ptr[e.first + 1]++;
        for (int g = 0; g < ng; g++) adj_ptr[g + 1] += adj_ptr[g];
        adj.resize(edges.size());
        std::vector<int> fill(adj_ptr.begin(), adj_ptr.end() - 1);
        for (auto &e : edges) adj[fill[e.first]++] = e.second;
    }
    // ---- 3. stage of each group ----------------------------------------------------
    std::vector<int> stage(ng, 0);
    int nstages = 0;
    if (mode == 0) {
        // LEX: wavefront level in the reference ordering
        for (int g = 0; g < ng; g++) {
            int lv = 0;
            for (int p = adj_ptr[g]; p < adj_ptr[g + 1]; p++) {
                int h = adj[p];
                if (h < g) lv = std::max(lv, stage[h] + 1);
            }
            stage[g] = lv;
            nstages = std::max(nstages, lv + 1);
        }
    } else {
        // MC: greedy colouring in the reference ordering (smallest colour unused by neighbours)
        std::vector<int> mark;
        for (int g = 0; g < ng; g++) {
            int deg = adj_ptr[g + 1] - adj_ptr[g];
            if ((int)mark.size() < deg + 2) mark.resize(deg + 2, -1);
            for (int p = adj_ptr[g]; p < adj_ptr[g + 1]; p++) {
                int h = adj[p];
                if (h < g && stage[h] <= deg) mark[stage[h]] = g;
            }
            int c = 0;
            while (c <= deg && mark[c] == g) c++;
            stage[g] = c;
            nstages = std::max(nstages, c + 1);
        }
    }
    // ---- 4. order groups by (stage, reference order): counting sort -----------------
    out.stage_group.assign(nstages + 1, 0);
    for (int g = 0; g < ng; g++) out.stage_group[stage[g] + 1]++;
    for (int s = 0; s < nstages; s++) out.stage_group[s + 1] += out.stage_group[s];
    std::vector<int> gorder(ng);
    {
        std::vector<int> fill(out.stage_group.begin(), out.stage_group.end() - 1);
        for (int g = 0; g < ng; g++) gorder[fill[stage[g]]++] = g;
    }
    out.perm.resize(n);
    out.iperm.resize(n);
    out.group_start.resize(ng + 1);
    int row = 0;
    for (int k = 0; k < ng; k++) {
        int g = gorder[k];
        out.group_start[k] = row;
        for (int i = gfirst[g]; i < gfirst[g + 1]; i++) {
            out.perm[row] = i;
            out.iperm[i] = row;
            row++;
        }
    }
    out.group_start[ng] = n;
    return true;
}

// Block-diagonal level (a batch of subdomains): the blocks are planned independently, in parallel, and merged --
// stage s of the level is stage s of every block, block after block.  This is exactly the plan of the whole level
// (the greedy colouring / the wavefronts of a block-diagonal pattern do not see the other blocks, and groups keep
// their original order inside a stage), obtained nsub times faster.
static bool merge_sub_plans(int n, const std::vector<int> &sub_off, const std::vector<LevelPlan> &part, LevelPlan &out)
{
    const int nsub = (int)part.size();
    int nstages = 0;
    long ng = 0;
    for (const LevelPlan &p : part) { nstages = std::max(nstages, p.nstages()); ng += p.ngroups(); }
    out = LevelPlan();
    out.n = n;
    out.perm.resize(n);
    out.iperm.resize(n);
    out.group_start.reserve(ng + 1);
    out.stage_group.assign(nstages + 1, 0);
    int row = 0;
    for (int st = 0; st < nstages; st++) {
        out.stage_group[st] = (int)out.group_start.size();
        for (int s = 0; s < nsub; s++) {
            const LevelPlan &p = part[s];
            if (st >= p.nstages()) continue;
            const int r0 = sub_off[s];
            for (int g = p.stage_group[st]; g < p.stage_group[st + 1]; g++) {
                out.group_start.push_back(row);
                for (int i = p.group_start[g]; i < p.group_start[g + 1]; i++) {
                    const int old = p.perm[i] + r0;
                    out.perm[row] = old;
                    out.iperm[old] = row;
                    row++;
                }
            }
        }
    }
    out.stage_group[nstages] = (int)out.group_start.size();
    out.group_start.push_back(n);
    return true;
}

bool build_level_plan_blocks(int n, const int *rp, const int *ci, int mode, const std::vector<int> &sub_off, LevelPlan &out, std::string &err)
{
    const int nsub = (int)sub_off.size() - 1;
    if (mode < 0 || nsub <= 1) return build_level_plan(n, rp, ci, mode, out, err);
    std::vector<LevelPlan> part(nsub);
    std::vector<std::string> errs(nsub);
    std::vector<char> ok(nsub, 1);
#pragma omp parallel for schedule(dynamic, 1)
    for (int s = 0; s < nsub; s++) {
        const int r0 = sub_off[s], nr = sub_off[s + 1] - r0;
        const int p0 = rp[r0];
        std::vector<int> lrp(nr + 1), lci((size_t)(rp[r0 + nr] - p0));
        for (int i = 0; i <= nr; i++) lrp[i] = rp[r0 + i] - p0;
        bool inside = true;
        for (size_t p = 0; p < lci.size(); p++) { lci[p] = ci[p0 + p] - r0; inside = inside && lci[p] >= 0 && lci[p] < nr; }
        if (!inside) { ok[s] = 0; errs[s] = "subdomain " + std::to_string(s) + ": the level is not block diagonal"; continue; }
        if (!build_level_plan(nr, lrp.data(), lci.data(), mode, part[s], errs[s])) ok[s] = 0;
    }
    for (int s = 0; s < nsub; s++) if (!ok[s]) { err = "subdomain " + std::to_string(s) + ": " + errs[s]; return false; }
    return merge_sub_plans(n, sub_off, part, out);
}

bool build_level_plan_subs(const CsrBlocks &B, int mode, LevelPlan &out, std::string &err)
{
    const int nsub = B.nsub, n = B.rows();
    if (nsub == 1) return build_level_plan(n, B.rp[0], B.ci[0], mode, out, err);
    if (mode < 0) {   // identity plan of the whole level: one stage, single-row groups (level 0)
        out = LevelPlan();
        out.n = n;
        out.perm.resize(n); out.iperm.resize(n); out.group_start.resize(n + 1);
        for (int i = 0; i < n; i++) out.perm[i] = out.iperm[i] = out.group_start[i] = i;
        out.group_start[n] = n;
        out.stage_group = {0, n};
        return true;
    }
    std::vector<LevelPlan> part(nsub);
    std::vector<std::string> errs(nsub);
    std::vector<char> ok(nsub, 1);
#pragma omp parallel for schedule(dynamic, 1)
    for (int s = 0; s < nsub; s++) {
        const int nr = B.row_off[s + 1] - B.row_off[s];
        if (!build_level_plan(nr, B.rp[s], B.ci[s], mode, part[s], errs[s])) ok[s] = 0;
    }
    for (int s = 0; s < nsub; s++) if (!ok[s]) { err = "subdomain " + std::to_string(s) + ": " + errs[s]; return false; }
    return merge_sub_plans(n, B.row_off, part, out);
}

// Wavefront plan of a unit triangular factor I + L (and of I + L^T, which has the same one): single-row groups,
// stage[i] = 1 + max stage[j] over the strictly-lower couplings j of row i.  This is what build_level_plan(LEX)
// returns for these operators, without building and symmetrising a 2 x nnz edge list.
void build_tri_plan(int n, const int *Lrp, const int *Lci, LevelPlan &out)
{
    out = LevelPlan();
    out.n = n;
    std::vector<int> stage(n, 0);
    int nstages = n ? 1 : 0;
    for (int i = 0; i < n; i++) {
        int lv = 0;
        for (int p = Lrp[i]; p < Lrp[i + 1]; p++) lv = std::max(lv, stage[Lci[p]] + 1);
        stage[i] = lv;
        nstages = std::max(nstages, lv + 1);
    }
    out.stage_group.assign(nstages + 1, 0);
    for (int i = 0; i < n; i++) out.stage_group[stage[i] + 1]++;
    for (int s = 0; s < nstages; s++) out.stage_group[s + 1] += out.stage_group[s];
    out.perm.resize(n);
    out.iperm.resize(n);
    std::vector<int> fill(out.stage_group.begin(), out.stage_group.end() - 1);
    for (int i = 0; i < n; i++) { const int r = fill[stage[i]]++; out.perm[r] = i; out.iperm[i] = r; }
    out.group_start.resize(n + 1);
    std::iota(out.group_start.begin(), out.group_start.end(), 0);
}

void permute_csr_blocks(const CsrBlocks &B, const std::vector<int> &prow, const std::vector<int> &icol, CsrHost &out)
{
    const int rows = B.rows(), nsub = B.nsub;
    out.rows = rows;
    out.cols = (int)icol.size();
    out.rp.assign(rows + 1, 0);
    // block of every output row (prow[i] = row of the whole operator)
    auto sub_of = [&](int old) { return nsub == 1 ? 0 : (int)(std::upper_bound(B.row_off.begin(), B.row_off.end(), old) - B.row_off.begin()) - 1; };
    for (int i = 0; i < rows; i++) {
        const int old = prow[i], s = sub_of(old), lr = old - B.row_off[s];
        out.rp[i + 1] = out.rp[i] + (B.rp[s][lr + 1] - B.rp[s][lr]);
    }
    const long nnz = out.rp[rows];
    out.ci.resize(nnz);
    out.v.resize(nnz);
#pragma omp parallel
    {
        // keys = (new column << 32 | position in the row): one integer sort per row, the values follow by position
        std::vector<unsigned long long> key;
#pragma omp for schedule(dynamic, 1024)
        for (int i = 0; i < rows; i++) {
            const int old = prow[i], s = sub_of(old), lr = old - B.row_off[s];
            const int p0 = B.rp[s][lr], len = B.rp[s][lr + 1] - p0;
            const int *cs = B.ci[s] + p0;
            const double *vs = B.v[s] + p0;
            const int coff = B.col_off[s];
            key.resize(len);
            bool sorted = true;
            for (int k = 0; k < len; k++) {
                key[k] = ((unsigned long long)(unsigned)icol[cs[k] + coff] << 32) | (unsigned)k;
                sorted = sorted && (k == 0 || key[k - 1] < key[k]);
            }
            if (!sorted) std::sort(key.begin(), key.end());
            const int base = out.rp[i];
            for (int k = 0; k < len; k++) { out.ci[base + k] = (int)(key[k] >> 32); out.v[base + k] = vs[(unsigned)key[k]]; }
        }
    }
}

void permute_csr(int rows, int cols, const int *rp, const int *ci, const double *v,
                 const std::vector<int> &prow, const std::vector<int> &icol, CsrHost &out)
{
    permute_csr_blocks(CsrBlocks::single(rows, cols, rp, ci, v), prow, icol, out);
}

void transpose_csr(const CsrHost &A, CsrHost &out)
{
    out.rows = A.cols;
    out.cols = A.rows;
    const long nnz = A.nnz();
    const int ncol = A.cols;
    int T = 1;
#ifdef _OPENMP
    T = omp_get_max_threads();
#endif
    // Row ranges of A go to threads; each counts its entries per column, a prefix over (column, thread) turns the
    // counts into write positions, each thread then scatters its rows in order: columns of the result come out sorted
    // and the result does not depend on the thread count.  The histograms cost T * ncol ints: only when the operator
    // is large enough to pay for them.
    if (T > 1 && nnz >= (1L << 16) && (long)T * ncol <= 4 * nnz) {
        std::vector<int> rbeg(T + 1);
        for (int t = 0; t <= T; t++) rbeg[t] = (int)((long)A.rows * t / T);
        RawVec<int> cnt((size_t)T * ncol);
#pragma omp parallel num_threads(T)
        {
            const int t = omp_get_thread_num();
            int *c = cnt.data() + (size_t)t * ncol;
            std::memset(c, 0, sizeof(int) * (size_t)ncol);
            for (long p = A.rp[rbeg[t]]; p < A.rp[rbeg[t + 1]]; p++) c[A.ci[p]]++;
        }
        out.rp.assign(ncol + 1, 0);
#pragma omp parallel for schedule(static)
        for (int j = 0; j < ncol; j++) {
            int tot = 0;
            for (int t = 0; t < T; t++) tot += cnt[(size_t)t * ncol + j];
            out.rp[j + 1] = tot;
        }
        for (int j = 0; j < ncol; j++) out.rp[j + 1] += out.rp[j];
#pragma omp parallel for schedule(static)
        for (int j = 0; j < ncol; j++) {
            int run = out.rp[j];
            for (int t = 0; t < T; t++) { const int c = cnt[(size_t)t * ncol + j]; cnt[(size_t)t * ncol + j] = run; run += c; }
        }
        out.ci.resize(nnz);
        out.v.resize(nnz);
#pragma omp parallel num_threads(T)
        {
            const int t = omp_get_thread_num();
            int *pos = cnt.data() + (size_t)t * ncol;
            for (int i = rbeg[t]; i < rbeg[t + 1]; i++)
                for (int p = A.rp[i]; p < A.rp[i + 1]; p++) {
                    const int q = pos[A.ci[p]]++;
                    out.ci[q] = i;
                    out.v[q] = A.v[p];
                }
        }
        return;
    }
    out.rp.assign(ncol + 1, 0);
    for (long p = 0; p < nnz; p++) out.rp[A.ci[p] + 1]++;
    for (int j = 0; j < ncol; j++) out.rp[j + 1] += out.rp[j];
    out.ci.resize(nnz);
    out.v.resize(nnz);
    std::vector<int> fill(out.rp.begin(), out.rp.end() - 1);
    for (int i = 0; i < A.rows; i++)
        for (int p = A.rp[i]; p < A.rp[i + 1]; p++) {
            int q = fill[A.ci[p]]++;
            out.ci[q] = i;   // rows visited in increasing order => sorted columns
            out.v[q] = A.v[p];
        }
}

long drop_zeros_csr(CsrHost &A)
{
    const long nnz = A.nnz();
    const int n = A.rows;
    // count per row, prefix, compact into fresh arrays (parallel over rows)
    std::vector<int> nrp(n + 1, 0);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        int c = 0;
        for (int p = A.rp[i]; p < A.rp[i + 1]; p++) c += (A.v[p] != 0.0);
        nrp[i + 1] = c;
    }
    for (int i = 0; i < n; i++) nrp[i + 1] += nrp[i];
    const long w = nrp[n];
    if (w == nnz) return 0;
    RawVec<int> ci(w);
    RawVec<double> v(w);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        int q = nrp[i];
        for (int p = A.rp[i]; p < A.rp[i + 1]; p++)
            if (A.v[p] != 0.0) { ci[q] = A.ci[p]; v[q] = A.v[p]; q++; }
    }
    A.rp.swap(nrp);
    A.ci.swap(ci);
    A.v.swap(v);
    return nnz - w;
}

}  // namespace ddpca
