// plan_sanitize.cpp -- TEST INFRASTRUCTURE (CPU).  Runs the host planning code of the device layout
// (ddpca-admm_b200/csrc/plan.cpp: row groups, LEX wavefronts / MC colours, block-wise planning, triangular-factor plans,
// symmetric permutation, transpose, zero compaction) over every level of a DDPK hierarchy dump, compiled with
// -fsanitize=address,undefined by tests/test_plan.py: out-of-bounds accesses, signed overflow or misaligned loads in the
// set-up path show up here instead of as a corrupted device layout.  Also re-checks the plan's invariants natively.
//   plan_sanitize file.ddpk   (uncompressed; written by oracle/ref_drivers/ddpk_io.h or ddpca_b200.ddpk.save)
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../ddpca-admm_b200/csrc/plan.h"

using namespace ddpca;

struct Rec { uint32_t dtype; uint64_t count; std::vector<char> data; };

static bool read_ddpk(const char *path, std::map<std::string, Rec> &out)
{
    FILE *f = std::fopen(path, "rb");
    if (!f) return false;
    char magic[8];
    if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "DDPK0001", 8) != 0) { std::fclose(f); return false; }
    size_t off = 8;
    for (;;) {
        uint32_t nl;
        if (std::fread(&nl, 4, 1, f) != 1) break;
        std::string name(nl, '\0');
        Rec r;
        if (std::fread(&name[0], 1, nl, f) != nl || std::fread(&r.dtype, 4, 1, f) != 1 || std::fread(&r.count, 8, 1, f) != 1) { std::fclose(f); return false; }
        off += 4 + nl + 4 + 8;
        const size_t pad = (8 - off % 8) % 8;
        char z[8];
        if (pad && std::fread(z, 1, pad, f) != pad) { std::fclose(f); return false; }
        off += pad;
        const size_t bytes = r.count * (r.dtype == 1 ? 4 : 8);
        r.data.resize(bytes);
        if (bytes && std::fread(r.data.data(), 1, bytes, f) != bytes) { std::fclose(f); return false; }
        off += bytes;
        out[name] = std::move(r);
    }
    std::fclose(f);
    return true;
}

static int check_plan(int n, const int *rp, const int *ci, const LevelPlan &pl, const char *what)
{
    if ((int)pl.perm.size() != n || (int)pl.iperm.size() != n) { std::fprintf(stderr, "%s: perm size\n", what); return 1; }
    for (int i = 0; i < n; i++) if (pl.perm[i] < 0 || pl.perm[i] >= n || pl.iperm[pl.perm[i]] != i) { std::fprintf(stderr, "%s: perm not a bijection\n", what); return 1; }
    if (pl.group_start.front() != 0 || pl.group_start.back() != n || pl.stage_group.front() != 0 || pl.stage_group.back() != pl.ngroups()) { std::fprintf(stderr, "%s: tables\n", what); return 1; }
    // no coupling between different groups of one stage
    std::vector<int> group_of(n), stage_of(pl.ngroups());
    for (int g = 0; g < pl.ngroups(); g++) for (int i = pl.group_start[g]; i < pl.group_start[g + 1]; i++) group_of[i] = g;
    for (int s = 0; s < pl.nstages(); s++) for (int g = pl.stage_group[s]; g < pl.stage_group[s + 1]; g++) stage_of[g] = s;
    for (int i = 0; i < n; i++) {
        const int old = pl.perm[i];
        for (int p = rp[old]; p < rp[old + 1]; p++) {
            const int j = pl.iperm[ci[p]];
            if (group_of[i] != group_of[j] && stage_of[group_of[i]] == stage_of[group_of[j]]) { std::fprintf(stderr, "%s: rows %d and %d coupled inside stage %d\n", what, i, j, stage_of[group_of[i]]); return 1; }
        }
    }
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: plan_sanitize file.ddpk\n"); return 2; }
    std::map<std::string, Rec> d;
    if (!read_ddpk(argv[1], d)) { std::fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    const long L = *reinterpret_cast<const long *>(d.at("maxiLeve").data.data());
    long rows_total = 0;
    for (long l = 0; l <= L; l++) {
        const std::string k = "consStif" + std::to_string(l);
        const long *shape = reinterpret_cast<const long *>(d.at(k + ".shape").data.data());
        const int n = (int)shape[0];
        const int *rp = reinterpret_cast<const int *>(d.at(k + ".rowptr").data.data());
        const int *ci = reinterpret_cast<const int *>(d.at(k + ".colidx").data.data());
        const double *v = reinterpret_cast<const double *>(d.at(k + ".val").data.data());
        std::string err;
        for (int mode = 0; mode <= 1; mode++) {
            LevelPlan pl;
            if (!build_level_plan(n, rp, ci, mode, pl, err)) { std::fprintf(stderr, "level %ld mode %d: %s\n", l, mode, err.c_str()); return 1; }
            if (check_plan(n, rp, ci, pl, mode ? "MC" : "LEX")) return 1;
            // the same level as one block, and as a 1-block "batch"
            LevelPlan pb, ps;
            std::vector<int> off = {0, n};
            if (!build_level_plan_blocks(n, rp, ci, mode, off, pb, err) || pb.perm != pl.perm || pb.stage_group != pl.stage_group) { std::fprintf(stderr, "level %ld: block plan differs\n", l); return 1; }
            CsrBlocks B = CsrBlocks::single(n, n, rp, ci, v);
            if (!build_level_plan_subs(B, mode, ps, err) || ps.perm != pl.perm || ps.group_start != pl.group_start) { std::fprintf(stderr, "level %ld: subs plan differs\n", l); return 1; }
            CsrHost Ap, Ab, At, Att;
            permute_csr(n, n, rp, ci, v, pl.perm, pl.iperm, Ap);
            permute_csr_blocks(B, pl.perm, pl.iperm, Ab);
            if (Ap.rp != Ab.rp || Ap.nnz() != Ab.nnz() || std::memcmp(Ap.ci.data(), Ab.ci.data(), sizeof(int) * Ap.nnz()) || std::memcmp(Ap.v.data(), Ab.v.data(), sizeof(double) * Ap.nnz())) { std::fprintf(stderr, "level %ld: permute_csr_blocks differs\n", l); return 1; }
            transpose_csr(Ap, At);
            transpose_csr(At, Att);
            if (Att.rp != Ap.rp || std::memcmp(Att.ci.data(), Ap.ci.data(), sizeof(int) * Ap.nnz()) || std::memcmp(Att.v.data(), Ap.v.data(), sizeof(double) * Ap.nnz())) { std::fprintf(stderr, "level %ld: transpose twice differs\n", l); return 1; }
            const long before = Ap.nnz();
            long zeros = 0;
            for (long p = 0; p < before; p++) zeros += (Ap.v[p] == 0.0);
            if (drop_zeros_csr(Ap) != zeros || Ap.nnz() != before - zeros) { std::fprintf(stderr, "level %ld: drop_zeros\n", l); return 1; }
        }
        {   // strictly-lower part as a triangular factor pattern: wavefront plan
            std::vector<int> lrp(n + 1, 0), lci;
            for (int i = 0; i < n; i++) { for (int p = rp[i]; p < rp[i + 1]; p++) if (ci[p] < i) lci.push_back(ci[p]); lrp[i + 1] = (int)lci.size(); }
            LevelPlan pt;
            build_tri_plan(n, lrp.data(), lci.data(), pt);
            if ((int)pt.perm.size() != n) { std::fprintf(stderr, "level %ld: tri plan\n", l); return 1; }
        }
        if (l < L) {   // prolongation: rectangular permutation + transpose + zero compaction
            const std::string kp = "realProl" + std::to_string(l);
            const long *ps = reinterpret_cast<const long *>(d.at(kp + ".shape").data.data());
            const int *prp = reinterpret_cast<const int *>(d.at(kp + ".rowptr").data.data());
            const int *pci = reinterpret_cast<const int *>(d.at(kp + ".colidx").data.data());
            const double *pv = reinterpret_cast<const double *>(d.at(kp + ".val").data.data());
            std::vector<int> idr((size_t)ps[0]), idc((size_t)ps[1]);
            for (size_t i = 0; i < idr.size(); i++) idr[i] = (int)(idr.size() - 1 - i);
            for (size_t i = 0; i < idc.size(); i++) idc[i] = (int)(idc.size() - 1 - i);
            CsrHost Pp, Rt;
            permute_csr((int)ps[0], (int)ps[1], prp, pci, pv, idr, idc, Pp);
            drop_zeros_csr(Pp);
            transpose_csr(Pp, Rt);
            if (Rt.rows != (int)ps[1] || Rt.nnz() != Pp.nnz()) { std::fprintf(stderr, "level %ld: transfer transpose\n", l); return 1; }
        }
        rows_total += n;
    }
    std::printf("OK levels %ld rows %ld\n", L + 1, rows_total);
    return 0;
}
