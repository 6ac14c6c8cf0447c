// mg.cu -- device-resident multigrid hierarchy and the MG-PCG driver behind the C ABI
// of include/ddpca_b200.h (class MGPIS of the reference, MGPIS.h:8-225).
//
// Layout in HBM, per level l (all in the stage-permuted numbering of plan.h):
//   A_l   group layout ("GCSR", kernels.cuh): one column pattern per row group (<= 3 rows of
//         one mesh node) + FP64 values per row.  The reference keeps three copies
//         (consLowe/consDiag/consUppe, MGPIS.h:29-33); here the strictly-lower / in-group /
//         strictly-upper parts are position ranges [0,kd) [kd,kd+gs) [kd+gs,len) of a pattern.
//   P_l-1 realProl[l-1] as CSR (n_l x n_l-1) and its explicit transpose R (gather-based
//         restriction, no atomics).
//   x,b,p1,r work vectors.   Level 0 additionally holds the dense inverse of consStif[0].
// Scalars of the CG recurrence live in a PcgState in HBM; one CG iteration is one CUDA
// graph launch; the host only polls a `done` flag with a lag (no per-iteration sync).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <functional>
#include <string>
#include <vector>

#include "../../include/ddpca_b200.h"
#include "kernels.cuh"
#include "kernels2.cuh"
#include "plan.h"

using namespace ddpca;

static thread_local std::string g_err;
static int fail(const std::string &m) { g_err = m; return 1; }

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" +  \
                    std::to_string(__LINE__) + ")";                                               \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

// inside the host-driven solvers: error exit through the function's cleanup() lambda
#define CUX(call)                                                                                 \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" +  \
                    std::to_string(__LINE__) + ")";                                               \
            cleanup();                                                                            \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

namespace {

struct DevCsr {
    int rows = 0, cols = 0;
    long nnz = 0;
    int *rp = nullptr, *ci = nullptr;
    double *v = nullptr;
    CsrView view() const { return CsrView{rows, rp, ci, v}; }
    double avg_row() const { return rows ? (double)nnz / rows : 0.0; }
};

// transfer operator in node-triple form (kernels.cuh, k_trip_spmv) + CSR remainder
struct DevTrip {
    int rows = 0, cols = 0, ntrip = 0, nrest = 0;
    long npairs = 0, nnz_rest = 0;
    int *trow = nullptr, *tptr = nullptr, *tcol = nullptr, *rest_row = nullptr;
    double *tw = nullptr;
    DevCsr rest;
    bool ok = false;        // built (otherwise the plain CSR operator is used)
    double avg_pairs() const { return ntrip ? (double)npairs / ntrip : 0.0; }
};

struct Segment {
    int multi;   // 0: one stage, one launch, one warp per group; 1: run of stages in one CTA
    int s0, s1;  // stage range
    int g0, g1;  // group range
    double bytes_lo, bytes_up;  // algorithmic bytes of the strictly-lower / -upper halves + vectors
};

struct Level {
    int n = 0;
    long nnz = 0;          // stored entries of consStif[l] (reference count, without padding)
    LevelPlan plan;
    DevCsr A;              // plain CSR, kept for level 0 only (input of the dense inversion)
    // group layout (kernels.cuh: GroupMeta / LvlView)
    int ng = 0;
    GroupMeta *meta = nullptr;
    int *gci = nullptr;
    double *gv = nullptr;
    long pat_entries = 0;  // pattern entries stored (sum of padded len over groups)
    long val_entries = 0;  // values stored (sum of gs * padded len)
    int *stage_group = nullptr, *perm = nullptr;
    std::vector<Segment> segs;
    bool wide_rows = false;   // triangular factors: one full warp per row in the single-CTA stage runs
    DevCsr P, R;  // level l <-> l-1 (l >= 1)
    DevTrip Pt, Rt;  // the same in node-triple form when the operators have that structure
    double *x = nullptr, *b = nullptr, *p1 = nullptr, *r = nullptr, *dinv = nullptr;
    double bytes_lower = 0, bytes_upper = 0, bytes_full = 0;  // algorithmic bytes of one pass over a half / the whole level
    LvlView view() const { return LvlView{n, ng, meta, gci, gv}; }
    // v2: split lower/upper storage + chunk table (kernels2.cuh); used when every stage is large
    bool v2 = false;
    GroupMeta2 *meta2 = nullptr;
    int *CL = nullptr, *CU = nullptr;
    double *VL = nullptr, *VU = nullptr, *BD = nullptr;
    // one chunk table per pass type (V2_TAB_FULL / _LO / _UP): the half passes cut their chunks by bytes
    ChunkDesc *chunks[3] = {nullptr, nullptr, nullptr};
    int *stage_chunk[3] = {nullptr, nullptr, nullptr};
    int nchunks[3] = {0, 0, 0}, max_stage_chunks[3] = {0, 0, 0};
    size_t buf[3] = {0, 0, 0};   // shared-memory bytes of one chunk buffer
    ChunkDesc *empty_desc = nullptr;   // one zero descriptor (ng = 0) standing in for chunks of frozen subdomains
    const PcgState *sub_state = nullptr;   // per-sub CG states of the owning batch when skipping is on (not owned)
    unsigned *gbar = nullptr;   // counters of the consumer grid barrier
    int fuse_bwd_last = 0;       // the forward sweeps also do the backward step of the last colour (see Lvl2View)
    double bytes_bwd_skip = 0;   // algorithmic bytes the backward sweep no longer touches then
    Lvl2View view2(int tab = V2_TAB_FULL) const { return Lvl2View{n, ng, nchunks[tab], plan.nstages(), fuse_bwd_last, meta2, CL, VL, CU, VU, BD, chunks[tab], stage_chunk[tab], empty_desc, sub_state}; }
};

struct ProfRec {
    int kclass, level;
    double bytes;
    cudaEvent_t a, b;
};

}  // namespace

// Launch context shared by every handle type: device, stream, launch accounting, per-class timing.
struct Engine {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    int sms = 148;
    long launches = 0;
    bool capturing = false;
    long captured_nodes = 0;
    bool profile = false;
    std::string launch_err;   // first failed launch since the last check (cooperative launches report through their return code)
    std::vector<ProfRec> prof;
    double prof_ms[DDPCA_K_COUNT][16];
    long prof_n[DDPCA_K_COUNT][16];
    double prof_bytes[DDPCA_K_COUNT][16];
    Engine() { prof_reset(); }
    void prof_reset()
    {
        std::memset(prof_ms, 0, sizeof(prof_ms));
        std::memset(prof_n, 0, sizeof(prof_n));
        std::memset(prof_bytes, 0, sizeof(prof_bytes));
    }
    void pre(int kclass, int level, double bytes)
    {
        if (capturing) { captured_nodes++; return; }
        launches++;
        if (profile) {
            ProfRec r{kclass, level, bytes, nullptr, nullptr};
            cudaEventCreate(&r.a);
            cudaEventCreate(&r.b);
            cudaEventRecord(r.a, stream);
            prof.push_back(r);
        }
    }
    void post()
    {
        if (profile && !capturing) cudaEventRecord(prof.back().b, stream);
    }
    void prof_collect()
    {
        if (prof.empty()) return;
        cudaStreamSynchronize(stream);
        for (auto &r : prof) {
            float ms = 0;
            cudaEventElapsedTime(&ms, r.a, r.b);
            int l = std::min(std::max(r.level, 0), 15);
            prof_ms[r.kclass][l] += ms;
            prof_n[r.kclass][l] += 1;
            prof_bytes[r.kclass][l] += r.bytes;
            cudaEventDestroy(r.a);
            cudaEventDestroy(r.b);
        }
        prof.clear();
    }
};

struct ddpca_mg : Engine {
    int mode = DDPCA_SMOOTH_MC;
    int nlev = 0;
    std::vector<Level> lev;
    // nsub >= 1 independent subdomain hierarchies held as ONE block-diagonal hierarchy (kernels.cuh, "batched PCG"):
    // level kernels work on all of them at once, the CG recurrence is per sub
    int nsub = 1;
    std::vector<std::vector<int>> sub_off;   // [nlev][nsub+1] first row of each sub on each level (reference numbering)
    std::vector<int> sub_n;                  // finest-level rows of each sub
    double *Binv = nullptr;       // dense inverses of the level-0 blocks, sub after sub
    int n0 = 0;                   // rows of level 0 (all subs)
    double coarse_bytes = 0;      // algorithmic bytes of one level-0 solve
    int *sub0_off_d = nullptr;    // [nsub+1] rows of level 0
    double **binv_ptr_d = nullptr;   // [nsub] dense inverse of each sub's level 0
    // segmented reductions on the finest level (device numbering)
    int nseg = 0;
    SegChunk *seg_d = nullptr;
    int *sub_chunk_d = nullptr, *sub_n_d = nullptr;
    // finest-level CG vectors (device numbering) + staging in reference numbering
    double *cg_r = nullptr, *cg_p = nullptr, *cg_q = nullptr, *cg_z = nullptr, *cg_x = nullptr;
    double *stage_a = nullptr, *stage_b = nullptr;  // max-n staging buffers
    PcgState *st = nullptr;       // [nsub]
    PcgState *st_host = nullptr;  // pinned [nsub]
    BatchFlags *fl = nullptr;
    BatchFlags *fl_host = nullptr;   // pinned, ring of kDepth+2
    double *partial[3] = {nullptr, nullptr, nullptr};
    KrylovState *ks = nullptr;    // scalars of MULT_SOLV / GMRES_SOLV / BiCGSTAB_SOLV (allocated on first use)
    cudaGraphExec_t iter_graph[2] = {nullptr, nullptr};  // per preconditioner
    long iter_graph_nodes[2] = {0, 0};
    // whole solve as ONE graph: set-up nodes + a WHILE conditional node around the iteration body
    cudaGraphExec_t solve_graph[2] = {nullptr, nullptr};
    long solve_init_nodes[2] = {0, 0}, solve_iter_nodes[2] = {0, 0};
    int while_state[2] = {0, 0};   // 0 untried, 1 available, -1 unavailable (host-polled loop is used)
    bool pending_while = false;
    int pending_prec = 1;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    double t_solve = 0, t_h2d = 0, t_d2h = 0;
    const int *done_flag() const { return &fl->done_all; }
};

#define KL(h, kc, lvl, bytes, ...) \
    do {                           \
        (h)->pre(kc, lvl, bytes);  \
        __VA_ARGS__;               \
        (h)->post();               \
    } while (0)

static const int kDepth = 2;  // CG iterations kept in flight ahead of the host's done-poll

// ------------------------------------------------------------------------------------------
static int upload_csr(const CsrHost &h, DevCsr &d)
{
    d.rows = h.rows; d.cols = h.cols; d.nnz = h.nnz();
    CU(cudaMalloc(&d.rp, sizeof(int) * (h.rows + 1)));
    CU(cudaMalloc(&d.ci, sizeof(int) * std::max<long>(1, d.nnz)));
    CU(cudaMalloc(&d.v, sizeof(double) * std::max<long>(1, d.nnz)));
    CU(cudaMemcpy(d.rp, h.rp.data(), sizeof(int) * (h.rows + 1), cudaMemcpyHostToDevice));
    if (d.nnz) {
        CU(cudaMemcpy(d.ci, h.ci.data(), sizeof(int) * d.nnz, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(d.v, h.v.data(), sizeof(double) * d.nnz, cudaMemcpyHostToDevice));
    }
    return 0;
}
static void free_csr(DevCsr &d)
{
    cudaFree(d.rp); cudaFree(d.ci); cudaFree(d.v);
    d = DevCsr();
}
template <class T, class Al>
static int upload_vec(const std::vector<T, Al> &h, T **d)
{
    CU(cudaMalloc(d, sizeof(T) * std::max<size_t>(1, h.size())));
    if (!h.empty()) CU(cudaMemcpy(*d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
    return 0;
}

// Work vector of n entries + ONE always-zero slot at [n], all zero-initialised.  The padding entries of
// the group layouts (value 0) point at column n, so a gather never multiplies 0 by stale or non-finite
// memory; nothing ever writes slot n.  The reference starts every solve from fresh zero vectors.
static int alloc_vec(double **d, long n)
{
    CU(cudaMalloc(d, sizeof(double) * (size_t)(n + 1)));
    CU(cudaMemset(*d, 0, sizeof(double) * (size_t)(n + 1)));
    return 0;
}

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }
// DDPCA_VERBOSE: wall-clock stages of the set-up on stderr
struct StageTimer {
    bool on;
    double t0, last;
    const char *what;
    static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
    explicit StageTimer(const char *w) : on(std::getenv("DDPCA_VERBOSE") != nullptr), t0(now()), last(t0), what(w) {}
    void lap(const char *stage) { if (!on) return; double t = now(); std::fprintf(stderr, "ddpca set-up [%s] %-28s %8.3f s\n", what, stage, t - last); last = t; }
    ~StageTimer() { if (on) std::fprintf(stderr, "ddpca set-up [%s] total %8.3f s\n", what, now() - t0); }
};

// Node-triple form of a transfer operator: runs of three consecutive rows that are shifted copies of each other
// (same length, same values, columns +1, +2) are stored once; everything else goes to the CSR remainder.  Worth it
// when most rows are in triples (no nodal rotations); returns with t.ok == false otherwise.
struct TripHost {
    bool ok = false;
    std::vector<int> trow, tptr, tcol, rest_row;
    std::vector<double> tw;
    CsrHost R;
};
static void build_trip_host(const CsrHost &A, TripHost &T)
{
    T = TripHost();
    T.tptr.assign(1, 0);
    T.R.rp.push_back(0);
    long in_trip = 0;
    for (int i = 0; i < A.rows;) {
        bool trip = false;
        if (i + 2 < A.rows) {
            const int len = A.rp[i + 1] - A.rp[i];
            trip = (A.rp[i + 2] - A.rp[i + 1] == len) && (A.rp[i + 3] - A.rp[i + 2] == len);
            for (int k = 0; trip && k < len; k++) {
                const int c = A.ci[A.rp[i] + k];
                const double w = A.v[A.rp[i] + k];
                trip = A.ci[A.rp[i + 1] + k] == c + 1 && A.ci[A.rp[i + 2] + k] == c + 2 && A.v[A.rp[i + 1] + k] == w && A.v[A.rp[i + 2] + k] == w;
            }
        }
        if (trip) {
            T.trow.push_back(i);
            for (int p = A.rp[i]; p < A.rp[i + 1]; p++) { T.tcol.push_back(A.ci[p]); T.tw.push_back(A.v[p]); }
            T.tptr.push_back((int)T.tcol.size());
            in_trip += 3;
            i += 3;
        } else {
            T.rest_row.push_back(i);
            for (int p = A.rp[i]; p < A.rp[i + 1]; p++) { T.R.ci.push_back(A.ci[p]); T.R.v.push_back(A.v[p]); }
            T.R.rp.push_back((int)T.R.ci.size());
            i += 1;
        }
    }
    T.R.rows = (int)T.rest_row.size(); T.R.cols = A.cols;
    T.ok = !(A.rows == 0 || in_trip < 0.8 * A.rows || std::getenv("DDPCA_NO_TRIP"));
}
static int upload_trip(const CsrHost &A, const TripHost &T, DevTrip &t)
{
    t.rows = A.rows; t.cols = A.cols;
    if (!T.ok) return 0;
    t.ntrip = (int)T.trow.size(); t.nrest = T.R.rows; t.npairs = (long)T.tcol.size(); t.nnz_rest = T.R.nnz();
    if (upload_vec(T.trow, &t.trow) || upload_vec(T.tptr, &t.tptr) || upload_vec(T.tcol, &t.tcol) || upload_vec(T.tw, &t.tw) ||
        upload_vec(T.rest_row, &t.rest_row) || upload_csr(T.R, t.rest)) return 1;
    t.ok = true;
    return 0;
}
// ---- kernel launch helpers (all on h->stream) ----------------------------------------------
static void launch_spmv(Engine *h, int kclass, int lvl, const DevCsr &A, const double *x, double *y, bool add,
                        const double *dotw, double *partial, const int *done, double alpha = 1.0)
{
    // algorithmic bytes, SURVEY.md §8(d): 12 nnz + 4 (rows+1) + 8 cols (x once) + 8 rows (y) [+8 rows for +=]
    double bytes = 12.0 * A.nnz + 4.0 * (A.rows + 1) + 8.0 * A.cols + 8.0 * A.rows * (add ? 2 : 1) + (dotw ? 8.0 * A.rows : 0.0);
    const double avg = A.avg_row();
    const int T = 256;
    auto grid_for = [&](int lanes) {
        long need = cdiv((long)A.rows * lanes, T);
        long cap = dotw ? kNumPart : (long)h->sms * 64;
        return (int)std::max<long>(1, std::min(need, cap));
    };
    if (dotw) {
        int g = grid_for(32);
        // stale slots of the partial buffer beyond g are never read: consumers use np = g
        KL(h, kclass, lvl, bytes, (k_spmv<32, false, true><<<g, T, 0, h->stream>>>(A.view(), alpha, x, y, dotw, partial, done)));
        return;
    }
#define SPMV_CASE(L)                                                                                                   \
    do {                                                                                                               \
        int g = grid_for(L);                                                                                           \
        if (add) KL(h, kclass, lvl, bytes, (k_spmv<L, true, false><<<g, T, 0, h->stream>>>(A.view(), alpha, x, y, nullptr, nullptr, done)));  \
        else KL(h, kclass, lvl, bytes, (k_spmv<L, false, false><<<g, T, 0, h->stream>>>(A.view(), alpha, x, y, nullptr, nullptr, done)));     \
    } while (0)
    if (avg > 40) SPMV_CASE(32);
    else if (avg > 20) SPMV_CASE(16);
    else if (avg > 10) SPMV_CASE(8);
    else SPMV_CASE(4);
#undef SPMV_CASE
}
// y (=|+=) T x with the node-triple form when available, else the CSR operator
static void launch_transfer(Engine *h, int kclass, int lvl, const DevCsr &A, const DevTrip &T, const double *x, double *y, bool add, const int *done)
{
    if (!T.ok) { launch_spmv(h, kclass, lvl, A, x, y, add, nullptr, nullptr, done); return; }
    // algorithmic bytes of the layout actually stored: 12 B per node pair + 8 B per triple (row, pointer) + remainder CSR + vectors
    const double bytes = 12.0 * T.npairs + 8.0 * T.ntrip + 12.0 * T.nnz_rest + 8.0 * T.nrest + 8.0 * T.cols + 8.0 * T.rows * (add ? 2 : 1);
    h->pre(kclass, lvl, bytes);
    const int TB = 256;
    if (T.ntrip) {
        if (T.avg_pairs() > 12.0) {
            const int g = cdiv((long)T.ntrip * 8, TB);
            if (add) k_trip_spmv<8, true><<<g, TB, 0, h->stream>>>(T.ntrip, T.trow, T.tptr, T.tcol, T.tw, x, y, done);
            else k_trip_spmv<8, false><<<g, TB, 0, h->stream>>>(T.ntrip, T.trow, T.tptr, T.tcol, T.tw, x, y, done);
        } else {
            const int g = cdiv((long)T.ntrip * 2, TB);
            if (add) k_trip_spmv<2, true><<<g, TB, 0, h->stream>>>(T.ntrip, T.trow, T.tptr, T.tcol, T.tw, x, y, done);
            else k_trip_spmv<2, false><<<g, TB, 0, h->stream>>>(T.ntrip, T.trow, T.tptr, T.tcol, T.tw, x, y, done);
        }
    }
    if (T.nrest) {
        if (h->capturing) h->captured_nodes++; else h->launches++;
        if (add) k_rowmap_spmv<true><<<cdiv(T.nrest, TB), TB, 0, h->stream>>>(T.nrest, T.rest_row, T.rest.view(), x, y, done);
        else k_rowmap_spmv<false><<<cdiv(T.nrest, TB), TB, 0, h->stream>>>(T.nrest, T.rest_row, T.rest.view(), x, y, done);
    }
    h->post();
}

// ---- v2 launches (kernels2.cuh) ------------------------------------------------------------------
template <int MODE>
static int v2_blocks_per_sm(size_t dyn)
{
    // function attributes are per device: a process may hold handles on several devices (and threads)
    static std::atomic<unsigned long long> attr_set{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(attr_set.load() & bit)) {
        cudaError_t e1 = cudaFuncSetAttribute(k_level_pass<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        // the chunk rings are the only consumers of the unified L1/shared array: take all of it,
        // otherwise the driver sizes the carve-out for ~5 CTAs and the half passes lose occupancy
        cudaError_t e2 = cudaFuncSetAttribute(k_level_pass<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e1 != cudaSuccess || e2 != cudaSuccess) {
            std::fprintf(stderr, "ddpca: cudaFuncSetAttribute(k_level_pass<%d>) failed: %s\n", MODE, cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
            cudaGetLastError();
            return 0;   // the caller reports the launch as impossible
        }
        attr_set.fetch_or(bit);
    }
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_level_pass<MODE>, v2_threads(MODE), dyn);
    return std::max(nb, 1);   // 0 is reserved for "attributes could not be set" above
}
// returns the grid used (number of partial sums for SPMV with a dot)
template <int MODE>
static int launch_v2(Engine *h, Level &L, int kclass, int l, double bytes, const double *b, double *x, double *p1, double *y,
                     const double *w, double *partial, const int *done)
{
    const bool staged = (MODE == V2_FWD_ZERO || MODE == V2_FWD_FULL || MODE == V2_BWD);
    const int nst = L.plan.nstages() - (MODE == V2_BWD ? L.fuse_bwd_last : 0);   // stages this launch walks
    if (nst <= 0) return 0;
    constexpr int tab = v2_table(MODE);
    size_t buf = L.buf[tab];
    size_t dyn = (size_t)kV2Bufs * buf;
    int per_sm = v2_blocks_per_sm<MODE>(dyn);
    if (per_sm <= 0) { if (h->launch_err.empty()) h->launch_err = "k_level_pass: shared-memory attributes could not be set"; return 0; }
    int cap = per_sm * h->sms;
    int grid = std::min(staged ? L.max_stage_chunks[tab] : L.nchunks[tab], cap);
    if (MODE == V2_SPMV && w) grid = std::min(grid, kNumPart);
    grid = std::max(grid, 1);
    if (std::getenv("DDPCA_VERBOSE")) {
        static int shown = 0;
        if (shown++ < 40) std::fprintf(stderr, "ddpca v2: level n=%d mode=%d buf=%zu B x%d, %d CTA/SM, grid=%d, chunks=%d (max/stage %d)\n", L.n, MODE, buf, kV2Bufs, per_sm, grid, L.nchunks[tab], L.max_stage_chunks[tab]);
    }
    Lvl2View A = L.view2(tab);
    unsigned *gbar = L.gbar;
    h->pre(kclass, l, bytes);
    if (staged && nst > 1) {
        void *args[] = {(void *)&A, (void *)&buf, (void *)&gbar, (void *)&b, (void *)&x, (void *)&p1, (void *)&y, (void *)&w, (void *)&partial, (void *)&done};
        cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_level_pass<MODE>, dim3(grid), dim3(v2_threads(MODE)), args, dyn, h->stream);
        if (e != cudaSuccess && h->launch_err.empty())
            h->launch_err = std::string("cooperative launch of k_level_pass failed (grid ") + std::to_string(grid) + " too large for co-residency?): " + cudaGetErrorString(e);
    } else {
        k_level_pass<MODE><<<grid, v2_threads(MODE), dyn, h->stream>>>(A, buf, gbar, b, x, p1, y, w, partial, done);
    }
    h->post();
    return grid;
}

// y = consStif[l] x on the group layout; returns the grid (= number of partial sums when dotw)
static int launch_level_spmv(Engine *h, Level &L, int l, const double *x, double *y, const double *dotw, double *partial, const int *done)
{
    if (L.v2) return launch_v2<V2_SPMV>(h, L, DDPCA_K_SPMV, l, L.bytes_full + (dotw ? 8.0 * L.n : 0.0), nullptr, const_cast<double *>(x), nullptr, y, dotw, partial, done);
    long need = cdiv((long)L.ng * GL, 256);
    int grid = (int)std::max<long>(1, std::min<long>(need, dotw ? kNumPart : (long)h->sms * 32));
    double bytes = L.bytes_full + (dotw ? 8.0 * L.n : 0.0);
    if (dotw) KL(h, DDPCA_K_SPMV, l, bytes, (k_spmv_group<true><<<grid, 256, 0, h->stream>>>(L.view(), x, y, dotw, partial, done)));
    else KL(h, DDPCA_K_SPMV, l, bytes, (k_spmv_group<false><<<grid, 256, 0, h->stream>>>(L.view(), x, y, nullptr, nullptr, done)));
    return grid;
}

static void sweep_fwd(Engine *h, Level &L, int l, const double *b, double *x, bool zero_x, const int *done)
{
    if (L.v2) {
        if (zero_x) launch_v2<V2_FWD_ZERO>(h, L, DDPCA_K_SWEEP_FWD0, l, L.bytes_lower, b, x, L.p1, nullptr, nullptr, nullptr, done);
        else launch_v2<V2_FWD_FULL>(h, L, DDPCA_K_SWEEP_FWD, l, L.bytes_lower + L.bytes_upper, b, x, L.p1, nullptr, nullptr, nullptr, done);
        return;
    }
    const int kc = zero_x ? DDPCA_K_SWEEP_FWD0 : DDPCA_K_SWEEP_FWD;
    for (const Segment &s : L.segs) {
        double bytes = s.bytes_lo + (zero_x ? 0.0 : s.bytes_up);
        if (!s.multi) {
            int ng = s.g1 - s.g0;
            int grid = cdiv((long)ng * GL, 256);
            if (zero_x) KL(h, kc, l, bytes, (k_sweep_fwd_stage<true><<<grid, 256, 0, h->stream>>>(L.view(), s.g0, s.g1, b, x, L.p1, done)));
            else KL(h, kc, l, bytes, (k_sweep_fwd_stage<false><<<grid, 256, 0, h->stream>>>(L.view(), s.g0, s.g1, b, x, L.p1, done)));
        } else {
            if (L.wide_rows) {
                if (zero_x) KL(h, kc, l, bytes, (k_sweep_fwd_multi<true, 32><<<1, 512, 0, h->stream>>>(L.view(), L.stage_group, s.s0, s.s1, b, x, L.p1, done)));
                else KL(h, kc, l, bytes, (k_sweep_fwd_multi<false, 32><<<1, 512, 0, h->stream>>>(L.view(), L.stage_group, s.s0, s.s1, b, x, L.p1, done)));
            } else {
                if (zero_x) KL(h, kc, l, bytes, (k_sweep_fwd_multi<true, GL><<<1, 512, 0, h->stream>>>(L.view(), L.stage_group, s.s0, s.s1, b, x, L.p1, done)));
                else KL(h, kc, l, bytes, (k_sweep_fwd_multi<false, GL><<<1, 512, 0, h->stream>>>(L.view(), L.stage_group, s.s0, s.s1, b, x, L.p1, done)));
            }
        }
    }
}
static void sweep_bwd(Engine *h, Level &L, int l, double *x, const int *done)
{
    if (L.v2) {
        launch_v2<V2_BWD>(h, L, DDPCA_K_SWEEP_BWD, l, L.bytes_upper - L.bytes_bwd_skip, nullptr, x, L.p1, nullptr, nullptr, nullptr, done);
        return;
    }
    for (int k = (int)L.segs.size() - 1; k >= 0; k--) {
        const Segment &s = L.segs[k];
        if (!s.multi) {
            int ng = s.g1 - s.g0;
            int grid = cdiv((long)ng * GL, 256);
            KL(h, DDPCA_K_SWEEP_BWD, l, s.bytes_up, (k_sweep_bwd_stage<<<grid, 256, 0, h->stream>>>(L.view(), s.g0, s.g1, L.p1, x, done)));
        } else {
            if (L.wide_rows) KL(h, DDPCA_K_SWEEP_BWD, l, s.bytes_up, (k_sweep_bwd_multi<32><<<1, 512, 0, h->stream>>>(L.view(), L.stage_group, s.s0, s.s1, L.p1, x, done)));
            else KL(h, DDPCA_K_SWEEP_BWD, l, s.bytes_up, (k_sweep_bwd_multi<GL><<<1, 512, 0, h->stream>>>(L.view(), L.stage_group, s.s0, s.s1, L.p1, x, done)));
        }
    }
}

// MGPIS::MULT_VCYC (MGPIS.h:55-128) on device vectors in device numbering
static void vcycle_dev(ddpca_mg *h, int l, const double *b, double *x, bool zero_x, const int *done)
{
    if (l == 0) {
        KL(h, DDPCA_K_COARSE, 0, h->coarse_bytes, (k_dense_gemv_batch<<<cdiv((long)h->n0 * 32, 256), 256, 0, h->stream>>>(h->nsub, h->sub0_off_d, h->binv_ptr_d, b, x, done)));
        return;
    }
    Level &L = h->lev[l];
    Level &C = h->lev[l - 1];
    sweep_fwd(h, L, l, b, x, zero_x, done);  // :65-72
    sweep_bwd(h, L, l, x, done);             // :73-76
    if (L.v2) launch_v2<V2_RESID>(h, L, DDPCA_K_RESID, l, L.bytes_lower + 8.0 * L.n, b, x, L.p1, L.r, nullptr, nullptr, done);
    else KL(h, DDPCA_K_RESID, l, L.bytes_lower + 8.0 * L.n, (k_resid_lower<<<cdiv((long)L.ng * GL, 256), 256, 0, h->stream>>>(L.view(), b, L.p1, x, L.r, done)));
    launch_transfer(h, DDPCA_K_RESTRICT, l, L.R, L.Rt, L.r, C.b, false, done);  // :96
    vcycle_dev(h, l - 1, C.b, C.x, true, done);                                          // :93-99
    launch_transfer(h, DDPCA_K_PROLONG, l, L.P, L.Pt, C.x, x, true, done);               // :100
    sweep_fwd(h, L, l, b, x, false, done);  // :102-109
    sweep_bwd(h, L, l, x, done);            // :110-113
}

static void precondition(ddpca_mg *h, int prec, const double *r, double *z, const int *done)
{
    int Lf = h->nlev - 1;
    Level &L = h->lev[Lf];
    if (prec == 0) {
        KL(h, DDPCA_K_VECTOR, Lf, 24.0 * L.n, (k_jacobi<<<cdiv(L.n, 256), 256, 0, h->stream>>>(L.n, L.dinv, r, z, done)));
    } else {
        vcycle_dev(h, Lf, r, z, true, done);
    }
}

static int vec_grid(const Engine *h, int n) { return std::max(1, std::min(cdiv(n, 256), std::min(kNumPart, h->sms * 8))); }

// body of one CG iteration, MGPIS.h:199-219, for every sub of the batch (segmented reductions, kernels.cuh)
static void enqueue_iteration(ddpca_mg *h, int prec, cudaGraphConditionalHandle cond = 0, int use_cond = 0)
{
    int Lf = h->nlev - 1;
    Level &L = h->lev[Lf];
    const int *done = h->done_flag();
    const int n = L.n, ns = h->nseg;
    launch_level_spmv(h, L, Lf, h->cg_p, h->cg_q, nullptr, nullptr, done);  // :200
    KL(h, DDPCA_K_VECTOR, Lf, 16.0 * n, (k_seg_dot<<<ns, 256, 0, h->stream>>>(h->seg_d, h->cg_p, h->cg_q, h->partial[0], done)));  // p.q of :201
    KL(h, DDPCA_K_VECTOR, Lf, 48.0 * n, (k_seg_update_xr<<<ns, 256, 0, h->stream>>>(h->seg_d, h->sub_chunk_d, h->st, h->partial[0], h->cg_p, h->cg_q, h->cg_x, h->cg_r, h->partial[1], done)));  // :201-203
    precondition(h, prec, h->cg_r, h->cg_z, done);  // :204-210
    KL(h, DDPCA_K_VECTOR, Lf, 16.0 * n, (k_seg_dot<<<ns, 256, 0, h->stream>>>(h->seg_d, h->cg_r, h->cg_z, h->partial[2], done)));  // :212
    KL(h, DDPCA_K_VECTOR, Lf, 0.0, (k_seg_beta_next<<<1, 1024, 0, h->stream>>>(h->nsub, h->st, h->sub_chunk_d, h->partial[1], h->partial[2], h->fl, cond, use_cond)));  // :211-213,219,198
    KL(h, DDPCA_K_VECTOR, Lf, 24.0 * n, (k_seg_update_p<<<ns, 256, 0, h->stream>>>(h->seg_d, h->st, h->cg_z, h->cg_p, done)));   // :214
}

// set-up of a solve after r = b, x = 0 (MGPIS.h:174-197): tolerance, first preconditioner application, delta_new
static void enqueue_setup(ddpca_mg *h, int prec, cudaGraphConditionalHandle cond = 0, int use_cond = 0)
{
    int Lf = h->nlev - 1;
    int n = h->lev[Lf].n;
    const int *done = h->done_flag();
    const int ns = h->nseg;
    KL(h, DDPCA_K_VECTOR, Lf, 8.0 * n, (k_seg_dot<<<ns, 256, 0, h->stream>>>(h->seg_d, h->cg_r, h->cg_r, h->partial[0], nullptr)));
    KL(h, DDPCA_K_VECTOR, Lf, 0.0, (k_seg_init<<<1, 1024, 0, h->stream>>>(h->nsub, h->st, h->sub_chunk_d, h->partial[0], h->fl)));   // :174-175
    precondition(h, prec, h->cg_r, h->cg_p, done);                                                                                    // :191-196
    KL(h, DDPCA_K_VECTOR, Lf, 16.0 * n, (k_seg_dot<<<ns, 256, 0, h->stream>>>(h->seg_d, h->cg_r, h->cg_p, h->partial[2], done)));
    KL(h, DDPCA_K_VECTOR, Lf, 0.0, (k_seg_delta0<<<1, 1024, 0, h->stream>>>(h->nsub, h->st, h->sub_chunk_d, h->partial[2], h->fl, cond, use_cond)));  // :197
}

// One graph per solve: [set-up] -> WHILE(cond){ CG iteration }.  The loop condition of MGPIS.h:198
// is evaluated on the device (k_s_delta0 / k_s_beta_next call cudaGraphSetConditional): no host
// polling, no iterations issued past convergence.  Returns 0 and sets while_state[prec] = -1 when
// the driver refuses (then the host-polled per-iteration graph is used).
static int build_solve_graph(ddpca_mg *h, int prec)
{
    if (h->while_state[prec] != 0) return 0;
    h->while_state[prec] = -1;
    if (std::getenv("DDPCA_NO_WHILE_GRAPH")) return 0;
    cudaGraph_t g = nullptr;
    cudaStream_t saved = h->stream;
    h->stream = h->own_stream;
    bool ok = false;
    do {
        if (cudaGraphCreate(&g, 0) != cudaSuccess) break;
        cudaGraphConditionalHandle cond;
        if (cudaGraphConditionalHandleCreate(&cond, g, 0, cudaGraphCondAssignDefault) != cudaSuccess) break;
        // 1. set-up nodes
        h->capturing = true; h->captured_nodes = 0;
        if (cudaStreamBeginCaptureToGraph(h->stream, g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { h->capturing = false; break; }
        enqueue_setup(h, prec, cond, 1);
        cudaStreamCaptureStatus cs; const cudaGraphNode_t *deps = nullptr; size_t ndeps = 0;
        cudaError_t e1 = cudaStreamGetCaptureInfo(h->stream, &cs, nullptr, nullptr, &deps, &ndeps);
        std::vector<cudaGraphNode_t> depv(deps, deps + (e1 == cudaSuccess ? ndeps : 0));
        cudaGraph_t gout = nullptr;
        cudaError_t e2 = cudaStreamEndCapture(h->stream, &gout);
        h->capturing = false;
        h->solve_init_nodes[prec] = h->captured_nodes;
        if (e1 != cudaSuccess || e2 != cudaSuccess) break;
        // 2. WHILE node after the set-up
        cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
        np.conditional.handle = cond;
        np.conditional.type = cudaGraphCondTypeWhile;
        np.conditional.size = 1;
        cudaGraphNode_t wnode;
        if (cudaGraphAddNode(&wnode, g, depv.data(), depv.size(), &np) != cudaSuccess) break;
        cudaGraph_t body = np.conditional.phGraph_out[0];
        // 3. iteration body
        h->capturing = true; h->captured_nodes = 0;
        if (cudaStreamBeginCaptureToGraph(h->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { h->capturing = false; break; }
        enqueue_iteration(h, prec, cond, 1);
        cudaError_t e3 = cudaStreamEndCapture(h->stream, &gout);
        h->capturing = false;
        h->solve_iter_nodes[prec] = h->captured_nodes;
        if (e3 != cudaSuccess) break;
        if (cudaGraphInstantiate(&h->solve_graph[prec], g, 0) != cudaSuccess) break;
        ok = true;
    } while (0);
    h->capturing = false;
    h->stream = saved;
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();   // clear any sticky-free error of the attempt
    if (ok) h->while_state[prec] = 1;
    else if (std::getenv("DDPCA_VERBOSE")) std::fprintf(stderr, "ddpca: WHILE-graph unavailable, using the host-polled CG loop\n");
    return 0;
}

static int build_iter_graph(ddpca_mg *h, int prec)
{
    if (h->iter_graph[prec]) return 0;
    cudaGraph_t g = nullptr;
    // capture on the handle's own stream (an external stream may be in use by others)
    cudaStream_t saved = h->stream;
    h->stream = h->own_stream;
    h->capturing = true;
    h->captured_nodes = 0;
    cudaError_t e = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
        enqueue_iteration(h, prec);
        e = cudaStreamEndCapture(h->stream, &g);
    }
    h->capturing = false;
    h->stream = saved;
    if (e != cudaSuccess) return fail(std::string("graph capture: ") + cudaGetErrorString(e));
    h->iter_graph_nodes[prec] = h->captured_nodes;
    CU(cudaGraphInstantiate(&h->iter_graph[prec], g, 0));
    CU(cudaGraphDestroy(g));
    return 0;
}

// wait for a solve enqueued by pcg_device and read its scalars back (iters / resid / tol_abs: sub 0, or the
// largest iteration count / residual of a batch; per-sub values through ddpca_mg_batch_result)
static int pcg_finish(ddpca_mg *h, long *iters, double *resid, double *tol_abs)
{
    CU(cudaStreamSynchronize(h->stream));
    if (!h->launch_err.empty()) { std::string m = h->launch_err; h->launch_err.clear(); cudaGetLastError(); return fail(m); }
    if (h->profile) h->prof_collect();
    if (h->pending_while) h->launches += h->solve_init_nodes[h->pending_prec] + (long)h->fl_host[0].it_max * h->solve_iter_nodes[h->pending_prec];
    long it = 0;
    double rr = 0.0;
    for (int s = 0; s < h->nsub; s++) { it = std::max(it, (long)h->st_host[s].it); rr = std::max(rr, h->st_host[s].rr); }
    if (iters) *iters = it;
    if (resid) *resid = std::sqrt(rr);
    if (tol_abs) *tol_abs = h->st_host[0].tol;
    CU(cudaGetLastError());
    return 0;
}

// CG_SOLV on device vectors: b_ref/x_ref in REFERENCE numbering (the subs' vectors one after the other),
// resident on the device.  maxit <= 0: the rows of each sub (MGPIS.h:178).
static int pcg_device(ddpca_mg *h, int prec, const double *b_ref, double *x_ref, double rel_tol, long maxit,
                      long *iters, double *resid, double *tol_abs, bool no_wait = false)
{
    if (prec != 0 && prec != 1) return fail("prec must be 0 (Jacobi) or 1 (V-cycle)");
    if (prec == 1 && !h->Binv) return fail("this hierarchy was built without a level-0 direct solver: Jacobi preconditioning only");
    int Lf = h->nlev - 1;
    Level &L = h->lev[Lf];
    int n = L.n;
    if (prec == 0 && !L.dinv) {
        CU(cudaMalloc(&L.dinv, sizeof(double) * n));
        if (L.v2) k_extract_diag_inv2<<<cdiv(L.ng, 256), 256, 0, h->stream>>>(L.view2(), L.dinv);
        else k_extract_diag_inv<<<cdiv(L.ng, 256), 256, 0, h->stream>>>(L.view(), L.dinv);
    }
    if (!h->profile) {
        build_solve_graph(h, prec);
        if (h->while_state[prec] != 1 && build_iter_graph(h, prec)) return 1;
    }
    const bool use_while = !h->profile && h->while_state[prec] == 1;
    // r = b (device numbering), x = 0                                  MGPIS.h:173,189
    KL(h, DDPCA_K_VECTOR, Lf, 20.0 * n, (k_gather<<<cdiv(n, 256), 256, 0, h->stream>>>(n, L.perm, b_ref, h->cg_r)));
    CU(cudaMemsetAsync(h->cg_x, 0, sizeof(double) * n, h->stream));
    CU(cudaMemsetAsync(h->cg_p, 0, sizeof(double) * n, h->stream));
    KL(h, DDPCA_K_VECTOR, Lf, 0.0, (k_seg_params<<<cdiv(h->nsub, 128), 128, 0, h->stream>>>(h->nsub, h->st, h->sub_n_d, rel_tol, (long long)maxit)));
    if (use_while) {
        CU(cudaGraphLaunch(h->solve_graph[prec], h->stream));
    } else {
        enqueue_setup(h, prec);
        // main loop: enqueue iterations ahead, poll the batch's done flag with a lag of `depth` iterations.  Under the
        // per-launch profiler the lag is zero, so that no early-exiting launch is timed (and booked with full bytes).
        const int depth = h->profile ? 0 : kDepth;
        cudaEvent_t evs[kDepth + 1];
        for (int k = 0; k <= kDepth; k++) CU(cudaEventCreateWithFlags(&evs[k], cudaEventDisableTiming));
        long issued = 0;
        bool finished = false;
        // state after setup (covers zero RHS: done already set, MGPIS.h:198 never entered)
        CU(cudaMemcpyAsync(&h->fl_host[0], h->fl, sizeof(BatchFlags), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaEventRecord(evs[0], h->stream));
        CU(cudaEventSynchronize(evs[0]));
        finished = h->fl_host[0].done_all != 0;
        while (!finished) {
            int slot = (int)(issued % (kDepth + 1));
            if (h->profile) {
                enqueue_iteration(h, prec);
            } else {
                CU(cudaGraphLaunch(h->iter_graph[prec], h->stream));
                h->launches += h->iter_graph_nodes[prec];
            }
            CU(cudaMemcpyAsync(&h->fl_host[slot], h->fl, sizeof(BatchFlags), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaEventRecord(evs[slot], h->stream));
            issued++;
            if (issued >= depth) {
                int old = (int)((issued - std::max(depth, 1)) % (kDepth + 1));
                CU(cudaEventSynchronize(evs[old]));
                if (h->fl_host[old].done_all) finished = true;
            }
            if (h->profile && (issued % 8) == 0) h->prof_collect();
        }
        for (int k = 0; k <= kDepth; k++) cudaEventDestroy(evs[k]);
    }
    // x (device numbering) -> reference numbering
    KL(h, DDPCA_K_VECTOR, Lf, 20.0 * n, (k_scatter<<<cdiv(n, 256), 256, 0, h->stream>>>(n, L.perm, h->cg_x, x_ref)));
    CU(cudaMemcpyAsync(h->st_host, h->st, sizeof(PcgState) * h->nsub, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(&h->fl_host[0], h->fl, sizeof(BatchFlags), cudaMemcpyDeviceToHost, h->stream));
    h->pending_while = use_while;
    h->pending_prec = prec;
    if (no_wait) return 0;   // the caller overlaps several solves and calls pcg_finish() later
    return pcg_finish(h, iters, resid, tol_abs);
}

// ------------------------------------------------------------------------------------------
// Host construction of the group layout of one (already permuted) level + the sweep schedule.
struct GroupLayoutHost {
    std::vector<GroupMeta> meta;
    std::vector<int> ci;
    std::vector<double> v;
    long pat_total = 0, val_total = 0;   // sizes of ci / v (the arrays are only filled when the level uses this layout)
};

static void fill_group_arrays(const CsrHost &Ap, GroupLayoutHost &out);
// descriptors of the group layout; the pattern / value arrays are filled by fill_group_arrays when the level keeps this
// layout (levels that go on to the split v2 layout only need the descriptors: skipping the copy saves a pass over the
// whole operator at set-up)
static bool build_group_layout(const CsrHost &Ap, const LevelPlan &pl, GroupLayoutHost &out, std::string &err)
{
    int ng = pl.ngroups();
    out.meta.resize(ng);
    long pat = 0, val = 0;
    for (int g = 0; g < ng; g++) {
        int r0 = pl.group_start[g], gs = pl.group_start[g + 1] - r0;
        int len = Ap.rp[r0 + 1] - Ap.rp[r0];
        int lenp = len + (len & 1);
        const int *c0 = Ap.ci.data() + Ap.rp[r0];
        int kd = (int)(std::lower_bound(c0, c0 + len, r0) - c0);
        if (kd + gs > len || c0[kd] != r0 || c0[kd + gs - 1] != r0 + gs - 1) { err = "group " + std::to_string(g) + ": in-group block incomplete"; return false; }
        GroupMeta m;
        m.row0 = r0; m.gs = gs; m.cptr = (int)pat; m.len = lenp; m.voff = val; m.kd = kd; m.pad = len;
        out.meta[g] = m;
        pat += lenp;
        val += (long)gs * lenp;
        if (pat > 0x7fffffffL) { err = "pattern index overflow"; return false; }
    }
    out.pat_total = pat;
    out.val_total = val;
    return true;
}
static void fill_group_arrays(const CsrHost &Ap, GroupLayoutHost &out)
{
    const int ng = (int)out.meta.size();
    out.ci.resize(out.pat_total);
    out.v.assign(out.val_total, 0.0);
#pragma omp parallel for schedule(static)
    for (int g = 0; g < ng; g++) {
        const GroupMeta &m = out.meta[g];
        int len = m.pad;
        const int *c0 = Ap.ci.data() + Ap.rp[m.row0];
        for (int k = 0; k < len; k++) out.ci[m.cptr + k] = c0[k];
        if (m.len > len) out.ci[m.cptr + len] = Ap.rows;      // padding: value 0 times the always-zero slot x[n] (alloc_vec)
        for (int r = 0; r < m.gs; r++) {
            const double *vr = Ap.v.data() + Ap.rp[m.row0 + r];
            double *dst = out.v.data() + m.voff + (long)r * m.len;
            for (int k = 0; k < len; k++) dst[k] = vr[k];
        }
    }
}

// ---- v2 layout (kernels2.cuh): split lower / upper arrays in stage order + chunk table -----------
struct Layout2Host {
    std::vector<GroupMeta2> meta;
    RawVec<int> CL, CU;       // written in full (pads included) by the parallel fill of build_layout2
    RawVec<double> VL, VU;
    std::vector<double> BD;
    // per chunk table (V2_TAB_*)
    std::vector<ChunkDesc> chunks[3];
    std::vector<int> stage_chunk[3];
    size_t buf[3] = {0, 0, 0};
    int max_stage_chunks[3] = {0, 0, 0};
};
static inline int round4(int v) { return (v + 3) & ~3; }

// Chunk tables of the v2 layout: consecutive groups, never across a stage boundary.  Whole-row passes
// take kChunkGroups groups per chunk; the half-pass tables take groups while the streamed bytes stay
// within 1/kHalfBudgetDiv of the largest whole-row chunk (at most kHalfGroups groups).
// (Re-cutting a colour into a multiple of the grid of equal chunks, so that no CTA waits for a last,
// nearly empty round, was measured and is slower: more, smaller chunks cost more than the tails.)
static void build_chunks2(const LevelPlan &pl, Layout2Host &out, const std::vector<int> *group_sub)
{
    const int ns = pl.nstages();
    auto make_desc = [&](int g0, int g1) {
        ChunkDesc d{};
        d.g0 = g0; d.ng = g1 - g0;
        d.sub = group_sub ? (*group_sub)[g0] : 0;
        const GroupMeta2 &a = out.meta[g0], &z = out.meta[g1 - 1];
        d.cl0 = a.cl; d.ncl = z.cl + z.nl - a.cl;
        d.cu0 = a.cu; d.ncu = z.cu + z.nu - a.cu;
        d.vl0 = a.vl; d.nvl = z.vl + z.gs * z.nl / 2 - a.vl;
        d.vu0 = a.vu; d.nvu = z.vu + z.gs * z.nu / 2 - a.vu;
        return d;
    };
    size_t budget = 0;   // streamed bytes (patterns + values) of the largest whole-row chunk
    for (int tab = 0; tab < 3; tab++) {
        const int mode = tab == V2_TAB_FULL ? V2_FWD_FULL : (tab == V2_TAB_LO ? V2_FWD_ZERO : V2_BWD);
        const int cap = v2_groups(mode);
        const size_t fixed = v2_chunk_bytes(mode, 0, 0, 0, 0);
        auto data_bytes = [&](const ChunkDesc &d) { return v2_chunk_bytes(mode, d.ncl, d.nvl, d.ncu, d.nvu) - fixed; };
        std::vector<ChunkDesc> &ch = out.chunks[tab];
        std::vector<int> &sc = out.stage_chunk[tab];
        ch.clear();
        out.buf[tab] = 0; out.max_stage_chunks[tab] = 0;
        sc.assign(ns + 1, 0);
        const size_t lim = tab == V2_TAB_FULL ? (size_t)-1 : budget / kHalfBudgetDiv;
        std::vector<ChunkDesc> st;
        for (int s = 0; s < ns; s++) {
            sc[s] = (int)ch.size();
            const int gbeg = pl.stage_group[s], gend = pl.stage_group[s + 1];
            st.clear();
            for (int g0 = gbeg; g0 < gend;) {
                int g1 = g0 + 1;
                while (g1 < gend && g1 - g0 < cap && (!group_sub || (*group_sub)[g1] == (*group_sub)[g0]) && data_bytes(make_desc(g0, g1 + 1)) <= lim) g1++;
                st.push_back(make_desc(g0, g1));
                g0 = g1;
            }
            for (const ChunkDesc &d : st) {
                ch.push_back(d);
                const size_t b = fixed + data_bytes(d);
                out.buf[tab] = std::max(out.buf[tab], b);
                if (tab == V2_TAB_FULL) budget = std::max(budget, b - fixed);
            }
            out.max_stage_chunks[tab] = std::max(out.max_stage_chunks[tab], (int)ch.size() - sc[s]);
        }
        sc[ns] = (int)ch.size();
        out.buf[tab] = (out.buf[tab] + 127) & ~(size_t)127;
    }
}

static bool build_layout2(const CsrHost &Ap, const LevelPlan &pl, Layout2Host &out, std::string &err, const std::vector<int> *group_sub)
{
    const int ng = pl.ngroups();
    out.meta.resize(ng);
    long cl = 0, cu = 0, vl = 0, vu = 0;   // ints / 16-byte units
    for (int g = 0; g < ng; g++) {
        int r0 = pl.group_start[g], gs = pl.group_start[g + 1] - r0;
        int len = Ap.rp[r0 + 1] - Ap.rp[r0];
        const int *c0 = Ap.ci.data() + Ap.rp[r0];
        int kd = (int)(std::lower_bound(c0, c0 + len, r0) - c0);
        if (kd + gs > len || c0[kd] != r0 || c0[kd + gs - 1] != r0 + gs - 1) { err = "group " + std::to_string(g) + ": in-group block incomplete"; return false; }
        GroupMeta2 m;
        m.row0 = r0; m.gs = gs;
        m.nl = round4(kd); m.nu = round4(len - kd - gs);
        m.cl = (int)cl; m.cu = (int)cu; m.vl = (int)vl; m.vu = (int)vu;
        out.meta[g] = m;
        cl += m.nl; cu += m.nu;
        vl += (long)gs * m.nl / 2; vu += (long)gs * m.nu / 2;
        if (cl > 0x7ffffff0L || cu > 0x7ffffff0L || vl > 0x7ffffff0L || vu > 0x7ffffff0L) { err = "level too large for 32-bit chunk offsets"; return false; }
    }
    out.CL.resize(cl); out.CU.resize(cu);
    out.VL.resize(vl * 2); out.VU.resize(vu * 2);
    out.BD.assign((size_t)ng * kBlkStride, 0.0);
#pragma omp parallel for schedule(static)
    for (int g = 0; g < ng; g++) {
        const GroupMeta2 &m = out.meta[g];
        const int r0 = m.row0, gs = m.gs;
        const int len = Ap.rp[r0 + 1] - Ap.rp[r0];
        const int *c0 = Ap.ci.data() + Ap.rp[r0];
        const int kd = (int)(std::lower_bound(c0, c0 + len, r0) - c0);
        const int nlr = kd, nur = len - kd - gs;
        // pads: value 0 times the always-zero slot x[n] of every gathered vector (alloc_vec) -- never a row
        // of the level, whose entry may be stale or, after a failed solve, not finite (0 * NaN = NaN)
        for (int k = 0; k < m.nl; k++) out.CL[m.cl + k] = k < nlr ? c0[k] : Ap.rows;
        for (int k = 0; k < m.nu; k++) out.CU[m.cu + k] = k < nur ? c0[kd + gs + k] : Ap.rows;
        for (int r = 0; r < gs; r++) {
            const double *vr = Ap.v.data() + Ap.rp[r0 + r];
            double *dl = out.VL.data() + (size_t)m.vl * 2 + (size_t)r * m.nl;
            double *du = out.VU.data() + (size_t)m.vu * 2 + (size_t)r * m.nu;
            for (int k = 0; k < nlr; k++) dl[k] = vr[k];
            for (int k = nlr; k < m.nl; k++) dl[k] = 0.0;
            for (int k = 0; k < nur; k++) du[k] = vr[kd + gs + k];
            for (int k = nur; k < m.nu; k++) du[k] = 0.0;
            for (int c = 0; c < gs; c++) out.BD[(size_t)g * kBlkStride + r * 3 + c] = vr[kd + c];
            out.BD[(size_t)g * kBlkStride + 9 + r] = 1.0 / vr[kd + r];
        }
    }
    build_chunks2(pl, out, group_sub);
    return true;
}

static int build_segments(Level &L, const GroupLayoutHost &G)
{
    const LevelPlan &pl = L.plan;
    const int kBigStage = 48;  // groups; smaller stages are merged into single-CTA runs
    int ns = pl.nstages();
    // algorithmic bytes of relaxing groups [ga,gb) over the lower / upper half:
    // 8 B per value + 4 B per pattern entry of the half + per row: 8 (rhs) + 16 (x read+write) + 8 (p1)
    // + 32 B group descriptor + the gs*gs in-group block
    auto seg_bytes = [&](int ga, int gb, double &lo, double &up) {
        lo = up = 0;
        for (int g = ga; g < gb; g++) {
            const GroupMeta &m = G.meta[g];
            double common = 32.0 + 8.0 * m.gs * m.gs + 32.0 * m.gs;
            lo += (8.0 * m.gs + 4.0) * m.kd + common;
            up += (8.0 * m.gs + 4.0) * (m.pad - m.kd - m.gs) + common;
        }
    };
    int s = 0;
    while (s < ns) {
        int ngs = pl.stage_group[s + 1] - pl.stage_group[s];
        Segment seg{};
        if (ngs >= kBigStage) {
            seg.multi = 0; seg.s0 = s; seg.s1 = s + 1;
        } else {
            int e = s + 1;
            while (e < ns && pl.stage_group[e + 1] - pl.stage_group[e] < kBigStage) e++;
            seg.multi = 1; seg.s0 = s; seg.s1 = e;
        }
        seg.g0 = pl.stage_group[seg.s0];
        seg.g1 = pl.stage_group[seg.s1];
        seg_bytes(seg.g0, seg.g1, seg.bytes_lo, seg.bytes_up);
        L.segs.push_back(seg);
        s = seg.s1;
    }
    return 0;
}

// Set-up of one operator level in two halves: everything the host computes (plan, permuted operator, layouts, chunk
// tables, byte counts -- no CUDA call, also run by ddpca_mg_setup_dryrun) and the uploads (device must be current).
struct SetupClock {   // seconds spent per stage of the hierarchy set-up (reported under DDPCA_VERBOSE and by the dry run)
    double concat = 0, plan = 0, permute = 0, layout = 0, upload = 0, tr_permute = 0, tr_transpose = 0, tr_trip = 0, tr_upload = 0;
};
static thread_local SetupClock g_clock;
struct LevelHost {
    CsrHost Ap;   // operator in device numbering
    GroupLayoutHost G;
    Layout2Host H2;
    bool group_layout = false;
};
static int setup_level_host(Level &L, const CsrBlocks &B, int mode, bool group_layout, const LevelPlan *given_plan, LevelHost &D,
                            const std::vector<int> *whole_sub_off = nullptr)
{
    // whole_sub_off: B is ONE block holding an already concatenated batch whose subs start at these rows (cross-check path)
    const int n = B.rows();
    const std::vector<int> *sub_off = whole_sub_off ? whole_sub_off : (B.nsub > 1 ? &B.row_off : nullptr);
    L.n = n;
    std::string err;
    double t0 = StageTimer::now();
    if (given_plan) L.plan = *given_plan;
    else if (whole_sub_off) { if (!build_level_plan_blocks(n, B.rp[0], B.ci[0], mode, *whole_sub_off, L.plan, err)) return fail(err); }
    else if (!build_level_plan_subs(B, mode, L.plan, err)) return fail(err);
    double t1 = StageTimer::now();
    g_clock.plan += t1 - t0;
    CsrHost &Ap = D.Ap;
    permute_csr_blocks(B, L.plan.perm, L.plan.iperm, Ap);
    L.nnz = Ap.nnz();
    double t2 = StageTimer::now();
    g_clock.permute += t2 - t1;
    D.group_layout = group_layout;
    if (group_layout) {
        GroupLayoutHost &G = D.G;
        if (!build_group_layout(Ap, L.plan, G, err)) return fail(err);
        if (mode >= 0) build_segments(L, G);
        // v2 (kernels2.cuh) when every stage is large enough to be worth a grid-wide pass
        bool want_v2 = mode >= 0 && !L.segs.empty() && !std::getenv("DDPCA_NO_V2");
        for (const Segment &sg : L.segs) if (sg.multi) want_v2 = false;
        Layout2Host &H2 = D.H2;
        if (want_v2) {
            // subdomain of every row group (batched hierarchies: chunks are cut at subdomain boundaries)
            std::vector<int> group_sub;
            if (sub_off && sub_off->size() > 2) {
                const int ngp = L.plan.ngroups();
                group_sub.resize(ngp);
                for (int g = 0; g < ngp; g++) {
                    const int old = L.plan.perm[L.plan.group_start[g]];
                    group_sub[g] = (int)(std::upper_bound(sub_off->begin(), sub_off->end(), old) - sub_off->begin()) - 1;
                }
            }
            if (!build_layout2(Ap, L.plan, H2, err, group_sub.empty() ? nullptr : &group_sub)) return fail(err);
            for (int tab = 0; tab < 3; tab++)
                if ((size_t)kV2Bufs * H2.buf[tab] > 200 * 1024) want_v2 = false;   // the chunk ring must fit in shared memory
        }
        L.v2 = want_v2;
        L.ng = (int)G.meta.size();
        L.pat_entries = G.pat_total;
        L.val_entries = G.val_total;
        for (const GroupMeta &m : G.meta) {
            double common = 32.0 + 8.0 * m.gs * m.gs + 32.0 * m.gs;
            L.bytes_lower += (8.0 * m.gs + 4.0) * m.kd + common;
            L.bytes_upper += (8.0 * m.gs + 4.0) * (m.pad - m.kd - m.gs) + common;
            L.bytes_full += (8.0 * m.gs + 4.0) * m.pad + 32.0 + 16.0 * m.gs;   // + x once, y once
        }
        if (L.v2) {
            // fwd -> bwd junction: the last colour has nothing above its groups (it is ordered last)
            const int ns = L.plan.nstages();
            bool none_above = ns >= 1 && !std::getenv("DDPCA_NO_FUSE_BWD");
            for (int g = L.plan.stage_group[ns - 1]; none_above && g < L.plan.stage_group[ns]; g++) none_above = (H2.meta[g].nu == 0);
            L.fuse_bwd_last = none_above ? 1 : 0;
            if (none_above)
                for (int g = L.plan.stage_group[ns - 1]; g < L.plan.stage_group[ns]; g++) {
                    const int gs = H2.meta[g].gs;
                    L.bytes_bwd_skip += 32.0 + 8.0 * gs * gs + 32.0 * gs;
                }
            for (int tab = 0; tab < 3; tab++) {
                L.nchunks[tab] = (int)H2.chunks[tab].size();
                L.max_stage_chunks[tab] = H2.max_stage_chunks[tab];
                L.buf[tab] = H2.buf[tab];
            }
        } else {
            fill_group_arrays(Ap, G);
        }
    }
    g_clock.layout += StageTimer::now() - t2;
    return 0;
}
static int setup_level_upload(Level &L, const LevelHost &D, bool keep_csr)
{
    const double t0 = StageTimer::now();
    const int n = L.n;
    if (keep_csr) { if (upload_csr(D.Ap, L.A)) return 1; }
    if (D.group_layout) {
        if (L.v2) {
            const Layout2Host &H2 = D.H2;
            for (int tab = 0; tab < 3; tab++)
                if (upload_vec(H2.chunks[tab], &L.chunks[tab]) || upload_vec(H2.stage_chunk[tab], &L.stage_chunk[tab])) return 1;
            CU(cudaMalloc(&L.empty_desc, sizeof(ChunkDesc)));
            CU(cudaMemset(L.empty_desc, 0, sizeof(ChunkDesc)));
            CU(cudaMalloc(&L.gbar, kGbarWords * sizeof(unsigned)));
            CU(cudaMemset(L.gbar, 0, kGbarWords * sizeof(unsigned)));
            if (upload_vec(H2.meta, &L.meta2) || upload_vec(H2.CL, &L.CL) || upload_vec(H2.CU, &L.CU) || upload_vec(H2.VL, &L.VL) ||
                upload_vec(H2.VU, &L.VU) || upload_vec(H2.BD, &L.BD)) return 1;
        } else {
            if (upload_vec(D.G.meta, &L.meta) || upload_vec(D.G.ci, &L.gci) || upload_vec(D.G.v, &L.gv)) return 1;
        }
    }
    if (upload_vec(L.plan.stage_group, &L.stage_group) || upload_vec(L.plan.perm, &L.perm)) return 1;
    if (alloc_vec(&L.x, n) || alloc_vec(&L.b, n) || alloc_vec(&L.p1, n) || alloc_vec(&L.r, n)) return 1;
    g_clock.upload += StageTimer::now() - t0;
    return 0;
}
static int setup_level(Level &L, int n, const int *rp, const int *ci, const double *v, int mode, bool keep_csr, bool group_layout,
                       const std::vector<int> *sub_off = nullptr, const LevelPlan *given_plan = nullptr)
{
    LevelHost D;
    (void)sub_off;
    if (setup_level_host(L, CsrBlocks::single(n, n, rp, ci, v), mode, group_layout, given_plan, D)) return 1;
    return setup_level_upload(L, D, keep_csr);
}
static void free_trip(DevTrip &t)
{
    cudaFree(t.trow); cudaFree(t.tptr); cudaFree(t.tcol); cudaFree(t.tw); cudaFree(t.rest_row);
    free_csr(t.rest);
    t = DevTrip();
}
static void free_level(Level &L)
{
    free_csr(L.A); free_csr(L.P); free_csr(L.R);
    free_trip(L.Pt); free_trip(L.Rt);
    cudaFree(L.meta); cudaFree(L.gci); cudaFree(L.gv); cudaFree(L.stage_group); cudaFree(L.perm);
    cudaFree(L.x); cudaFree(L.b); cudaFree(L.p1); cudaFree(L.r); cudaFree(L.dinv);
    cudaFree(L.meta2); cudaFree(L.CL); cudaFree(L.CU); cudaFree(L.VL); cudaFree(L.VU); cudaFree(L.BD); cudaFree(L.gbar); L.gbar = nullptr; cudaFree(L.empty_desc); L.empty_desc = nullptr;
    for (int tab = 0; tab < 3; tab++) { cudaFree(L.chunks[tab]); cudaFree(L.stage_chunk[tab]); L.chunks[tab] = nullptr; L.stage_chunk[tab] = nullptr; }
    L.meta2 = nullptr; L.CL = L.CU = nullptr; L.VL = L.VU = L.BD = nullptr;
    L.meta = nullptr; L.gci = nullptr; L.gv = nullptr; L.stage_group = nullptr; L.perm = nullptr;
    L.x = L.b = L.p1 = L.r = L.dinv = nullptr;
}

// One set-up stream per device: the dense inversions of an ADMM set-up (tens of interface mass matrices) are enqueued
// on it back to back and share one panel workspace in stream order.  The host does not wait for an inversion: while
// the device works, it allocates and uploads the next operator.  Consumers order themselves behind the stream with
// an event (ddpca_ldlt_create_dense) or wait for it (setup_stream_sync: ddpca_admm_finalize, ddpca_ldlt_solve*).
static cudaStream_t setup_stream(int dev)
{
    static std::mutex m;
    static cudaStream_t st[64] = {nullptr};
    std::lock_guard<std::mutex> lk(m);
    if (!st[dev & 63] && cudaStreamCreateWithFlags(&st[dev & 63], cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return st[dev & 63];
}
static int setup_stream_sync(int dev)
{
    cudaStream_t st = setup_stream(dev);
    if (!st) return fail("stream creation failed");
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return 0;
}

// In-place inversion + symmetrisation of a dense SPD matrix already on the device (set-up only):
// blocked Gauss-Jordan, three launches per 64-wide panel (kernels.cuh).  `wait`: return when the result is there;
// otherwise the work is only enqueued -- allowed on the device's set-up stream alone, whose order protects the workspace.
static int dense_invert_inplace(cudaStream_t st, int n, double *B, bool wait = true)
{
    static std::atomic<unsigned long long> attr_set{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    const size_t smem = sizeof(double) * 2 * kGjB * (kGjB + 1);
    if (!(attr_set.load() & bit)) {
        CU(cudaFuncSetAttribute(k_bgj_panels, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(cudaFuncSetAttribute(k_bgj_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set.fetch_or(bit);
    }
    // panel workspace (Dinv 64 x 64, C and R n x 64): one per device, grown on demand and kept -- an ADMM set-up
    // inverts hundreds of small interface mass matrices, and cudaMalloc / cudaFree (which synchronises) per call
    // cost more than the arithmetic
    static std::mutex ws_mtx;
    static double *ws_buf[64] = {nullptr};
    static size_t ws_len[64] = {0};
    std::lock_guard<std::mutex> ws_lock(ws_mtx);
    auto cleanup = [&]() {};
    const size_t need = (size_t)kGjB * kGjB + 2 * (size_t)n * kGjB;
    if (ws_len[dev & 63] < need) {
        CUX(cudaDeviceSynchronize());   // nothing may still be using the old workspace
        cudaFree(ws_buf[dev & 63]);
        ws_buf[dev & 63] = nullptr; ws_len[dev & 63] = 0;
        CUX(cudaMalloc(&ws_buf[dev & 63], sizeof(double) * need));
        ws_len[dev & 63] = need;
    }
    double *Dinv = ws_buf[dev & 63], *C = Dinv + kGjB * kGjB, *R = C + (size_t)n * kGjB;
    const int nt = cdiv(n, kGjB);
    // DDPCA_VERBOSE: device time per kernel class of this inversion (events around every launch)
    const bool timed = std::getenv("DDPCA_VERBOSE") != nullptr && n >= 1024;
    cudaStream_t sst = setup_stream(dev);
    if (st != sst || timed) wait = true;
    if (st != sst && sst) CUX(cudaStreamSynchronize(sst));   // inversions still pending there use the same workspace
    std::vector<cudaEvent_t> evs;
    auto mark = [&]() { if (timed) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); evs.push_back(e); } };
    for (int k0 = 0; k0 < n; k0 += kGjB) {
        const int nb = std::min(kGjB, n - k0);
        mark();
        k_bgj_diag<<<1, 1024, 0, st>>>(n, k0, nb, B, Dinv);
        mark();
        k_bgj_panels<<<nt, 256, smem, st>>>(n, k0, nb, B, Dinv, C, R);
        mark();
        k_bgj_update<<<dim3(nt, nt), 256, smem, st>>>(n, k0, nb, B, Dinv, C, R);
    }
    mark();
    dim3 g2(cdiv(n, 256), n);
    k_symmetrize<<<g2, 256, 0, st>>>(n, B);
    CUX(cudaGetLastError());
    if (!wait) return 0;
    CUX(cudaStreamSynchronize(st));   // the workspace is free again when the lock is released
    CUX(cudaGetLastError());
    if (timed) {
        double t[3] = {0, 0, 0};
        for (size_t k = 0; k + 1 < evs.size(); k++) { float ms = 0; cudaEventElapsedTime(&ms, evs[k], evs[k + 1]); t[k % 3] += ms; }
        for (cudaEvent_t e : evs) cudaEventDestroy(e);
        const double tot = t[0] + t[1] + t[2];
        std::fprintf(stderr, "ddpca set-up [dense inverse] n %d: diagonal blocks %.2f ms, panels %.2f ms, updates %.2f ms (%.2f TFLOP/s over all)\n",
                     n, t[0], t[1], t[2], tot > 0 ? 2.0 * n * (double)n * n / (tot * 1e9) : 0.0);
    }
    return 0;
}
// Dense inverse of the diagonal block [r0, r0+n) of an SPD operator given as device CSR, written to Bblk (n x n)
static int dense_spd_inverse_block(cudaStream_t st, const DevCsr &A, int r0, int n, double *Bblk, bool wait = true)
{
    CU(cudaMemsetAsync(Bblk, 0, sizeof(double) * (size_t)n * n, st));
    k_csr_to_dense<<<cdiv(n, 128), 128, 0, st>>>(A.view(), r0, n, Bblk);
    return dense_invert_inplace(st, n, Bblk, wait);
}
// ------------------------------------------------------------------------------------------
// Sparse direct solver with a host-computed factorisation: x = P^T L^-T D^-1 L^-1 P b, the
// solve phase of Eigen::SimplicialLDLT (SimplicialCholesky.h:148-171) that the reference uses
// for the interface mass matrices (MCONTACT.h:2677,2696) and the macroscopic problem (:2553).
// The two triangular solves reuse the stage machinery: I+L and I+L^T are "levels" whose LEX
// stages are the exact dependency wavefronts; forward substitution is the forward sweep from
// zero, back substitution the backward sweep.
struct ddpca_ldlt : Engine {
    int n = 0;
    long nnzL = 0;
    Level lo, up;
    int *m_in = nullptr, *m_mid = nullptr, *m_out = nullptr;  // composed index maps
    double *dinv_lo = nullptr;                                 // 1/D in lo numbering
    double *Binv = nullptr;                                    // small SPD operators: dense inverse, solve = one GEMV
    bool single_rows = false;                                  // every group is one row: k_tri_multi applies
    // dense tail: the last T rows of the elimination (one row per wavefront: a purely sequential chain)
    // are replaced by the dense inverse of their Schur complement M = L22 D2 L22^T -> one GEMV
    int tail_T = 0, tail_stage = 0, tail_g0 = 0, tail_n1 = 0;
    double *tail_Minv = nullptr, *tail_rhs = nullptr;
    int *tail_k1 = nullptr;
    // the staged solve is hundreds of tiny launches: replayed as ONE graph when called again with the same operands
    cudaGraphExec_t solve_graph = nullptr;
    const double *graph_b = nullptr;
    double *graph_x = nullptr;
    long graph_nodes = 0;
};

static void ldlt_solve_enqueue(Engine *e, ddpca_ldlt *s, const double *b_dev, double *x_dev, const int *done);
static void ldlt_solve_on(Engine *e, ddpca_ldlt *s, const double *b_dev, double *x_dev, const int *done)
{
    int n = s->n;
    if (s->Binv) {
        KL(e, DDPCA_K_COARSE, 0, 8.0 * n * (double)n + 16.0 * n, (k_dense_gemv<<<cdiv((long)n * 32, 256), 256, 0, e->stream>>>(n, s->Binv, b_dev, x_dev, done)));
        return;
    }
    if (e->profile || e->capturing || done != nullptr || std::getenv("DDPCA_NO_LDLT_GRAPH")) { ldlt_solve_enqueue(e, s, b_dev, x_dev, done); return; }
    if (s->solve_graph && (s->graph_b != b_dev || s->graph_x != x_dev)) { cudaGraphExecDestroy(s->solve_graph); s->solve_graph = nullptr; }
    if (!s->solve_graph) {
        cudaGraph_t g = nullptr;
        e->capturing = true;
        const long before = e->captured_nodes;
        cudaError_t err = cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal);
        if (err == cudaSuccess) {
            ldlt_solve_enqueue(e, s, b_dev, x_dev, nullptr);
            err = cudaStreamEndCapture(e->stream, &g);
        }
        e->capturing = false;
        s->graph_nodes = e->captured_nodes - before;
        if (err == cudaSuccess && cudaGraphInstantiate(&s->solve_graph, g, 0) == cudaSuccess) { s->graph_b = b_dev; s->graph_x = x_dev; }
        else { s->solve_graph = nullptr; cudaGetLastError(); }
        if (g) cudaGraphDestroy(g);
        if (!s->solve_graph) { ldlt_solve_enqueue(e, s, b_dev, x_dev, nullptr); return; }
    }
    if (cudaGraphLaunch(s->solve_graph, e->stream) != cudaSuccess) { if (e->launch_err.empty()) e->launch_err = "ddpca_ldlt: graph launch failed"; return; }
    e->launches += s->graph_nodes;
}
static void ldlt_solve_enqueue(Engine *e, ddpca_ldlt *s, const double *b_dev, double *x_dev, const int *done)
{
    int n = s->n;
    double tri_bytes = 12.0 * s->nnzL + 44.0 * n;
    KL(e, DDPCA_K_VECTOR, 0, 20.0 * n, (k_scatter<<<cdiv(n, 256), 256, 0, e->stream>>>(n, s->m_in, b_dev, s->lo.b)));
    (void)tri_bytes;
    if (s->single_rows) {
        // runs of small wavefronts: one CTA, rows split over warps (k_tri_multi); large wavefronts: one launch each
        const int slim = s->tail_T ? s->tail_stage : s->lo.plan.nstages();   // stages >= slim belong to the dense tail
        for (const Segment &sg0 : s->lo.segs) {
            Segment sg = sg0;
            if (sg.s0 >= slim) break;
            if (sg.s1 > slim) sg.s1 = slim;   // only multi segments can straddle the tail boundary (tail stages hold one row)
            if (sg.multi) KL(e, DDPCA_K_SWEEP_FWD, 0, sg.bytes_lo, (k_tri_multi<true><<<1, 1024, 0, e->stream>>>(s->lo.view(), s->lo.stage_group, sg.s0, sg.s1, s->lo.b, s->lo.x, done)));
            else if (s->lo.wide_rows) KL(e, DDPCA_K_SWEEP_FWD, 0, sg.bytes_lo, (k_tri_stage<true><<<sg.g1 - sg.g0, 128, 0, e->stream>>>(s->lo.view(), sg.g0, s->lo.b, s->lo.x, done)));
            else KL(e, DDPCA_K_SWEEP_FWD, 0, sg.bytes_lo, (k_sweep_fwd_stage<true><<<cdiv((long)(sg.g1 - sg.g0) * GL, 256), 256, 0, e->stream>>>(s->lo.view(), sg.g0, sg.g1, s->lo.b, s->lo.x, s->lo.p1, done)));
        }
    } else {
        sweep_fwd(e, s->lo, 0, s->lo.b, s->lo.x, true, done);
    }
    KL(e, DDPCA_K_VECTOR, 0, 28.0 * n, (k_scatter_scaled<<<cdiv(n, 256), 256, 0, e->stream>>>(n, s->m_mid, s->dinv_lo, s->lo.x, s->up.p1)));
    if (s->tail_T) {
        // x2 = (L22 D2 L22^T)^-1 (b2 - L21 y1): replaces the T sequential forward / backward steps of the tail
        const int T = s->tail_T;
        KL(e, DDPCA_K_SWEEP_FWD, 0, 0.0, (k_tail_rhs<<<cdiv((long)T * 32, 256), 256, 0, e->stream>>>(s->lo.view(), s->tail_g0, T, s->tail_k1, s->lo.b, s->lo.x, s->tail_rhs, done)));
        KL(e, DDPCA_K_COARSE, 0, 8.0 * T * (double)T, (k_dense_gemv<<<cdiv((long)T * 32, 256), 256, 0, e->stream>>>(T, s->tail_Minv, s->tail_rhs, s->up.x + s->tail_n1, done)));
    }
    if (s->single_rows) {
        const int slim = s->tail_T ? s->tail_stage : s->up.plan.nstages();
        for (int k = (int)s->up.segs.size() - 1; k >= 0; k--) {
            Segment sg = s->up.segs[k];
            if (sg.s0 >= slim) continue;
            if (sg.s1 > slim) sg.s1 = slim;
            if (sg.multi) KL(e, DDPCA_K_SWEEP_BWD, 0, sg.bytes_up, (k_tri_multi<false><<<1, 1024, 0, e->stream>>>(s->up.view(), s->up.stage_group, sg.s0, sg.s1, s->up.p1, s->up.x, done)));
            else if (s->up.wide_rows) KL(e, DDPCA_K_SWEEP_BWD, 0, sg.bytes_up, (k_tri_stage<false><<<sg.g1 - sg.g0, 128, 0, e->stream>>>(s->up.view(), sg.g0, s->up.p1, s->up.x, done)));
            else KL(e, DDPCA_K_SWEEP_BWD, 0, sg.bytes_up, (k_sweep_bwd_stage<<<cdiv((long)(sg.g1 - sg.g0) * GL, 256), 256, 0, e->stream>>>(s->up.view(), sg.g0, sg.g1, s->up.p1, s->up.x, done)));
        }
    } else {
        sweep_bwd(e, s->up, 0, s->up.x, done);
    }
    KL(e, DDPCA_K_VECTOR, 0, 20.0 * n, (k_gather<<<cdiv(n, 256), 256, 0, e->stream>>>(n, s->m_out, s->up.x, x_dev)));
}

static void ldlt_free(ddpca_ldlt *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    free_level(s->lo); free_level(s->up);
    if (s->solve_graph) cudaGraphExecDestroy(s->solve_graph);
    cudaFree(s->m_in); cudaFree(s->m_mid); cudaFree(s->m_out); cudaFree(s->dinv_lo); cudaFree(s->Binv); cudaFree(s->tail_Minv); cudaFree(s->tail_rhs); cudaFree(s->tail_k1);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    delete s;
}

static int ldlt_build(int device, int n, const int *perm, const int *Lrp, const int *Lci, const double *Lv, const double *D, ddpca_ldlt **out)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail("no CUDA device: libddpca_b200 has no CPU fallback"); }
    if (device < 0 || device >= ndev) return fail("device index out of range");
    CU(cudaSetDevice(device));
    StageTimer tm("ldlt");
    // T_lo = I + L (diagonal last in each row), T_up = I + L^T (diagonal first)
    CsrHost Tlo, Lh, Lt, Tup;
    Lh.rows = Lh.cols = n;
    Lh.rp.assign(Lrp, Lrp + n + 1);
    Lh.ci.assign(Lci, Lci + Lrp[n]);
    Lh.v.assign(Lv, Lv + Lrp[n]);
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int i = 0; i < n; i++)
        for (int p = Lrp[i]; p < Lrp[i + 1]; p++)
            if (Lci[p] >= i || Lci[p] < 0 || (p > Lrp[i] && Lci[p] <= Lci[p - 1])) bad |= 1;
    if (bad) return fail("ddpca_ldlt_create: L must be strictly lower with sorted rows");
    transpose_csr(Lh, Lt);
    auto add_diag = [&](const CsrHost &T, bool diag_last, CsrHost &o) {
        o.rows = o.cols = n;
        o.rp.assign(n + 1, 0);
        for (int i = 0; i < n; i++) o.rp[i + 1] = o.rp[i] + (T.rp[i + 1] - T.rp[i]) + 1;
        o.ci.resize(o.rp[n]);
        o.v.resize(o.rp[n]);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) {
            int q = o.rp[i];
            if (!diag_last) { o.ci[q] = i; o.v[q] = 1.0; q++; }
            for (int p = T.rp[i]; p < T.rp[i + 1]; p++, q++) { o.ci[q] = T.ci[p]; o.v[q] = T.v[p]; }
            if (diag_last) { o.ci[q] = i; o.v[q] = 1.0; }
        }
    };
    add_diag(Lh, true, Tlo);
    add_diag(Lt, false, Tup);
    ddpca_ldlt *s = new ddpca_ldlt();
    s->device = device;
    s->n = n;
    s->nnzL = Lrp[n];
    cudaDeviceGetAttribute(&s->sms, cudaDevAttrMultiProcessorCount, device);
    if (cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete s; return fail("stream creation failed"); }
    s->stream = s->own_stream;
    s->lo.wide_rows = s->up.wide_rows = ((double)Lrp[n] / n > 48.0);   // long fill-in rows
    LevelPlan tri;
    build_tri_plan(n, Lrp, Lci, tri);   // the wavefronts of both factors (DDPCA_LDLT_GENERIC_PLAN: through the general planner)
    const LevelPlan *given = std::getenv("DDPCA_LDLT_GENERIC_PLAN") ? nullptr : &tri;
    tm.lap("triangular operators + wavefront plan");
    if (setup_level(s->lo, n, Tlo.rp.data(), Tlo.ci.data(), Tlo.v.data(), DDPCA_SMOOTH_LEX, false, true, nullptr, given) ||
        setup_level(s->up, n, Tup.rp.data(), Tup.ci.data(), Tup.v.data(), DDPCA_SMOOTH_LEX, false, true, nullptr, given)) { ldlt_free(s); return 1; }
    tm.lap("two staged triangular levels");
    s->single_rows = (s->lo.plan.ngroups() == n && s->up.plan.ngroups() == n);
    std::vector<int> m_in(n), m_mid(n), m_out(n);
    std::vector<double> dinv(n);
    for (int i = 0; i < n; i++) {
        if (perm[i] < 0 || perm[i] >= n) { ldlt_free(s); return fail("ddpca_ldlt_create: bad permutation"); }
        m_in[i] = s->lo.plan.iperm[perm[i]];    // (P b)[perm[i]] = b[i], then into lo's stage numbering
        m_out[i] = s->up.plan.iperm[perm[i]];   // x[i] = y[perm[i]]
    }
    for (int j = 0; j < n; j++) {
        int o = s->lo.plan.perm[j];
        m_mid[j] = s->up.plan.iperm[o];
        dinv[j] = 1.0 / D[o];
    }
    if (upload_vec(m_in, &s->m_in) || upload_vec(m_mid, &s->m_mid) || upload_vec(m_out, &s->m_out) || upload_vec(dinv, &s->dinv_lo)) { ldlt_free(s); return 1; }
    if (s->single_rows && s->lo.plan.perm == s->up.plan.perm && !std::getenv("DDPCA_NO_DENSE_TAIL")) {
        // Trailing run of narrow wavefronts (the separators eliminated last: a nearly sequential chain of
        // thousands of stages of a few rows).  BLOCK's macroscopic factor (27 802 rows, 6 198 wavefronts):
        // the last 8 901 rows sit in 5 929 of them; as one dense block they cost one 0.63 GB GEMV per solve
        // (~0.1 ms) instead of 2 x 5 929 dependent steps, and 269 wide wavefronts remain.  The bounds keep
        // the set-up (an unblocked Gauss-Jordan inversion of the block, ~T^3 * 16 bytes of traffic) at seconds.
        const LevelPlan &pl = s->lo.plan;
        // (up to 24 576 rows = 4.8 GB: BEAM DD with 3.5 M DOF has a coarse factor of 76 860 rows whose last 17 874 rows sit in
        // 16 314 of its 16 632 wavefronts; with the round-1 limit of 9 216 rows 7 416 wavefronts remained and a solve took 36 ms)
        const int kTailWidth = 16, kTailMaxRows = 24576;
        int ns = pl.nstages(), st = ns, T = 0;
        while (st > 0) {
            const int wdt = pl.stage_group[st] - pl.stage_group[st - 1];   // single_rows: groups are rows
            if (wdt > kTailWidth || T + wdt > kTailMaxRows) break;
            T += wdt;
            st--;
        }
        if (T >= 512) {
            int g0 = pl.stage_group[st], n1 = pl.group_start[g0];
            // unit-lower L22 (the tail rows' couplings among themselves) as CSR and the number k1 of couplings of every tail
            // row to the sparse part, from the permuted I+L -- only the T tail rows are visited; the dense form is
            // assembled on the device
            CsrHost Tail;
            Tail.rows = Tail.cols = T;
            Tail.rp.assign(T + 1, 0);
            std::vector<double> D2(T);
            std::vector<int> k1(T);
#pragma omp parallel for schedule(static)
            for (int t = 0; t < T; t++) {
                const int old = pl.perm[n1 + t];
                int kk = 0, in = 0;
                for (int p = Tlo.rp[old]; p < Tlo.rp[old + 1]; p++) { if (pl.iperm[Tlo.ci[p]] < n1) kk++; else in++; }
                k1[t] = kk;
                Tail.rp[t + 1] = in;
                D2[t] = D[old];
            }
            for (int t = 0; t < T; t++) Tail.rp[t + 1] += Tail.rp[t];
            Tail.ci.resize(Tail.rp[T]);
            Tail.v.resize(Tail.rp[T]);
#pragma omp parallel for schedule(static)
            for (int t = 0; t < T; t++) {
                const int old = pl.perm[n1 + t];
                int q = Tail.rp[t];
                for (int p = Tlo.rp[old]; p < Tlo.rp[old + 1]; p++) {
                    const int c = pl.iperm[Tlo.ci[p]];
                    if (c >= n1) { Tail.ci[q] = c - n1; Tail.v[q] = Tlo.v[p]; q++; }   // includes the unit diagonal
                }
            }
            tm.lap("dense tail: host assembly");
            // assembled, multiplied and inverted on the device's set-up stream without waiting for it (the host goes on
            // with the rest of the set-up); this solver's stream is ordered behind the result by an event, temporaries are
            // stream-ordered allocations
            double *dL = nullptr, *dD = nullptr;
            DevCsr dTail;
            cudaStream_t sst = setup_stream(device);
            const size_t tnnz = (size_t)Tail.nnz();
            dTail.rows = dTail.cols = T; dTail.nnz = (long)tnnz;
            bool ok = sst != nullptr &&
                      cudaMalloc(&s->tail_Minv, sizeof(double) * (size_t)T * T) == cudaSuccess && cudaMalloc(&s->tail_rhs, sizeof(double) * T) == cudaSuccess &&
                      cudaMallocAsync((void **)&dL, sizeof(double) * (size_t)T * T, sst) == cudaSuccess && cudaMallocAsync((void **)&dD, sizeof(double) * T, sst) == cudaSuccess &&
                      cudaMallocAsync((void **)&dTail.rp, sizeof(int) * (T + 1), sst) == cudaSuccess &&
                      cudaMallocAsync((void **)&dTail.ci, sizeof(int) * std::max<size_t>(1, tnnz), sst) == cudaSuccess &&
                      cudaMallocAsync((void **)&dTail.v, sizeof(double) * std::max<size_t>(1, tnnz), sst) == cudaSuccess &&
                      cudaMemcpyAsync(dTail.rp, Tail.rp.data(), sizeof(int) * (T + 1), cudaMemcpyHostToDevice, sst) == cudaSuccess &&
                      cudaMemcpyAsync(dTail.ci, Tail.ci.data(), sizeof(int) * tnnz, cudaMemcpyHostToDevice, sst) == cudaSuccess &&
                      cudaMemcpyAsync(dTail.v, Tail.v.data(), sizeof(double) * tnnz, cudaMemcpyHostToDevice, sst) == cudaSuccess;
            if (ok) {
                cudaMemsetAsync(dL, 0, sizeof(double) * (size_t)T * T, sst);
                k_csr_to_dense<<<cdiv(T, 128), 128, 0, sst>>>(dTail.view(), 0, T, dL);
                cudaMemcpyAsync(dD, D2.data(), sizeof(double) * T, cudaMemcpyHostToDevice, sst);
                const size_t smem = sizeof(double) * 2 * kGjB * (kGjB + 1);
                ok = cudaFuncSetAttribute(k_ldl_tail_product64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess;
                dim3 gt(cdiv(T, kGjB), cdiv(T, kGjB));
                if (ok) k_ldl_tail_product64<<<gt, 256, smem, sst>>>(T, dL, dD, s->tail_Minv);
                ok = ok && dense_invert_inplace(sst, T, s->tail_Minv, /*wait=*/false) == 0 && upload_vec(k1, &s->tail_k1) == 0;
            }
            if (sst) {
                if (dTail.rp) cudaFreeAsync(dTail.rp, sst);
                if (dTail.ci) cudaFreeAsync(dTail.ci, sst);
                if (dTail.v) cudaFreeAsync(dTail.v, sst);
                if (dL) cudaFreeAsync(dL, sst);
                if (dD) cudaFreeAsync(dD, sst);
                cudaEvent_t ev;
                if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) {
                    ok = ok && cudaEventRecord(ev, sst) == cudaSuccess && cudaStreamWaitEvent(s->stream, ev, 0) == cudaSuccess;
                    cudaEventDestroy(ev);
                } else ok = false;
            }
            dTail = DevCsr();
            if (!ok && sst) cudaStreamSynchronize(sst);   // before the buffers of a failed attempt are released below
            tm.lap("dense tail: product + inverse (enqueued)");
            if (ok) { s->tail_T = T; s->tail_stage = st; s->tail_g0 = g0; s->tail_n1 = n1; }
            else { cudaFree(s->tail_Minv); cudaFree(s->tail_rhs); s->tail_Minv = s->tail_rhs = nullptr; cudaGetLastError(); }
        }
    }
    *out = s;
    return 0;
}

// ==========================================================================================
extern "C" {

const char *ddpca_last_error(void) { return g_err.c_str(); }
int ddpca_abi_version(void) { return DDPCA_ABI_VERSION; }
int ddpca_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

struct ddpca_plan { LevelPlan p; };

int ddpca_plan_create(int n, const int *rowptr, const int *colidx, int smoother_mode, ddpca_plan **out)
{
    if (!out || !rowptr || !colidx || n < 0) return fail("ddpca_plan_create: bad argument");
    ddpca_plan *pl = new ddpca_plan();
    std::string err;
    if (!build_level_plan(n, rowptr, colidx, smoother_mode, pl->p, err)) { delete pl; return fail(err); }
    *out = pl;
    return 0;
}
int ddpca_plan_create_blocks(int n, const int *rowptr, const int *colidx, int smoother_mode, int nsub, const int *sub_off, ddpca_plan **out)
{
    if (!out || !rowptr || !colidx || n < 0 || nsub < 1 || !sub_off || sub_off[0] != 0 || sub_off[nsub] != n) return fail("ddpca_plan_create_blocks: bad argument");
    ddpca_plan *pl = new ddpca_plan();
    std::string err;
    std::vector<int> off(sub_off, sub_off + nsub + 1);
    if (!build_level_plan_blocks(n, rowptr, colidx, smoother_mode, off, pl->p, err)) { delete pl; return fail(err); }
    *out = pl;
    return 0;
}
int ddpca_plan_create_tri(int n, const int *L_rowptr, const int *L_colidx, ddpca_plan **out)
{
    if (!out || !L_rowptr || !L_colidx || n < 0) return fail("ddpca_plan_create_tri: bad argument");
    for (int i = 0; i < n; i++)
        for (int p = L_rowptr[i]; p < L_rowptr[i + 1]; p++)
            if (L_colidx[p] < 0 || L_colidx[p] >= i) return fail("ddpca_plan_create_tri: L must be strictly lower");
    ddpca_plan *pl = new ddpca_plan();
    build_tri_plan(n, L_rowptr, L_colidx, pl->p);
    *out = pl;
    return 0;
}
int ddpca_plan_sizes(const ddpca_plan *p, int *n, int *ngroups, int *nstages)
{
    if (!p) return fail("null plan");
    if (n) *n = p->p.n;
    if (ngroups) *ngroups = p->p.ngroups();
    if (nstages) *nstages = p->p.nstages();
    return 0;
}
int ddpca_plan_get(const ddpca_plan *p, int *perm, int *group_start, int *stage_start)
{
    if (!p) return fail("null plan");
    if (perm) std::memcpy(perm, p->p.perm.data(), sizeof(int) * p->p.n);
    if (group_start) std::memcpy(group_start, p->p.group_start.data(), sizeof(int) * p->p.group_start.size());
    if (stage_start)
        for (int s = 0; s <= p->p.nstages(); s++) stage_start[s] = p->p.group_start[p->p.stage_group[s]];
    return 0;
}
int ddpca_plan_destroy(ddpca_plan *p) { delete p; return 0; }

int ddpca_mg_destroy(ddpca_mg *h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    for (auto &L : h->lev) free_level(L);
    cudaFree(h->Binv);
    cudaFree(h->cg_r); cudaFree(h->cg_p); cudaFree(h->cg_q); cudaFree(h->cg_z); cudaFree(h->cg_x);
    cudaFree(h->stage_a); cudaFree(h->stage_b);
    cudaFree(h->st); cudaFree(h->fl);
    cudaFree(h->sub0_off_d); cudaFree(h->binv_ptr_d); cudaFree(h->seg_d); cudaFree(h->sub_chunk_d); cudaFree(h->sub_n_d);
    if (h->st_host) cudaFreeHost(h->st_host);
    if (h->fl_host) cudaFreeHost(h->fl_host);
    for (int k = 0; k < 3; k++) cudaFree(h->partial[k]);
    cudaFree(h->ks);
    for (int k = 0; k < 2; k++) if (h->iter_graph[k]) cudaGraphExecDestroy(h->iter_graph[k]);
    for (int k = 0; k < 2; k++) if (h->solve_graph[k]) cudaGraphExecDestroy(h->solve_graph[k]);
    for (int k = 0; k < 4; k++) if (h->ev[k]) cudaEventDestroy(h->ev[k]);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return 0;
}

// block-diagonal concatenation of the subs' operators of one level (rows r[s], columns c[s])
static int concat_block_diag(int nsub, const int *r, const int *c, const int *const *rp, const int *const *ci, const double *const *v,
                             CsrHost &out)
{
    long rows = 0, cols = 0, nnz = 0;
    for (int s = 0; s < nsub; s++) { rows += r[s]; cols += c[s]; nnz += rp[s][r[s]]; }
    if (rows > 0x7ffffff0L || cols > 0x7ffffff0L || nnz > 0x7ffffff0L) return fail("batch too large for 32-bit indices: split it into several batches");
    out.rows = (int)rows; out.cols = (int)cols;
    out.rp.resize(rows + 1);
    out.ci.resize(nnz);
    out.v.resize(nnz);
    long r0 = 0, c0 = 0, p0 = 0;
    out.rp[0] = 0;
    for (int s = 0; s < nsub; s++) {
        const int nr = r[s];
        const long nz = rp[s][nr];
        for (int i = 0; i < nr; i++) out.rp[r0 + i + 1] = (int)(p0 + rp[s][i + 1]);
        const int *cs = ci[s];
        const double *vs = v[s];
        const int coff = (int)c0;
#pragma omp parallel for schedule(static)
        for (long p = 0; p < nz; p++) { out.ci[p0 + p] = cs[p] + coff; out.v[p0 + p] = vs[p]; }
        r0 += nr; c0 += c[s]; p0 += nz;
    }
    return 0;
}

// The caller's hierarchies as handed to ddpca_mg_create[_batch]: arrays indexed [s * nlevels + l] (prolongations [s * (nlevels-1) + l])
struct HierInput {
    int nsub, nlevels;
    const int *n;
    const int *const *rowptr, *const *colidx;
    const double *const *val;
    const int *const *P_rowptr, *const *P_colidx;
    const double *const *P_val;
};
// host half of level l: the subs' operators as the blocks of one block-diagonal level (nothing is concatenated)
static int level_host_pass(const HierInput &in, int l, int smoother_mode, const std::vector<std::vector<int>> &sub_off, Level &L, LevelHost &D)
{
    const int nsub = in.nsub, nlevels = in.nlevels;
    const int mode = (l == 0) ? -1 : smoother_mode;   // level 0 is only ever solved directly
    const bool coarse_only = (l == 0 && nlevels > 1);
    CsrBlocks B;
    B.nsub = nsub;
    B.row_off = sub_off[l]; B.col_off = sub_off[l];
    for (int s = 0; s < nsub; s++) { B.rp.push_back(in.rowptr[s * nlevels + l]); B.ci.push_back(in.colidx[s * nlevels + l]); B.v.push_back(in.val[s * nlevels + l]); }
    if (B.nnz() > 0x7ffffff0L) return fail("batch too large for 32-bit indices: split it into several batches");
    if (nsub > 1 && std::getenv("DDPCA_SETUP_CONCAT")) {   // cross-check: plan and permute the explicitly concatenated level
        CsrHost cat;
        std::vector<int> rs(nsub);
        for (int s = 0; s < nsub; s++) rs[s] = in.n[s * nlevels + l];
        const double ta = StageTimer::now();
        if (concat_block_diag(nsub, rs.data(), rs.data(), B.rp.data(), B.ci.data(), B.v.data(), cat)) return 1;
        g_clock.concat += StageTimer::now() - ta;
        return setup_level_host(L, CsrBlocks::single(cat.rows, cat.cols, cat.rp.data(), cat.ci.data(), cat.v.data()), mode, !coarse_only, nullptr, D, &sub_off[l]);
    }
    return setup_level_host(L, B, mode, /*group_layout=*/!coarse_only, nullptr, D);
}
// host half of the transfer pair of level l >= 1: realProl[l-1] (n_l x n_{l-1}) with rows in level l's numbering and
// columns in level l-1's, its transpose, and the node-triple forms of both
struct TransferHost { CsrHost Pp, Rp; TripHost Pt, Rt; };
static int transfer_host_pass(const HierInput &in, int l, const std::vector<std::vector<int>> &sub_off, const LevelPlan &fine, const LevelPlan &coarse, TransferHost &T)
{
    const int nsub = in.nsub, nlevels = in.nlevels;
    const double t1 = StageTimer::now();
    CsrBlocks B;
    B.nsub = nsub;
    B.row_off = sub_off[l]; B.col_off = sub_off[l - 1];
    for (int s = 0; s < nsub; s++) {
        B.rp.push_back(in.P_rowptr[s * (nlevels - 1) + l - 1]); B.ci.push_back(in.P_colidx[s * (nlevels - 1) + l - 1]); B.v.push_back(in.P_val[s * (nlevels - 1) + l - 1]);
    }
    if (B.nnz() > 0x7ffffff0L) return fail("batch too large for 32-bit indices: split it into several batches");
    permute_csr_blocks(B, fine.perm, coarse.iperm, T.Pp);
    // the reference stores realProl with explicit zeros (a 3x3 block per node pair, diagonal only
    // non-zero unless the nodes carry rotations): x + 0 * y == x, so they are dropped
    drop_zeros_csr(T.Pp);
    const double t2 = StageTimer::now();
    g_clock.tr_permute += t2 - t1;
    transpose_csr(T.Pp, T.Rp);
    const double t3 = StageTimer::now();
    g_clock.tr_transpose += t3 - t2;
    build_trip_host(T.Pp, T.Pt);
    build_trip_host(T.Rp, T.Rt);
    g_clock.tr_trip += StageTimer::now() - t3;
    return 0;
}

// nsub hierarchies of `nlevels` levels each; every array is indexed [s * nlevels + l] (prolongations [s * (nlevels-1) + l])
static int mg_create_impl(int device, int nsub, int nlevels, const int *n, const int *const *rowptr, const int *const *colidx,
                          const double *const *val, const int *const *P_rowptr, const int *const *P_colidx,
                          const double *const *P_val, int smoother_mode, ddpca_mg **out, bool no_direct = false)
{
    if (!out || nsub < 1 || nlevels < 1 || nlevels > 16 || !n || !rowptr || !colidx || !val) return fail("ddpca_mg_create: bad argument");
    if (smoother_mode != DDPCA_SMOOTH_LEX && smoother_mode != DDPCA_SMOOTH_MC) return fail("unknown smoother mode");
    if (nlevels > 1 && (!P_rowptr || !P_colidx || !P_val)) return fail("prolongation operators missing");
    int ndev = ddpca_device_count();
    if (ndev == 0) return fail("no CUDA device: libddpca_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail("device index out of range");
    CU(cudaSetDevice(device));
    ddpca_mg *h = new ddpca_mg();
    h->device = device;
    h->mode = smoother_mode;
    h->nlev = nlevels;
    h->nsub = nsub;
    h->lev.resize(nlevels);
    h->sub_off.assign(nlevels, std::vector<int>(nsub + 1, 0));
    cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, device);
#define FAILC(expr) do { if (expr) { ddpca_mg_destroy(h); return 1; } } while (0)
#define CUC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_err = std::string(#call) + ": " + cudaGetErrorString(e_); ddpca_mg_destroy(h); return 1; } } while (0)
    CUC(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    for (int k = 0; k < 4; k++) CUC(cudaEventCreate(&h->ev[k]));
    for (int l = 0; l < nlevels; l++) {
        long acc = 0;
        for (int s = 0; s < nsub; s++) { h->sub_off[l][s] = (int)acc; acc += n[s * nlevels + l]; }
        if (acc > 0x7ffffff0L) { ddpca_mg_destroy(h); return fail("batch too large for 32-bit indices: split it into several batches"); }
        h->sub_off[l][nsub] = (int)acc;
    }
    StageTimer tm("hierarchy");
    g_clock = SetupClock();
    int nmax = 0;
    HierInput in{nsub, nlevels, n, rowptr, colidx, val, P_rowptr, P_colidx, P_val};
    for (int l = 0; l < nlevels; l++) {
        Level &L = h->lev[l];
        nmax = std::max(nmax, h->sub_off[l][nsub]);
        LevelHost D;
        if (level_host_pass(in, l, smoother_mode, h->sub_off, L, D)) {
            g_err = "level " + std::to_string(l) + ": " + g_err;
            ddpca_mg_destroy(h);
            return 1;
        }
        FAILC(setup_level_upload(L, D, /*keep_csr=*/l == 0 && !no_direct));
        if (l >= 1) {
            TransferHost T;
            FAILC(transfer_host_pass(in, l, h->sub_off, L.plan, h->lev[l - 1].plan, T));
            const double tu = StageTimer::now();
            FAILC(upload_csr(T.Pp, L.P));
            FAILC(upload_csr(T.Rp, L.R));
            FAILC(upload_trip(T.Pp, T.Pt, L.Pt) || upload_trip(T.Rp, T.Rt, L.Rt));
            g_clock.tr_upload += StageTimer::now() - tu;
        }
    }
    if (tm.on) {
        const SetupClock &c = g_clock;
        std::fprintf(stderr, "ddpca set-up [hierarchy] %d subs, %d levels, %d rows: concatenate %.3f, plan %.3f, permute %.3f, layout %.3f, upload %.3f | transfers: permute %.3f, transpose %.3f, triples %.3f, upload %.3f s\n",
                     nsub, nlevels, nmax, c.concat, c.plan, c.permute, c.layout, c.upload, c.tr_permute, c.tr_transpose, c.tr_trip, c.tr_upload);
    }
    tm.lap("levels");
    FAILC(alloc_vec(&h->cg_r, nmax) || alloc_vec(&h->cg_p, nmax) || alloc_vec(&h->cg_q, nmax) || alloc_vec(&h->cg_z, nmax) ||
          alloc_vec(&h->cg_x, nmax) || alloc_vec(&h->stage_a, nmax) || alloc_vec(&h->stage_b, nmax));
    CUC(cudaMalloc(&h->st, sizeof(PcgState) * nsub));
    CUC(cudaMemset(h->st, 0, sizeof(PcgState) * nsub));
    if (nsub > 1 && !std::getenv("DDPCA_NO_SKIP_FROZEN"))
        for (Level &L : h->lev) L.sub_state = h->st;   // v2 passes skip the chunks of converged subdomains
    CUC(cudaMalloc(&h->fl, sizeof(BatchFlags)));
    CUC(cudaMemset(h->fl, 0, sizeof(BatchFlags)));
    CUC(cudaMallocHost(&h->st_host, sizeof(PcgState) * nsub));
    CUC(cudaMallocHost(&h->fl_host, sizeof(BatchFlags) * (kDepth + 2)));
    std::memset(h->st_host, 0, sizeof(PcgState) * nsub);
    std::memset(h->fl_host, 0, sizeof(BatchFlags) * (kDepth + 2));
    {   // ---- segmented-reduction tables of the finest level: rows of one sub are contiguous inside a stage ----
        const int Lf = nlevels - 1;
        const Level &L = h->lev[Lf];
        const std::vector<int> &off = h->sub_off[Lf];
        h->sub_n.resize(nsub);
        for (int s = 0; s < nsub; s++) h->sub_n[s] = off[s + 1] - off[s];
        std::vector<std::vector<SegChunk>> per_sub(nsub);
        int cur = -1, start = 0, sub = 0;
        auto flush = [&](int end) {
            for (int r0 = start; r0 < end; r0 += kSegRows) per_sub[cur].push_back(SegChunk{r0, std::min(kSegRows, end - r0), cur, 0});
        };
        for (int i = 0; i < L.n; i++) {
            const int old = L.plan.perm[i];
            if (nsub > 1) { while (old >= off[sub + 1]) sub++; while (old < off[sub]) sub--; } else sub = 0;
            if (sub != cur) { if (cur >= 0) flush(i); cur = sub; start = i; }
        }
        if (cur >= 0) flush(L.n);
        std::vector<SegChunk> seg;
        std::vector<int> sub_chunk(nsub + 1, 0);
        for (int s = 0; s < nsub; s++) { sub_chunk[s] = (int)seg.size(); seg.insert(seg.end(), per_sub[s].begin(), per_sub[s].end()); }
        sub_chunk[nsub] = (int)seg.size();
        h->nseg = (int)seg.size();
        if (h->nseg == 0) { seg.push_back(SegChunk{0, 0, 0, 0}); }
        FAILC(upload_vec(seg, &h->seg_d) || upload_vec(sub_chunk, &h->sub_chunk_d) || upload_vec(h->sub_n, &h->sub_n_d));
        for (int k = 0; k < 3; k++) {
            CUC(cudaMalloc(&h->partial[k], sizeof(double) * std::max(h->nseg, kNumPart)));
            CUC(cudaMemset(h->partial[k], 0, sizeof(double) * std::max(h->nseg, kNumPart)));
        }
    }
    if (!no_direct) {   // ---- level 0: dense inverse of every sub's block (the reference re-factorises in every CG_SOLV call, MGPIS.h:185) ----
        Level &L0 = h->lev[0];
        h->n0 = L0.n;
        std::vector<long long> boff(nsub);
        long long tot = 0;
        for (int s = 0; s < nsub; s++) {
            const long long ns = h->sub_off[0][s + 1] - h->sub_off[0][s];
            if (ns > 32768) { ddpca_mg_destroy(h); return fail("level 0 has " + std::to_string(ns) + " rows; the dense direct solver supports <= 32768"); }
            boff[s] = tot;
            tot += ns * ns + (ns & 1);   // keep every block 16-byte aligned
            h->coarse_bytes += 8.0 * ns * (double)ns + 16.0 * ns;
        }
        CUC(cudaMalloc(&h->Binv, sizeof(double) * (size_t)std::max<long long>(tot, 1)));
        for (int s = 0; s < nsub; s++)
            FAILC(dense_spd_inverse_block(h->stream, L0.A, h->sub_off[0][s], h->sub_off[0][s + 1] - h->sub_off[0][s], h->Binv + boff[s]));
        std::vector<double *> bptr(nsub);
        for (int s = 0; s < nsub; s++) bptr[s] = h->Binv + boff[s];
        FAILC(upload_vec(h->sub_off[0], &h->sub0_off_d) || upload_vec(bptr, &h->binv_ptr_d));
    }
    tm.lap("level-0 dense inverses");
    // capture + instantiate the V-cycle-preconditioned solve graph now (set-up time), not in the first solve
    // (no_direct: a Jacobi-PCG-only handle, e.g. the interface mass matrices of the ADMM loop; its graph is built on first use)
    if (!no_direct) {
        build_solve_graph(h, 1);
        if (h->while_state[1] != 1) FAILC(build_iter_graph(h, 1));
    }
#undef FAILC
#undef CUC
    *out = h;
    return 0;
}

int ddpca_mg_create(int device, int nlevels, const int *n, const int *const *rowptr, const int *const *colidx,
                    const double *const *val, const int *const *P_rowptr, const int *const *P_colidx,
                    const double *const *P_val, int smoother_mode, ddpca_mg **out)
{
    return mg_create_impl(device, 1, nlevels, n, rowptr, colidx, val, P_rowptr, P_colidx, P_val, smoother_mode, out);
}

int ddpca_mg_create_batch(int device, int nsub, int nlevels, const int *n, const int *const *rowptr, const int *const *colidx,
                          const double *const *val, const int *const *P_rowptr, const int *const *P_colidx,
                          const double *const *P_val, int smoother_mode, ddpca_mg **out)
{
    return mg_create_impl(device, nsub, nlevels, n, rowptr, colidx, val, P_rowptr, P_colidx, P_val, smoother_mode, out);
}

// The host half of ddpca_mg_create_batch without a device: plans, permuted operators, layouts, chunk tables and
// transfer forms are built exactly as for a real handle and dropped.  Reports what would be uploaded and where the
// host time goes; seconds[9] = concatenate, plan, permute, layout, (upload: 0), transfers: permute+drop, transpose,
// triples, (upload: 0).
int ddpca_mg_setup_dryrun(int nsub, int nlevels, const int *n, const int *const *rowptr, const int *const *colidx,
                          const double *const *val, const int *const *P_rowptr, const int *const *P_colidx,
                          const double *const *P_val, int smoother_mode, double *seconds, long *device_bytes, int *v2_levels,
                          unsigned long long *checksum)
{
    if (nsub < 1 || nlevels < 1 || nlevels > 16 || !n || !rowptr || !colidx || !val) return fail("ddpca_mg_setup_dryrun: bad argument");
    if (smoother_mode != DDPCA_SMOOTH_LEX && smoother_mode != DDPCA_SMOOTH_MC) return fail("unknown smoother mode");
    if (nlevels > 1 && (!P_rowptr || !P_colidx || !P_val)) return fail("prolongation operators missing");
    std::vector<std::vector<int>> sub_off(nlevels, std::vector<int>(nsub + 1, 0));
    for (int l = 0; l < nlevels; l++) {
        long acc = 0;
        for (int s = 0; s < nsub; s++) { sub_off[l][s] = (int)acc; acc += n[s * nlevels + l]; }
        if (acc > 0x7ffffff0L) return fail("batch too large for 32-bit indices: split it into several batches");
        sub_off[l][nsub] = (int)acc;
    }
    g_clock = SetupClock();
    HierInput in{nsub, nlevels, n, rowptr, colidx, val, P_rowptr, P_colidx, P_val};
    std::vector<Level> lev(nlevels);
    long bytes = 0;
    int nv2 = 0;
    unsigned long long hash = 1469598103934665603ULL;   // FNV-1a over everything that would be uploaded, 8 bytes at a time
    auto vb = [&](const auto &v) {
        const size_t nb = v.size() * sizeof(v[0]);
        if (checksum) {
            const unsigned char *p = reinterpret_cast<const unsigned char *>(v.data());
            size_t k = 0;
            for (; k + 8 <= nb; k += 8) { unsigned long long w; std::memcpy(&w, p + k, 8); hash = (hash ^ w) * 1099511628211ULL; }
            for (; k < nb; k++) hash = (hash ^ p[k]) * 1099511628211ULL;
        }
        return (long)nb;
    };
    for (int l = 0; l < nlevels; l++) {
        LevelHost D;
        if (level_host_pass(in, l, smoother_mode, sub_off, lev[l], D)) { g_err = "level " + std::to_string(l) + ": " + g_err; return 1; }
        if (D.group_layout) {
            if (lev[l].v2) {
                nv2++;
                bytes += vb(D.H2.meta) + vb(D.H2.CL) + vb(D.H2.CU) + vb(D.H2.VL) + vb(D.H2.VU) + vb(D.H2.BD);
                for (int tab = 0; tab < 3; tab++) bytes += vb(D.H2.chunks[tab]) + vb(D.H2.stage_chunk[tab]);
            } else bytes += vb(D.G.meta) + vb(D.G.ci) + vb(D.G.v);
        } else bytes += vb(D.Ap.rp) + vb(D.Ap.ci) + vb(D.Ap.v);
        bytes += vb(lev[l].plan.perm) + vb(lev[l].plan.stage_group);
        bytes += 4L * 8 * (lev[l].n + 1);
        if (l >= 1) {
            TransferHost T;
            if (transfer_host_pass(in, l, sub_off, lev[l].plan, lev[l - 1].plan, T)) return 1;
            bytes += vb(T.Pp.rp) + vb(T.Pp.ci) + vb(T.Pp.v) + vb(T.Rp.rp) + vb(T.Rp.ci) + vb(T.Rp.v);
            for (const TripHost *t : {&T.Pt, &T.Rt})
                if (t->ok) bytes += vb(t->trow) + vb(t->tptr) + vb(t->tcol) + vb(t->tw) + vb(t->rest_row) + vb(t->R.ci) + vb(t->R.v);
        }
    }
    if (checksum) *checksum = hash;
    if (seconds) {
        const SetupClock &c = g_clock;
        const double t[9] = {c.concat, c.plan, c.permute, c.layout, 0.0, c.tr_permute, c.tr_transpose, c.tr_trip, 0.0};
        for (int k = 0; k < 9; k++) seconds[k] = t[k];
    }
    if (device_bytes) *device_bytes = bytes;
    if (v2_levels) *v2_levels = nv2;
    return 0;
}

int ddpca_mg_batch_result(const ddpca_mg *h, int *nsub, long *iters, double *resid, double *tol_abs)
{
    if (!h) return fail("null handle");
    if (nsub) *nsub = h->nsub;
    for (int s = 0; s < h->nsub; s++) {
        if (iters) iters[s] = (long)h->st_host[s].it;
        if (resid) resid[s] = std::sqrt(h->st_host[s].rr);
        if (tol_abs) tol_abs[s] = h->st_host[s].tol;
    }
    return 0;
}

int ddpca_mg_set_stream(ddpca_mg *h, void *stream)
{
    if (!h) return fail("null handle");
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    return 0;
}

int ddpca_mg_pcg_dev(ddpca_mg *h, int prec, const double *b_dev, double *x_dev, double rel_tol, long maxit,
                     long *iters, double *resid, double *tol_abs)
{
    if (!h || !b_dev || !x_dev) return fail("ddpca_mg_pcg_dev: bad argument");
    CU(cudaSetDevice(h->device));
    CU(cudaEventRecord(h->ev[0], h->stream));
    if (pcg_device(h, prec, b_dev, x_dev, rel_tol, maxit, iters, resid, tol_abs)) return 1;
    CU(cudaEventRecord(h->ev[1], h->stream));
    CU(cudaEventSynchronize(h->ev[1]));
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
    h->t_solve = ms; h->t_h2d = 0; h->t_d2h = 0;
    return 0;
}

int ddpca_mg_pcg(ddpca_mg *h, int prec, const double *b, double *x, double rel_tol, long maxit, long *iters,
                 double *resid, double *tol_abs)
{
    if (!h || !b || !x) return fail("ddpca_mg_pcg: bad argument");
    CU(cudaSetDevice(h->device));
    int n = h->lev[h->nlev - 1].n;
    CU(cudaEventRecord(h->ev[0], h->stream));
    CU(cudaMemcpyAsync(h->stage_a, b, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
    CU(cudaEventRecord(h->ev[1], h->stream));
    if (pcg_device(h, prec, h->stage_a, h->stage_b, rel_tol, maxit, iters, resid, tol_abs)) return 1;
    CU(cudaEventRecord(h->ev[2], h->stream));
    CU(cudaMemcpyAsync(x, h->stage_b, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaEventRecord(h->ev[3], h->stream));
    CU(cudaEventSynchronize(h->ev[3]));
    float a = 0, s = 0, d = 0;
    cudaEventElapsedTime(&a, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&s, h->ev[1], h->ev[2]);
    cudaEventElapsedTime(&d, h->ev[2], h->ev[3]);
    h->t_h2d = a; h->t_solve = s; h->t_d2h = d;
    return 0;
}

// helper for the single-operator entry points: host vector(s) in reference numbering of
// level `lin` -> device numbering, run, result of level `lout` back
static int to_dev(ddpca_mg *h, int l, const double *host, double *dev_perm)
{
    Level &L = h->lev[l];
    CU(cudaMemcpyAsync(h->stage_a, host, sizeof(double) * L.n, cudaMemcpyHostToDevice, h->stream));
    k_gather<<<cdiv(L.n, 256), 256, 0, h->stream>>>(L.n, L.perm, h->stage_a, dev_perm);
    CU(cudaMemsetAsync(dev_perm + L.n, 0, sizeof(double), h->stream));   // the always-zero slot of THIS level (the vectors are sized for the largest)
    return 0;
}
static int to_host(ddpca_mg *h, int l, const double *dev_perm, double *host)
{
    Level &L = h->lev[l];
    k_scatter<<<cdiv(L.n, 256), 256, 0, h->stream>>>(L.n, L.perm, dev_perm, h->stage_b);
    CU(cudaMemcpyAsync(host, h->stage_b, sizeof(double) * L.n, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return 0;
}

int ddpca_mg_vcycle(ddpca_mg *h, int level, const double *b, double *x)
{
    if (!h || level < 0 || level >= h->nlev || !b || !x) return fail("ddpca_mg_vcycle: bad argument");
    CU(cudaSetDevice(h->device));
    // use cg_r / cg_z as the top-level b / x so that lev[level].b/x stay free for the recursion
    if (to_dev(h, level, b, h->cg_r)) return 1;
    if (to_dev(h, level, x, h->cg_z)) return 1;
    vcycle_dev(h, level, h->cg_r, h->cg_z, false, nullptr);
    if (h->profile) h->prof_collect();
    return to_host(h, level, h->cg_z, x);
}

int ddpca_mg_spmv(ddpca_mg *h, int level, const double *x, double *y)
{
    if (!h || level < 0 || level >= h->nlev || !x || !y) return fail("ddpca_mg_spmv: bad argument");
    CU(cudaSetDevice(h->device));
    if (to_dev(h, level, x, h->cg_p)) return 1;
    if (h->lev[level].meta || h->lev[level].v2) launch_level_spmv(h, h->lev[level], level, h->cg_p, h->cg_q, nullptr, nullptr, nullptr);
    else launch_spmv(h, DDPCA_K_SPMV, level, h->lev[level].A, h->cg_p, h->cg_q, false, nullptr, nullptr, nullptr);
    if (h->profile) h->prof_collect();
    return to_host(h, level, h->cg_q, y);
}

int ddpca_mg_restrict(ddpca_mg *h, int level, const double *r_fine, double *r_coarse)
{
    if (!h || level < 0 || level + 1 >= h->nlev || !r_fine || !r_coarse) return fail("ddpca_mg_restrict: bad argument");
    CU(cudaSetDevice(h->device));
    if (to_dev(h, level + 1, r_fine, h->cg_p)) return 1;
    launch_transfer(h, DDPCA_K_RESTRICT, level + 1, h->lev[level + 1].R, h->lev[level + 1].Rt, h->cg_p, h->cg_q, false, nullptr);
    if (h->profile) h->prof_collect();
    return to_host(h, level, h->cg_q, r_coarse);
}

int ddpca_mg_prolong_add(ddpca_mg *h, int level, const double *e_coarse, double *x_fine)
{
    if (!h || level < 0 || level + 1 >= h->nlev || !e_coarse || !x_fine) return fail("ddpca_mg_prolong_add: bad argument");
    CU(cudaSetDevice(h->device));
    if (to_dev(h, level, e_coarse, h->cg_p)) return 1;
    if (to_dev(h, level + 1, x_fine, h->cg_q)) return 1;
    launch_transfer(h, DDPCA_K_PROLONG, level + 1, h->lev[level + 1].P, h->lev[level + 1].Pt, h->cg_p, h->cg_q, true, nullptr);
    if (h->profile) h->prof_collect();
    return to_host(h, level + 1, h->cg_q, x_fine);
}

int ddpca_mg_coarse_solve(ddpca_mg *h, const double *b, double *x)
{
    if (!h || !b || !x) return fail("ddpca_mg_coarse_solve: bad argument");
    CU(cudaSetDevice(h->device));
    if (to_dev(h, 0, b, h->cg_p)) return 1;
    vcycle_dev(h, 0, h->cg_p, h->cg_q, true, nullptr);
    if (h->profile) h->prof_collect();
    return to_host(h, 0, h->cg_q, x);
}

// ---- host-driven drivers on the same kernels: scalars come back to the host after every
// reduction.  Kept as the cross-check of the device-resident drivers further down (DDPCA_KRYLOV_HOSTLOOP=1).
static int dev_dot(ddpca_mg *h, int n, const double *a, const double *b, double *out)
{
    int gv = vec_grid(h, n);
    KL(h, DDPCA_K_VECTOR, h->nlev - 1, 16.0 * n, (k_dot<<<gv, 256, 0, h->stream>>>(n, a, b, h->partial[0], nullptr)));
    std::vector<double> part(gv);
    CU(cudaMemcpyAsync(part.data(), h->partial[0], sizeof(double) * gv, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    double s = 0.0;
    for (int k = 0; k < gv; k++) s += part[k];
    *out = s;
    return 0;
}
static void dev_axpby(ddpca_mg *h, int n, double a, const double *x, double b, double *y)
{
    KL(h, DDPCA_K_VECTOR, h->nlev - 1, 24.0 * n, (k_axpby<<<cdiv(n, 256), 256, 0, h->stream>>>(n, a, x, b, y)));
}

// MGPIS::MULT_SOLV, MGPIS.h:130-160: V-cycle iteration until the residual norm stagnates
static int mult_solv_hostloop(ddpca_mg *h, const double *b, double *x, long *iters, double *resid)
{
    int Lf = h->nlev - 1;
    Level &L = h->lev[Lf];
    int n = L.n;
    double *bd = h->cg_r, *xd = h->cg_x, *rd = h->cg_q, *ax = h->cg_p;
    if (to_dev(h, Lf, b, bd)) return 1;
    CU(cudaMemsetAsync(xd, 0, sizeof(double) * n, h->stream));     // :133
    const long maxiNumb = 10000;                                   // :134
    double moni[5] = {0, 0, 0, 0, 0};
    long it = 0;
    while (it < maxiNumb) {
        vcycle_dev(h, Lf, bd, xd, false, nullptr);                 // :143
        launch_level_spmv(h, L, Lf, xd, ax, nullptr, nullptr, nullptr);
        CU(cudaMemcpyAsync(rd, bd, sizeof(double) * n, cudaMemcpyDeviceToDevice, h->stream));
        dev_axpby(h, n, -1.0, ax, 1.0, rd);                        // :144
        double rr;
        if (dev_dot(h, n, rd, rd, &rr)) return 1;
        moni[it % 5] = std::sqrt(rr);                              // :146
        if (it >= 4) {                                             // :147-153
            double mx = *std::max_element(moni, moni + 5), mn = *std::min_element(moni, moni + 5);
            if (mx - mn < 0.1 * ((mx + mn) / 2.0)) break;
        }
        it++;
    }
    if (h->profile) h->prof_collect();
    if (iters) *iters = it;
    if (resid) *resid = moni[it % 5];
    return to_host(h, Lf, xd, x);
}

// MGPIS::GMRES_SOLV, MGPIS.h:227-348: left-preconditioned restarted GMRES(10); Arnoldi vectors and
// all matrix / preconditioner work on the device, the 11x10 Hessenberg algebra on the host.
static int gmres_hostloop(ddpca_mg *h, int prec, const double *b, double *x, long *iters, double *resid, double *tol_abs)
{
    enum { STAG = 10 };
    int Lf = h->nlev - 1;
    Level &L = h->lev[Lf];
    int n = L.n;
    if (prec == 0 && !L.dinv) {
        CU(cudaMalloc(&L.dinv, sizeof(double) * n));
        if (L.v2) k_extract_diag_inv2<<<cdiv(L.ng, 256), 256, 0, h->stream>>>(L.view2(), L.dinv);
        else k_extract_diag_inv<<<cdiv(L.ng, 256), 256, 0, h->stream>>>(L.view(), L.dinv);
    }
    // Arnoldi vectors with stride n+1: each keeps the always-zero slot the padded gathers rely on (alloc_vec)
    const size_t ldv = (size_t)n + 1;
    double *V = nullptr, *x0 = nullptr, *bd = nullptr;
    auto cleanup = [&]() { cudaFree(V); cudaFree(x0); cudaFree(bd); };
    if (alloc_vec(&V, (long)(ldv * (STAG + 1)) - 1) || alloc_vec(&x0, n) || alloc_vec(&bd, n)) { cleanup(); return 1; }
    double *r = h->cg_r, *w0 = h->cg_q, *w = h->cg_z, *xd = h->cg_x;
    if (to_dev(h, Lf, b, bd)) { cleanup(); return 1; }
    CUX(cudaMemsetAsync(xd, 0, sizeof(double) * n, h->stream));                       // :248
    double bb;
    if (dev_dot(h, n, bd, bd, &bb)) { cleanup(); return 1; }
    const double tol = 1.0E-12 * std::sqrt(bb);                                      // :250
    const long maxiNumb = n;                                                         // :249
    double H[STAG + 1][STAG], Q[STAG + 1][STAG], R[STAG][STAG], moni[STAG];
    std::memset(moni, 0, sizeof(moni));
    double normR0 = 0.0;
    long it = 0;
    int rc = 0;
    while (it < maxiNumb) {                                                          // :261
        int k = (int)(it % STAG);
        if (k == 0) {                                                                // restart, :263-276
            CUX(cudaMemcpyAsync(x0, xd, sizeof(double) * n, cudaMemcpyDeviceToDevice, h->stream));
            launch_level_spmv(h, L, Lf, x0, w0, nullptr, nullptr, nullptr);
            CUX(cudaMemcpyAsync(r, bd, sizeof(double) * n, cudaMemcpyDeviceToDevice, h->stream));
            dev_axpby(h, n, -1.0, w0, 1.0, r);
            precondition(h, prec, r, w, nullptr);
            double ww;
            if ((rc = dev_dot(h, n, w, w, &ww))) break;
            normR0 = std::sqrt(ww);
            dev_axpby(h, n, 1.0 / normR0, w, 0.0, V);
            std::memset(H, 0, sizeof(H)); std::memset(Q, 0, sizeof(Q)); std::memset(R, 0, sizeof(R));
        }
        launch_level_spmv(h, L, Lf, V + (size_t)k * ldv, w0, nullptr, nullptr, nullptr);   // :278
        precondition(h, prec, w0, w, nullptr);                                           // :279-285
        for (int j = 0; j <= k && !rc; j++) rc = dev_dot(h, n, V + (size_t)j * ldv, w, &H[j][k]);   // :286
        if (rc) break;
        for (int j = 0; j <= k; j++) dev_axpby(h, n, -H[j][k], V + (size_t)j * ldv, 1.0, w);        // :287
        double qq;
        if ((rc = dev_dot(h, n, w, w, &qq))) break;
        const double nq = std::sqrt(qq);                                                 // :288
        H[k + 1][k] = nq;
        dev_axpby(h, n, 1.0 / nq, w, 0.0, V + (size_t)(k + 1) * ldv);                      // :294-296
        {   // QR of the Hessenberg matrix by Gram-Schmidt, one column per step, :297-316
            double col[STAG + 1];
            for (int i = 0; i <= k + 1; i++) col[i] = H[i][k];
            for (int j = 0; j < k; j++) {
                double s = 0.0;
                for (int i = 0; i <= k + 1; i++) s += Q[i][j] * H[i][k];
                R[j][k] = s;
            }
            for (int j = 0; j < k; j++) for (int i = 0; i <= k + 1; i++) col[i] -= Q[i][j] * R[j][k];
            double nc = 0.0;
            for (int i = 0; i <= k + 1; i++) nc += col[i] * col[i];
            nc = std::sqrt(nc);
            R[k][k] = nc;
            for (int i = 0; i <= k + 1; i++) Q[i][k] = col[i] / nc;
        }
        double y[STAG];                                                                  // :317-324
        for (int j = k; j >= 0; j--) {
            double s = normR0 * Q[0][j];
            for (int c = j + 1; c <= k; c++) s -= R[j][c] * y[c];
            y[j] = s / R[j][j];
        }
        CUX(cudaMemcpyAsync(xd, x0, sizeof(double) * n, cudaMemcpyDeviceToDevice, h->stream));   // :325
        for (int j = 0; j <= k; j++) dev_axpby(h, n, y[j], V + (size_t)j * ldv, 1.0, xd);
        launch_level_spmv(h, L, Lf, xd, w0, nullptr, nullptr, nullptr);                  // :326
        CUX(cudaMemcpyAsync(r, bd, sizeof(double) * n, cudaMemcpyDeviceToDevice, h->stream));
        dev_axpby(h, n, -1.0, w0, 1.0, r);
        double rr;
        if ((rc = dev_dot(h, n, r, r, &rr))) break;
        moni[k] = std::sqrt(rr);                                                         // :328
        if (it >= STAG - 1) {                                                            // :333-341
            double mx = *std::max_element(moni, moni + STAG), mn = *std::min_element(moni, moni + STAG);
            if (moni[k] <= tol || (moni[k] <= 1.0E2 * tol && (mx - mn) < 0.1 * ((mx + mn) / 2.0))) break;
        }
        it++;
    }
    if (h->profile) h->prof_collect();
    if (rc) { cleanup(); return 1; }
    if (iters) *iters = it;
    if (resid) *resid = moni[it % STAG];
    if (tol_abs) *tol_abs = tol;
    rc = to_host(h, Lf, xd, x);
    cleanup();
    return rc;
}

// MGPIS::BiCGSTAB_SOLV, MGPIS.h:350-432
static int bicgstab_hostloop(ddpca_mg *h, int prec, const double *b, double *x, double rel_tol, long maxit, long *iters,
                             double *resid, double *tol_abs)
{
    int Lf = h->nlev - 1;
    Level &L = h->lev[Lf];
    int n = L.n;
    if (prec == 0 && !L.dinv) {
        CU(cudaMalloc(&L.dinv, sizeof(double) * n));
        if (L.v2) k_extract_diag_inv2<<<cdiv(L.ng, 256), 256, 0, h->stream>>>(L.view2(), L.dinv);
        else k_extract_diag_inv<<<cdiv(L.ng, 256), 256, 0, h->stream>>>(L.view(), L.dinv);
    }
    // work vectors: r, rhat, p, v, s, t, phat, shat, x
    std::vector<double *> w(4, nullptr);
    auto cleanup = [&]() { for (auto q : w) cudaFree(q); };
    for (auto &p : w) if (alloc_vec(&p, n)) { cleanup(); return 1; }
    double *r = h->cg_r, *rhat = w[0], *p = h->cg_p, *v = h->cg_q, *s = w[1], *t = w[2], *phat = h->cg_z, *shat = w[3], *xd = h->cg_x;
    if (to_dev(h, Lf, b, r)) { cleanup(); return 1; }
    CUX(cudaMemsetAsync(xd, 0, sizeof(double) * n, h->stream));                                   // :361
    CUX(cudaMemcpyAsync(rhat, r, sizeof(double) * n, cudaMemcpyDeviceToDevice, h->stream));        // :378
    double bb, rr;
    if (dev_dot(h, n, r, r, &bb)) { cleanup(); return 1; }
    const double tol = rel_tol * std::sqrt(bb);                                                  // :363
    rr = bb;
    double rho_old = 1.0, rho_new = 1.0, alph = 1.0, omeg = 1.0;
    long it = 0;
    while (it < maxit && std::sqrt(rr) > tol) {                                                  // :382
        if (dev_dot(h, n, rhat, r, &rho_new)) { cleanup(); return 1; }                           // :383
        if (std::fabs(rho_new) == 0.0) break;                                                    // :384-387
        if (it == 0) {
            CUX(cudaMemcpyAsync(p, r, sizeof(double) * n, cudaMemcpyDeviceToDevice, h->stream));   // :389
        } else {
            double beta = (rho_new / rho_old) * (alph / omeg);                                   // :392-393
            dev_axpby(h, n, -omeg, v, 1.0, p);                                                   // p - omeg v
            dev_axpby(h, n, 1.0, r, beta, p);                                                    // :394
        }
        precondition(h, prec, p, phat, nullptr);                                                 // :396-402
        launch_level_spmv(h, L, Lf, phat, v, nullptr, nullptr, nullptr);                         // :403
        double rv;
        if (dev_dot(h, n, rhat, v, &rv)) { cleanup(); return 1; }
        alph = rho_new / rv;                                                                     // :404
        CUX(cudaMemcpyAsync(s, r, sizeof(double) * n, cudaMemcpyDeviceToDevice, h->stream));
        dev_axpby(h, n, -alph, v, 1.0, s);                                                       // :405
        double ss;
        if (dev_dot(h, n, s, s, &ss)) { cleanup(); return 1; }
        if (std::sqrt(ss) <= 0.0) {                                                              // :406-409
            dev_axpby(h, n, alph, phat, 1.0, xd);
            break;
        }
        precondition(h, prec, s, shat, nullptr);                                                 // :410-416
        launch_level_spmv(h, L, Lf, shat, t, nullptr, nullptr, nullptr);                         // :417
        double ts, tt;
        if (dev_dot(h, n, t, s, &ts) || dev_dot(h, n, t, t, &tt)) { cleanup(); return 1; }
        omeg = ts / tt;                                                                          // :418
        dev_axpby(h, n, alph, phat, 1.0, xd);
        dev_axpby(h, n, omeg, shat, 1.0, xd);                                                    // :419
        CUX(cudaMemcpyAsync(r, s, sizeof(double) * n, cudaMemcpyDeviceToDevice, h->stream));
        dev_axpby(h, n, -omeg, t, 1.0, r);                                                       // :420
        if (dev_dot(h, n, r, r, &rr)) { cleanup(); return 1; }
        rho_old = rho_new;
        it++;
    }
    if (h->profile) h->prof_collect();
    if (iters) *iters = it;
    if (resid) *resid = std::sqrt(rr);
    if (tol_abs) *tol_abs = tol;
    int rc = to_host(h, Lf, xd, x);
    cleanup();
    return rc;
}

// ---- the same three drivers with every scalar on the device (kernels.cuh, KrylovState) -----------------------
// The host enqueues whole iterations ahead and polls the done flag with a lag of kDepth iterations; launches
// issued past the stop leave at their first instruction.  No reduction result crosses to the host inside the loop.
static int krylov_run(ddpca_mg *h, const std::function<int(long)> &enqueue_iter)
{
    const int depth = h->profile ? 0 : kDepth;
    cudaEvent_t evs[kDepth + 1];
    for (int k = 0; k <= kDepth; k++) CU(cudaEventCreateWithFlags(&evs[k], cudaEventDisableTiming));
    CU(cudaMemcpyAsync(&h->fl_host[0], h->fl, sizeof(BatchFlags), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaEventRecord(evs[0], h->stream));
    CU(cudaEventSynchronize(evs[0]));
    bool finished = h->fl_host[0].done_all != 0;
    long issued = 0;
    while (!finished) {
        const int slot = (int)(issued % (kDepth + 1));
        if (enqueue_iter(issued)) return 1;
        CU(cudaMemcpyAsync(&h->fl_host[slot], h->fl, sizeof(BatchFlags), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaEventRecord(evs[slot], h->stream));
        issued++;
        if (issued >= depth) {
            const int old = (int)((issued - std::max(depth, 1)) % (kDepth + 1));
            CU(cudaEventSynchronize(evs[old]));
            if (h->fl_host[old].done_all) finished = true;
        }
        if (h->profile && (issued % 8) == 0) h->prof_collect();
    }
    for (int k = 0; k <= kDepth; k++) cudaEventDestroy(evs[k]);
    return 0;
}
static int krylov_prepare(ddpca_mg *h, int prec)
{
    if (h->nsub != 1) return fail("MULT_SOLV / GMRES_SOLV / BiCGSTAB_SOLV take one hierarchy (not a batch)");
    if (prec == 1 && !h->Binv) return fail("this hierarchy was built without a level-0 direct solver: Jacobi preconditioning only");
    Level &L = h->lev[h->nlev - 1];
    if (prec == 0 && !L.dinv) {
        CU(cudaMalloc(&L.dinv, sizeof(double) * L.n));
        if (L.v2) k_extract_diag_inv2<<<cdiv(L.ng, 256), 256, 0, h->stream>>>(L.view2(), L.dinv);
        else k_extract_diag_inv<<<cdiv(L.ng, 256), 256, 0, h->stream>>>(L.view(), L.dinv);
    }
    if (!h->ks) CU(cudaMalloc(&h->ks, sizeof(KrylovState)));
    return 0;
}
static int krylov_finish(ddpca_mg *h, KrylovState *hs)
{
    CU(cudaMemcpyAsync(hs, h->ks, sizeof(KrylovState), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (!h->launch_err.empty()) { std::string m = h->launch_err; h->launch_err.clear(); cudaGetLastError(); return fail(m); }
    if (h->profile) h->prof_collect();
    CU(cudaGetLastError());
    return 0;
}
#define KSEGDOT(a, b, part, dn) KL(h, DDPCA_K_VECTOR, Lf, 16.0 * n, (k_seg_dot<<<ns, 256, 0, h->stream>>>(h->seg_d, a, b, part, dn)))

// MGPIS::MULT_SOLV, MGPIS.h:130-160
static int mult_solv_device(ddpca_mg *h, const double *b, double *x, long *iters, double *resid)
{
    if (krylov_prepare(h, 1)) return 1;
    const int Lf = h->nlev - 1;
    Level &L = h->lev[Lf];
    const int n = L.n, ns = h->nseg, gv = vec_grid(h, n);
    const int *done = h->done_flag();
    double *bd = h->cg_r, *xd = h->cg_x, *rd = h->cg_q, *ax = h->cg_p;
    if (to_dev(h, Lf, b, bd)) return 1;
    CU(cudaMemsetAsync(xd, 0, sizeof(double) * n, h->stream));                                           // :133
    KSEGDOT(bd, bd, h->partial[0], nullptr);
    KL(h, DDPCA_K_VECTOR, Lf, 0.0, (k_kry_init<<<1, 32, 0, h->stream>>>(h->ks, h->st, h->fl, h->partial[0], ns, 0.0, 10000LL, 0)));   // :134
    if (krylov_run(h, [&](long) -> int {
            vcycle_dev(h, Lf, bd, xd, false, done);                                                      // :143
            launch_level_spmv(h, L, Lf, xd, ax, nullptr, nullptr, done);
            KL(h, DDPCA_K_VECTOR, Lf, 24.0 * n, (k_kry_sub<<<gv, 256, 0, h->stream>>>(n, bd, ax, rd, done)));    // :144
            KSEGDOT(rd, rd, h->partial[0], done);
            KL(h, DDPCA_K_VECTOR, Lf, 0.0, (k_ms_next<<<1, 32, 0, h->stream>>>(h->ks, h->st, h->fl, h->partial[0], ns)));   // :146-154
            return 0;
        })) return 1;
    KrylovState hs;
    if (krylov_finish(h, &hs)) return 1;
    if (iters) *iters = (long)hs.it;
    if (resid) *resid = hs.moni[hs.it % 5];
    return to_host(h, Lf, xd, x);
}

// MGPIS::GMRES_SOLV, MGPIS.h:227-348: left-preconditioned restarted GMRES(10); the 11x10 Hessenberg algebra runs in
// a one-warp kernel (k_gm_hess), the Arnoldi coefficients never leave the device
static int gmres_device(ddpca_mg *h, int prec, const double *b, double *x, long *iters, double *resid, double *tol_abs)
{
    if (krylov_prepare(h, prec)) return 1;
    const int Lf = h->nlev - 1;
    Level &L = h->lev[Lf];
    const int n = L.n, ns = h->nseg, gv = vec_grid(h, n);
    const int *done = h->done_flag();
    // Arnoldi vectors with stride n+1: each keeps the always-zero slot the padded gathers rely on (alloc_vec)
    const size_t ldv = (size_t)n + 1;
    double *V = nullptr, *x0 = nullptr, *bd = nullptr, *hp = nullptr;
    auto cleanup = [&]() { cudaFree(V); cudaFree(x0); cudaFree(bd); cudaFree(hp); };
    if (alloc_vec(&V, (long)(ldv * (kGmStag + 1)) - 1) || alloc_vec(&x0, n) || alloc_vec(&bd, n) || alloc_vec(&hp, (long)kGmStag * ns)) { cleanup(); return 1; }
    double *r = h->cg_r, *w0 = h->cg_q, *w = h->cg_z, *xd = h->cg_x;
    if (to_dev(h, Lf, b, bd)) { cleanup(); return 1; }
    CUX(cudaMemsetAsync(xd, 0, sizeof(double) * n, h->stream));                                           // :248
    KSEGDOT(bd, bd, h->partial[0], nullptr);
    KL(h, DDPCA_K_VECTOR, Lf, 0.0, (k_kry_init<<<1, 32, 0, h->stream>>>(h->ks, h->st, h->fl, h->partial[0], ns, 1.0E-12, (long long)n, 0)));   // :249-250
    int rc = krylov_run(h, [&](long it) -> int {
        const int k = (int)(it % kGmStag);
        if (k == 0) {                                                                                     // restart, :263-276
            if (cudaMemcpyAsync(x0, xd, sizeof(double) * n, cudaMemcpyDeviceToDevice, h->stream) != cudaSuccess) return fail("GMRES restart: device copy failed");
            launch_level_spmv(h, L, Lf, x0, w0, nullptr, nullptr, done);
            KL(h, DDPCA_K_VECTOR, Lf, 24.0 * n, (k_kry_sub<<<gv, 256, 0, h->stream>>>(n, bd, w0, r, done)));
            precondition(h, prec, r, w, done);
            KSEGDOT(w, w, h->partial[0], done);
            KL(h, DDPCA_K_VECTOR, Lf, 0.0, (k_gm_restart<<<1, 32, 0, h->stream>>>(h->ks, h->fl, h->partial[0], ns)));
            KL(h, DDPCA_K_VECTOR, Lf, 16.0 * n, (k_gm_scale<<<gv, 256, 0, h->stream>>>(n, h->ks, 0, 0, w, V, done)));
        }
        launch_level_spmv(h, L, Lf, V + (size_t)k * ldv, w0, nullptr, nullptr, done);                     // :278
        precondition(h, prec, w0, w, done);                                                               // :279-285
        KL(h, DDPCA_K_VECTOR, Lf, 8.0 * n * (k + 2), (k_gm_dots<<<ns, 256, 0, h->stream>>>(h->seg_d, V, ldv, k, w, hp, ns, done)));   // :286
        KL(h, DDPCA_K_VECTOR, Lf, 0.0, (k_gm_hcol<<<1, 32, 0, h->stream>>>(h->ks, h->fl, hp, ns, k)));
        KL(h, DDPCA_K_VECTOR, Lf, 8.0 * n * (k + 3), (k_gm_orth<<<gv, 256, 0, h->stream>>>(n, h->ks, V, ldv, k, w, done)));            // :287
        KSEGDOT(w, w, h->partial[0], done);                                                               // :288
        KL(h, DDPCA_K_VECTOR, Lf, 0.0, (k_gm_hess<<<1, 32, 0, h->stream>>>(h->ks, h->fl, h->partial[0], ns, k)));                       // :289-324
        KL(h, DDPCA_K_VECTOR, Lf, 16.0 * n, (k_gm_scale<<<gv, 256, 0, h->stream>>>(n, h->ks, 1, k, w, V + (size_t)(k + 1) * ldv, done)));   // :294-296
        KL(h, DDPCA_K_VECTOR, Lf, 8.0 * n * (k + 3), (k_gm_update_x<<<gv, 256, 0, h->stream>>>(n, h->ks, V, ldv, k, x0, xd, done)));     // :325
        launch_level_spmv(h, L, Lf, xd, w0, nullptr, nullptr, done);                                      // :326
        KL(h, DDPCA_K_VECTOR, Lf, 24.0 * n, (k_kry_sub<<<gv, 256, 0, h->stream>>>(n, bd, w0, r, done)));
        KSEGDOT(r, r, h->partial[0], done);
        KL(h, DDPCA_K_VECTOR, Lf, 0.0, (k_gm_next<<<1, 32, 0, h->stream>>>(h->ks, h->st, h->fl, h->partial[0], ns, k)));               // :328-342
        return 0;
    });
    KrylovState hs;
    if (rc || krylov_finish(h, &hs)) { cleanup(); return 1; }
    if (iters) *iters = (long)hs.it;
    if (resid) *resid = hs.moni[hs.it % kGmStag];
    if (tol_abs) *tol_abs = hs.tol;
    rc = to_host(h, Lf, xd, x);
    cleanup();
    return rc;
}

// MGPIS::BiCGSTAB_SOLV, MGPIS.h:350-432
static int bicgstab_device(ddpca_mg *h, int prec, const double *b, double *x, double rel_tol, long maxit, long *iters,
                           double *resid, double *tol_abs)
{
    if (krylov_prepare(h, prec)) return 1;
    const int Lf = h->nlev - 1;
    Level &L = h->lev[Lf];
    const int n = L.n, ns = h->nseg, gv = vec_grid(h, n);
    const int *done = h->done_flag();
    std::vector<double *> w(4, nullptr);
    auto cleanup = [&]() { for (auto q : w) cudaFree(q); };
    for (auto &p : w) if (alloc_vec(&p, n)) { cleanup(); return 1; }
    double *r = h->cg_r, *rhat = w[0], *p = h->cg_p, *v = h->cg_q, *s = w[1], *t = w[2], *phat = h->cg_z, *shat = w[3], *xd = h->cg_x;
    if (to_dev(h, Lf, b, r)) { cleanup(); return 1; }
    CUX(cudaMemsetAsync(xd, 0, sizeof(double) * n, h->stream));                                           // :361
    CUX(cudaMemcpyAsync(rhat, r, sizeof(double) * n, cudaMemcpyDeviceToDevice, h->stream));                // :378
    KSEGDOT(r, r, h->partial[0], nullptr);
    KL(h, DDPCA_K_VECTOR, Lf, 0.0, (k_kry_init<<<1, 32, 0, h->stream>>>(h->ks, h->st, h->fl, h->partial[0], ns, rel_tol, (long long)maxit, 1)));   // :362-363,382
    int rc = krylov_run(h, [&](long) -> int {
        KSEGDOT(rhat, r, h->partial[0], done);                                                            // :383
        KL(h, DDPCA_K_VECTOR, Lf, 0.0, (k_bi_rho<<<1, 32, 0, h->stream>>>(h->ks, h->st, h->fl, h->partial[0], ns)));                   // :384-393
        KL(h, DDPCA_K_VECTOR, Lf, 32.0 * n, (k_bi_update_p<<<gv, 256, 0, h->stream>>>(n, h->ks, r, v, p, done)));                     // :388-395
        precondition(h, prec, p, phat, done);                                                             // :396-402
        launch_level_spmv(h, L, Lf, phat, v, nullptr, nullptr, done);                                     // :403
        KSEGDOT(rhat, v, h->partial[0], done);
        KL(h, DDPCA_K_VECTOR, Lf, 0.0, (k_bi_alpha<<<1, 32, 0, h->stream>>>(h->ks, h->fl, h->partial[0], ns)));                        // :404
        KL(h, DDPCA_K_VECTOR, Lf, 24.0 * n, (k_bi_s<<<gv, 256, 0, h->stream>>>(n, h->ks, r, v, s, done)));                            // :405
        KSEGDOT(s, s, h->partial[0], done);
        KL(h, DDPCA_K_VECTOR, Lf, 0.0, (k_bi_scheck<<<1, 32, 0, h->stream>>>(h->ks, h->fl, h->partial[0], ns)));                       // :406-409
        precondition(h, prec, s, shat, done);                                                             // :410-416
        launch_level_spmv(h, L, Lf, shat, t, nullptr, nullptr, done);                                     // :417
        KSEGDOT(t, s, h->partial[0], done);
        KSEGDOT(t, t, h->partial[1], done);
        KL(h, DDPCA_K_VECTOR, Lf, 0.0, (k_bi_omega<<<1, 64, 0, h->stream>>>(h->ks, h->fl, h->partial[0], h->partial[1], ns)));         // :418
        KL(h, DDPCA_K_VECTOR, Lf, 64.0 * n, (k_bi_update_xr<<<ns, 256, 0, h->stream>>>(h->seg_d, h->ks, phat, shat, s, t, xd, r, h->partial[2], done)));   // :419-420
        KL(h, DDPCA_K_VECTOR, Lf, 0.0, (k_bi_next<<<1, 32, 0, h->stream>>>(h->ks, h->st, h->fl, h->partial[2], ns)));                  // :382
        return 0;
    });
    KrylovState hs;
    if (rc || krylov_finish(h, &hs)) { cleanup(); return 1; }
    if (iters) *iters = (long)hs.it;
    if (resid) *resid = std::sqrt(hs.rr);
    if (tol_abs) *tol_abs = hs.tol;
    rc = to_host(h, Lf, xd, x);
    cleanup();
    return rc;
}
#undef KSEGDOT

static bool krylov_hostloop() { return std::getenv("DDPCA_KRYLOV_HOSTLOOP") != nullptr; }

int ddpca_mg_mult_solv(ddpca_mg *h, const double *b, double *x, long *iters, double *resid)
{
    if (!h || !b || !x) return fail("ddpca_mg_mult_solv: bad argument");
    CU(cudaSetDevice(h->device));
    return krylov_hostloop() ? mult_solv_hostloop(h, b, x, iters, resid) : mult_solv_device(h, b, x, iters, resid);
}
int ddpca_mg_gmres(ddpca_mg *h, int prec, const double *b, double *x, long *iters, double *resid, double *tol_abs)
{
    if (!h || !b || !x || (prec != 0 && prec != 1)) return fail("ddpca_mg_gmres: bad argument");
    CU(cudaSetDevice(h->device));
    return krylov_hostloop() ? gmres_hostloop(h, prec, b, x, iters, resid, tol_abs) : gmres_device(h, prec, b, x, iters, resid, tol_abs);
}
int ddpca_mg_bicgstab(ddpca_mg *h, int prec, const double *b, double *x, double rel_tol, long maxit, long *iters,
                      double *resid, double *tol_abs)
{
    if (!h || !b || !x || (prec != 0 && prec != 1)) return fail("ddpca_mg_bicgstab: bad argument");
    CU(cudaSetDevice(h->device));
    return krylov_hostloop() ? bicgstab_hostloop(h, prec, b, x, rel_tol, maxit, iters, resid, tol_abs)
                             : bicgstab_device(h, prec, b, x, rel_tol, maxit, iters, resid, tol_abs);
}

int ddpca_ldlt_create(int device, int n, const int *perm, const int *L_rowptr, const int *L_colidx, const double *L_val,
                      const double *D, ddpca_ldlt **out)
{
    if (!out || n < 1 || !perm || !L_rowptr || !L_colidx || !L_val || !D) return fail("ddpca_ldlt_create: bad argument");
    return ldlt_build(device, n, perm, L_rowptr, L_colidx, L_val, D, out);
}
int ddpca_ldlt_create_dense(int device, int n, const int *rowptr, const int *colidx, const double *val, ddpca_ldlt **out)
{
    if (!out || n < 1 || n > 32768 || !rowptr || !colidx || !val) return fail("ddpca_ldlt_create_dense: bad argument (n must be <= 32768)");
    int ndev = ddpca_device_count();
    if (ndev == 0) return fail("no CUDA device: libddpca_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail("device index out of range");
    CU(cudaSetDevice(device));
    StageTimer tm("dense solver");
    CsrHost A;
    A.rows = A.cols = n;
    A.rp.assign(rowptr, rowptr + n + 1);
    A.ci.assign(colidx, colidx + rowptr[n]);
    A.v.assign(val, val + rowptr[n]);
    ddpca_ldlt *s = new ddpca_ldlt();
    s->device = device;
    s->n = n;
    s->nnzL = rowptr[n];
    cudaDeviceGetAttribute(&s->sms, cudaDevAttrMultiProcessorCount, device);
    if (cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete s; return fail("stream creation failed"); }
    s->stream = s->own_stream;
    // operator upload, inversion and release of the upload are enqueued on the device's set-up stream; this solver's own
    // stream is ordered behind them by an event, the host goes on to the next operator
    cudaStream_t sst = setup_stream(device);
    if (!sst) { ldlt_free(s); return fail("stream creation failed"); }
    DevCsr dA;
    dA.rows = dA.cols = n; dA.nnz = A.nnz();
    const size_t nnz = (size_t)A.nnz();
    bool ok = cudaMallocAsync((void **)&dA.rp, sizeof(int) * (n + 1), sst) == cudaSuccess &&
              cudaMallocAsync((void **)&dA.ci, sizeof(int) * std::max<size_t>(1, nnz), sst) == cudaSuccess &&
              cudaMallocAsync((void **)&dA.v, sizeof(double) * std::max<size_t>(1, nnz), sst) == cudaSuccess &&
              cudaMemcpyAsync(dA.rp, A.rp.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice, sst) == cudaSuccess &&
              (nnz == 0 || (cudaMemcpyAsync(dA.ci, A.ci.data(), sizeof(int) * nnz, cudaMemcpyHostToDevice, sst) == cudaSuccess &&
                            cudaMemcpyAsync(dA.v, A.v.data(), sizeof(double) * nnz, cudaMemcpyHostToDevice, sst) == cudaSuccess));
    auto drop_upload = [&]() { if (dA.rp) cudaFreeAsync(dA.rp, sst); if (dA.ci) cudaFreeAsync(dA.ci, sst); if (dA.v) cudaFreeAsync(dA.v, sst); dA = DevCsr(); };
    if (!ok) { drop_upload(); ldlt_free(s); cudaGetLastError(); return fail("operator upload failed (out of device memory?)"); }
    tm.lap("operator upload (enqueued)");
    if (cudaMalloc(&s->Binv, sizeof(double) * (size_t)n * n) != cudaSuccess) { drop_upload(); ldlt_free(s); cudaGetLastError(); return fail("out of device memory"); }
    tm.lap("allocation of the inverse");
    if (dense_spd_inverse_block(sst, dA, 0, n, s->Binv, /*wait=*/false)) { drop_upload(); ldlt_free(s); return 1; }
    drop_upload();
    {
        cudaEvent_t ev;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(ev, sst) != cudaSuccess ||
            cudaStreamWaitEvent(s->stream, ev, 0) != cudaSuccess) { ldlt_free(s); cudaGetLastError(); return fail("event set-up failed"); }
        cudaEventDestroy(ev);   // released by the runtime once the wait has been satisfied
    }
    tm.lap("inversion (enqueued)");
    // work vectors for the host-pointer entry point
    if (cudaMalloc(&s->lo.r, sizeof(double) * n) != cudaSuccess || cudaMalloc(&s->up.r, sizeof(double) * n) != cudaSuccess) { ldlt_free(s); return fail("out of device memory"); }
    tm.lap("release + work vectors");
    *out = s;
    return 0;
}
int ddpca_ldlt_destroy(ddpca_ldlt *s) { ldlt_free(s); return 0; }
int ddpca_ldlt_solve_dev(ddpca_ldlt *s, const double *b_dev, double *x_dev)
{
    if (!s || !b_dev || !x_dev) return fail("ddpca_ldlt_solve_dev: bad argument");
    CU(cudaSetDevice(s->device));
    ldlt_solve_on(s, s, b_dev, x_dev, nullptr);
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    return 0;
}
int ddpca_ldlt_solve(ddpca_ldlt *s, const double *b, double *x)
{
    if (!s || !b || !x) return fail("ddpca_ldlt_solve: bad argument");
    CU(cudaSetDevice(s->device));
    // lo.r / up.r are free work vectors of length n
    CU(cudaMemcpyAsync(s->lo.r, b, sizeof(double) * s->n, cudaMemcpyHostToDevice, s->stream));
    ldlt_solve_on(s, s, s->lo.r, s->up.r, nullptr);
    CU(cudaMemcpyAsync(x, s->up.r, sizeof(double) * s->n, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    return 0;
}
int ddpca_ldlt_info(const ddpca_ldlt *s, int *n, long *nnzL, int *stages_fwd, int *stages_bwd)
{
    if (!s) return fail("null handle");
    if (n) *n = s->n;
    if (nnzL) *nnzL = s->nnzL;
    if (stages_fwd) *stages_fwd = s->lo.plan.nstages();
    if (stages_bwd) *stages_bwd = s->up.plan.nstages();
    return 0;
}

int ddpca_mg_level_info(const ddpca_mg *h, int level, long *n, long *nnz, int *ngroups, int *nstages)
{
    if (!h || level < 0 || level >= h->nlev) return fail("ddpca_mg_level_info: bad argument");
    const Level &L = h->lev[level];
    if (n) *n = L.n;
    if (nnz) *nnz = L.nnz;
    if (ngroups) *ngroups = L.plan.ngroups();
    if (nstages) *nstages = L.plan.nstages();
    return 0;
}
long ddpca_mg_launch_count(ddpca_mg *h, int reset)
{
    if (!h) return -1;
    long v = h->launches;
    if (reset) h->launches = 0;
    return v;
}
int ddpca_mg_profile(ddpca_mg *h, int enable)
{
    if (!h) return fail("null handle");
    h->profile = enable != 0;
    if (enable) h->prof_reset();
    return 0;
}
int ddpca_mg_profile_get(ddpca_mg *h, int kclass, int level, double *ms, long *launches, double *bytes)
{
    if (!h || kclass < 0 || kclass >= DDPCA_K_COUNT || level < 0 || level >= 16) return fail("ddpca_mg_profile_get: bad argument");
    if (ms) *ms = h->prof_ms[kclass][level];
    if (launches) *launches = h->prof_n[kclass][level];
    if (bytes) *bytes = h->prof_bytes[kclass][level];
    return 0;
}
int ddpca_mg_last_timing(ddpca_mg *h, double *solve_ms, double *h2d_ms, double *d2h_ms)
{
    if (!h) return fail("null handle");
    if (solve_ms) *solve_ms = h->t_solve;
    if (h2d_ms) *h2d_ms = h->t_h2d;
    if (d2h_ms) *d2h_ms = h->t_d2h;
    return 0;
}

}  // extern "C"

#include "admm.inl"
#include "group.inl"
