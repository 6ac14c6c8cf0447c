// admm.inl -- device-resident ADMM iteration of MCONTACT::CONTACT_ANALYSIS (MCONTACT.h:2493-2723).
// Included at the end of mg.cu (one translation unit: the kernels of kernels.cuh are shared).
//
// Design (round 2): the loop body is written for MANY subdomains per GPU.  The reference iterates
// `#pragma omp parallel for` over bodies and interfaces (MCONTACT.h:2511,2629,2689), a dozen small sparse
// products each.  Here all local bodies and interface sides are CONCATENATED:
//   * state vectors live in one array   STATE = [ inteAuxi | inteLagr | resuDisp | inpoGamm ]
//     (all local sides / bodies / interfaces one after the other), previous iterates in STATE_PREV;
//   * every "sum over sides of operator x vector" of the loop body is ONE sparse product with a stacked
//     operator assembled once on the host from the reference's own matrices (a13):
//        OPA  addiForc   = [ systTran_pena | -systTran ] STATE                       (:2515-2523)
//        F/FT ADDITIONAL_FORCE / OUTP_SUB1 of all bodies, block diagonal              (:2524,2533)
//        OPG  globForc   = [ -globTran_pena | globTran | globTran_D ] STATE           (:2541-2549)
//        OPT  2 gamma+gap= [ +-inpoLagr | +-pemaInpo_r ] STATE                        (:2632-2635)
//        OPF1 / OPF2     right-hand sides of the auxiliary / multiplier updates       (:2671-2675,:2691-2694)
//   * the body solves are ONE batched MG-PCG per group of bodies with equal level count (mg.cu: the bodies'
//     hierarchies as one block-diagonal hierarchy, per-body CG scalars);
//   * the interface mass solves are one block-diagonal dense product (dense inverses built at set-up), sides
//     beyond the dense limit fall back to the staged sparse LDL^T solve.
// An iteration is ~25 launches + the solve graphs whatever the number of subdomains; per-body and per-side
// results do not depend on what else is in the batch (row-wise products, per-sub fixed-order reductions), so
// runs on 1, 2, 4, 8 GPUs produce the same bits.
// Multi-GPU (one process per GPU): a rank holds its bodies and their sides; per iteration it exchanges the
// signed side traces of cross-rank interfaces pairwise with the owning peer, and all-reduces the coarse
// right-hand side and the MONITOR sums.

namespace {

struct HostSide {
    CsrHost op[DDPCA_OP_COUNT];
    bool has[DDPCA_OP_COUNT] = {};
    ddpca_ldlt *mass = nullptr, *mass_pena = nullptr;
    int nc = 0;          // d * n_c
    bool iterative = false;   // mass solves by Jacobi-PCG on the device instead of a factor (MCONTACT.h:2678-2683,2698-2703)
    bool local = true;   // side lives with its body
    int coff = -1;       // offset in the AUX / LAGR parts of STATE
};
struct HostIface {
    int body[2] = {-1, -1};
    double fric = 0.0;
    int nip = 0, d = 1;
    std::vector<double> gap;
    HostSide side[2];
    bool set = false;
    bool cross = false;   // the two sides live on different ranks
    int peer = -1;        // the other rank of a cross interface
    int goff = -1;        // offset in the GAMMA part of STATE (and in T, GAP, STAT); -1: no local side
};
struct HostBody {
    bool set = false, local = true;
    int nlev = 0;
    std::vector<int> n;
    std::vector<const int *> rp, ci, prp, pci;   // caller's arrays: valid until ddpca_admm_finalize
    std::vector<const double *> v, pv;
    int nfull = 0, nred = 0;
    CsrHost F, accuProl, globTran_D_1;
    std::vector<double> consForc, dispCons;
    int foff = -1, roff = -1;   // offsets in the DISP part of STATE / in RHS and U
    int batch = -1, sub = -1;
};
struct Batch {
    ddpca_mg *mg = nullptr;
    std::vector<int> bodies;
    int roff = 0, nred = 0;
    cudaEvent_t ev_done = nullptr;
};
struct IfaceMeta { int ip0, nip, d, goff; double fric; };   // device table of the projection kernel

}  // namespace

// gamma = 0.5 (t - gapTerm) + contact projection for all active interfaces at once (MCONTACT.h:2636-2668);
// one thread per integration point.  fric < 0 tied, = 0 frictionless (one component), > 0 Coulomb.
__global__ void __launch_bounds__(256) k_gamma_project_all(int nif, const IfaceMeta *__restrict__ meta, const double *__restrict__ t,
                                                           const double *__restrict__ gap, double *__restrict__ gamma, int *__restrict__ stat)
{
    const int ip = blockIdx.x * blockDim.x + threadIdx.x;
    if (nif == 0 || ip >= meta[nif - 1].ip0 + meta[nif - 1].nip) return;
    int lo = 0, hi = nif;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (meta[mid].ip0 <= ip) lo = mid; else hi = mid; }
    const IfaceMeta m = meta[lo];
    const int k = ip - m.ip0;
    if (m.d == 1) {
        const int i = m.goff + k;
        double g = 0.5 * (t[i] - gap[i]);
        if (m.fric >= 0.0) g = fmax(0.0, g);
        gamma[i] = g;
        stat[i] = 0;
        return;
    }
    const int i = m.goff + 3 * k;
    double gn = 0.5 * (t[i] - gap[i]);
    double g1 = 0.5 * (t[i + 1] - gap[i + 1]);
    double g2 = 0.5 * (t[i + 2] - gap[i + 2]);
    int st = 0;
    if (m.fric >= 0.0) gn = fmax(0.0, gn);
    if (m.fric > 0.0) {
        if (gn > 0.0) {
            const double slid = m.fric * gn;
            const double nrm = sqrt(g1 * g1 + g2 * g2);
            if (nrm >= slid) { const double f = slid / nrm; g1 = f * g1; g2 = f * g2; st = 1; }
            else st = 2;
        } else { g1 = 0.0; g2 = 0.0; st = 0; }
    }
    gamma[i] = gn; gamma[i + 1] = g1; gamma[i + 2] = g2;
    stat[i] = 0; stat[i + 1] = st; stat[i + 2] = 0;
}
// MONITOR sums (MCONTACT.h:2737-2833): per chunk of one state vector, sum (cur-prev)^2 and sum cur^2
__global__ void __launch_bounds__(256) k_moni_seg(const SegChunk *__restrict__ ch, const double *__restrict__ cur, const double *__restrict__ prev,
                                                  double *__restrict__ part)
{
    const SegChunk c = ch[blockIdx.x];
    double a = 0.0, b = 0.0;
    for (int i = c.row0 + threadIdx.x; i < c.row0 + c.nrows; i += 256) {
        const double v = cur[i], dlt = v - prev[i];
        a += dlt * dlt;
        b += v * v;
    }
    __shared__ double sa[8], sb[8];
    a = warp_sum(a); b = warp_sum(b);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sa[w] = a; sb[w] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
        for (int k = 0; k < 8; k++) { ta += sa[k]; tb += sb[k]; }
        part[2 * blockIdx.x] = ta;
        part[2 * blockIdx.x + 1] = tb;
    }
}
// out[2s], out[2s+1] = sums of the chunks of slot s, fixed order (slots without local data: 0)
__global__ void k_moni_slots(int nslots, const int *__restrict__ slot_chunk, const double *__restrict__ part, double *__restrict__ out)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    double ta = 0.0, tb = 0.0;
    for (int c = slot_chunk[s]; c < slot_chunk[s + 1]; c++) { ta += part[2 * c]; tb += part[2 * c + 1]; }
    out[2 * s] = ta;
    out[2 * s + 1] = tb;
}

struct ddpca_admm : Engine {
    int nb = 0, ni = 0, muscSett = 0;
    int smoother_mode = -1;   // -1: DDPCA_SMOOTH_MC unless the environment says DDPCA_SMOOTHER=lex
    std::vector<HostBody> body;
    std::vector<HostIface> iface;
    std::vector<Batch> batch;
    // sizes of the concatenated vectors
    int NC = 0, NF = 0, NG = 0, NR = 0, NGloc = 0;   // NGloc: GAMMA entries of interfaces with both sides here
    int offAux() const { return 0; }
    int offLagr() const { return NC; }
    int offDisp() const { return 2 * NC; }
    int offGamma() const { return 2 * NC + NF; }
    double *state = nullptr, *state_prev = nullptr;
    double *addi = nullptr, *rhs = nullptr, *u = nullptr, *force = nullptr, *tmp = nullptr, *consForc = nullptr, *dispCons = nullptr;
    double *t = nullptr, *gap = nullptr;
    int *stat = nullptr;
    DevCsr OPA, Fall, FTall, OPG, OPG1, ACCU, OPT, OPF1, OPF2;
    IfaceMeta *meta_d = nullptr;
    int nact = 0, nip_tot = 0;
    // interface mass solves: dense blocks in one launch, the rest one by one
    int nls = 0;
    int *side_off_d = nullptr;
    double **mass_ptr_d = nullptr, **pena_ptr_d = nullptr;
    std::vector<std::pair<int, int>> sparse_mass, sparse_pena;   // (ts, tv) of sides whose solver is not dense
    // sides without a factor (the reference: rows >= DIRE_MAXI, Eigen CG with the diagonal preconditioner): one batched
    // Jacobi-PCG over all of them per update; they are numbered last, so their vectors are one contiguous range
    ddpca_mg *pcg_mass = nullptr, *pcg_pena = nullptr;
    int it_off = 0, it_rows = 0;
    ddpca_mg *coar1_mg = nullptr;   // interface-eliminated problem by MG-PCG (globCoup_1 beyond DIRE_MAXI rows, MCONTACT.h:2593-2595)
    // coarse problems
    int nglob = 0;
    std::vector<long> baseReco;
    ddpca_ldlt *coar = nullptr;
    ddpca_mg *coar_mg = nullptr;   // macroscopic problem solved by MG-PCG instead (globCoup beyond DIRE_MAXI rows)
    long macro_cg_iters = 0;
    double *globForc = nullptr, *globSolu = nullptr;
    int nglob1 = 0;
    ddpca_ldlt *coar1 = nullptr;
    double *globForc1_const = nullptr, *globSolu1 = nullptr;
    // MONITOR
    int nslots = 0, nmchunks = 0;
    SegChunk *moni_chunks_d = nullptr;
    int *slot_chunk_d = nullptr;
    double *moni_part = nullptr, *moni_out = nullptr, *moni_host = nullptr;
    bool finalized = false;
    cudaEvent_t ev_fork = nullptr;
    // multi-rank: ownership, pairwise trace exchange, externally provided exchange buffers (device memory)
    std::vector<int> body_rank;
    int my_rank = 0;
    std::vector<int> peers;            // ranks this one shares interfaces with, ascending
    std::vector<long> peer_off;        // [npeers+1] ranges of the cross part of T, one per peer
    double *x_glob = nullptr, *x_send = nullptr, *x_recv = nullptr, *x_moni = nullptr;   // external (not owned) or null
    long cg_iters = 0;        // CG iterations of the last step, all bodies
    double cg_dof_iters = 0;  // sum over bodies of n_L * iterations, last step
    bool bodies_pending = false;
};

static void admm_free(ddpca_admm *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    for (auto &b : h->batch) {
        if (b.mg) ddpca_mg_destroy(b.mg);
        if (b.ev_done) cudaEventDestroy(b.ev_done);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    for (auto &f : h->iface)
        for (auto &s : f.side) { ldlt_free(s.mass); ldlt_free(s.mass_pena); }
    cudaFree(h->state); cudaFree(h->state_prev); cudaFree(h->addi); cudaFree(h->rhs); cudaFree(h->u); cudaFree(h->force); cudaFree(h->tmp);
    cudaFree(h->consForc); cudaFree(h->dispCons); cudaFree(h->t); cudaFree(h->gap); cudaFree(h->stat);
    for (DevCsr *m : {&h->OPA, &h->Fall, &h->FTall, &h->OPG, &h->OPG1, &h->ACCU, &h->OPT, &h->OPF1, &h->OPF2}) free_csr(*m);
    cudaFree(h->meta_d); cudaFree(h->side_off_d); cudaFree(h->mass_ptr_d); cudaFree(h->pena_ptr_d);
    ldlt_free(h->coar);
    if (h->coar_mg) ddpca_mg_destroy(h->coar_mg);
    ldlt_free(h->coar1);
    if (h->coar1_mg) ddpca_mg_destroy(h->coar1_mg);
    if (h->pcg_mass) ddpca_mg_destroy(h->pcg_mass);
    if (h->pcg_pena) ddpca_mg_destroy(h->pcg_pena);
    cudaFree(h->globForc); cudaFree(h->globSolu); cudaFree(h->globForc1_const); cudaFree(h->globSolu1);
    cudaFree(h->moni_chunks_d); cudaFree(h->slot_chunk_d); cudaFree(h->moni_part); cudaFree(h->moni_out);
    if (h->moni_host) cudaFreeHost(h->moni_host);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

static int host_csr(int rows, int cols, const int *rp, const int *ci, const double *v, CsrHost &o)
{
    if (rows < 0 || cols < 0 || !rp) return fail("bad CSR argument");
    o.rows = rows; o.cols = cols;
    o.rp.assign(rp, rp + rows + 1);
    long nnz = rp[rows];
    if (nnz && (!ci || !v)) return fail("bad CSR argument");
    o.ci.assign(ci, ci + nnz);
    o.v.assign(v, v + nnz);
    for (long p = 0; p < nnz; p++) if (ci[p] < 0 || ci[p] >= cols) return fail("CSR column index out of range");
    return 0;
}

static int dev_vec(const double *host, long n, double **d)
{
    CU(cudaMalloc(d, sizeof(double) * std::max<long>(1, n)));
    if (host && n) CU(cudaMemcpy(*d, host, sizeof(double) * n, cudaMemcpyHostToDevice));
    else CU(cudaMemset(*d, 0, sizeof(double) * std::max<long>(1, n)));
    return 0;
}

// ---- host assembly of stacked operators ------------------------------------------------------------
// out = sum of blocks: scale * A (or A^T) placed at (row_off, col_off).  Entries are kept in block order
// inside a row (k_spmv does not need sorted rows).
struct OpBlock { const CsrHost *A; int row_off, col_off; double scale; bool transpose; };
static int stack_blocks(int rows, int cols, const std::vector<OpBlock> &blocks, CsrHost &out)
{
    out.rows = rows; out.cols = cols;
    std::vector<long> cnt(rows + 1, 0);
    std::vector<CsrHost> tr(blocks.size());
    std::vector<const CsrHost *> src(blocks.size());
    // first entry of block k in every row it touches: blocks keep their order inside a row, the fill below is then
    // independent per (block, row)
    std::vector<std::vector<long>> first(blocks.size());
    for (size_t k = 0; k < blocks.size(); k++) {
        const OpBlock &b = blocks[k];
        if (b.transpose) { transpose_csr(*b.A, tr[k]); src[k] = &tr[k]; } else src[k] = b.A;
        const CsrHost &A = *src[k];
        if (b.row_off < 0 || b.col_off < 0 || b.row_off + A.rows > rows || b.col_off + A.cols > cols) return fail("stack_blocks: block outside the stacked operator");
        first[k].resize(A.rows);
        for (int i = 0; i < A.rows; i++) { first[k][i] = cnt[b.row_off + i + 1]; cnt[b.row_off + i + 1] += A.rp[i + 1] - A.rp[i]; }
    }
    for (int i = 0; i < rows; i++) cnt[i + 1] += cnt[i];
    if (cnt[rows] > 0x7ffffff0L) return fail("stacked interface operator too large for 32-bit indices");
    out.rp.resize(rows + 1);
    for (int i = 0; i <= rows; i++) out.rp[i] = (int)cnt[i];
    out.ci.resize(cnt[rows]);
    out.v.resize(cnt[rows]);
    for (size_t k = 0; k < blocks.size(); k++) {
        const OpBlock &b = blocks[k];
        const CsrHost &A = *src[k];
        const std::vector<long> &f = first[k];
#pragma omp parallel for schedule(static)
        for (int i = 0; i < A.rows; i++) {
            long q = cnt[b.row_off + i] + f[i];
            for (int p = A.rp[i]; p < A.rp[i + 1]; p++, q++) {
                out.ci[q] = A.ci[p] + b.col_off;
                out.v[q] = b.scale * A.v[p];
            }
        }
    }
    return 0;
}
static int stack_upload(int rows, int cols, const std::vector<OpBlock> &blocks, DevCsr &d)
{
    CsrHost H;
    if (stack_blocks(rows, cols, blocks, H)) return 1;
    return upload_csr(H, d);
}

#define ADMM_SPMV(A, x, y, add, alpha) launch_spmv(h, DDPCA_K_VECTOR, 15, (A), (x), (y), (add), nullptr, nullptr, nullptr, (alpha))

// ---- the loop body in phases; between phases a multi-rank caller exchanges one buffer ---------
enum { PH_BODIES = 0, PH_MACRO_PARTIAL = 1, PH_MACRO_APPLY = 2, PH_TRACES = 3, PH_INTERFACE = 4, PH_MONITOR = 5, PH_MACRO1_PARTIAL = 6, PH_MACRO1_APPLY = 7 };

static double *glob_buf(ddpca_admm *h) { return h->x_glob ? h->x_glob : h->globForc; }
static double *moni_buf(ddpca_admm *h) { return h->x_moni ? h->x_moni : h->moni_out; }

// wait for the batched body solves of this iteration and account their CG iterations
static int admm_bodies_finish(ddpca_admm *h)
{
    for (ddpca_mg *m : {h->pcg_pena, h->pcg_mass})
        if (m) { if (pcg_finish(m, nullptr, nullptr, nullptr)) return 1; h->launches += ddpca_mg_launch_count(m, 1); }
    if (!h->bodies_pending) return 0;
    h->bodies_pending = false;
    for (Batch &b : h->batch) {
        if (pcg_finish(b.mg, nullptr, nullptr, nullptr)) return 1;
        h->launches += ddpca_mg_launch_count(b.mg, 1);
        for (int s = 0; s < b.mg->nsub; s++) {
            const long it = (long)b.mg->st_host[s].it;
            h->cg_iters += it;
            h->cg_dof_iters += (double)it * b.mg->sub_n[s];
        }
    }
    return 0;
}

// body balance, MCONTACT.h:2507-2538, all local bodies at once
static int admm_bodies(ddpca_admm *h)
{
    cudaStream_t st = h->stream;
    h->cg_iters = 0;
    h->cg_dof_iters = 0;
    CU(cudaMemcpyAsync(h->state_prev, h->state, sizeof(double) * (size_t)(2 * h->NC + h->NF), cudaMemcpyDeviceToDevice, st));   // :2507-2509
    if (h->NR == 0) return 0;
    ADMM_SPMV(h->OPA, h->state, h->addi, false, 1.0);                                  // :2514-2523
    CU(cudaMemcpyAsync(h->rhs, h->consForc, sizeof(double) * h->NR, cudaMemcpyDeviceToDevice, st));
    ADMM_SPMV(h->Fall, h->addi, h->rhs, true, 1.0);                                    // ADDITIONAL_FORCE :2524 ; consForc + addiForc :2531
    // one batched MG-PCG per group of bodies with equal level count (MG-PCG for every body, :2531); several
    // groups run on their own streams and join afterwards
    const bool fork = h->batch.size() > 1 && !std::getenv("DDPCA_SERIAL_BODIES");
    if (fork) {
        if (!h->ev_fork) CU(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        CU(cudaEventRecord(h->ev_fork, st));
    }
    for (Batch &b : h->batch) {
        cudaStream_t bs = fork ? b.mg->own_stream : st;
        ddpca_mg_set_stream(b.mg, (void *)bs);
        if (fork) CU(cudaStreamWaitEvent(bs, h->ev_fork, 0));
        if (pcg_device(b.mg, 1, h->rhs + b.roff, h->u + b.roff, 1.0e-14, 0, nullptr, nullptr, nullptr, /*no_wait=*/true)) return 1;
        if (fork) {
            if (!b.ev_done) CU(cudaEventCreateWithFlags(&b.ev_done, cudaEventDisableTiming));
            CU(cudaEventRecord(b.ev_done, bs));
            CU(cudaStreamWaitEvent(st, b.ev_done, 0));
        }
    }
    h->bodies_pending = true;
    double *disp = h->state + h->offDisp();
    CU(cudaMemcpyAsync(disp, h->dispCons, sizeof(double) * h->NF, cudaMemcpyDeviceToDevice, st));
    ADMM_SPMV(h->FTall, h->u, disp, true, 1.0);                                        // OUTP_SUB1 :2533
    return 0;
}

// macroscopic problem, :2541-2549: this rank's part of globForc
static int admm_macro_partial(ddpca_admm *h)
{
    if (!h->coar && !h->coar_mg) return fail("macroscopic problem requested but not set");
    ADMM_SPMV(h->OPG, h->state, glob_buf(h), false, 1.0);
    return 0;
}
// resuDisp[v] += OUTP_SUB1(accuProl[v] * globSolu[baseReco[v] ...]) for the local bodies (:2564-2570, :2599-2604)
static int admm_coarse_correction(ddpca_admm *h, const double *globSolu)
{
    if (h->NR == 0) return 0;
    cudaStream_t st = h->stream;
    double *disp = h->state + h->offDisp();
    ADMM_SPMV(h->ACCU, globSolu, h->u, false, 1.0);                                                     // :2564-2567
    CU(cudaMemcpyAsync(h->addi, h->dispCons, sizeof(double) * h->NF, cudaMemcpyDeviceToDevice, st));    // OUTP_SUB1 = FT u + dispCons ...
    ADMM_SPMV(h->FTall, h->u, h->addi, true, 1.0);
    KL(h, DDPCA_K_VECTOR, 15, 24.0 * h->NF, (k_axpy<<<cdiv(h->NF, 256), 256, 0, st>>>(h->NF, 1.0, h->addi, disp)));   // ... added to resuDisp (:2569-2570)
    return 0;
}
// :2553-2572: replicated coarse solve, correction of the local bodies
static int admm_macro_apply(ddpca_admm *h)
{
    cudaStream_t st = h->stream;
    if (h->coar_mg) {
        // :2560-2562  mgpi.CG_SOLV(1, globForc, globSolu)
        long it = 0;
        ddpca_mg_set_stream(h->coar_mg, (void *)st);
        if (pcg_device(h->coar_mg, 1, glob_buf(h), h->globSolu, 1.0e-14, h->nglob, &it, nullptr, nullptr)) return 1;
        h->launches += ddpca_mg_launch_count(h->coar_mg, 1);
        h->macro_cg_iters += it;
    } else {
        ldlt_solve_on(h, h->coar, glob_buf(h), h->globSolu, nullptr);   // :2553
    }
    return admm_coarse_correction(h, h->globSolu);
}
// muscSett bit 1, :2576-2584: this rank's part of globForc - globForc_1
static int admm_macro1_partial(ddpca_admm *h)
{
    if (!h->coar1 && !h->coar1_mg) return fail("interface-eliminated coarse problem requested but not set");
    ADMM_SPMV(h->OPG1, h->state, glob_buf(h), false, 1.0);   // :2579, :2583
    return 0;
}
// :2576 (constant part, added once after the sum over ranks), :2585-2606
static int admm_macro1_apply(ddpca_admm *h)
{
    cudaStream_t st = h->stream;
    double *gf = glob_buf(h);
    KL(h, DDPCA_K_VECTOR, 15, 24.0 * h->nglob1, (k_axpy<<<cdiv(h->nglob1, 256), 256, 0, st>>>(h->nglob1, 1.0, h->globForc1_const, gf)));
    if (h->coar1_mg) {
        // :2593-2595  mgpi_1.CG_SOLV(1, globForc, globSolu)
        long it = 0;
        ddpca_mg_set_stream(h->coar1_mg, (void *)st);
        if (pcg_device(h->coar1_mg, 1, gf, h->globSolu1, 1.0e-14, h->nglob1, &it, nullptr, nullptr)) return 1;
        h->launches += ddpca_mg_launch_count(h->coar1_mg, 1);
        h->macro_cg_iters += it;
    } else {
        ldlt_solve_on(h, h->coar1, gf, h->globSolu1, nullptr);   // :2588
    }
    return admm_coarse_correction(h, h->globSolu1);
}
// t = inpoLagr0 l0 - inpoLagr1 l1 + pemaInpo_r0 u0 - pemaInpo_r1 u1 (:2632-2635); for cross-rank interfaces the local
// side's signed part, copied to the send buffer of the pairwise exchange
static int admm_traces(ddpca_admm *h)
{
    if (h->NG == 0) return 0;
    ADMM_SPMV(h->OPT, h->state, h->t, false, 1.0);
    const long ncross = h->NG - h->NGloc;
    if (ncross && h->x_send) CU(cudaMemcpyAsync(h->x_send, h->t + h->NGloc, sizeof(double) * ncross, cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}
// interface balance :2636-2685 and multiplier update :2689-2704 for the local sides
static int admm_interface(ddpca_admm *h)
{
    if (h->NG == 0) return 0;
    cudaStream_t st = h->stream;
    const long ncross = h->NG - h->NGloc;
    // own part + the peer's part: a + b == b + a bit for bit, both owners project the same gamma
    if (ncross && h->x_recv) KL(h, DDPCA_K_VECTOR, 15, 24.0 * ncross, (k_axpy<<<cdiv(ncross, 256), 256, 0, st>>>((int)ncross, 1.0, h->x_recv, h->t + h->NGloc)));
    double *gamma = h->state + h->offGamma();
    KL(h, DDPCA_K_VECTOR, 15, 28.0 * h->NG, (k_gamma_project_all<<<cdiv(std::max(h->nip_tot, 1), 256), 256, 0, st>>>(h->nact, h->meta_d, h->t, h->gap, gamma, h->stat)));  // :2636-2668
    if (h->NC == 0) return 0;
    double *aux = h->state + h->offAux(), *lagr = h->state + h->offLagr();
    ADMM_SPMV(h->OPF1, h->state, h->force, false, 1.0);                                 // :2671-2675
    KL(h, DDPCA_K_COARSE, 15, 0.0, (k_dense_gemv_batch<<<cdiv((long)h->NC * 32, 256), 256, 0, st>>>(h->nls, h->side_off_d, h->pena_ptr_d, h->force, aux, nullptr)));   // :2677
    for (auto &p : h->sparse_pena) { HostSide &s = h->iface[p.first].side[p.second]; ldlt_solve_on(h, s.mass_pena, h->force + s.coff, aux + s.coff, nullptr); }
    if (h->pcg_pena) {   // :2678-2683: Eigen CG with its default diagonal preconditioner -> batched Jacobi-PCG
        ddpca_mg_set_stream(h->pcg_pena, (void *)st);
        if (pcg_device(h->pcg_pena, 0, h->force + h->it_off, aux + h->it_off, 1.0e-15, 0, nullptr, nullptr, nullptr, /*no_wait=*/true)) return 1;
    }
    ADMM_SPMV(h->OPF2, h->state, h->force, false, 1.0);                                 // :2691-2694
    KL(h, DDPCA_K_COARSE, 15, 0.0, (k_dense_gemv_batch<<<cdiv((long)h->NC * 32, 256), 256, 0, st>>>(h->nls, h->side_off_d, h->mass_ptr_d, h->force, h->tmp, nullptr)));   // :2696
    for (auto &p : h->sparse_mass) { HostSide &s = h->iface[p.first].side[p.second]; ldlt_solve_on(h, s.mass, h->force + s.coff, h->tmp + s.coff, nullptr); }
    if (h->pcg_mass) {   // :2698-2703
        ddpca_mg_set_stream(h->pcg_mass, (void *)st);
        if (pcg_device(h->pcg_mass, 0, h->force + h->it_off, h->tmp + h->it_off, 1.0e-15, 0, nullptr, nullptr, nullptr, /*no_wait=*/true)) return 1;
    }
    KL(h, DDPCA_K_VECTOR, 15, 24.0 * h->NC, (k_axpy<<<cdiv(h->NC, 256), 256, 0, st>>>(h->NC, 1.0, h->tmp, lagr)));
    return 0;
}
// MONITOR sums, :2737-2833: slots of remote bodies / sides are zero (the caller all-reduces)
static int admm_monitor(ddpca_admm *h)
{
    if (h->nmchunks) KL(h, DDPCA_K_VECTOR, 15, 16.0 * (2 * h->NC + h->NF), (k_moni_seg<<<h->nmchunks, 256, 0, h->stream>>>(h->moni_chunks_d, h->state, h->state_prev, h->moni_part)));
    KL(h, DDPCA_K_VECTOR, 15, 0.0, (k_moni_slots<<<cdiv(h->nslots, 128), 128, 0, h->stream>>>(h->nslots, h->slot_chunk_d, h->moni_part, moni_buf(h))));
    return 0;
}
// row of resuMoni.txt (:2742-2743, :2777-2778, :2807-2808, :2835) from the sums in h->moni_host
static void admm_format_row(ddpca_admm *h, double *monitor_row)
{
    if (!monitor_row) return;
    double convValu = 0.0, convCrit = 0.0;
    int c = 0;
    for (int v = 0; v < h->nb; v++) {
        monitor_row[c++] = h->moni_host[2 * v];
        monitor_row[c++] = h->moni_host[2 * v + 1];
        convValu += h->moni_host[2 * v];
        convCrit += h->moni_host[2 * v + 1];
    }
    for (int ts = 0; ts < h->ni; ts++)
        for (int tv = 0; tv < 2; tv++) {
            int slot = h->nb + 4 * ts + 2 * tv;   // tempIndi of :2771
            monitor_row[c++] = h->moni_host[2 * slot];
            monitor_row[c++] = h->moni_host[2 * slot + 1];
            convValu += h->moni_host[2 * slot];
            convCrit += h->moni_host[2 * slot + 1];
            monitor_row[c++] = h->moni_host[2 * (slot + 1)];
            monitor_row[c++] = h->moni_host[2 * (slot + 1) + 1];
        }
    monitor_row[c++] = convValu;
    monitor_row[c++] = convCrit;
}
// wait for the iteration, fetch the (all-reduced) MONITOR sums, assemble the row
static int admm_row(ddpca_admm *h, double *monitor_row)
{
    CU(cudaMemcpyAsync(h->moni_host, moni_buf(h), sizeof(double) * 2 * h->nslots, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    if (admm_bodies_finish(h)) return 1;
    if (!h->launch_err.empty()) { std::string m = h->launch_err; h->launch_err.clear(); return fail(m); }
    if (h->profile) h->prof_collect();
    admm_format_row(h, monitor_row);
    return 0;
}

static int admm_step(ddpca_admm *h, int apply_macro, double *monitor_row)
{
    if (h->NG - h->NGloc > 0 && !h->x_recv) return fail("ddpca_admm_step: this rank shares interfaces with other ranks; drive the phases and exchange the traces (ddpca_admm_phase)");
    if (admm_bodies(h)) return 1;
    if (apply_macro && (h->muscSett & 1) && (admm_macro_partial(h) || admm_macro_apply(h))) return 1;
    if (apply_macro && (h->muscSett & 2) && (admm_macro1_partial(h) || admm_macro1_apply(h))) return 1;
    if (admm_traces(h) || admm_interface(h) || admm_monitor(h)) return 1;
    return admm_row(h, monitor_row);
}

extern "C" {

int ddpca_admm_create(int device, int nbody, int niface, int muscSett, ddpca_admm **out)
{
    if (!out || nbody < 1 || niface < 0) return fail("ddpca_admm_create: bad argument");
    if (muscSett & ~3) return fail("ddpca_admm_create: muscSett has bits 0 (macroscopic problem) and 1 (interface-eliminated problem) only");
    int ndev = ddpca_device_count();
    if (ndev == 0) return fail("no CUDA device: libddpca_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail("device index out of range");
    CU(cudaSetDevice(device));
    ddpca_admm *h = new ddpca_admm();
    h->device = device;
    h->nb = nbody; h->ni = niface; h->muscSett = muscSett;
    h->body.resize(nbody);
    h->iface.resize(niface);
    cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, device);
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return fail("stream creation failed"); }
    h->stream = h->own_stream;
    *out = h;
    return 0;
}

int ddpca_admm_destroy(ddpca_admm *h) { admm_free(h); return 0; }

int ddpca_admm_set_body(ddpca_admm *h, int v, int nlevels, const int *n, const int *const *rowptr, const int *const *colidx,
                        const double *const *val, const int *const *P_rowptr, const int *const *P_colidx, const double *const *P_val,
                        int nfull, const double *consForc, const int *F_rowptr, const int *F_colidx, const double *F_val, const double *dispCons)
{
    if (!h || v < 0 || v >= h->nb || nlevels < 1 || nlevels > 16 || !n || !rowptr || !colidx || !val || nfull < 1 || !consForc || !F_rowptr || !dispCons)
        return fail("ddpca_admm_set_body: bad argument");
    if (nlevels > 1 && (!P_rowptr || !P_colidx || !P_val)) return fail("ddpca_admm_set_body: prolongation operators missing");
    if (h->finalized) return fail("ddpca_admm_set_body after ddpca_admm_finalize");
    HostBody &b = h->body[v];
    if (b.set) return fail("ddpca_admm_set_body: body already set");
    if (!b.local) return fail("ddpca_admm_set_body: body " + std::to_string(v) + " belongs to another rank");
    b.nlev = nlevels;
    b.n.assign(n, n + nlevels);
    b.rp.assign(rowptr, rowptr + nlevels); b.ci.assign(colidx, colidx + nlevels); b.v.assign(val, val + nlevels);
    if (nlevels > 1) { b.prp.assign(P_rowptr, P_rowptr + nlevels - 1); b.pci.assign(P_colidx, P_colidx + nlevels - 1); b.pv.assign(P_val, P_val + nlevels - 1); }
    b.nfull = nfull;
    b.nred = n[nlevels - 1];
    if (host_csr(b.nred, nfull, F_rowptr, F_colidx, F_val, b.F)) return 1;
    b.consForc.assign(consForc, consForc + b.nred);
    b.dispCons.assign(dispCons, dispCons + nfull);
    b.set = true;
    return 0;
}

int ddpca_admm_set_body_accuprol(ddpca_admm *h, int v, int rows, int cols, const int *rowptr, const int *colidx, const double *val)
{
    if (!h || v < 0 || v >= h->nb || !h->body[v].set || h->finalized) return fail("ddpca_admm_set_body_accuprol: bad argument");
    if (rows != h->body[v].nred) return fail("accuProl must have n_L rows");
    return host_csr(rows, cols, rowptr, colidx, val, h->body[v].accuProl);
}

int ddpca_admm_set_interface(ddpca_admm *h, int ts, int body0, int body1, double fricCoef, int nip, const double *gapTerm)
{
    if (!h || ts < 0 || ts >= h->ni || body0 < 0 || body0 >= h->nb || body1 < 0 || body1 >= h->nb || nip < 0 || !gapTerm || h->finalized)
        return fail("ddpca_admm_set_interface: bad argument");
    HostIface &f = h->iface[ts];
    if (f.set) return fail("interface already set");
    f.body[0] = body0; f.body[1] = body1;
    f.side[0].local = h->body[body0].local;
    f.side[1].local = h->body[body1].local;
    f.cross = !h->body_rank.empty() && h->body_rank[body0] != h->body_rank[body1];
    if (f.cross) f.peer = f.side[0].local ? h->body_rank[body1] : h->body_rank[body0];
    f.fric = fricCoef;
    f.nip = nip;
    f.d = (fricCoef == 0.0) ? 1 : 3;   // MCONTACT.h:886-893
    f.gap.assign(gapTerm, gapTerm + (size_t)f.d * nip);
    f.set = true;
    return 0;
}

int ddpca_admm_set_side_op(ddpca_admm *h, int ts, int tv, int op, int rows, int cols, const int *rowptr, const int *colidx, const double *val)
{
    if (!h || ts < 0 || ts >= h->ni || tv < 0 || tv > 1 || op < 0 || op >= DDPCA_OP_COUNT || !h->iface[ts].set || h->finalized) return fail("ddpca_admm_set_side_op: bad argument");
    HostSide &s = h->iface[ts].side[tv];
    if (!s.local) return fail("ddpca_admm_set_side_op: this side belongs to another rank");
    if (host_csr(rows, cols, rowptr, colidx, val, s.op[op])) return 1;
    s.has[op] = true;
    if (op == DDPCA_OP_INTEMASS) s.nc = rows;
    return 0;
}

int ddpca_admm_set_side_solver(ddpca_admm *h, int ts, int tv, int which, ddpca_ldlt *sol)
{
    if (!h || ts < 0 || ts >= h->ni || tv < 0 || tv > 1 || !sol || (which != DDPCA_SOLVER_MASS && which != DDPCA_SOLVER_MASS_PENA)) return fail("ddpca_admm_set_side_solver: bad argument");
    if (sol->device != h->device) return fail("solver lives on another device");
    if (!h->iface[ts].set) return fail("ddpca_admm_set_side_solver: interface not set");
    if (h->finalized) return fail("ddpca_admm_set_side_solver after ddpca_admm_finalize");
    HostSide &s = h->iface[ts].side[tv];
    if (!s.local) return fail("ddpca_admm_set_side_solver: this side belongs to another rank");
    if (which == DDPCA_SOLVER_MASS) { ldlt_free(s.mass); s.mass = sol; }
    else { ldlt_free(s.mass_pena); s.mass_pena = sol; }
    return 0;
}

int ddpca_admm_set_side_iterative(ddpca_admm *h, int ts, int tv)
{
    if (!h || ts < 0 || ts >= h->ni || tv < 0 || tv > 1 || !h->iface[ts].set || h->finalized) return fail("ddpca_admm_set_side_iterative: bad argument");
    HostSide &s = h->iface[ts].side[tv];
    if (!s.local) return fail("ddpca_admm_set_side_iterative: this side belongs to another rank");
    s.iterative = true;
    return 0;
}

int ddpca_admm_set_macro(ddpca_admm *h, int nglob, const long *baseReco, ddpca_ldlt *coarSolv)
{
    if (!h || nglob < 1 || !baseReco || !coarSolv) return fail("ddpca_admm_set_macro: bad argument");
    if (coarSolv->n != nglob) return fail("coarse solver size does not match globCoup");
    h->nglob = nglob;
    h->baseReco.assign(baseReco, baseReco + h->nb + 1);
    ldlt_free(h->coar);
    h->coar = coarSolv;
    return 0;
}

int ddpca_admm_set_macro_mg(ddpca_admm *h, int nglob, const long *baseReco, ddpca_mg *mgpi)
{
    if (!h || nglob < 1 || !baseReco || !mgpi) return fail("ddpca_admm_set_macro_mg: bad argument");
    if (mgpi->nsub != 1 || mgpi->lev[mgpi->nlev - 1].n != nglob) return fail("finest level of the macroscopic hierarchy does not match globCoup");
    if (mgpi->device != h->device) return fail("macroscopic hierarchy lives on another device");
    h->nglob = nglob;
    h->baseReco.assign(baseReco, baseReco + h->nb + 1);
    ldlt_free(h->coar);
    h->coar = nullptr;
    if (h->coar_mg && h->coar_mg != mgpi) ddpca_mg_destroy(h->coar_mg);
    h->coar_mg = mgpi;
    return 0;
}

int ddpca_admm_set_body_globtran_d1(ddpca_admm *h, int v, int rows, int cols, const int *rowptr, const int *colidx, const double *val)
{
    if (!h || v < 0 || v >= h->nb || !h->body[v].set || h->finalized) return fail("ddpca_admm_set_body_globtran_d1: bad argument");
    if (cols != h->body[v].nfull) return fail("globTran_D_1 must have 3 n_nodes columns");
    return host_csr(rows, cols, rowptr, colidx, val, h->body[v].globTran_D_1);
}

int ddpca_admm_set_macro1(ddpca_admm *h, int nglob1, const long *baseReco, const double *globForc_1, ddpca_ldlt *coarSolv_D_1)
{
    if (!h || nglob1 < 1 || !baseReco || !globForc_1 || !coarSolv_D_1) return fail("ddpca_admm_set_macro1: bad argument");
    if (coarSolv_D_1->n != nglob1) return fail("coarse solver size does not match globCoup_1");
    CU(cudaSetDevice(h->device));
    h->nglob1 = nglob1;
    h->baseReco.assign(baseReco, baseReco + h->nb + 1);
    ldlt_free(h->coar1);
    h->coar1 = coarSolv_D_1;
    cudaFree(h->globForc1_const);
    h->globForc1_const = nullptr;
    return dev_vec(globForc_1, nglob1, &h->globForc1_const);
}

int ddpca_admm_set_macro1_mg(ddpca_admm *h, int nglob1, const long *baseReco, const double *globForc_1, ddpca_mg *mgpi_1)
{
    if (!h || nglob1 < 1 || !baseReco || !globForc_1 || !mgpi_1) return fail("ddpca_admm_set_macro1_mg: bad argument");
    if (mgpi_1->nsub != 1 || mgpi_1->lev[mgpi_1->nlev - 1].n != nglob1) return fail("finest level of the hierarchy does not match globCoup_1");
    if (mgpi_1->device != h->device) return fail("hierarchy lives on another device");
    CU(cudaSetDevice(h->device));
    h->nglob1 = nglob1;
    h->baseReco.assign(baseReco, baseReco + h->nb + 1);
    ldlt_free(h->coar1);
    h->coar1 = nullptr;
    if (h->coar1_mg && h->coar1_mg != mgpi_1) ddpca_mg_destroy(h->coar1_mg);
    h->coar1_mg = mgpi_1;
    cudaFree(h->globForc1_const);
    h->globForc1_const = nullptr;
    return dev_vec(globForc_1, nglob1, &h->globForc1_const);
}

// Builds the device side of the loop: batches of body hierarchies, concatenated state, stacked operators.
int ddpca_admm_finalize(ddpca_admm *h)
{
    if (!h) return fail("null handle");
    if (h->finalized) return 0;
    CU(cudaSetDevice(h->device));
    const int nb = h->nb, ni = h->ni;
    StageTimer tm("admm finalize");
    // ---- completeness --------------------------------------------------------------------------
    for (int v = 0; v < nb; v++) {
        HostBody &b = h->body[v];
        if (!b.local) continue;
        if (!b.set) return fail("body " + std::to_string(v) + " not set");
        if ((h->muscSett & 3) && b.accuProl.rows == 0) return fail("body " + std::to_string(v) + ": accuProl missing");
        if ((h->muscSett & 1) && h->baseReco.size() == (size_t)nb + 1 && h->baseReco[v] + b.accuProl.cols > h->nglob) return fail("baseReco out of range");
        if ((h->muscSett & 2) && h->baseReco.size() == (size_t)nb + 1 && h->baseReco[v] + b.accuProl.cols > h->nglob1) return fail("baseReco out of range (globCoup_1)");
        if ((h->muscSett & 2) && (b.globTran_D_1.rp.empty() || b.globTran_D_1.rows != h->nglob1)) return fail("body " + std::to_string(v) + ": globTran_D_1 missing");
    }
    if ((h->muscSett & 1) && !h->coar && !h->coar_mg) return fail("macroscopic problem not set");
    if ((h->muscSett & 2) && !h->coar1 && !h->coar1_mg) return fail("interface-eliminated coarse problem not set");
    for (int ts = 0; ts < ni; ts++) {
        HostIface &f = h->iface[ts];
        if (!f.set) return fail("interface " + std::to_string(ts) + " not set");
        for (int tv = 0; tv < 2; tv++) {
            HostSide &s = f.side[tv];
            if (!s.local) continue;
            int need[] = {DDPCA_OP_SYSTTRAN, DDPCA_OP_SYSTTRAN_PENA, DDPCA_OP_INTEMASS, DDPCA_OP_INTEMASS_PENA, DDPCA_OP_INPOLAGR, DDPCA_OP_INTEINPO, DDPCA_OP_PEMAINPO_R};
            for (int o : need) if (!s.has[o]) return fail("interface " + std::to_string(ts) + " side " + std::to_string(tv) + ": operator " + std::to_string(o) + " missing");
            if (h->muscSett & 1)
                for (int o : {DDPCA_OP_GLOBTRAN, DDPCA_OP_GLOBTRAN_PENA, DDPCA_OP_GLOBTRAN_D}) if (!s.has[o]) return fail("macroscopic transfer operator missing");
            if ((h->muscSett & 2) && !s.has[DDPCA_OP_GLOBTRAN_1]) return fail("globTran_1 missing");
            if (!s.iterative && (!s.mass || !s.mass_pena)) return fail("interface " + std::to_string(ts) + " side " + std::to_string(tv) + ": mass solvers missing");
            int ng = f.d * f.nip, nfull = h->body[f.body[tv]].nfull;
            if (s.op[DDPCA_OP_SYSTTRAN].rows != nfull || s.op[DDPCA_OP_SYSTTRAN].cols != s.nc || s.op[DDPCA_OP_INPOLAGR].rows != ng ||
                s.op[DDPCA_OP_INTEINPO].rows != s.nc || s.op[DDPCA_OP_INTEINPO].cols != ng || s.op[DDPCA_OP_PEMAINPO_R].rows != ng ||
                s.op[DDPCA_OP_PEMAINPO_R].cols != nfull || (!s.iterative && (s.mass->n != s.nc || s.mass_pena->n != s.nc)))
                return fail("interface " + std::to_string(ts) + " side " + std::to_string(tv) + ": operator shapes are inconsistent");
        }
    }
    // ---- numbering: bodies grouped by level count (one batch each), sides in (ts, tv) order, interfaces with both
    //      sides here first, then the cross-rank ones grouped by peer (their part of T is what is exchanged) ----------
    std::vector<int> lb;
    for (int v = 0; v < nb; v++) if (h->body[v].local) lb.push_back(v);
    std::stable_sort(lb.begin(), lb.end(), [&](int a, int b) { return h->body[a].nlev < h->body[b].nlev; });
    long NF = 0, NR = 0;
    for (size_t k = 0; k < lb.size(); k++) {
        HostBody &b = h->body[lb[k]];
        if (k == 0 || b.nlev != h->body[lb[k - 1]].nlev) { Batch bt; bt.roff = (int)NR; h->batch.push_back(bt); }
        Batch &bt = h->batch.back();
        b.batch = (int)h->batch.size() - 1;
        b.sub = (int)bt.bodies.size();
        bt.bodies.push_back(lb[k]);
        b.foff = (int)NF; b.roff = (int)NR;
        NF += b.nfull; NR += b.nred;
        bt.nred += b.nred;
    }
    long NC = 0;
    h->nls = 0;
    std::vector<HostSide *> sorder;   // local sides in numbering order: those with a direct solver, then the iterative ones
    for (int pass = 0; pass < 2; pass++)
        for (int ts = 0; ts < ni; ts++)
            for (int tv = 0; tv < 2; tv++) {
                HostSide &s = h->iface[ts].side[tv];
                if (!s.local || (int)s.iterative != pass) continue;
                if (pass == 1 && h->it_rows == 0) h->it_off = (int)NC;
                s.coff = (int)NC;
                NC += s.nc;
                if (pass == 1) h->it_rows += s.nc;
                h->nls++;
                sorder.push_back(&s);
            }
    std::vector<int> act;   // active interfaces in T order
    for (int ts = 0; ts < ni; ts++) if (h->iface[ts].side[0].local && h->iface[ts].side[1].local) act.push_back(ts);
    long NG = 0;
    for (int ts : act) { h->iface[ts].goff = (int)NG; NG += (long)h->iface[ts].d * h->iface[ts].nip; }
    h->NGloc = (int)NG;
    h->peers.clear();
    for (int ts = 0; ts < ni; ts++) { HostIface &f = h->iface[ts]; if (f.cross && (f.side[0].local || f.side[1].local)) h->peers.push_back(f.peer); }
    std::sort(h->peers.begin(), h->peers.end());
    h->peers.erase(std::unique(h->peers.begin(), h->peers.end()), h->peers.end());
    h->peer_off.assign(h->peers.size() + 1, 0);
    for (size_t k = 0; k < h->peers.size(); k++) {
        h->peer_off[k] = NG - h->NGloc;
        for (int ts = 0; ts < ni; ts++) {
            HostIface &f = h->iface[ts];
            if (!(f.cross && (f.side[0].local || f.side[1].local) && f.peer == h->peers[k])) continue;
            f.goff = (int)NG;
            NG += (long)f.d * f.nip;
            act.push_back(ts);
        }
    }
    h->peer_off[h->peers.size()] = NG - h->NGloc;
    if (2 * NC + NF + NG > 0x7ffffff0L) return fail("state too large for 32-bit indices");
    h->NC = (int)NC; h->NF = (int)NF; h->NR = (int)NR; h->NG = (int)NG;
    const int ncols = 2 * h->NC + h->NF + h->NG;
    const int cA = h->offAux(), cL = h->offLagr(), cD = h->offDisp(), cG = h->offGamma();
    // ---- batched hierarchies ----------------------------------------------------------------------
    for (Batch &bt : h->batch) {
        const int ns = (int)bt.bodies.size(), nl = h->body[bt.bodies[0]].nlev;
        std::vector<int> n((size_t)ns * nl);
        std::vector<const int *> rp((size_t)ns * nl), ci((size_t)ns * nl), prp((size_t)ns * std::max(1, nl - 1)), pci((size_t)ns * std::max(1, nl - 1));
        std::vector<const double *> vv((size_t)ns * nl), pv((size_t)ns * std::max(1, nl - 1));
        for (int s = 0; s < ns; s++) {
            HostBody &b = h->body[bt.bodies[s]];
            for (int l = 0; l < nl; l++) { n[s * nl + l] = b.n[l]; rp[s * nl + l] = b.rp[l]; ci[s * nl + l] = b.ci[l]; vv[s * nl + l] = b.v[l]; }
            for (int l = 0; l + 1 < nl; l++) { prp[s * (nl - 1) + l] = b.prp[l]; pci[s * (nl - 1) + l] = b.pci[l]; pv[s * (nl - 1) + l] = b.pv[l]; }
        }
        int mode = std::getenv("DDPCA_SMOOTHER") && std::string(std::getenv("DDPCA_SMOOTHER")) == "lex" ? DDPCA_SMOOTH_LEX : DDPCA_SMOOTH_MC;
        if (h->smoother_mode >= 0) mode = h->smoother_mode;
        if (mg_create_impl(h->device, ns, nl, n.data(), rp.data(), ci.data(), vv.data(), prp.data(), pci.data(), pv.data(), mode, &bt.mg)) {
            g_err = "bodies of batch " + std::to_string(&bt - h->batch.data()) + ": " + g_err;
            return 1;
        }
    }
    tm.lap("batched hierarchies");
    // ---- vectors -------------------------------------------------------------------------------------
    if (dev_vec(nullptr, ncols, &h->state) || dev_vec(nullptr, 2 * h->NC + h->NF, &h->state_prev) || dev_vec(nullptr, h->NF, &h->addi) ||
        dev_vec(nullptr, h->NR, &h->rhs) || dev_vec(nullptr, h->NR, &h->u) || dev_vec(nullptr, h->NC, &h->force) || dev_vec(nullptr, h->NC, &h->tmp) ||
        dev_vec(nullptr, h->NG, &h->t)) return 1;
    {
        std::vector<double> cf(h->NR), dc(h->NF), gp(h->NG);
        for (int v : lb) {
            HostBody &b = h->body[v];
            std::copy(b.consForc.begin(), b.consForc.end(), cf.begin() + b.roff);
            std::copy(b.dispCons.begin(), b.dispCons.end(), dc.begin() + b.foff);
        }
        for (int ts : act) std::copy(h->iface[ts].gap.begin(), h->iface[ts].gap.end(), gp.begin() + h->iface[ts].goff);
        if (dev_vec(cf.data(), h->NR, &h->consForc) || dev_vec(dc.data(), h->NF, &h->dispCons) || dev_vec(gp.data(), h->NG, &h->gap)) return 1;
        CU(cudaMalloc(&h->stat, sizeof(int) * std::max(1, h->NG)));
        CU(cudaMemset(h->stat, 0, sizeof(int) * std::max(1, h->NG)));
    }
    // ---- stacked operators -----------------------------------------------------------------------------
    {
        std::vector<OpBlock> A, F, FT, G, G1, AC, T, F1, F2;
        for (int v : lb) {
            HostBody &b = h->body[v];
            F.push_back({&b.F, b.roff, b.foff, 1.0, false});
            FT.push_back({&b.F, b.foff, b.roff, 1.0, true});
            if (h->muscSett & 3) AC.push_back({&b.accuProl, b.roff, (int)h->baseReco[v], 1.0, false});
            if (h->muscSett & 2) G1.push_back({&b.globTran_D_1, 0, cD + b.foff, -1.0, false});          // :2583
        }
        for (int ts = 0; ts < ni; ts++) {
            HostIface &f = h->iface[ts];
            for (int tv = 0; tv < 2; tv++) {
                HostSide &s = f.side[tv];
                if (!s.local) continue;
                HostBody &b = h->body[f.body[tv]];
                const double sg = tv == 0 ? 1.0 : -1.0;
                A.push_back({&s.op[DDPCA_OP_SYSTTRAN_PENA], b.foff, cA + s.coff, 1.0, false});          // :2520
                A.push_back({&s.op[DDPCA_OP_SYSTTRAN], b.foff, cL + s.coff, -1.0, false});              // :2521
                if (h->muscSett & 1) {
                    G.push_back({&s.op[DDPCA_OP_GLOBTRAN], 0, cL + s.coff, 1.0, false});                // :2545
                    G.push_back({&s.op[DDPCA_OP_GLOBTRAN_PENA], 0, cA + s.coff, -1.0, false});          // :2546
                    G.push_back({&s.op[DDPCA_OP_GLOBTRAN_D], 0, cD + b.foff, 1.0, false});              // :2547
                }
                if (h->muscSett & 2) G1.push_back({&s.op[DDPCA_OP_GLOBTRAN_1], 0, cL + s.coff, 1.0, false});   // :2579
                T.push_back({&s.op[DDPCA_OP_INPOLAGR], f.goff, cL + s.coff, sg, false});                // :2632-2633
                T.push_back({&s.op[DDPCA_OP_PEMAINPO_R], f.goff, cD + b.foff, sg, false});              // :2634-2635
                F1.push_back({&s.op[DDPCA_OP_SYSTTRAN_PENA], s.coff, cD + b.foff, 1.0, true});          // :2673
                F1.push_back({&s.op[DDPCA_OP_INTEMASS], s.coff, cL + s.coff, 1.0, false});              // :2674
                F1.push_back({&s.op[DDPCA_OP_INTEINPO], s.coff, cG + f.goff, 1.0, false});              // :2675
                F2.push_back({&s.op[DDPCA_OP_SYSTTRAN_PENA], s.coff, cD + b.foff, 1.0, true});          // :2693
                F2.push_back({&s.op[DDPCA_OP_INTEMASS_PENA], s.coff, cA + s.coff, -1.0, false});        // :2694
            }
        }
        if (stack_upload(h->NF, ncols, A, h->OPA) || stack_upload(h->NR, h->NF, F, h->Fall) || stack_upload(h->NF, h->NR, FT, h->FTall) ||
            stack_upload(h->NG, ncols, T, h->OPT) || stack_upload(h->NC, ncols, F1, h->OPF1) || stack_upload(h->NC, ncols, F2, h->OPF2)) return 1;
        if ((h->muscSett & 1) && stack_upload(h->nglob, ncols, G, h->OPG)) return 1;
        if ((h->muscSett & 2) && stack_upload(h->nglob1, ncols, G1, h->OPG1)) return 1;
        if ((h->muscSett & 3) && stack_upload(h->NR, std::max(h->nglob, h->nglob1), AC, h->ACCU)) return 1;
    }
    tm.lap("stacked operators");
    if (h->muscSett & 3) { if (dev_vec(nullptr, std::max(h->nglob, h->nglob1), &h->globForc)) return 1; }
    if (h->muscSett & 1) { if (dev_vec(nullptr, h->nglob, &h->globSolu)) return 1; }
    if (h->muscSett & 2) { if (dev_vec(nullptr, h->nglob1, &h->globSolu1)) return 1; }
    // ---- projection table ------------------------------------------------------------------------------
    {
        std::vector<IfaceMeta> meta;
        int ip0 = 0;
        for (int ts : act) { HostIface &f = h->iface[ts]; meta.push_back(IfaceMeta{ip0, f.nip, f.d, f.goff, f.fric}); ip0 += f.nip; }
        h->nact = (int)meta.size();
        h->nip_tot = ip0;
        if (meta.empty()) meta.push_back(IfaceMeta{0, 0, 1, 0, 0.0});
        if (upload_vec(meta, &h->meta_d)) return 1;
    }
    // ---- interface mass solvers: dense inverses in one block-diagonal product ----------------------------
    {
        std::vector<int> off(h->nls + 1, 0);
        std::vector<double *> pm(std::max(1, h->nls), nullptr), pp(std::max(1, h->nls), nullptr);
        int k = 0;
        std::vector<HostSide *> its;
        for (HostSide *sp : sorder) {
            HostSide &s = *sp;
            off[k] = s.coff; off[k + 1] = s.coff + s.nc;
            if (s.iterative) its.push_back(sp);
            else { pm[k] = s.mass->Binv; pp[k] = s.mass_pena->Binv; }
            k++;
        }
        for (int ts = 0; ts < ni; ts++)
            for (int tv = 0; tv < 2; tv++) {
                HostSide &s = h->iface[ts].side[tv];
                if (!s.local || s.iterative) continue;
                if (!s.mass->Binv) h->sparse_mass.push_back({ts, tv});
                if (!s.mass_pena->Binv) h->sparse_pena.push_back({ts, tv});
            }
        if (upload_vec(off, &h->side_off_d) || upload_vec(pm, &h->mass_ptr_d) || upload_vec(pp, &h->pena_ptr_d)) return 1;
        if (!its.empty()) {
            // the mass matrices themselves, each a one-level "hierarchy", as two batches for Jacobi-PCG
            for (int which = 0; which < 2; which++) {
                const int op = which == 0 ? DDPCA_OP_INTEMASS : DDPCA_OP_INTEMASS_PENA;
                std::vector<int> n(its.size());
                std::vector<const int *> rp(its.size()), ci(its.size());
                std::vector<const double *> vv(its.size());
                for (size_t q = 0; q < its.size(); q++) { const CsrHost &M = its[q]->op[op]; n[q] = M.rows; rp[q] = M.rp.data(); ci[q] = M.ci.data(); vv[q] = M.v.data(); }
                ddpca_mg **dst = which == 0 ? &h->pcg_mass : &h->pcg_pena;
                if (mg_create_impl(h->device, (int)its.size(), 1, n.data(), rp.data(), ci.data(), vv.data(), nullptr, nullptr, nullptr, DDPCA_SMOOTH_MC, dst, /*no_direct=*/true)) {
                    g_err = "iterative interface mass solvers: " + g_err;
                    return 1;
                }
            }
        }
    }
    // ---- MONITOR tables: slot of body v is v, of side (ts, tv) nb + 4 ts + 2 tv (aux) and + 1 (lagr), :2771 ---
    {
        h->nslots = nb + 4 * ni;
        std::vector<std::vector<SegChunk>> per_slot(h->nslots);
        auto cut = [&](int slot, int row0, int n) { for (int r = 0; r < n; r += kSegRows) per_slot[slot].push_back(SegChunk{row0 + r, std::min(kSegRows, n - r), slot, 0}); };
        for (int v : lb) cut(v, cD + h->body[v].foff, h->body[v].nfull);
        for (int ts = 0; ts < ni; ts++)
            for (int tv = 0; tv < 2; tv++) {
                HostSide &s = h->iface[ts].side[tv];
                if (!s.local) continue;
                cut(nb + 4 * ts + 2 * tv, cA + s.coff, s.nc);
                cut(nb + 4 * ts + 2 * tv + 1, cL + s.coff, s.nc);
            }
        std::vector<SegChunk> ch;
        std::vector<int> sc(h->nslots + 1, 0);
        for (int s = 0; s < h->nslots; s++) { sc[s] = (int)ch.size(); ch.insert(ch.end(), per_slot[s].begin(), per_slot[s].end()); }
        sc[h->nslots] = (int)ch.size();
        h->nmchunks = (int)ch.size();
        if (ch.empty()) ch.push_back(SegChunk{0, 0, 0, 0});
        if (upload_vec(ch, &h->moni_chunks_d) || upload_vec(sc, &h->slot_chunk_d)) return 1;
        CU(cudaMalloc(&h->moni_part, sizeof(double) * 2 * std::max(1, h->nmchunks)));
        CU(cudaMalloc(&h->moni_out, sizeof(double) * 2 * h->nslots));
        CU(cudaMallocHost(&h->moni_host, sizeof(double) * 2 * h->nslots));
    }
    // host copies of the operators are no longer needed
    for (auto &f : h->iface) for (auto &s : f.side) for (auto &o : s.op) o = CsrHost();
    for (auto &b : h->body) { b.F = CsrHost(); b.accuProl = CsrHost(); b.globTran_D_1 = CsrHost(); b.rp.clear(); b.ci.clear(); b.v.clear(); b.prp.clear(); b.pci.clear(); b.pv.clear(); }
    // the dense inverses of the side / coarse solvers were only enqueued at their creation (mg.cu, setup_stream): the
    // device has been working on them while the host built the hierarchies and the stacked operators above
    if (setup_stream_sync(h->device)) return 1;
    tm.lap("pending dense inversions");
    h->finalized = true;
    return 0;
}

int ddpca_admm_step(ddpca_admm *h, int apply_macro, double *monitor_row, long *cg_iters, double *cg_dof_iters)
{
    if (!h || !h->finalized) return fail("ddpca_admm_step: handle not finalized");
    CU(cudaSetDevice(h->device));
    if (admm_step(h, apply_macro, monitor_row)) return 1;
    if (cg_iters) *cg_iters = h->cg_iters;
    if (cg_dof_iters) *cg_dof_iters = h->cg_dof_iters;
    return 0;
}

int ddpca_admm_set_partition(ddpca_admm *h, const int *body_rank, int my_rank)
{
    if (!h || !body_rank) return fail("ddpca_admm_set_partition: bad argument");
    for (int v = 0; v < h->nb; v++) if (h->body[v].set) return fail("ddpca_admm_set_partition must precede ddpca_admm_set_body");
    for (int ts = 0; ts < h->ni; ts++) if (h->iface[ts].set) return fail("ddpca_admm_set_partition must precede ddpca_admm_set_interface");
    h->body_rank.assign(body_rank, body_rank + h->nb);
    h->my_rank = my_rank;
    for (int v = 0; v < h->nb; v++) h->body[v].local = (body_rank[v] == my_rank);
    return 0;
}
int ddpca_admm_exchange_sizes(const ddpca_admm *h, long *nglob, long *ntrace, long *nmoni)
{
    if (!h) return fail("null handle");
    // from the declared interfaces; the layout inside the trace buffers is fixed by ddpca_admm_finalize
    long nt = 0;
    for (int ts = 0; ts < h->ni; ts++) { const HostIface &f = h->iface[ts]; if (f.set && f.cross && (f.side[0].local || f.side[1].local)) nt += (long)f.d * f.nip; }
    if (nglob) *nglob = std::max(h->nglob, h->nglob1);
    if (ntrace) *ntrace = nt;
    if (nmoni) *nmoni = 2L * (h->nb + 4 * h->ni);
    return 0;
}
int ddpca_admm_set_exchange(ddpca_admm *h, double *globForc_dev, double *trace_send_dev, double *trace_recv_dev, double *moni_dev)
{
    if (!h) return fail("null handle");
    h->x_glob = globForc_dev; h->x_send = trace_send_dev; h->x_recv = trace_recv_dev; h->x_moni = moni_dev;
    return 0;
}
int ddpca_admm_exchange_peers(const ddpca_admm *h, int *npeers, int *peer_rank, long *offset, long *count)
{
    if (!h || !h->finalized) return fail("ddpca_admm_exchange_peers: handle not finalized");
    if (npeers) *npeers = (int)h->peers.size();
    for (size_t k = 0; k < h->peers.size(); k++) {
        if (peer_rank) peer_rank[k] = h->peers[k];
        if (offset) offset[k] = h->peer_off[k];
        if (count) count[k] = h->peer_off[k + 1] - h->peer_off[k];
    }
    return 0;
}
int ddpca_admm_phase(ddpca_admm *h, int phase)
{
    if (!h || !h->finalized) return fail("ddpca_admm_phase: handle not finalized");
    CU(cudaSetDevice(h->device));
    switch (phase) {
    case PH_BODIES: return admm_bodies(h);
    case PH_MACRO_PARTIAL: return admm_macro_partial(h);
    case PH_MACRO_APPLY: return admm_macro_apply(h);
    case PH_TRACES: return admm_traces(h);
    case PH_INTERFACE: return admm_interface(h);
    case PH_MONITOR: return admm_monitor(h);
    case PH_MACRO1_PARTIAL: return admm_macro1_partial(h);
    case PH_MACRO1_APPLY: return admm_macro1_apply(h);
    }
    return fail("ddpca_admm_phase: unknown phase");
}
int ddpca_admm_monitor_row(ddpca_admm *h, double *monitor_row, long *cg_iters, double *cg_dof_iters)
{
    if (!h || !h->finalized) return fail("ddpca_admm_monitor_row: handle not finalized");
    CU(cudaSetDevice(h->device));
    if (admm_row(h, monitor_row)) return 1;
    if (cg_iters) *cg_iters = h->cg_iters;
    if (cg_dof_iters) *cg_dof_iters = h->cg_dof_iters;
    return 0;
}
int ddpca_admm_set_stream(ddpca_admm *h, void *stream)
{
    if (!h) return fail("null handle");
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    return 0;
}
int ddpca_admm_set_smoother(ddpca_admm *h, int smoother_mode)
{
    if (!h || h->finalized || (smoother_mode != DDPCA_SMOOTH_LEX && smoother_mode != DDPCA_SMOOTH_MC)) return fail("ddpca_admm_set_smoother: bad argument");
    h->smoother_mode = smoother_mode;
    return 0;
}

// zero initial state (MCONTACT.h:875-894) again: the same handle then repeats the analysis
int ddpca_admm_reset(ddpca_admm *h)
{
    if (!h || !h->finalized) return fail("ddpca_admm_reset: handle not finalized");
    CU(cudaSetDevice(h->device));
    CU(cudaMemsetAsync(h->state, 0, sizeof(double) * (size_t)(2 * h->NC + h->NF + h->NG), h->stream));
    CU(cudaMemsetAsync(h->state_prev, 0, sizeof(double) * (size_t)(2 * h->NC + h->NF), h->stream));
    CU(cudaMemsetAsync(h->stat, 0, sizeof(int) * std::max(1, h->NG), h->stream));
    return 0;
}
// new load vector of body v (multGrid[v].consForc, host memory, n_L doubles): enqueued on the handle's stream
int ddpca_admm_set_consforc(ddpca_admm *h, int v, const double *consForc)
{
    if (!h || !h->finalized || v < 0 || v >= h->nb || !consForc) return fail("ddpca_admm_set_consforc: bad argument");
    if (!h->body[v].local) return fail("ddpca_admm_set_consforc: body belongs to another rank");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(h->consForc + h->body[v].roff, consForc, sizeof(double) * h->body[v].nred, cudaMemcpyHostToDevice, h->stream));
    return 0;
}

int ddpca_admm_row_length(const ddpca_admm *h) { return h ? 2 * h->nb + 8 * h->ni + 2 : -1; }

int ddpca_admm_get_disp(ddpca_admm *h, int v, double *out)
{
    if (!h || !h->finalized || v < 0 || v >= h->nb || !out) return fail("ddpca_admm_get_disp: bad argument");
    CU(cudaSetDevice(h->device));
    if (!h->body[v].local) return fail("ddpca_admm_get_disp: body belongs to another rank");
    // ordered after everything enqueued on the handle's stream (which is non-blocking w.r.t. the legacy stream)
    CU(cudaMemcpyAsync(out, h->state + h->offDisp() + h->body[v].foff, sizeof(double) * h->body[v].nfull, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}
int ddpca_admm_get_side(ddpca_admm *h, int ts, int tv, double *aux, double *lagr)
{
    if (!h || !h->finalized || ts < 0 || ts >= h->ni || tv < 0 || tv > 1) return fail("ddpca_admm_get_side: bad argument");
    CU(cudaSetDevice(h->device));
    HostSide &s = h->iface[ts].side[tv];
    if (!s.local) return fail("ddpca_admm_get_side: side belongs to another rank");
    if (aux) CU(cudaMemcpyAsync(aux, h->state + h->offAux() + s.coff, sizeof(double) * s.nc, cudaMemcpyDeviceToHost, h->stream));
    if (lagr) CU(cudaMemcpyAsync(lagr, h->state + h->offLagr() + s.coff, sizeof(double) * s.nc, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}
int ddpca_admm_get_gamma(ddpca_admm *h, int ts, double *gamma, int *fricStat)
{
    if (!h || !h->finalized || ts < 0 || ts >= h->ni) return fail("ddpca_admm_get_gamma: bad argument");
    CU(cudaSetDevice(h->device));
    HostIface &f = h->iface[ts];
    if (f.goff < 0) return fail("ddpca_admm_get_gamma: no side of this interface lives on this rank");
    if (gamma) CU(cudaMemcpyAsync(gamma, h->state + h->offGamma() + f.goff, sizeof(double) * f.d * f.nip, cudaMemcpyDeviceToHost, h->stream));
    if (fricStat) CU(cudaMemcpyAsync(fricStat, h->stat + f.goff, sizeof(int) * f.d * f.nip, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}
long ddpca_admm_launch_count(ddpca_admm *h, int reset)
{
    if (!h) return -1;
    long v = h->launches;
    if (reset) h->launches = 0;
    return v;
}
// kernel-level entry (parity tests): the projection kernel of the loop on host arrays
int ddpca_gamma_project(int device, int nip, int d, double fricCoef, const double *t, const double *gapTerm, double *inpoGamm, int *fricStat)
{
    if (nip < 0 || (d != 1 && d != 3) || !t || !gapTerm || !inpoGamm || !fricStat) return fail("ddpca_gamma_project: bad argument");
    int ndev = ddpca_device_count();
    if (ndev == 0) return fail("no CUDA device: libddpca_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail("device index out of range");
    CU(cudaSetDevice(device));
    const size_t n = (size_t)d * nip;
    double *dt = nullptr, *dg = nullptr, *dgam = nullptr;
    int *dst = nullptr;
    IfaceMeta *dm = nullptr;
    auto cleanup = [&]() { cudaFree(dt); cudaFree(dg); cudaFree(dgam); cudaFree(dst); cudaFree(dm); };
    IfaceMeta m{0, nip, d, 0, fricCoef};
    CUX(cudaMalloc(&dt, sizeof(double) * std::max<size_t>(1, n)));
    CUX(cudaMalloc(&dg, sizeof(double) * std::max<size_t>(1, n)));
    CUX(cudaMalloc(&dgam, sizeof(double) * std::max<size_t>(1, n)));
    CUX(cudaMalloc(&dst, sizeof(int) * std::max<size_t>(1, n)));
    CUX(cudaMalloc(&dm, sizeof(IfaceMeta)));
    CUX(cudaMemcpy(dt, t, sizeof(double) * n, cudaMemcpyHostToDevice));
    CUX(cudaMemcpy(dg, gapTerm, sizeof(double) * n, cudaMemcpyHostToDevice));
    CUX(cudaMemcpy(dm, &m, sizeof(IfaceMeta), cudaMemcpyHostToDevice));
    k_gamma_project_all<<<cdiv(std::max(nip, 1), 256), 256>>>(1, dm, dt, dg, dgam, dst);
    CUX(cudaMemcpy(inpoGamm, dgam, sizeof(double) * n, cudaMemcpyDeviceToHost));
    CUX(cudaMemcpy(fricStat, dst, sizeof(int) * n, cudaMemcpyDeviceToHost));
    CUX(cudaGetLastError());
    cleanup();
    return 0;
}

// per-kernel-class timing of the batched body solves (ddpca_mg_profile on every batch): enable, run steps, read
int ddpca_admm_profile(ddpca_admm *h, int enable)
{
    if (!h || !h->finalized) return fail("ddpca_admm_profile: handle not finalized");
    for (Batch &b : h->batch) if (ddpca_mg_profile(b.mg, enable)) return 1;
    return 0;
}
int ddpca_admm_profile_get(ddpca_admm *h, int kclass, int level, double *ms, long *launches, double *bytes)
{
    if (!h || !h->finalized) return fail("ddpca_admm_profile_get: handle not finalized");
    double m = 0.0, by = 0.0;
    long n = 0;
    for (Batch &b : h->batch) {
        double m1 = 0.0, b1 = 0.0;
        long n1 = 0;
        if (level < b.mg->nlev && ddpca_mg_profile_get(b.mg, kclass, level, &m1, &n1, &b1)) return 1;
        m += m1; n += n1; by += b1;
    }
    if (ms) *ms = m;
    if (launches) *launches = n;
    if (bytes) *bytes = by;
    return 0;
}
// number of batches and, for every body, its CG iteration count in the last step (0 for remote bodies)
int ddpca_admm_body_iters(const ddpca_admm *h, int *nbatches, long *iters)
{
    if (!h || !h->finalized) return fail("ddpca_admm_body_iters: handle not finalized");
    if (nbatches) *nbatches = (int)h->batch.size();
    if (iters) {
        for (int v = 0; v < h->nb; v++) iters[v] = 0;
        for (const Batch &b : h->batch)
            for (size_t s = 0; s < b.bodies.size(); s++) iters[b.bodies[s]] = (long)b.mg->st_host[s].it;
    }
    return 0;
}

}  // extern "C"
