// admm.inl -- device-resident ADMM iteration of MCONTACT::CONTACT_ANALYSIS (MCONTACT.h:2493-2723).
// Included at the end of mg.cu (one translation unit: the kernels of kernels.cuh are shared).
//
// State lives in HBM for the whole loop: resuDisp[v], inteAuxi[ts][tv],
// inteLagr[ts][tv] and their previous-iteration copies (MONITOR needs the differences).
// Operators (a13) are the reference's own matrices, uploaded once after MCONTACT::ESTABLISH.
// One ddpca_admm_step() = one pass of the loop body (:2505-2704) + the MONITOR norms
// (:2737-2833); the stopping logic itself (ring buffers, VECT_MEDI_OSCI, MULT_MAXI) is scalar
// host logic and stays with the caller (host mirror of MCONTACT).

namespace {

struct AdmmBody {
    ddpca_mg *mg = nullptr;
    int nfull = 0, nred = 0;
    double *consForc = nullptr, *dispCons = nullptr;
    DevCsr F, FT;          // forcOper (nred x nfull) and its transpose (OUTP_SUB1 = FT u + dispCons)
    DevCsr accuProl;       // coarse-space corrections only (muscSett bits 0 / 1)
    DevCsr globTran_D_1;   // muscSett bit 1 only (MCONTACT.h:2583)
    double *disp = nullptr, *disp_prev = nullptr, *addi = nullptr, *rhs = nullptr, *u = nullptr;
    bool set = false;
    bool local = true;     // owned by this rank (multi-GPU: one process per GPU)
    cudaEvent_t ev_done = nullptr;   // end of this body's solve on its own stream
};
struct AdmmSide {
    DevCsr op[DDPCA_OP_COUNT];
    DevCsr systTran_penaT;
    ddpca_ldlt *mass = nullptr, *mass_pena = nullptr;
    double *aux = nullptr, *lagr = nullptr, *aux_prev = nullptr, *lagr_prev = nullptr, *force = nullptr, *tmp = nullptr;
    int nc = 0;   // d * n_c
    bool local = true;     // side lives with its body
    double *trace = nullptr;   // inpoLagr*lambda + pemaInpo_r*u of this side, d*n_ip (private or inside the exchange buffer)
};
struct AdmmIface {
    int body[2] = {-1, -1};
    double fric = 0.0;
    int nip = 0, d = 1;
    double *gap = nullptr, *t = nullptr, *gamma = nullptr;
    int *stat = nullptr;
    AdmmSide side[2];
    bool set = false;
    bool cross = false;    // the two sides live on different ranks: traces go through the exchange buffer
    long trace_off[2] = {-1, -1};
};

}  // namespace

struct ddpca_admm : Engine {
    int nb = 0, ni = 0, muscSett = 0;
    std::vector<AdmmBody> body;
    std::vector<AdmmIface> iface;
    int nglob = 0;
    std::vector<long> baseReco;
    ddpca_ldlt *coar = nullptr;
    ddpca_mg *coar_mg = nullptr;   // macroscopic problem solved by MG-PCG instead (globCoup beyond DIRE_MAXI rows)
    long macro_cg_iters = 0;
    double *globForc = nullptr, *globSolu = nullptr;
    // interface-eliminated coarse problem (muscSett bit 1, MCONTACT.h:2575-2607)
    int nglob1 = 0;
    ddpca_ldlt *coar1 = nullptr;
    double *globForc1_const = nullptr, *globForc1 = nullptr, *globSolu1 = nullptr;
    double *moni_part = nullptr, *moni_out = nullptr;
    double *moni_host = nullptr;
    int nslots = 0;
    bool finalized = false;
    cudaEvent_t ev_fork = nullptr;
    // multi-rank: ownership + externally provided exchange buffers (device memory the caller all-reduces)
    std::vector<int> body_rank;
    int my_rank = 0;
    long trace_total = 0;             // doubles in the packed trace buffer (cross-rank interfaces only)
    double *x_glob = nullptr, *x_trace = nullptr, *x_moni = nullptr;   // external buffers (not owned) or null
    double *own_trace = nullptr;      // internal packed trace buffer when none is provided
    long cg_iters = 0;        // CG iterations of the last step, all bodies
    double cg_dof_iters = 0;  // sum over bodies of n_L * iterations, last step
};

static void admm_free(ddpca_admm *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    for (auto &b : h->body) {
        if (b.mg) ddpca_mg_destroy(b.mg);
        cudaFree(b.consForc); cudaFree(b.dispCons); free_csr(b.F); free_csr(b.FT); free_csr(b.accuProl); free_csr(b.globTran_D_1);
        cudaFree(b.disp); cudaFree(b.disp_prev); cudaFree(b.addi); cudaFree(b.rhs); cudaFree(b.u);
        if (b.ev_done) cudaEventDestroy(b.ev_done);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    for (auto &f : h->iface) {
        cudaFree(f.gap); cudaFree(f.t); cudaFree(f.gamma); cudaFree(f.stat);
        for (auto &s : f.side) {
            for (auto &o : s.op) free_csr(o);
            free_csr(s.systTran_penaT);
            ldlt_free(s.mass); ldlt_free(s.mass_pena);
            cudaFree(s.aux); cudaFree(s.lagr); cudaFree(s.aux_prev); cudaFree(s.lagr_prev); cudaFree(s.force); cudaFree(s.tmp);
            if (!f.cross) cudaFree(s.trace);
        }
    }
    ldlt_free(h->coar);
    if (h->coar_mg) ddpca_mg_destroy(h->coar_mg);
    ldlt_free(h->coar1);
    cudaFree(h->globForc1_const); cudaFree(h->globForc1); cudaFree(h->globSolu1);
    cudaFree(h->globForc); cudaFree(h->globSolu); cudaFree(h->moni_part); cudaFree(h->moni_out); cudaFree(h->own_trace);
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

static int dev_vec(const double *host, int n, double **d)
{
    CU(cudaMalloc(d, sizeof(double) * std::max(1, n)));
    if (host) CU(cudaMemcpy(*d, host, sizeof(double) * n, cudaMemcpyHostToDevice));
    else CU(cudaMemset(*d, 0, sizeof(double) * std::max(1, n)));
    return 0;
}

#define ADMM_SPMV(A, x, y, add, alpha) launch_spmv(h, DDPCA_K_VECTOR, 15, (A), (x), (y), (add), nullptr, nullptr, nullptr, (alpha))

// MONITOR's two sums for one state vector -> slot
static void admm_moni(ddpca_admm *h, int slot, int n, const double *cur, const double *prev)
{
    KL(h, DDPCA_K_VECTOR, 15, 16.0 * n, (k_moni_partial<<<kMoniBlocks, 256, 0, h->stream>>>(n, cur, prev, h->moni_part + (size_t)slot * 2 * kMoniBlocks)));
}

// ---- the loop body in phases; between phases a multi-rank caller all-reduces one buffer ---------
enum { PH_BODIES = 0, PH_MACRO_PARTIAL = 1, PH_MACRO_APPLY = 2, PH_TRACES = 3, PH_INTERFACE = 4, PH_MONITOR = 5, PH_MACRO1_PARTIAL = 6, PH_MACRO1_APPLY = 7 };

static double *glob_buf(ddpca_admm *h) { return h->x_glob ? h->x_glob : h->globForc; }
static double *glob_buf1(ddpca_admm *h) { return h->x_glob ? h->x_glob : h->globForc1; }   // the exchange buffer holds max(nglob, nglob1)
static double *moni_buf(ddpca_admm *h) { return h->x_moni ? h->x_moni : h->moni_out; }

// body balance, MCONTACT.h:2511-2538 (local bodies).  The reference runs this loop under
// `#pragma omp parallel for` (:2511); here every body is enqueued on its own stream (right-hand side,
// the whole MG-PCG solve as one graph launch, expansion to nodal displacements) so that small
// subdomains, which cannot fill the GPU alone, overlap; the ADMM stream joins them afterwards.
static int admm_bodies(ddpca_admm *h)
{
    cudaStream_t st = h->stream;
    h->cg_iters = 0;
    h->cg_dof_iters = 0;
    if (!h->ev_fork) CU(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CU(cudaEventRecord(h->ev_fork, st));
    const bool overlap = !h->profile && !std::getenv("DDPCA_SERIAL_BODIES");
    for (int v = 0; v < h->nb; v++) {
        AdmmBody &b = h->body[v];
        if (!b.local) continue;
        cudaStream_t bs = overlap ? b.mg->own_stream : st;
        if (overlap) {
            CU(cudaStreamWaitEvent(bs, h->ev_fork, 0));
            if (!b.ev_done) CU(cudaEventCreateWithFlags(&b.ev_done, cudaEventDisableTiming));
        }
        h->stream = bs;   // the helper kernels of this body go to its stream
        ddpca_mg_set_stream(b.mg, (void *)bs);
        int rc = 0;
        do {
            if (cudaMemcpyAsync(b.disp_prev, b.disp, sizeof(double) * b.nfull, cudaMemcpyDeviceToDevice, bs) != cudaSuccess) { rc = 1; break; }  // :2507
            if (cudaMemsetAsync(b.addi, 0, sizeof(double) * b.nfull, bs) != cudaSuccess) { rc = 1; break; }                                       // :2514
            for (int ts = 0; ts < h->ni; ts++)
                for (int ti = 0; ti < 2; ti++) {
                    AdmmIface &f = h->iface[ts];
                    if (f.body[ti] != v) continue;
                    ADMM_SPMV(f.side[ti].op[DDPCA_OP_SYSTTRAN_PENA], f.side[ti].aux, b.addi, true, 1.0);   // :2520
                    ADMM_SPMV(f.side[ti].op[DDPCA_OP_SYSTTRAN], f.side[ti].lagr, b.addi, true, -1.0);      // :2521
                }
            if (cudaMemcpyAsync(b.rhs, b.consForc, sizeof(double) * b.nred, cudaMemcpyDeviceToDevice, bs) != cudaSuccess) { rc = 1; break; }
            ADMM_SPMV(b.F, b.addi, b.rhs, true, 1.0);   // ADDITIONAL_FORCE :2524 ; consForc + addiForc :2531
            if (pcg_device(b.mg, 1, b.rhs, b.u, 1.0e-14, b.nred, nullptr, nullptr, nullptr, /*no_wait=*/true)) { rc = 1; break; }   // :2531 (MG-PCG for every body)
            if (cudaMemcpyAsync(b.disp, b.dispCons, sizeof(double) * b.nfull, cudaMemcpyDeviceToDevice, bs) != cudaSuccess) { rc = 1; break; }
            ADMM_SPMV(b.FT, b.u, b.disp, true, 1.0);    // OUTP_SUB1 :2533
            if (overlap && cudaEventRecord(b.ev_done, bs) != cudaSuccess) { rc = 1; break; }
        } while (0);
        h->stream = st;
        if (rc) return fail("ddpca_admm: enqueue of body " + std::to_string(v) + " failed: " + cudaGetErrorString(cudaGetLastError()) + " " + g_err);
    }
    for (int v = 0; v < h->nb; v++) {
        AdmmBody &b = h->body[v];
        if (!b.local) continue;
        long it = 0;
        if (pcg_finish(b.mg, &it, nullptr, nullptr)) return 1;
        if (overlap) CU(cudaStreamWaitEvent(st, b.ev_done, 0));
        h->launches += ddpca_mg_launch_count(b.mg, 1);
        h->cg_iters += it;
        h->cg_dof_iters += (double)it * b.nred;
    }
    return 0;
}

// macroscopic problem, :2541-2549: this rank's part of globForc (sum over its sides)
static int admm_macro_partial(ddpca_admm *h)
{
    if (!h->coar && !h->coar_mg) return fail("macroscopic problem requested but not set");
    double *gf = glob_buf(h);
    CU(cudaMemsetAsync(gf, 0, sizeof(double) * h->nglob, h->stream));
    for (int ts = 0; ts < h->ni; ts++)
        for (int tv = 0; tv < 2; tv++) {
            AdmmIface &f = h->iface[ts];
            AdmmSide &s = f.side[tv];
            if (!s.local) continue;
            ADMM_SPMV(s.op[DDPCA_OP_GLOBTRAN], s.lagr, gf, true, 1.0);                       // :2545
            ADMM_SPMV(s.op[DDPCA_OP_GLOBTRAN_PENA], s.aux, gf, true, -1.0);                  // :2546
            ADMM_SPMV(s.op[DDPCA_OP_GLOBTRAN_D], h->body[f.body[tv]].disp, gf, true, 1.0);   // :2547
        }
    return 0;
}
// resuDisp[v] += OUTP_SUB1(accuProl[v] * globSolu[baseReco[v] ...]) for the local bodies (:2564-2570, :2599-2604)
static int admm_coarse_correction(ddpca_admm *h, const double *globSolu)
{
    cudaStream_t st = h->stream;
    for (int v = 0; v < h->nb; v++) {
        AdmmBody &b = h->body[v];
        if (!b.local) continue;
        ADMM_SPMV(b.accuProl, globSolu + h->baseReco[v], b.u, false, 1.0);      // :2564-2567
        ADMM_SPMV(b.FT, b.u, b.disp, true, 1.0);                                 // :2569-2570 (OUTP_SUB1 ...
        KL(h, DDPCA_K_VECTOR, 15, 24.0 * b.nfull, (k_axpy<<<cdiv(b.nfull, 256), 256, 0, st>>>(b.nfull, 1.0, b.dispCons, b.disp)));  // ... re-adds prescribed values)
    }
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
    if (!h->coar1) return fail("interface-eliminated coarse problem requested but not set");
    double *gf = glob_buf1(h);
    CU(cudaMemsetAsync(gf, 0, sizeof(double) * h->nglob1, h->stream));
    for (int ts = 0; ts < h->ni; ts++)
        for (int tv = 0; tv < 2; tv++) {
            AdmmSide &s = h->iface[ts].side[tv];
            if (!s.local) continue;
            ADMM_SPMV(s.op[DDPCA_OP_GLOBTRAN_1], s.lagr, gf, true, 1.0);   // :2579
        }
    for (int v = 0; v < h->nb; v++) {
        AdmmBody &b = h->body[v];
        if (!b.local) continue;
        ADMM_SPMV(b.globTran_D_1, b.disp, gf, true, -1.0);                 // :2583
    }
    return 0;
}
// :2576 (constant part, added once after the sum over ranks), :2585-2606
static int admm_macro1_apply(ddpca_admm *h)
{
    cudaStream_t st = h->stream;
    double *gf = glob_buf1(h);
    KL(h, DDPCA_K_VECTOR, 15, 24.0 * h->nglob1, (k_axpy<<<cdiv(h->nglob1, 256), 256, 0, st>>>(h->nglob1, 1.0, h->globForc1_const, gf)));
    ldlt_solve_on(h, h->coar1, gf, h->globSolu1, nullptr);   // :2588
    return admm_coarse_correction(h, h->globSolu1);
}
// side traces inpoLagr*lambda + pemaInpo_r*u (:2632-2635); remote sides of cross-rank interfaces are zero-filled
static int admm_traces(ddpca_admm *h)
{
    // slots of interfaces this rank does not touch must contribute zero to the all-reduce
    if (h->trace_total) CU(cudaMemsetAsync(h->x_trace ? h->x_trace : h->own_trace, 0, sizeof(double) * h->trace_total, h->stream));
    for (int ts = 0; ts < h->ni; ts++) {
        AdmmIface &f = h->iface[ts];
        if (!f.side[0].local && !f.side[1].local) continue;
        int ng = f.d * f.nip;
        for (int tv = 0; tv < 2; tv++) {
            AdmmSide &s = f.side[tv];
            if (s.local) {
                ADMM_SPMV(s.op[DDPCA_OP_INPOLAGR], s.lagr, s.trace, false, 1.0);
                ADMM_SPMV(s.op[DDPCA_OP_PEMAINPO_R], h->body[f.body[tv]].disp, s.trace, true, 1.0);
            } else {
                CU(cudaMemsetAsync(s.trace, 0, sizeof(double) * ng, h->stream));
            }
        }
    }
    return 0;
}
// interface balance :2636-2685 and multiplier update :2689-2704 for the local sides
static int admm_interface(ddpca_admm *h)
{
    cudaStream_t st = h->stream;
    for (int ts = 0; ts < h->ni; ts++) {
        AdmmIface &f = h->iface[ts];
        if (!f.side[0].local && !f.side[1].local) continue;
        int ng = f.d * f.nip;
        CU(cudaMemcpyAsync(f.t, f.side[0].trace, sizeof(double) * ng, cudaMemcpyDeviceToDevice, st));
        KL(h, DDPCA_K_VECTOR, 15, 24.0 * ng, (k_axpy<<<cdiv(ng, 256), 256, 0, st>>>(ng, -1.0, f.side[1].trace, f.t)));   // :2632-2635
        KL(h, DDPCA_K_VECTOR, 15, 28.0 * ng, (k_gamma_project<<<cdiv(f.nip, 256), 256, 0, st>>>(f.nip, f.d, f.fric, f.t, f.gap, f.gamma, f.stat)));  // :2636-2668
        for (int tv = 0; tv < 2; tv++) {
            AdmmSide &s = f.side[tv];
            if (!s.local) continue;
            ADMM_SPMV(s.systTran_penaT, h->body[f.body[tv]].disp, s.force, false, 1.0);   // :2673
            ADMM_SPMV(s.op[DDPCA_OP_INTEMASS], s.lagr, s.force, true, 1.0);               // :2674
            ADMM_SPMV(s.op[DDPCA_OP_INTEINPO], f.gamma, s.force, true, 1.0);              // :2675
            CU(cudaMemcpyAsync(s.aux_prev, s.aux, sizeof(double) * s.nc, cudaMemcpyDeviceToDevice, st));   // :2508
            ldlt_solve_on(h, s.mass_pena, s.force, s.aux, nullptr);                       // :2677
        }
    }
    for (int ts = 0; ts < h->ni; ts++) {
        AdmmIface &f = h->iface[ts];
        for (int tv = 0; tv < 2; tv++) {
            AdmmSide &s = f.side[tv];
            if (!s.local) continue;
            ADMM_SPMV(s.systTran_penaT, h->body[f.body[tv]].disp, s.force, false, 1.0);   // :2693
            ADMM_SPMV(s.op[DDPCA_OP_INTEMASS_PENA], s.aux, s.force, true, -1.0);          // :2694
            ldlt_solve_on(h, s.mass, s.force, s.tmp, nullptr);                            // :2696
            CU(cudaMemcpyAsync(s.lagr_prev, s.lagr, sizeof(double) * s.nc, cudaMemcpyDeviceToDevice, st));   // :2509
            KL(h, DDPCA_K_VECTOR, 15, 24.0 * s.nc, (k_axpy<<<cdiv(s.nc, 256), 256, 0, st>>>(s.nc, 1.0, s.tmp, s.lagr)));
        }
    }
    return 0;
}
// MONITOR sums, :2737-2833: slots of remote bodies / sides stay zero (the caller all-reduces)
static int admm_monitor(ddpca_admm *h)
{
    CU(cudaMemsetAsync(h->moni_part, 0, sizeof(double) * 2 * kMoniBlocks * h->nslots, h->stream));
    for (int v = 0; v < h->nb; v++)
        if (h->body[v].local) admm_moni(h, v, h->body[v].nfull, h->body[v].disp, h->body[v].disp_prev);
    for (int ts = 0; ts < h->ni; ts++)
        for (int tv = 0; tv < 2; tv++) {
            AdmmSide &s = h->iface[ts].side[tv];
            if (!s.local) continue;
            int slot = h->nb + 4 * ts + 2 * tv;   // tempIndi of :2771
            admm_moni(h, slot, s.nc, s.aux, s.aux_prev);
            admm_moni(h, slot + 1, s.nc, s.lagr, s.lagr_prev);
        }
    KL(h, DDPCA_K_VECTOR, 15, 0.0, (k_moni_final<<<h->nslots, 32, 0, h->stream>>>(h->moni_part, moni_buf(h))));
    return 0;
}
// row of resuMoni.txt (:2742-2743, :2777-2778, :2807-2808, :2835) from the (all-reduced) sums
static int admm_row(ddpca_admm *h, double *monitor_row)
{
    CU(cudaMemcpyAsync(h->moni_host, moni_buf(h), sizeof(double) * 2 * h->nslots, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    if (h->profile) h->prof_collect();
    if (!monitor_row) return 0;
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
            int slot = h->nb + 4 * ts + 2 * tv;
            monitor_row[c++] = h->moni_host[2 * slot];
            monitor_row[c++] = h->moni_host[2 * slot + 1];
            convValu += h->moni_host[2 * slot];
            convCrit += h->moni_host[2 * slot + 1];
            monitor_row[c++] = h->moni_host[2 * (slot + 1)];
            monitor_row[c++] = h->moni_host[2 * (slot + 1) + 1];
        }
    monitor_row[c++] = convValu;
    monitor_row[c++] = convCrit;
    return 0;
}

static int admm_step(ddpca_admm *h, int apply_macro, double *monitor_row)
{
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

int ddpca_admm_set_body(ddpca_admm *h, int v, ddpca_mg *mg, int nfull, const double *consForc, const int *F_rowptr,
                        const int *F_colidx, const double *F_val, const double *dispCons)
{
    if (!h || v < 0 || v >= h->nb || !mg || nfull < 1 || !consForc || !F_rowptr || !dispCons) return fail("ddpca_admm_set_body: bad argument");
    if (mg->device != h->device) return fail("ddpca_admm_set_body: hierarchy lives on another device");
    CU(cudaSetDevice(h->device));
    AdmmBody &b = h->body[v];
    if (b.set) return fail("ddpca_admm_set_body: body already set");
    if (!b.local) return fail("ddpca_admm_set_body: body " + std::to_string(v) + " belongs to another rank");
    b.mg = mg;
    b.nfull = nfull;
    b.nred = mg->lev[mg->nlev - 1].n;
    CsrHost F, FT;
    if (host_csr(b.nred, nfull, F_rowptr, F_colidx, F_val, F)) return 1;
    transpose_csr(F, FT);
    if (upload_csr(F, b.F) || upload_csr(FT, b.FT)) return 1;
    if (dev_vec(consForc, b.nred, &b.consForc) || dev_vec(dispCons, nfull, &b.dispCons)) return 1;
    if (dev_vec(nullptr, nfull, &b.disp) || dev_vec(nullptr, nfull, &b.disp_prev) || dev_vec(nullptr, nfull, &b.addi)) return 1;
    if (dev_vec(nullptr, b.nred, &b.rhs) || dev_vec(nullptr, b.nred, &b.u)) return 1;
    b.set = true;
    return 0;
}

int ddpca_admm_set_body_accuprol(ddpca_admm *h, int v, int rows, int cols, const int *rowptr, const int *colidx, const double *val)
{
    if (!h || v < 0 || v >= h->nb || !h->body[v].set) return fail("ddpca_admm_set_body_accuprol: bad argument");
    if (rows != h->body[v].nred) return fail("accuProl must have n_L rows");
    CU(cudaSetDevice(h->device));
    CsrHost A;
    if (host_csr(rows, cols, rowptr, colidx, val, A)) return 1;
    return upload_csr(A, h->body[v].accuProl);
}

int ddpca_admm_set_interface(ddpca_admm *h, int ts, int body0, int body1, double fricCoef, int nip, const double *gapTerm)
{
    if (!h || ts < 0 || ts >= h->ni || body0 < 0 || body0 >= h->nb || body1 < 0 || body1 >= h->nb || nip < 0 || !gapTerm)
        return fail("ddpca_admm_set_interface: bad argument");
    CU(cudaSetDevice(h->device));
    AdmmIface &f = h->iface[ts];
    if (f.set) return fail("interface already set");
    f.body[0] = body0; f.body[1] = body1;
    f.side[0].local = h->body[body0].local;
    f.side[1].local = h->body[body1].local;
    f.cross = !h->body_rank.empty() && h->body_rank[body0] != h->body_rank[body1];
    f.fric = fricCoef;
    f.nip = nip;
    f.d = (fricCoef == 0.0) ? 1 : 3;   // MCONTACT.h:886-893
    int ng = f.d * nip;
    if (dev_vec(gapTerm, ng, &f.gap) || dev_vec(nullptr, ng, &f.t) || dev_vec(nullptr, ng, &f.gamma)) return 1;
    CU(cudaMalloc(&f.stat, sizeof(int) * std::max(1, ng)));
    CU(cudaMemset(f.stat, 0, sizeof(int) * std::max(1, ng)));
    f.set = true;
    return 0;
}

int ddpca_admm_set_side_op(ddpca_admm *h, int ts, int tv, int op, int rows, int cols, const int *rowptr, const int *colidx, const double *val)
{
    if (!h || ts < 0 || ts >= h->ni || tv < 0 || tv > 1 || op < 0 || op >= DDPCA_OP_COUNT || !h->iface[ts].set) return fail("ddpca_admm_set_side_op: bad argument");
    CU(cudaSetDevice(h->device));
    AdmmSide &s = h->iface[ts].side[tv];
    if (!s.local) return fail("ddpca_admm_set_side_op: this side belongs to another rank");
    CsrHost A;
    if (host_csr(rows, cols, rowptr, colidx, val, A)) return 1;
    if (op == DDPCA_OP_INTEMASS) s.nc = rows;
    if (upload_csr(A, s.op[op])) return 1;
    if (op == DDPCA_OP_SYSTTRAN_PENA) {
        CsrHost T;
        transpose_csr(A, T);
        if (upload_csr(T, s.systTran_penaT)) return 1;
    }
    return 0;
}

int ddpca_admm_set_side_solver(ddpca_admm *h, int ts, int tv, int which, ddpca_ldlt *sol)
{
    if (!h || ts < 0 || ts >= h->ni || tv < 0 || tv > 1 || !sol || (which != DDPCA_SOLVER_MASS && which != DDPCA_SOLVER_MASS_PENA)) return fail("ddpca_admm_set_side_solver: bad argument");
    if (sol->device != h->device) return fail("solver lives on another device");
    if (!h->iface[ts].set) return fail("ddpca_admm_set_side_solver: interface not set");
    AdmmSide &s = h->iface[ts].side[tv];
    if (!s.local) return fail("ddpca_admm_set_side_solver: this side belongs to another rank");
    if (which == DDPCA_SOLVER_MASS) { ldlt_free(s.mass); s.mass = sol; }
    else { ldlt_free(s.mass_pena); s.mass_pena = sol; }
    return 0;
}

int ddpca_admm_set_macro(ddpca_admm *h, int nglob, const long *baseReco, ddpca_ldlt *coarSolv)
{
    if (!h || nglob < 1 || !baseReco || !coarSolv) return fail("ddpca_admm_set_macro: bad argument");
    if (coarSolv->n != nglob) return fail("coarse solver size does not match globCoup");
    CU(cudaSetDevice(h->device));
    h->nglob = nglob;
    h->baseReco.assign(baseReco, baseReco + h->nb + 1);
    ldlt_free(h->coar);
    h->coar = coarSolv;
    if (dev_vec(nullptr, nglob, &h->globForc) || dev_vec(nullptr, nglob, &h->globSolu)) return 1;
    return 0;
}

int ddpca_admm_set_macro_mg(ddpca_admm *h, int nglob, const long *baseReco, ddpca_mg *mgpi)
{
    if (!h || nglob < 1 || !baseReco || !mgpi) return fail("ddpca_admm_set_macro_mg: bad argument");
    if (mgpi->lev[mgpi->nlev - 1].n != nglob) return fail("finest level of the macroscopic hierarchy does not match globCoup");
    if (mgpi->device != h->device) return fail("macroscopic hierarchy lives on another device");
    CU(cudaSetDevice(h->device));
    h->nglob = nglob;
    h->baseReco.assign(baseReco, baseReco + h->nb + 1);
    ldlt_free(h->coar);
    h->coar = nullptr;
    if (h->coar_mg && h->coar_mg != mgpi) ddpca_mg_destroy(h->coar_mg);
    h->coar_mg = mgpi;
    if (dev_vec(nullptr, nglob, &h->globForc) || dev_vec(nullptr, nglob, &h->globSolu)) return 1;
    return 0;
}

int ddpca_admm_set_body_globtran_d1(ddpca_admm *h, int v, int rows, int cols, const int *rowptr, const int *colidx, const double *val)
{
    if (!h || v < 0 || v >= h->nb || !h->body[v].set) return fail("ddpca_admm_set_body_globtran_d1: bad argument");
    if (cols != h->body[v].nfull) return fail("globTran_D_1 must have 3 n_nodes columns");
    CU(cudaSetDevice(h->device));
    CsrHost A;
    if (host_csr(rows, cols, rowptr, colidx, val, A)) return 1;
    return upload_csr(A, h->body[v].globTran_D_1);
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
    if (dev_vec(globForc_1, nglob1, &h->globForc1_const) || dev_vec(nullptr, nglob1, &h->globForc1) || dev_vec(nullptr, nglob1, &h->globSolu1)) return 1;
    return 0;
}

int ddpca_admm_finalize(ddpca_admm *h)
{
    if (!h) return fail("null handle");
    CU(cudaSetDevice(h->device));
    for (int v = 0; v < h->nb; v++) {
        if (!h->body[v].local) continue;
        if (!h->body[v].set) return fail("body " + std::to_string(v) + " not set");
        if ((h->muscSett & 3) && h->body[v].accuProl.rows == 0) return fail("body " + std::to_string(v) + ": accuProl missing");
        if ((h->muscSett & 1) && h->baseReco.size() == (size_t)h->nb + 1 && h->baseReco[v] + h->body[v].accuProl.cols > h->nglob) return fail("baseReco out of range");
        if ((h->muscSett & 2) && h->baseReco.size() == (size_t)h->nb + 1 && h->baseReco[v] + h->body[v].accuProl.cols > h->nglob1) return fail("baseReco out of range (globCoup_1)");
        if ((h->muscSett & 2) && (h->body[v].globTran_D_1.rp == nullptr || h->body[v].globTran_D_1.rows != h->nglob1)) return fail("body " + std::to_string(v) + ": globTran_D_1 missing");
    }
    if ((h->muscSett & 1) && !h->coar && !h->coar_mg) return fail("macroscopic problem not set");
    if ((h->muscSett & 2) && !h->coar1) return fail("interface-eliminated coarse problem not set");
    for (int ts = 0; ts < h->ni; ts++) {
        AdmmIface &f = h->iface[ts];
        if (!f.set) return fail("interface " + std::to_string(ts) + " not set");
        for (int tv = 0; tv < 2; tv++) {
            AdmmSide &s = f.side[tv];
            if (!s.local) continue;
            int need[] = {DDPCA_OP_SYSTTRAN, DDPCA_OP_SYSTTRAN_PENA, DDPCA_OP_INTEMASS, DDPCA_OP_INTEMASS_PENA, DDPCA_OP_INPOLAGR, DDPCA_OP_INTEINPO, DDPCA_OP_PEMAINPO_R};
            for (int o : need) if (s.op[o].rows == 0 && s.op[o].rp == nullptr) return fail("interface " + std::to_string(ts) + " side " + std::to_string(tv) + ": operator " + std::to_string(o) + " missing");
            if (h->muscSett & 1)
                for (int o : {DDPCA_OP_GLOBTRAN, DDPCA_OP_GLOBTRAN_PENA, DDPCA_OP_GLOBTRAN_D}) if (s.op[o].rp == nullptr) return fail("macroscopic transfer operator missing");
            if ((h->muscSett & 2) && s.op[DDPCA_OP_GLOBTRAN_1].rp == nullptr) return fail("globTran_1 missing");
            if (!s.mass || !s.mass_pena) return fail("interface " + std::to_string(ts) + " side " + std::to_string(tv) + ": mass solvers missing");
            int ng = f.d * f.nip, nfull = h->body[f.body[tv]].nfull;
            if (s.op[DDPCA_OP_SYSTTRAN].rows != nfull || s.op[DDPCA_OP_SYSTTRAN].cols != s.nc || s.op[DDPCA_OP_INPOLAGR].rows != ng ||
                s.op[DDPCA_OP_INTEINPO].rows != s.nc || s.op[DDPCA_OP_INTEINPO].cols != ng || s.op[DDPCA_OP_PEMAINPO_R].rows != ng ||
                s.op[DDPCA_OP_PEMAINPO_R].cols != nfull || s.mass->n != s.nc || s.mass_pena->n != s.nc)
                return fail("interface " + std::to_string(ts) + " side " + std::to_string(tv) + ": operator shapes are inconsistent");
            if (!s.aux) {
                // zero initial state, MCONTACT.h:875-894
                if (dev_vec(nullptr, s.nc, &s.aux) || dev_vec(nullptr, s.nc, &s.lagr) || dev_vec(nullptr, s.nc, &s.aux_prev) ||
                    dev_vec(nullptr, s.nc, &s.lagr_prev) || dev_vec(nullptr, s.nc, &s.force) || dev_vec(nullptr, s.nc, &s.tmp)) return 1;
            }
        }
    }
    // trace buffers: cross-rank interfaces live in one packed buffer with the same layout on every rank
    h->trace_total = 0;
    for (int ts = 0; ts < h->ni; ts++) {
        AdmmIface &f = h->iface[ts];
        if (!f.cross) continue;
        for (int tv = 0; tv < 2; tv++) { f.trace_off[tv] = h->trace_total; h->trace_total += (long)f.d * f.nip; }
    }
    if (h->trace_total && !h->x_trace && !h->own_trace) { if (dev_vec(nullptr, (int)h->trace_total, &h->own_trace)) return 1; }
    for (int ts = 0; ts < h->ni; ts++) {
        AdmmIface &f = h->iface[ts];
        if (!f.side[0].local && !f.side[1].local) continue;
        for (int tv = 0; tv < 2; tv++) {
            AdmmSide &s = f.side[tv];
            if (f.cross) s.trace = (h->x_trace ? h->x_trace : h->own_trace) + f.trace_off[tv];
            else if (!s.trace) { if (dev_vec(nullptr, f.d * f.nip, &s.trace)) return 1; }
        }
    }
    h->nslots = h->nb + 4 * h->ni;
    if (!h->moni_part) {
        CU(cudaMalloc(&h->moni_part, sizeof(double) * 2 * kMoniBlocks * h->nslots));
        CU(cudaMalloc(&h->moni_out, sizeof(double) * 2 * h->nslots));
        CU(cudaMallocHost(&h->moni_host, sizeof(double) * 2 * h->nslots));
    }
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
    long nt = 0;
    for (int ts = 0; ts < h->ni; ts++) if (h->iface[ts].cross) nt += 2L * h->iface[ts].d * h->iface[ts].nip;
    if (nglob) *nglob = std::max(h->nglob, h->nglob1);
    if (ntrace) *ntrace = nt;
    if (nmoni) *nmoni = 2L * (h->nb + 4 * h->ni);
    return 0;
}
int ddpca_admm_set_exchange(ddpca_admm *h, double *globForc_dev, double *traces_dev, double *moni_dev)
{
    if (!h) return fail("null handle");
    if (h->finalized) return fail("ddpca_admm_set_exchange must precede ddpca_admm_finalize");
    h->x_glob = globForc_dev; h->x_trace = traces_dev; h->x_moni = moni_dev;
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

int ddpca_admm_row_length(const ddpca_admm *h) { return h ? 2 * h->nb + 8 * h->ni + 2 : -1; }

int ddpca_admm_get_disp(ddpca_admm *h, int v, double *out)
{
    if (!h || v < 0 || v >= h->nb || !out) return fail("ddpca_admm_get_disp: bad argument");
    CU(cudaSetDevice(h->device));
    if (!h->body[v].local) return fail("ddpca_admm_get_disp: body belongs to another rank");
    // ordered after everything enqueued on the handle's stream (which is non-blocking w.r.t. the legacy stream)
    CU(cudaMemcpyAsync(out, h->body[v].disp, sizeof(double) * h->body[v].nfull, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}
int ddpca_admm_get_side(ddpca_admm *h, int ts, int tv, double *aux, double *lagr)
{
    if (!h || ts < 0 || ts >= h->ni || tv < 0 || tv > 1) return fail("ddpca_admm_get_side: bad argument");
    CU(cudaSetDevice(h->device));
    AdmmSide &s = h->iface[ts].side[tv];
    if (!s.local) return fail("ddpca_admm_get_side: side belongs to another rank");
    if (aux) CU(cudaMemcpyAsync(aux, s.aux, sizeof(double) * s.nc, cudaMemcpyDeviceToHost, h->stream));
    if (lagr) CU(cudaMemcpyAsync(lagr, s.lagr, sizeof(double) * s.nc, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}
int ddpca_admm_get_gamma(ddpca_admm *h, int ts, double *gamma, int *fricStat)
{
    if (!h || ts < 0 || ts >= h->ni) return fail("ddpca_admm_get_gamma: bad argument");
    CU(cudaSetDevice(h->device));
    AdmmIface &f = h->iface[ts];
    if (gamma) CU(cudaMemcpyAsync(gamma, f.gamma, sizeof(double) * f.d * f.nip, cudaMemcpyDeviceToHost, h->stream));
    if (fricStat) CU(cudaMemcpyAsync(fricStat, f.stat, sizeof(int) * f.d * f.nip, cudaMemcpyDeviceToHost, h->stream));
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

}  // extern "C"
