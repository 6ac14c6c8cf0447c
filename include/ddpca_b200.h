/*
 * ddpca_b200.h -- C ABI of libddpca_b200.so: the B200 (sm_100a, FP64) hot path of
 * DDPCA-ADMM behind plain pointers and sizes.
 *
 * The reference has no FFI today: the path is header-inline C++ (class MGPIS,
 * MGPIS.h:8-38; MCONTACT::CONTACT_ANALYSIS, MCONTACT.h:2493-2723).  Each entry
 * point below names the reference interface it replaces; the C++ overlay in
 * ddpca-admm_b200/host/ forwards the reference's own class surface to these
 * calls.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message is
 *     available from ddpca_last_error() (thread-local).  There is NO CPU
 *     fallback: without a CUDA device every compute call fails.
 *   - sparse operators are Eigen RowMajor compressed storage
 *     (SparseMatrix<double,RowMajor>::outerIndexPtr/innerIndexPtr/valuePtr):
 *     int32 rowptr[rows+1], int32 colidx[nnz] sorted inside a row, double val[nnz].
 *   - pointers are HOST pointers unless the function name ends in _dev.
 *   - a handle is bound to one device and owns one non-blocking stream; calls
 *     on different handles may run concurrently from different host threads
 *     (the reference calls CG_SOLV from inside an OpenMP loop, MCONTACT.h:2511-2532).
 */
#ifndef DDPCA_B200_H
#define DDPCA_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define DDPCA_ABI_VERSION 2

/* Smoother ordering of the multigrid V-cycle.
 * LEX: the reference's lexicographic symmetric Gauss-Seidel (MGPIS.h:65-77),
 *      reproduced exactly by wavefront scheduling -- parity / debug mode.
 * MC : the same symmetric Gauss-Seidel algebra applied after a multicolour
 *      symmetric permutation of each level -- throughput mode; same fixed
 *      point, different (still SPD) preconditioner. */
enum { DDPCA_SMOOTH_LEX = 0, DDPCA_SMOOTH_MC = 1 };

/* kernel classes, for ddpca_mg_profile() / ddpca_mg_bench_kernel() */
enum {
    DDPCA_K_SPMV = 0,      /* K1  y = A x                       MGPIS.h:200        */
    DDPCA_K_SWEEP_FWD = 1, /* K3  forward  Gauss-Seidel sweep   MGPIS.h:66-72      */
    DDPCA_K_SWEEP_BWD = 2, /* K4  backward Gauss-Seidel sweep   MGPIS.h:73-76      */
    DDPCA_K_RESID = 3,     /* K2  r = b - (p1 + L x)            MGPIS.h:92         */
    DDPCA_K_RESTRICT = 4,  /* K5  r_c = P^T r                   MGPIS.h:96         */
    DDPCA_K_PROLONG = 5,   /* K6  x += P e                      MGPIS.h:100        */
    DDPCA_K_COARSE = 6,    /* K7  x0 = A0^-1 b0                 MGPIS.h:58         */
    DDPCA_K_VECTOR = 7,    /* K8  axpy / dot / norm             MGPIS.h:197-214    */
    DDPCA_K_SWEEP_FWD0 = 8,/* K3  forward sweep from x = 0 (lower half only)   MGPIS.h:93,204 + :66-72 */
    DDPCA_K_COUNT = 9
};

typedef struct ddpca_mg ddpca_mg;
typedef struct ddpca_plan ddpca_plan;

const char *ddpca_last_error(void);
int ddpca_abi_version(void);
/* number of visible CUDA devices (0 when there is none; never fails) */
int ddpca_device_count(void);

/* ---- host-side planning (no GPU needed) ------------------------------------
 * The per-level reordering the device uses: row groups (rows of one mesh node
 * share a column pattern), stages (sets of mutually independent groups) and the
 * symmetric permutation that makes every stage a contiguous row range.
 * Exposed so that the ordering can be inspected and tested on a CPU-only box. */
int ddpca_plan_create(int n, const int *rowptr, const int *colidx, int smoother_mode, ddpca_plan **out);
/* the same for a block-diagonal level (a batch of subdomains, blocks = row ranges [sub_off[s], sub_off[s+1])): the
 * blocks are planned in parallel and merged; the result is identical to ddpca_plan_create on the whole level */
int ddpca_plan_create_blocks(int n, const int *rowptr, const int *colidx, int smoother_mode, int nsub, const int *sub_off, ddpca_plan **out);
/* wavefront plan of the unit triangular factors I + L and I + L^T of a sparse LDL^T factorisation (L strictly lower,
 * CSR), as ddpca_ldlt_create uses it; equal to ddpca_plan_create(LEX) on either operator */
int ddpca_plan_create_tri(int n, const int *L_rowptr, const int *L_colidx, ddpca_plan **out);
int ddpca_plan_sizes(const ddpca_plan *, int *n, int *ngroups, int *nstages);
/* perm[new] = old ; group_start[ngroups+1] and stage_start[nstages+1] are in NEW row numbering */
int ddpca_plan_get(const ddpca_plan *, int *perm, int *group_start, int *stage_start);
int ddpca_plan_destroy(ddpca_plan *);

/* ---- MGPIS: multigrid hierarchy + solvers ----------------------------------
 * ddpca_mg_create replaces MGPIS::ESTABLISH (MGPIS.h:40-53) plus the
 * direSolv.compute(consStif[0]) that CG_SOLV repeats on every call (MGPIS.h:185):
 * it uploads consStif[0..nlevels-1] and realProl[0..nlevels-2], builds the
 * diagonal split, the stage schedule and the level-0 direct solver on `device`. */
int ddpca_mg_create(int device, int nlevels, const int *n,
                    const int *const *rowptr, const int *const *colidx, const double *const *val,
                    const int *const *P_rowptr, const int *const *P_colidx, const double *const *P_val,
                    int smoother_mode, ddpca_mg **out);
/* The same for nsub independent subdomain hierarchies of `nlevels` levels each, held as ONE block-diagonal
 * hierarchy: every level kernel (smoother sweeps, residual, transfers, product, level-0 solve) works on all
 * subdomains at once, the CG recurrence of MGPIS::CG_SOLV stays per subdomain (own alpha, beta, tolerance,
 * iteration count; a converged subdomain freezes).  This is how the body loop of MCONTACT::CONTACT_ANALYSIS
 * (`#pragma omp parallel for` over multGrid[tv].mgpi.CG_SOLV, MCONTACT.h:2511-2532) maps onto one GPU: small
 * subdomains share launches instead of queueing.  Arrays are indexed [s * nlevels + l] (prolongations
 * [s * (nlevels-1) + l]); vectors of ddpca_mg_pcg* are the subdomains' vectors one after the other. */
int ddpca_mg_create_batch(int device, int nsub, int nlevels, const int *n,
                          const int *const *rowptr, const int *const *colidx, const double *const *val,
                          const int *const *P_rowptr, const int *const *P_colidx, const double *const *P_val,
                          int smoother_mode, ddpca_mg **out);
/* The host half of ddpca_mg_create_batch without touching a device (no GPU needed): the sweep plans, the
 * permuted operators, the kernel layouts, chunk tables and transfer forms are built exactly as for a real
 * handle and dropped.  A planning / diagnostic aid: device_bytes = what the hierarchy would occupy in HBM,
 * v2_levels = levels that qualify for the TMA-staged layout, seconds[9] = host time per set-up stage
 * (concatenate, plan, permute, layout, 0, transfers: permute, transpose, triples, 0), checksum = a hash
 * of every array that would be uploaded (regression anchor for the host code).  Outputs may be NULL. */
int ddpca_mg_setup_dryrun(int nsub, int nlevels, const int *n,
                          const int *const *rowptr, const int *const *colidx, const double *const *val,
                          const int *const *P_rowptr, const int *const *P_colidx, const double *const *P_val,
                          int smoother_mode, double *seconds, long *device_bytes, int *v2_levels,
                          unsigned long long *checksum);
/* per-subdomain results of the last ddpca_mg_pcg* call: iters[nsub], resid[nsub], tol_abs[nsub] (any may be null) */
int ddpca_mg_batch_result(const ddpca_mg *, int *nsub, long *iters, double *resid, double *tol_abs);
int ddpca_mg_destroy(ddpca_mg *);

/* MGPIS::CG_SOLV(precSwit, totaForc, resuSolu), MGPIS.h:163-225.
 * prec 0 = Jacobi (DIAG_PREC, PREP.h:393-401), 1 = one V-cycle.  x0 = 0,
 * stop when it >= maxit or ||r||_2 <= rel_tol*||b||_2 (reference: 1e-14, n; maxit <= 0 means n, per subdomain).
 * iters = the reference's iterNumb (it prints iterNumb-1); for a batch the largest count and residual, see
 * ddpca_mg_batch_result.  b, x: host, length n_L (sum over the subdomains of a batch). */
int ddpca_mg_pcg(ddpca_mg *, int prec, const double *b, double *x, double rel_tol, long maxit,
                 long *iters, double *resid, double *tol_abs);
/* same, operands resident in HBM on the handle's device (reference numbering) */
int ddpca_mg_pcg_dev(ddpca_mg *, int prec, const double *b_dev, double *x_dev, double rel_tol, long maxit,
                     long *iters, double *resid, double *tol_abs);

/* MGPIS::MULT_VCYC(level, righHand, resuSolu, direSolv), MGPIS.h:55-128; x is in/out */
int ddpca_mg_vcycle(ddpca_mg *, int level, const double *b, double *x);
/* consStif[level] * x, MGPIS.h:189,200 */
int ddpca_mg_spmv(ddpca_mg *, int level, const double *x, double *y);
/* realProl[level]^T * r  (n_{level+1} -> n_level), MGPIS.h:96 */
int ddpca_mg_restrict(ddpca_mg *, int level, const double *r_fine, double *r_coarse);
/* x_fine += realProl[level] * e_coarse, MGPIS.h:100 */
int ddpca_mg_prolong_add(ddpca_mg *, int level, const double *e_coarse, double *x_fine);
/* direSolv.solve(b) on consStif[0], MGPIS.h:58 */
int ddpca_mg_coarse_solve(ddpca_mg *, const double *b, double *x);

/* MGPIS::MULT_SOLV (MGPIS.h:130-160), BiCGSTAB_SOLV (:350-432), GMRES_SOLV (:227-348) on the same
 * kernels, one hierarchy per handle (not a batch).  The recurrences are device-resident: every
 * reduction, coefficient and stopping test (incl. the 11x10 Hessenberg QR of GMRES) stays on the
 * device, the host enqueues iterations ahead and polls one flag.  DDPCA_KRYLOV_HOSTLOOP=1 selects
 * the older drivers that read every scalar back (kept as a cross-check). */
int ddpca_mg_mult_solv(ddpca_mg *, const double *b, double *x, long *iters, double *resid);
int ddpca_mg_bicgstab(ddpca_mg *, int prec, const double *b, double *x, double rel_tol, long maxit,
                      long *iters, double *resid, double *tol_abs);
/* MGPIS::GMRES_SOLV (MGPIS.h:227-348): restarted GMRES(10), tolerance 1e-12 ||b|| and the
 * reference's stagnation stop; iters = iterNumb as the reference prints it */
int ddpca_mg_gmres(ddpca_mg *, int prec, const double *b, double *x, long *iters, double *resid, double *tol_abs);

/* ---- DIRE_SOLV: sparse direct solves with a host-computed factorisation --------
 * The reference factorises the interface mass matrices and the macroscopic problem
 * once, on the host, with Eigen::SimplicialLDLT (typedef DIRE_SOLV, PREP.h:107;
 * MCONTACT.h:837-847, :1229-1230) -- setup, out of scope -- and calls .solve() in
 * every ADMM iteration (MCONTACT.h:2553,2677,2696).  ddpca_ldlt_* is that solve phase
 * on the device: x = P^T L^-T D^-1 L^-1 P b (SimplicialCholesky.h:148-171).
 *   perm[i]  = solver.permutationP().indices()[i]
 *   L        = strictly-lower unit factor, RowMajor CSR (solver.matrixL() without diagonal)
 *   D        = solver.vectorD()                                                      */
typedef struct ddpca_ldlt ddpca_ldlt;
int ddpca_ldlt_create(int device, int n, const int *perm, const int *L_rowptr, const int *L_colidx, const double *L_val,
                      const double *D, ddpca_ldlt **out);
/* SPD operators up to 32768 rows (interface mass matrices; the coarse problems, which every rank solves once per
 * iteration): no host factor needed -- the dense inverse is built on the device (blocked Gauss-Jordan, 8 n^2 bytes)
 * and every solve is one product. */
int ddpca_ldlt_create_dense(int device, int n, const int *rowptr, const int *colidx, const double *val, ddpca_ldlt **out);
int ddpca_ldlt_solve(ddpca_ldlt *, const double *b, double *x);
int ddpca_ldlt_solve_dev(ddpca_ldlt *, const double *b_dev, double *x_dev);
int ddpca_ldlt_info(const ddpca_ldlt *, int *n, long *nnzL, int *stages_fwd, int *stages_bwd);
int ddpca_ldlt_destroy(ddpca_ldlt *);

/* ---- MCONTACT: the ADMM loop ----------------------------------------------------
 * ddpca_admm_* keeps the state of MCONTACT::CONTACT_ANALYSIS (resuDisp, inteAuxi,
 * inteLagr, MCONTACT.h:25-27) and every operator the loop consumes (MCONTACT.h:29-46,
 * built by MCONTACT::ESTABLISH on the host) in HBM.  ddpca_admm_step() is one pass of the
 * loop body (MCONTACT.h:2505-2704) plus the norms MCONTACT::MONITOR needs (:2737-2833);
 * the stopping logic (ring buffers, VECT_MEDI_OSCI, MULT_MAXI, :2838-2843) is scalar host
 * code and stays with the caller.  Every body solve is MG-PCG on the device (the reference
 * switches to a host LDLT below 50 000 rows, MCONTACT.h:2527; both are solves to 1e-14). */
typedef struct ddpca_admm ddpca_admm;
enum {
    DDPCA_OP_SYSTTRAN = 0,      /* systTran[ts][tv]       3 n_nodes x d n_c   MCONTACT.h:2521 */
    DDPCA_OP_SYSTTRAN_PENA = 1, /* systTran_pena[ts][tv]  3 n_nodes x d n_c   :2520,2673     */
    DDPCA_OP_INTEMASS = 2,      /* inteMass[ts][tv]       d n_c x d n_c       :2674          */
    DDPCA_OP_INTEMASS_PENA = 3, /* inteMass_pena[ts][tv]                      :2694          */
    DDPCA_OP_INPOLAGR = 4,      /* inpoLagr[ts][tv]       d n_ip x d n_c      :2632          */
    DDPCA_OP_INTEINPO = 5,      /* inteInpo[ts][tv]       d n_c x d n_ip      :2675          */
    DDPCA_OP_PEMAINPO_R = 6,    /* pemaInpo_r[ts][tv]     d n_ip x 3 n_nodes  :2634          */
    DDPCA_OP_GLOBTRAN = 7,      /* globTran[ts][tv]       n_glob x d n_c      :2545          */
    DDPCA_OP_GLOBTRAN_PENA = 8, /* globTran_pena[ts][tv]                      :2546          */
    DDPCA_OP_GLOBTRAN_D = 9,    /* globTran_D[ts][tv]     n_glob x 3 n_nodes  :2547          */
    DDPCA_OP_GLOBTRAN_1 = 10,   /* globTran_1[ts][tv]     n_glob1 x d n_c     :2579 (muscSett bit 1) */
    DDPCA_OP_COUNT = 11
};
enum { DDPCA_SOLVER_MASS = 0 /* inteDiso */, DDPCA_SOLVER_MASS_PENA = 1 /* inteDiso_pena */ };

/* muscSett: bit 0 = macroscopic problem (MCONTACT.h:858-860), bit 1 = interface-eliminated coarse problem (:861-863) */
int ddpca_admm_create(int device, int nbody, int niface, int muscSett, ddpca_admm **out);
/* Body v: the hierarchy of multGrid[v].mgpi (consStif[0..nlevels-1], realProl[0..nlevels-2], as for
 * ddpca_mg_create -- the arrays must stay valid until ddpca_admm_finalize, which builds ONE batched hierarchy
 * per group of bodies with equal level count) and the body's loop operators: forcOper (n_L x 3 n_nodes) is
 * MULTIGRID::ADDITIONAL_FORCE as one operator, consOper[L] prolOper[L]^T earlTran^T
 * (MULTIGRID.h:1257-1261); MULTIGRID::OUTP_SUB1 is forcOper^T u + dispCons
 * (MULTIGRID.h:1263-1281, dispCons = OUTP_SUB1(0)); consForc is multGrid[v].consForc. */
int ddpca_admm_set_body(ddpca_admm *, int v, int nlevels, const int *n,
                        const int *const *rowptr, const int *const *colidx, const double *const *val,
                        const int *const *P_rowptr, const int *const *P_colidx, const double *const *P_val,
                        int nfull, const double *consForc,
                        const int *F_rowptr, const int *F_colidx, const double *F_val, const double *dispCons);
/* smoother ordering of the bodies' hierarchies (default DDPCA_SMOOTH_MC; environment DDPCA_SMOOTHER=lex); before finalize */
int ddpca_admm_set_smoother(ddpca_admm *, int smoother_mode);
/* accuProl[v] (MCONTACT.h:864-872), needed when muscSett bit 0 or bit 1 is set */
int ddpca_admm_set_body_accuprol(ddpca_admm *, int v, int rows, int cols, const int *rowptr, const int *colidx, const double *val);
/* Interface ts between contBody[ts][0..1]; fricCoef < 0 tied, = 0 frictionless, > 0 Coulomb
 * (MCONTACT.h:15-18); gapTerm = pemaInpo[ts] * inpoNgap[ts] (MCONTACT.h:2636), length d n_ip
 * with d = 1 for fricCoef == 0 and 3 otherwise (MCONTACT.h:886-893). */
int ddpca_admm_set_interface(ddpca_admm *, int ts, int body0, int body1, double fricCoef, int nip, const double *gapTerm);
int ddpca_admm_set_side_op(ddpca_admm *, int ts, int tv, int op, int rows, int cols, const int *rowptr, const int *colidx, const double *val);
/* takes ownership of the solver: inteDiso[ts][tv] / inteDiso_pena[ts][tv] (MCONTACT.h:837-847) */
int ddpca_admm_set_side_solver(ddpca_admm *, int ts, int tv, int which, ddpca_ldlt *solver);
/* No factor for this side: its two mass systems are solved by Jacobi-preconditioned CG on the device (all such sides of
 * a rank in one batched solve per update) -- what the reference does with Eigen::ConjugateGradient and its default
 * diagonal preconditioner when inteMass has DIRE_MAXI rows or more (MCONTACT.h:2678-2683, :2698-2703).  Replaces both
 * ddpca_admm_set_side_solver calls of the side. */
int ddpca_admm_set_side_iterative(ddpca_admm *, int ts, int tv);
/* macroscopic problem: coarSolv_D (factorised globCoup, MCONTACT.h:1229-1230) and baseReco[nbody+1] (:850-857) */
int ddpca_admm_set_macro(ddpca_admm *, int nglob, const long *baseReco, ddpca_ldlt *coarSolv);
/* the same for a macroscopic problem beyond DIRE_MAXI rows (PREP.h:69): the reference then solves it with
 * MCONTACT's own multigrid hierarchy, mgpi.CG_SOLV(1, globForc, globSolu) (MCONTACT.h:2560-2562, hierarchy
 * built by DOUBLE_M, :1538-1670).  Takes ownership of the hierarchy (finest level = globCoup). */
int ddpca_admm_set_macro_mg(ddpca_admm *, int nglob, const long *baseReco, ddpca_mg *mgpi);
/* Interface-eliminated coarse problem (muscSett bit 1; built by MCONTACT::MULTISCALE_1, MCONTACT.h:1672-2343,
 * applied at :2575-2607):  globForc = globForc_1 + sum globTran_1[ts][tv] inteLagr[ts][tv] - sum globTran_D_1[v] resuDisp[v],
 * globSolu = coarSolv_D_1.solve(globForc), then the same correction of the bodies through accuProl / baseReco as bit 0.
 * globTran_1 is side operator DDPCA_OP_GLOBTRAN_1; accuProl is needed as for bit 0.  Takes ownership of the solver. */
int ddpca_admm_set_body_globtran_d1(ddpca_admm *, int v, int rows, int cols, const int *rowptr, const int *colidx, const double *val);
int ddpca_admm_set_macro1(ddpca_admm *, int nglob1, const long *baseReco, const double *globForc_1, ddpca_ldlt *coarSolv_D_1);
/* the same with globCoup_1 beyond DIRE_MAXI rows: mgpi_1.CG_SOLV(1, globForc, globSolu) (MCONTACT.h:2593-2595); takes
 * ownership of the hierarchy (finest level = globCoup_1) */
int ddpca_admm_set_macro1_mg(ddpca_admm *, int nglob1, const long *baseReco, const double *globForc_1, ddpca_mg *mgpi_1);
/* Multi-GPU, one process per GPU: body_rank[v] = owning rank; a rank uploads only
 * its own bodies and their interface sides (set_body / set_side_op / set_side_solver), but declares
 * EVERY interface (set_interface) and the macroscopic solver.  Call before ddpca_admm_set_body. */
int ddpca_admm_set_partition(ddpca_admm *, const int *body_rank, int my_rank);
/* Device buffers of the per-iteration exchange between ranks (e.g. torch tensors over NCCL):
 *   globForc   [nglob]   all-reduced (sum) after DDPCA_PH_MACRO_PARTIAL / _MACRO1_PARTIAL (MCONTACT.h:2541-2549: coarse RHS)
 *   trace_send [ntrace]  after DDPCA_PH_TRACES: this rank's signed side traces of its cross-rank interfaces
 *   trace_recv [ntrace]  (:2632-2635), grouped by peer (ddpca_admm_exchange_peers): range k is SENT to peer k and the
 *                        same range of trace_recv RECEIVED from it -- a pairwise swap, half the bytes of a reduction;
 *                        both owners then add the two parts (a + b == b + a bit for bit) and project the same gamma
 *   moni       [nmoni]   all-reduced (sum) after DDPCA_PH_MONITOR (MONITOR sums, :2737-2833)
 * Not needed on a single rank.  Sizes are known once every interface is declared; call set_exchange before the
 * first iteration, exchange_peers after ddpca_admm_finalize. */
int ddpca_admm_exchange_sizes(const ddpca_admm *, long *nglob, long *ntrace, long *nmoni);
int ddpca_admm_set_exchange(ddpca_admm *, double *globForc_dev, double *trace_send_dev, double *trace_recv_dev, double *moni_dev);
int ddpca_admm_exchange_peers(const ddpca_admm *, int *npeers, int *peer_rank, long *offset, long *count);
int ddpca_admm_set_stream(ddpca_admm *, void *stream);
enum { DDPCA_PH_BODIES = 0, DDPCA_PH_MACRO_PARTIAL = 1, DDPCA_PH_MACRO_APPLY = 2, DDPCA_PH_TRACES = 3, DDPCA_PH_INTERFACE = 4, DDPCA_PH_MONITOR = 5,
       DDPCA_PH_MACRO1_PARTIAL = 6, DDPCA_PH_MACRO1_APPLY = 7 /* muscSett bit 1: between MACRO_APPLY and TRACES, same globForc exchange buffer */ };
/* enqueue one phase of the loop body on the handle's stream (ddpca_admm_step = all of them in order) */
int ddpca_admm_phase(ddpca_admm *, int phase);
/* after the moni buffer has been all-reduced: synchronise and assemble the resuMoni row */
int ddpca_admm_monitor_row(ddpca_admm *, double *monitor_row, long *cg_iters, double *cg_dof_iters);
/* checks completeness, allocates the zero initial state (MCONTACT.h:875-894) */
int ddpca_admm_finalize(ddpca_admm *);
/* One iteration.  apply_macro = ((muscSett & 3) != 0 && tc <= MULT_MAXI) (MCONTACT.h:2540,2575) as evaluated by the
 * caller (MCONTACT.h:2540).  monitor_row (host, ddpca_admm_row_length() doubles) is the line
 * MONITOR appends to resuMoni.txt.  cg_iters / cg_dof_iters: CG iterations of this step summed
 * over bodies, and sum of n_L * iterations. */
int ddpca_admm_step(ddpca_admm *, int apply_macro, double *monitor_row, long *cg_iters, double *cg_dof_iters);
int ddpca_admm_row_length(const ddpca_admm *);
/* Repeat the analysis on the same handle: zero initial state again (MCONTACT.h:875-894), optionally new load
 * vectors multGrid[v].consForc (host, n_L doubles; copied on the handle's stream). */
int ddpca_admm_reset(ddpca_admm *);
int ddpca_admm_set_consforc(ddpca_admm *, int v, const double *consForc);
/* kernel-level entry for parity tests: the contact projection of the interface block (MCONTACT.h:2636-2668) on host
 * arrays: inpoGamm = proj(0.5 (t - gapTerm)), d = 1 (fricCoef == 0) or 3 components per integration point, fricStat as
 * OUTPUT_PRTR writes it (0 open / 1 slide / 2 stick in the second component's slot, MCONTACT.h:118,2656-2665) */
int ddpca_gamma_project(int device, int nip, int d, double fricCoef, const double *t, const double *gapTerm, double *inpoGamm, int *fricStat);
/* per-kernel-class timing of the batched body solves, as ddpca_mg_profile / ddpca_mg_profile_get (summed over batches) */
int ddpca_admm_profile(ddpca_admm *, int enable);
int ddpca_admm_profile_get(ddpca_admm *, int kclass, int level, double *ms, long *launches, double *bytes);
/* number of batched hierarchies and the CG iteration count of every body in the last step (0 for remote bodies) */
int ddpca_admm_body_iters(const ddpca_admm *, int *nbatches, long *iters);
int ddpca_admm_get_disp(ddpca_admm *, int v, double *resuDisp);
int ddpca_admm_get_side(ddpca_admm *, int ts, int tv, double *inteAuxi, double *inteLagr);
/* inpoGamm[ts] and fricStat of the last iteration (what OUTPUT_PRTR writes, MCONTACT.h:97-123) */
int ddpca_admm_get_gamma(ddpca_admm *, int ts, double *inpoGamm, int *fricStat);
long ddpca_admm_launch_count(ddpca_admm *, int reset);
int ddpca_admm_destroy(ddpca_admm *);

/* ---- one process, several GPUs -----------------------------------------------------
 * The reference spreads the bodies of the loop over host threads (`#pragma omp parallel for`, MCONTACT.h:2511,2629,2689).
 * A group is that idea with devices: member k is an ordinary ddpca_admm handle on devices[k] that owns the bodies with
 * body_rank[v] == k (and their interface sides).  Fill the members with the setters above -- set_body / set_side_op /
 * set_side_solver on the owner (solvers created on the owner's device), set_interface and the coarse solvers on EVERY
 * member -- then ddpca_admm_group_finalize.  ddpca_admm_group_step issues one iteration on all members and moves the
 * exchanges itself over NVLink peer copies ordered by events: coarse right-hand side to member 0, fixed-order sum, back;
 * signed side traces of shared interfaces straight into the partner's receive buffer (pairwise); MONITOR sums to the host.
 * State is read from the owner: ddpca_admm_get_disp(ddpca_admm_group_member(g, ddpca_admm_group_owner(g, v)), v, out). */
typedef struct ddpca_admm_group ddpca_admm_group;
/* greedy bin packing of bodies onto ranks by weight (e.g. nnz of the finest operator); contBody[2 ts + 0..1] */
int ddpca_partition_bodies(int nbody, const double *weight, int niface, const int *contBody, int nranks, int *body_rank);
int ddpca_admm_group_create(int ndev, const int *devices, int nbody, int niface, int muscSett, const int *body_rank, ddpca_admm_group **out);
int ddpca_admm_group_size(const ddpca_admm_group *);
ddpca_admm *ddpca_admm_group_member(ddpca_admm_group *, int k);
int ddpca_admm_group_owner(const ddpca_admm_group *, int v);      /* member index of body v */
int ddpca_admm_group_device(const ddpca_admm_group *, int k);     /* CUDA device of member k */
int ddpca_admm_group_finalize(ddpca_admm_group *);
int ddpca_admm_group_step(ddpca_admm_group *, int apply_macro, double *monitor_row, long *cg_iters, double *cg_dof_iters);
long ddpca_admm_group_launch_count(ddpca_admm_group *, int reset);
int ddpca_admm_group_destroy(ddpca_admm_group *);                 /* destroys the members, too */

/* ---- introspection / measurement -------------------------------------------*/
/* rows, nnz, number of row groups and stages of a level's device layout */
int ddpca_mg_level_info(const ddpca_mg *, int level, long *n, long *nnz, int *ngroups, int *nstages);
/* number of kernels this handle has launched since creation (or last reset) */
long ddpca_mg_launch_count(ddpca_mg *, int reset);
/* Run the handle on an external stream (e.g. torch's current stream), so that
 * CUDA events recorded there bracket the work.  stream = cudaStream_t. */
int ddpca_mg_set_stream(ddpca_mg *, void *stream);
/* Per-kernel-class timing: when enabled, pcg/vcycle run un-captured and every
 * launch is bracketed by CUDA events on the handle's stream. */
int ddpca_mg_profile(ddpca_mg *, int enable);
/* accumulated ms, launches and algorithmic bytes per class since enable */
int ddpca_mg_profile_get(ddpca_mg *, int kclass, int level, double *ms, long *launches, double *bytes);
/* device time of the last ddpca_mg_pcg / ddpca_mg_pcg_dev call (CUDA events on the handle's
 * stream): solve only, and upload/download legs of the host variant */
int ddpca_mg_last_timing(ddpca_mg *, double *solve_ms, double *h2d_ms, double *d2h_ms);

#ifdef __cplusplus
}
#endif
#endif
