/*
 * mgpis_oracle.c -- CPU restatement of the reference's multigrid-preconditioned
 * iterative solver, in plain C.  TEST INFRASTRUCTURE ONLY: this file is the
 * parity checker for the CUDA path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the product
 * (ddpca-admm_b200/) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function
 * below against vectors produced by the untouched reference compiled from
 * /root/reference (oracle/ref_drivers/beam_nodd.cpp -> tests/golden/*.ddpk).
 *
 * Each function cites the reference lines it follows (paths relative to
 * /root/reference).  Matrices are Eigen RowMajor compressed storage: int32
 * rowptr[n+1], int32 colidx[nnz] sorted per row, double val[nnz].
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int n, m;            /* rows, cols */
    const int *rp, *ci;
    const double *v;
} csr_t;

typedef struct {
    int nlev;            /* maxiLeve + 1 */
    csr_t *A;            /* consStif[0..L]          MGPIS.h:15 */
    csr_t *P;            /* realProl[0..L-1]        MGPIS.h:13 */
    int **dpos;          /* position of the diagonal entry in each row: the
                            L/D/U split of MGPIS::ESTABLISH, MGPIS.h:40-53,
                            kept as an index instead of three matrix copies */
    double *ldl;         /* dense LDL^T factor of consStif[0] (unit L below the
                            diagonal, D on it); stands for SimplicialLDLT,
                            MGPIS.h:185 + Eigen SimplicialCholesky.h:148-171 */
    int n0;
} orc_mg;

/* ---- dense LDL^T on level 0 (exact direct solve, like DIRE_SOLV) ---------- */
static int dense_ldlt(double *a, int n)
{
    /* right-looking, lower storage a[i*n+j], j<=i */
    for (int k = 0; k < n; k++) {
        double d = a[(size_t)k * n + k];
        if (d == 0.0) return -1;
        for (int i = k + 1; i < n; i++) a[(size_t)i * n + k] /= d;
        for (int i = k + 1; i < n; i++) {
            double lik_d = a[(size_t)i * n + k] * d;
            if (lik_d == 0.0) continue;
            double *ai = a + (size_t)i * n;
            for (int j = k + 1; j <= i; j++) ai[j] -= lik_d * a[(size_t)j * n + k];
        }
    }
    return 0;
}

static void dense_ldlt_solve(const double *a, int n, const double *b, double *x)
{
    memcpy(x, b, sizeof(double) * n);
    for (int i = 0; i < n; i++) {
        double s = x[i];
        const double *ai = a + (size_t)i * n;
        for (int j = 0; j < i; j++) s -= ai[j] * x[j];
        x[i] = s;
    }
    for (int i = 0; i < n; i++) x[i] /= a[(size_t)i * n + i];
    for (int i = n - 1; i >= 0; i--) {
        double s = x[i];
        for (int j = i + 1; j < n; j++) s -= a[(size_t)j * n + i] * x[j];
        x[i] = s;
    }
}

/* ---- construction --------------------------------------------------------- */
orc_mg *orc_mg_create(int nlev, const int *n, const int *const *rp, const int *const *ci,
                      const double *const *v, const int *const *prp, const int *const *pci,
                      const double *const *pv)
{
    orc_mg *h = (orc_mg *)calloc(1, sizeof(orc_mg));
    h->nlev = nlev;
    h->A = (csr_t *)calloc(nlev, sizeof(csr_t));
    h->P = (csr_t *)calloc(nlev > 1 ? nlev - 1 : 1, sizeof(csr_t));
    h->dpos = (int **)calloc(nlev, sizeof(int *));
    for (int l = 0; l < nlev; l++) {
        h->A[l].n = h->A[l].m = n[l];
        h->A[l].rp = rp[l]; h->A[l].ci = ci[l]; h->A[l].v = v[l];
        /* MGPIS::ESTABLISH (MGPIS.h:44-51): strictly lower / diagonal / strictly upper */
        h->dpos[l] = (int *)malloc(sizeof(int) * n[l]);
        for (int i = 0; i < n[l]; i++) {
            int p = rp[l][i];
            while (p < rp[l][i + 1] && ci[l][p] < i) p++;
            h->dpos[l][i] = (p < rp[l][i + 1] && ci[l][p] == i) ? p : -1;
        }
    }
    for (int l = 0; l + 1 < nlev; l++) {
        h->P[l].n = n[l + 1]; h->P[l].m = n[l];
        h->P[l].rp = prp[l]; h->P[l].ci = pci[l]; h->P[l].v = pv[l];
    }
    h->n0 = n[0];
    return h;
}

/* direSolv.compute(consStif[0]) -- MGPIS.h:185 (done once per handle here; the
 * reference redoes it on every CG_SOLV call, which does not change results) */
int orc_mg_factor(orc_mg *h)
{
    if (h->ldl) return 0;
    int n = h->n0;
    h->ldl = (double *)calloc((size_t)n * n, sizeof(double));
    const csr_t *A = &h->A[0];
    for (int i = 0; i < n; i++)
        for (int p = A->rp[i]; p < A->rp[i + 1]; p++)
            if (A->ci[p] <= i) h->ldl[(size_t)i * n + A->ci[p]] = A->v[p];
    return dense_ldlt(h->ldl, n);
}

void orc_mg_destroy(orc_mg *h)
{
    if (!h) return;
    for (int l = 0; l < h->nlev; l++) free(h->dpos[l]);
    free(h->dpos); free(h->A); free(h->P); free(h->ldl); free(h);
}

/* ---- kernels --------------------------------------------------------------- */
/* y = A x : Eigen row-major sparse * dense (SparseDenseProduct.h:34-82) */
void orc_spmv(int n, const int *rp, const int *ci, const double *v, const double *x, double *y)
{
    for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int p = rp[i]; p < rp[i + 1]; p++) s += v[p] * x[ci[p]];
        y[i] = s;
    }
}

/* y = A^T x for a row-major A (n rows, m cols) -- realProl^T * r, MGPIS.h:96 */
void orc_spmv_t(int n, int m, const int *rp, const int *ci, const double *v, const double *x, double *y)
{
    memset(y, 0, sizeof(double) * m);
    for (int i = 0; i < n; i++) {
        double xi = x[i];
        for (int p = rp[i]; p < rp[i + 1]; p++) y[ci[p]] += v[p] * xi;
    }
}

void orc_coarse_solve(orc_mg *h, const double *b, double *x)
{
    orc_mg_factor(h);
    dense_ldlt_solve(h->ldl, h->n0, b, x);
}

/* one symmetric Gauss-Seidel pass, MGPIS.h:65-77 (pre) == :102-114 (post).
 * p1 (length n) receives D x - p0, which the caller needs for the residual. */
static void sgs_pass(const csr_t *A, const int *dpos, const double *b, double *x, double *p1)
{
    int n = A->n;
    double *bpp0 = (double *)malloc(sizeof(double) * n);
    double *p0 = (double *)malloc(sizeof(double) * n);
    /* p_0 = - consUppe * x ; bpp0 = b + p_0                         :66-67 */
    for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int p = dpos[i] + 1; p < A->rp[i + 1]; p++) s += A->v[p] * x[A->ci[p]];
        p0[i] = -s;
        bpp0[i] = b[i] + p0[i];
    }
    /* forward sweep with consLowe                                    :68-71 */
    for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int p = A->rp[i]; p < dpos[i]; p++) s += A->v[p] * x[A->ci[p]];
        x[i] = (bpp0[i] - s) / A->v[dpos[i]];
    }
    /* p_1 = consDiag * x - p_0                                       :72 */
    for (int i = 0; i < n; i++) p1[i] = A->v[dpos[i]] * x[i] - p0[i];
    /* backward sweep with consUppe                                   :73-76 */
    for (int i = n - 1; i >= 0; i--) {
        double s = 0.0;
        for (int p = dpos[i] + 1; p < A->rp[i + 1]; p++) s += A->v[p] * x[A->ci[p]];
        x[i] = (p1[i] - s) / A->v[dpos[i]];
    }
    free(bpp0); free(p0);
}

/* MGPIS::MULT_VCYC, MGPIS.h:55-128.  x is in/out exactly like resuSolu. */
void orc_vcycle(orc_mg *h, int lev, const double *b, double *x)
{
    if (lev == 0) {                                  /* :57-60 */
        orc_coarse_solve(h, b, x);
        return;
    }
    const csr_t *A = &h->A[lev];
    const int *dpos = h->dpos[lev];
    int n = A->n, nc = h->A[lev - 1].n;
    double *p1 = (double *)malloc(sizeof(double) * n);
    sgs_pass(A, dpos, b, x, p1);                     /* :65-77 */
    /* resiErro = b - (p_1 + consLowe * x)              :92 */
    double *r = (double *)malloc(sizeof(double) * n);
    for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int p = A->rp[i]; p < dpos[i]; p++) s += A->v[p] * x[A->ci[p]];
        r[i] = b[i] - (p1[i] + s);
    }
    /* recursion with realProl^T r, zero initial guess  :93-99 */
    const csr_t *P = &h->P[lev - 1];
    double *rc = (double *)malloc(sizeof(double) * nc);
    double *ec = (double *)calloc(nc, sizeof(double));
    orc_spmv_t(P->n, P->m, P->rp, P->ci, P->v, r, rc);
    orc_vcycle(h, lev - 1, rc, ec);
    /* x = x + realProl * e                             :100 */
    for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int p = P->rp[i]; p < P->rp[i + 1]; p++) s += P->v[p] * ec[P->ci[p]];
        x[i] += s;
    }
    sgs_pass(A, dpos, b, x, p1);                     /* :102-114 */
    free(p1); free(r); free(rc); free(ec);
}

static double dot(int n, const double *a, const double *b)
{
    double s = 0.0;
    for (int i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}

/* MGPIS::CG_SOLV, MGPIS.h:163-225.  precSwit 0: DIAG_PREC (PREP.h:393-401),
 * 1: one V-cycle.  Returns iterNumb (the reference prints iterNumb-1, :221).
 * resid_out = final ||r||_2 of the recurrence residual, tol_out = 1e-14*||b||. */
long orc_cg_solv(orc_mg *h, long precSwit, const double *b, double *x, double *resid_out, double *tol_out)
{
    int L = h->nlev - 1;
    const csr_t *A = &h->A[L];
    int n = A->n;
    memset(x, 0, sizeof(double) * n);                       /* :173 */
    long maxiNumb = n;                                      /* :174 */
    double toleLimi = 1.0E-14 * sqrt(dot(n, b, b));         /* :175 */
    double *dinv = NULL;
    if (precSwit == 0) {                                    /* :182 */
        dinv = (double *)malloc(sizeof(double) * n);
        for (int i = 0; i < n; i++) dinv[i] = 1.0 / A->v[h->dpos[L][i]];
    } else {
        orc_mg_factor(h);                                   /* :185 */
    }
    double *r = (double *)malloc(sizeof(double) * n);
    double *p = (double *)calloc(n, sizeof(double));
    double *q = (double *)malloc(sizeof(double) * n);
    double *z = (double *)malloc(sizeof(double) * n);
    memcpy(r, b, sizeof(double) * n);                       /* :189, x = 0 */
    if (precSwit == 0) for (int i = 0; i < n; i++) p[i] = dinv[i] * r[i];   /* :192 */
    else orc_vcycle(h, L, r, p);                            /* :195 */
    double delt_new = dot(n, r, p);                         /* :197 */
    long it = 0;
    while (it < maxiNumb && sqrt(dot(n, r, r)) > toleLimi) { /* :198 */
        orc_spmv(n, A->rp, A->ci, A->v, p, q);              /* :200 */
        double alph = delt_new / dot(n, p, q);              /* :201 */
        for (int i = 0; i < n; i++) x[i] += alph * p[i];    /* :202 */
        for (int i = 0; i < n; i++) r[i] -= alph * q[i];    /* :203 */
        if (precSwit == 0) for (int i = 0; i < n; i++) z[i] = dinv[i] * r[i];  /* :206 */
        else { memset(z, 0, sizeof(double) * n); orc_vcycle(h, L, r, z); }     /* :204,209 */
        double delt_old = delt_new;                         /* :211 */
        delt_new = dot(n, r, z);                            /* :212 */
        double beta = delt_new / delt_old;                  /* :213 */
        for (int i = 0; i < n; i++) p[i] = z[i] + beta * p[i];  /* :214 */
        it++;
    }
    if (resid_out) *resid_out = sqrt(dot(n, r, r));
    if (tol_out) *tol_out = toleLimi;
    free(r); free(p); free(q); free(z); free(dinv);
    return it;
}

/* MGPIS::MULT_SOLV, MGPIS.h:130-160.  Returns iterNumb; *resid_out = moniErro[iterNumb % 5]. */
long orc_mult_solv(orc_mg *h, const double *b, double *x, double *resid_out)
{
    int L = h->nlev - 1;
    const csr_t *A = &h->A[L];
    int n = A->n;
    memset(x, 0, sizeof(double) * n);                       /* :133 */
    long maxiNumb = 10000;                                  /* :134 */
    orc_mg_factor(h);                                       /* :137 */
    double *r = (double *)malloc(sizeof(double) * n);
    double moni[5] = {0, 0, 0, 0, 0};
    long it = 0;
    while (it < maxiNumb) {                                 /* :142 */
        orc_vcycle(h, L, b, x);                             /* :143 */
        orc_spmv(n, A->rp, A->ci, A->v, x, r);
        for (int i = 0; i < n; i++) r[i] = b[i] - r[i];     /* :144 */
        moni[it % 5] = sqrt(dot(n, r, r));                  /* :146 */
        if (it >= 4) {                                      /* :147-153, VECT_MEDI_OSCI PREP.h:147-153 */
            double mx = moni[0], mn = moni[0];
            for (int k = 1; k < 5; k++) { if (moni[k] > mx) mx = moni[k]; if (moni[k] < mn) mn = moni[k]; }
            if (mx - mn < 0.1 * ((mx + mn) / 2.0)) break;
        }
        it++;
    }
    if (resid_out) *resid_out = moni[it % 5];
    free(r);
    return it;
}

/* MGPIS::BiCGSTAB_SOLV, MGPIS.h:350-432.  Returns iterNumb. */
long orc_bicgstab(orc_mg *h, long precSwit, const double *b, double *x, double *resid_out, double *tol_out)
{
    int L = h->nlev - 1;
    const csr_t *A = &h->A[L];
    int n = A->n;
    memset(x, 0, sizeof(double) * n);                       /* :361 */
    long maxiNumb = n;                                      /* :362 */
    double toleLimi = 1.0E-14 * sqrt(dot(n, b, b));         /* :363 */
    double *dinv = NULL;
    if (precSwit == 0) {
        dinv = (double *)malloc(sizeof(double) * n);
        for (int i = 0; i < n; i++) dinv[i] = 1.0 / A->v[h->dpos[L][i]];
    } else orc_mg_factor(h);                                /* :373 */
    size_t nb = sizeof(double) * n;
    double *r = (double *)malloc(nb), *rh = (double *)malloc(nb), *p = (double *)calloc(n, sizeof(double));
    double *v = (double *)calloc(n, sizeof(double)), *s = (double *)malloc(nb), *t = (double *)malloc(nb);
    double *ph = (double *)malloc(nb), *sh = (double *)malloc(nb);
    memcpy(r, b, nb);                                       /* :377 */
    memcpy(rh, r, nb);                                      /* :378 */
    double rho[2] = {0, 0}, alph = 0, omeg = 0;
    long it = 0;
    while (it < maxiNumb && sqrt(dot(n, r, r)) > toleLimi) {            /* :382 */
        rho[(it + 1) % 2] = dot(n, rh, r);                              /* :383 */
        if (fabs(rho[(it + 1) % 2]) == 0.0) break;                      /* :384-387 */
        if (it == 0) memcpy(p, r, nb);                                  /* :389 */
        else {
            double beta = (rho[(it + 1) % 2] / rho[it % 2]) * (alph / omeg);   /* :392-393 */
            for (int i = 0; i < n; i++) p[i] = r[i] + beta * (p[i] - omeg * v[i]);   /* :394 */
        }
        if (precSwit == 0) for (int i = 0; i < n; i++) ph[i] = dinv[i] * p[i];
        else { memset(ph, 0, nb); orc_vcycle(h, L, p, ph); }            /* :396-402 */
        orc_spmv(n, A->rp, A->ci, A->v, ph, v);                         /* :403 */
        alph = rho[(it + 1) % 2] / dot(n, rh, v);                       /* :404 */
        for (int i = 0; i < n; i++) s[i] = r[i] - alph * v[i];          /* :405 */
        if (sqrt(dot(n, s, s)) <= 0.0) {                                /* :406-409 */
            for (int i = 0; i < n; i++) x[i] += alph * ph[i];
            break;
        }
        if (precSwit == 0) for (int i = 0; i < n; i++) sh[i] = dinv[i] * s[i];
        else { memset(sh, 0, nb); orc_vcycle(h, L, s, sh); }            /* :410-416 */
        orc_spmv(n, A->rp, A->ci, A->v, sh, t);                         /* :417 */
        omeg = dot(n, t, s) / dot(n, t, t);                             /* :418 */
        for (int i = 0; i < n; i++) x[i] += alph * ph[i] + omeg * sh[i];/* :419 */
        for (int i = 0; i < n; i++) r[i] = s[i] - omeg * t[i];          /* :420 */
        it++;
    }
    if (resid_out) *resid_out = sqrt(dot(n, r, r));
    if (tol_out) *tol_out = toleLimi;
    free(r); free(rh); free(p); free(v); free(s); free(t); free(ph); free(sh); free(dinv);
    return it;
}

/* MGPIS::GMRES_SOLV, MGPIS.h:227-348: left-preconditioned GMRES, restart 10, classical Gram-Schmidt
 * Arnoldi, QR of the Hessenberg matrix by Gram-Schmidt, true residual every step, stop when
 * ||r|| <= tol or (||r|| <= 100 tol and the last 10 residuals stagnate).  Returns iterNumb. */
long orc_gmres(orc_mg *h, long precSwit, const double *b, double *x, double *resid_out, double *tol_out)
{
    enum { STAG = 10 };
    int L = h->nlev - 1;
    const csr_t *A = &h->A[L];
    int n = A->n;
    double *dinv = NULL;
    if (precSwit == 0) {
        dinv = (double *)malloc(sizeof(double) * n);
        for (int i = 0; i < n; i++) dinv[i] = 1.0 / A->v[h->dpos[L][i]];
    } else orc_mg_factor(h);                                 /* :245 */
    memset(x, 0, sizeof(double) * n);                        /* :248 */
    long maxiNumb = n;                                       /* :249 */
    double toleLimi = 1.0E-12 * sqrt(dot(n, b, b));          /* :250 */
    size_t nb = sizeof(double) * n;
    double *V = (double *)malloc(nb * (STAG + 1));           /* orthBasi */
    double *x0 = (double *)malloc(nb), *r = (double *)malloc(nb), *w0 = (double *)malloc(nb), *w = (double *)malloc(nb);
    double H[STAG + 1][STAG], Q[STAG + 1][STAG], R[STAG][STAG], moni[STAG];
    double normR0 = 0.0;
    memset(moni, 0, sizeof(moni));
    long it = 0;
    while (it < maxiNumb) {                                  /* :261 */
        int k = (int)(it % STAG);
        if (k == 0) {                                        /* restart :263-276 */
            memcpy(x0, x, nb);
            orc_spmv(n, A->rp, A->ci, A->v, x0, r);
            for (int i = 0; i < n; i++) r[i] = b[i] - r[i];
            if (precSwit == 0) for (int i = 0; i < n; i++) w[i] = dinv[i] * r[i];
            else { memset(w, 0, nb); orc_vcycle(h, L, r, w); }
            normR0 = sqrt(dot(n, w, w));
            for (int i = 0; i < n; i++) V[i] = w[i] / normR0;
            memset(H, 0, sizeof(H)); memset(Q, 0, sizeof(Q)); memset(R, 0, sizeof(R));
        }
        orc_spmv(n, A->rp, A->ci, A->v, V + (size_t)k * n, w0);           /* :278 */
        if (precSwit == 0) for (int i = 0; i < n; i++) w[i] = dinv[i] * w0[i];
        else { memset(w, 0, nb); orc_vcycle(h, L, w0, w); }               /* :279-285 */
        for (int j = 0; j <= k; j++) H[j][k] = dot(n, V + (size_t)j * n, w);   /* b_i :286 */
        for (int j = 0; j <= k; j++) for (int i = 0; i < n; i++) w[i] -= H[j][k] * V[(size_t)j * n + i];   /* :287 */
        double nq = sqrt(dot(n, w, w));                                   /* :288 */
        H[k + 1][k] = nq;                                                 /* :289-293 */
        for (int i = 0; i < n; i++) V[(size_t)(k + 1) * n + i] = w[i] / nq;   /* :294-296 */
        /* QR of H by Gram-Schmidt, one column per step  :297-316 */
        {
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
            nc = sqrt(nc);
            R[k][k] = nc;
            for (int i = 0; i <= k + 1; i++) Q[i][k] = col[i] / nc;
        }
        /* y = R^-1 (normR0 * Q.row(0)^T)  :317-324 */
        double y[STAG];
        for (int j = k; j >= 0; j--) {
            double s = normR0 * Q[0][j];
            for (int c = j + 1; c <= k; c++) s -= R[j][c] * y[c];
            y[j] = s / R[j][j];
        }
        memcpy(x, x0, nb);                                                /* :325 */
        for (int j = 0; j <= k; j++) for (int i = 0; i < n; i++) x[i] += y[j] * V[(size_t)j * n + i];
        orc_spmv(n, A->rp, A->ci, A->v, x, r);                            /* :326 */
        for (int i = 0; i < n; i++) r[i] = b[i] - r[i];
        moni[k] = sqrt(dot(n, r, r));                                     /* :328 */
        if (it >= STAG - 1) {                                             /* :333-341 */
            double mx = moni[0], mn = moni[0];
            for (int j = 1; j < STAG; j++) { if (moni[j] > mx) mx = moni[j]; if (moni[j] < mn) mn = moni[j]; }
            double medi = (mx + mn) / 2.0, osci = mx - mn;
            if (moni[k] <= toleLimi || (moni[k] <= 1.0E2 * toleLimi && osci < 0.1 * medi)) break;
        }
        it++;
    }
    if (resid_out) *resid_out = moni[it % STAG];
    if (tol_out) *tol_out = toleLimi;
    free(V); free(x0); free(r); free(w0); free(w); free(dinv);
    return it;
}
