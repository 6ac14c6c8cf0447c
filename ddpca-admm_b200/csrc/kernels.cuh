// kernels.cuh -- sm_100a FP64 kernels of the MG-PCG hot path.
//
// All of these are HBM-bound sparse/streaming kernels: no tensor cores.  The
// design rules that matter (blackwell_cuda_programming.md G1,G2,G7,G13,G14):
// coalesced streaming of (val,col) pairs, read-only path for operators,
// enough bytes in flight per SM, deterministic two-stage reductions, grids
// sized from the 148-SM machine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ddpca {

constexpr int kNumPart = 1024;   // slots of a partial-sum buffer (one per CTA of a reducing kernel)

// One multigrid level in its device (stage-permuted) layout.
struct LvlView {
    int n;
    const int *__restrict__ rp;      // [n+1]
    const int *__restrict__ ci;      // [nnz] sorted per row
    const double *__restrict__ v;    // [nnz]
    const int *__restrict__ dpos;    // [n] position of the diagonal entry of each row
    const int *__restrict__ gstart;  // [ngroups+1] first row of each group (<= 3 rows)
};

struct CsrView {
    int rows;
    const int *__restrict__ rp;
    const int *__restrict__ ci;
    const double *__restrict__ v;
};

// Scalars of one PCG solve, resident in HBM (MGPIS.h:173-214).
struct PcgState {
    double delta_new, delta_old, pq, rr, bb, tol, rel_tol, alpha, beta, rz;
    long long it, maxit;
    int done;      // 1 once the loop condition of MGPIS.h:198 is false
    int pad;
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int LANES>
__device__ __forceinline__ double subwarp_sum(double v)
{
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming (read-once) operator loads: read-only path, do not pollute L1
__device__ __forceinline__ double ld_stream(const double *p)
{
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ int ld_stream(const int *p)
{
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// block-wide deterministic sum -> partial[blockIdx.x]; blockDim.x multiple of 32, <= 1024
__device__ __forceinline__ void block_sum_to_partial(double acc, double *partial)
{
    __shared__ double sm[32];
    acc = warp_sum(acc);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sm[w] = acc;
    __syncthreads();
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        double t = lane < nw ? sm[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) partial[blockIdx.x] = t;
    }
}

// sum of partial[0..np) by one warp, fixed order
__device__ __forceinline__ double warp_reduce_partials(const double *partial, int np)
{
    int lane = threadIdx.x & 31;
    double t = 0.0;
    for (int k = lane; k < np; k += 32) t += partial[k];
    return warp_sum(t);
}

// ------------------------------------------------------------------------------------
// K3: forward Gauss-Seidel relaxation of one row group (MGPIS.h:66-72 on the permuted level)
//   x_i   = (b_i - sum_{j<i} a_ij x_j(new) - sum_{j>i} a_ij x_j(old)) / a_ii
//   p1_i  = a_ii x_i + sum_{j>i} a_ij x_j(old)            ( = b - L x, MGPIS.h:72 )
// ZERO_X: x is known to be zero on entry (first smoothing of a V-cycle, MGPIS.h:93,204):
//         the strictly-upper half is not read at all.
// ------------------------------------------------------------------------------------
template <bool ZERO_X, bool NC_X>
__device__ __forceinline__ void group_fwd(const LvlView &A, int g, const double *__restrict__ b,
                                          double *x, double *__restrict__ p1, int lane)
{
    const int r0 = A.gstart[g];
    const int gs = A.gstart[g + 1] - r0;
    double sL[3] = {0.0, 0.0, 0.0}, sU[3] = {0.0, 0.0, 0.0};
    double blk[3][3], bb[3], xo[3], dg[3];
    int pdv[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        if (r < gs) {
            const int i = r0 + r;
            const int pb = A.rp[i], pd = A.dpos[i], pe = A.rp[i + 1];
            pdv[r] = pd;
            bb[r] = b[i];
            xo[r] = ZERO_X ? 0.0 : x[i];
#pragma unroll
            for (int c = 0; c < 3; c++) blk[r][c] = (c < gs) ? A.v[pd - r + c] : 0.0;
            dg[r] = blk[r][r];
            for (int p = pb + lane; p < pd - r; p += 32) {
                const int c = ld_stream(A.ci + p);
                const double a = ld_stream(A.v + p);
                sL[r] += a * (NC_X ? __ldg(x + c) : x[c]);
            }
            if (!ZERO_X) {
                for (int p = pd + (gs - r) + lane; p < pe; p += 32) {
                    const int c = ld_stream(A.ci + p);
                    const double a = ld_stream(A.v + p);
                    sU[r] += a * (NC_X ? __ldg(x + c) : x[c]);
                }
            }
        }
    }
    (void)pdv;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        sL[r] = warp_sum(sL[r]);
        if (!ZERO_X) sU[r] = warp_sum(sU[r]);
    }
    // sequential in-group solve, done redundantly by every lane (no divergence); lane 0 stores
    double xn[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int r = 0; r < 3; r++) {
        if (r < gs) {
            double inL = 0.0, inU = 0.0;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                if (c < r) inL += blk[r][c] * xn[c];
                if (c > r && c < gs) inU += blk[r][c] * xo[c];
            }
            const double up = sU[r] + inU;
            xn[r] = (bb[r] - sL[r] - inL - up) / dg[r];
            if (lane == 0) {
                x[r0 + r] = xn[r];
                p1[r0 + r] = dg[r] * xn[r] + up;
            }
        }
    }
}

// K4: backward relaxation of one row group (MGPIS.h:73-76):  x_i = (p1_i - sum_{j>i} a_ij x_j) / a_ii
template <bool NC_X>
__device__ __forceinline__ void group_bwd(const LvlView &A, int g, const double *__restrict__ p1,
                                          double *x, int lane)
{
    const int r0 = A.gstart[g];
    const int gs = A.gstart[g + 1] - r0;
    double sU[3] = {0.0, 0.0, 0.0};
    double blk[3][3], pp[3], dg[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        if (r < gs) {
            const int i = r0 + r;
            const int pd = A.dpos[i], pe = A.rp[i + 1];
            pp[r] = p1[i];
#pragma unroll
            for (int c = 0; c < 3; c++) blk[r][c] = (c < gs && c >= r) ? A.v[pd - r + c] : 0.0;
            dg[r] = blk[r][r];
            for (int p = pd + (gs - r) + lane; p < pe; p += 32) {
                const int c = ld_stream(A.ci + p);
                const double a = ld_stream(A.v + p);
                sU[r] += a * (NC_X ? __ldg(x + c) : x[c]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 3; r++) sU[r] = warp_sum(sU[r]);
    double xn[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int r = 2; r >= 0; r--) {
        if (r < gs) {
            double inU = 0.0;
#pragma unroll
            for (int c = 0; c < 3; c++)
                if (c > r && c < gs) inU += blk[r][c] * xn[c];
            xn[r] = (pp[r] - sU[r] - inU) / dg[r];
            if (lane == 0) x[r0 + r] = xn[r];
        }
    }
}

// one stage = groups [g0,g1): mutually independent, one warp per group
template <bool ZERO_X>
__global__ void __launch_bounds__(256) k_sweep_fwd_stage(LvlView A, int g0, int g1, const double *__restrict__ b,
                                                         double *x, double *__restrict__ p1, const int *done)
{
    if (done && *done) return;
    const int g = g0 + (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (g >= g1) return;
    group_fwd<ZERO_X, true>(A, g, b, x, p1, threadIdx.x & 31);
}

__global__ void __launch_bounds__(256) k_sweep_bwd_stage(LvlView A, int g0, int g1, const double *__restrict__ p1,
                                                         double *x, const int *done)
{
    if (done && *done) return;
    const int g = g0 + (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (g >= g1) return;
    group_bwd<true>(A, g, p1, x, threadIdx.x & 31);
}

// a run of small stages [s0,s1) relaxed by ONE CTA, __syncthreads() between stages
// (latency-bound regime: LEX wavefronts, coarse levels)
template <bool ZERO_X>
__global__ void __launch_bounds__(1024) k_sweep_fwd_multi(LvlView A, const int *__restrict__ stage_group, int s0, int s1,
                                                          const double *__restrict__ b, double *x, double *p1, const int *done)
{
    if (done && *done) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int s = s0; s < s1; s++) {
        const int ga = stage_group[s], gb = stage_group[s + 1];
        for (int g = ga + w; g < gb; g += nw) group_fwd<ZERO_X, false>(A, g, b, x, p1, lane);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) k_sweep_bwd_multi(LvlView A, const int *__restrict__ stage_group, int s0, int s1,
                                                          const double *p1, double *x, const int *done)
{
    if (done && *done) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int s = s1 - 1; s >= s0; s--) {
        const int ga = stage_group[s], gb = stage_group[s + 1];
        for (int g = ga + w; g < gb; g += nw) group_bwd<false>(A, g, p1, x, lane);
        __syncthreads();
    }
}

// K2: r = b - (p1 + L x)   (MGPIS.h:92) -- strictly-lower half only, LANES lanes per row
template <int LANES>
__global__ void __launch_bounds__(256) k_resid_lower(LvlView A, const double *__restrict__ b, const double *__restrict__ p1,
                                                     const double *__restrict__ x, double *__restrict__ r, const int *done)
{
    if (done && *done) return;
    const int sub = threadIdx.x % LANES;
    const int i = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) / LANES);
    double s = 0.0;
    if (i < A.n) {
        const int pb = A.rp[i], pd = A.dpos[i];
        for (int p = pb + sub; p < pd; p += LANES) s += ld_stream(A.v + p) * __ldg(x + ld_stream(A.ci + p));
    }
    s = subwarp_sum<LANES>(s);
    if (i < A.n && sub == 0) r[i] = b[i] - (p1[i] + s);
}

// K1/K5/K6: y (=|+=) A x, LANES lanes per row; DOT: partial[blockIdx] = sum_i w_i * y_i
// (fused p.q of MGPIS.h:201).  Grid-stride so that DOT stays deterministic for a fixed grid.
template <int LANES, bool ADD, bool DOT>
__global__ void __launch_bounds__(256) k_spmv(CsrView A, const double *__restrict__ x, double *y,
                                              const double *__restrict__ w, double *partial, const int *done)
{
    if (done && *done) return;
    const int sub = threadIdx.x % LANES;
    const int rows_per_pass = (int)((gridDim.x * (unsigned)blockDim.x) / LANES);
    double acc = 0.0;
    for (int base = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) / LANES);; base += rows_per_pass) {
        // all lanes of a warp leave together: rows are assigned warp-uniformly
        const int first_row_of_warp = base - (int)((threadIdx.x & 31) / LANES);
        if (first_row_of_warp >= A.rows) break;
        const int i = base;
        double s = 0.0;
        if (i < A.rows) {
            const int pb = A.rp[i], pe = A.rp[i + 1];
            for (int p = pb + sub; p < pe; p += LANES) s += ld_stream(A.v + p) * __ldg(x + ld_stream(A.ci + p));
        }
        s = subwarp_sum<LANES>(s);
        if (i < A.rows && sub == 0) {
            if (ADD) s += y[i];
            y[i] = s;
            if (DOT) acc += w[i] * s;
        }
    }
    if (DOT) block_sum_to_partial(acc, partial);
}

// K7: level-0 direct solve as a dense symmetric GEMV with the precomputed inverse
__global__ void __launch_bounds__(256) k_dense_gemv(int n, const double *__restrict__ B, const double *__restrict__ x,
                                                    double *__restrict__ y, const int *done)
{
    if (done && *done) return;
    const int lane = threadIdx.x & 31;
    const int i = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (i >= n) return;
    const double *row = B + (size_t)i * n;
    double s = 0.0;
    for (int j = lane; j < n; j += 32) s += ld_stream(row + j) * __ldg(x + j);
    s = warp_sum(s);
    if (lane == 0) y[i] = s;
}

// ---- dense in-place Gauss-Jordan inversion of the SPD level-0 operator (setup only) ----
__global__ void k_gj_pivot(int n, int k, const double *__restrict__ a, double *__restrict__ rowk, double *__restrict__ colk)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double d = 1.0 / a[(size_t)k * n + k];
    rowk[j] = (j == k) ? d : a[(size_t)k * n + j] * d;
    colk[j] = a[(size_t)j * n + k];
}
__global__ void k_gj_update(int n, int k, double *__restrict__ a, const double *__restrict__ rowk, const double *__restrict__ colk)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= n) return;
    const double d = rowk[k];
    double *e = a + (size_t)i * n + j;
    if (i == k) *e = rowk[j];
    else if (j == k) *e = -colk[i] * d;
    else *e -= colk[i] * rowk[j];
}
__global__ void k_symmetrize(int n, double *__restrict__ a)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= n || j <= i) return;
    const double m = 0.5 * (a[(size_t)i * n + j] + a[(size_t)j * n + i]);
    a[(size_t)i * n + j] = m;
    a[(size_t)j * n + i] = m;
}
__global__ void k_csr_to_dense(CsrView A, int n, double *__restrict__ a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.rows) return;
    for (int p = A.rp[i]; p < A.rp[i + 1]; p++) a[(size_t)i * n + A.ci[p]] = A.v[p];
}

// ---- K8/K9 vector kernels ------------------------------------------------------------
// out[i] = in[perm[i]]  (reference numbering -> device numbering)
__global__ void k_gather(int n, const int *__restrict__ perm, const double *__restrict__ in, double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[perm[i]];
}
// out[perm[i]] = in[i]  (device numbering -> reference numbering)
__global__ void k_scatter(int n, const int *__restrict__ perm, const double *__restrict__ in, double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[perm[i]] = in[i];
}
__global__ void k_fill(int n, double *__restrict__ x, double val, const int *done)
{
    if (done && *done) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = val;
}
__global__ void k_extract_diag_inv(LvlView A, double *__restrict__ dinv)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < A.n) dinv[i] = 1.0 / A.v[A.dpos[i]];
}
// z = dinv .* r   (DIAG_PREC, PREP.h:393-401; MGPIS.h:192,206)
__global__ void k_jacobi(int n, const double *__restrict__ dinv, const double *__restrict__ r, double *__restrict__ z, const int *done)
{
    if (done && *done) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) z[i] = dinv[i] * r[i];
}
// partial[blockIdx] = sum a_i b_i  (grid-stride, fixed grid => deterministic)
__global__ void __launch_bounds__(256) k_dot(int n, const double *__restrict__ a, const double *__restrict__ b, double *partial, const int *done)
{
    if (done && *done) return;
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc += a[i] * b[i];
    block_sum_to_partial(acc, partial);
}
// x += alpha p ; r -= alpha q ; partial = r.r   (MGPIS.h:202-203 and the norm of :198)
__global__ void __launch_bounds__(256) k_update_xr(int n, const PcgState *st, const double *__restrict__ p, const double *__restrict__ q,
                                                   double *__restrict__ x, double *__restrict__ r, double *partial)
{
    if (st->done) return;
    const double alpha = st->alpha;
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * q[i];
        r[i] = ri;
        acc += ri * ri;
    }
    block_sum_to_partial(acc, partial);
}
// p = z + beta p   (MGPIS.h:214)
__global__ void __launch_bounds__(256) k_update_p(int n, const PcgState *st, const double *__restrict__ z, double *__restrict__ p)
{
    if (st->done) return;
    const double beta = st->beta;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = z[i] + beta * p[i];
}

// ---- scalar kernels: one warp finishes a reduction and advances the PCG state ----------
// after r = b: bb = b.b, tol = rel_tol*||b||, rr = bb ; loop condition of MGPIS.h:198 at it = 0
__global__ void k_s_init(PcgState *st, const double *partial, int np, double rel_tol, long long maxit)
{
    const double bb = warp_reduce_partials(partial, np);
    if (threadIdx.x == 0) {
        st->bb = bb; st->rr = bb; st->rel_tol = rel_tol; st->tol = rel_tol * sqrt(bb);
        st->it = 0; st->maxit = maxit; st->alpha = 0.0; st->beta = 0.0; st->delta_old = 0.0; st->delta_new = 0.0;
        st->done = !(0 < maxit && sqrt(bb) > st->tol);
    }
}
// delta_new = r.z after the first preconditioner application (MGPIS.h:197)
__global__ void k_s_delta0(PcgState *st, const double *partial, int np)
{
    if (st->done) return;
    const double d = warp_reduce_partials(partial, np);
    if (threadIdx.x == 0) st->delta_new = d;
}
// alpha = delta_new / (p.q)   (MGPIS.h:201)
__global__ void k_s_alpha(PcgState *st, const double *partial, int np)
{
    if (st->done) return;
    const double pq = warp_reduce_partials(partial, np);
    if (threadIdx.x == 0) { st->pq = pq; st->alpha = st->delta_new / pq; }
}
// rr = r.r
__global__ void k_s_rr(PcgState *st, const double *partial, int np)
{
    if (st->done) return;
    const double rr = warp_reduce_partials(partial, np);
    if (threadIdx.x == 0) st->rr = rr;
}
// delta_old = delta_new; delta_new = r.z; beta = delta_new/delta_old   (MGPIS.h:211-213)
__global__ void k_s_beta(PcgState *st, const double *partial, int np)
{
    if (st->done) return;
    const double rz = warp_reduce_partials(partial, np);
    if (threadIdx.x == 0) { st->delta_old = st->delta_new; st->delta_new = rz; st->beta = rz / st->delta_old; }
}
// end of an iteration: it++ and re-evaluate `it < maxit && ||r|| > tol` (MGPIS.h:219,198)
__global__ void k_s_next(PcgState *st)
{
    if (st->done) return;
    st->it += 1;
    if (!(st->it < st->maxit && sqrt(st->rr) > st->tol)) st->done = 1;
}

}  // namespace ddpca
