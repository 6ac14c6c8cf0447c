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

constexpr int kNumPart = 1184;   // 148 SMs x 8:   // slots of a partial-sum buffer (one per CTA of a reducing kernel)

// One multigrid level in its device layout: stage-permuted, stored per ROW GROUP.
// The <= 3 rows of a group share one column pattern (plan.h), so the pattern is stored once
// per group and each row only adds its values ("GCSR"):
//   meta[g]            : {row0, gs, cptr, len, voff, kd}
//   ci[cptr .. +len]   : pattern (sorted new column numbers), len padded to an even count
//   v[voff + r*len ..] : values of row r of the group (r < gs), same order as the pattern
//   pattern positions [0,kd) are strictly-lower couplings (earlier stages), [kd,kd+gs) the
//   in-group block (diagonal block), [kd+gs,len) strictly-upper couplings (later stages).
// HBM traffic per stored entry: 8 B value + 4/gs B index (9.33 B for the usual gs = 3)
// instead of CSR's 12 B.
struct __align__(16) GroupMeta {
    int row0, gs, cptr, len;
    long long voff;
    int kd, pad;
};
struct LvlView {
    int n, ng;
    const GroupMeta *__restrict__ meta;  // [ng]
    const int *__restrict__ ci;
    const double *__restrict__ v;
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

__device__ __forceinline__ double2 ld_stream2(const double *p)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ int2 ld_stream2(const int *p)
{
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ GroupMeta ld_meta(const GroupMeta *p)
{
    const int4 a = __ldg(reinterpret_cast<const int4 *>(p));
    const int4 b = __ldg(reinterpret_cast<const int4 *>(p) + 1);
    GroupMeta m;
    m.row0 = a.x; m.gs = a.y; m.cptr = a.z; m.len = a.w;
    m.voff = (long long)(((unsigned long long)(unsigned)b.y << 32) | (unsigned)b.x);
    m.kd = b.z; m.pad = b.w;
    return m;
}

// A row group is processed by a SUB-WARP of GL lanes (GL = 8: four groups per warp), so that
// enough groups -- i.e. enough independent (descriptor -> pattern/values -> x gather) load
// chains -- are in flight per SM to cover HBM latency; each lane streams 16-byte value loads.
#ifndef DDPCA_GROUP_LANES
#define DDPCA_GROUP_LANES 8
#endif
constexpr int GL = DDPCA_GROUP_LANES;
#ifndef DDPCA_MIN_BLOCKS
#define DDPCA_MIN_BLOCKS 3
#endif
#ifndef DDPCA_UNROLL
#define DDPCA_UNROLL 3
#endif
#define DDPCA_PRAGMA(x) _Pragma(#x)
#define DDPCA_UNROLL_LOOP(n) DDPCA_PRAGMA(unroll n)

template <int LN>
__device__ __forceinline__ unsigned subwarp_mask_t()
{
    const unsigned lane = threadIdx.x & 31u;
    return (LN == 32) ? 0xffffffffu : (((1u << (LN & 31)) - 1u) << (lane & ~(unsigned)(LN - 1)));
}
template <int LN>
__device__ __forceinline__ double group_sum_t(double v, unsigned mask)
{
#pragma unroll
    for (int o = LN / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}
__device__ __forceinline__ unsigned subwarp_mask() { return subwarp_mask_t<GL>(); }
__device__ __forceinline__ double group_sum(double v, unsigned mask) { return group_sum_t<GL>(v, mask); }

// Partial sums of one row group over the pattern range selected by LOWER / UPPER:
//   sL[r] = sum_{k <  kd}      a_r[k] x[c_k]      sU[r] = sum_{k >= kd+gs} a_r[k] x[c_k]
// Lane sl of the sub-warp owns pattern positions (2*sl, 2*sl+1) + 2*GL*it: one 8-byte index
// load, gs 16-byte value loads and one x gather pair per step.
template <bool LOWER, bool UPPER, bool NC_X, int LN = GL>
__device__ __forceinline__ void group_partial(const LvlView &A, const GroupMeta &m, const double *x, int sl,
                                              double (&sL)[3], double (&sU)[3])
{
    const int *ci = A.ci + m.cptr;
    const double *v = A.v + m.voff;
    const int len = m.len, kd = m.kd, ku = m.kd + m.gs, gs = m.gs;
    const int kbeg = LOWER ? 0 : (ku & ~1);
    const int kend = UPPER ? len : kd;   // exclusive; entries >= kend are never needed
    DDPCA_UNROLL_LOOP(DDPCA_UNROLL)
    for (int k = kbeg + 2 * sl; k < kend; k += 2 * LN) {
        const int2 c = ld_stream2(ci + k);
        double2 a[3];
#pragma unroll
        for (int r = 0; r < 3; r++)
            if (r < gs) a[r] = ld_stream2(v + (size_t)r * len + k);
        const bool l0 = k < kd, l1 = k + 1 < kd, u0 = k >= ku, u1 = k + 1 >= ku;
        double x0 = 0.0, x1 = 0.0;
        if ((LOWER && l0) || (UPPER && u0)) x0 = NC_X ? __ldg(x + c.x) : x[c.x];
        if ((LOWER && l1) || (UPPER && u1)) x1 = NC_X ? __ldg(x + c.y) : x[c.y];
        const double xl0 = (LOWER && l0) ? x0 : 0.0, xl1 = (LOWER && l1) ? x1 : 0.0;
        const double xu0 = (UPPER && u0) ? x0 : 0.0, xu1 = (UPPER && u1) ? x1 : 0.0;
#pragma unroll
        for (int r = 0; r < 3; r++)
            if (r < gs) {
                if (LOWER) sL[r] += a[r].x * xl0 + a[r].y * xl1;
                if (UPPER) sU[r] += a[r].x * xu0 + a[r].y * xu1;
            }
    }
}

// ------------------------------------------------------------------------------------
// K3: forward Gauss-Seidel relaxation of one row group (MGPIS.h:66-72 on the permuted level)
//   x_i   = (b_i - sum_{j<i} a_ij x_j(new) - sum_{j>i} a_ij x_j(old)) / a_ii
//   p1_i  = a_ii x_i + sum_{j>i} a_ij x_j(old)            ( = b - L x, MGPIS.h:72 )
// ZERO_X: x is known to be zero on entry (first smoothing of a V-cycle, MGPIS.h:93,204):
//         the strictly-upper half is not read at all.
// ------------------------------------------------------------------------------------
template <bool ZERO_X, bool NC_X, int LN = GL>
__device__ __forceinline__ void group_fwd(const LvlView &A, int g, const double *__restrict__ b,
                                          double *x, double *__restrict__ p1, int sl, unsigned mask)
{
    const GroupMeta m = ld_meta(A.meta + g);
    const int gs = m.gs, r0 = m.row0;
    double sL[3] = {0.0, 0.0, 0.0}, sU[3] = {0.0, 0.0, 0.0};
    double blk[3][3], bb[3], xo[3];
    const double *vb = A.v + m.voff + m.kd;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        bb[r] = 0.0; xo[r] = 0.0;
        if (r < gs) {
            bb[r] = b[r0 + r];
            if (!ZERO_X) xo[r] = x[r0 + r];
        }
#pragma unroll
        for (int c = 0; c < 3; c++) blk[r][c] = (r < gs && c < gs) ? __ldg(vb + (size_t)r * m.len + c) : 0.0;
    }
    group_partial<true, !ZERO_X, NC_X, LN>(A, m, x, sl, sL, sU);
#pragma unroll
    for (int r = 0; r < 3; r++) {
        sL[r] = group_sum_t<LN>(sL[r], mask);
        if (!ZERO_X) sU[r] = group_sum_t<LN>(sU[r], mask);
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
                if (c > r) inU += blk[r][c] * xo[c];
            }
            const double up = sU[r] + inU;
            xn[r] = (bb[r] - sL[r] - inL - up) / blk[r][r];
            if (sl == 0) {
                x[r0 + r] = xn[r];
                p1[r0 + r] = blk[r][r] * xn[r] + up;
            }
        }
    }
}

// K4: backward relaxation of one row group (MGPIS.h:73-76):  x_i = (p1_i - sum_{j>i} a_ij x_j) / a_ii
template <bool NC_X, int LN = GL>
__device__ __forceinline__ void group_bwd(const LvlView &A, int g, const double *__restrict__ p1,
                                          double *x, int sl, unsigned mask)
{
    const GroupMeta m = ld_meta(A.meta + g);
    const int gs = m.gs, r0 = m.row0;
    double sL[3] = {0.0, 0.0, 0.0}, sU[3] = {0.0, 0.0, 0.0};
    double blk[3][3], pp[3];
    const double *vb = A.v + m.voff + m.kd;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        pp[r] = (r < gs) ? p1[r0 + r] : 0.0;
#pragma unroll
        for (int c = 0; c < 3; c++) blk[r][c] = (r < gs && c < gs && c >= r) ? __ldg(vb + (size_t)r * m.len + c) : 0.0;
    }
    group_partial<false, true, NC_X, LN>(A, m, x, sl, sL, sU);
#pragma unroll
    for (int r = 0; r < 3; r++) sU[r] = group_sum_t<LN>(sU[r], mask);
    double xn[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int r = 2; r >= 0; r--) {
        if (r < gs) {
            double inU = 0.0;
#pragma unroll
            for (int c = 0; c < 3; c++)
                if (c > r) inU += blk[r][c] * xn[c];
            xn[r] = (pp[r] - sU[r] - inU) / blk[r][r];
            if (sl == 0) x[r0 + r] = xn[r];
        }
    }
}

// one stage = groups [g0,g1): mutually independent, one sub-warp per group
// (LN lanes per group: GL for multigrid levels, 32 for the long fill-in rows of triangular factors)
template <bool ZERO_X, int LN = GL>
__global__ void __launch_bounds__(256, DDPCA_MIN_BLOCKS) k_sweep_fwd_stage(LvlView A, int g0, int g1, const double *__restrict__ b,
                                                         double *x, double *__restrict__ p1, const int *done)
{
    if (done && *done) return;
    const int g = g0 + (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) / LN);
    if (g >= g1) return;
    group_fwd<ZERO_X, true, LN>(A, g, b, x, p1, threadIdx.x % LN, subwarp_mask_t<LN>());
}

template <int LN = GL>
__global__ void __launch_bounds__(256, DDPCA_MIN_BLOCKS) k_sweep_bwd_stage(LvlView A, int g0, int g1, const double *__restrict__ p1,
                                                         double *x, const int *done)
{
    if (done && *done) return;
    const int g = g0 + (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) / LN);
    if (g >= g1) return;
    group_bwd<true, LN>(A, g, p1, x, threadIdx.x % LN, subwarp_mask_t<LN>());
}

// a run of small stages [s0,s1) relaxed by ONE CTA, __syncthreads() between stages
// (latency-bound regime: LEX wavefronts, coarse levels, triangular solves)
// LN lanes per row group: 8 for multigrid levels (rows of <= 81 entries), 32 for the triangular
// factors of the direct solves (fill-in rows of thousands of entries on a sequential dependency chain)
template <bool ZERO_X, int LN>
__global__ void __launch_bounds__(512) k_sweep_fwd_multi(LvlView A, const int *__restrict__ stage_group, int s0, int s1,
                                                          const double *__restrict__ b, double *x, double *p1, const int *done)
{
    if (done && *done) return;
    const int sl = threadIdx.x % LN, w = threadIdx.x / LN, nw = blockDim.x / LN;
    const unsigned mask = subwarp_mask_t<LN>();
    for (int s = s0; s < s1; s++) {
        const int ga = stage_group[s], gb = stage_group[s + 1];
        for (int g = ga + w; g < gb; g += nw) group_fwd<ZERO_X, false, LN>(A, g, b, x, p1, sl, mask);
        __syncthreads();
    }
}

template <int LN>
__global__ void __launch_bounds__(512) k_sweep_bwd_multi(LvlView A, const int *__restrict__ stage_group, int s0, int s1,
                                                          const double *p1, double *x, const int *done)
{
    if (done && *done) return;
    const int sl = threadIdx.x % LN, w = threadIdx.x / LN, nw = blockDim.x / LN;
    const unsigned mask = subwarp_mask_t<LN>();
    for (int s = s1 - 1; s >= s0; s--) {
        const int ga = stage_group[s], gb = stage_group[s + 1];
        for (int g = ga + w; g < gb; g += nw) group_bwd<false, LN>(A, g, p1, x, sl, mask);
        __syncthreads();
    }
}

// Triangular solve with a unit-diagonal factor stored as a LEX-staged level of single-row groups
// (I+L forward, I+L^T backward): one CTA of 32 warps walks the dependency wavefronts.  Stages with
// many rows give each warp its own rows; stages with few rows -- the long fill-in rows at the end of
// the elimination, which sit on ONE sequential dependency chain -- split every row over several
// warps, so a 3000-entry row costs one load round trip instead of fifty.
//   FWD: x_i = b_i - sum_{j<i} l_ij x_j          BWD: x_i = b_i - sum_{j>i} l_ji x_j
template <bool FWD>
__global__ void __launch_bounds__(1024) k_tri_multi(LvlView A, const int *__restrict__ stage_group, int s0, int s1,
                                                     const double *__restrict__ b, double *x, const int *done)
{
    if (done && *done) return;
    __shared__ double part[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int si = 0; si < s1 - s0; si++) {
        const int s = FWD ? (s0 + si) : (s1 - 1 - si);
        const int ga = stage_group[s], ng = stage_group[s + 1] - ga;
        if (ng >= nw) {
            for (int g = ga + w; g < ga + ng; g += nw) {
                const GroupMeta m = ld_meta(A.meta + g);
                const int kb = FWD ? 0 : ((m.kd + 1) & ~1), ke = FWD ? m.kd : m.len;
                const int *ci = A.ci + m.cptr;
                const double *v = A.v + m.voff;
                double acc = 0.0;
#pragma unroll 2
                for (int k = kb + 2 * lane; k < ke; k += 64) {
                    const int2 c = ld_stream2(ci + k);
                    const double2 a = ld_stream2(v + k);
                    if (FWD) { acc += a.x * x[c.x]; if (k + 1 < ke) acc += a.y * x[c.y]; }
                    else { if (k > m.kd) acc += a.x * x[c.x]; acc += a.y * x[c.y]; }
                }
                acc = warp_sum(acc);
                if (lane == 0) x[m.row0] = b[m.row0] - acc;
            }
            __syncthreads();
        } else {
            // wpg warps per row
            int wpg = 1;
            while (wpg * 2 * ng <= nw) wpg *= 2;
            const int gi = w / wpg, slice = w % wpg;
            double acc = 0.0;
            GroupMeta m;
            m.row0 = 0;
            if (gi < ng) {
                m = ld_meta(A.meta + ga + gi);
                const int kb = FWD ? 0 : ((m.kd + 1) & ~1), ke = FWD ? m.kd : m.len;
                const int *ci = A.ci + m.cptr;
                const double *v = A.v + m.voff;
#pragma unroll 2
                for (int k = kb + 2 * (slice * 32 + lane); k < ke; k += 64 * wpg) {
                    const int2 c = ld_stream2(ci + k);
                    const double2 a = ld_stream2(v + k);
                    if (FWD) { acc += a.x * x[c.x]; if (k + 1 < ke) acc += a.y * x[c.y]; }
                    else { if (k > m.kd) acc += a.x * x[c.x]; acc += a.y * x[c.y]; }
                }
                acc = warp_sum(acc);
            }
            if (lane == 0) part[w] = acc;
            __syncthreads();
            if (gi < ng && slice == 0 && lane == 0) {
                double t = 0.0;
                for (int q = 0; q < wpg; q++) t += part[gi * wpg + q];
                x[m.row0] = b[m.row0] - t;
            }
            __syncthreads();
        }
    }
}

// One WIDE wavefront of such a triangular solve: one CTA of 128 threads per row.  The fill-in rows of a factor hold
// hundreds to thousands of entries and every wavefront is one step of a dependency chain, so what counts is the
// latency of a row, not occupancy: 128 threads stream a 2 000-entry row in 8 steps where one warp needs 31.
template <bool FWD>
__global__ void __launch_bounds__(128) k_tri_stage(LvlView A, int g0, const double *__restrict__ b, double *x, const int *done)
{
    if (done && *done) return;
    __shared__ double part[4];
    const GroupMeta m = ld_meta(A.meta + g0 + blockIdx.x);
    const int kb = FWD ? 0 : ((m.kd + 1) & ~1), ke = FWD ? m.kd : m.len;
    const int *ci = A.ci + m.cptr;
    const double *v = A.v + m.voff;
    double acc = 0.0;
#pragma unroll 2
    for (int k = kb + 2 * (int)threadIdx.x; k < ke; k += 256) {
        const int2 c = ld_stream2(ci + k);
        const double2 a = ld_stream2(v + k);
        if (FWD) { acc += a.x * x[c.x]; if (k + 1 < ke) acc += a.y * x[c.y]; }
        else { if (k > m.kd) acc += a.x * x[c.x]; acc += a.y * x[c.y]; }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) x[m.row0] = b[m.row0] - ((part[0] + part[1]) + (part[2] + part[3]));
}

// K2: r = b - (p1 + L x)   (MGPIS.h:92) -- strictly-lower half only, one sub-warp per group
__global__ void __launch_bounds__(256, DDPCA_MIN_BLOCKS) k_resid_lower(LvlView A, const double *__restrict__ b, const double *__restrict__ p1,
                                                     const double *__restrict__ x, double *__restrict__ r, const int *done)
{
    if (done && *done) return;
    const int sl = threadIdx.x % GL;
    const int g = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) / GL);
    if (g >= A.ng) return;
    const unsigned mask = subwarp_mask();
    const GroupMeta m = ld_meta(A.meta + g);
    double sL[3] = {0.0, 0.0, 0.0}, sU[3] = {0.0, 0.0, 0.0};
    group_partial<true, false, true>(A, m, x, sl, sL, sU);
#pragma unroll
    for (int q = 0; q < 3; q++) sL[q] = group_sum(sL[q], mask);
    if (sl < m.gs) {
        const int i = m.row0 + sl;
        double s = sl == 0 ? sL[0] : (sl == 1 ? sL[1] : sL[2]);
        const double *vb = A.v + m.voff + (size_t)sl * m.len + m.kd;
        for (int c = 0; c < sl; c++) s += vb[c] * x[m.row0 + c];   // in-group strictly-lower part
        r[i] = b[i] - (p1[i] + s);
    }
}

// K1: y = A x on the group layout, one sub-warp per group, grid-stride; DOT: partial[blockIdx] = sum_i w_i y_i
// (the p.q of MGPIS.h:201 fused into the product of :200).  Fixed grid => deterministic.
template <bool DOT>
__global__ void __launch_bounds__(256, DDPCA_MIN_BLOCKS) k_spmv_group(LvlView A, const double *__restrict__ x, double *__restrict__ y,
                                                    const double *__restrict__ w, double *partial, const int *done)
{
    if (done && *done) return;
    const int sl = threadIdx.x % GL;
    const unsigned mask = subwarp_mask();
    const int nsub = (int)((gridDim.x * (unsigned)blockDim.x) / GL);
    double acc = 0.0;
    for (int g = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) / GL); g < A.ng; g += nsub) {
        const GroupMeta m = ld_meta(A.meta + g);
        const int *ci = A.ci + m.cptr;
        const double *v = A.v + m.voff;
        double s[3] = {0.0, 0.0, 0.0};
        DDPCA_UNROLL_LOOP(DDPCA_UNROLL)
        for (int k = 2 * sl; k < m.len; k += 2 * GL) {
            const int2 c = ld_stream2(ci + k);
            double2 a[3];
#pragma unroll
            for (int r = 0; r < 3; r++)
                if (r < m.gs) a[r] = ld_stream2(v + (size_t)r * m.len + k);
            const double x0 = __ldg(x + c.x), x1 = __ldg(x + c.y);
#pragma unroll
            for (int r = 0; r < 3; r++)
                if (r < m.gs) s[r] += a[r].x * x0 + a[r].y * x1;
        }
#pragma unroll
        for (int r = 0; r < 3; r++) s[r] = group_sum(s[r], mask);
        if (sl < m.gs) {
            const double yi = sl == 0 ? s[0] : (sl == 1 ? s[1] : s[2]);
            y[m.row0 + sl] = yi;
            if (DOT) acc += w[m.row0 + sl] * yi;
        }
    }
    if (DOT) block_sum_to_partial(acc, partial);
}

// K5/K6 and the ADMM interface operators: y (=|+=) alpha A x, LANES lanes per row; DOT: partial[blockIdx] = sum_i w_i * y_i
// (fused p.q of MGPIS.h:201).  Grid-stride so that DOT stays deterministic for a fixed grid.
template <int LANES, bool ADD, bool DOT>
__global__ void __launch_bounds__(256) k_spmv(CsrView A, double alpha, const double *__restrict__ x, double *y,
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
            s *= alpha;
            if (ADD) s += y[i];
            y[i] = s;
            if (DOT) acc += w[i] * s;
        }
    }
    if (DOT) block_sum_to_partial(acc, partial);
}

// K5/K6 on the NODE structure of the transfer operators.  realProl couples fine node i to coarse node j with one
// scalar weight for all three displacement components unless the nodes carry rotations (MULTIGRID.h:1143-1176), so
// after the explicit zeros are dropped the three rows of a node are shifted copies of each other:
// cols(row k) = cols(row 0) + k, same values.  Such row triples are stored ONCE (first column + weight per node
// pair: 12 B instead of 36 B of CSR); rows that do not fit (constrained or rotated nodes) stay in a small CSR
// remainder.  y (=|+=) T x for the triples, LANES lanes per triple (1-2 for prolongation rows of ~3 pairs, 8 for
// restriction rows of ~27).
template <int LANES, bool ADD>
__global__ void __launch_bounds__(256) k_trip_spmv(int ntrip, const int *__restrict__ trow, const int *__restrict__ tptr,
                                                   const int *__restrict__ tcol, const double *__restrict__ tw,
                                                   const double *__restrict__ x, double *y, const int *done)
{
    if (done && *done) return;
    const int sub = threadIdx.x % LANES;
    const int g = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) / LANES);
    const bool live = g < ntrip;   // whole sub-warps are live or not (256 % LANES == 0)
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    int r = 0;
    if (live) {
        r = trow[g];
        const int pb = tptr[g], pe = tptr[g + 1];
        for (int p = pb + sub; p < pe; p += LANES) {
            const int c = ld_stream(tcol + p);
            const double w = ld_stream(tw + p);
            s0 += w * __ldg(x + c);
            s1 += w * __ldg(x + c + 1);
            s2 += w * __ldg(x + c + 2);
        }
    }
    s0 = subwarp_sum<LANES>(s0); s1 = subwarp_sum<LANES>(s1); s2 = subwarp_sum<LANES>(s2);
    if (live && sub == 0) {
        if (ADD) { y[r] += s0; y[r + 1] += s1; y[r + 2] += s2; }
        else { y[r] = s0; y[r + 1] = s1; y[r + 2] = s2; }
    }
}
// remainder rows (compact CSR + row map): y[row[k]] (=|+=) sum_p v_p x[c_p], one thread per row
template <bool ADD>
__global__ void __launch_bounds__(256) k_rowmap_spmv(int nrows, const int *__restrict__ row, CsrView A, const double *__restrict__ x, double *y, const int *done)
{
    if (done && *done) return;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nrows) return;
    double s = 0.0;
    for (int p = A.rp[k]; p < A.rp[k + 1]; p++) s += A.v[p] * __ldg(x + A.ci[p]);
    const int i = row[k];
    if (ADD) y[i] += s; else y[i] = s;
}

// K7: level-0 direct solve as a dense symmetric GEMV with the precomputed inverse.  One warp per row;
// 16-byte streaming loads, four independent partial sums per lane (eight loads in flight) -- the rows
// are short (a few thousand entries), so memory-level parallelism per warp is what sets the rate.
__global__ void __launch_bounds__(256) k_dense_gemv(int n, const double *__restrict__ B, const double *__restrict__ x,
                                                    double *__restrict__ y, const int *done)
{
    if (done && *done) return;
    const int lane = threadIdx.x & 31;
    const int i = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (i >= n) return;
    const double *row = B + (size_t)i * n;
    double s = 0.0;
    if ((n & 1) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0) {
        const int n2 = n >> 1;
        const double2 *x2 = reinterpret_cast<const double2 *>(x);
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int j = lane;
        for (; j + 96 < n2; j += 128) {
            const double2 a0 = ld_stream2(row + 2 * j), a1 = ld_stream2(row + 2 * (j + 32));
            const double2 a2 = ld_stream2(row + 2 * (j + 64)), a3 = ld_stream2(row + 2 * (j + 96));
            const double2 b0 = __ldg(x2 + j), b1 = __ldg(x2 + j + 32), b2 = __ldg(x2 + j + 64), b3 = __ldg(x2 + j + 96);
            s0 += a0.x * b0.x + a0.y * b0.y;
            s1 += a1.x * b1.x + a1.y * b1.y;
            s2 += a2.x * b2.x + a2.y * b2.y;
            s3 += a3.x * b3.x + a3.y * b3.y;
        }
        for (; j < n2; j += 32) {
            const double2 a0 = ld_stream2(row + 2 * j);
            const double2 b0 = __ldg(x2 + j);
            s0 += a0.x * b0.x + a0.y * b0.y;
        }
        s = (s0 + s1) + (s2 + s3);
    } else {
        for (int j = lane; j < n; j += 32) s += ld_stream(row + j) * __ldg(x + j);
    }
    s = warp_sum(s);
    if (lane == 0) y[i] = s;
}

// ---- dense tail of a sparse LDL^T factor (setup): M = L22 diag(D2) L22^T, 32x32 tiles ----------
__global__ void __launch_bounds__(1024) k_ldl_tail_product(int T, const double *__restrict__ L22, const double *__restrict__ D2, double *__restrict__ M)
{
    __shared__ double sa[32][33], sb[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i = blockIdx.y * 32 + ty, j = blockIdx.x * 32 + tx;
    if (blockIdx.x > blockIdx.y) return;   // lower triangle of tiles; mirrored below
    double acc = 0.0;
    const int kmax = min(blockIdx.x * 32 + 32, T);   // L22 is lower triangular: k <= j
    for (int k0 = 0; k0 < kmax; k0 += 32) {
        const int ia = blockIdx.y * 32 + ty, ka = k0 + tx;
        sa[ty][tx] = (ia < T && ka < T) ? L22[(size_t)ia * T + ka] * D2[ka] : 0.0;
        const int jb = blockIdx.x * 32 + ty, kb = k0 + tx;
        sb[ty][tx] = (jb < T && kb < T) ? L22[(size_t)jb * T + kb] : 0.0;
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; k++) acc += sa[ty][k] * sb[tx][k];
        __syncthreads();
    }
    if (i < T && j < T) {
        M[(size_t)i * T + j] = acc;
        M[(size_t)j * T + i] = acc;
    }
}
// tail right-hand side: out[i - n1] = b[i] - sum_{k < k1[i-n1]} l_ik y_k   (couplings to the sparse part)
__global__ void __launch_bounds__(256) k_tail_rhs(LvlView A, int g0, int T, const int *__restrict__ k1, const double *__restrict__ b,
                                                   const double *__restrict__ y, double *__restrict__ out, const int *done)
{
    if (done && *done) return;
    const int lane = threadIdx.x & 31;
    const int t = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (t >= T) return;
    const GroupMeta m = ld_meta(A.meta + g0 + t);
    const int *ci = A.ci + m.cptr;
    const double *v = A.v + m.voff;
    const int ke = k1[t];
    double acc = 0.0;
    for (int k = 2 * lane; k < ke; k += 64) {
        const int2 c = ld_stream2(ci + k);
        const double2 a = ld_stream2(v + k);
        acc += a.x * y[c.x];
        if (k + 1 < ke) acc += a.y * y[c.y];
    }
    acc = warp_sum(acc);
    if (lane == 0) out[t] = b[m.row0] - acc;
}

// ---- dense in-place inversion of an SPD operator (set-up only): BLOCKED Gauss-Jordan ------------------
// Panel K = [k0, k0+nb), nb <= 64 (no pivoting: SPD).  With D = A[K,K], C = A[:,K], R = D^-1 A[K,:]:
//   A[i,j] -= C[i,:] R[:,j]  (i,j not in K)     A[K,j] = R[:,j]     A[i,K] = -C[i,:] D^-1     A[K,K] = D^-1
// -- the scalar elimination step with matrices for numbers.  Three launches per panel instead of two per
// ROW, and the trailing update is a tiled FP64 product (64x64 tiles, 4x4 register blocks) instead of a
// rank-1 sweep over the whole matrix: n = 9 216 moves 1.6 TFLOP through DFMA pipes instead of 12 TB through HBM.
constexpr int kGjB = 64;
// (a) Dinv = A[K,K]^-1, one CTA, scalar Gauss-Jordan in shared memory
__global__ void __launch_bounds__(1024) k_bgj_diag(int n, int k0, int nb, const double *__restrict__ A, double *__restrict__ Dinv)
{
    // The 64 x 64 block lives in registers (4 entries per thread: rows ib + 16 m of column j); a step needs the pivot
    // row, the pivot column and the reciprocal pivot, which the owners of row / column k+1 publish while they apply
    // step k -- double-buffered, ONE barrier per step (three barriers and a round trip through shared memory per
    // step made this kernel half of the inversion time of small interface matrices).  A partial last block is
    // padded with the identity.
    __shared__ double rowb[2][kGjB], colb[2][kGjB], pivb[2];
    const int j = threadIdx.x & (kGjB - 1), ib = threadIdx.x >> 6;
    double v[4];
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const int i = ib + 16 * m;
        v[m] = (i < nb && j < nb) ? A[(size_t)(k0 + i) * n + k0 + j] : (i == j ? 1.0 : 0.0);
        if (i == 0) rowb[0][j] = v[m];
        if (j == 0) { colb[0][i] = v[m]; if (i == 0) pivb[0] = 1.0 / v[m]; }
    }
    __syncthreads();
    for (int k = 0; k < nb; k++) {
        const int cur = k & 1, nxt = cur ^ 1;
        const double piv = pivb[cur];
        const double rj = (j == k) ? piv : rowb[cur][j] * piv;
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const int i = ib + 16 * m;
            const double c = colb[cur][i];
            const double nv = (i == k) ? rj : ((j == k) ? -c * piv : v[m] - c * rj);
            v[m] = nv;
            if (i == k + 1) rowb[nxt][j] = nv;
            if (j == k + 1) { colb[nxt][i] = nv; if (i == k + 1) pivb[nxt] = 1.0 / nv; }
        }
        __syncthreads();
    }
#pragma unroll
    for (int m = 0; m < 4; m++) Dinv[(ib + 16 * m) * kGjB + j] = v[m];
}
// 64x64 (x64) tile product helper: acc[4][4] of thread (ty, tx) in a 16x16 thread block:
// rows ty*4 + r, columns tx + 16*c of  As (64 x kk, row-major in shared) times Bs (kk x 64).  The strided column
// assignment makes the 16 lanes of a row read 16 CONSECUTIVE doubles of Bs (no bank conflicts; tx*4 + c gave a 4-way
// conflict on every Bs load and held the kernel at ~3 TFLOP/s); the As reads are broadcasts.
__device__ __forceinline__ int gj_col(int tx, int c) { return tx + 16 * c; }
__device__ __forceinline__ void gj_tile_mma(const double (*As)[kGjB + 1], const double (*Bs)[kGjB + 1], int kk, int ty, int tx, double (&acc)[4][4])
{
#pragma unroll 4
    for (int k = 0; k < kk; k++) {
        double a[4], b[4];
#pragma unroll
        for (int r = 0; r < 4; r++) a[r] = As[ty * 4 + r][k];
#pragma unroll
        for (int c = 0; c < 4; c++) b[c] = Bs[k][gj_col(tx, c)];
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[r][c] += a[r] * b[c];
    }
}
// (b) C = A[:,K] (n x 64, ld 64) and R = Dinv A[K,:] (64 x n, ld n); one CTA per 64 columns / rows
__global__ void __launch_bounds__(256) k_bgj_panels(int n, int k0, int nb, const double *__restrict__ A, const double *__restrict__ Dinv,
                                                    double *__restrict__ C, double *__restrict__ R)
{
    extern __shared__ double gj_sm[];
    double (*As)[kGjB + 1] = reinterpret_cast<double (*)[kGjB + 1]>(gj_sm);
    double (*Bs)[kGjB + 1] = As + kGjB;
    const int t0 = blockIdx.x * kGjB, tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    // column panel copy: rows t0..t0+63
    for (int e = threadIdx.x; e < kGjB * kGjB; e += 256) {
        const int i = t0 + e / kGjB, c = e % kGjB;
        if (i < n) C[(size_t)i * kGjB + c] = (c < nb) ? A[(size_t)i * n + k0 + c] : 0.0;
    }
    // R[:, t0..t0+63] = Dinv (nb x nb) * A[K, t0..]
    for (int e = threadIdx.x; e < kGjB * kGjB; e += 256) {
        const int r = e / kGjB, c = e % kGjB;
        As[r][c] = (r < nb && c < nb) ? Dinv[r * kGjB + c] : 0.0;
        Bs[r][c] = (r < nb && t0 + c < n) ? A[(size_t)(k0 + r) * n + t0 + c] : 0.0;
    }
    __syncthreads();
    double acc[4][4] = {};
    gj_tile_mma(As, Bs, nb, ty, tx, acc);
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int i = ty * 4 + r, j = t0 + gj_col(tx, c);
            if (i < nb && j < n) R[(size_t)i * n + j] = acc[r][c];
        }
}
// (c) the update of every 64x64 tile of A
__global__ void __launch_bounds__(256) k_bgj_update(int n, int k0, int nb, double *__restrict__ A, const double *__restrict__ Dinv,
                                                    const double *__restrict__ C, const double *__restrict__ R)
{
    extern __shared__ double gj_sm[];
    double (*As)[kGjB + 1] = reinterpret_cast<double (*)[kGjB + 1]>(gj_sm);
    double (*Bs)[kGjB + 1] = As + kGjB;
    const int i0 = blockIdx.y * kGjB, j0 = blockIdx.x * kGjB, tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const bool rowK = (i0 == k0), colK = (j0 == k0);
    if (rowK && colK) {
        for (int e = threadIdx.x; e < nb * nb; e += 256) A[(size_t)(k0 + e / nb) * n + k0 + e % nb] = Dinv[(e / nb) * kGjB + e % nb];
        return;
    }
    if (rowK) {
        for (int e = threadIdx.x; e < kGjB * kGjB; e += 256) {
            const int r = e / kGjB, j = j0 + e % kGjB;
            if (r < nb && j < n) A[(size_t)(k0 + r) * n + j] = R[(size_t)r * n + j];
        }
        return;
    }
    // As = C[i0.., :]   Bs = colK ? Dinv : R[:, j0..]
    for (int e = threadIdx.x; e < kGjB * kGjB; e += 256) {
        const int r = e / kGjB, c = e % kGjB;
        As[r][c] = (i0 + r < n && c < nb) ? C[(size_t)(i0 + r) * kGjB + c] : 0.0;
        if (colK) Bs[r][c] = (r < nb && c < nb) ? Dinv[r * kGjB + c] : 0.0;
        else Bs[r][c] = (r < nb && j0 + c < n) ? R[(size_t)r * n + j0 + c] : 0.0;
    }
    __syncthreads();
    double acc[4][4] = {};
    gj_tile_mma(As, Bs, nb, ty, tx, acc);
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int i = i0 + ty * 4 + r, j = j0 + gj_col(tx, c);
            if (i >= n) continue;
            if (colK) { if (gj_col(tx, c) < nb) A[(size_t)i * n + j] = -acc[r][c]; }
            else if (j < n) A[(size_t)i * n + j] -= acc[r][c];
        }
}
// dense tail of a sparse LDL^T factor (set-up): M = L22 diag(D2) L22^T on 64x64 tiles with the 4x4 register blocks of the
// inversion kernels.  L22 is unit lower triangular: only k-tiles up to the column tile contribute; lower tiles are
// computed and mirrored.
__global__ void __launch_bounds__(256) k_ldl_tail_product64(int T, const double *__restrict__ L22, const double *__restrict__ D2, double *__restrict__ M)
{
    if (blockIdx.x > blockIdx.y) return;   // lower triangle of tiles
    extern __shared__ double gj_sm[];
    double (*As)[kGjB + 1] = reinterpret_cast<double (*)[kGjB + 1]>(gj_sm);
    double (*Bs)[kGjB + 1] = As + kGjB;
    const int i0 = blockIdx.y * kGjB, j0 = blockIdx.x * kGjB, tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4] = {};
    const int kmax = min(j0 + kGjB, T);
    for (int k0 = 0; k0 < kmax; k0 += kGjB) {
        for (int e = threadIdx.x; e < kGjB * kGjB; e += 256) {
            const int r = e / kGjB, c = e % kGjB;
            const int k = k0 + c;
            As[r][c] = (i0 + r < T && k < T) ? L22[(size_t)(i0 + r) * T + k] * D2[k] : 0.0;      // (L22 D)[i, k]
            Bs[c][r] = (j0 + r < T && k < T) ? L22[(size_t)(j0 + r) * T + k] : 0.0;              // L22^T[k, j]
        }
        __syncthreads();
        gj_tile_mma(As, Bs, kGjB, ty, tx, acc);
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int i = i0 + ty * 4 + r, j = j0 + gj_col(tx, c);
            if (i < T && j < T) {
                M[(size_t)i * T + j] = acc[r][c];
                if (blockIdx.x != blockIdx.y) M[(size_t)j * T + i] = acc[r][c];
            }
        }
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
// dense copy of the diagonal block [r0, r0+n) x [r0, r0+n) of a CSR operator (a zero-initialised)
__global__ void k_csr_to_dense(CsrView A, int r0, int n, double *__restrict__ a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int p = A.rp[r0 + i]; p < A.rp[r0 + i + 1]; p++) {
        const int c = A.ci[p] - r0;
        if (c >= 0 && c < n) a[(size_t)i * n + c] = A.v[p];
    }
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
// out[idx[i]] = in[i] * scale[i]
__global__ void k_scatter_scaled(int n, const int *__restrict__ idx, const double *__restrict__ scale, const double *__restrict__ in, double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[idx[i]] = in[i] * scale[i];
}
__global__ void k_fill(int n, double *__restrict__ x, double val, const int *done)
{
    if (done && *done) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = val;
}
__global__ void k_extract_diag_inv(LvlView A, double *__restrict__ dinv)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= A.ng) return;
    const GroupMeta m = A.meta[g];
    for (int r = 0; r < m.gs; r++) dinv[m.row0 + r] = 1.0 / A.v[m.voff + (size_t)r * m.len + m.kd + r];
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
// alpha = delta_new / (p.q) ; x += alpha p ; r -= alpha q ; partial = r.r   (MGPIS.h:201-203 and the norm of :198)
// Every CTA finishes the p.q reduction itself (same fixed order => same alpha everywhere); CTA 0 records it.
__global__ void __launch_bounds__(256) k_update_xr(int n, PcgState *st, const double *__restrict__ pq_partial, int npq,
                                                   const double *__restrict__ p, const double *__restrict__ q,
                                                   double *__restrict__ x, double *__restrict__ r, double *partial)
{
    if (st->done) return;
    __shared__ double s_alpha;
    if (threadIdx.x < 32) {
        const double pq = warp_reduce_partials(pq_partial, npq);
        if (threadIdx.x == 0) {
            s_alpha = st->delta_new / pq;
            if (blockIdx.x == 0) { st->pq = pq; st->alpha = s_alpha; }
        }
    }
    __syncthreads();
    const double alpha = s_alpha;
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
// rel_tol / maxit of the coming solve (kept out of the captured graphs)
__global__ void k_s_params(PcgState *st, double rel_tol, long long maxit)
{
    st->rel_tol = rel_tol;
    st->maxit = maxit;
}
__global__ void k_s_init(PcgState *st, const double *partial, int np)
{
    const double bb = warp_reduce_partials(partial, np);
    if (threadIdx.x == 0) {
        st->bb = bb; st->rr = bb; st->tol = st->rel_tol * sqrt(bb);
        st->it = 0; st->alpha = 0.0; st->beta = 0.0; st->delta_old = 0.0; st->delta_new = 0.0;
        st->done = !(0 < st->maxit && sqrt(bb) > st->tol);
    }
}
// delta_new = r.z after the first preconditioner application (MGPIS.h:197)
// `cond`: conditional handle of the WHILE node that repeats the CG iteration (device-side loop
// control, one graph launch per solve); ignored when use_cond == 0.
__global__ void k_s_delta0(PcgState *st, const double *partial, int np, cudaGraphConditionalHandle cond, int use_cond)
{
    if (!st->done) {
        const double d = warp_reduce_partials(partial, np);
        if (threadIdx.x == 0) st->delta_new = d;
    }
    if (use_cond && threadIdx.x == 0) cudaGraphSetConditional(cond, st->done ? 0u : 1u);
}
// rr = r.r ; delta_old = delta_new ; delta_new = r.z ; beta = delta_new/delta_old ; it++ and the loop
// condition `it < maxit && ||r|| > tol` (MGPIS.h:211-213,219,198) -- one launch, two warps reduce.
// p = z + beta p (k_update_p) runs after it and is skipped once done is set: p is dead by then.
__global__ void k_s_beta_next(PcgState *st, const double *rr_partial, const double *rz_partial, int np,
                              cudaGraphConditionalHandle cond, int use_cond)
{
    __shared__ double s_rr;
    if (!st->done) {   // uniform
        const double t = warp_reduce_partials(threadIdx.x < 32 ? rr_partial : rz_partial, np);
        if (threadIdx.x == 0) s_rr = t;
        __syncthreads();
        if (threadIdx.x == 32) {
            const double rr = s_rr;
            st->rr = rr;
            st->delta_old = st->delta_new; st->delta_new = t; st->beta = t / st->delta_old;
            st->it += 1;
            if (!(st->it < st->maxit && sqrt(rr) > st->tol)) st->done = 1;
        }
        __syncthreads();
    }
    if (use_cond && threadIdx.x == 0) cudaGraphSetConditional(cond, st->done ? 0u : 1u);
}
// ---- batched PCG: several independent hierarchies ("subs") advance in lock-step ---------------------
// One ddpca_mg may hold nsub >= 1 subdomain hierarchies as ONE block-diagonal hierarchy: every level
// kernel (sweeps, residual, transfers, product) then works on all subdomains at once -- small subdomains,
// which cannot fill 148 SMs alone, share launches instead of queueing behind each other.  The CG recurrence
// stays per subdomain (MGPIS.h:163-225 is called once per subdomain, MCONTACT.h:2531): each sub has its own
// PcgState, reductions are segmented, a converged sub freezes (x, r, p are no longer touched, its iteration
// count stops), the loop ends when every sub is done.
// Rows of one sub are contiguous inside every stage (colour) of the permuted finest level; these ranges are
// cut into CHUNKS of at most kSegRows rows, numbered sub-major, so the partial sums of sub s are the
// contiguous range [sub_chunk[s], sub_chunk[s+1]) -- fixed order, bit-reproducible.
constexpr int kSegRows = 2048;
struct __align__(16) SegChunk { int row0, nrows, sub, pad; };
struct BatchFlags { int done_all; int pad; long long it_max; };

// partial[chunk] = sum_{i in chunk} a_i b_i
__global__ void __launch_bounds__(256) k_seg_dot(const SegChunk *__restrict__ ch, const double *__restrict__ a, const double *__restrict__ b,
                                                 double *partial, const int *done)
{
    if (done && *done) return;
    const SegChunk c = ch[blockIdx.x];
    const double *pa = a + c.row0, *pb = b + c.row0;
    double acc = 0.0;
    for (int i = threadIdx.x; i < c.nrows; i += 256) acc += pa[i] * pb[i];
    block_sum_to_partial(acc, partial);
}
// rel_tol / maxit of the coming solve (kept out of the captured graphs); maxit <= 0: rows of the sub (MGPIS.h:178)
__global__ void k_seg_params(int nsub, PcgState *st, const int *__restrict__ sub_n, double rel_tol, long long maxit)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsub) return;
    st[s].rel_tol = rel_tol;
    st[s].maxit = maxit > 0 ? maxit : (long long)sub_n[s];
}
// after r = b: bb = b.b, tol = rel_tol ||b||, loop condition of MGPIS.h:198 at it = 0 -- one warp per sub
__global__ void __launch_bounds__(1024) k_seg_init(int nsub, PcgState *st, const int *__restrict__ sub_chunk, const double *__restrict__ bb_partial,
                                                    BatchFlags *fl)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int all = 1;
    for (int s = w; s < nsub; s += 32) {
        const double bb = warp_reduce_partials(bb_partial + sub_chunk[s], sub_chunk[s + 1] - sub_chunk[s]);
        int dn = 0;
        if (lane == 0) {
            PcgState *p = st + s;
            p->bb = bb; p->rr = bb; p->tol = p->rel_tol * sqrt(bb);
            p->it = 0; p->alpha = 0.0; p->beta = 0.0; p->delta_old = 0.0; p->delta_new = 0.0;
            dn = !(0 < p->maxit && sqrt(bb) > p->tol);
            p->done = dn;
        }
        dn = __shfl_sync(0xffffffffu, dn, 0);
        all &= dn;
    }
    all = __syncthreads_and(all);
    if (threadIdx.x == 0) { fl->done_all = all; fl->it_max = 0; }
}
// delta_new = r.z after the first preconditioner application (MGPIS.h:197); `cond`: WHILE node of the solve graph
__global__ void __launch_bounds__(1024) k_seg_delta0(int nsub, PcgState *st, const int *__restrict__ sub_chunk, const double *__restrict__ rz_partial,
                                                      const BatchFlags *fl, cudaGraphConditionalHandle cond, int use_cond)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int s = w; s < nsub; s += 32) {
        if (st[s].done) continue;   // warp-uniform
        const double d = warp_reduce_partials(rz_partial + sub_chunk[s], sub_chunk[s + 1] - sub_chunk[s]);
        if (lane == 0) st[s].delta_new = d;
    }
    if (use_cond && threadIdx.x == 0) cudaGraphSetConditional(cond, fl->done_all ? 0u : 1u);
}
// alpha = delta_new / (p.q) ; x += alpha p ; r -= alpha q ; partial = r.r   (MGPIS.h:201-203 and the norm of :198)
// per chunk; every CTA finishes the p.q reduction of ITS sub itself (same fixed order => same alpha in all of them)
__global__ void __launch_bounds__(256) k_seg_update_xr(const SegChunk *__restrict__ ch, const int *__restrict__ sub_chunk, PcgState *st,
                                                       const double *__restrict__ pq_partial, const double *__restrict__ p,
                                                       const double *__restrict__ q, double *__restrict__ x, double *__restrict__ r,
                                                       double *rr_partial, const int *done)
{
    if (*done) return;
    const SegChunk c = ch[blockIdx.x];
    PcgState *ps = st + c.sub;
    if (ps->done) return;   // frozen sub: its partial sums are not consulted any more
    __shared__ double s_alpha;
    if (threadIdx.x < 32) {
        const int c0 = sub_chunk[c.sub];
        const double pq = warp_reduce_partials(pq_partial + c0, sub_chunk[c.sub + 1] - c0);
        if (threadIdx.x == 0) {
            s_alpha = ps->delta_new / pq;
            if ((int)blockIdx.x == c0) { ps->pq = pq; ps->alpha = s_alpha; }
        }
    }
    __syncthreads();
    const double alpha = s_alpha;
    double acc = 0.0;
    for (int i = c.row0 + threadIdx.x; i < c.row0 + c.nrows; i += 256) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * q[i];
        r[i] = ri;
        acc += ri * ri;
    }
    block_sum_to_partial(acc, rr_partial);
}
// per sub: rr = r.r ; delta_old = delta_new ; delta_new = r.z ; beta ; it++ ; loop condition (MGPIS.h:211-213,219,198);
// the batch is done when every sub is
__global__ void __launch_bounds__(1024) k_seg_beta_next(int nsub, PcgState *st, const int *__restrict__ sub_chunk, const double *__restrict__ rr_partial,
                                                         const double *__restrict__ rz_partial, BatchFlags *fl, cudaGraphConditionalHandle cond, int use_cond)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int all = 1;
    if (!fl->done_all) {
        for (int s = w; s < nsub; s += 32) {
            PcgState *p = st + s;
            int dn = p->done;   // warp-uniform
            if (!dn) {
                const int c0 = sub_chunk[s], nc = sub_chunk[s + 1] - c0;
                const double rr = warp_reduce_partials(rr_partial + c0, nc);
                const double rz = warp_reduce_partials(rz_partial + c0, nc);
                if (lane == 0) {
                    p->rr = rr;
                    p->delta_old = p->delta_new; p->delta_new = rz; p->beta = rz / p->delta_old;
                    p->it += 1;
                    if (!(p->it < p->maxit && sqrt(rr) > p->tol)) { p->done = 1; dn = 1; }
                }
                dn = __shfl_sync(0xffffffffu, dn, 0);
            }
            all &= dn;
        }
        all = __syncthreads_and(all);
        if (threadIdx.x == 0) { fl->it_max += 1; if (all) fl->done_all = 1; }
    }
    if (use_cond && threadIdx.x == 0) cudaGraphSetConditional(cond, all ? 0u : 1u);
}
// p = z + beta p   (MGPIS.h:214); skipped for frozen subs (p is dead by then)
__global__ void __launch_bounds__(256) k_seg_update_p(const SegChunk *__restrict__ ch, const PcgState *__restrict__ st, const double *__restrict__ z,
                                                      double *__restrict__ p, const int *done)
{
    if (*done) return;
    const SegChunk c = ch[blockIdx.x];
    if (st[c.sub].done) return;
    const double beta = st[c.sub].beta;
    for (int i = c.row0 + threadIdx.x; i < c.row0 + c.nrows; i += 256) p[i] = z[i] + beta * p[i];
}
// ---- the other MGPIS drivers with their scalars on the device --------------------
// MULT_SOLV (MGPIS.h:130-160), GMRES_SOLV (:227-348) and BiCGSTAB_SOLV (:350-432) of ONE hierarchy
// (nsub == 1): every reduction ends in the KrylovState below, every vector kernel reads its coefficients
// from there, the stopping tests run in one-warp kernels that raise the same done flag the level kernels
// honour (BatchFlags::done_all, and PcgState::done of sub 0 for the chunk producers).  The host only
// enqueues iterations ahead and polls the flag with a lag (mg.cu: krylov_run).
constexpr int kGmStag = 10;   // iterStag, MGPIS.h:255
struct KrylovState {
    double bb, rr, tol;
    long long it, maxit;
    // BiCGSTAB
    double rho_old, rho_new, alph, omeg, beta;
    int fin, pad;
    // GMRES(10): supeHess, Q, R of MGPIS.h:288-316, y of :317-324, moniErro (also MULT_SOLV's 5 entries)
    double normR0;
    double H[kGmStag + 1][kGmStag], Q[kGmStag + 1][kGmStag], R[kGmStag][kGmStag], y[kGmStag], moni[kGmStag];
};
__device__ __forceinline__ void krylov_stop(PcgState *st, BatchFlags *fl) { st->done = 1; fl->done_all = 1; }

// y = b - w (residual from a finished product); grid-stride
__global__ void __launch_bounds__(256) k_kry_sub(int n, const double *__restrict__ b, const double *__restrict__ w, double *__restrict__ y, const int *done)
{
    if (*done) return;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) y[i] = b[i] - w[i];
}
// common start: bb = b.b, tol = rel_tol ||b||, it = 0, loop condition `it < maxit [&& ||r|| > tol]` at it = 0
__global__ void __launch_bounds__(32) k_kry_init(KrylovState *ks, PcgState *st, BatchFlags *fl, const double *__restrict__ bb_partial, int np,
                                                  double rel_tol, long long maxit, int test_norm)
{
    const double bb = warp_reduce_partials(bb_partial, np);
    if (threadIdx.x == 0) {
        ks->bb = bb; ks->rr = bb; ks->tol = rel_tol * sqrt(bb); ks->it = 0; ks->maxit = maxit;
        ks->rho_old = 1.0; ks->rho_new = 1.0; ks->alph = 1.0; ks->omeg = 1.0; ks->beta = 0.0; ks->fin = 0;
        ks->normR0 = 0.0;
        for (int k = 0; k < kGmStag; k++) ks->moni[k] = 0.0;
        const int dn = !(0 < maxit && (!test_norm || sqrt(bb) > ks->tol));
        st->done = dn; st->it = 0; st->rr = bb; st->tol = ks->tol;
        fl->done_all = dn; fl->it_max = 0;
    }
}

// -- BiCGSTAB ------------------------------------------------------------------------------------
// rho = rhat.r, leave when it vanishes (MGPIS.h:383-387); beta of :392-393
__global__ void __launch_bounds__(32) k_bi_rho(KrylovState *ks, PcgState *st, BatchFlags *fl, const double *__restrict__ partial, int np)
{
    if (fl->done_all) return;
    const double rho = warp_reduce_partials(partial, np);
    if (threadIdx.x == 0) {
        ks->rho_new = rho;
        if (fabs(rho) == 0.0) krylov_stop(st, fl);
        else ks->beta = ks->it == 0 ? 0.0 : (rho / ks->rho_old) * (ks->alph / ks->omeg);
    }
}
// p = r in the first iteration, r + beta (p - omeg v) afterwards (MGPIS.h:388-395)
__global__ void __launch_bounds__(256) k_bi_update_p(int n, const KrylovState *__restrict__ ks, const double *__restrict__ r,
                                                     const double *__restrict__ v, double *__restrict__ p, const int *done)
{
    if (*done) return;
    const bool first = ks->it == 0;
    const double beta = ks->beta, omeg = ks->omeg;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) p[i] = first ? r[i] : r[i] + beta * (p[i] - omeg * v[i]);
}
// alph = rho / (rhat.v) (MGPIS.h:404)
__global__ void __launch_bounds__(32) k_bi_alpha(KrylovState *ks, const BatchFlags *fl, const double *__restrict__ partial, int np)
{
    if (fl->done_all) return;
    const double rv = warp_reduce_partials(partial, np);
    if (threadIdx.x == 0) ks->alph = ks->rho_new / rv;
}
// s = r - alph v (MGPIS.h:405)
__global__ void __launch_bounds__(256) k_bi_s(int n, const KrylovState *__restrict__ ks, const double *__restrict__ r,
                                              const double *__restrict__ v, double *__restrict__ s, const int *done)
{
    if (*done) return;
    const double alph = ks->alph;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) s[i] = r[i] - alph * v[i];
}
// ||s|| <= 0: the reference adds alph phat to x and leaves (MGPIS.h:406-409).  Here the iteration is finished with
// omeg = 0 (shat = t = 0 then, so x += alph phat and r = 0 exactly) without counting it, and the loop ends.
__global__ void __launch_bounds__(32) k_bi_scheck(KrylovState *ks, const BatchFlags *fl, const double *__restrict__ ss_partial, int np)
{
    if (fl->done_all) return;
    const double ss = warp_reduce_partials(ss_partial, np);
    if (threadIdx.x == 0) ks->fin = sqrt(ss) <= 0.0;
}
// omeg = t.s / t.t (MGPIS.h:418)
__global__ void __launch_bounds__(64) k_bi_omega(KrylovState *ks, const BatchFlags *fl, const double *__restrict__ ts_partial,
                                                 const double *__restrict__ tt_partial, int np)
{
    if (fl->done_all) return;
    __shared__ double s_ts;
    const double t = warp_reduce_partials(threadIdx.x < 32 ? ts_partial : tt_partial, np);
    if (threadIdx.x == 0) s_ts = t;
    __syncthreads();
    if (threadIdx.x == 32) ks->omeg = ks->fin ? 0.0 : s_ts / t;
}
// x += alph phat + omeg shat ; r = s - omeg t ; partial = r.r (MGPIS.h:419-420 and the norm of :382), per chunk
__global__ void __launch_bounds__(256) k_bi_update_xr(const SegChunk *__restrict__ ch, const KrylovState *__restrict__ ks,
                                                      const double *__restrict__ phat, const double *__restrict__ shat,
                                                      const double *__restrict__ s, const double *__restrict__ t,
                                                      double *__restrict__ x, double *__restrict__ r, double *rr_partial, const int *done)
{
    if (*done) return;
    const SegChunk c = ch[blockIdx.x];
    const double alph = ks->alph, omeg = ks->omeg;
    double acc = 0.0;
    for (int i = c.row0 + threadIdx.x; i < c.row0 + c.nrows; i += 256) {
        x[i] += alph * phat[i] + omeg * shat[i];
        const double ri = s[i] - omeg * t[i];
        r[i] = ri;
        acc += ri * ri;
    }
    block_sum_to_partial(acc, rr_partial);
}
// rr = r.r ; rho_old = rho ; it++ ; loop condition of MGPIS.h:382
__global__ void __launch_bounds__(32) k_bi_next(KrylovState *ks, PcgState *st, BatchFlags *fl, const double *__restrict__ rr_partial, int np)
{
    if (fl->done_all) return;
    const double rr = warp_reduce_partials(rr_partial, np);
    if (threadIdx.x == 0) {
        ks->rr = rr; ks->rho_old = ks->rho_new;
        if (!ks->fin) ks->it += 1;
        fl->it_max = ks->it;
        if (ks->fin || !(ks->it < ks->maxit && sqrt(rr) > ks->tol)) krylov_stop(st, fl);
    }
}

// -- MULT_SOLV -----------------------------------------------------------------------------------
// moniErro[it % 5] = ||b - A x|| ; stagnation test of MGPIS.h:147-153 (VECT_MEDI_OSCI, PREP.h:147-153) ; it++
__global__ void __launch_bounds__(32) k_ms_next(KrylovState *ks, PcgState *st, BatchFlags *fl, const double *__restrict__ rr_partial, int np)
{
    if (fl->done_all) return;
    const double rr = warp_reduce_partials(rr_partial, np);
    if (threadIdx.x == 0) {
        ks->rr = rr;
        ks->moni[ks->it % 5] = sqrt(rr);
        bool stop = false;
        if (ks->it >= 4) {
            double mx = ks->moni[0], mn = ks->moni[0];
            for (int k = 1; k < 5; k++) { mx = fmax(mx, ks->moni[k]); mn = fmin(mn, ks->moni[k]); }
            stop = mx - mn < 0.1 * ((mx + mn) / 2.0);
        }
        if (!stop) { ks->it += 1; stop = !(ks->it < ks->maxit); }
        fl->it_max = ks->it;
        if (stop) krylov_stop(st, fl);
    }
}

// -- GMRES(10) -----------------------------------------------------------------------------------
// restart (MGPIS.h:263-276): normR_0 = ||M^-1 (b - A x_0)||, empty Hessenberg / Q / R
__global__ void __launch_bounds__(32) k_gm_restart(KrylovState *ks, const BatchFlags *fl, const double *__restrict__ ww_partial, int np)
{
    if (fl->done_all) return;
    const double ww = warp_reduce_partials(ww_partial, np);
    double *z = &ks->H[0][0];   // H, Q, R are contiguous members
    const int nz = (kGmStag + 1) * kGmStag * 2 + kGmStag * kGmStag;
    for (int k = threadIdx.x; k < nz; k += 32) z[k] = 0.0;
    if (threadIdx.x == 0) ks->normR0 = sqrt(ww);
}
// dst = src / c with c = normR_0 (which == 0, MGPIS.h:275) or the new column's sub-diagonal entry H[k+1][k] (:294)
__global__ void __launch_bounds__(256) k_gm_scale(int n, const KrylovState *__restrict__ ks, int which, int k, const double *__restrict__ src,
                                                  double *__restrict__ dst, const int *done)
{
    if (*done) return;
    const double c = which == 0 ? ks->normR0 : ks->H[k + 1][k];
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) dst[i] = src[i] / c;
}
// partial[j * np + chunk] = sum_{i in chunk} V_j[i] w[i] for j = 0..k  (b_i = orthBasi^T precV, MGPIS.h:286)
__global__ void __launch_bounds__(256) k_gm_dots(const SegChunk *__restrict__ ch, const double *__restrict__ V, size_t ldv, int k,
                                                 const double *__restrict__ w, double *partial, int np, const int *done)
{
    if (*done) return;
    __shared__ double sm[8];
    const SegChunk c = ch[blockIdx.x];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    for (int j = 0; j <= k; j++) {
        const double *vj = V + (size_t)j * ldv;
        double acc = 0.0;
        for (int i = c.row0 + threadIdx.x; i < c.row0 + c.nrows; i += 256) acc += vj[i] * w[i];
        acc = warp_sum(acc);
        if (lane == 0) sm[wp] = acc;
        __syncthreads();
        if (wp == 0) {
            double t = lane < 8 ? sm[lane] : 0.0;
            t = warp_sum(t);
            if (lane == 0) partial[(size_t)j * np + blockIdx.x] = t;
        }
        __syncthreads();
    }
}
// H[j][k] = b_i(j), j = 0..k
__global__ void __launch_bounds__(32) k_gm_hcol(KrylovState *ks, const BatchFlags *fl, const double *__restrict__ partial, int np, int k)
{
    if (fl->done_all) return;
    for (int j = 0; j <= k; j++) {
        const double t = warp_reduce_partials(partial + (size_t)j * np, np);
        if (threadIdx.x == 0) ks->H[j][k] = t;
    }
}
// w -= sum_j H[j][k] V_j  (q_ip1 = precV - orthBasi b_i, MGPIS.h:287)
__global__ void __launch_bounds__(256) k_gm_orth(int n, const KrylovState *__restrict__ ks, const double *__restrict__ V, size_t ldv, int k,
                                                 double *__restrict__ w, const int *done)
{
    if (*done) return;
    double hk[kGmStag];
#pragma unroll
    for (int j = 0; j < kGmStag; j++) hk[j] = j <= k ? ks->H[j][k] : 0.0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < kGmStag; j++) if (j <= k) s += V[(size_t)j * ldv + i] * hk[j];
        w[i] -= s;
    }
}
// H[k+1][k] = ||q_ip1|| ; one Gram-Schmidt column of the QR of the Hessenberg matrix (MGPIS.h:297-316) ; back
// substitution for y (:317-324)
__global__ void __launch_bounds__(32) k_gm_hess(KrylovState *ks, const BatchFlags *fl, const double *__restrict__ qq_partial, int np, int k)
{
    if (fl->done_all) return;
    const double qq = warp_reduce_partials(qq_partial, np);
    if (threadIdx.x != 0) return;
    ks->H[k + 1][k] = sqrt(qq);
    double col[kGmStag + 1];
    for (int i = 0; i <= k + 1; i++) col[i] = ks->H[i][k];
    for (int j = 0; j < k; j++) {
        double s = 0.0;
        for (int i = 0; i <= k + 1; i++) s += ks->Q[i][j] * ks->H[i][k];
        ks->R[j][k] = s;
    }
    for (int j = 0; j < k; j++) for (int i = 0; i <= k + 1; i++) col[i] -= ks->Q[i][j] * ks->R[j][k];
    double nc = 0.0;
    for (int i = 0; i <= k + 1; i++) nc += col[i] * col[i];
    nc = sqrt(nc);
    ks->R[k][k] = nc;
    for (int i = 0; i <= k + 1; i++) ks->Q[i][k] = col[i] / nc;
    for (int j = k; j >= 0; j--) {
        double s = 0.0;
        for (int c = j + 1; c <= k; c++) s += ks->R[j][c] * ks->y[c];
        ks->y[j] = (ks->normR0 * ks->Q[0][j] - s) / ks->R[j][j];
    }
}
// x = x_0 + sum_{j<=k} y_j V_j (MGPIS.h:325)
__global__ void __launch_bounds__(256) k_gm_update_x(int n, const KrylovState *__restrict__ ks, const double *__restrict__ V, size_t ldv, int k,
                                                     const double *__restrict__ x0, double *__restrict__ x, const int *done)
{
    if (*done) return;
    double yk[kGmStag];
#pragma unroll
    for (int j = 0; j < kGmStag; j++) yk[j] = j <= k ? ks->y[j] : 0.0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < kGmStag; j++) if (j <= k) s += V[(size_t)j * ldv + i] * yk[j];
        x[i] = x0[i] + s;
    }
}
// moniErro[k] = ||b - A x|| ; stopping tests of MGPIS.h:333-341 ; it++ ; `it < maxiNumb` of :261
__global__ void __launch_bounds__(32) k_gm_next(KrylovState *ks, PcgState *st, BatchFlags *fl, const double *__restrict__ rr_partial, int np, int k)
{
    if (fl->done_all) return;
    const double rr = warp_reduce_partials(rr_partial, np);
    if (threadIdx.x == 0) {
        ks->rr = rr;
        const double e = sqrt(rr);
        ks->moni[k] = e;
        bool stop = false;
        if (ks->it >= kGmStag - 1) {
            double mx = ks->moni[0], mn = ks->moni[0];
            for (int c = 1; c < kGmStag; c++) { mx = fmax(mx, ks->moni[c]); mn = fmin(mn, ks->moni[c]); }
            stop = e <= ks->tol || (e <= 1.0E2 * ks->tol && (mx - mn) < 0.1 * ((mx + mn) / 2.0));
        }
        if (!stop) { ks->it += 1; stop = !(ks->it < ks->maxit); }
        fl->it_max = ks->it;
        if (stop) krylov_stop(st, fl);
    }
}

// Block-diagonal dense solves: y_s = Binv_s b_s for every block s (level-0 direct solves of a batch, interface
// mass matrices of the ADMM loop).  One warp per row; off[s] = first row of block s, bptr[s] = its dense
// inverse (n_s x n_s, row-major) or null: that block is handled elsewhere and its rows are left alone.
__global__ void __launch_bounds__(256) k_dense_gemv_batch(int nsub, const int *__restrict__ off, const double *const *__restrict__ bptr,
                                                          const double *__restrict__ x, double *__restrict__ y, const int *done)
{
    if (done && *done) return;
    const int lane = threadIdx.x & 31;
    const int i = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (i >= off[nsub]) return;
    int lo = 0, hi = nsub;   // largest s with off[s] <= i
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= i) lo = mid; else hi = mid; }
    const int n = off[lo + 1] - off[lo];
    const double *Bs = bptr[lo];
    if (Bs == nullptr) return;
    const double *row = Bs + (size_t)(i - off[lo]) * n;
    const double *xs = x + off[lo];
    double s = 0.0;
    if ((n & 1) == 0 && (reinterpret_cast<uintptr_t>(xs) & 15) == 0 && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
        const int n2 = n >> 1;
        const double2 *x2 = reinterpret_cast<const double2 *>(xs);
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int j = lane;
        for (; j + 96 < n2; j += 128) {
            const double2 a0 = ld_stream2(row + 2 * j), a1 = ld_stream2(row + 2 * (j + 32));
            const double2 a2 = ld_stream2(row + 2 * (j + 64)), a3 = ld_stream2(row + 2 * (j + 96));
            const double2 b0 = __ldg(x2 + j), b1 = __ldg(x2 + j + 32), b2 = __ldg(x2 + j + 64), b3 = __ldg(x2 + j + 96);
            s0 += a0.x * b0.x + a0.y * b0.y;
            s1 += a1.x * b1.x + a1.y * b1.y;
            s2 += a2.x * b2.x + a2.y * b2.y;
            s3 += a3.x * b3.x + a3.y * b3.y;
        }
        for (; j < n2; j += 32) {
            const double2 a0 = ld_stream2(row + 2 * j);
            const double2 b0 = __ldg(x2 + j);
            s0 += a0.x * b0.x + a0.y * b0.y;
        }
        s = (s0 + s1) + (s2 + s3);
    } else {
        for (int j = lane; j < n; j += 32) s += ld_stream(row + j) * __ldg(xs + j);
    }
    s = warp_sum(s);
    if (lane == 0) y[i] = s;
}

// ---- ADMM interface kernels (MCONTACT.h:2632-2668, :2737-2833) ---------------------------------
// gamma = 0.5*(t - gapTerm) followed by the contact projection; t already holds
// inpoLagr0 l0 - inpoLagr1 l1 + pemaInpo_r0 u0 - pemaInpo_r1 u1.
//   fric <  0 : tied interface, no projection                               (:2637)
//   fric == 0 : one component per integration point, gamma = max(0, gamma)  (:2639-2641)
//   fric >  0 : (n, t1, t2) per point; normal clamp, Coulomb cone on the tangential pair,
//               status 0 open / 1 slide / 2 stick in stat[3*ip+1]            (:2643-2667)
__global__ void k_gamma_project(int nip, int d, double fric, const double *__restrict__ t, const double *__restrict__ gap,
                                double *__restrict__ gamma, int *__restrict__ stat)
{
    const int ip = blockIdx.x * blockDim.x + threadIdx.x;
    if (ip >= nip) return;
    if (d == 1) {
        double g = 0.5 * (t[ip] - gap[ip]);
        if (fric >= 0.0) g = fmax(0.0, g);
        gamma[ip] = g;
        stat[ip] = 0;
        return;
    }
    double gn = 0.5 * (t[3 * ip] - gap[3 * ip]);
    double g1 = 0.5 * (t[3 * ip + 1] - gap[3 * ip + 1]);
    double g2 = 0.5 * (t[3 * ip + 2] - gap[3 * ip + 2]);
    int st = 0;
    if (fric >= 0.0) gn = fmax(0.0, gn);
    if (fric > 0.0) {
        if (gn > 0.0) {
            const double slid = fric * gn;
            const double nrm = sqrt(g1 * g1 + g2 * g2);
            if (nrm >= slid) { const double f = slid / nrm; g1 = f * g1; g2 = f * g2; st = 1; }
            else st = 2;
        } else { g1 = 0.0; g2 = 0.0; st = 0; }
    }
    gamma[3 * ip] = gn; gamma[3 * ip + 1] = g1; gamma[3 * ip + 2] = g2;
    stat[3 * ip] = 0; stat[3 * ip + 1] = st; stat[3 * ip + 2] = 0;
}
// y = a*x + b*y
__global__ void k_axpby(int n, double a, const double *__restrict__ x, double b, double *__restrict__ y)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = (b == 0.0) ? a * x[i] : a * x[i] + b * y[i];   // b == 0: y may be uninitialised
}
// y += a*x
__global__ void k_axpy(int n, double a, const double *__restrict__ x, double *__restrict__ y)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += a * x[i];
}
constexpr int kMoniBlocks = 64;
// part[0..64) = partial sums of (cur-prev)^2, part[64..128) of cur^2   (MONITOR, :2738-2739)
__global__ void __launch_bounds__(256) k_moni_partial(int n, const double *__restrict__ cur, const double *__restrict__ prev, double *part)
{
    double a = 0.0, b = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double c = cur[i], dlt = c - prev[i];
        a += dlt * dlt;
        b += c * c;
    }
    __shared__ double sa[8], sb[8];
    a = warp_sum(a); b = warp_sum(b);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sa[w] = a; sb[w] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) { ta += sa[k]; tb += sb[k]; }
        part[blockIdx.x] = ta;
        part[kMoniBlocks + blockIdx.x] = tb;
    }
}
// out[2s], out[2s+1] = fixed-order sums of slot s
__global__ void k_moni_final(const double *__restrict__ part, double *__restrict__ out)
{
    const int s = blockIdx.x;
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
        for (int k = 0; k < kMoniBlocks; k++) { ta += part[s * 2 * kMoniBlocks + k]; tb += part[s * 2 * kMoniBlocks + kMoniBlocks + k]; }
        out[2 * s] = ta;
        out[2 * s + 1] = tb;
    }
}

}  // namespace ddpca
