// kernels2.cuh -- "v2" level kernels: split lower/upper storage streamed through shared memory
// by the TMA engine (cp.async.bulk + mbarrier), one cooperative launch per Gauss-Seidel sweep.
//
// Why (profiles/r1c): the v1 sweep kernels are latency-bound -- three dependent memory round
// trips per row group (descriptor -> pattern/values -> x gather), short stage launches, and
// 64-byte DRAM granularity over-fetch when only half of each row is needed.  Here
//   * the strictly-lower and strictly-upper couplings of a level live in two separate arrays
//     in stage order, so a half sweep streams ONE contiguous region;
//   * a CTA processes CHUNKS of up to 16 row groups (passes over one half cut theirs by bytes):
//     descriptor, group metadata, diagonal blocks, patterns and values of a chunk arrive in shared
//     memory by a handful of bulk copies
//     (UBLKCP), double-buffered, so the only latency left on the critical path is the x gather,
//     and all gathers of a group are issued back to back from a pattern that is already on chip;
//   * the stages (colours) of one sweep run inside one cooperative kernel with a grid barrier of the
//     consumer warps between them; the next stage's first chunks are already in flight during the
//     barrier.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace ddpca {

namespace cg = cooperative_groups;

#ifndef DDPCA_V2_GROUPS
#define DDPCA_V2_GROUPS 16
#endif
#ifndef DDPCA_V2_BUFS
#define DDPCA_V2_BUFS 2
#endif
#ifndef DDPCA_V2_LANES
#define DDPCA_V2_LANES 8
#endif
constexpr int GL2 = DDPCA_V2_LANES;                  // lanes per row group in the v2 kernels (8 or 16)
#ifndef DDPCA_V2_HALF_GROUPS
#define DDPCA_V2_HALF_GROUPS DDPCA_V2_GROUPS
#endif
constexpr int kChunkGroups = DDPCA_V2_GROUPS;        // row groups per chunk of a pass over whole rows (FWD_FULL, SPMV)
// Passes over one half of the rows (FWD_ZERO, BWD, RESID) have their own chunk tables, cut by BYTES:
// at most kHalfGroups groups and at most 1/kHalfBudgetDiv of the streamed bytes of the largest
// whole-row chunk.  With the default (16 groups, 1/2) a half-pass ring is half as large, 5 CTAs fit
// on a SM instead of 3, and colours whose couplings all lie in one half get chunks of ~8 groups.
constexpr int kHalfGroups = DDPCA_V2_HALF_GROUPS;
#ifndef DDPCA_V2_HALF_BUDGET_DIV
#define DDPCA_V2_HALF_BUDGET_DIV 2
#endif
constexpr int kHalfBudgetDiv = DDPCA_V2_HALF_BUDGET_DIV;   // half-pass chunk budget = largest whole-row chunk / this
constexpr int kV2Bufs = DDPCA_V2_BUFS;               // depth of the shared-memory ring
constexpr int kV2StagesSmem = 64;                     // stage tables up to this size are cached in shared memory
static_assert(GL2 == 8 || GL2 == 16 || GL2 == 32, "v2 kernels: 8, 16 or 32 lanes per row group");
static_assert((kChunkGroups * GL2) % 32 == 0 && (kHalfGroups * GL2) % 32 == 0, "consumer threads must fill whole warps");
static_assert(kHalfGroups >= kChunkGroups, "half-pass chunks hold at least as many groups as whole-row chunks");
__device__ __forceinline__ unsigned subwarp_mask2()
{
    const unsigned lane = threadIdx.x & 31u;
    return (GL2 == 32) ? 0xffffffffu : (((1u << GL2) - 1u) << (lane & ~(unsigned)(GL2 - 1)));
}
__device__ __forceinline__ double group_sum2(double v, unsigned mask)
{
#pragma unroll
    for (int o = GL2 / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}

// 32-byte group descriptor of the split layout
struct __align__(16) GroupMeta2 {
    int row0, gs;
    int nl, nu;      // lower / upper pattern lengths, padded to multiples of 4 (pads: value 0)
    int cl, cu;      // offsets into CL / CU (ints)
    int vl, vu;      // offsets into VL / VU in units of 16 bytes (2 doubles)
};
// 48-byte chunk descriptor
struct __align__(16) ChunkDesc {
    int g0, ng;
    int cl0, ncl;    // pattern range (ints) of the chunk's lower parts
    int cu0, ncu;
    int vl0, nvl;    // value range in 16-byte units
    int vu0, nvu;
    int sub, pad1;   // sub: subdomain of a batched hierarchy the chunk belongs to (chunks never straddle subdomains)
};
constexpr int kBlkStride = 12;     // doubles per in-group block record: the gs x gs block row-major (9) + 1/diagonal (3)
// The v2 sweeps multiply by the stored reciprocal of the diagonal instead of dividing (three dependent
// FP64 divisions per row group sit on the critical path of every chunk otherwise) and take
// p1 = b - L x_new directly instead of D x_new + U x_old (equal in exact arithmetic, MGPIS.h:70-72).
// Results move by <= 1 ulp per operation; DDPCA_V2_RECIP_DIAG=0 restores the divisions.
#ifndef DDPCA_V2_RECIP_DIAG
#define DDPCA_V2_RECIP_DIAG 1
#endif

struct Lvl2View {
    int n, ng, nchunks, nstages;
    int fuse_bwd_last;   // 1: no group of the last stage has couplings above it -- the forward sweeps finish that
                         // stage with its backward substitution and the backward sweep starts one stage earlier
    const GroupMeta2 *__restrict__ meta;
    const int *__restrict__ CL;
    const double *__restrict__ VL;
    const int *__restrict__ CU;
    const double *__restrict__ VU;
    const double *__restrict__ BD;          // [ng * kBlkStride]
    const ChunkDesc *__restrict__ chunks;   // [nchunks], stage after stage -- the table of the pass type (V2_TAB_*)
    const int *__restrict__ stage_chunk;    // [nstages+1]
    // batched PCG: subdomains whose CG has converged are frozen -- the producer replaces their chunks by an empty
    // descriptor (ng = 0: 48 bytes instead of ~20 KB, nothing to compute), the ring protocol stays as it is
    const ChunkDesc *__restrict__ empty_desc;   // one descriptor with ng = 0
    const PcgState *__restrict__ sub_state;     // [nsub] or null: nothing is skipped
};

enum { V2_FWD_ZERO = 0, V2_FWD_FULL = 1, V2_BWD = 2, V2_RESID = 3, V2_SPMV = 4 };
// chunk tables of a level: whole rows / lower halves / upper halves
enum { V2_TAB_FULL = 0, V2_TAB_LO = 1, V2_TAB_UP = 2 };
__host__ __device__ constexpr int v2_table(int mode) { return (mode == V2_FWD_FULL || mode == V2_SPMV) ? V2_TAB_FULL : (mode == V2_BWD ? V2_TAB_UP : V2_TAB_LO); }
__host__ __device__ constexpr int v2_groups(int mode) { return v2_table(mode) == V2_TAB_FULL ? kChunkGroups : kHalfGroups; }   // consumer sub-warps per CTA
__host__ __device__ constexpr int v2_consumers(int mode) { return v2_groups(mode) * GL2; }
__host__ __device__ constexpr int v2_threads(int mode) { return v2_consumers(mode) + 32; }   // + one producer warp (one lane issues the bulk copies)

// fixed offsets inside a shared-memory buffer (G = group capacity of the chunk table)
constexpr int kOffDesc = 0;                                   // 48 B (+16 pad)
constexpr int kOffMeta = 64;                                  // G * 32 B
__host__ __device__ constexpr int v2_off_blk(int G) { return kOffMeta + G * 32; }              // G * kBlkStride * 8 B
__host__ __device__ constexpr int v2_off_data(int G) { return v2_off_blk(G) + G * kBlkStride * 8; }

__host__ __device__ inline size_t v2_chunk_bytes(int mode, int ncl, int nvl, int ncu, int nvu)
{
    size_t b = v2_off_data(v2_groups(mode));
    const bool lo = (mode != V2_BWD), up = (mode == V2_FWD_FULL || mode == V2_BWD || mode == V2_SPMV);
    if (lo) b += (size_t)ncl * 4 + (size_t)nvl * 16;
    if (up) b += (size_t)ncu * 4 + (size_t)nvu * 16;
    return b;
}

// ---- mbarrier / bulk-copy PTX ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// barrier among the consumer threads only (the producer warp never joins)
template <int NCONS>
__device__ __forceinline__ void consumer_bar_sync()
{
    asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory");
}
// Grid-wide barrier among the consumer threads of all CTAs of a cooperative launch (all CTAs are
// co-resident).  Flat and monotonic: a CTA arrives with ONE fire-and-forget red.release on one of
// kGbarFan counters (separate 128-byte lines, so arrivals spread over L2 atomic units), and the first
// lanes of its warp 0 poll all counters with ld.acquire until each holds epoch * (its share of the
// grid) -- one L2 one-way trip for the arrival plus the poll instead of a chain of dependent atomic
// round trips.  The counters are never reset inside the sweep; the last CTA to leave the kernel zeroes
// them (consumer_grid_barrier_exit).  Layout: gbar[32 * k] counter k, gbar[32 * kGbarFan] exit count.
#ifndef DDPCA_GBAR_FAN
#define DDPCA_GBAR_FAN 8
#endif
#ifndef DDPCA_GBAR_POLL
#define DDPCA_GBAR_POLL 0
#endif
constexpr int kGbarFan = DDPCA_GBAR_FAN;
static_assert(kGbarFan >= 1 && kGbarFan <= 32, "one polling lane per counter");
constexpr int kGbarWords = 32 * (1 + kGbarFan);
__device__ __forceinline__ unsigned atom_add_acqrel(unsigned *p, unsigned v)
{
    unsigned old;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void red_add_release(unsigned *p, unsigned v)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned *p, unsigned v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// epoch = 1 for the first barrier of the launch, 2 for the second, ...  A CTA's stores are ordered
// before thread 0's release by the CTA barrier; the polling lanes' acquires are ordered before the
// other consumers' loads by the second CTA barrier.
template <int NCONS>
__device__ __forceinline__ void consumer_grid_barrier(unsigned *gbar, unsigned epoch)
{
    consumer_bar_sync<NCONS>();
    if (threadIdx.x < 32) {
        const unsigned nsub = gridDim.x < (unsigned)kGbarFan ? gridDim.x : (unsigned)kGbarFan;
        if (threadIdx.x == 0) red_add_release(gbar + 32 * (blockIdx.x % nsub), 1u);
        if (threadIdx.x < nsub) {
            const unsigned sub_size = (gridDim.x - threadIdx.x + nsub - 1) / nsub;
            const unsigned target = epoch * sub_size;
            const unsigned *cnt = gbar + 32 * threadIdx.x;
#if DDPCA_GBAR_POLL == 0
            while (ld_acquire(cnt) < target) { }
#else
            // relaxed polling + one acquire fence: an acquire load drops the SM's L1 lines on every
            // poll, which the CTAs still working on the stage would pay for with gather misses
            while (ld_relaxed(cnt) < target) { }
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
        }
        __syncwarp();
    }
    consumer_bar_sync<NCONS>();
}
// after the last barrier of the launch: the last CTA out zeroes the counters for the next launch
__device__ __forceinline__ void consumer_grid_barrier_exit(unsigned *gbar)
{
    if (threadIdx.x == 0) {
        unsigned *ex = gbar + 32 * kGbarFan;
        if (atom_add_acqrel(ex, 1u) == gridDim.x - 1) {
            const unsigned nsub = gridDim.x < (unsigned)kGbarFan ? gridDim.x : (unsigned)kGbarFan;
            for (unsigned k = 0; k < nsub; k++) st_relaxed(gbar + 32 * k, 0u);
            st_relaxed(ex, 0u);
        }
    }
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// issue all bulk copies of one chunk (one thread)
template <int MODE>
__device__ __forceinline__ void v2_issue_chunk(const Lvl2View &A, int c, const ChunkDesc &d, unsigned char *buf, uint64_t *bar, const PcgState *sub_state)
{
    if (sub_state != nullptr && sub_state[d.sub].done) {   // frozen subdomain: an empty chunk keeps the ring in step
        mbar_expect_tx(bar, 48);
        bulk_g2s(buf + kOffDesc, A.empty_desc, 48, bar);
        return;
    }
    constexpr bool LO = (MODE != V2_BWD);
    constexpr bool UP = (MODE == V2_FWD_FULL || MODE == V2_BWD || MODE == V2_SPMV);
    constexpr int kOffBlk = v2_off_blk(v2_groups(MODE)), kOffData = v2_off_data(v2_groups(MODE));
    uint32_t bytes = 48 + (uint32_t)d.ng * 32 + (uint32_t)d.ng * kBlkStride * 8;
    if (LO) bytes += (uint32_t)d.ncl * 4 + (uint32_t)d.nvl * 16;
    if (UP) bytes += (uint32_t)d.ncu * 4 + (uint32_t)d.nvu * 16;
    mbar_expect_tx(bar, bytes);
    bulk_g2s(buf + kOffDesc, A.chunks + c, 48, bar);
    bulk_g2s(buf + kOffMeta, A.meta + d.g0, (uint32_t)d.ng * 32, bar);
    bulk_g2s(buf + kOffBlk, A.BD + (size_t)d.g0 * kBlkStride, (uint32_t)d.ng * kBlkStride * 8, bar);
    unsigned char *p = buf + kOffData;
    if (LO) {
        if (d.ncl) {
            bulk_g2s(p, A.CL + d.cl0, (uint32_t)d.ncl * 4, bar);
            bulk_g2s(p + (size_t)d.ncl * 4, A.VL + (size_t)d.vl0 * 2, (uint32_t)d.nvl * 16, bar);
        }
        p += (size_t)d.ncl * 4 + (size_t)d.nvl * 16;
    }
    if (UP) {
        if (d.ncu) {
            bulk_g2s(p, A.CU + d.cu0, (uint32_t)d.ncu * 4, bar);
            bulk_g2s(p + (size_t)d.ncu * 4, A.VU + (size_t)d.vu0 * 2, (uint32_t)d.nvu * 16, bar);
        }
    }
}

// Two-phase evaluation of one half of a row group from shared memory:
//   phase 1 (v2_gather): every x gather of the half is issued back to back into registers
//           (v2_iters(MODE) steps of 2*GL pattern positions cover 64 or 80 entries; hexahedral meshes have
//           at most 40 lower / 40 upper couplings per node) -- ONE exposed L2 latency per group;
//   phase 2 (v2_fma):    values stream from shared memory into the row sums.
// Longer halves continue in a generic loop (v2_tail).
// Entries of one half covered by the unrolled steps.  A node of a hexahedral mesh couples to 26 others:
// at most 78 entries in one half.  The sweeps and the residual cover 80 (measured +2 % on the half
// passes: no second dependent gather for the colours whose couplings all lie on one side); the plain
// product keeps 64, where the extra registers cost more than the rare tail (measured 6.0 vs 5.3 TB/s).
__host__ __device__ constexpr int v2_iters(int mode) { return ((mode == V2_SPMV ? 64 : 80) + 2 * GL2 - 1) / (2 * GL2); }

// x gathers go through L1: neighbouring row groups share most of their columns, and an L1 hit
// takes the gather off the L2 round trip that bounds these kernels.
//   RESID, SPMV: x is constant for the whole launch -> read-only path (ld.global.nc);
//   staged sweeps: other CTAs write x between stages -> ordinary (weak, L1-allocating) loads.  They
//   are ordered after those writes by the grid barrier (release on arrival, acquire by the polling
//   lanes -- which drops the SM's L1 lines --, CTA barrier), and nothing a stage reads is written
//   during that stage (rows of one colour are contiguous after the permutation and mutually
//   uncoupled; a sector shared with the colour being written is only read for its stable part and
//   is dropped again at the next barrier).
// DDPCA_V2_GATHER_L1: 0 = L2 only everywhere, 1 = L1 for RESID/SPMV only, 2 = everywhere.
#ifndef DDPCA_V2_GATHER_L1
#define DDPCA_V2_GATHER_L1 2
#endif
template <bool RO>
__device__ __forceinline__ double v2_ldx(const double *p)
{
    if (RO && DDPCA_V2_GATHER_L1 >= 1) return __ldg(p);
    if (!RO && DDPCA_V2_GATHER_L1 >= 2) return *p;
    return __ldcg(p);
}
template <bool RO, int ITERS>
__device__ __forceinline__ void v2_gather(const int *__restrict__ pc, int n, const double *x, int sl, double (&xs)[2 * ITERS])
{
#pragma unroll
    for (int it = 0; it < ITERS; it++) {
        const int k = 2 * sl + it * 2 * GL2;
        xs[2 * it] = 0.0;
        xs[2 * it + 1] = 0.0;
        if (k < n) {
            const int2 c = *reinterpret_cast<const int2 *>(pc + k);
            xs[2 * it] = v2_ldx<RO>(x + c.x);
            xs[2 * it + 1] = v2_ldx<RO>(x + c.y);
        }
    }
}
template <int ITERS>
__device__ __forceinline__ void v2_fma(const double *__restrict__ pv, int n, int gs, int sl, const double (&xs)[2 * ITERS], double (&s)[3])
{
#pragma unroll
    for (int it = 0; it < ITERS; it++) {
        const int k = 2 * sl + it * 2 * GL2;
        if (k < n) {
#pragma unroll
            for (int r = 0; r < 3; r++)
                if (r < gs) {
                    const double2 a = *reinterpret_cast<const double2 *>(pv + (size_t)r * n + k);
                    s[r] += a.x * xs[2 * it] + a.y * xs[2 * it + 1];
                }
        }
    }
}
template <bool RO, int ITERS>
__device__ __forceinline__ void v2_tail(const int *__restrict__ pc, const double *__restrict__ pv, int n, int gs,
                                        const double *x, int sl, double (&s)[3])
{
    for (int k = 2 * sl + ITERS * 2 * GL2; k < n; k += 2 * GL2) {
        const int2 c = *reinterpret_cast<const int2 *>(pc + k);
        const double x0 = v2_ldx<RO>(x + c.x), x1 = v2_ldx<RO>(x + c.y);
#pragma unroll
        for (int r = 0; r < 3; r++)
            if (r < gs) {
                const double2 a = *reinterpret_cast<const double2 *>(pv + (size_t)r * n + k);
                s[r] += a.x * x0 + a.y * x1;
            }
    }
}

// One sweep (all stages, descending for V2_BWD) or one stage-less pass (RESID, SPMV) over a level.
//   FWD_*: b = right-hand side, x in/out, p1 out          (MGPIS.h:66-72)
//   BWD  : p1 in, x in/out                                 (MGPIS.h:73-76)
//   RESID: y = b - (p1 + L x)                              (MGPIS.h:92)
//   SPMV : y = A x ; if w: partial[blockIdx] = sum w_i y_i (MGPIS.h:200-201)
// Warp-specialised: one producer lane walks the CTA's chunk sequence kV2Bufs chunks ahead of the
// consumers (full / empty mbarriers per ring slot) and keeps crossing stage boundaries -- operator
// data does not depend on x --, the consumer warps meet the other CTAs at a grid barrier between
// stages.  Staged modes must be launched cooperatively (co-residency of all CTAs).
template <int MODE>
__global__ void __launch_bounds__(v2_threads(MODE)) k_level_pass(Lvl2View A, size_t buf_bytes, unsigned *gbar, const double *__restrict__ b,
                                                            double *x, double *p1, double *y, const double *__restrict__ w,
                                                            double *partial, const int *done)
{
    if (done && *done) return;   // uniform across the grid: no CTA reaches a barrier
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[kV2Bufs], empty[kV2Bufs];
    __shared__ int s_stage_chunk[kV2StagesSmem + 1];
    constexpr bool LO = (MODE != V2_BWD);
    constexpr bool UP = (MODE == V2_FWD_FULL || MODE == V2_BWD || MODE == V2_SPMV);
    constexpr bool STAGED = (MODE == V2_FWD_ZERO || MODE == V2_FWD_FULL || MODE == V2_BWD);
    constexpr int kV2Consumers = v2_consumers(MODE);
    constexpr int kOffBlk = v2_off_blk(v2_groups(MODE)), kOffData = v2_off_data(v2_groups(MODE));
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < kV2Bufs; i++) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kV2Consumers / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the stage table is consulted for every chunk by producer and consumers: keep it on chip
    const bool sc_smem = A.nstages <= kV2StagesSmem;
    if (sc_smem)
        for (int i = tid; i <= A.nstages; i += blockDim.x) s_stage_chunk[i] = A.stage_chunk[i];
    __syncthreads();
    const int *sc = sc_smem ? s_stage_chunk : A.stage_chunk;

    // this CTA's chunk sequence: stage after stage, chunks c = first(stage) + blockIdx.x + j * gridDim.x
    const int nst = STAGED ? (MODE == V2_BWD ? A.nstages - A.fuse_bwd_last : A.nstages) : 1;
    auto stage_of = [&](int si) { return STAGED ? (MODE == V2_BWD ? (nst - 1 - si) : si) : 0; };
    auto c_begin = [&](int si) { return STAGED ? sc[stage_of(si)] : 0; };
    auto c_end = [&](int si) { return STAGED ? sc[stage_of(si) + 1] : A.nchunks; };
    auto settle = [&](int &si, int &c) {   // advance (si, c) to this CTA's next chunk at or after c
        while (si < nst) {
            if (c < c_end(si)) return true;
            si++;
            if (si < nst) c = c_begin(si) + (int)blockIdx.x;
        }
        return false;
    };
    int si = 0, c = c_begin(0) + (int)blockIdx.x;
    double acc = 0.0;

    if (tid >= kV2Consumers) {
        // ------------------------------- producer warp ---------------------------------------
        if (tid == kV2Consumers) {
            // inside a PCG loop (done != null) the states of the batch's subdomains are current: skip the frozen ones;
            // stand-alone passes (ddpca_mg_vcycle, the host-driven solvers) never skip
            const PcgState *sub_state = done ? A.sub_state : nullptr;
            // the descriptor of chunk j+1 is fetched before the wait for chunk j's ring slot, so
            // that its global-memory latency never sits between a slot release and the next copy
            bool have = settle(si, c);
            ChunkDesc d_next;
            if (have) d_next = A.chunks[c];
            for (int j = 0; have; j++) {
                const ChunkDesc d = d_next;
                const int c_this = c;
                c += (int)gridDim.x;
                have = settle(si, c);
                if (have) d_next = A.chunks[c];
                const int slot = j % kV2Bufs;
                if (j >= kV2Bufs) mbar_wait(&empty[slot], (uint32_t)(((j / kV2Bufs) - 1) & 1));
                fence_proxy_async();
                v2_issue_chunk<MODE>(A, c_this, d, smem + (size_t)slot * buf_bytes, &full[slot], sub_state);
            }
        }
    } else {
        // ------------------------------- consumer warps --------------------------------------
        const int sl = tid % GL2, sw = tid / GL2;
        const unsigned mask = subwarp_mask2();
        int si_done = 0;   // stage boundaries passed so far
        for (int j = 0;; j++, c += (int)gridDim.x) {
            const bool have = settle(si, c);
            if (STAGED) {
                const int target = have ? si : nst - 1;
                while (si_done < target) { si_done++; consumer_grid_barrier<kV2Consumers>(gbar, (unsigned)si_done); }
            }
            if (!have) break;
            const int slot = j % kV2Bufs;
            mbar_wait(&full[slot], (uint32_t)((j / kV2Bufs) & 1));
            const unsigned char *buf = smem + (size_t)slot * buf_bytes;
            const ChunkDesc &d = *reinterpret_cast<const ChunkDesc *>(buf + kOffDesc);
            if (sw < d.ng) {
                const GroupMeta2 m = *reinterpret_cast<const GroupMeta2 *>(buf + kOffMeta + sw * 32);
                const double *blk = reinterpret_cast<const double *>(buf + kOffBlk) + sw * kBlkStride;
                const unsigned char *pdat = buf + kOffData;
                const int *cL = nullptr, *cU = nullptr;
                const double *vL = nullptr, *vU = nullptr;
                if (LO) {
                    cL = reinterpret_cast<const int *>(pdat) + (m.cl - d.cl0);
                    vL = reinterpret_cast<const double *>(pdat + (size_t)d.ncl * 4) + (size_t)(m.vl - d.vl0) * 2;
                    pdat += (size_t)d.ncl * 4 + (size_t)d.nvl * 16;
                }
                if (UP) {
                    cU = reinterpret_cast<const int *>(pdat) + (m.cu - d.cu0);
                    vU = reinterpret_cast<const double *>(pdat + (size_t)d.ncu * 4) + (size_t)(m.vu - d.vu0) * 2;
                }
                const int gs = m.gs, r0 = m.row0;
                double sL[3] = {0.0, 0.0, 0.0}, sU[3] = {0.0, 0.0, 0.0};
                // right-hand sides and old iterate of the own rows: issued before the gathers they overlap with
                double rhs[3] = {0.0, 0.0, 0.0}, xo[3] = {0.0, 0.0, 0.0};
                if (MODE == V2_FWD_ZERO || MODE == V2_FWD_FULL || MODE == V2_BWD) {
#pragma unroll
                    for (int r = 0; r < 3; r++)
                        if (r < gs) {
                            rhs[r] = (MODE == V2_BWD) ? p1[r0 + r] : b[r0 + r];
                            if (MODE == V2_FWD_FULL) xo[r] = __ldcg(x + r0 + r);
                        }
                } else {
                    // RESID / SPMV: the operands of the finishing lanes, too
#pragma unroll
                    for (int r = 0; r < 3; r++)
                        if (r < gs) xo[r] = v2_ldx<true>(x + r0 + r);
                    if (sl < gs) {
                        if (MODE == V2_RESID) { rhs[0] = b[r0 + sl]; rhs[1] = p1[r0 + sl]; }
                        else if (w) rhs[0] = w[r0 + sl];
                    }
                }
                {
                    constexpr int IT = v2_iters(MODE);
                    double xl[2 * IT], xu[2 * IT];
                    if (LO) v2_gather<!STAGED, IT>(cL, m.nl, x, sl, xl);
                    if (UP) v2_gather<!STAGED, IT>(cU, m.nu, x, sl, xu);
                    if (LO) v2_fma<IT>(vL, m.nl, gs, sl, xl, sL);
                    if (UP) v2_fma<IT>(vU, m.nu, gs, sl, xu, sU);
                    if (LO && m.nl > IT * 2 * GL2) v2_tail<!STAGED, IT>(cL, vL, m.nl, gs, x, sl, sL);
                    if (UP && m.nu > IT * 2 * GL2) v2_tail<!STAGED, IT>(cU, vU, m.nu, gs, x, sl, sU);
                }
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    if (LO) sL[r] = group_sum2(sL[r], mask);
                    if (UP) sU[r] = group_sum2(sU[r], mask);
                }
                if (MODE == V2_FWD_ZERO || MODE == V2_FWD_FULL) {
                    double xn[3] = {0.0, 0.0, 0.0};
#pragma unroll
                    for (int r = 0; r < 3; r++) {
                        if (r < gs) {
                            double inL = 0.0, inU = 0.0;
#pragma unroll
                            for (int cc = 0; cc < 3; cc++) {
                                if (cc < r) inL += blk[r * 3 + cc] * xn[cc];
                                if (cc > r) inU += blk[r * 3 + cc] * xo[cc];
                            }
                            const double up = sU[r] + inU;
#if DDPCA_V2_RECIP_DIAG
                            const double t = rhs[r] - sL[r] - inL;
                            xn[r] = (t - up) * blk[9 + r];
                            rhs[r] = t;                 // p1 (MGPIS.h:72)
#else
                            const double dg = blk[r * 3 + r];
                            xn[r] = (rhs[r] - sL[r] - inL - up) / dg;
                            rhs[r] = dg * xn[r] + up;   // p1 (MGPIS.h:72)
#endif
                            if (sl == 0) p1[r0 + r] = rhs[r];
                        }
                    }
                    if (A.fuse_bwd_last && si == nst - 1) {
                        // last colour: the backward sweep's step for these rows (MGPIS.h:73-76) only involves
                        // the group itself -- do it now, the backward launch starts at the previous colour
                        double xb[3] = {0.0, 0.0, 0.0};
#pragma unroll
                        for (int r = 2; r >= 0; r--) {
                            if (r < gs) {
                                double inU = 0.0;
#pragma unroll
                                for (int cc = 0; cc < 3; cc++)
                                    if (cc > r) inU += blk[r * 3 + cc] * xb[cc];
#if DDPCA_V2_RECIP_DIAG
                                xb[r] = (rhs[r] - sU[r] - inU) * blk[9 + r];
#else
                                xb[r] = (rhs[r] - sU[r] - inU) / blk[r * 3 + r];
#endif
                            }
                        }
#pragma unroll
                        for (int r = 0; r < 3; r++) xn[r] = xb[r];
                    }
                    if (sl == 0) {
#pragma unroll
                        for (int r = 0; r < 3; r++)
                            if (r < gs) x[r0 + r] = xn[r];
                    }
                } else if (MODE == V2_BWD) {
                    double xn[3] = {0.0, 0.0, 0.0};
#pragma unroll
                    for (int r = 2; r >= 0; r--) {
                        if (r < gs) {
                            double inU = 0.0;
#pragma unroll
                            for (int cc = 0; cc < 3; cc++)
                                if (cc > r) inU += blk[r * 3 + cc] * xn[cc];
#if DDPCA_V2_RECIP_DIAG
                            xn[r] = (rhs[r] - sU[r] - inU) * blk[9 + r];
#else
                            xn[r] = (rhs[r] - sU[r] - inU) / blk[r * 3 + r];
#endif
                            if (sl == 0) x[r0 + r] = xn[r];
                        }
                    }
                } else {
                    // RESID / SPMV: one lane per row finishes with the in-group block
                    if (sl < gs) {
                        const int i = r0 + sl;
                        double sacc = sl == 0 ? sL[0] : (sl == 1 ? sL[1] : sL[2]);
                        if (MODE == V2_RESID) {
#pragma unroll
                            for (int cc = 0; cc < 3; cc++)
                                if (cc < sl) sacc += blk[sl * 3 + cc] * xo[cc];
                            y[i] = rhs[0] - (rhs[1] + sacc);
                        } else {
                            sacc += sl == 0 ? sU[0] : (sl == 1 ? sU[1] : sU[2]);
#pragma unroll
                            for (int cc = 0; cc < 3; cc++)
                                if (cc < gs) sacc += blk[sl * 3 + cc] * xo[cc];
                            y[i] = sacc;
                            if (w) acc += rhs[0] * sacc;
                        }
                    }
                }
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty[slot]);   // this warp is done reading the slot
        }
        if (STAGED && nst > 1) consumer_grid_barrier_exit(gbar);
    }
    if (MODE == V2_SPMV && partial) block_sum_to_partial(acc, partial);   // all warps, producer included
}

__global__ void k_extract_diag_inv2(Lvl2View A, double *__restrict__ dinv)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= A.ng) return;
    const GroupMeta2 m = A.meta[g];
    for (int r = 0; r < m.gs; r++) dinv[m.row0 + r] = 1.0 / A.BD[(size_t)g * kBlkStride + r * 3 + r];
}

}  // namespace ddpca
