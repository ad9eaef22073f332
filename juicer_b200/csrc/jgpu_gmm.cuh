// jgpu_gmm.cuh — diagonal-GMM log-likelihood, bit-exact to HTKFlatModels::calcGMMOutput
// (src/HTKFlatModels.cpp:226-262) and its private logAdd (:266-293).
//
//   out[row][g] = logAdd_{c < nComp(g)} ( -0.5 * sum_{d<D} ((x_d - mu_gcd)^2 * ivar_gcd) + det_gc )
//
// Exactness rules:
//   * the d-sum is a sequential fp32 accumulation with separate sub, mul, mul, add — written
//     with round-to-nearest intrinsics so ptxas can never contract it into an FMA.  The sub and
//     the two muls of two neighbouring dimensions go through the packed f32x2 forms (SASS FADD2 /
//     FMUL2, jg_acc4 below): IEEE per element, so the bits do not change, but a register-operand
//     FP32 instruction issues every other cycle per scheduler on sm_100 and the packed form
//     does two elements in that slot;
//   * "-0.5*sumxmu + det" (double, narrowed once, :254) is evaluated as fadd(fmul(-0.5f, s), det), which is
//     bit-identical for every pair of floats (see the comment at the store);
//   * components are folded in order 0..n-1 by logAdd: float diff, double threshold -18.42,
//     double log(1.0 + exp(diff)), narrowed once (:289).  No tree reduction anywhere.
// Parallelism comes from (Gaussians) x (feature rows), never from inside one Gaussian.
//
// Mapping: CTA = 256 threads; thread <-> one Gaussian slot (c, g_local) holding its mu/ivar
// row in registers for a tile of RT feature rows staged in shared memory (broadcast float4
// reads).  Per-component values go through shared memory to phase 2, where thread <->
// (row, gmm) runs the serial logAdd chain with all lanes busy.  Parameters are stored
// transposed [d][c][g] so that both phases read and write coalesced.
#pragma once

#include "jgpu_device.cuh"
#include "jgpu_softplus_table.h"

#ifndef JG_GMM_RT
#define JG_GMM_RT 64          // feature rows per CTA tile
#endif
#ifndef JG_GMM_UNROLL
#define JG_GMM_UNROLL 4       // feature rows in flight per thread (independent accumulation chains)
#endif
#ifndef JG_LOGADD_BRANCHFREE
#define JG_LOGADD_BRANCHFREE 0   // (1 + JG_GMM_P2 = 4 measured 27 % slower: the logAdd phase is bound by the L1 wavefronts of its
                                  //  table gathers, not by the latency of a chain — profiles/r02_logadd.md)
#endif
#ifndef JG_LOGSUM_PREFILTER
#define JG_LOGSUM_PREFILTER 0     // jg_mix_logsum first finds the logAdd steps that cannot change the accumulator (exact, same bits;
                                  // measured 6 % slower on the synthetic mixtures, whose components overlap: profiles/r02_logadd.md)
#endif
#ifndef JG_SP_SHARED
#define JG_SP_SHARED 1            // the scorer keeps the softplus table in shared memory (80-byte rows: a thread's 16-byte reads
                                  // of a random row spread over 8 bank groups)
#endif
#ifndef JG_GMM_PACKED
#define JG_GMM_PACKED 1       // packed f32x2 arithmetic in the d-sum (0: scalar instructions, same bits)
#endif
#define JG_PRAGMA_(x) _Pragma(#x)
#define JG_PRAGMA_UNROLL(n) JG_PRAGMA_(unroll n)
#define JG_GMM_DMAX 64        // max feature dimension held in registers

struct GmmDev {
    const float* mu;          // [DP][C][Gpad], zero padded
    const float* iv;          // [DP][C][Gpad], zero padded
    const float* det;         // [C][Gpad]
    const int*   ncomp;       // [Gpad]
    int n_gmms, g_pad, C, D, gpb;   // gpb = GMMs per 256-thread CTA = 256 / C (also the padding unit of g_pad)
    const double* softplus;   // [JG_SP_INTERVALS][8] Taylor table of log(1+exp(d)), see jgpu_softplus_table.h
};

// log(1.0 + exp(d)) for d in [-18.42, 0] in double, from a degree-7 Taylor table (296 intervals of
// 1/16).  Exhaustively compared with glibc's log(1.0 + exp(d)) over all 1.1e9 float32 arguments of the
// range: max |difference| 2.2e-16 (one ulp of 0.693), i.e. as accurate as the libm composite the
// reference calls (src/HTKFlatModels.cpp:289), at ~1/6 of the instructions of exp() + log().
// dynamic shared memory of the dense scorer: feature tile + per-component values (in floats, rounded to 16 bytes),
// then the softplus table
__host__ __device__ inline size_t gmm_vals_floats(int RT, int DP, int C, int gpb)
{
    return ((size_t)RT * DP + (size_t)C * (RT * gpb + (gpb & 31)) + 3) & ~(size_t)3;
}
// SMEM: the table is a shared-memory copy with rows JG_SP_SROW doubles apart (see gmm_scores_body).
#define JG_SP_SROW 10
__host__ __device__ inline size_t gmm_smem_bytes(int RT, int DP, int C, int gpb)
{
    return gmm_vals_floats(RT, DP, C, gpb) * sizeof(float) + (JG_SP_SHARED ? (size_t)JG_SP_INTERVALS * JG_SP_SROW * sizeof(double) : 0);
}
template <bool SMEM>
__device__ __forceinline__ double jg_softplus(const double* __restrict__ tab, double d)
{
    int k = (int)(-d * 16.0);
    k = min(k, JG_SP_INTERVALS - 1);
    const double t = d + ((double)k + 0.5) * 0.0625;
    const double2* T = reinterpret_cast<const double2*>(tab + k * (SMEM ? JG_SP_SROW : 8));
    double2 c01, c23, c45, c67;
    if (SMEM) { c01 = T[0]; c23 = T[1]; c45 = T[2]; c67 = T[3]; }
    else { c01 = __ldg(T); c23 = __ldg(T + 1); c45 = __ldg(T + 2); c67 = __ldg(T + 3); }
    double r = c67.y;
    r = fma(r, t, c67.x);
    r = fma(r, t, c45.y);
    r = fma(r, t, c45.x);
    r = fma(r, t, c23.y);
    r = fma(r, t, c23.x);
    r = fma(r, t, c01.y);
    r = fma(r, t, c01.x);
    return r;
}

// s + sum over the four dimensions of xv, IN ORDER, of ((x - mu)^2 * ivar): what the loop at
// src/HTKFlatModels.cpp:247-251 adds for them.  nmu = -mu (x - mu == x + (-mu) bit for bit, signed zeros included).
__device__ __forceinline__ float jg_acc4(float s, float4 xv, float2 nmu01, float2 nmu23, float2 iv01, float2 iv23)
{
#if JG_GMM_PACKED
    float2 d01 = __fadd2_rn(make_float2(xv.x, xv.y), nmu01);
    float2 d23 = __fadd2_rn(make_float2(xv.z, xv.w), nmu23);
    d01 = __fmul2_rn(d01, d01);
    d23 = __fmul2_rn(d23, d23);
    d01 = __fmul2_rn(d01, iv01);
    d23 = __fmul2_rn(d23, iv23);
    s = __fadd_rn(s, d01.x);
    s = __fadd_rn(s, d01.y);
    s = __fadd_rn(s, d23.x);
    return __fadd_rn(s, d23.y);
#else
    const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
    const float na[4] = {nmu01.x, nmu01.y, nmu23.x, nmu23.y}, va[4] = {iv01.x, iv01.y, iv23.x, iv23.y};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float xmu = __fadd_rn(xa[e], na[e]);
        s = __fadd_rn(s, __fmul_rn(__fmul_rn(xmu, xmu), va[e]));
    }
    return s;
#endif
}

// logAdd (src/HTKFlatModels.cpp:266-293).  Branch-free: some thread of a warp needs the softplus in practically every
// call (ncu: its instructions execute as often as the call itself), and without the early return the compiler can
// interleave the JG_GMM_P2 independent chains a thread folds side by side — each chain is ~250 cycles of dependent
// conversions, fp64 arithmetic and a table gather per component, and the logAdd phase was 45 % of the scorer's
// stall samples at 17 % of its instructions.  A diff below the threshold (down to -inf) indexes the last table
// interval and its value is discarded.
template <bool SMEM = false>
__device__ __forceinline__ float jg_log_add(const double* __restrict__ tab, float x, float y)
{
#if JG_LOGADD_BRANCHFREE
    const float hi = x < y ? y : x, lo = x < y ? x : y;
    const float diff = __fsub_rn(lo, hi);
    const double dd = (double)diff;
    const float r = (float)((double)hi + jg_softplus<SMEM>(tab, dd));
    return dd < -18.42 ? hi : r;                               // MINUS_LOG_THRESHOLD, compared in double
#else
    if (x < y) { const float t = x; x = y; y = t; }
    const float diff = __fsub_rn(y, x);
    if ((double)diff < -18.42) return x;                       // MINUS_LOG_THRESHOLD, compared in double
    return (float)((double)x + jg_softplus<SMEM>(tab, (double)diff));
#endif
}

// logAdd over the nc component values v[0], v[stride], ... IN ORDER (src/HTKFlatModels.cpp:252-255).
// A step whose component lies more than the threshold below the running maximum of the EARLIER components is a no-op
// in the reference: the accumulator is >= that maximum (logAdd never returns less than its larger operand), so the
// float difference the reference forms is <= fl(v - max), and its branch `diff < -18.42` returns the accumulator
// unchanged (:272-276).  Those steps are found first with four fp32 instructions each and only the others run the
// fp64 logAdd, in the same order — the same sequence of roundings, at a fraction of the table gathers and fp64
// chains (a Gaussian mixture is dominated by a few components for any one frame).
// (double)f < -18.42  <=>  f < -18.419998 (0xc1935c28, the smallest float above the double threshold).
template <bool SMEM>
__device__ __forceinline__ float jg_mix_logsum(const double* __restrict__ tab, const float* __restrict__ v, int stride, int nc)
{
    float lp = JG_LZ;
#if JG_LOGSUM_PREFILTER
    if (nc <= 32) {
        const float thr = __uint_as_float(0xc1935c28u);
        float m = JG_LZ;
        unsigned todo = 0u;
        for (int cc = 0; cc < nc; ++cc) {
            const float x = v[cc * stride];
            if (!(__fsub_rn(x, m) < thr)) todo |= 1u << cc;
            m = fmaxf(m, x);
        }
        while (todo) {
            const int cc = __ffs(todo) - 1;
            todo &= todo - 1u;
            lp = jg_log_add<SMEM>(tab, lp, v[cc * stride]);
        }
        return lp;
    }
#endif
    for (int cc = 0; cc < nc; ++cc) lp = jg_log_add<SMEM>(tab, lp, v[cc * stride]);
    return lp;
}

// rows: list of feature-row indices into x (row-major [*, D]); -1 = skip.  Output row i of
// the list goes to out[(out_base + i) * n_gmms + g].
// DP = feature dimension padded to a multiple of 4 with mu = ivar = x = 0: the padded terms
// add (0-0)^2*0 = +0.0f to a non-negative sum, which is exact.
// Two launch forms of the same code:
//   PERSIST = false : one CTA per (GMM group, row tile), grid (n_bx, n_by), NT = 256 threads — the kernel timed alone;
//   PERSIST = true  : a 1-D grid of at most one CTA per SM, CTA k walks the contiguous tile range
//                     [k*T/grid, (k+1)*T/grid) with the row tile as the inner index, so a Gaussian's parameters stay
//                     in registers over consecutive tiles.  Launched on the low-priority scoring stream with NT = 192
//                     it leaves room on every SM for the search kernels of the previous frame block (section 4 of
//                     the scorer is bound by FP32 issue slots, the search by memory latency).
// gpb = NT / C Gaussian mixtures per CTA.
template <int DP, int NT, bool PERSIST>
__device__ __forceinline__ void
gmm_scores_body(const GmmDev& g, int gpb, const float* __restrict__ x, const int* __restrict__ rows, int n_rows,
                float* __restrict__ out, long long out_base, int n_by)
{
    JG_TRACE_SCOPE(JGPU_K_GMM, 0);
    constexpr int RT = JG_GMM_RT;
    constexpr int D = DP;
    extern __shared__ float smem[];
    float* xs = smem;                          // [RT][DP]
    float* vals = smem + RT * DP;              // [C][RT*gpb + pad]
    __shared__ int row_id[RT];
    // softplus table: the logAdd phase reads one 64-byte row per (thread, component) at a data-dependent index; from
    // global memory that is one L1 wavefront per distinct row and instruction (as many wavefronts as all the feature
    // tile loads of phase 1, ncu r02f), from shared memory with padded rows it is a bank-conflict count
    double* sp_tab = reinterpret_cast<double*>(smem + gmm_vals_floats(RT, DP, g.C, gpb));
    if (JG_SP_SHARED) {
        for (int i = threadIdx.x; i < JG_SP_INTERVALS * 4; i += NT) {
            const int k = i >> 2, j = i & 3;
            reinterpret_cast<double2*>(sp_tab + k * JG_SP_SROW)[j] = __ldg(reinterpret_cast<const double2*>(g.softplus + k * 8) + j);
        }
    }

    const int C = g.C;
    const int cstride = RT * gpb + (gpb & 31);   // consecutive components land on different banks
    const int tid = threadIdx.x;
    const int c = tid / gpb, gl = tid - c * gpb;
    float2 nmu[D / 2], iv[D / 2];                // -mean and inverse variance of dimensions (2k, 2k+1)
    float det = 0.0f;
    int loaded_bx = -1;

    int t0, t1;
    if (PERSIST) {
        const long long T = (long long)((g.n_gmms + gpb - 1) / gpb) * n_by;
        t0 = (int)(T * blockIdx.x / gridDim.x);
        t1 = (int)(T * (blockIdx.x + 1) / gridDim.x);
    } else {
        t0 = blockIdx.x * n_by + blockIdx.y;
        t1 = t0 + 1;
    }
    for (int t = t0; t < t1; ++t) {
        const int bx = t / n_by, by = t - bx * n_by;
        const int g0 = bx * gpb;
        const int r0 = by * RT;

        // stage the feature tile
        if (tid < RT) {
            const int i = r0 + tid;
            int rid = -1;
            if (i < n_rows) rid = rows[i];
            row_id[tid] = rid;
        }
        __syncthreads();
        bool any = false;
        for (int i = 0; i < RT; ++i) any |= row_id[i] >= 0;
        if (!any) {
            if (PERSIST) { __syncthreads(); continue; }
            return;
        }
        for (int i = tid; i < RT * DP; i += NT) {
            const int r = i / DP, dd = i - r * DP;
            const int rid = row_id[r];
            xs[i] = (rid >= 0 && dd < g.D) ? x[(size_t)rid * g.D + dd] : 0.0f;
        }

        // phase 1: thread <-> Gaussian slot
        const int gi = g0 + gl;
        const bool active = c < C && gi < g.n_gmms;
        if (active && loaded_bx != bx) {
            const size_t plane = (size_t)C * g.g_pad, off = (size_t)c * g.g_pad + gi;
#pragma unroll
            for (int k = 0; k < D / 2; ++k) {
                nmu[k] = make_float2(-__ldg(g.mu + (2 * k) * plane + off), -__ldg(g.mu + (2 * k + 1) * plane + off));
                iv[k] = make_float2(__ldg(g.iv + (2 * k) * plane + off), __ldg(g.iv + (2 * k + 1) * plane + off));
            }
            det = __ldg(g.det + off);
        }
        loaded_bx = bx;
        __syncthreads();
        if (active) {
JG_PRAGMA_UNROLL(JG_GMM_UNROLL)
            for (int r = 0; r < RT; ++r) {
                const float4* xr = reinterpret_cast<const float4*>(xs + r * DP);
                float s = 0.0f;
#pragma unroll
                for (int q = 0; q < DP / 4; ++q) s = jg_acc4(s, xr[q], nmu[2 * q], nmu[2 * q + 1], iv[2 * q], iv[2 * q + 1]);
                // the reference evaluates -0.5*s + det in double and narrows (:254).  The fp32 form below is the same
                // value bit for bit: -0.5f*s is exact, and the sum of two floats rounded to double and then to float
                // equals the sum rounded to float directly — exact in double when the exponents differ by <= 29, and
                // beyond that both roundings return the larger operand.
                vals[c * cstride + r * gpb + gl] = __fadd_rn(__fmul_rn(-0.5f, s), det);
            }
        }
        __syncthreads();

        // phase 2: thread <-> (row, gmm); serial logAdd chain in component order (jg_mix_logsum)
        for (int p = tid; p < RT * gpb; p += NT) {
            const int r = p / gpb, l = p - r * gpb;
            if (g0 + l < g.n_gmms && row_id[r] >= 0) {
                const int nc = __ldg(g.ncomp + g0 + l);
                const float lp = JG_SP_SHARED ? jg_mix_logsum<true>(sp_tab, vals + p, cstride, nc)
                                              : jg_mix_logsum<false>(g.softplus, vals + p, cstride, nc);
                if (nc > 0) out[(size_t)(out_base + r0 + r) * g.n_gmms + g0 + l] = lp;
            }
        }
        if (PERSIST) __syncthreads();            // row_id / xs / vals are rewritten by the next tile
    }
}

template <int DP>
__global__ void __launch_bounds__(256, 2)
k_gmm_scores(GmmDev g, int gpb, const float* __restrict__ x, const int* __restrict__ rows, int n_rows,
             float* __restrict__ out, long long out_base, int n_by)
{
    gmm_scores_body<DP, 256, false>(g, gpb, x, rows, n_rows, out, out_base, n_by);
}

// =========================================================================================
// k_gmm_lazy: the scorer of the decode loop.  The reference evaluates a GMM only when a live state asks for it
// (calcGMMOutput behind a per-frame cache, src/HTKFlatModels.cpp:226-262: ~3.1k of 6000 GMMs per frame on c3);
// here the (GMM, lane) pairs stamped for this step (Dev::need, see jgpu_device.cuh) are scored once per step,
// between k_boundary and k_internal — a superset of what k_internal will read, counted per lane (total_gmm_evals).
//
//   warp   <-> one GMM at a time, taken from a device-wide counter: the rows per GMM differ by the dozen, and a static
//              split leaves a third of an SM's warps idle at the tail of every CTA (measured)
//   thread <-> component c = lane % CPW of that GMM, its mean / inverse-variance rows in registers; the 32 / CPW
//              groups of CPW threads work on different feature rows, JG_LAZY_U rows in flight per thread
//   rows    = the lanes whose stamp is current, compacted by ballot into a per-warp list
// so no SIMT slot is spent on a (GMM, lane) pair nobody asked for.  Arithmetic and operand order are those of
// k_gmm_scores above (same helpers), so the two produce the same bits.
//
// Feature tile: k_boundary gathers the lanes' feature rows of the step into one dense zero-padded tile
// [lanes][DP]; every CTA needs all of it.  It is staged with the TMA bulk-copy engine (cp.async.bulk, SASS UBLKCP)
// completing on an mbarrier, and MULTICAST across a thread-block cluster: each CTA of the cluster fetches 1 / CL of
// the tile once and the copy lands in the shared memory of all CL CTAs (JG_LAZY_CLUSTER = 1: plain per-CTA copy).
// =========================================================================================
#ifndef JG_LAZY_U
#define JG_LAZY_U 2
#endif
#ifndef JG_LAZY_CLUSTER
#define JG_LAZY_CLUSTER 2
#endif
#define JG_LAZY_WARPS 8

__device__ __forceinline__ unsigned jg_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void jg_mbar_init(void* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(jg_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void jg_mbar_expect_tx(void* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(jg_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void jg_mbar_wait(void* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "JG_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra JG_DONE;\n"
        "bra JG_WAIT;\n"
        "JG_DONE:\n"
        "}\n" ::"r"(jg_smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared memory of this CTA / of every CTA in `mask` of the cluster (same offsets in each)
__device__ __forceinline__ void jg_bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(jg_smem_u32(dst)), "l"(src), "r"(bytes), "r"(jg_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void jg_bulk_g2s_multicast(void* dst, const void* src, unsigned bytes, void* bar, unsigned short mask)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(jg_smem_u32(dst)), "l"(src), "r"(bytes), "r"(jg_smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ unsigned jg_cluster_rank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void jg_cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n"
                 "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct LazyArgs {
    const float* mu;           // [n_gmms][C][DP]  (row of one component contiguous, 16 B aligned)
    const float* iv;
    const float* det;          // [n_gmms][C]
    const int*   ncomp;
    const double* softplus;
    int n_gmms, C, cpw;        // cpw = threads per GMM group: smallest power of two >= C (<= 32)
    const unsigned char* need; // [need_stride][need_gp]
    unsigned char* scored;     // same shape, or nullptr: stamp of the step in which the pair was scored (self-check)
    int* gmm_next;             // work counter, zeroed by k_boundary
    int need_stride, need_gp, n_lanes;
    LaneCtl* ctl;
    const int* lane_stamp;     // [need_stride] the stamp that is current for each lane in this step (k_boundary); 0x100 = none
    const float* xtile;        // [need_stride][DP]
    float* scores;             // [n_lanes][n_gmms]
};

template <int DP, int CL>
__global__ void __launch_bounds__(JG_LAZY_WARPS * 32, (DP <= 40 ? 2 : 1)) k_gmm_lazy(LazyArgs a)
{
    JG_TRACE_SCOPE(JGPU_K_GMM, 0);
    extern __shared__ __align__(128) float lz_smem[];
    __shared__ __align__(8) unsigned long long lz_bar;
    const int Lp = a.need_stride, L = a.n_lanes;
    float* xs = lz_smem;                                           // [Lp][DP]
    float* vals = xs + (size_t)Lp * DP;                            // [warps][32][cpw + 1]
    unsigned short* rowlist = reinterpret_cast<unsigned short*>(vals + JG_LAZY_WARPS * 32 * (a.cpw + 1));   // [warps][Lp]
    int* s_stamp = reinterpret_cast<int*>(rowlist + JG_LAZY_WARPS * Lp);                                    // [Lp]
    int* s_cnt = s_stamp + Lp;                                     // [Lp]

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const unsigned tile_bytes = (unsigned)((size_t)Lp * DP * sizeof(float));
    if (tid == 0) jg_mbar_init(&lz_bar, 1);
    __syncthreads();
    if (CL > 1) jg_cluster_sync();                                 // every CTA of the cluster has its barrier ready
    if (tid == 0) {
        jg_mbar_expect_tx(&lz_bar, tile_bytes);
        if (CL > 1) {
            // my 1 / CL of the tile, delivered to all CTAs of the cluster
            const unsigned share = ((tile_bytes / CL) + 15u) & ~15u;
            const unsigned r = jg_cluster_rank();
            const unsigned b0 = min(r * share, tile_bytes), b1 = min(b0 + share, tile_bytes);
            if (b1 > b0)
                jg_bulk_g2s_multicast(reinterpret_cast<char*>(xs) + b0, reinterpret_cast<const char*>(a.xtile) + b0, b1 - b0, &lz_bar,
                                      (unsigned short)((1u << CL) - 1u));
        } else {
            jg_bulk_g2s(xs, a.xtile, tile_bytes, &lz_bar);
        }
    }
    // meanwhile: which stamp is current for each lane (0x100 never equals a byte: idle and seeding lanes)
    for (int l = tid; l < Lp; l += blockDim.x) {
        s_stamp[l] = l < L ? a.lane_stamp[l] : 0x100;
        s_cnt[l] = 0;
    }
    __syncthreads();

    const int cpw = a.cpw, rpi = 32 / cpw;                         // rows per warp iteration
    const int c = lane % cpw, hh = lane / cpw;
    const int vstride = cpw + 1;
    float* vw = vals + (size_t)wid * 32 * vstride;
    unsigned short* rl = rowlist + (size_t)wid * Lp;
    bool tile_ready = false;
    for (;;) {
        int g = 0;
        if (lane == 0) g = atomicAdd(a.gmm_next, 1);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= a.n_gmms) break;
        // ---- rows that asked for GMM g ----
        int n = 0;
        const unsigned char* nd = a.need + g;
        for (int lb = 0; lb < Lp; lb += 256) {                     // 8 stamp loads in flight, then the ballots
            int st8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int l = lb + k * 32 + lane;
                st8[k] = l < Lp ? (int)nd[(size_t)l * a.need_gp] : -1;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int l = lb + k * 32 + lane;
                const bool want = l < Lp && st8[k] == s_stamp[l];
                const unsigned m = __ballot_sync(0xffffffffu, want);
                if (want) {
                    rl[n + __popc(m & ((1u << lane) - 1u))] = (unsigned short)l;
                    atomicAdd(&s_cnt[l], 1);
                }
                n += __popc(m);
            }
        }
        if (n == 0) continue;
        // ---- this thread's Gaussian ----
        const int nc = __ldg(a.ncomp + g);
        const bool active = c < nc;
        float2 nmu[DP / 2], iv[DP / 2];
        float det = JG_LZ;
        {
            const float4* pm = reinterpret_cast<const float4*>(a.mu + ((size_t)g * a.C + (active ? c : 0)) * DP);
            const float4* pv = reinterpret_cast<const float4*>(a.iv + ((size_t)g * a.C + (active ? c : 0)) * DP);
#pragma unroll
            for (int q = 0; q < DP / 4; ++q) {
                const float4 m4 = __ldg(pm + q), v4 = __ldg(pv + q);
                nmu[2 * q] = make_float2(-m4.x, -m4.y); nmu[2 * q + 1] = make_float2(-m4.z, -m4.w);
                iv[2 * q] = make_float2(v4.x, v4.y); iv[2 * q + 1] = make_float2(v4.z, v4.w);
            }
            if (active) det = __ldg(a.det + (size_t)g * a.C + c);
        }
        if (!tile_ready) { jg_mbar_wait(&lz_bar, 0); tile_ready = true; }
        __syncwarp();                                              // the row list is complete
        for (int b0 = 0; b0 < n; b0 += 32) {
            const int nb = min(32, n - b0);
            // phase 1: JG_LAZY_U rows in flight per thread; the d-sum of each is sequential (:247-251)
            for (int i0 = 0; i0 < nb; i0 += rpi * JG_LAZY_U) {
                float s[JG_LAZY_U];
                int xo[JG_LAZY_U];                             // row offsets in float4 units
                const float4* xs4 = reinterpret_cast<const float4*>(xs);
#pragma unroll
                for (int u = 0; u < JG_LAZY_U; ++u) {
                    const int i = i0 + u * rpi + hh;
                    xo[u] = (i < nb ? (int)rl[b0 + i] : 0) * (DP / 4);
                    s[u] = 0.0f;
                }
#pragma unroll
                for (int q = 0; q < DP / 4; ++q) {
#pragma unroll
                    for (int u = 0; u < JG_LAZY_U; ++u)
                        s[u] = jg_acc4(s[u], xs4[xo[u] + q], nmu[2 * q], nmu[2 * q + 1], iv[2 * q], iv[2 * q + 1]);
                }
#pragma unroll
                for (int u = 0; u < JG_LAZY_U; ++u) {
                    const int i = i0 + u * rpi + hh;
                    if (i < nb && active) vw[i * vstride + c] = __fadd_rn(__fmul_rn(-0.5f, s[u]), det);   // see k_gmm_scores
                }
            }
            __syncwarp();
            // phase 2: thread <-> (row, GMM): logAdd over the components IN ORDER (:252-255, :266-293)
            if (lane < nb) {
                const float lp = jg_mix_logsum<false>(a.softplus, vw + lane * vstride, 1, nc);
                const int row = (int)rl[b0 + lane];
                a.scores[(size_t)row * a.n_gmms + g] = lp;
                if (a.scored) a.scored[(size_t)row * a.need_gp + g] = (unsigned char)s_stamp[row];
            }
            __syncwarp();
        }
    }
    __syncthreads();
    for (int l = tid; l < L; l += blockDim.x)
        if (s_cnt[l]) atomicAdd(&a.ctl[l].c_gmm, s_cnt[l]);
    if (!tile_ready) jg_mbar_wait(&lz_bar, 0);                     // never leave with a copy into this CTA in flight
    if (CL > 1) jg_cluster_sync();                                 // ... or while a peer still waits for my share
}
