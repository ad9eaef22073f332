// jgpu_device.cuh — device-side data model of the B200 token-passing decoder.
//
// Reference path being replaced (all paths relative to the idiap/juicer tree):
//   WFSTDecoderLite::processFrame           src/WFSTDecoderLite.cpp:311-372
//   WFSTDecoderLite::HMMInternalPropagation src/WFSTDecoderLite.cpp:376-484
//   WFSTDecoderLite::propagateToken         src/WFSTDecoderLite.cpp:491-605
//   HTKFlatModels::calcGMMOutput / logAdd   src/HTKFlatModels.cpp:226-293
//   Histogram                               src/Histogram.cpp:64-158
//
// Layout in HBM (one decoder handle, L lanes = utterances decoded in lock-step):
//   static, shared by all lanes : arcs int4{to | MULTI, w, in, out}, every state's row ordered
//                                 [epsilon arcs][tee-model arcs][other model arcs]
//                                 | states int4{first, n arcs, final weight, n_eps | n_tee << 16}
//                                 | hmm_info 8 x i32 per HMM | trP/SE per transition-matrix class
//                                 | GMM parameters transposed [d][comp][gmm]
//   per lane, double buffered   : active-instance list  inst_meta[2][cap] int4{arc, hmm | FRESH, to | MULTI, out label}
//                                 token planes          tok[2][S-1][cap] float4{score,ac,lm,path}
//   per lane, dense             : slotmap[nArcs] u32 = epoch stamp << 20 | list position + 1 — the GPU form of
//                                   WFSTTransition::hook, valid iff the stamp is the current epoch
//                                 state_key[nMulti] u64 (per-state max of arriving tokens, self-cleaning), for the
//                                   states that can see more than one arrival per frame (numbered first)
//   per lane, per frame scratch : arrival records (token plane + {via, state, label} plane, stored round after round)
//   per lane, per utterance     : word-boundary arena paths[cap_paths] (32 B records)
//
// Cost model behind the layout (measured on B200, tools/ubench/randmem.cu): random 32 B-sector
// accesses to DRAM saturate at ~36 G/s (1.15 TB/s) and 64-bit atomics at ~20 G/s whatever the
// footprint, L2-resident random accesses run at ~270 G/s, coalesced streams at ~200 G sectors/s.
// So the search path keeps random tables small (4 B per arc), touches them row-contiguously, and
// streams everything else with evict-first hints so that the tables stay in the 126 MB L2.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/juicer_b200.h"

#define JG_LZ (-3.402823466e+38f)
#define JG_MAX_ROUNDS 16
#ifndef JG_THREADS
#define JG_THREADS 256
#endif
enum { JG_MODE_IDLE = 0, JG_MODE_SEED = 1, JG_MODE_FRAME = 2, JG_FLAG_FINISH = 4 };
enum { JG_ERR_ACTIVE = 1, JG_ERR_ARRIVALS = 2, JG_ERR_PATHS = 4, JG_ERR_HIST = 8, JG_ERR_HUGE = 16,
       JG_ERR_WORDS = 64,       // the result-word pool of the batch was too small (the host grows it and decodes again)
       JG_ERR_LAZY = 128 };     // (self-check, frame_stats builds) k_internal used an acoustic score nobody had asked for
// errors a second pass with larger per-lane arenas cures (jgpu_engine.cu: decode_common re-decodes those utterances)
#define JG_ERR_RETRYABLE (JG_ERR_ACTIVE | JG_ERR_ARRIVALS | JG_ERR_PATHS | JG_ERR_WORDS)

typedef unsigned long long u64;

struct PathRec {          // 32 B: replaces Path (src/WFSTDecoderLite.h:39-55) minus the GC links
    int   prev, frame, label;
    float score, ac, lm;
    int   mark, pad1;     // mark: 0 when written, stamp of the last collection that reached it, JG_PATH_FREE on the free list
};
#define JG_PATH_FREE 0x7fffffff

struct ResHdr {           // 32 B per utterance
    int   status, n_frames;
    float score, ac, lm;
    int   error, word_off, n_words;   // the words of the best path are res_words[word_off .. word_off + n_words)
};

#define JG_MULTI 0x80000000u      // arcs.x / Arrival.q / inst_meta.z flag: the destination state can receive more
                                  // than one arrival per frame, so arrivals are max-reduced through state_key
#define JG_ROUND 0x40000000u      // same fields: the destination state has work for the expansion rounds (it is MULTI,
                                  // final, or has epsilon / tee out-arcs); every other exit token is finished inside
                                  // k_internal (word-boundary record) and only meets the commit
#define JG_STATE_MASK 0x3fffffff
// slotmap entry = (epoch & slot_emask) << slot_bits | list position + 1 and state key = epoch | orderable score | arrival
// id: the field widths are chosen at create from the size of the network (Dev::slot_bits, key_id_bits), so that a lane
// can hold an instance on every arc — the reference has no ceiling either (attachNetInst, src/WFSTDecoderLite.cpp:751-774)

#define JG_LR_CLASS 0x40000000   // hmm_info[0] flag: plain left-to-right topology (no skips), nStates <= 5
#define JG_FRESH 0x40000000   // inst_meta.y flag: only the entry token of this instance is valid

struct LaneCtl {
    int n_cur, n_huge, n_paths, flip;
    int n_r0;                 // records of round 0 that need the expansion round (fused mode: indices in r0_list)
    unsigned epoch;           // advances every non-idle step of this lane, never repeats
    int n_next;               // (n_next, n_arr[0]) sit in one aligned 64-bit word: k_internal bumps both with one atomic
    int n_arr[JG_MAX_ROUNDS + 2];    // arrivals feeding expansion round k (records are stored back to back)
    unsigned best_int;        // orderable max of emitting scores of this frame   (WFSTDecoderLite.cpp:417-418)
    unsigned best_ext;        // orderable max of entry scores of this frame      (:572-573)
    u64      best_final;      // key of the best arrival at a final state          (:513-520)
    float norm, thr_emit, thr_start;
    int mode, srow, frame, utt, error;
    int final_rec;            // arrival record behind best_final (found by the commit pass; -1 = none)
    int c_active_emit, c_active_end, c_end_proc, c_arcs, c_entry;
    int hist_count;
    int final_valid;
    float4 final_tok;
    long long s_active_models, s_active_emit, s_active_end, s_proc_emit, s_proc_end, s_arcs, s_entry,
        s_paths, s_frames, s_gmm;
    long long b_stats[10];    // sums over the utterances finished on this lane since the last batch reset
    // word-boundary arena collection (kept behind the per-step fields: their offsets are tuned to two 128 B lines)
    int n_free;               // records on the lane's free list (filled by k_gc_sweep)
    int gc_gen, gc_do;        // mark stamp of the collection in progress / whether this lane takes part in it
    int paths_recycled;       // records served from the free list in this utterance (statistics)
    int c_gmm;                // (GMM, frame) scores computed for this lane in the current step (lazy scorer)
    int pad_tail_[3];         // keeps sizeof(LaneCtl) off a multiple of 128 B: every CTA of every kernel reads the same
                              // fields of all lanes at start-up, and line-aligned control blocks put those hot lines on
                              // a subset of the L2 slices (measured: +4 us per kernel launch at a 384 B stride)
};

static_assert(sizeof(LaneCtl) % 16 == 0 && sizeof(LaneCtl) % 128 != 0, "LaneCtl stride: 16 B aligned, not line aligned");

struct Dev {
    // static tables
    const int4*  arcs;
    const int4*  states;       // {first arc, n arcs, final weight bits, n_eps | n_tee << 16}
    const float* arc_tee;      // per-arc tee weight of the arc's HMM, nullptr when no tee model exists
    const int*   hmm_info;     // [n_hmms][8] : nst | class<<8, gmm of states 1..3 | tee bits, gmm of states 4..6
    const uint2* hmm8;         // [n_hmms] the same for <= 5-state sets with < 64 K GMMs and < 4 K classes, in 8 bytes:
                               // {nst | lr << 3 | class << 4 | gmm1 << 16, gmm2 | gmm3 << 16} — a quarter of the bytes k_internal's
                               // per-instance gather pulls through L1 (nullptr when the set does not fit)
    const float* trp;          // [n_class][S*S]
    const int2*  se;           // [n_class][S]
    const float4* lr;          // [n_class][2]: left-to-right classes {a01,a11,a12,a22 | a23,a33,a34,-}
    int n_lr;                  // float4 entries of lr (2 per class)
    int n_arcs, n_states, init_state, n_hmms, n_gmms, S;
    unsigned init_multi;       // JG_MULTI when the initial state can receive more than one arrival per frame
    int n_multi;               // states are renumbered so that the multi-arrival ones are 0 .. n_multi-1: state_key has
                               // n_multi entries per lane (c3: 11 % of the states) and stays L2-sized
    // settings
    float start_beam, main_beam, end_beam, word_beam;
    int max_hyps, hist_min, hist_max, hist_nbins;
    int n_lanes, cap, cap_arr, cap_paths, cap_huge, n_rounds, max_frames, frame_stats;
    int slot_bits;                   // slotmap: low bits = list position + 1, the rest = epoch stamp
    unsigned slot_emask;
    int key_id_bits;                 // state key: low bits = arrival id (2 * (via arc + 1) + pass-through flag)
    unsigned key_emask;
    int huge_deg;
    int gc_threshold;                // a lane's arena is collected when more than this many records are in use
    int fuse_exits;                  // no end / word beam: exit tokens become arrivals inside k_internal
    int grid_internal, grid_other, grid_walk;   // CTAs of the chunk-scheduled kernels
    // per-lane state
    LaneCtl*  ctl;
    int4*     inst_meta;
    float4*   tok;
    unsigned char* live;       // [n_lanes][2][cap] per instance: bit i = emitting-state plane i holds a live token
    unsigned* slotmap;
    u64*      state_key;
    float4*   arr_tok;         // arrival records, two planes: token | {via arc (-1 seed, -2 dropped), state | JG_MULTI, out label, -}
    int4*     arr_meta;
    int*      r0_list;         // [n_lanes][cap_arr] arrival records of round 0 whose state is JG_ROUND (fused mode)
    int2*     huge;            // [n_lanes][cap_huge] {state, arrival record} of hub-like rows met by k_commit
    PathRec*  paths;
    int*      path_free;       // [n_lanes][cap_paths] free list: indices of dead word-boundary records
    int*      hist;
    const float* scores;       // [ring rows][n_gmms]
    // lazy acoustic scoring (HTKFlatModels::calcOutput is only ever called for states a live token asks for,
    // src/HTKFlatModels.cpp:226-262): whoever writes a token that can make a state ask for its score in the NEXT step
    // stamps need[lane][gmm] with the low byte of the lane's next epoch; k_gmm_lazy scores the stamped pairs of a
    // step between k_boundary and k_internal.  A stale stamp only costs an evaluation, so nothing is ever cleared.
    int lazy;                  // 0 (default): every GMM is scored for every frame, 16 frames ahead (k_gmm_scores)
    unsigned char* need;       // [need_stride][need_gp]: a lane's stamps are one contiguous row
    unsigned char* scored;     // [need_stride][need_gp] (frame_stats handles only) stamp of the last step the pair was scored
    int need_stride;           // lanes rounded up to 32
    int need_gp;               // GMMs rounded up to 32
    int* gmm_next;             // the scorer's work counter: next GMM to be taken by a warp (k_boundary resets it)
    const int* arc_g1;         // per arc: GMM of the first emitting state of its HMM, or -1 - hmm when the entry state
                               // has other emitting successors (then every emitting state is stamped)
    int* lane_stamp;           // [need_stride] stamp current in this step per lane, 0x100 when the lane scores nothing
    float* xtile;              // [need_stride][DP] the lanes' feature rows of this step, zero padded (k_boundary)
    int xtile_dp, feat_dim;
    const float* const* feat_base;   // device cell holding the feature base pointer of the running schedule chunk
    const int4*  sched;        // [n_steps + 1][n_lanes] {feature row, score row, flags, utt}
    int*      lane_step;       // [n_lanes] next schedule row of the lane (k_boundary)
    ResHdr*   res_hdr;
    JgpuWord* res_words;       // one pool for the batch, handed out by finish_utterance (no per-utterance word limit)
    int*      res_used;        // words handed out
    int       res_words_cap;
    int*      fstat_cnt;       // [n_lanes][max_frames][4]
    float*    fstat_best;      // [n_lanes][max_frames]
};

// ---- optional CTA timeline (compile with -DJG_TRACE; see tools/trace_report.py) -----------
#ifdef JG_TRACE
#define JG_TRACE_MARKS 8
struct TraceRec { int kid, block, smid, aux; unsigned long long t0, t1; unsigned mark[JG_TRACE_MARKS]; };   // marks: SM cycles after CTA start
__device__ TraceRec* g_trace;
__device__ unsigned g_trace_n, g_trace_cap;
__device__ int g_trace_on;
__device__ __forceinline__ unsigned long long jg_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct TraceScope {
    unsigned long long t0; long long c0; int kid, aux; unsigned mark[JG_TRACE_MARKS];
    __device__ __forceinline__ TraceScope(int k, int a) : kid(k), aux(a)
    {
        t0 = jg_now();
        c0 = clock64();
        for (int i = 0; i < JG_TRACE_MARKS; ++i) mark[i] = 0;
    }
    // stamp `i` of thread 0 in SM cycles since CTA start (kept the first time it is reached; globaltimer ticks too coarsely)
    __device__ __forceinline__ void at(int i) { if (threadIdx.x == 0 && mark[i] == 0) mark[i] = (unsigned)(clock64() - c0); }
    __device__ __forceinline__ ~TraceScope()
    {
        __syncthreads();
        if (threadIdx.x == 0 && g_trace_on) {
            const unsigned long long t1 = jg_now();
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            const unsigned i = atomicAdd(&g_trace_n, 1u);
            if (i < g_trace_cap) {
                TraceRec r; r.kid = kid; r.block = blockIdx.x + blockIdx.y * gridDim.x; r.smid = (int)smid; r.aux = aux; r.t0 = t0; r.t1 = t1;
                for (int k = 0; k < JG_TRACE_MARKS; ++k) r.mark[k] = mark[k];
                g_trace[i] = r;
            }
        }
    }
};
#define JG_TRACE_SCOPE(k, a) TraceScope _trace_scope(k, a)
#define JG_TRACE_AT(i) _trace_scope.at(i)
#else
#define JG_TRACE_SCOPE(k, a)
#define JG_TRACE_AT(i)
#endif

// ---- small helpers ----------------------------------------------------------------------
__device__ __forceinline__ unsigned f2o(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float o2f(unsigned o)
{
    unsigned u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(u);
}
__device__ __forceinline__ float4 null_tok()
{
    return make_float4(JG_LZ, JG_LZ, JG_LZ, __int_as_float(-1));
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// slotmap entries: the GPU form of WFSTTransition::hook (src/WFSTNetwork.h:50-51), valid iff stamped with the lane's epoch
__device__ __forceinline__ unsigned slot_entry(const Dev& d, unsigned epoch, int pos)
{
    return ((epoch & d.slot_emask) << d.slot_bits) | ((unsigned)pos + 1u);
}
__device__ __forceinline__ int slot_lookup(const Dev& d, unsigned sm, unsigned epoch)     // list position or -1
{
    const unsigned p = sm & ((1u << d.slot_bits) - 1u);
    return ((sm >> d.slot_bits) == (epoch & d.slot_emask) && p != 0u) ? (int)p - 1 : -1;
}

// warp-aggregated slot allocation: every thread of the warp must call it
__device__ __forceinline__ int warp_alloc(int* counter, bool want)
{
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (m == 0) return -1;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane_id() == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return want ? base + __popc(m & ((1u << lane_id()) - 1u)) : -1;
}

// two allocations per warp-iteration with both atomics in flight before either result is used
__device__ __forceinline__ void warp_alloc2(int* c1, bool w1, int* c2, bool w2, int& p1, int& p2)
{
    const unsigned m1 = __ballot_sync(0xffffffffu, w1), m2 = __ballot_sync(0xffffffffu, w2);
    const int l1 = m1 ? __ffs(m1) - 1 : 0, l2 = m2 ? __ffs(m2) - 1 : 1;
    int b1 = 0, b2 = 0;
    if (m1 && lane_id() == l1) b1 = atomicAdd(c1, __popc(m1));
    if (m2 && lane_id() == l2) b2 = atomicAdd(c2, __popc(m2));
    b1 = __shfl_sync(0xffffffffu, b1, l1);
    b2 = __shfl_sync(0xffffffffu, b2, l2);
    const unsigned lt = (1u << lane_id()) - 1u;
    p1 = w1 ? b1 + __popc(m1 & lt) : -1;
    p2 = w2 ? b2 + __popc(m2 & lt) : -1;
}
