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
//   static, shared by all lanes : arcs int4{to,w,in,out} | states int2{first,n} | final f32
//                                 | hmm_info 8 x i32 per HMM | trP/SE per transition-matrix class
//                                 | GMM parameters transposed [d][comp][gmm]
//   per lane, double buffered   : active-instance list  inst_arc[2][cap]
//                                 token planes          tok[2][S-1][cap] float4{score,ac,lm,path}
//   per lane, dense, self-cleaning:
//                                 arc2slot[nArcs]  u32  (the GPU form of WFSTTransition::hook)
//                                 entry_key[nArcs] u64  (atomicMax recombination of entry tokens)
//                                 state_key[nStates] u64 (per-state max of arriving tokens)
//   per lane, per frame scratch : exit list, arrival records, frontier lists, commit list
//   per lane, per utterance     : word-boundary arena paths[cap_paths] (32 B records)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/juicer_b200.h"

#define JG_LZ (-3.402823466e+38f)
#define JG_MAX_ROUNDS 16
#define JG_THREADS 256

enum { JG_MODE_IDLE = 0, JG_MODE_SEED = 1, JG_MODE_FRAME = 2, JG_FLAG_FINISH = 4 };
enum { JG_ERR_ACTIVE = 1, JG_ERR_ARRIVALS = 2, JG_ERR_PATHS = 4, JG_ERR_HIST = 8, JG_ERR_HUGE = 16,
       JG_ERR_FRAMES = 32 };

typedef unsigned long long u64;

struct PathRec {          // 32 B: replaces Path (src/WFSTDecoderLite.h:39-55) minus the GC links
    int   prev, frame, label;
    float score, ac, lm;
    int   pad0, pad1;
};

struct ResHdr {           // 32 B per utterance
    int   status, n_frames;
    float score, ac, lm;
    int   error, pad1, pad2;
};

struct LaneCtl {
    int n_cur, n_next, n_exit, n_arr, n_commit, n_touched, n_paths, flip;
    int n_front[JG_MAX_ROUNDS + 1];
    int n_huge[JG_MAX_ROUNDS + 1];
    unsigned best_int;        // orderable max of emitting scores of this frame   (WFSTDecoderLite.cpp:417-418)
    unsigned best_ext;        // orderable max of entry scores of this frame      (:572-573)
    u64      best_final;      // key of the best arrival at a final state          (:513-520)
    float norm, thr_emit, thr_start;
    int mode, srow, frame, utt, error, dirty;
    int c_active_emit, c_active_end, c_end_proc, c_arcs;
    int hist_count;
    int final_valid;
    float4 final_tok;
    long long s_active_models, s_active_emit, s_active_end, s_proc_emit, s_proc_end, s_arcs, s_entry,
        s_paths, s_frames, s_gmm;
    long long b_stats[10];    // sums over the utterances finished on this lane since the last batch reset
};

struct Dev {
    // static tables
    const int4*  arcs;
    const int2*  states;
    const float* state_final;
    const float* arc_tee;      // per-arc tee weight of the arc's HMM, nullptr when no tee model exists
    const int*   hmm_info;     // [n_hmms][8] : nst | class<<8, tee bits, gmm of states 1..6
    const float* trp;          // [n_class][S*S]
    const int2*  se;           // [n_class][S]
    int n_arcs, n_states, init_state, n_hmms, n_gmms, S;
    // settings
    float start_beam, main_beam, end_beam, word_beam;
    int max_hyps, hist_min, hist_max, hist_nbins;
    int n_lanes, cap, cap_arr, cap_paths, cap_huge, n_rounds, max_frames, frame_stats, max_words;
    int small_deg, huge_deg;
    // per-lane state
    LaneCtl*  ctl;
    int*      inst_arc;
    float4*   tok;
    unsigned* arc2slot;
    u64*      entry_key;
    u64*      state_key;
    int*      exit_arc;
    float4*   exit_tok;
    float4*   arr_tok;
    int*      arr_via;
    int2*     front;
    int2*     huge;
    int*      commit_arc;
    int*      touched;
    PathRec*  paths;
    int*      hist;
    const float* scores;       // [ring rows][n_gmms]
    const int4*  sched;        // [n_steps + 1][n_lanes] {feature row, score row, flags, utt}
    ResHdr*   res_hdr;
    JgpuWord* res_words;
    int*      fstat_cnt;       // [n_lanes][max_frames][4]
    float*    fstat_best;      // [n_lanes][max_frames]
};

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
