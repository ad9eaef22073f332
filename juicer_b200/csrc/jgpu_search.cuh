// jgpu_search.cuh — the per-frame token-passing kernels (sm_100a).
//
// One frame step of every lane is the launch sequence
//   k_boundary -> k_internal [-> k_filter] -> k_walk<0> x n_rounds -> k_walk<1> [-> k_commit_huge]
// which restates WFSTDecoderLite::processFrame (src/WFSTDecoderLite.cpp:311-372) as
// data-parallel passes.  All float arithmetic on scores is plain fp32 add/sub in the
// reference's per-token order (compiled with -fmad=false; there are no multiplies), so
// every token carries bit-identical scores to the CPU decoder.
//
// Work distribution: the per-lane work lists (active instances, arrival records) have very
// different lengths, and the work per item is uneven (a fresh instance reads one token, a
// word-end record walks 44 arcs).  Every kernel therefore runs a fixed grid (a multiple of the
// 148 SMs) over CHUNKS of 256 items of one lane; chunk c of the concatenation of all lanes'
// chunk lists goes to CTA c mod grid, so neighbouring chunks — similar work — spread over CTAs.
// The per-lane parameters a chunk needs (thresholds, epoch, list bases) are read once per CTA
// into shared memory: no dependent global load sits in front of a chunk.
#pragma once

#include "jgpu_device.cuh"

#ifndef JG_MAX_LANES
#define JG_MAX_LANES 512
#endif
#ifndef JG_RUN
#define JG_RUN 1                  // consecutive chunks of k_internal handed to one CTA (L1 reuse of per-lane tables)
#endif
#ifndef JG_WALK_CTAS
#define JG_WALK_CTAS 5            // resident CTAs per SM of k_walk (51 registers per thread; at 6 the commit spills since the
                                  // arrival ids and the demand stamps were added, and 5 vs 6 measured the same in round 1)
#endif
#ifndef JG_COMMIT_PREFETCH
#define JG_COMMIT_PREFETCH 0     // k_walk<1>: the next chunk's arrival records, state rows and state keys are fetched by cp.async
                                 // while the current chunk's arcs are walked (two of the four dependent round trips of a chunk).
                                 // Bit-exact and slower (108.4 us per launch against 86.4): the 14 KB of staging per CTA come out
                                 // of the L1 that the arc-row, state-row and slotmap gathers live in.  What stays from it: the
                                 // state row and the state key are requested together (91.5 -> 86.4 us).
#endif
#ifndef JG_WALK_ILP
#define JG_WALK_ILP 2            // arcs in flight per thread in the flattened arc list of k_walk
#endif
#ifndef JG_HUGE_ILP
#define JG_HUGE_ILP 2            // arcs in flight per thread in k_commit_huge (1: 32.8 us, 2: 29.0 us, 4: 31.1 us — 54 registers, two waves)
#endif
#ifndef JG_LR_SH
#define JG_LR_SH 64              // left-to-right class constants kept in shared memory by k_internal (float4 entries, 2 per class)
#endif
#ifndef JG_SCORES_LATE
#define JG_SCORES_LATE 1          // k_internal: the next chunk's score gathers are issued behind the first barrier of a chunk
#endif
#ifndef JG_DEFER
#define JG_DEFER 0                // k_internal: a chunk's stores wait one chunk for their allocation atomics (see the kernel;
                                  // measured slower, profiles/r02_defer_experiment.md)
#endif
#ifndef JG_INT_CTAS
#define JG_INT_CTAS 3             // resident CTAs per SM of k_internal<5> (register budget 64 K / (256 * CTAs))
#endif
#define JG_CH JG_THREADS          // items per chunk

// Per-lane views -------------------------------------------------------------------------
struct LaneView {
    LaneCtl* c;
    int4* meta_cur; int4* meta_nxt;
    float4* tok_cur; float4* tok_nxt;
    unsigned* slotmap;
    u64* skey;
    float4* arr_tok; int4* arr_meta;
    int2* huge;
    PathRec* paths;
    int* hist;
};

__device__ __forceinline__ LaneView lane_view(const Dev& d, int lane, LaneCtl* ctl = nullptr)
{
    LaneView v;
    v.c = ctl ? ctl : d.ctl + lane;
    const int flip = v.c->flip;
    const size_t cap = (size_t)d.cap, P = (size_t)(d.S - 1);
    v.meta_cur = d.inst_meta + ((size_t)lane * 2 + flip) * cap;
    v.meta_nxt = d.inst_meta + ((size_t)lane * 2 + (flip ^ 1)) * cap;
    v.tok_cur = d.tok + ((size_t)lane * 2 + flip) * P * cap;
    v.tok_nxt = d.tok + ((size_t)lane * 2 + (flip ^ 1)) * P * cap;
    v.slotmap = d.slotmap + (size_t)lane * d.n_arcs;
    v.skey = d.state_key + (size_t)lane * d.n_multi;
    v.arr_tok = d.arr_tok + (size_t)lane * d.cap_arr;
    v.arr_meta = d.arr_meta + (size_t)lane * d.cap_arr;
    v.huge = d.huge + (size_t)lane * d.cap_huge;
    v.paths = d.paths + (size_t)lane * d.cap_paths;
    v.hist = d.hist + (size_t)lane * d.hist_nbins;
    return v;
}

// aggregated counter bump: the threads of the warp that are here together share one
// atomicAdd.  All of them must target the SAME counter (one lane per chunk).
__device__ __forceinline__ int agg_inc(int* counter)
{
    const unsigned peers = __activemask();
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane_id() == leader) base = atomicAdd(counter, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    return base + __popc(peers & ((1u << lane_id()) - 1u));
}

// Word-boundary records (Path, src/WFSTDecoderLite.h:39-55): `n` records for the calling group come first from
// the lane's free list (records the last garbage collection found unreachable — the reference's refcounted
// free, :630-657 / collectPaths :699-747), then from the bump pointer n_paths.  Called by ONE thread of the
// group; record j of the group is path_index(a, free_list, j).
// `has_free` = the lane's free list was non-empty when the kernel started (it only shrinks during a step), read
// once per CTA: the common case — nothing on the list — costs exactly one atomicAdd on the bump pointer.
struct PathAlloc { int from_free, free_top, bump_base; };
__device__ __forceinline__ PathAlloc path_alloc(LaneCtl* c, int n, bool has_free)
{
    PathAlloc a;
    a.from_free = 0; a.free_top = 0; a.bump_base = 0;
    if (has_free) {
        const int old = atomicSub(&c->n_free, n);
        const int got = min(max(old, 0), n);
        if (got < n) atomicAdd(&c->n_free, n - got);          // give back what was not there
        a.from_free = got; a.free_top = old;
        if (got) atomicAdd(&c->paths_recycled, got);
    }
    if (a.from_free < n) a.bump_base = atomicAdd(&c->n_paths, n - a.from_free);
    return a;
}
__device__ __forceinline__ int path_index(const PathAlloc& a, const int* __restrict__ free_list, int j)
{
    return j < a.from_free ? free_list[a.free_top - 1 - j] : a.bump_base + (j - a.from_free);
}
// one record for every thread of the warp that is here (all of the same lane), one allocation for the group
__device__ __forceinline__ int path_alloc_here(LaneCtl* c, const int* __restrict__ free_list, bool has_free)
{
    const unsigned peers = __activemask();
    const int leader = __ffs(peers) - 1;
    PathAlloc a;
    a.from_free = 0; a.free_top = 0; a.bump_base = 0;
    if (lane_id() == leader) a = path_alloc(c, __popc(peers), has_free);
    a.from_free = __shfl_sync(peers, a.from_free, leader);
    a.free_top = __shfl_sync(peers, a.free_top, leader);
    a.bump_base = __shfl_sync(peers, a.bump_base, leader);
    return path_index(a, free_list, __popc(peers & ((1u << lane_id()) - 1u)));
}

// streaming (evict-first) accessors for data that is written once and read once per step
__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ int4 ld_stream(const int4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(int4* p, int4 v) { __stcs(p, v); }

// first arrival record of expansion round k (arrivals of earlier rounds are final by then)
__device__ __forceinline__ int arr_base(const LaneCtl* c, int round)
{
    int b = 0;
    for (int j = 0; j < round; ++j) b += c->n_arr[j];
    return b;
}

// ---- chunk scheduling --------------------------------------------------------------------
struct LaneSh {                   // per-CTA shared copy of what a chunk needs to know about its lane
    int pref[JG_MAX_LANES + 1];   // exclusive prefix of the per-lane chunk counts
    int cnt[JG_MAX_LANES];        // items in the lane's list
    unsigned epoch[JG_MAX_LANES];
    float f0[JG_MAX_LANES], f1[JG_MAX_LANES], f2[JG_MAX_LANES];   // kernel-specific
    int i0[JG_MAX_LANES], i1[JG_MAX_LANES], i2[JG_MAX_LANES];
};
// bit 31 of LaneSh::epoch (only the low 11 bits of the epoch are ever used as a stamp): the lane's
// word-boundary free list was not empty when the kernel started
#define JG_SH_HAS_FREE 0x80000000u

// the per-lane item counts cnt(l) must be in shared memory (and __syncthreads() NOT yet called); fills pref[0..L] and
// returns the number of chunks
template <class CntF>
__device__ __forceinline__ int chunk_scan_by(int* pref, int L, CntF cnt, int items = JG_CH)
{
    __syncthreads();
    if (threadIdx.x < 32) {                                  // warp 0: scan L values, L/32 per thread
        const int per = (L + 31) / 32;
        const int b = threadIdx.x * per;
        int sum = 0;
        for (int i = 0; i < per; ++i)
            if (b + i < L) sum += (cnt(b + i) + items - 1) / items;
        int incl = sum;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)threadIdx.x >= o) incl += t;
        }
        int run = incl - sum;
        if (threadIdx.x == 0) pref[0] = 0;
        for (int i = 0; i < per; ++i)
            if (b + i < L) { run += (cnt(b + i) + items - 1) / items; pref[b + i + 1] = run; }
    }
    __syncthreads();
    return pref[L];
}
__device__ __forceinline__ int chunk_scan(LaneSh& sh, int L, int items = JG_CH)
{
    return chunk_scan_by(sh.pref, L, [&](int l) { return sh.cnt[l]; }, items);
}

// k_internal keeps its lane table in DYNAMIC shared memory, sized by the number of lanes, so that at up to 256 lanes
// three CTAs per SM still fit beside the chunk buffers (LaneSh is sized for JG_MAX_LANES).
struct IntLane { int cnt; unsigned epoch; float norm, thr_emit, thr_start; int srow, flip, frame; };
#define JG_PREF_PAD ((JG_MAX_LANES + 1 + 3) & ~3)            // ints reserved for the chunk prefix (keeps IntLane 16-byte aligned)
__host__ __device__ inline size_t internal_smem_bytes(int S, int n_lanes)
{
    return (size_t)(JG_DEFER ? 3 : 2) * S * JG_THREADS * sizeof(float4) + JG_PREF_PAD * sizeof(int) + (size_t)n_lanes * sizeof(IntLane);
}

// lane of chunk `ch` (largest l with pref[l] <= ch), walked upwards from the lane of the CTA's previous chunk:
// a CTA's chunks are gridDim.x apart, i.e. usually several lanes apart, and a linear walk over pref[] was 18 % of
// the instructions of the commit (ncu source view, 65 steps per warp).
__device__ __forceinline__ int lane_of_chunk(const LaneSh& sh, int L, int lane, int ch)
{
    while (lane + 16 < L && sh.pref[lane + 16] <= ch) lane += 16;     // (strides instead of a bisection: no extra registers)
    while (sh.pref[lane + 1] <= ch) ++lane;
    return lane;
}

__device__ __forceinline__ u64 warp_max_u64(u64 v)
{
    for (int o = 16; o > 0; o >>= 1) {
        const u64 t = __shfl_xor_sync(0xffffffffu, v, o);
        if (t > v) v = t;
    }
    return v;
}

// state key = epoch stamp | orderable score (32 bits) | arrival id (key_id_bits): keys of older steps always lose
// the atomicMax and never compare equal, so state_key needs no per-frame cleaning (the host wipes a lane's table
// when its epoch stamp wraps).
// Arrival id = 2 * (via arc + 1) + pass-through flag (0 for the utterance seed): a property of the NETWORK, not of
// the order in which records were allocated, so that an exact score tie between two arrivals at one state is broken
// the same way in every run (the larger id wins; the reference keeps whichever it met first in its list order,
// src/WFSTDecoderLite.cpp:563-571).  An arc delivers an exit token (flag 0) and, for a tee model, a pass-through
// token (flag 1) at most once per frame each, except a pass-through arc whose source state is expanded again in a
// later round: then the later arrival has a score >= the earlier one, and when it is EQUAL its key is equal too —
// process_arc<0> sees that in the value atomicMax returns and drops the later record, so keys stay unique.
__device__ __forceinline__ unsigned arrival_id(int via, int pass_through)
{
    return ((unsigned)(via + 1) << 1) | (unsigned)pass_through;
}
__device__ __forceinline__ u64 state_key_of(const Dev& d, unsigned epoch, float score, unsigned id)
{
    return ((u64)(epoch & d.key_emask) << (32 + d.key_id_bits)) | ((u64)f2o(score) << d.key_id_bits) | (u64)id;
}

// Lazy acoustic scoring: "state with GMM g of this lane may ask for its score in the lane's next step"; races only
// ever write the same value.  A plain fire-and-forget byte store: reading the stamp first (most are already there)
// put a dependent L2 round trip in front of every chunk's barrier and cost far more than the stores (measured).
__device__ __forceinline__ void mark_need(const Dev& d, int lane, int g, unsigned epoch)
{
    d.need[(size_t)lane * d.need_gp + g] = (unsigned char)((epoch + 1u) & 0xffu);
}

// =========================================================================================
// k_boundary: one warp per lane.  (A) closes the previous step: statistics, best final
// token, back-trace when the schedule says the utterance is over (recognitionFinish,
// src/WFSTDecoderLite.cpp:230-309).  (B) opens this step: buffer swap, pruning thresholds
// (processFrame :318-339 + Histogram::calcThresh, src/Histogram.cpp:134-158), counter reset.
// =========================================================================================
__device__ void finish_utterance(const Dev& d, const LaneView& v, int lane)
{
    LaneCtl* c = v.c;
    const int utt = c->utt;
    if (utt < 0) return;
    ResHdr h;
    h.status = -1; h.n_frames = c->frame; h.score = h.ac = h.lm = JG_LZ;
    h.error = c->error; h.word_off = 0; h.n_words = 0;
    if (c->error) {
        h.status = JGPU_E_CAPACITY - 10 - (c->error << 8);
    } else if (c->final_valid) {
        const float4 best = c->final_tok;
        int n = 0;
        for (int p = __float_as_int(best.w); p >= 0; p = v.paths[p].prev) ++n;
        if (n == 0) {
            h.status = -2;                                   // :273-306: no word label on the path
        } else {
            // the chain goes into the batch's word pool, oldest word first: no per-utterance limit (the reference
            // allocates one DecHypHist per word, :279-301)
            const int off = atomicAdd(d.res_used, n);
            h.n_words = n;
            if (off < 0 || off + n > d.res_words_cap) {
                h.error = JG_ERR_WORDS;                       // the host grows the pool and decodes this utterance again
                h.status = JGPU_E_CAPACITY - 10 - (JG_ERR_WORDS << 8);
            } else {
                h.status = n; h.score = best.x; h.ac = best.y; h.lm = best.z;
                h.word_off = off;
                JgpuWord* w = d.res_words + off;
                int k = n;
                for (int p = __float_as_int(best.w); p >= 0; p = v.paths[p].prev) {
                    --k;
                    const PathRec r = v.paths[p];
                    JgpuWord o;
                    o.label = r.label; o.time = r.frame; o.score = r.score; o.ac = r.ac; o.lm = r.lm;
                    if (k == n - 1) { o.score = best.x; o.ac = best.y; o.lm = best.z; }   // :293-295
                    w[k] = o;
                }
            }
        }
    }
    d.res_hdr[utt] = h;
    if (!(h.error & JG_ERR_RETRYABLE)) {                     // (an utterance that is decoded again counts once)
        c->b_stats[0] += c->s_frames;        c->b_stats[1] += c->s_active_models;
        c->b_stats[2] += c->s_active_emit;   c->b_stats[3] += c->s_active_end;
        c->b_stats[4] += c->s_proc_emit;     c->b_stats[5] += c->s_proc_end;
        c->b_stats[6] += c->s_gmm;           c->b_stats[7] += c->s_arcs;
        c->b_stats[8] += c->s_entry;         c->b_stats[9] += c->s_paths;
    }
}

__global__ void k_reset_batch_stats(Dev d)
{
    const int lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane < d.n_lanes)
        for (int i = 0; i < 10; ++i) d.ctl[lane].b_stats[i] = 0;
}

__device__ float hist_thresh_warp(const Dev& d, const LaneView& v)
{
    // Histogram::calcThresh (src/Histogram.cpp:134-158), bins scanned from the top.
    const int nb = d.hist_nbins, maxN = d.max_hyps;
    if (v.c->hist_count <= maxN) return (float)((float)(d.hist_min) - 0.5);
    const int chunk = (nb + 31) / 32;
    const int l = lane_id();
    const int hi = nb - 1 - l * chunk;                       // my chunk: bins hi, hi-1, ... hi-chunk+1
    int sum = 0;
    for (int i = 0; i < chunk; ++i) {
        const int b = hi - i;
        if (b >= 0) sum += v.hist[b];
    }
    int incl = sum;                                          // inclusive scan over lanes
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (l >= o) incl += t;
    }
    const int excl = incl - sum;
    const unsigned m = __ballot_sync(0xffffffffu, incl >= maxN);
    float thr = (float)(d.hist_min);
    if (m) {
        const int w = __ffs(m) - 1;
        int found = 0;
        if (l == w) {
            int total = excl;
            for (int i = 0; i < chunk; ++i) {
                const int b = hi - i;
                if (b < 0) break;
                total += v.hist[b];
                if (total >= maxN) { found = b; break; }
            }
        }
        found = __shfl_sync(0xffffffffu, found, w);
        thr = (float)((float)(found + d.hist_min) - 0.5);
    }
    return thr;
}

// The schedule row is addressed by a per-lane device counter (reset by the host with every schedule
// chunk), so that the launch has no per-step argument and a whole block of steps replays as one CUDA graph.
__global__ void __launch_bounds__(32) k_boundary(Dev d)
{
    const int lane = blockIdx.x;
    const int l = lane_id();
    int step = 0;
    if (l == 0) step = d.lane_step[lane]++;
    // the lane's control block is worked on in shared memory (one coalesced read, one coalesced write back):
    // this kernel is a long chain of dependent scalar accesses to it
    __shared__ LaneCtl sc;
    static_assert(sizeof(LaneCtl) % 4 == 0, "LaneCtl is copied word by word");
    {
        const int* src = reinterpret_cast<const int*>(d.ctl + lane);
        int* dst = reinterpret_cast<int*>(&sc);
        for (int i = l; i < (int)(sizeof(LaneCtl) / 4); i += 32) dst[i] = src[i];
    }
    step = __shfl_sync(0xffffffffu, step, 0);
    __syncwarp();
    LaneView v = lane_view(d, lane, &sc);
    LaneCtl* c = v.c;
    const int prev_mode = c->mode;
    const int4 s = d.sched[(size_t)step * d.n_lanes + lane];
    const int mode = s.z & 3;

    // ---- (A) close the previous step ----------------------------------------------------
    if (l == 0 && prev_mode != JG_MODE_IDLE) {
        const int n_after = min(c->n_next, d.cap);
        const int n_arr_total = arr_base(c, d.n_rounds + 1);
        if (c->n_next > d.cap) c->error |= JG_ERR_ACTIVE;
        if (n_arr_total > d.cap_arr) c->error |= JG_ERR_ARRIVALS;
        if (c->n_paths > d.cap_paths) c->error |= JG_ERR_PATHS;
        const u64 key = c->best_final;
        if (key && c->final_rec >= 0) {                      // (the commit pass found the record behind the key)
            const int fr = c->final_rec;
            float4 t = v.arr_tok[fr];
            const float fw = __int_as_float(d.states[v.arr_meta[fr].y & JG_STATE_MASK].z);
            t.x += fw;                                        // :517-518
            t.z += fw;
            c->final_tok = t;
            c->final_valid = 1;
        } else {
            c->final_valid = 0;
        }
        if (prev_mode == JG_MODE_FRAME) {
            c->s_active_models += n_after;
            c->s_active_emit += c->c_active_emit;
            c->s_active_end += c->c_active_end;
            c->s_proc_emit += c->c_active_emit;
            c->s_proc_end += c->c_end_proc;
            c->s_arcs += c->c_arcs;
            c->s_entry += c->c_entry;
            c->s_gmm += d.lazy ? c->c_gmm : d.n_gmms;
            c->s_frames += 1;
            if (d.frame_stats) {
                if (c->frame < d.max_frames) {
                    int* f = d.fstat_cnt + ((size_t)lane * d.max_frames + c->frame) * 4;
                    f[0] = n_after; f[1] = c->c_active_emit; f[2] = c->c_active_end; f[3] = c->c_end_proc;
                    const float bi = o2f(c->best_int), be = o2f(c->best_ext);
                    d.fstat_best[(size_t)lane * d.max_frames + c->frame] = bi > be ? bi : be;
                }
            }
            c->frame += 1;
        }
        c->s_paths = c->n_paths + c->paths_recycled;
    }
    __syncwarp();
    if (l == 0 && (s.z & JG_FLAG_FINISH)) finish_utterance(d, v, lane);
    __syncwarp();

    // ---- (B) open this step -------------------------------------------------------------
    // (the host wipes a lane's state_key / slotmap tables when its 11-bit epoch stamp wraps, run_schedule)
    float thr_emit = JG_LZ;
    if (mode == JG_MODE_FRAME && d.max_hyps > 0) {
        thr_emit = hist_thresh_warp(d, v);                   // whole warp
        for (int b = l; b < d.hist_nbins; b += 32) v.hist[b] = 0;   // Histogram::reset
    }
    __syncwarp();
    if (l == 0) {
        if (prev_mode != JG_MODE_IDLE) {                     // the list built last step becomes current
            c->flip ^= 1;
            c->n_cur = min(c->n_next, d.cap);
        }
        c->n_next = 0; c->n_huge = 0; c->n_r0 = 0;
        for (int i = 0; i <= JG_MAX_ROUNDS + 1; ++i) c->n_arr[i] = 0;
        c->best_final = 0;
        c->final_rec = -1;
        c->c_gmm = 0;
        c->c_active_emit = c->c_active_end = c->c_end_proc = c->c_arcs = c->c_entry = 0;
        c->mode = mode;
        if (mode != JG_MODE_IDLE) c->epoch += 1;             // invalidates every arcdyn.slot of older steps
        if (mode == JG_MODE_SEED) {                          // recognitionStart :139-228
            c->utt = s.w;
            c->frame = 0;
            c->error = 0;
            c->n_cur = 0;                                    // previous utterance's instances are dropped (:148-158)
            c->n_paths = 0; c->n_free = 0; c->paths_recycled = 0;
            c->best_int = f2o(JG_LZ);
            c->best_ext = f2o(JG_LZ);
            c->norm = 0.0f; c->thr_emit = JG_LZ; c->thr_start = JG_LZ;
            c->hist_count = 0;
            c->final_valid = 0;
            c->s_active_models = c->s_active_emit = c->s_active_end = c->s_proc_emit = c->s_proc_end = 0;
            c->s_arcs = c->s_entry = c->s_paths = c->s_frames = c->s_gmm = 0;
            // propagateToken(&zeroToken, NULL) (:221-226) = an arrival at the initial state, walked by the
            // expansion rounds of this step with every threshold at LOG_ZERO
            v.arr_tok[0] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1));
            v.arr_meta[0] = make_int4(-1, d.init_state | (int)d.init_multi, 0, 0);
            c->n_arr[0] = 1;
            d.r0_list[(size_t)lane * d.cap_arr] = 0;          // the seed always goes through the expansion rounds
            c->n_r0 = 1;
            if (d.init_multi) v.skey[d.init_state] = state_key_of(d, c->epoch, 0.0f, arrival_id(-1, 0));
        } else if (mode == JG_MODE_FRAME) {                  // processFrame :318-339
            const float bi = o2f(c->best_int), bx = o2f(c->best_ext);
            const float be = bi > bx ? bi : bx;              // bestEmitScore at the end of the last frame
            const float norm = (be > JG_LZ ? be : 0.0f);
            float te;
            if (d.max_hyps > 0) {
                te = thr_emit - norm;
                if (d.main_beam > 0.0f && te < -d.main_beam) te = -d.main_beam;
                c->hist_count = 0;
            } else {
                te = (d.main_beam > 0.0f ? -d.main_beam : JG_LZ);
            }
            c->norm = norm;
            c->thr_emit = te;
            c->thr_start = (d.start_beam > 0.0f ? (be - d.start_beam) : JG_LZ);
            c->best_int = f2o(JG_LZ);                        // :905
            c->best_ext = f2o(JG_LZ);
            c->srow = d.lazy ? lane : s.y;                   // row of this lane's scores (lazy: one row per lane)
        }
    }
    if (mode == JG_MODE_SEED && d.max_hyps > 0)
        for (int b = l; b < d.hist_nbins; b += 32) v.hist[b] = 0;
    if (d.lazy) {
        // the lane's feature row of this step goes into the dense tile the scorer stages with one bulk copy
        if (mode == JG_MODE_FRAME) {
            const float* x = *d.feat_base + (size_t)s.x * d.feat_dim;
            for (int dd = l; dd < d.xtile_dp; dd += 32) d.xtile[(size_t)lane * d.xtile_dp + dd] = dd < d.feat_dim ? x[dd] : 0.0f;
        }
        if (l == 0) d.lane_stamp[lane] = mode == JG_MODE_FRAME ? (int)(c->epoch & 0xffu) : 0x100;
        if (l == 0 && lane == 0) *d.gmm_next = 0;
    }
    __syncwarp();
    {
        const int* src = reinterpret_cast<const int*>(&sc);
        int* dst = reinterpret_cast<int*>(d.ctl + lane);
        for (int i = l; i < (int)(sizeof(LaneCtl) / 4); i += 32) dst[i] = src[i];
    }
}

// =========================================================================================
// k_internal: one thread per active instance.  HMMInternalPropagation
// (src/WFSTDecoderLite.cpp:376-484) + the list walk of doHMMInternalPropagation (:899-935),
// with survivors compacted by warp ballot into the next list.  A live exit token becomes an
// arrival record {token, arc, destination state, output label} straight away — the instance
// record carries the arc's destination and label, so nothing is gathered for it.  With no
// end / word beam (FUSE) the record is final here (doHMMExternalPropagation :946-962 passes
// every live exit token); otherwise k_filter applies the two beams once bestEmit is known.
// =========================================================================================
template <int S>
__device__ __forceinline__ float4 viterbi_into(const float4 (&src)[S], const float* __restrict__ trp,
                                               int2 se, int j, int nst)
{
    // res = argmax_i src[i].score + trP[i][j] over i in [se.x, se.y), first wins (:393-406)
    float4 res = null_tok();
    bool have = false;
#pragma unroll
    for (int i = 0; i < S - 1; ++i) {
        if (i >= se.x && i < se.y && i < nst - 1) {
            const float tr = __ldg(trp + i * S + j);
            if (!have) {
                res = src[i];
                res.x = res.x + tr;
                res.y = res.y + tr;
                have = true;
            } else {
                const float tmp = src[i].x + tr;
                if (tmp > res.x) {
                    res = src[i];
                    res.x = tmp;
                    res.y = res.y + tr;
                }
            }
        }
    }
    return res;
}

// cp.async helpers: 16-byte global -> shared copies that bypass L1 and the register file; a thread
// only ever reads back what it copied itself, so cp.async.wait_group is the only synchronisation.
// The copies carry an L2 evict-first policy: the instance lists stream through once per step (~2 GB/s per lane)
// and must not push the arrival records, state rows and arc rows of the expansion kernels out of L2.
__device__ __forceinline__ u64 l2_evict_first_policy()
{
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, u64 policy)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
#ifdef JG_NO_L2_HINT
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
#else
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem_src), "l"(policy) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// plain (L1-allocating) 16- and 8-byte copies
__device__ __forceinline__ void cp_async_ca16(void* smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_ca8(void* smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

// Software pipeline over the CTA's chunks (chunk i = the i-th chunk this CTA owns):
//   iteration i:  registers <- shared buffer (i & 1)          [chunk i: instance record + token planes]
//                 cp.async chunk i+2 -> shared buffer (i & 1)  [one chunk of HBM latency ahead]
//                 wait for chunk i+1's copies, read its HMM ids, issue its hmm_info gathers
//                 Viterbi + pruning of chunk i
//                 issue chunk i+1's acoustic-score gathers (its hmm_info has landed by now)
//                 block-wide slot allocation (2 atomics per chunk) + counters, stores of chunk i
// so the only memory latency a chunk still waits for is that of its own two allocation atomics.
// LAZY: the opt-in per-step scorer is on (demand stamps + self-check); compiled out of the default kernels, whose
// register budget (80 at three CTAs per SM) has no room for it
template <int S, bool FUSE, bool LAZY>
__global__ void __launch_bounds__(JG_THREADS, (S <= 5 ? JG_INT_CTAS : 2)) k_internal(Dev d)
{
    JG_TRACE_SCOPE(JGPU_K_INTERNAL, 0);
    constexpr int P = S - 1;
    constexpr int NW = JG_THREADS / 32;
    constexpr bool DEFER = JG_DEFER != 0;
    constexpr int NPAR = DEFER ? 2 : 1;
    __shared__ int sh_w[NPAR][NW][7];                         // per warp: survivors, exits, packed counters, best, path records, round-0 entries, exits (copy)
    __shared__ int sh_base[NPAR][6];
    // dynamic shared memory (internal_smem_bytes): [2][P + 1][JG_THREADS] chunk buffers (plane 0 = instance record),
    // DEFER: [P + 1][JG_THREADS] results of the previous chunk, then the lane table (sized by the number of lanes)
    extern __shared__ float4 stage[];
    float4* const defer = stage + 2 * (P + 1) * JG_THREADS;
    int* const pref = reinterpret_cast<int*>(stage + (DEFER ? 3 : 2) * (P + 1) * JG_THREADS);
    IntLane* const sh = reinterpret_cast<IntLane*>(pref + JG_PREF_PAD);
    __shared__ int s_dflags[DEFER ? JG_THREADS : 1];
    __shared__ float4 s_lr[JG_LR_SH];                         // left-to-right class constants: the Viterbi of a chunk starts with
                                                              // them (an LDS instead of an L1 round trip; the host clears the class
                                                              // flag of every HMM when the table does not fit)
    const int L = d.n_lanes;
    const int tid = threadIdx.x, wid = tid >> 5;
    if (tid < min(d.n_lr, JG_LR_SH)) s_lr[tid] = __ldg(d.lr + tid);
    for (int l = tid; l < L; l += blockDim.x) {
        const LaneCtl* c = d.ctl + l;
        IntLane r;
        r.cnt = c->mode == JG_MODE_FRAME ? c->n_cur : 0;
        r.norm = c->norm; r.thr_emit = c->thr_emit; r.thr_start = c->thr_start;
        r.srow = c->srow; r.flip = c->flip; r.frame = c->frame;
        r.epoch = (c->epoch & ~JG_SH_HAS_FREE) | (c->n_free > 0 ? JG_SH_HAS_FREE : 0u);
        sh[l] = r;
    }
    const int total = chunk_scan_by(pref, L, [&](int l) { return sh[l].cnt; });
    const size_t cap = (size_t)d.cap;
    const bool hist_on = d.max_hyps > 0;
    const int G = gridDim.x;
    const u64 l2_stream = l2_evict_first_policy();

    // Which emitting-state planes of an instance hold a live token is one byte per instance (Dev::live, written by
    // whoever listed the instance): only those planes are copied — on c3 1.7 of 3 — and only those are written back.
    __shared__ unsigned char s_mask[2][JG_THREADS];           // the mask the copies of a staged chunk were issued with
    auto live_of = [&](int ch, int ln) -> unsigned {
        if (ch >= total) return 0u;
        const int k = (ch - pref[ln]) * JG_CH + tid;
        if (k >= sh[ln].cnt) return 0u;
        return d.live[((size_t)ln * 2 + sh[ln].flip) * cap + k];
    };
    // copies of chunk `ch` (lane `ln`) into buffer `buf`: record, entry token and the planes in `mask`; returns whether
    // this thread has an instance there
    auto issue = [&](int ch, int ln, int buf, unsigned mask) -> bool {
        bool v = false;
        if (ch < total) {
            const int k = (ch - pref[ln]) * JG_CH + tid;
            if (k < sh[ln].cnt) {
                v = true;
                const int flip = sh[ln].flip;
                const int4* meta_cur = d.inst_meta + ((size_t)ln * 2 + flip) * cap;
                const float4* tok_cur = d.tok + ((size_t)ln * 2 + flip) * P * cap;
                float4* dst = stage + (size_t)buf * (P + 1) * JG_THREADS + tid;
                cp_async16(dst, meta_cur + k, l2_stream);
                cp_async16(dst + JG_THREADS, tok_cur + k, l2_stream);
#pragma unroll
                for (int i = 1; i < P; ++i)
                    if ((mask >> i) & 1u) cp_async16(dst + (i + 1) * JG_THREADS, tok_cur + (size_t)i * cap + k, l2_stream);
                s_mask[buf][tid] = (unsigned char)mask;
            }
        }
        cp_async_commit();
        return v;
    };
    // This CTA's i-th chunk: chunks are dealt to CTAs in RUNS of JG_RUN consecutive chunks (same lane, mostly), so
    // that the lane's acoustic-score row and the HMM table lines stay in L1 across the run; run r of the CTA is
    // global run blockIdx.x + r * G.
    auto chunk_of = [&](int i) -> int {
        const int c_i = ((int)blockIdx.x + (i / JG_RUN) * G) * JG_RUN + (i % JG_RUN);
        return c_i < total ? c_i : total;
    };
    // lane of the i-th chunk, looked up once
    __shared__ unsigned short my_lane[JG_THREADS];
    {
        const int c_t = chunk_of(tid);
        int lo = 0, hi = L;                                   // largest l with pref[l] <= c_t
        if (c_t < total)
            while (lo + 1 < hi) {
                const int mid = (lo + hi) >> 1;
                if (pref[mid] <= c_t) lo = mid; else hi = mid;
            }
        my_lane[tid] = (unsigned short)lo;
    }
    __syncthreads();
    auto lane_of = [&](int i, int ch, int ln) -> int {
        if (i < JG_THREADS) return my_lane[i];
        if (ch < total)
            while (pref[ln + 1] <= ch) ++ln;                  // (more than 256 chunks per CTA: walk on)
        return ln;
    };

    // first int4 of an HMM's hmm_info row, from the 8-byte table when the model set has one
    auto load_h0 = [&](int hmm) -> int4 {
        if (S == 5 && d.hmm8) {
            const uint2 p = __ldg(d.hmm8 + hmm);
            return make_int4((int)((p.x & 7u) | ((p.x >> 4 & 0xfffu) << 8) | ((p.x & 8u) ? (unsigned)JG_LR_CLASS : 0u)),
                             (int)(p.x >> 16), (int)(p.y & 0xffffu), (int)(p.y >> 16));
        }
        return __ldg(reinterpret_cast<const int4*>(d.hmm_info) + hmm * 2);
    };
    int ch = chunk_of(0);
    int lane = lane_of(0, ch, 0);
    int lane1 = lane_of(1, chunk_of(1), lane), lane2 = lane_of(2, chunk_of(2), lane1);
    bool valid, valid1;
    {
        const unsigned m0 = live_of(ch, lane), m1 = live_of(chunk_of(1), lane1);
        valid = issue(ch, lane, 0, m0);
        valid1 = issue(chunk_of(1), lane1, 1, m1);
    }
    unsigned mask_ahead = live_of(chunk_of(2), lane2);        // always one chunk ahead of the copies it steers
    // hmm_info + scores of the first chunk (exposed once per CTA)
    int4 h0 = make_int4(2, 0, 0, 0), h1 = make_int4(0, 0, 0, 0);
    float outp[S - 2];
#pragma unroll
    for (int j = 0; j < S - 2; ++j) outp[j] = 0.0f;
    cp_async_wait<1>();
    if (valid) {
        const int hmm = reinterpret_cast<const int4*>(stage)[tid].y & ~JG_FRESH;
        h0 = load_h0(hmm);
        if (S > 5) h1 = __ldg(reinterpret_cast<const int4*>(d.hmm_info) + hmm * 2 + 1);
        const float* __restrict__ scores = d.scores + (size_t)sh[lane].srow * d.n_gmms;
        const int gm[6] = {h0.y, h0.z, h0.w, h1.y, h1.z, h1.w};
        const int nst0 = h0.x & 0xff;
#pragma unroll
        for (int j = 1; j < S - 1; ++j) outp[j - 1] = j < nst0 - 1 ? __ldg(scores + gm[j - 1]) : 0.0f;
    }

    JG_TRACE_AT(0);                                           // setup + first chunk's loads done
    // DEFER: the allocation atomics of chunk i are issued after its barrier and their results are only read one chunk
    // later — the threads that issued them publish them just before chunk i+1's barrier, and chunk i's survivors and
    // exit tokens, parked in shared memory meanwhile, are stored behind it.  One barrier per chunk, and no warp waits
    // for an L2 round trip (the second barrier and the wait behind it were 20 % of the kernel's stall samples).
    // The loop runs one extra, empty iteration to store the last chunk.
    bool pending = false;                                     // a chunk's results are parked (CTA-uniform)
    int lane_d = 0;                                           // ... and this is its lane
    int held0 = 0, held1 = 0, held2 = 0;                      // allocation results on their way (allocating threads only)
    for (int it = 0; ch < total || (DEFER && pending); ++it, ch = chunk_of(it)) {
        const int buf = it & 1;
        const int par = DEFER ? (it & 1) : 0;
        const float norm = sh[lane].norm, thr_emit = sh[lane].thr_emit, thr_start = sh[lane].thr_start;
        const unsigned epoch = sh[lane].epoch;
        LaneCtl* c = d.ctl + lane;
        // ---- registers <- buffer (chunk i) ----
        int4 meta = make_int4(0, 0, 0, 0);
        float4 old[S];
#pragma unroll
        for (int i = 0; i < S; ++i) old[i] = null_tok();
        if (valid) {
            const float4* src = stage + (size_t)buf * (P + 1) * JG_THREADS + tid;
            meta = *reinterpret_cast<const int4*>(src);
            old[0] = src[JG_THREADS];
            const unsigned mask = s_mask[buf][tid];            // (0 for a FRESH instance: only its entry token exists)
#pragma unroll
            for (int i = 1; i < P; ++i)
                if ((mask >> i) & 1u) old[i] = src[(i + 1) * JG_THREADS];
        }
        // ---- chunk i+2 -> the buffer just read; chunk i+1 has landed: start its hmm_info gathers ----
        const int ch2 = chunk_of(it + 2);
        lane2 = lane_of(it + 2, ch2, lane1);
        const bool valid2 = issue(ch2, lane2, buf, mask_ahead);
        {
            const int ch3 = chunk_of(it + 3);
            mask_ahead = live_of(ch3, lane_of(it + 3, ch3, lane2));
        }
        JG_TRACE_AT(1);                                       // registers loaded, next copies issued
        cp_async_wait<1>();
        JG_TRACE_AT(2);                                       // chunk i+1 landed
        int4 n0 = make_int4(2, 0, 0, 0), n1 = make_int4(0, 0, 0, 0);
        if (valid1) {
            const int hmm = reinterpret_cast<const int4*>(stage + (size_t)(buf ^ 1) * (P + 1) * JG_THREADS)[tid].y & ~JG_FRESH;
            n0 = load_h0(hmm);
            if (S > 5) n1 = __ldg(reinterpret_cast<const int4*>(d.hmm_info) + hmm * 2 + 1);
        }

        // ---- HMMInternalPropagation of chunk i ----
        float best = JG_LZ;
        int cnt_emit = 0, cnt_end = 0, cnt_hist = 0;
        bool survive = false, has_exit = false;
        const int nst = h0.x & 0xff;
        float4 nt[S];
        float4 ex = null_tok();
#pragma unroll
        for (int i = 0; i < S; ++i) nt[i] = null_tok();
        if (valid) {
            const int cls = (h0.x & ~JG_LR_CLASS) >> 8;
#pragma unroll
            for (int i = 1; i < P; ++i)
                if (i >= nst - 1) old[i] = null_tok();        // planes beyond this HMM's states hold stale data
            if (old[0].x > JG_LZ && old[0].x < thr_start) old[0] = null_tok();   // :915-918
            int nlive = 0;
            const bool lr = S == 5 && (h0.x & JG_LR_CLASS);
            float lrc[8];
            const float* __restrict__ trp = d.trp + (size_t)cls * S * S;
            const int2* __restrict__ se = d.se + (size_t)cls * S;
            if (lr) {
                const float4 c0 = s_lr[cls * 2], c1 = s_lr[cls * 2 + 1];
                lrc[0] = c0.x; lrc[1] = c0.y; lrc[2] = c0.z; lrc[3] = c0.w;
                lrc[4] = c1.x; lrc[5] = c1.y; lrc[6] = c1.z; lrc[7] = c1.w;
            }
#pragma unroll
            for (int j = 1; j < S - 1; ++j) {
                if (j < nst - 1) {
                    float4 res;
                    if (lr) {
                        // SEIndex[j] = [j-1, j+1): i = j-1 first, then i = j with strict '>' (:393-406)
                        const float a = lrc[2 * (j - 1)], b = lrc[2 * (j - 1) + 1];
                        res = old[j - 1];
                        res.x = res.x + a;
                        res.y = res.y + a;
                        const float tmp = old[j].x + b;
                        if (tmp > res.x) {
                            res = old[j];
                            res.x = tmp;
                            res.y = res.y + b;
                        }
                    } else {
                        res = viterbi_into<S>(old, trp, __ldg(se + j), j, nst);
                    }
                    res.x = res.x - norm;                                          // :408
                    if (res.x > thr_emit) {
                        const float o = outp[j - 1];                               // calcOutput :411
                        if (LAZY && d.frame_stats) {                               // self-check: the scorer did score this pair
                            const int gm_chk[6] = {h0.y, h0.z, h0.w, h1.y, h1.z, h1.w};
                            if (d.scored[(size_t)lane * d.need_gp + gm_chk[j - 1]] != (unsigned char)(epoch & 0xffu))
                                atomicOr(&c->error, JG_ERR_LAZY);
                        }
                        res.x = res.x + o;
                        res.y = res.y + o;
                        if (hist_on) {                                             // Histogram::addScore
                            int sc;
                            if (res.x < 0.0f) sc = (int)((double)res.x - 0.5);
                            else sc = (int)((double)res.x + 0.5);
                            if (sc > d.hist_max) atomicOr(&c->error, JG_ERR_HIST);
                            else if (sc >= d.hist_min) { atomicAdd(d.hist + (size_t)lane * d.hist_nbins + (sc - d.hist_min), 1); ++cnt_hist; }
                        }
                        if (res.x > best) best = res.x;
                        if (res.x > JG_LZ) { ++nlive; nt[j] = res; }
                    }
                }
            }
            cnt_emit += nlive;
            survive = nlive > 0;
            // exit state from the NEW emitting tokens (:443-483)
            {
                float4 res;
                if (lr) {                                 // SEIndex[N-1] = [N-2, N-1)
                    const float a = nst == 5 ? lrc[6] : nst == 4 ? lrc[4] : lrc[2];
                    res = nst == 5 ? nt[3] : nst == 4 ? nt[2] : nt[1];
                    res.x = res.x + a;
                    res.y = res.y + a;
                } else {
                    res = viterbi_into<S>(nt, trp, __ldg(se + (nst - 1)), nst - 1, nst);
                }
                if (res.x > JG_LZ) { ex = res; has_exit = true; ++cnt_end; }
            }
        }
        if (LAZY && survive) {
            // which states can ask for their score next step: state j when one of its predecessors holds a live token
            // (left-to-right: j - 1 or j); other topologies stamp every emitting state of a surviving instance
            const int gm_now[6] = {h0.y, h0.z, h0.w, h1.y, h1.z, h1.w};
            const bool lr_c = S == 5 && (h0.x & JG_LR_CLASS);
            bool prev_live = false;
#pragma unroll
            for (int j = 1; j < S - 1; ++j)
                if (j < nst - 1) {
                    const bool live = nt[j].x > JG_LZ;
                    if (!lr_c || live || prev_live) mark_need(d, lane, gm_now[j - 1], epoch);
                    prev_live = live;
                }
        }
        JG_TRACE_AT(3);                                       // Viterbi done
#if !JG_SCORES_LATE
        // ---- chunk i+1: its hmm_info has landed, start its acoustic-score gathers ----
        h0 = n0; h1 = n1;
        if (valid1) {
            const float* __restrict__ scores = d.scores + (size_t)sh[lane1].srow * d.n_gmms;
            const int gm[6] = {h0.y, h0.z, h0.w, h1.y, h1.z, h1.w};
            const int nstn = h0.x & 0xff;
#pragma unroll
            for (int j = 1; j < S - 1; ++j) outp[j - 1] = j < nstn - 1 ? __ldg(scores + gm[j - 1]) : 0.0f;
        }
#endif
        // ---- block-wide allocation: survivors -> next list, exit tokens -> arrival records of round 0;
        //      instances that die simply stop being listed, their slotmap entry goes stale with the epoch (:924-925)
        JG_TRACE_AT(4);                                       // next scores issued
        // With no end / word beam an exit token whose destination has nothing to do in the expansion rounds
        // (single arrival, not final, no epsilon / tee arcs) is finished here: its word-boundary record
        // (:497-509) is written now and the arrival record only meets the commit.  The others are listed for round 0.
        const bool to_round = FUSE && has_exit && (meta.z & (int)JG_ROUND) != 0;
        const bool need_path = FUSE && has_exit && !to_round && meta.z >= 0 && meta.w != 0;   // (MULTI: the commit writes it)
        const unsigned m_s = __ballot_sync(0xffffffffu, survive), m_e = __ballot_sync(0xffffffffu, has_exit);
        const unsigned m_p = __ballot_sync(0xffffffffu, need_path), m_r = __ballot_sync(0xffffffffu, to_round);
        // warp totals by vote and shuffle: REDUX (the __reduce_*_sync instructions) answers through the uniform datapath,
        // and its consumer was the single largest stall of the kernel (14 % of the samples, ncu source view)
        unsigned packed = 0u;
#pragma unroll
        for (int j = 1; j < S - 1; ++j) packed += (unsigned)__popc(__ballot_sync(0xffffffffu, nt[j].x > JG_LZ));
        if (hist_on) {
            int ch_ = cnt_hist;
            for (int o = 16; o > 0; o >>= 1) ch_ += __shfl_xor_sync(0xffffffffu, ch_, o);
            packed |= (unsigned)ch_ << 16;
        }
        unsigned best_o = f2o(best);
        for (int o = 16; o > 0; o >>= 1) best_o = max(best_o, __shfl_xor_sync(0xffffffffu, best_o, o));
        if (lane_id() == 0) {
            sh_w[par][wid][0] = __popc(m_s); sh_w[par][wid][1] = __popc(m_e); sh_w[par][wid][2] = (int)packed; sh_w[par][wid][3] = (int)best_o;
            sh_w[par][wid][4] = __popc(m_p); sh_w[par][wid][5] = __popc(m_r); sh_w[par][wid][6] = __popc(m_e);
        }
        if (DEFER && pending) {                               // the previous chunk's allocation has had a whole chunk to come back
            if (wid == 0) {
                if (lane_id() < 3) sh_base[par ^ 1][lane_id() == 2 ? 3 : lane_id()] = held0;
            } else if (tid == 32) {
                sh_base[par ^ 1][2] = held0; sh_base[par ^ 1][4] = held1; sh_base[par ^ 1][5] = held2;
            }
        }
        __syncthreads();
        // The allocation counters are bumped by DIFFERENT threads so that their round trips to L2 overlap: one
        // thread doing them in turn waits for each result before it issues the next (ATOMG -> STS pairs in SASS),
        // i.e. up to five dependent round trips per chunk with the whole CTA parked at the barrier below.
        if (wid == 0) {
            const int li = lane_id();
            if (li < 3) {                                     // lanes 0..2: next list, round-0 arrivals, round-0 work list
                const int col = li == 2 ? 5 : li;
                int tot = 0;
                for (int w = 0; w < NW; ++w) { const int a = sh_w[par][w][col]; sh_w[par][w][col] = tot; tot += a; }   // exclusive offsets of the warps
                int* ctr = li == 0 ? &c->n_next : li == 1 ? &c->n_arr[0] : &c->n_r0;
                int base = 0;
                if (tot) base = atomicAdd(ctr, tot);          // one predicated ATOMG for the three lanes
                if (DEFER) held0 = base;
                else sh_base[0][li == 2 ? 3 : li] = base;
            }
        } else if (tid == 32) {                               // word-boundary records
            int np = 0;
            for (int w = 0; w < NW; ++w) { const int a = sh_w[par][w][4]; sh_w[par][w][4] = np; np += a; }
            if (np) {
                const PathAlloc pa = path_alloc(c, np, (epoch & JG_SH_HAS_FREE) != 0);
                if (DEFER) { held0 = pa.bump_base; held1 = pa.from_free; held2 = pa.free_top; }
                else { sh_base[0][2] = pa.bump_base; sh_base[0][4] = pa.from_free; sh_base[0][5] = pa.free_top; }
            }
        } else if (tid == 64) {                               // counters nobody waits for
            int ne = 0, n_emit = 0, n_hist = 0;
            unsigned bo = 0;
            for (int w = 0; w < NW; ++w) {
                ne += sh_w[par][w][6];
                n_emit += sh_w[par][w][2] & 0xffff; n_hist += (unsigned)sh_w[par][w][2] >> 16;
                bo = max(bo, (unsigned)sh_w[par][w][3]);
            }
            if (bo > f2o(JG_LZ)) atomicMax(&c->best_int, bo);
            if (n_emit) atomicAdd(&c->c_active_emit, n_emit);
            if (ne) atomicAdd(&c->c_active_end, ne);
            if (FUSE && ne) atomicAdd(&c->c_end_proc, ne);
            if (n_hist) atomicAdd(&c->hist_count, n_hist);
        }
#if JG_SCORES_LATE
        // ---- chunk i+1: its hmm_info gathers (issued at the top of the iteration) have had the whole Viterbi and the
        //      first barrier to land; its acoustic-score gathers go out here, in the shadow of the allocation atomics ----
        h0 = n0; h1 = n1;
        if (valid1) {
            const float* __restrict__ scores = d.scores + (size_t)sh[lane1].srow * d.n_gmms;
            const int gm[6] = {h0.y, h0.z, h0.w, h1.y, h1.z, h1.w};
            const int nstn = h0.x & 0xff;
#pragma unroll
            for (int j = 1; j < S - 1; ++j) outp[j - 1] = j < nstn - 1 ? __ldg(scores + gm[j - 1]) : 0.0f;
        }
#endif
        // ---- what is stored now: this chunk (after a second barrier, once its allocation is known) or, DEFER, the
        //      previous one, whose place in shared memory this chunk's results take ----
        int4 e_meta = meta;
        float4 e_ex = ex;
        float4 e_nt[S];
#pragma unroll
        for (int i = 1; i < P; ++i) e_nt[i] = nt[i];
        unsigned e_flags = (survive ? 1u : 0u) | (has_exit ? 2u : 0u) | (need_path ? 4u : 0u) | (to_round ? 8u : 0u);
        int e_lane = lane;
        if (DEFER) {
            const unsigned now = e_flags;
            float4* dq = defer + tid;
            e_flags = pending ? (unsigned)s_dflags[tid] : 0u;
            e_lane = lane_d;
            s_dflags[tid] = (int)now;
            if (e_flags) e_meta = *reinterpret_cast<const int4*>(dq);
            if (now) *reinterpret_cast<int4*>(dq) = meta;
            if (e_flags & 2u) e_ex = dq[P * JG_THREADS];
            if (now & 2u) dq[P * JG_THREADS] = ex;
#pragma unroll
            for (int i = 1; i < P; ++i) {
                if (e_flags & 1u) e_nt[i] = dq[i * JG_THREADS];
                if (now & 1u) dq[i * JG_THREADS] = nt[i];
            }
            lane_d = lane;
            pending = ch < total;
        } else {
            __syncthreads();
        }
        JG_TRACE_AT(5);                                       // allocation known
        const int pe = DEFER ? par ^ 1 : 0;
        const bool e_survive = e_flags & 1u, e_has_exit = e_flags & 2u, e_need_path = e_flags & 4u, e_to_round = e_flags & 8u;
        const unsigned q_s = DEFER ? __ballot_sync(0xffffffffu, e_survive) : m_s, q_e = DEFER ? __ballot_sync(0xffffffffu, e_has_exit) : m_e;
        const unsigned q_p = DEFER ? __ballot_sync(0xffffffffu, e_need_path) : m_p, q_r = DEFER ? __ballot_sync(0xffffffffu, e_to_round) : m_r;
        const unsigned lt = (1u << lane_id()) - 1u;
        const int pos = sh_base[pe][0] + sh_w[pe][wid][0] + __popc(q_s & lt);
        const int e = sh_base[pe][1] + sh_w[pe][wid][1] + __popc(q_e & lt);
        const unsigned e_epoch = sh[e_lane].epoch;
        if (e_survive && pos < d.cap) {
            const int flip = sh[e_lane].flip;
            int4* meta_nxt = d.inst_meta + ((size_t)e_lane * 2 + (flip ^ 1)) * cap;
            float4* tok_nxt = d.tok + ((size_t)e_lane * 2 + (flip ^ 1)) * P * cap;
            st_stream(meta_nxt + pos, make_int4(e_meta.x, e_meta.y & ~JG_FRESH, e_meta.z, e_meta.w));
            tok_nxt[pos] = null_tok();                    // entry token consumed (:426-435); k_walk<1> may overwrite it
            unsigned nm = 0u;
#pragma unroll
            for (int i = 1; i < P; ++i)
                if (e_nt[i].x > JG_LZ) { st_stream(tok_nxt + (size_t)i * cap + pos, e_nt[i]); nm |= 1u << i; }
            d.live[((size_t)e_lane * 2 + (flip ^ 1)) * cap + pos] = (unsigned char)nm;
            d.slotmap[(size_t)e_lane * d.n_arcs + e_meta.x] = slot_entry(d, e_epoch, pos);
        }
        if (e_has_exit && e < d.cap_arr) {
            int via = e_meta.x;
            if (e_need_path) {
                PathAlloc pa;
                pa.bump_base = sh_base[pe][2]; pa.from_free = sh_base[pe][4]; pa.free_top = sh_base[pe][5];
                const int p = path_index(pa, d.path_free + (size_t)e_lane * d.cap_paths, sh_w[pe][wid][4] + __popc(q_p & lt));
                if (p < d.cap_paths) {
                    PathRec* pr = d.paths + (size_t)e_lane * d.cap_paths + p;
                    st_stream(reinterpret_cast<int4*>(pr), make_int4(__float_as_int(e_ex.w), sh[e_lane].frame, e_meta.w, __float_as_int(e_ex.x)));
                    st_stream(reinterpret_cast<int4*>(pr) + 1, make_int4(__float_as_int(e_ex.y), __float_as_int(e_ex.z), 0, 0));
                    e_ex.w = __int_as_float(p);
                } else {
                    via = -2;                             // arena full: the record is dropped, k_boundary flags the lane
                }
            }
            if (e_to_round) d.r0_list[(size_t)e_lane * d.cap_arr + sh_base[pe][3] + sh_w[pe][wid][5] + __popc(q_r & lt)] = e;
            d.arr_tok[(size_t)e_lane * d.cap_arr + e] = e_ex;
            d.arr_meta[(size_t)e_lane * d.cap_arr + e] = make_int4(via, e_meta.z, e_meta.w, 0);
            if (FUSE && e_meta.z < 0)                     // destination can see several arrivals this frame
                atomicMax(d.state_key + (size_t)e_lane * d.n_multi + (e_meta.z & JG_STATE_MASK),
                          state_key_of(d, e_epoch, e_ex.x, arrival_id(e_meta.x, 0)));
        }
        JG_TRACE_AT(6);                                       // stores issued: end of the chunk
        // (no barrier needed here: a warp only rewrites its own sh_w row, the allocating threads rewrite the offsets and
        //  sh_base after / before the next chunk's first barrier, which every warp reaches after reading its positions,
        //  and a thread only ever touches its own entries of the parked results)
        lane = lane1; lane1 = lane2;
        valid = valid1; valid1 = valid2;
    }
    cp_async_wait<0>();
}

// =========================================================================================
// k_filter (only with an end or word beam): the exit-token test of doHMMExternalPropagation
// (:946-962) once bestEmit of the frame is complete.  Records that fail are marked dropped;
// the others are max-reduced per destination state where that is needed.
// =========================================================================================
__global__ void __launch_bounds__(JG_THREADS, 6) k_filter(Dev d)
{
    JG_TRACE_SCOPE(JGPU_K_SEED, 0);
    __shared__ LaneSh sh;
    const int L = d.n_lanes, tid = threadIdx.x;
    for (int l = tid; l < L; l += blockDim.x) {
        const LaneCtl* c = d.ctl + l;
        sh.cnt[l] = c->mode == JG_MODE_FRAME ? min(c->n_arr[0], d.cap_arr) : 0;
        const float be = o2f(c->best_int);
        sh.f0[l] = (d.end_beam > 0.0f ? (be - d.end_beam) : JG_LZ);     // :349
        sh.f1[l] = (d.word_beam > 0.0f ? (be - d.word_beam) : JG_LZ);   // :350
        sh.epoch[l] = c->epoch;
    }
    const int total = chunk_scan(sh, L);
    int lane = 0;
    for (int ch = blockIdx.x; ch < total; ch += gridDim.x) {
        lane = lane_of_chunk(sh, L, lane, ch);
        const int e = (ch - sh.pref[lane]) * JG_CH + tid;
        int proc = 0;
        if (e < sh.cnt[lane]) {
            const float score = d.arr_tok[(size_t)lane * d.cap_arr + e].x;
            const int4 m = d.arr_meta[(size_t)lane * d.cap_arr + e];
            const float thr = m.z == 0 ? sh.f0[lane] : sh.f1[lane];                 // :952-962
            if (score > thr) {
                proc = 1;
                if (m.y < 0)
                    atomicMax(d.state_key + (size_t)lane * d.n_multi + (m.y & JG_STATE_MASK),
                              state_key_of(d, sh.epoch[lane], score, arrival_id(m.x, 0)));
            } else {
                d.arr_meta[(size_t)lane * d.cap_arr + e].x = -2;
            }
        }
        proc = __popc(__ballot_sync(0xffffffffu, proc != 0));
        if (lane_id() == 0 && proc) atomicAdd(&d.ctl[lane].c_end_proc, proc);
    }
}

// =========================================================================================
// External propagation (doHMMExternalPropagation :937-982 + propagateToken :491-605) as
// level-synchronous rounds over WFST states.
//   arrival  = a token reaching state q through arc `via` (exit of an instance, epsilon arc,
//              or tee pass-through).  A state that can receive more than one arrival per
//              frame (JG_MULTI, decided on the static network) max-reduces them through a
//              fire-and-forget 64-bit atomicMax on state_key (score bits | record index) and
//              only the record that owns the key expands the state; every other state has at
//              most one arrival per frame and needs no key at all.
//   k_walk<0>(k) = expansion round k over the records created in round k-1: word-boundary
//              record (:497-509), final-state candidate (:513-520), and the state's
//              PASS-THROUGH arcs only — epsilon arcs (:533-540) and tee models (:584-600) —
//              which create the arrivals of round k+1.  Rows are stored
//              [epsilon | tee-model | other model arcs], so this is a prefix of the row.
//   k_walk<1>  = commit over the records of all rounds: the record that owns its state at the
//              END of the closure walks the model arcs of the row once and writes the entry
//              tokens (:542-582).  An arc only ever receives candidates from the owners of its
//              source state, and a later owner beats an earlier one on every arc (fl(s+w) is
//              monotone in s), so the final owner's token IS the recombination result: no
//              per-arc atomic is needed.  The arc's instance is found through slotmap[arc]
//              (existing slot if the instance survived the internal phase this step, else a
//              new FRESH instance, attachNetInst :751-774); a row's slotmap entries are
//              contiguous, so the lookup costs one sequential read per state.
// =========================================================================================
template <int PASS, bool LAZY>
__device__ __forceinline__ void process_arc(const Dev& d, int lane, LaneCtl* c, unsigned epoch, float thr_end, float thr_word,
                                            int out_round, int out_base, int flip, const float4 tok, int b, const int4 a,
                                            unsigned sm, float& best, int& n_entry)
{
    const float w = __int_as_float(a.y);
    const float s = tok.x + w;
    if (PASS == 0) {
        float4 t;
        bool go;
        int pt = 0;
        if (a.z == 0) {                                       // epsilon input: :533-540
            t = make_float4(s, tok.y, tok.z + w, tok.w);
            go = s > thr_end;
        } else {                                              // tee model: :584-600
            const float tee = __ldg(d.arc_tee + b);
            const float s2 = s + tee;
            t = make_float4(s2, tok.y + tee, tok.z + w, tok.w);
            go = s2 > (a.w != 0 ? thr_word : thr_end);
            pt = 1;
        }
        if (go) {
            const int r = out_base + agg_inc(&c->n_arr[out_round]);
            if (r < d.cap_arr) {                              // overflow is flagged by k_boundary
                int via = b;
                if (a.x < 0) {
                    // the same arc fired in an earlier round with the same score (its source state changed owner to
                    // an arrival that rounds to the same sum): the earlier record stays, this one is dropped
                    const u64 key = state_key_of(d, epoch, t.x, arrival_id(b, pt));
                    if (atomicMax(d.state_key + (size_t)lane * d.n_multi + (a.x & JG_STATE_MASK), key) == key) via = -2;
                }
                d.arr_tok[(size_t)lane * d.cap_arr + r] = t;
                d.arr_meta[(size_t)lane * d.cap_arr + r] = make_int4(via, a.x, a.w, pt);
            }
        }
    } else {
        if (s > JG_LZ) {                                      // model arc: :542-582
            const float4 t = make_float4(s, tok.y, tok.z + w, tok.w);   // :568-570
            if (s > best) best = s;
            ++n_entry;
            if (LAZY) {                                       // the entry token makes the model's first state(s) ask
                const int g1 = __ldg(d.arc_g1 + b);
                if (g1 >= 0) mark_need(d, lane, g1, epoch);
                else {
                    const int* hi = d.hmm_info + (size_t)(-1 - g1) * 8;
                    const int ns = __ldg(hi) & 0xff;
                    for (int j = 1; j < ns - 1; ++j) mark_need(d, lane, __ldg(hi + (j <= 3 ? j : j + 1)), epoch);
                }
            }
            const size_t cap = (size_t)d.cap;
            float4* tok_nxt = d.tok + ((size_t)lane * 2 + (flip ^ 1)) * (size_t)(d.S - 1) * cap;
            const int slot = slot_lookup(d, sm, epoch);
            if (slot >= 0) {
                tok_nxt[slot] = t;                            // the instance survived the internal phase: plane 0 = entry token
            } else {
                const int pos = agg_inc(&c->n_next);
                if (pos < d.cap) {                            // overflow is flagged by k_boundary
                    int4* meta_nxt = d.inst_meta + ((size_t)lane * 2 + (flip ^ 1)) * cap;
                    st_stream(meta_nxt + pos, make_int4(b, (a.z - 1) | JG_FRESH, a.x, a.w));
                    tok_nxt[pos] = t;
                    d.live[((size_t)lane * 2 + (flip ^ 1)) * cap + pos] = 0;     // no emitting plane holds a token yet
                    d.slotmap[(size_t)lane * d.n_arcs + b] = slot_entry(d, epoch, pos);
                }
            }
        }
    }
}

template <int PASS, bool LAZY>
__global__ void __launch_bounds__(JG_THREADS, JG_WALK_CTAS) k_walk(Dev d, int round)
{
    JG_TRACE_SCOPE(PASS ? JGPU_K_COMMIT : JGPU_K_EXPAND, round);
    __shared__ LaneSh sh;
    __shared__ int s_off[JG_THREADS + 1];
    __shared__ int s_first[JG_THREADS];
    __shared__ float4 s_tok[JG_THREADS];
    __shared__ int s_wsum[JG_THREADS / 32];
    const int L = d.n_lanes;
    const int tid = threadIdx.x, wid = tid >> 5;
    for (int l = tid; l < L; l += blockDim.x) {
        const LaneCtl* c = d.ctl + l;
        const int mode = c->mode;
        int n = 0, rec0 = 0, out_base = 0;
        if (mode != JG_MODE_IDLE) {
            if (PASS == 0) {
                rec0 = arr_base(c, round);
                n = max(0, min(c->n_arr[round], d.cap_arr - rec0));
                out_base = rec0 + c->n_arr[round];
                if (round == 0 && d.fuse_exits) n = min(c->n_r0, n);   // only the records listed by k_internal / k_boundary
            } else {
                n = min(arr_base(c, d.n_rounds + 1), d.cap_arr);
                rec0 = c->frame;                              // (PASS 1 walks from record 0: the slot carries the frame)
                out_base = c->n_arr[0];                       // records below this index are round-0 arrivals
            }
        }
        sh.cnt[l] = n;
        sh.i0[l] = rec0; sh.i1[l] = out_base; sh.i2[l] = PASS == 0 ? c->frame : c->flip;
        float te = JG_LZ, tw = JG_LZ;
        if (mode == JG_MODE_FRAME) {
            const float be = o2f(c->best_int);
            te = (d.end_beam > 0.0f ? (be - d.end_beam) : JG_LZ);
            tw = (d.word_beam > 0.0f ? (be - d.word_beam) : JG_LZ);
        }
        sh.f0[l] = te; sh.f1[l] = tw;
        if (PASS == 1) {                                      // the commit has no use for the two beams: the slots carry the
            const u64 bf = c->best_final;                     // lane's best-final key, whose record this pass looks up
            sh.f0[l] = __uint_as_float((unsigned)(bf >> 32));
            sh.f1[l] = __uint_as_float((unsigned)bf);
        }
        sh.epoch[l] = (c->epoch & ~JG_SH_HAS_FREE) | (c->n_free > 0 ? JG_SH_HAS_FREE : 0u);
    }
    const int total_chunks = chunk_scan(sh, L);
    // lane of this CTA's i-th chunk, bisected once by thread i (walking pref[] from the previous chunk's lane, all
    // threads in step, was 17 % of the commit's instructions: a CTA's chunks are gridDim.x apart, ~20 lanes on c3)
    __shared__ unsigned short s_lane_of[JG_THREADS];
    {
        const long long c_t = (long long)blockIdx.x + (long long)tid * gridDim.x;
        int lo = 0, hi = L;                                   // largest l with pref[l] <= c_t
        if (c_t < total_chunks)
            while (lo + 1 < hi) {
                const int mid = (lo + hi) >> 1;
                if (sh.pref[mid] <= (int)c_t) lo = mid; else hi = mid;
            }
        s_lane_of[tid] = (unsigned short)lo;
    }
    __syncthreads();
    // The commit's prefetch (PASS 1): a chunk is four dependent round trips — arrival record, state row + state key,
    // [prefix], arc row + slotmap, token.  While the arcs of chunk i are walked, the record of chunk i+1 (issued at
    // the top of chunk i) has landed, and its state row and key are fetched; a thread only ever touches its own slots.
    constexpr bool PF = PASS == 1 && JG_COMMIT_PREFETCH != 0;
    __shared__ float4 p_tok[PF ? JG_THREADS : 1];
    __shared__ int4 p_meta[PF ? JG_THREADS : 1], p_st[PF ? JG_THREADS : 1];
    __shared__ u64 p_key[PF ? JG_THREADS : 1];
    const bool pf_on = PF && (long long)total_chunks <= (long long)JG_THREADS * gridDim.x;   // (s_lane_of covers every chunk)
    // this thread's record of the CTA's it2-th chunk: lane and index, or -1
    auto pf_slot = [&](int it2, int& ln) -> int {
        const long long ch2 = (long long)blockIdx.x + (long long)it2 * gridDim.x;
        if (ch2 >= total_chunks) return -1;
        ln = (int)s_lane_of[it2];
        const int e2 = ((int)ch2 - sh.pref[ln]) * JG_CH + tid;
        return e2 < sh.cnt[ln] ? e2 : -1;
    };
    auto pf_record = [&](int it2) {
        int ln = 0;
        const int e2 = pf_slot(it2, ln);
        if (e2 >= 0) {
            cp_async_ca16(&p_tok[tid], d.arr_tok + (size_t)ln * d.cap_arr + e2);
            cp_async_ca16(&p_meta[tid], d.arr_meta + (size_t)ln * d.cap_arr + e2);
        }
        cp_async_commit();
    };
    auto pf_state = [&](int it2) {                            // (its record has landed)
        int ln = 0;
        const int e2 = pf_slot(it2, ln);
        if (e2 >= 0) {
            const int my = p_meta[tid].y;
            const int q = my & JG_STATE_MASK;
            cp_async_ca16(&p_st[tid], &d.states[q]);
            if (my < 0) cp_async_ca8(&p_key[tid], d.state_key + (size_t)ln * d.n_multi + q);
        }
        cp_async_commit();
    };
    if (pf_on) {
        pf_record(0);
        cp_async_wait<0>();
        pf_state(0);
    }
    JG_TRACE_AT(0);                                           // setup done

    int lane = 0, it = 0;
    for (int ch = blockIdx.x; ch < total_chunks; ch += gridDim.x, ++it) {
        lane = it < JG_THREADS ? (int)s_lane_of[it] : lane_of_chunk(sh, L, lane, ch);
        LaneCtl* c = d.ctl + lane;
        const unsigned epoch = sh.epoch[lane];
        const float thr_end = sh.f0[lane], thr_word = sh.f1[lane];
        const int out_base = sh.i1[lane];
        float4* arr_tok = d.arr_tok + (size_t)lane * d.cap_arr;
        int4* arr_meta = d.arr_meta + (size_t)lane * d.cap_arr;
        // ---- (A) one thread per record ----
        const int e = (ch - sh.pref[lane]) * JG_CH + tid;
        bool valid = e < sh.cnt[lane];
        unsigned r = (unsigned)((PASS == 0 ? sh.i0[lane] : 0) + e);
        if (PASS == 0 && round == 0 && d.fuse_exits && valid) r = (unsigned)d.r0_list[(size_t)lane * d.cap_arr + e];
        int first = 0, deg = 0, arcs_done = 0;
        float4 tok = null_tok();
        u64 fin = 0;
        if (pf_on) cp_async_wait<0>();                        // this chunk's record, state row and key are in shared memory
        int4 m = make_int4(0, 0, 0, 0), st = make_int4(0, 0, 0, 0);
        u64 key_now = 0;
        if (valid) {
            if (pf_on) {
                tok = p_tok[tid]; m = p_meta[tid]; st = p_st[tid];
                if (m.y < 0) key_now = p_key[tid];
            } else {
                tok = arr_tok[r];
                m = arr_meta[r];                              // {via, q | MULTI, olab, -}
            }
        }
        if (pf_on) pf_record(it + 1);                         // (the slots just read are free again)
        if (valid) {
            const int q = m.y & JG_STATE_MASK;
            JG_TRACE_AT(1);                                   // record loaded
            if (!pf_on) {
                st = __ldg(&d.states[q]);
                if (m.y < 0) key_now = d.state_key[(size_t)lane * d.n_multi + q];
            }
            valid = m.x != -2;
            if (PASS == 1 && valid && m.x >= 0 && __int_as_float(st.z) > JG_LZ) {
                // best final token (:513-520): the rounds max-reduced (score + final weight | arrival id); the record
                // that made the winning proposal is the one whose key it is (arrival ids are unique per state)
                const u64 bf = ((u64)__float_as_uint(sh.f0[lane]) << 32) | (u64)__float_as_uint(sh.f1[lane]);
                if (bf != 0 && (((u64)f2o(tok.x + __int_as_float(st.z)) << 32) | (u64)arrival_id(m.x, m.w)) == bf) c->final_rec = (int)r;
            }
            if (valid && m.y < 0)                             // still the best arrival of q?
                valid = key_now == state_key_of(d, epoch, tok.x, arrival_id(m.x, m.w));
            if (valid) {
                const int n_eps = st.w & 0xffff, n_tee = (unsigned)st.w >> 16;
                JG_TRACE_AT(2);                               // state row (and key) loaded
                if (PASS == 0) {
                    if (m.z != 0) {                           // word boundary record: :497-509
                                                const int p = path_alloc_here(c, d.path_free + (size_t)lane * d.cap_paths, (sh.epoch[lane] & JG_SH_HAS_FREE) != 0);
                        if (p < d.cap_paths) {
                            PathRec* pr = d.paths + (size_t)lane * d.cap_paths + p;
                            st_stream(reinterpret_cast<int4*>(pr), make_int4(__float_as_int(tok.w), sh.i2[lane], m.z, __float_as_int(tok.x)));
                            st_stream(reinterpret_cast<int4*>(pr) + 1, make_int4(__float_as_int(tok.y), __float_as_int(tok.z), 0, 0));
                            tok.w = __int_as_float(p);
                            arr_tok[r].w = tok.w;
                        } else {
                            valid = false;                    // flagged by k_boundary
                            arr_meta[r].x = -2;
                        }
                    }
                    const float fw = __int_as_float(st.z);
                    if (valid && fw > JG_LZ && m.x >= 0) fin = ((u64)f2o(tok.x + fw) << 32) | (u64)arrival_id(m.x, m.w);   // :513-520 (not for the seed: trans == NULL)
                    if (valid) { first = st.x; deg = n_eps + n_tee; }
                } else {
                    if (d.fuse_exits && m.z != 0 && m.y < 0 && !(m.y & (int)JG_ROUND) && m.x >= 0 && r < (unsigned)sh.i1[lane]) {
                        // round-0 arrival at a multi-arrival state that the rounds skipped: the word-boundary record
                        // (:497-509) of the arrival that owns the state in the end is written here
                                                const int p = path_alloc_here(c, d.path_free + (size_t)lane * d.cap_paths, (sh.epoch[lane] & JG_SH_HAS_FREE) != 0);
                        if (p < d.cap_paths) {
                            PathRec* pr = d.paths + (size_t)lane * d.cap_paths + p;
                            st_stream(reinterpret_cast<int4*>(pr), make_int4(__float_as_int(tok.w), sh.i0[lane], m.z, __float_as_int(tok.x)));
                            st_stream(reinterpret_cast<int4*>(pr) + 1, make_int4(__float_as_int(tok.y), __float_as_int(tok.z), 0, 0));
                            tok.w = __int_as_float(p);
                        } else {
                            valid = false;                    // flagged by k_boundary
                        }
                    }
                    first = st.x + n_eps;
                    deg = valid ? st.y - n_eps : 0;
                    arcs_done = st.y;
                    if (deg >= d.huge_deg) {                  // hub-like row: left to k_commit_huge
                        const int h = atomicAdd(&c->n_huge, 1);
                        if (h < d.cap_huge) d.huge[(size_t)lane * d.cap_huge + h] = make_int2(q, (int)r);
                        else atomicOr(&c->error, JG_ERR_HUGE);
                        arr_tok[r].w = tok.w;                 // (k_commit_huge reads the token from the record)
                        deg = 0;
                    }
                }
            }
        }
        if (PASS == 0) {
            if (__any_sync(0xffffffffu, fin != 0)) {
                fin = warp_max_u64(fin);
                if (lane_id() == 0) atomicMax(&c->best_final, fin);
            }
        }
        JG_TRACE_AT(3);                                       // per-record work done
        // ---- block-wide exclusive prefix of the row lengths ----
        int incl = deg;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane_id() >= o) incl += t;
        }
        if (lane_id() == 31) s_wsum[wid] = incl;
        s_first[tid] = first; s_tok[tid] = tok;
        __syncthreads();
        int woff;
        {                                                     // exclusive prefix of the warp totals: one load + shuffles
            int ws = lane_id() < JG_THREADS / 32 ? s_wsum[lane_id()] : 0, wi = ws;
            for (int o = 1; o < JG_THREADS / 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane_id() >= o) wi += t;
            }
            woff = __shfl_sync(0xffffffffu, wi - ws, wid);
        }
        s_off[tid] = woff + incl - deg;
        if (tid == JG_THREADS - 1) s_off[JG_THREADS] = woff + incl;
        __syncthreads();
        const int total = s_off[JG_THREADS];
        JG_TRACE_AT(4);                                       // prefix done
        // ---- (B) the rows of the chunk as ONE flattened arc list, two arcs in flight per thread ----
        if (pf_on) { cp_async_wait<0>(); pf_state(it + 1); }  // the next chunk's record has landed: fetch its state row and key
        int n_entry = 0;
        float best = JG_LZ;
        for (int jb = 0; jb < total; jb += JG_WALK_ILP * JG_THREADS) {
            int b[JG_WALK_ILP], src[JG_WALK_ILP];
            int4 a[JG_WALK_ILP];
            unsigned sm[JG_WALK_ILP];
#pragma unroll
            for (int u = 0; u < JG_WALK_ILP; ++u) {
                sm[u] = 0u;
                const int j = jb + u * JG_THREADS + tid;
                b[u] = -1; src[u] = 0;
                if (j < total) {
                    int lo = 0, hi = JG_THREADS;              // largest src with s_off[src] <= j
                    while (lo + 1 < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (s_off[mid] <= j) lo = mid; else hi = mid;
                    }
                    src[u] = lo;
                    b[u] = s_first[lo] + (j - s_off[lo]);
                    a[u] = __ldg(&d.arcs[b[u]]);
                    if (PASS == 1) sm[u] = d.slotmap[(size_t)lane * d.n_arcs + b[u]];
                }
            }
#pragma unroll
            for (int u = 0; u < JG_WALK_ILP; ++u)
                if (b[u] >= 0)
                    process_arc<PASS, LAZY>(d, lane, c, epoch, thr_end, thr_word, round + 1, out_base, sh.i2[lane], s_tok[src[u]], b[u],
                                      a[u], sm[u], best, n_entry);
        }
        if (PASS == 1) {
            for (int o = 16; o > 0; o >>= 1) {                // (shuffles, not REDUX: see k_internal)
                arcs_done += __shfl_xor_sync(0xffffffffu, arcs_done, o);
                n_entry += __shfl_xor_sync(0xffffffffu, n_entry, o);
                best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
            }
            if (lane_id() == 0) {
                if (arcs_done) atomicAdd(&c->c_arcs, arcs_done);
                if (n_entry) atomicAdd(&c->c_entry, n_entry);
                if (best > JG_LZ) atomicMax(&c->best_ext, f2o(best));           // :572-573
            }
        }
        __syncthreads();                                      // s_off / s_tok are rewritten by the next chunk
        JG_TRACE_AT(5);                                       // first chunk done
    }
}

// hub-like rows met by the commit: all CTAs of the lane stride over the row.
template <bool LAZY>
__global__ void __launch_bounds__(JG_THREADS) k_commit_huge(Dev d)
{
    JG_TRACE_SCOPE(JGPU_K_EXPAND_HUGE, 0);
    const int lane = blockIdx.y;
    LaneCtl* c = d.ctl + lane;
    if (c->mode == JG_MODE_IDLE) return;
    const int n = min(c->n_huge, d.cap_huge);
    if (n == 0) return;
    const unsigned epoch = c->epoch;
    const int flip = c->flip;
    int n_entry = 0;
    float best = JG_LZ;
    for (int h = 0; h < n; ++h) {
        const int2 qr = d.huge[(size_t)lane * d.cap_huge + h];
        const float4 tok = d.arr_tok[(size_t)lane * d.cap_arr + qr.y];
        const int4 st = __ldg(&d.states[qr.x]);
        const int n_eps = st.w & 0xffff;
        // JG_HUGE_ILP arcs in flight per thread: the arc and its slotmap entry are independent loads, the walk is
        // otherwise one dependent round trip per arc
        const int stride = gridDim.x * blockDim.x, end = st.x + st.y;
        for (int b0 = st.x + n_eps + blockIdx.x * blockDim.x + threadIdx.x; b0 < end; b0 += stride * JG_HUGE_ILP) {
            int4 a[JG_HUGE_ILP];
            unsigned sm[JG_HUGE_ILP];
#pragma unroll
            for (int u = 0; u < JG_HUGE_ILP; ++u) {
                const int b = b0 + u * stride;
                if (b < end) {
                    a[u] = __ldg(&d.arcs[b]);
                    sm[u] = d.slotmap[(size_t)lane * d.n_arcs + b];
                }
            }
#pragma unroll
            for (int u = 0; u < JG_HUGE_ILP; ++u) {
                const int b = b0 + u * stride;
                if (b < end) process_arc<1, LAZY>(d, lane, c, epoch, JG_LZ, JG_LZ, 0, 0, flip, tok, b, a[u], sm[u], best, n_entry);
            }
        }
    }
    __syncwarp();
    n_entry = __reduce_add_sync(0xffffffffu, n_entry);
    for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane_id() == 0) {
        if (n_entry) atomicAdd(&c->c_entry, n_entry);
        if (best > JG_LZ) atomicMax(&c->best_ext, f2o(best));
    }
}

// =========================================================================================
// Garbage collection of the word-boundary arena — the GPU form of refPath / deRefPath and
// collectPaths (src/WFSTDecoderLite.cpp:630-657, 699-747).  The reference ref-counts every Path
// and sweeps every 100 frames; here a record is live iff it is reachable from a token that is
// alive between two steps, so a collection is  mark (walk the prev chains from every live
// token, stop at records already stamped)  +  sweep (unstamped records go to the lane's free
// list, from which path_alloc serves them again).  Nothing moves: indices held by tokens stay
// valid.  The host launches the three kernels between two frame steps every few dozen steps;
// k_gc_decide lets a lane take part only when its arena is filling up.
// =========================================================================================
__global__ void k_gc_decide(Dev d)
{
    const int lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= d.n_lanes) return;
    LaneCtl* c = d.ctl + lane;
    const int used = c->n_paths - c->n_free;
    c->gc_do = c->mode != JG_MODE_IDLE && c->n_paths <= d.cap_paths && used > d.gc_threshold;
    if (c->gc_do) {
        c->gc_gen += 1;
        if (c->gc_gen >= JG_PATH_FREE) c->gc_gen = 1;         // (2^31 collections: never in practice)
    }
}

__device__ __forceinline__ void gc_mark_chain(PathRec* paths, int p, int gen)
{
    while (p >= 0) {
        const int old = atomicExch(&paths[p].mark, gen);
        if (old == gen) break;                                // the rest of the chain is already stamped
        p = paths[p].prev;
    }
}

// grid (CTAs per lane, n_lanes).  Roots: every live token of the list built by the last step (the lane's NEXT
// buffer: k_boundary has not flipped yet) and the pending best final arrival.
__global__ void __launch_bounds__(JG_THREADS) k_gc_mark(Dev d)
{
    const int lane = blockIdx.y;
    LaneCtl* c = d.ctl + lane;
    if (!c->gc_do) return;
    const int gen = c->gc_gen;
    const size_t cap = (size_t)d.cap;
    const int P = d.S - 1;
    const int flip = c->flip ^ 1;
    const float4* tok = d.tok + ((size_t)lane * 2 + flip) * P * cap;
    PathRec* paths = d.paths + (size_t)lane * d.cap_paths;
    const int n = min(c->n_next, d.cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 t0 = tok[i];
        if (t0.x > JG_LZ) gc_mark_chain(paths, __float_as_int(t0.w), gen);
        const unsigned lv = d.live[((size_t)lane * 2 + flip) * cap + i];   // planes that hold a live token
        for (int j = 1; j < P; ++j)
            if ((lv >> j) & 1u) gc_mark_chain(paths, __float_as_int(tok[(size_t)j * cap + i].w), gen);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // the pending best final arrival (resolved by the next k_boundary)
        if (c->best_final && c->final_rec >= 0) gc_mark_chain(paths, __float_as_int(d.arr_tok[(size_t)lane * d.cap_arr + c->final_rec].w), gen);
        if (c->final_valid) gc_mark_chain(paths, __float_as_int(c->final_tok.w), gen);
    }
}

__global__ void __launch_bounds__(JG_THREADS) k_gc_sweep(Dev d)
{
    const int lane = blockIdx.y;
    LaneCtl* c = d.ctl + lane;
    if (!c->gc_do) return;
    const int gen = c->gc_gen;
    PathRec* paths = d.paths + (size_t)lane * d.cap_paths;
    int* free_list = d.path_free + (size_t)lane * d.cap_paths;
    const int n = min(c->n_paths, d.cap_paths);
    for (int i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
        const int i = i0 + threadIdx.x;
        bool dead = false;
        if (i < n) {
            const int mk = paths[i].mark;
            dead = mk != gen && mk != JG_PATH_FREE;
        }
        const unsigned m = __ballot_sync(0xffffffffu, dead);
        if (m) {
            int base = 0;
            if (lane_id() == 0) base = atomicAdd(&c->n_free, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (dead) {
                free_list[base + __popc(m & ((1u << lane_id()) - 1u))] = i;
                paths[i].mark = JG_PATH_FREE;
            }
        }
    }
}

// =========================================================================================
// Streaming partial results — tracePartialPath / traceWinningPaths (src/WFSTDecoderLite.cpp:822-897) on demand.
// The word-boundary records every live hypothesis has in its history can no longer change; the reference finds
// the newest of them by walking back from one token per active instance (the first one that has a path, :843-850)
// and counting visits per record (jointCount == nActiveInsts, :853-855).  Same here, for one lane between two
// steps (jgpu_partial_result): reset the counters, count, find the newest common record, write its chain out.
// scratch[0] = roots (instances), [1] = an instance without any word in its history (the reference walks off its
// token array there, :846-850: reported as "not traceable"), [2] = head record + 1.
// =========================================================================================
#define JG_PT_FLAG 0x40000000

__global__ void k_partial_reset(Dev d, int lane, int* scratch)
{
    const LaneCtl* c = d.ctl + lane;
    PathRec* paths = d.paths + (size_t)lane * d.cap_paths;
    const int n = min(c->n_paths, d.cap_paths);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) paths[i].pad1 = 0;
}

__global__ void k_partial_count(Dev d, int lane, int* scratch)
{
    const LaneCtl* c = d.ctl + lane;
    const size_t cap = (size_t)d.cap;
    const int P = d.S - 1;
    const int flip = c->flip;                                 // the streaming interface ends every call with the closing
    const float4* tok = d.tok + ((size_t)lane * 2 + flip) * P * cap;       // k_boundary, which has made the last step's list current
    PathRec* paths = d.paths + (size_t)lane * d.cap_paths;
    const int n = min(c->n_cur, d.cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned lv = 1u | d.live[((size_t)lane * 2 + flip) * cap + i];     // entry plane + the live emitting planes
        int root = -1;
        bool live = false;
        for (int p = 0; p < P && root < 0; ++p) {
            if (!((lv >> p) & 1u)) continue;
            const float4 t = tok[(size_t)p * cap + i];
            if (t.x > JG_LZ) { live = true; if (__float_as_int(t.w) >= 0) root = __float_as_int(t.w); }
        }
        if (!live) continue;                                  // (cannot happen for a listed instance)
        atomicAdd(&scratch[0], 1);
        if (root < 0) { scratch[1] = 1; continue; }
        for (int p = root; p >= 0; p = paths[p].prev) atomicAdd(&paths[p].pad1, 1);
    }
}

// a record every root passes through flags its predecessor: the newest common record is the one nobody flags
__global__ void k_partial_flag(Dev d, int lane, int* scratch)
{
    const LaneCtl* c = d.ctl + lane;
    PathRec* paths = d.paths + (size_t)lane * d.cap_paths;
    const int n = min(c->n_paths, d.cap_paths), roots = scratch[0];
    if (roots == 0 || scratch[1]) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if ((paths[i].pad1 & ~JG_PT_FLAG) == roots && paths[i].mark != JG_PATH_FREE && paths[i].prev >= 0)
            atomicOr(&paths[paths[i].prev].pad1, JG_PT_FLAG);
}

__global__ void k_partial_head(Dev d, int lane, int* scratch)
{
    const LaneCtl* c = d.ctl + lane;
    const PathRec* paths = d.paths + (size_t)lane * d.cap_paths;
    const int n = min(c->n_paths, d.cap_paths), roots = scratch[0];
    if (roots == 0 || scratch[1]) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (paths[i].pad1 == roots && paths[i].mark != JG_PATH_FREE) scratch[2] = i + 1;      // common and not flagged: unique
}

__global__ void k_partial_emit(Dev d, int lane, int slot, int* scratch)
{
    const LaneCtl* c = d.ctl + lane;
    const PathRec* paths = d.paths + (size_t)lane * d.cap_paths;
    ResHdr h;
    h.status = 0; h.n_frames = c->frame; h.score = h.ac = h.lm = JG_LZ; h.error = 0; h.word_off = 0; h.n_words = 0;
    if (scratch[0] == 0) h.status = -1;                        // no live hypothesis
    else if (scratch[1]) h.status = -3;                        // not traceable (see above)
    else if (scratch[2] > 0) {
        const int head = scratch[2] - 1;
        int n = 0;
        for (int p = head; p >= 0; p = paths[p].prev) ++n;
        const int off = atomicAdd(d.res_used, n);
        h.n_words = n;
        if (off + n > d.res_words_cap) { h.error = JG_ERR_WORDS; h.status = JGPU_E_CAPACITY - 10 - (JG_ERR_WORDS << 8); }
        else {
            h.status = n; h.word_off = off;
            const PathRec r0 = paths[head];
            h.score = r0.score; h.ac = r0.ac; h.lm = r0.lm;
            int k = n;
            for (int p = head; p >= 0; p = paths[p].prev) {
                --k;
                const PathRec r = paths[p];
                JgpuWord o;
                o.label = r.label; o.time = r.frame; o.score = r.score; o.ac = r.ac; o.lm = r.lm;
                d.res_words[off + k] = o;
            }
        }
    }
    d.res_hdr[slot] = h;
}
