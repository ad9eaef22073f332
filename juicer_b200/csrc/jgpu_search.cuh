// jgpu_search.cuh — the per-frame token-passing kernels (sm_100a).
//
// One frame step of every lane is the launch sequence
//   k_boundary -> k_internal -> k_seed -> { k_walk<0> [-> k_walk_huge<0>] } x n_rounds
//              -> k_walk<1> [-> k_walk_huge<1>]
// which restates WFSTDecoderLite::processFrame (src/WFSTDecoderLite.cpp:311-372) as
// data-parallel passes.  All float arithmetic on scores is plain fp32 add/sub in the
// reference's per-token order (compiled with -fmad=false; there are no multiplies), so
// every token carries bit-identical scores to the CPU decoder.
//
// Work distribution: the per-lane work lists (active instances, exit tokens, frontier,
// commit list) have very different lengths (a lane whose hub state was just expanded holds
// 20k fresh instances, its neighbour 2k), so every kernel runs a fixed grid (a multiple of
// the 148 SMs) and each CTA takes an equal, contiguous slice of the CONCATENATION of all
// lanes' lists (prefix of the per-lane counts in shared memory).
#pragma once

#include "jgpu_device.cuh"

#define JG_MAX_LANES 1024

// Per-lane views -------------------------------------------------------------------------
struct LaneView {
    LaneCtl* c;
    int2* meta_cur; int2* meta_nxt;
    float4* tok_cur; float4* tok_nxt;
    ArcDyn* ad;
    u64* skey;
    int* exit_arc; float4* exit_tok;
    Arrival* arr;
    int2* huge;
    PathRec* paths;
    int* hist;
};

__device__ __forceinline__ LaneView lane_view(const Dev& d, int lane)
{
    LaneView v;
    v.c = d.ctl + lane;
    const int flip = v.c->flip;
    const size_t cap = (size_t)d.cap, P = (size_t)(d.S - 1);
    v.meta_cur = d.inst_meta + ((size_t)lane * 2 + flip) * cap;
    v.meta_nxt = d.inst_meta + ((size_t)lane * 2 + (flip ^ 1)) * cap;
    v.tok_cur = d.tok + ((size_t)lane * 2 + flip) * P * cap;
    v.tok_nxt = d.tok + ((size_t)lane * 2 + (flip ^ 1)) * P * cap;
    v.ad = d.arcdyn + (size_t)lane * d.n_arcs;
    v.skey = d.state_key + (size_t)lane * d.n_states;
    v.exit_arc = d.exit_arc + (size_t)lane * cap;
    v.exit_tok = d.exit_tok + (size_t)lane * cap;
    v.arr = d.arr + (size_t)lane * d.cap_arr;
    v.huge = d.huge + (size_t)lane * (JG_MAX_ROUNDS + 1) * d.cap_huge;
    v.paths = d.paths + (size_t)lane * d.cap_paths;
    v.hist = d.hist + (size_t)lane * d.hist_nbins;
    return v;
}

// aggregated counter bump: the threads of the warp that are here together share one
// atomicAdd.  All of them must target the SAME counter (one lane per CTA segment).
__device__ __forceinline__ int agg_inc(int* counter)
{
    const unsigned peers = __activemask();
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane_id() == leader) base = atomicAdd(counter, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    return base + __popc(peers & ((1u << lane_id()) - 1u));
}

// ---- lane-balanced slicing ---------------------------------------------------------------
enum { JG_CNT_CUR = 0, JG_CNT_EXIT, JG_CNT_ROUND, JG_CNT_ALL };

// first arrival record of expansion round k (arrivals of earlier rounds are final by then)
__device__ __forceinline__ int arr_base(const LaneCtl* c, int round)
{
    int b = 0;
    for (int j = 0; j < round; ++j) b += c->n_arr[j];
    return b;
}

__device__ __forceinline__ int lane_count(const Dev& d, int lane, int which, int round)
{
    const LaneCtl* c = d.ctl + lane;
    const int mode = c->mode;
    if (mode == JG_MODE_IDLE) return 0;
    switch (which) {
    case JG_CNT_CUR: return mode == JG_MODE_FRAME ? c->n_cur : 0;
    case JG_CNT_EXIT: return mode == JG_MODE_FRAME ? c->n_exit : 0;
    case JG_CNT_ROUND: {
        const int b = arr_base(c, round);
        return max(0, min(c->n_arr[round], d.cap_arr - b));
    }
    default: return min(arr_base(c, d.n_rounds + 1), d.cap_arr);
    }
}

// sh_pref[0..L] = exclusive prefix of the lane counts; returns this CTA's global slice [g0, g1)
__device__ __forceinline__ void balanced_slice(const Dev& d, int which, int round, int* sh_pref, int& g0, int& g1)
{
    const int L = d.n_lanes;
    for (int l = threadIdx.x; l < L; l += blockDim.x) sh_pref[l + 1] = lane_count(d, l, which, round);
    if (threadIdx.x == 0) sh_pref[0] = 0;
    __syncthreads();
    if (threadIdx.x < 32) {                                  // warp 0: scan L values, L/32 per thread
        const int per = (L + 31) / 32;
        const int b = threadIdx.x * per;
        int sum = 0;
        for (int i = 0; i < per; ++i)
            if (b + i < L) sum += sh_pref[b + i + 1];
        int incl = sum;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)threadIdx.x >= o) incl += t;
        }
        int run = incl - sum;
        for (int i = 0; i < per; ++i)
            if (b + i < L) { run += sh_pref[b + i + 1]; sh_pref[b + i + 1] = run; }
    }
    __syncthreads();
    const int total = sh_pref[L];
    int per = (total + gridDim.x - 1) / gridDim.x;
    per = (per + 31) & ~31;
    g0 = min(blockIdx.x * per, total);
    g1 = min(g0 + per, total);
}

__device__ __forceinline__ int first_lane_of(const int* sh_pref, int L, int g)
{
    int lo = 0, hi = L;                                      // largest l with sh_pref[l] <= g
    while (lo + 1 < hi) {
        const int mid = (lo + hi) >> 1;
        if (sh_pref[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
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
    h.error = c->error; h.pad1 = h.pad2 = 0;
    if (c->error) {
        h.status = JGPU_E_CAPACITY - 10 - (c->error << 8);
    } else if (c->final_valid) {
        const float4 best = c->final_tok;
        int n = 0;
        for (int p = __float_as_int(best.w); p >= 0; p = v.paths[p].prev) ++n;
        if (n == 0) {
            h.status = -2;                                   // :273-306: no word label on the path
        } else {
            h.status = n; h.score = best.x; h.ac = best.y; h.lm = best.z;
            JgpuWord* w = d.res_words + (size_t)utt * d.max_words;
            int k = n;
            for (int p = __float_as_int(best.w); p >= 0; p = v.paths[p].prev) {
                --k;
                if (k < d.max_words) {
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
    c->b_stats[0] += c->s_frames;        c->b_stats[1] += c->s_active_models;
    c->b_stats[2] += c->s_active_emit;   c->b_stats[3] += c->s_active_end;
    c->b_stats[4] += c->s_proc_emit;     c->b_stats[5] += c->s_proc_end;
    c->b_stats[6] += c->s_gmm;           c->b_stats[7] += c->s_arcs;
    c->b_stats[8] += c->s_entry;         c->b_stats[9] += c->s_paths;
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

__global__ void __launch_bounds__(32) k_boundary(Dev d, int step, int open)
{
    const int lane = blockIdx.x;
    LaneView v = lane_view(d, lane);
    LaneCtl* c = v.c;
    const int l = lane_id();
    const int prev_mode = c->mode;
    const int4 s = open ? d.sched[(size_t)step * d.n_lanes + lane] : make_int4(-1, 0, 0, -1);
    const int mode = s.z & 3;

    // ---- (A) close the previous step ----------------------------------------------------
    if (l == 0 && prev_mode != JG_MODE_IDLE) {
        const int n_after = min(c->n_next, d.cap);
        const int n_arr_total = arr_base(c, d.n_rounds + 1);
        if (c->n_next > d.cap) c->error |= JG_ERR_ACTIVE;
        if (n_arr_total > d.cap_arr) c->error |= JG_ERR_ARRIVALS;
        if (c->n_paths > d.cap_paths) c->error |= JG_ERR_PATHS;
        const u64 key = c->best_final;
        if (key) {
            const Arrival a = v.arr[(unsigned)key];
            float4 t = a.tok;
            const float fw = __int_as_float(d.states[a.q].z);
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
            c->s_gmm += d.n_gmms;
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
        c->s_paths = c->n_paths;
    }
    __syncwarp();
    if (l == 0 && (s.z & JG_FLAG_FINISH)) finish_utterance(d, v, lane);
    __syncwarp();

    // ---- (B) open this step -------------------------------------------------------------
    if (mode != JG_MODE_IDLE && ((c->epoch + 1u) & 0x7ffu) == 0u) {      // 11-bit epoch of the state keys wraps
        for (int i = l; i < d.n_states; i += 32) v.skey[i] = 0;
    }
    __syncwarp();
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
        c->n_next = 0; c->n_exit = 0;
        for (int i = 0; i <= JG_MAX_ROUNDS; ++i) { c->n_arr[i] = 0; c->n_huge[i] = 0; }
        c->n_arr[JG_MAX_ROUNDS + 1] = 0;
        c->best_final = 0;
        c->c_active_emit = c->c_active_end = c->c_end_proc = c->c_arcs = c->c_entry = 0;
        c->mode = mode;
        if (mode != JG_MODE_IDLE) c->epoch += 1;             // invalidates every arcdyn.slot of older steps
        if (mode == JG_MODE_SEED) {                          // recognitionStart :139-228
            c->utt = s.w;
            c->frame = 0;
            c->error = 0;
            c->n_cur = 0;                                    // previous utterance's instances are dropped (:148-158)
            c->n_paths = 0;
            c->best_int = f2o(JG_LZ);
            c->best_ext = f2o(JG_LZ);
            c->norm = 0.0f; c->thr_emit = JG_LZ; c->thr_start = JG_LZ;
            c->hist_count = 0;
            c->final_valid = 0;
            c->s_active_models = c->s_active_emit = c->s_active_end = c->s_proc_emit = c->s_proc_end = 0;
            c->s_arcs = c->s_entry = c->s_paths = c->s_frames = c->s_gmm = 0;
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
            c->srow = s.y;
            if (c->frame >= d.max_frames && d.frame_stats) c->error |= JG_ERR_FRAMES;
        }
    }
    if (mode == JG_MODE_SEED && d.max_hyps > 0)
        for (int b = l; b < d.hist_nbins; b += 32) v.hist[b] = 0;
}

// =========================================================================================
// k_internal: one thread per active instance.  HMMInternalPropagation
// (src/WFSTDecoderLite.cpp:376-484) + the list walk of doHMMInternalPropagation (:899-935),
// with survivors compacted by warp ballot into the next list.
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

template <int S>
__global__ void __launch_bounds__(JG_THREADS, (S <= 5 ? 4 : 2)) k_internal(Dev d)
{
    JG_TRACE_SCOPE(JGPU_K_INTERNAL, 0);
    __shared__ int sh_pref[JG_MAX_LANES + 1];
    __shared__ float sh_best[JG_THREADS / 32];
    __shared__ int sh_cnt[3][JG_THREADS / 32];
    int g0, g1;
    balanced_slice(d, JG_CNT_CUR, 0, sh_pref, g0, g1);
    if (g0 >= g1) return;
    const size_t cap = (size_t)d.cap;
    constexpr int P = S - 1;
    const bool hist_on = d.max_hyps > 0;
    const int L = d.n_lanes;

    for (int lane = first_lane_of(sh_pref, L, g0); lane < L && sh_pref[lane] < g1; ++lane) {
        const int i0 = max(g0, sh_pref[lane]) - sh_pref[lane];
        const int i1 = min(g1, sh_pref[lane + 1]) - sh_pref[lane];
        if (i1 <= i0) continue;
        LaneView v = lane_view(d, lane);
        LaneCtl* c = v.c;
        const float norm = c->norm, thr_emit = c->thr_emit, thr_start = c->thr_start;
        const unsigned epoch = c->epoch;
        const float* __restrict__ scores = d.scores + (size_t)c->srow * d.n_gmms;
        float best = JG_LZ;
        int cnt_emit = 0, cnt_end = 0, cnt_hist = 0;

        for (int base = i0; base < i1; base += blockDim.x) {
            const int k = base + threadIdx.x;
            const bool valid = k < i1;
            bool survive = false, has_exit = false;
            int arc = 0, nst = 2, hmm = 0;
            float4 nt[S];
            float4 ex = null_tok();
#pragma unroll
            for (int i = 0; i < S; ++i) nt[i] = null_tok();
            if (valid) {
                const int2 meta = v.meta_cur[k];
                arc = meta.x;
                hmm = meta.y & ~JG_FRESH;
                const bool fresh = (meta.y & JG_FRESH) != 0;
                const int4 h0 = __ldg(reinterpret_cast<const int4*>(d.hmm_info) + hmm * 2);
                const int4 h1 = __ldg(reinterpret_cast<const int4*>(d.hmm_info) + hmm * 2 + 1);
                nst = h0.x & 0xff;
                const int cls = (h0.x & ~JG_LR_CLASS) >> 8;
                const int gm[6] = {h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                float4 old[S];
                old[0] = v.tok_cur[k];
#pragma unroll
                for (int i = 1; i < P; ++i)
                    old[i] = (!fresh && i < nst - 1) ? v.tok_cur[(size_t)i * cap + k] : null_tok();
                old[S - 1] = null_tok();
                if (old[0].x > JG_LZ && old[0].x < thr_start) old[0] = null_tok();   // :915-918
                int nlive = 0;
                const bool lr = S == 5 && (h0.x & JG_LR_CLASS);
                float lrc[8];
                const float* __restrict__ trp = d.trp + (size_t)cls * S * S;
                const int2* __restrict__ se = d.se + (size_t)cls * S;
                if (lr) {
                    const float4 c0 = __ldg(d.lr + cls * 2), c1 = __ldg(d.lr + cls * 2 + 1);
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
                            const float o = __ldg(scores + gm[j - 1]);                 // calcOutput :411
                            res.x = res.x + o;
                            res.y = res.y + o;
                            if (hist_on) {                                             // Histogram::addScore
                                int sc;
                                if (res.x < 0.0f) sc = (int)((double)res.x - 0.5);
                                else sc = (int)((double)res.x + 0.5);
                                if (sc > d.hist_max) atomicOr(&c->error, JG_ERR_HIST);
                                else if (sc >= d.hist_min) { atomicAdd(&v.hist[sc - d.hist_min], 1); ++cnt_hist; }
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
            // survivors -> next list (warp-ballot compaction); instances that die simply stop
            // being listed: their arcdyn.slot goes stale with the epoch (:924-925)
            int pos, e;
            warp_alloc2(&c->n_next, survive, &c->n_exit, has_exit, pos, e);
            if (survive && pos < d.cap) {
                v.meta_nxt[pos] = make_int2(arc, hmm);
                v.tok_nxt[pos] = null_tok();                  // entry token consumed (:426-435)
#pragma unroll
                for (int i = 1; i < P; ++i)
                    if (i < nst - 1) v.tok_nxt[(size_t)i * cap + pos] = nt[i];
                *reinterpret_cast<uint2*>(&v.ad[arc].slot) = make_uint2((unsigned)pos + 1u, epoch);
            }
            if (has_exit) {                                   // n_exit <= n_cur <= cap
                v.exit_arc[e] = arc;
                v.exit_tok[e] = ex;
            }
        }
        // per-lane block reductions
        for (int o = 16; o > 0; o >>= 1) {
            best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
            cnt_emit += __shfl_xor_sync(0xffffffffu, cnt_emit, o);
            cnt_end += __shfl_xor_sync(0xffffffffu, cnt_end, o);
            cnt_hist += __shfl_xor_sync(0xffffffffu, cnt_hist, o);
        }
        const int w = threadIdx.x >> 5;
        __syncthreads();
        if (lane_id() == 0) { sh_best[w] = best; sh_cnt[0][w] = cnt_emit; sh_cnt[1][w] = cnt_end; sh_cnt[2][w] = cnt_hist; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 1; i < JG_THREADS / 32; ++i) {
                best = fmaxf(best, sh_best[i]);
                cnt_emit += sh_cnt[0][i]; cnt_end += sh_cnt[1][i]; cnt_hist += sh_cnt[2][i];
            }
            if (best > JG_LZ) atomicMax(&c->best_int, f2o(best));
            if (cnt_emit) atomicAdd(&c->c_active_emit, cnt_emit);
            if (cnt_end) atomicAdd(&c->c_active_end, cnt_end);
            if (cnt_hist) atomicAdd(&c->hist_count, cnt_hist);
        }
    }
}

// =========================================================================================
// External propagation (doHMMExternalPropagation :937-982 + propagateToken :491-605) as
// level-synchronous rounds over WFST states.
//   arrival  = a token reaching state q through arc `via` (exit of an instance, epsilon arc,
//              or tee pass-through).  Arrivals are max-reduced per state by a fire-and-forget
//              64-bit atomicMax on state_key (score bits | record index); the records of one
//              round are stored back to back, so round k simply walks its slice of records and
//              expands those that still own their state — once per round instead of the
//              reference's re-expansion per token.
//   pass 0   = expansion, per out-arc of q: epsilon arc -> arrival for the next round
//              (:533-540); model arc -> entry-token candidate, atomicMax on arcdyn.key
//              (:542-582); tee model -> additional pass-through arrival (:584-600).
//   pass 1   = commit: every record that owns its state walks its arc row again; the candidate
//              that owns arcdyn.key writes the entry token into the next list, attaching a new
//              instance when the arc had none (attachNetInst :751-774), and clears the key.
// No atomic in pass 0 returns a value the thread has to wait for, except the per-round record
// counter (one aggregated atomicAdd per warp).
// =========================================================================================
// state key = epoch (11 bits) | orderable score (32 bits) | arrival record (21 bits): keys of older
// steps always lose the atomicMax and never compare equal, so state_key needs no per-frame cleaning
// (k_boundary wipes a lane's table when its 11-bit epoch wraps, once every 2048 steps).
#define JG_R_BITS 21
__device__ __forceinline__ u64 state_key_of(unsigned epoch, float score, unsigned r)
{
    return ((u64)(epoch & 0x7ffu) << 53) | ((u64)f2o(score) << JG_R_BITS) | (u64)r;
}

__device__ __forceinline__ void arrive(const Dev& d, const LaneView& v, int q, int via, float4 tok, int out_round,
                                       int out_base)
{
    const int r = out_base + agg_inc(&v.c->n_arr[out_round]);
    if (r >= d.cap_arr) return;                              // flagged by k_boundary
    Arrival a;
    a.tok = tok; a.via = via; a.q = q; a.pad[0] = a.pad[1] = 0;
    v.arr[r] = a;
    atomicMax(&v.skey[q], state_key_of(v.c->epoch, tok.x, (unsigned)r));
}

__global__ void __launch_bounds__(JG_THREADS, 6) k_seed(Dev d)
{
    JG_TRACE_SCOPE(JGPU_K_SEED, 0);
    __shared__ int sh_pref[JG_MAX_LANES + 1];
    // utterance seeds: propagateToken(&zeroToken, NULL) (:221-226) = an arrival at the initial state
    if (blockIdx.x == 0 && threadIdx.x < 32)
        for (int lane = threadIdx.x; lane < d.n_lanes; lane += 32)
            if (d.ctl[lane].mode == JG_MODE_SEED) {
                LaneView v = lane_view(d, lane);
                Arrival a;
                a.tok = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1));
                a.via = -1; a.q = d.init_state; a.pad[0] = a.pad[1] = 0;
                v.arr[0] = a;                                 // seed lanes have no exit tokens: record 0 is free
                v.c->n_arr[0] = 1;
                v.skey[d.init_state] = state_key_of(v.c->epoch, 0.0f, 0u);
            }
    int g0, g1;
    balanced_slice(d, JG_CNT_EXIT, 0, sh_pref, g0, g1);
    if (g0 >= g1) return;
    const int L = d.n_lanes;
    for (int lane = first_lane_of(sh_pref, L, g0); lane < L && sh_pref[lane] < g1; ++lane) {
        const int i0 = max(g0, sh_pref[lane]) - sh_pref[lane];
        const int i1 = min(g1, sh_pref[lane + 1]) - sh_pref[lane];
        if (i1 <= i0) continue;
        LaneView v = lane_view(d, lane);
        LaneCtl* c = v.c;
        const float be = o2f(c->best_int);
        const float thr_end = (d.end_beam > 0.0f ? (be - d.end_beam) : JG_LZ);     // :349
        const float thr_word = (d.word_beam > 0.0f ? (be - d.word_beam) : JG_LZ);  // :350
        int proc = 0;
        for (int e = i0 + threadIdx.x; e < i1; e += blockDim.x) {
            const int arc = v.exit_arc[e];
            const float4 t = v.exit_tok[e];
            const int4 a = __ldg(&d.arcs[arc]);
            const float thr = a.w == 0 ? thr_end : thr_word;                        // :952-962
            if (t.x > thr) {
                ++proc;
                arrive(d, v, a.x, arc, t, 0, 0);
            }
        }
        __syncwarp();
        for (int o = 16; o > 0; o >>= 1) proc += __shfl_xor_sync(0xffffffffu, proc, o);
        if (lane_id() == 0 && proc) atomicAdd(&c->c_end_proc, proc);
    }
}

struct WalkCtx {
    float thr_end, thr_word;
    int out_round, out_base;
    unsigned epoch;
};

template <int PASS>
__device__ __forceinline__ void process_arc(const Dev& d, const LaneView& v, const WalkCtx& x, const float4 tok,
                                            unsigned r, int b, float& best, int& n_entry)
{
    const int4 a = __ldg(&d.arcs[b]);
    const float w = __int_as_float(a.y);
    const float s = tok.x + w;
    if (a.z == 0) {                                           // epsilon input: :533-540
        if (PASS == 0 && s > x.thr_end)
            arrive(d, v, a.x, b, make_float4(s, tok.y, tok.z + w, tok.w), x.out_round, x.out_base);
        return;
    }
    // model arc: :542-601
    if (s > JG_LZ) {
        const u64 key = ((u64)f2o(s) << 32) | r;
        if (PASS == 0) {
            atomicMax(&v.ad[b].key, key);                     // recombination: best entry candidate of the arc
        } else {
            const uint4 dyn = *reinterpret_cast<const uint4*>(&v.ad[b]);
            if (dyn.x == r && dyn.y == (unsigned)(key >> 32)) {             // this candidate won
                const float4 t = make_float4(s, tok.y, tok.z + w, tok.w);   // :568-570
                if (s > best) best = s;
                ++n_entry;
                if (dyn.w == x.epoch && dyn.z != 0) {         // the instance survived the internal phase
                    v.tok_nxt[dyn.z - 1] = t;                 // plane 0 = entry token
                    v.ad[b].key = 0;
                } else {
                    const int pos = agg_inc(&v.c->n_next);
                    if (pos < d.cap) {
                        v.meta_nxt[pos] = make_int2(b, (a.z - 1) | JG_FRESH);
                        v.tok_nxt[pos] = t;
                        *reinterpret_cast<uint4*>(&v.ad[b]) = make_uint4(0u, 0u, (unsigned)pos + 1u, x.epoch);
                    } else {
                        v.ad[b].key = 0;                      // overflow is flagged by k_boundary
                    }
                }
            }
        }
    }
    if (PASS == 0 && d.arc_tee) {
        const float tee = __ldg(d.arc_tee + b);
        if (tee > JG_LZ) {                                    // :584-600
            const float s2 = s + tee;
            const float thr = a.w != 0 ? x.thr_word : x.thr_end;
            if (s2 > thr)
                arrive(d, v, a.x, b, make_float4(s2, tok.y + tee, tok.z + w, tok.w), x.out_round, x.out_base);
        }
    }
}

__device__ __forceinline__ WalkCtx walk_ctx(const Dev& d, const LaneCtl* c, int round)
{
    WalkCtx x;
    x.thr_end = x.thr_word = JG_LZ;
    if (c->mode == JG_MODE_FRAME) {
        const float be = o2f(c->best_int);
        x.thr_end = (d.end_beam > 0.0f ? (be - d.end_beam) : JG_LZ);
        x.thr_word = (d.word_beam > 0.0f ? (be - d.word_beam) : JG_LZ);
    }
    x.out_round = round + 1;
    x.out_base = arr_base(c, round + 1);
    x.epoch = c->epoch;
    return x;
}

// PASS 0: expansion round `round`.  PASS 1: commit over the records of all rounds.
// Per chunk of 256 records: (A) one thread per record decides whether the record still owns its
// state and does the per-record work (word-boundary record, final-state candidate); (B) the arc
// rows of all 256 records are walked as ONE flattened list — thread t takes arcs t, t+256, ... and
// finds the owning record by binary search in the shared prefix of out-degrees — so every lane has
// an independent arc in flight regardless of how the out-degrees are distributed.
template <int PASS>
__global__ void __launch_bounds__(JG_THREADS, 6) k_walk(Dev d, int round)
{
    JG_TRACE_SCOPE(PASS ? JGPU_K_COMMIT : JGPU_K_EXPAND, round);
    __shared__ int sh_pref[JG_MAX_LANES + 1];
    __shared__ int s_off[JG_THREADS + 1];
    __shared__ int s_first[JG_THREADS];
    __shared__ unsigned s_r[JG_THREADS];
    __shared__ float4 s_tok[JG_THREADS];
    __shared__ int s_wsum[JG_THREADS / 32];
    int g0, g1;
    balanced_slice(d, PASS == 0 ? JG_CNT_ROUND : JG_CNT_ALL, round, sh_pref, g0, g1);
    if (g0 >= g1) return;
    const int L = d.n_lanes;
    const int tid = threadIdx.x, wid = tid >> 5;
    for (int lane = first_lane_of(sh_pref, L, g0); lane < L && sh_pref[lane] < g1; ++lane) {
        const int i0 = max(g0, sh_pref[lane]) - sh_pref[lane];
        const int i1 = min(g1, sh_pref[lane + 1]) - sh_pref[lane];
        if (i1 <= i0) continue;
        LaneView v = lane_view(d, lane);
        LaneCtl* c = v.c;
        const WalkCtx x = walk_ctx(d, c, round);
        const int rec0 = PASS == 0 ? arr_base(c, round) : 0;
        const int frame = c->frame;
        int arcs_done = 0, n_entry = 0;
        float best = JG_LZ;
        for (int base = i0; base < i1; base += blockDim.x) {
            // ---- (A) one thread per record ----
            const int e = base + tid;
            bool valid = e < i1;
            int q = 0, first = 0, deg = 0;
            const unsigned r = (unsigned)(rec0 + e);
            float4 tok = null_tok();
            if (valid) {
                const Arrival a = v.arr[r];
                q = a.q;
                tok = a.tok;
                valid = a.via != -2 && v.skey[q] == state_key_of(x.epoch, tok.x, r);   // still the best arrival of q?
                if (valid) {
                    const int4 st = __ldg(&d.states[q]);
                    first = st.x; deg = st.y;
                    if (PASS == 0 && a.via >= 0) {
                        const int olab = __ldg(&d.arcs[a.via]).w;
                        if (olab != 0) {                      // word boundary record: :497-509
                            const int p = agg_inc(&c->n_paths);
                            if (p < d.cap_paths) {
                                PathRec pr;
                                pr.prev = __float_as_int(tok.w); pr.frame = frame; pr.label = olab;
                                pr.score = tok.x; pr.ac = tok.y; pr.lm = tok.z; pr.pad0 = pr.pad1 = 0;
                                v.paths[p] = pr;
                                tok.w = __int_as_float(p);
                                v.arr[r].tok.w = tok.w;
                            } else {
                                valid = false;                // flagged by k_boundary
                                v.arr[r].via = -2;
                            }
                        }
                        const float fw = __int_as_float(st.z);
                        if (valid && fw > JG_LZ)              // :513-520
                            atomicMax(&c->best_final, ((u64)f2o(tok.x + fw) << 32) | r);
                    }
                    if (!valid) deg = 0;
                    if (PASS == 0) arcs_done += deg;
                    if (deg >= d.huge_deg) {                  // hub-like state: left to k_walk_huge
                        if (PASS == 0) {
                            const int h = atomicAdd(&c->n_huge[round], 1);
                            if (h < d.cap_huge) v.huge[(size_t)round * d.cap_huge + h] = make_int2(q, (int)r);
                            else atomicOr(&c->error, JG_ERR_HUGE);
                        }
                        deg = 0;
                    }
                }
            }
            // ---- block-wide exclusive prefix of the out-degrees ----
            int incl = deg;
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane_id() >= o) incl += t;
            }
            if (lane_id() == 31) s_wsum[wid] = incl;
            s_first[tid] = first; s_r[tid] = r; s_tok[tid] = tok;
            __syncthreads();
            int woff = 0;
            for (int w = 0; w < wid; ++w) woff += s_wsum[w];
            s_off[tid] = woff + incl - deg;
            if (tid == blockDim.x - 1) s_off[blockDim.x] = woff + incl;
            __syncthreads();
            const int total = s_off[blockDim.x];
            // ---- (B) flattened arc rows ----
            for (int j = tid; j < total; j += blockDim.x) {
                int lo = 0, hi = blockDim.x;                  // largest src with s_off[src] <= j
                while (lo + 1 < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (s_off[mid] <= j) lo = mid; else hi = mid;
                }
                process_arc<PASS>(d, v, x, s_tok[lo], s_r[lo], s_first[lo] + (j - s_off[lo]), best, n_entry);
            }
            __syncthreads();
        }
        __syncwarp();
        for (int o = 16; o > 0; o >>= 1) {
            arcs_done += __shfl_xor_sync(0xffffffffu, arcs_done, o);
            n_entry += __shfl_xor_sync(0xffffffffu, n_entry, o);
            best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
        }
        if (lane_id() == 0) {
            if (arcs_done) atomicAdd(&c->c_arcs, arcs_done);
            if (n_entry) atomicAdd(&c->c_entry, n_entry);
            if (best > JG_LZ) atomicMax(&c->best_ext, f2o(best));               // :572-573
        }
    }
}

// hub-like states: a group of CTAs per lane strides over the arc row.
// PASS 0: the states met in round `round`; PASS 1: those of every round.
template <int PASS>
__global__ void __launch_bounds__(JG_THREADS) k_walk_huge(Dev d, int round)
{
    JG_TRACE_SCOPE(JGPU_K_EXPAND_HUGE, round + 100 * PASS);
    const int lane = blockIdx.y;
    LaneCtl* c = d.ctl + lane;
    if (c->mode == JG_MODE_IDLE) return;
    const int r_lo = PASS == 0 ? round : 0, r_hi = PASS == 0 ? round + 1 : d.n_rounds;
    bool any = false;
    for (int k = r_lo; k < r_hi; ++k) any |= c->n_huge[k] > 0;
    if (!any) return;
    LaneView v = lane_view(d, lane);
    int n_entry = 0;
    float best = JG_LZ;
    for (int k = r_lo; k < r_hi; ++k) {
        const int n = min(c->n_huge[k], d.cap_huge);
        if (n == 0) continue;
        const WalkCtx x = walk_ctx(d, c, k);
        const int2* list = v.huge + (size_t)k * d.cap_huge;
        for (int h = 0; h < n; ++h) {
            const int2 qr = list[h];
            const float4 tok = v.arr[qr.y].tok;
            if (PASS == 1 && v.skey[qr.x] != state_key_of(x.epoch, tok.x, (unsigned)qr.y)) continue;   // re-expanded later by a better token
            const int4 st = __ldg(&d.states[qr.x]);
            for (int b = st.x + blockIdx.x * blockDim.x + threadIdx.x; b < st.x + st.y; b += gridDim.x * blockDim.x)
                process_arc<PASS>(d, v, x, tok, (unsigned)qr.y, b, best, n_entry);
        }
    }
    if (PASS == 1) {
        __syncwarp();
        for (int o = 16; o > 0; o >>= 1) {
            n_entry += __shfl_xor_sync(0xffffffffu, n_entry, o);
            best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
        }
        if (lane_id() == 0) {
            if (n_entry) atomicAdd(&c->c_entry, n_entry);
            if (best > JG_LZ) atomicMax(&c->best_ext, f2o(best));
        }
    }
}
