// jgpu_search.cuh — the per-frame token-passing kernels (sm_100a).
//
// One frame step of every lane is the launch sequence
//   k_boundary -> k_internal -> k_seed -> { k_expand [-> k_expand_huge] } x n_rounds -> k_commit
// which restates WFSTDecoderLite::processFrame (src/WFSTDecoderLite.cpp:311-372) as
// data-parallel passes.  All float arithmetic on scores is plain fp32 add/sub in the
// reference's per-token order (compiled with -fmad=false; there are no multiplies), so
// every token carries bit-identical scores to the CPU decoder.
#pragma once

#include "jgpu_device.cuh"

// Per-lane views -------------------------------------------------------------------------
struct LaneView {
    LaneCtl* c;
    int* arc_cur;  int* arc_nxt;
    float4* tok_cur; float4* tok_nxt;
    unsigned* a2s;
    u64* ekey; u64* skey;
    int* exit_arc; float4* exit_tok;
    float4* arr_tok; int* arr_via;
    int2* front; int2* huge;
    int* commit_arc; int* touched;
    PathRec* paths;
    int* hist;
};

__device__ __forceinline__ LaneView lane_view(const Dev& d, int lane)
{
    LaneView v;
    v.c = d.ctl + lane;
    const int flip = v.c->flip;
    const size_t cap = (size_t)d.cap, P = (size_t)(d.S - 1);
    v.arc_cur = d.inst_arc + ((size_t)lane * 2 + flip) * cap;
    v.arc_nxt = d.inst_arc + ((size_t)lane * 2 + (flip ^ 1)) * cap;
    v.tok_cur = d.tok + ((size_t)lane * 2 + flip) * P * cap;
    v.tok_nxt = d.tok + ((size_t)lane * 2 + (flip ^ 1)) * P * cap;
    v.a2s = d.arc2slot + (size_t)lane * d.n_arcs;
    v.ekey = d.entry_key + (size_t)lane * d.n_arcs;
    v.skey = d.state_key + (size_t)lane * d.n_states;
    v.exit_arc = d.exit_arc + (size_t)lane * cap;
    v.exit_tok = d.exit_tok + (size_t)lane * cap;
    v.arr_tok = d.arr_tok + (size_t)lane * d.cap_arr;
    v.arr_via = d.arr_via + (size_t)lane * d.cap_arr;
    v.front = d.front + (size_t)lane * 2 * d.cap_arr;
    v.huge = d.huge + (size_t)lane * 2 * d.cap_huge;
    v.commit_arc = d.commit_arc + (size_t)lane * cap;
    v.touched = d.touched + (size_t)lane * d.cap_arr;
    v.paths = d.paths + (size_t)lane * d.cap_paths;
    v.hist = d.hist + (size_t)lane * d.hist_nbins;
    return v;
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
        if (c->n_next > d.cap) c->error |= JG_ERR_ACTIVE;
        if (c->n_arr > d.cap_arr) c->error |= JG_ERR_ARRIVALS;
        if (c->n_paths > d.cap_paths) c->error |= JG_ERR_PATHS;
        const u64 key = c->best_final;
        if (key) {
            const unsigned r = (unsigned)key;
            float4 t = v.arr_tok[r];
            const int via = v.arr_via[r];
            const float fw = d.state_final[d.arcs[via].x];
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
            c->s_entry += min(c->n_commit, d.cap);
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
        c->n_next = 0; c->n_exit = 0; c->n_arr = 0; c->n_commit = 0; c->n_touched = 0;
        for (int i = 0; i <= JG_MAX_ROUNDS; ++i) { c->n_front[i] = 0; c->n_huge[i] = 0; }
        c->best_final = 0;
        c->c_active_emit = c->c_active_end = c->c_end_proc = c->c_arcs = 0;
        c->mode = mode;
        if (mode == JG_MODE_SEED) {                          // recognitionStart :139-228
            c->utt = s.w;
            c->frame = 0;
            c->dirty = c->error;
            c->error = 0;
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
__global__ void __launch_bounds__(JG_THREADS) k_internal(Dev d)
{
    const int lane = blockIdx.y;
    LaneView v = lane_view(d, lane);
    LaneCtl* c = v.c;
    const int mode = c->mode;
    if (mode == JG_MODE_IDLE) return;
    const int n = c->n_cur;
    const size_t cap = (size_t)d.cap;
    constexpr int P = S - 1;

    if (mode == JG_MODE_SEED) {
        // new utterance on this lane: drop the previous utterance's instances
        // (recognitionStart :148-158); after a failed utterance wipe the dense tables.
        const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
        if (c->dirty) {
            for (int i = gtid; i < d.n_arcs; i += gsz) { v.a2s[i] = 0; v.ekey[i] = 0; }
            for (int i = gtid; i < d.n_states; i += gsz) v.skey[i] = 0;
        } else {
            for (int k = gtid; k < n; k += gsz) v.a2s[v.arc_cur[k]] = 0;
        }
        return;
    }

    const float norm = c->norm, thr_emit = c->thr_emit, thr_start = c->thr_start;
    const float* __restrict__ scores = d.scores + (size_t)c->srow * d.n_gmms;
    const bool hist_on = d.max_hyps > 0;
    float best = JG_LZ;
    int cnt_emit = 0, cnt_end = 0, cnt_hist = 0;

    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const int k = base + threadIdx.x;
        const bool valid = k < n;
        bool survive = false, has_exit = false;
        int arc = 0, nst = 2;
        float4 nt[S];
        float4 ex = null_tok();
#pragma unroll
        for (int i = 0; i < S; ++i) nt[i] = null_tok();
        if (valid) {
            arc = v.arc_cur[k];
            const int4 a = __ldg(&d.arcs[arc]);
            const int hmm = a.z - 1;
            const int4 i0 = __ldg(reinterpret_cast<const int4*>(d.hmm_info) + hmm * 2);
            const int4 i1 = __ldg(reinterpret_cast<const int4*>(d.hmm_info) + hmm * 2 + 1);
            nst = i0.x & 0xff;
            const int cls = i0.x >> 8;
            const int gm[6] = {i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
            const float* __restrict__ trp = d.trp + (size_t)cls * S * S;
            const int2* __restrict__ se = d.se + (size_t)cls * S;
            float4 old[S];
#pragma unroll
            for (int i = 0; i < P; ++i) old[i] = (i < nst - 1) ? v.tok_cur[(size_t)i * cap + k] : null_tok();
            old[S - 1] = null_tok();
            if (old[0].x > JG_LZ && old[0].x < thr_start) old[0] = null_tok();   // :915-918
            int nlive = 0;
#pragma unroll
            for (int j = 1; j < S - 1; ++j) {
                if (j < nst - 1) {
                    float4 res = viterbi_into<S>(old, trp, __ldg(se + j), j, nst);
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
                float4 res = viterbi_into<S>(nt, trp, __ldg(se + (nst - 1)), nst - 1, nst);
                if (res.x > JG_LZ) { ex = res; has_exit = true; ++cnt_end; }
            }
        }
        // survivors -> next list (warp-ballot compaction)
        const int pos = warp_alloc(&c->n_next, survive);
        if (survive) {
            if (pos < d.cap) {
                v.arc_nxt[pos] = arc;
                v.tok_nxt[pos] = null_tok();                  // entry token consumed (:426-435)
#pragma unroll
                for (int i = 1; i < P; ++i)
                    if (i < nst - 1) v.tok_nxt[(size_t)i * cap + pos] = nt[i];
                v.a2s[arc] = (unsigned)pos + 1u;
            } else {
                v.a2s[arc] = 0;
            }
        } else if (valid) {
            v.a2s[arc] = 0;                                   // instance deactivated (:924-925)
        }
        const int e = warp_alloc(&c->n_exit, has_exit);
        if (has_exit) {                                       // n_exit <= n_cur <= cap
            v.exit_arc[e] = arc;
            v.exit_tok[e] = ex;
        }
    }
    // block reductions
    __shared__ float sh_best[JG_THREADS / 32];
    __shared__ int sh_cnt[3][JG_THREADS / 32];
    for (int o = 16; o > 0; o >>= 1) {
        best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
        cnt_emit += __shfl_xor_sync(0xffffffffu, cnt_emit, o);
        cnt_end += __shfl_xor_sync(0xffffffffu, cnt_end, o);
        cnt_hist += __shfl_xor_sync(0xffffffffu, cnt_hist, o);
    }
    const int w = threadIdx.x >> 5;
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

// =========================================================================================
// External propagation (doHMMExternalPropagation :937-982 + propagateToken :491-605) as
// level-synchronous rounds over WFST states.
//   arrival  = a token reaching state q through arc `via` (exit of an instance, epsilon arc,
//              or tee pass-through).  Arrivals are max-reduced per state (64-bit atomicMax on
//              state_key: score bits | record index); only the winner of a state is expanded,
//              once per round, instead of the reference's re-expansion per token.
//   expansion= per out-arc of q: epsilon arc -> arrival for the next round (:533-540);
//              model arc -> entry-token candidate, atomicMax on entry_key (:542-582);
//              tee model -> additional pass-through arrival (:584-600).
// =========================================================================================
__device__ __forceinline__ void arrive(const Dev& d, const LaneView& v, int q, int via, float4 tok, int out_round)
{
    const int r = atomicAdd(&v.c->n_arr, 1);
    if (r >= d.cap_arr) return;                              // flagged by k_boundary
    v.arr_tok[r] = tok;
    v.arr_via[r] = via;
    const u64 key = ((u64)f2o(tok.x) << 32) | (unsigned)r;
    const u64 old = atomicMax(&v.skey[q], key);
    if (old < key) {
        const int f = atomicAdd(&v.c->n_front[out_round], 1);   // <= n_arr <= cap_arr
        v.front[(size_t)(out_round & 1) * d.cap_arr + f] = make_int2(q, r);
    }
    if (old == 0) {
        const int t = atomicAdd(&v.c->n_touched, 1);
        v.touched[t] = q;
    }
}

__global__ void __launch_bounds__(JG_THREADS) k_seed(Dev d)
{
    const int lane = blockIdx.y;
    LaneView v = lane_view(d, lane);
    LaneCtl* c = v.c;
    const int mode = c->mode;
    if (mode == JG_MODE_IDLE) return;
    if (mode == JG_MODE_SEED) {
        // propagateToken(&zeroToken, NULL) (:221-226): an arrival at the initial state
        if (blockIdx.x == 0 && threadIdx.x == 0)
            arrive(d, v, d.init_state, -1, make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1)), 0);
        return;
    }
    const float be = o2f(c->best_int);
    const float thr_end = (d.end_beam > 0.0f ? (be - d.end_beam) : JG_LZ);     // :349
    const float thr_word = (d.word_beam > 0.0f ? (be - d.word_beam) : JG_LZ);  // :350
    const int n = c->n_exit;
    int proc = 0;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int arc = v.exit_arc[e];
        const float4 t = v.exit_tok[e];
        const int4 a = __ldg(&d.arcs[arc]);
        const float thr = a.w == 0 ? thr_end : thr_word;                        // :952-962
        if (t.x > thr) {
            ++proc;
            arrive(d, v, a.x, arc, t, 0);
        }
    }
    for (int o = 16; o > 0; o >>= 1) proc += __shfl_xor_sync(0xffffffffu, proc, o);
    if (lane_id() == 0 && proc) atomicAdd(&c->c_end_proc, proc);
}

__device__ __forceinline__ void process_arc(const Dev& d, const LaneView& v, const float4 tok, unsigned r, int b,
                                            float thr_end, float thr_word, int out_round)
{
    const int4 a = __ldg(&d.arcs[b]);
    const float w = __int_as_float(a.y);
    if (a.z == 0) {                                           // epsilon input: :533-540
        const float s = tok.x + w;
        if (s > thr_end) arrive(d, v, a.x, b, make_float4(s, tok.y, tok.z + w, tok.w), out_round);
    } else {                                                  // model arc: :542-601
        const float s = tok.x + w;
        if (s > JG_LZ) {
            const u64 key = ((u64)f2o(s) << 32) | r;
            const u64 old = atomicMax(&v.ekey[b], key);
            if (old == 0) {
                const int i = atomicAdd(&v.c->n_commit, 1);
                if (i < d.cap) v.commit_arc[i] = b;
                else { v.ekey[b] = 0; atomicOr(&v.c->error, JG_ERR_ACTIVE); }
            }
        }
        if (d.arc_tee) {
            const float tee = __ldg(d.arc_tee + b);
            if (tee > JG_LZ) {                                // :584-600
                const float s2 = s + tee;
                const float thr = a.w != 0 ? thr_word : thr_end;
                if (s2 > thr) arrive(d, v, a.x, b, make_float4(s2, tok.y + tee, tok.z + w, tok.w), out_round);
            }
        }
    }
}

__global__ void __launch_bounds__(JG_THREADS) k_expand(Dev d, int round)
{
    const int lane = blockIdx.y;
    LaneView v = lane_view(d, lane);
    LaneCtl* c = v.c;
    const int mode = c->mode;
    if (mode == JG_MODE_IDLE) return;
    const int n = min(c->n_front[round], d.cap_arr);
    if (n == 0) return;
    float thr_end = JG_LZ, thr_word = JG_LZ;
    if (mode == JG_MODE_FRAME) {
        const float be = o2f(c->best_int);
        thr_end = (d.end_beam > 0.0f ? (be - d.end_beam) : JG_LZ);
        thr_word = (d.word_beam > 0.0f ? (be - d.word_beam) : JG_LZ);
    }
    const int2* list = v.front + (size_t)(round & 1) * d.cap_arr;
    const int frame = c->frame;
    int arcs_done = 0;
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const int e = base + threadIdx.x;
        bool valid = e < n;
        int q = 0, first = 0, deg = 0;
        unsigned r = 0;
        float4 tok = null_tok();
        if (valid) {
            const int2 qr = list[e];
            q = qr.x; r = (unsigned)qr.y;
            valid = ((unsigned)v.skey[q] == r);              // still the best arrival of q?
        }
        if (valid) {
            tok = v.arr_tok[r];
            const int via = v.arr_via[r];
            if (via >= 0) {
                const int olab = __ldg(&d.arcs[via]).w;
                if (olab != 0) {                              // word boundary record: :497-509
                    const int p = atomicAdd(&c->n_paths, 1);
                    if (p < d.cap_paths) {
                        PathRec pr;
                        pr.prev = __float_as_int(tok.w); pr.frame = frame; pr.label = olab;
                        pr.score = tok.x; pr.ac = tok.y; pr.lm = tok.z; pr.pad0 = pr.pad1 = 0;
                        v.paths[p] = pr;
                        tok.w = __int_as_float(p);
                        v.arr_tok[r].w = tok.w;
                    } else {
                        valid = false;                        // flagged by k_boundary
                    }
                }
                const float fw = __ldg(d.state_final + q);
                if (valid && fw > JG_LZ) {                    // :513-520
                    const float cand = tok.x + fw;
                    atomicMax(&c->best_final, ((u64)f2o(cand) << 32) | r);
                }
            }
        }
        if (valid) {
            const int2 st = __ldg(&d.states[q]);
            first = st.x; deg = st.y;
            arcs_done += deg;
        }
        const bool small = valid && deg <= d.small_deg;
        const bool is_huge = valid && deg >= d.huge_deg;
        if (small) {
            for (int b = first; b < first + deg; ++b) process_arc(d, v, tok, r, b, thr_end, thr_word, round + 1);
        } else if (is_huge) {
            const int h = atomicAdd(&c->n_huge[round], 1);
            if (h < d.cap_huge) v.huge[(size_t)(round & 1) * d.cap_huge + h] = make_int2(q, (int)r);
            else atomicOr(&c->error, JG_ERR_HUGE);
        }
        // medium out-degree: the warp walks the arc row together
        unsigned mm = __ballot_sync(0xffffffffu, valid && !small && !is_huge);
        while (mm) {
            const int src = __ffs(mm) - 1;
            mm &= mm - 1;
            const int f_s = __shfl_sync(0xffffffffu, first, src);
            const int n_s = __shfl_sync(0xffffffffu, deg, src);
            const unsigned r_s = __shfl_sync(0xffffffffu, r, src);
            float4 t_s;
            t_s.x = __shfl_sync(0xffffffffu, tok.x, src);
            t_s.y = __shfl_sync(0xffffffffu, tok.y, src);
            t_s.z = __shfl_sync(0xffffffffu, tok.z, src);
            t_s.w = __shfl_sync(0xffffffffu, tok.w, src);
            for (int b = f_s + lane_id(); b < f_s + n_s; b += 32)
                process_arc(d, v, t_s, r_s, b, thr_end, thr_word, round + 1);
        }
    }
    for (int o = 16; o > 0; o >>= 1) arcs_done += __shfl_xor_sync(0xffffffffu, arcs_done, o);
    if (lane_id() == 0 && arcs_done) atomicAdd(&c->c_arcs, arcs_done);
}

// hub-like states: every block of the lane strides over the arc row
__global__ void __launch_bounds__(JG_THREADS) k_expand_huge(Dev d, int round)
{
    const int lane = blockIdx.y;
    LaneView v = lane_view(d, lane);
    LaneCtl* c = v.c;
    const int mode = c->mode;
    if (mode == JG_MODE_IDLE) return;
    const int n = min(c->n_huge[round], d.cap_huge);
    if (n == 0) return;
    float thr_end = JG_LZ, thr_word = JG_LZ;
    if (mode == JG_MODE_FRAME) {
        const float be = o2f(c->best_int);
        thr_end = (d.end_beam > 0.0f ? (be - d.end_beam) : JG_LZ);
        thr_word = (d.word_beam > 0.0f ? (be - d.word_beam) : JG_LZ);
    }
    const int2* list = v.huge + (size_t)(round & 1) * d.cap_huge;
    for (int h = 0; h < n; ++h) {
        const int2 qr = list[h];
        const float4 tok = v.arr_tok[qr.y];
        const int2 st = __ldg(&d.states[qr.x]);
        for (int b = st.x + blockIdx.x * blockDim.x + threadIdx.x; b < st.x + st.y; b += gridDim.x * blockDim.x)
            process_arc(d, v, tok, (unsigned)qr.y, b, thr_end, thr_word, round + 1);
    }
}

// =========================================================================================
// k_commit: winners of the entry-token recombination write their token into the next
// list, attaching a new instance when the arc had none (attachNetInst :751-774); the dense
// tables are cleaned for the next frame.
// =========================================================================================
__global__ void __launch_bounds__(JG_THREADS) k_commit(Dev d)
{
    const int lane = blockIdx.y;
    LaneView v = lane_view(d, lane);
    LaneCtl* c = v.c;
    if (c->mode == JG_MODE_IDLE) return;
    const size_t cap = (size_t)d.cap;
    const int n = min(c->n_commit, d.cap);
    float best = JG_LZ;
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const int i = base + threadIdx.x;
        const bool valid = i < n;
        int b = 0, nst = 2;
        unsigned slot = 0;
        float4 t = null_tok();
        if (valid) {
            b = v.commit_arc[i];
            const u64 key = v.ekey[b];
            v.ekey[b] = 0;
            const unsigned r = (unsigned)key;
            const float4 src = v.arr_tok[r];
            const int4 a = __ldg(&d.arcs[b]);
            const float w = __int_as_float(a.y);
            t = make_float4(src.x + w, src.y, src.z + w, src.w);               // :568-570
            if (t.x > best) best = t.x;
            slot = v.a2s[b];
            nst = __ldg(d.hmm_info + (size_t)(a.z - 1) * 8) & 0xff;
        }
        const int pos = warp_alloc(&c->n_next, valid && slot == 0);
        if (valid) {
            if (slot) {
                v.tok_nxt[slot - 1] = t;                      // plane 0 = entry token
            } else if (pos < d.cap) {
                v.arc_nxt[pos] = b;
                v.tok_nxt[pos] = t;
                for (int p = 1; p < nst - 1; ++p) v.tok_nxt[(size_t)p * cap + pos] = null_tok();
                v.a2s[b] = (unsigned)pos + 1u;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane_id() == 0 && best > JG_LZ) atomicMax(&c->best_ext, f2o(best));   // :572-573
    const int nt = min(c->n_touched, d.cap_arr);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nt; i += gridDim.x * blockDim.x)
        v.skey[v.touched[i]] = 0;
}
