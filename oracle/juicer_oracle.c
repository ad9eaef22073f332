/* TEST INFRASTRUCTURE ONLY — see juicer_oracle.h.  Plain C99, scalar, single thread.
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference).  Compile with -O2 -ffp-contract=off (no FMA): float expressions are
 * evaluated in float, the two mixed-precision spots of the reference are written out in
 * double exactly where the reference promotes. */
#define _POSIX_C_SOURCE 200809L
#include "juicer_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define LZ (-FLT_MAX) /* LOG_ZERO (Torch3 log_add.h) */
#define SMAX 8

typedef struct {
    float score; /* normalised each frame: src/WFSTDecoderLite.h:57-62 */
    float ac;
    float lm;
    int path; /* index into paths[], -1 = NULL */
} Tok;

typedef struct {
    int prev, frame, label;
    float score, ac, lm;
} PathRec; /* src/WFSTDecoderLite.h:39-55 without the GC links */

typedef struct {
    int next; /* active list link, -1 = end */
    int hmm, nst, nactive, arc;
    float tee;
    Tok st[SMAX];
} Inst; /* src/WFSTDecoderLite.h:64-75 */

struct jor_handle {
    /* copied tables */
    JgpuNet net;
    JgpuHmm hmm;
    JgpuGmm gmm;
    JgpuCfg cfg;
    /* histogram: src/Histogram.cpp:23-55 */
    int hist_on, hist_min, hist_max, hist_nbins, hist_count;
    int* hist_cnt;
    int hist_overflow;
    /* decoder state: src/WFSTDecoderLite.h:109-165 */
    Inst* insts;
    int n_insts, cap_insts;
    int* hook; /* arc -> inst index or -1 (WFSTTransition::hook) */
    int active_head, new_head, new_last;
    PathRec* paths;
    int n_paths, cap_paths;
    Tok best_final;
    float best_emit, norm;
    float thr_start, thr_end, thr_word, thr_emit;
    int frame;
    int n_active_insts, n_active_emit, n_active_end, n_emit_proc, n_end_proc;
    /* acoustic scorer */
    const float* x; /* current frame */
    float* gmm_cache;
    int* gmm_stamp;
    /* counters for the roofline formula */
    int* state_stamp;
    int* arc_stamp;
    JgpuStats stats;
    int frame_paths;
};

static const Tok NULL_TOK = {LZ, LZ, LZ, -1}; /* src/WFSTDecoderLite.cpp:35 */

static void* dup_mem(const void* p, size_t n)
{
    void* q = malloc(n ? n : 1);
    if (n) memcpy(q, p, n);
    return q;
}

/* ---------------------------------------------------------------------------------------
 * Histogram (src/Histogram.cpp)
 * ------------------------------------------------------------------------------------ */
static void hist_init(jor_handle* h, float minScore_, float maxScore_)
{
    /* :29-37 with binWidth = 1 */
    h->hist_min = (int)(minScore_ - 1.0);
    h->hist_max = (int)(maxScore_ + 1.0);
    h->hist_nbins = h->hist_max - h->hist_min + 1;
    h->hist_cnt = (int*)calloc((size_t)h->hist_nbins, sizeof(int));
    h->hist_count = 0;
}

static void hist_reset(jor_handle* h) /* :123-131 */
{
    h->hist_count = 0;
    memset(h->hist_cnt, 0, sizeof(int) * (size_t)h->hist_nbins);
}

static void hist_add(jor_handle* h, float score) /* :64-102 (oldScore = LOG_ZERO) */
{
    int sc;
    if (score < 0.0)
        sc = (int)(score - 0.5);
    else
        sc = (int)(score + 0.5);
    if (sc > h->hist_max) { /* :78-79 is fatal in the reference */
        h->hist_overflow = 1;
        return;
    }
    if (sc < h->hist_min) return;
    h->hist_cnt[sc - h->hist_min]++;
    h->hist_count++;
}

static float hist_thresh(jor_handle* h, int maxN) /* :134-158 */
{
    int total = 0, i;
    if (h->hist_count <= maxN) return (float)((float)(h->hist_min) - 0.5);
    for (i = h->hist_nbins - 1; i >= 0; i--) {
        total += h->hist_cnt[i];
        if (total >= maxN) return (float)((float)(i + h->hist_min) - 0.5);
    }
    return (float)(h->hist_min);
}

/* ---------------------------------------------------------------------------------------
 * HTKFlatModels::calcGMMOutput + logAdd (src/HTKFlatModels.cpp:226-293)
 * ------------------------------------------------------------------------------------ */
static float log_add(float x, float y) /* :266-293 */
{
    float diff;
    if (x < y) {
        float t = x;
        x = y;
        y = t;
    }
    diff = y - x;
    if (diff < -18.42) /* MINUS_LOG_THRESHOLD, compared in double */
        return x;
    return (float)(x + log(1.0 + exp(diff)));
}

static float gmm_eval(const JgpuGmm* g, int gi, const float* x) /* :243-255 */
{
    const int D = g->dim, C = g->max_comps, n = g->n_comps[gi];
    const float* means = g->means + (size_t)gi * C * D;
    const float* vars = g->ivars + (size_t)gi * C * D;
    const float* dets = g->dets + (size_t)gi * C;
    float logProb = LZ;
    int i, j;
    for (i = 0; i < n; ++i) {
        float sumxmu = 0.0f;
        for (j = 0; j < D; ++j) {
            float xmu = x[j] - means[j];
            sumxmu += xmu * xmu * vars[j];
        }
        means += D;
        vars += D;
        logProb = log_add(logProb, (float)(-0.5 * sumxmu + dets[i])); /* :254, double then narrowed */
    }
    return logProb;
}

static float calc_output(jor_handle* h, int hmm, int state) /* :179-200 */
{
    int gi = h->hmm.gmm[hmm * h->hmm.max_states + state];
    if (h->gmm_stamp[gi] != h->frame) {
        h->gmm_cache[gi] = gmm_eval(&h->gmm, gi, h->x);
        h->gmm_stamp[gi] = h->frame;
        h->stats.total_gmm_evals++;
    }
    return h->gmm_cache[gi];
}

/* ---------------------------------------------------------------------------------------
 * construction
 * ------------------------------------------------------------------------------------ */
jor_handle* jor_create(const JgpuNet* net, const JgpuHmm* hmm, const JgpuGmm* gmm, const JgpuCfg* cfg)
{
    jor_handle* h = (jor_handle*)calloc(1, sizeof(*h));
    const int A = net->n_arcs, S = net->n_states, H = hmm->n_hmms, M = hmm->max_states;
    const int G = gmm->n_gmms, C = gmm->max_comps, D = gmm->dim;
    int i;
    if (M > SMAX) {
        free(h);
        return NULL;
    }
    h->net = *net;
    h->net.arc_to = (const int32_t*)dup_mem(net->arc_to, sizeof(int32_t) * A);
    h->net.arc_weight = (const float*)dup_mem(net->arc_weight, sizeof(float) * A);
    h->net.arc_in = (const int32_t*)dup_mem(net->arc_in, sizeof(int32_t) * A);
    h->net.arc_out = (const int32_t*)dup_mem(net->arc_out, sizeof(int32_t) * A);
    h->net.state_first = (const int32_t*)dup_mem(net->state_first, sizeof(int32_t) * S);
    h->net.state_narcs = (const int32_t*)dup_mem(net->state_narcs, sizeof(int32_t) * S);
    h->net.state_final = (const float*)dup_mem(net->state_final, sizeof(float) * S);
    h->hmm = *hmm;
    h->hmm.n_states = (const int32_t*)dup_mem(hmm->n_states, sizeof(int32_t) * H);
    h->hmm.gmm = (const int32_t*)dup_mem(hmm->gmm, sizeof(int32_t) * H * M);
    h->hmm.trp = (const float*)dup_mem(hmm->trp, sizeof(float) * H * M * M);
    h->hmm.se = (const int32_t*)dup_mem(hmm->se, sizeof(int32_t) * H * M * 2);
    h->hmm.tee = (const float*)dup_mem(hmm->tee, sizeof(float) * H);
    h->gmm = *gmm;
    h->gmm.n_comps = (const int32_t*)dup_mem(gmm->n_comps, sizeof(int32_t) * G);
    h->gmm.dets = (const float*)dup_mem(gmm->dets, sizeof(float) * G * C);
    h->gmm.means = (const float*)dup_mem(gmm->means, sizeof(float) * (size_t)G * C * D);
    h->gmm.ivars = (const float*)dup_mem(gmm->ivars, sizeof(float) * (size_t)G * C * D);
    h->cfg = *cfg;
    /* src/WFSTDecoderLite.cpp:76-82 */
    h->hist_on = cfg->max_hyps > 0;
    if (h->hist_on) {
        if (cfg->main_beam > 0.0)
            hist_init(h, (float)(-cfg->main_beam - 800.0), (float)200.0);
        else
            hist_init(h, (float)-1000.0, (float)200.0);
    }
    h->hook = (int*)malloc(sizeof(int) * (size_t)(A ? A : 1));
    for (i = 0; i < A; ++i) h->hook[i] = -1;
    h->gmm_cache = (float*)malloc(sizeof(float) * (size_t)(G ? G : 1));
    h->gmm_stamp = (int*)malloc(sizeof(int) * (size_t)(G ? G : 1));
    h->state_stamp = (int*)malloc(sizeof(int) * (size_t)(S ? S : 1));
    h->arc_stamp = (int*)malloc(sizeof(int) * (size_t)(A ? A : 1));
    h->active_head = h->new_head = h->new_last = -1;
    return h;
}

void jor_destroy(jor_handle* h)
{
    if (!h) return;
    free((void*)h->net.arc_to); free((void*)h->net.arc_weight); free((void*)h->net.arc_in);
    free((void*)h->net.arc_out); free((void*)h->net.state_first); free((void*)h->net.state_narcs);
    free((void*)h->net.state_final);
    free((void*)h->hmm.n_states); free((void*)h->hmm.gmm); free((void*)h->hmm.trp);
    free((void*)h->hmm.se); free((void*)h->hmm.tee);
    free((void*)h->gmm.n_comps); free((void*)h->gmm.dets); free((void*)h->gmm.means); free((void*)h->gmm.ivars);
    free(h->hist_cnt); free(h->insts); free(h->hook); free(h->paths);
    free(h->gmm_cache); free(h->gmm_stamp); free(h->state_stamp); free(h->arc_stamp);
    free(h);
}

void jor_gmm_scores(jor_handle* h, const float* x, int n_rows, float* out)
{
    const int G = h->gmm.n_gmms, D = h->gmm.dim;
    int r, g;
    for (r = 0; r < n_rows; ++r)
        for (g = 0; g < G; ++g) out[(size_t)r * G + g] = gmm_eval(&h->gmm, g, x + (size_t)r * D);
}

/* ---------------------------------------------------------------------------------------
 * instances and paths
 * ------------------------------------------------------------------------------------ */
static int new_path(jor_handle* h) /* src/WFSTDecoderLite.cpp:608-620 (no GC: results do not depend on it) */
{
    if (h->n_paths == h->cap_paths) {
        h->cap_paths = h->cap_paths ? h->cap_paths * 2 : 4096;
        h->paths = (PathRec*)realloc(h->paths, sizeof(PathRec) * (size_t)h->cap_paths);
    }
    h->stats.total_paths++;
    h->frame_paths++;
    return h->n_paths++;
}

static void push_new(jor_handle* h, int ii) /* prepend: :551-555, :767-771 */
{
    h->insts[ii].next = h->new_head;
    h->new_head = ii;
    if (h->new_last < 0) h->new_last = ii;
    ++h->n_active_insts;
}

static int attach_inst(jor_handle* h, int arc) /* :751-774 */
{
    int ii, i, hmm = h->net.arc_in[arc] - 1;
    if (h->n_insts == h->cap_insts) {
        h->cap_insts = h->cap_insts ? h->cap_insts * 2 : 4096;
        h->insts = (Inst*)realloc(h->insts, sizeof(Inst) * (size_t)h->cap_insts);
    }
    ii = h->n_insts++;
    h->insts[ii].hmm = hmm;
    h->insts[ii].nst = h->hmm.n_states[hmm];
    for (i = 0; i < SMAX; ++i) h->insts[ii].st[i] = NULL_TOK;
    h->hook[arc] = ii;
    h->insts[ii].arc = arc;
    h->insts[ii].tee = h->hmm.tee[hmm];
    h->insts[ii].nactive = 0;
    push_new(h, ii);
    return ii;
}

static void join_new(jor_handle* h) /* :799-805 */
{
    if (h->new_head < 0) return;
    h->insts[h->new_last].next = h->active_head;
    h->active_head = h->new_head;
    h->new_head = h->new_last = -1;
}

/* returns next inst; unlinks `ii` whose predecessor is `prev` (-1 = head): :777-797 */
static int return_inst(jor_handle* h, int ii, int prev)
{
    int i, nx = h->insts[ii].next;
    if (prev < 0)
        h->active_head = nx;
    else
        h->insts[prev].next = nx;
    for (i = 0; i < h->insts[ii].nst; ++i) h->insts[ii].st[i] = NULL_TOK;
    --h->n_active_insts;
    return nx;
}

/* ---------------------------------------------------------------------------------------
 * propagateToken (src/WFSTDecoderLite.cpp:491-605); arc = -1 is the NULL transition
 * ------------------------------------------------------------------------------------ */
static void propagate(jor_handle* h, Tok* tok, int arc)
{
    const JgpuNet* n = &h->net;
    int state, first, cnt, b;
    if (arc >= 0) {
        if (n->arc_out[arc] != 0) { /* :497-509 */
            int p = new_path(h);
            h->paths[p].frame = h->frame;
            h->paths[p].score = tok->score;
            h->paths[p].lm = tok->lm;
            h->paths[p].ac = tok->ac;
            h->paths[p].label = n->arc_out[arc];
            h->paths[p].prev = tok->path;
            tok->path = p;
        }
        if (n->state_final[n->arc_to[arc]] > LZ) { /* :513-520 */
            float weight = n->state_final[n->arc_to[arc]];
            if (tok->score + weight > h->best_final.score) {
                h->best_final = *tok;
                h->best_final.score += weight;
                h->best_final.lm += weight;
            }
        }
    }
    state = arc < 0 ? n->init_state : n->arc_to[arc]; /* :525-527, WFSTNetwork.cpp:709-721 */
    first = n->state_first[state];
    cnt = n->state_narcs[state];
    if (h->state_stamp[state] != h->frame) {
        h->state_stamp[state] = h->frame;
        h->stats.total_arcs_expanded += cnt;
    }
    for (b = first; b < first + cnt; ++b) {
        if (n->arc_in[b] == 0) { /* :533-540 */
            Tok tmp = *tok;
            tmp.score += n->arc_weight[b];
            tmp.lm += n->arc_weight[b];
            if (tmp.score > h->thr_end) propagate(h, &tmp, b);
        } else { /* :542-601 */
            int ii = h->hook[b];
            Inst* inst;
            Tok* res;
            float newScore;
            if (ii < 0) {
                ii = attach_inst(h, b);
            } else if (h->insts[ii].nactive == 0) {
                push_new(h, ii);
            }
            inst = &h->insts[ii];
            res = &inst->st[0];
            newScore = tok->score + n->arc_weight[b];
            if (newScore > res->score) {
                if (res->score <= LZ) ++inst->nactive;
                *res = *tok;
                res->score = newScore;
                res->lm += n->arc_weight[b];
                if (newScore > h->best_emit) h->best_emit = newScore; /* :572-573, :579-580 */
                if (h->arc_stamp[b] != h->frame) {
                    h->arc_stamp[b] = h->frame;
                    h->stats.total_entry_writes++;
                }
            }
            if (inst->tee > LZ) { /* :584-600 */
                float teeWeight = inst->tee;
                Tok tmp;
                newScore += teeWeight;
                tmp = *tok;
                tmp.score = newScore;
                tmp.ac += teeWeight;
                tmp.lm += n->arc_weight[b];
                if (n->arc_out[b] != 0) {
                    if (newScore > h->thr_word) propagate(h, &tmp, b);
                } else {
                    if (newScore > h->thr_end) propagate(h, &tmp, b);
                }
                inst = &h->insts[ii]; /* pool may have been reallocated by the recursion */
            }
        }
    }
}

/* ---------------------------------------------------------------------------------------
 * HMMInternalPropagation (src/WFSTDecoderLite.cpp:376-484)
 * ------------------------------------------------------------------------------------ */
/* Instrumentation for DESIGN.md "lazy acoustic scoring" (not part of the restated algorithm): how many distinct
 * GMMs per frame have at least one live predecessor token at the start of the frame, i.e. the superset of the
 * evaluated GMMs that can be marked one step ahead, before the frame's normalisation and beam are known. */
static long long g_superset_total = 0;
static int* g_superset_stamp = NULL;
static int g_superset_n = 0;
long long jor_debug_gmm_superset(void) { return g_superset_total; }
void jor_debug_gmm_superset_reset(void) { g_superset_total = 0; }

static void hmm_internal(jor_handle* h, Inst* inst)
{
    const int M = h->hmm.max_states;
    const int N_1 = inst->nst - 1;
    const float* trP = h->hmm.trp + (size_t)inst->hmm * M * M; /* trP[i*M + j] */
    const int32_t* se = h->hmm.se + (size_t)inst->hmm * M * 2;
    Tok buf[SMAX];
    int i, j;
    if (g_superset_n != h->gmm.n_gmms) {
        free(g_superset_stamp);
        g_superset_n = h->gmm.n_gmms;
        g_superset_stamp = (int*)malloc(sizeof(int) * (size_t)(g_superset_n > 0 ? g_superset_n : 1));
        for (i = 0; i < g_superset_n; ++i) g_superset_stamp[i] = -1000;
    }
    if (h->frame == 0) for (i = 0; i < g_superset_n && inst == &h->insts[h->active_head]; ++i) g_superset_stamp[i] = -1000;
    for (j = 1; j < N_1; ++j) {
        int live = 0, gi = h->hmm.gmm[inst->hmm * M + j];
        for (i = se[j * 2 + 0]; i < se[j * 2 + 1]; ++i) live |= inst->st[i].score > LZ;
        if (live && gi >= 0 && g_superset_stamp[gi] != h->frame) { g_superset_stamp[gi] = h->frame; ++g_superset_total; }
    }
    buf[0] = NULL_TOK; /* :108 */
    for (j = 1; j < N_1; ++j) { /* :387-424 */
        Tok* res = &buf[j];
        int endi = se[j * 2 + 1];
        i = se[j * 2 + 0];
        *res = inst->st[i];
        res->score += trP[i * M + j];
        res->ac += trP[i * M + j];
        for (++i; i < endi; ++i) {
            float tmpScore = inst->st[i].score + trP[i * M + j];
            if (tmpScore > res->score) {
                *res = inst->st[i];
                res->score = tmpScore;
                res->ac += trP[i * M + j];
            }
        }
        res->score -= h->norm; /* :408 */
        if (res->score > h->thr_emit) {
            float outp;
            ++h->n_emit_proc;
            outp = calc_output(h, inst->hmm, j);
            res->score += outp;
            res->ac += outp;
            if (h->hist_on) hist_add(h, res->score);
            if (res->score > h->best_emit) h->best_emit = res->score;
        } else {
            *res = NULL_TOK;
        }
    }
    inst->nactive = 0; /* :428-436 */
    for (i = 0; i < N_1; ++i) {
        if (buf[i].score > LZ) ++inst->nactive;
        inst->st[i] = buf[i];
    }
    h->n_active_emit += inst->nactive;
    { /* exit state: :443-483, reads the NEW emitting tokens */
        Tok* res = &inst->st[N_1];
        int endi = se[N_1 * 2 + 1];
        i = se[N_1 * 2 + 0];
        *res = inst->st[i];
        res->score += trP[i * M + N_1];
        res->ac += trP[i * M + N_1];
        for (++i; i < endi; ++i) {
            float tmpScore = inst->st[i].score + trP[i * M + N_1];
            if (tmpScore > res->score) {
                *res = inst->st[i];
                res->score = tmpScore;
                res->ac += trP[i * M + N_1];
            }
        }
        if (res->score <= LZ) {
            *res = NULL_TOK;
        } else {
            ++inst->nactive;
            ++h->n_active_end;
        }
    }
}

static void do_internal(jor_handle* h) /* :899-935 */
{
    int prev = -1, ii = h->active_head;
    h->n_active_emit = h->n_active_end = h->n_emit_proc = h->n_end_proc = 0;
    h->best_emit = LZ;
    while (ii >= 0) {
        Inst* inst = &h->insts[ii];
        Tok* entry = &inst->st[0];
        if (entry->score > LZ && entry->score < h->thr_start) { /* :915-918 */
            *entry = NULL_TOK;
            --inst->nactive;
        }
        hmm_internal(h, inst);
        if (inst->nactive == 0) {
            ii = return_inst(h, ii, prev);
        } else {
            prev = ii;
            ii = inst->next;
        }
    }
    h->stats.total_active_emit_hyps += h->n_active_emit;
    h->stats.total_active_end_hyps += h->n_active_end;
    h->stats.total_proc_emit_hyps += h->n_emit_proc;
}

static void do_external(jor_handle* h) /* :937-982 */
{
    int prev = -1, ii = h->active_head;
    h->n_end_proc = 0;
    while (ii >= 0) {
        int arc = h->insts[ii].arc;
        int xs = h->insts[ii].nst - 1;
        if (h->insts[ii].st[xs].score > LZ) {
            Tok ex = h->insts[ii].st[xs]; /* propagate may grow the pool: work on a copy */
            float thr = h->net.arc_out[arc] == 0 ? h->thr_end : h->thr_word; /* :952-962 */
            if (ex.score > thr) {
                ++h->n_end_proc;
                propagate(h, &ex, arc);
            }
            h->insts[ii].st[xs] = NULL_TOK; /* :964 */
            if (--h->insts[ii].nactive == 0) {
                ii = return_inst(h, ii, prev);
            } else {
                prev = ii;
                ii = h->insts[ii].next;
            }
        } else {
            prev = ii;
            ii = h->insts[ii].next;
        }
    }
    h->stats.total_proc_end_hyps += h->n_end_proc;
    join_new(h);
    h->stats.total_active_models += h->n_active_insts;
}

/* ---------------------------------------------------------------------------------------
 * recognitionStart / processFrame / recognitionFinish
 * ------------------------------------------------------------------------------------ */
static void recognition_start(jor_handle* h) /* :139-228 */
{
    int i;
    Tok tmp;
    h->frame = 0;
    h->best_final = NULL_TOK;
    h->n_insts = 0; /* per-utterance: drop every instance and hook (results are unaffected
                       by the reference's lazy MaxAllocModels policy, :164-169) */
    for (i = 0; i < h->net.n_arcs; ++i) h->hook[i] = -1;
    h->active_head = h->new_head = h->new_last = -1;
    h->n_paths = 0;
    if (h->hist_on) hist_reset(h);
    h->hist_overflow = 0;
    h->norm = 0.0f;
    h->best_emit = LZ;
    h->thr_start = h->thr_end = h->thr_word = h->thr_emit = LZ;
    h->n_active_insts = h->n_active_emit = h->n_active_end = h->n_emit_proc = h->n_end_proc = 0;
    memset(&h->stats, 0, sizeof(h->stats));
    for (i = 0; i < h->gmm.n_gmms; ++i) h->gmm_stamp[i] = -1000;
    for (i = 0; i < h->net.n_states; ++i) h->state_stamp[i] = -1000;
    for (i = 0; i < h->net.n_arcs; ++i) h->arc_stamp[i] = -1000;
    tmp.score = 0.0f;
    tmp.ac = 0.0f;
    tmp.lm = 0.0f;
    tmp.path = -1;
    propagate(h, &tmp, -1); /* :221-227, with currFrame = 0 and every threshold = LOG_ZERO */
    join_new(h);
    memset(&h->stats, 0, sizeof(h->stats));
    for (i = 0; i < h->net.n_states; ++i) h->state_stamp[i] = -1000;
    for (i = 0; i < h->net.n_arcs; ++i) h->arc_stamp[i] = -1000;
}

static void process_frame(jor_handle* h, const float* x, int frame) /* :311-372 */
{
    const JgpuCfg* c = &h->cfg;
    h->frame = frame;
    h->x = x; /* newFrame: HTKFlatModels.cpp:295-306 */
    h->best_final = NULL_TOK;
    h->frame_paths = 0;
    h->norm = (h->best_emit > LZ ? h->best_emit : 0.0f); /* :321 */
    if (h->hist_on) { /* :322-329 */
        h->thr_emit = hist_thresh(h, c->max_hyps);
        h->thr_emit -= h->norm;
        if (c->main_beam > 0.0 && h->thr_emit < -c->main_beam) h->thr_emit = -c->main_beam;
        hist_reset(h);
    } else {
        h->thr_emit = (c->main_beam > 0.0 ? -c->main_beam : LZ); /* :331 */
    }
    h->thr_start = (c->start_beam > 0.0 ? (h->best_emit - c->start_beam) : LZ); /* :337 */
    do_internal(h);
    h->thr_end = (c->end_beam > 0.0 ? (h->best_emit - c->end_beam) : LZ);   /* :349 */
    h->thr_word = (c->word_beam > 0.0 ? (h->best_emit - c->word_beam) : LZ); /* :350 */
    do_external(h);
    h->stats.n_frames++;
}

static int recognition_finish(jor_handle* h, JgpuResult* out) /* :230-309 */
{
    Tok best = h->best_final;
    int n = 0, p, k;
    out->score = out->ac = out->lm = LZ;
    if (best.score == LZ) {
        out->status = -1; /* :264-267 NULL */
        return out->status;
    }
    for (p = best.path; p >= 0; p = h->paths[p].prev) ++n;
    if (n == 0) { /* :273-306: while body never runs, DecHyp keeps its ctor values */
        out->status = -2;
        return out->status;
    }
    out->score = best.score;
    out->ac = best.ac;
    out->lm = best.lm;
    k = n;
    for (p = best.path; p >= 0; p = h->paths[p].prev) {
        --k;
        if (k < out->max_words && out->words) {
            JgpuWord* w = &out->words[k];
            w->label = h->paths[p].label;
            w->time = h->paths[p].frame;
            w->score = h->paths[p].score;
            w->ac = h->paths[p].ac;
            w->lm = h->paths[p].lm;
            if (k == n - 1) { /* :293-295 newest record carries the final-weight-inclusive totals */
                w->lm = best.lm;
                w->ac = best.ac;
                w->score = best.score;
            }
        }
    }
    out->status = n;
    return n;
}

int jor_decode(jor_handle* h, const float* feats, int n_frames, JgpuResult* out, int* frame_cnt,
               float* frame_best, double* seconds)
{
    const int D = h->gmm.dim;
    struct timespec t0, t1;
    int t;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    recognition_start(h);
    for (t = 0; t < n_frames; ++t) {
        process_frame(h, feats + (size_t)t * D, t);
        if (frame_cnt) {
            int* c = frame_cnt + (size_t)t * 6;
            c[0] = h->n_active_insts;
            c[1] = h->n_active_emit;
            c[2] = h->n_active_end;
            c[3] = h->n_emit_proc;
            c[4] = h->n_end_proc;
            c[5] = h->frame_paths;
        }
        if (frame_best) frame_best[t] = h->best_emit;
    }
    out->n_frames = n_frames;
    if (n_frames > 0) h->frame = n_frames - 1;
    recognition_finish(h, out);
    if (h->hist_overflow) out->status = -20;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    return out->status;
}

void jor_stats(jor_handle* h, JgpuStats* out) { *out = h->stats; }
