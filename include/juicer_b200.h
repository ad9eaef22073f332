/* juicer_b200.h — C ABI of the B200-native WFST Viterbi token-passing decode path.
 *
 * One path of idiap/juicer is re-implemented here as hand-written sm_100a CUDA kernels:
 *   WFSTDecoderLite::processFrame        (reference src/WFSTDecoderLite.cpp:311-372)
 *   HTKFlatModels::calcOutput            (reference src/HTKFlatModels.cpp:179-262)
 * plus the per-utterance start/finish around it (recognitionStart :139-228,
 * recognitionFinish :230-309).  Everything crossing this boundary is a plain pointer, a
 * size or a POD struct; no C++ or torch types.  All functions return 0 on success, a
 * negative JGPU_E_* code otherwise; jgpu_last_error() gives the message.  There is no CPU
 * fallback: without a CUDA device jgpu_create fails with JGPU_E_CUDA.
 *
 * Conventions (identical to the reference):
 *   real == float;  LOG_ZERO == -FLT_MAX is the "dead token" sentinel (Torch3 log_add.h);
 *   arc weights are +log probabilities, already negated / LM-scaled / insertion-penalised
 *   the way WFSTNetwork's loader stores them (src/WFSTNetwork.cpp:481-486);
 *   an arc's input label is HMM index + 1 (src/WFSTDecoderLite.cpp:754), 0 = epsilon;
 *   an arc's output label is word id + 1, 0 = epsilon.
 */
#ifndef JUICER_B200_H
#define JUICER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JGPU_LOG_ZERO (-3.402823466e+38f)

enum {
    JGPU_OK          =  0,
    JGPU_E_ARG       = -1,   /* bad argument / inconsistent tables (validated at create) */
    JGPU_E_CUDA      = -2,   /* CUDA runtime failure, or no device: there is no CPU path */
    JGPU_E_CAPACITY  = -3,   /* a device arena overflowed (reported per utterance too) */
    JGPU_E_STATE     = -4,   /* call sequence error (e.g. push_frames on a lane not begun) */
    JGPU_E_IO        = -5    /* host loader could not read / parse a file */
};

/* Flattened WFSTNetwork.  Replaces: WFSTTransition / WFSTState / WFSTFinalState
 * (src/WFSTNetwork.h:41-68) as seen through getTransitions(prev,&next)
 * (src/WFSTNetwork.cpp:709-721: a state's out-arcs are arcs [first, first+n) in file
 * order), transGoesToFinalState / getFinalStateWeight (src/WFSTNetwork.h:161-167) and
 * getInitState (:126). */
typedef struct JgpuNet {
    int32_t        n_states;
    int32_t        n_arcs;
    int32_t        init_state;
    const int32_t* arc_to;        /* [n_arcs] destination state                       */
    const float*   arc_weight;    /* [n_arcs] log weight as stored by the loader      */
    const int32_t* arc_in;        /* [n_arcs] input label  (HMM index + 1, 0 = eps)   */
    const int32_t* arc_out;       /* [n_arcs] output label (word id + 1, 0 = eps)     */
    const int32_t* state_first;   /* [n_states] id of the state's first out-arc       */
    const int32_t* state_narcs;   /* [n_states] number of out-arcs                    */
    const float*   state_final;   /* [n_states] final weight, JGPU_LOG_ZERO if the
                                     state is not final                               */
} JgpuNet;

/* HMM topology view of IModels.  Replaces: getNumHMMs / getNumStates / getTransMat /
 * getSEIndex / getTeeLogProb (src/Models.h:51-66; tables built by
 * HTKModels::createTrPandSEIndex, src/HTKModels.cpp:2330-2390, tee :1358-1370) and the
 * state -> GMM map HMM::gmmInds (src/HTKModels.h:75-82).  stride = max_states. */
typedef struct JgpuHmm {
    int32_t        n_hmms;
    int32_t        max_states;    /* S: row stride of the per-HMM tables below (<= 8)  */
    const int32_t* n_states;      /* [n_hmms] incl. non-emitting entry (0) and exit    */
    const int32_t* gmm;           /* [n_hmms*S] GMM id per state, -1 if non-emitting   */
    const float*   trp;           /* [n_hmms*S*S] trP[i][j] = log a_ij or LOG_ZERO     */
    const int32_t* se;            /* [n_hmms*S*2] SEIndex[j] = {start,end} of preds    */
    const float*   tee;           /* [n_hmms] entry->exit log prob or LOG_ZERO         */
} JgpuHmm;

/* Flat diagonal-GMM parameters.  Replaces: HTKFlatModels::fMixtures / fDets / fMeans /
 * fVars (src/HTKFlatModels.h:43-63, filled at src/HTKFlatModels.cpp:94-177): every GMM is
 * padded to max_comps slots; ivars = 1/var; dets = gconst + log weight. */
typedef struct JgpuGmm {
    int32_t        n_gmms;
    int32_t        dim;           /* feature vector size (39 in all configs)           */
    int32_t        max_comps;     /* C: slots per GMM                                  */
    const int32_t* n_comps;       /* [n_gmms] used slots                               */
    const float*   dets;          /* [n_gmms*C]                                        */
    const float*   means;         /* [n_gmms*C*dim]                                    */
    const float*   ivars;         /* [n_gmms*C*dim]                                    */
} JgpuGmm;

/* Decoder settings.  The five pruning fields are the constructor arguments of
 * WFSTDecoderLite (src/WFSTDecoderLite.h:81-89; call site src/juicer.cpp:584-586) with
 * the same meaning: a beam <= 0 disables that pruning, max_hyps 0 disables the histogram. */
typedef struct JgpuCfg {
    float   start_beam;           /* phoneStartPruneWin */
    float   main_beam;            /* emitPruneWin       */
    float   end_beam;             /* phoneEndPruneWin   */
    float   word_beam;            /* wordPruneWin       */
    int32_t max_hyps;             /* maxEmitHyps        */
    int32_t device;               /* CUDA device ordinal */
    int32_t n_lanes;              /* utterances decoded in lock-step per launch (>=1)  */
    int32_t max_active;           /* first-pass capacity: active HMM instances per lane, 0 = auto (up to one per
                                     arc, bounded by 30 % of the free device memory over n_lanes)          */
    int32_t max_frames;           /* frames per utterance KEPT by frame_stats, 0 = 4096 (decoding itself has no
                                     frame limit)                                                          */
    int32_t max_paths;            /* capacity: word-boundary records per lane, 0 = auto*/
    int32_t frame_stats;          /* 1 = keep per-frame work counters (parity tests)   */
    int32_t reserved;
} JgpuCfg;

/* One word-boundary record of the best path.  Replaces DecHypHist
 * (src/DecHypHistPool.h:38-49): label = state (output label = word id + 1), time = frame
 * in which the word's arc was left, score = normalised running score, ac / lm =
 * un-normalised cumulative acoustic and LM scores. */
typedef struct JgpuWord {
    int32_t label;
    int32_t time;
    float   score;
    float   ac;
    float   lm;
} JgpuWord;

/* Result of one utterance.  Replaces DecHyp* returned by IDecoder::finish()
 * (src/Decoder.h:29; built at src/WFSTDecoderLite.cpp:262-308).
 *   status >= 0 : number of words on the best path.  words[] holds min(status, max_words) of them, oldest
 *                 first; when the caller's buffer is the smaller one it gets the NEWEST max_words words, so
 *                 the last record always carries the final-weight-inclusive totals.  The device side has
 *                 no per-utterance word limit (the reference allocates one DecHypHist per word);
 *   status == -1: no token reached a final state in the last frame (reference: NULL hyp);
 *   status == -2: a final token survived but its path has no word label (reference:
 *                 non-NULL but inactive DecHyp, WFSTDecoderLite.cpp:273-306);
 *   status <= -10: the utterance failed: status = JGPU_E_CAPACITY - 10 - (error bits << 8), bits: 1 active
 *                 list, 2 arrival records, 4 word-boundary arena, 8 histogram range (the reference calls
 *                 error() there, src/Histogram.cpp:75-79), 64 result-word pool.  The batch entry points
 *                 decode such an utterance again with larger per-lane arenas (fewer lanes in lock-step), up
 *                 to one instance per arc — what the reference can reach by allocating — so only the
 *                 streaming interface and a handle whose memory cannot hold one full-size lane report it. */
typedef struct JgpuResult {
    int32_t   status;
    int32_t   n_frames;
    float     score;
    float     ac;
    float     lm;
    int32_t   max_words;          /* in: capacity of words[]                           */
    JgpuWord* words;              /* in: caller-owned buffer                           */
} JgpuResult;

/* Per-utterance work counters, same definitions as the reference's statistics print
 * (src/WFSTDecoderLite.cpp:231-241), as sums over frames. */
typedef struct JgpuStats {
    int64_t n_frames;
    int64_t total_active_models;     /* sum of nActiveInsts after each frame            */
    int64_t total_active_emit_hyps;
    int64_t total_active_end_hyps;
    int64_t total_proc_emit_hyps;
    int64_t total_proc_end_hyps;
    int64_t total_gmm_evals;         /* distinct (GMM, frame) scores computed           */
    int64_t total_arcs_expanded;     /* X of SURVEY 8d: out-arcs of distinct states     */
    int64_t total_entry_writes;      /* W: distinct destination arcs written            */
    int64_t total_paths;             /* P: word-boundary records appended               */
} JgpuStats;

typedef struct jgpu_handle jgpu_handle;

const char* jgpu_last_error(void);
const char* jgpu_version(void);

/* Build device-resident tables and arenas.  Validates the reference's unchecked input
 * preconditions: label ranges, arc ranges, no epsilon/tee cycles.
 * Limits the reference does not have, each refused here with JGPU_E_ARG and a message: HMMs of at most 8 states
 * (entry and exit included); feature dimension <= 64; <= 256 components per GMM; n_lanes <= 512; epsilon / tee
 * chains of at most 15 arcs; < 65536 epsilon and < 65536 tee out-arcs per state; < 2^27 arcs and < 2^30 states.
 * Per-frame statistics (cfg.frame_stats) keep the first max_frames frames of an utterance; decoding has no
 * frame, instance or word limit (see JgpuResult.status). */
int jgpu_create(const JgpuNet* net, const JgpuHmm* hmm, const JgpuGmm* gmm, const JgpuCfg* cfg,
                jgpu_handle** out);
int jgpu_destroy(jgpu_handle* h);

/* Acoustic scorer alone: out[r*n_gmms + g] = log-likelihood of GMM g for feature row r.
 * Replaces IModels::calcOutput(int gmmInd) (src/HTKFlatModels.cpp:202-262) evaluated for
 * every GMM of every row.  x and out are HOST pointers. */
int jgpu_gmm_scores(jgpu_handle* h, const float* x, int32_t n_rows, float* out);

/* Streaming interface on one lane.  Replaces IDecoder::init / processFrame / finish
 * (src/Decoder.h:18-30): begin == init(); push_frames(x, n) == n consecutive
 * processFrame() calls on frames x[0..n); end == finish().  x is a HOST pointer to
 * n*dim floats. */
int jgpu_utt_begin(jgpu_handle* h, int32_t lane);
int jgpu_push_frames(jgpu_handle* h, int32_t lane, const float* x, int32_t n_frames);
int jgpu_utt_end(jgpu_handle* h, int32_t lane, JgpuResult* out);

/* Streaming partial result of the utterance running on `lane`, between two jgpu_push_frames calls: the
 * word-boundary records EVERY live hypothesis has in its history — they can no longer change.  Replaces
 * tracePartialPath / traceWinningPaths / partialPaths (src/WFSTDecoderLite.cpp:822-897, src/WFSTDecoderLite.h:200),
 * which the reference only runs behind its path collector (:358-370) and only prints at the end (:247-257).
 *   status >= 0 : number of converged words (words[] as in jgpu_utt_end, without final weights);
 *   status == -1: no live hypothesis / nothing decoded yet;
 *   status == -3: some live instance has no word in its history yet (the reference's trace walks off that
 *                 instance's token array there, :846-850), so there is no common record. */
int jgpu_partial_result(jgpu_handle* h, int32_t lane, JgpuResult* out);

/* Whole-utterance batch: decodes n_utts independent utterances, n_lanes at a time in
 * lock-step, refilling lanes as utterances finish.  feats[u] is a HOST pointer to
 * n_frames[u]*dim floats.  Replaces the per-file loop of DecoderBatchTest::run
 * (src/DecoderBatchTest.cpp:690-777) around DecoderSingleTest::decodeUtterance. */
int jgpu_decode_batch(jgpu_handle* h, const float* const* feats, const int32_t* n_frames,
                      int32_t n_utts, JgpuResult* out);

/* Same, with all features already resident in device memory as one packed buffer:
 * utterance u occupies rows [row_offset[u], row_offset[u] + n_frames[u]) of d_feats
 * (DEVICE pointer, row-major [rows, dim]).  No host<->device feature traffic. */
int jgpu_decode_batch_device(jgpu_handle* h, const float* d_feats, const int64_t* row_offset,
                             const int32_t* n_frames, int32_t n_utts, JgpuResult* out);

/* ---- whole-utterance work stealing across the GPUs of one node (BASELINE configs[3]) ----------------
 * The reference decodes a file list one utterance after the other (DecoderBatchTest::run,
 * src/DecoderBatchTest.cpp:690-777); here one process per GPU decodes the SAME list, and a rank whose lanes
 * run dry takes the next utterance of the list from a counter shared through POSIX shared memory (an
 * atomic fetch-add per claim; no collective on the data path).
 *   jgpu_queue_open   : `create` != 0 on ONE rank (creates the segment, counter = 0); the others open it
 *                       after a barrier of the caller's process group.  name: [A-Za-z0-9_.-]+, node-local.
 *   jgpu_queue_reset  : counter = 0 for the next list (one rank, between two barriers).
 *   jgpu_queue_claim  : returns the first of `n` consecutive list positions now owned by the caller.
 *   jgpu_queue_close  : unmaps; the creator also unlinks the segment.
 *   jgpu_decode_queue : decodes what this rank can claim of the n_utts utterances, visiting them in
 *                       `order` (n_utts indices, the same on every rank — e.g. longest first; NULL = by
 *                       decreasing length).  feats[u] are HOST pointers; the features of an utterance are
 *                       copied to the device when it is claimed.  out[u] is written for claimed utterances
 *                       only; claimed[k], k < *n_claimed, lists them.  busy_ms (optional) = device time between
 *                       this rank's first and last kernel.  The schedule is built at most ~2 x 64 frame steps
 *                       ahead of the device, so a claim reflects how far this GPU really is.
 *   jgpu_decode_queue_device : the same with all features resident on this rank's device as one packed
 *                       buffer (see jgpu_decode_batch_device). */
typedef struct jgpu_queue jgpu_queue;
int     jgpu_queue_open(const char* name, int32_t create, jgpu_queue** out);
int     jgpu_queue_reset(jgpu_queue* q);
int64_t jgpu_queue_claim(jgpu_queue* q, int64_t n);
int64_t jgpu_queue_position(jgpu_queue* q);
int     jgpu_queue_close(jgpu_queue* q);
int jgpu_decode_queue(jgpu_handle* h, jgpu_queue* q, const float* const* feats, const int32_t* n_frames,
                      int32_t n_utts, const int32_t* order, JgpuResult* out, int32_t* claimed,
                      int32_t* n_claimed, double* busy_ms);
int jgpu_decode_queue_device(jgpu_handle* h, jgpu_queue* q, const float* d_feats, const int64_t* row_offset,
                             const int32_t* n_frames, int32_t n_utts, const int32_t* order, JgpuResult* out,
                             int32_t* claimed, int32_t* n_claimed, double* busy_ms);

/* Counters of the most recent utterance decoded on `lane` (streaming) or summed over the
 * most recent batch call (lane = -1). */
int jgpu_stats(jgpu_handle* h, int32_t lane, JgpuStats* out);

/* Per-frame work counters of the most recent utterance on `lane` (needs cfg.frame_stats):
 * cnt[t*4 + {0,1,2,3}] = nActiveInsts, nActiveEmitHyps, nActiveEndHyps, nEndHypsProcessed
 * after frame t; best[t] = bestEmitScore after frame t.  Returns frames written. */
int jgpu_frame_stats(jgpu_handle* h, int32_t lane, int32_t* cnt, float* best, int32_t max_frames);

/* Measurement utility for bench.py's secondary roofline: the FP32 issue peak of this device for code that may
 * not contract multiplies and adds into FMAs (the scorer's exact operand order forbids it), in 10^12 scalar
 * operations per second, measured now on the handle's stream. */
int jgpu_ubench_fp32(jgpu_handle* h, double* tera_ops);

/* Number of kernel launches issued by this handle since creation (bench bookkeeping). */
int64_t jgpu_launch_count(jgpu_handle* h);

/* Second passes run by the batch entry points so far: utterances whose lane overflowed an arena are decoded again
 * with the pools viewed as fewer lanes with larger arenas (see JgpuResult.status). */
int64_t jgpu_retry_count(jgpu_handle* h);

/* Enqueue all work of this handle on a caller-owned CUDA stream (a cudaStream_t passed as
 * void*), e.g. so that the caller's CUDA events bracket the decoder's kernels. */
int jgpu_set_stream(jgpu_handle* h, void* cuda_stream);

/* Per-kernel device timing: while enabled, every launch is bracketed by CUDA events on the
 * launching stream.  jgpu_profile_read fills ms[k] / count[k] for k < JGPU_N_KERNELS (summed
 * since the last enable) and returns JGPU_N_KERNELS. */
enum { JGPU_K_GMM = 0, JGPU_K_BOUNDARY, JGPU_K_INTERNAL, JGPU_K_SEED /* k_filter */, JGPU_K_EXPAND,
       JGPU_K_EXPAND_HUGE /* k_commit_huge */, JGPU_K_COMMIT, JGPU_K_EXPAND_R1, JGPU_K_EXPAND_R2,
       JGPU_N_KERNELS };   /* EXPAND = expansion round 0, _R1 = round 1, _R2 = rounds >= 2 */
int jgpu_profile(jgpu_handle* h, int32_t enable);
int jgpu_profile_read(jgpu_handle* h, double* ms, int64_t* count);
const char* jgpu_kernel_name(int32_t kind);

/* ---- host-side loaders (C++ mirrors of the reference's file readers) ------------------
 * jgpu_load_fsm   : AT&T text network + symbol tables, same semantics as
 *                   WFSTNetwork(fsm, insyms, outsyms, lmScale, insPenalty, REMOVEBOTH)
 *                   (src/WFSTNetwork.cpp:371-616).
 * jgpu_load_jmbi  : JMBI model binary, same semantics as HTKFlatModels::readBinary
 *                   (src/HTKModels.cpp:1112-1245 + src/HTKFlatModels.cpp:94-177).
 * The returned structs point into library-owned host memory released by jgpu_free_*. */
int jgpu_load_fsm(const char* fsm, const char* insyms, const char* outsyms, float lm_scale,
                  float ins_penalty, JgpuNet* out);
/* jgpu_load_jwnt  : JWNT binary network, same semantics as WFSTNetwork::readBinary
 *                   (src/WFSTNetwork.cpp:1228-1370; alphabets :250-297): transition weights are scaled by
 *                   lm_scale and get ins_penalty on arcs with an output label after reading. */
int jgpu_load_jwnt(const char* path, float lm_scale, float ins_penalty, JgpuNet* out);
int jgpu_free_net(JgpuNet* net);
int jgpu_load_jmbi(const char* path, JgpuHmm* hmm, JgpuGmm* gmm);
/* jgpu_load_mmf   : HTK MMF text model definitions, same semantics as
 *                   HTKFlatModels::Load(htkModelsFName, removeInitialToFinalTransitions)
 *                   (src/HTKFlatModels.cpp:89-92 -> src/HTKModels.cpp:221-283): the tokens of
 *                   src/htkparse.l.lpp, the grammar and checks of src/htkparse.y.ypp (~o ~h ~s ~t ~m ~v
 *                   macros; ~u and other constructs are syntax errors as in the reference), then
 *                   initFromHTKParseResult (src/HTKModels.cpp:397-444, 519-974), createTrPandSEIndex and the
 *                   flat GMM tables.  remove_initial_to_final != 0 drops entry->exit (tee) transitions and
 *                   re-normalises the entry row (src/HTKModels.cpp:921-969; juicer's -removeTeeModels). */
int jgpu_load_mmf(const char* path, int32_t remove_initial_to_final, JgpuHmm* hmm, JgpuGmm* gmm);
int jgpu_free_models(JgpuHmm* hmm, JgpuGmm* gmm);

#ifdef __cplusplus
}
#endif
#endif /* JUICER_B200_H */
