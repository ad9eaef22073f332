/* TEST INFRASTRUCTURE ONLY — never linked into, imported by or shipped with the product.
 *
 * Thin extern "C" driver around the UNMODIFIED reference objects (compiled where they lie
 * under /root/reference/src by oracle/Makefile into oracle/_ref/liboracle_ref.so):
 *   WFSTDecoderLite   (src/WFSTDecoderLite.cpp)   the decoder under test
 *   HTKFlatModels     (src/HTKFlatModels.cpp)     the diagonal-GMM scorer
 *   WFSTNetwork       (src/WFSTNetwork.cpp)       AT&T-text network loader
 * It exists so that tests and bench.py's reference arm can (a) run the real reference on
 * the same files the product reads, (b) read back the reference's own flattened tables and
 * per-frame work counters (protected members, reached through subclasses), and (c) time the
 * reference CPU path.  Nothing here restates the algorithm; see juicer_oracle.c for that.
 */
#include <chrono>
#include <vector>
#include <cstdio>
#include <cstring>

#include "WFSTDecoderLite.h"
#include "HTKFlatModels.h"
#include "WFSTNetwork.h"
#include "LogFile.h"

using namespace Juicer;

namespace {

/* protected-member access: WFSTDecoderLite.h:109-165 */
class RefDecoder : public WFSTDecoderLite {
public:
    RefDecoder(WFSTNetwork* n, IModels* m, real sb, real mb, real eb, real wb, int mh)
        : WFSTDecoderLite(n, m, sb, mb, eb, wb, mh) {}
    int   cntActiveInsts()    const { return nActiveInsts; }
    int   cntActiveEmit()     const { return nActiveEmitHyps; }
    int   cntActiveEnd()      const { return nActiveEndHyps; }
    int   cntEmitProcessed()  const { return nEmitHypsProcessed; }
    int   cntEndProcessed()   const { return nEndHypsProcessed; }
    float best()              const { return bestEmitScore; }
    int   cntPaths()          const { return nPath; }
#ifdef PARTIAL_DECODING
    /* tracePartialPath (src/WFSTDecoderLite.cpp:822-871) on demand.  The reference picks, per instance, the first token
     * with a non-NULL path by walking states[] WITHOUT a bound (:846-850): an instance none of whose tokens has a
     * word in its history sends it past the array.  Returns -1 (and traces nothing) when such an instance exists. */
    int partialTrace(std::vector<int>* labels, std::vector<int>* frames)
    {
        for (NetInst* inst = activeNetInstList; inst; inst = inst->next) {
            bool any = false;
            for (int i = 0; i < inst->nStates && !any; ++i) any = inst->states[i].path != NULL;
            if (!any) return -1;
        }
        tracePartialPath();
        labels->clear(); frames->clear();
        for (size_t i = 0; i < partialPaths.size(); ++i) { labels->push_back(partialPaths[i]->label); frames->push_back(partialPaths[i]->frame); }
        return (int)partialPaths.size();
    }
#endif
};

/* protected-member access: HTKFlatModels.h:43-63, HTKModels.h:139-170 */
class RefModels : public HTKFlatModels {
public:
    int dimVec()     const { return vecSize; }
    int dimGMMs()    const { return nGMMs; }
    int dimHMMs()    const { return nHMMs; }
    int dimTMats()   const { return nTransMats; }
    int maxComps()   const { int m = 0; for (int i = 0; i < nMixtures; ++i) m = std::max(m, mixtures[i].nComps); return m; }
    int maxStates()  const { int m = 0; for (int i = 0; i < nHMMs; ++i) m = std::max(m, hMMs[i].nStates); return m; }
    const HMM& hmm(int i) const { return hMMs[i]; }
    const TransMatrix& tmat(int i) const { return transMats[i]; }
    int  ncomp(int g) const { return fMixtures[g].compNum; }
    const real* det(int g)  { return fDet(g); }
    const real* mean(int g) { return fMean(g); }
    const real* ivar(int g) { return fVar(g); }
    int stride() const { return fvecSize4; }
};

struct Handle {
    RefModels*   models;
    WFSTNetwork* net;
    RefDecoder*  dec;
    int          blockSize;
};

} // namespace

extern "C" {

struct OrefWord { int label; int time; float score; float ac; float lm; };

void* oref_create(const char* jmbi, const char* fsm, const char* insyms, const char* outsyms,
                  float lmScale, float insPenalty,
                  float startBeam, float mainBeam, float endBeam, float wordBeam, int maxHyps,
                  int blockSize)
{
    Handle* h = new Handle;
    h->blockSize = blockSize;
    h->models = new RefModels;
    h->models->setBlockSize(blockSize);          /* must precede readBinary: HTKFlatModels.cpp:308-314 */
    h->models->readBinary(jmbi);
    h->net = new WFSTNetwork(fsm, insyms, outsyms, lmScale, insPenalty, REMOVEBOTH);
    h->dec = new RefDecoder(h->net, h->models, startBeam, mainBeam, endBeam, wordBeam, maxHyps);
    return h;
}

/* MMF text models (HTKFlatModels::Load -> HTKModels::Load, src/HTKModels.cpp:221-283): the UNMODIFIED reference
 * semantics on top of oracle/shim/htkparse_rd.cpp (bison/flex are not available to generate the reference's own
 * parser).  oref_create_mmf is oref_create with the models read from MMF text; oref_models_from_mmf opens a
 * models-only handle (net and decoder stay NULL: only oref_model_dims / oref_dump_models / oref_write_models /
 * oref_gmm_scores / oref_destroy may be used on it); oref_models_from_jmbi is the same for a JMBI file. */
void* oref_create_mmf(const char* mmf, int removeTee, const char* fsm, const char* insyms, const char* outsyms,
                      float lmScale, float insPenalty,
                      float startBeam, float mainBeam, float endBeam, float wordBeam, int maxHyps, int blockSize)
{
    Handle* h = new Handle;
    h->blockSize = blockSize;
    h->models = new RefModels;
    h->models->setBlockSize(blockSize);
    h->models->Load(mmf, removeTee != 0);
    h->net = new WFSTNetwork(fsm, insyms, outsyms, lmScale, insPenalty, REMOVEBOTH);
    h->dec = new RefDecoder(h->net, h->models, startBeam, mainBeam, endBeam, wordBeam, maxHyps);
    return h;
}

void* oref_models_from_mmf(const char* mmf, int removeTee, int blockSize)
{
    Handle* h = new Handle;
    h->blockSize = blockSize; h->net = NULL; h->dec = NULL;
    h->models = new RefModels;
    h->models->setBlockSize(blockSize);
    h->models->Load(mmf, removeTee != 0);
    return h;
}

void* oref_models_from_jmbi(const char* jmbi, int blockSize)
{
    Handle* h = new Handle;
    h->blockSize = blockSize; h->net = NULL; h->dec = NULL;
    h->models = new RefModels;
    h->models->setBlockSize(blockSize);
    h->models->readBinary(jmbi);
    return h;
}

/* HTKModels::output (src/HTKModels.cpp:993-1109): binary != 0 writes JMBI, else the reference's MMF text form. */
int oref_write_models(void* hv, const char* path, int binary)
{
    Handle* h = (Handle*)hv;
    h->models->output(path, binary != 0);
    return 0;
}

/* dims[0..5] = vecSize nGMMs nHMMs nTransMats maxStates maxComps */
void oref_model_dims(void* hv, int* dims)
{
    Handle* h = (Handle*)hv;
    dims[0] = h->models->dimVec();   dims[1] = h->models->dimGMMs();
    dims[2] = h->models->dimHMMs();  dims[3] = h->models->dimTMats();
    dims[4] = h->models->maxStates(); dims[5] = h->models->maxComps();
}

/* A new decoder with other pruning settings on the SAME network and models (loading a 6M-arc text network takes
 * longer than decoding the test utterances).  The reference leaves its instances hooked to the network's
 * transitions when a decoder dies (WFSTTransition::hook, src/WFSTNetwork.h:50-51; attachNetInst asserts the hook is
 * NULL, src/WFSTDecoderLite.cpp:752), so the hooks are cleared first. */
void oref_set_decoder(void* hv, float startBeam, float mainBeam, float endBeam, float wordBeam, int maxHyps)
{
    Handle* h = (Handle*)hv;
    delete h->dec;
    const int n = h->net->getNumTransitions();
    for (int i = 0; i < n; ++i) h->net->getOneTransition(i)->hook = NULL;
    h->dec = new RefDecoder(h->net, h->models, startBeam, mainBeam, endBeam, wordBeam, maxHyps);
}

void oref_destroy(void* hv)
{
    Handle* h = (Handle*)hv;
    delete h->dec; delete h->net; delete h->models; delete h;
}

/* JWNT binary networks (WFSTNetwork::writeBinary / readBinary, src/WFSTNetwork.cpp:1106-1370): write the
 * network of a handle, or open a network-only handle from a JWNT file (models and decoder stay NULL; only
 * oref_net_dims / oref_dump_net / oref_destroy may be used on it). */
int oref_write_jwnt(void* hv, const char* path)
{
    Handle* h = (Handle*)hv;
    h->net->writeBinary(path);
    return 0;
}

void* oref_net_from_jwnt(const char* path, float lmScale, float insPenalty)
{
    Handle* h = new Handle;
    h->blockSize = 0; h->models = NULL; h->dec = NULL;
    h->net = new WFSTNetwork(lmScale, insPenalty);
    h->net->readBinary(path);
    return h;
}

void oref_net_dims(void* hv, int* out3)
{
    Handle* h = (Handle*)hv;
    out3[0] = h->net->getNumStates(); out3[1] = h->net->getNumTransitions(); out3[2] = h->net->getInitState();
}

/* dims[0..7] = vecSize nGMMs nHMMs nTransMats maxStates maxComps nNetStates nArcs ; dims[8]=initState */
void oref_dims(void* hv, int* dims)
{
    Handle* h = (Handle*)hv;
    dims[0] = h->models->dimVec();   dims[1] = h->models->dimGMMs();
    dims[2] = h->models->dimHMMs();  dims[3] = h->models->dimTMats();
    dims[4] = h->models->maxStates(); dims[5] = h->models->maxComps();
    dims[6] = h->net->getNumStates(); dims[7] = h->net->getNumTransitions();
    dims[8] = h->net->getInitState();
}

/* Reference's flattened model tables, for checking the product's host loader.
 * trP/se are per HMM, padded to S=maxStates: trP[h][i][j], se[h][j][0..1]. */
void oref_dump_models(void* hv, int* hmm_nstates, int* hmm_gmm, int* hmm_tmat, float* hmm_tee,
                      float* trP, int* se, int* gmm_ncomp, float* dets, float* means, float* ivars)
{
    Handle* h = (Handle*)hv; RefModels* m = h->models;
    const int S = m->maxStates(), C = m->maxComps(), D = m->dimVec();
    for (int i = 0; i < m->dimHMMs(); ++i) {
        const int n = m->getNumStates(i);
        hmm_nstates[i] = n;
        hmm_tmat[i] = m->hmm(i).transMatrixInd;
        hmm_tee[i] = m->getTeeLogProb(i);
        real** tp = m->getTransMat(i);
        SEIndex* s = m->getSEIndex(i);
        for (int a = 0; a < S; ++a) {
            hmm_gmm[i * S + a] = a < n ? m->hmm(i).gmmInds[a] : -1;
            for (int b = 0; b < S; ++b)
                trP[(i * S + a) * S + b] = (a < n && b < n) ? tp[a][b] : LOG_ZERO;
            se[(i * S + a) * 2 + 0] = (a >= 1 && a < n) ? s[a].start : 0;
            se[(i * S + a) * 2 + 1] = (a >= 1 && a < n) ? s[a].end : 0;
        }
    }
    for (int g = 0; g < m->dimGMMs(); ++g) {
        const int n = m->ncomp(g);
        gmm_ncomp[g] = n;
        for (int c = 0; c < C; ++c) {
            dets[g * C + c] = c < n ? m->det(g)[c] : LOG_ZERO;
            for (int d = 0; d < D; ++d) {
                means[(g * C + c) * D + d] = c < n ? m->mean(g)[c * m->stride() + d] : 0.f;
                ivars[(g * C + c) * D + d] = c < n ? m->ivar(g)[c * m->stride() + d] : 0.f;
            }
        }
    }
}

/* Reference's network as loaded (weights already negated/scaled, aux symbols rewritten),
 * through public getters only (WFSTNetwork.h:126-167). */
void oref_dump_net(void* hv, int* arc_to, float* arc_w, int* arc_in, int* arc_out,
                   int* st_first, int* st_n, float* st_final)
{
    Handle* h = (Handle*)hv; WFSTNetwork* n = h->net;
    for (int a = 0; a < n->getNumTransitions(); ++a) {
        WFSTTransition* t = n->getOneTransition(a);
        arc_to[a] = t->toState; arc_w[a] = t->weight; arc_in[a] = t->inLabel; arc_out[a] = t->outLabel;
    }
    for (int s = 0; s < n->getNumStates(); ++s) {
        st_n[s] = n->getNumTransitionsOfOneState(s);
        st_first[s] = st_n[s] > 0 ? n->getTransID(s, 0) : 0;
        st_final[s] = n->isFinalState(s) ? n->getFinalStateWeight(s) : LOG_ZERO;
    }
}

/* out[t * nGMM + g] = HTKFlatModels::calcOutput(g) at frame t (public, HTKFlatModels.h:38). */
void oref_gmm_scores(void* hv, const float* feats, int T, float* out)
{
    Handle* h = (Handle*)hv; RefModels* m = h->models;
    const int D = m->dimVec(), G = m->dimGMMs();
    std::vector<float*> ptr(T);
    for (int t = 0; t < T; ++t) ptr[t] = const_cast<float*>(feats) + (size_t)t * D;
    for (int t = 0; t < T; ++t) {
        m->newFrame(t, &ptr[t], std::min(20, T - t));
        for (int g = 0; g < G; ++g) out[(size_t)t * G + g] = m->calcOutput(g);
    }
}

/* Decode one utterance exactly the way DecoderSingleTest::decodeUtterance drives the decoder
 * (src/DecoderSingleTest.cpp:262-298): init, processFrame with min(20,remaining) look-ahead
 * pointers, finish.  Returns number of words, -1 if finish() returned NULL, -2 if the hyp is
 * inactive (score <= LOG_ZERO: final token without any word label, WFSTDecoderLite.cpp:273-306).
 * frame_cnt (optional, T x 6 ints): nActiveInsts nActiveEmit nActiveEnd nEmitProcessed nEndProcessed nPath
 * frame_best (optional, T floats): bestEmitScore after the frame. */
int oref_decode(void* hv, const float* feats, int T, OrefWord* words, int maxWords, float* totals,
                int* frame_cnt, float* frame_best, double* seconds)
{
    Handle* h = (Handle*)hv;
    const int D = h->models->dimVec();
    std::vector<float*> ptr(T);
    for (int t = 0; t < T; ++t) ptr[t] = const_cast<float*>(feats) + (size_t)t * D;

    auto t0 = std::chrono::steady_clock::now();
    h->dec->init();
    for (int t = 0; t < T; ++t) {
        h->dec->processFrame(&ptr[t], t, std::min(20, T - t));
        if (frame_cnt) {
            int* c = frame_cnt + (size_t)t * 6;
            c[0] = h->dec->cntActiveInsts(); c[1] = h->dec->cntActiveEmit(); c[2] = h->dec->cntActiveEnd();
            c[3] = h->dec->cntEmitProcessed(); c[4] = h->dec->cntEndProcessed(); c[5] = h->dec->cntPaths();
        }
        if (frame_best) frame_best[t] = h->dec->best();
    }
    DecHyp* hyp = h->dec->finish();
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();

    totals[0] = totals[1] = totals[2] = LOG_ZERO;
    if (hyp == NULL) return -1;
    totals[0] = hyp->score; totals[1] = hyp->acousticScore; totals[2] = hyp->lmScore;
    if (!(hyp->score > LOG_ZERO)) return -2;
    /* chain is newest first; emit oldest first */
    int n = 0;
    for (DecHypHist* p = hyp->hist; p; p = p->prev) ++n;
    int k = n;
    for (DecHypHist* p = hyp->hist; p; p = p->prev) {
        --k;
        if (k < maxWords) {
            words[k].label = p->state; words[k].time = p->time;
            words[k].score = p->score; words[k].ac = p->acousticScore; words[k].lm = p->lmScore;
        }
    }
    return n;
}

/* Streaming partial results: decode `feats`, and after every `every`-th frame ask the reference for the path records
 * all live hypotheses have converged on (partialPaths, src/WFSTDecoderLite.h:200).  Trace k (k < return value) was
 * taken after frame trace_frame[k] and found trace_len[k] records (-1: not traceable, see RefDecoder::partialTrace),
 * whose labels / frames are flat[off .. off + len) with off = sum of the earlier non-negative lengths. */
int oref_decode_partial(void* hv, const float* feats, int T, int every, int max_traces, int* trace_frame, int* trace_len,
                        int* flat_labels, int* flat_frames, int flat_cap)
{
    Handle* h = (Handle*)hv;
    const int D = h->models->dimVec();
    std::vector<float*> ptr(T);
    for (int t = 0; t < T; ++t) ptr[t] = const_cast<float*>(feats) + (size_t)t * D;
    h->dec->setPartialDecodeOptions(1 << 30);           /* the decoder keeps its list, but never traces on its own */
    h->dec->init();
    int n_tr = 0, off = 0;
    std::vector<int> labels, frames;
    for (int t = 0; t < T; ++t) {
        h->dec->processFrame(&ptr[t], t, std::min(20, T - t));
        if (every > 0 && (t + 1) % every == 0 && n_tr < max_traces) {
            const int n = h->dec->partialTrace(&labels, &frames);
            trace_frame[n_tr] = t;
            trace_len[n_tr] = n;
            if (n > 0) {
                if (off + n > flat_cap) break;
                for (int i = 0; i < n; ++i) { flat_labels[off + i] = labels[i]; flat_frames[off + i] = frames[i]; }
                off += n;
            }
            ++n_tr;
        }
    }
    h->dec->finish();
    h->dec->setPartialDecodeOptions(0);
    return n_tr;
}

} // extern "C"

/* Parse stage alone (oracle/shim/htkparse_rd.cpp behind the reference's `htkparse` symbol): 0 = accepted,
 * 1 = syntax error, 2 = a check of a grammar action failed (htkerror), -1 = cannot open.  Lets tests compare what
 * the two restatements of htkparse.l/.y accept without the reference's error() ending the process. */
#include "htkparse.h"
extern "C" int oref_mmf_parse_only(const char* path, int* counts5)
{
    FILE* fd = fopen(path, "rb");
    if (!fd) return -1;
    const int rc = htkparse((void*)fd);
    fclose(fd);
    if (rc == 0 && counts5) {
        counts5[0] = htk_def.n_hmms; counts5[1] = htk_def.n_sh_states; counts5[2] = htk_def.n_sh_transmats;
        counts5[3] = htk_def.n_mix_pools; counts5[4] = htk_def.global_opts.vec_size;
    }
    if (rc == 0) cleanHTKDef();      /* (after a failed parse the partial records are simply leaked: test process) */
    return rc;
}
