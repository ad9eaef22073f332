// GpuWFSTDecoder.cpp — see GpuWFSTDecoder.h.
#include "GpuWFSTDecoder.h"
#include "LogFile.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <log_add.h>

using namespace Torch;

namespace Juicer
{
    namespace
    {
        // Protected members of HTKFlatModels / HTKModels, reached through pointers-to-member that a
        // derived class may legally form (src/HTKFlatModels.h:43-63, src/HTKModels.h:139-170).
        struct FlatPeek : public HTKFlatModels
        {
            static FMixture* mixtures(HTKFlatModels* m) { return m->*(&FlatPeek::fMixtures); }
            static real* dets(HTKFlatModels* m) { return m->*(&FlatPeek::fDets); }
            static real* means(HTKFlatModels* m) { return m->*(&FlatPeek::fMeans); }
            static real* ivars(HTKFlatModels* m) { return m->*(&FlatPeek::fVars); }
            static int stride(HTKFlatModels* m) { return m->*(&FlatPeek::fvecSize4); }
            static int nGmms(HTKFlatModels* m) { return m->*(&FlatPeek::nGMMs); }
            static HMM* hmms(HTKFlatModels* m) { return m->*(&FlatPeek::hMMs); }
            static bool hybrid(HTKFlatModels* m) { return m->*(&FlatPeek::hybridMode); }
        };

        const int kBlockFrames = 16;         // frames buffered per device push (<= the harness look-ahead of 20)
    }

    GpuWFSTDecoder::GpuWFSTDecoder(
        WFSTNetwork* network , IModels* models ,
        real phoneStartPruneWin , real emitPruneWin , real phoneEndPruneWin , real wordPruneWin ,
        int maxEmitHyps
    )
    {
        handle = NULL;
        bestDecHyp = NULL;
        nextFrame = 0;
        nPending = 0;

        HTKFlatModels* flat = dynamic_cast<HTKFlatModels*>(models);
        if (flat == NULL)
            error("GpuWFSTDecoder - models must be HTKFlatModels (the flat diagonal-GMM scorer)");
        if (FlatPeek::hybrid(flat))
            error("GpuWFSTDecoder - hybrid (ANN posterior) models are not on the GMM decode path");

        // ---- network through public getters (src/WFSTNetwork.h:126-167) ----
        const int nStates = network->getNumStates();
        const int nArcs = network->getNumTransitions();
        std::vector<int32_t> arcTo(nArcs), arcIn(nArcs), arcOut(nArcs), stFirst(nStates, 0), stN(nStates, 0);
        std::vector<float> arcW(nArcs), stFinal(nStates, LOG_ZERO);
        for (int a = 0; a < nArcs; ++a) {
            WFSTTransition* t = network->getOneTransition(a);
            arcTo[a] = t->toState; arcW[a] = t->weight; arcIn[a] = t->inLabel; arcOut[a] = t->outLabel;
        }
        for (int s = 0; s < nStates; ++s) {
            stN[s] = network->getNumTransitionsOfOneState(s);
            if (stN[s] > 0) stFirst[s] = network->getTransID(s, 0);   // getTransitions(prev,&next): first arc + count
            if (network->isFinalState(s)) stFinal[s] = network->getFinalStateWeight(s);
        }
        JgpuNet net;
        net.n_states = nStates; net.n_arcs = nArcs; net.init_state = network->getInitState();
        net.arc_to = arcTo.data(); net.arc_weight = arcW.data(); net.arc_in = arcIn.data(); net.arc_out = arcOut.data();
        net.state_first = stFirst.data(); net.state_narcs = stN.data(); net.state_final = stFinal.data();

        // ---- HMM topology through IModels (src/Models.h:51-66) ----
        const int nHMMs = models->getNumHMMs();
        int S = 0;
        for (int h = 0; h < nHMMs; ++h) if (models->getNumStates(h) > S) S = models->getNumStates(h);
        std::vector<int32_t> hN(nHMMs), hGmm((size_t)nHMMs * S, -1), hSe((size_t)nHMMs * S * 2, 0);
        std::vector<float> hTrp((size_t)nHMMs * S * S, LOG_ZERO), hTee(nHMMs);
        HMM* hmms = FlatPeek::hmms(flat);
        for (int h = 0; h < nHMMs; ++h) {
            const int n = models->getNumStates(h);
            real** trP = models->getTransMat(h);
            SEIndex* se = models->getSEIndex(h);
            hN[h] = n;
            hTee[h] = models->getTeeLogProb(h);
            for (int i = 0; i < n; ++i) {
                hGmm[(size_t)h * S + i] = hmms[h].gmmInds[i];
                for (int j = 0; j < n; ++j) hTrp[((size_t)h * S + i) * S + j] = trP[i][j];
                if (i >= 1) { hSe[((size_t)h * S + i) * 2] = se[i].start; hSe[((size_t)h * S + i) * 2 + 1] = se[i].end; }
            }
        }
        JgpuHmm hm;
        hm.n_hmms = nHMMs; hm.max_states = S; hm.n_states = hN.data(); hm.gmm = hGmm.data();
        hm.trp = hTrp.data(); hm.se = hSe.data(); hm.tee = hTee.data();

        // ---- flat GMM parameters (src/HTKFlatModels.cpp:94-177) ----
        vecSize = models->getInputVecSize();
        const int nG = FlatPeek::nGmms(flat), stride = FlatPeek::stride(flat);
        FMixture* fm = FlatPeek::mixtures(flat);
        int C = 1;
        for (int g = 0; g < nG; ++g) if (fm[g].compNum > C) C = fm[g].compNum;
        std::vector<int32_t> gN(nG);
        std::vector<float> gDet((size_t)nG * C, LOG_ZERO), gMu((size_t)nG * C * vecSize, 0.0f), gIv((size_t)nG * C * vecSize, 0.0f);
        for (int g = 0; g < nG; ++g) {
            gN[g] = fm[g].compNum;
            for (int c = 0; c < fm[g].compNum; ++c) {
                const int ci = fm[g].compInd + c;
                gDet[(size_t)g * C + c] = FlatPeek::dets(flat)[ci];
                memcpy(&gMu[((size_t)g * C + c) * vecSize], FlatPeek::means(flat) + (size_t)ci * stride, sizeof(float) * vecSize);
                memcpy(&gIv[((size_t)g * C + c) * vecSize], FlatPeek::ivars(flat) + (size_t)ci * stride, sizeof(float) * vecSize);
            }
        }
        JgpuGmm gm;
        gm.n_gmms = nG; gm.dim = vecSize; gm.max_comps = C; gm.n_comps = gN.data();
        gm.dets = gDet.data(); gm.means = gMu.data(); gm.ivars = gIv.data();

        JgpuCfg cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.start_beam = phoneStartPruneWin; cfg.main_beam = emitPruneWin;
        cfg.end_beam = phoneEndPruneWin; cfg.word_beam = wordPruneWin; cfg.max_hyps = maxEmitHyps;
        cfg.n_lanes = 1;
        const char* dev = getenv("JUICER_B200_DEVICE");
        cfg.device = dev ? atoi(dev) : 0;
        if (jgpu_create(&net, &hm, &gm, &cfg, &handle) != JGPU_OK)
            error("GpuWFSTDecoder - %s", jgpu_last_error());        // Torch error(): message + exit, like the reference
        const char* pti = getenv("PartialTraceInterval");        // src/WFSTDecoderLite.cpp:117
        partialTraceInterval = 0;
        lastPartialTraceFrame = -1;
        setPartialDecodeOptions(pti ? atoi(pti) : 0);
        pending.resize((size_t)kBlockFrames * vecSize);
        words.resize(65536);       // the device side has no per-utterance word limit; this is the host buffer
    }

    GpuWFSTDecoder::~GpuWFSTDecoder()
    {
        delete bestDecHyp;
        jgpu_destroy(handle);
    }

    void GpuWFSTDecoder::setPartialDecodeOptions(int traceInterval)
    {
        partialTraceInterval = traceInterval > 0 ? traceInterval : 0;
        LogFile::printf("WFSTDecoderLite::partialTraceInterval = %d frames\n", partialTraceInterval);   // :893-897
    }

    // the words every live hypothesis has in its history (tracePartialPath, src/WFSTDecoderLite.cpp:822-871)
    void GpuWFSTDecoder::tracePartial()
    {
        JgpuResult res;
        memset(&res, 0, sizeof(res));
        res.max_words = (int)words.size();
        res.words = words.data();
        if (jgpu_partial_result(handle, 0, &res) != JGPU_OK) error("GpuWFSTDecoder::tracePartial - %s", jgpu_last_error());
        if (res.status > 0) partialWords.assign(words.begin(), words.begin() + (res.status < res.max_words ? res.status : res.max_words));
        lastPartialTraceFrame = nextFrame - 1;
    }

    void GpuWFSTDecoder::init()
    {
        partialWords.clear();
        lastPartialTraceFrame = -1;
        delete bestDecHyp;                   // result of the last utterance stays valid until here
        bestDecHyp = NULL;
        hist.clear();
        nextFrame = 0;
        nPending = 0;
        if (jgpu_utt_begin(handle, 0) != JGPU_OK) error("GpuWFSTDecoder::init - %s", jgpu_last_error());
    }

    void GpuWFSTDecoder::flush()
    {
        if (nPending == 0) return;
        if (jgpu_push_frames(handle, 0, pending.data(), nPending) != JGPU_OK)
            error("GpuWFSTDecoder::processFrame - %s", jgpu_last_error());
        nPending = 0;
    }

    // inputVec[0] is frame currFrame_; the look-ahead pointers inputVec[1..nFrames) are not needed
    // (the reference only uses them for its GMM block cache, src/HTKFlatModels.cpp:226-262).
    void GpuWFSTDecoder::processFrame(real** inputVec, int currFrame_, int nFrames)
    {
        if (currFrame_ != nextFrame)
            error("GpuWFSTDecoder::processFrame - invalid frame");   // HTKFlatModels::newFrame :296-297
        memcpy(&pending[(size_t)nPending * vecSize], inputVec[0], sizeof(float) * vecSize);
        ++nPending;
        ++nextFrame;
        if (nPending == kBlockFrames) {
            flush();                                                 // launches asynchronously
            if (partialTraceInterval > 0 && nextFrame - 1 - lastPartialTraceFrame > partialTraceInterval) tracePartial();
        }
    }

    DecHyp* GpuWFSTDecoder::finish()
    {
        flush();
        JgpuResult res;
        memset(&res, 0, sizeof(res));
        res.max_words = (int)words.size();
        res.words = words.data();
        if (jgpu_utt_end(handle, 0, &res) != JGPU_OK) error("GpuWFSTDecoder::finish - %s", jgpu_last_error());
        if (res.status <= -10) error("GpuWFSTDecoder::finish - device arena overflow (status %d)", res.status);
        if (partialTraceInterval > 0) {                              // :246-257: the last trace runs from the best token
            if (res.status > 0) partialWords.assign(words.begin(), words.begin() + (res.status < res.max_words ? res.status : res.max_words));
            LogFile::printf("Partial paths recovered at frames: ");
            for (size_t k = 0; k < partialWords.size(); ++k) LogFile::printf("%03d ", partialWords[k].time);
            LogFile::printf("\n");
        }
        if (res.status == -1) {
            fprintf(stderr, "WARNING: no token survived at the end of decoding\n");   // src/WFSTDecoderLite.cpp:264-267
            return NULL;
        }
        bestDecHyp = new DecHyp();
        if (res.status == -2) return bestDecHyp;                     // :273-306, inactive hypothesis
        const int n = res.status < res.max_words ? res.status : res.max_words;
        hist.resize(n);
        for (int k = 0; k < n; ++k) {                                // words[] is oldest first; chain is newest first
            DecHypHist& hh = hist[k];
            hh.type = DHHTYPE;
            hh.nConnect = 1;
            hh.prev = k > 0 ? &hist[k - 1] : NULL;
            hh.state = words[k].label;
            hh.time = words[k].time;
            hh.score = words[k].score;
            hh.acousticScore = words[k].ac;
            hh.lmScore = words[k].lm;
        }
        bestDecHyp->score = res.score;
        bestDecHyp->lmScore = res.lm;
        bestDecHyp->acousticScore = res.ac;
        bestDecHyp->hist = n > 0 ? &hist[n - 1] : NULL;
        return bestDecHyp;
    }

    JgpuStats GpuWFSTDecoder::stats()
    {
        JgpuStats s;
        memset(&s, 0, sizeof(s));
        jgpu_stats(handle, 0, &s);
        return s;
    }
}
