// GpuWFSTDecoder.h — drop-in Juicer::IDecoder backed by the juicer_b200 C ABI.
//
// Compiled INSIDE a Juicer build (it includes Juicer's own headers): same constructor signature
// as WFSTDecoderLite (src/WFSTDecoderLite.h:81-89), same init / processFrame / finish contract
// (src/Decoder.h:18-30), same DecHyp / DecHypHist result chain (src/DecHypHistPool.h:38-49,
// 146-165).  Replace `new WFSTDecoderLite(...)` at src/juicer.cpp:584-586 with
// `new GpuWFSTDecoder(...)`; see INTEGRATION.md.
//
// The network is flattened through WFSTNetwork's public getters only and is never written to
// (the reference stores per-decoder state in WFSTTransition::hook; this decoder does not).
// HMM topology comes through IModels' public getters; the flat GMM parameters are protected
// members of HTKFlatModels and are read through pointers-to-member of a derived class.
#ifndef JUICER_B200_GPUWFSTDECODER_H
#define JUICER_B200_GPUWFSTDECODER_H

#include <vector>

#include "Decoder.h"
#include "HTKFlatModels.h"
#include "WFSTNetwork.h"

#include "juicer_b200.h"

namespace Juicer
{
    class GpuWFSTDecoder : public IDecoder
    {
    public:
        GpuWFSTDecoder(
            WFSTNetwork* network_ ,
            IModels* models_ ,
            real phoneStartPruneWin_ ,
            real emitPruneWin_ ,
            real phoneEndPruneWin_ ,
            real wordPruneWin_ ,
            int maxEmitHyps_
        );
        virtual ~GpuWFSTDecoder();

        bool modelLevelOutput() { return false ; }
        WFSTLattice* getLattice() { return 0 ; }
        void init();
        void processFrame(real** inputVec, int currFrame_, int nFrames);
        DecHyp* finish();

        // work counters of the last utterance (reference prints them at src/WFSTDecoderLite.cpp:231-241)
        JgpuStats stats();

        // Streaming partial results (PARTIAL_DECODING, src/WFSTDecoderLite.h:196-206): same switch as the reference —
        // setPartialDecodeOptions(n) or the environment variable PartialTraceInterval.  With n > 0 the decoder asks the
        // device every n frames for the words all live hypotheses agree on (partialResult() returns the latest answer)
        // and finish() prints "Partial paths recovered at frames: ..." like src/WFSTDecoderLite.cpp:247-257.
        void setPartialDecodeOptions(int traceInterval);
        const std::vector<JgpuWord>& partialResult() const { return partialWords; }

    private:
        jgpu_handle* handle;
        int vecSize;
        int nextFrame;
        std::vector<float> pending;          // frames accepted but not yet pushed to the device
        int nPending;
        std::vector<JgpuWord> words;
        std::vector<DecHypHist> hist;        // result chain storage, valid until the next init()
        DecHyp* bestDecHyp;

        void flush();
        int partialTraceInterval, lastPartialTraceFrame;
        std::vector<JgpuWord> partialWords;  // converged prefix at the last trace
        void tracePartial();
    };
}

#endif
