/* TEST INFRASTRUCTURE ONLY.  Drop-in check of the C++ host adapter against the reference,
 * through the reference's OWN interfaces: both decoders are held as Juicer::IDecoder* built
 * from the same WFSTNetwork* / IModels*, and driven exactly like
 * DecoderSingleTest::decodeUtterance (src/DecoderSingleTest.cpp:262-298).
 *   usage: dropin_test models.jmbi net.fsm net.insyms net.outsyms feats.f32 mainBeam [endBeam wordBeam startBeam maxHyps]
 * feats.f32 = raw float32 rows of vecSize.  Prints both word chains; exit 0 iff identical. */
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "WFSTDecoderLite.h"
#include "HTKFlatModels.h"
#include "WFSTNetwork.h"
#include "LogFile.h"
#include "GpuWFSTDecoder.h"
#include <log_add.h>

using namespace Juicer;

static DecHyp* decode(IDecoder* dec, std::vector<float>& feats, int T, int D)
{
    std::vector<float*> ptr(T);
    for (int t = 0; t < T; ++t) ptr[t] = &feats[(size_t)t * D];
    dec->init();
    for (int t = 0; t < T; ++t) dec->processFrame(&ptr[t], t, std::min(20, T - t));
    return dec->finish();
}

static int chain(DecHyp* h, std::vector<DecHypHist>& out)
{
    if (h == NULL) return -1;
    if (!(h->score > LOG_ZERO)) return -2;    /* DecHypHistPool::isActiveHyp, src/DecHypHistPool.cpp:239-250 */
    for (DecHypHist* p = h->hist; p; p = p->prev) {
        if (p->type != DHHTYPE) return -3;
        out.push_back(*p);
    }
    return (int)out.size();
}

int main(int argc, char** argv)
{
    if (argc < 7) { fprintf(stderr, "usage: see source\n"); return 2; }
    const float mainBeam = atof(argv[6]);
    const float endBeam = argc > 7 ? atof(argv[7]) : 0.f, wordBeam = argc > 8 ? atof(argv[8]) : 0.f;
    const float startBeam = argc > 9 ? atof(argv[9]) : 0.f;
    const int maxHyps = argc > 10 ? atoi(argv[10]) : 0;
    /* PARTIAL_DECODING: both decoders read the environment variable in their constructors (src/WFSTDecoderLite.cpp:117)
     * and print "Partial paths recovered at frames: ..." through LogFile at finish() (:247-257) */
    if (getenv("DROPIN_PARTIAL")) { setenv("PartialTraceInterval", getenv("DROPIN_PARTIAL"), 1); LogFile::open("stdout"); }
    HTKFlatModels* models = new HTKFlatModels;
    models->setBlockSize(5);
    models->readBinary(argv[1]);
    WFSTNetwork* net = new WFSTNetwork(argv[2], argv[3], argv[4], 1.0, 0.0, REMOVEBOTH);
    const int D = models->getInputVecSize();
    FILE* f = fopen(argv[5], "rb");
    if (!f) { perror(argv[5]); return 2; }
    std::vector<float> feats;
    float buf[4096];
    size_t n;
    while ((n = fread(buf, sizeof(float), 4096, f)) > 0) feats.insert(feats.end(), buf, buf + n);
    fclose(f);
    const int T = (int)(feats.size() / D);

    IDecoder* gpu = new GpuWFSTDecoder(net, models, startBeam, mainBeam, endBeam, wordBeam, maxHyps);
    IDecoder* ref = new WFSTDecoderLite(net, models, startBeam, mainBeam, endBeam, wordBeam, maxHyps);
    int bad = 0;
    for (int rep = 0; rep < 2; ++rep) {          /* twice: results must survive re-use of both decoders */
        std::vector<DecHypHist> a, b;
        DecHyp* hg = decode(gpu, feats, T, D);
        const int ng = chain(hg, a);
        const float gs = hg ? hg->score : 0, ga = hg ? hg->acousticScore : 0, gl = hg ? hg->lmScore : 0;
        DecHyp* hr = decode(ref, feats, T, D);
        const int nr = chain(hr, b);
        printf("rep %d: T=%d  reference %d words, gpu %d words\n", rep, T, nr, ng);
        if (ng != nr) { ++bad; continue; }
        if (nr > 0 && (gs != hr->score || ga != hr->acousticScore || gl != hr->lmScore)) {
            printf("  totals differ: gpu %.6f %.6f %.6f  ref %.6f %.6f %.6f\n", gs, ga, gl, hr->score, hr->acousticScore, hr->lmScore);
            ++bad;
        }
        for (int k = 0; k < nr; ++k) {
            const bool same = a[k].state == b[k].state && a[k].time == b[k].time && a[k].score == b[k].score &&
                              a[k].acousticScore == b[k].acousticScore && a[k].lmScore == b[k].lmScore;
            if (!same || rep == 0)
                printf("  %s word %d: label %d/%d time %d/%d score %.6f/%.6f ac %.4f/%.4f lm %.6f/%.6f\n", same ? "ok " : "BAD",
                       nr - 1 - k, a[k].state, b[k].state, a[k].time, b[k].time, a[k].score, b[k].score,
                       a[k].acousticScore, b[k].acousticScore, a[k].lmScore, b[k].lmScore);
            if (!same) ++bad;
        }
    }
    printf("%s\n", bad ? "DROPIN FAIL" : "DROPIN PASS");
    delete gpu;
    delete ref;
    return bad ? 1 : 0;
}
