/* TEST INFRASTRUCTURE ONLY.
 * The reference's own harness around the decode path — DecoderBatchTest / DecoderSingleTest (src/DecoderBatchTest.cpp,
 * src/DecoderSingleTest.cpp: extended file names, the frame-feeding loop, result-word extraction, the five output
 * formats) and DecVocabulary — compiled UNMODIFIED against stand-ins for the two out-of-tree libraries they use
 * (Tracter's HTKSource / FrameSink behind the reference's own FrontEnd.h, Torch3's DiskXFile / EditDistance:
 * oracle/shim_harness/), driving the unmodified WFSTDecoderLite.  juicer_b200/harness.py is checked against the text
 * this writes (tests/test_harness.py).  What the shims assume about Tracter — an HTK parameter file is a 12-byte
 * big-endian header plus big-endian float32 frames, a [begin, end] range includes its end frame, time stamps are
 * 10 ms per frame — is the one part that is NOT pinned by the reference's code. */
#include <cstdio>
#include <cstdlib>

#include "DecoderBatchTest.h"
#include "WFSTDecoderLite.h"
#include "HTKFlatModels.h"
#include "WFSTNetwork.h"
#include "WFSTLattice.h"
#include "LogFile.h"

using namespace Juicer;

extern "C" {

/* format: 0 verbose, 1 trans, 2 ref, 3 mlf, 4 xmlf (DBTOutputFormat).  list_file: one (extended) feature file name per
 * line.  lexicon: "word phone ..." lines in alphabetical order (DecVocabulary), so that vocabulary index = output
 * label - 1.  Returns 0. */
int oref_harness_run(const char* jmbi, const char* fsm, const char* insyms, const char* outsyms, const char* lexicon,
                     const char* list_file, const char* out_file, int format, float start_beam, float main_beam,
                     float end_beam, float word_beam, int max_hyps, const char* sent_start, const char* sent_end,
                     int remove_sil, int frames_per_sec)
{
    HTKFlatModels* models = new HTKFlatModels;
    models->setBlockSize(5);
    models->readBinary(jmbi);
    WFSTNetwork* net = new WFSTNetwork(fsm, insyms, outsyms, 1.0, 0.0, REMOVEBOTH);
    IDecoder* dec = new WFSTDecoderLite(net, models, start_beam, main_beam, end_beam, word_beam, max_hyps);
    const int D = models->getInputVecSize();
    Tracter::HTKSource::sFrameSize() = D;
    DecVocabulary* vocab = new DecVocabulary(lexicon, '\0', (sent_start && sent_start[0]) ? sent_start : NULL,
                                             (sent_end && sent_end[0]) ? sent_end : NULL, NULL);
    FrontEnd* fe = new FrontEnd(D, FRONTEND_HTK);
    DecoderBatchTest* bt = new DecoderBatchTest(vocab, NULL, fe, dec, list_file, DST_FEATS_FACTORY, D, out_file,
                                                (DBTOutputFormat)format, NULL, remove_sil != 0, frames_per_sec);
    bt->run();
    delete bt;
    delete fe;
    delete vocab;
    delete dec;
    delete net;
    delete models;
    return 0;
}

} /* extern "C" */
