/* TEST INFRASTRUCTURE ONLY.  Stand-in for Tracter's HTKSource (out of tree): an HTK parameter file — 12-byte big-endian
 * header {nSamples i32, sampPeriod i32, sampSize i16, parmKind i16}, then big-endian float32 frames — opened for the frame
 * time range [begin, end] (nanoseconds, -1 = unbounded; FrameSink::TimeStamp(i) = i x 10 ms, the inverse of what Open does).  The frame
 * size every source reports before a file is open (FrontEnd's constructor asserts on it) is HTKSource::sFrameSize, set by
 * the driver. */
#ifndef ORACLE_SHIM_HARNESS_HTKSOURCE_H
#define ORACLE_SHIM_HARNESS_HTKSOURCE_H
#include <cstdio>
#include <cstring>
#include <vector>
#include <stdint.h>
#include "TracterObject.h"
namespace Tracter {
class FrameStore : public Component<float> {
public:
    std::vector<float> frames;     /* the opened range, host byte order */
    int size;                      /* floats per frame */
    FrameStore() : size(0) {}
    long count() const { return size ? (long)(frames.size() / size) : 0; }
};
class HTKSource : public FrameStore, public ISource {
public:
    static int& sFrameSize() { static int s = 39; return s; }
    long mBegin;                   /* first frame of the opened range in the file */
    HTKSource() : mBegin(0) { size = sFrameSize(); }
    void Open(const char* name, TimeType begin = -1, TimeType end = -1)
    {
        frames.clear();
        FILE* f = fopen(name, "rb");
        if (!f) { fprintf(stderr, "HTKSource: cannot open %s\n", name); exit(-1); }
        unsigned char h[12];
        if (fread(h, 1, 12, f) != 12) { fprintf(stderr, "HTKSource: short header in %s\n", name); exit(-1); }
        const long n = (long)((uint32_t)h[0] << 24 | (uint32_t)h[1] << 16 | (uint32_t)h[2] << 8 | h[3]);
        const int samp = (int)((unsigned)h[8] << 8 | h[9]);
        if (samp != size * 4) { fprintf(stderr, "HTKSource: %s has %d-byte frames, expected %d\n", name, samp, size * 4); exit(-1); }
        const long bf = begin >= 0 ? (long)(begin / ORACLE_FRAME_PERIOD_NS) : 0, ef = end >= 0 ? (long)(end / ORACLE_FRAME_PERIOD_NS) : n - 1;
        const long b = bf, e = ef < n ? ef : n - 1;
        mBegin = b;
        std::vector<unsigned char> raw((size_t)samp);
        fseek(f, 12 + b * samp, SEEK_SET);
        for (long t = b; t <= e; ++t) {
            if (fread(raw.data(), 1, samp, f) != (size_t)samp) break;
            for (int i = 0; i < size; ++i) {
                const uint32_t u = (uint32_t)raw[4 * i] << 24 | (uint32_t)raw[4 * i + 1] << 16 | (uint32_t)raw[4 * i + 2] << 8 | raw[4 * i + 3];
                float x; memcpy(&x, &u, 4);
                frames.push_back(x);
            }
        }
        fclose(f);
    }
};
}
#endif
