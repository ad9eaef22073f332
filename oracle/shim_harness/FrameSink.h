/* TEST INFRASTRUCTURE ONLY.  Stand-in for Tracter's FrameSink<T>: random access to the frames of the source it is
 * connected to; Read() returns 0 past the end; time stamps are frame indices. */
#ifndef ORACLE_SHIM_HARNESS_FRAMESINK_H
#define ORACLE_SHIM_HARNESS_FRAMESINK_H
#include "HTKSource.h"
namespace Tracter {
struct FrameInfo { int size; };
template <class T> class FrameSink {
public:
    explicit FrameSink(Component<T>* src) : mStore(dynamic_cast<FrameStore*>(src)) {}
    FrameInfo Frame() const { FrameInfo f; f.size = mStore ? mStore->size : 0; return f; }
    void Reset() {}
    const T* Read(IndexType i) { return (mStore && i >= 0 && i < mStore->count()) ? &mStore->frames[(size_t)i * mStore->size] : 0; }
    TimeType TimeStamp(IndexType i) const { return (TimeType)i * ORACLE_FRAME_PERIOD_NS; }
private:
    FrameStore* mStore;
};
}
#endif
