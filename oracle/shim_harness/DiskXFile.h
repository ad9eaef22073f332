/* TEST INFRASTRUCTURE ONLY.  Stand-in for Torch3's DiskXFile: DecoderBatchTest only wraps its output FILE* in one to hand
 * it to EditDistance::print (src/DecoderBatchTest.cpp:157), which is reached only with expected results. */
#ifndef ORACLE_SHIM_HARNESS_DISKXFILE_H
#define ORACLE_SHIM_HARNESS_DISKXFILE_H
#include <cstdio>
namespace Torch {
class XFile { public: virtual ~XFile() {} };
class DiskXFile : public XFile {
public:
    FILE* file;
    explicit DiskXFile(FILE* f) : file(f) {}
};
}
using Torch::XFile;
using Torch::DiskXFile;
#endif
