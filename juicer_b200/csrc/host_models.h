// host_models.h — the acoustic-model records as the reference holds them after loading
// (HTKModels' meanVecs / varVecs / mixtures / gMMs / transMats / hMMs, src/HTKModels.h:36-110),
// shared by the JMBI reader (host_loaders.cpp) and the MMF text reader (host_mmf.cpp).
// jgpu_finish_models turns them into the flat tables of the C ABI.  Host only.
#pragma once
#include <string>
#include <vector>

#include "../../include/juicer_b200.h"

struct RawTransMat {
    int n = 0;                                   // nStates
    std::vector<std::vector<int>>   sucs;        // successors of each state, in column order
    std::vector<std::vector<float>> logp;        // log transition probability of each successor
    std::string name;                            // "" = not shared
};

struct RawModels {
    int D = 0;                                   // vecSize
    std::vector<float> means, vars, gconst;      // [nMean][D], [nVar][D], [nVar] (sumLogVarPlusNObsLog2Pi)
    std::vector<std::vector<int>> mix_mean, mix_var;   // per mixture: mean / variance vector index of each component
    std::vector<std::string> mix_name;
    std::vector<int> gmm_mix;                    // per GMM: mixture index
    std::vector<std::vector<float>> gmm_logw;    // per GMM: log component weights
    std::vector<std::string> gmm_name;
    std::vector<RawTransMat> tms;
    std::vector<int> hmm_n, hmm_tm;              // per HMM: nStates, transition matrix index
    std::vector<std::vector<int>> hmm_g;         // per HMM: GMM index of each state (-1 for entry / exit)
};

int jgpu_io_fail(const char* fmt, ...);
int jgpu_finish_models(const RawModels& m, JgpuHmm* hmm, JgpuGmm* gmm);
