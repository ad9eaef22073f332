// host_loaders.cpp — host-side file readers feeding the device tables.
//
// These mirror, value for value, what the reference's own loaders leave in memory, so the
// CUDA path decodes the SAME numbers as WFSTDecoderLite would:
//   jgpu_load_fsm   <-> WFSTNetwork::WFSTNetwork(fsm, insyms, outsyms, scale, insPenalty,
//                       REMOVEBOTH)                         src/WFSTNetwork.cpp:371-616
//                       WFSTAlphabet::WFSTAlphabet(file)    src/WFSTNetwork.cpp:52-112
//                       removeAuxiliarySymbols(true)        src/WFSTNetwork.cpp:1421-1456
//   jgpu_load_jmbi  <-> HTKModels::readBinary               src/HTKModels.cpp:1112-1245
//                       createTrPandSEIndex                 src/HTKModels.cpp:2330-2390
//                       tee weight                          src/HTKModels.cpp:1358-1370
//                       HTKFlatModels::init                 src/HTKFlatModels.cpp:94-177
//   (jgpu_load_mmf, the MMF text reader, lives in host_mmf.cpp and ends in the same jgpu_finish_models)
// Host only: no CUDA calls here.
#include <cfloat>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#include "../../include/juicer_b200.h"
#include "jgpu_err.h"
#include "host_models.h"

std::string& jgpu_err_buf()
{
    static thread_local std::string buf;
    return buf;
}

namespace {

#define LZ (-FLT_MAX)

int io_fail(const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    jgpu_err_buf() = buf;
    return JGPU_E_IO;
}

struct Alphabet {
    int max_label = -1;
    int n_labels = 0;
    std::vector<char> present, aux;
    std::vector<std::string> names;
};

int read_alphabet(const char* fname, Alphabet* a)
{
    FILE* fd = fopen(fname, "rb");
    if (!fd) return io_fail("cannot open symbols file %s", fname);
    char line[10000], sym[10000];
    int id;
    while (fgets(line, sizeof(line), fd)) {
        if (sscanf(line, "%s %d", sym, &id) != 2) continue;
        if (id < 0) { fclose(fd); return io_fail("%s: negative symbol id", fname); }
        if (id > (1 << 28)) { fclose(fd); return io_fail("%s: symbol id %d too large", fname, id); }
        if (id >= (int)a->present.size()) {
            a->present.resize(id + 1000, 0);
            a->aux.resize(id + 1000, 0);
            a->names.resize(id + 1000);
        }
        if (a->present[id]) { fclose(fd); return io_fail("%s: duplicate symbol id %d", fname, id); }
        a->present[id] = 1;
        a->names[id] = sym;
        a->aux[id] = sym[0] == '#';                    // auxiliary symbols start with '#'
        a->n_labels++;
        if (id > a->max_label) a->max_label = id;
    }
    fclose(fd);
    return JGPU_OK;
}

template <typename T>
T* dup_vec(const std::vector<T>& v)
{
    T* p = (T*)malloc(sizeof(T) * (v.size() ? v.size() : 1));
    if (!v.empty()) memcpy(p, v.data(), sizeof(T) * v.size());
    return p;
}

bool rd(FILE* f, void* p, size_t sz, size_t n) { return fread(p, sz, n, f) == n; }

bool rd_tag_name(FILE* f, const char* tag)
{
    char id[5] = {0};
    if (!rd(f, id, 4, 1) || strcmp(id, tag) != 0) return false;
    int len;
    if (!rd(f, &len, 4, 1) || len < 0) return false;
    if (len > 0 && fseek(f, len, SEEK_CUR) != 0) return false;
    return true;
}

} // namespace

int jgpu_io_fail(const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    jgpu_err_buf() = buf;
    return JGPU_E_IO;
}

static int load_fsm_impl(const char* fsm, const char* insyms, const char* outsyms, float lm_scale,
                         float ins_penalty, JgpuNet* out)
{
    if (!fsm || !insyms || !outsyms || !out) return io_fail("jgpu_load_fsm: null argument (both symbol files are mandatory)");
    memset(out, 0, sizeof(*out));
    FILE* fd = fopen(fsm, "rb");
    if (!fd) return io_fail("cannot open network file %s", fsm);
    std::vector<int> to, in, ol, st_first, st_n, fin_id;
    std::vector<float> w, fin_w;
    int init_state = -1, max_state = -1, max_in = -1, max_out = -1;
    char line[10000];
    int from, t, i, o, fs;
    float weight;
    while (fgets(line, sizeof(line), fd)) {
        if (sscanf(line, "%d %d %d %d %f", &from, &t, &i, &o, &weight) != 5) {
            if (sscanf(line, "%d %d %d %d", &from, &t, &i, &o) != 4) {
                if (sscanf(line, "%d %f", &fs, &weight) != 2) {
                    if (sscanf(line, "%d", &fs) != 1) continue;
                    weight = 0.0;
                }
                fin_id.push_back(fs);
                fin_w.push_back((float)(-weight * lm_scale));
                continue;
            } else {
                weight = 0.0;
            }
        }
        if (from < 0 || t < 0 || i < 0 || o < 0) { fclose(fd); return io_fail("%s: negative field in arc line", fsm); }
        if (init_state < 0) init_state = from;          // source state of the first line
        if (from > max_state) max_state = from;
        if (t > max_state) max_state = t;
        float wt = (float)(-weight * lm_scale);          // FSM weights are -log
        if (o > 0) wt += ins_penalty;
        const int id = (int)to.size();
        to.push_back(t); in.push_back(i); ol.push_back(o); w.push_back(wt);
        if (i > max_in) max_in = i;
        if (o > max_out) max_out = o;
        if (max_state >= (int)st_n.size()) {
            st_n.resize((size_t)max_state * 2 + 1024, 0);
            st_first.resize(st_n.size(), 0);
        }
        // getTransitions(prev,&next) returns first arc + count (src/WFSTNetwork.cpp:709-721): a state whose arc lines
        // are interleaved with another state's would get an overlapping row — the reference then decodes a different
        // network without a word; here that is a load error
        if (st_n[from] > 0 && st_first[from] + st_n[from] != id) {
            fclose(fd);
            return io_fail("%s: arcs of state %d are not contiguous (arc line %d)", fsm, from, id + 1);
        }
        if (st_n[from]++ == 0) st_first[from] = id;
    }
    fclose(fd);
    if (init_state < 0) return io_fail("%s: no arcs", fsm);
    const int n_states = max_state + 1;
    st_n.resize(n_states);
    st_first.resize(n_states);
    std::vector<float> st_final(n_states, LZ);
    for (size_t k = 0; k < fin_id.size(); ++k) {
        // (a final state that no arc mentions is beyond maxState: the reference errors out as well, src/WFSTNetwork.cpp:560-566)
        if (fin_id[k] < 0 || fin_id[k] > max_state) return io_fail("%s: final state %d out of range", fsm, fin_id[k]);
        st_final[fin_id[k]] = fin_w[k];
    }
    // arcs of one state must be contiguous for the reference's getTransitions(prev,&next)
    for (int s = 0; s < n_states; ++s)
        if (st_n[s] > 0 && st_first[s] + st_n[s] > (int)to.size())
            return io_fail("%s: arcs of state %d are not contiguous", fsm, s);

    Alphabet ia, oa;
    int rc;
    if ((rc = read_alphabet(insyms, &ia))) return rc;
    if ((rc = read_alphabet(outsyms, &oa))) return rc;
    if (max_in > ia.max_label) return io_fail("maxInLab > inputAlphabet->getMaxLabel()");
    if (max_out > oa.max_label) return io_fail("maxOutLab=%d > outputAlphabet->getMaxLabel()=%d", max_out, oa.max_label);
    max_in = ia.max_label;
    max_out = oa.max_label;
    int word_end_marker = max_in + 1;
    if (word_end_marker <= max_out) word_end_marker = max_out + 1;
    for (size_t a = 0; a < to.size(); ++a) {             // removeAuxiliarySymbols(markAux = true)
        if (!ia.present[in[a]]) return io_fail("input label %d has no symbol-table entry", in[a]);
        if (!oa.present[ol[a]]) return io_fail("output label %d has no symbol-table entry", ol[a]);
        if (ia.aux[in[a]]) in[a] = word_end_marker;
        if (oa.aux[ol[a]]) ol[a] = word_end_marker;
    }
    for (int k = 0; k < ia.n_labels; ++k)                // the reference calls getLabel(i) for i < nLabels
        if (k > ia.max_label || !ia.present[k]) return io_fail("input symbol table has a hole at id %d", k);

    out->n_states = n_states;
    out->n_arcs = (int)to.size();
    out->init_state = init_state;
    out->arc_to = dup_vec(to);
    out->arc_weight = dup_vec(w);
    out->arc_in = dup_vec(in);
    out->arc_out = dup_vec(ol);
    out->state_first = dup_vec(st_first);
    out->state_narcs = dup_vec(st_n);
    out->state_final = dup_vec(st_final);
    return JGPU_OK;
}

// JWNT binary network: WFSTNetwork::readBinary (src/WFSTNetwork.cpp:1228-1370) with the alphabets of
// WFSTAlphabet::readBinary (:250-297).  Layout (native endianness, int = i32, real = f32, bool = 1 byte):
//   "JWNT" initState maxState nStates maxOutTransitions wordEndMarker silMarker spMarker
//   per state 0..maxState : label finalInd nTrans trans[nTrans]      (trans = ids into the transition array)
//   nFinalStates, then (id, weight) each                               (weights as stored: already scaled)
//   nTransitions, then (id, toState, weight, inLabel, outLabel) each   (weights unscaled, without insPenalty)
//   bool haveInputAlphabet [alphabet]  bool haveOutputAlphabet [alphabet]  "JWNT"
//   alphabet = "JWAL" maxLabel nLabels, per label 0..maxLabel: len [len bytes, nul-terminated], nAux, isAux[maxLabel+1]
// After reading, the reference multiplies every transition weight by the scaling factor and adds the insertion
// penalty to arcs with an output label (:1351-1365); final weights are used as read.
static bool rd(FILE* fd, void* p, size_t n) { return fread(p, 1, n, fd) == n; }

// size of an open file: every count read from a binary header is checked against it before anything is sized
// from it, so a corrupt file is an I/O error and not a std::bad_alloc across the C ABI
static long long file_size(FILE* fd)
{
    const long pos = ftell(fd);
    if (pos < 0 || fseek(fd, 0, SEEK_END) != 0) return -1;
    const long end = ftell(fd);
    fseek(fd, pos, SEEK_SET);
    return (long long)end;
}

static int skip_alphabet(FILE* fd, const char* fname)
{
    char id[5] = {0, 0, 0, 0, 0};
    int max_label, n_labels;
    if (!rd(fd, id, 4) || strcmp(id, "JWAL") != 0) return io_fail("%s: WFSTAlphabet::readBinary - invalid ID", fname);
    if (!rd(fd, &max_label, 4) || !rd(fd, &n_labels, 4) || (long long)max_label * 4 > file_size(fd)) return io_fail("%s: error reading alphabet header", fname);
    if (max_label >= 0) {
        for (int i = 0; i <= max_label; ++i) {
            int len;
            if (!rd(fd, &len, 4) || len < 0 || len > (1 << 20)) return io_fail("%s: error reading label length", fname);
            if (len > 0) {
                std::vector<char> buf(len);
                if (!rd(fd, buf.data(), len)) return io_fail("%s: error reading label string", fname);
                if (buf[len - 1] != '\0') return io_fail("%s: last char of label string was not nul", fname);
            }
        }
        int n_aux;
        std::vector<char> is_aux(max_label + 1);
        if (!rd(fd, &n_aux, 4) || !rd(fd, is_aux.data(), max_label + 1)) return io_fail("%s: error reading isAux array", fname);
    }
    return JGPU_OK;
}

static int load_jwnt_impl(const char* path, float lm_scale, float ins_penalty, JgpuNet* out)
{
    if (!path || !out) return io_fail("jgpu_load_jwnt: null argument");
    memset(out, 0, sizeof(*out));
    FILE* fd = fopen(path, "rb");
    if (!fd) return io_fail("WFSTNetwork::readBinary - error opening input file %s", path);
    struct Closer { FILE* f; ~Closer() { fclose(f); } } closer{fd};
    char id[5] = {0, 0, 0, 0, 0};
    int hdr[7];   // initState maxState nStates maxOutTransitions wordEndMarker silMarker spMarker
    if (!rd(fd, id, 4) || strcmp(id, "JWNT") != 0) return io_fail("%s: WFSTNetwork::readBinary - invalid ID", path);
    if (!rd(fd, hdr, sizeof(hdr))) return io_fail("%s: error reading header", path);
    const int init_state = hdr[0], max_state = hdr[1];
    if (max_state < 0 || init_state < 0 || init_state > max_state) return io_fail("%s: bad initState / maxState", path);
    const long long fsize = file_size(fd);
    if (fsize < 0 || (long long)max_state * 12 > fsize) return io_fail("%s: maxState %d does not fit the file size", path, max_state);
    const int n_states = max_state + 1;
    std::vector<int> st_first(n_states, 0), st_n(n_states, 0), st_fin(n_states, -1);
    std::vector<std::vector<int>> st_trans(n_states);
    for (int s = 0; s < n_states; ++s) {
        int rec[3];   // label finalInd nTrans
        if (!rd(fd, rec, sizeof(rec)) || rec[2] < 0 || (long long)rec[2] * 4 > fsize) return io_fail("%s: error reading states[%d]", path, s);
        st_fin[s] = rec[1];
        st_n[s] = rec[2];
        if (rec[2] > 0) {
            st_trans[s].resize(rec[2]);
            if (!rd(fd, st_trans[s].data(), (size_t)rec[2] * 4)) return io_fail("%s: error reading states[%d].trans array", path, s);
            st_first[s] = st_trans[s][0];                 // getTransitions(prev, &next): first arc + count (:709-721)
        }
    }
    int n_final;
    if (!rd(fd, &n_final, 4) || n_final < 0 || (long long)n_final * 8 > fsize) return io_fail("%s: error reading nFinalStates", path);
    std::vector<int> fin_id(n_final);
    std::vector<float> fin_w(n_final);
    for (int i = 0; i < n_final; ++i)
        if (!rd(fd, &fin_id[i], 4) || !rd(fd, &fin_w[i], 4)) return io_fail("%s: error reading finalStates", path);
    int n_arcs;
    if (!rd(fd, &n_arcs, 4) || n_arcs < 0 || (long long)n_arcs * 20 > fsize) return io_fail("%s: error reading nTransitions", path);
    std::vector<int> to(n_arcs), in(n_arcs), ol(n_arcs);
    std::vector<float> w(n_arcs);
    for (int a = 0; a < n_arcs; ++a) {
        int tid;
        if (!rd(fd, &tid, 4) || !rd(fd, &to[a], 4) || !rd(fd, &w[a], 4) || !rd(fd, &in[a], 4) || !rd(fd, &ol[a], 4))
            return io_fail("%s: error reading transitions", path);
        if (to[a] < 0 || to[a] > max_state || in[a] < 0 || ol[a] < 0) return io_fail("%s: transition %d out of range", path, a);
    }
    for (int k = 0; k < 2; ++k) {
        unsigned char have;
        if (!rd(fd, &have, 1)) return io_fail("%s: error reading haveAlphabet", path);
        if (have) { int rc = skip_alphabet(fd, path); if (rc) return rc; }
    }
    char id2[5] = {0, 0, 0, 0, 0};
    if (!rd(fd, id2, 4) || strcmp(id2, "JWNT") != 0) return io_fail("%s: WFSTNetwork::readBinary - invalid ID (2)", path);

    if (lm_scale != 1.0f)
        for (int a = 0; a < n_arcs; ++a) w[a] *= lm_scale;
    if (ins_penalty != 0.0f)
        for (int a = 0; a < n_arcs; ++a)
            if (ol[a] > 0) w[a] += ins_penalty;
    std::vector<float> st_final(n_states, LZ);
    for (int s = 0; s < n_states; ++s) {
        if (st_fin[s] >= 0) {
            if (st_fin[s] >= n_final) return io_fail("%s: state %d: finalInd out of range", path, s);
            st_final[s] = fin_w[st_fin[s]];
        }
        if (st_n[s] > 0) {
            if (st_first[s] < 0 || st_first[s] + st_n[s] > n_arcs) return io_fail("%s: arcs of state %d out of range", path, s);
            for (int k = 0; k < st_n[s]; ++k)             // the decoder takes nTrans arcs from the first one on
                if (st_trans[s][k] != st_first[s] + k) return io_fail("%s: arcs of state %d are not contiguous", path, s);
        }
    }
    out->n_states = n_states;
    out->n_arcs = n_arcs;
    out->init_state = init_state;
    out->arc_to = dup_vec(to);
    out->arc_weight = dup_vec(w);
    out->arc_in = dup_vec(in);
    out->arc_out = dup_vec(ol);
    out->state_first = dup_vec(st_first);
    out->state_narcs = dup_vec(st_n);
    out->state_final = dup_vec(st_final);
    return JGPU_OK;
}

extern "C" int jgpu_free_net(JgpuNet* n)
{
    if (!n) return JGPU_OK;
    free((void*)n->arc_to); free((void*)n->arc_weight); free((void*)n->arc_in); free((void*)n->arc_out);
    free((void*)n->state_first); free((void*)n->state_narcs); free((void*)n->state_final);
    memset(n, 0, sizeof(*n));
    return JGPU_OK;
}

static int load_jmbi_impl(const char* path, JgpuHmm* hmm, JgpuGmm* gmm)
{
    if (!path || !hmm || !gmm) return io_fail("jgpu_load_jmbi: null argument");
    memset(hmm, 0, sizeof(*hmm));
    memset(gmm, 0, sizeof(*gmm));
    FILE* f = fopen(path, "rb");
    if (!f) return io_fail("cannot open model file %s", path);
#define BAD(msg) do { fclose(f); return io_fail("%s: %s", path, msg); } while (0)
    char id[5] = {0};
    int hdr[7];
    if (!rd(f, id, 4, 1) || strcmp(id, "JMBI") != 0) BAD("not a JMBI file");
    if (!rd(f, hdr, 4, 7)) BAD("truncated header");
    const int D = hdr[0], nMean = hdr[1], nVar = hdr[2], nMix = hdr[3], nGMM = hdr[4], nTM = hdr[5], nHMM = hdr[6];
    if (D <= 0 || nMean < 0 || nVar < 0 || nMix < 0 || nGMM < 0 || nTM < 0 || nHMM < 0) BAD("bad counts");
    {
        // smallest possible record of each kind (tag + nameLen + payload) against the file size
        const long long fsize = file_size(f);
        const long long least = (long long)nMean * (8 + 4ll * D) + (long long)nVar * (12 + 8ll * D) + (long long)nMix * 12 +
                                (long long)nGMM * 16 + (long long)nTM * 12 + (long long)nHMM * 16;
        if (fsize < 0 || D > (1 << 20) || least > fsize) BAD("record counts do not fit the file size");
    }
    RawModels m;
    m.D = D;
    m.means.resize((size_t)nMean * D); m.vars.resize((size_t)nVar * D); m.gconst.resize(nVar);
    std::vector<float> skip(D);
    for (int i = 0; i < nMean; ++i) {
        if (!rd_tag_name(f, "JMMN") || !rd(f, &m.means[(size_t)i * D], 4, D)) BAD("bad mean vector record");
    }
    for (int i = 0; i < nVar; ++i) {
        if (!rd_tag_name(f, "JMVR") || !rd(f, &m.vars[(size_t)i * D], 4, D) || !rd(f, skip.data(), 4, D) ||
            !rd(f, &m.gconst[i], 4, 1))
            BAD("bad variance vector record");
    }
    m.mix_mean.resize(nMix); m.mix_var.resize(nMix); m.mix_name.resize(nMix);
    for (int i = 0; i < nMix; ++i) {
        int nc;
        if (!rd_tag_name(f, "JMMX") || !rd(f, &nc, 4, 1) || nc < 0 || nc > nMean) BAD("bad mixture record");
        m.mix_mean[i].resize(nc);
        m.mix_var[i].resize(nc);
        if (!rd(f, m.mix_mean[i].data(), 4, nc) || !rd(f, m.mix_var[i].data(), 4, nc)) BAD("bad mixture record");
        for (int c = 0; c < nc; ++c)
            if (m.mix_mean[i][c] < 0 || m.mix_mean[i][c] >= nMean || m.mix_var[i][c] < 0 || m.mix_var[i][c] >= nVar) BAD("mixture index out of range");
    }
    m.gmm_mix.resize(nGMM); m.gmm_logw.resize(nGMM); m.gmm_name.resize(nGMM);
    for (int i = 0; i < nGMM; ++i) {
        int nc;
        if (!rd_tag_name(f, "JMGM") || !rd(f, &m.gmm_mix[i], 4, 1) || !rd(f, &nc, 4, 1) || nc < 0 || nc > nMean) BAD("bad GMM record");
        std::vector<float> wts(nc);
        m.gmm_logw[i].resize(nc);
        if (!rd(f, wts.data(), 4, nc) || !rd(f, m.gmm_logw[i].data(), 4, nc)) BAD("bad GMM record");
        if (m.gmm_mix[i] < 0 || m.gmm_mix[i] >= nMix) BAD("GMM mixture index out of range");
    }
    m.tms.resize(nTM);
    for (int i = 0; i < nTM; ++i) {
        RawTransMat& t = m.tms[i];
        if (!rd_tag_name(f, "JMTM") || !rd(f, &t.n, 4, 1) || t.n <= 0 || t.n > 64) BAD("bad transition matrix record");
        std::vector<int> nsucs(t.n);
        if (!rd(f, nsucs.data(), 4, t.n)) BAD("bad transition matrix record");
        int total = 0;
        for (int s = 0; s < t.n; ++s) { if (nsucs[s] < 0 || nsucs[s] > t.n) BAD("bad successor count"); total += nsucs[s]; }
        std::vector<int> sucs(total);
        std::vector<float> probs(total), logp(total);
        if (!rd(f, sucs.data(), 4, total) || !rd(f, probs.data(), 4, total) || !rd(f, logp.data(), 4, total)) BAD("bad transition matrix record");
        t.sucs.resize(t.n);
        t.logp.resize(t.n);
        int k = 0;
        for (int s = 0; s < t.n; ++s)
            for (int j = 0; j < nsucs[s]; ++j, ++k) {
                if (sucs[k] < 0 || sucs[k] >= t.n) BAD("successor out of range");
                t.sucs[s].push_back(sucs[k]);
                t.logp[s].push_back(logp[k]);
            }
    }
    m.hmm_n.resize(nHMM); m.hmm_tm.resize(nHMM); m.hmm_g.resize(nHMM);
    for (int i = 0; i < nHMM; ++i) {
        if (!rd_tag_name(f, "JMHM") || !rd(f, &m.hmm_n[i], 4, 1) || m.hmm_n[i] <= 0 || m.hmm_n[i] > 64) BAD("bad HMM record");
        m.hmm_g[i].resize(m.hmm_n[i]);
        if (!rd(f, m.hmm_g[i].data(), 4, m.hmm_n[i]) || !rd(f, &m.hmm_tm[i], 4, 1)) BAD("bad HMM record");
        if (m.hmm_tm[i] < 0 || m.hmm_tm[i] >= nTM || m.tms[m.hmm_tm[i]].n != m.hmm_n[i]) BAD("HMM / transition matrix mismatch");
    }
    unsigned char hybrid = 0;
    if (!rd(f, &hybrid, 1, 1)) BAD("missing hybridMode flag");
    fclose(f);
#undef BAD
    if (hybrid) return io_fail("%s: hybrid (ANN posterior) models are outside the GMM decode path", path);
    return jgpu_finish_models(m, hmm, gmm);
}

// Flat tables from the loaded records: trP / SEIndex (createTrPandSEIndex, src/HTKModels.cpp:2330-2390), tee
// weight (:582-593 text load, :1358-1370 binary load) and the flat GMM parameters (HTKFlatModels::init,
// src/HTKFlatModels.cpp:94-177).  Both load paths of the reference end in exactly these steps.
int jgpu_finish_models(const RawModels& m, JgpuHmm* hmm, JgpuGmm* gmm)
{
    const int D = m.D, nMix = (int)m.mix_mean.size(), nGMM = (int)m.gmm_mix.size(), nHMM = (int)m.hmm_n.size();
    const std::vector<RawTransMat>& tms = m.tms;
    const std::vector<int>&hmm_n = m.hmm_n, &hmm_tm = m.hmm_tm, &gmm_mix = m.gmm_mix;
    const std::vector<std::vector<int>>&hmm_g = m.hmm_g, &mix_mean = m.mix_mean, &mix_var = m.mix_var;
    const std::vector<std::vector<float>>& gmm_logw = m.gmm_logw;
    const std::vector<float>&means = m.means, &vars = m.vars, &gconst = m.gconst;
    int maxS = 0, maxC = 0;
    for (int i = 0; i < nHMM; ++i) if (hmm_n[i] > maxS) maxS = hmm_n[i];
    for (int i = 0; i < nMix; ++i) if ((int)mix_mean[i].size() > maxC) maxC = (int)mix_mean[i].size();

    // ---- HMM view: trP / SEIndex / tee (src/HTKModels.cpp:2330-2390, :1358-1370) ----
    const int S = maxS;
    std::vector<int> o_n(hmm_n), o_gmm((size_t)nHMM * S, -1), o_se((size_t)nHMM * S * 2, 0);
    std::vector<float> o_trp((size_t)nHMM * S * S, LZ), o_tee(nHMM, LZ);
    for (int i = 0; i < nHMM; ++i) {
        const RawTransMat& t = tms[hmm_tm[i]];
        const int n = t.n;
        std::vector<float> trp((size_t)n * n, LZ);
        for (int j = 0; j < n; ++j)
            for (size_t k = 0; k < t.sucs[j].size(); ++k) trp[(size_t)j * n + t.sucs[j][k]] = t.logp[j][k];
        for (int a = 0; a < n; ++a) {
            o_gmm[(size_t)i * S + a] = hmm_g[i][a];
            for (int b = 0; b < n; ++b) o_trp[((size_t)i * S + a) * S + b] = trp[(size_t)a * n + b];
        }
        for (int j = 1; j < n; ++j) {
            int mn, mx;
            for (mn = (j == n - 1 ? 1 : 0); mn < n - 1; ++mn)
                if (trp[(size_t)mn * n + j] > LZ) break;
            for (mx = n - 1; mx >= 1; --mx)
                if (trp[(size_t)mx * n + j] > LZ) break;
            o_se[((size_t)i * S + j) * 2 + 0] = mn;
            o_se[((size_t)i * S + j) * 2 + 1] = mx + 1;
        }
        for (size_t k = 1; k < t.sucs[0].size(); ++k)     // tee: entry -> exit as successor index >= 1
            if (t.sucs[0][k] == n - 1) {
                if (o_tee[i] != LZ) return io_fail("more than one tee transition in HMM %d", i);
                o_tee[i] = t.logp[0][k];
            }
    }
    // ---- flat GMM parameters (src/HTKFlatModels.cpp:94-177) ----
    if (nGMM > nMix) return io_fail("HTKFlatModels needs one mixture per GMM (nGMMs=%d > nMixtures=%d)", nGMM, nMix);
    const int C = maxC > 0 ? maxC : 1;
    std::vector<int> o_nc(nGMM);
    std::vector<float> o_det((size_t)nGMM * C, LZ), o_mu((size_t)nGMM * C * D, 0.0f), o_iv((size_t)nGMM * C * D, 0.0f);
    for (int g = 0; g < nGMM; ++g) {
        // flat parameters of slot g come from mixture g (:143-165); log weights from GMM g's
        // own mixture (:169-176) — identical when gMMs[g].mixtureInd == g, which the reference assumes
        const int mi = g;
        const int nc = (int)mix_mean[mi].size();
        if ((int)gmm_logw[g].size() < (int)mix_mean[gmm_mix[g]].size()) return io_fail("GMM %d: weight count mismatch", g);
        o_nc[g] = nc;
        for (int c = 0; c < nc; ++c) {
            const float* mo = &means[(size_t)mix_mean[mi][c] * D];
            const float* vo = &vars[(size_t)mix_var[mi][c] * D];
            for (int k = 0; k < D; ++k) {
                o_mu[((size_t)g * C + c) * D + k] = mo[k];
                o_iv[((size_t)g * C + c) * D + k] = (float)(1.0 / vo[k]);    // :160
            }
            o_det[(size_t)g * C + c] = gconst[mix_var[mi][c]];                // :163
        }
        const int nc2 = (int)mix_mean[gmm_mix[g]].size();
        for (int c = 0; c < nc2 && c < C; ++c) o_det[(size_t)g * C + c] += gmm_logw[g][c];   // :174
    }
    hmm->n_hmms = nHMM; hmm->max_states = S;
    hmm->n_states = dup_vec(o_n); hmm->gmm = dup_vec(o_gmm); hmm->trp = dup_vec(o_trp);
    hmm->se = dup_vec(o_se); hmm->tee = dup_vec(o_tee);
    gmm->n_gmms = nGMM; gmm->dim = D; gmm->max_comps = C;
    gmm->n_comps = dup_vec(o_nc); gmm->dets = dup_vec(o_det); gmm->means = dup_vec(o_mu); gmm->ivars = dup_vec(o_iv);
    return JGPU_OK;
}

extern "C" int jgpu_free_models(JgpuHmm* hmm, JgpuGmm* gmm)
{
    if (hmm) {
        free((void*)hmm->n_states); free((void*)hmm->gmm); free((void*)hmm->trp); free((void*)hmm->se); free((void*)hmm->tee);
        memset(hmm, 0, sizeof(*hmm));
    }
    if (gmm) {
        free((void*)gmm->n_comps); free((void*)gmm->dets); free((void*)gmm->means); free((void*)gmm->ivars);
        memset(gmm, 0, sizeof(*gmm));
    }
    return JGPU_OK;
}

// ---- C ABI wrappers: nothing may unwind across the boundary (a corrupt file must be JGPU_E_IO, not std::terminate) ----
#define JG_GUARDED(call)                                                                  \
    try { return call; }                                                                  \
    catch (const std::exception& e) { return io_fail("host loader: %s", e.what()); }      \
    catch (...) { return io_fail("host loader: unknown exception"); }

extern "C" int jgpu_load_fsm(const char* fsm, const char* insyms, const char* outsyms, float lm_scale, float ins_penalty, JgpuNet* out)
{
    JG_GUARDED(load_fsm_impl(fsm, insyms, outsyms, lm_scale, ins_penalty, out))
}
extern "C" int jgpu_load_jwnt(const char* path, float lm_scale, float ins_penalty, JgpuNet* out)
{
    JG_GUARDED(load_jwnt_impl(path, lm_scale, ins_penalty, out))
}
extern "C" int jgpu_load_jmbi(const char* path, JgpuHmm* hmm, JgpuGmm* gmm)
{
    JG_GUARDED(load_jmbi_impl(path, hmm, gmm))
}
