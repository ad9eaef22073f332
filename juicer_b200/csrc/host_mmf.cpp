// host_mmf.cpp — HTK MMF (text model definition) reader: jgpu_load_mmf.
//
// Mirrors what HTKFlatModels::Load leaves in memory (src/HTKFlatModels.cpp:89-92):
//   tokens     src/htkparse.l.lpp:20-268   (flex rules: longest match, earlier rule wins a tie,
//                                           unknown characters are dropped one at a time)
//   grammar    src/htkparse.y.ypp:76-700   (macros ~o ~h ~s ~t ~m ~v, the checks of each action)
//   semantics  HTKModels::initFromHTKParseResult / addHMM / addGMM / addMixture / addMeanVec /
//              addVarVec / addTransMatrix        src/HTKModels.cpp:397-444, 519-974
//   then the steps shared with the JMBI reader (jgpu_finish_models, host_loaders.cpp).
// The reference's parser is bison/flex output; this is a hand-written scanner and recursive
// descent over the same token classes and productions.  Arithmetic follows the reference's
// C expressions operand type by operand type (float members updated with double right-hand
// sides narrow after every statement), so the tables are bit-identical given the same libm.
// Host only: no CUDA calls here.
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/juicer_b200.h"
#include "host_models.h"

namespace {

#define LZ (-FLT_MAX)
const double kLog2Pi = 1.83787706640934548356;    // Torch3 log_add.h LOG_2_PI (src/HTKModels.cpp:859)

// ------------------------------------------------------------------------------------------
// scanner (src/htkparse.l.lpp)
enum Tok {
    T_EOF, T_INTEGER, T_REAL, T_QSTRING, T_STRING, T_BEGINHMM, T_ENDHMM, T_NUMSTATES, T_STATE, T_NUMMIXES,
    T_MIXTURE, T_MEAN, T_VARIANCE, T_GCONST, T_TRANSP, T_HMMSETID, T_TMIX, T_VECSIZE, T_STREAMINFO,
    T_COVKIND, T_DURKIND, T_PARMKIND, T_HMACRO, T_SMACRO, T_MMACRO, T_TMACRO, T_VMACRO, T_OMACRO
};

inline bool is_w(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n'; }
inline bool is_d(char c) { return c >= '0' && c <= '9'; }
inline bool is_strch(char c)      // STR  [a-zA-Z0-9\+\-\$\^#@_\&\[\]:]   (:24)
{
    return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || is_d(c) || c == '+' || c == '-' || c == '$' ||
           c == '^' || c == '#' || c == '@' || c == '_' || c == '&' || c == '[' || c == ']' || c == ':';
}

struct Scanner {
    const char* s;
    size_t n, p = 0;
    int ival = 0; float fval = 0; std::string sval;

    size_t len_int(size_t q) const         // INT  -?{D}+   (:22)
    {
        size_t r = q;
        if (r < n && s[r] == '-') ++r;
        size_t d = r;
        while (r < n && is_d(s[r])) ++r;
        return r > d ? r - q : 0;
    }
    size_t len_re(size_t q) const          // RE  (\+|-)?({D}+\.?{D}*|\.{D}+)([eE](\+|-)?{D}+)?   (:23)
    {
        size_t r = q;
        if (r < n && (s[r] == '+' || s[r] == '-')) ++r;
        size_t d = r;
        while (r < n && is_d(s[r])) ++r;
        if (r > d) {
            if (r < n && s[r] == '.') { ++r; while (r < n && is_d(s[r])) ++r; }
        } else {
            if (!(r < n && s[r] == '.')) return 0;
            size_t f = ++r;
            while (r < n && is_d(s[r])) ++r;
            if (r == f) return 0;
        }
        if (r < n && (s[r] == 'e' || s[r] == 'E')) {
            size_t e = r + 1;
            if (e < n && (s[e] == '+' || s[e] == '-')) ++e;
            size_t f = e;
            while (e < n && is_d(s[e])) ++e;
            if (e > f) r = e;
        }
        return r - q;
    }
    size_t len_str(size_t q) const { size_t r = q; while (r < n && is_strch(s[r])) ++r; return r - q; }
    size_t len_qstr(size_t q) const        // QSTR  \"{STR}\"
    {
        if (!(q < n && s[q] == '"')) return 0;
        const size_t l = len_str(q + 1);
        if (l == 0 || !(q + 1 + l < n && s[q + 1 + l] == '"')) return 0;
        return l + 2;
    }
    size_t skip_w(size_t q) const { while (q < n && is_w(s[q])) ++q; return q; }

    // "<" one of three spellings ">" ; returns the position after '>' or 0
    size_t kw(size_t q, const char* a, const char* b, const char* c) const
    {
        const char* alt[3] = {a, b, c};
        for (const char* k : alt) {
            if (!k) continue;
            const size_t l = strlen(k);
            if (q + 1 + l < n && s[q] == '<' && memcmp(s + q + 1, k, l) == 0 && s[q + 1 + l] == '>') return q + l + 2;
        }
        return 0;
    }
    // <Keyword>{W}*{INT}
    bool kw_int(const char* a, const char* b, const char* c)
    {
        size_t q = kw(p, a, b, c);
        if (!q) return false;
        q = skip_w(q);
        const size_t l = len_int(q);
        if (!l) return false;
        ival = atoi(std::string(s + q, l).c_str());
        p = q + l;
        return true;
    }
    // <Keyword>{W}*({STR}|{QSTR})
    bool kw_name(const char* a, const char* b, const char* c)
    {
        size_t q = kw(p, a, b, c);
        if (!q) return false;
        q = skip_w(q);
        size_t l = len_str(q);
        if (l) { sval.assign(s + q, l); p = q + l; return true; }
        l = len_qstr(q);
        if (l) { sval.assign(s + q + 1, l - 2); p = q + l; return true; }
        return false;
    }
    bool parmkind()                        // \<({PkBase}|{PKBASE}|{pkbase})({PKQUAL}|{pkqual})*\>   (:25-29, :196)
    {
        static const char* base[] = {"Discrete", "LPCepstra", "FBank", "MelSpec", "LPRefC", "User",
                                     "DISCRETE", "LPCEPSTRA", "LPC", "MFCC", "PLP", "FBANK", "MELSPEC", "LPREFC", "USER",
                                     "discrete", "lpcepstra", "lpc", "mfcc", "plp", "fbank", "melspec", "lprefc", "user"};
        for (const char* b : base) {
            const size_t l = strlen(b);
            if (p + 1 + l >= n || memcmp(s + p + 1, b, l) != 0) continue;
            size_t q = p + 1 + l;
            while (q + 1 < n && s[q] == '_' && strchr("DATENZOVCK0datenzovck", s[q + 1]) && s[q + 1] != '\0') q += 2;
            if (q < n && s[q] == '>') { sval.assign(s + p + 1, q - p - 1); p = q + 1; return true; }
        }
        return false;
    }
    bool macro(char which)                 // ~x{W}+{QSTR}
    {
        if (!(p + 2 < n && s[p] == '~' && s[p + 1] == which && is_w(s[p + 2]))) return false;
        const size_t q = skip_w(p + 2);
        const size_t l = len_qstr(q);
        if (!l) return false;
        sval.assign(s + q + 1, l - 2);
        p = q + l;
        return true;
    }

    Tok next()
    {
        for (;;) {
            if (p >= n) return T_EOF;
            const char c = s[p];
            if (is_w(c)) { ++p; continue; }
            const size_t li = len_int(p), lr = len_re(p), ls = len_str(p);
            if (li || lr || ls) {
                if (li >= lr && li >= ls) { ival = atoi(std::string(s + p, li).c_str()); p += li; return T_INTEGER; }
                if (lr >= ls) { fval = (float)atof(std::string(s + p, lr).c_str()); p += lr; return T_REAL; }
                sval.assign(s + p, ls); p += ls; return T_STRING;
            }
            if (c == '"') {
                const size_t l = len_qstr(p);
                if (l) { sval.assign(s + p + 1, l - 2); p += l; return T_QSTRING; }
            } else if (c == '<') {
                size_t q;
                if ((q = kw(p, "BeginHMM", "BEGINHMM", "beginhmm"))) { p = q; return T_BEGINHMM; }
                if ((q = kw(p, "EndHMM", "ENDHMM", "endhmm"))) { p = q; return T_ENDHMM; }
                if (kw_int("NumStates", "NUMSTATES", "numstates")) return T_NUMSTATES;
                if (kw_int("State", "STATE", "state")) return T_STATE;
                if (kw_int("NumMixes", "NUMMIXES", "nummixes")) return T_NUMMIXES;
                if (kw_int("Mixture", "MIXTURE", "mixture")) return T_MIXTURE;
                if (kw_int("Mean", "MEAN", "mean")) return T_MEAN;
                if (kw_int("Variance", "VARIANCE", "variance")) return T_VARIANCE;
                if ((q = kw(p, "GConst", "GCONST", "gconst"))) {        // {W}*({RE}|{INT})
                    q = skip_w(q);
                    size_t l = len_re(q);
                    const size_t l2 = len_int(q);
                    if (l2 > l) l = l2;
                    if (l) { fval = (float)atof(std::string(s + q, l).c_str()); p = q + l; return T_GCONST; }
                }
                if (kw_int("TransP", "TRANSP", "transp")) return T_TRANSP;
                if (kw_name("HmmSetId", "HMMSETID", "hmmsetid")) return T_HMMSETID;
                if (kw_name("TMix", "TMIX", "tmix")) return T_TMIX;
                if (kw_int("VecSize", "VECSIZE", "vecsize")) return T_VECSIZE;
                if (kw_int("StreamInfo", "STREAMINFO", "streaminfo")) return T_STREAMINFO;
                if ((q = kw(p, "DiagC", "DIAGC", "diagc"))) { p = q; ival = 0; return T_COVKIND; }
                if ((q = kw(p, "InvDiagC", "INVDIAGC", "invdiagc"))) { p = q; ival = 1; return T_COVKIND; }
                if ((q = kw(p, "FullC", "FULLC", "fullc"))) { p = q; ival = 2; return T_COVKIND; }
                if ((q = kw(p, "LLTC", "lltc", nullptr))) { p = q; ival = 3; return T_COVKIND; }
                if ((q = kw(p, "XFormC", "XFORMC", "xformc"))) { p = q; ival = 4; return T_COVKIND; }
                if ((q = kw(p, "NullD", "NULLD", "nulld"))) { p = q; ival = 0; return T_DURKIND; }
                if ((q = kw(p, "PoissonD", "POISSOND", "poissond"))) { p = q; ival = 1; return T_DURKIND; }
                if ((q = kw(p, "GammaD", "GAMMAD", "gammad"))) { p = q; ival = 2; return T_DURKIND; }
                if ((q = kw(p, "GenD", "GEND", "gend"))) { p = q; ival = 3; return T_DURKIND; }
                if (parmkind()) return T_PARMKIND;
            } else if (c == '~') {
                if (macro('h')) return T_HMACRO;
                if (macro('s')) return T_SMACRO;
                if (macro('m')) return T_MMACRO;
                if (macro('t')) return T_TMACRO;
                if (macro('v')) return T_VMACRO;
                if (p + 2 < n && s[p + 1] == 'o' && is_w(s[p + 2])) { p = skip_w(p + 2); return T_OMACRO; }   // ~o{W}+
            }
            ++p;                                               // the catch-all rule drops the character (:268)
        }
    }
};

// ------------------------------------------------------------------------------------------
// parse result (src/htkparse.h:77-150)
struct PMix { int id = 0; float weight = 0; std::vector<float> means, vars; float gconst = 0; };
struct PMixList { std::vector<PMix> mixes; int pool_ind = -1; std::vector<float> weights; int n_mixes = 0; };
struct PState { std::string sh_name; bool named = false; int id = -1; PMixList ml; };
struct PTransMat { std::string sh_name; bool named = false; int n_states = 0; std::vector<float> transp; };
struct PHmm { std::string name; int n_states = 0; std::vector<PState> states; PTransMat tm; };
struct PPool { std::string name; std::vector<PMix> mixes; };

struct ParseError { std::string msg; };

struct Parser {
    Scanner sc;
    Tok tok = T_EOF;
    // global options (:181-270)
    std::string hmm_set_id, parm_kind; bool has_set_id = false, has_parm_kind = false;
    int n_streams = 0, vec_size = 0, cov_kind = -1, dur_kind = -1;
    std::vector<int> stream_widths;
    std::vector<PTransMat> sh_transmats;
    std::vector<PState> sh_states;
    std::vector<PPool> pools;
    std::vector<PHmm> hmms;

    [[noreturn]] void fail(const std::string& m) { throw ParseError{m}; }
    void advance() { tok = sc.next(); }
    void syntax(const char* where) { fail(std::string("syntax error in ") + where + " near byte " + std::to_string(sc.p)); }

    std::vector<float> rvector()           // rvector : (INTEGER | REAL)+   (:640-672)
    {
        std::vector<float> v;
        if (tok != T_INTEGER && tok != T_REAL) syntax("vector");
        while (tok == T_INTEGER || tok == T_REAL) {
            v.push_back(tok == T_INTEGER ? (float)sc.ival : sc.fval);
            advance();
        }
        return v;
    }
    std::vector<float> sized_vec(Tok which, const char* name)    // meanvec / variancevec (:560-588)
    {
        if (tok != which) syntax(name);
        const int n = sc.ival;
        if (n != vec_size) fail(std::string("HTKPARSE:") + name + " - value did not match global vec size");
        advance();
        std::vector<float> v = rvector();
        if ((int)v.size() != n) fail(std::string("HTKPARSE:") + name + " - n_elems did not match value");
        return v;
    }
    PMix mixpdf()                          // mixpdf : meanvec variancevec gconst   (:545-558, :589-597)
    {
        PMix m;
        m.means = sized_vec(T_MEAN, "meanvec");
        m.vars = sized_vec(T_VARIANCE, "variancevec");
        m.gconst = 0.0f;
        if (tok == T_GCONST) { m.gconst = sc.fval; advance(); }
        return m;
    }
    PMix mixturedef()                      // mixturedef : MIXTURE REAL mixpdf | mixpdf   (:531-544)
    {
        if (tok == T_MIXTURE) {
            const int id = sc.ival;
            advance();
            if (tok != T_REAL) syntax("mixturedef (the weight must be written as a real number)");
            const float w = sc.fval;
            advance();
            PMix m = mixpdf();
            m.id = id; m.weight = w;
            return m;
        }
        PMix m = mixpdf();
        m.id = 1; m.weight = 1.0f;
        return m;
    }
    PMixList mixtures()                    // mixtures : TMIX rvector | mixturelist   (:480-530)
    {
        PMixList l;
        if (tok == T_TMIX) {
            const std::string pool = sc.sval;
            advance();
            std::vector<float> w = rvector();
            size_t i = 0;
            for (; i < pools.size(); ++i) if (pools[i].name == pool) break;
            if (i >= pools.size()) fail("HTKPARSE:mixtures - TMIX string did not match the name of a mix pool");
            if (w.size() != pools[i].mixes.size()) fail("HTKPARSE:mixtures - tmixweights n_elems did not match n_mixes in mix pool");
            l.n_mixes = (int)w.size(); l.pool_ind = (int)i; l.weights = w;
            if (tok == T_MIXTURE || tok == T_MEAN) fail("mixture definitions after <TMix> weights are not supported");
            return l;
        }
        if (tok != T_MIXTURE && tok != T_MEAN) syntax("mixtures");
        while (tok == T_MIXTURE || tok == T_MEAN) l.mixes.push_back(mixturedef());
        l.n_mixes = (int)l.mixes.size();
        return l;
    }
    PTransMat transp()                     // transp : TRANSP rvector   (:620-638)
    {
        if (tok != T_TRANSP) syntax("transp");
        PTransMat t;
        const int n = sc.ival;
        advance();
        std::vector<float> v = rvector();
        if (n <= 0 || n != (int)v.size() / n) fail("HTKPARSE:transp - vec n_elems did not match TRANSP value");
        t.n_states = n;
        t.transp.assign(v.begin(), v.begin() + (size_t)n * n);
        return t;
    }
    void options()                         // options : option+   (:176-270)
    {
        bool any = false;
        for (;; any = true) {
            if (tok == T_HMMSETID) {
                if (has_set_id) { if (hmm_set_id != sc.sval) fail("HTKPARSE:option - hmm_set_id mismatch"); }
                else { hmm_set_id = sc.sval; has_set_id = true; }
                advance();
            } else if (tok == T_STREAMINFO) {
                const int ns = sc.ival;
                advance();
                std::vector<int> iv;
                if (tok != T_INTEGER) syntax("ivector");
                while (tok == T_INTEGER) { iv.push_back(sc.ival); advance(); }
                if ((int)iv.size() != ns) fail("HTKPARSE:option - STREAMINFO value does not match ivec size");
                if (vec_size > 0) {
                    int sum = 0;
                    for (int x : iv) sum += x;
                    if (sum != vec_size) fail("HTKPARSE:option - sum of stream widths does not equal vec_size");
                }
                if (n_streams > 0) { if (n_streams != ns) fail("HTKPARSE:option - n_streams mismatch"); }
                else { n_streams = ns; stream_widths = iv; }
            } else if (tok == T_VECSIZE) {
                const int vs = sc.ival;
                if (n_streams > 0) {
                    int sum = 0;
                    for (int x : stream_widths) sum += x;
                    if (sum != vs) fail("HTKPARSE:option - sum of stream widths does not equal NEW vec_size");
                }
                if (vec_size > 0) { if (vec_size != vs) fail("HTKPARSE:option - vec_size mismatch"); }
                else vec_size = vs;
                advance();
            } else if (tok == T_COVKIND) {
                if (cov_kind != -1) { if (cov_kind != sc.ival) fail("HTKPARSE:option - cov_kind mismatch"); }
                else cov_kind = sc.ival;
                advance();
            } else if (tok == T_DURKIND) {
                if (dur_kind != -1) { if (dur_kind != sc.ival) fail("HTKPARSE:option - dur_kind mismatch"); }
                else dur_kind = sc.ival;
                advance();
            } else if (tok == T_PARMKIND) {
                if (has_parm_kind) { if (parm_kind != sc.sval) fail("HTKPARSE:option - parm_kind_str already initialised"); }
                else { parm_kind = sc.sval; has_parm_kind = true; }
                advance();
            } else
                break;
        }
        if (!any) syntax("global options");
    }
    PState state_body(const char* what)    // [NUMMIXES] mixtures   (:305-338, :432-466)
    {
        PState st;
        if (tok == T_NUMMIXES) {           // the declared count is not checked against the list (:320, :451)
            advance();
            st.ml = mixtures();
        } else {
            st.ml = mixtures();
            if (st.ml.n_mixes != 1) fail(std::string("HTKPARSE:") + what + " - mixtures n_mixes value != 1");
        }
        return st;
    }
    void hmmdef()                          // HMACRO BEGINHMM NUMSTATES optglobopts states transmatdef ENDHMM   (:340-356)
    {
        PHmm h;
        h.name = sc.sval;
        advance();
        if (tok != T_BEGINHMM) syntax("hmmdef (<BeginHMM> expected)");
        advance();
        if (tok != T_NUMSTATES) syntax("hmmdef (<NumStates> expected)");
        h.n_states = sc.ival;
        advance();
        if (tok == T_OMACRO) { advance(); options(); }
        if (tok != T_STATE) syntax("hmmdef (<State> expected)");
        while (tok == T_STATE) {
            const int id = sc.ival;
            advance();
            PState st;
            if (tok == T_SMACRO) {          // statedef : STATE SMACRO   (:372-397)
                st.sh_name = sc.sval; st.named = true;
                size_t i = 0;
                for (; i < sh_states.size(); ++i) if (sh_states[i].sh_name == st.sh_name) break;
                if (i >= sh_states.size()) fail("HTKPARSE:statedef - SMACRO string not found in htk_def");
                advance();
            } else
                st = state_body("statedef");
            st.id = id;
            h.states.push_back(st);
        }
        if (tok == T_TMACRO) {              // transmatdef : TMACRO | transp   (:598-619)
            h.tm.sh_name = sc.sval; h.tm.named = true;
            size_t i = 0;
            for (; i < sh_transmats.size(); ++i) if (sh_transmats[i].sh_name == h.tm.sh_name) break;
            if (i >= sh_transmats.size()) fail("HTKPARSE:transmatdef - SMACRO string not found in htk_def");
            advance();
        } else
            h.tm = transp();
        if (tok != T_ENDHMM) syntax("hmmdef (<EndHMM> expected)");
        advance();
        if (h.n_states - 2 != (int)h.states.size()) fail("HTKPARSE:hmmdef - hmmstatelist n_elems did not match n_states");
        hmms.push_back(h);
    }
    void run()                             // htkdef : htkmacro+   (:76-170)
    {
        advance();
        if (tok == T_EOF) syntax("model file (no macro found)");
        while (tok != T_EOF) {
            switch (tok) {
            case T_OMACRO: advance(); options(); break;
            case T_HMACRO: hmmdef(); break;
            case T_TMACRO: {
                const std::string name = sc.sval;
                advance();
                PTransMat t = transp();
                t.sh_name = name; t.named = true;
                sh_transmats.push_back(t);
                break;
            }
            case T_SMACRO: {
                const std::string name = sc.sval;
                advance();
                PState st = state_body("shstatedef");
                st.sh_name = name; st.named = true;
                sh_states.push_back(st);
                break;
            }
            case T_VMACRO: {
                fprintf(stderr, "htkparse: ~v macros not supported - ignoring ~v \"%s\" definition\n", sc.sval.c_str());
                advance();
                sized_vec(T_VARIANCE, "variancevec");
                break;
            }
            case T_MMACRO: {               // ~m "<pool><n>" meanvec variancevec   (:118-170)
                const std::string macro = sc.sval;
                advance();
                const size_t len = strcspn(macro.c_str(), "0123456789");
                if (len == 0) fail("HTKPARSE:htkmacro - MMACRO pool name not found");
                const std::string pool = macro.substr(0, len);
                size_t i = 0;
                for (; i < pools.size(); ++i) if (pools[i].name == pool) break;
                if (i >= pools.size()) { pools.push_back(PPool()); pools.back().name = pool; }
                PMix m;
                m.means = sized_vec(T_MEAN, "meanvec");
                m.vars = sized_vec(T_VARIANCE, "variancevec");
                m.id = atoi(macro.c_str() + len);
                m.weight = 1.0f; m.gconst = 0.0f;
                if (m.id != (int)pools[i].mixes.size() + 1) fail("HTKPARSE:htkmacro - shmixdef mix id does not match pool n_mixes");
                pools[i].mixes.push_back(m);
                break;
            }
            default: syntax("model file (a ~o ~h ~s ~t ~m or ~v macro expected)");
            }
        }
    }
};

// ------------------------------------------------------------------------------------------
// HTKModels::initFromHTKParseResult and the add* methods it calls
struct Builder {
    const Parser& P;
    RawModels m;
    bool remove_tee;
    [[noreturn]] void fail(const std::string& s) { throw ParseError{s}; }

    int add_mean(const std::vector<float>& v)                        // addMeanVec   (:812-836)
    {
        m.means.insert(m.means.end(), v.begin(), v.begin() + m.D);
        return (int)(m.means.size() / m.D) - 1;
    }
    int add_var(const std::vector<float>& v)                         // addVarVec   (:839-875)
    {
        m.vars.insert(m.vars.end(), v.begin(), v.begin() + m.D);
        float sum = (float)(m.D * kLog2Pi);
        for (int i = 0; i < m.D; ++i) sum = (float)((double)sum + log((double)v[i]));
        sum = (float)((double)sum * -0.5);
        m.gconst.push_back(sum);
        return (int)m.gconst.size() - 1;
    }
    int add_mixture(const std::string& name, const std::vector<PMix>& comps, const char* who)   // addMixture   (:693-790)
    {
        std::vector<int> mi, vi;
        for (const PMix& c : comps) {
            if ((int)c.means.size() != m.D) fail(std::string("HTKModels::") + who + " - n_means != vecSize");
            mi.push_back(add_mean(c.means));
            if ((int)c.vars.size() != m.D) fail(std::string("HTKModels::") + who + " - n_vars != vecSize");
            vi.push_back(add_var(c.vars));
        }
        m.mix_mean.push_back(mi); m.mix_var.push_back(vi); m.mix_name.push_back(name);
        return (int)m.mix_mean.size() - 1;
    }
    static float logw(float w) { return w > 0.0f ? (float)log((double)w) : LZ; }
    int add_gmm(const PState& st)                                    // addGMM   (:596-676)
    {
        std::vector<float> lw;
        int mix;
        if (st.ml.pool_ind >= 0) {
            mix = -1;
            const std::string& pool = P.pools[st.ml.pool_ind].name;
            for (size_t i = 0; i < m.mix_name.size(); ++i)
                if (!m.mix_name[i].empty() && m.mix_name[i] == pool) { mix = (int)i; break; }
            if (mix < 0) fail("HTKModels::addGMM - shared mixture not found");
            const int nc = (int)m.mix_mean[mix].size();
            if (st.ml.n_mixes != nc) fail("HTKModels::addGMM - st->n_mixes != nComps");
            for (int i = 0; i < nc; ++i) lw.push_back(logw(st.ml.weights[i]));
            if (nc == 1 && st.ml.weights[0] != 1.0f) fail("HTKModels::addGMM - (nComps == 1) && (curr->compWeights[0] != 1.0)");
        } else {
            mix = add_mixture("", st.ml.mixes, "addMixture(2)");
            for (const PMix& c : st.ml.mixes) lw.push_back(logw(c.weight));
            if (st.ml.n_mixes == 1 && st.ml.mixes[0].weight != 1.0f)
                fail("HTKModels::addGMM - (st->n_mixes == 1) && (curr->compWeights[0] != 1.0)");
        }
        m.gmm_mix.push_back(mix); m.gmm_logw.push_back(lw); m.gmm_name.push_back(st.named ? st.sh_name : std::string());
        return (int)m.gmm_mix.size() - 1;
    }
    int add_transmat(const std::string& name, int n, const std::vector<float>& tr)   // addTransMatrix   (:878-974)
    {
        RawTransMat t;
        t.n = n; t.name = name;
        t.sucs.resize(n); t.logp.resize(n);
        std::vector<std::vector<float>> prob(n);
        float tee = 0.0f;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                const float p = tr[(size_t)i * n + j];
                if (!(p > 0.0f)) continue;
                if (i == 0 && j == n - 1 && remove_tee) {
                    if ((tee = p) >= 1.0f) fail("HTKModels::addTransMatrix - initial-final transition had prob. >= 1.0");
                    continue;
                }
                t.sucs[i].push_back(j);
                prob[i].push_back(p);
                t.logp[i].push_back((float)log((double)p));
            }
        if (tee > 0.0f)                      // re-normalise the entry state's row (:962-969)
            for (size_t i = 0; i < t.sucs[0].size(); ++i)
                t.logp[0][i] = (float)((double)t.logp[0][i] - log(1.0 - (double)tee));
        m.tms.push_back(t);
        return (int)m.tms.size() - 1;
    }
    void add_hmm(const PHmm& h)                                      // addHMM   (:519-593)
    {
        const int n = h.n_states;
        if (n < 2 || n > 64) fail("HMM \"" + h.name + "\": unsupported number of states");
        std::vector<int> g(n, -1);
        for (int i = 1; i < n - 1; ++i) {
            const PState& st = h.states[i - 1];
            if (st.named) {
                g[i] = -1;
                for (size_t k = 0; k < m.gmm_name.size(); ++k)
                    if (!m.gmm_name[k].empty() && m.gmm_name[k] == st.sh_name) { g[i] = (int)k; break; }
                if (g[i] < 0) fail("HTKModels::addHMM - shared state not found");
            } else
                g[i] = add_gmm(st);
        }
        int tm = -1;
        if (h.tm.named) {
            for (size_t k = 0; k < m.tms.size(); ++k)
                if (!m.tms[k].name.empty() && m.tms[k].name == h.tm.sh_name) { tm = (int)k; break; }
            if (tm < 0) fail("HTKModels::addHMM - shared transition matrix not found " + h.tm.sh_name);
            if (m.tms[tm].n != n) fail("HMM \"" + h.name + "\": shared transition matrix has a different number of states");
        } else {
            if (n != h.tm.n_states) fail("HTKModels::addHMM - curr->nStates != hmm->transmat->n_states");
            tm = add_transmat("", n, h.tm.transp);
        }
        m.hmm_n.push_back(n); m.hmm_g.push_back(g); m.hmm_tm.push_back(tm);
    }
    void run()                                                       // initFromHTKParseResult   (:397-444)
    {
        m.D = P.vec_size;
        if (m.D <= 0) fail("model file defines no <VecSize>");
        for (const PPool& p : P.pools) add_mixture(p.name, p.mixes, "addMixture");
        for (const PTransMat& t : P.sh_transmats) add_transmat(t.sh_name, t.n_states, t.transp);
        for (const PState& s : P.sh_states) add_gmm(s);
        for (const PHmm& h : P.hmms) add_hmm(h);
    }
};

} // namespace

extern "C" int jgpu_load_mmf(const char* path, int32_t remove_initial_to_final, JgpuHmm* hmm, JgpuGmm* gmm)
{
    if (!path || !hmm || !gmm) return jgpu_io_fail("jgpu_load_mmf: null argument");
    memset(hmm, 0, sizeof(*hmm));
    memset(gmm, 0, sizeof(*gmm));
    FILE* f = fopen(path, "rb");
    if (!f) return jgpu_io_fail("cannot open model file %s", path);
    std::string text;
    char buf[1 << 16];
    size_t k;
    while ((k = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, k);
    fclose(f);
    try {
        Parser p;
        p.sc.s = text.data(); p.sc.n = text.size();
        p.run();
        Builder b{p, RawModels(), remove_initial_to_final != 0};
        b.run();
        return jgpu_finish_models(b.m, hmm, gmm);
    } catch (const ParseError& e) {
        return jgpu_io_fail("%s: %s", path, e.msg.c_str());
    } catch (const std::exception& e) {                      // nothing unwinds across the C ABI
        return jgpu_io_fail("%s: %s", path, e.what());
    } catch (...) {
        return jgpu_io_fail("%s: unknown exception in the MMF loader", path);
    }
}
