/* TEST INFRASTRUCTURE ONLY (oracle) — never linked into, imported by or shipped with the product.
 *
 * htkparse() / cleanHTKDef() for the reference build in oracle/_ref.  In the reference these two
 * functions are bison/flex OUTPUT (src/htkparse.y.ypp, src/htkparse.l.lpp); bison and flex do not exist
 * in this image, so the generated parser cannot be produced.  This file restates the two inputs:
 *   - tokens: every flex rule of htkparse.l.lpp, in file order, as an anchored POSIX extended regular
 *     expression; the scanner takes the longest match, the earlier rule on a tie, and drops one
 *     character when nothing matches — flex's own disambiguation (POSIX regexec returns the
 *     leftmost-longest match of a pattern, which is what a flex rule matches);
 *   - grammar: the productions of htkparse.y.ypp as a recursive descent, each action filling the
 *     reference's own `htk_def` (src/htkparse.h) with malloc'ed records exactly as the bison action
 *     does, so that the UNMODIFIED HTKModels::initFromHTKParseResult (src/HTKModels.cpp:397-444)
 *     consumes it.
 * What this pins: everything the reference does with a parsed model set (its own compiled code).
 * What it cannot pin: bison's tables themselves ("grammar half unpinned").
 */
#include <regex.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <setjmp.h>

#include "htkparse.h"

HTKDef htk_def;

namespace {

enum {
    INTEGER = 1, REAL, QUOTEDSTRING, STRING, BEGINHMM, ENDHMM, NUMSTATES, STATE, NUMMIXES, MIXTURE, MEAN, VARIANCE,
    GCONST, TRANSP, HMMSETID, TMIX, VECSIZE, STREAMINFO, DIAGC, INVDIAGC, FULLC, LLTC, XFORMC, NULLD, POISSOND,
    GAMMAD, GEND, PARMKIND, HMACRO, SMACRO, MMACRO, TMACRO, VMACRO, OMACRO, WS, END = 0
};

#define D_   "[0-9]"
#define W_   "[ \t\r\n]"
#define INT_ "-?" D_ "+"
#define RE_  "(\\+|-)?(" D_ "+\\.?" D_ "*|\\." D_ "+)([eE](\\+|-)?" D_ "+)?"
#define STR_ "[]a-zA-Z0-9+$^#@_&:[-]+"
#define QSTR_ "\"" STR_ "\""
#define PKB1 "Discrete|LPCepstra|FBank|MelSpec|LPRefC|User"
#define PKB2 "DISCRETE|LPC|LPCEPSTRA|MFCC|PLP|FBANK|MELSPEC|LPREFC|USER"
#define PKB3 "discrete|lpc|lpcepstra|mfcc|plp|fbank|melspec|lprefc|user"
#define PKQ1 "_D|_A|_T|_E|_N|_Z|_O|_V|_C|_K|_0"
#define PKQ2 "_d|_a|_t|_e|_n|_z|_o|_v|_c|_k|_0"

struct Rule { int tok; const char* re; regex_t rx; };
Rule rules[] = {                                     /* htkparse.l.lpp:36-268, in file order */
    {INTEGER, "^" INT_},
    {REAL, "^" RE_},
    {QUOTEDSTRING, "^" QSTR_},
    {STRING, "^" STR_},
    {BEGINHMM, "^<(BeginHMM|BEGINHMM|beginhmm)>"},
    {ENDHMM, "^<(EndHMM|ENDHMM|endhmm)>"},
    {NUMSTATES, "^<(NumStates|NUMSTATES|numstates)>" W_ "*" INT_},
    {STATE, "^<(State|STATE|state)>" W_ "*" INT_},
    {NUMMIXES, "^<(NumMixes|NUMMIXES|nummixes)>" W_ "*" INT_},
    {MIXTURE, "^<(Mixture|MIXTURE|mixture)>" W_ "*" INT_},
    {MEAN, "^<(Mean|MEAN|mean)>" W_ "*" INT_},
    {VARIANCE, "^<(Variance|VARIANCE|variance)>" W_ "*" INT_},
    {GCONST, "^<(GConst|GCONST|gconst)>" W_ "*(" RE_ "|" INT_ ")"},
    {TRANSP, "^<(TransP|TRANSP|transp)>" W_ "*" INT_},
    {HMMSETID, "^<(HmmSetId|HMMSETID|hmmsetid)>" W_ "*(" STR_ "|" QSTR_ ")"},
    {TMIX, "^<(TMix|TMIX|tmix)>" W_ "*(" STR_ "|" QSTR_ ")"},
    {VECSIZE, "^<(VecSize|VECSIZE|vecsize)>" W_ "*" INT_},
    {STREAMINFO, "^<(StreamInfo|STREAMINFO|streaminfo)>" W_ "*" INT_},
    {DIAGC, "^<(DiagC|DIAGC|diagc)>"},
    {INVDIAGC, "^<(InvDiagC|INVDIAGC|invdiagc)>"},
    {FULLC, "^<(FullC|FULLC|fullc)>"},
    {LLTC, "^<(LLTC|lltc)>"},
    {XFORMC, "^<(XFormC|XFORMC|xformc)>"},
    {NULLD, "^<(NullD|NULLD|nulld)>"},
    {POISSOND, "^<(PoissonD|POISSOND|poissond)>"},
    {GAMMAD, "^<(GammaD|GAMMAD|gammad)>"},
    {GEND, "^<(GenD|GEND|gend)>"},
    {PARMKIND, "^<(" PKB1 "|" PKB2 "|" PKB3 ")(" PKQ1 "|" PKQ2 ")*>"},
    {HMACRO, "^~h" W_ "+" QSTR_},
    {SMACRO, "^~s" W_ "+" QSTR_},
    {MMACRO, "^~m" W_ "+" QSTR_},
    {TMACRO, "^~t" W_ "+" QSTR_},
    {VMACRO, "^~v" W_ "+" QSTR_},
    {OMACRO, "^~o" W_ "+"},
    {WS, "^" W_},
};
const int n_rules = sizeof(rules) / sizeof(rules[0]);
bool rules_ready = false;

char* text = NULL;          /* whole file, NUL terminated */
size_t text_len = 0, pos = 0;
int tok;                    /* look-ahead */
int ival; real fval; char* cptr;
jmp_buf bail;

char* dup(const char* s) { char* r = (char*)malloc(strlen(s) + 1); strcpy(r, s); return r; }

/* second strtok token, as the flex actions extract their values */
char* second(char* s, const char* delim) { strtok(s, delim); return strtok(NULL, delim); }

int lex()
{
    for (;;) {
        if (pos >= text_len) return END;
        /* a bounded window keeps regexec's strlen linear over the whole file */
        char win[2048];
        size_t w = text_len - pos < sizeof(win) - 1 ? text_len - pos : sizeof(win) - 1;
        memcpy(win, text + pos, w); win[w] = 0;
        for (size_t i = 0; i < w; ++i) if (win[i] == 0) { win[i] = 1; }
        int best = -1; regoff_t best_len = 0;
        for (int r = 0; r < n_rules; ++r) {
            regmatch_t m;
            if (regexec(&rules[r].rx, win, 1, &m, 0) == 0 && m.rm_so == 0 && m.rm_eo > best_len) { best = r; best_len = m.rm_eo; }
        }
        if (best < 0) { ++pos; continue; }               /* `.` rule: ignore */
        win[best_len] = 0;
        pos += best_len;
        const int t = rules[best].tok;
        switch (t) {
        case WS: continue;
        case INTEGER: ival = atoi(win); return t;
        case REAL: fval = (float)atof(win); return t;
        case QUOTEDSTRING: cptr = dup(strtok(win, "\"")); return t;
        case STRING: cptr = dup(win); return t;
        case NUMSTATES: case STATE: case NUMMIXES: case MIXTURE: case MEAN: case VARIANCE: case TRANSP: case VECSIZE:
        case STREAMINFO: ival = atoi(second(win, " \r\n\t>")); return t;
        case GCONST: fval = (float)atof(second(win, " \r\n\t>")); return t;
        case HMMSETID: case TMIX: cptr = dup(second(win, " \r\n\t\">")); return t;
        case PARMKIND: cptr = dup(strtok(win, "<>")); return t;
        case HMACRO: cptr = dup(second(win, " \"\r\n\t")); return t;
        case SMACRO: case MMACRO: case TMACRO: case VMACRO: { strtok(win, " \r\n\t"); cptr = dup(strtok(NULL, " \"\r\n\t")); return t; }
        default: return t;
        }
    }
}

void next() { tok = lex(); }
void htkerror(const char* s) { fprintf(stderr, "%s\n", s); longjmp(bail, 2); }   /* the reference exits; tests want a status */
void syntax() { fprintf(stderr, "htkparse: syntax error near byte %zu\n", pos); longjmp(bail, 1); }

RealVector* rvector()
{
    if (tok != INTEGER && tok != REAL) syntax();
    RealVector* v = (RealVector*)malloc(sizeof(RealVector));
    v->n_elems = 0; v->elems = NULL;
    while (tok == INTEGER || tok == REAL) {
        v->n_elems++;
        v->elems = (real*)realloc(v->elems, v->n_elems * sizeof(real));
        v->elems[v->n_elems - 1] = tok == INTEGER ? (real)ival : fval;
        next();
    }
    return v;
}

RealVector* meanvec()
{
    if (tok != MEAN) syntax();
    const int n = ival;
    if (n != htk_def.global_opts.vec_size) htkerror("HTKPARSE:meanvec - MEAN value did not match global vec size\n");
    next();
    RealVector* v = rvector();
    if (v->n_elems != n) htkerror("HTKPARSE:meanvec - n_elems did not match MEAN value\n");
    return v;
}

RealVector* variancevec()
{
    if (tok != VARIANCE) syntax();
    const int n = ival;
    if (n != htk_def.global_opts.vec_size) htkerror("HTKPARSE:variancevec - VARIANCE value did not match global vec size\n");
    next();
    RealVector* v = rvector();
    if (v->n_elems != n) htkerror("HTKPARSE:variancevec - n_elems did not match VARIANCE value\n");
    return v;
}

HTKMixture* mixpdf()
{
    RealVector* m = meanvec();
    RealVector* v = variancevec();
    HTKMixture* mix = (HTKMixture*)malloc(sizeof(HTKMixture));
    mix->n_means = m->n_elems; mix->means = m->elems;
    mix->n_vars = v->n_elems; mix->vars = v->elems;
    free(m); free(v);
    mix->gconst = 0.0;
    if (tok == GCONST) { mix->gconst = fval; next(); }
    return mix;
}

HTKMixture* mixturedef()
{
    if (tok == MIXTURE) {
        const int id = ival;
        next();
        if (tok != REAL) syntax();
        const real w = fval;
        next();
        HTKMixture* m = mixpdf();
        m->id = id; m->weight = w;
        return m;
    }
    HTKMixture* m = mixpdf();
    m->id = 1; m->weight = 1.0;
    return m;
}

HTKMixtureList* mixtures()
{
    HTKMixtureList* l = (HTKMixtureList*)malloc(sizeof(HTKMixtureList));
    if (tok == TMIX) {
        char* name = cptr;
        next();
        RealVector* w = rvector();
        int i;
        for (i = 0; i < htk_def.n_mix_pools; i++)
            if (strcmp(htk_def.mix_pools[i]->name, name) == 0) break;
        if (i >= htk_def.n_mix_pools) htkerror("HTKPARSE:mixtures - TMIX string did not match the name of a mix pool\n");
        if (w->n_elems != htk_def.mix_pools[i]->n_mixes) htkerror("HTKPARSE:mixtures - tmixweights n_elems did not match n_mixes in mix pool\n");
        l->n_mixes = w->n_elems; l->mixes = NULL; l->pool_ind = i; l->weights = w->elems;
        free(w);
    } else {
        if (tok != MIXTURE && tok != MEAN) syntax();
        l->n_mixes = 1; l->pool_ind = -1; l->weights = NULL;
        l->mixes = (HTKMixture**)malloc(sizeof(HTKMixture*));
        l->mixes[0] = mixturedef();
    }
    while (tok == MIXTURE || tok == MEAN) {                   /* mixturelist : mixtures mixturedef */
        HTKMixture* m = mixturedef();
        l->n_mixes++;
        l->mixes = (HTKMixture**)realloc(l->mixes, l->n_mixes * sizeof(HTKMixture*));
        l->mixes[l->n_mixes - 1] = m;
    }
    return l;
}

HTKTransMat* transp()
{
    if (tok != TRANSP) syntax();
    const int n = ival;
    next();
    RealVector* v = rvector();
    if (n == 0 || n != (v->n_elems / n)) htkerror("HTKPARSE:transp - vec n_elems did not match TRANSP value\n");
    HTKTransMat* tm = (HTKTransMat*)malloc(sizeof(HTKTransMat));
    tm->sh_name = NULL; tm->n_states = n;
    tm->transp = (real**)malloc(n * sizeof(real*));
    for (int i = 0, k = 0; i < n; i++) {
        tm->transp[i] = (real*)malloc(n * sizeof(real));
        for (int j = 0; j < n; j++) tm->transp[i][j] = v->elems[k++];
    }
    free(v->elems); free(v);
    return tm;
}

bool is_option() {
    return tok == HMMSETID || tok == STREAMINFO || tok == VECSIZE || tok == PARMKIND || (tok >= DIAGC && tok <= GEND);
}

void option()
{
    HTKGlobalOpts& g = htk_def.global_opts;
    if (tok == HMMSETID) {
        if (g.hmm_set_id != NULL) { if (strcmp(g.hmm_set_id, cptr) != 0) htkerror("HTKPARSE:option - hmm_set_id mismatch\n"); }
        else g.hmm_set_id = cptr;
        next();
    } else if (tok == STREAMINFO) {
        const int ns = ival;
        next();
        if (tok != INTEGER) syntax();
        IntVector iv; iv.n_elems = 0; iv.elems = NULL;
        while (tok == INTEGER) {
            iv.n_elems++;
            iv.elems = (int*)realloc(iv.elems, iv.n_elems * sizeof(int));
            iv.elems[iv.n_elems - 1] = ival;
            next();
        }
        int i, sum;
        if (iv.n_elems != ns) htkerror("HTKPARSE:option - STREAMINFO value does not match ivec size\n");
        if (g.vec_size > 0) {
            for (i = 0, sum = 0; i < iv.n_elems; i++) sum += iv.elems[i];
            if (sum != g.vec_size) htkerror("HTKPARSE:option - sum of stream widths does not equal vec_size\n");
        }
        if (g.n_streams > 0) {
            if (g.n_streams != ns) htkerror("HTKPARSE:option - n_streams mismatch\n");
            free(iv.elems);
        } else { g.n_streams = ns; g.stream_widths = iv.elems; }
    } else if (tok == VECSIZE) {
        int i, sum;
        if (g.n_streams > 0) {
            for (i = 0, sum = 0; i < g.n_streams; i++) sum += g.stream_widths[i];
            if (sum != ival) htkerror("HTKPARSE:option - sum of stream widths does not equal NEW vec_size\n");
        }
        if (g.vec_size > 0) { if (g.vec_size != ival) htkerror("HTKPARSE:option - vec_size mismatch\n"); }
        else g.vec_size = ival;
        next();
    } else if (tok >= DIAGC && tok <= XFORMC) {
        const CovKind k = (CovKind)(CK_DIAGC + (tok - DIAGC));
        if (g.cov_kind != CK_INVALID) { if (g.cov_kind != k) htkerror("HTKPARSE:option - cov_kind mismatch\n"); }
        else g.cov_kind = k;
        next();
    } else if (tok >= NULLD && tok <= GEND) {
        const DurKind k = (DurKind)(DK_NULLD + (tok - NULLD));
        if (g.dur_kind != DK_INVALID) { if (g.dur_kind != k) htkerror("HTKPARSE:option - dur_kind mismatch\n"); }
        else g.dur_kind = k;
        next();
    } else if (tok == PARMKIND) {
        if (g.parm_kind_str != NULL) { if (strcmp(g.parm_kind_str, cptr) != 0) htkerror("HTKPARSE:option - parm_kind_str already initialised\n"); }
        else g.parm_kind_str = cptr;
        next();
    } else
        syntax();
}

void globalopts()            /* OMACRO options */
{
    next();
    option();
    while (is_option()) option();
}

HTKHMMState* new_state(char* sh_name, int id, HTKMixtureList* l, bool counted, const char* msg)
{
    HTKHMMState* st = (HTKHMMState*)malloc(sizeof(HTKHMMState));
    if (!counted && l->n_mixes != 1) htkerror(msg);
    st->sh_name = sh_name; st->id = id;
    st->n_mixes = counted ? l->n_mixes : 1;
    st->mixes = l->mixes; st->pool_ind = l->pool_ind; st->weights = l->weights;
    free(l);
    return st;
}

HTKHMMState* statedef()
{
    const int id = ival;     /* STATE */
    next();
    if (tok == SMACRO) {
        HTKHMMState* st = (HTKHMMState*)malloc(sizeof(HTKHMMState));
        st->sh_name = cptr; st->id = id; st->n_mixes = 0; st->mixes = NULL; st->pool_ind = -1; st->weights = NULL;
        int i;
        for (i = 0; i < htk_def.n_sh_states; i++)
            if (strcmp(htk_def.sh_states[i]->sh_name, st->sh_name) == 0) break;
        if (i >= htk_def.n_sh_states) htkerror("HTKPARSE:statedef - SMACRO string not found in htk_def\n");
        next();
        return st;
    }
    if (tok == NUMMIXES) { next(); return new_state(NULL, id, mixtures(), true, ""); }
    return new_state(NULL, id, mixtures(), false, "HTKPARSE:statedef - mixtures n_mixes value != 1\n");
}

HTKHMM* hmmdef()
{
    char* name = cptr;       /* HMACRO */
    next();
    if (tok != BEGINHMM) syntax();
    next();
    if (tok != NUMSTATES) syntax();
    const int ns = ival;
    next();
    if (tok == OMACRO) globalopts();
    if (tok != STATE) syntax();
    HTKHMMStateList sl; sl.n_states = 0; sl.states = NULL;
    while (tok == STATE) {
        HTKHMMState* st = statedef();
        sl.n_states++;
        sl.states = (HTKHMMState**)realloc(sl.states, sl.n_states * sizeof(HTKHMMState*));
        sl.states[sl.n_states - 1] = st;
    }
    HTKTransMat* tm;
    if (tok == TMACRO) {
        tm = (HTKTransMat*)malloc(sizeof(HTKTransMat));
        tm->sh_name = cptr; tm->n_states = 0; tm->transp = NULL;
        int i;
        for (i = 0; i < htk_def.n_sh_transmats; i++)
            if (strcmp(htk_def.sh_transmats[i]->sh_name, tm->sh_name) == 0) break;
        if (i >= htk_def.n_sh_transmats) htkerror("HTKPARSE:transmatdef - SMACRO string not found in htk_def\n");
        next();
    } else
        tm = transp();
    if (tok != ENDHMM) syntax();
    next();
    HTKHMM* hmm = (HTKHMM*)malloc(sizeof(HTKHMM));
    hmm->name = name; hmm->n_states = ns;
    if ((hmm->n_states - 2) != sl.n_states) htkerror("HTKPARSE:hmmdef - hmmstatelist n_elems did not match n_states\n");
    hmm->emit_states = sl.states;
    hmm->transmat = tm;
    return hmm;
}

void mmacro()
{
    char* macro = cptr;
    next();
    RealVector* mv = meanvec();
    RealVector* vv = variancevec();
    int i, len;
    char name[100];
    HTKMixturePool* pool = NULL;
    if ((len = strcspn(macro, "0123456789")) == 0) htkerror("HTKPARSE:htkmacro - MMACRO pool name not found\n");
    strncpy(name, macro, len * sizeof(char));
    name[len] = '\0';
    for (i = 0; i < htk_def.n_mix_pools; i++)
        if (strcmp(htk_def.mix_pools[i]->name, name) == 0) { pool = htk_def.mix_pools[i]; break; }
    if (i >= htk_def.n_mix_pools) {
        htk_def.n_mix_pools++;
        htk_def.mix_pools = (HTKMixturePool**)realloc(htk_def.mix_pools, htk_def.n_mix_pools * sizeof(HTKMixturePool*));
        pool = (HTKMixturePool*)malloc(sizeof(HTKMixturePool));
        pool->name = dup(name); pool->n_mixes = 0; pool->mixes = NULL;
        htk_def.mix_pools[htk_def.n_mix_pools - 1] = pool;
    }
    HTKMixture* mix = (HTKMixture*)malloc(sizeof(HTKMixture));
    mix->id = atoi(macro + len);
    mix->weight = 1.0;
    mix->n_means = mv->n_elems; mix->means = mv->elems;
    mix->n_vars = vv->n_elems; mix->vars = vv->elems;
    free(mv); free(vv);
    mix->gconst = 0.0;
    if (mix->id != (pool->n_mixes + 1)) htkerror("HTKPARSE:htkmacro - shmixdef mix id does not match pool n_mixes\n");
    pool->n_mixes++;
    pool->mixes = (HTKMixture**)realloc(pool->mixes, pool->n_mixes * sizeof(HTKMixture*));
    pool->mixes[pool->n_mixes - 1] = mix;
    free(macro);
}

void htkmacro()
{
    switch (tok) {
    case OMACRO: globalopts(); break;
    case HMACRO: {
        HTKHMM* h = hmmdef();
        htk_def.n_hmms++;
        htk_def.hmms = (HTKHMM**)realloc(htk_def.hmms, htk_def.n_hmms * sizeof(HTKHMM*));
        htk_def.hmms[htk_def.n_hmms - 1] = h;
        break;
    }
    case TMACRO: {
        char* name = cptr;
        next();
        HTKTransMat* tm = transp();
        tm->sh_name = name;
        htk_def.n_sh_transmats++;
        htk_def.sh_transmats = (HTKTransMat**)realloc(htk_def.sh_transmats, htk_def.n_sh_transmats * sizeof(HTKTransMat*));
        htk_def.sh_transmats[htk_def.n_sh_transmats - 1] = tm;
        break;
    }
    case SMACRO: {
        char* name = cptr;
        next();
        HTKHMMState* st;
        if (tok == NUMMIXES) { next(); st = new_state(name, -1, mixtures(), true, ""); }
        else st = new_state(name, -1, mixtures(), false, "HTKPARSE:shstatedef - mixtures n_mixes value != 1\n");
        htk_def.n_sh_states++;
        htk_def.sh_states = (HTKHMMState**)realloc(htk_def.sh_states, htk_def.n_sh_states * sizeof(HTKHMMState*));
        htk_def.sh_states[htk_def.n_sh_states - 1] = st;
        break;
    }
    case VMACRO: {
        char* name = cptr;
        next();
        RealVector* v = variancevec();
        fprintf(stderr, "htkparse: ~v macros not supported - ignoring ~v \"%s\" definition\n", name);
        free(name);
        if (v->elems != NULL) free(v->elems);
        free(v);
        break;
    }
    case MMACRO: mmacro(); break;
    default: syntax();
    }
}

void initHTKDef()
{
    memset(&htk_def, 0, sizeof(htk_def));
    htk_def.global_opts.cov_kind = CK_INVALID;
    htk_def.global_opts.dur_kind = DK_INVALID;
}

void clean_mix(HTKMixture* m) { if (m->means) free(m->means); if (m->vars) free(m->vars); free(m); }
void clean_tm(HTKTransMat* tm)
{
    if (tm->sh_name) free(tm->sh_name);
    if (tm->transp) { for (int i = 0; i < tm->n_states; i++) free(tm->transp[i]); free(tm->transp); }
    free(tm);
}
void clean_state(HTKHMMState* st)
{
    if (st->sh_name) free(st->sh_name);
    if (st->mixes) { for (int i = 0; i < st->n_mixes; i++) clean_mix(st->mixes[i]); free(st->mixes); }
    if (st->weights) free(st->weights);
    free(st);
}

} // namespace

int htkparse(void* fd)
{
    if (!rules_ready) {
        for (int r = 0; r < n_rules; ++r)
            if (regcomp(&rules[r].rx, rules[r].re, REG_EXTENDED) != 0) { fprintf(stderr, "htkparse: bad rule %d\n", r); return 3; }
        rules_ready = true;
    }
    FILE* f = (FILE*)fd;
    free(text); text = NULL; text_len = 0; pos = 0;
    char buf[1 << 16];
    size_t k;
    while ((k = fread(buf, 1, sizeof(buf), f)) > 0) {
        text = (char*)realloc(text, text_len + k + 1);
        memcpy(text + text_len, buf, k);
        text_len += k;
    }
    if (text) text[text_len] = 0;
    initHTKDef();
    const int rc = setjmp(bail);
    if (rc != 0) return rc;
    next();
    if (tok == END) syntax();              /* htkmacros needs at least one macro */
    while (tok != END) htkmacro();
    return 0;
}

void cleanHTKDef()
{
    HTKGlobalOpts& g = htk_def.global_opts;
    if (g.hmm_set_id) free(g.hmm_set_id);
    if (g.stream_widths) free(g.stream_widths);
    if (g.parm_kind_str) free(g.parm_kind_str);
    for (int i = 0; i < htk_def.n_sh_transmats; i++) clean_tm(htk_def.sh_transmats[i]);
    free(htk_def.sh_transmats);
    for (int i = 0; i < htk_def.n_sh_states; i++) clean_state(htk_def.sh_states[i]);
    free(htk_def.sh_states);
    for (int i = 0; i < htk_def.n_mix_pools; i++) {
        HTKMixturePool* p = htk_def.mix_pools[i];
        if (p->name) free(p->name);
        for (int j = 0; j < p->n_mixes; j++) clean_mix(p->mixes[j]);
        free(p->mixes); free(p);
    }
    free(htk_def.mix_pools);
    for (int i = 0; i < htk_def.n_hmms; i++) {
        HTKHMM* h = htk_def.hmms[i];
        if (h->name) free(h->name);
        if (h->emit_states) { for (int j = 0; j < h->n_states - 2; j++) clean_state(h->emit_states[j]); free(h->emit_states); }
        if (h->transmat) clean_tm(h->transmat);
        free(h);
    }
    free(htk_def.hmms);
    initHTKDef();
}
