// jgpu_engine.cu — host engine + C ABI (include/juicer_b200.h) over the sm_100a kernels.
//
// The engine owns the device-resident network / model tables and the per-lane decoding
// state, turns utterances into a lock-step schedule (one table row per frame step and
// lane), and enqueues the per-step kernel sequence on one CUDA stream.  Nothing in the
// decode loop synchronises with the host: thresholds, list sizes and round counts all live
// in device memory; the host only waits in utt_end / at the end of a batch.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <queue>
#include <string>
#include <vector>

#include "jgpu_device.cuh"
#include "jgpu_err.h"
#include "host_queue.h"
#include "jgpu_gmm.cuh"
#define JG_HUGE_DEG 1024   // rows with at least this many model arcs are walked by all CTAs of the lane
#include "jgpu_search.cuh"

#ifdef JG_TRACE
#define JG_TRACING(h) ((h)->d_trace != nullptr)
#else
#define JG_TRACING(h) false
#endif

// non-FMA FP32 issue-rate microbenchmark (jgpu_ubench_fp32)
__global__ void __launch_bounds__(256) k_ubench_fp32(float* out, int iters, float m)
{
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 1.0f + threadIdx.x * 1e-6f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { a[i] = __fmul_rn(a[i], m); a[i] = __fmul_rn(a[i], 1.0009765625f); }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

namespace {

int fail(int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    jgpu_err_buf() = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(JGPU_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct LaneHost {
    bool begun = false;
    bool seeded = false;
    int frames = 0;
};

} // namespace

struct jgpu_handle {
    Dev d{};
    GmmDev g{};
    JgpuCfg cfg{};
    int device = 0;
    int S = 5;              // kernel template parameter (5 or 8)
    int DP = 40;            // padded feature dimension
    int dim = 39;
    int FB = 16;            // frames per GMM scoring block
    int bpl = 8;            // CTAs per lane for the search kernels
    bool has_huge = false;
    int max_deg = 0, n_huge_states = 0;
    cudaGraphExec_t graph_step = nullptr, graph_block = nullptr;   // one frame step / FB frame steps of all lanes
    bool use_graphs = true;
    std::vector<unsigned> host_epoch;    // mirror of LaneCtl::epoch (advances on every non-idle step of the lane)
    int gc_period = 64, steps_since_gc = 0;   // word-boundary arena collection every gc_period frame steps (0 = never)
    cudaStream_t stream = nullptr;       // everything is enqueued here
    std::vector<void*> allocs;
    size_t bytes = 0;
    // schedule
    int sched_chunk = 2048;
    int4* d_sched = nullptr;
    int* d_rows = nullptr;
    std::vector<int> rows_host;          // staging of the feature-row table of a chunk
    // features / scores
    float* d_feats = nullptr;
    size_t feats_cap = 0;   // rows
    float* d_stream_feats = nullptr;
    int stream_chunk = 256;
    float* d_scores = nullptr;
    float* d_gmm_out = nullptr;   // jgpu_gmm_scores scratch
    // lazy scorer (k_gmm_lazy): per-component parameter rows, the per-step feature tile and the demand stamps
    bool lazy = false;            // JUICER_B200_LAZY=1: score only the stamped (GMM, lane) pairs, once per step (see DESIGN.md:
                                  // measured slower than scoring everything 16 frames ahead on the BASELINE workloads)
    int lazy_cluster = JG_LAZY_CLUSTER;   // CTAs sharing one multicast feature tile (JUICER_B200_LAZY_CLUSTER = 1 | 2 | 4)
    LazyArgs lz{};
    const float** d_feat_base = nullptr;
    int* d_partial = nullptr;     // scratch of jgpu_partial_result
    int gmm_chunk = 1024;
    // results
    size_t res_cap = 0;                  // utterance headers
    size_t words_cap = 0;                // word pool of a batch (JgpuWord records)
    // views: the per-lane arenas are pools (n_lanes x capacity); a second pass over the utterances that overflowed
    // theirs looks at the same memory as FEWER lanes with LARGER arenas (decode_common)
    struct View { int n_lanes, cap, cap_arr, cap_paths; };
    View base{};                         // what jgpu_create sized
    size_t pool_inst = 0, pool_arr = 0, pool_paths = 0;   // pool sizes in records
    int cap_full = 0;                    // instances a lane can ever hold: one per arc (+1)
    int retry_passes = 0;                // second passes run so far (statistics)
    bool sticky = false;                 // the last batch mostly overflowed the base view: start with sticky_view
    View sticky_view{};
    unsigned epoch_wrap = 0x7ffu;        // a lane's stamped tables are wiped when (epoch & epoch_wrap) == 0
    std::vector<LaneHost> lanes;
    int64_t launches = 0;
    JgpuStats batch_stats{};
    char* static_base = nullptr;   // arcs | states | arc_tee | hmm tables
    size_t static_bytes = 0;
    bool own_stream = true;
    int n_sm = 148;
    // optional per-kernel timing (CUDA events on the launching stream)
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;
    std::vector<int> prof_kind;
    size_t prof_used = 0;
    double prof_ms[JGPU_N_KERNELS] = {0};
    int64_t prof_cnt[JGPU_N_KERNELS] = {0};

#ifdef JG_TRACE
    TraceRec* d_trace = nullptr;
    unsigned trace_cap = 0;
    int trace_s0 = 0, trace_s1 = 0, abs_step = 0;
    std::string trace_path;
#endif

    void prof_begin(int kind, cudaStream_t st = nullptr)
    {
        if (!prof_on) return;
        if (!st) st = stream;
        if (prof_used + 2 > prof_ev.size()) {
            const size_t n = prof_ev.size() + 4096;
            while (prof_ev.size() < n) {
                cudaEvent_t e;
                cudaEventCreate(&e);
                prof_ev.push_back(e);
            }
        }
        prof_kind.push_back(kind);
        cudaEventRecord(prof_ev[prof_used++], st);
    }
    void prof_end(cudaStream_t st = nullptr)
    {
        if (!prof_on) return;
        cudaEventRecord(prof_ev[prof_used++], st ? st : stream);
    }
    void prof_collect()
    {
        if (prof_used == 0) return;
        cudaStreamSynchronize(stream);
        for (size_t i = 0; i + 1 < prof_used; i += 2) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, prof_ev[i], prof_ev[i + 1]);
            const int k = prof_kind[i / 2];
            prof_ms[k] += ms;
            prof_cnt[k] += 1;
        }
        prof_used = 0;
        prof_kind.clear();
    }

    template <typename T>
    int alloc(T** p, size_t n, bool zero = true)
    {
        void* q = nullptr;
        const size_t b = std::max<size_t>(n, 1) * sizeof(T);
        cudaError_t e = cudaMalloc(&q, b);
        if (e != cudaSuccess)
            return fail(JGPU_E_CUDA, "cudaMalloc(%zu bytes) failed: %s (already holding %zu bytes)", b,
                        cudaGetErrorString(e), bytes);
        if (zero) {
            e = cudaMemsetAsync(q, 0, b, stream);
            if (e != cudaSuccess) return fail(JGPU_E_CUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
        }
        allocs.push_back(q);
        bytes += b;
        *p = (T*)q;
        return JGPU_OK;
    }
    void release(void* p)
    {
        if (!p) return;
        auto it = std::find(allocs.begin(), allocs.end(), p);
        if (it != allocs.end()) allocs.erase(it);
        cudaFree(p);
    }
};

namespace {

template <typename T>
int upload(jgpu_handle* h, T** dst, const std::vector<T>& src)
{
    int rc = h->alloc(dst, src.size(), false);
    if (rc) return rc;
    if (!src.empty()) CK(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    return JGPU_OK;
}

// longest chain of pass-through arcs (epsilon input or tee model); -1 on a cycle.
int passthrough_depth(const JgpuNet* n, const std::vector<char>& pass, std::vector<int>* topo = nullptr)
{
    const int S = n->n_states;
    std::vector<int> indeg(S, 0), depth(S, 0);
    for (int s = 0; s < S; ++s)
        for (int b = n->state_first[s]; b < n->state_first[s] + n->state_narcs[s]; ++b)
            if (pass[b]) ++indeg[n->arc_to[b]];
    std::vector<int> q;
    q.reserve(S);
    for (int s = 0; s < S; ++s)
        if (indeg[s] == 0) q.push_back(s);
    size_t head = 0;
    int best = 0;
    while (head < q.size()) {
        const int s = q[head++];
        for (int b = n->state_first[s]; b < n->state_first[s] + n->state_narcs[s]; ++b) {
            if (!pass[b]) continue;
            const int t = n->arc_to[b];
            depth[t] = std::max(depth[t], depth[s] + 1);
            best = std::max(best, depth[t]);
            if (--indeg[t] == 0) q.push_back(t);
        }
    }
    if ((int)q.size() != S) return -1;
    if (topo) *topo = q;
    return best;
}

int validate(const JgpuNet* n, const JgpuHmm* m, const JgpuGmm* g, const JgpuCfg* c)
{
    if (!n || !m || !g || !c) return fail(JGPU_E_ARG, "null argument");
    if (n->n_states <= 0 || n->n_arcs < 0) return fail(JGPU_E_ARG, "empty network");
    if (n->n_states > JG_STATE_MASK) return fail(JGPU_E_ARG, "more than 2^30 states");
    if (n->init_state < 0 || n->init_state >= n->n_states) return fail(JGPU_E_ARG, "init_state out of range");
    if (m->n_hmms <= 0 || m->max_states < 2 || m->max_states > 8)
        return fail(JGPU_E_ARG, "max_states=%d unsupported (2..8)", m->max_states);
    if (g->dim <= 0 || g->dim > JG_GMM_DMAX) return fail(JGPU_E_ARG, "feature dim %d unsupported (1..%d)", g->dim, JG_GMM_DMAX);
    if (g->max_comps <= 0 || g->max_comps > 256) return fail(JGPU_E_ARG, "max_comps %d unsupported (1..256)", g->max_comps);
    if (c->n_lanes < 1 || c->n_lanes > JG_MAX_LANES) return fail(JGPU_E_ARG, "n_lanes %d out of range (1..%d)", c->n_lanes, JG_MAX_LANES);
    for (int s = 0; s < n->n_states; ++s) {
        const int f = n->state_first[s], k = n->state_narcs[s];
        if (k < 0 || (k > 0 && (f < 0 || f + k > n->n_arcs))) return fail(JGPU_E_ARG, "state %d arc range invalid", s);
    }
    for (int a = 0; a < n->n_arcs; ++a) {
        if (n->arc_to[a] < 0 || n->arc_to[a] >= n->n_states) return fail(JGPU_E_ARG, "arc %d: toState out of range", a);
        // the reference never checks this (checkConsistency is dead code, src/juicer.cpp:1001-1061)
        if (n->arc_in[a] < 0 || n->arc_in[a] > m->n_hmms)
            return fail(JGPU_E_ARG, "arc %d: input label %d is not an HMM index + 1 (n_hmms=%d)", a, n->arc_in[a], m->n_hmms);
        if (n->arc_out[a] < 0) return fail(JGPU_E_ARG, "arc %d: negative output label", a);
    }
    for (int h = 0; h < m->n_hmms; ++h) {
        const int ns = m->n_states[h];
        if (ns < 2 || ns > m->max_states) return fail(JGPU_E_ARG, "hmm %d: n_states %d invalid", h, ns);
        for (int s = 1; s < ns - 1; ++s) {
            const int gi = m->gmm[h * m->max_states + s];
            if (gi < 0 || gi >= g->n_gmms) return fail(JGPU_E_ARG, "hmm %d state %d: gmm %d out of range", h, s, gi);
        }
    }
    for (int i = 0; i < g->n_gmms; ++i)
        if (g->n_comps[i] < 1 || g->n_comps[i] > g->max_comps) return fail(JGPU_E_ARG, "gmm %d: n_comps invalid", i);
    return JGPU_OK;
}

int build_tables(jgpu_handle* h, const JgpuNet* n, const JgpuHmm* m, const JgpuGmm* g)
{
    Dev& d = h->d;
    const int A = n->n_arcs, NS = n->n_states, H = m->n_hmms, M = m->max_states;
    const int S = M <= 5 ? 5 : 8;
    h->S = S;
    d.S = S;
    d.n_arcs = A; d.n_states = NS; d.init_state = n->init_state; d.n_hmms = H; d.n_gmms = g->n_gmms;

    // network.  pass-through arcs = epsilon input (:533-540) or tee model (:584-600)
    bool any_tee = false;
    for (int i = 0; i < H; ++i) any_tee |= m->tee[i] > JG_LZ;
    std::vector<char> pass(A);
    std::vector<float> tee_in(A, JG_LZ);                 // in the caller's arc order
    for (int a = 0; a < A; ++a) {
        const int in = n->arc_in[a];
        if (in > 0) tee_in[a] = m->tee[in - 1];
        pass[a] = (in == 0) || (tee_in[a] > JG_LZ);
    }
    std::vector<int> topo;
    const int depth = passthrough_depth(n, pass, &topo);
    if (depth < 0) return fail(JGPU_E_ARG, "network has a cycle of epsilon/tee arcs (the reference would recurse forever)");
    if (depth + 1 > JG_MAX_ROUNDS) return fail(JGPU_E_ARG, "epsilon/tee chains of depth %d exceed %d rounds", depth, JG_MAX_ROUNDS);
    d.n_rounds = depth + 1;
    // Which states can receive MORE THAN ONE arrival in a frame?  An arc delivers at most one token per frame
    // (one instance per arc; a tee model's arc delivers its exit token and its pass-through token), unless it is
    // a pass-through arc out of a state that is itself expanded more than once.  Only those states need the
    // per-state max-reduction of arrivals; all others are expanded by their single arrival.
    std::vector<int> n_in(NS, 0);
    n_in[n->init_state] += 1;                            // the utterance seed (recognitionStart :221-226)
    for (int a = 0; a < A; ++a) n_in[n->arc_to[a]] += (n->arc_in[a] > 0 && tee_in[a] > JG_LZ) ? 2 : 1;
    std::vector<char> multi(NS, 0);
    for (int s = 0; s < NS; ++s) multi[s] = n_in[s] > 1;
    for (int s : topo)                                   // topological order of the pass-through sub-graph
        if (multi[s])
            for (int b = n->state_first[s]; b < n->state_first[s] + n->state_narcs[s]; ++b)
                if (pass[b]) multi[n->arc_to[b]] = 1;
    d.init_multi = multi[n->init_state] ? JG_MULTI : 0u;
    // Internal state numbering: multi-arrival states first, so that the per-lane state_key table only spans them
    // (a dense table over all states is 3.5 MB per lane on c3 and every touch of it a DRAM miss).  State ids never
    // leave the engine.
    std::vector<int> new_id(NS);
    int n_multi = 0;
    for (int s = 0; s < NS; ++s) if (multi[s]) new_id[s] = n_multi++;
    {
        int k = n_multi;
        for (int s = 0; s < NS; ++s) if (!multi[s]) new_id[s] = k++;
    }
    d.n_multi = std::max(n_multi, 1);
    d.init_state = new_id[n->init_state];
    // Which states have work for the expansion rounds at all?  Final states (best final token, :513-520) and states
    // with epsilon / tee out-arcs.  An exit token reaching any other state skips the rounds: if the state has a single
    // arrival per frame its word-boundary record is written inside k_internal, otherwise by the commit for the arrival
    // that owns the state in the end.
    std::vector<char> round_work(NS, 0);
    for (int s = 0; s < NS; ++s) {
        round_work[s] = n->state_final[s] > JG_LZ;        // (a MULTI state without either is resolved by the commit alone)
        for (int b = n->state_first[s]; b < n->state_first[s] + n->state_narcs[s] && !round_work[s]; ++b) round_work[s] = pass[b];
    }
    // every row is re-ordered [epsilon | tee-model | other model arcs] (stable), so that the expansion rounds
    // walk a prefix of the row and the commit walks the rest; arc ids are internal to the engine.
    std::vector<int4> arcs(A);
    std::vector<float> arc_tee(A, JG_LZ);
    std::vector<int4> states(NS);
    int max_deg = 0, n_huge_states = 0;
    for (int s = 0; s < NS; ++s) {
        const int f = n->state_first[s], k = n->state_narcs[s];
        int pos = f, n_eps = 0, n_tee = 0;
        for (int cls = 0; cls < 3; ++cls)
            for (int b = f; b < f + k; ++b) {
                const int in = n->arc_in[b];
                const int c_b = in == 0 ? 0 : (tee_in[b] > JG_LZ ? 1 : 2);
                if (c_b != cls) continue;
                int wbits;
                memcpy(&wbits, &n->arc_weight[b], 4);
                const int to = n->arc_to[b];
                arcs[pos] = make_int4(new_id[to] | (multi[to] ? (int)JG_MULTI : 0) | (round_work[to] ? (int)JG_ROUND : 0), wbits, in, n->arc_out[b]);
                arc_tee[pos] = tee_in[b];
                ++pos;
                n_eps += cls == 0; n_tee += cls == 1;
            }
        if (n_eps > 0xffff || n_tee > 0xffff)
            return fail(JGPU_E_ARG, "state %d has %d epsilon / %d tee out-arcs (limit 65535 each)", s, n_eps, n_tee);
        int fbits;
        memcpy(&fbits, &n->state_final[s], 4);
        states[new_id[s]] = make_int4(f, k, fbits, n_eps | (n_tee << 16));
        max_deg = std::max(max_deg, k - n_eps);
        n_huge_states += (k - n_eps) >= JG_HUGE_DEG;
    }
    h->max_deg = max_deg;
    h->n_huge_states = n_huge_states;

    // HMM classes: deduplicate (nst, trP, SEIndex)
    std::map<std::string, int> cls_of;
    std::vector<float> trp;
    std::vector<int2> se;
    std::vector<int> info((size_t)H * 8, -1);
    std::vector<float4> lrtab;
    for (int i = 0; i < H; ++i) {
        const int ns = m->n_states[i];
        std::vector<float> t((size_t)S * S, JG_LZ);
        std::vector<int2> e(S, make_int2(0, 0));
        for (int a = 0; a < ns; ++a) {
            for (int b = 0; b < ns; ++b) t[a * S + b] = m->trp[((size_t)i * M + a) * M + b];
            e[a] = make_int2(m->se[((size_t)i * M + a) * 2], m->se[((size_t)i * M + a) * 2 + 1]);
        }
        std::string key((const char*)&ns, 4);
        key.append((const char*)t.data(), t.size() * 4);
        key.append((const char*)e.data(), e.size() * 8);
        // plain left-to-right class: SEIndex[j] = [j-1, j+1) for emitting j, [N-2, N-1) for the exit
        bool is_lr = S == 5 && ns >= 3 && ns <= 5;
        for (int j = 1; j < ns - 1 && is_lr; ++j) is_lr = e[j].x == j - 1 && e[j].y == j + 1;
        if (is_lr) is_lr = e[ns - 1].x == ns - 2 && e[ns - 1].y == ns - 1;
        auto it = cls_of.find(key);
        int cls;
        if (it == cls_of.end()) {
            cls = (int)cls_of.size();
            cls_of.emplace(key, cls);
            trp.insert(trp.end(), t.begin(), t.end());
            se.insert(se.end(), e.begin(), e.end());
            float c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (is_lr) {
                for (int j = 1; j < ns - 1; ++j) { c[2 * (j - 1)] = t[(j - 1) * S + j]; c[2 * (j - 1) + 1] = t[j * S + j]; }
                c[2 * (ns - 2)] = t[(ns - 2) * S + (ns - 1)];
            }
            lrtab.push_back(make_float4(c[0], c[1], c[2], c[3]));
            lrtab.push_back(make_float4(c[4], c[5], c[6], c[7]));
        } else {
            cls = it->second;
        }
        int teebits;
        memcpy(&teebits, &m->tee[i], 4);
        info[(size_t)i * 8 + 0] = ns | (cls << 8) | (is_lr ? JG_LR_CLASS : 0);
        info[(size_t)i * 8 + 4] = teebits;                // {nst | class, gmm 1..3 | tee, gmm 4..6}: <= 5-state HMMs need one int4
        for (int s = 1; s < ns - 1; ++s) info[(size_t)i * 8 + (s <= 3 ? s : s + 1)] = m->gmm[i * M + s];
    }

    // lazy scoring: the GMM an entry token makes ask for its score — the first emitting state when that is the entry
    // state's only emitting successor, else (-1 - hmm): the commit then stamps every emitting state of the model
    std::vector<int> arc_g1(A, 0);
    {
        std::vector<int> g1_of(H);
        for (int i = 0; i < H; ++i) {
            const int ns = m->n_states[i];
            int n_succ = 0, only = -1;
            for (int j = 1; j < ns - 1; ++j)
                if (m->trp[((size_t)i * M + 0) * M + j] > JG_LZ) { ++n_succ; only = j; }
            g1_of[i] = (n_succ == 1 && only == 1) ? m->gmm[i * M + 1] : (-1 - i);
            if (ns <= 2) g1_of[i] = -1 - i;              // no emitting state at all: nothing to stamp
        }
        for (int a = 0; a < A; ++a) arc_g1[a] = arcs[a].z > 0 ? g1_of[arcs[a].z - 1] : 0;
    }
    int rc;
    {
        int* d_g1 = nullptr;
        if ((rc = upload(h, &d_g1, arc_g1))) return rc;
        d.arc_g1 = d_g1;
    }
    // The static network + HMM tables live in ONE allocation (an L2 access-policy window over it was measured:
    // pinning them costs more L2 than it saves; the streams carry evict-first hints instead).
    auto pad256 = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t b_arcs = pad256(arcs.size() * sizeof(int4)), b_states = pad256(states.size() * sizeof(int4));
    const size_t b_tee = any_tee ? pad256(arc_tee.size() * sizeof(float)) : 0, b_info = pad256(info.size() * sizeof(int));
    const size_t b_trp = pad256(trp.size() * sizeof(float)), b_se = pad256(se.size() * sizeof(int2));
    // k_internal keeps the left-to-right constants in shared memory: a model set with more classes than fit there
    // takes the generic SEIndex / trP path for all of them
    if ((int)lrtab.size() > JG_LR_SH)
        for (int i = 0; i < H; ++i) info[(size_t)i * 8 + 0] &= ~JG_LR_CLASS;
    const size_t b_lr = pad256(lrtab.size() * sizeof(float4));
    // 8-byte form of hmm_info for k_internal<5> (see Dev::hmm8)
    std::vector<uint2> info8;
    {
        const char* off8 = getenv("JUICER_B200_HMM8");
        bool fits = S == 5 && g->n_gmms < 65536 && cls_of.size() < 4096 && !(off8 && atoi(off8) == 0);
        if (fits) {
            info8.resize(H);
            for (int i = 0; i < H; ++i) {
                const int* e = &info[(size_t)i * 8];
                const unsigned ns = e[0] & 0xff, cls = (unsigned)(e[0] & ~JG_LR_CLASS) >> 8, lr = (e[0] & JG_LR_CLASS) ? 1u : 0u;
                info8[i] = make_uint2(ns | (lr << 3) | (cls << 4) | ((unsigned)e[1] << 16), (unsigned)e[2] | ((unsigned)e[3] << 16));
            }
        }
    }
    const size_t b_info8 = pad256(info8.size() * sizeof(uint2));
    h->static_bytes = b_arcs + b_states + b_tee + b_info + b_trp + b_se + b_lr + b_info8;
    char* base = nullptr;
    if ((rc = h->alloc(&base, h->static_bytes, false))) return rc;
    h->static_base = base;
    size_t off = 0;
    auto put = [&](const void* src, size_t bytes, size_t padded) -> char* {
        char* p = base + off;
        if (bytes) cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, h->stream);
        off += padded;
        return p;
    };
    d.arcs = (const int4*)put(arcs.data(), arcs.size() * sizeof(int4), b_arcs);
    d.states = (const int4*)put(states.data(), states.size() * sizeof(int4), b_states);
    d.arc_tee = any_tee ? (const float*)put(arc_tee.data(), arc_tee.size() * sizeof(float), b_tee) : nullptr;
    int* d_info = (int*)put(info.data(), info.size() * sizeof(int), b_info);
    float* d_trp = (float*)put(trp.data(), trp.size() * sizeof(float), b_trp);
    int2* d_se = (int2*)put(se.data(), se.size() * sizeof(int2), b_se);
    d.lr = (const float4*)put(lrtab.data(), lrtab.size() * sizeof(float4), b_lr);
    d.n_lr = (int)lrtab.size();
    d.hmm8 = info8.empty() ? nullptr : (const uint2*)put(info8.data(), info8.size() * sizeof(uint2), b_info8);
    CK(cudaGetLastError());
    d.hmm_info = d_info; d.trp = d_trp; d.se = d_se;

    // GMM parameters, transposed [d][c][g], zero padded
    GmmDev& G = h->g;
    const int C = g->max_comps, D = g->dim;
    const int DP = (D + 3) & ~3;
    h->DP = DP <= 16 ? 16 : DP <= 28 ? 28 : DP <= 40 ? 40 : DP <= 52 ? 52 : 64;
    h->dim = D;
    G.n_gmms = g->n_gmms; G.C = C; G.D = D; G.gpb = std::max(1, 256 / C);
    G.g_pad = (g->n_gmms + G.gpb - 1) / G.gpb * G.gpb;
    const size_t plane = (size_t)C * G.g_pad;
    std::vector<float> mu((size_t)h->DP * plane, 0.0f), iv((size_t)h->DP * plane, 0.0f), det(plane, JG_LZ);
    std::vector<int> nc(G.g_pad, 0);
    for (int gi = 0; gi < g->n_gmms; ++gi) {
        nc[gi] = g->n_comps[gi];
        for (int c = 0; c < g->n_comps[gi]; ++c) {
            det[(size_t)c * G.g_pad + gi] = g->dets[(size_t)gi * C + c];
            for (int dd = 0; dd < D; ++dd) {
                mu[(size_t)dd * plane + (size_t)c * G.g_pad + gi] = g->means[((size_t)gi * C + c) * D + dd];
                iv[(size_t)dd * plane + (size_t)c * G.g_pad + gi] = g->ivars[((size_t)gi * C + c) * D + dd];
            }
        }
    }
    float *d_mu, *d_iv, *d_det; int* d_nc;
    if ((rc = upload(h, &d_mu, mu))) return rc;
    if ((rc = upload(h, &d_iv, iv))) return rc;
    if ((rc = upload(h, &d_det, det))) return rc;
    if ((rc = upload(h, &d_nc, nc))) return rc;
    G.mu = d_mu; G.iv = d_iv; G.det = d_det; G.ncomp = d_nc;
    if (const char* e = getenv("JUICER_B200_LAZY")) h->lazy = atoi(e) != 0;
    if (const char* e = getenv("JUICER_B200_LAZY_CLUSTER")) h->lazy_cluster = atoi(e);
    if (h->lazy_cluster != 1 && h->lazy_cluster != 2 && h->lazy_cluster != 4) h->lazy_cluster = JG_LAZY_CLUSTER;
    if (C > 32) h->lazy = false;                         // a component per thread of a warp at most
    if (h->lazy) {
        // the same parameters once more, one 16 B-aligned row of DP floats per component: [gmm][component][DP]
        const int DPl = h->DP;
        std::vector<float> mu2((size_t)g->n_gmms * C * DPl, 0.0f), iv2(mu2.size(), 0.0f), det2((size_t)g->n_gmms * C, JG_LZ);
        for (int gi = 0; gi < g->n_gmms; ++gi)
            for (int c = 0; c < g->n_comps[gi]; ++c) {
                det2[(size_t)gi * C + c] = g->dets[(size_t)gi * C + c];
                for (int dd = 0; dd < D; ++dd) {
                    mu2[((size_t)gi * C + c) * DPl + dd] = g->means[((size_t)gi * C + c) * D + dd];
                    iv2[((size_t)gi * C + c) * DPl + dd] = g->ivars[((size_t)gi * C + c) * D + dd];
                }
            }
        float *d_mu2, *d_iv2, *d_det2;
        if ((rc = upload(h, &d_mu2, mu2))) return rc;
        if ((rc = upload(h, &d_iv2, iv2))) return rc;
        if ((rc = upload(h, &d_det2, det2))) return rc;
        LazyArgs& z = h->lz;
        z.mu = d_mu2; z.iv = d_iv2; z.det = d_det2; z.ncomp = d_nc;
        z.n_gmms = g->n_gmms; z.C = C;
        z.cpw = 1;
        while (z.cpw < C) z.cpw *= 2;
    }
    {
        std::vector<double> sp(&JG_SOFTPLUS_TABLE[0][0], &JG_SOFTPLUS_TABLE[0][0] + JG_SP_INTERVALS * (JG_SP_DEG + 1));
        double* d_sp;
        if ((rc = upload(h, &d_sp, sp))) return rc;
        G.softplus = d_sp;
        h->lz.softplus = d_sp;
    }
    return JGPU_OK;
}

static int bits_for(unsigned long long v)     // smallest b with v < 2^b
{
    int b = 0;
    while ((v >> b) != 0ull) ++b;
    return b;
}

// bytes of per-lane state per instance of capacity: two list buffers (record + S-1 token planes each), the arrival
// records (two planes) and the round-0 work list at 2 per instance
static double bytes_per_instance(int S) { return 16.0 * 2 + 16.0 * 2 * (S - 1) + 2.0 + 2 * (32.0 + 4.0); }

int build_state(jgpu_handle* h)
{
    Dev& d = h->d;
    const JgpuCfg& c = h->cfg;
    const size_t L = c.n_lanes;
    d.start_beam = c.start_beam; d.main_beam = c.main_beam; d.end_beam = c.end_beam; d.word_beam = c.word_beam;
    d.max_hyps = c.max_hyps;
    d.hist_min = 0; d.hist_max = 0; d.hist_nbins = 1;
    if (c.max_hyps > 0) {                                 // src/WFSTDecoderLite.cpp:76-82, src/Histogram.cpp:29-37
        const float mn = c.main_beam > 0.0 ? (float)(-c.main_beam - 800.0) : (float)-1000.0;
        d.hist_min = (int)(mn - 1.0);
        d.hist_max = (int)((float)200.0 + 1.0);
        d.hist_nbins = d.hist_max - d.hist_min + 1;
    }
    d.n_lanes = c.n_lanes;
    // An arc holds at most one instance (attachNetInst, src/WFSTDecoderLite.cpp:751-774), so n_arcs + 1 is all a lane
    // can ever need — the reference simply keeps allocating.  The lists of all lanes get up to 30 % of the free
    // memory; when that is less than a full-size lane, an utterance that overflows its lane is decoded again in a
    // second pass that views the same pools as fewer lanes with larger arenas (decode_common).
    h->cap_full = d.n_arcs + 1;
    if (h->cap_full >= (1 << 28)) return fail(JGPU_E_ARG, "network too large: %d arcs", d.n_arcs);
    size_t free_b0 = 0, total_b0 = 0;
    CK(cudaMemGetInfo(&free_b0, &total_b0));
    if (c.max_active > 0) {
        d.cap = std::min(c.max_active, h->cap_full);
    } else {
        const double per_lane = 0.30 * (double)free_b0 / (double)L / bytes_per_instance(d.S);
        d.cap = (int)std::min<double>(h->cap_full, std::max(per_lane, 1024.0));
    }
    d.cap = std::max(d.cap, 64);
    d.cap_arr = 2 * d.cap + 1024;
    d.cap_paths = c.max_paths > 0 ? c.max_paths : (1 << 21);   // re-sized from free memory below when 0
    d.cap_huge = std::max(h->n_huge_states, 1);          // a state is committed at most once per frame
    d.max_frames = c.max_frames > 0 ? c.max_frames : 4096;
    d.frame_stats = c.frame_stats;
    d.huge_deg = JG_HUGE_DEG;
    d.fuse_exits = !(c.end_beam > 0.0f) && !(c.word_beam > 0.0f);
    // field widths of the stamped tables: positions up to a full-size lane, arrival ids up to 2 * (n_arcs + 1) + 1
    d.slot_bits = bits_for((unsigned long long)h->cap_full + 1ull);
    d.key_id_bits = bits_for(2ull * ((unsigned long long)d.n_arcs + 1ull) + 1ull);
    d.slot_emask = (1u << std::min(32 - d.slot_bits, 16)) - 1u;
    d.key_emask = (1u << std::min(32 - d.key_id_bits, 16)) - 1u;
    h->epoch_wrap = std::min(d.slot_emask, d.key_emask);    // both tables are wiped when the narrower stamp wraps
    if (h->epoch_wrap < 7u) return fail(JGPU_E_ARG, "network too large for the stamped tables: %d arcs", d.n_arcs);
    h->has_huge = h->n_huge_states > 0;
    h->bpl = std::max(2, std::min(64, (1184 * (256 / JG_THREADS) + c.n_lanes - 1) / c.n_lanes));   // k_commit_huge only
    {
        int n_sm = 148;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->device);
        d.grid_internal = n_sm * (h->S <= 5 ? JG_INT_CTAS : 2);   // __launch_bounds__ of k_internal
        d.grid_other = n_sm * 6;                       // __launch_bounds__(256, 6): k_filter
        d.grid_walk = n_sm * 6;                        // __launch_bounds__(256, 6): k_walk
    }

    const size_t cap = d.cap, P = d.S - 1;
    size_t need = L * (2 * cap * 16 + 2 * P * cap * 16 + (size_t)d.n_arcs * 4 + (size_t)d.n_multi * 8 +
                       (size_t)d.cap_arr * 36 + (size_t)d.cap_paths * 36);
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    if (c.max_paths <= 0) {
        // word-boundary arena + its free list: 4/5 of what is left (1M .. 32M records per lane).  An arena that
        // never fills to 3/4 is never garbage-collected, which keeps allocation sequential.
        const size_t fixed = need - L * (size_t)d.cap_paths * 36;
        const size_t budget = free_b > fixed + (2ull << 30) ? (free_b - fixed - (2ull << 30)) / 5 * 4 : 0;
        size_t per_lane = budget / (L * 36);
        per_lane = std::min<size_t>(std::max<size_t>(per_lane, 1u << 20), 1u << 25);
        d.cap_paths = (int)per_lane;
        need = fixed + L * (size_t)d.cap_paths * 36;
    }
    if (need + (1ull << 30) > free_b)
        return fail(JGPU_E_CAPACITY, "decoder state needs %.1f GB for %d lanes but only %.1f GB of device memory is free",
                    need / 1e9, c.n_lanes, free_b / 1e9);
    h->base = jgpu_handle::View{c.n_lanes, d.cap, d.cap_arr, d.cap_paths};
    h->pool_inst = L * cap; h->pool_arr = L * (size_t)d.cap_arr; h->pool_paths = L * (size_t)d.cap_paths;
    int rc;
    if ((rc = h->alloc(&d.ctl, L))) return rc;
    if ((rc = h->alloc(&d.inst_meta, 2 * h->pool_inst, false))) return rc;
    if ((rc = h->alloc(&d.tok, 2 * P * h->pool_inst, false))) return rc;
    if ((rc = h->alloc(&d.live, 2 * h->pool_inst))) return rc;
    if ((rc = h->alloc(&d.slotmap, L * d.n_arcs))) return rc;
    if ((rc = h->alloc(&d.state_key, L * d.n_multi))) return rc;
    if ((rc = h->alloc(&d.arr_tok, h->pool_arr, false))) return rc;
    if ((rc = h->alloc(&d.arr_meta, h->pool_arr, false))) return rc;
    if ((rc = h->alloc(&d.huge, L * d.cap_huge, false))) return rc;
    if ((rc = h->alloc(&d.r0_list, h->pool_arr, false))) return rc;
    if ((rc = h->alloc(&d.paths, h->pool_paths, false))) return rc;
    if ((rc = h->alloc(&d.path_free, h->pool_paths, false))) return rc;
    d.gc_threshold = d.cap_paths - d.cap_paths / 4;           // collect when 3/4 full: recycled slots are scattered, so
                                                              // an arena that is big enough is never collected at all
    h->gc_period = d.cap_paths < (1 << 20) ? 16 : 64;       // small arenas (tests, tight memory) are collected more often
    if (const char* e = getenv("JUICER_B200_GC_PERIOD")) h->gc_period = atoi(e);
    if ((rc = h->alloc(&d.hist, L * d.hist_nbins))) return rc;
    if ((rc = h->alloc(&d.fstat_cnt, d.frame_stats ? L * d.max_frames * 4 : 1))) return rc;
    if ((rc = h->alloc(&d.fstat_best, d.frame_stats ? L * d.max_frames : 1))) return rc;
    // schedule + score ring + streaming feature staging
    if ((rc = h->alloc(&h->d_sched, (size_t)(h->sched_chunk + 1) * L, false))) return rc;
    if ((rc = h->alloc(&h->d_rows, (size_t)h->sched_chunk * L, false))) return rc;
    if ((rc = h->alloc(&d.lane_step, L))) return rc;
    d.lazy = h->lazy ? 1 : 0;
    d.need_stride = ((int)L + 31) & ~31;
    d.need_gp = (d.n_gmms + 31) & ~31;
    d.xtile_dp = h->DP;
    d.feat_dim = h->dim;
    if ((rc = h->alloc(&h->d_feat_base, 1))) return rc;
    d.feat_base = h->d_feat_base;
    if (h->lazy) {
        if ((rc = h->alloc(&d.need, (size_t)d.need_gp * d.need_stride))) return rc;
        if (d.frame_stats && (rc = h->alloc(&d.scored, (size_t)d.need_gp * d.need_stride))) return rc;
        if ((rc = h->alloc(&d.gmm_next, 1))) return rc;
        if ((rc = h->alloc(&d.lane_stamp, (size_t)d.need_stride))) return rc;
        if ((rc = h->alloc(&d.xtile, (size_t)d.need_stride * h->DP))) return rc;
        std::vector<int> st0(d.need_stride, 0x100);
        CK(cudaMemcpyAsync(d.lane_stamp, st0.data(), st0.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if ((rc = h->alloc(&h->d_scores, L * d.n_gmms, false))) return rc;                    // one row per lane
        LazyArgs& z = h->lz;
        z.need = d.need; z.scored = d.scored; z.gmm_next = d.gmm_next; z.need_stride = d.need_stride; z.need_gp = d.need_gp;
        z.n_lanes = (int)L; z.ctl = d.ctl; z.lane_stamp = d.lane_stamp;
        z.xtile = d.xtile; z.scores = h->d_scores;
    } else if ((rc = h->alloc(&h->d_scores, (size_t)2 * h->FB * L * d.n_gmms, false))) return rc;   // two halves
    if ((rc = h->alloc(&h->d_stream_feats, L * h->stream_chunk * h->dim, false))) return rc;
    if ((rc = h->alloc(&h->d_gmm_out, (size_t)h->gmm_chunk * d.n_gmms, false))) return rc;
    d.scores = h->d_scores;
    d.sched = h->d_sched;
    // results: at least one header per lane for the streaming interface; the words of a batch share one pool
    h->res_cap = std::max<size_t>(L, 64);
    h->words_cap = std::max<size_t>(h->res_cap * 64, 1u << 16);
    if (const char* e = getenv("JUICER_B200_WORD_POOL")) h->words_cap = (size_t)std::max(1, atoi(e));
    if ((rc = h->alloc(&d.res_hdr, h->res_cap))) return rc;
    if ((rc = h->alloc(&d.res_words, h->words_cap))) return rc;
    if ((rc = h->alloc(&d.res_used, 1))) return rc;
    d.res_words_cap = (int)h->words_cap;
    {
        const int smem = (int)internal_smem_bytes(h->S, (int)L);
        cudaError_t e = cudaSuccess;
        void (*kerns[8])(Dev) = {k_internal<5, true, false>, k_internal<5, false, false>, k_internal<5, true, true>, k_internal<5, false, true>,
                                 k_internal<8, true, false>, k_internal<8, false, false>, k_internal<8, true, true>, k_internal<8, false, true>};
        for (int i = 0; i < 8 && e == cudaSuccess; ++i) e = cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        // Shared-memory carve-out: exactly what the resident CTAs need and no more — what is left of the 256 KB is the
        // L1 that the per-instance hmm_info / score gathers go through (measured on c3: 233.8 us per launch at 72 %,
        // 237.0 at the driver's default choice, 286.6 at 100 %).  JUICER_B200_INT_CARVEOUT overrides (percent).
        for (int i = 0; i < 8 && e == cudaSuccess; ++i) {
            cudaFuncAttributes fa;
            int occ = 0;
            if (cudaFuncGetAttributes(&fa, kerns[i]) != cudaSuccess ||
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kerns[i], JG_THREADS, smem) != cudaSuccess || occ < 1) { cudaGetLastError(); continue; }
            int pct = (int)std::min<size_t>(100, ((size_t)occ * (smem + fa.sharedSizeBytes + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
            if (const char* cv = getenv("JUICER_B200_INT_CARVEOUT")) pct = atoi(cv);
            e = cudaFuncSetAttribute(kerns[i], cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        }
        if (e != cudaSuccess) return fail(JGPU_E_CUDA, "k_internal shared memory opt-in: %s", cudaGetErrorString(e));
    }
    {
        // grids = SM count x the CTAs that are really resident (a few hundred bytes of shared memory more per
        // CTA can cost a whole CTA per SM, and a grid that no longer fits in one wave costs a second set-up)
        int n_sm = 148, occ = 0;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->device);
        const size_t smem_int = internal_smem_bytes(h->S, (int)L);
        cudaError_t e = h->S == 5 ? (h->lazy ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_internal<5, true, true>, JG_THREADS, smem_int)
                                             : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_internal<5, true, false>, JG_THREADS, smem_int))
                                  : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_internal<8, true, false>, JG_THREADS, smem_int);
        if (e == cudaSuccess && occ > 0) d.grid_internal = n_sm * occ;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_walk<1, false>, JG_THREADS, 0) == cudaSuccess && occ > 0) d.grid_walk = n_sm * occ;
        if (const char* cv = getenv("JUICER_B200_WALK_CARVEOUT")) {   // tuning: preferred shared-memory carve-out of the walk kernels, percent
            cudaFuncSetAttribute(k_walk<0, false>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(cv));
            cudaFuncSetAttribute(k_walk<1, false>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(cv));
            cudaFuncSetAttribute(k_commit_huge<false>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(cv));
        }
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_filter, JG_THREADS, 0) == cudaSuccess && occ > 0) d.grid_other = n_sm * occ;
        cudaGetLastError();
    }
    h->lanes.assign(L, LaneHost());
    h->host_epoch.assign(L, 0u);
    return JGPU_OK;
}

void drop_graphs(jgpu_handle* h)
{
    if (h->graph_step) cudaGraphExecDestroy(h->graph_step);
    if (h->graph_block) cudaGraphExecDestroy(h->graph_block);
    h->graph_step = h->graph_block = nullptr;
}

// room for n utterance headers and n_words result words (the pool only grows)
int ensure_results(jgpu_handle* h, size_t n, size_t n_words)
{
    Dev& d = h->d;
    if (n <= h->res_cap && n_words <= h->words_cap) return JGPU_OK;
    CK(cudaStreamSynchronize(h->stream));
    drop_graphs(h);                                       // Dev (kernel argument) changes
    int rc;
    if (n > h->res_cap) {
        h->release(d.res_hdr);
        d.res_hdr = nullptr;
        h->res_cap = n;
        if ((rc = h->alloc(&d.res_hdr, n))) return rc;
    }
    if (n_words > h->words_cap) {
        if (n_words > (size_t)1 << 30) return fail(JGPU_E_CAPACITY, "result word pool of %zu records", n_words);
        h->release(d.res_words);
        d.res_words = nullptr;
        h->words_cap = n_words;
        if ((rc = h->alloc(&d.res_words, n_words))) return rc;
        d.res_words_cap = (int)n_words;
    }
    return JGPU_OK;
}

// Looks at the pools as `v.n_lanes` lanes with arenas of v.cap instances / v.cap_arr arrivals / v.cap_paths
// word-boundary records.  Per-lane tables that do not depend on the arena sizes (slotmap, state keys, control
// blocks, histograms) keep their stride; their stamps stay valid because a lane's epoch never goes back.
int set_view(jgpu_handle* h, const jgpu_handle::View& v)
{
    Dev& d = h->d;
    if (d.n_lanes == v.n_lanes && d.cap == v.cap && d.cap_arr == v.cap_arr && d.cap_paths == v.cap_paths) return JGPU_OK;
    CK(cudaStreamSynchronize(h->stream));
    drop_graphs(h);
    d.n_lanes = v.n_lanes; d.cap = v.cap; d.cap_arr = v.cap_arr; d.cap_paths = v.cap_paths;
    h->lz.n_lanes = v.n_lanes;
    d.gc_threshold = d.cap_paths - d.cap_paths / 4;
    h->bpl = std::max(2, std::min(64, (1184 * (256 / JG_THREADS) + v.n_lanes - 1) / v.n_lanes));
    return JGPU_OK;
}

// The view for a second pass over `n_failed` utterances whose lanes overflowed (error bits or-ed in `errors`):
// four times the arena that was too small (at most what a lane can ever hold), on as many lanes as the pools
// then have room for.  Returns false when nothing can grow any more.
bool grow_view(const jgpu_handle* h, int errors, int n_failed, jgpu_handle::View* out)
{
    const Dev& d = h->d;
    jgpu_handle::View v{d.n_lanes, d.cap, d.cap_arr, d.cap_paths};
    long long cap = v.cap;
    if (errors & (JG_ERR_ACTIVE | JG_ERR_ARRIVALS)) cap = std::min<long long>(h->cap_full, 4ll * cap);
    long long lanes = std::min<long long>(v.n_lanes, std::max(1, n_failed));
    if (errors & JG_ERR_PATHS) lanes = std::max(1ll, std::min<long long>(lanes, v.n_lanes / 4));
    lanes = std::min<long long>(lanes, (long long)(h->pool_inst / (size_t)cap));
    lanes = std::min<long long>(lanes, (long long)(h->pool_arr / (size_t)(2 * cap + 1024)));
    if (lanes < 1) {                                      // not even one lane of that size: one lane takes the whole pool
        lanes = 1;
        cap = std::min<long long>((long long)h->pool_inst, ((long long)h->pool_arr - 1024) / 2);
    }
    v.n_lanes = (int)lanes;
    v.cap = (int)cap;
    v.cap_arr = (int)(2 * cap + 1024);
    v.cap_paths = (int)std::min<size_t>(h->pool_paths / (size_t)lanes, (size_t)1 << 30);
    const bool grew = v.cap > d.cap || v.cap_paths > d.cap_paths;
    *out = v;
    return grew;
}

int launch_gmm(jgpu_handle* h, const float* d_x, const int* d_rows, int n_rows, float* d_out, long long out_base)
{
    if (n_rows <= 0) return JGPU_OK;
    cudaStream_t st = h->stream;
    const GmmDev& G = h->g;
    const int gpb = G.gpb;
    const int n_bx = (G.n_gmms + gpb - 1) / gpb, n_by = (n_rows + JG_GMM_RT - 1) / JG_GMM_RT;
    const size_t smem = gmm_smem_bytes(JG_GMM_RT, h->DP, G.C, gpb);
    h->prof_begin(JGPU_K_GMM, st);
    switch (h->DP) {
#define GMM_CASE(DPV)                                                                                          \
    case DPV:                                                                                                  \
        CK(cudaFuncSetAttribute(k_gmm_scores<DPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
        k_gmm_scores<DPV><<<dim3(n_bx, n_by), 256, smem, st>>>(G, gpb, d_x, d_rows, n_rows, d_out, out_base, n_by); \
        break;
        GMM_CASE(16) GMM_CASE(28) GMM_CASE(40) GMM_CASE(52) GMM_CASE(64)
#undef GMM_CASE
    default: return fail(JGPU_E_ARG, "unsupported padded dim %d", h->DP);
    }
    h->prof_end(st);
    ++h->launches;
    CK(cudaGetLastError());
    return JGPU_OK;
}

// the stamped (GMM, lane) pairs of this step (see k_gmm_lazy); runs between k_boundary and k_internal
int launch_lazy(jgpu_handle* h)
{
    const LazyArgs& z = h->lz;
    const int CL = h->lazy_cluster;
    // persistent: as many CTAs as are resident (2 per SM up to DP 40), warps take GMMs from a counter
    int grid = std::min((z.n_gmms + JG_LAZY_WARPS - 1) / JG_LAZY_WARPS, h->n_sm * (h->DP <= 40 ? 2 : 1));
    grid = std::max(CL, grid / CL * CL);
    const size_t smem = (size_t)z.need_stride * h->DP * sizeof(float) + (size_t)JG_LAZY_WARPS * 32 * (z.cpw + 1) * sizeof(float) +
                        (size_t)JG_LAZY_WARPS * z.need_stride * sizeof(unsigned short) + (size_t)2 * z.need_stride * sizeof(int);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(JG_LAZY_WARPS * 32); cfg.dynamicSmemBytes = smem; cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = CL > 1 ? 1 : 0;
    h->prof_begin(JGPU_K_GMM);
    cudaError_t e = cudaErrorInvalidValue;
#define LAZY_LAUNCH(DPV, CLV)                                                                                              \
    if (h->DP == DPV && CL == CLV) {                                                                                       \
        e = cudaFuncSetAttribute(k_gmm_lazy<DPV, CLV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
        if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, k_gmm_lazy<DPV, CLV>, z);                                       \
    }
#define LAZY_DP(DPV) LAZY_LAUNCH(DPV, 1) LAZY_LAUNCH(DPV, 2) LAZY_LAUNCH(DPV, 4)
    LAZY_DP(16) LAZY_DP(28) LAZY_DP(40) LAZY_DP(52) LAZY_DP(64)
#undef LAZY_DP
#undef LAZY_LAUNCH
    h->prof_end();
    ++h->launches;
    if (e != cudaSuccess) return fail(JGPU_E_CUDA, "k_gmm_lazy launch (DP %d, cluster %d, %zu B shared): %s", h->DP, CL, smem, cudaGetErrorString(e));
    return JGPU_OK;
}

int launch_step(jgpu_handle* h)
{
    const Dev& d = h->d;
    const dim3 grid_huge(h->bpl, d.n_lanes);
    cudaStream_t st = h->stream;
#ifdef JG_TRACE
    if (h->d_trace) {
        static const int on = 1, off = 0;
        if (h->abs_step == h->trace_s0) cudaMemcpyToSymbolAsync(g_trace_on, &on, sizeof(int), 0, cudaMemcpyHostToDevice, h->stream);
        if (h->abs_step == h->trace_s1) cudaMemcpyToSymbolAsync(g_trace_on, &off, sizeof(int), 0, cudaMemcpyHostToDevice, h->stream);
        ++h->abs_step;
    }
#endif
    h->prof_begin(JGPU_K_BOUNDARY);
    k_boundary<<<d.n_lanes, 32, 0, st>>>(d);
    h->prof_end();
    if (h->lazy) { int rc = launch_lazy(h); if (rc) return rc; }
    h->prof_begin(JGPU_K_INTERNAL);
    {
        const size_t smem = internal_smem_bytes(h->S, d.n_lanes);   // chunk buffers (record + S-1 token planes) + lane table
        void (*kern)(Dev);
        if (h->S == 5) kern = h->lazy ? (d.fuse_exits ? k_internal<5, true, true> : k_internal<5, false, true>)
                                      : (d.fuse_exits ? k_internal<5, true, false> : k_internal<5, false, false>);
        else kern = h->lazy ? (d.fuse_exits ? k_internal<8, true, true> : k_internal<8, false, true>)
                            : (d.fuse_exits ? k_internal<8, true, false> : k_internal<8, false, false>);
        kern<<<d.grid_internal, JG_THREADS, smem, st>>>(d);
    }
    h->prof_end();
    if (!d.fuse_exits) {
        h->prof_begin(JGPU_K_SEED);
        k_filter<<<d.grid_other, JG_THREADS, 0, st>>>(d);
        h->prof_end();
        ++h->launches;
    }
    for (int r = 0; r < d.n_rounds; ++r) {
        h->prof_begin(r == 0 ? JGPU_K_EXPAND : r == 1 ? JGPU_K_EXPAND_R1 : JGPU_K_EXPAND_R2);
        // The expansion rounds have little to do (a few chunks per lane) and every CTA first reads the control block of
        // every lane: fewer CTAs per SM means less of that fixed cost (JUICER_B200_EXPAND_CTAS: CTAs per SM, rounds 0 / >= 1)
        static const int ex0 = getenv("JUICER_B200_EXPAND_CTAS0") ? atoi(getenv("JUICER_B200_EXPAND_CTAS0")) : 0;
        static const int ex1 = getenv("JUICER_B200_EXPAND_CTAS1") ? atoi(getenv("JUICER_B200_EXPAND_CTAS1")) : 0;
        const int per_sm = r == 0 ? ex0 : ex1;
        const int grid_r = per_sm > 0 ? std::min(d.grid_walk, h->n_sm * per_sm) : d.grid_walk;
        k_walk<0, false><<<grid_r, JG_THREADS, 0, st>>>(d, r);
        h->prof_end();
    }
    h->prof_begin(JGPU_K_COMMIT);
    if (h->lazy) k_walk<1, true><<<d.grid_walk, JG_THREADS, 0, st>>>(d, 0);
    else k_walk<1, false><<<d.grid_walk, JG_THREADS, 0, st>>>(d, 0);
    h->prof_end();
    if (h->has_huge) {
        h->prof_begin(JGPU_K_EXPAND_HUGE);
        if (h->lazy) k_commit_huge<true><<<grid_huge, JG_THREADS, 0, st>>>(d);
        else k_commit_huge<false><<<grid_huge, JG_THREADS, 0, st>>>(d);
        h->prof_end();
        ++h->launches;
    }
    h->launches += 3 + d.n_rounds;
    CK(cudaGetLastError());
    return JGPU_OK;
}

// The kernel sequence of `n` frame steps has no per-step argument (schedule rows, thresholds, list
// sizes and round counts all live in device memory), so it is captured once and replayed as a CUDA
// graph: the launch gaps between the 6-8 small kernels of a step shrink to the graph's node-to-node latency.
int launch_steps_graph(jgpu_handle* h, int n)
{
    cudaGraphExec_t& exec = n == 1 ? h->graph_step : h->graph_block;
    if (!exec) {
        cudaGraph_t g = nullptr;
        CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        const int64_t before = h->launches;
        int rc = JGPU_OK;
        for (int i = 0; i < n && !rc; ++i) rc = launch_step(h);
        h->launches = before;
        cudaError_t e = cudaStreamEndCapture(h->stream, &g);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) return fail(JGPU_E_CUDA, "graph capture: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&exec, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { exec = nullptr; return fail(JGPU_E_CUDA, "graph instantiate: %s", cudaGetErrorString(e)); }
    }
    CK(cudaGraphLaunch(exec, h->stream));
    h->launches += (int64_t)n * (3 + h->d.n_rounds + (h->d.fuse_exits ? 0 : 1) + (h->has_huge ? 1 : 0) + (h->lazy ? 1 : 0));
    return JGPU_OK;
}

// Enqueues one chunk of the lock-step schedule: `ns` frame / seed steps of all lanes (rows {feature row, -, flags,
// result slot}, at most sched_chunk of them) and, when `last`, the trailing close-only row that runs the pending
// finishes.  d_x: feature base.  `chunk` is modified (score-ring rows are filled in).
int submit_chunk(jgpu_handle* h, std::vector<int4>& chunk, int ns, bool last, const float* d_x)
{
    const Dev& d = h->d;
    const int L = d.n_lanes, FB = h->FB;
    if (ns + (last ? 1 : 0) == 0) return JGPU_OK;
    std::vector<int>& rows = h->rows_host;
    rows.resize((size_t)std::max(ns, 1) * L);
    for (int i = 0; i < ns; ++i)
        for (int l = 0; l < L; ++l) {
            int4& e = chunk[(size_t)i * L + l];
            e.y = (((i / FB) & 1) * FB + (i % FB)) * L + l;   // score-ring row of (step, lane): two halves
            rows[(size_t)i * L + l] = ((e.z & 3) == JG_MODE_FRAME) ? e.x : -1;
        }
    CK(cudaMemcpyAsync(h->d_sched, chunk.data(), (size_t)(ns + (last ? 1 : 0)) * L * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->d.lane_step, 0, (size_t)L * sizeof(int), h->stream));
    if (ns && !h->lazy) CK(cudaMemcpyAsync(h->d_rows, rows.data(), (size_t)ns * L * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_feat_base, &d_x, sizeof(const float*), cudaMemcpyHostToDevice, h->stream));   // k_boundary gathers from it
    // dense mode: the acoustic scores of frame block b+1 are enqueued ahead of the search of block b (two halves of the ring)
    auto issue_gmm = [&](int b) -> int {
        if (h->lazy) return JGPU_OK;
        const int b0 = b * FB, nb = std::min(FB, ns - b0), half = b & 1;
        return launch_gmm(h, d_x, h->d_rows + (size_t)b0 * L, nb * L, h->d_scores, (long long)half * FB * L);
    };
    const int n_blocks = (ns + FB - 1) / FB;
    if (n_blocks > 0) { int rc = issue_gmm(0); if (rc) return rc; }
    for (int b = 0; b < n_blocks; ++b) {
        int rc;
        if (b + 1 < n_blocks && (rc = issue_gmm(b + 1))) return rc;
        const int b0 = b * FB, nb = std::min(FB, ns - b0);
        // a lane whose epoch stamp wraps inside this block gets its stamped tables wiped right before that
        // step (every epoch_wrap + 1 steps: 2048 on a 2M-arc network); such a block is launched step by step
        bool wipe_in_block = false;
        {
            std::vector<unsigned> ep(h->host_epoch);
            for (int i = b0; i < b0 + nb && !wipe_in_block; ++i)
                for (int l = 0; l < L; ++l)
                    if ((chunk[(size_t)i * L + l].z & 3) != JG_MODE_IDLE && ((++ep[l]) & h->epoch_wrap) == 0u) { wipe_in_block = true; break; }
        }
        const bool graphs = h->use_graphs && !h->prof_on && !JG_TRACING(h);
        if (graphs && nb == FB && !wipe_in_block) {
            for (int i = b0; i < b0 + nb; ++i)
                for (int l = 0; l < L; ++l)
                    if ((chunk[(size_t)i * L + l].z & 3) != JG_MODE_IDLE) ++h->host_epoch[l];
            if ((rc = launch_steps_graph(h, nb))) return rc;
        } else {
            for (int i = b0; i < b0 + nb; ++i) {
                for (int l = 0; l < L; ++l) {
                    if ((chunk[(size_t)i * L + l].z & 3) == JG_MODE_IDLE) continue;
                    if (((++h->host_epoch[l]) & h->epoch_wrap) == 0u) {
                        CK(cudaMemsetAsync(h->d.state_key + (size_t)l * d.n_multi, 0, (size_t)d.n_multi * sizeof(u64), h->stream));
                        CK(cudaMemsetAsync(h->d.slotmap + (size_t)l * d.n_arcs, 0, (size_t)d.n_arcs * sizeof(unsigned), h->stream));
                    }
                }
                if ((rc = graphs ? launch_steps_graph(h, 1) : launch_step(h))) return rc;
            }
        }
        // word-boundary arena: mark + sweep between two frame steps, every gc_period steps (lanes whose arena
        // is less than half full skip it on the device)
        h->steps_since_gc += nb;
        if (h->gc_period > 0 && h->steps_since_gc >= h->gc_period) {
            h->steps_since_gc = 0;
            const dim3 grid_gc(std::max(2, std::min(32, 1184 / L)), L);
            k_gc_decide<<<(L + 127) / 128, 128, 0, h->stream>>>(d);
            k_gc_mark<<<grid_gc, JG_THREADS, 0, h->stream>>>(d);
            k_gc_sweep<<<grid_gc, JG_THREADS, 0, h->stream>>>(d);
            h->launches += 3;
            CK(cudaGetLastError());
        }
    }
    if (last) {
        h->prof_begin(JGPU_K_BOUNDARY);
        k_boundary<<<L, 32, 0, h->stream>>>(d);          // close the last step, run pending finishes
        h->prof_end();
        ++h->launches;
        CK(cudaGetLastError());
    }
    // (the host vectors may be reused at once: pageable copies are staged before cudaMemcpyAsync returns)
    return JGPU_OK;
}

// Runs a prebuilt schedule of `n_steps` rows (plus the trailing close-only row n_steps): the streaming interface.
int run_schedule(jgpu_handle* h, const std::vector<int4>& sched, int n_steps, const float* d_x)
{
    const int L = h->d.n_lanes, CH = h->sched_chunk;
    std::vector<int4> chunk;
    for (int s0 = 0; s0 <= n_steps; s0 += CH) {
        const int ns = std::min(CH, n_steps - s0);          // frame/seed steps in this chunk
        const bool last = s0 + ns == n_steps;
        chunk.assign(sched.begin() + (size_t)s0 * L, sched.begin() + (size_t)(s0 + ns + (last ? 1 : 0)) * L);
        int rc = submit_chunk(h, chunk, ns, last, d_x);
        if (rc) return rc;
        if (last) break;
    }
    return JGPU_OK;
}

// Where the scheduler gets its utterances from: the caller's list in a fixed order (one rank, or a second pass),
// or a queue shared by the ranks of a node (jgpu_decode_queue: whole-utterance work stealing).
struct UttSource {
    virtual ~UttSource() {}
    virtual int next() = 0;                               // index into the caller's arrays, -1 when exhausted
    virtual bool shared() const { return false; }
};
struct ListSource : UttSource {
    const std::vector<int>& ids;
    size_t pos = 0;
    explicit ListSource(const std::vector<int>& v) : ids(v) {}
    int next() override { return pos < ids.size() ? ids[pos++] : -1; }
};

// The lock-step scheduler.  Lanes are refilled as their utterances end: a free lane takes the next utterance of the
// source — with the list sorted longest first that is LPT assignment; with a shared queue it is work stealing, and
// then the schedule is built in short chunks that stay at most one chunk ahead of the device, so that a claim
// reflects how far THIS GPU really is.  slot_utt[k] = utterance decoded into result slot k.
int run_stream(jgpu_handle* h, UttSource* src, const float* d_feats, const int64_t* row_offset, const int32_t* n_frames,
               std::vector<int>* slot_utt, const std::function<int(int)>& on_claim)
{
    const int L = h->d.n_lanes;
    const int CH = src->shared() ? 4 * h->FB : h->sched_chunk;
    std::vector<int> cur(L, -1), pos(L, 0), slot(L, -1);
    std::vector<char> fin(L, 0);
    bool exhausted = false;
    std::vector<int4> chunk;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    if (src->shared())
        for (int i = 0; i < 2; ++i) CK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    int rc = JGPU_OK;
    for (int k = 0; !rc; ++k) {
        // (shared queue) chunk k-2 must be done before anything is claimed for chunk k: chunk k-1 keeps the GPU busy
        if (src->shared() && k >= 2 && cudaEventSynchronize(ev[k & 1]) != cudaSuccess) { rc = fail(JGPU_E_CUDA, "queue throttle: %s", cudaGetErrorString(cudaGetLastError())); break; }
        chunk.assign((size_t)(CH + 1) * L, make_int4(-1, 0, JG_MODE_IDLE, -1));
        int ns = 0;
        bool done = false;
        for (; ns < CH && !rc; ++ns) {
            bool any = false;
            for (int l = 0; l < L && !rc; ++l) {
                int4& e = chunk[(size_t)ns * L + l];
                if (fin[l]) { e.z |= JG_FLAG_FINISH; fin[l] = 0; }   // row after the last frame of the lane's previous utterance
                if (cur[l] < 0 && !exhausted) {
                    const int u = src->next();
                    if (u < 0) exhausted = true;
                    else {
                        cur[l] = u; pos[l] = -1;
                        slot[l] = (int)slot_utt->size();
                        slot_utt->push_back(u);
                        if (on_claim) rc = on_claim(u);
                    }
                }
                if (cur[l] < 0) continue;
                const int u = cur[l];
                any = true;
                if (pos[l] < 0) { e.z |= JG_MODE_SEED; e.w = slot[l]; pos[l] = 0; }
                else { e.x = (int)(row_offset[u] + pos[l]); e.z |= JG_MODE_FRAME; e.w = slot[l]; ++pos[l]; }
                if (pos[l] >= n_frames[u]) { fin[l] = 1; cur[l] = -1; }
            }
            if (!any) { done = true; break; }              // this row is the close-only row: pending finishes only
        }
        if (rc) break;
        rc = submit_chunk(h, chunk, ns, done, d_feats);
        if (!rc && src->shared()) cudaEventRecord(ev[k & 1], h->stream);
        if (done) break;
    }
    for (int i = 0; i < 2; ++i) if (ev[i]) cudaEventDestroy(ev[i]);
    return rc;
}

void fill_result(const ResHdr& hdr, const JgpuWord* pool, JgpuResult* out)
{
    out->status = hdr.status; out->n_frames = hdr.n_frames;
    out->score = hdr.score; out->ac = hdr.ac; out->lm = hdr.lm;
    if (hdr.status > 0 && out->words && out->max_words > 0) {
        // words[] holds min(status, max_words) records; when the caller's buffer is the smaller one it gets the NEWEST
        // words, so that the last record is always the one carrying the final-weight-inclusive totals
        const int n = std::min(hdr.status, out->max_words);
        memcpy(out->words, pool + hdr.word_off + (hdr.status - n), (size_t)n * sizeof(JgpuWord));
    }
}

int fetch_result(jgpu_handle* h, int slot, JgpuResult* out)
{
    const Dev& d = h->d;
    ResHdr hdr;
    CK(cudaMemcpy(&hdr, d.res_hdr + slot, sizeof(hdr), cudaMemcpyDeviceToHost));
    std::vector<JgpuWord> words;
    if (hdr.status > 0) {
        words.resize((size_t)hdr.word_off + hdr.status);
        CK(cudaMemcpy(words.data() + hdr.word_off, d.res_words + hdr.word_off, (size_t)hdr.status * sizeof(JgpuWord), cudaMemcpyDeviceToHost));
    }
    fill_result(hdr, words.data(), out);
    return JGPU_OK;
}

// One pass with the current view: the scheduler takes utterances from `src` (result slot k <- utterance
// slot_utt[k]), results go into out[utterance].  err[k] = device error bits of slot k, need_words[k] = the words
// its best path has.
int decode_pass(jgpu_handle* h, UttSource* src, const float* d_feats, const int64_t* row_offset, const int32_t* n_frames,
                JgpuResult* out, std::vector<int>* slot_utt, std::vector<int>* err, std::vector<int>* need_words,
                const std::function<int(int)>& on_claim)
{
    Dev& d = h->d;
    const int L = d.n_lanes;
    slot_utt->clear();
    k_reset_batch_stats<<<(L + 127) / 128, 128, 0, h->stream>>>(d);
    ++h->launches;
    CK(cudaMemsetAsync(d.res_used, 0, sizeof(int), h->stream));
#ifdef JG_TRACE
    if (h->d_trace) {
        static const unsigned zero = 0;
        static const int off = 0;
        h->abs_step = 0;
        cudaMemcpyToSymbolAsync(g_trace_n, &zero, sizeof(unsigned), 0, cudaMemcpyHostToDevice, h->stream);
        cudaMemcpyToSymbolAsync(g_trace_on, &off, sizeof(int), 0, cudaMemcpyHostToDevice, h->stream);
    }
#endif
    int rc = run_stream(h, src, d_feats, row_offset, n_frames, slot_utt, on_claim);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    const int n_utts = (int)slot_utt->size();
    const std::vector<int>& ids = *slot_utt;
#ifdef JG_TRACE
    if (h->d_trace && !h->prof_on) {                   // (the per-kernel event timing widens the launch gaps)
        unsigned n = 0;
        cudaMemcpyFromSymbol(&n, g_trace_n, sizeof(unsigned));
        n = std::min(n, h->trace_cap);
        std::vector<TraceRec> recs(n);
        if (n) cudaMemcpy(recs.data(), h->d_trace, (size_t)n * sizeof(TraceRec), cudaMemcpyDeviceToHost);
        if (FILE* f = fopen(h->trace_path.c_str(), "wb")) {
            fwrite(recs.data(), sizeof(TraceRec), n, f);
            fclose(f);
        }
    }
#endif
    {
        std::vector<LaneCtl> ctl(L);
        CK(cudaMemcpy(ctl.data(), d.ctl, (size_t)L * sizeof(LaneCtl), cudaMemcpyDeviceToHost));
        long long b[10] = {0};
        for (int l = 0; l < L; ++l)
            for (int i = 0; i < 10; ++i) b[i] += ctl[l].b_stats[i];
        JgpuStats& st = h->batch_stats;                      // (summed over the passes of a batch call)
        st.n_frames += b[0]; st.total_active_models += b[1]; st.total_active_emit_hyps += b[2];
        st.total_active_end_hyps += b[3]; st.total_proc_emit_hyps += b[4]; st.total_proc_end_hyps += b[5];
        st.total_gmm_evals += b[6]; st.total_arcs_expanded += b[7]; st.total_entry_writes += b[8]; st.total_paths += b[9];
    }
    // results
    std::vector<ResHdr> hdr(n_utts);
    if (n_utts) CK(cudaMemcpy(hdr.data(), d.res_hdr, (size_t)n_utts * sizeof(ResHdr), cudaMemcpyDeviceToHost));
    int used = 0;
    CK(cudaMemcpy(&used, d.res_used, sizeof(int), cudaMemcpyDeviceToHost));
    used = std::max(0, std::min(used, d.res_words_cap));
    std::vector<JgpuWord> words((size_t)used);
    bool want_words = false;
    for (int k = 0; k < n_utts; ++k) want_words |= (hdr[k].status > 0 && out[ids[k]].words && out[ids[k]].max_words > 0);
    if (want_words && used) CK(cudaMemcpy(words.data(), d.res_words, (size_t)used * sizeof(JgpuWord), cudaMemcpyDeviceToHost));
    err->assign(n_utts, 0);
    need_words->assign(n_utts, 0);
    for (int k = 0; k < n_utts; ++k) {
        fill_result(hdr[k], words.data(), &out[ids[k]]);
        (*err)[k] = hdr[k].error;
        (*need_words)[k] = hdr[k].n_words;
    }
    return JGPU_OK;
}

// A batch call: one pass with the handle's view; utterances whose lane overflowed an arena (wide beams on a large
// network) are then decoded again, from their first frame, with the pools viewed as fewer lanes with larger arenas,
// until nothing fails or nothing can grow: the reference has no capacity at all (it keeps allocating,
// src/WFSTDecoderLite.cpp:751-805), so a capacity failure must never be what the caller sees.
// `first` = the source of the first pass (nullptr: all n_utts utterances, longest first); `claimed`, when given,
// receives the utterances this call decoded (a shared queue hands every rank a different subset).
int decode_common(jgpu_handle* h, const float* d_feats, const int64_t* row_offset, const int32_t* n_frames,
                  int32_t n_utts, JgpuResult* out, UttSource* first = nullptr, std::vector<int>* claimed = nullptr,
                  const std::function<int(int)>& on_claim = nullptr)
{
    for (auto& l : h->lanes)
        if (l.begun) return fail(JGPU_E_STATE, "streaming utterance in flight: finish it before a batch call");
    for (int i = 0; i < n_utts; ++i)
        if (n_frames[i] < 0) return fail(JGPU_E_ARG, "utterance %d: negative frame count", i);
    size_t pool0 = std::max<size_t>((size_t)n_utts * 64, 1u << 16);
    if (const char* e = getenv("JUICER_B200_WORD_POOL")) pool0 = (size_t)std::max(1, atoi(e));   // (tests: force the pool to grow)
    int rc = ensure_results(h, (size_t)n_utts, pool0);
    if (rc) return rc;
    memset(&h->batch_stats, 0, sizeof(h->batch_stats));
    // longest first: a free lane takes the longest utterance left (LPT)
    std::vector<int> ids(n_utts), slot_utt, err, need_words;
    for (int i = 0; i < n_utts; ++i) ids[i] = i;
    std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) { return n_frames[a] > n_frames[b]; });
    if ((rc = set_view(h, h->sticky ? h->sticky_view : h->base))) return rc;
    for (int pass = 0; pass < 4; ++pass) {
        ListSource list(ids);
        UttSource* src = (pass == 0 && first) ? first : &list;
        if ((rc = decode_pass(h, src, d_feats, row_offset, n_frames, out, &slot_utt, &err, &need_words, on_claim))) break;
        if (pass == 0 && claimed) *claimed = slot_utt;
        std::vector<int> failed;
        int errors = 0;
        size_t words = 0;
        for (size_t k = 0; k < slot_utt.size(); ++k)
            if (err[k] & JG_ERR_RETRYABLE) {
                failed.push_back(slot_utt[k]);
                errors |= err[k];
                words += (size_t)std::max(need_words[k], 64);
            }
        if (failed.empty()) break;
        jgpu_handle::View v;
        const bool grew = grow_view(h, errors, (int)failed.size(), &v);
        const bool more_words = (errors & JG_ERR_WORDS) != 0;
        if (!grew && !more_words) break;                      // nothing left to enlarge: the statuses stand
        if (more_words && (rc = ensure_results(h, (size_t)n_utts, words + words / 2))) break;
        if (grew && (rc = set_view(h, v))) break;
        // a batch that mostly overflowed starts the next call with the larger arenas straight away
        if (grew && pass == 0 && failed.size() * 2 > slot_utt.size()) { h->sticky = true; h->sticky_view = v; }
        ++h->retry_passes;
        std::stable_sort(failed.begin(), failed.end(), [&](int a, int b) { return n_frames[a] > n_frames[b]; });
        ids.swap(failed);
    }
    const int rc2 = set_view(h, h->base);                     // the streaming interface always sees the base view
    return rc ? rc : rc2;
}

// device buffer for the packed features of a host-buffer batch call
int ensure_feats(jgpu_handle* h, size_t rows)
{
    if (rows > h->feats_cap) {
        CK(cudaStreamSynchronize(h->stream));
        if (h->d_feats) cudaFree(h->d_feats);
        h->d_feats = nullptr;
        h->feats_cap = 0;
        CK(cudaMalloc(&h->d_feats, std::max<size_t>(rows, 1) * h->dim * sizeof(float)));
        h->feats_cap = rows;
    }
    return JGPU_OK;
}

int lane_stats(jgpu_handle* h, int lane, JgpuStats* out)
{
    LaneCtl c;
    CK(cudaMemcpy(&c, h->d.ctl + lane, sizeof(c), cudaMemcpyDeviceToHost));
    out->n_frames += c.s_frames;
    out->total_active_models += c.s_active_models;
    out->total_active_emit_hyps += c.s_active_emit;
    out->total_active_end_hyps += c.s_active_end;
    out->total_proc_emit_hyps += c.s_proc_emit;
    out->total_proc_end_hyps += c.s_proc_end;
    out->total_gmm_evals += c.s_gmm;
    out->total_arcs_expanded += c.s_arcs;
    out->total_entry_writes += c.s_entry;
    out->total_paths += c.s_paths;
    return JGPU_OK;
}

} // namespace

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

const char* jgpu_last_error(void) { return jgpu_err_buf().c_str(); }
const char* jgpu_version(void) { return "juicer_b200 0.2 (sm_100a)"; }

int jgpu_create(const JgpuNet* net, const JgpuHmm* hmm, const JgpuGmm* gmm, const JgpuCfg* cfg, jgpu_handle** out)
{
    if (!out) return fail(JGPU_E_ARG, "null out");
    *out = nullptr;
    int rc = validate(net, hmm, gmm, cfg);
    if (rc) return rc;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(JGPU_E_CUDA, "no CUDA device (%s): juicer_b200 has no CPU path", cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(JGPU_E_ARG, "device %d out of range (%d devices)", cfg->device, ndev);
    CK(cudaSetDevice(cfg->device));
    jgpu_handle* h = new jgpu_handle;
    h->cfg = *cfg;
    h->device = cfg->device;
    cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, h->device);
    if (const char* g = getenv("JUICER_B200_GRAPHS")) h->use_graphs = atoi(g) != 0;
    e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete h; return fail(JGPU_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    rc = build_tables(h, net, hmm, gmm);
    if (!rc) rc = build_state(h);
#ifdef JG_TRACE
    if (!rc && getenv("JUICER_B200_TRACE")) {
        h->trace_path = getenv("JUICER_B200_TRACE");
        h->trace_cap = 1u << 21;
        h->trace_s0 = 400; h->trace_s1 = 408;
        if (const char* w = getenv("JUICER_B200_TRACE_STEPS")) sscanf(w, "%d:%d", &h->trace_s0, &h->trace_s1);
        rc = h->alloc(&h->d_trace, (size_t)h->trace_cap);
        if (!rc) {
            cudaMemcpyToSymbol(g_trace, &h->d_trace, sizeof(TraceRec*));
            cudaMemcpyToSymbol(g_trace_cap, &h->trace_cap, sizeof(unsigned));
        }
    }
#endif
    if (!rc) {
        e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail(JGPU_E_CUDA, "create sync: %s", cudaGetErrorString(e));
    }
    if (rc) { jgpu_destroy(h); return rc; }
    *out = h;
    return JGPU_OK;
}

int jgpu_destroy(jgpu_handle* h)
{
    if (!h) return JGPU_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    drop_graphs(h);
    for (void* p : h->allocs) cudaFree(p);
    if (h->d_feats) cudaFree(h->d_feats);
    for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
    if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return JGPU_OK;
}

int jgpu_gmm_scores(jgpu_handle* h, const float* x, int32_t n_rows, float* out)
{
    if (!h || !x || !out || n_rows < 0) return fail(JGPU_E_ARG, "bad argument");
    CK(cudaSetDevice(h->device));
    const int CH = h->gmm_chunk, G = h->d.n_gmms, D = h->dim;
    float* d_x = nullptr;
    int* d_rows = nullptr;
    CK(cudaMalloc(&d_x, (size_t)std::max(CH, 1) * D * sizeof(float)));
    CK(cudaMalloc(&d_rows, (size_t)CH * sizeof(int)));
    std::vector<int> rows(CH);
    for (int i = 0; i < CH; ++i) rows[i] = i;
    cudaMemcpyAsync(d_rows, rows.data(), CH * sizeof(int), cudaMemcpyHostToDevice, h->stream);
    int rc = JGPU_OK;
    for (int r0 = 0; r0 < n_rows && !rc; r0 += CH) {
        const int n = std::min(CH, n_rows - r0);
        cudaMemcpyAsync(d_x, x + (size_t)r0 * D, (size_t)n * D * sizeof(float), cudaMemcpyHostToDevice, h->stream);
        rc = launch_gmm(h, d_x, d_rows, n, h->d_gmm_out, 0);
        if (rc) break;
        cudaMemcpyAsync(out + (size_t)r0 * G, h->d_gmm_out, (size_t)n * G * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
        if (cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fail(JGPU_E_CUDA, "gmm_scores: %s", cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(d_x);
    cudaFree(d_rows);
    return rc;
}

int jgpu_utt_begin(jgpu_handle* h, int32_t lane)
{
    if (!h || lane < 0 || lane >= h->d.n_lanes) return fail(JGPU_E_ARG, "bad lane");
    h->lanes[lane] = LaneHost();
    h->lanes[lane].begun = true;
    return JGPU_OK;
}

int jgpu_push_frames(jgpu_handle* h, int32_t lane, const float* x, int32_t n_frames)
{
    if (!h || lane < 0 || lane >= h->d.n_lanes || n_frames < 0 || (n_frames > 0 && !x)) return fail(JGPU_E_ARG, "bad argument");
    LaneHost& lh = h->lanes[lane];
    if (!lh.begun) return fail(JGPU_E_STATE, "push_frames on lane %d without utt_begin", lane);
    CK(cudaSetDevice(h->device));
    const int L = h->d.n_lanes, D = h->dim;
    int f0 = 0;
    do {
        // later copies into the staging buffer are stream-ordered behind the kernels reading it
        const int n = std::min(h->stream_chunk, n_frames - f0);
        const int seed = lh.seeded ? 0 : 1;
        const int steps = seed + n;
        if (n > 0)
            CK(cudaMemcpyAsync(h->d_stream_feats + (size_t)lane * h->stream_chunk * D, x + (size_t)f0 * D,
                               (size_t)n * D * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        if (steps > 0) {
            std::vector<int4> sched((size_t)(steps + 1) * L, make_int4(-1, 0, JG_MODE_IDLE, -1));
            if (seed) sched[lane] = make_int4(-1, 0, JG_MODE_SEED, lane);
            for (int t = 0; t < n; ++t)
                sched[(size_t)(seed + t) * L + lane] = make_int4(lane * h->stream_chunk + t, 0, JG_MODE_FRAME, lane);
            int rc = run_schedule(h, sched, steps, h->d_stream_feats);
            if (rc) return rc;
        }
        lh.seeded = true;
        lh.frames += n;
        f0 += n;
    } while (f0 < n_frames);
    return JGPU_OK;
}

int jgpu_utt_end(jgpu_handle* h, int32_t lane, JgpuResult* out)
{
    if (!h || lane < 0 || lane >= h->d.n_lanes || !out) return fail(JGPU_E_ARG, "bad argument");
    LaneHost& lh = h->lanes[lane];
    if (!lh.begun) return fail(JGPU_E_STATE, "utt_end on lane %d without utt_begin", lane);
    CK(cudaSetDevice(h->device));
    int rc;
    if (!lh.seeded && (rc = jgpu_push_frames(h, lane, nullptr, 0))) return rc;
    const int L = h->d.n_lanes;
    std::vector<int4> sched((size_t)L, make_int4(-1, 0, JG_MODE_IDLE, -1));
    sched[lane].z |= JG_FLAG_FINISH;
    CK(cudaMemsetAsync(h->d.res_used, 0, sizeof(int), h->stream));   // one result at a time sits in the word pool
    if ((rc = run_schedule(h, sched, 0, h->d_stream_feats))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    lh.begun = false;
    return fetch_result(h, lane, out);
}

int jgpu_partial_result(jgpu_handle* h, int32_t lane, JgpuResult* out)
{
    if (!h || lane < 0 || lane >= h->d.n_lanes || !out) return fail(JGPU_E_ARG, "bad argument");
    LaneHost& lh = h->lanes[lane];
    if (!lh.begun) return fail(JGPU_E_STATE, "partial_result on lane %d without utt_begin", lane);
    CK(cudaSetDevice(h->device));
    if (!lh.seeded) {                                          // nothing decoded yet
        out->status = -1; out->n_frames = 0; out->score = out->ac = out->lm = JGPU_LOG_ZERO;
        return JGPU_OK;
    }
    int rc;
    if (!h->d_partial && (rc = h->alloc(&h->d_partial, 4))) return rc;
    const Dev& d = h->d;
    CK(cudaMemsetAsync(h->d_partial, 0, 4 * sizeof(int), h->stream));
    CK(cudaMemsetAsync(d.res_used, 0, sizeof(int), h->stream));
    k_partial_reset<<<128, JG_THREADS, 0, h->stream>>>(d, lane, h->d_partial);
    k_partial_count<<<128, JG_THREADS, 0, h->stream>>>(d, lane, h->d_partial);
    k_partial_flag<<<128, JG_THREADS, 0, h->stream>>>(d, lane, h->d_partial);
    k_partial_head<<<128, JG_THREADS, 0, h->stream>>>(d, lane, h->d_partial);
    k_partial_emit<<<1, 1, 0, h->stream>>>(d, lane, lane, h->d_partial);
    h->launches += 5;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    return fetch_result(h, lane, out);
}

int jgpu_decode_batch_device(jgpu_handle* h, const float* d_feats, const int64_t* row_offset, const int32_t* n_frames,
                             int32_t n_utts, JgpuResult* out)
{
    if (!h || n_utts < 0 || (n_utts > 0 && (!d_feats || !row_offset || !n_frames || !out))) return fail(JGPU_E_ARG, "bad argument");
    CK(cudaSetDevice(h->device));
    return decode_common(h, d_feats, row_offset, n_frames, n_utts, out);
}

int jgpu_decode_batch(jgpu_handle* h, const float* const* feats, const int32_t* n_frames, int32_t n_utts, JgpuResult* out)
{
    if (!h || n_utts < 0 || (n_utts > 0 && (!feats || !n_frames || !out))) return fail(JGPU_E_ARG, "bad argument");
    CK(cudaSetDevice(h->device));
    std::vector<int64_t> off(n_utts + 1, 0);
    for (int u = 0; u < n_utts; ++u) {
        if (n_frames[u] < 0) return fail(JGPU_E_ARG, "utterance %d: negative frame count", u);
        off[u + 1] = off[u] + n_frames[u];
    }
    int rc0 = ensure_feats(h, (size_t)off[n_utts]);
    if (rc0) return rc0;
    for (int u = 0; u < n_utts; ++u)
        if (n_frames[u] > 0)
            CK(cudaMemcpyAsync(h->d_feats + (size_t)off[u] * h->dim, feats[u], (size_t)n_frames[u] * h->dim * sizeof(float),
                               cudaMemcpyHostToDevice, h->stream));
    return decode_common(h, h->d_feats, off.data(), n_frames, n_utts, out);
}

namespace {
struct QueueSource : UttSource {
    jgpu_queue* q;
    std::vector<int> order;
    QueueSource(jgpu_queue* q_, const int32_t* ord, const int32_t* n_frames, int n) : q(q_), order(n)
    {
        for (int i = 0; i < n; ++i) order[i] = ord ? ord[i] : i;
        if (!ord) std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return n_frames[a] > n_frames[b]; });
    }
    int next() override
    {
        const int64_t i = jgpu_queue_claim(q, 1);
        return (i >= 0 && i < (int64_t)order.size()) ? order[(size_t)i] : -1;
    }
    bool shared() const override { return true; }
};

int decode_queue_common(jgpu_handle* h, jgpu_queue* q, const float* d_feats, const int64_t* row_offset, const int32_t* n_frames,
                        int32_t n_utts, const int32_t* order, JgpuResult* out, int32_t* claimed, int32_t* n_claimed,
                        double* busy_ms, const std::function<int(int)>& on_claim)
{
    if (order)
        for (int i = 0; i < n_utts; ++i)
            if (order[i] < 0 || order[i] >= n_utts) return fail(JGPU_E_ARG, "order[%d] = %d out of range", i, order[i]);
    QueueSource src(q, order, n_frames, n_utts);
    std::vector<int> mine;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (busy_ms) {
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0, h->stream));
    }
    int rc = decode_common(h, d_feats, row_offset, n_frames, n_utts, out, &src, &mine, on_claim);
    if (busy_ms) {
        cudaEventRecord(e1, h->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        *busy_ms = ms;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    if (n_claimed) *n_claimed = (int32_t)mine.size();
    if (claimed)
        for (size_t k = 0; k < mine.size(); ++k) claimed[k] = mine[k];
    return rc;
}
} // namespace

int jgpu_decode_queue_device(jgpu_handle* h, jgpu_queue* q, const float* d_feats, const int64_t* row_offset,
                             const int32_t* n_frames, int32_t n_utts, const int32_t* order, JgpuResult* out,
                             int32_t* claimed, int32_t* n_claimed, double* busy_ms)
{
    if (!h || !q || n_utts < 0 || (n_utts > 0 && (!d_feats || !row_offset || !n_frames || !out))) return fail(JGPU_E_ARG, "bad argument");
    CK(cudaSetDevice(h->device));
    return decode_queue_common(h, q, d_feats, row_offset, n_frames, n_utts, order, out, claimed, n_claimed, busy_ms, nullptr);
}

int jgpu_decode_queue(jgpu_handle* h, jgpu_queue* q, const float* const* feats, const int32_t* n_frames, int32_t n_utts,
                      const int32_t* order, JgpuResult* out, int32_t* claimed, int32_t* n_claimed, double* busy_ms)
{
    if (!h || !q || n_utts < 0 || (n_utts > 0 && (!feats || !n_frames || !out))) return fail(JGPU_E_ARG, "bad argument");
    CK(cudaSetDevice(h->device));
    std::vector<int64_t> off(n_utts + 1, 0);
    for (int u = 0; u < n_utts; ++u) {
        if (n_frames[u] < 0) return fail(JGPU_E_ARG, "utterance %d: negative frame count", u);
        off[u + 1] = off[u] + n_frames[u];
    }
    int rc = ensure_feats(h, (size_t)off[n_utts]);
    if (rc) return rc;
    // the features of an utterance cross the bus when it is claimed, stream-ordered ahead of the steps that read them
    auto on_claim = [&](int u) -> int {
        if (n_frames[u] > 0)
            CK(cudaMemcpyAsync(h->d_feats + (size_t)off[u] * h->dim, feats[u], (size_t)n_frames[u] * h->dim * sizeof(float),
                               cudaMemcpyHostToDevice, h->stream));
        return JGPU_OK;
    };
    return decode_queue_common(h, q, h->d_feats, off.data(), n_frames, n_utts, order, out, claimed, n_claimed, busy_ms, on_claim);
}

int jgpu_stats(jgpu_handle* h, int32_t lane, JgpuStats* out)
{
    if (!h || !out || lane < -1 || lane >= h->d.n_lanes) return fail(JGPU_E_ARG, "bad argument");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    memset(out, 0, sizeof(*out));
    if (lane >= 0) return lane_stats(h, lane, out);
    *out = h->batch_stats;
    return JGPU_OK;
}

int jgpu_frame_stats(jgpu_handle* h, int32_t lane, int32_t* cnt, float* best, int32_t max_frames)
{
    if (!h || lane < 0 || lane >= h->d.n_lanes) return fail(JGPU_E_ARG, "bad argument");
    if (!h->d.frame_stats) return fail(JGPU_E_STATE, "handle created without cfg.frame_stats");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    LaneCtl c;
    CK(cudaMemcpy(&c, h->d.ctl + lane, sizeof(c), cudaMemcpyDeviceToHost));
    const int n = std::min(std::min(c.frame, h->d.max_frames), max_frames);
    if (n > 0 && cnt) CK(cudaMemcpy(cnt, h->d.fstat_cnt + (size_t)lane * h->d.max_frames * 4, (size_t)n * 4 * sizeof(int), cudaMemcpyDeviceToHost));
    if (n > 0 && best) CK(cudaMemcpy(best, h->d.fstat_best + (size_t)lane * h->d.max_frames, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    return n;
}

// Measurement utility (bench.py: roofline.secondary): the FP32 issue peak for code that may not use FMA — what bounds
// the scorer, whose exact operand order forbids contraction.  16 independent multiply chains per
// thread, 8 CTAs of 256 threads per SM; returns 10^12 scalar FP32 operations per second, best of 5 launches.
int jgpu_ubench_fp32(jgpu_handle* h, double* tera_ops)
{
    if (!h || !tera_ops) return fail(JGPU_E_ARG, "bad argument");
    CK(cudaSetDevice(h->device));
    float* out = nullptr;
    const int grid = h->n_sm * 8, iters = 1 << 14;
    CK(cudaMalloc(&out, (size_t)grid * 256 * sizeof(float)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0, h->stream);
        k_ubench_fp32<<<grid, 256, 0, h->stream>>>(out, iters, 0.999f);
        cudaEventRecord(e1, h->stream);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)grid * 256 * (double)iters * 32.0;
        if (rep > 0 && ms > 0.f) best = std::max(best, ops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    CK(cudaGetLastError());
    *tera_ops = best;
    return JGPU_OK;
}

int64_t jgpu_launch_count(jgpu_handle* h) { return h ? h->launches : 0; }
int64_t jgpu_retry_count(jgpu_handle* h) { return h ? h->retry_passes : 0; }

int jgpu_set_stream(jgpu_handle* h, void* cuda_stream)
{
    if (!h) return fail(JGPU_E_ARG, "null handle");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)cuda_stream;
    h->own_stream = false;
    return JGPU_OK;
}

int jgpu_profile(jgpu_handle* h, int32_t enable)
{
    if (!h) return fail(JGPU_E_ARG, "null handle");
    CK(cudaSetDevice(h->device));
    h->prof_collect();
    h->prof_on = enable != 0;
    if (enable) {
        for (int k = 0; k < JGPU_N_KERNELS; ++k) { h->prof_ms[k] = 0; h->prof_cnt[k] = 0; }
    }
    return JGPU_OK;
}

int jgpu_profile_read(jgpu_handle* h, double* ms, int64_t* count)
{
    if (!h || !ms || !count) return fail(JGPU_E_ARG, "bad argument");
    CK(cudaSetDevice(h->device));
    h->prof_collect();
    for (int k = 0; k < JGPU_N_KERNELS; ++k) { ms[k] = h->prof_ms[k]; count[k] = h->prof_cnt[k]; }
    return JGPU_N_KERNELS;
}

const char* jgpu_kernel_name(int32_t kind)
{
    // k_expand / _r1 / _r2 = k_walk<0> by round (0, 1, >= 2), k_commit = k_walk<1>, k_filter only with an end / word beam
    static const char* names[JGPU_N_KERNELS] = {"k_gmm_scores", "k_boundary", "k_internal", "k_filter", "k_expand",
                                                "k_commit_huge", "k_commit", "k_expand_r1", "k_expand_r2"};
    return (kind >= 0 && kind < JGPU_N_KERNELS) ? names[kind] : "";
}

} // extern "C"
