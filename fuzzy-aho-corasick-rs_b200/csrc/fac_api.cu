// fac_api.cu -- the C ABI of include/fac.h over the sm_100a kernels.
//
// Host orchestration only: device memory, one stream per in-flight call (workspaces are pooled so
// an engine is re-entrant), kernel launches, and the small synchronous read-backs the control flow
// needs (classification flag, counters).  There is NO CPU implementation of the search path here:
// every entry point fails with FAC_CUDA_ERROR when no device is usable.
#include <cub/cub.cuh>
#include <atomic>

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <memory>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/fac.h"
#include "fac_apply.cuh"
#include "fac_builder.h"
#include "fac_kernels.cuh"
#include "fac_beam.cuh"
#include "fac_fastreduce.cuh"
#include "fac_segment.cuh"
#include "fac_succinct.cuh"
#include "fac_stack.cuh"
#include "fac_beam2.cuh"
#include "fac_bitap.cuh"

#define FAC_TABLE_QUAL static const
#include "unicode_tables.h"

namespace {

thread_local std::string g_err;
thread_local uint64_t g_last_graphemes = 0;

void set_err(const std::string &s) { g_err = s; }

#define CK(call)                                                                                                  \
    do {                                                                                                          \
        cudaError_t e_ = (call);                                                                                  \
        if (e_ != cudaSuccess) {                                                                                  \
            set_err(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
            return e_ == cudaErrorMemoryAllocation ? FAC_OOM : FAC_CUDA_ERROR;                                    \
        }                                                                                                         \
    } while (0)
#define CKS(call)                     \
    do {                              \
        fac_status s_ = (call);       \
        if (s_ != FAC_OK) return s_;  \
    } while (0)

struct DBuf {
    void *p = nullptr;
    size_t cap = 0;
    fac_status ensure(size_t bytes) {
        if (bytes <= cap && p) return FAC_OK;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { set_err(std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e)); p = nullptr; return FAC_OOM; }
        cap = want;
        return FAC_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};

inline uint32_t next_pow2(uint64_t v) { uint64_t p = 1; while (p < v) p <<= 1; return (uint32_t)std::min<uint64_t>(p, 1ull << 31); }
inline uint32_t cdiv(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

struct U8ToU32 {
    __host__ __device__ __forceinline__ uint32_t operator()(const uint8_t &v) const { return v; }
};
struct U8ToU64 {
    __host__ __device__ __forceinline__ unsigned long long operator()(const uint8_t &v) const { return v; }
};

struct SearchStats {
    uint64_t states = 0;
    uint64_t dirty_windows = 0;
    uint64_t pf_slices = 0, pf_hits = 0;
    double device_ms = 0, expand_ms = 0;
    uint32_t launches = 0;
};

struct Workspace {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, evk0 = nullptr, evk1 = nullptr;
    DBuf flat_n, flat_e;
    DBuf hay, mark, gidx, first, gid, off, pfsym, srec, cov, covcnt, covoff, sl_gs, sl_ge, bp_k;
    DBuf queue, nxt, hslot, gtab_rep, gtab_head, gtab_min;
    uint32_t grid = 0, qcap = 0, gtab_size = 0;
    DBuf cands, counters, failed_tiles, failed_bitmap, tiles;
    DBuf best_rep, best_val, cslot, tab_sim, tab_cmin, tab_cmax, tab_first, dirty;
    DBuf keys_a, keys_b;
    DBuf m_a, m_b, idx_a, idx_b, winend, st_a, st_b, flags8, sel, nsel, outm, keep8, windows, misc, cubtmp, used;
    uint64_t *h_counters = nullptr;  // pinned
    uint32_t *h_flags = nullptr;     // pinned
    fac_status init() {
        CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CK(cudaEventCreate(&ev_begin)); CK(cudaEventCreate(&ev_end)); CK(cudaEventCreate(&evk0)); CK(cudaEventCreate(&evk1));
        CK(cudaMallocHost(&h_counters, 16 * sizeof(uint64_t)));
        CK(cudaMallocHost(&h_flags, 16 * sizeof(uint32_t)));
        return FAC_OK;
    }
    void destroy() {
        for (DBuf *b : {&flat_n, &flat_e, &hay, &mark, &gidx, &first, &gid, &off, &pfsym, &srec, &cov, &covcnt, &covoff, &sl_gs, &sl_ge, &bp_k, &queue, &nxt, &hslot, &gtab_rep, &gtab_head, &gtab_min, &cands, &counters,
                        &failed_tiles, &failed_bitmap, &tiles, &best_rep, &best_val, &cslot, &tab_sim, &tab_cmin, &tab_cmax, &tab_first, &dirty, &keys_a, &keys_b, &m_a, &m_b, &idx_a, &idx_b, &winend, &st_a, &st_b,
                        &flags8, &sel, &nsel, &outm, &keep8, &windows, &misc, &cubtmp, &used})
            b->release();
        if (h_counters) cudaFreeHost(h_counters);
        if (h_flags) cudaFreeHost(h_flags);
        if (ev_begin) cudaEventDestroy(ev_begin);
        if (ev_end) cudaEventDestroy(ev_end);
        if (evk0) cudaEventDestroy(evk0);
        if (evk1) cudaEventDestroy(evk1);
        if (stream) cudaStreamDestroy(stream);
    }
};

// ---- pinned result buffers ---------------------------------------------------------------------------
// Match lists can be hundreds of MB (cfg2: ~0.17 matches per haystack byte).  Copying them into pageable
// memory runs at a few GB/s; the result vector therefore allocates page-locked memory from a process-wide
// pool (page-locking is slow, so buffers are recycled) and does not zero-fill on resize.
struct PinnedPool {
    std::mutex mu;
    std::multimap<size_t, void *> free_;
    std::map<void *, size_t> live_;  // pinned pointers handed out -> class size
    size_t pooled_bytes = 0;
    static constexpr size_t kMinPinned = 1u << 20, kMaxPooled = (size_t)24 << 30, kBig = (size_t)256 << 20;
    void *get(size_t bytes) {
        if (bytes < kMinPinned) return malloc(bytes ? bytes : 1);
        size_t cls = kMinPinned;
        if (bytes >= kBig) cls = (bytes + kBig - 1) / kBig * kBig;   // big lists: 256 MiB granularity
        else while (cls < bytes) cls <<= 1;
        {
            std::lock_guard<std::mutex> g(mu);
            auto it = free_.lower_bound(cls);   // any recycled buffer that is large enough (and not wastefully so)
            if (it != free_.end() && it->first <= cls + cls / 2 + kBig) {
                void *p = it->second; const size_t got = it->first;
                free_.erase(it); pooled_bytes -= got; live_[p] = got; return p;
            }
        }
        void *p = nullptr;
        if (cudaHostAlloc(&p, cls, cudaHostAllocPortable) != cudaSuccess) {
            // pageable fallback: the device-to-host copy of the result list drops to a few GB/s -- say so (once per process)
            cudaGetLastError();
            static std::atomic<bool> warned{false};
            if (!warned.exchange(true))
                fprintf(stderr, "libfacgpu: warning: cudaHostAlloc(%zu bytes) failed; match lists are returned in pageable memory (slow device-to-host copies)\n", cls);
            return malloc(bytes);
        }
        std::lock_guard<std::mutex> g(mu);
        live_[p] = cls;
        return p;
    }
    void put(void *p) {
        if (!p) return;
        size_t cls = 0;
        {
            std::lock_guard<std::mutex> g(mu);
            auto it = live_.find(p);
            if (it != live_.end()) {
                cls = it->second; live_.erase(it);
                if (pooled_bytes + cls <= kMaxPooled) { free_.emplace(cls, p); pooled_bytes += cls; return; }
            }
        }
        if (cls) cudaFreeHost(p); else free(p);
    }
};
PinnedPool &pinned_pool() { static PinnedPool *p = new PinnedPool(); return *p; }

template <class T>
struct PinnedAlloc {
    using value_type = T;
    PinnedAlloc() = default;
    template <class U> PinnedAlloc(const PinnedAlloc<U> &) {}
    T *allocate(size_t n) { T *p = (T *)pinned_pool().get(n * sizeof(T)); if (!p) throw std::bad_alloc(); return p; }
    void deallocate(T *p, size_t) { pinned_pool().put(p); }
    template <class U> void construct(U *) {}                                   // resize() without zero-fill
    template <class U, class... A> void construct(U *p, A &&...a) { ::new ((void *)p) U(std::forward<A>(a)...); }
    template <class U> bool operator==(const PinnedAlloc<U> &) const { return true; }
    template <class U> bool operator!=(const PinnedAlloc<U> &) const { return false; }
};
using MatchVec = std::vector<fac_match, PinnedAlloc<fac_match>>;

// Device-resident result lists (FAC_RESULT_ON_DEVICE) are handed to the caller as whole buffers; freed lists go back
// to a small per-process pool so that a multi-GB list is not cudaMalloc'ed / cudaFree'd on every call.
struct DevicePool {
    struct Item { int device; void *p; size_t cap; };
    std::mutex mu;
    std::vector<Item> free_;
    static constexpr size_t kMaxItems = 6;
    bool get(int device, size_t bytes, DBuf &out) {
        std::lock_guard<std::mutex> g(mu);
        size_t best = free_.size();
        for (size_t i = 0; i < free_.size(); i++)
            if (free_[i].device == device && free_[i].cap >= bytes && (best == free_.size() || free_[i].cap < free_[best].cap)) best = i;
        if (best == free_.size()) return false;
        out.p = free_[best].p; out.cap = free_[best].cap;
        free_.erase(free_.begin() + best);
        return true;
    }
    void put(int device, DBuf &b) {
        if (!b.p) return;
        Item victim{device, b.p, b.cap};
        b.p = nullptr; b.cap = 0;
        bool drop = false;
        {
            std::lock_guard<std::mutex> g(mu);
            free_.push_back(victim);
            if (free_.size() > kMaxItems) {  // drop the smallest
                size_t k = 0;
                for (size_t i = 1; i < free_.size(); i++) if (free_[i].cap < free_[k].cap) k = i;
                victim = free_[k]; free_.erase(free_.begin() + k); drop = true;
            }
        }
        if (drop) { int cur = 0; cudaGetDevice(&cur); cudaSetDevice(victim.device); cudaFree(victim.p); cudaSetDevice(cur); }
    }
};
DevicePool &device_pool() { static DevicePool *p = new DevicePool(); return *p; }

}  // namespace

struct fac_engine {
    int device = 0;
    fac::HostAutomaton host;
    AutomatonView dview;
    std::vector<void *> dallocs;
    UnicodeTables dU;
    FacSymbol *d_symbols = nullptr;
    uint32_t sym_mask = 0;
    uint8_t *d_pool = nullptr;
    uint32_t *d_pat_bytes = nullptr;
    uint32_t *d_pat_uid_dense = nullptr;
    uint32_t n_uid = 0;
    int sm_count = 148;
    uint32_t lookahead = 0;
    uint32_t default_tile = 0;  // 0 = adaptive
    uint32_t qcap = 1u << 17;
    uint32_t smem_tab = 1024;
    int ctas_per_sm = 6;
    int use_tma = 1;
    // bitap pre-filter (fac_bitap.cuh)
    const uint64_t *d_bp_mask = nullptr;      // [P][128] ASCII bytes (case folding baked in)
    const uint64_t *d_bp_symmask = nullptr;   // [P][alphabet + 1] symbol ids (non-ASCII haystacks)
    const uint8_t *d_bp_m = nullptr;
    const FacSymbol *d_pf_symbols = nullptr;  // folded grapheme -> pre-filter symbol id (transcode, prefilter.rs:262-281)
    const uint8_t *d_pf_pool = nullptr;
    uint32_t pf_mask = 0;
    // succinct-trie fast kernel (fac_succinct.cuh)
    bool succ_ok = false, succ_generic_ok = false;
    const void *d_s_bm = nullptr;   // u32 [N] (narrow) or u64 [N] with bit 63 = has outputs (wide)
    const uint32_t *d_s_fc = nullptr, *d_s_out_idx = nullptr, *d_s_out2 = nullptr;
    const float *d_s_plen = nullptr, *d_s_plow = nullptr, *d_s_subpen = nullptr;
    const uint8_t *d_s_symof = nullptr;
    const void *d_s_masks = nullptr;   // transposed survivor masks (SuccGMDev): gmT then gm2T
    uint32_t succ_gm2_off = 0, succ_gm3_off = 0, succ_pm3_off = 0, succ_pm2_off = 0, succ_pm4_off = 0;
    const uint32_t *d_s_node_lim = nullptr;
    uint32_t succ_nt = 1024, succ_tile = 4096, succ_stack = 0, succ_min_stack = 64, succ_feed = 1;
    bool radix_unsorted = true;   // FAC_RADIX_UNSORTED=0: Order::Unsorted always through the comparison merge sort
    int smem_optin = 0;
    const uint32_t *d_flat_nrec = nullptr, *d_flat_erec = nullptr, *d_flat_ooff = nullptr, *d_flat_olist = nullptr, *d_flat_gm_row = nullptr;
    const uint64_t *d_flat_gm = nullptr, *d_flat_pm = nullptr, *d_flat_px = nullptr;
    const uint32_t *d_flat_px_row = nullptr;   // static parts of the merged records (fac_flat.h)
    bool beam2_ok = false;          // shared-memory beamed kernel (fac_beam2.cuh)
    uint32_t beam2_warps = 5, beam2_vcap = 1024, beam2_ctas = 8;
    uint32_t max_fan = 0;           // most children one state can push (2 * widest node + mapping transitions + 3)
    bool stack_ok = false;          // general stack-machine kernel (fac_stack.cuh) for fast engines outside the succinct domain
    uint32_t stack_cap = 384, stack_tile = 1024;
    bool fast_ok = false;  // FAST kernel allowed (fast-path edit ceiling, no beam); FAC_FAITHFUL=1 forces the order-faithful kernel
    mutable std::mutex mu;
    mutable std::vector<Workspace *> pool;
    // fac_engine_create_multi: the same automaton on further devices (the stream entry points deal window batches over them)
    std::vector<fac_engine *> replicas;
};

struct fac_matches {
    MatchVec v;
    SearchStats stats;
    // FAC_RESULT_ON_DEVICE: the list stays in device memory (d_out[0..n_dev))
    bool on_device = false;
    int device = 0;
    DBuf d_out;
    size_t n_dev = 0;
    ~fac_matches() { if (d_out.p) device_pool().put(device, d_out); }
};

namespace {

template <class T>
fac_status upload(fac_engine *E, const std::vector<T> &v, const T **out) {
    void *p = nullptr;
    const size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
    CK(cudaMalloc(&p, bytes));
    E->dallocs.push_back(p);
    if (!v.empty()) CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T *)p;
    return FAC_OK;
}
template <class T>
fac_status upload_raw(fac_engine *E, const T *src, size_t n, const T **out) {
    std::vector<T> v(src, src + n);
    return upload(E, v, out);
}

Workspace *acquire_ws(const fac_engine *E, fac_status &st) {
    {
        std::lock_guard<std::mutex> g(E->mu);
        if (!E->pool.empty()) { Workspace *w = E->pool.back(); E->pool.pop_back(); st = FAC_OK; return w; }
    }
    Workspace *w = new Workspace();
    st = w->init();
    if (st != FAC_OK) { w->destroy(); delete w; return nullptr; }
    return w;
}
void release_ws(const fac_engine *E, Workspace *w) {
    std::lock_guard<std::mutex> g(E->mu);
    E->pool.push_back(w);
}

inline bool host_is_ascii(const uint8_t *p, size_t n) {
    size_t i = 0;
    uint64_t acc = 0;
    for (; i + 8 <= n; i += 8) { uint64_t v; memcpy(&v, p + i, 8); acc |= v; }
    for (; i < n; i++) acc |= p[i];
    return (acc & 0x8080808080808080ull) == 0;
}

int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

template <bool ASCII, bool MAPP, bool FAST>
fac_status launch_expand_t(const ExpandParams &P, uint32_t grid, size_t smem, cudaStream_t s) {
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_expand<ASCII, MAPP, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_expand<ASCII, MAPP, FAST><<<grid, FAC_BLOCK, smem, s>>>(P);
    CK(cudaGetLastError());
    return FAC_OK;
}
template <bool FAST>
fac_status launch_expand_f(const ExpandParams &P, uint32_t grid, size_t smem, cudaStream_t s) {
    const bool ascii = P.tv.ascii != 0, mapp = P.A.has_mappings != 0;
    if (ascii && !mapp) return launch_expand_t<true, false, FAST>(P, grid, smem, s);
    if (ascii && mapp) return launch_expand_t<true, true, FAST>(P, grid, smem, s);
    if (!ascii && !mapp) return launch_expand_t<false, false, FAST>(P, grid, smem, s);
    return launch_expand_t<false, true, FAST>(P, grid, smem, s);
}
fac_status launch_expand(const ExpandParams &P, uint32_t grid, size_t smem, cudaStream_t s, bool fast) {
    return fast ? launch_expand_f<true>(P, grid, smem, s) : launch_expand_f<false>(P, grid, smem, s);
}

// ---- succinct fast kernel launch (fac_succinct.cuh) ----
template <int NT, bool LIM, bool W>
fac_status launch_succ_tt(const SuccParams &P, uint32_t grid, size_t smem, cudaStream_t s) {
    CK(cudaFuncSetAttribute(k_expand_succinct<NT, LIM, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_expand_succinct<NT, LIM, W><<<grid, NT, smem, s>>>(P);
    CK(cudaGetLastError());
    return FAC_OK;
}
template <int NT>
fac_status launch_succ_t(const SuccParams &P, bool wide, uint32_t grid, size_t smem, cudaStream_t s) {
    if (wide) return P.K.lim ? launch_succ_tt<NT, true, true>(P, grid, smem, s) : launch_succ_tt<NT, false, true>(P, grid, smem, s);
    return P.K.lim ? launch_succ_tt<NT, true, false>(P, grid, smem, s) : launch_succ_tt<NT, false, false>(P, grid, smem, s);
}
fac_status launch_succinct(const fac_engine *E, Workspace *ws, const TextView &tv, float thr, uint32_t seg_begin, uint32_t seg_end,
                           uint32_t text_end, FacCand *cands, uint32_t cand_cap, cudaStream_t s, const uint4 *d_tiles = nullptr,
                           uint32_t n_explicit = 0) {
    const fac::HostSuccinct &S = E->host.succ;
    const uint32_t N = (uint32_t)S.bm.size();
    CKS(ws->srec.ensure((size_t)N * 16));
    if (S.wide) k_succ_prepare<true><<<cdiv(N, 256), 256, 0, s>>>(E->d_s_bm, E->d_s_fc, E->d_s_plen, E->d_s_plow, E->d_s_out_idx, thr, N, ws->srec.as<uint4>());
    else k_succ_prepare<false><<<cdiv(N, 256), 256, 0, s>>>(E->d_s_bm, E->d_s_fc, E->d_s_plen, E->d_s_plow, E->d_s_out_idx, thr, N, ws->srec.as<uint4>());
    CK(cudaGetLastError());
    SuccParams P;
    memset(&P, 0, sizeof(P));
    P.rec = ws->srec.as<uint4>(); P.out2 = (const uint4 *)E->d_s_out2; P.sub_pen = E->d_s_subpen; P.sym_of = E->d_s_symof; P.text = tv.bytes; P.first = tv.ascii ? nullptr : tv.first;
    P.n_nodes = N;
    P.K.thr = thr;
    P.K.maxpen = S.prune_len[0] - S.prune_low[0] * thr;  // search.rs:487 (host compiled without contraction)
    P.K.pen_ins = E->host.pen_ins; P.K.pen_del = E->host.pen_del; P.K.pen_swap = E->host.pen_swap; P.K.mef = E->host.mef;
    if (S.limits_mode) {   // per-state permissions from the limits tables; mef = bound on a state's total edits
        P.K.mef = (int32_t)S.edit_bound; P.K.lim = E->dview.lim; P.K.node_lim = E->d_s_node_lim; P.K.has_global = E->host.has_global_limits;
    }
    P.exact_only = S.exact_only ? 1 : 0;
    P.K.out_idx = E->d_s_out_idx;
    P.ci = E->host.ci; P.wskip = E->host.wskip; P.first_mask = S.first_mask; P.second_mask = S.second_mask;
    P.seg_begin = seg_begin; P.seg_end = seg_end; P.text_end = text_end;
    P.tile = E->succ_tile; P.n_tiles = cdiv((uint64_t)seg_end - seg_begin, P.tile); P.lookahead = E->lookahead;
    if (d_tiles) { P.tiles = d_tiles; P.n_tiles = n_explicit; }
    // deeper edit budgets push whole sibling sets of non-final states: fewer warps, deeper stacks
    const bool deep = S.limits_mode || E->host.mef > 2;
    const uint32_t nt = deep ? std::min<uint32_t>(E->succ_nt, 512u) : E->succ_nt;
    const uint32_t nw = nt / 32;
    P.stack_cap = E->succ_stack ? E->succ_stack : std::max(!deep ? 128u : 384u, E->succ_min_stack);
    P.feed = E->succ_feed;
    P.text_cap = (P.tile + P.lookahead + 32u + 15u) & ~15u;   // + alignment lead (< 16) + 3 positions of context look-ahead
    P.masks = E->d_s_masks; P.gm_nodes = S.gm_nodes; P.gm2_nodes = S.gm2_nodes; P.gm2_off = E->succ_gm2_off;
    if (!S.wide) {
        P.n3 = S.n3; P.np2 = S.np2; P.n4 = S.n4; P.r3 = S.r3;
        P.gm3_off = E->succ_gm3_off; P.pm3_off = E->succ_pm3_off; P.pm2_off = E->succ_pm2_off; P.pm4_off = E->succ_pm4_off;
    }
    const size_t fixed = (size_t)nw * (P.stack_cap + SUCC_WQ_CAP) * 16 + (S.wide ? 64u : 32u) * SUCC_SP_STRIDE * 4 + (size_t)P.text_cap * (tv.ascii ? 5 : 8) + 256;   // context words + raw tile
    const size_t budget = (size_t)E->smem_optin - 1024;  // static shared + reserve
    if (fixed + 16 * 64 > budget) { set_err("succinct kernel: shared-memory budget too small for the configured stack / tile"); return FAC_UNSUPPORTED; }
    P.n_smem_nodes = (uint32_t)std::min<size_t>(N, (budget - fixed) / 16);
    P.cands = cands; P.cand_cap = cand_cap;
    P.counters = ws->counters.as<unsigned long long>();
    P.dirty = ws->dirty.as<uint32_t>();
    const size_t smem = fixed + (size_t)P.n_smem_nodes * 16;
    const uint32_t grid = std::min<uint32_t>((uint32_t)E->sm_count, P.n_tiles);
    switch (nt) {
        case 1024: return launch_succ_t<1024>(P, S.wide, grid, smem, s);
        case 512: return launch_succ_t<512>(P, S.wide, grid, smem, s);
        default: return launch_succ_t<768>(P, S.wide, grid, smem, s);
    }
}

// ---- general stack-machine kernel launch (fac_stack.cuh) ----
template <bool ASCII, bool MAPP>
fac_status launch_stack_t(const StackParams &SP, uint32_t grid, size_t smem, cudaStream_t s) {
    CK(cudaFuncSetAttribute(k_expand_stack<ASCII, MAPP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_expand_stack<ASCII, MAPP><<<grid, STK_THREADS, smem, s>>>(SP);
    CK(cudaGetLastError());
    return FAC_OK;
}
size_t stack_smem_bytes(const fac_engine *E, bool ascii, uint32_t text_cap) {
    const size_t tb = ascii ? ((text_cap + 31u) & ~15u) : ((text_cap * 4u + 31u) & ~15u);
    const size_t mult = (!ascii && E->host.has_mappings) ? 2 : 1;
    return tb * mult + (size_t)(STK_THREADS / 32) * (E->stack_cap + STK_WQ_CAP) * 16;
}
// per-call merged records (the node ceilings depend on the threshold)
fac_status prepare_flat(const fac_engine *E, Workspace *ws, float thr, FlatView &F, cudaStream_t s) {
    const uint32_t N = E->host.n_nodes(), NE = (uint32_t)E->host.edge_char.size();
    CKS(ws->flat_n.ensure((size_t)std::max<uint32_t>(N, 1) * 16));
    CKS(ws->flat_e.ensure((size_t)std::max<uint32_t>(NE, 1) * 16));
    k_flat_prepare_nodes<<<cdiv(N, 256), 256, 0, s>>>((const uint4 *)E->d_flat_nrec, E->dview.node_prune_len, E->dview.node_prune_low, thr, N, ws->flat_n.as<uint4>());
    if (NE) k_flat_prepare_edges<<<cdiv(NE, 256), 256, 0, s>>>((const uint4 *)E->d_flat_erec, ws->flat_n.as<uint4>(), NE, ws->flat_e.as<uint4>());
    CK(cudaGetLastError());
    F.nrec = ws->flat_n.as<FlatRec>(); F.erec = ws->flat_e.as<FlatRec>(); F.ooff = E->d_flat_ooff; F.olist = E->d_flat_olist;
    F.gm_row = E->d_flat_gm_row; F.gm = (const unsigned long long *)E->d_flat_gm;
    F.pm_root = E->host.flat_pm_g ? (const unsigned long long *)E->d_flat_pm : nullptr; F.pm_g = E->host.flat_pm_g; F.pm_words = E->host.flat_pm_words; F.pm_k = E->host.flat_pm_k;
    F.px_row = E->host.flat_px_row.empty() ? nullptr : E->d_flat_px_row; F.px_bits = (const unsigned long long *)E->d_flat_px;
    return FAC_OK;
}
fac_status launch_beam2(const fac_engine *E, Workspace *ws, const ExpandParams &P, uint32_t bw, uint32_t n_tiles, cudaStream_t s) {
    Beam2Params BP;
    BP.E = P;
    CKS(prepare_flat(E, ws, P.thr, BP.F, s));
    const uint32_t warps = E->beam2_warps, vcap = E->beam2_vcap;
    const size_t smem = (size_t)warps * BM2_SMEM_PER_WARP(vcap);
    const uint32_t per_sm = (uint32_t)std::max<size_t>(1, std::min<size_t>(E->beam2_ctas, ((size_t)227 * 1024) / (smem + 1024)));
    const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)E->sm_count * per_sm, cdiv(n_tiles, warps));
    if (vcap == 512) {
        CK(cudaFuncSetAttribute(k_beam_warp<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_beam_warp<512><<<grid, warps * 32, smem, s>>>(BP, bw);
    } else {
        CK(cudaFuncSetAttribute(k_beam_warp<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_beam_warp<1024><<<grid, warps * 32, smem, s>>>(BP, bw);
    }
    CK(cudaGetLastError());
    return FAC_OK;
}
fac_status launch_stack(const fac_engine *E, Workspace *ws, const ExpandParams &P, uint32_t *dirty, uint32_t n_tiles, cudaStream_t s) {
    StackParams SP;
    SP.E = P; SP.stack_cap = E->stack_cap; SP.dirty = dirty;
    CKS(prepare_flat(E, ws, P.thr, SP.F, s));
    SP.feed_below = 32u;
    const bool ascii = P.tv.ascii != 0, mapp = P.A.has_mappings != 0;
    const size_t smem = stack_smem_bytes(E, ascii, P.smem_text_cap);
    if (smem + 1024 > (size_t)E->smem_optin) { set_err("stack kernel: shared-memory budget too small for the configured stack / tile"); return FAC_UNSUPPORTED; }
    const uint32_t per_sm = (uint32_t)std::max<size_t>(1, std::min<size_t>(8, ((size_t)228 * 1024) / (smem + 1024)));
    const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)E->sm_count * per_sm, n_tiles);
    if (ascii && !mapp) return launch_stack_t<true, false>(SP, grid, smem, s);
    if (ascii && mapp) return launch_stack_t<true, true>(SP, grid, smem, s);
    if (!ascii && !mapp) return launch_stack_t<false, false>(SP, grid, smem, s);
    return launch_stack_t<false, true>(SP, grid, smem, s);
}

size_t expand_smem_bytes(const fac_engine *E, bool ascii, uint32_t text_cap) {
    const size_t tb = ascii ? ((text_cap + 31u) & ~15u) : ((text_cap * 4u + 31u) & ~15u);
    const size_t mult = (!ascii && E->host.has_mappings) ? 2 : 1;
    return tb * mult + (size_t)E->smem_tab * 8;
}

// Make sure the per-CTA expansion scratch exists for `grid` CTAs of `qcap` states.
fac_status ensure_scratch(const fac_engine *E, Workspace *ws, uint32_t grid, uint32_t qcap) {
    const uint32_t gts = next_pow2((uint64_t)qcap * 2);
    const bool fresh_tab = ws->gtab_rep.cap < (size_t)grid * gts * 4 || ws->grid != grid || ws->qcap != qcap;
    CKS(ws->queue.ensure((size_t)grid * qcap * sizeof(FacState)));
    CKS(ws->nxt.ensure((size_t)grid * qcap * 4));
    CKS(ws->hslot.ensure((size_t)grid * qcap * 4));
    CKS(ws->gtab_rep.ensure((size_t)grid * gts * 4));
    CKS(ws->gtab_head.ensure((size_t)grid * gts * 4));
    CKS(ws->gtab_min.ensure((size_t)grid * gts * 4));
    if (fresh_tab) {
        CK(cudaMemsetAsync(ws->gtab_rep.p, 0xFF, (size_t)grid * gts * 4, ws->stream));
        CK(cudaMemsetAsync(ws->gtab_head.p, 0xFF, (size_t)grid * gts * 4, ws->stream));
        k_fill_u32<<<1184, 256, 0, ws->stream>>>(ws->gtab_min.as<uint32_t>(), 0x7F800000u, (uint64_t)grid * gts);  // +inf
        CK(cudaGetLastError());
    }
    ws->grid = grid; ws->qcap = qcap; ws->gtab_size = gts;
    return FAC_OK;
}

// m_a must be able to grow while keeping its contents: ensure() frees, so grow through a copy.
fac_status grow_keep(DBuf &b, size_t keep_bytes, size_t want_bytes, cudaStream_t s) {
    if (want_bytes <= b.cap) return FAC_OK;
    DBuf nb;
    CKS(nb.ensure(want_bytes * 2));
    if (keep_bytes) CK(cudaMemcpyAsync(nb.p, b.p, keep_bytes, cudaMemcpyDeviceToDevice, s));
    CK(cudaStreamSynchronize(s));
    b.release();
    b = nb;
    return FAC_OK;
}

struct ExpandRun {
    // description of what to expand
    TextView tv;
    std::vector<uint4> tiles;     // mode 1 descriptors (empty => mode 0)
    uint32_t seg_begin = 0, seg_end = 0, text_end = 0;
    const FacWindow *d_windows = nullptr;
    float thr = 0.f;
    uint32_t *d_per_window = nullptr;
    bool fast = false;            // FAST kernel (exhausted chains walked in place) + tie detection + faithful redo
    bool count_states = true;
    bool beam = false;            // use the one-window-per-CTA beamed kernel (bw == 0: exact)
    uint32_t bw = 0;
    // called after the expansion has completed (stream synchronised) and before the reduction;
    // returns the exclusive upper bound of the start windows whose candidates count (auto_beam), default: all
    std::function<uint32_t(uint64_t states)> limit_after_expand;
    // pre-filter mode: the merged slices (gs, ge) the explicit tiles were cut from; a start window's
    // text_end is the end of its slice (each slice is searched as its own haystack, prefilter.rs:346-350)
    const std::vector<std::pair<uint32_t, uint32_t>> *slices = nullptr;
    bool slice_tags = false;  // stream batches: slice i is haystack window i (the candidate tag)
};

// Run K3 (+ retry of failed tiles) and the best-per-span reduction; appends WMatch records to
// ws->m_a starting at *n_matches.
fac_status expand_and_reduce(const fac_engine *E, Workspace *ws, const ExpandRun &R, uint32_t tile, uint64_t *n_matches, SearchStats &stats,
                             double *states_per_window_out) {
    cudaStream_t s = ws->stream;
    const bool ascii = R.tv.ascii != 0;
    const bool explicit_tiles = !R.tiles.empty();
    const uint64_t n_windows_total = explicit_tiles ? 0 : (uint64_t)(R.seg_end - R.seg_begin);
    const bool succ_text = ascii || E->host.succ.unicode_text_ok;   // K1 first-char stream works too
    const bool use_succ = R.fast && (E->succ_ok || E->succ_generic_ok) && succ_text && (!explicit_tiles || R.slices) && !R.beam && !R.d_per_window;
    const bool use_stack = R.fast && E->stack_ok && !use_succ && !R.beam && !R.d_per_window;
    if (use_stack && !explicit_tiles) tile = E->stack_tile;   // the stack machines stream windows: large tiles, one staging per tile
    uint32_t n_tiles = explicit_tiles ? (uint32_t)R.tiles.size() : cdiv(n_windows_total, tile);
    if (n_tiles == 0) return FAC_OK;
    uint32_t max_count = tile;
    if (explicit_tiles) { max_count = 1; for (auto &t : R.tiles) max_count = std::max(max_count, t.y); }

    // beamed windows run one per CTA: by default one WARP per window (32 windows in flight per SM, small scratch),
    // windows that overflow that scratch are redone by 256-thread CTAs with the large queue
    // beam widths whose un-popped queue fits the shared-memory ring run on the one-warp-per-window shared-memory kernel
    const bool use_beam2 = R.beam && R.bw > 0 && E->beam2_ok && !R.d_per_window && 2ull * R.bw + 2ull * E->max_fan + 64ull <= BM2_QCAP;
    const uint32_t beam_bs = R.beam ? (uint32_t)(env_int("FAC_BEAM_BLOCK", 32) == 256 ? 256 : 32) : 0;
    const uint32_t ctas = beam_bs == 32 ? 32u : (uint32_t)E->ctas_per_sm;
    const uint32_t run_qcap = beam_bs == 32 ? (uint32_t)std::max(4096, env_int("FAC_BEAM_QCAP", 16384)) : E->qcap;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)E->sm_count * ctas, n_tiles);
    if (!use_succ && !use_stack && !use_beam2) CKS(ensure_scratch(E, ws, (uint32_t)E->sm_count * ctas, run_qcap));   // the stack machines keep their frontier in shared memory
    uint32_t cand_cap = (uint32_t)std::max<size_t>(ws->cands.cap / sizeof(FacCand), 1u << 20);
    // dense-match workloads emit ~0.2-0.4 candidates per start window: size the first attempt so it need not be redone
    if (R.fast && !explicit_tiles) cand_cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(cand_cap, n_windows_total / 2 + (1u << 20)), 0x7FFFFFF0u);
    CKS(ws->cands.ensure((size_t)cand_cap * sizeof(FacCand)));
    cand_cap = (uint32_t)std::min<size_t>(ws->cands.cap / sizeof(FacCand), 0x7FFFFFF0u);
    CKS(ws->counters.ensure(16 * 8));
    CKS(ws->failed_tiles.ensure((size_t)4 * std::max<uint32_t>(n_tiles, 1)));
    CKS(ws->failed_bitmap.ensure((size_t)4 * (n_tiles / 32 + 1)));
    if (explicit_tiles) {
        CKS(ws->tiles.ensure(R.tiles.size() * sizeof(uint4)));
        CK(cudaMemcpyAsync(ws->tiles.p, R.tiles.data(), R.tiles.size() * sizeof(uint4), cudaMemcpyHostToDevice, s));
    }

    ExpandParams P;
    memset(&P, 0, sizeof(P));
    P.A = E->dview; P.tv = R.tv; P.thr = R.thr;
    P.maxpen = E->host.node_prune_len[0] - E->host.node_prune_low[0] * R.thr;  // search.rs:487 (f32, host compiled without contraction)
    P.mode = explicit_tiles ? 1 : 0;
    P.seg_begin = R.seg_begin; P.seg_end = R.seg_end; P.text_end = R.text_end; P.tile = tile; P.n_tiles = n_tiles;
    P.tiles = ws->tiles.as<uint4>();
    P.lookahead = E->lookahead;
    P.queue = ws->queue.as<FacState>(); P.nxt = ws->nxt.as<uint32_t>(); P.hslot = ws->hslot.as<uint32_t>(); P.qcap = ws->qcap;
    P.gtab_rep = ws->gtab_rep.as<uint32_t>(); P.gtab_head = ws->gtab_head.as<uint32_t>(); P.gtab_min = ws->gtab_min.as<float>();
    P.gtab_size = ws->gtab_size; P.smem_tab_size = E->smem_tab;
    P.smem_text_cap = max_count + E->lookahead + 32;
    P.cands = ws->cands.as<FacCand>(); P.cand_cap = cand_cap;
    P.counters = ws->counters.as<unsigned long long>();
    P.failed_tiles = ws->failed_tiles.as<uint32_t>(); P.failed_cap = n_tiles;
    P.failed_bitmap = ws->failed_bitmap.as<uint32_t>();
    P.per_window = R.d_per_window;
    P.use_tma = E->use_tma;
    const size_t smem = expand_smem_bytes(E, ascii, P.smem_text_cap);
    // an exact-only engine is "fast" only through the succinct kernel; the generic FAST kernel needs an edit budget
    const bool fast_run = R.fast && (E->fast_ok || use_succ);
    if (use_succ && explicit_tiles && max_count > E->succ_tile) { set_err("internal: slice tile larger than the succinct tile"); return FAC_INVALID_ARGUMENT; }
    if (use_stack && max_count > FAC_MAX_TILE) { set_err("internal: tile larger than the 12-bit window field"); return FAC_INVALID_ARGUMENT; }
    const uint32_t n_win_seg = R.seg_end - R.seg_begin;
    const uint32_t dirty_words = n_win_seg / 32 + 1;
    if (fast_run) CKS(ws->dirty.ensure((size_t)dirty_words * 4));

    for (int attempt = 0; attempt < 3; attempt++) {
        CK(cudaMemsetAsync(ws->counters.p, 0, 16 * 8, s));
        if (fast_run) CK(cudaMemsetAsync(ws->dirty.p, 0, (size_t)dirty_words * 4, s));
        CK(cudaMemsetAsync(ws->failed_bitmap.p, 0, (size_t)4 * (n_tiles / 32 + 1), s));
        P.pass = 0; P.cand_cap = cand_cap; P.cands = ws->cands.as<FacCand>();
        CK(cudaEventRecord(ws->evk0, s));
        if (use_beam2) { CKS(launch_beam2(E, ws, P, R.bw, n_tiles, s)); stats.launches += 2; }
        else if (R.beam) {
            if (beam_bs == 32) k_expand_beam<32><<<grid, 32, 0, s>>>(P, R.bw);
            else k_expand_beam<256><<<grid, 256, 0, s>>>(P, R.bw);
            CK(cudaGetLastError());
        } else if (use_succ) {
            CKS(launch_succinct(E, ws, R.tv, R.thr, R.seg_begin, R.seg_end, R.text_end, ws->cands.as<FacCand>(), cand_cap, s,
                                explicit_tiles ? ws->tiles.as<uint4>() : nullptr, n_tiles));
            stats.launches++;
        } else if (use_stack) { CKS(launch_stack(E, ws, P, ws->dirty.as<uint32_t>(), n_tiles, s)); stats.launches += 2; }
        else CKS(launch_expand(P, grid, smem, s, fast_run));
        CK(cudaEventRecord(ws->evk1, s));
        stats.launches++;
        CK(cudaMemcpyAsync(ws->h_counters, ws->counters.p, 8 * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ws->evk0, ws->evk1));
        stats.expand_ms += ms;
        uint64_t n_cand = ws->h_counters[1], n_failed = ws->h_counters[3];
        uint64_t states = ws->h_counters[2];
        if (R.beam && n_failed && beam_bs != 32 && !use_beam2) {
            set_err("a beamed start window exceeded the per-window queue capacity (" + std::to_string(ws->qcap / 4) + " states)");
            return FAC_UNSUPPORTED;
        }
        if (n_cand > cand_cap) {  // candidate buffer too small: grow to the exact need and redo the segment
            cand_cap = (uint32_t)std::min<uint64_t>(n_cand + n_cand / 4 + 1024, 0x7FFFFFF0u);
            if (n_cand > cand_cap) { set_err("candidate volume exceeds the 31-bit candidate index space; search a smaller haystack slice"); return FAC_UNSUPPORTED; }
            CKS(ws->cands.ensure((size_t)cand_cap * sizeof(FacCand)));
            continue;
        }
        if (n_failed) {
            // retry the windows of failed tiles one window per tile on a few CTAs with a large queue
            std::vector<uint32_t> failed(n_failed);
            CK(cudaMemcpy(failed.data(), ws->failed_tiles.p, n_failed * 4, cudaMemcpyDeviceToHost));
            std::vector<uint4> rt;
            for (uint32_t t : failed) {
                uint32_t start, count, tend, tag;
                if (explicit_tiles) { start = R.tiles[t].x; count = R.tiles[t].y; tend = R.tiles[t].z; tag = R.tiles[t].w; }
                else { start = R.seg_begin + t * tile; count = std::min(tile, R.seg_end - start); tend = R.text_end; tag = 0; }
                for (uint32_t w = 0; w < count; w++) rt.push_back(make_uint4(start + w, 1, tend, tag));
            }
            const uint32_t big_q = R.beam ? std::max<uint32_t>(E->qcap, 1u << 19) : 1u << 21;
            const uint32_t rgrid = (uint32_t)std::min<size_t>(R.beam ? 128 : 16, rt.size());
            // the retry uses its own scratch sizing; rebuild the regular scratch afterwards
            const uint32_t keep_grid = ws->grid, keep_q = ws->qcap;
            CKS(ensure_scratch(E, ws, rgrid, big_q));
            DBuf rtiles;
            CKS(rtiles.ensure(rt.size() * sizeof(uint4)));
            CK(cudaMemcpyAsync(rtiles.p, rt.data(), rt.size() * sizeof(uint4), cudaMemcpyHostToDevice, s));
            ExpandParams Q = P;
            Q.mode = 1; Q.tiles = rtiles.as<uint4>(); Q.n_tiles = (uint32_t)rt.size(); Q.pass = 1;
            Q.queue = ws->queue.as<FacState>(); Q.nxt = ws->nxt.as<uint32_t>(); Q.hslot = ws->hslot.as<uint32_t>(); Q.qcap = big_q;
            Q.gtab_rep = ws->gtab_rep.as<uint32_t>(); Q.gtab_head = ws->gtab_head.as<uint32_t>(); Q.gtab_min = ws->gtab_min.as<float>();
            Q.gtab_size = ws->gtab_size;
            Q.failed_bitmap = nullptr; Q.failed_cap = 0;
            // keep counters[1] (candidates) running; reset tile / failed counters
            CK(cudaMemsetAsync(ws->counters.p, 0, 8, s));
            CK(cudaMemsetAsync((uint8_t *)ws->counters.p + 24, 0, 8, s));
            CK(cudaEventRecord(ws->evk0, s));
            if (R.beam) { k_expand_beam<256><<<rgrid, 256, 0, s>>>(Q, R.bw); CK(cudaGetLastError()); }
            else CKS(launch_expand(Q, rgrid, smem, s, fast_run));
            CK(cudaEventRecord(ws->evk1, s));
            stats.launches++;
            CK(cudaMemcpyAsync(ws->h_counters, ws->counters.p, 8 * 8, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            CK(cudaEventElapsedTime(&ms, ws->evk0, ws->evk1));
            stats.expand_ms += ms;
            rtiles.release();
            CKS(ensure_scratch(E, ws, keep_grid, keep_q));
            if (ws->h_counters[3]) {
                set_err("a start window expands more than " + std::to_string(R.beam ? big_q / 4 : big_q) + " states; lower the edit limits / raise the threshold or use a beam");
                return FAC_UNSUPPORTED;
            }
            if (ws->h_counters[1] > cand_cap) {
                cand_cap = (uint32_t)std::min<uint64_t>(ws->h_counters[1] * 2, 0x7FFFFFF0u);
                CKS(ws->cands.ensure((size_t)cand_cap * sizeof(FacCand)));
                continue;
            }
            n_cand = ws->h_counters[1];
            states = ws->h_counters[2];
        }
        uint32_t sg_end = 0xFFFFFFFFu;
        if (R.limit_after_expand) sg_end = R.limit_after_expand(states);
        else if (R.count_states) stats.states += states;
        if (states_per_window_out && n_windows_total) *states_per_window_out = (double)ws->h_counters[6] / (double)n_windows_total;
        const uint64_t n_overflowed = (use_succ || use_stack) ? ws->h_counters[7] : 0;
        if (n_cand == 0 && n_overflowed == 0) return FAC_OK;

        // ---- best-per-span reduction ----
        const uint32_t tab = next_pow2(n_cand * 2 + 16);
        CKS(ws->best_rep.ensure((size_t)tab * 4));
        CKS(ws->best_val.ensure((size_t)tab * 8));
        CKS(ws->cslot.ensure((size_t)n_cand * 4));
        CK(cudaMemsetAsync(ws->best_rep.p, 0xFF, (size_t)tab * 4, s));
        CK(cudaMemsetAsync(ws->best_val.p, 0, (size_t)tab * 8, s));
        CKS(grow_keep(ws->m_a, (size_t)*n_matches * sizeof(WMatch), (size_t)(*n_matches + n_cand) * sizeof(WMatch), s));
        BestParams B;
        memset(&B, 0, sizeof(B));
        B.cands = ws->cands.as<FacCand>(); B.n_cands = (uint32_t)n_cand;
        B.tab_rep = ws->best_rep.as<uint32_t>(); B.tab_val = ws->best_val.as<unsigned long long>(); B.tab_size = tab;
        B.cslot = ws->cslot.as<uint32_t>();
        B.failed_bitmap = n_failed ? ws->failed_bitmap.as<uint32_t>() : nullptr;
        B.tv = R.tv; B.windows = R.d_windows; B.sg_end = sg_end;
        B.out = ws->m_a.as<WMatch>() + *n_matches;
        B.out_cap = (uint32_t)n_cand;
        CK(cudaMemsetAsync((uint8_t *)ws->counters.p + 32, 0, 8, s));
        B.out_count = ws->counters.as<unsigned long long>() + 4;
        if (!fast_run) {
            k_best_insert<<<cdiv(n_cand, 256), 256, 0, s>>>(B);
            k_best_select<<<cdiv(n_cand, 256), 256, 0, s>>>(B);
            CK(cudaGetLastError());
            stats.launches += 2;
            CK(cudaMemcpyAsync(ws->h_counters, ws->counters.p, 8 * 8, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            *n_matches += ws->h_counters[4];
            return FAC_OK;
        }
        // ---- FAST reduction: order-independent maximum, tie detection, faithful redo of dirty windows ----
        const uint32_t n_win = R.seg_end - R.seg_begin;
        const uint32_t dwords = n_win / 32 + 1;
        CKS(ws->tab_sim.ensure((size_t)tab * 4)); CKS(ws->tab_cmin.ensure((size_t)tab * 4));
        CKS(ws->tab_cmax.ensure((size_t)tab * 4)); CKS(ws->tab_first.ensure((size_t)tab * 4));
        CK(cudaMemsetAsync(ws->tab_sim.p, 0, (size_t)tab * 4, s));
        CK(cudaMemsetAsync(ws->tab_cmin.p, 0xFF, (size_t)tab * 4, s));
        CK(cudaMemsetAsync(ws->tab_cmax.p, 0, (size_t)tab * 4, s));
        CK(cudaMemsetAsync(ws->tab_first.p, 0xFF, (size_t)tab * 4, s));
        FBestParams F;
        F.B = B; F.tab_sim = ws->tab_sim.as<uint32_t>(); F.tab_cmin = ws->tab_cmin.as<uint32_t>(); F.tab_cmax = ws->tab_cmax.as<uint32_t>();
        F.tab_first = ws->tab_first.as<uint32_t>(); F.dirty = ws->dirty.as<uint32_t>(); F.dirty_base = R.seg_begin;
        const uint32_t gb = cdiv(n_cand, 256);
        if (gb) {
            k_fbest_max<<<gb, 256, 0, s>>>(F);
            k_fbest_minmax<<<gb, 256, 0, s>>>(F);
            k_fbest_mark<<<gb, 256, 0, s>>>(F);
            k_fbest_emit<<<gb, 256, 0, s>>>(F);
        }
        CKS(ws->tiles.ensure((size_t)n_win * sizeof(uint4) / 8 + 4096));  // dirty windows are rare; overflow is re-run below
        const uint32_t dcap = (uint32_t)(ws->tiles.cap / sizeof(uint4));
        CK(cudaMemsetAsync((uint8_t *)ws->counters.p + 40, 0, 8, s));
        k_dirty_tiles<<<cdiv(dwords, 256), 256, 0, s>>>(ws->dirty.as<uint32_t>(), dwords, R.seg_begin, R.text_end, ws->tiles.as<uint4>(), dcap,
                                                        ws->counters.as<unsigned long long>() + 5);
        CK(cudaGetLastError());
        stats.launches += 5;
        CK(cudaMemcpyAsync(ws->h_counters, ws->counters.p, 8 * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        *n_matches += ws->h_counters[4];
        uint64_t n_dirty = ws->h_counters[5];
        if (n_dirty) {
            if (n_dirty > dcap) {  // enlarge and list again
                CKS(ws->tiles.ensure((size_t)n_dirty * sizeof(uint4)));
                CK(cudaMemsetAsync((uint8_t *)ws->counters.p + 40, 0, 8, s));
                k_dirty_tiles<<<cdiv(dwords, 256), 256, 0, s>>>(ws->dirty.as<uint32_t>(), dwords, R.seg_begin, R.text_end, ws->tiles.as<uint4>(),
                                                                (uint32_t)n_dirty, ws->counters.as<unsigned long long>() + 5);
                CK(cudaStreamSynchronize(s));
            }
            ExpandRun R2;
            R2.tv = R.tv; R2.text_end = R.text_end; R2.d_windows = R.d_windows; R2.thr = R.thr; R2.fast = false; R2.count_states = false;
            R2.tiles.resize(n_dirty);
            CK(cudaMemcpy(R2.tiles.data(), ws->tiles.p, n_dirty * sizeof(uint4), cudaMemcpyDeviceToHost));
            if (R.slices)
                for (uint4 &t4 : R2.tiles) {  // the window's haystack ends where its slice ends
                    auto it = std::upper_bound(R.slices->begin(), R.slices->end(), std::make_pair(t4.x, 0xFFFFFFFFu));
                    --it;
                    t4.z = it->second;
                    if (R.slice_tags) t4.w = (uint32_t)(it - R.slices->begin());
                }
            stats.dirty_windows += n_dirty;
            CKS(expand_and_reduce(E, ws, R2, 1, n_matches, stats, nullptr));
        }
        return FAC_OK;
    }
    set_err("candidate buffer kept overflowing");
    return FAC_UNSUPPORTED;
}

template <class Cmp, class T>
fac_status merge_sort(Workspace *ws, T *d, uint32_t n, Cmp cmp) {
    size_t tb = 0;
    CK(cub::DeviceMergeSort::SortKeys((void *)nullptr, tb, d, (int)n, cmp, ws->stream));
    CKS(ws->cubtmp.ensure(tb));
    CK(cub::DeviceMergeSort::SortKeys(ws->cubtmp.p, tb, d, (int)n, cmp, ws->stream));
    return FAC_OK;
}

// Order::Unsorted (ascending (start, end, pattern), oracle rule U2) of a single-window list as an LSD radix sort of
// (64-bit key, index) pairs followed by one gather of the 32-byte records: ~4x less DRAM traffic than the comparison
// merge sort of the records themselves on the 10^8-match lists of dense workloads.  *done = false when the key fields do
// not fit 64 bits (the caller falls back to the merge sort).
fac_status radix_sort_unsorted(Workspace *ws, uint32_t n, bool *done, SearchStats &stats) {
    cudaStream_t s = ws->stream;
    *done = false;
    CKS(ws->misc.ensure(64));
    CK(cudaMemsetAsync(ws->misc.p, 0, 32, s));
    k_key_bounds<<<std::min<uint32_t>(cdiv(n, 256), 148 * 16), 256, 0, s>>>(ws->m_a.as<WMatch>(), n, ws->misc.as<unsigned long long>());
    unsigned long long hb[4];
    CK(cudaMemcpyAsync(hb, ws->misc.p, 32, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    auto bits = [](unsigned long long v) { uint32_t b = 0; while (v) { b++; v >>= 1; } return b; };
    const uint32_t pat_bits = std::max(1u, bits(hb[2])), len_bits = std::max(1u, bits(hb[1])), start_bits = std::max(1u, bits(hb[0]));
    if (hb[3] != 0 || pat_bits + len_bits + start_bits > 64) return FAC_OK;
    CKS(ws->keys_a.ensure((size_t)n * 8)); CKS(ws->keys_b.ensure((size_t)n * 8));
    CKS(ws->idx_a.ensure((size_t)n * 4)); CKS(ws->idx_b.ensure((size_t)n * 4));
    CKS(ws->m_b.ensure((size_t)n * sizeof(WMatch)));
    k_make_keys<<<cdiv(n, 256), 256, 0, s>>>(ws->m_a.as<WMatch>(), n, pat_bits, pat_bits + len_bits, ws->keys_a.as<unsigned long long>(), ws->idx_a.as<uint32_t>());
    cub::DoubleBuffer<unsigned long long> dk(ws->keys_a.as<unsigned long long>(), ws->keys_b.as<unsigned long long>());
    cub::DoubleBuffer<uint32_t> dv(ws->idx_a.as<uint32_t>(), ws->idx_b.as<uint32_t>());
    size_t tb = 0;
    const int end_bit = (int)(pat_bits + len_bits + start_bits);
    CK(cub::DeviceRadixSort::SortPairs((void *)nullptr, tb, dk, dv, (int)n, 0, end_bit, s));
    CKS(ws->cubtmp.ensure(tb));
    CK(cub::DeviceRadixSort::SortPairs(ws->cubtmp.p, tb, dk, dv, (int)n, 0, end_bit, s));
    k_gather_matches16<<<cdiv((uint64_t)n * 2, 256), 256, 0, s>>>(ws->m_a.as<WMatch>(), dv.Current(), n, ws->m_b.as<WMatch>());
    CK(cudaGetLastError());
    std::swap(ws->m_a, ws->m_b);
    stats.launches += 4 + (uint32_t)((end_bit + 7) / 8);
    *done = true;
    return FAC_OK;
}

// FuzzyMatches::apply on the device.  In: ws->m_a[0..n) (any order).  Out: ws->m_a[0..*n_out) in final order.
fac_status apply_device(const fac_engine *E, Workspace *ws, uint32_t n, int order, int overlap, uint32_t n_windows, uint32_t *n_out,
                        SearchStats &stats, bool presorted = false) {
    cudaStream_t s = ws->stream;
    *n_out = n;
    if (n == 0) return FAC_OK;
    RankLess rl{order, E->d_pat_bytes};
    if (!presorted) {
        bool done = false;
        if (order == 0 && n >= 4096u && E->radix_unsorted) CKS(radix_sort_unsorted(ws, n, &done, stats));
        if (!done) {
            CKS(merge_sort(ws, ws->m_a.as<WMatch>(), n, rl));
            stats.launches += 2;
        }
    }
    if (overlap == FAC_OVERLAP_KEEP) return FAC_OK;
    WMatch *R = ws->m_a.as<WMatch>();
    // position order
    CKS(ws->idx_a.ensure((size_t)n * 4));
    CKS(ws->winend.ensure((size_t)n * sizeof(WinEnd)));
    CKS(ws->st_a.ensure(n)); CKS(ws->st_b.ensure(n)); CKS(ws->flags8.ensure(n));
    CKS(ws->sel.ensure((size_t)n * 4)); CKS(ws->nsel.ensure(16)); CKS(ws->m_b.ensure((size_t)n * sizeof(WMatch)));
    uint32_t *Pp = ws->idx_a.as<uint32_t>();
    k_iota<<<cdiv(n, 256), 256, 0, s>>>(Pp, n);
    CKS(merge_sort(ws, Pp, n, PosLess{R}));
    k_gather_winend<<<cdiv(n, 256), 256, 0, s>>>(R, Pp, ws->winend.as<WinEnd>(), n);
    {
        size_t tb = 0;
        CK(cub::DeviceScan::InclusiveScan((void *)nullptr, tb, ws->winend.as<WinEnd>(), ws->winend.as<WinEnd>(), WinEndMax(), (int)n, s));
        CKS(ws->cubtmp.ensure(tb));
        CK(cub::DeviceScan::InclusiveScan(ws->cubtmp.p, tb, ws->winend.as<WinEnd>(), ws->winend.as<WinEnd>(), WinEndMax(), (int)n, s));
    }
    stats.launches += 5;
    uint8_t *st_final = nullptr;
    if (overlap == FAC_OVERLAP_NON_OVERLAPPING) {
        CK(cudaMemsetAsync(ws->st_a.p, 0, n, s));
        uint8_t *a = ws->st_a.as<uint8_t>(), *b = ws->st_b.as<uint8_t>();
        CKS(ws->misc.ensure(64));
        for (uint32_t round = 0;; round++) {
            CK(cudaMemsetAsync(ws->misc.p, 0, 4, s));
            // a few rounds per host check
            for (int k = 0; k < 4; k++) {
                k_overlap_round<<<cdiv(n, 256), 256, 0, s>>>(R, Pp, ws->winend.as<WinEnd>(), a, b, n, ws->misc.as<uint32_t>());
                std::swap(a, b);
                stats.launches++;
                if (k < 3) CK(cudaMemsetAsync(ws->misc.p, 0, 4, s));
            }
            CK(cudaMemcpyAsync(ws->h_flags, ws->misc.p, 4, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            if (ws->h_flags[0] == 0) break;
            if (round > n) { set_err("overlap resolution did not converge"); return FAC_CUDA_ERROR; }
        }
        st_final = a;
    } else {
        CKS(ws->idx_b.ensure((size_t)n * 4));
        CKS(ws->misc.ensure((size_t)(n_windows + 2) * 4));
        const uint32_t uid_words = E->n_uid / 32 + 1;
        const uint32_t ugrid = std::min<uint32_t>(n_windows, 1024);
        const bool fresh = ws->used.cap < (size_t)1024 * uid_words * 4;
        CKS(ws->used.ensure((size_t)1024 * uid_words * 4));
        if (fresh) CK(cudaMemsetAsync(ws->used.p, 0, ws->used.cap, s));
        k_posof<<<cdiv(n, 256), 256, 0, s>>>(Pp, ws->idx_b.as<uint32_t>(), n);
        k_win_rank_off<<<cdiv(n + 1, 256), 256, 0, s>>>(R, n, ws->misc.as<uint32_t>(), n_windows);
        CK(cudaMemsetAsync(ws->st_a.p, 0, n, s));
        k_unique_select<<<ugrid, 32, 0, s>>>(R, Pp, ws->idx_b.as<uint32_t>(), ws->winend.as<WinEnd>(), ws->misc.as<uint32_t>(), n_windows, n,
                                             E->d_pat_uid_dense, uid_words, ws->used.as<uint32_t>(), ws->st_a.as<uint8_t>());
        stats.launches += 3;
        st_final = ws->st_a.as<uint8_t>();
    }
    k_accept_flags<<<cdiv(n, 256), 256, 0, s>>>(st_final, ws->flags8.as<uint8_t>(), n);
    {
        size_t tb = 0;
        CK(cub::DeviceSelect::Flagged((void *)nullptr, tb, Pp, ws->flags8.as<uint8_t>(), ws->sel.as<uint32_t>(), ws->nsel.as<int>(), (int)n, s));
        CKS(ws->cubtmp.ensure(tb));
        CK(cub::DeviceSelect::Flagged(ws->cubtmp.p, tb, Pp, ws->flags8.as<uint8_t>(), ws->sel.as<uint32_t>(), ws->nsel.as<int>(), (int)n, s));
    }
    CK(cudaMemcpyAsync(ws->h_flags, ws->nsel.p, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const uint32_t kept = ws->h_flags[0];
    if (kept) {
        k_gather_matches<<<cdiv(kept, 256), 256, 0, s>>>(R, ws->sel.as<uint32_t>(), kept, ws->m_b.as<WMatch>());
        CK(cudaMemcpyAsync(R, ws->m_b.p, (size_t)kept * sizeof(WMatch), cudaMemcpyDeviceToDevice, s));
    }
    stats.launches += 4;
    CK(cudaGetLastError());
    *n_out = kept;
    return FAC_OK;
}

// WMatch list -> fac_match on the host (optionally dropping matches a stream window does not own).
fac_status finalize_and_fetch(Workspace *ws, uint32_t n, const FacWindow *d_windows, bool filter_commit, MatchVec &out,
                              SearchStats &stats, fac_matches *dev_sink = nullptr) {
    cudaStream_t s = ws->stream;
    if (n == 0) return FAC_OK;
    if (dev_sink && !filter_commit) {
        // the finalised list is written straight into a buffer the result handle owns (recycled through the pool)
        const size_t bytes = (size_t)n * sizeof(fac_match);
        if (!device_pool().get(dev_sink->device, bytes, dev_sink->d_out)) CKS(dev_sink->d_out.ensure(bytes));
        CKS(ws->keep8.ensure(n));
        k_finalize<<<cdiv(n, 256), 256, 0, s>>>(ws->m_a.as<WMatch>(), n, d_windows, dev_sink->d_out.as<fac_match>(), ws->keep8.as<uint8_t>());
        CK(cudaGetLastError());
        stats.launches++;
        dev_sink->n_dev = n; dev_sink->on_device = true;
        CK(cudaStreamSynchronize(s));
        return FAC_OK;
    }
    CKS(ws->outm.ensure((size_t)n * sizeof(fac_match)));
    CKS(ws->keep8.ensure(n));
    k_finalize<<<cdiv(n, 256), 256, 0, s>>>(ws->m_a.as<WMatch>(), n, d_windows, ws->outm.as<fac_match>(), ws->keep8.as<uint8_t>());
    stats.launches++;
    fac_match *src = ws->outm.as<fac_match>();
    uint32_t cnt = n;
    if (filter_commit) {
        CKS(ws->m_b.ensure((size_t)n * sizeof(fac_match)));
        CKS(ws->nsel.ensure(16));
        size_t tb = 0;
        CK(cub::DeviceSelect::Flagged((void *)nullptr, tb, src, ws->keep8.as<uint8_t>(), ws->m_b.as<fac_match>(), ws->nsel.as<int>(), (int)n, s));
        CKS(ws->cubtmp.ensure(tb));
        CK(cub::DeviceSelect::Flagged(ws->cubtmp.p, tb, src, ws->keep8.as<uint8_t>(), ws->m_b.as<fac_match>(), ws->nsel.as<int>(), (int)n, s));
        CK(cudaMemcpyAsync(ws->h_flags, ws->nsel.p, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        cnt = ws->h_flags[0];
        src = ws->m_b.as<fac_match>();
        stats.launches += 2;
    }
    if (dev_sink) {  // filtered list (stream windows): copy the kept records into the handle's buffer
        const size_t bytes = (size_t)std::max<uint32_t>(cnt, 1) * sizeof(fac_match);
        if (!device_pool().get(dev_sink->device, bytes, dev_sink->d_out)) CKS(dev_sink->d_out.ensure(bytes));
        if (cnt) CK(cudaMemcpyAsync(dev_sink->d_out.p, src, (size_t)cnt * sizeof(fac_match), cudaMemcpyDeviceToDevice, s));
        dev_sink->n_dev = cnt; dev_sink->on_device = true;
        CK(cudaStreamSynchronize(s));
        return FAC_OK;
    }
    const size_t old = out.size();
    out.resize(old + cnt);
    if (cnt) CK(cudaMemcpyAsync(out.data() + old, src, (size_t)cnt * sizeof(fac_match), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return FAC_OK;
}

// Segment a non-ASCII text range on the device (K1).  Fills ws->first/gid/off and *n_graphemes.
// Byte offsets are 32-bit for haystacks below 4 GiB and 64-bit above (*wide_out): the reference's limit is the number
// of graphemes (u32 positions, src/search.rs:198-202, 296-300), not the number of bytes.
fac_status segment_device(const fac_engine *E, Workspace *ws, const uint8_t *d_text, uint64_t len, bool want_pf, uint64_t *n_graphemes,
                          SearchStats &stats, bool *wide_out) {
    cudaStream_t s = ws->stream;
    const bool wide = len >= 0xFFFFFFF0ull;
    *wide_out = wide;
    CKS(ws->misc.ensure(64));
    CK(cudaMemsetAsync(ws->misc.p, 0, 8, s));
    k_validate_utf8<<<cdiv(len, 256), 256, 0, s>>>(d_text, len, ws->misc.as<uint32_t>());
    CKS(ws->mark.ensure(len + 16));
    CKS(ws->gidx.ensure((len + 16) * 4));
    k_seg_mark<<<cdiv(len, 256), 256, 0, s>>>(d_text, 0, len, E->dU, ws->mark.as<uint8_t>());
    {
        size_t tb = 0;
        cub::TransformInputIterator<uint32_t, U8ToU32, const uint8_t *> it(ws->mark.as<uint8_t>(), U8ToU32());
        CK(cub::DeviceScan::ExclusiveSum((void *)nullptr, tb, it, ws->gidx.as<uint32_t>(), (int64_t)len + 1, s));
        CKS(ws->cubtmp.ensure(tb));
        CK(cudaMemsetAsync(ws->mark.as<uint8_t>() + len, 0, 1, s));
        CK(cub::DeviceScan::ExclusiveSum(ws->cubtmp.p, tb, it, ws->gidx.as<uint32_t>(), (int64_t)len + 1, s));
    }
    if (wide) {  // the 32-bit scan may wrap: the true cluster count decides HaystackTooLarge
        size_t tb = 0;
        cub::TransformInputIterator<unsigned long long, U8ToU64, const uint8_t *> it64(ws->mark.as<uint8_t>(), U8ToU64());
        CK(cub::DeviceReduce::Sum((void *)nullptr, tb, it64, ws->misc.as<unsigned long long>() + 1, (int64_t)len, s));
        CKS(ws->cubtmp.ensure(tb));
        CK(cub::DeviceReduce::Sum(ws->cubtmp.p, tb, it64, ws->misc.as<unsigned long long>() + 1, (int64_t)len, s));
        CK(cudaMemcpyAsync(ws->h_counters, ws->misc.as<unsigned long long>() + 1, 8, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaMemcpyAsync(ws->h_flags, ws->misc.p, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ws->h_flags + 1, ws->gidx.as<uint32_t>() + len, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    stats.launches += 4;
    if (ws->h_flags[0]) { set_err("haystack is not valid UTF-8"); return FAC_INVALID_UTF8; }
    const uint64_t n = wide ? ws->h_counters[0] : (uint64_t)ws->h_flags[1];
    *n_graphemes = n;
    if (n > 0xFFFFFFFFull) return FAC_OK;   // the caller reports SearchError::HaystackTooLarge
    CKS(ws->first.ensure((n + 4) * 4));
    CKS(ws->off.ensure((n + 4) * (wide ? 8 : 4)));
    if (E->host.has_mappings) CKS(ws->gid.ensure((n + 4) * 4));
    if (want_pf) CKS(ws->pfsym.ensure(n + 16));
    SegEmitParams P;
    memset(&P, 0, sizeof(P));
    P.s = d_text; P.lo = 0; P.hi = len; P.mark = ws->mark.as<uint8_t>(); P.gidx = ws->gidx.as<uint32_t>(); P.g_base = 0; P.U = E->dU;
    if (E->host.has_mappings) { P.symbols = E->d_symbols; P.sym_mask = E->sym_mask; P.pool = E->d_pool; }
    if (want_pf) { P.pf_symbols = E->d_pf_symbols; P.pf_mask = E->pf_mask; P.pf_pool = E->d_pf_pool; P.pf_sym = ws->pfsym.as<uint8_t>(); }
    P.fold = E->host.ci; P.first = ws->first.as<uint32_t>(); P.gid = ws->gid.as<uint32_t>();
    if (wide) P.off64 = ws->off.as<uint64_t>(); else P.off32 = ws->off.as<uint32_t>();
    k_seg_emit<<<cdiv(len, 256), 256, 0, s>>>(P);
    const uint32_t len32 = (uint32_t)len;
    const uint64_t len64 = len;
    if (wide) CK(cudaMemcpyAsync(ws->off.as<uint64_t>() + n, &len64, 8, cudaMemcpyHostToDevice, s));
    else CK(cudaMemcpyAsync(ws->off.as<uint32_t>() + n, &len32, 4, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));  // the sentinel is a stack variable
    stats.launches++;
    CK(cudaGetLastError());
    return FAC_OK;
}

// Engines with a beam (src/search.rs:527-528, 577-589, 1096-1103).
//  * beam_width: every start window runs the beamed kernel.
//  * auto_beam(budget, width): windows are exact until the running total of queue.len() exceeds the
//    budget; the window that crosses is still exact, every later one is beamed.  Exact chunks are
//    run first with only their total; the chunk whose total crosses is redone with per-window
//    counts to find the crossing window, its later windows are dropped from the reduction.
fac_status search_beamed(const fac_engine *E, Workspace *ws, const TextView &tv, uint32_t g_begin, uint32_t g_end, float thr,
                         uint64_t *n_matches, SearchStats &stats) {
    cudaStream_t s = ws->stream;
    const uint64_t SEG = 1u << 22;
    auto run = [&](uint32_t a, uint32_t b, bool beamk, uint32_t bw, uint32_t *d_pw, std::function<uint32_t(uint64_t)> lim) -> fac_status {
        for (uint64_t pos = a; pos < b; pos += SEG) {
            ExpandRun R;
            R.tv = tv; R.seg_begin = (uint32_t)pos; R.seg_end = (uint32_t)std::min<uint64_t>(b, pos + SEG); R.text_end = tv.n;
            R.d_windows = ws->windows.as<FacWindow>(); R.thr = thr; R.beam = beamk; R.bw = bw; R.d_per_window = d_pw; R.limit_after_expand = lim;
            CKS(grow_keep(ws->m_a, *n_matches * sizeof(WMatch), (*n_matches + (1u << 20)) * sizeof(WMatch), s));
            CKS(expand_and_reduce(E, ws, R, beamk ? 1 : 8, n_matches, stats, nullptr));
        }
        return FAC_OK;
    };
    if (E->host.beam_width != 0) return run(g_begin, g_end, true, (uint32_t)std::min<uint64_t>(E->host.beam_width, 0x3FFFFFFFu), nullptr, nullptr);
    // auto_beam
    const uint64_t budget = E->host.ab_budget;
    const uint32_t width = (uint32_t)std::min<uint64_t>(E->host.ab_width, 0x3FFFFFFFu);
    uint64_t cum = 0;
    uint32_t pos = g_begin;
    uint32_t chunk = 1024;
    while (pos < g_end) {
        const uint32_t end = (uint32_t)std::min<uint64_t>(g_end, (uint64_t)pos + chunk);
        // 1) exact pass over [pos, end) with the tiled kernel; only the chunk total is needed unless it crosses
        bool crosses = false;
        uint64_t chunk_states = 0;
        auto lim1 = [&](uint64_t states) -> uint32_t {
            chunk_states = states;
            if (cum + states <= budget) return 0xFFFFFFFFu;  // the running total is monotone: no crossing inside
            crosses = true;
            return pos;  // discard this chunk's candidates; it is redone below with per-window counts
        };
        CKS(run(pos, end, false, 0, nullptr, lim1));
        if (!crosses) {
            stats.states += chunk_states;
            cum += chunk_states;
            pos = end;
            if (chunk < (1u << 20)) chunk *= 4;
            continue;
        }
        // 2) redo the chunk one window per CTA, recording every window's queue.len(), to find the crossing window
        CKS(ws->pfsym.ensure((size_t)(end - pos) * 4 + 16));
        CK(cudaMemsetAsync(ws->pfsym.p, 0, (size_t)(end - pos) * 4, s));
        uint32_t *d_pw = ws->pfsym.as<uint32_t>() - pos;  // the kernel indexes by absolute start window
        uint32_t crossing = end - 1;
        auto lim2 = [&](uint64_t) -> uint32_t {
            std::vector<uint32_t> pw(end - pos);
            cudaMemcpy(pw.data(), ws->pfsym.p, pw.size() * 4, cudaMemcpyDeviceToHost);
            uint64_t c = cum;
            for (uint32_t w = 0; w < pw.size(); w++) {
                c += pw[w];
                if (c > budget) { crossing = pos + w; break; }
            }
            stats.states += c - cum;
            return crossing + 1;
        };
        CKS(run(pos, end, true, 0, d_pw, lim2));
        if (crossing + 1 < g_end) CKS(run(crossing + 1, g_end, true, width, nullptr, nullptr));
        return FAC_OK;
    }
    return FAC_OK;
}

// ---- K2: bitap pre-filter (src/prefilter.rs:285-343) on an ASCII haystack -----------------------------
// k_for (prefilter.rs:285-302): false => the call falls back to the plain search.
bool bitap_k_for(const fac::HostBitap &B, size_t p, float thr, uint32_t &k_out) {
    const float n = (float)B.m[p];
    const float p_max = n * (1.0f - thr / B.weight[p]);
    uint64_t k_pen;
    if (p_max <= 0.0f) k_pen = 0;
    else {
        const float v = std::floor(p_max * B.edit_cost_mult);
        if (std::isnan(v)) k_pen = 0; else if (v >= 1.8e19f) k_pen = ~0ull; else k_pen = (uint64_t)v;  // Rust `as usize` saturates
    }
    const uint64_t k = B.k_limit[p] >= 0 ? std::min<uint64_t>(k_pen, (uint64_t)B.k_limit[p]) : k_pen;
    if (k > 24) return false;
    k_out = (uint32_t)k;
    return true;
}

template <int KMAX>
void launch_bitap_t(const BitapParams &P, dim3 grid, cudaStream_t s) {
    const size_t smem = (size_t)P.rows * 32 * sizeof(uint64_t);
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_bitap_scan<KMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_bitap_scan<KMAX><<<grid, BITAP_WARPS * 32, smem, s>>>(P);
}

// Scans, merges and returns the slices (gs, ge) in ascending order.  *fallback = true when a pattern's
// budget exceeds MAX_USEFUL_K (the reference then runs the plain search).
// `d_stream` = one byte per grapheme: the ASCII haystack itself, or (sym_stream) the K1 symbol-id stream.
fac_status prefilter_slices(const fac_engine *E, Workspace *ws, const uint8_t *d_stream, bool sym_stream, uint32_t n, float thr,
                            std::vector<std::pair<uint32_t, uint32_t>> &slices, bool *fallback, SearchStats &stats) {
    cudaStream_t s = ws->stream;
    const fac::HostBitap &B = E->host.bitap;
    const size_t P = B.m.size();
    std::vector<uint8_t> ks(P);
    uint32_t kmax = 0, warm = 0;
    *fallback = false;
    for (size_t p = 0; p < P; p++) {
        uint32_t k;
        if (!bitap_k_for(B, p, thr, k)) { *fallback = true; return FAC_OK; }
        ks[p] = (uint8_t)k; kmax = std::max(kmax, k); warm = std::max(warm, B.m[p] + k);
    }
    const uint32_t n_words = n / 32 + 1;
    CKS(ws->bp_k.ensure(P + 16));
    CKS(ws->cov.ensure((size_t)(n_words + 2) * 4));
    CKS(ws->covcnt.ensure((size_t)n_words * 4));
    CKS(ws->covoff.ensure((size_t)(n_words + 1) * 8));
    CKS(ws->misc.ensure(64));
    CK(cudaMemcpyAsync(ws->bp_k.p, ks.data(), P, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(ws->cov.p, 0, (size_t)(n_words + 2) * 4, s));
    CK(cudaMemsetAsync(ws->misc.p, 0, 16, s));
    BitapParams BP;
    BP.text = d_stream; BP.n = n; BP.m = E->d_bp_m; BP.k = ws->bp_k.as<uint8_t>();
    if (sym_stream) { BP.bytemask = E->d_bp_symmask; BP.rows = B.alphabet + 1; }
    else { BP.bytemask = E->d_bp_mask; BP.rows = 128; }
    BP.n_patterns = (uint32_t)P; BP.warm = warm; BP.cov = ws->cov.as<uint32_t>(); BP.hits = ws->misc.as<unsigned long long>();
    const dim3 grid(cdiv(n, (uint64_t)BITAP_SUB * BITAP_WARPS), cdiv(P, 32));
    if (kmax <= 2) launch_bitap_t<2>(BP, grid, s);
    else if (kmax <= 4) launch_bitap_t<4>(BP, grid, s);
    else if (kmax <= 8) launch_bitap_t<8>(BP, grid, s);
    else launch_bitap_t<24>(BP, grid, s);
    CK(cudaGetLastError());
    k_cov_edges<<<cdiv(n_words, 256), 256, 0, s>>>(ws->cov.as<uint32_t>(), n_words, ws->covcnt.as<uint32_t>());
    {
        size_t tb = 0;
        cub::TransformInputIterator<unsigned long long, CovCountToU64, const uint32_t *> it(ws->covcnt.as<uint32_t>(), CovCountToU64());
        CK(cub::DeviceScan::ExclusiveSum((void *)nullptr, tb, it, ws->covoff.as<unsigned long long>(), (int64_t)n_words, s));
        CKS(ws->cubtmp.ensure(tb));
        CK(cub::DeviceScan::ExclusiveSum(ws->cubtmp.p, tb, it, ws->covoff.as<unsigned long long>(), (int64_t)n_words, s));
    }
    // total = off[last] + cnt[last]
    unsigned long long last_off = 0; uint32_t last_cnt = 0;
    CK(cudaMemcpyAsync(&last_off, ws->covoff.as<unsigned long long>() + (n_words - 1), 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&last_cnt, ws->covcnt.as<uint32_t>() + (n_words - 1), 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ws->h_counters, ws->misc.p, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    stats.launches += 3;
    stats.pf_hits = ws->h_counters[0];
    const uint32_t n_runs = (uint32_t)(last_off & 0xFFFFFFFFull) + (last_cnt & 0xFFFFu);
    slices.clear();
    if (n_runs == 0) return FAC_OK;
    CKS(ws->sl_gs.ensure((size_t)n_runs * 4));
    CKS(ws->sl_ge.ensure((size_t)n_runs * 4));
    k_cov_emit<<<cdiv(n_words, 256), 256, 0, s>>>(ws->cov.as<uint32_t>(), n_words, ws->covoff.as<unsigned long long>(), ws->sl_gs.as<uint32_t>(),
                                                  ws->sl_ge.as<uint32_t>());
    CK(cudaGetLastError());
    std::vector<uint32_t> gs(n_runs), ge(n_runs);
    CK(cudaMemcpyAsync(gs.data(), ws->sl_gs.p, (size_t)n_runs * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ge.data(), ws->sl_ge.p, (size_t)n_runs * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    stats.launches++;
    slices.resize(n_runs);
    for (uint32_t i = 0; i < n_runs; i++) slices[i] = {gs[i], std::min(ge[i], n)};  // ge.min(n), prefilter.rs:348
    stats.pf_slices = n_runs;
    return FAC_OK;
}

// The whole-haystack search on device-resident text: classification, (K1), K3 over segments of
// start windows, reduction, apply.  Restricts start windows to the byte range [own_begin, own_end).
fac_status search_resident(const fac_engine *E, Workspace *ws, const uint8_t *d_text, uint64_t len, float thr, int order, int overlap,
                           uint64_t own_begin, uint64_t own_end, uint64_t base, uint64_t commit, bool apply, MatchVec &out, SearchStats &stats,
                           bool use_prefilter = false, fac_matches *dev_sink = nullptr, bool force_unicode = false) {
    cudaStream_t s = ws->stream;
    if (len == 0) return FAC_OK;
    CKS(ws->misc.ensure(64));
    bool ascii = false;
    if (!force_unicode) {
        CK(cudaMemsetAsync(ws->misc.p, 0, 4, s));
        k_scan_bytes<<<(unsigned)std::min<uint64_t>(cdiv(len, 256 * 64), 148 * 8), 256, 0, s>>>(d_text, len, ws->misc.as<uint32_t>());
        CK(cudaMemcpyAsync(ws->h_flags, ws->misc.p, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        stats.launches++;
        ascii = ws->h_flags[0] == 0;
    }
    // the pre-filter applies to whole-haystack calls of beam-less engines whose configuration reduced to the bit model
    const bool want_pf = use_prefilter && E->host.bitap.active && own_begin == 0 && own_end >= len && E->host.beam_width == 0 && !E->host.has_auto_beam;
    TextView tv;
    memset(&tv, 0, sizeof(tv));
    tv.n_bytes = len; tv.ascii = ascii;
    uint64_t n = len;
    if (ascii) tv.bytes = d_text;
    else {
        bool wide = false;
        CKS(segment_device(E, ws, d_text, len, want_pf, &n, stats, &wide));
        tv.first = ws->first.as<uint32_t>(); tv.gid = ws->gid.as<uint32_t>();
        if (wide) tv.off64 = ws->off.as<uint64_t>(); else tv.off32 = ws->off.as<uint32_t>();
    }
    if (n > 0xFFFFFFFFull) {  // SearchError::HaystackTooLarge, search.rs:198-202
        g_last_graphemes = n;
        set_err("haystack has " + std::to_string(n) + " grapheme clusters, exceeding the u32 position space");
        return FAC_HAYSTACK_TOO_LARGE;
    }
    tv.n = (uint32_t)n;
    if (n == 0) return FAC_OK;
    // owned grapheme range
    uint32_t g_begin = 0, g_end = (uint32_t)n;
    if (own_begin > 0 || own_end < len) {
        if (ascii) { g_begin = (uint32_t)std::min<uint64_t>(own_begin, n); g_end = (uint32_t)std::min<uint64_t>(own_end, n); }
        else {
            // first grapheme with offset >= own_begin / own_end
            if (tv.off64) {
                std::vector<uint64_t> off(n + 1);
                CK(cudaMemcpy(off.data(), ws->off.p, (n + 1) * 8, cudaMemcpyDeviceToHost));
                g_begin = (uint32_t)(std::lower_bound(off.begin(), off.begin() + n, std::min<uint64_t>(own_begin, len)) - off.begin());
                g_end = (uint32_t)(std::lower_bound(off.begin(), off.begin() + n, std::min<uint64_t>(own_end, len)) - off.begin());
            } else {
                std::vector<uint32_t> off(n + 1);
                CK(cudaMemcpy(off.data(), ws->off.p, (n + 1) * 4, cudaMemcpyDeviceToHost));
                g_begin = (uint32_t)(std::lower_bound(off.begin(), off.begin() + n, (uint32_t)std::min<uint64_t>(own_begin, len)) - off.begin());
                g_end = (uint32_t)(std::lower_bound(off.begin(), off.begin() + n, (uint32_t)std::min<uint64_t>(own_end, len)) - off.begin());
            }
        }
    }
    FacWindow w;
    memset(&w, 0, sizeof(w));
    w.byte_begin = 0; w.byte_end = len; w.base = base; w.commit = commit; w.g_begin = 0; w.g_end = (uint32_t)n;
    CKS(ws->windows.ensure(sizeof(FacWindow)));
    CK(cudaMemcpyAsync(ws->windows.p, &w, sizeof(w), cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));

    uint64_t n_matches = 0;
    if (E->host.beam_width != 0 || E->host.has_auto_beam) {
        CKS(search_beamed(E, ws, tv, g_begin, g_end, thr, &n_matches, stats));
        uint32_t n_fin = (uint32_t)n_matches;
        if (apply) CKS(apply_device(E, ws, (uint32_t)n_matches, order, overlap, 1, &n_fin, stats));
        CKS(finalize_and_fetch(ws, n_fin, ws->windows.as<FacWindow>(), commit != ~0ull, out, stats, dev_sink));
        return FAC_OK;
    }
    // Prefiltered::search (src/prefilter.rs:135-155, 304-374): bitap scan -> merged slices -> the engine on every
    // slice as its own haystack.  ASCII haystacks are scanned byte by byte (Offsets::Identity); non-ASCII haystacks
    // through the symbol-id stream K1 emitted (transcode, prefilter.rs:262-281), slices then are grapheme ranges.
    // Engines with a beam answer with the plain search (result-neutral by the reference's contract, prefilter.rs:1-21).
    if (want_pf) {
        std::vector<std::pair<uint32_t, uint32_t>> slices;
        bool fallback = false;
        CKS(prefilter_slices(E, ws, ascii ? d_text : ws->pfsym.as<uint8_t>(), !ascii, (uint32_t)n, thr, slices, &fallback, stats));
        if (!fallback) {
            // Per-slice `is_ascii` (prefilter.rs:346-350 -> search.rs:196): an all-ASCII slice of a non-ASCII haystack is
            // searched with the byte-per-grapheme storage.  That differs from the K1 streams only when the slice holds a
            // CR LF pair (one cluster in UAX #29, two graphemes as bytes): those slices are searched from the bytes.
            std::vector<std::pair<uint32_t, uint32_t>> byte_slices;   // (byte begin, byte end) of ASCII slices holding CR LF
            if (!ascii && !slices.empty()) {
                const uint32_t ns = (uint32_t)slices.size();
                CKS(ws->misc.ensure((size_t)ns * 4 + 64));
                k_slice_classify<<<cdiv((uint64_t)ns * 32, 256), 256, 0, s>>>(d_text, tv.off32, tv.off64, ws->sl_gs.as<uint32_t>(), ws->sl_ge.as<uint32_t>(), ns,
                                                                             ws->misc.as<uint32_t>());
                CK(cudaGetLastError());
                std::vector<uint32_t> fl(ns);
                CK(cudaMemcpyAsync(fl.data(), ws->misc.p, (size_t)ns * 4, cudaMemcpyDeviceToHost, s));
                CK(cudaStreamSynchronize(s));
                stats.launches++;
                std::vector<uint32_t> special;
                for (uint32_t i = 0; i < ns; i++) if (fl[i] == 2u) special.push_back(i);
                if (!special.empty()) {
                    // byte ranges of those slices: read the two offsets of each back
                    std::vector<std::pair<uint32_t, uint32_t>> keep;
                    for (uint32_t i : special) {
                        uint64_t b0 = 0, b1 = 0;
                        if (tv.off64) {
                            CK(cudaMemcpy(&b0, tv.off64 + slices[i].first, 8, cudaMemcpyDeviceToHost));
                            CK(cudaMemcpy(&b1, tv.off64 + slices[i].second, 8, cudaMemcpyDeviceToHost));
                        } else {
                            uint32_t a0 = 0, a1 = 0;
                            CK(cudaMemcpy(&a0, tv.off32 + slices[i].first, 4, cudaMemcpyDeviceToHost));
                            CK(cudaMemcpy(&a1, tv.off32 + slices[i].second, 4, cudaMemcpyDeviceToHost));
                            b0 = a0; b1 = a1;
                        }
                        if (b1 > 0xFFFFFFF0ull) { set_err("an ASCII pre-filter slice lies beyond the 32-bit byte index space"); return FAC_UNSUPPORTED; }
                        byte_slices.emplace_back((uint32_t)b0, (uint32_t)b1);
                    }
                    size_t k = 0;
                    for (uint32_t i = 0; i < ns; i++) { if (k < special.size() && special[k] == i) { k++; continue; } keep.push_back(slices[i]); }
                    slices.swap(keep);
                }
            }
            auto run_slices = [&](const TextView &stv, const std::vector<std::pair<uint32_t, uint32_t>> &sl, uint32_t stream_len) -> fac_status {
                const bool s_ascii = stv.ascii != 0;
                const bool succ_here = (E->succ_ok || E->succ_generic_ok) && (s_ascii || E->host.succ.unicode_text_ok);
                const uint32_t tile_w = succ_here ? E->succ_tile : (E->stack_ok ? E->stack_tile : 64u);
                size_t si = 0;
                while (si < sl.size()) {
                    // batches of slices bounded by a window budget so the candidate buffers stay modest
                    ExpandRun R;
                    R.tv = stv; R.d_windows = ws->windows.as<FacWindow>(); R.thr = thr; R.fast = E->fast_ok || (E->succ_generic_ok && succ_here); R.slices = &sl;
                    uint64_t wins = 0;
                    const size_t s0 = si;
                    for (; si < sl.size() && wins < (1u << 25); si++) {
                        const uint32_t gs = sl[si].first, ge = sl[si].second;
                        for (uint32_t a = gs; a < ge; a += tile_w) R.tiles.push_back(make_uint4(a, std::min(tile_w, ge - a), ge, 0u));
                        wins += ge - gs;
                    }
                    R.seg_begin = sl[s0].first; R.seg_end = sl[si - 1].second; R.text_end = stream_len;
                    CKS(grow_keep(ws->m_a, n_matches * sizeof(WMatch), (n_matches + (1u << 20)) * sizeof(WMatch), s));
                    CKS(expand_and_reduce(E, ws, R, tile_w, &n_matches, stats, nullptr));
                }
                return FAC_OK;
            };
            CKS(run_slices(tv, slices, (uint32_t)n));
            if (!byte_slices.empty()) {
                TextView btv;
                memset(&btv, 0, sizeof(btv));
                btv.bytes = d_text; btv.n_bytes = len; btv.n = (uint32_t)std::min<uint64_t>(len, 0xFFFFFFF0ull); btv.ascii = 1;
                CKS(run_slices(btv, byte_slices, btv.n));
            }
            uint32_t n_final = (uint32_t)n_matches;
            if (apply) CKS(apply_device(E, ws, (uint32_t)n_matches, order, overlap, 1, &n_final, stats));
            CKS(finalize_and_fetch(ws, n_final, ws->windows.as<FacWindow>(), commit != ~0ull, out, stats, dev_sink));
            return FAC_OK;
        }
    }
    const uint64_t SEG = (uint64_t)env_int("FAC_SEGMENT_WINDOWS", 1 << 25);
    uint32_t tile = E->default_tile;
    bool calibrated = tile != 0;
    if ((E->succ_ok || E->succ_generic_ok) && (ascii || E->host.succ.unicode_text_ok)) { calibrated = true; if (!tile) tile = 8; }  // the succinct kernel tiles by itself
    if (E->stack_ok && !calibrated) { calibrated = true; tile = 8; }   // the stack kernel tiles by itself; 8 = faithful redo tile
    if (!calibrated) tile = 4;
    uint64_t pos = g_begin;
    while (pos < g_end) {
        uint64_t seg = std::min<uint64_t>(SEG, g_end - pos);
        if (!calibrated) seg = std::min<uint64_t>(seg, 1 << 16);  // small calibration segment decides the tile size
        ExpandRun R;
        R.tv = tv; R.seg_begin = (uint32_t)pos; R.seg_end = (uint32_t)(pos + seg); R.text_end = (uint32_t)n;
        R.d_windows = ws->windows.as<FacWindow>(); R.thr = thr; R.fast = E->fast_ok || (E->succ_generic_ok && (ascii || E->host.succ.unicode_text_ok));
        double spw = 0;
        // keep what is already in m_a: grow by copy before the reduction writes
        CKS(grow_keep(ws->m_a, n_matches * sizeof(WMatch), (n_matches + (1u << 20)) * sizeof(WMatch), s));
        const uint64_t before = n_matches;
        (void)before;
        CKS(expand_and_reduce(E, ws, R, tile, &n_matches, stats, &spw));
        if (!calibrated) {
            calibrated = true;
            // aim at ~1/6 of the queue per tile on average (levels are bursty); power of two, 1..256
            const double target = (double)E->qcap / 6.0;
            uint32_t t = 1;
            while (t < 256 && (double)(t * 2) * std::max(spw, 1.0) <= target) t <<= 1;
            tile = t;
        }
        pos += seg;
    }
    uint32_t n_final = (uint32_t)n_matches;
    if (apply) CKS(apply_device(E, ws, (uint32_t)n_matches, order, overlap, 1, &n_final, stats));
    CKS(finalize_and_fetch(ws, n_final, ws->windows.as<FacWindow>(), commit != ~0ull, out, stats, dev_sink));
    return FAC_OK;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char *fac_last_error_string(void) { return g_err.c_str(); }
int fac_abi_version(void) { return FAC_ABI_VERSION; }
#ifndef FAC_SOURCE_HASH
#define FAC_SOURCE_HASH "unstamped"
#endif
static const char fac_source_stamp[] = "FAC_SOURCE_STAMP:" FAC_SOURCE_HASH;   // also read straight from the file by build()
const char *fac_build_source_hash(void) { return fac_source_stamp + 17; }
uint64_t fac_last_haystack_graphemes(void) { return g_last_graphemes; }

fac_status fac_engine_create_on(int device, const fac_config *cfg, const fac_pattern *patterns, size_t n_patterns, fac_engine **out) {
    if (!out) { set_err("null out pointer"); return FAC_INVALID_ARGUMENT; }
    *out = nullptr;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        set_err(std::string("no usable CUDA device (") + cudaGetErrorString(ce) + "); this library has no CPU fallback");
        return FAC_CUDA_ERROR;
    }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= ndev) { set_err("device index out of range"); return FAC_INVALID_ARGUMENT; }
    CK(cudaSetDevice(device));
    fac_engine *E = new fac_engine();
    E->device = device;
    std::string err;
    fac_status st = fac::build_automaton(cfg, patterns, n_patterns, E->host, err);
    if (st != FAC_OK) { set_err(err); delete E; return st; }
    auto fail = [&](fac_status s) { fac_engine_free(E); return s; };
    const fac::HostAutomaton &H = E->host;
    AutomatonView &V = E->dview;
    V = H.host_view();
#define UP(field, vec) do { fac_status s_ = upload(E, vec, &V.field); if (s_ != FAC_OK) return fail(s_); } while (0)
    UP(node_edge_off, H.node_edge_off); UP(node_prune_len, H.node_prune_len); UP(node_prune_low, H.node_prune_low);
    UP(node_out_off, H.node_out_off); UP(node_bitmap, H.node_bitmap); UP(node_lim, H.node_lim); UP(node_map_off, H.node_map_off);
    UP(edge_char, H.edge_char); UP(edge_next, H.edge_next); UP(edge_sym, H.edge_sym); UP(trans, H.trans); UP(out_pat, H.out_pat);
    UP(pat_glen, H.pat_glen); UP(pat_weight, H.pat_weight); UP(pat_lim, H.pat_lim); UP(lim, H.lim);
    UP(sim_ascii, H.sim_ascii); UP(sim_keys, H.sim_keys); UP(sim_vals, H.sim_vals);
    UP(map_hay_off, H.map_hay_off); UP(map_hay_gid, H.map_hay_gid); UP(map_next, H.map_next); UP(map_pen, H.map_pen);
    UP(ascii_gid, H.ascii_gid);
#undef UP
    {
        const uint32_t *pb = nullptr;
        if ((st = upload(E, H.pat_bytes, &pb)) != FAC_OK) return fail(st);
        E->d_pat_bytes = (uint32_t *)pb;
        // dense pattern identities for non_overlapping_unique
        std::vector<int64_t> keys(H.pat_uid);
        std::sort(keys.begin(), keys.end());
        keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
        std::vector<uint32_t> dense(H.pat_uid.size());
        for (size_t i = 0; i < dense.size(); i++) dense[i] = (uint32_t)(std::lower_bound(keys.begin(), keys.end(), H.pat_uid[i]) - keys.begin());
        E->n_uid = (uint32_t)keys.size();
        if ((st = upload(E, dense, &pb)) != FAC_OK) return fail(st);
        E->d_pat_uid_dense = (uint32_t *)pb;
        const uint16_t *s1; const uint8_t *s2; const uint32_t *lk, *lv;
        if ((st = upload_raw(E, FAC_GCB_STAGE1, 0x1100, &s1)) != FAC_OK) return fail(st);
        if ((st = upload_raw(E, FAC_GCB_STAGE2, (size_t)FAC_GCB_NBLOCKS * 256, &s2)) != FAC_OK) return fail(st);
        if ((st = upload_raw(E, FAC_LOWER_KEYS, (size_t)FAC_LOWER_N, &lk)) != FAC_OK) return fail(st);
        if ((st = upload_raw(E, FAC_LOWER_VALS, (size_t)FAC_LOWER_N, &lv)) != FAC_OK) return fail(st);
        E->dU = UnicodeTables{s1, s2, lk, lv, FAC_LOWER_N};
        const fac::HostSymbol *sy; const uint8_t *pool;
        if ((st = upload(E, H.symbols, &sy)) != FAC_OK) return fail(st);
        if ((st = upload(E, H.symbol_pool, &pool)) != FAC_OK) return fail(st);
        E->d_symbols = (FacSymbol *)sy; E->sym_mask = (uint32_t)H.symbols.size() - 1; E->d_pool = (uint8_t *)pool;
    }
    if (H.bitap.active) {
        const fac::HostBitap &B = H.bitap;
        const size_t P = B.m.size();
        std::vector<uint64_t> bm(P * 128, 0);
        std::vector<uint8_t> mm(P);
        for (size_t pi = 0; pi < P; pi++) {
            mm[pi] = (uint8_t)B.m[pi];
            for (int b = 0; b < 128; b++) bm[pi * 128 + b] = B.masks[pi * (size_t)(B.alphabet + 1) + B.ascii_id[b]];
        }
        if ((st = upload(E, bm, &E->d_bp_mask)) != FAC_OK) return fail(st);
        if ((st = upload(E, mm, &E->d_bp_m)) != FAC_OK) return fail(st);
        if ((st = upload(E, B.masks, &E->d_bp_symmask)) != FAC_OK) return fail(st);
        const fac::HostSymbol *psy; const uint8_t *ppool;
        if ((st = upload(E, B.symbols, &psy)) != FAC_OK) return fail(st);
        if ((st = upload(E, B.symbol_pool, &ppool)) != FAC_OK) return fail(st);
        E->d_pf_symbols = (const FacSymbol *)psy; E->pf_mask = (uint32_t)B.symbols.size() - 1; E->d_pf_pool = ppool;
    }
    if (H.succ.ok) {
        const fac::HostSuccinct &S = H.succ;
        const size_t Nn = S.bm.size();
        std::vector<uint8_t> symof(S.sym_of, S.sym_of + 256);
        std::vector<uint32_t> fcsym(Nn);
        for (size_t i = 0; i < Nn; i++) fcsym[i] = S.fc[i] | ((uint32_t)S.insym[i] << (S.wide ? 26 : 27));
        // survivor masks, transposed so that the node index is fastest (SuccGMDev): gmT[y][node], then gm2T[y1][y2][node]
        const uint32_t ROWS = S.wide ? 64u : 32u, n1 = S.gm_nodes, n2 = S.gm2_nodes;
        const size_t off2 = (size_t)ROWS * n1;
        // deep tables (narrow layout only): gm3T[y1][y2][y3][node], pm3T[a][b][c][node], pm2T[a][b][node], pm4T[a][b][c][d][node]
        const uint32_t n3 = S.wide ? 0u : S.n3, np2 = S.wide ? 0u : S.np2, n4 = S.wide ? 0u : S.n4, R3 = S.r3;
        const size_t r2 = (size_t)R3 * R3, r3 = r2 * R3, r4 = r3 * R3;
        const size_t off3 = off2 + (size_t)ROWS * ROWS * n2, offp3 = off3 + r3 * n3, offp2 = offp3 + r3 * n3, offp4 = offp2 + r2 * np2;
        const size_t total = offp4 + r4 * n4;
        if (total >= 0xFFFFFFFFull) { set_err("survivor-mask tables exceed the 32-bit entry index"); return fail(FAC_UNSUPPORTED); }
        E->succ_gm2_off = (uint32_t)off2; E->succ_gm3_off = (uint32_t)off3; E->succ_pm3_off = (uint32_t)offp3; E->succ_pm2_off = (uint32_t)offp2;
        E->succ_pm4_off = (uint32_t)offp4;
        auto fill_masks = [&](auto *dst) {
            typedef typename std::remove_pointer<decltype(dst)>::type T;
            for (uint32_t h = 0; h < n1; h++)
                for (uint32_t y = 0; y < ROWS; y++) dst[(size_t)y * n1 + h] = (T)S.gmask[(size_t)h * ROWS + y];
            for (uint32_t h = 0; h < n2; h++)
                for (uint32_t y1 = 0; y1 < ROWS; y1++)
                    for (uint32_t y2 = 0; y2 < ROWS; y2++)
                        dst[off2 + ((size_t)y1 * ROWS + y2) * n2 + h] = (T)S.gmask2[((size_t)h * ROWS + y1) * ROWS + y2];
            for (uint32_t h = 0; h < n3; h++)
                for (size_t k = 0; k < r3; k++) { dst[off3 + k * n3 + h] = (T)S.gmask3[h * r3 + k]; dst[offp3 + k * n3 + h] = (T)S.pmask3[h * r3 + k]; }
            for (uint32_t h = 0; h < np2; h++)
                for (size_t k = 0; k < r2; k++) dst[offp2 + k * np2 + h] = (T)S.pmask2[h * r2 + k];
            for (uint32_t h = 0; h < n4; h++)
                for (size_t k = 0; k < r4; k++) dst[offp4 + k * n4 + h] = (T)S.pmask4[h * r4 + k];
        };
        if (S.wide) {
            std::vector<uint64_t> bm(Nn), masks(std::max<size_t>(total, 1));
            for (size_t i = 0; i < Nn; i++) bm[i] = S.bm[i] | (S.out_idx[i] != FAC_NONE ? 1ull << 63 : 0ull);
            fill_masks(masks.data());
            const uint64_t *p64;
            if ((st = upload(E, bm, &p64)) != FAC_OK) return fail(st);
            E->d_s_bm = p64;
            if ((st = upload(E, masks, &p64)) != FAC_OK) return fail(st);
            E->d_s_masks = p64;
        } else {
            std::vector<uint32_t> bm(Nn), masks(std::max<size_t>(total, 1));
            for (size_t i = 0; i < Nn; i++) bm[i] = (uint32_t)S.bm[i];
            fill_masks(masks.data());
            const uint32_t *p32;
            if ((st = upload(E, bm, &p32)) != FAC_OK) return fail(st);
            E->d_s_bm = p32;
            if ((st = upload(E, masks, &p32)) != FAC_OK) return fail(st);
            E->d_s_masks = p32;
        }
        if ((st = upload(E, fcsym, &E->d_s_fc)) != FAC_OK) return fail(st);
        if ((st = upload(E, S.out_idx, &E->d_s_out_idx)) != FAC_OK) return fail(st);
        if ((st = upload(E, S.out2, &E->d_s_out2)) != FAC_OK) return fail(st);
        if ((st = upload(E, S.prune_len, &E->d_s_plen)) != FAC_OK) return fail(st);
        if ((st = upload(E, S.prune_low, &E->d_s_plow)) != FAC_OK) return fail(st);
        if ((st = upload(E, S.sub_pen, &E->d_s_subpen)) != FAC_OK) return fail(st);
        if ((st = upload(E, symof, &E->d_s_symof)) != FAC_OK) return fail(st);
        if ((st = upload(E, S.node_lim, &E->d_s_node_lim)) != FAC_OK) return fail(st);
    }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    CK(cudaDeviceGetAttribute(&E->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    E->sm_count = prop.multiProcessorCount;
    E->lookahead = (uint32_t)H.max_match_graphemes + H.max_map_hay + 4;   // + 1: the deep tables of the succinct kernel read the context word of j + 1
    E->default_tile = (uint32_t)env_int("FAC_TILE", 0);
    E->qcap = (uint32_t)env_int("FAC_QCAP", 1 << 17);
    E->smem_tab = (uint32_t)env_int("FAC_SMEM_TAB", 1024);
    E->ctas_per_sm = env_int("FAC_CTAS_PER_SM", 6);
    E->use_tma = env_int("FAC_USE_TMA", 1);
    E->fast_ok = H.mef != 255 && H.beam_width == 0 && !H.has_auto_beam && env_int("FAC_FAITHFUL", 0) == 0;
    E->succ_ok = E->fast_ok && H.succ.ok && env_int("FAC_SUCCINCT", 1) != 0;
    // engines on the reference's generic path (no limits at all, or per-pattern / per-type limits): order-independent
    // only through the succinct kernel (exact-chain shortcut / limits mode)
    E->succ_generic_ok = H.succ.ok && (H.succ.exact_only || H.succ.limits_mode) && H.beam_width == 0 && !H.has_auto_beam && env_int("FAC_FAITHFUL", 0) == 0 &&
                    env_int("FAC_SUCCINCT", 1) != 0;
    if (H.flat_ok) {
        if ((st = upload(E, H.flat_nrec, &E->d_flat_nrec)) != FAC_OK) return fail(st);
        if ((st = upload(E, H.flat_erec, &E->d_flat_erec)) != FAC_OK) return fail(st);
        if ((st = upload(E, H.flat_ooff, &E->d_flat_ooff)) != FAC_OK) return fail(st);
        if ((st = upload(E, H.flat_olist, &E->d_flat_olist)) != FAC_OK) return fail(st);
        if ((st = upload(E, H.flat_gm_row, &E->d_flat_gm_row)) != FAC_OK) return fail(st);
        if ((st = upload(E, H.flat_gm, &E->d_flat_gm)) != FAC_OK) return fail(st);
        if ((st = upload(E, H.flat_pm, &E->d_flat_pm)) != FAC_OK) return fail(st);
        if (!H.flat_px_row.empty()) {
            if ((st = upload(E, H.flat_px_row, &E->d_flat_px_row)) != FAC_OK) return fail(st);
            if ((st = upload(E, H.flat_px, &E->d_flat_px)) != FAC_OK) return fail(st);
        }
    }
    E->stack_ok = E->fast_ok && H.flat_ok && env_int("FAC_STACK", 1) != 0;
    {
        uint32_t maxdeg = 0, maxmaps = 0;
        for (uint32_t i = 0; i < H.n_nodes(); i++) {
            maxdeg = std::max(maxdeg, H.node_edge_off[i + 1] - H.node_edge_off[i]);
            maxmaps = std::max(maxmaps, H.node_map_off[i + 1] - H.node_map_off[i]);
        }
        E->max_fan = 2 * maxdeg + maxmaps + 3;
    }
    E->beam2_warps = (uint32_t)std::min(8, std::max(1, env_int("FAC_BEAM2_WARPS", 5)));
    E->beam2_vcap = env_int("FAC_BEAM2_VCAP", 1024) == 512 ? 512u : 1024u;
    E->beam2_ctas = (uint32_t)std::min(8, std::max(1, env_int("FAC_BEAM2_CTAS", 8)));
    if (env_int("FAC_STACK_CAP", 0) == 0) E->stack_cap = std::max<uint32_t>(256u, (E->max_fan + 40u + 31u) & ~31u);   // the widest state must fit on top of a fed stack
    if (E->stack_cap > 1400u) E->stack_ok = false;   // such a fan-out does not fit the per-warp shared-memory stacks: generic kernel
    E->beam2_ok = H.flat_ok && H.mef != 255 && (H.beam_width != 0 || H.has_auto_beam) && env_int("FAC_BEAM2", 1) != 0;
    E->stack_cap = (uint32_t)std::min(2048, std::max(64, env_int("FAC_STACK_CAP", 256)));
    E->stack_tile = (uint32_t)std::min(4096, std::max(32, env_int("FAC_STACK_TILE", 1024)));
    E->succ_nt = (uint32_t)env_int("FAC_SUCC_THREADS", 1024);
    E->succ_tile = (uint32_t)std::min(4096, std::max(32, env_int("FAC_SUCC_TILE", 4096)));  // 12-bit window field of a state
    E->succ_stack = (uint32_t)env_int("FAC_SUCC_STACK", 0);
    E->radix_unsorted = env_int("FAC_RADIX_UNSORTED", 1) != 0;
    // start windows fed at once: engines whose roots' children are on their last edit push few states per root (the
    // productivity masks leave ~20 of 55), deeper budgets fan out again
    E->succ_feed = (uint32_t)std::min(32, std::max(1, env_int("FAC_SUCC_FEED", (H.mef <= 2 && !H.succ.limits_mode) ? 8 : 1)));
    if (H.succ.ok) {
        // a popped state that is not on its last edit reserves 2 * children + 3 stack slots and roots are fed while
        // fewer than 32 states are stacked: the stack must hold the widest node on top of those 32
        uint32_t maxdeg = 0;
        for (uint64_t b : H.succ.bm) maxdeg = std::max<uint32_t>(maxdeg, (uint32_t)__builtin_popcountll(b & 0x7FFFFFFFFFFFFFFFull));
        E->succ_min_stack = (2 * maxdeg + 3 + 32 + 7) & ~7u;
    }
    if (E->succ_nt != 512 && E->succ_nt != 768 && E->succ_nt != 1024) {
        set_err("FAC_SUCC_THREADS must be 512, 768 or 1024"); fac_engine_free(E); return FAC_INVALID_ARGUMENT;
    }
    // memory safety needs room for the 32-root feed on top of < 32 stacked states; below succ_min_stack results stay
    // exact but windows whose widest state does not fit are redone by the order-faithful kernel (tests use that)
    if (E->succ_stack != 0 && (E->succ_stack < 40 || E->succ_stack > 4096)) {
        set_err("FAC_SUCC_STACK must be 0 (automatic) or between 40 and 4096");
        fac_engine_free(E); return FAC_INVALID_ARGUMENT;
    }
    *out = E;
    return FAC_OK;
}

fac_status fac_engine_create(const fac_config *cfg, const fac_pattern *patterns, size_t n_patterns, fac_engine **out) {
    return fac_engine_create_on(-1, cfg, patterns, n_patterns, out);
}

fac_status fac_engine_create_multi(const int *devices, size_t n_devices, const fac_config *cfg, const fac_pattern *patterns, size_t n_patterns,
                                   fac_engine **out) {
    if (!out) { set_err("null out pointer"); return FAC_INVALID_ARGUMENT; }
    *out = nullptr;
    if (!devices || n_devices == 0) { set_err("fac_engine_create_multi needs at least one device"); return FAC_INVALID_ARGUMENT; }
    for (size_t i = 0; i < n_devices; i++)
        for (size_t j = 0; j < i; j++)
            if (devices[i] == devices[j]) { set_err("fac_engine_create_multi: duplicate device index"); return FAC_INVALID_ARGUMENT; }
    fac_engine *primary = nullptr;
    CKS(fac_engine_create_on(devices[0], cfg, patterns, n_patterns, &primary));
    for (size_t i = 1; i < n_devices; i++) {
        fac_engine *r = nullptr;
        const fac_status st = fac_engine_create_on(devices[i], cfg, patterns, n_patterns, &r);
        if (st != FAC_OK) { fac_engine_free(primary); return st; }
        primary->replicas.push_back(r);
    }
    cudaSetDevice(primary->device);
    *out = primary;
    return FAC_OK;
}

void fac_engine_free(fac_engine *E) {
    if (!E) return;
    for (fac_engine *r : E->replicas) fac_engine_free(r);
    E->replicas.clear();
    cudaSetDevice(E->device);
    for (Workspace *w : E->pool) { w->destroy(); delete w; }
    for (void *p : E->dallocs) cudaFree(p);
    delete E;
}

size_t fac_engine_max_match_graphemes(const fac_engine *E) { return E ? E->host.max_match_graphemes : 0; }
int fac_engine_prefilter_active(const fac_engine *E) { return E && E->host.bitap.active ? 1 : 0; }
size_t fac_engine_num_nodes(const fac_engine *E) { return E ? E->host.n_nodes() : 0; }
size_t fac_engine_num_patterns(const fac_engine *E) { return E ? E->host.patterns.size() : 0; }
int fac_engine_device(const fac_engine *E) { return E ? E->device : -1; }
size_t fac_engine_num_devices(const fac_engine *E) { return E ? 1 + E->replicas.size() : 0; }

static fac_status search_common(const fac_engine *E, const uint8_t *hay, size_t len, bool on_device, float thr, int order, int overlap,
                                int use_prefilter, size_t own_begin, size_t own_end, uint64_t base, bool apply, fac_matches **out,
                                uint32_t flags = 0) {
    if (!E || !out || (len && !hay)) { set_err("null argument"); return FAC_INVALID_ARGUMENT; }
    *out = nullptr;
    if (own_end > len) own_end = len;
    if (own_begin > own_end) { set_err("own_begin lies behind own_end"); return FAC_INVALID_ARGUMENT; }
    const bool partial = own_begin > 0 || own_end < len;
    if (partial && E->host.has_auto_beam) {
        // the auto_beam budget accumulates over ALL start windows of the haystack (src/search.rs:1096-1103): a shard
        // cannot know where the running total stands, so a sharded call would switch to the beam at other windows
        set_err("auto_beam engines cannot search a shard of a haystack: the state budget is cumulative over the whole call");
        return FAC_UNSUPPORTED;
    }
    if (partial && overlap != FAC_OVERLAP_KEEP) {
        set_err("overlap resolution is global: gather the shard lists and call fac_matches_apply[_device]");
        return FAC_INVALID_ARGUMENT;
    }
    CK(cudaSetDevice(E->device));
    fac_status st;
    Workspace *ws = acquire_ws(E, st);
    if (!ws) return st;
    fac_matches *M = new fac_matches();
    M->device = E->device;
    auto body = [&]() -> fac_status {
        CK(cudaEventRecord(ws->ev_begin, ws->stream));
        const uint8_t *d_text = hay;
        if (!on_device && len) {
            CKS(ws->hay.ensure(len + 64));
            CK(cudaMemcpyAsync(ws->hay.p, hay, len, cudaMemcpyHostToDevice, ws->stream));
            CK(cudaMemsetAsync((uint8_t *)ws->hay.p + len, 0, 64, ws->stream));
            d_text = ws->hay.as<uint8_t>();
        }
        CKS(search_resident(E, ws, d_text, len, thr, order, overlap, own_begin, own_end, base, ~0ull, apply, M->v, M->stats, use_prefilter != 0,
                            (flags & FAC_RESULT_ON_DEVICE) ? M : nullptr, (flags & FAC_TEXT_IS_UNICODE) != 0));
        if (flags & FAC_RESULT_ON_DEVICE) M->on_device = true;
        CK(cudaEventRecord(ws->ev_end, ws->stream));
        CK(cudaEventSynchronize(ws->ev_end));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ws->ev_begin, ws->ev_end));
        M->stats.device_ms = ms;
        return FAC_OK;
    };
    st = body();
    release_ws(E, ws);
    if (st != FAC_OK) { delete M; return st; }
    *out = M;
    return FAC_OK;
}

fac_status fac_search(const fac_engine *E, const uint8_t *haystack, size_t len, float threshold, fac_order order, fac_overlap overlap,
                      int use_prefilter, fac_matches **out) {
    return search_common(E, haystack, len, false, threshold, order, overlap, use_prefilter, 0, len, 0, true, out);
}
fac_status fac_search_device(const fac_engine *E, const uint8_t *d_haystack, size_t len, float threshold, fac_order order,
                             fac_overlap overlap, int use_prefilter, fac_matches **out) {
    return search_common(E, d_haystack, len, true, threshold, order, overlap, use_prefilter, 0, len, 0, true, out);
}
fac_status fac_search_shard(const fac_engine *E, const uint8_t *haystack, size_t len, size_t own_begin, size_t own_end, uint64_t base,
                            float threshold, int on_device, fac_matches **out) {
    return search_common(E, haystack, len, on_device != 0, threshold, FAC_ORDER_UNSORTED, FAC_OVERLAP_KEEP, 0, own_begin, own_end, base, true, out);
}
fac_status fac_search_ex(const fac_engine *E, const fac_search_args *a, fac_matches **out) {
    if (!a) { set_err("null argument"); return FAC_INVALID_ARGUMENT; }
    if ((unsigned)a->order > FAC_ORDER_COVERAGE_WEIGHTED || (unsigned)a->overlap > FAC_OVERLAP_NON_OVERLAPPING_UNIQUE) {
        set_err("invalid order / overlap"); return FAC_INVALID_ARGUMENT;
    }
    return search_common(E, a->haystack, a->len, (a->flags & FAC_HAYSTACK_ON_DEVICE) != 0, a->threshold, a->order, a->overlap, a->use_prefilter,
                         a->own_begin, a->own_end, a->base, true, out, a->flags);
}

// Shard plan: cuts on extended-grapheme-cluster boundaries of the WHOLE haystack (a cut inside a cluster would give
// the next shard a start window the whole-haystack search never has), halo counted in clusters.
fac_status fac_plan_shards(size_t max_match_graphemes, const uint8_t *hay, size_t len, size_t n_shards, fac_shard *out) {
    if (!out || n_shards == 0) { set_err("null argument"); return FAC_INVALID_ARGUMENT; }
    const size_t halo = max_match_graphemes + 3;
    const bool ascii = hay == nullptr || host_is_ascii(hay, len);
    const UnicodeTables &U = fac::host_unicode_tables();
    auto snap = [&](size_t c) -> size_t {  // first cluster boundary at or after c
        if (ascii || c >= len) return std::min(c, len);
        while (c < len && (fac_is_cont(hay[c]) || !fac_break_before(U, hay, 0, c))) c++;
        return c;
    };
    auto advance = [&](size_t c, size_t clusters) -> size_t {  // `clusters` boundaries further right
        if (ascii) return std::min(len, c + clusters);
        for (size_t k = 0; k < clusters && c < len; k++) c = snap(c + 1);
        return c;
    };
    size_t prev = 0;
    for (size_t r = 0; r < n_shards; r++) {
        size_t end = r + 1 == n_shards ? len : snap((size_t)(((unsigned __int128)len * (r + 1)) / n_shards));
        if (end < prev) end = prev;
        out[r].own_begin = prev; out[r].own_end = end; out[r].read_end = advance(end, halo);
        prev = end;
    }
    return FAC_OK;
}

fac_status fac_matches_apply(const fac_engine *E, const fac_match *in, size_t n, fac_order order, fac_overlap overlap, fac_matches **out) {
    if (!E || !out || (n && !in)) { set_err("null argument"); return FAC_INVALID_ARGUMENT; }
    *out = nullptr;
    CK(cudaSetDevice(E->device));
    fac_status st;
    Workspace *ws = acquire_ws(E, st);
    if (!ws) return st;
    fac_matches *M = new fac_matches();
    auto body = [&]() -> fac_status {
        if (n == 0) return FAC_OK;
        std::vector<WMatch> w(n);
        for (size_t i = 0; i < n; i++) {
            w[i].start = in[i].start; w[i].end = in[i].end; w[i].pat = in[i].pattern_index; w[i].sim = in[i].similarity; w[i].win = 0;
            w[i].cnt = in[i].insertions | (in[i].deletions << 8) | (in[i].substitutions << 16) | ((uint32_t)in[i].swaps << 24);
            if (in[i].pattern_index >= E->host.patterns.size()) { set_err("pattern index out of range"); return FAC_INVALID_ARGUMENT; }
        }
        CKS(ws->m_a.ensure(n * sizeof(WMatch)));
        CK(cudaMemcpyAsync(ws->m_a.p, w.data(), n * sizeof(WMatch), cudaMemcpyHostToDevice, ws->stream));
        FacWindow fw;
        memset(&fw, 0, sizeof(fw));
        fw.commit = ~0ull;
        CKS(ws->windows.ensure(sizeof(FacWindow)));
        CK(cudaMemcpyAsync(ws->windows.p, &fw, sizeof(fw), cudaMemcpyHostToDevice, ws->stream));
        CK(cudaStreamSynchronize(ws->stream));
        uint32_t kept = 0;
        CKS(apply_device(E, ws, (uint32_t)n, order, overlap, 1, &kept, M->stats));
        CKS(finalize_and_fetch(ws, kept, ws->windows.as<FacWindow>(), false, M->v, M->stats));
        return FAC_OK;
    };
    st = body();
    release_ws(E, ws);
    if (st != FAC_OK) { delete M; return st; }
    *out = M;
    return FAC_OK;
}

fac_status fac_matches_apply_device(const fac_engine *E, const fac_match *d_in, size_t n, fac_order order, fac_overlap overlap, uint32_t flags,
                                    fac_matches **out) {
    if (!E || !out || (n && !d_in)) { set_err("null argument"); return FAC_INVALID_ARGUMENT; }
    *out = nullptr;
    if (n > 0xFFFFFFF0ull) { set_err("match list exceeds the 32-bit rank space"); return FAC_UNSUPPORTED; }
    CK(cudaSetDevice(E->device));
    fac_status st;
    Workspace *ws = acquire_ws(E, st);
    if (!ws) return st;
    fac_matches *M = new fac_matches();
    M->device = E->device;
    const bool to_device = (flags & FAC_RESULT_ON_DEVICE) != 0;
    auto body = [&]() -> fac_status {
        cudaStream_t s = ws->stream;
        CK(cudaEventRecord(ws->ev_begin, s));
        if (to_device) M->on_device = true;
        if (n) {
            if ((flags & FAC_APPLY_PRESORTED) && overlap == FAC_OVERLAP_KEEP) {
                // nothing to rank or select: the answer is the input list
                if (to_device) {
                    const size_t bytes = n * sizeof(fac_match);
                    if (!device_pool().get(M->device, bytes, M->d_out)) CKS(M->d_out.ensure(bytes));
                    CK(cudaMemcpyAsync(M->d_out.p, d_in, bytes, cudaMemcpyDeviceToDevice, s));
                    M->n_dev = n;
                } else {
                    M->v.resize(n);
                    CK(cudaMemcpyAsync(M->v.data(), d_in, n * sizeof(fac_match), cudaMemcpyDeviceToHost, s));
                }
            } else {
                CKS(ws->m_a.ensure(n * sizeof(WMatch)));
                CKS(ws->misc.ensure(64));
                CK(cudaMemsetAsync(ws->misc.p, 0, 4, s));
                k_unfinalize<<<cdiv(n, 256), 256, 0, s>>>(d_in, (uint32_t)n, (uint32_t)E->host.patterns.size(), ws->m_a.as<WMatch>(), ws->misc.as<uint32_t>());
                CK(cudaGetLastError());
                FacWindow fw;
                memset(&fw, 0, sizeof(fw));
                fw.commit = ~0ull;
                CKS(ws->windows.ensure(sizeof(FacWindow)));
                CK(cudaMemcpyAsync(ws->windows.p, &fw, sizeof(fw), cudaMemcpyHostToDevice, s));
                CK(cudaMemcpyAsync(ws->h_flags, ws->misc.p, 4, cudaMemcpyDeviceToHost, s));
                CK(cudaStreamSynchronize(s));
                if (ws->h_flags[0]) { set_err("pattern index out of range"); return FAC_INVALID_ARGUMENT; }
                uint32_t kept = 0;
                CKS(apply_device(E, ws, (uint32_t)n, order, overlap, 1, &kept, M->stats, (flags & FAC_APPLY_PRESORTED) != 0));
                CKS(finalize_and_fetch(ws, kept, ws->windows.as<FacWindow>(), false, M->v, M->stats, to_device ? M : nullptr));
            }
        }
        CK(cudaEventRecord(ws->ev_end, s));
        CK(cudaEventSynchronize(ws->ev_end));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ws->ev_begin, ws->ev_end));
        M->stats.device_ms = ms;
        return FAC_OK;
    };
    st = body();
    release_ws(E, ws);
    if (st != FAC_OK) { delete M; return st; }
    *out = M;
    return FAC_OK;
}

const fac_match *fac_matches_data(const fac_matches *m) { return m && !m->on_device && !m->v.empty() ? m->v.data() : nullptr; }
const fac_match *fac_matches_device_data(const fac_matches *m) { return m && m->on_device && m->n_dev ? m->d_out.as<fac_match>() : nullptr; }
size_t fac_matches_len(const fac_matches *m) { return m ? (m->on_device ? m->n_dev : m->v.size()) : 0; }
uint64_t fac_matches_states_pushed(const fac_matches *m) { return m ? m->stats.states : 0; }
double fac_matches_device_ms(const fac_matches *m) { return m ? m->stats.device_ms : 0; }
double fac_matches_expand_ms(const fac_matches *m) { return m ? m->stats.expand_ms : 0; }
uint32_t fac_matches_kernel_launches(const fac_matches *m) { return m ? m->stats.launches : 0; }
void fac_matches_free(fac_matches *m) { delete m; }

}  // extern "C"

#include "fac_stream.inc"
