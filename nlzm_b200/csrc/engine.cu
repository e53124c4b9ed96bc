// Engine: owns the HBM-resident input, the work buffers and two result slots; runs the stages
// S (suffix ranks) -> T (BT4 divide and conquer) -> H (HT2/HT3) -> R (RK256) -> M (merge) for a
// range of positions and exposes the C ABI of include/nlzm_mf.h.
#include "../../include/nlzm_mf.h"
#include "platform.cuh"
#ifdef NLZM_EMU
#include "emu_runtime.hpp"
#endif
#include "common.cuh"
#include "prim.cuh"
#include "suffix_rank.cuh"
#include "dc_levels.cuh"
#include "bt_short.cuh"
#include "ht_rows.cuh"
#include "rk256.cuh"
#include "merge_steps.cuh"

#include <atomic>
#include <condition_variable>
#include <map>
#include <mutex>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <vector>
#include <stdio.h>

#ifdef NLZM_EMU
thread_local EmuCta nlzm_emu_cta;
#ifdef NLZM_EMU_STATS
unsigned long long nlzm_stats[16];
extern "C" unsigned long long *nlzm_emu_stats() { return nlzm_stats; }
#endif
#endif

// ---- launch accounting / per-kernel timing ----------------------------------------------------
static std::atomic<unsigned long long> g_launches{0};
struct KernelProf {
    struct Pending { const char *name; cudaEvent_t a, b; };
    std::atomic<bool> on{false};
    std::mutex mu;
    std::map<std::string, std::pair<u64, double>> acc;     // name -> (launches, ms)
};
static KernelProf g_prof;
// launches of one find are issued by one host thread: the begin/end pairing stays inside that thread
static thread_local std::vector<KernelProf::Pending> tl_pending;
static void prof_resolve() {                               // call after the stream is synchronized
    if (tl_pending.empty()) return;
    std::lock_guard<std::mutex> l(g_prof.mu);
    for (auto &p : tl_pending) {
        float ms = 0;
        if (p.b && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            auto &e = g_prof.acc[p.name];
            e.first += 1;
            e.second += ms;
        }
        cudaEventDestroy(p.a);
        if (p.b) cudaEventDestroy(p.b);
    }
    tl_pending.clear();
}
void nlzm_launch_begin(const char *name, cudaStream_t st) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_prof.on.load(std::memory_order_relaxed)) return;
    KernelProf::Pending p;
    p.name = name;
    p.b = nullptr;
    cudaEventCreate(&p.a);
    cudaEventRecord(p.a, st);
    tl_pending.push_back(p);
}
void nlzm_launch_end(cudaStream_t st) {
    if (tl_pending.empty() || tl_pending.back().b) return;
    cudaEventCreate(&tl_pending.back().b);
    cudaEventRecord(tl_pending.back().b, st);
}

// CUDA event that cannot leak on an early return
struct Ev {
    cudaEvent_t e = nullptr;
    Ev() { cudaEventCreate(&e); }
    ~Ev() { if (e) cudaEventDestroy(e); }
    Ev(const Ev &) = delete;
    Ev &operator=(const Ev &) = delete;
    operator cudaEvent_t() const { return e; }
};

static std::string g_create_error;

// Scalars the host needs between stages (counts that size the next launch) are written by a one-thread kernel
// straight into mapped pinned host memory: a cudaMemcpy would queue behind the previous range's multi-gigabyte
// device->host copy on the same copy engine and stall the stages for its whole duration.
#ifndef NLZM_EMU
__global__ void k_words_to_host(const u32 *src, volatile u32 *dst, u32 n) { for (u32 i = 0; i < n; i++) dst[i] = src[i]; }
#endif

// device scalars live in one small buffer (u32 slots)
enum { SC_SUM = 0 /* u64 */, SC_RK_HITS = 8, SC_RK_INTERVALS = 9, SC_RK_VALID = 10, SC_RK_OK = 11 };

#define CK(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) return fail((int)e_, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)
#define CKI(expr)                                                                                  \
    do {                                                                                           \
        int r_ = (expr);                                                                           \
        if (r_ != 0) return r_ < 0 ? r_ : fail(r_, std::string(#expr) + ": " + cudaGetErrorString((cudaError_t)r_)); \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    bool shared = false;      // an IPC handle of this allocation was handed out: other processes may keep it mapped,
                              // so it is never freed before the engine goes (it just keeps cycling through the pool)
    template <class T> T *as() const { return (T *)p; }
};

struct Slot {
    DevBuf d_offsets, d_steps;
    void *h_offsets = nullptr, *h_steps = nullptr;
    size_t h_offsets_bytes = 0, h_steps_bytes = 0;
    u64 begin = 0, end = 0, n_steps = 0;
    std::thread worker;
    bool pending = false;
    int status = 0;
};

// A retained segment: one sorted block of a finished stage T (level array + pointers). A later find
// whose range follows queries it instead of re-ranking and re-merging the window behind its range
// (dc_xmerge_tile); it can also come from another engine / GPU (nlzm_mf_import_segment).
struct SegBufs {
    nlzm_mf *owner = nullptr;
    DevBuf el, ptr;
    ~SegBufs();
};
struct Segment {
    std::shared_ptr<SegBufs> bufs;
    u64 u0 = 0;          // element positions are relative to this absolute offset
    u64 elem_off = 0;    // first element inside bufs->el
    u32 n_elems = 0;     // valid elements (positions past the producing find's range are cut off)
    u64 ptr_pos0 = 0;    // absolute position of bufs->ptr[0]
    u64 pos_b = 0, pos_e = 0;    // absolute positions covered
    bool imported = false;       // copy of another engine's segment: used by the next find, never exported or kept
    const Elem *elems() const { return bufs->el.as<Elem>() + elem_off; }
    Elem *elems_rw() const { return bufs->el.as<Elem>() + elem_off; }
    const PtrEntry *ptrs() const {       // indexable by (position - u0)
        return (const PtrEntry *)((uintptr_t)bufs->ptr.p - (uintptr_t)((ptr_pos0 - u0) * sizeof(PtrEntry)));
    }
};

struct Turn;
struct nlzm_mf {
    Geom g;
    int device = 0;
    u32 mask = NLZM_MF_ALL;
    u64 max_range = 0;
    cudaStream_t st = 0, st_copy = 0;
    DevBuf x;
    bool have_input = false;

    DevBuf k64[2], v32[2], rank, ptr, aux0, aux1, el[2], part;     // stages S/T
    DevBuf tk[2], tv[2], tcount, keep, out_idx;                    // tuples / merge
    DevBuf e_k[2], e_v[2], e_inv;                                  // BT short-length bucket sort (small windows)
    DevBuf ht_snap, ht_tab, ht_gmax, ht_coarse, ht_cfirst, ht_clast, ht_ccount, ht_ps, ht_pl, ht_pr;                 // HT: per-tile last-access tables, PS/PL/PR
    DevBuf hblk, sl_k[2], sl_v[2], sl_cnt, sl_off;                 // RK table
    DevBuf hit_k[2], hit_v[2], hit_len, iv, val_k, val_v;          // RK hits / carry intervals
    DevBuf scalars;                                                // misc device scalars
    u32 *h_words = nullptr, *d_words = nullptr;                    // mapped pinned host memory and its device alias
    PrimTemp tmp;
    DevBuf tmpbuf;
    u32 tuple_cap_mult = 6;
    u64 tuple_cap_extra = 1u << 20;        // options "tuple_cap_mult" / "tuple_cap_extra": candidate tuple capacity = range * mult + extra
    u64 ht_margin = NLZM_HT_MARGIN;        // options (nlzm_mf_set_option): tuning / test knobs
    u32 ht_coarse_log = NLZM_HT_COARSE_LOG;
    bool rk_all_hits = false, rk_overflowed = false;
    u64 rk_restart = 1u << 18;             // option "rk_restart": how far before a range stage R starts looking for hits
    bool retain = true;                    // option "retain": keep the sorted blocks of a find for the next one
    u32 max_segments = 8;                  // option "max_segments": more retained segments than this behind a range => halo mode

    cudaEvent_t import_ev = nullptr;       // last asynchronous segment import on st_copy (nlzm_mf_import_segment, via | 0x100)
    bool imports_pending = false;
    std::vector<Segment> segs;             // retained, ascending by position
    std::vector<Segment> fresh;            // sorted blocks of the find in progress (own universe)
    u64 fresh_u0 = 0;
    bool prepared = false;                 // nlzm_mf_prepare ran stages S/T for [prep_b, prep_e): find continues from there
    u64 prep_b = 0, prep_e = 0;
    std::vector<Segment> prep_fresh;       // ... with these blocks and these candidate tuples (other ranges may be
    DevBuf prep_tk, prep_tv;               //     found in between: the first range of a shard finishes last)
    u32 prep_nt = 0;
    float prep_ms_rank = 0, prep_ms_levels = 0, prep_ms = 0;
    std::map<std::string, void *> ipc_open;  // importer: IPC handles already mapped into this process (closed at destroy)
    std::map<void *, std::string> ipc_made;  // exporter: handle of an allocation (made once)
    std::vector<DevBuf> pool;              // level-array / pointer buffers waiting to be reused
    std::mutex pool_mu;

    Slot slot[2];
    // One call computes at a time, in the order the calls were MADE: a submit takes its turn in the caller's thread,
    // so the set_input / prepare / find that follow it cannot overtake the worker thread it spawned.
    std::mutex turn_mu;
    std::condition_variable turn_cv;
    u64 turn_next = 0, turn_serving = 0;
    u64 take_turn() { std::lock_guard<std::mutex> l(turn_mu); return turn_next++; }
    void wait_turn(u64 t) { std::unique_lock<std::mutex> l(turn_mu); turn_cv.wait(l, [&] { return turn_serving == t; }); }
    void end_turn() { { std::lock_guard<std::mutex> l(turn_mu); ++turn_serving; } turn_cv.notify_all(); }
    std::mutex stats_mu;
    std::string err;
    nlzm_mf_stats stats{};

    int fail(int code, const std::string &msg) {
        err = msg;
        return code;
    }

    int ensure(DevBuf &b, size_t bytes) {
        if (bytes <= b.bytes) return 0;
        if (b.p && b.shared) to_pool(b);
        if (b.p) cudaFree(b.p);
        b.p = nullptr;
        b.bytes = 0;
        size_t want = bytes + (bytes >> 4) + 256;
        cudaError_t e = cudaMalloc(&b.p, want);
        if (e != cudaSuccess) {
            b.p = nullptr;
            return fail(NLZM_MF_E_NOMEM, "cudaMalloc(" + std::to_string(want) + "): " + cudaGetErrorString(e));
        }
        b.bytes = want;
        return 0;
    }
    // big level-array / pointer buffers move in and out of retained segments: take a fitting one from the pool first
    int ensure_pooled(DevBuf &b, size_t bytes) {
        if (bytes <= b.bytes) return 0;
        {
            std::lock_guard<std::mutex> l(pool_mu);
            int best = -1;
            for (size_t i = 0; i < pool.size(); i++)
                if (pool[i].bytes >= bytes && (best < 0 || pool[i].bytes < pool[(size_t)best].bytes)) best = (int)i;
            if (best >= 0) {
                std::swap(b, pool[(size_t)best]);
                if (!pool[(size_t)best].p) pool.erase(pool.begin() + best);
                return 0;
            }
        }
        return ensure(b, bytes);
    }
    void to_pool(DevBuf &b) {
        if (!b.p) return;
        std::lock_guard<std::mutex> l(pool_mu);
        size_t held = 0, n_free = 0;
        for (const DevBuf &q : pool) { held += q.bytes; n_free += q.shared ? 0 : 1; }
        if (!b.shared && (n_free >= 8 || held + b.bytes > (64ull << 30))) {     // keep a few, and the largest of them
            size_t small = pool.size();
            for (size_t i = 0; i < pool.size(); i++)
                if (!pool[i].shared && (small == pool.size() || pool[i].bytes < pool[small].bytes)) small = i;
            if (small < pool.size() && pool[small].bytes < b.bytes) std::swap(pool[small], b);
            ipc_made.erase(b.p);
            cudaFree(b.p);
        } else {
            pool.push_back(b);
        }
        b.p = nullptr;
        b.bytes = 0;
        b.shared = false;
    }
    void release(DevBuf &b) {
        if (b.p) cudaFree(b.p);
        b.p = nullptr;
        b.bytes = 0;
    }

    // n 32-bit words at device address `dev` -> host, through the mapped buffer (stream-ordered, then synchronised)
    int fetch_words(const void *dev, void *host, u32 n) {
#ifndef NLZM_EMU
        nlzm_launch_begin("k_words_to_host", st);
        k_words_to_host<<<1, 1, 0, st>>>((const u32 *)dev, d_words, n);
        nlzm_launch_end(st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return fail((int)e, std::string("fetch_words: ") + cudaGetErrorString(e));
        memcpy(host, h_words, (size_t)n * 4);
#else
        memcpy(host, dev, (size_t)n * 4);
#endif
        return 0;
    }

    int ensure_prim(u64 n) {
        size_t need = prim_temp_bytes(n);
        int r = ensure(tmpbuf, need);
        if (r) return r;
        tmp.ptr = tmpbuf.p;
        tmp.bytes = tmpbuf.bytes;
        return 0;
    }

    TupleSink sink() {
        TupleSink s;
        s.keys = tk[0].as<u64>();
        s.vals = tv[0].as<u32>();
        s.count = tcount.as<u32>();
        s.cap = (u32)(tk[0].bytes / 8 < tv[0].bytes / 4 ? tk[0].bytes / 8 : tv[0].bytes / 4);
        return s;
    }

    bool covered_by_segments(u64 b, std::vector<Segment> &use) const;
    int stage_bt4_own(u64 own_b, u64 own_e, u64 u0);
    int stage_bt4_cross(u64 own_b, u64 own_e, const std::vector<Segment> &behind);
    void retain_fresh(u64 own_e);
    void add_segments(const std::vector<Segment> &v);
    int trim_segments(u64 from);
    int publish_segments(u64 from, nlzm_mf_segment *out, uint32_t cap, uint32_t *n_out);
    DevBuf xstage;                         // export buffer: ONE allocation (one IPC handle, opened once by every importer)
                                           // that holds the published copies of this engine's segments
    int stage_ht(u64 own_b, u64 own_e, const HtCfg &c);
    int stage_rk(u64 own_b, u64 own_e);
    int stage_rk_from(u64 own_b, u64 own_e, u64 rk_e, u64 rk_b, bool need_restart, bool &done);
    int stage_merge(u64 own_b, u64 own_e, Slot &s);
    int compute(u64 b, u64 e, Slot &s);
    int find_impl(u64 b, u64 e, int slot, bool to_host, u64 ticket);
    int prepare_impl(u64 b, u64 e);
};

// scope of one call's turn on the engine
struct Turn {
    nlzm_mf *mf;
    bool held = true;
    static const u64 NONE = ~0ull;
    Turn(nlzm_mf *m, u64 ticket = NONE) : mf(m) { mf->wait_turn(ticket == NONE ? mf->take_turn() : ticket); }
    void release() { if (held) { held = false; mf->end_turn(); } }
    ~Turn() { release(); }
};

SegBufs::~SegBufs() {
    if (owner) { owner->to_pool(el); owner->to_pool(ptr); }
    else { if (el.p) cudaFree(el.p); if (ptr.p) cudaFree(ptr.p); }
}

static inline u32 bits_for(u64 v) {
    u32 b = 0;
    while ((1ull << b) < v) ++b;
    return b;
}

// Retained segments that cover the window behind position b without a gap, nearest first.
bool nlzm_mf::covered_by_segments(u64 b, std::vector<Segment> &use) const {
    use.clear();
    const u64 need = b > (u64)(g.W - 1) ? b - (g.W - 1) : 0;
    u64 at = b;
    while (at > need) {
        const Segment *hit = nullptr;                  // the longest segment that ends here
        for (const Segment &s : segs) if (s.pos_e == at && s.pos_b < at && (!hit || s.pos_b < hit->pos_b)) hit = &s;
        if (!hit) return false;
        use.push_back(*hit);
        if (use.size() > max_segments) return false;
        at = hit->pos_b;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// Stage S + T over the universe [u0, own_e + pad) (+ the 2..3-byte bucket-collision candidates of small windows)
// ------------------------------------------------------------------------------------------------
int nlzm_mf::stage_bt4_own(u64 own_b, u64 own_e, u64 u0) {
    fresh.clear();
    fresh_u0 = u0;
    const u64 u1 = own_e + 1024 < g.flen ? own_e + 1024 : g.flen;
    const u64 n = u1 - u0;
    if (n < 2) return 0;
    CKI(ensure(k64[0], n * 8)); CKI(ensure(k64[1], n * 8));
    CKI(ensure(v32[0], n * 4)); CKI(ensure(v32[1], n * 4));
    CKI(ensure(rank, n * 4)); CKI(ensure(aux0, n * 4)); CKI(ensure(aux1, n * 4));
    CKI(ensure_pooled(ptr, n * sizeof(PtrEntry)));
    CKI(ensure_prim(n));
    u64 *sum_dev = scalars.as<u64>();

    Ev ev0, ev1, ev2;
    cudaEventRecord(ev0, st);

    // --- S: prefix doubling 8 -> 16 -> ... -> >= 264 bytes; rounds after the first only re-sort the
    //        positions whose group is still ambiguous. Scratch comes out of the (not yet used) level arrays.
    CKI(ensure_pooled(el[0], n * sizeof(Elem))); CKI(ensure_pooled(el[1], n * sizeof(Elem)));
    u32 *scratch = el[1].as<u32>();
    u32 *new_grp = scratch, *act_flag = scratch + n, *act_idx = scratch + 2 * n;
    u32 *grp_buf[2] = {scratch + 3 * n, scratch + 4 * n}, *pos_buf[2] = {scratch + 5 * n, scratch + 6 * n};
    const u32 nb = bits_for(n + 2);
    {
        RankInitParams ip{x.as<u8>(), u0, k64[0].as<u64>(), v32[0].as<u32>()};
        launch_rank_init(ip, n, st);
    }
    u64 m = n, depth = 8;
    int ab = 0;
    bool round0 = true;
    while (true) {
        int sel = 0;
        CKI(prim_sort_pairs64(tmp, k64[0].as<u64>(), k64[1].as<u64>(), v32[0].as<u32>(), v32[1].as<u32>(), m, 0,
                              round0 ? 64 : (int)(2 * nb), st, &sel));
        RankHeadParams hp{k64[sel].as<u64>(), aux0.as<u32>(), aux1.as<u32>(), nb, round0 ? 1u : 0u};
        launch_rank_head(hp, m, st);
        CKI(prim_inclusive_max(tmp, aux0.as<u32>(), aux0.as<u32>(), m, st));
        CKI(prim_inclusive_max(tmp, aux1.as<u32>(), aux1.as<u32>(), m, st));
        RankAssignParams ap{k64[sel].as<u64>(), v32[sel].as<u32>(), aux0.as<u32>(), aux1.as<u32>(), rank.as<u32>(),
                            new_grp, act_flag, m, nb, round0 ? 1u : 0u};
        launch_rank_assign(ap, m, st);
        depth *= 2;                                       // ranks now order the first `depth` bytes
        if (depth / 2 >= NLZM_MATCH_MAX) break;
        CKI(prim_sum(tmp, act_flag, sum_dev, m, st));
        u64 m_next = 0;
        CKI(fetch_words(sum_dev, &m_next, 2));
        if (m_next == 0) break;
        CKI(prim_exclusive_sum(tmp, act_flag, act_idx, m, st));
        RankCompactParams cp{act_flag, act_idx, new_grp, v32[sel].as<u32>(), grp_buf[ab], pos_buf[ab]};
        launch_rank_compact(cp, m, st);
        m = m_next;
        RankKeysParams kp{grp_buf[ab], pos_buf[ab], rank.as<u32>(), k64[0].as<u64>(), v32[0].as<u32>(), n, depth / 2, nb};
        launch_rank_keys(kp, m, st);
        ab ^= 1;
        round0 = false;
    }
    cudaEventRecord(ev1, st);

    // --- T: first NLZM_BASE_LOG levels in shared memory, then one merge-path pass per level
    const u64 tiles = (n + NLZM_MT_TILE - 1) / NLZM_MT_TILE;
    CKI(ensure(part, (tiles + 2) * 4));
    DcParams dp{};
    dp.x = x.as<u8>();
    dp.g = g;
    dp.u0 = u0;
    dp.own_b = own_b;
    dp.own_e = own_e;
    dp.n = (u32)n;
    dp.n_valid = (u32)(own_e - u0);
    dp.h = 0;
    dp.rank = rank.as<u32>();
    dp.cur = nullptr;
    dp.nxt = el[0].as<Elem>();
    dp.corank = nullptr;
    dp.part = part.as<u32>();
    dp.ptr = ptr.as<PtrEntry>();
    dp.sink = sink();
    CKI(launch_dc_base(dp, (n + NLZM_BASE_TILE - 1) / NLZM_BASE_TILE, NLZM_BASE_SMEM, st));
    dp.corank = rank.as<u32>();          // ranks now live inside the element keys
    dp.origin = 0;
    dp.cross = 0;
    int cur = 0;
    // binary levels only up to window-sized segments: nothing further back than W-1 can be a candidate
    const u64 h_end = n < (u64)g.W ? n : (u64)g.W;
    for (u64 h = NLZM_BASE_TILE; h < h_end; h <<= 1) {
        dp.cur = el[cur].as<Elem>();
        dp.nxt = el[cur ^ 1].as<Elem>();
        dp.h = (u32)h;
        launch_dc_partition(dp, tiles, st);
        CKI(launch_dc_merge_tile(dp, tiles, NLZM_MT_SMEM, st));
        launch_dc_link(dp, n, st);
        cur ^= 1;
    }
    // ... then every window-sized block queries the block before it (two passes: even and odd pairs);
    // these passes store nothing but the best lengths: no level above them exists
    if (n > (u64)g.W && (u64)g.W >= NLZM_BASE_TILE) {
        for (u32 parity = 0; parity < 2; parity++) {
            const u64 origin = (u64)parity * g.W;
            if (n <= origin + g.W) break;
            DcParams cp = dp;
            cp.origin = (u32)origin;
            cp.cross = 1;
            cp.n = (u32)(n - origin);
            cp.h = g.W;
            cp.cur = el[cur].as<Elem>() + origin;
            cp.nxt = nullptr;
            cp.corank = nullptr;
            const u64 ctiles = (cp.n + NLZM_MT_TILE - 1) / NLZM_MT_TILE;
            launch_dc_partition(cp, ctiles, st);
            CKI(launch_dc_merge_tile(cp, ctiles, NLZM_MT_SMEM, st));
        }
    }

    // --- lengths 2..3 inside a bucket (only windows < 2^19)
    if (g.bt_bits < 16) {
        const u64 pe = own_e < g.flen - 3 ? own_e : g.flen - 3;      // callers guarantee flen >= 4 here
        if (pe > own_b) {
            const u64 s0 = own_b > 4095 ? own_b - 4095 : 0;
            const u64 m = pe - s0;
            CKI(ensure(e_k[0], m * 4)); CKI(ensure(e_k[1], m * 4));
            CKI(ensure(e_v[0], m * 4)); CKI(ensure(e_v[1], m * 4));
            CKI(ensure(e_inv, m * 4));
            CKI(ensure_prim(m));
            BtBucketParams bp{x.as<u8>(), s0, 32 - g.bt_bits, e_k[0].as<u32>(), e_v[0].as<u32>()};
            launch_bt_bucket(bp, m, st);
            int sel = 0;
            CKI(prim_sort_pairs32(tmp, e_k[0].as<u32>(), e_k[1].as<u32>(), e_v[0].as<u32>(), e_v[1].as<u32>(), m, 0, (int)g.bt_bits, st, &sel));
            HtInvParams vp{e_v[sel].as<u32>(), e_inv.as<u32>(), 0u};
            launch_ht_inv(vp, m, st);
            BtShortParams sp{x.as<u8>(), g, e_k[sel].as<u32>(), e_v[sel].as<u32>(), e_inv.as<u32>(), s0, own_b, sink()};
            launch_bt_short(sp, pe - own_b, st);
        }
    }
    cudaEventRecord(ev2, st);
    CK(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&stats.ms_rank, ev0, ev1);
    cudaEventElapsedTime(&stats.ms_levels, ev1, ev2);

    // The final level array (sorted window-sized blocks, or one sorted array) and the pointers become segments.
    // Positions past own_e carry the pad rank and sit at the end of their block: they are cut off.
    auto bufs = std::make_shared<SegBufs>();
    bufs->owner = this;
    std::swap(bufs->el, el[cur]);
    std::swap(bufs->ptr, ptr);
    const u64 n_valid = own_e - u0;
    const u64 blk = n > (u64)g.W ? (u64)g.W : n;
    for (u64 o = 0; o < n_valid; o += blk) {
        Segment sg;
        sg.bufs = bufs;
        sg.u0 = u0;
        sg.elem_off = o;
        sg.n_elems = (u32)((n_valid - o) < blk ? (n_valid - o) : blk);
        sg.ptr_pos0 = u0;
        sg.pos_b = u0 + o;
        sg.pos_e = sg.pos_b + sg.n_elems;
        fresh.push_back(sg);
    }
    return 0;
}

// The first block of the own universe queries the retained segments behind it, nearest first (a longer
// match further back only counts if nothing nearer is at least as long: best lengths carry over).
int nlzm_mf::stage_bt4_cross(u64 own_b, u64 own_e, const std::vector<Segment> &behind) {
    stats.segments_queried = 0;
    if (fresh.empty() || behind.empty()) return 0;
    const Segment &own = fresh.front();
    Ev ev0, ev1;
    cudaEventRecord(ev0, st);
    for (const Segment &sg : behind) {
        if (sg.n_elems == 0) continue;
        const u64 total = (u64)sg.n_elems + own.n_elems;
        const u64 tiles = (total + NLZM_MT_TILE - 1) / NLZM_MT_TILE;
        CKI(ensure(part, (tiles + 2) * 4));
        DcParams cp{};
        cp.x = x.as<u8>();
        cp.g = g;
        cp.u0 = own.u0;
        cp.own_b = own_b;
        cp.own_e = own_e;
        cp.part = part.as<u32>();
        cp.sink = sink();
        cp.cross = 2;
        cp.seg = sg.elems();
        cp.seg_ptr = sg.ptrs();
        cp.seg_u0 = sg.u0;
        cp.seg_last = sg.pos_e - 1;
        cp.seg_len = sg.n_elems;
        cp.own = own.elems_rw();
        cp.own_len = own.n_elems;
        stats.segments_queried += 1;
        launch_dc_xpartition(cp, tiles, st);
        CKI(launch_dc_xmerge_tile(cp, tiles, NLZM_MT_SMEM, st));
    }
    cudaEventRecord(ev1, st);
    CK(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&stats.ms_cross, ev0, ev1);
    return 0;
}

// new blocks replace whatever the list holds for the same positions
void nlzm_mf::add_segments(const std::vector<Segment> &v) {
    if (v.empty()) return;
    const u64 lo = v.front().pos_b, hi = v.back().pos_e;
    std::vector<Segment> kept;
    for (Segment &o : segs) if (o.pos_e <= lo || o.pos_b >= hi) kept.push_back(o);
    for (const Segment &sg : v) kept.push_back(sg);
    segs.swap(kept);
}

// Only positions >= from stay in the retained list. A segment that straddles `from` is replaced by the sub-sequence of
// its elements at those positions: greater-position pointers never lead to an earlier position, so the later part of
// a segment is a complete segment of its own (what a neighbour imports is then at most one window, however the
// exporting shard cut its blocks).
int nlzm_mf::trim_segments(u64 from) {
    std::vector<Segment> kept;
    for (Segment &sg : segs) {
        if (sg.pos_e <= from) continue;
        if (sg.pos_b >= from) { kept.push_back(sg); continue; }
        const u64 n = sg.n_elems;
        CKI(ensure(aux0, n * 4)); CKI(ensure(aux1, n * 4));
        CKI(ensure_prim(n));
        auto bufs = std::make_shared<SegBufs>();
        bufs->owner = this;
        const u64 n_keep = sg.pos_e - from;
        CKI(ensure_pooled(bufs->el, n_keep * sizeof(Elem)));
        CKI(ensure_pooled(bufs->ptr, n_keep * sizeof(PtrEntry)));
        SegTrimParams tp{sg.elems(), bufs->el.as<Elem>(), aux0.as<u32>(), aux1.as<u32>(), (u32)(from - sg.u0)};
        launch_seg_trim_flag(tp, n, st);
        CKI(prim_exclusive_sum(tmp, aux0.as<u32>(), aux1.as<u32>(), n, st));
        launch_seg_trim_move(tp, n, st);
        CK(cudaMemcpyAsync(bufs->ptr.p, (const u8 *)sg.bufs->ptr.p + (from - sg.ptr_pos0) * sizeof(PtrEntry), n_keep * sizeof(PtrEntry),
                           cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
        Segment t;
        t.bufs = bufs;
        t.u0 = sg.u0;
        t.elem_off = 0;
        t.n_elems = (u32)n_keep;
        t.ptr_pos0 = from;
        t.pos_b = from;
        t.pos_e = sg.pos_e;
        kept.push_back(t);
    }
    segs.swap(kept);
    stats.segments_retained = (u32)segs.size();
    return 0;
}

// Copies of this engine's own segments, cut down to positions >= from, into the export buffer, and their
// descriptors. Importers in other processes map that one allocation once; what changes from round to round is
// only its content. (Handing out the level-array buffers themselves would mean a new IPC mapping on the importing
// side whenever the buffers rotate: hundreds of milliseconds for allocations of several gigabytes.)
int nlzm_mf::publish_segments(u64 from, nlzm_mf_segment *out, uint32_t cap, uint32_t *n_out) {
    struct Plan { const Segment *sg; u64 first, n_keep, el_off, pt_off; };
    std::vector<Plan> plan;
    u64 bytes = 0;
    for (const Segment &sg : segs) {
        if (sg.imported || sg.pos_e <= from) continue;
        Plan pl;
        pl.sg = &sg;
        pl.first = sg.pos_b > from ? sg.pos_b : from;
        pl.n_keep = sg.pos_e - pl.first;                       // one element per position
        pl.el_off = bytes; bytes += (pl.n_keep * sizeof(Elem) + 255) & ~255ull;
        pl.pt_off = bytes; bytes += (pl.n_keep * sizeof(PtrEntry) + 255) & ~255ull;
        plan.push_back(pl);
    }
    *n_out = (uint32_t)plan.size();
    if (!out || cap < plan.size()) return 0;
    if (bytes > xstage.bytes) { ipc_made.erase(xstage.p); CKI(ensure(xstage, bytes + (bytes >> 3))); }
    u8 *base = xstage.as<u8>();
    for (const Plan &pl : plan) {
        const Segment &sg = *pl.sg;
        if (pl.first == sg.pos_b) {
            CK(cudaMemcpyAsync(base + pl.el_off, sg.elems(), pl.n_keep * sizeof(Elem), cudaMemcpyDeviceToDevice, st));
        } else {
            const u64 n = sg.n_elems;
            CKI(ensure(aux0, n * 4)); CKI(ensure(aux1, n * 4));
            CKI(ensure_prim(n));
            SegTrimParams tp{sg.elems(), (Elem *)(base + pl.el_off), aux0.as<u32>(), aux1.as<u32>(), (u32)(pl.first - sg.u0)};
            launch_seg_trim_flag(tp, n, st);
            CKI(prim_exclusive_sum(tmp, aux0.as<u32>(), aux1.as<u32>(), n, st));
            launch_seg_trim_move(tp, n, st);
        }
        CK(cudaMemcpyAsync(base + pl.pt_off, (const u8 *)sg.bufs->ptr.p + (pl.first - sg.ptr_pos0) * sizeof(PtrEntry),
                           pl.n_keep * sizeof(PtrEntry), cudaMemcpyDeviceToDevice, st));
    }
    CK(cudaStreamSynchronize(st));
    std::string handle;
#ifndef NLZM_EMU
    {
        auto it = ipc_made.find(xstage.p);
        if (it == ipc_made.end()) {
            cudaIpcMemHandle_t h;
            if (cudaIpcGetMemHandle(&h, xstage.p) == cudaSuccess) it = ipc_made.emplace(xstage.p, std::string((const char *)&h, sizeof h)).first;
            else cudaGetLastError();
        }
        if (it != ipc_made.end()) handle = it->second;
    }
#endif
    for (size_t i = 0; i < plan.size(); i++) {
        const Plan &pl = plan[i];
        nlzm_mf_segment &d = out[i];
        memset(&d, 0, sizeof d);
        d.pos_begin = pl.first;
        d.pos_end = pl.sg->pos_e;
        d.origin = pl.sg->u0;
        d.n_elems = pl.n_keep;
        d.elems_offset_bytes = pl.el_off;
        d.elems_bytes = pl.n_keep * sizeof(Elem);
        d.ptrs_offset_bytes = pl.pt_off;
        d.ptrs_bytes = pl.n_keep * sizeof(PtrEntry);
        d.elems_alloc = xstage.p;
        d.ptrs_alloc = xstage.p;
        d.device = device;
        if (!handle.empty()) { memcpy(d.ipc_elems, handle.data(), handle.size()); memcpy(d.ipc_ptrs, handle.data(), handle.size()); d.flags = 3u; }
    }
    return 0;
}

// after a find: its blocks join the retained list; whatever a range starting at own_e could not reach goes
void nlzm_mf::retain_fresh(u64 own_e) {
    if (!retain) { fresh.clear(); segs.clear(); stats.segments_retained = 0; return; }
    add_segments(fresh);
    fresh.clear();
    const u64 keep_from = own_e > (u64)(g.W - 1) ? own_e - (g.W - 1) : 0;
    std::vector<Segment> kept;
    for (Segment &sg : segs)
        if (!sg.imported && sg.pos_e > keep_from && (sg.pos_e <= own_e || (prepared && sg.pos_b >= prep_b && sg.pos_e <= prep_e))) kept.push_back(sg);
    segs.swap(kept);
    stats.segments_retained = (u32)segs.size();
}

// ------------------------------------------------------------------------------------------------
// Stage H
// ------------------------------------------------------------------------------------------------
int nlzm_mf::stage_ht(u64 own_b, u64 own_e, const HtCfg &c) {
    if (g.flen < 4) return 0;
    const u64 n_acc = own_e < g.flen - 3 ? own_e : g.flen - 3;       // accesses happen while 4 bytes are visible
    if (n_acc <= own_b) return 0;
    const u64 nc = 1ull << c.bits;
    // far prefix [0, pos0): one coarse table per 2^20 positions; near part [pos0, n_acc): fine tiles + ordered walk
    u64 pos0 = own_b > ht_margin ? own_b - ht_margin : 0;
    pos0 &= ~((1ull << ht_coarse_log) - 1);
    const u64 n_coarse = pos0 >> ht_coarse_log;
    const u64 n_tiles = (n_acc - pos0 + NLZM_HT_TILE - 1) / NLZM_HT_TILE;
    CKI(ensure(ht_coarse, (n_coarse + 1) * nc * 4));
    CKI(ensure(ht_cfirst, (n_coarse + 1) * nc * 4)); CKI(ensure(ht_clast, (n_coarse + 1) * nc * 4));
    CKI(ensure(ht_ccount, (n_coarse + 1) * nc * 4));
    CKI(ensure(ht_tab, n_tiles * nc * 4));
    CKI(ensure(ht_ps, (n_acc - pos0) * 4));
    if (c.rows == 2) { CKI(ensure(ht_pl, (n_acc - pos0) * 4)); CKI(ensure(ht_pr, (n_acc - pos0) * 4)); }
    u32 *base_row = ht_coarse.as<u32>() + n_coarse * nc;             // table at pos0
    const u64 max_tiles = n_tiles > n_coarse ? n_tiles : n_coarse;
    const u64 n_groups_max = (max_tiles + NLZM_HT_GROUP - 1) / NLZM_HT_GROUP;
    CKI(ensure(ht_gmax, n_groups_max * nc * 4));
    HtTableParams tp;
    tp.first_rows = nullptr; tp.last_rows = nullptr; tp.count_rows = nullptr;
    tp.x = x.as<u8>(); tp.c = c; tp.x_limit = g.flen + NLZM_X_PAD;
    tp.ps = ht_ps.as<u32>(); tp.pl = ht_pl.as<u32>(); tp.pr = ht_pr.as<u32>();
    HtScanParams sp;
    sp.group_max = ht_gmax.as<u32>(); sp.nc = (u32)nc;
    if (n_coarse) {
        tp.pos0 = 0; tp.n_acc = pos0; tp.tile_log = ht_coarse_log; tp.n_tiles = (u32)n_coarse;
        tp.tile_last = ht_coarse.as<u32>();
        tp.first_rows = ht_cfirst.as<u32>(); tp.last_rows = ht_clast.as<u32>(); tp.count_rows = ht_ccount.as<u32>();
        CKI(launch_ht_tile_last(tp, n_coarse, nc * 12, st));
        sp.tile_last = ht_coarse.as<u32>(); sp.init = nullptr; sp.final_row = base_row;
        sp.n_tiles = (u32)n_coarse; sp.n_groups = (u32)((n_coarse + NLZM_HT_GROUP - 1) / NLZM_HT_GROUP);
        launch_ht_tile_scan(sp, st);
    }
    tp.pos0 = pos0; tp.n_acc = n_acc; tp.tile_log = NLZM_HT_TILE_LOG; tp.n_tiles = (u32)n_tiles;
    tp.tile_last = ht_tab.as<u32>();
    tp.first_rows = nullptr; tp.last_rows = nullptr;
    CKI(launch_ht_tile_last(tp, n_tiles, nc * 4, st));
    sp.tile_last = ht_tab.as<u32>(); sp.init = n_coarse ? base_row : nullptr; sp.final_row = nullptr;
    sp.n_tiles = (u32)n_tiles; sp.n_groups = (u32)((n_tiles + NLZM_HT_GROUP - 1) / NLZM_HT_GROUP);
    launch_ht_tile_scan(sp, st);
    CKI(launch_ht_prev(tp, n_tiles, nc * 2 + NLZM_HT_STAGE + 16, st));
    HtFindParams fp{x.as<u8>(), g, c, ht_ps.as<u32>(), ht_pl.as<u32>(), ht_pr.as<u32>(), pos0, ht_coarse.as<u32>(), ht_cfirst.as<u32>(), ht_clast.as<u32>(), ht_ccount.as<u32>(), ht_coarse_log,
                    0u, nullptr, own_b, (mask & NLZM_MF_BT4) ? 1u : 0u, sink()};
    if (n_coarse) {
        // what every cell holds at pos0, resolved once per cell over the far prefix
        CKI(ensure(ht_snap, (nc + 1) * 4));
        HtSnapParams sn{fp, base_row, ht_snap.as<u32>()};
#ifndef NLZM_EMU
        launch_ht_snapshot(sn, (nc + 1) * 32, st);           // one warp per cell
#else
        launch_ht_snapshot(sn, nc + 1, st);
#endif
        fp.snap = ht_snap.as<u32>();
    }
    launch_ht_find(fp, n_acc - own_b, st);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Stage R
// ------------------------------------------------------------------------------------------------
int nlzm_mf::stage_rk(u64 own_b, u64 own_e) {
    if (g.flen < NLZM_RK_BLOCK) return 0;
    const u64 rk_e = own_e < g.flen - 255 ? own_e : g.flen - 255;    // RK is called while 256 bytes are visible
    if (rk_e <= own_b) return 0;
    // the carry state machine restarts cleanly at a ring shift: the last one at or before own_b ...
    u64 rk_b_shift = 0;
    const u32 ep = geom_epoch(g, own_b);
    if (ep > 0) {
        u64 k = (((u64)(ep + 1) << g.hb) + g.cs - 1) / g.cs;
        rk_b_shift = k * g.cs;
    }
    // ... and behind any 65536 positions without a valid hit (no carry outlives them). First try: look hits up only from
    // a little before the range and let the chain kernel find such a stretch; if there is none (dense hits), redo from
    // the ring shift. A late range then costs what its own positions cost, not what the prefix since the shift costs.
    const bool try_short = own_b > rk_b_shift + rk_restart + 2 * NLZM_RK_CARRY_MAX;
    if (try_short) {
        bool done = false;
        CKI(stage_rk_from(own_b, own_e, rk_e, own_b - rk_restart, true, done));
        if (done) return 0;
    }
    bool done = false;
    return stage_rk_from(own_b, own_e, rk_e, rk_b_shift, false, done);
}

// hits looked up from rk_b on; need_restart: the state at rk_b is unknown, *done = false if no restart point exists
int nlzm_mf::stage_rk_from(u64 own_b, u64 own_e, u64 rk_e, u64 rk_b, bool need_restart, bool &done) {
    done = true;
    const u64 n_blk = (rk_e + NLZM_RK_BLOCK - 1) / NLZM_RK_BLOCK;    // blocks starting before rk_e (each has 256 bytes)
    const u64 n_slots = 1ull << g.rk_bits;
    CKI(ensure(hblk, n_blk * 4));
    CKI(ensure(sl_k[0], n_blk * 4)); CKI(ensure(sl_k[1], n_blk * 4));
    CKI(ensure(sl_v[0], n_blk * 4)); CKI(ensure(sl_v[1], n_blk * 4));
    CKI(ensure(sl_cnt, (n_slots + 1) * 4)); CKI(ensure(sl_off, (n_slots + 1) * 4));
    const u64 range = rk_e - rk_b;
    u64 hit_cap = rk_all_hits ? range + 1024 : range / 4 + (1u << 20);
    CKI(ensure(hit_k[0], hit_cap * 8)); CKI(ensure(hit_k[1], hit_cap * 8));
    CKI(ensure(hit_v[0], hit_cap * 4)); CKI(ensure(hit_v[1], hit_cap * 4));
    CKI(ensure(hit_len, hit_cap * 4));
    CKI(ensure(iv, hit_cap * sizeof(RkInterval)));
    CKI(ensure_prim(n_blk > n_slots + 1 ? n_blk : n_slots + 1));
    CKI(ensure_prim(hit_cap));
    u32 *hit_count = scalars.as<u32>() + SC_RK_HITS;
    u32 *n_iv = scalars.as<u32>() + SC_RK_INTERVALS;

    CK(cudaMemsetAsync(sl_cnt.p, 0, (n_slots + 1) * 4, st));
    CK(cudaMemsetAsync(hit_count, 0, 8, st));
    RkBlockParams bp{x.as<u8>(), hblk.as<u32>(), sl_k[0].as<u32>(), sl_v[0].as<u32>(), sl_cnt.as<u32>(), 32 - g.rk_bits};
    launch_rk_block(bp, n_blk, st);
    int sel = 0;
    CKI(prim_sort_pairs32(tmp, sl_k[0].as<u32>(), sl_k[1].as<u32>(), sl_v[0].as<u32>(), sl_v[1].as<u32>(), n_blk, 0, (int)g.rk_bits, st, &sel));
    CKI(prim_exclusive_sum(tmp, sl_cnt.as<u32>(), sl_off.as<u32>(), n_slots + 1, st));

    RkLookupParams lp;
    lp.x = x.as<u8>(); lp.g = g; lp.hblk = hblk.as<u32>(); lp.slot_off = sl_off.as<u32>(); lp.slot_blk = sl_v[sel].as<u32>();
    lp.rk_b = rk_b; lp.rk_e = rk_e; lp.span = 64;
    lp.hit_keys = hit_k[0].as<u64>(); lp.hit_vals = hit_v[0].as<u32>(); lp.hit_count = hit_count; lp.hit_cap = (u32)hit_cap;
    launch_rk_lookup(lp, (range + lp.span - 1) / lp.span, st);
    u32 n_hits = 0;
    CKI(fetch_words(hit_count, &n_hits, 1));
    if (n_hits > hit_cap) {                      // dense hits (e.g. zero runs): redo with room for one hit per position
        rk_all_hits = true;
        rk_overflowed = true;
        return fail(NLZM_MF_E_OVERFLOW, "RK raw-hit buffer overflow");
    }
    if (n_hits == 0) return 0;                   // no raw hit in the looked-up prefix either: nothing is carried into the range
    // extension of every raw hit (any order); the few that are real hits are compacted, sorted by
    // position and fed to the sequential carry state machine
    u32 *n_valid = scalars.as<u32>() + SC_RK_VALID;
    CK(cudaMemsetAsync(n_valid, 0, 4, st));
    RkExtendParams xp{x.as<u8>(), g, hit_k[0].as<u64>(), hit_v[0].as<u32>(), hit_len.as<u32>(),
                      hit_k[1].as<u64>(), hit_v[1].as<u32>(), n_valid};
#ifndef NLZM_EMU
    launch_rk_extend_warp(xp, n_hits, st);
#else
    launch_rk_extend(xp, n_hits, st);
#endif
    u32 nv = 0;
    CKI(fetch_words(n_valid, &nv, 1));
    if (nv == 0) return 0;
    CKI(ensure(val_k, (u64)nv * 8)); CKI(ensure(val_v, (u64)nv * 4));
    int hsel = 0;
    CKI(prim_sort_pairs64(tmp, hit_k[1].as<u64>(), val_k.as<u64>(), hit_v[1].as<u32>(), val_v.as<u32>(), nv, 0, (int)bits_for(g.flen + 1), st, &hsel));
    RkChainParams cp{g, hsel ? val_k.as<u64>() : hit_k[1].as<u64>(), hsel ? val_v.as<u32>() : hit_v[1].as<u32>(),
                     hit_v[0].as<u32>(), hit_len.as<u32>(), n_valid, iv.as<RkInterval>(), n_iv,
                     rk_b, own_b, need_restart ? 1u : 0u, scalars.as<u32>() + SC_RK_OK};
    launch_rk_chain(cp, 1, st);
    u32 res[3] = {0, 0, 0};                      // n_iv, n_valid, ok (consecutive scalars)
    CKI(fetch_words(n_iv, res, 3));
    if (need_restart && !res[2]) { done = false; return 0; }     // hits too dense for a restart point: from the ring shift
    RkExpandParams ex{g, iv.as<RkInterval>(), own_b, own_e, (mask & NLZM_MF_BT4) ? 1u : 0u, sink()};
    launch_rk_expand(ex, res[0], st);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Stage M
// ------------------------------------------------------------------------------------------------
int nlzm_mf::stage_merge(u64 own_b, u64 own_e, Slot &s) {
    const u64 n_own = own_e - own_b;
    u32 nt = 0;
    CKI(fetch_words(tcount.p, &nt, 1));
    stats.tuples_last = nt;
    const u32 cap = sink().cap;
    if (nt > cap) return fail(NLZM_MF_E_OVERFLOW, "candidate tuple buffer overflow");
    CKI(ensure(s.d_offsets, (n_own + 1) * 4));
    CK(cudaMemsetAsync(s.d_offsets.p, 0, (n_own + 1) * 4, st));
    s.n_steps = 0;
    if (nt == 0) return 0;
    CKI(ensure(aux0, (n_own + 1) * 4));
    const u64 n_bins = (n_own + NLZM_BIN - 1) >> NLZM_BIN_LOG;
    CKI(ensure(keep, (n_bins + 2) * 4));
    CKI(ensure_prim(nt > n_own + 1 ? nt : n_own + 1));
    // sort by bin only (the position bits above the bin), finish every bin in shared memory
    int sel = 0;
    const int lo_bit = 9 + (int)NLZM_BIN_LOG, hi_bit = (int)(9 + bits_for(n_own + 1));
    if (hi_bit > lo_bit) {
        CKI(prim_sort_pairs64(tmp, tk[0].as<u64>(), tk[1].as<u64>(), tv[0].as<u32>(), tv[1].as<u32>(), nt, lo_bit, hi_bit, st, &sel));
    }
    u32 *count = aux0.as<u32>();
    CK(cudaMemsetAsync(count, 0, (n_own + 1) * 4, st));
    BinBoundsParams bb{tk[sel].as<u64>(), nt, keep.as<u32>()};
    launch_bin_bounds(bb, n_bins + 1, st);
    Step *staging = (Step *)tk[sel ^ 1].p;                        // the sort's other buffer is free now (8 bytes per tuple >= 6)
    BinFinishParams fp{tk[sel].as<u64>(), tv[sel].as<u32>(), keep.as<u32>(), (u32)n_own, count, staging};
    CKI(launch_bin_finish(fp, n_bins, NLZM_BIN_SMEM, st));
    CKI(prim_exclusive_sum(tmp, count, s.d_offsets.as<u32>(), n_own + 1, st));
    u32 total = 0;
    CKI(fetch_words(s.d_offsets.as<u32>() + n_own, &total, 1));
    s.n_steps = total;
    CKI(ensure(s.d_steps, (u64)(total ? total : 1) * sizeof(Step)));
    BinPlaceParams pp{staging, keep.as<u32>(), s.d_offsets.as<u32>(), (u32)n_own, s.d_steps.as<Step>()};
    CKI(launch_bin_place(pp, n_bins, 0, st));
    return 0;
}

int nlzm_mf::compute(u64 b, u64 e, Slot &s) {
    s.begin = b; s.end = e; s.n_steps = 0;
    const u64 n_own = e - b;
    const bool was_prepared = prepared && prep_b == b && prep_e == e;
    if (was_prepared) prepared = false;
    for (int attempt = 0; attempt < 4; attempt++) {
        const u64 cap = n_own * tuple_cap_mult + tuple_cap_extra;
        if (cap >= 0xFFFFFFF0ull) return fail(NLZM_MF_E_OVERFLOW, "candidate tuple capacity exceeds 2^32");
        const bool continue_prepared = was_prepared && attempt == 0;
        CKI(ensure(tk[0], cap * 8)); CKI(ensure(tk[1], cap * 8));
        CKI(ensure(tv[0], cap * 4)); CKI(ensure(tv[1], cap * 4));
        if (!continue_prepared) {
            CK(cudaMemsetAsync(tcount.p, 0, 4, st));
            stats.ms_rank = stats.ms_levels = 0;
        } else {
            // stages S/T of this range ran in nlzm_mf_prepare: take its blocks and candidate tuples back
            fresh = prep_fresh;
            prep_fresh.clear();
            fresh_u0 = b;
            if (prep_nt) {
                CK(cudaMemcpyAsync(tk[0].p, prep_tk.p, (size_t)prep_nt * 8, cudaMemcpyDeviceToDevice, st));
                CK(cudaMemcpyAsync(tv[0].p, prep_tv.p, (size_t)prep_nt * 4, cudaMemcpyDeviceToDevice, st));
            }
            CK(cudaMemcpyAsync(tcount.p, &prep_nt, 4, cudaMemcpyHostToDevice, st));
            CK(cudaStreamSynchronize(st));
            stats.ms_rank = prep_ms_rank; stats.ms_levels = prep_ms_levels;
        }
        stats.ms_cross = 0;
        stats.segments_queried = 0;
        Ev e0, e1, e2, e3, e4;
        cudaEventRecord(e0, st);
        int r = 0;
        std::vector<Segment> behind_all;
        bool have_behind = false;
        if (n_own > 0) {
            if ((mask & NLZM_MF_BT4) && g.flen >= 4) {
                // window behind the range: from retained segments when they cover it, else re-ranked with the range.
                // A prepared range has no halo by construction: its segments must have been imported by now.
                std::vector<Segment> &behind = behind_all;
                const bool covered = covered_by_segments(b, behind);
                if (was_prepared && !covered)
                    return fail(NLZM_MF_E_STATE, "prepared range: the window behind it is not covered by imported segments");
                const u64 halo_b = b > (u64)(g.W - 1) ? b - (g.W - 1) : 0;
                if (!continue_prepared) r = stage_bt4_own(b, e, covered ? b : halo_b);
                if (!covered) behind.clear();
                have_behind = true;
            }
            // stages H and R do not depend on the segments behind the range: they run while asynchronous imports
            // are still crossing NVLink; the cross passes wait for the copies
            cudaEventRecord(e1, st);
            if (r == 0 && (mask & NLZM_MF_HT2)) { HtCfg c{1, 12, 2}; r = stage_ht(b, e, c); }
            if (r == 0 && (mask & NLZM_MF_HT3)) { HtCfg c{2, g.ht3_bits, 3}; r = stage_ht(b, e, c); }
            cudaEventRecord(e2, st);
            if (r == 0 && (mask & NLZM_MF_RK256)) r = stage_rk(b, e);
            cudaEventRecord(e3, st);
            if (r == 0 && have_behind && !behind_all.empty()) {
#ifndef NLZM_EMU
                if (imports_pending) { CK(cudaStreamWaitEvent(st, import_ev, 0)); imports_pending = false; }
#endif
                r = stage_bt4_cross(b, e, behind_all);
            }
        } else {
            cudaEventRecord(e1, st); cudaEventRecord(e2, st); cudaEventRecord(e3, st);
        }
        if (r == 0) r = stage_merge(b, e, s);
        cudaEventRecord(e4, st);
        cudaStreamSynchronize(st);
        if (r == 0) {
            std::lock_guard<std::mutex> l(stats_mu);
            cudaEventElapsedTime(&stats.ms_ht, e1, e2);
            cudaEventElapsedTime(&stats.ms_rk, e2, e3);
            cudaEventElapsedTime(&stats.ms_merge, e3, e4);
            cudaEventElapsedTime(&stats.ms_total, e0, e4);
            if (continue_prepared) stats.ms_total += prep_ms;
        }
        if (r == NLZM_MF_E_OVERFLOW && attempt < 3) {      // rare: dense candidates; grow and redo the range
            if (rk_overflowed) rk_overflowed = false; else tuple_cap_mult *= 2;
            fresh.clear();
            continue;
        }
        if (r) { fresh.clear(); return r; }
        break;
    }
#ifndef NLZM_EMU
    if (imports_pending) { cudaStreamSynchronize(st_copy); imports_pending = false; }   // imported but not queried: let the copies land before the buffers go
#endif
    retain_fresh(e);
    return 0;
}

// Stages S/T of [b, e) alone (no window behind it): what other engines need before they can import this
// range's segments, and what nlzm_mf_find(b, e) then continues from.
int nlzm_mf::prepare_impl(u64 b, u64 e) {
    Turn turn(this);
    if (!have_input) return fail(NLZM_MF_E_STATE, "prepare before set_input");
    if (b > e || e > g.flen) return fail(NLZM_MF_E_ARG, "bad range");
    if (e - b > (1ull << 28) || e - b > max_range) return fail(NLZM_MF_E_ARG, "range too large");
    if (!(mask & NLZM_MF_BT4) || g.flen < 4 || e == b) { prepared = true; prep_b = b; prep_e = e; prep_nt = 0; prep_fresh.clear(); prep_ms = prep_ms_rank = prep_ms_levels = 0; return 0; }
#ifndef NLZM_EMU
    CK(cudaSetDevice(device));
#endif
    prepared = false;
#ifndef NLZM_EMU
    if (imports_pending) { cudaStreamSynchronize(st_copy); imports_pending = false; }   // imports nobody waited for
#endif
    stats.ms_import = 0;
    stats.bytes_imported = 0;
    {
        std::vector<Segment> own_only;                 // imports of an earlier round are stale by definition
        for (Segment &sg : segs) if (!sg.imported) own_only.push_back(sg);
        segs.swap(own_only);
    }
    const u64 n_own = e - b;
    for (int attempt = 0; attempt < 4; attempt++) {
        const u64 cap = n_own * tuple_cap_mult + tuple_cap_extra;
        if (cap >= 0xFFFFFFF0ull) return fail(NLZM_MF_E_OVERFLOW, "candidate tuple capacity exceeds 2^32");
        CKI(ensure(tk[0], cap * 8)); CKI(ensure(tk[1], cap * 8));
        CKI(ensure(tv[0], cap * 4)); CKI(ensure(tv[1], cap * 4));
        CK(cudaMemsetAsync(tcount.p, 0, 4, st));
        Ev e0, e1;
        cudaEventRecord(e0, st);
        int r = stage_bt4_own(b, e, b);
        u32 nt = 0;
        if (r == 0) {
            cudaEventRecord(e1, st);
            CKI(fetch_words(tcount.p, &nt, 1));
            cudaEventElapsedTime(&stats.ms_prepare, e0, e1);
            if (nt > sink().cap) r = NLZM_MF_E_OVERFLOW;
        }
        if (r == NLZM_MF_E_OVERFLOW && attempt < 3) { tuple_cap_mult *= 2; fresh.clear(); continue; }
        if (r) { fresh.clear(); return r == NLZM_MF_E_OVERFLOW ? fail(r, "candidate tuple buffer overflow") : r; }
        break;
    }
    if (g_prof.on) prof_resolve();
    // the blocks are exported from the retained list; find(b, e) picks them up again as its own universe
    u32 nt = 0;
    CKI(fetch_words(tcount.p, &nt, 1));
    CKI(ensure(prep_tk, (size_t)(nt ? nt : 1) * 8)); CKI(ensure(prep_tv, (size_t)(nt ? nt : 1) * 4));
    if (nt) {
        CK(cudaMemcpyAsync(prep_tk.p, tk[0].p, (size_t)nt * 8, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(prep_tv.p, tv[0].p, (size_t)nt * 4, cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
    }
    prep_nt = nt;
    prep_fresh = fresh;
    prep_ms_rank = stats.ms_rank; prep_ms_levels = stats.ms_levels; prep_ms = stats.ms_prepare;
    add_segments(fresh);
    fresh.clear();
    prepared = true; prep_b = b; prep_e = e;
    return 0;
}

int nlzm_mf::find_impl(u64 b, u64 e, int si, bool to_host, u64 ticket) {
    Turn turn(this, ticket);
    if (!have_input) return fail(NLZM_MF_E_STATE, "find before set_input");
    if (si < 0 || si > 1) return fail(NLZM_MF_E_ARG, "slot must be 0 or 1");
    if (b > e || e > g.flen) return fail(NLZM_MF_E_ARG, "bad range");
    if (e - b > (1ull << 28)) return fail(NLZM_MF_E_ARG, "range larger than 2^28 positions: split it");
    if (e - b > max_range) return fail(NLZM_MF_E_ARG, "range larger than config.max_range");
#ifndef NLZM_EMU
    CK(cudaSetDevice(device));
#endif
    Slot &s = slot[si];
    CKI(compute(b, e, s));
    const u64 n_own = e - b;
    {
        std::lock_guard<std::mutex> l(stats_mu);
        stats.kernel_launches = g_launches.load();
        stats.ms_d2h = 0;
    }
    if (g_prof.on) prof_resolve();
    if (!to_host) return 0;
    // Device -> host on the copy stream, outside the compute lock: the next range's kernels run while this
    // range's records cross PCIe. The slot's device buffers stay untouched until its next find.
    size_t ob = (n_own + 1) * 4, sb = (size_t)(s.n_steps ? s.n_steps : 1) * sizeof(Step);
    if (ob > s.h_offsets_bytes) {
        if (s.h_offsets) cudaFreeHost(s.h_offsets);
        s.h_offsets = nullptr; s.h_offsets_bytes = 0;
        CK(cudaMallocHost(&s.h_offsets, ob + (ob >> 3)));
        s.h_offsets_bytes = ob + (ob >> 3);
    }
    if (sb > s.h_steps_bytes) {
        if (s.h_steps) cudaFreeHost(s.h_steps);
        s.h_steps = nullptr; s.h_steps_bytes = 0;
        CK(cudaMallocHost(&s.h_steps, sb + (sb >> 3)));
        s.h_steps_bytes = sb + (sb >> 3);
    }
    const u64 n_steps = s.n_steps;
    void *d_off = s.d_offsets.p, *d_steps = s.d_steps.p;
    turn.release();
    Ev e0, e1;
    cudaEventRecord(e0, st_copy);
    cudaError_t ce = cudaMemcpyAsync(s.h_offsets, d_off, ob, cudaMemcpyDeviceToHost, st_copy);
    if (ce == cudaSuccess && n_steps) ce = cudaMemcpyAsync(s.h_steps, d_steps, n_steps * sizeof(Step), cudaMemcpyDeviceToHost, st_copy);
    cudaEventRecord(e1, st_copy);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st_copy);
    if (ce != cudaSuccess) return fail((int)ce, std::string("device->host copy: ") + cudaGetErrorString(ce));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::lock_guard<std::mutex> l(stats_mu);
    stats.ms_d2h = ms;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int nlzm_mf_abi_version(void) { return NLZM_MF_ABI_VERSION; }

int nlzm_mf_get_geometry(uint64_t file_len, uint32_t hist_bits, nlzm_mf_geometry *out) {
    if (!out) return NLZM_MF_E_ARG;
    Geom g = make_geom(file_len, hist_bits);
    out->hist_bits = g.hb;
    out->window = g.W;
    out->frame_bits = nlzm_clampu(g.hb - 2, 14, 17);
    out->chunk_size = g.cs;
    out->feed_size = g.cs + NLZM_MATCH_MAX + 1;
    out->ht2_bits = 12;
    out->ht3_bits = g.ht3_bits;
    out->bt4_bits = g.bt_bits;
    out->rk_bits = g.rk_bits;
    return 0;
}

int nlzm_mf_create(const nlzm_mf_config *cfg, nlzm_mf **out) {
    if (!cfg || !out || cfg->struct_size != sizeof(nlzm_mf_config)) { g_create_error = "bad config"; return NLZM_MF_E_ARG; }
    if (cfg->file_len >= (1ull << 31)) { g_create_error = "file_len must be < 2^31"; return NLZM_MF_E_ARG; }
    *out = nullptr;
#ifndef NLZM_EMU
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device (") + cudaGetErrorString(e) + "); this engine has no CPU fallback";
        return NLZM_MF_E_NODEVICE;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "bad device ordinal"; return NLZM_MF_E_ARG; }
    cudaSetDevice(cfg->device);
#endif
    nlzm_mf *mf = new (std::nothrow) nlzm_mf();
    if (!mf) { g_create_error = "out of host memory"; return NLZM_MF_E_NOMEM; }
    mf->g = make_geom(cfg->file_len, cfg->hist_bits);
    mf->device = cfg->device;
    mf->mask = cfg->finder_mask ? cfg->finder_mask : (u32)NLZM_MF_ALL;
    mf->max_range = cfg->max_range ? cfg->max_range : cfg->file_len;
#ifndef NLZM_EMU
    if (cudaStreamCreateWithFlags(&mf->st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&mf->st_copy, cudaStreamNonBlocking) != cudaSuccess) {
        g_create_error = "cudaStreamCreate failed";
        delete mf;
        return NLZM_MF_E_NODEVICE;
    }
#endif
#ifndef NLZM_EMU
    if (cudaHostAlloc((void **)&mf->h_words, 256, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void **)&mf->d_words, mf->h_words, 0) != cudaSuccess) {
        g_create_error = "cudaHostAlloc(mapped) failed";
        nlzm_mf_destroy(mf);
        return NLZM_MF_E_NOMEM;
    }
#endif
    int r = mf->ensure(mf->x, cfg->file_len + NLZM_X_PAD + 8);
    if (r == 0) r = mf->ensure(mf->tcount, 64);
    if (r == 0) r = mf->ensure(mf->scalars, 256);
    if (r) { g_create_error = mf->err; nlzm_mf_destroy(mf); return r; }
    *out = mf;
    return 0;
}

void nlzm_mf_destroy(nlzm_mf *mf) {
    if (!mf) return;
#ifndef NLZM_EMU
    cudaSetDevice(mf->device);
#endif
    for (auto &s : mf->slot) if (s.worker.joinable()) s.worker.join();
#ifndef NLZM_EMU
    if (mf->st_copy) cudaStreamSynchronize(mf->st_copy);
#endif
    mf->fresh.clear();
    mf->prep_fresh.clear();
    mf->segs.clear();                                  // buffers go back to the pool, which is freed below
#ifndef NLZM_EMU
    for (auto &kv : mf->ipc_open) cudaIpcCloseMemHandle(kv.second);
#endif
    mf->ipc_open.clear();
    for (auto &b : mf->pool) if (b.p) cudaFree(b.p);
    mf->pool.clear();
    for (auto &s : mf->slot) {
        if (s.worker.joinable()) s.worker.join();
        mf->release(s.d_offsets); mf->release(s.d_steps);
        if (s.h_offsets) cudaFreeHost(s.h_offsets);
        if (s.h_steps) cudaFreeHost(s.h_steps);
    }
    DevBuf *all[] = {&mf->x, &mf->k64[0], &mf->k64[1], &mf->v32[0], &mf->v32[1], &mf->rank, &mf->ptr, &mf->el[0], &mf->el[1], &mf->part, &mf->aux0,
                     &mf->aux1, &mf->tk[0], &mf->tk[1], &mf->tv[0], &mf->tv[1], &mf->tcount, &mf->keep, &mf->out_idx,
                     &mf->e_k[0], &mf->e_k[1], &mf->e_v[0], &mf->e_v[1], &mf->e_inv, &mf->ht_snap, &mf->ht_tab, &mf->ht_gmax, &mf->ht_coarse, &mf->ht_cfirst, &mf->ht_clast, &mf->ht_ccount, &mf->ht_ps, &mf->ht_pl, &mf->ht_pr, &mf->hblk, &mf->sl_k[0], &mf->sl_k[1],
                     &mf->sl_v[0], &mf->sl_v[1], &mf->sl_cnt, &mf->sl_off, &mf->hit_k[0], &mf->hit_k[1], &mf->hit_v[0],
                     &mf->hit_v[1], &mf->hit_len, &mf->iv, &mf->val_k, &mf->val_v, &mf->scalars, &mf->tmpbuf, &mf->prep_tk, &mf->prep_tv, &mf->xstage};
    for (DevBuf *b : all) mf->release(*b);
#ifndef NLZM_EMU
    if (mf->import_ev) cudaEventDestroy(mf->import_ev);
    if (mf->st) cudaStreamDestroy(mf->st);
    if (mf->st_copy) cudaStreamDestroy(mf->st_copy);
    if (mf->h_words) cudaFreeHost(mf->h_words);
#endif
    delete mf;
}

const char *nlzm_mf_last_error(const nlzm_mf *mf) { return mf ? mf->err.c_str() : g_create_error.c_str(); }

static int set_input_common(nlzm_mf *mf, const void *src, uint64_t len, bool from_device) {
    if (!mf || (!src && len)) return NLZM_MF_E_ARG;
    if (len != mf->g.flen) return mf->fail(NLZM_MF_E_ARG, "set_input length differs from config.file_len");
    Turn turn(mf);
#ifndef NLZM_EMU
    cudaSetDevice(mf->device);
#endif
    cudaError_t e = cudaMemsetAsync((u8 *)mf->x.p + len, 0, NLZM_X_PAD + 8, mf->st);
    if (e == cudaSuccess && len)
        e = cudaMemcpyAsync(mf->x.p, src, len, from_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, mf->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(mf->st);
    if (e != cudaSuccess) return mf->fail((int)e, std::string("set_input: ") + cudaGetErrorString(e));
#ifndef NLZM_EMU
    if (mf->imports_pending) { cudaStreamSynchronize(mf->st_copy); mf->imports_pending = false; }
#endif
    mf->segs.clear();                                  // segments describe the previous bytes
    mf->fresh.clear();
    mf->prep_fresh.clear();
    mf->prepared = false;
    mf->have_input = true;
    return 0;
}

int nlzm_mf_set_input(nlzm_mf *mf, const uint8_t *host_data, uint64_t len) { return set_input_common(mf, host_data, len, false); }
int nlzm_mf_set_input_device(nlzm_mf *mf, const void *device_data, uint64_t len) { return set_input_common(mf, device_data, len, true); }

static void fill_view(nlzm_mf *mf, int slot, bool host, nlzm_mf_view *out) {
    Slot &s = mf->slot[slot];
    out->begin = s.begin;
    out->end = s.end;
    out->n_steps = s.n_steps;
    out->offsets = host ? (const uint32_t *)s.h_offsets : s.d_offsets.as<uint32_t>();
    out->steps = host ? (const nlzm_mf_step *)s.h_steps : (const nlzm_mf_step *)s.d_steps.p;
}

int nlzm_mf_find(nlzm_mf *mf, uint64_t begin, uint64_t end, int slot, nlzm_mf_view *out) {
    if (!mf || !out) return NLZM_MF_E_ARG;
    int r = mf->find_impl(begin, end, slot, true, Turn::NONE);
    if (r == 0) fill_view(mf, slot, true, out);
    return r;
}

int nlzm_mf_find_device(nlzm_mf *mf, uint64_t begin, uint64_t end, int slot, nlzm_mf_view *out) {
    if (!mf || !out) return NLZM_MF_E_ARG;
    int r = mf->find_impl(begin, end, slot, false, Turn::NONE);
    if (r == 0) fill_view(mf, slot, false, out);
    return r;
}

int nlzm_mf_submit(nlzm_mf *mf, uint64_t begin, uint64_t end, int slot) {
    if (!mf || slot < 0 || slot > 1) return NLZM_MF_E_ARG;
    Slot &s = mf->slot[slot];
    if (s.pending) return mf->fail(NLZM_MF_E_STATE, "slot already has a pending submit");
    if (s.worker.joinable()) s.worker.join();
    const u64 ticket = mf->take_turn();            // the worker computes in the order of the calls, not of the threads
    try {
        s.worker = std::thread([mf, begin, end, slot, ticket]() {
            int r;
            try { r = mf->find_impl(begin, end, slot, true, ticket); }
            catch (const std::bad_alloc &) { r = NLZM_MF_E_NOMEM; }
            catch (...) { r = NLZM_MF_E_STATE; }
            mf->slot[slot].status = r;
        });
    } catch (const std::exception &ex) {
        mf->wait_turn(ticket);
        mf->end_turn();
        return mf->fail(NLZM_MF_E_NOMEM, std::string("cannot start the submit thread: ") + ex.what());
    }
    s.pending = true;
    return 0;
}

int nlzm_mf_fetch(nlzm_mf *mf, int slot, nlzm_mf_view *out) {
    if (!mf || !out || slot < 0 || slot > 1) return NLZM_MF_E_ARG;
    Slot &s = mf->slot[slot];
    if (!s.pending) return mf->fail(NLZM_MF_E_STATE, "fetch without submit");
    s.worker.join();
    s.pending = false;
    if (s.status == 0) fill_view(mf, slot, true, out);
    return s.status;
}

int nlzm_mf_prepare(nlzm_mf *mf, uint64_t begin, uint64_t end) {
    if (!mf) return NLZM_MF_E_ARG;
    return mf->prepare_impl(begin, end);
}

int nlzm_mf_export_segments(nlzm_mf *mf, nlzm_mf_segment *out, uint32_t cap, uint32_t *n_out) {
    if (!mf || !n_out) return NLZM_MF_E_ARG;
    Turn turn(mf);
#ifndef NLZM_EMU
    cudaSetDevice(mf->device);
#endif
    uint32_t i = 0;
    for (const Segment &sg : mf->segs) {
        if (sg.imported) continue;                     // only what this engine computed itself
        if (out && i < cap) {
            nlzm_mf_segment &d = out[i];
            memset(&d, 0, sizeof d);
            d.pos_begin = sg.pos_b;
            d.pos_end = sg.pos_e;
            d.origin = sg.u0;
            d.n_elems = sg.n_elems;
            d.elems_offset_bytes = sg.elem_off * sizeof(Elem);
            d.elems_bytes = (u64)sg.n_elems * sizeof(Elem);
            d.ptrs_offset_bytes = (sg.pos_b - sg.ptr_pos0) * sizeof(PtrEntry);
            d.ptrs_bytes = (sg.pos_e - sg.pos_b) * sizeof(PtrEntry);
            d.elems_alloc = sg.bufs->el.p;
            d.ptrs_alloc = sg.bufs->ptr.p;
            d.device = mf->device;
#ifndef NLZM_EMU
            auto handle_of = [&](DevBuf &b, uint8_t *dst) -> bool {
                auto it = mf->ipc_made.find(b.p);
                if (it == mf->ipc_made.end()) {
                    cudaIpcMemHandle_t h;
                    if (cudaIpcGetMemHandle(&h, b.p) != cudaSuccess) { cudaGetLastError(); return false; }
                    it = mf->ipc_made.emplace(b.p, std::string((const char *)&h, sizeof h)).first;
                }
                b.shared = true;
                memcpy(dst, it->second.data(), sizeof(cudaIpcMemHandle_t));
                return true;
            };
            if (handle_of(sg.bufs->el, d.ipc_elems)) d.flags |= 1u;
            if (handle_of(sg.bufs->ptr, d.ipc_ptrs)) d.flags |= 2u;
#endif
        }
        ++i;
    }
    *n_out = i;
    return 0;
}

int nlzm_mf_import_segment(nlzm_mf *mf, const nlzm_mf_segment *d, int via_ipc) {
    if (!mf || !d || d->pos_end <= d->pos_begin || d->pos_begin < d->origin) return NLZM_MF_E_ARG;
    if (d->elems_bytes != d->n_elems * sizeof(Elem) || d->ptrs_bytes != (d->pos_end - d->pos_begin) * sizeof(PtrEntry))
        return mf->fail(NLZM_MF_E_ARG, "segment descriptor sizes do not match this library's element layout");
    Turn turn(mf);
    if (d->pos_end > mf->g.flen) return mf->fail(NLZM_MF_E_ARG, "segment lies outside this engine's input");
#ifndef NLZM_EMU
    cudaSetDevice(mf->device);
#endif
    auto bufs = std::make_shared<SegBufs>();
    bufs->owner = mf;
    // the level-array work buffers are idle between finds: lend them to the imports (stage T takes them back)
    mf->to_pool(mf->el[0]); mf->to_pool(mf->el[1]); mf->to_pool(mf->ptr);
    int r = mf->ensure_pooled(bufs->el, (size_t)d->elems_bytes);
    if (r == 0) r = mf->ensure_pooled(bufs->ptr, (size_t)d->ptrs_bytes);
    if (r) return r;
    const bool async = (via_ipc & 0x100) != 0;           // copies go to the copy stream; the next find waits for them
    via_ipc &= 0xFF;
    cudaStream_t cst = async ? mf->st_copy : mf->st;
    (void)cst;
    const u8 *src_el = (const u8 *)d->elems_alloc, *src_ptr = (const u8 *)d->ptrs_alloc;
    if (via_ipc == 2) {
        // elems_alloc / ptrs_alloc are HOST copies of the two slices (nlzm_mf_read_segment on the exporting side)
        cudaError_t e = cudaMemcpyAsync(bufs->el.p, src_el, (size_t)d->elems_bytes, cudaMemcpyHostToDevice, mf->st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(bufs->ptr.p, src_ptr, (size_t)d->ptrs_bytes, cudaMemcpyHostToDevice, mf->st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(mf->st);
        if (e != cudaSuccess) return mf->fail((int)e, std::string("segment upload: ") + cudaGetErrorString(e));
    } else {
#ifndef NLZM_EMU
    void *open_el = nullptr, *open_ptr = nullptr;
    if (via_ipc) {
        if ((d->flags & 3u) != 3u) return mf->fail(NLZM_MF_E_ARG, "segment descriptor carries no IPC handles");
        // a mapping stays open for the life of the engine: the exporter keeps the allocation alive (DevBuf::shared)
        auto mapped = [&](const uint8_t *raw, void **out) -> cudaError_t {
            const std::string key((const char *)raw, sizeof(cudaIpcMemHandle_t));
            auto it = mf->ipc_open.find(key);
            if (it != mf->ipc_open.end()) { *out = it->second; return cudaSuccess; }
            cudaIpcMemHandle_t h;
            memcpy(&h, raw, sizeof h);
            cudaError_t e = cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
            if (e == cudaSuccess) mf->ipc_open.emplace(key, *out);
            return e;
        };
        cudaError_t e = mapped(d->ipc_elems, &open_el);
        if (e == cudaSuccess) e = mapped(d->ipc_ptrs, &open_ptr);
        if (e != cudaSuccess) return mf->fail((int)e, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
        src_el = (const u8 *)open_el;
        src_ptr = (const u8 *)open_ptr;
    }
    Ev c0, c1;
    cudaEventRecord(c0, cst);
    cudaError_t e;
    if (via_ipc) {
        // an IPC mapping is an ordinary device pointer of this process (unified addressing finds the owner GPU)
        e = cudaMemcpyAsync(bufs->el.p, src_el + d->elems_offset_bytes, (size_t)d->elems_bytes, cudaMemcpyDefault, cst);
        if (e == cudaSuccess) e = cudaMemcpyAsync(bufs->ptr.p, src_ptr + d->ptrs_offset_bytes, (size_t)d->ptrs_bytes, cudaMemcpyDefault, cst);
    } else {
        e = cudaMemcpyPeerAsync(bufs->el.p, mf->device, src_el + d->elems_offset_bytes, d->device, (size_t)d->elems_bytes, cst);
        if (e == cudaSuccess) e = cudaMemcpyPeerAsync(bufs->ptr.p, mf->device, src_ptr + d->ptrs_offset_bytes, d->device, (size_t)d->ptrs_bytes, cst);
    }
    cudaEventRecord(c1, cst);
    if (async) {
        if (e == cudaSuccess) {
            if (!mf->import_ev) cudaEventCreateWithFlags(&mf->import_ev, cudaEventDisableTiming);
            cudaEventRecord(mf->import_ev, cst);
            mf->imports_pending = true;
            std::lock_guard<std::mutex> l(mf->stats_mu);
            mf->stats.bytes_imported += d->elems_bytes + d->ptrs_bytes;
        }
    } else {
        if (e == cudaSuccess) e = cudaStreamSynchronize(cst);
        if (e == cudaSuccess) {
            float ms = 0;
            cudaEventElapsedTime(&ms, c0, c1);
            std::lock_guard<std::mutex> l(mf->stats_mu);
            mf->stats.ms_import += ms;
            mf->stats.bytes_imported += d->elems_bytes + d->ptrs_bytes;
        }
    }
    if (e != cudaSuccess) return mf->fail((int)e, std::string("segment copy: ") + cudaGetErrorString(e));
#else
    (void)via_ipc;
    memcpy(bufs->el.p, src_el + d->elems_offset_bytes, (size_t)d->elems_bytes);
    memcpy(bufs->ptr.p, src_ptr + d->ptrs_offset_bytes, (size_t)d->ptrs_bytes);
#endif
    }
    Segment sg;
    sg.bufs = bufs;
    sg.u0 = d->origin;
    sg.elem_off = 0;
    sg.n_elems = (u32)d->n_elems;
    sg.ptr_pos0 = d->pos_begin;
    sg.pos_b = d->pos_begin;
    sg.pos_e = d->pos_end;
    sg.imported = true;
    for (const Segment &o : mf->segs) if (o.pos_b == sg.pos_b && o.pos_e == sg.pos_e) return 0;   // already there
    mf->segs.push_back(sg);
    return 0;
}

// copy of a retained segment's two slices into host memory (elems_bytes / ptrs_bytes of its descriptor):
// the transport of last resort when neither a peer copy nor CUDA IPC is possible
int nlzm_mf_read_segment(nlzm_mf *mf, uint32_t index, void *elems_host, void *ptrs_host) {
    if (!mf || !elems_host || !ptrs_host) return NLZM_MF_E_ARG;
    Turn turn(mf);
    const Segment *found = nullptr;                    // index counts the segments nlzm_mf_export_segments lists
    uint32_t k = 0;
    for (const Segment &o : mf->segs) if (!o.imported && k++ == index) { found = &o; break; }
    if (!found) return mf->fail(NLZM_MF_E_ARG, "no such segment");
#ifndef NLZM_EMU
    cudaSetDevice(mf->device);
#endif
    const Segment &sg = *found;
    cudaError_t e = cudaMemcpyAsync(elems_host, sg.elems(), (size_t)sg.n_elems * sizeof(Elem), cudaMemcpyDeviceToHost, mf->st);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(ptrs_host, (const u8 *)sg.bufs->ptr.p + (sg.pos_b - sg.ptr_pos0) * sizeof(PtrEntry),
                            (size_t)(sg.pos_e - sg.pos_b) * sizeof(PtrEntry), cudaMemcpyDeviceToHost, mf->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(mf->st);
    if (e != cudaSuccess) return mf->fail((int)e, std::string("read_segment: ") + cudaGetErrorString(e));
    return 0;
}

int nlzm_mf_trim_segments(nlzm_mf *mf, uint64_t from_pos) {
    if (!mf) return NLZM_MF_E_ARG;
    Turn turn(mf);
#ifndef NLZM_EMU
    cudaSetDevice(mf->device);
#endif
    return mf->trim_segments(from_pos);
}

int nlzm_mf_publish_segments(nlzm_mf *mf, uint64_t from_pos, nlzm_mf_segment *out, uint32_t cap, uint32_t *n_out) {
    if (!mf || !n_out) return NLZM_MF_E_ARG;
    Turn turn(mf);
#ifndef NLZM_EMU
    cudaSetDevice(mf->device);
#endif
    return mf->publish_segments(from_pos, out, cap, n_out);
}

int nlzm_mf_drop_segments(nlzm_mf *mf) {
    if (!mf) return NLZM_MF_E_ARG;
    Turn turn(mf);
#ifndef NLZM_EMU
    if (mf->imports_pending) { cudaStreamSynchronize(mf->st_copy); mf->imports_pending = false; }
#endif
    mf->segs.clear();
    mf->fresh.clear();
    mf->prep_fresh.clear();
    mf->prepared = false;
    return 0;
}

int nlzm_mf_profile(int enable) {
    g_prof.on.store(enable != 0);
    if (!enable) {
        std::lock_guard<std::mutex> l(g_prof.mu);
        g_prof.acc.clear();
    }
    return 0;
}

int nlzm_mf_get_kernel_times(nlzm_mf_kernel_time *out, uint32_t cap, uint32_t *n_out) {
    if (!n_out) return NLZM_MF_E_ARG;
    std::lock_guard<std::mutex> l(g_prof.mu);
    uint32_t i = 0;
    for (auto &kv : g_prof.acc) {
        if (out && i < cap) {
            memset(&out[i], 0, sizeof out[i]);
            strncpy(out[i].name, kv.first.c_str(), sizeof(out[i].name) - 1);
            out[i].launches = kv.second.first;
            out[i].ms = kv.second.second;
        }
        ++i;
    }
    *n_out = i;
    return 0;
}

int nlzm_mf_set_option(nlzm_mf *mf, const char *key, uint64_t value) {
    if (!mf || !key) return NLZM_MF_E_ARG;
    Turn turn(mf);
    const std::string k(key);
    if (k == "ht_margin") { mf->ht_margin = value; return 0; }
    if (k == "rk_restart") { mf->rk_restart = value; return 0; }
    if (k == "tuple_cap_mult") { mf->tuple_cap_mult = value ? (u32)value : 1u; return 0; }
    if (k == "tuple_cap_extra") { mf->tuple_cap_extra = value; return 0; }
    if (k == "retain") { mf->retain = value != 0; if (!mf->retain) mf->segs.clear(); return 0; }
    if (k == "max_segments") { mf->max_segments = (u32)value; return 0; }
    if (k == "ht_coarse_log") {
        if (value < NLZM_HT_TILE_LOG && value < 10) return mf->fail(NLZM_MF_E_ARG, "ht_coarse_log too small");
        if (value > 30) return mf->fail(NLZM_MF_E_ARG, "ht_coarse_log too large");
        mf->ht_coarse_log = (u32)value;
        return 0;
    }
    return mf->fail(NLZM_MF_E_ARG, "unknown option " + k);
}

int nlzm_mf_get_stats(const nlzm_mf *mf, nlzm_mf_stats *out) {
    if (!mf || !out) return NLZM_MF_E_ARG;
    {
        std::lock_guard<std::mutex> l(const_cast<nlzm_mf *>(mf)->stats_mu);
        *out = mf->stats;
    }
    out->kernel_launches = g_launches.load();
    return 0;
}

} // extern "C"
