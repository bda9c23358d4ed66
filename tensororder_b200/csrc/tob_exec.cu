// C-ABI entry points (include/tob200.h) and the device executor: arena, leaf upload, slice loop
// as a CUDA graph, per-op profiling, stand-alone tensordot / permute.
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cmath>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "tob_kernels.cuh"

namespace tob {
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
}  // namespace tob

using namespace tob;

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                    \
            return (_e == cudaErrorMemoryAllocation) ? TOB_E_OOM : TOB_E_CUDA;                \
        }                                                                                     \
    } while (0)

constexpr int kMaxLanes = 2;        // slices in flight at once (each lane: own arena, workspace, stream)
constexpr int kMaxResults = 4096;   // per-slice results buffered on the device before the ordered sum
constexpr int kMaxBranches = 32;    // streams per lane for the DAG schedule (Op::branch); branch 0 is the lane's stream

struct Lane {
    cudaStream_t stream = nullptr;  // lane 0: the plan's stream (own or caller's); lane >= 1: own stream
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;  // lane 0: run start / end;  lane >= 1: fork / done
    double* d_arena = nullptr;
    double* d_ws = nullptr;
    unsigned* d_sk_flags = nullptr;  // stream-K counter + flags (zeroed at upload, self-cleaning afterwards)
    long long* d_leaf_off = nullptr;
    DevState* d_state = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    cudaGraph_t inv_graph = nullptr;          // lane 0: the slice-invariant prologue as a graph (reused plans)
    cudaGraphExec_t inv_graph_exec = nullptr;
    // DAG schedule: branch b >= 1 runs on branch_stream[b]; branch_ev[b] joins it back, ev_fork releases it
    int n_branches = 1;
    cudaStream_t branch_stream[kMaxBranches] = {nullptr};
    cudaEvent_t branch_ev[kMaxBranches] = {nullptr};
    cudaEvent_t branch_ev_spare[kMaxBranches] = {nullptr};
    cudaEvent_t ev_fork = nullptr;
    std::vector<cudaEvent_t> op_ev[2];        // [list][op index]: recorded after ops other branches wait for
};

struct tob_plan {
    Program prog;
    bool uploaded = false;
    int device = 0;
    int n_lanes = 1;
    Lane lane[kMaxLanes];
    // one device block: [states | acc | results | leaf_off x lanes | slice tables | micro programs | leaves |
    //                    arena x lanes | workspace x lanes]
    void* d_block = nullptr;
    size_t d_block_size = 0;
    double* d_leaves = nullptr;
    double* d_acc = nullptr;
    double* d_results = nullptr;
    int32_t* d_term_start = nullptr;
    uint8_t *d_id_bit = nullptr, *d_addr_bit = nullptr;
    std::vector<MicroOpDev*> d_micro_ops;   // per micro stage
    std::vector<int32_t*> d_micro_start;
    // one pinned block mirroring the prefix of the device block up to the end of the leaves
    void* h_block = nullptr;
    size_t h_block_size = 0;
    DevState* h_state = nullptr;   // upload slots, one per lane
    double* h_readback = nullptr;  // result slot
    double* h_leaves = nullptr;    // pinned mirror of the leaf region (tob_plan_update_leaves refills it)
    bool has_terms = false;
    double last_ms = 0, last_issue_ms = 0;
    int64_t last_launches = 0;
    int64_t graph_launches_per_slice = 0, inv_graph_launches = 0;
    int64_t runs = 0;
    std::vector<cudaEvent_t> gemm_events;  // pairs
    std::vector<double> gemm_event_flops;
    double last_gemm_ms = 0, last_gemm_flops = 0;
    int64_t last_gemm_launches = 0;
    double slice_flops = 0, invariant_flops = 0;
    bool time_gemm = false;
    bool in_flight = false;  // a run or profile pass has been issued and not yet waited for
    // tob_plan_run_async -> tob_plan_wait
    bool pending = false;
    int pending_launches = 0;
    size_t pending_gemm = 0;
    cudaEvent_t ev_after = nullptr;  // orders an async run after the caller's stream
    double modulus = 0.0;  // exact mode: prime modulus (< 2^23), 0 = float64 arithmetic
};

// CUDA events released on every exit path of the stand-alone entry points
struct EventSet {
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    ~EventSet() {
        for (cudaEvent_t e : ev)
            if (e) cudaEventDestroy(e);
    }
};

static bool g_configured[64] = {false};  // per device: kernel attributes live in the device's context

// ------------------------------------------------------------------------------------------------
// Process-wide caches.  A call of B200API.contract_sliced creates, uploads, runs and destroys a plan;
// cudaMalloc / cudaMallocHost / cudaFree / stream + event creation cost milliseconds, more than the
// kernels of a small instance.  Blocks and streams are therefore recycled (device blocks above
// 4 GiB are returned to the driver at once; at most 8 GiB stay cached per device).
// ------------------------------------------------------------------------------------------------
namespace {
struct Block { void* ptr; size_t size; int device; bool pinned; };
struct StreamSet { cudaStream_t stream; cudaEvent_t ev0, ev1; int device; };
std::mutex g_pool_mu;
std::vector<Block> g_free_blocks;
std::vector<StreamSet> g_free_streams;
const size_t kMaxCachedBlock = (size_t)4 << 30;
const size_t kMaxCachedTotal = (size_t)8 << 30;
const size_t kMaxCachedPinned = (size_t)256 << 20;  // idle pinned host memory kept for reuse

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// cudaFree / cudaFreeHost wait for EVERYTHING in flight on the device, so with several host threads running contractions
// (the bench's e2e arm, any caller with a thread pool) a free issued under the pool lock stalled every other thread at its next
// pool call until the GPU drained — measured as 86 -> 263 ms steps.  Rules: (1) nothing is freed on a cache MISS (the old
// "drop the largest block that was too small" rule), (2) sizes are rounded up to 4 significant bits so plans of similar
// size exchange blocks instead of missing, (3) blocks over the cap are taken off the list under the lock and freed after it.
size_t size_class(size_t bytes) {
    bytes = std::max<size_t>(bytes, 4096);
    int e = 0;
    while (((size_t)1 << (e + 1)) <= bytes) e++;
    const size_t step = (size_t)1 << std::max(e - 3, 12);
    return align_up(bytes, step);
}

void free_blocks(const std::vector<Block>& victims) {
    for (const Block& b : victims) {
        if (b.pinned) cudaFreeHost(b.ptr); else cudaFree(b.ptr);
    }
}

// takes cached device blocks of `device` off the list (largest first) until at most keep_bytes stay; the caller frees them
void pool_trim_locked(int device, size_t keep_bytes, std::vector<Block>* victims) {
    size_t total = 0;
    for (const Block& b : g_free_blocks)
        if (!b.pinned && b.device == device) total += b.size;
    while (total > keep_bytes) {
        int big = -1;
        for (size_t i = 0; i < g_free_blocks.size(); i++)
            if (!g_free_blocks[i].pinned && g_free_blocks[i].device == device &&
                (big < 0 || g_free_blocks[i].size > g_free_blocks[big].size)) big = (int)i;
        if (big < 0) break;
        victims->push_back(g_free_blocks[big]);
        total -= g_free_blocks[big].size;
        g_free_blocks.erase(g_free_blocks.begin() + big);
    }
}

cudaError_t pool_acquire(size_t bytes, int device, bool pinned, Block* out) {
    bytes = size_class(bytes);
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        int best = -1;
        for (size_t i = 0; i < g_free_blocks.size(); i++) {
            const Block& b = g_free_blocks[i];
            if (b.pinned != pinned || (!pinned && b.device != device) || b.size < bytes) continue;
            if (best < 0 || b.size < g_free_blocks[best].size) best = (int)i;
        }
        // a cached block up to twice the size is taken as it is; a larger one stays for a plan that needs it
        if (best >= 0 && g_free_blocks[best].size <= 2 * bytes) {
            *out = g_free_blocks[best];
            g_free_blocks.erase(g_free_blocks.begin() + best);
            return cudaSuccess;
        }
    }
    void* ptr = nullptr;
    cudaError_t e = pinned ? cudaMallocHost(&ptr, bytes) : cudaMalloc(&ptr, bytes);
    if (e == cudaErrorMemoryAllocation && !pinned) {  // out of memory: give everything cached back, then try once more
        cudaGetLastError();
        std::vector<Block> victims;
        {
            std::lock_guard<std::mutex> lock(g_pool_mu);
            pool_trim_locked(device, 0, &victims);
        }
        free_blocks(victims);
        e = cudaMalloc(&ptr, bytes);
    }
    if (e != cudaSuccess) return e;
    *out = Block{ptr, bytes, device, pinned};
    return cudaSuccess;
}

void pool_release(const Block& b) {
    if (!b.ptr) return;
    std::vector<Block> victims;
    if ((!b.pinned && b.size > kMaxCachedBlock) || (b.pinned && b.size > kMaxCachedPinned / 4)) {
        victims.push_back(b);
    } else {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        g_free_blocks.push_back(b);
        if (!b.pinned) {
            pool_trim_locked(b.device, kMaxCachedTotal, &victims);
        } else {
            size_t pinned_total = 0;  // idle pinned blocks: oldest go first once the cap is exceeded
            for (const Block& f : g_free_blocks)
                if (f.pinned) pinned_total += f.size;
            for (size_t i = 0; i < g_free_blocks.size() && pinned_total > kMaxCachedPinned;) {
                if (!g_free_blocks[i].pinned) { i++; continue; }
                pinned_total -= g_free_blocks[i].size;
                victims.push_back(g_free_blocks[i]);
                g_free_blocks.erase(g_free_blocks.begin() + i);
            }
        }
    }
    free_blocks(victims);
}

// (Stream priorities by plan size — shortest job first across the contractions several host threads keep in flight — were
// measured and change nothing: 82.3-83.9 vs 81.9-85.9 ms per e2e bench step; a small plan's kernels cannot co-reside with two
// 96-register GEMM CTAs per SM either way.  What shortens the step is more host threads: profiles/r02i_e2e_threads.md.)
cudaError_t streams_acquire(int device, StreamSet* out) {
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        for (size_t i = 0; i < g_free_streams.size(); i++)
            if (g_free_streams[i].device == device) {
                *out = g_free_streams[i];
                g_free_streams.erase(g_free_streams.begin() + i);
                return cudaSuccess;
            }
    }
    StreamSet s{nullptr, nullptr, nullptr, device};
    cudaError_t e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&s.ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&s.ev1);
    if (e != cudaSuccess) return e;
    *out = s;
    return cudaSuccess;
}

void streams_release(const StreamSet& s) {
    if (!s.stream) return;
    std::lock_guard<std::mutex> lock(g_pool_mu);
    g_free_streams.push_back(s);
}

// ordering-only events (no timing) for the DAG schedule, recycled per device
std::vector<std::pair<int, cudaEvent_t>> g_free_events;
cudaError_t event_acquire(int device, cudaEvent_t* out) {
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        for (size_t i = g_free_events.size(); i-- > 0;)
            if (g_free_events[i].first == device) {
                *out = g_free_events[i].second;
                g_free_events.erase(g_free_events.begin() + i);
                return cudaSuccess;
            }
    }
    return cudaEventCreateWithFlags(out, cudaEventDisableTiming);
}
void event_release(int device, cudaEvent_t e) {
    if (!e) return;
    std::lock_guard<std::mutex> lock(g_pool_mu);
    g_free_events.emplace_back(device, e);
}

}  // namespace

static int ensure_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        set_error(std::string("no usable CUDA device: ") + cudaGetErrorString(e));
        return TOB_E_NODEVICE;
    }
    if (device < 0 || device >= n) {
        set_error("device ordinal out of range");
        return TOB_E_INVALID;
    }
    CUDA_TRY(cudaSetDevice(device));
    if (device < 64 && !g_configured[device]) {
        CUDA_TRY(configure_kernels());
        g_configured[device] = true;
    }
    return TOB_OK;
}

extern "C" {

void tob_default_options(tob_options* opt) {
    opt->device = 0;
    opt->use_graph = 2;
    opt->kernel_policy = 0;
    opt->hoist_invariant = 1;
    opt->mem_limit_bytes = 0;
    opt->use_microtree = 1;
    opt->slice_lanes = 0;
    opt->dag_branches = 0;
}

const char* tob_last_error(void) { return g_error.c_str(); }

int tob_tuning_set(const char* key, double value) {
    if (tuning_set(key, value)) return TOB_OK;
    set_error(std::string("unknown tuning key: ") + (key ? key : "(null)"));
    return TOB_E_INVALID;
}
double tob_gemm_time_model_us(int32_t m, int32_t n, int32_t k, int32_t tm_log2, int32_t tn_log2, int32_t c) {
    return gemm_time_model_us(m, n, k, tm_log2, tn_log2, c);
}
int tob_tuning_get(const char* key, double* value) {
    if (value && tuning_get(key, value)) return TOB_OK;
    set_error(std::string("unknown tuning key: ") + (key ? key : "(null)"));
    return TOB_E_INVALID;
}
const char* tob_version(void) { return "tob200 0.1 (sm_100a)"; }

int tob_warm(int32_t device) {
    int rc = ensure_device(device);  // context + kernel attributes
    if (rc != TOB_OK) return rc;
    CUDA_TRY(cudaFree(nullptr));
    // first-use costs that would otherwise land in the first contraction: a small pinned + device block and a stream set
    Block hb, db;
    if (pool_acquire(1 << 20, device, true, &hb) == cudaSuccess) pool_release(hb);
    if (pool_acquire(8 << 20, device, false, &db) == cudaSuccess) pool_release(db);
    StreamSet ss;
    if (streams_acquire(device, &ss) == cudaSuccess) streams_release(ss);
    cudaGetLastError();
    return TOB_OK;
}

int tob_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

// the SM count feeds the compiler's wave model and the persistent grids; queried once, without creating a context
static void query_num_sms(int device) {
    static bool done = false;
    if (done) return;
    done = true;
    int n = 0, sms = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return; }
    if (device < 0 || device >= n) device = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess) set_num_sms(sms);
    else cudaGetLastError();
}

int tob_plan_create(const tob_plan_desc* desc, const tob_options* opt, tob_plan** out) {
    if (!out) { set_error("out is NULL"); return TOB_E_INVALID; }
    *out = nullptr;
    query_num_sms(opt ? opt->device : 0);
    tob_plan* p = new tob_plan();
    std::string err;
    int rc = compile(desc, opt, &p->prog, &err);
    if (rc != TOB_OK) {
        set_error(err);
        delete p;
        return rc;
    }
    p->device = p->prog.opt.device;
    *out = p;
    return TOB_OK;
}

int64_t tob_plan_peak_bytes(const tob_plan* p) {
    size_t tables = 4096 + 32 * p->prog.leaves.size();
    for (const MicroProgram& mp : p->prog.micro) tables += mp.ops.size() * sizeof(MicroOpDev) + mp.cta_start.size() * 4 + 1024;
    for (const LeafInfo& L : p->prog.leaves) tables += 2 * L.slice_id_bit.size();
    tables += kMaxResults * 8;
    const int64_t lanes = p->prog.lanes;
    return 8 * (p->prog.leaf_doubles + lanes * (p->prog.arena_doubles + p->prog.ws_doubles)) + (int64_t)tables;
}

uint64_t tob_plan_num_slices(const tob_plan* p) { return (uint64_t)1 << p->prog.n_slice_groups; }

int64_t tob_plan_describe(const tob_plan* p, char* buf, int64_t cap) {
    std::string s = describe(p->prog);
    if (buf && cap > 0) {
        int64_t n = std::min<int64_t>(cap - 1, (int64_t)s.size());
        memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return (int64_t)s.size();
}

int64_t tob_plan_num_ops(const tob_plan* p) { return (int64_t)(p->prog.invariant_ops.size() + p->prog.slice_ops.size()); }
int tob_plan_work(const tob_plan* p, double* slice_flops, double* invariant_flops, int64_t* slice_launches) {
    if (!p) { set_error("null plan"); return TOB_E_INVALID; }
    double sf = 0, inv = 0;
    for (const Op& op : p->prog.slice_ops) sf += op.flops;
    for (const Op& op : p->prog.invariant_ops) inv += op.flops;
    if (slice_flops) *slice_flops = sf;
    if (invariant_flops) *invariant_flops = inv;
    if (slice_launches) *slice_launches = (int64_t)p->prog.slice_ops.size() + 1;
    return TOB_OK;
}

static void release_lanes(tob_plan* p) {
    for (int l = 0; l < kMaxLanes; l++) {
        Lane& L = p->lane[l];
        // a completed run has joined every branch into the lane's stream and waited for it: only a run that
        // failed half-way can have left work behind
        if (p->in_flight) {
            if (L.own_stream) cudaStreamSynchronize(L.own_stream);
            for (int b = 1; b < L.n_branches; b++)
                if (L.branch_stream[b]) cudaStreamSynchronize(L.branch_stream[b]);
        }
        if (L.graph_exec) cudaGraphExecDestroy(L.graph_exec);
        if (L.graph) cudaGraphDestroy(L.graph);
        if (L.inv_graph_exec) cudaGraphExecDestroy(L.inv_graph_exec);
        if (L.inv_graph) cudaGraphDestroy(L.inv_graph);
        if (L.own_stream) streams_release(StreamSet{L.own_stream, L.ev_a, L.ev_b, p->device});
        for (int b = 1; b < L.n_branches; b++)
            if (L.branch_stream[b]) streams_release(StreamSet{L.branch_stream[b], L.branch_ev[b], L.branch_ev_spare[b], p->device});
        event_release(p->device, L.ev_fork);
        for (int w = 0; w < 2; w++)
            for (cudaEvent_t e : L.op_ev[w]) event_release(p->device, e);
        L = Lane();
    }
}

static void release_device(tob_plan* p) {
    release_lanes(p);
    pool_release(Block{p->d_block, p->d_block_size, p->device, false});
    pool_release(Block{p->h_block, p->h_block_size, p->device, true});
    for (cudaEvent_t e : p->gemm_events) cudaEventDestroy(e);
    p->gemm_events.clear();
    event_release(p->device, p->ev_after);
    p->ev_after = nullptr;
    p->pending = false;
    p->d_block = nullptr; p->h_block = nullptr;
    p->d_term_start = nullptr; p->d_id_bit = nullptr; p->d_addr_bit = nullptr;
    p->h_state = nullptr; p->h_readback = nullptr; p->h_leaves = nullptr;
    p->d_micro_ops.clear();
    p->d_micro_start.clear();
    p->uploaded = false;
    p->runs = 0;
}

void tob_plan_destroy(tob_plan* p) {
    if (!p) return;
    if (p->uploaded || p->lane[0].own_stream) {
        cudaSetDevice(p->device);
        release_device(p);
    }
    delete p;
}

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// leaves permuted on the host into canonical order (LeafInfo::src_bit), zero-padded to the arena alignment
static void fill_leaves(const Program& G, const double* leaf_data, double* h_leaves) {
    for (const LeafInfo& Lf : G.leaves) {
        const int64_t n = (int64_t)1 << Lf.rank;
        const double* src = leaf_data + Lf.src_offset;
        double* dst = h_leaves + Lf.dev_offset;
        for (int64_t e = 0; e < n; e++) {
            int64_t sidx = 0;
            for (int q = 0; q < Lf.rank; q++) sidx |= ((e >> q) & 1) << Lf.src_bit[q];
            dst[e] = src[sidx];
        }
        for (int64_t e = n; e < (int64_t)align_up((size_t)n, 32); e++) dst[e] = 0.0;
    }
}

int tob_plan_upload(tob_plan* p, const double* leaf_data, int64_t n_doubles) {
    if (!p || !leaf_data) { set_error("NULL argument"); return TOB_E_INVALID; }
    static const bool trace = getenv("TOB_TRACE") != nullptr;
    const double t_begin = now_ms();
    Program& G = p->prog;
    if (n_doubles != G.src_leaf_len) { set_error("leaf buffer length does not match the plan"); return TOB_E_INVALID; }
    int rc = ensure_device(p->device);
    if (rc != TOB_OK) return rc;
    if (p->uploaded || p->lane[0].own_stream) release_device(p);

    const int64_t need = tob_plan_peak_bytes(p);
    if (G.opt.mem_limit_bytes > 0 && need > G.opt.mem_limit_bytes) {
        set_error("plan needs " + std::to_string(need) + " bytes, above mem_limit_bytes");
        return TOB_E_OOM;
    }
    const double t_dev = now_ms();
    p->n_lanes = G.lanes;
    for (int l = 0; l < p->n_lanes; l++) {
        StreamSet ss;
        CUDA_TRY(streams_acquire(p->device, &ss));
        p->lane[l].own_stream = ss.stream;
        p->lane[l].stream = ss.stream;
        p->lane[l].ev_a = ss.ev0;
        p->lane[l].ev_b = ss.ev1;
    }
    // DAG schedule: branch streams + one ordering event per op another branch waits for
    for (int l = 0; l < p->n_lanes; l++) {
        Lane& Ln = p->lane[l];
        Ln.n_branches = std::min<int>(std::max<int>(G.branches, 1), kMaxBranches);
        if (Ln.n_branches > 1) CUDA_TRY(event_acquire(p->device, &Ln.ev_fork));
        for (int b = 1; b < Ln.n_branches; b++) {
            StreamSet ss;
            CUDA_TRY(streams_acquire(p->device, &ss));
            Ln.branch_stream[b] = ss.stream;
            Ln.branch_ev[b] = ss.ev0;
            Ln.branch_ev_spare[b] = ss.ev1;
        }
        for (int w = 0; w < 2; w++) {
            const std::vector<Op>& list = w ? G.slice_ops : G.invariant_ops;
            if (w == 0 && l > 0) continue;  // the invariant prologue runs on lane 0 only
            Ln.op_ev[w].assign(list.size(), nullptr);
            if (Ln.n_branches > 1)
                for (size_t j = 0; j < list.size(); j++)
                    if (list[j].signal) CUDA_TRY(event_acquire(p->device, &Ln.op_ev[w][j]));
        }
    }
    p->slice_flops = 0;
    for (const Op& op : G.slice_ops) p->slice_flops += op.flops;
    p->invariant_flops = 0;
    for (const Op& op : G.invariant_ops) p->invariant_flops += op.flops;

    // ---- host-side tables ----
    const int L = (int)G.leaves.size();
    std::vector<int32_t> term_start(L + 1, 0);
    std::vector<uint8_t> id_bit, addr_bit;
    for (int l = 0; l < L; l++) {
        term_start[l] = (int32_t)id_bit.size();
        for (size_t j = 0; j < G.leaves[l].slice_id_bit.size(); j++) {
            id_bit.push_back((uint8_t)G.leaves[l].slice_id_bit[j]);
            addr_bit.push_back((uint8_t)G.leaves[l].slice_addr_bit[j]);
        }
    }
    term_start[L] = (int32_t)id_bit.size();
    p->has_terms = !id_bit.empty();

    // ---- layout of the single device block (byte offsets, 256-B aligned sections) ----
    size_t off = 0;
    auto section = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const int NL = p->n_lanes;
    const size_t o_state = section(sizeof(DevState) * kMaxLanes);
    const size_t o_acc = section(64);
    const size_t o_results = section(sizeof(double) * kMaxResults);
    size_t o_leaf_off[kMaxLanes];
    for (int l = 0; l < NL; l++) o_leaf_off[l] = section(sizeof(long long) * (L + 1));
    const size_t o_term = section(sizeof(int32_t) * (L + 1));
    const size_t o_idb = section(id_bit.size() + 1);
    const size_t o_adb = section(addr_bit.size() + 1);
    const int NM = (int)G.micro.size();
    std::vector<size_t> o_mops(NM), o_mstart(NM);
    for (int w = 0; w < NM; w++) {
        o_mops[w] = section(G.micro[w].ops.size() * sizeof(MicroOpDev) + 8);
        o_mstart[w] = section(G.micro[w].cta_start.size() * sizeof(int32_t) + 8);
    }
    size_t o_skf[kMaxLanes];
    for (int l = 0; l < NL; l++) o_skf[l] = section(kSkFlagBytes);  // inside the zero-initialised prefix
    const size_t o_leaves = section((size_t)G.leaf_doubles * 8);
    const size_t prefix_bytes = off;  // everything up to here is initialised from the pinned mirror
    size_t o_arena[kMaxLanes], o_ws[kMaxLanes];
    for (int l = 0; l < NL; l++) o_arena[l] = section((size_t)G.arena_doubles * 8);
    for (int l = 0; l < NL; l++) o_ws[l] = section((size_t)G.ws_doubles * 8);
    const size_t total_bytes = off + 256;

    const double t_tables = now_ms();
    Block db, hb;
    {
        cudaError_t e = pool_acquire(total_bytes, p->device, false, &db);
        if (e == cudaErrorMemoryAllocation) {
            cudaGetLastError();
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            set_error("plan needs " + std::to_string(total_bytes) + " bytes, device has " + std::to_string(free_b) + " free");
            release_lanes(p);
            return TOB_E_OOM;
        }
        CUDA_TRY(e);
    }
    p->d_block = db.ptr;
    p->d_block_size = db.size;
    CUDA_TRY(pool_acquire(prefix_bytes + 256, p->device, true, &hb));
    p->h_block = hb.ptr;
    p->h_block_size = hb.size;
    char* d = static_cast<char*>(p->d_block);
    char* h = static_cast<char*>(p->h_block);
    for (int l = 0; l < NL; l++) {
        p->lane[l].d_state = reinterpret_cast<DevState*>(d + o_state) + l;
        p->lane[l].d_leaf_off = reinterpret_cast<long long*>(d + o_leaf_off[l]);
        p->lane[l].d_arena = reinterpret_cast<double*>(d + o_arena[l]);
        p->lane[l].d_ws = reinterpret_cast<double*>(d + o_ws[l]);
        p->lane[l].d_sk_flags = reinterpret_cast<unsigned*>(d + o_skf[l]);
    }
    p->d_acc = reinterpret_cast<double*>(d + o_acc);
    p->d_results = reinterpret_cast<double*>(d + o_results);
    p->d_term_start = reinterpret_cast<int32_t*>(d + o_term);
    p->d_id_bit = reinterpret_cast<uint8_t*>(d + o_idb);
    p->d_addr_bit = reinterpret_cast<uint8_t*>(d + o_adb);
    p->d_micro_ops.resize(NM);
    p->d_micro_start.resize(NM);
    for (int w = 0; w < NM; w++) {
        p->d_micro_ops[w] = reinterpret_cast<MicroOpDev*>(d + o_mops[w]);
        p->d_micro_start[w] = reinterpret_cast<int32_t*>(d + o_mstart[w]);
    }
    p->d_leaves = reinterpret_cast<double*>(d + o_leaves);
    p->h_state = reinterpret_cast<DevState*>(h + o_state);
    p->h_readback = reinterpret_cast<double*>(h + prefix_bytes);

    const double t_alloc = now_ms();
    // ---- fill the pinned mirror: tables, micro programs, leaves permuted into canonical order ----
    memset(h, 0, o_leaves);
    memcpy(h + o_term, term_start.data(), sizeof(int32_t) * (L + 1));
    if (!id_bit.empty()) {
        memcpy(h + o_idb, id_bit.data(), id_bit.size());
        memcpy(h + o_adb, addr_bit.data(), addr_bit.size());
    }
    for (int w = 0; w < NM; w++) {
        const MicroProgram& mp = G.micro[w];
        MicroOpDev* mo = reinterpret_cast<MicroOpDev*>(h + o_mops[w]);
        for (size_t j = 0; j < mp.ops.size(); j++) {
            const Op& op = mp.ops[j];
            MicroOpDev m;
            memset(&m, 0, sizeof(m));
            m.a_off = op.a.offset; m.b_off = op.b.offset; m.c_off = op.c_offset;
            m.a_leaf = op.a.leaf; m.b_leaf = op.b.leaf;
            m.mask_m = (uint16_t)op.mask_m;
            m.a_space = (uint8_t)op.a.space; m.b_space = (uint8_t)op.b.space;
            m.m = (uint8_t)op.m; m.n = (uint8_t)op.n; m.k = (uint8_t)op.k;
            mo[j] = m;
        }
        // per CTA: forward each result to the next join when that join consumes it, stage small leaf
        // operands in the CTA's shared leaf cache
        for (size_t c = 0; c + 1 < mp.cta_start.size(); c++) {
            int cache_used = 0;
            for (int j = mp.cta_start[c]; j < mp.cta_start[c + 1]; j++) {
                const Op& op = mp.ops[j];
                if (j > mp.cta_start[c]) {
                    const Op& prev = mp.ops[j - 1];
                    const bool small = ((int64_t)1 << (prev.m + prev.n)) <= (mp.threads > 256 ? kMicroFwdMaxBig : kMicroFwdMax);
                    if (small && op.a.space != 0 && op.a.node == prev.node) { mo[j].a_src = 1; mo[j - 1].fwd_out = 1; }
                    if (small && op.b.space != 0 && op.b.node == prev.node) { mo[j].b_src = 1; mo[j - 1].fwd_out = 1; }
                }
                const int a_sz = 1 << (op.m + op.k), b_sz = 1 << (op.n + op.k);
                if (op.a.space == 0 && a_sz <= kMicroStageMax && cache_used + a_sz <= kMicroLeafCache) {
                    mo[j].a_src = 2; mo[j].a_soff = (uint16_t)cache_used; cache_used += std::max(a_sz, 2);
                }
                if (op.b.space == 0 && b_sz <= kMicroStageMax && cache_used + b_sz <= kMicroLeafCache) {
                    mo[j].b_src = 2; mo[j].b_soff = (uint16_t)cache_used; cache_used += std::max(b_sz, 2);
                }
            }
        }
        if (!mp.cta_start.empty()) memcpy(h + o_mstart[w], mp.cta_start.data(), mp.cta_start.size() * sizeof(int32_t));
    }
    p->h_leaves = reinterpret_cast<double*>(h + o_leaves);
    fill_leaves(G, leaf_data, p->h_leaves);
    const double t_fill = now_ms();
    // ---- ONE pinned host->device copy: state, tables, micro programs, leaves ----
    p->in_flight = true;
    CUDA_TRY(cudaMemcpyAsync(p->d_block, p->h_block, prefix_bytes, cudaMemcpyHostToDevice, p->lane[0].stream));
    CUDA_TRY(cudaStreamSynchronize(p->lane[0].stream));
    p->in_flight = false;
    if (trace)
        fprintf(stderr, "[tob] upload: ensure_device %.3f  streams+tables %.3f  alloc %.3f  fill %.3f  h2d+sync %.3f ms (%zu B)\n",
                t_dev - t_begin, t_tables - t_dev, t_alloc - t_tables, t_fill - t_alloc, now_ms() - t_fill, prefix_bytes);
    p->uploaded = true;
    return TOB_OK;
}

int tob_plan_update_leaves(tob_plan* p, const double* leaf_data, int64_t n_doubles) {
    if (!p || !leaf_data) { set_error("NULL argument"); return TOB_E_INVALID; }
    if (!p->uploaded) return tob_plan_upload(p, leaf_data, n_doubles);
    if (n_doubles != p->prog.src_leaf_len) { set_error("leaf buffer length does not match the plan"); return TOB_E_INVALID; }
    int rc = ensure_device(p->device);
    if (rc != TOB_OK) return rc;
    fill_leaves(p->prog, leaf_data, p->h_leaves);
    p->in_flight = true;
    CUDA_TRY(cudaMemcpyAsync(p->d_leaves, p->h_leaves, (size_t)p->prog.leaf_doubles * 8, cudaMemcpyHostToDevice, p->lane[0].stream));
    CUDA_TRY(cudaStreamSynchronize(p->lane[0].stream));
    p->in_flight = false;
    return TOB_OK;
}

int tob_plan_release(tob_plan* p) {
    if (!p) { set_error("NULL argument"); return TOB_E_INVALID; }
    if (p->uploaded || p->lane[0].own_stream) {
        cudaSetDevice(p->device);
        release_device(p);
    }
    return TOB_OK;
}

int tob_pool_trim(int32_t device) {
    std::vector<Block> victims;
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        for (size_t i = g_free_blocks.size(); i-- > 0;) {
            const Block b = g_free_blocks[i];
            if (device >= 0 && !b.pinned && b.device != device) continue;
            victims.push_back(b);
            g_free_blocks.erase(g_free_blocks.begin() + i);
        }
    }
    for (const Block& b : victims) {
        if (!b.pinned) cudaSetDevice(b.device);
        if (b.pinned) cudaFreeHost(b.ptr); else cudaFree(b.ptr);
    }
    return TOB_OK;
}

static const double* operand_ptr(const tob_plan* p, const Lane& L, const OperandRef& r) {
    // space 0: leaf region; 1: this lane's arena; 2: slice-invariant tensor, always in lane 0's arena
    const double* base = r.space == 0 ? p->d_leaves : (r.space == 2 ? p->lane[0].d_arena : L.d_arena);
    return base + r.offset;
}

static KParams make_params(const tob_plan* p, const Lane& L, const Op& op) {
    KParams k;
    memset(&k, 0, sizeof(k));
    k.a = operand_ptr(p, L, op.a);
    k.b = operand_ptr(p, L, op.b);
    k.c = L.d_arena + op.c_offset;
    k.ws = L.d_ws;
    k.sk_flags = L.d_sk_flags;
    k.streamk = op.streamk;
    k.leaf_off = L.d_leaf_off;
    k.a_leaf = op.a.leaf;
    k.b_leaf = op.b.leaf;
    k.m = op.m;
    k.n = op.n;
    k.k = op.k;
    k.ksplit_log2 = op.ksplit_log2;
    k.mask_m = op.mask_m;
    const int tot = op.m + op.n;
    k.mask_n = ~op.mask_m & (tot >= 64 ? ~0ull : ((1ull << tot) - 1ull));
    k.runs_m = make_runs(k.mask_m);
    k.runs_n = make_runs(k.mask_n);
    k.modp = p->modulus;
    k.inv_modp = p->modulus > 0.0 ? 1.0 / p->modulus : 0.0;
    return k;
}

static cudaError_t launch_op(tob_plan* p, const Lane& L, const Op& op, int* launches, cudaStream_t stream = nullptr) {
    if (!stream) stream = L.stream;
    if (op.kind == OP_ACCUM) {
        (*launches)++;
        return launch_accum(L.d_state, operand_ptr(p, L, op.a), L.d_leaf_off, op.a.leaf, p->d_results, stream);
    }
    if (op.kind == OP_MICRO) {
        const int w = op.micro_which;
        (*launches)++;
        const MicroProgram& mp = p->prog.micro[w];
        int max_ops = 1;
        for (size_t c = 0; c + 1 < mp.cta_start.size(); c++) max_ops = std::max(max_ops, mp.cta_start[c + 1] - mp.cta_start[c]);
        const int smem_ops = std::min(max_ops, (int)(kMicroDescBytes / sizeof(MicroOpDev)));
        return launch_microtree(p->d_micro_ops[w], p->d_micro_start[w], (int)mp.cta_start.size() - 1, smem_ops, mp.threads,
                                p->d_leaves, L.d_arena, p->lane[0].d_arena, L.d_leaf_off, p->modulus, stream);
    }
    KParams k = make_params(p, L, op);
    return launch_contract(op, k, stream, launches);
}

// Issues one op list along its DAG schedule (Op::branch / waits / signal, tob_compile.cpp).  Works the same
// under stream capture: the event records and waits become the graph's dependency edges.
static cudaError_t run_list(tob_plan* p, Lane& L, const std::vector<Op>& list, int which,
                            const std::function<cudaError_t(const Op&, cudaStream_t)>& launch) {
    cudaError_t e = cudaSuccess;
    const bool multi = L.n_branches > 1;
    bool forked[kMaxBranches] = {false}, dirty[kMaxBranches] = {false};
    auto join_all = [&]() -> cudaError_t {
        for (int b = 1; b < L.n_branches; b++) {
            if (!dirty[b]) { forked[b] = false; continue; }
            cudaError_t e2 = cudaEventRecord(L.branch_ev[b], L.branch_stream[b]);
            if (e2 == cudaSuccess) e2 = cudaStreamWaitEvent(L.stream, L.branch_ev[b], 0);
            if (e2 != cudaSuccess) return e2;
            dirty[b] = forked[b] = false;
        }
        return cudaSuccess;
    };
    if (multi && (e = cudaEventRecord(L.ev_fork, L.stream)) != cudaSuccess) return e;
    for (size_t j = 0; j < list.size(); j++) {
        const Op& op = list[j];
        if (op.kind != OP_GENERIC && op.kind != OP_GEMM) {  // barrier ops run on the lane's stream after a full join
            if ((e = join_all()) != cudaSuccess) return e;
            if ((e = launch(op, L.stream)) != cudaSuccess) return e;
            if (multi && (e = cudaEventRecord(L.ev_fork, L.stream)) != cudaSuccess) return e;
            continue;
        }
        const int b = multi ? std::min<int>(op.branch, L.n_branches - 1) : 0;
        cudaStream_t s = b ? L.branch_stream[b] : L.stream;
        if (b && !forked[b]) {
            if ((e = cudaStreamWaitEvent(s, L.ev_fork, 0)) != cudaSuccess) return e;
            forked[b] = true;
        }
        if (multi)
            for (int32_t w : op.waits)
                if ((e = cudaStreamWaitEvent(s, L.op_ev[which][w], 0)) != cudaSuccess) return e;
        if ((e = launch(op, s)) != cudaSuccess) return e;
        if (multi && op.signal && (e = cudaEventRecord(L.op_ev[which][j], s)) != cudaSuccess) return e;
        if (b) dirty[b] = true;
    }
    return join_all();
}

static cudaError_t launch_begin(tob_plan* p, const Lane& L, int* launches) {
    if (!p->has_terms) return cudaSuccess;
    SliceTables t{p->d_term_start, p->d_id_bit, p->d_addr_bit, L.d_leaf_off, (int32_t)p->prog.leaves.size()};
    (*launches)++;
    return launch_begin_slice(L.d_state, t, L.stream);
}

static cudaError_t launch_slice(tob_plan* p, Lane& L, int* launches) {
    cudaError_t e = launch_begin(p, L, launches);
    if (e != cudaSuccess) return e;
    return run_list(p, L, p->prog.slice_ops, 1,
                    [&](const Op& op, cudaStream_t s) { return launch_op(p, L, op, launches, s); });
}

int tob_plan_run(tob_plan* p, uint64_t first, uint64_t count, uint64_t stride, double* result) {
    return tob_plan_run_ex(p, first, count, stride, 0.0, 0, result);
}

static int run_issue(tob_plan* p, uint64_t first, uint64_t count, uint64_t stride, double initial, int32_t flags,
                     bool async, cudaStream_t after);
static int run_finish(tob_plan* p, double* result);

int tob_plan_run_ex(tob_plan* p, uint64_t first, uint64_t count, uint64_t stride, double initial, int32_t flags,
                    double* result) {
    if (!p || !result) { set_error("NULL argument"); return TOB_E_INVALID; }
    int rc = run_issue(p, first, count, stride, initial, flags, false, nullptr);
    if (rc != TOB_OK) return rc;
    return run_finish(p, result);
}

int tob_plan_run_async(tob_plan* p, uint64_t first, uint64_t count, uint64_t stride, double initial, int32_t flags,
                       void* after_stream) {
    if (!p) { set_error("NULL argument"); return TOB_E_INVALID; }
    if (count > (uint64_t)kMaxResults) {
        set_error("tob_plan_run_async: at most " + std::to_string(kMaxResults) + " slices per call");
        return TOB_E_INVALID;
    }
    return run_issue(p, first, count, stride, initial, flags, true, (cudaStream_t)after_stream);
}

int tob_plan_join(tob_plan* p, void* stream) {
    if (!p || !p->pending) { set_error("tob_plan_join: no run in flight"); return TOB_E_INVALID; }
    CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, p->lane[0].ev_b, 0));
    return TOB_OK;
}

int tob_plan_wait(tob_plan* p, double* result) {
    if (!p || !result) { set_error("NULL argument"); return TOB_E_INVALID; }
    if (!p->pending) { set_error("tob_plan_wait: no run in flight"); return TOB_E_INVALID; }
    return run_finish(p, result);
}

static int run_issue(tob_plan* p, uint64_t first, uint64_t count, uint64_t stride, double initial, int32_t flags,
                     bool async, cudaStream_t after) {
    if (!p->uploaded) { set_error("tob_plan_upload has not been called"); return TOB_E_INVALID; }
    if (p->pending) { set_error("a run is already in flight on this plan (tob_plan_wait first)"); return TOB_E_INVALID; }
    const uint64_t nslices = tob_plan_num_slices(p);
    if (count > 0 && (first >= nslices || first + (count - 1) * stride >= nslices)) {
        set_error("slice range exceeds the number of slices");
        return TOB_E_INVALID;
    }
    int rc = ensure_device(p->device);
    if (rc != TOB_OK) return rc;
    int launches = 0;
    Lane& L0 = p->lane[0];
    const int ug = p->prog.opt.use_graph;
    // auto: launch-bound slices replay as a graph, but only once the plan is being reused (second run or
    // several slices): a plan that runs a single slice once never pays capture + instantiate
    const bool as_graph = (ug == 1) || (ug == 2 && p->slice_flops < 2e9 && (p->runs > 0 || count >= 4));
    const bool skip_invariant = (flags & TOB_RUN_SKIP_INVARIANT) != 0 && p->runs > 0;
    size_t n_gemm = 0;
    cudaEvent_t last_gemm_end = nullptr;
    cudaStream_t last_gemm_stream = nullptr;
    // plain stream launches: bracket every DMMA GEMM with CUDA events (per-kernel roofline, bench.py).
    // GEMMs on different streams (lanes, DAG branches) are chained (a GEMM waits for the previous GEMM):
    // two tensor-pipe-bound kernels gain nothing from sharing the SMs, and each event pair then brackets
    // one GEMM running alone with only small kernels of other streams beside it.
    auto timed_op = [&](const Lane& L, const Op& op, cudaStream_t s) -> cudaError_t {
        if (op.kind != OP_GEMM || !p->time_gemm) return launch_op(p, L, op, &launches, s);
        cudaError_t e = cudaSuccess;
        if (p->gemm_events.size() < 2 * (n_gemm + 1)) {
            cudaEvent_t a, b;
            if ((e = cudaEventCreate(&a)) != cudaSuccess) return e;
            if ((e = cudaEventCreate(&b)) != cudaSuccess) return e;
            p->gemm_events.push_back(a);
            p->gemm_events.push_back(b);
            p->gemm_event_flops.push_back(0.0);
        }
        if (last_gemm_end && last_gemm_stream != s && (e = cudaStreamWaitEvent(s, last_gemm_end, 0)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(p->gemm_events[2 * n_gemm], s)) != cudaSuccess) return e;
        if ((e = launch_op(p, L, op, &launches, s)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(p->gemm_events[2 * n_gemm + 1], s)) != cudaSuccess) return e;
        last_gemm_end = p->gemm_events[2 * n_gemm + 1];
        last_gemm_stream = s;
        p->gemm_event_flops[n_gemm] = op.flops;
        n_gemm++;
        return cudaSuccess;
    };
    // the slice-invariant prologue replays as a graph too once the plan is reused and it is launch-bound
    // (its GEMMs then go untimed, so not while per-GEMM timing wants them)
    bool inv_has_gemm = false;
    for (const Op& op : p->prog.invariant_ops) inv_has_gemm |= (op.kind == OP_GEMM);
    const bool inv_as_graph = (ug == 1 || (ug == 2 && p->invariant_flops < 2e9 && p->runs > 0)) &&
                              !(p->time_gemm && inv_has_gemm) && !p->prog.invariant_ops.empty();

    const double t_issue0 = now_ms();
    p->in_flight = true;
    if (after && after != L0.stream) {  // start behind everything already enqueued on the caller's stream
        if (!p->ev_after) CUDA_TRY(event_acquire(p->device, &p->ev_after));
        CUDA_TRY(cudaEventRecord(p->ev_after, after));
        CUDA_TRY(cudaStreamWaitEvent(L0.stream, p->ev_after, 0));
    }
    CUDA_TRY(cudaEventRecord(L0.ev_a, L0.stream));
    uint64_t done = 0;
    bool first_batch = true;
    do {
        const uint64_t batch = std::min<uint64_t>(count - done, kMaxResults);
        if (!first_batch) CUDA_TRY(cudaStreamSynchronize(L0.stream));  // the pinned state slots are about to be rewritten
        const int lanes = (int)std::min<uint64_t>((uint64_t)p->n_lanes, std::max<uint64_t>(batch, 1));
        for (int l = 0; l < lanes; l++) {
            p->h_state[l].next_slice = first + (done + l) * stride;
            p->h_state[l].stride = stride * lanes;
            p->h_state[l].slot = l;
            p->h_state[l].slot_stride = lanes;
        }
        CUDA_TRY(cudaMemcpyAsync(L0.d_state, p->h_state, sizeof(DevState) * lanes, cudaMemcpyHostToDevice, L0.stream));
        if (first_batch && batch > 0 && !skip_invariant) {
            if (inv_as_graph) {
                if (!L0.inv_graph_exec) {
                    int n_inv = 0;
                    CUDA_TRY(cudaStreamBeginCapture(L0.stream, cudaStreamCaptureModeThreadLocal));
                    cudaError_t e = run_list(p, L0, p->prog.invariant_ops, 0,
                                             [&](const Op& op, cudaStream_t s) { return launch_op(p, L0, op, &n_inv, s); });
                    cudaError_t e2 = cudaStreamEndCapture(L0.stream, &L0.inv_graph);
                    CUDA_TRY(e);
                    CUDA_TRY(e2);
                    CUDA_TRY(cudaGraphInstantiate(&L0.inv_graph_exec, L0.inv_graph, 0));
                    p->inv_graph_launches = n_inv;
                }
                CUDA_TRY(cudaGraphLaunch(L0.inv_graph_exec, L0.stream));
                launches += (int)p->inv_graph_launches;
            } else {
                CUDA_TRY(run_list(p, L0, p->prog.invariant_ops, 0,
                                  [&](const Op& op, cudaStream_t s) { return timed_op(L0, op, s); }));
            }
        }
        // fork: the other lanes start after the state upload and the slice-invariant prologue
        for (int l = 1; l < lanes; l++) {
            CUDA_TRY(cudaEventRecord(p->lane[l].ev_a, L0.stream));
            CUDA_TRY(cudaStreamWaitEvent(p->lane[l].stream, p->lane[l].ev_a, 0));
        }
        if (as_graph) {
            for (int l = 0; l < lanes; l++) {
                Lane& L = p->lane[l];
                if (L.graph_exec) continue;
                int per_slice = 0;
                CUDA_TRY(cudaStreamBeginCapture(L.stream, cudaStreamCaptureModeThreadLocal));
                cudaError_t e = launch_slice(p, L, &per_slice);
                cudaError_t e2 = cudaStreamEndCapture(L.stream, &L.graph);
                CUDA_TRY(e);
                CUDA_TRY(e2);
                CUDA_TRY(cudaGraphInstantiate(&L.graph_exec, L.graph, 0));
                p->graph_launches_per_slice = per_slice;
            }
            for (uint64_t j = 0; j < batch; j++) CUDA_TRY(cudaGraphLaunch(p->lane[j % lanes].graph_exec, p->lane[j % lanes].stream));
            launches += (int)(p->graph_launches_per_slice * batch);
        } else {
            for (uint64_t j = 0; j < batch; j++) {
                Lane& L = p->lane[j % lanes];
                CUDA_TRY(launch_begin(p, L, &launches));
                CUDA_TRY(run_list(p, L, p->prog.slice_ops, 1,
                                  [&](const Op& op, cudaStream_t s) { return timed_op(L, op, s); }));
            }
        }
        // join, then the ordered sum of this batch's per-slice results (sequential, slice order)
        for (int l = 1; l < lanes; l++) {
            CUDA_TRY(cudaEventRecord(p->lane[l].ev_b, p->lane[l].stream));
            CUDA_TRY(cudaStreamWaitEvent(L0.stream, p->lane[l].ev_b, 0));
        }
        CUDA_TRY(launch_final_sum(p->d_acc, p->d_results, (int)batch, initial, first_batch ? 0 : 1, p->modulus, L0.stream));
        launches++;
        done += batch;
        first_batch = false;
    } while (done < count);
    CUDA_TRY(cudaEventRecord(L0.ev_b, L0.stream));
    CUDA_TRY(cudaMemcpyAsync(p->h_readback, p->d_acc, sizeof(double), cudaMemcpyDeviceToHost, L0.stream));
    p->last_issue_ms = now_ms() - t_issue0;  // host time spent issuing the run (everything before the final wait)
    p->pending = true;
    p->pending_launches = launches;
    p->pending_gemm = n_gemm;
    (void)async;
    return TOB_OK;
}

static int run_finish(tob_plan* p, double* result) {
    Lane& L0 = p->lane[0];
    p->pending = false;
    CUDA_TRY(cudaStreamSynchronize(L0.stream));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, L0.ev_a, L0.ev_b));
    p->last_ms = ms;
    p->last_launches = p->pending_launches;
    p->last_gemm_ms = 0;
    p->last_gemm_flops = 0;
    const size_t n_gemm = p->pending_gemm;
    for (size_t i = 0; i < n_gemm; i++) {
        float g = 0;
        CUDA_TRY(cudaEventElapsedTime(&g, p->gemm_events[2 * i], p->gemm_events[2 * i + 1]));
        p->last_gemm_ms += g;
        p->last_gemm_flops += p->gemm_event_flops[i];
    }
    p->last_gemm_launches = (int64_t)n_gemm;
    *result = *p->h_readback;
    p->runs++;
    p->in_flight = false;
    return TOB_OK;
}

int tob_plan_set_modulus(tob_plan* p, double modulus) {
    if (!p) { set_error("NULL argument"); return TOB_E_INVALID; }
    if (modulus != 0.0 && !(modulus >= 2.0 && modulus < 8388608.0 && modulus == floor(modulus))) {
        set_error("modulus must be 0 or an integer in [2, 2^23)");
        return TOB_E_INVALID;
    }
    if (modulus != p->modulus) {  // captured graphs carry the modulus in their kernel parameters
        for (int l = 0; l < kMaxLanes; l++) {
            Lane& L = p->lane[l];
            if (L.graph_exec) { cudaGraphExecDestroy(L.graph_exec); L.graph_exec = nullptr; }
            if (L.graph) { cudaGraphDestroy(L.graph); L.graph = nullptr; }
            if (L.inv_graph_exec) { cudaGraphExecDestroy(L.inv_graph_exec); L.inv_graph_exec = nullptr; }
            if (L.inv_graph) { cudaGraphDestroy(L.inv_graph); L.inv_graph = nullptr; }
        }
    }
    p->modulus = modulus;
    return TOB_OK;
}

int tob_plan_set_gemm_timing(tob_plan* p, int32_t on) {
    if (!p) { set_error("NULL argument"); return TOB_E_INVALID; }
    p->time_gemm = on != 0;
    return TOB_OK;
}

int tob_plan_last_gemm(const tob_plan* p, double* ms, double* flops, int64_t* launches) {
    if (!p) { set_error("NULL argument"); return TOB_E_INVALID; }
    if (ms) *ms = p->last_gemm_ms;
    if (flops) *flops = p->last_gemm_flops;
    if (launches) *launches = p->last_gemm_launches;
    return TOB_OK;
}

int tob_plan_set_stream(tob_plan* p, void* stream) {
    if (!p || !p->uploaded) { set_error("tob_plan_set_stream: plan is not uploaded"); return TOB_E_INVALID; }
    Lane& L0 = p->lane[0];
    if (L0.graph_exec) { cudaGraphExecDestroy(L0.graph_exec); L0.graph_exec = nullptr; }
    if (L0.graph) { cudaGraphDestroy(L0.graph); L0.graph = nullptr; }
    if (L0.inv_graph_exec) { cudaGraphExecDestroy(L0.inv_graph_exec); L0.inv_graph_exec = nullptr; }
    if (L0.inv_graph) { cudaGraphDestroy(L0.inv_graph); L0.inv_graph = nullptr; }
    L0.stream = stream ? (cudaStream_t)stream : L0.own_stream;
    return TOB_OK;
}

double tob_plan_last_ms(const tob_plan* p) { return p->last_ms; }
double tob_plan_last_issue_ms(const tob_plan* p) { return p->last_issue_ms; }
int64_t tob_plan_last_launches(const tob_plan* p) { return p->last_launches; }

int tob_plan_profile(tob_plan* p, uint64_t slice, float* ms_per_op, int64_t n_ops, double* result) {
    if (!p || !ms_per_op) { set_error("NULL argument"); return TOB_E_INVALID; }
    if (!p->uploaded) { set_error("tob_plan_upload has not been called"); return TOB_E_INVALID; }
    if (n_ops != tob_plan_num_ops(p)) { set_error("n_ops mismatch"); return TOB_E_INVALID; }
    if (slice >= tob_plan_num_slices(p)) { set_error("slice out of range"); return TOB_E_INVALID; }
    int rc = ensure_device(p->device);
    if (rc != TOB_OK) return rc;
    int launches = 0;
    Lane& L0 = p->lane[0];
    p->in_flight = true;
    p->h_state[0].next_slice = slice;
    p->h_state[0].stride = 1;
    p->h_state[0].slot = 0;
    p->h_state[0].slot_stride = 1;
    CUDA_TRY(cudaMemcpyAsync(L0.d_state, p->h_state, sizeof(DevState), cudaMemcpyHostToDevice, L0.stream));
    std::vector<cudaEvent_t> ev(n_ops + 1);
    for (auto& e : ev) CUDA_TRY(cudaEventCreate(&e));
    int64_t i = 0;
    for (const Op& op : p->prog.invariant_ops) {
        CUDA_TRY(cudaEventRecord(ev[i], L0.stream));
        CUDA_TRY(launch_op(p, L0, op, &launches));
        i++;
    }
    CUDA_TRY(launch_begin(p, L0, &launches));
    for (const Op& op : p->prog.slice_ops) {
        CUDA_TRY(cudaEventRecord(ev[i], L0.stream));
        CUDA_TRY(launch_op(p, L0, op, &launches));
        i++;
    }
    CUDA_TRY(cudaEventRecord(ev[i], L0.stream));
    CUDA_TRY(launch_final_sum(p->d_acc, p->d_results, 1, 0.0, 0, p->modulus, L0.stream));
    CUDA_TRY(cudaMemcpyAsync(p->h_readback, p->d_acc, sizeof(double), cudaMemcpyDeviceToHost, L0.stream));
    CUDA_TRY(cudaStreamSynchronize(L0.stream));
    for (int64_t j = 0; j < n_ops; j++) CUDA_TRY(cudaEventElapsedTime(&ms_per_op[j], ev[j], ev[j + 1]));
    for (auto& e : ev) cudaEventDestroy(e);
    if (result) *result = *p->h_readback;
    p->in_flight = false;
    return TOB_OK;
}

// ------------------------------------------------------------------------------------------------
// per-op verification (tools/verify_ops.py, tests): run a prefix of one slice's program sequentially on lane 0,
// then read tensors back, so every join's output can be held against the numpy interpreter of the same program
// ------------------------------------------------------------------------------------------------
int tob_plan_debug_run(tob_plan* p, uint64_t slice, int64_t n_ops) {
    if (!p || !p->uploaded) { set_error("tob_plan_debug_run: plan is not uploaded"); return TOB_E_INVALID; }
    if (slice >= tob_plan_num_slices(p) || n_ops < 0 || n_ops > tob_plan_num_ops(p)) { set_error("bad debug arguments"); return TOB_E_INVALID; }
    int rc = ensure_device(p->device);
    if (rc != TOB_OK) return rc;
    int launches = 0;
    Lane& L0 = p->lane[0];
    p->in_flight = true;
    p->h_state[0].next_slice = slice;
    p->h_state[0].stride = 1;
    p->h_state[0].slot = 0;
    p->h_state[0].slot_stride = 1;
    CUDA_TRY(cudaMemcpyAsync(L0.d_state, p->h_state, sizeof(DevState), cudaMemcpyHostToDevice, L0.stream));
    int64_t i = 0;
    for (const Op& op : p->prog.invariant_ops) {
        if (i++ >= n_ops) break;
        CUDA_TRY(launch_op(p, L0, op, &launches));
    }
    CUDA_TRY(launch_begin(p, L0, &launches));
    for (const Op& op : p->prog.slice_ops) {
        if (i++ >= n_ops) break;
        CUDA_TRY(launch_op(p, L0, op, &launches));
    }
    CUDA_TRY(cudaStreamSynchronize(L0.stream));
    p->in_flight = false;
    return TOB_OK;
}

int tob_plan_debug_read(tob_plan* p, int32_t space, int64_t offset, int64_t n, double* out) {
    if (!p || !p->uploaded || !out || offset < 0 || n < 0) { set_error("bad debug arguments"); return TOB_E_INVALID; }
    const int64_t limit = space == 0 ? p->prog.leaf_doubles : p->prog.arena_doubles;
    if (offset + n > limit) { set_error("tob_plan_debug_read: range outside the region"); return TOB_E_INVALID; }
    int rc = ensure_device(p->device);
    if (rc != TOB_OK) return rc;
    const double* base = space == 0 ? p->d_leaves : p->lane[0].d_arena;
    CUDA_TRY(cudaMemcpy(out, base + offset, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return TOB_OK;
}

// ------------------------------------------------------------------------------------------------
// stand-alone tensordot / permute
// ------------------------------------------------------------------------------------------------
int tob_permute_device(const double* in, double* out, int32_t rank, const int32_t* perm, void* stream_v, float* ms) {
    if (!in || !out || rank < 0 || rank > 40 || (rank > 0 && !perm)) { set_error("bad permute arguments"); return TOB_E_INVALID; }
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    int rc = ensure_device(dev);
    if (rc != TOB_OK) return rc;
    cudaStream_t stream = (cudaStream_t)stream_v;
    // numpy: out axis j = in axis perm[j]  =>  out address bit (r-1-j) <- in address bit (r-1-perm[j])
    int32_t src_bit[64];
    std::vector<char> seen(rank, 0);
    for (int j = 0; j < rank; j++) {
        if (perm[j] < 0 || perm[j] >= rank || seen[perm[j]]) { set_error("perm is not a permutation"); return TOB_E_INVALID; }
        seen[perm[j]] = 1;
        src_bit[rank - 1 - j] = rank - 1 - perm[j];
    }
    EventSet es;
    if (ms) {
        CUDA_TRY(cudaEventCreate(&es.ev[0]));
        CUDA_TRY(cudaEventCreate(&es.ev[1]));
        CUDA_TRY(cudaEventRecord(es.ev[0], stream));
    }
    CUDA_TRY(launch_permute(in, out, rank, src_bit, stream));
    if (ms) {
        CUDA_TRY(cudaEventRecord(es.ev[1], stream));
        CUDA_TRY(cudaEventSynchronize(es.ev[1]));
        CUDA_TRY(cudaEventElapsedTime(ms, es.ev[0], es.ev[1]));
    }
    return TOB_OK;
}

int tob_tensordot_device(const double* a, int32_t rank_a, const double* b, int32_t rank_b, const int32_t* axes_a,
                         const int32_t* axes_b, int32_t n_axes, double* c, double* workspace, int64_t workspace_bytes,
                         int32_t kernel_policy, void* stream_v, float* ms) {
    if (!a || !b || !c || rank_a < 0 || rank_b < 0 || rank_a > 40 || rank_b > 40 || n_axes < 0 || n_axes > rank_a ||
        n_axes > rank_b) {
        set_error("bad tensordot arguments");
        return TOB_E_INVALID;
    }
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    int rc = ensure_device(dev);
    if (rc != TOB_OK) return rc;
    cudaStream_t stream = (cudaStream_t)stream_v;
    const int k = n_axes, m = rank_a - k, n = rank_b - k;
    // canonical operand = transpose(free axes in order ++ contracted axes in pair order), C-ordered
    auto canonical_perm = [&](int rank, const int32_t* axes, std::vector<int32_t>* perm) -> bool {
        std::vector<char> is_k(rank, 0);
        for (int i = 0; i < k; i++) {
            if (axes[i] < 0 || axes[i] >= rank || is_k[axes[i]]) return false;
            is_k[axes[i]] = 1;
        }
        for (int j = 0; j < rank; j++)
            if (!is_k[j]) perm->push_back(j);
        for (int i = 0; i < k; i++) perm->push_back(axes[i]);
        return true;
    };
    std::vector<int32_t> perm_a, perm_b;
    if (!canonical_perm(rank_a, axes_a, &perm_a) || !canonical_perm(rank_b, axes_b, &perm_b)) {
        set_error("axes are not valid");
        return TOB_E_INVALID;
    }
    auto is_identity = [](const std::vector<int32_t>& v) {
        for (size_t i = 0; i < v.size(); i++)
            if (v[i] != (int32_t)i) return false;
        return true;
    };
    const bool pa = !is_identity(perm_a), pb = !is_identity(perm_b);
    Op op;
    op.m = m; op.n = n; op.k = k;
    op.mask_m = (m >= 64 ? ~0ull : ((1ull << m) - 1ull)) << n;  // numpy output order: a's free axes are the high bits
    const double* A = a;
    const double* B = b;
    bool swapped = false;
    // workspace carve-up: [permuted a][permuted b][split-K partials]
    int64_t need = 0;
    int64_t off_a = 0, off_b = 0, off_ws = 0;
    if (pa) { off_a = need; need += (int64_t)8 << rank_a; }
    if (pb) { off_b = need; need += (int64_t)8 << rank_b; }
    if (n > m) {  // keep the larger free side as M
        swapped = true;
        const int tot0 = m + n;
        const uint64_t full = tot0 >= 64 ? ~0ull : ((1ull << tot0) - 1ull);
        std::swap(op.m, op.n);
        op.mask_m = full & ~op.mask_m;
    }
    const int64_t need_perm = need;   // what the operand permutations alone require
    need = (need + 255) & ~(int64_t)255;
    const int64_t off_flags = need;  // stream-K counter + flags, zeroed below when the join runs on that kernel
    need += kSkFlagBytes;
    off_ws = need;
    const int64_t ws_avail = (workspace ? workspace_bytes : 0) - need;
    choose_kernel(&op, kernel_policy, true);
    if (op.streamk > 0 && (int64_t)op.streamk * kSkSlotDoubles * 8 > ws_avail) op.streamk = 0;  // no room for the partial tiles: one tile per CTA
    while (op.ksplit_log2 > 0 && ((int64_t)8 << (op.m + op.n + op.ksplit_log2)) > ws_avail) op.ksplit_log2--;
    if ((pa || pb) && (!workspace || workspace_bytes < need_perm)) {
        set_error("tensordot needs a workspace of at least " + std::to_string(need_perm) + " bytes for the operand permutations");
        return TOB_E_INVALID;
    }
    EventSet es;
    cudaEvent_t* ev = es.ev;
    if (ms) {
        for (int i = 0; i < 3; i++) CUDA_TRY(cudaEventCreate(&ev[i]));
        CUDA_TRY(cudaEventRecord(ev[0], stream));
    }
    auto run_permute = [&](const double* src, double* dst, int rank, const std::vector<int32_t>& perm) -> cudaError_t {
        int32_t src_bit[64];
        for (int j = 0; j < rank; j++) src_bit[rank - 1 - j] = rank - 1 - perm[j];
        return launch_permute(src, dst, rank, src_bit, stream);
    };
    if (pa) {
        double* t = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + off_a);
        CUDA_TRY(run_permute(a, t, rank_a, perm_a));
        A = t;
    }
    if (pb) {
        double* t = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + off_b);
        CUDA_TRY(run_permute(b, t, rank_b, perm_b));
        B = t;
    }
    if (ms) CUDA_TRY(cudaEventRecord(ev[1], stream));
    KParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.a = swapped ? B : A;
    kp.b = swapped ? A : B;
    kp.c = c;
    kp.ws = workspace ? reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + off_ws) : nullptr;
    kp.leaf_off = nullptr;
    kp.a_leaf = kp.b_leaf = -1;
    kp.m = op.m; kp.n = op.n; kp.k = op.k;
    kp.ksplit_log2 = op.ksplit_log2;
    kp.streamk = op.streamk;
    if (op.streamk > 0) {
        kp.sk_flags = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(workspace) + off_flags);
        CUDA_TRY(cudaMemsetAsync(kp.sk_flags, 0, kSkFlagBytes, stream));
    }
    kp.mask_m = op.mask_m;
    const int tot = op.m + op.n;
    kp.mask_n = tot == 0 ? 0ull : (~op.mask_m & (tot >= 64 ? ~0ull : ((1ull << tot) - 1ull)));
    kp.runs_m = make_runs(kp.mask_m);
    kp.runs_n = make_runs(kp.mask_n);
    kp.modp = 0.0;
    kp.inv_modp = 0.0;
    int launches = 0;
    CUDA_TRY(launch_contract(op, kp, stream, &launches));
    if (ms) {
        CUDA_TRY(cudaEventRecord(ev[2], stream));
        CUDA_TRY(cudaEventSynchronize(ev[2]));
        CUDA_TRY(cudaEventElapsedTime(&ms[0], ev[0], ev[1]));  // permutations
        CUDA_TRY(cudaEventElapsedTime(&ms[1], ev[1], ev[2]));  // contraction
        ms[2] = (float)op.kind;
    }
    return TOB_OK;
}

int tob_tensordot_host(const double* a, int32_t rank_a, const double* b, int32_t rank_b, const int32_t* axes_a,
                       const int32_t* axes_b, int32_t n_axes, double* c) {
    if (rank_a < 0 || rank_b < 0 || rank_a > 32 || rank_b > 32 || n_axes < 0 || n_axes > rank_a || n_axes > rank_b) {
        set_error("bad tensordot arguments");
        return TOB_E_INVALID;
    }
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    int rc = ensure_device(dev);
    if (rc != TOB_OK) return rc;
    const int rank_c = rank_a + rank_b - 2 * n_axes;
    const size_t na = (size_t)8 << rank_a, nb = (size_t)8 << rank_b, nc = (size_t)8 << rank_c;
    const size_t nws = na + nb + (nc << 4) + 256;
    // ONE block from the pool (operands, result, workspace): cudaMalloc / cudaFree per call cost milliseconds, and cudaFree
    // waits for everything else in flight on the device (other host threads' contractions)
    const size_t oa = 0, ob = align_up(oa + na, 256), oc = align_up(ob + nb, 256), ows = align_up(oc + nc, 256);
    Block blk;
    cudaError_t e = pool_acquire(ows + nws, dev, false, &blk);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? TOB_E_OOM : TOB_E_CUDA;
    }
    char* base = static_cast<char*>(blk.ptr);
    double* da = reinterpret_cast<double*>(base + oa);
    double* db = reinterpret_cast<double*>(base + ob);
    double* dc = reinterpret_cast<double*>(base + oc);
    double* dws = reinterpret_cast<double*>(base + ows);
    rc = TOB_OK;
    if ((e = cudaMemcpy(da, a, na, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(db, b, nb, cudaMemcpyHostToDevice)) != cudaSuccess) {
        set_error(std::string("cudaMemcpy: ") + cudaGetErrorString(e));
        rc = TOB_E_CUDA;
    }
    if (rc == TOB_OK) rc = tob_tensordot_device(da, rank_a, db, rank_b, axes_a, axes_b, n_axes, dc, dws, (int64_t)nws, 0, nullptr, nullptr);
    if (rc == TOB_OK && (e = cudaMemcpy(c, dc, nc, cudaMemcpyDeviceToHost)) != cudaSuccess) {
        set_error(std::string("cudaMemcpy: ") + cudaGetErrorString(e));
        rc = TOB_E_CUDA;
    }
    pool_release(blk);
    return rc;
}

}  // extern "C"
