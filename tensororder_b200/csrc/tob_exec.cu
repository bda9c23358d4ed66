// C-ABI entry points (include/tob200.h) and the device executor: arena, leaf upload, slice loop
// as a CUDA graph, per-op profiling, stand-alone tensordot / permute.
#include <cstdio>
#include <cstdlib>
#include <cstring>
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

struct tob_plan {
    Program prog;
    bool uploaded = false;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double* d_block = nullptr;  // [leaves | arena | workspace]
    double *d_leaves = nullptr, *d_arena = nullptr, *d_ws = nullptr;
    DevState* d_state = nullptr;
    long long* d_leaf_off = nullptr;
    int32_t* d_term_start = nullptr;
    uint8_t *d_id_bit = nullptr, *d_addr_bit = nullptr;
    DevState* h_state = nullptr;  // pinned
    double* h_stage = nullptr;    // pinned leaf staging
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    bool has_terms = false;
    double last_ms = 0;
    int64_t last_launches = 0;
    int64_t graph_launches_per_slice = 0;
    cudaStream_t own_stream = nullptr;
    std::vector<cudaEvent_t> gemm_events;  // pairs
    double last_gemm_ms = 0, last_gemm_flops = 0;
    int64_t last_gemm_launches = 0;
    double slice_flops = 0;
    MicroOpDev* d_micro_ops[2] = {nullptr, nullptr};
    int32_t* d_micro_start[2] = {nullptr, nullptr};
};

static bool g_configured = false;

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
    if (!g_configured) {
        CUDA_TRY(configure_kernels());
        g_configured = true;
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
    opt->reserved = 0;
}

const char* tob_last_error(void) { return g_error.c_str(); }
const char* tob_version(void) { return "tob200 0.1 (sm_100a)"; }

int tob_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int tob_plan_create(const tob_plan_desc* desc, const tob_options* opt, tob_plan** out) {
    if (!out) { set_error("out is NULL"); return TOB_E_INVALID; }
    *out = nullptr;
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
    return 8 * (p->prog.leaf_doubles + p->prog.arena_doubles + p->prog.ws_doubles) + 4096;
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

static void release_device(tob_plan* p) {
    if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
    if (p->graph) cudaGraphDestroy(p->graph);
    p->graph_exec = nullptr;
    p->graph = nullptr;
    if (p->d_block) cudaFree(p->d_block);
    if (p->d_state) cudaFree(p->d_state);
    if (p->d_leaf_off) cudaFree(p->d_leaf_off);
    if (p->d_term_start) cudaFree(p->d_term_start);
    if (p->d_id_bit) cudaFree(p->d_id_bit);
    if (p->d_addr_bit) cudaFree(p->d_addr_bit);
    if (p->h_state) cudaFreeHost(p->h_state);
    if (p->h_stage) cudaFreeHost(p->h_stage);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    for (int w = 0; w < 2; w++) {
        if (p->d_micro_ops[w]) cudaFree(p->d_micro_ops[w]);
        if (p->d_micro_start[w]) cudaFree(p->d_micro_start[w]);
        p->d_micro_ops[w] = nullptr;
        p->d_micro_start[w] = nullptr;
    }
    if (p->own_stream) cudaStreamDestroy(p->own_stream);
    for (cudaEvent_t e : p->gemm_events) cudaEventDestroy(e);
    p->gemm_events.clear();
    p->own_stream = nullptr;
    p->d_block = nullptr; p->d_state = nullptr; p->d_leaf_off = nullptr; p->d_term_start = nullptr;
    p->d_id_bit = nullptr; p->d_addr_bit = nullptr; p->h_state = nullptr; p->h_stage = nullptr;
    p->ev0 = p->ev1 = nullptr; p->stream = nullptr;
    p->uploaded = false;
}

void tob_plan_destroy(tob_plan* p) {
    if (!p) return;
    if (p->uploaded || p->stream) {
        cudaSetDevice(p->device);
        release_device(p);
    }
    delete p;
}

int tob_plan_upload(tob_plan* p, const double* leaf_data, int64_t n_doubles) {
    if (!p || !leaf_data) { set_error("NULL argument"); return TOB_E_INVALID; }
    Program& G = p->prog;
    if (n_doubles != G.src_leaf_len) { set_error("leaf buffer length does not match the plan"); return TOB_E_INVALID; }
    int rc = ensure_device(p->device);
    if (rc != TOB_OK) return rc;
    if (p->uploaded || p->stream) release_device(p);

    const int64_t need = tob_plan_peak_bytes(p);
    if (G.opt.mem_limit_bytes > 0 && need > G.opt.mem_limit_bytes) {
        set_error("plan needs " + std::to_string(need) + " bytes, above mem_limit_bytes");
        return TOB_E_OOM;
    }
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    if ((size_t)need > free_b) {
        set_error("plan needs " + std::to_string(need) + " bytes, device has " + std::to_string(free_b) + " free");
        return TOB_E_OOM;
    }
    CUDA_TRY(cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking));
    p->stream = p->own_stream;
    p->slice_flops = 0;
    for (const Op& op : G.slice_ops) p->slice_flops += op.flops;
    CUDA_TRY(cudaEventCreate(&p->ev0));
    CUDA_TRY(cudaEventCreate(&p->ev1));
    const int64_t block = G.leaf_doubles + G.arena_doubles + G.ws_doubles + 32;
    CUDA_TRY(cudaMalloc(&p->d_block, (size_t)block * 8));
    p->d_leaves = p->d_block;
    p->d_arena = p->d_block + G.leaf_doubles;
    p->d_ws = p->d_arena + G.arena_doubles;
    CUDA_TRY(cudaMalloc(&p->d_state, sizeof(DevState)));
    CUDA_TRY(cudaMallocHost(&p->h_state, sizeof(DevState)));

    // slice term tables
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
    CUDA_TRY(cudaMalloc(&p->d_leaf_off, sizeof(long long) * (L + 1)));
    CUDA_TRY(cudaMemsetAsync(p->d_leaf_off, 0, sizeof(long long) * (L + 1), p->stream));
    CUDA_TRY(cudaMalloc(&p->d_term_start, sizeof(int32_t) * (L + 1)));
    CUDA_TRY(cudaMemcpyAsync(p->d_term_start, term_start.data(), sizeof(int32_t) * (L + 1), cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaMalloc(&p->d_id_bit, id_bit.size() + 1));
    CUDA_TRY(cudaMalloc(&p->d_addr_bit, addr_bit.size() + 1));
    if (!id_bit.empty()) {
        CUDA_TRY(cudaMemcpyAsync(p->d_id_bit, id_bit.data(), id_bit.size(), cudaMemcpyHostToDevice, p->stream));
        CUDA_TRY(cudaMemcpyAsync(p->d_addr_bit, addr_bit.data(), addr_bit.size(), cudaMemcpyHostToDevice, p->stream));
    }
    // micro-subtree programs
    std::vector<MicroOpDev> micro_host[2];
    for (int w = 0; w < 2; w++) {
        const MicroProgram& mp = G.micro[w];
        if (mp.ops.empty()) continue;
        for (const Op& op : mp.ops) {
            MicroOpDev d;
            memset(&d, 0, sizeof(d));
            d.a_off = op.a.offset; d.b_off = op.b.offset; d.c_off = op.c_offset;
            d.a_leaf = op.a.leaf; d.b_leaf = op.b.leaf;
            d.mask_m = (uint16_t)op.mask_m;
            d.a_space = (uint8_t)op.a.space; d.b_space = (uint8_t)op.b.space;
            d.m = (uint8_t)op.m; d.n = (uint8_t)op.n; d.k = (uint8_t)op.k;
            micro_host[w].push_back(d);
        }
        CUDA_TRY(cudaMalloc(&p->d_micro_ops[w], micro_host[w].size() * sizeof(MicroOpDev)));
        CUDA_TRY(cudaMemcpyAsync(p->d_micro_ops[w], micro_host[w].data(), micro_host[w].size() * sizeof(MicroOpDev),
                                 cudaMemcpyHostToDevice, p->stream));
        CUDA_TRY(cudaMalloc(&p->d_micro_start[w], mp.cta_start.size() * sizeof(int32_t)));
        CUDA_TRY(cudaMemcpyAsync(p->d_micro_start[w], mp.cta_start.data(), mp.cta_start.size() * sizeof(int32_t),
                                 cudaMemcpyHostToDevice, p->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(p->stream));  // the std::vectors above die at scope exit

    // leaves: permute on the host into the canonical device layout, one H2D copy
    CUDA_TRY(cudaMallocHost(&p->h_stage, (size_t)std::max<int64_t>(G.leaf_doubles, 1) * 8));
    memset(p->h_stage, 0, (size_t)G.leaf_doubles * 8);
    for (const LeafInfo& Lf : G.leaves) {
        const int64_t n = (int64_t)1 << Lf.rank;
        const double* src = leaf_data + Lf.src_offset;
        double* dst = p->h_stage + Lf.dev_offset;
        for (int64_t d = 0; d < n; d++) {
            int64_t s = 0;
            for (int q = 0; q < Lf.rank; q++) s |= ((d >> q) & 1) << Lf.src_bit[q];
            dst[d] = src[s];
        }
    }
    if (G.leaf_doubles > 0)
        CUDA_TRY(cudaMemcpyAsync(p->d_leaves, p->h_stage, (size_t)G.leaf_doubles * 8, cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    p->uploaded = true;
    return TOB_OK;
}

static KParams make_params(const tob_plan* p, const Op& op) {
    KParams k;
    memset(&k, 0, sizeof(k));
    auto base = [&](const OperandRef& r) -> const double* { return (r.space == 0 ? p->d_leaves : p->d_arena) + r.offset; };
    k.a = base(op.a);
    k.b = base(op.b);
    k.c = p->d_arena + op.c_offset;
    k.ws = p->d_ws;
    k.leaf_off = p->d_leaf_off;
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
    return k;
}

static cudaError_t launch_op(tob_plan* p, const Op& op, int* launches) {
    if (op.kind == OP_ACCUM) {
        const double* root = (op.a.space == 0 ? p->d_leaves : p->d_arena) + op.a.offset;
        (*launches)++;
        return launch_accum(p->d_state, root, p->d_leaf_off, op.a.leaf, p->stream);
    }
    if (op.kind == OP_MICRO) {
        const int w = op.micro_which;
        (*launches)++;
        return launch_microtree(p->d_micro_ops[w], p->d_micro_start[w], (int)p->prog.micro[w].cta_start.size() - 1,
                                p->d_leaves, p->d_arena, p->d_leaf_off, p->stream);
    }
    KParams k = make_params(p, op);
    return launch_contract(op, k, p->stream, launches);
}

static cudaError_t launch_slice(tob_plan* p, int* launches) {
    cudaError_t e;
    if (p->has_terms) {
        SliceTables t{p->d_term_start, p->d_id_bit, p->d_addr_bit, p->d_leaf_off, (int32_t)p->prog.leaves.size()};
        e = launch_begin_slice(p->d_state, t, p->stream);
        (*launches)++;
        if (e != cudaSuccess) return e;
    }
    for (const Op& op : p->prog.slice_ops) {
        e = launch_op(p, op, launches);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

int tob_plan_run(tob_plan* p, uint64_t first, uint64_t count, uint64_t stride, double* result) {
    if (!p || !result) { set_error("NULL argument"); return TOB_E_INVALID; }
    if (!p->uploaded) { set_error("tob_plan_upload has not been called"); return TOB_E_INVALID; }
    const uint64_t nslices = tob_plan_num_slices(p);
    if (count > 0 && (first >= nslices || first + (count - 1) * stride >= nslices)) {
        set_error("slice range exceeds the number of slices");
        return TOB_E_INVALID;
    }
    int rc = ensure_device(p->device);
    if (rc != TOB_OK) return rc;
    int launches = 0;
    p->h_state->next_slice = first;
    p->h_state->stride = stride;
    p->h_state->acc = 0.0;
    p->h_state->pad = 0.0;
    CUDA_TRY(cudaMemcpyAsync(p->d_state, p->h_state, sizeof(DevState), cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaEventRecord(p->ev0, p->stream));
    const int ug = p->prog.opt.use_graph;
    const bool as_graph = (ug == 1) || (ug == 2 && p->slice_flops < 2e9);
    size_t n_gemm = 0;
    double gemm_flops = 0;
    // plain stream launches: bracket every DMMA GEMM with CUDA events (per-kernel roofline, bench.py)
    auto timed_op = [&](const Op& op) -> int {
        if (op.kind != OP_GEMM || as_graph) {
            CUDA_TRY(launch_op(p, op, &launches));
            return TOB_OK;
        }
        if (p->gemm_events.size() < 2 * (n_gemm + 1)) {
            cudaEvent_t a, b;
            CUDA_TRY(cudaEventCreate(&a));
            CUDA_TRY(cudaEventCreate(&b));
            p->gemm_events.push_back(a);
            p->gemm_events.push_back(b);
        }
        CUDA_TRY(cudaEventRecord(p->gemm_events[2 * n_gemm], p->stream));
        CUDA_TRY(launch_op(p, op, &launches));
        CUDA_TRY(cudaEventRecord(p->gemm_events[2 * n_gemm + 1], p->stream));
        n_gemm++;
        gemm_flops += op.flops;
        return TOB_OK;
    };
    if (count > 0) {
        for (const Op& op : p->prog.invariant_ops) {
            int rc2 = timed_op(op);
            if (rc2 != TOB_OK) return rc2;
        }
        if (as_graph) {
            if (!p->graph_exec) {
                int per_slice = 0;
                CUDA_TRY(cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal));
                cudaError_t e = launch_slice(p, &per_slice);
                cudaError_t e2 = cudaStreamEndCapture(p->stream, &p->graph);
                CUDA_TRY(e);
                CUDA_TRY(e2);
                CUDA_TRY(cudaGraphInstantiate(&p->graph_exec, p->graph, 0));
                p->graph_launches_per_slice = per_slice;
            }
            for (uint64_t s = 0; s < count; s++) CUDA_TRY(cudaGraphLaunch(p->graph_exec, p->stream));
            launches += (int)(p->graph_launches_per_slice * count);
        } else {
            for (uint64_t s = 0; s < count; s++) {
                if (p->has_terms) {
                    SliceTables t{p->d_term_start, p->d_id_bit, p->d_addr_bit, p->d_leaf_off, (int32_t)p->prog.leaves.size()};
                    CUDA_TRY(launch_begin_slice(p->d_state, t, p->stream));
                    launches++;
                }
                for (const Op& op : p->prog.slice_ops) {
                    int rc2 = timed_op(op);
                    if (rc2 != TOB_OK) return rc2;
                }
            }
        }
    }
    CUDA_TRY(cudaEventRecord(p->ev1, p->stream));
    CUDA_TRY(cudaMemcpyAsync(p->h_state, p->d_state, sizeof(DevState), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, p->ev0, p->ev1));
    p->last_ms = ms;
    p->last_launches = launches;
    p->last_gemm_ms = 0;
    for (size_t i = 0; i < n_gemm; i++) {
        float g = 0;
        CUDA_TRY(cudaEventElapsedTime(&g, p->gemm_events[2 * i], p->gemm_events[2 * i + 1]));
        p->last_gemm_ms += g;
    }
    p->last_gemm_flops = gemm_flops;
    p->last_gemm_launches = (int64_t)n_gemm;
    *result = p->h_state->acc;
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
    if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }
    if (p->graph) { cudaGraphDestroy(p->graph); p->graph = nullptr; }
    p->stream = stream ? (cudaStream_t)stream : p->own_stream;
    return TOB_OK;
}

double tob_plan_last_ms(const tob_plan* p) { return p->last_ms; }
int64_t tob_plan_last_launches(const tob_plan* p) { return p->last_launches; }

int tob_plan_profile(tob_plan* p, uint64_t slice, float* ms_per_op, int64_t n_ops, double* result) {
    if (!p || !ms_per_op) { set_error("NULL argument"); return TOB_E_INVALID; }
    if (!p->uploaded) { set_error("tob_plan_upload has not been called"); return TOB_E_INVALID; }
    if (n_ops != tob_plan_num_ops(p)) { set_error("n_ops mismatch"); return TOB_E_INVALID; }
    if (slice >= tob_plan_num_slices(p)) { set_error("slice out of range"); return TOB_E_INVALID; }
    int rc = ensure_device(p->device);
    if (rc != TOB_OK) return rc;
    int launches = 0;
    p->h_state->next_slice = slice;
    p->h_state->stride = 1;
    p->h_state->acc = 0.0;
    CUDA_TRY(cudaMemcpyAsync(p->d_state, p->h_state, sizeof(DevState), cudaMemcpyHostToDevice, p->stream));
    std::vector<cudaEvent_t> ev(n_ops + 1);
    for (auto& e : ev) CUDA_TRY(cudaEventCreate(&e));
    int64_t i = 0;
    for (const Op& op : p->prog.invariant_ops) {
        CUDA_TRY(cudaEventRecord(ev[i], p->stream));
        CUDA_TRY(launch_op(p, op, &launches));
        i++;
    }
    if (p->has_terms) {
        SliceTables t{p->d_term_start, p->d_id_bit, p->d_addr_bit, p->d_leaf_off, (int32_t)p->prog.leaves.size()};
        CUDA_TRY(launch_begin_slice(p->d_state, t, p->stream));
    }
    for (const Op& op : p->prog.slice_ops) {
        CUDA_TRY(cudaEventRecord(ev[i], p->stream));
        CUDA_TRY(launch_op(p, op, &launches));
        i++;
    }
    CUDA_TRY(cudaEventRecord(ev[i], p->stream));
    CUDA_TRY(cudaMemcpyAsync(p->h_state, p->d_state, sizeof(DevState), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    for (int64_t j = 0; j < n_ops; j++) CUDA_TRY(cudaEventElapsedTime(&ms_per_op[j], ev[j], ev[j + 1]));
    for (auto& e : ev) cudaEventDestroy(e);
    if (result) *result = p->h_state->acc;
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
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ms) {
        CUDA_TRY(cudaEventCreate(&e0));
        CUDA_TRY(cudaEventCreate(&e1));
        CUDA_TRY(cudaEventRecord(e0, stream));
    }
    CUDA_TRY(launch_permute(in, out, rank, src_bit, stream));
    if (ms) {
        CUDA_TRY(cudaEventRecord(e1, stream));
        CUDA_TRY(cudaEventSynchronize(e1));
        CUDA_TRY(cudaEventElapsedTime(ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
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
    off_ws = need;
    const int64_t ws_avail = (workspace ? workspace_bytes : 0) - need;
    choose_kernel(&op, kernel_policy, true);
    while (op.ksplit_log2 > 0 && ((int64_t)8 << (op.m + op.n + op.ksplit_log2)) > ws_avail) op.ksplit_log2--;
    if ((pa || pb) && (!workspace || workspace_bytes < need)) {
        set_error("tensordot needs a workspace of at least " + std::to_string(need) + " bytes for the operand permutations");
        return TOB_E_INVALID;
    }
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (ms) {
        for (auto& e : ev) CUDA_TRY(cudaEventCreate(&e));
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
    kp.mask_m = op.mask_m;
    const int tot = op.m + op.n;
    kp.mask_n = tot == 0 ? 0ull : (~op.mask_m & (tot >= 64 ? ~0ull : ((1ull << tot) - 1ull)));
    kp.runs_m = make_runs(kp.mask_m);
    kp.runs_n = make_runs(kp.mask_n);
    int launches = 0;
    CUDA_TRY(launch_contract(op, kp, stream, &launches));
    if (ms) {
        CUDA_TRY(cudaEventRecord(ev[2], stream));
        CUDA_TRY(cudaEventSynchronize(ev[2]));
        CUDA_TRY(cudaEventElapsedTime(&ms[0], ev[0], ev[1]));  // permutations
        CUDA_TRY(cudaEventElapsedTime(&ms[1], ev[1], ev[2]));  // contraction
        ms[2] = (float)op.kind;
        for (auto& e : ev) if (e) cudaEventDestroy(e);
    }
    return TOB_OK;
}

int tob_tensordot_host(const double* a, int32_t rank_a, const double* b, int32_t rank_b, const int32_t* axes_a,
                       const int32_t* axes_b, int32_t n_axes, double* c) {
    if (rank_a < 0 || rank_b < 0 || rank_a > 32 || rank_b > 32 || n_axes < 0 || n_axes > rank_a || n_axes > rank_b) {
        set_error("bad tensordot arguments");
        return TOB_E_INVALID;
    }
    int rc = ensure_device(0);
    if (rc != TOB_OK) return rc;
    const int rank_c = rank_a + rank_b - 2 * n_axes;
    const size_t na = (size_t)8 << rank_a, nb = (size_t)8 << rank_b, nc = (size_t)8 << rank_c;
    const size_t nws = na + nb + (nc << 4) + 256;
    double *da = nullptr, *db = nullptr, *dc = nullptr, *dws = nullptr;
    auto cleanup = [&]() { cudaFree(da); cudaFree(db); cudaFree(dc); cudaFree(dws); };
    cudaError_t e;
    if ((e = cudaMalloc(&da, na)) != cudaSuccess || (e = cudaMalloc(&db, nb)) != cudaSuccess ||
        (e = cudaMalloc(&dc, nc)) != cudaSuccess || (e = cudaMalloc(&dws, nws)) != cudaSuccess) {
        cleanup();
        set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? TOB_E_OOM : TOB_E_CUDA;
    }
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
    cleanup();
    return rc;
}

}  // extern "C"
