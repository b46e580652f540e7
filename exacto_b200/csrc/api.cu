// api.cu -- C ABI (include/exacto_b200.h): context construction (host-side
// precomputation of plans and HPS constants), device memory plumbing, and the
// batched hot-path entry points that sequence the kernels of kernels.cu.
//
// The host-side restatements of the reference (params, plans, dispatch, dBFV plan)
// live in host_setup.cpp; this file adds the CUDA side: uploads, workspaces, streams
// and the launch sequence, incl. reduction::reduce (dbfv/reduction.rs:15-60).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/exacto_b200.h"
#include "kernels.cuh"

using namespace exb;


// ---------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

#define EXB_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(EXB_CUDA_ERROR, std::string(#call) + ": " + cudaGetErrorString(e_));    \
    } while (0)

// ---------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------
// Workspace slots.  Device-resident entry points take a slot of the `ws` pool for the duration of their
// enqueue (two host threads on two streams get two slots and overlap on the host and on the GPU); the
// host-buffer pipeline owns the `hs` slots, each with its own stream and staging buffers.
static const int kSlots = 4;
static const int kHostSlots = 8;                        // allocated; Tuning::host_slots of them rotate

struct Workspace {
    std::mutex mu;                                        // held while a call enqueues work that uses the slot
    cudaStream_t stream = nullptr;                        // host pipeline slots only (owned)
    cudaEvent_t done = nullptr;                           // recorded after the last launch that used the slot
    const void *last_tag = nullptr;                       // stream handle of that work: compared, never dereferenced
    bool used = false;
    u64 *ext = nullptr, *r01 = nullptr, *excess = nullptr, *wide = nullptr;
    void *digits = nullptr;
    size_t ext_b = 0, r01_b = 0, excess_b = 0, digits_b = 0, wide_b = 0;
    u64 *in1 = nullptr, *in2 = nullptr, *out = nullptr;   // staging for the *_host entry points
    size_t in_b = 0, out_b = 0;
};

struct StageEvents {
    cudaEvent_t ev[6];
    bool has_reduce, has_c2;
};

// One asynchronous host-buffer call in flight: the events that mark its last chunk on every slot it used.
struct Ticket {
    std::vector<cudaEvent_t> events;
};

struct Tuning {                                           // exb_context_set_option
    size_t device_chunk_bytes = (size_t)4 << 30;          // workspace budget of one device-resident chunk
    size_t host_chunk_products = 0;                       // products per chunk of the host-buffer pipeline; 0 = auto:
                                                          // 1024 for a synchronous call (short fill / drain), 3996 when
                                                          // calls are pipelined (measured optimum, tools/sweep_e2e.py)
    int host_slots = 4;                                   // pipeline depth (streams / staging sets in rotation)
    int kshard_kernel_stores = 0;                         // k-shard exchange: 0 = copy engines (peer DMA), 1 = stores from
                                                          // the relin kernel's epilogue (measured slower: SM-issued NVLink writes)
};

struct exb_context : HostSetup {
    int device = 0;
    bool profiling = false;
    Tuning tune;
    std::vector<Tw *> d_tables;       // owned device twiddle tables
    Workspace ws[kSlots];             // device-resident calls
    Workspace hs[kHostSlots];         // host-buffer pipeline
    std::mutex mu;                    // slot selection, profiling record, tickets
    std::mutex host_mu;               // one host-buffer call enqueues at a time (its chunks stay in order)
    cudaStream_t xs[kMaxPeers] = {};  // k-shard exchange: one copy stream per peer (created on first use)
    cudaEvent_t xev[kMaxPeers + 1] = {};
    std::vector<StageEvents> events;
    unsigned rr = 0, host_rr = 0;
    uint64_t next_ticket = 1;
    std::unordered_map<uint64_t, Ticket> tickets;

    // true when reduce() (dbfv/reduction.rs:28-52) has nothing to add: every small representative is zero
    bool excess_free(uint32_t d, uint64_t base, uint64_t pm) const {
        if (d < 2 || d > (uint32_t)kMaxDigits || base < 2) return true;
        std::vector<int64_t> reps((size_t)(d - 1) * d, 0);
        std::string err;
        if (host_small_reps(base, d, pm, reps.data(), &err) != EXB_OK) return true;   // the plan reports the error
        for (int64_t r : reps) if (r != 0) return false;
        return true;
    }
};

// Take a workspace slot for work that is about to be enqueued on `stream`: the slot this stream used last if it
// is free, else any free slot, else wait for one.  Work left in the slot by ANOTHER stream is ordered before the
// new work with the slot's own event (cudaStreamWaitEvent): no host blocking, and a user stream that has since
// been destroyed is never touched.
static int acquire(exb_context *c, cudaStream_t stream, std::unique_lock<std::mutex> *held, Workspace **out) {
    Workspace *w = nullptr;
    {
        std::lock_guard<std::mutex> g(c->mu);
        for (int pass = 0; pass < 2 && !w; pass++)
            for (int i = 0; i < kSlots && !w; i++) {
                Workspace &cand = c->ws[(c->rr + i) % kSlots];
                if (pass == 0 && !(cand.used && cand.last_tag == (const void *)stream)) continue;
                std::unique_lock<std::mutex> l(cand.mu, std::try_to_lock);
                if (l.owns_lock()) { *held = std::move(l); w = &cand; }
            }
        if (!w) w = &c->ws[c->rr++ % kSlots];
    }
    if (!held->owns_lock()) *held = std::unique_lock<std::mutex>(w->mu);
    if (w->used && w->last_tag != (const void *)stream) EXB_CUDA(cudaStreamWaitEvent(stream, w->done, 0));
    *out = w;
    return EXB_OK;
}

// The slot's work has been enqueued on `stream`.
static int release(Workspace *w, cudaStream_t stream) {
    EXB_CUDA(cudaEventRecord(w->done, stream));
    w->last_tag = (const void *)stream;
    w->used = true;
    return EXB_OK;
}

struct exb_relin_key {
    exb_context *ctx = nullptr;
    u64 *d_mont = nullptr;            // [num_keys][2][n], Montgomery form
    u32 num_keys = 0;
};

static int grow(void **p, size_t *have, size_t want) {
    if (*have >= want) return EXB_OK;
    if (*p) EXB_CUDA(cudaFree(*p));
    *p = nullptr; *have = 0;
    EXB_CUDA(cudaMalloc(p, want));
    *have = want;
    return EXB_OK;
}

// in2 and out share the growth bookkeeping of out_b
static int grow_in2_out(Workspace &w, size_t want) {
    if (w.out_b >= want) return EXB_OK;
    if (w.in2) EXB_CUDA(cudaFree(w.in2));
    if (w.out) EXB_CUDA(cudaFree(w.out));
    w.in2 = w.out = nullptr; w.out_b = 0;
    EXB_CUDA(cudaMalloc((void **)&w.in2, want));
    EXB_CUDA(cudaMalloc((void **)&w.out, want));
    w.out_b = want;
    return EXB_OK;
}

extern "C" const char *exb_last_error(void) { return g_err.c_str(); }
extern "C" const char *exb_version(void) { return "exacto_b200 0.1 (sm_100a)"; }
extern "C" unsigned long long exb_launch_count(void) { return exb::g_launch_count.load(std::memory_order_relaxed); }

static void free_workspace(Workspace &w) {
    cudaFree(w.ext); cudaFree(w.r01); cudaFree(w.excess); cudaFree(w.digits); cudaFree(w.wide);
    cudaFree(w.in1); cudaFree(w.in2); cudaFree(w.out);
    if (w.done) cudaEventDestroy(w.done);
    if (w.stream) cudaStreamDestroy(w.stream);
}

extern "C" void exb_context_destroy(exb_context *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (StageEvents &se : c->events) for (auto &e : se.ev) cudaEventDestroy(e);
    for (auto &kv : c->tickets) for (cudaEvent_t e : kv.second.events) cudaEventDestroy(e);
    for (Tw *t : c->d_tables) cudaFree(t);
    for (Workspace &w : c->ws) free_workspace(w);
    for (Workspace &w : c->hs) free_workspace(w);
    for (cudaStream_t s : c->xs) if (s) cudaStreamDestroy(s);
    for (cudaEvent_t e : c->xev) if (e) cudaEventDestroy(e);
    delete c;
}

extern "C" int exb_context_create_ex(const exb_bfv_params *p, int device, uint32_t flags, exb_context **out) {
    if (!p || !out) return fail(EXB_INVALID_PARAM, "null argument");
    *out = nullptr;
    if (flags & ~(uint32_t)EXB_CTX_REFERENCE_AUX_BASIS) return fail(EXB_INVALID_PARAM, "unknown context flag");
    exb_context *c = new exb_context();
    c->device = device;
    std::string err;
    int rc = host_setup_build(p, c, &err, flags);
    if (rc != EXB_OK) { delete c; return fail(rc, err); }
    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) {
        delete c;
        return fail(EXB_CUDA_ERROR, std::string("cudaSetDevice: ") + cudaGetErrorString(ce));
    }
    const u32 n = c->n;
    for (int b = 0; b < kMaxBases; b++) {
        if (!c->has_plan[b]) continue;
        Tw *df = nullptr, *di = nullptr;
        if (cudaMalloc(&df, sizeof(Tw) * n) != cudaSuccess || cudaMalloc(&di, sizeof(Tw) * n) != cudaSuccess ||
            cudaMemcpy(df, c->twf[b].data(), sizeof(Tw) * n, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(di, c->twi[b].data(), sizeof(Tw) * n, cudaMemcpyHostToDevice) != cudaSuccess) {
            std::string m = std::string("twiddle upload: ") + cudaGetErrorString(cudaGetLastError());
            cudaFree(df); cudaFree(di);
            exb_context_destroy(c);
            return fail(EXB_CUDA_ERROR, m);
        }
        c->d_tables.push_back(df); c->d_tables.push_back(di);
        c->P.twf[b] = df; c->P.twi[b] = di;
    }
    for (u32 i = 0; c->P.sb.enabled && i < c->P.sb.K; i++) {
        Tw32 *df = nullptr, *di = nullptr;
        if (cudaMalloc(&df, sizeof(Tw32) * n) != cudaSuccess || cudaMalloc(&di, sizeof(Tw32) * n) != cudaSuccess ||
            cudaMemcpy(df, c->twf32[i].data(), sizeof(Tw32) * n, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(di, c->twi32[i].data(), sizeof(Tw32) * n, cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaFree(df); cudaFree(di);
            exb_context_destroy(c);
            return fail(EXB_CUDA_ERROR, "small-basis twiddle upload failed");
        }
        c->d_tables.push_back(reinterpret_cast<Tw *>(df)); c->d_tables.push_back(reinterpret_cast<Tw *>(di));
        c->P.sb.twf[i] = df; c->P.sb.twi[i] = di;
    }
    if (c->rns_enabled) {
        auto up = [&](const std::vector<Tw> &h, const Tw **dst) {
            Tw *d = nullptr;
            if (cudaMalloc(&d, sizeof(Tw) * n) != cudaSuccess ||
                cudaMemcpy(d, h.data(), sizeof(Tw) * n, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return false; }
            c->d_tables.push_back(d);
            *dst = d;
            return true;
        };
        bool ok = true;
        for (u32 l = 0; ok && l < c->R.L; l++) ok = up(c->rns_twf_q[l], &c->T.twf_q[l]) && up(c->rns_twi_q[l], &c->T.twi_q[l]);
        for (u32 k = 0; ok && k < c->R.K; k++) ok = up(c->rns_twf_e[k], &c->T.twf_e[k]) && up(c->rns_twi_e[k], &c->T.twi_e[k]);
        if (!ok) {
            exb_context_destroy(c);
            return fail(EXB_CUDA_ERROR, "multi-prime twiddle upload failed");
        }
    }
    for (Workspace &w : c->ws)
        if (cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming) != cudaSuccess) {
            exb_context_destroy(c);
            return fail(EXB_CUDA_ERROR, "cudaEventCreate failed");
        }
    for (Workspace &w : c->hs)
        if (cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming) != cudaSuccess) {
            exb_context_destroy(c);
            return fail(EXB_CUDA_ERROR, "cudaStreamCreate failed");
        }
    launch_prepare(device);
    *out = c;
    return EXB_OK;
}

extern "C" int exb_context_create(const exb_bfv_params *p, int device, exb_context **out) {
    return exb_context_create_ex(p, device, 0, out);
}

// Tuning knobs (tests and lab sweeps use them; the defaults are the measured optimum).
extern "C" int exb_context_set_option(exb_context *c, const char *name, int64_t value) {
    if (!c || !name) return fail(EXB_INVALID_PARAM, "null argument");
    std::lock_guard<std::mutex> g(c->mu);
    const std::string k(name);
    if (k == "device_chunk_bytes" && value > 0) c->tune.device_chunk_bytes = (size_t)value;
    else if (k == "host_chunk_products" && value > 0) c->tune.host_chunk_products = (size_t)value;
    else if (k == "host_slots" && value >= 2 && value <= kHostSlots) {
        std::lock_guard<std::mutex> hl(c->host_mu);
        cudaSetDevice(c->device);
        for (Workspace &w : c->hs) cudaStreamSynchronize(w.stream);
        c->tune.host_slots = (int)value;
    }
    else if (k == "tensor_per_product") c->P.tensor_per_product = value ? 1u : 0u;
    else if (k == "relin_narrow") c->P.relin_narrow = value ? 1u : 0u;
    else if (k == "kshard_kernel_stores") c->tune.kshard_kernel_stores = value ? 1 : 0;
    else if (k == "ntt_cp_async") g_ntt_path.store(value ? 1 : 0);     // process-wide: A/B of the two n = 4096 transform kernels
    else return fail(EXB_INVALID_PARAM, "unknown option or value out of range: " + k);
    return EXB_OK;
}

extern "C" int exb_profile_enable(exb_context *c, int on) {
    if (!c) return fail(EXB_INVALID_PARAM, "null context");
    c->profiling = on != 0;
    return EXB_OK;
}

extern "C" int exb_profile_read(exb_context *c, double *ms, unsigned long long *launches) {
    if (!c || !ms || !launches) return fail(EXB_INVALID_PARAM, "null argument");
    std::vector<StageEvents> evs;
    {
        std::lock_guard<std::mutex> lock(c->mu);
        evs.swap(c->events);
    }
    EXB_CUDA(cudaSetDevice(c->device));
    for (StageEvents &se : evs) {
        EXB_CUDA(cudaEventSynchronize(se.ev[5]));
        for (int k = 0; k < 5; k++) {
            float t = 0.f;
            EXB_CUDA(cudaEventElapsedTime(&t, se.ev[k], se.ev[k + 1]));
            if ((k != 2 || se.has_c2) && (k != 4 || se.has_reduce)) { ms[k] += t; launches[k] += 1; }
        }
        for (auto &e : se.ev) cudaEventDestroy(e);
    }
    return EXB_OK;
}

extern "C" int exb_context_gadget(const exb_context *c, uint64_t *base, uint32_t *digits) {
    if (!c) return fail(EXB_INVALID_PARAM, "null context");
    if (base) *base = c->gadget_base;
    if (digits) *digits = c->gadget_digits;
    return EXB_OK;
}

extern "C" int exb_context_psi(const exb_context *c, uint32_t idx, uint64_t *psi) {
    if (!c || !psi) return fail(EXB_INVALID_PARAM, "null argument");
    if (idx >= c->psi.size()) return fail(EXB_INVALID_PARAM, "modulus index out of range");
    *psi = c->psi[idx];
    return EXB_OK;
}

extern "C" int exb_ntt_format_id(const exb_context *c, uint32_t idx, uint64_t *id) {
    if (!c || !id) return fail(EXB_INVALID_PARAM, "null argument");
    if (idx >= c->psi.size()) return fail(EXB_INVALID_PARAM, "modulus index out of range");
    std::vector<u64> all;
    all.push_back(c->ct_moduli[0]);
    for (u64 a : c->aux_moduli) all.push_back(a);
    for (size_t i = 1; i < c->ct_moduli.size(); i++) all.push_back(c->ct_moduli[i]);
    u64 h = 0xcbf29ce484222325ull;                          // FNV-1a over (n, q, psi, ordering tag)
    const u64 words[4] = {c->n, all[idx], c->psi[idx], 0x4354626974726576ull /* "CTbitrev" */};
    for (u64 w : words) for (int b = 0; b < 8; b++) { h ^= (w >> (8 * b)) & 0xff; h *= 0x100000001b3ull; }
    *id = ((u64)EXB_NTT_FORMAT_VERSION << 56) ^ (h & 0x00ffffffffffffffull);
    return EXB_OK;
}

// ---------------------------------------------------------------------------------
// plumbing
// ---------------------------------------------------------------------------------
extern "C" int exb_device_alloc(exb_context *c, size_t bytes, void **p) {
    if (!c || !p) return fail(EXB_INVALID_PARAM, "null argument");
    EXB_CUDA(cudaSetDevice(c->device));
    EXB_CUDA(cudaMalloc(p, bytes ? bytes : 8));
    return EXB_OK;
}
extern "C" int exb_device_free(exb_context *c, void *p) {
    if (!c) return fail(EXB_INVALID_PARAM, "null context");
    EXB_CUDA(cudaSetDevice(c->device));
    EXB_CUDA(cudaFree(p));
    return EXB_OK;
}
extern "C" int exb_copy_to_device(exb_context *c, void *dst, const void *src, size_t bytes, void *stream) {
    if (!c) return fail(EXB_INVALID_PARAM, "null context");
    EXB_CUDA(cudaSetDevice(c->device));
    EXB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return EXB_OK;
}
extern "C" int exb_copy_to_host(exb_context *c, void *dst, const void *src, size_t bytes, void *stream) {
    if (!c) return fail(EXB_INVALID_PARAM, "null context");
    EXB_CUDA(cudaSetDevice(c->device));
    EXB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return EXB_OK;
}
extern "C" int exb_synchronize(exb_context *c, void *stream) {
    if (!c) return fail(EXB_INVALID_PARAM, "null context");
    EXB_CUDA(cudaSetDevice(c->device));
    EXB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return EXB_OK;
}

static int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(EXB_CUDA_ERROR, std::string(what) + ": " + cudaGetErrorString(e));
    return EXB_OK;
}

// Modulus indices: 0 = q_0, 1..A = aux primes, then the remaining ciphertext primes q_1.. (multi-prime sets).
static int rns_prime_of(const exb_context *c, u32 idx) {          // ciphertext prime l >= 1 behind `idx`, or -1
    const u32 A = c->user_aux;
    if (!c->rns_enabled || idx <= A) return -1;
    const u32 l = idx - A;
    return l < c->R.L ? (int)l : -1;
}
static int check_base(const exb_context *c, u32 idx) {
    if (!c) return fail(EXB_INVALID_PARAM, "null context");
    const u32 A = c->user_aux;                       // an internal auxiliary pair is not addressable
    if (rns_prime_of(c, idx) >= 0) return EXB_OK;
    if (idx != 0 && (A > (u32)kMaxAux || idx > A))
        return fail(EXB_MODULUS_MISMATCH, "modulus index " + std::to_string(idx) + " has no device plan");
    return EXB_OK;
}
static const Modulus &modulus_of(const exb_context *c, u32 idx) {
    const int l = rns_prime_of(c, idx);
    return l >= 0 ? c->R.q[l] : c->P.mod[idx];
}
static void ntt_any(exb_context *c, u32 idx, bool fwd, const u64 *in, u64 *out, size_t count, cudaStream_t s) {
    const int l = rns_prime_of(c, idx);
    if (l >= 0) launch_ntt_plan(c->R.q[l], fwd ? c->T.twf_q[l] : c->T.twi_q[l], fwd ? c->T.headf_q[l] : c->T.headi_q[l],
                                c->logn, fwd, in, out, count, s);
    else if (fwd) launch_ntt_fwd(c->P, (int)idx, in, out, count, s);
    else launch_ntt_inv(c->P, (int)idx, in, out, count, s);
}

// ---------------------------------------------------------------------------------
// ring ops
// ---------------------------------------------------------------------------------
extern "C" int exb_ntt_forward(exb_context *c, uint32_t idx, const uint64_t *in, uint64_t *out, size_t count,
                               void *stream) {
    int rc = check_base(c, idx);
    if (rc) return rc;
    EXB_CUDA(cudaSetDevice(c->device));
    ntt_any(c, idx, true, in, out, count, (cudaStream_t)stream);
    return check_launch("ntt_fwd");
}
extern "C" int exb_ntt_inverse(exb_context *c, uint32_t idx, const uint64_t *in, uint64_t *out, size_t count,
                               void *stream) {
    int rc = check_base(c, idx);
    if (rc) return rc;
    EXB_CUDA(cudaSetDevice(c->device));
    ntt_any(c, idx, false, in, out, count, (cudaStream_t)stream);
    return check_launch("ntt_inv");
}

static int ntt_host(exb_context *c, u32 idx, const u64 *in, u64 *out, size_t count, bool fwd) {
    int rc = check_base(c, idx);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(c->host_mu);
    if (count == 0) return EXB_OK;
    EXB_CUDA(cudaSetDevice(c->device));
    Workspace &w = c->hs[0];
    const size_t bytes = count * c->n * sizeof(u64);
    rc = grow((void **)&w.in1, &w.in_b, bytes);
    if (rc) return rc;
    EXB_CUDA(cudaMemcpyAsync(w.in1, in, bytes, cudaMemcpyHostToDevice, w.stream));
    ntt_any(c, idx, fwd, w.in1, w.in1, count, w.stream);
    rc = check_launch("ntt_host");
    if (rc) return rc;
    EXB_CUDA(cudaMemcpyAsync(out, w.in1, bytes, cudaMemcpyDeviceToHost, w.stream));
    EXB_CUDA(cudaStreamSynchronize(w.stream));
    return EXB_OK;
}
extern "C" int exb_ntt_forward_host(exb_context *c, uint32_t idx, const uint64_t *in, uint64_t *out, size_t count) {
    return ntt_host(c, idx, in, out, count, true);
}
extern "C" int exb_ntt_inverse_host(exb_context *c, uint32_t idx, const uint64_t *in, uint64_t *out, size_t count) {
    return ntt_host(c, idx, in, out, count, false);
}

static int poly_op(exb_context *c, u32 idx, PolyOp op, const u64 *a, const u64 *b, u64 scalar, u64 *out,
                   size_t words, void *stream) {
    int rc = check_base(c, idx);
    if (rc) return rc;
    EXB_CUDA(cudaSetDevice(c->device));
    launch_poly_op(modulus_of(c, idx), op, a, b, scalar, out, words, (cudaStream_t)stream);
    return check_launch("poly_op");
}
extern "C" int exb_poly_add(exb_context *c, uint32_t i, const uint64_t *a, const uint64_t *b, uint64_t *o, size_t w, void *s) {
    return poly_op(c, i, OP_ADD, a, b, 0, o, w, s);
}
extern "C" int exb_poly_sub(exb_context *c, uint32_t i, const uint64_t *a, const uint64_t *b, uint64_t *o, size_t w, void *s) {
    return poly_op(c, i, OP_SUB, a, b, 0, o, w, s);
}
extern "C" int exb_poly_neg(exb_context *c, uint32_t i, const uint64_t *a, uint64_t *o, size_t w, void *s) {
    return poly_op(c, i, OP_NEG, a, nullptr, 0, o, w, s);
}
extern "C" int exb_poly_mul(exb_context *c, uint32_t i, const uint64_t *a, const uint64_t *b, uint64_t *o, size_t w, void *s) {
    return poly_op(c, i, OP_MUL, a, b, 0, o, w, s);
}
extern "C" int exb_poly_scalar_mul(exb_context *c, uint32_t i, const uint64_t *a, uint64_t scalar, uint64_t *o, size_t w, void *s) {
    int rc = check_base(c, i);
    if (rc) return rc;
    return poly_op(c, i, OP_SCALAR_MUL, a, nullptr, scalar % modulus_of(c, i).m, o, w, s);   // ring/ntt.rs:133
}
extern "C" int exb_bfv_add(exb_context *c, const uint64_t *a, const uint64_t *b, uint64_t *o, size_t batch, void *s) {
    if (!c) return fail(EXB_INVALID_PARAM, "null context");
    return poly_op(c, 0, OP_ADD, a, b, 0, o, batch * 2 * (size_t)c->n, s);
}

// ---------------------------------------------------------------------------------
// relinearisation key
// ---------------------------------------------------------------------------------
extern "C" void exb_relin_key_destroy(exb_relin_key *k) {
    if (!k) return;
    if (k->ctx) cudaSetDevice(k->ctx->device);
    cudaFree(k->d_mont);
    delete k;
}

static int relin_key_make(exb_context *c, const u64 *src, bool src_is_host, u32 num_keys, cudaStream_t stream,
                          exb_relin_key **out) {
    if (!c || !out || (!src && num_keys)) return fail(EXB_INVALID_PARAM, "null argument");
    int rc = check_base(c, 0);
    if (rc) return rc;
    EXB_CUDA(cudaSetDevice(c->device));
    exb_relin_key *k = new exb_relin_key();
    k->ctx = c; k->num_keys = num_keys;
    const size_t L = c->ct_moduli.size();
    const size_t words = (size_t)num_keys * 2 * L * c->n;             // multi-prime keys are [G][2][L][n]
    if (L > 1 && !c->rns_enabled) { delete k; return fail(c->mul_status ? c->mul_status : EXB_NOT_IMPLEMENTED, c->mul_error); }
    if (cudaMalloc(&k->d_mont, words ? words * 8 : 8) != cudaSuccess) { delete k; return fail(EXB_CUDA_ERROR, "cudaMalloc relin key"); }
    if (words) {
        cudaError_t e = cudaMemcpyAsync(k->d_mont, src, words * 8, src_is_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, stream);
        if (e != cudaSuccess) { exb_relin_key_destroy(k); return fail(EXB_CUDA_ERROR, cudaGetErrorString(e)); }
        if (L > 1) launch_rns_to_mont(c->R, k->d_mont, (size_t)num_keys * 2 * L, stream);
        else launch_poly_op(c->P.mod[0], OP_TO_MONT, k->d_mont, nullptr, 0, k->d_mont, words, stream);
        rc = check_launch("relin key to Montgomery");
        if (rc) { exb_relin_key_destroy(k); return rc; }
        e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) { exb_relin_key_destroy(k); return fail(EXB_CUDA_ERROR, cudaGetErrorString(e)); }
    }
    *out = k;
    return EXB_OK;
}
extern "C" int exb_relin_key_load(exb_context *c, const uint64_t *rlk_host, uint32_t num_keys, exb_relin_key **out) {
    return relin_key_make(c, rlk_host, true, num_keys, nullptr, out);
}
extern "C" int exb_relin_key_load_device(exb_context *c, const uint64_t *rlk_dev, uint32_t num_keys, void *stream,
                                         exb_relin_key **out) {
    return relin_key_make(c, rlk_dev, false, num_keys, (cudaStream_t)stream, out);
}

// ---------------------------------------------------------------------------------
// dBFV plan: which products / limbs are needed (dbfv/eval.rs:109-114, reduction.rs:28-52)
// ---------------------------------------------------------------------------------
extern "C" int exb_dbfv_small_reps(uint64_t base, uint32_t d, uint64_t pm, int64_t *reps) {
    std::string err;
    int rc = host_small_reps(base, d, pm, reps, &err);
    return rc ? fail(rc, err) : EXB_OK;
}

static int build_plan(u32 d, u64 base, u64 pm, u32 flags, u32 limb_mask, HostPlan *hp) {
    std::string err;
    int rc = host_build_plan(d, base, pm, flags, limb_mask, hp, &err);
    return rc ? fail(rc, err) : EXB_OK;
}

static size_t ws_bytes_per_pair(const exb_context *c, const HostPlan &hp, u32 G, size_t *ext_b, size_t *r01_b,
                                size_t *dig_b, size_t *exc_b) {
    const size_t n = c->n, A = c->aux_moduli.size(), d = hp.M.d;
    *ext_b = 2 * d * 2 * (1 + A) * n * 8;
    if (c->P.sb.enabled && c->logn == 12) *ext_b = 2 * d * 2 * (size_t)c->P.sb.K * n * 4;
    *r01_b = (size_t)hp.M.num_products * 2 * n * 8;
    *dig_b = (size_t)hp.M.num_products * (G ? G : 1) * n * c->digit_bytes();
    *exc_b = (size_t)(hp.M.num_limbs - hp.num_low) * 2 * n * 8;
    return *ext_b + *r01_b + *dig_b + *exc_b;
}

// Run the pipeline for `pairs` pairs whose inputs/outputs are on the device.  The caller holds the slot.
// The n = 4096 kernels move a thread's 8 words with 256-bit accesses: device buffers handed to the multiplication
// entry points must be 32-byte aligned (cudaMalloc / exb_device_alloc / torch allocations are; rows are 32 KiB).
static bool misaligned32(const void *a, const void *b = nullptr, const void *c = nullptr) {
    return (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 31u) != 0;
}
static const char *const kAlignMsg = "device buffers must be 32-byte aligned";

static int run_pairs(exb_context *c, Workspace &w, const HostPlan &hp, const exb_relin_key *rlk, const u64 *ct1,
                     const u64 *ct2, u64 *out, size_t pairs, cudaStream_t stream, bool pipelined = false) {
    if (misaligned32(ct1, ct2, out)) return fail(EXB_INVALID_PARAM, kAlignMsg);
    for (uint32_t p = 0; p < hp.M.num_peers; p++)
        if (misaligned32(hp.M.peer_out[p])) return fail(EXB_INVALID_PARAM, kAlignMsg);
    DeviceParams P = c->P;
    P.pipelined = pipelined ? 1u : 0u;
    const u32 G = rlk->num_keys < c->gadget_digits ? rlk->num_keys : c->gadget_digits;   // keyswitch.rs:86-89
    P.gadget_digits = G;
    size_t eb, rb, db, xb;
    ws_bytes_per_pair(c, hp, G, &eb, &rb, &db, &xb);
    int rc;
    if ((rc = grow((void **)&w.ext, &w.ext_b, eb * pairs))) return rc;
    if ((rc = grow((void **)&w.r01, &w.r01_b, rb * pairs))) return rc;
    if ((rc = grow(&w.digits, &w.digits_b, db * pairs))) return rc;
    if (xb && (rc = grow((void **)&w.excess, &w.excess_b, xb * pairs))) return rc;
    StageEvents se;
    se.has_reduce = false;
    const bool prof = c->profiling;
    if (prof) {
        for (auto &e : se.ev) EXB_CUDA(cudaEventCreate(&e));
        EXB_CUDA(cudaEventRecord(se.ev[0], stream));
    }
    launch_lift(P, hp.M, ct1, ct2, w.ext, pairs, stream);
    if (prof) EXB_CUDA(cudaEventRecord(se.ev[1], stream));
    se.has_c2 = tensor_sums_per_limb(P, hp.M, pairs);
    launch_tensor(P, hp.M, ct1, ct2, w.ext, w.r01, w.digits, c->digit_kind(), pairs, stream, prof ? se.ev[2] : nullptr);
    if (prof) EXB_CUDA(cudaEventRecord(se.ev[3], stream));
    u64 *wide = nullptr;
    if (relin_goes_wide(P, hp.M, pairs)) {
        if ((rc = grow((void **)&w.wide, &w.wide_b, relin_wide_scratch_bytes(P, hp.M, pairs)))) return rc;
        wide = w.wide;
    }
    launch_relin(P, hp.M, w.r01, w.digits, c->digit_kind(), rlk->d_mont, out, w.excess, pairs, stream, wide);
    if (prof) EXB_CUDA(cudaEventRecord(se.ev[4], stream));
    // reduction::reduce for non-zero small representatives (dbfv/reduction.rs:34-52)
    const u32 d = hp.M.d;
    const size_t n = c->n, nx = hp.M.num_limbs - hp.num_low;
    const u64 q = c->ct_moduli[0];
    for (u32 j = d; j + 1 < 2 * d; j++) {
        if (hp.excess_index[j] < 0) continue;
        for (u32 i = 0; i < d; i++) {
            const int64_t rep = hp.reps[(size_t)(j - d) * d + i];
            if (rep == 0) continue;
            bool computed = false;
            for (u32 l = 0; l < hp.M.num_limbs; l++) if (hp.M.limb_k[l] == i) computed = true;
            if (!computed) continue;
            const u64 mag = rep < 0 ? (u64)(-(rep + 1)) + 1 : (u64)rep;
            launch_reduce_mac(P, out + (size_t)i * 2 * n, w.excess + (size_t)hp.excess_index[j] * 2 * n, mag % q,
                              rep < 0, (size_t)d * 2 * n, nx * 2 * n, pairs, stream);
            se.has_reduce = true;
        }
    }
    if (prof) {
        EXB_CUDA(cudaEventRecord(se.ev[5], stream));
        std::lock_guard<std::mutex> g(c->mu);
        c->events.push_back(se);
    }
    return check_launch("ct-mul pipeline");
}

static int mul_precheck(exb_context *c, const exb_relin_key *rlk) {
    if (!c || !rlk) return fail(EXB_INVALID_PARAM, "null argument");
    if (rlk->ctx != c) return fail(EXB_INVALID_PARAM, "relinearisation key belongs to another context");
    if (c->mul_status != EXB_OK) return fail(c->mul_status, c->mul_error);
    return EXB_OK;
}

static size_t device_chunk_pairs(const exb_context *c, const HostPlan &hp, u32 G) {
    size_t eb, rb, db, xb;
    const size_t per = ws_bytes_per_pair(c, hp, G, &eb, &rb, &db, &xb);
    size_t chunk = c->tune.device_chunk_bytes / (per ? per : 1);
    // whole waves: the per-limb kernels (tensor01, relin) run num_limbs CTAs per pair on 2 x #SM resident CTAs, so a
    // chunk that is a multiple of slots / gcd(slots, limbs) pairs leaves no partially filled last wave
    const size_t slots = (size_t)num_sms() * 2, limbs = hp.M.num_limbs ? hp.M.num_limbs : 1;
    size_t a = slots, b = limbs;
    while (b) { const size_t t = a % b; a = b; b = t; }
    const size_t wave = slots / a;
    if (chunk > wave) chunk -= chunk % wave;
    return chunk ? chunk : 1;
}

// Multi-prime ciphertext modulus (bfv_mul_generic_rns, bfv/eval.rs:113-147, and the L > 1 relinearize):
// ct [batch][d][2][L][n].  mode 0: multiply + relinearise + per-k sums; mode 1: bfv_mul_no_relin -> [batch][3][L][n].
static int rns_check_plan(const HostPlan &hp) {
    for (int64_t r : hp.reps)
        if (r != 0)
            return fail(EXB_NOT_IMPLEMENTED, "multi-prime dbfv_mul on the device path needs all-zero small representatives (p = b^d)");
    if (hp.M.num_limbs != hp.num_low)
        return fail(EXB_NOT_IMPLEMENTED, "multi-prime dbfv_mul on the device path computes the limbs k < d only");
    return EXB_OK;
}

// The caller holds workspace `w`; inputs / outputs are on the device.
static int rns_mul_on(exb_context *c, Workspace *w, const HostPlan &hp, const exb_relin_key *rlk, const u64 *ct1,
                      const u64 *ct2, u64 *out, size_t batch, int mode, cudaStream_t st) {
    const u32 G = rlk ? (rlk->num_keys < c->gadget_digits ? rlk->num_keys : c->gadget_digits) : 0;
    const size_t per = rns_workspace_words(c->R, hp.M, 1) * 8;
    size_t chunk = c->tune.device_chunk_bytes / (per ? per : 1);
    if (chunk < 1) chunk = 1;
    const size_t in_stride = (size_t)hp.M.d * 2 * c->R.L * c->n;
    const size_t out_stride = mode == 1 ? (size_t)3 * c->R.L * c->n : in_stride;
    int rc = EXB_OK;
    for (size_t off = 0; off < batch && !rc; off += chunk) {
        const size_t cnt = batch - off < chunk ? batch - off : chunk;
        if ((rc = grow((void **)&w->ext, &w->ext_b, per * cnt))) break;
        launch_rns_mul(c->R, c->T, hp.M, ct1 + off * in_stride, ct2 + off * in_stride, rlk ? rlk->d_mont : nullptr, G,
                       w->ext, out + off * out_stride, cnt, mode, st);
        rc = check_launch("multi-prime ct-mul pipeline");
    }
    return rc;
}

static int rns_mul(exb_context *c, const HostPlan &hp, const exb_relin_key *rlk, const u64 *ct1, const u64 *ct2,
                   u64 *out, size_t batch, int mode, cudaStream_t st) {
    int rc = rns_check_plan(hp);
    if (rc) return rc;
    if (batch == 0) return EXB_OK;
    EXB_CUDA(cudaSetDevice(c->device));
    std::unique_lock<std::mutex> held;
    Workspace *w = nullptr;
    if ((rc = acquire(c, st, &held, &w))) return rc;
    rc = rns_mul_on(c, w, hp, rlk, ct1, ct2, out, batch, mode, st);
    const int rc2 = release(w, st);
    return rc ? rc : rc2;
}

// k-shard exchange on the copy engines: the owned output limbs (runs of consecutive k are one 2-D copy) go from this
// rank's output into every peer's, one stream per peer so the NVLink DMAs run side by side; `st` continues only
// after all of them (fork / join with events), so whatever the caller enqueues next -- its barrier -- is ordered.
static int scatter_limbs(exb_context *c, const HostPlan &hp, const u64 *out, uint64_t *const *peers, uint32_t num_peers,
                         size_t batch, cudaStream_t st) {
    const size_t n = c->n, d = hp.M.d, pitch = d * 2 * n * 8;
    std::lock_guard<std::mutex> g(c->mu);
    for (uint32_t p = 0; p < num_peers; p++)
        if (!c->xs[p]) {
            EXB_CUDA(cudaStreamCreateWithFlags(&c->xs[p], cudaStreamNonBlocking));
            EXB_CUDA(cudaEventCreateWithFlags(&c->xev[p], cudaEventDisableTiming));
        }
    if (!c->xev[kMaxPeers]) EXB_CUDA(cudaEventCreateWithFlags(&c->xev[kMaxPeers], cudaEventDisableTiming));
    EXB_CUDA(cudaEventRecord(c->xev[kMaxPeers], st));
    for (uint32_t p = 0; p < num_peers; p++) {
        EXB_CUDA(cudaStreamWaitEvent(c->xs[p], c->xev[kMaxPeers], 0));
        for (u32 l = 0; l < hp.M.num_low;) {
            u32 e = l;
            while (e + 1 < hp.M.num_low && hp.M.limb_k[e + 1] == hp.M.limb_k[e] + 1) e++;
            const size_t k0 = hp.M.limb_k[l], run = e - l + 1;
            EXB_CUDA(cudaMemcpy2DAsync(peers[p] + k0 * 2 * n, pitch, out + k0 * 2 * n, pitch, run * 2 * n * 8, batch,
                                       cudaMemcpyDefault, c->xs[p]));
            l = e + 1;
        }
        EXB_CUDA(cudaEventRecord(c->xev[p], c->xs[p]));
        EXB_CUDA(cudaStreamWaitEvent(st, c->xev[p], 0));
    }
    return EXB_OK;
}

static int dbfv_mul_device(exb_context *c, uint64_t base, uint32_t d, uint64_t pm, const uint64_t *ct1,
                           const uint64_t *ct2, const exb_relin_key *rlk, uint64_t *out, uint64_t *const *peers,
                           uint32_t num_peers, size_t batch, uint32_t flags, uint32_t limb_mask, void *stream) {
    int rc = mul_precheck(c, rlk);
    if (rc) return rc;
    if (base < 2) return fail(EXB_INVALID_PARAM, "base must be >= 2");
    if (num_peers > (uint32_t)kMaxPeers) return fail(EXB_INVALID_PARAM, "at most 7 peer output buffers");
    if (num_peers && !peers) return fail(EXB_INVALID_PARAM, "null peer list");
    HostPlan hp;
    if ((rc = build_plan(d, base, pm, flags, limb_mask, &hp))) return rc;
    if (c->rns_enabled) {
        if (num_peers) return fail(EXB_NOT_IMPLEMENTED, "k-sharded dbfv_mul needs a single ciphertext prime");
        return rns_mul(c, hp, rlk, ct1, ct2, out, batch, 0, (cudaStream_t)stream);
    }
    if (batch == 0) return EXB_OK;
    EXB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    std::unique_lock<std::mutex> held;
    Workspace *w = nullptr;
    if ((rc = acquire(c, st, &held, &w))) return rc;
    const size_t stride = (size_t)d * 2 * c->n;
    const size_t chunk = device_chunk_pairs(c, hp, c->gadget_digits);
    const bool kernel_stores = num_peers && c->tune.kshard_kernel_stores;
    hp.M.num_peers = kernel_stores ? num_peers : 0;
    for (size_t off = 0; off < batch && !rc; off += chunk) {
        const size_t cnt = batch - off < chunk ? batch - off : chunk;
        for (uint32_t p = 0; p < hp.M.num_peers; p++) hp.M.peer_out[p] = peers[p] + off * stride;
        rc = run_pairs(c, *w, hp, rlk, ct1 + off * stride, ct2 + off * stride, out + off * stride, cnt, st);
    }
    if (!rc && num_peers && !kernel_stores) rc = scatter_limbs(c, hp, out, peers, num_peers, batch, st);
    const int rc2 = release(w, st);
    return rc ? rc : rc2;
}

extern "C" int exb_dbfv_mul(exb_context *c, uint64_t base, uint32_t d, uint64_t pm, const uint64_t *ct1,
                            const uint64_t *ct2, const exb_relin_key *rlk, uint64_t *out, size_t batch,
                            uint32_t flags, uint32_t limb_mask, void *stream) {
    return dbfv_mul_device(c, base, d, pm, ct1, ct2, rlk, out, nullptr, 0, batch, flags, limb_mask, stream);
}

// k-sharded dbfv_mul: this rank computes the output limbs of `limb_mask` and stores each finished limb into its
// own `out` AND into the `num_peers` peer buffers (other GPUs' outputs, mapped with exb_ipc_open) from the
// relinearisation kernel's epilogue -- the gather of dbfv/eval.rs:125-136's per-k sums rides on the compute.
extern "C" int exb_dbfv_mul_scatter(exb_context *c, uint64_t base, uint32_t d, uint64_t pm, const uint64_t *ct1,
                                    const uint64_t *ct2, const exb_relin_key *rlk, uint64_t *out,
                                    uint64_t *const *peer_outs, uint32_t num_peers, size_t batch, uint32_t flags,
                                    uint32_t limb_mask, void *stream) {
    if (c && !c->excess_free(d, base, pm))
        return fail(EXB_NOT_IMPLEMENTED, "k-sharded dbfv_mul needs all-zero small representatives (p = b^d): the general "
                                         "reduction adds excess limbs held by other ranks");
    return dbfv_mul_device(c, base, d, pm, ct1, ct2, rlk, out, peer_outs, num_peers, batch, flags, limb_mask, stream);
}

// CUDA IPC plumbing for the peer buffers (one process per GPU): export the handle of an exb_device_alloc
// allocation, open a peer's handle (peer access is enabled lazily), close it.
extern "C" int exb_ipc_export(exb_context *c, void *dev_ptr, uint8_t handle[64]) {
    if (!c || !dev_ptr || !handle) return fail(EXB_INVALID_PARAM, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    EXB_CUDA(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    EXB_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle, &h, 64);
    return EXB_OK;
}
extern "C" int exb_ipc_open(exb_context *c, const uint8_t handle[64], void **dev_ptr) {
    if (!c || !dev_ptr || !handle) return fail(EXB_INVALID_PARAM, "null argument");
    EXB_CUDA(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    EXB_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return EXB_OK;
}
extern "C" int exb_ipc_close(exb_context *c, void *dev_ptr) {
    if (!c) return fail(EXB_INVALID_PARAM, "null context");
    EXB_CUDA(cudaSetDevice(c->device));
    EXB_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return EXB_OK;
}

extern "C" int exb_bfv_mul_and_relin(exb_context *c, const uint64_t *ct1, const uint64_t *ct2,
                                     const exb_relin_key *rlk, uint64_t *out, size_t batch, void *stream) {
    // one product per pair: the d = 1 case of the same pipeline (base is irrelevant)
    return exb_dbfv_mul(c, 2, 1, 0, ct1, ct2, rlk, out, batch, 0, 0, stream);
}

// ---- host memory for the *_host entry points --------------------------------------------------------
// The reference owns plain Vec<u64> (bfv/mod.rs:19-24).  Page-locked memory is what lets the copies of the
// host-buffer pipeline run asynchronously at PCIe speed: allocate ciphertext storage with exb_host_alloc, or
// pin an existing allocation in place with exb_host_register.
extern "C" int exb_host_alloc_ex(exb_context *c, size_t bytes, uint32_t flags, void **p) {
    if (!c || !p) return fail(EXB_INVALID_PARAM, "null argument");
    if (flags & ~(uint32_t)EXB_HOST_WRITE_COMBINED) return fail(EXB_INVALID_PARAM, "unknown host allocation flag");
    EXB_CUDA(cudaSetDevice(c->device));
    EXB_CUDA(cudaHostAlloc(p, bytes ? bytes : 8,
                           cudaHostAllocPortable | ((flags & EXB_HOST_WRITE_COMBINED) ? cudaHostAllocWriteCombined : 0)));
    return EXB_OK;
}
extern "C" int exb_host_alloc(exb_context *c, size_t bytes, void **p) { return exb_host_alloc_ex(c, bytes, 0, p); }
extern "C" int exb_host_free(exb_context *c, void *p) {
    if (!c) return fail(EXB_INVALID_PARAM, "null context");
    EXB_CUDA(cudaSetDevice(c->device));
    EXB_CUDA(cudaFreeHost(p));
    return EXB_OK;
}
extern "C" int exb_host_register(exb_context *c, void *p, size_t bytes) {
    if (!c || !p) return fail(EXB_INVALID_PARAM, "null argument");
    EXB_CUDA(cudaSetDevice(c->device));
    EXB_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return EXB_OK;
}
extern "C" int exb_host_unregister(exb_context *c, void *p) {
    if (!c || !p) return fail(EXB_INVALID_PARAM, "null argument");
    EXB_CUDA(cudaSetDevice(c->device));
    EXB_CUDA(cudaHostUnregister(p));
    return EXB_OK;
}

// ---- host-buffer pipeline -------------------------------------------------------------------------------
// A call is cut into chunks; chunk i goes to host slot (host_rr + i) % kHostSlots, whose stream runs
// H2D(ct1, ct2) -> kernels -> D2H(out).  Slots keep rotating ACROSS calls, so with the asynchronous entry the
// next call's first H2D overlaps this call's last kernels and D2H: the copy engines and the SMs stay busy and
// only the very first fill and the very last drain of a sequence of calls are exposed.
static int host_pipeline(exb_context *c, uint64_t base, uint32_t d, uint64_t pm, const uint64_t *ct1,
                         const uint64_t *ct2, const exb_relin_key *rlk, uint64_t *out, size_t batch, uint32_t flags,
                         bool taper, Ticket *tk) {
    int rc = mul_precheck(c, rlk);
    if (rc) return rc;
    if (base < 2) return fail(EXB_INVALID_PARAM, "base must be >= 2");
    HostPlan hp;
    if ((rc = build_plan(d, base, pm, flags, 0, &hp))) return rc;
    if (c->rns_enabled && (rc = rns_check_plan(hp))) return rc;
    if (batch == 0) return EXB_OK;
    std::lock_guard<std::mutex> lock(c->host_mu);
    EXB_CUDA(cudaSetDevice(c->device));
    if (c->rns_enabled) {
        // multi-prime ciphertext modulus: [batch][d][2][L][n], staged through one slot in bounded chunks
        Workspace &w = c->hs[0];
        const size_t stride = (size_t)d * 2 * c->R.L * c->n;
        size_t chunk = ((size_t)256 << 20) / (stride * 8);
        if (chunk < 1) chunk = 1;
        for (size_t off = 0; off < batch; off += chunk) {
            const size_t cnt = batch - off < chunk ? batch - off : chunk, bytes = cnt * stride * 8;
            if ((rc = grow((void **)&w.in1, &w.in_b, chunk * stride * 8))) return rc;
            if ((rc = grow_in2_out(w, chunk * stride * 8))) return rc;
            EXB_CUDA(cudaMemcpyAsync(w.in1, ct1 + off * stride, bytes, cudaMemcpyHostToDevice, w.stream));
            EXB_CUDA(cudaMemcpyAsync(w.in2, ct2 + off * stride, bytes, cudaMemcpyHostToDevice, w.stream));
            if ((rc = rns_mul_on(c, &w, hp, rlk, w.in1, w.in2, w.out, cnt, 0, w.stream))) return rc;
            EXB_CUDA(cudaMemcpyAsync(out + off * stride, w.out, bytes, cudaMemcpyDeviceToHost, w.stream));
        }
        cudaEvent_t e;
        EXB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        tk->events.push_back(e);
        EXB_CUDA(cudaEventRecord(e, w.stream));
        return EXB_OK;
    }
    const size_t stride = (size_t)d * 2 * c->n;
    // chunk size: enough CTAs to fill the GPU, small enough that H2D / kernels / D2H of consecutive chunks overlap
    const size_t chunk_products = c->tune.host_chunk_products ? c->tune.host_chunk_products : (taper ? 1024 : 3996);
    size_t chunk = chunk_products / (hp.M.num_products ? hp.M.num_products : 1);
    if (chunk < 1) chunk = 1;
    const int slots = c->tune.host_slots;
    if (batch < chunk * (size_t)(slots - 1)) chunk = (batch + slots - 2) / (size_t)(slots - 1);
    bool used[kHostSlots] = {};
    auto enqueue = [&]() -> int {
        size_t ci = 0;
        for (size_t off = 0; off < batch; ci++) {
            const int si = (int)(c->host_rr++ % (unsigned)slots);
            Workspace &w = c->hs[si];
            const size_t left = batch - off;
            size_t cnt = chunk;
            if (taper) {
                // a synchronous call exposes the first chunk's H2D and the last chunk's kernels + D2H: taper both
                // ends (half-size first chunk, remainder split over the last two)
                if (ci == 0 && batch > 2 * chunk) cnt = (chunk + 1) / 2;
                else if (left <= chunk) cnt = left;
                else if (left < 2 * chunk) cnt = (left + 1) / 2 + (left + 1) / 8;
            }
            if (cnt > left) cnt = left;
            if (cnt > chunk) cnt = chunk;
            const size_t bytes = cnt * stride * 8;
            int r;
            if ((r = grow((void **)&w.in1, &w.in_b, chunk * stride * 8))) return r;
            if ((r = grow_in2_out(w, chunk * stride * 8))) return r;
            used[si] = true;
            EXB_CUDA(cudaMemcpyAsync(w.in1, ct1 + off * stride, bytes, cudaMemcpyHostToDevice, w.stream));
            EXB_CUDA(cudaMemcpyAsync(w.in2, ct2 + off * stride, bytes, cudaMemcpyHostToDevice, w.stream));
            if ((r = run_pairs(c, w, hp, rlk, w.in1, w.in2, w.out, cnt, w.stream, true))) return r;
            EXB_CUDA(cudaMemcpyAsync(out + off * stride, w.out, bytes, cudaMemcpyDeviceToHost, w.stream));
            off += cnt;
        }
        for (int si = 0; si < kHostSlots; si++) {
            if (!used[si]) continue;
            cudaEvent_t e;
            EXB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            tk->events.push_back(e);
            EXB_CUDA(cudaEventRecord(e, c->hs[si].stream));
        }
        return EXB_OK;
    };
    rc = enqueue();
    if (rc) {   // nothing of a failed call may still be reading the caller's buffers when we return
        const std::string msg = g_err;
        for (int si = 0; si < kHostSlots; si++) if (used[si]) cudaStreamSynchronize(c->hs[si].stream);
        g_err = msg;
    }
    return rc;
}

static int wait_ticket(Ticket &tk) {
    cudaError_t first = cudaSuccess;
    for (cudaEvent_t e : tk.events) {
        const cudaError_t r = cudaEventSynchronize(e);
        if (r != cudaSuccess && first == cudaSuccess) first = r;
        cudaEventDestroy(e);
    }
    tk.events.clear();
    if (first != cudaSuccess) return fail(EXB_CUDA_ERROR, std::string("host pipeline: ") + cudaGetErrorString(first));
    return EXB_OK;
}

extern "C" int exb_dbfv_mul_host(exb_context *c, uint64_t base, uint32_t d, uint64_t pm, const uint64_t *ct1,
                                 const uint64_t *ct2, const exb_relin_key *rlk, uint64_t *out, size_t batch,
                                 uint32_t flags) {
    Ticket tk;
    const int rc = host_pipeline(c, base, d, pm, ct1, ct2, rlk, out, batch, flags, true, &tk);
    const int rc2 = wait_ticket(tk);
    return rc ? rc : rc2;
}

extern "C" int exb_dbfv_mul_host_async(exb_context *c, uint64_t base, uint32_t d, uint64_t pm, const uint64_t *ct1,
                                       const uint64_t *ct2, const exb_relin_key *rlk, uint64_t *out, size_t batch,
                                       uint32_t flags, exb_ticket *ticket) {
    if (!ticket) return fail(EXB_INVALID_PARAM, "null ticket");
    *ticket = 0;
    Ticket tk;
    const int rc = host_pipeline(c, base, d, pm, ct1, ct2, rlk, out, batch, flags, false, &tk);
    if (rc) { wait_ticket(tk); return rc; }
    std::lock_guard<std::mutex> g(c->mu);
    *ticket = c->next_ticket++;
    c->tickets.emplace(*ticket, std::move(tk));
    return EXB_OK;
}

extern "C" int exb_bfv_mul_and_relin_host(exb_context *c, const uint64_t *ct1, const uint64_t *ct2,
                                          const exb_relin_key *rlk, uint64_t *out, size_t batch) {
    return exb_dbfv_mul_host(c, 2, 1, 0, ct1, ct2, rlk, out, batch, 0);
}

extern "C" int exb_bfv_mul_and_relin_host_async(exb_context *c, const uint64_t *ct1, const uint64_t *ct2,
                                                const exb_relin_key *rlk, uint64_t *out, size_t batch,
                                                exb_ticket *ticket) {
    return exb_dbfv_mul_host_async(c, 2, 1, 0, ct1, ct2, rlk, out, batch, 0, ticket);
}

// Block until the call behind `ticket` has delivered its output to host memory.  A ticket is waited once.
extern "C" int exb_wait(exb_context *c, exb_ticket ticket) {
    if (!c) return fail(EXB_INVALID_PARAM, "null context");
    if (ticket == 0) return EXB_OK;                      // a call that had nothing to do
    Ticket tk;
    {
        std::lock_guard<std::mutex> g(c->mu);
        auto it = c->tickets.find(ticket);
        if (it == c->tickets.end()) return fail(EXB_INVALID_PARAM, "unknown or already waited ticket");
        tk = std::move(it->second);
        c->tickets.erase(it);
    }
    return wait_ticket(tk);
}

// ---- Galois automorphism + key switch (bfv/eval.rs:512-561) -------------------------------------
static int galois_precheck(exb_context *c, const exb_relin_key *gk, uint64_t element) {
    if (!c || !gk) return fail(EXB_INVALID_PARAM, "null argument");
    if (gk->ctx != c) return fail(EXB_INVALID_PARAM, "Galois key belongs to another context");
    if (c->ct_moduli.size() != 1) return fail(EXB_NOT_IMPLEMENTED, "automorphism on the device path needs a single ciphertext prime");
    if (!(element & 1))   // sigma_k is a ring automorphism only for odd k (bfv/keygen.rs:216-217)
        return fail(EXB_INVALID_PARAM, "Galois element must be odd");
    if (gk->num_keys < c->gadget_digits)
        return fail(EXB_INVALID_PARAM, "Galois key holds fewer than gadget_digits key pairs");
    return EXB_OK;
}

extern "C" int exb_bfv_apply_automorphism(exb_context *c, const uint64_t *ct, uint64_t element,
                                          const exb_relin_key *gk, uint64_t *out, size_t batch, void *stream) {
    int rc = galois_precheck(c, gk, element);
    if (rc) return rc;
    if (batch == 0) return EXB_OK;
    const size_t words = batch * 2 * (size_t)c->n;
    if (ct < out + words && out < ct + words) return fail(EXB_INVALID_PARAM, "automorphism output must not alias its input");
    EXB_CUDA(cudaSetDevice(c->device));
    launch_galois(c->P, ct, gk->d_mont, (u32)(element % (2 * (uint64_t)c->n)), out, batch, (cudaStream_t)stream);
    return check_launch("galois");
}

extern "C" int exb_bfv_apply_automorphism_host(exb_context *c, const uint64_t *ct, uint64_t element,
                                               const exb_relin_key *gk, uint64_t *out, size_t batch) {
    int rc = galois_precheck(c, gk, element);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(c->host_mu);
    if (batch == 0) return EXB_OK;
    EXB_CUDA(cudaSetDevice(c->device));
    const size_t stride = 2 * (size_t)c->n;
    size_t chunk = 1024;                                   // 64 MiB of ciphertexts at n = 4096
    if (batch < chunk * kHostSlots) chunk = (batch + kHostSlots - 1) / kHostSlots;
    size_t ci = 0;
    for (size_t off = 0; off < batch; off += chunk, ci++) {
        Workspace &w = c->hs[ci % kHostSlots];
        const size_t cnt = batch - off < chunk ? batch - off : chunk;
        if ((rc = grow((void **)&w.in1, &w.in_b, chunk * stride * 8))) return rc;
        if ((rc = grow_in2_out(w, chunk * stride * 8))) return rc;
        EXB_CUDA(cudaMemcpyAsync(w.in1, ct + off * stride, cnt * stride * 8, cudaMemcpyHostToDevice, w.stream));
        launch_galois(c->P, w.in1, gk->d_mont, (u32)(element % (2 * (uint64_t)c->n)), w.out, cnt, w.stream);
        if ((rc = check_launch("galois"))) return rc;
        EXB_CUDA(cudaMemcpyAsync(out + off * stride, w.out, cnt * stride * 8, cudaMemcpyDeviceToHost, w.stream));
    }
    for (Workspace &w : c->hs) EXB_CUDA(cudaStreamSynchronize(w.stream));
    return EXB_OK;
}

// ---- decrypt (bfv/encrypt.rs:111-178) -------------------------------------------------------------
static int decrypt_precheck(exb_context *c, uint32_t ncomp) {
    if (!c) return fail(EXB_INVALID_PARAM, "null argument");
    if (ncomp < 1) return fail(EXB_INVALID_PARAM, "ciphertext has no components");
    if (c->ct_moduli.size() != 1) {
        if (!c->rns_enabled) return fail(c->mul_status ? c->mul_status : EXB_NOT_IMPLEMENTED, c->mul_error);
        return EXB_OK;                                    // multi-prime: BigUint CRT of bfv/encrypt.rs:136-170 on the device
    }
    if (c->plain >= c->ct_moduli[0]) return fail(EXB_NOT_IMPLEMENTED, "decrypt on the device path needs plain_modulus < q");
    return EXB_OK;
}

extern "C" int exb_bfv_decrypt(exb_context *c, const uint64_t *ct, uint32_t ncomp, const uint64_t *sk_ntt,
                               uint64_t *out, size_t batch, void *stream) {
    int rc = decrypt_precheck(c, ncomp);
    if (rc) return rc;
    if (batch == 0) return EXB_OK;
    EXB_CUDA(cudaSetDevice(c->device));
    if (c->rns_enabled) {
        cudaStream_t st = (cudaStream_t)stream;
        std::unique_lock<std::mutex> held;
        Workspace *w = nullptr;
        if ((rc = acquire(c, st, &held, &w))) return rc;
        rc = grow((void **)&w->ext, &w->ext_b, batch * c->R.L * c->n * 8);
        if (!rc) {
            launch_rns_decrypt(c->R, c->T, ct, ncomp, sk_ntt, w->ext, out, batch, st);
            rc = check_launch("multi-prime decrypt");
        }
        const int rc2 = release(w, st);
        return rc ? rc : rc2;
    }
    launch_decrypt(c->P, ct, ncomp, sk_ntt, out, batch, (cudaStream_t)stream);
    return check_launch("decrypt");
}

extern "C" int exb_bfv_decrypt_host(exb_context *c, const uint64_t *ct, uint32_t ncomp, const uint64_t *sk_ntt,
                                    uint64_t *out, size_t batch) {
    int rc = decrypt_precheck(c, ncomp);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(c->host_mu);
    if (batch == 0) return EXB_OK;
    EXB_CUDA(cudaSetDevice(c->device));
    Workspace &w = c->hs[0];
    const size_t Lq = c->ct_moduli.size();
    const size_t n = c->n, in_bytes = batch * ncomp * Lq * n * 8, out_bytes = batch * n * 8, sk_bytes = Lq * n * 8;
    if ((rc = grow((void **)&w.in1, &w.in_b, in_bytes))) return rc;
    if ((rc = grow_in2_out(w, out_bytes > sk_bytes ? out_bytes : sk_bytes))) return rc;
    EXB_CUDA(cudaMemcpyAsync(w.in1, ct, in_bytes, cudaMemcpyHostToDevice, w.stream));
    EXB_CUDA(cudaMemcpyAsync(w.in2, sk_ntt, sk_bytes, cudaMemcpyHostToDevice, w.stream));
    if (c->rns_enabled) {
        if ((rc = grow((void **)&w.ext, &w.ext_b, batch * Lq * n * 8))) return rc;
        launch_rns_decrypt(c->R, c->T, w.in1, ncomp, w.in2, w.ext, w.out, batch, w.stream);
    } else
    launch_decrypt(c->P, w.in1, ncomp, w.in2, w.out, batch, w.stream);
    if ((rc = check_launch("decrypt"))) return rc;
    EXB_CUDA(cudaMemcpyAsync(out, w.out, out_bytes, cudaMemcpyDeviceToHost, w.stream));
    EXB_CUDA(cudaStreamSynchronize(w.stream));
    return EXB_OK;
}

// ---- bfv_mul_no_relin (bfv/eval.rs:89-108), relinearize (bfv/keyswitch.rs:59-101), gadget_decompose (:11-52)
// as stand-alone entry points (the fused pipeline above is what bfv_mul_and_relin / dbfv_mul use) ----------
extern "C" int exb_bfv_mul_no_relin(exb_context *c, const uint64_t *ct1, const uint64_t *ct2, uint64_t *out3,
                                    size_t batch, void *stream) {
    if (!c) return fail(EXB_INVALID_PARAM, "null argument");
    if (c->mul_status != EXB_OK) return fail(c->mul_status, c->mul_error);
    HostPlan hp;
    int rc = build_plan(1, 2, 0, 0, 0, &hp);
    if (rc) return rc;
    if (c->rns_enabled) return rns_mul(c, hp, nullptr, ct1, ct2, out3, batch, 1, (cudaStream_t)stream);
    if (batch == 0) return EXB_OK;
    if (misaligned32(ct1, ct2, out3)) return fail(EXB_INVALID_PARAM, kAlignMsg);
    EXB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    std::unique_lock<std::mutex> held;
    Workspace *wp = nullptr;
    if ((rc = acquire(c, st, &held, &wp))) return rc;
    Workspace &w = *wp;
    const size_t n = c->n;
    size_t eb, rb, db, xb;
    ws_bytes_per_pair(c, hp, c->gadget_digits, &eb, &rb, &db, &xb);
    const size_t per = eb + 3 * n * 8;
    size_t chunk = ((size_t)4 << 30) / per;
    if (chunk < 1) chunk = 1;
    for (size_t off = 0; off < batch; off += chunk) {
        const size_t cnt = batch - off < chunk ? batch - off : chunk;
        if ((rc = grow((void **)&w.ext, &w.ext_b, eb * cnt))) return rc;
        if ((rc = grow((void **)&w.r01, &w.r01_b, 3 * n * 8 * cnt))) return rc;
        launch_lift(c->P, hp.M, ct1 + off * 2 * n, ct2 + off * 2 * n, w.ext, cnt, st);
        launch_tensor(c->P, hp.M, ct1 + off * 2 * n, ct2 + off * 2 * n, w.ext, w.r01, nullptr, c->digit_kind(), cnt, st, nullptr, true);
        launch_ntt_fwd(c->P, 0, w.r01, out3 + off * 3 * n, 3 * cnt, st);      // hps_scale ends with from_coeff_poly (:412)
    }
    if ((rc = release(&w, st))) return rc;
    return check_launch("bfv_mul_no_relin");
}

extern "C" int exb_bfv_relinearize(exb_context *c, const uint64_t *ct, uint32_t ncomp, const exb_relin_key *rlk,
                                   uint64_t *out, size_t batch, void *stream) {
    if (!c || !rlk) return fail(EXB_INVALID_PARAM, "null argument");
    if (rlk->ctx != c) return fail(EXB_INVALID_PARAM, "relinearisation key belongs to another context");
    if (ncomp > 3) return fail(EXB_INVALID_PARAM, "relinearization only supports degree-2 ciphertexts");   // :66-70
    if (c->ct_moduli.size() != 1 && !c->rns_enabled) return fail(c->mul_status ? c->mul_status : EXB_NOT_IMPLEMENTED, c->mul_error);
    if (batch == 0) return EXB_OK;
    EXB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = c->n * c->ct_moduli.size();                               // words per component
    if (ncomp < 3) {                                                           // :63-65 already degree 1: unchanged
        if (out != ct) EXB_CUDA(cudaMemcpyAsync(out, ct, batch * ncomp * n * 8, cudaMemcpyDeviceToDevice, st));
        return EXB_OK;
    }
    if (c->rns_enabled) {
        HostPlan hp1;
        int rc1 = build_plan(1, 2, 0, 0, 0, &hp1);
        if (rc1) return rc1;
        std::unique_lock<std::mutex> held1;
        Workspace *w1 = nullptr;
        if ((rc1 = acquire(c, st, &held1, &w1))) return rc1;
        const u32 G1 = rlk->num_keys < c->gadget_digits ? rlk->num_keys : c->gadget_digits;
        const size_t per = rns_relin_workspace_words(c->R, 1) * 8;
        size_t chunk1 = c->tune.device_chunk_bytes / per;
        if (chunk1 < 1) chunk1 = 1;
        for (size_t off = 0; off < batch && !rc1; off += chunk1) {
            const size_t cnt = batch - off < chunk1 ? batch - off : chunk1;
            if ((rc1 = grow((void **)&w1->ext, &w1->ext_b, per * cnt))) break;
            launch_rns_relinearize(c->R, c->T, hp1.M, ct + off * 3 * n, rlk->d_mont, G1, w1->ext, out + off * 2 * n, cnt, st);
            rc1 = check_launch("multi-prime relinearize");
        }
        const int rc2 = release(w1, st);
        return rc1 ? rc1 : rc2;
    }
    if (c->gadget_base < 2 || c->gadget_base > (1ull << 32))
        return fail(EXB_NOT_IMPLEMENTED, "device path supports gadget bases in [2, 2^32]");
    if (misaligned32(ct, out)) return fail(EXB_INVALID_PARAM, kAlignMsg);
    HostPlan hp;
    int rc = build_plan(1, 2, 0, 0, 0, &hp);
    if (rc) return rc;
    DeviceParams P = c->P;
    const u32 G = rlk->num_keys < c->gadget_digits ? rlk->num_keys : c->gadget_digits;                    // :86-89
    P.gadget_digits = G;
    std::unique_lock<std::mutex> held;
    Workspace *wp = nullptr;
    if ((rc = acquire(c, st, &held, &wp))) return rc;
    Workspace &w = *wp;
    const size_t dig_b = (size_t)(G ? G : 1) * n * c->digit_bytes();
    const size_t per = n * 8 + 2 * n * 8 + dig_b + (size_t)(G + 1) * 2 * n * 8;
    size_t chunk = ((size_t)4 << 30) / per;
    if (chunk < 1) chunk = 1;
    for (size_t off = 0; off < batch; off += chunk) {
        const size_t cnt = batch - off < chunk ? batch - off : chunk;
        if ((rc = grow((void **)&w.ext, &w.ext_b, cnt * n * 8))) return rc;    // c2 in the coefficient domain
        if ((rc = grow((void **)&w.r01, &w.r01_b, cnt * 2 * n * 8))) return rc;
        if ((rc = grow(&w.digits, &w.digits_b, cnt * dig_b))) return rc;
        const uint64_t *src = ct + off * 3 * n;
        EXB_CUDA(cudaMemcpy2DAsync(w.r01, 2 * n * 8, src, 3 * n * 8, 2 * n * 8, cnt, cudaMemcpyDeviceToDevice, st));
        EXB_CUDA(cudaMemcpy2DAsync(w.ext, n * 8, src + 2 * n, 3 * n * 8, n * 8, cnt, cudaMemcpyDeviceToDevice, st));
        launch_ntt_inv(P, 0, w.ext, w.ext, cnt, st);                           // :76
        launch_gadget_digits(P, w.ext, w.digits, c->digit_kind() == 2 ? 3 : c->digit_kind(), cnt, st);   // :79
        u64 *wide = nullptr;
        if (relin_goes_wide(P, hp.M, cnt)) {
            if ((rc = grow((void **)&w.wide, &w.wide_b, relin_wide_scratch_bytes(P, hp.M, cnt)))) return rc;
            wide = w.wide;
        }
        launch_relin(P, hp.M, w.r01, w.digits, c->digit_kind(), rlk->d_mont, out + off * 2 * n, nullptr, cnt, st, wide, true);
    }
    if ((rc = release(&w, st))) return rc;
    return check_launch("relinearize");
}

extern "C" int exb_gadget_decompose(exb_context *c, const uint64_t *coeffs, uint64_t *out, size_t count, void *stream) {
    if (!c) return fail(EXB_INVALID_PARAM, "null argument");
    if (c->gadget_base < 2 || c->gadget_base > (1ull << 32))
        return fail(EXB_NOT_IMPLEMENTED, "device path supports gadget bases in [2, 2^32]");
    if (count == 0) return EXB_OK;
    EXB_CUDA(cudaSetDevice(c->device));
    launch_gadget_digits(c->P, coeffs, out, 2, count, (cudaStream_t)stream);
    return check_launch("gadget_decompose");
}
