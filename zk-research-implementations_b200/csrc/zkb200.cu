// zkb200.cu -- host driver and C ABI (include/zkb200.h) of the B200 sumcheck /
// GKR prover engine.  All table arithmetic runs in the kernels of kernels.cuh;
// this file owns device memory, the round loop, the mailbox through which the
// (d+1) round evaluations reach the host, the host transcript, and the optional
// NCCL communicator for tables sharded on low index bits.
//
// There is no CPU fallback: every table entry point needs a ctx, and a ctx can
// only be created on a CUDA device.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/zkb200.h"
#include "host_math.hpp"
#include "kzg_impl.cuh"

using namespace zkb;

// ------------------------------------------------------------------ host portability (x86-64 and aarch64 hosts)
namespace {
inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#elif defined(__aarch64__)
    asm volatile("yield" ::: "memory");
#endif
}
// A monotonic tick counter for the trace (ZKB200_TRACE); ticks_per_us() calibrates it against CLOCK_MONOTONIC.
inline unsigned long long host_ticks() {
#if defined(__x86_64__) || defined(__i386__)
    return __builtin_ia32_rdtsc();
#elif defined(__aarch64__)
    unsigned long long v;
    asm volatile("mrs %0, cntvct_el0" : "=r"(v));
    return v;
#else
    timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (unsigned long long)t.tv_sec * 1000000000ull + (unsigned long long)t.tv_nsec;
#endif
}
// Flags and sequence numbers written by the device into mapped host memory (or by the host for the device) are read
// with acquire / written with release semantics: on a weakly ordered host (Grace) a plain load of the payload could
// otherwise be satisfied before the load of the flag that guards it.
inline unsigned int load_acquire(const volatile unsigned int* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void store_release(volatile unsigned int* p, unsigned int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
}  // namespace

// ------------------------------------------------------------------ NCCL (lazy)
// Resolved with dlopen at zkb_ctx_comm_init so that (a) the library loads on a
// machine without NCCL/GPU and (b) a process that already carries NCCL (torch)
// shares that copy instead of loading a second one.
namespace {
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (h) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
        for (int i = 0; names[i] && !h; ++i) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if (!h) return false;
        GetUniqueId = (decltype(GetUniqueId))dlsym(h, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(h, "ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(h, "ncclAllReduce");
        AllGather = (decltype(AllGather))dlsym(h, "ncclAllGather");
        GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
        return GetUniqueId && CommInitRank && CommDestroy && AllReduce && AllGather;
    }
};
NcclApi g_nccl;

const FieldKernels* kernels_for(int field) {
    switch (field) {
        case ZKB_FIELD_BN254_FR: return field_kernels_bn254_fr();
        case ZKB_FIELD_BN254_FQ: return field_kernels_bn254_fq();
        case ZKB_FIELD_BLS12_381_FR: return field_kernels_bls12_381_fr();
    }
    return nullptr;
}

inline Fe fe_from_u64x4(const uint64_t* p) {
    Fe r;
    std::memcpy(r.l, p, 32);
    return r;
}
inline void fe_to_u64x4(const Fe& v, uint64_t* p) { std::memcpy(p, v.l, 32); }
inline int ilog2_u64(uint64_t x) {
    int r = 0;
    while (x > 1) {
        x >>= 1;
        ++r;
    }
    return r;
}
inline bool is_pow2(uint64_t x) { return x && !(x & (x - 1)); }

// A device table in the planar layout of kernels.cuh.
struct Table {
    uint4* base = nullptr;
    uint64_t n = 0;       // live entries
    uint64_t stride = 0;  // distance between the two limb planes (uint4 units) = allocated entries
    int refs = 0;         // SumPolys that read this table (zkb_sumpoly_create .. zkb_sumpoly_free)
    bool dead = false;    // zkb_mle_free was called while refs > 0: the memory goes when the last SumPoly does
    TabRef ref() const { return TabRef{base, stride}; }
};

struct SumPolyState {
    int P = 0, D = 0;             // declared shape (SumPoly of P ProductPolys of D factors)
    int kind = KIND_PROD;
    int kP = 0, kD = 0, npts = 0; // what the round kernel multiplies / how many points it returns
    std::vector<int> sel;         // tables the round kernel reads (kernel order)
    std::vector<int> rest;        // tables that are only folded (compat mode)
    std::vector<Table> src;       // caller's tables, never written
    std::vector<uint64_t> src_handles;  // their handles (reference counts, see Table::refs)
    std::vector<Table> work;      // private folded copies
    std::vector<Table> gath;      // after the multi-GPU gather
    uint64_t n0 = 0;              // local entries per table before any bind
    uint64_t cur_n = 0;           // local entries per table now
    Fe last_evals[MAXPTS];        // s(0..d) of the current round (the claim chain: s(1) = claim - s(0))
    bool have_evals = false;
    int state = 0;                // 0 = unbound (src), 1 = in work, 2 = in gath
    bool sharded = false;         // partial sums need the allreduce
    Table& cur(int t) { return state == 0 ? src[t] : (state == 1 ? work[t] : gath[t]); }
};

struct CircuitState {
    int L = 0;
    std::vector<uint32_t> gates;  // input side first
    std::vector<size_t> opoff;
    std::vector<uint8_t> h_ops;
    uint8_t* d_ops = nullptr;
    Table inputs;
    std::vector<Table> vals;  // per layer outputs
    Table H1, HA2, coef;      // phase tables (max size)
    Table gtmp;               // general wiring: one per-gate product (coef*W[in2] / coef*eq(u,in1))
    Table eq[8];              // split eq tables: rb hi/lo, rc hi/lo, u hi/lo, w hi/lo
    SumPolyState sp;          // the XYZ sumcheck state (work tables reused across layers)
    void* aos_stage = nullptr;
    size_t aos_stage_bytes = 0;
    // general wiring (zkb_circuit_create_wired; an extension beyond the reference's fixed (2i, 2i+1) wiring)
    bool wired = false;
    uint64_t n_inputs = 0;
    std::vector<uint64_t> width;  // wires below each layer
    std::vector<size_t> csroff;   // offset of each layer's CSR row pointers (width + 1 each)
    uint32_t *d_in1 = nullptr, *d_in2 = nullptr, *d_lst1 = nullptr, *d_lst2 = nullptr, *d_off1 = nullptr, *d_off2 = nullptr;
};
}  // namespace


// ------------------------------------------------ intra-node exchange of round sums
// The per-round combine of the ranks' partial sums is latency-critical and tiny ((d+1) x 32 bytes per
// rank), and the host needs the result anyway for the transcript.  On one box (the deployment this engine
// targets: the 8 GPUs of a node) the ranks' host threads therefore exchange the sums through a POSIX
// shared-memory segment with sequence numbers (about 1 us), and NCCL over NVLink carries the bulk step
// (the all-gather of the shrunken tables).  If the segment cannot be opened the engine falls back to one
// ncclAllReduce per round over zero-extended limbs.
struct ShmSlot {
    volatile uint64_t seq;  // rounds posted by this rank
    uint64_t pad[7];
    uint64_t v[2][MAXPTS][4];
    uint64_t pad2[(512 - 64 - 2 * MAXPTS * 32) / 8];
};
static_assert(sizeof(ShmSlot) == 512, "one slot = 512 bytes");
struct ShmHeader {
    volatile uint32_t attached;
    uint32_t pad[127];
};
struct ShmComm {
    ShmHeader* hdr = nullptr;
    ShmSlot* slots = nullptr;
    size_t bytes = 0;
    uint64_t round = 0;
    int rank = 0, world = 1;
    bool open(const uint8_t id[128], int rank_, int world_) {
        rank = rank_;
        world = world_;
        char name[64];
        unsigned long long h = 1469598103934665603ull;
        for (int i = 0; i < 128; ++i) h = (h ^ id[i]) * 1099511628211ull;
        snprintf(name, sizeof name, "/zkb200_%016llx", h);
        bytes = sizeof(ShmHeader) + sizeof(ShmSlot) * (size_t)world;
        int fd = -1;
        if (rank == 0) {
            shm_unlink(name);
            fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
            if (fd < 0 || ftruncate(fd, (off_t)bytes) != 0) {
                if (fd >= 0) close(fd);
                return false;
            }
        } else {
            for (int tries = 0; tries < 20000 && fd < 0; ++tries) {  // up to ~20 s for rank 0 to create it
                fd = shm_open(name, O_RDWR, 0600);
                struct stat st;
                if (fd >= 0 && (fstat(fd, &st) != 0 || (size_t)st.st_size < bytes)) {
                    close(fd);
                    fd = -1;
                }
                if (fd < 0) usleep(1000);
            }
            if (fd < 0) return false;
        }
        void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (p == MAP_FAILED) return false;
        hdr = (ShmHeader*)p;
        slots = (ShmSlot*)((uint8_t*)p + sizeof(ShmHeader));
        __atomic_add_fetch(&hdr->attached, 1u, __ATOMIC_SEQ_CST);
        for (long tries = 0; hdr->attached < (uint32_t)world; ++tries) {
            if (tries > 20000) return false;
            usleep(1000);
        }
        // every rank has it mapped: the name can go (the memory lives until the last unmap).  Rank 0 waits a
        // little so that no rank is still between shm_open and mmap -- they all passed `attached` already.
        if (rank == 0) shm_unlink(name);
        return true;
    }
    void close_() {
        if (hdr) munmap((void*)hdr, bytes);
        hdr = nullptr;
        slots = nullptr;
    }
    // vals[0..n) of every rank are added (mod p, in rank order, so every rank gets identical bits)
    bool allreduce(const HostField& H, Fe* vals, int n) {
        const uint64_t k = ++round;
        ShmSlot& mine = slots[rank];
        for (int i = 0; i < n; ++i) std::memcpy((void*)mine.v[k & 1][i], vals[i].l, 32);
        __atomic_store_n(&mine.seq, k, __ATOMIC_RELEASE);
        Fe acc[MAXPTS];
        for (int i = 0; i < n; ++i) acc[i] = H.zero();
        for (int r = 0; r < world; ++r) {
            ShmSlot& s = slots[r];
            uint64_t spins = 0;
            while (__atomic_load_n(&s.seq, __ATOMIC_ACQUIRE) < k) {
                if (++spins > (1ull << 33)) return false;  // a peer died
                cpu_relax();
            }
            for (int i = 0; i < n; ++i) {
                Fe v;
                std::memcpy(v.l, (const void*)s.v[k & 1][i], 32);
                acc[i] = H.add(acc[i], v);
            }
        }
        for (int i = 0; i < n; ++i) vals[i] = acc[i];
        return true;
    }
};

// Multilinear KZG setup (G1 side): the Lagrange basis and its folds, affine, on the device.
struct KzgState {
    uint32_t n_vars = 0;
    std::vector<G1Affine*> basis;  // level k: 2^(n_vars - k) points; level 0 = g1 * eq(taus, .)
};

// Keccak Merkle tree (merkle_tree/src/merkle_tree.rs:24-29): leaves then levels in one device buffer of Montgomery residues
struct MerkleState {
    uint32_t depth = 0;
    Fe* tree = nullptr;
};

struct zkb_transcript {
    TranscriptImpl impl;
};

struct zkb_ctx {
    int field = 0, device = 0, mode = 0;
    const FieldKernels* K = nullptr;
    HostField H;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    uint64_t launches = 0;
    std::string last_error;
    // mailbox (mapped, pinned host memory)
    Fe* h_res = nullptr;
    volatile unsigned int* h_flag = nullptr;
    unsigned int seq = 0;
    // device scratch of the reducing kernels
    Fe* d_partials = nullptr;
    size_t partials_cap = 0;
    unsigned int* d_ticket = nullptr;
    Fe* d_res = nullptr;
    unsigned long long* d_wide = nullptr;
    unsigned long long* h_wide = nullptr;
    // staging for uploads / downloads
    void* stage = nullptr;
    size_t stage_bytes = 0;
    void* hash_pin = nullptr;  // 2 x 2 MiB pinned slots for absorb_table_bytes
    cudaEvent_t hash_ev[2] = {nullptr, nullptr};
    // double-buffered upload: H2D copies on their own stream overlap the layout kernels
    cudaStream_t copy_stream = nullptr;
    void* up_stage[2] = {nullptr, nullptr};
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr}, ev_entry = nullptr;
    // handles
    uint64_t next_handle = 1;
    std::unordered_map<uint64_t, Table> mles;
    std::unordered_map<uint64_t, std::unique_ptr<SumPolyState>> sps;
    std::unordered_map<uint64_t, std::unique_ptr<CircuitState>> circs;
    std::unordered_map<uint64_t, std::unique_ptr<KzgState>> kzgs;
    std::unordered_map<uint64_t, std::unique_ptr<MerkleState>> merkles;
    G1Affine* g1_table = nullptr;  // 32 x 256 multiples of the generator (fixed-base windows), made on first use
    cudaStream_t kzg_streams[4] = {nullptr, nullptr, nullptr, nullptr};  // side streams of get_proof (made on first use)
    cudaEvent_t kzg_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // multi-GPU
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, log2world = 0;
    uint32_t gather_log2 = 0;                // 0 = automatic (gather_threshold_n)
    ShmComm shm;
    bool use_shm = false;
    RoundInterpolator interp[MAXPTS + 1];
    FixedMulBuilder fmb;
    // tensor-core folds (tcfold.cuh): on for the throughput-bound rounds of products of >= 2 factors; ZKB200_NO_TC=1 keeps
    // every fold on the CUDA cores
    TcMatsBuilder tcm;
    Fe* d_cpow8 = nullptr;
    bool tc_enabled = true;
    int cluster_max = 1;        // CTAs of the largest thread-block cluster k_sc_small is launched with (ZKB200_CLUSTER_MAX caps it; 1 = single CTA)
    bool tc_tail_ok = false;    // two CTAs per SM of the tensor-core persistent kernel are co-resident (probed at creation)
    // persistent round kernel
    TailMailbox* mb = nullptr;   // mapped pinned host memory
    TailRelay* d_relay = nullptr;
    unsigned int tail_seq = 0;
    uint32_t tail_log2 = 40;                 // every unsharded round after the first runs in a persistent kernel
    uint32_t small_bytes = SMALL_SMEM_MAX;   // shared-memory budget of k_sc_small (0 = off)
    bool dt_enabled = false;                 // k_sc_small derives the challenges itself (device transcript), host replays;
                                             // off by default: measured slower than the host path
    DtRound* dt_rounds = nullptr;            // mapped host memory, DT_MAX_ROUNDS records
    uint64_t dt_launches = 0, dt_rounds_checked = 0;
    // ZKB200_TRACE=1: where a persistent-kernel round spends its time (printed at zkb_ctx_destroy)
    bool trace = false, trace_verbose = false, dbg_done = false;
    unsigned long long* d_dbg = nullptr;
    int dbg_grid = 0;
    double tr_n = 0, tr_host = 0, tr_rtt = 0, tr_relay = 0, tr_spread = 0, tr_pass = 0, tr_reduce = 0;
    // per table size (log2 of the entries bound in the round): rounds seen, device round trip, wait for the peer ranks
    double trs_n[48] = {0}, trs_rtt[48] = {0}, trs_wait[48] = {0}, trs_host[48] = {0};
    std::unordered_map<std::string, int> occ_cache;
    // per-launch event timing (zkb_ctx_profile)
    bool prof = false;
    struct ProfRec { int kind; cudaEvent_t e0, e1; double bytes; bool open; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> ev_pool;
    uint64_t prof_launches[ZKB_K_COUNT] = {0};
    double prof_ms[ZKB_K_COUNT] = {0}, prof_bytes[ZKB_K_COUNT] = {0};
    int cur_kind = ZKB_K_OTHER;
    double cur_bytes = 0;
};

namespace {

#define ZK_CUDA(ctx, call)                                                                    \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            (ctx)->last_error = std::string(#call) + ": " + cudaGetErrorString(e_);           \
            return e_ == cudaErrorMemoryAllocation ? ZKB_ERR_OOM : ZKB_ERR_CUDA;               \
        }                                                                                     \
    } while (0)
#define ZK_NCCL(ctx, call)                                                                    \
    do {                                                                                      \
        ncclResult_t e_ = (call);                                                             \
        if (e_ != ncclSuccess) {                                                              \
            (ctx)->last_error = std::string(#call) + ": " +                                   \
                                (g_nccl.GetErrorString ? g_nccl.GetErrorString(e_) : "nccl"); \
            return ZKB_ERR_NCCL;                                                              \
        }                                                                                     \
    } while (0)
#define ZK_TRY(expr)                 \
    do {                             \
        int32_t s_ = (expr);         \
        if (s_ != ZKB_OK) return s_; \
    } while (0)
#define ZK_FAIL(ctx, code, msg)    \
    do {                           \
        (ctx)->last_error = (msg); \
        return (code);             \
    } while (0)

cudaEvent_t prof_event(zkb_ctx* c) {
    if (!c->ev_pool.empty()) {
        cudaEvent_t e = c->ev_pool.back();
        c->ev_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
// Call right before a launch: names the kernel and its algorithmic bytes for the profile.
inline void prof_begin(zkb_ctx* c, int kind, double bytes) {
    if (!c->prof) return;
    zkb_ctx::ProfRec r{kind, prof_event(c), prof_event(c), bytes, true};
    cudaEventRecord(r.e0, c->stream);
    c->prof_recs.push_back(r);
}
void prof_drain(zkb_ctx* c) {
    for (auto& r : c->prof_recs) {
        float ms = 0;
        if (!r.open && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
            c->prof_launches[r.kind] += 1;
            c->prof_ms[r.kind] += ms;
            c->prof_bytes[r.kind] += r.bytes;
        }
        c->ev_pool.push_back(r.e0);
        c->ev_pool.push_back(r.e1);
    }
    c->prof_recs.clear();
}

int32_t check_launch(zkb_ctx* c, const char* what) {
    if (c->prof && !c->prof_recs.empty() && c->prof_recs.back().open) {
        cudaEventRecord(c->prof_recs.back().e1, c->stream);
        c->prof_recs.back().open = false;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        c->last_error = std::string(what) + ": " + cudaGetErrorString(e);
        return ZKB_ERR_CUDA;
    }
    ++c->launches;
    return ZKB_OK;
}

int grid_for(const zkb_ctx* c, uint64_t items, int ctas_per_sm) {
    uint64_t need = (items + BLOCK - 1) / BLOCK;
    uint64_t cap = (uint64_t)c->sm_count * (uint64_t)(ctas_per_sm > 0 ? ctas_per_sm : 1);
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

int32_t alloc_table(zkb_ctx* c, uint64_t n, Table* t) {
    void* p = nullptr;
    ZK_CUDA(c, cudaMallocAsync(&p, (size_t)n * 32, c->stream));
    t->base = (uint4*)p;
    t->n = n;
    t->stride = n;
    return ZKB_OK;
}
void free_table(zkb_ctx* c, Table* t) {
    if (t->base) cudaFreeAsync(t->base, c->stream);
    t->base = nullptr;
    t->n = t->stride = 0;
}
int32_t ensure_stage(zkb_ctx* c, size_t bytes) {
    if (c->stage_bytes >= bytes) return ZKB_OK;
    if (c->stage) cudaFreeAsync(c->stage, c->stream);
    c->stage = nullptr;
    c->stage_bytes = 0;
    ZK_CUDA(c, cudaMallocAsync(&c->stage, bytes, c->stream));
    c->stage_bytes = bytes;
    return ZKB_OK;
}
int32_t ensure_partials(zkb_ctx* c, size_t n_fe) {
    if (c->partials_cap >= n_fe) return ZKB_OK;
    if (c->d_partials) cudaFreeAsync(c->d_partials, c->stream);
    c->d_partials = nullptr;
    ZK_CUDA(c, cudaMallocAsync((void**)&c->d_partials, n_fe * sizeof(Fe), c->stream));
    c->partials_cap = n_fe;
    return ZKB_OK;
}

Table* find_mle(zkb_ctx* c, zkb_mle h) {
    auto it = c->mles.find(h);
    return it == c->mles.end() || it->second.dead ? nullptr : &it->second;
}
zkb_mle put_mle(zkb_ctx* c, const Table& t) {
    zkb_mle h = c->next_handle++;
    c->mles[h] = t;
    return h;
}

// Wait for the mailbox flag to reach `seq` (spin on mapped host memory; a stream
// query every few thousand spins turns a faulted kernel into an error instead of
// a hang).
int32_t wait_mailbox(zkb_ctx* c, unsigned int seq) {
    uint32_t spins = 0;
    while (load_acquire(c->h_flag) != seq) {
        if ((++spins & 0x3fff) == 0) {
            cudaError_t e = cudaStreamQuery(c->stream);
            if (e == cudaSuccess) {
                if (load_acquire(c->h_flag) == seq) break;
                ZK_FAIL(c, ZKB_ERR_CUDA, "mailbox: stream drained without a result");
            }
            if (e != cudaErrorNotReady) {
                c->last_error = std::string("mailbox: ") + cudaGetErrorString(e);
                return ZKB_ERR_CUDA;
            }
        }
        cpu_relax();
    }
    return ZKB_OK;  // the acquire load above orders the reads of the results after the flag
}

// Fill the FinishArgs of a reducing launch of `grid` CTAs returning `npts` sums.
int32_t prep_finish(zkb_ctx* c, int grid, int npts, bool sharded, FinishArgs* f) {
    ZK_TRY(ensure_partials(c, (size_t)grid * npts));
    f->partials = c->d_partials;
    f->ticket = c->d_ticket;
    f->stamp = nullptr;
    f->chk = nullptr;
    if (sharded && !c->use_shm) {
        f->result = c->d_res;
        f->result_wide = c->d_wide;
        f->flag = nullptr;
        f->seq = 0;
    } else {
        f->result = c->h_res;
        f->result_wide = nullptr;
        f->flag = c->h_flag;
        f->seq = ++c->seq;
    }
    return ZKB_OK;
}
// Bring the `npts` sums of the launch prepared by prep_finish to the host.
int32_t collect(zkb_ctx* c, int npts, bool sharded, const FinishArgs& f, Fe* out) {
    if (!sharded || c->use_shm) {
        ZK_TRY(wait_mailbox(c, f.seq));
        for (int p = 0; p < npts; ++p) out[p] = c->h_res[p];
        if (sharded && !c->shm.allreduce(c->H, out, npts)) ZK_FAIL(c, ZKB_ERR_NCCL, "shared-memory exchange: a peer rank stopped answering");
        return ZKB_OK;
    }
    // C1: exact integer sum of the ranks' residues on zero-extended limbs, reduced mod p on the host
    ZK_NCCL(c, g_nccl.AllReduce(c->d_wide, c->d_wide, (size_t)npts * 8, ncclUint64, ncclSum, c->comm, c->stream));
    ZK_CUDA(c, cudaMemcpyAsync(c->h_wide, c->d_wide, (size_t)npts * 8 * sizeof(unsigned long long),
                               cudaMemcpyDeviceToHost, c->stream));
    ZK_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int p = 0; p < npts; ++p) out[p] = c->H.from_wide_limbs(c->h_wide + 8 * p);
    return ZKB_OK;
}

int sc_occ(zkb_ctx* c, int fused, int kind, int D, int npts);
// Tensor-core folds for a round pass whose smallest output tables have n_out_min entries: throughput-bound sizes only
// (a tile is 128 quad positions; below 2^16 entries the TMA -> MMA -> TMEM pipeline is latency, not bandwidth).
// smallest round (entries written per table) the tensor-core kernels take; ZKB200_TC_MIN_LOG2 overrides (tuning)
static const uint64_t TC_MIN_N_OUT = 1ull << (getenv("ZKB200_TC_MIN_LOG2") ? std::atoi(getenv("ZKB200_TC_MIN_LOG2")) : 15);
// ... and the smallest the PERSISTENT tensor-core launch continues to once it runs (ZKB200_TC_TAIL_MIN_LOG2): a launch
// boundary (~35 us) costs more than its small rounds lose against the CUDA-core launch (15 vs 16 us per round), so it
// runs down to wherever the on-chip kernel / the gather of the shards takes over (a tile is 128 quad positions: >= 2^9).
static const uint64_t TC_TAIL_MIN_N_OUT = 1ull << (getenv("ZKB200_TC_TAIL_MIN_LOG2") ? std::atoi(getenv("ZKB200_TC_TAIL_MIN_LOG2")) : 9);
bool tc_round_ok(zkb_ctx* c, int kind, int D, int npts, uint64_t n_out_min, int fused, uint64_t floor_n = TC_MIN_N_OUT) {
    return c->tc_enabled && tc_shape(kind, D, npts) && n_out_min >= floor_n && n_out_min >= 256 && ((n_out_min >> 1) & 127u) == 0 && sc_occ(c, fused, kind, D, npts) > 0;
}
int sc_occ(zkb_ctx* c, int fused, int kind, int D, int npts) {
    char key[64];
    snprintf(key, sizeof key, "%d/%d/%d/%d", fused, kind, D, npts);
    auto it = c->occ_cache.find(key);
    if (it != c->occ_cache.end()) return it->second;
    int o = c->K->sc_occupancy(fused, kind, D, npts);
    c->occ_cache[key] = o;
    return o;
}

// --------------------------------------------------------------- SumPoly core
int32_t sp_configure(zkb_ctx* c, SumPolyState* sp, int P, int D, int kind) {
    sp->P = P;
    sp->D = D;
    sp->kind = kind;
    sp->sel.clear();
    sp->rest.clear();
    const int T = P * D;
    if (kind == KIND_XYZ) {
        sp->kP = 1;
        sp->kD = 2;
        sp->npts = 3;
        for (int t = 0; t < 3; ++t) sp->sel.push_back(t);
        return ZKB_OK;
    }
    if (T > MAXT || D + 1 > MAXPTS || D < 1 || P < 1) ZK_FAIL(c, ZKB_ERR_UNSUPPORTED, "sumpoly: more than 16 tables or degree > 4");
    sp->npts = D + 1;
    if (c->mode == ZKB_MODE_COMPAT && !(P == 1 && D == 1)) {
        // SumPoly::reduce adds products 0 and 1 of ProductPoly::reduce = factor0 * factor1 (composed_polynomial.rs:52-54,88-99)
        if (P < 2 || D < 2) ZK_FAIL(c, ZKB_ERR_COMPAT_SHAPE, "compat mode needs >= 2 products of >= 2 factors");
        sp->kP = 2;
        sp->kD = 2;
        for (int p = 0; p < 2; ++p)
            for (int f = 0; f < 2; ++f) sp->sel.push_back(p * D + f);
        for (int t = 0; t < T; ++t) {
            bool used = (t / D < 2) && (t % D < 2);
            if (!used) sp->rest.push_back(t);
        }
    } else {
        sp->kP = P;
        sp->kD = D;
        for (int t = 0; t < T; ++t) sp->sel.push_back(t);
    }
    if (sc_occ(c, 0, sp->kind, sp->kD, sp->npts) <= 0) ZK_FAIL(c, ZKB_ERR_UNSUPPORTED, "sumpoly: no kernel for this (degree, points) shape");
    return ZKB_OK;
}

void sp_release(zkb_ctx* c, SumPolyState* sp) {
    for (auto& t : sp->work) free_table(c, &t);
    for (auto& t : sp->gath) free_table(c, &t);
    sp->work.clear();
    sp->gath.clear();
}
void sp_reset(zkb_ctx* c, SumPolyState* sp) {
    sp->state = 0;
    sp->have_evals = false;
    sp->cur_n = sp->n0;
    sp->sharded = c->comm != nullptr && c->world > 1;
    for (auto& t : sp->gath) free_table(c, &t);
    sp->gath.clear();
}

int32_t sp_round_evals(zkb_ctx* c, SumPolyState* sp, Fe* evals) {
    if (sp->cur_n < 2) ZK_FAIL(c, ZKB_ERR_ARITY, "round_evals: no variable left");
    ScArgs a;
    std::memset(&a, 0, sizeof a);
    for (size_t i = 0; i < sp->sel.size(); ++i) a.in[i] = sp->cur(sp->sel[i]).ref();
    a.n_tables = (int)sp->sel.size();
    a.n_products = sp->kP;
    a.n_out = sp->cur_n;
    // products of two factors, throughput-bound sizes: the sums of products are a Gram matrix on the tensor cores
    const int tc_occ = c->tc_enabled && sp->cur_n >= (1ull << 17) && ((sp->cur_n >> 1) & 127u) == 0 ? sc_occ(c, 5, sp->kind, sp->kD, sp->npts) : 0;
    const bool tc = tc_occ > 0;
    const int grid = grid_for(c, sp->cur_n / 2, tc ? tc_occ : sc_occ(c, 0, sp->kind, sp->kD, sp->npts));
    ZK_TRY(prep_finish(c, grid, sp->npts, sp->sharded, &a.fin));
    prof_begin(c, ZKB_K_SC_EVAL, 32.0 * (double)sp->sel.size() * (double)sp->cur_n);
    if (tc) {
        if (!c->K->sc_eval_tc(sp->kind, sp->kD, sp->npts, a, grid, c->stream)) ZK_FAIL(c, ZKB_ERR_UNSUPPORTED, "sc_eval_tc: shape not instantiated");
    } else if (!c->K->sc_eval(sp->kind, sp->kD, sp->npts, a, grid, c->stream)) ZK_FAIL(c, ZKB_ERR_UNSUPPORTED, "sc_eval: shape not instantiated");
    ZK_TRY(check_launch(c, "k_sc_eval"));
    ZK_TRY(collect(c, sp->npts, sp->sharded, a.fin, evals));
    for (int i = 0; i < sp->npts; ++i) sp->last_evals[i] = evals[i];
    sp->have_evals = true;
    return ZKB_OK;
}

int32_t sp_ensure_work(zkb_ctx* c, SumPolyState* sp) {
    if (!sp->work.empty()) return ZKB_OK;
    const int T = (int)sp->src.size();
    sp->work.resize(T);
    for (int t = 0; t < T; ++t) ZK_TRY(alloc_table(c, sp->n0 / 2 ? sp->n0 / 2 : 1, &sp->work[t]));
    return ZKB_OK;
}

// C2: gather every rank's shard of every table and interleave to global order.
// Local table size at which the shards are gathered (C2).  Automatic: as soon as the REPLICATED table fits the on-chip
// kernel, so the rounds after the gather are one k_sc_small launch; with the per-round NCCL fallback (no shared
// memory between the ranks' hosts) every sharded round is a launch and an all-reduce, so gather early instead.
// Largest table (entries) one CTA of k_sc_small holds for `tables` tables / the largest its cluster holds.
uint64_t small_cap_cta(const zkb_ctx* c, size_t tables) {
    if (!c->small_bytes || !tables) return 0;
    const uint64_t budget = c->small_bytes < (uint32_t)SMALL_SMEM_MAX ? c->small_bytes : (uint32_t)SMALL_SMEM_MAX;
    uint64_t n = 1;
    while (2 * n * tables * 32 <= budget) n *= 2;
    return n >= 2 ? n : 0;
}
// (the device-side transcript, an opt-in experiment, exists for the single-CTA kernel only)
inline int small_cluster_ctas(const zkb_ctx* c) { return c->dt_enabled ? 1 : c->cluster_max; }
uint64_t small_cap_cluster(const zkb_ctx* c, size_t tables) {
    const uint64_t n = small_cap_cta(c, tables), k = (uint64_t)small_cluster_ctas(c);
    // (a CTA of a cluster keeps >= 2 x cluster entries per table: the collection step of k_sc_small stages behind them)
    return n >= 2 * k ? n * k : n;
}
uint64_t gather_threshold_n(const zkb_ctx* c, const SumPolyState* sp) {
    if (c->gather_log2) return 1ull << c->gather_log2;
    if (c->use_shm && c->small_bytes && sp->rest.empty() && !sp->sel.empty()) {
        // (the single-CTA size, not the cluster's: gathering 16 x larger shards earlier was not faster at 2 GPUs, 1.12 vs
        // 1.13 ms, and the all-gather of 8 ranks grows with it)
        const uint64_t n = small_cap_cta(c, sp->sel.size());
        const uint64_t g = n >> c->log2world;
        if (g >= 2) return g;
    }
    return 1ull << 14;
}
int32_t sp_gather(zkb_ctx* c, SumPolyState* sp) {
    // ONE all-gather for all tables (the shards are a few KiB by now: the call's latency is what counts): every rank sends
    // [table][plane][nl] and receives the same block of every rank
    const int T = (int)sp->src.size();
    const uint64_t nl = sp->cur_n, G = (uint64_t)c->world;
    const uint64_t blk = (uint64_t)T * 2 * nl;  // uint4 per rank
    ZK_TRY(ensure_stage(c, (size_t)(G + 1) * blk * 16));
    uint4* send = (uint4*)c->stage;
    uint4* recv = send + blk;
    for (int t = 0; t < T; ++t) {
        Table& cur = sp->cur(t);
        ZK_CUDA(c, cudaMemcpyAsync(send + (uint64_t)t * 2 * nl, cur.base, nl * 16, cudaMemcpyDeviceToDevice, c->stream));
        ZK_CUDA(c, cudaMemcpyAsync(send + (uint64_t)t * 2 * nl + nl, cur.base + cur.stride, nl * 16, cudaMemcpyDeviceToDevice, c->stream));
    }
    ZK_NCCL(c, g_nccl.AllGather(send, recv, (size_t)blk * 16, ncclUint8, c->comm, c->stream));
    sp->gath.resize(T);
    for (int t = 0; t < T; ++t) {
        Table g;
        ZK_TRY(alloc_table(c, nl * G, &g));
        TabRef gr{recv + (uint64_t)t * 2 * nl, nl};
        c->K->interleave_shards(gr, blk, g.ref(), nl, (uint32_t)c->log2world, grid_for(c, nl * G, 8), c->stream);
        ZK_TRY(check_launch(c, "k_interleave_shards"));
        sp->gath[t] = g;
    }
    sp->state = 2;
    sp->cur_n = nl * G;
    sp->sharded = false;
    return ZKB_OK;
}

// Fold every table with r; if `evals` != NULL also return the next round's evaluations (one pass).
int32_t sp_bind_and_next(zkb_ctx* c, SumPolyState* sp, const Fe& r, Fe* evals, Fe* final_vals) {
    if (sp->cur_n < 2) ZK_FAIL(c, ZKB_ERR_ARITY, "bind: no variable left");
    if (sp->sharded && sp->cur_n <= gather_threshold_n(c, sp)) ZK_TRY(sp_gather(c, sp));
    const int T = (int)sp->src.size();
    if (sp->state == 0) ZK_TRY(sp_ensure_work(c, sp));
    const uint64_t n_out = sp->cur_n / 2;
    auto dst = [&](int t) -> Table& { return sp->state == 2 ? sp->gath[t] : sp->work[t]; };

    if (n_out == 1 && !sp->sharded) {  // last bind: publish the bound values
        FoldTablesArgs fa;
        std::memset(&fa, 0, sizeof fa);
        for (int t = 0; t < T; ++t) {
            fa.in[t] = sp->cur(t).ref();
            fa.out[t] = dst(t).ref();
        }
        fa.n_tables = T;
        fa.n_out = 1;
        c->fmb.make(c->H, r, &fa.rt);
        sp->have_evals = false;
        unsigned int seq = ++c->seq;
        prof_begin(c, ZKB_K_FINAL_BIND, 96.0 * T);
        c->K->final_bind(fa, c->h_res, c->h_flag, seq, c->stream);
        ZK_TRY(check_launch(c, "k_final_bind"));
        if (sp->state == 0) sp->state = 1;
        sp->cur_n = 1;
        ZK_TRY(wait_mailbox(c, seq));
        if (final_vals)
            for (int t = 0; t < T; ++t) final_vals[t] = c->h_res[t];
        if (evals) ZK_FAIL(c, ZKB_ERR_ARITY, "bind_and_next: no next round after the last variable");
        return ZKB_OK;
    }

    if (!sp->rest.empty() || !evals || n_out < 2) {
        // tables outside the round kernel (compat mode) or a plain fold request
        FoldTablesArgs fa;
        std::memset(&fa, 0, sizeof fa);
        int k = 0;
        const bool all = (!evals || n_out < 2);
        for (int t = 0; t < T; ++t) {
            bool in_rest = false;
            for (int x : sp->rest) in_rest |= (x == t);
            if (!all && !in_rest) continue;
            fa.in[k] = sp->cur(t).ref();
            fa.out[k] = dst(t).ref();
            ++k;
        }
        fa.n_tables = k;
        fa.n_out = n_out;
        c->fmb.make(c->H, r, &fa.rt);
        prof_begin(c, ZKB_K_FOLD_TABLES, 96.0 * (double)k * (double)n_out);
        c->K->fold_tables(fa, grid_for(c, n_out * (uint64_t)k, 8), c->stream);
        ZK_TRY(check_launch(c, "k_fold_tables"));
        if (all) {
            if (sp->state == 0) sp->state = 1;
            sp->cur_n = n_out;
            sp->have_evals = false;
            if (evals) return sp_round_evals(c, sp, evals);  // only reached for sharded 1-entry shards
            return ZKB_OK;
        }
    }
    if (!sp->have_evals) {  // bind requested before this round's evaluations were taken: take them now
        Fe tmp[MAXPTS];
        ZK_TRY(sp_round_evals(c, sp, tmp));
    }
    ScArgs a;
    std::memset(&a, 0, sizeof a);
    for (size_t i = 0; i < sp->sel.size(); ++i) {
        a.in[i] = sp->cur(sp->sel[i]).ref();
        a.out[i] = dst(sp->sel[i]).ref();
    }
    a.n_tables = (int)sp->sel.size();
    a.n_products = sp->kP;
    a.n_out = n_out;
    c->fmb.make(c->H, r, &a.rt);
    // the claim chain: this round's s(0) + s(1) equals the previous round polynomial at r
    const RoundInterpolator& ip = c->interp[sp->npts];
    Fe co[MAXPTS];
    const int colen = ip.interpolate(sp->last_evals, co);
    const Fe claim = uni_evaluate(c->H, co, colen, r);
    const bool tc = tc_round_ok(c, sp->kind, sp->kD, sp->npts, n_out, 3);
    const int grid = grid_for(c, n_out / 2, sc_occ(c, tc ? 3 : 1, sp->kind, sp->kD, sp->npts));
    ZK_TRY(prep_finish(c, grid, sp->npts - 1, sp->sharded, &a.fin));
    if (tc) {
        ScArgsTc at;
        at.s = a;
        c->tcm.make(c->H, r, &at.mats);
        prof_begin(c, ZKB_K_SC_FOLD_EVAL, 96.0 * (double)sp->sel.size() * (double)n_out);  // after the host arithmetic: the interval is the launch
        if (!c->K->sc_fold_eval_tc(sp->kind, sp->kD, sp->npts, at, grid, c->stream)) ZK_FAIL(c, ZKB_ERR_UNSUPPORTED, "sc_fold_eval_tc: shape not instantiated");
    } else {
        prof_begin(c, ZKB_K_SC_FOLD_EVAL, 96.0 * (double)sp->sel.size() * (double)n_out);
        if (!c->K->sc_fold_eval(sp->kind, sp->kD, sp->npts, a, grid, c->stream)) ZK_FAIL(c, ZKB_ERR_UNSUPPORTED, "sc_fold_eval: shape not instantiated");
    }
    ZK_TRY(check_launch(c, "k_sc_fold_eval"));
    if (sp->state == 0) sp->state = 1;
    sp->cur_n = n_out;
    Fe got[MAXPTS];
    ZK_TRY(collect(c, sp->npts - 1, sp->sharded, a.fin, got));
    evals[0] = got[0];
    evals[1] = c->H.sub(claim, got[0]);
    for (int t = 2; t < sp->npts; ++t) evals[t] = got[t - 1];
    for (int i = 0; i < sp->npts; ++i) sp->last_evals[i] = evals[i];
    return ZKB_OK;
}

int32_t sp_final_values(zkb_ctx* c, SumPolyState* sp, Fe* vals) {
    if (sp->cur_n != 1) ZK_FAIL(c, ZKB_ERR_ARITY, "final_values: variables left to bind");
    const int T = (int)sp->src.size();
    GatherArgs g;
    std::memset(&g, 0, sizeof g);
    for (int t = 0; t < T; ++t) g.t[t] = sp->cur(t).ref();
    g.n = T;
    g.idx = 0;
    g.out = c->d_res;
    launch_gather_elems(g, c->stream);
    ZK_TRY(check_launch(c, "k_gather_elems"));
    ZK_CUDA(c, cudaMemcpyAsync(c->h_res, c->d_res, sizeof(Fe) * T, cudaMemcpyDeviceToHost, c->stream));
    ZK_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int t = 0; t < T; ++t) vals[t] = c->h_res[t];
    return ZKB_OK;
}


// ------------------------------------------------------------- round driver
// One sumcheck from the first round to the last bind.  Three regimes, chosen by table size:
//   * sharded tables (multi-GPU): one launch per round + NCCL all-reduce (sp_bind_and_next);
//   * large tables: the persistent cooperative kernel k_sc_tail (all rounds in one launch,
//     challenges through the host mailbox);
//   * tables that fit in shared memory: the single-CTA kernel k_sc_small, which can also
//     produce round 0 itself, so a small sumcheck is exactly one launch.
// table size below which the rounds are latency-bound (CUDA-core launch of the persistent kernel, or the on-chip kernel
// k_sc_small where its cluster holds the tables: 2^15 entries for two tables); ZKB200_MID_LOG2 overrides (tuning).
// Measured at 2^24 x 2 tables: tensor-core launch down to 2^17 + CUDA-core launch to 2^15 + k_sc_small 1.063 ms; tensor-core
// launch down to 2^15 + k_sc_small 1.037 ms (a launch boundary costs more than two 18 us rounds).
static const uint64_t MID_N = 1ull << (getenv("ZKB200_MID_LOG2") ? std::atoi(getenv("ZKB200_MID_LOG2")) : 15);
struct RoundDriver {
    zkb_ctx* c;
    SumPolyState* sp;
    bool live = false;       // a persistent kernel is running
    bool small = false;      // ... and it is k_sc_small
    unsigned int base = 0;   // its mailbox sequence base
    unsigned int sent = 0;   // challenges delivered through the mailbox so far
    unsigned int pubs = 0;   // messages consumed from it so far
    uint64_t stop_n = 0;     // k_sc_tail leaves once the tables have <= stop_n entries
    unsigned long long t_recv = 0;
    static double tsc_per_us() {
        static double v = 0;
        if (v == 0) {
            timespec a, b;
            clock_gettime(CLOCK_MONOTONIC, &a);
            unsigned long long t0 = host_ticks();
            do clock_gettime(CLOCK_MONOTONIC, &b);
            while ((b.tv_sec - a.tv_sec) * 1e9 + (b.tv_nsec - a.tv_nsec) < 2e6);
            v = (double)(host_ticks() - t0) / (((b.tv_sec - a.tv_sec) * 1e9 + (b.tv_nsec - a.tv_nsec)) * 1e-3);
        }
        return v;
    }
    Fe poly[MAXPTS];         // coefficients of the current round polynomial, if the caller interpolated it already
    int poly_len = -1;
    TranscriptImpl* tr = nullptr;  // the caller's transcript: lets k_sc_small continue it on the device (device transcript)
    bool dt = false;               // the live k_sc_small derives its own challenges; the host only checks them

    RoundDriver(zkb_ctx* ctx, SumPolyState* s) : c(ctx), sp(s) {}
    ~RoundDriver() { abort(); }

    void abort() {
        if (!live) return;
        c->mb->abort = 1;
        cudaStreamSynchronize(c->stream);
        c->mb->abort = 0;
        live = false;
    }
    uint64_t small_cap() const {  // largest table (entries) k_sc_small takes for this shape
        if (!c->small_bytes || sp->sharded || !sp->rest.empty()) return 0;
        return small_cap_cluster(c, sp->sel.size());
    }
    bool small_ok() const { return sp->cur_n >= 2 && sp->cur_n <= small_cap(); }
    uint64_t gather_n() const { return gather_threshold_n(c, sp); }
    bool tail_ok() const {
        // products of >= 3 factors: the persistent kernel spills with the challenge table in shared memory,
        // so their large (throughput-bound) rounds stay one launch each
        if (sp->kind == KIND_PROD && sp->kD >= 3 && sp->cur_n > (1ull << 16)) return false;
        // sharded tables can use it when the ranks' hosts exchange the sums through shared memory
        if (sp->sharded && (!c->use_shm || sp->cur_n <= gather_n())) return false;
        return c->tail_log2 > 0 && sp->rest.empty() && sp->have_evals && sp->cur_n >= 2 &&
               sp->cur_n <= (1ull << (c->tail_log2 > 62 ? 62 : c->tail_log2)) && sc_occ(c, 2, sp->kind, sp->kD, sp->npts) > 0;
    }
    int32_t wait_dev(unsigned int want) {
        uint32_t spins = 0;
        // "not yet reached": with the device transcript the kernel does not wait for the host and may already be
        // several messages ahead, so the sequence number is compared as a counter, not for equality
        auto behind = [&]() { return (int32_t)(load_acquire(&c->mb->dev_seq) - want) < 0; };
        while (behind()) {
            if ((++spins & 0x3fff) == 0) {
                cudaError_t e = cudaStreamQuery(c->stream);
                if (e == cudaSuccess && behind()) {
                    live = false;
                    ZK_FAIL(c, ZKB_ERR_CUDA, c->mb->dev_error ? "persistent round kernel timed out waiting for the host" : "persistent round kernel exited early");
                }
                if (e != cudaSuccess && e != cudaErrorNotReady) {
                    live = false;
                    c->last_error = std::string("persistent round kernel: ") + cudaGetErrorString(e);
                    return ZKB_ERR_CUDA;
                }
            }
            cpu_relax();
        }
        return ZKB_OK;  // (acquire load in behind(): mb->finals / evals are read after dev_seq)
    }
    // The persistent kernels publish a round's sums without a system fence (kernels.cuh FinishArgs::chk): the sequence
    // number may become visible before the sums, so they are taken only once their checksum matches.
    int32_t read_msg(Fe* out, int n, unsigned int seq) {
        for (uint32_t spins = 0;; ++spins) {
            for (int i = 0; i < n; ++i) out[i] = c->mb->evals[i];
            const unsigned int g0 = c->mb->dev_chk[0], g1 = c->mb->dev_chk[1];
            unsigned int c0, c1;
            msg_checksum(out, n, seq, &c0, &c1);
            if (c0 == g0 && c1 == g1) return ZKB_OK;
            if (spins > 50000000u) {
                abort();
                ZK_FAIL(c, ZKB_ERR_CUDA, "round message from the device never became consistent");
            }
            cpu_relax();
        }
    }
    void begin_mailbox() {
        c->tail_seq += 64;
        base = c->tail_seq;
        sent = pubs = 0;
        c->mb->abort = 0;
        c->mb->dev_error = 0;
    }
    void send(const Fe& r) {
        const unsigned int want = base + (++sent);
        uint32_t x = 0;
        for (int k = 0; k < 8; ++k) {
            c->mb->r[k] = r.l[k];
            x ^= r.l[k];
        }
        c->mb->chk = x ^ (want * 0x9E3779B9u);
        store_release(&c->mb->host_seq, want);  // the payload is visible before the sequence number (and the device checks chk)
    }
    int32_t launch_small(bool first_eval, const Fe& r, const Fe* claim = nullptr) {
        if (sp->state == 0) ZK_TRY(sp_ensure_work(c, sp));
        SmallArgs a;
        std::memset(&a, 0, sizeof a);
        // Device transcript: the sponge as it stands now (after the challenge r was drawn, or before round 0) moves
        // to the kernel; the host keeps its own copy and replays every round behind the device.
        dt = false;
        const uint64_t cta_cap = small_cap_cta(c, sp->sel.size());
        int nc = 1;
        while (cta_cap && sp->cur_n / (uint64_t)nc > cta_cap && nc < small_cluster_ctas(c)) nc *= 2;
        if (tr && c->dt_enabled && c->dt_rounds && nc == 1 && sp->npts == 3 && (first_eval || claim) && ilog2_u64(sp->cur_n) + 1 <= DT_MAX_ROUNDS &&
            tr->hasher.snapshot(a.dt.st, &a.dt.fill_words)) {
            dt = true;
            a.dt.enabled = 1;
            if (claim) a.dt.claim0 = *claim;
            a.dt.rounds = c->dt_rounds;
            ++c->dt_launches;
        }
        for (size_t i = 0; i < sp->sel.size(); ++i) {
            a.in[i] = sp->cur(sp->sel[i]).ref();
            a.out[i] = (sp->state == 2 ? sp->gath[sp->sel[i]] : sp->work[sp->sel[i]]).ref();
        }
        a.n_tables = (int)sp->sel.size();
        a.n_products = sp->kP;
        a.n_in = (uint32_t)sp->cur_n;
        a.nc = nc;
        a.first_eval = first_eval ? 1 : 0;
        a.r0 = r;
        a.mb = c->mb;
        begin_mailbox();
        a.base_seq = base;
        a.timeout_clocks = 6000000000ll;
        prof_begin(c, ZKB_K_SC_SMALL, 32.0 * (double)sp->sel.size() * (double)sp->cur_n);
        int e = c->K->sc_small(sp->kind, sp->kD, sp->npts, a, c->stream);
        if (e < 0) ZK_FAIL(c, ZKB_ERR_UNSUPPORTED, "sc_small: shape not instantiated");
        if (e != 0) {
            c->last_error = std::string("k_sc_small launch: ") + cudaGetErrorString((cudaError_t)e);
            return ZKB_ERR_CUDA;
        }
        ZK_TRY(check_launch(c, "k_sc_small"));
        live = small = true;
        if (sp->state == 0) sp->state = 1;
        return ZKB_OK;
    }
    int32_t launch_tail(const Fe& r, bool first_eval = false) {
        if (sp->state == 0) ZK_TRY(sp_ensure_work(c, sp));
        TailArgs a;
        std::memset(&a, 0, sizeof a);
        for (size_t i = 0; i < sp->sel.size(); ++i) {
            a.in[i] = sp->cur(sp->sel[i]).ref();
            a.out[i] = (sp->state == 2 ? sp->gath[sp->sel[i]] : sp->work[sp->sel[i]]).ref();
        }
        a.n_tables = (int)sp->sel.size();
        a.n_products = sp->kP;
        a.n_in = sp->cur_n;
        if (!first_eval) c->fmb.make(c->H, r, &a.rt0);
        a.first_eval = first_eval ? 1 : 0;
        for (int i = 0; i < 8; ++i) a.cpow[i] = c->fmb.c[i];
        a.mb = c->mb;
        a.relay = c->d_relay;
        a.ticket = c->d_ticket;
        begin_mailbox();
        a.base_seq = base;
        a.timeout_clocks = 6000000000ll;  // ~3 s
        // Rounds above MID_N entries and the latency-bound rounds below run as two launches of the same
        // kernel, so that each launch (and its profile entry) belongs to one regime.
        stop_n = sp->sharded ? gather_n() : small_cap();
        static const uint64_t shard_mid = getenv("ZKB200_TC_SHARD_MIN_LOG2") ? MID_N : (1ull << 17);
        const uint64_t mid_n = sp->sharded && MID_N < shard_mid ? shard_mid : MID_N;
        const bool big = sp->cur_n > mid_n;
        // tensor-core folds when every round of this launch is large enough for them (its last round writes stop_n entries)
        bool tc = false;
        if (big && !first_eval && c->tc_tail_ok) {
            // sharded rounds (lock-step with the other ranks): the split of round 2's measured 8-GPU runs stays -- tensor cores
            // down to 2^17 entries, the CUDA-core launch below
            static const uint64_t shard_floor = 1ull << (getenv("ZKB200_TC_SHARD_MIN_LOG2") ? std::atoi(getenv("ZKB200_TC_SHARD_MIN_LOG2")) : 17);
            const uint64_t floor_n = sp->sharded ? shard_floor : TC_TAIL_MIN_N_OUT;
            const uint64_t tstop = stop_n < floor_n ? floor_n : stop_n;
            if (tc_round_ok(c, sp->kind, sp->kD, sp->npts, tstop, 4, floor_n)) {
                tc = true;
                stop_n = tstop;
            }
        }
        if (big && !tc && stop_n < mid_n) stop_n = mid_n;
        a.stop_n = stop_n;
        if (tc) {
            a.cpow8 = c->d_cpow8;
            c->tcm.make(c->H, r, &a.mats0);
        }
        const uint64_t quads = first_eval ? sp->cur_n / 2 : (sp->cur_n / 4 ? sp->cur_n / 4 : 1);
        const int grid = grid_for(c, quads, sc_occ(c, tc ? 4 : 2, sp->kind, sp->kD, sp->npts));
        ZK_TRY(ensure_partials(c, (size_t)grid * MAXPTS));
        a.partials = c->d_partials;
        ZK_CUDA(c, cudaMemsetAsync(&c->d_relay->seq, 0, 2 * sizeof(unsigned int), c->stream));
        if (c->trace_verbose && big && !c->dbg_done) {
            ZK_CUDA(c, cudaMalloc((void**)&c->d_dbg, sizeof(unsigned long long) * 2 * 1024));
            ZK_CUDA(c, cudaMemsetAsync(c->d_dbg, 0, sizeof(unsigned long long) * 2 * 1024, c->stream));
            a.dbg = c->d_dbg;
            c->dbg_grid = grid;
            c->dbg_done = true;
        }
        prof_begin(c, big ? ZKB_K_SC_TAIL : ZKB_K_SC_TAIL_MID, 96.0 * (double)sp->sel.size() * (double)(sp->cur_n - (stop_n ? stop_n : 1)) +
                                         (first_eval ? 32.0 * (double)sp->sel.size() * (double)sp->cur_n : 0.0));
        int e = c->K->sc_tail(sp->kind, sp->kD, sp->npts, tc, a, grid, c->stream);
        if (e < 0) ZK_FAIL(c, ZKB_ERR_UNSUPPORTED, "sc_tail: shape not instantiated");
        if (e != 0) {
            c->last_error = std::string("k_sc_tail launch: ") + cudaGetErrorString((cudaError_t)e);
            return ZKB_ERR_CUDA;
        }
        ZK_TRY(check_launch(c, "k_sc_tail"));
        live = true;
        small = false;
        if (sp->state == 0) sp->state = 1;
        return ZKB_OK;
    }
    // Round 0: s(0..d) of the unbound tables.
    int32_t first(Fe* evals) {
        if (sp->cur_n >= 2 && small_ok() && sc_occ(c, 0, sp->kind, sp->kD, sp->npts) > 0) {
            ZK_TRY(launch_small(true, c->H.zero()));
            ZK_TRY(wait_dev(base + (++pubs)));
            if (dt) for (int i = 0; i < sp->npts; ++i) evals[i] = c->dt_rounds[pubs - 1].evals[i];
            else ZK_TRY(read_msg(evals, sp->npts, base + pubs));
            for (int i = 0; i < sp->npts; ++i) sp->last_evals[i] = evals[i];
            sp->have_evals = true;
            return ZKB_OK;
        }
        // round 0 inside the persistent kernel: one launch for the whole sumcheck
        sp->have_evals = true;  // (tail_ok asks for it; set for real below)
        // (only for latency-bound sizes: for large tables the stand-alone k_sc_eval measured 3% faster)
        const bool tail_first = sp->cur_n >= 4 && sp->cur_n <= (1ull << 18) && tail_ok();
        sp->have_evals = false;
        if (tail_first) {
            ZK_TRY(launch_tail(c->H.zero(), true));
            ZK_TRY(wait_dev(base + (++pubs)));
            ZK_TRY(read_msg(evals, sp->npts, base + pubs));
            if (sp->sharded && !c->shm.allreduce(c->H, evals, sp->npts)) {
                abort();
                ZK_FAIL(c, ZKB_ERR_NCCL, "shared-memory exchange: a peer rank stopped answering");
            }
            for (int i = 0; i < sp->npts; ++i) sp->last_evals[i] = evals[i];
            sp->have_evals = true;
            return ZKB_OK;
        }
        return sp_round_evals(c, sp, evals);
    }
    // Bind r; evals != NULL: the next round's s(0..d); NULL: this was the last variable, finals get the bound values.
    int32_t next(const Fe& r, Fe* evals, Fe* finals) {
        if (!live && sp->sharded && sp->cur_n <= gather_n()) ZK_TRY(sp_gather(c, sp));  // C2, then local rounds
        if (!live && !sp->have_evals) return sp_bind_and_next(c, sp, r, evals, finals);
        const bool go_small = !live && small_ok();
        if (!live && !go_small && !tail_ok()) return sp_bind_and_next(c, sp, r, evals, finals);
        const bool was_live = live;
        const bool tracing = c->trace && was_live && !small;
        unsigned long long t_send = 0;
        const unsigned long long t_recv_prev = t_recv;
        if (tracing) {
            t_send = host_ticks();
            if (t_recv) c->tr_host += (double)(t_send - t_recv) / tsc_per_us();
        }
        if (was_live && small && dt) {
            // the device drew this challenge itself after the round the host has just absorbed: they must agree
            if (!c->H.eq(c->dt_rounds[pubs - 1].chal, r)) {
                abort();
                ZK_FAIL(c, ZKB_ERR_CUDA, "device transcript diverged from the host transcript");
            }
            ++c->dt_rounds_checked;
            if (c->trace_verbose) {
                const long long* t = c->dt_rounds[pubs - 1].t;
                fprintf(stderr, "[zkb200 trace] device transcript step: interpolate+serialise %lld, absorb+keccak %lld, challenge %lld cycles; claim+record %lld, system fence %lld\n",
                        t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4]);
            }
        } else if (was_live) send(r);  // first, so the device works while the host finishes the claim chain
        if (poly_len < 0) poly_len = c->interp[sp->npts].interpolate(sp->last_evals, poly);
        const Fe claim = uni_evaluate(c->H, poly, poly_len, r);
        poly_len = -1;
        if (!was_live) {
            if (go_small) ZK_TRY(launch_small(false, r, &claim));
            else ZK_TRY(launch_tail(r));
        }
        ZK_TRY(wait_dev(base + (++pubs)));
        t_recv = c->trace ? host_ticks() : 0;
        if (tracing) {
            const unsigned long long* ts = c->mb->ts;
            // CTA 0's stamps are not ordered with the publishing CTA's flag: they may belong to the previous round,
            // so only differences of stamps written by the same thread are exact; the others are clamped
            auto d = [](unsigned long long a, unsigned long long b) {
                const long long x = (long long)(a - b);
                return x < 0 || x > 100000000ll ? 0.0 : (double)x * 1e-3;
            };
            c->tr_n += 1;
            c->tr_rtt += (double)(t_recv - t_send) / tsc_per_us();
            c->tr_relay += d(ts[1], ts[0]);
            c->tr_spread += d(ts[2], ts[1]);
            c->tr_pass += d(ts[3], ts[2]);
            c->tr_reduce += d(ts[4], ts[3]);
            if (c->trace_verbose)
                fprintf(stderr, "[zkb200 trace] n_in=2^%d send->result %.2f us: relay %.2f fan-out %.2f pass(CTA0) %.2f reduce/publish %.2f\n",
                        ilog2_u64(sp->cur_n), (double)(t_recv - t_send) / tsc_per_us(), d(ts[1], ts[0]), d(ts[2], ts[1]), d(ts[3], ts[2]), d(ts[4], ts[3]));
        }
        sp->cur_n /= 2;
        if (sp->cur_n == 1) {
            live = false;  // the kernel leaves after publishing the bound values
            const int T = (int)sp->sel.size();
            if (finals)
                for (int t = 0; t < T; ++t) finals[sp->sel[t]] = c->mb->finals[t];
            sp->have_evals = false;
            if (evals) ZK_FAIL(c, ZKB_ERR_ARITY, "bind_and_next: no next round after the last variable");
            return ZKB_OK;
        }
        if (!small && sp->cur_n <= stop_n) live = false;  // k_sc_tail handed over to k_sc_small
        if (!evals) {  // caller stops early: not supported inside a persistent kernel
            abort();
            ZK_FAIL(c, ZKB_ERR_BAD_ARG, "round driver: early stop inside the persistent kernel");
        }
        Fe got[MAXPTS];
        if (small && dt) for (int t = 0; t < sp->npts - 1; ++t) got[t] = c->dt_rounds[pubs - 1].evals[t];
        else ZK_TRY(read_msg(got, sp->npts - 1, base + pubs));
        const unsigned long long t_x0 = tracing ? host_ticks() : 0;
        if (sp->sharded && !c->shm.allreduce(c->H, got, sp->npts - 1)) {
            abort();
            ZK_FAIL(c, ZKB_ERR_NCCL, "shared-memory exchange: a peer rank stopped answering");
        }
        if (tracing) {
            const int lg = ilog2_u64(sp->cur_n) + 1;  // the round that has just been answered bound tables of 2^lg entries
            if (lg < 48) {
                c->trs_n[lg] += 1;
                c->trs_rtt[lg] += (double)(t_recv - t_send) / tsc_per_us();
                c->trs_wait[lg] += (double)(host_ticks() - t_x0) / tsc_per_us();
                if (t_recv_prev) c->trs_host[lg] += (double)(t_send - t_recv_prev) / tsc_per_us();
            }
        }
        evals[0] = got[0];
        evals[1] = c->H.sub(claim, evals[0]);
        for (int t = 2; t < sp->npts; ++t) evals[t] = got[t - 1];
        for (int i = 0; i < sp->npts; ++i) sp->last_evals[i] = evals[i];
        return ZKB_OK;
    }
};

// The composed sumcheck loop (sum_check_protocol.rs:86-115) over a configured state.
// coeffs: rounds x slots elements; returns challenges and the T bound values.
int32_t sp_prove(zkb_ctx* c, SumPolyState* sp, TranscriptImpl* tr, int slots, uint64_t* coeffs, int32_t* lens,
                 uint64_t* challenges, Fe* final_vals, bool as_evals) {
    sp_reset(c, sp);
    const int n_rounds = ilog2_u64(sp->n0) + (sp->sharded ? c->log2world : 0);
    const RoundInterpolator& ip = c->interp[sp->npts];
    Fe evals[MAXPTS], co[MAXPTS];
    if (n_rounds == 0) return sp_final_values(c, sp, final_vals);
    RoundDriver drv(c, sp);
    if (!as_evals) drv.tr = tr;
    ZK_TRY(drv.first(evals));
    for (int k = 0; k < n_rounds; ++k) {
        int len;
        if (as_evals) {  // plain sumcheck: the message is the evaluations themselves (:168-175)
            len = sp->npts;
            for (int i = 0; i < len; ++i) co[i] = evals[i];
        } else {
            len = ip.interpolate(evals, co);
            for (int i = 0; i < len; ++i) drv.poly[i] = co[i];
            drv.poly_len = len;
        }
        tr->append_elements(co, (size_t)len);
        if (lens) lens[k] = len;
        for (int i = 0; i < slots; ++i) fe_to_u64x4(i < len ? co[i] : c->H.zero(), coeffs + ((size_t)k * slots + i) * 4);
        Fe r = tr->challenge();
        if (challenges) fe_to_u64x4(r, challenges + (size_t)k * 4);
        ZK_TRY(drv.next(r, k + 1 < n_rounds ? evals : nullptr, final_vals));
    }
    return ZKB_OK;
}

// ------------------------------------------------------------------ MLE helpers
constexpr uint64_t UP_CHUNK = 1ull << 19;  // elements per upload chunk (16 MiB)
int32_t upload_aos(zkb_ctx* c, const uint64_t* aos, uint64_t n_src, uint64_t first, uint64_t stride, uint64_t n_dst,
                   int conv, Table* out) {
    ZK_TRY(alloc_table(c, n_dst, out));
    (void)n_src;
    if (stride > 1) {
        // strided shard: copy only this rank's elements with a 2D copy (32-byte rows)
        ZK_TRY(ensure_stage(c, (size_t)n_dst * 32));
        ZK_CUDA(c, cudaMemcpy2DAsync(c->stage, 32, (const uint8_t*)aos + first * 32, (size_t)stride * 32, 32, n_dst,
                                     cudaMemcpyHostToDevice, c->stream));
        prof_begin(c, ZKB_K_LAYOUT, 64.0 * (double)n_dst);
        c->K->aos_to_planar(c->stage, out->ref(), n_dst, 0, 1, conv, grid_for(c, n_dst, 8), c->stream);
        ZK_TRY(check_launch(c, "k_aos_to_planar"));
        return ZKB_OK;
    }
    if (n_dst <= UP_CHUNK) {  // small table: one copy, one kernel, one stream
        ZK_TRY(ensure_stage(c, (size_t)n_dst * 32));
        ZK_CUDA(c, cudaMemcpyAsync(c->stage, (const uint8_t*)aos + first * 32, (size_t)n_dst * 32, cudaMemcpyHostToDevice, c->stream));
        prof_begin(c, ZKB_K_LAYOUT, 64.0 * (double)n_dst);
        c->K->aos_to_planar(c->stage, out->ref(), n_dst, 0, 1, conv, grid_for(c, n_dst, 8), c->stream);
        ZK_TRY(check_launch(c, "k_aos_to_planar"));
        return ZKB_OK;
    }
    // large table: 16 MiB chunks through two staging buffers; the copy engine never waits for a layout kernel
    if (!c->copy_stream) {
        ZK_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            ZK_CUDA(c, cudaMalloc(&c->up_stage[b], (size_t)UP_CHUNK * 32));
            ZK_CUDA(c, cudaEventCreateWithFlags(&c->ev_copied[b], cudaEventDisableTiming));
            ZK_CUDA(c, cudaEventCreateWithFlags(&c->ev_free[b], cudaEventDisableTiming));
        }
        ZK_CUDA(c, cudaEventCreateWithFlags(&c->ev_entry, cudaEventDisableTiming));
    }
    ZK_CUDA(c, cudaEventRecord(c->ev_entry, c->stream));  // earlier users of the staging buffers are ordered before us
    ZK_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_entry, 0));
    uint64_t i = 0;
    for (uint64_t off = 0; off < n_dst; off += UP_CHUNK, ++i) {
        const uint64_t m = (n_dst - off < UP_CHUNK) ? n_dst - off : UP_CHUNK;
        const int b = (int)(i & 1);
        if (i >= 2) ZK_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_free[b], 0));
        ZK_CUDA(c, cudaMemcpyAsync(c->up_stage[b], (const uint8_t*)aos + (first + off) * 32, (size_t)m * 32, cudaMemcpyHostToDevice,
                                   c->copy_stream));
        ZK_CUDA(c, cudaEventRecord(c->ev_copied[b], c->copy_stream));
        ZK_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
        TabRef dst{out->base + off, out->stride};
        prof_begin(c, ZKB_K_LAYOUT, 64.0 * (double)m);
        c->K->aos_to_planar(c->up_stage[b], dst, m, 0, 1, conv, grid_for(c, m, 8), c->stream);
        ZK_TRY(check_launch(c, "k_aos_to_planar"));
        ZK_CUDA(c, cudaEventRecord(c->ev_free[b], c->stream));
    }
    return ZKB_OK;
}

int32_t download_aos(zkb_ctx* c, const Table& t, void* host, int conv) {
    const uint64_t chunk = t.n < (1ull << 21) ? t.n : (1ull << 21);
    ZK_TRY(ensure_stage(c, (size_t)chunk * 32));
    for (uint64_t off = 0; off < t.n; off += chunk) {
        const uint64_t m = (t.n - off < chunk) ? t.n - off : chunk;
        TabRef src{t.base + off, t.stride};
        prof_begin(c, ZKB_K_LAYOUT, 64.0 * (double)m);
        c->K->planar_to_aos(src, c->stage, m, conv, grid_for(c, m, 8), c->stream);
        ZK_TRY(check_launch(c, "k_planar_to_aos"));
        ZK_CUDA(c, cudaMemcpyAsync((uint8_t*)host + off * 32, c->stage, (size_t)m * 32, cudaMemcpyDeviceToHost, c->stream));
    }
    ZK_CUDA(c, cudaStreamSynchronize(c->stream));
    return ZKB_OK;
}

// evaluate / multi_partial_evaluate: bind the k leading variables of `src` into a fresh table.
int32_t multi_fold(zkb_ctx* c, const Table& src, const Fe* rs, uint32_t k, Table* out) {
    const int nv = ilog2_u64(src.n);
    if ((int)k > nv) ZK_FAIL(c, ZKB_ERR_ARITY, "Invalid number of values");
    Table w;
    if (k == 0) {
        ZK_TRY(alloc_table(c, src.n, &w));
        ZK_CUDA(c, cudaMemcpyAsync(w.base, src.base, src.n * 16, cudaMemcpyDeviceToDevice, c->stream));
        ZK_CUDA(c, cudaMemcpyAsync(w.base + w.stride, src.base + src.stride, src.n * 16, cudaMemcpyDeviceToDevice, c->stream));
        *out = w;
        return ZKB_OK;
    }
    // K3: three variables per pass while possible (36 B of traffic per input entry instead of 96)
    uint64_t n = src.n;
    const uint32_t first = k >= 3 ? 3 : k;
    ZK_TRY(alloc_table(c, src.n >> first, &w));
    for (uint32_t i = 0; i < k;) {
        const uint32_t kk = k - i >= 3 ? 3 : k - i;
        if (kk == 3 && c->tc_enabled && n >= (1ull << 17)) {
            // the bound entry is a linear combination of 8 inputs with weights prod_l (bit_l ? r_l : 1 - r_l): tensor cores
            auto ta = std::make_unique<MultiFoldTcArgs>();
            ta->in = i == 0 ? src.ref() : w.ref();
            ta->out = w.ref();
            ta->n_out = n >> 3;
            Fe omr[3];
            for (int l = 0; l < 3; ++l) omr[l] = c->H.sub(c->tcm.c[32], rs[i + l]);
            for (int x = 0; x < 8; ++x) {
                Fe wgt = (x & 4) ? rs[i] : omr[0];
                wgt = c->H.mul(wgt, (x & 2) ? rs[i + 1] : omr[1]);
                wgt = c->H.mul(wgt, (x & 1) ? rs[i + 2] : omr[2]);
                c->tcm.make_weight(c->H, wgt, ta->mats[x]);
            }
            prof_begin(c, ZKB_K_FOLD_TABLES, 32.0 * (double)n + 32.0 * (double)(n >> 3));
            const uint64_t tiles = (n >> 3) >> 7;
            c->K->multifold_tc(*ta, (int)(tiles < (uint64_t)(2 * c->sm_count) ? tiles : (uint64_t)(2 * c->sm_count)), c->stream);
            ZK_TRY(check_launch(c, "k_multifold_tc"));
            n >>= 3;
            i += 3;
            continue;
        }
        MultiFoldArgs fa;
        std::memset(&fa, 0, sizeof fa);
        fa.in = i == 0 ? src.ref() : w.ref();
        fa.out = w.ref();
        fa.n_out = n >> kk;
        for (uint32_t l = 0; l < kk; ++l) c->fmb.make(c->H, rs[i + l], &fa.rt[l]);
        prof_begin(c, ZKB_K_FOLD_TABLES, 32.0 * (double)n + 32.0 * (double)(n >> kk));
        c->K->multifold((int)kk, fa, grid_for(c, n >> kk, 2), c->stream);
        ZK_TRY(check_launch(c, "k_multifold"));
        n >>= kk;
        i += kk;
    }
    w.n = n;
    *out = w;
    return ZKB_OK;
}

int32_t read_elems(zkb_ctx* c, const Table& t, uint64_t first, int count, Fe* out) {
    GatherArgs g;
    for (int i = 0; i < count; i += MAXT) {
        std::memset(&g, 0, sizeof g);
        int m = count - i < MAXT ? count - i : MAXT;
        for (int j = 0; j < m; ++j) {
            g.t[j] = TabRef{t.base + first + i + j, t.stride};
        }
        g.n = m;
        g.idx = 0;
        g.out = c->d_res;
        launch_gather_elems(g, c->stream);
        ZK_TRY(check_launch(c, "k_gather_elems"));
        ZK_CUDA(c, cudaMemcpyAsync(c->h_res, c->d_res, sizeof(Fe) * m, cudaMemcpyDeviceToHost, c->stream));
        ZK_CUDA(c, cudaStreamSynchronize(c->stream));
        for (int j = 0; j < m; ++j) out[i + j] = c->h_res[j];
    }
    return ZKB_OK;
}

// Full evaluation of a (possibly sharded) table at k = log2(global size) values.
int32_t evaluate_table(zkb_ctx* c, const Table& src, const Fe* rs, uint32_t k, Fe* out) {
    const int nv_local = ilog2_u64(src.n);
    const bool sharded = c->comm && c->world > 1;
    const int nv = nv_local + (sharded ? c->log2world : 0);
    if ((int)k != nv) ZK_FAIL(c, ZKB_ERR_ARITY, "Invalid number of values");
    Table w;
    ZK_TRY(multi_fold(c, src, rs, (uint32_t)nv_local, &w));
    Fe v;
    int32_t st = read_elems(c, w, 0, 1, &v);
    free_table(c, &w);
    ZK_TRY(st);
    if (sharded) {
        // the remaining log2(world) variables index the rank: gather the per-rank values, finish on the host
        Fe* dbuf = c->d_res;
        ZK_CUDA(c, cudaMemcpyAsync(dbuf + c->world, &v, sizeof(Fe), cudaMemcpyHostToDevice, c->stream));
        ZK_NCCL(c, g_nccl.AllGather(dbuf + c->world, dbuf, sizeof(Fe), ncclUint8, c->comm, c->stream));
        ZK_CUDA(c, cudaMemcpyAsync(c->h_res, dbuf, sizeof(Fe) * c->world, cudaMemcpyDeviceToHost, c->stream));
        ZK_CUDA(c, cudaStreamSynchronize(c->stream));
        std::vector<Fe> vals(c->h_res, c->h_res + c->world);
        for (int i = 0; i < c->log2world; ++i) {
            size_t h = vals.size() / 2;
            const Fe& r = rs[nv_local + i];
            for (size_t j = 0; j < h; ++j) vals[j] = c->H.add(vals[j], c->H.mul(r, c->H.sub(vals[j + h], vals[j])));
            vals.resize(h);
        }
        v = vals[0];
    }
    *out = v;
    return ZKB_OK;
}

// Canonical little-endian bytes of a whole table into the transcript.  Keccak is sequential and runs on the host
// (about 0.4 GB/s), so the table is converted and copied in 2 MiB pieces through two pinned slots while the
// previous piece is being hashed; no table-sized host allocation.
int32_t absorb_table_bytes(zkb_ctx* c, const Table& t, TranscriptImpl* tr) {
    const uint64_t chunk = 1ull << 16;  // elements per piece (2 MiB)
    ZK_TRY(ensure_stage(c, (size_t)2 * chunk * 32));
    if (!c->hash_pin) {
        ZK_CUDA(c, cudaHostAlloc(&c->hash_pin, (size_t)2 * chunk * 32, cudaHostAllocDefault));
        for (auto& e : c->hash_ev) ZK_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const uint64_t pieces = (t.n + chunk - 1) / chunk;
    for (uint64_t k = 0; k <= pieces; ++k) {
        if (k < pieces) {
            const uint64_t off = k * chunk, m = t.n - off < chunk ? t.n - off : chunk;
            uint4* dslot = (uint4*)c->stage + (k & 1) * chunk * 2;
            TabRef src{t.base + off, t.stride};
            prof_begin(c, ZKB_K_LAYOUT, 64.0 * (double)m);
            c->K->planar_to_aos(src, dslot, m, 2, grid_for(c, m, 8), c->stream);
            ZK_TRY(check_launch(c, "k_planar_to_aos"));
            ZK_CUDA(c, cudaMemcpyAsync((uint8_t*)c->hash_pin + (k & 1) * chunk * 32, dslot, (size_t)m * 32, cudaMemcpyDeviceToHost, c->stream));
            ZK_CUDA(c, cudaEventRecord(c->hash_ev[k & 1], c->stream));
        }
        if (k > 0) {
            const uint64_t j = k - 1, off = j * chunk, m = t.n - off < chunk ? t.n - off : chunk;
            ZK_CUDA(c, cudaEventSynchronize(c->hash_ev[j & 1]));
            tr->append((const uint8_t*)c->hash_pin + (j & 1) * chunk * 32, (size_t)m * 32);
        }
    }
    return ZKB_OK;
}

// ------------------------------------------------------------------- GKR core
uint32_t layer_rounds(uint32_t gates) {
    int nb = ilog2_u64(2ull * gates);
    if (nb < 1) nb = 1;
    return 2u * (uint32_t)nb;
}

int32_t circuit_check_shape(zkb_ctx* c, uint32_t n_layers, const uint32_t* g) {
    if (n_layers < 1) ZK_FAIL(c, ZKB_ERR_CIRCUIT_SHAPE, "circuit: no layers");
    if (g[n_layers - 1] != 1 && g[n_layers - 1] != 2) ZK_FAIL(c, ZKB_ERR_CIRCUIT_SHAPE, "circuit: output layer must have 1 or 2 gates (gkr_protocol.rs:235)");
    for (uint32_t l = 0; l + 1 < n_layers; ++l)
        if (g[l] != 2 * g[l + 1]) ZK_FAIL(c, ZKB_ERR_CIRCUIT_SHAPE, "circuit: each layer must have twice the gates of the next (gkr_circuit.rs:76-78)");
    return ZKB_OK;
}

int32_t circuit_run(zkb_ctx* c, CircuitState* cs, const uint64_t* inputs_mont, uint64_t n_inputs) {
    if (cs->wired) {
        if (n_inputs != cs->n_inputs) ZK_FAIL(c, ZKB_ERR_LENGTH_MISMATCH, "circuit: wrong number of inputs");
    } else if (n_inputs != 2ull * cs->gates[0])
        ZK_FAIL(c, ZKB_ERR_LENGTH_MISMATCH, "circuit: inputs must be 2 x gates of the first layer");
    // inputs: AoS Montgomery -> planar
    if (cs->aos_stage_bytes < n_inputs * 32) {
        if (cs->aos_stage) cudaFreeAsync(cs->aos_stage, c->stream);
        ZK_CUDA(c, cudaMallocAsync(&cs->aos_stage, n_inputs * 32, c->stream));
        cs->aos_stage_bytes = n_inputs * 32;
    }
    ZK_CUDA(c, cudaMemcpyAsync(cs->aos_stage, inputs_mont, n_inputs * 32, cudaMemcpyHostToDevice, c->stream));
    c->K->aos_to_planar(cs->aos_stage, cs->inputs.ref(), n_inputs, 0, 1, 0, grid_for(c, n_inputs, 8), c->stream);
    ZK_TRY(check_launch(c, "k_aos_to_planar"));
    for (int l = 0; l < cs->L; ++l) {
        const Table& in = l == 0 ? cs->inputs : cs->vals[l - 1];
        if (cs->wired) {
            c->K->layer_eval_w(in.ref(), cs->vals[l].ref(), cs->d_ops + cs->opoff[l], cs->d_in1 + cs->opoff[l], cs->d_in2 + cs->opoff[l],
                               cs->gates[l], grid_for(c, cs->gates[l], 8), c->stream);
            ZK_TRY(check_launch(c, "k_layer_eval_w"));
            continue;
        }
        c->K->layer_eval(in.ref(), cs->vals[l].ref(), cs->d_ops + cs->opoff[l], cs->gates[l], grid_for(c, cs->gates[l], 8), c->stream);
        ZK_TRY(check_launch(c, "k_layer_eval"));
    }
    return ZKB_OK;
}

int32_t eq_tables(zkb_ctx* c, const Fe* r, int n, Table* hi, Table* lo, int* n_lo_out) {
    ChalList cl;
    std::memset(&cl, 0, sizeof cl);
    for (int i = 0; i < n; ++i) cl.r[i] = r[i];
    const int n_hi = n / 2, n_lo = n - n_hi;
    c->K->eq_split(cl, n, n_hi, hi->ref(), lo->ref(), grid_for(c, (1ull << n_hi) + (1ull << n_lo), 8), c->stream);
    ZK_TRY(check_launch(c, "k_eq_split"));
    *n_lo_out = n_lo;
    return ZKB_OK;
}

// One phase (nb rounds) of the two-phase layer sumcheck over X*Y + Z.
int32_t xyz_phase(zkb_ctx* c, CircuitState* cs, const Table& X, const Table& Y, const Table& Z, uint64_t nw,
                  TranscriptImpl* tr, uint64_t* coeffs, int32_t* lens, uint64_t* chals, Fe* point, Fe* x_final) {
    SumPolyState* sp = &cs->sp;
    sp->src.resize(3);
    sp->src[0] = X;
    sp->src[1] = Y;
    sp->src[2] = Z;
    for (auto& t : sp->src) t.n = nw;
    sp->n0 = nw;
    sp_reset(c, sp);
    sp->sharded = false;  // GKR layers are never sharded (replicas only)
    const int nb = ilog2_u64(nw);
    const RoundInterpolator& ip = c->interp[3];
    Fe evals[3], co[3], fin[3];
    RoundDriver drv(c, sp);
    drv.tr = tr;
    ZK_TRY(drv.first(evals));
    for (int k = 0; k < nb; ++k) {
        int len = ip.interpolate(evals, co);
        for (int i = 0; i < len; ++i) drv.poly[i] = co[i];
        drv.poly_len = len;
        tr->append_elements(co, (size_t)len);
        lens[k] = len;
        for (int i = 0; i < 3; ++i) fe_to_u64x4(i < len ? co[i] : c->H.zero(), coeffs + ((size_t)k * 3 + i) * 4);
        Fe r = tr->challenge();
        point[k] = r;
        if (chals) fe_to_u64x4(r, chals + (size_t)k * 4);
        ZK_TRY(drv.next(r, k + 1 < nb ? evals : nullptr, fin));
    }
    *x_final = fin[0];
    return ZKB_OK;
}

struct GkrLayerCtx {
    Fe r0, alpha, beta;
    std::vector<Fe> rb, rc;
};

int32_t gkr_prove_impl(zkb_ctx* c, CircuitState* cs, const uint64_t* inputs_mont, uint64_t n_inputs, uint64_t* w0_out,
                       uint64_t* coeffs, int32_t* lens, uint64_t* challenges, uint64_t* claimed, uint64_t* final_openings,
                       uint32_t* n_rounds_out) {
    ZK_TRY(circuit_run(c, cs, inputs_mont, n_inputs));
    const int L = cs->L;
    TranscriptImpl tr;
    tr.H = c->H;
    // w_0 padded to two entries (gkr_protocol.rs:34-39); initiate_protocol (:229-241)
    Fe w0[2] = {c->H.zero(), c->H.zero()};
    ZK_TRY(read_elems(c, cs->vals[L - 1], 0, (int)cs->gates[L - 1], w0));
    fe_to_u64x4(w0[0], w0_out);
    fe_to_u64x4(w0[1], w0_out + 4);
    tr.append_elements(w0, 2);
    Fe r0 = tr.challenge();
    Fe m0 = c->H.add(w0[0], c->H.mul(r0, c->H.sub(w0[1], w0[0])));
    tr.append_elements(&m0, 1);

    Fe alpha = c->H.zero(), beta = c->H.zero();
    std::vector<Fe> rb, rc;
    size_t round_base = 0;
    Fe o1 = c->H.zero(), o2 = c->H.zero();
    for (int idx = 0; idx < L; ++idx) {
        const int l = L - 1 - idx;
        const uint64_t G = cs->gates[l], nw = 2 * G;
        const int nb = ilog2_u64(nw);
        const Table& W = l == 0 ? cs->inputs : cs->vals[l - 1];
        GkrP1Args p1;
        std::memset(&p1, 0, sizeof p1);
        p1.W = W.ref();
        p1.H1 = cs->H1.ref();
        p1.HA2 = cs->HA2.ref();
        p1.coef = cs->coef.ref();
        p1.ops = cs->d_ops + cs->opoff[l];
        p1.n_gates = G;
        p1.first_layer = idx == 0;
        p1.r0 = r0;
        p1.alpha = alpha;
        p1.beta = beta;
        if (idx > 0) {
            if ((1ull << rb.size()) != G) ZK_FAIL(c, ZKB_ERR_CIRCUIT_SHAPE, "gkr: challenge count does not match the layer width");
            int n_lo = 0;
            ZK_TRY(eq_tables(c, rb.data(), (int)rb.size(), &cs->eq[0], &cs->eq[1], &n_lo));
            ZK_TRY(eq_tables(c, rc.data(), (int)rc.size(), &cs->eq[2], &cs->eq[3], &n_lo));
            p1.eb_hi = cs->eq[0].ref();
            p1.eb_lo = cs->eq[1].ref();
            p1.ec_hi = cs->eq[2].ref();
            p1.ec_lo = cs->eq[3].ref();
            p1.n_lo = n_lo;
        }
        prof_begin(c, ZKB_K_GKR_BUILD, 32.0 * (double)G * 6.0);
        c->K->gkr_phase1(p1, grid_for(c, G, 8), c->stream);
        ZK_TRY(check_launch(c, "k_gkr_phase1"));
        std::vector<Fe> u(nb), v(nb);
        Fe Wu, Wv;
        ZK_TRY(xyz_phase(c, cs, W, cs->H1, cs->HA2, nw, &tr, coeffs + round_base * 12, lens + round_base,
                         challenges ? challenges + round_base * 4 : nullptr, u.data(), &Wu));
        // phase 2: b bound to u
        int n_lo_u = 0;
        ZK_TRY(eq_tables(c, u.data(), nb, &cs->eq[4], &cs->eq[5], &n_lo_u));
        GkrP2Args p2;
        std::memset(&p2, 0, sizeof p2);
        p2.C = cs->H1.ref();
        p2.D = cs->HA2.ref();
        p2.coef = cs->coef.ref();
        p2.eu_hi = cs->eq[4].ref();
        p2.eu_lo = cs->eq[5].ref();
        p2.n_lo = n_lo_u;
        p2.ops = cs->d_ops + cs->opoff[l];
        p2.n_gates = G;
        p2.Wu = Wu;
        prof_begin(c, ZKB_K_GKR_BUILD, 32.0 * (double)G * 5.0);
        c->K->gkr_phase2(p2, grid_for(c, G, 8), c->stream);
        ZK_TRY(check_launch(c, "k_gkr_phase2"));
        ZK_TRY(xyz_phase(c, cs, W, cs->H1, cs->HA2, nw, &tr, coeffs + (round_base + nb) * 12, lens + round_base + nb,
                         challenges ? challenges + (round_base + nb) * 4 : nullptr, v.data(), &Wv));
        round_base += 2 * (size_t)nb;
        rb = u;
        rc = v;
        o1 = Wu;
        o2 = Wv;
        if (idx < L - 1) {  // :80-89
            tr.append_elements(&o1, 1);
            alpha = tr.challenge();
            tr.append_elements(&o2, 1);
            beta = tr.challenge();
            fe_to_u64x4(o1, claimed + (size_t)idx * 8);
            fe_to_u64x4(o2, claimed + (size_t)idx * 8 + 4);
        }
    }
    fe_to_u64x4(o1, final_openings);
    fe_to_u64x4(o2, final_openings + 4);
    if (n_rounds_out) *n_rounds_out = (uint32_t)round_base;
    return ZKB_OK;
}

// ------------------------------------------------------------------ general wiring (extension)
// The reference fixes gate i's inputs to wires (2i, 2i+1) (gkr_circuit.rs:76-78,132), which forces every layer to
// be half as wide as the one below and the output layer to 1-2 gates.  BASELINE configs[2] ("2^20 gates per layer
// and 16 layers") needs arbitrary wiring in1[g], in2[g] and a wide output layer; the protocol is the reference's
// (same transcript order, same round polynomials, claim merge and openings), with the single output challenge
// r0 replaced by log2(outputs) consecutive challenges.  With in1 = 2g, in2 = 2g+1 and <= 2 outputs this path
// produces the bytes of gkr_prove_impl (asserted in tests/test_gpu_parity.py).
void wired_coef_args(CircuitState* cs, int idx, const Fe& alpha, const Fe& beta, int n_lo, WiredCoef* wc) {
    std::memset(wc, 0, sizeof *wc);
    wc->a1_hi = cs->eq[0].ref();
    wc->a1_lo = cs->eq[1].ref();
    wc->a2_hi = cs->eq[2].ref();
    wc->a2_lo = cs->eq[3].ref();
    wc->n_lo = n_lo;
    wc->two = idx > 0;
    wc->alpha = alpha;
    wc->beta = beta;
}

int32_t gkr_prove_wired_impl(zkb_ctx* c, CircuitState* cs, const uint64_t* inputs_mont, uint64_t n_inputs, uint64_t* w0_out,
                             uint64_t* coeffs, int32_t* lens, uint64_t* challenges, uint64_t* claimed, uint64_t* final_openings,
                             uint32_t* n_rounds_out) {
    ZK_TRY(circuit_run(c, cs, inputs_mont, n_inputs));
    const int L = cs->L;
    const HostField& H = c->H;
    TranscriptImpl tr;
    tr.H = H;
    // initiate_protocol (:229-241) with one challenge per output variable
    const uint64_t Gout = cs->gates[L - 1], n0 = Gout < 2 ? 2 : Gout;
    const int k0 = ilog2_u64(n0);
    std::vector<Fe> r0(k0);
    Fe m0;
    if (Gout == 1) {
        Fe w0[2] = {H.zero(), H.zero()};
        ZK_TRY(read_elems(c, cs->vals[L - 1], 0, 1, w0));
        fe_to_u64x4(w0[0], w0_out);
        fe_to_u64x4(w0[1], w0_out + 4);
        tr.append_elements(w0, 2);
        r0[0] = tr.challenge();
        m0 = H.add(w0[0], H.mul(r0[0], H.sub(w0[1], w0[0])));
    } else {
        ZK_TRY(download_aos(c, cs->vals[L - 1], w0_out, 0));
        ZK_TRY(absorb_table_bytes(c, cs->vals[L - 1], &tr));  // fq_vec_to_bytes on the device, Keccak on the host
        for (int i = 0; i < k0; ++i) r0[i] = tr.challenge();
        ZK_TRY(evaluate_table(c, cs->vals[L - 1], r0.data(), (uint32_t)k0, &m0));
    }
    tr.append_elements(&m0, 1);

    Fe alpha = H.zero(), beta = H.zero();
    std::vector<Fe> rb, rc;
    size_t round_base = 0;
    Fe o1 = H.zero(), o2 = H.zero();
    for (int idx = 0; idx < L; ++idx) {
        const int l = L - 1 - idx;
        const uint64_t G = cs->gates[l], nw = cs->width[l];
        const int nb = ilog2_u64(nw);
        const Table& W = l == 0 ? cs->inputs : cs->vals[l - 1];
        int n_lo = 0;
        if (idx == 0) {
            ZK_TRY(eq_tables(c, r0.data(), k0, &cs->eq[0], &cs->eq[1], &n_lo));
        } else {
            if ((1ull << rb.size()) != G) ZK_FAIL(c, ZKB_ERR_CIRCUIT_SHAPE, "gkr: challenge count does not match the layer width");
            ZK_TRY(eq_tables(c, rb.data(), (int)rb.size(), &cs->eq[0], &cs->eq[1], &n_lo));
            ZK_TRY(eq_tables(c, rc.data(), (int)rc.size(), &cs->eq[2], &cs->eq[3], &n_lo));
        }
        GkrW1Args p1;
        std::memset(&p1, 0, sizeof p1);
        p1.W = W.ref();
        p1.H1 = cs->H1.ref();
        p1.HA2 = cs->HA2.ref();
        p1.coef = cs->coef.ref();
        wired_coef_args(cs, idx, alpha, beta, n_lo, &p1.wc);
        p1.ops = cs->d_ops + cs->opoff[l];
        p1.in2 = cs->d_in2 + cs->opoff[l];
        p1.off1 = cs->d_off1 + cs->csroff[l];
        p1.lst1 = cs->d_lst1 + cs->opoff[l];
        p1.tmp = cs->gtmp.ref();
        p1.width = nw;
        p1.n_gates = G;
        prof_begin(c, ZKB_K_GKR_BUILD, 32.0 * ((double)G * 5.0 + (double)nw * 2.0));
        c->K->gkr_w_phase1(p1, grid_for(c, G, 8), grid_for(c, nw, 8), c->stream);
        ++c->launches;
        ZK_TRY(check_launch(c, "k_gkr_w_phase1"));
        std::vector<Fe> u(nb), v(nb);
        Fe Wu, Wv;
        ZK_TRY(xyz_phase(c, cs, W, cs->H1, cs->HA2, nw, &tr, coeffs + round_base * 12, lens + round_base,
                         challenges ? challenges + round_base * 4 : nullptr, u.data(), &Wu));
        int n_lo_u = 0;
        ZK_TRY(eq_tables(c, u.data(), nb, &cs->eq[4], &cs->eq[5], &n_lo_u));
        GkrW2Args p2;
        std::memset(&p2, 0, sizeof p2);
        p2.C = cs->H1.ref();
        p2.D = cs->HA2.ref();
        p2.coef = cs->coef.ref();
        p2.eu_hi = cs->eq[4].ref();
        p2.eu_lo = cs->eq[5].ref();
        p2.n_lo = n_lo_u;
        p2.ops = cs->d_ops + cs->opoff[l];
        p2.in1 = cs->d_in1 + cs->opoff[l];
        p2.off2 = cs->d_off2 + cs->csroff[l];
        p2.lst2 = cs->d_lst2 + cs->opoff[l];
        p2.tmp = cs->gtmp.ref();
        p2.width = nw;
        p2.n_gates = G;
        p2.Wu = Wu;
        prof_begin(c, ZKB_K_GKR_BUILD, 32.0 * ((double)G * 3.0 + (double)nw * 2.0));
        c->K->gkr_w_phase2(p2, grid_for(c, G, 8), grid_for(c, nw, 8), c->stream);
        ++c->launches;
        ZK_TRY(check_launch(c, "k_gkr_w_phase2"));
        ZK_TRY(xyz_phase(c, cs, W, cs->H1, cs->HA2, nw, &tr, coeffs + (round_base + nb) * 12, lens + round_base + nb,
                         challenges ? challenges + (round_base + nb) * 4 : nullptr, v.data(), &Wv));
        round_base += 2 * (size_t)nb;
        rb = u;
        rc = v;
        o1 = Wu;
        o2 = Wv;
        if (idx < L - 1) {  // :80-89
            tr.append_elements(&o1, 1);
            alpha = tr.challenge();
            tr.append_elements(&o2, 1);
            beta = tr.challenge();
            fe_to_u64x4(o1, claimed + (size_t)idx * 8);
            fe_to_u64x4(o2, claimed + (size_t)idx * 8 + 4);
        }
    }
    fe_to_u64x4(o1, final_openings);
    fe_to_u64x4(o2, final_openings + 4);
    if (n_rounds_out) *n_rounds_out = (uint32_t)round_base;
    return ZKB_OK;
}

// Untrusted proof elements must be canonical Montgomery residues (limbs < p): the host arithmetic and H.eq assume it,
// and a proof carrying c + p in place of c would otherwise hash as c and could be accepted (malleability).  ark-ff
// cannot produce such a value; a C caller can.
bool all_canonical(const HostField& H, const uint64_t* v, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        const uint64_t* x = v + 4 * i;
        bool lt = false;
        for (int k = 3; k >= 0; --k) {
            if (x[k] != H.p[k]) {
                lt = x[k] < H.p[k];
                break;
            }
        }
        if (!lt) return false;
    }
    return true;
}
bool coeffs_canonical(const HostField& H, const uint64_t* coeffs, const int32_t* lens, uint32_t n_rounds, uint32_t slots) {
    for (uint32_t k = 0; k < n_rounds; ++k) {
        if (lens[k] < 0 || (uint32_t)lens[k] > slots) return false;
        if (!all_canonical(H, coeffs + (size_t)k * slots * 4, (size_t)lens[k])) return false;
    }
    return true;
}

// gkr_verify (sum_check_protocol.rs:117-150) on the host.
bool sc_verify_rounds(const HostField& H, TranscriptImpl* tr, uint32_t n_rounds, uint32_t slots, const uint64_t* coeffs,
                      const int32_t* lens, Fe claim, Fe* final_claim, Fe* chals) {
    const Fe one = H.one();
    for (uint32_t k = 0; k < n_rounds; ++k) {
        Fe co[MAXPTS + 3];
        const int len = lens[k];
        for (int i = 0; i < len; ++i) co[i] = fe_from_u64x4(coeffs + ((size_t)k * slots + i) * 4);
        Fe p0 = uni_evaluate(H, co, len, H.zero());
        Fe p1 = uni_evaluate(H, co, len, one);
        if (!H.eq(H.add(p0, p1), claim)) return false;
        tr->append_elements(co, (size_t)len);
        Fe r = tr->challenge();
        chals[k] = r;
        claim = uni_evaluate(H, co, len, r);
    }
    *final_claim = claim;
    return true;
}


// ------------------------------------------------------------------- KZG core (kzg_impl.cuh)
void kzg_release(zkb_ctx* c, KzgState* k) {
    for (auto* p : k->basis)
        if (p) cudaFreeAsync(p, c->stream);
    k->basis.clear();
}
MsmPlan msm_plan(uint64_t n) {
    int lg = ilog2_u64(n);
    int cbits = lg - 4;
    if (cbits < 4) cbits = 4;
    if (cbits > 16) cbits = 16;
    MsmPlan pl;
    pl.c = (uint32_t)cbits;
    pl.windows = (255 + pl.c - 1) / pl.c;
    const uint32_t digits = 1u << pl.c;
    pl.chunks = digits / 64 ? digits / 64 : 1;
    return pl;
}
// sum_i bases[i] * scalars[i] -> 96 canonical affine bytes in d_out (device).  scalars: a Montgomery Fr table.
// (st: the stream the whole MSM runs on, scratch included -- the quotient commitments of get_proof run on side streams)
int32_t msm_run(zkb_ctx* c, const Table& scalars, const G1Affine* bases, uint64_t n, uint8_t* d_out, cudaStream_t st) {
    if (n <= 256) {
        G1Jac* scratch = nullptr;
        ZK_CUDA(c, cudaMallocAsync((void**)&scratch, sizeof(G1Jac) * 256, st));
        k_msm_small<<<1, 256, 0, st>>>(scalars.ref(), bases, (uint32_t)n, scratch, d_out);
        ZK_TRY(check_launch(c, "k_msm_small"));
        cudaFreeAsync(scratch, st);
        return ZKB_OK;
    }
    const MsmPlan pl = msm_plan(n);
    const uint64_t m = (uint64_t)pl.windows * n;
    if (m >= (1ull << 32)) ZK_FAIL(c, ZKB_ERR_UNSUPPORTED, "msm: more than 2^32 (window, point) pairs");
    const uint32_t nb = pl.windows << pl.c;
    uint32_t *keys = nullptr, *vals = nullptr, *keys2 = nullptr, *vals2 = nullptr, *start = nullptr, *heavy = nullptr;
    G1Jac *buckets = nullptr, *parts = nullptr, *wsum = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    int key_bits = (int)pl.c;
    while ((1u << (key_bits - (int)pl.c)) < pl.windows) ++key_bits;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, vals, vals2, (int)m, 0, key_bits, st);
    ZK_CUDA(c, cudaMallocAsync((void**)&keys, m * 4, st));
    ZK_CUDA(c, cudaMallocAsync((void**)&vals, m * 4, st));
    ZK_CUDA(c, cudaMallocAsync((void**)&keys2, m * 4, st));
    ZK_CUDA(c, cudaMallocAsync((void**)&vals2, m * 4, st));
    ZK_CUDA(c, cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 16, st));
    ZK_CUDA(c, cudaMallocAsync((void**)&start, (size_t)nb * 8, st));  // start[nb], end[nb]
    ZK_CUDA(c, cudaMallocAsync((void**)&heavy, ((size_t)nb + 1) * 4, st));  // count, then the list
    ZK_CUDA(c, cudaMallocAsync((void**)&buckets, sizeof(G1Jac) * nb, st));
    ZK_CUDA(c, cudaMallocAsync((void**)&parts, sizeof(G1Jac) * pl.windows * pl.chunks, st));
    ZK_CUDA(c, cudaMallocAsync((void**)&wsum, sizeof(G1Jac) * pl.windows, st));
    ZK_CUDA(c, cudaMemsetAsync(start, 0, (size_t)nb * 8, st));
    ZK_CUDA(c, cudaMemsetAsync(heavy, 0, 4, st));
    uint32_t* end = start + nb;
    k_msm_digits<<<grid_for(c, n, 8), BLOCK, 0, st>>>(scalars.ref(), n, pl, keys, vals);
    ZK_TRY(check_launch(c, "k_msm_digits"));
    if (cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, vals, vals2, (int)m, 0, key_bits, st) != cudaSuccess)
        ZK_FAIL(c, ZKB_ERR_CUDA, "msm: radix sort failed");
    k_msm_bounds<<<grid_for(c, m, 8), BLOCK, 0, st>>>(keys2, m, start, end);
    ZK_TRY(check_launch(c, "k_msm_bounds"));
    k_msm_buckets<<<(nb + 127) / 128, 128, 0, st>>>(vals2, start, end, bases, pl, buckets, heavy + 1, heavy);
    ZK_TRY(check_launch(c, "k_msm_buckets"));
    k_msm_heavy<<<c->sm_count, 256, sizeof(G1Jac) * 256, st>>>(vals2, start, end, bases, buckets, heavy + 1, heavy);
    ZK_TRY(check_launch(c, "k_msm_heavy"));
    k_msm_window_chunks<<<(pl.windows * pl.chunks + 127) / 128, 128, 0, st>>>(buckets, pl, parts);
    ZK_TRY(check_launch(c, "k_msm_window_chunks"));
    k_msm_window_sum<<<pl.windows, 128, 0, st>>>(parts, pl, wsum);
    ZK_TRY(check_launch(c, "k_msm_window_sum"));
    k_msm_horner<<<1, 32, 0, st>>>(wsum, pl, d_out);
    ZK_TRY(check_launch(c, "k_msm_horner"));
    for (void* q : {(void*)keys, (void*)vals, (void*)keys2, (void*)vals2, tmp, (void*)start, (void*)heavy, (void*)buckets, (void*)parts, (void*)wsum})
        cudaFreeAsync(q, st);
    return ZKB_OK;
}
int32_t kzg_require_field(zkb_ctx* c) {
    if (c->field != ZKB_FIELD_BLS12_381_FR) ZK_FAIL(c, ZKB_ERR_UNSUPPORTED, "kzg: the commitment lives on BLS12-381; create the ctx with ZKB_FIELD_BLS12_381_FR");
    if (c->comm && c->world > 1) ZK_FAIL(c, ZKB_ERR_UNSUPPORTED, "kzg: sharded tables are not supported (GKR input layers are replicas only)");
    return ZKB_OK;
}
int32_t jac_to_affine(zkb_ctx* c, const G1Jac* jac, G1Affine* aff, uint64_t n) {
    const uint64_t threads = (n + KZG_BATCH - 1) / KZG_BATCH;
    k_g1_batch_affine<<<(unsigned)((threads + 127) / 128), 128, 0, c->stream>>>(jac, aff, n);
    return check_launch(c, "k_g1_batch_affine");
}
}  // namespace

// =================================================================== C ABI
extern "C" {

const char* zkb_strerror(int32_t s) {
    switch (s) {
        case ZKB_OK: return "ok";
        case ZKB_ERR_BAD_ARG: return "bad argument";
        case ZKB_ERR_NOT_POW2: return "Invalid evaluations";
        case ZKB_ERR_ARITY: return "Invalid number of values";
        case ZKB_ERR_LENGTH_MISMATCH: return "all evaluations must have same length";
        case ZKB_ERR_DEGREE_MISMATCH: return "all product polys must have same degree";
        case ZKB_ERR_CUDA: return "CUDA error";
        case ZKB_ERR_NCCL: return "NCCL error";
        case ZKB_ERR_OOM: return "out of device memory";
        case ZKB_ERR_UNSUPPORTED: return "shape outside the instantiated kernels";
        case ZKB_ERR_COMPAT_SHAPE: return "compat mode needs >= 2 products of >= 2 factors";
        case ZKB_ERR_CIRCUIT_SHAPE: return "circuit shape not expressible by the reference wiring";
    }
    return "unknown status";
}
#ifndef ZKB_SRC_HASH
#define ZKB_SRC_HASH "unknown"
#endif
const char* zkb_version(void) { return "zkb200 0.2 (sm_100a) src:" ZKB_SRC_HASH; }

int32_t zkb_ctx_create(int32_t field_id, int32_t device, int32_t mode, zkb_ctx** out) {
    if (!out) return ZKB_ERR_BAD_ARG;
    *out = nullptr;
    const FieldKernels* K = kernels_for(field_id);
    if (!K || (mode != ZKB_MODE_COMPAT && mode != ZKB_MODE_FULL)) return ZKB_ERR_BAD_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return ZKB_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return ZKB_ERR_CUDA;
    std::unique_ptr<zkb_ctx> c(new zkb_ctx);
    c->field = field_id;
    c->device = device;
    c->mode = mode;
    c->K = K;
    c->H = HostField::make(K);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return ZKB_ERR_CUDA;
    c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return ZKB_ERR_CUDA;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    void* hp = nullptr;
    if (cudaHostAlloc(&hp, sizeof(Fe) * 64 + 64, cudaHostAllocMapped) != cudaSuccess) return ZKB_ERR_CUDA;
    std::memset(hp, 0, sizeof(Fe) * 64 + 64);
    c->h_res = (Fe*)hp;
    c->h_flag = (volatile unsigned int*)((uint8_t*)hp + sizeof(Fe) * 64);
    if (cudaHostAlloc((void**)&c->h_wide, sizeof(unsigned long long) * 8 * MAXPTS, cudaHostAllocDefault) != cudaSuccess) return ZKB_ERR_CUDA;
    if (cudaMalloc((void**)&c->d_ticket, sizeof(unsigned int)) != cudaSuccess) return ZKB_ERR_CUDA;
    cudaMemset(c->d_ticket, 0, sizeof(unsigned int));
    if (cudaMalloc((void**)&c->d_res, sizeof(Fe) * 64) != cudaSuccess) return ZKB_ERR_CUDA;
    if (cudaMalloc((void**)&c->d_wide, sizeof(unsigned long long) * 8 * MAXPTS) != cudaSuccess) return ZKB_ERR_CUDA;
    if (cudaHostAlloc((void**)&c->mb, sizeof(TailMailbox), cudaHostAllocMapped) != cudaSuccess) return ZKB_ERR_CUDA;
    std::memset((void*)c->mb, 0, sizeof(TailMailbox));
    if (cudaMalloc((void**)&c->d_relay, sizeof(TailRelay)) != cudaSuccess) return ZKB_ERR_CUDA;
    cudaMemset(c->d_relay, 0, sizeof(TailRelay));
    for (int n = 2; n <= MAXPTS; ++n) c->interp[n].init(c->H, n);
    c->fmb.init(c->H);
    c->tcm.init(c->H);
    if (cudaMalloc((void**)&c->d_cpow8, sizeof(Fe) * 33) != cudaSuccess) return ZKB_ERR_CUDA;
    if (cudaMemcpy(c->d_cpow8, c->tcm.c, sizeof(Fe) * 33, cudaMemcpyHostToDevice) != cudaSuccess) return ZKB_ERR_CUDA;
    c->tc_enabled = getenv("ZKB200_NO_TC") == nullptr;
    if (c->tc_enabled) {  // co-residency probe for the plain-launched persistent variant (tcfold.cuh k_tc_probe)
        unsigned int* d = c->d_ticket;  // zero between launches; [1] does not exist: use d_res as the failure word
        unsigned int* failed = reinterpret_cast<unsigned int*>(c->d_res);
        cudaMemsetAsync(failed, 0, sizeof(unsigned int), c->stream);
        const int e = launch_tc_probe(d, failed, 2 * c->sm_count, TailSmemTc<4>::bytes, 100000000ll, c->stream);  // ~50 ms
        unsigned int f = 1;
        if (e == 0 && cudaMemcpyAsync(&f, failed, sizeof f, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess && cudaStreamSynchronize(c->stream) == cudaSuccess)
            c->tc_tail_ok = f == 0;
        else
            cudaGetLastError();
        cudaMemsetAsync(d, 0, sizeof(unsigned int), c->stream);
    }
    if (const char* e = getenv("ZKB200_GATHER_LOG2")) c->gather_log2 = (uint32_t)std::atoi(e);  // tuning: as zkb_ctx_set_gather_threshold
    {
        const char* e = getenv("ZKB200_CLUSTER_MAX");
        int want = e ? std::atoi(e) : SMALL_MAX_CLUSTER;
        if (want > SMALL_MAX_CLUSTER) want = SMALL_MAX_CLUSTER;
        c->cluster_max = 1;
        if (want > 1) {
            const int have = c->K->sc_small_max_cluster();
            while (c->cluster_max * 2 <= want && c->cluster_max * 2 <= have) c->cluster_max *= 2;
        }
    }
    // Under Nsight Compute every launch is made synchronous, so a kernel that waits for the host's next
    // challenge can never be answered: profile with one launch per round (the same round_pass code).
    extern char** environ;
    bool profiler = getenv("ZKB200_NO_PERSISTENT") != nullptr;
    for (char** e = environ; e && *e && !profiler; ++e)
        if (std::strncmp(*e, "CUDA_INJECTION64_PATH=", 22) == 0 || std::strncmp(*e, "NV_NSIGHT_INJECTION", 19) == 0 ||
            std::strncmp(*e, "NV_COMPUTE_PROFILER", 19) == 0)
            profiler = true;
    if (profiler) {
        c->tail_log2 = 0;
        c->small_bytes = 0;
    }
    c->trace = getenv("ZKB200_TRACE") != nullptr;
    c->trace_verbose = c->trace && std::atoi(getenv("ZKB200_TRACE")) >= 2;
    // device transcript of k_sc_small: per-round records in mapped host memory
    {
        void* hp = nullptr;
        if (cudaHostAlloc(&hp, sizeof(DtRound) * DT_MAX_ROUNDS, cudaHostAllocMapped) != cudaSuccess) {
            cudaGetLastError();
            c->dt_enabled = false;
        } else {
            std::memset(hp, 0, sizeof(DtRound) * DT_MAX_ROUNDS);
            c->dt_rounds = (DtRound*)hp;
        }
        if (getenv("ZKB200_DT")) c->dt_enabled = std::atoi(getenv("ZKB200_DT")) != 0 && c->dt_rounds;
    }
    *out = c.release();
    return ZKB_OK;
}

int32_t zkb_ctx_destroy(zkb_ctx* c) {
    if (!c) return ZKB_ERR_BAD_ARG;
    if (c->d_dbg) {
        std::vector<unsigned long long> h(2 * 1024);
        cudaMemcpy(h.data(), c->d_dbg, h.size() * 8, cudaMemcpyDeviceToHost);
        unsigned long long t0 = ~0ull;
        for (int b = 0; b < c->dbg_grid; ++b) t0 = h[2 * b] < t0 ? h[2 * b] : t0;
        fprintf(stderr, "[zkb200 trace] per-CTA pass of the second round (start, end in us after the first start):\n");
        for (int b = 0; b < c->dbg_grid; ++b)
            fprintf(stderr, "%s%d:%.0f-%.0f", b % 12 ? " " : "\n  ", b, (double)(h[2 * b] - t0) * 1e-3, (double)(h[2 * b + 1] - t0) * 1e-3);
        fprintf(stderr, "\n");
        cudaFree(c->d_dbg);
    }
    if (c->trace && c->tr_n > 0)
        fprintf(stderr, "[zkb200 trace] k_sc_tail rounds=%.0f  per round (us): host %.2f | send->result %.2f = relay %.2f + fan-out %.2f + pass(CTA0) %.2f + reduce/publish %.2f + pcie/poll %.2f\n",
                c->tr_n, c->tr_host / c->tr_n, c->tr_rtt / c->tr_n, c->tr_relay / c->tr_n, c->tr_spread / c->tr_n, c->tr_pass / c->tr_n,
                c->tr_reduce / c->tr_n, (c->tr_rtt - c->tr_relay - c->tr_spread - c->tr_pass - c->tr_reduce) / c->tr_n);
    if (c->trace) {
        for (int lg = 47; lg >= 0; --lg)
            if (c->trs_n[lg] > 0)
                fprintf(stderr, "[zkb200 trace] rank %d round 2^%-2d -> 2^%-2d  x%-4.0f device round trip %8.2f us | wait for the other ranks %7.2f us | host (transcript, claim) %5.2f us\n",
                        c->rank, lg, lg - 1, c->trs_n[lg], c->trs_rtt[lg] / c->trs_n[lg], c->trs_wait[lg] / c->trs_n[lg], c->trs_host[lg] / c->trs_n[lg]);
    }
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& kv : c->sps) sp_release(c, kv.second.get());
    for (auto& kv : c->kzgs) kzg_release(c, kv.second.get());
    if (c->g1_table) cudaFree(c->g1_table);
    for (auto& kv : c->circs) {
        CircuitState* cs = kv.second.get();
        sp_release(c, &cs->sp);
        free_table(c, &cs->inputs);
        for (auto& t : cs->vals) free_table(c, &t);
        free_table(c, &cs->H1);
        free_table(c, &cs->HA2);
        free_table(c, &cs->coef);
        free_table(c, &cs->gtmp);
        for (auto& t : cs->eq) free_table(c, &t);
        if (cs->d_ops) cudaFreeAsync(cs->d_ops, c->stream);
        for (uint32_t* q : {cs->d_in1, cs->d_in2, cs->d_lst1, cs->d_lst2, cs->d_off1, cs->d_off2})
            if (q) cudaFreeAsync(q, c->stream);
        if (cs->aos_stage) cudaFreeAsync(cs->aos_stage, c->stream);
    }
    for (auto& kv : c->mles) free_table(c, &kv.second);
    if (c->stage) cudaFreeAsync(c->stage, c->stream);
    if (c->hash_pin) {
        cudaFreeHost(c->hash_pin);
        for (auto& e : c->hash_ev)
            if (e) cudaEventDestroy(e);
    }
    if (c->d_partials) cudaFreeAsync(c->d_partials, c->stream);
    cudaStreamSynchronize(c->stream);
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        for (int b = 0; b < 2; ++b) {
            cudaFree(c->up_stage[b]);
            cudaEventDestroy(c->ev_copied[b]);
            cudaEventDestroy(c->ev_free[b]);
        }
        cudaEventDestroy(c->ev_entry);
        cudaStreamDestroy(c->copy_stream);
    }
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    c->shm.close_();
    prof_drain(c);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    cudaFree(c->d_ticket);
    cudaFree(c->d_res);
    cudaFree(c->d_wide);
    cudaFreeHost(c->h_res);
    cudaFreeHost(c->h_wide);
    cudaFreeHost((void*)c->mb);
    if (c->dt_rounds) cudaFreeHost(c->dt_rounds);
    cudaFree(c->d_relay);
    cudaFree(c->d_cpow8);
    for (auto& kv : c->merkles) cudaFree(kv.second->tree);
    for (int i = 0; i < 4; ++i) {
        if (c->kzg_streams[i]) cudaStreamDestroy(c->kzg_streams[i]);
        if (c->kzg_ev[i]) cudaEventDestroy(c->kzg_ev[i]);
    }
    cudaStreamDestroy(c->stream);
    delete c;
    return ZKB_OK;
}
const char* zkb_ctx_last_error(const zkb_ctx* c) { return c ? c->last_error.c_str() : "null ctx"; }
void* zkb_ctx_stream(const zkb_ctx* c) { return c ? (void*)c->stream : nullptr; }
uint64_t zkb_ctx_launch_count(const zkb_ctx* c) { return c ? c->launches : 0; }
int32_t zkb_ctx_sync(zkb_ctx* c) {
    if (!c) return ZKB_ERR_BAD_ARG;
    ZK_CUDA(c, cudaStreamSynchronize(c->stream));
    return ZKB_OK;
}

int32_t zkb_ctx_profile(zkb_ctx* c, int32_t enable) {
    if (!c) return ZKB_ERR_BAD_ARG;
    ZK_CUDA(c, cudaStreamSynchronize(c->stream));
    prof_drain(c);
    if (enable && !c->prof) {
        for (int k = 0; k < ZKB_K_COUNT; ++k) {
            c->prof_launches[k] = 0;
            c->prof_ms[k] = c->prof_bytes[k] = 0;
        }
    }
    c->prof = enable != 0;
    return ZKB_OK;
}
int32_t zkb_ctx_profile_read(zkb_ctx* c, int32_t k, uint64_t* launches, double* ms, double* alg_bytes) {
    if (!c || k < 0 || k >= ZKB_K_COUNT) return ZKB_ERR_BAD_ARG;
    ZK_CUDA(c, cudaStreamSynchronize(c->stream));
    prof_drain(c);
    if (launches) *launches = c->prof_launches[k];
    if (ms) *ms = c->prof_ms[k];
    if (alg_bytes) *alg_bytes = c->prof_bytes[k];
    return ZKB_OK;
}
const char* zkb_kernel_name(int32_t k) {
    static const char* names[ZKB_K_COUNT] = {"k_sc_eval", "k_sc_fold_eval", "k_fold_tables", "k_final_bind", "k_fold",
                                             "k_aos_to_planar/k_planar_to_aos", "k_gkr_phase1/2", "other", "k_sc_tail", "k_sc_small",
                                             "k_sc_tail (latency-bound launch)"};
    return (k >= 0 && k < ZKB_K_COUNT) ? names[k] : "?";
}

int32_t zkb_comm_unique_id(uint8_t out[128]) {
    if (!g_nccl.load()) return ZKB_ERR_NCCL;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return ZKB_ERR_NCCL;
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(out, &id, 128);
    return ZKB_OK;
}
int32_t zkb_ctx_comm_init(zkb_ctx* c, int32_t rank, int32_t world, const uint8_t unique_id[128]) {
    if (!c || world < 1 || rank < 0 || rank >= world || !is_pow2((uint64_t)world)) return ZKB_ERR_BAD_ARG;
    if (world == 1) return ZKB_OK;
    if (!g_nccl.load()) ZK_FAIL(c, ZKB_ERR_NCCL, "cannot load libnccl.so.2");
    ncclUniqueId id;
    std::memcpy(&id, unique_id, 128);
    ZK_CUDA(c, cudaSetDevice(c->device));
    ZK_NCCL(c, g_nccl.CommInitRank(&c->comm, world, id, rank));
    c->rank = rank;
    c->world = world;
    c->log2world = ilog2_u64((uint64_t)world);
    c->use_shm = getenv("ZKB200_NO_SHM") == nullptr && c->shm.open(unique_id, rank, world);
    return ZKB_OK;
}
int32_t zkb_ctx_tensor_cores(const zkb_ctx* c, int32_t* enabled, int32_t* persistent) {
    if (!c || !enabled || !persistent) return ZKB_ERR_BAD_ARG;
    *enabled = c->tc_enabled ? 1 : 0;
    *persistent = c->tc_enabled && c->tc_tail_ok ? 1 : 0;
    return ZKB_OK;
}
int32_t zkb_ctx_small_cluster_max(const zkb_ctx* c, int32_t* ctas) {
    if (!c || !ctas) return ZKB_ERR_BAD_ARG;
    *ctas = c->cluster_max;
    return ZKB_OK;
}
int32_t zkb_ctx_set_tail_threshold(zkb_ctx* c, uint32_t log2_entries) {
    if (!c) return ZKB_ERR_BAD_ARG;
    c->tail_log2 = log2_entries;
    return ZKB_OK;
}
int32_t zkb_ctx_set_device_transcript(zkb_ctx* c, int32_t enable) {
    if (!c) return ZKB_ERR_BAD_ARG;
    c->dt_enabled = enable != 0 && c->dt_rounds != nullptr;
    return ZKB_OK;
}
int32_t zkb_ctx_device_transcript_stats(const zkb_ctx* c, uint64_t* launches, uint64_t* rounds_checked) {
    if (!c) return ZKB_ERR_BAD_ARG;
    if (launches) *launches = c->dt_launches;
    if (rounds_checked) *rounds_checked = c->dt_rounds_checked;
    return ZKB_OK;
}
int32_t zkb_ctx_set_small_threshold(zkb_ctx* c, uint32_t smem_bytes) {
    if (!c) return ZKB_ERR_BAD_ARG;
    c->small_bytes = smem_bytes;
    return ZKB_OK;
}
int32_t zkb_ctx_set_gather_threshold(zkb_ctx* c, uint32_t log2_local_entries) {
    if (!c) return ZKB_ERR_BAD_ARG;
    c->gather_log2 = log2_local_entries;  // 0 = automatic
    return ZKB_OK;
}

// --------------------------------------------------------------- MultilinearPoly
int32_t zkb_mle_upload(zkb_ctx* c, const uint64_t* aos, uint64_t len, zkb_mle* out) {
    if (!c || !aos || !out) return ZKB_ERR_BAD_ARG;
    if (!is_pow2(len)) ZK_FAIL(c, ZKB_ERR_NOT_POW2, "Invalid evaluations");
    Table t;
    ZK_TRY(upload_aos(c, aos, len, 0, 1, len, 0, &t));
    *out = put_mle(c, t);
    return ZKB_OK;
}
int32_t zkb_mle_upload_shard(zkb_ctx* c, const uint64_t* aos, uint64_t len_full, zkb_mle* out) {
    if (!c || !aos || !out) return ZKB_ERR_BAD_ARG;
    if (!is_pow2(len_full) || len_full < (uint64_t)c->world * 2) ZK_FAIL(c, ZKB_ERR_NOT_POW2, "Invalid evaluations");
    Table t;
    ZK_TRY(upload_aos(c, aos, len_full, (uint64_t)c->rank, (uint64_t)c->world, len_full / c->world, 0, &t));
    *out = put_mle(c, t);
    return ZKB_OK;
}
int32_t zkb_mle_generate(zkb_ctx* c, uint64_t seed, uint64_t table_id, uint32_t n_vars, zkb_mle* out) {
    if (!c || !out || n_vars > 40) return ZKB_ERR_BAD_ARG;
    if ((int)n_vars < c->log2world + 1 && c->world > 1) return ZKB_ERR_BAD_ARG;
    const uint64_t n_local = (1ull << n_vars) >> c->log2world;
    Table t;
    ZK_TRY(alloc_table(c, n_local, &t));
    c->K->generate(t.ref(), n_local, seed, table_id, (uint64_t)c->rank, (uint64_t)c->world, grid_for(c, n_local, 8), c->stream);
    ZK_TRY(check_launch(c, "k_generate"));
    *out = put_mle(c, t);
    return ZKB_OK;
}
int32_t zkb_mle_download(zkb_ctx* c, zkb_mle m, uint64_t* aos) {
    Table* t = c ? find_mle(c, m) : nullptr;
    if (!t || !aos) return ZKB_ERR_BAD_ARG;
    return download_aos(c, *t, aos, 0);
}
int32_t zkb_mle_download_canonical(zkb_ctx* c, zkb_mle m, uint8_t* bytes) {
    Table* t = c ? find_mle(c, m) : nullptr;
    if (!t || !bytes) return ZKB_ERR_BAD_ARG;
    return download_aos(c, *t, bytes, 2);
}
int32_t zkb_mle_clone(zkb_ctx* c, zkb_mle m, zkb_mle* out) {
    Table* t = c ? find_mle(c, m) : nullptr;
    if (!t || !out) return ZKB_ERR_BAD_ARG;
    Table w;
    ZK_TRY(multi_fold(c, *t, nullptr, 0, &w));
    *out = put_mle(c, w);
    return ZKB_OK;
}
int32_t zkb_mle_free(zkb_ctx* c, zkb_mle m) {
    Table* t = c ? find_mle(c, m) : nullptr;
    if (!t) return ZKB_ERR_BAD_ARG;
    if (t->refs > 0) {  // a SumPoly still reads it (the reference's SumPoly owns clones): free when the last one goes
        t->dead = true;
        return ZKB_OK;
    }
    free_table(c, t);
    c->mles.erase(m);
    return ZKB_OK;
}
int32_t zkb_mle_num_vars(zkb_ctx* c, zkb_mle m, uint32_t* n_vars) {
    Table* t = c ? find_mle(c, m) : nullptr;
    if (!t || !n_vars) return ZKB_ERR_BAD_ARG;
    *n_vars = (uint32_t)ilog2_u64(t->n);
    return ZKB_OK;
}
int32_t zkb_mle_partial_evaluate(zkb_ctx* c, zkb_mle in, uint32_t bit, const uint64_t value[4], zkb_mle* out) {
    Table* t = c ? find_mle(c, in) : nullptr;
    if (!t || !value || !out) return ZKB_ERR_BAD_ARG;
    const int nv = ilog2_u64(t->n);
    if ((int)bit >= nv) ZK_FAIL(c, ZKB_ERR_ARITY, "partial_evaluate: bit out of range");
    Table w;
    ZK_TRY(alloc_table(c, t->n / 2, &w));
    FixedMul rt;
    c->fmb.make(c->H, fe_from_u64x4(value), &rt);
    prof_begin(c, ZKB_K_FOLD, 96.0 * (double)(t->n / 2));
    c->K->fold(t->ref(), w.ref(), t->n / 2, (uint32_t)(nv - 1 - (int)bit), rt, grid_for(c, t->n / 2, 8), c->stream);
    ZK_TRY(check_launch(c, "k_fold"));
    *out = put_mle(c, w);
    return ZKB_OK;
}
int32_t zkb_mle_multi_partial_evaluate(zkb_ctx* c, zkb_mle in, const uint64_t* values, uint32_t k, zkb_mle* out) {
    Table* t = c ? find_mle(c, in) : nullptr;
    if (!t || (!values && k) || !out) return ZKB_ERR_BAD_ARG;
    std::vector<Fe> rs(k);
    for (uint32_t i = 0; i < k; ++i) rs[i] = fe_from_u64x4(values + 4 * i);
    Table w;
    ZK_TRY(multi_fold(c, *t, rs.data(), k, &w));
    *out = put_mle(c, w);
    return ZKB_OK;
}
int32_t zkb_mle_evaluate(zkb_ctx* c, zkb_mle m, const uint64_t* values, uint32_t k, uint64_t out[4]) {
    Table* t = c ? find_mle(c, m) : nullptr;
    if (!t || (!values && k) || !out) return ZKB_ERR_BAD_ARG;
    std::vector<Fe> rs(k);
    for (uint32_t i = 0; i < k; ++i) rs[i] = fe_from_u64x4(values + 4 * i);
    Fe v;
    ZK_TRY(evaluate_table(c, *t, rs.data(), k, &v));
    fe_to_u64x4(v, out);
    return ZKB_OK;
}
int32_t zkb_mle_sum_halves(zkb_ctx* c, zkb_mle m, uint64_t out[8]) {
    Table* t = c ? find_mle(c, m) : nullptr;
    if (!t || !out) return ZKB_ERR_BAD_ARG;
    if (t->n < 2) ZK_FAIL(c, ZKB_ERR_ARITY, "sum_halves: table has no variable");
    SumPolyState sp;
    ZK_TRY(sp_configure(c, &sp, 1, 1, KIND_PROD));
    sp.src.push_back(*t);
    sp.n0 = t->n;
    sp_reset(c, &sp);
    Fe ev[2];
    ZK_TRY(sp_round_evals(c, &sp, ev));
    fe_to_u64x4(ev[0], out);
    fe_to_u64x4(ev[1], out + 4);
    return ZKB_OK;
}
int32_t zkb_mle_scale(zkb_ctx* c, zkb_mle in, const uint64_t value[4], zkb_mle* out) {
    Table* t = c ? find_mle(c, in) : nullptr;
    if (!t || !value || !out) return ZKB_ERR_BAD_ARG;
    Table w;
    ZK_TRY(alloc_table(c, t->n, &w));
    TabRef none{nullptr, 0};
    c->K->axpby(t->ref(), none, w.ref(), t->n, fe_from_u64x4(value), c->H.zero(), grid_for(c, t->n, 8), c->stream);
    ZK_TRY(check_launch(c, "k_axpby"));
    *out = put_mle(c, w);
    return ZKB_OK;
}
int32_t zkb_mle_binary(zkb_ctx* c, zkb_mle a, zkb_mle b, int32_t op, zkb_mle* out) {
    Table* x = c ? find_mle(c, a) : nullptr;
    Table* y = c ? find_mle(c, b) : nullptr;
    if (!x || !y || !out || op < ZKB_OP_ADD || op > ZKB_OP_SUB) return ZKB_ERR_BAD_ARG;
    if (x->n != y->n) ZK_FAIL(c, ZKB_ERR_LENGTH_MISMATCH, "all evaluations must have same length");
    Table w;
    ZK_TRY(alloc_table(c, x->n, &w));
    const int kop = op == ZKB_OP_ADD ? 0 : (op == ZKB_OP_SUB ? 1 : 2);
    c->K->vec_op(x->ref(), y->ref(), w.ref(), x->n, kop, grid_for(c, x->n, 8), c->stream);
    ZK_TRY(check_launch(c, "k_vec_op"));
    *out = put_mle(c, w);
    return ZKB_OK;
}
int32_t zkb_mle_tensor(zkb_ctx* c, zkb_mle a, zkb_mle b, int32_t op, zkb_mle* out) {
    Table* x = c ? find_mle(c, a) : nullptr;
    Table* y = c ? find_mle(c, b) : nullptr;
    if (!x || !y || !out || (op != ZKB_OP_ADD && op != ZKB_OP_MUL)) return ZKB_ERR_BAD_ARG;
    Table w;
    ZK_TRY(alloc_table(c, x->n * y->n, &w));
    c->K->tensor(x->ref(), y->ref(), w.ref(), x->n, y->n, op == ZKB_OP_ADD ? 0 : 1, grid_for(c, x->n * y->n, 8), c->stream);
    ZK_TRY(check_launch(c, "k_tensor"));
    *out = put_mle(c, w);
    return ZKB_OK;
}

// ------------------------------------------------------------ ProductPoly / SumPoly
int32_t zkb_sumpoly_create(zkb_ctx* c, const zkb_mle* tables, uint32_t n_products, uint32_t degree, zkb_sp* out) {
    if (!c || !tables || !out || n_products < 1 || degree < 1) return ZKB_ERR_BAD_ARG;
    if (c->mode == ZKB_MODE_COMPAT && (n_products < 2 || degree < 2))
        ZK_FAIL(c, ZKB_ERR_COMPAT_SHAPE, "compat mode needs >= 2 products of >= 2 factors");
    std::unique_ptr<SumPolyState> sp(new SumPolyState);
    ZK_TRY(sp_configure(c, sp.get(), (int)n_products, (int)degree, KIND_PROD));
    for (uint32_t i = 0; i < n_products * degree; ++i) {
        Table* t = find_mle(c, tables[i]);
        if (!t) return ZKB_ERR_BAD_ARG;
        if (i && t->n != sp->src[0].n) ZK_FAIL(c, ZKB_ERR_LENGTH_MISMATCH, "all evaluations must have same length");
        sp->src.push_back(*t);
    }
    for (uint32_t i = 0; i < n_products * degree; ++i) {
        sp->src_handles.push_back(tables[i]);
        ++c->mles[tables[i]].refs;
    }
    sp->n0 = sp->src[0].n;
    sp_reset(c, sp.get());
    zkb_sp h = c->next_handle++;
    c->sps[h] = std::move(sp);
    *out = h;
    return ZKB_OK;
}
int32_t zkb_sumpoly_free(zkb_ctx* c, zkb_sp h) {
    if (!c) return ZKB_ERR_BAD_ARG;
    auto it = c->sps.find(h);
    if (it == c->sps.end()) return ZKB_ERR_BAD_ARG;
    sp_release(c, it->second.get());
    for (uint64_t mh : it->second->src_handles) {  // drop the references; tables freed meanwhile go now
        auto mt = c->mles.find(mh);
        if (mt == c->mles.end()) continue;
        if (--mt->second.refs <= 0 && mt->second.dead) {
            free_table(c, &mt->second);
            c->mles.erase(mt);
        }
    }
    c->sps.erase(it);
    return ZKB_OK;
}
int32_t zkb_sumpoly_reset(zkb_ctx* c, zkb_sp h) {
    if (!c) return ZKB_ERR_BAD_ARG;
    auto it = c->sps.find(h);
    if (it == c->sps.end()) return ZKB_ERR_BAD_ARG;
    sp_reset(c, it->second.get());
    return ZKB_OK;
}
int32_t zkb_sumpoly_evaluate(zkb_ctx* c, zkb_sp h, const uint64_t* values, uint32_t k, uint64_t out[4]) {
    if (!c || !out) return ZKB_ERR_BAD_ARG;
    auto it = c->sps.find(h);
    if (it == c->sps.end()) return ZKB_ERR_BAD_ARG;
    SumPolyState* sp = it->second.get();
    std::vector<Fe> rs(k);
    for (uint32_t i = 0; i < k; ++i) rs[i] = fe_from_u64x4(values + 4 * i);
    Fe sum = c->H.zero();
    for (int p = 0; p < sp->P; ++p) {
        Fe prod = c->H.one();
        for (int f = 0; f < sp->D; ++f) {
            Fe v;
            ZK_TRY(evaluate_table(c, sp->src[p * sp->D + f], rs.data(), k, &v));
            prod = c->H.mul(prod, v);
        }
        sum = c->H.add(sum, prod);
    }
    fe_to_u64x4(sum, out);
    return ZKB_OK;
}
int32_t zkb_sc_round_evals(zkb_ctx* c, zkb_sp h, uint64_t* evals) {
    if (!c || !evals) return ZKB_ERR_BAD_ARG;
    auto it = c->sps.find(h);
    if (it == c->sps.end()) return ZKB_ERR_BAD_ARG;
    Fe ev[MAXPTS];
    ZK_TRY(sp_round_evals(c, it->second.get(), ev));
    for (int i = 0; i < it->second->npts; ++i) fe_to_u64x4(ev[i], evals + 4 * i);
    return ZKB_OK;
}
int32_t zkb_sc_bind_and_next(zkb_ctx* c, zkb_sp h, const uint64_t r[4], uint64_t* evals) {
    if (!c || !r) return ZKB_ERR_BAD_ARG;
    auto it = c->sps.find(h);
    if (it == c->sps.end()) return ZKB_ERR_BAD_ARG;
    SumPolyState* sp = it->second.get();
    Fe ev[MAXPTS];
    const bool last = sp->cur_n == 2 && !sp->sharded;
    if (last && evals) evals = nullptr;
    ZK_TRY(sp_bind_and_next(c, sp, fe_from_u64x4(r), evals ? ev : nullptr, nullptr));
    if (evals)
        for (int i = 0; i < sp->npts; ++i) fe_to_u64x4(ev[i], evals + 4 * i);
    return ZKB_OK;
}
int32_t zkb_sc_final_values(zkb_ctx* c, zkb_sp h, uint64_t* values) {
    if (!c || !values) return ZKB_ERR_BAD_ARG;
    auto it = c->sps.find(h);
    if (it == c->sps.end()) return ZKB_ERR_BAD_ARG;
    SumPolyState* sp = it->second.get();
    std::vector<Fe> v(sp->src.size());
    ZK_TRY(sp_final_values(c, sp, v.data()));
    for (size_t t = 0; t < v.size(); ++t) fe_to_u64x4(v[t], values + 4 * t);
    return ZKB_OK;
}

// ------------------------------------------------------------------ Transcript
int32_t zkb_transcript_new(int32_t field_id, zkb_transcript** out) {
    const FieldKernels* K = kernels_for(field_id);
    if (!K || !out) return ZKB_ERR_BAD_ARG;
    zkb_transcript* t = new zkb_transcript;
    t->impl.H = HostField::make(K);
    *out = t;
    return ZKB_OK;
}
int32_t zkb_transcript_free(zkb_transcript* t) {
    delete t;
    return ZKB_OK;
}
int32_t zkb_transcript_append(zkb_transcript* t, const uint8_t* bytes, size_t len) {
    if (!t || (!bytes && len)) return ZKB_ERR_BAD_ARG;
    t->impl.append(bytes, len);
    return ZKB_OK;
}
int32_t zkb_transcript_append_elements(zkb_transcript* t, const uint64_t* mont, size_t n) {
    if (!t || (!mont && n)) return ZKB_ERR_BAD_ARG;
    for (size_t i = 0; i < n; ++i) {
        Fe v = fe_from_u64x4(mont + 4 * i);
        t->impl.append_elements(&v, 1);
    }
    return ZKB_OK;
}
int32_t zkb_transcript_challenge(zkb_transcript* t, uint64_t out_mont[4]) {
    if (!t || !out_mont) return ZKB_ERR_BAD_ARG;
    fe_to_u64x4(t->impl.challenge(), out_mont);
    return ZKB_OK;
}
int32_t zkb_keccak256(const uint8_t* bytes, size_t len, uint8_t out[32]) {
    if ((!bytes && len) || !out) return ZKB_ERR_BAD_ARG;
    Keccak256 k;
    k.update(bytes, len);
    k.finalize_reset(out);
    return ZKB_OK;
}

int32_t zkb_fe_to_mont(int32_t field_id, const uint64_t* canonical, uint64_t* mont, size_t n) {
    const FieldKernels* K = kernels_for(field_id);
    if (!K || (!canonical && n) || (!mont && n)) return ZKB_ERR_BAD_ARG;
    HostField H = HostField::make(K);
    for (size_t i = 0; i < n; ++i) fe_to_u64x4(H.to_mont(fe_from_u64x4(canonical + 4 * i)), mont + 4 * i);
    return ZKB_OK;
}
int32_t zkb_fe_from_mont(int32_t field_id, const uint64_t* mont, uint64_t* canonical, size_t n) {
    const FieldKernels* K = kernels_for(field_id);
    if (!K || (!canonical && n) || (!mont && n)) return ZKB_ERR_BAD_ARG;
    HostField H = HostField::make(K);
    for (size_t i = 0; i < n; ++i) fe_to_u64x4(H.from_mont(fe_from_u64x4(mont + 4 * i)), canonical + 4 * i);
    return ZKB_OK;
}

int32_t zkb_tc_fold_matrices(int32_t field_id, const uint64_t r_mont[4], uint8_t out[2048]) {
    if (!r_mont || !out || field_id < 0 || field_id > 2) return ZKB_ERR_BAD_ARG;
    const FieldKernels* K = field_id == 0 ? field_kernels_bn254_fr() : (field_id == 1 ? field_kernels_bn254_fq() : field_kernels_bls12_381_fr());
    const HostField H = HostField::make(K);
    TcMatsBuilder b;
    b.init(H);
    TcFoldMats m;
    b.make(H, fe_from_u64x4(r_mont), &m);
    std::memcpy(out, m.b, 2048);
    return ZKB_OK;
}
int32_t zkb_fe_reduce_wide(int32_t field_id, const uint64_t* wide, uint64_t* out, size_t n) {
    const FieldKernels* K = kernels_for(field_id);
    if (!K || (!wide && n) || (!out && n)) return ZKB_ERR_BAD_ARG;
    HostField H = HostField::make(K);
    for (size_t i = 0; i < n; ++i) fe_to_u64x4(H.from_wide_limbs((const unsigned long long*)wide + 8 * i), out + 4 * i);
    return ZKB_OK;
}

// -------------------------------------------------------------- UnivariatePoly
int32_t zkb_uni_interpolate(int32_t field_id, const uint64_t* xs, const uint64_t* ys, uint32_t n, uint64_t* coeffs, uint32_t* len) {
    const FieldKernels* K = kernels_for(field_id);
    if (!K || !xs || !ys || !coeffs || !len || n > 64) return ZKB_ERR_BAD_ARG;
    HostField H = HostField::make(K);
    std::vector<Fe> x(n), y(n), co(n);
    for (uint32_t i = 0; i < n; ++i) {
        x[i] = fe_from_u64x4(xs + 4 * i);
        y[i] = fe_from_u64x4(ys + 4 * i);
    }
    int l = uni_interpolate(H, x.data(), y.data(), (int)n, co.data());
    for (int i = 0; i < l; ++i) fe_to_u64x4(co[i], coeffs + 4 * i);
    *len = (uint32_t)l;
    return ZKB_OK;
}
int32_t zkb_uni_evaluate(int32_t field_id, const uint64_t* coeffs, uint32_t len, const uint64_t x[4], uint64_t out[4]) {
    const FieldKernels* K = kernels_for(field_id);
    if (!K || (!coeffs && len) || !x || !out) return ZKB_ERR_BAD_ARG;
    HostField H = HostField::make(K);
    std::vector<Fe> co(len);
    for (uint32_t i = 0; i < len; ++i) co[i] = fe_from_u64x4(coeffs + 4 * i);
    fe_to_u64x4(uni_evaluate(H, co.data(), (int)len, fe_from_u64x4(x)), out);
    return ZKB_OK;
}

// ------------------------------------------------------------ proof wire format (host only)
// A stable byte encoding of the sumcheck proofs: every field element as fq_vec_to_bytes writes it into the transcript
// (32-byte little-endian canonical integer, fiat_shamir_transcript.rs:32-37), behind a 12-byte header.
//   "ZKBP" | u8 version = 1 | u8 field_id | u8 kind (1 = Proof of sum_check::prove, 2 = GkrProof of gkr_prove) | u8 0 |
//   u32 n_rounds (LE) | claimed_sum (32 B) | per round: u8 len, then len x 32 B  (kind 1: len = 2, the evaluations
//   [s(0), s(1)]; kind 2: the trimmed ascending coefficients, len <= slots)
int32_t zkb_proof_encode(int32_t field_id, int32_t kind, uint32_t n_rounds, uint32_t slots, const uint64_t* msgs, const int32_t* lens,
                         const uint64_t claimed_sum[4], uint8_t* out, size_t cap, size_t* len) {
    const FieldKernels* K = kernels_for(field_id);
    if (!K || (kind != 1 && kind != 2) || (!msgs && n_rounds) || !claimed_sum || !len || slots > 255 || (kind == 2 && !lens && n_rounds)) return ZKB_ERR_BAD_ARG;
    if (kind == 1 && slots != 2) return ZKB_ERR_BAD_ARG;
    HostField H = HostField::make(K);
    size_t need = 12 + 32;
    for (uint32_t k = 0; k < n_rounds; ++k) {
        const int32_t l = kind == 1 ? 2 : lens[k];
        if (l < 0 || (uint32_t)l > slots) return ZKB_ERR_BAD_ARG;
        need += 1 + 32 * (size_t)l;
    }
    *len = need;
    if (!out) return ZKB_OK;  // size query
    if (cap < need) return ZKB_ERR_BAD_ARG;
    if (!all_canonical(H, claimed_sum, 1)) return ZKB_ERR_BAD_ARG;
    uint8_t* w = out;
    std::memcpy(w, "ZKBP", 4);
    w[4] = 1;
    w[5] = (uint8_t)field_id;
    w[6] = (uint8_t)kind;
    w[7] = 0;
    for (int i = 0; i < 4; ++i) w[8 + i] = (uint8_t)(n_rounds >> (8 * i));
    w += 12;
    auto put = [&](const uint64_t* mont) {
        const Fe v = H.from_mont(fe_from_u64x4(mont));
        std::memcpy(w, v.l, 32);
        w += 32;
    };
    put(claimed_sum);
    for (uint32_t k = 0; k < n_rounds; ++k) {
        const int32_t l = kind == 1 ? 2 : lens[k];
        if (!all_canonical(H, msgs + (size_t)k * slots * 4, (size_t)l)) return ZKB_ERR_BAD_ARG;
        *w++ = (uint8_t)l;
        for (int32_t i = 0; i < l; ++i) put(msgs + ((size_t)k * slots + i) * 4);
    }
    return ZKB_OK;
}
int32_t zkb_proof_decode(const uint8_t* bytes, size_t len, int32_t* field_id, int32_t* kind, uint32_t* n_rounds, uint32_t slots,
                         uint64_t* msgs, int32_t* lens, uint64_t claimed_sum[4]) {
    if (!bytes || len < 44 || !field_id || !kind || !n_rounds) return ZKB_ERR_BAD_ARG;
    if (std::memcmp(bytes, "ZKBP", 4) != 0 || bytes[4] != 1 || bytes[7] != 0) return ZKB_ERR_BAD_ARG;
    const FieldKernels* K = kernels_for(bytes[5]);
    if (!K || (bytes[6] != 1 && bytes[6] != 2)) return ZKB_ERR_BAD_ARG;
    HostField H = HostField::make(K);
    *field_id = bytes[5];
    *kind = bytes[6];
    uint32_t n = 0;
    for (int i = 0; i < 4; ++i) n |= (uint32_t)bytes[8 + i] << (8 * i);
    // first pass: structure and canonical values (every element < p), nothing is written on a malformed proof
    size_t off = 44;
    auto canon_ok = [&](const uint8_t* p) {
        uint64_t v[4];
        std::memcpy(v, p, 32);
        for (int k = 3; k >= 0; --k)
            if (v[k] != H.p[k]) return v[k] < H.p[k];
        return false;
    };
    if (!canon_ok(bytes + 12)) return ZKB_ERR_BAD_ARG;
    for (uint32_t k = 0; k < n; ++k) {
        if (off >= len) return ZKB_ERR_BAD_ARG;
        const uint32_t l = bytes[off++];
        if ((bytes[6] == 1 && l != 2) || off + 32 * (size_t)l > len) return ZKB_ERR_BAD_ARG;
        for (uint32_t i = 0; i < l; ++i)
            if (!canon_ok(bytes + off + 32 * (size_t)i)) return ZKB_ERR_BAD_ARG;
        if (msgs && l > slots) return ZKB_ERR_BAD_ARG;
        off += 32 * (size_t)l;
    }
    if (off != len) return ZKB_ERR_BAD_ARG;
    *n_rounds = n;
    if (!msgs) return ZKB_OK;  // header query (the caller sizes its buffers from n_rounds and its slot count)
    if (!claimed_sum || !lens) return ZKB_ERR_BAD_ARG;
    auto get = [&](const uint8_t* p, uint64_t* mont) {
        Fe v;
        std::memcpy(v.l, p, 32);
        fe_to_u64x4(H.to_mont(v), mont);
    };
    get(bytes + 12, claimed_sum);
    off = 44;
    for (uint32_t k = 0; k < n; ++k) {
        const uint32_t l = bytes[off++];
        lens[k] = (int32_t)l;
        for (uint32_t i = 0; i < slots; ++i) {
            if (i < l) get(bytes + off + 32 * (size_t)i, msgs + ((size_t)k * slots + i) * 4);
            else std::memset(msgs + ((size_t)k * slots + i) * 4, 0, 32);
        }
        off += 32 * (size_t)l;
    }
    return ZKB_OK;
}

// ---------------------------------------------------------- sum_check_protocol
static int32_t absorb_table(zkb_ctx* c, const Table& t, TranscriptImpl* tr) {
    // fq_vec_to_bytes(&polynomial.evaluation) (sum_check_protocol.rs:27): canonical bytes made on the
    // device, hashed on the host.  Sharded tables are not supported here (the reference order needs the whole table).
    if (c->comm && c->world > 1) ZK_FAIL(c, ZKB_ERR_UNSUPPORTED, "absorbing a sharded table into the transcript is not supported");
    return absorb_table_bytes(c, t, tr);
}

int32_t zkb_sumcheck_prove(zkb_ctx* c, zkb_mle poly, uint32_t flags, uint64_t claimed_sum[4], uint64_t* msgs, uint64_t* challenges) {
    Table* t = c ? find_mle(c, poly) : nullptr;
    if (!t || !claimed_sum) return ZKB_ERR_BAD_ARG;
    TranscriptImpl tr;
    tr.H = c->H;
    if (flags & 1u) ZK_TRY(absorb_table(c, *t, &tr));
    SumPolyState sp;
    ZK_TRY(sp_configure(c, &sp, 1, 1, KIND_PROD));
    sp.src.push_back(*t);
    sp.n0 = t->n;
    sp_reset(c, &sp);
    const int n_rounds = ilog2_u64(sp.n0) + (sp.sharded ? c->log2world : 0);
    int32_t st = ZKB_OK;
    Fe ev[2], fin[1];
    if (n_rounds == 0) {
        Fe v;
        ZK_TRY(read_elems(c, *t, 0, 1, &v));
        fe_to_u64x4(v, claimed_sum);
        return ZKB_OK;
    }
    if (!msgs) return ZKB_ERR_BAD_ARG;
    RoundDriver drv(c, &sp);
    st = drv.first(ev);
    if (st == ZKB_OK) {
        Fe claimed = c->H.add(ev[0], ev[1]);  // :29 the claimed sum is the sum of the two halves
        fe_to_u64x4(claimed, claimed_sum);
        tr.append_elements(&claimed, 1);
        for (int k = 0; k < n_rounds && st == ZKB_OK; ++k) {
            tr.append_elements(ev, 2);
            fe_to_u64x4(ev[0], msgs + (size_t)k * 8);
            fe_to_u64x4(ev[1], msgs + (size_t)k * 8 + 4);
            Fe r = tr.challenge();
            if (challenges) fe_to_u64x4(r, challenges + (size_t)k * 4);
            st = drv.next(r, k + 1 < n_rounds ? ev : nullptr, fin);
        }
    }
    drv.abort();
    sp_release(c, &sp);
    return st;
}

int32_t zkb_sumcheck_verify(zkb_ctx* c, zkb_mle poly, uint32_t flags, const uint64_t claimed_sum[4], const uint64_t* msgs,
                            uint32_t n_msgs, int32_t* accepted) {
    Table* t = c ? find_mle(c, poly) : nullptr;
    if (!t || !claimed_sum || !accepted || (!msgs && n_msgs)) return ZKB_ERR_BAD_ARG;
    *accepted = 0;
    if (!all_canonical(c->H, claimed_sum, 1) || !all_canonical(c->H, msgs, 2 * (size_t)n_msgs)) return ZKB_OK;  // rejected
    TranscriptImpl tr;
    tr.H = c->H;
    if (flags & 1u) ZK_TRY(absorb_table(c, *t, &tr));
    Fe expected = fe_from_u64x4(claimed_sum);
    tr.append_elements(&expected, 1);
    std::vector<Fe> chals;
    for (uint32_t k = 0; k < n_msgs; ++k) {
        Fe s[2] = {fe_from_u64x4(msgs + (size_t)k * 8), fe_from_u64x4(msgs + (size_t)k * 8 + 4)};
        if (!c->H.eq(c->H.add(s[0], s[1]), expected)) return ZKB_OK;  // :66-68
        tr.append_elements(s, 2);
        Fe r = tr.challenge();
        chals.push_back(r);
        expected = c->H.add(s[0], c->H.mul(r, c->H.sub(s[1], s[0])));  // :73-74
    }
    // :81 polynomial.evaluate(challenges) -- panics (arity) if the proof has the wrong number of rounds
    Fe v;
    ZK_TRY(evaluate_table(c, *t, chals.data(), (uint32_t)chals.size(), &v));
    *accepted = c->H.eq(v, expected) ? 1 : 0;
    return ZKB_OK;
}

int32_t zkb_gkr_sumcheck_prove(zkb_ctx* c, zkb_transcript* t, const uint64_t claimed_sum[4], zkb_sp h, uint64_t* coeffs,
                               int32_t* lens, uint64_t* challenges, uint64_t* final_values) {
    (void)claimed_sum;  // the reference echoes it without using it (sum_check_protocol.rs:87,112)
    if (!c || !t || !coeffs || !lens) return ZKB_ERR_BAD_ARG;
    auto it = c->sps.find(h);
    if (it == c->sps.end()) return ZKB_ERR_BAD_ARG;
    SumPolyState* sp = it->second.get();
    std::vector<Fe> fin(sp->src.size());
    ZK_TRY(sp_prove(c, sp, &t->impl, sp->npts, coeffs, lens, challenges, fin.data(), false));
    if (final_values)
        for (size_t i = 0; i < fin.size(); ++i) fe_to_u64x4(fin[i], final_values + 4 * i);
    return ZKB_OK;
}

int32_t zkb_gkr_sumcheck_verify(zkb_transcript* t, uint32_t n_rounds, uint32_t slots, const uint64_t* coeffs, const int32_t* lens,
                                const uint64_t claimed_sum[4], int32_t* accepted, uint64_t final_claim[4], uint64_t* challenges) {
    if (!t || (!coeffs && n_rounds) || (!lens && n_rounds) || !claimed_sum || !accepted || !final_claim || !challenges) return ZKB_ERR_BAD_ARG;
    for (uint32_t k = 0; k < n_rounds; ++k)
        if (lens[k] < 0 || (uint32_t)lens[k] > slots || lens[k] > MAXPTS + 3) return ZKB_ERR_BAD_ARG;
    const HostField& H = t->impl.H;
    std::vector<Fe> ch(n_rounds ? n_rounds : 1);
    Fe fin;
    if (!all_canonical(H, claimed_sum, 1) || !coeffs_canonical(H, coeffs, lens, n_rounds, slots) ||
        !sc_verify_rounds(H, &t->impl, n_rounds, slots, coeffs, lens, fe_from_u64x4(claimed_sum), &fin, ch.data())) {
        *accepted = 0;  // :129-133
        std::memset(final_claim, 0, 32);
        std::memset(challenges, 0, 32);
        return ZKB_OK;
    }
    *accepted = 1;
    fe_to_u64x4(fin, final_claim);
    for (uint32_t k = 0; k < n_rounds; ++k) fe_to_u64x4(ch[k], challenges + 4 * (size_t)k);
    return ZKB_OK;
}

// ------------------------------------------------------------ gkr_circuit / gkr
uint32_t zkb_gkr_total_rounds(uint32_t n_layers, const uint32_t* gates) {
    uint32_t tot = 0;
    for (uint32_t l = 0; l < n_layers; ++l) tot += layer_rounds(gates[l]);
    return tot;
}

int32_t zkb_circuit_create(zkb_ctx* c, uint32_t n_layers, const uint32_t* gates, const uint8_t* ops, zkb_circ* out) {
    if (!c || !gates || !ops || !out) return ZKB_ERR_BAD_ARG;
    ZK_TRY(circuit_check_shape(c, n_layers, gates));
    std::unique_ptr<CircuitState> cs(new CircuitState);
    cs->L = (int)n_layers;
    cs->gates.assign(gates, gates + n_layers);
    size_t tot = 0;
    for (uint32_t l = 0; l < n_layers; ++l) {
        cs->opoff.push_back(tot);
        tot += gates[l];
    }
    cs->h_ops.assign(ops, ops + tot);
    for (size_t i = 0; i < tot; ++i)
        if (ops[i] != ZKB_OP_ADD && ops[i] != ZKB_OP_MUL) return ZKB_ERR_BAD_ARG;
    ZK_CUDA(c, cudaMallocAsync((void**)&cs->d_ops, tot, c->stream));
    ZK_CUDA(c, cudaMemcpyAsync(cs->d_ops, cs->h_ops.data(), tot, cudaMemcpyHostToDevice, c->stream));
    ZK_CUDA(c, cudaStreamSynchronize(c->stream));
    const uint64_t n_in = 2ull * gates[0];
    ZK_TRY(alloc_table(c, n_in, &cs->inputs));
    cs->vals.resize(n_layers);
    for (uint32_t l = 0; l < n_layers; ++l) ZK_TRY(alloc_table(c, gates[l], &cs->vals[l]));
    ZK_TRY(alloc_table(c, n_in, &cs->H1));
    ZK_TRY(alloc_table(c, n_in, &cs->HA2));
    ZK_TRY(alloc_table(c, gates[0], &cs->coef));
    const int nmax = ilog2_u64(n_in);
    for (auto& t : cs->eq) ZK_TRY(alloc_table(c, 1ull << ((nmax + 1) / 2 + 1), &t));
    ZK_TRY(sp_configure(c, &cs->sp, 1, 3, KIND_XYZ));
    cs->sp.work.resize(3);
    for (auto& t : cs->sp.work) ZK_TRY(alloc_table(c, n_in / 2, &t));
    zkb_circ h = c->next_handle++;
    c->circs[h] = std::move(cs);
    *out = h;
    return ZKB_OK;
}
int32_t zkb_circuit_free(zkb_ctx* c, zkb_circ h) {
    if (!c) return ZKB_ERR_BAD_ARG;
    auto it = c->circs.find(h);
    if (it == c->circs.end()) return ZKB_ERR_BAD_ARG;
    CircuitState* cs = it->second.get();
    sp_release(c, &cs->sp);
    free_table(c, &cs->inputs);
    for (auto& t : cs->vals) free_table(c, &t);
    free_table(c, &cs->H1);
    free_table(c, &cs->HA2);
    free_table(c, &cs->coef);
    free_table(c, &cs->gtmp);
    for (auto& t : cs->eq) free_table(c, &t);
    if (cs->d_ops) cudaFreeAsync(cs->d_ops, c->stream);
    for (uint32_t* q : {cs->d_in1, cs->d_in2, cs->d_lst1, cs->d_lst2, cs->d_off1, cs->d_off2})
        if (q) cudaFreeAsync(q, c->stream);
    if (cs->aos_stage) cudaFreeAsync(cs->aos_stage, c->stream);
    c->circs.erase(it);
    return ZKB_OK;
}
int32_t zkb_circuit_evaluate(zkb_ctx* c, zkb_circ h, const uint64_t* inputs, uint64_t n_inputs, uint64_t* outputs) {
    if (!c || !inputs) return ZKB_ERR_BAD_ARG;
    auto it = c->circs.find(h);
    if (it == c->circs.end()) return ZKB_ERR_BAD_ARG;
    CircuitState* cs = it->second.get();
    ZK_TRY(circuit_run(c, cs, inputs, n_inputs));
    if (outputs) {
        size_t off = 0;
        for (int l = 0; l < cs->L; ++l) {
            ZK_TRY(download_aos(c, cs->vals[l], outputs + off * 4, 0));
            off += cs->gates[l];
        }
    } else {
        ZK_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return ZKB_OK;
}
int32_t zkb_layer_add_mul_i(zkb_ctx* c, const uint8_t* ops, uint32_t n_gates, int32_t op, zkb_mle* out) {
    if (!c || !ops || !out || (op != ZKB_OP_ADD && op != ZKB_OP_MUL)) return ZKB_ERR_BAD_ARG;
    if (!is_pow2(n_gates)) ZK_FAIL(c, ZKB_ERR_CIRCUIT_SHAPE, "get_add_mul_i: the gate count must be a power of two");
    const int w = ilog2_u64(n_gates);
    const int bits = n_gates == 1 ? 3 : 3 * w + 2;
    if (bits > 30) ZK_FAIL(c, ZKB_ERR_UNSUPPORTED, "get_add_mul_i: dense table too large; zkb_gkr_prove uses the sparse form");
    Table t;
    ZK_TRY(alloc_table(c, 1ull << bits, &t));
    ZK_CUDA(c, cudaMemsetAsync(t.base, 0, (size_t)t.n * 32, c->stream));
    uint8_t* d_ops = nullptr;
    ZK_CUDA(c, cudaMallocAsync((void**)&d_ops, n_gates, c->stream));
    ZK_CUDA(c, cudaMemcpyAsync(d_ops, ops, n_gates, cudaMemcpyHostToDevice, c->stream));
    c->K->add_mul_i(d_ops, n_gates, op, w, t.ref(), c->stream);
    ZK_TRY(check_launch(c, "k_add_mul_i"));
    ZK_CUDA(c, cudaFreeAsync(d_ops, c->stream));
    ZK_CUDA(c, cudaStreamSynchronize(c->stream));  // `ops` is the caller's pageable memory
    *out = put_mle(c, t);
    return ZKB_OK;
}
int32_t zkb_gkr_prove(zkb_ctx* c, zkb_circ h, const uint64_t* inputs, uint64_t n_inputs, uint64_t w0[8], uint64_t* coeffs,
                      int32_t* lens, uint64_t* challenges, uint64_t* claimed, uint64_t final_openings[8], uint32_t* n_rounds) {
    if (!c || !inputs || !w0 || !coeffs || !lens || !final_openings) return ZKB_ERR_BAD_ARG;
    auto it = c->circs.find(h);
    if (it == c->circs.end()) return ZKB_ERR_BAD_ARG;
    if (it->second->L > 1 && !claimed) return ZKB_ERR_BAD_ARG;
    return gkr_prove_impl(c, it->second.get(), inputs, n_inputs, w0, coeffs, lens, challenges, claimed, final_openings, n_rounds);
}

int32_t zkb_gkr_verify(zkb_ctx* c, zkb_circ h, const uint64_t* inputs, uint64_t n_inputs, const uint64_t w0_in[8],
                       const uint64_t* coeffs, const int32_t* lens, const uint64_t* claimed, const uint64_t final_openings[8],
                       int32_t* accepted) {
    if (!c || !inputs || !w0_in || !coeffs || !lens || !final_openings || !accepted) return ZKB_ERR_BAD_ARG;
    auto it = c->circs.find(h);
    if (it == c->circs.end()) return ZKB_ERR_BAD_ARG;
    CircuitState* cs = it->second.get();
    *accepted = 0;
    const int L = cs->L;
    if (n_inputs != 2ull * cs->gates[0]) ZK_FAIL(c, ZKB_ERR_LENGTH_MISMATCH, "circuit: inputs must be 2 x gates of the first layer");
    const HostField& H = c->H;
    {   // untrusted proof elements must be canonical residues (see all_canonical)
        uint32_t tot = 0;
        for (int l = 0; l < L; ++l) tot += layer_rounds(cs->gates[l]);
        if (!all_canonical(H, w0_in, 2) || !all_canonical(H, final_openings, 2) || !coeffs_canonical(H, coeffs, lens, tot, 3) ||
            (L > 1 && (!claimed || !all_canonical(H, claimed, 2 * (size_t)(L - 1)))))
            return ZKB_OK;  // *accepted == 0
    }
    TranscriptImpl tr;
    tr.H = H;
    Fe w0[2] = {fe_from_u64x4(w0_in), fe_from_u64x4(w0_in + 4)};
    tr.append_elements(w0, 2);
    Fe r0 = tr.challenge();
    Fe claim = H.add(w0[0], H.mul(r0, H.sub(w0[1], w0[0])));
    tr.append_elements(&claim, 1);
    // the input MLE on the device (the reference checks a KZG opening instead, gkr_protocol.rs:157-183)
    Table in_tab;
    ZK_TRY(upload_aos(c, inputs, n_inputs, 0, 1, n_inputs, 0, &in_tab));
    Fe alpha = H.zero(), beta = H.zero();
    std::vector<Fe> prb, prc;
    size_t round_base = 0;
    int32_t st = ZKB_OK;
    bool ok = true;
    for (int idx = 0; idx < L && ok && st == ZKB_OK; ++idx) {
        const int l = L - 1 - idx;
        const uint64_t G = cs->gates[l], nw = 2 * G;
        const int nb = ilog2_u64(nw);
        const uint32_t nr = 2 * (uint32_t)nb;
        std::vector<Fe> cur(nr);
        Fe fin;
        for (uint32_t k = 0; k < nr; ++k)
            if (lens[round_base + k] < 0 || lens[round_base + k] > 3) { ok = false; break; }
        if (!ok) break;
        if (!sc_verify_rounds(H, &tr, nr, 3, coeffs + round_base * 12, lens + round_base, claim, &fin, cur.data())) { ok = false; break; }
        round_base += nr;
        std::vector<Fe> u(cur.begin(), cur.begin() + nb), w(cur.begin() + nb, cur.end());
        Fe o1, o2;
        if (idx == L - 1) {
            st = evaluate_table(c, in_tab, u.data(), (uint32_t)nb, &o1);
            if (st == ZKB_OK) st = evaluate_table(c, in_tab, w.data(), (uint32_t)nb, &o2);
            if (st != ZKB_OK) break;
            if (!H.eq(o1, fe_from_u64x4(final_openings)) || !H.eq(o2, fe_from_u64x4(final_openings + 4))) { ok = false; break; }
        } else {
            o1 = fe_from_u64x4(claimed + (size_t)idx * 8);
            o2 = fe_from_u64x4(claimed + (size_t)idx * 8 + 4);
        }
        // wiring predicates at (prev point, u, w) in O(G)
        GkrWiringArgs wa;
        std::memset(&wa, 0, sizeof wa);
        wa.ops = cs->d_ops + cs->opoff[l];
        wa.n_gates = G;
        wa.first_layer = idx == 0;
        wa.r0 = r0;
        wa.alpha = alpha;
        wa.beta = beta;
        int n_lo = 0;
        if (idx > 0) {
            if ((1ull << prb.size()) != G) { ok = false; break; }
            st = eq_tables(c, prb.data(), (int)prb.size(), &cs->eq[0], &cs->eq[1], &n_lo);
            if (st == ZKB_OK) st = eq_tables(c, prc.data(), (int)prc.size(), &cs->eq[2], &cs->eq[3], &n_lo);
            if (st != ZKB_OK) break;
            wa.eb_hi = cs->eq[0].ref();
            wa.eb_lo = cs->eq[1].ref();
            wa.ec_hi = cs->eq[2].ref();
            wa.ec_lo = cs->eq[3].ref();
            wa.n_lo_a = n_lo;
        }
        st = eq_tables(c, u.data(), nb, &cs->eq[4], &cs->eq[5], &n_lo);
        if (st == ZKB_OK) st = eq_tables(c, w.data(), nb, &cs->eq[6], &cs->eq[7], &n_lo);
        if (st != ZKB_OK) break;
        wa.eu_hi = cs->eq[4].ref();
        wa.eu_lo = cs->eq[5].ref();
        wa.ew_hi = cs->eq[6].ref();
        wa.ew_lo = cs->eq[7].ref();
        wa.n_lo_w = n_lo;
        const int grid = grid_for(c, G, 4);
        st = prep_finish(c, grid, 2, false, &wa.fin);
        if (st != ZKB_OK) break;
        c->K->gkr_wiring(wa, grid, c->stream);
        st = check_launch(c, "k_gkr_wiring");
        if (st != ZKB_OK) break;
        Fe am[2];
        st = collect(c, 2, false, wa.fin, am);
        if (st != ZKB_OK) break;
        Fe expected = H.add(H.mul(am[0], H.add(o1, o2)), H.mul(am[1], H.mul(o1, o2)));  // :211,313,340
        if (!H.eq(expected, fin)) { ok = false; break; }
        prb = u;
        prc = w;
        tr.append_elements(&o1, 1);  // :217-221 (the verifier absorbs after every layer)
        alpha = tr.challenge();
        tr.append_elements(&o2, 1);
        beta = tr.challenge();
        claim = H.add(H.mul(alpha, o1), H.mul(beta, o2));
    }
    free_table(c, &in_tab);
    if (st != ZKB_OK) return st;
    *accepted = ok ? 1 : 0;
    return ZKB_OK;
}

// ------------------------------------------------------------ general wiring (extension beyond the reference)
int32_t zkb_circuit_create_wired(zkb_ctx* c, uint32_t n_layers, const uint32_t* gates, uint64_t n_inputs, const uint8_t* ops,
                                 const uint32_t* in1, const uint32_t* in2, zkb_circ* out) {
    if (!c || !gates || !ops || !in1 || !in2 || !out) return ZKB_ERR_BAD_ARG;
    if (n_layers < 1) ZK_FAIL(c, ZKB_ERR_CIRCUIT_SHAPE, "circuit: no layers");
    if (!is_pow2(n_inputs) || n_inputs < 2 || n_inputs > (1ull << 30)) ZK_FAIL(c, ZKB_ERR_CIRCUIT_SHAPE, "circuit: the input count must be a power of two in [2, 2^30]");
    std::unique_ptr<CircuitState> cs(new CircuitState);
    cs->wired = true;
    cs->L = (int)n_layers;
    cs->n_inputs = n_inputs;
    cs->gates.assign(gates, gates + n_layers);
    size_t tot = 0, csr = 0;
    uint64_t wmax = n_inputs, gmax = 1;
    for (uint32_t l = 0; l < n_layers; ++l) {
        const uint64_t w = l == 0 ? n_inputs : gates[l - 1];
        if (!is_pow2(gates[l]) || gates[l] > (1u << 30)) ZK_FAIL(c, ZKB_ERR_CIRCUIT_SHAPE, "circuit: gates per layer must be a power of two");
        if (w < 2) ZK_FAIL(c, ZKB_ERR_CIRCUIT_SHAPE, "circuit: only the output layer may have a single gate");
        cs->width.push_back(w);
        cs->opoff.push_back(tot);
        cs->csroff.push_back(csr);
        tot += gates[l];
        csr += (size_t)w + 1;
        if (w > wmax) wmax = w;
        if (gates[l] > gmax) gmax = gates[l];
    }
    cs->h_ops.assign(ops, ops + tot);
    // CSR lists of the gates by first / second input (counting sort; gate order kept inside a row)
    std::vector<uint32_t> off1(csr, 0), off2(csr, 0), lst1(tot), lst2(tot);
    for (uint32_t l = 0; l < n_layers; ++l) {
        const uint64_t w = cs->width[l];
        const size_t go = cs->opoff[l], co = cs->csroff[l];
        for (uint32_t g = 0; g < gates[l]; ++g) {
            if (ops[go + g] != ZKB_OP_ADD && ops[go + g] != ZKB_OP_MUL) return ZKB_ERR_BAD_ARG;
            if (in1[go + g] >= w || in2[go + g] >= w) ZK_FAIL(c, ZKB_ERR_CIRCUIT_SHAPE, "circuit: wire index out of range");
            ++off1[co + in1[go + g] + 1];
            ++off2[co + in2[go + g] + 1];
        }
        for (uint64_t b = 0; b < w; ++b) {
            off1[co + b + 1] += off1[co + b];
            off2[co + b + 1] += off2[co + b];
        }
        std::vector<uint32_t> p1(off1.begin() + co, off1.begin() + co + w), p2(off2.begin() + co, off2.begin() + co + w);
        for (uint32_t g = 0; g < gates[l]; ++g) {
            lst1[go + p1[in1[go + g]]++] = g;
            lst2[go + p2[in2[go + g]]++] = g;
        }
    }
    auto put = [&](const void* src, size_t bytes, void** dst) -> int32_t {
        ZK_CUDA(c, cudaMallocAsync(dst, bytes, c->stream));
        ZK_CUDA(c, cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
        return ZKB_OK;
    };
    ZK_TRY(put(cs->h_ops.data(), tot, (void**)&cs->d_ops));
    ZK_TRY(put(in1, tot * 4, (void**)&cs->d_in1));
    ZK_TRY(put(in2, tot * 4, (void**)&cs->d_in2));
    ZK_TRY(put(lst1.data(), tot * 4, (void**)&cs->d_lst1));
    ZK_TRY(put(lst2.data(), tot * 4, (void**)&cs->d_lst2));
    ZK_TRY(put(off1.data(), csr * 4, (void**)&cs->d_off1));
    ZK_TRY(put(off2.data(), csr * 4, (void**)&cs->d_off2));
    ZK_CUDA(c, cudaStreamSynchronize(c->stream));
    ZK_TRY(alloc_table(c, n_inputs, &cs->inputs));
    cs->vals.resize(n_layers);
    for (uint32_t l = 0; l < n_layers; ++l) ZK_TRY(alloc_table(c, gates[l], &cs->vals[l]));
    ZK_TRY(alloc_table(c, wmax, &cs->H1));
    ZK_TRY(alloc_table(c, wmax, &cs->HA2));
    ZK_TRY(alloc_table(c, gmax, &cs->coef));
    ZK_TRY(alloc_table(c, gmax, &cs->gtmp));
    const int nmax = ilog2_u64(wmax > gmax ? wmax : gmax);
    for (auto& t : cs->eq) ZK_TRY(alloc_table(c, 1ull << ((nmax + 1) / 2 + 1), &t));
    ZK_TRY(sp_configure(c, &cs->sp, 1, 3, KIND_XYZ));
    cs->sp.work.resize(3);
    for (auto& t : cs->sp.work) ZK_TRY(alloc_table(c, wmax / 2, &t));
    zkb_circ h = c->next_handle++;
    c->circs[h] = std::move(cs);
    *out = h;
    return ZKB_OK;
}

int32_t zkb_circuit_total_rounds(zkb_ctx* c, zkb_circ h, uint32_t* n_rounds) {
    if (!c || !n_rounds) return ZKB_ERR_BAD_ARG;
    auto it = c->circs.find(h);
    if (it == c->circs.end()) return ZKB_ERR_BAD_ARG;
    CircuitState* cs = it->second.get();
    uint32_t tot = 0;
    for (int l = 0; l < cs->L; ++l) tot += cs->wired ? 2u * (uint32_t)ilog2_u64(cs->width[l]) : layer_rounds(cs->gates[l]);
    *n_rounds = tot;
    return ZKB_OK;
}

int32_t zkb_gkr_prove_wired(zkb_ctx* c, zkb_circ h, const uint64_t* inputs, uint64_t n_inputs, uint64_t* w0, uint64_t n_w0,
                            uint64_t* coeffs, int32_t* lens, uint64_t* challenges, uint64_t* claimed, uint64_t final_openings[8],
                            uint32_t* n_rounds) {
    if (!c || !inputs || !w0 || !coeffs || !lens || !final_openings) return ZKB_ERR_BAD_ARG;
    auto it = c->circs.find(h);
    if (it == c->circs.end() || !it->second->wired) return ZKB_ERR_BAD_ARG;
    CircuitState* cs = it->second.get();
    if (cs->L > 1 && !claimed) return ZKB_ERR_BAD_ARG;
    const uint64_t Gout = cs->gates[cs->L - 1];
    if (n_w0 != (Gout < 2 ? 2 : Gout)) ZK_FAIL(c, ZKB_ERR_LENGTH_MISMATCH, "gkr: w0 must hold max(outputs, 2) elements");
    return gkr_prove_wired_impl(c, cs, inputs, n_inputs, w0, coeffs, lens, challenges, claimed, final_openings, n_rounds);
}

int32_t zkb_gkr_verify_wired(zkb_ctx* c, zkb_circ h, const uint64_t* inputs, uint64_t n_inputs, const uint64_t* w0_in, uint64_t n_w0,
                             const uint64_t* coeffs, const int32_t* lens, const uint64_t* claimed, const uint64_t final_openings[8],
                             int32_t* accepted) {
    if (!c || !inputs || !w0_in || !coeffs || !lens || !final_openings || !accepted) return ZKB_ERR_BAD_ARG;
    auto it = c->circs.find(h);
    if (it == c->circs.end() || !it->second->wired) return ZKB_ERR_BAD_ARG;
    CircuitState* cs = it->second.get();
    *accepted = 0;
    const int L = cs->L;
    if (n_inputs != cs->n_inputs) ZK_FAIL(c, ZKB_ERR_LENGTH_MISMATCH, "circuit: wrong number of inputs");
    const uint64_t Gout = cs->gates[L - 1], n0 = Gout < 2 ? 2 : Gout;
    if (n_w0 != n0) ZK_FAIL(c, ZKB_ERR_LENGTH_MISMATCH, "gkr: w0 must hold max(outputs, 2) elements");
    if (L > 1 && !claimed) return ZKB_ERR_BAD_ARG;
    const HostField& H = c->H;
    {   // untrusted proof elements must be canonical residues (see all_canonical)
        uint32_t tot = 0;
        for (int l = 0; l < L; ++l) tot += 2u * (uint32_t)ilog2_u64(cs->width[l]);
        if (!all_canonical(H, w0_in, (size_t)n0) || !all_canonical(H, final_openings, 2) || !coeffs_canonical(H, coeffs, lens, tot, 3) ||
            (L > 1 && !all_canonical(H, claimed, 2 * (size_t)(L - 1))))
            return ZKB_OK;  // *accepted == 0
    }
    TranscriptImpl tr;
    tr.H = H;
    const int k0 = ilog2_u64(n0);
    std::vector<Fe> r0(k0);
    Fe claim;
    if (n0 <= 4096) {
        std::vector<Fe> w0(n0);
        for (uint64_t i = 0; i < n0; ++i) w0[i] = fe_from_u64x4(w0_in + i * 4);
        tr.append_elements(w0.data(), (size_t)n0);
        for (int i = 0; i < k0; ++i) r0[i] = tr.challenge();
        for (int i = 0; i < k0; ++i) {
            const size_t hlf = w0.size() / 2;
            for (size_t j = 0; j < hlf; ++j) w0[j] = H.add(w0[j], H.mul(r0[i], H.sub(w0[j + hlf], w0[j])));
            w0.resize(hlf);
        }
        claim = w0[0];
    } else {  // wide output layer: canonical bytes and the MLE evaluation both come from the device copy
        Table t0;
        ZK_TRY(upload_aos(c, w0_in, n0, 0, 1, n0, 0, &t0));
        int32_t st0 = absorb_table_bytes(c, t0, &tr);
        if (st0 == ZKB_OK) {
            for (int i = 0; i < k0; ++i) r0[i] = tr.challenge();
            st0 = evaluate_table(c, t0, r0.data(), (uint32_t)k0, &claim);
        }
        free_table(c, &t0);
        ZK_TRY(st0);
    }
    tr.append_elements(&claim, 1);
    Table in_tab;
    ZK_TRY(upload_aos(c, inputs, n_inputs, 0, 1, n_inputs, 0, &in_tab));
    Fe alpha = H.zero(), beta = H.zero();
    std::vector<Fe> prb, prc;
    size_t round_base = 0;
    int32_t st = ZKB_OK;
    bool ok = true;
    for (int idx = 0; idx < L && ok && st == ZKB_OK; ++idx) {
        const int l = L - 1 - idx;
        const uint64_t G = cs->gates[l], nw = cs->width[l];
        const int nb = ilog2_u64(nw);
        const uint32_t nr = 2 * (uint32_t)nb;
        std::vector<Fe> cur(nr);
        Fe fin;
        for (uint32_t k = 0; k < nr; ++k)
            if (lens[round_base + k] < 0 || lens[round_base + k] > 3) { ok = false; break; }
        if (!ok) break;
        if (!sc_verify_rounds(H, &tr, nr, 3, coeffs + round_base * 12, lens + round_base, claim, &fin, cur.data())) { ok = false; break; }
        round_base += nr;
        std::vector<Fe> u(cur.begin(), cur.begin() + nb), w(cur.begin() + nb, cur.end());
        Fe o1, o2;
        if (idx == L - 1) {
            st = evaluate_table(c, in_tab, u.data(), (uint32_t)nb, &o1);
            if (st == ZKB_OK) st = evaluate_table(c, in_tab, w.data(), (uint32_t)nb, &o2);
            if (st != ZKB_OK) break;
            if (!H.eq(o1, fe_from_u64x4(final_openings)) || !H.eq(o2, fe_from_u64x4(final_openings + 4))) { ok = false; break; }
        } else {
            o1 = fe_from_u64x4(claimed + (size_t)idx * 8);
            o2 = fe_from_u64x4(claimed + (size_t)idx * 8 + 4);
        }
        GkrWWiringArgs wa;
        std::memset(&wa, 0, sizeof wa);
        int n_lo = 0;
        if (idx == 0) {
            st = eq_tables(c, r0.data(), k0, &cs->eq[0], &cs->eq[1], &n_lo);
        } else {
            if ((1ull << prb.size()) != G) { ok = false; break; }
            st = eq_tables(c, prb.data(), (int)prb.size(), &cs->eq[0], &cs->eq[1], &n_lo);
            if (st == ZKB_OK) st = eq_tables(c, prc.data(), (int)prc.size(), &cs->eq[2], &cs->eq[3], &n_lo);
        }
        if (st != ZKB_OK) break;
        wired_coef_args(cs, idx, alpha, beta, n_lo, &wa.wc);
        st = eq_tables(c, u.data(), nb, &cs->eq[4], &cs->eq[5], &n_lo);
        if (st == ZKB_OK) st = eq_tables(c, w.data(), nb, &cs->eq[6], &cs->eq[7], &n_lo);
        if (st != ZKB_OK) break;
        wa.eu_hi = cs->eq[4].ref();
        wa.eu_lo = cs->eq[5].ref();
        wa.ew_hi = cs->eq[6].ref();
        wa.ew_lo = cs->eq[7].ref();
        wa.n_lo_w = n_lo;
        wa.ops = cs->d_ops + cs->opoff[l];
        wa.in1 = cs->d_in1 + cs->opoff[l];
        wa.in2 = cs->d_in2 + cs->opoff[l];
        wa.n_gates = G;
        const int grid = grid_for(c, G, 4);
        st = prep_finish(c, grid, 2, false, &wa.fin);
        if (st != ZKB_OK) break;
        c->K->gkr_w_wiring(wa, grid, c->stream);
        st = check_launch(c, "k_gkr_w_wiring");
        if (st != ZKB_OK) break;
        Fe am[2];
        st = collect(c, 2, false, wa.fin, am);
        if (st != ZKB_OK) break;
        Fe expected = H.add(H.mul(am[0], H.add(o1, o2)), H.mul(am[1], H.mul(o1, o2)));  // :211,313,340
        if (!H.eq(expected, fin)) { ok = false; break; }
        prb = u;
        prc = w;
        tr.append_elements(&o1, 1);
        alpha = tr.challenge();
        tr.append_elements(&o2, 1);
        beta = tr.challenge();
        claim = H.add(H.mul(alpha, o1), H.mul(beta, o2));
    }
    free_table(c, &in_tab);
    if (st != ZKB_OK) return st;
    *accepted = ok ? 1 : 0;
    return ZKB_OK;
}

// ------------------------------------------------------------- microbenchmarks
// ------------------------------------------------------------ multilinear KZG (input-layer commitment)
int32_t zkb_kzg_setup(zkb_ctx* c, uint32_t n_vars, const uint64_t* taus, zkb_kzg* out) {
    if (!c || !taus || !out) return ZKB_ERR_BAD_ARG;
    ZK_TRY(kzg_require_field(c));
    if (n_vars < 1 || n_vars > 26) ZK_FAIL(c, ZKB_ERR_ARITY, "Invalid num of vars for lagrange basis");  // kzg.rs:184-186
    const uint64_t N = 1ull << n_vars;
    if (!c->g1_table) {
        ZK_CUDA(c, cudaMalloc((void**)&c->g1_table, sizeof(G1Affine) * 32 * 256));
        k_g1_window_table<<<64, 128, 0, c->stream>>>(c->g1_table);
        ZK_TRY(check_launch(c, "k_g1_window_table"));
    }
    // Lagrange scalars eq(taus, .) (kzg.rs:183-207)
    std::vector<Fe> tv(n_vars);
    for (uint32_t i = 0; i < n_vars; ++i) tv[i] = fe_from_u64x4(taus + 4 * i);
    const int n_hi = (int)n_vars / 2, n_lo = (int)n_vars - n_hi;
    Table hi, lo, L;
    ZK_TRY(alloc_table(c, 1ull << n_hi, &hi));
    ZK_TRY(alloc_table(c, 1ull << n_lo, &lo));
    ZK_TRY(alloc_table(c, N, &L));
    int nlo_out = 0;
    ZK_TRY(eq_tables(c, tv.data(), (int)n_vars, &hi, &lo, &nlo_out));
    k_kzg_eq_full<<<grid_for(c, N, 8), BLOCK, 0, c->stream>>>(hi.ref(), lo.ref(), nlo_out, L.ref(), N);
    ZK_TRY(check_launch(c, "k_kzg_eq_full"));
    std::unique_ptr<KzgState> ks(new KzgState);
    ks->n_vars = n_vars;
    ks->basis.assign(n_vars + 1, nullptr);
    G1Jac* jac = nullptr;
    ZK_CUDA(c, cudaMallocAsync((void**)&jac, sizeof(G1Jac) * N, c->stream));
    ZK_CUDA(c, cudaMallocAsync((void**)&ks->basis[0], sizeof(G1Affine) * N, c->stream));
    k_g1_fixed_base<<<grid_for(c, N, 8) * 2, 128, 0, c->stream>>>(L.ref(), c->g1_table, jac, N);  // g1 * scalar, kzg.rs:209-212
    ZK_TRY(check_launch(c, "k_g1_fixed_base"));
    ZK_TRY(jac_to_affine(c, jac, ks->basis[0], N));
    for (uint32_t k = 1; k <= n_vars; ++k) {  // folded bases: level k[j] = level k-1[j] + level k-1[j + half]
        const uint64_t half = N >> k;
        ZK_CUDA(c, cudaMallocAsync((void**)&ks->basis[k], sizeof(G1Affine) * half, c->stream));
        k_g1_fold_basis<<<grid_for(c, half, 8) * 2, 128, 0, c->stream>>>(ks->basis[k - 1], jac, half);
        ZK_TRY(check_launch(c, "k_g1_fold_basis"));
        ZK_TRY(jac_to_affine(c, jac, ks->basis[k], half));
    }
    cudaFreeAsync(jac, c->stream);
    free_table(c, &hi);
    free_table(c, &lo);
    free_table(c, &L);
    zkb_kzg h = c->next_handle++;
    c->kzgs[h] = std::move(ks);
    *out = h;
    return ZKB_OK;
}
int32_t zkb_kzg_free(zkb_ctx* c, zkb_kzg h) {
    if (!c) return ZKB_ERR_BAD_ARG;
    auto it = c->kzgs.find(h);
    if (it == c->kzgs.end()) return ZKB_ERR_BAD_ARG;
    kzg_release(c, it->second.get());
    c->kzgs.erase(it);
    return ZKB_OK;
}
int32_t zkb_kzg_basis(zkb_ctx* c, zkb_kzg h, uint32_t level, uint64_t first, uint64_t count, uint8_t* out) {
    if (!c || !out) return ZKB_ERR_BAD_ARG;
    auto it = c->kzgs.find(h);
    if (it == c->kzgs.end() || level > it->second->n_vars) return ZKB_ERR_BAD_ARG;
    const uint64_t n = 1ull << (it->second->n_vars - level);
    if (first + count > n) return ZKB_ERR_BAD_ARG;
    if (!count) return ZKB_OK;
    uint8_t* d = nullptr;
    ZK_CUDA(c, cudaMallocAsync((void**)&d, count * 96, c->stream));
    k_g1_export<<<(unsigned)((count + 127) / 128), 128, 0, c->stream>>>(it->second->basis[level] + first, count, d);
    ZK_TRY(check_launch(c, "k_g1_export"));
    ZK_CUDA(c, cudaMemcpyAsync(out, d, count * 96, cudaMemcpyDeviceToHost, c->stream));
    ZK_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFreeAsync(d, c->stream);
    return ZKB_OK;
}
int32_t zkb_kzg_commit(zkb_ctx* c, zkb_kzg h, zkb_mle poly, uint8_t out[96]) {
    if (!c || !out) return ZKB_ERR_BAD_ARG;
    ZK_TRY(kzg_require_field(c));
    auto it = c->kzgs.find(h);
    Table* t = find_mle(c, poly);
    if (it == c->kzgs.end() || !t) return ZKB_ERR_BAD_ARG;
    if (t->n != (1ull << it->second->n_vars)) ZK_FAIL(c, ZKB_ERR_LENGTH_MISMATCH, "invalid polynomial or lagrange basis");  // kzg.rs:135-137
    uint8_t* d = nullptr;
    ZK_CUDA(c, cudaMallocAsync((void**)&d, 96, c->stream));
    ZK_TRY(msm_run(c, *t, it->second->basis[0], t->n, d, c->stream));
    ZK_CUDA(c, cudaMemcpyAsync(out, d, 96, cudaMemcpyDeviceToHost, c->stream));
    ZK_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFreeAsync(d, c->stream);
    return ZKB_OK;
}
int32_t zkb_kzg_open(zkb_ctx* c, zkb_kzg h, zkb_mle poly, const uint64_t* opening_values, uint32_t n, uint64_t out[4]) {
    if (!c) return ZKB_ERR_BAD_ARG;
    if (c->kzgs.find(h) == c->kzgs.end()) return ZKB_ERR_BAD_ARG;
    return zkb_mle_evaluate(c, poly, opening_values, n, out);  // kzg.rs:55-57
}
int32_t zkb_kzg_get_proof(zkb_ctx* c, zkb_kzg h, zkb_mle poly, const uint64_t opened_value[4], const uint64_t* opening_values, uint32_t n,
                          uint8_t* out) {
    (void)opened_value;  // the quotients do not depend on the constant shift poly - v (see kzg_impl.cuh)
    if (!c || !opening_values || !out) return ZKB_ERR_BAD_ARG;
    ZK_TRY(kzg_require_field(c));
    auto it = c->kzgs.find(h);
    Table* t = find_mle(c, poly);
    if (it == c->kzgs.end() || !t) return ZKB_ERR_BAD_ARG;
    KzgState* ks = it->second.get();
    if (t->n != (1ull << ks->n_vars)) ZK_FAIL(c, ZKB_ERR_LENGTH_MISMATCH, "invalid polynomial or lagrange basis");
    if (n > ks->n_vars) ZK_FAIL(c, ZKB_ERR_ARITY, "Invalid number of values");
    uint8_t* d = nullptr;
    ZK_CUDA(c, cudaMallocAsync((void**)&d, 96 * (size_t)(n ? n : 1), c->stream));
    // The n quotient commitments are independent of each other (only the remainder chain is sequential), and a small MSM
    // is latency: a single-warp window combination and a few CTAs of bucket sums.  Every level gets its own slice of one
    // quotient table and its MSM runs on one of four side streams behind an event, so the serial tails overlap.
    for (int i = 0; i < 4; ++i) {
        if (!c->kzg_streams[i]) ZK_CUDA(c, cudaStreamCreateWithFlags(&c->kzg_streams[i], cudaStreamNonBlocking));
        if (!c->kzg_ev[i]) ZK_CUDA(c, cudaEventCreateWithFlags(&c->kzg_ev[i], cudaEventDisableTiming));
    }
    Table cur = *t, work, q;
    ZK_TRY(alloc_table(c, t->n / 2, &work));
    ZK_TRY(alloc_table(c, t->n, &q));  // level k (half_k entries) at offset n - 2 half_k: the slices tile [0, n)
    cudaEvent_t ready = nullptr;
    ZK_CUDA(c, cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    for (uint32_t k = 0; k < n; ++k) {
        const uint64_t half = cur.n / 2;
        Table qk = q;
        qk.base = q.base + (t->n - 2 * half);
        qk.n = half;
        k_kzg_quotient<<<grid_for(c, half, 8), BLOCK, 0, c->stream>>>(cur.ref(), qk.ref(), half);  // kzg.rs:152-163
        ZK_TRY(check_launch(c, "k_kzg_quotient"));
        cudaStream_t side = c->kzg_streams[k & 3];
        ZK_CUDA(c, cudaEventRecord(ready, c->stream));
        ZK_CUDA(c, cudaStreamWaitEvent(side, ready, 0));
        ZK_TRY(msm_run(c, qk, ks->basis[k + 1], half, d + 96 * (size_t)k, side));                    // :80-83 on the folded basis
        FixedMul rt;                                                                               // remainder: partial_evaluate(0, z_k), :146-150
        c->fmb.make(c->H, fe_from_u64x4(opening_values + 4 * k), &rt);
        c->K->fold(cur.ref(), work.ref(), half, (uint32_t)ilog2_u64(half), rt, grid_for(c, half, 8), c->stream);
        ZK_TRY(check_launch(c, "k_fold"));
        cur = work;
        cur.n = half;
    }
    for (int i = 0; i < 4; ++i) {  // the main stream continues after every side stream
        ZK_CUDA(c, cudaEventRecord(c->kzg_ev[i], c->kzg_streams[i]));
        ZK_CUDA(c, cudaStreamWaitEvent(c->stream, c->kzg_ev[i], 0));
    }
    cudaEventDestroy(ready);
    ZK_CUDA(c, cudaMemcpyAsync(out, d, 96 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    ZK_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFreeAsync(d, c->stream);
    free_table(c, &work);
    free_table(c, &q);
    return ZKB_OK;
}

// ------------------------------------------------------------ fft/src/fft.rs
namespace {
// ark-ff: TWO_ADIC_ROOT_OF_UNITY = GENERATOR^((p - 1) / 2^TWO_ADICITY); get_root_of_unity(n) squares it TWO_ADICITY - log n times.
// Generators and two-adicities of ark-bn254 0.5.0 Fr (5, 28), Fq (3, 1) and ark-bls12-381 0.5.0 Fr (7, 32).
int32_t root_of_unity(zkb_ctx* c, int log_n, Fe* omega) {
    static const uint32_t gen[3] = {5, 3, 7};
    static const int adicity[3] = {28, 1, 32};
    const int s = adicity[c->field];
    if (log_n > s) ZK_FAIL(c, ZKB_ERR_UNSUPPORTED, "the field has no root of unity of that order");
    Fe e = c->H.modulus();  // (p - 1) >> s: p is odd, so p - 1 clears bit 0 and the shift drops the s zero bits
    e.l[0] &= ~1u;
    for (int k = 0; k < s; ++k) {
        for (int i = 0; i < 7; ++i) e.l[i] = (e.l[i] >> 1) | (e.l[i + 1] << 31);
        e.l[7] >>= 1;
    }
    const Fe g = c->H.from_u64(gen[c->field]);
    Fe r = c->H.one();
    for (int i = 255; i >= 0; --i) {
        r = c->H.mul(r, r);
        if ((e.l[i / 32] >> (i % 32)) & 1) r = c->H.mul(r, g);
    }
    for (int k = log_n; k < s; ++k) r = c->H.mul(r, r);
    *omega = r;
    return ZKB_OK;
}
// y[j] = sum_i x_i w^(i j) (inverse: w^-1 and a final n^-1), natural order in and out (fft.rs:6-60)
int32_t ntt_table(zkb_ctx* c, const Table& src, bool inverse, Table* out) {
    const uint64_t n = src.n;
    const int log_n = ilog2_u64(n);
    Fe omega;
    ZK_TRY(root_of_unity(c, log_n, &omega));
    if (n == 1) return multi_fold(c, src, nullptr, 0, out);
    if (inverse) omega = c->H.inv(omega);
    NttPows pw;
    pw.w[0] = omega;
    for (int i = 1; i < 32; ++i) pw.w[i] = c->H.mul(pw.w[i - 1], pw.w[i - 1]);
    const uint32_t lo_bits = (uint32_t)(log_n - 1 < 12 ? log_n - 1 : 12);
    const uint64_t n_hi = (n >> 1) >> lo_bits;
    Table w_lo, w_hi, dst;
    ZK_TRY(alloc_table(c, 1ull << lo_bits, &w_lo));
    ZK_TRY(alloc_table(c, n_hi, &w_hi));
    ZK_TRY(alloc_table(c, n, &dst));
    c->K->ntt_twiddles(w_lo.ref(), lo_bits, w_hi.ref(), n_hi, pw, grid_for(c, (1ull << lo_bits) + n_hi, 8), c->stream);
    ZK_TRY(check_launch(c, "k_ntt_twiddles"));
    NttArgs a;
    std::memset(&a, 0, sizeof a);
    a.in = src.ref();
    a.data = dst.ref();
    a.log_n = (uint32_t)log_n;
    a.w_lo = w_lo.ref();
    a.w_hi = w_hi.ref();
    a.lo_bits = lo_bits;
    a.scale = c->H.inv(c->H.from_u64(n));
    for (int s0 = 0; s0 < log_n;) {
        const int g = s0 == 0 ? (log_n < NTT_TILE_LOG ? log_n : NTT_TILE_LOG) : (log_n - s0 < 6 ? log_n - s0 : 6);
        a.s0 = (uint32_t)s0;
        a.g = (uint32_t)g;
        a.do_scale = inverse && s0 + g == log_n;
        const uint64_t tiles = n >> (s0 == 0 ? g : NTT_TILE_LOG);
        prof_begin(c, ZKB_K_OTHER, 64.0 * (double)n);
        c->K->ntt_pass(a, (int)(tiles < (uint64_t)c->sm_count * 8 ? tiles : (uint64_t)c->sm_count * 8), c->stream);
        ZK_TRY(check_launch(c, "k_ntt_pass"));
        s0 += g;
    }
    free_table(c, &w_lo);
    free_table(c, &w_hi);
    *out = dst;
    return ZKB_OK;
}
int32_t fft_host(zkb_ctx* c, const uint64_t* in, uint64_t n, uint64_t* out, bool inverse) {
    if (!c || !in || !out) return ZKB_ERR_BAD_ARG;
    if (!is_pow2(n)) ZK_FAIL(c, ZKB_ERR_NOT_POW2, "Length must be a power of 2");
    Table t, r;
    ZK_TRY(upload_aos(c, in, n, 0, 1, n, 0, &t));
    int32_t st = ntt_table(c, t, inverse, &r);
    free_table(c, &t);
    ZK_TRY(st);
    st = download_aos(c, r, out, 0);
    free_table(c, &r);
    return st;
}
}  // namespace
int32_t zkb_fft_evaluate(zkb_ctx* c, const uint64_t* coeffs, uint64_t n, uint64_t* evals) { return fft_host(c, coeffs, n, evals, false); }
int32_t zkb_fft_interpolate(zkb_ctx* c, const uint64_t* evals, uint64_t n, uint64_t* coeffs) { return fft_host(c, evals, n, coeffs, true); }
int32_t zkb_mle_ntt(zkb_ctx* c, zkb_mle in, int32_t inverse, zkb_mle* out) {
    Table* t = c ? find_mle(c, in) : nullptr;
    if (!t || !out) return ZKB_ERR_BAD_ARG;
    Table r;
    ZK_TRY(ntt_table(c, *t, inverse != 0, &r));
    *out = put_mle(c, r);
    return ZKB_OK;
}

// ------------------------------------------------------------ merkle_tree/src/merkle_tree.rs
namespace {
MerkleState* find_merkle(zkb_ctx* c, zkb_merkle h) {
    if (!c) return nullptr;
    auto it = c->merkles.find(h);
    return it == c->merkles.end() ? nullptr : it->second.get();
}
// compute_hash / hash_pair on the host (:201-214)
Fe merkle_hash_host(const HostField& H, const Fe& a, const Fe* b) {
    Keccak256 k;
    Fe ca = H.from_mont(a);
    k.update(reinterpret_cast<const uint8_t*>(ca.l), 32);
    if (b) {
        Fe cb = H.from_mont(*b);
        k.update(reinterpret_cast<const uint8_t*>(cb.l), 32);
    }
    uint8_t dg[32];
    k.finalize_reset(dg);
    return H.from_le_bytes_mod_order(dg);
}
}  // namespace
int32_t zkb_merkle_build(zkb_ctx* c, const uint64_t* inputs, uint64_t n_inputs, uint32_t depth, zkb_merkle* out) {
    if (!c || !out || (n_inputs && !inputs)) return ZKB_ERR_BAD_ARG;
    if (depth == 0 || depth > 32) ZK_FAIL(c, ZKB_ERR_BAD_ARG, "merkle tree depth must be 1..32");
    const uint64_t n_leaves = 1ull << depth;
    if (n_inputs > n_leaves) ZK_FAIL(c, ZKB_ERR_BAD_ARG, "Too many inputs for tree depth");
    auto ms = std::make_unique<MerkleState>();
    ms->depth = depth;
    ZK_CUDA(c, cudaMalloc((void**)&ms->tree, sizeof(Fe) * ((2ull << depth) - 1)));
    Fe* d_in = nullptr;
    if (n_inputs) {
        ZK_CUDA(c, cudaMallocAsync((void**)&d_in, sizeof(Fe) * n_inputs, c->stream));
        ZK_CUDA(c, cudaMemcpyAsync(d_in, inputs, sizeof(Fe) * n_inputs, cudaMemcpyHostToDevice, c->stream));
    }
    c->K->merkle_leaves(d_in, n_inputs, ms->tree, n_leaves, grid_for(c, n_leaves, 8), c->stream);
    ZK_TRY(check_launch(c, "k_merkle_leaves"));
    for (uint32_t level = 0; level < depth; ++level) {
        const uint64_t n_next = n_leaves >> (level + 1);
        c->K->merkle_level(ms->tree + merkle_level_off(depth, level), ms->tree + merkle_level_off(depth, level + 1), n_next, grid_for(c, n_next, 8), c->stream);
        ZK_TRY(check_launch(c, "k_merkle_level"));
    }
    if (d_in) cudaFreeAsync(d_in, c->stream);
    ZK_CUDA(c, cudaStreamSynchronize(c->stream));  // the caller's input buffer is free again
    const uint64_t h = c->next_handle++;
    c->merkles[h] = std::move(ms);
    *out = h;
    return ZKB_OK;
}
int32_t zkb_merkle_free(zkb_ctx* c, zkb_merkle h) {
    MerkleState* ms = find_merkle(c, h);
    if (!ms) return ZKB_ERR_BAD_ARG;
    cudaFree(ms->tree);
    c->merkles.erase(h);
    return ZKB_OK;
}
int32_t zkb_merkle_nodes(zkb_ctx* c, zkb_merkle h, uint32_t level, uint64_t first, uint64_t count, uint64_t* out) {
    MerkleState* ms = find_merkle(c, h);
    if (!ms || !out || level > ms->depth || first + count > (1ull << (ms->depth - level))) return ZKB_ERR_BAD_ARG;
    ZK_CUDA(c, cudaMemcpyAsync(out, ms->tree + merkle_level_off(ms->depth, level) + first, sizeof(Fe) * count, cudaMemcpyDeviceToHost, c->stream));
    ZK_CUDA(c, cudaStreamSynchronize(c->stream));
    return ZKB_OK;
}
int32_t zkb_merkle_root(zkb_ctx* c, zkb_merkle h, uint64_t out[4]) {
    MerkleState* ms = find_merkle(c, h);
    if (!ms || !out) return ZKB_ERR_BAD_ARG;
    return zkb_merkle_nodes(c, h, ms->depth, 0, 1, out);
}
int32_t zkb_merkle_update_leaf(zkb_ctx* c, zkb_merkle h, uint64_t leaf_id, const uint64_t data[4], int32_t is_hash) {
    MerkleState* ms = find_merkle(c, h);
    if (!ms || !data) return ZKB_ERR_BAD_ARG;
    if (leaf_id >= (1ull << ms->depth)) ZK_FAIL(c, ZKB_ERR_BAD_ARG, "Invalid leaf ID");
    c->K->merkle_path(ms->tree, ms->depth, leaf_id, fe_from_u64x4(data), is_hash, 0, nullptr, nullptr, c->stream);
    ZK_TRY(check_launch(c, "k_merkle_path"));
    return ZKB_OK;
}
int32_t zkb_merkle_create_proof(zkb_ctx* c, zkb_merkle h, const uint64_t data[4], uint64_t leaf_id, uint64_t* sibling_hashes, uint8_t* sides) {
    MerkleState* ms = find_merkle(c, h);
    if (!ms || !data || !sibling_hashes || !sides) return ZKB_ERR_BAD_ARG;
    if (leaf_id >= (1ull << ms->depth)) ZK_FAIL(c, ZKB_ERR_BAD_ARG, "Invalid leaf ID");
    Fe* d_sib = c->d_res;  // 64 elements: depth <= 32
    c->K->merkle_path(ms->tree, ms->depth, leaf_id, fe_from_u64x4(data), 0, 1, d_sib, c->d_ticket, c->stream);
    ZK_TRY(check_launch(c, "k_merkle_path"));
    unsigned int status = 0;
    ZK_CUDA(c, cudaMemcpyAsync(&status, c->d_ticket, sizeof status, cudaMemcpyDeviceToHost, c->stream));
    ZK_CUDA(c, cudaMemcpyAsync(sibling_hashes, d_sib, sizeof(Fe) * ms->depth, cudaMemcpyDeviceToHost, c->stream));
    ZK_CUDA(c, cudaMemsetAsync(c->d_ticket, 0, sizeof(unsigned int), c->stream));
    ZK_CUDA(c, cudaStreamSynchronize(c->stream));
    if (status) ZK_FAIL(c, ZKB_ERR_BAD_ARG, "Data does not match the leaf hash");
    for (uint32_t level = 0; level < ms->depth; ++level) sides[level] = ((leaf_id >> level) & 1) ? 0 : 1;  // even index: the sibling is on the Right (:160-164)
    return ZKB_OK;
}
int32_t zkb_merkle_verify(zkb_ctx* c, zkb_merkle h, const uint64_t data[4], const uint64_t* sibling_hashes, const uint8_t* sides, uint32_t n, int32_t* ok) {
    MerkleState* ms = find_merkle(c, h);
    if (!ms || !data || !ok || (n && (!sibling_hashes || !sides))) return ZKB_ERR_BAD_ARG;
    uint64_t root[4];
    ZK_TRY(zkb_merkle_root(c, h, root));
    Fe cur = merkle_hash_host(c->H, fe_from_u64x4(data), nullptr);
    for (uint32_t i = 0; i < n; ++i) {
        const Fe sib = fe_from_u64x4(sibling_hashes + 4 * i);
        cur = sides[i] == 0 ? merkle_hash_host(c->H, sib, &cur) : merkle_hash_host(c->H, cur, &sib);  // Left: (sibling, current), :190-193
    }
    *ok = c->H.eq(cur, fe_from_u64x4(root)) ? 1 : 0;
    return ZKB_OK;
}

int32_t zkb_bench_modmul(zkb_ctx* c, int32_t variant, uint32_t iters, double* out) {
    if (!c || !out || variant < 0 || variant > 3) return ZKB_ERR_BAD_ARG;
    const int grid = c->sm_count * 8;
    Fe* buf = nullptr;
    ZK_CUDA(c, cudaMallocAsync((void**)&buf, (size_t)grid * BLOCK * sizeof(Fe), c->stream));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    c->K->bench_mul(variant, buf, iters / 8 + 1, grid, c->stream);  // warm-up
    cudaEventRecord(e0, c->stream);
    c->K->bench_mul(variant, buf, iters, grid, c->stream);
    cudaEventRecord(e1, c->stream);
    ZK_CUDA(c, cudaStreamSynchronize(c->stream));
    c->launches += 2;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFreeAsync(buf, c->stream);
    const int ilp = (variant & 1) ? 2 : 1;
    *out = (double)grid * BLOCK * (double)iters * ilp / (ms * 1e-3);
    return ZKB_OK;
}
int32_t zkb_bench_imad(zkb_ctx* c, int32_t mode, uint32_t iters, double* out) {
    if (!c || !out || mode < 0 || mode > 3) return ZKB_ERR_BAD_ARG;
    const int grid = c->sm_count * 8;
    uint64_t* buf = nullptr;
    ZK_CUDA(c, cudaMallocAsync((void**)&buf, (size_t)grid * BLOCK * sizeof(uint64_t), c->stream));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch_bench_imad(mode, buf, iters / 8 + 1, grid, c->stream);
    cudaEventRecord(e0, c->stream);
    launch_bench_imad(mode, buf, iters, grid, c->stream);
    cudaEventRecord(e1, c->stream);
    ZK_CUDA(c, cudaStreamSynchronize(c->stream));
    c->launches += 2;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFreeAsync(buf, c->stream);
    *out = (double)grid * BLOCK * (double)iters * 32.0 / (ms * 1e-3);
    return ZKB_OK;
}

}  // extern "C"
