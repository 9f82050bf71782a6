// comm.cu — data-parallel exchanges for one-process-per-GPU training on one NVLink/NVSwitch box.
//
// The reference has no distributed code; what must cross GPUs follows from its batch-global semantics: gradient SUMS
// (models.py:200-205 never divides by the batch), the whole-tensor L2 norms behind every clip (models.py:214,
// preprocessing.py:220-224), BatchNormalization's per-position batch statistics (layers.py:1237-1238) and the two batch
// reductions of its backward (layers.py:1261-1270). All of them are sums, so one primitive serves: an in-place
// sum-allreduce of a device buffer, in two implementations:
//
//   * NCCL (ncclAllReduce over NVLink 5 / NVSwitch) for the gradient buckets — bandwidth-bound payloads. libnccl is
//     resolved at run time from the copy the process already loaded (PyTorch bundles it); nothing links against it.
//   * a ONE-SHOT peer-memory exchange for the latency-bound payloads (1-5 fp64 clip norms, BatchNorm statistic
//     vectors): every rank owns a "symmetric" buffer mapped into every peer with CUDA IPC; a rank STORES its
//     contribution straight into each peer's inbox over NVLink, raises a per-(site, rank, CTA) flag with a
//     system-scope release, spins on its own flags with system-scope acquires and then sums the inbox in RANK
//     ORDER — so every rank computes bit-identical sums (the clip multipliers derived from them must agree exactly
//     or the replicas drift apart). One kernel per exchange, no host involvement, capturable into the step graph;
//     inboxes are double-buffered on the parity of a device-resident round counter so that a fast rank can never
//     overwrite data a slow rank is still summing.
#include <dlfcn.h>
#include <stddef.h>
#include <string.h>

#include "comm.cuh"   // nccl.h types and prototypes only; every call goes through dlsym'd pointers

namespace b200nn {

struct NcclApi {
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    void* handle = nullptr;
};
static NcclApi g_nccl;

static int load_nccl(const char* path) {
    if (g_nccl.AllReduce != nullptr) return B200NN_OK;
    const char* cands[] = {path, "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* c : cands) {
        if (c == nullptr || c[0] == 0) continue;
        h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (h != nullptr) break;
    }
    if (h == nullptr) {
        set_error("libnccl.so.2 could not be loaded: %s", dlerror());
        return B200NN_ENOTSUP;
    }
    g_nccl.handle = h;
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(h, "ncclAllReduce"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    g_nccl.GetVersion = reinterpret_cast<decltype(g_nccl.GetVersion)>(dlsym(h, "ncclGetVersion"));
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
        set_error("libnccl does not export the expected entry points");
        g_nccl = NcclApi();
        return B200NN_ENOTSUP;
    }
    return B200NN_OK;
}

#define B200NN_NCCL(expr)                                                                                  \
    do {                                                                                                   \
        ncclResult_t _r = (expr);                                                                          \
        if (_r != ncclSuccess) {                                                                           \
            ::b200nn::set_error("%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?"); \
            return B200NN_ECUDA;                                                                           \
        }                                                                                                  \
    } while (0)

__device__ __forceinline__ uint4 ld_inbox16(const uint4* p) {
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
template <class T>
__device__ __forceinline__ void add16(uint4& acc, const uint4& v) {
    constexpr int N = 16 / sizeof(T);
    T* a = reinterpret_cast<T*>(&acc);
    const T* b = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] += b[i];
}

// In-place sum of buf[0..n) over the ranks. Grid <= MAX_XCTAS CTAs (all resident: a CTA only ever waits for the
// CTA with the same index on the other GPUs, which pushes before it waits). The payload moves as 16-byte vectors
// (one NVLink store per peer per vector); `buf` must be 16-byte aligned, a tail shorter than 16 bytes goes scalar in CTA 0.
template <class T>
__global__ void __launch_bounds__(256) exchange_sum_kernel(PeerPtrs peers, int rank, int world, int site, size_t data_off,
                                                           size_t rank_stride, T* __restrict__ buf, long long n, int vec_ok) {
    constexpr int EPV = 16 / sizeof(T);
    SymHeader* me = reinterpret_cast<SymHeader*>(peers.p[rank]);
    const unsigned seq = *reinterpret_cast<volatile unsigned*>(&me->seq[site]) + 1u;
    const unsigned par = seq & 1u;
    const long long nvec = vec_ok ? n / EPV : 0;   // an unaligned (small) buffer goes scalar through CTA 0
    const long long per = (nvec + gridDim.x - 1) / gridDim.x;
    const long long lo = min(nvec, (long long)blockIdx.x * per), hi = min(nvec, lo + per);
    const size_t my_off = data_off + ((size_t)par * world + rank) * rank_stride;
    // push this CTA's slice into every rank's inbox (own rank included: the reduction below then reads one layout)
    const uint4* src4 = reinterpret_cast<const uint4*>(buf);
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const uint4 v = src4[i];
        for (int p = 0; p < world; ++p) reinterpret_cast<uint4*>(peers.p[p] + my_off)[i] = v;
    }
    if (blockIdx.x == 0) {
        for (long long i = nvec * EPV + threadIdx.x; i < n; i += blockDim.x) {
            const T v = buf[i];
            for (int p = 0; p < world; ++p) reinterpret_cast<T*>(peers.p[p] + my_off)[i] = v;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        SymHeader* peer = reinterpret_cast<SymHeader*>(peers.p[threadIdx.x]);
        st_release_sys(&peer->flags[site][par][rank][blockIdx.x], seq);
        const unsigned* mine = &me->flags[site][par][threadIdx.x][blockIdx.x];
        wait_flag(mine, seq, &me->timeouts);
    }
    __syncthreads();
    const char* inbox = peers.p[rank] + data_off + (size_t)par * world * rank_stride;
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        uint4 acc = ld_inbox16(reinterpret_cast<const uint4*>(inbox) + i);
        for (int p = 1; p < world; ++p)                                                   // fixed rank order
            add16<T>(acc, ld_inbox16(reinterpret_cast<const uint4*>(inbox + (size_t)p * rank_stride) + i));
        reinterpret_cast<uint4*>(buf)[i] = acc;
    }
    if (blockIdx.x == 0) {
        for (long long i = nvec * EPV + threadIdx.x; i < n; i += blockDim.x) {
            T s2 = ld_inbox(reinterpret_cast<const T*>(inbox) + i);
            for (int p = 1; p < world; ++p) s2 += ld_inbox(reinterpret_cast<const T*>(inbox + (size_t)p * rank_stride) + i);
            buf[i] = s2;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&me->done[site], 1u) == gridDim.x - 1) {
            me->done[site] = 0u;
            __threadfence();
            *reinterpret_cast<volatile unsigned*>(&me->seq[site]) = seq;
        }
    }
}

// Latency path for <= 4 KB (the clip-norm slots, small gradient buckets): low-latency protocol -- every 4-byte word travels
// as ONE 8-byte store {bits, round number} into each peer's inbox and is polled by the receiving thread (no fence, no flag,
// no barrier); sums are formed in rank order. Inbox: word[2 parity][world][rank_words] of 8 bytes (the site's stride is
// twice the payload).
constexpr int LL_MAX_BYTES = 4096;                       // up to 1024 four-byte words = one CTA
template <class T>
__global__ void __launch_bounds__(1024) exchange_sum_ll_kernel(PeerPtrs peers, int rank, int world, int site, size_t data_off,
                                                               size_t rank_words, T* __restrict__ buf, int n) {
    constexpr int WPE = sizeof(T) / 4;                   // 4-byte words per element
    SymHeader* me = reinterpret_cast<SymHeader*>(peers.p[rank]);
    const unsigned seq = *reinterpret_cast<volatile unsigned*>(&me->seq[site]) + 1u;
    const unsigned par = seq & 1u;
    const int lane = threadIdx.x;                        // word index
    const int nwords = n * WPE;
    unsigned mine = 0;
    if (lane < nwords) mine = reinterpret_cast<const unsigned*>(buf)[lane];
    if (lane < nwords) {
        for (int p = 0; p < world; ++p) {
            uint2* dst = reinterpret_cast<uint2*>(peers.p[p] + data_off) + ((size_t)par * world + rank) * rank_words + lane;
            asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(mine), "r"(seq) : "memory");
        }
    }
    // every rank's word is polled in the same sweep (loads in flight together: one round trip per sweep, not per rank)
    unsigned wv[MAX_WORLD];
#pragma unroll
    for (int r = 0; r < MAX_WORLD; ++r) wv[r] = 0;
    if (lane < nwords) {
        const uint2* base = reinterpret_cast<const uint2*>(peers.p[rank] + data_off) + (size_t)par * world * rank_words + lane;
        unsigned got = 0, spins = 0;
        const unsigned full = (1u << world) - 1u;
        unsigned long long t0 = 0;
        for (;;) {
            unsigned vx[MAX_WORLD], vy[MAX_WORLD];
#pragma unroll
            for (int r = 0; r < MAX_WORLD; ++r) {
                vx[r] = 0; vy[r] = 0;
                if (r < world && !((got >> r) & 1u))
                    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(vx[r]), "=r"(vy[r]) : "l"(base + (size_t)r * rank_words) : "memory");
            }
#pragma unroll
            for (int r = 0; r < MAX_WORLD; ++r)
                if (r < world && !((got >> r) & 1u) && vy[r] == seq) { got |= 1u << r; wv[r] = vx[r]; }
            if (got == full) break;
            if ((++spins & 0x3ffu) == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > WAIT_TIMEOUT_NS) { atomicAdd(&me->timeouts, 1u); break; }
            }
        }
    }
    __syncwarp();
    T total = T(0);
#pragma unroll
    for (int r = 0; r < MAX_WORLD; ++r) {          // sums in rank order: bit-identical on every rank
        if (r < world) {
            if (WPE == 1) {
                total += static_cast<T>(__uint_as_float(wv[r]));
            } else {                                 // lane 2i holds the low word, lane 2i+1 the high word
                const unsigned hi = __shfl_down_sync(0xffffffffu, wv[r], 1);
                total += static_cast<T>(__hiloint2double((int)hi, (int)wv[r]));
            }
        }
    }
    if (WPE == 1) { if (lane < n) buf[lane] = total; }
    else if ((lane & 1) == 0 && (lane >> 1) < n) buf[lane >> 1] = total;
    __syncthreads();
    if (lane == 0) {
        __threadfence();
        *reinterpret_cast<volatile unsigned*>(&me->seq[site]) = seq;
    }
}

// Two payloads in ONE latency exchange: n_a float32 words followed (from an even word index) by n_b float64 values. The train step's
// last small gradient bucket and its clip slots / loss accumulators then cross the GPUs in one NVLink round trip instead of two
// back-to-back kernels (~9 us per step at config 2).
__global__ void __launch_bounds__(1024) exchange_sum_ll2_kernel(PeerPtrs peers, int rank, int world, int site, size_t data_off,
                                                                size_t rank_words, float* __restrict__ a, int n_a,
                                                                double* __restrict__ b, int n_b) {
    SymHeader* me = reinterpret_cast<SymHeader*>(peers.p[rank]);
    const unsigned seq = *reinterpret_cast<volatile unsigned*>(&me->seq[site]) + 1u;
    const unsigned par = seq & 1u;
    const int lane = threadIdx.x;
    const int b0 = (n_a + 1) & ~1;                       // first word of the float64 segment (even, so pairs sit in one warp)
    const int nwords = b0 + 2 * n_b;
    const bool in_a = lane < n_a, in_b = lane >= b0 && lane < nwords, live = in_a || in_b;
    unsigned mine = 0;
    if (in_a) mine = __float_as_uint(a[lane]);
    if (in_b) mine = reinterpret_cast<const unsigned*>(b)[lane - b0];
    if (live) {
        for (int p = 0; p < world; ++p) {
            uint2* dst = reinterpret_cast<uint2*>(peers.p[p] + data_off) + ((size_t)par * world + rank) * rank_words + lane;
            asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(mine), "r"(seq) : "memory");
        }
    }
    unsigned wv[MAX_WORLD];
#pragma unroll
    for (int r = 0; r < MAX_WORLD; ++r) wv[r] = 0;
    if (live) {
        const uint2* base = reinterpret_cast<const uint2*>(peers.p[rank] + data_off) + (size_t)par * world * rank_words + lane;
        unsigned got = 0, spins = 0;
        const unsigned full = (1u << world) - 1u;
        unsigned long long t0 = 0;
        for (;;) {
            unsigned vx[MAX_WORLD], vy[MAX_WORLD];
#pragma unroll
            for (int r = 0; r < MAX_WORLD; ++r) {
                vx[r] = 0; vy[r] = 0;
                if (r < world && !((got >> r) & 1u))
                    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(vx[r]), "=r"(vy[r]) : "l"(base + (size_t)r * rank_words) : "memory");
            }
#pragma unroll
            for (int r = 0; r < MAX_WORLD; ++r)
                if (r < world && !((got >> r) & 1u) && vy[r] == seq) { got |= 1u << r; wv[r] = vx[r]; }
            if (got == full) break;
            if ((++spins & 0x3ffu) == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > WAIT_TIMEOUT_NS) { atomicAdd(&me->timeouts, 1u); break; }
            }
        }
    }
    __syncwarp();
    float tf = 0.0f;
    double td = 0.0;
#pragma unroll
    for (int r = 0; r < MAX_WORLD; ++r) {
        if (r < world) {
            tf += __uint_as_float(wv[r]);
            const unsigned hi = __shfl_down_sync(0xffffffffu, wv[r], 1);      // word 2i low, 2i + 1 high (b0 is even)
            td += __hiloint2double((int)hi, (int)wv[r]);
        }
    }
    if (in_a) a[lane] = tf;
    if (in_b && ((lane - b0) & 1) == 0) b[(lane - b0) >> 1] = td;
    __syncthreads();
    if (lane == 0) {
        __threadfence();
        *reinterpret_cast<volatile unsigned*>(&me->seq[site]) = seq;
    }
}

}  // namespace b200nn

using namespace b200nn;

extern "C" {

int b200nn_comm_load_nccl(const char* path) { return load_nccl(path); }

int b200nn_comm_nccl_version(void) {
    int v = 0;
    if (g_nccl.GetVersion != nullptr) g_nccl.GetVersion(&v);
    return v;
}

int b200nn_comm_unique_id(void* out128) {
    B200NN_REQUIRE(out128 != nullptr, "comm_unique_id: null buffer");
    if (int rc = load_nccl(nullptr)) return rc;
    ncclUniqueId id;
    B200NN_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return B200NN_OK;
}

int b200nn_comm_create(const void* unique_id128, int rank, int world, b200nn_comm** out) {
    B200NN_REQUIRE(out != nullptr && unique_id128 != nullptr, "comm_create: null argument");
    B200NN_REQUIRE(world >= 1 && world <= MAX_WORLD && rank >= 0 && rank < world, "comm_create: bad rank %d / world %d (max %d)",
                   rank, world, MAX_WORLD);
    if (int rc = load_nccl(nullptr)) return rc;
    b200nn_comm* c = new b200nn_comm();
    memset(c, 0, sizeof(*c));
    c->rank = rank;
    c->world = world;
    B200NN_CUDA(cudaGetDevice(&c->device));
    ncclUniqueId id;
    memcpy(&id, unique_id128, sizeof(id));
    B200NN_NCCL(g_nccl.CommInitRank(&c->nccl, world, id, rank));
    *out = c;
    return B200NN_OK;
}

int b200nn_comm_destroy(b200nn_comm* c) {
    if (c == nullptr) return B200NN_OK;
    for (int p = 0; p < c->world; ++p)
        if (c->peer_ready && p != c->rank && c->peer[p] != nullptr) cudaIpcCloseMemHandle(c->peer[p]);
    if (c->sym_local != nullptr) cudaFree(c->sym_local);
    if (c->nccl != nullptr && g_nccl.CommDestroy != nullptr) g_nccl.CommDestroy(c->nccl);
    delete c;
    return B200NN_OK;
}

size_t b200nn_comm_ipc_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

int b200nn_comm_peer_alloc(b200nn_comm* c, size_t bytes, void* handle_out) {
    B200NN_REQUIRE(c != nullptr && handle_out != nullptr, "comm_peer_alloc: null argument");
    B200NN_REQUIRE(c->sym_local == nullptr, "comm_peer_alloc: already allocated");
    B200NN_REQUIRE(bytes >= sizeof(SymHeader) + (1u << 20), "comm_peer_alloc: need at least %zu bytes", sizeof(SymHeader) + (1u << 20));
    void* p = nullptr;
    B200NN_CUDA(cudaMalloc(&p, bytes));
    B200NN_CUDA(cudaMemset(p, 0, bytes));
    B200NN_CUDA(cudaDeviceSynchronize());
    c->sym_local = static_cast<char*>(p);
    c->sym_bytes = bytes;
    c->bump = (sizeof(SymHeader) + 255) & ~size_t(255);
    cudaIpcMemHandle_t h;
    B200NN_CUDA(cudaIpcGetMemHandle(&h, p));
    memcpy(handle_out, &h, sizeof(h));
    return B200NN_OK;
}

int b200nn_comm_peer_open(b200nn_comm* c, const void* all_handles) {
    B200NN_REQUIRE(c != nullptr && all_handles != nullptr && c->sym_local != nullptr, "comm_peer_open: allocate first");
    const char* hs = static_cast<const char*>(all_handles);
    for (int p = 0; p < c->world; ++p) {
        if (p == c->rank) {
            c->peer[p] = c->sym_local;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, hs + (size_t)p * sizeof(h), sizeof(h));
        void* ptr = nullptr;
        B200NN_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer[p] = static_cast<char*>(ptr);
    }
    c->peer_ready = true;
    return B200NN_OK;
}

int b200nn_comm_has_peer(const b200nn_comm* c) { return (c != nullptr && c->peer_ready) ? 1 : 0; }

// number of peer waits that timed out so far (synchronises the device); non-zero means an exchange produced garbage
int b200nn_comm_timeouts(const b200nn_comm* c) {
    if (c == nullptr || c->sym_local == nullptr) return 0;
    unsigned v = 0;
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpy(&v, c->sym_local + offsetof(SymHeader, timeouts), sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int)v;
}

int b200nn_comm_site_create(b200nn_comm* c, size_t payload_bytes, int* site_out) {
    B200NN_REQUIRE(c != nullptr && site_out != nullptr, "comm_site_create: null argument");
    B200NN_REQUIRE(c->peer_ready, "comm_site_create: peer memory is not mapped");
    B200NN_REQUIRE(c->n_sites < MAX_SITES, "comm_site_create: more than %d exchange sites", MAX_SITES);
    // payloads of the latency path carry an 8-byte word per 4 bytes of payload
    const size_t stride = ((payload_bytes <= LL_MAX_BYTES ? 2 * payload_bytes : payload_bytes) + 255) & ~size_t(255);
    const size_t need = 2 * (size_t)c->world * stride;
    B200NN_REQUIRE(c->bump + need <= c->sym_bytes, "comm_site_create: symmetric buffer exhausted (%zu + %zu > %zu bytes)", c->bump, need,
                   c->sym_bytes);
    const int s = c->n_sites++;
    c->site_off[s] = c->bump;
    c->site_stride[s] = stride;
    c->bump += need;
    *site_out = s;
    return B200NN_OK;
}

// A raw, zero-initialised region of the symmetric buffer (same offset on every rank) plus a site id for its round counter:
// for kernels that exchange over peer memory THEMSELVES (conv_block.cu's peer-fused mode).
int b200nn_comm_region_create(b200nn_comm* c, size_t bytes, int* site_out, size_t* offset_out) {
    B200NN_REQUIRE(c != nullptr && site_out != nullptr && offset_out != nullptr, "comm_region_create: null argument");
    B200NN_REQUIRE(c->peer_ready, "comm_region_create: peer memory is not mapped");
    B200NN_REQUIRE(c->n_sites < MAX_SITES, "comm_region_create: more than %d exchange sites", MAX_SITES);
    const size_t need = (bytes + 255) & ~size_t(255);
    B200NN_REQUIRE(c->bump + need <= c->sym_bytes, "comm_region_create: symmetric buffer exhausted (%zu + %zu > %zu bytes)", c->bump, need,
                   c->sym_bytes);
    const int s = c->n_sites++;
    c->site_off[s] = c->bump;
    c->site_stride[s] = 0;
    *site_out = s;
    *offset_out = c->bump;
    c->bump += need;
    return B200NN_OK;
}

// dtype: 0 = float32, 1 = float64. site >= 0 selects the one-shot peer-memory exchange (the payload must fit the
// site); site < 0 uses NCCL.
int b200nn_comm_allreduce_sum(b200nn_comm* c, int site, void* buf, long long n, int dtype, void* stream) {
    B200NN_REQUIRE(c != nullptr && buf != nullptr && n > 0, "comm_allreduce_sum: bad argument");
    B200NN_REQUIRE(dtype == 0 || dtype == 1, "comm_allreduce_sum: dtype %d", dtype);
    if (c->world == 1) return B200NN_OK;
    const size_t esz = dtype == 0 ? 4 : 8;
    if (site < 0) {
        B200NN_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, dtype == 0 ? ncclFloat32 : ncclFloat64, ncclSum, c->nccl, as_stream(stream)));
        count_launch();
        return B200NN_OK;
    }
    B200NN_REQUIRE(c->peer_ready && site < c->n_sites, "comm_allreduce_sum: unknown site %d", site);
    B200NN_REQUIRE((size_t)n * esz <= c->site_stride[site], "comm_allreduce_sum: payload %lld x %zu exceeds site %d (%zu bytes)", n, esz,
                   site, c->site_stride[site]);
    PeerPtrs pp;
    for (int p = 0; p < MAX_WORLD; ++p) pp.p[p] = p < c->world ? c->peer[p] : nullptr;
    if ((size_t)n * esz <= LL_MAX_BYTES && c->site_stride[site] >= 2 * (size_t)n * esz) {   // latency path
        const int threads = (int)(((size_t)n * esz / 4 + 31) / 32 * 32);
        const size_t rank_words = c->site_stride[site] / 8;
        if (dtype == 0)
            exchange_sum_ll_kernel<float><<<1, threads, 0, as_stream(stream)>>>(pp, c->rank, c->world, site, c->site_off[site], rank_words,
                                                                               static_cast<float*>(buf), (int)n);
        else
            exchange_sum_ll_kernel<double><<<1, threads, 0, as_stream(stream)>>>(pp, c->rank, c->world, site, c->site_off[site], rank_words,
                                                                                static_cast<double*>(buf), (int)n);
        B200NN_LAUNCH_CHECK("exchange_sum_ll");
        return B200NN_OK;
    }
    const int vec_ok = (reinterpret_cast<uintptr_t>(buf) & 15) == 0 ? 1 : 0;
    B200NN_REQUIRE(vec_ok || n <= 4096, "comm_allreduce_sum: large payloads must be 16-byte aligned");
    const long long nvec = vec_ok ? (long long)((size_t)n * esz / 16) : 0;
    long long ctas = (nvec + 255) / 256;      // one 16-byte vector per thread when it fits
    if (ctas > MAX_XCTAS) ctas = MAX_XCTAS;
    if (ctas < 1) ctas = 1;
    const int threads = nvec <= 32 ? 32 : 256;
    if (dtype == 0)
        exchange_sum_kernel<float><<<(int)ctas, threads, 0, as_stream(stream)>>>(pp, c->rank, c->world, site, c->site_off[site],
                                                                                c->site_stride[site], static_cast<float*>(buf), n, vec_ok);
    else
        exchange_sum_kernel<double><<<(int)ctas, threads, 0, as_stream(stream)>>>(pp, c->rank, c->world, site, c->site_off[site],
                                                                                 c->site_stride[site], static_cast<double*>(buf), n, vec_ok);
    B200NN_LAUNCH_CHECK("exchange_sum");
    return B200NN_OK;
}

int b200nn_comm_allreduce_sum2(b200nn_comm* c, int site, float* a, int n_a, double* b, int n_b, void* stream) {
    B200NN_REQUIRE(c != nullptr && a != nullptr && b != nullptr && n_a > 0 && n_b > 0, "comm_allreduce_sum2: bad argument");
    if (c->world == 1) return B200NN_OK;
    const size_t words = (size_t)((n_a + 1) & ~1) + 2 * (size_t)n_b;
    B200NN_REQUIRE(c->peer_ready && site >= 0 && site < c->n_sites && words <= 1024 && c->site_stride[site] >= 8 * words,
                   "comm_allreduce_sum2: %zu words do not fit the latency path of site %d", words, site);
    PeerPtrs pp;
    for (int p = 0; p < MAX_WORLD; ++p) pp.p[p] = p < c->world ? c->peer[p] : nullptr;
    const int threads = (int)((words + 31) / 32 * 32);
    exchange_sum_ll2_kernel<<<1, threads, 0, as_stream(stream)>>>(pp, c->rank, c->world, site, c->site_off[site], c->site_stride[site] / 8,
                                                                a, n_a, b, n_b);
    B200NN_LAUNCH_CHECK("exchange_sum_ll2");
    return B200NN_OK;
}

}  // extern "C"
