// comm.cuh — shared between comm.cu (communicator, exchange kernel) and the kernels that exchange over peer memory
// themselves (conv_block.cu): symmetric-buffer header, peer pointer table, system-scope load / store helpers.
#pragma once
#include <nccl.h>   // types only
#include "common.cuh"

namespace b200nn {

constexpr int MAX_WORLD = 8;
constexpr int MAX_SITES = 96;
constexpr int MAX_XCTAS = 64;

struct SymHeader {                                       // offset 0 of every rank's symmetric buffer
    unsigned int flags[MAX_SITES][2][MAX_WORLD][MAX_XCTAS];
    unsigned int seq[MAX_SITES];                          // completed rounds per site (local)
    unsigned int done[MAX_SITES];                         // CTAs that finished the current round (local)
    unsigned int timeouts;                                // waits that gave up (a peer never arrived): results are invalid
};
constexpr unsigned long long WAIT_TIMEOUT_NS = 10ull * 1000ull * 1000ull * 1000ull;   // a lost peer must not hang the GPU

struct PeerPtrs {
    char* p[MAX_WORLD];
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float ld_inbox(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_inbox(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}


// spin until *flag == seq (system-scope acquire); gives up after WAIT_TIMEOUT_NS and counts the failure
__device__ __forceinline__ void wait_flag(const unsigned* flag, unsigned seq, unsigned* timeouts) {
    unsigned long long t0 = 0;
    unsigned spins = 0;
    while (ld_acquire_sys(flag) != seq) {
        if ((++spins & 0x3ffu) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > WAIT_TIMEOUT_NS) {
                atomicAdd(timeouts, 1u);
                break;
            }
        }
    }
}

// what a kernel needs to exchange through a raw region of the symmetric buffer (b200nn_comm_region_create)
struct PeerRegion {
    PeerPtrs peers;
    int rank, world, site;
    size_t off;          // byte offset of the region in every rank's symmetric buffer
    unsigned stagger_ns; // start offset between CTAs that share an SM (peer-fused kernels)
    int poll_all;        // 1: poll every rank's word in the same sweep (loads in flight together); 0: rank by rank
};

}  // namespace b200nn

struct b200nn_comm {
    int rank, world, device;
    ncclComm_t nccl;
    char* sym_local;
    size_t sym_bytes;
    char* peer[b200nn::MAX_WORLD];
    bool peer_ready;
    size_t bump;
    int n_sites;
    size_t site_off[b200nn::MAX_SITES], site_stride[b200nn::MAX_SITES];
};

