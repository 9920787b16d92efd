// peer.cuh -- device helpers for kernels that store into peer GPUs' memory and signal completion (see peer.cu).
#pragma once
#include "common.cuh"

namespace ac {

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// every CTA calls this after its peer stores: the last one to arrive publishes `seq` to all destinations
__device__ __forceinline__ void peer_publish_when_grid_done(const ac_peer_table &t, uint32_t seq, unsigned int *counter) {
    __threadfence_system();                       // this thread's peer stores are ordered before what follows
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(counter, 1u);
        if (prev == gridDim.x * gridDim.y - 1) {
            *counter = 0;                         // next launch on this stream starts from zero
            __threadfence_system();
            for (int p = 0; p < t.world; ++p) st_release_sys(t.flag[p] + t.rank, seq);
        }
    }
}

}  // namespace ac
