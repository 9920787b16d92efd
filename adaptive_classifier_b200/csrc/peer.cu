// peer.cu -- the exchange steps of the row-sharded prototype search written straight into peer memory over NVLink.
//
// with the prototype rows sharded over G GPUs, a step needs (1) every rank's unit embeddings on
// every rank and (2) every shard's per-query candidates (d, global id) on the rank that owns the query.  Round 1 did both
// with NCCL (all-gather, two all-to-alls: ~0.9 ms of a 16.5 ms step at G = 8, almost all of it launch + synchronisation,
// the payloads are 1.5 MB and 245 KB).  Here the producer kernels store directly into the consumers' buffers:
//
//   ac_peer_scatter     one kernel copies a local block into slot `rank` of EVERY peer's buffer (embeddings), or block g
//                       into slot `rank` of peer g's buffer (candidates), with plain st.global on NVLink-mapped peer
//                       pointers (torch symmetric memory provides the mapping: plumbing, like torch.distributed);
//                       the last CTA to finish publishes a sequence number in flag[rank] of every destination
//                       (__threadfence_system + st.release.sys).
//   ac_peer_wait        spins (ld.acquire.sys) until the local flags of all sources reached the sequence number.
//   cls_normalize_scatter (encoder.cu epilogue, through ac_encoder_forward_cls_scatter): the last kernel of the encoder
//                       writes the unit CLS rows to all peers itself, so step (1) costs no kernel of its own.
//
// Buffers are double-buffered by sequence parity by the caller (adaptive_classifier_b200/parallel.py explains why two
// buffers suffice).  Status: written after the round-1 GPU budget was spent; compiles for sm_100a, NOT yet run on hardware
// (needs >= 2 GPUs); the NCCL path stays the default.
#include "peer.cuh"

namespace ac {

// mode 0: the same `n16` 16-byte units of src go to every peer (slot offset dst_off);  mode 1: block p of src (blocks are
// block_stride16 units apart) goes to peer p
__global__ void peer_scatter_kernel(const uint4 *__restrict__ src, size_t n16, ac_peer_table t, size_t dst_off, int mode,
                                    size_t block_stride16, uint32_t seq, unsigned int *counter) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (int p = 0; p < t.world; ++p) {
        const uint4 *s = src + (mode ? static_cast<size_t>(p) * block_stride16 : 0);
        uint4 *d = reinterpret_cast<uint4 *>(static_cast<uint8_t *>(t.buf[p]) + dst_off);
        for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) d[i] = s[i];
    }
    peer_publish_when_grid_done(t, seq, counter);
}

__global__ void peer_wait_kernel(const uint32_t *flags, int n, uint32_t seq) {
    const int i = threadIdx.x;
    if (i >= n) return;
    unsigned long long spins = 0;
    // sequence numbers only grow (wrap-safe comparison)
    while (static_cast<int32_t>(ld_acquire_sys(flags + i) - seq) < 0) {
        __nanosleep(200);
        if (++spins > (1ull << 24)) {             // ~ seconds: a peer died or the protocol is broken -> fail loudly
            printf("ac: peer wait watchdog fired (flag %d: have %u, want %u)\n", i, ld_acquire_sys(flags + i), seq);
            __trap();
        }
    }
}

}  // namespace ac

using namespace ac;

static int check_table(const ac_peer_table *t, const char *who) {
    AC_REQUIRE(t && t->world >= 1 && t->world <= AC_MAX_PEERS && t->rank >= 0 && t->rank < t->world, "%s: bad peer table", who);
    for (int p = 0; p < t->world; ++p) AC_REQUIRE(t->buf[p] && t->flag[p], "%s: null peer pointer %d", who, p);
    return AC_OK;
}

extern "C" int ac_peer_scatter(const void *src, size_t bytes_per_dst, int blocks_mode, const ac_peer_table *t,
                               size_t dst_offset_bytes, uint32_t seq, uint32_t *counter, ac_stream_t stream) {
    int rc = check_table(t, "ac_peer_scatter");
    if (rc) return rc;
    AC_REQUIRE(src && counter, "ac_peer_scatter: null argument");
    AC_REQUIRE(bytes_per_dst % 16 == 0 && dst_offset_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0,
               "ac_peer_scatter: 16-byte granularity");
    if (bytes_per_dst == 0) return AC_OK;
    const size_t n16 = bytes_per_dst / 16;
    int blocks = static_cast<int>((n16 + 255) / 256);
    if (blocks > 2 * sm_count()) blocks = 2 * sm_count();
    peer_scatter_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4 *>(src), n16, *t, dst_offset_bytes,
                                                                             blocks_mode ? 1 : 0, n16, seq, counter);
    AC_LAUNCH_CHECK();
    return AC_OK;
}

extern "C" int ac_peer_wait(const uint32_t *flags_local, int n_flags, uint32_t seq, ac_stream_t stream) {
    AC_REQUIRE(flags_local && n_flags >= 1 && n_flags <= 1024, "ac_peer_wait: bad arguments");
    peer_wait_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(flags_local, n_flags, seq);
    AC_LAUNCH_CHECK();
    return AC_OK;
}
