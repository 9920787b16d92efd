// tc_emul.h -- functional model of the Blackwell pieces the kernels use (mbarrier, TMA 2-D tile loads with the 128-byte
// swizzle, tcgen05.mma kind::f16 with K-major SWIZZLE_128B shared-memory descriptors, TMEM, tcgen05.ld/commit), with the
// same names as the PTX wrappers of csrc/common.cuh.  TEST INFRASTRUCTURE ONLY (tests/cpu_shim).
//
// Purpose: kernels written without GPU access (attention_pipe_kernel) can be RUN on the CPU next to kernels that were
// verified on a B200 (attention_kernel).  The model is first checked against a verified kernel -- it must reproduce a plain
// reference through that kernel's TMA / descriptor / TMEM addressing -- and then used to demand that the new kernel gives the
// same bits.  Everything asynchronous completes at issue (TMA bytes land immediately, MMAs execute immediately, commits
// arrive immediately); ordering bugs are the business of tests/test_sim_attention_pipe.py, this model checks data paths:
// buffer offsets, descriptors, swizzle, TMEM columns, barrier counts and parities (a wrong parity deadlocks here too).
#pragma once
#include "cuda_shim.h"

typedef struct CUtensorMap_st {
    const void *ptr;
    int elem_bytes;
    uint64_t rows, cols, stride_bytes;
    uint32_t box_rows, box_cols;
} CUtensorMap;
#define __grid_constant__

static inline void __trap() { printf("__trap() reached\n"); abort(); }
static inline uint32_t shim_ballot(bool pred) {
    uint32_t mine = pred ? 1u : 0u, out = 0;
    for (int l = 0; l < 32; ++l) out |= (shim_shfl(mine, l) & 1u) << l;
    return out;
}
#define __ballot_sync(mask, pred) shim_ballot(pred)

namespace shim {
// per-block dynamic shared memory (1024-byte aligned so that address bits [7:9] used by the swizzle are those of the offset)
// and TMEM live in the Block (cuda_shim.h); all blocks of one run form the cluster
inline uint8_t *dyn_smem() { return g_cur->blk->dyn_smem; }
inline Block *cluster_block(int rank) { return (*g_cur->blk->cluster)[rank]; }
}  // namespace shim
#define __cluster_dims__(...)

static inline void shim_st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    uint32_t v[4] = {a, b, c, d};
    memcpy(shim::dyn_smem() + addr, v, 16);
}

namespace ac {

static inline uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(static_cast<const uint8_t *>(p) - shim::dyn_smem()); }

// ---- mbarrier: the 64-bit word holds {count:16, pending:16, phase:1, tx bytes:31}
struct MbarBits { uint64_t count : 16, pending : 16, phase : 1; int64_t tx : 31; };   // tx may go transiently negative (bytes of a peer land before expect_tx)
static_assert(sizeof(MbarBits) == 8, "mbarrier model must fit the 64-bit word");
static inline MbarBits *mb(uint64_t *bar) { return reinterpret_cast<MbarBits *>(bar); }
static inline void mbar_flip_if_complete(uint64_t *bar) {
    MbarBits *b = mb(bar);
    if (b->pending == 0 && b->tx == 0) {
        b->phase ^= 1;
        b->pending = b->count;
        ++shim::g_progress;
    }
}
static inline void mbar_init(uint64_t *bar, uint32_t count) { MbarBits z{}; z.count = count; z.pending = count; *mb(bar) = z; }
static inline void fence_mbar_init() {}
static inline void mbar_arrive(uint64_t *bar) {
    if (mb(bar)->pending == 0) { printf("mbarrier: more arrivals than its count\n"); abort(); }
    mb(bar)->pending -= 1;
    mbar_flip_if_complete(bar);
}
static inline void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    mb(bar)->tx += static_cast<int64_t>(bytes);
    mbar_arrive(bar);
}
static inline void mbar_complete_tx(uint64_t *bar, uint32_t bytes) {
    mb(bar)->tx -= static_cast<int64_t>(bytes);
    mbar_flip_if_complete(bar);
}
static inline bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    if (mb(bar)->phase != parity) return true;       // the phase with this parity has completed
    shim::yield();                                   // let the other threads run; a pass without any progress is a deadlock
    return false;
}

// ---- TMA: 2-D tile, 128-byte swizzle (16-byte chunk index XOR (row & 7)), out-of-range elements are zero
static inline void tma_prefetch_desc(const CUtensorMap *) {}
static inline void tma_load_2d(void *smem_dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1) {
    uint8_t *dst = static_cast<uint8_t *>(smem_dst);
    if ((smem_u32(dst) & 1023u) != 0) { printf("TMA destination is not 1024-byte aligned\n"); abort(); }
    const int eb = m->elem_bytes;
    for (uint32_t r = 0; r < m->box_rows; ++r) {
        const int64_t grow = static_cast<int64_t>(c1) + r;
        for (uint32_t ch = 0; ch < 8; ++ch) {
            uint8_t tmp[16];
            for (int e = 0; e < 16 / eb; ++e) {
                const int64_t gcol = static_cast<int64_t>(c0) + ch * (16 / eb) + e;
                if (grow >= 0 && static_cast<uint64_t>(grow) < m->rows && gcol >= 0 && static_cast<uint64_t>(gcol) < m->cols)
                    memcpy(tmp + e * eb, static_cast<const uint8_t *>(m->ptr) + grow * m->stride_bytes + gcol * eb, eb);
                else
                    memset(tmp + e * eb, 0, eb);
            }
            memcpy(dst + r * 128 + ((ch ^ (r & 7)) << 4), tmp, 16);
        }
    }
    mbar_complete_tx(bar, m->box_rows * 128);
}
static inline void fence_proxy_async_smem() {}

// ---- TMEM
static inline void tmem_alloc(uint32_t *smem_dst, uint32_t) { *smem_dst = 0; }
static inline void tmem_relinquish() {}
static inline void tmem_dealloc(uint32_t, uint32_t) {}
static inline void tc_fence_before() {}
static inline void tc_fence_after() {}
static inline void tc_commit(uint64_t *bar) { mbar_arrive(bar); }       // every MMA issued so far has already executed
static inline void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    const int lane = static_cast<int>(taddr >> 16) + shim::g_cur->lane, col = static_cast<int>(taddr & 0xffffu);
    if ((static_cast<int>(taddr >> 16) & 31) != 0 || (static_cast<int>(taddr >> 16) >> 5) != ((threadIdx.x >> 5) & 3)) {
        printf("tcgen05.ld: warp %u reads TMEM lane base %u (a warp may only touch lanes 32*(warp %% 4)..)\n", threadIdx.x >> 5, taddr >> 16);
        abort();
    }
    for (int j = 0; j < 32; ++j) memcpy(&r[j], &shim::g_cur->blk->tmem[lane][col + j], 4);
}
static inline uint32_t tmem_ld_32x1(uint32_t taddr) {
    uint32_t v;
    memcpy(&v, &shim::g_cur->blk->tmem[static_cast<int>(taddr >> 16) + shim::g_cur->lane][taddr & 0xffffu], 4);
    return v;
}
static inline void tmem_ld_wait() {}

// ---- tcgen05.mma kind::f16, A and B K-major in the SWIZZLE_128B layout, one instruction = K 16
static inline float umma_operand(const uint8_t *smem, uint32_t start_byte, int row, int k /*0..15*/) {
    // address = start + row * 128 + k * 2; the swizzle XORs address bits [4:6] with bits [7:9]
    uint32_t addr = start_byte + static_cast<uint32_t>(row) * 128u + static_cast<uint32_t>(k) * 2u;
    addr ^= ((addr >> 7) & 7u) << 4;
    __half h;
    memcpy(&h, smem + addr, 2);
    return __half2float(h);
}
static inline void umma_tf32(uint32_t, uint64_t, uint64_t, uint32_t, uint32_t) { printf("umma model: kind::tf32 is not modelled\n"); abort(); }
static inline void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    const int M = static_cast<int>((idesc >> 24) & 31u) << 4, N = static_cast<int>((idesc >> 17) & 63u) << 3;
    if (M != 128 || ((idesc >> 7) & 7u) != 0 || ((a_desc >> 61) & 7u) != 2 || ((b_desc >> 61) & 7u) != 2 || (d_tmem >> 16) != 0) {
        printf("umma model: unsupported instruction (M %d, fmt %u, swizzle %u/%u)\n", M, (idesc >> 7) & 7u, unsigned((a_desc >> 61) & 7u), unsigned((b_desc >> 61) & 7u));
        abort();
    }
    const uint32_t a0 = static_cast<uint32_t>(a_desc & 0x3FFFu) << 4, b0 = static_cast<uint32_t>(b_desc & 0x3FFFu) << 4;
    const int col0 = static_cast<int>(d_tmem & 0xffffu);
    const uint8_t *smem = shim::dyn_smem();
    float (*tm)[512] = shim::g_cur->blk->tmem;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = accumulate ? tm[m][col0 + n] : 0.f;
            for (int k = 0; k < 16; ++k) s = fmaf(umma_operand(smem, a0, m, k), umma_operand(smem, b0, n, k), s);
            tm[m][col0 + n] = s;
        }
}

// ---------------------------------------------------------------- CTA pair (cluster of 2, cta_group::2)
// shared::cluster addresses: the CTA rank + 1 in bits [24..], the shared::cta offset below
static inline uint32_t cluster_ctarank() { return static_cast<uint32_t>(shim::g_cur->blk->cluster_rank); }
static inline uint32_t mapa_shared(uint32_t addr, uint32_t rank) { return addr | ((rank + 1u) << 24); }
static inline uint64_t *cluster_bar(uint32_t caddr) {
    const uint32_t r = caddr >> 24;
    shim::Block *b = r ? shim::cluster_block(static_cast<int>(r) - 1) : shim::g_cur->blk;
    return reinterpret_cast<uint64_t *>(b->dyn_smem + (caddr & 0xFFFFFFu));
}
static inline void cluster_sync_all() { shim::barrier_wait(shim::g_grid_bar); }        // one run = one cluster
static inline void mbar_arrive_cluster(uint32_t caddr) { mbar_arrive(cluster_bar(caddr)); }
static inline bool mbar_try_wait_cluster(uint64_t *bar, uint32_t parity) { return mbar_try_wait(bar, parity); }
static inline void tma_load_2d_pair(void *smem_dst, const CUtensorMap *m, uint32_t bar_caddr, int c0, int c1) {
    uint64_t scratch;                                   // tma_load_2d signals a barrier: give it a private one, then forward
    mbar_init(&scratch, 1);
    mb(&scratch)->tx = static_cast<int64_t>(m->box_rows) * 128;
    tma_load_2d(smem_dst, m, &scratch, c0, c1);
    mbar_complete_tx(cluster_bar(bar_caddr), m->box_rows * 128);
}
static inline void tmem_alloc_pair(uint32_t *smem_dst, uint32_t) { *smem_dst = 0; }
static inline void tmem_relinquish_pair() {}
static inline void tmem_dealloc_pair(uint32_t, uint32_t) {}
static inline void tc_commit_pair(uint64_t *bar, uint16_t cta_mask) {
    const uint32_t off = smem_u32(bar);
    for (int r = 0; r < 2; ++r)
        if (cta_mask & (1u << r)) mbar_arrive(reinterpret_cast<uint64_t *>(shim::cluster_block(r)->dyn_smem + off));
}
// one instruction of the pair: D (256 x N) = A (256 x 16) B^T (N x 16).  CTA r holds rows 128r.. of A and of D (its TMEM)
// and rows (N/2) r .. of B, each at the SAME shared-memory offsets (this is what ran bit-identically on the B200)
static inline void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    const int M = static_cast<int>((idesc >> 24) & 31u) << 4, N = static_cast<int>((idesc >> 17) & 63u) << 3;
    if (M != 256 || ((idesc >> 7) & 7u) != 0 || cluster_ctarank() != 0) { printf("umma pair model: M %d or not the leader\n", M); abort(); }
    const uint32_t a0 = static_cast<uint32_t>(a_desc & 0x3FFFu) << 4, b0 = static_cast<uint32_t>(b_desc & 0x3FFFu) << 4;
    const int col0 = static_cast<int>(d_tmem & 0xffffu);
    for (int r = 0; r < 2; ++r) {
        shim::Block *cta = shim::cluster_block(r);
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                const shim::Block *bsrc = shim::cluster_block(n / (N / 2));
                float s = accumulate ? cta->tmem[m][col0 + n] : 0.f;
                for (int k = 0; k < 16; ++k) s = fmaf(umma_operand(cta->dyn_smem, a0, m, k), umma_operand(bsrc->dyn_smem, b0, n % (N / 2), k), s);
                cta->tmem[m][col0 + n] = s;
            }
    }
}
static inline void umma_tf32_pair(uint32_t, uint64_t, uint64_t, uint32_t, uint32_t) { printf("umma model: kind::tf32 is not modelled\n"); abort(); }

}  // namespace ac
